"""Host-side mirror of the reference's interface for the correction path.

Same names and argument meaning as /root/reference/TranscriptClean/TranscriptClean.py:
  prep_refs(options, transcripts, sam_header)            TC.py:203-255
  batch_correct(sam_transcripts, options, refs, outfiles, buffer_size)   TC.py:350-403
  correct_transcript(transcript_line, options, refs)     TC.py:406-460
  dryRun(sam, options, outfiles)                         TC.py:1615-1678
`options` is any object with the reference's option attributes (correct* are the strings
"true"/"false" as in the reference, booleans are accepted too); `refs` carries the device
context instead of pyfaidx / pyranges / dict objects.
"""
import warnings

from . import engine as E
from . import render, sam
from .genome import GenomeImage
from .tables import SjTables, VariantTables, build_sj_tables, build_variant_tables


class Struct(dict):
    """dstruct.Struct of the reference (dstruct.py:1-11)."""

    def __init__(self, **kw):
        dict.__init__(self, kw)
        self.__dict__ = self


def _truthy(v):
    return v is True or (isinstance(v, str) and v.lower() == "true")


def _opt(options, name, default=None):
    return getattr(options, name, default) if not isinstance(options, dict) else options.get(name, default)


def make_refs(genome, sj_tables=None, var_tables=None, device=0, engine=None):
    """Builds the `refs` Struct around an Engine.  `genome` is a GenomeImage."""
    refs = Struct()
    refs.genome = genome
    refs.engine = engine if engine is not None else E.Engine(device)
    refs.engine.load_genome(genome)
    refs.sj = sj_tables if sj_tables is not None else SjTables()
    refs.var = var_tables if var_tables is not None else VariantTables()
    refs.engine.load_junctions(refs.sj)
    refs.engine.load_variants(refs.var)
    refs.sjAnnot = refs.sj          # len(refs.sjAnnot) > 0 keeps its meaning (TC.py:443)
    return refs


def build_tables(options, chroms, genome):
    """The splice-junction and variant tables of prep_refs (TC.py:224-253) as host arrays: built ONCE, whatever the
    number of GPUs (`make_refs` uploads the same arrays to every engine)."""
    sj = var = None
    sj_file, vcf = _opt(options, "sjAnnotFile"), _opt(options, "variantFile")
    if not _truthy(_opt(options, "correctSJs", "true")):
        print("SJ correction turned off in options. Skipping reference SJ parsing.")
    elif sj_file is not None:
        print("Processing annotated splice junctions ...")
        sj = build_sj_tables(sj_file, chroms, genome.index)
    else:
        print("No splice annotation provided. Will skip splice junction correction.")
    if vcf is not None:
        print("Processing variant file .................")
        var = build_variant_tables(vcf, int(_opt(options, "maxLenIndel", 5)), genome.index, read_chroms=chroms)
    else:
        print("No variant file provided. Transcript correction will not be variant-aware.")
    return sj, var


def prep_refs(options, transcripts, sam_header=None, device=0, engine=None, genome=None):
    """TC.py:203-255: genome, splice annotation restricted to the chromosomes of `transcripts`,
    variant catalogues.  No temp files are written."""
    if genome is None:
        print("Reading genome ..............................")
        try:
            lib = engine._lib if engine is not None else E.load_library()
        except Exception:
            lib = None
        genome = GenomeImage.from_fasta_cached(_opt(options, "refGenome"), lib=lib)      # packed by the library, cached next to the FASTA
    chroms = set(t.split("\t", 3)[2] for t in transcripts)
    sj, var = build_tables(options, chroms, genome)
    return make_refs(genome, sj, var, device=device, engine=engine)


def _set_engine_options(options, refs):
    refs.engine.set_options(
        maxLenIndel=int(_opt(options, "maxLenIndel", 5)), maxSJOffset=int(_opt(options, "maxSJOffset", 5)),
        correctMismatches=_truthy(_opt(options, "correctMismatches", "true")),
        correctIndels=_truthy(_opt(options, "correctIndels", "true")),
        correctSJs=_truthy(_opt(options, "correctSJs", "true")),
        sj_ignore_strand=bool(_opt(options, "sjIgnoreStrand", False)),
        sj_tie_right=bool(_opt(options, "sjTieRight", False)))


def correct_batch_text(sam_transcripts, options, refs):
    """Corrects a list of SAM lines on the GPU and returns the four output texts."""
    _set_engine_options(options, refs)
    pb = sam.parse_lines(sam_transcripts, refs.genome)
    res = refs.engine.correct(pb.cols)
    return render.render_batch(pb, res, bool(_opt(options, "primaryOnly", False)), bool(_opt(options, "canonOnly", False)),
                               warn=print)


def correct_sam_text(text, options, refs, dry=False, consume=None):
    """Fast path for whole SAM texts (bytes): parsing and rendering in the library's C++ host code,
    correction on the GPU.  Returns (header str, {"sam","fa","log","te"} as bytes).  With
    `consume(header, views)` the four texts are handed over as memoryviews of the library's own
    buffers (no copies; only valid during the call) and (header, None) is returned."""
    _set_engine_options(options, refs)
    st = refs.engine.parse_sam(text, refs.genome)
    try:
        raw = refs.engine.raw_dryrun(st.batch) if dry else refs.engine.raw_correct(st.batch)
        out = st.render(raw, bool(_opt(options, "primaryOnly", False)), bool(_opt(options, "canonOnly", False)), dry,
                        copy=consume is None)
        if consume is not None:
            consume(st.header, out)
            for v in out.values():
                v.release()
            out = None
        return st.header, out
    finally:
        st.close()


def correct_sam_stream(blocks, options, refs_list, sink, dry=False, n_threads=None):
    """Pipelined text path: `blocks` yields SAM text (bytes, whole lines) in input order; every entry of `refs_list` is a
    worker with its own Engine (several per GPU are fine - they are what overlaps parsing, the GPU call and rendering of
    successive blocks, the way the reference overlaps its worker processes, TC.py:48-62).  `sink(header, views)` gets the
    four texts of every block in input order, each as a list of pieces (memoryviews of the formatter threads' own buffers:
    nothing is copied together; valid during the call).
    Host threads for the C++ parser / renderer are split between the workers."""
    import os
    import threading
    n_workers = len(refs_list)
    per = max(1, int(n_threads or min(32, os.cpu_count() or 1)) // n_workers)
    it = enumerate(blocks)
    lock = threading.Lock()
    cond = threading.Condition()
    ready, state = {}, {"next": 0, "err": None, "end": None, "live": n_workers}

    def take():
        with lock:
            try:
                return next(it)
            except StopIteration:
                return None

    timing = os.environ.get("TC_STREAM_TIMING") == "1"           # per-worker seconds by stage, to stderr
    import time

    def work(refs):
        tm = {"parse": 0.0, "abi": 0.0, "render": 0.0, "wait_writer": 0.0, "close": 0.0, "blocks": 0}
        try:
            _set_engine_options(options, refs)
            while state["err"] is None:
                item = take()
                if item is None:
                    return
                idx, text = item
                t0 = time.perf_counter()
                st = refs.engine.parse_sam(text, refs.genome, per)
                t1 = time.perf_counter()
                try:
                    raw = refs.engine.raw_dryrun(st.batch) if dry else refs.engine.raw_correct(st.batch)
                    t2 = time.perf_counter()
                    out = st.render_parts(raw, bool(_opt(options, "primaryOnly", False)), bool(_opt(options, "canonOnly", False)), dry)
                    t3 = time.perf_counter()
                    done = threading.Event()
                    with cond:
                        ready[idx] = (st.header, out, done)
                        cond.notify_all()
                    done.wait()                                    # the writer is through with this block's buffers
                    t4 = time.perf_counter()
                    for vs in out.values():
                        for v in vs:
                            v.release()
                finally:
                    st.close()
                t5 = time.perf_counter()
                for k, v in (("parse", t1 - t0), ("abi", t2 - t1), ("render", t3 - t2), ("wait_writer", t4 - t3), ("close", t5 - t4)):
                    tm[k] += v
                tm["blocks"] += 1
        except BaseException as e:                                 # surfaced by the writer loop
            with cond:
                state["err"] = state["err"] or e
                cond.notify_all()
        finally:
            with cond:
                state["live"] -= 1
                cond.notify_all()                                  # (the writer also waits for the last worker to leave)
            if timing:
                import sys
                print("correct_sam_stream worker (%d threads): %s" % (per, {k: round(v, 4) for k, v in tm.items()}), file=sys.stderr)

    threads = [threading.Thread(target=work, args=(r,), daemon=True) for r in refs_list]
    for t in threads:
        t.start()
    try:
        while True:
            with cond:
                while state["next"] not in ready and state["err"] is None and state["live"] > 0:
                    cond.wait(0.05)
                if state["err"] is not None:
                    raise state["err"]
                if state["next"] not in ready:
                    break                                          # every worker has finished and nothing is pending
                header, out, done = ready.pop(state["next"])
            try:
                sink(header, out)
            finally:
                done.set()
            state["next"] += 1
    finally:
        with cond:
            state["err"] = state["err"] or RuntimeError("stream aborted")
            for _, _, d in ready.values():
                d.set()
        for t in threads:
            t.join(timeout=5)
    return state["next"]


def _seam_write(sam_transcripts, options, refs, files, dry=False):
    """Corrects a list of SAM lines - what the reference's seam functions are handed - and writes the texts to `files`
    ({"sam","fa","log","te"} -> open text file; missing keys are skipped).  Lists of any size go through the library's C++
    parser / renderer (the lines are joined into one text: the fast path of the command line) and its buffers are written without
    a copy; a handful of lines, or TC_PY_TEXT=1, go through the Python twins sam.py / render.py.  Same texts either way
    (tests/test_host_fast_path.py)."""
    import os
    lines = sam_transcripts if isinstance(sam_transcripts, list) else list(sam_transcripts)
    if len(lines) < 16 or os.environ.get("TC_PY_TEXT") == "1" or not hasattr(refs.engine, "parse_sam"):
        out = dry_run_text(lines, refs) if dry else correct_batch_text(lines, options, refs)
        for k, f in files.items():
            f.write(out[k])
        return
    text = "\n".join(lines) + "\n"                       # (lines that end in a newline leave blank lines, which the parser skips)

    def sink(_header, views):
        for k, f in files.items():
            _write_text(f, views[k])
    correct_sam_text(text.encode(), options if options is not None else Struct(), refs, dry=dry, consume=sink)


def _write_text(f, data):
    """Writes a text (str, or bytes in the file's own encoding) to an open text file.  Bytes go straight to the binary file
    under a text-mode file object (after a flush of what the caller wrote before) - decoding several GB to str only for the
    file object to encode them again is most of the seam's time."""
    import os
    if isinstance(data, str):
        f.write(data)
        return
    raw = getattr(f, "buffer", None)
    enc = str(getattr(f, "encoding", "") or "").lower().replace("-", "").replace("_", "")
    if raw is not None and os.linesep == "\n" and enc in ("utf8", "ascii", "latin1", "iso88591", "ansix3.41968", "usascii", "646"):
        f.flush()
        raw.write(data)
    else:
        f.write(bytes(data).decode())


def batch_correct(sam_transcripts, options, refs, outfiles, buffer_size=100):
    """TC.py:350-403.  `buffer_size` only staggers disk writes in the reference; the whole list
    is one device batch here."""
    print("Correcting transcripts...")
    _seam_write(sam_transcripts, options, refs, {"sam": outfiles.sam, "fa": outfiles.fasta, "log": outfiles.log, "te": outfiles.TElog})


def correct_transcript(transcript_line, options, refs):
    """TC.py:406-460: (Transcript-like or None, logInfo, TE_entries) for one SAM line."""
    _set_engine_options(options, refs)
    pb = sam.parse_lines([transcript_line], refs.genome)
    res = refs.engine.correct(pb.cols)
    render.check_errors(pb, res)
    st = int(res.status[0])
    cls = st & E.ST_CLASS_MASK
    log = Struct()
    log.TranscriptID = transcript_line.split("\t", 1)[0]
    log.Mapping = render.CLASS_NAME[cls]
    for k, name in enumerate(("corrected_deletions", "uncorrected_deletions", "variant_deletions",
                              "corrected_insertions", "uncorrected_insertions", "variant_insertions",
                              "corrected_mismatches", "uncorrected_mismatches", "corrected_NC_SJs",
                              "uncorrected_NC_SJs")):
        c = int(res.counters[0][k])
        log[name] = "NA" if c < 0 else c
    if cls != 0:
        return None, log, ""
    if st & E.ST_FALLBACK:
        warnings.warn("Problem encountered while correcting transcript with ID %s. Will output original version."
                      % log.TranscriptID)
    view = render.ReadView(pb, res, 0)
    return view, log, render.te_rows(view.QNAME, view.CHROM, res.te[res.te_off[0]:res.te_off[1]])


def dry_run_text(sam_transcripts, refs):
    pb = sam.parse_lines(sam_transcripts, refs.genome)
    res = refs.engine.dryrun(pb.cols)
    return render.render_dry(pb, res)


def dryRun(sam_lines, options, outfiles, refs=None, device=0):
    """TC.py:1615-1678."""
    print("Dry run mode: Identifying indels and mismatches.........")
    if refs is None:
        refs = make_refs(GenomeImage.from_fasta_cached(_opt(options, "refGenome"), lib=E.load_library()), device=device)
    _seam_write(sam_lines, options, refs, {"log": outfiles.log, "te": outfiles.TElog}, dry=True)

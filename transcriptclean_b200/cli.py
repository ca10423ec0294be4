"""`transcriptclean` command line on top of the B200 engine.

Option surface, defaults, prefix abbreviations (optparse) and output files are those of the
reference (getOptions / cleanup_options / combine_outputs, TranscriptClean.py:72-200, 604-663).
`--threads` keeps its place in the interface; here it is the number of GPUs to shard over
(capped by the GPUs present), the analogue of the reference's worker processes.
"""
import os
import sys
from optparse import OptionParser

from . import api
from .genome import GenomeImage

LOG_HEADER = "\t".join(["TranscriptID", "Mapping", "corrected_deletions", "uncorrected_deletions", "variant_deletions",
                        "corrected_insertions", "uncorrected_insertions", "variant_insertions", "corrected_mismatches",
                        "uncorrected_mismatches", "corrected_NC_SJs", "uncorrected_NC_SJs"]) + "\n"
TE_HEADER = "\t".join(["TranscriptID", "Position", "ErrorType", "Size", "Corrected", "ReasonNotCorrected"]) + "\n"


def get_options(argv=None):
    p = OptionParser()
    p.add_option("--sam", "-s", dest="sam", metavar="FILE", type="string", default="",
                 help="Input SAM file containing transcripts to correct. Must contain a header.")
    p.add_option("--genome", "-g", dest="refGenome", metavar="FILE", type="string", default="",
                 help="Reference genome fasta file.")
    p.add_option("--threads", "-t", dest="n_threads", type="int", default=1,
                 help="Number of GPUs to shard the reads over (the reference's worker count).")
    p.add_option("--spliceJns", "-j", dest="sjAnnotFile", metavar="FILE", type="string", default=None,
                 help="STAR-format splice junction file.")
    p.add_option("--variants", "-v", dest="variantFile", metavar="FILE", type="string", default=None,
                 help="VCF of variants to avoid correcting away (optional).")
    p.add_option("--maxLenIndel", dest="maxLenIndel", type="int", default=5, help="Maximum size indel to correct (Default: 5 bp)")
    p.add_option("--maxSJOffset", dest="maxSJOffset", type="int", default=5,
                 help="Maximum distance from annotated splice junction to correct (Default: 5 bp)")
    p.add_option("--outprefix", "-o", dest="outprefix", metavar="FILE", type="string", default="TC",
                 help="Output file prefix. '_clean' plus a file extension will be added to the end.")
    p.add_option("--correctMismatches", "-m", dest="correctMismatches", type="string", default="true")
    p.add_option("--correctIndels", "-i", dest="correctIndels", type="string", default="true")
    p.add_option("--correctSJs", dest="correctSJs", type="string", default="true")
    p.add_option("--dryRun", dest="dryRun", action="store_true", default=False)
    p.add_option("--primaryOnly", dest="primaryOnly", action="store_true", default=False)
    p.add_option("--canonOnly", dest="canonOnly", action="store_true", default=False)
    p.add_option("--tmpDir", dest="tmp_path", default=None)
    p.add_option("--bufferSize", dest="buffer_size", type="int", default=100)
    p.add_option("--deleteTmp", dest="delete_tmp", action="store_true", default=False)
    options, _ = p.parse_args(argv)
    return cleanup_options(options)


def cleanup_options(options):
    """TC.py:158-200: case-fold and validate the true/false strings, derive TC_tmp and the prefix."""
    for name, flag in (("correctMismatches", "--correctMismatches/-m"), ("correctIndels", "--correctIndels/-i"),
                       ("correctSJs", "--correctSJs")):
        v = getattr(options, name).lower()
        if v not in ("true", "false"):
            raise RuntimeError("Invalid choice for %s option. Valid choices are 'true' and 'false'" % flag)
        setattr(options, name, v)
    if options.tmp_path is not None:
        parts = options.tmp_path.split("/")
        options.tmp_dir = "/".join((parts if os.path.isdir(options.tmp_path) else parts[0:-1]) + ["TC_tmp/"])
    else:
        options.tmp_dir = "/".join(options.outprefix.split("/")[0:-1] + ["TC_tmp/"])
    if os.path.isdir(options.outprefix):
        options.outprefix = "/".join(options.outprefix.split("/") + ["TC"])
    return options


def split_sam(path):
    """TC.py:519-541: header lines, chromosome set, alignment lines (stripped)."""
    header, chroms, lines = [], set(), []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if line.startswith("@"):
                header.append(line)
            elif line:
                lines.append(line)
                chroms.add(line.split("\t")[2])
    return header, chroms, lines


def vcf_chrom_set(vcf_file, lib=None):
    """The set validate_chroms collects from the VCF (TC.py:497-505): first column of every line that is not blank and does not
    start with '#'.  With `lib`, one pass in C++ (tc_vcf_chroms) instead of a Python loop over millions of records."""
    import gzip
    opener = gzip.open if str(vcf_file).endswith(".gz") else open
    if lib is None:
        with opener(vcf_file, "rt") as f:
            return set(l.split("\t", 1)[0] for l in f if l.strip() and not l.startswith("#"))
    import ctypes as C
    with opener(vcf_file, "rb") as f:
        text = f.read()
    n = C.c_int64()
    p = lib.tc_vcf_chroms(text, len(text), C.byref(n))
    if not p:
        raise MemoryError("tc_vcf_chroms")
    try:
        return set(x.decode() for x in C.string_at(p, n.value).split(b"\0")[:-1])
    finally:
        lib.tc_free(p)


def validate_chroms(genome, vcf_file, sam_chroms, lib=None):
    """TC.py:463-516, same messages.  With `lib`, the VCF's chromosome column is collected by the library (tc_vcf_chroms: one
    pass in C++ instead of a Python loop over millions of records)."""
    fasta_chroms = set(genome.keys())
    sam_chroms = set(sam_chroms)
    sam_chroms.discard("*")
    if not sam_chroms.issubset(fasta_chroms):
        fmt = lambda s: "{" + ", ".join('"' + str(x) + '"' for x in s) + "}"
        raise RuntimeError("One or more SAM chromosomes were not found in the fasta reference.\nSAM chromosomes:\n"
                           + fmt(sam_chroms) + "\nFASTA chromosomes:\n" + fmt(fasta_chroms) + "\n"
                           "One common cause of this problem is when the fasta headers contain more than one word. "
                           "If this is the case, try trimming the headers to include only the chromosome name (i.e. '>chr1').")
    if vcf_file is not None:
        try:
            vcf_chroms = vcf_chrom_set(vcf_file, lib)
        except Exception as e:
            print(e)
            raise RuntimeError("Problem reading the provided VCF file. Please make sure that the filename is correct and "
                               "that your VCF file is properly formatted, including a header.")
        if len(sam_chroms & vcf_chroms) == 0:
            fmt = lambda s: "{" + ", ".join('"' + str(x) + '"' for x in s) + "}"
            raise RuntimeError("None of the chromosomes included in the VCF file matched the chromosomes in the provided "
                               "SAM file.\nSAM chromosomes:\n" + fmt(sam_chroms) + "\nVCF chromosomes:\n" + fmt(vcf_chroms) + "\n")


def sam_blocks(path, block_bytes):
    """The SAM file as blocks of about block_bytes that end at line boundaries: slices of a private mapping of the file
    (no copy: the parser threads read the page cache), or bytes objects read in a loop where the file cannot be mapped."""
    import mmap
    mm = None
    try:
        with open(path, "rb") as f:
            if os.fstat(f.fileno()).st_size > 0:
                mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_COPY)
    except (OSError, ValueError):
        mm = None
    if mm is not None:
        mv, n, a = memoryview(mm), len(mm), 0
        while a < n:
            b = n if a + block_bytes >= n else mm.rfind(b"\n", a, a + block_bytes) + 1
            if b <= a:                                    # a line longer than a block
                b = mm.find(b"\n", a + block_bytes) + 1 or n
            yield mv[a:b]
            a = b
        return
    with open(path, "rb") as f:
        rest = b""
        while True:
            buf = f.read(block_bytes)
            if not buf:
                if rest:
                    yield rest
                return
            cut = buf.rfind(b"\n")
            if cut < 0:
                rest += buf
                continue
            yield rest + buf[:cut + 1]
            rest = buf[cut + 1:]


def n_gpus_available(lib=None):
    """GPUs the library can open (asked of the CUDA runtime through the C ABI: importing torch for this cost seconds)."""
    try:
        from . import engine as E
        return int((lib or E.load_library()).tc_device_count())
    except Exception:
        return 1


def _hide_unused_gpus(want):
    """The CUDA runtime brings up every visible device when it starts (seconds on an eight-GPU box, during which the scan of the
    SAM file stalls as well): before the first CUDA call, CUDA_VISIBLE_DEVICES is narrowed to the `want` devices this run will
    open - the first ones of the current list, or of /dev/nvidia<N> when the variable is not set."""
    import re
    cur = os.environ.get("CUDA_VISIBLE_DEVICES")
    if cur is not None:
        ids = [x.strip() for x in cur.split(",") if x.strip()]
    else:
        try:
            ids = [str(k) for k in sorted(int(m.group(1)) for m in (re.fullmatch(r"nvidia(\d+)", f) for f in os.listdir("/dev")) if m)]
        except OSError:
            ids = []
    if len(ids) > want:
        os.environ["CUDA_VISIBLE_DEVICES"] = ",".join(ids[:want])


class _FileWriters:
    """One writer thread per output file: the four texts of a block are written side by side (the page-cache copies of
    the clean SAM, the FASTA and the TE log are what the writing end of the pipeline spends its time on)."""

    def __init__(self, files):
        import queue
        import threading
        self.files, self.q, self.err = files, {k: queue.Queue() for k in files}, []
        self.threads = [threading.Thread(target=self._run, args=(k,), daemon=True) for k in files]
        for t in self.threads:
            t.start()

    def _run(self, k):
        f, q = self.files[k], self.q[k]
        while True:
            item = q.get()
            if item is None:
                return
            pieces, done = item
            try:
                for piece in pieces:
                    f.write(piece)
            except BaseException as e:
                self.err.append(e)
            finally:
                done.set()

    def write(self, views):
        """Writes views[k] (a list of buffers) to file k for every k; returns when all of it is written."""
        import threading
        evs = []
        for k in self.files:
            ev = threading.Event()
            self.q[k].put((views[k], ev))
            evs.append(ev)
        for ev in evs:
            ev.wait()
        if self.err:
            raise self.err[0]

    def close(self):
        for k in self.files:
            self.q[k].put(None)
        for t in self.threads:
            t.join()


def split_text(text, n):
    """Cuts a SAM text (bytes) into n pieces at line boundaries, balanced by bytes.  Any
    contiguous cut reproduces the single-worker output order when the pieces are concatenated."""
    cuts = [0]
    for k in range(1, n):
        p = text.find(b"\n", len(text) * k // n)
        cuts.append(len(text) if p < 0 else p + 1)
    cuts.append(len(text))
    for k in range(1, len(cuts)):
        cuts[k] = max(cuts[k], cuts[k - 1])
    return [text[cuts[k]:cuts[k + 1]] for k in range(n)]


def run(options, engines=None):
    """engines: optional list of pre-built Engine objects (tests inject CPU stand-ins here).
    The SAM file is streamed twice in blocks (TC_CLI_BLOCK_MB, default 12): a scan for the header lines and the
    chromosome names (what split_SAM collects, TC.py:519-541), then the correction.  Blocks are handed to worker
    threads in input order - TC_CLI_WORKERS_PER_GPU (default: up to 4) workers per GPU, each with its own engine context, so
    that parsing block k+1, correcting block k and rendering block k-1 overlap - and their four texts are appended to
    the output files in input order (combine_outputs, TC.py:604-663: headers first; no SAM / FASTA in --dryRun).
    SAM parsing and text rendering run in the library's multi-threaded C++ host code (csrc/tc_host.cpp); the SJ / VCF
    tables are parsed once (csrc/tc_tables.cpp) and uploaded to every context."""
    from . import engine as E
    import time
    t_start = time.perf_counter()
    timing = os.environ.get("TC_CLI_TIMING") == "1"

    def stamp(what):                                      # TC_CLI_TIMING=1: seconds since the start of run(), to stderr
        if timing:
            print("cli %-28s %8.3f s" % (what, time.perf_counter() - t_start), file=sys.stderr)
    block = int(os.environ.get("TC_CLI_BLOCK_BYTES", "0")) or max(1, int(os.environ.get("TC_CLI_BLOCK_MB", "12"))) << 20
    lib = E.load_library() if engines is None else engines[0]._lib
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=max(2, min(16, os.cpu_count() or 2)))
    made = None
    if engines is None:                                   # the device contexts come up while the SAM file is scanned
        # GPUs: the text path is bound by the host's cores (parser / formatter threads, file writes), one GPU keeps about a
        # dozen of them busy, and opening a device context costs most of a second: one GPU per 12 host cores, at most
        # --threads of the visible ones (TC_CLI_GPUS overrides).  profiles/r02f_cli_wallclock_8gpu_box.json
        cores = os.cpu_count() or 4
        want = int(os.environ.get("TC_CLI_GPUS", "0")) or max(1, min(int(options.n_threads), max(1, cores // 12)))
        _hide_unused_gpus(want)
        n = max(1, min(want, max(1, n_gpus_available(lib))))
        # workers per GPU: four where the host has the cores for them (each worker's parser / formatter threads are the host's
        # cores divided by the workers), fewer on a box with many GPUs and few cores
        per_gpu = int(os.environ.get("TC_CLI_WORKERS_PER_GPU", "0")) or max(1, min(4, ((os.cpu_count() or 4) // 4) // n))
        made = [pool.submit(E.Engine, d) for d in range(n) for _ in range(per_gpu)]
    chroms, header = set(), ""
    pending = []

    def scan(blk):                                        # (C calls: the pool's threads run side by side)
        return E.sam_rnames(lib, blk), E.sam_header_lines(lib, blk)

    def drain(keep):
        nonlocal header
        while len(pending) > keep:
            names, head = pending.pop(0).result()
            chroms.update(names)
            header += head
    for blk in sam_blocks(options.sam, block):
        pending.append(pool.submit(scan, blk))
        drain(24)                                         # bounded: at most 24 blocks of the file in memory
    drain(0)
    stamp("scan of the SAM file")
    print("Reading genome ..............................")
    genome = GenomeImage.from_fasta_cached(options.refGenome, lib=lib)
    stamp("genome image")
    if made is not None:
        engines = [f.result() for f in made]
    pool.shutdown(wait=False)
    stamp("engine contexts")
    validate_chroms(genome, options.variantFile, chroms, lib=lib)
    os.makedirs(options.tmp_dir, exist_ok=True)          # kept for interface compatibility (--tmpDir / --deleteTmp)
    dry = bool(options.dryRun)
    # tables: parsed once and uploaded to every context; a dry run never opens the SJ / VCF files (TC.py:589-601)
    sj, var = (None, None) if dry else api.build_tables(options, chroms, genome)
    stamp("tables built")
    refs_list = [api.make_refs(genome, sj, var, engine=e) for e in engines]
    stamp("tables and genome on the GPUs")
    pre = options.outprefix
    names = {"log": pre + "_clean.log", "te": pre + "_clean.TE.log"}
    if not dry:
        names.update({"sam": pre + "_clean.sam", "fa": pre + "_clean.fa"})
    files = {k: open(v + ".partial", "wb") for k, v in names.items()}     # renamed on success: an abort leaves no truncated final file
    ok = False
    try:
        files["log"].write(LOG_HEADER.encode())
        files["te"].write(TE_HEADER.encode())
        if not dry:
            files["sam"].write(header.encode())

        writers = _FileWriters(files)

        def sink(_header, views):                         # the formatter threads' own buffers, written without a copy
            writers.write(views)

        try:
            api.correct_sam_stream(sam_blocks(options.sam, block), options, refs_list, sink, dry=dry)
        finally:
            writers.close()
        stamp("correction pass")
        ok = True
    finally:
        for f in files.values():
            f.close()
        for v in names.values():
            if ok:
                os.replace(v + ".partial", v)
            elif os.path.exists(v + ".partial"):
                os.remove(v + ".partial")
    stamp("files closed")
    if options.delete_tmp:
        import shutil
        shutil.rmtree(options.tmp_dir, ignore_errors=True)


def main(argv=None):
    run(get_options(argv))
    if argv is None:
        # the command line itself: the output files are closed and renamed; leaving through the interpreter's teardown would
        # spend another second unpinning host buffers and taking the device contexts down one by one
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()

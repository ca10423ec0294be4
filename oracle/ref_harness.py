"""Runs the UNMODIFIED reference in-process under oracle/stubs.

TEST / BENCH INFRASTRUCTURE ONLY.  The reference is looked for at /root/reference (the build container) and then at
oracle/_ref (the pip-installed copy `__graft_entry__.build()` makes there; git-ignored, but it travels to the GPU box,
where `bench.py --impl reference` times it).  Used by oracle/make_golden*.py to produce tests/golden/*.json.gz, by
tests/test_oracle_vs_reference.py (auto-skipped when no reference is found) and by bench.py's reference arm.
"""
import io
import os
import sys
import tempfile
import contextlib

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("TC_REF_ROOT") or "/root/reference"
if not os.path.isdir(os.path.join(REF_ROOT, "TranscriptClean")):
    REF_ROOT = os.path.join(_HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "TranscriptClean", "TranscriptClean.py"))


def load():
    """Import the reference package with the stub third-party modules in front."""
    stubs = os.path.join(_HERE, "stubs")
    for p in (REF_ROOT, stubs):
        if p not in sys.path:
            sys.path.insert(0, p)
    import pyfaidx, pyranges                      # noqa: the stubs
    import TranscriptClean.TranscriptClean as TC  # noqa
    from TranscriptClean.dstruct import Struct
    return TC, Struct, pyfaidx, pyranges


class _Mem:
    """In-memory genome for the pyfaidx stub: dict name -> str, or a SparseGenome."""

    def __init__(self, seqs):
        self.seqs = seqs

    def names(self):
        return list(self.seqs.keys())

    def fetch(self, name, start, end):
        s = self.seqs[name]
        if hasattr(s, "fetch"):
            return s.fetch(start, end)
        return s[start - 1:max(end, start - 1)]


def run_reference(lines, genome_seqs, sj_file=None, vcf_file=None, maxLenIndel=5, maxSJOffset=5,
                  correctMismatches=True, correctIndels=True, correctSJs=True, primaryOnly=False,
                  canonOnly=False, dryRun=False, strand_mode="same", tie="left", header=()):
    """One `--threads 1` worker of the reference: prep_refs + batch_correct (TC.py:560-586) or
    dryRun (TC.py:589-601).  Returns {"sam","fa","log","te"} texts (no headers)."""
    TC, Struct, pyfaidx, pyranges = load()
    pyranges.STRAND_MODE, pyranges.TIE = strand_mode, tie
    tmp = tempfile.mkdtemp(prefix="tc_ref_") + "/"
    gname = tmp + "genome.mem"
    pyfaidx.REGISTRY[gname] = _Mem(genome_seqs)
    o = Struct()
    o.refGenome, o.variantFile, o.sjAnnotFile, o.tmp_dir = gname, vcf_file, sj_file, tmp
    o.maxLenIndel, o.maxSJOffset = maxLenIndel, maxSJOffset
    tf = lambda b: "true" if b else "false"
    o.correctMismatches, o.correctIndels, o.correctSJs = tf(correctMismatches), tf(correctIndels), tf(correctSJs)
    o.primaryOnly, o.canonOnly, o.dryRun, o.buffer_size = primaryOnly, canonOnly, dryRun, 100
    out = Struct()
    out.sam, out.fasta, out.log, out.TElog = io.StringIO(), io.StringIO(), io.StringIO(), io.StringIO()
    sink = io.StringIO()
    import warnings
    with contextlib.redirect_stdout(sink), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(2)
        os.dup2(devnull, 2)                       # `samtools: not found` from os.system
        try:
            if dryRun:
                TC.dryRun(list(lines), o, out)
            else:
                refs = TC.prep_refs(o, list(lines), list(header))
                TC.batch_correct(list(lines), o, refs, out, buffer_size=100)
        finally:
            os.dup2(saved, 2)
            os.close(devnull)
            os.close(saved)
    del pyfaidx.REGISTRY[gname]
    os.system("rm -rf " + tmp)
    return {"sam": out.sam.getvalue(), "fa": out.fasta.getvalue(), "log": out.log.getvalue(),
            "te": out.TElog.getvalue()}


class Worker:
    """One worker process of the reference as run_chunk sets it up (TC.py:560-586): prep_refs once, then batch_correct on
    whatever lines it is given.  `correct(lines)` returns the four texts."""

    def __init__(self, genome_seqs, sj_file=None, vcf_file=None, chrom_lines=(), **opts):
        TC, Struct, pyfaidx, pyranges = load()
        self.TC, self.Struct = TC, Struct
        pyranges.STRAND_MODE, pyranges.TIE = opts.pop("strand_mode", "same"), opts.pop("tie", "left")
        self.tmp = tempfile.mkdtemp(prefix="tc_ref_") + "/"
        gname = self.tmp + "genome.mem"
        pyfaidx.REGISTRY[gname] = genome_seqs if hasattr(genome_seqs, "fetch") and hasattr(genome_seqs, "names") else _Mem(genome_seqs)
        o = Struct()
        o.refGenome, o.variantFile, o.sjAnnotFile, o.tmp_dir = gname, vcf_file, sj_file, self.tmp
        o.maxLenIndel, o.maxSJOffset = opts.get("maxLenIndel", 5), opts.get("maxSJOffset", 5)
        tf = lambda b: "true" if b else "false"
        o.correctMismatches, o.correctIndels, o.correctSJs = (tf(opts.get(k, True)) for k in ("correctMismatches", "correctIndels", "correctSJs"))
        o.primaryOnly, o.canonOnly, o.dryRun, o.buffer_size = opts.get("primaryOnly", False), opts.get("canonOnly", False), False, 100
        self.o = o
        with self._quiet():
            self.refs = TC.prep_refs(o, list(chrom_lines), [])

    @contextlib.contextmanager
    def _quiet(self):
        import warnings
        sink = io.StringIO()
        with contextlib.redirect_stdout(sink), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            devnull = os.open(os.devnull, os.O_WRONLY)
            saved = os.dup(2)
            os.dup2(devnull, 2)
            try:
                yield
            finally:
                os.dup2(saved, 2)
                os.close(devnull)
                os.close(saved)

    def correct(self, lines):
        out = self.Struct()
        out.sam, out.fasta, out.log, out.TElog = io.StringIO(), io.StringIO(), io.StringIO(), io.StringIO()
        with self._quiet():
            self.TC.batch_correct(list(lines), self.o, self.refs, out, buffer_size=100)
        return {"sam": out.sam.getvalue(), "fa": out.fasta.getvalue(), "log": out.log.getvalue(), "te": out.TElog.getvalue()}

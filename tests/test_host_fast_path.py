"""The C++ host code (csrc/tc_host.cpp: SAM parse + text rendering) against its Python twin
(sam.py / render.py) and against the oracle, on CPU through the hostsim engine."""
import numpy as np
import pytest

import pipeline
import synth
from hostsim_engine import HostSimEngine
from transcriptclean_b200 import api, engine as E, sam
from transcriptclean_b200.genome import GenomeImage
from transcriptclean_b200.tables import build_sj_tables, build_variant_tables


@pytest.fixture(scope="module")
def eng():
    e = HostSimEngine()
    yield e
    e.close()


def _text(case):
    return ("@HD\tVN:1.0\n@SQ\tSN:chr1\tLN:1\n" + "".join(l + "\n" for l in case.lines)).encode()


@pytest.mark.parametrize("seed", [401, 402, 403])
def test_parse_matches_python_parser(eng, seed):
    case = synth.make_case(seed, n_reads=150)
    gi = GenomeImage.from_dict(case.genome)
    st = eng.parse_sam(_text(case), gi, n_threads=3)
    want = sam.parse_lines(case.lines, gi).cols
    got = st.cols()
    assert st.header == "@HD\tVN:1.0\n@SQ\tSN:chr1\tLN:1\n"
    for k in want:
        assert np.array_equal(got[k], want[k]), k
    assert E.sam_rnames(eng._lib, _text(case)) == set(l.split("\t")[2] for l in case.lines)


@pytest.mark.parametrize("seed,opts", [(411, {}), (412, {"canonOnly": True, "maxSJOffset": 8}), (413, {"primaryOnly": True})])
def test_fast_path_texts_match_oracle(eng, seed, opts):
    case = synth.make_case(seed, n_reads=160)
    gi = GenomeImage.from_dict(case.genome)
    sjf, vcf = pipeline.write_tables(case.sj_rows, case.vcf_rows)
    chroms = set(l.split("\t")[2] for l in case.lines)
    refs = api.make_refs(gi, build_sj_tables(sjf, chroms, gi.index), build_variant_tables(vcf, 5, gi.index), engine=eng)
    o = api.Struct(maxSJOffset=opts.get("maxSJOffset", 5), primaryOnly=opts.get("primaryOnly", False), canonOnly=opts.get("canonOnly", False))
    header, got = api.correct_sam_text(_text(case), o, refs)
    want = pipeline.oracle_texts(case.genome, case.lines, sjf, vcf, opts)
    got = {k: v.decode() for k, v in got.items()}
    d = pipeline.first_diff(got, want)
    assert d is None, d


def test_fast_path_dry_run(eng):
    case = synth.make_case(421, n_reads=120, quirks=False)
    gi = GenomeImage.from_dict(case.genome)
    refs = api.make_refs(gi, engine=eng)
    _, got = api.correct_sam_text(_text(case), api.Struct(), refs, dry=True)
    want = pipeline.oracle_texts(case.genome, case.lines, None, None, {}, dry=True)
    assert {k: v.decode() for k, v in got.items()} == want


@pytest.mark.parametrize("bad,msg", [
    ("r\t0\tchr1\t10\t60\t5M\t*\t0\t0\tACGT\t*", "does not match CIGAR"),
    ("r\t0\tchr1\t10\t60\t4=\t*\t0\t0\tACGT\t*", "unsupported CIGAR"),
    ("r\t0\tchr1\t10\t60\t*\t*\t0\t0\tACGT\t*", "unsupported CIGAR"),
    ("r\t0\tchr1\t0\t60\t4M\t*\t0\t0\tACGT\t*", "runs off contig"),
    ("r\t0\tchr1\t999999999\t60\t4M\t*\t0\t0\tACGT\t*", "runs off contig"),
    ("r\tx\tchr1\t10\t60\t4M\t*\t0\t0\tACGT\t*", "FLAG is not an integer"),
    ("r\t0\tchr1\t10\t60\t4M", "fewer than 11 fields"),
    ("r\t0\tchr1\t10\t60\t2M5N2M\t*\t0\t0\tACGT\t*\tjI:B:i,12", "jI tag"),
    ("r\t0\tchr1\t10\t60\t4M\t*\t0\t0\tACGT\t*\tNM:i:0\tMD:4", "malformed MD"),
])
def test_unsupported_records_raise_the_same_error_in_both_parsers(eng, bad, msg):
    gi = GenomeImage.from_dict({"chr1": "ACGT" * 50})
    with pytest.raises(sam.SamFormatError, match=msg):
        sam.parse_lines([bad], gi)
    with pytest.raises(sam.SamFormatError, match=msg):
        eng.parse_sam((bad + "\n").encode(), gi)


def test_seam_functions_write_the_same_files_through_both_paths(eng, monkeypatch):
    """api.batch_correct / api.dryRun (the reference's seam: a list of SAM lines in, four open files written) through the C++ text
    path (any list of 16 lines or more) and through the Python twins (TC_PY_TEXT=1): same bytes, equal to the oracle's texts.
    Lines handed over with their line ends (readlines()) are taken as well."""
    import io
    case, clean = synth.make_case(431, n_reads=140), synth.make_case(432, n_reads=120, quirks=False)
    sjf, vcf = pipeline.write_tables(case.sj_rows, case.vcf_rows)
    chroms = set(l.split("\t")[2] for l in case.lines)
    gi = GenomeImage.from_dict(case.genome)
    refs = api.make_refs(gi, build_sj_tables(sjf, chroms, gi.index), build_variant_tables(vcf, 5, gi.index), engine=eng)
    want = pipeline.oracle_texts(case.genome, case.lines, sjf, vcf, {})

    def run(lines, refs, dry=False):
        out = api.Struct(sam=io.StringIO(), fasta=io.StringIO(), log=io.StringIO(), TElog=io.StringIO())
        if dry:
            api.dryRun(lines, api.Struct(), out, refs=refs)
        else:
            api.batch_correct(lines, api.Struct(), refs, out)
        return {"sam": out.sam.getvalue(), "fa": out.fasta.getvalue(), "log": out.log.getvalue(), "te": out.TElog.getvalue()}
    called = []
    real = api.correct_sam_text
    monkeypatch.setattr(api, "correct_sam_text", lambda *a, **k: (called.append(1), real(*a, **k))[1])
    fast = run(case.lines, refs)
    assert len(called) == 1 and pipeline.first_diff(fast, want) is None
    assert run([l + "\n" for l in case.lines], refs) == fast
    few = run(case.lines[:5], refs)                                    # a handful of lines: the twins
    assert len(called) == 2 and few["log"] == "".join(fast["log"].splitlines(True)[:5])
    monkeypatch.setenv("TC_PY_TEXT", "1")
    assert run(case.lines, refs) == fast and len(called) == 2
    monkeypatch.delenv("TC_PY_TEXT")
    # real text-mode files: the bytes go to the binary file underneath, behind what the caller wrote before
    import tempfile
    d = tempfile.mkdtemp(prefix="tc_seam_")
    out = api.Struct(sam=open(d + "/s", "w"), fasta=open(d + "/f", "w"), log=open(d + "/l", "w"), TElog=open(d + "/t", "w"))
    out.sam.write("@HD\tVN:1.0\n"); out.log.write("header\n")
    api.batch_correct(case.lines, api.Struct(), refs, out)
    out.sam.write("tail\n")
    for f in out.values():
        f.close()
    assert open(d + "/s").read() == "@HD\tVN:1.0\n" + fast["sam"] + "tail\n" and open(d + "/l").read() == "header\n" + fast["log"]
    assert open(d + "/f").read() == fast["fa"] and open(d + "/t").read() == fast["te"]
    # dry run (a case without the records the reference's dryRun dies on)
    refs_d = api.make_refs(GenomeImage.from_dict(clean.genome), engine=eng)
    want_dry = pipeline.oracle_texts(clean.genome, clean.lines, None, None, {}, dry=True)
    fast_dry = run(clean.lines, refs_d, dry=True)
    assert len(called) == 4 and (fast_dry["log"], fast_dry["te"]) == (want_dry["log"], want_dry["te"]) and fast_dry["sam"] == fast_dry["fa"] == ""
    monkeypatch.setenv("TC_PY_TEXT", "1")
    assert run(clean.lines, refs_d, dry=True) == fast_dry and len(called) == 4

"""GPU parity on the BASELINE.json workload shapes (tools/synth_gen.c) at sizes the oracle
finishes in seconds, and size-independent properties at a large size."""
import os
import tempfile

import numpy as np
import pytest

import goldens
import pipeline

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    from tools.workload import Workload
    from transcriptclean_b200.engine import Engine
    w = Workload(seed=77, chrom_len=24_000_000, n_genes=300)
    gi = w.genome_image()
    d = tempfile.mkdtemp(prefix="tc_wl_")
    sjf = goldens.write_rows(w.sj_rows(), d + "/sj.tab")
    eng = Engine(0)
    yield w, gi, sjf, eng, d
    eng.close()
    w.close()


def _compare(setup, shape, n, opts, vcf_rows=None, dry=False):
    w, gi, sjf, eng, d = setup
    cols, nm = w.reads(shape, n, seed_offset=3)
    lines = w.sam_lines(cols, nm, 0, n)
    vcf = goldens.write_rows(vcf_rows, d + "/v_%s.vcf" % shape) if vcf_rows else None
    got = pipeline.product_texts(eng, gi, lines, sjf, vcf, opts, dry=dry)
    want = pipeline.oracle_texts({w.chrom: w.genome.tobytes().decode()}, lines, sjf, vcf, opts, dry=dry)
    diff = pipeline.first_diff(got, want)
    assert diff is None, diff
    return want


def test_config1_pacbio_shape(setup):
    want = _compare(setup, "pacbio", 5000, {})
    assert want["te"].count("NC_SJ_boundary\t") > 50 and want["te"].count("Corrected") > 10000


def test_config2_variant_aware(setup):
    from tools.workload import variants_from_reads
    w = setup[0]
    cols, nm = w.reads("pacbio", 4000, seed_offset=3)
    rows = variants_from_reads(w, cols, 0, 4000, frac=0.3)
    want = _compare(setup, "pacbio", 4000, {}, vcf_rows=rows)
    assert want["te"].count("VariantMatch") > 3000


def test_config3_ont_shape_canon_only(setup):
    want = _compare(setup, "ont", 3000, {"canonOnly": True})
    assert want["te"].count("TooLarge") > 500 and want["sam"].count("\n") < 3000


def test_config4_dry_run_then_full(setup):
    _compare(setup, "ont", 2000, {}, dry=True)
    _compare(setup, "ont", 2000, {"primaryOnly": True})


def test_large_batch_properties(setup):
    """300k reads: the pipelined host-buffer path and the device-resident path agree byte for
    byte; counters account for every TE row; every new SEQ has the length its CIGAR implies."""
    from transcriptclean_b200 import engine as E
    from transcriptclean_b200.tables import build_sj_tables
    w, gi, sjf, eng, d = setup
    eng.set_options()
    eng.load_genome(gi)
    eng.load_junctions(build_sj_tables(sjf, {w.chrom}, gi.index))
    from transcriptclean_b200.tables import VariantTables
    eng.load_variants(VariantTables())
    n = 300000
    cols, nm = w.reads("pacbio", n, seed_offset=9)
    eng.set_pipeline(50000)
    a = eng.correct(cols)                  # pipelined, 6 chunks
    eng.set_pipeline(0)
    b = eng.correct(cols)                  # one piece
    eng.set_pipeline(131072)
    for name in ("status", "counters", "nm", "seq_len", "cig_op", "cig_len", "jm", "ji", "te_off", "cig_off", "jn_off", "delta_len"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    assert np.array_equal(a.te, b.te)
    assert np.array_equal(a.md_len, b.md_len)
    # SEQ rows (padding bytes excluded) and MD values compared through their offsets
    idx = np.flatnonzero(a.seq_len > 0)[::97]
    for i in idx:
        assert a.read_seq(i) == b.read_seq(i) and len(a.read_seq(i)) == int(a.seq_len[i])
        assert a.md[a.md_off[i]:a.md_off[i] + a.md_len[i]].tobytes() == b.md[b.md_off[i]:b.md_off[i] + b.md_len[i]].tobytes()
    primary = (a.status & E.ST_CLASS_MASK) == 0
    ok = primary & ((a.status & E.ST_FALLBACK) == 0)
    c = np.where(a.counters < 0, 0, a.counters)
    assert np.array_equal(c[ok].sum(axis=1), np.diff(a.te_off)[ok])
    qlen = np.where(np.isin(a.cig_op, np.frombuffer(b"MIS", dtype=np.uint8)), a.cig_len, 0).astype(np.int64)
    per = np.add.reduceat(np.concatenate((qlen, [0])), np.minimum(a.cig_off[:-1], len(qlen)))
    per = np.where(np.diff(a.cig_off) > 0, per, 0)
    rew = primary & ((a.status & E.ST_CIGAR_REWRITTEN) != 0)        # (an untouched CIGAR is not sent back)
    assert np.array_equal(per[rew], a.seq_len[rew].astype(np.int64)) and int(rew.sum()) > 0.9 * int(primary.sum())
    assert int((a.status & (E.ST_ERR_INTRON | E.ST_ERR_MERGE)).sum()) == 0


def test_sam_text_fast_path_on_gpu(setup):
    """SAM text in, four texts out: C++ host parse/render around the CUDA path equals the oracle."""
    from transcriptclean_b200 import api
    from transcriptclean_b200.tables import build_sj_tables
    w, gi, sjf, eng, d = setup
    cols, nm = w.reads("ont", 2500, seed_offset=5)
    lines = w.sam_lines(cols, nm, 0, 2500)
    refs = api.make_refs(gi, build_sj_tables(sjf, {w.chrom}, gi.index), None, engine=eng)
    header, got = api.correct_sam_text(("@HD\tVN:1.0\n" + "\n".join(lines) + "\n").encode(), api.Struct(canonOnly=True), refs)
    want = pipeline.oracle_texts({w.chrom: w.genome.tobytes().decode()}, lines, sjf, None, {"canonOnly": True})
    assert header == "@HD\tVN:1.0\n"
    diff = pipeline.first_diff({k: v.decode() for k, v in got.items()}, want)
    assert diff is None, diff


def test_plane_offsets_beyond_2_31(setup):
    """hg38-scale layout: the contig the reads map to starts 2.3 Gb into the genome planes, so every
    plane index exceeds 2^31 (chromosome id 1 in the table keys).  Results must equal the ones of
    the same reads on the small single-contig image."""
    from tools.workload import variants_from_reads
    from transcriptclean_b200.genome import GenomeImage
    from transcriptclean_b200.tables import build_sj_tables, build_variant_tables
    w, gi, sjf, eng, d = setup
    n = 20000
    cols, nm = w.reads("ont", n, seed_offset=11)
    vcf = goldens.write_rows(variants_from_reads(w, cols, 0, 3000, frac=0.3), d + "/v_big.vcf")

    def run(image, cols):
        eng.set_options()
        eng.load_genome(image)
        eng.load_junctions(build_sj_tables(sjf, {w.chrom}, image.index))
        eng.load_variants(build_variant_tables(vcf, 5, image.index))
        return eng.correct(cols)

    small = run(gi, cols)
    big = GenomeImage(["filler", w.chrom], [2_300_000_000, w.chrom_len])
    big._put(int(big.chrom_off[1]), w.genome)
    assert int(big.chrom_off[1]) > 2 ** 31
    cols2 = dict(cols)
    cols2["chrom"] = np.where(cols["chrom"] == 0, 1, cols["chrom"]).astype(np.int32)
    large = run(big, cols2)
    for name in ("status", "counters", "nm", "seq_len", "delta_len", "cig_off", "cig_op", "cig_len", "jn_off", "jm", "ji", "te_off", "md_len"):
        assert np.array_equal(getattr(small, name), getattr(large, name)), name
    assert np.array_equal(small.te, large.te)
    for i in range(0, n, 7):
        assert small.read_seq(i) == large.read_seq(i), i
        assert small.md[small.md_off[i]:small.md_off[i] + small.md_len[i]].tobytes() == large.md[large.md_off[i]:large.md_off[i] + large.md_len[i]].tobytes(), i
    assert int((np.asarray(small.counters)[:, 2] > 0).sum()) + int((np.asarray(small.counters)[:, 5] > 0).sum()) > 100   # variant hits happened


def test_command_line_on_the_gpu(monkeypatch):
    """The `transcriptclean` command line itself on the GPU: files in, the four files out, blocks of a few kB
    through three pipelined workers per GPU on every visible GPU; --dryRun writes the two logs only."""
    import synth
    from transcriptclean_b200 import cli
    case = synth.make_case(4711, n_reads=400)
    d = tempfile.mkdtemp(prefix="tc_cli_gpu_")
    case.write_fasta(d + "/g.fa"); case.write_sam(d + "/in.sam"); case.write_sj(d + "/sj.tab"); case.write_vcf(d + "/v.vcf")
    monkeypatch.setenv("TC_CLI_BLOCK_BYTES", "20000")
    o = cli.get_options(["-s", d + "/in.sam", "-g", d + "/g.fa", "-j", d + "/sj.tab", "-v", d + "/v.vcf", "-t", "8", "-o", d + "/TC", "--canonOnly"])
    cli.run(o)
    want = pipeline.oracle_texts(case.genome, case.lines, d + "/sj.tab", d + "/v.vcf", {"canonOnly": True})
    sam = open(d + "/TC_clean.sam").read()
    assert sam.startswith("@HD")
    assert "".join(l + "\n" for l in sam.splitlines() if not l.startswith("@")) == want["sam"]
    assert open(d + "/TC_clean.fa").read() == want["fa"]
    assert open(d + "/TC_clean.log").read() == cli.LOG_HEADER + want["log"]
    assert open(d + "/TC_clean.TE.log").read() == cli.TE_HEADER + want["te"]
    assert not [f for f in os.listdir(d) if f.endswith(".partial")]
    o = cli.get_options(["-s", d + "/in.sam", "-g", d + "/g.fa", "--dryRun", "-o", d + "/DRY"])
    clean = synth.make_case(4711, n_reads=400, quirks=False)
    clean.write_sam(d + "/in.sam")
    cli.run(o)
    want = pipeline.oracle_texts(clean.genome, clean.lines, None, None, {}, dry=True)
    assert open(d + "/DRY_clean.log").read() == cli.LOG_HEADER + want["log"]
    assert open(d + "/DRY_clean.TE.log").read() == cli.TE_HEADER + want["te"]
    assert not os.path.exists(d + "/DRY_clean.sam")

"""Parity checks shared by the CPU (hostsim) and GPU test modules: every check takes an engine
and compares the product pipeline with (a) outputs of the unmodified reference stored under
tests/golden and (b) the oracle on freshly generated seeded cases."""
import goldens
import pipeline
import synth
from transcriptclean_b200.genome import GenomeImage


def check_fixture(engine, rec):
    gi = GenomeImage.from_windows(rec["chrom_len"], rec["windows"])
    have_sj = len(rec["sj_rows"]) > 0
    sjf, vcf = pipeline.write_tables(rec["sj_rows"] if have_sj else None, rec["vcf_rows"])
    got = pipeline.product_texts(engine, gi, rec["lines"], sjf, vcf, rec["opts"])
    d = pipeline.first_diff(got, rec["expected"])
    assert d is None, d
    got = pipeline.product_texts(engine, gi, rec["lines"], None, None, {}, dry=True)
    d = pipeline.first_diff(got, rec["expected_dry"])
    assert d is None, d
    if rec["answer_matches"]:
        assert got is not None and pipeline.product_texts(engine, gi, rec["lines"], sjf, vcf, rec["opts"])["sam"].split() \
            == rec["ref_answer_line"].split()


def check_synth_golden(engine, seed):
    rec = goldens.load_synth(seed)
    gi = GenomeImage.from_dict(rec["genome"])
    sjf, vcf = pipeline.write_tables(rec["sj_rows"], rec["vcf_rows"])
    for run in rec["runs"]:
        got = pipeline.product_texts(engine, gi, rec["lines"], sjf if run["use_sj"] else None,
                                     vcf if run["use_vcf"] else None, run["opts"])
        d = pipeline.first_diff(got, run["expected"])
        assert d is None, (run["opts"], d)
    got = pipeline.product_texts(engine, GenomeImage.from_dict(rec["dry_genome"]), rec["dry_lines"], None, None, {}, dry=True)
    d = pipeline.first_diff(got, rec["expected_dry"])
    assert d is None, d


MODES = (
    (dict(), True, False),
    (dict(), True, True),
    (dict(primaryOnly=True, correctMismatches=False), False, False),
    (dict(canonOnly=True, maxSJOffset=8, maxLenIndel=3), True, True),
    (dict(correctIndels=False), True, True),
    (dict(correctSJs=False), True, True),
)
UNPINNED_MODES = (
    (dict(sjIgnoreStrand=True, sjTieRight=True), True, True),
    (dict(sjIgnoreStrand=False, sjTieRight=True), True, False),
    (dict(sjIgnoreStrand=True, sjTieRight=False), True, False),
)


def check_generated(engine, seed, n_reads=150, unpinned=False, **kw):
    case = synth.make_case(seed, n_reads=n_reads, unpinned=unpinned, **kw)
    sjf, vcf = pipeline.write_tables(case.sj_rows, case.vcf_rows)
    gi = GenomeImage.from_dict(case.genome)
    for opts, us, uv in (MODES + UNPINNED_MODES if unpinned else MODES):
        got = pipeline.product_texts(engine, gi, case.lines, sjf if us else None, vcf if uv else None, opts)
        want = pipeline.oracle_texts(case.genome, case.lines, sjf if us else None, vcf if uv else None, opts)
        d = pipeline.first_diff(got, want)
        assert d is None, (seed, opts, d)
    clean = synth.make_case(seed, n_reads=n_reads, quirks=False, **kw)
    got = pipeline.product_texts(engine, GenomeImage.from_dict(clean.genome), clean.lines, None, None, {}, dry=True)
    want = pipeline.oracle_texts(clean.genome, clean.lines, None, None, {}, dry=True)
    d = pipeline.first_diff(got, want)
    assert d is None, (seed, "dry", d)


def check_seq_forms(engine, seed, n_reads=160):
    """SEQ as 4-bit codes (tc_batch_in.seq_poff) and as ASCII give the same four texts, through the
    column path and through the C++ SAM-text path; a SEQ byte outside the alphabet keeps the batch ASCII."""
    import numpy as np
    from transcriptclean_b200 import api, sam
    from transcriptclean_b200.tables import build_sj_tables, build_variant_tables
    case = synth.make_case(seed, n_reads=n_reads)
    sjf, vcf = pipeline.write_tables(case.sj_rows, case.vcf_rows)
    gi = GenomeImage.from_dict(case.genome)
    want = pipeline.oracle_texts(case.genome, case.lines, sjf, vcf, {})
    chroms = set(l.split("\t")[2] for l in case.lines)
    text = "".join(l + "\n" for l in case.lines).encode()
    try:
        for seq2, seq4 in ((True, True), (False, True), (False, False)):          # 2-bit + listed bases, 4-bit codes, ASCII
            engine.seq2, engine.seq4, engine.seq4_text = seq2, seq4, seq4
            got = pipeline.product_texts(engine, gi, case.lines, sjf, vcf, {})
            d = pipeline.first_diff(got, want)
            assert d is None, (seq2, seq4, d)
            refs = api.make_refs(gi, build_sj_tables(sjf, chroms, gi.index), build_variant_tables(vcf, 5, gi.index), engine=engine)
            _, got = api.correct_sam_text(text, api.Struct(), refs)
            d = pipeline.first_diff({k: v.decode() for k, v in got.items()}, want)
            assert d is None, (seq4, "text path", d)
        engine.seq4 = True
        cols = sam.parse_lines(case.lines, gi).cols
        p2 = engine.pack_cols2(cols, max_exc_frac=1.0)      # explicit 2-bit columns in: every base outside ACGT is listed
        assert p2 is not None and p2["seq_bits"] == 2 and p2["seq"].size < cols["seq"].size * 0.27 + 4 * len(case.lines) + 64
        r2 = engine.correct(p2)
        packed = engine.pack_cols(cols)
        if engine.alphabet is None:                     # a genome with more than 16 distinct bytes: ASCII only
            assert packed is None and len(set("".join(case.genome.values())) | set("ACGTacgt")) > 16
            return False
        assert packed is not None and packed["seq"].size < cols["seq"].size * 0.52 + 4 * len(case.lines) + 64
        res = engine.correct(packed)                    # explicit packed columns in, ASCII view out
        engine.seq4 = False
        ref = engine.correct(cols)
        assert np.array_equal(res.seq_len, ref.seq_len) and np.array_equal(res.delta_len, ref.delta_len)
        for i in range(len(case.lines)):
            assert res.read_seq(i) == ref.read_seq(i) == r2.read_seq(i) and len(res.read_seq(i)) == int(ref.seq_len[i]), i
        assert np.array_equal(r2.status, ref.status) and np.array_equal(r2.te_words, ref.te_words) and np.array_equal(r2.nm, ref.nm)
        engine.cig16 = True                              # CIGAR as 16-bit words (introns and other long ops listed) against op + length arrays
        pc16 = engine.pack_cigar(cols)
        assert "cig16" in pc16 and len(pc16["cig_exc_idx"]) < len(cols["cig_op"])     # (introns of 4,096 bases and more are listed: tools/workload shapes)
        r16 = engine.correct(pc16)
        engine.cig16 = False
        r40 = engine.correct(cols)
        for name in ("status", "counters", "nm", "te_words", "cig_op", "cig_len", "jm", "ji", "delta_len", "md_len"):
            assert np.array_equal(getattr(r16, name), getattr(r40, name)), name
        engine.seq4 = True
        odd = dict(cols)
        odd["seq"] = cols["seq"].copy()
        if odd["seq"].size:
            odd["seq"][0] = ord("!")
        assert engine.pack_cols(odd) is None
        return True
    finally:
        engine.seq2, engine.seq4, engine.seq4_text, engine.cig16 = True, True, False, True


def check_plane_and_ascii_paths(engine, seed, n_reads=300, **kw):
    """The NM / MD check on the 2-bit plane (k_verify) and the ASCII path alone (a context made under TC_NO_PLANE=1: k_emit
    decides every read) give the same results, for SEQ sent as 2-bit rows, as 4-bit codes and as ASCII; the plane path really
    settles reads (fewer reads reach the ASCII path than the batch holds)."""
    import os
    import numpy as np
    from transcriptclean_b200 import sam
    from transcriptclean_b200.engine import Engine
    from transcriptclean_b200.tables import build_sj_tables, build_variant_tables
    case = synth.make_case(seed, n_reads=n_reads, **kw)
    sjf, vcf = pipeline.write_tables(case.sj_rows, case.vcf_rows)
    gi = GenomeImage.from_dict(case.genome)
    chroms = set(l.split("\t")[2] for l in case.lines)
    cols = sam.parse_lines(case.lines, gi).cols
    os.environ["TC_NO_PLANE"] = "1"
    try:
        plain = Engine(engine.device)
    finally:
        del os.environ["TC_NO_PLANE"]
    names = ("status", "counters", "nm", "te_words", "cig_op", "cig_len", "jm", "ji", "delta_len", "md_len", "seq_len")
    try:
        for e in (engine, plain):
            e.load_genome(gi); e.load_junctions(build_sj_tables(sjf, chroms, gi.index)); e.load_variants(build_variant_tables(vcf, 5, gi.index))
        settled = 0
        for seq2, seq4 in ((True, True), (False, True), (False, False)):
            for e in (engine, plain):
                e.seq2, e.seq4 = seq2, seq4
            a, b = engine.correct(cols), plain.correct(cols)
            ta, tb = engine.timing(), plain.timing()
            assert tb["ascii_path_reads"] == len(case.lines) and ta["ascii_path_reads"] <= len(case.lines)
            settled += len(case.lines) - ta["ascii_path_reads"]
            for name in names:
                assert np.array_equal(getattr(a, name), getattr(b, name)), (seq2, seq4, name)
            for i in range(len(case.lines)):                       # (MD values lie in arbitrary order in the pool)
                ma = a.md[int(a.md_off[i]):int(a.md_off[i]) + int(a.md_len[i])].tobytes()
                mb = b.md[int(b.md_off[i]):int(b.md_off[i]) + int(b.md_len[i])].tobytes()
                assert ma == mb and a.read_seq(i) == b.read_seq(i), (seq2, seq4, i, ma, mb)
        return settled
    finally:
        plain.close()
        engine.seq2, engine.seq4 = True, True

"""The C++ table builders (csrc/tc_tables.cpp: tc_vcf_parse, tc_sj_parse) against their Python twins and against
the oracle's own builders, on seeded cases with multi-allelic records, gzip input, rows on foreign contigs."""
import gzip
import os
import tempfile

import numpy as np
import pytest

import pipeline
import synth
from oracle import tc_oracle as oc
from transcriptclean_b200 import tables
from transcriptclean_b200.genome import GenomeImage


def _case(seed):
    case = synth.make_case(seed, n_reads=120, unpinned=(seed % 2 == 0))
    gi = GenomeImage.from_dict(case.genome)
    sjf, vcf = pipeline.write_tables(case.sj_rows + [["chrZZ", 5, 900, 1, 1, 1, 3, 0, 20]], case.vcf_rows + [["chrZZ", 10, ".", "A", "T", ".", "PASS", "."]])
    return case, gi, sjf, vcf


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_sj_tables_cpp_equals_python(seed):
    case, gi, sjf, _ = _case(seed)
    chroms = set(l.split("\t")[2] for l in case.lines)
    a, b = tables.build_sj_tables(sjf, chroms, gi.index), tables.build_sj_tables_py(sjf, chroms, gi.index)
    for k in ("don_s", "acc_s", "don_u", "acc_u", "ann_k1", "ann_k2"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert len(a) == len(b) > 0
    one = {sorted(c for c in chroms if c in gi.index)[0]}
    a, b = tables.build_sj_tables(sjf, one, gi.index), tables.build_sj_tables_py(sjf, one, gi.index)
    assert np.array_equal(a.don_s, b.don_s) and np.array_equal(a.ann_k2, b.ann_k2) and 0 < len(a.don_s)


@pytest.mark.parametrize("seed,max_len", [(11, 5), (12, 3), (13, 8)])
def test_variant_tables_cpp_equals_python(seed, max_len):
    case, gi, _, vcf = _case(seed)
    gz = vcf + ".gz"
    with open(vcf, "rb") as f, gzip.open(gz, "wb") as g:
        g.write(f.read())
    b = tables.build_variant_tables_py(vcf, max_len, gi.index)
    for path in (vcf, gz):
        a = tables.build_variant_tables(path, max_len, gi.index)
        assert sorted(zip(a.snp_key.tolist(), a.snp_alt.tolist())) == sorted(zip(b.snp_key.tolist(), b.snp_alt.tolist()))
        assert np.array_equal(a.snp_key, np.sort(a.snp_key))
        def ins_rows(v):
            return sorted((int(k), int(s), v.ins_bytes[v.ins_aoff[i]:v.ins_aoff[i + 1]].tobytes()) for i, (k, s) in enumerate(zip(v.ins_key, v.ins_size)))
        assert ins_rows(a) == ins_rows(b) and np.array_equal(a.ins_key, b.ins_key) and np.array_equal(a.ins_size, b.ins_size)
        assert np.array_equal(a.del_key, b.del_key) and np.array_equal(a.del_size, b.del_size)
    # the oracle's own dicts hold the same keys (TC.py:814-854)
    ov = oc.VariantTables.from_file(vcf, max_len)
    names = gi.names
    assert {(names[k >> 32], k & 0xffffffff) for k in b.snp_key.tolist()} == {k for k in ov.snps if k[0] in gi.index}
    assert {(names[k >> 32], k & 0xffffffff, (k & 0xffffffff) + s) for k, s in zip(b.del_key.tolist(), b.del_size.tolist())} == \
        {k for k in ov.dels if k[0] in gi.index}
    only = tables.build_variant_tables(vcf, max_len, gi.index, read_chroms={names[0]})
    assert 0 < len(only.snp_key) < len(b.snp_key) and set((only.snp_key >> np.uint64(32)).tolist()) == {0}


def test_bad_rows_raise():
    d = tempfile.mkdtemp()
    gi = GenomeImage.from_dict({"chr1": "ACGT" * 50})
    open(d + "/bad.vcf", "w").write("#h\nchr1\tx\t.\tA\tT\n")
    with pytest.raises(ValueError):
        tables.build_variant_tables(d + "/bad.vcf", 5, gi.index)
    open(d + "/bad.tab", "w").write("chr1\t5\tx\t1\t1\t1\t1\t1\t1\n")
    with pytest.raises(ValueError):
        tables.build_sj_tables(d + "/bad.tab", {"chr1"}, gi.index)
    open(d + "/short.tab", "w").write("chr1\t5\t9\t1\n")
    with pytest.raises(ValueError):
        tables.build_sj_tables(d + "/short.tab", {"chr1"}, gi.index)
    open(d + "/other.tab", "w").write("chr9\t5\tx\t1\n")             # a contig without reads is dropped before anything is parsed (TC.py:685)
    assert len(tables.build_sj_tables(d + "/other.tab", {"chr1"}, gi.index)) == 0


def _merged(runs):
    out = []
    for s, e, c in sorted(runs):
        if out and out[-1][1] == s and out[-1][2] == c:
            out[-1] = (out[-1][0], e, c)
        else:
            out.append((s, e, c))
    return out


@pytest.mark.parametrize("eol,width,threads", [("\n", 60, 1), ("\r\n", 80, 4), ("\n", 7, 3), ("\n", 1000, 8)])
def test_fasta_packer_cpp_equals_numpy(eol, width, threads):
    """tc_fasta_open / tc_fasta_pack (the library's multi-threaded FASTA -> genome image) against GenomeImage.from_fasta: names,
    lengths, both planes bit for bit, the literal runs, and get_seq windows - soft-masked stretches, N / IUPAC runs, contigs of
    0 / 1 / 15..65 bases (word boundaries of both planes), CRLF line ends, no newline at the end of the file."""
    import random
    from transcriptclean_b200 import engine
    lib = engine.load_library()
    rng = random.Random(hash((eol, width)) & 0xffff)

    def seq(n):
        out, have = [], 0
        while have < n:
            k, L = rng.random(), rng.randint(1, 300)
            piece = ("".join(rng.choice("ACGT") for _ in range(L)) if k < 0.55 else "".join(rng.choice("acgt") for _ in range(L)) if k < 0.8
                     else "N" * rng.randint(1, 90) if k < 0.9 else "n" * rng.randint(1, 40) if k < 0.95 else "".join(rng.choice("RYKMSWrykm") for _ in range(rng.randint(1, 5))))
            out.append(piece); have += len(piece)
        return "".join(out)[:n]
    recs = [("chr%d some description" % k, seq(n)) for k, n in enumerate([0, 1, 15, 16, 17, 31, 32, 33, 63, 64, 65, 1000, 5000, 33333])]
    d = tempfile.mkdtemp(prefix="tc_fa_")
    with open(d + "/g.fa", "w", newline="") as f:
        for name, s in recs:
            f.write(">" + name + eol)
            for i in range(0, len(s), width):
                f.write(s[i:i + width] + eol)
        f.write(">last\tx" + eol + "ACGTnnRY")                      # no line end behind the last base
    a, b = GenomeImage.from_fasta(d + "/g.fa"), GenomeImage.from_fasta_lib(d + "/g.fa", lib, threads)
    assert a.names == b.names == ["chr%d" % k for k in range(14)] + ["last"]
    assert np.array_equal(a.chrom_len, b.chrom_len) and [int(x) for x in b.chrom_len[:14]] == [len(s) for _, s in recs]
    assert np.array_equal(a.bases2, b.bases2) and np.array_equal(a.lower1, b.lower1)
    assert _merged(a._exc) == _merged(b._exc) and len(b._exc) > 3
    for name, s in recs[-3:]:
        c = name.split()[0]
        for _ in range(20):
            x = rng.randint(1, len(s)); y = min(len(s), x + rng.randint(0, 200))
            assert b.get_seq(c, x, y) == s[x - 1:y]
    assert b.get_seq("last", 1, 8) == "ACGTnnRY"


def test_fasta_packer_errors_and_cache():
    from transcriptclean_b200 import engine
    lib = engine.load_library()
    d = tempfile.mkdtemp(prefix="tc_fa_")
    with pytest.raises(ValueError):
        GenomeImage.from_fasta_lib(d + "/missing.fa", lib)
    open(d + "/bad.fa", "w").write(">\nACGT\n")
    with pytest.raises(ValueError):
        GenomeImage.from_fasta_lib(d + "/bad.fa", lib)
    open(d + "/empty.fa", "w").close()
    assert GenomeImage.from_fasta_lib(d + "/empty.fa", lib).names == []
    open(d + "/g.fa", "w").write(">c1\nACGTacgtNN\n>c2\nGG\n")
    g = GenomeImage.from_fasta_cached(d + "/g.fa", lib=lib)          # built by the library, written next to the FASTA ...
    assert os.path.exists(d + "/g.fa.tcb200.npz") and g.get_seq("c1", 1, 10) == "ACGTacgtNN"
    h = GenomeImage.from_fasta_cached(d + "/g.fa", lib=lib)          # ... and loaded the second time
    assert np.array_equal(g.bases2, h.bases2) and h.get_seq("c2", 1, 2) == "GG"

"""get_sjs_from_gtf against the output of the unmodified reference script
(tests/golden/gtf_sjs.json.gz, made by oracle/make_golden_gtf.py)."""
import gzip
import json
import os
import tempfile

import pytest

from transcriptclean_b200 import get_sjs_from_gtf as G
from transcriptclean_b200.genome import GenomeImage

HERE = os.path.dirname(os.path.abspath(__file__))
REC = json.load(gzip.open(os.path.join(HERE, "golden", "gtf_sjs.json.gz"), "rt"))


@pytest.mark.parametrize("min_intron", [21, 13])
def test_rows_equal_the_reference_script(min_intron):
    gi = GenomeImage.from_dict(REC["genome"])
    rows = list(G.junction_rows(REC["gtf_lines"], gi, min_intron))
    want = REC["expected"][str(min_intron)]
    assert "".join(r + "\n" for r in rows) == want
    assert len({r.split("\t")[4] for r in rows}) >= 5          # several motif codes occur


def test_command_line(tmp_path):
    d = str(tmp_path)
    with open(d + "/g.fa", "w") as f:
        for name, s in REC["genome"].items():
            f.write(">" + name + " x\n" + "\n".join(s[i:i + 60] for i in range(0, len(s), 60)) + "\n")
    with open(d + "/a.gtf", "w") as f:
        f.write("\n".join(REC["gtf_lines"]) + "\n")
    G.main(["--f", d + "/a.gtf", "--g", d + "/g.fa", "--o", d + "/sj.tab"])
    assert open(d + "/sj.tab").read() == REC["expected"]["21"]


def test_genome_image_get_seq_semantics():
    gi = GenomeImage.from_dict({"c": "ACGTNnacgtRYACGT"})
    assert gi.get_seq("c", 1, 4) == "ACGT" and gi.get_seq("c", 5, 12) == "NnacgtRY"
    assert gi.get_seq("c", 15, 99) == "GT" and gi.get_seq("c", 17, 20) == ""        # truncated / empty past the end
    with pytest.raises(IndexError):
        gi.get_seq("c", 0, 3)
    with pytest.raises(KeyError):
        gi.get_seq("nope", 1, 3)


def _device_rows(engine):
    gi = GenomeImage.from_dict(REC["genome"])
    engine.load_genome(gi)
    for min_intron in (21, 13):
        rows = list(G.junction_rows_device(REC["gtf_lines"], gi, engine, min_intron))
        assert "".join(r + "\n" for r in rows) == REC["expected"][str(min_intron)]
    # the motif fetch against the host image on random introns, contig ends and coordinates below 1 included
    import numpy as np
    rng = np.random.default_rng(3)
    name = gi.names[0]
    L = int(gi.chrom_len[0])
    st = np.concatenate((rng.integers(1, L - 5, 500), [0, 1, 1, L - 1, L, L - 2])).astype(np.int32)
    en = np.concatenate((st[:500] + rng.integers(2, 400, 500), [50, 2, 1, L + 5, L + 9, L])).astype(np.int32)
    codes = engine.intron_motifs(np.zeros(len(st), dtype=np.int32), st, en)
    for s, e, c in zip(st.tolist(), en.tolist(), codes.tolist()):
        if s < 1 or e - 1 < 1:
            assert c == -1, (s, e, c)
            continue
        assert str(20 + c) == G.intron_motif(gi, name, s, e), (s, e, c)


def test_device_motif_fetch_on_the_cpu_twin():
    import hostsim_engine
    eng = hostsim_engine.HostSimEngine()
    try:
        _device_rows(eng)
    finally:
        eng.close()


@pytest.mark.gpu
def test_device_motif_fetch_on_the_gpu():
    from transcriptclean_b200.engine import Engine
    eng = Engine(0)
    try:
        _device_rows(eng)
    finally:
        eng.close()

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

"""Loaders for tests/golden/*.json.gz (written by oracle/make_golden.py from the reference)."""
import gzip
import json
import os

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class SparseSeq:
    """A contig given as [start, bases] windows over an all-'N' background (oracle side)."""

    def __init__(self, length, windows):
        self.length = length
        self.windows = sorted((int(s), b) for s, b in windows)

    def fetch(self, start, end):
        end = min(end, self.length)
        if end < start:
            return ""
        out = ["N"] * (end - start + 1)
        for s, b in self.windows:
            e = s + len(b) - 1
            if e < start or s > end:
                continue
            lo, hi = max(s, start), min(e, end)
            out[lo - start:hi - start + 1] = b[lo - s:hi - s + 1]
        return "".join(out)


def load_fixtures():
    with gzip.open(os.path.join(GOLDEN_DIR, "fixtures.json.gz"), "rt") as f:
        return json.load(f)


def load_synth(seed):
    with gzip.open(os.path.join(GOLDEN_DIR, "synth_%d.json.gz" % seed), "rt") as f:
        return json.load(f)


SYNTH_SEEDS = (101, 102, 103, 104)


def write_rows(rows, path, header=None):
    with open(path, "w") as f:
        if header:
            f.write(header)
        for r in rows:
            f.write("\t".join(str(x) for x in r) + "\n")
    return path

"""Golden vectors for introns that a junction rescue shrinks to zero or below: the four output
texts of the UNMODIFIED reference (oracle/ref_harness.py) on tests/synth.make_micro_intron_case.

TEST INFRASTRUCTURE ONLY; runs in the build container (needs /root/reference).  Usage:
    python -m oracle.make_golden_micro        -> tests/golden/micro_introns.json.gz
"""
import gzip
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_harness as rh   # noqa: E402
import pipeline                        # noqa: E402
import synth                           # noqa: E402

SEEDS = (7005, 7047, 7083, 7091, 7101, 7102)
OPTION_SETS = (dict(), dict(maxSJOffset=10), dict(maxSJOffset=10, correctIndels=False))


def main():
    recs = []
    for seed in SEEDS:
        case = synth.make_micro_intron_case(seed)
        sjf, _ = pipeline.write_tables(case.sj_rows, None)
        runs = [{"opts": o, "expected": rh.run_reference(case.lines, case.genome, sj_file=sjf, **o)} for o in OPTION_SETS]
        recs.append({"seed": seed, "runs": runs})
    with gzip.open(os.path.join(ROOT, "tests", "golden", "micro_introns.json.gz"), "wt") as f:
        json.dump(recs, f)
    neg = sum(r["expected"]["sam"].count("-") for rec in recs for r in rec["runs"])
    print("wrote %d cases; '-' characters in the reference's SAM texts: %d" % (len(recs), neg))


if __name__ == "__main__":
    main()

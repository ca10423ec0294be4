"""Introns that a junction rescue shrinks to length zero or below (tc_plan.cuh `intron_text_rules`).
The reference emits these reads - a "0N" op, or a "-3N" whose minus sign then sticks to the op in front of it when the
CIGAR text is split again - or leaves the junction uncorrected ("Other").  Pinned by texts the UNMODIFIED reference
produced (tests/golden/micro_introns.json.gz, oracle/make_golden_micro.py)."""
import gzip
import json
import os

import pytest

import goldens
import pipeline
import synth
from transcriptclean_b200.genome import GenomeImage

with gzip.open(os.path.join(goldens.GOLDEN_DIR, "micro_introns.json.gz"), "rt") as f:
    RECS = json.load(f)


def _check(engine, rec):
    case = synth.make_micro_intron_case(rec["seed"])
    sjf, _ = pipeline.write_tables(case.sj_rows, None)
    gi = GenomeImage.from_dict(case.genome)
    for run in rec["runs"]:
        want = run["expected"]
        d = pipeline.first_diff(pipeline.oracle_texts(case.genome, case.lines, sjf, None, run["opts"]), want)
        assert d is None, ("oracle", rec["seed"], run["opts"], d)
        if engine is not None:
            d = pipeline.first_diff(pipeline.product_texts(engine, gi, case.lines, sjf, None, run["opts"]), want)
            assert d is None, ("engine", rec["seed"], run["opts"], d)


def test_goldens_hold_the_cases():
    cig = [l.split("\t")[5] for rec in RECS for run in rec["runs"] for l in run["expected"]["sam"].splitlines()]
    assert sum("-" in c for c in cig) >= 3 and sum("M0N" in c or "D0N" in c or "I0N" in c for c in cig) >= 3
    assert sum(run["expected"]["te"].count("Other") for rec in RECS for run in rec["runs"]) >= 20


@pytest.mark.parametrize("rec", RECS, ids=[str(r["seed"]) for r in RECS])
def test_oracle_and_host_twin_reproduce_the_reference(rec):
    import hostsim_engine
    eng = hostsim_engine.HostSimEngine()
    try:
        _check(eng, rec)
    finally:
        eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("rec", RECS, ids=[str(r["seed"]) for r in RECS])
def test_gpu_reproduces_the_reference(rec):
    from transcriptclean_b200.engine import Engine
    eng = Engine(0)
    try:
        _check(eng, rec)
    finally:
        eng.close()


@pytest.mark.gpu
def test_gpu_more_seeds_vs_oracle():
    from transcriptclean_b200.engine import Engine
    eng = Engine(0)
    try:
        for seed in range(7300, 7340):
            case = synth.make_micro_intron_case(seed)
            sjf, _ = pipeline.write_tables(case.sj_rows, None)
            gi = GenomeImage.from_dict(case.genome)
            for opts in (dict(), dict(maxSJOffset=10, correctIndels=False)):
                d = pipeline.first_diff(pipeline.product_texts(eng, gi, case.lines, sjf, None, opts),
                                        pipeline.oracle_texts(case.genome, case.lines, sjf, None, opts))
                assert d is None, (seed, opts, d)
    finally:
        eng.close()

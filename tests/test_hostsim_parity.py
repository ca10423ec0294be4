"""CPU-only checks of the planning code (csrc/tc_plan.cuh, shared verbatim with the kernels)
and of the host-side parser / table builders / renderer, through the hostsim build of the ABI.
The GPU parity tests proper are in test_gpu_parity.py."""
import pytest

import goldens
import parity_cases as pc
from hostsim_engine import HostSimEngine

FIX = goldens.load_fixtures()


@pytest.fixture(scope="module")
def engine():
    e = HostSimEngine()
    yield e
    e.close()


@pytest.mark.parametrize("rec", FIX, ids=[r["name"] for r in FIX])
def test_reference_fixture(engine, rec):
    pc.check_fixture(engine, rec)


@pytest.mark.parametrize("seed", goldens.SYNTH_SEEDS)
def test_reference_synthetic(engine, seed):
    pc.check_synth_golden(engine, seed)


@pytest.mark.parametrize("seed", range(200, 212))
def test_generated_vs_oracle(engine, seed):
    pc.check_generated(engine, seed, unpinned=(seed % 3 == 0))


def test_seq_forms_agree(engine):
    assert sum(bool(pc.check_seq_forms(engine, seed)) for seed in (231, 232, 1232)) >= 1


def test_fully_tagged_input(engine):
    """Every read carries NM and MD: the dry run then does not upload the SEQ column at all."""
    pc.check_generated(engine, 1204, n_reads=200, tagless_frac=0.0)


def test_long_cigars(engine):
    """15 % errors: CIGARs of well over 64 ops, where plan_read also leaves the intron coordinates."""
    pc.check_generated(engine, 5003, n_reads=60, exon_len=(200, 900), err=0.15)

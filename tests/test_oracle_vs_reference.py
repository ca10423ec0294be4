"""Live differential pin of the oracle: oracle/tc_oracle.py against the UNMODIFIED reference run in-process
under oracle/stubs (oracle/ref_harness.py), on freshly generated seeded cases.  Needs /root/reference, i.e.
runs in the build container only (auto-skipped elsewhere; the committed golden vectors cover the rest)."""
import pytest

import pipeline
import synth
from oracle import ref_harness

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason="/root/reference is not present on this machine")

OPTION_SETS = (
    dict(),
    dict(canonOnly=True, maxSJOffset=8, maxLenIndel=3),
    dict(primaryOnly=True, correctMismatches=False),
    dict(correctIndels=False, correctSJs=False),
)


@pytest.mark.parametrize("seed", list(range(9101, 9111)))
def test_oracle_equals_reference(seed):
    case = synth.make_case(seed, n_reads=80, unpinned=False)
    sjf, vcf = pipeline.write_tables(case.sj_rows, case.vcf_rows)
    for k, opts in enumerate(OPTION_SETS):
        use_vcf = vcf if k % 2 == 0 else None
        want = ref_harness.run_reference(case.lines, case.genome, sj_file=sjf, vcf_file=use_vcf, **opts)
        got = pipeline.oracle_texts(case.genome, case.lines, sjf, use_vcf, opts)
        d = pipeline.first_diff(got, want)
        assert d is None, (seed, opts, d)


def test_oracle_dry_run_equals_reference():
    case = synth.make_case(9104, n_reads=60, quirks=False)
    want = ref_harness.run_reference(case.lines, case.genome, dryRun=True)
    got = pipeline.oracle_texts(case.genome, case.lines, None, None, {}, dry=True)
    assert {k: got[k] for k in ("log", "te")} == {k: want[k] for k in ("log", "te")}


@pytest.mark.parametrize("seed", (7201, 7202, 7203))
def test_oracle_equals_reference_on_micro_introns(seed):
    """Rescues that shrink an intron to zero or below (tests/test_micro_introns.py)."""
    case = synth.make_micro_intron_case(seed)
    sjf, _ = pipeline.write_tables(case.sj_rows, None)
    for opts in (dict(), dict(maxSJOffset=10, correctIndels=False)):
        want = ref_harness.run_reference(case.lines, case.genome, sj_file=sjf, **opts)
        d = pipeline.first_diff(pipeline.oracle_texts(case.genome, case.lines, sjf, None, opts), want)
        assert d is None, (seed, opts, d)

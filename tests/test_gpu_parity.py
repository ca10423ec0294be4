"""GPU parity: the CUDA path (through the C ABI) against outputs of the unmodified reference
(tests/golden) and against the oracle on seeded cases.  Bit-exact on all four output texts."""
import pytest

import goldens
import parity_cases as pc

pytestmark = pytest.mark.gpu
FIX = goldens.load_fixtures()


@pytest.fixture(scope="module")
def engine():
    from transcriptclean_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("rec", FIX, ids=[r["name"] for r in FIX])
def test_reference_fixture(engine, rec):
    pc.check_fixture(engine, rec)


@pytest.mark.parametrize("seed", goldens.SYNTH_SEEDS)
def test_reference_synthetic(engine, seed):
    pc.check_synth_golden(engine, seed)


@pytest.mark.parametrize("seed", range(300, 330))
def test_generated_vs_oracle(engine, seed):
    pc.check_generated(engine, seed, unpinned=(seed % 3 == 0))


def test_long_reads_and_many(engine):
    pc.check_generated(engine, 999, n_reads=1500, err=0.08, shift_frac=0.4)


def test_reads_longer_than_the_shared_memory_stage(engine):
    """Reads of 5-20 kb (new SEQ > EMIT_CAP) and reads with hundreds of slices take the generic
    paths of the output kernel."""
    pc.check_generated(engine, 1201, n_reads=120, chrom_len=400000, genes_per_chrom=10, exon_len=(900, 3000), err=0.03)
    pc.check_generated(engine, 1202, n_reads=150, err=0.12, shift_frac=0.5)


def test_tiny_reads_and_heavy_errors(engine):
    """Reads of a few dozen bases (every row alignment of the packed / unpacked SEQ buffers, chunks that hold
    several M runs) and reads with 15 % errors (many slices, deletion tokens in the clean-read MD)."""
    pc.check_generated(engine, 5001, n_reads=300, exon_len=(12, 40), err=0.05)
    pc.check_generated(engine, 5002, n_reads=300, exon_len=(12, 25), err=0.10, shift_frac=0.5)
    pc.check_generated(engine, 5003, n_reads=60, exon_len=(200, 900), err=0.15)


def test_fully_tagged_input(engine):
    """Every read carries NM and MD: the dry run then does not upload the SEQ column at all."""
    pc.check_generated(engine, 1204, n_reads=300, tagless_frac=0.0)


def test_tagless_input_at_scale(engine):
    """No NM/MD/jM/jI tags at all: NM/MD are bootstrapped on the device for every read (tr.py:47-48)."""
    pc.check_generated(engine, 1203, n_reads=400, tagless_frac=1.0)


def test_seq_forms_agree(engine):
    """4-bit and ASCII SEQ transport give the same texts; at least one of the cases really travels packed."""
    assert sum(bool(pc.check_seq_forms(engine, seed, n_reads=300)) for seed in (1231, 1232, 1233)) >= 1


def test_seq_forms_agree_when_pipelined(engine):
    engine.set_pipeline(41)
    try:
        pc.check_seq_forms(engine, 1234, n_reads=300)
    finally:
        engine.set_pipeline(131072)


def test_pipelined_chunks(engine):
    """tc_correct_batch cut into small chunks (3-stream pipeline, rebased offsets) gives the same texts."""
    for chunk in (16, 37, 100):
        engine.set_pipeline(chunk)
        try:
            pc.check_generated(engine, 777 + chunk, n_reads=260, unpinned=False)
        finally:
            engine.set_pipeline(131072)


def test_empty_batch(engine):
    import numpy as np
    from transcriptclean_b200.genome import GenomeImage
    from transcriptclean_b200 import api
    refs = api.make_refs(GenomeImage.from_dict({"chr1": "ACGT" * 100}), engine=engine)
    out = api.correct_batch_text([], api.Struct(), refs)
    assert out == {"sam": "", "fa": "", "log": "", "te": ""}


def test_plane_path_and_ascii_path_agree(engine):
    """k_verify (NM / MD on the 2-bit plane of the input SEQ) against k_emit deciding every read (TC_NO_PLANE=1), every SEQ form;
    with variants in play (reads that keep a mismatch go to the ASCII path) and with heavy errors."""
    assert pc.check_plane_and_ascii_paths(engine, 4101, n_reads=300) > 0
    assert pc.check_plane_and_ascii_paths(engine, 4102, n_reads=200, err=0.10, shift_frac=0.5) > 0
    pc.check_plane_and_ascii_paths(engine, 4103, n_reads=200, exon_len=(12, 40), err=0.05)

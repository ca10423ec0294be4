"""Host-side logic on CPU: the `transcriptclean` option surface, file outputs, and the sharded
(N>1) paths - threads and a world_size-2 gloo group - through the hostsim engine."""
import os
import sys
import tempfile

import pytest

import pipeline
import synth
from hostsim_engine import HostSimEngine
from transcriptclean_b200 import cli, shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _case_files(seed, n_reads=120):
    case = synth.make_case(seed, n_reads=n_reads)
    d = tempfile.mkdtemp(prefix="tc_cli_")
    case.write_fasta(d + "/g.fa"); case.write_sam(d + "/in.sam"); case.write_sj(d + "/sj.tab"); case.write_vcf(d + "/v.vcf")
    return case, d


def test_shard_bounds_match_split_input():
    assert shard.shard_bounds(10, 3) == [(0, 4), (4, 8), (8, 10)]
    assert shard.shard_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert shard.shard_bounds(0, 2) == [(0, 0), (0, 0)]


def test_option_surface_and_prefix_abbreviations():
    d = tempfile.mkdtemp()
    o = cli.get_options(["--sam", "x.sam", "--g", "g.fa", "-j", "sj", "--maxLenIndel", "3", "--correctSJs", "False",
                         "--primaryOnly", "--canonOnly", "--o", d + "/out/TC"])
    assert (o.refGenome, o.sjAnnotFile, o.maxLenIndel, o.correctSJs, o.primaryOnly, o.canonOnly) == \
        ("g.fa", "sj", 3, "false", True, True)
    assert o.tmp_dir == d + "/out/TC_tmp/"
    with pytest.raises(RuntimeError):
        cli.get_options(["--correctIndels", "maybe"])
    assert cli.get_options(["--o", d]).outprefix == d + "/TC"          # a directory gets the default prefix


@pytest.mark.parametrize("n_engines", [1, 3])
def test_cli_end_to_end_files(n_engines):
    case, d = _case_files(31)
    o = cli.get_options(["-s", d + "/in.sam", "-g", d + "/g.fa", "-j", d + "/sj.tab", "-v", d + "/v.vcf",
                         "-o", d + "/TC", "--deleteTmp"])
    engines = [HostSimEngine() for _ in range(n_engines)]
    cli.run(o, engines=engines)
    want = pipeline.oracle_texts(case.genome, case.lines, d + "/sj.tab", d + "/v.vcf", {})
    sam = open(d + "/TC_clean.sam").read()
    assert sam.startswith("@HD") and sam.split("\n", 1)[0] == "@HD\tVN:1.0\tSO:unsorted"
    body = "".join(l + "\n" for l in sam.splitlines() if not l.startswith("@"))
    assert body == want["sam"]
    assert open(d + "/TC_clean.fa").read() == want["fa"]
    assert open(d + "/TC_clean.log").read() == cli.LOG_HEADER + want["log"]
    assert open(d + "/TC_clean.TE.log").read() == cli.TE_HEADER + want["te"]
    assert not os.path.exists(d + "/TC_tmp")


@pytest.mark.parametrize("n_engines", [1, 2])
def test_cli_streams_the_sam_file_in_blocks(n_engines, monkeypatch):
    """Blocks of a few kB (dozens per file, header lines only in the first): the same four files."""
    case, d = _case_files(34, n_reads=150)
    monkeypatch.setenv("TC_CLI_BLOCK_BYTES", "6000")
    assert len(list(cli.sam_blocks(d + "/in.sam", 6000))) > 10
    assert b"".join(cli.sam_blocks(d + "/in.sam", 6000)) == open(d + "/in.sam", "rb").read()
    o = cli.get_options(["-s", d + "/in.sam", "-g", d + "/g.fa", "-j", d + "/sj.tab", "-v", d + "/v.vcf", "-o", d + "/TC"])
    cli.run(o, engines=[HostSimEngine() for _ in range(n_engines)])
    want = pipeline.oracle_texts(case.genome, case.lines, d + "/sj.tab", d + "/v.vcf", {})
    sam = open(d + "/TC_clean.sam").read()
    assert sam.startswith("@HD")
    assert "".join(l + "\n" for l in sam.splitlines() if not l.startswith("@")) == want["sam"]
    assert open(d + "/TC_clean.fa").read() == want["fa"]
    assert open(d + "/TC_clean.log").read() == cli.LOG_HEADER + want["log"]
    assert open(d + "/TC_clean.TE.log").read() == cli.TE_HEADER + want["te"]


def test_cli_dry_run_writes_only_logs():
    case = synth.make_case(32, n_reads=80, quirks=False)
    d = tempfile.mkdtemp(prefix="tc_cli_")
    case.write_fasta(d + "/g.fa"); case.write_sam(d + "/in.sam")
    o = cli.get_options(["-s", d + "/in.sam", "-g", d + "/g.fa", "--dryRun", "-o", d + "/TC"])
    cli.run(o, engines=[HostSimEngine()])
    want = pipeline.oracle_texts(case.genome, case.lines, None, None, {}, dry=True)
    assert open(d + "/TC_clean.log").read() == cli.LOG_HEADER + want["log"]
    assert open(d + "/TC_clean.TE.log").read() == cli.TE_HEADER + want["te"]
    assert not os.path.exists(d + "/TC_clean.sam") and not os.path.exists(d + "/TC_clean.fa")


def test_unknown_chromosome_is_reported():
    case, d = _case_files(33, 20)
    with open(d + "/in.sam", "a") as f:
        f.write("x\t0\tchrNope\t1\t60\t4M\t*\t0\t0\tACGT\t*\n")
    o = cli.get_options(["-s", d + "/in.sam", "-g", d + "/g.fa", "-o", d + "/TC"])
    with pytest.raises(RuntimeError, match="not found in the fasta reference"):
        cli.run(o, engines=[HostSimEngine()])


_WORKER = r"""
import os, sys, json
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import torch.distributed as dist
import synth, pipeline
from hostsim_engine import HostSimEngine
from transcriptclean_b200 import api, shard
from transcriptclean_b200.genome import GenomeImage
from transcriptclean_b200.tables import build_sj_tables
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
case = synth.make_case(55, n_reads=90)
sjf, _ = pipeline.write_tables(case.sj_rows, None)
gi = GenomeImage.from_dict(case.genome)
chroms = set(l.split("\t")[2] for l in case.lines)
refs = api.make_refs(gi, build_sj_tables(sjf, chroms, gi.index), None, engine=HostSimEngine())
out = shard.correct_sharded_dist(case.lines, api.Struct(), refs, api.correct_batch_text, dist)
if dist.get_rank() == 0:
    want = pipeline.oracle_texts(case.genome, case.lines, sjf, None, {})
    assert out == want, pipeline.first_diff(out, want)
    print("RANK0_OK")
dist.barrier()
dist.destroy_process_group()
"""


def test_world_size_2_gloo_shards_concatenate_in_order():
    import socket
    import subprocess
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    code = _WORKER % {"root": ROOT, "port": port}
    procs = [subprocess.Popen([sys.executable, "-c", code, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=300)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "RANK0_OK" in outs[0]


def test_genome_image_cache_round_trip():
    """The packed genome image saved next to the FASTA loads back identical (planes, literal runs, contig table)
    and is rebuilt when it was not made from this very FASTA file."""
    import numpy as np
    from transcriptclean_b200.genome import GenomeImage
    case, d = _case_files(35, 10)
    a = GenomeImage.from_fasta_cached(d + "/g.fa")
    assert os.path.exists(d + "/g.fa.tcb200.npz")
    b = GenomeImage.from_fasta_cached(d + "/g.fa")          # from the image
    assert a.names == b.names and np.array_equal(a.chrom_len, b.chrom_len) and np.array_equal(a.chrom_off, b.chrom_off)
    assert np.array_equal(a.bases2, b.bases2) and np.array_equal(a.lower1, b.lower1)
    assert all(np.array_equal(x, y) for x, y in zip(a.exceptions(), b.exceptions()))
    assert b.source == GenomeImage.fasta_fingerprint(d + "/g.fa")
    # another genome copied over the same path with an OLDER time stamp (cp -p, rsync -t): the image is not trusted
    other = synth.make_case(36, n_reads=5)
    st = os.stat(d + "/g.fa")
    other.write_fasta(d + "/g.fa")
    os.utime(d + "/g.fa", ns=(st.st_atime_ns - 10 ** 12, st.st_mtime_ns - 10 ** 12))
    c = GenomeImage.from_fasta_cached(d + "/g.fa")
    fresh = GenomeImage.from_fasta(d + "/g.fa")
    assert np.array_equal(c.bases2, fresh.bases2) and not np.array_equal(c.bases2, a.bases2)
    assert GenomeImage.load(d + "/g.fa.tcb200.npz").source == GenomeImage.fasta_fingerprint(d + "/g.fa")


def test_unused_gpus_are_hidden_before_cuda_starts(monkeypatch):
    """cli._hide_unused_gpus narrows CUDA_VISIBLE_DEVICES to the first `want` devices of the current list (or of /dev/nvidia<N>)
    and never widens it."""
    monkeypatch.setenv("CUDA_VISIBLE_DEVICES", "3,5,6,7")
    cli._hide_unused_gpus(2)
    assert os.environ["CUDA_VISIBLE_DEVICES"] == "3,5"
    cli._hide_unused_gpus(8)
    assert os.environ["CUDA_VISIBLE_DEVICES"] == "3,5"
    monkeypatch.delenv("CUDA_VISIBLE_DEVICES")
    monkeypatch.setattr(os, "listdir", lambda p: ["nvidia0", "nvidia1", "nvidia10", "nvidia2", "nvidiactl", "nvidia-uvm", "null"])
    cli._hide_unused_gpus(3)
    assert os.environ["CUDA_VISIBLE_DEVICES"] == "0,1,2"
    monkeypatch.delenv("CUDA_VISIBLE_DEVICES")
    monkeypatch.setattr(os, "listdir", lambda p: ["null", "zero"])
    cli._hide_unused_gpus(1)
    assert "CUDA_VISIBLE_DEVICES" not in os.environ


def test_file_writers_write_every_piece_in_order_and_surface_errors():
    d = tempfile.mkdtemp(prefix="tc_wr_")
    files = {k: open(d + "/" + k, "wb") for k in ("sam", "fa", "log", "te")}
    w = cli._FileWriters(files)
    for blk in range(5):
        w.write({k: [memoryview(("%s%d-%d|" % (k, blk, p)).encode()) for p in range(3)] for k in files})
    w.close()
    for f in files.values():
        f.close()
    assert open(d + "/fa").read() == "".join("fa%d-%d|" % (b, p) for b in range(5) for p in range(3))
    closed = {k: open(d + "/" + k + "2", "wb") for k in ("log", "te")}
    w = cli._FileWriters(closed)
    closed["te"].close()
    with pytest.raises(ValueError):
        w.write({"log": [b"x"], "te": [b"y"]})
    w.close()


def test_sam_blocks_are_slices_of_a_mapping_and_cover_the_file():
    d = tempfile.mkdtemp(prefix="tc_blk_")
    lines = [("read%d\t0\tchr1\t%d\t60\t10M\t*\t0\t0\tACGTACGTAC\t*" % (k, k + 1)).encode() for k in range(500)]
    for tail in (b"\n", b""):                             # with and without a newline at the end of the file
        open(d + "/x.sam", "wb").write(b"\n".join(lines) + tail)
        blocks = list(cli.sam_blocks(d + "/x.sam", 700))
        assert len(blocks) > 20 and all(isinstance(b, memoryview) for b in blocks)
        assert b"".join(blocks) == b"\n".join(lines) + tail
        assert all(bytes(b).endswith(b"\n") for b in blocks[:-1])
    open(d + "/long.sam", "wb").write(b"A" * 5000 + b"\n" + b"B" * 10 + b"\n")      # a line longer than a block
    assert [len(b) for b in cli.sam_blocks(d + "/long.sam", 1000)] == [5001, 11]
    open(d + "/empty.sam", "wb").close()
    assert list(cli.sam_blocks(d + "/empty.sam", 1000)) == []


def test_vcf_chromosomes_from_the_library_equal_the_python_scan():
    import gzip
    case, d = _case_files(41)
    lib = HostSimEngine()._lib
    rows = open(d + "/v.vcf").read()
    extra = "##meta\n#CHROM\tPOS\n\n   \nchrQ\t5\t.\tA\tT\nchrQ\t9\t.\tA\tT\nnotab\nchr1\t7\t.\tC\tG\nlast_without_newline"
    open(d + "/w.vcf", "w").write(rows + extra)
    with gzip.open(d + "/w.vcf.gz", "wt") as f:
        f.write(rows + extra)
    for path in (d + "/w.vcf", d + "/w.vcf.gz"):
        got, want = cli.vcf_chrom_set(path, lib), cli.vcf_chrom_set(path)
        assert got == want and {"chrQ", "chr1", "notab\n", "last_without_newline"} <= got
    genome = type("G", (), {"keys": lambda self: ["chr1", "chr2", "chr3", "chrQ"]})()
    cli.validate_chroms(genome, d + "/w.vcf", {"chr1"}, lib=lib)
    with pytest.raises(RuntimeError, match="None of the chromosomes included in the VCF file"):
        cli.validate_chroms(genome, d + "/w.vcf.gz", {"chr3"}, lib=lib)
    with pytest.raises(RuntimeError, match="Problem reading the provided VCF file"):
        cli.validate_chroms(genome, d + "/missing.vcf", {"chr1"}, lib=lib)

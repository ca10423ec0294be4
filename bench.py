#!/usr/bin/env python
"""bench.py - reads corrected per second on the BASELINE.json workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the correction path (TranscriptClean.py:406-460 and callees) over one
batch of synthetic reads.  N=1 workload = BASELINE.json configs[1]: synthetic PacBio Iso-seq
reads (~2 kb, ~1 % error, 6 exons) against a synthetic 248,956,422-bp "chr1" with a STAR-format
SJ table (tools/synth_gen.c; the real hg38 / GM12878 inputs are not shipped).  With N>1 every
rank corrects its own shard of the same shape (weak scaling, no data-path collective).

  value  device-resident throughput: inputs already in HBM, tc_batch_run() only
  e2e    through the C ABI with pinned host buffers: tc_correct_batch() = H2D + kernels + D2H, pipelined in chunks;
         the SEQ column travels as 2-bit codes + the listed bases outside ACGT (--seq4: 4-bit codes, --seq8: ASCII),
         the CIGAR as 16-bit words + the listed long ops (--cig40: op + length arrays)
  roofline  the dominant kernel's algorithmic bytes / its CUDA-event time vs measured HBM peak
            (traffic: DRAM bytes of that kernel from the committed ncu --set full capture); roofline_step: the whole step
            against SURVEY.md 8(d)'s bytes per read
  dry_run   the --dryRun inventory (tc_dryrun_batch) of the same batch, device-resident and end to end
  sam_text_path  SAM text in -> the four output texts (C++ host parse / render around the GPU), N=1 only
  seam_batch_correct  the reference's own seam as a maintainer calls it: api.batch_correct(list of line strings, ..., outfiles), N=1 only
  cpu_baseline  the oracle port of the reference (oracle/tc_oracle.py) on a bounded sample, one core
  --impl reference  the unmodified reference (pip-installed into oracle/_ref by build(), run under oracle/stubs) on all host
                    cores; the oracle port where that install is missing
  --shape ont       BASELINE.json configs[3] shape;  --resident-only  profiling aid (see profiles/README.md)
"""
import argparse
import contextlib
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOAD = "configs[1]: synthetic PacBio Iso-seq reads (~2 kb, ~1% error, 6 exons) vs synthetic 248,956,422-bp chr1 + STAR SJ table"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    def __init__(self, device):
        self.path = tempfile.mktemp(prefix="tc_clocks_", suffix=".csv")
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


def oracle_setup(w, sj_path):
    from oracle import tc_oracle as oc
    genome = oc.Genome({w.chrom: w.genome.tobytes().decode("ascii")})
    sj = oc.SjTables.from_file(sj_path, {w.chrom})
    return oc, genome, sj


_ORACLE = {}


def _oracle_chunk(lines):
    oc, genome, sj = _ORACLE["ctx"]
    out = oc.correct_lines(lines, oc.Options(), genome, sj, None)
    if _ORACLE.get("keep"):
        _ORACLE["texts"] = out           # the parity check of the timed batch compares these with the GPU's texts
    return len(out["log"])


def parity_check(w, gi, cols, nm, res, n_check, sj_path, texts=None):
    """The GPU results of the first n_check reads of the TIMED batch, rendered to the four output texts,
    against the oracle's texts for the same reads.  Returns None, or a description of the first difference."""
    from transcriptclean_b200 import render, sam
    lines = w.sam_lines(cols, nm, 0, n_check)
    if texts is None:
        oc, genome, sj = _ORACLE.get("ctx") or oracle_setup(w, sj_path)
        texts = oc.correct_lines(lines, oc.Options(), genome, sj, None)
    got = render.render_batch(sam.parse_lines(lines, gi), res, False, False)
    for k in ("sam", "fa", "log", "te"):
        if got[k] != texts[k]:
            a, b = got[k].split("\n"), texts[k].split("\n")
            for i, (x, y) in enumerate(zip(a, b)):
                if x != y:
                    return "%s line %d: got %r want %r" % (k, i, x[:300], y[:300])
            return "%s: %d vs %d lines" % (k, len(a), len(b))
    return None


def oracle_rate(w, sj_path, lines, threads, steps=1, warmup=0):
    """Reads/s of the oracle port on `lines` with `threads` forked workers."""
    import multiprocessing as mp
    _ORACLE["ctx"] = oracle_setup(w, sj_path)
    _ORACLE["ctx"][0].FAITHFUL_PER_BASE = True     # per-base genome walk like transcript.py:296-318
    chunks = [lines[i::threads] for i in range(threads)]
    times = []
    if threads == 1:
        for s in range(warmup + steps):
            t0 = time.perf_counter(); _oracle_chunk(lines); t1 = time.perf_counter()
            if s >= warmup:
                times.append(t1 - t0)
    else:
        ctx = mp.get_context("fork")
        with ctx.Pool(threads) as pool:
            for s in range(warmup + steps):
                t0 = time.perf_counter(); pool.map(_oracle_chunk, chunks); t1 = time.perf_counter()
                if s >= warmup:
                    times.append(t1 - t0)
    return len(lines) / (sum(times) / len(times)), sum(times) / len(times)


def write_sj(w):
    path = tempfile.mktemp(prefix="tc_sj_", suffix=".tab")
    with open(path, "w") as f:
        for r in w.sj_rows():
            f.write("\t".join(str(x) for x in r) + "\n")
    return path


_REF = {}


def _ref_init(w_genome, chrom, sj_path, chrom_line):
    """Pool initializer: one reference worker per process, as run_chunk does (prep_refs per worker, TC.py:560-586)."""
    from oracle import ref_harness
    _REF["w"] = ref_harness.Worker({chrom: w_genome}, sj_file=sj_path, chrom_lines=[chrom_line])


def _ref_chunk(lines):
    return len(_REF["w"].correct(lines)["log"])


def reference_rate(w, sj_path, lines, threads, steps=1, warmup=0):
    """Reads/s of the UNMODIFIED reference (oracle/_ref or /root/reference, under oracle/stubs): `threads` forked workers,
    each with its own prep_refs; the timed region is the batch_correct calls (TC.py:350-403), like the GPU arm's
    timed region excludes table set-up."""
    import multiprocessing as mp
    genome = w.genome.tobytes().decode("ascii")          # shared with the forked workers (copy-on-write pages)
    chunks = [lines[i::threads] for i in range(threads)]
    times = []
    ctx = mp.get_context("fork")
    with ctx.Pool(threads, initializer=_ref_init, initargs=(genome, w.chrom, sj_path, lines[0])) as pool:
        pool.map(_ref_chunk, [c[:2] for c in chunks])    # every worker is up (prep_refs done) before the clock starts
        for s in range(warmup + steps):
            t0 = time.perf_counter(); pool.map(_ref_chunk, chunks, chunksize=1); t1 = time.perf_counter()
            if s >= warmup:
                times.append(t1 - t0)
    return len(lines) / (sum(times) / len(times)), sum(times) / len(times)


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores, on a bounded sample of
    the same workload.  The UNMODIFIED reference (pip-installed into oracle/_ref by __graft_entry__.build(), run under
    oracle/stubs) when it is there, else the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tools.workload import Workload
    from oracle import ref_harness
    cores = os.cpu_count() or 1
    threads = min(cores, 64)
    w = Workload(seed=args.seed, chrom_len=args.chrom_len, n_genes=args.genes)
    sj_path = write_sj(w)
    real = ref_harness.available() and not args.ref_port
    n_sample = args.ref_reads if args.ref_reads else threads * (300 if real else 600)
    cols, nm = w.reads(args.shape, n_sample)
    lines = w.sam_lines(cols, nm, 0, n_sample)
    if real:
        rate, sec = reference_rate(w, sj_path, lines, threads, steps=args.steps, warmup=min(args.warmup, 1))
        kind, what = "reference", "the unmodified reference (TranscriptClean v2.1 installed in oracle/_ref, prep_refs + batch_correct under oracle/stubs)"
    else:
        rate, sec = oracle_rate(w, sj_path, lines, threads, steps=args.steps, warmup=min(args.warmup, 1))
        kind, what = "port", "oracle/tc_oracle.py with the reference's per-base genome walk"
    line = {"impl": "reference", "metric": "reads corrected/sec", "value": rate, "unit": "reads/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "shape": args.shape, "reads_per_step": n_sample},
            "cpu_baseline": {"value": rate, "unit": "reads/s", "cores": threads, "kind": kind,
                             "sample": "first %d reads of the workload per step, %d forked workers, %s" % (n_sample, threads, what)},
            "e2e": {"value": rate, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    # stdout carries exactly one JSON line: everything else (NCCL banner, library chatter) goes to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--reads", type=int, default=1000000, help="reads per GPU per step")
    ap.add_argument("--shape", default="pacbio")
    ap.add_argument("--chrom-len", type=int, default=248956422)
    ap.add_argument("--genes", type=int, default=3000)
    ap.add_argument("--seed", type=int, default=20261019)
    ap.add_argument("--cpu-sample", type=int, default=3000)
    ap.add_argument("--ref-reads", type=int, default=0)
    ap.add_argument("--ref-port", action="store_true", help="--impl reference: time the oracle port even when oracle/_ref holds the reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--text-sample", type=int, default=200000, help="reads for the SAM-text path measurement (0 = skip)")
    ap.add_argument("--text-block-mb", type=int, default=12, help="block size of the SAM-text path measurement (the CLI's TC_CLI_BLOCK_MB)")
    ap.add_argument("--text-workers", type=int, default=4, help="pipelined workers (engine contexts) of the SAM-text path measurement")
    ap.add_argument("--resident-only", action="store_true", help="profiling aid: only the device-resident steps (no pipelined e2e calls), so that a "
                    "launch list holds exactly the kernels of the timed region")
    ap.add_argument("--no-dry", action="store_true", help="skip the --dryRun measurement")
    ap.add_argument("--seq8", action="store_true", help="send SEQ as ASCII instead of a packed form")
    ap.add_argument("--cig40", action="store_true", help="send the CIGAR as op + length arrays instead of 16-bit words")
    ap.add_argument("--seq4", action="store_true", help="send SEQ as 4-bit codes instead of the 2-bit form")
    ap.add_argument("--debug-mask", type=int, default=0, help="profiling aid: skip phases of k_emit (results invalid)")
    ap.add_argument("--parity-reads", type=int, default=-1, help="reads of the timed batch whose texts are compared with the oracle's "
                    "(default: --cpu-sample at N=1, 500 per rank at N>1; 0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from tools.workload import Workload
    from transcriptclean_b200 import engine as E
    from transcriptclean_b200.tables import build_sj_tables

    steps, warmup = args.steps, max(args.warmup, 3)
    w = Workload(seed=args.seed, chrom_len=args.chrom_len, n_genes=args.genes)
    sj_path = write_sj(w)
    cols, nm = w.reads(args.shape, args.reads, seed_offset=rank)
    gi = w.genome_image()
    eng = E.Engine(local)
    eng.set_options(debug=args.debug_mask)
    eng.load_genome(gi)
    eng.load_junctions(build_sj_tables(sj_path, {w.chrom}, gi.index))

    # inputs in pinned host memory; SEQ travels in the ABI's 4-bit form (two bases per byte) unless --seq8
    hcols = cols
    if not args.seq8:
        hcols = (None if args.seq4 else eng.pack_cols2(cols)) or eng.pack_cols(cols)
        if hcols is None:
            raise SystemExit("the synthetic SEQ column holds bytes outside the engine's alphabet")
    if not args.cig40:
        hcols = eng.pack_cigar(hcols)
        if "cig16" in hcols:              # the op / length arrays stay on the host
            hcols = {k: v for k, v in hcols.items() if k not in ("cig_op", "cig_len")}
    seq_form = "ASCII" if args.seq8 else ("2-bit codes + a list of the bases outside ACGT" if hcols.get("seq_bits") == 2 else "4-bit codes")
    pinned = {}
    for k, a in hcols.items():
        if not isinstance(a, np.ndarray):
            continue
        t = torch.empty(max(a.nbytes, 8), dtype=torch.uint8, pin_memory=True)
        if a.nbytes:
            t[:a.nbytes].copy_(torch.from_numpy(a.view(np.uint8).reshape(-1)))
        pinned[k] = t
    b = E._BatchIn()
    b.n_reads = args.reads
    for k in pinned:
        setattr(b, k, pinned[k].data_ptr())
    b.seq_bits = int(hcols.get("seq_bits", 4 if "seq_poff" in hcols else 8))
    b.n_exc = len(hcols["exc_pos"]) if "exc_pos" in hcols else 0
    b.n_cig_exc = len(hcols["cig_exc_idx"]) if "cig_exc_idx" in hcols else 0
    h2d_bytes = int(sum(a.nbytes for a in hcols.values() if isinstance(a, np.ndarray)))

    stream = torch.cuda.ExternalStream(eng.stream())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.resident_only:
        eng.set_pipeline(0)
    for _ in range(1 if args.resident_only else warmup):
        eng.raw_correct(b)
    t_last = eng.timing()
    d2h_bytes = int(t_last["d2h_bytes"])

    # ---- device-resident: upload once, time K runs
    eng._check(eng._lib.tc_batch_upload(eng._ctx, C.byref(b)), "upload")
    for _ in range(warmup if args.resident_only else 1):
        eng.run()
    sampler = ClockSampler(local)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern = {k: 0.0 for k in ("k_nmmd_init_ms", "k_measure_ms", "k_plan_ms", "k_junction_ms", "k_emit_ms", "kernels_ms")}
    launches = 0
    e0.record(stream)
    for _ in range(steps):
        eng.run()
        t = eng.timing()
        for k in kern:
            kern[k] += t[k]
        launches += t["launches"]
        exact_reads = t["exact_nmmd_reads"]; ascii_reads = t["ascii_path_reads"]
    e1.record(stream)
    barrier()
    ms_run = e0.elapsed_time(e1)
    # ---- end to end through the ABI with host buffers
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    for _ in range(1 if args.resident_only else steps):
        out = eng.raw_correct(b)
    f1.record(stream)
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    clocks = sampler.stop()

    res = E.BatchResult(out, False, eng.alphabet, hcols)  # host copies of the last corrected batch (the dry run below reuses the buffers)
    # ---- --dryRun (dryRun(), TC.py:1615-1678: the positional inventory of mismatches / indels) on the same batch
    ms_dry = ms_dry_e2e = 0.0
    if not args.no_dry and not args.resident_only:
        dsteps = max(1, min(steps, 5))
        eng._check(eng._lib.tc_batch_upload(eng._ctx, C.byref(b)), "upload")
        for _ in range(2):
            eng.run(dry=True)
        barrier()
        g0, g1, g2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        g0.record(stream)
        for _ in range(dsteps):
            eng.run(dry=True)
        g1.record(stream)
        for _ in range(dsteps):
            eng.raw_dryrun(b)
        g2.record(stream)
        barrier()
        ms_dry, ms_dry_e2e = g0.elapsed_time(g1) / dsteps, g1.elapsed_time(g2) / dsteps

    tm = torch.tensor([ms_run, ms_e2e, ms_dry, ms_dry_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms_run, ms_e2e, ms_dry, ms_dry_e2e = (float(x) for x in tm)

    # ---- algorithmic bytes of the dominant kernel (DESIGN.md "Kernels and rooflines")
    st = res.status
    primary = (st & E.ST_CLASS_MASK) == 0
    refreshed = primary & ((st & E.ST_NMMD_DEVICE) != 0)
    cig_m = np.where(res.cig_op == ord("M"), res.cig_len, 0).astype(np.int64)
    m_per_read = np.add.reduceat(np.concatenate((cig_m, [0])), np.minimum(res.cig_off[:-1], len(cig_m)))
    m_per_read = np.where(np.diff(res.cig_off) > 0, m_per_read, 0)
    # k_emit: SEQ in + SEQ out (16-byte padded rows) + 2-bit genome words under the verified M bases
    # + the planned CIGAR (8 B per op) + MD + the 128-byte plan row.  (TE rows, the CIGAR copy and jM / jI
    # are moved by k_small; the segment lists k_emit reads are not counted.)
    seq_out_bytes = int(((res.seq_len.astype(np.int64) + 15) & ~15).sum())
    bytes_emit = (int(cols["seq"].nbytes) + seq_out_bytes + 0.25 * float(m_per_read[refreshed].sum())
                  + 8.0 * len(res.cig_op) + float(res.md_len.sum()) + 128.0 * int(primary.sum()))
    bytes_plan = (5.0 * len(cols["cig_op"]) + float(cols["md"].nbytes) + 8.0 * len(res.cig_op) + 16.0 * len(res.te)
                  + 96.0 * int(primary.sum()))
    # whole step against SURVEY.md 8(d)'s bytes per read (B_read): 4-bit SEQ in, CIGAR / MD in, the fixed record, 2-bit + case plane under
    # the exonic reference bases, ASCII SEQ out, CIGAR' / MD' / junction / TE rows out  (MD in is counted by its bytes)
    n_jn_total = int(len(res.jm)) if hasattr(res, "jm") else 0
    m_total = float(m_per_read[primary].sum())
    bytes_step = (0.5 * float(cols["seq"].nbytes) + 4.0 * len(cols["cig_op"]) + float(cols["md"].nbytes) + 32.0 * args.reads
                  + 0.375 * m_total + float(res.seq_len.astype(np.int64).sum()) + 4.0 * len(res.cig_op) + float(res.md_len.sum())
                  + 9.0 * n_jn_total + 16.0 * len(res.te) + 24.0 * args.reads)
    per = {k: v / steps for k, v in kern.items()}
    dom = max(("k_emit_ms", "k_plan_ms", "k_junction_ms", "k_nmmd_init_ms", "k_measure_ms"), key=lambda k: per[k])
    dom_bytes = bytes_emit if dom == "k_emit_ms" else bytes_plan
    peak, peak_src = peaks()
    achieved = dom_bytes / (per[dom] * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r02_kernel_traffic.json")
    if os.path.exists(tpath):            # DRAM bytes per read of that kernel from the committed ncu --set full captures
        per_read = json.load(open(tpath))["dram_bytes_per_read"].get(args.shape, {}).get(dom.replace("_ms", ""))
        traffic = per_read * args.reads if per_read is not None else None

    reads_total = args.reads * world
    line = {
        "metric": "reads corrected/sec", "value": reads_total * steps / (ms_run * 1e-3), "unit": "reads/s",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_run / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "shape": args.shape, "reads_per_gpu_per_step": args.reads,
                   "mean_read_len": float(cols["seq"].nbytes) / args.reads,
                   "cigar_form": "16-bit words + listed long ops" if "cig16" in hcols else "op + length arrays",
                   "seq_form": seq_form + " on the host side of the ABI (expanded on the device); the new SEQ comes back as a delta against the input SEQ",
                   "l2": "inputs (%.2f GB per step) are larger than L2; no flush needed" % (h2d_bytes / 1e9),
                   "parallelism": "reads sharded by input-order chunk, tables replicated, no collective"},
        "e2e": {"value": reads_total * steps / (ms_e2e * 1e-3), "unit": "reads/s", "ms_per_step": ms_e2e / steps,
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
        "gpu_launches": int(launches),
        "kernel_ms_per_step": per, "exact_nmmd_reads_per_step": int(exact_reads), "ascii_path_reads_per_step": int(ascii_reads),
        "roofline": {"bound": "hbm", "kernel": dom.replace("_ms", ""), "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": dom_bytes,
                     "note": "k_emit = the output stage (k_verify + k_emit over the reads it lists); the byte formula is SURVEY.md 8(d)'s "
                             "(ASCII SEQ in and out counted) although the stage itself now reads the 2-bit plane and writes no SEQ row: "
                             "`traffic` is what it moves"},
        "roofline_step": {"bound": "hbm", "achieved": bytes_step / (ms_run / steps * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                          "frac": bytes_step / (ms_run / steps * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_step": bytes_step,
                          "note": "every kernel of the step against SURVEY.md 8(d)'s B_read x reads"},
        "clocks": clocks,
    }
    if ms_dry > 0:
        line["dry_run"] = {"value": args.reads * world / (ms_dry * 1e-3), "e2e": args.reads * world / (ms_dry_e2e * 1e-3), "unit": "reads/s",
                           "ms_per_step": ms_dry, "e2e_ms_per_step": ms_dry_e2e,
                           "note": "--dryRun inventory (tc_dryrun_batch) of the same batch: device-resident and through the ABI with host buffers"}
    if world == 1 and rank == 0 and args.text_sample > 0:
        # SAM text in -> four output texts, as the CLI runs it: blocks of the text go through workers with
        # their own engine context (parse k+1 | GPU k | render k-1 overlap), texts delivered in input order
        from transcriptclean_b200 import api
        from transcriptclean_b200.cli import split_text
        ns = min(args.text_sample, args.reads)
        sam_lines = w.sam_lines(cols, nm, 0, ns)
        text = ("\n".join(sam_lines) + "\n").encode()
        n_workers = max(1, args.text_workers)
        sj_tab = build_sj_tables(sj_path, {w.chrom}, gi.index)
        refs_list = [api.make_refs(gi, sj_tab, None, engine=eng)] + [api.make_refs(gi, sj_tab, None, device=local) for _ in range(n_workers - 1)]
        blocks = split_text(text, max(n_workers, len(text) // (args.text_block_mb << 20)))
        sizes = {"sam": 0, "fa": 0, "log": 0, "te": 0}
        devnull = open(os.devnull, "wb")

        def sink(header, views):            # what the CLI does with the four texts: write them out
            for k, vs in views.items():
                for v in vs:
                    devnull.write(v)
                    sizes[k] += len(v)
        api.correct_sam_stream(blocks, api.Struct(), refs_list, sink)
        for k in sizes:
            sizes[k] = 0
        t0 = time.perf_counter()
        api.correct_sam_stream(blocks, api.Struct(), refs_list, sink)
        t1 = time.perf_counter()
        devnull.close()
        # the reference's own seam, as a maintainer would call it (option A): a list of SAM line strings in,
        # four open text files written - api.batch_correct on one context (str join / encode / decode included)
        import io
        import tempfile
        seam_n = min(100000, ns)
        seam_lines = sam_lines[:seam_n]
        seam_s = 1e30
        with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
            for _ in range(2):
                outs = api.Struct(sam=open(td + "/c.sam", "w"), fasta=open(td + "/c.fa", "w"), log=open(td + "/c.log", "w"), TElog=open(td + "/c.TE.log", "w"))
                ts0 = time.perf_counter()
                with contextlib.redirect_stdout(io.StringIO()):
                    api.batch_correct(seam_lines, api.Struct(), refs_list[0], outs)
                for fo in outs.values():
                    fo.close()
                seam_s = min(seam_s, time.perf_counter() - ts0)
        line["seam_batch_correct"] = {"value": seam_n / seam_s, "unit": "reads/s",
                                      "sample": "api.batch_correct(list of %d SAM line strings, options, refs, outfiles) on one engine context, "
                                                "the four text files written and closed, best of 2" % seam_n}
        for r in refs_list[1:]:
            r.engine.close()
        line["sam_text_path"] = {"value": ns / (t1 - t0), "unit": "reads/s", "in_bytes": len(text),
                                 "out_bytes": sum(sizes.values()),
                                 "sample": "%d reads as SAM text -> clean SAM + FASTA + log + TE.log texts; %d blocks through %d pipelined workers "
                                           "(own engine context each, one GPU), host threads=%d" % (ns, len(blocks), n_workers, min(32, os.cpu_count() or 1))}
    texts, n_texts = None, 0
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        ns = min(args.cpu_sample, args.reads)
        lines = w.sam_lines(cols, nm, 0, ns)
        _ORACLE["keep"] = True
        rate, sec = oracle_rate(w, sj_path, lines, 1)
        texts, n_texts = _ORACLE.get("texts"), ns
        line["cpu_baseline"] = {"value": rate, "unit": "reads/s", "cores": 1, "kind": "port",
                                "sample": "first %d reads of the same batch, 1 thread, oracle/tc_oracle.py with the "
                                          "reference's per-base genome walk (%.1f s)" % (ns, sec)}
    # ---- parity of the TIMED batch: the texts of its first reads (every rank checks its own shard) against the oracle
    n_par = args.parity_reads if args.parity_reads >= 0 else (min(args.cpu_sample, args.reads) if world == 1 else min(500, args.reads))
    bad = None
    if n_par > 0 and not args.debug_mask:
        if texts is not None and n_texts != n_par:
            texts = None
        bad = parity_check(w, gi, cols, nm, res, n_par, sj_path, texts)
        if bad:
            sys.stderr.write("PARITY FAILURE on rank %d: %s\n" % (rank, bad))
    flag = torch.tensor([1 if bad else 0, n_par if not args.debug_mask else 0], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.SUM)
    line["parity_checked_reads"] = int(flag[1])
    line["parity"] = "FAILED" if int(flag[0]) else ("ok: texts of the first %d reads of the timed batch per rank equal the oracle's" % n_par if n_par > 0 and not args.debug_mask else "not checked")
    if rank == 0:
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    if int(flag[0]):
        raise SystemExit(3)


if __name__ == "__main__":
    main()

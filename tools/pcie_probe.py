"""Host-link probe: pinned H2D / D2H bandwidth per GPU, alone and with both directions (and every GPU of the job) busy at once.

    python tools/pcie_probe.py                                   one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py   N GPUs at once

Prints one JSON line (rank 0): per-direction GB/s per GPU (min / mean over ranks) and the aggregate over the box - the ceiling
of bench.py's `e2e` figure is aggregate / (bytes per read over the host links)."""
import json
import os
import time

import torch
import torch.distributed as dist

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
h_in, h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True), torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_a, d_b = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d(); d2h()


def chunks(k):                      # the pipeline's pattern: many medium copies per direction
    c = n // k

    def f():
        for i in range(k):
            with torch.cuda.stream(s1):
                d_a[i * c:(i + 1) * c].copy_(h_in[i * c:(i + 1) * c], non_blocking=True)
            with torch.cuda.stream(s2):
                h_out[i * c:(i + 1) * c].copy_(d_b[i * c:(i + 1) * c], non_blocking=True)
    return f


def gbs(fn, rep=4):
    best = 1e9
    for _ in range(rep):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return n / best / 1e9


mine = torch.tensor([gbs(h2d), gbs(d2h), gbs(both), gbs(chunks(64))], dtype=torch.float64, device="cuda")
allr = [torch.zeros_like(mine) for _ in range(world)]
if world > 1:
    dist.all_gather(allr, mine)
else:
    allr = [mine]
if rank == 0:
    m = torch.stack(allr).cpu()
    names = ["h2d_alone", "d2h_alone", "both_each_direction", "both_64_chunks_each_direction"]
    out = {"gpus_busy_at_once": world, "GBs_per_gpu_min": {k: float(m[:, i].min()) for i, k in enumerate(names)},
           "GBs_per_gpu_mean": {k: float(m[:, i].mean()) for i, k in enumerate(names)},
           "aggregate_GBs_both_directions_all_gpus": float(2 * m[:, 2].sum())}
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()

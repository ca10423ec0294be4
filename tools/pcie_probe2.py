"""Two processes, one per GPU, both moving pinned host buffers both ways at once: per-GPU GB/s."""
import os, sys, time, json, torch
import torch.distributed as dist
r = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(r)
dist.init_process_group("nccl", device_id=torch.device("cuda", r))
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
best = 1e9
for _ in range(4):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter(); both(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
print(json.dumps({"rank": r, "each_direction_GBs": n / best / 1e9, "world": dist.get_world_size()}), flush=True)
dist.destroy_process_group()

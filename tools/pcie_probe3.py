"""Does the copy-size mix of one pipeline chunk explain the e2e bandwidth?  Both directions at once,
8 chunks, each chunk = the copy sizes of a 131072-read chunk (MB)."""
import torch, time, json
down = [60.3, 38.4, 5.2, 5.7, 1.4, 5.2, 0.65, 1.05] + [1.05] * 5 + [0.52] * 4
up = [120.0, 19.0, 7.5, 3.8] + [1.05] * 5 + [0.52] * 3 + [0.13]
def bufs(sizes):
    return [(torch.empty(int(s * 1e6), dtype=torch.uint8, pin_memory=True), torch.empty(int(s * 1e6), dtype=torch.uint8, device="cuda")) for s in sizes]
bd, bu = bufs(down), bufs(up)
big_h = torch.empty(int(sum(down) * 1e6), dtype=torch.uint8, pin_memory=True); big_d = torch.empty_like(big_h, device="cuda")
bigu_h = torch.empty(int(sum(up) * 1e6), dtype=torch.uint8, pin_memory=True); bigu_d = torch.empty_like(bigu_h, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def mix():
    for _ in range(8):
        with torch.cuda.stream(s1):
            for h, d in bu: d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            for h, d in bd: h.copy_(d, non_blocking=True)
def single():
    for _ in range(8):
        with torch.cuda.stream(s1): bigu_d.copy_(bigu_h, non_blocking=True)
        with torch.cuda.stream(s2): big_h.copy_(big_d, non_blocking=True)
def t(fn):
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best
tm, ts = t(mix), t(single)
print(json.dumps({"mix_ms": tm * 1e3, "single_ms": ts * 1e3, "down_MB_per_chunk": sum(down), "up_MB_per_chunk": sum(up),
                  "mix_down_GBs": 8 * sum(down) / 1e3 / tm, "single_down_GBs": 8 * sum(down) / 1e3 / ts}))

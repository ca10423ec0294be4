"""PCIe probe: pinned H2D / D2H bandwidth alone and concurrently (sizes like one bench step)."""
import torch, time, json
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, rep=3):
    best = 1e9
    for _ in range(rep):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best
def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
def both(): h2d(); d2h()
def chunks(k):
    c = n // k
    def f():
        for i in range(k):
            with torch.cuda.stream(s1): d_a[i*c:(i+1)*c].copy_(h_in[i*c:(i+1)*c], non_blocking=True)
            with torch.cuda.stream(s2): h_out[i*c:(i+1)*c].copy_(d_b[i*c:(i+1)*c], non_blocking=True)
    return f
r = {"h2d_GBs": n / t(h2d) / 1e9, "d2h_GBs": n / t(d2h) / 1e9, "both_each_GBs": n / t(both) / 1e9,
     "both_64chunks_each_GBs": n / t(chunks(64)) / 1e9, "both_1024chunks_each_GBs": n / t(chunks(1024)) / 1e9}
print(json.dumps(r))

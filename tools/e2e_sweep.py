"""e2e (host buffers through tc_correct_batch) over pipeline chunk sizes; wall-clock per call."""
import os, sys, time, json
import numpy as np, torch, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from tools.workload import Workload
from transcriptclean_b200 import engine as E
from transcriptclean_b200.tables import build_sj_tables
n = int(os.environ.get("N", "1000000"))
w = Workload(seed=20261019, chrom_len=248956422, n_genes=3000)
sj = bench.write_sj(w)
shape = os.environ.get("SHAPE", "pacbio")
cols, nm = w.reads(shape, n, seed_offset=0)
gi = w.genome_image()
eng = E.Engine(0); eng.load_genome(gi); eng.load_junctions(build_sj_tables(sj, {w.chrom}, gi.index))
if not os.environ.get("SEQ8"):
    cols = (None if os.environ.get("SEQ4") else eng.pack_cols2(cols)) or eng.pack_cols(cols)
if not os.environ.get("CIG40"):
    cols = eng.pack_cigar(cols)
    if "cig16" in cols: cols = {k: v for k, v in cols.items() if k not in ("cig_op", "cig_len")}
pinned = {}
b = E._BatchIn(); b.n_reads = n
for k, a in cols.items():
    if not isinstance(a, np.ndarray): continue
    t = torch.empty(max(a.nbytes, 8), dtype=torch.uint8, pin_memory=True)
    if a.nbytes: t[:a.nbytes].copy_(torch.from_numpy(a.view(np.uint8).reshape(-1)))
    pinned[k] = t; setattr(b, k, t.data_ptr())
b.seq_bits = int(cols.get("seq_bits", 4 if "seq_poff" in cols else 8)); b.n_exc = len(cols["exc_pos"]) if "exc_pos" in cols else 0; b.n_cig_exc = len(cols["cig_exc_idx"]) if "cig_exc_idx" in cols else 0
res = {}
for ch in [int(x) for x in os.environ.get("CHUNKS", "32768,65536,131072,262144,524288").split(",")]:
    eng.set_pipeline(ch)
    for _ in range(2): eng.raw_correct(b)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): eng.raw_correct(b)
    torch.cuda.synchronize(); res[ch] = (time.perf_counter() - t0) / 3 * 1e3
t = eng.timing()
print(json.dumps({"shape": shape, "reads": n, "ms_per_call": res, "reads_per_s": {k: n / (v * 1e-3) for k, v in res.items()},
                  "exact_nmmd_reads": t["exact_nmmd_reads"], "h2d_bytes": t["h2d_bytes"], "d2h_bytes": t["d2h_bytes"]}))

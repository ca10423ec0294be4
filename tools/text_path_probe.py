"""Where the SAM-text path spends its time on the GPU box: parse / pack / correct (pageable host buffers) / render."""
import os, sys, time, json, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from tools.workload import Workload
from transcriptclean_b200 import engine as E
from transcriptclean_b200.tables import build_sj_tables
n = int(os.environ.get("N", "200000"))
w = Workload(seed=20261019, chrom_len=248956422, n_genes=3000)
sj = bench.write_sj(w)
cols, nm = w.reads("pacbio", n)
gi = w.genome_image()
eng = E.Engine(0); eng.load_genome(gi); eng.load_junctions(build_sj_tables(sj, {w.chrom}, gi.index))
text = ("\n".join(w.sam_lines(cols, nm, 0, n)) + "\n").encode()
lib = eng._lib
names = (C.c_char_p * len(gi.names))(*[x.encode() for x in gi.names]); err = C.create_string_buffer(1024)
nt = min(32, os.cpu_count() or 1)
res = []
for rep in range(4):
    t0 = time.perf_counter()
    st = eng.parse_sam(text, gi)                      # parse + pack
    t1 = time.perf_counter()
    raw = eng.raw_correct(st.batch)
    t2 = time.perf_counter()
    out = st.render(raw, copy=False)
    t3 = time.perf_counter()
    for v in out.values(): v.release()
    st.close()
    t4 = time.perf_counter()
    res.append({"parse_pack_ms": (t1 - t0) * 1e3, "correct_ms": (t2 - t1) * 1e3, "render_ms": (t3 - t2) * 1e3, "close_ms": (t4 - t3) * 1e3})
print(json.dumps({"reads": n, "threads": nt, "text_MB": len(text) >> 20, "runs": res}))

"""Host side of the SAM-text path by thread count: tc_sam_parse and tc_sam_render_parts on one block of PacBio-shape reads
(N reads, default 100,000), NT = 1..all cores; us*thread per read shows how the C++ threads scale on the box."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from tools.workload import Workload
from transcriptclean_b200 import engine as E
from transcriptclean_b200.tables import build_sj_tables
n = int(os.environ.get("N", "100000"))
w = Workload(seed=20261019, chrom_len=248956422, n_genes=3000)
sj = bench.write_sj(w)
cols, nm = w.reads("pacbio", n)
gi = w.genome_image()
eng = E.Engine(0); eng.load_genome(gi); eng.load_junctions(build_sj_tables(sj, {w.chrom}, gi.index))
text = ("\n".join(w.sam_lines(cols, nm, 0, n)) + "\n").encode()
res = {"reads": n, "text_bytes": len(text), "cores": os.cpu_count(), "by_threads": {}}
nts = [int(x) for x in os.environ.get("NTS", "1,2,4,8,16,32").split(",") if int(x) <= (os.cpu_count() or 1)]
for nt in nts:
    best = {}
    for rep in range(3):
        t0 = time.perf_counter(); st = eng.parse_sam(text, gi, nt); t1 = time.perf_counter()
        raw = eng.raw_correct(st.batch); t2 = time.perf_counter()
        out = st.render_parts(raw, False, False, False); t3 = time.perf_counter()
        for vs in out.values():
            for v in vs: v.release()
        st.close(); t4 = time.perf_counter()
        cur = {"parse_ms": (t1 - t0) * 1e3, "abi_ms": (t2 - t1) * 1e3, "render_ms": (t3 - t2) * 1e3, "close_ms": (t4 - t3) * 1e3}
        for k, v in cur.items(): best[k] = min(best.get(k, 1e30), v)
    best["parse_us_thread_per_read"] = best["parse_ms"] * 1e3 * nt / n; best["render_us_thread_per_read"] = best["render_ms"] * 1e3 * nt / n
    res["by_threads"][nt] = best
print(json.dumps(res))

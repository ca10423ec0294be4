"""Copies the measured artefacts of a gpurun_out/ tag into profiles/ (tracked) with readable summaries.
usage: python tools/make_profiles.py TAG ROUNDTAG   e.g.  r1b r01b"""
import csv, io, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, out = sys.argv[1], sys.argv[2]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
for src, dst in ((f"bench_{tag}.json", f"{out}_bench.json"), (f"bench_{tag}_ont.json", f"{out}_bench_ont_shape.json"),
                 (f"bench_{tag}_seq8.json", f"{out}_bench_ascii_seq.json"), (f"bench_{tag}_reference.json", f"{out}_bench_reference_arm.json"),
                 (f"bench_{tag}_n2.json", f"{out}_bench_2gpu.json"), (f"launches_{tag}.csv", f"{out}_launches.csv")):
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))
# launch list summary
lp = os.path.join(G, f"launches_{tag}.csv")
if os.path.exists(lp):
    rows = [r for r in csv.reader(l for l in open(lp) if not l.startswith("=="))]
    h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = {}
    for r in rows[1:]:
        if len(r) <= vi or not r[vi].replace(".", "").replace(",", "").isdigit(): continue
        k = r[ki].split("(")[0][:60]; v = float(r[vi].replace(",", ""))
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    unit = rows[1][h.index("Metric Unit")] if "Metric Unit" in h else "ns"
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(P, f"{out}_launches_summary.txt"), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 800 : python bench.py --no-cpu-baseline --text-sample 0 --steps 2 --warmup 3 --resident-only\n"
                "(1M PacBio-shape reads per launch; --resident-only = the device-resident steps of bench.py without the pipelined e2e calls,\n"
                " i.e. exactly the kernels behind `value` / `kernel_ms_per_step`; k_seq_unpack / k_seq_pack belong to the one upload / download)\n"
                "per-launch times are cold-cache and serialised; compare shares\n")
        f.write("%-62s %8s %12s %7s\n" % ("kernel", "launches", "total_us", "share"))
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("%-62s %8d %12.1f %6.1f%%\n" % (k, a[0], a[1] * scale, 100 * a[1] / tot))
# ncu --set full summary of k_emit
rp = os.path.join(G, f"prof_{tag}_emit.ncu-rep")
if os.path.exists(rp):
    raw = subprocess.run(["ncu", "-i", rp, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw))); h = r[0]; row = r[2]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
            "launch__occupancy_limit_registers", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio", "l1tex__t_sector_hit_rate.pct",
            "lts__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
            "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
    nreads = 120000
    val = lambda n: float(row[h.index(n)].replace(",", ""))
    unit = lambda n: r[1][h.index(n)]
    tobytes = lambda n: val(n) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[unit(n)]
    dram = tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")
    with open(os.path.join(P, f"{out}_ncu_full_k_emit_summary.txt"), "w") as f:
        f.write("ncu --set full --clock-control none --import-source on -k regex:^k_emit -s 3 -c 1 : python bench.py --no-cpu-baseline --text-sample 0 --steps 1 --warmup 3 --reads 120000\n"
                "(one launch over 120,000 PacBio-shape reads)\n")
        for w in want:
            if w in h: f.write("%-78s %s %s\n" % (w, row[h.index(w)], r[1][h.index(w)]))
        for i, name in enumerate(h):
            if name.startswith("smsp__average_warps_issue_stalled") and name.endswith("per_issue_active.ratio") and float(row[i] or 0) > 0.3:
                f.write("%-78s %s\n" % (name, row[i]))
        f.write("\nDRAM traffic per launch = %.1f MB = %.0f bytes per read\n" % (dram / 1e6, dram / nreads))
    json.dump({"kernel": "k_emit", "reads_in_profiled_launch": nreads, "dram_bytes_per_launch": dram, "dram_bytes_per_read": dram / nreads,
               "source": f"profiles/{out}_ncu_full_k_emit_summary.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"},
              open(os.path.join(P, "r01_k_emit_traffic.json"), "w"), indent=1)
print("ok")

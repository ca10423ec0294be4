"""Readable summary of an ncu --set full capture (one section per kernel launch): python tools/prof_summary.py REP N_READS [title] > profiles/...txt"""
import csv, io, subprocess, sys
rep, nreads = sys.argv[1], float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(out))); h = r[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
if len(sys.argv) > 3: print(sys.argv[3])
def num(row, n):
    try: return float(row[h.index(n)].replace(",", ""))
    except Exception: return float("nan")
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
for row in r[2:]:
    print("\n== %s" % row[h.index('Kernel Name')])
    for w in want:
        if w in h: print("%-72s %s %s" % (w, row[h.index(w)], r[1][h.index(w)]))
    for i, name in enumerate(h):
        if name.startswith('smsp__average_warps_issue_stalled') and name.endswith('per_issue_active.ratio') and float(row[i] or 0) > 0.3:
            print("%-72s %s" % (name.replace('smsp__average_warps_issue_stalled_', 'stalled warps per issue: ').replace('_per_issue_active.ratio', ''), row[i]))
    dram = sum(num(row, n) * scale.get(r[1][h.index(n)], 1.0) for n in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
    us = num(row, 'gpu__time_duration.sum') * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}.get(r[1][h.index('gpu__time_duration.sum')], 1.0)
    ld_r, ld_s = num(row, 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum'), num(row, 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum')
    st_r, st_s = num(row, 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum'), num(row, 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum')
    print("-> per read: %.0f DRAM bytes, %.0f warp-instructions; %.2f ms per 1M reads; DRAM %.0f GB/s; sectors per request: loads %.1f, stores %.1f; lanes active per instruction %.1f of 32"
          % (dram / nreads, num(row, 'smsp__inst_executed.sum') / nreads, us / nreads * 1e3, dram / us / 1e3,
             ld_s / ld_r if ld_r else 0, st_s / st_r if st_r else 0, num(row, 'smsp__thread_inst_executed_per_inst_executed.ratio')))

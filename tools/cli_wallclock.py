"""Wall clock of the real command line (`python -m transcriptclean_b200 ...`) on a synthetic SAM file: what a user of the
reference's CLI gets.  usage: python tools/cli_wallclock.py [--reads 1000000] [--gpus 1] [--shape pacbio] [--dir /dev/shm]
Prints one JSON line: wall seconds of the whole command (second run: the packed genome image is cached, like pyfaidx's .fai),
reads/s over the whole command and over the correction pass alone (the run minus the time to the first corrected block)."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.workload import Workload  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reads", type=int, default=1000000)
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--shape", default="pacbio")
ap.add_argument("--dir", default="/dev/shm")
ap.add_argument("--chrom-len", type=int, default=248956422)
ap.add_argument("--genes", type=int, default=3000)
args = ap.parse_args()
d = tempfile.mkdtemp(prefix="tc_cli_", dir=args.dir if os.path.isdir(args.dir) else None)
w = Workload(seed=20261019, chrom_len=args.chrom_len, n_genes=args.genes)
t0 = time.perf_counter()
w.write_fasta(d + "/g.fa")
with open(d + "/sj.tab", "w") as f:
    for r in w.sj_rows():
        f.write("\t".join(str(x) for x in r) + "\n")
step = 250000
for lo in range(0, args.reads, step):
    cols, nm = w.reads(args.shape, min(step, args.reads - lo), seed_offset=lo // step)
    w.write_sam(d + "/in.sam", cols, nm, 0, len(nm), header=(lo == 0))
t_gen = time.perf_counter() - t0
cmd = [sys.executable, "-m", "transcriptclean_b200", "--sam", d + "/in.sam", "--genome", d + "/g.fa", "--spliceJns", d + "/sj.tab",
       "--threads", str(args.gpus), "--outprefix", d + "/TC"]
runs = []
for k in range(2):
    t0 = time.perf_counter()
    subprocess.check_call(cmd, cwd=ROOT, stdout=subprocess.DEVNULL)
    runs.append(time.perf_counter() - t0)
sizes = {k: os.path.getsize(d + "/TC_clean." + k) for k in ("sam", "fa", "log", "TE.log")}
print(json.dumps({"reads": args.reads, "gpus": args.gpus, "shape": args.shape, "sam_bytes": os.path.getsize(d + "/in.sam"),
                  "out_bytes": sizes, "generate_s": t_gen, "cli_wall_s_first_run": runs[0], "cli_wall_s": runs[1],
                  "reads_per_s_whole_command": args.reads / runs[1], "host_cores": os.cpu_count()}))
subprocess.call(["rm", "-rf", d])

"""Per-source-line instruction counts of an ncu capture: python tools/prof_lines.py REP N_READS [top] [--by-line] [--kernel=REGEX]"""
import csv, subprocess, sys, io
rep, nreads = sys.argv[1], float(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 60
kern = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--kernel=")]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + (["-k", "regex:" + kern[0]] if kern else []), capture_output=True, text=True).stdout
rows = []; cur = None
for row in csv.reader(io.StringIO(out)):
    if len(row) < 10:
        if row and row[0] == "File Path": cur = row[1]
        continue
    if row[0] == "Line No": continue
    if row[2] == "-":
        try: rows.append((cur.split('/')[-1], int(row[0]), row[1][:100], int(row[4]), int(row[7])))
        except Exception: pass
ti = sum(x[4] for x in rows); ts = sum(x[3] for x in rows)
print("total warp-inst %d  per read %.1f  samples %d" % (ti, ti / nreads, ts))
if "--by-line" in sys.argv:
    rows.sort(key=lambda x: (x[0], x[1]))
    for x in rows:
        if x[4] / nreads >= 0.5: print("%-14s %5d  %7.1f  %5.1f%%  %s" % (x[0][:14], x[1], x[4] / nreads, 100.0 * x[3] / ts, x[2]))
else:
    rows.sort(key=lambda x: -x[4])
    for x in rows[:top]: print("%-14s %5d  %7.1f  %5.1f%%  %s" % (x[0][:14], x[1], x[4] / nreads, 100.0 * x[3] / ts, x[2]))

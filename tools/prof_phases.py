"""Phase sums of k_emit from a prof_lines --by-line dump: python tools/prof_phases.py LINES.txt"""
import re, sys
import os
src = open(os.environ.get('TC_EMIT_SRC', '/root/repo/transcriptclean_b200/csrc/tc_emit.cuh')).read().splitlines()
def find(pat, start=0):
    for i in range(start, len(src)):
        if pat in src[i]: return i + 1
    raise KeyError(pat)
pf = find("// the next read's segment list and SEQ lines")
marks = [("loop/plan row", find("k_emit(Params P, OutDev O")), ("setup2", find("(TE rows, the CIGAR and the jM")),
         ("SEQ setup", find("// ---- SEQ")), ("(A) seg list", find("// (A) segment list")),
         ("prefetch", pf), ("exc+chunk_bp", find("if (fast) {", pf)), ("(B)", find("// (B) bulk copy")), ("(C)", find("// (C) the part")),
         ("(D)", find("// (D) genome slices")), ("(E)", find("// (E) ")), ("T1 scan", find("// ---- NM / MD")),
         ("T1 verify", find("// The staged SEQ is verified")), ("T1 md", find("// every aligned base equals the genome")),
         ("exact+finish", find("if (!done) {")), ("end", len(src) + 1)]
rows = []
for l in open(sys.argv[1]).read().splitlines()[1:]:
    m = re.match(r"(\S+)\s+(\d+)\s+([\d.]+)\s+([\d.]+)%", l)
    if m: rows.append((m.group(1), int(m.group(2)), float(m.group(3)), float(m.group(4))))
for (n, a), (n2, b) in zip(marks, marks[1:]):
    sel = [r for r in rows if r[0] == "tc_emit.cuh" and a <= r[1] < b]
    print("%-16s %4d-%4d  inst/read %7.1f  samples %5.1f%%" % (n, a, b, sum(r[2] for r in sel), sum(r[3] for r in sel)))
oth = [r for r in rows if not (r[0] == "tc_emit.cuh" and marks[0][1] <= r[1] < marks[-1][1])]
print("other            inst/read %7.1f  samples %5.1f%%" % (sum(r[2] for r in oth), sum(r[3] for r in oth)))
for r in oth:
    if r[2] > 8 or r[3] > 1: print("   ", r)

"""Golden vector for transcriptclean_b200/get_sjs_from_gtf.py: runs the UNMODIFIED reference script
(/root/reference/TranscriptClean/accessory_scripts/get_SJs_from_gtf.py) under the pyfaidx stub on a seeded
synthetic GTF + genome and stores inputs and output under tests/golden/.  Build container only.

    python oracle/make_golden_gtf.py
"""
import gzip
import json
import os
import random
import runpy
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/TranscriptClean/accessory_scripts/get_SJs_from_gtf.py"


def make_inputs(seed=20261020):
    rng = random.Random(seed)
    genome = {}
    for name, n in (("chr1", 30000), ("chr2", 12000)):
        s = [rng.choice("ACGT") for _ in range(n)]
        for _ in range(6):                                   # soft-masked stretches and N runs
            a = rng.randrange(0, n - 400); b = a + rng.randrange(50, 400)
            for k in range(a, b):
                s[k] = s[k].lower()
        a = rng.randrange(1000, n - 300)
        for k in range(a, a + 120):
            s[k] = "N"
        genome[name] = s
    motifs = ["GTAG", "CTAC", "GCAG", "CTGC", "ATAC", "GTAT", "AAAA", "gtag", "GtAg"]
    lines = ["##description: synthetic", "#!genome-build test"]
    tid = 0
    for chrom in genome:
        n = len(genome[chrom])
        for g in range(10):
            strand = rng.choice("+-")
            start = rng.randrange(100, n - 6000)
            exons, pos = [], start
            for _ in range(rng.randrange(1, 7)):
                e0 = pos; e1 = e0 + rng.randrange(30, 300)
                exons.append((e0, e1))
                pos = e1 + 1 + rng.choice([5, 12, 20, 21, 22, 60, 300, 900])
            for (a0, a1), (b0, b1) in zip(exons, exons[1:]):      # plant a motif in each intron (a1+1 .. b0-1)
                m = rng.choice(motifs)
                i0, i1 = a1 + 1, b0 - 1
                if i1 - i0 >= 3:
                    genome[chrom][i0 - 1], genome[chrom][i0] = m[0], m[1]
                    genome[chrom][i1 - 2], genome[chrom][i1 - 1] = m[2], m[3]
            for rep in range(rng.choice([1, 1, 2])):              # some transcripts twice (duplicate junction rows)
                tid += 1
                t = "T%d" % tid
                lines.append("\t".join([chrom, "syn", "transcript", str(exons[0][0]), str(exons[-1][1]), ".", strand, ".",
                                        'gene_id "G%d"; transcript_id "%s";' % (g, t)]))
                order = exons if strand == "+" else exons[::-1]
                for k, (e0, e1) in enumerate(order):
                    lines.append("\t".join([chrom, "syn", "exon", str(e0), str(e1), ".", strand, ".",
                                            'gene_id "G%d"; transcript_id "%s"; exon_number "%d";' % (g, t, k + 1)]))
                if rng.random() < 0.3:
                    lines.append("\t".join([chrom, "syn", "CDS", str(exons[0][0]), str(exons[0][1]), ".", strand, "0",
                                            'gene_id "G%d"; transcript_id "%s";' % (g, t)]))
        lines.append("\t".join([chrom, "syn", "exon", "50", "90", ".", "+", ".", 'gene_id "Gx";']))   # no transcript_id: skipped
    return {k: "".join(v) for k, v in genome.items()}, lines


def main():
    sys.path.insert(0, os.path.join(HERE, "stubs"))
    genome, lines = make_inputs()
    d = tempfile.mkdtemp(prefix="tc_gtf_")
    with open(d + "/g.fa", "w") as f:
        for name, s in genome.items():
            f.write(">" + name + " synthetic\n")
            for i in range(0, len(s), 70):
                f.write(s[i:i + 70] + "\n")
    with open(d + "/a.gtf", "w") as f:
        f.write("\n".join(lines) + "\n")
    out = {}
    for m in (21, 13):
        argv = sys.argv
        sys.argv = [REF, "--f", d + "/a.gtf", "--g", d + "/g.fa", "--o", d + "/out_%d.txt" % m, "-m", str(m)]
        try:
            runpy.run_path(REF, run_name="__main__")
        finally:
            sys.argv = argv
        out[str(m)] = open(d + "/out_%d.txt" % m).read()
    rec = {"genome": genome, "gtf_lines": lines, "expected": out,
           "made_by": "oracle/make_golden_gtf.py: unmodified accessory_scripts/get_SJs_from_gtf.py under oracle/stubs/pyfaidx.py"}
    path = os.path.join(ROOT, "tests", "golden", "gtf_sjs.json.gz")
    with gzip.open(path, "wt") as f:
        json.dump(rec, f)
    print("wrote", path, {k: v.count("\n") for k, v in out.items()})


if __name__ == "__main__":
    main()

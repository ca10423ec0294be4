"""Seeded small-scale generator of SAM reads + genome + SJ table + VCF for parity tests.

Covers the behaviours listed in SURVEY.md 3.4 (Q1-Q25): soft-masked and N genome runs,
mismatches / insertions / deletions of both correctable and too-large size, junctions shifted
off their annotated position, soft and hard clips, minus strand, unmapped / secondary records,
reads with and without NM/MD/jM/jI tags, MD strings without `0` separators, MD that overruns
the CIGAR, variants that match planted errors, multi-allelic records.

By default no two annotated sites are equidistant from a shifted bound and no annotated site
exists only on the opposite strand within reach (the two pyranges semantics the reference's
tests do not pin); `unpinned=True` adds such cases.
"""
import random

BASES = "ACGT"
_MOTIFS_PLUS = [("GT", "AG")] * 12 + [("GC", "AG"), ("AT", "AC")]
_RC = str.maketrans("ACGTacgt", "TGCAtgca")


def _rc(s):
    return s.translate(_RC)[::-1]


class Case:
    """genome: dict name->str; sj_rows: list of 9-col rows; vcf_rows: list of 5+ col rows;
    lines: SAM lines (no header)."""

    def __init__(self):
        self.genome, self.sj_rows, self.vcf_rows, self.lines = {}, [], [], []

    def write_sj(self, path):
        with open(path, "w") as f:
            for r in self.sj_rows:
                f.write("\t".join(str(x) for x in r) + "\n")

    def write_vcf(self, path):
        with open(path, "w") as f:
            f.write("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
            for r in self.vcf_rows:
                f.write("\t".join(str(x) for x in r) + "\n")

    def write_fasta(self, path):
        with open(path, "w") as f:
            for name, s in self.genome.items():
                f.write(">" + name + "\n")
                for i in range(0, len(s), 60):
                    f.write(s[i:i + 60] + "\n")

    def write_sam(self, path, header=True):
        with open(path, "w") as f:
            if header:
                f.write("@HD\tVN:1.0\tSO:unsorted\n")
                for name, s in self.genome.items():
                    f.write("@SQ\tSN:%s\tLN:%d\n" % (name, len(s)))
            for l in self.lines:
                f.write(l + "\n")


def _genome(rng, n_chrom, length):
    g = {}
    for c in range(n_chrom):
        s = [rng.choice(BASES) for _ in range(length)]
        for _ in range(length // 5000):                # soft-masked runs
            a = rng.randrange(length)
            for i in range(a, min(length, a + rng.randrange(50, 1500))):
                s[i] = s[i].lower()
        for _ in range(max(1, length // 40000)):       # N runs and stray IUPAC codes
            a = rng.randrange(length)
            for i in range(a, min(length, a + rng.randrange(5, 300))):
                s[i] = "N"
        for _ in range(length // 20000):
            s[rng.randrange(length)] = rng.choice("RYKMSWnry")
        g["chr%d" % (c + 1) if c < 2 else "chrUn_KI27%d" % c] = s
    return g


def _genes(rng, g, per_chrom, sj_rows, strand0_frac=0.05, exon_len=(30, 400)):
    """Lay out genes (exon blocks) and plant splice motifs; returns list of gene dicts."""
    genes = []
    for chrom, s in g.items():
        L = len(s)
        pos = 2000
        for _ in range(per_chrom):
            strand = rng.choice("+-")
            n_ex = rng.randrange(2, 9)
            exons, p = [], pos
            for e in range(n_ex):
                el = rng.randrange(exon_len[0], exon_len[1])
                exons.append((p, p + el - 1))           # 1-based inclusive
                p += el + rng.randrange(60, 3000)
            if p + 3000 >= L:
                break
            pos = p + rng.randrange(200, 3000)
            for (a0, a1), (b0, b1) in zip(exons[:-1], exons[1:]):
                i0, i1 = a1 + 1, b0 - 1                 # intron, 1-based inclusive
                d, a = rng.choice(_MOTIFS_PLUS)
                if rng.random() < 0.06:
                    d, a = rng.choice(BASES) + rng.choice(BASES), rng.choice(BASES) + rng.choice(BASES)
                left, right = (d, a) if strand == "+" else (_rc(a), _rc(d))
                for k, ch in enumerate(left):
                    s[i0 - 1 + k] = ch if rng.random() > 0.1 else ch.lower()
                for k, ch in enumerate(right):
                    s[i1 - 2 + k] = ch
                col4 = (1 if strand == "+" else 2) if rng.random() > strand0_frac else 0
                sj_rows.append([chrom, i0, i1, col4, 1, rng.randrange(2), rng.randrange(50),
                                rng.randrange(5), rng.randrange(10, 60)])
            genes.append({"chrom": chrom, "strand": strand, "exons": exons})
    return genes


def _std_md(read_seq_aligned):
    """MD/NM from a list of alignment columns ('M', readbase, refbase) / ('D', refbases) /
    ('I', n); standard SAM semantics."""
    md, run, nm = [], 0, 0
    for col in read_seq_aligned:
        if col[0] == "M":
            if col[1].upper() == col[2].upper():
                run += 1
            else:
                md.append(str(run) + col[2])
                run = 0
                nm += 1
        elif col[0] == "D":
            md.append(str(run) + "^" + col[1])
            run = 0
            nm += len(col[1])
        elif col[0] == "I":
            nm += col[1]
    md.append(str(run))
    return nm, "".join(md)


def make_case(seed, n_reads=200, n_chrom=2, chrom_len=120000, genes_per_chrom=12, err=0.02,
              shift_frac=0.25, unpinned=False, variant_frac=0.3, tagless_frac=0.2,
              quirks=True, exon_len=(30, 400)):
    rng = random.Random(seed)
    case = Case()
    g = _genome(rng, n_chrom, chrom_len)
    genes = _genes(rng, g, genes_per_chrom, case.sj_rows, exon_len=exon_len)
    case.genome = {k: "".join(v) for k, v in g.items()}
    annotated = {}
    for r in case.sj_rows:
        if r[3] != 0:
            annotated.setdefault(r[0], []).extend([r[1], r[2]])
    vcf = []
    for n in range(n_reads):
        gene = rng.choice(genes)
        chrom, gs = gene["chrom"], case.genome[gene["chrom"]]
        ex = gene["exons"]
        i = rng.randrange(len(ex))
        j = rng.randrange(i, len(ex))
        blocks = [list(b) for b in ex[i:j + 1]]
        blocks[0][0] += rng.randrange(0, max(1, (blocks[0][1] - blocks[0][0]) // 2))
        blocks[-1][1] -= rng.randrange(0, max(1, (blocks[-1][1] - blocks[-1][0]) // 2))
        # shift junction bounds (the aligner put them a few bases off)
        for k in range(len(blocks) - 1):
            if rng.random() < shift_frac:
                mode = rng.randrange(4)
                d0 = rng.randrange(-7, 8) if mode in (0, 2) else 0
                d1 = rng.randrange(-7, 8) if mode in (1, 2) else 0
                if mode == 3:
                    d0 = d1 = rng.randrange(-12, 13)
                if blocks[k][1] + d0 - blocks[k][0] > 8 and blocks[k + 1][1] - (blocks[k + 1][0] + d1) > 8:
                    blocks[k][1] += d0
                    blocks[k + 1][0] += d1
        strand = gene["strand"] if rng.random() > 0.03 else rng.choice("+-")
        # build alignment columns with errors
        cols, cigar = [], []

        def push(op, ct):
            if ct <= 0:
                return
            if cigar and cigar[-1][0] == op:
                cigar[-1][1] += ct
            else:
                cigar.append([op, ct])

        seq = []
        e_here = err * rng.choice([0, 0.5, 1, 1, 2, 4])
        for bi, (b0, b1) in enumerate(blocks):
            p = b0
            while p <= b1:
                r = rng.random()
                edge = (p - b0 < 3) or (b1 - p < 3)
                if r < e_here / 3 and not edge:                       # deletion
                    dl = rng.choice([1, 1, 1, 2, 2, 3, 4, 5, 6, 9])
                    dl = min(dl, b1 - p - 2)
                    if dl > 0:
                        ref = gs[p - 1:p - 1 + dl]
                        cols.append(("D", ref))
                        push("D", dl)
                        if rng.random() < variant_frac:
                            anchor = gs[p - 2].upper()
                            vcf.append([chrom, p - 1, ".", anchor + ref.upper(), anchor])
                        p += dl
                        continue
                if r > 1 - e_here / 3 and not edge:                   # insertion
                    il = rng.choice([1, 1, 1, 2, 2, 3, 4, 5, 6, 8])
                    ins = "".join(rng.choice(BASES) for _ in range(il))
                    seq.append(ins)
                    cols.append(("I", il))
                    push("I", il)
                    if rng.random() < variant_frac:
                        anchor = gs[p - 2].upper()
                        alts = [anchor + ins]
                        if rng.random() < 0.3:
                            alts.insert(0, anchor + "".join(rng.choice(BASES) for _ in range(il)))
                        if rng.random() < 0.3:   # same key, wrong allele -> still corrected
                            alts = [anchor + ("A" * il if ins != "A" * il else "C" * il)]
                        vcf.append([chrom, p - 1, ".", anchor, ",".join(alts)])
                refb = gs[p - 1]
                last = (bi == len(blocks) - 1 and p == b1)
                if e_here / 3 <= r < 2 * e_here / 3 and (quirks or not last):   # mismatch
                    rb = rng.choice([b for b in BASES if b != refb.upper()])
                    if rng.random() < variant_frac:
                        alt = rb if rng.random() < 0.7 else rng.choice(BASES)
                        alts = [alt] + ([rng.choice(BASES)] if rng.random() < 0.2 else [])
                        vcf.append([chrom, p, ".", refb.upper(), ",".join(alts)])
                else:
                    rb = refb.upper() if rng.random() > 0.02 else refb
                    if (rb in "Nn" or rb not in "ACGTacgt") and (quirks or not last):
                        rb = rng.choice(BASES)
                seq.append(rb)
                cols.append(("M", rb, refb))
                push("M", 1)
                p += 1
            if bi + 1 < len(blocks):
                push("N", blocks[bi + 1][0] - b1 - 1)
        pos = blocks[0][0]
        nm, md = _std_md(cols)
        # clips
        left_clip = right_clip = ""
        if rng.random() < 0.25:
            left_clip = "".join(rng.choice(BASES) for _ in range(rng.randrange(1, 30)))
            cigar.insert(0, ["S", len(left_clip)])
        if rng.random() < 0.25:
            right_clip = "".join(rng.choice(BASES) for _ in range(rng.randrange(1, 30)))
            cigar.append(["S", len(right_clip)])
        if quirks and rng.random() < 0.03:
            cigar.insert(0, ["H", rng.randrange(1, 20)])
        if quirks and rng.random() < 0.03:
            cigar.append(["H", rng.randrange(1, 20)])
        seq_s = left_clip + "".join(seq) + right_clip
        cig_s = "".join("%d%s" % (c, o) for o, c in cigar)
        if quirks and rng.random() < 0.03 and len(cigar) > 1 and cigar[0][0] == "M" and cigar[0][1] > 4:
            c0 = cigar[0][1]                                         # adjacent M ops (verbatim CIGAR, Q10)
            cig_s = "%dM%dM" % (2, c0 - 2) + "".join("%d%s" % (c, o) for o, c in cigar[1:])
        flag = 0 if strand == "+" else 16
        # junction tags
        ji, jm_ok = [], []
        gp = pos
        for o, c in cigar:
            if o == "N":
                ji += [gp, gp + c - 1]
            if o not in ("S", "I"):
                gp += c
        tags = ["NH:i:1", "HI:i:%d" % rng.randrange(1, 3)]
        r = rng.random()
        if r >= tagless_frac:
            if quirks and rng.random() < 0.04:
                md = md.replace("0", "", 1) if "0" in md[1:-1] else md   # missing 0 separator (Q5)
            if quirks and rng.random() < 0.02:
                md = md + "A0"                                          # MD overruns CIGAR (Q14)
            if quirks and rng.random() < 0.02:
                md = md.lower()
            tags += ["NM:i:%d" % nm, "MD:Z:" + md]
        elif r < tagless_frac / 4:
            tags += ["NM:i:%d" % nm]                                    # only one of the two (Q3)
        r = rng.random()
        ji_s = "jI:B:i," + (",".join(str(x) for x in ji) if ji else "-1")
        if r < 0.4:
            pass                                                        # neither jM nor jI
        elif r < 0.55:
            tags += [ji_s]
        elif r < 0.6 and quirks and ji:
            # trusted jI that disagrees with the CIGAR (Q3/Q17 iii): one extra junction
            tags += ["jM:B:c," + ",".join("0" for _ in range(len(ji) // 2 + 1)),
                     ji_s + ",%d,%d" % (ji[-1] + 500, ji[-1] + 900)]
        else:
            tags += ["jM:B:c," + (",".join("1" for _ in range(len(ji) // 2)) if ji else "-1"), ji_s]
        if rng.random() < 0.3:
            rng.shuffle(tags)
        qual = "*" if rng.random() < 0.5 else "".join(chr(33 + rng.randrange(40)) for _ in seq_s)
        line = "\t".join(["read%d/%d" % (seed, n), str(flag), chrom, str(pos), str(rng.choice([255, 60, 3])),
                          cig_s, "*", "0", "0", seq_s, qual] + tags)
        case.lines.append(line)
        # non-primary / unmapped companions
        r = rng.random()
        if r < 0.03:
            case.lines.append("\t".join(["unm%d/%d" % (seed, n), "4", "*", "0", "0", "*", "*", "0", "0",
                                         seq_s[:50], "*"]))
        elif r < 0.06:
            f2 = rng.choice([256, 272, 2048, 2064, 20])
            case.lines.append("\t".join(["sec%d/%d" % (seed, n), str(f2), chrom, str(pos), "0", cig_s,
                                         "*", "0", "0", "*", "*", "NH:i:2"]))
    # VCF: planted + random records, sorted
    for _ in range(len(vcf) // 2 + 5):
        chrom = rng.choice(list(case.genome))
        p = rng.randrange(10, len(case.genome[chrom]) - 10)
        ref = case.genome[chrom][p - 1].upper()
        kind = rng.randrange(4)
        if kind == 0:
            vcf.append([chrom, p, ".", ref, rng.choice(BASES)])
        elif kind == 1:
            vcf.append([chrom, p, ".", ref, ref + "".join(rng.choice(BASES) for _ in range(rng.randrange(1, 8)))])
        elif kind == 2:
            n = rng.randrange(1, 8)
            vcf.append([chrom, p, ".", case.genome[chrom][p - 1:p + n].upper(), ref])
        else:
            vcf.append([chrom, p, ".", case.genome[chrom][p - 1:p + 1].upper(), "".join(rng.choice(BASES) for _ in range(2))])
    vcf.sort(key=lambda r: (r[0], r[1]))
    case.vcf_rows = [r + [".", "PASS", "."] for r in vcf]
    if unpinned:
        _add_unpinned(rng, case)
    return case


def _add_unpinned(rng, case):
    """Equidistant annotated sites and opposite-strand-only sites next to existing junctions."""
    extra = []
    for r in case.sj_rows[: max(4, len(case.sj_rows) // 4)]:
        chrom, i0, i1, col4 = r[0], r[1], r[2], r[3]
        if col4 == 0:
            continue
        other = 3 - col4
        extra.append([chrom, i0 + 4, i1 + 4, col4] + r[4:])
        extra.append([chrom, i0 - 4, i1 - 4, col4] + r[4:])
        extra.append([chrom, i0 + 2, i1 - 2, other] + r[4:])
    case.sj_rows += extra


def make_micro_intron_case(seed, n_reads=40):
    """Reads whose junctions sit on introns of 1-10 bases next to annotated sites up to 9 bases away, so that the
    rescue drives intron lengths to zero and below (tc_plan.cuh `intron_text_rules`).  The op in
    front of the intron is an M, a D or an I, the op behind it an M, a D or an I.  Returns a Case (one contig,
    no VCF); use with maxSJOffset 5 and 10, with and without the indel stage."""
    rng = random.Random(seed)
    length = 60000
    gs = "".join(rng.choice(BASES) for _ in range(length))
    case = Case()
    case.genome = {"chr1": gs}
    p = 8000
    for r in range(n_reads):
        strand = rng.choice("+-")
        l1, l2 = rng.randrange(20, 60), rng.randrange(20, 60)
        m = rng.randrange(1, 11)
        pre, post = rng.choice("MMDI"), rng.choice("MMDI")
        ops, seq, md = [], [], []
        state = {"gp": p, "run": 0}

        def match(n_):
            ops.append(("M", n_)); seq.append(gs[state["gp"] - 1:state["gp"] - 1 + n_]); state["run"] += n_; state["gp"] += n_

        def dele(n_):
            ops.append(("D", n_)); md.append(str(state["run"]) + "^" + gs[state["gp"] - 1:state["gp"] - 1 + n_])
            state["run"] = 0; state["gp"] += n_

        def ins(n_):
            ops.append(("I", n_)); seq.append("".join(rng.choice(BASES) for _ in range(n_)))

        match(l1)
        if pre == "D":
            dele(rng.randrange(1, 9))                # longer than maxLenIndel now and then: the op survives the indel stage
        elif pre == "I":
            ins(rng.randrange(1, 9))
        i0 = state["gp"]
        ops.append(("N", m)); state["gp"] += m
        i1 = state["gp"] - 1
        if post == "D":
            dele(rng.randrange(1, 9))
        elif post == "I":
            ins(rng.randrange(1, 9))
        match(l2)
        if rng.random() < 0.4:                       # an ordinary second intron behind, shifted off its annotated site now and then
            a0 = state["gp"]; il = rng.randrange(80, 400)
            ops.append(("N", il)); state["gp"] += il
            sh = rng.choice([0, 0, 2, -3])
            case.sj_rows.append(["chr1", a0 + sh, a0 + il - 1, 1 if strand == "+" else 2, 1, 1, 5, 0, 30])
            match(rng.randrange(20, 50))
        md.append(str(state["run"]))
        cig = "".join("%d%s" % (c, o) for o, c in ops)
        nm = sum(c for o, c in ops if o in "DI")
        col4 = 1 if strand == "+" else 2
        dd, da = rng.randrange(-9, 10), rng.randrange(-9, 10)
        case.sj_rows.append(["chr1", i0 + dd, i0 + dd + 3000 + r, col4, 1, 1, 5, 0, 30])
        case.sj_rows.append(["chr1", i1 + da - 3000 - r, i1 + da, col4, 1, 1, 5, 0, 30])
        case.lines.append("\t".join(["mi%d_%d" % (seed, r), "0" if strand == "+" else "16", "chr1", str(p), "60", cig, "*", "0", "0",
                                     "".join(seq), "*", "NM:i:%d" % nm, "MD:Z:" + "".join(md)]))
        p = state["gp"] + 200
    return case

"""CPU oracle for TranscriptClean's per-read correction path.  TEST INFRASTRUCTURE ONLY.

This module restates, in plain Python over op lists, what the reference does in
`correct_transcript()` (/root/reference/TranscriptClean/TranscriptClean.py:406-460) and
everything it calls.  It exists to *check* the CUDA path: only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it.  The product package
`transcriptclean_b200` never does.

Pinning (oracle/make_golden.py, tests/test_oracle_golden.py): this restatement is compared
byte-for-byte with outputs of the unmodified reference (run in the build container under the
three stub modules in oracle/stubs/) on the reference's own fixtures and on seeded synthetic
reads; the vectors are committed under tests/golden/.

Third-party semantics the reference relies on and that are restated here:
  pyfaidx (>0.5)          Genome.fetch(): 1-based inclusive, case kept, truncated past the end,
                          FetchError when start < 1.
  pyranges (0.0.127)      nearest(): see SjTables.nearest(); the two unpinned choices
                          (strand handling, equidistant tie) are switches, defaults as in
                          oracle/stubs/pyranges.py.
  pybedtools/bedtools     VCF pre-filter is an identity for lookups (TC.py:806).

Every function cites the reference lines it follows (TC = TranscriptClean.py,
tr = transcript.py, sj = spliceJunction.py, ib = intronBound.py).
"""
import gzip
import re

NA = "NA"
# bench.py sets this to time the reference's own algorithm: one genome fetch per aligned base
# (transcript.py:296-318) instead of the whole-segment compare used to keep the tests fast.
FAITHFUL_PER_BASE = False


class FetchError(IndexError):
    pass


class Genome:
    """In-memory genome: name -> str.  pyfaidx.get_seq semantics (see module docstring)."""

    def __init__(self, seqs):
        self.seqs = seqs

    def fetch(self, name, start, end):
        if name not in self.seqs:
            raise FetchError(name)
        if start < 1:
            raise FetchError("start < 1")
        s = self.seqs[name]
        if hasattr(s, "fetch"):          # sparse contig (tests/goldens.py)
            return s.fetch(start, end)
        return s[start - 1:max(end, start - 1)]


class Options:
    """The option fields the path reads (TC.py:423-449, 375, 381)."""

    def __init__(self, maxLenIndel=5, maxSJOffset=5, correctMismatches=True, correctIndels=True,
                 correctSJs=True, primaryOnly=False, canonOnly=False, dryRun=False):
        self.maxLenIndel, self.maxSJOffset = maxLenIndel, maxSJOffset
        self.correctMismatches, self.correctIndels, self.correctSJs = \
            correctMismatches, correctIndels, correctSJs
        self.primaryOnly, self.canonOnly, self.dryRun = primaryOnly, canonOnly, dryRun


# --------------------------------------------------------------------------- reference tables

class SjTables:
    """Donor / acceptor position tables and the annotated-junction set (TC.py:666-766)."""

    def __init__(self, strand_mode="same", tie="left"):
        self.donors = {}      # (chrom, strand) and (chrom, None) -> sorted list of 0-based starts
        self.acceptors = {}
        self.annot = set()    # (chrom, start, end, strand)
        self.strand_mode, self.tie = strand_mode, tie

    @staticmethod
    def from_file(path, read_chroms, strand_mode="same", tie="left"):
        t = SjTables(strand_mode, tie)
        don, acc = set(), set()
        with open(path) as f:
            for line in f:
                p = line.strip().split("\t")
                chrom = p[0]
                if chrom not in read_chroms:                       # TC.py:685
                    continue
                start, end = int(p[1]), int(p[2])
                strand = {"1": "+", "2": "-"}.get(p[3], ".")       # TC.py:691-695
                int(p[4]); int(p[5]); p[8]                         # TC.py:697-701 (must parse / exist)
                if strand == ".":                                  # TC.py:714-715
                    continue
                first, second = (don, acc) if strand == "+" else (acc, don)   # TC.py:704-713
                first.add((chrom, start - 1, strand))
                second.add((chrom, end - 1, strand))
                t.annot.add((chrom, start, end, strand))           # TC.py:728-730
        for src, dst in ((don, t.donors), (acc, t.acceptors)):
            for chrom, s, strand in src:
                dst.setdefault((chrom, strand), []).append(s)
                dst.setdefault((chrom, None), []).append(s)
            for k in dst:
                dst[k] = sorted(set(dst[k]))
        return t

    def nearest(self, table, chrom, pos, strand):
        """Signed distance ref.Start - query.Start to the closest site (TC.py:1234-1261).
        Raises IndexError when the table has nothing on this chromosome (TC.py:1245, Q23)."""
        import bisect
        key = (chrom, strand) if self.strand_mode == "same" else (chrom, None)
        cand = table.get(key, [])
        q = pos - 1
        if not cand:
            raise IndexError("no annotated site on " + chrom)
        i = bisect.bisect_left(cand, q)
        best = cand[i] if i < len(cand) else None
        if i > 0:
            left = cand[i - 1]
            if best is None:
                best = left
            else:
                dl, dr = q - left, best - q
                if dl < dr or (dl == dr and self.tie == "left"):
                    best = left
        return best - q


class VariantTables:
    """SNP / insertion / deletion catalogues keyed as the reference keys them (TC.py:814-856)."""

    def __init__(self):
        self.snps, self.ins, self.dels = {}, {}, set()

    @staticmethod
    def from_file(path, max_len):
        v = VariantTables()
        opener = gzip.open if str(path).endswith(".gz") else open
        with opener(path, "rt") as f:
            for line in f:
                if line.startswith("#") or not line.strip():
                    continue
                p = line.rstrip("\n").split("\t")
                v.add(p[0], int(p[1]), p[3], p[4], max_len)
        return v

    def add(self, chrom, pos, ref, alt_field, max_len):
        for allele in alt_field.split(","):                        # TC.py:819-826
            if len(ref) == len(allele) == 1:                       # TC.py:829-834
                self.snps.setdefault((chrom, pos), []).append(allele)
                continue
            size = abs(len(ref) - len(allele))                     # TC.py:837-846
            if size > max_len:
                continue
            key = (chrom, pos, pos + size)
            if len(ref) < len(allele):
                self.ins.setdefault(key, []).append(allele[1:])
            elif len(ref) > len(allele):
                self.dels.add(key)


# --------------------------------------------------------------------------- small helpers

def split_cigar(cigar):
    """tr.py:579-588 / TC.py:1504-1512: two regex passes, then zip."""
    ops = re.sub("[0-9]", " ", cigar).split()
    cts = [int(x) for x in re.sub("[A-Z]", " ", cigar).split()]
    return list(zip(ops, cts))


def join_cigar(pairs):
    return "".join(str(c) + o for o, c in pairs)


def split_md(md_field):
    """tr.py:128-157: value = third ':'-separated piece; runs of digits vs non-digits."""
    val = str(md_field).split(":")[2]
    out = []
    for m in re.finditer(r"[0-9]+|[^0-9]+", val):
        tok = m.group(0)
        if tok[0].isdigit():
            out.append(["M", int(tok)])
        elif tok.startswith("^"):
            out.append(["D", len(tok) - 1])
        else:
            out.append(["X", len(tok)])
    return out


def merge_md_cigar(cigar, md_field):
    """tr.py:159-213.  IndexError escapes exactly where the reference's list indexing would."""
    cg = [list(x) for x in split_cigar(cigar)]
    md = split_md(md_field)
    out, ci, mi = [], 0, 0
    while mi < len(md) or ci < len(cg):
        cop, cct = cg[ci]                       # IndexError when CIGAR is exhausted first
        if cop in ("H", "S", "I", "N"):
            out.append((cop, cct))
            ci += 1
            continue
        mop, mct = md[mi]                       # IndexError when MD is exhausted first
        if cct < mct:
            md[mi][1] = mct - cct
            out.append((cop, cct))
            ci += 1
        elif cct > mct:
            cg[ci][1] = cct - mct
            out.append((mop, mct))
            mi += 1
        else:
            out.append((mop, mct))
            mi += 1
            ci += 1
    return out


_MOTIF = {"GTAG": 1, "CTAC": 2, "GCAG": 3, "CTGC": 4, "ATAC": 5, "GTAT": 6}   # sj.py:71-89
_COMP = str.maketrans("ACGTNacgtn*", "TGCANtgcan*")                              # tr.py:541-576


class Junction:
    """SpliceJunction + its two IntronBounds (sj.py:6-68, ib.py:4-45).  `bounds` is a mutable
    2-list shared by every shallow copy of the read, which is what produces Q18."""

    def __init__(self, chrom, start, end, strand, genome, annot):
        self.chrom, self.start, self.end, self.strand = chrom, start, end, strand
        self.bounds = [start, end]
        self.code = None
        self.canonical = None
        self.check_motif(genome, annot)

    def donor_side(self):          # sj.py:29-34
        return 0 if self.strand == "+" else 1

    def acceptor_side(self):       # sj.py:36-41
        return 1 if self.strand == "+" else 0

    def check_motif(self, genome, annot):      # sj.py:48-68, ib.py:29-45
        a = genome.fetch(self.chrom, self.bounds[0], self.bounds[0] + 1)
        b = genome.fetch(self.chrom, self.bounds[1] - 1, self.bounds[1])
        code = 20 if (self.chrom, self.start, self.end, self.strand) in annot else 0
        code += _MOTIF.get((a + b).upper(), 0)
        self.code = code
        self.canonical = code not in (0, 20)


class Read:
    """Transcript (tr.py:10-72)."""

    def __init__(self, fields, genome, annot):
        self.qname, self.flag, self.chrom = fields[0], int(fields[1]), fields[2]
        self.pos, self.mapq, self.cigar = int(fields[3]), fields[4], str(fields[5])
        self.rnext, self.pnext, self.tlen, self.seq = fields[6], fields[7], fields[8], fields[9]
        self.nm = self.md = self.jm = self.ji = ""
        other = []
        for f in fields[11:]:                                       # tr.py:33-43
            if f.startswith("NM"):
                self.nm = f
            elif f.startswith("MD"):
                self.md = f
            elif f.startswith("jM"):
                self.jm = f
            elif f.startswith("jI"):
                self.ji = f
            else:
                other.append(f)
        if self.nm == "" or self.md == "":                          # tr.py:47-48
            self.nm, self.md = nm_md_tags(self, genome)
        self.strand = "-" if self.flag in (16, 272) else "+"        # tr.py:51-53
        if self.ji == "":                                           # tr.py:56-57
            self.ji = ji_from_cigar(self)
        if "N" in self.cigar:                                       # tr.py:59-66
            vals = self.ji.split(":")[-1].split(",")[1:]
            self.junctions = []
            k = 0
            while k < len(vals):
                self.junctions.append(Junction(self.chrom, int(vals[k]), int(vals[k + 1]),
                                               self.strand, genome, annot))
                k += 2
            self.canonical = all(j.canonical for j in self.junctions)
            self.all_annot = all(j.code >= 20 for j in self.junctions)
        else:
            self.junctions, self.canonical, self.all_annot = [], True, True
        if self.jm == "":                                           # tr.py:69-70
            self.jm, self.ji = junction_tags(self)
        self.other = "\t".join(other)

    def clone(self):
        """copy.copy(): attributes rebound, junction objects shared (TC.py:420, 1205)."""
        c = Read.__new__(Read)
        c.__dict__.update(self.__dict__)
        return c

    def sam_line(self):                                             # tr.py:244-255
        parts = [self.qname, self.flag, self.chrom, self.pos, self.mapq, self.cigar, self.rnext,
                 self.pnext, self.tlen, self.seq, "*", self.other, self.nm, self.md, self.jm,
                 self.ji]
        return "\t".join(str(p) for p in parts if p != "" and p is not None).strip()

    def fasta_record(self):                                         # tr.py:257-268
        s = str(self.seq)
        if self.strand == "-":
            s = s.translate(_COMP)[::-1]
        return ">" + self.qname + "\n" + "\n".join(s[i:i + 80] for i in range(0, len(s), 80))


def nm_md_tags(read, genome):
    """tr.py:280-342 (Q8, Q9).  Whole-segment compare first; the per-base walk only runs over
    segments that contain a mismatch, which gives the same strings."""
    nm, md, run, qpos, gpos = 0, ["MD:Z:"], 0, 0, read.pos
    for op, ct in split_cigar(read.cigar):
        if op == "M":
            ref = genome.fetch(read.chrom, gpos, gpos + ct - 1) if ct > 0 else ""
            seg = read.seq[qpos:qpos + ct]
            if len(seg) < ct:
                raise IndexError("string index out of range")      # tr.py:299
            if FAITHFUL_PER_BASE:
                for i in range(ct):
                    refb = genome.fetch(read.chrom, gpos + i, gpos + i)
                    if refb == "":
                        return None, None
                    if seg[i].upper() != refb.upper():
                        md.append(str(run))
                        md.append(refb)
                        run = 0
                        nm += 1
                    else:
                        run += 1
            elif len(ref) == ct and seg.upper() == ref.upper():
                run += ct
            else:
                for i in range(ct):
                    if i >= len(ref):
                        return None, None                          # tr.py:303-304
                    if seg[i].upper() != ref[i].upper():
                        md.append(str(run))
                        md.append(ref[i])
                        run = 0
                        nm += 1
                    else:
                        run += 1
            qpos += ct
            gpos += ct
        elif op == "D":                                             # tr.py:319-331
            md.append(str(run))
            run = 0
            ref = genome.fetch(read.chrom, gpos, gpos + ct - 1)
            if ref == "":
                return None, None
            md.append("^" + ref)
            nm += ct
            gpos += ct
        elif op in ("I", "S"):                                      # tr.py:333-336
            qpos += ct
            if op == "I":
                nm += ct
        elif op in ("N", "H"):                                      # tr.py:337-338
            gpos += ct
    if run > 0:
        md.append(str(run))
    return "NM:i:" + str(nm), "".join(md)


def ji_from_cigar(read):
    """tr.py:368-393."""
    out, gpos = ["jI:B:i"], read.pos
    for op, ct in split_cigar(read.cigar):
        if op == "N":
            out += [str(gpos), str(gpos + ct - 1)]
        if op not in ("S", "I"):
            gpos += ct
    if len(out) == 1:
        out.append("-1")
    return ",".join(out)


def junction_tags(read):
    """tr.py:344-366."""
    jm, ji = ["jM:B:c"], ["jI:B:i"]
    for j in read.junctions:
        jm.append(str(j.code))
        ji += [str(j.bounds[0]), str(j.bounds[1])]
    if len(jm) == 1:
        jm.append("-1")
        ji.append("-1")
    return ",".join(jm), ",".join(ji)


def seq_matches_cigar(seq, cigar):
    """tr.py:591-605."""
    return len(seq) == sum(c for o, c in split_cigar(cigar) if o in ("M", "I", "S"))


class _CigarWriter:
    """newCIGAR / MVal / endMatch bookkeeping shared by the three correctors (TC.py:1171-1183)."""

    def __init__(self):
        self.out, self.run = [], 0

    def match(self, n):
        self.run += n

    def flush(self):
        if self.run > 0:
            self.out.append(("M", self.run))
            self.run = 0

    def op(self, o, c):
        self.flush()
        self.out.append((o, c))

    def done(self):
        self.flush()
        return join_cigar(self.out)


def new_log(fields):
    """TC.py:1558-1588 (Q1, Q12)."""
    flag, chrom = int(fields[1]), fields[2]
    mapping = "unmapped" if (chrom == "*" or flag == 4) else ("non-primary" if flag > 16 else "primary")
    return {"id": fields[0], "mapping": mapping, "cD": NA, "uD": NA, "vD": NA, "cI": NA, "uI": NA,
            "vI": NA, "cM": NA, "uM": NA, "cJ": NA, "uJ": NA}


def log_row(g):
    """TC.py:1591-1604."""
    return "\t".join(str(g[k]) for k in ("id", "mapping", "cD", "uD", "vD", "cI", "uI", "vI",
                                         "cM", "uM", "cJ", "uJ"))


def te_row(qname, ident, kind, size, status, reason):
    return "\t".join([qname, ident, kind, str(size), status, reason]) + "\n"


# --------------------------------------------------------------------------- the four stages

def fix_mismatches(read, genome, snps, log):
    """TC.py:1077-1168 (Q5, Q6, Q11)."""
    log["uM"] = log["cM"] = 0
    if not any(c in read.md.upper() for c in "ACTGN"):              # TC.py:1090
        return ""
    te, w, seq, qpos, gpos = "", _CigarWriter(), [], 0, read.pos
    for op, ct in merge_md_cigar(read.cigar, read.md):
        if op == "M":
            seq.append(read.seq[qpos:qpos + ct])
            w.match(ct)
            qpos += ct
            gpos += ct
        if op == "X":
            ident = read.chrom + "_" + str(gpos)
            if (read.chrom, gpos) in snps and read.seq[qpos] in snps[(read.chrom, gpos)]:
                te += te_row(read.qname, ident, "Mismatch", ct, "Uncorrected", "VariantMatch")
                log["uM"] += 1
                seq.append(read.seq[qpos:qpos + ct])
            else:
                te += te_row(read.qname, ident, "Mismatch", ct, "Corrected", NA)
                log["cM"] += 1
                seq.append(genome.fetch(read.chrom, gpos, gpos + ct - 1))
            w.match(ct)
            qpos += ct
            gpos += ct
        if op in ("S", "I"):
            w.op(op, ct)
            seq.append(read.seq[qpos:qpos + ct])
            qpos += ct
        if op in ("N", "H", "D"):
            w.op(op, ct)
            gpos += ct
    read.cigar, read.seq = w.done(), "".join(seq)
    read.nm, read.md = nm_md_tags(read, genome)                     # TC.py:1166
    return te


def fix_insertions(read, ins, max_len, log):
    """TC.py:859-964 (Q7: no NM/MD refresh)."""
    log["uI"] = log["cI"] = log["vI"] = 0
    ops = split_cigar(read.cigar)
    if "I" not in read.cigar:                                       # TC.py:875
        return ""
    te, w, seq, qpos, gpos = "", _CigarWriter(), [], 0, read.pos
    for op, ct in ops:
        if op == "M":
            seq.append(read.seq[qpos:qpos + ct])
            w.match(ct)
            qpos += ct
            gpos += ct
        if op == "I":
            key = (read.chrom, gpos - 1, gpos + ct - 1)
            ident = "_".join([read.chrom, str(key[1]), str(key[2])])
            if ct > max_len:
                te += te_row(read.qname, ident, "Insertion", ct, "Uncorrected", "TooLarge")
                log["uI"] += 1
                keep = True
            elif key in ins and read.seq[qpos:qpos + ct] in ins[key]:
                te += te_row(read.qname, ident, "Insertion", ct, "Uncorrected", "VariantMatch")
                log["vI"] += 1
                keep = True
            else:
                te += te_row(read.qname, ident, "Insertion", ct, "Corrected", NA)
                log["cI"] += 1
                keep = False
            if keep:
                w.op(op, ct)
                seq.append(read.seq[qpos:qpos + ct])
            qpos += ct
        if op == "S":
            w.op(op, ct)
            seq.append(read.seq[qpos:qpos + ct])
            qpos += ct
        if op in ("N", "H", "D"):
            w.op(op, ct)
            gpos += ct
    read.cigar, read.seq = w.done(), "".join(seq)
    return te


def fix_deletions(read, genome, dels, max_len, log):
    """TC.py:967-1074."""
    log["uD"] = log["cD"] = log["vD"] = 0
    ops = split_cigar(read.cigar)
    if "D" not in read.cigar:                                       # TC.py:984
        return ""
    te, w, seq, qpos, gpos = "", _CigarWriter(), [], 0, read.pos
    for op, ct in ops:
        if op == "M":
            seq.append(read.seq[qpos:qpos + ct])
            w.match(ct)
            qpos += ct
            gpos += ct
        if op == "D":
            key = (read.chrom, gpos - 1, gpos + ct - 1)
            ident = "_".join([read.chrom, str(key[1]), str(key[2])])
            if ct > max_len:
                te += te_row(read.qname, ident, "Deletion", ct, "Uncorrected", "TooLarge")
                log["uD"] += 1
                w.op(op, ct)
            elif key in dels:
                genome.fetch(read.chrom, gpos, gpos + ct - 1)       # TC.py:1016 (may raise)
                te += te_row(read.qname, ident, "Deletion", ct, "Uncorrected", "VariantMatch")
                log["vD"] += 1
                w.op(op, ct)
            else:
                te += te_row(read.qname, ident, "Deletion", ct, "Corrected", NA)
                log["cD"] += 1
                seq.append(genome.fetch(read.chrom, gpos, gpos + ct - 1))
                w.match(ct)
            gpos += ct
        if op in ("S", "I"):
            w.op(op, ct)
            seq.append(read.seq[qpos:qpos + ct])
            qpos += ct
        if op in ("N", "H"):
            w.op(op, ct)
            gpos += ct
    read.cigar, read.seq = w.done(), "".join(seq)
    read.nm, read.md = nm_md_tags(read, genome)                     # TC.py:1073
    return te


def combined_dist(d0, d1):
    """TC.py:1347-1376."""
    if (d0 < 0) != (d1 < 0):
        return abs(d0) + abs(d1)
    return abs(abs(d0) - abs(d1))


def _exon_groups(cigar, seq):
    """TC.py:1379-1412."""
    ex_c, in_c, ex_s = [], [], []
    cur_c, cur_s, q = [], "", 0
    for op, ct in split_cigar(cigar):
        if op == "N":
            ex_s.append(cur_s)
            in_c.append(ct)
            ex_c.append(cur_c)
            cur_c, cur_s = [], ""
        elif op in ("M", "S", "I"):
            cur_s += seq[q:q + ct]
            cur_c.append((op, ct))
            q += ct
        else:
            cur_c.append((op, ct))
    ex_s.append(cur_s)
    ex_c.append(cur_c)
    return ex_c, in_c, ex_s


def _edit_exon(ops, at_end, n):
    """TC.py:1515-1555: add (n >= 0) or remove (n < 0) bases at one end; removal eats ops of
    any type (Q17)."""
    ops = list(ops[::-1] if at_end else ops)
    if n >= 0:
        if ops[0][0] == "M":                  # IndexError on an empty exon (TC.py:1526)
            ops[0] = ("M", ops[0][1] + n)
        else:
            ops.insert(0, ("M", n))
    else:
        need, i = -n, 0
        while i < len(ops):
            rem = ops[i][1] - need
            if rem > 0:
                ops[i] = (ops[i][0], rem)
                break
            ops[i] = None
            if rem == 0:
                break
            need = -rem
            i += 1
        ops = [o for o in ops if o is not None]
    return ops[::-1] if at_end else ops


def _fix_side(read, k, junction, side, d, genome):
    """TC.py:1415-1479.  Mutates junction.bounds[side] *before* the length check (Q18)."""
    if d == 0:
        return
    ex_c, in_c, ex_s = _exon_groups(read.cigar, read.seq)
    if side == 0:
        exon = ex_s[k]
        if d > 0:
            end = junction.bounds[0] - 1
            ex_s[k] = exon + genome.fetch(read.chrom, end + 1, end + d)
        else:
            ex_s[k] = exon[0:d]
        junction.bounds[0] += d
        in_c[k] -= d
        ex_c[k] = _edit_exon(ex_c[k], True, d)
    else:
        exon = ex_s[k + 1]
        if d < 0:
            start = junction.bounds[1] + 1
            ex_s[k + 1] = genome.fetch(read.chrom, start + d, start - 1) + exon
        else:
            ex_s[k + 1] = exon[d:]
        junction.bounds[1] += d
        ex_c[k + 1] = _edit_exon(ex_c[k + 1], False, -d)
        in_c[k] += d
    seq = "".join(ex_s)
    cigar = ""
    for i in range(len(in_c)):
        cigar += join_cigar(ex_c[i]) + str(in_c[i]) + "N"
    cigar += join_cigar(ex_c[-1])
    if not seq_matches_cigar(seq, cigar):
        raise RuntimeError("CIGAR string and sequence are not the same length")
    read.seq, read.cigar = seq, cigar


def try_fix_junction(read, k, genome, sj, max_dist):
    """TC.py:1297-1344, 1264-1294."""
    j = read.junctions[k]
    ds, as_ = j.donor_side(), j.acceptor_side()
    d_don = sj.nearest(sj.donors, j.chrom, j.bounds[ds], j.strand)      # may raise (Q23)
    d_acc = sj.nearest(sj.acceptors, j.chrom, j.bounds[as_], j.strand)
    dist = combined_dist(d_don, d_acc)
    if dist > max_dist or abs(d_don) > 2 * max_dist or abs(d_acc) > 2 * max_dist:
        return False, "TooFarFromAnnotJn", dist
    try:
        _fix_side(read, k, j, ds, d_don, genome)
        _fix_side(read, k, j, as_, d_acc, genome)
        j.start, j.end = j.bounds                                    # sj.py:43-46
        j.check_motif(genome, sj.annot)
        read.nm, read.md = nm_md_tags(read, genome)
        read.jm, read.ji = junction_tags(read)
        read.canonical = all(x.canonical for x in read.junctions)
        read.all_annot = all(x.code >= 20 for x in read.junctions)
    except Exception:
        return False, "Other", dist
    return True, NA, dist


def fix_junctions(read, genome, sj, max_dist, log):
    """TC.py:1186-1231 (Q16-Q18)."""
    log["cJ"] = log["uJ"] = 0
    te = ""
    if read.canonical:
        return read, te
    for k in range(len(read.junctions)):
        j = read.junctions[k]
        if j.canonical:
            continue
        trial = read.clone()
        ident = "_".join([j.chrom, str(j.start), str(j.end)])
        ok, reason, dist = try_fix_junction(trial, k, genome, sj, max_dist)
        if ok:
            read = trial
            log["cJ"] += 1
        else:
            log["uJ"] += 1
        te += te_row(read.qname, ident, "NC_SJ_boundary", dist,
                     "Corrected" if ok else "Uncorrected", reason)
    return read, te


def correct_line(line, opts, genome, sj, var):
    """TC.py:406-460, 312-347.  Returns (Read or None, log dict, TE text)."""
    fields = line.split("\t")
    log = new_log(fields)
    if log["mapping"] != "primary":
        return None, log, ""
    annot = sj.annot if sj is not None else set()
    orig = Read(fields, genome, annot)
    try:
        read, te = orig.clone(), ""
        if opts.correctMismatches:
            te += fix_mismatches(read, genome, var.snps, log)
        if opts.correctIndels:
            te += fix_insertions(read, var.ins, opts.maxLenIndel, log)
            te += fix_deletions(read, genome, var.dels, opts.maxLenIndel, log)
        if len(annot) > 0 and opts.correctSJs:
            read, te2 = fix_junctions(read, genome, sj, opts.maxSJOffset, log)
            te += te2
    except Exception:
        return orig, log, ""                                        # Q13
    return read, log, te


def correct_lines(lines, opts, genome, sj=None, var=None):
    """batch_correct (TC.py:350-403, Q19).  Returns dict of the four output texts."""
    var = var or VariantTables()
    sam, fa, lg, te = [], [], [], []
    for line in lines:
        read, log, tes = correct_line(line, opts, genome, sj, var)
        if read is None:
            lg.append(log_row(log) + "\n")
            if not opts.primaryOnly and not opts.canonOnly:
                sam.append(line + "\n")
        else:
            if not opts.canonOnly or read.canonical or read.all_annot:
                sam.append(read.sam_line() + "\n")
                fa.append(read.fasta_record() + "\n")
            lg.append(log_row(log) + "\n")
            te.append(tes)
    return {"sam": "".join(sam), "fa": "".join(fa), "log": "".join(lg), "te": "".join(te)}


def dry_run_lines(lines, genome):
    """TC.py:1615-1678 (positional inventory, no correction)."""
    lg, te = [], []
    for line in lines:
        fields = line.split("\t")
        log = new_log(fields)
        if log["mapping"] != "primary":
            lg.append(log_row(log) + "\n")
            continue
        read = Read(fields, genome, set())
        log["uD"] = log["uI"] = log["uM"] = 0
        gpos = read.pos
        for op, ct in merge_md_cigar(read.cigar, read.md):
            if op == "M":
                gpos += ct
            if op == "X":
                te.append(te_row(read.qname, read.chrom + "_" + str(gpos), "Mismatch", ct,
                                 "Uncorrected", "DryRun"))
                log["uM"] += 1
                gpos += ct
            if op == "D":
                te.append(te_row(read.qname, "_".join([read.chrom, str(gpos - 1), str(gpos + ct - 1)]),
                                 "Deletion", ct, "Uncorrected", "DryRun"))
                log["uD"] += 1
                gpos += ct
            if op == "I":
                te.append(te_row(read.qname, "_".join([read.chrom, str(gpos - 1), str(gpos + ct - 1)]),
                                 "Insertion", ct, "Uncorrected", "DryRun"))
                log["uI"] += 1
            if op in ("N", "H"):
                gpos += ct
        lg.append(log_row(log) + "\n")
    return {"sam": "", "fa": "", "log": "".join(lg), "te": "".join(te)}

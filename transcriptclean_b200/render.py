"""Text rendering of the engine's columnar output: clean SAM, FASTA, .log and .TE.log rows.

Follows printableSAM / printableFa (transcript.py:244-268), reverseComplement
(transcript.py:541-576), create_log_string (TranscriptClean.py:1591-1604), the TE row formats
(TC.py:911-936, 1017-1046, 1119-1134, 1227-1228, 1650-1668) and the emission rules of
batch_correct (TC.py:373-385, Q19).
"""
from . import engine as E

_COMP = bytes.maketrans(b"ACGTNacgtn*", b"TGCANtgcan*")
_KNOWN = set(b"ACGTNacgtn*")
TE_TYPE = ("Mismatch", "Insertion", "Deletion", "NC_SJ_boundary")
TE_STATUS = ("Corrected", "Uncorrected")
TE_REASON = ("NA", "VariantMatch", "TooLarge", "TooFarFromAnnotJn", "Other", "DryRun")
CLASS_NAME = ("primary", "unmapped", "non-primary")


def reverse_complement(seq_bytes, warn=None):
    if warn is not None:
        for b in set(seq_bytes) - _KNOWN:
            for _ in range(seq_bytes.count(bytes([b]))):
                warn("Warning: reverse complement function encountered unknown base '%s'" % chr(b))
    return seq_bytes.translate(_COMP)[::-1]


def log_row(qname, cls, counters):
    return "\t".join([qname, CLASS_NAME[cls]] + ["NA" if c < 0 else str(int(c)) for c in counters])


def te_rows(qname, chrom, te):
    out = []
    for r in te:
        t = int(r["type"])
        ident = "%s_%d" % (chrom, r["a"]) if t == 0 else "%s_%d_%d" % (chrom, r["a"], r["b"])
        out.append("\t".join([qname, ident, TE_TYPE[t], str(int(r["size"])), TE_STATUS[int(r["status"])],
                              TE_REASON[int(r["reason"])]]) + "\n")
    return "".join(out)


class ReadView:
    """One corrected read, with the attribute names of the reference's Transcript object."""

    def __init__(self, pb, res, i):
        st = int(res.status[i])
        f = pb.fields[i]
        nm_in, md_in, jm_in, ji_in = pb.tag_str[i]
        self.QNAME, self.FLAG, self.CHROM, self.POS, self.MAPQ = f[0], int(f[1]), f[2], int(f[3]), f[4]
        self.RNEXT, self.PNEXT, self.TLEN, self.QUAL = f[6], f[7], f[8], "*"
        self.strand = "-" if self.FLAG == 16 else "+"
        self.SEQ = res.read_seq(i).decode("ascii")
        if st & E.ST_CIGAR_REWRITTEN:
            a, b = res.cig_off[i], res.cig_off[i + 1]
            self.CIGAR = "".join("%d%s" % (l, chr(o).upper()) for o, l in zip(res.cig_op[a:b], res.cig_len[a:b]))   # 'd' = "D-" (tc_plan.cuh)
        else:
            self.CIGAR = pb.cigar_str[i]
        if st & E.ST_NMMD_DEVICE:
            if st & E.ST_NMMD_NONE:
                self.NM = self.MD = None
            else:
                self.NM = "NM:i:%d" % res.nm[i]
                self.MD = "MD:Z:" + res.md[res.md_off[i]:res.md_off[i] + res.md_len[i]].tobytes().decode("ascii")
        else:
            self.NM, self.MD = nm_in, md_in
        a, b = res.jn_off[i], res.jn_off[i + 1]
        if st & E.ST_JM_DEVICE:
            self.jM = "jM:B:c," + (",".join(str(int(x)) for x in res.jm[a:b]) if b > a else "-1")
        else:
            self.jM = jm_in
        if st & E.ST_JI_DEVICE:
            self.jI = "jI:B:i," + (",".join(str(int(x)) for x in res.ji[2 * a:2 * b]) if b > a else "-1")
        else:
            self.jI = ji_in
        self.otherFields = pb.other[i]
        self.isCanonical = bool(st & E.ST_CANONICAL)
        self.allJnsAnnotated = bool(st & E.ST_ALL_ANNOT)

    def printableSAM(self):
        parts = [self.QNAME, self.FLAG, self.CHROM, self.POS, self.MAPQ, self.CIGAR, self.RNEXT, self.PNEXT,
                 self.TLEN, self.SEQ, self.QUAL, self.otherFields, self.NM, self.MD, self.jM, self.jI]
        return "\t".join(str(p) for p in parts if p != "" and p is not None).strip()

    def printableFa(self, warn=None):
        s = self.SEQ.encode("ascii")
        if self.strand == "-":
            s = reverse_complement(s, warn)
        s = s.decode("ascii")
        return ">" + self.QNAME + "\n" + "\n".join(s[i:i + 80] for i in range(0, len(s), 80))


def check_errors(pb, res):
    bad = (res.status & E.ST_ERR_MERGE) != 0       # (ST_ERR_INTRON is informational: that junction stayed uncorrected)
    if bad.any():
        i = int(bad.argmax())
        name = pb.lines[i].split("\t")[0]
        raise E.EngineError("read %s: MD tag and CIGAR disagree (the reference's --dryRun crashes on this read)" % name)


def render_batch(pb, res, primary_only=False, canon_only=False, warn=None):
    """The four output texts of batch_correct for one batch (TC.py:350-403)."""
    check_errors(pb, res)
    sam, fa, lg, te = [], [], [], []
    for i in range(pb.n):
        st = int(res.status[i])
        cls = st & E.ST_CLASS_MASK
        qname = pb.lines[i].split("\t", 1)[0]
        if cls != 0:
            lg.append(log_row(qname, cls, res.counters[i]) + "\n")
            if not primary_only and not canon_only:
                sam.append(pb.lines[i] + "\n")
            continue
        r = ReadView(pb, res, i)
        if not canon_only or r.isCanonical or r.allJnsAnnotated:
            sam.append(r.printableSAM() + "\n")
            fa.append(r.printableFa(warn) + "\n")
        lg.append(log_row(qname, 0, res.counters[i]) + "\n")
        te.append(te_rows(qname, r.CHROM, res.te[res.te_off[i]:res.te_off[i + 1]]))
    return {"sam": "".join(sam), "fa": "".join(fa), "log": "".join(lg), "te": "".join(te)}


def render_dry(pb, res):
    check_errors(pb, res)
    lg, te = [], []
    for i in range(pb.n):
        cls = int(res.status[i]) & E.ST_CLASS_MASK
        qname = pb.lines[i].split("\t", 1)[0]
        lg.append(log_row(qname, cls, res.counters[i]) + "\n")
        if cls == 0:
            te.append(te_rows(qname, pb.fields[i][2], res.te[res.te_off[i]:res.te_off[i + 1]]))
    return {"sam": "", "fa": "", "log": "".join(lg), "te": "".join(te)}

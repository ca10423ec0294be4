"""SAM records -> columnar host buffers (the tc_batch_in of include/tc_b200.h).

Mirrors the field handling of Transcript.__init__ (transcript.py:10-72) and init_log_info
(TranscriptClean.py:1558-1588).  Records the reference cannot process without crashing its
worker (malformed CIGAR on a primary record, SEQ/CIGAR length mismatch,
alignment running past the contig, odd jI list ...) raise SamFormatError with the read name
instead of undefined behaviour.
"""
import re

import numpy as np

_CIGAR_RE = re.compile(r"([0-9]+)([A-Z])")
_CIGAR_FULL = re.compile(r"^(?:[0-9]+[A-Z])+$")

TAG_NM, TAG_MD, TAG_JM, TAG_JI = 1, 2, 4, 8


class SamFormatError(ValueError):
    pass


class ParsedBatch:
    """Columnar arrays + the per-read strings the renderer passes through verbatim."""

    def __init__(self):
        self.n = 0
        self.lines = []        # original line (non-primary pass-through, Q19)
        self.fields = []       # first 9 mandatory fields as given
        self.other = []        # optional tags other than NM/MD/jM/jI, tab-joined (tr.py:33-43)
        self.tag_str = []      # (NM, MD, jM, jI) field strings or ""
        self.cigar_str = []
        self.cols = {}


def read_class(flag, chrom):
    """TC.py:1570-1575: 1 unmapped, 2 non-primary, 0 primary."""
    if chrom == "*" or flag == 4:
        return 1
    if flag > 16:
        return 2
    return 0


def parse_lines(lines, genome):
    """lines: SAM alignment lines (no header, no trailing newline); genome: GenomeImage."""
    pb = ParsedBatch()
    n = len(lines)
    flag = np.zeros(n, dtype=np.int32)
    chrom = np.full(n, -1, dtype=np.int32)
    pos = np.zeros(n, dtype=np.int32)
    tags = np.zeros(n, dtype=np.uint8)
    seq_parts, md_parts = [], []
    seq_off = np.zeros(n + 1, dtype=np.int64)
    cig_off = np.zeros(n + 1, dtype=np.int64)
    md_off = np.zeros(n + 1, dtype=np.int64)
    ji_off = np.zeros(n + 1, dtype=np.int64)
    cig_op, cig_len, ji_vals = [], [], []
    so = co = mo = jo = 0
    for i, line in enumerate(lines):
        f = line.split("\t")
        if len(f) < 11:
            raise SamFormatError("SAM record with fewer than 11 fields: %r" % line[:80])
        qname = f[0]
        try:
            fl = int(f[1])
        except ValueError:
            raise SamFormatError("read %s: FLAG is not an integer" % qname)
        flag[i] = fl
        cls = read_class(fl, f[2])
        pb.lines.append(line)
        if cls != 0:
            pb.fields.append(None); pb.other.append(""); pb.tag_str.append(("", "", "", "")); pb.cigar_str.append("")
            seq_off[i + 1], cig_off[i + 1], md_off[i + 1], ji_off[i + 1] = so, co, mo, jo
            chrom[i] = -1 if cls == 1 else 0      # the class only depends on RNAME == '*' and FLAG (TC.py:1570)
            continue
        if f[2] not in genome.index:
            raise SamFormatError("read %s: chromosome %s is not in the reference genome" % (qname, f[2]))
        c = genome.index[f[2]]
        chrom[i] = c
        try:
            p = int(f[3])
        except ValueError:
            raise SamFormatError("read %s: POS is not an integer" % qname)
        cigar, seq = f[5], f[9]
        if not _CIGAR_FULL.match(cigar):
            raise SamFormatError("read %s: unsupported CIGAR %r" % (qname, cigar[:40]))
        ops = _CIGAR_RE.findall(cigar)
        qlen = 0
        gend = p
        for ct, op in ops:
            ct = int(ct)
            if ct < 1:
                raise SamFormatError("read %s: zero-length CIGAR operation" % qname)
            cig_op.append(ord(op)); cig_len.append(ct)
            if op in "MIS":
                qlen += ct
            if op in "MDNH":
                gend += ct
        if qlen != len(seq):
            raise SamFormatError("read %s: SEQ length %d does not match CIGAR (%d)" % (qname, len(seq), qlen))
        if p < 1 or gend - 1 > int(genome.chrom_len[c]):
            raise SamFormatError("read %s: alignment runs off contig %s" % (qname, f[2]))
        nm = md = jm = ji = ""
        other = []
        for t in f[11:]:
            if t.startswith("NM"):
                nm = t
            elif t.startswith("MD"):
                md = t
            elif t.startswith("jM"):
                jm = t
            elif t.startswith("jI"):
                ji = t
            else:
                other.append(t)
        tb = 0
        if nm:
            tb |= TAG_NM
        if md:
            tb |= TAG_MD
            parts = md.split(":")
            if len(parts) < 3:
                raise SamFormatError("read %s: malformed MD tag" % qname)
            if nm:
                mdv = parts[2].encode("ascii")
                md_parts.append(mdv); mo += len(mdv)
        if jm:
            tb |= TAG_JM
        if ji:
            tb |= TAG_JI
            if "N" in cigar:
                try:
                    vals = [int(x) for x in ji.split(":")[-1].split(",")[1:]]
                except ValueError:
                    raise SamFormatError("read %s: malformed jI tag" % qname)
                if len(vals) % 2 or any(v < 2 or v > int(genome.chrom_len[c]) for v in vals):
                    raise SamFormatError("read %s: jI tag does not describe intron pairs on the contig" % qname)
                ji_vals.extend(vals); jo += len(vals)
        tags[i] = tb
        pos[i] = p
        seq_parts.append(seq.encode("ascii")); so += len(seq)
        co += len(ops)
        seq_off[i + 1], cig_off[i + 1], md_off[i + 1], ji_off[i + 1] = so, co, mo, jo
        pb.fields.append(f[:9]); pb.other.append("\t".join(other)); pb.tag_str.append((nm, md, jm, ji))
        pb.cigar_str.append(cigar)
    pb.n = n
    pb.cols = {
        "flag": flag, "chrom": chrom, "pos": pos, "tags": tags,
        "seq_off": seq_off, "seq": np.frombuffer(b"".join(seq_parts), dtype=np.uint8).copy(),
        "cig_off": cig_off, "cig_op": np.asarray(cig_op, dtype=np.uint8), "cig_len": np.asarray(cig_len, dtype=np.int32),
        "md_off": md_off, "md": np.frombuffer(b"".join(md_parts), dtype=np.uint8).copy(),
        "ji_off": ji_off, "ji": np.asarray(ji_vals, dtype=np.int32),
    }
    return pb

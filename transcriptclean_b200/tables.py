"""Sorted lookup tables for the device, built from the STAR SJ file and the VCF.

Host-side replacement of processSpliceAnnotation (TranscriptClean.py:666-766) and processVCF
(TranscriptClean.py:769-856): same table *contents*, no BED / BAM temp files, no pyranges /
pybedtools.  The bedtools pre-filter of the VCF (TC.py:781-806) never changes a lookup result
and is not reproduced.
"""
import gzip
import warnings

import numpy as np


class SjTables:
    def __init__(self):
        e = np.zeros(0, dtype=np.uint64)
        self.don_s = self.acc_s = self.don_u = self.acc_u = self.ann_k1 = self.ann_k2 = e
        self.n_annot = 0

    def __len__(self):
        return self.n_annot


def build_sj_tables(path, read_chroms, chrom_index):
    """STAR SJ.out.tab -> donor / acceptor tables + annotated set.  Rows on chromosomes absent
    from `read_chroms` are dropped (TC.py:685); strand column 0 rows are skipped (TC.py:714);
    column 6 ("annotated") is ignored, as in the reference (Q24)."""
    don, acc, ann = set(), set(), set()
    with open(path) as f:
        for line in f:
            p = line.strip().split("\t")
            chrom = p[0]
            if chrom not in read_chroms:
                continue
            start, end = int(p[1]), int(p[2])
            strand = 0 if p[3] == "1" else (1 if p[3] == "2" else -1)
            int(p[4]); int(p[5]); p[8]        # the reference parses / touches these (TC.py:697-701)
            if strand < 0 or chrom not in chrom_index:
                continue
            c = chrom_index[chrom]
            first, second = (don, acc) if strand == 0 else (acc, don)     # TC.py:704-713
            first.add((c, strand, start - 1))
            second.add((c, strand, end - 1))
            ann.add((c, start, end, strand))
    t = SjTables()

    def keys_s(s):
        return np.asarray(sorted((c << 33) | (st << 32) | p for c, st, p in s), dtype=np.uint64)

    def keys_u(s):
        return np.asarray(sorted(set((c << 32) | p for c, st, p in s)), dtype=np.uint64)

    t.don_s, t.acc_s, t.don_u, t.acc_u = keys_s(don), keys_s(acc), keys_u(don), keys_u(acc)
    a = sorted(((c << 32) | s, (e << 1) | st) for c, s, e, st in ann)
    t.ann_k1 = np.asarray([x[0] for x in a], dtype=np.uint64)
    t.ann_k2 = np.asarray([x[1] for x in a], dtype=np.uint64)
    t.n_annot = len(a)
    if len(don) == 0 or len(acc) == 0:                                    # TC.py:761-764
        warnings.warn("Warning: No splice donors or acceptors found on chromosomes: %s. If this is "
                      "unexpected, check your SJ annotation file." % (", ".join(read_chroms)))
    return t


class VariantTables:
    def __init__(self):
        self.snp_key = np.zeros(0, dtype=np.uint64)
        self.snp_alt = np.zeros(0, dtype=np.uint8)
        self.ins_key = np.zeros(0, dtype=np.uint64)
        self.ins_size = np.zeros(0, dtype=np.int32)
        self.ins_aoff = np.zeros(1, dtype=np.int64)
        self.ins_bytes = np.zeros(0, dtype=np.uint8)
        self.del_key = np.zeros(0, dtype=np.uint64)
        self.del_size = np.zeros(0, dtype=np.int32)


def build_variant_tables(path, max_len, chrom_index):
    """VCF (plain or .gz) -> SNP / insertion / deletion tables keyed as TC.py:814-854 (Q11, Q25)."""
    snps, ins, dels = [], [], set()
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rt") as f:
        for line in f:
            if line.startswith("#") or not line.strip():
                continue
            p = line.rstrip("\n").split("\t")
            chrom, pos, ref = p[0], int(p[1]), p[3]
            if chrom not in chrom_index:
                continue
            c = chrom_index[chrom]
            for allele in p[4].split(","):
                if len(ref) == len(allele) == 1:
                    snps.append(((c << 32) | pos, ord(allele)))
                    continue
                size = abs(len(ref) - len(allele))
                if size > max_len:
                    continue
                if len(ref) < len(allele):
                    ins.append(((c << 32) | pos, size, allele[1:].encode("ascii")))
                elif len(ref) > len(allele):
                    dels.add(((c << 32) | pos, size))
    v = VariantTables()
    snps.sort(key=lambda x: x[0])
    v.snp_key = np.asarray([x[0] for x in snps], dtype=np.uint64)
    v.snp_alt = np.asarray([x[1] for x in snps], dtype=np.uint8)
    ins.sort(key=lambda x: (x[0], x[1]))
    v.ins_key = np.asarray([x[0] for x in ins], dtype=np.uint64)
    v.ins_size = np.asarray([x[1] for x in ins], dtype=np.int32)
    lens = np.asarray([len(x[2]) for x in ins], dtype=np.int64)
    v.ins_aoff = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    v.ins_bytes = np.frombuffer(b"".join(x[2] for x in ins), dtype=np.uint8).copy()
    d = sorted(dels)
    v.del_key = np.asarray([x[0] for x in d], dtype=np.uint64)
    v.del_size = np.asarray([x[1] for x in d], dtype=np.int32)
    return v

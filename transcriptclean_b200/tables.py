"""Sorted lookup tables for the device, built from the STAR SJ file and the VCF.

Host-side replacement of processSpliceAnnotation (TranscriptClean.py:666-766) and processVCF
(TranscriptClean.py:769-856): same table *contents*, no BED / BAM temp files, no pyranges /
pybedtools.  The bedtools pre-filter of the VCF (TC.py:781-806) only narrows the catalogue to
the reads' neighbourhood; here it is narrowed to the reads' chromosomes (`read_chroms`), which
never changes a lookup result either.

`build_sj_tables` / `build_variant_tables` parse the files in the library's multi-threaded C++
(csrc/tc_tables.cpp: a 5 M-record catalogue in a few seconds); `build_*_py` are the same logic in
Python, kept as the readable twin the tests compare with.
"""
import ctypes as C
import gzip
import os
import warnings

import numpy as np

_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        from . import engine
        _LIB = engine.load_library()
    return _LIB


def _names(chrom_index):
    names = sorted(chrom_index, key=lambda k: chrom_index[k])
    assert [chrom_index[n] for n in names] == list(range(len(names)))
    return names, (C.c_char_p * len(names))(*[n.encode() for n in names])


def _use_mask(names, read_chroms):
    if read_chroms is None:
        return None
    return np.asarray([1 if n in read_chroms else 0 for n in names], dtype=np.uint8)


def _read_bytes(path):
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as f:
        return f.read()


def _copy(ptr, dtype, count):
    if not ptr or count == 0:
        return np.zeros(0, dtype=dtype)
    return np.frombuffer((C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr), dtype=dtype, count=count).copy()


class SjTables:
    def __init__(self):
        e = np.zeros(0, dtype=np.uint64)
        self.don_s = self.acc_s = self.don_u = self.acc_u = self.ann_k1 = self.ann_k2 = e
        self.n_annot = 0

    def __len__(self):
        return self.n_annot


def build_sj_tables(path, read_chroms, chrom_index, lib=None):
    """STAR SJ.out.tab -> donor / acceptor tables + annotated set.  Rows on chromosomes absent
    from `read_chroms` are dropped (TC.py:685); strand column 0 rows are skipped (TC.py:714);
    column 6 ("annotated") is ignored, as in the reference (Q24).  Parsed by tc_sj_parse (C++)."""
    lib = lib or _lib()
    names, cnames = _names(chrom_index)
    use = _use_mask(names, set(read_chroms))
    text = _read_bytes(path)
    err = C.create_string_buffer(512)
    h = lib.tc_sj_parse(text, len(text), len(names), cnames, use.ctypes.data if len(names) else None, err, 512)
    if not h:
        raise ValueError(err.value.decode())
    try:
        n = [C.c_int64() for _ in range(5)]
        p = [C.c_void_p() for _ in range(6)]
        lib.tc_sj_tables(h, C.byref(n[0]), C.byref(p[0]), C.byref(n[1]), C.byref(p[1]), C.byref(n[2]), C.byref(p[2]),
                         C.byref(n[3]), C.byref(p[3]), C.byref(n[4]), C.byref(p[4]), C.byref(p[5]))
        t = SjTables()
        t.don_s, t.acc_s = _copy(p[0].value, np.uint64, n[0].value), _copy(p[1].value, np.uint64, n[1].value)
        t.don_u, t.acc_u = _copy(p[2].value, np.uint64, n[2].value), _copy(p[3].value, np.uint64, n[3].value)
        t.ann_k1, t.ann_k2 = _copy(p[4].value, np.uint64, n[4].value), _copy(p[5].value, np.uint64, n[4].value)
        t.n_annot = int(n[4].value)
    finally:
        lib.tc_sj_free(h)
    if len(t.don_s) == 0 or len(t.acc_s) == 0:                            # TC.py:761-764
        warnings.warn("Warning: No splice donors or acceptors found on chromosomes: %s. If this is "
                      "unexpected, check your SJ annotation file." % (", ".join(read_chroms)))
    return t


def build_sj_tables_py(path, read_chroms, chrom_index):
    """The Python twin of tc_sj_parse (same tables; see build_sj_tables)."""
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


def build_variant_tables(path, max_len, chrom_index, read_chroms=None, lib=None, n_threads=None):
    """VCF (plain or .gz) -> SNP / insertion / deletion tables keyed as TC.py:814-854 (Q11, Q25).  Parsed by
    tc_vcf_parse (C++, multi-threaded); `read_chroms` (optional) keeps only the contigs reads map to."""
    lib = lib or _lib()
    names, cnames = _names(chrom_index)
    use = _use_mask(names, None if read_chroms is None else set(read_chroms))
    text = _read_bytes(path)
    err = C.create_string_buffer(512)
    h = lib.tc_vcf_parse(text, len(text), len(names), cnames, use.ctypes.data if use is not None and len(names) else None, int(max_len),
                         int(n_threads or min(32, os.cpu_count() or 1)), err, 512)
    if not h:
        raise ValueError(err.value.decode())
    try:
        n = [C.c_int64() for _ in range(3)]
        p = [C.c_void_p() for _ in range(8)]
        lib.tc_vcf_tables(h, C.byref(n[0]), C.byref(p[0]), C.byref(p[1]), C.byref(n[1]), C.byref(p[2]), C.byref(p[3]), C.byref(p[4]),
                          C.byref(p[5]), C.byref(n[2]), C.byref(p[6]), C.byref(p[7]))
        v = VariantTables()
        v.snp_key, v.snp_alt = _copy(p[0].value, np.uint64, n[0].value), _copy(p[1].value, np.uint8, n[0].value)
        v.ins_key, v.ins_size = _copy(p[2].value, np.uint64, n[1].value), _copy(p[3].value, np.int32, n[1].value)
        v.ins_aoff = _copy(p[4].value, np.int64, n[1].value + 1)
        v.ins_bytes = _copy(p[5].value, np.uint8, int(v.ins_aoff[-1]))
        v.del_key, v.del_size = _copy(p[6].value, np.uint64, n[2].value), _copy(p[7].value, np.int32, n[2].value)
        v.n_records = int(lib.tc_vcf_n_records(h))
    finally:
        lib.tc_vcf_free(h)
    return v


def build_variant_tables_py(path, max_len, chrom_index):
    """The Python twin of tc_vcf_parse (same tables up to the order of rows that share a key; see build_variant_tables)."""
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

"""Python wrapper of tools/synth_gen.c: the synthetic workloads named in BASELINE.json
(`configs`), produced directly as the columnar buffers of include/tc_b200.h.  Bench / test
tooling - not part of the correction path."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libtcsynth.so")


class _Params(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_reads", C.c_int64), ("mean_len", C.c_int32), ("sd_len", C.c_int32),
                ("n_exons", C.c_int32), ("err", C.c_double), ("p_mm", C.c_double), ("p_del", C.c_double),
                ("p_ins", C.c_double), ("indel_geometric", C.c_int32), ("jn_shift_frac", C.c_double),
                ("jn_shift_max", C.c_int32), ("nonprimary_frac", C.c_double), ("with_nm_md", C.c_int32)]


class _Genes(C.Structure):
    _fields_ = [("n_genes", C.c_int64), ("gene_exon_off", C.POINTER(C.c_int64)), ("gene_strand", C.POINTER(C.c_int32)),
                ("gene_chrom", C.POINTER(C.c_int32)), ("exon_start", C.POINTER(C.c_int64)),
                ("exon_end", C.POINTER(C.c_int64)), ("n_exons_total", C.c_int64)]


class _Reads(C.Structure):
    _fields_ = [("n", C.c_int64), ("flag", C.c_void_p), ("chrom", C.c_void_p), ("pos", C.c_void_p), ("tags", C.c_void_p),
                ("seq_off", C.c_void_p), ("seq", C.c_void_p), ("cig_off", C.c_void_p), ("cig_op", C.c_void_p),
                ("cig_len", C.c_void_p), ("md_off", C.c_void_p), ("md", C.c_void_p), ("nm", C.c_void_p)]


def _lib():
    if not os.path.exists(_SO):
        import subprocess
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", _SO,
                               os.path.join(_HERE, "synth_gen.c"), "-lm"])
    lib = C.CDLL(_SO)
    lib.tcsynth_genome.argtypes = [C.c_uint64, C.c_int64, C.c_void_p]
    lib.tcsynth_genes.argtypes = [C.c_uint64, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int32, C.c_double,
                                  C.POINTER(_Genes)]
    lib.tcsynth_reads.argtypes = [C.POINTER(_Params), C.c_void_p, C.c_int64, C.POINTER(_Genes), C.POINTER(_Reads)]
    lib.tcsynth_free_reads.argtypes = [C.POINTER(_Reads)]
    lib.tcsynth_free_genes.argtypes = [C.POINTER(_Genes)]
    return lib


SHAPES = {
    # BASELINE.json configs[1]: PacBio Iso-seq, ~2 kb, ~1 % error, 6 exons
    "pacbio": dict(mean_len=2000, sd_len=300, n_exons=6, err=0.01, p_mm=0.5, p_del=0.25, p_ins=0.25,
                   indel_geometric=0, jn_shift_frac=0.05, jn_shift_max=8, nonprimary_frac=0.02),
    # BASELINE.json configs[3]: ONT, ~1.5 kb, ~8 % error, frequent microindels and noncanonical junctions
    "ont": dict(mean_len=1500, sd_len=400, n_exons=5, err=0.08, p_mm=1 / 3, p_del=1 / 3, p_ins=1 / 3,
                indel_geometric=1, jn_shift_frac=0.20, jn_shift_max=12, nonprimary_frac=0.02),
}


class Workload:
    """genome (ASCII uint8), gene models, SJ rows, and a generator of read batches."""

    def __init__(self, seed=20261019, chrom_len=248956422, n_genes=3000, chrom="chr1", minus_frac=0.10, min_exons=6):
        self.lib = _lib()
        self.seed, self.chrom, self.chrom_len = seed, chrom, chrom_len
        self.genome = np.empty(chrom_len, dtype=np.uint8)
        self.lib.tcsynth_genome(seed, chrom_len, self.genome.ctypes.data)
        self.genes = _Genes()
        self.lib.tcsynth_genes(seed, self.genome.ctypes.data, chrom_len, 0, n_genes, min_exons, minus_frac,
                               C.byref(self.genes))

    def sj_rows(self, unannotated_frac=0.05):
        """STAR SJ.out.tab rows (9 columns) for the planted introns."""
        g = self.genes
        off = np.ctypeslib.as_array(g.gene_exon_off, (g.n_genes + 1,))
        st = np.ctypeslib.as_array(g.exon_start, (g.n_exons_total,))
        en = np.ctypeslib.as_array(g.exon_end, (g.n_exons_total,))
        strand = np.ctypeslib.as_array(g.gene_strand, (g.n_genes,))
        rng = np.random.default_rng(self.seed)
        rows = []
        for k in range(g.n_genes):
            for e in range(off[k], off[k + 1] - 1):
                if rng.random() < unannotated_frac:
                    continue
                rows.append([self.chrom, int(en[e]) + 1, int(st[e + 1]) - 1, 2 if strand[k] else 1, 1, 1, 10, 0, 30])
        return rows

    def genome_image(self):
        from transcriptclean_b200.genome import GenomeImage
        g = GenomeImage([self.chrom], [self.chrom_len])
        g._put(0, self.genome)
        return g

    def reads(self, shape, n_reads, seed_offset=0, with_nm_md=True):
        p = _Params(seed=self.seed + 1000 * seed_offset + 1, n_reads=n_reads, with_nm_md=int(with_nm_md), **SHAPES[shape])
        out = _Reads()
        self.lib.tcsynth_reads(C.byref(p), self.genome.ctypes.data, self.chrom_len, C.byref(self.genes), C.byref(out))
        n = out.n

        def arr(ptr, dtype, count):
            if count == 0:
                return np.zeros(0, dtype=dtype)
            buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
            return np.frombuffer(buf, dtype=dtype, count=count).copy()

        seq_off, cig_off, md_off = arr(out.seq_off, np.int64, n + 1), arr(out.cig_off, np.int64, n + 1), arr(out.md_off, np.int64, n + 1)
        cols = {
            "flag": arr(out.flag, np.int32, n), "chrom": arr(out.chrom, np.int32, n), "pos": arr(out.pos, np.int32, n),
            "tags": arr(out.tags, np.uint8, n), "seq_off": seq_off, "seq": arr(out.seq, np.uint8, int(seq_off[-1])),
            "cig_off": cig_off, "cig_op": arr(out.cig_op, np.uint8, int(cig_off[-1])),
            "cig_len": arr(out.cig_len, np.int32, int(cig_off[-1])), "md_off": md_off, "md": arr(out.md, np.uint8, int(md_off[-1])),
            "ji_off": np.zeros(n + 1, dtype=np.int64), "ji": np.zeros(0, dtype=np.int32),
        }
        nm = arr(out.nm, np.int32, n)
        self.lib.tcsynth_free_reads(C.byref(out))
        return cols, nm

    def sam_lines(self, cols, nm, lo, hi):
        """SAM text of reads [lo, hi) of a columnar batch (for the oracle / the CLI)."""
        lines = []
        for i in range(lo, hi):
            fl, ch = int(cols["flag"][i]), int(cols["chrom"][i])
            name = "r%d" % i
            if ch < 0 or fl == 4:
                lines.append("\t".join([name, str(fl), "*", "0", "0", "*", "*", "0", "0", "*", "*"]))
                continue
            if fl > 16:
                lines.append("\t".join([name, str(fl), self.chrom, "1", "0", "10M", "*", "0", "0", "*", "*"]))
                continue
            a, b = cols["cig_off"][i], cols["cig_off"][i + 1]
            cigar = "".join("%d%s" % (l, chr(o)) for o, l in zip(cols["cig_op"][a:b], cols["cig_len"][a:b]))
            seq = cols["seq"][cols["seq_off"][i]:cols["seq_off"][i + 1]].tobytes().decode()
            f = [name, str(fl), self.chrom, str(int(cols["pos"][i])), "60", cigar, "*", "0", "0", seq, "*"]
            if cols["tags"][i] & 3 == 3:
                md = cols["md"][cols["md_off"][i]:cols["md_off"][i + 1]].tobytes().decode()
                f += ["NM:i:%d" % nm[i], "MD:Z:" + md]
            lines.append("\t".join(f))
        return lines

    def write_sam(self, path, cols, nm, lo, hi, header=True):
        """SAM file with reads [lo, hi) (the lines of sam_lines(), written by the C generator)."""
        if header:                                   # (without a header the reads are appended to the file)
            with open(path, "w") as f:
                f.write("@HD\tVN:1.0\tSO:unsorted\n@SQ\tSN:%s\tLN:%d\n" % (self.chrom, self.chrom_len))
        self.lib.tcsynth_write_sam.restype = C.c_int64
        self.lib.tcsynth_write_sam.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int64, C.c_int64] + [C.c_void_p] * 12
        a = [np.ascontiguousarray(cols[k]) for k in ("flag", "chrom", "pos", "tags", "seq_off", "seq", "cig_off", "cig_op", "cig_len", "md_off", "md")]
        nm = np.ascontiguousarray(nm, dtype=np.int32)
        r = self.lib.tcsynth_write_sam(path.encode(), b"a", self.chrom.encode(), lo, hi, *[x.ctypes.data for x in a], nm.ctypes.data)
        if r < 0:
            raise IOError("cannot write " + path)
        return r

    def write_fasta(self, path, width=60):
        n = len(self.genome)
        pad = (-n) % width
        body = np.concatenate((self.genome, np.full(pad, ord("\n"), dtype=np.uint8))).reshape(-1, width)
        lines = np.concatenate((body, np.full((body.shape[0], 1), ord("\n"), dtype=np.uint8)), axis=1).tobytes()
        with open(path, "wb") as f:
            f.write((">%s\n" % self.chrom).encode())
            f.write(lines[:len(lines) - pad] if pad == 0 else lines[:-(pad + 1)] + b"\n")

    def close(self):
        self.lib.tcsynth_free_genes(C.byref(self.genes))


def variants_from_reads(w, cols, lo, hi, frac=0.3, seed=1):
    """VCF rows (CHROM POS ID REF ALT) that whitelist a fraction of the errors planted in reads
    [lo, hi): the variant-aware configuration of BASELINE.json (configs[2]) at test scale."""
    import re
    rng = np.random.default_rng(seed)
    g = w.genome
    rows = []
    for i in range(lo, hi):
        if cols["chrom"][i] < 0 or cols["flag"][i] > 16 or not (cols["tags"][i] & 2):
            continue
        a, b = cols["cig_off"][i], cols["cig_off"][i + 1]
        md = cols["md"][cols["md_off"][i]:cols["md_off"][i + 1]].tobytes().decode()
        seq = cols["seq"][cols["seq_off"][i]:cols["seq_off"][i + 1]].tobytes().decode()
        toks = re.findall(r"[0-9]+|\^[A-Za-z]+|[A-Za-z]", md)
        # walk CIGAR; mismatch positions from MD
        mm = []
        ref_i = 0
        for t in toks:
            if t[0].isdigit():
                ref_i += int(t)
            elif t[0] == "^":
                ref_i += len(t) - 1
            else:
                mm.append(ref_i); ref_i += 1
        gp, q, ref_i, k = int(cols["pos"][i]), 0, 0, 0
        for op, ct in zip(cols["cig_op"][a:b], cols["cig_len"][a:b]):
            op, ct = chr(op), int(ct)
            if op == "M":
                while k < len(mm) and mm[k] < ref_i + ct:
                    off = mm[k] - ref_i
                    if rng.random() < frac:
                        rows.append([w.chrom, gp + off, ".", chr(g[gp + off - 1]).upper(), seq[q + off]])
                    k += 1
                gp += ct; q += ct; ref_i += ct
            elif op == "D":
                if rng.random() < frac:
                    rows.append([w.chrom, gp - 1, ".", g[gp - 2:gp - 1 + ct].tobytes().decode().upper(), chr(g[gp - 2]).upper()])
                gp += ct; ref_i += ct
            elif op == "I":
                if rng.random() < frac:
                    anchor = chr(g[gp - 2]).upper()
                    rows.append([w.chrom, gp - 1, ".", anchor, anchor + seq[q:q + ct]])
                q += ct
            elif op == "S":
                q += ct
            elif op in "NH":
                gp += ct
    rows.sort(key=lambda r: r[1])
    return rows

"""ctypes binding of libtc_b200.so (C ABI: include/tc_b200.h).  No CPU fallback: if the CUDA
library is missing or no GPU is usable, constructing an Engine raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TC_B200_LIB") or os.path.join(_HERE, "csrc", "libtc_b200.so")   # (TC_B200_LIB: tuning builds)

EXPORTS = ["tc_device_count", "tc_ctx_create", "tc_ctx_destroy", "tc_last_error", "tc_ctx_set_options", "tc_ctx_stream",
           "tc_ctx_load_genome", "tc_ctx_load_junctions", "tc_ctx_load_variants", "tc_correct_batch",
           "tc_dryrun_batch", "tc_batch_upload", "tc_batch_run", "tc_batch_download", "tc_get_timing",
           "tc_ctx_set_pipeline", "tc_sam_parse", "tc_sam_free", "tc_sam_n_reads", "tc_sam_header", "tc_sam_batch", "tc_sam_render", "tc_sam_render_parts",
           "tc_sam_rnames", "tc_free", "tc_ctx_set_alphabet", "tc_seq_pack", "tc_seq_packed_size", "tc_seq_pack2", "tc_seq_packed2_size", "tc_cig_pack16", "tc_sam_pack_seq", "tc_sam_header_lines",
           "tc_vcf_parse", "tc_vcf_tables", "tc_vcf_n_records", "tc_vcf_free", "tc_sj_parse", "tc_sj_tables", "tc_sj_free", "tc_intron_motifs",
           "tc_vcf_chroms", "tc_fasta_open", "tc_fasta_n_contigs", "tc_fasta_contig", "tc_fasta_pack", "tc_fasta_exceptions", "tc_fasta_free"]

# status bits (include/tc_b200.h)
ST_CLASS_MASK, ST_FALLBACK, ST_CIGAR_REWRITTEN = 3, 1 << 2, 1 << 3
ST_NMMD_DEVICE, ST_NMMD_NONE, ST_JM_DEVICE, ST_JI_DEVICE = 1 << 4, 1 << 5, 1 << 6, 1 << 7
ST_CANONICAL, ST_ALL_ANNOT, ST_ERR_INTRON, ST_ERR_MERGE = 1 << 8, 1 << 9, 1 << 10, 1 << 11


class EngineError(RuntimeError):
    pass


class _Options(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("max_len_indel", "max_sj_offset", "correct_mismatches", "correct_indels",
                                          "correct_sjs", "sj_ignore_strand", "sj_tie_right", "reserved")]


class _BatchIn(C.Structure):
    _fields_ = [("n_reads", C.c_int64)] + [(n, C.c_void_p) for n in (
        "flag", "chrom", "pos", "tags", "seq_off", "seq", "cig_off", "cig_op", "cig_len", "md_off", "md", "ji_off", "ji",
        "seq_poff")] + [("seq_bits", C.c_int32), ("reserved0", C.c_int32), ("n_exc", C.c_int64), ("exc_pos", C.c_void_p), ("exc_byte", C.c_void_p),
                     ("cig16", C.c_void_p), ("n_cig_exc", C.c_int64), ("cig_exc_idx", C.c_void_p), ("cig_exc_op", C.c_void_p), ("cig_exc_len", C.c_void_p)]


class _BatchOut(C.Structure):
    _fields_ = [("n_reads", C.c_int64)] + [(n, C.c_void_p) for n in (
        "status", "counters", "nm", "delta_off", "delta_len", "seq_len", "delta", "cig_off", "cig_op", "cig_len", "md_off", "md_len", "md",
        "jn_off", "jm", "ji", "te_off", "te")]


class Timing(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("h2d_ms", "kernels_ms", "d2h_ms", "k_nmmd_init_ms", "k_measure_ms",
                                         "k_plan_ms", "k_junction_ms", "k_emit_ms")] + \
               [("launches", C.c_int32), ("md_pool_retries", C.c_int32), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("exact_nmmd_reads", C.c_int64), ("ascii_path_reads", C.c_int64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


TE_DTYPE = np.dtype([("a", "<i4"), ("b", "<i4"), ("size", "<i4"), ("type", "u1"), ("status", "u1"),
                     ("reason", "u1"), ("pad", "u1")])


def load_library(path=LIB_PATH):
    if not os.path.exists(path):
        raise EngineError("CUDA library %s is missing - run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a).  There is no CPU fallback." % path)
    lib = C.CDLL(path)
    _declare(lib)
    return lib


def _declare(lib):
    lib.tc_device_count.restype = C.c_int
    lib.tc_device_count.argtypes = []
    lib.tc_ctx_create.restype = C.c_void_p
    lib.tc_ctx_create.argtypes = [C.c_int]
    lib.tc_ctx_destroy.argtypes = [C.c_void_p]
    lib.tc_ctx_destroy.restype = None
    lib.tc_last_error.restype = C.c_char_p
    lib.tc_last_error.argtypes = [C.c_void_p]
    lib.tc_ctx_stream.restype = C.c_void_p
    lib.tc_ctx_stream.argtypes = [C.c_void_p]
    lib.tc_ctx_set_options.argtypes = [C.c_void_p, C.POINTER(_Options)]
    lib.tc_ctx_load_genome.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                       C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.tc_ctx_load_junctions.argtypes = [C.c_void_p] + [C.c_int64, C.c_void_p] * 4 + [C.c_int64, C.c_void_p, C.c_void_p]
    lib.tc_ctx_load_variants.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    for f in ("tc_correct_batch", "tc_dryrun_batch"):
        getattr(lib, f).argtypes = [C.c_void_p, C.POINTER(_BatchIn), C.POINTER(_BatchOut)]
    lib.tc_batch_upload.argtypes = [C.c_void_p, C.POINTER(_BatchIn)]
    lib.tc_batch_run.argtypes = [C.c_void_p, C.c_int]
    lib.tc_batch_download.argtypes = [C.c_void_p, C.POINTER(_BatchOut)]
    lib.tc_get_timing.argtypes = [C.c_void_p, C.POINTER(Timing)]
    lib.tc_ctx_set_pipeline.argtypes = [C.c_void_p, C.c_int64]
    lib.tc_sam_parse.restype = C.c_void_p
    lib.tc_sam_parse.argtypes = [C.c_char_p, C.c_int64, C.c_int32, C.POINTER(C.c_char_p), C.c_void_p, C.c_int32, C.c_char_p, C.c_int64]
    lib.tc_sam_free.argtypes = [C.c_void_p]
    lib.tc_sam_free.restype = None
    lib.tc_sam_rnames.argtypes = [C.c_char_p, C.c_int64, C.POINTER(C.c_int64)]
    lib.tc_sam_rnames.restype = C.c_void_p
    lib.tc_sam_header_lines.argtypes = [C.c_char_p, C.c_int64, C.POINTER(C.c_int64)]
    lib.tc_sam_header_lines.restype = C.c_void_p
    lib.tc_free.argtypes = [C.c_void_p]
    lib.tc_free.restype = None
    lib.tc_sam_n_reads.argtypes = [C.c_void_p]
    lib.tc_sam_n_reads.restype = C.c_int64
    lib.tc_sam_header.argtypes = [C.c_void_p, C.POINTER(C.c_char_p)]
    lib.tc_sam_header.restype = C.c_int64
    lib.tc_sam_batch.argtypes = [C.c_void_p, C.POINTER(_BatchIn)]
    lib.tc_sam_batch.restype = None
    lib.tc_ctx_set_alphabet.argtypes = [C.c_void_p, C.c_char_p, C.c_int32]
    lib.tc_seq_pack.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_char_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]
    lib.tc_seq_pack.restype = C.c_int64
    lib.tc_seq_packed_size.argtypes = [C.c_void_p, C.c_int64]
    lib.tc_seq_packed_size.restype = C.c_int64
    lib.tc_seq_packed2_size.argtypes = [C.c_void_p, C.c_int64]
    lib.tc_seq_packed2_size.restype = C.c_int64
    lib.tc_seq_pack2.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int32]
    lib.tc_seq_pack2.restype = C.c_int64
    lib.tc_cig_pack16.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
    lib.tc_cig_pack16.restype = C.c_int64
    lib.tc_sam_pack_seq.argtypes = [C.c_void_p, C.c_char_p, C.c_int32, C.c_int32]
    lib.tc_sam_render.argtypes = [C.c_void_p, C.POINTER(_BatchOut)] + [C.c_int32] * 4 + [C.POINTER(C.c_void_p), C.POINTER(C.c_int64)] * 4 + [C.c_char_p, C.c_int64]
    lib.tc_sam_render_parts.argtypes = [C.c_void_p, C.POINTER(_BatchOut)] + [C.c_int32] * 4 + [C.POINTER(C.c_int32), C.c_void_p, C.c_void_p, C.c_int32, C.c_char_p, C.c_int64]
    lib.tc_intron_motifs.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.tc_vcf_parse.restype = C.c_void_p
    lib.tc_vcf_parse.argtypes = [C.c_char_p, C.c_int64, C.c_int32, C.POINTER(C.c_char_p), C.c_void_p, C.c_int32, C.c_int32, C.c_char_p, C.c_int64]
    lib.tc_vcf_tables.restype = None
    lib.tc_vcf_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 11
    lib.tc_vcf_n_records.restype = C.c_int64
    lib.tc_vcf_n_records.argtypes = [C.c_void_p]
    lib.tc_vcf_free.restype = None
    lib.tc_vcf_free.argtypes = [C.c_void_p]
    lib.tc_sj_parse.restype = C.c_void_p
    lib.tc_sj_parse.argtypes = [C.c_char_p, C.c_int64, C.c_int32, C.POINTER(C.c_char_p), C.c_void_p, C.c_char_p, C.c_int64]
    lib.tc_sj_tables.restype = None
    lib.tc_sj_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 11
    lib.tc_sj_free.restype = None
    lib.tc_sj_free.argtypes = [C.c_void_p]
    lib.tc_vcf_chroms.restype = C.c_void_p
    lib.tc_vcf_chroms.argtypes = [C.c_char_p, C.c_int64, C.POINTER(C.c_int64)]
    lib.tc_fasta_open.restype = C.c_void_p
    lib.tc_fasta_open.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int64]
    lib.tc_fasta_n_contigs.restype = C.c_int32
    lib.tc_fasta_n_contigs.argtypes = [C.c_void_p]
    lib.tc_fasta_contig.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
    lib.tc_fasta_pack.restype = C.c_int64
    lib.tc_fasta_pack.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32]
    lib.tc_fasta_exceptions.restype = C.c_int64
    lib.tc_fasta_exceptions.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.tc_fasta_free.argtypes = [C.c_void_p]
    lib.tc_fasta_free.restype = None
    for f in EXPORTS:
        if f not in ("tc_vcf_chroms", "tc_fasta_open", "tc_fasta_n_contigs", "tc_fasta_pack", "tc_fasta_exceptions", "tc_fasta_free", "tc_vcf_parse", "tc_vcf_tables", "tc_vcf_n_records", "tc_vcf_free", "tc_sj_parse", "tc_sj_tables", "tc_sj_free", "tc_device_count", "tc_ctx_create", "tc_ctx_destroy", "tc_last_error", "tc_ctx_stream", "tc_sam_parse", "tc_sam_free",
                     "tc_sam_n_reads", "tc_sam_header", "tc_sam_batch", "tc_sam_rnames", "tc_free", "tc_seq_pack", "tc_seq_packed_size", "tc_seq_pack2", "tc_seq_packed2_size", "tc_cig_pack16", "tc_sam_header_lines"):
            getattr(lib, f).restype = C.c_int


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


def decode_te_words(words, te_off):
    """tc_batch_out.te (8-byte words, tc_te_decode in include/tc_b200.h) -> (rows as TE_DTYPE, row offsets per read)."""
    n = len(te_off) - 1
    if len(words) == 0:
        return np.zeros(0, dtype=TE_DTYPE), np.zeros(n + 1, dtype=np.int64)
    nw = np.diff(te_off)
    read = np.repeat(np.arange(n), nw)
    idx = np.arange(len(words)) - np.repeat(te_off[:-1], nw)
    typ = ((words >> np.uint64(56)) & np.uint64(3)).astype(np.int64)
    # junction rows (type 3) close a read's list and take two words each: the first one marks where the pairs begin
    first3 = np.full(n, np.iinfo(np.int64).max)
    np.minimum.at(first3, read[typ == 3], idx[typ == 3])
    f3 = first3[read]
    is_row = (idx < f3) | (((idx - f3) & 1) == 0)
    w = words[is_row]
    rows = np.zeros(len(w), dtype=TE_DTYPE)
    rows["a"] = (w & np.uint64(0xffffffff)).astype(np.uint32).view(np.int32)
    rows["size"] = ((w >> np.uint64(32)) & np.uint64(0xffffff)).astype(np.int32)
    rows["type"] = ((w >> np.uint64(56)) & np.uint64(3)).astype(np.uint8)
    rows["status"] = ((w >> np.uint64(58)) & np.uint64(1)).astype(np.uint8)
    rows["reason"] = ((w >> np.uint64(59)) & np.uint64(7)).astype(np.uint8)
    nxt = np.concatenate((words[1:], np.zeros(1, dtype=np.uint64)))[is_row]
    b3 = (nxt & np.uint64(0xffffffff)).astype(np.uint32).view(np.int32)
    rows["b"] = np.where(rows["type"] == 3, b3, np.where(rows["type"] == 0, 0, rows["a"] + rows["size"]))
    per_read = np.bincount(read[is_row], minlength=n)
    return rows, np.concatenate(([0], np.cumsum(per_read))).astype(np.int64)


def apply_delta(in_seq, delta):
    """tc_seq_apply_delta (include/tc_b200.h): input SEQ bytes + delta bytes -> new SEQ bytes."""
    out, q, k, n_d = bytearray(), 0, 0, len(delta)
    while k < n_d:
        kind, n = delta[k] >> 6, delta[k] & 63
        k += 1
        if n == 62:
            n = delta[k] | (delta[k + 1] << 8); k += 2
        elif n == 63:
            n = delta[k] | (delta[k + 1] << 8) | (delta[k + 2] << 16) | (delta[k + 3] << 24); k += 4
        if kind == 1:
            q += n
        elif kind == 0:
            if q + n > len(in_seq):
                raise EngineError("SEQ delta runs past the input SEQ")
            out += in_seq[q:q + n]; q += n
        elif kind == 2:
            out += delta[k:k + n]; k += n
        else:
            raise EngineError("malformed SEQ delta")
    return bytes(out)


class BatchResult:
    """numpy copies of a tc_batch_out.  The new SEQ of a read comes back as a delta against its input SEQ
    (`read_seq(i)` applies it to the `cols` the batch was made from); TE rows are decoded to TE_DTYPE."""

    def __init__(self, out, dry, alphabet=None, cols=None):
        n = out.n_reads
        self.n = n
        self.dry = dry
        self._cols, self._alphabet = cols, alphabet

        def arr(p, dtype, count):
            if not p or count == 0:
                return np.zeros(0, dtype=dtype)
            buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(p)
            return np.frombuffer(buf, dtype=dtype, count=count).copy()

        self.status = arr(out.status, np.uint32, n)
        self.counters = arr(out.counters, np.int32, 10 * n).reshape(n, 10)
        self.te_word_off = arr(out.te_off, np.int64, n + 1) if n else np.zeros(1, np.int64)
        self.te_words = arr(out.te, np.uint64, int(self.te_word_off[-1]))
        self.te, self.te_off = decode_te_words(self.te_words, self.te_word_off)
        if dry:
            return
        self.nm = arr(out.nm, np.int32, n)
        self.seq_len = arr(out.seq_len, np.int32, n)
        self.delta_off = arr(out.delta_off, np.int64, n)
        self.delta_len = arr(out.delta_len, np.int32, n)
        d_total = int((self.delta_off + np.maximum(self.delta_len, 0)).max()) if n else 0
        self.delta = arr(out.delta, np.uint8, d_total)
        self.cig_off = arr(out.cig_off, np.int64, n + 1) if n else np.zeros(1, np.int64)
        self.cig_op = arr(out.cig_op, np.uint8, int(self.cig_off[-1]))
        self.cig_len = arr(out.cig_len, np.int32, int(self.cig_off[-1]))
        self.md_off = arr(out.md_off, np.int64, n)
        self.md_len = arr(out.md_len, np.int32, n)
        md_total = int((self.md_off + self.md_len).max()) if n else 0
        self.md = arr(out.md, np.uint8, md_total)
        self.jn_off = arr(out.jn_off, np.int64, n + 1) if n else np.zeros(1, np.int64)
        self.jm = arr(out.jm, np.int8, int(self.jn_off[-1]))
        self.ji = arr(out.ji, np.int32, 2 * int(self.jn_off[-1]))

    def input_seq(self, i):
        """The input SEQ of read i as bytes (ASCII), from the columns the batch was made from."""
        c = self._cols
        if c is None:
            raise EngineError("this BatchResult was made without the input columns: read_seq() needs them")
        a, b = int(c["seq_off"][i]), int(c["seq_off"][i + 1])
        if c.get("seq_poff") is None:
            return c["seq"][a:b].tobytes()
        p = int(c["seq_poff"][i])
        if int(c.get("seq_bits", 4)) == 2:
            pk = c["seq"][p:p + (b - a + 3) // 4]
            out = np.frombuffer(b"ACGT", dtype=np.uint8)[(pk[:, None] >> (2 * np.arange(4, dtype=np.uint8))) & 3].reshape(-1)[:b - a].copy()
            lo, hi = np.searchsorted(c["exc_pos"], a), np.searchsorted(c["exc_pos"], b)
            out[c["exc_pos"][lo:hi] - a] = c["exc_byte"][lo:hi]
            return out.tobytes()
        pk = c["seq"][p:p + (b - a + 1) // 2]
        al = np.frombuffer(bytes(self._alphabet).ljust(16, b"\0"), dtype=np.uint8)
        out = np.empty(2 * len(pk), dtype=np.uint8)
        out[0::2], out[1::2] = al[pk & 15], al[pk >> 4]
        return out[:b - a].tobytes()

    def read_seq(self, i):
        """The corrected SEQ of read i as bytes."""
        s = self.input_seq(i)
        if self.delta_len[i] < -1:
            raise EngineError("read %d: the engine flagged its SEQ delta as inconsistent" % i)
        if self.delta_len[i] < 0:
            return s
        a = int(self.delta_off[i])
        out = apply_delta(s, self.delta[a:a + int(self.delta_len[i])].tobytes())
        if len(out) != int(self.seq_len[i]):
            raise EngineError("read %d: SEQ delta gives %d bases, seq_len says %d" % (i, len(out), int(self.seq_len[i])))
        return out


def _c_text(text):
    """(argument for a `const char*` parameter, length) of a SAM text: bytes as they are, any other writable buffer (a slice
    of the CLI's file mapping) without a copy."""
    if isinstance(text, bytes):
        return text, len(text)
    mv = memoryview(text)
    return ((C.c_char * mv.nbytes).from_buffer(mv) if mv.nbytes and not mv.readonly else bytes(mv)), mv.nbytes


def sam_rnames(lib, text):
    """Set of RNAME strings of the alignment lines of a SAM text (bytes)."""
    n = C.c_int64()
    text, n_text = _c_text(text)
    p = lib.tc_sam_rnames(text, n_text, C.byref(n))
    if not p:
        raise MemoryError("tc_sam_rnames")
    try:
        return set(C.string_at(p, n.value).decode().split("\n")) - {""}
    finally:
        lib.tc_free(p)


def sam_header_lines(lib, text):
    """The '@' lines of a SAM text (bytes), newline-terminated, as str."""
    n = C.c_int64()
    text, n_text = _c_text(text)
    p = lib.tc_sam_header_lines(text, n_text, C.byref(n))
    if not p:
        raise MemoryError("tc_sam_header_lines")
    try:
        return C.string_at(p, n.value).decode()
    finally:
        lib.tc_free(p)


class SamText:
    """SAM text parsed by the library (tc_sam_parse): header, columnar batch, and the renderer of the
    four output texts.  Host-side only; the same logic as sam.py / render.py, in multi-threaded C++."""

    def __init__(self, lib, text, genome, n_threads=None, alphabet=None):
        self._lib = lib
        self._text, n_text = _c_text(text)                                 # kept alive: fields are referenced
        names = (C.c_char_p * len(genome.names))(*[n.encode() for n in genome.names])
        err = C.create_string_buffer(1024)
        nt = int(n_threads or min(32, os.cpu_count() or 1))
        self._nt = nt
        self._h = lib.tc_sam_parse(self._text, n_text, len(genome.names), names, genome.chrom_len.ctypes.data, nt, err, 1024)
        if not self._h:
            from .sam import SamFormatError
            raise SamFormatError(err.value.decode())
        self.n = lib.tc_sam_n_reads(self._h)
        p = C.c_char_p()
        n = lib.tc_sam_header(self._h, C.byref(p))
        self.header = C.string_at(p, n).decode() if n else ""
        self._ascii = _BatchIn()
        lib.tc_sam_batch(self._h, C.byref(self._ascii))
        self.batch = self._ascii
        if alphabet and lib.tc_sam_pack_seq(self._h, bytes(alphabet), len(alphabet), nt) == 1:    # SEQ column in 4-bit form
            self.batch = _BatchIn()
            lib.tc_sam_batch(self._h, C.byref(self.batch))

    def cols(self):
        """numpy views of the parsed columns (valid while this object lives)."""
        b, n = self._ascii, self.n

        def view(ptr, dtype, count):
            if not ptr or count == 0:
                return np.zeros(0, dtype=dtype)
            return np.frombuffer((C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr), dtype=dtype, count=count)

        so, co = view(b.seq_off, np.int64, n + 1), view(b.cig_off, np.int64, n + 1)
        mo, jo = view(b.md_off, np.int64, n + 1), view(b.ji_off, np.int64, n + 1)
        if n == 0:
            so = co = mo = jo = np.zeros(1, dtype=np.int64)
        return {"flag": view(b.flag, np.int32, n), "chrom": view(b.chrom, np.int32, n), "pos": view(b.pos, np.int32, n),
                "tags": view(b.tags, np.uint8, n), "seq_off": so, "seq": view(b.seq, np.uint8, int(so[-1])), "cig_off": co,
                "cig_op": view(b.cig_op, np.uint8, int(co[-1])), "cig_len": view(b.cig_len, np.int32, int(co[-1])), "md_off": mo,
                "md": view(b.md, np.uint8, int(mo[-1])), "ji_off": jo, "ji": view(b.ji, np.int32, int(jo[-1]))}

    def render(self, raw_out, primary_only=False, canon_only=False, dry=False, copy=True):
        """raw_out: the _BatchOut of Engine.raw_correct / raw_dryrun for this batch -> four bytes objects
        (copy=False: memoryviews of the library's buffers, valid until this object is closed)."""
        ptr = [C.c_void_p() for _ in range(4)]
        ln = [C.c_int64() for _ in range(4)]
        err = C.create_string_buffer(1024)
        args = []
        for a, b in zip(ptr, ln):
            args += [C.byref(a), C.byref(b)]
        rc = self._lib.tc_sam_render(self._h, C.byref(raw_out), int(primary_only), int(canon_only), int(dry), self._nt, *args, err, 1024)
        if rc != 0:
            raise EngineError(err.value.decode())
        if copy:
            return {k: C.string_at(p, l.value) if l.value else b"" for k, p, l in zip(("sam", "fa", "log", "te"), ptr, ln)}
        return {k: memoryview((C.c_char * l.value).from_address(p.value)).cast("B") if l.value else memoryview(b"")
                for k, p, l in zip(("sam", "fa", "log", "te"), ptr, ln)}

    def render_parts(self, raw_out, primary_only=False, canon_only=False, dry=False):
        """Like render(copy=False) but without the library's final join: {"sam","fa","log","te"} -> list of memoryviews
        (the formatter threads' pieces, in order), valid until this object renders again or is closed."""
        cap = 64
        ptr, ln, n = (C.c_void_p * (4 * cap))(), (C.c_int64 * (4 * cap))(), C.c_int32()
        err = C.create_string_buffer(1024)
        rc = self._lib.tc_sam_render_parts(self._h, C.byref(raw_out), int(primary_only), int(canon_only), int(dry), self._nt,
                                           C.byref(n), ptr, ln, cap, err, 1024)
        if rc != 0:
            raise EngineError(err.value.decode())
        out = {}
        for q, name in enumerate(("sam", "fa", "log", "te")):
            out[name] = [memoryview((C.c_char * ln[4 * k + q]).from_address(ptr[4 * k + q])).cast("B") for k in range(n.value) if ln[4 * k + q]]
        return out

    def close(self):
        if getattr(self, "_h", None):
            self._lib.tc_sam_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """One correction context on one GPU."""

    def __init__(self, device=0):
        self._setup(load_library(), device)

    def _setup(self, lib, device):
        self._lib = lib
        self._ctx = lib.tc_ctx_create(int(device))
        if not self._ctx:
            raise EngineError("cannot create a CUDA context on device %d: %s (no CPU fallback)"
                              % (device, lib.tc_last_error(None).decode()))
        self.device = device
        self._keep = []
        self.alphabet = None       # bytes of the 4-bit SEQ codes once a genome is loaded (None: SEQ travels as ASCII)
        self.cig16 = True          # CIGAR as 16-bit words + a list of the ops that do not fit (column path)
        self.seq2 = True           # ... and in 2-bit form + a list of the bases outside ACGT when those are few (sequencer reads)
        self.seq4 = True           # send SEQ in 4-bit form whenever the batch allows it (column path: PCIe is what limits it)
        self.seq4_text = False     # SAM-text path: ASCII (the host's parse / render threads limit it; packing would only add to them)

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.tc_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise EngineError("%s failed (%d): %s" % (what, rc, self._lib.tc_last_error(self._ctx).decode()))

    def stream(self):
        return self._lib.tc_ctx_stream(self._ctx)

    def set_options(self, maxLenIndel=5, maxSJOffset=5, correctMismatches=True, correctIndels=True,
                    correctSJs=True, sj_ignore_strand=False, sj_tie_right=False, debug=0):
        o = _Options(int(maxLenIndel), int(maxSJOffset), int(bool(correctMismatches)), int(bool(correctIndels)),
                     int(bool(correctSJs)), int(bool(sj_ignore_strand)), int(bool(sj_tie_right)), int(debug))
        self._check(self._lib.tc_ctx_set_options(self._ctx, C.byref(o)), "tc_ctx_set_options")

    def set_pipeline(self, chunk_reads):
        self._check(self._lib.tc_ctx_set_pipeline(self._ctx, int(chunk_reads)), "tc_ctx_set_pipeline")

    def load_genome(self, g):
        st, en, by = g.exceptions()
        self._check(self._lib.tc_ctx_load_genome(
            self._ctx, len(g.names), _ptr(g.chrom_len), _ptr(g.chrom_off), int(g.n_pad), _ptr(g.bases2), _ptr(g.lower1),
            len(st), _ptr(st), _ptr(en), _ptr(by)), "tc_ctx_load_genome")
        # alphabet of the 4-bit SEQ form: the usual bases plus whatever else the genome holds
        al = bytearray(b"ACGTacgt")
        for c in sorted(set(np.asarray(by, dtype=np.uint8).tolist())) + [ord("N"), ord("n")]:
            if c not in al and (len(al) < 16 or c not in b"Nn"):
                al.append(c)
        self.alphabet = bytes(al) if len(al) <= 16 else None       # None: too many distinct genome bytes, SEQ stays ASCII
        if self.alphabet:
            self._check(self._lib.tc_ctx_set_alphabet(self._ctx, self.alphabet, len(self.alphabet)), "tc_ctx_set_alphabet")

    def load_junctions(self, t):
        self._check(self._lib.tc_ctx_load_junctions(
            self._ctx, len(t.don_s), _ptr(t.don_s), len(t.acc_s), _ptr(t.acc_s), len(t.don_u), _ptr(t.don_u),
            len(t.acc_u), _ptr(t.acc_u), len(t.ann_k1), _ptr(t.ann_k1), _ptr(t.ann_k2)), "tc_ctx_load_junctions")

    def load_variants(self, v):
        self._check(self._lib.tc_ctx_load_variants(
            self._ctx, len(v.snp_key), _ptr(v.snp_key), _ptr(v.snp_alt), len(v.ins_key), _ptr(v.ins_key), _ptr(v.ins_size),
            _ptr(v.ins_aoff), _ptr(v.ins_bytes), len(v.del_key), _ptr(v.del_key), _ptr(v.del_size)), "tc_ctx_load_variants")

    def pack_cols2(self, cols, n_threads=None, max_exc_frac=0.02):
        """cols with the SEQ column in 2-bit form (`seq` packed, `seq_poff`, `seq_bits` = 2, `exc_pos` / `exc_byte` = the bases
        outside ACGT), or None when more than max_exc_frac of the bases would have to be listed."""
        n = len(cols["flag"])
        so = np.ascontiguousarray(cols["seq_off"], dtype=np.int64)
        out = np.zeros(int(self._lib.tc_seq_packed2_size(so.ctypes.data, n)) + 64, dtype=np.uint8)
        poff = np.zeros(n + 1, dtype=np.int64)
        cap = int(max_exc_frac * int(so[-1] if n else 0)) + 1024
        epos, ebyte, ne = np.zeros(cap, dtype=np.int64), np.zeros(cap, dtype=np.uint8), C.c_int64()
        seq = cols["seq"]
        r = self._lib.tc_seq_pack2(seq.ctypes.data if seq.size else None, so.ctypes.data, n, out.ctypes.data, poff.ctypes.data, cap,
                                   epos.ctypes.data, ebyte.ctypes.data, C.byref(ne), int(n_threads or min(32, os.cpu_count() or 1)))
        if r < 0:
            return None
        packed = dict(cols)
        packed["seq"], packed["seq_poff"], packed["seq_off"] = out, poff, so
        packed["seq_bits"], packed["exc_pos"], packed["exc_byte"] = 2, epos[:ne.value].copy(), ebyte[:ne.value].copy()
        return packed

    def pack_cigar(self, cols, n_threads=None, max_exc_frac=0.25):
        """cols with the CIGAR as 16-bit words (`cig16`) + the listed ops that do not fit (`cig_exc_*`); cols unchanged when
        more than max_exc_frac of the ops would have to be listed."""
        n_ops = len(cols["cig_op"])
        if n_ops == 0 or "cig16" in cols:
            return cols
        cap = int(max_exc_frac * n_ops) + 64
        w = np.zeros(n_ops + 8, dtype=np.uint16)
        eidx, eop, elen = np.zeros(cap, dtype=np.int64), np.zeros(cap, dtype=np.uint8), np.zeros(cap, dtype=np.int32)
        op, ln = np.ascontiguousarray(cols["cig_op"]), np.ascontiguousarray(cols["cig_len"], dtype=np.int32)
        m = self._lib.tc_cig_pack16(op.ctypes.data, ln.ctypes.data, n_ops, w.ctypes.data, cap, eidx.ctypes.data, eop.ctypes.data, elen.ctypes.data,
                                    int(n_threads or min(32, os.cpu_count() or 1)))
        if m < 0:
            return cols
        out = dict(cols)
        out["cig16"], out["cig_exc_idx"], out["cig_exc_op"], out["cig_exc_len"] = w, eidx[:m].copy(), eop[:m].copy(), elen[:m].copy()
        return out

    def pack_cols(self, cols, n_threads=None):
        """cols with the SEQ column in 4-bit form (`seq` packed, `seq_poff` added), or None when a SEQ
        byte is outside the alphabet."""
        if not self.alphabet:
            return None
        n = len(cols["flag"])
        so = np.ascontiguousarray(cols["seq_off"], dtype=np.int64)
        out = np.zeros(int(self._lib.tc_seq_packed_size(so.ctypes.data, n)) + 64, dtype=np.uint8)
        poff = np.zeros(n + 1, dtype=np.int64)
        seq = cols["seq"]
        r = self._lib.tc_seq_pack(seq.ctypes.data if seq.size else None, so.ctypes.data, n, self.alphabet, len(self.alphabet),
                                  out.ctypes.data, poff.ctypes.data, int(n_threads or min(32, os.cpu_count() or 1)))
        if r < 0:
            return None
        packed = dict(cols)
        packed["seq"], packed["seq_poff"], packed["seq_off"] = out, poff, so
        return packed

    def _batch_in(self, cols):
        if self.seq4 and "seq_poff" not in cols:
            cols = (self.pack_cols2(cols) if self.seq2 else None) or self.pack_cols(cols) or cols
        if self.cig16:
            cols = self.pack_cigar(cols)
        b = _BatchIn()
        b.n_reads = len(cols["flag"])
        for k in ("flag", "chrom", "pos", "tags", "seq_off", "seq", "cig_off", "cig_op", "cig_len", "md_off", "md",
                  "ji_off", "ji", "seq_poff", "exc_pos", "exc_byte", "cig16", "cig_exc_idx", "cig_exc_op", "cig_exc_len"):
            a = cols.get(k)
            if a is None:
                continue
            assert a.flags["C_CONTIGUOUS"]
            setattr(b, k, a.ctypes.data if a.size else None)
        b.seq_bits = int(cols.get("seq_bits", 4 if cols.get("seq_poff") is not None else 8))
        b.n_exc = len(cols["exc_pos"]) if cols.get("exc_pos") is not None else 0
        b.n_cig_exc = len(cols["cig_exc_idx"]) if cols.get("cig_exc_idx") is not None else 0
        self._keep = [cols]          # the arrays behind the pointers
        return b

    def correct(self, cols):
        """tc_correct_batch: host columns in -> BatchResult (host copies)."""
        b, out = self._batch_in(cols), _BatchOut()
        self._check(self._lib.tc_correct_batch(self._ctx, C.byref(b), C.byref(out)), "tc_correct_batch")
        return BatchResult(out, False, self.alphabet, self._keep[0])

    def dryrun(self, cols):
        b, out = self._batch_in(cols), _BatchOut()
        self._check(self._lib.tc_dryrun_batch(self._ctx, C.byref(b), C.byref(out)), "tc_dryrun_batch")
        return BatchResult(out, True, self.alphabet, self._keep[0])

    # split calls for device-resident measurement
    def upload(self, cols):
        b = self._batch_in(cols)
        self._check(self._lib.tc_batch_upload(self._ctx, C.byref(b)), "tc_batch_upload")

    def run(self, dry=False):
        self._check(self._lib.tc_batch_run(self._ctx, int(dry)), "tc_batch_run")

    def download(self, dry=False, copy=True):
        out = _BatchOut()
        self._check(self._lib.tc_batch_download(self._ctx, C.byref(out)), "tc_batch_download")
        return BatchResult(out, dry, self.alphabet, self._keep[0] if self._keep else None) if copy else out

    def raw_correct(self, b):
        """tc_correct_batch on a prepared _BatchIn; returns the raw _BatchOut (no numpy copies)."""
        out = _BatchOut()
        self._check(self._lib.tc_correct_batch(self._ctx, C.byref(b), C.byref(out)), "tc_correct_batch")
        return out

    def raw_dryrun(self, b):
        out = _BatchOut()
        self._check(self._lib.tc_dryrun_batch(self._ctx, C.byref(b), C.byref(out)), "tc_dryrun_batch")
        return out

    def parse_sam(self, text, genome, n_threads=None):
        return SamText(self._lib, text, genome, n_threads, self.alphabet if (self.seq4 and self.seq4_text) else None)

    def intron_motifs(self, chrom, start, end):
        """Motif codes (0-6, -1 where a coordinate is below 1) of introns [start, end] (1-based inclusive) on contig ids `chrom`,
        with the junction kernel's motif fetch on the device."""
        c, s, e = (np.ascontiguousarray(x, dtype=np.int32) for x in (chrom, start, end))
        out = np.zeros(len(c), dtype=np.int8)
        self._check(self._lib.tc_intron_motifs(self._ctx, len(c), _ptr(c), _ptr(s), _ptr(e), _ptr(out)), "tc_intron_motifs")
        return out

    def timing(self):
        t = Timing()
        self._check(self._lib.tc_get_timing(self._ctx, C.byref(t)), "tc_get_timing")
        return t.as_dict()

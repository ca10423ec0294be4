"""Genome image for the device: 2-bit base plane + lower-case plane + literal runs.

Replaces `pyfaidx.Fasta` as used by the reference (TranscriptClean.py:221; fetches at
transcript.py:300,323, intronBound.py:36-44, TC.py:1016,1036,1142,1441,1457).  The image keeps
every byte of the FASTA recoverable - case and N / IUPAC codes leak into SEQ and MD in the
reference - in 0.375 byte per base:
  bases2  uint32, 16 bases per word, base i in bits 2*(i%16); A=0 C=1 G=2 T=3
  lower1  uint32, 32 bases per word, bit set = lower case
  exc_*   sorted runs [start,end) of one literal byte for everything outside ACGTacgt
Contigs are laid out back to back, each starting at a multiple of 64 bases.
"""
import numpy as np

_CODE = np.zeros(256, dtype=np.uint8)
for _i, _c in enumerate("ACGT"):
    _CODE[ord(_c)] = _i
    _CODE[ord(_c.lower())] = _i
_IS_ACGT = np.zeros(256, dtype=bool)
for _c in "ACGTacgt":
    _IS_ACGT[ord(_c)] = True
_IS_LOWER_ACGT = np.zeros(256, dtype=bool)
for _c in "acgt":
    _IS_LOWER_ACGT[ord(_c)] = True


def _pad64(n):
    return (n + 63) // 64 * 64


class GenomeImage:
    def __init__(self, names, lens):
        self.names = list(names)
        self.index = {n: i for i, n in enumerate(self.names)}
        self.chrom_len = np.asarray(lens, dtype=np.int64)
        offs, o = [], 0
        for L in lens:
            offs.append(o)
            o += _pad64(int(L)) + 64
        self.chrom_off = np.asarray(offs, dtype=np.int64)
        self.n_pad = max(64, o)
        self.bases2 = np.zeros(self.n_pad // 16, dtype=np.uint32)
        self.lower1 = np.zeros(self.n_pad // 32, dtype=np.uint32)
        self._exc = []          # (start, end, byte) in plane coordinates

    # ------------------------------------------------------------------ builders
    def _put(self, off, a):
        """Write ASCII bytes `a` (uint8 array) at plane offset `off` (any alignment)."""
        n = len(a)
        if n == 0:
            return
        if off % 32 == 0:
            self._put_aligned(off, a)
        else:
            head = min(n, 32 - off % 32)
            self._put_slow(off, a[:head])
            if n > head:
                self._put_aligned(off + head, a[head:])

    def _put_slow(self, off, a):
        pos = off + np.arange(len(a), dtype=np.int64)
        codes = _CODE[a].astype(np.uint32)
        w, s = pos >> 4, ((pos & 15) * 2).astype(np.uint32)
        np.bitwise_and.at(self.bases2, w, ~(np.uint32(3) << s))
        np.bitwise_or.at(self.bases2, w, codes << s)
        lw, ls = pos >> 5, (pos & 31).astype(np.uint32)
        np.bitwise_and.at(self.lower1, lw, ~(np.uint32(1) << ls))
        np.bitwise_or.at(self.lower1, lw, _IS_LOWER_ACGT[a].astype(np.uint32) << ls)
        self._scan_exc(off, a)

    def _put_aligned(self, off, a):
        n = len(a)
        full = n // 32 * 32
        if full:
            body = a[:full]
            c = _CODE[body].reshape(-1, 4)
            b = (c[:, 0] | (c[:, 1] << 2) | (c[:, 2] << 4) | (c[:, 3] << 6)).astype(np.uint8)
            self.bases2[off // 16: off // 16 + full // 16] = b.view("<u4")
            lo = np.packbits(_IS_LOWER_ACGT[body], bitorder="little")
            self.lower1[off // 32: off // 32 + full // 32] = lo.view("<u4")
            self._scan_exc(off, body)
        if n > full:
            self._put_slow(off + full, a[full:])

    def _scan_exc(self, off, a):
        bad = ~_IS_ACGT[a]
        if not bad.any():
            return
        idx = np.flatnonzero(bad)
        vals = a[idx]
        brk = np.flatnonzero((np.diff(idx) != 1) | (np.diff(vals) != 0)) + 1
        starts = np.concatenate(([0], brk))
        ends = np.concatenate((brk, [len(idx)]))
        for s, e in zip(starts, ends):
            self._exc.append((off + int(idx[s]), off + int(idx[e - 1]) + 1, int(vals[s])))

    @classmethod
    def from_dict(cls, seqs):
        names = list(seqs.keys())
        g = cls(names, [len(seqs[n]) for n in names])
        for i, n in enumerate(names):
            s = seqs[n]
            a = np.frombuffer(s.encode("ascii") if isinstance(s, str) else bytes(s), dtype=np.uint8)
            g._put(int(g.chrom_off[i]), a)
        return g

    @classmethod
    def from_fasta(cls, path):
        """Contig name = first word of the header (pyfaidx key semantics, TC.py:469-470)."""
        with open(path, "rb") as f:
            data = f.read()
        recs = []
        pos = 0
        while True:
            h = data.find(b">", pos)
            if h < 0:
                break
            nl = data.find(b"\n", h)
            nxt = data.find(b"\n>", nl)
            end = len(data) if nxt < 0 else nxt + 1
            name = data[h + 1:nl].split()[0].decode()
            body = np.frombuffer(data, dtype=np.uint8, count=end - nl - 1, offset=nl + 1)
            body = body[(body != 10) & (body != 13)]
            recs.append((name, body))
            pos = end
        g = cls([r[0] for r in recs], [len(r[1]) for r in recs])
        for i, (_, body) in enumerate(recs):
            g._put(int(g.chrom_off[i]), body)
        return g

    @classmethod
    def from_fasta_lib(cls, path, lib, n_threads=None):
        """from_fasta through the library's C++ packer (tc_fasta_open / tc_fasta_pack, csrc/tc_tables.cpp): the file is mapped
        and packed by all host threads - seconds for a 3-Gb genome where the numpy passes above take minutes.  Same image."""
        import ctypes as C
        import os
        nt = int(n_threads or min(32, os.cpu_count() or 1))
        err = C.create_string_buffer(512)
        h = lib.tc_fasta_open(os.fsencode(str(path)), nt, err, 512)
        if not h:
            raise ValueError("%s: %s" % (path, err.value.decode() or "cannot read the FASTA file"))
        try:
            names, lens = [], []
            p, n, L = C.c_void_p(), C.c_int32(), C.c_int64()
            for k in range(lib.tc_fasta_n_contigs(h)):
                lib.tc_fasta_contig(h, k, C.byref(p), C.byref(n), C.byref(L))
                names.append(C.string_at(p, n.value).decode())
                lens.append(int(L.value))
            g = cls(names, lens)
            n_exc = lib.tc_fasta_pack(h, g.chrom_off.ctypes.data, g.n_pad, g.bases2.ctypes.data, g.lower1.ctypes.data, nt)
            if n_exc < 0:
                raise ValueError("%s: the contig table does not fit the planes" % path)
            st, en, by = np.zeros(n_exc, dtype=np.int64), np.zeros(n_exc, dtype=np.int64), np.zeros(n_exc, dtype=np.uint8)
            if n_exc:
                lib.tc_fasta_exceptions(h, st.ctypes.data, en.ctypes.data, by.ctypes.data)
            g._exc = list(zip(st.tolist(), en.tolist(), by.tolist()))
            return g
        finally:
            lib.tc_fasta_free(h)

    # ------------------------------------------------------------------ on-disk image
    def save(self, path, source=None):
        """The packed image as one .npz (what the device gets; 0.375 bytes per base).  `source` = fingerprint of the FASTA
        it was made from (fasta_fingerprint), kept so that a cache is never trusted for another file."""
        st, en, by = self.exceptions()
        with open(path, "wb") as f:
            np.savez(f, names=np.asarray(self.names), chrom_len=self.chrom_len, bases2=self.bases2, lower1=self.lower1,
                     exc_start=st, exc_end=en, exc_byte=by, source=np.asarray(source if source is not None else [], dtype=np.int64))

    @classmethod
    def load(cls, path):
        z = np.load(path, allow_pickle=False)
        g = cls([str(n) for n in z["names"]], [int(x) for x in z["chrom_len"]])
        if z["bases2"].shape != g.bases2.shape or z["lower1"].shape != g.lower1.shape:
            raise ValueError("genome image %s does not match its own contig table" % path)
        g.bases2, g.lower1 = z["bases2"], z["lower1"]
        g._exc = list(zip(z["exc_start"].tolist(), z["exc_end"].tolist(), z["exc_byte"].tolist()))
        g.source = z["source"].tolist() if "source" in z.files else []
        return g

    @staticmethod
    def fasta_fingerprint(path):
        """[size, mtime_ns, crc32 of the first MB, crc32 of the last MB]: cheap, and different for a different genome copied
        over the same path whatever its time stamps say."""
        import os
        import zlib
        st = os.stat(path)
        with open(path, "rb") as f:
            head = f.read(1 << 20)
            f.seek(max(0, st.st_size - (1 << 20)))
            tail = f.read(1 << 20)
        return [int(st.st_size), int(st.st_mtime_ns), zlib.crc32(head), zlib.crc32(tail)]

    @classmethod
    def from_fasta_cached(cls, path, cache=None, lib=None):
        """from_fasta, keeping the packed image next to the FASTA (<fasta>.tcb200.npz, like pyfaidx keeps its .fai
        there): loading the image is faster still than packing the FASTA (with `lib`, the library's C++ packer; else numpy).  The image is rebuilt unless it was made from
        this very file (size, mtime and checksums of its first and last MB); a read-only directory just means no cache.
        TC_GENOME_CACHE=0 turns it off."""
        import os
        build = (lambda: cls.from_fasta_lib(path, lib)) if lib is not None else (lambda: cls.from_fasta(path))
        if os.environ.get("TC_GENOME_CACHE", "1") == "0":
            return build()
        cache = cache or str(path) + ".tcb200.npz"
        fp = cls.fasta_fingerprint(path)
        try:
            if os.path.exists(cache):
                g = cls.load(cache)
                if g.source == fp:
                    print("Using the packed genome image %s" % cache)
                    return g
        except Exception:
            pass
        g = build()
        try:
            tmp = cache + ".tmp%d" % os.getpid()
            g.save(tmp, source=fp)
            os.replace(tmp, cache)
            print("Wrote the packed genome image %s (reused while the FASTA stays the same file)" % cache)
        except Exception:
            pass
        return g

    @classmethod
    def from_windows(cls, chrom_len, windows, fill="N"):
        """Contigs that are `fill` everywhere except the given {name: [[start(1-based), bases]]}."""
        names = list(chrom_len.keys())
        g = cls(names, [chrom_len[n] for n in names])
        fb = ord(fill)
        for i, n in enumerate(names):
            off, L = int(g.chrom_off[i]), int(chrom_len[n])
            cur = 1
            for s, b in sorted((int(s), b) for s, b in windows.get(n, [])):
                if s > cur:
                    g._exc.append((off + cur - 1, off + s - 1, fb))
                a = np.frombuffer(b.encode("ascii"), dtype=np.uint8)
                g._put_slow(off + s - 1, a)
                cur = s + len(b)
            if cur <= L:
                g._exc.append((off + cur - 1, off + L, fb))
        return g

    # ------------------------------------------------------------------ accessors
    def exceptions(self):
        ex = sorted(self._exc)
        st = np.asarray([e[0] for e in ex], dtype=np.int64)
        en = np.asarray([e[1] for e in ex], dtype=np.int64)
        by = np.asarray([e[2] for e in ex], dtype=np.uint8)
        return st, en, by

    def keys(self):
        return list(self.names)

    def get_seq(self, chrom, start, end):
        """pyfaidx `get_seq(chrom, start, end).seq` from the packed image: 1-based inclusive, case and literal
        (non-ACGT) bases preserved, silently truncated past the contig end; start < 1 or an unknown contig raise."""
        import bisect
        if chrom not in self.index:
            raise KeyError("Requested rname %s does not exist!" % chrom)
        if start < 1:
            raise IndexError("Requested start coordinate must be greater than 0.")
        c = self.index[chrom]
        end = min(int(end), int(self.chrom_len[c]))
        if end < start:
            return ""
        if getattr(self, "_exc_sorted_n", -1) != len(self._exc):
            self._exc_sorted = sorted(self._exc)
            self._exc_starts = [e[0] for e in self._exc_sorted]
            self._exc_sorted_n = len(self._exc)
        off = int(self.chrom_off[c]) - 1
        out = bytearray()
        for p in range(off + start, off + end + 1):
            k = bisect.bisect_right(self._exc_starts, p) - 1
            if k >= 0 and p < self._exc_sorted[k][1]:
                out.append(self._exc_sorted[k][2])
                continue
            code = (int(self.bases2[p >> 4]) >> ((p & 15) * 2)) & 3
            low = (int(self.lower1[p >> 5]) >> (p & 31)) & 1
            out.append(b"ACGT"[code] | (low << 5))
        return out.decode("ascii")

    def nbytes(self):
        return self.bases2.nbytes + self.lower1.nbytes

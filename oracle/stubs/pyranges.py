"""Stand-in for `pyranges` (pinned ==0.0.127 by the reference, absent from this image).

TEST INFRASTRUCTURE ONLY.  Covers exactly the calls the reference makes:
  TranscriptClean.py:746-758  read_bed(f).drop_duplicate_positions(), PyRanges()
  TranscriptClean.py:761      .length
  TranscriptClean.py:1241-45  PyRanges(chromosomes,starts,ends,strands).nearest(ref).df.iloc[0]
  TranscriptClean.py:1256-59  row.Chromosome / .Start / .Start_b / .End_b

PARITY UNPINNED for two points no reference test discriminates:
  STRAND_MODE  "same"   - nearest() only considers intervals on the query's strand
                          (pyranges documents strandedness=None as "auto": same-strand
                          when both sides carry +/- strands, which is always the case here;
                          it is also what the bedtools call it replaced did: closest -s)
               "ignore" - strand is ignored
  TIE          "left"   - when the previous and next site are equidistant the lower
                          coordinate wins (bedtools closest -t first); "right" = the higher.
The same two switches exist in oracle/tc_oracle.py and in the CUDA junction kernel.
"""
import bisect

STRAND_MODE = "same"
TIE = "left"


class _Row:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class _ILoc:
    def __init__(self, rows):
        self.rows = rows

    def __getitem__(self, i):
        return self.rows[i]  # IndexError on an empty frame, like pandas


class _DF:
    def __init__(self, rows):
        self.iloc = _ILoc(rows)


class PyRanges:
    def __init__(self, chromosomes=None, starts=None, ends=None, strands=None):
        self.rows = []
        if chromosomes is not None:
            strands = strands if strands is not None else ["."] * len(chromosomes)
            self.rows = list(zip(chromosomes, starts, ends, strands))
        self._index = None

    @property
    def length(self):
        return sum(e - s for _, s, e, _ in self.rows)

    def __len__(self):
        return len(self.rows)

    @property
    def df(self):
        return _DF([_Row(Chromosome=c, Start=s, End=e, Strand=st) for c, s, e, st in self.rows])

    def drop_duplicate_positions(self):
        out = PyRanges()
        seen = set()
        for r in self.rows:
            if r not in seen:
                seen.add(r)
                out.rows.append(r)
        return out

    def _build(self):
        idx = {}
        for c, s, e, st in self.rows:
            for key in ((c, st), (c, None)):
                idx.setdefault(key, []).append((s, e))
        for k in idx:
            idx[k].sort()
        self._index = idx

    def nearest(self, other):
        if other._index is None:
            other._build()
        out = []
        for c, s, e, st in self.rows:
            key = (c, st) if STRAND_MODE == "same" else (c, None)
            cand = other._index.get(key, [])
            if not cand:
                continue
            i = bisect.bisect_left(cand, (s, -1))
            best = None
            if i < len(cand):
                best = cand[i]
            if i > 0:
                left = cand[i - 1]
                if best is None:
                    best = left
                else:
                    dl, dr = s - left[0], best[0] - s
                    if dl < dr or (dl == dr and TIE == "left"):
                        best = left
            out.append(_Row(Chromosome=c, Start=s, End=e, Strand=st,
                            Start_b=best[0], End_b=best[1]))
        res = PyRanges()
        res._rows_b = out
        res.__class__ = _Joined
        return res


class _Joined(PyRanges):
    @property
    def df(self):
        return _DF(self._rows_b)


def read_bed(path):
    rows = []
    with open(path) as f:
        for line in f:
            p = line.rstrip("\n").split("\t")
            if len(p) < 6:
                continue
            rows.append((p[0], int(p[1]), int(p[2]), p[5]))
    if not rows:
        raise IndexError("empty bed file")  # what the reference catches (TC.py:752)
    return PyRanges([r[0] for r in rows], [r[1] for r in rows],
                    [r[2] for r in rows], [r[3] for r in rows])

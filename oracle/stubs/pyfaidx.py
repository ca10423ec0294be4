"""Stand-in for `pyfaidx` (pinned >0.5 / ==0.5.8 by the reference, absent from this image).

TEST INFRASTRUCTURE ONLY.  Lets the *unmodified* reference under /root/reference
import and run in the build container so that golden vectors can be generated
(oracle/make_golden.py) and the restatement in oracle/tc_oracle.py can be pinned.

Published pyfaidx semantics restated here (call sites: TranscriptClean.py:221,469,
1016,1036,1142,1441,1457,1622; transcript.py:300,323; intronBound.py:36-44):
  * Fasta(path).keys()            -> first word of each header line
  * get_seq(name, start, end).seq -> 1-based inclusive slice, case preserved,
                                     silently truncated / empty past the contig end
  * start < 1                     -> FetchError ("Requested start coordinate must be greater than 0")
  * unknown contig                -> FetchError
A genome can also be registered in memory: REGISTRY[path] = object with
.names() and .fetch(name, start, end) (used for sparse reconstructed genomes).
"""

REGISTRY = {}


class FetchError(IndexError):
    pass


class Sequence:
    __slots__ = ("name", "seq", "start", "end")

    def __init__(self, name, seq, start, end):
        self.name, self.seq, self.start, self.end = name, seq, start, end

    def __str__(self):
        return self.seq


class _DictGenome:
    def __init__(self, seqs):
        self.seqs = seqs

    def names(self):
        return list(self.seqs.keys())

    def fetch(self, name, start, end):
        s = self.seqs[name]
        return s[start - 1:max(end, start - 1)]


def _load_fasta(path):
    seqs, name, chunks = {}, None, []
    with open(path) as f:
        for line in f:
            if line.startswith(">"):
                if name is not None:
                    seqs[name] = "".join(chunks)
                name, chunks = line[1:].split()[0], []
            else:
                chunks.append(line.strip())
    if name is not None:
        seqs[name] = "".join(chunks)
    return _DictGenome(seqs)


class Fasta:
    def __init__(self, filename, **kw):
        self.filename = filename
        if filename not in REGISTRY:
            REGISTRY[filename] = _load_fasta(filename)
        self._g = REGISTRY[filename]

    def keys(self):
        return self._g.names()

    def get_seq(self, name, start, end, rc=False):
        if name not in self._g.names():
            raise FetchError("Requested rname {0} does not exist!".format(name))
        if start < 1:
            raise FetchError("Requested start coordinate must be greater than 0.")
        return Sequence(name, self._g.fetch(name, start, end), start, end)

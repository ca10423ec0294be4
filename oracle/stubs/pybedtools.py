"""Stand-in for `pybedtools` (+ the bedtools / samtools binaries), absent from this image.

TEST INFRASTRUCTURE ONLY.  The reference uses it to iterate VCF records
(TranscriptClean.py:495-504, 797-819) and to pre-filter them to those that overlap the
reads (`intersect(bam, u=True)`, TC.py:806).  The pre-filter never changes a lookup result
(a variant that a read can hit necessarily overlaps that read), so intersect() is identity.
"""
import gzip


class _Rec:
    def __init__(self, fields):
        self.fields = fields
        self.chrom = fields[0]

    def __getitem__(self, i):
        return self.fields[i]


class BedTool:
    def __init__(self, path):
        self.path = path
        opener = gzip.open if str(path).endswith(".gz") else open
        self.recs = []
        with opener(path, "rt") as f:
            for line in f:
                if line.startswith("#") or not line.strip():
                    continue
                self.recs.append(_Rec(line.rstrip("\n").split("\t")))

    def intersect(self, *a, **kw):
        return self

    def __iter__(self):
        return iter(self.recs)

    def __len__(self):
        return len(self.recs)

"""Generates tests/golden/*.json.gz by running the UNMODIFIED reference (oracle/ref_harness.py).

TEST INFRASTRUCTURE ONLY; runs in the build container (needs /root/reference).  Usage:
    python -m oracle.make_golden

Two families of vectors:

1. `fixtures.json.gz` - the reference's own fixture reads:
   * tests/input_files/sams/deletion_insertion_mismatch_nc.sam (tagged twin under archived/)
     vs testing_suite/answer_files/DIM_nc_full/TC_clean.sam and the log / TE rows asserted in
     testing_suite/test_example_transcripts.py:166-190;
   * the 12 input/answer pairs of archived/test_TranscriptClean/sam_files (run in variant-aware
     mode like archived/test_TranscriptClean/test_TranscriptClean.py:120-133 does);
   * testing_suite/input_files/vcf_test/chr11_variantDel_andMismatch.sam with the expectations
     of testing_suite/test_example_varDM_monoexon.py:50-85 and test_full_dryrun.py:39-64.
   The hg38 FASTA blobs are not shipped (.MISSING_LARGE_BLOBS), so the genome under each read is
   reconstructed from the fixture itself: exon bases from SEQ + MD of the input line and of the
   golden answer line, intron motifs from the answer's jM codes; every other base is 'N'.  The
   real GM12878_SJs_chr1.tab and GM12878_chr1.vcf.gz are used (rows near the reads are copied
   into the fixture).  `ref_answer_line` keeps the reference's golden answer and
   `answer_matches` records whether the reference run reproduces it byte-for-byte.

2. `synth_<seed>.json.gz` - seeded synthetic cases (tests/synth.py) with the reference's four
   output texts for several option sets, and the --dryRun texts.
"""
import gzip
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_harness as rh   # noqa: E402

TS = rh.REF_ROOT + "/TranscriptClean/testing_suite"
AR = rh.REF_ROOT + "/TranscriptClean/archived/test_TranscriptClean/sam_files"
EX = rh.REF_ROOT + "/TranscriptClean/example_files"
MOTIFS = {1: ("GT", "AG"), 2: ("CT", "AC"), 3: ("GC", "AG"), 4: ("CT", "GC"), 5: ("AT", "AC"),
          6: ("GT", "AT")}


class SparseGenome:
    """pos -> base with 'N' elsewhere; fetch() as the pyfaidx stub expects."""

    def __init__(self, length):
        self.length, self.known = length, {}

    def put(self, pos, base, overwrite=True):
        if overwrite or pos not in self.known:
            self.known[pos] = base

    def fetch(self, start, end):
        end = min(end, self.length)
        return "".join(self.known.get(p, "N") for p in range(start, end + 1))

    def windows(self):
        out, cur, start, prev = [], [], None, None
        for p in sorted(self.known):
            if prev is not None and p == prev + 1:
                cur.append(self.known[p])
            else:
                if cur:
                    out.append([start, "".join(cur)])
                start, cur = p, [self.known[p]]
            prev = p
        if cur:
            out.append([start, "".join(cur)])
        return out


def first_read(path):
    with open(path) as f:
        for line in f:
            if not line.startswith("@") and line.strip():
                return line.strip()


def learn_exons(g, line, overwrite):
    """Reference bases under a tagged SAM line (CIGAR + MD + SEQ)."""
    f = line.split("\t")
    pos, cigar, seq = int(f[3]), f[5], f[9]
    md = [t for t in f[11:] if t.startswith("MD:Z:")]
    if not md:
        return
    md = md[0][5:]
    toks = re.findall(r"[0-9]+|\^[A-Za-z]+|[A-Za-z]", md)
    # expand MD into per-reference-base stream: None = match, letter = mismatch ref, ('D', x)
    stream = []
    for t in toks:
        if t[0].isdigit():
            stream += [None] * int(t)
        elif t[0] == "^":
            stream += [("D", c) for c in t[1:]]
        else:
            stream.append(t)
    si, q, gp = 0, 0, pos
    for ct, op in re.findall(r"([0-9]+)([A-Z])", cigar):
        ct = int(ct)
        if op == "M":
            for _ in range(ct):
                s = stream[si]
                si += 1
                g.put(gp, seq[q] if s is None else s, overwrite)
                q += 1
                gp += 1
        elif op == "D":
            for _ in range(ct):
                g.put(gp, stream[si][1], overwrite)
                si += 1
                gp += 1
        elif op in ("I", "S"):
            q += ct
        elif op in ("N", "H"):
            gp += ct


def learn_motifs(g, line, overwrite):
    f = line.split("\t")
    jm = [t for t in f[11:] if t.startswith("jM:B:c,")]
    ji = [t for t in f[11:] if t.startswith("jI:B:i,")]
    if not jm or not ji:
        return
    codes = [int(x) for x in jm[0].split(",")[1:]]
    pos = [int(x) for x in ji[0].split(",")[1:]]
    for k, code in enumerate(codes):
        if code < 0 or code % 20 == 0:
            continue
        left, right = MOTIFS[code % 20]
        a, b = pos[2 * k], pos[2 * k + 1]
        g.put(a, left[0], overwrite)
        g.put(a + 1, left[1], overwrite)
        g.put(b - 1, right[0], overwrite)
        g.put(b, right[1], overwrite)


def near_rows(path, chrom, lo, hi, col, opener=open, skip_hash=False, pad=200000):
    out = []
    with opener(path, "rt") as f:
        for line in f:
            if skip_hash and line.startswith("#"):
                continue
            p = line.rstrip("\n").split("\t")
            if p[0] == chrom and lo - pad <= int(p[col]) <= hi + pad:
                out.append(p)
    return out


def span(line):
    f = line.split("\t")
    pos = int(f[3])
    end = pos + sum(int(c) for c, o in re.findall(r"([0-9]+)([A-Z])", f[5]) if o in "MDNH")
    return f[2], pos, end


def write_tmp(rows, path):
    with open(path, "w") as fh:
        for r in rows:
            fh.write("\t".join(r) + "\n")
    return path


def run_fixture(name, in_line, answer_line, sj_path, vcf_path, opts, chrom_len=250000000,
                expect=None):
    chrom, lo, hi = span(in_line)
    g = SparseGenome(chrom_len)
    if answer_line:
        learn_motifs(g, answer_line, True)
        learn_exons(g, answer_line, True)
    learn_motifs(g, in_line, False)
    learn_exons(g, in_line, True if not answer_line else False)
    sj_rows = near_rows(sj_path, chrom, lo, hi, 1) if sj_path else []
    vcf_rows = []
    if vcf_path:
        op = gzip.open if vcf_path.endswith(".gz") else open
        vcf_rows = [r[:5] for r in near_rows(vcf_path, chrom, lo, hi, 1, op, True, pad=1000)]
    tmp = "/tmp/tc_golden_%d" % os.getpid()
    os.makedirs(tmp, exist_ok=True)
    sjf = write_tmp(sj_rows, tmp + "/sj.tab") if sj_path else None
    vcf = write_tmp(vcf_rows, tmp + "/v.vcf") if vcf_path else None
    out = rh.run_reference([in_line], {chrom: g}, sjf, vcf, **opts)
    dry = rh.run_reference([in_line], {chrom: g}, dryRun=True)
    rec = {"name": name, "chrom_len": {chrom: chrom_len}, "windows": {chrom: g.windows()},
           "sj_rows": sj_rows, "vcf_rows": vcf_rows, "lines": [in_line], "opts": opts,
           "expected": out, "expected_dry": dry, "ref_answer_line": answer_line,
           "answer_matches": None, "expect_fields": expect}
    if answer_line:
        rec["answer_matches"] = (out["sam"].rstrip("\n").split() == answer_line.split())
    return rec


def fixtures():
    sj = TS + "/input_files/GM12878_SJs_chr1.tab"
    vcf = EX + "/GM12878_chr1.vcf.gz"
    recs = []
    base = dict(maxLenIndel=5, maxSJOffset=5)
    # DIM_nc_full (test_example_transcripts.py:97-198): --primaryOnly --canonOnly, no variants
    recs.append(run_fixture(
        "DIM_nc_full", first_read(AR + "/deletion_insertion_mismatch_nc.sam"),
        first_read(TS + "/answer_files/DIM_nc_full/TC_clean.sam"), sj, None,
        dict(base, primaryOnly=True, canonOnly=True),
        expect={"log": "c34150/f1p1/3707\tprimary\t2\t0\t0\t1\t0\t0\t2\t0\t1\t0",
                "te": ["c34150/f1p1/3707\tchr1_150944919\tMismatch\t1\tCorrected\tNA",
                       "c34150/f1p1/3707\tchr1_150944920\tMismatch\t1\tCorrected\tNA",
                       "c34150/f1p1/3707\tchr1_150943033_150943034\tInsertion\t1\tCorrected\tNA",
                       "c34150/f1p1/3707\tchr1_150949256_150949257\tDeletion\t1\tCorrected\tNA",
                       "c34150/f1p1/3707\tchr1_150950682_150950683\tDeletion\t1\tCorrected\tNA",
                       "c34150/f1p1/3707\tchr1_150943994_150944918\tNC_SJ_boundary\t1\tCorrected\tNA"]}))
    pairs = [("mismatch_input", "mismatch_correctAnswer"), ("insertion_input", "insertion_correctAnswer"),
             ("deletion_input", "deletion_correctAnswer"),
             ("deletion_insertion_mismatch", "deletion_insertion_mismatch_correctAnswer"),
             ("deletion_insertion_mismatch_nc", "deletion_insertion_mismatch_nc_correctAnswer"),
             ("deletion_variant", "deletion_variant_correctAnswer"),
             ("nc_case1", "nc_case1_correctAnswer"), ("nc_case2", "nc_case2_correctAnswer"),
             ("nc_case3", "nc_case3_correctAnswer"), ("nc_case4", "nc_case4_correctAnswer"),
             ("nc_case5", "nc_case5_correctAnswer"), ("insertion_variant", None)]
    for a, b in pairs:
        ans = first_read(AR + "/" + b + ".sam") if b else None
        inp = first_read(AR + "/" + a + ".sam")
        if a == "insertion_variant":
            ans = inp     # archived test: variant insertion must be left alone (answer = input)
        recs.append(run_fixture("archived/" + a, inp, ans, sj, vcf, dict(base)))
    # tag-less twins: NM/MD/jM/jI must be generated (test_compute_MD_NM.py, test_compute_jI_jM.py)
    for a, tagged in (("deletion_insertion_mismatch_nc_noExtraTags", "deletion_insertion_mismatch_nc"),
                      ("perfectReferenceMatch_twoIntrons_noExtraTags", "perfectReferenceMatch_twoIntrons"),
                      ("perfectReferenceMatch_noIntrons_noExtraTags", "perfectReferenceMatch_noIntrons")):
        chrom, lo, hi = span(first_read(AR + "/" + tagged + ".sam"))
        g = SparseGenome(250000000)
        learn_motifs(g, first_read(AR + "/" + tagged + ".sam"), True)
        learn_exons(g, first_read(AR + "/" + tagged + ".sam"), True)
        line = first_read(AR + "/" + a + ".sam")
        sj_rows = near_rows(sj, chrom, lo, hi, 1)
        sjf = write_tmp(sj_rows, "/tmp/tc_golden_sj.tab")
        tagged_line = first_read(AR + "/" + tagged + ".sam")
        expect = None
        if "nc_noExtraTags" in a:   # NM/MD must come out as in the tagged twin (test_compute_MD_NM.py)
            opts = dict(base, correctMismatches=False, correctIndels=False, correctSJs=False)
            tags = tagged_line.split("\t")[11:]
            expect = {"nm": [t for t in tags if t.startswith("NM")][0],
                      "md": [t for t in tags if t.startswith("MD")][0]}
        else:
            opts = dict(base)
        out = rh.run_reference([line], {chrom: g}, sjf, None, **opts)
        recs.append({"name": "tagless/" + a, "chrom_len": {chrom: 250000000},
                     "windows": {chrom: g.windows()}, "sj_rows": sj_rows, "vcf_rows": [],
                     "lines": [line], "opts": opts, "expected": out,
                     "expected_dry": rh.run_reference([line], {chrom: g}, dryRun=True),
                     "ref_answer_line": first_read(AR + "/" + tagged + ".sam"),
                     "answer_matches": out["sam"].split() == first_read(AR + "/" + tagged + ".sam").split(),
                     "expect_fields": expect})
    # variant-aware mono-exon read (test_example_varDM_monoexon.py, test_full_dryrun.py)
    line = first_read(TS + "/input_files/vcf_test/chr11_variantDel_andMismatch.sam")
    f = line.split("\t")
    # answer assembled from the assertions of test_example_varDM_monoexon.py:50-85 (the real
    # genome has C at chr11:207698 although the hand-written MD says T)
    ans = "\t".join(f[:9] + ["GCGGGCC", "*", "NM:i:1", "MD:Z:2^G5", "jM:B:c,-1", "jI:B:i,-1"])
    recs.append(run_fixture("varDM_monoexon", line, ans, TS + "/input_files/chr11_sjs.txt",
                            TS + "/input_files/vcf_test/chr11_variantDel.vcf",
                            dict(base, primaryOnly=True), chrom_len=135086622,
                            expect={"log": "m54284_181015_235905/15205058/29_3462_CCS-PB72-derived\tprimary\t0\t0\t1\t0\t0\t0\t1\t0\t0\t0",
                                    "dry_log": "m54284_181015_235905/15205058/29_3462_CCS-PB72-derived\tprimary\tNA\t1\tNA\tNA\t0\tNA\tNA\t1\tNA\tNA",
                                    "dry_te": ["m54284_181015_235905/15205058/29_3462_CCS-PB72-derived\tchr11_207693_207694\tDeletion\t1\tUncorrected\tDryRun",
                                               "m54284_181015_235905/15205058/29_3462_CCS-PB72-derived\tchr11_207698\tMismatch\t1\tUncorrected\tDryRun"]}))
    return recs


def synth_case(seed, n_reads):
    import synth
    case = synth.make_case(seed, n_reads=n_reads, unpinned=False)
    tmp = "/tmp/tc_golden_%d" % os.getpid()
    os.makedirs(tmp, exist_ok=True)
    case.write_sj(tmp + "/sj.tab")
    case.write_vcf(tmp + "/v.vcf")
    runs = []
    for opts, use_sj, use_vcf in (
            (dict(maxLenIndel=5, maxSJOffset=5), True, False),
            (dict(maxLenIndel=5, maxSJOffset=5), True, True),
            (dict(maxLenIndel=3, maxSJOffset=8, canonOnly=True), True, True),
            (dict(maxLenIndel=5, maxSJOffset=5, primaryOnly=True, correctMismatches=False), False, False),
            (dict(maxLenIndel=5, maxSJOffset=5, correctIndels=False, correctSJs=False), True, True)):
        out = rh.run_reference(case.lines, case.genome, tmp + "/sj.tab" if use_sj else None,
                               tmp + "/v.vcf" if use_vcf else None, **opts)
        runs.append({"opts": opts, "use_sj": use_sj, "use_vcf": use_vcf, "expected": out})
    clean = synth.make_case(seed, n_reads=n_reads, quirks=False)
    dry = rh.run_reference(clean.lines, clean.genome, dryRun=True)
    return {"seed": seed, "n_reads": n_reads, "genome": case.genome,
            "sj_rows": [[str(x) for x in r] for r in case.sj_rows],
            "vcf_rows": [[str(x) for x in r] for r in case.vcf_rows], "lines": case.lines,
            "runs": runs, "dry_lines": clean.lines, "dry_genome": clean.genome, "expected_dry": dry}


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    recs = fixtures()
    for r in recs:
        print("%-55s answer_matches=%s" % (r["name"], r["answer_matches"]))
    with gzip.open(os.path.join(out_dir, "fixtures.json.gz"), "wt") as f:
        json.dump(recs, f)
    for seed in (101, 102, 103, 104):
        rec = synth_case(seed, 60)
        with gzip.open(os.path.join(out_dir, "synth_%d.json.gz" % seed), "wt") as f:
            json.dump(rec, f)
        print("synth", seed, "reads", len(rec["lines"]))


if __name__ == "__main__":
    main()

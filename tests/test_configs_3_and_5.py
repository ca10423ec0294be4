"""BASELINE.json configs[2] and configs[4] against the ORACLE (not GPU against GPU):

  configs[2]  variant-aware mode over an hg38-scale image: four contigs, 2.4 Gb of them filler so that every plane
              index of the reads' contigs lies beyond 2^31, and a catalogue of more than a million VCF records
              (85 % SNPs, 7.5 % insertions, 7.5 % deletions, multi-allelic rows; 30 % of the reads' own errors planted)
  configs[4]  --dryRun inventory, then the full correction, of a PacBio + ONT mix on that image, sharded over every
              visible GPU by shard_bounds (split_input's rule, TC.py:544-557), the shards' texts concatenated in order

The same drivers run on the CPU twin of the library with small fillers (`-m "not gpu"`), so that the test code itself
is exercised in the build container."""
import os
import tempfile

import numpy as np
import pytest

import goldens
import pipeline
from tools.workload import Workload, variants_from_reads
from transcriptclean_b200 import api, shard
from transcriptclean_b200.genome import GenomeImage
from transcriptclean_b200.tables import build_sj_tables, build_variant_tables


class Setup:
    def __init__(self, filler, n_random_variants):
        self.wa = Workload(seed=301, chrom_len=6_000_000, n_genes=80, chrom="chr1")
        self.wb = Workload(seed=302, chrom_len=5_000_000, n_genes=70, chrom="chr2")
        self.names = ["chrFillA", "chr1", "chrFillB", "chr2"]
        self.lens = [filler, self.wa.chrom_len, filler - 1000, self.wb.chrom_len]
        self.gi = GenomeImage(self.names, self.lens)            # fillers: all 'A' (zero words)
        self.gi._put(int(self.gi.chrom_off[1]), self.wa.genome)
        self.gi._put(int(self.gi.chrom_off[3]), self.wb.genome)
        self.oracle_genome = {"chr1": self.wa.genome.tobytes().decode(), "chr2": self.wb.genome.tobytes().decode()}
        self.dir = tempfile.mkdtemp(prefix="tc_cfg_")
        self.sjf = goldens.write_rows(self.wa.sj_rows() + self.wb.sj_rows(), self.dir + "/sj.tab")
        self.n_random_variants = n_random_variants

    def reads(self, shape_a, shape_b, n, seed_offset):
        """n reads of contig chr1 in shape_a interleaved with n reads of chr2 in shape_b -> (lines, planted VCF rows)."""
        ca, nma = self.wa.reads(shape_a, n, seed_offset=seed_offset)
        cb, nmb = self.wb.reads(shape_b, n, seed_offset=seed_offset + 1)
        la, lb = self.wa.sam_lines(ca, nma, 0, n), self.wb.sam_lines(cb, nmb, 0, n)
        lines = [x for pair in zip(la, lb) for x in pair]
        rows = variants_from_reads(self.wa, ca, 0, n, frac=0.3, seed=5) + variants_from_reads(self.wb, cb, 0, n, frac=0.3, seed=6)
        return lines, rows

    def vcf(self, planted):
        """planted rows + n_random_variants random records over all four contigs, in the proportions of GM12878_chr1.vcf.gz."""
        rng = np.random.default_rng(99)
        n = self.n_random_variants
        contig = rng.integers(0, 4, n)
        pos = (rng.random(n) * (np.asarray(self.lens)[contig] - 20)).astype(np.int64) + 5
        kind = rng.random(n)
        bases = np.frombuffer(b"ACGT", dtype=np.uint8)
        rows = []
        ref_b = bases[rng.integers(0, 4, n)]
        alt_b = bases[rng.integers(0, 4, n)]
        size = rng.integers(1, 8, n)
        multi = rng.random(n) < 0.006
        for k in range(n):
            c, p, r, a = self.names[contig[k]], int(pos[k]), chr(ref_b[k]), chr(alt_b[k])
            if kind[k] < 0.85:
                alt = a + ("," + "ACGT"[k & 3] if multi[k] else "")
                rows.append((contig[k], p, "%s\t%d\t.\t%s\t%s\t.\tPASS\t.\n" % (c, p, r, alt)))
            elif kind[k] < 0.925:
                rows.append((contig[k], p, "%s\t%d\t.\t%s\t%s\t.\tPASS\t.\n" % (c, p, r, r + "ACGTACG"[:size[k]])))
            else:
                rows.append((contig[k], p, "%s\t%d\t.\t%s\t%s\t.\tPASS\t.\n" % (c, p, r + "TGCATGC"[:size[k]], r)))
        idx = {nm: i for i, nm in enumerate(self.names)}
        for r in planted:
            rows.append((idx[r[0]], int(r[1]), "\t".join(str(x) for x in r) + "\t.\tPASS\t.\n"))
        rows.sort(key=lambda x: (x[0], x[1]))
        path = self.dir + "/catalogue.vcf"
        with open(path, "w") as f:
            f.write("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
            f.writelines(r[2] for r in rows)
        return path, len(rows)

    def close(self):
        self.wa.close(); self.wb.close()


def run_config3(engine, filler, n_random_variants, n_reads):
    s = Setup(filler, n_random_variants)
    try:
        lines, planted = s.reads("pacbio", "pacbio", n_reads // 2, 21)
        vcf, n_rec = s.vcf(planted)
        var = build_variant_tables(vcf, 5, s.gi.index)
        assert var.n_records == n_rec and len(var.snp_key) > 0.8 * n_random_variants
        refs = api.make_refs(s.gi, build_sj_tables(s.sjf, {"chr1", "chr2"}, s.gi.index), var, engine=engine)
        got = api.correct_batch_text(lines, api.Struct(), refs)
        want = pipeline.oracle_texts(s.oracle_genome, lines, s.sjf, vcf, {})
        d = pipeline.first_diff(got, want)
        assert d is None, d
        assert want["te"].count("VariantMatch") > n_reads and want["te"].count("Corrected") > 3 * n_reads
        return s.gi
    finally:
        s.close()


def run_config5(engines, filler, n_reads):
    s = Setup(filler, 0)
    try:
        lines, _ = s.reads("pacbio", "ont", n_reads // 2, 31)
        sj = build_sj_tables(s.sjf, {"chr1", "chr2"}, s.gi.index)
        refs_list = [api.make_refs(s.gi, sj, None, engine=e) for e in engines]
        assert shard.shard_bounds(len(lines), len(engines))[-1][1] == len(lines)
        dry = shard.correct_sharded_threads(lines, None, refs_list, lambda ls, o, refs: api.dry_run_text(ls, refs))
        want = pipeline.oracle_texts(s.oracle_genome, lines, None, None, {}, dry=True)
        assert (dry["log"], dry["te"]) == (want["log"], want["te"])
        opts = api.Struct(canonOnly=True)
        full = shard.correct_sharded_threads(lines, opts, refs_list, api.correct_batch_text)
        want = pipeline.oracle_texts(s.oracle_genome, lines, s.sjf, None, {"canonOnly": True})
        d = pipeline.first_diff(full, want)
        assert d is None, d
        assert want["te"].count("TooLarge") > n_reads // 20
    finally:
        s.close()


def test_config3_driver_on_the_cpu_twin():
    import hostsim_engine
    eng = hostsim_engine.HostSimEngine()
    try:
        run_config3(eng, 3_000_000, 20000, 300)
    finally:
        eng.close()


def test_config5_driver_on_the_cpu_twin():
    import hostsim_engine
    engines = [hostsim_engine.HostSimEngine() for _ in range(3)]
    try:
        run_config5(engines, 3_000_000, 240)
    finally:
        for e in engines:
            e.close()


@pytest.mark.gpu
def test_config3_variant_aware_hg38_scale_image_vs_oracle():
    from transcriptclean_b200.engine import Engine
    eng = Engine(0)
    try:
        gi = run_config3(eng, 1_200_000_000, 1_100_000, 4000)
        assert int(gi.chrom_off[3]) > 2 ** 31 and gi.n_pad > 2_400_000_000
    finally:
        eng.close()


@pytest.mark.gpu
def test_config5_dry_run_then_correction_sharded_over_all_gpus_vs_oracle():
    import torch
    from transcriptclean_b200.engine import Engine
    engines = [Engine(d) for d in range(max(1, torch.cuda.device_count()))]
    try:
        run_config5(engines, 1_200_000_000, 4000)
    finally:
        for e in engines:
            e.close()

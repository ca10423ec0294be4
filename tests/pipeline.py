"""Shared test driver: runs a case through the product host code + an engine, and through the
oracle, and returns both sets of texts."""
import os
import tempfile

import goldens
from oracle import tc_oracle as oc
from transcriptclean_b200 import api
from transcriptclean_b200.genome import GenomeImage
from transcriptclean_b200.tables import build_sj_tables, build_variant_tables


def write_tables(sj_rows, vcf_rows):
    d = tempfile.mkdtemp(prefix="tc_case_")
    sjf = goldens.write_rows(sj_rows, d + "/sj.tab") if sj_rows is not None else None
    vcf = goldens.write_rows(vcf_rows, d + "/v.vcf") if vcf_rows else None
    return sjf, vcf


def product_texts(engine, genome_image, lines, sjf, vcf, opts, dry=False):
    chroms = set(l.split("\t")[2] for l in lines)
    max_len = opts.get("maxLenIndel", 5)
    sj = var = None
    if sjf and opts.get("correctSJs", True):
        sj = build_sj_tables(sjf, chroms, genome_image.index)
    if vcf:
        var = build_variant_tables(vcf, max_len, genome_image.index)
    refs = api.make_refs(genome_image, sj, var, engine=engine)
    if dry:
        return api.dry_run_text(lines, refs)
    o = api.Struct(maxLenIndel=max_len, maxSJOffset=opts.get("maxSJOffset", 5),
                   correctMismatches=opts.get("correctMismatches", True), correctIndels=opts.get("correctIndels", True),
                   correctSJs=opts.get("correctSJs", True), primaryOnly=opts.get("primaryOnly", False),
                   canonOnly=opts.get("canonOnly", False), sjIgnoreStrand=opts.get("sjIgnoreStrand", False),
                   sjTieRight=opts.get("sjTieRight", False))
    return api.correct_batch_text(lines, o, refs)


def oracle_texts(genome_seqs, lines, sjf, vcf, opts, dry=False):
    g = oc.Genome(genome_seqs)
    if dry:
        return oc.dry_run_lines(lines, g)
    chroms = set(l.split("\t")[2] for l in lines)
    sj = var = None
    if sjf and opts.get("correctSJs", True):
        sj = oc.SjTables.from_file(sjf, chroms, "ignore" if opts.get("sjIgnoreStrand") else "same",
                                   "right" if opts.get("sjTieRight") else "left")
    if vcf:
        var = oc.VariantTables.from_file(vcf, opts.get("maxLenIndel", 5))
    oo = {k: v for k, v in opts.items() if k not in ("sjIgnoreStrand", "sjTieRight")}
    return oc.correct_lines(lines, oc.Options(**oo), g, sj, var)


def first_diff(a, b):
    for k in ("sam", "fa", "log", "te"):
        if a[k] != b[k]:
            x, y = a[k].split("\n"), b[k].split("\n")
            for i, (p, q) in enumerate(zip(x, y)):
                if p != q:
                    return "%s line %d:\n  got  %s\n  want %s" % (k, i, p[:600], q[:600])
            return "%s: %d vs %d lines" % (k, len(x), len(y))
    return None

/* synth_gen.c - seeded generator of the BASELINE.json synthetic workloads (bench tooling).
 *
 * Produces (a) an ASCII genome with soft-masked and N runs, (b) gene models whose introns carry
 * planted splice motifs (rows of a STAR SJ.out.tab), and (c) long reads in the columnar layout
 * of include/tc_b200.h: PacBio-like (1 % error, 1-3 bp indels) or ONT-like (8 % error,
 * geometric indel lengths), with junctions shifted off their annotated position, minus-strand
 * genes, unmapped / secondary records, NM + MD tags in minimap2 --MD style.
 * Deterministic for a given seed regardless of the thread count (one RNG stream per read).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t s; } rng_t;
static inline uint64_t rnd(rng_t* r) {            /* splitmix64 */
    uint64_t z = (r->s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static inline double rndu(rng_t* r) { return (double)(rnd(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline int64_t rndi(rng_t* r, int64_t lo, int64_t hi) { return lo + (int64_t)(rnd(r) % (uint64_t)(hi - lo + 1)); }
static double rndn(rng_t* r) { double u = rndu(r) + 1e-12, v = rndu(r); return sqrt(-2.0 * log(u)) * cos(6.283185307179586 * v); }

typedef struct {
    uint64_t seed;
    int64_t n_reads;
    int32_t mean_len, sd_len, n_exons;
    double err, p_mm, p_del, p_ins;
    int32_t indel_geometric;       /* 0: uniform 1..3, 1: geometric, ~10 % longer than 5 */
    double jn_shift_frac; int32_t jn_shift_max;
    double nonprimary_frac;
    int32_t with_nm_md;
} gen_params;

typedef struct {
    int64_t n_genes;
    int64_t* gene_exon_off;        /* n_genes+1 */
    int32_t* gene_strand;          /* 0 '+', 1 '-' */
    int32_t* gene_chrom;
    int64_t* exon_start; int64_t* exon_end;   /* 1-based inclusive, contig coordinates */
    int64_t n_exons_total;
} gene_models;

typedef struct {
    int64_t n; int32_t* flag; int32_t* chrom; int32_t* pos; uint8_t* tags;
    int64_t* seq_off; uint8_t* seq; int64_t* cig_off; uint8_t* cig_op; int32_t* cig_len;
    int64_t* md_off; uint8_t* md; int32_t* nm;
} read_cols;

static const char BASES[4] = {'A', 'C', 'G', 'T'};

/* ASCII genome: random bases, ~half soft-masked in runs, a few N runs. */
void tcsynth_genome(uint64_t seed, int64_t len, uint8_t* out) {
    rng_t r = {seed * 7919 + 1};
    int64_t i = 0;
    while (i < len) {
        uint64_t w = rnd(&r);
        for (int k = 0; k < 32 && i < len; k++, i++) out[i] = (uint8_t)BASES[(w >> (2 * k)) & 3];
    }
    i = 0;
    while (i < len) {               /* alternating upper / lower segments */
        int64_t seg = rndi(&r, 200, 6000); int lower = rnd(&r) & 1;
        if (lower) for (int64_t k = i; k < i + seg && k < len; k++) out[k] |= 0x20;
        i += seg;
    }
}

static char comp(char c) { switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; default: return 'A'; } }

/* Gene models + motif planting (+ N runs between genes).  Returns malloc'd arrays in *gm. */
void tcsynth_genes(uint64_t seed, uint8_t* genome, int64_t len, int32_t chrom, int64_t n_genes, int32_t min_exons,
                   double minus_frac, gene_models* gm) {
    rng_t r = {seed * 104729 + 3};
    gm->n_genes = n_genes;
    gm->gene_exon_off = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_genes + 1));
    gm->gene_strand = (int32_t*)malloc(sizeof(int32_t) * (size_t)n_genes);
    gm->gene_chrom = (int32_t*)malloc(sizeof(int32_t) * (size_t)n_genes);
    int64_t cap = n_genes * (min_exons + 8);
    gm->exon_start = (int64_t*)malloc(sizeof(int64_t) * (size_t)cap);
    gm->exon_end = (int64_t*)malloc(sizeof(int64_t) * (size_t)cap);
    int64_t spacing = len / n_genes, ne = 0;
    for (int64_t g = 0; g < n_genes; g++) {
        int nex = min_exons + (int)rndi(&r, 0, 6);
        int strand = rndu(&r) < minus_frac ? 1 : 0;
        int64_t p = g * spacing + 1000 + rndi(&r, 0, spacing / 8);
        gm->gene_exon_off[g] = ne; gm->gene_strand[g] = strand; gm->gene_chrom[g] = chrom;
        int64_t limit = (g + 1) * spacing - 2000;
        for (int e = 0; e < nex; e++) {
            int64_t el = rndi(&r, 200, 470);
            if (p + el + 400 >= limit) break;
            gm->exon_start[ne] = p; gm->exon_end[ne] = p + el - 1; ne++;
            double lg = log(300.0) + rndu(&r) * (log(20000.0) - log(300.0));
            int64_t il = (int64_t)exp(lg);
            if (p + el + il + 700 >= limit) il = 300 + rndi(&r, 0, 50);
            p += el + il;
        }
        if (ne - gm->gene_exon_off[g] < 2) { /* degenerate: force two exons */
            ne = gm->gene_exon_off[g];
            int64_t q = g * spacing + 1000;
            gm->exon_start[ne] = q; gm->exon_end[ne] = q + 299; ne++;
            gm->exon_start[ne] = q + 800; gm->exon_end[ne] = q + 1099; ne++;
        }
        for (int64_t e = gm->gene_exon_off[g]; e + 1 < ne; e++) {       /* plant motifs */
            int64_t i0 = gm->exon_end[e] + 1, i1 = gm->exon_start[e + 1] - 1;
            double u = rndu(&r);
            const char* d = u < 0.96 ? "GT" : (u < 0.99 ? "GC" : "AT");
            const char* a = u < 0.99 ? "AG" : "AC";
            char l0, l1, r0, r1;
            if (!strand) { l0 = d[0]; l1 = d[1]; r0 = a[0]; r1 = a[1]; }
            else { l0 = comp(a[1]); l1 = comp(a[0]); r0 = comp(d[1]); r1 = comp(d[0]); }
            genome[i0 - 1] = (uint8_t)l0 | (genome[i0 - 1] & 0x20); genome[i0] = (uint8_t)l1 | (genome[i0] & 0x20);
            genome[i1 - 2] = (uint8_t)r0 | (genome[i1 - 2] & 0x20); genome[i1 - 1] = (uint8_t)r1 | (genome[i1 - 1] & 0x20);
        }
        /* an N run in the gap after the gene, now and then */
        if ((rnd(&r) & 15) == 0) {
            int64_t a = limit + 200, b = a + rndi(&r, 50, 1500);
            for (int64_t k = a; k < b && k < len; k++) genome[k] = 'N';
        }
    }
    gm->gene_exon_off[n_genes] = ne; gm->n_exons_total = ne;
}

typedef struct { uint8_t* p; int64_t n, cap; } bbuf;
static void bb_need(bbuf* b, int64_t extra, size_t el) {
    if (b->n + extra > b->cap) { b->cap = (b->n + extra) * 2 + 1024; b->p = (uint8_t*)realloc(b->p, (size_t)b->cap * el); }
}

typedef struct { bbuf seq, cop, clen, md; } tbuf;

static int indel_len(rng_t* r, int geometric) {
    if (!geometric) return (int)rndi(r, 1, 3);
    int l = 1; while (rndu(r) < 0.631 && l < 40) l++;      /* P(l > 5) = 0.631^5 = 0.1 */
    return l;
}

static void push_op(tbuf* t, int64_t* nops, int op, int32_t ct) {
    if (ct <= 0) return;
    if (*nops > 0 && t->cop.p[t->cop.n - 1] == (uint8_t)op) { ((int32_t*)t->clen.p)[t->clen.n - 1] += ct; return; }
    bb_need(&t->cop, 1, 1); bb_need(&t->clen, 1, 4);
    t->cop.p[t->cop.n++] = (uint8_t)op; ((int32_t*)t->clen.p)[t->clen.n++] = ct; (*nops)++;
}
static void md_num(tbuf* t, int64_t v) {
    char tmp[24]; int n = 0; do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    bb_need(&t->md, n, 1); while (n) t->md.p[t->md.n++] = (uint8_t)tmp[--n];
}

/* One read into the thread buffers; returns 0 for a non-primary record. */
static int gen_read(const gen_params* P, const uint8_t* genome, const gene_models* gm, int64_t idx, tbuf* t,
                    int32_t* flag, int32_t* chrom, int32_t* pos, uint8_t* tags, int32_t* nm_out, int64_t* nops_out) {
    rng_t r = {P->seed * 1000003ull + (uint64_t)idx * 0x2545F4914F6CDD1Dull + 17};
    *nops_out = 0; *nm_out = 0; *tags = 0;
    if (rndu(&r) < P->nonprimary_frac) {
        if (rnd(&r) & 1) { *flag = 4; *chrom = -1; *pos = 0; } else { *flag = 256; *chrom = 0; *pos = 1; }
        return 0;
    }
    int64_t g = rndi(&r, 0, gm->n_genes - 1);
    int64_t e0 = gm->gene_exon_off[g], e1 = gm->gene_exon_off[g + 1];
    int nex = (int)(e1 - e0), want = P->n_exons < nex ? P->n_exons : nex;
    int64_t first = e0 + rndi(&r, 0, nex - want);
    int64_t bs[64], be[64];
    if (want > 64) want = 64;
    int64_t total = 0;
    for (int k = 0; k < want; k++) { bs[k] = gm->exon_start[first + k]; be[k] = gm->exon_end[first + k]; total += be[k] - bs[k] + 1; }
    int64_t L = (int64_t)(P->mean_len + P->sd_len * rndn(&r)); if (L < 200) L = 200;
    if (total > L) {
        int64_t trim = total - L, a = be[0] - bs[0] + 1 - 30, b = be[want - 1] - bs[want - 1] + 1 - 30;
        int64_t ta = trim / 2 < a ? trim / 2 : a; if (want == 1) { b = a - ta; }
        int64_t tb = trim - ta < b ? trim - ta : b;
        if (ta > 0) bs[0] += ta;
        if (tb > 0) be[want - 1] -= tb;
    }
    for (int k = 0; k + 1 < want; k++) {                     /* aligner put the junction a few bases off */
        if (rndu(&r) < P->jn_shift_frac) {
            int mode = (int)(rnd(&r) % 3); int d = (int)rndi(&r, 1, P->jn_shift_max); if (rnd(&r) & 1) d = -d;
            int64_t d0 = (mode == 0 || mode == 2) ? d : 0, d1 = (mode == 1 || mode == 2) ? d : 0;
            if (be[k] + d0 - bs[k] > 20 && be[k + 1] - (bs[k + 1] + d1) > 20) { be[k] += d0; bs[k + 1] += d1; }
        }
    }
    const int strand = gm->gene_strand[g];
    *flag = strand ? 16 : 0; *chrom = gm->gene_chrom[g]; *pos = (int32_t)bs[0];
    int64_t run = 0; int nm = 0; int64_t nops = 0;
    const double pe = P->err, pd = pe * P->p_del, pi = pe * P->p_ins, pm = pe * P->p_mm;
    for (int k = 0; k < want; k++) {
        int64_t p = bs[k];
        while (p <= be[k]) {
            double u = rndu(&r);
            int edge = (p - bs[k] < 3) || (be[k] - p < 3);
            if (!edge && u < pd) {
                int dl = indel_len(&r, P->indel_geometric); if (dl > be[k] - p - 2) dl = (int)(be[k] - p - 2);
                if (dl > 0) {
                    if (P->with_nm_md) { md_num(t, run); bb_need(&t->md, dl + 1, 1); t->md.p[t->md.n++] = '^';
                        for (int j = 0; j < dl; j++) t->md.p[t->md.n++] = genome[p - 1 + j] & 0xDF; }
                    run = 0; nm += dl; push_op(t, &nops, 'D', dl); p += dl; continue;
                }
            }
            if (!edge && u >= pd && u < pd + pi) {
                int il = indel_len(&r, P->indel_geometric);
                bb_need(&t->seq, il, 1);
                for (int j = 0; j < il; j++) t->seq.p[t->seq.n++] = (uint8_t)BASES[rnd(&r) & 3];
                nm += il; push_op(t, &nops, 'I', il);
            }
            uint8_t ref = genome[p - 1], refu = ref & 0xDF, rb = refu;
            int last = (k == want - 1 && p == be[k]);
            if (refu == 'N' && !last) rb = (uint8_t)BASES[rnd(&r) & 3];
            else if (u >= pd + pi && u < pd + pi + pm && !last) { do { rb = (uint8_t)BASES[rnd(&r) & 3]; } while (rb == refu); }
            bb_need(&t->seq, 1, 1); t->seq.p[t->seq.n++] = rb;
            if (rb != refu) { if (P->with_nm_md) { md_num(t, run); bb_need(&t->md, 1, 1); t->md.p[t->md.n++] = refu; } run = 0; nm++; }
            else run++;
            push_op(t, &nops, 'M', 1); p++;
        }
        if (k + 1 < want) push_op(t, &nops, 'N', (int32_t)(bs[k + 1] - be[k] - 1));
    }
    if (P->with_nm_md) { md_num(t, run); *tags = 3; }
    *nm_out = nm; *nops_out = nops;
    return 1;
}

/* Generates P->n_reads reads; arrays in *out are malloc'd (free with tcsynth_free_reads). */
int tcsynth_reads(const gen_params* P, const uint8_t* genome, int64_t genome_len, const gene_models* gm, read_cols* out) {
    (void)genome_len;
    const int64_t n = P->n_reads;
    out->n = n;
    out->flag = (int32_t*)malloc(4 * (size_t)(n + 1)); out->chrom = (int32_t*)malloc(4 * (size_t)(n + 1));
    out->pos = (int32_t*)malloc(4 * (size_t)(n + 1)); out->tags = (uint8_t*)malloc((size_t)(n + 1));
    out->nm = (int32_t*)malloc(4 * (size_t)(n + 1));
    out->seq_off = (int64_t*)calloc((size_t)(n + 1), 8); out->cig_off = (int64_t*)calloc((size_t)(n + 1), 8);
    out->md_off = (int64_t*)calloc((size_t)(n + 1), 8);
    int nchunks = 256; if (nchunks > n) nchunks = (int)(n > 0 ? n : 1);
    tbuf* tb = (tbuf*)calloc((size_t)nchunks, sizeof(tbuf));
    int64_t per = (n + nchunks - 1) / nchunks;
#pragma omp parallel for schedule(dynamic, 1)
    for (int c = 0; c < nchunks; c++) {
        int64_t lo = c * per, hi = lo + per < n ? lo + per : n;
        for (int64_t i = lo; i < hi; i++) {
            int64_t s0 = tb[c].seq.n, c0 = tb[c].cop.n, m0 = tb[c].md.n, nops;
            gen_read(P, genome, gm, i, &tb[c], &out->flag[i], &out->chrom[i], &out->pos[i], &out->tags[i], &out->nm[i], &nops);
            out->seq_off[i + 1] = tb[c].seq.n - s0; out->cig_off[i + 1] = tb[c].cop.n - c0; out->md_off[i + 1] = tb[c].md.n - m0;
        }
    }
    for (int64_t i = 0; i < n; i++) { out->seq_off[i + 1] += out->seq_off[i]; out->cig_off[i + 1] += out->cig_off[i]; out->md_off[i + 1] += out->md_off[i]; }
    out->seq = (uint8_t*)malloc((size_t)out->seq_off[n] + 8); out->cig_op = (uint8_t*)malloc((size_t)out->cig_off[n] + 8);
    out->cig_len = (int32_t*)malloc(4 * (size_t)out->cig_off[n] + 8); out->md = (uint8_t*)malloc((size_t)out->md_off[n] + 8);
#pragma omp parallel for schedule(dynamic, 1)
    for (int c = 0; c < nchunks; c++) {
        int64_t lo = c * per; if (lo >= n) continue;
        memcpy(out->seq + out->seq_off[lo], tb[c].seq.p, (size_t)tb[c].seq.n);
        memcpy(out->cig_op + out->cig_off[lo], tb[c].cop.p, (size_t)tb[c].cop.n);
        memcpy(out->cig_len + out->cig_off[lo], tb[c].clen.p, (size_t)tb[c].clen.n * 4);
        memcpy(out->md + out->md_off[lo], tb[c].md.p, (size_t)tb[c].md.n);
    }
    for (int c = 0; c < nchunks; c++) { free(tb[c].seq.p); free(tb[c].cop.p); free(tb[c].clen.p); free(tb[c].md.p); }
    free(tb);
    return 0;
}

void tcsynth_free_reads(read_cols* o) {
    free(o->flag); free(o->chrom); free(o->pos); free(o->tags); free(o->nm); free(o->seq_off); free(o->seq);
    free(o->cig_off); free(o->cig_op); free(o->cig_len); free(o->md_off); free(o->md);
    memset(o, 0, sizeof(*o));
}
void tcsynth_free_genes(gene_models* g) {
    free(g->gene_exon_off); free(g->gene_strand); free(g->gene_chrom); free(g->exon_start); free(g->exon_end);
    memset(g, 0, sizeof(*g));
}

/* SAM text of reads [lo, hi) of a columnar batch, as tools/workload.py sam_lines() formats them (bench tooling: a 1 M-read file in
 * seconds).  Returns the number of bytes written, -1 on an I/O error. */
#include <stdio.h>
int64_t tcsynth_write_sam(const char* path, const char* mode, const char* chrom_name, int64_t lo, int64_t hi, const int32_t* flag,
                          const int32_t* chrom, const int32_t* pos, const uint8_t* tags, const int64_t* seq_off, const uint8_t* seq,
                          const int64_t* cig_off, const uint8_t* cig_op, const int32_t* cig_len, const int64_t* md_off,
                          const uint8_t* md, const int32_t* nm) {
    FILE* f = fopen(path, mode);
    if (!f) return -1;
    static char buf[1 << 22];
    setvbuf(f, buf, _IOFBF, sizeof(buf));
    int64_t total = 0;
    for (int64_t i = lo; i < hi; i++) {
        int n;
        if (chrom[i] < 0 || flag[i] == 4) { n = fprintf(f, "r%lld\t%d\t*\t0\t0\t*\t*\t0\t0\t*\t*\n", (long long)i, flag[i]); total += n; continue; }
        if (flag[i] > 16) { n = fprintf(f, "r%lld\t%d\t%s\t1\t0\t10M\t*\t0\t0\t*\t*\n", (long long)i, flag[i], chrom_name); total += n; continue; }
        total += fprintf(f, "r%lld\t%d\t%s\t%d\t60\t", (long long)i, flag[i], chrom_name, pos[i]);
        for (int64_t k = cig_off[i]; k < cig_off[i + 1]; k++) total += fprintf(f, "%d%c", cig_len[k], (char)cig_op[k]);
        total += fprintf(f, "\t*\t0\t0\t");
        total += (int64_t)fwrite(seq + seq_off[i], 1, (size_t)(seq_off[i + 1] - seq_off[i]), f);
        total += fprintf(f, "\t*");
        if ((tags[i] & 3) == 3) {
            total += fprintf(f, "\tNM:i:%d\tMD:Z:", nm[i]);
            total += (int64_t)fwrite(md + md_off[i], 1, (size_t)(md_off[i + 1] - md_off[i]), f);
        }
        total += fprintf(f, "\n");
    }
    if (fclose(f) != 0) return -1;
    return total;
}

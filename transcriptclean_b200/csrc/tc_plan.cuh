// tc_plan.cuh - per-read planning logic of the correction path (one thread per read).
//
// These functions decide *what* happens to a read - which mismatches / indels are rewritten,
// which junctions move, the new CIGAR, the TE.log records, the counters - and describe the new
// SEQ as a short list of segments (slices of the input SEQ or of the genome).  No bulk bytes
// move here; tc_emit (warp per read) materialises SEQ and regenerates NM/MD afterwards.
//
// Semantics follow the reference line by line (citations: TC = TranscriptClean.py,
// tr = transcript.py, sj = spliceJunction.py, ib = intronBound.py; Qn = SURVEY.md 3.4).
#pragma once
#include <stdint.h>
#include "../../include/tc_b200.h"

#ifdef TC_HOSTSIM
#define TC_HD inline
#else
#define TC_HD __device__ inline
#endif

typedef unsigned long long u64;

// ---------------------------------------------------------------------------------- tables

struct GenomeDev {
    const uint32_t* g2;        // 2-bit bases, 16 per word
    const uint32_t* lo1;       // lower-case plane, 32 per word
    const uint32_t* excblk;    // 1 bit per 256-base block that holds a non-ACGT base
    const int64_t* chrom_off;
    const int64_t* chrom_len;
    const int64_t* exc_start;  // literal runs: [start, end) in plane coordinates
    const int64_t* exc_end;
    const uint8_t* exc_byte;
    int64_t n_exc;
    int32_t n_chrom;
};

struct SjDev {
    const u64 *don_s, *acc_s, *don_u, *acc_u, *ann_k1, *ann_k2;
    int64_t n_don_s, n_acc_s, n_don_u, n_acc_u, n_ann;
    const int32_t *rng_don_s, *rng_acc_s, *rng_don_u, *rng_acc_u;   // first row of each key prefix (n_prefix + 1 entries)
    int32_t n_prefix;
    const struct AnnSlot* ann_hash; uint32_t ann_mask;              // open-addressing table over (k1, k2); 0 slots = use the sorted arrays
};
struct alignas(16) AnnSlot { u64 k1, k2; };                         // empty slot: k1 == ~0
#ifdef TC_HOSTSIM
inline
#else
__host__ __device__ inline
#endif
uint32_t ann_hash_of(u64 k1, u64 k2) {
    u64 h = (k1 ^ (k2 * 0x9E3779B97F4A7C15ull)) * 0xD6E8FEB86659FD93ull;
    return (uint32_t)(h >> 32);
}

struct VarDev {
    const u64* snp_key; const uint8_t* snp_alt; int64_t n_snp;
    const u64* ins_key; const int32_t* ins_size; const int64_t* ins_aoff; const uint8_t* ins_bytes; int64_t n_ins;
    const u64* del_key; const int32_t* del_size; int64_t n_del;
};

struct BatchDev {
    int64_t n;
    const int32_t *flag, *chrom, *pos;
    const uint8_t* tags;
    const int64_t* seq_off; const uint8_t* seq;
    const int64_t* cig_off; const uint8_t* cig_op; const int32_t* cig_len;
    const int64_t* md_off; const uint8_t* md;
    const int64_t* ji_off; const int32_t* ji;
};

// everything tc_emit needs to know about a read: one 128-byte row, one word per lane.
// Words 0-15 are written by sizes_read, 16-31 by pack_read once the output offsets are scanned.
struct alignas(16) EmitPlan { int32_t w[32]; };
enum { EP_FLAGS, EP_NCIG, EP_NSEG, EP_NTE, EP_TEMM, EP_TEINS, EP_TEDEL, EP_NJN, EP_OUTLEN, EP_NMEXTRA, EP_CIGCAP, EP_DLEN, EP_ODL_LO,
       EP_STATUS, EP_CHROM, EP_POS,
       EP_WS_LO = 16, EP_WS_HI, EP_OSEQ_LO, EP_OSEQ_HI, EP_OCIG_LO, EP_OCIG_HI, EP_OTE_LO, EP_OTE_HI, EP_OJN_LO, EP_OJN_HI,
       EP_INB_LO, EP_INB_HI, EP_LQ, EP_CIG0_LO, EP_CIG0_HI, EP_ODL_HI };
// EP_FLAGS: bit0 refresh, bits 2-3 cig_buf+1, bits 4-5 seg_buf+1.  EP_DLEN: bytes of the SEQ delta; EP_ODL: its place in the delta output.
// The other two list capacities follow from EP_CIGCAP and EP_NJN (measure_read): segments cc + 2 nj + 2, TE rows cc - nj - 1.
#define EP_SEGCAP_OF(cc, nj) ((cc) + 2 * (nj) + 2)
#define EP_TECAP_OF(cc, nj) ((cc) - (nj) - 1)

// per-read state (structure of arrays) + workspace
struct WorkDev {
    uint32_t* status;
    int32_t* counters;       // 10 per read
    // NM/MD computed at init for reads that lack either tag (tr.py:47-48)
    const uint8_t* md_init_pool; int64_t* md_init_off; int32_t* md_init_len; int32_t* nm_init;
    // measure
    int32_t* n_mdtok; int32_t* n_jn; uint8_t* mflags;
    uint32_t* md_tok;        // input MD tags split into tokens (op << 30 | count), same CSR offsets as B.md
    int32_t *cig_cap, *seg_cap, *te_cap;
    int64_t* ws_off;         // n+1 (exclusive scan of the need)
    u64* ws;
    // current planned state
    int8_t* cig_buf; int32_t* n_cig_cur;    // -1 = the input CIGAR
    int8_t* seg_buf; int32_t* n_seg_cur;    // -1 = the input SEQ
    int32_t* n_te; int32_t* te_cnt;         // te_cnt: 3 per read (mismatch, insertion, deletion rows)
    int32_t* nm_extra; uint8_t* refresh;
    // exact output sizes, turned into offsets (n+1) by exclusive scans
    int64_t *o_seq, *o_cig, *o_te, *o_jn, *o_delta;
    int32_t* delta_size;     // bytes of the read's SEQ delta (kept current by plan_read / the junction rescue, like seq_len_out)
    EmitPlan* plan;
    int32_t* seq_len_out;    // exact length of the new SEQ (o_seq offsets are padded to 16 bytes)
};

enum { MF_HAS_I = 1, MF_HAS_D = 2, MF_HAS_N = 4, MF_MD_ACGTN = 8, MF_MD_BYTES = 16, MF_JI_DONE = 32 };
// MF_MD_BYTES: a count does not fit a token; MF_JI_DONE: plan_read already left the intron coordinates (compute_jI) in the junction records
enum { TAG_NM = 1, TAG_MD = 2, TAG_JM = 4, TAG_JI = 8 };

struct Params {
    GenomeDev G; SjDev S; VarDev V; BatchDev B; WorkDev W; tc_options O;
    int32_t sj_enabled;      // len(refs.sjAnnot) > 0 (TC.py:443)
};

// ---------------------------------------------------------------------------------- packing

#define SEG_LEN_BITS 24
#define SEG_LEN_MASK ((1ull << SEG_LEN_BITS) - 1)
TC_HD u64 seg_make(int kind, int64_t src, int64_t len) { return ((u64)kind << 63) | ((u64)src << SEG_LEN_BITS) | (u64)len; }
TC_HD int seg_kind(u64 s) { return (int)(s >> 63); }
TC_HD int64_t seg_src(u64 s) { return (int64_t)((s & ~(1ull << 63)) >> SEG_LEN_BITS); }
TC_HD int64_t seg_len(u64 s) { return (int64_t)(s & SEG_LEN_MASK); }
TC_HD u64 cg_make(int op, int32_t len) { return ((u64)(uint32_t)op << 32) | (uint32_t)len; }
TC_HD int cg_op(u64 c) { return (int)(c >> 32); }
TC_HD int32_t cg_len(u64 c) { return (int32_t)(uint32_t)c; }

// .TE.log row as it crosses the ABI (include/tc_b200.h, tc_te_decode): a | size << 32 | type << 56 | status << 58 | reason << 59;
// a junction row (type 3) is followed by one more word that holds b.  (Insertion / Deletion: b = a + size.)
TC_HD u64 te_word(int32_t a, int32_t size, int type, int status, int reason) {
    return (u64)(uint32_t)a | ((u64)((uint32_t)size & 0xffffffu) << 32) | ((u64)(type & 3) << 56) | ((u64)(status & 1) << 58) | ((u64)(reason & 7) << 59);
}
// New SEQ as an edit script against the input SEQ (include/tc_b200.h, "SEQ delta"): ops COPY n / SKIP n / LIT n + n bytes;
// header byte = kind << 6 | n6, n6 < 62: n = n6; 62: n in the next 2 bytes; 63: n in the next 4 bytes (little endian).
enum { DL_COPY = 0, DL_SKIP = 1, DL_LIT = 2 };
TC_HD int dl_hdr_len(int64_t n) { return n < 62 ? 1 : n < 65536 ? 3 : 5; }
TC_HD int dl_put_hdr(uint8_t* p, int kind, int64_t n) {
    if (n < 62) { p[0] = (uint8_t)((kind << 6) | (int)n); return 1; }
    if (n < 65536) { p[0] = (uint8_t)((kind << 6) | 62); p[1] = (uint8_t)(n & 255); p[2] = (uint8_t)(n >> 8); return 3; }
    p[0] = (uint8_t)((kind << 6) | 63); p[1] = (uint8_t)(n & 255); p[2] = (uint8_t)((n >> 8) & 255); p[3] = (uint8_t)((n >> 16) & 255); p[4] = (uint8_t)((n >> 24) & 255);
    return 5;
}

// bytes one segment adds to the delta; `in_end` = end of the latest input slice before it (updated)
TC_HD int dl_seg_size(u64 sg, int64_t& in_end) {
    const int64_t len = seg_len(sg);
    if (len <= 0) return 0;
    if (seg_kind(sg)) return dl_hdr_len(len) + (int)len;
    const int64_t src = seg_src(sg), gap = src - in_end;
    in_end = src + len;
    return (gap > 0 ? dl_hdr_len(gap) : 0) + dl_hdr_len(len);
}
TC_HD int64_t dl_list_size(const u64* segs, int ns) {
    int64_t n = 0, in_end = 0;
    for (int s = 0; s < ns; s++) n += dl_seg_size(segs[s], in_end);
    return n;
}

struct WsView { u64* cig[3]; u64* seg[3]; tc_te* te; int32_t* jn; };
// junction record, 8 int32: b0 b1 start end code snap_ji0 snap_ji1 snap_jm
#define JN_STRIDE 8

TC_HD int64_t ws_need_units(int cig_cap, int seg_cap, int te_cap, int n_jn) {
    return 3ll * cig_cap + 3ll * seg_cap + 2ll * te_cap + 4ll * n_jn;
}
TC_HD WsView ws_view(const WorkDev& W, int64_t i) {
    WsView v; u64* p = W.ws + W.ws_off[i];
    int cc = W.cig_cap[i], sc = W.seg_cap[i], tc = W.te_cap[i];
    for (int k = 0; k < 3; k++) { v.seg[k] = p; p += sc; }      // segment lists first: tc_emit prefetches them
    for (int k = 0; k < 3; k++) { v.cig[k] = p; p += cc; }
    v.te = (tc_te*)p; p += 2ll * tc;
    v.jn = (int32_t*)p;
    return v;
}

// ---------------------------------------------------------------------------------- genome

TC_HD uint8_t genome_byte_fast(const GenomeDev& G, int64_t g) {
    uint32_t code = (G.g2[g >> 4] >> ((g & 15) * 2)) & 3u;
    uint32_t low = (G.lo1[g >> 5] >> (g & 31)) & 1u;
    return (uint8_t)((0x54474341u >> (code * 8)) & 0xff) | (uint8_t)(low << 5);   // "ACGT"
}
TC_HD bool genome_blk_has_exc(const GenomeDev& G, int64_t g) {
    int64_t b = g >> 8;
    return (G.excblk[b >> 5] >> (b & 31)) & 1u;
}
TC_HD uint8_t genome_byte_exc(const GenomeDev& G, int64_t g) {   // g lies in a flagged block
    int64_t lo = 0, hi = G.n_exc;
    while (lo < hi) { int64_t m = (lo + hi) >> 1; if (G.exc_start[m] <= g) lo = m + 1; else hi = m; }
    if (lo > 0 && g < G.exc_end[lo - 1]) return G.exc_byte[lo - 1];
    return genome_byte_fast(G, g);
}
TC_HD uint8_t genome_byte(const GenomeDev& G, int64_t g) {
    return genome_blk_has_exc(G, g) ? genome_byte_exc(G, g) : genome_byte_fast(G, g);
}
TC_HD uint8_t up8(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }

// pyfaidx get_seq(chrom, s, e): number of bases actually returned (truncated at the contig end)
TC_HD int fetch_len(const GenomeDev& G, int chrom, int64_t s, int64_t e) {
    int64_t L = G.chrom_len[chrom];
    if (e > L) e = L;
    return e >= s ? (int)(e - s + 1) : 0;
}

// ---------------------------------------------------------------------------------- searches

TC_HD int64_t lower_bound_u64(const u64* a, int64_t lo, int64_t hi, u64 key) {
    while (lo < hi) { int64_t m = (lo + hi) >> 1; if (a[m] < key) lo = m + 1; else hi = m; }
    return lo;
}

// nearest annotated site (pyranges.nearest as used at TC.py:1241-1259).  false = nothing on this
// chromosome/strand -> the reference's .iloc[0] raises (Q23).
TC_HD bool sj_nearest(const u64* tab, const int32_t* rng, int32_t n_prefix, u64 prefix, int64_t q, bool tie_right, int& dist) {
    if (prefix >= (u64)n_prefix) return false;
    const int64_t lo = rng[prefix], hi = rng[prefix + 1];
    if (lo == hi) return false;
    int64_t i = lower_bound_u64(tab, lo, hi, (prefix << 32) | (u64)(uint32_t)q);
    int64_t best;
    if (i < hi) {
        best = (int64_t)(tab[i] & 0xffffffffull);
        if (i > lo) {
            int64_t left = (int64_t)(tab[i - 1] & 0xffffffffull);
            int64_t dl = q - left, dr = best - q;
            if (dl < dr || (dl == dr && !tie_right)) best = left;
        }
    } else {
        best = (int64_t)(tab[i - 1] & 0xffffffffull);
    }
    dist = (int)(best - q);
    return true;
}

TC_HD bool sj_annotated(const SjDev& S, int chrom, int32_t start, int32_t end, int strand) {
    const u64 k1 = ((u64)chrom << 32) | (uint32_t)start, k2 = ((u64)(uint32_t)end << 1) | (u64)strand;
    if (S.ann_hash) {                                    // one or two probes instead of ~15 dependent search steps
        for (uint32_t h = ann_hash_of(k1, k2) & S.ann_mask;; h = (h + 1) & S.ann_mask) {
#ifdef TC_HOSTSIM
            const AnnSlot e = S.ann_hash[h];
#else
            const ulonglong2 raw = *reinterpret_cast<const ulonglong2*>(S.ann_hash + h); AnnSlot e; e.k1 = raw.x; e.k2 = raw.y;
#endif
            if (e.k1 == k1 && e.k2 == k2) return true;
            if (e.k1 == ~0ull) return false;
        }
    }
    int64_t lo = 0, hi = S.n_ann;                        // one lexicographic search over (k1, k2)
    while (lo < hi) {
        const int64_t m = (lo + hi) >> 1; const u64 a = S.ann_k1[m];
        if (a < k1 || (a == k1 && S.ann_k2[m] < k2)) lo = m + 1; else hi = m;
    }
    return lo < S.n_ann && S.ann_k1[lo] == k1 && S.ann_k2[lo] == k2;
}

TC_HD bool snp_match(const VarDev& V, int chrom, int64_t pos, uint8_t base) {   // TC.py:1114-1118
    u64 key = ((u64)chrom << 32) | (uint32_t)pos;
    for (int64_t j = lower_bound_u64(V.snp_key, 0, V.n_snp, key); j < V.n_snp && V.snp_key[j] == key; j++)
        if (V.snp_alt[j] == base) return true;
    return false;
}
TC_HD int64_t indel_lower(const u64* key, const int32_t* size, int64_t n, u64 k, int32_t sz) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t m = (lo + hi) >> 1;
        if (key[m] < k || (key[m] == k && size[m] < sz)) lo = m + 1; else hi = m;
    }
    return lo;
}
TC_HD bool ins_match(const VarDev& V, int chrom, int64_t pos, int32_t ct, const uint8_t* s) {   // TC.py:905-909
    u64 key = ((u64)chrom << 32) | (uint32_t)pos;
    for (int64_t j = indel_lower(V.ins_key, V.ins_size, V.n_ins, key, ct);
         j < V.n_ins && V.ins_key[j] == key && V.ins_size[j] == ct; j++) {
        int64_t a = V.ins_aoff[j], b = V.ins_aoff[j + 1];
        if (b - a != ct) continue;
        bool eq = true;
        for (int t = 0; t < ct; t++) if (V.ins_bytes[a + t] != s[t]) { eq = false; break; }
        if (eq) return true;
    }
    return false;
}
TC_HD bool del_match(const VarDev& V, int chrom, int64_t pos, int32_t ct) {   // TC.py:1012
    u64 key = ((u64)chrom << 32) | (uint32_t)pos;
    int64_t j = indel_lower(V.del_key, V.del_size, V.n_del, key, ct);
    return j < V.n_del && V.del_key[j] == key && V.del_size[j] == ct;
}

// ---------------------------------------------------------------------------------- MD tokens

struct MdCur { const uint8_t* p; int len; int i; const uint32_t* tok; };   // tok != 0: pre-split tokens, len of them
TC_HD bool is_dig(uint8_t c) { return c >= '0' && c <= '9'; }
#define MD_TOK_MAX 0x3fffffff
// splitMD (tr.py:128-157): runs of digits -> ('M', value); other runs -> '^...' = ('D', len-1),
// else ('X', len).  One token starting at byte s of the tag; returns the byte after it.
TC_HD int md_token_at(const uint8_t* p, int len, int s, int& op, int64_t& cnt) {
    int k = s; const uint8_t c = p[s];
    if (is_dig(c)) {
        int64_t v = 0;
        while (k < len && is_dig(p[k])) { v = v * 10 + (p[k] - '0'); k++; }
        op = 'M'; cnt = v;
    } else {
        while (k < len && !is_dig(p[k])) k++;
        const int L = k - s;
        if (c == '^') { op = 'D'; cnt = L - 1; } else { op = 'X'; cnt = L; }
    }
    return k;
}
TC_HD uint32_t md_tok_pack(int op, int64_t cnt) { return ((op == 'M' ? 0u : op == 'X' ? 1u : 2u) << 30) | (uint32_t)cnt; }
TC_HD bool md_next(MdCur& t, int& op, int32_t& cnt) {
    if (t.i >= t.len) return false;
    if (t.tok) {
        const uint32_t w = t.tok[t.i++]; const uint32_t k = w >> 30;
        op = k == 0 ? 'M' : k == 1 ? 'X' : 'D'; cnt = (int32_t)(w & MD_TOK_MAX);
        return true;
    }
    int64_t v; t.i = md_token_at(t.p, t.len, t.i, op, v); cnt = (int32_t)v;
    return true;
}

// splitMD over a whole tag, reading it in aligned 8-byte words (the buffers that hold MD tags are
// 8-byte aligned and padded, so the words around [p, p+len) are readable).  Stores the tokens when
// `tok` is given; returns their number; sets MF_MD_ACGTN (TC.py:1090 looks for a base letter in the
// tag) and MF_MD_BYTES (a count that does not fit a token).
TC_HD int md_tokenize(const uint8_t* p, int len, uint32_t* tok, uint8_t& mf) {
    // One straight-line step per byte (the lanes of a warp are in different states, so branches on the
    // state would all be walked): a token ends where the digit / non-digit class changes; the ended
    // token's word is `cur` for a number and kind << 30 | count for a letter run.
    int ntok = 0; uint32_t v = 0, L = 0, kind = 0; bool prev_dig = false, any = false, acgtn = false, big = false;   // v > MD_TOK_MAX: too big (sticky)
    const int a = (int)((uintptr_t)p & 7);
    const u64* wp = reinterpret_cast<const u64*>(p - a);
    const int end = a + len;                                          // bytes [a, end) of the word stream
    for (int w0 = 0; w0 < end; w0 += 8) {
        const u64 word = wp[w0 >> 3];
        const uint32_t half[2] = {(uint32_t)word, (uint32_t)(word >> 32)};
#pragma unroll
        for (int k = 0; k < 8; k++) {                                 // constant shifts; bytes outside [a, end) are skipped
            if (w0 + k < a || w0 + k >= end) continue;
            const uint32_t c = (half[k >> 2] >> (8 * (k & 3))) & 255u, d = c - '0';
            const bool dig = d < 10u;
            if (any && dig != prev_dig) {                             // the previous token ends here
                const uint32_t w = prev_dig ? v : (kind << 30) | (L - (kind == 2u ? 1u : 0u));
                if (prev_dig && v > MD_TOK_MAX) big = true; else if (tok) tok[ntok] = w;
                ntok++;
            }
            if (!any || dig != prev_dig) { v = 0; L = 0; kind = c == '^' ? 2u : 1u; }
            v = !dig ? v : (v > MD_TOK_MAX / 10 ? 0xffffffffu : v * 10u + d);
            L += dig ? 0u : 1u;
            const uint32_t u = c & 0xDFu;
            acgtn |= !dig && (u == 'A' || u == 'C' || u == 'G' || u == 'T' || u == 'N');
            prev_dig = dig; any = true;
        }
    }
    if (any) {
        const uint32_t w = prev_dig ? v : (kind << 30) | (L - (kind == 2u ? 1u : 0u));
        if (prev_dig && v > MD_TOK_MAX) big = true; else if (tok) tok[ntok] = w;
        ntok++;
    }
    if (acgtn) mf |= MF_MD_ACGTN;
    if (big) mf |= MF_MD_BYTES;
    return ntok;
}

// `B` may be a copy of P.B whose cig_op / cig_len / md pointers address a staged (shared-memory) copy
TC_HD void md_effective(const Params& P, const BatchDev& B, int64_t i, const uint8_t*& p, int& len) {
    if (P.W.status[i] & TC_ST_NMMD_DEVICE) { p = P.W.md_init_pool + P.W.md_init_off[i]; len = P.W.md_init_len[i]; }
    else { p = B.md + B.md_off[i]; len = (int)(B.md_off[i + 1] - B.md_off[i]); }
}
// cursor over the read's MD tokens: the pre-split tokens of an input tag when the measure pass stored them
TC_HD void md_cursor(const Params& P, const BatchDev& B, int64_t i, MdCur& t) {
    t.i = 0; t.tok = 0;
    md_effective(P, B, i, t.p, t.len);
    if (P.W.md_tok && !(P.W.status[i] & TC_ST_NMMD_DEVICE) && !(P.W.mflags[i] & MF_MD_BYTES)) { t.tok = P.W.md_tok + B.md_off[i]; t.len = P.W.n_mdtok[i]; }
}

TC_HD int read_class(int32_t flag, int32_t chrom) {           // TC.py:1570-1575 (Q1)
    if (chrom < 0 || flag == 4) return 1;
    if (flag > 16) return 2;
    return 0;
}

// ---------------------------------------------------------------------------------- K0 measure

// Sizes the per-read workspace.  Runs after the init NM/MD pass so that reads without tags are
// measured on their generated MD.
TC_HD void measure_read(const Params& P, int64_t i) {
    const BatchDev& B = P.B; const WorkDev& W = P.W;
    uint32_t st = W.status[i];
    // (the ten counters of every read start at -1 = "NA": one memset of the whole array by the caller)
    W.cig_buf[i] = -1; W.seg_buf[i] = -1; W.n_cig_cur[i] = 0; W.n_seg_cur[i] = 0;
    W.n_te[i] = 0; W.te_cnt[3 * i] = W.te_cnt[3 * i + 1] = W.te_cnt[3 * i + 2] = 0;
    W.nm_extra[i] = 0; W.refresh[i] = 0; W.delta_size[i] = 0;
    W.n_mdtok[i] = 0; W.n_jn[i] = 0; W.mflags[i] = 0;
    W.cig_cap[i] = W.seg_cap[i] = W.te_cap[i] = 0;
    if ((st & TC_ST_CLASS_MASK) != 0) { W.ws_off[i] = 0; return; }
    int64_t c0 = B.cig_off[i], c1 = B.cig_off[i + 1];
    int nc = (int)(c1 - c0), nN = 0; uint8_t mf = 0;
    {   // op bytes in aligned 8-byte words (the op array is 8-byte aligned and padded like the MD buffer)
        const uint8_t* p = B.cig_op + c0; const int a = (int)((uintptr_t)p & 7), end = a + nc;
        const u64* wp = reinterpret_cast<const u64*>(p - a);
        for (int w0 = 0; w0 < end; w0 += 8) {
            const u64 word = wp[w0 >> 3];
            const int k0 = w0 < a ? a - w0 : 0, k1 = end - w0 < 8 ? end - w0 : 8;
            for (int k = k0; k < k1; k++) {
                const uint32_t op = (uint32_t)(word >> (8 * k)) & 255u;
                if (op == 'I') mf |= MF_HAS_I; else if (op == 'D') mf |= MF_HAS_D; else if (op == 'N') { mf |= MF_HAS_N; nN++; }
            }
        }
    }
    int ntok = 0;
    if (!(st & TC_ST_NMMD_NONE)) {
        const uint8_t* p; int len; md_effective(P, B, i, p, len);
        uint32_t* tok = (W.md_tok && !(st & TC_ST_NMMD_DEVICE)) ? W.md_tok + B.md_off[i] : 0;
        ntok = md_tokenize(p, len, tok, mf);
    }
    int nj = 0;
    if (mf & MF_HAS_N) nj = (B.tags[i] & TAG_JI) ? (int)((B.ji_off[i + 1] - B.ji_off[i]) / 2) : nN;   // tr.py:56-62
    W.n_mdtok[i] = ntok; W.n_jn[i] = nj; W.mflags[i] = mf;
    int cc = nc + ntok + 2 * nj + 2, sc = nc + ntok + 4 * nj + 4, tc = nc + ntok + nj + 1;
    W.cig_cap[i] = cc; W.seg_cap[i] = sc; W.te_cap[i] = tc;
    W.ws_off[i] = ws_need_units(cc, sc, tc, nj);      // need; scanned into offsets afterwards
}

// ---------------------------------------------------------------------------------- K1 plan

struct Emit {                 // newCIGAR / newSeq / MVal bookkeeping (endMatch, TC.py:1171-1183)
    u64* cig; int nc; u64* seg; int ns; int32_t run; int64_t seqlen; u64 last;   // last = seg[ns - 1]
    TC_HD void match(int32_t n) { run += n; }
    TC_HD void flush() { if (run > 0) { cig[nc++] = cg_make('M', run); run = 0; } }
    TC_HD void op(int o, int32_t n) { flush(); cig[nc++] = cg_make(o, n); }
    TC_HD void push(int kind, int64_t src, int64_t len) {
        if (len <= 0) return;
        seqlen += len;
        if (ns > 0 && seg_kind(last) == kind && seg_src(last) + seg_len(last) == src && seg_len(last) + len <= (int64_t)SEG_LEN_MASK) {
            last = seg_make(kind, seg_src(last), seg_len(last) + len); seg[ns - 1] = last; return;
        }
        last = seg_make(kind, src, len); seg[ns++] = last;
    }
};

TC_HD void te_put(tc_te* te, int& n, int type, int32_t a, int32_t b, int32_t size, int status, int reason) {
    // The TE area of a read starts 3*(cig_cap + seg_cap) units into its workspace; cig_cap + seg_cap and every
    // read's workspace size are even numbers of 8-byte units (measure_read), so a row is 16-byte aligned.
    const u64 w0 = (u64)(uint32_t)a | ((u64)(uint32_t)b << 32);
    const u64 w1 = (u64)(uint32_t)size | ((u64)(uint32_t)(type & 255) << 32) | ((u64)(uint32_t)(status & 255) << 40) | ((u64)(uint32_t)(reason & 255) << 48);
#ifdef TC_HOSTSIM
    u64* p = reinterpret_cast<u64*>(te + n); p[0] = w0; p[1] = w1;
#else
    *reinterpret_cast<ulonglong2*>(te + n) = make_ulonglong2(w0, w1);
#endif
    n++;
}

// Stages 1-3 in one left-to-right pass.  The reference runs correctMismatches (TC.py:1077-1168),
// correctInsertions (859-964) and correctDeletions (967-1074) as three passes; because each only
// rewrites its own kind of op, keeps the genome cursor of every op, and joins M runs the same
// way, one pass over the MD x CIGAR merge (tr.py:159-213) yields the same CIGAR, SEQ, counters
// and TE rows (rows are re-ordered by stage in tc_emit).
TC_HD void plan_read(const Params& P, int64_t i, const BatchDev& B) {
    const WorkDev& W = P.W; const tc_options& O = P.O;
    uint32_t st = W.status[i];
    if ((st & TC_ST_CLASS_MASK) != 0) return;
    int32_t* C = W.counters + 10 * i;
    const uint8_t mf = W.mflags[i];
    if (O.correct_mismatches) {
        C[TC_C_COR_MM] = 0; C[TC_C_UNC_MM] = 0;                      // TC.py:1081-1082
        if (st & TC_ST_NMMD_NONE) { W.status[i] = st | TC_ST_FALLBACK; return; }   // None.upper() (TC.py:1090)
    }
    bool stage1 = O.correct_mismatches && (mf & MF_MD_ACGTN);
    bool indel = O.correct_indels != 0;
    bool rewritten = stage1 || (indel && (mf & (MF_HAS_I | MF_HAS_D)));
    if (!rewritten) {
        if (indel) for (int k = TC_C_COR_DEL; k <= TC_C_VAR_INS; k++) C[k] = 0;
        return;
    }
    WsView v = ws_view(W, i);
    Emit E; E.cig = v.cig[0]; E.nc = 0; E.seg = v.seg[0]; E.ns = 0; E.run = 0; E.seqlen = 0; E.last = 0;
    const int chrom = B.chrom[i];
    const int64_t goff = P.G.chrom_off[chrom] - 1;       // plane index of 1-based position p is goff + p
    const uint8_t* seq = B.seq + B.seq_off[i];
    int64_t q = 0, g = B.pos[i];
    int nte = 0; u64 cnt_cor = 0, cnt_var = 0, cnt_big = 0;   // 21-bit fields: mismatches, insertions, deletions
    int32_t ins_removed = 0; bool sawD = false;
    // intron coordinates for reads without a jI tag: computed on this walk when the CIGAR is long (k_jn_init would
    // otherwise walk its ~150 ops again; for short CIGARs its own walk is cheaper than a fourth store stream here)
    const bool want_ji = (mf & MF_HAS_N) && !(B.tags[i] & TAG_JI) && B.cig_off[i + 1] - B.cig_off[i] > 64; int kN = 0;
    // merge state (tr.py:171-211)
    // The next CIGAR op and the next MD token are fetched one step ahead, and a pre-split token
    // stays an undecoded word until the step that looks at it, so the loads have a whole step to land.
    int64_t ci = B.cig_off[i]; const int64_t cend = B.cig_off[i + 1];
    int cop = ci < cend ? B.cig_op[ci] : 0; int32_t crem = ci < cend ? B.cig_len[ci] : 0;
    MdCur md; md.i = 0; md.len = 0; md.p = 0; md.tok = 0;
    int mop = 0; int32_t mrem = 0; bool have_md = false; uint32_t mw = 0; bool mraw = false;
    if (stage1) {
        md_cursor(P, B, i, md);
        if (md.tok) { have_md = md.i < md.len; if (have_md) { mw = md.tok[md.i++]; mraw = true; } } else have_md = md_next(md, mop, mrem);
    }
    const bool have_cat = P.V.n_snp > 0 || P.V.n_ins > 0 || P.V.n_del > 0;
    const int32_t max_indel = O.max_len_indel;
    bool fail = false;
    // The 32 reads of a warp sit on different kinds of op at any moment, so every `if` in this loop would be walked by
    // the whole warp once per branch taken (the first version spent four fifths of its issue slots that way: 13 of 32
    // lanes active on average).  The step is therefore written as one straight block - selects and predicated stores -
    // with branches left only for the loop exits and the variant catalogue searches.
    while (true) {
        if (ci >= cend) { fail = stage1 && have_md; break; }          // MD left over: cigarOperation[cigarIndex] IndexError (Q14)
        const bool pass = !stage1 || cop == 'H' || cop == 'S' || cop == 'I' || cop == 'N';
        if (!pass && !have_md) { fail = true; break; }                // mdCount[mdIndex] IndexError
        if (mraw) { const uint32_t k = mw >> 30; mop = k == 0 ? 'M' : k == 1 ? 'X' : 'D'; mrem = (int32_t)(mw & MD_TOK_MAX); mraw = false; }
        // ---- one step of the merge (tr.py:171-211): the shorter of the two current entries is the piece; ties take the MD entry
        const bool take_c = pass || crem < mrem, tie = !pass && crem == mrem;
        const bool adv_c = take_c || tie, adv_m = !take_c;
        const int op = take_c ? cop : mop; const int32_t ct = take_c ? crem : mrem;
        { const int32_t m0 = mrem; mrem = (!pass && take_c) ? mrem - crem : mrem; crem = (adv_m && !tie) ? crem - m0 : crem; }
        if (adv_c) { ci++; if (ci < cend) { cop = B.cig_op[ci]; crem = B.cig_len[ci]; } }              // fetched one step ahead of their use
        if (adv_m) {
            if (md.tok) { have_md = md.i < md.len; if (have_md) { mw = md.tok[md.i++]; mraw = true; } }   // (stays an undecoded word until the next step)
            else have_md = md_next(md, mop, mrem);
        }
        // ---- the piece:
        //   M            SEQ slice, joins the match run
        //   X            TE row; SNP catalogue hit -> kept as is (TC.py:1115-1130, Q11), else genome bases verbatim (Q6)
        //   S            CIGAR op + SEQ slice
        //   I (indels)   TE row; too long / catalogue hit -> kept, else dropped (TC.py:898-940)
        //   D (indels)   TE row; too long / catalogue hit -> kept, else filled from the genome (TC.py:1007-1051)
        //   N, H         CIGAR op;   anything else vanishes (Q10)
        const bool isM = op == 'M', isX = op == 'X', isS = op == 'S', isI = op == 'I', isD = op == 'D', isNH = op == 'N' || op == 'H';
        const bool edit = isX || ((isI || isD) && indel);                  // gets a TE row
        const bool toolarge = edit && !isX && ct > max_indel;
        bool variant = false;
        if (have_cat && edit && !toolarge) {                               // SEQ is only touched when a catalogue exists
            if (isX) variant = P.V.n_snp > 0 && snp_match(P.V, chrom, g, seq[q]);
            else variant = isI ? (P.V.n_ins > 0 && ins_match(P.V, chrom, g - 1, ct, seq + q)) : (P.V.n_del > 0 && del_match(P.V, chrom, g - 1, ct));
        }
        const bool kept = variant || toolarge, corrected = edit && !kept;
        if (edit) te_put(v.te, nte, isX ? 0 : isI ? 1 : 2, (int32_t)(isX ? g : g - 1), isX ? 0 : (int32_t)(g + ct - 1), ct, kept ? 1 : 0, toolarge ? 2 : variant ? 1 : 0);
        {
            const u64 inc = edit ? 1ull << (isX ? 0 : isI ? 21 : 42) : 0ull;
            cnt_big += toolarge ? inc : 0ull; cnt_var += (variant && !toolarge) ? inc : 0ull; cnt_cor += corrected ? inc : 0ull;
        }
        sawD |= isD;
        ins_removed += (isI && corrected) ? ct : 0;
        const bool from_genome = (isX || isD) && corrected;
        {   // Emit.push: the slice joins the segment before it when it continues it
            const bool do_push = ct > 0 && (isM || isS || (isX && !corrected) || (isI && !corrected) || from_genome);
            const int kind = from_genome ? 1 : 0; const int64_t src = from_genome ? goff + g : q;
            const int64_t l_src = seg_src(E.last), l_len = seg_len(E.last); const int l_kind = seg_kind(E.last);
            const bool merge = do_push && E.ns > 0 && l_kind == kind && l_src + l_len == src && l_len + ct <= (int64_t)SEG_LEN_MASK;
            const bool fresh = do_push && !merge;
            E.seqlen += do_push ? ct : 0;
            const u64 nl = merge ? seg_make(kind, l_src, l_len + ct) : seg_make(kind, src, ct);
            if (do_push) { E.seg[merge ? E.ns - 1 : E.ns] = nl; E.last = nl; }
            E.ns += fresh ? 1 : 0;
        }
        {   // Emit.match / Emit.op (endMatch, TC.py:1171-1183)
            const bool join = isM || isX || (isD && corrected);
            const bool isop = !join && (isS || isNH || ((isI || isD) && !corrected));
            const bool fl = isop && E.run > 0;
            if (fl) E.cig[E.nc] = cg_make('M', E.run);
            E.nc += fl ? 1 : 0;
            if (isop) E.cig[E.nc] = cg_make(op, ct);
            E.nc += isop ? 1 : 0;
            E.run = isop ? 0 : E.run + (join ? ct : 0);
        }
        q += (isM || isX || isS || isI) ? ct : 0;
        if (want_ji && op == 'N') {                                   // compute_jI (tr.py:368-393) on the way; one 8-byte store
#ifdef TC_HOSTSIM
            v.jn[kN * JN_STRIDE] = (int32_t)g; v.jn[kN * JN_STRIDE + 1] = (int32_t)(g + ct - 1);
#else
            *reinterpret_cast<int2*>(v.jn + kN * JN_STRIDE) = make_int2((int32_t)g, (int32_t)(g + ct - 1));
#endif
            kN++;
        }
        g += (isM || isX || isD || isNH) ? ct : 0;
    }
    if (fail) {                                                       // whole-read fallback (Q13): stage 1 raised
        W.status[i] = st | TC_ST_FALLBACK;                            // before stages 2-3 zeroed their counters
        return;
    }
    E.flush();
    const int cM = (int)(cnt_cor & 0x1fffff), cI = (int)((cnt_cor >> 21) & 0x1fffff), cD = (int)((cnt_cor >> 42) & 0x1fffff);
    const int uM = (int)(cnt_var & 0x1fffff), vI = (int)((cnt_var >> 21) & 0x1fffff), vD = (int)((cnt_var >> 42) & 0x1fffff);
    const int uI = (int)((cnt_big >> 21) & 0x1fffff), uD = (int)((cnt_big >> 42) & 0x1fffff);
    const int nteM = cM + uM, nteI = cI + uI + vI, nteD = cD + uD + vD;
    C[TC_C_COR_MM] = O.correct_mismatches ? cM : -1; C[TC_C_UNC_MM] = O.correct_mismatches ? uM : -1;
    if (indel) { C[TC_C_COR_INS] = cI; C[TC_C_UNC_INS] = uI; C[TC_C_VAR_INS] = vI;
                 C[TC_C_COR_DEL] = cD; C[TC_C_UNC_DEL] = uD; C[TC_C_VAR_DEL] = vD; }
    W.cig_buf[i] = 0; W.n_cig_cur[i] = E.nc; W.seg_buf[i] = 0; W.n_seg_cur[i] = E.ns; W.seq_len_out[i] = (int32_t)E.seqlen;
    W.delta_size[i] = (int32_t)dl_list_size(E.seg, E.ns);              // bytes of the SEQ delta: one more walk over the finished list (cheaper than booking every step)
    W.n_te[i] = nte; W.te_cnt[3 * i] = nteM; W.te_cnt[3 * i + 1] = nteI; W.te_cnt[3 * i + 2] = nteD;
    // NM/MD refresh points (Q7): end of correctDeletions when the CIGAR holds a D, else end of
    // correctMismatches; insertion removal never refreshes, so NM keeps counting removed bases.
    if (indel && sawD) { W.refresh[i] = 1; W.nm_extra[i] = 0; }
    else if (stage1) { W.refresh[i] = 1; W.nm_extra[i] = ins_removed; }
    if (want_ji) W.mflags[i] = mf | MF_JI_DONE;
    W.status[i] = st | TC_ST_CIGAR_REWRITTEN;
}

TC_HD void plan_read(const Params& P, int64_t i) { plan_read(P, i, P.B); }

// ---------------------------------------------------------------------------------- K2 junctions

TC_HD int motif_code(uint8_t a0, uint8_t a1, uint8_t b0, uint8_t b1) {   // sj.py:71-89
    uint32_t m = ((uint32_t)up8(a0) << 24) | ((uint32_t)up8(a1) << 16) | ((uint32_t)up8(b0) << 8) | up8(b1);
    switch (m) {
    case 0x47544147u: return 1;   // GTAG
    case 0x43544143u: return 2;   // CTAC
    case 0x47434147u: return 3;   // GCAG
    case 0x43544743u: return 4;   // CTGC
    case 0x41544143u: return 5;   // ATAC
    case 0x47544154u: return 6;   // GTAT
    }
    return 0;
}

// checkSpliceMotif (sj.py:48-68).  false = pyfaidx would raise (start < 1).
TC_HD bool junction_code(const Params& P, int chrom, int strand, int32_t b0, int32_t b1, int32_t start,
                         int32_t end, int& code) {
    if (b0 < 1 || b1 - 1 < 1) return false;
    const GenomeDev& G = P.G; int64_t goff = G.chrom_off[chrom] - 1;
    int n0 = fetch_len(G, chrom, b0, (int64_t)b0 + 1), n1 = fetch_len(G, chrom, (int64_t)b1 - 1, b1);
    int mc = 0;
    if (n0 == 2 && n1 == 2) {
        const int64_t ga = goff + b0, gb = goff + b1 - 1;
        if (genome_blk_has_exc(G, ga) | genome_blk_has_exc(G, ga + 1) | genome_blk_has_exc(G, gb) | genome_blk_has_exc(G, gb + 1)) {
            mc = motif_code(genome_byte(G, ga), genome_byte(G, ga + 1), genome_byte(G, gb), genome_byte(G, gb + 1));
        } else {                                        // case does not matter for the motif: 2-bit plane only
            const uint32_t c = ((G.g2[ga >> 4] >> ((ga & 15) * 2)) & 3u) | (((G.g2[(ga + 1) >> 4] >> (((ga + 1) & 15) * 2)) & 3u) << 2) |
                               (((G.g2[gb >> 4] >> ((gb & 15) * 2)) & 3u) << 4) | (((G.g2[(gb + 1) >> 4] >> (((gb + 1) & 15) * 2)) & 3u) << 6);
            mc = c == 142u ? 1 : c == 77u ? 2 : c == 134u ? 3 : c == 109u ? 4 : c == 76u ? 5 : c == 206u ? 6 : 0;   // GTAG CTAC GCAG CTGC ATAC GTAT
        }
    }
    code = mc + (sj_annotated(P.S, chrom, start, end, strand) ? 20 : 0);
    return true;
}

// the motif part of checkSpliceMotif alone (sj.py:48-68, 71-89): -1 = pyfaidx would raise (a coordinate below 1)
TC_HD int intron_motif(const GenomeDev& G, int chrom, int32_t b0, int32_t b1) {
    if (b0 < 1 || b1 - 1 < 1) return -1;
    if (fetch_len(G, chrom, b0, (int64_t)b0 + 1) != 2 || fetch_len(G, chrom, (int64_t)b1 - 1, b1) != 2) return 0;
    const int64_t goff = G.chrom_off[chrom] - 1, ga = goff + b0, gb = goff + b1 - 1;
    return motif_code(genome_byte(G, ga), genome_byte(G, ga + 1), genome_byte(G, gb), genome_byte(G, gb + 1));
}

struct ExonLoc { int lo, hi; int64_t q0, q1; int nidx; };
// group_CIGAR_by_exon_and_intron (TC.py:1379-1412): exon t = ops between the t-th and (t+1)-th N;
// its sequence = the M/S/I bases in that range.
TC_HD bool locate_exon(const u64* cig, int n, int t, ExonLoc& e, int& nN) {
    int exon = 0, lo = 0; int64_t q = 0, q0 = 0; bool found = false;
    for (int k = 0; k < n; k++) {
        int op = cg_op(cig[k]);
        if (op == 'N') {
            if (exon == t) { e.lo = lo; e.hi = k; e.q0 = q0; e.q1 = q; e.nidx = k; found = true; }
            exon++; lo = k + 1; q0 = q;
        } else if (op == 'M' || op == 'S' || op == 'I') q += cg_len(cig[k]);
    }
    if (exon == t) { e.lo = lo; e.hi = n; e.q0 = q0; e.q1 = q; e.nidx = -1; found = true; }
    nN = exon;
    return found;
}
TC_HD int64_t cig_query_len(const u64* cig, int n) {             // tr.py:591-605
    int64_t q = 0;
    for (int k = 0; k < n; k++) { int op = cg_op(cig[k]); if (op == 'M' || op == 'S' || op == 'I') q += cg_len(cig[k]); }
    return q;
}

// ---- introns of length <= 0 after a rescue.
// fix_one_side_of_junction pastes the new CIGAR together as TEXT (TC.py:1466-1470) and every later consumer splits that
// text again with two regular expressions (tr.py:579-588, TC.py:1504-1512).  An intron whose length went negative prints
// as "-3N"; on the next split the '-' sticks to the op letter in front of it ("20M-" is no known op any more) while the
// count parses as -3.  What follows from that, in terms of the op list (probe-verified against the
// unmodified reference, tests/test_micro_introns.py):
//   length 0                   "0N": an ordinary op, everything downstream is integer arithmetic
//   < 0 behind M / S / I       the length check at TC.py:1476 no longer counts that op -> RuntimeError -> "Other"
//   < 0 behind D               the check passes ("D-" was never counted).  The op is kept as 'd' here: no consumer knows it
//                              (getNMandMDFlags skips it, tr.py:296-334), it prints as 'D', and the '-' it carries now
//                              belongs to the text: a later re-serialisation prints "D-" + str(v) + "N", which parses as
//                              -v for v >= 0 and raises ValueError ("--3") for v < 0
//   < 0 behind anything else   (an intron, a clip, the start of the CIGAR) the split lists go out of step; not modelled:
//                              the junction is left uncorrected ("Other") and TC_ST_ERR_INTRON is set on the read
#define CG_OP_DMINUS 'd'
// dst[idx] is the intron whose arithmetic length is L.  Returns false when the reference raises.
TC_HD bool intron_text_rules(u64* dst, int idx, int32_t L, bool& unsupported) {
    const int prev = idx > 0 ? cg_op(dst[idx - 1]) : 0;
    if (prev == CG_OP_DMINUS) {
        if (L < 0) return false;                                      // "5D--3N": int("--3")
        dst[idx] = cg_make('N', -L); return true;                     // "5D-2N" parses as -2
    }
    if (L >= 0) { dst[idx] = cg_make('N', L); return true; }
    if (prev == 'M' || prev == 'S' || prev == 'I') return false;      // the op's bases drop out of the length check
    if (prev == 'D') { dst[idx - 1] = cg_make(CG_OP_DMINUS, cg_len(dst[idx - 1])); dst[idx] = cg_make('N', L); return true; }
    unsupported = true; return false;
}
// every other "D-" of the list is printed again with its intron behind it: a negative one raises ("--")
TC_HD bool other_dminus_ok(const u64* c, int n, int edited_idx) {
    for (int k = 0; k + 1 < n; k++)
        if (cg_op(c[k]) == CG_OP_DMINUS && k + 1 != edited_idx && cg_op(c[k + 1]) == 'N' && cg_len(c[k + 1]) < 0) return false;
    return true;
}

// editExonCIGAR (TC.py:1515-1555) applied to ops [lo,hi) of src, plus the intron length change.
// Returns the new op count, or -1 for the IndexError on an empty exon (TC.py:1526); `raised` is set when the
// text rules above make the reference raise.
TC_HD int cig_edit(const u64* src, int n, u64* dst, int lo, int hi, bool at_end, int nb,
                   int intron_idx, int intron_delta, bool& raised, bool& unsupported) {
    int m = 0, intron_dst = -1; int32_t intron_len = 0;
    for (int k = 0; k < lo; k++) {
        u64 c = src[k];
        if (k == intron_idx) { intron_dst = m; intron_len = cg_len(c) + intron_delta; }
        dst[m++] = c;
    }
    int ne = hi - lo;
    if (nb >= 0) {
        if (ne == 0) return -1;
        if (at_end) {
            for (int k = lo; k < hi - 1; k++) dst[m++] = src[k];
            u64 last = src[hi - 1];
            if (cg_op(last) == 'M') dst[m++] = cg_make('M', cg_len(last) + nb);
            else { dst[m++] = last; dst[m++] = cg_make('M', nb); }
        } else {
            u64 first = src[lo];
            if (cg_op(first) == 'M') dst[m++] = cg_make('M', cg_len(first) + nb);
            else { dst[m++] = cg_make('M', nb); dst[m++] = first; }
            for (int k = lo + 1; k < hi; k++) dst[m++] = src[k];
        }
    } else {
        int64_t need = -(int64_t)nb;
        if (at_end) {
            int k = hi - 1; int32_t keep_len = -1;
            for (; k >= lo; k--) {
                int64_t rem = (int64_t)cg_len(src[k]) - need;
                if (rem > 0) { keep_len = (int32_t)rem; break; }
                if (rem == 0) { k--; break; }
                need = -rem;
            }
            // ops lo..k survive; op k shortened when keep_len >= 0
            for (int j = lo; j <= k; j++) dst[m++] = (j == k && keep_len >= 0) ? cg_make(cg_op(src[j]), keep_len) : src[j];
        } else {
            int k = lo; int32_t keep_len = -1;
            for (; k < hi; k++) {
                int64_t rem = (int64_t)cg_len(src[k]) - need;
                if (rem > 0) { keep_len = (int32_t)rem; break; }
                if (rem == 0) { k++; break; }
                need = -rem;
            }
            for (int j = k; j < hi; j++) dst[m++] = (j == k && keep_len >= 0) ? cg_make(cg_op(src[j]), keep_len) : src[j];
        }
    }
    for (int k = hi; k < n; k++) {
        u64 c = src[k];
        if (k == intron_idx) { intron_dst = m; intron_len = cg_len(c) + intron_delta; }
        dst[m++] = c;
    }
    if (intron_dst >= 0 && (!intron_text_rules(dst, intron_dst, intron_len, unsupported) || !other_dminus_ok(dst, m, intron_dst))) raised = true;
    return m;
}

TC_HD int seg_insert(const u64* s, int n, u64* d, int64_t qpos, u64 ns) {
    int m = 0; int64_t q = 0; bool done = seg_len(ns) == 0;
    for (int j = 0; j < n; j++) {
        int64_t len = seg_len(s[j]);
        if (!done && qpos <= q) { d[m++] = ns; done = true; }
        if (!done && qpos < q + len) {
            int64_t l1 = qpos - q;
            d[m++] = seg_make(seg_kind(s[j]), seg_src(s[j]), l1);
            d[m++] = ns;
            d[m++] = seg_make(seg_kind(s[j]), seg_src(s[j]) + l1, len - l1);
            done = true; q += len; continue;
        }
        d[m++] = s[j]; q += len;
    }
    if (!done) d[m++] = ns;
    return m;
}
TC_HD int seg_delete(const u64* s, int n, u64* d, int64_t qa, int64_t qb) {
    int m = 0; int64_t q = 0;
    for (int j = 0; j < n; j++) {
        int64_t len = seg_len(s[j]), a = q, b = q + len;
        q = b;
        if (b <= qa || a >= qb || qa >= qb) { d[m++] = s[j]; continue; }
        if (a < qa) d[m++] = seg_make(seg_kind(s[j]), seg_src(s[j]), qa - a);
        if (b > qb) d[m++] = seg_make(seg_kind(s[j]), seg_src(s[j]) + (qb - a), b - qb);
    }
    return m;
}

struct JnState {              // the transcript copy being edited by attempt_jn_correction
    const u64* cig; int nc; const u64* seg; int ns; int64_t seqlen;
};

// fix_one_side_of_junction (TC.py:1415-1479).  `bound` is mutated before the checks that can
// still fail, exactly like intronBound.pos (Q18).  Returns false for every path on which the
// reference raises inside the try at TC.py:1323-1342.
TC_HD bool fix_side(const Params& P, int chrom, int k, int side, int d, int32_t* bound, JnState& S,
                    u64* cig_dst, u64* seg_dst, bool& unsupported) {
    if (d == 0) return true;
    bool raised = false;
    ExonLoc e; int nN; int t = side == 0 ? k : k + 1;
    if (!locate_exon(S.cig, S.nc, t, e, nN)) return false;            // exonSeqs[targetExon] IndexError
    const GenomeDev& G = P.G; int64_t goff = G.chrom_off[chrom] - 1;
    int ns2; int64_t newlen = S.seqlen;
    if (side == 0) {
        if (d > 0) {                                                  // case 1: extend exon end from the genome
            int32_t b = *bound;
            if (b < 1) return false;
            int gl = fetch_len(G, chrom, b, (int64_t)b + d - 1);
            ns2 = seg_insert(S.seg, S.ns, seg_dst, e.q1, seg_make(1, goff + b, gl)); newlen += gl;
        } else {                                                      // case 3: exon[0:d]
            int64_t m = e.q1 - e.q0; if (m > -d) m = -d;
            ns2 = seg_delete(S.seg, S.ns, seg_dst, e.q1 - m, e.q1); newlen -= m;
        }
        *bound += d;
        if (k >= nN) return false;                                    // intronCIGARs[jn_number] IndexError
        int ncn = cig_edit(S.cig, S.nc, cig_dst, e.lo, e.hi, true, d, e.nidx, -d, raised, unsupported);
        if (ncn < 0 || raised) return false;
        S.nc = ncn;
    } else {
        if (d < 0) {                                                  // case 2: prepend to the next exon
            int64_t start = (int64_t)*bound + 1 + d;
            if (start < 1) return false;
            int gl = fetch_len(G, chrom, start, (int64_t)*bound);
            ns2 = seg_insert(S.seg, S.ns, seg_dst, e.q0, seg_make(1, goff + start, gl)); newlen += gl;
        } else {                                                      // case 4: exon[d:]
            int64_t m = e.q1 - e.q0; if (m > d) m = d;
            ns2 = seg_delete(S.seg, S.ns, seg_dst, e.q0, e.q0 + m); newlen -= m;
        }
        *bound += d;
        // the intron before exon k+1 is the k-th N op
        ExonLoc ek; ek.nidx = -1; int nn2; locate_exon(S.cig, S.nc, k, ek, nn2);
        int ncn = cig_edit(S.cig, S.nc, cig_dst, e.lo, e.hi, false, -d, ek.nidx, d, raised, unsupported);
        if (ncn < 0 || raised) return false;
        S.nc = ncn;
    }
    S.cig = cig_dst; S.seg = seg_dst; S.ns = ns2; S.seqlen = newlen;
    return newlen == cig_query_len(S.cig, S.nc);                      // TC.py:1476-1477
}

// Transcript.__init__: junction objects, motif codes, canonical / annotated flags (tr.py:56-70, sj.py:8-68).
// Returns true when cleanNoncanonical has work to do for this read.
TC_HD bool junction_init_read(const Params& P, int64_t i) {
    const BatchDev& B = P.B; const WorkDev& W = P.W; const tc_options& O = P.O;
    uint32_t st = W.status[i];
    if ((st & TC_ST_CLASS_MASK) != 0) return false;
    const int chrom = B.chrom[i]; const int strand = (B.flag[i] == 16) ? 1 : 0;     // tr.py:51-53
    const int nj = W.n_jn[i]; const uint8_t tags = B.tags[i];
    WsView v = ws_view(W, i); int32_t* J = v.jn;
    // ---- Transcript.__init__: junction objects from jI (input tag, else from the CIGAR)
    if (nj > 0) {
        if (tags & TAG_JI) {
            const int32_t* ji = B.ji + B.ji_off[i];
            for (int k = 0; k < nj; k++) { J[k * JN_STRIDE] = ji[2 * k]; J[k * JN_STRIDE + 1] = ji[2 * k + 1]; }
        } else if (!(W.mflags[i] & MF_JI_DONE)) {                     // compute_jI (tr.py:368-393), unless plan_read did it on its walk
            int64_t g = B.pos[i]; int k = 0;
            for (int64_t c = B.cig_off[i]; c < B.cig_off[i + 1]; c++) {
                int op = B.cig_op[c]; int32_t ct = B.cig_len[c];
                if (op == 'N') { J[k * JN_STRIDE] = (int32_t)g; J[k * JN_STRIDE + 1] = (int32_t)(g + ct - 1); k++; }
                if (op != 'S' && op != 'I') g += ct;
            }
        }
    }
    bool canon = true, annot = true;
    for (int k0 = 0; k0 < nj; k0 += 4) {                // four junctions at a time: their genome / table loads overlap
        int32_t b0[4], b1[4]; int code[4];
#pragma unroll
        for (int u = 0; u < 4; u++) if (k0 + u < nj) { b0[u] = J[(k0 + u) * JN_STRIDE]; b1[u] = J[(k0 + u) * JN_STRIDE + 1]; }
#pragma unroll
        for (int u = 0; u < 4; u++) if (k0 + u < nj) { code[u] = 0; junction_code(P, chrom, strand, b0[u], b1[u], b0[u], b1[u], code[u]); }
#pragma unroll
        for (int u = 0; u < 4; u++) if (k0 + u < nj) {
            int32_t* r = J + (k0 + u) * JN_STRIDE;
            r[2] = b0[u]; r[3] = b1[u]; r[4] = code[u]; r[5] = b0[u]; r[6] = b1[u]; r[7] = code[u];
            if (code[u] == 0 || code[u] == 20) canon = false;
            if (code[u] < 20) annot = false;
        }
    }
    if (!(tags & TAG_JM)) st |= TC_ST_JM_DEVICE | TC_ST_JI_DEVICE;     // tr.py:69-70
    else if (!(tags & TAG_JI)) st |= TC_ST_JI_DEVICE;                  // tr.py:56-57
    int32_t* C = W.counters + 10 * i;
    const bool run = !(st & TC_ST_FALLBACK) && O.correct_sjs && P.sj_enabled;
    if (run) { C[TC_C_COR_SJ] = 0; C[TC_C_UNC_SJ] = 0; }                // TC.py:1192-1193
    if (canon) st |= TC_ST_CANONICAL;
    if (annot) st |= TC_ST_ALL_ANNOT;
    W.status[i] = st;
    return run && !canon;
}

// cleanNoncanonical (TC.py:1186-1231) for a read whose junction objects exist and that is not canonical.
TC_HD void junction_fix_read(const Params& P, int64_t i) {
    const BatchDev& B = P.B; const WorkDev& W = P.W; const tc_options& O = P.O;
    uint32_t st = W.status[i];
    const int chrom = B.chrom[i]; const int strand = (B.flag[i] == 16) ? 1 : 0;
    const int nj = W.n_jn[i];
    WsView v = ws_view(W, i); int32_t* J = v.jn;
    int32_t* C = W.counters + 10 * i;
    bool canon = false, annot = (st & TC_ST_ALL_ANNOT) != 0;
    st &= ~(uint32_t)(TC_ST_CANONICAL | TC_ST_ALL_ANNOT);
    {
        {
            int nte = W.n_te[i]; int cJ = 0, uJ = 0;
            const int64_t Lq = B.seq_off[i + 1] - B.seq_off[i];
            // committed transcript state
            int cb = W.cig_buf[i], sb = W.seg_buf[i]; int nc = W.n_cig_cur[i], ns = W.n_seg_cur[i];
            if (cb < 0) {                                             // still the input: bring it into the workspace
                nc = 0; for (int64_t c = B.cig_off[i]; c < B.cig_off[i + 1]; c++) v.cig[0][nc++] = cg_make(B.cig_op[c], B.cig_len[c]);
                ns = 0; if (Lq > 0) v.seg[0][ns++] = seg_make(0, 0, Lq);
                cb = 0; sb = 0;
            }
            int64_t seqlen = 0; for (int s = 0; s < ns; s++) seqlen += seg_len(v.seg[sb][s]);
            bool whole_fail = false, unsupported = false, any_ok = false;
            const bool tie_right = O.sj_tie_right != 0;
            for (int k = 0; k < nj && !whole_fail; k++) {
                int32_t* r = J + k * JN_STRIDE;
                if (!(r[4] == 0 || r[4] == 20)) continue;
                const int32_t id_a = r[2], id_b = r[3];
                const int ds = strand == 0 ? 0 : 1, as = 1 - ds;       // sj.py:29-41
                int d_don, d_acc; bool ok_t;
                if (O.sj_ignore_strand) {
                    ok_t = sj_nearest(P.S.don_u, P.S.rng_don_u, P.S.n_prefix, (u64)chrom, (int64_t)r[ds] - 1, tie_right, d_don) &&
                           sj_nearest(P.S.acc_u, P.S.rng_acc_u, P.S.n_prefix, (u64)chrom, (int64_t)r[as] - 1, tie_right, d_acc);
                } else {
                    u64 pre = ((u64)chrom << 1) | (u64)strand;
                    ok_t = sj_nearest(P.S.don_s, P.S.rng_don_s, P.S.n_prefix, pre, (int64_t)r[ds] - 1, tie_right, d_don) &&
                           sj_nearest(P.S.acc_s, P.S.rng_acc_s, P.S.n_prefix, pre, (int64_t)r[as] - 1, tie_right, d_acc);
                }
                if (!ok_t) { whole_fail = true; break; }              // Q23: raises outside the inner try
                int ad = d_don < 0 ? -d_don : d_don, aa = d_acc < 0 ? -d_acc : d_acc;
                int dist = ((d_don < 0) != (d_acc < 0)) ? ad + aa : (ad > aa ? ad - aa : aa - ad);   // TC.py:1371-1375
                if (dist > O.max_sj_offset || ad > 2 * O.max_sj_offset || aa > 2 * O.max_sj_offset) {
                    uJ++; te_put(v.te, nte, 3, id_a, id_b, dist, 1, 3); continue;
                }
                // the two scratch buffers that are not the committed one
                int t1 = (cb + 1) % 3, t2 = (cb + 2) % 3, u1 = (sb + 1) % 3, u2 = (sb + 2) % 3;
                JnState S; S.cig = v.cig[cb]; S.nc = nc; S.seg = v.seg[sb]; S.ns = ns; S.seqlen = seqlen;
                int used_c = cb, used_s = sb;
                bool ok = true;
                if (d_don != 0) { ok = fix_side(P, chrom, k, ds, d_don, &r[ds], S, v.cig[t1], v.seg[u1], unsupported); used_c = t1; used_s = u1; }
                if (ok && d_acc != 0) {
                    int tc = used_c == cb ? t1 : t2, ts = used_s == sb ? u1 : u2;
                    ok = fix_side(P, chrom, k, as, d_acc, &r[as], S, v.cig[tc], v.seg[ts], unsupported); used_c = tc; used_s = ts;
                }
                if (ok) {                                             // update_post_ncsj_correction (TC.py:1279-1294)
                    r[2] = r[0]; r[3] = r[1];
                    int code;
                    if (!junction_code(P, chrom, strand, r[0], r[1], r[2], r[3], code)) ok = false;
                    else r[4] = code;
                }
                if (ok) {
                    cb = used_c; sb = used_s; nc = S.nc; ns = S.ns; seqlen = S.seqlen; any_ok = true;
                    canon = true; annot = true;
                    for (int m = 0; m < nj; m++) {                     // tags + flags from the shared junction objects
                        int32_t* rr = J + m * JN_STRIDE;
                        rr[5] = rr[0]; rr[6] = rr[1]; rr[7] = rr[4];
                        if (rr[4] == 0 || rr[4] == 20) canon = false;
                        if (rr[4] < 20) annot = false;
                    }
                    cJ++; te_put(v.te, nte, 3, id_a, id_b, dist, 0, 0);
                } else {
                    uJ++; te_put(v.te, nte, 3, id_a, id_b, dist, 1, 4);
                }
            }
            if (whole_fail) {                                         // Q13/Q23: original read, TE dropped
                st |= TC_ST_FALLBACK; st &= ~(uint32_t)TC_ST_CIGAR_REWRITTEN;
                W.cig_buf[i] = -1; W.seg_buf[i] = -1; W.n_te[i] = 0;
                W.te_cnt[3 * i] = W.te_cnt[3 * i + 1] = W.te_cnt[3 * i + 2] = 0;
                W.refresh[i] = 0; W.nm_extra[i] = 0;
                // junction flags go back to the init values (no fix was committed before the raise)
                canon = true; annot = true;
                for (int m = 0; m < nj; m++) { int c = J[m * JN_STRIDE + 4]; if (c == 0 || c == 20) canon = false; if (c < 20) annot = false; }
            } else {
                C[TC_C_COR_SJ] = cJ; C[TC_C_UNC_SJ] = uJ;
                W.n_te[i] = nte;
                if (any_ok) {
                    W.cig_buf[i] = (int8_t)cb; W.seg_buf[i] = (int8_t)sb; W.n_cig_cur[i] = nc; W.n_seg_cur[i] = ns; W.seq_len_out[i] = (int32_t)seqlen;
                    W.delta_size[i] = (int32_t)dl_list_size(v.seg[sb], ns);
                    W.refresh[i] = 1; W.nm_extra[i] = 0;
                    st |= TC_ST_CIGAR_REWRITTEN | TC_ST_JM_DEVICE | TC_ST_JI_DEVICE;
                }
                if (unsupported) st |= TC_ST_ERR_INTRON;
            }
        }
    }
    if (canon) st |= TC_ST_CANONICAL;
    if (annot) st |= TC_ST_ALL_ANNOT;
    W.status[i] = st;
}

TC_HD void junction_read(const Params& P, int64_t i) { if (junction_init_read(P, i)) junction_fix_read(P, i); }

// exact output sizes of a read (input to the four exclusive scans that place tc_emit's output)
TC_HD void sizes_read(const Params& P, int64_t i) {
    const BatchDev& B = P.B; const WorkDev& W = P.W;
    uint32_t st = W.status[i];
    if ((st & TC_ST_CLASS_MASK) != 0) { W.o_seq[i] = W.o_cig[i] = W.o_te[i] = W.o_jn[i] = W.o_delta[i] = 0; W.seq_len_out[i] = 0; W.plan[i].w[EP_STATUS] = (int32_t)st; return; }
    int64_t sl;
    if (W.seg_buf[i] < 0) sl = B.seq_off[i + 1] - B.seq_off[i];
    else sl = W.seq_len_out[i];                    // kept current by plan_read / the junction rescue (the sum of the segment lengths)
    W.o_seq[i] = (sl + 15) & ~15ll;                 // 16-byte aligned rows: tc_emit stores whole uint4
    W.seq_len_out[i] = (int32_t)sl;
    const int ncig = W.cig_buf[i] < 0 ? (int)(B.cig_off[i + 1] - B.cig_off[i]) : W.n_cig_cur[i];
    W.o_cig[i] = W.cig_buf[i] < 0 ? 0 : ncig;      // an untouched CIGAR is not sent back: the caller still holds its text
    W.o_te[i] = 2 * W.n_te[i] - (W.te_cnt[3 * i] + W.te_cnt[3 * i + 1] + W.te_cnt[3 * i + 2]);   // 8-byte words: a junction row takes two (tc_te_word)
    W.o_jn[i] = W.n_jn[i];
    const int32_t dlen = W.seg_buf[i] < 0 ? -1 : W.delta_size[i];     // -1: the SEQ is the input SEQ
    W.o_delta[i] = dlen < 0 ? 0 : dlen;
    EmitPlan ep;
    for (int k = 0; k < 16; k++) ep.w[k] = 0;
    ep.w[EP_FLAGS] = (W.refresh[i] ? 1 : 0) | ((W.cig_buf[i] + 1) << 2) | ((W.seg_buf[i] + 1) << 4);
    ep.w[EP_NCIG] = ncig; ep.w[EP_NSEG] = W.n_seg_cur[i]; ep.w[EP_NTE] = W.n_te[i];
    ep.w[EP_TEMM] = W.te_cnt[3 * i]; ep.w[EP_TEINS] = W.te_cnt[3 * i + 1]; ep.w[EP_TEDEL] = W.te_cnt[3 * i + 2];
    ep.w[EP_NJN] = W.n_jn[i]; ep.w[EP_OUTLEN] = (int32_t)sl; ep.w[EP_NMEXTRA] = W.nm_extra[i];
    ep.w[EP_CIGCAP] = W.cig_cap[i]; ep.w[EP_DLEN] = dlen;
    ep.w[EP_STATUS] = (int32_t)st; ep.w[EP_CHROM] = B.chrom[i]; ep.w[EP_POS] = B.pos[i];
#ifndef TC_HOSTSIM
    int4* dst = reinterpret_cast<int4*>(W.plan[i].w);
    const int4* src = reinterpret_cast<const int4*>(ep.w);
    for (int k = 0; k < 4; k++) dst[k] = src[k];
#else
    for (int k = 0; k < 16; k++) W.plan[i].w[k] = ep.w[k];
#endif
}

// ---------------------------------------------------------------------------------- dry run

TC_HD void dry_sizes(const Params& P, int64_t i) {
    const WorkDev& W = P.W;
    W.o_seq[i] = W.o_cig[i] = W.o_jn[i] = W.o_delta[i] = 0; W.o_te[i] = W.n_te[i]; W.seq_len_out[i] = 0;
    EmitPlan ep;
    for (int k = 0; k < 16; k++) ep.w[k] = 0;
    ep.w[EP_NTE] = W.n_te[i]; ep.w[EP_CIGCAP] = W.cig_cap[i]; ep.w[EP_NJN] = W.n_jn[i]; ep.w[EP_DLEN] = -1;
    ep.w[EP_STATUS] = (int32_t)W.status[i];
    for (int k = 0; k < 16; k++) W.plan[i].w[k] = ep.w[k];
}

// second half of the plan row: offsets that exist only after the exclusive scans
TC_HD void pack_read(const Params& P, int64_t i) {
    const BatchDev& B = P.B; const WorkDev& W = P.W;
    int32_t* w = W.plan[i].w;
    const int64_t ws = W.ws_off[i], os = W.o_seq[i], oc = W.o_cig[i], ot = W.o_te[i], oj = W.o_jn[i], ib = B.seq_off[i], c0 = B.cig_off[i];
    w[EP_WS_LO] = (int32_t)(uint32_t)ws; w[EP_WS_HI] = (int32_t)(ws >> 32);
    w[EP_OSEQ_LO] = (int32_t)(uint32_t)os; w[EP_OSEQ_HI] = (int32_t)(os >> 32);
    w[EP_OCIG_LO] = (int32_t)(uint32_t)oc; w[EP_OCIG_HI] = (int32_t)(oc >> 32);
    w[EP_OTE_LO] = (int32_t)(uint32_t)ot; w[EP_OTE_HI] = (int32_t)(ot >> 32);
    w[EP_OJN_LO] = (int32_t)(uint32_t)oj; w[EP_OJN_HI] = (int32_t)(oj >> 32);
    w[EP_INB_LO] = (int32_t)(uint32_t)ib; w[EP_INB_HI] = (int32_t)(ib >> 32);
    w[EP_LQ] = (int32_t)(B.seq_off[i + 1] - ib);
    w[EP_CIG0_LO] = (int32_t)(uint32_t)c0; w[EP_CIG0_HI] = (int32_t)(c0 >> 32);
    const int64_t od = W.o_delta[i];
    w[EP_ODL_LO] = (int32_t)(uint32_t)od; w[EP_ODL_HI] = (int32_t)(od >> 32);
}

// dryRun (TC.py:1615-1678): positional inventory of X / D / I from the MD x CIGAR merge.
TC_HD void dry_read(const Params& P, int64_t i) {
    const BatchDev& B = P.B; const WorkDev& W = P.W;
    uint32_t st = W.status[i];
    if ((st & TC_ST_CLASS_MASK) != 0) return;
    int32_t* C = W.counters + 10 * i;
    C[TC_C_UNC_DEL] = C[TC_C_UNC_INS] = C[TC_C_UNC_MM] = 0;
    if (st & TC_ST_NMMD_NONE) { W.status[i] = st | TC_ST_ERR_MERGE; return; }
    WsView v = ws_view(W, i);
    int64_t g = B.pos[i]; int nte = 0, uM = 0, uI = 0, uD = 0;
    int64_t ci = B.cig_off[i]; const int64_t cend = B.cig_off[i + 1];
    int32_t crem = ci < cend ? B.cig_len[ci] : 0;
    MdCur md; md_cursor(P, B, i, md);
    int mop = 0; int32_t mrem = 0; bool have_md = md_next(md, mop, mrem);
    while (true) {
        int op; int32_t ct;
        if (!have_md && ci >= cend) break;
        if (ci >= cend) { W.status[i] = st | TC_ST_ERR_MERGE; return; }
        int cop = B.cig_op[ci];
        if (cop == 'H' || cop == 'S' || cop == 'I' || cop == 'N') { op = cop; ct = crem; ci++; if (ci < cend) crem = B.cig_len[ci]; }
        else {
            if (!have_md) { W.status[i] = st | TC_ST_ERR_MERGE; return; }
            if (crem < mrem) { mrem -= crem; op = cop; ct = crem; ci++; if (ci < cend) crem = B.cig_len[ci]; }
            else if (crem > mrem) { crem -= mrem; op = mop; ct = mrem; have_md = md_next(md, mop, mrem); }
            else { op = mop; ct = mrem; have_md = md_next(md, mop, mrem); ci++; if (ci < cend) crem = B.cig_len[ci]; }
        }
        if (op == 'M') g += ct;
        else if (op == 'X') { te_put(v.te, nte, 0, (int32_t)g, 0, ct, 1, 5); uM++; g += ct; }
        else if (op == 'D') { te_put(v.te, nte, 2, (int32_t)(g - 1), (int32_t)(g + ct - 1), ct, 1, 5); uD++; g += ct; }
        else if (op == 'I') { te_put(v.te, nte, 1, (int32_t)(g - 1), (int32_t)(g + ct - 1), ct, 1, 5); uI++; }
        else if (op == 'N' || op == 'H') g += ct;
    }
    C[TC_C_UNC_DEL] = uD; C[TC_C_UNC_INS] = uI; C[TC_C_UNC_MM] = uM;
    W.n_te[i] = nte;
}

// tc_b200.cu - kernels and C ABI of the B200 correction engine (see include/tc_b200.h).
//
// Pipeline of one batch (all on the context's stream):
//   k_init      warp/read   status word; NM/MD regeneration for reads lacking a tag (tr.py:47-48)
//   k_measure   thread/read workspace sizing                      -> exclusive scan
//   k_plan      thread/read mismatch + indel stages (tc_plan.cuh)
//   k_junction  thread/read junction objects + noncanonical junction rescue
//   k_sizes     thread/read exact output sizes                    -> 4 exclusive scans
//   k_emit      warp/read   CIGAR / TE / jM,jI compaction, SEQ materialisation (genome gather),
//                           NM/MD regeneration (genome gather + compare)
// With -DTC_HOSTSIM the same ABI is built for the CPU with serial loops instead of kernels;
// that build lives under tests/hostsim and is TEST INFRASTRUCTURE for the host-side code
// (parser, renderer, CLI) on machines without a GPU.  The product never loads it.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "tc_plan.cuh"

#ifndef TC_HOSTSIM
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>
#endif

// ---------------------------------------------------------------------------------- outputs

struct OutDev {
    int32_t* nm; int64_t* md_off; int32_t* md_len;
    uint8_t* md_pool; unsigned long long* md_cursor; int64_t md_cap;
    int64_t md_inline;       // MD values of <= 8 bytes live at md_inline + 8*i (no allocation); longer ones in the pool behind
    uint8_t* seq; uint8_t* cig_op; int32_t* cig_len; tc_te* te; int8_t* jm; int32_t* ji;
    int32_t* exact_list; unsigned* exact_count;   // reads whose NM/MD need the token-emitting walk (k_nmmd_exact)
};
struct InitDev { uint8_t* pool; unsigned long long* cursor; int64_t cap; };

TC_HD int ndigits(uint32_t v) {
    return v < 10u ? 1 : v < 100u ? 2 : v < 1000u ? 3 : v < 10000u ? 4 : v < 100000u ? 5 : v < 1000000u ? 6 : v < 10000000u ? 7 :
           v < 100000000u ? 8 : v < 1000000000u ? 9 : 10;
}
// decimal digits of v (nd of them, nd <= 8) packed little-endian into one 64-bit word
TC_HD unsigned long long dec_word(uint32_t v, int nd) {
    unsigned long long w = 0;
    for (int k = nd - 1; k >= 0; k--) { w |= (unsigned long long)('0' + v % 10) << (8 * k); v /= 10; }
    return w;
}
// same for v < 10000 (nd <= 4), without a loop
TC_HD unsigned long long dec_word4(uint32_t v, int nd) {
    const uint32_t d3 = v / 1000u, r3 = v - d3 * 1000u, d2 = r3 / 100u, r2 = r3 - d2 * 100u, d1 = r2 / 10u, d0 = r2 - d1 * 10u;
    const uint32_t w = (d3 | (d2 << 8) | (d1 << 16) | (d0 << 24)) + 0x30303030u;
    return nd > 0 ? (unsigned long long)(w >> (8 * (4 - nd))) : 0ull;
}
TC_HD void write_dec(uint8_t* dst, uint32_t v, int nd) {
    for (int k = nd - 1; k >= 0; k--) { dst[k] = (uint8_t)('0' + v % 10); v /= 10; }
}

struct CigSrc { const uint8_t* op; const int32_t* len; const u64* packed; int n; };
TC_HD void cig_get(const CigSrc& c, int k, int& op, int32_t& len) {
    if (c.packed) { op = cg_op(c.packed[k]); len = cg_len(c.packed[k]); }
    else { op = c.op[k]; len = c.len[k]; }
}

// TE rows back into the reference's order (stage order: mismatches, insertions, deletions, junctions;
// positional inside a stage - the planner wrote them positionally), the final CIGAR and the jM / jI
// snapshot of one read -> their places in the dense outputs.  Thread per read (k_pack): these are a few
// hundred bytes per read, cheaper here than as warp-wide phases of the output kernel.
TC_HD void small_copies_read(const Params& P, const OutDev& O, int64_t i, int dry) {
    const BatchDev& B = P.B; const WorkDev& W = P.W;
    if ((W.status[i] & TC_ST_CLASS_MASK) != 0) return;
    const WsView v = ws_view(W, i);
    const int nte = W.n_te[i];
    const u64* src = reinterpret_cast<const u64*>(v.te); u64* dst = reinterpret_cast<u64*>(O.te + W.o_te[i]);
    if (dry) { for (int k = 0; k < 2 * nte; k++) dst[k] = src[k]; return; }
    int slot[4]; slot[0] = 0; slot[1] = W.te_cnt[3 * i]; slot[2] = slot[1] + W.te_cnt[3 * i + 1]; slot[3] = slot[2] + W.te_cnt[3 * i + 2];
    for (int k = 0; k < nte; k++) {
        const u64 a = src[2 * k], b = src[2 * k + 1];
        const int t = (int)((b >> 32) & 3u);                          // tc_te.type
        const int s = t == 0 ? slot[0]++ : t == 1 ? slot[1]++ : t == 2 ? slot[2]++ : slot[3]++;
        dst[2 * s] = a; dst[2 * s + 1] = b;
    }
    const int cb = W.cig_buf[i]; const int64_t oc = W.o_cig[i];
    if (cb < 0) {
        const int64_t c0 = B.cig_off[i]; const int n = (int)(B.cig_off[i + 1] - c0);
        for (int k = 0; k < n; k++) { O.cig_op[oc + k] = B.cig_op[c0 + k]; O.cig_len[oc + k] = B.cig_len[c0 + k]; }
    } else {
        const u64* c = v.cig[cb]; const int n = W.n_cig_cur[i];
        for (int k = 0; k < n; k++) { O.cig_op[oc + k] = (uint8_t)cg_op(c[k]); O.cig_len[oc + k] = cg_len(c[k]); }
    }
    const int nj = W.n_jn[i]; const int64_t oj = W.o_jn[i];
    for (int k = 0; k < nj; k++) { const int32_t* r = v.jn + k * JN_STRIDE; O.jm[oj + k] = (int8_t)r[7]; O.ji[2 * (oj + k)] = r[5]; O.ji[2 * (oj + k) + 1] = r[6]; }
}

#ifndef TC_HOSTSIM
// ================================================================================== CUDA

#define FULL 0xffffffffu
// fields of a read's plan row (one word per lane, `pw`)
#define PF(k) __shfl_sync(FULL, pw, (k))
#define PF64(k) ((int64_t)(((u64)(uint32_t)PF((k) + 1) << 32) | (u64)(uint32_t)PF(k)))
#define WARPS_PER_BLOCK 8
#define EMIT_CAP 4096            // reads whose new SEQ fits are staged in shared memory
#define EMIT_MAXBP 192
#define EMIT_MAXGI 192
#define EMIT_MAXD 44

#define EMIT_MAXRUN 32          // M runs of a read that tier 1 handles (one lane per run)
struct __align__(16) EmitSmem {  // per warp
    uint8_t seq[EMIT_CAP + 32];
    uint8_t chunk_bp[(EMIT_CAP >> 4) + 16];                             // slice in force at the start of each 16-byte chunk
    union {
        struct {                                                        // while SEQ is built
            uint16_t bp_pos[EMIT_MAXBP]; int16_t bp_shift[EMIT_MAXBP]; // starts of input-SEQ slices: output offset, src-dst shift
            uint16_t gi_out[EMIT_MAXGI]; uint16_t gi_len[EMIT_MAXGI]; int32_t gi_src[EMIT_MAXGI];   // genome slices
        };
        struct {                                                        // while NM/MD is verified (tier 1)
            uint16_t run_q[EMIT_MAXRUN], run_ct[EMIT_MAXRUN]; int32_t run_g[EMIT_MAXRUN];   // M runs: query offset, length, reference offset
            int32_t rb_w[EMIT_MAXRUN], rb_coff[EMIT_MAXRUN], rb_goff[EMIT_MAXRUN];         // runs that own whole chunks: first item, chunk and genome offsets
            int32_t d_g[EMIT_MAXD], d_ct[EMIT_MAXD], d_run[EMIT_MAXD];    // D ops: reference offset, length, match run before
        };
    };
};
#define EMIT_SMEM_BYTES (1024 + WARPS_PER_BLOCK * sizeof(EmitSmem))

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int& total) {
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int y = __shfl_up_sync(FULL, x, d); if (lane >= d) x += y; }
    total = __shfl_sync(FULL, x, 31);
    return x - v;
}

__device__ __forceinline__ bool warp_range_has_exc(const GenomeDev& G, int64_t g0, int64_t g1, int lane) {
    const int64_t b0 = g0 >> 8; const int nblk = (int)(((g1 - 1) >> 8) - b0) + 1;
    bool hit = false;
    for (int t = lane; t < nblk; t += 32) { const int64_t b = b0 + t; hit |= (G.excblk[b >> 5] >> (b & 31)) & 1u; }
    return __any_sync(FULL, hit);
}

struct NmMd { int nm, len, ntok, carry; bool none; };

// getNMandMDFlags (tr.py:280-342, Q8/Q9).  One warp walks the CIGAR (ops are fetched 32 at a
// time, one per lane, and broadcast with shuffles; cursors are 32-bit offsets from POS).
//  FAST (SEQ staged in shared memory, 16-byte aligned): every lane takes 16 bases of an M run -
//    five LDS.32 + funnel shifts for the read bytes, two coalesced words of the 2-bit plane,
//    a 256-entry code->ASCII table, one XOR per 4 bases; a 512-base chunk without difference
//    costs ~50 instructions.  A chunk with any difference (or an M run that touches a literal
//    run of the genome) is redone by the exact path below.
//  exact path: lanes take consecutive bases, gather the genome byte (2-bit + case plane,
//    literal runs), compare case-insensitively, and compact the mismatch tokens with
//    ballot / popc / prefix sums.
// WRITE=false measures NM and the MD length; WRITE=true writes the MD bytes.
template <bool WRITE, bool FAST>
__device__ NmMd warp_nmmd(const GenomeDev& G, const uint8_t* seq, const uint32_t* lut, const CigSrc& cs, int chrom,
                          int64_t pos, int lane, uint8_t* dst) {
    const int64_t goff = G.chrom_off[chrom] - 1, clen = G.chrom_len[chrom];
    const int64_t gbase = goff + pos;                   // plane index of reference offset 0
    NmMd R;
    int nm = 0, len = 0, carry = 0, ntok = 0, q = 0, gr = 0; bool none = false;
    // reference span: one literal-run check and one contig-end check for the whole read
    bool span_exc, careful;
    {
        int64_t span = 0;
        for (int k = lane; k < cs.n; k += 32) { int op; int32_t ct; cig_get(cs, k, op, ct); if (op == 'M' || op == 'D' || op == 'N' || op == 'H') span += ct; }
#pragma unroll
        for (int d = 16; d; d >>= 1) span += __shfl_xor_sync(FULL, span, d);
        careful = pos < 1 || pos + span - 1 > clen || span > 0x7fffffffll;
        span_exc = careful || (span > 0 && warp_range_has_exc(G, gbase, gbase + span, lane));
    }
    for (int k0 = 0; k0 < cs.n && !none; k0 += 32) {
        int op_l = 0; int32_t ct_l = 0;
        if (k0 + lane < cs.n) cig_get(cs, k0 + lane, op_l, ct_l);
        const int nround = min(32, cs.n - k0);
        for (int j = 0; j < nround; j++) {
            const int op = __shfl_sync(FULL, op_l, j); const int32_t ct = __shfl_sync(FULL, ct_l, j);
            if (op == 'M') {
                if (ct <= 0) continue;
                if (careful) { const int64_t g = pos + gr; if (g < 1 || g + ct - 1 > clen) { none = true; break; } }   // tr.py:303-304
                const int64_t g0 = gbase + gr;
                const bool exc = span_exc && warp_range_has_exc(G, g0, g0 + ct, lane);
                for (int32_t c0 = 0; c0 < ct; c0 += (FAST ? 512 : 32)) {
                    const int span = min(FAST ? 512 : 32, ct - c0);
                    if (FAST) {
                        bool bad = exc;
                        if (!exc) {
                            const int32_t lo = c0 + 16 * lane; int nb = ct - lo; nb = nb < 0 ? 0 : (nb > 16 ? 16 : nb);
                            if (nb > 0) {
                                const int a = q + lo;
                                const uint32_t* sp = reinterpret_cast<const uint32_t*>(seq) + (a >> 2); const int sh = (a & 3) * 8;
                                const uint32_t x0 = sp[0], x1 = sp[1], x2 = sp[2], x3 = sp[3], x4 = sp[4];
                                const int64_t gb = g0 + lo; const uint32_t* gp = G.g2 + (gb >> 4);
                                const uint32_t gw = __funnelshift_r(gp[0], gp[1], (int)(gb & 15) * 2);
                                uint32_t d0 = (__funnelshift_r(x0, x1, sh) & 0xDFDFDFDFu) ^ lut[gw & 255u];
                                uint32_t d1 = (__funnelshift_r(x1, x2, sh) & 0xDFDFDFDFu) ^ lut[(gw >> 8) & 255u];
                                uint32_t d2 = (__funnelshift_r(x2, x3, sh) & 0xDFDFDFDFu) ^ lut[(gw >> 16) & 255u];
                                uint32_t d3 = (__funnelshift_r(x3, x4, sh) & 0xDFDFDFDFu) ^ lut[gw >> 24];
                                if (nb < 16) {                          // the lane that holds the end of the run
                                    const uint32_t m0 = nb >= 4 ? ~0u : ((1u << (8 * nb)) - 1u);
                                    const uint32_t m1 = nb >= 8 ? ~0u : (nb <= 4 ? 0u : ((1u << (8 * (nb - 4))) - 1u));
                                    const uint32_t m2 = nb >= 12 ? ~0u : (nb <= 8 ? 0u : ((1u << (8 * (nb - 8))) - 1u));
                                    const uint32_t m3 = nb <= 12 ? 0u : ((1u << (8 * (nb - 12))) - 1u);
                                    d0 &= m0; d1 &= m1; d2 &= m2; d3 &= m3;
                                }
                                bad = (d0 | d1 | d2 | d3) != 0u;
                            }
                        }
                        if (!__any_sync(FULL, bad)) { carry += span; continue; }
                    }
                    for (int32_t c = c0; c < c0 + span; c += 32) {       // exact path
                        const int32_t idx = c + lane; const bool valid = idx < c0 + span;
                        uint8_t rb = 0, gb = 0;
                        if (valid) { rb = seq[q + idx]; gb = exc ? genome_byte(G, g0 + idx) : genome_byte_fast(G, g0 + idx); }
                        const bool mism = valid && up8(rb) != up8(gb);
                        const unsigned mask = __ballot_sync(FULL, mism);
                        const int nvalid = min(32, c0 + span - c);
                        if (mask == 0) { carry += nvalid; continue; }
                        nm += __popc(mask); ntok += __popc(mask);
                        const unsigned below = mask & ((1u << lane) - 1u);
                        const int run_before = below ? (lane - (31 - __clz(below)) - 1) : carry + lane;
                        const int nd = ndigits((uint32_t)run_before);
                        int total; const int off = warp_excl_scan(mism ? nd + 1 : 0, lane, total);
                        if (WRITE && mism) { write_dec(dst + len + off, (uint32_t)run_before, nd); dst[len + off + nd] = gb; }
                        len += total;
                        carry = nvalid - 1 - (31 - __clz(mask));
                    }
                }
                q += ct; gr += ct;
            } else if (op == 'D') {
                int gl = ct;
                if (careful) {
                    const int64_t g = pos + gr;
                    if (g < 1) { none = true; break; }
                    gl = fetch_len(G, chrom, g, g + ct - 1);
                    if (gl == 0) { none = true; break; }                // tr.py:327-328
                }
                const int nd = ndigits((uint32_t)carry);
                if (WRITE) {
                    if (lane == 0) { write_dec(dst + len, (uint32_t)carry, nd); dst[len + nd] = '^'; }
                    for (int t = lane; t < gl; t += 32) dst[len + nd + 1 + t] = genome_byte(G, gbase + gr + t);
                }
                len += nd + 1 + gl; carry = 0; nm += ct; gr += ct; ntok++;
            } else if (op == 'I') { q += ct; nm += ct; }
            else if (op == 'S') q += ct;
            else if (op == 'N' || op == 'H') gr += ct;
        }
    }
    R.carry = carry;
    if (!none && carry > 0) {
        const int nd = ndigits((uint32_t)carry);
        if (WRITE && lane == 0) write_dec(dst + len, (uint32_t)carry, nd);
        len += nd;
    }
    R.nm = nm; R.len = len; R.none = none; R.ntok = ntok;
    return R;
}

// grabs `len` bytes of an MD pool for the warp; returns the offset or -1 when the pool is full
// (the cursor still advances, so the host learns the exact size needed for the retry)
__device__ __forceinline__ int64_t pool_grab(unsigned long long* cursor, int64_t cap, int len, int lane) {
    unsigned long long off = 0;
    if (lane == 0) off = atomicAdd(cursor, (unsigned long long)len);
    off = __shfl_sync(FULL, off, 0);
    return ((int64_t)off + len <= cap) ? (int64_t)off : -1;
}

// measure, allocate, write: MD of one read into a pool.  Reads without any token (the usual
// case after correction) only need the final run length.
template <bool FAST>
__device__ __forceinline__ void nmmd_to_pool(const GenomeDev& G, const uint8_t* seq, const uint32_t* lut, const CigSrc& cs,
                                             int chrom, int64_t pos, int lane, unsigned long long* cursor, int64_t cap,
                                             uint8_t* pool, int64_t inline_off, int& nm, int& len, int64_t& off, bool& none) {
    NmMd r = warp_nmmd<false, FAST>(G, seq, lut, cs, chrom, pos, lane, 0);
    none = r.none; nm = r.nm; len = r.none ? 0 : r.len;
    off = (inline_off >= 0 && len <= 8) ? inline_off : pool_grab(cursor, cap, len, lane);
    if (off >= 0 && len > 0) {
        if (r.ntok == 0) { if (lane == 0) write_dec(pool + off, (uint32_t)r.carry, len); }
        else warp_nmmd<true, FAST>(G, seq, lut, cs, chrom, pos, lane, pool + off);
    }
}

// status word of every read (thread per read); reads that lack NM or MD (tr.py:47-48) are listed for k_init_nmmd
__global__ void k_init(Params P, int32_t* list, unsigned* count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.B.n) return;
    const int cls = read_class(P.B.flag[i], P.B.chrom[i]);
    const uint8_t tags = P.B.tags[i];
    P.W.status[i] = (uint32_t)cls;
    if (cls == 0 && (!(tags & TAG_NM) || !(tags & TAG_MD))) list[atomicAdd(count, 1u)] = (int32_t)i;
}
// NM/MD bootstrap of the listed reads, one warp per read
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) k_init_nmmd(Params P, InitDev I, int32_t* nm_init, const int32_t* list, const unsigned* count) {
    const int lane = threadIdx.x & 31;
    const int64_t nw = (int64_t)gridDim.x * WARPS_PER_BLOCK, nl = (int64_t)*count;
    for (int64_t t = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5); t < nl; t += nw) {
        const int64_t i = list[t];
        CigSrc cs; cs.packed = 0; cs.op = P.B.cig_op + P.B.cig_off[i]; cs.len = P.B.cig_len + P.B.cig_off[i];
        cs.n = (int)(P.B.cig_off[i + 1] - P.B.cig_off[i]);
        int nm, len; int64_t off; bool none;
        nmmd_to_pool<false>(P.G, P.B.seq + P.B.seq_off[i], 0, cs, P.B.chrom[i], P.B.pos[i], lane, I.cursor, I.cap, I.pool, -1, nm, len, off, none);
        uint32_t st = P.W.status[i] | TC_ST_NMMD_DEVICE;
        if (none) { st |= TC_ST_NMMD_NONE; nm = 0; }
        if (lane == 0) { P.W.md_init_off[i] = off < 0 ? 0 : off; P.W.md_init_len[i] = off < 0 ? 0 : len; nm_init[i] = nm; P.W.status[i] = st; }
    }
}

__global__ void k_measure(Params P) { int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < P.B.n) measure_read(P, i); }
__global__ void __launch_bounds__(128, 8) k_plan(Params P) { int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < P.B.n) plan_read(P, i); }
// junction objects for every read; reads that need cleanNoncanonical are appended to `list`
__global__ void __launch_bounds__(128, 8) k_jn_init(Params P, int32_t* list, unsigned* count) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P.B.n && junction_init_read(P, i)) list[atomicAdd(count, 1u)] = (int32_t)i;
}
// ---------------------------------------------------------------------------------- junction rescue, one warp per listed read
// cleanNoncanonical (TC.py:1186-1231) with the control flow of junction_fix_read (tc_plan.cuh) kept
// warp-uniform; what is a loop over a list there - the copies of seg_insert / seg_delete / cig_edit,
// locate_exon, the length sums, the nearest-site searches - runs across the lanes here.

__device__ __forceinline__ int64_t w_seg_total(const u64* s, int n, int lane) {
    int t = 0;
    for (int j = lane; j < n; j += 32) t += (int)seg_len(s[j]);
    return (int64_t)__reduce_add_sync(FULL, t);
}
__device__ __forceinline__ int64_t w_cig_qlen(const u64* c, int n, int lane) {                    // tr.py:591-605
    int t = 0;
    for (int k = lane; k < n; k += 32) { const int op = cg_op(c[k]); if (op == 'M' || op == 'S' || op == 'I') t += cg_len(c[k]); }
    return (int64_t)__reduce_add_sync(FULL, t);
}
// group_CIGAR_by_exon_and_intron (TC.py:1379-1412): exon t = the ops between the (t-1)-th and the t-th N
__device__ bool w_locate_exon(const u64* cig, int n, int t, ExonLoc& e, int& nN, int lane) {
    const unsigned lt = (1u << lane) - 1u;
    int ncount = 0, qcur = 0; int lo = 0, hi = -1; int q0 = 0, q1 = 0;
    for (int k0 = 0; k0 < n; k0 += 32) {
        const int k = k0 + lane; int op = 0, len = 0;
        if (k < n) { const u64 c = cig[k]; op = cg_op(c); len = cg_len(c); }
        const bool isN = op == 'N';
        int qt; const int qs = qcur + warp_excl_scan((op == 'M' || op == 'S' || op == 'I') ? len : 0, lane, qt);
        const unsigned mN = __ballot_sync(FULL, isN);
        const int nb = ncount + __popc(mN & lt);
        const unsigned mHi = __ballot_sync(FULL, isN && nb == t), mLo = __ballot_sync(FULL, isN && nb == t - 1);
        if (mHi) { const int sl = __ffs(mHi) - 1; hi = k0 + sl; q1 = __shfl_sync(FULL, qs, sl); }
        if (mLo) { const int sl = __ffs(mLo) - 1; lo = k0 + sl + 1; q0 = __shfl_sync(FULL, qs, sl); }
        ncount += __popc(mN); qcur += qt;
    }
    nN = ncount;
    if (t > ncount || t < 0) return false;
    e.lo = lo; e.q0 = q0;
    if (hi >= 0) { e.hi = hi; e.q1 = q1; e.nidx = hi; } else { e.hi = n; e.q1 = qcur; e.nidx = -1; }
    return true;
}
__device__ int w_seg_insert(const u64* s, int n, u64* d, int64_t qpos, u64 ns, int lane) {
    if (seg_len(ns) == 0) { for (int j = lane; j < n; j += 32) d[j] = s[j]; return n; }
    bool found = false; int jstar = 0, shift = 0; int64_t qcur = 0;
    for (int j0 = 0; j0 < n; j0 += 32) {
        const int j = j0 + lane; const bool valid = j < n;
        const u64 sg = valid ? s[j] : 0ull; const int len = valid ? (int)seg_len(sg) : 0;
        int tot; const int64_t qs = qcur + warp_excl_scan(len, lane, tot);
        int64_t q_at = 0;
        if (!found) {
            const unsigned mC = __ballot_sync(FULL, valid && qpos < qs + len);
            if (mC) {
                const int sl = __ffs(mC) - 1; jstar = j0 + sl;
                q_at = ((int64_t)__shfl_sync(FULL, (int)(qs >> 32), sl) << 32) | (uint32_t)__shfl_sync(FULL, (int)(uint32_t)qs, sl);
                found = true; shift = qpos > q_at ? 2 : 1;
            }
        }
        if (valid) {
            if (!found || j < jstar) d[j] = sg;
            else if (j > jstar) d[j + shift] = sg;
            else if (shift == 2) {
                const int64_t l1 = qpos - qs;
                d[j] = seg_make(seg_kind(sg), seg_src(sg), l1); d[j + 1] = ns; d[j + 2] = seg_make(seg_kind(sg), seg_src(sg) + l1, len - l1);
            } else { d[j] = ns; d[j + 1] = sg; }
        }
        qcur += tot;
    }
    if (!found) { if (lane == 0) d[n] = ns; return n + 1; }
    return n + shift;
}
__device__ int w_seg_delete(const u64* s, int n, u64* d, int64_t qa, int64_t qb, int lane) {
    if (qa >= qb) { for (int j = lane; j < n; j += 32) d[j] = s[j]; return n; }
    int m = 0; int64_t qcur = 0;
    for (int j0 = 0; j0 < n; j0 += 32) {
        const int j = j0 + lane; const bool valid = j < n;
        const u64 sg = valid ? s[j] : 0ull; const int len = valid ? (int)seg_len(sg) : 0;
        int tot; const int64_t a = qcur + warp_excl_scan(len, lane, tot), b = a + len;
        const bool keep = b <= qa || a >= qb;
        const int cnt = !valid ? 0 : keep ? 1 : (a < qa ? 1 : 0) + (b > qb ? 1 : 0);
        int ctot; int p = m + warp_excl_scan(cnt, lane, ctot);
        if (valid) {
            if (keep) d[p] = sg;
            else {
                if (a < qa) d[p++] = seg_make(seg_kind(sg), seg_src(sg), qa - a);
                if (b > qb) d[p] = seg_make(seg_kind(sg), seg_src(sg) + (qb - a), b - qb);
            }
        }
        m += ctot; qcur += tot;
    }
    return m;
}
// editExonCIGAR (TC.py:1515-1555) on ops [lo,hi) plus the intron length change: the few ops the edit
// touches are found by a (uniform) scalar walk, the list is copied across the lanes
__device__ int w_cig_edit(const u64* src, int n, u64* dst, int lo, int hi, bool at_end, int nb, int intron_idx, int intron_delta,
                          bool& degenerate, int lane) {
    int sa = lo, sb = hi, km = -1, newlen = 0, ins_front = 0, ins_back = 0;
    if (nb >= 0) {
        if (hi - lo == 0) return -1;
        if (at_end) { const u64 last = src[hi - 1]; if (cg_op(last) == 'M') { km = hi - 1; newlen = cg_len(last) + nb; } else ins_back = 1; }
        else { const u64 first = src[lo]; if (cg_op(first) == 'M') { km = lo; newlen = cg_len(first) + nb; } else ins_front = 1; }
    } else {
        int64_t need = -(int64_t)nb;
        if (at_end) {
            int k = hi - 1; int32_t keep_len = -1;
            for (; k >= lo; k--) {
                const int64_t rem = (int64_t)cg_len(src[k]) - need;
                if (rem > 0) { keep_len = (int32_t)rem; break; }
                if (rem == 0) { k--; break; }
                need = -rem;
            }
            sb = k + 1; if (keep_len >= 0) { km = k; newlen = keep_len; }
        } else {
            int k = lo; int32_t keep_len = -1;
            for (; k < hi; k++) {
                const int64_t rem = (int64_t)cg_len(src[k]) - need;
                if (rem > 0) { keep_len = (int32_t)rem; break; }
                if (rem == 0) { k++; break; }
                need = -rem;
            }
            sa = k; if (keep_len >= 0) { km = k; newlen = keep_len; }
        }
        if (sb < sa) sb = sa;
    }
    const int nsurv = sb - sa; bool deg = false;
    for (int k = lane; k < n; k += 32) {
        u64 c = src[k];
        if (k < lo || k >= hi) {
            if (k == intron_idx) { const int32_t L = cg_len(c) + intron_delta; if (L <= 0) deg = true; c = cg_make('N', L); }
            dst[k < lo ? k : lo + ins_front + nsurv + ins_back + (k - hi)] = c;
        } else if (k >= sa && k < sb) {
            if (k == km) c = cg_make(cg_op(c), newlen);
            dst[lo + ins_front + (k - sa)] = c;
        }
    }
    if (lane == 0) { if (ins_front) dst[lo] = cg_make('M', nb); if (ins_back) dst[lo + nsurv] = cg_make('M', nb); }
    if (__any_sync(FULL, deg)) degenerate = true;
    return lo + ins_front + nsurv + ins_back + (n - hi);
}
// sj_nearest (tc_plan.cuh) with a 32-way search: three probes instead of fifteen for ~17,000 sites
__device__ bool w_sj_nearest(const u64* tab, const int32_t* rng, int32_t n_prefix, u64 prefix, int64_t q, bool tie_right, int& dist, int lane) {
    if (prefix >= (u64)n_prefix) return false;
    int lo = rng[prefix], hi = rng[prefix + 1];
    if (lo == hi) return false;
    const int lo0 = lo, hi0 = hi;
    const u64 key = (prefix << 32) | (u64)(uint32_t)q;
    while (hi - lo > 32) {
        const int step = (hi - lo + 31) >> 5; const int idx = lo + lane * step;
        const bool less = idx < hi && tab[idx] < key;
        const int c = __popc(__ballot_sync(FULL, less));             // the probes below the key are the first c lanes
        if (c == 0) { hi = lo; break; }
        const int nlo = lo + (c - 1) * step + 1, nhi = lo + c * step;
        lo = nlo; hi = nhi < hi ? nhi : hi;
    }
    { const int idx = lo + lane; const bool less = idx < hi && tab[idx] < key; lo += __popc(__ballot_sync(FULL, less)); }
    const int i = lo; int64_t best;                                    // lower bound of the key inside [lo0, hi0)
    if (i < hi0) {
        best = (int64_t)(tab[i] & 0xffffffffull);
        if (i > lo0) {
            const int64_t left = (int64_t)(tab[i - 1] & 0xffffffffull);
            const int64_t dl = q - left, dr = best - q;
            if (dl < dr || (dl == dr && !tie_right)) best = left;
        }
    } else best = (int64_t)(tab[i - 1] & 0xffffffffull);
    dist = (int)(best - q);
    return true;
}
// fix_one_side_of_junction (TC.py:1415-1479); see fix_side in tc_plan.cuh
__device__ bool w_fix_side(const Params& P, int chrom, int k, int side, int d, int32_t* bound_slot, int32_t& bound, JnState& S,
                           u64* cig_dst, u64* seg_dst, bool& degenerate, int lane) {
    if (d == 0) return true;
    ExonLoc e; int nN; const int t = side == 0 ? k : k + 1;
    if (!w_locate_exon(S.cig, S.nc, t, e, nN, lane)) return false;            // exonSeqs[targetExon] IndexError
    const GenomeDev& G = P.G; const int64_t goff = G.chrom_off[chrom] - 1;
    int ns2; int64_t newlen = S.seqlen;
    if (side == 0) {
        if (d > 0) {                                                  // case 1: extend exon end from the genome
            const int32_t b = bound;
            if (b < 1) return false;
            const int gl = fetch_len(G, chrom, b, (int64_t)b + d - 1);
            ns2 = w_seg_insert(S.seg, S.ns, seg_dst, e.q1, seg_make(1, goff + b, gl), lane); newlen += gl;
        } else {                                                      // case 3: exon[0:d]
            int64_t m = e.q1 - e.q0; if (m > -d) m = -d;
            ns2 = w_seg_delete(S.seg, S.ns, seg_dst, e.q1 - m, e.q1, lane); newlen -= m;
        }
        bound += d; if (lane == 0) *bound_slot = bound;
        if (k >= nN) return false;                                    // intronCIGARs[jn_number] IndexError
        const int ncn = w_cig_edit(S.cig, S.nc, cig_dst, e.lo, e.hi, true, d, e.nidx, -d, degenerate, lane);
        if (ncn < 0) return false;
        S.nc = ncn;
    } else {
        if (d < 0) {                                                  // case 2: prepend to the next exon
            const int64_t start = (int64_t)bound + 1 + d;
            if (start < 1) return false;
            const int gl = fetch_len(G, chrom, start, (int64_t)bound);
            ns2 = w_seg_insert(S.seg, S.ns, seg_dst, e.q0, seg_make(1, goff + start, gl), lane); newlen += gl;
        } else {                                                      // case 4: exon[d:]
            int64_t m = e.q1 - e.q0; if (m > d) m = d;
            ns2 = w_seg_delete(S.seg, S.ns, seg_dst, e.q0, e.q0 + m, lane); newlen -= m;
        }
        bound += d; if (lane == 0) *bound_slot = bound;
        ExonLoc ek; ek.nidx = -1; int nn2; w_locate_exon(S.cig, S.nc, k, ek, nn2, lane);   // the intron before exon k+1 is the k-th N op
        const int ncn = w_cig_edit(S.cig, S.nc, cig_dst, e.lo, e.hi, false, -d, ek.nidx, d, degenerate, lane);
        if (ncn < 0) return false;
        S.nc = ncn;
    }
    __syncwarp();                                                     // the new lists are read by other lanes from here on
    S.cig = cig_dst; S.seg = seg_dst; S.ns = ns2; S.seqlen = newlen;
    return newlen == w_cig_qlen(S.cig, S.nc, lane);                   // TC.py:1476-1477
}

__device__ void junction_fix_warp(const Params& P, int64_t i, int lane) {
    const BatchDev& B = P.B; const WorkDev& W = P.W; const tc_options& O = P.O;
    uint32_t st = W.status[i];
    const int chrom = B.chrom[i]; const int strand = (B.flag[i] == 16) ? 1 : 0;
    const int nj = W.n_jn[i];
    WsView v = ws_view(W, i); int32_t* J = v.jn;
    int32_t* C = W.counters + 10 * i;
    bool canon = false, annot = (st & TC_ST_ALL_ANNOT) != 0;
    st &= ~(uint32_t)(TC_ST_CANONICAL | TC_ST_ALL_ANNOT);
    int nte = W.n_te[i]; int cJ = 0, uJ = 0;
    const int64_t Lq = B.seq_off[i + 1] - B.seq_off[i];
    int cb = W.cig_buf[i], sb = W.seg_buf[i]; int nc = W.n_cig_cur[i], ns = W.n_seg_cur[i];
    if (cb < 0) {                                                     // still the input: bring it into the workspace
        const int64_t c0 = B.cig_off[i]; nc = (int)(B.cig_off[i + 1] - c0);
        for (int k = lane; k < nc; k += 32) v.cig[0][k] = cg_make(B.cig_op[c0 + k], B.cig_len[c0 + k]);
        ns = 0; if (Lq > 0) { if (lane == 0) v.seg[0][0] = seg_make(0, 0, Lq); ns = 1; }
        cb = 0; sb = 0;
        __syncwarp();
    }
    int64_t seqlen = w_seg_total(v.seg[sb], ns, lane);
    bool whole_fail = false, degenerate = false, any_ok = false;
    const bool tie_right = O.sj_tie_right != 0;
    for (int k = 0; k < nj && !whole_fail; k++) {
        int32_t* r = J + k * JN_STRIDE;
        const int code0 = r[4];
        if (!(code0 == 0 || code0 == 20)) continue;
        const int32_t id_a = r[2], id_b = r[3];
        const int ds = strand == 0 ? 0 : 1, as = 1 - ds;             // sj.py:29-41
        int32_t bnd[2] = {r[0], r[1]};
        int dd[2] = {0, 0}; bool ok_t = true;                         // dd[0]: distance to the nearest donor, dd[1]: acceptor
        const u64 pre = O.sj_ignore_strand ? (u64)chrom : (((u64)chrom << 1) | (u64)strand);
#pragma unroll 1
        for (int w = 0; w < 2 && ok_t; w++) {
            const u64* tab = O.sj_ignore_strand ? (w == 0 ? P.S.don_u : P.S.acc_u) : (w == 0 ? P.S.don_s : P.S.acc_s);
            const int32_t* rng = O.sj_ignore_strand ? (w == 0 ? P.S.rng_don_u : P.S.rng_acc_u) : (w == 0 ? P.S.rng_don_s : P.S.rng_acc_s);
            ok_t = w_sj_nearest(tab, rng, P.S.n_prefix, pre, (int64_t)bnd[w == 0 ? ds : as] - 1, tie_right, dd[w], lane);
        }
        const int d_don = dd[0], d_acc = dd[1];
        if (!ok_t) { whole_fail = true; break; }                      // Q23: raises outside the inner try
        const int ad = d_don < 0 ? -d_don : d_don, aa = d_acc < 0 ? -d_acc : d_acc;
        const int dist = ((d_don < 0) != (d_acc < 0)) ? ad + aa : (ad > aa ? ad - aa : aa - ad);   // TC.py:1371-1375
        if (dist > O.max_sj_offset || ad > 2 * O.max_sj_offset || aa > 2 * O.max_sj_offset) {
            uJ++; if (lane == 0) { int n2 = nte; te_put(v.te, n2, 3, id_a, id_b, dist, 1, 3); } nte++; continue;
        }
        const int t1 = (cb + 1) % 3, t2 = (cb + 2) % 3, u1 = (sb + 1) % 3, u2 = (sb + 2) % 3;   // the scratch buffers that are not the committed one
        JnState S; S.cig = v.cig[cb]; S.nc = nc; S.seg = v.seg[sb]; S.ns = ns; S.seqlen = seqlen;
        int used_c = cb, used_s = sb;
        bool ok = true;
#pragma unroll 1
        for (int w = 0; w < 2 && ok; w++) {                           // donor side, then acceptor side
            const int side = w == 0 ? ds : as, d = w == 0 ? d_don : d_acc;
            if (d == 0) continue;
            const int tc = used_c == cb ? t1 : t2, ts = used_s == sb ? u1 : u2;
            ok = w_fix_side(P, chrom, k, side, d, &r[side], bnd[side], S, v.cig[tc], v.seg[ts], degenerate, lane); used_c = tc; used_s = ts;
        }
        int code = code0;
        if (ok) {                                                     // update_post_ncsj_correction (TC.py:1279-1294)
            if (!junction_code(P, chrom, strand, bnd[0], bnd[1], bnd[0], bnd[1], code)) ok = false;
            if (lane == 0) { r[2] = bnd[0]; r[3] = bnd[1]; if (ok) r[4] = code; }
        }
        __syncwarp();
        if (ok) {
            cb = used_c; sb = used_s; nc = S.nc; ns = S.ns; seqlen = S.seqlen; any_ok = true;
            bool cn = true, an = true;
            for (int m = lane; m < nj; m += 32) {                     // tags + flags from the shared junction objects
                int32_t* rr = J + m * JN_STRIDE;
                const int c4 = rr[4];
                rr[5] = rr[0]; rr[6] = rr[1]; rr[7] = c4;
                if (c4 == 0 || c4 == 20) cn = false;
                if (c4 < 20) an = false;
            }
            canon = __all_sync(FULL, cn); annot = __all_sync(FULL, an);
            cJ++; if (lane == 0) { int n2 = nte; te_put(v.te, n2, 3, id_a, id_b, dist, 0, 0); } nte++;
        } else {
            uJ++; if (lane == 0) { int n2 = nte; te_put(v.te, n2, 3, id_a, id_b, dist, 1, 4); } nte++;
        }
        __syncwarp();
    }
    if (whole_fail) {                                                 // Q13/Q23: original read, TE dropped
        st |= TC_ST_FALLBACK; st &= ~(uint32_t)TC_ST_CIGAR_REWRITTEN;
        bool cn = true, an = true;
        for (int m = lane; m < nj; m += 32) { const int c = J[m * JN_STRIDE + 4]; if (c == 0 || c == 20) cn = false; if (c < 20) an = false; }
        canon = __all_sync(FULL, cn); annot = __all_sync(FULL, an);   // junction flags go back to the init values
        if (lane == 0) {
            W.cig_buf[i] = -1; W.seg_buf[i] = -1; W.n_te[i] = 0;
            W.te_cnt[3 * i] = W.te_cnt[3 * i + 1] = W.te_cnt[3 * i + 2] = 0;
            W.refresh[i] = 0; W.nm_extra[i] = 0;
        }
    } else {
        if (any_ok) { st |= TC_ST_CIGAR_REWRITTEN | TC_ST_JM_DEVICE | TC_ST_JI_DEVICE; if (degenerate) st |= TC_ST_ERR_INTRON; }
        if (lane == 0) {
            C[TC_C_COR_SJ] = cJ; C[TC_C_UNC_SJ] = uJ;
            W.n_te[i] = nte;
            if (any_ok) {
                W.cig_buf[i] = (int8_t)cb; W.seg_buf[i] = (int8_t)sb; W.n_cig_cur[i] = nc; W.n_seg_cur[i] = ns; W.seq_len_out[i] = (int32_t)seqlen;
                W.refresh[i] = 1; W.nm_extra[i] = 0;
            }
        }
    }
    if (canon) st |= TC_ST_CANONICAL;
    if (annot) st |= TC_ST_ALL_ANNOT;
    if (lane == 0) W.status[i] = st;
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, 4) k_jn_fix(Params P, const int32_t* list, const unsigned* count) {
    const int lane = threadIdx.x & 31;
    const int64_t nw = (int64_t)gridDim.x * WARPS_PER_BLOCK, nl = (int64_t)*count;
    for (int64_t t = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5); t < nl; t += nw) junction_fix_warp(P, list[t], lane);
}
__global__ void k_sizes(Params P) { int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < P.B.n) sizes_read(P, i); }
__global__ void k_dry(Params P) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P.B.n) { dry_read(P, i); dry_sizes(P, i); }
}

__global__ void k_pack(Params P) { int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < P.B.n) pack_read(P, i); }

// small_copies_read as one warp per read: TE rows into stage order by a ballot partition, CIGAR and
// jM / jI copies, all from the read's 128-byte plan row.  A light kernel of its own (no shared memory,
// few registers, 48 warps of an SM resident): these phases are latency, not bytes, and inside
// k_emit they held its 30 resident warps up for a fifth of its time.
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, 6) k_small(Params P, OutDev O, int dry) {
    const int lane = threadIdx.x & 31; const unsigned lt = (1u << lane) - 1u;
    const BatchDev& B = P.B; const WorkDev& W = P.W;
    const int64_t nw = (int64_t)gridDim.x * WARPS_PER_BLOCK;
    for (int64_t i = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5); i < B.n; i += nw) {
        const int pw = W.plan[i].w[lane];
        if (((uint32_t)PF(EP_STATUS) & TC_ST_CLASS_MASK) != 0) continue;
        const int pflags = PF(EP_FLAGS), cig_cap = PF(EP_CIGCAP), seg_cap = PF(EP_SEGCAP), te_cap = PF(EP_TECAP);
        const int ncig = PF(EP_NCIG), nte = PF(EP_NTE), nj = PF(EP_NJN);
        const int te_mm = PF(EP_TEMM), te_ins = PF(EP_TEINS), te_del = PF(EP_TEDEL);
        const int64_t ws_off = PF64(EP_WS_LO), o_cig = PF64(EP_OCIG_LO), o_te = PF64(EP_OTE_LO), o_jn = PF64(EP_OJN_LO), cig0 = PF64(EP_CIG0_LO);
        const u64* const ws = W.ws + ws_off;
        const tc_te* const ws_te = reinterpret_cast<const tc_te*>(ws + 3ll * seg_cap + 3ll * cig_cap);
        tc_te* dst = O.te + o_te;
        if (dry) { for (int k = lane; k < nte; k += 32) dst[k] = ws_te[k]; continue; }
        int base0 = 0, base1 = te_mm, base2 = te_mm + te_ins, base3 = te_mm + te_ins + te_del;
        for (int c = 0; c < nte; c += 32) {
            const int k = c + lane; const bool valid = k < nte;
            tc_te r; int t = -1;
            if (valid) { r = ws_te[k]; t = r.type; }
            const unsigned m0 = __ballot_sync(FULL, t == 0), m1 = __ballot_sync(FULL, t == 1), m2 = __ballot_sync(FULL, t == 2), m3 = __ballot_sync(FULL, t == 3);
            if (valid) {
                const int slot = t == 0 ? base0 + __popc(m0 & lt) : t == 1 ? base1 + __popc(m1 & lt) : t == 2 ? base2 + __popc(m2 & lt) : base3 + __popc(m3 & lt);
                dst[slot] = r;
            }
            base0 += __popc(m0); base1 += __popc(m1); base2 += __popc(m2); base3 += __popc(m3);
        }
        const int cb = ((pflags >> 2) & 3) - 1;
        CigSrc cs; cs.n = ncig;
        if (cb < 0) { cs.packed = 0; cs.op = B.cig_op + cig0; cs.len = B.cig_len + cig0; }
        else { cs.packed = ws + 3ll * seg_cap + (int64_t)cb * cig_cap; cs.op = 0; cs.len = 0; }
        uint8_t* dop = O.cig_op + o_cig; int32_t* dl = O.cig_len + o_cig;
        for (int k = lane; k < ncig; k += 32) { int op; int32_t l; cig_get(cs, k, op, l); dop[k] = (uint8_t)op; dl[k] = l; }
        if (nj > 0) {
            int8_t* djm = O.jm + o_jn; int32_t* dji = O.ji + 2 * o_jn;
            const int32_t* jn = reinterpret_cast<const int32_t*>(ws + 3ll * seg_cap + 3ll * cig_cap + 2ll * te_cap);
            for (int k = lane; k < nj; k += 32) { const int32_t* r = jn + k * JN_STRIDE; djm[k] = (int8_t)r[7]; dji[2 * k] = r[5]; dji[2 * k + 1] = r[6]; }
        }
    }
}

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Output stage: one warp per read.  [seq_lo, seq_hi) = 32-bit words of B.seq that may be read
// (the batch's SEQ buffer on the device, padded).
//   1. plan row (128 B, one word per lane) -> everything about the read; the next read's row,
//      segment list and SEQ lines are prefetched into L2 meanwhile
//   2. TE rows back into stage order (ballot partition), CIGAR and jM/jI copies
//   3. SEQ: built in shared memory - (A) segment list -> slice table, (B) 16-byte chunks gathered
//      from the input SEQ with the shift in force, (C) fix-ups after slice starts, (D) genome
//      slices byte-exact, (E) one coalesced 16-byte-vector store
//   4. NM/MD (getNMandMDFlags): tier 1 stages the 2-bit genome words of every M run in shared
//      memory with cp.async (all loads in flight at once) and verifies 16 bases per lane against
//      the staged SEQ; a read that verifies clean needs no token (MD = total match length).
//      Anything else (a difference, a D op, a literal run, the contig end) takes warp_nmmd.
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, 4) k_emit(Params P, OutDev O, int dry, int64_t seq_lo, int64_t seq_hi) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t* lut = reinterpret_cast<uint32_t*>(smem_raw);             // 4 bases (8 bits of the 2-bit plane) -> 4 ASCII bytes
    EmitSmem& SM = reinterpret_cast<EmitSmem*>(smem_raw + 1024)[threadIdx.x >> 5];
    for (int t = threadIdx.x; t < 256; t += blockDim.x) {
        const uint32_t acgt = 0x54474341u;
        lut[t] = ((acgt >> (8 * (t & 3))) & 255u) | (((acgt >> (8 * ((t >> 2) & 3))) & 255u) << 8) |
                 (((acgt >> (8 * ((t >> 4) & 3))) & 255u) << 16) | (((acgt >> (8 * ((t >> 6) & 3))) & 255u) << 24);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t nw = (int64_t)gridDim.x * WARPS_PER_BLOCK;
    const BatchDev& B = P.B; const WorkDev& W = P.W;
    const uint32_t* S32 = reinterpret_cast<const uint32_t*>(B.seq);
    const int dbg = P.O.reserved;     // profiling aid: bit0 skip small copies, bit1 skip SEQ, bit2 skip NM/MD (results invalid)
    int64_t i = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    int pw = i < B.n ? W.plan[i].w[lane] : 0;
    for (; i < B.n; i += nw) {
        uint32_t st = (uint32_t)PF(EP_STATUS);
        const int pflags = PF(EP_FLAGS), cig_cap = PF(EP_CIGCAP), seg_cap = PF(EP_SEGCAP), nm_extra = PF(EP_NMEXTRA);
        const int ncig = PF(EP_NCIG), out_len = PF(EP_OUTLEN), ns = PF(EP_NSEG);
        const int chrom = PF(EP_CHROM); const int64_t pos = PF(EP_POS);
        const int64_t ws_off = PF64(EP_WS_LO), o_seq = PF64(EP_OSEQ_LO);
        const int64_t in_base = PF64(EP_INB_LO), cig0 = PF64(EP_CIG0_LO); const int Lq = PF(EP_LQ);
        // next read: fetch its row now (consumed at the end of this iteration), and pull its segment
        // list and SEQ lines towards L2
        const int64_t inext = i + nw;
        const int pw_next = inext < B.n ? W.plan[inext].w[lane] : 0;
        if ((st & TC_ST_CLASS_MASK) != 0) { if (lane == 0) { O.md_off[i] = 0; O.md_len[i] = 0; O.nm[i] = 0; } pw = pw_next; continue; }
        u64* const ws = W.ws + ws_off;
        const int cb = ((pflags >> 2) & 3) - 1, sb = ((pflags >> 4) & 3) - 1;
        if (dry) { if (lane == 0) { O.md_off[i] = 0; O.md_len[i] = 0; O.nm[i] = 0; } pw = pw_next; continue; }
        // (TE rows, the CIGAR and the jM / jI snapshot were copied by k_pack)
        CigSrc cs; cs.n = ncig;
        if (cb < 0) { cs.packed = 0; cs.op = B.cig_op + cig0; cs.len = B.cig_len + cig0; }
        else { cs.packed = ws + 3ll * seg_cap + (int64_t)cb * cig_cap; cs.op = 0; cs.len = 0; }
        if (dbg & 2) { pw = pw_next; continue; }
        // ---- SEQ
        const uint8_t* in_seq = B.seq + in_base;
        uint8_t* out_seq = O.seq + o_seq;
        const int64_t goff = P.G.chrom_off[chrom] - 1, clen = P.G.chrom_len[chrom];
        const int64_t gbase = goff + pos;                                // plane index of the read's first reference base
        const u64* const ws_seg = ws + (int64_t)(sb < 0 ? 0 : sb) * seg_cap;
        bool fast = out_len <= EMIT_CAP;
        int nbp = 0, ngi = 0;
        if (fast) {
            // (A) segment list -> starts of input slices (output offset, shift) + genome slices
            if (sb < 0) { if (lane == 0) { SM.bp_pos[0] = 0; SM.bp_shift[0] = 0; } nbp = 1; }
            else {
                const u64* segs = ws_seg; int obase = 0; bool bad = false;
                int last_shift = 0x7fffffff;                           // shift of the latest input slice (none yet)
                for (int s0 = 0; s0 < ns; s0 += 32) {
                    const int s = s0 + lane; const bool valid = s < ns;
                    const u64 sg = valid ? segs[s] : 0ull;
                    const int len = valid ? (int)seg_len(sg) : 0; const int64_t src = seg_src(sg); const int kind = seg_kind(sg);
                    int total; const int ostart = obase + warp_excl_scan(len, lane, total);
                    const bool isR = valid && kind == 0, isG = valid && kind == 1;
                    const unsigned mR0 = __ballot_sync(FULL, isR), mG = __ballot_sync(FULL, isG);
                    // an input slice that continues the previous one's shift (only genome slices lie between them:
                    // corrected mismatches) is not a new slice - the genome bytes are patched in afterwards
                    const int64_t shift = src - ostart;
                    const bool shift_ok = shift >= -32768 && shift <= 32767;
                    const int sh32 = shift_ok ? (int)shift : 0x7ffffffe;
                    const unsigned prevR = mR0 & lt;
                    const int prev_shift = __shfl_sync(FULL, sh32, prevR ? 31 - __clz(prevR) : lane);
                    const bool isBP = isR && (!shift_ok || sh32 != (prevR ? prev_shift : last_shift));
                    const unsigned mR = __ballot_sync(FULL, isBP);
                    if (mR0) last_shift = __shfl_sync(FULL, sh32, 31 - __clz(mR0));
                    if (isBP) {
                        const int idx = nbp + __popc(mR & lt);
                        if (idx < EMIT_MAXBP && shift_ok) { SM.bp_pos[idx] = (uint16_t)ostart; SM.bp_shift[idx] = (int16_t)shift; }
                        else bad = true;
                    }
                    if (isG) {
                        const int idx = ngi + __popc(mG & lt); const int64_t rel = src - gbase;
                        if (idx < EMIT_MAXGI && len <= 65535 && rel >= -2147483647ll && rel <= 2147483647ll) {
                            SM.gi_out[idx] = (uint16_t)ostart; SM.gi_len[idx] = (uint16_t)len; SM.gi_src[idx] = (int32_t)rel;
                        } else bad = true;
                    }
                    nbp += __popc(mR); ngi += __popc(mG); obase += total;
                }
                if (__any_sync(FULL, bad)) fast = false;
            }
        }
        // the next read's segment list and SEQ lines -> L2 while this read is being built
        if (inext < B.n) {
            const int64_t nws = ((int64_t)__shfl_sync(FULL, pw_next, EP_WS_HI) << 32) | (uint32_t)__shfl_sync(FULL, pw_next, EP_WS_LO);
            const int64_t nin = ((int64_t)__shfl_sync(FULL, pw_next, EP_INB_HI) << 32) | (uint32_t)__shfl_sync(FULL, pw_next, EP_INB_LO);
            const int nlq = __shfl_sync(FULL, pw_next, EP_LQ);
            if (lane < 2) prefetch_l2(W.ws + nws + lane * 16);
            else if ((lane - 2) * 128 < nlq) prefetch_l2(B.seq + nin + (lane - 2) * 128);
        }
        if (fast) {
            __syncwarp();
            bool read_exc = false;
            for (int k = lane; k < ngi; k += 32) { const int64_t src = gbase + SM.gi_src[k]; read_exc |= genome_blk_has_exc(P.G, src) | genome_blk_has_exc(P.G, src + SM.gi_len[k] - 1) | (SM.gi_len[k] > 256); }
            read_exc = __any_sync(FULL, read_exc);
            const int nchunk = (out_len + 15) >> 4;
            // each slice writes its index into the chunks that start inside it (slice k is in force from the
            // first chunk boundary at or after its start up to the next slice's first chunk)
            for (int k = lane; k < nbp; k += 32) {
                const int c0 = k == 0 ? 0 : ((int)SM.bp_pos[k] + 15) >> 4;
                const int c1 = k + 1 < nbp ? ((int)SM.bp_pos[k + 1] + 15) >> 4 : nchunk;
                for (int c = c0; c < c1; c++) SM.chunk_bp[c] = (uint8_t)k;
            }
            __syncwarp();
            // (B) bulk copy, 16 output bytes per lane, with the shift in force at the chunk start
            // (32-bit offsets from the read's first word; the clamps keep the gather inside the SEQ buffer)
            const uint32_t* const rp = S32 + (in_base >> 2); const int rsub = (int)(in_base & 3);
            const int64_t wlo64 = seq_lo - (in_base >> 2), whi64 = seq_hi - 5 - (in_base >> 2);
            const int wlo = wlo64 < -0x40000000ll ? -0x40000000 : (int)wlo64, whi = whi64 > 0x40000000ll ? 0x40000000 : (int)whi64;
            auto src_of = [&](int c, int& sh) {                       // first source word of output chunk c, and its byte shift
                const int k = SM.chunk_bp[c];
                const int a = rsub + (c << 4) + (nbp ? (int)SM.bp_shift[k] : 0);
                int wi = a >> 2; sh = (a & 3) * 8;
                wi = wi < wlo ? wlo : wi; wi = wi > whi ? whi : wi;
                return rp + wi;
            };
            for (int c = lane; c < nchunk; c += 64) {                  // two chunks per trip: ten loads in flight per lane
                int sh0, sh1 = 0; const int c1 = c + 32; const bool two = c1 < nchunk;
                const uint32_t* xp = src_of(c, sh0); const uint32_t* yp = two ? src_of(c1, sh1) : xp;
                const uint32_t x0 = xp[0], x1 = xp[1], x2 = xp[2], x3 = xp[3], x4 = xp[4];
                const uint32_t y0 = yp[0], y1 = yp[1], y2 = yp[2], y3 = yp[3], y4 = yp[4];
                uint4 o; o.x = __funnelshift_r(x0, x1, sh0); o.y = __funnelshift_r(x1, x2, sh0); o.z = __funnelshift_r(x2, x3, sh0); o.w = __funnelshift_r(x3, x4, sh0);
                *reinterpret_cast<uint4*>(SM.seq + (c << 4)) = o;
                if (two) {
                    o.x = __funnelshift_r(y0, y1, sh1); o.y = __funnelshift_r(y1, y2, sh1); o.z = __funnelshift_r(y2, y3, sh1); o.w = __funnelshift_r(y3, y4, sh1);
                    *reinterpret_cast<uint4*>(SM.seq + (c1 << 4)) = o;
                }
            }
            __syncwarp();
            // (C) the part of a chunk that follows a slice start inside it: same 5-word gather, byte stores
            for (int k = lane; k < nbp; k += 32) {
                const int b = SM.bp_pos[k];
                if (b & 15) {
                    int end = (b | 15) + 1; if (end > out_len) end = out_len;
                    if (k + 1 < nbp && (int)SM.bp_pos[k + 1] < end) end = SM.bp_pos[k + 1];
                    const int a = rsub + b + (int)SM.bp_shift[k];
                    int wi = a >> 2; const int sh = (a & 3) * 8;
                    wi = wi < wlo ? wlo : wi; wi = wi > whi ? whi : wi;
                    const uint32_t* xp = rp + wi;
                    const uint32_t x0 = xp[0], x1 = xp[1], x2 = xp[2], x3 = xp[3], x4 = xp[4];
                    uint32_t w[4]; w[0] = __funnelshift_r(x0, x1, sh); w[1] = __funnelshift_r(x1, x2, sh); w[2] = __funnelshift_r(x2, x3, sh); w[3] = __funnelshift_r(x3, x4, sh);
                    const int n = end - b;
#pragma unroll
                    for (int j = 0; j < 15; j++) if (j < n) SM.seq[b + j] = (uint8_t)(w[j >> 2] >> (8 * (j & 3)));
                }
            }
            __syncwarp();
            // (D) genome slices (corrected mismatches, deletion fills, junction extensions): byte exact (Q6)
            for (int k = lane; k < ngi; k += 32) {
                const int o = SM.gi_out[k], len = SM.gi_len[k]; const int64_t src = gbase + SM.gi_src[k];
                if (read_exc) { for (int j = 0; j < len; j++) SM.seq[o + j] = genome_byte(P.G, src + j); }
                else { for (int j = 0; j < len; j++) SM.seq[o + j] = genome_byte_fast(P.G, src + j); }
            }
            __syncwarp();
            // (E) one coalesced store of the row (rows are padded to 16 bytes)
            for (int c = lane; c < nchunk; c += 32) reinterpret_cast<uint4*>(out_seq)[c] = *reinterpret_cast<const uint4*>(SM.seq + (c << 4));
        } else if (sb < 0) {
            for (int64_t j = lane; j < Lq; j += 32) out_seq[j] = in_seq[j];
        } else {
            const u64* segs = ws_seg; int64_t o = 0;
            for (int s = 0; s < ns; s++) {
                const u64 sg = segs[s]; const int64_t src = seg_src(sg), len = seg_len(sg);
                if (seg_kind(sg) == 0) { for (int64_t j = lane; j < len; j += 32) out_seq[o + j] = in_seq[src + j]; }
                else { for (int64_t j = lane; j < len; j += 32) out_seq[o + j] = genome_byte(P.G, src + j); }
                o += len;
            }
        }
        // ---- NM / MD
        if ((pflags & 1) && !(dbg & 4)) {
            __syncwarp();                                              // SM tables are dead from here on: their space holds the run list
            int nm = 0, len = 0; int64_t off = 0; bool done = false;
            if (fast && ncig <= 32 * 4) {
                // tier 1: run list (query offset, reference offset, length) of the M ops, lane-parallel
                uint16_t* run_q = SM.run_q; uint16_t* run_ct = SM.run_ct; int32_t* run_g = SM.run_g;
                int nrun = 0, qcur = 0, nD = 0, ibases = 0, dbases = 0, mcur = 0, m_at_last_d = 0, qsum = 0; int64_t gcur = 0; bool ovf = false;
                for (int k0 = 0; k0 < ncig; k0 += 32) {
                    int op = 0; int32_t ct = 0;
                    if (k0 + lane < ncig) cig_get(cs, k0 + lane, op, ct);
                    const int qadv = (op == 'M' || op == 'I' || op == 'S') ? ct : 0;
                    const int gadv = (op == 'M' || op == 'D' || op == 'N' || op == 'H') ? ct : 0;
                    // query offsets and matched-base counts share one scan (16 bits each: both are bounded by
                    // the query length, which the REDUX total below checks against out_len <= EMIT_CAP)
                    qsum += __reduce_add_sync(FULL, qadv);
                    int qmtot, gtot; const int qm = warp_excl_scan((qadv << 16) | (op == 'M' ? (ct & 0xffff) : 0), lane, qmtot);
                    const int gs = warp_excl_scan(gadv, lane, gtot);
                    const int qs = (int)((unsigned)qm >> 16), ms = qm & 0xffff;              // ms: matched bases before this op
                    const int qtot = (int)((unsigned)qmtot >> 16), mtot = qmtot & 0xffff;
                    const bool isM = op == 'M' && ct > 0, isD = op == 'D';
                    const unsigned mM = __ballot_sync(FULL, isM), mD = __ballot_sync(FULL, isD);
                    {   // MD token of a D op: <matches since the previous token>^<bases>
                        const unsigned prev = mD & lt;                  // nearest earlier D in this round, else the carried one
                        const int cand = __shfl_sync(FULL, mcur + ms, prev ? 31 - __clz(prev) : lane);
                        if (isD) {
                            const int idx = nD + __popc(prev);
                            if (idx < EMIT_MAXD) { SM.d_g[idx] = (int32_t)(gcur + gs); SM.d_ct[idx] = ct; SM.d_run[idx] = mcur + ms - (prev ? cand : m_at_last_d); }
                            else ovf = true;
                        }
                        const int lastd = __shfl_sync(FULL, mcur + ms, mD ? 31 - __clz(mD) : 0);
                        if (mD) m_at_last_d = lastd;
                    }
                    nD += __popc(mD);
                    ibases += op == 'I' ? ct : 0; dbases += isD ? ct : 0;
                    if (isM) {
                        const int idx = nrun + __popc(mM & lt);
                        if (idx < EMIT_MAXRUN && gcur + gs < 0x7fff0000ll) { run_q[idx] = (uint16_t)(qcur + qs); run_g[idx] = (int32_t)(gcur + gs); run_ct[idx] = (uint16_t)ct; }
                        else ovf = true;
                        // literal runs (N, IUPAC) under this M run?  (blocks of 256 bases; a run spans at most 17)
                        const int64_t ga = gbase + gcur + gs;
                        const uint32_t b1 = (uint32_t)((ga + ct - 1) >> 8);
                        for (uint32_t b = (uint32_t)(ga >> 8); b <= b1; b++) ovf |= (P.G.excblk[b >> 5] >> (b & 31)) & 1u;
                    }
                    nrun += __popc(mM); qcur += qtot; gcur += gtot; mcur += mtot;
                }
                ibases = __reduce_add_sync(FULL, ibases); dbases = __reduce_add_sync(FULL, dbases);
                const bool careful = pos < 1 || pos + gcur - 1 > clen || gbase < 16;   // (chunks may start up to 15 bases before a run)
                bool tier1 = !careful && qsum == out_len && nD <= EMIT_MAXD && nrun <= EMIT_MAXRUN && !__any_sync(FULL, ovf);
                if (tier1) {
                    // The staged SEQ is verified in aligned 16-byte chunks (one LDS.128 per item; the 2-bit genome
                    // words carry the shift).  Pass 1: chunks that lie wholly inside an M run, as one flat item
                    // space over the runs that have any (lane r keeps run r's item offset / chunk offset / genome
                    // offset; an item finds its run with REDUX.OR + POPC and fetches the two offsets by shuffle).
                    // Pass 2: the partial chunk at the head and at the tail of every run, one lane each, masked.
                    __syncwarp();
                    bool bad = false; const int total_m = mcur;
                    int my_w = 0x7fffffff, my_coff = 0, my_goff = 0, nit = 0;
                    {
                        int cnt = 0, c0 = 0, q0 = 0, g0 = 0;
                        if (lane < nrun) { q0 = run_q[lane]; g0 = run_g[lane]; c0 = (q0 + 15) >> 4; cnt = ((q0 + (int)run_ct[lane]) >> 4) - c0; cnt = cnt < 0 ? 0 : cnt; }
                        const int w = warp_excl_scan(cnt, lane, nit);
                        const unsigned has = __ballot_sync(FULL, cnt > 0);
                        if (cnt > 0) { const int k = __popc(has & lt); SM.rb_w[k] = w; SM.rb_coff[k] = c0 - w; SM.rb_goff[k] = g0 - q0; }
                        __syncwarp();
                        if (lane < __popc(has)) { my_w = SM.rb_w[lane]; my_coff = SM.rb_coff[lane]; my_goff = SM.rb_goff[lane]; }
                    }
                    int rbase = 0;
                    const uint32_t* const gw0p = P.G.g2 + (gbase >> 4); const int gsub = (int)(gbase & 15);
                    auto locate = [&](int t0, int& p, const uint32_t*& gp, int& gsh) {   // item t0 + lane: chunk start, genome words, shift
                        const unsigned word = __reduce_or_sync(FULL, (my_w >= t0 && my_w < t0 + 32) ? 1u << (my_w - t0) : 0u);
                        const int r = rbase + __popc(word & (0xffffffffu >> (31 - lane))) - 1;
                        rbase += __popc(word);
                        const int coff = __shfl_sync(FULL, my_coff, r & 31), goff = __shfl_sync(FULL, my_goff, r & 31);
                        p = (t0 + lane + coff) << 4;                             // chunk start (query offset)
                        const int gb = gsub + p + goff;                          // bases from the read's first genome word to chunk byte 0
                        gp = gw0p + (gb >> 4); gsh = (gb & 15) * 2;
                    };
                    auto differs = [&](int p, uint32_t g0, uint32_t g1, int gsh) {
                        const uint32_t gw = __funnelshift_r(g0, g1, gsh);
                        const uint4 x = *reinterpret_cast<const uint4*>(SM.seq + p);
                        const uint32_t d0 = (x.x & 0xDFDFDFDFu) ^ lut[gw & 255u], d1 = (x.y & 0xDFDFDFDFu) ^ lut[(gw >> 8) & 255u];
                        const uint32_t d2 = (x.z & 0xDFDFDFDFu) ^ lut[(gw >> 16) & 255u], d3 = (x.w & 0xDFDFDFDFu) ^ lut[gw >> 24];
                        return (d0 | d1 | d2 | d3) != 0u;
                    };
                    for (int t0 = 0; t0 < nit; t0 += 64) {                 // two items per trip: four genome loads in flight per lane
                        int pA, pB = 0, shA, shB = 0; const uint32_t* gA; const uint32_t* gB = gw0p;
                        locate(t0, pA, gA, shA);
                        const bool two = t0 + 32 < nit;                         // (uniform)
                        if (two) locate(t0 + 32, pB, gB, shB);
                        const bool vA = t0 + lane < nit, vB = two && t0 + 32 + lane < nit;
                        uint32_t a0 = 0, a1 = 0, b0 = 0, b1 = 0;
                        if (vA) { a0 = gA[0]; a1 = gA[1]; }
                        if (vB) { b0 = gB[0]; b1 = gB[1]; }
                        if (vA) bad |= differs(pA, a0, a1, shA);
                        if (vB) bad |= differs(pB, b0, b1, shB);
                    }
                    for (int j = lane; j < 2 * nrun; j += 32) {
                        const int r = j >> 1, a = run_q[r], b = a + (int)run_ct[r];
                        const bool head = (a & 15) != 0;
                        int s, e;
                        if (!(j & 1)) { s = a; e = (a | 15) + 1; e = e < b ? e : b; if (!head) e = s; }
                        else { s = b & ~15; e = b; if (!(b & 15) || (head && (b >> 4) == (a >> 4))) e = s; }
                        if (e > s) {
                            const int p = s & ~15, l8 = 8 * (s - p), h8 = 8 * (e - p);   // chunk bytes [lo, hi) count
                            const int64_t gb = gbase + run_g[r] + (p - a);
                            const uint32_t* gp = P.G.g2 + (gb >> 4);
                            const uint32_t gw = __funnelshift_r(gp[0], gp[1], (int)(gb & 15) * 2);
                            const uint4 x = *reinterpret_cast<const uint4*>(SM.seq + p);
                            uint32_t d0 = (x.x & 0xDFDFDFDFu) ^ lut[gw & 255u], d1 = (x.y & 0xDFDFDFDFu) ^ lut[(gw >> 8) & 255u];
                            uint32_t d2 = (x.z & 0xDFDFDFDFu) ^ lut[(gw >> 16) & 255u], d3 = (x.w & 0xDFDFDFDFu) ^ lut[gw >> 24];
#define TC_BITS_BELOW(b, w) ((b) >= 32 * (w) + 32 ? ~0u : ((b) <= 32 * (w) ? 0u : ((1u << ((b) - 32 * (w))) - 1u)))
                            d0 &= TC_BITS_BELOW(h8, 0) & ~TC_BITS_BELOW(l8, 0);
                            d1 &= TC_BITS_BELOW(h8, 1) & ~TC_BITS_BELOW(l8, 1);
                            d2 &= TC_BITS_BELOW(h8, 2) & ~TC_BITS_BELOW(l8, 2);
                            d3 &= TC_BITS_BELOW(h8, 3) & ~TC_BITS_BELOW(l8, 3);
#undef TC_BITS_BELOW
                            bad |= (d0 | d1 | d2 | d3) != 0u;
                        }
                    }
                    if (!__any_sync(FULL, bad)) {                      // every aligned base equals the genome
                        nm = ibases + dbases;
                        if (nD == 0) {                                 // MD is one number, at most 5 digits: the read's inline slot
                            len = total_m > 0 ? ndigits((uint32_t)total_m) : 0;
                            off = O.md_inline + 8 * i;
                            if (lane == 0) *reinterpret_cast<unsigned long long*>(O.md_pool + off) = dec_word4((uint32_t)total_m, len);   // slot is 8-byte aligned; total_m <= EMIT_CAP
                        } else {                                       // one token per D op, then the trailing run
                            __syncwarp();
                            int32_t* d_off = SM.run_g;                 // the run list is no longer needed
                            int tlen = 0;
                            for (int k0 = 0; k0 < nD; k0 += 32) {
                                const int k = k0 + lane;
                                const int t = k < nD ? ndigits((uint32_t)SM.d_run[k]) + 1 + SM.d_ct[k] : 0;
                                int tot; const int to = tlen + warp_excl_scan(t, lane, tot);
                                if (k < nD) d_off[k] = to;
                                tlen += tot;
                            }
                            const int final_run = total_m - m_at_last_d, ndf = final_run > 0 ? ndigits((uint32_t)final_run) : 0;
                            len = tlen + ndf;
                            off = len <= 8 ? O.md_inline + 8 * i : pool_grab(O.md_cursor, O.md_cap, len, lane);
                            __syncwarp();
                            if (off >= 0) {
                                uint8_t* out = O.md_pool + off;
                                for (int k = lane; k < nD; k += 32) { const int nd = ndigits((uint32_t)SM.d_run[k]); write_dec(out + d_off[k], (uint32_t)SM.d_run[k], nd); out[d_off[k] + nd] = '^'; }
                                for (int k = 0; k < nD; k++) {
                                    uint8_t* bp = out + d_off[k] + ndigits((uint32_t)SM.d_run[k]) + 1; const int64_t g = gbase + SM.d_g[k]; const int ct = SM.d_ct[k];
                                    for (int j = lane; j < ct; j += 32) bp[j] = genome_byte(P.G, g + j);
                                }
                                if (ndf > 0 && lane == 0) write_dec(out + tlen, (uint32_t)final_run, ndf);
                            }
                        }
                        done = true;
                    }
                }
            }
            if (!done) {                                               // a difference to the genome, a literal run, the contig end,
                if (lane == 0) O.exact_list[atomicAdd(O.exact_count, 1u)] = (int32_t)i;   // a long read: k_nmmd_exact walks it
            } else {
                st = (st & ~(uint32_t)TC_ST_NMMD_NONE) | TC_ST_NMMD_DEVICE;
                if (lane == 0) { O.nm[i] = nm + nm_extra; O.md_off[i] = off < 0 ? 0 : off; O.md_len[i] = off < 0 ? 0 : len; W.status[i] = st; }
            }
        } else if (st & TC_ST_NMMD_DEVICE) {                           // generated at init and never refreshed
            const int len = W.md_init_len[i];
            const int64_t off = pool_grab(O.md_cursor, O.md_cap, len, lane);
            if (off >= 0) { const uint8_t* s = W.md_init_pool + W.md_init_off[i]; for (int j = lane; j < len; j += 32) O.md_pool[off + j] = s[j]; }
            if (lane == 0) { O.nm[i] = W.nm_init[i]; O.md_off[i] = off < 0 ? 0 : off; O.md_len[i] = off < 0 ? 0 : len; }
        } else if (lane == 0) { O.nm[i] = 0; O.md_off[i] = 0; O.md_len[i] = 0; }
        __syncwarp();
        pw = pw_next;
    }
}

// getNMandMDFlags (tr.py:280-342) for the reads k_emit could not settle with its clean-read check:
// the token-emitting walk over the final SEQ and CIGAR (both already in the output arrays).
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) k_nmmd_exact(Params P, OutDev O) {
    __shared__ uint32_t lut[256];
    for (int t = threadIdx.x; t < 256; t += blockDim.x) {
        const uint32_t acgt = 0x54474341u;
        lut[t] = ((acgt >> (8 * (t & 3))) & 255u) | (((acgt >> (8 * ((t >> 2) & 3))) & 255u) << 8) |
                 (((acgt >> (8 * ((t >> 4) & 3))) & 255u) << 16) | (((acgt >> (8 * ((t >> 6) & 3))) & 255u) << 24);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t nw = (int64_t)gridDim.x * WARPS_PER_BLOCK, nl = (int64_t)*O.exact_count;
    for (int64_t t = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5); t < nl; t += nw) {
        const int64_t i = O.exact_list[t];
        const int32_t* w = P.W.plan[i].w;
        const int64_t o_seq = ((int64_t)w[EP_OSEQ_HI] << 32) | (uint32_t)w[EP_OSEQ_LO], o_cig = ((int64_t)w[EP_OCIG_HI] << 32) | (uint32_t)w[EP_OCIG_LO];
        CigSrc cs; cs.packed = 0; cs.op = O.cig_op + o_cig; cs.len = O.cig_len + o_cig; cs.n = w[EP_NCIG];
        int nm, len; int64_t off; bool none;
        nmmd_to_pool<true>(P.G, O.seq + o_seq, lut, cs, w[EP_CHROM], (int64_t)w[EP_POS], lane, O.md_cursor, O.md_cap, O.md_pool, O.md_inline + 8 * i, nm, len, off, none);
        uint32_t st = ((uint32_t)w[EP_STATUS] & ~(uint32_t)TC_ST_NMMD_NONE) | TC_ST_NMMD_DEVICE;
        if (none) st |= TC_ST_NMMD_NONE;
        if (lane == 0) { O.nm[i] = nm + w[EP_NMEXTRA]; O.md_off[i] = off < 0 ? 0 : off; O.md_len[i] = off < 0 ? 0 : len; P.W.status[i] = st; }
    }
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { ctx->err = std::string(#x) + ": " + cudaGetErrorString(e_); return TC_E_CUDA; } } while (0)

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct HostBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 256;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocPortable | cudaHostAllocMapped);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};
#else
// ================================================================================== HOSTSIM
#define CK(x) do { int e_ = (x); if (e_ != 0) { ctx->err = #x; return TC_E_CUDA; } } while (0)
struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    int ensure(size_t n) { if (n <= cap) return 0; free(p); cap = n + n / 8 + 256; p = calloc(cap, 1); return p ? 0 : 1; }
    void release() { free(p); p = nullptr; cap = 0; }
};
typedef DevBuf HostBuf;

// scalar twin of warp_nmmd (test infrastructure; tr.py:280-342)
static void host_nmmd(const GenomeDev& G, const uint8_t* seq, const CigSrc& cs, int chrom, int64_t pos,
                      int& nm_out, bool& none_out, std::string& md) {
    const int64_t goff = G.chrom_off[chrom] - 1, clen = G.chrom_len[chrom];
    int nm = 0; int64_t run = 0, q = 0, g = pos; none_out = false; md.clear();
    for (int k = 0; k < cs.n; k++) {
        int op; int32_t ct; cig_get(cs, k, op, ct);
        if (op == 'M') {
            for (int32_t j = 0; j < ct; j++) {
                if (g < 1 || g > clen) { none_out = true; nm_out = 0; md.clear(); return; }
                uint8_t gb = genome_byte(G, goff + g), rb = seq[q];
                if (up8(gb) != up8(rb)) { md += std::to_string(run); md.push_back((char)gb); run = 0; nm++; } else run++;
                q++; g++;
            }
        } else if (op == 'D') {
            int gl = g < 1 ? 0 : fetch_len(G, chrom, g, g + ct - 1);
            if (gl == 0) { none_out = true; nm_out = 0; md.clear(); return; }
            md += std::to_string(run); md.push_back('^'); run = 0;
            for (int j = 0; j < gl; j++) md.push_back((char)genome_byte(G, goff + g + j));
            nm += ct; g += ct;
        } else if (op == 'I') { q += ct; nm += ct; }
        else if (op == 'S') q += ct;
        else if (op == 'N' || op == 'H') g += ct;
    }
    if (run > 0) md += std::to_string(run);
    nm_out = nm;
}
#endif

// ---------------------------------------------------------------------------------- context

// Everything that belongs to one batch (or one chunk of a pipelined batch) on the device.
struct Slot {
    DevBuf b_seqp, b_seq_poff, o_seqp;   // 4-bit SEQ: packed input rows + their offsets, packed output
    DevBuf b_flag, b_chrom, b_pos, b_tags, b_seq_off, b_seq, b_cig_off, b_cig_op, b_cig_len, b_md_off, b_md, b_ji_off, b_ji;
    BatchDev B; int64_t n = 0, n_seq = 0, n_cig = 0, n_md = 0, n_ji = 0, md_base = 0;
    int64_t seq_lo_word = 0, seq_hi_word = 0;       // words of B.seq (as uint32) that may be dereferenced
    DevBuf w_status, w_counters, w_md_init_off, w_md_init_len, w_nm_init, w_n_mdtok, w_md_tok, w_n_jn, w_mflags,
        w_cig_cap, w_seg_cap, w_te_cap, w_ws_off, w_ws, w_cig_buf, w_n_cig_cur, w_seg_buf, w_n_seg_cur, w_n_te,
        w_te_cnt, w_nm_extra, w_refresh, w_o_seq, w_o_cig, w_o_te, w_o_jn, w_seq_len, w_plan, w_fix_list, init_pool, cursors, scan_tmp;
    WorkDev W;
    DevBuf o_nm, o_md_off, o_md_len, o_md_pool, o_seq, o_cig_op, o_cig_len, o_te, o_jm, o_ji;
    int64_t tot_seq = 0, tot_cig = 0, tot_te = 0, tot_jn = 0, md_used = 0;
    int64_t base_seq = 0, base_cig = 0, base_te = 0, base_jn = 0, base_md = 0;   // position of this chunk in the batch output
    HostBuf h_tot;
    bool ran = false, ran_dry = false;
#ifndef TC_HOSTSIM
    cudaEvent_t ev_in = nullptr, ev_cmp = nullptr, ev_out = nullptr;
#endif
    void release() {
        DevBuf* d[] = {&b_flag, &b_chrom, &b_pos, &b_tags, &b_seq_off, &b_seq, &b_cig_off, &b_cig_op, &b_cig_len, &b_md_off, &b_md, &b_ji_off, &b_ji,
                       &w_status, &w_counters, &w_md_init_off, &w_md_init_len, &w_nm_init, &w_n_mdtok, &w_md_tok, &w_n_jn, &w_mflags, &w_cig_cap, &w_seg_cap,
                       &w_te_cap, &w_ws_off, &w_ws, &w_cig_buf, &w_n_cig_cur, &w_seg_buf, &w_n_seg_cur, &w_n_te, &w_te_cnt, &w_nm_extra, &w_refresh,
                       &w_o_seq, &w_o_cig, &w_o_te, &w_o_jn, &w_seq_len, &w_plan, &w_fix_list, &init_pool, &cursors, &scan_tmp,
                       &o_nm, &o_md_off, &o_md_len, &o_md_pool, &o_seq, &o_cig_op, &o_cig_len, &o_te, &o_jm, &o_ji};
        for (DevBuf* b : d) b->release();
        b_seqp.release(); b_seq_poff.release(); o_seqp.release();
        h_tot.release();
    }
};


struct tc_ctx {
    int device = 0;
    std::string err;
    tc_options opt;
#ifndef TC_HOSTSIM
    cudaStream_t stream = nullptr, s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev[12];
#endif
    bool have_genome = false;
    DevBuf g2, lo1, excblk, chrom_off, chrom_len, exc_start, exc_end, exc_byte;
    DevBuf don_s, acc_s, don_u, acc_u, ann_k1, ann_k2, sj_rng, ann_hash;
    DevBuf snp_key, snp_alt, ins_key, ins_size, ins_aoff, ins_bytes, del_key, del_size;
    GenomeDev G; SjDev S; VarDev V;
    Slot slot[2];
    int64_t n = 0; bool ran = false, ran_dry = false; bool piped = false;
    int64_t chunk_reads = 131072;   // reads per chunk of the pipelined tc_correct_batch (0 = never pipeline)
    uint8_t alphabet[16] = {'A', 'C', 'G', 'T', 'N', 'a', 'c', 'g', 't', 'n', 0, 0, 0, 0, 0, 0}; int n_alpha = 10;   // 4-bit SEQ codes
    bool genome_byte[256] = {};     // bytes the loaded genome can put into a read
    bool packed = false;            // the current batch came with 4-bit SEQ (results go back packed)
    // pinned host outputs (whole batch)
    HostBuf h_status, h_counters, h_nm, h_seq_off, h_seq, h_cig_off, h_cig_op, h_cig_len, h_md_off, h_md_len, h_md,
        h_jn_off, h_jm, h_ji, h_te_off, h_te, h_seq_len;
    tc_timing tm;
};

static std::string g_create_err;

#ifndef TC_HOSTSIM
__global__ void k_set_u64(unsigned long long* p, unsigned long long v) { *p = v; }

struct Alphabet { uint8_t c[16]; };
// 4-bit SEQ rows -> ASCII rows, one warp per read: a lane turns one 32-bit word (8 bases) into 8 bytes
// through a 256-entry table (packed byte -> two characters).
__global__ void __launch_bounds__(256) k_seq_unpack(int64_t n, const int64_t* seq_off, const int64_t* poff, const uint8_t* packed,
                                                    uint8_t* seq, Alphabet A) {
    __shared__ uint16_t lut[256];
    for (int t = threadIdx.x; t < 256; t += blockDim.x) lut[t] = (uint16_t)(A.c[t & 15] | (A.c[t >> 4] << 8));
    __syncthreads();
    const int lane = threadIdx.x & 31; const int64_t nw = (int64_t)gridDim.x * 8;
    for (int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += nw) {
        const int64_t o = seq_off[i]; const int L = (int)(seq_off[i + 1] - o);
        const uint32_t* src = reinterpret_cast<const uint32_t*>(packed + poff[i]);
        uint8_t* dst = seq + o;
        const int mis = (int)(reinterpret_cast<uintptr_t>(dst) & 3);   // the same for every 8-byte group of the row
        for (int w = lane; w * 8 < L; w += 32) {
            const uint32_t x = src[w];
            const uint32_t lo = lut[x & 255u] | ((uint32_t)lut[(x >> 8) & 255u] << 16), hi = lut[(x >> 16) & 255u] | ((uint32_t)lut[x >> 24] << 16);
            const int left = L - w * 8; uint8_t* d = dst + w * 8;
            if (left < 8) {
#pragma unroll
                for (int k = 0; k < 7; k++) if (k < left) d[k] = (uint8_t)((k < 4 ? lo : hi) >> (8 * (k & 3)));
            } else if (mis == 0) { reinterpret_cast<uint32_t*>(d)[0] = lo; reinterpret_cast<uint32_t*>(d)[1] = hi; }
            else if (mis == 2) {                       // widest naturally aligned stores: 2 + 4 + 2
                *reinterpret_cast<uint16_t*>(d) = (uint16_t)lo; *reinterpret_cast<uint32_t*>(d + 2) = (lo >> 16) | (hi << 16); *reinterpret_cast<uint16_t*>(d + 6) = (uint16_t)(hi >> 16);
            } else if (mis == 1) {                     // 1 + 2 + 4 + 1
                d[0] = (uint8_t)lo; *reinterpret_cast<uint16_t*>(d + 1) = (uint16_t)(lo >> 8); *reinterpret_cast<uint32_t*>(d + 3) = (lo >> 24) | (hi << 8); d[7] = (uint8_t)(hi >> 24);
            } else {                                   // 1 + 4 + 2 + 1
                d[0] = (uint8_t)lo; *reinterpret_cast<uint32_t*>(d + 1) = (lo >> 8) | (hi << 24); *reinterpret_cast<uint16_t*>(d + 5) = (uint16_t)(hi >> 8); d[7] = (uint8_t)(hi >> 24);
            }
        }
    }
}
// ASCII rows -> 4-bit rows over the whole dense output buffer (rows are 16-byte aligned, so packed byte j
// holds bases 2j and 2j+1 whatever the row): 16 bytes in, 8 bytes out per thread
__global__ void __launch_bounds__(256) k_seq_pack(int64_t n16, const uint4* seq, uint2* out, Alphabet A) {
    __shared__ uint8_t inv[256];
    inv[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x < 16 && (threadIdx.x == 0 || A.c[threadIdx.x] != 0)) inv[A.c[threadIdx.x]] = (uint8_t)threadIdx.x;
    __syncthreads();
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n16; t += (int64_t)gridDim.x * blockDim.x) {
        const uint4 v = seq[t];
        auto pk = [&](uint32_t w) { return (uint32_t)inv[w & 255u] | ((uint32_t)inv[(w >> 8) & 255u] << 4) | ((uint32_t)inv[(w >> 16) & 255u] << 8) | ((uint32_t)inv[w >> 24] << 12); };
        uint2 o; o.x = pk(v.x) | (pk(v.y) << 16); o.y = pk(v.z) | (pk(v.w) << 16);
        out[t] = o;
    }
}

static int be_h2d(tc_ctx* ctx, cudaStream_t st, DevBuf& d, size_t dst_off, const void* src, size_t bytes, size_t slack = 0) {
    CK(d.ensure(dst_off + bytes + slack + 8));
    if (bytes) CK(cudaMemcpyAsync((char*)d.p + dst_off, src, bytes, cudaMemcpyHostToDevice, st));
    ctx->tm.h2d_bytes += (int64_t)bytes;
    return 0;
}
// grows a pinned host buffer, keeping its content; `busy` = a stream that may still be writing into it
static int host_reserve(tc_ctx* ctx, HostBuf& h, size_t need, size_t keep, cudaStream_t busy) {
    if (need <= h.cap) return 0;
    if (busy) CK(cudaStreamSynchronize(busy));
    void* old = h.p; h.p = nullptr; h.cap = 0;
    size_t want = need + need / 4 + 4096;
    CK(cudaHostAlloc(&h.p, want, cudaHostAllocDefault)); h.cap = want;
    if (old) { if (keep) memcpy(h.p, old, keep); cudaFreeHost(old); }
    return 0;
}
static int be_d2h(tc_ctx* ctx, cudaStream_t st, HostBuf& h, size_t dst_off, const void* src, size_t bytes) {
    if (bytes) CK(cudaMemcpyAsync((char*)h.p + dst_off, src, bytes, cudaMemcpyDeviceToHost, st));
    ctx->tm.d2h_bytes += (int64_t)bytes;
    return 0;
}
static int be_sync(tc_ctx* ctx, cudaStream_t st) { CK(cudaStreamSynchronize(st)); return 0; }
static int be_zero(tc_ctx* ctx, cudaStream_t st, void* p, size_t bytes) { CK(cudaMemsetAsync(p, 0, bytes, st)); return 0; }
// in-place exclusive prefix sum over n1 int64 values, starting at `init`
static int be_scan(tc_ctx* ctx, cudaStream_t st, DevBuf& tmpbuf, int64_t* a, int64_t n1, int64_t init) {
    size_t tmp = 0;
    CK(cub::DeviceScan::ExclusiveScan(nullptr, tmp, a, a, ::cuda::std::plus<>{}, init, (int)n1, st));
    CK(tmpbuf.ensure(tmp));
    CK(cub::DeviceScan::ExclusiveScan(tmpbuf.p, tmp, a, a, ::cuda::std::plus<>{}, init, (int)n1, st));
    ctx->tm.launches += 2;
    return 0;
}
// One 8-byte total back to the host.  `host` lies in pinned memory, which the device can address
// (unified addressing), so a one-thread kernel stores it: a cudaMemcpyAsync would queue on the
// device-to-host copy engine behind the bulk download of the previous chunk and stall the pipeline.
__global__ void k_publish_i64(const int64_t* dev, int64_t* host) { *host = *dev; __threadfence_system(); }
static int be_read_i64(tc_ctx* ctx, cudaStream_t st, const int64_t* dev, int64_t* host) {
    k_publish_i64<<<1, 1, 0, st>>>(dev, host);
    CK(cudaGetLastError());
    return 0;
}
typedef cudaStream_t stream_t;
#else
typedef int stream_t;
static int be_h2d(tc_ctx* ctx, stream_t, DevBuf& d, size_t dst_off, const void* src, size_t bytes, size_t slack = 0) {
    CK(d.ensure(dst_off + bytes + slack + 8)); if (bytes) memcpy((char*)d.p + dst_off, src, bytes); ctx->tm.h2d_bytes += (int64_t)bytes; return 0;
}
static int host_reserve(tc_ctx* ctx, HostBuf& h, size_t need, size_t keep, stream_t) {
    if (need <= h.cap) return 0;
    void* old = h.p; size_t want = need + need / 4 + 4096;
    h.p = calloc(want, 1); h.cap = want;
    if (!h.p) { ctx->err = "out of host memory"; return 1; }
    if (old) { if (keep) memcpy(h.p, old, keep); free(old); }
    return 0;
}
static int be_d2h(tc_ctx* ctx, stream_t, HostBuf& h, size_t dst_off, const void* src, size_t bytes) {
    if (bytes) memcpy((char*)h.p + dst_off, src, bytes); ctx->tm.d2h_bytes += (int64_t)bytes; return 0;
}
static int be_sync(tc_ctx*, stream_t) { return 0; }
static int be_zero(tc_ctx*, stream_t, void* p, size_t bytes) { memset(p, 0, bytes); return 0; }
static int be_scan(tc_ctx*, stream_t, DevBuf&, int64_t* a, int64_t n1, int64_t init) { int64_t s = init; for (int64_t k = 0; k < n1; k++) { int64_t v = a[k]; a[k] = s; s += v; } return 0; }
static int be_read_i64(tc_ctx*, stream_t, const int64_t* dev, int64_t* host) { *host = *dev; return 0; }
#endif

extern "C" {

tc_ctx* tc_ctx_create(int device) {
    tc_ctx* ctx = new tc_ctx();
    ctx->device = device;
    memset(&ctx->opt, 0, sizeof(ctx->opt));
    ctx->opt.max_len_indel = 5; ctx->opt.max_sj_offset = 5;
    ctx->opt.correct_mismatches = ctx->opt.correct_indels = ctx->opt.correct_sjs = 1;
    memset(&ctx->G, 0, sizeof(ctx->G)); memset(&ctx->S, 0, sizeof(ctx->S)); memset(&ctx->V, 0, sizeof(ctx->V));
    memset(&ctx->tm, 0, sizeof(ctx->tm));
#ifndef TC_HOSTSIM
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking);
    for (int k = 0; k < 12 && e == cudaSuccess; k++) e = cudaEventCreate(&ctx->ev[k]);
    for (int k = 0; k < 2 && e == cudaSuccess; k++) {
        e = cudaEventCreateWithFlags(&ctx->slot[k].ev_in, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->slot[k].ev_cmp, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->slot[k].ev_out, cudaEventDisableTiming);
    }
    if (e != cudaSuccess) { g_create_err = std::string("tc_ctx_create: ") + cudaGetErrorString(e); delete ctx; return nullptr; }
#endif
    return ctx;
}

void tc_ctx_destroy(tc_ctx* ctx) {
    if (!ctx) return;
#ifndef TC_HOSTSIM
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream); cudaStreamSynchronize(ctx->s_in); cudaStreamSynchronize(ctx->s_out);
#endif
    DevBuf* d[] = {&ctx->g2, &ctx->lo1, &ctx->excblk, &ctx->chrom_off, &ctx->chrom_len, &ctx->exc_start, &ctx->exc_end, &ctx->exc_byte,
                   &ctx->don_s, &ctx->acc_s, &ctx->don_u, &ctx->acc_u, &ctx->ann_k1, &ctx->ann_k2, &ctx->snp_key, &ctx->snp_alt,
                   &ctx->ins_key, &ctx->ins_size, &ctx->ins_aoff, &ctx->ins_bytes, &ctx->del_key, &ctx->del_size};
    for (DevBuf* b : d) b->release();
    for (int k = 0; k < 2; k++) ctx->slot[k].release();
    HostBuf* h[] = {&ctx->h_status, &ctx->h_counters, &ctx->h_nm, &ctx->h_seq_off, &ctx->h_seq, &ctx->h_cig_off, &ctx->h_cig_op, &ctx->h_cig_len,
                    &ctx->h_md_off, &ctx->h_md_len, &ctx->h_md, &ctx->h_jn_off, &ctx->h_jm, &ctx->h_ji, &ctx->h_te_off, &ctx->h_te, &ctx->h_seq_len};
    for (HostBuf* b : h) b->release();
#ifndef TC_HOSTSIM
    for (int k = 0; k < 12; k++) cudaEventDestroy(ctx->ev[k]);
    for (int k = 0; k < 2; k++) { cudaEventDestroy(ctx->slot[k].ev_in); cudaEventDestroy(ctx->slot[k].ev_cmp); cudaEventDestroy(ctx->slot[k].ev_out); }
    cudaStreamDestroy(ctx->stream); cudaStreamDestroy(ctx->s_in); cudaStreamDestroy(ctx->s_out);
#endif
    delete ctx;
}

const char* tc_last_error(tc_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int tc_ctx_set_options(tc_ctx* ctx, const tc_options* o) {
    if (!ctx || !o) return TC_E_ARG;
    if (o->max_len_indel < 0 || o->max_sj_offset < 0) { ctx->err = "negative maxLenIndel / maxSJOffset"; return TC_E_ARG; }
    ctx->opt = *o;
    return TC_OK;
}

void* tc_ctx_stream(tc_ctx* ctx) {
#ifndef TC_HOSTSIM
    return ctx ? (void*)ctx->stream : nullptr;
#else
    (void)ctx; return nullptr;
#endif
}

static stream_t main_stream(tc_ctx* ctx) {
#ifndef TC_HOSTSIM
    return ctx->stream;
#else
    (void)ctx; return 0;
#endif
}

int tc_ctx_load_genome(tc_ctx* ctx, int32_t n_chrom, const int64_t* chrom_len, const int64_t* chrom_off,
                       int64_t n_pad, const uint32_t* bases2, const uint32_t* lower1, int64_t n_exc,
                       const int64_t* exc_start, const int64_t* exc_end, const uint8_t* exc_byte) {
    if (!ctx || n_chrom <= 0 || n_pad <= 0 || (n_pad & 63)) { if (ctx) ctx->err = "tc_ctx_load_genome: bad sizes"; return TC_E_ARG; }
#ifndef TC_HOSTSIM
    CK(cudaSetDevice(ctx->device));
#endif
    stream_t st = main_stream(ctx);
    // block bitmap of literal runs (1 bit per 256 bases)
    int64_t nblk = (n_pad >> 8) + 2, nwords = (nblk + 31) / 32 + 1;
    std::vector<uint32_t> blk((size_t)nwords, 0u);
    for (int64_t k = 0; k < n_exc; k++) {
        if (exc_end[k] <= exc_start[k] || exc_start[k] < 0 || exc_end[k] > n_pad) { ctx->err = "tc_ctx_load_genome: bad literal run"; return TC_E_ARG; }
        for (int64_t b = exc_start[k] >> 8; b <= (exc_end[k] - 1) >> 8; b++) blk[(size_t)(b >> 5)] |= 1u << (b & 31);
    }
    if (be_h2d(ctx, st, ctx->g2, 0, bases2, (size_t)(n_pad / 16) * 4, 256) || be_h2d(ctx, st, ctx->lo1, 0, lower1, (size_t)(n_pad / 32) * 4, 256) ||
        be_h2d(ctx, st, ctx->excblk, 0, blk.data(), (size_t)nwords * 4) || be_h2d(ctx, st, ctx->chrom_off, 0, chrom_off, (size_t)n_chrom * 8) ||
        be_h2d(ctx, st, ctx->chrom_len, 0, chrom_len, (size_t)n_chrom * 8) || be_h2d(ctx, st, ctx->exc_start, 0, exc_start, (size_t)n_exc * 8) ||
        be_h2d(ctx, st, ctx->exc_end, 0, exc_end, (size_t)n_exc * 8) || be_h2d(ctx, st, ctx->exc_byte, 0, exc_byte, (size_t)n_exc) || be_sync(ctx, st)) return TC_E_CUDA;
    GenomeDev& G = ctx->G;
    G.g2 = (const uint32_t*)ctx->g2.p; G.lo1 = (const uint32_t*)ctx->lo1.p; G.excblk = (const uint32_t*)ctx->excblk.p;
    G.chrom_off = (const int64_t*)ctx->chrom_off.p; G.chrom_len = (const int64_t*)ctx->chrom_len.p;
    G.exc_start = (const int64_t*)ctx->exc_start.p; G.exc_end = (const int64_t*)ctx->exc_end.p;
    G.exc_byte = (const uint8_t*)ctx->exc_byte.p; G.n_exc = n_exc; G.n_chrom = n_chrom;
    memset(ctx->genome_byte, 0, sizeof(ctx->genome_byte));
    for (const char* c = "ACGTacgt"; *c; c++) ctx->genome_byte[(uint8_t)*c] = true;
    for (int64_t k = 0; k < n_exc; k++) ctx->genome_byte[exc_byte[k]] = true;
    ctx->have_genome = true;
    return TC_OK;
}

int tc_ctx_load_junctions(tc_ctx* ctx, int64_t n_don_s, const uint64_t* don_s, int64_t n_acc_s, const uint64_t* acc_s,
                          int64_t n_don_u, const uint64_t* don_u, int64_t n_acc_u, const uint64_t* acc_u,
                          int64_t n_annot, const uint64_t* k1, const uint64_t* k2) {
    if (!ctx) return TC_E_ARG;
#ifndef TC_HOSTSIM
    CK(cudaSetDevice(ctx->device));
#endif
    stream_t st = main_stream(ctx);
    if (be_h2d(ctx, st, ctx->don_s, 0, don_s, (size_t)n_don_s * 8) || be_h2d(ctx, st, ctx->acc_s, 0, acc_s, (size_t)n_acc_s * 8) ||
        be_h2d(ctx, st, ctx->don_u, 0, don_u, (size_t)n_don_u * 8) || be_h2d(ctx, st, ctx->acc_u, 0, acc_u, (size_t)n_acc_u * 8) ||
        be_h2d(ctx, st, ctx->ann_k1, 0, k1, (size_t)n_annot * 8) || be_h2d(ctx, st, ctx->ann_k2, 0, k2, (size_t)n_annot * 8) || be_sync(ctx, st)) return TC_E_CUDA;
    // first row of every chromosome(/strand) prefix in the four site tables, so that the kernels
    // search only inside the prefix's rows: rng[t] has np+1 entries, prefix p owns [rng[t][p], rng[t][p+1])
    const uint64_t* tabs[4] = {don_s, acc_s, don_u, acc_u}; const int64_t ns[4] = {n_don_s, n_acc_s, n_don_u, n_acc_u};
    int64_t np = 0;
    for (int t = 0; t < 4; t++) if (ns[t] > 0) { int64_t top = (int64_t)(tabs[t][ns[t] - 1] >> 32) + 1; if (top > np) np = top; }
    std::vector<int32_t> rng((size_t)(4 * (np + 1)), 0);
    for (int t = 0; t < 4; t++) {
        int32_t* r = rng.data() + (size_t)t * (np + 1); int64_t k = 0;
        for (int64_t pfx = 0; pfx <= np; pfx++) { while (k < ns[t] && (int64_t)(tabs[t][k] >> 32) < pfx) k++; r[pfx] = (int32_t)k; }
    }
    if (be_h2d(ctx, st, ctx->sj_rng, 0, rng.data(), rng.size() * 4) || be_sync(ctx, st)) return TC_E_CUDA;
    // annotated set as an open-addressing table (load factor <= 1/2, linear probing)
    uint32_t slots = 16; while ((int64_t)slots < 2 * n_annot) slots <<= 1;
    std::vector<AnnSlot> tab(slots); for (auto& e : tab) { e.k1 = ~0ull; e.k2 = 0; }
    for (int64_t k = 0; k < n_annot; k++) {
        uint32_t h = ann_hash_of(k1[k], k2[k]) & (slots - 1);
        while (tab[h].k1 != ~0ull && !(tab[h].k1 == k1[k] && tab[h].k2 == k2[k])) h = (h + 1) & (slots - 1);
        tab[h].k1 = k1[k]; tab[h].k2 = k2[k];
    }
    if (be_h2d(ctx, st, ctx->ann_hash, 0, tab.data(), tab.size() * sizeof(AnnSlot)) || be_sync(ctx, st)) return TC_E_CUDA;
    SjDev& S = ctx->S;
    S.ann_hash = n_annot > 0 ? (const AnnSlot*)ctx->ann_hash.p : nullptr; S.ann_mask = slots - 1;
    S.n_prefix = (int32_t)np;
    S.rng_don_s = (const int32_t*)ctx->sj_rng.p; S.rng_acc_s = S.rng_don_s + (np + 1); S.rng_don_u = S.rng_acc_s + (np + 1); S.rng_acc_u = S.rng_don_u + (np + 1);
    S.don_s = (const u64*)ctx->don_s.p; S.acc_s = (const u64*)ctx->acc_s.p; S.don_u = (const u64*)ctx->don_u.p; S.acc_u = (const u64*)ctx->acc_u.p;
    S.ann_k1 = (const u64*)ctx->ann_k1.p; S.ann_k2 = (const u64*)ctx->ann_k2.p;
    S.n_don_s = n_don_s; S.n_acc_s = n_acc_s; S.n_don_u = n_don_u; S.n_acc_u = n_acc_u; S.n_ann = n_annot;
    return TC_OK;
}

int tc_ctx_load_variants(tc_ctx* ctx, int64_t n_snp, const uint64_t* snp_key, const uint8_t* snp_alt, int64_t n_ins,
                         const uint64_t* ins_key, const int32_t* ins_size, const int64_t* ins_aoff, const uint8_t* ins_bytes,
                         int64_t n_del, const uint64_t* del_key, const int32_t* del_size) {
    if (!ctx) return TC_E_ARG;
#ifndef TC_HOSTSIM
    CK(cudaSetDevice(ctx->device));
#endif
    stream_t st = main_stream(ctx);
    int64_t nab = n_ins ? ins_aoff[n_ins] : 0;
    if (be_h2d(ctx, st, ctx->snp_key, 0, snp_key, (size_t)n_snp * 8) || be_h2d(ctx, st, ctx->snp_alt, 0, snp_alt, (size_t)n_snp) ||
        be_h2d(ctx, st, ctx->ins_key, 0, ins_key, (size_t)n_ins * 8) || be_h2d(ctx, st, ctx->ins_size, 0, ins_size, (size_t)n_ins * 4) ||
        be_h2d(ctx, st, ctx->ins_aoff, 0, ins_aoff, (size_t)(n_ins ? n_ins + 1 : 0) * 8) || be_h2d(ctx, st, ctx->ins_bytes, 0, ins_bytes, (size_t)nab) ||
        be_h2d(ctx, st, ctx->del_key, 0, del_key, (size_t)n_del * 8) || be_h2d(ctx, st, ctx->del_size, 0, del_size, (size_t)n_del * 4) || be_sync(ctx, st)) return TC_E_CUDA;
    VarDev& V = ctx->V;
    V.snp_key = (const u64*)ctx->snp_key.p; V.snp_alt = (const uint8_t*)ctx->snp_alt.p; V.n_snp = n_snp;
    V.ins_key = (const u64*)ctx->ins_key.p; V.ins_size = (const int32_t*)ctx->ins_size.p; V.ins_aoff = (const int64_t*)ctx->ins_aoff.p;
    V.ins_bytes = (const uint8_t*)ctx->ins_bytes.p; V.n_ins = n_ins;
    V.del_key = (const u64*)ctx->del_key.p; V.del_size = (const int32_t*)ctx->del_size.p; V.n_del = n_del;
    return TC_OK;
}

}  // extern "C"

// 4-bit SEQ rows of reads [r0, r1): upload, then expand to the ASCII rows the kernels read (B.seq at d0)
static int seq_upload_packed(tc_ctx* ctx, Slot& S, const tc_batch_in* in, int64_t r0, int64_t r1, size_t d0, stream_t st) {
    const int64_t n = r1 - r0;
    CK(S.b_seq.ensure(d0 + (size_t)S.n_seq + 256 + 8));
    if (n == 0) return 0;
    const int64_t p0 = in->seq_poff[r0], pbytes = in->seq_poff[r1] - p0;
    if (be_h2d(ctx, st, S.b_seqp, 0, in->seq + p0, (size_t)pbytes, 64) || be_h2d(ctx, st, S.b_seq_poff, 0, in->seq_poff + r0, (size_t)(n + 1) * 8)) return TC_E_CUDA;
    const int64_t s0 = in->seq_off[r0];
    uint8_t* dst = (uint8_t*)S.b_seq.p + d0 - s0; const uint8_t* src = (const uint8_t*)S.b_seqp.p - p0;
#ifndef TC_HOSTSIM
    Alphabet A; memcpy(A.c, ctx->alphabet, 16);
    int64_t wb = (n + 7) / 8; int g = (int)(wb < 148 * 8 ? wb : 148 * 8);
    k_seq_unpack<<<g, 256, 0, st>>>(n, (const int64_t*)S.b_seq_off.p, (const int64_t*)S.b_seq_poff.p, src, dst, A); ctx->tm.launches++;
    CK(cudaGetLastError());
#else
    const int64_t* so = (const int64_t*)S.b_seq_off.p; const int64_t* po = (const int64_t*)S.b_seq_poff.p;
    for (int64_t i = 0; i < n; i++)
        for (int64_t k = 0; k < so[i + 1] - so[i]; k++) { const uint8_t b = src[po[i] + (k >> 1)]; dst[so[i] + k] = ctx->alphabet[(k & 1) ? b >> 4 : b & 15]; }
#endif
    return 0;
}

// uploads reads [r0, r1) of `in` into slot S.  CSR offsets stay the batch's absolute offsets; the
// data pointers are shifted so that they can be applied unchanged.
static int slot_upload(tc_ctx* ctx, Slot& S, const tc_batch_in* in, int64_t r0, int64_t r1, stream_t st) {
    const int64_t n = r1 - r0;
    S.n = n; S.ran = false;
    static const int64_t zero_off[1] = {0};
    const int64_t s0 = n ? in->seq_off[r0] : 0, c0 = n ? in->cig_off[r0] : 0, m0 = n ? in->md_off[r0] : 0, j0 = n ? in->ji_off[r0] : 0;
    S.n_seq = n ? in->seq_off[r1] - s0 : 0; S.n_cig = n ? in->cig_off[r1] - c0 : 0;
    S.n_md = n ? in->md_off[r1] - m0 : 0; S.n_ji = n ? in->ji_off[r1] - j0 : 0; S.md_base = m0;
    const size_t d0 = (size_t)(s0 & 15);            // keeps (device pointer - s0) 16-byte aligned
    if (be_h2d(ctx, st, S.b_flag, 0, in->flag + r0, (size_t)n * 4) || be_h2d(ctx, st, S.b_chrom, 0, in->chrom + r0, (size_t)n * 4) ||
        be_h2d(ctx, st, S.b_pos, 0, in->pos + r0, (size_t)n * 4) || be_h2d(ctx, st, S.b_tags, 0, in->tags + r0, (size_t)n) ||
        be_h2d(ctx, st, S.b_seq_off, 0, n ? in->seq_off + r0 : zero_off, (size_t)(n + 1) * 8) ||
        (in->seq_poff ? seq_upload_packed(ctx, S, in, r0, r1, d0, st) : be_h2d(ctx, st, S.b_seq, d0, n ? in->seq + s0 : nullptr, (size_t)S.n_seq, 256)) ||
        be_h2d(ctx, st, S.b_cig_off, 0, n ? in->cig_off + r0 : zero_off, (size_t)(n + 1) * 8) ||
        be_h2d(ctx, st, S.b_cig_op, 0, n ? in->cig_op + c0 : nullptr, (size_t)S.n_cig) ||
        be_h2d(ctx, st, S.b_cig_len, 0, n ? in->cig_len + c0 : nullptr, (size_t)S.n_cig * 4) ||
        be_h2d(ctx, st, S.b_md_off, 0, n ? in->md_off + r0 : zero_off, (size_t)(n + 1) * 8) ||
        be_h2d(ctx, st, S.b_md, 0, n ? in->md + m0 : nullptr, (size_t)S.n_md) ||
        be_h2d(ctx, st, S.b_ji_off, 0, n ? in->ji_off + r0 : zero_off, (size_t)(n + 1) * 8) ||
        be_h2d(ctx, st, S.b_ji, 0, n ? in->ji + j0 : nullptr, (size_t)S.n_ji * 4)) return TC_E_CUDA;
    BatchDev& B = S.B;
    B.n = n; B.flag = (const int32_t*)S.b_flag.p; B.chrom = (const int32_t*)S.b_chrom.p; B.pos = (const int32_t*)S.b_pos.p;
    B.tags = (const uint8_t*)S.b_tags.p; B.seq_off = (const int64_t*)S.b_seq_off.p; B.seq = (const uint8_t*)S.b_seq.p + d0 - s0;
    B.cig_off = (const int64_t*)S.b_cig_off.p; B.cig_op = (const uint8_t*)S.b_cig_op.p - c0; B.cig_len = (const int32_t*)S.b_cig_len.p - c0;
    B.md_off = (const int64_t*)S.b_md_off.p; B.md = (const uint8_t*)S.b_md.p - m0; B.ji_off = (const int64_t*)S.b_ji_off.p; B.ji = (const int32_t*)S.b_ji.p - j0;
    S.seq_lo_word = (s0 - (int64_t)d0) / 4; S.seq_hi_word = S.seq_lo_word + (int64_t)(S.b_seq.cap / 4);
    return TC_OK;
}

static int alloc_work(tc_ctx* ctx, Slot& S) {
    const size_t n = (size_t)S.n, n1 = n + 1;
    CK(S.w_status.ensure(n1 * 4)); CK(S.w_counters.ensure(n1 * 40)); CK(S.w_md_init_off.ensure(n1 * 8));
    CK(S.w_md_init_len.ensure(n1 * 4)); CK(S.w_nm_init.ensure(n1 * 4)); CK(S.w_n_mdtok.ensure(n1 * 4));
    CK(S.w_n_jn.ensure(n1 * 4)); CK(S.w_mflags.ensure(n1)); CK(S.w_cig_cap.ensure(n1 * 4)); CK(S.w_seg_cap.ensure(n1 * 4));
    CK(S.w_te_cap.ensure(n1 * 4)); CK(S.w_ws_off.ensure(n1 * 8)); CK(S.w_cig_buf.ensure(n1)); CK(S.w_n_cig_cur.ensure(n1 * 4));
    CK(S.w_seg_buf.ensure(n1)); CK(S.w_n_seg_cur.ensure(n1 * 4)); CK(S.w_n_te.ensure(n1 * 4)); CK(S.w_te_cnt.ensure(n1 * 12));
    CK(S.w_nm_extra.ensure(n1 * 4)); CK(S.w_refresh.ensure(n1)); CK(S.w_o_seq.ensure(n1 * 8)); CK(S.w_o_cig.ensure(n1 * 8));
    CK(S.w_o_te.ensure(n1 * 8)); CK(S.w_o_jn.ensure(n1 * 8)); CK(S.w_seq_len.ensure(n1 * 4)); CK(S.w_plan.ensure(n1 * sizeof(EmitPlan))); CK(S.cursors.ensure(64));
    CK(S.o_nm.ensure(n1 * 4)); CK(S.o_md_off.ensure(n1 * 8)); CK(S.o_md_len.ensure(n1 * 4));
    WorkDev& W = S.W;
    W.status = (uint32_t*)S.w_status.p; W.counters = (int32_t*)S.w_counters.p;
    W.md_init_off = (int64_t*)S.w_md_init_off.p; W.md_init_len = (int32_t*)S.w_md_init_len.p; W.nm_init = (int32_t*)S.w_nm_init.p;
    W.n_mdtok = (int32_t*)S.w_n_mdtok.p; W.n_jn = (int32_t*)S.w_n_jn.p; W.mflags = (uint8_t*)S.w_mflags.p;
    CK(S.w_md_tok.ensure((size_t)(S.n_md + 1) * 4)); W.md_tok = (uint32_t*)S.w_md_tok.p - S.md_base;   // indexed by the batch's absolute md offsets
    W.cig_cap = (int32_t*)S.w_cig_cap.p; W.seg_cap = (int32_t*)S.w_seg_cap.p; W.te_cap = (int32_t*)S.w_te_cap.p;
    W.ws_off = (int64_t*)S.w_ws_off.p; W.cig_buf = (int8_t*)S.w_cig_buf.p; W.n_cig_cur = (int32_t*)S.w_n_cig_cur.p;
    W.seg_buf = (int8_t*)S.w_seg_buf.p; W.n_seg_cur = (int32_t*)S.w_n_seg_cur.p; W.n_te = (int32_t*)S.w_n_te.p;
    W.te_cnt = (int32_t*)S.w_te_cnt.p; W.nm_extra = (int32_t*)S.w_nm_extra.p; W.refresh = (uint8_t*)S.w_refresh.p;
    W.o_seq = (int64_t*)S.w_o_seq.p; W.o_cig = (int64_t*)S.w_o_cig.p; W.o_te = (int64_t*)S.w_o_te.p; W.o_jn = (int64_t*)S.w_o_jn.p;
    W.seq_len_out = (int32_t*)S.w_seq_len.p; W.plan = (EmitPlan*)S.w_plan.p;
    return 0;
}

static Params make_params(tc_ctx* ctx, Slot& S) {
    Params P; P.G = ctx->G; P.S = ctx->S; P.V = ctx->V; P.B = S.B; P.W = S.W; P.O = ctx->opt;
    P.sj_enabled = ctx->S.n_ann > 0 ? 1 : 0;
    return P;
}

static OutDev make_out(Slot& S, unsigned long long* cur, int64_t md_cap) {
    OutDev O; O.nm = (int32_t*)S.o_nm.p; O.md_off = (int64_t*)S.o_md_off.p; O.md_len = (int32_t*)S.o_md_len.p;
    O.md_pool = (uint8_t*)S.o_md_pool.p - S.base_md; O.md_cursor = cur; O.md_cap = S.base_md + md_cap; O.md_inline = S.base_md;
    O.seq = (uint8_t*)S.o_seq.p - S.base_seq; O.cig_op = (uint8_t*)S.o_cig_op.p - S.base_cig; O.cig_len = (int32_t*)S.o_cig_len.p - S.base_cig;
    O.te = (tc_te*)S.o_te.p - S.base_te; O.jm = (int8_t*)S.o_jm.p - S.base_jn; O.ji = (int32_t*)S.o_ji.p - 2 * S.base_jn;
    O.exact_list = (int32_t*)S.w_fix_list.p; O.exact_count = cur ? (unsigned*)(cur + 1) : nullptr;   // (the junction stage is done with the list)
    return O;
}

#ifdef TC_HOSTSIM
static void host_init_read(Params& P, int64_t i, std::vector<uint8_t>& pool) {
    int cls = read_class(P.B.flag[i], P.B.chrom[i]); uint32_t st = (uint32_t)cls; uint8_t tags = P.B.tags[i];
    if (cls == 0 && (!(tags & TAG_NM) || !(tags & TAG_MD))) {
        CigSrc cs; cs.packed = 0; cs.op = P.B.cig_op + P.B.cig_off[i]; cs.len = P.B.cig_len + P.B.cig_off[i]; cs.n = (int)(P.B.cig_off[i + 1] - P.B.cig_off[i]);
        int nm; bool none; std::string md;
        host_nmmd(P.G, P.B.seq + P.B.seq_off[i], cs, P.B.chrom[i], P.B.pos[i], nm, none, md);
        st |= TC_ST_NMMD_DEVICE; if (none) st |= TC_ST_NMMD_NONE;
        P.W.md_init_off[i] = (int64_t)pool.size(); P.W.md_init_len[i] = (int32_t)md.size(); P.W.nm_init[i] = nm;
        pool.insert(pool.end(), md.begin(), md.end());
    }
    P.W.status[i] = st;
}
// scalar twin of k_emit (test infrastructure)
static void host_emit_read(Params& P, OutDev& O, int64_t i, int dry, std::vector<uint8_t>& pool) {
    const BatchDev& B = P.B; const WorkDev& W = P.W;
    uint32_t st = W.status[i]; O.md_off[i] = 0; O.md_len[i] = 0; O.nm[i] = 0;
    if ((st & TC_ST_CLASS_MASK) != 0) return;
    WsView v = ws_view(W, i);
    small_copies_read(P, O, i, dry);
    if (dry) return;
    CigSrc cs;
    if (W.cig_buf[i] < 0) { cs.packed = 0; cs.op = B.cig_op + B.cig_off[i]; cs.len = B.cig_len + B.cig_off[i]; cs.n = (int)(B.cig_off[i + 1] - B.cig_off[i]); }
    else { cs.packed = v.cig[W.cig_buf[i]]; cs.op = 0; cs.len = 0; cs.n = W.n_cig_cur[i]; }
    const uint8_t* in_seq = B.seq + B.seq_off[i]; uint8_t* out_seq = O.seq + W.o_seq[i];
    if (W.seg_buf[i] < 0) memcpy(out_seq, in_seq, (size_t)(B.seq_off[i + 1] - B.seq_off[i]));
    else { int64_t o = 0; const u64* segs = v.seg[W.seg_buf[i]];
        for (int s = 0; s < W.n_seg_cur[i]; s++) { int64_t src = seg_src(segs[s]), len = seg_len(segs[s]);
            for (int64_t j = 0; j < len; j++) out_seq[o + j] = seg_kind(segs[s]) ? genome_byte(P.G, src + j) : in_seq[src + j];
            o += len; } }
    if (W.refresh[i]) {
        int nm; bool none; std::string md; host_nmmd(P.G, out_seq, cs, B.chrom[i], B.pos[i], nm, none, md);
        st = (st & ~(uint32_t)TC_ST_NMMD_NONE) | TC_ST_NMMD_DEVICE; if (none) st |= TC_ST_NMMD_NONE;
        O.nm[i] = nm + W.nm_extra[i]; O.md_off[i] = (int64_t)pool.size(); O.md_len[i] = (int32_t)md.size(); pool.insert(pool.end(), md.begin(), md.end());
        W.status[i] = st;
    } else if (st & TC_ST_NMMD_DEVICE) {
        O.nm[i] = W.nm_init[i]; O.md_off[i] = (int64_t)pool.size(); O.md_len[i] = W.md_init_len[i];
        pool.insert(pool.end(), W.md_init_pool + W.md_init_off[i], W.md_init_pool + W.md_init_off[i] + W.md_init_len[i]);
    }
}
#endif


// runs the kernel pipeline of one slot on stream `st`.  Output offsets start at the slot's
// base_* (its position in the whole batch).
static int slot_run(tc_ctx* ctx, Slot& S, int dry, stream_t st, bool timed) {
    const int64_t n = S.n;
    if (alloc_work(ctx, S)) return TC_E_CUDA;
    CK(S.h_tot.ensure(128)); int64_t* tot = (int64_t*)S.h_tot.p;
    unsigned long long* cur = (unsigned long long*)S.cursors.p;
    Params P = make_params(ctx, S);
    S.tot_seq = S.tot_cig = S.tot_te = S.tot_jn = 0; S.md_used = 0;
    if (n == 0) { S.ran = true; S.ran_dry = dry != 0; return TC_OK; }
#ifndef TC_HOSTSIM
    const int TPB = 128; const int gb = (int)((n + TPB - 1) / TPB);
    int nsm = 148; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device);
    int64_t wb = (n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK; int gw = (int)(wb < (int64_t)nsm * 8 ? wb : (int64_t)nsm * 8);
    if (timed) CK(cudaEventRecord(ctx->ev[2], st));
    // ---- init NM/MD (exact retry on overflow)
    int64_t init_cap = S.init_pool.cap > (1 << 20) + 64 ? (int64_t)S.init_pool.cap - 64 : (1 << 20);
    for (int attempt = 0;; attempt++) {
        CK(S.init_pool.ensure((size_t)init_cap + 64)); init_cap = (int64_t)S.init_pool.cap - 64;   // 64 spare bytes: md_tokenize reads whole 8-byte words
        if (be_zero(ctx, st, cur, 64)) return TC_E_CUDA;
        InitDev I; I.pool = (uint8_t*)S.init_pool.p; I.cursor = cur; I.cap = init_cap;
        P.W.md_init_pool = I.pool; S.W.md_init_pool = I.pool;
        CK(S.w_fix_list.ensure((size_t)(n + 1) * 4));                // (the junction stage reuses this list later)
        k_init<<<gb, TPB, 0, st>>>(P, (int32_t*)S.w_fix_list.p, (unsigned*)(cur + 2)); ctx->tm.launches++;
        CK(cudaGetLastError());
        k_init_nmmd<<<gw, WARPS_PER_BLOCK * 32, 0, st>>>(P, I, S.W.nm_init, (const int32_t*)S.w_fix_list.p, (const unsigned*)(cur + 2)); ctx->tm.launches++;
        CK(cudaGetLastError());
        CK(cudaMemsetAsync(S.W.counters, 0xff, (size_t)n * 40, st));
        k_measure<<<gb, TPB, 0, st>>>(P); ctx->tm.launches++;        // measure only needs lengths: safe before the overflow check
        CK(cudaGetLastError());
        if (be_zero(ctx, st, S.W.ws_off + n, 8) || be_scan(ctx, st, S.scan_tmp, S.W.ws_off, n + 1, 0) ||
            be_read_i64(ctx, st, (const int64_t*)cur, tot) || be_read_i64(ctx, st, S.W.ws_off + n, tot + 1) || be_sync(ctx, st)) return TC_E_CUDA;
        if (tot[0] <= init_cap) break;
        if (attempt >= 2) { ctx->err = "init MD pool overflow"; return TC_E_CUDA; }
        init_cap = tot[0] + 1024; ctx->tm.md_pool_retries++;
    }
    if (timed) CK(cudaEventRecord(ctx->ev[4], st));
    CK(S.w_ws.ensure((size_t)(tot[1] + 8) * 8));
    S.W.ws = (u64*)S.w_ws.p; P.W.ws = S.W.ws;
    if (dry) { k_dry<<<gb, TPB, 0, st>>>(P); ctx->tm.launches++; CK(cudaGetLastError()); if (timed) { CK(cudaEventRecord(ctx->ev[5], st)); CK(cudaEventRecord(ctx->ev[6], st)); } }
    else {
        k_plan<<<gb, TPB, 0, st>>>(P); ctx->tm.launches++; CK(cudaGetLastError());
        if (timed) CK(cudaEventRecord(ctx->ev[5], st));
        CK(S.w_fix_list.ensure((size_t)(n + 1) * 4));
        if (be_zero(ctx, st, cur + 2, 8)) return TC_E_CUDA;
        k_jn_init<<<gb, TPB, 0, st>>>(P, (int32_t*)S.w_fix_list.p, (unsigned*)(cur + 2)); ctx->tm.launches++; CK(cudaGetLastError());
        int occ_j = 1; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_j, k_jn_fix, WARPS_PER_BLOCK * 32, 0));
        const int gj = (int)(wb < (int64_t)nsm * occ_j ? wb : (int64_t)nsm * occ_j);            // persistent grid = the resident blocks
        k_jn_fix<<<gj, WARPS_PER_BLOCK * 32, 0, st>>>(P, (const int32_t*)S.w_fix_list.p, (const unsigned*)(cur + 2)); ctx->tm.launches++; CK(cudaGetLastError());
        if (timed) CK(cudaEventRecord(ctx->ev[6], st));
        k_sizes<<<gb, TPB, 0, st>>>(P); ctx->tm.launches++; CK(cudaGetLastError());
    }
    int64_t* sc[4] = {S.W.o_seq, S.W.o_cig, S.W.o_te, S.W.o_jn};
    const int64_t bases[4] = {S.base_seq, S.base_cig, S.base_te, S.base_jn};
    for (int k = 0; k < 4; k++) { if (be_zero(ctx, st, sc[k] + n, 8) || be_scan(ctx, st, S.scan_tmp, sc[k], n + 1, bases[k]) || be_read_i64(ctx, st, sc[k] + n, tot + 1 + k)) return TC_E_CUDA; }
    k_pack<<<gb, TPB, 0, st>>>(P); ctx->tm.launches++; CK(cudaGetLastError());
    if (be_sync(ctx, st)) return TC_E_CUDA;
    S.tot_seq = tot[1] - bases[0]; S.tot_cig = tot[2] - bases[1]; S.tot_te = tot[3] - bases[2]; S.tot_jn = tot[4] - bases[3];
    CK(S.o_seq.ensure((size_t)S.tot_seq + 64)); CK(S.o_cig_op.ensure((size_t)S.tot_cig + 64)); CK(S.o_cig_len.ensure((size_t)S.tot_cig * 4 + 64));
    CK(S.o_te.ensure((size_t)S.tot_te * 16 + 64)); CK(S.o_jm.ensure((size_t)S.tot_jn + 64)); CK(S.o_ji.ensure((size_t)S.tot_jn * 8 + 64));
    CK(cudaFuncSetAttribute(k_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EMIT_SMEM_BYTES));
    int occ = 1; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_emit, WARPS_PER_BLOCK * 32, EMIT_SMEM_BYTES));
    int ge = (int)(wb < (int64_t)nsm * occ ? wb : (int64_t)nsm * occ);
    int64_t md_cap = S.o_md_pool.cap > (size_t)(8 * n + (1 << 20)) ? (int64_t)S.o_md_pool.cap : (72 * n + 8 * S.n_cig + (1 << 20));
    {   // persistent grid: exactly the resident blocks, so that no block waits for a second wave
        int occ_s = 1; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_s, k_small, WARPS_PER_BLOCK * 32, 0));
        const int gs = (int)(wb < (int64_t)nsm * occ_s ? wb : (int64_t)nsm * occ_s);
        k_small<<<gs, WARPS_PER_BLOCK * 32, 0, st>>>(P, make_out(S, cur, md_cap), dry); ctx->tm.launches++; CK(cudaGetLastError());   // TE rows, CIGAR, jM / jI
    }
    if (timed) CK(cudaEventRecord(ctx->ev[7], st));
    for (int attempt = 0;; attempt++) {
        CK(S.o_md_pool.ensure((size_t)md_cap)); md_cap = (int64_t)S.o_md_pool.cap;
        k_set_u64<<<1, 1, 0, st>>>(cur + 1, 0ull);
        k_set_u64<<<1, 1, 0, st>>>(cur, (unsigned long long)(S.base_md + 8 * n));      // the pool starts behind the inline slots
        OutDev O = make_out(S, cur, md_cap);
        if (!dry) { k_emit<<<ge, WARPS_PER_BLOCK * 32, EMIT_SMEM_BYTES, st>>>(P, O, dry, S.seq_lo_word, S.seq_hi_word); ctx->tm.launches++; }   // (a dry run ends with k_pack)
        ctx->tm.launches += 2;
        CK(cudaGetLastError());
        if (timed && attempt == 0) CK(cudaEventRecord(ctx->ev[8], st));
        if (!dry) { k_nmmd_exact<<<gw, WARPS_PER_BLOCK * 32, 0, st>>>(P, O); ctx->tm.launches++; CK(cudaGetLastError()); }
        if (be_read_i64(ctx, st, (const int64_t*)cur, tot) || be_read_i64(ctx, st, (const int64_t*)cur + 1, tot + 6) || be_sync(ctx, st)) return TC_E_CUDA;
        if (tot[0] - S.base_md <= md_cap) { S.md_used = tot[0] - S.base_md; ctx->tm.exact_nmmd_reads += tot[6]; break; }
        if (attempt >= 2) { ctx->err = "MD pool overflow"; return TC_E_CUDA; }
        md_cap = tot[0] - S.base_md + 1024; ctx->tm.md_pool_retries++;
    }
    if (timed) {
        CK(cudaEventRecord(ctx->ev[9], st));
        CK(cudaStreamSynchronize(st));
        float t;
        cudaEventElapsedTime(&t, ctx->ev[2], ctx->ev[9]); ctx->tm.kernels_ms += t;
        cudaEventElapsedTime(&t, ctx->ev[2], ctx->ev[4]); ctx->tm.k_nmmd_init_ms += t;     // k_init + k_measure + scan
        ctx->tm.k_measure_ms += 0.f;
        cudaEventElapsedTime(&t, ctx->ev[4], ctx->ev[5]); ctx->tm.k_plan_ms += t;
        cudaEventElapsedTime(&t, ctx->ev[5], ctx->ev[6]); ctx->tm.k_junction_ms += t;
        cudaEventElapsedTime(&t, ctx->ev[7], ctx->ev[8]); ctx->tm.k_emit_ms += t;
    }
#else
    (void)timed; (void)st;
    std::vector<uint8_t> ipool, opool;
    for (int64_t i = 0; i < n; i++) host_init_read(P, i, ipool);
    CK(S.init_pool.ensure(ipool.size() + 8)); memcpy(S.init_pool.p, ipool.data(), ipool.size());
    P.W.md_init_pool = (const uint8_t*)S.init_pool.p; S.W.md_init_pool = P.W.md_init_pool;
    memset(S.W.counters, 0xff, (size_t)n * 40);
    for (int64_t i = 0; i < n; i++) measure_read(P, i);
    S.W.ws_off[n] = 0; be_scan(ctx, st, S.scan_tmp, S.W.ws_off, n + 1, 0);
    CK(S.w_ws.ensure((size_t)(S.W.ws_off[n] + 8) * 8)); S.W.ws = (u64*)S.w_ws.p; P.W.ws = S.W.ws;
    for (int64_t i = 0; i < n; i++) {
        if (dry) { dry_read(P, i); dry_sizes(P, i); }
        else { plan_read(P, i); junction_read(P, i); sizes_read(P, i); }
    }
    int64_t* sc[4] = {S.W.o_seq, S.W.o_cig, S.W.o_te, S.W.o_jn};
    const int64_t bases[4] = {S.base_seq, S.base_cig, S.base_te, S.base_jn};
    for (int k = 0; k < 4; k++) { sc[k][n] = 0; be_scan(ctx, st, S.scan_tmp, sc[k], n + 1, bases[k]); tot[1 + k] = sc[k][n] - bases[k]; }
    S.tot_seq = tot[1]; S.tot_cig = tot[2]; S.tot_te = tot[3]; S.tot_jn = tot[4];
    CK(S.o_seq.ensure((size_t)S.tot_seq + 64)); CK(S.o_cig_op.ensure((size_t)S.tot_cig + 64)); CK(S.o_cig_len.ensure((size_t)S.tot_cig * 4 + 64));
    CK(S.o_te.ensure((size_t)S.tot_te * 16 + 64)); CK(S.o_jm.ensure((size_t)S.tot_jn + 64)); CK(S.o_ji.ensure((size_t)S.tot_jn * 8 + 64));
    OutDev O = make_out(S, nullptr, 0);
    for (int64_t i = 0; i < n; i++) host_emit_read(P, O, i, dry, opool);
    for (int64_t i = 0; i < n; i++) O.md_off[i] += S.base_md;
    CK(S.o_md_pool.ensure(opool.size() + 8)); memcpy(S.o_md_pool.p, opool.data(), opool.size()); S.md_used = (int64_t)opool.size();
    (void)cur;
#endif
    S.ran = true; S.ran_dry = dry != 0;
    return TC_OK;
}

// reserves the pinned host arrays for the batch rows [0, r1) and output positions up to the slot's end
static int host_reserve_all(tc_ctx* ctx, Slot& S, int64_t r0, int64_t n_total, bool dry, stream_t busy) {
    const size_t n1 = (size_t)n_total + 1; (void)r0;
    if (host_reserve(ctx, ctx->h_status, n1 * 4, 0, busy) || host_reserve(ctx, ctx->h_counters, n1 * 40, 0, busy) ||
        host_reserve(ctx, ctx->h_te_off, n1 * 8, 0, busy) ||
        host_reserve(ctx, ctx->h_te, (size_t)(S.base_te + S.tot_te) * 16 + 64, (size_t)S.base_te * 16, busy)) return TC_E_CUDA;
    if (dry) return 0;
    if (host_reserve(ctx, ctx->h_nm, n1 * 4, 0, busy) || host_reserve(ctx, ctx->h_seq_off, n1 * 8, 0, busy) || host_reserve(ctx, ctx->h_seq_len, n1 * 4, 0, busy) ||
        host_reserve(ctx, ctx->h_cig_off, n1 * 8, 0, busy) || host_reserve(ctx, ctx->h_md_off, n1 * 8, 0, busy) || host_reserve(ctx, ctx->h_md_len, n1 * 4, 0, busy) ||
        host_reserve(ctx, ctx->h_jn_off, n1 * 8, 0, busy) ||
        host_reserve(ctx, ctx->h_seq, (size_t)(S.base_seq + S.tot_seq) + 64, (size_t)S.base_seq, busy) ||
        host_reserve(ctx, ctx->h_cig_op, (size_t)(S.base_cig + S.tot_cig) + 64, (size_t)S.base_cig, busy) ||
        host_reserve(ctx, ctx->h_cig_len, (size_t)(S.base_cig + S.tot_cig) * 4 + 64, (size_t)S.base_cig * 4, busy) ||
        host_reserve(ctx, ctx->h_md, (size_t)(S.base_md + S.md_used) + 64, (size_t)S.base_md, busy) ||
        host_reserve(ctx, ctx->h_jm, (size_t)(S.base_jn + S.tot_jn) + 64, (size_t)S.base_jn, busy) ||
        host_reserve(ctx, ctx->h_ji, (size_t)(S.base_jn + S.tot_jn) * 8 + 64, (size_t)S.base_jn * 8, busy)) return TC_E_CUDA;
    return 0;
}

// new SEQ rows of a slot -> host; packed to 4 bits first when the batch came in that form
static int seq_download(tc_ctx* ctx, Slot& S, stream_t st) {
    if (!ctx->packed) return be_d2h(ctx, st, ctx->h_seq, (size_t)S.base_seq, S.o_seq.p, (size_t)S.tot_seq);
    if (S.tot_seq == 0) return 0;
    CK(S.o_seqp.ensure((size_t)S.tot_seq / 2 + 64));
#ifndef TC_HOSTSIM
    Alphabet A; memcpy(A.c, ctx->alphabet, 16);
    const int64_t n16 = S.tot_seq / 16; int64_t nb = (n16 + 255) / 256; int g = (int)(nb < 148 * 16 ? nb : 148 * 16);
    k_seq_pack<<<g, 256, 0, st>>>(n16, (const uint4*)S.o_seq.p, (uint2*)S.o_seqp.p, A); ctx->tm.launches++;
    CK(cudaGetLastError());
#else
    uint8_t inv[256] = {0}; for (int k = 0; k < ctx->n_alpha; k++) inv[ctx->alphabet[k]] = (uint8_t)k;
    const uint8_t* a = (const uint8_t*)S.o_seq.p; uint8_t* o = (uint8_t*)S.o_seqp.p;
    for (int64_t j = 0; j < S.tot_seq / 2; j++) o[j] = (uint8_t)(inv[a[2 * j]] | (inv[a[2 * j + 1]] << 4));
#endif
    return be_d2h(ctx, st, ctx->h_seq, (size_t)S.base_seq / 2, S.o_seqp.p, (size_t)S.tot_seq / 2);
}

// enqueues the device-to-host copies of a slot's results into the batch-wide pinned arrays
static int slot_download(tc_ctx* ctx, Slot& S, int64_t r0, stream_t st) {
    const int64_t n = S.n; const size_t n1 = (size_t)n + 1;
    if (n == 0) return 0;
    if (be_d2h(ctx, st, ctx->h_status, (size_t)r0 * 4, S.W.status, (size_t)n * 4) || be_d2h(ctx, st, ctx->h_counters, (size_t)r0 * 40, S.W.counters, (size_t)n * 40) ||
        be_d2h(ctx, st, ctx->h_te_off, (size_t)r0 * 8, S.W.o_te, n1 * 8) || be_d2h(ctx, st, ctx->h_te, (size_t)S.base_te * 16, S.o_te.p, (size_t)S.tot_te * 16)) return TC_E_CUDA;
    if (S.ran_dry) return 0;
    if (be_d2h(ctx, st, ctx->h_nm, (size_t)r0 * 4, S.o_nm.p, (size_t)n * 4) || be_d2h(ctx, st, ctx->h_seq_off, (size_t)r0 * 8, S.W.o_seq, n1 * 8) ||
        be_d2h(ctx, st, ctx->h_seq_len, (size_t)r0 * 4, S.W.seq_len_out, (size_t)n * 4) || seq_download(ctx, S, st) ||
        be_d2h(ctx, st, ctx->h_cig_off, (size_t)r0 * 8, S.W.o_cig, n1 * 8) || be_d2h(ctx, st, ctx->h_cig_op, (size_t)S.base_cig, S.o_cig_op.p, (size_t)S.tot_cig) ||
        be_d2h(ctx, st, ctx->h_cig_len, (size_t)S.base_cig * 4, S.o_cig_len.p, (size_t)S.tot_cig * 4) ||
        be_d2h(ctx, st, ctx->h_md_off, (size_t)r0 * 8, S.o_md_off.p, (size_t)n * 8) || be_d2h(ctx, st, ctx->h_md_len, (size_t)r0 * 4, S.o_md_len.p, (size_t)n * 4) ||
        be_d2h(ctx, st, ctx->h_md, (size_t)S.base_md, S.o_md_pool.p, (size_t)S.md_used) || be_d2h(ctx, st, ctx->h_jn_off, (size_t)r0 * 8, S.W.o_jn, n1 * 8) ||
        be_d2h(ctx, st, ctx->h_jm, (size_t)S.base_jn, S.o_jm.p, (size_t)S.tot_jn) || be_d2h(ctx, st, ctx->h_ji, (size_t)S.base_jn * 8, S.o_ji.p, (size_t)S.tot_jn * 8)) return TC_E_CUDA;
    return 0;
}

static void fill_out(tc_ctx* ctx, tc_batch_out* out, int64_t n) {
    memset(out, 0, sizeof(*out));
    out->n_reads = n;
    out->status = (const uint32_t*)ctx->h_status.p; out->counters = (const int32_t*)ctx->h_counters.p; out->nm = (const int32_t*)ctx->h_nm.p;
    out->seq_off = (const int64_t*)ctx->h_seq_off.p; out->seq_len = (const int32_t*)ctx->h_seq_len.p; out->seq = (const uint8_t*)ctx->h_seq.p;
    out->cig_off = (const int64_t*)ctx->h_cig_off.p; out->cig_op = (const uint8_t*)ctx->h_cig_op.p; out->cig_len = (const int32_t*)ctx->h_cig_len.p;
    out->md_off = (const int64_t*)ctx->h_md_off.p; out->md_len = (const int32_t*)ctx->h_md_len.p; out->md = (const uint8_t*)ctx->h_md.p;
    out->jn_off = (const int64_t*)ctx->h_jn_off.p; out->jm = (const int8_t*)ctx->h_jm.p; out->ji = (const int32_t*)ctx->h_ji.p;
    out->te_off = (const int64_t*)ctx->h_te_off.p; out->te = (const tc_te*)ctx->h_te.p;
    out->seq_bits = ctx->packed ? 4 : 8;
}

// notes the SEQ form of the batch; a 4-bit batch needs an alphabet that holds every genome byte
static int check_packed(tc_ctx* ctx, const tc_batch_in* in) {
    ctx->packed = in->seq_poff != nullptr;
    if (!ctx->packed) return 0;
    bool have[256] = {false};
    for (int k = 0; k < ctx->n_alpha; k++) have[ctx->alphabet[k]] = true;
    for (int b = 0; b < 256; b++) if (ctx->genome_byte[b] && !have[b]) {
        ctx->err = "4-bit SEQ: the alphabet lacks genome byte " + std::to_string(b) + " (tc_ctx_set_alphabet)"; return 1;
    }
    return 0;
}

static void reset_bases(Slot& S) { S.base_seq = S.base_cig = S.base_te = S.base_jn = S.base_md = 0; }

extern "C" {

int tc_batch_upload(tc_ctx* ctx, const tc_batch_in* in) {
    if (!ctx || !in || in->n_reads < 0) return TC_E_ARG;
    if (!ctx->have_genome) { ctx->err = "tc_batch_upload: load a genome first"; return TC_E_STATE; }
    if (in->n_reads >= (1ll << 31) - 2) { ctx->err = "tc_batch_upload: batch too large"; return TC_E_ARG; }
#ifndef TC_HOSTSIM
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
#endif
    ctx->tm.h2d_bytes = 0;
    ctx->n = in->n_reads; ctx->ran = false;
    if (check_packed(ctx, in)) return TC_E_ARG;
    int rc = slot_upload(ctx, ctx->slot[0], in, 0, in->n_reads, main_stream(ctx));
#ifndef TC_HOSTSIM
    if (!rc) CK(cudaEventRecord(ctx->ev[1], ctx->stream));
#endif
    return rc;
}

int tc_batch_run(tc_ctx* ctx, int dry) {
    if (!ctx) return TC_E_ARG;
    if (!ctx->have_genome) { ctx->err = "tc_batch_run: load a genome first"; return TC_E_STATE; }
#ifndef TC_HOSTSIM
    CK(cudaSetDevice(ctx->device));
#endif
    ctx->tm.launches = 0; ctx->tm.md_pool_retries = 0; ctx->tm.exact_nmmd_reads = 0;
    ctx->tm.kernels_ms = ctx->tm.k_nmmd_init_ms = ctx->tm.k_measure_ms = ctx->tm.k_plan_ms = ctx->tm.k_junction_ms = ctx->tm.k_emit_ms = 0.f;
    reset_bases(ctx->slot[0]);
    int rc = slot_run(ctx, ctx->slot[0], dry, main_stream(ctx), true);
    if (!rc) { ctx->ran = true; ctx->ran_dry = dry != 0; ctx->piped = false; }
    return rc;
}

int tc_batch_download(tc_ctx* ctx, tc_batch_out* out) {
    if (!ctx || !out) return TC_E_ARG;
    if (!ctx->ran) { ctx->err = "tc_batch_download: nothing has run"; return TC_E_STATE; }
    const int64_t n = ctx->n;
    if (!ctx->piped) {
#ifndef TC_HOSTSIM
        CK(cudaSetDevice(ctx->device));
        CK(cudaEventRecord(ctx->ev[10], ctx->stream));
#endif
        ctx->tm.d2h_bytes = 0;
        Slot& S = ctx->slot[0];
        if (host_reserve_all(ctx, S, 0, n, ctx->ran_dry, main_stream(ctx)) || slot_download(ctx, S, 0, main_stream(ctx))) return TC_E_CUDA;
#ifndef TC_HOSTSIM
        CK(cudaEventRecord(ctx->ev[11], ctx->stream));
#endif
        if (be_sync(ctx, main_stream(ctx))) return TC_E_CUDA;
#ifndef TC_HOSTSIM
        cudaEventElapsedTime(&ctx->tm.d2h_ms, ctx->ev[10], ctx->ev[11]);
        cudaEventElapsedTime(&ctx->tm.h2d_ms, ctx->ev[0], ctx->ev[1]);
#endif
    }
    fill_out(ctx, out, n);
    return TC_OK;
}

// host columns in -> host columns out.  Large batches are cut into chunks of chunk_reads
// reads and pipelined over three streams: upload of chunk c+1, kernels of chunk c and download
// of chunk c-1 overlap (PCIe is full duplex); two slots alternate.
static int run_all(tc_ctx* ctx, const tc_batch_in* in, tc_batch_out* out, int dry) {
    if (!ctx || !in || !out || in->n_reads < 0) return TC_E_ARG;
    if (!ctx->have_genome) { ctx->err = "load a genome first"; return TC_E_STATE; }
    const int64_t n = in->n_reads;
#ifndef TC_HOSTSIM
    const int64_t CH = ctx->chunk_reads;
    if (CH > 0 && n > CH + CH / 2) {
        CK(cudaSetDevice(ctx->device));
        ctx->tm.h2d_bytes = ctx->tm.d2h_bytes = 0; ctx->tm.launches = 0; ctx->tm.md_pool_retries = 0; ctx->tm.exact_nmmd_reads = 0;
        ctx->tm.kernels_ms = ctx->tm.k_nmmd_init_ms = ctx->tm.k_measure_ms = ctx->tm.k_plan_ms = ctx->tm.k_junction_ms = ctx->tm.k_emit_ms = 0.f;
        ctx->n = n; ctx->ran = false;
        if (check_packed(ctx, in)) return TC_E_ARG;
        const int64_t nchunk = (n + CH - 1) / CH;
        int64_t bs = 0, bc = 0, bt = 0, bj = 0, bm = 0;
        CK(cudaEventRecord(ctx->ev[0], ctx->stream));
        const bool trace = getenv("TC_PIPE_TRACE") != nullptr;        // per-chunk stream timestamps on stderr (debugging aid)
        std::vector<cudaEvent_t> tev;
        auto mark = [&](cudaStream_t st) { if (trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); tev.push_back(e); } };
        mark(ctx->s_in);
        if (slot_upload(ctx, ctx->slot[0], in, 0, CH < n ? CH : n, ctx->s_in)) return TC_E_CUDA;
        CK(cudaEventRecord(ctx->slot[0].ev_in, ctx->s_in));
        mark(ctx->s_in);
        for (int64_t c = 0; c < nchunk; c++) {
            Slot& S = ctx->slot[c & 1];
            const int64_t r0 = c * CH;
            if (c + 1 < nchunk) {                                   // next upload: its slot's kernels (chunk c-1) are done
                Slot& N = ctx->slot[(c + 1) & 1];
                const int64_t a = (c + 1) * CH, b = a + CH < n ? a + CH : n;
                mark(ctx->s_in);
                if (slot_upload(ctx, N, in, a, b, ctx->s_in)) return TC_E_CUDA;
                CK(cudaEventRecord(N.ev_in, ctx->s_in));
                mark(ctx->s_in);
            }
            CK(cudaStreamWaitEvent(ctx->stream, S.ev_in, 0));
            if (c >= 2) CK(cudaStreamWaitEvent(ctx->stream, S.ev_out, 0));   // the slot's previous results have left the device
            S.base_seq = bs; S.base_cig = bc; S.base_te = bt; S.base_jn = bj; S.base_md = bm;
            mark(ctx->stream);
            int rc = slot_run(ctx, S, dry, ctx->stream, false);
            if (rc) return rc;
            CK(cudaEventRecord(S.ev_cmp, ctx->stream));
            mark(ctx->stream);
            if (host_reserve_all(ctx, S, r0, n, dry != 0, ctx->s_out)) return TC_E_CUDA;
            CK(cudaStreamWaitEvent(ctx->s_out, S.ev_cmp, 0));
            mark(ctx->s_out);
            if (slot_download(ctx, S, r0, ctx->s_out)) return TC_E_CUDA;
            CK(cudaEventRecord(S.ev_out, ctx->s_out));
            mark(ctx->s_out);
            bs += S.tot_seq; bc += S.tot_cig; bt += S.tot_te; bj += S.tot_jn; bm += S.md_used;
        }
        CK(cudaStreamSynchronize(ctx->s_out));
        CK(cudaEventRecord(ctx->ev[1], ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        cudaEventElapsedTime(&ctx->tm.kernels_ms, ctx->ev[0], ctx->ev[1]);
        if (trace) {      // order of marks: up0 b/e, then per chunk: [up(c+1) b/e], run b/e, down b/e
            size_t k = 0; auto t = [&](size_t j) { float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[0], tev[j]); return ms; };
            fprintf(stderr, "pipe: upload 0: %.2f-%.2f\n", t(0), t(1)); k = 2;
            for (int64_t c = 0; c < nchunk; c++) {
                if (c + 1 < nchunk) { fprintf(stderr, "pipe: upload %lld: %.2f-%.2f\n", (long long)(c + 1), t(k), t(k + 1)); k += 2; }
                fprintf(stderr, "pipe: run %lld: %.2f-%.2f   download %lld: %.2f-%.2f\n", (long long)c, t(k), t(k + 1), (long long)c, t(k + 2), t(k + 3)); k += 4;
            }
            for (auto e : tev) cudaEventDestroy(e);
        }
        ctx->ran = true; ctx->ran_dry = dry != 0; ctx->piped = true;
        fill_out(ctx, out, n);
        return TC_OK;
    }
#endif
    int rc = tc_batch_upload(ctx, in);
    if (rc) return rc;
    rc = tc_batch_run(ctx, dry);
    if (rc) return rc;
    return tc_batch_download(ctx, out);
}
int tc_correct_batch(tc_ctx* ctx, const tc_batch_in* in, tc_batch_out* out) { return run_all(ctx, in, out, 0); }
int tc_dryrun_batch(tc_ctx* ctx, const tc_batch_in* in, tc_batch_out* out) { return run_all(ctx, in, out, 1); }

int tc_ctx_set_alphabet(tc_ctx* ctx, const uint8_t* alphabet, int32_t n) {
    if (!ctx || !alphabet || n < 1 || n > 16) return TC_E_ARG;
    for (int a = 0; a < n; a++) for (int b = a + 1; b < n; b++) if (alphabet[a] == alphabet[b]) { ctx->err = "tc_ctx_set_alphabet: duplicate byte"; return TC_E_ARG; }
    memset(ctx->alphabet, 0, 16); memcpy(ctx->alphabet, alphabet, (size_t)n); ctx->n_alpha = n;
    return TC_OK;
}

int tc_ctx_set_pipeline(tc_ctx* ctx, int64_t chunk_reads) {
    if (!ctx || chunk_reads < 0) return TC_E_ARG;
    ctx->chunk_reads = chunk_reads;
    return TC_OK;
}

int tc_get_timing(tc_ctx* ctx, tc_timing* t) { if (!ctx || !t) return TC_E_ARG; *t = ctx->tm; return TC_OK; }

}  // extern "C"

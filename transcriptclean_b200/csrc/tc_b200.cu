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
    uint8_t* seq; uint8_t* cig_op; int32_t* cig_len; tc_te* te; int8_t* jm; int32_t* ji;
};
struct InitDev { uint8_t* pool; unsigned long long* cursor; int64_t cap; };

TC_HD int ndigits(uint32_t v) {
    int n = 1;
    while (v >= 10) { v /= 10; n++; }
    return n;
}
TC_HD void write_dec(uint8_t* dst, uint32_t v, int nd) {
    for (int k = nd - 1; k >= 0; k--) { dst[k] = (uint8_t)('0' + v % 10); v /= 10; }
}

struct CigSrc { const uint8_t* op; const int32_t* len; const u64* packed; int n; };
TC_HD void cig_get(const CigSrc& c, int k, int& op, int32_t& len) {
    if (c.packed) { op = cg_op(c.packed[k]); len = cg_len(c.packed[k]); }
    else { op = c.op[k]; len = c.len[k]; }
}

#ifndef TC_HOSTSIM
// ================================================================================== CUDA

#define FULL 0xffffffffu
#define WARPS_PER_BLOCK 8
#define EMIT_CAP 4096            // reads whose new SEQ fits are staged in shared memory
#define EMIT_MAXBP 192
#define EMIT_MAXGI 192

struct __align__(16) EmitSmem {  // per warp
    uint8_t seq[EMIT_CAP + 32];
    uint16_t bp_pos[EMIT_MAXBP]; int16_t bp_shift[EMIT_MAXBP];     // starts of input-SEQ slices: output offset, src-dst shift
    uint16_t gi_out[EMIT_MAXGI]; uint16_t gi_len[EMIT_MAXGI]; int32_t gi_src[EMIT_MAXGI];   // genome slices
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
                                             uint8_t* pool, int& nm, int& len, int64_t& off, bool& none) {
    NmMd r = warp_nmmd<false, FAST>(G, seq, lut, cs, chrom, pos, lane, 0);
    none = r.none; nm = r.nm; len = r.none ? 0 : r.len;
    off = pool_grab(cursor, cap, len, lane);
    if (off >= 0 && len > 0) {
        if (r.ntok == 0) { if (lane == 0) write_dec(pool + off, (uint32_t)r.carry, len); }
        else warp_nmmd<true, FAST>(G, seq, lut, cs, chrom, pos, lane, pool + off);
    }
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) k_init(Params P, InitDev I, int32_t* nm_init) {
    const int lane = threadIdx.x & 31;
    const int64_t nw = (int64_t)gridDim.x * WARPS_PER_BLOCK;
    for (int64_t i = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5); i < P.B.n; i += nw) {
        const int cls = read_class(P.B.flag[i], P.B.chrom[i]);
        uint32_t st = (uint32_t)cls;
        const uint8_t tags = P.B.tags[i];
        if (cls == 0 && (!(tags & TAG_NM) || !(tags & TAG_MD))) {
            CigSrc cs; cs.packed = 0; cs.op = P.B.cig_op + P.B.cig_off[i]; cs.len = P.B.cig_len + P.B.cig_off[i];
            cs.n = (int)(P.B.cig_off[i + 1] - P.B.cig_off[i]);
            int nm, len; int64_t off; bool none;
            nmmd_to_pool<false>(P.G, P.B.seq + P.B.seq_off[i], 0, cs, P.B.chrom[i], P.B.pos[i], lane, I.cursor, I.cap, I.pool, nm, len, off, none);
            st |= TC_ST_NMMD_DEVICE;
            if (none) { st |= TC_ST_NMMD_NONE; nm = 0; }
            if (lane == 0) { P.W.md_init_off[i] = off < 0 ? 0 : off; P.W.md_init_len[i] = off < 0 ? 0 : len; nm_init[i] = nm; }
        }
        if (lane == 0) P.W.status[i] = st;
    }
}

__global__ void k_measure(Params P) { int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < P.B.n) measure_read(P, i); }
__global__ void __launch_bounds__(128, 8) k_plan(Params P) { int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < P.B.n) plan_read(P, i); }
__global__ void __launch_bounds__(128, 8) k_junction(Params P) { int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < P.B.n) junction_read(P, i); }
__global__ void k_sizes(Params P) { int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < P.B.n) sizes_read(P, i); }
__global__ void k_dry(Params P) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P.B.n) { dry_read(P, i); dry_sizes(P, i); }
}

// Output stage: one warp per read.  `seq_words` = number of 32-bit words that may be read from
// the input SEQ array (the buffer is padded).
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, 4) k_emit(Params P, OutDev O, int dry, int64_t seq_words) {
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
    for (int64_t i = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5); i < B.n; i += nw) {
        uint32_t st = W.status[i];
        if ((st & TC_ST_CLASS_MASK) != 0) { if (lane == 0) { O.md_off[i] = 0; O.md_len[i] = 0; O.nm[i] = 0; } continue; }
        // the read's plan row: one word per lane, fields broadcast by shuffle
        const int pw = lane < 16 ? W.plan[i].w[lane] : 0;
#define PF(k) __shfl_sync(FULL, pw, (k))
        const int pflags = PF(EP_FLAGS), cig_cap = PF(EP_CIGCAP), seg_cap = PF(EP_SEGCAP), nm_extra = PF(EP_NMEXTRA);
        const int cb = ((pflags >> 2) & 3) - 1, sb = ((pflags >> 4) & 3) - 1;
        u64* const ws = W.ws + W.ws_off[i];
        const tc_te* const ws_te = reinterpret_cast<const tc_te*>(ws + 3ll * cig_cap + 3ll * seg_cap);
        // ---- TE rows: stage order (mismatch, insertion, deletion, junction), positional inside
        const int nte = PF(EP_NTE);
        if (nte > 0) {
            tc_te* dst = O.te + W.o_te[i];
            if (dry) { for (int k = lane; k < nte; k += 32) dst[k] = ws_te[k]; }
            else {
                int base[4]; base[0] = 0; base[1] = PF(EP_TEMM); base[2] = base[1] + PF(EP_TEINS); base[3] = base[2] + PF(EP_TEDEL);
                for (int c = 0; c < nte; c += 32) {
                    const int k = c + lane; const bool valid = k < nte;
                    tc_te r; int t = -1;
                    if (valid) { r = ws_te[k]; t = r.type; }
#pragma unroll
                    for (int ty = 0; ty < 4; ty++) {
                        const unsigned m = __ballot_sync(FULL, t == ty);
                        if (t == ty) dst[base[ty] + __popc(m & lt)] = r;
                        base[ty] += __popc(m);
                    }
                }
            }
        }
        if (dry) { if (lane == 0) { O.md_off[i] = 0; O.md_len[i] = 0; O.nm[i] = 0; } continue; }
        // ---- CIGAR
        CigSrc cs; cs.n = PF(EP_NCIG);
        if (cb < 0) { cs.packed = 0; cs.op = B.cig_op + B.cig_off[i]; cs.len = B.cig_len + B.cig_off[i]; }
        else { cs.packed = ws + (int64_t)cb * cig_cap; cs.op = 0; cs.len = 0; }
        {
            const int64_t oc = W.o_cig[i]; uint8_t* dop = O.cig_op + oc; int32_t* dl = O.cig_len + oc;
            for (int k = lane; k < cs.n; k += 32) { int op; int32_t l; cig_get(cs, k, op, l); dop[k] = (uint8_t)op; dl[k] = l; }
        }
        // ---- jM / jI snapshot
        const int nj = PF(EP_NJN);
        if (nj > 0) {
            const int64_t oj = W.o_jn[i]; int8_t* djm = O.jm + oj; int32_t* dji = O.ji + 2 * oj;
            const int32_t* jn = reinterpret_cast<const int32_t*>(ws + 3ll * cig_cap + 3ll * seg_cap + 2ll * PF(EP_TECAP));
            for (int k = lane; k < nj; k += 32) {
                const int32_t* r = jn + k * JN_STRIDE;
                djm[k] = (int8_t)r[7]; dji[2 * k] = r[5]; dji[2 * k + 1] = r[6];
            }
        }
        // ---- SEQ
        const int64_t in_base = B.seq_off[i];
        const uint8_t* in_seq = B.seq + in_base;
        const int64_t Lq = B.seq_off[i + 1] - in_base;
        const int out_len = PF(EP_OUTLEN);
        uint8_t* out_seq = O.seq + W.o_seq[i];
        const int chrom = B.chrom[i]; const int64_t pos = B.pos[i];
        const int64_t gbase = P.G.chrom_off[chrom] - 1 + pos;            // plane index of the read's first reference base
        const u64* const ws_seg = ws + 3ll * cig_cap + (int64_t)(sb < 0 ? 0 : sb) * seg_cap;
        const int ns = PF(EP_NSEG);
        bool fast = out_len <= EMIT_CAP;
        int nbp = 0, ngi = 0;
        if (fast) {
            // (A) segment list -> starts of input slices (output offset, shift) + genome slices
            if (sb < 0) { if (lane == 0) { SM.bp_pos[0] = 0; SM.bp_shift[0] = 0; } nbp = 1; }
            else {
                const u64* segs = ws_seg; int obase = 0; bool bad = false;
                for (int s0 = 0; s0 < ns; s0 += 32) {
                    const int s = s0 + lane; const bool valid = s < ns;
                    const u64 sg = valid ? segs[s] : 0ull;
                    const int len = valid ? (int)seg_len(sg) : 0; const int64_t src = seg_src(sg); const int kind = seg_kind(sg);
                    int total; const int ostart = obase + warp_excl_scan(len, lane, total);
                    const bool isR = valid && kind == 0, isG = valid && kind == 1;
                    const unsigned mR = __ballot_sync(FULL, isR), mG = __ballot_sync(FULL, isG);
                    if (isR) {
                        const int idx = nbp + __popc(mR & lt); const int64_t shift = src - ostart;
                        if (idx < EMIT_MAXBP && shift >= -32768 && shift <= 32767) { SM.bp_pos[idx] = (uint16_t)ostart; SM.bp_shift[idx] = (int16_t)shift; }
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
        if (fast) {
            __syncwarp();
            bool read_exc = false;
            for (int k = lane; k < ngi; k += 32) { const int64_t src = gbase + SM.gi_src[k]; read_exc |= genome_blk_has_exc(P.G, src) | genome_blk_has_exc(P.G, src + SM.gi_len[k] - 1) | (SM.gi_len[k] > 256); }
            read_exc = __any_sync(FULL, read_exc);
            const int nchunk = (out_len + 15) >> 4;
            int top = 1; while (top * 2 < nbp) top <<= 1; if (nbp <= 1) top = 0;
            // (B) bulk copy, 16 output bytes per lane, with the shift in force at the chunk start
            for (int c = lane; c < nchunk; c += 32) {
                const int p = c << 4; int k = 0;
                for (int step = top; step; step >>= 1) { const int t = k + step; if (t < nbp && (int)SM.bp_pos[t] <= p) k = t; }
                int64_t a = in_base + p + (nbp ? (int)SM.bp_shift[k] : 0);
                if (a < 0) a = 0;
                int64_t wi = a >> 2; const int sh = (int)(a & 3) * 8;
                if (wi + 4 >= seq_words) wi = seq_words - 5;
                const uint32_t x0 = S32[wi], x1 = S32[wi + 1], x2 = S32[wi + 2], x3 = S32[wi + 3], x4 = S32[wi + 4];
                uint4 o; o.x = __funnelshift_r(x0, x1, sh); o.y = __funnelshift_r(x1, x2, sh); o.z = __funnelshift_r(x2, x3, sh); o.w = __funnelshift_r(x3, x4, sh);
                *reinterpret_cast<uint4*>(SM.seq + p) = o;
            }
            __syncwarp();
            // (C) the part of a chunk that follows a slice start inside it: same 5-word gather, byte stores
            for (int k = lane; k < nbp; k += 32) {
                const int b = SM.bp_pos[k];
                if (b & 15) {
                    int end = (b | 15) + 1; if (end > out_len) end = out_len;
                    if (k + 1 < nbp && (int)SM.bp_pos[k + 1] < end) end = SM.bp_pos[k + 1];
                    int64_t a = in_base + b + (int)SM.bp_shift[k];
                    int64_t wi = a >> 2; const int sh = (int)(a & 3) * 8;
                    if (wi + 4 >= seq_words) wi = seq_words - 5;
                    const uint32_t x0 = S32[wi], x1 = S32[wi + 1], x2 = S32[wi + 2], x3 = S32[wi + 3], x4 = S32[wi + 4];
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
        if (pflags & 1) {
            __syncwarp();
            int nm, len; int64_t off; bool none;
            if (fast) nmmd_to_pool<true>(P.G, SM.seq, lut, cs, chrom, pos, lane, O.md_cursor, O.md_cap, O.md_pool, nm, len, off, none);
            else nmmd_to_pool<false>(P.G, out_seq, lut, cs, chrom, pos, lane, O.md_cursor, O.md_cap, O.md_pool, nm, len, off, none);
            st = (st & ~(uint32_t)TC_ST_NMMD_NONE) | TC_ST_NMMD_DEVICE;
            if (none) st |= TC_ST_NMMD_NONE;
            if (lane == 0) { O.nm[i] = nm + nm_extra; O.md_off[i] = off < 0 ? 0 : off; O.md_len[i] = off < 0 ? 0 : len; W.status[i] = st; }
        } else if (st & TC_ST_NMMD_DEVICE) {                           // generated at init and never refreshed
            const int len = W.md_init_len[i];
            const int64_t off = pool_grab(O.md_cursor, O.md_cap, len, lane);
            if (off >= 0) { const uint8_t* s = W.md_init_pool + W.md_init_off[i]; for (int j = lane; j < len; j += 32) O.md_pool[off + j] = s[j]; }
            if (lane == 0) { O.nm[i] = W.nm_init[i]; O.md_off[i] = off < 0 ? 0 : off; O.md_len[i] = off < 0 ? 0 : len; }
        } else if (lane == 0) { O.nm[i] = 0; O.md_off[i] = 0; O.md_len[i] = 0; }
        __syncwarp();
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
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
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

struct tc_ctx {
    int device = 0;
    std::string err;
    tc_options opt;
#ifndef TC_HOSTSIM
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[12];
#endif
    bool have_genome = false;
    // tables
    DevBuf g2, lo1, excblk, chrom_off, chrom_len, exc_start, exc_end, exc_byte;
    DevBuf don_s, acc_s, don_u, acc_u, ann_k1, ann_k2;
    DevBuf snp_key, snp_alt, ins_key, ins_size, ins_aoff, ins_bytes, del_key, del_size;
    GenomeDev G; SjDev S; VarDev V;
    std::vector<int64_t> h_chrom_len;
    // batch input on the device
    DevBuf b_flag, b_chrom, b_pos, b_tags, b_seq_off, b_seq, b_cig_off, b_cig_op, b_cig_len, b_md_off, b_md, b_ji_off, b_ji;
    BatchDev B; int64_t n = 0; int64_t n_seq = 0, n_cig = 0, n_md = 0, n_ji = 0;
    // work
    DevBuf w_status, w_counters, w_md_init_off, w_md_init_len, w_nm_init, w_n_mdtok, w_n_jn, w_mflags,
        w_cig_cap, w_seg_cap, w_te_cap, w_ws_off, w_ws, w_cig_buf, w_n_cig_cur, w_seg_buf, w_n_seg_cur, w_n_te,
        w_te_cnt, w_nm_extra, w_refresh, w_o_seq, w_o_cig, w_o_te, w_o_jn, w_seq_len, w_plan, init_pool, cursors, scan_tmp;
    WorkDev W;
    // dense outputs on the device
    DevBuf o_nm, o_md_off, o_md_len, o_md_pool, o_seq, o_cig_op, o_cig_len, o_te, o_jm, o_ji;
    int64_t tot_seq = 0, tot_cig = 0, tot_te = 0, tot_jn = 0, md_used = 0;
    bool ran = false, ran_dry = false;
    // pinned host outputs
    HostBuf h_status, h_counters, h_nm, h_seq_off, h_seq, h_cig_off, h_cig_op, h_cig_len, h_md_off, h_md_len, h_md,
        h_jn_off, h_jm, h_ji, h_te_off, h_te, h_tot, h_seq_len;
    tc_timing tm;
};

static std::string g_create_err;

#ifndef TC_HOSTSIM
static int be_h2d(tc_ctx* ctx, DevBuf& d, const void* src, size_t bytes) {
    CK(d.ensure(bytes ? bytes : 8));
    if (bytes) CK(cudaMemcpyAsync(d.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ctx->tm.h2d_bytes += (int64_t)bytes;
    return 0;
}
static int be_d2h(tc_ctx* ctx, HostBuf& h, const void* src, size_t bytes) {
    CK(h.ensure(bytes ? bytes : 8));
    if (bytes) CK(cudaMemcpyAsync(h.p, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->tm.d2h_bytes += (int64_t)bytes;
    return 0;
}
static int be_sync(tc_ctx* ctx) { CK(cudaStreamSynchronize(ctx->stream)); return 0; }
static int be_zero(tc_ctx* ctx, void* p, size_t bytes) { CK(cudaMemsetAsync(p, 0, bytes, ctx->stream)); return 0; }
// in-place exclusive prefix sum over n1 int64 values
static int be_scan(tc_ctx* ctx, int64_t* a, int64_t n1) {
    size_t tmp = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, a, a, (int)n1, ctx->stream));
    CK(ctx->scan_tmp.ensure(tmp));
    CK(cub::DeviceScan::ExclusiveSum(ctx->scan_tmp.p, tmp, a, a, (int)n1, ctx->stream));
    ctx->tm.launches += 2;
    return 0;
}
static int be_read_i64(tc_ctx* ctx, const int64_t* dev, int64_t* host) {
    CK(cudaMemcpyAsync(host, dev, 8, cudaMemcpyDeviceToHost, ctx->stream));
    return 0;
}
#else
static int be_h2d(tc_ctx* ctx, DevBuf& d, const void* src, size_t bytes) {
    CK(d.ensure(bytes ? bytes : 8)); if (bytes) memcpy(d.p, src, bytes); ctx->tm.h2d_bytes += (int64_t)bytes; return 0;
}
static int be_d2h(tc_ctx* ctx, HostBuf& h, const void* src, size_t bytes) {
    CK(h.ensure(bytes ? bytes : 8)); if (bytes) memcpy(h.p, src, bytes); ctx->tm.d2h_bytes += (int64_t)bytes; return 0;
}
static int be_sync(tc_ctx*) { return 0; }
static int be_zero(tc_ctx*, void* p, size_t bytes) { memset(p, 0, bytes); return 0; }
static int be_scan(tc_ctx*, int64_t* a, int64_t n1) { int64_t s = 0; for (int64_t k = 0; k < n1; k++) { int64_t v = a[k]; a[k] = s; s += v; } return 0; }
static int be_read_i64(tc_ctx*, const int64_t* dev, int64_t* host) { *host = *dev; return 0; }
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
    for (int k = 0; k < 12 && e == cudaSuccess; k++) e = cudaEventCreate(&ctx->ev[k]);
    if (e != cudaSuccess) { g_create_err = std::string("tc_ctx_create: ") + cudaGetErrorString(e); delete ctx; return nullptr; }
#endif
    return ctx;
}

void tc_ctx_destroy(tc_ctx* ctx) {
    if (!ctx) return;
#ifndef TC_HOSTSIM
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
#endif
    DevBuf* d[] = {&ctx->g2, &ctx->lo1, &ctx->excblk, &ctx->chrom_off, &ctx->chrom_len, &ctx->exc_start, &ctx->exc_end, &ctx->exc_byte,
                   &ctx->don_s, &ctx->acc_s, &ctx->don_u, &ctx->acc_u, &ctx->ann_k1, &ctx->ann_k2, &ctx->snp_key, &ctx->snp_alt,
                   &ctx->ins_key, &ctx->ins_size, &ctx->ins_aoff, &ctx->ins_bytes, &ctx->del_key, &ctx->del_size,
                   &ctx->b_flag, &ctx->b_chrom, &ctx->b_pos, &ctx->b_tags, &ctx->b_seq_off, &ctx->b_seq, &ctx->b_cig_off, &ctx->b_cig_op,
                   &ctx->b_cig_len, &ctx->b_md_off, &ctx->b_md, &ctx->b_ji_off, &ctx->b_ji,
                   &ctx->w_status, &ctx->w_counters, &ctx->w_md_init_off, &ctx->w_md_init_len, &ctx->w_nm_init, &ctx->w_n_mdtok, &ctx->w_n_jn,
                   &ctx->w_mflags, &ctx->w_cig_cap, &ctx->w_seg_cap, &ctx->w_te_cap, &ctx->w_ws_off, &ctx->w_ws, &ctx->w_cig_buf, &ctx->w_n_cig_cur,
                   &ctx->w_seg_buf, &ctx->w_n_seg_cur, &ctx->w_n_te, &ctx->w_te_cnt, &ctx->w_nm_extra, &ctx->w_refresh, &ctx->w_o_seq, &ctx->w_o_cig,
                   &ctx->w_o_te, &ctx->w_o_jn, &ctx->w_seq_len, &ctx->w_plan, &ctx->init_pool, &ctx->cursors, &ctx->scan_tmp,
                   &ctx->o_nm, &ctx->o_md_off, &ctx->o_md_len, &ctx->o_md_pool, &ctx->o_seq, &ctx->o_cig_op, &ctx->o_cig_len, &ctx->o_te, &ctx->o_jm, &ctx->o_ji};
    for (DevBuf* b : d) b->release();
    HostBuf* h[] = {&ctx->h_status, &ctx->h_counters, &ctx->h_nm, &ctx->h_seq_off, &ctx->h_seq, &ctx->h_cig_off, &ctx->h_cig_op, &ctx->h_cig_len,
                    &ctx->h_md_off, &ctx->h_md_len, &ctx->h_md, &ctx->h_jn_off, &ctx->h_jm, &ctx->h_ji, &ctx->h_te_off, &ctx->h_te, &ctx->h_tot, &ctx->h_seq_len};
    for (HostBuf* b : h) b->release();
#ifndef TC_HOSTSIM
    for (int k = 0; k < 12; k++) cudaEventDestroy(ctx->ev[k]);
    cudaStreamDestroy(ctx->stream);
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

int tc_ctx_load_genome(tc_ctx* ctx, int32_t n_chrom, const int64_t* chrom_len, const int64_t* chrom_off,
                       int64_t n_pad, const uint32_t* bases2, const uint32_t* lower1, int64_t n_exc,
                       const int64_t* exc_start, const int64_t* exc_end, const uint8_t* exc_byte) {
    if (!ctx || n_chrom <= 0 || n_pad <= 0 || (n_pad & 63)) { if (ctx) ctx->err = "tc_ctx_load_genome: bad sizes"; return TC_E_ARG; }
#ifndef TC_HOSTSIM
    CK(cudaSetDevice(ctx->device));
#endif
    // block bitmap of literal runs (1 bit per 256 bases)
    int64_t nblk = (n_pad >> 8) + 2, nwords = (nblk + 31) / 32 + 1;
    std::vector<uint32_t> blk((size_t)nwords, 0u);
    for (int64_t k = 0; k < n_exc; k++) {
        if (exc_end[k] <= exc_start[k] || exc_start[k] < 0 || exc_end[k] > n_pad) { ctx->err = "tc_ctx_load_genome: bad literal run"; return TC_E_ARG; }
        for (int64_t b = exc_start[k] >> 8; b <= (exc_end[k] - 1) >> 8; b++) blk[(size_t)(b >> 5)] |= 1u << (b & 31);
    }
    if (be_h2d(ctx, ctx->g2, bases2, (size_t)(n_pad / 16) * 4)) return TC_E_CUDA;
    if (be_h2d(ctx, ctx->lo1, lower1, (size_t)(n_pad / 32) * 4)) return TC_E_CUDA;
    if (be_h2d(ctx, ctx->excblk, blk.data(), (size_t)nwords * 4)) return TC_E_CUDA;
    if (be_h2d(ctx, ctx->chrom_off, chrom_off, (size_t)n_chrom * 8)) return TC_E_CUDA;
    if (be_h2d(ctx, ctx->chrom_len, chrom_len, (size_t)n_chrom * 8)) return TC_E_CUDA;
    if (be_h2d(ctx, ctx->exc_start, exc_start, (size_t)n_exc * 8)) return TC_E_CUDA;
    if (be_h2d(ctx, ctx->exc_end, exc_end, (size_t)n_exc * 8)) return TC_E_CUDA;
    if (be_h2d(ctx, ctx->exc_byte, exc_byte, (size_t)n_exc)) return TC_E_CUDA;
    if (be_sync(ctx)) return TC_E_CUDA;
    GenomeDev& G = ctx->G;
    G.g2 = (const uint32_t*)ctx->g2.p; G.lo1 = (const uint32_t*)ctx->lo1.p; G.excblk = (const uint32_t*)ctx->excblk.p;
    G.chrom_off = (const int64_t*)ctx->chrom_off.p; G.chrom_len = (const int64_t*)ctx->chrom_len.p;
    G.exc_start = (const int64_t*)ctx->exc_start.p; G.exc_end = (const int64_t*)ctx->exc_end.p;
    G.exc_byte = (const uint8_t*)ctx->exc_byte.p; G.n_exc = n_exc; G.n_chrom = n_chrom;
    ctx->h_chrom_len.assign(chrom_len, chrom_len + n_chrom);
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
    if (be_h2d(ctx, ctx->don_s, don_s, (size_t)n_don_s * 8) || be_h2d(ctx, ctx->acc_s, acc_s, (size_t)n_acc_s * 8) ||
        be_h2d(ctx, ctx->don_u, don_u, (size_t)n_don_u * 8) || be_h2d(ctx, ctx->acc_u, acc_u, (size_t)n_acc_u * 8) ||
        be_h2d(ctx, ctx->ann_k1, k1, (size_t)n_annot * 8) || be_h2d(ctx, ctx->ann_k2, k2, (size_t)n_annot * 8) || be_sync(ctx)) return TC_E_CUDA;
    SjDev& S = ctx->S;
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
    int64_t nab = n_ins ? ins_aoff[n_ins] : 0;
    if (be_h2d(ctx, ctx->snp_key, snp_key, (size_t)n_snp * 8) || be_h2d(ctx, ctx->snp_alt, snp_alt, (size_t)n_snp) ||
        be_h2d(ctx, ctx->ins_key, ins_key, (size_t)n_ins * 8) || be_h2d(ctx, ctx->ins_size, ins_size, (size_t)n_ins * 4) ||
        be_h2d(ctx, ctx->ins_aoff, ins_aoff, (size_t)(n_ins ? n_ins + 1 : 0) * 8) || be_h2d(ctx, ctx->ins_bytes, ins_bytes, (size_t)nab) ||
        be_h2d(ctx, ctx->del_key, del_key, (size_t)n_del * 8) || be_h2d(ctx, ctx->del_size, del_size, (size_t)n_del * 4) || be_sync(ctx)) return TC_E_CUDA;
    VarDev& V = ctx->V;
    V.snp_key = (const u64*)ctx->snp_key.p; V.snp_alt = (const uint8_t*)ctx->snp_alt.p; V.n_snp = n_snp;
    V.ins_key = (const u64*)ctx->ins_key.p; V.ins_size = (const int32_t*)ctx->ins_size.p; V.ins_aoff = (const int64_t*)ctx->ins_aoff.p;
    V.ins_bytes = (const uint8_t*)ctx->ins_bytes.p; V.n_ins = n_ins;
    V.del_key = (const u64*)ctx->del_key.p; V.del_size = (const int32_t*)ctx->del_size.p; V.n_del = n_del;
    return TC_OK;
}

int tc_batch_upload(tc_ctx* ctx, const tc_batch_in* in) {
    if (!ctx || !in || in->n_reads < 0) return TC_E_ARG;
    if (!ctx->have_genome) { ctx->err = "tc_batch_upload: load a genome first"; return TC_E_STATE; }
    if (in->n_reads >= (1ll << 31) - 2) { ctx->err = "tc_batch_upload: batch too large"; return TC_E_ARG; }
#ifndef TC_HOSTSIM
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
#endif
    const int64_t n = in->n_reads;
    ctx->tm.h2d_bytes = 0;
    ctx->n = n;
    ctx->n_seq = n ? in->seq_off[n] : 0; ctx->n_cig = n ? in->cig_off[n] : 0;
    ctx->n_md = n ? in->md_off[n] : 0; ctx->n_ji = n ? in->ji_off[n] : 0;
    static const int64_t zero_off[1] = {0};
    if (be_h2d(ctx, ctx->b_flag, in->flag, (size_t)n * 4) || be_h2d(ctx, ctx->b_chrom, in->chrom, (size_t)n * 4) ||
        be_h2d(ctx, ctx->b_pos, in->pos, (size_t)n * 4) || be_h2d(ctx, ctx->b_tags, in->tags, (size_t)n) ||
        be_h2d(ctx, ctx->b_seq_off, n ? in->seq_off : zero_off, (size_t)(n + 1) * 8) || be_h2d(ctx, ctx->b_seq, in->seq, (size_t)ctx->n_seq) ||
        be_h2d(ctx, ctx->b_cig_off, n ? in->cig_off : zero_off, (size_t)(n + 1) * 8) || be_h2d(ctx, ctx->b_cig_op, in->cig_op, (size_t)ctx->n_cig) ||
        be_h2d(ctx, ctx->b_cig_len, in->cig_len, (size_t)ctx->n_cig * 4) ||
        be_h2d(ctx, ctx->b_md_off, n ? in->md_off : zero_off, (size_t)(n + 1) * 8) || be_h2d(ctx, ctx->b_md, in->md, (size_t)ctx->n_md) ||
        be_h2d(ctx, ctx->b_ji_off, n ? in->ji_off : zero_off, (size_t)(n + 1) * 8) || be_h2d(ctx, ctx->b_ji, in->ji, (size_t)ctx->n_ji * 4)) return TC_E_CUDA;
#ifndef TC_HOSTSIM
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
#endif
    BatchDev& B = ctx->B;
    B.n = n; B.flag = (const int32_t*)ctx->b_flag.p; B.chrom = (const int32_t*)ctx->b_chrom.p; B.pos = (const int32_t*)ctx->b_pos.p;
    B.tags = (const uint8_t*)ctx->b_tags.p; B.seq_off = (const int64_t*)ctx->b_seq_off.p; B.seq = (const uint8_t*)ctx->b_seq.p;
    B.cig_off = (const int64_t*)ctx->b_cig_off.p; B.cig_op = (const uint8_t*)ctx->b_cig_op.p; B.cig_len = (const int32_t*)ctx->b_cig_len.p;
    B.md_off = (const int64_t*)ctx->b_md_off.p; B.md = (const uint8_t*)ctx->b_md.p; B.ji_off = (const int64_t*)ctx->b_ji_off.p; B.ji = (const int32_t*)ctx->b_ji.p;
    ctx->ran = false;
    return TC_OK;
}

static int alloc_work(tc_ctx* ctx) {
    const size_t n = (size_t)ctx->n, n1 = n + 1;
    CK(ctx->w_status.ensure(n1 * 4)); CK(ctx->w_counters.ensure(n1 * 40)); CK(ctx->w_md_init_off.ensure(n1 * 8));
    CK(ctx->w_md_init_len.ensure(n1 * 4)); CK(ctx->w_nm_init.ensure(n1 * 4)); CK(ctx->w_n_mdtok.ensure(n1 * 4));
    CK(ctx->w_n_jn.ensure(n1 * 4)); CK(ctx->w_mflags.ensure(n1)); CK(ctx->w_cig_cap.ensure(n1 * 4)); CK(ctx->w_seg_cap.ensure(n1 * 4));
    CK(ctx->w_te_cap.ensure(n1 * 4)); CK(ctx->w_ws_off.ensure(n1 * 8)); CK(ctx->w_cig_buf.ensure(n1)); CK(ctx->w_n_cig_cur.ensure(n1 * 4));
    CK(ctx->w_seg_buf.ensure(n1)); CK(ctx->w_n_seg_cur.ensure(n1 * 4)); CK(ctx->w_n_te.ensure(n1 * 4)); CK(ctx->w_te_cnt.ensure(n1 * 12));
    CK(ctx->w_nm_extra.ensure(n1 * 4)); CK(ctx->w_refresh.ensure(n1)); CK(ctx->w_o_seq.ensure(n1 * 8)); CK(ctx->w_o_cig.ensure(n1 * 8));
    CK(ctx->w_o_te.ensure(n1 * 8)); CK(ctx->w_o_jn.ensure(n1 * 8)); CK(ctx->w_seq_len.ensure(n1 * 4)); CK(ctx->w_plan.ensure(n1 * sizeof(EmitPlan))); CK(ctx->cursors.ensure(64));
    CK(ctx->o_nm.ensure(n1 * 4)); CK(ctx->o_md_off.ensure(n1 * 8)); CK(ctx->o_md_len.ensure(n1 * 4));
    WorkDev& W = ctx->W;
    W.status = (uint32_t*)ctx->w_status.p; W.counters = (int32_t*)ctx->w_counters.p;
    W.md_init_off = (int64_t*)ctx->w_md_init_off.p; W.md_init_len = (int32_t*)ctx->w_md_init_len.p; W.nm_init = (int32_t*)ctx->w_nm_init.p;
    W.n_mdtok = (int32_t*)ctx->w_n_mdtok.p; W.n_jn = (int32_t*)ctx->w_n_jn.p; W.mflags = (uint8_t*)ctx->w_mflags.p;
    W.cig_cap = (int32_t*)ctx->w_cig_cap.p; W.seg_cap = (int32_t*)ctx->w_seg_cap.p; W.te_cap = (int32_t*)ctx->w_te_cap.p;
    W.ws_off = (int64_t*)ctx->w_ws_off.p; W.cig_buf = (int8_t*)ctx->w_cig_buf.p; W.n_cig_cur = (int32_t*)ctx->w_n_cig_cur.p;
    W.seg_buf = (int8_t*)ctx->w_seg_buf.p; W.n_seg_cur = (int32_t*)ctx->w_n_seg_cur.p; W.n_te = (int32_t*)ctx->w_n_te.p;
    W.te_cnt = (int32_t*)ctx->w_te_cnt.p; W.nm_extra = (int32_t*)ctx->w_nm_extra.p; W.refresh = (uint8_t*)ctx->w_refresh.p;
    W.o_seq = (int64_t*)ctx->w_o_seq.p; W.o_cig = (int64_t*)ctx->w_o_cig.p; W.o_te = (int64_t*)ctx->w_o_te.p; W.o_jn = (int64_t*)ctx->w_o_jn.p; W.seq_len_out = (int32_t*)ctx->w_seq_len.p; W.plan = (EmitPlan*)ctx->w_plan.p;
    return 0;
}

static Params make_params(tc_ctx* ctx) {
    Params P; P.G = ctx->G; P.S = ctx->S; P.V = ctx->V; P.B = ctx->B; P.W = ctx->W; P.O = ctx->opt;
    P.sj_enabled = ctx->S.n_ann > 0 ? 1 : 0;
    return P;
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
    int nte = W.n_te[i]; tc_te* dst = O.te + W.o_te[i];
    if (dry) { for (int k = 0; k < nte; k++) dst[k] = v.te[k]; return; }
    int m = 0; for (int ty = 0; ty < 4; ty++) for (int k = 0; k < nte; k++) if (v.te[k].type == ty) dst[m++] = v.te[k];
    CigSrc cs;
    if (W.cig_buf[i] < 0) { cs.packed = 0; cs.op = B.cig_op + B.cig_off[i]; cs.len = B.cig_len + B.cig_off[i]; cs.n = (int)(B.cig_off[i + 1] - B.cig_off[i]); }
    else { cs.packed = v.cig[W.cig_buf[i]]; cs.op = 0; cs.len = 0; cs.n = W.n_cig_cur[i]; }
    for (int k = 0; k < cs.n; k++) { int op; int32_t l; cig_get(cs, k, op, l); O.cig_op[W.o_cig[i] + k] = (uint8_t)op; O.cig_len[W.o_cig[i] + k] = l; }
    for (int k = 0; k < W.n_jn[i]; k++) { const int32_t* r = v.jn + k * JN_STRIDE; O.jm[W.o_jn[i] + k] = (int8_t)r[7]; O.ji[2 * (W.o_jn[i] + k)] = r[5]; O.ji[2 * (W.o_jn[i] + k) + 1] = r[6]; }
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

int tc_batch_run(tc_ctx* ctx, int dry) {
    if (!ctx) return TC_E_ARG;
    if (!ctx->have_genome) { ctx->err = "tc_batch_run: load a genome first"; return TC_E_STATE; }
    const int64_t n = ctx->n;
#ifndef TC_HOSTSIM
    CK(cudaSetDevice(ctx->device));
#endif
    if (alloc_work(ctx)) return TC_E_CUDA;
    ctx->tm.launches = 0; ctx->tm.md_pool_retries = 0;
    HostBuf& ht = ctx->h_tot; CK(ht.ensure(128)); int64_t* tot = (int64_t*)ht.p;
    unsigned long long* cur = (unsigned long long*)ctx->cursors.p;
    Params P = make_params(ctx);
    ctx->tot_seq = ctx->tot_cig = ctx->tot_te = ctx->tot_jn = 0; ctx->md_used = 0;
    if (n == 0) { ctx->ran = true; ctx->ran_dry = dry != 0; return TC_OK; }
#ifndef TC_HOSTSIM
    const int TPB = 128; const int gb = (int)((n + TPB - 1) / TPB);
    int nsm = 148; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device);
    int64_t wb = (n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK; int gw = (int)(wb < (int64_t)nsm * 8 ? wb : (int64_t)nsm * 8);
    CK(cudaEventRecord(ctx->ev[2], ctx->stream));
    // ---- init NM/MD (pool sized from the reads that lack a tag; exact retry on overflow)
    int64_t init_cap = ctx->init_pool.cap > (1 << 20) ? (int64_t)ctx->init_pool.cap : (1 << 20);
    for (int attempt = 0;; attempt++) {
        CK(ctx->init_pool.ensure((size_t)init_cap)); init_cap = (int64_t)ctx->init_pool.cap;
        if (be_zero(ctx, cur, 64)) return TC_E_CUDA;
        InitDev I; I.pool = (uint8_t*)ctx->init_pool.p; I.cursor = cur; I.cap = init_cap;
        P.W.md_init_pool = I.pool; ctx->W.md_init_pool = I.pool;
        k_init<<<gw, WARPS_PER_BLOCK * 32, 0, ctx->stream>>>(P, I, ctx->W.nm_init); ctx->tm.launches++;
        CK(cudaGetLastError());
        if (be_read_i64(ctx, (const int64_t*)cur, tot) || be_sync(ctx)) return TC_E_CUDA;
        if (tot[0] <= init_cap) break;
        if (attempt >= 2) { ctx->err = "init MD pool overflow"; return TC_E_CUDA; }
        init_cap = tot[0] + 1024; ctx->tm.md_pool_retries++;
    }
    CK(cudaEventRecord(ctx->ev[3], ctx->stream));
    // ---- measure + workspace
    k_measure<<<gb, TPB, 0, ctx->stream>>>(P); ctx->tm.launches++;
    CK(cudaGetLastError());
    if (be_zero(ctx, ctx->W.ws_off + n, 8) || be_scan(ctx, ctx->W.ws_off, n + 1) || be_read_i64(ctx, ctx->W.ws_off + n, tot) || be_sync(ctx)) return TC_E_CUDA;
    CK(ctx->w_ws.ensure((size_t)(tot[0] + 8) * 8));
    ctx->W.ws = (u64*)ctx->w_ws.p; P.W.ws = ctx->W.ws;
    CK(cudaEventRecord(ctx->ev[4], ctx->stream));
    if (dry) { k_dry<<<gb, TPB, 0, ctx->stream>>>(P); ctx->tm.launches++; CK(cudaGetLastError()); CK(cudaEventRecord(ctx->ev[5], ctx->stream)); CK(cudaEventRecord(ctx->ev[6], ctx->stream)); }
    else {
        k_plan<<<gb, TPB, 0, ctx->stream>>>(P); ctx->tm.launches++; CK(cudaGetLastError());
        CK(cudaEventRecord(ctx->ev[5], ctx->stream));
        k_junction<<<gb, TPB, 0, ctx->stream>>>(P); ctx->tm.launches++; CK(cudaGetLastError());
        CK(cudaEventRecord(ctx->ev[6], ctx->stream));
        k_sizes<<<gb, TPB, 0, ctx->stream>>>(P); ctx->tm.launches++; CK(cudaGetLastError());
    }
    int64_t* sc[4] = {ctx->W.o_seq, ctx->W.o_cig, ctx->W.o_te, ctx->W.o_jn};
    for (int k = 0; k < 4; k++) { if (be_zero(ctx, sc[k] + n, 8) || be_scan(ctx, sc[k], n + 1) || be_read_i64(ctx, sc[k] + n, tot + 1 + k)) return TC_E_CUDA; }
    if (be_sync(ctx)) return TC_E_CUDA;
    ctx->tot_seq = tot[1]; ctx->tot_cig = tot[2]; ctx->tot_te = tot[3]; ctx->tot_jn = tot[4];
    CK(ctx->o_seq.ensure((size_t)ctx->tot_seq + 64)); CK(ctx->o_cig_op.ensure((size_t)ctx->tot_cig + 64)); CK(ctx->o_cig_len.ensure((size_t)ctx->tot_cig * 4 + 64));
    CK(ctx->o_te.ensure((size_t)ctx->tot_te * 16 + 64)); CK(ctx->o_jm.ensure((size_t)ctx->tot_jn + 64)); CK(ctx->o_ji.ensure((size_t)ctx->tot_jn * 8 + 64));
    CK(cudaFuncSetAttribute(k_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EMIT_SMEM_BYTES));
    int occ = 1; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_emit, WARPS_PER_BLOCK * 32, EMIT_SMEM_BYTES));
    int ge = (int)(wb < (int64_t)nsm * occ ? wb : (int64_t)nsm * occ);
    const int64_t seq_words = (int64_t)(ctx->b_seq.cap / 4);
    int64_t md_cap = ctx->o_md_pool.cap > 0 ? (int64_t)ctx->o_md_pool.cap : (64 * n + 8 * ctx->n_cig + (1 << 20));
    CK(cudaEventRecord(ctx->ev[7], ctx->stream));
    for (int attempt = 0;; attempt++) {
        CK(ctx->o_md_pool.ensure((size_t)md_cap)); md_cap = (int64_t)ctx->o_md_pool.cap;
        if (be_zero(ctx, cur, 64)) return TC_E_CUDA;
        OutDev O; O.nm = (int32_t*)ctx->o_nm.p; O.md_off = (int64_t*)ctx->o_md_off.p; O.md_len = (int32_t*)ctx->o_md_len.p;
        O.md_pool = (uint8_t*)ctx->o_md_pool.p; O.md_cursor = cur; O.md_cap = md_cap; O.seq = (uint8_t*)ctx->o_seq.p;
        O.cig_op = (uint8_t*)ctx->o_cig_op.p; O.cig_len = (int32_t*)ctx->o_cig_len.p; O.te = (tc_te*)ctx->o_te.p; O.jm = (int8_t*)ctx->o_jm.p; O.ji = (int32_t*)ctx->o_ji.p;
        k_emit<<<ge, WARPS_PER_BLOCK * 32, EMIT_SMEM_BYTES, ctx->stream>>>(P, O, dry, seq_words); ctx->tm.launches++;
        CK(cudaGetLastError());
        if (attempt == 0) CK(cudaEventRecord(ctx->ev[8], ctx->stream));
        if (be_read_i64(ctx, (const int64_t*)cur, tot) || be_sync(ctx)) return TC_E_CUDA;
        if (tot[0] <= md_cap) { ctx->md_used = tot[0]; break; }
        if (attempt >= 2) { ctx->err = "MD pool overflow"; return TC_E_CUDA; }
        md_cap = tot[0] + 1024; ctx->tm.md_pool_retries++;
    }
    CK(cudaEventRecord(ctx->ev[9], ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->tm.kernels_ms, ctx->ev[2], ctx->ev[9]);
    cudaEventElapsedTime(&ctx->tm.k_nmmd_init_ms, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&ctx->tm.k_measure_ms, ctx->ev[3], ctx->ev[4]);
    cudaEventElapsedTime(&ctx->tm.k_plan_ms, ctx->ev[4], ctx->ev[5]);
    cudaEventElapsedTime(&ctx->tm.k_junction_ms, ctx->ev[5], ctx->ev[6]);
    cudaEventElapsedTime(&ctx->tm.k_emit_ms, ctx->ev[7], ctx->ev[8]);
#else
    std::vector<uint8_t> ipool, opool;
    {   // init needs the pool pointer before measure; build it first
        for (int64_t i = 0; i < n; i++) host_init_read(P, i, ipool);
        CK(ctx->init_pool.ensure(ipool.size() + 8)); memcpy(ctx->init_pool.p, ipool.data(), ipool.size());
        P.W.md_init_pool = (const uint8_t*)ctx->init_pool.p; ctx->W.md_init_pool = P.W.md_init_pool;
    }
    for (int64_t i = 0; i < n; i++) measure_read(P, i);
    ctx->W.ws_off[n] = 0; be_scan(ctx, ctx->W.ws_off, n + 1);
    CK(ctx->w_ws.ensure((size_t)(ctx->W.ws_off[n] + 8) * 8)); ctx->W.ws = (u64*)ctx->w_ws.p; P.W.ws = ctx->W.ws;
    for (int64_t i = 0; i < n; i++) {
        if (dry) { dry_read(P, i); dry_sizes(P, i); }
        else { plan_read(P, i); junction_read(P, i); sizes_read(P, i); }
    }
    int64_t* sc[4] = {ctx->W.o_seq, ctx->W.o_cig, ctx->W.o_te, ctx->W.o_jn};
    for (int k = 0; k < 4; k++) { sc[k][n] = 0; be_scan(ctx, sc[k], n + 1); tot[1 + k] = sc[k][n]; }
    ctx->tot_seq = tot[1]; ctx->tot_cig = tot[2]; ctx->tot_te = tot[3]; ctx->tot_jn = tot[4];
    CK(ctx->o_seq.ensure((size_t)ctx->tot_seq + 64)); CK(ctx->o_cig_op.ensure((size_t)ctx->tot_cig + 64)); CK(ctx->o_cig_len.ensure((size_t)ctx->tot_cig * 4 + 64));
    CK(ctx->o_te.ensure((size_t)ctx->tot_te * 16 + 64)); CK(ctx->o_jm.ensure((size_t)ctx->tot_jn + 64)); CK(ctx->o_ji.ensure((size_t)ctx->tot_jn * 8 + 64));
    OutDev O; O.nm = (int32_t*)ctx->o_nm.p; O.md_off = (int64_t*)ctx->o_md_off.p; O.md_len = (int32_t*)ctx->o_md_len.p; O.md_pool = 0; O.md_cursor = 0; O.md_cap = 0;
    O.seq = (uint8_t*)ctx->o_seq.p; O.cig_op = (uint8_t*)ctx->o_cig_op.p; O.cig_len = (int32_t*)ctx->o_cig_len.p; O.te = (tc_te*)ctx->o_te.p; O.jm = (int8_t*)ctx->o_jm.p; O.ji = (int32_t*)ctx->o_ji.p;
    for (int64_t i = 0; i < n; i++) host_emit_read(P, O, i, dry, opool);
    CK(ctx->o_md_pool.ensure(opool.size() + 8)); memcpy(ctx->o_md_pool.p, opool.data(), opool.size()); ctx->md_used = (int64_t)opool.size();
    (void)cur;
#endif
    ctx->ran = true; ctx->ran_dry = dry != 0;
    return TC_OK;
}

int tc_batch_download(tc_ctx* ctx, tc_batch_out* out) {
    if (!ctx || !out) return TC_E_ARG;
    if (!ctx->ran) { ctx->err = "tc_batch_download: nothing has run"; return TC_E_STATE; }
    const int64_t n = ctx->n; const size_t n1 = (size_t)n + 1;
#ifndef TC_HOSTSIM
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventRecord(ctx->ev[10], ctx->stream));
#endif
    ctx->tm.d2h_bytes = 0;
    memset(out, 0, sizeof(*out));
    out->n_reads = n;
    if (n > 0) {
        if (be_d2h(ctx, ctx->h_status, ctx->W.status, (size_t)n * 4) || be_d2h(ctx, ctx->h_counters, ctx->W.counters, (size_t)n * 40) ||
            be_d2h(ctx, ctx->h_te_off, ctx->W.o_te, n1 * 8) || be_d2h(ctx, ctx->h_te, ctx->o_te.p, (size_t)ctx->tot_te * 16)) return TC_E_CUDA;
        if (!ctx->ran_dry) {
            if (be_d2h(ctx, ctx->h_nm, ctx->o_nm.p, (size_t)n * 4) || be_d2h(ctx, ctx->h_seq_off, ctx->W.o_seq, n1 * 8) || be_d2h(ctx, ctx->h_seq_len, ctx->W.seq_len_out, (size_t)n * 4) ||
                be_d2h(ctx, ctx->h_seq, ctx->o_seq.p, (size_t)ctx->tot_seq) || be_d2h(ctx, ctx->h_cig_off, ctx->W.o_cig, n1 * 8) ||
                be_d2h(ctx, ctx->h_cig_op, ctx->o_cig_op.p, (size_t)ctx->tot_cig) || be_d2h(ctx, ctx->h_cig_len, ctx->o_cig_len.p, (size_t)ctx->tot_cig * 4) ||
                be_d2h(ctx, ctx->h_md_off, ctx->o_md_off.p, (size_t)n * 8) || be_d2h(ctx, ctx->h_md_len, ctx->o_md_len.p, (size_t)n * 4) ||
                be_d2h(ctx, ctx->h_md, ctx->o_md_pool.p, (size_t)ctx->md_used) || be_d2h(ctx, ctx->h_jn_off, ctx->W.o_jn, n1 * 8) ||
                be_d2h(ctx, ctx->h_jm, ctx->o_jm.p, (size_t)ctx->tot_jn) || be_d2h(ctx, ctx->h_ji, ctx->o_ji.p, (size_t)ctx->tot_jn * 8)) return TC_E_CUDA;
        }
#ifndef TC_HOSTSIM
        CK(cudaEventRecord(ctx->ev[11], ctx->stream));
#endif
        if (be_sync(ctx)) return TC_E_CUDA;
#ifndef TC_HOSTSIM
        cudaEventElapsedTime(&ctx->tm.d2h_ms, ctx->ev[10], ctx->ev[11]);
        cudaEventElapsedTime(&ctx->tm.h2d_ms, ctx->ev[0], ctx->ev[1]);
#endif
    }
    out->status = (const uint32_t*)ctx->h_status.p; out->counters = (const int32_t*)ctx->h_counters.p; out->nm = (const int32_t*)ctx->h_nm.p;
    out->seq_off = (const int64_t*)ctx->h_seq_off.p; out->seq_len = (const int32_t*)ctx->h_seq_len.p; out->seq = (const uint8_t*)ctx->h_seq.p;
    out->cig_off = (const int64_t*)ctx->h_cig_off.p; out->cig_op = (const uint8_t*)ctx->h_cig_op.p; out->cig_len = (const int32_t*)ctx->h_cig_len.p;
    out->md_off = (const int64_t*)ctx->h_md_off.p; out->md_len = (const int32_t*)ctx->h_md_len.p; out->md = (const uint8_t*)ctx->h_md.p;
    out->jn_off = (const int64_t*)ctx->h_jn_off.p; out->jm = (const int8_t*)ctx->h_jm.p; out->ji = (const int32_t*)ctx->h_ji.p;
    out->te_off = (const int64_t*)ctx->h_te_off.p; out->te = (const tc_te*)ctx->h_te.p;
    return TC_OK;
}

static int run_all(tc_ctx* ctx, const tc_batch_in* in, tc_batch_out* out, int dry) {
    int rc = tc_batch_upload(ctx, in);
    if (rc) return rc;
    rc = tc_batch_run(ctx, dry);
    if (rc) return rc;
    return tc_batch_download(ctx, out);
}
int tc_correct_batch(tc_ctx* ctx, const tc_batch_in* in, tc_batch_out* out) { return run_all(ctx, in, out, 0); }
int tc_dryrun_batch(tc_ctx* ctx, const tc_batch_in* in, tc_batch_out* out) { return run_all(ctx, in, out, 1); }

int tc_get_timing(tc_ctx* ctx, tc_timing* t) { if (!ctx || !t) return TC_E_ARG; *t = ctx->tm; return TC_OK; }

}  // extern "C"

// tc_b200.cu - C ABI and batch orchestration of the B200 correction engine (see include/tc_b200.h).
//
// Pipeline of one batch (all on the context's stream; kernels in the .cuh files of this directory):
//   k_init / k_init_nmmd   status word; NM/MD bootstrap of reads lacking a tag (tr.py:47-48)      tc_nmmd.cuh
//   k_measure              workspace sizing, MD tags split into tokens        -> exclusive scan    tc_plan.cuh
//   k_plan                 mismatch + indel stages                                                 tc_plan.cuh
//   k_jn_init / k_jn_fix   junction objects; noncanonical junction rescue (warp per listed read)   tc_junction.cuh
//   k_sizes, k_pack        exact output sizes -> 4 exclusive scans; one plan row per read          tc_plan.cuh
//   k_small                TE rows / CIGAR / jM,jI into the dense outputs                          tc_emit.cuh
//   k_emit                 new SEQ (genome gather) + the clean-read NM/MD check                    tc_emit.cuh
//   k_nmmd_exact           token-emitting NM/MD walk for the reads k_emit lists                    tc_nmmd.cuh
//   k_seq_unpack / k_seq_pack   4-bit SEQ <-> ASCII rows on the copy streams (this file)
// With -DTC_HOSTSIM the same ABI is built for the CPU with serial loops instead of kernels;
// that build lives under tests/hostsim and is TEST INFRASTRUCTURE for the host-side code
// (parser, renderer, CLI) on machines without a GPU.  The product never loads it.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
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
    int64_t* delta_off; int32_t* delta_len;                              // SEQ delta of every read (-1: SEQ is the input SEQ)
    uint8_t* delta_pool;                                                 // dense: a read's delta lies at its delta_off
    uint8_t* seq;            // device scratch: SEQ rows of the reads whose NM/MD need the exact walk (never downloaded)
    uint8_t* cig_op; int32_t* cig_len; u64* te; int8_t* jm; int32_t* ji;
    int32_t* exact_list; unsigned* exact_count;   // reads whose NM/MD need the token-emitting walk (k_nmmd_exact)
};
struct InitDev { uint8_t* pool; unsigned long long* cursor; int64_t cap; };

TC_HD int ndigits(uint32_t v) {
    return v < 10u ? 1 : v < 100u ? 2 : v < 1000u ? 3 : v < 10000u ? 4 : v < 100000u ? 5 : v < 1000000u ? 6 : v < 10000000u ? 7 :
           v < 100000000u ? 8 : v < 1000000000u ? 9 : 10;
}
// decimal digits of v < 10000 (nd of them, nd <= 4) packed little-endian into one word, without a loop
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

// SEQ delta of one read from its segment list (serial; the warp version is in k_small).  dst == 0: size only.
TC_HD int64_t delta_encode(const GenomeDev& G, const u64* segs, int ns, uint8_t* dst) {
    int64_t n = 0, cur = 0;
    for (int s = 0; s < ns; s++) {
        const int64_t src = seg_src(segs[s]), len = seg_len(segs[s]);
        if (len <= 0) continue;
        if (seg_kind(segs[s]) == 0) {
            const int64_t gap = src - cur;
            if (gap > 0) { if (dst) dl_put_hdr(dst + n, DL_SKIP, gap); n += dl_hdr_len(gap); }
            if (dst) dl_put_hdr(dst + n, DL_COPY, len);
            n += dl_hdr_len(len); cur = src + len;
        } else {
            if (dst) { dl_put_hdr(dst + n, DL_LIT, len); for (int64_t j = 0; j < len; j++) dst[n + dl_hdr_len(len) + j] = genome_byte(G, src + j); }
            n += dl_hdr_len(len) + len;
        }
    }
    return n;
}

// TE rows back into the reference's order (stage order: mismatches, insertions, deletions, junctions;
// positional inside a stage - the planner wrote them positionally) as compact words, the final CIGAR and the jM / jI
// snapshot of one read -> their places in the dense outputs.  Serial twin of k_small (CPU test build).
TC_HD void small_copies_read(const Params& P, const OutDev& O, int64_t i, int dry) {
    const WorkDev& W = P.W;
    if ((W.status[i] & TC_ST_CLASS_MASK) != 0) return;
    const WsView v = ws_view(W, i);
    const int nte = W.n_te[i];
    u64* dst = O.te + W.o_te[i];
    int slot[4]; slot[0] = 0; slot[1] = W.te_cnt[3 * i]; slot[2] = slot[1] + W.te_cnt[3 * i + 1]; slot[3] = slot[2] + W.te_cnt[3 * i + 2];
    const int base3 = slot[3]; int k3 = 0;
    for (int k = 0; k < nte; k++) {
        const tc_te r = v.te[k];
        const u64 w = te_word(r.a, r.size, r.type, r.status, r.reason);
        if (dry) { dst[k] = w; continue; }                            // (a dry run lists the rows positionally; no junction rows)
        if (r.type == 3) { dst[base3 + 2 * k3] = w; dst[base3 + 2 * k3 + 1] = (u64)(uint32_t)r.b; k3++; }
        else dst[slot[r.type]++] = w;
    }
    if (dry) return;
    const int cb = W.cig_buf[i]; const int64_t oc = W.o_cig[i];
    if (cb >= 0) {
        const u64* c = v.cig[cb]; const int n = W.n_cig_cur[i];
        for (int k = 0; k < n; k++) { O.cig_op[oc + k] = (uint8_t)cg_op(c[k]); O.cig_len[oc + k] = cg_len(c[k]); }
    }
    const int nj = W.n_jn[i]; const int64_t oj = W.o_jn[i];
    for (int k = 0; k < nj; k++) { const int32_t* r = v.jn + k * JN_STRIDE; O.jm[oj + k] = (int8_t)r[7]; O.ji[2 * (oj + k)] = r[5]; O.ji[2 * (oj + k) + 1] = r[6]; }
}

#ifndef TC_HOSTSIM
// ================================================================================== CUDA

#include "tc_warp.cuh"
#include "tc_nmmd.cuh"

__global__ void k_measure(Params P) { int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < P.B.n) measure_read(P, i); }
#ifndef TC_PLAN_BLOCKS
#define TC_PLAN_BLOCKS 5      // registers over occupancy: 8 blocks (62 registers, spills) was 7 % slower; 5 with the junction objects fused in: 1.54 -> 1.50 ms
#endif
// (Handing the reads of a 256-read tile out in order of their work - a counting sort in shared memory, so that the 32 reads of a warp
//  are neighbours in list length - was measured in round 2 and lost: k_plan 1.22 -> 1.37 ms, ONT 5.71 -> 5.97 ms per 1 M reads.  The
//  lanes that idle are not what bounds these kernels; the scattered per-read rows the permutation adds cost more.)
// TC_FUSE_JNINIT: the read's junction objects (junction_init_read: motif fetches and annotation lookups, all latency) are made by the
// thread that planned the read, so that their waits are filled by the other warps' planning instead of a kernel of their own
#ifndef TC_FUSE_JNINIT
#define TC_FUSE_JNINIT 1
#endif
__global__ void __launch_bounds__(128, TC_PLAN_BLOCKS) k_plan(Params P, int32_t* list, unsigned* count) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.B.n) return;
    plan_read(P, i);
#if TC_FUSE_JNINIT
    if (junction_init_read(P, i)) list[atomicAdd(count, 1u)] = (int32_t)i;
#else
    (void)list; (void)count;
#endif
}
#include "tc_junction.cuh"

__global__ void k_sizes(Params P) { int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < P.B.n) sizes_read(P, i); }
__global__ void k_dry(Params P) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P.B.n) { dry_read(P, i); dry_sizes(P, i); }
}

__global__ void k_pack(Params P) { int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < P.B.n) pack_read(P, i); }

#include "tc_emit.cuh"
#include "tc_fast.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { ctx->err = std::string(#x) + ": " + cudaGetErrorString(e_); return TC_E_CUDA; } } while (0)

// page-locked memory for the host parser's column buffers (tc_host.cpp)
extern void* (*tc_pinned_alloc)(size_t);
extern void (*tc_pinned_free)(void*);
static struct PinnedHooks {
    PinnedHooks() {
        tc_pinned_alloc = [](size_t n) -> void* { void* p = nullptr; if (cudaHostAlloc(&p, n, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; } return p; };
        tc_pinned_free = [](void* p) { cudaFreeHost(p); };
    }
} g_pinned_hooks;

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
        size_t want = n + n / 8 + 256; void* q = nullptr;
        cudaError_t e = cudaHostAlloc(&q, want, cudaHostAllocPortable | cudaHostAllocMapped);   // new buffer first: a failed grow keeps the old one
        if (e != cudaSuccess) return e;
        if (p) cudaFreeHost(p);
        p = q; cap = want;
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
    DevBuf b_seqp, b_seq_poff, b_exc_pos, b_exc_byte;   // 4-bit / 2-bit SEQ: packed input rows + their offsets; the 2-bit form's listed bases
    DevBuf b_cig16, b_cexc_idx, b_cexc_op, b_cexc_len; bool cig_unpack = false; int64_t cig_nexc = 0, cig_c0 = 0;   // 16-bit CIGAR form
    DevBuf b_plane, b_plane_ok, w_rest_list; int plane_mode = 0; bool plane_ready = false; int64_t plane_s0 = 0, plane_p0 = 0, plane_nexc = 0;   // 2-bit plane of the input SEQ (tc_fast.cuh); plane_mode 2: the caller's 2-bit rows, 1: made from the ASCII rows, 0: none
    int unpack_bits = 0; int64_t unpack_p0 = 0, unpack_s0 = 0, unpack_nexc = 0; size_t unpack_d0 = 0;   // expansion still to be launched (slot_run does it on its own stream)
    DevBuf b_flag, b_chrom, b_pos, b_tags, b_seq_off, b_seq, b_cig_off, b_cig_op, b_cig_len, b_md_off, b_md, b_ji_off, b_ji;
    BatchDev B; int64_t n = 0, n_seq = 0, n_cig = 0, n_md = 0, n_ji = 0, md_base = 0;
    int64_t seq_lo_word = 0, seq_hi_word = 0;       // words of B.seq (as uint32) that may be dereferenced
    DevBuf w_status, w_counters, w_md_init_off, w_md_init_len, w_nm_init, w_n_mdtok, w_md_tok, w_n_jn, w_mflags,
        w_cig_cap, w_seg_cap, w_te_cap, w_ws_off, w_ws, w_cig_buf, w_n_cig_cur, w_seg_buf, w_n_seg_cur, w_n_te,
        w_te_cnt, w_nm_extra, w_refresh, w_o_seq, w_o_cig, w_o_te, w_o_jn, w_o_delta, w_delta_size, w_seq_len, w_plan, w_fix_list, init_pool, cursors, scan_tmp;
    WorkDev W;
    DevBuf o_nm, o_md_off, o_md_len, o_md_pool, o_delta_off, o_delta_len, o_delta_pool, o_seq, o_cig_op, o_cig_len, o_te, o_jm, o_ji;
    int64_t tot_seq = 0, tot_cig = 0, tot_te = 0, tot_jn = 0, md_used = 0, delta_used = 0;
    int64_t base_cig = 0, base_te = 0, base_jn = 0, base_md = 0, base_delta = 0;   // position of this chunk in the batch output
    HostBuf h_tot;
    bool ran = false, ran_dry = false;
#ifndef TC_HOSTSIM
    cudaEvent_t ev_in = nullptr, ev_cmp = nullptr, ev_out = nullptr;
#endif
    void release() {
        DevBuf* d[] = {&b_flag, &b_chrom, &b_pos, &b_tags, &b_seq_off, &b_seq, &b_cig_off, &b_cig_op, &b_cig_len, &b_md_off, &b_md, &b_ji_off, &b_ji,
                       &w_status, &w_counters, &w_md_init_off, &w_md_init_len, &w_nm_init, &w_n_mdtok, &w_md_tok, &w_n_jn, &w_mflags, &w_cig_cap, &w_seg_cap,
                       &w_te_cap, &w_ws_off, &w_ws, &w_cig_buf, &w_n_cig_cur, &w_seg_buf, &w_n_seg_cur, &w_n_te, &w_te_cnt, &w_nm_extra, &w_refresh,
                       &w_o_seq, &w_o_cig, &w_o_te, &w_o_jn, &w_o_delta, &w_delta_size, &w_seq_len, &w_plan, &w_fix_list, &init_pool, &cursors, &scan_tmp,
                       &o_nm, &o_md_off, &o_md_len, &o_md_pool, &o_delta_off, &o_delta_len, &o_delta_pool, &o_seq, &o_cig_op, &o_cig_len, &o_te, &o_jm, &o_ji};
        for (DevBuf* b : d) b->release();
        b_seqp.release(); b_seq_poff.release(); b_exc_pos.release(); b_exc_byte.release();
        b_cig16.release(); b_cexc_idx.release(); b_cexc_op.release(); b_cexc_len.release();
        b_plane.release(); b_plane_ok.release(); w_rest_list.release();
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
    bool packed = false;            // the current batch came with 4-bit SEQ
    bool no_plane = false;          // TC_NO_PLANE=1: k_emit decides every read (the ASCII path alone; for measurements and tests)
    bool skip_seq = false;          // dry run over reads that all carry NM and MD: the SEQ column is never read, so it is not uploaded
    // pinned host outputs (whole batch)
    HostBuf h_status, h_counters, h_nm, h_delta_off, h_delta_len, h_delta, h_cig_off, h_cig_op, h_cig_len, h_md_off, h_md_len, h_md,
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
// 8 characters (lo, hi) to d, whose address mod 4 is `mis` (the same for every 8-byte group of a row): the widest naturally aligned stores
__device__ __forceinline__ void store8(uint8_t* d, uint32_t lo, uint32_t hi, int left, int mis) {
    if (left < 8) {
#pragma unroll
        for (int k = 0; k < 7; k++) if (k < left) d[k] = (uint8_t)((k < 4 ? lo : hi) >> (8 * (k & 3)));
    } else if (mis == 0) { reinterpret_cast<uint32_t*>(d)[0] = lo; reinterpret_cast<uint32_t*>(d)[1] = hi; }
    else if (mis == 2) { *reinterpret_cast<uint16_t*>(d) = (uint16_t)lo; *reinterpret_cast<uint32_t*>(d + 2) = (lo >> 16) | (hi << 16); *reinterpret_cast<uint16_t*>(d + 6) = (uint16_t)(hi >> 16); }
    else if (mis == 1) { d[0] = (uint8_t)lo; *reinterpret_cast<uint16_t*>(d + 1) = (uint16_t)(lo >> 8); *reinterpret_cast<uint32_t*>(d + 3) = (lo >> 24) | (hi << 8); d[7] = (uint8_t)(hi >> 24); }
    else { d[0] = (uint8_t)lo; *reinterpret_cast<uint32_t*>(d + 1) = (lo >> 8) | (hi << 24); *reinterpret_cast<uint16_t*>(d + 5) = (uint16_t)(hi >> 8); d[7] = (uint8_t)(hi >> 24); }
}
// 2-bit SEQ rows -> ASCII rows, one warp per read: a lane turns two bytes (8 bases) into 8 characters
__global__ void __launch_bounds__(256) k_seq_unpack2(int64_t n, const int64_t* seq_off, const int64_t* poff, const uint8_t* packed, uint8_t* seq) {
    __shared__ uint32_t lut[256];
    for (int t = threadIdx.x; t < 256; t += blockDim.x) {
        const uint32_t acgt = 0x54474341u;
        lut[t] = ((acgt >> (8 * (t & 3))) & 255u) | (((acgt >> (8 * ((t >> 2) & 3))) & 255u) << 8) |
                 (((acgt >> (8 * ((t >> 4) & 3))) & 255u) << 16) | (((acgt >> (8 * ((t >> 6) & 3))) & 255u) << 24);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31; const int64_t nw = (int64_t)gridDim.x * 8;
    for (int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += nw) {
        const int64_t o = seq_off[i]; const int L = (int)(seq_off[i + 1] - o);
        const uint16_t* src = reinterpret_cast<const uint16_t*>(packed + poff[i]);     // rows start at multiples of 4 bytes
        uint8_t* dst = seq + o;
        const int mis = (int)(reinterpret_cast<uintptr_t>(dst) & 3);
        for (int w = lane; w * 8 < L; w += 32) { const uint32_t x = src[w]; store8(dst + w * 8, lut[x & 255u], lut[x >> 8], L - w * 8, mis); }
    }
}
// 16-bit CIGAR words -> the op / length arrays the kernels read (flat over the ops), then the listed ops
__global__ void k_cig_unpack(int64_t n_ops, const uint16_t* w, uint8_t* op, int32_t* len) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_ops; k += (int64_t)gridDim.x * blockDim.x) {
        const unsigned x = w[k];
        op[k] = (uint8_t)"MIDNSHP=XB??????"[x & 15u]; len[k] = (int32_t)(x >> 4);
    }
}
__global__ void k_cig_patch(int64_t n_exc, const int64_t* idx, const uint8_t* eop, const int32_t* elen, uint8_t* op, int32_t* len) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_exc; k += (int64_t)gridDim.x * blockDim.x) { op[idx[k]] = eop[k]; len[idx[k]] = elen[k]; }
}
__global__ void k_seq_patch(int64_t n_exc, const int64_t* pos, const uint8_t* byte, uint8_t* seq) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_exc; k += (int64_t)gridDim.x * blockDim.x) seq[pos[k]] = byte[k];
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
    { const char* np = getenv("TC_NO_PLANE"); ctx->no_plane = np && np[0] == '1'; }
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
    HostBuf* h[] = {&ctx->h_status, &ctx->h_counters, &ctx->h_nm, &ctx->h_delta_off, &ctx->h_delta_len, &ctx->h_delta, &ctx->h_cig_off, &ctx->h_cig_op, &ctx->h_cig_len,
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

int tc_device_count(void) {
#ifndef TC_HOSTSIM
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
#else
    return 1;
#endif
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

// packed (4-bit or 2-bit) SEQ rows of reads [r0, r1): upload, then expand to the ASCII rows the kernels read (B.seq at d0)
static int seq_upload_packed(tc_ctx* ctx, Slot& S, const tc_batch_in* in, int64_t r0, int64_t r1, size_t d0, stream_t st) {
    const int64_t n = r1 - r0;
    CK(S.b_seq.ensure(d0 + (size_t)S.n_seq + 256 + 8));
    if (n == 0) return 0;
    const bool two = in->seq_bits == 2;
    const int64_t p0 = in->seq_poff[r0], pbytes = in->seq_poff[r1] - p0;
    if (be_h2d(ctx, st, S.b_seqp, 0, in->seq + p0, (size_t)pbytes, 64) || be_h2d(ctx, st, S.b_seq_poff, 0, in->seq_poff + r0, (size_t)(n + 1) * 8)) return TC_E_CUDA;
    const int64_t s0 = in->seq_off[r0], s1 = in->seq_off[r1];
    int64_t e0 = 0, e1 = 0;                                            // listed bases of these reads (exc_pos is ascending)
    if (two && in->n_exc > 0) {
        e0 = std::lower_bound(in->exc_pos, in->exc_pos + in->n_exc, s0) - in->exc_pos; e1 = std::lower_bound(in->exc_pos, in->exc_pos + in->n_exc, s1) - in->exc_pos;
        if (e1 > e0 && (be_h2d(ctx, st, S.b_exc_pos, 0, in->exc_pos + e0, (size_t)(e1 - e0) * 8) || be_h2d(ctx, st, S.b_exc_byte, 0, in->exc_byte + e0, (size_t)(e1 - e0)))) return TC_E_CUDA;
    }
#ifndef TC_HOSTSIM
    // the expansion itself is launched by slot_run on the compute stream: the upload stream carries copies only, so the next
    // chunk's bytes follow at once (the kernel stream has slack, the host link has none)
    S.unpack_bits = two ? 2 : 4; S.unpack_p0 = p0; S.unpack_s0 = s0; S.unpack_nexc = e1 - e0; S.unpack_d0 = d0;
#else
    uint8_t* dst = (uint8_t*)S.b_seq.p + d0 - s0; const uint8_t* src = (const uint8_t*)S.b_seqp.p - p0;
    const int64_t* so = (const int64_t*)S.b_seq_off.p; const int64_t* po = (const int64_t*)S.b_seq_poff.p;
    for (int64_t i = 0; i < n; i++)
        for (int64_t k = 0; k < so[i + 1] - so[i]; k++) {
            if (two) dst[so[i] + k] = (uint8_t)"ACGT"[(src[po[i] + (k >> 2)] >> (2 * (k & 3))) & 3];
            else { const uint8_t b = src[po[i] + (k >> 1)]; dst[so[i] + k] = ctx->alphabet[(k & 1) ? b >> 4 : b & 15]; }
        }
    for (int64_t k = e0; k < e1; k++) dst[in->exc_pos[k]] = in->exc_byte[k];
#endif
    return 0;
}

// CIGAR ops [c0, c0 + S.n_cig): as op / length arrays, or as 16-bit words + listed ops that slot_run expands (cig_unpack)
static int cig_upload(tc_ctx* ctx, Slot& S, const tc_batch_in* in, int64_t c0, stream_t st) {
    S.cig_unpack = false;
    if (!in->cig16) {
        return be_h2d(ctx, st, S.b_cig_op, 0, S.n_cig ? in->cig_op + c0 : nullptr, (size_t)S.n_cig) ||
               be_h2d(ctx, st, S.b_cig_len, 0, S.n_cig ? in->cig_len + c0 : nullptr, (size_t)S.n_cig * 4);
    }
    CK(S.b_cig_op.ensure((size_t)S.n_cig + 72)); CK(S.b_cig_len.ensure((size_t)S.n_cig * 4 + 72));
    if (S.n_cig == 0) return 0;
    if (be_h2d(ctx, st, S.b_cig16, 0, in->cig16 + c0, (size_t)S.n_cig * 2)) return TC_E_CUDA;
    int64_t e0 = 0, e1 = 0;
    if (in->n_cig_exc > 0) {
        e0 = std::lower_bound(in->cig_exc_idx, in->cig_exc_idx + in->n_cig_exc, c0) - in->cig_exc_idx;
        e1 = std::lower_bound(in->cig_exc_idx, in->cig_exc_idx + in->n_cig_exc, c0 + S.n_cig) - in->cig_exc_idx;
        if (e1 > e0 && (be_h2d(ctx, st, S.b_cexc_idx, 0, in->cig_exc_idx + e0, (size_t)(e1 - e0) * 8) || be_h2d(ctx, st, S.b_cexc_op, 0, in->cig_exc_op + e0, (size_t)(e1 - e0)) ||
                        be_h2d(ctx, st, S.b_cexc_len, 0, in->cig_exc_len + e0, (size_t)(e1 - e0) * 4))) return TC_E_CUDA;
    }
#ifndef TC_HOSTSIM
    S.cig_unpack = true; S.cig_nexc = e1 - e0; S.cig_c0 = c0;
#else
    uint8_t* op = (uint8_t*)S.b_cig_op.p; int32_t* len = (int32_t*)S.b_cig_len.p; const uint16_t* w = (const uint16_t*)S.b_cig16.p;
    for (int64_t k = 0; k < S.n_cig; k++) { op[k] = (uint8_t)"MIDNSHP=XB??????"[w[k] & 15u]; len[k] = (int32_t)(w[k] >> 4); }
    for (int64_t k = e0; k < e1; k++) { op[in->cig_exc_idx[k] - c0] = in->cig_exc_op[k]; len[in->cig_exc_idx[k] - c0] = in->cig_exc_len[k]; }
#endif
    return 0;
}

// uploads reads [r0, r1) of `in` into slot S.  CSR offsets stay the batch's absolute offsets; the
// data pointers are shifted so that they can be applied unchanged.
static int slot_upload(tc_ctx* ctx, Slot& S, const tc_batch_in* in, int64_t r0, int64_t r1, stream_t st) {
    const int64_t n = r1 - r0;
    S.n = n; S.ran = false; S.unpack_bits = 0;
    static const int64_t zero_off[1] = {0};
    const int64_t s0 = n ? in->seq_off[r0] : 0, c0 = n ? in->cig_off[r0] : 0, m0 = n ? in->md_off[r0] : 0, j0 = n ? in->ji_off[r0] : 0;
    S.n_seq = n ? in->seq_off[r1] - s0 : 0; S.n_cig = n ? in->cig_off[r1] - c0 : 0;
    S.n_md = n ? in->md_off[r1] - m0 : 0; S.n_ji = n ? in->ji_off[r1] - j0 : 0; S.md_base = m0;
    const size_t d0 = (size_t)(s0 & 15);            // keeps (device pointer - s0) 16-byte aligned
    if (be_h2d(ctx, st, S.b_flag, 0, in->flag + r0, (size_t)n * 4) || be_h2d(ctx, st, S.b_chrom, 0, in->chrom + r0, (size_t)n * 4) ||
        be_h2d(ctx, st, S.b_pos, 0, in->pos + r0, (size_t)n * 4) || be_h2d(ctx, st, S.b_tags, 0, in->tags + r0, (size_t)n) ||
        be_h2d(ctx, st, S.b_seq_off, 0, n ? in->seq_off + r0 : zero_off, (size_t)(n + 1) * 8) ||
        (ctx->skip_seq ? be_h2d(ctx, st, S.b_seq, d0, nullptr, 0, (size_t)S.n_seq + 256)
                       : in->seq_poff ? seq_upload_packed(ctx, S, in, r0, r1, d0, st) : be_h2d(ctx, st, S.b_seq, d0, n ? in->seq + s0 : nullptr, (size_t)S.n_seq, 256)) ||
        be_h2d(ctx, st, S.b_cig_off, 0, n ? in->cig_off + r0 : zero_off, (size_t)(n + 1) * 8) ||
        cig_upload(ctx, S, in, c0, st) ||
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
    // 2-bit plane of the SEQ column for k_verify: the caller's own 2-bit rows when they came that way, else made from the ASCII rows (slot_run)
    S.plane_ready = false; S.plane_s0 = s0;
    S.plane_mode = (ctx->skip_seq || n == 0) ? 0 : (in->seq_poff && in->seq_bits == 2) ? 2 : 1;
    if (S.plane_mode == 2) { S.plane_p0 = in->seq_poff[r0]; S.plane_nexc = S.unpack_nexc; }
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
    CK(S.w_o_te.ensure(n1 * 8)); CK(S.w_o_jn.ensure(n1 * 8)); CK(S.w_o_delta.ensure(n1 * 8)); CK(S.w_delta_size.ensure(n1 * 4)); CK(S.w_seq_len.ensure(n1 * 4)); CK(S.w_plan.ensure(n1 * sizeof(EmitPlan))); CK(S.cursors.ensure(64));
    CK(S.o_nm.ensure(n1 * 4)); CK(S.o_md_off.ensure(n1 * 8)); CK(S.o_md_len.ensure(n1 * 4)); CK(S.o_delta_off.ensure(n1 * 8)); CK(S.o_delta_len.ensure(n1 * 4));
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
    W.o_delta = (int64_t*)S.w_o_delta.p; W.delta_size = (int32_t*)S.w_delta_size.p;
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
    O.delta_off = (int64_t*)S.o_delta_off.p; O.delta_len = (int32_t*)S.o_delta_len.p; O.delta_pool = (uint8_t*)S.o_delta_pool.p - S.base_delta;
    O.seq = (uint8_t*)S.o_seq.p; O.cig_op = (uint8_t*)S.o_cig_op.p - S.base_cig; O.cig_len = (int32_t*)S.o_cig_len.p - S.base_cig;
    O.te = (u64*)S.o_te.p - S.base_te; O.jm = (int8_t*)S.o_jm.p - S.base_jn; O.ji = (int32_t*)S.o_ji.p - 2 * S.base_jn;
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
// scalar twin of k_small + k_emit (test infrastructure)
static void host_emit_read(Params& P, OutDev& O, int64_t i, int dry, std::vector<uint8_t>& pool) {
    const BatchDev& B = P.B; const WorkDev& W = P.W;
    uint32_t st = W.status[i]; O.md_off[i] = 0; O.md_len[i] = 0; O.nm[i] = 0; O.delta_off[i] = 0; O.delta_len[i] = -1;
    if ((st & TC_ST_CLASS_MASK) != 0) return;
    WsView v = ws_view(W, i);
    small_copies_read(P, O, i, dry);
    if (dry) return;
    CigSrc cs;
    if (W.cig_buf[i] < 0) { cs.packed = 0; cs.op = B.cig_op + B.cig_off[i]; cs.len = B.cig_len + B.cig_off[i]; cs.n = (int)(B.cig_off[i + 1] - B.cig_off[i]); }
    else { cs.packed = v.cig[W.cig_buf[i]]; cs.op = 0; cs.len = 0; cs.n = W.n_cig_cur[i]; }
    const uint8_t* in_seq = B.seq + B.seq_off[i];
    std::vector<uint8_t> row((size_t)W.seq_len_out[i] + 16);
    uint8_t* out_seq = row.data();
    if (W.seg_buf[i] < 0) memcpy(out_seq, in_seq, (size_t)(B.seq_off[i + 1] - B.seq_off[i]));
    else { int64_t o = 0; const u64* segs = v.seg[W.seg_buf[i]];
        for (int s = 0; s < W.n_seg_cur[i]; s++) { int64_t src = seg_src(segs[s]), len = seg_len(segs[s]);
            for (int64_t j = 0; j < len; j++) out_seq[o + j] = seg_kind(segs[s]) ? genome_byte(P.G, src + j) : in_seq[src + j];
            o += len; }
        const int64_t dn = delta_encode(P.G, segs, W.n_seg_cur[i], nullptr);
        O.delta_off[i] = W.o_delta[i]; O.delta_len[i] = dn == W.delta_size[i] ? (int32_t)dn : -2;   // (-2: the planner's size is wrong; the renderers refuse it)
        if (dn == W.delta_size[i]) delta_encode(P.G, segs, W.n_seg_cur[i], O.delta_pool + W.o_delta[i]); }
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
    S.tot_seq = S.tot_cig = S.tot_te = S.tot_jn = 0; S.md_used = 0; S.delta_used = 0;
    if (n == 0) { S.ran = true; S.ran_dry = dry != 0; return TC_OK; }
#ifndef TC_HOSTSIM
#ifndef TC_TPB
#define TC_TPB 128
#endif
    const int TPB = TC_TPB; const int gb = (int)((n + TPB - 1) / TPB);
    if (S.cig_unpack) {                                                // 16-bit CIGAR words -> op / length arrays
        k_cig_unpack<<<(int)std::min<int64_t>((S.n_cig + 255) / 256, 148 * 16), 256, 0, st>>>(S.n_cig, (const uint16_t*)S.b_cig16.p, (uint8_t*)S.b_cig_op.p, (int32_t*)S.b_cig_len.p); ctx->tm.launches++;
        if (S.cig_nexc > 0) {
            k_cig_patch<<<(int)std::min<int64_t>((S.cig_nexc + 255) / 256, 1184), 256, 0, st>>>(S.cig_nexc, (const int64_t*)S.b_cexc_idx.p, (const uint8_t*)S.b_cexc_op.p, (const int32_t*)S.b_cexc_len.p,
                                                                                              (uint8_t*)S.b_cig_op.p - S.cig_c0, (int32_t*)S.b_cig_len.p - S.cig_c0); ctx->tm.launches++;
        }
        CK(cudaGetLastError());
        S.cig_unpack = false;
    }
    if (S.unpack_bits) {                                               // packed SEQ rows -> the ASCII rows the kernels read
        uint8_t* dst = (uint8_t*)S.b_seq.p + S.unpack_d0 - S.unpack_s0; const uint8_t* src = (const uint8_t*)S.b_seqp.p - S.unpack_p0;
        int64_t wb0 = (n + 7) / 8; int g = (int)(wb0 < 148 * 8 ? wb0 : 148 * 8);
        if (S.unpack_bits == 2) {
            k_seq_unpack2<<<g, 256, 0, st>>>(n, (const int64_t*)S.b_seq_off.p, (const int64_t*)S.b_seq_poff.p, src, dst); ctx->tm.launches++;
            if (S.unpack_nexc > 0) { k_seq_patch<<<(int)std::min<int64_t>((S.unpack_nexc + 255) / 256, 1184), 256, 0, st>>>(S.unpack_nexc, (const int64_t*)S.b_exc_pos.p, (const uint8_t*)S.b_exc_byte.p, dst); ctx->tm.launches++; }
        } else {
            Alphabet A; memcpy(A.c, ctx->alphabet, 16);
            k_seq_unpack<<<g, 256, 0, st>>>(n, (const int64_t*)S.b_seq_off.p, (const int64_t*)S.b_seq_poff.p, src, dst, A); ctx->tm.launches++;
        }
        CK(cudaGetLastError());
        S.unpack_bits = 0;
    }
    if (!dry && S.plane_mode && !S.plane_ready && !ctx->no_plane) {
        CK(S.b_plane_ok.ensure((size_t)n + 8)); CK(S.w_rest_list.ensure((size_t)(n + 1) * 4));
        if (S.plane_mode == 2) {
            CK(cudaMemsetAsync(S.b_plane_ok.p, 1, (size_t)n, st));
            if (S.plane_nexc > 0) { k_seq_plane_flags<<<(int)std::min<int64_t>((S.plane_nexc + 255) / 256, 1184), 256, 0, st>>>(S.plane_nexc, (const int64_t*)S.b_exc_pos.p, (const int64_t*)S.b_seq_off.p, n, (uint8_t*)S.b_plane_ok.p); ctx->tm.launches++; }
        } else {
            CK(S.b_plane.ensure(((size_t)S.n_seq / 16 + (size_t)n + 8) * 4));
            int64_t wb0 = (n + 7) / 8; int g = (int)(wb0 < 148 * 8 ? wb0 : 148 * 8);
            k_seq_plane<<<g, 256, 0, st>>>(n, (const int64_t*)S.b_seq_off.p, S.B.seq, S.plane_s0, (uint32_t*)S.b_plane.p, (uint8_t*)S.b_plane_ok.p); ctx->tm.launches++;
        }
        CK(cudaGetLastError());
        S.plane_ready = true;
    }
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
#ifndef TC_HOSTSIM
    if (S.w_ws.ensure((size_t)(tot[1] + 8) * 8) != cudaSuccess) {
        cudaGetLastError();
        ctx->err = "out of device memory for the per-read workspace (" + std::to_string((tot[1] + 8) * 8 >> 20) + " MB for " + std::to_string(n) +
                   " reads): run smaller batches, or tc_correct_batch / tc_dryrun_batch, which pipeline a batch in chunks (tc_ctx_set_pipeline)";
        return TC_E_CUDA;
    }
#else
    CK(S.w_ws.ensure((size_t)(tot[1] + 8) * 8));
#endif
    S.W.ws = (u64*)S.w_ws.p; P.W.ws = S.W.ws;
    if (dry) { k_dry<<<gb, TPB, 0, st>>>(P); ctx->tm.launches++; CK(cudaGetLastError()); if (timed) { CK(cudaEventRecord(ctx->ev[5], st)); CK(cudaEventRecord(ctx->ev[6], st)); } }
    else {
        CK(S.w_fix_list.ensure((size_t)(n + 1) * 4));
        if (be_zero(ctx, st, cur + 2, 8)) return TC_E_CUDA;
        k_plan<<<gb, TPB, 0, st>>>(P, (int32_t*)S.w_fix_list.p, (unsigned*)(cur + 2)); ctx->tm.launches++; CK(cudaGetLastError());
        if (timed) CK(cudaEventRecord(ctx->ev[5], st));
#if !TC_FUSE_JNINIT
        k_jn_init<<<gb, TPB, 0, st>>>(P, (int32_t*)S.w_fix_list.p, (unsigned*)(cur + 2)); ctx->tm.launches++; CK(cudaGetLastError());
#endif
        int occ_j = 1; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_j, k_jn_fix, WARPS_PER_BLOCK * 32, 0));
        const int gj = (int)(wb < (int64_t)nsm * occ_j ? wb : (int64_t)nsm * occ_j);            // persistent grid = the resident blocks
        k_jn_fix<<<gj, WARPS_PER_BLOCK * 32, 0, st>>>(P, (const int32_t*)S.w_fix_list.p, (const unsigned*)(cur + 2)); ctx->tm.launches++; CK(cudaGetLastError());
        if (timed) CK(cudaEventRecord(ctx->ev[6], st));
        k_sizes<<<gb, TPB, 0, st>>>(P); ctx->tm.launches++; CK(cudaGetLastError());
    }
    int64_t* sc[5] = {S.W.o_seq, S.W.o_cig, S.W.o_te, S.W.o_jn, S.W.o_delta};
    const int64_t bases[5] = {0, S.base_cig, S.base_te, S.base_jn, S.base_delta};   // (SEQ rows are device scratch: slot-local offsets)
    for (int k = 0; k < 5; k++) { if (be_zero(ctx, st, sc[k] + n, 8) || be_scan(ctx, st, S.scan_tmp, sc[k], n + 1, bases[k]) || be_read_i64(ctx, st, sc[k] + n, tot + 1 + k)) return TC_E_CUDA; }
    k_pack<<<gb, TPB, 0, st>>>(P); ctx->tm.launches++; CK(cudaGetLastError());
    if (be_sync(ctx, st)) return TC_E_CUDA;
    S.tot_seq = tot[1] - bases[0]; S.tot_cig = tot[2] - bases[1]; S.tot_te = tot[3] - bases[2]; S.tot_jn = tot[4] - bases[3]; S.delta_used = tot[5] - bases[4];
    CK(S.o_seq.ensure((size_t)S.tot_seq + 64)); CK(S.o_cig_op.ensure((size_t)S.tot_cig + 64)); CK(S.o_cig_len.ensure((size_t)S.tot_cig * 4 + 64));
    CK(S.o_te.ensure((size_t)S.tot_te * 8 + 64)); CK(S.o_jm.ensure((size_t)S.tot_jn + 64)); CK(S.o_ji.ensure((size_t)S.tot_jn * 8 + 64));
    CK(S.o_delta_pool.ensure((size_t)S.delta_used + 64));
    CK(cudaFuncSetAttribute(k_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EMIT_SMEM_BYTES));
    int occ = 1; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_emit, WARPS_PER_BLOCK * 32, EMIT_SMEM_BYTES));
    int ge = (int)(wb < (int64_t)nsm * occ ? wb : (int64_t)nsm * occ);
    int64_t md_cap = S.o_md_pool.cap > (size_t)(8 * n + (1 << 20)) ? (int64_t)S.o_md_pool.cap : (72 * n + 8 * S.n_cig + (1 << 20));
    {   // persistent grid: exactly the resident blocks, so that no block waits for a second wave
        int occ_s = 1; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_s, k_small, WARPS_PER_BLOCK * 32, 0));
        const int gs = (int)(wb < (int64_t)nsm * occ_s ? wb : (int64_t)nsm * occ_s);
        k_small<<<gs, WARPS_PER_BLOCK * 32, 0, st>>>(P, make_out(S, cur, md_cap), dry); ctx->tm.launches++; CK(cudaGetLastError());   // TE rows, CIGAR, jM / jI, SEQ delta
    }
    if (timed) CK(cudaEventRecord(ctx->ev[7], st));
    for (int attempt = 0;; attempt++) {
        CK(S.o_md_pool.ensure((size_t)md_cap)); md_cap = (int64_t)S.o_md_pool.cap;
        k_set_u64<<<1, 1, 0, st>>>(cur + 1, 0ull);
        k_set_u64<<<1, 1, 0, st>>>(cur, (unsigned long long)(S.base_md + 8 * n));      // the pool starts behind the inline slots
        k_set_u64<<<1, 1, 0, st>>>(cur + 3, 0ull);
        OutDev O = make_out(S, cur, md_cap);
        if (!dry && S.plane_ready && !ctx->no_plane) {                // NM / MD on the 2-bit plane; what that cannot settle goes to k_emit as a list
            SeqPlane SP; SP.ok = (const uint8_t*)S.b_plane_ok.p; SP.s0 = S.plane_s0; SP.by_poff = S.plane_mode == 2; SP.poff = (const int64_t*)S.b_seq_poff.p;
            SP.words = S.plane_mode == 2 ? (const uint32_t*)S.b_seqp.p - (S.plane_p0 >> 2) : (const uint32_t*)S.b_plane.p;
            int occ_v = 1; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_v, k_verify, WARPS_PER_BLOCK * 32, 0));
            const int gv = (int)(wb < (int64_t)nsm * occ_v ? wb : (int64_t)nsm * occ_v);
            k_verify<<<gv, WARPS_PER_BLOCK * 32, 0, st>>>(P, O, SP, (int32_t*)S.w_rest_list.p, (unsigned*)(cur + 3)); ctx->tm.launches++;
            k_emit<<<ge, WARPS_PER_BLOCK * 32, EMIT_SMEM_BYTES, st>>>(P, O, dry, S.seq_lo_word, S.seq_hi_word, (const int32_t*)S.w_rest_list.p, (const unsigned*)(cur + 3)); ctx->tm.launches++;
        } else if (!dry) { k_emit<<<ge, WARPS_PER_BLOCK * 32, EMIT_SMEM_BYTES, st>>>(P, O, dry, S.seq_lo_word, S.seq_hi_word, nullptr, nullptr); ctx->tm.launches++; }   // (a dry run ends with k_small)
        ctx->tm.launches += 3;
        CK(cudaGetLastError());
        if (timed && attempt == 0) CK(cudaEventRecord(ctx->ev[8], st));
        if (!dry) { k_nmmd_exact<<<gw, WARPS_PER_BLOCK * 32, 0, st>>>(P, O); ctx->tm.launches++; CK(cudaGetLastError()); }
        if (be_read_i64(ctx, st, (const int64_t*)cur, tot) || be_read_i64(ctx, st, (const int64_t*)cur + 1, tot + 6) || be_read_i64(ctx, st, (const int64_t*)cur + 3, tot + 7) || be_sync(ctx, st)) return TC_E_CUDA;
        if (tot[0] - S.base_md <= md_cap) { S.md_used = tot[0] - S.base_md; ctx->tm.exact_nmmd_reads += tot[6]; ctx->tm.ascii_path_reads += (S.plane_ready && !ctx->no_plane) ? tot[7] : n; break; }
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
    int64_t* sc[5] = {S.W.o_seq, S.W.o_cig, S.W.o_te, S.W.o_jn, S.W.o_delta};
    const int64_t bases[5] = {0, S.base_cig, S.base_te, S.base_jn, S.base_delta};
    for (int k = 0; k < 5; k++) { sc[k][n] = 0; be_scan(ctx, st, S.scan_tmp, sc[k], n + 1, bases[k]); tot[1 + k] = sc[k][n] - bases[k]; }
    S.tot_seq = tot[1]; S.tot_cig = tot[2]; S.tot_te = tot[3]; S.tot_jn = tot[4]; S.delta_used = tot[5];
    CK(S.o_seq.ensure((size_t)S.tot_seq + 64)); CK(S.o_cig_op.ensure((size_t)S.tot_cig + 64)); CK(S.o_cig_len.ensure((size_t)S.tot_cig * 4 + 64));
    CK(S.o_te.ensure((size_t)S.tot_te * 8 + 64)); CK(S.o_jm.ensure((size_t)S.tot_jn + 64)); CK(S.o_ji.ensure((size_t)S.tot_jn * 8 + 64));
    CK(S.o_delta_pool.ensure((size_t)S.delta_used + 64));
    for (int64_t i = 0; i < n; i++) pack_read(P, i);
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
        host_reserve(ctx, ctx->h_te, (size_t)(S.base_te + S.tot_te) * 8 + 64, (size_t)S.base_te * 8, busy)) return TC_E_CUDA;
    if (dry) return 0;
    if (host_reserve(ctx, ctx->h_nm, n1 * 4, 0, busy) || host_reserve(ctx, ctx->h_delta_off, n1 * 8, 0, busy) || host_reserve(ctx, ctx->h_delta_len, n1 * 4, 0, busy) ||
        host_reserve(ctx, ctx->h_seq_len, n1 * 4, 0, busy) ||
        host_reserve(ctx, ctx->h_cig_off, n1 * 8, 0, busy) || host_reserve(ctx, ctx->h_md_off, n1 * 8, 0, busy) || host_reserve(ctx, ctx->h_md_len, n1 * 4, 0, busy) ||
        host_reserve(ctx, ctx->h_jn_off, n1 * 8, 0, busy) ||
        host_reserve(ctx, ctx->h_delta, (size_t)(S.base_delta + S.delta_used) + 64, (size_t)S.base_delta, busy) ||
        host_reserve(ctx, ctx->h_cig_op, (size_t)(S.base_cig + S.tot_cig) + 64, (size_t)S.base_cig, busy) ||
        host_reserve(ctx, ctx->h_cig_len, (size_t)(S.base_cig + S.tot_cig) * 4 + 64, (size_t)S.base_cig * 4, busy) ||
        host_reserve(ctx, ctx->h_md, (size_t)(S.base_md + S.md_used) + 64, (size_t)S.base_md, busy) ||
        host_reserve(ctx, ctx->h_jm, (size_t)(S.base_jn + S.tot_jn) + 64, (size_t)S.base_jn, busy) ||
        host_reserve(ctx, ctx->h_ji, (size_t)(S.base_jn + S.tot_jn) * 8 + 64, (size_t)S.base_jn * 8, busy)) return TC_E_CUDA;
    return 0;
}

// enqueues the device-to-host copies of a slot's results into the batch-wide pinned arrays
static int slot_download(tc_ctx* ctx, Slot& S, int64_t r0, stream_t st) {
    const int64_t n = S.n; const size_t n1 = (size_t)n + 1;
    if (n == 0) return 0;
    if (be_d2h(ctx, st, ctx->h_status, (size_t)r0 * 4, S.W.status, (size_t)n * 4) || be_d2h(ctx, st, ctx->h_counters, (size_t)r0 * 40, S.W.counters, (size_t)n * 40) ||
        be_d2h(ctx, st, ctx->h_te_off, (size_t)r0 * 8, S.W.o_te, n1 * 8) || be_d2h(ctx, st, ctx->h_te, (size_t)S.base_te * 8, S.o_te.p, (size_t)S.tot_te * 8)) return TC_E_CUDA;
    if (S.ran_dry) return 0;
    if (be_d2h(ctx, st, ctx->h_nm, (size_t)r0 * 4, S.o_nm.p, (size_t)n * 4) || be_d2h(ctx, st, ctx->h_delta_off, (size_t)r0 * 8, S.o_delta_off.p, (size_t)n * 8) ||
        be_d2h(ctx, st, ctx->h_delta_len, (size_t)r0 * 4, S.o_delta_len.p, (size_t)n * 4) ||
        be_d2h(ctx, st, ctx->h_delta, (size_t)S.base_delta, S.o_delta_pool.p, (size_t)S.delta_used) ||
        be_d2h(ctx, st, ctx->h_seq_len, (size_t)r0 * 4, S.W.seq_len_out, (size_t)n * 4) ||
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
    out->delta_off = (const int64_t*)ctx->h_delta_off.p; out->delta_len = (const int32_t*)ctx->h_delta_len.p; out->delta = (const uint8_t*)ctx->h_delta.p;
    out->seq_len = (const int32_t*)ctx->h_seq_len.p;
    out->cig_off = (const int64_t*)ctx->h_cig_off.p; out->cig_op = (const uint8_t*)ctx->h_cig_op.p; out->cig_len = (const int32_t*)ctx->h_cig_len.p;
    out->md_off = (const int64_t*)ctx->h_md_off.p; out->md_len = (const int32_t*)ctx->h_md_len.p; out->md = (const uint8_t*)ctx->h_md.p;
    out->jn_off = (const int64_t*)ctx->h_jn_off.p; out->jm = (const int8_t*)ctx->h_jm.p; out->ji = (const int32_t*)ctx->h_ji.p;
    out->te_off = (const int64_t*)ctx->h_te_off.p; out->te = (const uint64_t*)ctx->h_te.p;
}

// notes the SEQ form of the batch; a 4-bit batch needs an alphabet that holds every genome byte
static int check_packed(tc_ctx* ctx, const tc_batch_in* in) {
    ctx->packed = in->seq_poff != nullptr;
    if (in->seq_poff && in->seq_bits == 2) return 0;                   // 2-bit rows carry their own list of other bytes: no alphabet involved
    if (!ctx->packed) return 0;
    bool have[256] = {false};
    for (int k = 0; k < ctx->n_alpha; k++) have[ctx->alphabet[k]] = true;
    for (int b = 0; b < 256; b++) if (ctx->genome_byte[b] && !have[b]) {
        ctx->err = "4-bit SEQ: the alphabet lacks genome byte " + std::to_string(b) + " (tc_ctx_set_alphabet)"; return 1;
    }
    return 0;
}

static void reset_bases(Slot& S) { S.base_cig = S.base_te = S.base_jn = S.base_md = S.base_delta = 0; }

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
    ctx->tm.launches = 0; ctx->tm.md_pool_retries = 0; ctx->tm.exact_nmmd_reads = 0; ctx->tm.ascii_path_reads = 0;
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
    // dryRun() only reads SEQ to bootstrap NM/MD (tr.py:47-48); when every primary read carries both tags the
    // column - four fifths of the upload - stays on the host
    bool all_tagged = dry != 0;
    for (int64_t i = 0; all_tagged && i < n; i++)
        if (!(in->chrom[i] < 0 || in->flag[i] == 4 || in->flag[i] > 16) && (in->tags[i] & (TAG_NM | TAG_MD)) != (TAG_NM | TAG_MD)) all_tagged = false;   // (read_class == 0)
    struct SkipGuard { tc_ctx* c; ~SkipGuard() { c->skip_seq = false; } } guard{ctx};
    ctx->skip_seq = all_tagged;
#ifndef TC_HOSTSIM
    const int64_t CH = ctx->chunk_reads;
    if (CH > 0 && n > CH + CH / 2) {
        CK(cudaSetDevice(ctx->device));
        ctx->tm.h2d_bytes = ctx->tm.d2h_bytes = 0; ctx->tm.launches = 0; ctx->tm.md_pool_retries = 0; ctx->tm.exact_nmmd_reads = 0; ctx->tm.ascii_path_reads = 0;
        ctx->tm.kernels_ms = ctx->tm.k_nmmd_init_ms = ctx->tm.k_measure_ms = ctx->tm.k_plan_ms = ctx->tm.k_junction_ms = ctx->tm.k_emit_ms = 0.f;
        ctx->n = n; ctx->ran = false;
        if (check_packed(ctx, in)) return TC_E_ARG;
        // chunk bounds: the first two chunks are a quarter and a half of the rest, so that the download
        // stream (the bottleneck) starts early; the tail chunk is whatever remains
        std::vector<int64_t> cut(1, 0);
        for (int64_t a = 0, k = 0; a < n; k++) { const int64_t sz = k == 0 ? (CH + 3) / 4 : k == 1 ? (CH + 1) / 2 : CH; a = a + sz < n ? a + sz : n; cut.push_back(a); }
        const int64_t nchunk = (int64_t)cut.size() - 1;
        int64_t bc = 0, bt = 0, bj = 0, bm = 0, bd = 0;
        CK(cudaEventRecord(ctx->ev[0], ctx->stream));
        const bool trace = getenv("TC_PIPE_TRACE") != nullptr;        // per-chunk stream timestamps on stderr (debugging aid)
        std::vector<cudaEvent_t> tev;
        auto mark = [&](cudaStream_t st) { if (trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); tev.push_back(e); } };
        mark(ctx->s_in);
        if (slot_upload(ctx, ctx->slot[0], in, 0, cut[1], ctx->s_in)) return TC_E_CUDA;
        CK(cudaEventRecord(ctx->slot[0].ev_in, ctx->s_in));
        mark(ctx->s_in);
        for (int64_t c = 0; c < nchunk; c++) {
            Slot& S = ctx->slot[c & 1];
            const int64_t r0 = cut[(size_t)c];
            if (c + 1 < nchunk) {                                   // next upload: its slot's kernels (chunk c-1) are done
                Slot& N = ctx->slot[(c + 1) & 1];
                const int64_t a = cut[(size_t)c + 1], b = cut[(size_t)c + 2];
                mark(ctx->s_in);
                if (slot_upload(ctx, N, in, a, b, ctx->s_in)) return TC_E_CUDA;
                CK(cudaEventRecord(N.ev_in, ctx->s_in));
                mark(ctx->s_in);
            }
            CK(cudaStreamWaitEvent(ctx->stream, S.ev_in, 0));
            if (c >= 2) CK(cudaStreamWaitEvent(ctx->stream, S.ev_out, 0));   // the slot's previous results have left the device
            S.base_cig = bc; S.base_te = bt; S.base_jn = bj; S.base_md = bm; S.base_delta = bd;
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
            bc += S.tot_cig; bt += S.tot_te; bj += S.tot_jn; bm += S.md_used; bd += S.delta_used;
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

#ifndef TC_HOSTSIM
__global__ void k_motifs(GenomeDev G, int64_t n, const int32_t* chrom, const int32_t* start, const int32_t* end, int8_t* code) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
        code[k] = (chrom[k] < 0 || chrom[k] >= G.n_chrom) ? (int8_t)-1 : (int8_t)intron_motif(G, chrom[k], start[k], end[k]);
}
#endif
int tc_intron_motifs(tc_ctx* ctx, int64_t n, const int32_t* chrom, const int32_t* start, const int32_t* end, int8_t* code) {
    if (!ctx || n < 0 || (n && (!chrom || !start || !end || !code))) return TC_E_ARG;
    if (!ctx->have_genome) { ctx->err = "tc_intron_motifs: load a genome first"; return TC_E_STATE; }
    if (n == 0) return TC_OK;
#ifndef TC_HOSTSIM
    CK(cudaSetDevice(ctx->device));
    DevBuf dc, ds, de, dk;
    stream_t st = main_stream(ctx);
    int rc = TC_OK;
    if (be_h2d(ctx, st, dc, 0, chrom, (size_t)n * 4) || be_h2d(ctx, st, ds, 0, start, (size_t)n * 4) || be_h2d(ctx, st, de, 0, end, (size_t)n * 4) ||
        dk.ensure((size_t)n + 8) != cudaSuccess) rc = TC_E_CUDA;
    if (!rc) {
        k_motifs<<<(int)std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(ctx->G, n, (const int32_t*)dc.p, (const int32_t*)ds.p, (const int32_t*)de.p, (int8_t*)dk.p);
        if (cudaGetLastError() != cudaSuccess || cudaMemcpyAsync(code, dk.p, (size_t)n, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess) { ctx->err = "tc_intron_motifs: CUDA error"; rc = TC_E_CUDA; }
    }
    dc.release(); ds.release(); de.release(); dk.release();
    return rc;
#else
    for (int64_t k = 0; k < n; k++) code[k] = (chrom[k] < 0 || chrom[k] >= ctx->G.n_chrom) ? (int8_t)-1 : (int8_t)intron_motif(ctx->G, chrom[k], start[k], end[k]);
    return TC_OK;
#endif
}

int tc_get_timing(tc_ctx* ctx, tc_timing* t) { if (!ctx || !t) return TC_E_ARG; *t = ctx->tm; return TC_OK; }

}  // extern "C"

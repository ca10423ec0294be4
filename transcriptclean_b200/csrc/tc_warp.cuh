// tc_warp.cuh - warp-level helpers and the shared-memory stage of the output kernel
// Part of the CUDA build of tc_b200.cu (included there; not a translation unit of its own).
#pragma once

#define FULL 0xffffffffu
// fields of a read's plan row (one word per lane, `pw`)
#define PF(k) __shfl_sync(FULL, pw, (k))
#define PF64(k) ((int64_t)(((u64)(uint32_t)PF((k) + 1) << 32) | (u64)(uint32_t)PF(k)))
#define PF64_2(lo, hi) ((int64_t)(((u64)(uint32_t)PF(hi) << 32) | (u64)(uint32_t)PF(lo)))
#ifndef WARPS_PER_BLOCK
#define WARPS_PER_BLOCK 8
#endif
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


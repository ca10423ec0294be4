// tc_emit.cuh - the output stage: k_small (TE / CIGAR / jM,jI copies) and k_emit (new SEQ + the clean-read NM/MD check)
// Part of the CUDA build of tc_b200.cu (included there; not a translation unit of its own).
#pragma once

// small_copies_read as one warp per read: TE rows into stage order by a ballot partition (as compact words), CIGAR and
// jM / jI copies, and the SEQ delta (the new SEQ as an edit script against the input SEQ, from the read's segment list),
// all from the read's 128-byte plan row.  A light kernel of its own (no shared memory, few registers, many resident
// warps): these phases are latency, not bytes, and inside k_emit they held its 30 resident warps up for a fifth of its time.
#ifndef TC_SMALL_BLOCKS
#define TC_SMALL_BLOCKS 4
#endif
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, TC_SMALL_BLOCKS) k_small(Params P, OutDev O, int dry) {
    __shared__ uint32_t lut[256];                                      // 4 bases (8 bits of the 2-bit plane) -> 4 ASCII bytes
    for (int t = threadIdx.x; t < 256; t += blockDim.x) {
        const uint32_t acgt = 0x54474341u;
        lut[t] = ((acgt >> (8 * (t & 3))) & 255u) | (((acgt >> (8 * ((t >> 2) & 3))) & 255u) << 8) |
                 (((acgt >> (8 * ((t >> 4) & 3))) & 255u) << 16) | (((acgt >> (8 * ((t >> 6) & 3))) & 255u) << 24);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31; const unsigned lt = (1u << lane) - 1u;
    const BatchDev& B = P.B; const WorkDev& W = P.W;
    const int64_t nw = (int64_t)gridDim.x * WARPS_PER_BLOCK;
    for (int64_t i = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5); i < B.n; i += nw) {
        const int pw = W.plan[i].w[lane];
        if (((uint32_t)PF(EP_STATUS) & TC_ST_CLASS_MASK) != 0) { if (lane == 0 && !dry) { O.delta_off[i] = 0; O.delta_len[i] = -1; } continue; }
        const int pflags = PF(EP_FLAGS), cig_cap = PF(EP_CIGCAP);
        const int ncig = PF(EP_NCIG), nte = PF(EP_NTE), nj = PF(EP_NJN), ns = PF(EP_NSEG);
        const int seg_cap = EP_SEGCAP_OF(cig_cap, nj), te_cap = EP_TECAP_OF(cig_cap, nj);
        const int te_mm = PF(EP_TEMM), te_ins = PF(EP_TEINS), te_del = PF(EP_TEDEL);
        const int64_t ws_off = PF64(EP_WS_LO), o_cig = PF64(EP_OCIG_LO), o_te = PF64(EP_OTE_LO), o_jn = PF64(EP_OJN_LO);
        const u64* const ws = W.ws + ws_off;
        const tc_te* const ws_te = reinterpret_cast<const tc_te*>(ws + 3ll * seg_cap + 3ll * cig_cap);
        u64* dst = O.te + o_te;
        if (dry) { for (int k = lane; k < nte; k += 32) { const tc_te r = ws_te[k]; dst[k] = te_word(r.a, r.size, r.type, r.status, r.reason); } continue; }
        const int cb = ((pflags >> 2) & 3) - 1, sb = ((pflags >> 4) & 3) - 1;
        // The lists of a read are independent of one another, but every loop below uses what it loads at once, so taken one
        // after the other they are five memory round trips in a row per read (ncu: 9 stalled warps per issue, all on loads).
        // The first round of every list is therefore asked for here, together, and each loop asks for its next round before
        // it works on the current one.
        const u64* const cigs = ws + 3ll * seg_cap + (int64_t)(cb < 0 ? 0 : cb) * cig_cap;
        const u64* const segs = ws + (int64_t)(sb < 0 ? 0 : sb) * seg_cap;
        const int32_t* const jn = reinterpret_cast<const int32_t*>(ws + 3ll * seg_cap + 3ll * cig_cap + 2ll * te_cap);
        const ulonglong2* const te2 = reinterpret_cast<const ulonglong2*>(ws_te);
        ulonglong2 te_nx = lane < nte ? te2[lane] : make_ulonglong2(0ull, 0ull);
        u64 cg_nx = (cb >= 0 && lane < ncig) ? cigs[lane] : 0ull;
        u64 sg_nx = (sb >= 0 && lane < ns) ? segs[lane] : 0ull;
        int j5 = 0, j6 = 0, j7 = 0;
        if (lane < nj) { const int32_t* r = jn + lane * JN_STRIDE; j5 = r[5]; j6 = r[6]; j7 = r[7]; }
        int base0 = 0, base1 = te_mm, base2 = te_mm + te_ins, base3 = te_mm + te_ins + te_del;
        for (int c = 0; c < nte; c += 32) {
            const int k = c + lane; const bool valid = k < nte;
            const ulonglong2 raw = te_nx;
            if (k + 32 < nte) te_nx = te2[k + 32];
            // (a workspace row: a | b << 32, size | type << 32 | status << 40 | reason << 48 - te_put in tc_plan.cuh)
            const int32_t ra = (int32_t)(uint32_t)raw.x, rb = (int32_t)(uint32_t)(raw.x >> 32), rsize = (int32_t)(uint32_t)raw.y;
            const int t = valid ? (int)((raw.y >> 32) & 255u) : -1, rstatus = (int)((raw.y >> 40) & 255u), rreason = (int)((raw.y >> 48) & 255u);
            const unsigned m0 = __ballot_sync(FULL, t == 0), m1 = __ballot_sync(FULL, t == 1), m2 = __ballot_sync(FULL, t == 2), m3 = __ballot_sync(FULL, t == 3);
            if (valid) {
                const u64 w = te_word(ra, rsize, t, rstatus, rreason);
                if (t == 3) { const int slot = base3 + 2 * __popc(m3 & lt); dst[slot] = w; dst[slot + 1] = (u64)(uint32_t)rb; }   // a junction row: two words
                else dst[t == 0 ? base0 + __popc(m0 & lt) : t == 1 ? base1 + __popc(m1 & lt) : base2 + __popc(m2 & lt)] = w;
            }
            base0 += __popc(m0); base1 += __popc(m1); base2 += __popc(m2); base3 += 2 * __popc(m3);
        }
        if (cb >= 0) {                                                 // (an untouched CIGAR is not sent back)
            uint8_t* dop = O.cig_op + o_cig; int32_t* dl = O.cig_len + o_cig;
            for (int k = lane; k < ncig; k += 32) { const u64 x = cg_nx; if (k + 32 < ncig) cg_nx = cigs[k + 32]; dop[k] = (uint8_t)cg_op(x); dl[k] = cg_len(x); }
        }
        if (nj > 0) {
            int8_t* djm = O.jm + o_jn; int32_t* dji = O.ji + 2 * o_jn;
            for (int k = lane; k < nj; k += 32) {
                if (k >= 32) { const int32_t* r = jn + k * JN_STRIDE; j5 = r[5]; j6 = r[6]; j7 = r[7]; }
                djm[k] = (int8_t)j7; dji[2 * k] = j5; dji[2 * k + 1] = j6;
            }
        }
        // ---- SEQ delta (delta_encode in tc_b200.cu is the serial twin): one pass over the segment list.  Its size is known
        //      (EP_DLEN, kept by the planner) and its place came out of the scans, so every segment's bytes go straight out.
        const int dlen = PF(EP_DLEN); const int64_t o_dl = PF64_2(EP_ODL_LO, EP_ODL_HI);
        if (lane == 0) { O.delta_off[i] = o_dl; O.delta_len[i] = dlen; }
        if (sb < 0 || dlen <= 0) continue;
        uint8_t* const dout = O.delta_pool + o_dl;
        int cur = 0, off0 = 0;                                         // input cursor (end of the latest input slice), bytes so far
        for (int s0 = 0; s0 < ns; s0 += 32) {
            const int s = s0 + lane; const bool valid = s < ns;
            const u64 sg = valid ? sg_nx : 0ull;
            if (s + 32 < ns) sg_nx = segs[s + 32];
            const int len = valid ? (int)seg_len(sg) : 0; const int64_t src = seg_src(sg);
            const bool isR = len > 0 && seg_kind(sg) == 0, isG = len > 0 && seg_kind(sg) == 1;
            const unsigned mR = __ballot_sync(FULL, isR);
            const int my_end = (int)src + len;                         // (input offsets fit 32 bits: reads are shorter than 2^24 bases)
            const unsigned prevR = mR & lt;
            const int prev_end = __shfl_sync(FULL, my_end, prevR ? 31 - __clz(prevR) : lane);
            const int gap = isR ? (int)src - (prevR ? prev_end : cur) : 0;
            const int hl = len < 62 ? 1 : len < 65536 ? 3 : 5, hg = gap <= 0 ? 0 : gap < 62 ? 1 : gap < 65536 ? 3 : 5;
            const int sz = isR ? hg + hl : isG ? hl + len : 0;
            // genome slices: the plane words are asked for now, so that they travel while the warp scans the sizes
            uint32_t code = 0, low = 0; bool plain = false;
            if (isG && len <= 16) {
                const int64_t b0 = src >> 8, b1 = (src + len - 1) >> 8;
                plain = !(((P.G.excblk[b0 >> 5] >> (b0 & 31)) | (P.G.excblk[b1 >> 5] >> (b1 & 31))) & 1u);
                const uint32_t* g2 = P.G.g2 + (src >> 4); const uint32_t* lo = P.G.lo1 + (src >> 5);
                code = __funnelshift_r(g2[0], g2[1], (int)(src & 15) * 2); low = __funnelshift_r(lo[0], lo[1], (int)(src & 31));
            }
            int tot; const int off = off0 + warp_excl_scan(sz, lane, tot);
            if (sz > 0 && off + sz <= dlen) {
                uint8_t* p = dout + off;
                if (isR) {
                    if (gap > 0) { if (hg == 1) *p++ = (uint8_t)((DL_SKIP << 6) | gap); else p += dl_put_hdr(p, DL_SKIP, gap); }
                    if (hl == 1) *p = (uint8_t)((DL_COPY << 6) | len); else dl_put_hdr(p, DL_COPY, len);
                } else {                                               // genome bases, byte exact (Q6), four at a time
                    if (hl == 1) *p++ = (uint8_t)((DL_LIT << 6) | len); else p += dl_put_hdr(p, DL_LIT, len);
                    if (plain) {
                        for (int j = 0; j < len; j += 4) {
                            const uint32_t w4 = lut[(code >> (2 * j)) & 255u] | (((((low >> j) & 15u) * 0x00204081u) & 0x01010101u) << 5);
                            const int m = len - j;
                            p[j] = (uint8_t)w4; if (m > 1) p[j + 1] = (uint8_t)(w4 >> 8); if (m > 2) p[j + 2] = (uint8_t)(w4 >> 16); if (m > 3) p[j + 3] = (uint8_t)(w4 >> 24);
                        }
                    } else for (int j = 0; j < len; j++) p[j] = genome_byte(P.G, src + j);
                }
            }
            if (mR) cur = __shfl_sync(FULL, my_end, 31 - __clz(mR));
            off0 += tot;
        }
        if (off0 != dlen && lane == 0) O.delta_len[i] = -2;            // the planner's size and the list disagree: the host refuses the batch
    }
}

// Output stage: one warp per read.  [seq_lo, seq_hi) = 32-bit words of B.seq that may be read
// (the batch's SEQ buffer on the device, padded).
//   1. plan row (128 B, one word per lane) -> everything about the read; the next read's row,
//      segment list and SEQ lines are prefetched into L2 meanwhile
//   2. (TE rows, CIGAR and jM / jI are copied by k_small above)
//   3. SEQ: built in shared memory - (A) segment list -> slice table, (B) 16-byte chunks gathered
//      from the input SEQ with the shift in force, (C) fix-ups after slice starts, (D) genome
//      slices byte-exact, (E) one coalesced 16-byte-vector store
//   4. NM/MD (getNMandMDFlags), clean-read check: a packed scan over the CIGAR gives the M runs; the
//      staged SEQ is compared with the 2-bit genome words in aligned 16-byte chunks.  A read that
//      verifies clean needs no token walk (MD = match lengths + the deletion tokens of the CIGAR);
//      anything else (a difference, a literal run, the contig end, a long read) is listed for k_nmmd_exact.
#ifndef TC_EMIT_BLOCKS
#define TC_EMIT_BLOCKS (32 / WARPS_PER_BLOCK)
#endif
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, TC_EMIT_BLOCKS) k_emit(Params P, OutDev O, int dry, int64_t seq_lo, int64_t seq_hi, const int32_t* list, const unsigned* list_n) {
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
    // `list` != 0: only the reads k_verify (tc_fast.cuh) could not settle, in the list's order
    const int64_t n_items = list ? (int64_t)*list_n : B.n;
    int64_t it = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    int64_t i = it < n_items ? (list ? (int64_t)list[it] : it) : 0;
    int pw = it < n_items ? W.plan[i].w[lane] : 0;
    for (; it < n_items; it += nw) {
        uint32_t st = (uint32_t)PF(EP_STATUS);
        const int pflags = PF(EP_FLAGS), cig_cap = PF(EP_CIGCAP), seg_cap = EP_SEGCAP_OF(PF(EP_CIGCAP), PF(EP_NJN)), nm_extra = PF(EP_NMEXTRA);
        const int ncig = PF(EP_NCIG), out_len = PF(EP_OUTLEN), ns = PF(EP_NSEG);
        const int chrom = PF(EP_CHROM); const int64_t pos = PF(EP_POS);
        const int64_t ws_off = PF64(EP_WS_LO), o_seq = PF64(EP_OSEQ_LO);
        const int64_t in_base = PF64(EP_INB_LO), cig0 = PF64(EP_CIG0_LO); const int Lq = PF(EP_LQ);
        // next read: fetch its row now (consumed at the end of this iteration), and pull its segment
        // list and SEQ lines towards L2
        const bool more = it + nw < n_items;
        const int64_t inext = more ? (list ? (int64_t)list[it + nw] : it + nw) : 0;
        const int pw_next = more ? W.plan[inext].w[lane] : 0;
#define TC_EMIT_NEXT() do { pw = pw_next; i = inext; } while (0)
        if ((st & TC_ST_CLASS_MASK) != 0) { if (lane == 0) { O.md_off[i] = 0; O.md_len[i] = 0; O.nm[i] = 0; } TC_EMIT_NEXT(); continue; }
        u64* const ws = W.ws + ws_off;
        const int cb = ((pflags >> 2) & 3) - 1, sb = ((pflags >> 4) & 3) - 1;
        if (dry) { if (lane == 0) { O.md_off[i] = 0; O.md_len[i] = 0; O.nm[i] = 0; } TC_EMIT_NEXT(); continue; }
        // (TE rows, the CIGAR and the jM / jI snapshot were copied by k_pack)
        CigSrc cs; cs.n = ncig;
        if (cb < 0) { cs.packed = 0; cs.op = B.cig_op + cig0; cs.len = B.cig_len + cig0; }
        else { cs.packed = ws + 3ll * seg_cap + (int64_t)cb * cig_cap; cs.op = 0; cs.len = 0; }
        if (dbg & 2) { TC_EMIT_NEXT(); continue; }
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
        if (more) {
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
            // (E) the row itself only leaves the SM for the reads whose NM/MD need the exact walk (below): the caller gets
            //     the new SEQ as a delta against the input SEQ (k_small)
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
                    ovf |= ct < 0;                                      // an intron of negative length (tc_plan.cuh, intron_text_rules): the exact walk
                    const unsigned mM = __ballot_sync(FULL, isM), mD = __ballot_sync(FULL, isD);
                    if (mD) {   // MD token of a D op: <matches since the previous token>^<bases>   (uniform branch: most reads keep no D)
                        const unsigned prev = mD & lt;                  // nearest earlier D in this round, else the carried one
                        const int cand = __shfl_sync(FULL, mcur + ms, prev ? 31 - __clz(prev) : lane);
                        if (isD) {
                            const int idx = nD + __popc(prev);
                            if (idx < EMIT_MAXD) { SM.d_g[idx] = (int32_t)(gcur + gs); SM.d_ct[idx] = ct; SM.d_run[idx] = mcur + ms - (prev ? cand : m_at_last_d); }
                            else ovf = true;
                        }
                        m_at_last_d = __shfl_sync(FULL, mcur + ms, 31 - __clz(mD));
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
                if (fast) {                                            // a long read: k_nmmd_exact walks it, on the row stored here
                    const int nchunk = (out_len + 15) >> 4;            // (rows are padded to 16 bytes; the tier-1 tables never overlap SM.seq)
                    for (int c = lane; c < nchunk; c += 32) reinterpret_cast<uint4*>(out_seq)[c] = *reinterpret_cast<const uint4*>(SM.seq + (c << 4));
                }
                if (lane == 0) O.exact_list[atomicAdd(O.exact_count, 1u)] = (int32_t)i;
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
        TC_EMIT_NEXT();
    }
#undef TC_EMIT_NEXT
}


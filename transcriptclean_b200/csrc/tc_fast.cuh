// tc_fast.cuh - k_verify: getNMandMDFlags (tr.py:280-342) for the reads whose new SEQ equals the genome under every M op,
// decided on the 2-bit plane of the input SEQ without building the new SEQ at all.  k_seq_plane / k_seq_plane_flags make that plane.
// Part of the CUDA build of tc_b200.cu (included there; not a translation unit of its own).
#pragma once

// The input SEQ column as 2-bit codes (A C G T = 0 1 2 3, the genome plane's code), 16 bases per word, every read's row on a
// word of its own.  ok[i] == 0: the row holds a byte outside ACGT (N, IUPAC, lower case) - such reads take k_emit's ASCII path.
//   by_poff != 0: the rows are the caller's own 2-bit upload (tc_seq_pack2): row i starts at byte poff[i] of `words`;
//   by_poff == 0: the plane was made on the device from the ASCII rows: row i starts at word ((seq_off[i] - s0) >> 4) + i.
struct SeqPlane { const uint32_t* words; const int64_t* poff; int64_t s0; const uint8_t* ok; int32_t by_poff; };

// ASCII rows -> plane rows + ok flags, one warp per read: a lane turns 16 bytes into one word
__global__ void __launch_bounds__(256) k_seq_plane(int64_t n, const int64_t* seq_off, const uint8_t* seq, int64_t s0, uint32_t* words, uint8_t* ok) {
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
        uint32_t* row = words + ((o - s0) >> 4) + i;
        const uint8_t* src = seq + o;
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(src) & ~(uintptr_t)3);
        const int sh = (int)(reinterpret_cast<uintptr_t>(src) & 3) * 8;
        bool good = true;
        for (int w = lane; w * 16 < L; w += 32) {
            const uint32_t* p = sw + w * 4;                               // (the SEQ buffer is padded: reading past the row end is safe)
            const uint32_t x0 = p[0], x1 = p[1], x2 = p[2], x3 = p[3], x4 = p[4];
            uint32_t c[4] = {__funnelshift_r(x0, x1, sh), __funnelshift_r(x1, x2, sh), __funnelshift_r(x2, x3, sh), __funnelshift_r(x3, x4, sh)};
            const int left = L - w * 16; uint32_t out = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int m = left - 4 * k;                                // bytes of this word inside the row
                if (m < 4) { const uint32_t keep = m <= 0 ? 0u : (1u << (8 * m)) - 1u; c[k] = (c[k] & keep) | (0x41414141u & ~keep); }   // 'A' behind the row end
                const uint32_t code = (((c[k] >> 1) ^ (c[k] >> 2)) & 0x03030303u) * 0x01041040u >> 24;                                   // 4 x 2 bits
                good &= lut[code] == c[k];
                out |= code << (8 * k);
            }
            row[w] = out;
        }
        good = __all_sync(0xffffffffu, good);
        if (lane == 0) ok[i] = good ? 1 : 0;
    }
}
// the caller's 2-bit upload: reads that own a listed base (exc_pos: absolute offsets into the SEQ column, seq_off likewise) are not ok
__global__ void k_seq_plane_flags(int64_t n_exc, const int64_t* exc_pos, const int64_t* seq_off, int64_t n, uint8_t* ok) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_exc; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = exc_pos[k]; int64_t lo = 0, hi = n;             // last read with seq_off[i] <= p
        while (hi - lo > 1) { const int64_t m = (lo + hi) >> 1; if (seq_off[m] <= p) lo = m; else hi = m; }
        ok[lo] = 0;
    }
}

#define FAST_MAXSEG 512
struct __align__(16) FastSmem {      // per warp
    int4 it[32];                                                         // intervals of the current round: first item, first input base, reference offset, bases
    uint16_t run_q[EMIT_MAXRUN], run_ct[EMIT_MAXRUN]; int32_t run_g[EMIT_MAXRUN];   // M ops: query offset, length, reference offset
    int32_t d_g[EMIT_MAXD], d_ct[EMIT_MAXD], d_run[EMIT_MAXD];          // D ops: reference offset, length, match run before
    int32_t d_off[EMIT_MAXD];
    uint16_t seg_o[FAST_MAXSEG];                                         // output offset of every segment
};

// One warp per read.  A read is settled here when
//   * its row of the plane is exact (ok), its new SEQ has at most EMIT_CAP bases, its CIGAR at most 128 ops, 32 M ops, EMIT_MAXD D ops,
//     no literal run of the genome lies under an M op, and the alignment stays inside the contig;
//   * every genome slice of its segment list lies inside one M op at its own reference position (equal by construction), and
//   * every stretch of an input slice under an M op equals the genome: 16 bases per lane and step, one XOR of the row's word with
//     the two genome words funnel-shifted onto it, masked at the stretch's ends.
// Then NM = inserted + deleted bases and MD = match lengths + the deletion tokens (as in k_emit's tier 1, whose CIGAR scan this
// repeats).  Every other read is listed for k_emit, which builds the row in ASCII and decides again.
#ifndef TC_FAST_BLOCKS
#define TC_FAST_BLOCKS 4
#endif
#ifndef TC_FAST_UNROLL
#define TC_FAST_UNROLL 2
#endif
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, TC_FAST_BLOCKS) k_verify(Params P, OutDev O, SeqPlane SP, int32_t* rest_list, unsigned* rest_count) {
    __shared__ FastSmem smem[WARPS_PER_BLOCK];
    FastSmem& SM = smem[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t nw = (int64_t)gridDim.x * WARPS_PER_BLOCK;
    const BatchDev& B = P.B; const WorkDev& W = P.W;
    int64_t i = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    int pw = i < B.n ? W.plan[i].w[lane] : 0;
    int64_t poff_cur = (SP.by_poff && i < B.n) ? SP.poff[i] : 0, poff_next = 0;
    for (; i < B.n; i += nw, poff_cur = poff_next) {
        uint32_t st = (uint32_t)PF(EP_STATUS);
        const int pflags = PF(EP_FLAGS), cig_cap = PF(EP_CIGCAP), seg_cap = EP_SEGCAP_OF(PF(EP_CIGCAP), PF(EP_NJN)), nm_extra = PF(EP_NMEXTRA);
        const int ncig = PF(EP_NCIG), out_len = PF(EP_OUTLEN); int ns = PF(EP_NSEG);
        const int chrom = PF(EP_CHROM); const int64_t pos = PF(EP_POS);
        const int64_t ws_off = PF64(EP_WS_LO);
        const int64_t in_base = PF64(EP_INB_LO), cig0 = PF64(EP_CIG0_LO); const int Lq = PF(EP_LQ);
        const int64_t inext = i + nw;
        const int pw_next = inext < B.n ? W.plan[inext].w[lane] : 0;
        poff_next = (SP.by_poff && inext < B.n) ? SP.poff[inext] : 0;
        if ((st & TC_ST_CLASS_MASK) != 0) { if (lane == 0) { O.md_off[i] = 0; O.md_len[i] = 0; O.nm[i] = 0; } pw = pw_next; continue; }
        if (!(pflags & 1)) {                                           // NM / MD are not refreshed
            if (st & TC_ST_NMMD_DEVICE) {                              // generated at init and kept
                const int len = W.md_init_len[i];
                const int64_t off = pool_grab(O.md_cursor, O.md_cap, len, lane);
                if (off >= 0) { const uint8_t* s = W.md_init_pool + W.md_init_off[i]; for (int j = lane; j < len; j += 32) O.md_pool[off + j] = s[j]; }
                if (lane == 0) { O.nm[i] = W.nm_init[i]; O.md_off[i] = off < 0 ? 0 : off; O.md_len[i] = off < 0 ? 0 : len; }
            } else if (lane == 0) { O.nm[i] = 0; O.md_off[i] = 0; O.md_len[i] = 0; }
            pw = pw_next; continue;
        }
        const u64* const ws = W.ws + ws_off;
        const int cb = ((pflags >> 2) & 3) - 1, sb = ((pflags >> 4) & 3) - 1;
        CigSrc cs; cs.n = ncig;
        if (cb < 0) { cs.packed = 0; cs.op = B.cig_op + cig0; cs.len = B.cig_len + cig0; }
        else { cs.packed = ws + 3ll * seg_cap + (int64_t)cb * cig_cap; cs.op = 0; cs.len = 0; }
        const u64* const segs = ws + (int64_t)(sb < 0 ? 0 : sb) * seg_cap;
        const int64_t goff = P.G.chrom_off[chrom] - 1, clen = P.G.chrom_len[chrom];
        const int64_t gbase = goff + pos;                              // plane index of the read's first reference base
        const uint32_t* const row = SP.by_poff ? SP.words + (poff_cur >> 2) : SP.words + ((in_base - SP.s0) >> 4) + i;
        // the first round of both lists and the row are asked for before the CIGAR scan needs its own loads
        u64 sg_nx = (sb >= 0 && lane < ns) ? segs[lane] : 0ull;
        bool settled = false; int nm = 0, len = 0; int64_t off = 0;
        if (SP.ok[i] && out_len <= EMIT_CAP && ncig <= 32 * 4 && (sb < 0 || ns <= FAST_MAXSEG)) {
            // ---- run list (query offset, reference offset, length) of the M ops and the D tokens, lane-parallel (k_emit, tier 1)
            int nrun = 0, qcur = 0, nD = 0, ibases = 0, dbases = 0, mcur = 0, m_at_last_d = 0, qsum = 0; int64_t gcur = 0; bool ovf = false;
            for (int k0 = 0; k0 < ncig; k0 += 32) {
                int op = 0; int32_t ct = 0;
                if (k0 + lane < ncig) cig_get(cs, k0 + lane, op, ct);
                const int qadv = (op == 'M' || op == 'I' || op == 'S') ? ct : 0;
                const int gadv = (op == 'M' || op == 'D' || op == 'N' || op == 'H') ? ct : 0;
                qsum += __reduce_add_sync(FULL, qadv);
                int qmtot, gtot; const int qm = warp_excl_scan((qadv << 16) | (op == 'M' ? (ct & 0xffff) : 0), lane, qmtot);
                const int gs = warp_excl_scan(gadv, lane, gtot);
                const int qs = (int)((unsigned)qm >> 16), ms = qm & 0xffff;
                const int qtot = (int)((unsigned)qmtot >> 16), mtot = qmtot & 0xffff;
                const bool isM = op == 'M' && ct > 0, isD = op == 'D';
                ovf |= ct < 0;
                const unsigned mM = __ballot_sync(FULL, isM), mD = __ballot_sync(FULL, isD);
                if (mD) {
                    const unsigned prev = mD & lt;
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
                    if (idx < EMIT_MAXRUN && gcur + gs < 0x7fff0000ll && ct <= 0xffff) { SM.run_q[idx] = (uint16_t)(qcur + qs); SM.run_g[idx] = (int32_t)(gcur + gs); SM.run_ct[idx] = (uint16_t)ct; }
                    else ovf = true;
                    // literal runs (N, IUPAC) under this M op?  Blocks of 256 bases, one bit each; an op of at most 4,096 bases spans
                    // at most 17 of them, i.e. bits of at most two words of the bitmap
                    const int64_t ga = gbase + gcur + gs;
                    const uint32_t b0 = (uint32_t)(ga >> 8), b1 = (uint32_t)((ga + ct - 1) >> 8);
                    const uint32_t w0 = P.G.excblk[b0 >> 5], w1 = P.G.excblk[b1 >> 5];
                    const uint32_t m0 = 0xffffffffu << (b0 & 31), m1 = 0xffffffffu >> (31 - (b1 & 31));
                    ovf |= ((b0 >> 5) == (b1 >> 5) ? (w0 & m0 & m1) : ((w0 & m0) | (w1 & m1))) != 0u;
                }
                nrun += __popc(mM); qcur += qtot; gcur += gtot; mcur += mtot;
            }
            if (inext < B.n) {                                             // the next read's lists, row and first genome words -> L2 meanwhile
                const int64_t nws = ((int64_t)__shfl_sync(FULL, pw_next, EP_WS_HI) << 32) | (uint32_t)__shfl_sync(FULL, pw_next, EP_WS_LO);
                const int64_t nin = ((int64_t)__shfl_sync(FULL, pw_next, EP_INB_HI) << 32) | (uint32_t)__shfl_sync(FULL, pw_next, EP_INB_LO);
                const int nlq = __shfl_sync(FULL, pw_next, EP_LQ), nchrom = __shfl_sync(FULL, pw_next, EP_CHROM), npos = __shfl_sync(FULL, pw_next, EP_POS);
                if (lane < 4) prefetch_l2(W.ws + nws + lane * 16);
                else if (lane < 12) { if ((lane - 4) * 512 < nlq) prefetch_l2((SP.by_poff ? SP.words + (poff_next >> 2) : SP.words + ((nin - SP.s0) >> 4) + inext) + (lane - 4) * 32); }
                else if (lane < 14) { if (nchrom >= 0 && nchrom < P.G.n_chrom) prefetch_l2(P.G.g2 + ((P.G.chrom_off[nchrom] - 1 + npos) >> 4) + (lane - 12) * 32); }
            }
            ibases = __reduce_add_sync(FULL, ibases); dbases = __reduce_add_sync(FULL, dbases);
            const bool careful = pos < 1 || pos + gcur - 1 > clen || gbase < 16;
            bool good = !careful && qsum == out_len && nD <= EMIT_MAXD && nrun <= EMIT_MAXRUN && !__any_sync(FULL, ovf);
            if (good) {
                __syncwarp();
                bool bad = false;
                // the stretches of one round (at most one per lane) as one flat space of 16-base items
                auto verify_round = [&](int cnt, int src0, int ref0, int nb) {
                    int nit; const int w = warp_excl_scan(cnt, lane, nit);
                    if (nit == 0) return;
                    const unsigned has = __ballot_sync(FULL, cnt > 0);
                    if (cnt > 0) SM.it[__popc(has & lt)] = make_int4(w, src0, ref0, nb);
                    __syncwarp();
                    const int my_w = lane < __popc(has) ? SM.it[lane].x : 0x7fffffff;
                    int rbase = 0;
                    for (int t0 = 0; t0 < nit; t0 += 32 * TC_FAST_UNROLL) {     // the loads of TC_FAST_UNROLL steps are in flight together
                        uint32_t x[TC_FAST_UNROLL], g0[TC_FAST_UNROLL], g1[TC_FAST_UNROLL], mk[TC_FAST_UNROLL]; int sh[TC_FAST_UNROLL];
#pragma unroll
                        for (int u = 0; u < TC_FAST_UNROLL; u++) {
                            const int t = t0 + 32 * u;
                            x[u] = g0[u] = g1[u] = mk[u] = 0u; sh[u] = 0;
                            if (t < nit) {                                                    // (uniform)
                                const unsigned word = __reduce_or_sync(FULL, (my_w >= t && my_w < t + 32) ? 1u << (my_w - t) : 0u);
                                const int r = rbase + __popc(word & (0xffffffffu >> (31 - lane))) - 1;
                                rbase += __popc(word);
                                if (t + lane < nit) {
                                    const int4 it = SM.it[r];
                                    const int wi = (it.y >> 4) + (t + lane - it.x);           // word of the row
                                    const int rel = (wi << 4) - it.y;                         // its first base, counted from the stretch's first base
                                    x[u] = row[wi];
                                    const int64_t ga = gbase + it.z + rel;
                                    const uint32_t* gp = P.G.g2 + (ga >> 4);
                                    g0[u] = gp[0]; g1[u] = gp[1]; sh[u] = (int)(ga & 15) * 2;
                                    const int lo = max(-rel, 0), hi = min(it.w - rel, 16);    // bases [lo, hi) of the word belong to the stretch (0 <= lo < hi)
                                    mk[u] = (0xffffffffu >> (32 - 2 * hi)) & (0xffffffffu << (2 * lo));
                                }
                            }
                        }
#pragma unroll
                        for (int u = 0; u < TC_FAST_UNROLL; u++) bad |= ((x[u] ^ __funnelshift_r(g0[u], g1[u], sh[u])) & mk[u]) != 0u;
                    }
                    __syncwarp();
                };
                // ---- segments: output offsets; genome slices by position; input slices from their start to the end of the M op there
                int obase = 0;
                const int run_step = nrun > 0 ? 1 << (31 - __clz(nrun)) : 0;                // (highest power of two <= nrun reaches every index from -1)
                if (sb < 0) { ns = 1; sg_nx = lane == 0 ? seg_make(0, 0, Lq) : 0ull; }     // the input SEQ as it is
                for (int s0 = 0; s0 < ns; s0 += 32) {
                    const int s = s0 + lane; const bool valid = s < ns;
                    const u64 sg = valid ? sg_nx : 0ull;
                    if (s + 32 < ns) sg_nx = segs[s + 32];
                    const int slen = valid ? (int)seg_len(sg) : 0; const int64_t src = seg_src(sg); const int kind = seg_kind(sg);
                    int total; const int o = obase + warp_excl_scan(slen, lane, total);
                    if (valid) SM.seg_o[s] = (uint16_t)o;
                    int r = -1;                                        // last M op that starts at or before o
                    for (int step = run_step; step >= 1; step >>= 1) { const int t = r + step; if (t < nrun && (int)SM.run_q[t] <= o) r = t; }
                    int cnt = 0, ref0 = 0, nb = 0;
                    if (slen > 0) {
                        const int rq = r >= 0 ? (int)SM.run_q[r] : 0, rend = r >= 0 ? rq + (int)SM.run_ct[r] : 0;
                        const bool inside = r >= 0 && o < rend;
                        if (inside) ref0 = SM.run_g[r] + (o - rq);
                        if (kind == 1) bad |= !(inside && o + slen <= rend && src == gbase + ref0);
                        else if (inside) { nb = min(slen, rend - o); cnt = (((int)src + nb - 1) >> 4) - ((int)src >> 4) + 1; }
                    }
                    verify_round(cnt, (int)src, ref0, nb);
                    obase += total;
                }
                bad |= obase != out_len;
                // ---- M ops that start strictly inside an input slice: from there to the end of the op or of the slice
                {
                    int cnt = 0, src0 = 0, ref0 = 0, nb = 0;
                    __syncwarp();
                    if (lane < nrun) {
                        const int q = SM.run_q[lane];
                        int s = -1;                                    // last segment that starts at or before q
                        for (int step = ns > 0 ? 1 << (31 - __clz(ns)) : 0; step >= 1; step >>= 1) { const int t = s + step; if (t < ns && (int)SM.seg_o[t] <= q) s = t; }
                        if (s >= 0) {
                            const u64 sg = sb < 0 ? seg_make(0, 0, Lq) : segs[s]; const int o = SM.seg_o[s], slen = (int)seg_len(sg);
                            if (seg_kind(sg) == 0 && q > o && q < o + slen) {
                                nb = min(o + slen, q + (int)SM.run_ct[lane]) - q; src0 = (int)seg_src(sg) + (q - o); ref0 = SM.run_g[lane];
                                cnt = ((src0 + nb - 1) >> 4) - (src0 >> 4) + 1;
                            }
                        }
                    }
                    verify_round(cnt, src0, ref0, nb);
                }
                if (!__any_sync(FULL, bad)) {                          // every aligned base equals the genome
                    const int total_m = mcur;
                    nm = ibases + dbases;
                    if (nD == 0) {                                     // MD is one number of at most 4 digits: the read's inline slot
                        len = total_m > 0 ? ndigits((uint32_t)total_m) : 0;
                        off = O.md_inline + 8 * i;
                        if (lane == 0) *reinterpret_cast<unsigned long long*>(O.md_pool + off) = dec_word4((uint32_t)total_m, len);
                    } else {                                           // one token per D op, then the trailing run
                        int tlen = 0;
                        for (int k0 = 0; k0 < nD; k0 += 32) {
                            const int k = k0 + lane;
                            const int t = k < nD ? ndigits((uint32_t)SM.d_run[k]) + 1 + SM.d_ct[k] : 0;
                            int tot; const int to = tlen + warp_excl_scan(t, lane, tot);
                            if (k < nD) SM.d_off[k] = to;
                            tlen += tot;
                        }
                        const int final_run = total_m - m_at_last_d, ndf = final_run > 0 ? ndigits((uint32_t)final_run) : 0;
                        len = tlen + ndf;
                        off = len <= 8 ? O.md_inline + 8 * i : pool_grab(O.md_cursor, O.md_cap, len, lane);
                        __syncwarp();
                        if (off >= 0) {
                            uint8_t* out = O.md_pool + off;
                            for (int k = lane; k < nD; k += 32) { const int nd = ndigits((uint32_t)SM.d_run[k]); write_dec(out + SM.d_off[k], (uint32_t)SM.d_run[k], nd); out[SM.d_off[k] + nd] = '^'; }
                            for (int k = 0; k < nD; k++) {
                                uint8_t* bp = out + SM.d_off[k] + ndigits((uint32_t)SM.d_run[k]) + 1; const int64_t g = gbase + SM.d_g[k]; const int ct = SM.d_ct[k];
                                for (int j = lane; j < ct; j += 32) bp[j] = genome_byte(P.G, g + j);
                            }
                            if (ndf > 0 && lane == 0) write_dec(out + tlen, (uint32_t)final_run, ndf);
                        }
                    }
                    settled = true;
                }
            }
        }
        if (settled) {
            st = (st & ~(uint32_t)TC_ST_NMMD_NONE) | TC_ST_NMMD_DEVICE;
            if (lane == 0) { O.nm[i] = nm + nm_extra; O.md_off[i] = off < 0 ? 0 : off; O.md_len[i] = off < 0 ? 0 : len; W.status[i] = st; }
        } else if (lane == 0) rest_list[atomicAdd(rest_count, 1u)] = (int32_t)i;
        __syncwarp();
        pw = pw_next;
    }
}

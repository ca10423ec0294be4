// tc_nmmd.cuh - getNMandMDFlags (transcript.py:280-342) as warp code: the token-emitting walk, the NM/MD bootstrap of tag-less reads, the exact pass behind k_emit
// Part of the CUDA build of tc_b200.cu (included there; not a translation unit of its own).
#pragma once

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
        int64_t span = 0; bool neg = false;                 // (negative lengths - a rescued micro-intron - only widen the span that is checked)
        for (int k = lane; k < cs.n; k += 32) { int op; int32_t ct; cig_get(cs, k, op, ct); if (op == 'M' || op == 'D' || op == 'N' || op == 'H') span += ct > 0 ? ct : 0; neg |= ct < 0; }
#pragma unroll
        for (int d = 16; d; d >>= 1) span += __shfl_xor_sync(FULL, span, d);
        careful = pos < 1 || pos + span - 1 > clen || span > 0x7fffffffll || __any_sync(FULL, neg);
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
        if (((w[EP_FLAGS] >> 2) & 3) == 0) {                           // the input CIGAR (never copied to the outputs)
            const int64_t c0 = ((int64_t)w[EP_CIG0_HI] << 32) | (uint32_t)w[EP_CIG0_LO];
            cs.op = P.B.cig_op + c0; cs.len = P.B.cig_len + c0;
        }
        int nm, len; int64_t off; bool none;
        nmmd_to_pool<true>(P.G, O.seq + o_seq, lut, cs, w[EP_CHROM], (int64_t)w[EP_POS], lane, O.md_cursor, O.md_cap, O.md_pool, O.md_inline + 8 * i, nm, len, off, none);
        uint32_t st = ((uint32_t)w[EP_STATUS] & ~(uint32_t)TC_ST_NMMD_NONE) | TC_ST_NMMD_DEVICE;
        if (none) st |= TC_ST_NMMD_NONE;
        if (lane == 0) { O.nm[i] = nm + w[EP_NMEXTRA]; O.md_off[i] = off < 0 ? 0 : off; O.md_len[i] = off < 0 ? 0 : len; P.W.status[i] = st; }
    }
}


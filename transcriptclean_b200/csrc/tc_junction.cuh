// tc_junction.cuh - junction objects (k_jn_init) and the warp-per-read noncanonical junction rescue (k_jn_fix)
// Part of the CUDA build of tc_b200.cu (included there; not a translation unit of its own).
#pragma once

// junction objects for every read; reads that need cleanNoncanonical are appended to `list`
#ifndef TC_JNINIT_BLOCKS
#define TC_JNINIT_BLOCKS 6
#endif
__global__ void __launch_bounds__(128, TC_JNINIT_BLOCKS) k_jn_init(Params P, int32_t* list, unsigned* count) {
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
// dl_list_size (tc_plan.cuh) across the lanes: bytes of the SEQ delta of a segment list
__device__ int w_delta_size(const u64* s, int n, int lane) {
    const unsigned lt = (1u << lane) - 1u;
    int total = 0, cur = 0;
    for (int j0 = 0; j0 < n; j0 += 32) {
        const int j = j0 + lane; const u64 sg = j < n ? s[j] : 0ull;
        const int len = j < n ? (int)seg_len(sg) : 0; const int src = (int)seg_src(sg);
        const bool isR = len > 0 && seg_kind(sg) == 0, isG = len > 0 && seg_kind(sg) == 1;
        const unsigned mR = __ballot_sync(FULL, isR), prevR = mR & lt;
        const int my_end = src + len;
        const int prev_end = __shfl_sync(FULL, my_end, prevR ? 31 - __clz(prevR) : lane);
        const int gap = isR ? src - (prevR ? prev_end : cur) : 0;
        total += isR ? (gap > 0 ? dl_hdr_len(gap) : 0) + dl_hdr_len(len) : isG ? dl_hdr_len(len) + len : 0;
        if (mR) cur = __shfl_sync(FULL, my_end, 31 - __clz(mR));
    }
    return __reduce_add_sync(FULL, total);
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
                          bool& raised, bool& unsupported, int lane) {
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
    const int nsurv = sb - sa;
    for (int k = lane; k < n; k += 32) {
        u64 c = src[k];
        if (k < lo || k >= hi) dst[k < lo ? k : lo + ins_front + nsurv + ins_back + (k - hi)] = c;
        else if (k >= sa && k < sb) {
            if (k == km) c = cg_make(cg_op(c), newlen);
            dst[lo + ins_front + (k - sa)] = c;
        }
    }
    if (lane == 0) { if (ins_front) dst[lo] = cg_make('M', nb); if (ins_back) dst[lo + nsurv] = cg_make('M', nb); }
    const int m = lo + ins_front + nsurv + ins_back + (n - hi);
    __syncwarp();
    if (intron_idx >= 0) {                                             // the intron whose length changes, under the text rules (tc_plan.cuh)
        const int idx = intron_idx < lo ? intron_idx : lo + ins_front + nsurv + ins_back + (intron_idx - hi);
        const int32_t L = cg_len(src[intron_idx]) + intron_delta;
        bool ok = true, uns = false;
        if (lane == 0) ok = intron_text_rules(dst, idx, L, uns);
        ok = __shfl_sync(FULL, ok ? 1 : 0, 0) != 0; uns = __shfl_sync(FULL, uns ? 1 : 0, 0) != 0;
        __syncwarp();
        bool bad = false;                                              // other_dminus_ok across the lanes
        for (int k = lane; k + 1 < m; k += 32)
            bad |= cg_op(dst[k]) == CG_OP_DMINUS && k + 1 != idx && cg_op(dst[k + 1]) == 'N' && cg_len(dst[k + 1]) < 0;
        if (!ok || __any_sync(FULL, bad)) raised = true;
        if (uns) unsupported = true;
    }
    return m;
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
                           u64* cig_dst, u64* seg_dst, bool& unsupported, int lane) {
    if (d == 0) return true;
    bool raised = false;
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
        const int ncn = w_cig_edit(S.cig, S.nc, cig_dst, e.lo, e.hi, true, d, e.nidx, -d, raised, unsupported, lane);
        if (ncn < 0 || raised) return false;
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
        const int ncn = w_cig_edit(S.cig, S.nc, cig_dst, e.lo, e.hi, false, -d, ek.nidx, d, raised, unsupported, lane);
        if (ncn < 0 || raised) return false;
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
    bool whole_fail = false, unsupported = false, any_ok = false;
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
            ok = w_fix_side(P, chrom, k, side, d, &r[side], bnd[side], S, v.cig[tc], v.seg[ts], unsupported, lane); used_c = tc; used_s = ts;
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
        if (any_ok) st |= TC_ST_CIGAR_REWRITTEN | TC_ST_JM_DEVICE | TC_ST_JI_DEVICE;
        if (unsupported) st |= TC_ST_ERR_INTRON;
        const int dsize = any_ok ? w_delta_size(v.seg[sb], ns, lane) : 0;
        if (lane == 0) {
            C[TC_C_COR_SJ] = cJ; C[TC_C_UNC_SJ] = uJ;
            W.n_te[i] = nte;
            if (any_ok) {
                W.cig_buf[i] = (int8_t)cb; W.seg_buf[i] = (int8_t)sb; W.n_cig_cur[i] = nc; W.n_seg_cur[i] = ns; W.seq_len_out[i] = (int32_t)seqlen;
                W.delta_size[i] = dsize;
                W.refresh[i] = 1; W.nm_extra[i] = 0;
            }
        }
    }
    if (canon) st |= TC_ST_CANONICAL;
    if (annot) st |= TC_ST_ALL_ANNOT;
    if (lane == 0) W.status[i] = st;
}

#ifndef TC_JNFIX_BLOCKS
#define TC_JNFIX_BLOCKS (32 / WARPS_PER_BLOCK)
#endif
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, TC_JNFIX_BLOCKS) k_jn_fix(Params P, const int32_t* list, const unsigned* count) {
    const int lane = threadIdx.x & 31;
    const int64_t nw = (int64_t)gridDim.x * WARPS_PER_BLOCK, nl = (int64_t)*count;
    for (int64_t t = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5); t < nl; t += nw) junction_fix_warp(P, list[t], lane);
}

// tc_host.cpp - host-side SAM ingest and text rendering (C++, multi-threaded).
//
// the callers either side of the correction path.  Replaces, with identical
// results, the Python-level work of the reference around correct_transcript():
//   split_SAM                      TranscriptClean.py:519-541   header / alignment lines
//   init_log_info                  TranscriptClean.py:1558-1588 mapping class
//   Transcript.__init__            transcript.py:10-72          field and tag handling
//   printableSAM / printableFa     transcript.py:244-268, reverseComplement 541-576
//   create_log_string              TranscriptClean.py:1591-1604
//   TE rows                        TC.py:911-936, 1017-1046, 1119-1134, 1227-1228, 1650-1668
//   batch_correct emission rules   TC.py:373-385 (Q19)
// transcriptclean_b200/sam.py and render.py are the same logic in Python; tests compare the two.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/tc_b200.h"

// page-locked host memory is the CUDA translation unit's business: tc_b200.cu installs these two (null: plain malloc)
void* (*tc_pinned_alloc)(size_t) = nullptr;
void (*tc_pinned_free)(void*) = nullptr;

namespace {

struct Span { int64_t s; int32_t n; };

// std::vector whose resize() leaves new elements uninitialised: the column buffers are resized to their final size and then
// filled by the parser threads, and value-initialising them first was a serial pass over every byte of the batch (most of
// tc_sam_parse's time on 16 threads: profiles/r02f_host_text_probe.json)
template <class T> struct DefaultInitAlloc : std::allocator<T> {
    template <class U> struct rebind { using other = DefaultInitAlloc<U>; };
    using std::allocator<T>::allocator;
    template <class U, class... A> void construct(U* p, A&&... a) {
        if constexpr (sizeof...(A) == 0) ::new ((void*)p) U; else ::new ((void*)p) U(std::forward<A>(a)...);
    }
};
template <class T> using UV = std::vector<T, DefaultInitAlloc<T>>;

// The same for the batch's large columns (SEQ, CIGAR, MD), in page-locked memory: tc_correct_batch then copies them to the device
// straight from where the parser put them (a pageable source goes through the driver's staging buffer first: two more passes
// over every byte on the host).  The objects are pooled, so the cost of page-locking is paid once per pipeline worker.
// Falls back to malloc when no device is there (or the hooks are not installed); a 64-byte header in front of the block says which it was.
template <class T> struct PinnedAlloc : DefaultInitAlloc<T> {
    template <class U> struct rebind { using other = PinnedAlloc<U>; };
    PinnedAlloc() = default;
    template <class U> PinnedAlloc(const PinnedAlloc<U>&) {}
    T* allocate(size_t n) {
        const size_t bytes = n * sizeof(T) + 64; void* p = nullptr; uint64_t kind = 0;
        if (bytes >= ((size_t)1 << 20) && tc_pinned_alloc && (p = tc_pinned_alloc(bytes)) != nullptr) kind = 1;
        if (!p) { p = malloc(bytes); if (!p) throw std::bad_alloc(); }
        *reinterpret_cast<uint64_t*>(p) = kind;
        return reinterpret_cast<T*>(reinterpret_cast<char*>(p) + 64);
    }
    void deallocate(T* q, size_t) {
        void* p = reinterpret_cast<char*>(q) - 64;
        if (*reinterpret_cast<uint64_t*>(p) == 1) { tc_pinned_free(p); return; }
        free(p);
    }
};
template <class T> using PV = std::vector<T, PinnedAlloc<T>>;

struct Part {                       // what one parser thread produces
    UV<int32_t> flag, chrom, pos; UV<uint8_t> tags, cls;
    UV<int64_t> seq_len, cig_n, md_n, ji_n;
    UV<uint8_t> cig_op, md; UV<int32_t> cig_len, ji;
    UV<Span> seq_at; size_t seq_bytes = 0;   // where every read's SEQ lies in the text (copied once, straight into the batch's column)
    UV<Span> line, f[9], cigar, tg[4];
    UV<int64_t> other_off; UV<Span> other;
    UV<Span> header;
    std::string err;
    void reset() {                   // empty, capacity kept: the parts of a host thread are reused from batch to batch
        flag.clear(); chrom.clear(); pos.clear(); tags.clear(); cls.clear(); seq_len.clear(); cig_n.clear(); md_n.clear(); ji_n.clear();
        seq_at.clear(); seq_bytes = 0; cig_op.clear(); md.clear(); cig_len.clear(); ji.clear(); line.clear(); cigar.clear(); other_off.clear(); other.clear();
        for (auto& v : f) v.clear();
        for (auto& v : tg) v.clear();
        header.clear(); err.clear();
    }
};

template <class V> inline void put_bytes(V& v, const char* p, int64_t n) {
    const size_t o = v.size(); v.resize(o + (size_t)n); if (n) memcpy(v.data() + o, p, (size_t)n);
}
inline bool is_ws(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

// Python int(): optional surrounding whitespace, optional sign, digits (underscores not supported)
bool py_int(const char* p, int n, int64_t& out) {
    int a = 0, b = n;
    while (a < b && is_ws(p[a])) a++;
    while (b > a && is_ws(p[b - 1])) b--;
    if (a >= b) return false;
    bool neg = false;
    if (p[a] == '+' || p[a] == '-') { neg = p[a] == '-'; a++; }
    if (a >= b) return false;
    int64_t v = 0;
    for (int k = a; k < b; k++) { if (p[k] < '0' || p[k] > '9') return false; v = v * 10 + (p[k] - '0'); if (v > (1ll << 40)) return false; }
    out = neg ? -v : v;
    return true;
}

}  // namespace

namespace { struct SamBufs; }
struct tc_sam {
    const char* text = nullptr; int64_t len = 0;
    SamBufs* bufs = nullptr;         // the parser threads' parts and the formatter threads' texts: storage that travels with the object through the pool
    ~tc_sam();
    int64_t n = 0;
    UV<int32_t> flag, chrom, pos; UV<uint8_t> tags, cls;
    UV<int64_t> seq_off, cig_off, md_off, ji_off;
    PV<uint8_t> seq, cig_op, md; PV<int32_t> cig_len; UV<int32_t> ji;
    UV<Span> line, f[9], cigar, tg[4];
    UV<int64_t> other_off; UV<Span> other;
    std::string header;
    struct RawBuf {                  // output text: grown without initialising (a std::string resize would touch it twice)
        char* p = nullptr; size_t cap = 0, n = 0;
        bool ensure(size_t want) {
            if (want <= cap) return true;
            const size_t c2 = want + want / 8 + 4096; char* q = (char*)malloc(c2);
            if (!q) return false;                                     // (the old buffer stays as it was)
            free(p); p = q; cap = c2; return true;
        }
        ~RawBuf() { free(p); }
    } out[4];                        // sam, fa, log, te
    // 4-bit SEQ form (tc_sam_pack_seq): packed rows, their offsets, the alphabet they index
    UV<uint8_t> seqp; UV<int64_t> seq_poff; uint8_t alphabet[16] = {0}; int n_alpha = 0; bool packed = false;
};

namespace {

void set_err(char* err, int64_t cap, const std::string& m) {
    if (err && cap > 0) { snprintf(err, (size_t)cap, "%s", m.c_str()); }
}

void parse_range_body(const char* T, int64_t a, int64_t b, const std::unordered_map<std::string, int>& cmap, const int64_t* clen, Part& P) {
    int64_t p = a;
    std::vector<Span> fld;
    {   // sized from the byte range: a long-read SAM line is mostly SEQ (and QUAL when present)
        const size_t bytes = (size_t)(b - a), est = bytes / 1500 + 16;
        P.seq_at.reserve(est); P.cig_op.reserve(est * 40); P.cig_len.reserve(est * 40); P.md.reserve(est * 80);
        P.flag.reserve(est); P.chrom.reserve(est); P.pos.reserve(est); P.tags.reserve(est); P.cls.reserve(est); P.line.reserve(est);
        for (int k = 0; k < 9; k++) P.f[k].reserve(est);
        for (int k = 0; k < 4; k++) P.tg[k].reserve(est);
        P.cigar.reserve(est); P.other.reserve(est * 3); P.other_off.reserve(est);
        P.seq_len.reserve(est); P.cig_n.reserve(est); P.md_n.reserve(est); P.ji_n.reserve(est);
    }
    Span last_chrom = {0, 0}; int last_cid = -1;
    while (p < b && P.err.empty()) {
        const char* nlp = (const char*)memchr(T + p, '\n', (size_t)(b - p));
        int64_t e = nlp ? nlp - T : b;
        int64_t s = p, t = e;
        p = e + 1;
        while (s < t && is_ws(T[s])) s++;
        while (t > s && is_ws(T[t - 1])) t--;
        if (s >= t) continue;
        if (T[s] == '@') { P.header.push_back({s, (int32_t)(t - s)}); continue; }
        fld.clear();
        for (int64_t q = s;;) {
            const char* tp = (const char*)memchr(T + q, '\t', (size_t)(t - q));
            const int64_t k = tp ? tp - T : t;
            fld.push_back({q, (int32_t)(k - q)});
            if (!tp) break;
            q = k + 1;
        }
        if (fld.size() < 11) { P.err = "SAM record with fewer than 11 fields: " + std::string(T + s, (size_t)std::min<int64_t>(80, t - s)); break; }
#define qname std::string(T + fld[0].s, (size_t)fld[0].n)
        int64_t fl;
        if (!py_int(T + fld[1].s, fld[1].n, fl)) { P.err = "read " + qname + ": FLAG is not an integer"; break; }
        const bool star = fld[2].n == 1 && T[fld[2].s] == '*';
        const int cls = (star || fl == 4) ? 1 : (fl > 16 ? 2 : 0);
        P.flag.push_back((int32_t)fl); P.cls.push_back((uint8_t)cls); P.line.push_back({s, (int32_t)(t - s)});
        for (int k = 0; k < 9; k++) P.f[k].push_back(fld[k]);
        if (cls != 0) {
            P.chrom.push_back(cls == 1 ? -1 : 0); P.pos.push_back(0); P.tags.push_back(0);
            P.seq_len.push_back(0); P.cig_n.push_back(0); P.md_n.push_back(0); P.ji_n.push_back(0);
            P.cigar.push_back({0, 0}); for (int k = 0; k < 4; k++) P.tg[k].push_back({0, 0});
            P.other_off.push_back((int64_t)P.other.size());
            continue;
        }
        int c;
        if (last_cid >= 0 && last_chrom.n == fld[2].n && memcmp(T + last_chrom.s, T + fld[2].s, (size_t)fld[2].n) == 0) c = last_cid;
        else {
            auto it = cmap.find(std::string(T + fld[2].s, (size_t)fld[2].n));
            if (it == cmap.end()) { P.err = "read " + qname + ": chromosome " + std::string(T + fld[2].s, (size_t)fld[2].n) + " is not in the reference genome"; break; }
            c = it->second; last_cid = c; last_chrom = fld[2];
        }
        int64_t posv;
        if (!py_int(T + fld[3].s, fld[3].n, posv)) { P.err = "read " + qname + ": POS is not an integer"; break; }
        // CIGAR: ^(?:[0-9]+[A-Z])+$
        const char* cg = T + fld[5].s; const int cn = fld[5].n;
        int64_t qlen = 0, gend = posv; bool okc = cn > 0, hasN = false; int nops = 0;
        for (int k = 0; k < cn && okc;) {
            int64_t v = 0; int d = 0;
            while (k < cn && cg[k] >= '0' && cg[k] <= '9') { v = v * 10 + (cg[k] - '0'); k++; d++; if (v > (1ll << 40)) okc = false; }
            if (d == 0 || k >= cn || cg[k] < 'A' || cg[k] > 'Z') { okc = false; break; }
            const char op = cg[k++];
            if (v < 1) { P.err = "read " + qname + ": zero-length CIGAR operation"; break; }
            P.cig_op.push_back((uint8_t)op); P.cig_len.push_back((int32_t)v); nops++;
            if (op == 'M' || op == 'I' || op == 'S') qlen += v;
            if (op == 'M' || op == 'D' || op == 'N' || op == 'H') gend += v;
            if (op == 'N') hasN = true;
        }
        if (!P.err.empty()) break;
        if (!okc) { P.err = "read " + qname + ": unsupported CIGAR '" + std::string(cg, (size_t)std::min(cn, 40)) + "'"; break; }
        if (qlen != fld[9].n) { P.err = "read " + qname + ": SEQ length " + std::to_string(fld[9].n) + " does not match CIGAR (" + std::to_string(qlen) + ")"; break; }
        if (posv < 1 || gend - 1 > clen[c]) { P.err = "read " + qname + ": alignment runs off contig " + std::string(T + fld[2].s, (size_t)fld[2].n); break; }
        Span tg[4] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};
        P.other_off.push_back((int64_t)P.other.size());
        for (size_t k = 11; k < fld.size(); k++) {
            const char* x = T + fld[k].s; const int n = fld[k].n;
            int which = -1;
            if (n >= 2) {
                if (x[0] == 'N' && x[1] == 'M') which = 0; else if (x[0] == 'M' && x[1] == 'D') which = 1;
                else if (x[0] == 'j' && x[1] == 'M') which = 2; else if (x[0] == 'j' && x[1] == 'I') which = 3;
            }
            if (which >= 0) tg[which] = fld[k]; else P.other.push_back(fld[k]);
        }
        uint8_t tb = 0; int64_t mdn = 0, jin = 0;
        if (tg[0].n) tb |= 1;
        if (tg[1].n) {
            tb |= 2;
            const char* x = T + tg[1].s; const int n = tg[1].n;
            int c1 = -1, c2 = -1, c3 = n;
            for (int k = 0; k < n; k++) if (x[k] == ':') { if (c1 < 0) c1 = k; else if (c2 < 0) c2 = k; else { c3 = k; break; } }
            if (c2 < 0) { P.err = "read " + qname + ": malformed MD tag"; break; }
            if (tg[0].n) { put_bytes(P.md, x + c2 + 1, c3 - c2 - 1); mdn = c3 - c2 - 1; }
        }
        if (tg[2].n) tb |= 4;
        if (tg[3].n) {
            tb |= 8;
            if (hasN) {
                const char* x = T + tg[3].s; const int n = tg[3].n;
                int lc = -1; for (int k = 0; k < n; k++) if (x[k] == ':') lc = k;
                int k = lc + 1; int idx = 0; bool bad = false;
                while (k <= n) {
                    int e2 = k; while (e2 < n && x[e2] != ',') e2++;
                    if (idx > 0) {
                        int64_t v;
                        if (!py_int(x + k, e2 - k, v)) { P.err = "read " + qname + ": malformed jI tag"; bad = true; break; }
                        if (v < 2 || v > clen[c]) { P.err = "read " + qname + ": jI tag does not describe intron pairs on the contig"; bad = true; break; }
                        P.ji.push_back((int32_t)v); jin++;
                    }
                    idx++; k = e2 + 1;
                }
                if (bad) break;
                if (jin % 2) { P.err = "read " + qname + ": jI tag does not describe intron pairs on the contig"; break; }
            }
        }
        P.chrom.push_back(c); P.pos.push_back((int32_t)posv); P.tags.push_back(tb);
        P.seq_at.push_back(fld[9]); P.seq_bytes += (size_t)fld[9].n;
        P.seq_len.push_back(fld[9].n); P.cig_n.push_back(nops); P.md_n.push_back(mdn); P.ji_n.push_back(jin);
        P.cigar.push_back(fld[5]); for (int k = 0; k < 4; k++) P.tg[k].push_back(tg[k]);
    }
#undef qname
}

// (a thread body: an allocation failure must come back as the part's error, not end the process)
void parse_range(const char* T, int64_t a, int64_t b, const std::unordered_map<std::string, int>& cmap, const int64_t* clen, Part& P) {
    try { parse_range_body(T, a, b, cmap, clen, P); }
    catch (const std::exception& e) { P.err = std::string("tc_sam_parse: ") + e.what(); }
}

template <class T> void append(std::vector<T>& d, const std::vector<T>& s) { d.insert(d.end(), s.begin(), s.end()); }

const char* const CLASS_NAME[3] = {"primary", "unmapped", "non-primary"};
const char* const TE_TYPE[4] = {"Mismatch", "Insertion", "Deletion", "NC_SJ_boundary"};
const char* const TE_STATUS[2] = {"Corrected", "Uncorrected"};
const char* const TE_REASON[6] = {"NA", "VariantMatch", "TooLarge", "TooFarFromAnnotJn", "Other", "DryRun"};
const size_t TE_TYPE_N[4] = {8, 9, 8, 14};
struct TeTails {                    // "<Corrected|Uncorrected>\t<reason>\n" for every (status, reason)
    std::string s[2][8];
    TeTails() { for (int a = 0; a < 2; a++) for (int b = 0; b < 8; b++) s[a][b] = std::string(TE_STATUS[a]) + "\t" + (b < 6 ? TE_REASON[b] : "NA") + "\n"; }
};
const TeTails TE_TAILS;
#define TE_TAIL TE_TAILS_P
#define TE_TAIL_N TE_TAILS_N
struct TeTailsP { const char* p[2][8]; size_t n[2][8]; TeTailsP() { for (int a = 0; a < 2; a++) for (int b = 0; b < 8; b++) { p[a][b] = TE_TAILS.s[a][b].data(); n[a][b] = TE_TAILS.s[a][b].size(); } } };
const TeTailsP TE_TP;
#define TE_TAILS_P TE_TP.p
#define TE_TAILS_N TE_TP.n

// Text under construction.  The string is only storage (its size is the capacity in use); `n` is the length
// written so far.  need(k) makes room for k more bytes once per record; the writers below do not check.
struct Out {
    std::string buf; size_t n = 0;
    void clear() { n = 0; }
    void need(size_t k) { if (n + k > buf.size()) buf.resize(std::max(buf.size() * 2, n + k + ((size_t)1 << 20))); }
    void put(const char* p, size_t k) { memcpy(&buf[n], p, k); n += k; }
    void ch(char c) { buf[n++] = c; }
    void str(const char* z) { put(z, strlen(z)); }
    void num(int64_t v) {                                        // digits written in place, two at a time
        static const char D2[201] = "00010203040506070809101112131415161718192021222324252627282930313233343536373839404142434445464748495051525354555657585960616263646566676869707172737475767778798081828384858687888990919293949596979899";
        uint64_t u = v < 0 ? (uint64_t)(-v) : (uint64_t)v;
        if (v < 0) buf[n++] = '-';
        int nd = 1; for (uint64_t t = u; t >= 10; t /= 10) nd++;
        char* e = &buf[n] + nd; n += (size_t)nd;
        while (u >= 100) { const unsigned r = (unsigned)(u % 100); u /= 100; e -= 2; e[0] = D2[2 * r]; e[1] = D2[2 * r + 1]; }
        if (u >= 10) { e -= 2; e[0] = D2[2 * u]; e[1] = D2[2 * u + 1]; } else { e[-1] = (char)('0' + u); }
    }
};

struct Comp { uint8_t t[256]; Comp() { for (int k = 0; k < 256; k++) t[k] = (uint8_t)k; const char* a = "ACGTNacgtn*"; const char* b = "TGCANtgcan*"; for (int k = 0; a[k]; k++) t[(uint8_t)a[k]] = (uint8_t)b[k]; } };
const Comp COMP;

// tc_seq_apply_delta (include/tc_b200.h) with block copies
int64_t apply_delta(const uint8_t* in, int64_t in_len, const uint8_t* dl, int64_t dn, uint8_t* out, int64_t cap) {
    if (dn < 0) { if (in_len > cap) return -1; memcpy(out, in, (size_t)in_len); return in_len; }
    int64_t q = 0, o = 0, k = 0;
    while (k < dn) {
        const int kind = dl[k] >> 6; int64_t n = dl[k] & 63; k++;
        if (n == 62) { if (k + 2 > dn) return -1; n = (int64_t)dl[k] | ((int64_t)dl[k + 1] << 8); k += 2; }
        else if (n == 63) { if (k + 4 > dn) return -1; n = (int64_t)dl[k] | ((int64_t)dl[k + 1] << 8) | ((int64_t)dl[k + 2] << 16) | ((int64_t)dl[k + 3] << 24); k += 4; }
        if (kind == 1) { q += n; continue; }
        if (o + n > cap) return -1;
        if (kind == 0) { if (q + n > in_len) return -1; memcpy(out + o, in + q, (size_t)n); q += n; }
        else if (kind == 2) { if (k + n > dn) return -1; memcpy(out + o, dl + k, (size_t)n); k += n; }
        else return -1;
        o += n;
    }
    return o;
}

void render_range_body(const tc_sam* S, const tc_batch_out* R, int64_t a, int64_t b, bool primary_only, bool canon_only, bool dry,
                       Out& sam, Out& fa, Out& lg, Out& te, int64_t& bad_delta) {
    const char* T = S->text;
    for (int64_t i = a; i < b; i++) {
        const uint32_t st = R->status[i]; const int cls = (int)(st & TC_ST_CLASS_MASK);
        const Span q = S->f[0][i], ch = S->f[2][i];
        // log row: TranscriptID, Mapping, ten counters ("NA" = -1)
        lg.need((size_t)q.n + 160);
        lg.put(T + q.s, (size_t)q.n); lg.ch('\t'); lg.str(CLASS_NAME[cls]);
        for (int k = 0; k < 10; k++) { lg.ch('\t'); const int32_t c = R->counters[10 * i + k]; if (c < 0) lg.put("NA", 2); else lg.num(c); }
        lg.ch('\n');
        if (cls != 0) {
            if (!dry && !primary_only && !canon_only) { sam.need((size_t)S->line[i].n + 1); sam.put(T + S->line[i].s, (size_t)S->line[i].n); sam.ch('\n'); }
            continue;
        }
        const int64_t t0 = R->te_off[i], t1 = R->te_off[i + 1];
        te.need((size_t)(t1 - t0) * ((size_t)q.n + (size_t)ch.n + 96));
        if (t1 > t0) {                                                // every row of the read starts with "<QNAME>\t<chrom>_"
            const size_t pre0 = te.n, pre_n = (size_t)q.n + (size_t)ch.n + 2;
            te.put(T + q.s, (size_t)q.n); te.ch('\t'); te.put(T + ch.s, (size_t)ch.n); te.ch('_');
            for (int64_t k = t0; k < t1;) {
                tc_te r; k += tc_te_decode(R->te + k, &r);
                if (te.n != pre0 + pre_n) { memcpy(&te.buf[te.n], &te.buf[pre0], pre_n); te.n += pre_n; }
                te.num(r.a);
                if (r.type != 0) { te.ch('_'); te.num(r.b); }
                te.ch('\t'); te.put(TE_TYPE[r.type], TE_TYPE_N[r.type]); te.ch('\t'); te.num(r.size); te.ch('\t');
                te.put(TE_TAIL[r.status][r.reason], TE_TAIL_N[r.status][r.reason]);
            }
        }
        if (dry) continue;
        if (canon_only && !(st & (TC_ST_CANONICAL | TC_ST_ALL_ANNOT))) continue;
        // ---- SAM line (printableSAM): empty fields are dropped, the line is stripped.  Written in place.
        const int32_t sl = R->seq_len[i];
        const int64_t c0 = R->cig_off[i], c1 = R->cig_off[i + 1], j0 = R->jn_off[i], j1 = R->jn_off[i + 1];
        sam.need((size_t)S->line[i].n + (size_t)sl + (size_t)(c1 - c0) * 12 + (size_t)R->md_len[i] + (size_t)(j1 - j0) * 40 + 128);
        const size_t row0 = sam.n;
        auto add = [&](const char* p, int64_t n) { if (n > 0) { if (sam.n > row0) sam.ch('\t'); sam.put(p, (size_t)n); } };
        auto add_span = [&](const Span& s) { add(T + s.s, s.n); };
        auto add_num = [&](int64_t v) { if (sam.n > row0) sam.ch('\t'); sam.num(v); };
        add_span(q); add_num(S->flag[i]); add_span(ch); add_num(S->pos[i]); add_span(S->f[4][i]);
        if (st & TC_ST_CIGAR_REWRITTEN) {
            if (c1 > c0) {
                if (sam.n > row0) sam.ch('\t');
                for (int64_t k = c0; k < c1; k++) { sam.num(R->cig_len[k]); sam.ch(R->cig_op[k] == 'd' ? 'D' : (char)R->cig_op[k]); }   // 'd': tc_plan.cuh, intron_text_rules
            }
        } else add_span(S->cigar[i]);
        add_span(S->f[6][i]); add_span(S->f[7][i]); add_span(S->f[8][i]);
        // SEQ: the input SEQ with the read's delta applied (include/tc_b200.h), written in place
        const uint8_t* sq = reinterpret_cast<const uint8_t*>(&sam.buf[sam.n]) + (sam.n > row0 && sl > 0 ? 1 : 0);
        if (sl > 0) {
            if (sam.n > row0) sam.ch('\t');
            const uint8_t* in = S->seq.data() + S->seq_off[i]; const int64_t in_len = S->seq_off[i + 1] - S->seq_off[i];
            uint8_t* d = reinterpret_cast<uint8_t*>(&sam.buf[sam.n]);
            if (R->delta_len[i] < -1 || apply_delta(in, in_len, R->delta + R->delta_off[i], R->delta_len[i], d, sl) != sl) { bad_delta = i; return; }
            sam.n += (size_t)sl;
        }
        add("*", 1);
        for (int64_t k = S->other_off[i]; k < S->other_off[i + 1]; k++) {
            // otherFields is one tab-joined field: dropped only when the whole join is empty
            if (k == S->other_off[i]) { bool any = false; for (int64_t m = k; m < S->other_off[i + 1]; m++) any |= S->other[m].n > 0 || m > k; if (!any) break; if (sam.n > row0) sam.ch('\t'); }
            else sam.ch('\t');
            sam.put(T + S->other[k].s, (size_t)S->other[k].n);
        }
        if (st & TC_ST_NMMD_DEVICE) {
            if (!(st & TC_ST_NMMD_NONE)) {
                if (sam.n > row0) sam.ch('\t');
                sam.put("NM:i:", 5); sam.num(R->nm[i]); sam.put("\tMD:Z:", 6);
                sam.put((const char*)R->md + R->md_off[i], (size_t)R->md_len[i]);
            }
        } else { add_span(S->tg[0][i]); add_span(S->tg[1][i]); }
        if (st & TC_ST_JM_DEVICE) {
            if (sam.n > row0) sam.ch('\t');
            sam.put("jM:B:c", 6);
            if (j1 > j0) for (int64_t k = j0; k < j1; k++) { sam.ch(','); sam.num(R->jm[k]); } else sam.put(",-1", 3);
        } else add_span(S->tg[2][i]);
        if (st & TC_ST_JI_DEVICE) {
            if (sam.n > row0) sam.ch('\t');
            sam.put("jI:B:i", 6);
            if (j1 > j0) for (int64_t k = 2 * j0; k < 2 * j1; k++) { sam.ch(','); sam.num(R->ji[k]); } else sam.put(",-1", 3);
        } else add_span(S->tg[3][i]);
        {   // strip
            size_t x = row0, y = sam.n;
            while (x < y && is_ws(sam.buf[x])) x++;
            while (y > x && is_ws(sam.buf[y - 1])) y--;
            if (x > row0) memmove(&sam.buf[row0], &sam.buf[x], y - x);
            sam.n = row0 + (y - x);
        }
        sam.ch('\n');
        // ---- FASTA record (printableFa): reverse complement on the minus strand, 80 columns
        fa.need((size_t)q.n + (size_t)sl + (size_t)sl / 80 + 8);
        fa.ch('>'); fa.put(T + q.s, (size_t)q.n); fa.ch('\n');
        if (S->flag[i] == 16) {
            for (int32_t k = 0; k < sl; k += 80) {
                if (k) fa.ch('\n');
                const int32_t m = std::min(80, sl - k); char* d = &fa.buf[fa.n];
                for (int32_t t = 0; t < m; t++) d[t] = (char)COMP.t[sq[sl - 1 - (k + t)]];
                fa.n += (size_t)m;
            }
        } else {
            for (int32_t k = 0; k < sl; k += 80) { if (k) fa.ch('\n'); fa.put((const char*)sq + k, (size_t)std::min(80, sl - k)); }
        }
        fa.ch('\n');
    }
}

struct SamBufs { std::vector<Part> parts; std::vector<Out> ps, pf, pl, pt; };
std::mutex g_pool_mu; std::vector<tc_sam*> g_pool;
tc_sam* sam_acquire() {
    std::lock_guard<std::mutex> g(g_pool_mu);
    if (g_pool.empty()) { tc_sam* S = new tc_sam(); S->bufs = new SamBufs(); return S; }
    tc_sam* S = g_pool.back(); g_pool.pop_back();
    return S;
}
void sam_release(tc_sam* S) {
    if (!S) return;
    S->text = nullptr; S->len = 0; S->n = 0; S->packed = false; S->n_alpha = 0;
    S->flag.clear(); S->chrom.clear(); S->pos.clear(); S->tags.clear(); S->cls.clear(); S->seq_off.clear(); S->cig_off.clear(); S->md_off.clear();
    S->ji_off.clear(); S->seq.clear(); S->cig_op.clear(); S->md.clear(); S->cig_len.clear(); S->ji.clear(); S->line.clear(); S->cigar.clear();
    for (auto& v : S->f) v.clear();
    for (auto& v : S->tg) v.clear();
    S->other_off.clear(); S->other.clear(); S->header.clear(); S->seqp.clear(); S->seq_poff.clear();
    for (auto& b : S->out) b.n = 0;
    std::lock_guard<std::mutex> g(g_pool_mu);
    if (g_pool.size() < 8) g_pool.push_back(S); else delete S;   // (one object per pipeline worker is in use at a time)
}

}  // namespace

tc_sam::~tc_sam() { delete bufs; }

extern "C" {

tc_sam* tc_sam_parse(const char* text, int64_t len, int32_t n_chrom, const char* const* chrom_names, const int64_t* chrom_len,
                     int32_t n_threads, char* err, int64_t err_cap) {
    if (!text || len < 0) { set_err(err, err_cap, "tc_sam_parse: no text"); return nullptr; }
    std::unordered_map<std::string, int> cmap;
    for (int k = 0; k < n_chrom; k++) cmap.emplace(chrom_names[k], k);
    int T = n_threads < 1 ? 1 : n_threads;
    if (len < (1 << 20)) T = 1;
    std::vector<int64_t> cut(T + 1, len);
    cut[0] = 0;
    for (int k = 1; k < T; k++) { int64_t p = len / T * k; while (p < len && text[p] != '\n') p++; cut[k] = p < len ? p + 1 : len; }
    for (int k = 1; k <= T; k++) if (cut[k] < cut[k - 1]) cut[k] = cut[k - 1];
    // Buffers are recycled (per calling thread / through a small pool of tc_sam objects): touching
    // fresh pages costs more than the parsing itself.
    tc_sam* S = nullptr;
  try {
    S = sam_acquire();
    std::vector<Part>& parts = S->bufs->parts;
    if (parts.size() < (size_t)T) parts.resize((size_t)T);
    for (auto& p : parts) p.reset();
    std::vector<std::thread> th;
    for (int k = 0; k < T; k++) th.emplace_back(parse_range, text, cut[k], cut[k + 1], std::cref(cmap), chrom_len, std::ref(parts[(size_t)k]));
    for (auto& t : th) t.join();
    for (int k = 0; k < T; k++) if (!parts[(size_t)k].err.empty()) { set_err(err, err_cap, parts[(size_t)k].err); sam_release(S); return nullptr; }
    S->text = text; S->len = len;
    // totals, then every part is copied to its place by its own thread
    std::vector<size_t> r0((size_t)T + 1, 0), s0((size_t)T + 1, 0), c0((size_t)T + 1, 0), m0((size_t)T + 1, 0), j0((size_t)T + 1, 0), o0((size_t)T + 1, 0);
    for (int k = 0; k < T; k++) {
        const Part& p = parts[(size_t)k];
        r0[k + 1] = r0[k] + p.flag.size(); s0[k + 1] = s0[k] + p.seq_bytes; c0[k + 1] = c0[k] + p.cig_op.size();
        m0[k + 1] = m0[k] + p.md.size(); j0[k + 1] = j0[k] + p.ji.size(); o0[k + 1] = o0[k] + p.other.size();
        for (auto& h : p.header) { S->header.append(text + h.s, (size_t)h.n); S->header.push_back('\n'); }
    }
    const size_t n = r0[T];
    S->flag.resize(n); S->chrom.resize(n); S->pos.resize(n); S->tags.resize(n); S->cls.resize(n);
    S->seq_off.resize(n + 1); S->cig_off.resize(n + 1); S->md_off.resize(n + 1); S->ji_off.resize(n + 1); S->other_off.resize(n);
    // (page-locking a new block is expensive: grow with headroom, so that blocks of about the same size never reallocate)
    auto fit = [](auto& v, size_t want) { if (v.capacity() < want) { v.clear(); v.reserve(want + want / 4 + ((size_t)1 << 16)); } v.resize(want); };
    fit(S->seq, s0[T] + 64); fit(S->cig_op, c0[T] + 8); fit(S->cig_len, c0[T] + 8); fit(S->md, m0[T] + 8); S->ji.resize(j0[T] + 8);
    S->line.resize(n); for (int k = 0; k < 9; k++) S->f[k].resize(n); S->cigar.resize(n); for (int k = 0; k < 4; k++) S->tg[k].resize(n);
    S->other.resize(o0[T]);
    auto place = [&](int k) {
        Part& p = parts[(size_t)k]; const size_t r = r0[k], m = p.flag.size();
        auto cp = [](auto& dst, size_t at, const auto& src) { if (!src.empty()) memcpy(dst.data() + at, src.data(), src.size() * sizeof(src[0])); };
        cp(S->flag, r, p.flag); cp(S->chrom, r, p.chrom); cp(S->pos, r, p.pos); cp(S->tags, r, p.tags); cp(S->cls, r, p.cls);
        cp(S->cig_op, c0[k], p.cig_op); cp(S->cig_len, c0[k], p.cig_len); cp(S->md, m0[k], p.md); cp(S->ji, j0[k], p.ji);
        cp(S->line, r, p.line); for (int q = 0; q < 9; q++) cp(S->f[q], r, p.f[q]); cp(S->cigar, r, p.cigar); for (int q = 0; q < 4; q++) cp(S->tg[q], r, p.tg[q]);
        cp(S->other, o0[k], p.other);
        int64_t so = (int64_t)s0[k], co = (int64_t)c0[k], mo = (int64_t)m0[k], jo = (int64_t)j0[k];
        size_t si = 0;                                                   // (seq_at holds the primary reads only)
        for (size_t i = 0; i < m; i++) {
            S->seq_off[r + i] = so; S->cig_off[r + i] = co; S->md_off[r + i] = mo; S->ji_off[r + i] = jo; S->other_off[r + i] = (int64_t)o0[k] + p.other_off[i];
            if (p.cls[i] == 0) { const Span sp = p.seq_at[si++]; if (sp.n > 0) memcpy(S->seq.data() + so, text + sp.s, (size_t)sp.n); }
            so += p.seq_len[i]; co += p.cig_n[i]; mo += p.md_n[i]; jo += p.ji_n[i];
        }
    };
    {
        std::vector<std::thread> tp;
        for (int k = 0; k < T; k++) tp.emplace_back(place, k);
        for (auto& t : tp) t.join();
    }
    S->seq_off[n] = (int64_t)s0[T]; S->cig_off[n] = (int64_t)c0[T]; S->md_off[n] = (int64_t)m0[T]; S->ji_off[n] = (int64_t)j0[T];
    S->other_off.push_back((int64_t)S->other.size());
    S->n = (int64_t)S->flag.size();
    return S;
  } catch (const std::exception& e) {                                  // allocation failure: an error return, not the end of the process
    set_err(err, err_cap, std::string("tc_sam_parse: ") + e.what());
    if (S) { try { sam_release(S); } catch (...) {} }
    return nullptr;
  }
}

// distinct RNAME values of the alignment lines, newline-joined (what split_SAM collects for
// validate_chroms, TC.py:535-536).  The caller releases the string with tc_free.
char* tc_sam_rnames(const char* text, int64_t len, int64_t* out_len) {
    std::unordered_map<std::string, int> seen; std::string out;
    int64_t p = 0;
    while (p < len) {
        const char* nl = (const char*)memchr(text + p, '\n', (size_t)(len - p));
        int64_t e = nl ? nl - text : len, s = p;
        p = e + 1;
        while (s < e && is_ws(text[s])) s++;
        if (s >= e || text[s] == '@') continue;
        int tabs = 0; int64_t a = -1, b = -1;
        for (int64_t k = s; k < e; k++) if (text[k] == '\t') { tabs++; if (tabs == 2) a = k + 1; else if (tabs == 3) { b = k; break; } }
        if (a < 0) continue;
        if (b < 0) { b = e; while (b > a && is_ws(text[b - 1])) b--; }
        if (seen.emplace(std::string(text + a, (size_t)(b - a)), 1).second) { out.append(text + a, (size_t)(b - a)); out.push_back('\n'); }
    }
    char* r = (char*)malloc(out.size() + 1);
    if (!r) { if (out_len) *out_len = 0; return nullptr; }           // (callers treat NULL as "out of memory")
    memcpy(r, out.data(), out.size()); r[out.size()] = 0;
    if (out_len) *out_len = (int64_t)out.size();
    return r;
}
// the '@' lines of a SAM text, trimmed and newline-terminated (what tc_sam_parse sets aside as the header);
// lets a caller that streams a file in blocks collect the header before any record is written
char* tc_sam_header_lines(const char* text, int64_t len, int64_t* out_len) {
    std::string out;
    int64_t p = 0;
    while (p < len) {
        const char* nl = (const char*)memchr(text + p, '\n', (size_t)(len - p));
        int64_t e = nl ? nl - text : len, s = p;
        p = e + 1;
        while (s < e && is_ws(text[s])) s++;
        while (e > s && is_ws(text[e - 1])) e--;
        if (s < e && text[s] == '@') { out.append(text + s, (size_t)(e - s)); out.push_back('\n'); }
    }
    char* r = (char*)malloc(out.size() + 1);
    if (!r) { if (out_len) *out_len = 0; return nullptr; }           // (callers treat NULL as "out of memory")
    memcpy(r, out.data(), out.size()); r[out.size()] = 0;
    if (out_len) *out_len = (int64_t)out.size();
    return r;
}
void tc_free(void* p) { free(p); }

void tc_sam_free(tc_sam* S) { sam_release(S); }
int64_t tc_sam_n_reads(const tc_sam* S) { return S ? S->n : 0; }
int64_t tc_sam_header(const tc_sam* S, const char** text) { if (!S) return 0; if (text) *text = S->header.data(); return (int64_t)S->header.size(); }

int64_t tc_seq_packed_size(const int64_t* seq_off, int64_t n) {
    int64_t t = 0;
    for (int64_t i = 0; i < n; i++) t += (seq_off[i + 1] - seq_off[i] + 7) / 8 * 4;
    return t;
}

int64_t tc_seq_pack(const uint8_t* seq, const int64_t* seq_off, int64_t n, const uint8_t* alphabet, int32_t n_alpha,
                    uint8_t* out, int64_t* poff, int32_t n_threads) {
    if (n_alpha < 1 || n_alpha > 16) return -1;
    uint8_t inv[256]; memset(inv, 0xff, sizeof(inv));
    for (int k = 0; k < n_alpha; k++) inv[alphabet[k]] = (uint8_t)k;
    poff[0] = 0;
    for (int64_t i = 0; i < n; i++) poff[i + 1] = poff[i] + (seq_off[i + 1] - seq_off[i] + 7) / 8 * 4;
    int T = n_threads < 1 ? 1 : n_threads; if (n < 4096) T = 1;
    std::vector<int> bad((size_t)T, 0);
    auto work = [&](int k) {
        for (int64_t i = n * k / T; i < n * (k + 1) / T; i++) {
            const uint8_t* r = seq + seq_off[i]; const int64_t L = seq_off[i + 1] - seq_off[i]; uint8_t* o = out + poff[i];
            unsigned any = 0;
            for (int64_t j = 0; j + 1 < L; j += 2) { const unsigned a = inv[r[j]], b = inv[r[j + 1]]; any |= a | b; o[j >> 1] = (uint8_t)(a | (b << 4)); }
            if (L & 1) { const unsigned a = inv[r[L - 1]]; any |= a; o[L >> 1] = (uint8_t)(a & 15); }
            for (int64_t j = (L + 1) / 2; j < poff[i + 1] - poff[i]; j++) o[j] = 0;
            if (any & 0xf0) bad[(size_t)k] = 1;
        }
    };
    std::vector<std::thread> th;
    for (int k = 0; k < T; k++) th.emplace_back(work, k);
    for (auto& t : th) t.join();
    for (int b : bad) if (b) return -1;
    return poff[n];
}

// packs the parsed SEQ column to 4 bits per base; 1 = done (tc_sam_batch now yields the packed form),
int64_t tc_seq_packed2_size(const int64_t* seq_off, int64_t n) {
    int64_t t = 0;
    for (int64_t i = 0; i < n; i++) t += (seq_off[i + 1] - seq_off[i] + 15) / 16 * 4;
    return t;
}
int64_t tc_seq_pack2(const uint8_t* seq, const int64_t* seq_off, int64_t n, uint8_t* out, int64_t* poff,
                     int64_t exc_cap, int64_t* exc_pos, uint8_t* exc_byte, int64_t* n_exc, int32_t n_threads) {
    uint8_t code[256]; memset(code, 0xff, sizeof(code));
    code[(uint8_t)'A'] = 0; code[(uint8_t)'C'] = 1; code[(uint8_t)'G'] = 2; code[(uint8_t)'T'] = 3;
    poff[0] = 0;
    for (int64_t i = 0; i < n; i++) poff[i + 1] = poff[i] + (seq_off[i + 1] - seq_off[i] + 15) / 16 * 4;
    int T = n_threads < 1 ? 1 : n_threads; if (n < 4096) T = 1;
    std::vector<std::vector<std::pair<int64_t, uint8_t>>> exc((size_t)T);
    auto work = [&](int k) {
        auto& ex = exc[(size_t)k];
        for (int64_t i = n * k / T; i < n * (k + 1) / T; i++) {
            const uint8_t* r = seq + seq_off[i]; const int64_t L = seq_off[i + 1] - seq_off[i]; uint8_t* o = out + poff[i];
            const int64_t nb = poff[i + 1] - poff[i];
            for (int64_t b = 0; b < nb; b++) {
                unsigned v = 0;
                for (int t = 0; t < 4; t++) {
                    const int64_t j = 4 * b + t;
                    if (j < L) { unsigned c = code[r[j]]; if (c > 3) { ex.emplace_back(seq_off[i] + j, r[j]); c = 0; } v |= c << (2 * t); }
                }
                o[b] = (uint8_t)v;
            }
        }
    };
    std::vector<std::thread> th;
    for (int k = 0; k < T; k++) th.emplace_back(work, k);
    for (auto& t : th) t.join();
    int64_t m = 0; for (auto& e : exc) m += (int64_t)e.size();
    *n_exc = m;
    if (m > exc_cap) return -1;
    m = 0;
    for (auto& e : exc) for (auto& x : e) { exc_pos[m] = x.first; exc_byte[m] = x.second; m++; }   // (thread ranges are ascending read ranges)
    return poff[n];
}
int64_t tc_cig_pack16(const uint8_t* cig_op, const int32_t* cig_len, int64_t n_ops, uint16_t* out, int64_t exc_cap,
                      int64_t* exc_idx, uint8_t* exc_op, int32_t* exc_len, int32_t n_threads) {
    uint8_t code[256]; memset(code, 0xff, sizeof(code));
    const char* letters = "MIDNSHP=XB";
    for (int k = 0; letters[k]; k++) code[(uint8_t)letters[k]] = (uint8_t)k;
    int T = n_threads < 1 ? 1 : n_threads; if (n_ops < (1 << 16)) T = 1;
    std::vector<std::vector<int64_t>> exc((size_t)T);
    auto work = [&](int t) {
        for (int64_t k = n_ops * t / T; k < n_ops * (t + 1) / T; k++) {
            const unsigned c = code[cig_op[k]]; const int32_t l = cig_len[k];
            if (c > 9 || l < 1 || l > 4095) { out[k] = 0xFFFFu; exc[(size_t)t].push_back(k); } else out[k] = (uint16_t)(c | ((unsigned)l << 4));
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < T; t++) th.emplace_back(work, t);
    for (auto& x : th) x.join();
    int64_t m = 0; for (auto& e : exc) m += (int64_t)e.size();
    if (m > exc_cap) return -1;
    m = 0;
    for (auto& e : exc) for (int64_t k : e) { exc_idx[m] = k; exc_op[m] = cig_op[k]; exc_len[m] = cig_len[k]; m++; }
    return m;
}
// 0 = a SEQ byte lies outside the alphabet (the batch stays ASCII)
int tc_sam_pack_seq(tc_sam* S, const uint8_t* alphabet, int32_t n_alpha, int32_t n_threads) {
    if (!S || !alphabet || n_alpha < 1 || n_alpha > 16) return 0;
    memset(S->alphabet, 0, 16); memcpy(S->alphabet, alphabet, (size_t)n_alpha); S->n_alpha = n_alpha;
    if (S->packed) return 1;
    S->seq_poff.assign((size_t)S->n + 1, 0);
    S->seqp.resize((size_t)tc_seq_packed_size(S->seq_off.data(), S->n) + 64);
    if (tc_seq_pack(S->seq.data(), S->seq_off.data(), S->n, alphabet, n_alpha, S->seqp.data(), S->seq_poff.data(), n_threads) < 0) {
        UV<uint8_t>().swap(S->seqp); return 0;
    }
    S->packed = true;
    return 1;
}

void tc_sam_batch(const tc_sam* S, tc_batch_in* b) {
    memset(b, 0, sizeof(*b));
    b->n_reads = S->n; b->flag = S->flag.data(); b->chrom = S->chrom.data(); b->pos = S->pos.data(); b->tags = S->tags.data();
    b->seq_off = S->seq_off.data(); b->seq = S->seq.data(); b->cig_off = S->cig_off.data(); b->cig_op = S->cig_op.data();
    b->cig_len = S->cig_len.data(); b->md_off = S->md_off.data(); b->md = S->md.data(); b->ji_off = S->ji_off.data(); b->ji = S->ji.data();
    if (S->packed) { b->seq = S->seqp.data(); b->seq_poff = S->seq_poff.data(); }
}

// the four texts as the parts of the T formatter threads (ps / pf / pl / pt, per calling thread, storage kept between batches)
// (a thread body: -2 in bad_delta = out of host memory)
static void render_range(const tc_sam* S, const tc_batch_out* R, int64_t a, int64_t b, bool primary_only, bool canon_only, bool dry,
                         Out& sam, Out& fa, Out& lg, Out& te, int64_t& bad_delta) {
    try { render_range_body(S, R, a, b, primary_only, canon_only, dry, sam, fa, lg, te, bad_delta); }
    catch (const std::exception&) { bad_delta = -2; }
}

static int render_parts(tc_sam* S, const tc_batch_out* R, int32_t primary_only, int32_t canon_only, int32_t dry, int32_t n_threads,
                        int& T_out, char* err, int64_t err_cap) {
    if (!S || !R || R->n_reads != S->n) { set_err(err, err_cap, "tc_sam_render: result does not belong to this batch"); return TC_E_ARG; }
    for (int64_t i = 0; i < S->n; i++) {
        const uint32_t st = R->status[i];
        if (st & TC_ST_ERR_MERGE) {                                    // dry run only: the reference's dryRun() dies on this read
            std::string q(S->text + S->f[0][i].s, (size_t)S->f[0][i].n);
            set_err(err, err_cap, "read " + q + ": MD tag and CIGAR disagree (the reference's --dryRun crashes on this read)");
            return TC_E_INPUT;
        }
    }
    int T = n_threads < 1 ? 1 : n_threads;
    if (S->n < 4096) T = 1;
    std::vector<Out>&ps = S->bufs->ps, &pf = S->bufs->pf, &pl = S->bufs->pl, &pt = S->bufs->pt;
    for (auto* v : {&ps, &pf, &pl, &pt}) { if (v->size() < (size_t)T) v->resize((size_t)T); for (auto& x : *v) x.clear(); }
    std::vector<std::thread> th; std::vector<int64_t> bad((size_t)T, -1);
    for (int k = 0; k < T; k++) {
        const int64_t a = S->n * k / T, b = S->n * (k + 1) / T;
        th.emplace_back(render_range, S, R, a, b, primary_only != 0, canon_only != 0, dry != 0, std::ref(ps[(size_t)k]), std::ref(pf[(size_t)k]),
                        std::ref(pl[(size_t)k]), std::ref(pt[(size_t)k]), std::ref(bad[(size_t)k]));
    }
    for (auto& t : th) t.join();
    for (int k = 0; k < T; k++) if (bad[(size_t)k] == -2) { set_err(err, err_cap, "tc_sam_render: out of host memory"); return TC_E_ARG; }
    for (int k = 0; k < T; k++) if (bad[(size_t)k] >= 0) {
        set_err(err, err_cap, "read " + std::string(S->text + S->f[0][bad[(size_t)k]].s, (size_t)S->f[0][bad[(size_t)k]].n) + ": the SEQ delta does not fit the input SEQ (result of another batch?)");
        return TC_E_ARG;
    }
    T_out = T;
    return TC_OK;
}

int tc_sam_render(tc_sam* S, const tc_batch_out* R, int32_t primary_only, int32_t canon_only, int32_t dry, int32_t n_threads,
                  const char** sam, int64_t* sam_len, const char** fa, int64_t* fa_len, const char** lg, int64_t* lg_len,
                  const char** te, int64_t* te_len, char* err, int64_t err_cap) {
    struct timespec ts0; clock_gettime(CLOCK_MONOTONIC, &ts0); const double t_enter = ts0.tv_sec + 1e-9 * ts0.tv_nsec;
    int T = 1;
    const int rc = render_parts(S, R, primary_only, canon_only, dry, n_threads, T, err, err_cap);
    if (rc != TC_OK) return rc;
    const bool tm = getenv("TC_HOST_TIMING") != nullptr;
    auto now = [] { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; };
    const double t_r = now();
    // the parts of the T formatter threads -> one text each, copied by T threads again
    std::vector<Out>* parts[4] = {&S->bufs->ps, &S->bufs->pf, &S->bufs->pl, &S->bufs->pt};
    std::vector<size_t> off((size_t)(4 * (T + 1)), 0);
    for (int q = 0; q < 4; q++) {
        for (int k = 0; k < T; k++) off[(size_t)(q * (T + 1) + k + 1)] = off[(size_t)(q * (T + 1) + k)] + (*parts[q])[(size_t)k].n;
        if (!S->out[q].ensure(off[(size_t)(q * (T + 1) + T)] + 1)) { set_err(err, err_cap, "tc_sam_render: out of host memory"); return TC_E_ARG; }
        S->out[q].n = off[(size_t)(q * (T + 1) + T)];
    }
    {
        std::vector<std::thread> tj;
        for (int k = 0; k < T; k++) tj.emplace_back([&, k] {
            for (int q = 0; q < 4; q++) { const Out& src = (*parts[q])[(size_t)k]; if (src.n) memcpy(S->out[q].p + off[(size_t)(q * (T + 1) + k)], src.buf.data(), src.n); }
        });
        for (auto& t : tj) t.join();
    }
    if (tm) fprintf(stderr, "tc_sam_render: formatting %.3f s (%d threads), join %.3f s\n", t_r - t_enter, T, now() - t_r);
    *sam = S->out[0].p; *sam_len = (int64_t)S->out[0].n; *fa = S->out[1].p; *fa_len = (int64_t)S->out[1].n;
    *lg = S->out[2].p; *lg_len = (int64_t)S->out[2].n; *te = S->out[3].p; *te_len = (int64_t)S->out[3].n;
    return TC_OK;
}

// tc_sam_render without the final copy: every text comes as n_parts pieces (the formatter threads' own buffers, in
// order), ptr[4 * k + q] / len[4 * k + q] for part k of text q (0 sam, 1 fa, 2 log, 3 te).  The pieces stay valid until
// the tc_sam is rendered again or freed.
int tc_sam_render_parts(tc_sam* S, const tc_batch_out* R, int32_t primary_only, int32_t canon_only, int32_t dry, int32_t n_threads,
                        int32_t* n_parts, const char** ptr, int64_t* len, int32_t max_parts, char* err, int64_t err_cap) {
    int T = 1;
    if (n_threads > max_parts) n_threads = max_parts;
    const int rc = render_parts(S, R, primary_only, canon_only, dry, n_threads, T, err, err_cap);
    if (rc != TC_OK) return rc;
    std::vector<Out>* parts[4] = {&S->bufs->ps, &S->bufs->pf, &S->bufs->pl, &S->bufs->pt};
    for (int k = 0; k < T; k++) for (int q = 0; q < 4; q++) { ptr[4 * k + q] = (*parts[q])[(size_t)k].buf.data(); len[4 * k + q] = (int64_t)(*parts[q])[(size_t)k].n; }
    *n_parts = T;
    return TC_OK;
}

}  // extern "C"

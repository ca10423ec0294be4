// tc_tables.cpp - host-side builders of the lookup tables the device searches (C++, multi-threaded).
//
// the reference builds these once per worker in Python -
//   processVCF                 TranscriptClean.py:769-856   SNP / insertion / deletion dicts
//   processSpliceAnnotation    TranscriptClean.py:666-766   donor / acceptor tables + annotated set
// - here one pass over the file's bytes yields the sorted key arrays of tc_ctx_load_variants /
// tc_ctx_load_junctions (include/tc_b200.h).  transcriptclean_b200/tables.py holds the same logic in Python
// (`build_*_py`); tests compare the two.  No GPU involved.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/tc_b200.h"

namespace tc_tables_detail {

void set_err(char* err, int64_t cap, const std::string& m) {
    if (!err || cap <= 0) return;
    size_t n = m.size() < (size_t)cap - 1 ? m.size() : (size_t)cap - 1;
    memcpy(err, m.data(), n); err[n] = 0;
}
inline bool ws(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }
// Python int() of a field: surrounding whitespace, optional sign, digits
bool field_int(const char* p, int64_t n, int64_t& out) {
    int64_t a = 0, b = n;
    while (a < b && tc_tables_detail::ws(p[a])) a++;
    while (b > a && tc_tables_detail::ws(p[b - 1])) b--;
    if (a >= b) return false;
    bool neg = false;
    if (p[a] == '+' || p[a] == '-') { neg = p[a] == '-'; a++; }
    if (a >= b) return false;
    int64_t v = 0;
    for (int64_t k = a; k < b; k++) { if (p[k] < '0' || p[k] > '9') return false; v = v * 10 + (p[k] - '0'); if (v > (1ll << 40)) return false; }
    out = neg ? -v : v;
    return true;
}
struct ChromMap {
    std::unordered_map<std::string, int> m; const char* last_p = nullptr; int64_t last_n = -1; int last_id = -1; std::string last;
    int find(const char* p, int64_t n) {                            // files are sorted by chromosome: one string compare per row
        if (n == last_n && last_n >= 0 && memcmp(p, last.data(), (size_t)n) == 0) return last_id;
        auto it = m.find(std::string(p, (size_t)n));
        last.assign(p, (size_t)n); last_n = n; last_id = it == m.end() ? -1 : it->second;
        return last_id;
    }
};
// lines of [a, b): calls f(line_start, line_len) for every line (without its '\n')
template <class F> void each_line(const char* t, int64_t a, int64_t b, F f) {
    while (a < b) {
        const char* nl = (const char*)memchr(t + a, '\n', (size_t)(b - a));
        const int64_t e = nl ? nl - t : b;
        f(t + a, e - a);
        a = e + 1;
    }
}
std::vector<int64_t> line_cuts(const char* t, int64_t len, int T) {
    std::vector<int64_t> cut(1, 0);
    for (int k = 1; k < T; k++) {
        int64_t p = len * k / T; if (p < cut.back()) p = cut.back();
        const char* nl = p < len ? (const char*)memchr(t + p, '\n', (size_t)(len - p)) : nullptr;
        cut.push_back(nl ? nl - t + 1 : len);
    }
    cut.push_back(len);
    return cut;
}

}  // namespace tc_tables_detail


struct tc_vcf {
    std::vector<uint64_t> snp_key; std::vector<uint8_t> snp_alt;
    std::vector<uint64_t> ins_key; std::vector<int32_t> ins_size; std::vector<int64_t> ins_aoff; std::vector<uint8_t> ins_bytes;
    std::vector<uint64_t> del_key; std::vector<int32_t> del_size;
    int64_t n_records = 0;
};
struct tc_sj {
    std::vector<uint64_t> don_s, acc_s, don_u, acc_u, ann_k1, ann_k2;
    int64_t n_don = 0, n_acc = 0;
};

extern "C" {

// processVCF (TC.py:814-854): CHROM POS . REF ALT[,ALT...]; per ALT: both one base -> SNP keyed (chrom, POS) with the ALT byte;
// else size = |len(REF) - len(ALT)| <= maxLenIndel: insertion keyed (chrom, POS) with size and ALT[1:], deletion keyed
// (chrom, POS) with size (Q11: the anchor base's position).  `use[c]` = 0 drops contig c (the reads touch no such contig;
// the reference narrows the catalogue to the reads with bedtools, TC.py:781-806 - a space saver that never changes a lookup).
tc_vcf* tc_vcf_parse(const char* text, int64_t len, int32_t n_chrom, const char* const* chrom_names, const uint8_t* use,
                     int32_t max_len_indel, int32_t n_threads, char* err, int64_t err_cap) {
    if (!text || len < 0 || n_chrom < 0) { tc_tables_detail::set_err(err, err_cap, "tc_vcf_parse: bad arguments"); return nullptr; }
    int T = n_threads < 1 ? 1 : (n_threads > 64 ? 64 : n_threads);
    if (len < (1 << 20)) T = 1;
    struct Ins { uint64_t key; int32_t size; int64_t off; int32_t n; };
    struct Part { std::vector<uint64_t> snp; std::vector<Ins> ins; std::vector<uint8_t> ibytes; std::vector<std::pair<uint64_t, int32_t>> del; int64_t nrec = 0; std::string err; };
    std::vector<Part> parts((size_t)T);
    const std::vector<int64_t> cut = tc_tables_detail::line_cuts(text, len, T);
    auto work = [&](int t) {
        Part& P = parts[(size_t)t]; tc_tables_detail::ChromMap cm;
        for (int c = 0; c < n_chrom; c++) cm.m.emplace(chrom_names[c], c);
        tc_tables_detail::each_line(text, cut[(size_t)t], cut[(size_t)t + 1], [&](const char* p, int64_t n) {
            if (!P.err.empty() || n == 0 || p[0] == '#') return;
            bool blank = true; for (int64_t k = 0; k < n && blank; k++) blank = tc_tables_detail::ws(p[k]);
            if (blank) return;
            const char* f[6]; int nf = 0; f[nf++] = p;                 // starts of the first five fields (+ end of the fifth)
            for (int64_t k = 0; k < n && nf < 6; k++) if (p[k] == '\t') f[nf++] = p + k + 1;
            if (nf < 5) { P.err = "VCF record with fewer than 5 columns: " + std::string(p, (size_t)(n < 60 ? n : 60)); return; }
            const char* end5 = nf == 6 ? f[5] - 1 : p + n;
            // (rstrip("\n") only: a '\r' stays part of the last field, as in the reference)
            P.nrec++;
            const int c = cm.find(f[0], f[1] - 1 - f[0]);
            int64_t pos;
            if (!tc_tables_detail::field_int(f[1], f[2] - 1 - f[1], pos)) { P.err = "VCF record with a non-integer POS: " + std::string(p, (size_t)(n < 60 ? n : 60)); return; }
            if (c < 0 || (use && !use[c])) return;
            const int64_t rl = f[4] - 1 - f[3];
            const uint64_t key = ((uint64_t)c << 32) | (uint32_t)pos;
            for (const char* a = f[4]; a <= end5;) {
                const char* e = (const char*)memchr(a, ',', (size_t)(end5 - a)); if (!e) e = end5;
                const int64_t al = e - a;
                if (rl == 1 && al == 1) P.snp.push_back((key << 8) | (uint8_t)a[0]);
                else {
                    const int64_t size = rl > al ? rl - al : al - rl;
                    if (size <= max_len_indel) {
                        if (rl < al) { P.ins.push_back({key, (int32_t)size, (int64_t)P.ibytes.size(), (int32_t)(al - 1)}); P.ibytes.insert(P.ibytes.end(), a + 1, e); }
                        else if (rl > al) P.del.emplace_back(key, (int32_t)size);
                    }
                }
                a = e + 1;
            }
        });
    };
    if (T == 1) work(0);
    else { std::vector<std::thread> th; for (int t = 0; t < T; t++) th.emplace_back(work, t); for (auto& x : th) x.join(); }
    for (auto& P : parts) if (!P.err.empty()) { tc_tables_detail::set_err(err, err_cap, P.err); return nullptr; }
    tc_vcf* V = new tc_vcf();
    // SNPs: one sort over (key << 8 | ALT); rows of one key stay together, which is all the lookup needs (TC.py:1114-1118)
    size_t ns = 0, ni = 0, nd = 0; for (auto& P : parts) { ns += P.snp.size(); ni += P.ins.size(); nd += P.del.size(); V->n_records += P.nrec; }
    std::vector<uint64_t> all; all.reserve(ns);
    for (auto& P : parts) { all.insert(all.end(), P.snp.begin(), P.snp.end()); std::vector<uint64_t>().swap(P.snp); }
    std::sort(all.begin(), all.end());
    V->snp_key.resize(ns); V->snp_alt.resize(ns);
    for (size_t k = 0; k < ns; k++) { V->snp_key[k] = all[k] >> 8; V->snp_alt[k] = (uint8_t)all[k]; }
    std::vector<uint64_t>().swap(all);
    // insertions by (key, size), input order inside (stable)
    struct Ref { uint64_t key; int32_t size; int32_t part; int64_t off; int32_t n; };
    std::vector<Ref> ir; ir.reserve(ni);
    for (int t = 0; t < T; t++) for (auto& x : parts[(size_t)t].ins) ir.push_back({x.key, x.size, t, x.off, x.n});
    std::stable_sort(ir.begin(), ir.end(), [](const Ref& a, const Ref& b) { return a.key < b.key || (a.key == b.key && a.size < b.size); });
    V->ins_key.resize(ni); V->ins_size.resize(ni); V->ins_aoff.assign(ni + 1, 0);
    for (size_t k = 0; k < ni; k++) { V->ins_key[k] = ir[k].key; V->ins_size[k] = ir[k].size; V->ins_aoff[k + 1] = V->ins_aoff[k] + ir[k].n; }
    V->ins_bytes.resize((size_t)V->ins_aoff[ni]);
    for (size_t k = 0; k < ni; k++) if (ir[k].n) memcpy(V->ins_bytes.data() + V->ins_aoff[k], parts[(size_t)ir[k].part].ibytes.data() + ir[k].off, (size_t)ir[k].n);
    // deletions: a set of (key, size)
    std::vector<std::pair<uint64_t, int32_t>> dl; dl.reserve(nd);
    for (auto& P : parts) dl.insert(dl.end(), P.del.begin(), P.del.end());
    std::sort(dl.begin(), dl.end()); dl.erase(std::unique(dl.begin(), dl.end()), dl.end());
    V->del_key.resize(dl.size()); V->del_size.resize(dl.size());
    for (size_t k = 0; k < dl.size(); k++) { V->del_key[k] = dl[k].first; V->del_size[k] = dl[k].second; }
    return V;
}
void tc_vcf_free(tc_vcf* v) { delete v; }
int64_t tc_vcf_n_records(const tc_vcf* v) { return v ? v->n_records : 0; }
void tc_vcf_tables(const tc_vcf* v, int64_t* n_snp, const uint64_t** snp_key, const uint8_t** snp_alt, int64_t* n_ins,
                   const uint64_t** ins_key, const int32_t** ins_size, const int64_t** ins_aoff, const uint8_t** ins_bytes,
                   int64_t* n_del, const uint64_t** del_key, const int32_t** del_size) {
    *n_snp = (int64_t)v->snp_key.size(); *snp_key = v->snp_key.data(); *snp_alt = v->snp_alt.data();
    *n_ins = (int64_t)v->ins_key.size(); *ins_key = v->ins_key.data(); *ins_size = v->ins_size.data(); *ins_aoff = v->ins_aoff.data(); *ins_bytes = v->ins_bytes.data();
    *n_del = (int64_t)v->del_key.size(); *del_key = v->del_key.data(); *del_size = v->del_size.data();
}

// processSpliceAnnotation (TC.py:683-716): chrom, intron start, intron end (1-based inclusive), strand column {0,1,2};
// rows on contigs without reads (`use[c]` = 0 or a name the genome does not hold... the reference tests the reads' set first,
// TC.py:685) are dropped before anything else is parsed; strand 0 rows are skipped (TC.py:714); '+' : donor = start,
// acceptor = end, '-' the other way round (TC.py:704-713); positions are kept 0-based (the BED starts pyranges sees).
tc_sj* tc_sj_parse(const char* text, int64_t len, int32_t n_chrom, const char* const* chrom_names, const uint8_t* use,
                   char* err, int64_t err_cap) {
    if (!text || len < 0) { tc_tables_detail::set_err(err, err_cap, "tc_sj_parse: bad arguments"); return nullptr; }
    tc_tables_detail::ChromMap cm; for (int c = 0; c < n_chrom; c++) cm.m.emplace(chrom_names[c], c);
    std::vector<uint64_t> don, acc; std::vector<std::pair<uint64_t, uint64_t>> ann; std::string e;
    tc_tables_detail::each_line(text, 0, len, [&](const char* p, int64_t n) {
        if (!e.empty()) return;
        int64_t a = 0, b = n;                                         // line.strip()
        while (a < b && tc_tables_detail::ws(p[a])) a++;
        while (b > a && tc_tables_detail::ws(p[b - 1])) b--;
        if (a >= b) return;                                            // ''.split('\t')[0] = '' is in no chromosome set
        const char* f[10]; int nf = 0; f[nf++] = p + a;
        for (int64_t k = a; k < b && nf < 10; k++) if (p[k] == '\t') f[nf++] = p + k + 1;
        const char* endp = p + b;
        auto fend = [&](int k) { return k + 1 < nf ? f[k + 1] - 1 : endp; };
        const int c = cm.find(f[0], fend(0) - f[0]);
        if (c < 0 || (use && !use[c])) return;                        // chrom not among the reads' chromosomes
        if (nf < 9) { e = "SJ row with fewer than 9 columns: " + std::string(p, (size_t)(n < 60 ? n : 60)); return; }
        int64_t s, en, x;
        if (!tc_tables_detail::field_int(f[1], fend(1) - f[1], s) || !tc_tables_detail::field_int(f[2], fend(2) - f[2], en) || !tc_tables_detail::field_int(f[4], fend(4) - f[4], x) ||
            !tc_tables_detail::field_int(f[5], fend(5) - f[5], x)) { e = "SJ row with a non-integer column: " + std::string(p, (size_t)(n < 60 ? n : 60)); return; }
        const int64_t sl = fend(3) - f[3];
        const int strand = (sl == 1 && f[3][0] == '1') ? 0 : (sl == 1 && f[3][0] == '2') ? 1 : -1;
        if (strand < 0) return;
        const uint64_t ks = ((uint64_t)c << 33) | ((uint64_t)strand << 32) | (uint32_t)(s - 1), ke = ((uint64_t)c << 33) | ((uint64_t)strand << 32) | (uint32_t)(en - 1);
        if (strand == 0) { don.push_back(ks); acc.push_back(ke); } else { acc.push_back(ks); don.push_back(ke); }
        ann.emplace_back(((uint64_t)c << 32) | (uint32_t)s, ((uint64_t)(uint32_t)en << 1) | (uint64_t)strand);
    });
    if (!e.empty()) { tc_tables_detail::set_err(err, err_cap, e); return nullptr; }
    tc_sj* S = new tc_sj();
    auto uniq = [](std::vector<uint64_t>& v) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); };
    uniq(don); uniq(acc);
    S->don_s = don; S->acc_s = acc; S->n_don = (int64_t)don.size(); S->n_acc = (int64_t)acc.size();
    auto unstrand = [&](const std::vector<uint64_t>& v, std::vector<uint64_t>& o) {
        o.resize(v.size());
        for (size_t k = 0; k < v.size(); k++) o[k] = ((v[k] >> 33) << 32) | (v[k] & 0xffffffffull);
        uniq(o);
    };
    unstrand(don, S->don_u); unstrand(acc, S->acc_u);
    std::sort(ann.begin(), ann.end()); ann.erase(std::unique(ann.begin(), ann.end()), ann.end());
    S->ann_k1.resize(ann.size()); S->ann_k2.resize(ann.size());
    for (size_t k = 0; k < ann.size(); k++) { S->ann_k1[k] = ann[k].first; S->ann_k2[k] = ann[k].second; }
    return S;
}
void tc_sj_free(tc_sj* s) { delete s; }
void tc_sj_tables(const tc_sj* s, int64_t* n_don_s, const uint64_t** don_s, int64_t* n_acc_s, const uint64_t** acc_s,
                  int64_t* n_don_u, const uint64_t** don_u, int64_t* n_acc_u, const uint64_t** acc_u,
                  int64_t* n_annot, const uint64_t** ann_k1, const uint64_t** ann_k2) {
    *n_don_s = (int64_t)s->don_s.size(); *don_s = s->don_s.data(); *n_acc_s = (int64_t)s->acc_s.size(); *acc_s = s->acc_s.data();
    *n_don_u = (int64_t)s->don_u.size(); *don_u = s->don_u.data(); *n_acc_u = (int64_t)s->acc_u.size(); *acc_u = s->acc_u.data();
    *n_annot = (int64_t)s->ann_k1.size(); *ann_k1 = s->ann_k1.data(); *ann_k2 = s->ann_k2.data();
}

// ---------------------------------------------------------------------------------- FASTA -> genome image
// Replaces pyfaidx.Fasta's indexing of the reference FASTA (TranscriptClean.py:221) for the device: the file is mapped, cut into
// records ('>' of the first header, then every "\n>") and packed by all host threads into the planes of tc_ctx_load_genome
// (2-bit bases, lower-case bits, literal runs for every byte outside ACGTacgt).  GenomeImage.from_fasta (genome.py) is the same
// logic in numpy; tests compare the two bit by bit.
}  // extern "C"  (reopened below)

#include <atomic>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

struct tc_fasta {
    const char* data = nullptr; int64_t len = 0; bool mapped = false; std::string held;
    struct Rec { int64_t name_s; int32_t name_n; int64_t body_s, body_e, length; };
    std::vector<Rec> recs;
    struct Task { int32_t rec; int64_t a, b, base0; };                  // bytes [a, b) of a record's body; base0 = bases of the record before a
    std::vector<Task> tasks;
    std::vector<int64_t> exc_start, exc_end; std::vector<uint8_t> exc_byte;
    ~tc_fasta() { if (mapped && data) munmap((void*)data, (size_t)len); }
};

namespace tc_tables_detail {
template <class F> void run_tasks(size_t n, int T, F f) {
    std::atomic<size_t> next(0);
    auto work = [&]() { for (size_t k = next.fetch_add(1); k < n; k = next.fetch_add(1)) f(k); };
    std::vector<std::thread> th;
    for (int t = 1; t < T; t++) th.emplace_back(work);
    work();
    for (auto& x : th) x.join();
}
}  // namespace tc_tables_detail

extern "C" {

// distinct first columns of the record lines of a VCF text (what validate_chroms collects, TranscriptClean.py:497-505: every line
// that is not blank and does not start with '#', cut at its first tab), each followed by a NUL, in a malloc'ed buffer (tc_free)
char* tc_vcf_chroms(const char* text, int64_t len, int64_t* out_len) {
    std::vector<std::string> seen; std::string last; bool have_last = false;
    tc_tables_detail::each_line(text, 0, len, [&](const char* p, int64_t n) {
        if (n > 0 && p[0] == '#') return;
        int64_t a = 0; while (a < n && tc_tables_detail::ws(p[a])) a++;
        if (a == n) return;                                             // blank
        const char* tp = (const char*)memchr(p, '\t', (size_t)n);
        const int64_t k = tp ? tp - p : (p + n < text + len ? n + 1 : n);   // (no tab: Python's split keeps the line end in the name)
        if (have_last && (int64_t)last.size() == k && memcmp(last.data(), p, (size_t)k) == 0) return;
        last.assign(p, (size_t)k); have_last = true;
        if (std::find(seen.begin(), seen.end(), last) == seen.end()) seen.push_back(last);
    });
    size_t total = 0; for (auto& x : seen) total += x.size() + 1;
    char* out = (char*)malloc(total + 1);
    if (!out) return nullptr;
    size_t o = 0; for (auto& x : seen) { memcpy(out + o, x.data(), x.size()); o += x.size(); out[o++] = 0; }
    out[o] = 0; *out_len = (int64_t)o;
    return out;
}

tc_fasta* tc_fasta_open(const char* path, int32_t n_threads, char* err, int64_t err_cap) {
    const int T = n_threads < 1 ? 1 : n_threads;
    tc_fasta* F = new tc_fasta();
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { tc_tables_detail::set_err(err, err_cap, std::string("cannot open ") + path); delete F; return nullptr; }
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); tc_tables_detail::set_err(err, err_cap, std::string("cannot stat ") + path); delete F; return nullptr; }
    F->len = (int64_t)st.st_size;
    if (F->len > 0) {
        void* m = mmap(nullptr, (size_t)F->len, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) { close(fd); tc_tables_detail::set_err(err, err_cap, std::string("cannot map ") + path); delete F; return nullptr; }
        F->data = (const char*)m; F->mapped = true;
    }
    close(fd);
    const char* D = F->data; const int64_t n = F->len;
    // records: '>' of the first header, then every "\n>" behind a header line
    int64_t pos = 0;
    while (pos < n) {
        const char* hp = (const char*)memchr(D + pos, '>', (size_t)(n - pos));
        if (!hp) break;
        const int64_t h = hp - D;
        const char* nlp = (const char*)memchr(D + h, '\n', (size_t)(n - h));
        const int64_t nl = nlp ? nlp - D : n;
        int64_t end = n;
        if (nl < n) { const char* nx = (const char*)memmem(D + nl, (size_t)(n - nl), "\n>", 2); if (nx) end = (nx - D) + 1; }
        int64_t a = h + 1; while (a < nl && tc_tables_detail::ws(D[a])) a++;
        int64_t b = a; while (b < nl && !tc_tables_detail::ws(D[b])) b++;
        if (b == a) { tc_tables_detail::set_err(err, err_cap, "FASTA header without a name"); delete F; return nullptr; }
        tc_fasta::Rec r; r.name_s = a; r.name_n = (int32_t)(b - a); r.body_s = nl < n ? nl + 1 : n; r.body_e = end; r.length = 0;
        if (r.body_e < r.body_s) r.body_e = r.body_s;
        F->recs.push_back(r);
        pos = end;
    }
    // bases per record: every body byte that is not a line end; counted by all threads over pieces of 4 MB
    const int64_t piece = (int64_t)4 << 20;
    for (size_t k = 0; k < F->recs.size(); k++)
        for (int64_t a = F->recs[k].body_s; a < F->recs[k].body_e; a += piece)
            F->tasks.push_back({(int32_t)k, a, std::min(a + piece, F->recs[k].body_e), 0});
    std::vector<int64_t> cnt(F->tasks.size(), 0);
    tc_tables_detail::run_tasks(F->tasks.size(), T, [&](size_t k) {
        const tc_fasta::Task& t = F->tasks[k]; int64_t le = 0;
        for (int64_t j = t.a; j < t.b; j++) le += (D[j] == '\n') | (D[j] == '\r');
        cnt[k] = (t.b - t.a) - le;
    });
    for (size_t k = 0; k < F->tasks.size(); k++) {
        tc_fasta::Rec& r = F->recs[(size_t)F->tasks[k].rec];
        F->tasks[k].base0 = r.length; r.length += cnt[k];
    }
    return F;
}
int32_t tc_fasta_n_contigs(const tc_fasta* f) { return f ? (int32_t)f->recs.size() : 0; }
int tc_fasta_contig(const tc_fasta* f, int32_t k, const char** name, int32_t* name_len, int64_t* length) {
    if (!f || k < 0 || (size_t)k >= f->recs.size()) return TC_E_ARG;
    *name = f->data + f->recs[(size_t)k].name_s; *name_len = f->recs[(size_t)k].name_n; *length = f->recs[(size_t)k].length;
    return TC_OK;
}
// planes must be zero; chrom_off[k] = plane index of contig k's first base.  Returns the number of literal runs (kept inside; tc_fasta_exceptions).
int64_t tc_fasta_pack(tc_fasta* F, const int64_t* chrom_off, int64_t n_pad, uint32_t* bases2, uint32_t* lower1, int32_t n_threads) {
    if (!F || !chrom_off || !bases2 || !lower1) return -1;
    for (size_t k = 0; k < F->recs.size(); k++) if (chrom_off[k] < 0 || chrom_off[k] + F->recs[k].length > n_pad) return -1;
    uint8_t code[256], kind[256];                                       // kind: 0 ACGT, 1 acgt, 2 line end, 3 literal
    for (int c = 0; c < 256; c++) { code[c] = 0; kind[c] = 3; }
    const char* up = "ACGT"; const char* lo = "acgt";
    for (int k = 0; k < 4; k++) { code[(uint8_t)up[k]] = (uint8_t)k; kind[(uint8_t)up[k]] = 0; code[(uint8_t)lo[k]] = (uint8_t)k; kind[(uint8_t)lo[k]] = 1; }
    kind[10] = kind[13] = 2;
    struct Run { int64_t s, e; uint8_t b; };
    std::vector<std::vector<Run>> runs(F->tasks.size());
    const char* D = F->data;
    tc_tables_detail::run_tasks(F->tasks.size(), n_threads < 1 ? 1 : n_threads, [&](size_t ti) {
        const tc_fasta::Task& t = F->tasks[ti];
        int64_t p = chrom_off[t.rec] + t.base0;                            // plane index of the next base
        const int64_t p_first = p;
        uint32_t w2 = 0, wl = 0; std::vector<Run>& R = runs[ti];
        // a word goes out when the plane index leaves it - OR-ed in, because the first and the last word of a piece are shared with
        // the neighbouring pieces (the planes start as zeros)
        auto put2 = [&](int64_t word, uint32_t v) { if (v) __atomic_fetch_or(&bases2[word], v, __ATOMIC_RELAXED); };
        auto putl = [&](int64_t word, uint32_t v) { if (v) __atomic_fetch_or(&lower1[word], v, __ATOMIC_RELAXED); };
        for (int64_t j = t.a; j < t.b; j++) {
            const uint8_t c = (uint8_t)D[j]; const uint8_t k = kind[c];
            if (k == 2) continue;
            w2 |= (uint32_t)code[c] << ((p & 15) * 2);
            if (k == 1) wl |= 1u << (p & 31);
            else if (k == 3) { if (!R.empty() && R.back().e == p && R.back().b == c) R.back().e = p + 1; else R.push_back({p, p + 1, c}); }
            p++;
            if ((p & 15) == 0) { put2((p - 1) >> 4, w2); w2 = 0; if ((p & 31) == 0) { putl((p - 1) >> 5, wl); wl = 0; } }
        }
        if (p > p_first) { if (p & 15) put2(p >> 4, w2); if (p & 31) putl(p >> 5, wl); }
    });
    F->exc_start.clear(); F->exc_end.clear(); F->exc_byte.clear();
    for (auto& R : runs) for (const Run& r : R) {                       // (pieces are in plane order: runs that meet at a piece boundary are joined)
        if (!F->exc_start.empty() && F->exc_end.back() == r.s && F->exc_byte.back() == r.b) F->exc_end.back() = r.e;
        else { F->exc_start.push_back(r.s); F->exc_end.push_back(r.e); F->exc_byte.push_back(r.b); }
    }
    return (int64_t)F->exc_start.size();
}
int64_t tc_fasta_exceptions(const tc_fasta* f, int64_t* start, int64_t* end, uint8_t* byte) {
    if (!f) return 0;
    const size_t n = f->exc_start.size();
    if (start && end && byte && n) { memcpy(start, f->exc_start.data(), n * 8); memcpy(end, f->exc_end.data(), n * 8); memcpy(byte, f->exc_byte.data(), n); }
    return (int64_t)n;
}
void tc_fasta_free(tc_fasta* f) { delete f; }

}  // extern "C"

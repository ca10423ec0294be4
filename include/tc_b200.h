/* tc_b200.h - C ABI of the B200 correction engine (libtc_b200.so).
 *
 * The reference (mortazavilab/TranscriptClean v2.1) has no FFI: its seam is the Python call
 *   batch_correct(sam_transcripts, options, refs, outfiles, buffer_size)   TranscriptClean.py:350-403
 * which loops correct_transcript(line, options, refs)                       TranscriptClean.py:406-460.
 * This library replaces everything between "N parsed SAM records in" and "corrected
 * records / log counters / TE records out" of that call.  Plain pointers and sizes only.
 *
 * Conventions: every function returns 0 on success, a negative TC_E_* code otherwise;
 * tc_last_error() gives the text.  One context per GPU, one host thread per context.
 * All `const T*` inputs are HOST pointers unless the function name ends in `_device`.
 * Output buffers are owned by the context (pinned host memory) and stay valid until the next
 * tc_correct_batch / tc_dryrun_batch / tc_ctx_destroy on that context.
 */
#ifndef TC_B200_H
#define TC_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tc_ctx tc_ctx;

enum {
    TC_OK = 0,
    TC_E_CUDA = -1,       /* CUDA runtime error (no device, OOM, launch failure) */
    TC_E_ARG = -2,        /* bad argument */
    TC_E_STATE = -3,      /* call order (e.g. batch before genome) */
    TC_E_INPUT = -4       /* a read the engine flags as unsupported (see TC_ST_ERR_*) */
};

/* options.maxLenIndel / maxSJOffset / correct* of the reference (TranscriptClean.py:100-126),
 * plus the two pyranges.nearest() semantics the reference's tests leave unpinned. */
typedef struct {
    int32_t max_len_indel;       /* --maxLenIndel, default 5 */
    int32_t max_sj_offset;       /* --maxSJOffset, default 5 */
    int32_t correct_mismatches;  /* --correctMismatches true/false */
    int32_t correct_indels;      /* --correctIndels */
    int32_t correct_sjs;         /* --correctSJs (only effective when junctions are loaded) */
    int32_t sj_ignore_strand;    /* 0: nearest site on the read's strand (default); 1: any strand */
    int32_t sj_tie_right;        /* 0: equidistant -> lower coordinate (default); 1: higher */
    int32_t reserved;
} tc_options;

/* per-read status word */
enum {
    TC_ST_CLASS_MASK = 3u,        /* 0 primary, 1 unmapped, 2 non-primary   (TC.py:1570-1575) */
    TC_ST_FALLBACK = 1u << 2,     /* exception path: original read is output (TC.py:450-456) */
    TC_ST_CIGAR_REWRITTEN = 1u << 3, /* a corrector re-serialised the CIGAR; else verbatim input */
    TC_ST_NMMD_DEVICE = 1u << 4,  /* NM/MD come from the engine (nm[], md[]); else input tags */
    TC_ST_NMMD_NONE = 1u << 5,    /* NM/MD are None (alignment ran past the contig end) */
    TC_ST_JM_DEVICE = 1u << 6,    /* jM must be rendered from jm[]; else input tag */
    TC_ST_JI_DEVICE = 1u << 7,    /* jI must be rendered from ji[]; else input tag */
    TC_ST_CANONICAL = 1u << 8,    /* transcript.isCanonical */
    TC_ST_ALL_ANNOT = 1u << 9,    /* transcript.allJnsAnnotated */
    TC_ST_ERR_INTRON = 1u << 10,  /* informational: a junction rescue met the one text-CIGAR corner that is not modelled (a negative
                                     intron behind an op other than M/S/I/D); that junction was left uncorrected ("Other") */
    TC_ST_ERR_MERGE = 1u << 11    /* dry run only: MD/CIGAR merge ran off a list (reference crashes) */
};

/* indices into counters[10*i + k]; -1 encodes the reference's "NA" (TC.py:1577-1586) */
enum { TC_C_COR_DEL, TC_C_UNC_DEL, TC_C_VAR_DEL, TC_C_COR_INS, TC_C_UNC_INS, TC_C_VAR_INS,
       TC_C_COR_MM, TC_C_UNC_MM, TC_C_COR_SJ, TC_C_UNC_SJ };

/* one row of the .TE.log (TC.py:911-936, 1017-1046, 1119-1134, 1227-1228, 1650-1668), decoded */
typedef struct {
    int32_t a;        /* Mismatch: position; Insertion/Deletion: g-1; NC_SJ: intron start */
    int32_t b;        /* Insertion/Deletion: g+size-1; NC_SJ: intron end; Mismatch: unused */
    int32_t size;
    uint8_t type;     /* 0 Mismatch 1 Insertion 2 Deletion 3 NC_SJ_boundary */
    uint8_t status;   /* 0 Corrected 1 Uncorrected */
    uint8_t reason;   /* 0 NA 1 VariantMatch 2 TooLarge 3 TooFarFromAnnotJn 4 Other 5 DryRun */
    uint8_t pad;
} tc_te;

/* The rows cross the ABI as 8-byte words (tc_batch_out.te): bits 0-31 a, 32-55 size, 56-57 type, 58 status, 59-61 reason.
 * An NC_SJ row is followed by one more word whose low 32 bits hold b; for an Insertion / Deletion b = a + size.
 * Decodes the row at w[0] and returns the number of words it occupies (1 or 2). */
static inline int tc_te_decode(const uint64_t* w, tc_te* row) {
    row->a = (int32_t)(uint32_t)w[0]; row->size = (int32_t)((w[0] >> 32) & 0xffffffu);
    row->type = (uint8_t)((w[0] >> 56) & 3u); row->status = (uint8_t)((w[0] >> 58) & 1u); row->reason = (uint8_t)((w[0] >> 59) & 7u); row->pad = 0;
    if (row->type == 3) { row->b = (int32_t)(uint32_t)w[1]; return 2; }
    row->b = row->type == 0 ? 0 : row->a + row->size;
    return 1;
}

/* SEQ delta: the corrected SEQ of a read as an edit script against its INPUT SEQ (the caller holds it), tc_batch_out.delta.
 * A stream of ops, each a header byte kind << 6 | n6 (n6 < 62: n = n6; 62: n = the next 2 bytes; 63: n = the next 4 bytes,
 * little endian): kind 0 COPY n input bases, 1 SKIP n input bases, 2 LIT: the n bytes that follow (genome bases, case and
 * N / IUPAC kept).  delta_len < 0: the SEQ is the input SEQ.  Writes the new SEQ to `out` (capacity out_cap) and returns
 * its length, or -1 when the script runs past the input, the output or its own end. */
static inline int64_t tc_seq_apply_delta(const uint8_t* in_seq, int64_t in_len, const uint8_t* delta, int64_t delta_len,
                                         uint8_t* out, int64_t out_cap) {
    int64_t q = 0, o = 0, k = 0, j;
    if (delta_len < 0) { if (in_len > out_cap) return -1; for (j = 0; j < in_len; j++) out[j] = in_seq[j]; return in_len; }
    while (k < delta_len) {
        const int kind = delta[k] >> 6; int64_t n = delta[k] & 63; k++;
        if (n == 62) { if (k + 2 > delta_len) return -1; n = (int64_t)delta[k] | ((int64_t)delta[k + 1] << 8); k += 2; }
        else if (n == 63) { if (k + 4 > delta_len) return -1; n = (int64_t)delta[k] | ((int64_t)delta[k + 1] << 8) | ((int64_t)delta[k + 2] << 16) | ((int64_t)delta[k + 3] << 24); k += 4; }
        if (kind == 1) { q += n; continue; }
        if (o + n > out_cap) return -1;
        if (kind == 0) { if (q + n > in_len) return -1; for (j = 0; j < n; j++) out[o + j] = in_seq[q + j]; q += n; }
        else if (kind == 2) { if (k + n > delta_len) return -1; for (j = 0; j < n; j++) out[o + j] = delta[k + j]; k += n; }
        else return -1;
        o += n;
    }
    return o;
}

/* Columnar batch of SAM records (what Transcript.__init__ parses, transcript.py:10-72). */
typedef struct {
    int64_t n_reads;
    const int32_t* flag;      /* FLAG */
    const int32_t* chrom;     /* index of RNAME in the loaded genome, -1 for '*' */
    const int32_t* pos;       /* POS, 1-based */
    const uint8_t* tags;      /* bit0 NM present, bit1 MD present, bit2 jM present, bit3 jI present */
    const int64_t* seq_off;   /* n+1 */
    const uint8_t* seq;       /* SEQ bytes (ASCII) */
    const int64_t* cig_off;   /* n+1 */
    const uint8_t* cig_op;    /* CIGAR op letters */
    const int32_t* cig_len;   /* CIGAR op lengths */
    const int64_t* md_off;    /* n+1 */
    const uint8_t* md;        /* value of the MD tag (third ':' piece), when present */
    const int64_t* ji_off;    /* n+1 */
    const int32_t* ji;        /* integers of the input jI tag, when present */
    /* 4-bit SEQ (optional): when seq_poff != NULL, `seq` holds two bases per byte (low nibble first),
     * codes index the alphabet of tc_ctx_set_alphabet, row i starts at byte seq_poff[i] (a multiple
     * of 4) and has seq_off[i+1]-seq_off[i] bases; seq_off keeps the dense base offsets.  (The results
     * never carry SEQ rows: the new SEQ comes back as a delta against the input SEQ, tc_batch_out.delta.) */
    const int64_t* seq_poff;  /* n+1, or NULL: SEQ is ASCII */
    /* 2-bit SEQ (optional, seq_bits == 2; seq_bits 0 / 4 / 8 = the forms above): `seq` holds four bases per byte (base j of a row in
     * bits 2*(j%4) of its byte j/4; A=0 C=1 G=2 T=3), rows at seq_poff[i] (multiples of 4) as for the 4-bit form.  Every base that is
     * not one of ACGT (upper case) is listed: exc_pos[k] = its index in the dense SEQ column (seq_off units, ascending), exc_byte[k] =
     * the byte.  Sequencer reads are ACGT almost everywhere, so this is a quarter of the ASCII column on the host link. */
    int32_t seq_bits; int32_t reserved0;
    int64_t n_exc; const int64_t* exc_pos; const uint8_t* exc_byte;
    /* 16-bit CIGAR (optional, cig16 != NULL; cig_op / cig_len are then not read): one word per op, op code in bits 0-3 (index into
     * "MIDNSHP=XB"), length in bits 4-15; ops with a length outside 1..4095 or another letter have the word 0xFFFF and are listed:
     * cig_exc_idx[k] = the op's index in the dense CIGAR column (cig_off units, ascending), cig_exc_op / cig_exc_len = the op.
     * (tc_cig_pack16.)  Long-read CIGARs are mostly 1-5-base ops: two fifths of the bytes of the op + length arrays. */
    const uint16_t* cig16; int64_t n_cig_exc; const int64_t* cig_exc_idx; const uint8_t* cig_exc_op; const int32_t* cig_exc_len;
} tc_batch_in;

typedef struct {
    int64_t n_reads;
    const uint32_t* status;   /* n */
    const int32_t* counters;  /* 10 n */
    const int32_t* nm;        /* n; valid with TC_ST_NMMD_DEVICE */
    const int64_t* delta_off; /* n; start of the read's SEQ delta in delta[] (arbitrary order) */
    const int32_t* delta_len; /* n; < 0: the SEQ is the input SEQ */
    const int32_t* seq_len;   /* n; length of the read's new SEQ */
    const uint8_t* delta;     /* SEQ deltas (tc_seq_apply_delta) */
    const int64_t* cig_off;   /* n+1; a read whose CIGAR was not rewritten (no TC_ST_CIGAR_REWRITTEN) has no ops here */
    const uint8_t* cig_op;    /* op letters; 'd' prints as 'D' (a "D-" of the reference's text CIGAR) */
    const int32_t* cig_len;
    const int64_t* md_off;    /* n; start of the read's MD value in md[] (arbitrary order) */
    const int32_t* md_len;    /* n */
    const uint8_t* md;
    const int64_t* jn_off;    /* n+1; junction objects of the read */
    const int8_t* jm;         /* motif codes */
    const int32_t* ji;        /* 2 per junction */
    const int64_t* te_off;    /* n+1; in 8-byte words */
    const uint64_t* te;       /* .TE.log rows as words (tc_te_decode), in reference order (stage order, positional inside a stage) */
} tc_batch_out;

/* device timing of the last batch call, milliseconds (CUDA events on the context's stream) */
typedef struct {
    float h2d_ms, kernels_ms, d2h_ms;
    float k_nmmd_init_ms, k_measure_ms, k_plan_ms, k_junction_ms, k_emit_ms;
    int32_t launches;         /* kernels of this library launched by the call */
    int32_t md_pool_retries;
    int64_t h2d_bytes, d2h_bytes;
    int64_t exact_nmmd_reads; /* reads whose NM/MD needed the exact path (a difference to the genome, a D op, ...) */
    int64_t ascii_path_reads; /* reads the 2-bit check (k_verify) could not settle: their new SEQ was built in ASCII by k_emit */
} tc_timing;

int tc_device_count(void);                  /* GPUs this process can open (0: none); the CLI sizes its worker set with it (--threads, TC.py:560-586) */
tc_ctx* tc_ctx_create(int device);
void tc_ctx_destroy(tc_ctx* ctx);
const char* tc_last_error(tc_ctx* ctx);     /* ctx may be NULL: error of the failed create */
int tc_ctx_set_options(tc_ctx* ctx, const tc_options* opts);
void* tc_ctx_stream(tc_ctx* ctx);           /* cudaStream_t the context launches on */

/* refs.genome (pyfaidx.Fasta in the reference, TC.py:221): 2-bit base plane (16 bases per
 * word, base i in bits 2*(i%16), A=0 C=1 G=2 T=3), lower-case plane (32 bases per word), and
 * literal runs for every base outside ACGTacgt.  chrom_off[c] is the contig's first base in
 * the planes (a multiple of 64). */
int tc_ctx_load_genome(tc_ctx* ctx, int32_t n_chrom, const int64_t* chrom_len,
                       const int64_t* chrom_off, int64_t n_bases_padded,
                       const uint32_t* bases2, const uint32_t* lower1,
                       int64_t n_exc, const int64_t* exc_start, const int64_t* exc_end,
                       const uint8_t* exc_byte);

/* refs.donors / refs.acceptors / refs.sjAnnot (processSpliceAnnotation, TC.py:666-766).
 * *_s: sorted keys (chrom<<33 | strand<<32 | pos0), strand 0 '+', 1 '-';
 * *_u: sorted, de-duplicated keys (chrom<<32 | pos0);
 * annotated junctions: sorted pairs (k1 = chrom<<32 | start, k2 = end<<1 | strand). */
int tc_ctx_load_junctions(tc_ctx* ctx, int64_t n_don_s, const uint64_t* don_s,
                          int64_t n_acc_s, const uint64_t* acc_s,
                          int64_t n_don_u, const uint64_t* don_u,
                          int64_t n_acc_u, const uint64_t* acc_u,
                          int64_t n_annot, const uint64_t* ann_k1, const uint64_t* ann_k2);

/* refs.snps / refs.insertions / refs.deletions (processVCF, TC.py:769-856).
 * snp: sorted keys (chrom<<32 | POS), one row per ALT base.
 * ins: rows sorted by (key = chrom<<32 | POS, size), allele bytes in CSR form.
 * del: rows sorted by (key, size). */
int tc_ctx_load_variants(tc_ctx* ctx, int64_t n_snp, const uint64_t* snp_key, const uint8_t* snp_alt,
                         int64_t n_ins, const uint64_t* ins_key, const int32_t* ins_size,
                         const int64_t* ins_aoff, const uint8_t* ins_bytes,
                         int64_t n_del, const uint64_t* del_key, const int32_t* del_size);

/* Alphabet of the 4-bit SEQ form: code k stands for byte alphabet[k] (n <= 16).  It must hold every
 * byte of the loaded genome (ACGTacgt and the literal-run bytes), because corrected reads copy
 * genome bytes verbatim (Q6); batches with other SEQ bytes are sent as ASCII. */
int tc_ctx_set_alphabet(tc_ctx* ctx, const uint8_t* alphabet, int32_t n);
/* host helpers (tc_host.cpp): ASCII rows -> packed rows (returns the packed size, -1 when a byte is
 * outside the alphabet; poff gets n+1 offsets; out needs sum(ceil(len/8)*4) bytes) and back. */
int64_t tc_seq_pack(const uint8_t* seq, const int64_t* seq_off, int64_t n, const uint8_t* alphabet, int32_t n_alpha,
                    uint8_t* out, int64_t* poff, int32_t n_threads);
int64_t tc_seq_packed_size(const int64_t* seq_off, int64_t n);
/* ASCII rows -> 2-bit rows + the list of bases outside ACGT (at most exc_cap of them, else -1 is returned and the caller
 * uses another form); out needs sum(ceil(len/16)*4) bytes (tc_seq_packed2_size), poff n+1 entries. */
int64_t tc_seq_pack2(const uint8_t* seq, const int64_t* seq_off, int64_t n, uint8_t* out, int64_t* poff,
                     int64_t exc_cap, int64_t* exc_pos, uint8_t* exc_byte, int64_t* n_exc, int32_t n_threads);
int64_t tc_seq_packed2_size(const int64_t* seq_off, int64_t n);
/* CIGAR op + length arrays -> 16-bit words + the list of ops that do not fit (at most exc_cap, else -1 is returned). */
int64_t tc_cig_pack16(const uint8_t* cig_op, const int32_t* cig_len, int64_t n_ops, uint16_t* out, int64_t exc_cap,
                      int64_t* exc_idx, uint8_t* exc_op, int32_t* exc_len, int32_t n_threads);

/* batch_correct() minus text rendering: host columns in, host columns out (H2D, kernels, D2H). */
int tc_correct_batch(tc_ctx* ctx, const tc_batch_in* in, tc_batch_out* out);
/* dryRun() (TC.py:1615-1678): only status, counters, te_off, te are filled. */
int tc_dryrun_batch(tc_ctx* ctx, const tc_batch_in* in, tc_batch_out* out);

/* tc_correct_batch / tc_dryrun_batch cut batches larger than 1.5 x chunk_reads into chunks and
 * overlap upload, kernels and download on three streams (default 131072; 0 = one piece). */
int tc_ctx_set_pipeline(tc_ctx* ctx, int64_t chunk_reads);

/* The same work split for device-resident measurement: upload once, run many, download. */
int tc_batch_upload(tc_ctx* ctx, const tc_batch_in* in);
int tc_batch_run(tc_ctx* ctx, int dry_run);
int tc_batch_download(tc_ctx* ctx, tc_batch_out* out);

int tc_get_timing(tc_ctx* ctx, tc_timing* t);

/* Splice motif of n introns with the junction kernel's motif fetch (getSJMotifCode, spliceJunction.py:71-89; getIntronMotif of
 * accessory_scripts/get_SJs_from_gtf.py:44-66): code[k] = 1 GTAG, 2 CTAC, 3 GCAG, 4 CTGC, 5 ATAC, 6 GTAT, 0 anything else
 * (including an intron end past the contig), -1 where pyfaidx would raise (a coordinate below 1).  Host arrays in and out. */
int tc_intron_motifs(tc_ctx* ctx, int64_t n, const int32_t* chrom, const int32_t* start, const int32_t* end, int8_t* code);

/* ---- host side: SAM text <-> columns (csrc/tc_host.cpp, multi-threaded C++; no GPU involved) ----
 * tc_sam_parse replaces split_SAM (TranscriptClean.py:519-541) + the field / tag handling of
 * Transcript.__init__ (transcript.py:10-72): header lines are set aside, alignment lines become a
 * tc_batch_in.  `text` must stay alive as long as the tc_sam (fields are referenced, not copied).
 * Returns NULL and fills `err` for records the reference would crash on (see sam.py). */
typedef struct tc_sam tc_sam;
tc_sam* tc_sam_parse(const char* text, int64_t len, int32_t n_chrom, const char* const* chrom_names,
                     const int64_t* chrom_len, int32_t n_threads, char* err, int64_t err_cap);
void tc_sam_free(tc_sam* s);
/* distinct RNAMEs of the alignment lines, newline-joined (split_SAM's chromosome set, TC.py:535); free with tc_free */
char* tc_sam_rnames(const char* text, int64_t len, int64_t* out_len);
void tc_free(void* p);
/* the '@' lines of a SAM text (trimmed, newline-terminated), for callers that stream a file in blocks; free with tc_free */
char* tc_sam_header_lines(const char* text, int64_t len, int64_t* out_len);
int64_t tc_sam_n_reads(const tc_sam* s);
int64_t tc_sam_header(const tc_sam* s, const char** text);      /* '@' lines, newline-terminated */
void tc_sam_batch(const tc_sam* s, tc_batch_in* out);            /* view of the parsed columns */
/* packs the SEQ column to the 4-bit form; 1 = packed, 0 = a SEQ byte is outside the alphabet (stays ASCII) */
int tc_sam_pack_seq(tc_sam* s, const uint8_t* alphabet, int32_t n_alpha, int32_t n_threads);
/* printableSAM / printableFa (transcript.py:244-268), create_log_string (TC.py:1591-1604), the TE
 * rows and the emission rules of batch_correct (TC.py:373-385).  The four texts are owned by `s`. */
int tc_sam_render(tc_sam* s, const tc_batch_out* res, int32_t primary_only, int32_t canon_only, int32_t dry,
                  int32_t n_threads, const char** sam, int64_t* sam_len, const char** fa, int64_t* fa_len,
                  const char** log, int64_t* log_len, const char** te, int64_t* te_len, char* err, int64_t err_cap);
/* the same texts without the final copy: n_parts pieces per text, in order (the formatter threads' own buffers);
 * ptr[4 * k + q] / len[4 * k + q] = piece k of text q (0 sam, 1 fa, 2 log, 3 te); ptr / len hold 4 * max_parts entries.
 * Valid until the tc_sam is rendered again or freed (the pieces are its own storage). */
int tc_sam_render_parts(tc_sam* s, const tc_batch_out* res, int32_t primary_only, int32_t canon_only, int32_t dry,
                        int32_t n_threads, int32_t* n_parts, const char** ptr, int64_t* len, int32_t max_parts,
                        char* err, int64_t err_cap);

/* ---- host side: lookup tables from their files (csrc/tc_tables.cpp, multi-threaded C++; no GPU involved) ----
 * tc_vcf_parse replaces processVCF (TranscriptClean.py:769-856): VCF text (already decompressed) -> the arrays of
 * tc_ctx_load_variants.  tc_sj_parse replaces processSpliceAnnotation (TranscriptClean.py:666-766): STAR SJ.out.tab
 * text -> the arrays of tc_ctx_load_junctions.  `use[c]` = 0 drops rows of contig c (contigs no read maps to, TC.py:685;
 * NULL keeps all).  NULL + `err` for rows the reference's int() / indexing would raise on. */
typedef struct tc_vcf tc_vcf;
tc_vcf* tc_vcf_parse(const char* text, int64_t len, int32_t n_chrom, const char* const* chrom_names, const uint8_t* use,
                     int32_t max_len_indel, int32_t n_threads, char* err, int64_t err_cap);
void tc_vcf_tables(const tc_vcf* v, int64_t* n_snp, const uint64_t** snp_key, const uint8_t** snp_alt, int64_t* n_ins,
                   const uint64_t** ins_key, const int32_t** ins_size, const int64_t** ins_aoff, const uint8_t** ins_bytes,
                   int64_t* n_del, const uint64_t** del_key, const int32_t** del_size);
int64_t tc_vcf_n_records(const tc_vcf* v);
/* distinct CHROM values of a VCF text, each followed by a NUL (validate_chroms, TranscriptClean.py:497-505); tc_free */
char* tc_vcf_chroms(const char* text, int64_t len, int64_t* out_len);
void tc_vcf_free(tc_vcf* v);
typedef struct tc_sj tc_sj;
tc_sj* tc_sj_parse(const char* text, int64_t len, int32_t n_chrom, const char* const* chrom_names, const uint8_t* use,
                   char* err, int64_t err_cap);
void tc_sj_tables(const tc_sj* s, int64_t* n_don_s, const uint64_t** don_s, int64_t* n_acc_s, const uint64_t** acc_s,
                  int64_t* n_don_u, const uint64_t** don_u, int64_t* n_acc_u, const uint64_t** acc_u,
                  int64_t* n_annot, const uint64_t** ann_k1, const uint64_t** ann_k2);
void tc_sj_free(tc_sj* s);

/* refs.genome from its FASTA (pyfaidx.Fasta(options.refGenome), TranscriptClean.py:221): the file is mapped and packed by all host
 * threads into the planes of tc_ctx_load_genome.  tc_fasta_open finds the contigs (name = first word of the header line, pyfaidx's
 * key, TC.py:469-470) and their lengths; tc_fasta_pack fills zeroed planes (chrom_off[k] = plane index of contig k's first base,
 * n_pad = bases the planes hold) and returns the number of literal runs (every byte outside ACGTacgt), which tc_fasta_exceptions
 * copies out.  GenomeImage.from_fasta (genome.py) is the same logic in numpy. */
typedef struct tc_fasta tc_fasta;
tc_fasta* tc_fasta_open(const char* path, int32_t n_threads, char* err, int64_t err_cap);
int32_t tc_fasta_n_contigs(const tc_fasta* f);
int tc_fasta_contig(const tc_fasta* f, int32_t k, const char** name, int32_t* name_len, int64_t* length);
int64_t tc_fasta_pack(tc_fasta* f, const int64_t* chrom_off, int64_t n_pad, uint32_t* bases2, uint32_t* lower1, int32_t n_threads);
int64_t tc_fasta_exceptions(const tc_fasta* f, int64_t* start, int64_t* end, uint8_t* byte);
void tc_fasta_free(tc_fasta* f);

#ifdef __cplusplus
}
#endif
#endif

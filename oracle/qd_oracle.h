/* quickdna hot-path ORACLE -- CPU restatement of the reference's algorithms.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under quickdna_b200/ may include, link or load
 * this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / the timed CPU baseline.
 *
 * Parity status: PINNED to the reference's golden artefacts -- the regenerated LUT
 * must hash to the sha256 of /root/reference/src/tables.dat, and every known-answer
 * vector in the reference's own tests is replayed against
 * these functions in tests/test_oracle_golden.py.  The reference itself (Rust) cannot
 * be compiled in this image (no cargo/rustc), so this is a "port", not the reference.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference/).
 */
#ifndef QD_ORACLE_H
#define QD_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QDO_CODONS_PER_TABLE 4096 /* src/trans_table.rs:107 */
#define QDO_N_TRANS_TABLES 27     /* src/trans_table.rs:109 */

/* src/errors.rs:17-29 */
enum {
    QDO_OK = 0,
    QDO_NON_ASCII_BYTE = 1,
    QDO_BAD_NUCLEOTIDE = 2,
    QDO_UNEXPECTED_AMBIGUOUS = 3,
    QDO_BAD_TABLE = 4
};

typedef struct {
    int32_t kind;  /* one of the enum above */
    uint8_t byte;  /* payload: offending byte (upper-cased letter for kind 3), or table id */
    uint64_t pos;  /* index of the offending byte in the input (derived from loop order) */
} qdo_err;

/* Tables -------------------------------------------------------------------------- */
const uint8_t *qdo_tables(void);      /* 27*4096 bytes, same layout as src/tables.dat */
int qdo_table_slot(int ncbi_id);      /* src/trans_table.rs:114-146, 230-267; -1 if invalid */

/* Alphabet ------------------------------------------------------------------------ */
/* src/nucleotide.rs:66-101, 268-315. Returns QDO_OK and *mask, or the error kind and *payload. */
int qdo_decode_byte(uint8_t b, int strict, uint8_t *mask, uint8_t *payload);
uint8_t qdo_complement_mask(uint8_t m); /* src/nucleotide.rs:125-132, 195-213 */
uint8_t qdo_mask_to_ascii(uint8_t m);   /* src/nucleotide.rs:138-145, 219-237 */
/* src/errors.rs:19-28 Display text; returns length written (excluding NUL) */
int qdo_err_message(const qdo_err *e, char *buf, size_t cap);

/* Translation / reverse complement ----------------------------------------------------- */
/* src/trans_table.rs:176-203 (+ python_api.rs:33-51: table checked first). out: n/3 bytes. */
int qdo_translate_bytes(int ncbi_id, int strict, const uint8_t *dna, size_t n, uint8_t *out,
                        qdo_err *err);
/* src/trans_table.rs:205-227. out: n/3 bytes. */
void qdo_translate_masks(int slot, const uint8_t *masks, size_t n, uint8_t *out);
/* src/trans_table.rs:269-278. out: n bytes. */
int qdo_revcomp_bytes(int strict, const uint8_t *dna, size_t n, uint8_t *out, qdo_err *err);
/* src/trans_table.rs:284-286. out: n bytes. */
void qdo_revcomp_masks(const uint8_t *masks, size_t n, uint8_t *out);

/* Frames. `out` receives the frames back to back (f0|f1|f2[|r0|r1|r2]); lens[k] the length of
 * frame k (0 for an omitted frame); returns the number of frames the reference returns
 * (src/rust_api.rs:227-255: 3 if n>=5, 2 if n==4, 1 if n==3, else 0; :263-270 doubles it). */
int qdo_self_frames_masks(int slot, const uint8_t *masks, size_t n, uint8_t *out, size_t lens[3]);
int qdo_all_frames_masks(int slot, const uint8_t *masks, size_t n, uint8_t *out, size_t lens[6]);
/* Python path: quickdna/__init__.py:128-185 over python_api.rs (ASCII, fallible; the
 * reverse complement inside translate_all_frames is NON-strict even when strict=1, :182).
 * Returns 0 and *n_frames, or the error kind (err filled; pos relative to dna). */
int qdo_py_self_frames(int ncbi_id, int strict, const uint8_t *dna, size_t n, uint8_t *out,
                       size_t lens[3], int *n_frames, qdo_err *err);
int qdo_py_all_frames(int ncbi_id, int strict, const uint8_t *dna, size_t n, uint8_t *out,
                      size_t lens[6], int *n_frames, qdo_err *err);

/* src/rust_api.rs:310-322: ASCII -> masks, dropping ' ' and '\t'. masks: cap n. */
int qdo_parse(int strict, const uint8_t *ascii, size_t n, uint8_t *masks, size_t *n_masks,
              qdo_err *err);

/* Canonicalisation (strict masks in {1,2,4,8}) ----------------------------------------- */
void qdo_forward_canonical(const uint8_t *masks, size_t n, uint8_t *out); /* src/canonical.rs:71-118 */
void qdo_canonical(const uint8_t *masks, size_t n, uint8_t *out); /* src/canonical.rs:17-55,129-179 */
/* every length-w window canonicalised, benches/canonical.rs:29-34; out: (n-w+1)*w */
size_t qdo_canonical_windows(const uint8_t *masks, size_t n, size_t w, uint8_t *out);

/* Expansions ---------------------------------------------------------------------- */
/* src/expansions.rs:96-107 for a fresh iterator: returns 1 and *count, or 0 on overflow
 * ((usize::MAX, None)). */
int qdo_expansions_size_hint(const uint8_t *masks, size_t n, uint64_t *count);
/* src/expansions.rs:37-94: writes up to max_out expansions of n masks each, in iteration
 * order; returns how many were written. */
size_t qdo_expansions(const uint8_t *masks, size_t n, uint8_t *out, size_t max_out);

/* Windows ------------------------------------------------------------------------- */
/* src/rust_api.rs:100-104, 277-279: all length-w windows copied out; returns their count. */
size_t qdo_windows(const uint8_t *seq, size_t n, size_t w, uint8_t *out);
/* benches/all_windows.rs:78-102 + :39-75 (SequenceWindows::from_dna + enumerate) for strict
 * masks: DNA windows (ASCII), RC-DNA windows (ASCII), then 6 frames' aa windows, each with its
 * origin index.  dna_out/aa_out may be NULL to count only.
 *   dna_out:  n_dna_windows * w_dna bytes (fwd windows then rc windows)
 *   aa_out:   n_aa_windows * w_aa bytes (frames 0..5 in order)
 *   origin:   one uint64 per window, in enumerate() order
 * counts[0]=fwd dna windows, [1]=rc dna windows, [2..7]=aa windows per frame. */
size_t qdo_windows_all(const uint8_t *masks, size_t n, size_t w_dna, size_t w_aa, uint8_t *dna_out,
                       uint8_t *aa_out, uint64_t *origin, size_t counts[8]);

/* FASTA ---------------------------------------------------------------------------- */
/* FastaParser::<DnaSequence<T>>::parse, src/fasta.rs:207-415 (state machine) + :425-480 (driver): lines split as
 * BufRead::lines does ('\n', with one '\r' before it dropped), header lines start with '>' or ';' (:482-491),
 * content lines go through TryFrom<&[u8]> for DnaSequence (src/rust_api.rs:310-322).
 * A record: header text bytes [header_begin, header_end) of `text` (first header line start to the end of the
 * last header line, line break excluded; header_begin == header_end == UINT64_MAX for the headerless record made of
 * preceding content), masks [content_begin, content_end) of `masks`, line_range [line_begin, line_end) 1-based.
 * On a content error: err->pos = byte offset in `text`, *err_line = its 1-based line number.
 * Returns the number of records (records beyond max_records are counted, not stored). */
typedef struct {
    uint64_t header_begin, header_end, content_begin, content_end, line_begin, line_end;
} qdo_fasta_record;
int qdo_fasta_scan(int strict, int concatenate_headers, int allow_preceding_comment, const uint8_t *text, size_t n,
                   uint8_t *masks, size_t *n_masks, qdo_fasta_record *records, size_t max_records, size_t *n_records,
                   qdo_err *err, uint64_t *err_line);

/* Timed CPU baseline helpers (bench.py cpu_baseline / --impl reference) ------------ */
/* Batch of n_reads reads of read_len ASCII nts each: per read, the Rust-API work of
 * src/rust_api.rs:263-270 after the parse of :310-322 (decode, 3 passes, materialised
 * reverse complement, 3 more passes), split over `threads` pthreads by read count.
 * out: n_reads * 2*(read_len-2) bytes.  Returns seconds of wall time. */
double qdo_bench_all_frames_batch(int slot, const uint8_t *dna, size_t n_reads, size_t read_len,
                                  uint8_t *out, int threads, int reps);
double qdo_bench_revcomp(const uint8_t *dna, size_t n, uint8_t *out, int threads, int reps);

/* Synthetic input + whole-config checks (tests and bench.py's checksum asserts) ------------- */
/* quickdna_b200/synth.py synth_np, byte for byte: the counter-based stream every config's input is cut from */
void qdo_synth(uint64_t seed, uint64_t start, uint64_t count, const uint8_t *alphabet, size_t alen, uint8_t *out);
/* The position-weighted linear checksum qd_checksum64_dev computes (include/quickdna_b200.h): `buf` holds bytes
 * [base, base + n) of a byte stream; sum over stream words J of word_J * (mix64(J * GOLD + 1) | 1) mod 2^64. */
uint64_t qdo_checksum64(const uint8_t *buf, size_t n, uint64_t base);
/* What the GPU path must produce, as that checksum, computed with the functions above on `threads` threads:
 *  - frames of uniform reads [first_read, first_read + n_reads) of the seed's ACGT stream, slot layout
 *    (the output stream starts at read first_read);
 *  - reverse_complement_bytes of stream bytes [start, start + n);
 *  - canonical w-nt windows (kind 0) or the all_windows output (kind 1: dna_out stream then aa_out stream) of n_seq
 *    strict ASCII sequences, in the layouts of qd_canonical_windows_batch_dev / qd_windows_all_batch_dev;
 *  - expansion rows of n_win ASCII windows (windows with more than `cap` expansions skipped). */
uint64_t qdo_check_frames_uniform(uint64_t seed, uint64_t first_read, uint64_t n_reads, size_t read_len, int slot, int frames,
                                  int threads);
uint64_t qdo_check_revcomp(uint64_t seed, uint64_t start, uint64_t n, int strict, int threads);
uint64_t qdo_check_windows(int kind, const uint8_t *dna, size_t n_seq, size_t seq_len, size_t w, size_t w_aa, int threads);
uint64_t qdo_check_expansions(const uint8_t *windows, size_t n_win, size_t w, uint64_t cap, int threads, uint64_t *total_rows);

#ifdef __cplusplus
}
#endif
#endif

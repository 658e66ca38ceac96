/* quickdna_b200 -- C ABI of the B200-native quickdna sequence hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / CUDA types in the
 * signatures (a CUDA stream crosses as `void*`).  Each entry point names the reference
 * interface whose body it replaces (paths relative to the SecureDNA/quickdna checkout).
 * INTEGRATION.md shows the Rust `extern "C"` block and the PyO3 shims that bind it.
 *
 * Conventions kept from the reference:
 *   - inputs are borrowed; outputs are caller-allocated (every size is a closed form of n);
 *   - the translation table id is validated BEFORE any DNA byte (src/python_api.rs:35,48);
 *   - errors carry the offending byte exactly as src/errors.rs:17-29 does (`kind` + `byte`
 *     rebuild the reference's message, see qd_err_message); `pos` / `seq_index` are the
 *     extra "first error position" the north star asks for;
 *   - on error the output contents are unspecified (the reference drops its Vec).
 *
 * There is NO CPU fallback: every compute entry point runs sm_100a kernels and returns
 * QD_ERR_CUDA if no usable device / kernel image is present.
 */
#ifndef QUICKDNA_B200_H
#define QUICKDNA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QD_ABI_VERSION 1

/* ---- status codes (return value of every int-returning entry point) ------------------ */
enum {
    QD_OK = 0,
    QD_ERR_NON_ASCII_BYTE = 1,       /* TranslationError::NonAsciiByte(u8)                    */
    QD_ERR_BAD_NUCLEOTIDE = 2,       /* TranslationError::BadNucleotide(char)                 */
    QD_ERR_UNEXPECTED_AMBIGUOUS = 3, /* TranslationError::UnexpectedAmbiguousNucleotide(char) */
    QD_ERR_BAD_TABLE = 4,            /* TranslationError::BadTranslationTable(u8)             */
    QD_ERR_BAD_MASK = 5,             /* MASK encoding only: byte is not a valid IUPAC mask     */
    QD_ERR_INVALID_ARGUMENT = 90,
    QD_ERR_CUDA = 100                /* CUDA runtime failure; qd_last_cuda_error() has text   */
};

/* ---- encodings ------------------------------------------------------------------------ */
enum {
    QD_ENC_ASCII = 0, /* one ASCII letter per nucleotide (PyO3 path, src/python_api.rs)        */
    QD_ENC_MASK = 1   /* one byte holding the 4-bit IUPAC set A=1,T=2,C=4,G=8 -- the in-memory
                         value of `Nucleotide` / `NucleotideAmbiguous` (src/nucleotide.rs:20-53),
                         i.e. a `&[T]` of the typed Rust API reinterpreted as bytes             */
};

/* ---- frame selection -------------------------------------------------------------------- */
enum {
    QD_FRAMES_ONE = 1,  /* translate:               src/rust_api.rs:216-219                    */
    QD_FRAMES_SELF = 3, /* translate_self_frames:   src/rust_api.rs:227-255                    */
    QD_FRAMES_ALL = 6   /* translate_all_frames:    src/rust_api.rs:263-270                    */
};

typedef struct qd_ctx qd_ctx;

typedef struct qd_err {
    int32_t kind;       /* QD_OK or one of QD_ERR_* above                                      */
    uint8_t byte;       /* payload of the reference's error variant (offending byte; upper-case
                           letter for kind 3; table id for kind 4)                              */
    uint64_t pos;       /* index of the first offending byte within its sequence               */
    uint64_t seq_index; /* sequence index within a batch (0 for single-sequence calls)         */
} qd_err;

/* ---- context ----------------------------------------------------------------------------
 * One context owns one device, one stream, the device-resident LUTs and grow-only scratch
 * buffers.  Entry points taking a context are serialised on an internal mutex (the
 * reference's functions are pure and re-entrant; use one context per thread to run in
 * parallel).  Scratch and the device-side error key belong to the context, not to a stream:
 * a device-pointer call leaves an event behind and the next call on the same context that
 * runs on ANOTHER stream waits for it (host-pointer calls wait on the host), so calls on one
 * context execute in the order they were issued whatever streams they use; work that should
 * overlap needs contexts of its own.  Every entry point makes the context's device current
 * for its duration and restores the caller's device before returning. */
int qd_ctx_create(qd_ctx **out, int device);
void qd_ctx_destroy(qd_ctx *ctx);
int qd_ctx_device(const qd_ctx *ctx);
void *qd_ctx_stream(const qd_ctx *ctx); /* the context's cudaStream_t */
int qd_ctx_sm_count(const qd_ctx *ctx);
const char *qd_last_cuda_error(const qd_ctx *ctx);
int qd_abi_version(void);
/* input bytes per chunk of the host batch pipeline (default 64 MiB) */
int qd_ctx_set_chunk_bytes(qd_ctx *ctx, size_t bytes);
/* Host entry points take any host pointer.  Pageable buffers (a Rust Vec<u8>, Python bytes, a numpy array) of large calls
 * (>= 32 MiB in + out) are staged through pinned buffers of the context, 16 MiB of input per chunk, filled and drained by
 * `threads` persistent host threads, so the PCIe copies stay asynchronous; pinned (cudaHostAlloc / cudaHostRegister)
 * buffers are copied from and to directly.  threads: -1 = half the host's hardware threads, between 2 and 8 (default);
 * 0 = no staging (the driver's own staged copies). */
int qd_ctx_set_stage_threads(qd_ctx *ctx, int threads);
/* Pinned (page-locked) host memory for callers that want the direct path: buffers from qd_host_alloc are copied from and to
 * without staging (24.7 against 13.3 Gnt/s for a six-frame batch).  Portable across the devices of the process.
 * qd_host_free(NULL) is a no-op. */
void *qd_host_alloc(size_t bytes);
void qd_host_free(void *p);
/* tuning / test knob: frame launches of at least `runs` 48-nt runs use the 1024-thread tile
 * (default 2 * 1024 * SM count); 0 = always, UINT64_MAX = never */
int qd_ctx_set_big_tile_min_runs(qd_ctx *ctx, uint64_t runs);

/* ---- tables / errors (host-only helpers) --------------------------------------------------
 * qd_table_slot:  TranslationTable::try_from(u8) + table_index(), src/trans_table.rs:114-146,
 *                 230-267.  Returns the LUT slot 0..26 or -1 (=> BadTranslationTable).
 * qd_table_data:  the 4096-byte LUT of a slot in the reference's layout, index
 *                 (m0<<8)|(m1<<4)|m2 (src/trans_table.rs:78-85); all 27 are contiguous, so
 *                 qd_table_data(0) addresses a byte-for-byte image of src/tables.dat.
 * qd_err_message: Display of TranslationError, src/errors.rs:19-28 (what PyO3 puts into the
 *                 ValueError, src/python_api.rs:15-19).  Returns the text length. */
int qd_table_slot(int ncbi_id);
const uint8_t *qd_table_data(int slot);
int qd_err_message(const qd_err *err, char *buf, size_t cap);

/* ---- output geometry -----------------------------------------------------------------------
 * Frame k (0..2, either strand) of an n-nt sequence has floor((n-k)/3) residues when n > k
 * (src/rust_api.rs:227-255); the reference returns only non-empty frames.
 *
 * Device / batch layout ("slots"): a sequence is cut into runs of 48 nt; with
 * R = ceil(n/48) each of the F frames (F = 1, 3 or 6) owns a slot of 16*R bytes, slots back to
 * back, so one sequence occupies 16*R*F bytes and every kernel store is an aligned 128-bit
 * store.  Forward frame k starts at the beginning of slot k; reverse-complement frame k ENDS at
 * the end of slot 3+k (it is produced back to front).  Bytes of a slot outside its frame are 0.
 */
size_t qd_frame_len(size_t n, int k);
size_t qd_runs(size_t n);                                           /* R = ceil(n/48)        */
size_t qd_slots_bytes(size_t n, int frames);                        /* 16*R*frames           */
size_t qd_slot_frame_offset(size_t n, int frames, int frame_index); /* first residue's offset */

/* ======================= host-pointer entry points (synchronous) =========================== */

/* translate_dna_bytes::<T> / translate_dna::<T>: src/trans_table.rs:176-203, 205-227
 * (PyO3 _translate/_translate_strict, src/python_api.rs:33-51; DnaSequence::translate,
 * src/rust_api.rs:216-219).  out: floor(n/3) bytes.  ASCII: validates the first 3*floor(n/3)
 * bytes; strict rejects ambiguity codes. */
int qd_translate(qd_ctx *ctx, int table, int enc, int strict, const uint8_t *dna, size_t n,
                 uint8_t *out, qd_err *err);

/* translate_self_frames / translate_all_frames: src/rust_api.rs:227-255, 263-270 (the Python
 * wrappers at quickdna/__init__.py:128-185 make 3 / 7 native calls for the same result).
 * frames = QD_FRAMES_SELF or QD_FRAMES_ALL.  out: frames written back to back, tightly
 * (f0|f1|f2[|r0|r1|r2]); out_len[k] = residues of frame k (0 when the reference omits it);
 * *n_frames = number of frames the reference returns (6/4/2/0 or 3/2/1/0).  Capacity needed:
 * sum of qd_frame_len(n,k) over the frames.  ASCII: all n bytes are validated when n >= 3,
 * none when n < 3. */
int qd_translate_frames(qd_ctx *ctx, int table, int enc, int strict, int frames,
                        const uint8_t *dna, size_t n, uint8_t *out, size_t out_len[6],
                        int *n_frames, qd_err *err);

/* reverse_complement_bytes::<T> / reverse_complement::<T>: src/trans_table.rs:269-278, 284-286
 * (PyO3 _reverse_complement[_strict], src/python_api.rs:58-74; DnaSequence::reverse_complement,
 * src/rust_api.rs:273-275).  out: n bytes, same encoding as the input; ASCII output is upper
 * case.  Every byte is validated. */
int qd_revcomp(qd_ctx *ctx, int enc, int strict, const uint8_t *dna, size_t n, uint8_t *out,
               qd_err *err);

/* Batch of n_seq sequences stored back to back in `dna` with CSR offsets[n_seq+1]
 * (offsets[0] = 0): what a caller looping DnaSequence::translate* over a Vec of reads does,
 * in one call.  out: slot layout above, sequence i at out_offsets[i] = 16*frames*sum_{j<i}R_j;
 * out_offsets (n_seq+1 entries) is filled when non-NULL.  Capacity: qd_batch_out_bytes().
 * Work is pipelined in chunks of whole sequences (H2D, kernel, D2H overlap).
 * Offsets that decrease: QD_ERR_INVALID_ARGUMENT, err->seq_index = the first such sequence, nothing is launched. */
size_t qd_batch_out_bytes(const uint64_t *offsets, size_t n_seq, int frames);
int qd_translate_frames_batch(qd_ctx *ctx, int table, int enc, int strict, int frames,
                              const uint8_t *dna, const uint64_t *offsets, size_t n_seq,
                              uint8_t *out, uint64_t *out_offsets, qd_err *err);
/* Same for n_seq sequences of identical length (no offsets array; closed-form geometry:
 * sequence i at dna + i*seq_len and out + i*qd_slots_bytes(seq_len, frames)). */
int qd_translate_frames_batch_uniform(qd_ctx *ctx, int table, int enc, int strict, int frames,
                                      const uint8_t *dna, size_t n_seq, size_t seq_len,
                                      uint8_t *out, qd_err *err);

/* TryFrom<&[u8]> / FromStr for DnaSequence<T>: src/rust_api.rs:310-338.  ASCII in, one mask byte per kept
 * nucleotide out: ' ' and '\t' are dropped, every other byte must be a nucleotide letter of the alphabet
 * (strict != 0: Nucleotide, else NucleotideAmbiguous).  masks: capacity n; *n_masks = masks written.
 * On error `pos` is the index of the offending byte in the INPUT. */
int qd_parse(qd_ctx *ctx, int strict, const uint8_t *ascii, size_t n, uint8_t *masks, size_t *n_masks,
             qd_err *err);

/* FastaParser::<DnaSequence<T>>::parse, src/fasta.rs:207-491 (the format feeding batches): header lines start with
 * '>' or ';', lines end at '\n' (a '\r' right before it is dropped), content lines go through the parse above.
 * masks: one mask per nucleotide of all records back to back (capacity n); records[i] = header bytes
 * [header_begin, header_end) of `text` (first header line start to the end of the last header line; both UINT64_MAX
 * for the headerless record made of content before the first header), masks [content_begin, content_end) -- so
 * content_begin/_end are the CSR offsets the batch entry points take -- and line_range [line_begin, line_end).
 * concatenate_headers / allow_preceding_comment: FastaParseSettings, src/fasta.rs:67-110 (defaults 1 / 0).
 * Call with masks == records == NULL to get *n_masks and *n_records only; a call whose max_records is too small stores
 * the needed *n_records and returns QD_ERR_INVALID_ARGUMENT (so one call with masks[n] and a generous guess, e.g.
 * n / 64 + 16 records, is the usual case).  On a content error err->pos is the byte
 * offset in `text` and err->seq_index the 1-based line number (Located::line_number, src/errors.rs:8-15). */
typedef struct qd_fasta_record {
    uint64_t header_begin, header_end, content_begin, content_end, line_begin, line_end;
} qd_fasta_record;
int qd_fasta_scan(qd_ctx *ctx, int strict, int concatenate_headers, int allow_preceding_comment,
                  const uint8_t *text, size_t n, uint8_t *masks, size_t *n_masks, qd_fasta_record *records,
                  size_t max_records, size_t *n_records, qd_err *err);

/* qd_fasta_scan on device buffers (asynchronous up to ONE synchronisation of `stream`: the totals of the first pass size the
 * second): text_dev in; masks_dev (capacity n), records_dev (capacity max_records) and, optionally, offsets_dev
 * (max_records + 1 entries: the CSR offsets qd_translate_frames_batch_dev takes with QD_ENC_MASK) out.  *n_masks / *n_records
 * are host values.  masks_dev == records_dev == NULL counts only. */
int qd_fasta_scan_dev(qd_ctx *ctx, int strict, int concatenate_headers, int allow_preceding_comment,
                      const uint8_t *text_dev, size_t n, uint8_t *masks_dev, size_t *n_masks,
                      qd_fasta_record *records_dev, size_t max_records, size_t *n_records, uint64_t *offsets_dev,
                      qd_err *err, void *stream);

/* FASTA text in, the frames of every record out: FastaParser::parse (src/fasta.rs:207-491) followed by
 * translate / translate_self_frames / translate_all_frames (src/rust_api.rs:227-270) of every record, in one call.  The
 * scanner's masks and CSR offsets feed the frame kernels on the device: only the text goes up, only records and frames
 * come back.  records[i] as qd_fasta_scan (content_begin/_end = the record's nucleotide range); out = the slot layout of
 * qd_translate_frames_batch over the records (record i at out_offsets[i], frames at qd_slot_frame_offset(len_i, frames, f)),
 * out_offsets: max_records + 1 entries or NULL.  *out_bytes = bytes needed / written (at most
 * 16 * frames * (n / 48 + n_records)).  Call with records == out == NULL for *n_records and *out_bytes only; too small a
 * capacity stores the needed values and returns QD_ERR_INVALID_ARGUMENT.  Errors as qd_fasta_scan (table id first). */
int qd_fasta_translate_frames(qd_ctx *ctx, int table, int strict, int frames, int concatenate_headers,
                              int allow_preceding_comment, const uint8_t *text, size_t n, qd_fasta_record *records,
                              size_t max_records, size_t *n_records, uint8_t *out, size_t out_capacity,
                              uint64_t *out_offsets, size_t *out_bytes, qd_err *err);

/* DnaSequence::<Nucleotide>::canonical over every length-w window: src/canonical.rs:17-55,
 * 71-118, 129-179 applied as in benches/canonical.rs:29-34.  Strict alphabet only.
 * forward_only != 0 gives ForwardCanonical.  out: (n-w+1)*w bytes (0 windows if n < w),
 * window i at out + i*w, same encoding as the input (ASCII output upper case). 1 <= w <= 64. */
int qd_canonical_windows(qd_ctx *ctx, int enc, int forward_only, const uint8_t *dna, size_t n,
                         size_t w, uint8_t *out, qd_err *err);

/* DnaSequence::<Nucleotide>::canonical of ONE sequence of any length: src/rust_api.rs:374-377 over
 * src/canonical.rs:17-55 (min_lex(fw(seq), fw(reverse(seq))), A<T<C<G).  out: n bytes, input encoding. */
int qd_canonical(qd_ctx *ctx, int enc, int forward_only, const uint8_t *dna, size_t n, uint8_t *out,
                 qd_err *err);

/* DnaSequence::<NucleotideAmbiguous>::expansions over a batch of n_win windows of w nt each
 * (benches/expansions.rs:126-130 shape): src/expansions.rs:35-108.
 *   counts[i]      = Expansions::size_hint() of window i (UINT64_MAX on overflow == (usize::MAX, None))
 *   out_index[i]   = index of window i's first expansion in `out`, or UINT64_MAX when
 *                    counts[i] > max_expansions (such windows are skipped, as
 *                    benches/expansions.rs:57-66 does); out_index[n_win] = total expansions
 *   out            = expansions in iteration order, w bytes each (strict alphabet, input encoding)
 * Call with out == NULL to get counts / out_index only.  out capacity:
 * out_index[n_win]*w, at most n_win*max_expansions*w.  w >= 1 (n_win = 1, w = n is DnaSequence::expansions()). */
int qd_expansion_windows(qd_ctx *ctx, int enc, const uint8_t *windows, size_t n_win, size_t w,
                         uint64_t max_expansions, uint64_t *counts, uint64_t *out_index,
                         uint8_t *out, qd_err *err);

/* DnaSequence::windows / ProteinSequence::windows, src/rust_api.rs:100-104, 277-279:
 * out + i*w holds seq[i..i+w); returns the window count through *n_windows. */
int qd_windows(qd_ctx *ctx, const uint8_t *seq, size_t n, size_t w, uint8_t *out,
               size_t *n_windows);

/* benches/all_windows.rs:78-102 (SequenceWindows::from_dna, table Ncbi1) + :39-75 (enumerate)
 * for one strict sequence: forward DNA windows, reverse-complement DNA windows (ASCII), then the
 * aa windows of the frames the reference returns, each with its origin index.
 *   counts[0..1] = DNA / RC-DNA window counts, counts[2..7] = aa windows per returned frame
 *   dna_out: (counts[0]+counts[1])*w_dna bytes; aa_out: sum(counts[2..7])*w_aa bytes;
 *   origin: one uint64 per window in enumerate() order.  Any of the three may be NULL. */
int qd_windows_all(qd_ctx *ctx, int enc, const uint8_t *dna, size_t n, size_t w_dna, size_t w_aa,
                   uint8_t *dna_out, uint8_t *aa_out, uint64_t *origin, size_t counts[8],
                   qd_err *err);

/* =================== device-pointer entry points (asynchronous) ============================
 * Same arithmetic on buffers already resident in HBM; launched on `stream` (NULL = the
 * legacy default stream, as in the CUDA runtime) and NOT synchronised.  `dna` may have any alignment; kernels read whole
 * 16-byte granules, so the granules containing the first and last byte must be readable (true
 * for any cudaMalloc / torch allocation).  `out` must be 16-byte aligned and sized by the slot
 * layout (qd_slots_bytes / qd_batch_out_bytes).  Errors are reported through a device-side
 * key: call qd_dev_error_reset before a launch sequence and qd_dev_error_fetch (which
 * synchronises the stream) after it. */
int qd_dev_error_reset(qd_ctx *ctx, void *stream);
/* dna/offsets: the device buffers the launches ran on (offsets NULL for a single sequence or
 * uniform batch with seq_len given); fills err like the host entry points do. */
int qd_dev_error_fetch(qd_ctx *ctx, void *stream, int enc, int strict, const uint8_t *dna_dev,
                       const uint64_t *offsets_dev, size_t n_seq, size_t uniform_len, qd_err *err);

int qd_revcomp_dev(qd_ctx *ctx, int enc, int strict, const uint8_t *dna_dev, size_t n,
                   uint8_t *out_dev, void *stream);
/* single sequence, slot layout */
int qd_translate_frames_dev(qd_ctx *ctx, int table, int enc, int strict, int frames,
                            const uint8_t *dna_dev, size_t n, uint8_t *out_dev, void *stream);
int qd_translate_frames_batch_uniform_dev(qd_ctx *ctx, int table, int enc, int strict, int frames,
                                          const uint8_t *dna_dev, size_t n_seq, size_t seq_len,
                                          uint8_t *out_dev, void *stream);
/* variable lengths: CSR offsets on the device, offsets_dev[0] == 0 (sequence i is dna_dev[offsets[i] .. offsets[i+1]));
 * the run prefix, the padded offsets copy and the tile maps the kernel needs are built on `stream` in context-owned
 * scratch.  The host cannot look at device offsets, so the scan checks them: a sequence whose offsets decrease or point
 * past total_nt is skipped (no byte outside dna_dev[0 .. total_nt) is read) and qd_dev_error_fetch returns
 * QD_ERR_INVALID_ARGUMENT with err->seq_index = the first such sequence -- unless a bad nucleotide was found, which is
 * reported first. */
int qd_translate_frames_batch_dev(qd_ctx *ctx, int table, int enc, int strict, int frames,
                                  const uint8_t *dna_dev, const uint64_t *offsets_dev,
                                  size_t n_seq, size_t total_nt, uint8_t *out_dev, void *stream);
/* n_masks_dev: device uint64 receiving the mask count (may be NULL).  Input without blanks and tabs whose output shares
 * the input's 16-byte phase (both 16-byte aligned, say) is finished by one streaming pass; anything else runs the scan and
 * the compacting pass behind it (three launches either way; the last two return at once when not needed). */
int qd_parse_dev(qd_ctx *ctx, int strict, const uint8_t *ascii_dev, size_t n, uint8_t *masks_dev,
                 uint64_t *n_masks_dev, void *stream);
int qd_canonical_dev(qd_ctx *ctx, int enc, int forward_only, const uint8_t *dna_dev, size_t n,
                     uint8_t *out_dev, void *stream);
int qd_canonical_windows_dev(qd_ctx *ctx, int enc, int forward_only, const uint8_t *dna_dev,
                             size_t n, size_t w, uint8_t *out_dev, void *stream);
/* the same over n_seq sequences of seq_len nucleotides each (benches/canonical.rs shape scaled up): windows never
 * span two sequences; out holds n_seq * (seq_len - w + 1) windows of w bytes, tightly packed */
int qd_canonical_windows_batch_dev(qd_ctx *ctx, int enc, int forward_only, const uint8_t *dna_dev,
                                   size_t n_seq, size_t seq_len, size_t w, uint8_t *out_dev, void *stream);
/* ---- canonical window KEYS ------------------------------------------------
 * One fixed-width key per canonical window instead of its w materialised bytes -- the form a downstream
 * set lookup (the consumer benches/all_windows.rs:15-102 was adapted from) wants, and 5x-42x fewer output
 * bytes.  The window arithmetic is exactly qd_canonical_windows' (src/canonical.rs:17-55, 71-118, 129-179).
 *   QD_CANON_KEY_PACKED  the canonical window itself, lossless, as two bit planes: PW = ceil(w / 32) uint32
 *                        words with the LOW bit of each nucleotide's code (A, T, C, G = 0, 1, 2, 3: the
 *                        reference's Ord; nucleotide j at bit j % 32 of word j / 32), then PW words with the
 *                        HIGH bits; unused bits zero.  8 bytes per window for w <= 32, else 16.
 *   QD_CANON_KEY_FP64    a 64-bit fingerprint of those words: h = 0x9e3779b97f4a7c15 * w; for k < PW:
 *                        h = mix64(h ^ (lo[k] | (uint64) hi[k] << 32)), mix64 = the splitmix64 finaliser
 *                        (z ^= z >> 30; z *= 0xbf58476d1ce4e5b9; z ^= z >> 27; z *= 0x94d049bb133111eb;
 *                        z ^= z >> 31).  Injective for w <= 32.
 * qd_canonical_key_bytes: bytes per key (0 = bad kind / w).  qd_canonical_fingerprint: the FP64 key of one
 * PACKED key, on the host (what the kernel computes).  Keys are in window order, tightly packed; keys_dev
 * 16-byte aligned.  Windows never span two sequences. */
enum { QD_CANON_KEY_PACKED = 0, QD_CANON_KEY_FP64 = 1 };
size_t qd_canonical_key_bytes(int kind, size_t w);
uint64_t qd_canonical_fingerprint(const uint32_t *packed, size_t w);
int qd_canonical_window_keys(qd_ctx *ctx, int enc, int forward_only, int kind, const uint8_t *dna, size_t n,
                             size_t w, void *keys, qd_err *err);
int qd_canonical_window_keys_dev(qd_ctx *ctx, int enc, int forward_only, int kind, const uint8_t *dna_dev,
                                 size_t n_seq, size_t seq_len, size_t w, void *keys_dev, void *stream);
/* expansions on device buffers (w <= 64, max_expansions < 128): counts_dev[n_win], out_index_dev[n_win + 1] (device
 * uint64; skipped windows repeat the next index), rows into out_dev while they fit out_capacity_rows
 * (out_index_dev[n_win] tells how many there are).  out_dev may be NULL: counts and indices only.
 * ONE kernel, one read of the windows: the row index is a chained scan over 32-window tiles inside the kernel that
 * writes the rows (rows still land in window order, whatever order the tiles finish in). */
int qd_expansion_windows_dev(qd_ctx *ctx, int enc, const uint8_t *windows_dev, size_t n_win, size_t w,
                             uint64_t max_expansions, uint64_t *counts_dev, uint64_t *out_index_dev,
                             uint8_t *out_dev, uint64_t out_capacity_rows, void *stream);
/* windows of a batch: n_seq sequences of seq_len bytes back to back; out_dev = [n_seq][seq_len - w + 1][w], tightly packed.
 * mode: what a window's bytes are -- */
enum {
    QD_WIN_COPY = 0,          /* the bytes as they are (ProteinSequence::windows, typed DnaSequence::windows)      */
    QD_WIN_DNA_ASCII = 1,     /* strict DNA letters -> upper-case letters (benches/all_windows.rs:84-86)           */
    QD_WIN_DNA_ASCII_RC = 2,  /* windows of the reverse complement (:87-92), never materialised: window i, byte j =
                                 complement(seq[len - 1 - i - j])                                                   */
    QD_WIN_MASK_ASCII = 3,    /* strict masks -> letters                                                            */
    QD_WIN_MASK_ASCII_RC = 4
};
int qd_windows_batch_dev(qd_ctx *ctx, const uint8_t *seq_dev, size_t n_seq, size_t seq_len, size_t w, int mode,
                         uint8_t *out_dev, void *stream);
/* benches/all_windows.rs:78-102 (SequenceWindows::from_dna, table Ncbi1) for a batch of strict sequences:
 *   dna_out_dev = [forward windows: n_seq x counts[0] x w_dna][reverse-complement windows: same size]
 *   aa_out_dev  = six blocks, block f (frame order fwd0..2, rc0..2) = [n_seq][counts[2+f]][w_aa]
 * counts come from qd_windows_all_batch_counts (returns their sum); the origin index of every window is the closed
 * form of benches/all_windows.rs:39-75 (see qd_windows_all).  Either output may be NULL. */
size_t qd_windows_all_batch_counts(size_t seq_len, size_t w_dna, size_t w_aa, size_t counts[8]);
int qd_windows_all_batch_dev(qd_ctx *ctx, int enc, const uint8_t *dna_dev, size_t n_seq, size_t seq_len, size_t w_dna,
                             size_t w_aa, uint8_t *dna_out_dev, uint8_t *aa_out_dev, void *stream);
int qd_windows_dev(qd_ctx *ctx, const uint8_t *seq_dev, size_t n, size_t w, uint8_t *out_dev,
                   void *stream);

/* ======================= several GPUs of one box, one call (host pointers, synchronous) =======================
 * north_star: "batches of sequences, or one long chromosome split at codon-aligned offsets with a halo, are partitioned
 * by byte count across the GPUs of one box, each on its own stream; outputs are concatenated on the host" -- the
 * caller loop over DnaSequence::translate_all_frames (src/rust_api.rs:263-270) / reverse_complement (:273-275).
 * ctxs: n_ctx DISTINCT contexts, normally one per device (two contexts of one device also work).  One host thread per
 * context drives that context's pipe streams; every device copies its results to the right offset of the ONE output
 * buffer; nothing is exchanged between devices.  Results and errors are exactly those of the single-context calls
 * (the first error in input order is the one reported).
 *   _batch_multi : offsets != NULL -> qd_translate_frames_batch (cut at the sequence boundaries nearest k*total/n_ctx);
 *                  offsets == NULL -> qd_translate_frames_batch_uniform with seq_len (cut by count)
 *   _multi       : ONE sequence, qd_translate_frames: cut at multiples of 48, each piece read with a 2-nt halo and
 *                  translated as a stand-alone sequence; forward local frame k = global frame k from residue start/3,
 *                  reverse local frame k = global reverse frame (k + d) % 3 from residue (k + d) / 3, d = nucleotides
 *                  to the right of the piece
 *   qd_revcomp_multi : piece [a, b) lands at out[n - b, n - a)
 * A single context already pipelines sequences of 8 MiB and more this way (chunks of n/16, 4 .. 64 MiB). */
int qd_translate_frames_batch_multi(qd_ctx *const *ctxs, int n_ctx, int table, int enc, int strict, int frames, const uint8_t *dna,
                                    const uint64_t *offsets, size_t n_seq, size_t seq_len, uint8_t *out, uint64_t *out_offsets,
                                    qd_err *err);
int qd_translate_frames_multi(qd_ctx *const *ctxs, int n_ctx, int table, int enc, int strict, int frames, const uint8_t *dna, size_t n,
                              uint8_t *out, size_t out_len[6], int *n_frames, qd_err *err);
int qd_revcomp_multi(qd_ctx *const *ctxs, int n_ctx, int enc, int strict, const uint8_t *dna, size_t n, uint8_t *out, qd_err *err);

/* ---- checksum of a device buffer ---------------------------------------------------------------------------------
 * Position-weighted LINEAR checksum: the buffer holds bytes [base, base + n) of a byte stream; stream word J (bytes
 * 8J .. 8J+7, little endian) weighs K(J) = mix64(J * 0x9e3779b97f4a7c15 + 1) | 1 (mix64 = the splitmix64 finaliser above,
 * qd_checksum64_weight returns K(J)); the checksum is sum_J word_J * K(J) mod 2^64 and is ADDED to *acc_dev.  Linear in
 * the data: the chunks of an output ring, or the shards the GPUs of a box produced, add up to the checksum of the whole
 * output whatever their byte alignment -- so a 336 GB window enumeration is checked without ever being resident.
 * One HBM read of the buffer.  Any alignment of buf_dev. */
int qd_checksum64_dev(qd_ctx *ctx, const uint8_t *buf_dev, size_t n, uint64_t base, uint64_t *acc_dev, void *stream);
uint64_t qd_checksum64_weight(uint64_t word_index);

/* ---- chunked ("ring") window enumeration ---------------------------------------------------------------------------
 * The window kernels write 42 - 124 bytes per input byte (benches/canonical.rs:29-34, benches/all_windows.rs:78-102,
 * benches/expansions.rs:126-130): the output of an 8 GB shard (336 GB of canonical rows) does not fit HBM.  These entry
 * points walk the batch in chunks of `chunk_seqs` sequences (`chunk_windows` windows), chunk k written to slot
 * k % n_slots of the caller's ring (n_slots = ring_bytes / slot size, slot size = the chunk's bytes rounded up to 256;
 * sizes from the qd_*_chunk_bytes helpers).  After each chunk, on the same stream:
 *   - checksums_dev[k] (optional, qd_chunk_count() entries) receives the chunk's qd_checksum64:
 *       canonical   : base = first window of the chunk * w  (summing all chunks gives the checksum of the whole
 *                     [n_seq][windows][w] output, whatever the chunking)
 *       windows_all : chunk-relative -- the chunk's DNA windows [forward block][reverse block] at base 0, then its
 *                     amino-acid windows (six frame blocks) at base = the DNA bytes of the chunk
 *       expansions  : chunk-relative, base 0, over the chunk_rows_dev[k] rows the chunk produced
 *   - `consume` (optional) is called on the host with the slot; whatever it enqueues on `stream` runs before the slot
 *     is overwritten.  n_rows_dev is NULL unless the row count is only known on the device (expansions).
 * Slot layout: canonical = qd_canonical_windows_batch_dev of the chunk; windows_all = dna_out at 0, aa_out at
 * *aa_offset (qd_windows_all_chunk_bytes); expansions = rows, tightly packed. */
typedef void (*qd_chunk_fn)(void *user, size_t chunk, size_t first_unit, size_t n_units, const uint8_t *slot_dev, size_t slot_bytes,
                            const uint64_t *n_rows_dev, void *stream);
size_t qd_chunk_count(size_t n_units, size_t chunk_units);
size_t qd_canonical_chunk_bytes(size_t chunk_seqs, size_t seq_len, size_t w);
size_t qd_windows_all_chunk_bytes(size_t chunk_seqs, size_t seq_len, size_t w_dna, size_t w_aa, size_t *aa_offset);
size_t qd_expansion_chunk_bytes(size_t chunk_windows, size_t w, uint64_t max_expansions);
int qd_canonical_windows_chunked_dev(qd_ctx *ctx, int enc, int forward_only, const uint8_t *dna_dev, size_t n_seq, size_t seq_len,
                                     size_t w, uint8_t *ring_dev, size_t ring_bytes, size_t chunk_seqs, uint64_t *checksums_dev,
                                     qd_chunk_fn consume, void *user, void *stream);
int qd_windows_all_chunked_dev(qd_ctx *ctx, int enc, const uint8_t *dna_dev, size_t n_seq, size_t seq_len, size_t w_dna, size_t w_aa,
                               uint8_t *ring_dev, size_t ring_bytes, size_t chunk_seqs, uint64_t *checksums_dev, qd_chunk_fn consume,
                               void *user, void *stream);
/* counts_dev: size_hint of every window (n_win entries) or NULL; chunk_rows_dev[k] = rows chunk k produced (required) */
int qd_expansion_windows_chunked_dev(qd_ctx *ctx, int enc, const uint8_t *windows_dev, size_t n_win, size_t w, uint64_t max_expansions,
                                     uint8_t *ring_dev, size_t ring_bytes, size_t chunk_windows, uint64_t *counts_dev,
                                     uint64_t *chunk_rows_dev, uint64_t *checksums_dev, qd_chunk_fn consume, void *user, void *stream);

/* Number of kernel launches this context has issued (bench.py's gpu_launches claim). */
uint64_t qd_launch_count(const qd_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* QUICKDNA_B200_H */

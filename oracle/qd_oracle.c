/* quickdna hot-path ORACLE -- see qd_oracle.h.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain C, single-threaded scalar loops shaped like the reference's (so that it can also
 * serve as the timed "reference CPU path").  Paths cited are relative to /root/reference/.
 */
#define _POSIX_C_SOURCE 200809L
#include "qd_oracle.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "ncbi_codes.h"

/* ---------------------------------------------------------------- alphabet -------- */

/* src/nucleotide.rs:20-28, 31-53: the in-memory value IS the 4-bit IUPAC set. */
enum { M_A = 1, M_T = 2, M_C = 4, M_G = 8 };

/* src/nucleotide.rs:66-101: ASCII -> mask, 0 = None, as the reference's 256-entry ASCII_TO_NUCLEOTIDE table
 * (its PERF comment: 256 entries so that any u8 indexes it without a bounds check). */
#define BOTH_CASES(ch, m) [ch] = (m), [(ch) + 32] = (m)
static const uint8_t ASCII_TO_MASK[256] = {
    BOTH_CASES('A', M_A), BOTH_CASES('T', M_T), BOTH_CASES('C', M_C), BOTH_CASES('G', M_G),
    BOTH_CASES('N', M_A | M_T | M_C | M_G), BOTH_CASES('M', M_A | M_C), BOTH_CASES('R', M_A | M_G),
    BOTH_CASES('W', M_A | M_T), BOTH_CASES('S', M_C | M_G), BOTH_CASES('Y', M_T | M_C),
    BOTH_CASES('K', M_T | M_G), BOTH_CASES('V', M_A | M_C | M_G), BOTH_CASES('H', M_A | M_T | M_C),
    BOTH_CASES('D', M_A | M_T | M_G), BOTH_CASES('B', M_T | M_C | M_G),
};
#undef BOTH_CASES

/* src/nucleotide.rs:138-145, 219-237: always upper case. */
static const uint8_t MASK_TO_ASCII[16] = {'?', 'A', 'T', 'W', 'C', 'M', 'Y', 'H', 'G', 'R', 'K', 'D', 'S', 'V', 'B', 'N'};
static inline uint8_t mask_to_ascii(uint8_t m) { return MASK_TO_ASCII[m & 15]; }
uint8_t qdo_mask_to_ascii(uint8_t m) { return mask_to_ascii(m); }

/* src/nucleotide.rs:125-132, 195-213: the reference's match (A<->T, C<->G, M<->K, R<->Y, H<->D, V<->B; W, S, N fixed)
 * as the table a compiler makes of it; 0 for a byte that is no mask. */
static const uint8_t COMPLEMENT[16] = {0, 2, 1, 3, 8, 10, 9, 11, 4, 6, 5, 7, 12, 14, 13, 15};
static inline uint8_t complement_mask(uint8_t m) { return m < 16 ? COMPLEMENT[m] : 0; }
uint8_t qdo_complement_mask(uint8_t m) { return complement_mask(m); }

static inline int popcount4(uint8_t m) { return (m & 1) + ((m >> 1) & 1) + ((m >> 2) & 1) + ((m >> 3) & 1); }

/* src/nucleotide.rs:285-299 (strict) / :301-315 (ambiguous) / :268-283 (ambiguous -> strict). */
static inline int decode_byte(uint8_t b, int strict, uint8_t *mask, uint8_t *payload) {
    if (b >= 128) {
        *payload = b;
        return QDO_NON_ASCII_BYTE;
    }
    uint8_t m = ASCII_TO_MASK[b];
    if (m == 0) {
        *payload = b; /* BadNucleotide(u.into()): the char as given, case preserved */
        return QDO_BAD_NUCLEOTIDE;
    }
    if (strict && (m & (m - 1))) {
        *payload = mask_to_ascii(m); /* other.into(): upper-case letter of the code */
        return QDO_UNEXPECTED_AMBIGUOUS;
    }
    *mask = m;
    return QDO_OK;
}
int qdo_decode_byte(uint8_t b, int strict, uint8_t *mask, uint8_t *payload) { return decode_byte(b, strict, mask, payload); }

/* Rust `{:?}` of a char (core::char::EscapeDebug) for ASCII inputs. */
static int fmt_char_debug(uint8_t c, char *buf, size_t cap) {
    switch (c) {
        case '\0': return snprintf(buf, cap, "'\\0'");
        case '\t': return snprintf(buf, cap, "'\\t'");
        case '\r': return snprintf(buf, cap, "'\\r'");
        case '\n': return snprintf(buf, cap, "'\\n'");
        case '\'': return snprintf(buf, cap, "'\\''");
        case '\\': return snprintf(buf, cap, "'\\\\'");
        default:
            if (c < 0x20 || c == 0x7f) return snprintf(buf, cap, "'\\u{%x}'", c);
            return snprintf(buf, cap, "'%c'", c);
    }
}

/* src/errors.rs:19-28 */
int qdo_err_message(const qdo_err *e, char *buf, size_t cap) {
    char ch[16];
    switch (e->kind) {
        case QDO_NON_ASCII_BYTE: return snprintf(buf, cap, "non-ascii byte: %x", e->byte);
        case QDO_BAD_NUCLEOTIDE:
            fmt_char_debug(e->byte, ch, sizeof ch);
            return snprintf(buf, cap, "bad nucleotide: %s", ch);
        case QDO_UNEXPECTED_AMBIGUOUS:
            fmt_char_debug(e->byte, ch, sizeof ch);
            return snprintf(buf, cap, "unexpected ambiguous nucleotide: %s", ch);
        case QDO_BAD_TABLE: return snprintf(buf, cap, "not a ncbi translation table: %u", e->byte);
        default: return snprintf(buf, cap, "ok");
    }
}

/* ---------------------------------------------------------------- tables ---------- */

static uint8_t g_tables[QDO_N_TRANS_TABLES * QDO_CODONS_PER_TABLE];
static pthread_once_t g_tables_once = PTHREAD_ONCE_INIT;

/* src/trans_table.rs:78-85 */
static inline size_t codon_idx(uint8_t a, uint8_t b, uint8_t c) {
    return ((size_t)a << 8) | ((size_t)b << 4) | (size_t)c;
}

/* src/nucleotide.rs:171-189: possibilities() in A,T,C,G order == ascending mask bit. */
static int possibilities(uint8_t m, uint8_t out[4]) {
    int k = 0;
    for (int bit = 1; bit <= 8; bit <<= 1)
        if (m & bit) out[k++] = (uint8_t)bit;
    return k;
}

/* src/bin/gen_table.rs:83-115 */
static uint8_t ambiguous_codon_protein(uint8_t a, uint8_t b, uint8_t c, const uint8_t *table) {
    uint8_t seen[3];
    int n_seen = 0;
    uint8_t pa[4], pb[4], pc[4];
    int na = possibilities(a, pa), nb = possibilities(b, pb), nc = possibilities(c, pc);
    for (int i = 0; i < na; i++)
        for (int j = 0; j < nb; j++)
            for (int k = 0; k < nc; k++) {
                uint8_t aa = table[codon_idx(pa[i], pb[j], pc[k])];
                int known = 0;
                for (int s = 0; s < n_seen; s++) known |= (seen[s] == aa);
                if (!known) {
                    seen[n_seen++] = aa;
                    if (n_seen >= 3) return 'X';
                }
            }
    if (n_seen == 1) return seen[0];
    uint8_t lo = seen[0] < seen[1] ? seen[0] : seen[1];
    uint8_t hi = seen[0] < seen[1] ? seen[1] : seen[0];
    if (lo == 'D' && hi == 'N') return 'B';
    if (lo == 'E' && hi == 'Q') return 'Z';
    if (lo == 'I' && hi == 'L') return 'J';
    return 'X';
}

/* src/bin/gen_table.rs:117-157 */
static void gen_tables(void) {
    static const uint8_t TCAG[4] = {M_T, M_C, M_A, M_G};
    memset(g_tables, '*', sizeof g_tables);
    for (int slot = 0; slot < QDO_N_TRANS_TABLES; slot++) {
        uint8_t *t = g_tables + (size_t)slot * QDO_CODONS_PER_TABLE;
        const char *aas = QDO_SLOT_AAS[slot];
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++)
                for (int k = 0; k < 4; k++)
                    t[codon_idx(TCAG[i], TCAG[j], TCAG[k])] = (uint8_t)aas[16 * i + 4 * j + k];
        for (uint8_t a = 1; a < 16; a++)
            for (uint8_t b = 1; b < 16; b++)
                for (uint8_t c = 1; c < 16; c++)
                    if (popcount4(a) > 1 || popcount4(b) > 1 || popcount4(c) > 1)
                        t[codon_idx(a, b, c)] = ambiguous_codon_protein(a, b, c, t);
    }
}

const uint8_t *qdo_tables(void) {
    pthread_once(&g_tables_once, gen_tables);
    return g_tables;
}

/* src/trans_table.rs:230-267 (valid ids) and :114-146 (slot; 8 -> slot of 1, 7 -> slot of 4). */
int qdo_table_slot(int id) {
    if (id == 8) return 0;
    if (id == 7) return 3;
    for (int s = 0; s < QDO_N_TRANS_TABLES; s++)
        if (QDO_SLOT_NCBI_ID[s] == id) return s;
    return -1;
}

/* ---------------------------------------------------------------- translate ------- */

/* src/trans_table.rs:205-227 */
void qdo_translate_masks(int slot, const uint8_t *m, size_t n, uint8_t *out) {
    const uint8_t *t = qdo_tables() + (size_t)slot * QDO_CODONS_PER_TABLE;
    size_t k = 0;
    for (size_t i = 0; i + 3 <= n; i += 3) out[k++] = t[codon_idx(m[i], m[i + 1], m[i + 2])];
}

/* src/python_api.rs:33-51 then src/trans_table.rs:176-203 */
int qdo_translate_bytes(int ncbi_id, int strict, const uint8_t *dna, size_t n, uint8_t *out,
                        qdo_err *err) {
    int slot = (ncbi_id < 0 || ncbi_id > 255) ? -1 : qdo_table_slot(ncbi_id);
    if (slot < 0) {
        err->kind = QDO_BAD_TABLE;
        err->byte = (uint8_t)ncbi_id;
        err->pos = 0;
        return QDO_BAD_TABLE;
    }
    const uint8_t *t = qdo_tables() + (size_t)slot * QDO_CODONS_PER_TABLE;
    size_t k = 0;
    for (size_t i = 0; i + 3 <= n; i += 3) { /* chunks_exact(3): the tail is never looked at */
        uint8_t m[3];
        for (int j = 0; j < 3; j++) {
            uint8_t payload = 0;
            int rc = decode_byte(dna[i + j], strict, &m[j], &payload);
            if (rc != QDO_OK) {
                err->kind = rc;
                err->byte = payload;
                err->pos = i + j;
                return rc;
            }
        }
        out[k++] = t[codon_idx(m[0], m[1], m[2])];
    }
    err->kind = QDO_OK;
    return QDO_OK;
}

/* src/trans_table.rs:269-278 */
int qdo_revcomp_bytes(int strict, const uint8_t *dna, size_t n, uint8_t *out, qdo_err *err) {
    memset(out, 0, n); /* vec![0u8; n] */
    for (size_t i = 0; i < n; i++) {
        uint8_t m = 0, payload = 0;
        int rc = decode_byte(dna[i], strict, &m, &payload);
        if (rc != QDO_OK) {
            err->kind = rc;
            err->byte = payload;
            err->pos = i;
            return rc;
        }
        out[n - 1 - i] = mask_to_ascii(complement_mask(m));
    }
    err->kind = QDO_OK;
    return QDO_OK;
}

/* src/trans_table.rs:284-286 */
void qdo_revcomp_masks(const uint8_t *m, size_t n, uint8_t *out) {
    for (size_t i = 0; i < n; i++) out[i] = complement_mask(m[n - 1 - i]);
}

/* src/rust_api.rs:227-255 */
int qdo_self_frames_masks(int slot, const uint8_t *m, size_t n, uint8_t *out, size_t lens[3]) {
    int frames = n >= 5 ? 3 : n == 4 ? 2 : n == 3 ? 1 : 0;
    lens[0] = lens[1] = lens[2] = 0;
    for (int k = 0; k < frames; k++) {
        qdo_translate_masks(slot, m + k, n - k, out);
        lens[k] = (n - k) / 3;
        out += lens[k];
    }
    return frames;
}

/* src/rust_api.rs:263-270: self frames, then self frames of the MATERIALISED reverse complement */
int qdo_all_frames_masks(int slot, const uint8_t *m, size_t n, uint8_t *out, size_t lens[6]) {
    int f = qdo_self_frames_masks(slot, m, n, out, lens);
    size_t used = lens[0] + lens[1] + lens[2];
    uint8_t *rc = (uint8_t *)malloc(n ? n : 1);
    qdo_revcomp_masks(m, n, rc);
    int r = qdo_self_frames_masks(slot, rc, n, out + used, lens + 3);
    free(rc);
    return f + r;
}

/* quickdna/__init__.py:128-161 over python_api.rs:33-51 */
int qdo_py_self_frames(int ncbi_id, int strict, const uint8_t *dna, size_t n, uint8_t *out,
                       size_t lens[3], int *n_frames, qdo_err *err) {
    int frames = n >= 5 ? 3 : n == 4 ? 2 : n == 3 ? 1 : 0;
    lens[0] = lens[1] = lens[2] = 0;
    *n_frames = 0;
    for (int k = 0; k < frames; k++) {
        int rc = qdo_translate_bytes(ncbi_id, strict, dna + k, n - k, out, err);
        if (rc != QDO_OK) {
            if (rc != QDO_BAD_TABLE) err->pos += (uint64_t)k;
            return rc;
        }
        lens[k] = (n - k) / 3;
        out += lens[k];
    }
    *n_frames = frames;
    err->kind = QDO_OK;
    return QDO_OK;
}

/* quickdna/__init__.py:163-185: note reverse_complement() is called WITHOUT strict (:182) */
int qdo_py_all_frames(int ncbi_id, int strict, const uint8_t *dna, size_t n, uint8_t *out,
                      size_t lens[6], int *n_frames, qdo_err *err) {
    int nf = 0, nr = 0;
    *n_frames = 0;
    lens[3] = lens[4] = lens[5] = 0;
    int rc = qdo_py_self_frames(ncbi_id, strict, dna, n, out, lens, &nf, err);
    if (rc != QDO_OK) return rc;
    size_t used = lens[0] + lens[1] + lens[2];
    uint8_t *rcseq = (uint8_t *)malloc(n ? n : 1);
    rc = qdo_revcomp_bytes(0, dna, n, rcseq, err);
    if (rc == QDO_OK) {
        rc = qdo_py_self_frames(ncbi_id, strict, rcseq, n, out + used, lens + 3, &nr, err);
        if (rc != QDO_OK && rc != QDO_BAD_TABLE) err->pos = (uint64_t)n - 1 - err->pos; /* index in dna */
    }
    free(rcseq);
    if (rc != QDO_OK) return rc;
    *n_frames = nf + nr;
    return QDO_OK;
}

/* src/rust_api.rs:310-322 */
int qdo_parse(int strict, const uint8_t *ascii, size_t n, uint8_t *masks, size_t *n_masks,
              qdo_err *err) {
    size_t k = 0;
    for (size_t i = 0; i < n; i++) {
        uint8_t b = ascii[i];
        if (b != ' ' && b != '\t') {
            uint8_t m = 0, payload = 0;
            int rc = decode_byte(b, strict, &m, &payload);
            if (rc != QDO_OK) {
                err->kind = rc;
                err->byte = payload;
                err->pos = i;
                return rc;
            }
            masks[k++] = m;
        }
    }
    *n_masks = k;
    err->kind = QDO_OK;
    return QDO_OK;
}

/* ---------------------------------------------------------------- FASTA ----------- */

/* src/fasta.rs:207-415.  The three parser states carry: the record's first line, where its header text starts and
 * ends in `text`, and where its content starts in `masks`. */
enum { FA_START_OF_FILE, FA_IN_HEADER, FA_IN_RECORD };

int qdo_fasta_scan(int strict, int concatenate_headers, int allow_preceding_comment, const uint8_t *text, size_t n,
                   uint8_t *masks, size_t *n_masks, qdo_fasta_record *records, size_t max_records, size_t *n_records,
                   qdo_err *err, uint64_t *err_line) {
    size_t k = 0, n_rec = 0;
    int state = FA_START_OF_FILE;
    uint64_t start_line = 1, header_begin = UINT64_MAX, header_end = UINT64_MAX, content_begin = 0;
    uint64_t line_number = 0;
    size_t pos = 0;
    err->kind = QDO_OK;
#define FA_EMIT(line_end_)                                                                        \
    do {                                                                                          \
        if (n_rec < max_records) {                                                                \
            qdo_fasta_record r = {header_begin, header_end, content_begin, k, start_line, (line_end_)}; \
            records[n_rec] = r;                                                                   \
        }                                                                                         \
        n_rec++;                                                                                  \
    } while (0)
    while (pos < n) { /* BufRead::lines: one line per '\n'; a final line without '\n' still counts */
        size_t eol = pos;
        while (eol < n && text[eol] != '\n') eol++;
        size_t line_end = eol; /* exclusive, before the line break */
        if (eol < n && line_end > pos && text[line_end - 1] == '\r') line_end--; /* CRLF */
        line_number++;
        const int is_header = line_end > pos && (text[pos] == '>' || text[pos] == ';'); /* :482-491 */
        if (is_header) {
            if (state == FA_START_OF_FILE) { /* :240-262 */
                if (!allow_preceding_comment && k > 0) FA_EMIT(line_number);
                content_begin = k;
                start_line = line_number, header_begin = pos, header_end = line_end, state = FA_IN_HEADER;
            } else if (state == FA_IN_HEADER) { /* :276-304 */
                if (concatenate_headers) {
                    header_end = line_end;
                } else {
                    FA_EMIT(line_number);
                    start_line = line_number, header_begin = pos, header_end = line_end;
                }
            } else { /* InRecord: :322-340 */
                FA_EMIT(line_number);
                content_begin = k;
                start_line = line_number, header_begin = pos, header_end = line_end, state = FA_IN_HEADER;
            }
        } else {
            /* StartOfFile keeps content only if it may be emitted (:265-272); InHeader -> InRecord (:306-320);
             * InRecord extends (:342-358) */
            if (!(state == FA_START_OF_FILE && allow_preceding_comment)) {
                for (size_t i = pos; i < line_end; i++) {
                    const uint8_t b = text[i];
                    if (b == ' ' || b == '\t') continue; /* src/rust_api.rs:316 */
                    uint8_t m = 0, payload = 0;
                    const int rc = decode_byte(b, strict, &m, &payload);
                    if (rc != QDO_OK) {
                        err->kind = rc;
                        err->byte = payload;
                        err->pos = i;
                        *err_line = line_number;
                        return rc;
                    }
                    masks[k++] = m;
                }
            }
            if (state == FA_IN_HEADER) state = FA_IN_RECORD;
        }
        pos = eol < n ? eol + 1 : n;
    }
    /* advance_eof, :372-411 */
    if (state == FA_START_OF_FILE) {
        if (!allow_preceding_comment && k > 0) FA_EMIT(line_number + 1);
    } else {
        FA_EMIT(line_number + 1);
    }
#undef FA_EMIT
    *n_masks = k;
    *n_records = n_rec;
    return QDO_OK;
}

/* ---------------------------------------------------------------- canonical ------- */

/* src/canonical.rs:95-118: streaming first-occurrence relabelling. `step` = +1 / -1 walks the
 * sequence forwards / backwards (ForwardCanonical over iter.rev()). */
typedef struct {
    const uint8_t *p;
    ptrdiff_t step;
    size_t left;
    uint8_t perm[4]; /* 0 = None */
    int next_unmapped;
} fwd_canon_iter;

static const uint8_t NUC_ALL[4] = {M_A, M_T, M_C, M_G}; /* src/nucleotide.rs:104 */

static void fci_init(fwd_canon_iter *it, const uint8_t *first, ptrdiff_t step, size_t n) {
    it->p = first;
    it->step = step;
    it->left = n;
    memset(it->perm, 0, sizeof it->perm);
    it->next_unmapped = 0;
}

/* returns 0 when exhausted (None), else the relabelled nucleotide mask */
static uint8_t fci_next(fwd_canon_iter *it) {
    if (it->left == 0) return 0;
    uint8_t nuc = *it->p;
    it->p += it->step;
    it->left--;
    int idx = nuc == M_A ? 0 : nuc == M_T ? 1 : nuc == M_C ? 2 : 3; /* canonical.rs:101-106 */
    if (it->perm[idx] == 0) /* get_or_insert_with(|| unmapped.next().unwrap_or(A)) */
        it->perm[idx] = it->next_unmapped < 4 ? NUC_ALL[it->next_unmapped++] : M_A;
    return it->perm[idx];
}

void qdo_forward_canonical(const uint8_t *m, size_t n, uint8_t *out) {
    fwd_canon_iter it;
    fci_init(&it, m, 1, n);
    for (size_t i = 0; i < n; i++) out[i] = fci_next(&it);
}

/* src/canonical.rs:25-33 + LexicalMin :146-170.  Ord on Nucleotide is the discriminant
 * (A=1 < T=2 < C=4 < G=8, src/nucleotide.rs:20-28), and None < Some(_). */
void qdo_canonical(const uint8_t *m, size_t n, uint8_t *out) {
    fwd_canon_iter fw, rev;
    fci_init(&fw, m, 1, n);
    fci_init(&rev, n ? m + n - 1 : m, -1, n);
    int order = 0; /* Ordering::Equal */
    for (size_t i = 0; i < n; i++) {
        uint8_t v;
        if (order < 0) {
            v = fci_next(&fw);
        } else if (order > 0) {
            v = fci_next(&rev);
        } else {
            uint8_t a = fci_next(&fw), b = fci_next(&rev);
            order = (a < b) ? -1 : (a > b) ? 1 : 0;
            v = order < 0 ? a : b;
        }
        out[i] = v;
    }
}

size_t qdo_canonical_windows(const uint8_t *m, size_t n, size_t w, uint8_t *out) {
    if (w == 0 || n < w) return 0;
    size_t count = n - w + 1;
    for (size_t i = 0; i < count; i++) qdo_canonical(m + i, w, out + i * w);
    return count;
}

/* ---------------------------------------------------------------- expansions ------ */

/* src/expansions.rs:96-107 with began=false and all digits 0 */
int qdo_expansions_size_hint(const uint8_t *m, size_t n, uint64_t *count) {
    uint64_t size = 0;
    for (size_t i = 0; i < n; i++) {
        uint64_t states = (uint64_t)popcount4(m[i]);
        if (states <= 1) continue; /* only ambiguities are tracked, :39-47 */
        uint64_t remaining = states - 0 - 1;
        if (size != 0 && states > UINT64_MAX / size) return 0; /* checked_mul */
        size *= states;
        if (size > UINT64_MAX - remaining) return 0; /* checked_add */
        size += remaining;
    }
    if (size == UINT64_MAX) return 0; /* checked_add(!began) */
    *count = size + 1;
    return 1;
}

/* src/expansions.rs:37-94 */
size_t qdo_expansions(const uint8_t *m, size_t n, uint8_t *out, size_t max_out) {
    size_t n_amb = 0;
    size_t *amb_index = (size_t *)malloc((n ? n : 1) * sizeof(size_t));
    uint8_t *digit = (uint8_t *)calloc(n ? n : 1, 1);
    uint8_t *buf = (uint8_t *)malloc(n ? n : 1);
    for (size_t i = 0; i < n; i++) {
        uint8_t poss[4];
        possibilities(m[i], poss);
        buf[i] = poss[0]; /* first possibility of every position, :48-51 */
        if (popcount4(m[i]) > 1) amb_index[n_amb++] = i;
    }
    size_t produced = 0;
    int began = 0;
    for (;;) {
        if (!began) {
            began = 1;
        } else {
            int advanced = 0;
            for (size_t a = n_amb; a-- > 0;) { /* ambiguities.iter_mut().rev() */
                size_t idx = amb_index[a];
                uint8_t poss[4];
                int states = possibilities(m[idx], poss);
                digit[a] = (uint8_t)((digit[a] + 1) % states);
                buf[idx] = poss[digit[a]];
                if (digit[a] > 0) {
                    advanced = 1;
                    break;
                }
            }
            if (!advanced) break;
        }
        if (produced >= max_out) break;
        memcpy(out + produced * n, buf, n);
        produced++;
    }
    free(amb_index);
    free(digit);
    free(buf);
    return produced;
}

/* ---------------------------------------------------------------- windows --------- */

/* src/rust_api.rs:100-104, 277-279 (slice::windows: none if n < w; w == 0 panics there) */
size_t qdo_windows(const uint8_t *seq, size_t n, size_t w, uint8_t *out) {
    if (w == 0 || n < w) return 0;
    size_t count = n - w + 1;
    if (out)
        for (size_t i = 0; i < count; i++) memcpy(out + i * w, seq + i, w);
    return count;
}

/* benches/all_windows.rs:78-102 (from_dna, table Ncbi1) and :39-75 (enumerate) */
size_t qdo_windows_all(const uint8_t *m, size_t n, size_t w_dna, size_t w_aa, uint8_t *dna_out,
                       uint8_t *aa_out, uint64_t *origin, size_t counts[8]) {
    uint8_t *ascii = (uint8_t *)malloc(n ? n : 1);
    uint8_t *rc = (uint8_t *)malloc(n ? n : 1);
    uint8_t *rc_ascii = (uint8_t *)malloc(n ? n : 1);
    uint8_t *aa = (uint8_t *)malloc(2 * n + 16);
    size_t lens[6];
    for (size_t i = 0; i < n; i++) ascii[i] = mask_to_ascii(m[i]);
    qdo_revcomp_masks(m, n, rc);
    for (size_t i = 0; i < n; i++) rc_ascii[i] = mask_to_ascii(rc[i]);
    int n_frames = qdo_all_frames_masks(0 /* Ncbi1 */, m, n, aa, lens);

    size_t total = 0;
    counts[0] = qdo_windows(ascii, n, w_dna, dna_out);
    for (size_t i = 0; i < counts[0]; i++)
        if (origin) origin[total + i] = i;
    total += counts[0];
    counts[1] = qdo_windows(rc_ascii, n, w_dna, dna_out ? dna_out + counts[0] * w_dna : NULL);
    for (size_t i = 0; i < counts[1]; i++)
        if (origin) origin[total + i] = counts[1] - i - 1;
    total += counts[1];

    /* translations.iter(): only the frames the reference returned, `operation` = list index.
     * The returned list drops empty frames, so compact lens[] the same way. */
    size_t list_len[6], list_off[6];
    int k = 0;
    size_t off = 0;
    for (int f = 0; f < 6; f++) {
        if (lens[f] > 0) {
            list_len[k] = lens[f];
            list_off[k] = off;
            k++;
        }
        off += lens[f];
    }
    (void)n_frames;
    size_t aa_written = 0;
    for (int op = 0; op < 6; op++) {
        counts[2 + op] = 0;
        if (op >= k) continue;
        size_t c = qdo_windows(aa + list_off[op], list_len[op], w_aa,
                               aa_out ? aa_out + aa_written * w_aa : NULL);
        counts[2 + op] = c;
        for (size_t j = 0; j < c; j++) {
            uint64_t idx;
            if (op <= 2)
                idx = 3 * j + (size_t)op;
            else
                idx = n - ((j + w_aa) * 3 + (size_t)(op % 3));
            if (origin) origin[total + j] = idx;
        }
        aa_written += c;
        total += c;
    }
    free(ascii);
    free(rc);
    free(rc_ascii);
    free(aa);
    return total;
}

/* ---------------------------------------------------------------- thread pool ----- */

/* One set of threads per call: job k of n_jobs goes to thread k % threads (jobs are equal-sized by construction). */
typedef void (*job_fn)(void *ctx, size_t job);
typedef struct {
    job_fn fn;
    void *ctx;
    size_t first, n_jobs, stride;
} pool_arg;

static void *pool_thread(void *a) {
    pool_arg *p = (pool_arg *)a;
    for (size_t j = p->first; j < p->n_jobs; j += p->stride) p->fn(p->ctx, j);
    return NULL;
}

static void run_jobs(job_fn fn, void *ctx, size_t n_jobs, int threads) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > n_jobs) threads = n_jobs ? (int)n_jobs : 1;
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    pool_arg *arg = (pool_arg *)malloc(sizeof(pool_arg) * (size_t)threads);
    for (int t = 0; t < threads; t++) {
        arg[t].fn = fn, arg[t].ctx = ctx, arg[t].first = (size_t)t, arg[t].n_jobs = n_jobs, arg[t].stride = (size_t)threads;
        if (t) pthread_create(&tid[t], NULL, pool_thread, &arg[t]);
    }
    pool_thread(&arg[0]);
    for (int t = 1; t < threads; t++) pthread_join(tid[t], NULL);
    free(tid);
    free(arg);
}

/* ---------------------------------------------------------------- timed baseline -- */

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

typedef struct {
    int slot, threads, reps;
    const uint8_t *dna;
    size_t n_reads, read_len;
    uint8_t *out;
} frames_job;

/* one thread's share of the batch, `reps` times over: per read the parse of rust_api.rs:310-322 (validation lives
 * there on the Rust path), then translate_all_frames (:263-270).  Scratch is per thread, not per read (the reference
 * allocates seven Vecs per read; none of that is charged here). */
static void frames_worker(void *ctx, size_t t) {
    const frames_job *j = (const frames_job *)ctx;
    const size_t L = j->read_len, out_per_read = L >= 2 ? 2 * (L - 2) : 0;
    const size_t r0 = j->n_reads * t / (size_t)j->threads, r1 = j->n_reads * (t + 1) / (size_t)j->threads;
    const uint8_t *table = qdo_tables() + (size_t)j->slot * QDO_CODONS_PER_TABLE;
    uint8_t *masks = (uint8_t *)malloc(2 * (L ? L : 1)), *rc = masks + (L ? L : 1);
    for (int rep = 0; rep < j->reps; rep++)
        for (size_t r = r0; r < r1; r++) {
            qdo_err e;
            size_t nm = 0;
            if (qdo_parse(0, j->dna + r * L, L, masks, &nm, &e) != QDO_OK) continue;
            uint8_t *o = j->out + r * out_per_read;
            const int frames = nm >= 5 ? 3 : nm == 4 ? 2 : nm == 3 ? 1 : 0;
            for (int k = 0; k < frames; k++)
                for (size_t i = (size_t)k; i + 3 <= nm; i += 3) *o++ = table[codon_idx(masks[i], masks[i + 1], masks[i + 2])];
            qdo_revcomp_masks(masks, nm, rc);
            for (int k = 0; k < frames; k++)
                for (size_t i = (size_t)k; i + 3 <= nm; i += 3) *o++ = table[codon_idx(rc[i], rc[i + 1], rc[i + 2])];
        }
    free(masks);
}

double qdo_bench_all_frames_batch(int slot, const uint8_t *dna, size_t n_reads, size_t read_len,
                                  uint8_t *out, int threads, int reps) {
    if (threads < 1) threads = 1;
    qdo_tables();
    frames_job j = {slot, threads, reps, dna, n_reads, read_len, out};
    const double t0 = now_s();
    run_jobs(frames_worker, &j, (size_t)threads, threads);
    return now_s() - t0;
}

typedef struct {
    const uint8_t *dna;
    size_t n;
    uint8_t *out;
    int threads, reps;
} rc_job;

/* one thread's share of the loop at src/trans_table.rs:273-276 (zero-filled Vec, then the reversed indexed store) */
static void rc_worker(void *ctx, size_t t) {
    const rc_job *j = (const rc_job *)ctx;
    const size_t i0 = j->n * t / (size_t)j->threads, i1 = j->n * (t + 1) / (size_t)j->threads;
    for (int rep = 0; rep < j->reps; rep++) {
        memset(j->out + (j->n - i1), 0, i1 - i0);
        for (size_t i = i0; i < i1; i++) {
            uint8_t m = 0, payload = 0;
            if (decode_byte(j->dna[i], 0, &m, &payload) != QDO_OK) break;
            j->out[j->n - 1 - i] = mask_to_ascii(complement_mask(m));
        }
    }
}

double qdo_bench_revcomp(const uint8_t *dna, size_t n, uint8_t *out, int threads, int reps) {
    if (threads < 1) threads = 1;
    rc_job j = {dna, n, out, threads, reps};
    const double t0 = now_s();
    run_jobs(rc_worker, &j, (size_t)threads, threads);
    return now_s() - t0;
}

/* ---------------------------------------------------------------- synthetic input + checksum ------------------- */

static inline uint64_t mix64(uint64_t z) { /* splitmix64 finaliser */
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

/* quickdna_b200/synth.py synth_np: nt[i] = alphabet[(mix64(seed * 0x1000003D + i + GOLD) >> 33) % alen] */
void qdo_synth(uint64_t seed, uint64_t start, uint64_t count, const uint8_t *alphabet, size_t alen, uint8_t *out) {
    const uint64_t off = seed * 0x1000003Dull + 0x9e3779b97f4a7c15ull + start;
    if (alen == 4)
        for (uint64_t i = 0; i < count; i++) out[i] = alphabet[(mix64(off + i) >> 33) & 3];
    else
        for (uint64_t i = 0; i < count; i++) out[i] = alphabet[(mix64(off + i) >> 33) % alen];
}

/* Position-weighted linear checksum (the one qd_checksum64_dev computes, include/quickdna_b200.h): the buffer is a
 * piece of a byte stream starting at stream offset `base`; stream word J (bytes 8J..8J+7, little endian) weighs
 * K(J) = mix64(J * GOLD + 1) | 1, and the checksum is sum_J word_J * K(J) mod 2^64.  Linear in the data, so pieces of
 * any byte alignment add up to the checksum of the whole. */
static inline uint64_t weight64(uint64_t word_index) { return mix64(word_index * 0x9e3779b97f4a7c15ull + 1) | 1; }

uint64_t qdo_checksum64(const uint8_t *buf, size_t n, uint64_t base) {
    uint64_t sum = 0;
    size_t i = 0;
    while (i < n && ((base + i) & 7)) { /* head: bytes of a partly covered word */
        sum += ((uint64_t)buf[i] << (8 * ((base + i) & 7))) * weight64((base + i) >> 3);
        i++;
    }
    for (; i + 8 <= n; i += 8) {
        uint64_t w;
        memcpy(&w, buf + i, 8);
        sum += w * weight64((base + i) >> 3);
    }
    for (; i < n; i++) sum += ((uint64_t)buf[i] << (8 * ((base + i) & 7))) * weight64((base + i) >> 3);
    return sum;
}

/* ---- whole-config checks: what the GPU path must produce for the synthetic configs, as a checksum ------------- */

typedef struct {
    uint64_t seed, first_read, n_reads;
    size_t read_len;
    int slot, frames, n_jobs;
    uint64_t *partial;
} chk_frames;

/* frames of read r in the slot layout of include/quickdna_b200.h (R = ceil(n/48) runs, 16*R bytes per frame slot,
 * forward frame k from the start of slot k, reverse frame k ending at the end of slot 3+k, zero padding) */
static void chk_frames_job(void *ctx, size_t job) {
    const chk_frames *c = (const chk_frames *)ctx;
    const size_t L = c->read_len, R = (L + 47) / 48, slot_bytes = 16 * R, per_read = slot_bytes * (size_t)c->frames;
    const uint64_t r0 = c->n_reads * job / (uint64_t)c->n_jobs, r1 = c->n_reads * (job + 1) / (uint64_t)c->n_jobs;
    uint8_t *dna = (uint8_t *)malloc(3 * L + 2 * L + 16 + per_read), *masks = dna + L, *rc = masks + L, *aa = rc + L, *img = aa + 2 * L + 16;
    uint64_t sum = 0;
    for (uint64_t r = r0; r < r1; r++) {
        qdo_synth(c->seed, (c->first_read + r) * L, L, (const uint8_t *)"ACGT", 4, dna);
        qdo_err e;
        size_t nm = 0, lens[6] = {0, 0, 0, 0, 0, 0};
        if (qdo_parse(0, dna, L, masks, &nm, &e) != QDO_OK) continue;
        if (c->frames == 1) {
            qdo_translate_masks(c->slot, masks, nm, aa);
            lens[0] = nm / 3;
        } else if (c->frames == 3) {
            qdo_self_frames_masks(c->slot, masks, nm, aa, lens);
        } else {
            qdo_all_frames_masks(c->slot, masks, nm, aa, lens);
        }
        memset(img, 0, per_read);
        size_t off = 0;
        for (int f = 0; f < c->frames; f++) {
            if (f < 3) memcpy(img + (size_t)f * slot_bytes, aa + off, lens[f]);
            else memcpy(img + (size_t)(f + 1) * slot_bytes - lens[f], aa + off, lens[f]);
            off += lens[f];
        }
        sum += qdo_checksum64(img, per_read, r * per_read);
    }
    c->partial[job] = sum;
    free(dna);
}

/* checksum of the slot-layout output of reads [first_read, first_read + n_reads) of the seed's uniform ACGT stream
 * (read r = stream bytes [r * read_len, (r + 1) * read_len)); the output stream starts at read first_read */
uint64_t qdo_check_frames_uniform(uint64_t seed, uint64_t first_read, uint64_t n_reads, size_t read_len, int slot, int frames,
                                  int threads) {
    qdo_tables();
    const int n_jobs = threads < 1 ? 1 : threads * 8;
    uint64_t *partial = (uint64_t *)calloc((size_t)n_jobs, 8);
    chk_frames c = {seed, first_read, n_reads, read_len, slot, frames, n_jobs, partial};
    run_jobs(chk_frames_job, &c, (size_t)n_jobs, threads);
    uint64_t sum = 0;
    for (int i = 0; i < n_jobs; i++) sum += partial[i];
    free(partial);
    return sum;
}

typedef struct {
    uint64_t seed, start, n;
    int n_jobs, strict;
    uint64_t *partial;
    int failed;
} chk_rc;

static void chk_rc_job(void *ctx, size_t job) {
    chk_rc *c = (chk_rc *)ctx;
    /* output range [o0, o1), cut at multiples of 64 bytes; input range [n - o1, n - o0) */
    const uint64_t blocks = (c->n + 63) / 64;
    uint64_t o0 = blocks * job / (uint64_t)c->n_jobs * 64, o1 = blocks * (job + 1) / (uint64_t)c->n_jobs * 64;
    if (o1 > c->n) o1 = c->n;
    uint64_t sum = 0;
    const uint64_t step = 1 << 20;
    uint8_t *in = (uint8_t *)malloc(2 * step), *out = in + step;
    for (uint64_t o = o0; o < o1; o += step) {
        const uint64_t len = o1 - o < step ? o1 - o : step;
        qdo_synth(c->seed, c->start + c->n - (o + len), len, (const uint8_t *)"ACGT", 4, in);
        qdo_err e;
        if (qdo_revcomp_bytes(c->strict, in, len, out, &e) != QDO_OK) c->failed = 1;
        sum += qdo_checksum64(out, len, o);
    }
    c->partial[job] = sum;
    free(in);
}

/* checksum of reverse_complement_bytes of stream bytes [start, start + n) of the seed's uniform ACGT stream */
uint64_t qdo_check_revcomp(uint64_t seed, uint64_t start, uint64_t n, int strict, int threads) {
    const int n_jobs = threads < 1 ? 1 : threads * 4;
    uint64_t *partial = (uint64_t *)calloc((size_t)n_jobs, 8);
    chk_rc c = {seed, start, n, n_jobs, strict, partial, 0};
    run_jobs(chk_rc_job, &c, (size_t)n_jobs, threads);
    uint64_t sum = 0;
    for (int i = 0; i < n_jobs; i++) sum += partial[i];
    free(partial);
    return c.failed ? ~sum : sum;
}

typedef struct {
    const uint8_t *dna; /* ASCII, n_seq sequences of seq_len back to back */
    size_t n_seq, seq_len, w, w_aa;
    int kind; /* 0 canonical rows, 1 all_windows */
    uint64_t *partial;
} chk_win;

static void chk_win_job(void *ctx, size_t s) {
    const chk_win *c = (const chk_win *)ctx;
    const size_t L = c->seq_len, w = c->w;
    const uint8_t *src = c->dna + s * L;
    uint8_t *masks = (uint8_t *)calloc(L ? L : 1, 1);
    for (size_t i = 0; i < L; i++) masks[i] = ASCII_TO_MASK[src[i]];
    uint64_t sum = 0;
    if (c->kind == 0) {
        /* qd_canonical_windows_batch_dev layout: [n_seq][n_win][w], input encoding (ASCII, upper case) */
        const size_t nw = L >= w ? L - w + 1 : 0;
        uint8_t row[64], asc[64];
        for (size_t i = 0; i < nw; i++) {
            qdo_canonical(masks + i, w, row);
            for (size_t k = 0; k < w; k++) asc[k] = mask_to_ascii(row[k]);
            sum += qdo_checksum64(asc, w, (uint64_t)(s * nw + i) * w);
        }
    } else {
        /* qd_windows_all_batch_dev layout: dna_out = [fwd: n_seq x c0 x w][rc: n_seq x c0 x w];
         * aa_out = six blocks [n_seq][c_f][w_aa]; the aa stream is checksummed after the dna stream */
        size_t counts[8];
        const size_t total = qdo_windows_all(masks, L, w, c->w_aa, NULL, NULL, NULL, counts);
        (void)total;
        uint8_t *dna_out = (uint8_t *)malloc((counts[0] + counts[1]) * w + 1);
        size_t n_aa = 0;
        for (int f = 0; f < 6; f++) n_aa += counts[2 + f];
        uint8_t *aa_out = (uint8_t *)malloc(n_aa * c->w_aa + 1);
        qdo_windows_all(masks, L, w, c->w_aa, dna_out, aa_out, NULL, counts);
        const uint64_t fwd_block = (uint64_t)c->n_seq * counts[0] * w;
        sum += qdo_checksum64(dna_out, counts[0] * w, (uint64_t)s * counts[0] * w);
        sum += qdo_checksum64(dna_out + counts[0] * w, counts[1] * w, fwd_block + (uint64_t)s * counts[1] * w);
        uint64_t aa_base = 2 * fwd_block;
        size_t at = 0;
        for (int f = 0; f < 6; f++) {
            sum += qdo_checksum64(aa_out + at * c->w_aa, counts[2 + f] * c->w_aa, aa_base + (uint64_t)s * counts[2 + f] * c->w_aa);
            aa_base += (uint64_t)c->n_seq * counts[2 + f] * c->w_aa;
            at += counts[2 + f];
        }
        free(dna_out);
        free(aa_out);
    }
    c->partial[s] = sum;
    free(masks);
}

/* checksum of the canonical w-nt windows (kind 0) / the all_windows output (kind 1; w = DNA window, w_aa = amino-acid window)
 * of n_seq strict ASCII sequences of seq_len each, in the layouts of qd_canonical_windows_batch_dev /
 * qd_windows_all_batch_dev (dna_out stream followed by aa_out stream) */
uint64_t qdo_check_windows(int kind, const uint8_t *dna, size_t n_seq, size_t seq_len, size_t w, size_t w_aa, int threads) {
    qdo_tables();
    if (w > 64 || w_aa > 64) return 0;
    uint64_t *partial = (uint64_t *)calloc(n_seq ? n_seq : 1, 8);
    chk_win c = {dna, n_seq, seq_len, w, w_aa, kind, partial};
    run_jobs(chk_win_job, &c, n_seq, threads);
    uint64_t sum = 0;
    for (size_t i = 0; i < n_seq; i++) sum += partial[i];
    free(partial);
    return sum;
}

typedef struct {
    const uint8_t *win; /* ASCII windows, n_win x w */
    size_t n_win, w, n_jobs;
    uint64_t cap;
    uint64_t *rows, *sums; /* per job */
} chk_exp;

/* rows of job j are checksummed relative to the job's first row; qdo_check_expansions rebases them */
static void chk_exp_job(void *ctx, size_t job) {
    chk_exp *c = (chk_exp *)ctx;
    const size_t i0 = c->n_win * job / c->n_jobs, i1 = c->n_win * (job + 1) / c->n_jobs, w = c->w;
    uint8_t *masks = (uint8_t *)malloc(w + 1);
    uint64_t n_rows = 0;
    /* first pass: rows of this job (needed before positions are known) */
    for (size_t i = i0; i < i1; i++) {
        for (size_t k = 0; k < w; k++) masks[k] = ASCII_TO_MASK[c->win[i * w + k]];
        uint64_t cnt = 0;
        if (qdo_expansions_size_hint(masks, w, &cnt) && cnt <= c->cap) n_rows += cnt;
    }
    c->rows[job] = n_rows;
    free(masks);
}

static void chk_exp_job2(void *ctx, size_t job) {
    chk_exp *c = (chk_exp *)ctx;
    const size_t i0 = c->n_win * job / c->n_jobs, i1 = c->n_win * (job + 1) / c->n_jobs, w = c->w;
    uint8_t *masks = (uint8_t *)malloc(w + (size_t)c->cap * w * 2 + 1), *rows = masks + w, *asc = rows + (size_t)c->cap * w;
    uint64_t row = c->rows[job]; /* exclusive prefix by now */
    uint64_t sum = 0;
    for (size_t i = i0; i < i1; i++) {
        for (size_t k = 0; k < w; k++) masks[k] = ASCII_TO_MASK[c->win[i * w + k]];
        uint64_t cnt = 0;
        if (!(qdo_expansions_size_hint(masks, w, &cnt) && cnt <= c->cap)) continue; /* benches/expansions.rs:57-66 */
        const size_t got = qdo_expansions(masks, w, rows, (size_t)cnt);
        for (size_t k = 0; k < got * w; k++) asc[k] = mask_to_ascii(rows[k]);
        sum += qdo_checksum64(asc, got * w, row * w);
        row += got;
    }
    c->sums[job] = sum;
    free(masks);
}

/* checksum of the expansion rows (ASCII, iteration order, windows with more than `cap` expansions skipped) of n_win
 * ASCII windows of w nt; *total_rows receives the row count */
uint64_t qdo_check_expansions(const uint8_t *windows, size_t n_win, size_t w, uint64_t cap, int threads, uint64_t *total_rows) {
    const size_t n_jobs = (size_t)(threads < 1 ? 1 : threads * 4);
    uint64_t *rows = (uint64_t *)calloc(2 * n_jobs, 8), *sums = rows + n_jobs;
    chk_exp c = {windows, n_win, w, n_jobs, cap, rows, sums};
    run_jobs(chk_exp_job, &c, n_jobs, threads);
    uint64_t acc = 0;
    for (size_t j = 0; j < n_jobs; j++) {
        const uint64_t r = rows[j];
        rows[j] = acc;
        acc += r;
    }
    if (total_rows) *total_rows = acc;
    run_jobs(chk_exp_job2, &c, n_jobs, threads);
    uint64_t sum = 0;
    for (size_t j = 0; j < n_jobs; j++) sum += sums[j];
    free(rows);
    return sum;
}

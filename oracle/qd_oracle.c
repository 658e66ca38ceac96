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

/* src/nucleotide.rs:66-101: ASCII -> mask, 0 = None. Upper and lower case both accepted. */
static uint8_t ascii_to_mask(uint8_t b) {
    switch (b) {
        case 'a': case 'A': return M_A;
        case 't': case 'T': return M_T;
        case 'c': case 'C': return M_C;
        case 'g': case 'G': return M_G;
        case 'n': case 'N': return M_A | M_T | M_C | M_G;
        case 'm': case 'M': return M_A | M_C;
        case 'r': case 'R': return M_A | M_G;
        case 'w': case 'W': return M_A | M_T;
        case 's': case 'S': return M_C | M_G;
        case 'y': case 'Y': return M_T | M_C;
        case 'k': case 'K': return M_T | M_G;
        case 'v': case 'V': return M_A | M_C | M_G;
        case 'h': case 'H': return M_A | M_T | M_C;
        case 'd': case 'D': return M_A | M_T | M_G;
        case 'b': case 'B': return M_T | M_C | M_G;
        default: return 0;
    }
}

/* src/nucleotide.rs:138-145, 219-237: always upper case. */
uint8_t qdo_mask_to_ascii(uint8_t m) {
    static const char LETTERS[16] = {'?', 'A', 'T', 'W', 'C', 'M', 'Y', 'H',
                                     'G', 'R', 'K', 'D', 'S', 'V', 'B', 'N'};
    return (uint8_t)LETTERS[m & 15];
}

/* src/nucleotide.rs:125-132, 195-213, written out as the reference's match. */
uint8_t qdo_complement_mask(uint8_t m) {
    switch (m) {
        case 1: return 2;    /* A -> T */
        case 2: return 1;    /* T -> A */
        case 3: return 3;    /* W -> W */
        case 4: return 8;    /* C -> G */
        case 5: return 10;   /* M -> K */
        case 6: return 9;    /* Y -> R */
        case 7: return 11;   /* H -> D */
        case 8: return 4;    /* G -> C */
        case 9: return 6;    /* R -> Y */
        case 10: return 5;   /* K -> M */
        case 11: return 7;   /* D -> H */
        case 12: return 12;  /* S -> S */
        case 13: return 14;  /* V -> B */
        case 14: return 13;  /* B -> V */
        case 15: return 15;  /* N -> N */
        default: return 0;
    }
}

static int popcount4(uint8_t m) { return (m & 1) + ((m >> 1) & 1) + ((m >> 2) & 1) + ((m >> 3) & 1); }

/* src/nucleotide.rs:285-299 (strict) / :301-315 (ambiguous) / :268-283 (ambiguous -> strict). */
int qdo_decode_byte(uint8_t b, int strict, uint8_t *mask, uint8_t *payload) {
    if (b >= 128) {
        *payload = b;
        return QDO_NON_ASCII_BYTE;
    }
    uint8_t m = ascii_to_mask(b);
    if (m == 0) {
        *payload = b; /* BadNucleotide(u.into()): the char as given, case preserved */
        return QDO_BAD_NUCLEOTIDE;
    }
    if (strict && popcount4(m) > 1) {
        *payload = qdo_mask_to_ascii(m); /* other.into(): upper-case letter of the code */
        return QDO_UNEXPECTED_AMBIGUOUS;
    }
    *mask = m;
    return QDO_OK;
}

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
            int rc = qdo_decode_byte(dna[i + j], strict, &m[j], &payload);
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
        int rc = qdo_decode_byte(dna[i], strict, &m, &payload);
        if (rc != QDO_OK) {
            err->kind = rc;
            err->byte = payload;
            err->pos = i;
            return rc;
        }
        out[n - 1 - i] = qdo_mask_to_ascii(qdo_complement_mask(m));
    }
    err->kind = QDO_OK;
    return QDO_OK;
}

/* src/trans_table.rs:284-286 */
void qdo_revcomp_masks(const uint8_t *m, size_t n, uint8_t *out) {
    for (size_t i = 0; i < n; i++) out[i] = qdo_complement_mask(m[n - 1 - i]);
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
            int rc = qdo_decode_byte(b, strict, &m, &payload);
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
                    const int rc = qdo_decode_byte(b, strict, &m, &payload);
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
    for (size_t i = 0; i < n; i++) ascii[i] = qdo_mask_to_ascii(m[i]);
    qdo_revcomp_masks(m, n, rc);
    for (size_t i = 0; i < n; i++) rc_ascii[i] = qdo_mask_to_ascii(rc[i]);
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

/* ---------------------------------------------------------------- timed baseline -- */

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

typedef struct {
    int slot;
    const uint8_t *dna;
    size_t r0, r1, read_len;
    uint8_t *out;
} frames_job;

static void *frames_worker(void *arg) {
    frames_job *j = (frames_job *)arg;
    size_t L = j->read_len;
    size_t out_per_read = L >= 2 ? 2 * (L - 2) : 0;
    uint8_t *masks = (uint8_t *)malloc(L ? L : 1);
    size_t lens[6];
    for (size_t r = j->r0; r < j->r1; r++) {
        const uint8_t *src = j->dna + r * L;
        /* the parse of rust_api.rs:310-322 (validation lives here on the Rust path) */
        qdo_err e;
        size_t nm = 0;
        if (qdo_parse(0, src, L, masks, &nm, &e) != QDO_OK) continue;
        qdo_all_frames_masks(j->slot, masks, nm, j->out + r * out_per_read, lens);
    }
    free(masks);
    return NULL;
}

double qdo_bench_all_frames_batch(int slot, const uint8_t *dna, size_t n_reads, size_t read_len,
                                  uint8_t *out, int threads, int reps) {
    if (threads < 1) threads = 1;
    qdo_tables();
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    frames_job *jobs = (frames_job *)malloc(sizeof(frames_job) * (size_t)threads);
    double t0 = now_s();
    for (int rep = 0; rep < reps; rep++) {
        for (int t = 0; t < threads; t++) {
            jobs[t].slot = slot;
            jobs[t].dna = dna;
            jobs[t].read_len = read_len;
            jobs[t].out = out;
            jobs[t].r0 = n_reads * (size_t)t / (size_t)threads;
            jobs[t].r1 = n_reads * (size_t)(t + 1) / (size_t)threads;
            pthread_create(&tid[t], NULL, frames_worker, &jobs[t]);
        }
        for (int t = 0; t < threads; t++) pthread_join(tid[t], NULL);
    }
    double dt = now_s() - t0;
    free(tid);
    free(jobs);
    return dt;
}

typedef struct {
    const uint8_t *dna;
    size_t n, i0, i1;
    uint8_t *out;
} rc_job;

/* one thread's share of the loop at src/trans_table.rs:273-276 */
static void *rc_worker(void *arg) {
    rc_job *j = (rc_job *)arg;
    memset(j->out + (j->n - j->i1), 0, j->i1 - j->i0);
    for (size_t i = j->i0; i < j->i1; i++) {
        uint8_t m = 0, payload = 0;
        if (qdo_decode_byte(j->dna[i], 0, &m, &payload) != QDO_OK) break;
        j->out[j->n - 1 - i] = qdo_mask_to_ascii(qdo_complement_mask(m));
    }
    return NULL;
}

double qdo_bench_revcomp(const uint8_t *dna, size_t n, uint8_t *out, int threads, int reps) {
    if (threads < 1) threads = 1;
    pthread_t *tid = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    rc_job *jobs = (rc_job *)malloc(sizeof(rc_job) * (size_t)threads);
    double t0 = now_s();
    for (int rep = 0; rep < reps; rep++) {
        for (int t = 0; t < threads; t++) {
            jobs[t].dna = dna;
            jobs[t].n = n;
            jobs[t].out = out;
            jobs[t].i0 = n * (size_t)t / (size_t)threads;
            jobs[t].i1 = n * (size_t)(t + 1) / (size_t)threads;
            pthread_create(&tid[t], NULL, rc_worker, &jobs[t]);
        }
        for (int t = 0; t < threads; t++) pthread_join(tid[t], NULL);
    }
    double dt = now_s() - t0;
    free(tid);
    free(jobs);
    return dt;
}

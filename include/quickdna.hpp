// quickdna.hpp -- C++17 host mirror of the reference's Rust API for the sequence hot path
// (src/rust_api.rs, src/trans_table.rs, src/nucleotide.rs, src/errors.rs of SecureDNA/quickdna),
// header-only, over the C ABI of quickdna_b200.h.
//
// The reference is a Rust crate; Rust is not available in this build environment, so this header plays the
// role the rebound crate would: the same type and method names, argument meaning, frame rules and error
// vocabulary, every method body one call into libquickdna_b200.so (INTEGRATION.md shows the Rust side).
// A `DnaSequence<T>` stores one byte per nucleotide whose value IS the 4-bit IUPAC mask, exactly as
// `Vec<Nucleotide>` / `Vec<NucleotideAmbiguous>` do (src/nucleotide.rs:20-53), so buffers cross the
// boundary as QD_ENC_MASK without conversion.  There is no CPU fallback: constructing a Context without a
// usable GPU throws.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <string_view>
#include <utility>
#include <type_traits>
#include <vector>

#include "quickdna_b200.h"

namespace quickdna {

// src/nucleotide.rs:20-28
enum class Nucleotide : uint8_t { A = 1, T = 2, C = 4, G = 8 };
// src/nucleotide.rs:31-53
enum class NucleotideAmbiguous : uint8_t {
    A = 1, T = 2, W = 3, C = 4, M = 5, Y = 6, H = 7, G = 8, R = 9, K = 10, D = 11, S = 12, V = 13, B = 14, N = 15
};

template <class T>
inline constexpr bool is_nucleotide_like = std::is_same_v<T, Nucleotide> || std::is_same_v<T, NucleotideAmbiguous>;
template <class T>
inline constexpr int strict_of = std::is_same_v<T, Nucleotide> ? 1 : 0;

// src/nucleotide.rs:138-145, 219-237
template <class T>
inline char to_ascii(T n) { return "?ATWCMYHGRKDSVBN"[static_cast<uint8_t>(n) & 15]; }
// src/nucleotide.rs:125-132, 195-213
template <class T>
inline T complement(T n) {
    const uint8_t m = static_cast<uint8_t>(n);
    return static_cast<T>(((m & 5) << 1) | ((m & 10) >> 1));
}

// src/trans_table.rs:11-73 (ids accepted by TryFrom<u8>, :230-267)
enum class TranslationTable : uint8_t {
    Ncbi1 = 1, Ncbi2, Ncbi3, Ncbi4, Ncbi5, Ncbi6, Ncbi7, Ncbi8, Ncbi9, Ncbi10, Ncbi11, Ncbi12, Ncbi13, Ncbi14, Ncbi15, Ncbi16,
    Ncbi21 = 21, Ncbi22, Ncbi23, Ncbi24, Ncbi25, Ncbi26, Ncbi27, Ncbi28, Ncbi29, Ncbi30, Ncbi31, Ncbi32, Ncbi33
};

// src/errors.rs:17-29; what() is the reference's Display text
class TranslationError : public std::runtime_error {
  public:
    enum Kind { NonAsciiByte = 1, BadNucleotide = 2, UnexpectedAmbiguousNucleotide = 3, BadTranslationTable = 4 };
    TranslationError(const qd_err& e, std::string msg) : std::runtime_error(std::move(msg)), kind(static_cast<Kind>(e.kind)), byte(e.byte), pos(e.pos) {}
    Kind kind;
    uint8_t byte;   // payload of the variant
    uint64_t pos;   // extra: index of the offending byte
};

class RuntimeFailure : public std::runtime_error {
    using std::runtime_error::runtime_error;
};

// One qd_ctx per thread (the reference's functions are pure and re-entrant).
class Context {
  public:
    explicit Context(int device = 0) {
        if (qd_ctx_create(&ctx_, device) != QD_OK) throw RuntimeFailure("quickdna_b200: no usable CUDA device (there is no CPU fallback)");
    }
    ~Context() { qd_ctx_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    qd_ctx* get() const { return ctx_; }
    static Context& current() {
        thread_local Context c;
        return c;
    }
    void check(int rc, const qd_err& e) const {
        if (rc == QD_OK) return;
        if (rc >= QD_ERR_NON_ASCII_BYTE && rc <= QD_ERR_BAD_TABLE) {
            char buf[96];
            qd_err_message(&e, buf, sizeof buf);
            throw TranslationError(e, buf);
        }
        throw RuntimeFailure(std::string("quickdna_b200 status ") + std::to_string(rc) + ": " + qd_last_cuda_error(ctx_));
    }

  private:
    qd_ctx* ctx_ = nullptr;
};

// TranslationTable::try_from(u8), src/trans_table.rs:230-267
inline TranslationTable table_try_from(uint8_t id) {
    if (qd_table_slot(id) < 0) {
        qd_err e{QD_ERR_BAD_TABLE, id, 0, 0};
        char buf[96];
        qd_err_message(&e, buf, sizeof buf);
        throw TranslationError(e, buf);
    }
    return static_cast<TranslationTable>(id);
}

// ---- src/trans_table.rs free functions ----------------------------------------------------------------

// TranslationTable::translate_dna_bytes::<T>, src/trans_table.rs:176-203
template <class T>
std::vector<uint8_t> translate_dna_bytes(TranslationTable table, std::string_view dna) {
    static_assert(is_nucleotide_like<T>);
    std::vector<uint8_t> out(dna.size() / 3);
    qd_err e{};
    Context& c = Context::current();
    c.check(qd_translate(c.get(), static_cast<int>(table), QD_ENC_ASCII, strict_of<T>, reinterpret_cast<const uint8_t*>(dna.data()),
                         dna.size(), out.data(), &e), e);
    return out;
}

// TranslationTable::translate_dna::<T>, src/trans_table.rs:205-227
template <class T>
std::vector<uint8_t> translate_dna(TranslationTable table, const T* dna, size_t n) {
    static_assert(is_nucleotide_like<T> && sizeof(T) == 1);
    std::vector<uint8_t> out(n / 3);
    qd_err e{};
    Context& c = Context::current();
    c.check(qd_translate(c.get(), static_cast<int>(table), QD_ENC_MASK, 0, reinterpret_cast<const uint8_t*>(dna), n, out.data(), &e), e);
    return out;
}

// reverse_complement_bytes::<T>, src/trans_table.rs:269-278
template <class T>
std::vector<uint8_t> reverse_complement_bytes(std::string_view dna) {
    std::vector<uint8_t> out(dna.size());
    qd_err e{};
    Context& c = Context::current();
    c.check(qd_revcomp(c.get(), QD_ENC_ASCII, strict_of<T>, reinterpret_cast<const uint8_t*>(dna.data()), dna.size(), out.data(), &e), e);
    return out;
}

// reverse_complement::<T>, src/trans_table.rs:284-286
template <class T>
std::vector<T> reverse_complement(const T* dna, size_t n) {
    std::vector<T> out(n);
    qd_err e{};
    Context& c = Context::current();
    c.check(qd_revcomp(c.get(), QD_ENC_MASK, 0, reinterpret_cast<const uint8_t*>(dna), n, reinterpret_cast<uint8_t*>(out.data()), &e), e);
    return out;
}

// ---- src/rust_api.rs ---------------------------------------------------------------------------------

// src/rust_api.rs:77-161
class ProteinSequence {
  public:
    ProteinSequence() = default;
    explicit ProteinSequence(std::vector<uint8_t> aa) : amino_acids_(std::move(aa)) {}
    // TryFrom<&[u8]> / FromStr, :132-161: ASCII only, upper-cased
    static ProteinSequence from_str(std::string_view s) {
        std::vector<uint8_t> v(s.begin(), s.end());
        for (uint8_t& b : v) {
            if (b >= 128) {
                qd_err e{QD_ERR_NON_ASCII_BYTE, b, 0, 0};
                char buf[96];
                qd_err_message(&e, buf, sizeof buf);
                throw TranslationError(e, buf);
            }
            if (b >= 'a' && b <= 'z') b = static_cast<uint8_t>(b - 32);
        }
        return ProteinSequence(std::move(v));
    }
    size_t len() const { return amino_acids_.size(); }
    bool is_empty() const { return amino_acids_.empty(); }
    const std::vector<uint8_t>& as_slice() const { return amino_acids_; }
    void push(uint8_t aa) { amino_acids_.push_back(aa); }
    uint8_t operator[](size_t i) const { return amino_acids_[i]; }
    std::string to_string() const { return std::string(amino_acids_.begin(), amino_acids_.end()); }
    bool operator==(const ProteinSequence& o) const { return amino_acids_ == o.amino_acids_; }
    bool operator!=(const ProteinSequence& o) const { return !(*this == o); }
    // :100-104 -- every length-`length` window as an owned sequence
    std::vector<ProteinSequence> windows(size_t length) const {
        std::vector<ProteinSequence> res;
        if (length == 0 || amino_acids_.size() < length) return res;
        const size_t count = amino_acids_.size() - length + 1;
        std::vector<uint8_t> flat(count * length);
        size_t got = 0;
        Context& c = Context::current();
        c.check(qd_windows(c.get(), amino_acids_.data(), amino_acids_.size(), length, flat.data(), &got), qd_err{});
        res.reserve(got);
        for (size_t i = 0; i < got; i++) res.emplace_back(std::vector<uint8_t>(flat.begin() + i * length, flat.begin() + (i + 1) * length));
        return res;
    }

  private:
    std::vector<uint8_t> amino_acids_;
};

// src/rust_api.rs:163-387
template <class T>
class DnaSequence {
    static_assert(is_nucleotide_like<T> && sizeof(T) == 1);

  public:
    DnaSequence() = default;
    explicit DnaSequence(std::vector<T> dna) : dna_(std::move(dna)) {}  // :210-212

    // TryFrom<&[u8]> / FromStr, :310-338: ' ' and '\t' are skipped, everything else validated
    static DnaSequence from_str(std::string_view s) {
        std::vector<T> v(s.size());
        size_t k = 0;
        qd_err e{};
        Context& c = Context::current();
        c.check(qd_parse(c.get(), strict_of<T>, reinterpret_cast<const uint8_t*>(s.data()), s.size(), reinterpret_cast<uint8_t*>(v.data()), &k, &e), e);
        v.resize(k);
        return DnaSequence(std::move(v));
    }

    size_t len() const { return dna_.size(); }
    bool is_empty() const { return dna_.empty(); }
    const std::vector<T>& as_slice() const { return dna_; }
    void push(T n) { dna_.push_back(n); }  // :281-283
    T operator[](size_t i) const { return dna_[i]; }
    bool operator==(const DnaSequence& o) const { return dna_ == o.dna_; }
    bool operator!=(const DnaSequence& o) const { return !(*this == o); }
    std::string to_string() const {  // Display, :301-308
        std::string s(dna_.size(), '?');
        for (size_t i = 0; i < dna_.size(); i++) s[i] = to_ascii(dna_[i]);
        return s;
    }

    // :216-219
    ProteinSequence translate(TranslationTable table) const { return ProteinSequence(translate_dna(table, dna_.data(), dna_.size())); }

    // :227-255 -- 3 frames if len >= 5, 2 if len == 4, 1 if len == 3, else none
    std::vector<ProteinSequence> translate_self_frames(TranslationTable table) const { return frames(table, QD_FRAMES_SELF); }
    // :263-270 -- forward frames then the frames of the reverse complement; one pass over the input
    std::vector<ProteinSequence> translate_all_frames(TranslationTable table) const { return frames(table, QD_FRAMES_ALL); }

    // :273-275
    DnaSequence reverse_complement() const { return DnaSequence(quickdna::reverse_complement(dna_.data(), dna_.size())); }

    // :277-279
    std::vector<DnaSequence> windows(size_t length) const {
        std::vector<DnaSequence> res;
        if (length == 0 || dna_.size() < length) return res;
        const size_t count = dna_.size() - length + 1;
        std::vector<uint8_t> flat(count * length);
        size_t got = 0;
        Context& c = Context::current();
        c.check(qd_windows(c.get(), bytes(), dna_.size(), length, flat.data(), &got), qd_err{});
        res.reserve(got);
        for (size_t i = 0; i < got; i++) {
            const T* w = reinterpret_cast<const T*>(flat.data() + i * length);
            res.emplace_back(std::vector<T>(w, w + length));
        }
        return res;
    }

    // :374-377 (DnaSequence<Nucleotide> only): min_lex(fw(seq), fw(reverse(seq))) under A<T<C<G
    template <class U = T, class = std::enable_if_t<std::is_same_v<U, Nucleotide>>>
    DnaSequence canonical() const {
        std::vector<T> out(dna_.size());
        qd_err e{};
        Context& c = Context::current();
        c.check(qd_canonical(c.get(), QD_ENC_MASK, 0, bytes(), dna_.size(), reinterpret_cast<uint8_t*>(out.data()), &e), e);
        return DnaSequence(std::move(out));
    }

    // Not in the reference: one fixed-width key per canonical length-`w` window, in window order --
    // what a set lookup downstream of `windows(w)` + `canonical()` needs, without materialising the windows.
    // kind = QD_CANON_KEY_FP64: one uint64 per window; QD_CANON_KEY_PACKED: qd_canonical_key_bytes(kind, w) / 8 uint64 per window.
    template <class U = T, class = std::enable_if_t<std::is_same_v<U, Nucleotide>>>
    std::vector<uint64_t> canonical_window_keys(size_t w, int kind = QD_CANON_KEY_FP64) const {
        const size_t kb = qd_canonical_key_bytes(kind, w);
        if (kb == 0) throw std::invalid_argument("quickdna: bad key kind or window length");
        std::vector<uint64_t> keys(dna_.size() >= w ? (dna_.size() - w + 1) * (kb / 8) : 0);
        qd_err e{};
        Context& c = Context::current();
        uint64_t none = 0;
        c.check(qd_canonical_window_keys(c.get(), QD_ENC_MASK, 0, kind, bytes(), dna_.size(), w, keys.empty() ? &none : keys.data(), &e), e);
        return keys;
    }

    // :385-387 (DnaSequence<NucleotideAmbiguous> only): every concrete expansion, lexicographic, materialised.
    // size_hint() of the reference is `expansion_count()`; UINT64_MAX == (usize::MAX, None).
    template <class U = T, class = std::enable_if_t<std::is_same_v<U, NucleotideAmbiguous>>>
    uint64_t expansion_count() const {
        if (dna_.empty()) return 1;
        uint64_t count = 0, index[2] = {0, 0};
        qd_err e{};
        Context& c = Context::current();
        c.check(qd_expansion_windows(c.get(), QD_ENC_MASK, bytes(), 1, dna_.size(), 0, &count, index, nullptr, &e), e);
        return count;
    }
    template <class U = T, class = std::enable_if_t<std::is_same_v<U, NucleotideAmbiguous>>>
    std::vector<DnaSequence<Nucleotide>> expansions(uint64_t max_expansions = uint64_t(1) << 24) const {
        std::vector<DnaSequence<Nucleotide>> res;
        if (dna_.empty()) {  // one empty expansion (src/expansions.rs:37-58 on an empty slice)
            res.emplace_back();
            return res;
        }
        uint64_t count = 0, index[2] = {0, 0};
        qd_err e{};
        Context& c = Context::current();
        c.check(qd_expansion_windows(c.get(), QD_ENC_MASK, bytes(), 1, dna_.size(), max_expansions, &count, index, nullptr, &e), e);
        if (count > max_expansions) throw std::length_error("quickdna: more expansions than max_expansions");
        std::vector<uint8_t> flat(static_cast<size_t>(count) * dna_.size());
        c.check(qd_expansion_windows(c.get(), QD_ENC_MASK, bytes(), 1, dna_.size(), max_expansions, &count, index, flat.data(), &e), e);
        res.reserve(count);
        for (uint64_t i = 0; i < count; i++) {
            const Nucleotide* w = reinterpret_cast<const Nucleotide*>(flat.data() + i * dna_.size());
            res.emplace_back(std::vector<Nucleotide>(w, w + dna_.size()));
        }
        return res;
    }

  private:
    const uint8_t* bytes() const { return reinterpret_cast<const uint8_t*>(dna_.data()); }
    std::vector<ProteinSequence> frames(TranslationTable table, int which) const {
        const size_t n = dna_.size();
        size_t cap = 0;
        for (int k = 0; k < 3; k++) cap += qd_frame_len(n, k);
        std::vector<uint8_t> buf((which == QD_FRAMES_ALL ? 2 : 1) * cap + 1);
        size_t lens[6] = {0, 0, 0, 0, 0, 0};
        int n_frames = 0;
        qd_err e{};
        Context& c = Context::current();
        c.check(qd_translate_frames(c.get(), static_cast<int>(table), QD_ENC_MASK, 0, which, bytes(), n, buf.data(), lens, &n_frames, &e), e);
        std::vector<ProteinSequence> res;
        size_t off = 0;
        for (size_t l : lens)
            if (l) {
                res.emplace_back(std::vector<uint8_t>(buf.begin() + off, buf.begin() + off + l));
                off += l;
            }
        return res;
    }
    std::vector<T> dna_;
};

// ---- src/iter.rs, src/trans_table.rs:165-174: per-element lazy adaptors ----------------------------------------
//
// The reference's `NucleotideIter` extension trait composes adaptors over any nucleotide iterator; every use it has is
// over a slice, so the mirror is a pair of O(1) views over a BORROWED span (no copies, no allocation per element):
// `NucleotideView` (a strand: forward / reversed, plain / complemented) and `CodonsView` (one reading frame of it).
// One element per call they are host-side by nature, exactly like the Rust iterators.  What the reference's callers
// DO with them is collect a whole frame -- `frame.map(table.to_fn()).collect()` (iter.rs:21-74, benches/all_windows.rs:
// 113-149) -- and those bulk forms run on the device in ONE launch: CodonsView::translate, NucleotideView::
// translate_all_frames / translate_self_frames, translate_codons (to_fn over a slice of codons), and
// NucleotideView::collect_bulk for the reverse-complement strand.  `to_fn` itself reads the same 27 x 4096-byte LUT image
// the device tables are built from.

// N::Codon, src/nucleotide.rs:354, 396
template <class T>
struct Codon {
    T n[3];
    bool operator==(const Codon& o) const { return n[0] == o.n[0] && n[1] == o.n[1] && n[2] == o.n[2]; }
    bool operator!=(const Codon& o) const { return !(*this == o); }
    std::string to_string() const { return {to_ascii(n[0]), to_ascii(n[1]), to_ascii(n[2])}; }  // Display, :380-384
};

// TranslationTable::to_fn, src/trans_table.rs:165-174: codon -> amino acid over one 4096-byte slice,
// index (m0 << 8) | (m1 << 4) | m2 (CodonIdx, :75-103)
inline auto to_fn(TranslationTable table) {
    const uint8_t* slice = qd_table_data(qd_table_slot(static_cast<int>(table)));
    return [slice](const auto& codon) -> uint8_t {
        return slice[(static_cast<unsigned>(codon.n[0]) << 8) | (static_cast<unsigned>(codon.n[1]) << 4) | static_cast<unsigned>(codon.n[2])];
    };
}

template <class T>
class CodonsView;

// `codons.iter().map(table.to_fn()).collect()` over a slice of codons (src/trans_table.rs:165-174 applied in bulk): a
// Codon is three mask bytes, so the slice IS a mask sequence of 3 n nucleotides -- one qd_translate launch.
template <class T>
std::vector<uint8_t> translate_codons(TranslationTable table, const std::vector<Codon<T>>& codons) {
    static_assert(sizeof(Codon<T>) == 3);
    return translate_dna(table, reinterpret_cast<const T*>(codons.data()), 3 * codons.size());
}

template <class T>
class NucleotideView {
    static_assert(is_nucleotide_like<T>);

  public:
    NucleotideView() = default;
    NucleotideView(const T* p, size_t n) : p_(p), n_(n) {}
    explicit NucleotideView(const std::vector<T>& v) : p_(v.data()), n_(v.size()) {}
    explicit NucleotideView(const DnaSequence<T>& s) : NucleotideView(s.as_slice()) {}

    size_t len() const { return n_; }
    bool is_empty() const { return n_ == 0; }
    T operator[](size_t j) const {
        const T v = p_[rev_ ? n_ - 1 - j : j];
        return comp_ ? quickdna::complement(v) : v;
    }
    // iter.rs:148-157
    NucleotideView complement() const {
        NucleotideView r = *this;
        r.comp_ = !comp_;
        return r;
    }
    // iter.rs:171-177 (`self.rev().complement()`)
    NucleotideView reverse_complement() const {
        NucleotideView r = *this;
        r.rev_ = !rev_;
        r.comp_ = !comp_;
        return r;
    }
    // `next()` k times / `next_back()` k times
    NucleotideView skip(size_t k) const {
        NucleotideView r = *this;
        k = k < n_ ? k : n_;
        if (!rev_) r.p_ += k;
        r.n_ -= k;
        return r;
    }
    NucleotideView skip_back(size_t k) const {
        NucleotideView r = *this;
        k = k < n_ ? k : n_;
        if (rev_) r.p_ += k;
        r.n_ -= k;
        return r;
    }
    // trim_to_codon / trimmed_to_codon, iter.rs:253-283
    NucleotideView trimmed_to_codon() const { return skip_back(n_ % 3); }
    // iter.rs:125-134
    CodonsView<T> codons() const;
    // iter.rs:221-236: frames at offsets 0, 1, 2; empty ones dropped
    std::vector<CodonsView<T>> self_reading_frames() const;
    // iter.rs:75-103: forward frames, then the frames of the reverse complement; empty ones dropped
    std::vector<CodonsView<T>> all_reading_frames() const;
    // `.collect()`, one element at a time on the host like the Rust iterator
    std::vector<T> collect() const {
        std::vector<T> v(n_);
        for (size_t j = 0; j < n_; j++) v[j] = (*this)[j];
        return v;
    }
    // `.collect()` in bulk: the reverse-complement strand of a span is ONE qd_revcomp launch (src/trans_table.rs:284-286);
    // the forward strand is a copy; the two one-sided adaptors (reversed only / complemented only) have no bulk form
    std::vector<T> collect_bulk() const {
        if (rev_ && comp_) return quickdna::reverse_complement(p_, n_);
        return collect();
    }
    // `self_reading_frames().map(|f| f.map(table.to_fn()).collect())` / the same for all_reading_frames (iter.rs:75-103,
    // 221-236) in ONE launch (qd_translate_frames); empty frames are dropped like the iterators drop them.
    std::vector<std::vector<uint8_t>> translate_self_frames(TranslationTable table) const { return bulk_frames(table, QD_FRAMES_SELF); }
    std::vector<std::vector<uint8_t>> translate_all_frames(TranslationTable table) const { return bulk_frames(table, QD_FRAMES_ALL); }
    bool is_forward() const { return !rev_ && !comp_; }
    bool is_reverse_complement() const { return rev_ && comp_; }
    const T* data() const { return p_; }

  private:
    std::vector<std::vector<uint8_t>> bulk_frames(TranslationTable table, int which) const {
        // the strand as a contiguous mask buffer: the span itself, or (rc / one-sided views) its collected form
        std::vector<T> own;
        const T* src = p_;
        if (!is_forward()) {
            own = collect_bulk();
            src = own.data();
        }
        size_t cap = 0;
        for (int k = 0; k < 3; k++) cap += qd_frame_len(n_, k);
        std::vector<uint8_t> buf((which == QD_FRAMES_ALL ? 2 : 1) * cap + 1);
        size_t lens[6] = {0, 0, 0, 0, 0, 0};
        int n_frames = 0;
        qd_err e{};
        Context& c = Context::current();
        c.check(qd_translate_frames(c.get(), static_cast<int>(table), QD_ENC_MASK, 0, which, reinterpret_cast<const uint8_t*>(src), n_, buf.data(), lens,
                                    &n_frames, &e), e);
        std::vector<std::vector<uint8_t>> res;
        size_t off = 0;
        for (size_t l : lens)
            if (l) {
                res.emplace_back(buf.begin() + off, buf.begin() + off + l);
                off += l;
            }
        return res;
    }
    const T* p_ = nullptr;
    size_t n_ = 0;
    bool rev_ = false, comp_ = false;
};

// Codons<N, I> / ForwardOrRcCodons<N, I>, iter.rs:292-448: one view type serves both, the strand is a property of the
// NucleotideView underneath.  Double-ended like the reference: next() takes from the front, next_back() from the back
// (after trimming the incomplete trailing codon, :322-323).
template <class T>
class CodonsView {
  public:
    CodonsView() = default;
    explicit CodonsView(NucleotideView<T> v) : v_(v), front_(0), back_(v.len() / 3) {}
    size_t len() const { return back_ - front_; }  // ExactSizeIterator, :335-343
    bool is_empty() const { return len() == 0; }
    Codon<T> operator[](size_t i) const {
        const size_t j = 3 * (front_ + i);
        return Codon<T>{{v_[j], v_[j + 1], v_[j + 2]}};
    }
    bool next(Codon<T>* out) {
        if (front_ == back_) return false;
        *out = (*this)[0];
        front_++;
        return true;
    }
    bool next_back(Codon<T>* out) {
        if (front_ == back_) return false;
        *out = (*this)[len() - 1];
        back_--;
        return true;
    }
    std::vector<Codon<T>> collect() const {
        std::vector<Codon<T>> r;
        r.reserve(len());
        for (size_t i = 0; i < len(); i++) r.push_back((*this)[i]);
        return r;
    }
    std::vector<Codon<T>> collect_rev() const {  // `.rev().collect()`
        std::vector<Codon<T>> r;
        r.reserve(len());
        for (size_t i = len(); i-- > 0;) r.push_back((*this)[i]);
        return r;
    }
    // `.map(f).collect()`, e.g. f = to_fn(table): per element on the host, as the Rust iterator does
    template <class F>
    std::vector<uint8_t> map(F f) const {
        std::vector<uint8_t> r(len());
        for (size_t i = 0; i < len(); i++) r[i] = f((*this)[i]);
        return r;
    }
    // `.map(table.to_fn()).collect()` in bulk: the frame's nucleotides are contiguous masks (the forward strand in place,
    // any other strand collected first -- the reverse complement by one qd_revcomp launch), so it is one qd_translate launch
    std::vector<uint8_t> translate(TranslationTable table) const {
        const size_t n = 3 * len();
        if (v_.is_forward()) return translate_dna(table, v_.data() + 3 * front_, n);
        const std::vector<T> strand = v_.collect_bulk();
        return translate_dna(table, strand.data() + 3 * front_, n);
    }

  private:
    NucleotideView<T> v_;
    size_t front_ = 0, back_ = 0;
};

template <class T>
CodonsView<T> NucleotideView<T>::codons() const {
    return CodonsView<T>(*this);
}
template <class T>
std::vector<CodonsView<T>> NucleotideView<T>::self_reading_frames() const {
    std::vector<CodonsView<T>> frames;
    for (size_t k = 0; k < 3; k++)
        if (CodonsView<T> f = skip(k).codons(); f.len() > 0) frames.push_back(f);
    return frames;
}
template <class T>
std::vector<CodonsView<T>> NucleotideView<T>::all_reading_frames() const {
    std::vector<CodonsView<T>> frames = self_reading_frames();
    for (const CodonsView<T>& f : reverse_complement().self_reading_frames()) frames.push_back(f);
    return frames;
}

// ---- src/fasta.rs ------------------------------------------------------------------------------------

// FastaParseSettings, src/fasta.rs:67-110
struct FastaParseSettings {
    bool concatenate_headers = true;
    bool allow_preceding_comment = false;
};

// FastaRecord<DnaSequence<T>>, src/fasta.rs:17-29
template <class T>
struct FastaRecord {
    std::string header;                      // without the leading '>' / ';'; concatenated header lines joined by '\n'
    DnaSequence<T> contents;
    std::pair<size_t, size_t> line_range;    // 1-based, end exclusive
};

// Located<FastaParseError::ParseError(TranslationError)>, src/fasta.rs:413-423 + src/errors.rs:8-15
class FastaParseError : public TranslationError {
  public:
    FastaParseError(const qd_err& e, const std::string& msg)
        : TranslationError(e, "on line " + std::to_string(e.seq_index) + ": error parsing record: " + msg), line_number(e.seq_index) {}
    uint64_t line_number;
};

// FastaParser<DnaSequence<T>>, src/fasta.rs:425-480
template <class T>
class FastaParser {
  public:
    explicit FastaParser(FastaParseSettings settings = {}) : settings_(settings) {}
    std::vector<FastaRecord<T>> parse_str(std::string_view s) const {
        Context& c = Context::current();
        const uint8_t* text = reinterpret_cast<const uint8_t*>(s.data());
        size_t n_masks = 0, n_records = 0;
        qd_err e{};
        auto call = [&](uint8_t* masks, qd_fasta_record* recs, size_t cap) {  // false: `cap` records were not enough
            const int rc = qd_fasta_scan(c.get(), strict_of<T>, settings_.concatenate_headers, settings_.allow_preceding_comment, text, s.size(), masks,
                                         &n_masks, recs, cap, &n_records, &e);
            if (rc >= QD_ERR_NON_ASCII_BYTE && rc <= QD_ERR_UNEXPECTED_AMBIGUOUS) {
                char buf[96];
                qd_err_message(&e, buf, sizeof buf);
                throw FastaParseError(e, buf);
            }
            if (rc == QD_ERR_INVALID_ARGUMENT && n_records > cap) return false;
            c.check(rc, e);
            return true;
        };
        // one call in the usual case: there are at most s.size() masks, and a record of ordinary FASTA is longer than 64 bytes;
        // only a file of very short records needs the exact count (the refused call has reported it) and a second call
        std::vector<uint8_t> masks(s.size() + 1);
        std::vector<qd_fasta_record> recs(s.size() / 64 + 16);
        if (!call(masks.data(), recs.data(), recs.size())) {
            recs.resize(n_records + 1);
            call(masks.data(), recs.data(), n_records);
        }
        std::vector<FastaRecord<T>> out;
        out.reserve(n_records);
        for (size_t i = 0; i < n_records; i++) {
            const qd_fasta_record& r = recs[i];
            FastaRecord<T> rec;
            if (r.header_begin != UINT64_MAX) {  // header lines without their first byte, line breaks normalised to '\n'
                size_t pos = r.header_begin;
                while (pos < r.header_end) {
                    size_t eol = pos;
                    while (eol < r.header_end && s[eol] != '\n') eol++;
                    size_t end = (eol < r.header_end && eol > pos && s[eol - 1] == '\r') ? eol - 1 : eol;
                    if (!rec.header.empty() || pos != r.header_begin) rec.header.push_back('\n');
                    rec.header.append(s.substr(pos + 1, end - pos - 1));
                    pos = eol + 1;
                }
            }
            const T* m = reinterpret_cast<const T*>(masks.data() + r.content_begin);
            rec.contents = DnaSequence<T>(std::vector<T>(m, m + (r.content_end - r.content_begin)));
            rec.line_range = {static_cast<size_t>(r.line_begin), static_cast<size_t>(r.line_end)};
            out.push_back(std::move(rec));
        }
        return out;
    }

  private:
    FastaParseSettings settings_;
};

using DnaSequenceStrict = DnaSequence<Nucleotide>;              // src/rust_api.rs:163
using DnaSequenceAmbiguous = DnaSequence<NucleotideAmbiguous>;  // src/rust_api.rs:164

}  // namespace quickdna

// Bulk (device) forms of the lazy adaptor mirror in include/quickdna.hpp against their per-element host forms, which
// tests/cpp/test_iter_views.cpp pins to the doctests of src/iter.rs: `frame.map(table.to_fn()).collect()` for every
// frame of all_reading_frames (iter.rs:21-74, benches/all_windows.rs:113-149), to_fn over a slice of codons
// (src/trans_table.rs:165-174) and `.reverse_complement().collect()` (iter.rs:161-177).  Needs a GPU.
// Run by tests/test_cpp_mirror.py (-m gpu).
#include <cstdio>
#include <cstdint>
#include <vector>

#include "quickdna.hpp"

using namespace quickdna;

static int failures = 0;
#define CHECK(cond)                                                     \
    do {                                                                \
        if (!(cond)) {                                                  \
            std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            failures++;                                                 \
        }                                                               \
    } while (0)

static uint64_t rng_state = 0x9e3779b97f4a7c15ull;
static uint32_t rnd() {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return (uint32_t)(rng_state >> 32);
}

template <class T>
static void check_sequence(const std::vector<T>& dna, TranslationTable table) {
    const NucleotideView<T> v(dna);
    const auto fn = to_fn(table);
    // every frame of all_reading_frames / self_reading_frames: bulk == per element
    const auto frames = v.all_reading_frames();
    const auto bulk = v.translate_all_frames(table);
    CHECK(frames.size() == bulk.size());
    for (size_t f = 0; f < frames.size() && f < bulk.size(); f++) {
        CHECK(frames[f].map(fn) == bulk[f]);
        CHECK(frames[f].translate(table) == bulk[f]);
    }
    const auto self_bulk = v.translate_self_frames(table);
    const auto self_frames = v.self_reading_frames();
    CHECK(self_frames.size() == self_bulk.size());
    for (size_t f = 0; f < self_frames.size() && f < self_bulk.size(); f++) CHECK(self_frames[f].map(fn) == self_bulk[f]);
    // the frames of the reverse-complement VIEW are the reverse frames of the sequence
    const auto rc_bulk = v.reverse_complement().translate_self_frames(table);
    for (size_t f = 0; f < rc_bulk.size(); f++) CHECK(self_frames.size() + f < bulk.size() && rc_bulk[f] == bulk[self_frames.size() + f]);
    // strands
    CHECK(v.reverse_complement().collect_bulk() == v.reverse_complement().collect());
    CHECK(v.collect_bulk() == dna);
    CHECK(v.complement().collect_bulk() == v.complement().collect());
    // to_fn over a slice of codons, and the same DnaSequence methods
    const std::vector<Codon<T>> codons = v.codons().collect();
    CHECK(translate_codons(table, codons) == v.codons().map(fn));
    const DnaSequence<T> seq(dna);
    const auto seq_frames = seq.translate_all_frames(table);
    CHECK(seq_frames.size() == bulk.size());
    for (size_t f = 0; f < bulk.size() && f < seq_frames.size(); f++) CHECK(seq_frames[f].as_slice() == bulk[f]);
    // a frame advanced by next() / next_back() translates what is left
    if (!frames.empty() && frames[0].len() >= 3) {
        CodonsView<T> part = frames[0];
        Codon<T> c;
        part.next(&c);
        part.next_back(&c);
        CHECK(part.translate(table) == part.map(fn));
        CodonsView<T> rpart = frames.back();
        rpart.next(&c);
        CHECK(rpart.translate(table) == rpart.map(fn));
    }
}

int main() {
    const Nucleotide strict[4] = {Nucleotide::A, Nucleotide::T, Nucleotide::C, Nucleotide::G};
    const TranslationTable tables[4] = {TranslationTable::Ncbi1, TranslationTable::Ncbi2, TranslationTable::Ncbi11, TranslationTable::Ncbi22};
    for (size_t n : {0, 1, 2, 3, 4, 5, 6, 8, 47, 48, 49, 50, 999, 1000, 1001, 30011, 300007}) {
        std::vector<Nucleotide> a(n);
        std::vector<NucleotideAmbiguous> b(n);
        for (size_t i = 0; i < n; i++) {
            a[i] = strict[rnd() & 3];
            b[i] = static_cast<NucleotideAmbiguous>(1 + rnd() % 15);
        }
        check_sequence(a, tables[n % 4]);
        check_sequence(b, tables[(n + 1) % 4]);
    }
    if (failures) {
        std::printf("%d checks FAILED\n", failures);
        return 1;
    }
    std::printf("all checks passed\n");
    return 0;
}

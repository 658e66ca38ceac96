// Host-only checks of the lazy adaptor mirror in include/quickdna.hpp (NucleotideView / CodonsView / to_fn) against
// the reference's doctests and unit tests for src/iter.rs (:21-74, :112-124, :138-147, :161-170, :186-220, :244-252,
// :455-462) and src/trans_table.rs:155-164.  Needs no GPU: the views never call a kernel.
// Run by tests/test_cpp_mirror.py (not gpu).
#include <cstdio>
#include <initializer_list>
#include <vector>

#include "quickdna.hpp"

using namespace quickdna;
using N = Nucleotide;
using NA = NucleotideAmbiguous;

static int failures = 0;
#define CHECK(cond)                                                     \
    do {                                                                \
        if (!(cond)) {                                                  \
            std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            failures++;                                                 \
        }                                                               \
    } while (0)

static Codon<N> cod(N a, N b, N c) { return Codon<N>{{a, b, c}}; }
static bool eq(const CodonsView<N>& f, std::initializer_list<Codon<N>> want) { return f.collect() == std::vector<Codon<N>>(want); }

int main() {
    const N A = N::A, T = N::T, C = N::C, G = N::G;
    const std::vector<N> dna = {C, G, A, T, C, G, A, T};
    const NucleotideView<N> v(dna);

    // all_reading_frames, iter.rs:21-74
    {
        auto f = v.all_reading_frames();
        CHECK(f.size() == 6);
        CHECK(eq(f[0], {cod(C, G, A), cod(T, C, G)}));
        CHECK(eq(f[1], {cod(G, A, T), cod(C, G, A)}));
        CHECK(eq(f[2], {cod(A, T, C), cod(G, A, T)}));
        CHECK(eq(f[3], {cod(A, T, C), cod(G, A, T)}));
        CHECK(eq(f[4], {cod(T, C, G), cod(A, T, C)}));
        CHECK(eq(f[5], {cod(C, G, A), cod(T, C, G)}));
        auto g = NucleotideView<N>(dna.data(), 4).all_reading_frames();  // third frames would hold 2 nucleotides
        CHECK(g.size() == 4);
        CHECK(eq(g[0], {cod(C, G, A)}));
        CHECK(eq(g[1], {cod(G, A, T)}));
        CHECK(eq(g[2], {cod(A, T, C)}));
        CHECK(eq(g[3], {cod(T, C, G)}));
        CHECK(NucleotideView<N>(dna.data(), 2).all_reading_frames().empty());
        CHECK(NucleotideView<N>(dna.data(), 3).all_reading_frames().size() == 2);
        CHECK(NucleotideView<N>(dna.data(), 0).all_reading_frames().empty());
    }
    // codons, iter.rs:112-124
    {
        const std::vector<N> d = {A, A, T, T, C, C, G, G};
        CHECK(eq(NucleotideView<N>(d).codons(), {cod(A, A, T), cod(T, C, C)}));
        // test_reverse_codons, iter.rs:455-462: `.codons().rev()` trims the incomplete codon first
        CHECK(NucleotideView<N>(d).codons().collect_rev() == (std::vector<Codon<N>>{cod(T, C, C), cod(A, A, T)}));
        CodonsView<N> it = NucleotideView<N>(d).codons();
        Codon<N> c{};
        CHECK(it.len() == 2 && it.next_back(&c) && c == cod(T, C, C));
        CHECK(it.len() == 1 && it.next(&c) && c == cod(A, A, T));
        CHECK(it.len() == 0 && !it.next(&c) && !it.next_back(&c));
    }
    // complement / reverse_complement, iter.rs:138-147, :161-170
    {
        const std::vector<N> d = {C, G, A, T};
        CHECK(NucleotideView<N>(d).complement().collect() == (std::vector<N>{G, C, T, A}));
        CHECK(NucleotideView<N>(d).reverse_complement().collect() == (std::vector<N>{A, T, C, G}));
        CHECK(NucleotideView<N>(d).reverse_complement().reverse_complement().collect() == d);
        const std::vector<NA> amb = {NA::A, NA::M, NA::R, NA::H, NA::V, NA::N, NA::W, NA::S};
        CHECK(NucleotideView<NA>(amb).reverse_complement().collect() == (std::vector<NA>{NA::S, NA::W, NA::N, NA::B, NA::D, NA::Y, NA::K, NA::T}));
    }
    // self_reading_frames, iter.rs:186-220
    {
        auto f = v.self_reading_frames();
        CHECK(f.size() == 3);
        CHECK(eq(f[0], {cod(C, G, A), cod(T, C, G)}));
        CHECK(eq(f[1], {cod(G, A, T), cod(C, G, A)}));
        CHECK(eq(f[2], {cod(A, T, C), cod(G, A, T)}));
        auto g = NucleotideView<N>(dna.data(), 4).self_reading_frames();
        CHECK(g.size() == 2);
        CHECK(eq(g[0], {cod(C, G, A)}));
        CHECK(eq(g[1], {cod(G, A, T)}));
        CHECK(NucleotideView<N>(dna.data(), 2).self_reading_frames().empty());
    }
    // trim_to_codon, iter.rs:244-252
    {
        CHECK(v.trimmed_to_codon().collect() == (std::vector<N>{C, G, A, T, C, G}));
        CHECK(v.reverse_complement().trimmed_to_codon().collect() == (std::vector<N>{A, T, C, G, A, T}));
        CHECK(v.skip(2).trimmed_to_codon().len() == 6 && v.skip(3).trimmed_to_codon().len() == 3);
    }
    // to_fn, trans_table.rs:155-164: ATCGATCG -> "ID"
    {
        const std::vector<N> d = {A, T, C, G, A, T, C, G};
        const auto ncbi1 = to_fn(TranslationTable::Ncbi1);
        CHECK(NucleotideView<N>(d).codons().map(ncbi1) == (std::vector<uint8_t>{'I', 'D'}));
        // ambiguous codons through the same closure: TTR -> L, TTV -> X, RAT -> B (rust_api.rs:472-482, README.md:31-37)
        const std::vector<NA> amb = {NA::T, NA::T, NA::R, NA::T, NA::T, NA::V, NA::R, NA::A, NA::T};
        CHECK(NucleotideView<NA>(amb).codons().map(ncbi1) == (std::vector<uint8_t>{'L', 'X', 'B'}));
        // table 22 reads TCA as a stop (README.md:64-71)
        const std::vector<N> tca = {T, C, A};
        CHECK(NucleotideView<N>(tca).codons().map(to_fn(TranslationTable::Ncbi22)) == (std::vector<uint8_t>{'*'}));
        CHECK(NucleotideView<N>(tca).codons().map(ncbi1) == (std::vector<uint8_t>{'S'}));
        CHECK(cod(T, C, A).to_string() == "TCA");
        // the six frames of AAAGGGAAA through the lazy path equal the bulk answer KGK,KG,RE,FPF,FP,SL (rust_api.rs:484-496)
        const std::vector<N> k = {A, A, A, G, G, G, A, A, A};
        const char* want[6] = {"KGK", "KG", "RE", "FPF", "FP", "SL"};
        auto frames = NucleotideView<N>(k).all_reading_frames();
        CHECK(frames.size() == 6);
        for (size_t i = 0; i < frames.size(); i++) {
            const auto aa = frames[i].map(ncbi1);
            CHECK(std::string(aa.begin(), aa.end()) == want[i]);
        }
    }
    if (failures == 0) std::printf("all checks passed\n");
    return failures ? 1 : 0;
}

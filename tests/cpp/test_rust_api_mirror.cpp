// Exercises include/quickdna.hpp (the C++ mirror of the reference's Rust API) through the C ABI on a GPU.
// The cases restate, as data, the reference's own Rust unit tests for this path
// (src/rust_api.rs:421-769, src/expansions.rs:208-305, src/canonical.rs:200-214 and :406-412, src/rust_api.rs:362-373,
// tests/integration_tests.rs:24-39, src/trans_table.rs:155-164).  Run by tests/test_cpp_mirror.py (-m gpu).
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "quickdna.hpp"

using namespace quickdna;

static int failures = 0;
#define CHECK(cond)                                                          \
    do {                                                                     \
        if (!(cond)) {                                                       \
            std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);      \
            failures++;                                                      \
        }                                                                    \
    } while (0)

static std::pair<size_t, size_t> lr(size_t a, size_t b) { return {a, b}; }
static DnaSequenceAmbiguous dna(const char* s) { return DnaSequenceAmbiguous::from_str(s); }
static DnaSequenceStrict dna_strict(const char* s) { return DnaSequenceStrict::from_str(s); }
static ProteinSequence protein(const char* s) { return ProteinSequence::from_str(s); }
static std::vector<ProteinSequence> proteins(std::initializer_list<const char*> l) {
    std::vector<ProteinSequence> v;
    for (const char* s : l) v.push_back(protein(s));
    return v;
}
template <class F>
static bool throws_translation_error(F f, int kind, uint8_t byte, const char* msg) {
    try {
        f();
    } catch (const TranslationError& e) {
        if (e.kind != kind || e.byte != byte || std::strcmp(e.what(), msg) != 0) {
            std::printf("  got kind %d byte %u '%s'\n", (int)e.kind, (unsigned)e.byte, e.what());
            return false;
        }
        return true;
    }
    return false;
}

int main() {
    const TranslationTable T1 = TranslationTable::Ncbi1;
    // accepted alphabets (src/rust_api.rs:421-457)
    for (int c = 0; c < 128; c++) {
        const std::string s(1, (char)c);
        bool ok = true, ok_strict = true;
        try { DnaSequenceAmbiguous::from_str(s); } catch (const TranslationError&) { ok = false; }
        try { DnaSequenceStrict::from_str(s); } catch (const TranslationError&) { ok_strict = false; }
        CHECK(ok == (c != 0 && std::strchr("aAtTcCgGmMrRwWsSyYkKvVhHdDbBnN \t", c) != nullptr));
        CHECK(ok_strict == (c != 0 && std::strchr("aAtTcCgG \t", c) != nullptr));
    }
    // translate (src/rust_api.rs:460-482)
    CHECK(dna("AAAGGGAAA").translate(T1) == protein("KGK"));
    CHECK(dna_strict("AAAGGGAAA").translate(T1) == protein("KGK"));
    CHECK(dna("TTR TTV").translate(T1) == protein("LX"));
    // self frames incl. short inputs (src/rust_api.rs:484-532)
    CHECK(dna("AAAGGGAAA").translate_self_frames(T1) == proteins({"KGK", "KG", "RE"}));
    CHECK(dna_strict("AAAGGGAAA").translate_self_frames(T1) == proteins({"KGK", "KG", "RE"}));
    CHECK(dna("GGGG").translate_self_frames(T1) == proteins({"G", "G"}));
    CHECK(dna("GGG").translate_self_frames(T1) == proteins({"G"}));
    CHECK(dna("GG").translate_self_frames(T1).empty());
    CHECK(dna("G").translate_self_frames(T1).empty());
    CHECK(dna("").translate_self_frames(T1).empty());
    // all frames (src/rust_api.rs:534-632)
    CHECK(dna("AAAGGGAAA").translate_all_frames(T1) == proteins({"KGK", "KG", "RE", "FPF", "FP", "SL"}));
    CHECK(dna_strict("AAAGGGAAA").translate_all_frames(T1) == proteins({"KGK", "KG", "RE", "FPF", "FP", "SL"}));
    CHECK(dna("GGGG").translate_all_frames(T1) == proteins({"G", "G", "P", "P"}));
    CHECK(dna("GGG").translate_all_frames(T1) == proteins({"G", "P"}));
    CHECK(dna("GG").translate_all_frames(T1).empty());
    CHECK(dna("").translate_all_frames(T1).empty());
    // README six-frame example
    CHECK(dna("taatcaagactattcaaccaa").translate_all_frames(T1) == proteins({"*SRLFNQ", "NQDYST", "IKTIQP", "LVE*S*L", "WLNSLD", "G*IVLI"}));
    CHECK(dna("taatcaagactattcaaccaa").translate(TranslationTable::Ncbi22) == protein("**RLFNQ"));
    // reverse complement (tests/integration_tests.rs:24-39; src/iter.rs:163-170)
    CHECK(dna("CATTAG").reverse_complement().translate(T1) == protein("LM"));
    CHECK(dna_strict("CATTAG").reverse_complement().translate(T1) == protein("LM"));
    CHECK(dna("CGAT").reverse_complement() == dna("ATCG"));
    CHECK(dna("AAABBBGGG").reverse_complement().to_string() == "CCCVVVTTT");
    // bytes API + errors (src/python_api.rs:57; src/fasta.rs:1156-1169, 1230-1288, 1388-1396)
    {
        auto v = reverse_complement_bytes<NucleotideAmbiguous>("AAAAABCCC");
        CHECK(std::string(v.begin(), v.end()) == "GGGVTTTTT");
        auto t = translate_dna_bytes<NucleotideAmbiguous>(T1, "atcgatcg");
        CHECK(std::string(t.begin(), t.end()) == "ID");
    }
    CHECK(throws_translation_error([] { dna_strict("ABCD"); }, TranslationError::UnexpectedAmbiguousNucleotide, 'B', "unexpected ambiguous nucleotide: 'B'"));
    CHECK(throws_translation_error([] { dna("AAAelephant"); }, TranslationError::BadNucleotide, 'e', "bad nucleotide: 'e'"));
    CHECK(throws_translation_error([] { dna("AA\xc4\x8d" "CCG"); }, TranslationError::NonAsciiByte, 0xc4, "non-ascii byte: c4"));
    CHECK(throws_translation_error([] { reverse_complement_bytes<Nucleotide>("AAABBBGGG"); }, TranslationError::UnexpectedAmbiguousNucleotide, 'B',
                                   "unexpected ambiguous nucleotide: 'B'"));
    CHECK(throws_translation_error([] { table_try_from(17); }, TranslationError::BadTranslationTable, 17, "not a ncbi translation table: 17"));
    CHECK(table_try_from(8) == TranslationTable::Ncbi8);
    // windows (src/rust_api.rs:702-769)
    {
        const char* want[11] = {"gcantaccta", "cantacctaa", "antacctaan", "ntacctaang", "tacctaangt", "acctaangtn",
                                "cctaangtna", "ctaangtnat", "taangtnatt", "aangtnatta", "angtnattag"};
        auto w = dna("gcantacctaangtnattag").windows(10);
        auto pw = protein("gcantacctaangtnattag").windows(10);
        CHECK(w.size() == 11 && pw.size() == 11);
        for (size_t i = 0; i < w.size() && i < 11; i++) {
            CHECK(w[i] == dna(want[i]));
            CHECK(pw[i] == protein(want[i]));
        }
        CHECK(dna("antg").windows(10).empty());
        CHECK(dna_strict("actg").windows(10).empty());
        CHECK(protein("antg").windows(10).empty());
    }
    // whitespace (src/rust_api.rs:771-785) and as_ref (:787-792)
    CHECK(dna("  gcantac\tctaangtnattag ") == dna("gcantacctaangtnattag"));
    CHECK(dna_strict(" gca ctac ctaacgtcattag \t") == dna_strict("gcactacctaacgtcattag"));
    CHECK(dna("ac").as_slice() == (std::vector<NucleotideAmbiguous>{NucleotideAmbiguous::A, NucleotideAmbiguous::C}));
    // canonical (src/rust_api.rs:362-373; src/canonical.rs:200-214, 406-412)
    CHECK(dna_strict("CATTAG").canonical() == dna_strict("ATCCTG"));
    CHECK(dna_strict("TTGT").canonical() == dna_strict("AATA"));
    CHECK(dna_strict("TGTT").canonical() == dna_strict("AATA"));
    CHECK(dna_strict("ATCGCCAT").canonical() == dna_strict("ATCCGCAT"));
    CHECK(dna_strict("").canonical() == dna_strict(""));
    // canonical window keys (not in the reference: SURVEY 8(f) rank 3) agree with windows(w) + canonical()
    {
        const auto x = dna_strict("TGCGAGTGTAGCGAGATGTAGCGTAGAGTCTGAGATGCAGTACATTAGCATTAG");
        for (size_t w : {size_t(6), size_t(33), size_t(42)}) {
            const auto wins = x.windows(w);
            const auto fp = x.canonical_window_keys(w);
            const auto packed = x.canonical_window_keys(w, QD_CANON_KEY_PACKED);
            const size_t pw = (w + 31) / 32;
            CHECK(fp.size() == wins.size() && packed.size() == wins.size() * pw);
            for (size_t i = 0; i < wins.size() && i < fp.size(); i++) {
                const auto c = wins[i].canonical();
                uint32_t words[4] = {0, 0, 0, 0};  // lo plane words, then hi plane words
                for (size_t j = 0; j < w; j++) {
                    const unsigned code = c[j] == Nucleotide::A ? 0 : c[j] == Nucleotide::T ? 1 : c[j] == Nucleotide::C ? 2 : 3;
                    words[j / 32] |= (code & 1u) << (j % 32);
                    words[pw + j / 32] |= (code >> 1) << (j % 32);
                }
                CHECK(fp[i] == qd_canonical_fingerprint(words, w));
                CHECK(std::memcmp(&packed[i * pw], words, 8 * pw) == 0);
            }
        }
        CHECK(dna_strict("ACG").canonical_window_keys(42).empty());
    }
    {
        const DnaSequenceStrict x = dna_strict("GATTACAGATTACACATTAGGATCCAGTTTACGATCAGGACT"), rc = x.reverse_complement();
        CHECK(x.canonical() == rc.reverse_complement().canonical());
        CHECK(x.canonical().canonical() == x.canonical());
        std::vector<Nucleotide> rev(x.as_slice().rbegin(), x.as_slice().rend());
        CHECK(DnaSequenceStrict(rev).canonical() == x.canonical());
    }
    // expansions (src/expansions.rs:208-233, 283-305)
    {
        auto e = dna("ATBCGYAC").expansions();
        const char* want[6] = {"ATTCGTAC", "ATTCGCAC", "ATCCGTAC", "ATCCGCAC", "ATGCGTAC", "ATGCGCAC"};
        CHECK(e.size() == 6 && dna("ATBCGYAC").expansion_count() == 6);
        for (size_t i = 0; i < e.size() && i < 6; i++) CHECK(e[i] == dna_strict(want[i]));
        auto w = dna("WWW").expansions();
        const char* want_w[8] = {"AAA", "AAT", "ATA", "ATT", "TAA", "TAT", "TTA", "TTT"};
        CHECK(w.size() == 8);
        for (size_t i = 0; i < w.size() && i < 8; i++) CHECK(w[i] == dna_strict(want_w[i]));
        auto n4 = dna("ATCGNGCTA").expansions();
        const char* want_n[4] = {"ATCGAGCTA", "ATCGTGCTA", "ATCGCGCTA", "ATCGGGCTA"};
        CHECK(n4.size() == 4);
        for (size_t i = 0; i < n4.size() && i < 4; i++) CHECK(n4[i] == dna_strict(want_n[i]));
        CHECK(dna("NNNNNNNNNNNNNNNNNNBBBBBBBBBBBBBBBBBY").expansion_count() == 17748888853923495936ull);
        CHECK(dna("NNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNN").expansion_count() == UINT64_MAX);
        CHECK(dna("ACGT").expansions().size() == 1 && dna("ACGT").expansions()[0] == dna_strict("ACGT"));
    }
    // a long sequence end to end: six frames of x == three frames of x + three frames of revcomp(x)
    {
        std::string s;
        uint64_t z = 12345;
        for (int i = 0; i < 100003; i++) {
            z = z * 6364136223846793005ull + 1442695040888963407ull;
            s.push_back("ACGTNRY"[(z >> 33) % 7]);
        }
        const DnaSequenceAmbiguous x = dna(s.c_str());
        auto all = x.translate_all_frames(TranslationTable::Ncbi11);
        auto fwd = x.translate_self_frames(TranslationTable::Ncbi11), rev = x.reverse_complement().translate_self_frames(TranslationTable::Ncbi11);
        CHECK(all.size() == 6 && fwd.size() == 3 && rev.size() == 3);
        for (int k = 0; k < 3 && all.size() == 6; k++) {
            CHECK(all[k] == fwd[k]);
            CHECK(all[3 + k] == rev[k]);
        }
        CHECK(x.reverse_complement().reverse_complement() == x);
    }
    // FASTA (src/fasta.rs:1131-1210, 1342-1396)
    {
        auto recs = FastaParser<Nucleotide>().parse_str(">Virus1\nAAAA\nAAAA\n>Virus2\nCCCC\nCCCC\n");
        CHECK(recs.size() == 2);
        if (recs.size() == 2) {
            CHECK(recs[0].header == "Virus1" && recs[0].contents == dna_strict("AAAAAAAA") && recs[0].line_range == lr(1, 4));
            CHECK(recs[1].header == "Virus2" && recs[1].contents == dna_strict("CCCCCCCC") && recs[1].line_range == lr(4, 7));
        }
        auto many = FastaParser<Nucleotide>().parse_str(">Virus1\nAC\nT\n>Empty\n\n>Virus2\n>with many\n>comment lines\nC  AT");
        CHECK(many.size() == 3);
        if (many.size() == 3) {
            CHECK(many[1].header == "Empty" && many[1].contents.is_empty());
            CHECK(many[2].header == "Virus2\nwith many\ncomment lines" && many[2].contents == dna_strict("CAT"));
        }
        auto amb = FastaParser<NucleotideAmbiguous>().parse_str(">Virus1\n  AAAA\tBCD \t\n");
        CHECK(amb.size() == 1 && amb[0].contents == dna("AAAABCD"));
        auto noconcat = FastaParser<Nucleotide>(FastaParseSettings{false, false}).parse_str(">a\n>b\ntcga");
        CHECK(noconcat.size() == 2 && noconcat[0].header == "a" && noconcat[0].contents.is_empty() && noconcat[1].contents == dna_strict("TCGA"));
        auto pre = FastaParser<Nucleotide>().parse_str("ACGT\nGGCC\n\n>Virus\n\n");
        CHECK(pre.size() == 2 && pre[0].header.empty() && pre[0].contents == dna_strict("ACGTGGCC") && pre[0].line_range == lr(1, 4));
        bool threw = false;
        try {
            FastaParser<Nucleotide>().parse_str(">Virus1\nAAA\nCCCxGGG");
        } catch (const FastaParseError& e) {
            threw = e.line_number == 3 && std::string(e.what()) == "on line 3: error parsing record: bad nucleotide: 'x'";
        }
        CHECK(threw);
        // more records than the first capacity guess (one per 64 bytes of text): the exactly sized second call
        std::string tiny;
        for (int i = 0; i < 300; i++) tiny += ">" + std::to_string(i) + "\nAC\n";
        auto many_tiny = FastaParser<Nucleotide>().parse_str(tiny);
        CHECK(many_tiny.size() == 300 && many_tiny[299].header == "299" && many_tiny[7].contents == dna_strict("AC") &&
              many_tiny[299].line_range == lr(599, 601));
    }
    if (failures) {
        std::printf("%d check(s) failed\n", failures);
        return 1;
    }
    std::printf("rust-api mirror: all checks passed\n");
    return 0;
}

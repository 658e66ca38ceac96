// Latency of the tiny host calls from C (no Python in the way): g++ -O2 -I include tools/latency_c.cpp -o tools/latency_c -L quickdna_b200 -lquickdna_b200 -Wl,-rpath,$PWD/quickdna_b200
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "quickdna_b200.h"

int main() {
    qd_ctx* c = nullptr;
    if (qd_ctx_create(&c, 0) != 0) return 1;
    const size_t n = 29901;
    std::vector<uint8_t> dna(n), out(3 * n);
    uint64_t x = 88172645463325252ull;
    for (size_t i = 0; i < n; i++) {
        x ^= x << 13, x ^= x >> 7, x ^= x << 17;
        dna[i] = "ACGT"[x & 3];
    }
    qd_err err;
    size_t lens[6];
    int nf;
    auto time_it = [&](const char* name, auto fn) {
        for (int i = 0; i < 50; i++) fn();
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < 2000; i++) fn();
        double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / 2000;
        printf("%-28s %7.2f us per call\n", name, us);
    };
    time_it("qd_translate", [&] { qd_translate(c, 1, QD_ENC_ASCII, 0, dna.data(), n, out.data(), &err); });
    time_it("qd_revcomp", [&] { qd_revcomp(c, QD_ENC_ASCII, 0, dna.data(), n, out.data(), &err); });
    time_it("qd_translate_frames (6)", [&] { qd_translate_frames(c, 1, QD_ENC_ASCII, 0, QD_FRAMES_ALL, dna.data(), n, out.data(), lens, &nf, &err); });
    size_t k;
    time_it("qd_parse", [&] { qd_parse(c, 0, dna.data(), n, out.data(), &k, &err); });
    // the window calls on the same 30-kb genome
    std::vector<uint8_t> rows((n - 41) * 42 + 64);
    std::vector<uint64_t> keys(2 * (n - 41) + 8);
    size_t nw = 0;
    time_it("qd_canonical_windows (42)", [&] { qd_canonical_windows(c, QD_ENC_ASCII, 0, dna.data(), n, 42, rows.data(), &err); });
    time_it("qd_canonical_window_keys", [&] { qd_canonical_window_keys(c, QD_ENC_ASCII, 0, 0, dna.data(), n, 42, keys.data(), &err); });
    time_it("qd_canonical", [&] { qd_canonical(c, QD_ENC_ASCII, 0, dna.data(), n, out.data(), &err); });
    time_it("qd_windows (42)", [&] { qd_windows(c, dna.data(), n, 42, rows.data(), &nw); });
    {
        const size_t n_win = 500;
        std::vector<uint8_t> win(n_win * 42), ex(n_win * 16 * 42);
        std::vector<uint64_t> counts(n_win), index(n_win + 1);
        for (size_t i = 0; i < n_win * 42; i++) win[i] = dna[i];
        for (size_t i = 0; i < n_win; i++) win[i * 42 + 7] = 'N', win[i * 42 + 30] = 'R';
        time_it("qd_expansion_windows (500)", [&] { qd_expansion_windows(c, QD_ENC_ASCII, win.data(), n_win, 42, 16, counts.data(), index.data(), ex.data(), &err); });
    }
    {
        const size_t n_seq = 100, len = 299;
        std::vector<uint64_t> offs(n_seq + 1), oo2(n_seq + 1);
        for (size_t i = 0; i <= n_seq; i++) offs[i] = i * len;
        std::vector<uint8_t> fr(qd_batch_out_bytes(offs.data(), n_seq, 6) + 64);
        time_it("qd_translate_frames_batch", [&] { qd_translate_frames_batch(c, 1, QD_ENC_ASCII, 0, QD_FRAMES_ALL, dna.data(), offs.data(), n_seq, fr.data(), oo2.data(), &err); });
        time_it("qd_translate_frames_batch_u", [&] { qd_translate_frames_batch_uniform(c, 1, QD_ENC_ASCII, 0, QD_FRAMES_ALL, dna.data(), n_seq, len, fr.data(), &err); });
    }
    // a 30-kb FASTA file: one record, 70-column lines
    std::vector<uint8_t> text;
    for (const char* h = ">covid\n"; *h; h++) text.push_back((uint8_t)*h);
    for (size_t i = 0; i < n; i++) {
        text.push_back(dna[i]);
        if (i % 70 == 69 || i + 1 == n) text.push_back('\n');
    }
    std::vector<uint8_t> masks(text.size());
    std::vector<qd_fasta_record> recs(16);
    std::vector<uint64_t> oo(17);
    std::vector<uint8_t> frames(qd_slots_bytes(n, 6) + 4096);
    size_t n_masks = 0, n_recs = 0, out_bytes = 0;
    time_it("qd_fasta_scan", [&] { qd_fasta_scan(c, 0, 1, 0, text.data(), text.size(), masks.data(), &n_masks, recs.data(), recs.size(), &n_recs, &err); });
    time_it("qd_fasta_translate_frames", [&] {
        qd_fasta_translate_frames(c, 1, 0, QD_FRAMES_ALL, 1, 0, text.data(), text.size(), recs.data(), recs.size(), &n_recs, frames.data(), frames.size(),
                                  oo.data(), &out_bytes, &err);
    });
    printf("records %zu masks %zu frame bytes %zu\n", n_recs, n_masks, out_bytes);
    qd_ctx_destroy(c);
    return 0;
}

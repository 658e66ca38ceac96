// Kernel-only harness for tuning: runs translate_frames_kernel<KB_FRAMES, KB_THREADS, KB_DIRECT> on a uniform batch and revcomp_kernel on one
// sequence, device-resident, CUDA-event timed.  Not product code, not a parity check (tests/ does that) --
// it exists so kernel variants (compile-time -D knobs) can be compared back to back on ONE box.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/kbench tools/kbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <algorithm>
#include <vector>

#include "../quickdna_b200/csrc/qd_kernels.cuh"
#include "../quickdna_b200/csrc/qd_tables.hpp"

using namespace qd;

#ifndef KB_RC_MISALIGN
#define KB_RC_MISALIGN 0
#endif
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void synth_kernel(uint8_t* out, size_t n, uint64_t seed) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t z = seed + i + 0x9e3779b97f4a7c15ull;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        z ^= z >> 31;
        out[i] = "ACGT"[z & 3];
    }
}
__global__ void xorsum_kernel(const uint32_t* p, size_t n, unsigned long long* acc) {
    unsigned long long s = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s += p[i] * (unsigned long long)(i % 1000003 + 1);
    atomicAdd(acc, s);
}

int main(int argc, char** argv) {
    const size_t n_reads = argc > 1 ? strtoull(argv[1], 0, 10) : 4000000;
    const size_t L = argc > 2 ? strtoull(argv[2], 0, 10) : 1000;
    const int reps = argc > 3 ? atoi(argv[3]) : 10;
    const size_t n_rc = 250000000;
    static HostTables T;
    build_host_tables(T);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const size_t total = n_reads * L, runs = (L + RUN_NT - 1) / RUN_NT, out_bytes = n_reads * runs * 16 * 6;
    uint8_t *dna, *out, *d_decode;
    uint16_t *d_lut, *d_pair;
    unsigned long long *d_err, *d_acc;
    CK(cudaMalloc(&dna, total + 64));
    CK(cudaMalloc(&out, out_bytes + 64));
    CK(cudaMalloc(&d_lut, LUT_ENTRIES * 2));
    CK(cudaMalloc(&d_pair, sizeof(T.pair)));
    CK(cudaMalloc(&d_decode, 256));
    CK(cudaMalloc(&d_err, 8));
    CK(cudaMalloc(&d_acc, 8));
    CK(cudaMemcpy(d_lut, T.folded.data(), LUT_ENTRIES * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_pair, T.pair, sizeof(T.pair), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_decode, T.decode_ascii, 256, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_err, 0xff, 8));
    CK(cudaMemset(d_acc, 0, 8));
    synth_kernel<<<sms * 8, 256>>>(dna, total, 3);
    CK(cudaDeviceSynchronize());

    FramesParams p{};
    p.dna = dna;
    p.total_nt = total;
    p.n_seq = n_reads;
    p.uniform_len = L;
    p.uniform_runs = (uint32_t)runs;
    p.uniform_magic = runs == 1 ? 0 : (~0ull / runs) + 1;
    p.uniform_mod3 = (uint32_t)(L % 3);
    p.total_runs = runs * n_reads;
#ifndef KB_THREADS
#define KB_THREADS 1024
#endif
#ifndef KB_FRAMES
#define KB_FRAMES 6
#endif
    constexpr int TILE_THREADS = KB_THREADS, TILE_RUNS = TileGeom<KB_THREADS, frames_decoupled(KB_FRAMES)>::RUNS;
    constexpr bool KB_DIRECT = KB_FRAMES == 1;
    if (KB_DIRECT) {
        std::vector<uint8_t> host;
        build_direct_table(T, 0, true, false, host);
        uint8_t* d_direct;
        CK(cudaMalloc(&d_direct, DIRECT_ENTRIES));
        CK(cudaMemcpy(d_direct, host.data(), DIRECT_ENTRIES, cudaMemcpyHostToDevice));
        p.direct = d_direct;
        p.direct_k1 = DIRECT_K1_ASCII;
        p.direct_k2 = DIRECT_K2_ASCII;
    }
    p.n_tiles = (uint32_t)((p.total_runs + TILE_RUNS - 1) / TILE_RUNS);
#ifndef KB_NO_ALIGNED_TILES
    if (TILE_RUNS / runs >= 1 && (TILE_RUNS / runs) * runs * 16 >= (size_t)TILE_RUNS * 15) {  // whole sequences per tile, as frames_uniform_dev does
        p.seqs_per_tile = (uint32_t)(TILE_RUNS / runs);
        p.tile_magic = (uint32_t)(((1u << 20) + runs - 1) / runs);
        p.n_tiles = (uint32_t)((n_reads + p.seqs_per_tile - 1) / p.seqs_per_tile);
    }
#endif
    p.out = out;
    p.lut = d_lut;
    p.decode = d_decode;
    p.pair = d_pair;
    p.pair_coef = 2u | ((2u * PAIR_K1_ASCII) << 16);
    p.and_need = 0x40404040u;
    p.or_forbid = 0x80808080u;
    p.err_key = d_err;
    int occ = 0;
    using Smem = FramesSmem<KB_THREADS, frames_stages(KB_FRAMES)>;
    const size_t smem = KB_DIRECT ? ((sizeof(Smem) + 127) & ~(size_t)127) + DIRECT_ENTRIES : sizeof(Smem);
    CK(cudaFuncSetAttribute(translate_frames_kernel<KB_FRAMES, KB_THREADS, KB_DIRECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, translate_frames_kernel<KB_FRAMES, KB_THREADS, KB_DIRECT>, TILE_THREADS, smem));
#ifdef KB_OCC
    occ = KB_OCC;
#endif
#ifdef KB_GRID
    const unsigned grid = KB_GRID;
#else
    const unsigned grid = (unsigned)(sms * occ);
#endif
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; i++) translate_frames_kernel<KB_FRAMES, KB_THREADS, KB_DIRECT><<<grid, TILE_THREADS, smem>>>(p);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; i++) translate_frames_kernel<KB_FRAMES, KB_THREADS, KB_DIRECT><<<grid, TILE_THREADS, smem>>>(p);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    if (getenv("KB_DISTURB")) {
        // what does the kernel do after the GPU has idled / run something else?  per-CTA start, end and SM of one launch each
        unsigned long long* d_trace;
        CK(cudaMalloc(&d_trace, grid * 64));
        std::vector<unsigned long long> h(grid * 8);
        auto traced = [&](const char* label) {
            p.trace = d_trace;
            CK(cudaMemset(d_trace, 0, grid * 64));
            CK(cudaEventRecord(e0));
            translate_frames_kernel<KB_FRAMES, KB_THREADS, KB_DIRECT><<<grid, TILE_THREADS, smem>>>(p);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float t;
            CK(cudaEventElapsedTime(&t, e0, e1));
            CK(cudaMemcpy(h.data(), d_trace, grid * 64, cudaMemcpyDeviceToHost));
            unsigned long long t0 = ~0ull, t1 = 0, late = 0, dur_min = ~0ull, dur_max = 0;
            std::vector<int> per_sm(256, 0);
            for (unsigned b = 0; b < grid; b++) t0 = std::min(t0, h[8 * b]);
            for (unsigned b = 0; b < grid; b++) {
                t1 = std::max(t1, h[8 * b + 1]);
                late += h[8 * b] - t0 > 100000;  // started more than 100 us after the first CTA
                dur_min = std::min(dur_min, h[8 * b + 1] - h[8 * b]);
                dur_max = std::max(dur_max, h[8 * b + 1] - h[8 * b]);
                per_sm[h[8 * b + 2] & 255]++;
            }
            int sms_used = 0, max_per_sm = 0;
            for (int v : per_sm) sms_used += v > 0, max_per_sm = std::max(max_per_sm, v);
            printf("  %-28s %.3f ms  span %.3f ms  CTA time %.3f..%.3f ms  CTAs started late %llu  SMs used %d  max CTAs on one SM %d\n", label, t,
                   (t1 - t0) * 1e-6, dur_min * 1e-6, dur_max * 1e-6, late, sms_used, max_per_sm);
#ifdef QD_TRACE_DETAIL
            for (unsigned b : {0u, 1u, 2u, 3u})
                printf("    block %u sm %llu: %.3f ms, producer waited for empty %.3f ms, staging %.3f ms; consumer warp 0 waited for full %.3f ms, warp 16 %.3f ms\n", b,
                       h[8 * b + 2], (h[8 * b + 1] - h[8 * b]) * 1e-6, h[8 * b + 4] * 1e-6, h[8 * b + 5] * 1e-6, h[8 * b + 6] * 1e-6, h[8 * b + 7] * 1e-6);
#endif
            if (getenv("KB_DISTURB_DETAIL") && dur_max > 1.5 * dur_min) {
                printf("    slow CTAs (block:sm:ms):");
                for (unsigned b = 0; b < grid; b++)
                    if (h[8 * b + 1] - h[8 * b] > 1.3 * dur_min) printf(" %u:%llu:%.2f", b, h[8 * b + 2], (h[8 * b + 1] - h[8 * b]) * 1e-6);
                printf("\n");
            }
            p.trace = nullptr;
        };
        traced("back to back");
        traced("again");
        usleep(500000);
        traced("after 0.5 s idle");
        traced("again");
        void* big;
        CK(cudaMalloc(&big, size_t(1) << 30));
        CK(cudaMemset(big, 1, size_t(1) << 30));
        traced("after cudaMemset 1 GB");
        traced("again");
    }
    size_t algo = L;
    for (int k = 0; k < (KB_FRAMES == 1 ? 1 : 3); k++) algo += (KB_FRAMES == 6 ? 2 : 1) * ((L - k) / 3);
    xorsum_kernel<<<sms * 8, 256>>>((const uint32_t*)out, out_bytes / 4, d_acc);
    unsigned long long acc = 0, err = 0;
    CK(cudaMemcpy(&acc, d_acc, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&err, d_err, 8, cudaMemcpyDeviceToHost));
    printf("frames%d thr %d occ %d grid %u  %zu x %zu  %.3f ms  %.1f GB/s algorithmic  %.3f Tnt/s  checksum %016llx err %llx\n", KB_FRAMES, KB_THREADS, occ, grid, n_reads, L,
           ms, n_reads * (double)algo / ms / 1e6, total / (double)ms / 1e9, acc, err);

    // reverse complement, rotating 4 buffer sets inside the big allocations
    if (total >= 4 * n_rc && out_bytes >= 4 * n_rc) {
        uint8_t* d_tab;
        uint16_t* d_rcp;
        CK(cudaMalloc(&d_tab, 256));
        CK(cudaMalloc(&d_rcp, PAIR_ENTRIES * 2));
        CK(cudaMemcpy(d_tab, T.rc_ascii[0], 256, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_rcp, T.rc_pair[RC_PAIR_ASCII], PAIR_ENTRIES * 2, cudaMemcpyHostToDevice));
        int occ_rc = 0;
        constexpr int RCT = RC_BIG_THREADS;
        const size_t rc_smem = sizeof(RevcompSmem<RCT>);
        CK(cudaFuncSetAttribute(revcomp_kernel<RCT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rc_smem));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_rc, revcomp_kernel<RCT>, RCT, rc_smem));
#ifdef KB_RC_OCC
        occ_rc = KB_RC_OCC;
#endif
        const uint64_t tiles = ((n_rc >> 4) + RevcompSmem<RCT>::TILE_CHUNKS - 1) / RevcompSmem<RCT>::TILE_CHUNKS;
        const unsigned g2 = (unsigned)std::min<uint64_t>(tiles, (uint64_t)sms * occ_rc);
        auto mk = [&](int i) {
            RevcompParams q{};
            q.dna = dna + (size_t)i * n_rc + KB_RC_MISALIGN;
            q.n = n_rc - 16;
            q.out = out + (size_t)i * n_rc;
            q.table = d_tab;
            q.pair = d_rcp;
            q.pair_coef = 2u | ((2u * PAIR_K1_ASCII) << 16);
            q.and_need = 0x40404040u;
            q.or_forbid = 0x80808080u;
            q.filler = 0x41u;
            q.err_key = d_err;
            return q;
        };
        for (int i = 0; i < 4; i++) revcomp_kernel<RCT><<<g2, RCT, rc_smem>>>(mk(i));
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int i = 0; i < 20; i++) {
            revcomp_kernel<RCT><<<g2, RCT, rc_smem>>>(mk(i & 3));
        }
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= 20;
        CK(cudaMemset(d_acc, 0, 8));
        xorsum_kernel<<<sms * 8, 256>>>((const uint32_t*)out, n_rc / 4, d_acc);
        CK(cudaMemcpy(&acc, d_acc, 8, cudaMemcpyDeviceToHost));
        printf("revcomp  thr %d unroll %d occ %d grid %u  %zu nt  %.4f ms  %.1f GB/s algorithmic  checksum %016llx\n", RCT, RC_UNROLL, occ_rc, g2, n_rc, ms, 2.0 * n_rc / ms / 1e6, acc);
    }
    // (canonical windows / keys are timed through the library: tools/canon_keys_bench.py)
    CK(cudaDeviceSynchronize());
    return 0;
}

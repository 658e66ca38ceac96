// quickdna_b200 -- sm_100a kernels for the quickdna sequence hot path.
//
// All kernels are HBM-bound byte/integer work (no tensor cores).  Design, data layout and the
// algorithmic bytes per unit are described in DESIGN.md; reference citations are relative to
// the SecureDNA/quickdna checkout.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace qd {

// ------------------------------------------------------------------------------------------------
// Geometry
// ------------------------------------------------------------------------------------------------
constexpr int RUN_NT = 48;        // nucleotides per run: 16 residues per frame = one 128-bit store
constexpr int TILE_RUNS = 256;    // runs per tile = threads per CTA (one run per thread in phase 2)
constexpr int TILE_THREADS = 256;
constexpr int TILE_MAX_BYTES = TILE_RUNS * RUN_NT + 2 + 15 + 15;  // span + halo + alignment slack
constexpr int TILE_MAX_CHUNKS = (TILE_MAX_BYTES + 15) / 16 + 1;
constexpr int NIB_WORDS = TILE_MAX_CHUNKS * 2 + 16;  // 8 nibble codes per 32-bit word, + window slack
constexpr int RAW_BYTES = TILE_MAX_CHUNKS * 16;      // one staged tile of raw input bytes
constexpr int SEARCH_CAP = 512;   // reads of one tile whose run prefix is cached in smem

// Internal 4-bit nucleotide code (NOT the reference's mask): concrete bases are 0..3 in the
// reference's A<T<C<G order, 4..14 are the IUPAC ambiguity codes, 15 = invalid byte / past the end.
constexpr int CODE_INVALID = 15;

// LUT index of a codon (c0 = first nucleotide): idx = c0 + 16*c1 + 244*c2.  It is injective on
// codes 0..14, a codon whose LAST nucleotide is 15 (anything past the end of a sequence) lands in
// [3660, 3916) where every entry is 0, and -- found by exhaustive search -- the 64 concrete codons
// occupy 32 words in 32 distinct shared-memory banks, so random ACGT codons never conflict.
// The kernel gets 2*idx (the byte offset of a 16-bit entry) from ONE dp4a over four consecutive
// doubled codes with coefficients (1, 16, 244, 0).  Entry = forward aa | (reverse-strand aa << 8);
// the reverse aa is LUT[comp(c2), comp(c1), comp(c0)] of the reference table.
constexpr uint32_t LUT_K1 = 16, LUT_K2 = 244;
constexpr uint32_t LUT_ENTRIES = 4096;  // 15 + 16*15 + 244*15 = 3915 is the largest index
constexpr uint32_t LUT_DP4A_COEF = 1u | (LUT_K1 << 8) | (LUT_K2 << 16);
__host__ __device__ __forceinline__ uint32_t lut_index(uint32_t c0, uint32_t c1, uint32_t c2) {
    return c0 + LUT_K1 * c1 + LUT_K2 * c2;
}

// ------------------------------------------------------------------------------------------------
// Small PTX helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg_cached(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(void* p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// mbarrier + TMA bulk copy (global -> shared::cta), used by the TMA-staged tile loader.
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// Error key: the lowest flat input offset of an offending byte, kept with atomicMin.
// ------------------------------------------------------------------------------------------------
constexpr unsigned long long NO_ERROR_KEY = ~0ull;

// ------------------------------------------------------------------------------------------------
// Frame translation (1, 3 or 6 frames from ONE read of the input)
// ------------------------------------------------------------------------------------------------
// Where a tile starts: flat offset of its first run, and the sequence that run belongs to.
struct TileInfo {
    uint64_t flat;
    uint32_t first_seq;
    uint32_t pad_;
};

struct FramesParams {
    const uint8_t* dna;          // flat input (all sequences back to back)
    uint64_t total_nt;           // bytes in dna
    const uint64_t* offsets;     // CSR offsets [n_seq+1] or nullptr (uniform batch / single sequence)
    const uint64_t* run_prefix;  // exclusive prefix of ceil(len/48) [n_seq+1] or nullptr (uniform)
    const TileInfo* tiles;       // per tile (variable mode), or nullptr (uniform: closed form)
    uint64_t n_seq;
    uint64_t uniform_len;        // sequence length in uniform mode
    uint64_t uniform_magic;      // ceil(2^64 / uniform_runs): g / uniform_runs == umul64hi(g, magic) for g < 2^32
    uint32_t uniform_runs;       // ceil(uniform_len / 48)
    uint32_t uniform_mod3;       // uniform_len % 3
    uint64_t total_runs;         // uniform mode; variable mode reads run_prefix[n_seq]
    uint32_t n_tiles;            // upper bound on tiles (tiles past total_runs exit)
    uint8_t* out;                // slot layout, 16-byte aligned
    const uint16_t* lut;         // LUT_ENTRIES entries of the selected table, lut_index() layout
    const uint8_t* decode;       // 256-byte ASCII (or mask) -> internal code table
    unsigned long long* err_key;
    uint64_t flat_base;          // value of offsets[] that corresponds to dna[0] (CSR chunks)
    uint64_t key_base;           // added to a flat offset to form the error key (chunked host calls)
    int strict;
};

struct RunDesc {
    uint64_t flat_start;  // flat offset of the run's first nucleotide
    uint64_t out_base;    // byte offset of the sequence's slot block in `out`
    uint64_t left_nt;     // nucleotides from the run start to the end of its sequence
    uint32_t m;           // run index within the sequence
    uint32_t runs;        // R = ceil(n/48) of the sequence
    uint32_t mod3;        // n % 3
    bool tiny;            // n < 3 (frames modes validate nothing)
};

// First-error search over one run's window (only on tiles the decode pass flagged).
__device__ __forceinline__ void validate_run(const uint32_t (&W)[7], int vcnt, int strict, uint64_t key0,
                                             unsigned long long* err_key) {
#pragma unroll
    for (int k = 0; k < 6; k++) {
        int left = vcnt - 8 * k;
        if (left <= 0) break;
        uint32_t x = W[k];
        uint32_t f = strict ? ((x | (x >> 1)) & 0x44444444u) >> 2 : (x & (x >> 1) & (x >> 2) & (x >> 3) & 0x11111111u);
        if (left < 8) f &= (1u << (4 * left)) - 1u;
        if (f) {
            uint32_t pos = 8 * k + ((__ffs(f) - 1) >> 2);
            atomicMin(err_key, (unsigned long long)(key0 + pos));
            break;
        }
    }
}

// Shared memory of one CTA.  raw[] is the TMA landing zone: while tile T is decoded and translated,
// the bytes of the CTA's next tile stream into the other buffer (cp.async.bulk + mbarrier), so the
// DRAM latency of a tile is hidden behind the previous tile's arithmetic.
struct FramesSmem {
    alignas(256) uint8_t decode[256];        // input byte -> code; 256-aligned so base | byte is its address
    alignas(16) uint16_t lut[LUT_ENTRIES];   // codon -> (forward aa, reverse aa)
    alignas(16) uint32_t nib[NIB_WORDS];     // the tile's nucleotides as 4-bit codes, 8 per word
    alignas(128) uint8_t raw[2][RAW_BYTES];  // double-buffered raw input of a tile, 2-nt halo included
    alignas(8) uint64_t full[2];             // mbarriers: raw[b] has landed
    uint64_t span[2][2];                     // flat [start, end) of the input span staged in raw[b]
    uint32_t search[SEARCH_CAP + 1];
    uint32_t suspicious[2];                  // per buffer: the decode pass saw a code that needs the exact check
};

// flat offset of the first run of tile t (t * TILE_RUNS < total_runs)
__device__ __forceinline__ uint64_t tile_flat_start(const FramesParams& p, uint64_t t) {
    if (p.tiles) return p.tiles[t].flat;
    const uint32_t g = (uint32_t)(t * TILE_RUNS);
    const uint32_t r = p.uniform_runs == 1 ? g : (uint32_t)__umul64hi((uint64_t)g, p.uniform_magic);
    return (uint64_t)r * p.uniform_len + (uint64_t)(g - r * p.uniform_runs) * RUN_NT;
}

// One thread: publish the span of tile t in S.span[b] and start its bulk copy into S.raw[b].
// Only whole 16-byte granules inside [dna, dna + total) are copied by TMA; the (at most two) granules
// that straddle the ends of the whole buffer are read byte-wise by the decode pass instead.
__device__ __forceinline__ void stage_tile(FramesSmem& S, const FramesParams& p, uint64_t t, uint64_t total_runs, int b) {
    const uint64_t lo = tile_flat_start(p, t);
    uint64_t hi = p.total_nt;
    if ((t + 1) * TILE_RUNS < total_runs) hi = min(p.total_nt, tile_flat_start(p, t + 1) + 2);  // 2-nt halo
    S.span[b][0] = lo;
    S.span[b][1] = hi;
    const uintptr_t dna_lo = reinterpret_cast<uintptr_t>(p.dna), dna_hi = dna_lo + p.total_nt;
    const uintptr_t a0 = (dna_lo + lo) & ~(uintptr_t)15;
    const uintptr_t a1 = (dna_lo + hi + 15) & ~(uintptr_t)15;
    const uintptr_t t0 = max(a0, (dna_lo + 15) & ~(uintptr_t)15);
    const uintptr_t t1 = min(a1, dna_hi & ~(uintptr_t)15);
    if (t1 > t0) {
        const uint32_t bytes = (uint32_t)(t1 - t0);
        mbar_expect_tx(&S.full[b], bytes);
        tma_bulk_g2s(S.raw[b] + (t0 - a0), reinterpret_cast<const void*>(t0), bytes, &S.full[b]);
    } else {
        mbar_arrive(&S.full[b]);
    }
}

template <int FRAMES>
__global__ void __launch_bounds__(TILE_THREADS, 4) translate_frames_kernel(const FramesParams p) {
    __shared__ FramesSmem S;
    const int tid = threadIdx.x;
    const bool uniform = (p.offsets == nullptr);
    const uint64_t total_runs = uniform ? p.total_runs : p.run_prefix[p.n_seq];
    const uint64_t n_tiles = min((uint64_t)p.n_tiles, (total_runs + TILE_RUNS - 1) / TILE_RUNS);

    // Per-CTA tables (persistent CTA: loaded once, reused for every tile) and the first bulk copy.
    if (tid == 0) {
        mbar_init(&S.full[0], 1);
        mbar_init(&S.full[1], 1);
        fence_barrier_init();
        S.suspicious[0] = 0;
        if (blockIdx.x < n_tiles) stage_tile(S, p, blockIdx.x, total_runs, 0);
    }
    {
        const uint4* src = reinterpret_cast<const uint4*>(p.lut);
        uint4* dst = reinterpret_cast<uint4*>(S.lut);
        for (int i = tid; i < (int)(LUT_ENTRIES * 2 / 16); i += TILE_THREADS) dst[i] = ldg_cached(src + i);
        if (tid < 16) reinterpret_cast<uint4*>(S.decode)[tid] = ldg_cached(reinterpret_cast<const uint4*>(p.decode) + tid);
    }
    const uintptr_t dna_lo = reinterpret_cast<uintptr_t>(p.dna);
    const uintptr_t dna_hi = dna_lo + p.total_nt;
    // shared-window addresses of the two tables: folded into the address-forming instruction itself
    // (dp4a accumulator / byte permute), so a lookup is [extract, load] with no separate add
    const uint32_t lut_base = smem_u32(S.lut);
    const uint32_t dec_base = smem_u32(S.decode);
    __syncthreads();

    uint32_t it = 0;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint64_t g0 = tile * TILE_RUNS;
        const uint32_t cnt = (uint32_t)min((uint64_t)TILE_RUNS, total_runs - g0);
        if (tid == 0) {
            S.suspicious[buf ^ 1] = 0;  // last read in the previous iteration, next set in the following one
            // raw[buf ^ 1] was last read during the previous iteration's decode pass (barrier since)
            if (tile + gridDim.x < n_tiles) stage_tile(S, p, tile + gridDim.x, total_runs, buf ^ 1);
        }

        // ---- phase 0: which run is mine --------------------------------------------------------
        RunDesc rd;
        const bool active = tid < (int)cnt;
        if (uniform) {
            const uint32_t g = (uint32_t)g0 + (active ? tid : 0);
            const uint32_t r = p.uniform_runs == 1 ? g : (uint32_t)__umul64hi((uint64_t)g, p.uniform_magic);
            rd.m = g - r * p.uniform_runs;
            rd.runs = p.uniform_runs;
            rd.mod3 = p.uniform_mod3;
            rd.tiny = p.uniform_len < 3;
            const uint64_t in_seq = (uint64_t)rd.m * RUN_NT;
            rd.left_nt = p.uniform_len - in_seq;
            rd.flat_start = (uint64_t)r * p.uniform_len + in_seq;
            rd.out_base = (uint64_t)r * ((uint64_t)p.uniform_runs * (16u * FRAMES));
        } else {
            // cache the run prefix of the sequences this tile touches, then upper_bound in smem
            const uint32_t r_first = p.tiles[tile].first_seq;
            for (uint32_t i = tid; i <= SEARCH_CAP; i += TILE_THREADS) {
                uint64_t r = (uint64_t)r_first + i;
                uint64_t v = (r <= p.n_seq) ? p.run_prefix[r] : ~0ull;
                uint64_t rel = v >= g0 ? v - g0 : 0;
                S.search[i] = rel > 0xffffffffull ? 0xffffffffu : (uint32_t)rel;
            }
            __syncthreads();
            const uint32_t gl = active ? tid : 0;  // run index relative to g0
            uint32_t lo = 0, hi = SEARCH_CAP;      // largest i with search[i] <= gl
            while (lo < hi) {
                uint32_t mid = (lo + hi + 1) >> 1;
                if (S.search[mid] <= gl) lo = mid; else hi = mid - 1;
            }
            uint64_t r = (uint64_t)r_first + lo;
            if (lo == SEARCH_CAP) {  // more than SEARCH_CAP sequences start inside this tile: finish in global
                uint64_t a = r, b = p.n_seq;  // invariant: run_prefix[a] <= g
                const uint64_t g = g0 + gl;
                while (a < b) {
                    uint64_t mid = (a + b + 1) >> 1;
                    if (p.run_prefix[mid] <= g) a = mid; else b = mid - 1;
                }
                r = a;
            }
            const uint64_t rp = p.run_prefix[r];
            const uint64_t o0 = p.offsets[r] - p.flat_base;
            const uint64_t n = p.offsets[r + 1] - p.flat_base - o0;
            rd.m = (uint32_t)(g0 + gl - rp);
            rd.runs = (uint32_t)((n + RUN_NT - 1) / RUN_NT);
            rd.mod3 = (n >> 32) ? (uint32_t)(n % 3) : ((uint32_t)n % 3u);
            rd.tiny = n < 3;
            const uint64_t in_seq = (uint64_t)rd.m * RUN_NT;
            rd.left_nt = n - in_seq;
            rd.flat_start = o0 + in_seq;
            rd.out_base = rp * (16u * FRAMES);
        }
        const int have = (int)min((uint64_t)(RUN_NT + 2), rd.left_nt);  // nucleotides this run may look at

        // ---- phase 1: the staged bytes -> 4-bit codes in shared memory --------------------------
        const uintptr_t span_lo = dna_lo + S.span[buf][0];
        const uintptr_t span_hi = dna_lo + S.span[buf][1];
        const uintptr_t a0 = span_lo & ~(uintptr_t)15;
        const uint32_t n_chunks = (uint32_t)((span_hi - a0 + 15) >> 4);
        mbar_wait(&S.full[buf], (it >> 1) & 1);
        uint32_t flagged = 0;
        for (uint32_t c = tid; c < n_chunks; c += TILE_THREADS) {
            const uintptr_t addr = a0 + ((uintptr_t)c << 4);
            uint32_t w[4];
            if (addr >= dna_lo && addr + 16 <= dna_hi) {
                const uint4 v = *reinterpret_cast<const uint4*>(S.raw[buf] + ((size_t)c << 4));
                w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
            } else {  // first / last granule of the whole buffer: byte-wise, nothing outside [dna, dna+total)
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    uint32_t x = 0;
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        uintptr_t a = addr + q * 4 + b;
                        uint32_t byte = (a >= dna_lo && a < dna_hi) ? *reinterpret_cast<const uint8_t*>(a) : 0u;
                        x |= byte << (8 * b);
                    }
                    w[q] = x;
                }
            }
            uint32_t nw[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t wa = w[2 * h], wb = w[2 * h + 1];
                // one permute per nucleotide extracts the byte INTO the table's base address; the *16
                // chain that packs the codes runs on the FMA pipe
                uint32_t acc = lds_u8(__byte_perm(wb, dec_base, 0x7653));
                acc = acc * 16u + lds_u8(__byte_perm(wb, dec_base, 0x7652));
                acc = acc * 16u + lds_u8(__byte_perm(wb, dec_base, 0x7651));
                acc = acc * 16u + lds_u8(__byte_perm(wb, dec_base, 0x7650));
                acc = acc * 16u + lds_u8(__byte_perm(wa, dec_base, 0x7653));
                acc = acc * 16u + lds_u8(__byte_perm(wa, dec_base, 0x7652));
                acc = acc * 16u + lds_u8(__byte_perm(wa, dec_base, 0x7651));
                acc = acc * 16u + lds_u8(__byte_perm(wa, dec_base, 0x7650));
                nw[h] = acc;
            }
            // conservative filter: strict -> any code >= 4; else any code >= 12 (exact check in phase 2)
            const uint32_t both = nw[0] | nw[1];
            flagged |= p.strict ? (both & 0xCCCCCCCCu) : ((nw[0] & (nw[0] >> 1) & 0x44444444u) | (nw[1] & (nw[1] >> 1) & 0x44444444u));
            *reinterpret_cast<uint2*>(&S.nib[2 * c]) = make_uint2(nw[0], nw[1]);
        }
        if (flagged) S.suspicious[buf] = 1;
        __syncthreads();

        // ---- phase 2: one run per thread: 48 codon starts -> 16 residues per frame per strand -----
        if (active) {
            const uint32_t o = (uint32_t)(dna_lo + rd.flat_start - a0);  // local nucleotide offset of the run
            const uint32_t wi = o >> 3, sh = (o & 7u) * 4u;
            uint32_t W[7];
            {
                uint32_t raw[8];
#pragma unroll
                for (int k = 0; k < 8; k++) raw[k] = S.nib[wi + k];
#pragma unroll
                for (int k = 0; k < 7; k++) W[k] = __funnelshift_r(raw[k], raw[k + 1], sh);
            }
            if (S.suspicious[buf]) {
                // validated range: translate -> first 3*floor(n/3) bytes (src/trans_table.rs:191);
                // frames -> every byte when n >= 3 (union of the frames), none otherwise
                uint64_t vleft = rd.left_nt;  // validated nucleotides from the run start on
                if (FRAMES == 1) vleft = vleft >= rd.mod3 ? vleft - rd.mod3 : 0;
                else if (rd.tiny) vleft = 0;
                if (vleft) validate_run(W, (int)min((uint64_t)RUN_NT, vleft), p.strict, p.key_base + rd.flat_start, p.err_key);
            }
            if (have < RUN_NT + 2) {  // past the end of the sequence: code 15 -> LUT entry 0
#pragma unroll
                for (int k = 0; k < 7; k++) {
                    const int left = have - 8 * k;
                    if (left <= 0) W[k] = 0xffffffffu;
                    else if (left < 8) W[k] |= 0xffffffffu << (4 * left);
                }
            }
            // codes -> doubled codes, one per byte: Bq[q] holds nucleotides 4q..4q+3
            uint32_t Bq[13];
#pragma unroll
            for (int k = 0; k < 7; k++) {
                const uint32_t lo2 = (W[k] << 1) & 0x1e1e1e1eu;  // nibbles 0,2,4,6 doubled
                const uint32_t hi2 = (W[k] >> 3) & 0x1e1e1e1eu;  // nibbles 1,3,5,7 doubled
                Bq[2 * k] = __byte_perm(lo2, hi2, 0x5140);
                if (2 * k + 1 < 13) Bq[2 * k + 1] = __byte_perm(lo2, hi2, 0x7362);
            }
            uint8_t* const out_seq = p.out + rd.out_base;
            const uint32_t slot = rd.runs * 16u;
            uint8_t* const fwd_dst = out_seq + (size_t)rd.m * 16u;
            uint8_t* const rc_dst = out_seq + 4 * (size_t)slot - (size_t)(rd.m + 1) * 16u;
#pragma unroll
            for (int i = 0; i < (FRAMES == 1 ? 1 : 3); i++) {
                uint32_t F[4], R[4];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    uint32_t r4[4];
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        const int j = 3 * (4 * q + e) + i;  // codon start within the run
                        const int wq = j >> 2, rb = j & 3;
                        const uint32_t x = rb == 0 ? Bq[wq] : __funnelshift_r(Bq[wq], Bq[wq + 1], 8 * rb);
                        r4[e] = lds_u16(__dp4a(x, LUT_DP4A_COEF, lut_base));  // base + 2 * lut_index(c0, c1, c2)
                    }
                    // r = fwd | rc << 8.  Pack pairs with a multiply-add (FMA pipe), finish with two permutes.
                    const uint32_t t01 = r4[1] * 65536u + r4[0];  // [f0, r0, f1, r1]
                    const uint32_t t23 = r4[3] * 65536u + r4[2];  // [f2, r2, f3, r3]
                    F[q] = __byte_perm(t01, t23, 0x6420);         // f0 f1 f2 f3
                    R[q] = __byte_perm(t01, t23, 0x1357);         // r3 r2 r1 r0
                }
                stg_stream(fwd_dst + (size_t)i * slot, make_uint4(F[0], F[1], F[2], F[3]));
                if (FRAMES == 6) {
                    // codon start p = 48m + 3jj + i is reverse frame (n - i) mod 3, residue (n-3-i)/3 - 16m - jj
                    uint32_t kq = rd.mod3 + 3u - i;
                    kq = kq >= 3u ? kq - 3u : kq;
                    stg_stream(rc_dst + (size_t)kq * slot, make_uint4(R[3], R[2], R[1], R[0]));
                }
            }
        }
        __syncthreads();  // nib / span / search are rewritten by the next tile
    }
}

// ------------------------------------------------------------------------------------------------
// Reverse complement of one sequence: out[n-1-i] = comp[in[i]]
// ------------------------------------------------------------------------------------------------
struct RevcompParams {
    const uint8_t* dna;
    uint64_t n;
    uint8_t* out;            // 16-byte aligned
    const uint8_t* table;    // 256-byte complement table, 0 = invalid byte
    unsigned long long* err_key;
};

constexpr int RC_THREADS = 256;
constexpr int RC_UNROLL = 4;
constexpr int RC_TILE_CHUNKS = RC_THREADS * RC_UNROLL;  // 16-byte output chunks per tile (16 KiB)
constexpr int RC_RAW_BYTES = (RC_TILE_CHUNKS + 1) * 16;  // + one granule for the source misalignment

__device__ __forceinline__ uint32_t rc_map_word(uint32_t w, uint32_t tab) {
    // input bytes b0..b3 (ascending address) -> output word whose ascending bytes are c(b3)..c(b0)
    // `tab` is 256-byte aligned: one permute extracts a byte straight into the table address
    uint32_t r = lds_u8(__byte_perm(w, tab, 0x7650));
    r = r * 256u + lds_u8(__byte_perm(w, tab, 0x7651));
    r = r * 256u + lds_u8(__byte_perm(w, tab, 0x7652));
    r = r * 256u + lds_u8(__byte_perm(w, tab, 0x7653));
    return r;
}
__device__ __forceinline__ uint32_t has_zero_byte(uint32_t v) { return (v - 0x01010101u) & ~v & 0x80808080u; }

struct RevcompSmem {
    alignas(256) uint8_t tab[256];
    alignas(128) uint8_t raw[2][RC_RAW_BYTES];  // double-buffered source bytes of a tile (TMA landing zone)
    alignas(8) uint64_t full[2];
};

// Tile t produces output chunks [t*RC_TILE_CHUNKS, ...): out[16j, 16j+16) = in[n-16j-16, n-16j) reversed and
// complemented.  Its source is one contiguous span that ends at in_end - 16*j0; `d` (the misalignment of
// in_end) is the same for every chunk, so the span is staged as whole 16-byte granules starting at a0.
__device__ __forceinline__ void rc_stage_tile(RevcompSmem& S, const RevcompParams& p, uint64_t t, uint64_t full_chunks, int b) {
    const uint64_t j0 = t * RC_TILE_CHUNKS;
    const uint32_t cnt = (uint32_t)min((uint64_t)RC_TILE_CHUNKS, full_chunks - j0);
    const uintptr_t dna_lo = reinterpret_cast<uintptr_t>(p.dna), dna_hi = dna_lo + p.n;
    const uint32_t d = (uint32_t)(dna_hi & 15u);
    const uintptr_t a0 = dna_hi - 16 * (j0 + cnt) - d;  // granule holding the tile's lowest source byte
    const uintptr_t a1 = a0 + 16 * (uintptr_t)(cnt + (d ? 1 : 0));
    const uintptr_t t0 = max(a0, (dna_lo + 15) & ~(uintptr_t)15);
    const uintptr_t t1 = min(a1, dna_hi & ~(uintptr_t)15);
    if (t1 > t0) {
        const uint32_t bytes = (uint32_t)(t1 - t0);
        mbar_expect_tx(&S.full[b], bytes);
        tma_bulk_g2s(S.raw[b] + (t0 - a0), reinterpret_cast<const void*>(t0), bytes, &S.full[b]);
    } else {
        mbar_arrive(&S.full[b]);
    }
}

// granule g of the staged tile; granules that straddle the ends of the whole buffer were not staged
__device__ __forceinline__ uint4 rc_granule(const RevcompSmem& S, int b, uintptr_t a0, uint32_t g, uintptr_t dna_lo, uintptr_t dna_hi) {
    const uintptr_t addr = a0 + 16 * (uintptr_t)g;
    if (addr >= dna_lo && addr + 16 <= dna_hi) return *reinterpret_cast<const uint4*>(S.raw[b] + 16 * (size_t)g);
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        uint32_t x = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uintptr_t a = addr + 4 * q + k;
            const uint32_t byte = (a >= dna_lo && a < dna_hi) ? *reinterpret_cast<const uint8_t*>(a) : 0x41u;  // filler never used
            x |= byte << (8 * k);
        }
        w[q] = x;
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void __launch_bounds__(RC_THREADS, 4) revcomp_kernel(const RevcompParams p) {
    __shared__ RevcompSmem S;
    const int tid = threadIdx.x;
    const uint64_t n = p.n;
    const uint64_t full = n >> 4;  // full 16-byte output chunks
    const uint64_t n_tiles = (full + RC_TILE_CHUNKS - 1) / RC_TILE_CHUNKS;
    if (tid == 0) {
        mbar_init(&S.full[0], 1);
        mbar_init(&S.full[1], 1);
        fence_barrier_init();
        if (blockIdx.x < n_tiles) rc_stage_tile(S, p, blockIdx.x, full, 0);
    }
    if (tid < 16) reinterpret_cast<uint4*>(S.tab)[tid] = ldg_cached(reinterpret_cast<const uint4*>(p.table) + tid);
    __syncthreads();
    const uint32_t tab = smem_u32(S.tab);
    const uintptr_t dna_lo = reinterpret_cast<uintptr_t>(p.dna), dna_hi = dna_lo + n;
    const uint32_t d = (uint32_t)(dna_hi & 15u);  // misalignment of every chunk's source
    const uint32_t ws = d >> 2, bs = (d & 3u) * 8u;

    uint32_t it = 0;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        if (tid == 0 && tile + gridDim.x < n_tiles) rc_stage_tile(S, p, tile + gridDim.x, full, buf ^ 1);
        const uint64_t j0 = tile * RC_TILE_CHUNKS;
        const uint32_t cnt = (uint32_t)min((uint64_t)RC_TILE_CHUNKS, full - j0);
        const uintptr_t a0 = dna_hi - 16 * (j0 + cnt) - d;
        mbar_wait(&S.full[buf], (it >> 1) & 1);
#pragma unroll
        for (int u = 0; u < RC_UNROLL; u++) {
            const uint32_t l = (uint32_t)u * RC_THREADS + tid;  // chunk within the tile
            if (l < cnt) {
                const uint64_t j = j0 + l;
                const uint32_t g = cnt - 1 - l;  // granule holding in[n-16j-16] (at byte offset d)
                uint32_t x[4];
                if (d == 0) {
                    const uint4 v = rc_granule(S, buf, a0, g, dna_lo, dna_hi);
                    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
                } else {
                    const uint4 lo = rc_granule(S, buf, a0, g, dna_lo, dna_hi);
                    const uint4 hi = rc_granule(S, buf, a0, g + 1, dna_lo, dna_hi);
                    const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
                    switch (ws) {  // uniform across the grid
                        case 0:
#pragma unroll
                            for (int k = 0; k < 4; k++) x[k] = __funnelshift_r(w[k], w[k + 1], bs);
                            break;
                        case 1:
#pragma unroll
                            for (int k = 0; k < 4; k++) x[k] = __funnelshift_r(w[k + 1], w[k + 2], bs);
                            break;
                        case 2:
#pragma unroll
                            for (int k = 0; k < 4; k++) x[k] = __funnelshift_r(w[k + 2], w[k + 3], bs);
                            break;
                        default:
#pragma unroll
                            for (int k = 0; k < 4; k++) x[k] = __funnelshift_r(w[k + 3], w[(k + 4) & 7], bs);
                            break;
                    }
                }
                // x[0..3] = in[n-16j-16 .. n-16j) ; output ascending = input descending
                uint4 o;
                o.x = rc_map_word(x[3], tab);
                o.y = rc_map_word(x[2], tab);
                o.z = rc_map_word(x[1], tab);
                o.w = rc_map_word(x[0], tab);
                if (has_zero_byte(o.x) | has_zero_byte(o.y) | has_zero_byte(o.z) | has_zero_byte(o.w)) {
                    // lowest failing INPUT index of this chunk = highest failing output byte
                    const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
                    for (int t = 15; t >= 0; t--)
                        if (((ow[t >> 2] >> (8 * (t & 3))) & 0xffu) == 0) {
                            atomicMin(p.err_key, (unsigned long long)(n - 1 - (16 * j + t)));
                            break;
                        }
                }
                stg_stream(p.out + 16 * j, o);
            }
        }
        __syncthreads();  // raw[buf] may be restaged two iterations from now by thread 0
    }
    // tail: out[16*full, n) = in[0, n mod 16) reversed
    if (blockIdx.x == 0 && tid < (int)(n & 15u)) {
        const uint64_t i = tid;
        const uint32_t v = S.tab[p.dna[i]];
        if (v == 0) atomicMin(p.err_key, (unsigned long long)i);
        p.out[n - 1 - i] = (uint8_t)v;
    }
}

// ------------------------------------------------------------------------------------------------
// Windows: out[i*w + j] = seq[i + j]   (DnaSequence::windows / ProteinSequence::windows)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) windows_kernel(const uint8_t* __restrict__ seq, uint64_t n_windows, uint32_t w,
                                                      uint8_t* __restrict__ out) {
    // one thread per aligned 16-byte chunk of the output (out is 16-byte aligned)
    const uint64_t total = n_windows * w;
    const uint64_t n_chunks = (total + 15) >> 4;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t b0 = c << 4;
        uint64_t i = b0 / w;
        uint32_t j = (uint32_t)(b0 - i * w);
        uint32_t word[4] = {0, 0, 0, 0};
        const int lim = (int)min((uint64_t)16, total - b0);
        for (int t = 0; t < lim; t++) {
            word[t >> 2] |= (uint32_t)__ldg(seq + i + j) << (8 * (t & 3));
            if (++j == w) { j = 0; ++i; }
        }
        if (lim == 16) stg_stream(out + b0, make_uint4(word[0], word[1], word[2], word[3]));
        else for (int t = 0; t < lim; t++) out[b0 + t] = (uint8_t)(word[t >> 2] >> (8 * (t & 3)));
    }
}

// origin indices of benches/all_windows.rs:39-75 for one segment of windows
//   kind 0: i ; kind 1: count-1-i ; kind 2: 3*i + f ; kind 3: L - ((i + w_aa)*3 + f)
__global__ void origin_kernel(uint64_t* __restrict__ out, uint64_t count, int kind, uint64_t f, uint64_t L, uint64_t w_aa) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t v;
        switch (kind) {
            case 0: v = i; break;
            case 1: v = count - 1 - i; break;
            case 2: v = 3 * i + f; break;
            default: v = L - ((i + w_aa) * 3 + f); break;
        }
        out[i] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// Canonical windows (strict alphabet): per window, min_lex(fw(window), fw(reverse(window)))
// ------------------------------------------------------------------------------------------------
struct CanonParams {
    const uint8_t* dna;
    uint64_t n;
    uint32_t w;
    int forward_only;
    uint8_t* out;            // n_windows * w bytes, 16-byte aligned
    const uint8_t* decode;   // 256-byte table -> internal code (0..3 concrete)
    uint32_t letters;        // output byte of codes 0..3 packed little-endian (ASCII "ATCG" or masks 1,2,4,8)
    unsigned long long* err_key;
};

constexpr int CANON_THREADS = 128;  // windows per CTA
constexpr int CANON_MAX_W = 64;

// Relabel by first occurrence (src/canonical.rs:95-118) on 2-bit codes held in two 64-bit words
// (nucleotide j at bits 2j of lo for j < 32, of hi otherwise).
__device__ __forceinline__ void fw_canon_packed(uint64_t lo, uint64_t hi, uint32_t w, uint64_t& olo, uint64_t& ohi) {
    uint32_t perm = 0, seen = 0, next = 0;
    olo = 0; ohi = 0;
    for (uint32_t j = 0; j < w; j++) {
        const uint32_t x = (uint32_t)(((j < 32) ? (lo >> (2 * j)) : (hi >> (2 * (j - 32)))) & 3u);
        if (!((seen >> x) & 1u)) {
            seen |= 1u << x;
            perm |= next << (2 * x);
            next++;
        }
        const uint64_t y = (perm >> (2 * x)) & 3u;
        if (j < 32) olo |= y << (2 * j); else ohi |= y << (2 * (j - 32));
    }
}

__global__ void __launch_bounds__(CANON_THREADS) canonical_windows_kernel(const CanonParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: codes[CANON_THREADS + CANON_MAX_W] | stage[16 + CANON_THREADS * w]
    uint8_t* codes = smem_raw;
    uint8_t* stage = smem_raw + ((CANON_THREADS + CANON_MAX_W + 15) & ~15);
    const int tid = threadIdx.x;
    const uint32_t w = p.w;
    const uint64_t n_windows = p.n - w + 1;
    for (uint64_t blk = blockIdx.x; blk * CANON_THREADS < n_windows; blk += gridDim.x) {
        const uint64_t i0 = blk * CANON_THREADS;
        const uint32_t nwin = (uint32_t)min((uint64_t)CANON_THREADS, n_windows - i0);
        const uint32_t nnt = nwin + w - 1;
        __syncthreads();
        for (uint32_t t = tid; t < nnt; t += CANON_THREADS) {
            const uint32_t c = __ldg(p.decode + __ldg(p.dna + i0 + t));
            if (c > 3) atomicMin(p.err_key, (unsigned long long)(i0 + t));
            codes[t] = (uint8_t)(c & 3u);
        }
        __syncthreads();
        const uint64_t out0 = i0 * w;                 // byte offset of this block's first window
        const uint32_t phase = (uint32_t)(out0 & 15u);  // stage mirrors the output's 16-byte phase
        if (tid < (int)nwin) {
            uint64_t lo = 0, hi = 0, rlo = 0, rhi = 0;
            for (uint32_t j = 0; j < w; j++) {
                const uint64_t x = codes[tid + j];
                if (j < 32) lo |= x << (2 * j); else hi |= x << (2 * (j - 32));
                const uint32_t jr = w - 1 - j;
                if (jr < 32) rlo |= x << (2 * jr); else rhi |= x << (2 * (jr - 32));
            }
            uint64_t flo, fhi;
            fw_canon_packed(lo, hi, w, flo, fhi);
            if (!p.forward_only) {
                uint64_t glo, ghi;
                fw_canon_packed(rlo, rhi, w, glo, ghi);
                // lexical order, first nucleotide most significant, under A<T<C<G == code order
                bool rev_less = false;
                for (uint32_t j = 0; j < w; j++) {
                    const uint32_t a = (uint32_t)(((j < 32) ? (flo >> (2 * j)) : (fhi >> (2 * (j - 32)))) & 3u);
                    const uint32_t b = (uint32_t)(((j < 32) ? (glo >> (2 * j)) : (ghi >> (2 * (j - 32)))) & 3u);
                    if (a != b) { rev_less = b < a; break; }
                }
                if (rev_less) { flo = glo; fhi = ghi; }
            }
            uint8_t* dst = stage + phase + (size_t)tid * w;
            for (uint32_t j = 0; j < w; j++) {
                const uint32_t y = (uint32_t)(((j < 32) ? (flo >> (2 * j)) : (fhi >> (2 * (j - 32)))) & 3u);
                dst[j] = (uint8_t)(p.letters >> (8 * y));
            }
        }
        __syncthreads();
        // coalesced copy-out: head bytes to the first 16-byte boundary, aligned uint4 body, tail bytes
        const uint32_t nbytes = nwin * w;
        uint8_t* gdst = p.out + out0;
        const uint32_t head = min(nbytes, (16u - phase) & 15u);
        const uint32_t body = (nbytes - head) >> 4;
        if (tid < (int)head) gdst[tid] = stage[phase + tid];
        for (uint32_t c = tid; c < body; c += CANON_THREADS)
            stg_stream(gdst + head + 16 * c, *reinterpret_cast<const uint4*>(stage + phase + head + 16 * c));
        const uint32_t done = head + 16 * body;
        if (tid < (int)(nbytes - done)) gdst[done + tid] = stage[phase + done + tid];
    }
}

// ------------------------------------------------------------------------------------------------
// Expansions of ambiguous windows (src/expansions.rs:35-108)
// ------------------------------------------------------------------------------------------------
struct ExpandParams {
    const uint8_t* windows;  // n_win * w bytes
    uint64_t n_win;
    uint32_t w;
    uint64_t max_expansions;
    const uint8_t* to_mask;  // 256-byte table: input byte -> 4-bit IUPAC mask (0 = invalid)
    uint32_t letters;        // output byte for bases A,T,C,G (ascending mask bit), packed little-endian
    uint64_t* counts;        // size_hint per window, UINT64_MAX on overflow
    uint64_t* emit;          // expansions this window writes (0 if skipped)
    const uint64_t* out_index;  // exclusive scan of emit (phase 2)
    uint8_t* out;
    unsigned long long* err_key;
};

// size_hint for a fresh iterator: prod(popcount) with checked arithmetic (src/expansions.rs:96-107)
__global__ void __launch_bounds__(256) expansion_count_kernel(const ExpandParams p) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n_win; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint8_t* win = p.windows + i * p.w;
        uint64_t size = 0;
        bool overflow = false;
        for (uint32_t j = 0; j < p.w; j++) {
            const uint32_t m = __ldg(p.to_mask + __ldg(win + j));
            if (m == 0) atomicMin(p.err_key, (unsigned long long)(i * p.w + j));
            const uint64_t states = (uint64_t)__popc(m);
            if (states > 1 && !overflow) {
                if (size != 0 && states > 0xffffffffffffffffull / size) overflow = true;
                else {
                    size *= states;
                    if (size > 0xffffffffffffffffull - (states - 1)) overflow = true;
                    else size += states - 1;
                }
            }
        }
        if (!overflow && size == 0xffffffffffffffffull) overflow = true;
        const uint64_t count = overflow ? 0xffffffffffffffffull : size + 1;
        p.counts[i] = count;
        p.emit[i] = (count <= p.max_expansions) ? count : 0;
    }
}

// one thread per (window, expansion): digit of ambiguity a = (k / prod(states of later ambiguities)) % states_a,
// i.e. the odometer of src/expansions.rs:71-94 with the LAST ambiguity fastest
__global__ void __launch_bounds__(256) expansion_write_kernel(const ExpandParams p, uint64_t total_rows) {
    for (uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; row < total_rows; row += (uint64_t)gridDim.x * blockDim.x) {
        // window owning this row: largest i with out_index[i] <= row
        uint64_t lo = 0, hi = p.n_win - 1;
        while (lo < hi) {
            uint64_t mid = (lo + hi + 1) >> 1;
            if (p.out_index[mid] <= row) lo = mid; else hi = mid - 1;
        }
        const uint64_t i = lo;
        uint64_t k = row - p.out_index[i];
        const uint8_t* win = p.windows + i * p.w;
        uint8_t* dst = p.out + row * p.w;
        for (int j = (int)p.w - 1; j >= 0; j--) {
            const uint32_t m = __ldg(p.to_mask + __ldg(win + j));
            const uint32_t states = __popc(m);
            uint32_t digit = 0;
            if (states > 1) { digit = (uint32_t)(k % states); k /= states; }
            // digit-th set bit of m, ascending (possibilities() order A,T,C,G)
            uint32_t mm = m;
            for (uint32_t s = 0; s < digit; s++) mm &= mm - 1;
            const uint32_t base = __ffs(mm) - 1;  // 0..3 = A,T,C,G
            dst[j] = (uint8_t)(p.letters >> (8 * base));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Batch geometry pre-pass (variable-length batches): runs per sequence -> exclusive scan, and the
// sequence that holds the first run of every tile.
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;  // items per thread
constexpr int SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint64_t block_exclusive_scan(uint64_t v, uint64_t* total, uint64_t* warp_sums) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint64_t s = (lane < (int)(blockDim.x >> 5)) ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint64_t y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += y;
        }
        warp_sums[lane] = s;  // inclusive
    }
    __syncthreads();
    const uint64_t warp_off = wid ? warp_sums[wid - 1] : 0;
    *total = warp_sums[(blockDim.x >> 5) - 1];
    return warp_off + x - v;
}

// value(i): 0 = runs of sequence i from CSR offsets (ceil(len/48)); 1 = raw uint64 values
template <int MODE>
__device__ __forceinline__ uint64_t scan_value(const uint64_t* src, uint64_t i) {
    if (MODE == 0) return (src[i + 1] - src[i] + RUN_NT - 1) / RUN_NT;
    return src[i];
}

template <int MODE>
__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums_kernel(const uint64_t* src, uint64_t n, uint64_t* block_sums) {
    __shared__ uint64_t warp_sums[32];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_BLOCK + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
        if (base + k < n) s += scan_value<MODE>(src, base + k);
    uint64_t total;
    block_exclusive_scan(s, &total, warp_sums);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single CTA: exclusive scan of the block sums in place
__global__ void __launch_bounds__(SCAN_THREADS) scan_spine_kernel(uint64_t* block_sums, uint64_t n_blocks) {
    __shared__ uint64_t warp_sums[32];
    __shared__ uint64_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint64_t base = 0; base < n_blocks; base += SCAN_THREADS) {
        const uint64_t i = base + threadIdx.x;
        const uint64_t v = i < n_blocks ? block_sums[i] : 0;
        uint64_t total;
        const uint64_t ex = block_exclusive_scan(v, &total, warp_sums);
        const uint64_t carry = carry_s;
        if (i < n_blocks) block_sums[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
}

// dst[0..n] = exclusive scan (dst[n] = grand total)
template <int MODE>
__global__ void __launch_bounds__(SCAN_THREADS) scan_finish_kernel(const uint64_t* src, uint64_t n, const uint64_t* block_sums,
                                                                   uint64_t* dst) {
    __shared__ uint64_t warp_sums[32];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_BLOCK + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint64_t v[SCAN_ITEMS];
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = (base + k < n) ? scan_value<MODE>(src, base + k) : 0;
        s += v[k];
    }
    uint64_t total;
    uint64_t ex = block_exclusive_scan(s, &total, warp_sums) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n) dst[base + k] = ex;
        ex += v[k];
        if (base + k + 1 == n) dst[n] = ex;
    }
    if (n == 0 && blockIdx.x == 0 && threadIdx.x == 0) dst[0] = 0;
}

// tiles[t] = {flat offset of run t*TILE_RUNS, sequence containing it}
__global__ void __launch_bounds__(256) tile_info_kernel(const uint64_t* __restrict__ run_prefix, const uint64_t* __restrict__ offsets,
                                                        uint64_t flat_base, uint64_t n_seq, TileInfo* __restrict__ tiles,
                                                        uint64_t n_tiles_cap) {
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_seq; r += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t a = run_prefix[r], b = run_prefix[r + 1];
        if (b == a) continue;
        const uint64_t o0 = offsets[r] - flat_base;
        uint64_t t = (a + TILE_RUNS - 1) / TILE_RUNS;
        for (; t * TILE_RUNS < b && t < n_tiles_cap; t++) {
            TileInfo ti;
            ti.flat = o0 + (t * TILE_RUNS - a) * RUN_NT;
            ti.first_seq = (uint32_t)r;
            ti.pad_ = 0;
            tiles[t] = ti;
        }
    }
}

}  // namespace qd

// quickdna_b200 -- sm_100a kernels for the quickdna sequence hot path.
//
// Byte/integer streaming work throughout (no tensor cores): the frame and reverse-complement kernels are HBM-bound, the
// window kernels issue- or shared-memory-bound.  Design, data layout and the algorithmic bytes per unit are described in
// DESIGN.md; reference citations are relative to
// the SecureDNA/quickdna checkout.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace qd {

// ------------------------------------------------------------------------------------------------
// Geometry
// ------------------------------------------------------------------------------------------------
constexpr int RUN_NT = 48;        // nucleotides per run: 16 residues per frame = one 128-bit store
#ifndef QD_ST_HINT
#define QD_ST_HINT 0  // 0: st.global.L1::no_allocate, 1: st.global.cs, 2: st.global
#endif
// A tile = one run per thread of the CTA.  Two pipeline forms (frames_decoupled()):
//   lockstep   one __syncthreads per tile; thread 0 refills the landing buffer behind it (six frames)
//   decoupled  no CTA-wide barrier in the steady state: a warp that holds its bytes of a landing buffer in registers
//              counts itself off on the buffer, and the LAST warp to do so issues the buffer's next bulk copy -- no
//              dedicated producer, nobody waits on an "empty" barrier, warps drift apart (one and three frames; r01:
//              with a __syncthreads per tile "barrier" was the top stall reason of those kernels)
// Measured on B200 (tools/kbench.cu): for the 1 read : 2 write mix of the six-frame kernel, ONE 1024-thread CTA per SM
// (48 KB contiguous in, 96 KB out per tile) reaches 5.75 TB/s where four 256-thread CTAs per SM reach 5.05 TB/s; small
// inputs use the 256-thread form so they still spread over the SMs.
constexpr int BIG_THREADS = 1024;
constexpr int SMALL_THREADS = 256;
// TMA landing buffers per CTA (three measured no better once the refill left the critical path; the one-frame kernel's
// 37 KB codon table leaves no room for a third anyway)
#ifndef QD_FRAME_STAGES
#define QD_FRAME_STAGES 2
#endif
__host__ __device__ constexpr int frames_stages(int /*frames*/) { return QD_FRAME_STAGES; }
// Which pipeline form a frame count uses: all of them the decoupled one.  Measured on B200, 4 M x 1000-nt reads, fraction of
// the 6548.5 GB/s copy peak, lockstep -> decoupled (tools/kbench.cu, also after other kernels / idle time in between):
// one frame 0.79 -> 0.86, three frames 0.77 -> 0.80, six frames 0.90 -> 0.93.  -DQD_FRAMES_LOCKSTEP=1 builds the lockstep
// form for A/B runs.
#ifndef QD_FRAMES_LOCKSTEP
#define QD_FRAMES_LOCKSTEP 0
#endif
__host__ __device__ constexpr bool frames_decoupled(int /*frames*/) { return !QD_FRAMES_LOCKSTEP; }
template <int THREADS, bool SPEC = false>
struct TileGeom {
    static constexpr int RUNS = THREADS;                           // runs per tile: one per thread
    static constexpr int WQ = RUNS / 32;                           // consumer warps
    static constexpr int WQ_STRIDE = ((THREADS / 32) + 3) & ~3;    // entries of the 32-run map per tile (16-byte aligned slices)
    static constexpr int MAX_BYTES = RUNS * RUN_NT + 2 + 15 + 15;  // span + halo + alignment slack
    static constexpr int MAX_CHUNKS = (MAX_BYTES + 15) / 16 + 1;
    static constexpr int RAW_BYTES = MAX_CHUNKS * 16 + 64;         // one staged tile (+ slack: a run reads 80 B)
    static constexpr int SEARCH_CAP = RUNS;                        // sequences of one tile whose geometry is cached in smem
    static constexpr int CTAS_PER_SM = THREADS >= 1024 ? 1 : (THREADS >= 512 ? 2 : 4);
};

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
#if QD_ST_HINT == 1
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
#elif QD_ST_HINT == 2
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
#else
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
#endif
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
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((uint16_t)v) : "memory"); }
__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 lds_u64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds_u128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// d = c + a.lo16 * b.byte0 + a.hi16 * b.byte1   (.hi: b.byte2, b.byte3) -- 16-bit coefficients, 8-bit data
__device__ __forceinline__ uint32_t dp2a_lo(uint32_t coef, uint32_t data, uint32_t acc) {
    uint32_t r;
    asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(coef), "r"(data), "r"(acc));
    return r;
}
__device__ __forceinline__ uint32_t dp2a_hi(uint32_t coef, uint32_t data, uint32_t acc) {
    uint32_t r;
    asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(coef), "r"(data), "r"(acc));
    return r;
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
#ifndef QD_WAIT_MODE
#define QD_WAIT_MODE 0  // 0: try_wait in a tight loop; 1: nanosleep back-off between polls; 2: try_wait with a suspend-time hint
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#if QD_WAIT_MODE == 1
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(64);
    }
#elif QD_WAIT_MODE == 2
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(0x989680u)
        : "memory");
#else
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
#endif
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// one elected lane of a converged warp (the pattern ptxas recognises: a uniform-datapath instruction behind it needs no loop
// over the lanes' operand values)
__device__ __forceinline__ bool elect_one_lane() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void bulk_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
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
// Programmatic dependent launch: a kernel launched with the programmatic-stream-serialization attribute (launch_pdl in
// qd_api.cu) may become resident while its predecessor in the stream is still running -- after every CTA of the predecessor
// has called pdl_trigger() or exited -- and must call pdl_wait() before it touches anything the predecessor wrote (the wait
// returns once the predecessor has completed and its writes are visible).  Both are no-ops in a plain launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

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
    const uint32_t* warp_first;  // variable mode: sequence holding run 32 q, for every q (bounds each thread's search to 32 entries)
    uint64_t n_seq;
    uint64_t uniform_len;        // sequence length in uniform mode
    uint64_t uniform_magic;      // ceil(2^64 / uniform_runs): g / uniform_runs == umul64hi(g, magic) for g < 2^32
    uint32_t uniform_runs;       // ceil(uniform_len / 48)
    uint32_t uniform_mod3;       // uniform_len % 3
    uint64_t total_runs;         // uniform mode; variable mode reads run_prefix[n_seq]
    // uniform mode, short sequences: a tile is `seqs_per_tile` WHOLE sequences (0 = a tile is TILE_RUNS consecutive runs of
    // the batch).  Thread t of a tile then owns run t % runs of the tile's sequence t / runs -- 32-bit arithmetic, no
    // 64-bit quotient -- at the price of TILE_RUNS - seqs_per_tile * runs idle threads (1.6 % for 1000-nt reads).
    uint32_t seqs_per_tile;
    uint32_t tile_magic;         // ceil(2^20 / uniform_runs): t / runs == (t * tile_magic) >> 20 for t < 1024
    uint32_t n_tiles;            // upper bound on tiles (tiles past total_runs exit)
    uint8_t* out;                // slot layout, 16-byte aligned
    const uint16_t* lut;         // LUT_ENTRIES entries of the selected table, lut_index() layout
    const uint8_t* decode;       // 256-byte ASCII (or mask) -> internal code table (exact error search only)
    const uint16_t* pair;        // PAIR_ENTRIES letter-pair entries for (encoding, strict)
    const uint8_t* direct;       // one-frame kernel: DIRECT_ENTRIES bytes for (table, encoding, strict)
    uint32_t direct_k1, direct_k2;
    uint32_t pair_coef;          // dp2a coefficients {2, 2 * k1} of the pair table
    uint32_t and_need;           // bits every input byte must have (ASCII letters: 0x40 per byte)
    uint32_t or_forbid;          // bits no input byte may have (ASCII: 0x80, mask: 0xe0 per byte)
    unsigned long long* err_key;
    uint64_t flat_base;          // value of offsets[] that corresponds to dna[0] (CSR chunks)
    uint64_t key_base;           // added to a flat offset to form the error key (chunked host calls)
    int strict;
    unsigned long long* trace;   // diagnostics (tools/kbench.cu): per CTA 8 words {start ns, end ns, SM id, tiles done, producer wait-empty ns, stage ns, warp 0 / warp 16 wait-full ns}, or nullptr
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

// Pair table: 5-bit letter pair (u0, u1), u = byte & 31, at index u0 + PAIR_K1 * u1 -> two "doubled code" bytes
// (2 * code | PAIR_FLAG).  flag = the letter is invalid (or ambiguous in a strict table); the code of a VALID letter is
// always exact, whatever its neighbour is.  The strides were picked by search so that the 16 concrete pairs fall in
// 16 distinct shared-memory banks (random ACGT never conflicts); 31 * (1 + k1) + 1 <= PAIR_ENTRIES.
constexpr uint32_t PAIR_K1_ASCII = 40, PAIR_K1_MASK = 38;
constexpr uint32_t PAIR_ENTRIES = 1280;
// Direct codon table of the one-frame kernel: three 5-bit letters -> amino acid, at u0 + K1 u1 + K2 u2 (0x80 = a letter is
// invalid / ambiguous in a strict table).  Strides found by search: injective on every valid codon, never aliasing an
// invalid one, and about two shared-memory wavefronts per lookup on random ACGT.
constexpr uint32_t DIRECT_K1_ASCII = 35, DIRECT_K2_ASCII = 1169, DIRECT_K1_MASK = 31, DIRECT_K2_MASK = 999;
constexpr uint32_t DIRECT_ENTRIES = 37376;  // >= 31 * (1 + K1 + K2) + 1, a multiple of 16
constexpr uint32_t DIRECT_INVALID = 0x80;
// The flag sits above the doubled code (bits 1..4) so a flagged letter still forms an even, in-allocation LUT address
// (its residue is garbage, but a flagged letter inside a validated range always ends in an error).
constexpr uint32_t PAIR_FLAG = 0x20;

// Exact first-error search over one run (only threads whose conservative flags fired get here).
__device__ __noinline__ void exact_check_run(const uint8_t* __restrict__ run_bytes /* global */, int vcnt, const uint8_t* __restrict__ decode, int strict,
                                             uint64_t key0, unsigned long long* err_key) {
    if ((unsigned long long)key0 >= *reinterpret_cast<volatile unsigned long long*>(err_key)) return;  // cannot lower the minimum
    for (int i = 0; i < vcnt; i++) {
        const uint32_t c = __ldg(decode + __ldg(run_bytes + i));
        if (c == (uint32_t)CODE_INVALID || (strict && c >= 4u)) {
            atomicMin(err_key, (unsigned long long)(key0 + i));
            return;
        }
    }
}

// Shared memory of one CTA.  raw[] is the TMA landing zone: while tile T is translated, the bytes of the CTA's next
// tiles stream into the other buffer (cp.async.bulk + mbarrier), so DRAM latency hides behind arithmetic.
// what the next bulk copies into a landing buffer are (plan_tile / issue_tile)
struct TilePlan {
    uint64_t lo, hi;      // flat [start, end) of the tile's input span, 2-nt halo included
    uint64_t tile;
    uint32_t first_seq;   // variable-length batches: the tile's first sequence
    uint32_t pad_;
};

template <int THREADS, int STAGES>
struct FramesSmem {
    alignas(16) uint16_t lut[LUT_ENTRIES];   // codon -> (forward aa, reverse aa)
    alignas(16) uint16_t pair[PAIR_ENTRIES]; // letter pair -> two doubled codes
    alignas(16) uint4 lowmask[17];           // lowmask[k]: the first k bytes of a 16-byte vector are 0xff
    alignas(128) uint8_t raw[STAGES][TileGeom<THREADS>::RAW_BYTES];  // multi-buffered raw input of a tile, 2-nt halo included
    alignas(8) uint64_t full[STAGES];             // mbarriers: raw[b] has landed
    uint32_t released[STAGES];                    // decoupled form: warps that hold their bytes of raw[b] in registers
    alignas(16) TilePlan plan[STAGES];            // the refill of raw[b], planned by the first warp to release it
    uint64_t span[STAGES][2];                     // flat [start, end) of the input span staged in raw[b]
    // variable-length batches: run prefix and CSR offsets of the sequences a tile touches and the per-warp sequence map,
    // bulk-copied together with the tile's bytes (same mbarrier), so no dependent global load sits between two tiles
    alignas(16) uint64_t rp_cache[STAGES][TileGeom<THREADS>::SEARCH_CAP + 6];
    alignas(16) uint64_t off_cache[STAGES][TileGeom<THREADS>::SEARCH_CAP + 6];
    alignas(16) uint32_t warp_first[STAGES][TileGeom<THREADS>::WQ_STRIDE];
    uint32_t first_seq[STAGES];  // first sequence of the tile staged in buffer b
    // The kernel parameters once more, for the (not inlined) staging function: handing it the __grid_constant__ struct by
    // reference made every thread keep a 200-byte copy in LOCAL memory, and the staging thread read its fields back with a
    // chain of local loads per tile -- L2 or DRAM round trips on the critical path of the TMA pipeline (ncu: the consumers
    // starved on the full barriers, DRAM at 29 %; how bad depended on what L2 happened to hold).
    alignas(16) FramesParams params;
};

using FramesSmemSmallest = FramesSmem<SMALL_THREADS, 2>;
static_assert(offsetof(FramesSmemSmallest, lut) == 0 && (2 * 14 + PAIR_FLAG) * (1 + LUT_K1 + LUT_K2) + 2 <= sizeof(FramesSmemSmallest),
              "a flagged codon must still address shared memory of this CTA");

// flat offset of the first run of tile t (t * RUNS < total_runs)
template <int THREADS, bool SPEC>
__device__ __forceinline__ uint64_t tile_flat_start(const FramesParams& p, uint64_t t) {
    if (p.tiles) return p.tiles[t].flat;
    if (p.seqs_per_tile) return min(p.total_nt, t * p.seqs_per_tile * p.uniform_len);
    const uint32_t g = (uint32_t)(t * TileGeom<THREADS, SPEC>::RUNS);
    const uint32_t r = p.uniform_runs == 1 ? g : (uint32_t)__umul64hi((uint64_t)g, p.uniform_magic);
    return (uint64_t)r * p.uniform_len + (uint64_t)(g - r * p.uniform_runs) * RUN_NT;
}

// One thread: publish the span of tile t in S.span[b] and start its bulk copy into S.raw[b].
// Whole 16-byte granules inside [dna, dna + total) are copied by TMA; the (at most two) granules that straddle the
// ends of the whole buffer are copied byte-wise by this thread, so nothing outside the buffer is ever read.
// Staging a tile has two halves.  plan_tile works out WHAT to copy (a 64-bit quotient or, for variable-length batches,
// dependent global loads): anybody may do it at any time, it touches only S.plan[b].  issue_tile starts the copies
// described by S.plan[b] into landing buffer b: a handful of instructions, to be run once the buffer is free.
template <int THREADS, bool SPEC, int STAGES>
__device__ __noinline__ void plan_tile(FramesSmem<THREADS, STAGES>& S, const FramesParams& p, uint64_t t, uint64_t total_runs, int b) {
    constexpr int TILE_RUNS = TileGeom<THREADS, SPEC>::RUNS;
    const uint64_t lo = tile_flat_start<THREADS, SPEC>(p, t);
    uint64_t hi = p.total_nt;
    if (p.seqs_per_tile ? (t + 1) * p.seqs_per_tile < p.n_seq : (t + 1) * TILE_RUNS < total_runs)
        hi = min(p.total_nt, tile_flat_start<THREADS, SPEC>(p, t + 1) + 2);  // 2-nt halo
    TilePlan& P = S.plan[b];
    P.lo = lo;
    P.hi = hi;
    P.tile = t;
    P.first_seq = p.tiles ? p.tiles[t].first_seq : 0u;
}

template <int THREADS, bool SPEC, int STAGES>
__device__ __forceinline__ void issue_tile(FramesSmem<THREADS, STAGES>& S, const FramesParams& p, int b) {
    const TilePlan P = S.plan[b];
    const uint64_t lo = P.lo, hi = P.hi;
    S.span[b][0] = lo;
    S.span[b][1] = hi;
    const uintptr_t dna_lo = reinterpret_cast<uintptr_t>(p.dna), dna_hi = dna_lo + p.total_nt;
    const uintptr_t a0 = (dna_lo + lo) & ~(uintptr_t)15;
    const uintptr_t a1 = (dna_lo + hi + 15) & ~(uintptr_t)15;
    const uintptr_t t0 = max(a0, (dna_lo + 15) & ~(uintptr_t)15);
    const uintptr_t t1 = min(a1, dna_hi & ~(uintptr_t)15);
    if (a0 < t0 || t1 < a1) {  // edge granules of the whole buffer (the first and the last tile of a launch only)
#pragma unroll 1
        for (uintptr_t a = a0; a < min(t0, a1); a++) S.raw[b][a - a0] = (a >= dna_lo && a < dna_hi) ? *reinterpret_cast<const uint8_t*>(a) : (uint8_t)0;
#pragma unroll 1
        for (uintptr_t a = max(t1, a0); a < a1; a++) S.raw[b][a - a0] = (a >= dna_lo && a < dna_hi) ? *reinterpret_cast<const uint8_t*>(a) : (uint8_t)0;
    }
    const uint32_t bytes = t1 > t0 ? (uint32_t)(t1 - t0) : 0u;
    if (p.tiles) {
        // geometry of a variable-length tile: windows of run_prefix / offsets starting at the tile's first sequence (the host
        // keeps 16-byte aligned, ~0-padded copies of both arrays, so an even index is an aligned source) and the tile's
        // slice of the 32-run sequence map
        constexpr uint32_t CAP = TileGeom<THREADS>::SEARCH_CAP;
        const uint32_t r_first = P.first_seq, rs = r_first & 1u;
        constexpr uint32_t WIN_BYTES = (CAP + 4) * 8, MAP_BYTES = TileGeom<THREADS>::WQ_STRIDE * 4;
        static_assert(WIN_BYTES % 16 == 0 && MAP_BYTES % 16 == 0, "bulk copies move whole 16-byte granules");
        S.first_seq[b] = r_first;
        mbar_expect_tx(&S.full[b], bytes + 2 * WIN_BYTES + MAP_BYTES);  // release: plain stores above are visible with the phase
        tma_bulk_g2s(S.rp_cache[b], p.run_prefix + (r_first - rs), WIN_BYTES, &S.full[b]);
        tma_bulk_g2s(S.off_cache[b], p.offsets + (r_first - rs), WIN_BYTES, &S.full[b]);
        tma_bulk_g2s(S.warp_first[b], p.warp_first + P.tile * TileGeom<THREADS>::WQ_STRIDE, MAP_BYTES, &S.full[b]);
        if (bytes) tma_bulk_g2s(S.raw[b] + (t0 - a0), reinterpret_cast<const void*>(t0), bytes, &S.full[b]);
    } else if (bytes) {
        mbar_expect_tx(&S.full[b], bytes);  // release: the byte stores above are visible to whoever observes the phase
        tma_bulk_g2s(S.raw[b] + (t0 - a0), reinterpret_cast<const void*>(t0), bytes, &S.full[b]);
    } else {
        mbar_arrive(&S.full[b]);
    }
}

template <int THREADS, bool SPEC, int STAGES>
__device__ __forceinline__ void stage_tile(FramesSmem<THREADS, STAGES>& S, const FramesParams& p, uint64_t t, uint64_t total_runs, int b) {
    plan_tile<THREADS, SPEC, STAGES>(S, p, t, total_runs, b);
    issue_tile<THREADS, SPEC, STAGES>(S, p, b);
}

template <int FRAMES, int THREADS, bool DIRECT = false>
__global__ void __launch_bounds__(THREADS, TileGeom<THREADS>::CTAS_PER_SM) translate_frames_kernel(const FramesParams p) {
    static_assert(!DIRECT || FRAMES == 1, "the direct codon table serves the one-frame kernel only");
    constexpr bool SPEC = frames_decoupled(FRAMES);
    constexpr int STAGES = frames_stages(FRAMES);
    constexpr int TILE_THREADS = THREADS, TILE_RUNS = TileGeom<THREADS, SPEC>::RUNS, SEARCH_CAP = TileGeom<THREADS>::SEARCH_CAP;
    extern __shared__ __align__(128) unsigned char frames_smem_raw[];  // sizeof(FramesSmem<THREADS, STAGES>) dynamic bytes
    using Smem = FramesSmem<THREADS, frames_stages(FRAMES)>;
    Smem& S = *reinterpret_cast<Smem*>(frames_smem_raw);
    uint8_t* const direct_s = frames_smem_raw + ((sizeof(Smem) + 127) & ~(size_t)127);  // DIRECT only
    const int tid = threadIdx.x;
    const bool uniform = (p.offsets == nullptr);
    pdl_wait();  // variable-length batches launch this kernel behind the map kernels as a dependent (no-op otherwise)
    if (p.trace && tid == 0) {
        unsigned long long t0;
        uint32_t smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        p.trace[8 * blockIdx.x] = t0;
        p.trace[8 * blockIdx.x + 2] = smid;
    }
    const uint64_t total_runs = uniform ? p.total_runs : p.run_prefix[p.n_seq];
    const uint64_t n_tiles = (uniform && p.seqs_per_tile) ? (uint64_t)p.n_tiles : min((uint64_t)p.n_tiles, (total_runs + TILE_RUNS - 1) / TILE_RUNS);

    // Per-CTA tables (persistent CTA: loaded once, reused for every tile)
    if (tid == 0) {
        S.params = p;
        for (int b = 0; b < STAGES; b++) {
            mbar_init(&S.full[b], 1);
            S.released[b] = 0;
        }
        fence_barrier_init();
        {   // the first STAGES tiles
            for (int b = 0; b < STAGES; b++)
                if (blockIdx.x + b * (uint64_t)gridDim.x < n_tiles) stage_tile<THREADS, SPEC, STAGES>(S, S.params, blockIdx.x + b * (uint64_t)gridDim.x, total_runs, b);
        }
    }
    if (DIRECT) {
        const uint4* src = reinterpret_cast<const uint4*>(p.direct);
        uint4* dst = reinterpret_cast<uint4*>(direct_s);
        for (int i = tid; i < (int)(DIRECT_ENTRIES / 16); i += TILE_THREADS) dst[i] = ldg_cached(src + i);
    } else {
        const uint4* src = reinterpret_cast<const uint4*>(p.lut);
        uint4* dst = reinterpret_cast<uint4*>(S.lut);
        for (int i = tid; i < (int)(LUT_ENTRIES * 2 / 16); i += TILE_THREADS) dst[i] = ldg_cached(src + i);
        const uint4* psrc = reinterpret_cast<const uint4*>(p.pair);
        uint4* pdst = reinterpret_cast<uint4*>(S.pair);
        for (int i = tid; i < (int)(PAIR_ENTRIES * 2 / 16); i += TILE_THREADS) pdst[i] = ldg_cached(psrc + i);
    }
    if (tid < 17) {
        auto word = [](int k) { return k >= 4 ? 0xffffffffu : (k <= 0 ? 0u : (1u << (8 * k)) - 1u); };
        S.lowmask[tid] = make_uint4(word(tid), word(tid - 4), word(tid - 8), word(tid - 12));
    }
    const uintptr_t dna_lo = reinterpret_cast<uintptr_t>(p.dna);
    // shared-window addresses of the two tables: folded into the address-forming dot products as the accumulator,
    // so a lookup is [dot product, load] with no separate add
    const uint32_t lut_base = smem_u32(S.lut);
    const uint32_t pair_base = smem_u32(S.pair);
    const uint32_t pair_coef = p.pair_coef;
    __syncthreads();  // the only CTA-wide barrier

    uint32_t it = 0;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int buf = it % STAGES;
        const bool aligned_tiles = uniform && p.seqs_per_tile != 0;
        const uint64_t g0 = tile * TILE_RUNS;
        const uint32_t cnt = aligned_tiles ? (uint32_t)min((uint64_t)p.seqs_per_tile, p.n_seq - tile * p.seqs_per_tile) * p.uniform_runs
                                           : (uint32_t)min((uint64_t)TILE_RUNS, total_runs - g0);

        // ---- phase 0: which run is mine --------------------------------------------------------
        RunDesc rd;
        const bool active = tid < (int)cnt;
        if (aligned_tiles) {
            // whole sequences per tile: everything is 32-bit (sequences here are at most 1024 runs = 49 152 nt long)
            const uint32_t t = active ? (uint32_t)tid : 0u;
            const uint32_t rl = (t * p.tile_magic) >> 20;
            const uint32_t r = (uint32_t)tile * p.seqs_per_tile + rl;  // < n_seq < 2^32
            const uint32_t len = (uint32_t)p.uniform_len;
            rd.m = t - rl * p.uniform_runs;
            rd.runs = p.uniform_runs;
            rd.mod3 = p.uniform_mod3;
            rd.tiny = len < 3u;
            const uint32_t in_seq = rd.m * RUN_NT;
            rd.left_nt = len - in_seq;
            rd.flat_start = (uint64_t)r * len + in_seq;
            rd.out_base = (uint64_t)r * (p.uniform_runs * (16u * FRAMES));
        } else if (uniform) {
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
            // the geometry windows of this tile arrived with its bytes (same mbarrier)
            mbar_wait(&S.full[buf], (it / STAGES) & 1);
            const uint32_t first_cur = S.first_seq[buf];
            const uint64_t* rpc = S.rp_cache[buf] + (first_cur & 1u);
            const uint64_t* offc = S.off_cache[buf] + (first_cur & 1u);
            const uint32_t gl = active ? tid : 0;  // run index relative to g0
            uint64_t g = g0 + gl;
            // largest i with run_prefix[r_first + i] <= g: it is at or after the sequence of my warp's first run, and at most
            // 32 entries further unless empty sequences intervene (then the whole window is searched)
            uint32_t lo = 0, hi = SEARCH_CAP;
            bool found = false;
            if ((uint32_t)(tid & ~31) < cnt) {  // (warp-uniform) my warp has runs in this tile
                // Together: lane L looks at the L-th sequence after the one holding the warp's first run.  Those that start
                // inside the warp's 32 runs set a bit at their first run; my sequence is as many steps further as there are
                // bits at or below my run.  (Empty sequences repeat a prefix value and would share a bit: search instead.)
                const uint32_t lo0 = S.warp_first[buf][tid >> 5] - first_cur, lane = tid & 31;
                if (lo0 + 33u <= (uint32_t)SEARCH_CAP) {
                    const uint64_t gw = g0 + (uint32_t)(tid & ~31);
                    const uint64_t v = rpc[lo0 + 1 + lane], vnext = rpc[lo0 + 2 + lane];  // all > gw: warp_first is the LAST sequence at or before run gw
                    const bool inside = v <= gw + 31;
                    const uint32_t repeats = __ballot_sync(0xffffffffu, inside && v == vnext);
                    const uint32_t starts = __reduce_or_sync(0xffffffffu, inside ? 1u << (uint32_t)(v - gw) : 0u);
                    const bool all_seen = __shfl_sync(0xffffffffu, inside ? 0 : 1, 31) != 0;  // the 32nd candidate starts past the warp's runs
                    if (repeats == 0 && all_seen) {
                        lo = lo0 + __popc(starts & (0xffffffffu >> (31u - lane)));
                        if (!active) {  // lanes past the last run: the geometry of the warp's first run (nothing is written)
                            lo = lo0;
                            g = gw;
                        }
                        found = true;
                    }
                }
            }
            if (!found) {
                if (active) {
                    lo = S.warp_first[buf][tid >> 5] - first_cur;
                    const uint32_t near = min(lo + 32u, (uint32_t)SEARCH_CAP);
                    if (lo > (uint32_t)SEARCH_CAP) lo = SEARCH_CAP;  // window too small for this tile: global search below
                    else if (rpc[near] > g) hi = near;
                }
                while (lo < hi) {
                    const uint32_t mid = (lo + hi + 1) >> 1;
                    if (rpc[mid] <= g) lo = mid; else hi = mid - 1;
                }
            }
            uint64_t rp = rpc[lo], off_lo = offc[lo], off_hi = offc[lo + 1];
            if (lo == SEARCH_CAP) {  // more than SEARCH_CAP sequences start inside this tile: finish in global
                uint64_t a = (uint64_t)first_cur + lo, b = p.n_seq;  // invariant: run_prefix[a] <= g
                while (a < b) {
                    uint64_t mid = (a + b + 1) >> 1;
                    if (p.run_prefix[mid] <= g) a = mid; else b = mid - 1;
                }
                rp = p.run_prefix[a];
                off_lo = p.offsets[a];
                off_hi = p.offsets[a + 1];
            }
            const uint64_t o0 = off_lo - p.flat_base;
            const uint64_t n = off_hi - off_lo;
            rd.m = (uint32_t)(g - rp);
            if ((n >> 32) == 0) {  // the usual case: 32-bit arithmetic
                const uint32_t n32 = (uint32_t)n, in_seq = rd.m * RUN_NT;
                rd.runs = (n32 + RUN_NT - 1) / RUN_NT;
                rd.mod3 = n32 % 3u;
                rd.left_nt = n32 - in_seq;
                rd.flat_start = o0 + in_seq;
            } else {
                const uint64_t in_seq = (uint64_t)rd.m * RUN_NT;
                rd.runs = (uint32_t)((n + RUN_NT - 1) / RUN_NT);
                rd.mod3 = (uint32_t)(n % 3);
                rd.left_nt = n - in_seq;
                rd.flat_start = o0 + in_seq;
            }
            rd.tiny = n < 3;
            rd.out_base = rp * (16u * FRAMES);
        }

        // ---- phase 1: my run's 52 raw bytes, from the staged tile into registers ------------------
#ifdef QD_TRACE_DETAIL
        unsigned long long tw0, tw1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tw0));
#endif
        mbar_wait(&S.full[buf], (it / STAGES) & 1);
#ifdef QD_TRACE_DETAIL
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tw1));
        if (p.trace && (tid == 0 || tid == 512)) atomicAdd(&p.trace[8 * blockIdx.x + (tid ? 7 : 6)], tw1 - tw0);
#endif
        const uintptr_t a0 = (dna_lo + S.span[buf][0]) & ~(uintptr_t)15;
        const uint32_t o = (uint32_t)(dna_lo + rd.flat_start - a0);  // local byte offset of the run
        uint32_t W[13];
        if (__all_sync(0xffffffffu, (o & 7u) == 0u)) {
            // every run of the warp starts on an 8-byte boundary of the staged tile (sequences whose length is a multiple of
            // 8 in an 8-byte aligned buffer -- 1000-nt reads): seven 64-bit loads, no realignment
            const uint32_t sa = smem_u32(S.raw[buf]) + o;
#pragma unroll
            for (int k = 0; k < 6; k++) {
                const uint2 v = lds_u64(sa + 8 * k);
                W[2 * k] = v.x;
                W[2 * k + 1] = v.y;
            }
            W[12] = lds_u32(sa + 48);
        } else {
            const uint4* src = reinterpret_cast<const uint4*>(S.raw[buf] + (o & ~15u));
            const uint4 v0 = src[0], v1 = src[1], v2 = src[2], v3 = src[3], v4 = src[4];
            const uint32_t w[20] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y,
                                    v2.z, v2.w, v3.x, v3.y, v3.z, v3.w, v4.x, v4.y, v4.z, v4.w};
            const uint32_t bs = (o & 3u) * 8u;
            switch ((o >> 2) & 3u) {  // the same for every run of a sequence (48 = 0 mod 16)
                case 0:
#pragma unroll
                    for (int k = 0; k < 13; k++) W[k] = __funnelshift_r(w[k], w[k + 1], bs);
                    break;
                case 1:
#pragma unroll
                    for (int k = 0; k < 13; k++) W[k] = __funnelshift_r(w[k + 1], w[k + 2], bs);
                    break;
                case 2:
#pragma unroll
                    for (int k = 0; k < 13; k++) W[k] = __funnelshift_r(w[k + 2], w[k + 3], bs);
                    break;
                default:
#pragma unroll
                    for (int k = 0; k < 13; k++) W[k] = __funnelshift_r(w[k + 3], w[k + 4], bs);
                    break;
            }
        }
        if (SPEC) {
            // every lane holds its bytes (and is done with the tile's geometry windows): the warp counts itself off on raw[buf].
            // A warp that finds the count still at zero PLANS the buffer's refill before it counts itself off (it is ahead of
            // the others: the planning -- a 64-bit quotient, or dependent global loads for variable-length batches -- costs
            // nobody anything; several early warps may plan, they write the same values); the LAST warp to count itself off
            // knows the buffer is free and only has to issue the planned copies.  Release / acquire runs through the counter.
            __syncwarp();
            if ((tid & 31) == 0) {
                const uint64_t next = tile + STAGES * (uint64_t)gridDim.x;
                if (next < n_tiles && *reinterpret_cast<volatile uint32_t*>(&S.released[buf]) == 0u)
                    plan_tile<THREADS, SPEC, STAGES>(S, S.params, next, total_runs, buf);
                uint32_t before;
                asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(before) : "r"(smem_u32(&S.released[buf])) : "memory");
                if (before == (uint32_t)(THREADS / 32 - 1)) {
                    S.released[buf] = 0;  // published with the refill's mbarrier phase
                    if (next < n_tiles) issue_tile<THREADS, SPEC, STAGES>(S, S.params, buf);
                }
            }
        } else {
            __syncthreads();  // every thread holds its bytes: raw[buf] can be refilled
            if (tid == 0 && tile + STAGES * (uint64_t)gridDim.x < n_tiles) stage_tile<THREADS, SPEC, STAGES>(S, S.params, tile + STAGES * (uint64_t)gridDim.x, total_runs, buf);
        }

        // ---- phase 2: one run per thread: 48 codon starts -> 16 residues per frame per strand -----
        if (DIRECT && active) {
            // one frame: 16 codons, each three letters -> ONE table byte; the index comes from two dp2a over the letters
            // where they lie (codon e starts at byte 3e: within a word, or across two), no realignment, no code stage
            const uint32_t dbase = smem_u32(direct_s);
            const uint32_t c01 = 1u | (p.direct_k1 << 16), c2x = p.direct_k2, cx0 = 1u << 16, c12 = p.direct_k1 | (p.direct_k2 << 16);
            uint32_t U[13], acc_and = 0xffffffffu, acc_or = 0u;
#pragma unroll
            for (int k = 0; k < 13; k++) {
                if (k & 1) {
                    acc_and &= W[k - 1] & W[k];
                    acc_or |= W[k - 1] | W[k];
                } else if (k == 12) {
                    acc_and &= W[k];
                    acc_or |= W[k];
                }
                U[k] = W[k] & 0x1f1f1f1fu;
            }
            uint32_t F[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint32_t r4[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int j = 3 * (4 * q + e), wq = j >> 2, rb = j & 3;
                    uint32_t a;
                    if (rb == 0) a = dp2a_hi(c2x, U[wq], dp2a_lo(c01, U[wq], dbase));
                    else if (rb == 1) a = dp2a_hi(c12, U[wq], dp2a_lo(cx0, U[wq], dbase));
                    else if (rb == 2) a = dp2a_lo(c2x, U[wq + 1], dp2a_hi(c01, U[wq], dbase));
                    else a = dp2a_lo(c12, U[wq + 1], dp2a_hi(cx0, U[wq], dbase));
                    r4[e] = lds_u8(a);
                }
                F[q] = __byte_perm(__byte_perm(r4[0], r4[1], 0x0040), __byte_perm(r4[2], r4[3], 0x0040), 0x5410);
            }
            const bool suspicious = ((F[0] | F[1] | F[2] | F[3]) & (DIRECT_INVALID * 0x01010101u)) | ((acc_and & p.and_need) ^ p.and_need) | (acc_or & p.or_forbid);
            if (suspicious) {
                uint64_t vleft = rd.left_nt >= rd.mod3 ? rd.left_nt - rd.mod3 : 0;  // translate validates 3*floor(n/3) bytes
                if (vleft) exact_check_run(p.dna + rd.flat_start, (int)min((uint64_t)RUN_NT, vleft), p.decode, p.strict, p.key_base + rd.flat_start, p.err_key);
            }
            uint4 fv = make_uint4(F[0], F[1], F[2], F[3]);
            if (rd.left_nt < (uint64_t)(RUN_NT + 2)) {
                const uint32_t keep = min(16u, ((uint32_t)rd.left_nt * 171u) >> 9);
                const uint4 lm = S.lowmask[keep];
                fv.x &= lm.x; fv.y &= lm.y; fv.z &= lm.z; fv.w &= lm.w;
            }
            stg_stream(p.out + rd.out_base + (size_t)rd.m * 16u, fv);
        }
        if (!DIRECT && active) {
            // letters -> doubled codes, two letters per table lookup; Bq[q] holds nucleotides 4q..4q+3
            uint32_t Bq[13];
            uint32_t acc_and = 0xffffffffu, acc_or = 0u;
#pragma unroll
            for (int k = 0; k < 13; k++) {
                if (k & 1) {  // one 3-input LOP3 per two words
                    acc_and &= W[k - 1] & W[k];
                    acc_or |= W[k - 1] | W[k];
                } else if (k == 12) {
                    acc_and &= W[k];
                    acc_or |= W[k];
                }
                const uint32_t m = W[k] & 0x1f1f1f1fu;
                const uint32_t e_lo = lds_u16(dp2a_lo(pair_coef, m, pair_base));
                const uint32_t e_hi = lds_u16(dp2a_hi(pair_coef, m, pair_base));
                Bq[k] = e_hi * 65536u + e_lo;
            }
            uint32_t flags = 0;
#pragma unroll
            for (int k = 0; k < 12; k += 2) flags |= Bq[k] | Bq[k + 1];
            flags |= Bq[12];
            // conservative: a flagged letter, or a byte outside the letter range of the encoding
            const bool suspicious = (flags & (PAIR_FLAG * 0x01010101u)) | ((acc_and & p.and_need) ^ p.and_need) | (acc_or & p.or_forbid);
            if (suspicious) {
                // validated range: translate -> first 3*floor(n/3) bytes (src/trans_table.rs:191);
                // frames -> every byte when n >= 3 (union of the frames), none otherwise
                uint64_t vleft = rd.left_nt;  // validated nucleotides from the run start on
                if (FRAMES == 1) vleft = vleft >= rd.mod3 ? vleft - rd.mod3 : 0;
                else if (rd.tiny) vleft = 0;
                if (vleft) exact_check_run(p.dna + rd.flat_start, (int)min((uint64_t)RUN_NT, vleft), p.decode, p.strict, p.key_base + rd.flat_start, p.err_key);
            }
            uint8_t* const out_seq = p.out + rd.out_base;
            const uint32_t slot = rd.runs * 16u;
            uint8_t* const fwd_dst = out_seq + (size_t)rd.m * 16u;
            uint8_t* const rc_dst = out_seq + 4 * (size_t)slot - (size_t)(rd.m + 1) * 16u;
            const bool partial = rd.left_nt < (uint64_t)(RUN_NT + 2);  // last run of its sequence
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
                    // r = fwd | rc << 8
                    const uint32_t t01 = __byte_perm(r4[0], r4[1], 0x5410);  // [f0, r0, f1, r1]
                    const uint32_t t23 = __byte_perm(r4[2], r4[3], 0x5410);  // [f2, r2, f3, r3]
                    F[q] = __byte_perm(t01, t23, 0x6420);                    // f0 f1 f2 f3
                    R[q] = __byte_perm(t01, t23, 0x1357);                    // r3 r2 r1 r0
                }
                uint4 fv = make_uint4(F[0], F[1], F[2], F[3]);
                uint4 rv = make_uint4(R[3], R[2], R[1], R[0]);
                if (partial) {  // codon starts past the end of the sequence produce nothing: those bytes are 0
                    const uint32_t left = max((uint32_t)rd.left_nt, (uint32_t)i) - i;  // < 50
                    const uint32_t keep = min(16u, (left * 171u) >> 9);                // floor(left / 3), exact below 64
                    const uint4 lm = S.lowmask[keep], hm = S.lowmask[16u - keep];
                    fv.x &= lm.x; fv.y &= lm.y; fv.z &= lm.z; fv.w &= lm.w;            // forward: the first `keep` residues
                    rv.x &= ~hm.x; rv.y &= ~hm.y; rv.z &= ~hm.z; rv.w &= ~hm.w;        // reverse: the last `keep` bytes
                }
                stg_stream(fwd_dst + (size_t)i * slot, fv);
                if (FRAMES == 6) {
                    // codon start p = 48m + 3jj + i is reverse frame (n - i) mod 3, residue (n-3-i)/3 - 16m - jj
                    uint32_t kq = rd.mod3 + 3u - i;
                    kq = kq >= 3u ? kq - 3u : kq;
                    stg_stream(rc_dst + (size_t)kq * slot, rv);
                }
            }
        }
    }
    if (p.trace && tid == 0) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        p.trace[8 * blockIdx.x + 1] = t1;
        p.trace[8 * blockIdx.x + 3] = it;
    }
}

// ------------------------------------------------------------------------------------------------
// Reverse complement of one sequence: out[n-1-i] = comp[in[i]]
// ------------------------------------------------------------------------------------------------
struct RevcompParams {
    const uint8_t* dna;
    uint64_t n;
    uint8_t* out;            // 16-byte aligned
    const uint8_t* table;    // 256-byte complement table, 0 = invalid byte (tail bytes and the exact error search)
    const uint16_t* pair;    // PAIR_ENTRIES entries: letters (u0, u1) at u0 + k1 * u1 -> comp(u1) | comp(u0) << 8, invalid -> 0x80
    uint32_t pair_coef;      // dp2a coefficients {2, 2 * k1}
    uint32_t and_need;       // bits every input byte must have
    uint32_t or_forbid;      // bits no input byte may have
    uint32_t filler;         // a valid input byte replicated 4x: stands in for bytes outside the buffer in its edge granules
    unsigned long long* err_key;
    uint64_t key_base;       // added to an input index to form the error key (chunks of a longer sequence)
};

#ifndef QD_RC_THREADS
#define QD_RC_THREADS 1024
#endif
#ifndef QD_RC_UNROLL
#define QD_RC_UNROLL 2
#endif
#ifndef QD_RC_TMA_STORE
#define QD_RC_TMA_STORE 0  // 1: stage a tile's output in shared memory and write it with one cp.async.bulk
#endif
constexpr int RC_BIG_THREADS = QD_RC_THREADS, RC_SMALL_THREADS = 256;
constexpr int RC_STAGES = 2;  // TMA landing buffers of the reverse-complement kernel
constexpr int RC_UNROLL = QD_RC_UNROLL;  // 16-byte output granules per thread and tile
constexpr uint32_t RC_INVALID = 0x80;    // complement-table flag (outputs are ASCII letters or masks: bit 7 is free)

template <int THREADS>
struct RevcompSmem {
    static constexpr int TILE_CHUNKS = THREADS * RC_UNROLL;        // 16-byte output granules per tile
    static constexpr int RAW_BYTES = (TILE_CHUNKS + 1) * 16;       // + one granule for the source misalignment
    alignas(16) uint16_t pair[PAIR_ENTRIES];
    alignas(128) uint8_t raw[RC_STAGES][RAW_BYTES];  // source bytes of a tile (TMA landing zone)
#if QD_RC_TMA_STORE
    alignas(128) uint8_t outb[2][TILE_CHUNKS * 16];  // a tile's output, drained by one bulk store
#endif
    alignas(8) uint64_t full[RC_STAGES];
    alignas(16) RevcompParams params;  // for the staging function (see FramesSmem::params)
};

// Tile t produces output granules [t*TILE_CHUNKS, ...): out[16j, 16j+16) = in[n-16j-16, n-16j) reversed and
// complemented.  Its source is one contiguous span that ends at in_end - 16*j0; `d` (the misalignment of
// in_end) is the same for every granule, so the span is staged as whole 16-byte granules starting at a0.
// Granules that straddle the ends of the whole buffer are filled byte-wise by the staging thread.
template <int THREADS>
__device__ __noinline__ void rc_stage_tile(RevcompSmem<THREADS>& S, const RevcompParams& p, uint64_t t, uint64_t full_chunks, int b) {
    constexpr int TILE_CHUNKS = RevcompSmem<THREADS>::TILE_CHUNKS;
    const uint64_t j0 = t * TILE_CHUNKS;
    const uint32_t cnt = (uint32_t)min((uint64_t)TILE_CHUNKS, full_chunks - j0);
    const uintptr_t dna_lo = reinterpret_cast<uintptr_t>(p.dna), dna_hi = dna_lo + p.n;
    const uint32_t d = (uint32_t)(dna_hi & 15u);
    const uintptr_t a0 = dna_hi - 16 * (j0 + cnt) - d;  // granule holding the tile's lowest source byte
    const uintptr_t a1 = a0 + 16 * (uintptr_t)(cnt + (d ? 1 : 0));
    const uintptr_t t0 = max(a0, (dna_lo + 15) & ~(uintptr_t)15);
    const uintptr_t t1 = min(a1, dna_hi & ~(uintptr_t)15);
    if (a0 < t0 || t1 < a1) {
#pragma unroll 1
        for (uintptr_t a = a0; a < min(t0, a1); a++) S.raw[b][a - a0] = (a >= dna_lo && a < dna_hi) ? *reinterpret_cast<const uint8_t*>(a) : (uint8_t)p.filler;
#pragma unroll 1
        for (uintptr_t a = max(t1, a0); a < a1; a++) S.raw[b][a - a0] = (a >= dna_lo && a < dna_hi) ? *reinterpret_cast<const uint8_t*>(a) : (uint8_t)p.filler;
    }
    if (t1 > t0) {
        const uint32_t bytes = (uint32_t)(t1 - t0);
        mbar_expect_tx(&S.full[b], bytes);
        tma_bulk_g2s(S.raw[b] + (t0 - a0), reinterpret_cast<const void*>(t0), bytes, &S.full[b]);
    } else {
        mbar_arrive(&S.full[b]);
    }
}

// input word (bytes b0..b3, ascending address) -> output word whose ascending bytes are c(b3) c(b2) c(b1) c(b0):
// two letter-pair lookups, each returning its two complemented letters already swapped
__device__ __forceinline__ uint32_t rc_map_word(uint32_t w, uint32_t coef, uint32_t base) {
    const uint32_t m = w & 0x1f1f1f1fu;
    const uint32_t e_lo = lds_u16(dp2a_lo(coef, m, base));  // c(b1) | c(b0) << 8
    const uint32_t e_hi = lds_u16(dp2a_hi(coef, m, base));  // c(b3) | c(b2) << 8
    return e_lo * 65536u + e_hi;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS >= 1024 ? 1 : 4) revcomp_kernel(const RevcompParams p) {
    constexpr int TILE_CHUNKS = RevcompSmem<THREADS>::TILE_CHUNKS;
    extern __shared__ __align__(128) unsigned char rc_smem_raw[];
    RevcompSmem<THREADS>& S = *reinterpret_cast<RevcompSmem<THREADS>*>(rc_smem_raw);
    const int tid = threadIdx.x;
    const uint64_t n = p.n;
    const uint64_t full = n >> 4;  // full 16-byte output granules
    const uint64_t n_tiles = (full + TILE_CHUNKS - 1) / TILE_CHUNKS;
    if (tid == 0) {
        S.params = p;
        for (int b = 0; b < RC_STAGES; b++) mbar_init(&S.full[b], 1);
        fence_barrier_init();
        for (int b = 0; b < RC_STAGES; b++)
            if (blockIdx.x + b * (uint64_t)gridDim.x < n_tiles) rc_stage_tile(S, S.params, blockIdx.x + b * (uint64_t)gridDim.x, full, b);
    }
    {
        const uint4* psrc = reinterpret_cast<const uint4*>(p.pair);
        uint4* pdst = reinterpret_cast<uint4*>(S.pair);
        for (int i = tid; i < (int)(PAIR_ENTRIES * 2 / 16); i += THREADS) pdst[i] = ldg_cached(psrc + i);
    }
    __syncthreads();
    const uint32_t pair_base = smem_u32(S.pair), coef = p.pair_coef;
    const uintptr_t dna_hi = reinterpret_cast<uintptr_t>(p.dna) + n;
    const uint32_t d = (uint32_t)(dna_hi & 15u);  // misalignment of every granule's source
    const uint32_t ws = d >> 2, bs = (d & 3u) * 8u;

    uint32_t it = 0;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int buf = it % RC_STAGES;
        const uint64_t j0 = tile * TILE_CHUNKS;
        const uint32_t cnt = (uint32_t)min((uint64_t)TILE_CHUNKS, full - j0);
        mbar_wait(&S.full[buf], (it / RC_STAGES) & 1);
        // ---- my granules' 16 source bytes each, into registers ----------------------------------------
        uint32_t x[RC_UNROLL][4];
#pragma unroll
        for (int u = 0; u < RC_UNROLL; u++) {
            const uint32_t l = (uint32_t)u * THREADS + tid;  // output granule within the tile
            const uint32_t g = l < cnt ? cnt - 1 - l : 0;    // staged granule holding in[n-16j-16] (at byte offset d)
            const uint4* src = reinterpret_cast<const uint4*>(S.raw[buf]) + g;
            const uint4 lo = src[0];
            if (d == 0) {
                x[u][0] = lo.x; x[u][1] = lo.y; x[u][2] = lo.z; x[u][3] = lo.w;
            } else {
                const uint4 hi = src[1];
                const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
                switch (ws) {  // uniform across the grid
                    case 0:
#pragma unroll
                        for (int k = 0; k < 4; k++) x[u][k] = __funnelshift_r(w[k], w[k + 1], bs);
                        break;
                    case 1:
#pragma unroll
                        for (int k = 0; k < 4; k++) x[u][k] = __funnelshift_r(w[k + 1], w[k + 2], bs);
                        break;
                    case 2:
#pragma unroll
                        for (int k = 0; k < 4; k++) x[u][k] = __funnelshift_r(w[k + 2], w[k + 3], bs);
                        break;
                    default:
#pragma unroll
                        for (int k = 0; k < 4; k++) x[u][k] = __funnelshift_r(w[k + 3], w[(k + 4) & 7], bs);
                        break;
                }
            }
        }
        __syncthreads();  // every thread holds its bytes: raw[buf] can be refilled
        if (tid == 0 && tile + RC_STAGES * (uint64_t)gridDim.x < n_tiles) rc_stage_tile(S, S.params, tile + RC_STAGES * (uint64_t)gridDim.x, full, buf);
        // ---- complement + reverse + validate + store ----------------------------------------------------
#pragma unroll
        for (int u = 0; u < RC_UNROLL; u++) {
            const uint32_t l = (uint32_t)u * THREADS + tid;
            if (l < cnt) {
                const uint64_t j = j0 + l;
                // x[u][0..3] = in[n-16j-16 .. n-16j) ; output ascending = input descending
                uint4 o;
                o.x = rc_map_word(x[u][3], coef, pair_base);
                o.y = rc_map_word(x[u][2], coef, pair_base);
                o.z = rc_map_word(x[u][1], coef, pair_base);
                o.w = rc_map_word(x[u][0], coef, pair_base);
                const uint32_t a = x[u][0] & x[u][1] & x[u][2] & x[u][3], r = x[u][0] | x[u][1] | x[u][2] | x[u][3];
                const uint32_t bad = ((o.x | o.y | o.z | o.w) & (RC_INVALID * 0x01010101u)) | ((a & p.and_need) ^ p.and_need) | (r & p.or_forbid);
                if (bad) {  // exact: the lowest failing input index of this granule
                    const uint64_t i0 = n - 16 * j - 16;
                    if ((unsigned long long)(p.key_base + i0) < *reinterpret_cast<volatile unsigned long long*>(p.err_key)) {
                        for (int t = 0; t < 16; t++)
                            if (__ldg(p.table + __ldg(p.dna + i0 + t)) == 0) {
                                atomicMin(p.err_key, (unsigned long long)(p.key_base + i0 + t));
                                break;
                            }
                    }
                }
#if QD_RC_TMA_STORE
                *reinterpret_cast<uint4*>(S.outb[it & 1] + 16 * l) = o;
#else
                stg_stream(p.out + 16 * j, o);
#endif
            }
        }
#if QD_RC_TMA_STORE
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            bulk_s2g(p.out + 16 * j0, S.outb[it & 1], cnt * 16);
            bulk_commit();
            bulk_wait_read_1();  // the other output buffer (next tile's) has been read out
        }
#endif
    }
#if QD_RC_TMA_STORE
    if (tid == 0) bulk_wait_all();
#endif
    // tail: out[16*full, n) = in[0, n mod 16) reversed
    if (blockIdx.x == 0 && tid < (int)(n & 15u)) {
        const uint64_t i = tid;
        const uint32_t v = __ldg(p.table + __ldg(p.dna + i));
        if (v == 0) atomicMin(p.err_key, (unsigned long long)(p.key_base + i));
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

// ------------------------------------------------------------------------------------------------
// Windows of a batch (DnaSequence::windows / ProteinSequence::windows over many sequences, and the DNA / amino-acid
// window sets of benches/all_windows.rs:78-102): out[s][i][j] = map(seq_s[i + j]), or with REVERSE the windows of the
// reverse complement, out[s][i][j] = map(seq_s[len - 1 - i - j]) -- taken straight from the forward strand, the reverse
// complement is never materialised.  `map` is a 256-byte table (identity when NULL; 0 = reject, for the strict DNA
// alphabets).  Sequence s starts at src + s * src_stride and has `len` bytes.
// A warp owns 64 consecutive windows of ONE sequence: the <= 64 + w - 1 source bytes go through the table into shared
// memory once (reversed if REVERSE), every lane re-aligns its window with funnel shifts and stores it into a staging
// row, and the staged rows leave with aligned 128-bit stores.
// ------------------------------------------------------------------------------------------------
struct WindowsBatchParams {
    const uint8_t* src;
    uint64_t src_stride;
    uint64_t n_seq;
    uint32_t len;  // bytes per sequence (<= 2^32 - 1)
    uint32_t w;    // 1 <= w <= 64
    int reverse;
    const uint8_t* map;  // 256-byte table or NULL
    uint8_t* out;        // n_seq * (len - w + 1) * w bytes
    uint64_t key_stride;  // error key of byte i of sequence s = s * key_stride + i
    unsigned long long* err_key;
    // Launched as a programmatic dependent (launch_pdl): 1 = src was written by the kernel before, wait for it.  A launch
    // with 0 here starts reading at once -- legal only when everything it reads was complete before the FIRST kernel of the
    // chain of dependents started.  Dependents behind this kernel are released after that wait, never before it.
    uint32_t dep_wait;
};
#ifndef QD_WIN_TMA_STORE
#define QD_WIN_TMA_STORE 1
#endif
constexpr int WIN_WPT = 64, WIN_WARPS = 8;
// windows per warp-tile: 64, or for a compile-time window length as many more groups of 64 (at most 4) as fit the stage,
// which is sized for 64 windows of 64 bytes -- the fixed cost of a tile is spread over more bytes of short windows
template <int W>
struct WinTile {
    static constexpr int GROUPS = W == 0 ? 1 : (64 / W < 1 ? 1 : (64 / W > 4 ? 4 : 64 / W));
    static constexpr int WPT = GROUPS * WIN_WPT;
};
struct WinWarpSmem {
    alignas(16) uint8_t srcb[4 + 4 * WIN_WPT + 64 + 24];
    alignas(16) uint8_t stage[WIN_WPT * 64 + 32];
};

// One window = srcb[4 + l .. 4 + l + w) -> stage row at byte D.  Words are formed ALIGNED TO THE DESTINATION: destination word m
// (bytes (D & ~3) + 4m ..) holds source bytes l - A + 4m .., A = D & 3; the middle words are whole 32-bit stores, only the
// first and the last word of a row can be partial.  With W and A compile-time constants everything unrolls.
template <int W, int A>
__device__ __forceinline__ void windows_copy_row(const uint8_t* srcb, uint8_t* stage, uint32_t l, uint32_t D, uint32_t w_rt) {
    const uint32_t w = W ? (uint32_t)W : w_rt;
    const uint32_t b0 = l + 4u - A;  // source offset of destination word 0, in srcb coordinates (4-byte front pad)
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(srcb) + (b0 >> 2);
    const uint32_t sh = (b0 & 3u) * 8u;
    uint32_t* dw = reinterpret_cast<uint32_t*>(stage + (D & ~3u));
    const uint32_t n_words = (A + w + 3u) >> 2, tail = (A + w) & 3u, whole_end = (A + w) >> 2;
    uint32_t prev = sw[0], m = 0;
    if (A != 0) {  // first word: bytes A .. min(4, A + w) - 1
        const uint32_t next = sw[1];
        const uint32_t word = __funnelshift_r(prev, next, sh);
        prev = next;
        uint8_t* db = reinterpret_cast<uint8_t*>(dw);
        const uint32_t hi = min(4u, A + w);
        if (A == 2 && hi == 4) {
            *reinterpret_cast<uint16_t*>(db + 2) = (uint16_t)(word >> 16);
        } else {
            if (A <= 1 && hi > 1) db[1] = (uint8_t)(word >> 8);
            if (A <= 2 && hi > 2) db[2] = (uint8_t)(word >> 16);
            if (hi > 3) db[3] = (uint8_t)(word >> 24);
        }
        m = 1;
    }
#pragma unroll
    for (uint32_t k = 0; k < 17; k++) {  // whole words (at most 64 / 4 + 1)
        if (m + k < whole_end) {
            const uint32_t next = sw[m + k + 1];
            dw[m + k] = __funnelshift_r(prev, next, sh);
            prev = next;
        }
    }
    m = max(m, whole_end);
    if (tail != 0 && (A == 0 || n_words > 1)) {  // last word: bytes 0 .. tail - 1
        const uint32_t word = __funnelshift_r(prev, sw[m + 1], sh);
        uint8_t* db = reinterpret_cast<uint8_t*>(dw + m);
        if (tail == 2) {
            *reinterpret_cast<uint16_t*>(db) = (uint16_t)word;
        } else {
            db[0] = (uint8_t)word;
            if (tail > 2) {
                db[1] = (uint8_t)(word >> 8);
                db[2] = (uint8_t)(word >> 16);
            }
        }
    }
}

// where tiles start so that their rows fill whole 128-byte lines of the output (see the kernel)
struct WinAlign {
    uint32_t g_shift, A, P, inv;
    __host__ __device__ constexpr explicit WinAlign(uint32_t w) : g_shift(0), A(0), P(0), inv(0) {
        uint32_t g = 0;
        while (g < 7 && ((w >> g) & 1u) == 0) g++;
        g_shift = g;
        A = (64u << g) < 128u ? (64u << g) : 128u;
        P = A >> g;
        const uint32_t x = w >> g;  // odd; its inverse modulo 2^12 by two Newton steps
        uint32_t i = x;
        i *= 2u - x * i;
        i *= 2u - x * i;
        inv = i;
    }
};

// MAP: bytes go through the 256-byte table p.map (DNA letters: upper case / complement + validation); plain copies skip it
template <int W, bool MAP = true>
__global__ void __launch_bounds__(32 * WIN_WARPS, W ? 6 : 4) windows_batch_kernel(const WindowsBatchParams p) {
    __shared__ WinWarpSmem SM[WIN_WARPS];
    __shared__ uint8_t map_s[256];
    for (int i = threadIdx.x; i < 256; i += 32 * WIN_WARPS) map_s[i] = MAP ? __ldg(p.map + i) : (uint8_t)i;
    if (p.dep_wait) pdl_wait();
    pdl_trigger();
    __syncthreads();
    const int lane = threadIdx.x & 31;
    WinWarpSmem& S = SM[threadIdx.x >> 5];
    constexpr uint32_t WPT = WinTile<W>::WPT;
    constexpr bool TMA_OUT = QD_WIN_TMA_STORE && WinTile<W>::GROUPS > 1;  // bulk-copy the rows out when a tile is large enough to pay for it
    const uint32_t w = W ? (uint32_t)W : p.w, nw = p.len - w + 1;
    // Tiles are cut where the OUTPUT crosses a 128-byte line: window g of the batch starts at byte g * w, so every tile but a
    // sequence's first starts at a batch-wide window index that is a multiple of P = A / gcd(w, A), A = 128 (64 for odd w:
    // P must divide the tile); the first tile of a sequence takes the 0 .. P - 1 windows in front of that.  Then a tile's
    // rows fill whole lines, nothing is written byte-wise but a sequence's first and last granule, and no two warps share
    // a line.  Measured, w = 42, 512 x ~100 000 nt: 0.78-0.81 of the copy peak when tiles start wherever a sequence starts
    // (or on 16-byte granules only), 0.88-0.89 when they start on lines.
    //   (out + g * w) % A == 0  <=>  g = -(phase / G) * inverse(w / G)  mod P,  G = gcd(w, A) = 1 << g_shift -- when G divides
    //   the phase of the buffer at all (else tiles start with the sequences)
    // (compile-time constants for the unrolled window lengths: the kernel has no register to spare -- at 48 instead of 40
    // registers, five instead of six CTAs per SM, w = 42 drops from 0.89 to 0.63 of the copy peak)
    const WinAlign al = W ? WinAlign(W) : WinAlign(w);
    const uint32_t g_shift = al.g_shift, A = al.A, P = al.P, inv = al.inv;
    const uint32_t tiles_per_seq = 1u + (nw + WPT - 1) / WPT;  // the short first tile (may be empty) + the aligned ones (the last may be empty)
    // (sequence, tile) of this warp's work, stepped by the number of warps in the grid; 32-bit: n_seq < 2^32 (checked by the host)
    const uint32_t warps_total = gridDim.x * WIN_WARPS, first = blockIdx.x * WIN_WARPS + (threadIdx.x >> 5);
    uint32_t s = first / tiles_per_seq, t = first % tiles_per_seq;
    const uint32_t step_s = warps_total / tiles_per_seq, step_t = warps_total % tiles_per_seq, n_seq = (uint32_t)p.n_seq;
    for (; s < n_seq;) {
        // windows in front of the sequence's first line boundary
        const uint32_t seq_phase = ((uint32_t)reinterpret_cast<uintptr_t>(p.out) + ((s * nw) & (A - 1u)) * w) & (A - 1u);
        const uint32_t lead = (seq_phase & ((1u << g_shift) - 1u)) ? 0u : ((((A - seq_phase) >> g_shift) * inv) & (P - 1u));
        const uint32_t i0 = t == 0 ? 0u : min(nw, lead + (t - 1u) * WPT), i1 = t == 0 ? min(nw, lead) : min(nw, lead + t * WPT);
        const uint32_t cnt = i1 - i0, span = cnt ? cnt + w - 1 : 0u;  // (an empty tile runs through the code below without touching anything)
        const uint8_t* seq = p.src + s * p.src_stride;
        // source bytes of the tile, mapped, in window order: srcb[4 + k] = map(seq[i0 + k]) or map(seq[len - 1 - i0 - k])
        // (a word-granular version of this loop was measured slower: 3.4 vs 2.8 ms on the all_windows bench leg)
        for (uint32_t k = lane; k < span; k += 32) {
            const uint32_t at = p.reverse ? p.len - 1 - i0 - k : i0 + k;
            uint32_t v = __ldg(seq + at);
            if (MAP) {  // (plain copies -- amino-acid windows, typed windows -- skip the table: one shared-memory access less per byte)
                v = map_s[v];
                if (v == 0) atomicMin(p.err_key, (unsigned long long)(s * p.key_stride + at));
            }
            S.srcb[4 + k] = (uint8_t)v;  // 4-byte front pad: a row's first destination word may start before its window
        }
        if (TMA_OUT && lane == 0) bulk_wait_read_0();  // the previous tile's rows have left the stage
        __syncwarp();
        const uint64_t out0 = ((uint64_t)s * nw + i0) * (uint64_t)w;
        uint8_t* gdst = p.out + out0;
        const uint32_t phase = (uint32_t)(reinterpret_cast<uintptr_t>(gdst) & 15u);  // stage mirrors the output's 16-byte phase
        // within every 64 windows lane handles windows 2 * lane and 2 * lane + 1: for even w all rows of one pass share their
        // alignment in the stage, so the switch below does not diverge
#pragma unroll 1
        for (uint32_t pass = 0; pass < WPT / 32; pass++) {
            const uint32_t l = 64u * (pass >> 1) + 2u * lane + (pass & 1u);
            if (l < cnt) {
                const uint32_t D = phase + l * w;
                switch (D & 3u) {
                    case 0: windows_copy_row<W, 0>(S.srcb, S.stage, l, D, w); break;
                    case 1: windows_copy_row<W, 1>(S.srcb, S.stage, l, D, w); break;
                    case 2: windows_copy_row<W, 2>(S.srcb, S.stage, l, D, w); break;
                    default: windows_copy_row<W, 3>(S.srcb, S.stage, l, D, w); break;
                }
            }
        }
        if (TMA_OUT) fence_proxy_async_smem();  // my rows, written through the generic proxy, are visible to the bulk copy
        __syncwarp();
        const uint32_t nbytes = cnt * w;
        const uint32_t head = min(nbytes, (16u - phase) & 15u), body = (nbytes - head) >> 4;
        if (lane < (int)head) gdst[lane] = S.stage[phase + lane];
        if (TMA_OUT) {
            // the aligned body leaves as ONE bulk copy shared -> global: no per-granule LDS.128 / STG.128 through the LSU pipe,
            // which is what bounds this kernel (measured: +8 % for 192-window tiles of 20 bytes, -13 % for 64-window tiles of 42)
            if (lane == 0 && body) {
                bulk_s2g(gdst + head, S.stage + phase + head, 16 * body);
                bulk_commit();
            }
        } else {
            for (uint32_t c = lane; c < body; c += 32) stg_stream(gdst + head + 16 * c, *reinterpret_cast<const uint4*>(S.stage + phase + head + 16 * c));
        }
        const uint32_t done = head + 16 * body;
        if (lane < (int)(nbytes - done)) gdst[done + lane] = S.stage[phase + done + lane];
        __syncwarp();
        s += step_s;
        t += step_t;
        if (t >= tiles_per_seq) { t -= tiles_per_seq; s++; }
    }
    if (TMA_OUT && lane == 0) bulk_wait_all();  // shared memory must outlive the last copy
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
// (src/canonical.rs:17-55, 71-118, 129-179 applied per window as benches/canonical.rs:29-34 does).
// ------------------------------------------------------------------------------------------------
struct CanonParams {
    const uint8_t* dna;      // n_seq sequences of seq_len nucleotides, back to back
    uint64_t n_seq;
    uint64_t seq_len;
    uint32_t w;
    int forward_only;
    uint8_t* out;            // n_seq * (seq_len - w + 1) windows, w bytes each, tightly packed, 16-byte aligned
    const uint8_t* decode;   // 256-byte table -> internal code (0..3 concrete); exact error search only
    const uint8_t* pair;     // PAIR_ENTRIES bytes: letter pair -> code0 | code1 << 2, | CKEY_PAIR_FLAG if either is not a concrete base
    uint32_t pair_coef;      // dp2a coefficients {1, k1} (byte entries)
    uint32_t and_need, or_forbid, filler;  // as in the frame / reverse-complement kernels
    uint32_t letters;        // output byte of codes 0..3 packed little-endian (ASCII "ATCG" or masks 1,2,4,8)
    unsigned long long* err_key;
};

constexpr int CANON_MAX_W = 64;

// ------------------------------------------------------------------------------------------------
// Canonical windows, tile kernel (src/canonical.rs:17-55, 71-118, 129-179 per window, as benches/canonical.rs:29-34 does).
// One kernel body, three outputs:
//   CTILE_ROWS        the canonical window's w bytes (letters or masks): qd_canonical_windows*
//   CTILE_KEY_PACKED  the canonical window as two BIT PLANES, lossless: PW = ceil(w / 32) words holding the low bit of each
//                     nucleotide's code (A, T, C, G = 0..3, the reference's order; nucleotide j at bit j % 32 of word j / 32),
//                     then PW words holding the high bits; unused bits zero            
//   CTILE_KEY_FP64    canon_fingerprint() of those words: 64 bits
//
// Bit planes make the relabelling two LOP3 per 32 nucleotides (new_lo = (lo & A) ^ (hi & B) ^ T0, new_hi likewise: every
// permutation of the four codes is affine over GF(2)^2) and the reversal one BREV per word.  What is left, finding the
// relabelling itself, is shared between windows: the order in which the four codes first occur reading FORWARD from
// position j is the same for every window that starts at j, and follows from the order at j + 1 by moving code[j] to
// the front; the order reading BACKWARD from j serves every window that ends at j.  (Codes that do not occur inside a
// window may sit anywhere in the order: they are never relabelled, so the scan need not stop at the window's end.)
// So after the decode a warp runs one pass over its tile's positions -- each lane takes 8 consecutive positions, starts
// from the order found directly in the 32 positions past its chunk, and steps through a 25 x 4 move-to-front table --
// leaving one byte per position and direction; a window then needs two byte loads and two table rows (its six
// coefficient masks each) instead of two searches.  Windows whose chunk start was not decided within 32 positions
// (low-complexity sequence) search their own w positions.
//
// Warps are independent (no block barrier after the tables are loaded): a warp owns `wpt` consecutive windows of the
// whole batch per iteration.  Keys leave straight from registers, consecutive lanes -> consecutive keys.  Rows are
// formed 4 bytes per PRMT from the RAW codes kept as nibbles (forward and reversed stream) through a 4-letter register
// that holds the winning relabelling, staged in shared memory in output order, and copied out with 128-bit stores.
// ------------------------------------------------------------------------------------------------
constexpr int CTILE_THREADS = 256;
constexpr int CTILE_WARPS = CTILE_THREADS / 32;
constexpr int CTILE_KEY_PACKED = 0, CTILE_KEY_FP64 = 1, CTILE_ROWS = 2;
constexpr uint32_t CKEY_PAIR_FLAG = 0x80;
constexpr uint32_t CKEY_RANKS = 25;  // 0 = not known, 1..24 = the orders of {0, 1, 2, 3} in lexicographic order
constexpr uint32_t CKEY_PAD = 32;    // positions of zero padding in front of the plane streams (backward look-behind)

struct CanonKeyTables {
    uint8_t rank_of[64];            // c0 | c1 << 2 | c2 << 4 (pairwise distinct) -> rank of the order (c0, c1, c2, the fourth); else 0
    uint8_t mtf[128];               // rank * 4 + code -> rank of the order with `code` moved to the front (rank 0 stays 0)
    uint32_t masks[CKEY_RANKS + 7][8];  // rank -> A, B, C, D, T0, T1 (0 or ~0), perm (2 bits per code: its place in the order), 0
};

// CTA-wide shared memory: the tables, the letter-pair image, the coefficient masks once more in a conflict-free layout,
// and the four output letters of every order
constexpr uint32_t CKEY_SHARED_TABLE_BYTES = ((uint32_t)(sizeof(CanonKeyTables) + PAIR_ENTRIES + 15) & ~15u) + 2 * CKEY_RANKS * 128 + 128;

// per-warp shared memory for a tile of `wpt` windows touching at most `pos_cap` positions (a multiple of 32)
struct CanonTileLayout {
    uint32_t plane_words, cw_bytes, state_bytes, nib_words, stage_bytes, total_bytes;
    __host__ __device__ CanonTileLayout(uint32_t pos_cap, uint32_t wpt, uint32_t w, bool rows) {
        plane_words = ((pos_cap + CKEY_PAD) / 32 + 4 + 3) & ~3u;  // every part a multiple of 16 bytes
        cw_bytes = (pos_cap / 4 + 15) & ~15u;  // 2 bits per position
        state_bytes = pos_cap;
        nib_words = rows ? ((pos_cap / 8 + 16 + 3) & ~3u) : 0u;  // 4 bits per position: one pad word in front, + what a window's last load may touch
        stage_bytes = rows ? ((64 * w + 15) & ~15u) : 0u;        // rows leave 64 at a time
        total_bytes = (8 * plane_words + cw_bytes + 2 * state_bytes + 8 * nib_words + stage_bytes + 15) & ~15u;
    }
};
// positions a warp-tile of `wpt` windows can touch: its windows, w - 1 more per sequence it reaches into, alignment slack
__host__ __device__ inline uint64_t ckey_span_bound(uint64_t wpt, uint64_t w, uint64_t nw) {
    const uint64_t seqs = (wpt - 1 + nw - 1) / nw + 1;
    return wpt + seqs * (w - 1) + 32;
}

__host__ __device__ __forceinline__ uint64_t canon_mix64(uint64_t z) {  // the splitmix64 finaliser: a bijection of 64 bits
    z ^= z >> 30;
    z *= 0xbf58476d1ce4e5b9ull;
    z ^= z >> 27;
    z *= 0x94d049bb133111ebull;
    z ^= z >> 31;
    return z;
}
// h = GOLDEN * w;  for k < PW:  h = mix64(h ^ (lo[k] | hi[k] << 32)).
// Every step is a bijection of h for a fixed pair and of the pair for a fixed h, so windows of w <= 32 never collide.
template <int PW>
__host__ __device__ __forceinline__ uint64_t canon_fingerprint(const uint32_t (&lo)[PW], const uint32_t (&hi)[PW], uint32_t w) {
    static_assert(PW == 1 || PW == 2, "at most 64 nucleotides");
    uint64_t h = 0x9e3779b97f4a7c15ull * w;
    h = canon_mix64(h ^ ((uint64_t)lo[0] | (uint64_t)hi[0] << 32));
    if (PW > 1) h = canon_mix64(h ^ ((uint64_t)lo[PW - 1] | (uint64_t)hi[PW - 1] << 32));
    return h;
}

// The order in which codes first occur in a window given as bit planes (V = the window's positions), as the index
// c0 | c1 << 2 | c2 << 4 of CanonKeyTables::rank_of.  A window with fewer than three distinct codes gets some valid order
// (the missing codes are never relabelled) unless STRICT3, which then returns 0xff.
template <int PW, bool STRICT3>
__device__ __forceinline__ uint32_t ckey_first_three(const uint32_t (&lo)[PW], const uint32_t (&hi)[PW], const uint32_t (&V)[PW]) {
    const uint32_t c0 = (lo[0] & 1u) | ((hi[0] & 1u) << 1);
    const uint32_t a0 = 0u - (lo[0] & 1u), b0 = 0u - (hi[0] & 1u);
    uint32_t ne1[PW];
#pragma unroll
    for (int k = 0; k < PW; k++) ne1[k] = ((lo[k] ^ a0) | (hi[k] ^ b0)) & V[k];  // positions whose code differs from c0
    uint32_t nsel = ne1[PW - 1], lsel = lo[PW - 1], hsel = hi[PW - 1];
    if (PW > 1 && ne1[0]) nsel = ne1[0], lsel = lo[0], hsel = hi[0];
    if (nsel == 0) return STRICT3 ? 0xffu : (c0 | ((c0 ^ 1u) << 2) | ((c0 ^ 2u) << 4));
    uint32_t bit = nsel & (0u - nsel);
    const uint32_t c1 = ((lsel & bit) ? 1u : 0u) | ((hsel & bit) ? 2u : 0u);
    const uint32_t a1 = 0u - (c1 & 1u), b1 = 0u - (c1 >> 1);
    uint32_t ne2[PW];
#pragma unroll
    for (int k = 0; k < PW; k++) ne2[k] = ne1[k] & ((lo[k] ^ a1) | (hi[k] ^ b1));  // ... and from c1
    nsel = ne2[PW - 1], lsel = lo[PW - 1], hsel = hi[PW - 1];
    if (PW > 1 && ne2[0]) nsel = ne2[0], lsel = lo[0], hsel = hi[0];
    if (nsel == 0) {
        if (STRICT3) return 0xffu;
        const uint32_t u = c0 ^ c1;
        return c0 | (c1 << 2) | ((c0 ^ (u == 1u ? 2u : 1u)) << 4);
    }
    bit = nsel & (0u - nsel);
    const uint32_t c2 = ((lsel & bit) ? 1u : 0u) | ((hsel & bit) ? 2u : 0u);
    return c0 | (c1 << 2) | (c2 << 4);
}

struct CanonKeyParams {
    CanonParams base;               // out = the rows or keys; pair = the ckey_pair image of the encoding
    const CanonKeyTables* tables;
    uint32_t wpt;                   // windows per warp-tile (a multiple of 64)
    uint32_t pos_cap;               // positions a warp-tile can touch, rounded up to 32 (sizes the shared memory)
};

// NW = ceil(w / 16) (the row emission works on 16-nucleotide units), W = the window length when it is a compile-time
// constant (42: benches/canonical.rs), 0 = p.w
template <int NW, int W, int MODE>
__global__ void __launch_bounds__(CTILE_THREADS) canonical_tile_kernel(const CanonKeyParams kp) {
    constexpr int PW = (NW + 1) / 2;  // plane words of a window
    constexpr bool ROWS = MODE == CTILE_ROWS;
    static_assert(W == 0 || (W + 15) / 16 == NW, "NW must be the 16-nucleotide unit count of W");
    extern __shared__ __align__(16) unsigned char ctile_smem[];
    const CanonParams& p = kp.base;
    const uint32_t w = W ? (uint32_t)W : p.w;
    const uint32_t WPT = kp.wpt;
    const CanonTileLayout lay(kp.pos_cap, WPT, w, ROWS);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // CTA-wide tables, then one private region per warp
    CanonKeyTables* const tabs = reinterpret_cast<CanonKeyTables*>(ctile_smem);
    uint8_t* const pair_s = ctile_smem + sizeof(CanonKeyTables);  // PAIR_ENTRIES bytes
    // the coefficient masks again, one 128-byte line per rank and part -- (A, B, C, D) and (T0, T1, T0 & V_last, T1 & V_last),
    // V_last = the positions of the window's last plane word -- replicated so that every lane of a quarter-warp reads its
    // own bank group whatever its rank: no conflicts
    uint32_t* const quad_s = reinterpret_cast<uint32_t*>(ctile_smem + CKEY_SHARED_TABLE_BYTES - 2 * CKEY_RANKS * 128 - 128);
    uint32_t* const duo_s = quad_s + CKEY_RANKS * 32;
    uint32_t* const letters_s = duo_s + CKEY_RANKS * 32;  // rank -> the output byte of raw code c at byte c
    unsigned char* const mine = ctile_smem + CKEY_SHARED_TABLE_BYTES + (size_t)warp * lay.total_bytes;
    uint32_t* const LO = reinterpret_cast<uint32_t*>(mine);
    uint32_t* const HI = LO + lay.plane_words;
    uint16_t* const CW = reinterpret_cast<uint16_t*>(HI + lay.plane_words);  // codes interleaved, 8 positions per half-word
    uint8_t* const stF = reinterpret_cast<uint8_t*>(CW) + lay.cw_bytes;      // order reading forward from each position
    uint8_t* const stB = stF + lay.state_bytes;                              // order reading backward from each position
    uint32_t* const NF = reinterpret_cast<uint32_t*>(stB + lay.state_bytes); // raw codes as nibbles, ascending positions
    uint32_t* const NR = NF + lay.nib_words;                                 // the same, descending
    uint8_t* const stage = reinterpret_cast<uint8_t*>(NR + lay.nib_words);   // the tile's rows in output order
    for (int i = tid; i < (int)(sizeof(CanonKeyTables) / 4); i += CTILE_THREADS)
        reinterpret_cast<uint32_t*>(tabs)[i] = __ldg(reinterpret_cast<const uint32_t*>(kp.tables) + i);
    for (int i = tid; i < (int)PAIR_ENTRIES; i += CTILE_THREADS) pair_s[i] = __ldg(p.pair + i);
    {
        const uint32_t left = w - 32u * (PW - 1), v_last = left >= 32u ? 0xffffffffu : ((1u << left) - 1u);
        for (int i = tid; i < (int)CKEY_RANKS * 32; i += CTILE_THREADS) {
            quad_s[i] = __ldg(&kp.tables->masks[i >> 5][i & 3]);
            duo_s[i] = __ldg(&kp.tables->masks[i >> 5][4 + (i & 1)]) & ((i & 2) ? v_last : 0xffffffffu);
        }
        if (tid < (int)CKEY_RANKS) {
            // letters[c] = p.letters[place of raw code c in the order]
            const uint32_t perm = __ldg(&kp.tables->masks[tid][6]);
            uint32_t t = (perm | (perm << 4)) & 0x0f0fu;
            t = (t | (t << 2)) & 0x3333u;
            letters_s[tid] = __byte_perm(p.letters, 0, t);
        }
    }
    __syncthreads();
    const uint64_t nw = p.seq_len - w + 1;  // windows per sequence
    const uint64_t total = p.n_seq * nw;    // windows in the batch
    const uint64_t n_tiles = (total + WPT - 1) / WPT;
    uint32_t V[PW];
#pragma unroll
    for (int k = 0; k < PW; k++) {
        const int left = (int)w - 32 * k;
        V[k] = left >= 32 ? 0xffffffffu : (left <= 0 ? 0u : ((1u << left) - 1u));
    }
    const uintptr_t dna_lo = reinterpret_cast<uintptr_t>(p.dna), dna_hi = dna_lo + p.n_seq * p.seq_len;
    const uint32_t pair_base = smem_u32(pair_s);
    const uint32_t mtf_base = smem_u32(tabs->mtf), rank_base = smem_u32(tabs->rank_of);
    const uint32_t quad_base = smem_u32(quad_s) + (lane & 7) * 16, duo_base = smem_u32(duo_s) + (lane & 7) * 16;
    const uint32_t lo_base = smem_u32(LO), hi_delta = 4u * lay.plane_words, stF_base = smem_u32(stF), stB_base = smem_u32(stB) + w - 1u;
    const uint32_t nf_base = smem_u32(NF), nr_base = smem_u32(NR), stage_base = smem_u32(stage), letters_base = smem_u32(letters_s);
    const bool long_seqs = nw >= WPT;  // a tile reaches into at most two sequences
    const uint32_t V32[1] = {0xffffffffu};
    const uint64_t warps_total = (uint64_t)gridDim.x * CTILE_WARPS;
    uint64_t tile = (uint64_t)blockIdx.x * CTILE_WARPS + warp;
    uint64_t s0 = (tile * WPT) / nw, i0 = (tile * WPT) % nw;
    const uint64_t step_s = (warps_total * WPT) / nw, step_i = (warps_total * WPT) % nw;
    for (; tile < n_tiles; tile += warps_total) {
        const uint64_t g0 = tile * WPT;
        const uint32_t cnt = (uint32_t)min((uint64_t)WPT, total - g0);
        uint64_t sl = s0, il = i0 + (cnt - 1);  // the tile's last window
        if (il >= nw) {
            if (nw >= WPT) { il -= nw; sl++; }
            else { sl += il / nw; il %= nw; }
        }
        const uint64_t flat_lo = s0 * p.seq_len + i0, flat_hi = sl * p.seq_len + il + w;
        const uintptr_t a0 = (dna_lo + flat_lo) & ~(uintptr_t)15;
        const uint32_t n_chunk = (uint32_t)((dna_lo + flat_hi - a0 + 7) >> 3);  // 8 positions each
        const uint32_t o_first = (uint32_t)(dna_lo + flat_lo - a0);             // position of the tile's first window
        const uint32_t l_seam = (uint32_t)min((uint64_t)WPT, nw - i0);          // the first window of the next sequence
        // ---- decode: item = 8 nucleotides -> one byte per plane, one half-word of interleaved codes (+ 2 nibble words) ----
        if (lane < 4) {
            reinterpret_cast<uint8_t*>(LO)[lane] = 0;  // the padding in front
            reinterpret_cast<uint8_t*>(HI)[lane] = 0;
        }
        if (ROWS && lane == 4) NF[0] = NR[0] = 0;
        for (uint32_t item = lane; item < n_chunk + 12; item += 32) {
            uint32_t plane_lo = 0, plane_hi = 0, cw = 0;
            if (item < n_chunk) {
                const uintptr_t a = a0 + 8 * (uintptr_t)item;
                uint32_t x0, x1;
                if (a >= dna_lo && a + 8 <= dna_hi) {
                    // bytes outside the tile's span are some other tile's: validated there too
                    const uint2 v = __ldg(reinterpret_cast<const uint2*>(a));
                    x0 = v.x;
                    x1 = v.y;
                } else {  // the chunks straddling the ends of the whole buffer: bytes outside it read as a valid filler
                    x0 = x1 = 0;
#pragma unroll 1
                    for (int t = 0; t < 8; t++) {
                        const uintptr_t b = a + t;
                        const uint32_t v = (b >= dna_lo && b < dna_hi) ? (uint32_t)*reinterpret_cast<const uint8_t*>(b) : (p.filler & 0xffu);
                        if (t < 4) x0 |= v << (8 * t);
                        else x1 |= v << (8 * (t - 4));
                    }
                }
                const uint32_t m0 = x0 & 0x1f1f1f1fu, m1 = x1 & 0x1f1f1f1fu;
                uint32_t e0 = lds_u8(dp2a_lo(p.pair_coef, m0, pair_base)), e1 = lds_u8(dp2a_hi(p.pair_coef, m0, pair_base));
                uint32_t e2 = lds_u8(dp2a_lo(p.pair_coef, m1, pair_base)), e3 = lds_u8(dp2a_hi(p.pair_coef, m1, pair_base));
                if (((e0 | e1 | e2 | e3) & CKEY_PAIR_FLAG) | (((x0 & x1) & p.and_need) ^ p.and_need) | ((x0 | x1) & p.or_forbid)) {
#pragma unroll 1
                    for (int t = 0; t < 8; t++) {  // exact: first byte of this chunk that is not a concrete base
                        const uintptr_t b = a + t;
                        if (b >= dna_lo && b < dna_hi && __ldg(p.decode + *reinterpret_cast<const uint8_t*>(b)) > 3u) {
                            atomicMin(p.err_key, (unsigned long long)(b - dna_lo));
                            break;
                        }
                    }
                    e0 &= 0xfu, e1 &= 0xfu, e2 &= 0xfu, e3 &= 0xfu;
                }
                cw = e0 | (e1 << 4) | (e2 << 8) | (e3 << 12);  // code of position t at bits 2t
                // even bits -> low plane byte, odd bits -> high plane byte (both halves of one word at once)
                uint32_t z = (cw & 0x5555u) | (((cw >> 1) & 0x5555u) << 16);
                z = (z | (z >> 1)) & 0x33333333u;
                z = (z | (z >> 2)) & 0x0f0f0f0fu;
                z = (z | (z >> 4)) & 0x00ff00ffu;
                plane_lo = z & 0xffu;
                plane_hi = z >> 16;
                if (ROWS) {
                    // the eight codes as nibbles, and once more in reversed order at the mirrored word
                    uint32_t n = (cw | (cw << 8)) & 0x00ff00ffu;
                    n = (n | (n << 4)) & 0x0f0f0f0fu;
                    n = (n | (n << 2)) & 0x33333333u;
                    NF[1 + item] = n;
                    const uint32_t b = __byte_perm(n, 0, 0x0123);
                    NR[n_chunk - item] = ((b >> 4) & 0x0f0f0f0fu) | ((b & 0x0f0f0f0fu) << 4);
                }
            } else if (ROWS) {
                NF[1 + item] = 0;  // the words past the end a window's last load may touch
                NR[1 + item] = 0;
            }
            // items past the tile's end: zeros (the look-ahead of the last chunks reads them)
            reinterpret_cast<uint8_t*>(LO)[CKEY_PAD / 8 + item] = (uint8_t)plane_lo;
            reinterpret_cast<uint8_t*>(HI)[CKEY_PAD / 8 + item] = (uint8_t)plane_hi;
            if (item < n_chunk) CW[item] = (uint16_t)cw;
        }
        __syncwarp();
        // ---- orders: lane takes chunk ch; positions 8 ch .. 8 ch + 7 --------------------------------------------
        for (uint32_t ch = lane; ch < n_chunk; ch += 32) {
            const uint32_t cw = CW[ch];
            {   // forward orders, from the back of the chunk to its front; start: the order found in positions 8 ch + 8 .. + 39
                const uint32_t q = CKEY_PAD + 8 * ch + 8, wi = q >> 5, sh = q & 31u;
                const uint32_t l[1] = {__funnelshift_r(LO[wi], LO[wi + 1], sh)}, h[1] = {__funnelshift_r(HI[wi], HI[wi + 1], sh)};
                const uint32_t idx = ckey_first_three<1, true>(l, h, V32);
                uint32_t r = idx == 0xffu ? 0u : lds_u8(rank_base + idx);
                uint32_t pk0 = 0, pk1 = 0;
#pragma unroll
                for (int t = 7; t >= 0; t--) {
                    r = lds_u8(mtf_base + r * 4u + ((cw >> (2 * t)) & 3u));
                    if (t >= 4) pk1 |= r << (8 * (t - 4));
                    else pk0 |= r << (8 * t);
                }
                *reinterpret_cast<uint2*>(stF + 8 * ch) = make_uint2(pk0, pk1);
            }
            {   // backward orders, front to back; start: the order found reading back from position 8 ch - 1
                const uint32_t q = CKEY_PAD + 8 * ch - 32, wi = q >> 5, sh = q & 31u;
                const uint32_t l[1] = {__brev(__funnelshift_r(LO[wi], LO[wi + 1], sh))}, h[1] = {__brev(__funnelshift_r(HI[wi], HI[wi + 1], sh))};
                const uint32_t idx = ckey_first_three<1, true>(l, h, V32);
                uint32_t r = idx == 0xffu ? 0u : lds_u8(rank_base + idx);
                uint32_t pk0 = 0, pk1 = 0;
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    r = lds_u8(mtf_base + r * 4u + ((cw >> (2 * t)) & 3u));
                    if (t >= 4) pk1 |= r << (8 * (t - 4));
                    else pk0 |= r << (8 * t);
                }
                *reinterpret_cast<uint2*>(stB + 8 * ch) = make_uint2(pk0, pk1);
            }
        }
        __syncwarp();
        // ---- my windows ------------------------------------------------------------------------------------------
        // keys: lane handles windows lane, lane + 32, ...: consecutive lanes write consecutive keys.
        // rows: within every 64 windows lane handles 2 lane and 2 lane + 1: for even w all rows of one pass share their
        //       alignment in the stage, and the rows of a pass are an odd number of words apart (no bank conflicts)
        const uint32_t n_it = ROWS ? 2u * ((cnt + 63u) >> 6) : (cnt + 31u) >> 5;
        uint32_t l = ROWS ? 2u * lane : lane;
#pragma unroll 1
        for (uint32_t it = 0; it < n_it; it++, l += ROWS ? ((it & 1u) ? 1u : 63u) : 32u) {
            if (l < cnt) {
            uint32_t o;  // position inside the streams: every sequence seam before the window skips w - 1 positions
            if (long_seqs) o = o_first + l + (l >= l_seam ? w - 1u : 0u);
            else o = o_first + l + (((uint32_t)i0 + l) / (uint32_t)nw) * (w - 1u);  // nw < WPT here: 32-bit division
            uint32_t rF = lds_u8(stF_base + o), rB = lds_u8(stB_base + o);
            uint32_t xl[PW], xh[PW], rl[PW], rh[PW];
            {
                const uint32_t q = CKEY_PAD + o, wa = lo_base + ((q >> 5) << 2), sh = q & 31u;
                uint32_t tl[PW + 1], th[PW + 1];
#pragma unroll
                for (int k = 0; k <= PW; k++) {
                    tl[k] = lds_u32(wa + 4 * k);
                    th[k] = lds_u32(wa + hi_delta + 4 * k);
                }
#pragma unroll
                for (int k = 0; k < PW; k++) {
                    xl[k] = __funnelshift_r(tl[k], tl[k + 1], sh) & V[k];
                    xh[k] = __funnelshift_r(th[k], th[k + 1], sh) & V[k];
                }
            }
            // the reversed window: position j <-> w - 1 - j
            if constexpr (PW == 1) {
                rl[0] = __brev(xl[0]) >> (32u - w);
                rh[0] = __brev(xh[0]) >> (32u - w);
            } else {
                const uint32_t sh = 64u - w;  // < 32: PW is the smallest that fits w
                const uint32_t bl0 = __brev(xl[PW - 1]), bl1 = __brev(xl[0]), bh0 = __brev(xh[PW - 1]), bh1 = __brev(xh[0]);
                rl[0] = __funnelshift_r(bl0, bl1, sh);
                rl[PW - 1] = bl1 >> sh;
                rh[0] = __funnelshift_r(bh0, bh1, sh);
                rh[PW - 1] = bh1 >> sh;
            }
            if (p.forward_only) rB = rF;
            if (__any_sync(__activemask(), (rF == 0u) | (rB == 0u))) {  // rare: the whole warp takes the detour together
                if (rF == 0) rF = lds_u8(rank_base + ckey_first_three<PW, false>(xl, xh, V));
                if (rB == 0) rB = lds_u8(rank_base + ckey_first_three<PW, false>(rl, rh, V));
            }
            const uint4 mf0 = lds_u128(quad_base + rF * 128u);
            const uint4 mf1 = lds_u128(duo_base + rF * 128u);  // T0, T1, and both restricted to the last word's positions
            uint32_t yl[PW], yh[PW];
#pragma unroll
            for (int k = 0; k < PW; k++) {
                yl[k] = (xl[k] & mf0.x) ^ (xh[k] & mf0.y) ^ (k == PW - 1 ? mf1.z : mf1.x);
                yh[k] = (xl[k] & mf0.z) ^ (xh[k] & mf0.w) ^ (k == PW - 1 ? mf1.w : mf1.y);
            }
            bool use_rev = false;
            if (!p.forward_only) {
                const uint4 mr0 = lds_u128(quad_base + rB * 128u);
                const uint4 mr1 = lds_u128(duo_base + rB * 128u);
                uint32_t zl[PW], zh[PW], lose[PW];
#pragma unroll
                for (int k = 0; k < PW; k++) {
                    zl[k] = (rl[k] & mr0.x) ^ (rh[k] & mr0.y) ^ (k == PW - 1 ? mr1.z : mr1.x);
                    zh[k] = (rl[k] & mr0.z) ^ (rh[k] & mr0.w) ^ (k == PW - 1 ? mr1.w : mr1.y);
                    // lexical order: the first differing position decides; there the smaller code wins, hi bit first.
                    // lose = positions that differ AND where the reversed window holds the larger code
                    const uint32_t d = (yl[k] ^ zl[k]) | (yh[k] ^ zh[k]);
                    const uint32_t hd = yh[k] ^ zh[k];
                    lose[k] = (d & (0u - d)) & ((hd & zh[k]) | (~hd & ~yl[k]));  // first differing position of this word, if lost there
                }
                if (PW == 1) {
                    use_rev = lose[0] == 0;  // (equal windows: either)
                } else {
                    const uint32_t d0 = (yl[0] ^ zl[0]) | (yh[0] ^ zh[0]);
                    use_rev = (d0 ? lose[0] : lose[PW - 1]) == 0;
                }
                if (!ROWS && use_rev) {
#pragma unroll
                    for (int k = 0; k < PW; k++) yl[k] = zl[k], yh[k] = zh[k];
                }
            }
            if constexpr (!ROWS) {
                const uint64_t g = g0 + l;
                if (MODE == CTILE_KEY_FP64) {
                    reinterpret_cast<uint64_t*>(p.out)[g] = canon_fingerprint<PW>(yl, yh, w);
                } else if (PW == 1) {
                    reinterpret_cast<uint2*>(p.out)[g] = make_uint2(yl[0], yh[0]);
                } else {
                    reinterpret_cast<uint4*>(p.out)[g] = make_uint4(yl[0], yl[PW - 1], yh[0], yh[PW - 1]);
                }
            } else {
                // the window's letters: T[raw code] = the letter of its place in the winning order
                const uint32_t T = lds_u32(letters_base + 4u * (use_rev ? rB : rF));
                // selectors: the window's raw codes as nibbles, from the forward or the reversed stream (one pad word in front).
                // A row of even length starts 4-byte aligned in the stage or 2 bytes past; in the second case the selectors
                // start 2 nucleotides early, so every word formed is an aligned stage word either way -- and the case is the
                // same for every row of a pass (rows 2 lane + pass), so the branch does not diverge.
                const uint32_t row = 2u * lane + (it & 1u);
                const uint32_t mis = (w & 1u) ? 0u : ((row * w) & 2u);
                const uint32_t on = 8u + (use_rev ? 8u * n_chunk - o - w : o) - mis;
                const uint32_t na = (use_rev ? nr_base : nf_base) + ((on >> 3) << 2), sh = (on & 7u) * 4u;
                constexpr int NN = 2 * NW;                              // nibble words of a window
                constexpr bool EXTRA = W == 0 || (W + 2 + 3) / 4 > 2 * NN;  // a misaligned row of 16 NW nucleotides reaches one word further
                constexpr int NL = 2 * NN + (EXTRA ? 1 : 0);
                uint32_t N[NN + 2];
#pragma unroll
                for (int k = 0; k <= NN + (EXTRA ? 1 : 0); k++) N[k] = lds_u32(na + 4 * k);
                const uint32_t dst = stage_base + row * w - mis;  // 4-byte aligned for even w
                uint32_t Lw[NL];  // the row's letters, four per PRMT
#pragma unroll
                for (int m = 0; m < NL; m += 2) {
                    const uint32_t sel = __funnelshift_r(N[m >> 1], N[(m >> 1) + 1], sh);  // nucleotides 4m .. 4m + 7 (- mis)
                    Lw[m] = __byte_perm(T, 0, sel);
                    if (m + 1 < NL) Lw[m + 1] = __byte_perm(T, 0, sel >> 16);
                }
                if ((w & 1u) == 0) {
                    if (mis == 0) {
#pragma unroll
                        for (int m = 0; m < NL; m++) {
                            const uint32_t j = 4 * m;  // row bytes j .. j + 3
                            if ((W && j + 4 <= (uint32_t)W) || j + 4 <= w) sts_u32(dst + j, Lw[m]);
                            else if (j < w) sts_u16(dst + j, Lw[m]);  // w % 4 == 2: the last two bytes
                        }
                    } else {
                        sts_u16(dst + 2, Lw[0] >> 16);  // row bytes 0, 1
#pragma unroll
                        for (int m = 1; m < NL; m++) {
                            const uint32_t j = 4 * m - 2;  // row bytes j .. j + 3
                            if ((W && j + 4 <= (uint32_t)W) || j + 4 <= w) sts_u32(dst + 4 * m, Lw[m]);
                            else if (j < w) sts_u16(dst + 4 * m, Lw[m]);
                        }
                    }
                } else {
#pragma unroll
                    for (int m = 0; m < NL; m++) {
                        const uint32_t j = 4 * m;
                        if (j < w) {
                            const uint32_t four = Lw[m], left = w - j;
                            sts_u8(dst + j, four);
                            if (left > 1) sts_u8(dst + j + 1, four >> 8);
                            if (left > 2) sts_u8(dst + j + 2, four >> 16);
                            if (left > 3) sts_u8(dst + j + 3, four >> 24);
                        }
                    }
                }
            }
            }  // l < cnt
            if constexpr (ROWS) {
                if ((it & 1u) != 0) {
                    // ---- copy-out of 64 rows: contiguous bytes at out + (g0 + first) * w (16-byte aligned: 64 w is a multiple of 16) ----
                    __syncwarp();
                    const uint32_t first = 64u * (it >> 1), nbytes = min(64u, cnt - first) * w, body = nbytes >> 4;
                    uint8_t* const gdst = p.out + (g0 + first) * w;
                    uint4* gp = reinterpret_cast<uint4*>(gdst) + lane;
                    uint32_t sa = stage_base + 16u * lane;
                    bool done = false;
                    if constexpr (W != 0) {
                        if (nbytes == 64u * (uint32_t)W) {  // a full group of a compile-time window length: 4 W granules, unrolled
                            constexpr uint32_t BODY = 4u * (uint32_t)W;
#pragma unroll
                            for (uint32_t k = 0; k < (BODY + 31u) / 32u; k++)
                                if (32u * k + 32u <= BODY || lane + 32u * k < BODY) stg_stream(gp + 32 * k, lds_u128(sa + 512u * k));
                            done = true;
                        }
                    }
                    if (!done) {
                        for (uint32_t c = lane; c < body; c += 32, gp += 32, sa += 512u) stg_stream(gp, lds_u128(sa));
                        for (uint32_t b = 16 * body + lane; b < nbytes; b += 32) gdst[b] = stage[b];
                    }
                    __syncwarp();
                }
            }
        }
        __syncwarp();
        s0 += step_s;
        i0 += step_i;
        if (i0 >= nw) { i0 -= nw; s0++; }
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
    uint64_t* desc;          // per window: its ambiguous positions, packed (fast path, see expansion_write_tiles_kernel)
    uint64_t row_cap;        // rows `out` can hold (fast path: rows past it are not written)
    uint8_t* out;
    unsigned long long* err_key;
};

constexpr uint32_t EXP_DESC_MAX = 6;  // ambiguous positions a 64-bit descriptor holds: 6 x (6-bit position, 4-bit mask) + count

// size_hint for a fresh iterator: prod(popcount) with checked arithmetic (src/expansions.rs:96-107)
__global__ void __launch_bounds__(256) expansion_count_kernel(const ExpandParams p) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n_win; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint8_t* win = p.windows + i * p.w;
        uint64_t size = 0, desc = 0;
        uint32_t n_amb = 0;
        bool overflow = false;
        for (uint32_t jj = 0; jj < p.w; jj++) {
            const uint32_t j = p.w - 1 - jj;  // last ambiguity first: it is the fastest odometer digit
            const uint32_t m = __ldg(p.to_mask + __ldg(win + j));
            if (m == 0) atomicMin(p.err_key, (unsigned long long)(i * p.w + j));
            const uint64_t states = (uint64_t)__popc(m);
            if (states > 1) {
                if (n_amb < EXP_DESC_MAX) desc |= (uint64_t)((j & 63u) | (m << 6)) << (10 * n_amb);
                n_amb++;
            }
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
        if (p.desc) p.desc[i] = desc | ((uint64_t)min(n_amb, 15u) << 60);
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

constexpr uint32_t PARSE_PAIR_SLOW = 0x40;  // letter-pair table of the parse (qd_tables.hpp): the letter is not a nucleotide of the alphabet
// ---- fast path (w <= 64, max_expansions < 128): warp tiles of 32 windows -------------------------------------------
constexpr int EXP_WARPS = 8;

// the tile's cnt * w source bytes into shared memory with aligned 128-bit loads; returns the offset of window 0 in `raw`
// (bytes before / after the buffer read as `filler`, a valid concrete base, and are never looked at)
// FOLD: clear bit 5 of every byte on the way (ASCII letters become upper case; no byte changes between valid and invalid)
template <bool FOLD = false>
__device__ __forceinline__ uint32_t expansion_stage_windows(const ExpandParams& p, uint64_t i0, uint32_t cnt, uint8_t* raw, int lane, uint32_t filler) {
    const uintptr_t lo = reinterpret_cast<uintptr_t>(p.windows), hi = lo + p.n_win * p.w;
    const uintptr_t src = lo + i0 * p.w, a0 = src & ~(uintptr_t)15;
    const uint32_t n_gran = (uint32_t)((src + (uintptr_t)cnt * p.w - a0 + 15) >> 4);
    for (uint32_t c = lane; c < n_gran; c += 32) {
        const uintptr_t a = a0 + 16 * (uintptr_t)c;
        uint4 v;
        if (a >= lo && a + 16 <= hi) {
            v = ldg_stream(reinterpret_cast<const void*>(a));
        } else {
            uint32_t wds[4] = {0, 0, 0, 0};
#pragma unroll 1
            for (int t = 0; t < 16; t++) {
                const uintptr_t b = a + t;
                wds[t >> 2] |= ((b >= lo && b < hi) ? (uint32_t)*reinterpret_cast<const uint8_t*>(b) : filler) << (8 * (t & 3));
            }
            v = make_uint4(wds[0], wds[1], wds[2], wds[3]);
        }
        if (FOLD) v = make_uint4(v.x & 0xdfdfdfdfu, v.y & 0xdfdfdfdfu, v.z & 0xdfdfdfdfu, v.w & 0xdfdfdfdfu);
        *reinterpret_cast<uint4*>(raw + 16 * c) = v;
    }
    return (uint32_t)(src - a0);
}

// size_hint and the ambiguity descriptor of ONE window whose bytes sit at raw[off ..] in shared memory
// (src/expansions.rs:96-107).  Four letters per step: masks from the letter-pair table (ASCII) or the bytes themselves
// (MASK); a word without an ambiguity code (m & (m - 1) == 0 in every byte) costs nothing more.
// W: the window length when it is a compile-time constant (42: benches/expansions.rs), 0 = p.w
// WIDE: the descriptor is two words of 16-bit entries -- position | (states - 1) << 6 | mask << 8, four in desc_out, two more and
// the count (bits 63:60) in desc_hi -- instead of one word of 10-bit entries (position | mask << 6) with the count on top
template <bool ASCII, int W, bool WIDE = false>
__device__ __forceinline__ void expansion_size_hint(const ExpandParams& p, const uint8_t* raw, uint32_t raw_base, uint32_t off, uint32_t pair_base,
                                                    uint64_t window, uint64_t& count, uint64_t& desc_out, uint64_t& desc_hi) {
    const uint32_t w = W ? (uint32_t)W : p.w, n_words = (w + 3) >> 2;
    const uint32_t pair_coef = 2u | ((2u * PAIR_K1_ASCII) << 16);  // 16-bit entries
    const uint32_t filler = ASCII ? 0x41u : 0x01u;
    const uint32_t sa = raw_base + (off & ~3u);
    const uint32_t sh = (off & 3u) * 8u;
    // phase A, the same instructions for every lane: a bit per position that holds an ambiguity code
    uint32_t amb_lo = 0, amb_hi = 0, bad = 0;
    uint32_t prev = lds_u32(sa);
#pragma unroll
    for (uint32_t k = 0; k < (W ? (uint32_t)(W + 3) / 4 : 16u); k++) {
        if (W == 0 && k >= n_words) break;
        const uint32_t next = lds_u32(sa + 4 * k + 4);
        uint32_t x = __funnelshift_r(prev, next, sh);
        prev = next;
        const uint32_t nb = min(4u, w - 4u * k);  // bytes of this word inside the window
        if (nb < 4) x = (x & ((1u << (8 * nb)) - 1u)) | ((filler * 0x01010101u) << (8 * nb));
        uint32_t m4;
        if (ASCII) {
            const uint32_t u = x & 0x1f1f1f1fu;
            m4 = lds_u16(dp2a_lo(pair_coef, u, pair_base)) | (lds_u16(dp2a_hi(pair_coef, u, pair_base)) << 16);
            bad |= (m4 & (PARSE_PAIR_SLOW * 0x01010101u)) | ((x & 0x40404040u) ^ 0x40404040u) | (x & 0x80808080u);
        } else {
            m4 = x;
            bad |= (x & 0xf0f0f0f0u) | ((x - 0x01010101u) & ~x & 0x80808080u);
        }
        // bytes with more than one bit -> one bit each -> a nibble (of a word with an invalid byte: garbage, the call fails anyway)
        uint32_t t4 = m4 & (m4 - 0x01010101u) & 0x0f0f0f0fu;
        t4 = (t4 | (t4 >> 1) | (t4 >> 2) | (t4 >> 3)) & 0x01010101u;
        const uint32_t nib = (t4 * 0x10204080u) >> 28;  // byte t -> bit t
        if (k < 8) amb_lo |= nib << (4 * k);
        else amb_hi |= nib << (4 * (k - 8));
    }
    if (bad) {  // exact: the first byte of the window that is not a nucleotide code (one branch per window, not per word)
        for (uint32_t j = 0; j < w; j++) {
            if (__ldg(p.to_mask + raw[off + j]) == 0) {
                atomicMin(p.err_key, (unsigned long long)(window * w + j));
                break;
            }
        }
        amb_lo = amb_hi = 0;  // (the call fails; the counts of this window do not matter)
    }
    // phase B: the few ambiguous positions, last first (the fastest odometer digit)
    uint64_t size = 1, desc = 0, hi = 0;
    uint32_t n_amb = 0;
    bool overflow = false;
    while (amb_lo | amb_hi) {
        const uint32_t j = amb_hi ? 63u - __clz(amb_hi) : 31u - __clz(amb_lo);
        if (j >= 32) amb_hi &= ~(1u << (j - 32));
        else amb_lo &= ~(1u << j);
        const uint32_t m = __ldg(p.to_mask + raw[off + j]);
        const uint32_t states = __popc(m);
        if (WIDE) {
            const uint64_t entry = (j & 63u) | ((states - 1u) << 6) | (m << 8);
            if (n_amb < 4) desc |= entry << (16 * n_amb);
            else if (n_amb < EXP_DESC_MAX) hi |= entry << (16 * (n_amb - 4));
        } else if (n_amb < EXP_DESC_MAX) {
            desc |= (uint64_t)((j & 63u) | (m << 6)) << (10 * n_amb);
        }
        n_amb++;
        overflow |= __umul64hi(size, (uint64_t)states) != 0;
        size *= states;
    }
    count = overflow ? 0xffffffffffffffffull : size;
    if (WIDE) {
        desc_out = desc;
        desc_hi = hi | ((uint64_t)min(n_amb, 15u) << 60);
    } else {
        desc_out = desc | ((uint64_t)min(n_amb, 15u) << 60);
    }
}

// size_hint, rows to emit and the ambiguity descriptor of every window: lane = window (two-pass form: the host entry
// point needs the row total before it can size the output; device callers with a buffer use expansion_fused_tiles_kernel)
template <bool ASCII, int W>
__global__ void __launch_bounds__(32 * EXP_WARPS) expansion_count_tiles_kernel(const ExpandParams p, const uint16_t* __restrict__ pair) {
    __shared__ __align__(16) uint8_t raw_s[EXP_WARPS][32 * 64 + 48];
    __shared__ __align__(16) uint16_t pair_s[ASCII ? PAIR_ENTRIES : 8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (ASCII) {
        for (int i = threadIdx.x; i < (int)PAIR_ENTRIES; i += 32 * EXP_WARPS) pair_s[i] = __ldg(pair + i);
        __syncthreads();
    }
    uint8_t* const raw = raw_s[warp];
    const uint32_t w = W ? (uint32_t)W : p.w;
    const uint32_t pair_base = smem_u32(pair_s), raw_base = smem_u32(raw);
    const uint32_t filler = ASCII ? 0x41u : 0x01u;
    const uint64_t n_tiles = (p.n_win + 31) / 32;
    for (uint64_t tile = (uint64_t)blockIdx.x * EXP_WARPS + warp; tile < n_tiles; tile += (uint64_t)gridDim.x * EXP_WARPS) {
        const uint64_t i0 = tile * 32;
        const uint32_t cnt = (uint32_t)min((uint64_t)32, p.n_win - i0);
        const uint32_t off0 = expansion_stage_windows(p, i0, cnt, raw, lane, filler);
        if (lane == 0) *reinterpret_cast<uint4*>(raw + ((off0 + cnt * w + 15) & ~15u)) = make_uint4(0, 0, 0, 0);  // the word past the last row
        __syncwarp();
        if (lane < (int)cnt) {
            uint64_t count, desc, unused;
            expansion_size_hint<ASCII, W>(p, raw, raw_base, off0 + lane * w, pair_base, i0 + lane, count, desc, unused);
            p.counts[i0 + lane] = count;
            p.emit[i0 + lane] = (count <= p.max_expansions) ? count : 0;
            p.desc[i0 + lane] = desc;
        }
        __syncwarp();
    }
}

// Rows of the fast path.  A warp owns 32 consecutive windows: their bytes, descriptors and first rows go to shared memory
// once; then 32 output rows at a time -- lane = row -- are copied from their window with whole 32-bit words aligned to the
// DESTINATION (upper-cased on the way: concrete positions are then final), patched at the ambiguous positions with the
// row's odometer digits, and drained with aligned 128-bit stores (head / tail bytes of the chunk individually).
// The window of a row comes from two ballots over the tile's non-empty windows instead of a search, the digits from
// 7-bit arithmetic (a row index inside a window is < 128).
// WMAX: bytes per window the buffers are sized for
template <int WMAX>
struct ExpWarpSmem {
    alignas(16) uint8_t win[(32 * WMAX + 48 + 15) & ~15];
    alignas(16) uint8_t stage[(32 * WMAX + 32 + 15) & ~15];
    uint64_t desc[32];
    uint32_t ne_first[32];  // first row (relative to the tile's first row) of the k-th window that emits rows
    uint32_t ne_win[32];    // ... and which window that is
};

#ifndef QD_EXP_SKEW
#define QD_EXP_SKEW 1  // 1: rows that start mid-word copy their second five words first (W = 42: no bank conflicts in the stage)
#endif

// the rows of one warp tile: S.win (window 0 at off0) and S.desc are filled; window `lane` emits `emit` rows, the first one
// `rel` rows after the tile's first row `row_lo`
// W: the window length when it is a compile-time constant with W % 4 == 2 (42: benches/expansions.rs), 0 = p.w
template <int W, int WMAX>
__device__ __forceinline__ void expansion_tile_rows(const ExpandParams& p, ExpWarpSmem<WMAX>& S, uint32_t letter_base, uint32_t off0, uint32_t cnt,
                                                    uint64_t row_lo, uint32_t rel, uint32_t emit, int lane) {
    static_assert(W == 0 || (W % 4 == 2 && W <= 64 && (W - 2) / 4 % 2 == 0), "the unrolled row copy is written for W = 2 mod 8");
    const uint32_t w = W ? (uint32_t)W : p.w;
    const uint32_t win_base = smem_u32(S.win), stage_base = smem_u32(S.stage);
    const uint32_t total = __shfl_sync(0xffffffffu, rel + emit, (int)cnt - 1);
    const uint32_t ne = __ballot_sync(0xffffffffu, emit != 0);
    const uint32_t m = __popc(ne);
    if (emit != 0) {
        const uint32_t k = __popc(ne & ((1u << lane) - 1u));
        S.ne_first[k] = rel;
        S.ne_win[k] = lane;
    }
    __syncwarp();
    const uint32_t my_ne_first = lane < (int)m ? S.ne_first[lane] : 0xffffffffu;
    uint8_t* gdst = p.out + row_lo * w;
    for (uint32_t r0 = 0; r0 < total; r0 += 32, gdst += 32 * w) {
        const uint32_t nrows = min(32u, total - r0);
        if (row_lo + r0 + nrows > p.row_cap) break;  // caller's buffer is too small: out_index[n_win] tells how much was needed
        const uint32_t phase = (uint32_t)(reinterpret_cast<uintptr_t>(gdst) & 15u);  // stage mirrors the output's 16-byte phase
        // the window of every row of the chunk: how many non-empty windows start at or before row r0, and a bit per
        // window that starts inside the chunk
        const uint32_t base_k = __popc(__ballot_sync(0xffffffffu, my_ne_first <= r0)) - 1u;
        const uint32_t starts = __reduce_or_sync(0xffffffffu, (my_ne_first > r0 && my_ne_first < r0 + 32u) ? 1u << (my_ne_first - r0) : 0u);
        if (lane < (int)nrows) {
            const uint32_t kidx = base_k + __popc(starts & (0xffffffffu >> (31 - lane)));
            const uint32_t j = S.ne_win[kidx];
            uint32_t k = r0 + lane - S.ne_first[kidx];  // the row's index inside its window: < 128
            const uint64_t d = S.desc[j];
            const uint32_t s0 = off0 + j * w, d0 = phase + lane * w;
            if (W != 0 && ((phase | off0) & 1u) == 0) {
                // rows and windows start at even addresses: the row is (W - 2) / 4 aligned destination words plus one
                // half-word, at its front (row starts mid-word) or at its back
                const uint32_t hb = d0 & 2u, hpos = hb ? 0u : (uint32_t)W - 2u;
                sts_u16(stage_base + d0 + hpos, lds_u16(win_base + s0 + hpos) & 0xdfdfu);
                const uint32_t sb = s0 + hb, sh = (sb & 3u) * 8u, sa = win_base + (sb & ~3u), da = stage_base + d0 + hb;
                constexpr int NQ = (W - 2) / 4;
                if (QD_EXP_SKEW) {
                    // rows W bytes apart: a word-aligned row and the mid-word row three rows on share their banks word for
                    // word; walking the halves of a row in opposite order keeps them apart
                    constexpr uint32_t HALF = 4u * (NQ / 2);
                    const uint32_t o1 = hb ? HALF : 0u, o2 = HALF - o1;
                    uint32_t prev = lds_u32(sa + o1);
#pragma unroll
                    for (int q = 0; q < NQ / 2; q++) {
                        const uint32_t next = lds_u32(sa + o1 + 4 * q + 4);
                        sts_u32(da + o1 + 4 * q, __funnelshift_r(prev, next, sh) & 0xdfdfdfdfu);
                        prev = next;
                    }
                    prev = lds_u32(sa + o2);
#pragma unroll
                    for (int q = 0; q < NQ / 2; q++) {
                        const uint32_t next = lds_u32(sa + o2 + 4 * q + 4);
                        sts_u32(da + o2 + 4 * q, __funnelshift_r(prev, next, sh) & 0xdfdfdfdfu);
                        prev = next;
                    }
                } else {
                    uint32_t prev = lds_u32(sa);
#pragma unroll
                    for (int q = 0; q < NQ; q++) {
                        const uint32_t next = lds_u32(sa + 4 * q + 4);
                        sts_u32(da + 4 * q, __funnelshift_r(prev, next, sh) & 0xdfdfdfdfu);
                        prev = next;
                    }
                }
            } else {
                // head bytes up to the first aligned destination word, whole words, tail bytes
                const uint32_t head = min(w, (4u - (d0 & 3u)) & 3u);
                for (uint32_t b = 0; b < head; b++) S.stage[d0 + b] = S.win[s0 + b] & 0xdfu;
                const uint32_t n_body = (w - head) >> 2, sb = s0 + head, sh = (sb & 3u) * 8u;
                const uint32_t* sw = reinterpret_cast<const uint32_t*>(S.win) + (sb >> 2);
                uint32_t* dw = reinterpret_cast<uint32_t*>(S.stage + d0 + head);
                uint32_t prev = sw[0];
#pragma unroll 1
                for (uint32_t q = 0; q < n_body; q++) {
                    const uint32_t next = sw[q + 1];
                    dw[q] = __funnelshift_r(prev, next, sh) & 0xdfdfdfdfu;
                    prev = next;
                }
                for (uint32_t b = head + 4u * n_body; b < w; b++) S.stage[d0 + b] = S.win[s0 + b] & 0xdfu;
            }
            const uint32_t n_amb = (uint32_t)(d >> 60);
#pragma unroll
            for (uint32_t e = 0; e < EXP_DESC_MAX; e++) {
                if (e < n_amb) {
                    const uint32_t f = (uint32_t)(d >> (10 * e)) & 0x3ffu, pos = f & 63u, mask = f >> 6;
                    const uint32_t states = __popc(mask);
                    const uint32_t quo = states == 2u ? k >> 1 : (states == 4u ? k >> 2 : (k * 171u) >> 9);  // k / 3 for k < 128
                    const uint32_t digit = k - quo * states;
                    k = quo;
                    sts_u8(stage_base + d0 + pos, lds_u8(letter_base + mask * 4u + digit));
                }
            }
        }
        __syncwarp();
        const uint32_t nbytes = nrows * w;
        const uint32_t hd = min(nbytes, (16u - phase) & 15u), body = (nbytes - hd) >> 4;
        if (lane < (int)hd) gdst[lane] = S.stage[phase + lane];
        {
            uint4* gp = reinterpret_cast<uint4*>(gdst + hd) + lane;
            uint32_t sa = stage_base + phase + hd + 16u * lane;
            for (uint32_t c = lane; c < body; c += 32, gp += 32, sa += 512u) stg_stream(gp, lds_u128(sa));
        }
        const uint32_t done = hd + 16 * body;
        if (lane < (int)(nbytes - done)) gdst[done + lane] = S.stage[phase + done + lane];
        __syncwarp();
    }
}

// mask * 4 + digit -> the letter of the digit-th base of the mask (possibilities() order A,T,C,G); needs a __syncthreads after
__device__ __forceinline__ void expansion_fill_letters(uint8_t* letter_of, uint32_t letters) {
    if (threadIdx.x < 64) {
        uint32_t mm = threadIdx.x >> 2;
        for (uint32_t t = 0; t < (threadIdx.x & 3u); t++) mm &= mm - 1u;
        letter_of[threadIdx.x] = mm ? (uint8_t)(letters >> (8 * (__ffs(mm) - 1))) : (uint8_t)0;
    }
}

// W: the window length when it is a compile-time constant (42: benches/expansions.rs), 0 = p.w
template <int W>
__global__ void __launch_bounds__(32 * EXP_WARPS) expansion_write_tiles_kernel(const ExpandParams p) {
    __shared__ ExpWarpSmem<64> SM[EXP_WARPS];
    __shared__ uint8_t letter_of[64];
    const int lane = threadIdx.x & 31;
    ExpWarpSmem<64>& S = SM[threadIdx.x >> 5];
    expansion_fill_letters(letter_of, p.letters);
    __syncthreads();
    const uint32_t w = W ? (uint32_t)W : p.w;
    const uint64_t n_tiles = (p.n_win + 31) / 32;
    for (uint64_t tile = (uint64_t)blockIdx.x * EXP_WARPS + (threadIdx.x >> 5); tile < n_tiles; tile += (uint64_t)gridDim.x * EXP_WARPS) {
        const uint64_t i0 = tile * 32;
        const uint32_t cnt = (uint32_t)min((uint64_t)32, p.n_win - i0);
        // geometry: first rows relative to the tile's first row (a tile emits fewer than 32 * 128 rows)
        const uint64_t my_first = p.out_index[min(i0 + lane, p.n_win)], my_next = p.out_index[min(i0 + lane + 1, p.n_win)];
        const uint64_t row_lo = __shfl_sync(0xffffffffu, my_first, 0);
        const uint32_t rel = (uint32_t)(my_first - row_lo), emit = lane < (int)cnt ? (uint32_t)(my_next - my_first) : 0u;
        S.desc[lane] = lane < (int)cnt ? p.desc[i0 + lane] : 0;
        const uint32_t off0 = expansion_stage_windows(p, i0, cnt, S.win, lane, 0x41u);
        __syncwarp();
        expansion_tile_rows<W, 64>(p, S, smem_u32(letter_of), off0, cnt, row_lo, rel, emit, lane);
        __syncwarp();
    }
    (void)w;
}

// ---- one pass: size_hint, row index and rows of a tile from ONE read of its windows ---------------------------------------
// The first row of a warp tile (32 windows) is the number of rows all tiles before it emit: a chained scan with decoupled
// look-back over one 64-bit status word per tile (bits 63:62 = 0 nothing yet, 1 the tile's own rows, 2 rows of all tiles
// up to and including it).  Warps are independent (no CTA barrier in the loop) and take tiles from a ticket counter, so
// a tile with a lower number was always started earlier, and the warp holding the lowest unpublished tile never waits:
// no deadlock, whatever part of the grid is resident.  A warp counts the windows of its NEXT tile and publishes that
// tile's rows before it looks back for the current one (two window buffers): by then -- a whole row phase later -- every
// tile before it has long been published, and all but the last few hundred carry their inclusive sum.
// `state` (ticket counter, then the status words) must be zero when the kernel starts.
// Measured on the way (2*10^7 windows of 42 nt, 1.3*10^8 rows; two-pass form 2.15 ms): tiles handed out statically, look-
// back right after the count: 2.99 ms (every warp walks back through hundreds of unfinished neighbours); CTA tiles of
// 8 x 32 windows with two __syncthreads per tile: 2.37 ms, a third of all warp time spent at the barriers; the same
// without barriers (flags in shared memory, count one tile ahead): 2.0 ms -- with tiles assigned statically the whole
// grid advances at the pace of the slowest warp of each generation.
#ifndef QD_EXP_FUSED_WARPS
#define QD_EXP_FUSED_WARPS 8
#endif
constexpr int EXP_FUSED_WARPS = QD_EXP_FUSED_WARPS;
constexpr uint64_t EXP_STATE_VALUE = (1ull << 62) - 1;
#ifndef QD_EXP_ELECT
#define QD_EXP_ELECT 1
#endif
#ifndef QD_EXP_LOOK
#define QD_EXP_LOOK 1  // 1 -> 1.74 ms, 2 -> 1.77, 4 -> 1.79, 8 -> 1.87 for 2e7 windows
#endif
constexpr int EXP_LOOK = QD_EXP_LOOK;  // status words per lane and look

__device__ __forceinline__ uint64_t ld_relaxed_gpu_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int WMAX>
struct ExpPipeWarpSmem {
    alignas(16) uint8_t win[2][(32 * WMAX + 48 + 15) & ~15];    // windows of the current and of the next tile
    alignas(16) uint8_t stage[2][(32 * WMAX + 32 + 15) & ~15];  // 32 rows on their way out, and the 32 before them
    ulonglong2 desc[2][32];  // ambiguous positions of each window (expansion_size_hint, WIDE form)
    uint16_t ne_first[32];   // first row (relative to the tile's first row) of the k-th window that emits rows
    uint8_t ne_win[32];      // ... and which window that is
};
template <int WMAX>
struct ExpPipeSmem {
    ExpPipeWarpSmem<WMAX> warp[EXP_FUSED_WARPS];
    alignas(16) uint16_t pair[PAIR_ENTRIES];
    uint8_t letter_of[64];
};

// rows the tiles before `tile` emit; every lane returns the sum.  A tile's own rows fit 32 bits (one REDUX per 32 status
// words); only the one inclusive word that ends the walk is 64 bits wide.
__device__ __forceinline__ uint64_t expansion_look_back(const uint64_t* status, uint64_t tile, int lane) {
    uint64_t acc = 0;
    uint32_t own = 0;  // own rows of the tiles walked over since `acc` was last topped up (a tile emits at most 4064 rows)
    int64_t idx = (int64_t)tile - 1 - lane;
    for (;;) {
        uint64_t v[EXP_LOOK];
#pragma unroll
        for (int g = 0; g < EXP_LOOK; g++) v[g] = idx - 32 * g >= 0 ? ld_relaxed_gpu_u64(status + idx - 32 * g) : 2ull << 62;  // in front of the first tile: nothing
#pragma unroll
        for (int g = 0; g < EXP_LOOK; g++) {
            while ((v[g] >> 62) == 0) v[g] = ld_relaxed_gpu_u64(status + idx - 32 * g);
        }
        bool found = false;
        uint64_t ends_with = 0;
#pragma unroll
        for (int g = 0; g < EXP_LOOK; g++) {
            if (!found) {
                const uint32_t done = __ballot_sync(0xffffffffu, (v[g] >> 62) == 2);
                const int first = done ? __ffs(done) - 1 : 32;
                own += __reduce_add_sync(0xffffffffu, lane < first ? (uint32_t)v[g] : 0u);
                if (done) {
                    found = true;
                    ends_with = __shfl_sync(0xffffffffu, v[g], first) & EXP_STATE_VALUE;
                }
            }
        }
        if (found) return acc + own + ends_with;
        if (own >= 0x80000000u) {  // (a tile emits at most 4064 rows: EXP_LOOK * 32 of them stay far below 2^31)
            acc += own;
            own = 0;
        }
        idx -= 32 * EXP_LOOK;
    }
}

// what a warp keeps of a tile between counting it and writing its rows
struct ExpTileRegs {
    uint64_t i0;
    uint32_t cnt, off0, emit, rel;
};

// W: the window length when it is a compile-time constant (42: benches/expansions.rs), 0 = p.w
template <bool ASCII, int W>
__global__ void __launch_bounds__(32 * EXP_FUSED_WARPS, W ? 4 : 3) expansion_fused_tiles_kernel(const ExpandParams p, const uint16_t* __restrict__ pair, uint64_t* state) {
    constexpr int WMAX = W ? W : 64;
    constexpr int FW = EXP_FUSED_WARPS;
    constexpr int WR = (W % 8 == 2) ? W : 0;  // the unrolled row copy exists for W = 2 mod 8
    extern __shared__ __align__(16) uint8_t exp_smem_raw[];
    ExpPipeSmem<WMAX>& SH = *reinterpret_cast<ExpPipeSmem<WMAX>*>(exp_smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    ExpPipeWarpSmem<WMAX>& S = SH.warp[warp];
    if (ASCII)
        for (int i = threadIdx.x; i < (int)PAIR_ENTRIES; i += 32 * FW) SH.pair[i] = __ldg(pair + i);
    expansion_fill_letters(SH.letter_of, p.letters);
    __syncthreads();
    const uint32_t w = W ? (uint32_t)W : p.w;
    const uint32_t pair_base = smem_u32(SH.pair), letter_base = smem_u32(SH.letter_of);
    const uint32_t filler = ASCII ? 0x41u : 0x01u;
    const uint64_t n_tiles = (p.n_win + 31) / 32;
    uint64_t* const out_index = const_cast<uint64_t*>(p.out_index);
    uint64_t* const status = state + 1;

    auto take_tile = [&]() -> uint64_t {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(reinterpret_cast<unsigned long long*>(state), 1ull);
        return __shfl_sync(0xffffffffu, t, 0);
    };
    // size_hint of the 32 windows of `tile` into buffer nb; publishes the rows the tile emits
    auto count_tile = [&](uint64_t tile, uint32_t nb, ExpTileRegs& T) {
        T.i0 = tile * 32;
        T.cnt = (uint32_t)min((uint64_t)32, p.n_win - T.i0);
        T.off0 = expansion_stage_windows<ASCII>(p, T.i0, T.cnt, S.win[nb], lane, filler);  // upper case from here on
        if (lane == 0) *reinterpret_cast<uint4*>(S.win[nb] + ((T.off0 + T.cnt * w + 15) & ~15u)) = make_uint4(0, 0, 0, 0);  // the word past the last row
        __syncwarp();
        uint64_t count = 0, desc = 0, desc_hi = 0;
        if (lane < (int)T.cnt) {
            expansion_size_hint<ASCII, W, true>(p, S.win[nb], smem_u32(S.win[nb]), T.off0 + lane * w, pair_base, T.i0 + lane, count, desc, desc_hi);
            p.counts[T.i0 + lane] = count;
        }
        T.emit = (lane < (int)T.cnt && count <= p.max_expansions) ? (uint32_t)count : 0u;
        S.desc[nb][lane] = make_ulonglong2(desc, desc_hi);
        uint32_t incl = T.emit;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        T.rel = incl - T.emit;
        if (lane == 31) st_relaxed_gpu_u64(status + tile, ((tile == 0 ? 2ull : 1ull) << 62) | incl);
    };

    // Per iteration: draw a ticket, count and PUBLISH that tile, then look back for the current tile and write its rows.
    // The order matters: a warp never sits in a look-back holding a tile it has drawn but not published.  (Drawing the
    // ticket and starting the window loads one step earlier, so that their latency passes behind the look-back, was
    // measured: publishes then queue behind look-backs that wait for other publishes -- 18.6 ms instead of 1.9.)
    ExpTileRegs cur{}, nxt{};
    uint32_t iter = 0, chunk_no = 0;
    uint64_t t_next = take_tile();
    if (t_next < n_tiles) count_tile(t_next, 0, nxt);
    for (; t_next < n_tiles; iter++) {
        const uint32_t b = iter & 1u;
        const uint64_t tile = t_next;
        cur = nxt;
        t_next = take_tile();
        if (t_next < n_tiles) count_tile(t_next, b ^ 1u, nxt);
        const uint32_t total = __shfl_sync(0xffffffffu, cur.rel + cur.emit, 31);
        uint64_t row_lo = 0;
        if (tile != 0) {
            row_lo = expansion_look_back(status, tile, lane);
            if (lane == 0) st_relaxed_gpu_u64(status + tile, (2ull << 62) | (row_lo + total));
        }
        if (lane < (int)cur.cnt) out_index[cur.i0 + lane] = row_lo + cur.rel;
        if (tile == n_tiles - 1 && lane == 0) out_index[p.n_win] = row_lo + total;
        if (!p.out) continue;

        // ---- rows of this warp tile: 32 at a time, lane = row ----
        // A chunk of 32 rows is a multiple of 16 bytes, so all chunks of a tile sit at the same 16-byte phase of the output;
        // the stage mirrors it.  The (phase) bytes a chunk leaves past its last whole granule are carried into the front
        // of the next chunk's stage, so only a tile's first and last granule are written byte-wise; everything else
        // leaves as one bulk copy per chunk (two stage buffers: the copy of chunk c reads while chunk c + 1 is built).
        const uint32_t win_base = smem_u32(S.win[b]);
        const uint32_t lim = (uint32_t)min((uint64_t)total, p.row_cap - min(p.row_cap, row_lo));  // rows the caller's buffer still holds
        const uint32_t rows_out = total <= lim ? total : (lim & ~31u);  // caller's buffer too small: whole chunks only (out_index[n_win] tells how much was needed)
        const uint32_t ne = __ballot_sync(0xffffffffu, cur.emit != 0);
        const uint32_t m = __popc(ne);
        if (cur.emit != 0) {
            const uint32_t k = __popc(ne & ((1u << lane) - 1u));
            S.ne_first[k] = (uint16_t)cur.rel;
            S.ne_win[k] = (uint8_t)lane;
        }
        __syncwarp();
        const uint32_t my_ne_first = lane < (int)m ? S.ne_first[lane] : 0xffffffffu;
        uint8_t* const gtile = p.out + row_lo * w;
        const uint32_t phase = (uint32_t)(reinterpret_cast<uintptr_t>(gtile) & 15u);
        uint8_t* galigned = gtile - phase;  // the granule the chunk's first row starts in
        const uint32_t d0 = phase + lane * w;
        const bool even = WR != 0 && ((phase | cur.off0) & 1u) == 0;
        for (uint32_t r0 = 0; r0 < rows_out; r0 += 32, galigned += 32 * w, chunk_no++) {
            const uint32_t nrows = min(32u, rows_out - r0);
            uint8_t* const stage = S.stage[chunk_no & 1u];
            const uint32_t stage_base = smem_u32(stage);
            if (lane == 0) {
                bulk_wait_read_1();  // the copy that last used this stage buffer has read it
                if (r0 != 0 && phase != 0)  // the previous chunk's last, partial granule
                    *reinterpret_cast<uint4*>(stage) = *reinterpret_cast<const uint4*>(S.stage[(chunk_no & 1u) ^ 1u] + 32 * w);
            }
            __syncwarp();
            // the window of every row of the chunk: how many non-empty windows start at or before row r0, and a bit per
            // window that starts inside the chunk
            const uint32_t base_k = __popc(__ballot_sync(0xffffffffu, my_ne_first <= r0)) - 1u;
            const uint32_t starts = __reduce_or_sync(0xffffffffu, (my_ne_first > r0 && my_ne_first < r0 + 32u) ? 1u << (my_ne_first - r0) : 0u);
            uint32_t n_amb = 0, k = 0;
            ulonglong2 dd = make_ulonglong2(0, 0);
            if (lane < (int)nrows) {
                const uint32_t kidx = base_k + __popc(starts & (0xffffffffu >> (31 - lane)));
                const uint32_t j = S.ne_win[kidx];
                k = r0 + lane - S.ne_first[kidx];  // the row's index inside its window: < 128
                dd = S.desc[b][j];
                n_amb = (uint32_t)(dd.y >> 60);
                const uint32_t s0 = cur.off0 + j * w;
                if (even) {
                    // rows and windows start at even addresses: the row is (W - 2) / 4 aligned destination words plus one
                    // half-word, at its front (row starts mid-word) or at its back
                    const uint32_t hb = d0 & 2u, hpos = hb ? 0u : (uint32_t)WR - 2u;
                    sts_u16(stage_base + d0 + hpos, lds_u16(win_base + s0 + hpos));
                    const uint32_t sb = s0 + hb, sh = (sb & 3u) * 8u, sa = win_base + (sb & ~3u), da = stage_base + d0 + hb;
                    constexpr int NQ = (WR - 2) / 4;
                    // rows W bytes apart: a word-aligned row and the mid-word row three rows on share their banks word for
                    // word; walking the halves of a row in opposite order keeps them apart
                    constexpr uint32_t HALF = 4u * (NQ / 2);
                    const uint32_t o1 = (QD_EXP_SKEW && hb) ? HALF : 0u, o2 = HALF - o1;
                    uint32_t prev = lds_u32(sa + o1);
#pragma unroll
                    for (int q = 0; q < NQ / 2; q++) {
                        const uint32_t next = lds_u32(sa + o1 + 4 * q + 4);
                        sts_u32(da + o1 + 4 * q, __funnelshift_r(prev, next, sh));
                        prev = next;
                    }
                    prev = lds_u32(sa + o2);
#pragma unroll
                    for (int q = 0; q < NQ / 2; q++) {
                        const uint32_t next = lds_u32(sa + o2 + 4 * q + 4);
                        sts_u32(da + o2 + 4 * q, __funnelshift_r(prev, next, sh));
                        prev = next;
                    }
                } else {
                    // head bytes up to the first aligned destination word, whole words, tail bytes
                    const uint8_t* const wsrc = S.win[b];
                    const uint32_t head = min(w, (4u - (d0 & 3u)) & 3u);
                    for (uint32_t q = 0; q < head; q++) stage[d0 + q] = wsrc[s0 + q];
                    const uint32_t n_body = (w - head) >> 2, sb = s0 + head, sh = (sb & 3u) * 8u;
                    const uint32_t* sw = reinterpret_cast<const uint32_t*>(wsrc) + (sb >> 2);
                    uint32_t* dw = reinterpret_cast<uint32_t*>(stage + d0 + head);
                    uint32_t prev = sw[0];
#pragma unroll 1
                    for (uint32_t q = 0; q < n_body; q++) {
                        const uint32_t next = sw[q + 1];
                        dw[q] = __funnelshift_r(prev, next, sh);
                        prev = next;
                    }
                    for (uint32_t q = head + 4u * n_body; q < w; q++) stage[d0 + q] = wsrc[s0 + q];
                }
            }
            // the row's odometer digits, fastest first: digit = k mod states, k /= states; states in {2, 3, 4} and
            // k < 128, so the quotient is one multiply: k / 2 = k * 128 >> 8, k / 3 = k * 86 >> 8, k / 4 = k * 64 >> 8
            const uint32_t patch_base = stage_base + d0;
            auto patch = [&](uint32_t f) {
                const uint32_t recip = (0x40568000u >> ((f >> 3) & 0x18u)) & 0xffu;
                const uint32_t quo = (k * recip) >> 8;
                const uint32_t digit = k - quo - quo * ((f >> 6) & 3u);
                k = quo;
                sts_u8(patch_base + (f & 63u), lds_u8(letter_base + ((f >> 6) & 0x3cu) + digit));
            };
            // (benches/expansions.rs puts two ambiguity codes into a window: the first two digits are straight-line code)
            if (n_amb > 0) patch((uint32_t)dd.x);
            if (n_amb > 1) patch((uint32_t)(dd.x >> 16));
            if (__any_sync(0xffffffffu, n_amb > 2)) {
                if (n_amb > 2) patch((uint32_t)(dd.x >> 32));
                if (n_amb > 3) patch((uint32_t)(dd.x >> 48));
                if (n_amb > 4) patch((uint32_t)dd.y);
                if (n_amb > 5) patch((uint32_t)(dd.y >> 16));
            }
            fence_proxy_async_smem();  // my rows, written through the generic proxy, are visible to the bulk copy
            __syncwarp();
            const bool last = r0 + 32u >= rows_out;
            const uint32_t start_off = r0 == 0 ? ((phase + 15u) & ~15u) : 0u;  // a tile's first granule may belong to the tile before
            const uint32_t end_off = phase + nrows * w, body_end = last ? (end_off & ~15u) : min(end_off & ~15u, 32u * w);
#if QD_EXP_ELECT
            {
                // warp-uniform operands through the uniform datapath (REDUX writes a uniform register)
                const uint64_t gaddr = reinterpret_cast<uint64_t>(galigned + start_off);
                const uint32_t g_lo = __reduce_max_sync(0xffffffffu, (uint32_t)gaddr), g_hi = __reduce_max_sync(0xffffffffu, (uint32_t)(gaddr >> 32));
                const uint32_t s_at = __reduce_max_sync(0xffffffffu, stage_base + start_off);
                const uint32_t n_at = __reduce_max_sync(0xffffffffu, body_end > start_off ? body_end - start_off : 0u);
                if (elect_one_lane()) {
                    if (n_at) asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(((uint64_t)g_hi << 32) | g_lo), "r"(s_at), "r"(n_at) : "memory");
                    bulk_commit();
                }
            }
#else
            if (lane == 0) {
                if (body_end > start_off) bulk_s2g(galigned + start_off, stage + start_off, body_end - start_off);
                bulk_commit();
            }
#endif
            if (r0 == 0) {  // (a branch the whole warp takes or skips)
                if (lane < (int)(min(start_off, end_off) - phase)) gtile[lane] = stage[phase + lane];
            }
            if (last) {
                const uint32_t tail0 = max(body_end, start_off);  // (a tile shorter than a granule has no body)
                if (tail0 + lane < end_off) galigned[tail0 + lane] = stage[tail0 + lane];
            }
        }
        __syncwarp();
    }
    if (lane == 0) bulk_wait_all();  // shared memory must outlive the last copy
}

// ------------------------------------------------------------------------------------------------
// Batch geometry pre-pass (variable-length batches): runs per sequence -> exclusive scan, and the
// sequence that holds the first run of every tile.
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;  // items per thread
constexpr int SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;

// OP 0: sum, OP 1: max (identity 0 for both)
template <int OP>
__device__ __forceinline__ uint64_t scan_comb(uint64_t a, uint64_t b) { return OP ? (a > b ? a : b) : a + b; }

template <int OP>
__device__ __forceinline__ uint64_t block_exclusive_scan_op(uint64_t v, uint64_t* total, uint64_t* warp_sums) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x = scan_comb<OP>(x, y);
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint64_t s = (lane < (int)(blockDim.x >> 5)) ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint64_t y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s = scan_comb<OP>(s, y);
        }
        warp_sums[lane] = s;  // inclusive
    }
    __syncthreads();
    const uint64_t warp_off = wid ? warp_sums[wid - 1] : 0;
    *total = warp_sums[(blockDim.x >> 5) - 1];
    uint64_t prev = __shfl_up_sync(0xffffffffu, x, 1);
    if (lane == 0) prev = 0;
    return scan_comb<OP>(warp_off, prev);
}
__device__ __forceinline__ uint64_t block_exclusive_scan(uint64_t v, uint64_t* total, uint64_t* warp_sums) {
    return block_exclusive_scan_op<0>(v, total, warp_sums);
}

// value(i): 0 = runs of sequence i from CSR offsets (ceil(len/48)); 1 = raw uint64 values
template <int MODE>
__device__ __forceinline__ uint64_t scan_value(const uint64_t* src, uint64_t i) {
    if (MODE == 0) return (src[i + 1] - src[i] + RUN_NT - 1) / RUN_NT;
    return src[i];
}

// a thread's SCAN_ITEMS consecutive values (0 past n): 128-bit loads for raw values in an aligned array
template <int MODE>
__device__ __forceinline__ void scan_load_items(const uint64_t* src, uint64_t n, uint64_t base, uint64_t (&v)[SCAN_ITEMS]) {
    static_assert(SCAN_ITEMS % 2 == 0, "pairs of values per 128-bit access");
    if (MODE == 1 && base + SCAN_ITEMS <= n && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k += 2) {
            const ulonglong2 x = *reinterpret_cast<const ulonglong2*>(src + base + k);
            v[k] = x.x;
            v[k + 1] = x.y;
        }
        return;
    }
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) v[k] = (base + k < n) ? scan_value<MODE>(src, base + k) : 0;
}

template <int MODE, int OP = 0>
// gate / gate_gen: the kernel returns at once unless *gate == gate_gen (nullptr: always runs); see ParseParams
// batch_stride: blockIdx.y selects one of several arrays of n values scanned in the same launches (array y at src / dst +
// y * batch_stride, its block sums at block_sums + y * blocks per array); 0 for a single array
__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums_kernel(const uint64_t* src, uint64_t n, uint64_t* block_sums,
                                                                       const unsigned long long* gate = nullptr, unsigned long long gate_gen = 0,
                                                                       uint64_t batch_stride = 0) {
    pdl_trigger();
    pdl_wait();  // (when launched as a dependent: src comes from the kernel before)
    if (gate != nullptr && *reinterpret_cast<const volatile unsigned long long*>(gate) != gate_gen) return;
    src += blockIdx.y * batch_stride;
    block_sums += (uint64_t)blockIdx.y * gridDim.x;
    __shared__ uint64_t warp_sums[32];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_BLOCK + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint64_t v[SCAN_ITEMS];
    scan_load_items<MODE>(src, n, base, v);
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) s = scan_comb<OP>(s, v[k]);
    uint64_t total;
    block_exclusive_scan_op<OP>(s, &total, warp_sums);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single CTA: exclusive scan of the block sums in place
template <int OP = 0>
__global__ void __launch_bounds__(SCAN_THREADS) scan_spine_kernel(uint64_t* block_sums, uint64_t n_blocks, const unsigned long long* gate = nullptr,
                                                                  unsigned long long gate_gen = 0) {
    pdl_trigger();
    pdl_wait();  // (a dependent launch: the block sums come from the kernel before)
    if (gate != nullptr && *reinterpret_cast<const volatile unsigned long long*>(gate) != gate_gen) return;
    block_sums += (uint64_t)blockIdx.x * n_blocks;  // one CTA per array of a batch
    __shared__ uint64_t warp_sums[32];
    __shared__ uint64_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint64_t base = 0; base < n_blocks; base += SCAN_BLOCK) {
        const uint64_t i0 = base + (uint64_t)threadIdx.x * SCAN_ITEMS;
        uint64_t v[SCAN_ITEMS], s = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            v[k] = i0 + k < n_blocks ? block_sums[i0 + k] : 0;
            s = scan_comb<OP>(s, v[k]);
        }
        uint64_t total;
        uint64_t ex = block_exclusive_scan_op<OP>(s, &total, warp_sums);
        const uint64_t carry = carry_s;
        ex = scan_comb<OP>(carry, ex);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            if (i0 + k < n_blocks) block_sums[i0 + k] = ex;
            ex = scan_comb<OP>(ex, v[k]);
        }
        __syncthreads();
        if (threadIdx.x == 0) carry_s = scan_comb<OP>(carry, total);
        __syncthreads();
    }
}

// dst[0..n] = exclusive scan (dst[n] = grand total)
template <int MODE, int OP = 0>
__global__ void __launch_bounds__(SCAN_THREADS) scan_finish_kernel(const uint64_t* src, uint64_t n, const uint64_t* block_sums,
                                                                   uint64_t* dst, const unsigned long long* gate = nullptr,
                                                                   unsigned long long gate_gen = 0, uint64_t batch_stride = 0) {
    pdl_trigger();
    pdl_wait();
    if (gate != nullptr && *reinterpret_cast<const volatile unsigned long long*>(gate) != gate_gen) return;
    src += blockIdx.y * batch_stride;
    dst += blockIdx.y * batch_stride;
    block_sums += (uint64_t)blockIdx.y * gridDim.x;
    __shared__ uint64_t warp_sums[32];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_BLOCK + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint64_t v[SCAN_ITEMS];
    scan_load_items<MODE>(src, n, base, v);
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) s = scan_comb<OP>(s, v[k]);
    uint64_t total;
    uint64_t ex = scan_comb<OP>(block_exclusive_scan_op<OP>(s, &total, warp_sums), block_sums[blockIdx.x]);
    if (base + SCAN_ITEMS <= n && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {  // whole group: 128-bit stores
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k += 2) {
            ulonglong2 o;
            o.x = ex;
            ex = scan_comb<OP>(ex, v[k]);
            o.y = ex;
            ex = scan_comb<OP>(ex, v[k + 1]);
            *reinterpret_cast<ulonglong2*>(dst + base + k) = o;
        }
        if (base + SCAN_ITEMS == n) dst[n] = ex;
        return;
    }
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n) dst[base + k] = ex;
        ex = scan_comb<OP>(ex, v[k]);
        if (base + k + 1 == n) dst[n] = ex;
    }
    if (n == 0 && blockIdx.x == 0 && threadIdx.x == 0) dst[0] = 0;
}

// ------------------------------------------------------------------------------------------------
// The maps a variable-length batch needs before the frame kernel, written by the LAST pass of the scan of its run
// counts (one kernel instead of scan-finish + two map kernels + a copy of the offsets + two fills):
//   run_prefix[r]           runs before sequence r (exclusive scan of ceil(len / 48)), [n_seq] = the total
//   off_copy[r]             the offsets once more, next to the prefix (the frame kernel bulk-copies windows of both)
//   both padded with ~0 up to arr_entries
//   tiles[t]                {flat offset of run t * tile_runs, the sequence that holds it}
//   warp_first[t * wq_stride + j]   the sequence that holds run t * tile_runs + 32 j
// Every 32nd run g lies in exactly one sequence (a <= g < b, b > a), so the sequence's own thread writes the entry: no
// search.  A sequence of up to CSR_THREAD_RUNS runs is mapped by its own thread; longer ones go on two lists and are mapped
// by csr_listed_kernel: up to CSR_WARP_RUNS runs by one warp each, a chromosome among reads (163 k entries) by one CTA.
// ------------------------------------------------------------------------------------------------
constexpr uint64_t CSR_THREAD_RUNS = 512, CSR_WARP_RUNS = 65536;

struct CsrMapParams {
    const uint64_t* offsets;  // n_seq + 1, absolute
    uint64_t n_seq;
    uint64_t flat_base;       // offsets[0]: TileInfo::flat counts from here
    uint64_t* run_prefix;
    uint64_t* off_copy;
    uint64_t arr_entries;     // entries of run_prefix / off_copy, pads included
    TileInfo* tiles;
    uint64_t n_tiles_cap;
    uint32_t tile_shift;      // tile_runs = 1 << tile_shift (a multiple of 32)
    uint32_t wq_stride;
    uint32_t* warp_first;
    uint64_t total_nt;          // nucleotides in the buffer: every sequence must lie inside [flat_base, flat_base + total_nt]
    unsigned long long* err_key;  // CSR_BAD_OFFSETS | sequence index for offsets that do not
    uint32_t* mid_list;         // sequences of more than CSR_THREAD_RUNS runs
    uint32_t* long_list;        // sequences of more than CSR_WARP_RUNS runs
    unsigned int* list_counts;  // entries of the two lists; zeroed before the launch
};

// Offsets come from the caller (device memory: the host cannot look at them): a sequence whose offsets decrease or leave the
// buffer counts as empty -- the frame kernel never touches it -- and is reported through the error key, above every
// nucleotide position (those keep precedence in the atomicMin) and below "no error".
constexpr unsigned long long CSR_BAD_OFFSETS = 1ull << 63;
__device__ __forceinline__ uint64_t csr_runs_checked(const CsrMapParams& p, uint64_t o0, uint64_t o1, bool& bad) {
    bad = o1 < o0 || o0 - p.flat_base > p.total_nt || o1 - p.flat_base > p.total_nt;
    return bad ? 0 : (o1 - o0 + RUN_NT - 1) / RUN_NT;
}

// first pass of the scan: runs per block of SCAN_BLOCK sequences
__global__ void __launch_bounds__(SCAN_THREADS) csr_block_sums_kernel(const CsrMapParams p, uint64_t* __restrict__ block_sums) {
    pdl_trigger();
    __shared__ uint64_t warp_sums[32];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_BLOCK + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint64_t o[SCAN_ITEMS + 1], s = 0;
#pragma unroll
    for (int k = 0; k <= SCAN_ITEMS; k++) o[k] = base + k <= p.n_seq ? p.offsets[base + k] : 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        bool bad;
        if (base + k < p.n_seq) s += csr_runs_checked(p, o[k], o[k + 1], bad);
    }
    uint64_t total;
    block_exclusive_scan(s, &total, warp_sums);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// the entry of run g (a multiple of 32) of sequence r = runs [a, ...), first nucleotide at o0 past flat_base
__device__ __forceinline__ void csr_map_entry(const CsrMapParams& p, uint64_t g, uint64_t a, uint64_t o0, uint32_t r) {
    const uint64_t t = g >> p.tile_shift;
    if (t >= p.n_tiles_cap) return;
    const uint32_t j = (uint32_t)(g - (t << p.tile_shift)) >> 5;
    p.warp_first[t * p.wq_stride + j] = r;
    if (j == 0) {
        TileInfo ti;
        ti.flat = o0 + (g - a) * RUN_NT;
        ti.first_seq = r;
        ti.pad_ = 0;
        p.tiles[t] = ti;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) csr_finish_kernel(const CsrMapParams p, const uint64_t* __restrict__ block_sums) {
    pdl_trigger();
    pdl_wait();
    __shared__ uint64_t warp_sums[32];
    const uint64_t n = p.n_seq;
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_BLOCK + (uint64_t)threadIdx.x * SCAN_ITEMS;
    const int lane = threadIdx.x & 31;
    uint64_t o[SCAN_ITEMS + 1], v[SCAN_ITEMS], a[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k <= SCAN_ITEMS; k++) o[k] = base + k <= n ? p.offsets[base + k] : 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        bool bad = false;
        v[k] = base + k < n ? csr_runs_checked(p, o[k], o[k + 1], bad) : 0;
        if (bad) atomicMin(p.err_key, CSR_BAD_OFFSETS | (base + k));
        s += v[k];
    }
    uint64_t total;
    uint64_t ex = block_exclusive_scan(s, &total, warp_sums) + block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        a[k] = ex;
        ex += v[k];
    }
    if (base + SCAN_ITEMS <= n) {  // whole group: 128-bit stores (both arrays 16-byte aligned, base a multiple of 8)
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k += 2) {
            *reinterpret_cast<ulonglong2*>(p.run_prefix + base + k) = make_ulonglong2(a[k], a[k + 1]);
            *reinterpret_cast<ulonglong2*>(p.off_copy + base + k) = make_ulonglong2(o[k], o[k + 1]);
        }
        if (base + SCAN_ITEMS == n) {
            p.run_prefix[n] = ex;
            p.off_copy[n] = o[SCAN_ITEMS];
        }
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            if (base + k < n) {
                p.run_prefix[base + k] = a[k];
                p.off_copy[base + k] = o[k];
                if (base + k + 1 == n) {
                    p.run_prefix[n] = a[k] + v[k];
                    p.off_copy[n] = o[k + 1];
                }
            }
        }
    }
    if (blockIdx.x == 0) {
        for (uint64_t i = n + 1 + threadIdx.x; i < p.arr_entries; i += SCAN_THREADS) {
            p.run_prefix[i] = ~0ull;
            p.off_copy[i] = ~0ull;
        }
    }
    // the maps
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        const uint64_t ak = a[k], vk = v[k], o0 = o[k] - p.flat_base;
        const uint32_t r = (uint32_t)(base + k);
        if (vk <= CSR_THREAD_RUNS) {
            for (uint64_t g = (ak + 31) & ~(uint64_t)31; g < ak + vk; g += 32) csr_map_entry(p, g, ak, o0, r);
        } else if (vk > CSR_WARP_RUNS) {
            p.long_list[atomicAdd(p.list_counts + 1, 1u)] = r;
        }
        // the middle list can take most of a batch (long reads): one atomic per warp
        const bool mid = vk > CSR_THREAD_RUNS && vk <= CSR_WARP_RUNS;
        const uint32_t mm = __ballot_sync(0xffffffffu, mid);
        if (mm) {
            uint32_t at = 0;
            if (lane == __ffs(mm) - 1) at = atomicAdd(p.list_counts, (unsigned int)__popc(mm));
            at = __shfl_sync(0xffffffffu, at, __ffs(mm) - 1);
            if (mid) p.mid_list[at + __popc(mm & ((1u << lane) - 1u))] = r;
        }
    }
}

// the listed sequences: one warp each for the middle ones, one CTA each for the long ones
__global__ void __launch_bounds__(256) csr_listed_kernel(const CsrMapParams p) {
    pdl_trigger();
    pdl_wait();
    const unsigned int n_mid = p.list_counts[0], n_long = p.list_counts[1];
    const int lane = threadIdx.x & 31;
    for (unsigned int i = blockIdx.x * 8u + (threadIdx.x >> 5); i < n_mid; i += gridDim.x * 8u) {
        const uint32_t r = p.mid_list[i];
        const uint64_t a = p.run_prefix[r], b = p.run_prefix[r + 1], o0 = p.off_copy[r] - p.flat_base;
        for (uint64_t g = ((a + 31) & ~(uint64_t)31) + 32u * lane; g < b; g += 32u * 32u) csr_map_entry(p, g, a, o0, r);
    }
    for (unsigned int i = blockIdx.x; i < n_long; i += gridDim.x) {
        const uint32_t r = p.long_list[i];
        const uint64_t a = p.run_prefix[r], b = p.run_prefix[r + 1], o0 = p.off_copy[r] - p.flat_base;
        for (uint64_t g = ((a + 31) & ~(uint64_t)31) + 32u * threadIdx.x; g < b; g += 32u * 256u) csr_map_entry(p, g, a, o0, r);
    }
}

// ------------------------------------------------------------------------------------------------
// ASCII -> typed parse (src/rust_api.rs:310-322): drop ' ' and '\t', validate every other byte, emit one mask per
// kept byte.  Stream compaction in two passes: kept-byte counts per 4 KiB block (+ an optimistic in-place store of
// every clean granule) -> device scan -> compacted, coalesced write of the blocks the guess was wrong for.
// ------------------------------------------------------------------------------------------------
constexpr int PARSE_THREADS = 256;
constexpr int PARSE_BLOCK_BYTES = PARSE_THREADS * 16;  // one aligned 16-byte granule per thread
constexpr uint32_t PARSE_SKIP = 0x80;                  // table value of ' ' and '\t' (0 = invalid, else the mask)

struct ParseParams {
    const uint8_t* in;
    uint64_t n;
    const uint8_t* table;        // 256 bytes: mask | PARSE_SKIP | 0
    const uint16_t* pair;        // letter pair -> mask0 | mask1 << 8 (PARSE_PAIR_SLOW in a byte: not a plain nucleotide)
    uint32_t pair_coef;
    uint64_t* counts;            // kept bytes per block (pass 1 out)
    const uint64_t* offsets;     // exclusive scan of counts, [n_blocks + 1] (pass 2 in)
    uint8_t* out;
    unsigned long long* err_key;
    // Input without whitespace whose masks land where the letters were is finished by ONE optimistic pass
    // (parse_pass1_kernel); that pass writes `gate_gen` to *gate when it drops a byte, and only then do the
    // scan and the compacting pass behind it run (gate == nullptr: they always run).
    unsigned long long* gate;
    unsigned long long gate_gen;
};
__device__ __forceinline__ bool parse_gate_closed(const unsigned long long* gate, unsigned long long gen) {
    return gate != nullptr && *reinterpret_cast<const volatile unsigned long long*>(gate) != gen;
}

// granule `g` (16 bytes at the 16-byte-aligned address a) of the input as table values; bytes outside [in, in+n) skip
__device__ __forceinline__ void parse_load(const ParseParams& p, const uint8_t* tab_s, uint64_t g, uint8_t (&v)[16], uint32_t& kept,
                                           int& first_bad) {
    const uintptr_t lo = reinterpret_cast<uintptr_t>(p.in), hi = lo + p.n;
    const uintptr_t a = (lo & ~(uintptr_t)15) + 16 * g;
    uint32_t w[4];
    if (a >= lo && a + 16 <= hi) {
        const uint4 x = ldg_stream(reinterpret_cast<const void*>(a));
        w[0] = x.x; w[1] = x.y; w[2] = x.z; w[3] = x.w;
    } else {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t x = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uintptr_t b = a + 4 * q + k;
                x |= ((b >= lo && b < hi) ? (uint32_t)*reinterpret_cast<const uint8_t*>(b) : 0x20u) << (8 * k);  // ' ' = skipped
            }
            w[q] = x;
        }
    }
    kept = 0;
    first_bad = -1;
#pragma unroll
    for (int k = 15; k >= 0; k--) {
        const uint32_t t = tab_s[(w[k >> 2] >> (8 * (k & 3))) & 0xffu];
        v[k] = (uint8_t)t;
        if (t == 0) first_bad = k;
        kept += (t != 0 && t != PARSE_SKIP) ? 1u : 0u;
    }
}

// Fast path for a granule that lies wholly inside the input: two letters per lookup (the frame kernel's dp2a pair
// index); returns false when any byte is whitespace, invalid or outside the letter range (then the byte path decides)
__device__ __forceinline__ bool parse_granule_fast(const ParseParams& p, uint32_t pair_base, uint64_t g, uint4& masks) {
    const uintptr_t lo = reinterpret_cast<uintptr_t>(p.in), hi = lo + p.n;
    const uintptr_t a = (lo & ~(uintptr_t)15) + 16 * g;
    if (a < lo || a + 16 > hi) return false;
    const uint4 x = ldg_stream(reinterpret_cast<const void*>(a));
    const uint32_t w[4] = {x.x, x.y, x.z, x.w};
    uint32_t m[4], flags = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint32_t u = w[q] & 0x1f1f1f1fu;
        const uint32_t e_lo = lds_u16(dp2a_lo(p.pair_coef, u, pair_base)), e_hi = lds_u16(dp2a_hi(p.pair_coef, u, pair_base));
        m[q] = e_hi * 65536u + e_lo;
        flags |= m[q];
    }
    const uint32_t a_and = w[0] & w[1] & w[2] & w[3], a_or = w[0] | w[1] | w[2] | w[3];
    if ((flags & (PARSE_PAIR_SLOW * 0x01010101u)) | ((a_and & 0x40404040u) ^ 0x40404040u) | (a_or & 0x80808080u)) return false;
    masks = make_uint4(m[0], m[1], m[2], m[3]);
    return true;
}

// Pass 1, a streaming kernel without barriers: every granule that is 16 plain nucleotide letters is translated and --
// STORE: input and output share their 16-byte phase -- stored where it lands if nothing before it is dropped; granules
// cut by the ends of the input go byte-wise.  Dropped bytes (blank, tab, invalid) are added to their 4-KiB block's
// entry of `counts` (all zero when the kernel starts, and all zero again when the call is over), the first invalid byte
// goes to the error key, and the gate opens for the scan and the compacting pass.  Input without whitespace is
// finished here: one launch that does work, one read of the input.
#ifndef QD_PARSE_UNROLL
#define QD_PARSE_UNROLL 1  // granules per thread in flight: 1 -> 0.0933 ms, 4 -> 0.0954 ms for 250 M clean letters (full occupancy hides the latency)
#endif
constexpr int PARSE_CLEAN_UNROLL = QD_PARSE_UNROLL;
template <bool STORE>
__global__ void __launch_bounds__(PARSE_THREADS) parse_pass1_kernel(const ParseParams p, uint64_t n_gran, uint64_t* n_masks) {
    __shared__ uint8_t tab_s[256];
    __shared__ __align__(16) uint16_t pair_s[PAIR_ENTRIES];
    pdl_trigger();  // the gated passes behind this one may take their places early (they wait for this grid before they look at the gate)
    tab_s[threadIdx.x] = __ldg(p.table + threadIdx.x);
    for (int i = threadIdx.x; i < (int)PAIR_ENTRIES; i += PARSE_THREADS) pair_s[i] = __ldg(p.pair + i);
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0 && n_masks) *n_masks = p.n;  // right unless the gate opens (then the compacting pass overwrites it)
    const uint32_t pair_base = smem_u32(pair_s);
    const uintptr_t lo = reinterpret_cast<uintptr_t>(p.in), hi = lo + p.n, a0 = lo & ~(uintptr_t)15;
    uint8_t* const out0 = p.out - (lo & 15);  // granule g of the input lands at out0 + 16 g
    bool dirty = !STORE;
    constexpr uint64_t STEP = (uint64_t)PARSE_THREADS * PARSE_CLEAN_UNROLL;
    for (uint64_t g0 = (uint64_t)blockIdx.x * STEP; g0 < n_gran; g0 += (uint64_t)gridDim.x * STEP) {
        uint4 x[PARSE_CLEAN_UNROLL];
        bool whole[PARSE_CLEAN_UNROLL];
#pragma unroll
        for (int u = 0; u < PARSE_CLEAN_UNROLL; u++) {
            const uint64_t g = g0 + (uint64_t)u * PARSE_THREADS + threadIdx.x;
            const uintptr_t a = a0 + 16 * g;
            whole[u] = g < n_gran && a >= lo && a + 16 <= hi;
            if (whole[u]) x[u] = ldg_stream(reinterpret_cast<const void*>(a));
        }
#pragma unroll
        for (int u = 0; u < PARSE_CLEAN_UNROLL; u++) {
            const uint64_t g = g0 + (uint64_t)u * PARSE_THREADS + threadIdx.x;
            uint32_t dropped = 0;
            int first_bad = -1;
            if (whole[u]) {
                const uint32_t w[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
                uint32_t m[4], flags = 0;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint32_t t = w[q] & 0x1f1f1f1fu;
                    m[q] = lds_u16(dp2a_hi(p.pair_coef, t, pair_base)) * 65536u + lds_u16(dp2a_lo(p.pair_coef, t, pair_base));
                    flags |= m[q];
                }
                const uint32_t a_and = w[0] & w[1] & w[2] & w[3], a_or = w[0] | w[1] | w[2] | w[3];
                if ((flags & (PARSE_PAIR_SLOW * 0x01010101u)) | ((a_and & 0x40404040u) ^ 0x40404040u) | (a_or & 0x80808080u)) {
#pragma unroll
                    for (int k = 15; k >= 0; k--) {
                        const uint32_t t = tab_s[(w[k >> 2] >> (8 * (k & 3))) & 0xffu];
                        if (t == 0) first_bad = k;
                        dropped += (t == 0 || t == PARSE_SKIP) ? 1u : 0u;
                    }
                    if (STORE && dropped == 0) stg_stream(out0 + 16 * g, make_uint4(m[0], m[1], m[2], m[3]));  // (letters the pair table leaves to the byte table)
                } else if (STORE) {
                    stg_stream(out0 + 16 * g, make_uint4(m[0], m[1], m[2], m[3]));
                }
            } else if (g < n_gran) {  // a granule cut by the start or the end of the input
                for (int k = 15; k >= 0; k--) {
                    const uintptr_t b = a0 + 16 * g + k;
                    if (b < lo || b >= hi) continue;
                    const uint32_t t = tab_s[*reinterpret_cast<const uint8_t*>(b)];
                    if (t == 0) first_bad = k;
                    if (t == 0 || t == PARSE_SKIP) dropped++;
                    else if (STORE) out0[16 * g + k] = (uint8_t)t;
                }
            }
            if (first_bad >= 0) atomicMin(p.err_key, (unsigned long long)((a0 + 16 * g + first_bad) - lo));
            if (__any_sync(0xffffffffu, dropped != 0)) {  // a warp's 32 granules lie in one 4-KiB block
                dirty = true;
                const uint32_t sum = __reduce_add_sync(0xffffffffu, dropped);
                if ((threadIdx.x & 31) == 0) atomicAdd(reinterpret_cast<unsigned long long*>(p.counts + g / PARSE_THREADS), (unsigned long long)sum);
            }
        }
    }
    if (dirty) *reinterpret_cast<volatile unsigned long long*>(p.gate) = p.gate_gen;
}

// offsets[0 .. n_blocks] = exclusive scan of the kept bytes per 4-KiB block (the block's bytes inside the input minus its
// dropped bytes): one CTA, 8192 blocks per round -- only input with whitespace comes here
constexpr int PARSE_SCAN_THREADS = 1024;
__global__ void __launch_bounds__(PARSE_SCAN_THREADS) parse_scan_kernel(const ParseParams p, uint64_t n_blocks) {
    pdl_trigger();
    pdl_wait();
    if (parse_gate_closed(p.gate, p.gate_gen)) return;
    __shared__ uint64_t warp_sums[32];
    __shared__ uint64_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const uint64_t lead = reinterpret_cast<uintptr_t>(p.in) & 15, end = lead + p.n;  // the input in block coordinates
    uint64_t* const offsets = const_cast<uint64_t*>(p.offsets);
    constexpr int ITEMS = 8;
    for (uint64_t base = 0; base < n_blocks; base += (uint64_t)PARSE_SCAN_THREADS * ITEMS) {
        const uint64_t i0 = base + (uint64_t)threadIdx.x * ITEMS;
        uint64_t v[ITEMS], sum = 0;
#pragma unroll
        for (int k = 0; k < ITEMS; k++) {
            const uint64_t i = i0 + k;
            v[k] = 0;
            if (i < n_blocks) {
                const uint64_t b0 = max(i * PARSE_BLOCK_BYTES, lead), b1 = min((i + 1) * PARSE_BLOCK_BYTES, end);
                v[k] = (b1 > b0 ? b1 - b0 : 0) - p.counts[i];
            }
            sum += v[k];
        }
        uint64_t total;
        uint64_t ex = block_exclusive_scan(sum, &total, warp_sums);
        const uint64_t carry = carry_s;
        ex += carry;
#pragma unroll
        for (int k = 0; k < ITEMS; k++) {
            if (i0 + k < n_blocks) offsets[i0 + k] = ex;
            ex += v[k];
        }
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[n_blocks] = carry_s;
}

__global__ void __launch_bounds__(PARSE_THREADS) parse_write_kernel(const ParseParams p, uint64_t n_blocks, uint64_t* n_masks) {
    pdl_wait();
    if (parse_gate_closed(p.gate, p.gate_gen)) return;
    if (blockIdx.x == 0 && threadIdx.x == 0 && n_masks) *n_masks = p.offsets[n_blocks];
    __shared__ uint8_t tab_s[256];
    __shared__ uint64_t warp_sums[32];
    __shared__ __align__(16) uint8_t stage[PARSE_BLOCK_BYTES + 16];
    __shared__ __align__(16) uint16_t pair_s[PAIR_ENTRIES];
    tab_s[threadIdx.x] = __ldg(p.table + threadIdx.x);
    for (int i = threadIdx.x; i < (int)PAIR_ENTRIES; i += PARSE_THREADS) pair_s[i] = __ldg(p.pair + i);
    __syncthreads();
    const uint32_t pair_base = smem_u32(pair_s);
    const uintptr_t in_lo = reinterpret_cast<uintptr_t>(p.in);
    const bool same_phase = ((reinterpret_cast<uintptr_t>(p.out) ^ in_lo) & 15u) == 0;
    // This CTA owns blocks blockIdx.x + k * gridDim.x.  Which of them the first pass already finished is decided for
    // PARSE_THREADS of them at a time, one block per thread (a CTA-wide loop with the check inside pays one dependent
    // global load per block: 57 us for 250 M clean letters, against 3 us this way).
    __shared__ uint8_t todo_s[PARSE_THREADS];
    for (uint64_t k0 = 0; blockIdx.x + k0 * gridDim.x < n_blocks; k0 += PARSE_THREADS) {
        const uint64_t mine = blockIdx.x + (k0 + threadIdx.x) * gridDim.x;
        bool todo = false;
        if (mine < n_blocks) {
            // dense block (every byte a nucleotide: its output is as long as its input) that lands exactly where the first
            // pass guessed (nothing was dropped before it): already written
            const uint64_t o0 = p.offsets[mine], o1 = p.offsets[mine + 1];
            todo = !(same_phase && o1 - o0 == PARSE_BLOCK_BYTES && o0 + (in_lo & 15) == mine * (uint64_t)PARSE_BLOCK_BYTES);
            p.counts[mine] = 0;  // the scan has used it: all zero again for the next call
        }
        __syncthreads();  // the previous round no longer reads todo_s
        todo_s[threadIdx.x] = todo ? 1 : 0;
        if (!__syncthreads_or(todo ? 1 : 0)) continue;
        for (int j = 0; j < PARSE_THREADS; j++) {
            if (!todo_s[j]) continue;
            const uint64_t blk = blockIdx.x + (k0 + j) * gridDim.x;
            // dense block landing on a 16-byte boundary: no compaction, no staging -- translate my granule and store it
            if (p.offsets[blk + 1] - p.offsets[blk] == PARSE_BLOCK_BYTES && ((reinterpret_cast<uintptr_t>(p.out) + p.offsets[blk]) & 15u) == 0) {
                uint4 fast;
                if (parse_granule_fast(p, pair_base, blk * PARSE_THREADS + threadIdx.x, fast)) {  // always true here
                    stg_stream(p.out + p.offsets[blk] + 16 * threadIdx.x, fast);
                    continue;
                }
            }
            uint8_t v[16];
            uint32_t kept;
            int bad;
            parse_load(p, tab_s, blk * PARSE_THREADS + threadIdx.x, v, kept, bad);
            uint64_t total;
            const uint32_t off = (uint32_t)block_exclusive_scan((uint64_t)kept, &total, warp_sums);
            const uint64_t out0 = p.offsets[blk];
            uint8_t* gdst = p.out + out0;
            const uint32_t phase = (uint32_t)(reinterpret_cast<uintptr_t>(gdst) & 15u);  // stage mirrors the output's 16-byte phase
            uint32_t at = phase + off;
#pragma unroll
            for (int k = 0; k < 16; k++)
                if (v[k] != 0 && v[k] != PARSE_SKIP) stage[at++] = v[k];
            __syncthreads();
            const uint32_t nbytes = (uint32_t)total;
            const uint32_t head = min(nbytes, (16u - phase) & 15u);
            const uint32_t body = (nbytes - head) >> 4;
            if (threadIdx.x < head) gdst[threadIdx.x] = stage[phase + threadIdx.x];
            for (uint32_t c = threadIdx.x; c < body; c += PARSE_THREADS)
                stg_stream(gdst + head + 16 * c, *reinterpret_cast<const uint4*>(stage + phase + head + 16 * c));
            const uint32_t done = head + 16 * body;
            if (threadIdx.x < nbytes - done) gdst[done + threadIdx.x] = stage[phase + done + threadIdx.x];
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Canonical form of ONE sequence of any length (DnaSequence::canonical, src/rust_api.rs:374-377;
// src/canonical.rs:17-55, 71-118, 129-179).  ForwardCanonical relabels bases by first occurrence, so the forward map
// is fixed by the FIRST index of each base and the map of the reversed sequence by the LAST index of each base:
//   pass 1  first / last occurrence of each base (+ strict validation)          -> state[0..7]
//   pass 2  first i with fw[x[i]] != rv[x[n-1-i]] (LexicalMin decides there)     -> state[8]
//   pass 3  write the winner
// state: [0..3] first index of code c (~0 = absent), [4..7] last index + 1 (0 = absent), [8] first difference (~0 = none)
// ------------------------------------------------------------------------------------------------
struct CanonSeqParams {
    const uint8_t* dna;
    uint64_t n;
    const uint8_t* decode;  // 256-byte table -> internal code (0..3 concrete)
    uint32_t letters;       // output byte of A,T,C,G packed little-endian
    int forward_only;
    unsigned long long* state;  // 9 x u64
    uint8_t* out;
    unsigned long long* err_key;
};

__global__ void __launch_bounds__(256) canon_seq_scan_kernel(const CanonSeqParams p) {
    pdl_trigger();
    unsigned long long first[4] = {~0ull, ~0ull, ~0ull, ~0ull}, last1[4] = {0, 0, 0, 0};
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t c = __ldg(p.decode + __ldg(p.dna + i));
        if (c > 3) {
            atomicMin(p.err_key, (unsigned long long)i);
            continue;
        }
#pragma unroll
        for (int b = 0; b < 4; b++)
            if (c == (uint32_t)b) {
                first[b] = min(first[b], (unsigned long long)i);
                last1[b] = max(last1[b], (unsigned long long)i + 1);
            }
    }
#pragma unroll
    for (int b = 0; b < 4; b++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            first[b] = min(first[b], __shfl_xor_sync(0xffffffffu, first[b], o));
            last1[b] = max(last1[b], __shfl_xor_sync(0xffffffffu, last1[b], o));
        }
        if ((threadIdx.x & 31) == 0) {
            if (first[b] != ~0ull) atomicMin(p.state + b, first[b]);
            if (last1[b] != 0) atomicMax(p.state + 4 + b, last1[b]);
        }
    }
}

// relabelling maps (2 bits per code) from the occurrence indices: forward = rank by first index, reverse = rank by
// descending last index; absent bases get rank 3 (never used)
__device__ __forceinline__ void canon_seq_maps(const unsigned long long* state, uint32_t& fw, uint32_t& rv) {
    unsigned long long f[4], l[4];
#pragma unroll
    for (int b = 0; b < 4; b++) {
        f[b] = state[b];
        l[b] = state[4 + b];
    }
    fw = 0;
    rv = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) {
        uint32_t rf = 0, rr = 0;
#pragma unroll
        for (int o = 0; o < 4; o++) {
            rf += (o != b && f[o] < f[b]) ? 1u : 0u;
            rr += (o != b && l[o] > l[b]) ? 1u : 0u;
        }
        fw |= (f[b] == ~0ull ? 3u : rf) << (2 * b);
        rv |= (l[b] == 0 ? 3u : rr) << (2 * b);
    }
}

__global__ void __launch_bounds__(256) canon_seq_diff_kernel(const CanonSeqParams p) {
    pdl_trigger();
    pdl_wait();
    uint32_t fw, rv;
    canon_seq_maps(p.state, fw, rv);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (uint64_t)gridDim.x * blockDim.x) {
        if (i >= *reinterpret_cast<volatile unsigned long long*>(p.state + 8)) return;  // a smaller difference is known
        const uint32_t a = __ldg(p.decode + __ldg(p.dna + i)) & 3u, b = __ldg(p.decode + __ldg(p.dna + (p.n - 1 - i))) & 3u;
        if (((fw >> (2 * a)) & 3u) != ((rv >> (2 * b)) & 3u)) {
            atomicMin(p.state + 8, (unsigned long long)i);
            return;
        }
    }
}

__global__ void __launch_bounds__(256) canon_seq_write_kernel(const CanonSeqParams p) {
    pdl_trigger();
    pdl_wait();
    uint32_t fw, rv;
    canon_seq_maps(p.state, fw, rv);
    bool use_rev = false;
    const unsigned long long d = p.state[8];
    if (!p.forward_only && d != ~0ull) {
        const uint32_t a = __ldg(p.decode + __ldg(p.dna + d)) & 3u, b = __ldg(p.decode + __ldg(p.dna + (p.n - 1 - d))) & 3u;
        use_rev = ((rv >> (2 * b)) & 3u) < ((fw >> (2 * a)) & 3u);
    }
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t c = __ldg(p.decode + __ldg(p.dna + (use_rev ? p.n - 1 - i : i))) & 3u;
        const uint32_t y = ((use_rev ? rv : fw) >> (2 * c)) & 3u;
        p.out[i] = (uint8_t)(p.letters >> (8 * y));
    }
}

// ------------------------------------------------------------------------------------------------
// FASTA record scanner (FastaParser::<DnaSequence<T>>::parse, src/fasta.rs:207-491): the line state machine of the
// reference as data-parallel passes over the text.
//   line type   a line is a header iff its first byte is '>' or ';'; the type of the line a byte belongs to is that of
//               the last line start at or before it -> "last writer wins" = a MAX scan over (position, type)
//   records     a header line starts a record unless headers are concatenated and the previous line is a header too;
//               content before the first header is a headerless record (or dropped, unvalidated, when a preceding
//               comment is allowed)
//   content     bytes of non-header lines minus line breaks ('\n', and a '\r' right before it), ' ' and '\t', each
//               validated and turned into a mask (src/rust_api.rs:310-322) -> stream compaction
// pass 1: last line start per 4 KiB block + first header position.  max-scan over blocks.
// pass 2: per block kept bytes / record starts / line breaks (+ first error).  three sum scans.
// pass 3: compacted masks + one entry per record.  pass 4: per record, the *_end fields and the header extent.
// ------------------------------------------------------------------------------------------------
struct FastaRecord {  // == qd_fasta_record
    uint64_t header_begin, header_end, content_begin, content_end, line_begin, line_end;
};
constexpr int FA_THREADS = 256;
constexpr int FA_BLOCK_BYTES = FA_THREADS * 16;
constexpr uint64_t FA_NONE = ~0ull;

struct FastaParams {
    const uint8_t* text;
    uint64_t n;
    const uint8_t* table;  // 256 bytes: mask | PARSE_SKIP | 0 (the parse table of the chosen alphabet)
    int concatenate_headers, allow_preceding_comment;
    uint64_t n_blocks;
    uint64_t* blk_last;         // pass 1: ((position + 1) << 2 | type) of the block's last line start, 0 if none
    const uint64_t* blk_state;  // exclusive max scan of blk_last
    unsigned long long* scalars;  // [0] position of the first header line (FA_NONE if none), [1] kept bytes before it
                                  // inside its own block
    uint64_t* cnt_kept;
    uint64_t* cnt_recs;
    uint64_t* cnt_nl;
    const uint64_t* pre_kept;  // exclusive sum scans, [n_blocks + 1]
    const uint64_t* pre_recs;
    const uint64_t* pre_nl;
    uint8_t* masks;
    FastaRecord* records;
    uint64_t max_records;
    unsigned long long* err_key;
};

__device__ __forceinline__ bool fa_is_header_byte(uint32_t c) { return c == '>' || c == ';'; }

// the 16 bytes of granule g (bytes outside [text, text + n) come back as 0 with in_range false), the byte before and
// the byte after
struct FaGranule {
    uint8_t b[16];
    uint64_t i0;        // text index of b[0] (may be "negative" = wrapped when the granule starts before the text)
    uint32_t lo, hi;    // valid k range [lo, hi)
    uint32_t prev, next;  // text[i0 + lo - 1] ('\n' at the file start), text[i0 + hi] (0 at the end)
};
__device__ __forceinline__ void fa_load(const FastaParams& p, uint64_t g, FaGranule& G) {
    const uintptr_t lo = reinterpret_cast<uintptr_t>(p.text), hi = lo + p.n;
    const uintptr_t a = (lo & ~(uintptr_t)15) + 16 * g;
    G.i0 = (uint64_t)(a - lo);
    G.lo = a < lo ? (uint32_t)(lo - a) : 0u;
    G.hi = a + 16 <= hi ? 16u : (a < hi ? (uint32_t)(hi - a) : 0u);
    if (G.lo == 0 && G.hi == 16) {
        const uint4 x = ldg_cached(reinterpret_cast<const void*>(a));
        const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int k = 0; k < 16; k++) G.b[k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
    } else {
#pragma unroll
        for (int k = 0; k < 16; k++) G.b[k] = (k >= (int)G.lo && k < (int)G.hi) ? *reinterpret_cast<const uint8_t*>(a + k) : (uint8_t)0;
    }
    G.prev = '\n';
    G.next = 0;
    if (G.lo < G.hi) {
        if (a + G.lo > lo) G.prev = *reinterpret_cast<const uint8_t*>(a + G.lo - 1);
        if (a + G.hi < hi) G.next = *reinterpret_cast<const uint8_t*>(a + G.hi);
    }
}

// ((position + 1) << 2 | type) of the last line start among the granule's bytes, 0 if none; type 3 = header, 2 = other
__device__ __forceinline__ uint64_t fa_last_line_start(const FaGranule& G, unsigned long long* first_header) {
    uint64_t last = 0;
    uint32_t prevc = G.prev;
    for (uint32_t k = G.lo; k < G.hi; k++) {
        const uint32_t c = G.b[k];
        if (prevc == '\n') {
            const bool h = fa_is_header_byte(c);
            last = ((G.i0 + k + 1) << 2) | (h ? 3u : 2u);
            if (h && first_header) atomicMin(first_header, (unsigned long long)(G.i0 + k));
        }
        prevc = c;
    }
    return last;
}

__global__ void __launch_bounds__(FA_THREADS) fasta_lines_kernel(const FastaParams p) {
    pdl_trigger();
    __shared__ uint64_t warp_sums[32];
    for (uint64_t blk = blockIdx.x; blk < p.n_blocks; blk += gridDim.x) {
        FaGranule G;
        fa_load(p, blk * FA_THREADS + threadIdx.x, G);
        const uint64_t last = fa_last_line_start(G, p.scalars);
        uint64_t total;
        block_exclusive_scan_op<1>(last, &total, warp_sums);
        if (threadIdx.x == 0) p.blk_last[blk] = total;
        __syncthreads();
    }
}

template <bool WRITE>
__global__ void __launch_bounds__(FA_THREADS) fasta_pass_kernel(const FastaParams p) {
    pdl_trigger();
    pdl_wait();
    __shared__ uint8_t tab_s[256];
    __shared__ uint64_t warp_sums[32];
    __shared__ __align__(16) uint8_t stage[FA_BLOCK_BYTES + 16];
    tab_s[threadIdx.x] = __ldg(p.table + threadIdx.x);
    __syncthreads();
    const uint64_t first_header = p.scalars[0];
    // records are shifted by one when content before the first header makes a headerless record
    uint64_t shift = 0;
    if (WRITE && !p.allow_preceding_comment) {
        const uint64_t before = first_header == FA_NONE ? p.pre_kept[p.n_blocks]
                                                        : p.pre_kept[(((reinterpret_cast<uintptr_t>(p.text) & 15) + first_header) / FA_BLOCK_BYTES)] + p.scalars[1];
        shift = before > 0 ? 1 : 0;
    }
    for (uint64_t blk = blockIdx.x; blk < p.n_blocks; blk += gridDim.x) {
        FaGranule G;
        fa_load(p, blk * FA_THREADS + threadIdx.x, G);
        uint64_t total;
        // type of the line my first byte's predecessor belongs to
        const uint64_t mine = fa_last_line_start(G, nullptr);
        uint64_t before = block_exclusive_scan_op<1>(mine, &total, warp_sums);
        __syncthreads();
        before = max(before, p.blk_state[blk]);
        uint32_t cur_type = before ? (uint32_t)(before & 3u) : 2u;
        // ---- walk my bytes: what is kept, where records start, how many lines end ----
        uint8_t keep[16];
        uint32_t n_keep = 0, n_rec = 0, n_nl = 0, rec_at[8], rec_keep[8], rec_nl[8];  // a record start needs >= 2 bytes: at most 8
        uint32_t prevc = G.prev;
        for (uint32_t k = G.lo; k < G.hi; k++) {
            const uint32_t c = G.b[k];
            const uint64_t i = G.i0 + k;
            if (prevc == '\n') {
                const uint32_t prev_type = cur_type;
                cur_type = fa_is_header_byte(c) ? 3u : 2u;
                if (cur_type == 3u && (!p.concatenate_headers || prev_type != 3u)) {
                    rec_at[n_rec] = k;
                    rec_keep[n_rec] = n_keep;
                    rec_nl[n_rec] = n_nl;
                    n_rec++;
                }
            }
            if (c == '\n') n_nl++;
            else if (cur_type == 2u) {
                const uint32_t nextc = k + 1 < G.hi ? (uint32_t)G.b[k + 1] : G.next;
                const bool line_break_cr = c == '\r' && nextc == '\n';
                const bool dropped_comment = p.allow_preceding_comment && i < first_header;
                if (!line_break_cr && !dropped_comment) {
                    const uint32_t t = tab_s[c];
                    if (t == 0) { if (!WRITE) atomicMin(p.err_key, (unsigned long long)i); }
                    else if (t != PARSE_SKIP) keep[n_keep++] = (uint8_t)t;
                }
            }
            prevc = c;
        }
        // ---- block-level prefixes ----
        const uint64_t ex_keep = block_exclusive_scan_op<0>(n_keep, &total, warp_sums);
        const uint64_t blk_keep = total;
        __syncthreads();
        const uint64_t ex_rec = block_exclusive_scan_op<0>(n_rec, &total, warp_sums);
        const uint64_t blk_rec = total;
        __syncthreads();
        const uint64_t ex_nl = block_exclusive_scan_op<0>(n_nl, &total, warp_sums);
        const uint64_t blk_nl = total;
        __syncthreads();
        if (!WRITE) {
            if (threadIdx.x == 0) {
                p.cnt_kept[blk] = blk_keep;
                p.cnt_recs[blk] = blk_rec;
                p.cnt_nl[blk] = blk_nl;
            }
            // kept bytes of this block before the first header line (it is the first record start of its thread)
            if (first_header != FA_NONE && first_header >= G.i0 + G.lo && first_header < G.i0 + G.hi && G.lo < G.hi) {
                uint32_t kb = 0;
                for (uint32_t r = 0; r < n_rec; r++)
                    if (G.i0 + rec_at[r] == first_header) kb = rec_keep[r];
                p.scalars[1] = ex_keep + kb;
            }
        } else {
            const uint64_t out0 = p.pre_kept[blk];
            uint8_t* gdst = p.masks + out0;
            const uint32_t phase = (uint32_t)(reinterpret_cast<uintptr_t>(gdst) & 15u);  // stage mirrors the output's 16-byte phase
            for (uint32_t j = 0; j < n_keep; j++) stage[phase + ex_keep + j] = keep[j];
            for (uint32_t r = 0; r < n_rec; r++) {
                const uint64_t idx = shift + p.pre_recs[blk] + ex_rec + r;
                if (idx < p.max_records) {
                    FastaRecord rec;
                    rec.header_begin = G.i0 + rec_at[r];
                    rec.header_end = 0;
                    rec.content_begin = out0 + ex_keep + rec_keep[r];
                    rec.content_end = 0;
                    rec.line_begin = p.pre_nl[blk] + ex_nl + rec_nl[r] + 1;
                    rec.line_end = 0;
                    p.records[idx] = rec;
                }
            }
            __syncthreads();
            const uint32_t nbytes = (uint32_t)blk_keep;
            const uint32_t head = min(nbytes, (16u - phase) & 15u), body = (nbytes - head) >> 4;
            if (threadIdx.x < head) gdst[threadIdx.x] = stage[phase + threadIdx.x];
            for (uint32_t c = threadIdx.x; c < body; c += FA_THREADS)
                stg_stream(gdst + head + 16 * c, *reinterpret_cast<const uint4*>(stage + phase + head + 16 * c));
            const uint32_t done = head + 16 * body;
            if (threadIdx.x < nbytes - done) gdst[done + threadIdx.x] = stage[phase + done + threadIdx.x];
            __syncthreads();
        }
    }
}

// after the count pass and its scans: what the host needs before it can size the second pass, gathered into five words (of
// mapped host memory: no copy) -- kept bytes, record starts, kept bytes before the first header line, the first invalid byte
__global__ void __launch_bounds__(32) fasta_totals_kernel(const FastaParams p, unsigned long long* __restrict__ out) {
    pdl_wait();
    if (threadIdx.x != 0) return;
    const uint64_t kept = p.pre_kept[p.n_blocks], first_header = p.scalars[0];
    uint64_t before = kept;
    if (first_header != FA_NONE) before = p.pre_kept[((reinterpret_cast<uintptr_t>(p.text) & 15) + first_header) / FA_BLOCK_BYTES] + p.scalars[1];
    out[0] = kept;
    out[1] = p.pre_recs[p.n_blocks];
    out[2] = before;
    out[3] = *p.err_key;
}

// pass 4: one thread per record: where its content and its lines end (= where the next record's begin), and the extent
// of its header lines in the text; record 0 is the headerless one when `shift` says so
__global__ void __launch_bounds__(256) fasta_finish_kernel(const FastaParams p, uint64_t n_records, uint64_t shift) {
    pdl_trigger();
    pdl_wait();
    const uint64_t total_kept = p.pre_kept[p.n_blocks];
    const uint64_t n_lines = p.pre_nl[p.n_blocks] + ((p.n > 0 && p.text[p.n - 1] != '\n') ? 1 : 0);
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_records && r < p.max_records; r += (uint64_t)gridDim.x * blockDim.x) {
        const bool has_next = r + 1 < n_records && r + 1 < p.max_records;
        const uint64_t next_content = has_next ? p.records[r + 1].content_begin : total_kept;
        const uint64_t next_line = has_next ? p.records[r + 1].line_begin : n_lines + 1;
        if (r == 0 && shift) {
            FastaRecord rec = {FA_NONE, FA_NONE, 0, next_content, 1, next_line};
            p.records[0].header_begin = rec.header_begin;
            p.records[0].header_end = rec.header_end;
            p.records[0].content_begin = 0;
            p.records[0].content_end = rec.content_end;
            p.records[0].line_begin = 1;
            p.records[0].line_end = rec.line_end;
            continue;
        }
        uint64_t pos = p.records[r].header_begin, header_end = pos;
        for (;;) {  // the header line starting at pos, then (concatenated headers) the header lines right after it
            uint64_t eol = pos;
            while (eol < p.n && p.text[eol] != '\n') eol++;
            header_end = (eol < p.n && eol > pos && p.text[eol - 1] == '\r') ? eol - 1 : eol;
            if (!p.concatenate_headers || eol + 1 >= p.n || !fa_is_header_byte(p.text[eol + 1])) break;
            pos = eol + 1;
        }
        p.records[r].header_end = header_end;
        p.records[r].content_end = next_content;
        p.records[r].line_end = next_line;
    }
}

// ------------------------------------------------------------------------------------------------
// Position-weighted linear checksum of a byte stream (qd_checksum64_dev): stream word J (bytes 8J .. 8J+7, little
// endian) weighs K(J) = mix64(J * GOLDEN + 1) | 1; the checksum is sum_J word_J * K(J) mod 2^64.  Linear in the data, so
// pieces of a stream -- the chunks of an output ring, the shards of the GPUs of a box -- add up to the checksum of the
// whole whatever their byte alignment.  HBM-bound read (8 B per mix64).
// CSR offsets of the records of a FASTA scan (what the batch entry points take): offsets[i] = records[i].content_begin,
// offsets[n_records] = the number of masks
__global__ void __launch_bounds__(256) fasta_offsets_kernel(const FastaRecord* __restrict__ records, uint64_t n_records, uint64_t total_kept,
                                                            uint64_t* __restrict__ offsets) {
    pdl_trigger();
    pdl_wait();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n_records; i += (uint64_t)gridDim.x * blockDim.x)
        offsets[i] = i < n_records ? records[i].content_begin : total_kept;
}

// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t checksum_weight(uint64_t word_index) {
    return canon_mix64(word_index * 0x9e3779b97f4a7c15ull + 1ull) | 1ull;
}

struct ChecksumParams {
    const uint8_t* buf;
    uint64_t n;                  // bytes (ignored when n_dev is set)
    const uint64_t* n_dev;       // optional: element count on the device ...
    uint64_t n_scale;            // ... times bytes per element
    uint64_t base;               // stream offset of buf[0]
    unsigned long long* acc;     // *acc += checksum
};

constexpr int CHECKSUM_THREADS = 256;
__global__ void __launch_bounds__(CHECKSUM_THREADS) checksum64_kernel(const ChecksumParams p) {
    const uint64_t n = p.n_dev ? *p.n_dev * p.n_scale : p.n;
    const uintptr_t lo = reinterpret_cast<uintptr_t>(p.buf), hi = lo + n;
    const uintptr_t b0 = min(hi, (lo + 15) & ~(uintptr_t)15), b1 = max(b0, hi & ~(uintptr_t)15);  // aligned body [b0, b1)
    uint64_t sum = 0;
    const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, gsize = (uint64_t)gridDim.x * blockDim.x;
    // head and tail bytes (fewer than 32 in all)
    if (gtid < 32) {
        const uint64_t n_head = b0 - lo, n_tail = hi - b1;
        if (gtid < n_head + n_tail) {
            const uintptr_t a = gtid < n_head ? lo + gtid : b1 + (gtid - n_head);
            const uint64_t g = p.base + (a - lo);
            sum += ((uint64_t)*reinterpret_cast<const uint8_t*>(a) << (8 * (g & 7))) * checksum_weight(g >> 3);
        }
    }
    const uint64_t g0 = p.base + (b0 - lo);       // stream offset of the body
    const uint32_t sh = (uint32_t)(g0 & 7) * 8u;  // the same for every body word
    const uint64_t n_gran = (b1 - b0) >> 4;
    const uint4* body = reinterpret_cast<const uint4*>(b0);
    for (uint64_t c = gtid; c < n_gran; c += gsize) {
        const uint4 v = ldg_stream(body + c);
        const uint64_t w0 = (uint64_t)v.x | ((uint64_t)v.y << 32), w1 = (uint64_t)v.z | ((uint64_t)v.w << 32);
        const uint64_t J = (g0 >> 3) + 2 * c;
        const uint64_t k0 = checksum_weight(J), k1 = checksum_weight(J + 1);
        if (sh == 0) {
            sum += w0 * k0 + w1 * k1;
        } else {
            const uint64_t k2 = checksum_weight(J + 2);
            sum += (w0 << sh) * k0 + ((w0 >> (64 - sh)) | (w1 << sh)) * k1 + (w1 >> (64 - sh)) * k2;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __shared__ unsigned long long warp_sum[CHECKSUM_THREADS / 32];
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long s = 0;
#pragma unroll
        for (int i = 0; i < CHECKSUM_THREADS / 32; i++) s += warp_sum[i];
        if (s) atomicAdd(p.acc, s);
    }
}

}  // namespace qd

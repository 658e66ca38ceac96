// quickdna_b200 -- C ABI implementation (include/quickdna_b200.h) over the sm_100a kernels.
//
// No CPU fallback: every compute entry point launches kernels and returns QD_ERR_CUDA when that
// is impossible.  Host code here only moves bytes, sizes buffers and classifies the one offending
// byte of an error the way src/nucleotide.rs:268-315 does.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/quickdna_b200.h"
#include "qd_kernels.cuh"
#include "qd_tables.hpp"

using namespace qd;

namespace {

const HostTables& host_tables() {
    static HostTables* T = [] {
        auto* t = new HostTables();
        build_host_tables(*t);
        return t;
    }();
    return *T;
}

// Grow-only scratch.  Growth is stream-ordered (cudaFreeAsync / cudaMallocAsync on the stream the buffer is about to be
// used on): the old block is released only after the work already enqueued on that stream, nothing synchronises the
// device, and a call whose buffers are warm allocates nothing.  `tls_stream` is the stream of the entry point running
// on this thread (set by its Guard).
thread_local cudaStream_t tls_stream = nullptr;
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeAsync(p, tls_stream);
        p = nullptr;
        cap = 0;
        size_t want = (bytes + (size_t(1) << 20)) & ~((size_t(1) << 20) - 1);  // 1 MiB granules, >= 16 B slack
        cudaError_t e = cudaMallocAsync(&p, want, tls_stream);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

constexpr int N_PIPE = 3;  // streams / buffer sets of the chunked host pipeline
// host calls whose device result is at most this big come back as ONE device->host copy [result | error key] into pinned
// memory (the reference's typical call is a 30-kb genome: launch and copy count is what it costs here)
constexpr size_t SMALL_CALL_BYTES = size_t(2) << 20;
// Tiny host calls (the reference's typical call: one 30-kb genome) skip the copies altogether: the kernel reads its input from
// and writes its result and error key to MAPPED pinned host memory of the context, so a call is one host memcpy in, ONE
// launch, one synchronise and one memcpy out instead of five runtime calls (H2D, memset, launch, D2H, synchronise).
constexpr size_t TINY_IN_BYTES = size_t(128) << 10, TINY_OUT_BYTES = size_t(288) << 10;  // input / result capacity of the mapped block

// Pageable host memory (a Rust Vec<u8>, a Python bytes object, a numpy array) cannot be the source or target of an
// asynchronous copy: the driver stages it through one internal buffer on the calling thread, and the chunk pipeline
// serialises (250 M nt reverse-complemented host to host: 39 ms against 6.7 ms from pinned buffers).  Large calls on
// pageable buffers therefore go through pinned staging buffers of the context, filled and drained by a few persistent
// host threads (HostPool::copy), so the PCIe copies stay asynchronous and overlap with the host-side memcpy of other chunks.
class HostPool {
public:
    explicit HostPool(int n) {
        try {
            for (int i = 0; i < n; i++) workers_.emplace_back([this, i, n] { loop(i, n); });
        } catch (...) {
            shutdown();
            throw;
        }
    }
    ~HostPool() { shutdown(); }
    // blocking memcpy, split over the workers
    void copy(void* dst, const void* src, size_t n) {
        if (n < (size_t(1) << 20) || workers_.empty()) {
            memcpy(dst, src, n);
            return;
        }
        std::unique_lock<std::mutex> lk(m_);
        dst_ = static_cast<uint8_t*>(dst);
        src_ = static_cast<const uint8_t*>(src);
        n_ = n;
        pending_ = workers_.size();
        gen_++;
        work_.notify_all();
        done_.wait(lk, [this] { return pending_ == 0; });
    }

private:
    void shutdown() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        work_.notify_all();
        for (auto& t : workers_) t.join();
        workers_.clear();
    }
    void loop(int i, int n_workers) {
        uint64_t seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(m_);
            work_.wait(lk, [&] { return stop_ || gen_ != seen; });
            if (stop_) return;
            seen = gen_;
            uint8_t* const dst = dst_;
            const uint8_t* const src = src_;
            const size_t n = n_;
            lk.unlock();
            const size_t piece = ((n + n_workers - 1) / n_workers + 63) & ~size_t(63);
            const size_t a = std::min(n, piece * i), b = std::min(n, a + piece);
            if (b > a) memcpy(dst + a, src + a, b - a);
            lk.lock();
            if (--pending_ == 0) done_.notify_one();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable work_, done_;
    uint8_t* dst_ = nullptr;
    const uint8_t* src_ = nullptr;
    size_t n_ = 0, pending_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
};

struct PinBuf {  // grow-only pinned host buffer
    uint8_t* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        const size_t want = (bytes + (size_t(1) << 20)) & ~((size_t(1) << 20) - 1);
        cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&p), want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

struct Pipe {
    cudaStream_t stream = nullptr;
    DevBuf in, out, offsets, run_prefix, tile_first, block_sums;
    PinBuf hin, hout;            // staging for pageable caller buffers
    cudaEvent_t done = nullptr;  // the chunk last staged through hin / hout has left the device
};

}  // namespace

struct qd_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    std::string last_error;
    uint64_t launches = 0;
    // device-resident tables
    uint16_t* d_lut = nullptr;  // 27 * 4096
    uint8_t* d_tables = nullptr;  // 256-byte tables, see TAB_* offsets
    uint16_t* d_pair = nullptr;   // [encoding][strict][PAIR_ENTRIES]
    uint16_t* d_rc_pair = nullptr;  // [RC_PAIR_COUNT][PAIR_ENTRIES]
    uint8_t* d_ckey_pair = nullptr;   // [encoding][PAIR_ENTRIES]
    CanonKeyTables* d_ckey = nullptr;
    uint16_t* d_parse_pair = nullptr;  // [strict][PAIR_ENTRIES]
    unsigned long long* d_gate = nullptr;  // generation of the last optimistic pass that failed (ParseParams::gate)
    unsigned long long gate_gen = 0;
    std::map<int, uint8_t*> d_direct;  // one-frame direct codon tables, built on first use: key = slot * 4 + ascii * 2 + strict
    unsigned long long* d_err = nullptr;
    unsigned long long* cur_err_key = nullptr;  // set by small_call for the length of one call: the key behind the result
    unsigned long long* h_err = nullptr;  // pinned
    uint64_t* h_scratch = nullptr;        // pinned, 8 x u64
    uint8_t* h_stage = nullptr;           // pinned staging for small host calls: [result | error key] in ONE copy
    uint8_t* h_tiny = nullptr;            // mapped pinned block of the tiny-call path: [input | error key | result]
    bool tiny_calls = true;               // QD_TINY_CALLS=0 turns the mapped path off (A/B measurements)
    DevBuf in, out, aux0, aux1, aux2, aux3, run_prefix, tile_first, block_sums, parse_counts, parse_offsets, canon_state, fasta_scratch;
    Pipe pipe[N_PIPE];
    size_t chunk_bytes = size_t(64) << 20;   // input bytes per pipeline chunk
    HostPool* pool = nullptr;                // host threads of the pageable staging path, started on first use
    int stage_threads = -1;                  // -1: pick from the host's thread count; 0: no staging (plain copies)
    size_t stage_chunk_bytes = size_t(16) << 20;  // input bytes per chunk when staging (bounds the pinned memory)
    int occ_frames[6] = {0, 0, 0, 0, 0, 0};  // [tile size][frames]
    uint64_t big_tile_min_runs = 0;           // launches with at least this many runs use 1024-thread tiles
    int occ_revcomp[2] = {0, 0};  // [small, big tile]
    // Scratch and the error key are per context, not per stream: a device-pointer call leaves an event behind, and the
    // next call that runs on ANOTHER stream waits for it first (host-pointer calls wait on the host), so two calls on one
    // context are ordered whatever streams they use.
    cudaEvent_t dev_done = nullptr;
    cudaStream_t dev_stream = nullptr;
    bool dev_pending = false;
    std::map<uint64_t, int> occ_cache;  // (kernel id, smem bytes) -> resident CTAs per SM, attribute already set
    bool pdl = true;  // QD_PDL=0: plain launches for the short dependent kernels (A/B measurements)
    bool aligned_tiles = true;  // QD_ALIGNED_TILES=0 turns the whole-sequence tiles of the frame kernels off (A/B measurements)
    uint32_t ctile_wpt[2] = {960, 448};  // canonical tile kernel: windows per warp-tile for keys / rows (QD_CTILE_WPT_* read once)
};

namespace {

enum {
    TAB_DECODE_ASCII = 0,
    TAB_DECODE_MASK,
    TAB_RC_ASCII,
    TAB_RC_ASCII_STRICT,
    TAB_RC_MASK,
    TAB_RC_MASK_STRICT,
    TAB_RC_MASK_TO_ASCII,
    TAB_RC_MASK_TO_ASCII_STRICT,
    TAB_UPPER_ASCII,
    TAB_UPPER_ASCII_STRICT,
    TAB_MASK_TO_ASCII,
    TAB_MASK_TO_ASCII_STRICT,
    TAB_TO_MASK_ASCII,
    TAB_TO_MASK_MASK,
    TAB_PARSE,
    TAB_PARSE_STRICT,
    TAB_COUNT
};

#define QD_CUDA(ctx, expr)                                                                    \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            (ctx)->last_error = std::string(#expr) + ": " + cudaGetErrorString(e__);          \
            return QD_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

// Every entry point: lock the context, make its device current (restored on return), clear the last error text, and
// order this call after a pending device-pointer call that ran on another stream.
struct Guard {
    qd_ctx* c;
    std::unique_lock<std::mutex> lk;
    int prev = -1;
    bool ok = true;
    bool dev;
    cudaStream_t s;
    // host-pointer entry points: Guard(c);  device-pointer entry points: Guard(c, stream)
    explicit Guard(qd_ctx* c_) : Guard(c_, nullptr, false) {}
    Guard(qd_ctx* c_, cudaStream_t stream, bool dev_ = true) : c(c_), lk(c_->mu), dev(dev_), s(dev_ ? stream : c_->stream) {
        c->last_error.clear();
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != c->device && cudaSetDevice(c->device) != cudaSuccess) {
            fail("cudaSetDevice");
            return;
        }
        tls_stream = s;
        if (c->dev_pending && (!dev || c->dev_stream != s)) {
            const cudaError_t e = dev ? cudaStreamWaitEvent(s, c->dev_done, 0) : cudaEventSynchronize(c->dev_done);
            if (e != cudaSuccess) {
                fail("waiting for the previous device-pointer call");
                return;
            }
            if (!dev) c->dev_pending = false;
        }
    }
    void fail(const char* what) {
        ok = false;
        c->last_error = std::string(what) + " failed";
    }
    ~Guard() {
        if (ok && dev) {
            if (cudaEventRecord(c->dev_done, s) == cudaSuccess) {
                c->dev_pending = true;
                c->dev_stream = s;
            }
        }
        if (prev >= 0 && prev != c->device) cudaSetDevice(prev);
    }
};

inline const uint8_t* tab(const qd_ctx* c, int which) { return c->d_tables + 256 * which; }

void clear_err(qd_err* e) {
    if (e) {
        e->kind = QD_OK;
        e->byte = 0;
        e->pos = 0;
        e->seq_index = 0;
    }
}

int set_err(qd_err* e, int kind, uint8_t byte, uint64_t pos, uint64_t seq) {
    if (e) {
        e->kind = kind;
        e->byte = byte;
        e->pos = pos;
        e->seq_index = seq;
    }
    return kind;
}

// src/nucleotide.rs:285-315 + :268-283 for the one byte the kernels singled out
int classify_byte(int enc, int strict, uint8_t b, uint8_t* payload) {
    if (enc == QD_ENC_ASCII) {
        if (b >= 128) {
            *payload = b;
            return QD_ERR_NON_ASCII_BYTE;
        }
        const uint8_t m = mask_of_ascii(b);
        if (!m) {
            *payload = b;
            return QD_ERR_BAD_NUCLEOTIDE;
        }
        if (strict && popcount4(m) > 1) {
            *payload = ascii_of_mask(m);
            return QD_ERR_UNEXPECTED_AMBIGUOUS;
        }
    } else {
        if (b < 1 || b > 15) {
            *payload = b;
            return QD_ERR_BAD_MASK;
        }
        if (strict && popcount4(b) > 1) {
            *payload = ascii_of_mask(b);
            return QD_ERR_UNEXPECTED_AMBIGUOUS;
        }
    }
    *payload = b;
    return QD_OK;
}

int check_table(int table, qd_err* err, int* slot) {
    *slot = (table < 0 || table > 255) ? -1 : qd_table_slot(table);
    if (*slot < 0) return set_err(err, QD_ERR_BAD_TABLE, (uint8_t)table, 0, 0);
    return QD_OK;
}

bool valid_frames(int f) { return f == QD_FRAMES_ONE || f == QD_FRAMES_SELF || f == QD_FRAMES_ALL; }

// direct codon table of the one-frame kernel for (slot, enc, strict), built and uploaded on first use
int direct_table(qd_ctx* c, int slot, int enc, int strict, cudaStream_t s, const uint8_t** out) {
    const bool ascii = enc == QD_ENC_ASCII;
    const int key = slot * 4 + (ascii ? 2 : 0) + (strict ? 1 : 0);
    auto it = c->d_direct.find(key);
    if (it == c->d_direct.end()) {
        std::vector<uint8_t> host;
        build_direct_table(host_tables(), slot, ascii, strict != 0, host);
        uint8_t* d = nullptr;
        QD_CUDA(c, cudaMalloc(&d, DIRECT_ENTRIES));
        QD_CUDA(c, cudaMemcpyAsync(d, host.data(), DIRECT_ENTRIES, cudaMemcpyHostToDevice, s));
        QD_CUDA(c, cudaStreamSynchronize(s));  // `host` goes away
        it = c->d_direct.emplace(key, d).first;
    }
    *out = it->second;
    return QD_OK;
}

void set_encoding(const qd_ctx* c, FramesParams& p, int enc, int strict) {
    const bool ascii = enc == QD_ENC_ASCII;
    p.decode = tab(c, ascii ? TAB_DECODE_ASCII : TAB_DECODE_MASK);
    p.pair = c->d_pair + ((ascii ? 0 : 2) + (strict ? 1 : 0)) * PAIR_ENTRIES;
    const uint32_t k1 = ascii ? PAIR_K1_ASCII : PAIR_K1_MASK;
    p.pair_coef = 2u | ((2u * k1) << 16);
    p.and_need = ascii ? 0x40404040u : 0u;
    p.or_forbid = ascii ? 0x80808080u : 0xe0e0e0e0u;
    p.direct_k1 = ascii ? DIRECT_K1_ASCII : DIRECT_K1_MASK;
    p.direct_k2 = ascii ? DIRECT_K2_ASCII : DIRECT_K2_MASK;
    p.strict = strict;
}

// Which CTA size a launch of `total_runs` runs uses: the 1024-thread tile once every SM gets at least two of them.
bool use_big_tiles(const qd_ctx* c, uint64_t total_runs) { return total_runs >= c->big_tile_min_runs; }

// launch with programmatic stream serialization: the kernel may become resident before its predecessor in the stream has
// finished and calls pdl_wait() before it reads what the predecessor wrote (qd_kernels.cuh)
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, unsigned block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

template <int FRAMES, int THREADS>
int launch_frames_t(qd_ctx* c, const FramesParams& p, cudaStream_t s, int occ_slot, bool dependent) {
    constexpr bool DIRECT = FRAMES == 1;  // the one-frame kernel looks codons up straight from the letters
    auto kern = translate_frames_kernel<FRAMES, THREADS, DIRECT>;
    using Smem = FramesSmem<THREADS, frames_stages(FRAMES)>;
    const size_t smem = DIRECT ? ((sizeof(Smem) + 127) & ~(size_t)127) + DIRECT_ENTRIES : sizeof(Smem);
    if (c->occ_frames[occ_slot] == 0) {
        int occ = 0;
        QD_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (const char* e = getenv("QD_CARVEOUT")) QD_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e)));
        QD_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
        c->occ_frames[occ_slot] = std::max(1, std::min(occ, TileGeom<THREADS>::CTAS_PER_SM));
    }
    if (p.n_tiles == 0) return QD_OK;
    const unsigned grid = (unsigned)std::min<uint64_t>(p.n_tiles, (uint64_t)c->sm_count * c->occ_frames[occ_slot]);
    if (dependent && c->pdl) QD_CUDA(c, launch_pdl(kern, grid, THREADS, smem, s, p));
    else kern<<<grid, THREADS, smem, s>>>(p);
    c->launches++;
    QD_CUDA(c, cudaGetLastError());
    return QD_OK;
}

// runs per tile of a launch = threads of its CTAs
static uint64_t tile_runs_of(int /*frames*/, bool big) { return big ? TileGeom<BIG_THREADS>::RUNS : TileGeom<SMALL_THREADS>::RUNS; }

// p.n_tiles must match the tile size (tile_runs_of)
// (two 512-thread CTAs per SM were measured for variable-length batches too: no better than one 1024-thread CTA)
// dependent: launched behind kernels that trigger early (programmatic dependent launch, see launch_pdl)
int launch_frames(qd_ctx* c, int frames, const FramesParams& p, cudaStream_t s, bool big, bool dependent = false) {
    if (big) {
        switch (frames) {
            case QD_FRAMES_ONE: return launch_frames_t<1, BIG_THREADS>(c, p, s, 3, dependent);
            case QD_FRAMES_SELF: return launch_frames_t<3, BIG_THREADS>(c, p, s, 4, dependent);
            default: return launch_frames_t<6, BIG_THREADS>(c, p, s, 5, dependent);
        }
    }
    switch (frames) {
        case QD_FRAMES_ONE: return launch_frames_t<1, SMALL_THREADS>(c, p, s, 0, dependent);
        case QD_FRAMES_SELF: return launch_frames_t<3, SMALL_THREADS>(c, p, s, 1, dependent);
        default: return launch_frames_t<6, SMALL_THREADS>(c, p, s, 2, dependent);
    }
}

// uniform batch (also the single-sequence case n_seq = 1) on device buffers
int frames_uniform_dev(qd_ctx* c, int slot, int enc, int strict, int frames, const uint8_t* dna_dev, size_t n_seq,
                       size_t seq_len, uint8_t* out_dev, cudaStream_t s, uint64_t key_base = 0, unsigned long long* err_key = nullptr,
                       bool dependent = false) {
    if (n_seq == 0 || seq_len == 0) return QD_OK;
    const uint64_t runs = (seq_len + RUN_NT - 1) / RUN_NT;
    const uint64_t total_runs = runs * n_seq;
    if (runs > 0xffffffffull || total_runs > 0xffffff00ull) {
        c->last_error = "batch too large for one launch (more than 2^32 runs)";
        return QD_ERR_INVALID_ARGUMENT;
    }
    FramesParams p{};
    p.dna = dna_dev;
    p.total_nt = (uint64_t)n_seq * seq_len;
    p.offsets = nullptr;
    p.run_prefix = nullptr;
    p.tiles = nullptr;
    p.n_seq = n_seq;
    p.uniform_len = seq_len;
    p.uniform_runs = (uint32_t)runs;
    p.uniform_magic = runs == 1 ? 0 : (~0ull / runs) + 1;  // ceil(2^64 / runs); the kernel special-cases runs == 1
    p.uniform_mod3 = (uint32_t)(seq_len % 3);
    p.total_runs = total_runs;
    const bool big = use_big_tiles(c, total_runs);
    const uint64_t tile_runs = tile_runs_of(frames, big);
    p.n_tiles = (uint32_t)((total_runs + tile_runs - 1) / tile_runs);
    // short sequences: tiles of WHOLE sequences when that idles at most 1/16 of a tile's threads (32-bit geometry, no quotient)
    const uint64_t k = tile_runs / runs;
    if (c->aligned_tiles && k >= 1 && k * runs * 16 >= tile_runs * 15 && n_seq <= 0xffffffffull) {
        p.seqs_per_tile = (uint32_t)k;
        p.tile_magic = (uint32_t)(((1u << 20) + runs - 1) / runs);
        p.n_tiles = (uint32_t)((n_seq + k - 1) / k);
    }
    p.out = out_dev;
    p.lut = c->d_lut + (size_t)slot * LUT_ENTRIES;
    set_encoding(c, p, enc, strict);
    if (frames == QD_FRAMES_ONE)
        if (int rc = direct_table(c, slot, enc, strict, s, &p.direct)) return rc;
    p.err_key = err_key ? err_key : c->d_err;
    p.key_base = key_base;
    return launch_frames(c, frames, p, s, big, dependent);
}

template <int MODE, int OP = 0>
int device_exclusive_scan(qd_ctx* c, const uint64_t* src, uint64_t n, uint64_t* dst, DevBuf& block_sums, cudaStream_t s,
                          const unsigned long long* gate = nullptr, unsigned long long gate_gen = 0, bool dependent = false, unsigned batch = 1,
                          uint64_t batch_stride = 0) {
    // batch > 1: `batch` arrays of n values, array y at src / dst + y * batch_stride, scanned by the same three launches
    const uint64_t n_blocks = std::max<uint64_t>(1, (n + SCAN_BLOCK - 1) / SCAN_BLOCK);
    QD_CUDA(c, block_sums.reserve(batch * n_blocks * sizeof(uint64_t)));
    const dim3 grid((unsigned)n_blocks, batch);
    uint64_t* const sums = block_sums.as<uint64_t>();
    // dependent: the kernel before this scan in the stream triggers early (every kernel of the scan waits for its predecessor)
    if (dependent && c->pdl) QD_CUDA(c, launch_pdl(scan_block_sums_kernel<MODE, OP>, grid, SCAN_THREADS, 0, s, src, n, sums, gate, gate_gen, batch_stride));
    else scan_block_sums_kernel<MODE, OP><<<grid, SCAN_THREADS, 0, s>>>(src, n, sums, gate, gate_gen, batch_stride);
    if (c->pdl) {
        QD_CUDA(c, launch_pdl(scan_spine_kernel<OP>, dim3(batch), SCAN_THREADS, 0, s, sums, n_blocks, gate, gate_gen));
        QD_CUDA(c, launch_pdl(scan_finish_kernel<MODE, OP>, grid, SCAN_THREADS, 0, s, src, n, (const uint64_t*)sums, dst, gate, gate_gen, batch_stride));
    } else {
        scan_spine_kernel<OP><<<batch, SCAN_THREADS, 0, s>>>(sums, n_blocks, gate, gate_gen);
        scan_finish_kernel<MODE, OP><<<grid, SCAN_THREADS, 0, s>>>(src, n, sums, dst, gate, gate_gen, batch_stride);
    }
    c->launches += 3;
    QD_CUDA(c, cudaGetLastError());
    return QD_OK;
}

uint64_t runs_upper_bound(uint64_t total_nt, uint64_t n_seq) { return total_nt / RUN_NT + n_seq; }

// variable-length batch on device buffers; offsets_dev are absolute, `flat_base` = offsets[0]
int frames_csr_dev(qd_ctx* c, int slot, int enc, int strict, int frames, const uint8_t* dna_dev, uint64_t total_nt,
                   const uint64_t* offsets_dev, size_t n_seq, DevBuf& run_prefix, DevBuf& tile_first, DevBuf& block_sums,
                   uint8_t* out_dev, cudaStream_t s, uint64_t flat_base = 0, uint64_t key_base = 0, unsigned long long* err_key = nullptr) {
    if (n_seq == 0) return QD_OK;
    if (n_seq >= 0xffffffffull || total_nt / RUN_NT + n_seq > 0xffffff00ull) {
        c->last_error = "batch too large for one launch";
        return QD_ERR_INVALID_ARGUMENT;
    }
    // the exact run count is only known on the device; the tile size is picked from its upper bound
    const bool big = use_big_tiles(c, total_nt / RUN_NT);
    const uint64_t tile_runs = tile_runs_of(frames, big);
    const uint32_t wq_stride = big ? TileGeom<BIG_THREADS>::WQ_STRIDE : TileGeom<SMALL_THREADS>::WQ_STRIDE;
    const size_t n_tiles = (size_t)((runs_upper_bound(total_nt, n_seq) + tile_runs - 1) / tile_runs);
    // The frame kernel bulk-copies windows of run_prefix / offsets (TMA: 16-byte aligned sources, reads up to a window
    // past the last entry), so both live in one own buffer: [run_prefix | pad of ~0][copy of the offsets | pad]
    constexpr size_t PAD = BIG_THREADS + 8;
    const size_t arr_entries = (n_seq + 1 + PAD + 1) & ~(size_t)1;  // even: the second array stays 16-byte aligned
    QD_CUDA(c, run_prefix.reserve(2 * arr_entries * sizeof(uint64_t)));
    uint64_t* const rp_dev = run_prefix.as<uint64_t>();
    uint64_t* const off_pad = rp_dev + arr_entries;
    // one buffer: the tile map, then the 32-run map (padded by one tile's worth of entries), then the lists of long sequences
    const size_t warp_map_off = ((n_tiles + 1) * sizeof(TileInfo) + 127) & ~(size_t)127;
    const uint64_t n_q = (uint64_t)(n_tiles + 1) * wq_stride;  // a slice of wq_stride entries per tile
    const size_t long_off = (warp_map_off + (n_q + 8) * sizeof(uint32_t) + 15) & ~(size_t)15;
    const size_t mid_cap = (size_t)(runs_upper_bound(total_nt, n_seq) / CSR_THREAD_RUNS) + 1;
    const size_t long_cap = (size_t)(runs_upper_bound(total_nt, n_seq) / CSR_WARP_RUNS) + 1;
    QD_CUDA(c, tile_first.reserve(long_off + 16 + (mid_cap + long_cap) * sizeof(uint32_t)));
    uint32_t* warp_first = reinterpret_cast<uint32_t*>(tile_first.as<uint8_t>() + warp_map_off);
    // scan of the run counts; its last pass writes the prefix, the padded copy of the offsets and both maps (csr_finish_kernel)
    CsrMapParams m{};
    m.offsets = offsets_dev;
    m.n_seq = n_seq;
    m.flat_base = flat_base;
    m.total_nt = total_nt;
    m.err_key = err_key ? err_key : c->d_err;
    m.run_prefix = rp_dev;
    m.off_copy = off_pad;
    m.arr_entries = arr_entries;
    m.tiles = tile_first.as<TileInfo>();
    m.n_tiles_cap = n_tiles;
    m.tile_shift = 0;
    while ((uint64_t(1) << m.tile_shift) < tile_runs) m.tile_shift++;
    if ((uint64_t(1) << m.tile_shift) != tile_runs) {
        c->last_error = "tile size is not a power of two";
        return QD_ERR_INVALID_ARGUMENT;
    }
    m.wq_stride = wq_stride;
    m.warp_first = warp_first;
    m.list_counts = reinterpret_cast<unsigned int*>(tile_first.as<uint8_t>() + long_off);
    m.mid_list = reinterpret_cast<uint32_t*>(tile_first.as<uint8_t>() + long_off + 16);
    m.long_list = m.mid_list + mid_cap;
    QD_CUDA(c, cudaMemsetAsync(m.list_counts, 0, 2 * sizeof(unsigned int), s));
    const uint64_t n_blocks = std::max<uint64_t>(1, (n_seq + SCAN_BLOCK - 1) / SCAN_BLOCK);
    QD_CUDA(c, block_sums.reserve(n_blocks * sizeof(uint64_t)));
    csr_block_sums_kernel<<<(unsigned)n_blocks, SCAN_THREADS, 0, s>>>(m, block_sums.as<uint64_t>());
    const unsigned g_listed = (unsigned)std::min<uint64_t>((mid_cap + 7) / 8 + long_cap, (uint64_t)c->sm_count * 4);
    if (c->pdl) {  // each kernel of the chain takes its place while the one before drains, and waits for it before it reads
        QD_CUDA(c, launch_pdl(scan_spine_kernel<0>, 1, SCAN_THREADS, 0, s, block_sums.as<uint64_t>(), n_blocks, nullptr, 0ull));
        QD_CUDA(c, launch_pdl(csr_finish_kernel, (unsigned)n_blocks, SCAN_THREADS, 0, s, m, block_sums.as<uint64_t>()));
        QD_CUDA(c, launch_pdl(csr_listed_kernel, g_listed, 256, 0, s, m));
    } else {
        scan_spine_kernel<0><<<1, SCAN_THREADS, 0, s>>>(block_sums.as<uint64_t>(), n_blocks, nullptr, 0);
        csr_finish_kernel<<<(unsigned)n_blocks, SCAN_THREADS, 0, s>>>(m, block_sums.as<uint64_t>());
        csr_listed_kernel<<<g_listed, 256, 0, s>>>(m);
    }
    c->launches += 4;
    QD_CUDA(c, cudaGetLastError());
    FramesParams p{};
    p.dna = dna_dev;
    p.total_nt = total_nt;
    p.offsets = off_pad;
    p.run_prefix = rp_dev;
    p.tiles = tile_first.as<TileInfo>();
    p.warp_first = warp_first;
    p.n_seq = n_seq;
    p.n_tiles = (uint32_t)n_tiles;
    p.out = out_dev;
    p.lut = c->d_lut + (size_t)slot * LUT_ENTRIES;
    set_encoding(c, p, enc, strict);
    if (frames == QD_FRAMES_ONE)
        if (int rc = direct_table(c, slot, enc, strict, s, &p.direct)) return rc;
    p.err_key = err_key ? err_key : c->d_err;
    p.flat_base = flat_base;
    p.key_base = key_base;
    return launch_frames(c, frames, p, s, big, /*dependent=*/true);
}

template <int THREADS>
int launch_revcomp_t(qd_ctx* c, const RevcompParams& p, cudaStream_t s, int occ_slot) {
    auto kern = revcomp_kernel<THREADS>;
    const size_t smem = sizeof(RevcompSmem<THREADS>);
    if (c->occ_revcomp[occ_slot] == 0) {
        int occ = 0;
        QD_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        QD_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
        c->occ_revcomp[occ_slot] = std::max(1, occ);
    }
    const uint64_t tiles = ((p.n >> 4) + RevcompSmem<THREADS>::TILE_CHUNKS - 1) / RevcompSmem<THREADS>::TILE_CHUNKS;
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(tiles, (uint64_t)c->sm_count * c->occ_revcomp[occ_slot]));
    kern<<<grid, THREADS, smem, s>>>(p);
    c->launches++;
    QD_CUDA(c, cudaGetLastError());
    return QD_OK;
}

// `which` = one of RC_PAIR_*: selects the complement table (256-byte form and letter-pair form)
int revcomp_dev_table(qd_ctx* c, int which, const uint8_t* dna_dev, size_t n, uint8_t* out_dev, cudaStream_t s,
                      unsigned long long* err_key = nullptr, uint64_t key_base = 0) {
    if (n == 0) return QD_OK;
    static const int byte_tab[RC_PAIR_COUNT] = {TAB_RC_ASCII, TAB_RC_ASCII_STRICT, TAB_RC_MASK, TAB_RC_MASK_STRICT,
                                                TAB_RC_MASK_TO_ASCII, TAB_RC_MASK_TO_ASCII_STRICT};
    const bool ascii_in = which < RC_PAIR_MASK;
    RevcompParams p{};
    p.dna = dna_dev;
    p.n = n;
    p.out = out_dev;
    p.table = tab(c, byte_tab[which]);
    p.pair = c->d_rc_pair + (size_t)which * PAIR_ENTRIES;
    p.pair_coef = 2u | ((2u * (ascii_in ? PAIR_K1_ASCII : PAIR_K1_MASK)) << 16);
    p.and_need = ascii_in ? 0x40404040u : 0u;
    p.or_forbid = ascii_in ? 0x80808080u : 0xe0e0e0e0u;
    p.filler = ascii_in ? 0x41u : 0x01u;
    p.err_key = err_key ? err_key : c->d_err;
    p.key_base = key_base;
    // big tiles once every SM gets two of them
    const bool big = (n >> 4) >= (uint64_t)2 * RevcompSmem<RC_BIG_THREADS>::TILE_CHUNKS * (uint64_t)c->sm_count && c->big_tile_min_runs != ~0ull;
    if (big || c->big_tile_min_runs == 0) return launch_revcomp_t<RC_BIG_THREADS>(c, p, s, 1);
    return launch_revcomp_t<RC_SMALL_THREADS>(c, p, s, 0);
}

int reset_err_key(qd_ctx* c, cudaStream_t s) {
    QD_CUDA(c, cudaMemsetAsync(c->d_err, 0xff, sizeof(unsigned long long), s));
    return QD_OK;
}

// returns QD_OK and *key (NO_ERROR_KEY when clean); synchronises `s`
int fetch_err_key(qd_ctx* c, cudaStream_t s, unsigned long long* key) {
    QD_CUDA(c, cudaMemcpyAsync(c->h_err, c->d_err, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    QD_CUDA(c, cudaStreamSynchronize(s));
    *key = *c->h_err;
    return QD_OK;
}

int report_single(int enc, int strict, const uint8_t* dna_host, unsigned long long key, qd_err* err) {
    uint8_t payload = 0;
    const int kind = classify_byte(enc, strict, dna_host[key], &payload);
    return set_err(err, kind == QD_OK ? QD_ERR_BAD_NUCLEOTIDE : kind, payload, key, 0);
}

}  // namespace

// =================================================================================================
// context
// =================================================================================================
extern "C" int qd_abi_version(void) { return QD_ABI_VERSION; }

extern "C" int qd_ctx_create(qd_ctx** out, int device) {
    if (!out) return QD_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) return QD_ERR_CUDA;
    int prev_dev = -1;
    if (cudaGetDevice(&prev_dev) != cudaSuccess) prev_dev = -1;
    struct Restore {
        int prev, dev;
        ~Restore() { if (prev >= 0 && prev != dev) cudaSetDevice(prev); }
    } restore{prev_dev, device};
    if (cudaSetDevice(device) != cudaSuccess) return QD_ERR_CUDA;
    qd_ctx* c = new qd_ctx();
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete c;
        return QD_ERR_CUDA;
    }
    c->sm_count = prop.multiProcessorCount;
    c->big_tile_min_runs = (uint64_t)2 * BIG_THREADS * (uint64_t)c->sm_count;  // every SM gets two big tiles
    // tuning knobs of the canonical tile kernel, read once per context (not on the launch path)
    if (const char* e = getenv("QD_ALIGNED_TILES")) c->aligned_tiles = atoi(e) != 0;
    if (const char* e = getenv("QD_PDL")) c->pdl = atoi(e) != 0;
    if (const char* e = getenv("QD_CTILE_WPT_KEYS")) c->ctile_wpt[0] = (uint32_t)std::max(64, atoi(e) & ~63);
    if (const char* e = getenv("QD_CTILE_WPT_ROWS")) c->ctile_wpt[1] = (uint32_t)std::max(64, atoi(e) & ~63);
    const HostTables& T = host_tables();
    bool ok = true;
    ok &= cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < N_PIPE; i++) ok &= cudaStreamCreateWithFlags(&c->pipe[i].stream, cudaStreamNonBlocking) == cudaSuccess;
    ok &= cudaMalloc(&c->d_lut, T.folded.size() * sizeof(uint16_t)) == cudaSuccess;
    ok &= cudaMalloc(&c->d_tables, 256 * TAB_COUNT) == cudaSuccess;
    ok &= cudaMalloc(&c->d_pair, sizeof(T.pair)) == cudaSuccess;
    ok &= cudaMalloc(&c->d_rc_pair, sizeof(T.rc_pair)) == cudaSuccess;
    ok &= cudaMalloc(&c->d_ckey_pair, sizeof(T.ckey_pair)) == cudaSuccess;
    ok &= cudaMalloc(&c->d_ckey, sizeof(T.ckey)) == cudaSuccess;
    ok &= cudaMalloc(&c->d_parse_pair, sizeof(T.parse_pair)) == cudaSuccess;
    ok &= cudaMalloc(&c->d_gate, 8) == cudaSuccess && cudaMemset(c->d_gate, 0, 8) == cudaSuccess;
    ok &= cudaMalloc(&c->d_err, sizeof(unsigned long long)) == cudaSuccess;
    ok &= cudaEventCreateWithFlags(&c->dev_done, cudaEventDisableTiming) == cudaSuccess;
    ok &= cudaHostAlloc(&c->h_err, sizeof(unsigned long long), cudaHostAllocDefault) == cudaSuccess;
    ok &= cudaHostAlloc(&c->h_scratch, 8 * sizeof(uint64_t), cudaHostAllocDefault) == cudaSuccess;
    ok &= cudaHostAlloc(&c->h_stage, SMALL_CALL_BYTES + 64, cudaHostAllocDefault) == cudaSuccess;
    ok &= cudaHostAlloc(&c->h_tiny, TINY_IN_BYTES + 64 + TINY_OUT_BYTES + 64, cudaHostAllocMapped) == cudaSuccess;
    if (const char* e = getenv("QD_TINY_CALLS")) c->tiny_calls = atoi(e) != 0;
    if (ok) {
        std::vector<uint8_t> tabs(256 * TAB_COUNT);
        auto put = [&](int which, const uint8_t* src) { memcpy(tabs.data() + 256 * which, src, 256); };
        put(TAB_DECODE_ASCII, T.decode_ascii);
        put(TAB_DECODE_MASK, T.decode_mask);
        put(TAB_RC_ASCII, T.rc_ascii[0]);
        put(TAB_RC_ASCII_STRICT, T.rc_ascii[1]);
        put(TAB_RC_MASK, T.rc_mask[0]);
        put(TAB_RC_MASK_STRICT, T.rc_mask[1]);
        put(TAB_RC_MASK_TO_ASCII, T.rc_mask_to_ascii[0]);
        put(TAB_RC_MASK_TO_ASCII_STRICT, T.rc_mask_to_ascii[1]);
        put(TAB_UPPER_ASCII, T.upper_ascii[0]);
        put(TAB_UPPER_ASCII_STRICT, T.upper_ascii[1]);
        put(TAB_MASK_TO_ASCII, T.mask_to_ascii[0]);
        put(TAB_MASK_TO_ASCII_STRICT, T.mask_to_ascii[1]);
        put(TAB_TO_MASK_ASCII, T.to_mask_ascii);
        put(TAB_TO_MASK_MASK, T.to_mask_mask);
        put(TAB_PARSE, T.parse[0]);
        put(TAB_PARSE_STRICT, T.parse[1]);
        ok &= cudaMemcpy(c->d_tables, tabs.data(), tabs.size(), cudaMemcpyHostToDevice) == cudaSuccess;
        ok &= cudaMemcpy(c->d_lut, T.folded.data(), T.folded.size() * sizeof(uint16_t), cudaMemcpyHostToDevice) == cudaSuccess;
        ok &= cudaMemcpy(c->d_pair, T.pair, sizeof(T.pair), cudaMemcpyHostToDevice) == cudaSuccess;
        ok &= cudaMemcpy(c->d_rc_pair, T.rc_pair, sizeof(T.rc_pair), cudaMemcpyHostToDevice) == cudaSuccess;
        ok &= cudaMemcpy(c->d_ckey_pair, T.ckey_pair, sizeof(T.ckey_pair), cudaMemcpyHostToDevice) == cudaSuccess;
        ok &= cudaMemcpy(c->d_ckey, &T.ckey, sizeof(T.ckey), cudaMemcpyHostToDevice) == cudaSuccess;
        ok &= cudaMemcpy(c->d_parse_pair, T.parse_pair, sizeof(T.parse_pair), cudaMemcpyHostToDevice) == cudaSuccess;
        ok &= cudaMemset(c->d_err, 0xff, sizeof(unsigned long long)) == cudaSuccess;
    }
    if (!ok) {
        qd_ctx_destroy(c);
        return QD_ERR_CUDA;
    }
    *out = c;
    return QD_OK;
}

extern "C" void qd_ctx_destroy(qd_ctx* c) {
    if (!c) return;
    int prev = -1;
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (DevBuf* b : {&c->in, &c->out, &c->aux0, &c->aux1, &c->aux2, &c->aux3, &c->run_prefix, &c->tile_first, &c->block_sums, &c->parse_counts,
                      &c->parse_offsets, &c->canon_state, &c->fasta_scratch}) b->release();
    for (int i = 0; i < N_PIPE; i++) {
        Pipe& P = c->pipe[i];
        for (DevBuf* b : {&P.in, &P.out, &P.offsets, &P.run_prefix, &P.tile_first, &P.block_sums}) b->release();
        P.hin.release();
        P.hout.release();
        if (P.done) cudaEventDestroy(P.done);
        if (P.stream) cudaStreamDestroy(P.stream);
    }
    delete c->pool;
    if (c->d_lut) cudaFree(c->d_lut);
    if (c->d_tables) cudaFree(c->d_tables);
    if (c->d_pair) cudaFree(c->d_pair);
    if (c->d_rc_pair) cudaFree(c->d_rc_pair);
    if (c->d_ckey_pair) cudaFree(c->d_ckey_pair);
    if (c->d_ckey) cudaFree(c->d_ckey);
    if (c->d_parse_pair) cudaFree(c->d_parse_pair);
    if (c->d_gate) cudaFree(c->d_gate);
    for (auto& kv : c->d_direct) cudaFree(kv.second);
    if (c->d_err) cudaFree(c->d_err);
    if (c->h_err) cudaFreeHost(c->h_err);
    if (c->h_scratch) cudaFreeHost(c->h_scratch);
    if (c->h_stage) cudaFreeHost(c->h_stage);
    if (c->h_tiny) cudaFreeHost(c->h_tiny);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->dev_done) cudaEventDestroy(c->dev_done);
    if (prev >= 0 && prev != c->device) cudaSetDevice(prev);
    delete c;
}

extern "C" int qd_ctx_device(const qd_ctx* c) { return c ? c->device : -1; }
extern "C" void* qd_ctx_stream(const qd_ctx* c) { return c ? (void*)c->stream : nullptr; }
extern "C" int qd_ctx_sm_count(const qd_ctx* c) { return c ? c->sm_count : 0; }
extern "C" const char* qd_last_cuda_error(const qd_ctx* c) { return c ? c->last_error.c_str() : "no context"; }
extern "C" uint64_t qd_launch_count(const qd_ctx* c) { return c ? c->launches : 0; }
extern "C" int qd_ctx_set_big_tile_min_runs(qd_ctx* c, uint64_t runs) {
    if (!c) return QD_ERR_INVALID_ARGUMENT;
    c->big_tile_min_runs = runs;
    return QD_OK;
}
extern "C" int qd_ctx_set_chunk_bytes(qd_ctx* c, size_t bytes) {
    if (!c || bytes < (size_t(1) << 16)) return QD_ERR_INVALID_ARGUMENT;
    c->chunk_bytes = bytes;
    return QD_OK;
}
extern "C" void* qd_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (bytes == 0 || cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
extern "C" void qd_host_free(void* p) {
    if (p) cudaFreeHost(p);
}
extern "C" int qd_ctx_set_stage_threads(qd_ctx* c, int threads) {
    if (!c || threads < -1 || threads > 64) return QD_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lk(c->mu);
    delete c->pool;
    c->pool = nullptr;
    c->stage_threads = threads;
    return QD_OK;
}

// =================================================================================================
// host-only helpers
// =================================================================================================
extern "C" int qd_table_slot(int id) {
    if (id == 8) return 0;  // alias of 1, src/trans_table.rs:118
    if (id == 7) return 3;  // alias of 4, src/trans_table.rs:122
    for (int s = 0; s < N_SLOTS; s++)
        if (SLOT_NCBI_ID[s] == id) return s;
    return -1;
}

extern "C" const uint8_t* qd_table_data(int slot) {
    if (slot < 0 || slot >= N_SLOTS) return nullptr;
    return host_tables().ref.data() + (size_t)slot * CODONS;
}

static int fmt_char_debug(uint8_t ch, char* buf, size_t cap) {  // Rust `{:?}` of an ASCII char
    switch (ch) {
        case 0: return snprintf(buf, cap, "'\\0'");
        case 9: return snprintf(buf, cap, "'\\t'");
        case 10: return snprintf(buf, cap, "'\\n'");
        case 13: return snprintf(buf, cap, "'\\r'");
        case 39: return snprintf(buf, cap, "'\\''");
        case 92: return snprintf(buf, cap, "'\\\\'");
        default: break;
    }
    if (ch < 32 || ch == 127) return snprintf(buf, cap, "'\\u{%x}'", (unsigned)ch);
    return snprintf(buf, cap, "'%c'", (char)ch);
}

extern "C" int qd_err_message(const qd_err* e, char* buf, size_t cap) {
    if (!e || !buf || !cap) return 0;
    char q[16];
    switch (e->kind) {
        case QD_OK: return snprintf(buf, cap, "ok");
        case QD_ERR_NON_ASCII_BYTE: return snprintf(buf, cap, "non-ascii byte: %x", (unsigned)e->byte);
        case QD_ERR_BAD_NUCLEOTIDE:
            fmt_char_debug(e->byte, q, sizeof q);
            return snprintf(buf, cap, "bad nucleotide: %s", q);
        case QD_ERR_UNEXPECTED_AMBIGUOUS:
            fmt_char_debug(e->byte, q, sizeof q);
            return snprintf(buf, cap, "unexpected ambiguous nucleotide: %s", q);
        case QD_ERR_BAD_TABLE: return snprintf(buf, cap, "not a ncbi translation table: %u", (unsigned)e->byte);
        case QD_ERR_BAD_MASK: return snprintf(buf, cap, "not a nucleotide mask: %u", (unsigned)e->byte);
        case QD_ERR_INVALID_ARGUMENT: return snprintf(buf, cap, "invalid argument");
        default: return snprintf(buf, cap, "cuda error");
    }
}

extern "C" size_t qd_frame_len(size_t n, int k) { return (k >= 0 && k < 3 && n > (size_t)k) ? (n - k) / 3 : 0; }
extern "C" size_t qd_runs(size_t n) { return (n + RUN_NT - 1) / RUN_NT; }
extern "C" size_t qd_slots_bytes(size_t n, int frames) { return 16 * qd_runs(n) * (size_t)frames; }
extern "C" size_t qd_slot_frame_offset(size_t n, int frames, int f) {
    const size_t slot = 16 * qd_runs(n);
    if (f < 3 || frames != QD_FRAMES_ALL) return (size_t)f * slot;
    return (size_t)(f + 1) * slot - qd_frame_len(n, f - 3);
}

extern "C" size_t qd_batch_out_bytes(const uint64_t* offsets, size_t n_seq, int frames) {
    size_t runs = 0;
    for (size_t i = 0; i < n_seq; i++) runs += qd_runs((size_t)(offsets[i + 1] - offsets[i]));
    return runs * 16 * (size_t)frames;
}

// =================================================================================================
// device-pointer entry points
// =================================================================================================
// device-pointer entry points run on exactly the stream they are given; NULL is the legacy default
// stream (what torch.cuda.current_stream() is unless the caller switched streams)
static cudaStream_t pick_stream(qd_ctx*, void* s) { return (cudaStream_t)s; }

extern "C" int qd_dev_error_reset(qd_ctx* c, void* stream) {
    if (!c) return QD_ERR_INVALID_ARGUMENT;
    Guard g(c, pick_stream(c, stream));
    if (!g.ok) return QD_ERR_CUDA;
    return reset_err_key(c, pick_stream(c, stream));
}

extern "C" int qd_dev_error_fetch(qd_ctx* c, void* stream, int enc, int strict, const uint8_t* dna_dev,
                                  const uint64_t* offsets_dev, size_t n_seq, size_t uniform_len, qd_err* err) {
    if (!c) return QD_ERR_INVALID_ARGUMENT;
    Guard g(c, pick_stream(c, stream));
    if (!g.ok) return QD_ERR_CUDA;
    clear_err(err);
    cudaStream_t s = pick_stream(c, stream);
    unsigned long long key = 0;
    int rc = fetch_err_key(c, s, &key);
    if (rc) return rc;
    if (key == NO_ERROR_KEY) return QD_OK;
    if (key & CSR_BAD_OFFSETS) {  // csr_finish_kernel: the offsets of this sequence decrease or leave the buffer
        c->last_error = "offsets of sequence " + std::to_string(key & ~CSR_BAD_OFFSETS) + " decrease or point outside the buffer";
        set_err(err, QD_ERR_INVALID_ARGUMENT, 0, 0, key & ~CSR_BAD_OFFSETS);
        return QD_ERR_INVALID_ARGUMENT;
    }
    uint8_t b = 0;
    QD_CUDA(c, cudaMemcpyAsync(c->h_scratch, dna_dev + key, 1, cudaMemcpyDeviceToHost, s));
    QD_CUDA(c, cudaStreamSynchronize(s));
    b = *reinterpret_cast<uint8_t*>(c->h_scratch);
    uint64_t seq = 0, pos = key;
    if (offsets_dev && n_seq) {
        // upper_bound over the device offsets by bisection with single-element copies (error path only)
        uint64_t lo = 0, hi = n_seq - 1;
        while (lo < hi) {
            uint64_t mid = (lo + hi + 1) >> 1, v = 0;
            QD_CUDA(c, cudaMemcpyAsync(c->h_scratch, offsets_dev + mid, 8, cudaMemcpyDeviceToHost, s));
            QD_CUDA(c, cudaStreamSynchronize(s));
            v = c->h_scratch[0];
            if (v <= key) lo = mid; else hi = mid - 1;
        }
        QD_CUDA(c, cudaMemcpyAsync(c->h_scratch, offsets_dev + lo, 8, cudaMemcpyDeviceToHost, s));
        QD_CUDA(c, cudaStreamSynchronize(s));
        seq = lo;
        pos = key - c->h_scratch[0];
    } else if (uniform_len) {
        seq = key / uniform_len;
        pos = key % uniform_len;
    }
    uint8_t payload = 0;
    int kind = classify_byte(enc, strict, b, &payload);
    if (kind == QD_OK) kind = QD_ERR_BAD_NUCLEOTIDE;
    return set_err(err, kind, payload, pos, seq);
}

extern "C" int qd_revcomp_dev(qd_ctx* c, int enc, int strict, const uint8_t* dna_dev, size_t n, uint8_t* out_dev, void* stream) {
    if (!c || (n && (!dna_dev || !out_dev)) || (reinterpret_cast<uintptr_t>(out_dev) & 15)) return QD_ERR_INVALID_ARGUMENT;
    Guard g(c, pick_stream(c, stream));
    if (!g.ok) return QD_ERR_CUDA;
    const int which = enc == QD_ENC_ASCII ? (strict ? RC_PAIR_ASCII_STRICT : RC_PAIR_ASCII) : (strict ? RC_PAIR_MASK_STRICT : RC_PAIR_MASK);
    return revcomp_dev_table(c, which, dna_dev, n, out_dev, pick_stream(c, stream));
}

extern "C" int qd_translate_frames_dev(qd_ctx* c, int table, int enc, int strict, int frames, const uint8_t* dna_dev, size_t n,
                                       uint8_t* out_dev, void* stream) {
    return qd_translate_frames_batch_uniform_dev(c, table, enc, strict, frames, dna_dev, n ? 1 : 0, n, out_dev, stream);
}

extern "C" int qd_translate_frames_batch_uniform_dev(qd_ctx* c, int table, int enc, int strict, int frames, const uint8_t* dna_dev,
                                                     size_t n_seq, size_t seq_len, uint8_t* out_dev, void* stream) {
    if (!c || !valid_frames(frames) || (reinterpret_cast<uintptr_t>(out_dev) & 15)) return QD_ERR_INVALID_ARGUMENT;
    int slot;
    if (check_table(table, nullptr, &slot)) return QD_ERR_BAD_TABLE;
    Guard g(c, pick_stream(c, stream));
    if (!g.ok) return QD_ERR_CUDA;
    return frames_uniform_dev(c, slot, enc, strict, frames, dna_dev, n_seq, seq_len, out_dev, pick_stream(c, stream));
}

extern "C" int qd_translate_frames_batch_dev(qd_ctx* c, int table, int enc, int strict, int frames, const uint8_t* dna_dev,
                                             const uint64_t* offsets_dev, size_t n_seq, size_t total_nt, uint8_t* out_dev,
                                             void* stream) {
    if (!c || !valid_frames(frames) || (reinterpret_cast<uintptr_t>(out_dev) & 15)) return QD_ERR_INVALID_ARGUMENT;
    int slot;
    if (check_table(table, nullptr, &slot)) return QD_ERR_BAD_TABLE;
    Guard g(c, pick_stream(c, stream));
    if (!g.ok) return QD_ERR_CUDA;
    return frames_csr_dev(c, slot, enc, strict, frames, dna_dev, total_nt, offsets_dev, n_seq, c->run_prefix, c->tile_first,
                          c->block_sums, out_dev, pick_stream(c, stream));
}

static CanonParams canonical_params(qd_ctx* c, int enc, int forward_only, const uint8_t* dna_dev, size_t n_seq, size_t seq_len, size_t w,
                                    uint8_t* out_dev) {
    CanonParams p{};
    p.dna = dna_dev;
    p.n_seq = n_seq;
    p.seq_len = seq_len;
    p.w = (uint32_t)w;
    p.forward_only = forward_only;
    p.out = out_dev;
    const bool ascii = enc == QD_ENC_ASCII;
    p.decode = tab(c, ascii ? TAB_DECODE_ASCII : TAB_DECODE_MASK);
    p.pair_coef = 1u | ((ascii ? PAIR_K1_ASCII : PAIR_K1_MASK) << 16);
    p.and_need = ascii ? 0x40404040u : 0u;
    p.or_forbid = ascii ? 0x80808080u : 0xe0e0e0e0u;
    p.filler = ascii ? 0x41u : 0x01u;
    p.letters = ascii ? ('A' | ('T' << 8) | ('C' << 16) | ((uint32_t)'G' << 24)) : (1u | (2u << 8) | (4u << 16) | (8u << 24));
    p.err_key = c->cur_err_key ? c->cur_err_key : c->d_err;
    return p;
}

// One launch of the tile kernel: rows (MODE = CTILE_ROWS) or keys.  The warp-tile is the largest multiple of 64 windows
// whose positions (and, for rows, staged bytes) fit the per-warp shared-memory budget.
template <int NW, int W, int MODE>
static int launch_canonical_tiles(qd_ctx* c, CanonKeyParams kp, cudaStream_t s) {
    auto kern = canonical_tile_kernel<NW, W, MODE>;
    const CanonParams& p = kp.base;
    const bool rows = MODE == CTILE_ROWS;
    const uint64_t nw = p.seq_len - p.w + 1;
    // 448 + 41 + slack = 16 x 32 positions for w = 42 is two full rounds of the per-position pass (192 and 704 fill one and
    // three; anything between wastes lanes of the last round).  Rows are staged and copied out 64 at a time whatever the tile,
    // so the larger tile only spreads the halo (w - 1 + look-ahead positions decoded per tile) over more windows: measured
    // on 768 x 99 999 nt, rows: 192 -> 0.586, 256 -> 0.563, 448 -> 0.629, 704 -> 0.633, 1024 -> 0.606 of the copy peak.
    uint32_t wpt = c->ctile_wpt[rows ? 1 : 0];
    while (wpt > 64 && ckey_span_bound(wpt, p.w, nw) > 2048) wpt -= 64;
    kp.wpt = wpt;
    kp.pos_cap = (uint32_t)((ckey_span_bound(wpt, p.w, nw) + 31) & ~(uint64_t)31);
    const size_t smem = CKEY_SHARED_TABLE_BYTES + (size_t)CanonTileLayout(kp.pos_cap, wpt, p.w, rows).total_bytes * CTILE_WARPS;
    // occupancy once per (instantiation, shared-memory size); the kernel's dynamic shared-memory limit only ever grows
    const uint64_t kern_key = ((uint64_t)(NW * 4 + MODE) << 40) | ((uint64_t)(W != 0) << 39);
    int& smem_limit = c->occ_cache[kern_key | ((uint64_t)1 << 38)];
    if ((int)smem > smem_limit) {
        QD_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_limit = (int)smem;
    }
    auto occ_it = c->occ_cache.find(kern_key | (uint64_t)smem);
    if (occ_it == c->occ_cache.end()) {
        int o = 0;
        QD_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, CTILE_THREADS, smem));
        occ_it = c->occ_cache.emplace(kern_key | (uint64_t)smem, o).first;
    }
    const int occ = occ_it->second;
    const uint64_t tiles = (p.n_seq * nw + (uint64_t)wpt * CTILE_WARPS - 1) / ((uint64_t)wpt * CTILE_WARPS);
    const unsigned grid = (unsigned)std::min<uint64_t>(tiles, (uint64_t)c->sm_count * std::max(1, occ));
    kern<<<grid, CTILE_THREADS, smem, s>>>(kp);
    c->launches++;
    QD_CUDA(c, cudaGetLastError());
    return QD_OK;
}

template <int MODE>
static int canonical_tiles_mode(qd_ctx* c, const CanonKeyParams& kp, cudaStream_t s) {
    if (kp.base.w == 42) return launch_canonical_tiles<3, 42, MODE>(c, kp, s);  // the window length of benches/canonical.rs: unrolled
    switch ((kp.base.w + 15) / 16) {
        case 1: return launch_canonical_tiles<1, 0, MODE>(c, kp, s);
        case 2: return launch_canonical_tiles<2, 0, MODE>(c, kp, s);
        case 3: return launch_canonical_tiles<3, 0, MODE>(c, kp, s);
        default: return launch_canonical_tiles<4, 0, MODE>(c, kp, s);
    }
}

static CanonKeyParams canonical_tile_params(qd_ctx* c, int enc, int forward_only, const uint8_t* dna_dev, size_t n_seq, size_t seq_len, size_t w,
                                            void* out_dev) {
    CanonKeyParams kp{};
    kp.base = canonical_params(c, enc, forward_only, dna_dev, n_seq, seq_len, w, static_cast<uint8_t*>(out_dev));
    kp.base.pair = c->d_ckey_pair + (enc == QD_ENC_ASCII ? 0 : PAIR_ENTRIES);
    kp.tables = c->d_ckey;
    return kp;
}

// every length-w window of each of n_seq sequences of seq_len nucleotides (n_seq = 1: one sequence)
static int canonical_dev(qd_ctx* c, int enc, int forward_only, const uint8_t* dna_dev, size_t n_seq, size_t seq_len, size_t w,
                         uint8_t* out_dev, cudaStream_t s) {
    if (w < 1 || w > CANON_MAX_W) return QD_ERR_INVALID_ARGUMENT;
    if (seq_len < w || n_seq == 0) return QD_OK;
    return canonical_tiles_mode<CTILE_ROWS>(c, canonical_tile_params(c, enc, forward_only, dna_dev, n_seq, seq_len, w, out_dev), s);
}

// one fixed-width key per canonical window
static int canonical_keys_dev(qd_ctx* c, int enc, int forward_only, int kind, const uint8_t* dna_dev, size_t n_seq, size_t seq_len, size_t w,
                              void* keys_dev, cudaStream_t s) {
    if (w < 1 || w > CANON_MAX_W || (kind != QD_CANON_KEY_PACKED && kind != QD_CANON_KEY_FP64)) return QD_ERR_INVALID_ARGUMENT;
    if (seq_len < w || n_seq == 0) return QD_OK;
    const CanonKeyParams kp = canonical_tile_params(c, enc, forward_only, dna_dev, n_seq, seq_len, w, keys_dev);
    return kind == QD_CANON_KEY_FP64 ? canonical_tiles_mode<CTILE_KEY_FP64>(c, kp, s) : canonical_tiles_mode<CTILE_KEY_PACKED>(c, kp, s);
}

extern "C" size_t qd_canonical_key_bytes(int kind, size_t w) {
    if (w < 1 || w > CANON_MAX_W) return 0;
    return kind == QD_CANON_KEY_FP64 ? 8 : (kind == QD_CANON_KEY_PACKED ? 8 * ((w + 31) / 32) : 0);
}

extern "C" uint64_t qd_canonical_fingerprint(const uint32_t* packed, size_t w) {
    if (!packed || w < 1 || w > CANON_MAX_W) return 0;
    if (w <= 32) {  // the same template the kernel instantiates
        const uint32_t lo[1] = {packed[0]}, hi[1] = {packed[1]};
        return canon_fingerprint<1>(lo, hi, (uint32_t)w);
    }
    const uint32_t lo[2] = {packed[0], packed[1]}, hi[2] = {packed[2], packed[3]};
    return canon_fingerprint<2>(lo, hi, (uint32_t)w);
}

extern "C" int qd_canonical_window_keys_dev(qd_ctx* c, int enc, int forward_only, int kind, const uint8_t* dna_dev, size_t n_seq,
                                            size_t seq_len, size_t w, void* keys_dev, void* stream) {
    if (!c || (reinterpret_cast<uintptr_t>(keys_dev) & 15)) return QD_ERR_INVALID_ARGUMENT;
    Guard g(c, pick_stream(c, stream));
    if (!g.ok) return QD_ERR_CUDA;
    return canonical_keys_dev(c, enc, forward_only, kind, dna_dev, n_seq, seq_len, w, keys_dev, pick_stream(c, stream));
}

extern "C" int qd_canonical_windows_batch_dev(qd_ctx* c, int enc, int forward_only, const uint8_t* dna_dev, size_t n_seq, size_t seq_len,
                                              size_t w, uint8_t* out_dev, void* stream) {
    if (!c || (reinterpret_cast<uintptr_t>(out_dev) & 15)) return QD_ERR_INVALID_ARGUMENT;
    Guard g(c, pick_stream(c, stream));
    if (!g.ok) return QD_ERR_CUDA;
    return canonical_dev(c, enc, forward_only, dna_dev, n_seq, seq_len, w, out_dev, pick_stream(c, stream));
}

extern "C" int qd_canonical_windows_dev(qd_ctx* c, int enc, int forward_only, const uint8_t* dna_dev, size_t n, size_t w,
                                        uint8_t* out_dev, void* stream) {
    if (!c || (reinterpret_cast<uintptr_t>(out_dev) & 15)) return QD_ERR_INVALID_ARGUMENT;
    Guard g(c, pick_stream(c, stream));
    if (!g.ok) return QD_ERR_CUDA;
    return canonical_dev(c, enc, forward_only, dna_dev, 1, n, w, out_dev, pick_stream(c, stream));
}

// ---- pageable caller buffers: staging through pinned memory (see HostPool) --------------------------------------------------
static bool host_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();  // (older drivers report unregistered memory as an error)
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}
static HostPool* stage_pool(qd_ctx* c) {
    if (!c->pool) {
        int t = c->stage_threads;
        if (t < 0) t = std::max(2, std::min(8, (int)std::thread::hardware_concurrency() / 2));
        try {
            c->pool = new HostPool(t);
        } catch (...) {  // no threads to be had: the calling thread copies alone (nothing may be thrown across the C ABI)
            c->pool = new HostPool(0);
        }
    }
    return c->pool;
}
static int pipe_event(qd_ctx* c, Pipe& P) {
    if (!P.done) QD_CUDA(c, cudaEventCreateWithFlags(&P.done, cudaEventDisableTiming));
    return QD_OK;
}
constexpr size_t STAGE_MIN_BYTES = size_t(32) << 20;  // smaller calls: the driver's own staging is as good

// One chunked host call.  Per chunk on pipe p:  begin(p) -- the chunk this pipe staged before has left the device, its
// output is copied to the caller --, h2d, kernels, d2h..., end(p).  finish() drains what is still staged.  With pinned
// caller buffers (or staging off) every call is the plain asynchronous copy.
struct Stager {
    struct Seg {
        uint8_t* dst;
        size_t off, n;
    };
    qd_ctx* c;
    bool in_staged = false, out_staged = false;
    std::vector<Seg> segs[N_PIPE];
    size_t out_used[N_PIPE] = {};
    bool busy[N_PIPE] = {};
    Stager(qd_ctx* c_, const void* in, const void* out, uint64_t total_bytes) : c(c_) {
        const bool on = c->stage_threads != 0 && total_bytes >= STAGE_MIN_BYTES;
        in_staged = on && in && host_pageable(in);
        out_staged = on && out && host_pageable(out);
    }
    bool active() const { return in_staged || out_staged; }
    size_t chunk_bytes() const { return active() ? std::min(c->chunk_bytes, c->stage_chunk_bytes) : c->chunk_bytes; }
    int begin(int p) {
        if (!busy[p]) return QD_OK;
        QD_CUDA(c, cudaEventSynchronize(c->pipe[p].done));
        for (const Seg& g : segs[p]) stage_pool(c)->copy(g.dst, c->pipe[p].hout.p + g.off, g.n);
        segs[p].clear();
        out_used[p] = 0;
        busy[p] = false;
        return QD_OK;
    }
    // the chunk's payload (once per chunk: the staging buffer holds one)
    int h2d(int p, void* dev, const uint8_t* host, size_t n) {
        Pipe& P = c->pipe[p];
        if (in_staged) {
            QD_CUDA(c, P.hin.reserve(n));
            stage_pool(c)->copy(P.hin.p, host, n);
            host = P.hin.p;
        }
        QD_CUDA(c, cudaMemcpyAsync(dev, host, n, cudaMemcpyHostToDevice, P.stream));
        return QD_OK;
    }
    int d2h_reserve(int p, size_t total, size_t pieces = 1) {
        if (out_staged) QD_CUDA(c, c->pipe[p].hout.reserve(total + 64 * pieces));
        return QD_OK;
    }
    int d2h(int p, uint8_t* host, const void* dev, size_t n) {
        Pipe& P = c->pipe[p];
        if (!out_staged) {
            QD_CUDA(c, cudaMemcpyAsync(host, dev, n, cudaMemcpyDeviceToHost, P.stream));
            return QD_OK;
        }
        QD_CUDA(c, cudaMemcpyAsync(P.hout.p + out_used[p], dev, n, cudaMemcpyDeviceToHost, P.stream));
        segs[p].push_back({host, out_used[p], n});
        out_used[p] += (n + 63) & ~size_t(63);
        return QD_OK;
    }
    int end(int p) {
        if (!active()) return QD_OK;
        Pipe& P = c->pipe[p];
        if (int rc = pipe_event(c, P)) return rc;
        QD_CUDA(c, cudaEventRecord(P.done, P.stream));
        busy[p] = true;
        return QD_OK;
    }
    int finish(int next_turn) {
        for (int k = 0; k < N_PIPE; k++)
            if (int rc = begin((next_turn + k) % N_PIPE)) return rc;  // oldest chunk first
        return QD_OK;
    }
};

// one large buffer up / down on stream s through the pipes' staging buffers (plain copy when pinned or small)
static int staged_upload(qd_ctx* c, void* dev, const uint8_t* host, size_t n, cudaStream_t s) {
    if (c->stage_threads == 0 || n < STAGE_MIN_BYTES || !host_pageable(host)) {
        QD_CUDA(c, cudaMemcpyAsync(dev, host, n, cudaMemcpyHostToDevice, s));
        return QD_OK;
    }
    const size_t chunk = c->stage_chunk_bytes;
    int k = 0;
    for (size_t off = 0; off < n; off += chunk, k++) {
        Pipe& P = c->pipe[k % N_PIPE];
        const size_t len = std::min(chunk, n - off);
        if (int rc = pipe_event(c, P)) return rc;
        QD_CUDA(c, cudaEventSynchronize(P.done));  // the copy that last read this staging buffer (none: returns at once)
        QD_CUDA(c, P.hin.reserve(len));
        stage_pool(c)->copy(P.hin.p, host + off, len);
        QD_CUDA(c, cudaMemcpyAsync(static_cast<uint8_t*>(dev) + off, P.hin.p, len, cudaMemcpyHostToDevice, s));
        QD_CUDA(c, cudaEventRecord(P.done, s));
    }
    return QD_OK;
}
// synchronises s
static int staged_download(qd_ctx* c, uint8_t* host, const void* dev, size_t n, cudaStream_t s) {
    if (c->stage_threads == 0 || n < STAGE_MIN_BYTES || !host_pageable(host)) {
        QD_CUDA(c, cudaMemcpyAsync(host, dev, n, cudaMemcpyDeviceToHost, s));
        QD_CUDA(c, cudaStreamSynchronize(s));
        return QD_OK;
    }
    const size_t chunk = c->stage_chunk_bytes;
    const size_t n_chunks = (n + chunk - 1) / chunk;
    for (size_t k = 0; k < n_chunks + N_PIPE; k++) {
        if (k >= N_PIPE) {  // drain chunk k - N_PIPE
            Pipe& Q = c->pipe[k % N_PIPE];
            const size_t off = (k - N_PIPE) * chunk;
            QD_CUDA(c, cudaEventSynchronize(Q.done));
            stage_pool(c)->copy(host + off, Q.hout.p, std::min(chunk, n - off));
        }
        if (k < n_chunks) {
            Pipe& P = c->pipe[k % N_PIPE];
            const size_t off = k * chunk, len = std::min(chunk, n - off);
            if (int rc = pipe_event(c, P)) return rc;
            QD_CUDA(c, P.hout.reserve(len));
            QD_CUDA(c, cudaMemcpyAsync(P.hout.p, static_cast<const uint8_t*>(dev) + off, len, cudaMemcpyDeviceToHost, s));
            QD_CUDA(c, cudaEventRecord(P.done, s));
        }
    }
    QD_CUDA(c, cudaStreamSynchronize(s));
    return QD_OK;
}

// src/rust_api.rs:310-322 on device buffers: counts per 4 KiB block, scan, compacted write.
// n_masks_dev receives the number of masks written (device uint64).
static int parse_dev(qd_ctx* c, int strict, const uint8_t* ascii_dev, size_t n, uint8_t* masks_dev, uint64_t* n_masks_dev,
                     cudaStream_t s, unsigned long long* err_key = nullptr) {
    if (n == 0) {
        if (n_masks_dev) QD_CUDA(c, cudaMemsetAsync(n_masks_dev, 0, 8, s));
        return QD_OK;
    }
    const uint64_t span = (reinterpret_cast<uintptr_t>(ascii_dev) & 15) + n;  // from the aligned start
    const uint64_t n_blocks = (span + PARSE_BLOCK_BYTES - 1) / PARSE_BLOCK_BYTES;
    // dropped bytes per block: all zero between calls (pass 1 adds, the compacting pass clears what it has used)
    if (c->parse_counts.cap < n_blocks * 8) {
        QD_CUDA(c, c->parse_counts.reserve(n_blocks * 8));
        QD_CUDA(c, cudaMemsetAsync(c->parse_counts.p, 0, c->parse_counts.cap, s));
    }
    QD_CUDA(c, c->parse_offsets.reserve((n_blocks + 1) * 8));
    ParseParams p{};
    p.in = ascii_dev;
    p.n = n;
    p.table = tab(c, strict ? TAB_PARSE_STRICT : TAB_PARSE);
    p.pair = c->d_parse_pair + (strict ? PAIR_ENTRIES : 0);
    p.pair_coef = 2u | ((2u * PAIR_K1_ASCII) << 16);
    p.counts = c->parse_counts.as<uint64_t>();
    p.offsets = c->parse_offsets.as<uint64_t>();
    p.out = masks_dev;
    p.err_key = err_key ? err_key : c->d_err;
    p.gate = c->d_gate;
    p.gate_gen = ++c->gate_gen;
    // One streaming pass finishes input without whitespace (when input and output share their 16-byte phase); the scan
    // and the compacting pass behind it return at once unless that pass dropped a byte.
    const bool same_phase = ((reinterpret_cast<uintptr_t>(ascii_dev) ^ reinterpret_cast<uintptr_t>(masks_dev)) & 15) == 0;
    const uint64_t n_gran = (span + 15) / 16;
    const uint64_t step = (uint64_t)PARSE_THREADS * PARSE_CLEAN_UNROLL;
    const unsigned g1 = (unsigned)std::min<uint64_t>((n_gran + step - 1) / step, (uint64_t)c->sm_count * 8);
    if (same_phase) parse_pass1_kernel<true><<<g1, PARSE_THREADS, 0, s>>>(p, n_gran, n_masks_dev);
    else parse_pass1_kernel<false><<<g1, PARSE_THREADS, 0, s>>>(p, n_gran, n_masks_dev);
    const unsigned grid = (unsigned)std::min<uint64_t>(n_blocks, (uint64_t)c->sm_count * 8);
    if (c->pdl) {
        QD_CUDA(c, launch_pdl(parse_scan_kernel, 1, PARSE_SCAN_THREADS, 0, s, p, n_blocks));
        QD_CUDA(c, launch_pdl(parse_write_kernel, grid, PARSE_THREADS, 0, s, p, n_blocks, n_masks_dev));
    } else {
        parse_scan_kernel<<<1, PARSE_SCAN_THREADS, 0, s>>>(p, n_blocks);
        parse_write_kernel<<<grid, PARSE_THREADS, 0, s>>>(p, n_blocks, n_masks_dev);
    }
    c->launches += 3;
    QD_CUDA(c, cudaGetLastError());
    return QD_OK;
}

extern "C" int qd_parse_dev(qd_ctx* c, int strict, const uint8_t* ascii_dev, size_t n, uint8_t* masks_dev, uint64_t* n_masks_dev,
                            void* stream) {
    if (!c || (n && (!ascii_dev || !masks_dev))) return QD_ERR_INVALID_ARGUMENT;
    Guard g(c, pick_stream(c, stream));
    if (!g.ok) return QD_ERR_CUDA;
    return parse_dev(c, strict, ascii_dev, n, masks_dev, n_masks_dev, pick_stream(c, stream));
}

extern "C" int qd_parse(qd_ctx* c, int strict, const uint8_t* ascii, size_t n, uint8_t* masks, size_t* n_masks, qd_err* err) {
    clear_err(err);
    if (n_masks) *n_masks = 0;
    if (!c || (n && (!ascii || !masks))) return QD_ERR_INVALID_ARGUMENT;
    Guard g(c);
    if (!g.ok) return QD_ERR_CUDA;
    if (n == 0) return QD_OK;
    cudaStream_t s = c->stream;
    if (c->tiny_calls && n <= TINY_IN_BYTES) {
        // tiny call: letters, mask count, error key and masks in mapped host memory -- three launches, one synchronise, no copies
        uint8_t* const h_in = c->h_tiny;
        unsigned long long* const h_cnt = reinterpret_cast<unsigned long long*>(c->h_tiny + TINY_IN_BYTES);
        unsigned long long* const h_key = reinterpret_cast<unsigned long long*>(c->h_tiny + TINY_IN_BYTES + 16);
        uint8_t* const h_out = c->h_tiny + TINY_IN_BYTES + 64;
        memcpy(h_in, ascii, n);
        *h_key = NO_ERROR_KEY;
        *h_cnt = 0;
        if (int rc = parse_dev(c, strict, h_in, n, h_out, reinterpret_cast<uint64_t*>(h_cnt), s, h_key)) return rc;
        QD_CUDA(c, cudaStreamSynchronize(s));
        const unsigned long long key = *h_key;
        if (key != NO_ERROR_KEY) return report_single(QD_ENC_ASCII, strict, ascii, key, err);
        const size_t k = (size_t)*h_cnt;
        memcpy(masks, h_out, k);
        if (n_masks) *n_masks = k;
        return QD_OK;
    }
    QD_CUDA(c, c->in.reserve(n + 16));
    QD_CUDA(c, c->out.reserve(n + 16));
    QD_CUDA(c, c->aux0.reserve(8));
    if (int rc = staged_upload(c, c->in.p, ascii, n, s)) return rc;
    if (int rc = reset_err_key(c, s)) return rc;
    if (int rc = parse_dev(c, strict, c->in.as<uint8_t>(), n, c->out.as<uint8_t>(), c->aux0.as<uint64_t>(), s)) return rc;
    QD_CUDA(c, cudaMemcpyAsync(c->h_scratch, c->aux0.p, 8, cudaMemcpyDeviceToHost, s));
    unsigned long long key = 0;
    if (int rc = fetch_err_key(c, s, &key)) return rc;  // synchronises
    if (key != NO_ERROR_KEY) return report_single(QD_ENC_ASCII, strict, ascii, key, err);
    const size_t k = (size_t)c->h_scratch[0];
    if (int rc = staged_download(c, masks, c->out.p, k, s)) return rc;
    if (n_masks) *n_masks = k;
    return QD_OK;
}

static int canonical_seq_dev(qd_ctx* c, int enc, int forward_only, const uint8_t* dna_dev, size_t n, uint8_t* out_dev, cudaStream_t s) {
    if (n == 0) return QD_OK;
    QD_CUDA(c, c->canon_state.reserve(9 * 8));
    unsigned long long* st = c->canon_state.as<unsigned long long>();
    QD_CUDA(c, cudaMemsetAsync(st, 0xff, 9 * 8, s));
    QD_CUDA(c, cudaMemsetAsync(st + 4, 0, 4 * 8, s));
    CanonSeqParams p{};
    p.dna = dna_dev;
    p.n = n;
    p.decode = tab(c, enc == QD_ENC_ASCII ? TAB_DECODE_ASCII : TAB_DECODE_MASK);
    p.letters = enc == QD_ENC_ASCII ? ('A' | ('T' << 8) | ('C' << 16) | ((uint32_t)'G' << 24)) : (1u | (2u << 8) | (4u << 16) | (8u << 24));
    p.forward_only = forward_only;
    p.state = st;
    p.out = out_dev;
    p.err_key = c->cur_err_key ? c->cur_err_key : c->d_err;
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, (uint64_t)c->sm_count * 8));
    canon_seq_scan_kernel<<<grid, 256, 0, s>>>(p);
    if (c->pdl) {  // the second and third pass take their places while the pass before drains, and wait for it
        if (!forward_only) QD_CUDA(c, launch_pdl(canon_seq_diff_kernel, grid, 256, 0, s, p));
        QD_CUDA(c, launch_pdl(canon_seq_write_kernel, grid, 256, 0, s, p));
    } else {
        if (!forward_only) canon_seq_diff_kernel<<<grid, 256, 0, s>>>(p);
        canon_seq_write_kernel<<<grid, 256, 0, s>>>(p);
    }
    c->launches += forward_only ? 2 : 3;
    QD_CUDA(c, cudaGetLastError());
    return QD_OK;
}

extern "C" int qd_canonical_dev(qd_ctx* c, int enc, int forward_only, const uint8_t* dna_dev, size_t n, uint8_t* out_dev, void* stream) {
    if (!c || (n && (!dna_dev || !out_dev))) return QD_ERR_INVALID_ARGUMENT;
    Guard g(c, pick_stream(c, stream));
    if (!g.ok) return QD_ERR_CUDA;
    return canonical_seq_dev(c, enc, forward_only, dna_dev, n, out_dev, pick_stream(c, stream));
}

// Small single-sequence host calls: result and error key lie next to each other on the device and come back in ONE copy
// through pinned memory; launch(d_in, d_out, stream) enqueues the kernels, which report to the key behind the result.
template <typename Launch>
static int small_call(qd_ctx* c, const uint8_t* in, size_t n, uint8_t* out, size_t out_bytes, unsigned long long* key, Launch launch) {
    cudaStream_t s = c->stream;
    const size_t key_off = (out_bytes + 15) & ~(size_t)15;
    QD_CUDA(c, c->in.reserve(n + 16));
    QD_CUDA(c, c->out.reserve(key_off + 16));
    uint8_t* const d = c->out.as<uint8_t>();
    if (int rc = staged_upload(c, c->in.p, in, n, s)) return rc;
    QD_CUDA(c, cudaMemsetAsync(d + key_off, 0xff, 8, s));
    c->cur_err_key = reinterpret_cast<unsigned long long*>(d + key_off);
    const int rc = launch(c->in.as<uint8_t>(), d, s);
    c->cur_err_key = nullptr;
    if (rc) return rc;
    QD_CUDA(c, cudaMemcpyAsync(c->h_stage, d, key_off + 8, cudaMemcpyDeviceToHost, s));
    QD_CUDA(c, cudaStreamSynchronize(s));
    memcpy(key, c->h_stage + key_off, 8);
    if (*key == NO_ERROR_KEY) memcpy(out, c->h_stage, out_bytes);
    return QD_OK;
}
static bool small_result(size_t out_bytes) { return out_bytes + 32 <= SMALL_CALL_BYTES; }

extern "C" int qd_canonical(qd_ctx* c, int enc, int forward_only, const uint8_t* dna, size_t n, uint8_t* out, qd_err* err) {
    clear_err(err);
    if (!c || (n && (!dna || !out))) return QD_ERR_INVALID_ARGUMENT;
    Guard g(c);
    if (!g.ok) return QD_ERR_CUDA;
    if (n == 0) return QD_OK;
    if (small_result(n)) {
        unsigned long long key = 0;
        if (int rc = small_call(c, dna, n, out, n, &key, [&](const uint8_t* d_in, uint8_t* d_out, cudaStream_t st) {
                return canonical_seq_dev(c, enc, forward_only, d_in, n, d_out, st);
            }))
            return rc;
        if (key != NO_ERROR_KEY) return report_single(enc, /*strict=*/1, dna, key, err);
        return QD_OK;
    }
    cudaStream_t s = c->stream;
    QD_CUDA(c, c->in.reserve(n + 16));
    QD_CUDA(c, c->out.reserve(n + 16));
    if (int rc = staged_upload(c, c->in.p, dna, n, s)) return rc;
    if (int rc = reset_err_key(c, s)) return rc;
    if (int rc = canonical_seq_dev(c, enc, forward_only, c->in.as<uint8_t>(), n, c->out.as<uint8_t>(), s)) return rc;
    QD_CUDA(c, cudaMemcpyAsync(out, c->out.p, n, cudaMemcpyDeviceToHost, s));
    unsigned long long key = 0;
    if (int rc = fetch_err_key(c, s, &key)) return rc;
    if (key != NO_ERROR_KEY) return report_single(enc, /*strict=*/1, dna, key, err);
    return QD_OK;
}

// ---- FASTA record scanner (src/fasta.rs:207-491) ------------------------------------------------------------------
static_assert(sizeof(qd_fasta_record) == sizeof(FastaRecord), "qd_fasta_record and the kernel's record must agree");

namespace {
struct FastaRun {
    FastaParams p{};
    uint64_t* pre_base = nullptr;  // device: blk_last | blk_state | cnt x3 | pre x3
};
}  // namespace

// passes 1-2 and the scans; totals stay on the device (p.pre_*[n_blocks], p.scalars)
static int fasta_count_dev(qd_ctx* c, int strict, int concat, int allow, const uint8_t* text_dev, size_t n, FastaRun& R, cudaStream_t s) {
    FastaParams& p = R.p;
    p.text = text_dev;
    p.n = n;
    p.table = tab(c, strict ? TAB_PARSE_STRICT : TAB_PARSE);
    p.concatenate_headers = concat;
    p.allow_preceding_comment = allow;
    const uint64_t span = (reinterpret_cast<uintptr_t>(text_dev) & 15) + n;
    p.n_blocks = std::max<uint64_t>(1, (span + FA_BLOCK_BYTES - 1) / FA_BLOCK_BYTES);
    const uint64_t nb1 = p.n_blocks + 1;
    QD_CUDA(c, c->fasta_scratch.reserve((8 * nb1 + 2) * 8));
    uint64_t* base = c->fasta_scratch.as<uint64_t>();
    p.scalars = reinterpret_cast<unsigned long long*>(base);
    p.blk_last = base + 2;
    uint64_t* blk_state = base + 2 + nb1;
    p.blk_state = blk_state;
    p.cnt_kept = base + 2 + 2 * nb1;
    p.cnt_recs = base + 2 + 3 * nb1;
    p.cnt_nl = base + 2 + 4 * nb1;
    uint64_t* pre_kept = base + 2 + 5 * nb1;
    uint64_t* pre_recs = base + 2 + 6 * nb1;
    uint64_t* pre_nl = base + 2 + 7 * nb1;
    p.pre_kept = pre_kept;
    p.pre_recs = pre_recs;
    p.pre_nl = pre_nl;
    p.err_key = c->d_err;
    QD_CUDA(c, cudaMemsetAsync(p.scalars, 0xff, 8, s));
    QD_CUDA(c, cudaMemsetAsync(p.scalars + 1, 0, 8, s));
    const unsigned grid = (unsigned)std::min<uint64_t>(p.n_blocks, (uint64_t)c->sm_count * 8);
    fasta_lines_kernel<<<grid, FA_THREADS, 0, s>>>(p);
    c->launches++;
    QD_CUDA(c, cudaGetLastError());
    // every kernel from here on is a programmatic dependent of the one before: it takes its place early and waits
    if (int rc = device_exclusive_scan<1, 1>(c, p.blk_last, p.n_blocks, blk_state, c->block_sums, s, nullptr, 0, true)) return rc;
    if (c->pdl) QD_CUDA(c, launch_pdl(fasta_pass_kernel<false>, grid, FA_THREADS, 0, s, p));
    else fasta_pass_kernel<false><<<grid, FA_THREADS, 0, s>>>(p);
    c->launches++;
    QD_CUDA(c, cudaGetLastError());
    // kept bytes, record starts and line ends per block: three arrays nb1 entries apart, scanned by one set of launches
    if (int rc = device_exclusive_scan<1>(c, p.cnt_kept, p.n_blocks, pre_kept, c->block_sums, s, nullptr, 0, true, 3, nb1)) return rc;
    return QD_OK;
}

static int fasta_write_dev(qd_ctx* c, FastaRun& R, uint8_t* masks_dev, FastaRecord* records_dev, uint64_t max_records, uint64_t n_records,
                           uint64_t shift, cudaStream_t s) {
    FastaParams& p = R.p;
    p.masks = masks_dev;
    p.records = records_dev;
    p.max_records = max_records;
    const unsigned grid = (unsigned)std::min<uint64_t>(p.n_blocks, (uint64_t)c->sm_count * 8);
    fasta_pass_kernel<true><<<grid, FA_THREADS, 0, s>>>(p);  // (behind a synchronisation: a plain launch)
    c->launches++;
    if (n_records) {
        const unsigned g4 = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n_records + 255) / 256, (uint64_t)c->sm_count * 8));
        if (c->pdl) QD_CUDA(c, launch_pdl(fasta_finish_kernel, g4, 256, 0, s, p, n_records, shift));
        else fasta_finish_kernel<<<g4, 256, 0, s>>>(p, n_records, shift);
        c->launches++;
    }
    QD_CUDA(c, cudaGetLastError());
    return QD_OK;
}

// after fasta_count_dev: totals and the first content error (synchronises the stream once).  `text` is the host copy of the
// input, or NULL when it only exists on the device (text_dev): then the bytes in front of an error are fetched to count its line.
static int fasta_totals(qd_ctx* c, int strict, int allow_preceding_comment, const uint8_t* text, const FastaRun& R, cudaStream_t s,
                        uint64_t* total_kept, uint64_t* total_records, uint64_t* shift_out, qd_err* err) {
    const FastaParams& p = R.p;
    // one single-warp kernel writes the four numbers into mapped host memory: no copies, one synchronisation
    unsigned long long* const h = reinterpret_cast<unsigned long long*>(c->h_tiny + TINY_IN_BYTES);
    if (c->pdl) QD_CUDA(c, launch_pdl(fasta_totals_kernel, 1, 32, 0, s, p, h));
    else fasta_totals_kernel<<<1, 32, 0, s>>>(p, h);
    c->launches++;
    QD_CUDA(c, cudaGetLastError());
    QD_CUDA(c, cudaStreamSynchronize(s));
    const uint64_t kept = h[0], rec_starts = h[1], before = h[2];
    const unsigned long long key = h[3];
    if (key != NO_ERROR_KEY) {
        std::vector<uint8_t> head;
        if (!text) {
            head.resize(key + 1);
            QD_CUDA(c, cudaMemcpyAsync(head.data(), p.text, key + 1, cudaMemcpyDeviceToHost, s));
            QD_CUDA(c, cudaStreamSynchronize(s));
            text = head.data();
        }
        uint64_t line = 1;  // Located::line_number (src/errors.rs:8-15): lines are counted as BufRead::lines does
        for (size_t i = 0; i < key; i++) line += text[i] == '\n';
        uint8_t payload = 0;
        int kind = classify_byte(QD_ENC_ASCII, strict, text[key], &payload);
        if (kind == QD_OK) kind = QD_ERR_BAD_NUCLEOTIDE;
        return set_err(err, kind, payload, key, line);
    }
    *shift_out = (!allow_preceding_comment && before > 0) ? 1 : 0;
    *total_kept = kept;
    *total_records = rec_starts + *shift_out;
    return QD_OK;
}

extern "C" int qd_fasta_scan(qd_ctx* c, int strict, int concatenate_headers, int allow_preceding_comment, const uint8_t* text, size_t n,
                             uint8_t* masks, size_t* n_masks, qd_fasta_record* records, size_t max_records, size_t* n_records, qd_err* err) {
    clear_err(err);
    if (n_masks) *n_masks = 0;
    if (n_records) *n_records = 0;
    if (!c || (n && !text)) return QD_ERR_INVALID_ARGUMENT;
    Guard g(c);
    if (!g.ok) return QD_ERR_CUDA;
    if (n == 0) return QD_OK;
    cudaStream_t s = c->stream;
    QD_CUDA(c, c->in.reserve(n + 16));
    if (int rc = staged_upload(c, c->in.p, text, n, s)) return rc;
    if (int rc = reset_err_key(c, s)) return rc;
    FastaRun R;
    if (int rc = fasta_count_dev(c, strict, concatenate_headers, allow_preceding_comment, c->in.as<uint8_t>(), n, R, s)) return rc;
    uint64_t total_kept = 0, total_records = 0, shift = 0;
    if (int rc = fasta_totals(c, strict, allow_preceding_comment, text, R, s, &total_kept, &total_records, &shift, err)) return rc;
    if (n_masks) *n_masks = (size_t)total_kept;
    if (n_records) *n_records = (size_t)total_records;
    if (!masks && !records) return QD_OK;  // counting call
    if (!masks || (total_records && !records) || max_records < total_records) return QD_ERR_INVALID_ARGUMENT;
    // small results: masks and records next to each other on the device, ONE copy into pinned memory, two host copies out
    const size_t rec_off = ((size_t)total_kept + 15) & ~(size_t)15, rec_bytes = (size_t)total_records * sizeof(FastaRecord);
    if (rec_off + rec_bytes <= SMALL_CALL_BYTES) {
        QD_CUDA(c, c->out.reserve(rec_off + rec_bytes + sizeof(FastaRecord)));
        if (int rc = fasta_write_dev(c, R, c->out.as<uint8_t>(), reinterpret_cast<FastaRecord*>(c->out.as<uint8_t>() + rec_off), total_records, total_records,
                                     shift, s))
            return rc;
        if (rec_off + rec_bytes) QD_CUDA(c, cudaMemcpyAsync(c->h_stage, c->out.p, rec_off + rec_bytes, cudaMemcpyDeviceToHost, s));
        QD_CUDA(c, cudaStreamSynchronize(s));
        if (total_kept) memcpy(masks, c->h_stage, total_kept);
        if (rec_bytes) memcpy(records, c->h_stage + rec_off, rec_bytes);
        return QD_OK;
    }
    QD_CUDA(c, c->out.reserve(total_kept + 16));
    QD_CUDA(c, c->aux0.reserve((total_records + 1) * sizeof(FastaRecord)));
    if (int rc = fasta_write_dev(c, R, c->out.as<uint8_t>(), c->aux0.as<FastaRecord>(), total_records, total_records, shift, s)) return rc;
    if (total_records) QD_CUDA(c, cudaMemcpyAsync(records, c->aux0.p, total_records * sizeof(FastaRecord), cudaMemcpyDeviceToHost, s));
    if (total_kept)
        if (int rc = staged_download(c, masks, c->out.p, total_kept, s)) return rc;
    QD_CUDA(c, cudaStreamSynchronize(s));
    return QD_OK;
}

// the scanner on device buffers: text_dev in; masks_dev (capacity n), records_dev (capacity max_records) and offsets_dev
// (capacity max_records + 1, optional) out.  Synchronises `stream` once (the totals size the second pass); on a content error
// err->pos is the byte offset and err->seq_index the 1-based line number, counted on the device copy of the text.
extern "C" int qd_fasta_scan_dev(qd_ctx* c, int strict, int concatenate_headers, int allow_preceding_comment, const uint8_t* text_dev, size_t n,
                                 uint8_t* masks_dev, size_t* n_masks, qd_fasta_record* records_dev, size_t max_records, size_t* n_records,
                                 uint64_t* offsets_dev, qd_err* err, void* stream) {
    clear_err(err);
    if (n_masks) *n_masks = 0;
    if (n_records) *n_records = 0;
    if (!c || (n && !text_dev)) return QD_ERR_INVALID_ARGUMENT;
    cudaStream_t s = pick_stream(c, stream);
    Guard g(c, s);
    if (!g.ok) return QD_ERR_CUDA;
    if (n == 0) return QD_OK;
    if (int rc = reset_err_key(c, s)) return rc;
    FastaRun R;
    if (int rc = fasta_count_dev(c, strict, concatenate_headers, allow_preceding_comment, text_dev, n, R, s)) return rc;
    // the totals; a content error is located with the text still on the device: its byte, and the newlines in front of it
    uint64_t total_kept = 0, total_records = 0, shift = 0;
    if (int rc = fasta_totals(c, strict, allow_preceding_comment, nullptr, R, s, &total_kept, &total_records, &shift, err)) return rc;
    if (n_masks) *n_masks = (size_t)total_kept;
    if (n_records) *n_records = (size_t)total_records;
    if (!masks_dev && !records_dev) return QD_OK;  // counting call
    if (!masks_dev || (total_records && !records_dev) || max_records < total_records) return QD_ERR_INVALID_ARGUMENT;
    if (int rc = fasta_write_dev(c, R, masks_dev, reinterpret_cast<FastaRecord*>(records_dev), total_records, total_records, shift, s)) return rc;
    if (offsets_dev) {
        const unsigned g1 = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((total_records + 256) / 256, (uint64_t)c->sm_count * 8));
        if (c->pdl) QD_CUDA(c, launch_pdl(fasta_offsets_kernel, g1, 256, 0, s, reinterpret_cast<const FastaRecord*>(records_dev), total_records, total_kept, offsets_dev));
        else fasta_offsets_kernel<<<g1, 256, 0, s>>>(reinterpret_cast<const FastaRecord*>(records_dev), total_records, total_kept, offsets_dev);
        c->launches++;
        QD_CUDA(c, cudaGetLastError());
    }
    return QD_OK;
}

// FASTA text in, the frames of every record out: the scanner's masks and CSR offsets feed the frame kernels on the device,
// nothing but the text goes up and nothing but records and frames comes back.
extern "C" int qd_fasta_translate_frames(qd_ctx* c, int table, int strict, int frames, int concatenate_headers, int allow_preceding_comment,
                                         const uint8_t* text, size_t n, qd_fasta_record* records, size_t max_records, size_t* n_records,
                                         uint8_t* out, size_t out_capacity, uint64_t* out_offsets, size_t* out_bytes, qd_err* err) {
    clear_err(err);
    if (n_records) *n_records = 0;
    if (out_bytes) *out_bytes = 0;
    if (!c || !valid_frames(frames) || (n && !text)) return QD_ERR_INVALID_ARGUMENT;
    int slot;
    if (int rc = check_table(table, err, &slot)) return rc;  // table first, as everywhere
    Guard g(c);
    if (!g.ok) return QD_ERR_CUDA;
    if (n == 0) return QD_OK;
    cudaStream_t s = c->stream;
    QD_CUDA(c, c->in.reserve(n + 16));
    if (int rc = staged_upload(c, c->in.p, text, n, s)) return rc;
    if (int rc = reset_err_key(c, s)) return rc;
    FastaRun R;
    if (int rc = fasta_count_dev(c, strict, concatenate_headers, allow_preceding_comment, c->in.as<uint8_t>(), n, R, s)) return rc;
    uint64_t total_kept = 0, total_records = 0, shift = 0;
    if (int rc = fasta_totals(c, strict, allow_preceding_comment, text, R, s, &total_kept, &total_records, &shift, err)) return rc;
    if (n_records) *n_records = (size_t)total_records;
    if (total_records == 0) return QD_OK;
    // masks and records stay on the device
    QD_CUDA(c, c->aux1.reserve(total_kept + 16));
    QD_CUDA(c, c->aux0.reserve((total_records + 1) * sizeof(FastaRecord)));
    QD_CUDA(c, c->aux2.reserve((total_records + 1) * sizeof(uint64_t)));
    if (int rc = fasta_write_dev(c, R, c->aux1.as<uint8_t>(), c->aux0.as<FastaRecord>(), total_records, total_records, shift, s)) return rc;
    const unsigned g1 = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((total_records + 256) / 256, (uint64_t)c->sm_count * 8));
    if (c->pdl) QD_CUDA(c, launch_pdl(fasta_offsets_kernel, g1, 256, 0, s, (const FastaRecord*)c->aux0.as<FastaRecord>(), total_records, total_kept, c->aux2.as<uint64_t>()));
    else fasta_offsets_kernel<<<g1, 256, 0, s>>>(c->aux0.as<FastaRecord>(), total_records, total_kept, c->aux2.as<uint64_t>());
    c->launches++;
    QD_CUDA(c, cudaGetLastError());
    // frames of all records (slot layout of qd_translate_frames_batch); the exact size is the run total the device finds
    const uint64_t stride = 16 * (uint64_t)frames;
    const uint64_t runs_bound = total_kept / RUN_NT + total_records;
    QD_CUDA(c, c->out.reserve(runs_bound * stride + 16));
    if (int rc = frames_csr_dev(c, slot, QD_ENC_MASK, strict, frames, c->aux1.as<uint8_t>(), total_kept, c->aux2.as<uint64_t>(), total_records,
                                c->run_prefix, c->tile_first, c->block_sums, c->out.as<uint8_t>(), s))
        return rc;
    // small results: records, run prefix and frames (up to the bound) into pinned memory behind the kernels, ONE synchronisation
    const size_t rec_bytes = (size_t)total_records * sizeof(FastaRecord), pre_off = (rec_bytes + 15) & ~(size_t)15;
    const size_t fr_off = (pre_off + (size_t)(total_records + 1) * 8 + 15) & ~(size_t)15;
    if (out && records && max_records >= total_records && fr_off + runs_bound * stride <= SMALL_CALL_BYTES) {
        QD_CUDA(c, cudaMemcpyAsync(c->h_stage, c->aux0.p, rec_bytes, cudaMemcpyDeviceToHost, s));
        QD_CUDA(c, cudaMemcpyAsync(c->h_stage + pre_off, c->run_prefix.p, (total_records + 1) * 8, cudaMemcpyDeviceToHost, s));
        if (runs_bound) QD_CUDA(c, cudaMemcpyAsync(c->h_stage + fr_off, c->out.p, runs_bound * stride, cudaMemcpyDeviceToHost, s));
        QD_CUDA(c, cudaStreamSynchronize(s));
        const uint64_t* pre = reinterpret_cast<const uint64_t*>(c->h_stage + pre_off);
        const uint64_t need = pre[total_records] * stride;
        if (out_bytes) *out_bytes = (size_t)need;
        if (out_capacity < need) return QD_ERR_INVALID_ARGUMENT;
        memcpy(records, c->h_stage, rec_bytes);
        if (out_offsets)
            for (uint64_t i = 0; i <= total_records; i++) out_offsets[i] = pre[i] * stride;
        memcpy(out, c->h_stage + fr_off, need);
        return QD_OK;
    }
    QD_CUDA(c, cudaMemcpyAsync(c->h_scratch, c->run_prefix.as<uint64_t>() + total_records, 8, cudaMemcpyDeviceToHost, s));
    QD_CUDA(c, cudaStreamSynchronize(s));
    const uint64_t need = c->h_scratch[0] * stride;
    if (out_bytes) *out_bytes = (size_t)need;
    if (!out && !records) return QD_OK;  // sizing call
    if (!out || !records || max_records < total_records || out_capacity < need) return QD_ERR_INVALID_ARGUMENT;
    QD_CUDA(c, cudaMemcpyAsync(records, c->aux0.p, total_records * sizeof(FastaRecord), cudaMemcpyDeviceToHost, s));
    if (out_offsets) {
        QD_CUDA(c, cudaMemcpyAsync(out_offsets, c->run_prefix.p, (total_records + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        QD_CUDA(c, cudaStreamSynchronize(s));
        for (uint64_t i = 0; i <= total_records; i++) out_offsets[i] *= stride;
    }
    if (need)
        if (int rc = staged_download(c, out, c->out.p, need, s)) return rc;
    QD_CUDA(c, cudaStreamSynchronize(s));
    return QD_OK;
}

static int windows_batch_dev(qd_ctx* c, const uint8_t* src_dev, uint64_t stride, size_t n_seq, size_t len, size_t w, int mode,
                             uint8_t* out_dev, uint64_t key_stride, cudaStream_t s, int dep = 0);
static int windows_all_batch_dev(qd_ctx* c, int enc, const uint8_t* dna_dev, size_t n_seq, size_t seq_len, size_t w_dna, size_t w_aa,
                                 uint8_t* dna_out_dev, uint8_t* aa_out_dev, cudaStream_t s);

static int windows_dev(qd_ctx* c, const uint8_t* seq_dev, size_t n, size_t w, uint8_t* out_dev, cudaStream_t s) {
    if (w == 0 || w > 0xffffffffull) return QD_ERR_INVALID_ARGUMENT;
    if (n < w) return QD_OK;
    // windows of up to 64 bytes: the warp-tile kernel of the batch form (rows staged in shared memory, 128-bit stores)
    if (w <= 64 && n <= 0xffffffffull) return windows_batch_dev(c, seq_dev, n, 1, n, w, QD_WIN_COPY, out_dev, n, s);
    const uint64_t n_windows = n - w + 1;
    const uint64_t chunks = (n_windows * w + 15) >> 4;
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((chunks + 255) / 256, (uint64_t)c->sm_count * 16));
    windows_kernel<<<grid, 256, 0, s>>>(seq_dev, n_windows, (uint32_t)w, out_dev);
    c->launches++;
    QD_CUDA(c, cudaGetLastError());
    return QD_OK;
}

// windows of n_seq sequences (sequence s at src + s * stride, `len` bytes each), see windows_batch_kernel
// dep: DEP_NONE = a plain launch; DEP_FREE / DEP_WAIT = a programmatic dependent of the kernel before it in the stream that
// may start while that kernel drains -- DEP_WAIT when it reads what that kernel wrote, DEP_FREE when everything it reads was
// complete before the first kernel of the chain started (WindowsBatchParams::dep_wait)
enum { DEP_NONE = 0, DEP_FREE = 1, DEP_WAIT = 2 };
template <typename K>
static cudaError_t launch_windows(qd_ctx* c, K kern, unsigned grid, cudaStream_t s, int dep, const WindowsBatchParams& p) {
    if (dep != DEP_NONE && c->pdl) return launch_pdl(kern, grid, 32 * WIN_WARPS, 0, s, p);
    kern<<<grid, 32 * WIN_WARPS, 0, s>>>(p);
    return cudaSuccess;
}
static int windows_batch_dev(qd_ctx* c, const uint8_t* src_dev, uint64_t stride, size_t n_seq, size_t len, size_t w, int mode,
                             uint8_t* out_dev, uint64_t key_stride, cudaStream_t s, int dep) {
    if (w < 1 || w > 64 || len > 0xffffffffull || n_seq >= 0xffffffffull || mode < QD_WIN_COPY || mode > QD_WIN_MASK_ASCII_RC) return QD_ERR_INVALID_ARGUMENT;
    if (len < w || n_seq == 0) return QD_OK;
    static const int map_tab[5] = {-1, TAB_UPPER_ASCII_STRICT, TAB_RC_ASCII_STRICT, TAB_MASK_TO_ASCII_STRICT, TAB_RC_MASK_TO_ASCII_STRICT};
    WindowsBatchParams p{};
    p.src = src_dev;
    p.src_stride = stride;
    p.n_seq = n_seq;
    p.len = (uint32_t)len;
    p.w = (uint32_t)w;
    p.reverse = (mode == QD_WIN_DNA_ASCII_RC || mode == QD_WIN_MASK_ASCII_RC) ? 1 : 0;
    p.map = map_tab[mode] < 0 ? nullptr : tab(c, map_tab[mode]);
    p.out = out_dev;
    p.key_stride = key_stride;
    p.err_key = c->d_err;
    p.dep_wait = (dep == DEP_WAIT && c->pdl) ? 1u : 0u;
    const uint64_t wpt = w == 42 ? WinTile<42>::WPT : (w == 20 ? WinTile<20>::WPT : WinTile<0>::WPT);
    const uint64_t tiles = (uint64_t)n_seq * (1 + (len - w + 1 + wpt - 1) / wpt);  // + the short first tile of every sequence
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((tiles + WIN_WARPS - 1) / WIN_WARPS, (uint64_t)c->sm_count * 6));
    // the two window lengths of benches/all_windows.rs get fully unrolled instantiations
    if (p.map) {
        if (w == 42) QD_CUDA(c, launch_windows(c, windows_batch_kernel<42, true>, grid, s, dep, p));
        else if (w == 20) QD_CUDA(c, launch_windows(c, windows_batch_kernel<20, true>, grid, s, dep, p));
        else QD_CUDA(c, launch_windows(c, windows_batch_kernel<0, true>, grid, s, dep, p));
    } else {  // plain copies: no table lookup per byte
        if (w == 42) QD_CUDA(c, launch_windows(c, windows_batch_kernel<42, false>, grid, s, dep, p));
        else if (w == 20) QD_CUDA(c, launch_windows(c, windows_batch_kernel<20, false>, grid, s, dep, p));
        else QD_CUDA(c, launch_windows(c, windows_batch_kernel<0, false>, grid, s, dep, p));
    }
    c->launches++;
    QD_CUDA(c, cudaGetLastError());
    return QD_OK;
}

extern "C" int qd_windows_batch_dev(qd_ctx* c, const uint8_t* seq_dev, size_t n_seq, size_t seq_len, size_t w, int mode, uint8_t* out_dev,
                                    void* stream) {
    if (!c) return QD_ERR_INVALID_ARGUMENT;
    Guard g(c, pick_stream(c, stream));
    if (!g.ok) return QD_ERR_CUDA;
    return windows_batch_dev(c, seq_dev, seq_len, n_seq, seq_len, w, mode, out_dev, seq_len, pick_stream(c, stream));
}

extern "C" size_t qd_windows_all_batch_counts(size_t seq_len, size_t w_dna, size_t w_aa, size_t counts[8]) {
    auto nwin = [](size_t len, size_t w) { return len + 1 - std::min(w, len + 1); };  // benches/all_windows.rs:209-211
    size_t total = 0;
    counts[0] = counts[1] = nwin(seq_len, w_dna);
    for (int f = 0; f < 6; f++) {
        const size_t fl = qd_frame_len(seq_len, f % 3);
        counts[2 + f] = fl ? nwin(fl, w_aa) : 0;
    }
    for (int i = 0; i < 8; i++) total += counts[i];
    return total;
}

extern "C" int qd_windows_all_batch_dev(qd_ctx* c, int enc, const uint8_t* dna_dev, size_t n_seq, size_t seq_len, size_t w_dna, size_t w_aa,
                                        uint8_t* dna_out_dev, uint8_t* aa_out_dev, void* stream) {
    if (!c || w_dna < 1 || w_aa < 1 || w_dna > 64 || w_aa > 64) return QD_ERR_INVALID_ARGUMENT;
    Guard g(c, pick_stream(c, stream));
    if (!g.ok) return QD_ERR_CUDA;
    return windows_all_batch_dev(c, enc, dna_dev, n_seq, seq_len, w_dna, w_aa, dna_out_dev, aa_out_dev, g.s);
}

static int windows_all_batch_dev(qd_ctx* c, int enc, const uint8_t* dna_dev, size_t n_seq, size_t seq_len, size_t w_dna, size_t w_aa,
                                 uint8_t* dna_out_dev, uint8_t* aa_out_dev, cudaStream_t s) {
    if (n_seq == 0 || seq_len == 0) return QD_OK;
    size_t counts[8];
    qd_windows_all_batch_counts(seq_len, w_dna, w_aa, counts);
    // forward and reverse-complement DNA windows straight from the input (strict alphabet: anything else is an error)
    const int fwd_mode = enc == QD_ENC_ASCII ? QD_WIN_DNA_ASCII : QD_WIN_MASK_ASCII;
    if (dna_out_dev && counts[0]) {
        // The eight window kernels and the frame kernel between them are one chain of programmatic dependents: each may
        // take its place on the SMs while the one before drains.  The first launch is a plain one, so everything the chain
        // reads from outside (the input; the scratch of the frames: no reader of an earlier call is left) is settled
        // before it starts; only the first window kernel over the frames has to wait for the kernel before it.
        if (int rc = windows_batch_dev(c, dna_dev, seq_len, n_seq, seq_len, w_dna, fwd_mode, dna_out_dev, seq_len, s, DEP_NONE)) return rc;
        if (int rc = windows_batch_dev(c, dna_dev, seq_len, n_seq, seq_len, w_dna, fwd_mode + 1, dna_out_dev + n_seq * counts[0] * w_dna, seq_len, s, DEP_FREE)) return rc;
    }
    if (aa_out_dev) {
        // six frames (table Ncbi1, as SequenceWindows::from_dna does) into context scratch, then their windows
        const size_t slots = qd_slots_bytes(seq_len, 6);
        QD_CUDA(c, c->aux2.reserve(n_seq * slots + 16));
        const bool chained = dna_out_dev && counts[0];  // (else the frame kernel is the plain first launch of the chain)
        if (int rc = frames_uniform_dev(c, 0, enc, 1, QD_FRAMES_ALL, dna_dev, n_seq, seq_len, c->aux2.as<uint8_t>(), s, 0, nullptr, chained)) return rc;
        size_t at = 0;
        int dep = DEP_WAIT;  // the first kernel over the frames waits for them; the others start behind it
        for (int f = 0; f < 6; f++) {
            if (!counts[2 + f]) continue;
            const uint8_t* frame0 = c->aux2.as<uint8_t>() + qd_slot_frame_offset(seq_len, 6, f);
            if (int rc = windows_batch_dev(c, frame0, slots, n_seq, qd_frame_len(seq_len, f % 3), w_aa, QD_WIN_COPY, aa_out_dev + at, seq_len, s, dep)) return rc;
            dep = DEP_FREE;
            at += n_seq * counts[2 + f] * w_aa;
        }
    }
    return QD_OK;
}

extern "C" int qd_windows_dev(qd_ctx* c, const uint8_t* seq_dev, size_t n, size_t w, uint8_t* out_dev, void* stream) {
    if (!c || (reinterpret_cast<uintptr_t>(out_dev) & 15)) return QD_ERR_INVALID_ARGUMENT;
    Guard g(c, pick_stream(c, stream));
    if (!g.ok) return QD_ERR_CUDA;
    return windows_dev(c, seq_dev, n, w, out_dev, pick_stream(c, stream));
}

// ---- checksum and chunked ("ring") window enumeration -----------------------------------------------------------
static int checksum_dev(qd_ctx* c, const uint8_t* buf_dev, uint64_t n, const uint64_t* n_dev, uint64_t n_scale, uint64_t base,
                        uint64_t* acc_dev, cudaStream_t s) {
    if (n == 0 && !n_dev) return QD_OK;
    ChecksumParams p{};
    p.buf = buf_dev;
    p.n = n;
    p.n_dev = n_dev;
    p.n_scale = n_scale;
    p.base = base;
    p.acc = reinterpret_cast<unsigned long long*>(acc_dev);
    const uint64_t gran = ((n_dev ? n * n_scale : n) >> 4) + 1;  // with a device-side count, `n` is the caller's bound on it (sizes the grid only)
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((gran + CHECKSUM_THREADS - 1) / CHECKSUM_THREADS, (uint64_t)c->sm_count * 8));
    checksum64_kernel<<<grid, CHECKSUM_THREADS, 0, s>>>(p);
    c->launches++;
    QD_CUDA(c, cudaGetLastError());
    return QD_OK;
}

extern "C" int qd_checksum64_dev(qd_ctx* c, const uint8_t* buf_dev, size_t n, uint64_t base, uint64_t* acc_dev, void* stream) {
    if (!c || !acc_dev || (n && !buf_dev)) return QD_ERR_INVALID_ARGUMENT;
    Guard g(c, pick_stream(c, stream));
    if (!g.ok) return QD_ERR_CUDA;
    return checksum_dev(c, buf_dev, n, nullptr, 0, base, acc_dev, g.s);
}

extern "C" uint64_t qd_checksum64_weight(uint64_t word_index) { return checksum_weight(word_index); }

extern "C" size_t qd_chunk_count(size_t n_units, size_t chunk_units) { return chunk_units ? (n_units + chunk_units - 1) / chunk_units : 0; }

extern "C" size_t qd_canonical_chunk_bytes(size_t chunk_seqs, size_t seq_len, size_t w) {
    return (seq_len >= w && w) ? chunk_seqs * (seq_len - w + 1) * w : 0;
}

extern "C" size_t qd_windows_all_chunk_bytes(size_t chunk_seqs, size_t seq_len, size_t w_dna, size_t w_aa, size_t* aa_offset) {
    size_t counts[8];
    qd_windows_all_batch_counts(seq_len, w_dna, w_aa, counts);
    const size_t dna_bytes = 2 * chunk_seqs * counts[0] * w_dna;
    size_t aa_bytes = 0;
    for (int f = 0; f < 6; f++) aa_bytes += chunk_seqs * counts[2 + f] * w_aa;
    const size_t aa_off = (dna_bytes + 15) & ~(size_t)15;
    if (aa_offset) *aa_offset = aa_off;
    return aa_off + aa_bytes;
}

namespace {
struct Ring {
    uint8_t* base;
    size_t slot_bytes, n_slots;
    uint8_t* slot(size_t chunk) const { return base + (chunk % n_slots) * slot_bytes; }
};
// slots of `need` bytes (rounded up to 256) cut from the caller's ring; false when not even one fits
bool make_ring(uint8_t* ring_dev, size_t ring_bytes, size_t need, Ring* r) {
    r->base = ring_dev;
    r->slot_bytes = (need + 255) & ~(size_t)255;
    r->n_slots = r->slot_bytes ? ring_bytes / r->slot_bytes : 0;
    return ring_dev && r->n_slots >= 1 && (reinterpret_cast<uintptr_t>(ring_dev) & 15) == 0;
}
}  // namespace

extern "C" int qd_canonical_windows_chunked_dev(qd_ctx* c, int enc, int forward_only, const uint8_t* dna_dev, size_t n_seq, size_t seq_len,
                                                size_t w, uint8_t* ring_dev, size_t ring_bytes, size_t chunk_seqs, uint64_t* checksums_dev,
                                                qd_chunk_fn consume, void* user, void* stream) {
    if (!c || w < 1 || w > CANON_MAX_W || chunk_seqs == 0) return QD_ERR_INVALID_ARGUMENT;
    Guard g(c, pick_stream(c, stream));
    if (!g.ok) return QD_ERR_CUDA;
    if (n_seq == 0 || seq_len < w) return QD_OK;
    Ring ring;
    const size_t per_seq = (seq_len - w + 1) * w;
    if (!make_ring(ring_dev, ring_bytes, std::min(chunk_seqs, n_seq) * per_seq, &ring)) {
        c->last_error = "ring smaller than one chunk (qd_canonical_chunk_bytes) or not 16-byte aligned";
        return QD_ERR_INVALID_ARGUMENT;
    }
    const size_t n_chunks = qd_chunk_count(n_seq, chunk_seqs);
    if (checksums_dev) QD_CUDA(c, cudaMemsetAsync(checksums_dev, 0, n_chunks * 8, g.s));
    for (size_t k = 0; k < n_chunks; k++) {
        const size_t s0 = k * chunk_seqs, cnt = std::min(chunk_seqs, n_seq - s0);
        uint8_t* out = ring.slot(k);
        if (int rc = canonical_dev(c, enc, forward_only, dna_dev + s0 * seq_len, cnt, seq_len, w, out, g.s)) return rc;
        // the error key of a chunk is relative to the chunk's first byte: callers that need positions run chunks of their own
        if (checksums_dev)
            if (int rc = checksum_dev(c, out, cnt * per_seq, nullptr, 0, (uint64_t)s0 * per_seq, checksums_dev + k, g.s)) return rc;
        if (consume) consume(user, k, s0, cnt, out, cnt * per_seq, nullptr, (void*)g.s);
    }
    return QD_OK;
}

extern "C" int qd_windows_all_chunked_dev(qd_ctx* c, int enc, const uint8_t* dna_dev, size_t n_seq, size_t seq_len, size_t w_dna, size_t w_aa,
                                          uint8_t* ring_dev, size_t ring_bytes, size_t chunk_seqs, uint64_t* checksums_dev, qd_chunk_fn consume,
                                          void* user, void* stream) {
    if (!c || w_dna < 1 || w_aa < 1 || w_dna > 64 || w_aa > 64 || chunk_seqs == 0) return QD_ERR_INVALID_ARGUMENT;
    Guard g(c, pick_stream(c, stream));
    if (!g.ok) return QD_ERR_CUDA;
    if (n_seq == 0 || seq_len == 0) return QD_OK;
    Ring ring;
    size_t aa_off = 0;
    if (!make_ring(ring_dev, ring_bytes, qd_windows_all_chunk_bytes(std::min(chunk_seqs, n_seq), seq_len, w_dna, w_aa, &aa_off), &ring)) {
        c->last_error = "ring smaller than one chunk (qd_windows_all_chunk_bytes) or not 16-byte aligned";
        return QD_ERR_INVALID_ARGUMENT;
    }
    const size_t n_chunks = qd_chunk_count(n_seq, chunk_seqs);
    if (checksums_dev) QD_CUDA(c, cudaMemsetAsync(checksums_dev, 0, n_chunks * 8, g.s));
    for (size_t k = 0; k < n_chunks; k++) {
        const size_t s0 = k * chunk_seqs, cnt = std::min(chunk_seqs, n_seq - s0);
        size_t off = 0;
        const size_t bytes = qd_windows_all_chunk_bytes(cnt, seq_len, w_dna, w_aa, &off);
        uint8_t* out = ring.slot(k);
        if (int rc = windows_all_batch_dev(c, enc, dna_dev + s0 * seq_len, cnt, seq_len, w_dna, w_aa, out, out + off, g.s)) return rc;
        if (checksums_dev) {
            // chunk-relative stream: the DNA windows, then (without the alignment gap) the amino-acid windows
            size_t counts[8];
            qd_windows_all_batch_counts(seq_len, w_dna, w_aa, counts);
            const uint64_t dna_bytes = 2 * (uint64_t)cnt * counts[0] * w_dna;
            if (int rc = checksum_dev(c, out, dna_bytes, nullptr, 0, 0, checksums_dev + k, g.s)) return rc;
            if (int rc = checksum_dev(c, out + off, bytes - off, nullptr, 0, dna_bytes, checksums_dev + k, g.s)) return rc;
        }
        if (consume) consume(user, k, s0, cnt, out, bytes, nullptr, (void*)g.s);
    }
    return QD_OK;
}

// =================================================================================================
// host-pointer entry points
// =================================================================================================
// ---- one long sequence through the pipe streams --------------------------------------------------------------------------
// Chunk size for a sequence of n bytes: a multiple of 48 (whole runs, whole codons, whole 16-byte granules), small enough
// that the pipeline fills quickly, at most the context's chunk size.
static uint64_t long_chunk_nt(const qd_ctx* c, uint64_t n, uint64_t chunk_bytes = 0) {
    uint64_t ch = std::min<uint64_t>(chunk_bytes ? chunk_bytes : c->chunk_bytes, std::max<uint64_t>(uint64_t(4) << 20, n / 16));
    return std::max<uint64_t>(RUN_NT, ch / RUN_NT * RUN_NT);
}
static bool long_call(const qd_ctx* c, uint64_t n) {
    return n >= std::min<uint64_t>(uint64_t(8) << 20, 2 * (uint64_t)c->chunk_bytes) && n >= 2 * long_chunk_nt(c, n);
}

// Frames of the codon starts in [lo, hi) of ONE n-nt host sequence (lo a multiple of 48, hi a multiple of 48 or n), chunk by
// chunk: a chunk [a, b) is translated as the stand-alone sequence dna[a, min(n, b + 2)) -- the 2-nt halo completes the codons
// that start before b -- and its frames are copied to their place in the global frames: forward local frame k is global
// frame k from residue a / 3 on; reverse local frame k is global reverse frame (k + d) % 3 from residue (k + d) / 3 on,
// d = nucleotides to the right of the piece.  frame_at[f] = where global frame f starts in `out`.
// Error keys are global positions.  The caller resets / fetches the error key and synchronises the pipe streams.
static int frames_long_pipeline(qd_ctx* c, int slot, int enc, int strict, int frames, const uint8_t* dna, uint64_t n, uint64_t lo, uint64_t hi,
                                uint8_t* out, const size_t frame_at[6]) {
    Stager st(c, dna + lo, out, 2 * (hi - lo));
    const uint64_t chunk = long_chunk_nt(c, n, st.chunk_bytes());
    int turn = 0;
    for (uint64_t a = lo; a < hi; a += chunk) {
        const uint64_t b = std::min(hi, a + chunk), read_end = std::min<uint64_t>(n, b + 2), n_local = read_end - a, d = n - read_end;
        const int pi = turn++ % N_PIPE;
        Pipe& P = c->pipe[pi];
        cudaStream_t s = P.stream;
        tls_stream = s;
        if (int rc = st.begin(pi)) return rc;
        const size_t slot_bytes = qd_slots_bytes(n_local, frames);
        QD_CUDA(c, P.in.reserve(n_local + 16));
        QD_CUDA(c, P.out.reserve(slot_bytes + 16));
        if (int rc = st.h2d(pi, P.in.p, dna + a, n_local)) return rc;
        if (int rc = frames_uniform_dev(c, slot, enc, strict, frames, P.in.as<uint8_t>(), 1, n_local, P.out.as<uint8_t>(), s, a)) return rc;
        const int n_slots = frames == QD_FRAMES_ONE ? 1 : frames;
        if (int rc = st.d2h_reserve(pi, slot_bytes, 6)) return rc;
        for (int f = 0; f < n_slots; f++) {
            const size_t len = qd_frame_len(n_local, f % 3);
            if (!len) continue;
            const uint8_t* src = P.out.as<uint8_t>() + qd_slot_frame_offset(n_local, frames, f);
            uint8_t* dst;
            if (f < 3) dst = out + frame_at[f] + a / 3;
            else dst = out + frame_at[3 + (f - 3 + d) % 3] + (f - 3 + d) / 3;
            if (int rc = st.d2h(pi, dst, src, len)) return rc;
        }
        if (int rc = st.end(pi)) return rc;
    }
    tls_stream = c->stream;
    return st.finish(turn);
}

// out[n - b, n - a) = reverse complement of dna[a, b) for the chunks [a, b) of [lo, hi)
static int revcomp_long_pipeline(qd_ctx* c, int which, const uint8_t* dna, uint64_t n, uint64_t lo, uint64_t hi, uint8_t* out) {
    Stager st(c, dna + lo, out, 2 * (hi - lo));
    const uint64_t chunk = long_chunk_nt(c, n, st.chunk_bytes());
    int turn = 0;
    for (uint64_t a = lo; a < hi; a += chunk) {
        const uint64_t b = std::min(hi, a + chunk), len = b - a;
        const int pi = turn++ % N_PIPE;
        Pipe& P = c->pipe[pi];
        cudaStream_t s = P.stream;
        tls_stream = s;
        if (int rc = st.begin(pi)) return rc;
        QD_CUDA(c, P.in.reserve(len + 16));
        QD_CUDA(c, P.out.reserve(len + 16));
        if (int rc = st.h2d(pi, P.in.p, dna + a, len)) return rc;
        if (int rc = revcomp_dev_table(c, which, P.in.as<uint8_t>(), len, P.out.as<uint8_t>(), s, nullptr, a)) return rc;
        if (int rc = st.d2h_reserve(pi, len)) return rc;
        if (int rc = st.d2h(pi, out + (n - b), P.out.p, len)) return rc;
        if (int rc = st.end(pi)) return rc;
    }
    tls_stream = c->stream;
    return st.finish(turn);
}

static int pipes_begin(qd_ctx* c) {
    for (int i = 0; i < N_PIPE; i++) QD_CUDA(c, cudaStreamSynchronize(c->pipe[i].stream));
    if (int rc = reset_err_key(c, c->stream)) return rc;
    QD_CUDA(c, cudaStreamSynchronize(c->stream));
    return QD_OK;
}
static int pipes_end(qd_ctx* c, unsigned long long* key) {
    for (int i = 0; i < N_PIPE; i++) QD_CUDA(c, cudaStreamSynchronize(c->pipe[i].stream));
    return fetch_err_key(c, c->stream, key);
}

// frame lengths / count / positions of the tight output f0|f1|f2[|r0|r1|r2] of an n-nt sequence
static int tight_frames(size_t n, int frames, size_t frame_at[6], size_t out_len[6]) {
    const int n_slots = frames == QD_FRAMES_ONE ? 1 : frames;
    size_t off = 0;
    int nf = 0;
    for (int f = 0; f < 6; f++) {
        const size_t len = f < n_slots ? qd_frame_len(n, f % 3) : 0;
        frame_at[f] = off;
        if (out_len) out_len[f] = len;
        off += len;
        nf += len != 0;
    }
    return nf;
}

extern "C" int qd_translate(qd_ctx* c, int table, int enc, int strict, const uint8_t* dna, size_t n, uint8_t* out, qd_err* err) {
    size_t lens[6];
    int nf = 0;
    if (!c) return QD_ERR_INVALID_ARGUMENT;
    // same path as the frames entry point with one frame; the reference returns an empty Vec for n < 3
    return qd_translate_frames(c, table, enc, strict, QD_FRAMES_ONE, dna, n, out, lens, &nf, err);
}

extern "C" int qd_translate_frames(qd_ctx* c, int table, int enc, int strict, int frames, const uint8_t* dna, size_t n, uint8_t* out,
                                   size_t out_len[6], int* n_frames, qd_err* err) {
    clear_err(err);
    if (!c || !valid_frames(frames) || (n && !dna)) return QD_ERR_INVALID_ARGUMENT;
    int slot;
    if (int rc = check_table(table, err, &slot)) return rc;  // table first: src/python_api.rs:35,48
    for (int k = 0; k < 6; k++)
        if (out_len) out_len[k] = 0;
    if (n_frames) *n_frames = 0;
    Guard g(c);
    if (!g.ok) return QD_ERR_CUDA;
    if (n == 0) return QD_OK;
    cudaStream_t s = c->stream;
    if (long_call(c, n)) {
        // a chromosome-sized sequence: chunks cut at multiples of 48 with a 2-nt halo, H2D / kernel / D2H overlapped on the
        // pipe streams, each chunk's frames copied straight to their place in the caller's buffer
        size_t frame_at[6];
        const int nf = tight_frames(n, frames, frame_at, out_len);
        if (int rc = pipes_begin(c)) return rc;
        if (int rc = frames_long_pipeline(c, slot, enc, strict, frames, dna, n, 0, n, out, frame_at)) return rc;
        unsigned long long key = 0;
        if (int rc = pipes_end(c, &key)) return rc;
        if (key != NO_ERROR_KEY) return report_single(enc, strict, dna, key, err);
        if (n_frames) *n_frames = nf;
        return QD_OK;
    }
    const size_t slot_bytes = qd_slots_bytes(n, frames);
    const int n_slots = frames == QD_FRAMES_ONE ? 1 : frames;
    if (c->tiny_calls && n <= TINY_IN_BYTES && slot_bytes <= TINY_OUT_BYTES) {
        // tiny call: input, error key and result in mapped host memory -- one launch, no copies
        uint8_t* const h_in = c->h_tiny;
        unsigned long long* const h_key = reinterpret_cast<unsigned long long*>(c->h_tiny + TINY_IN_BYTES + 16);
        uint8_t* const h_out = c->h_tiny + TINY_IN_BYTES + 64;
        memcpy(h_in, dna, n);
        *h_key = NO_ERROR_KEY;
        if (int rc = frames_uniform_dev(c, slot, enc, strict, frames, h_in, 1, n, h_out, s, 0, h_key)) return rc;
        QD_CUDA(c, cudaStreamSynchronize(s));
        const unsigned long long key = *h_key;
        if (key != NO_ERROR_KEY) return report_single(enc, strict, dna, key, err);
        size_t off = 0;
        int nf = 0;
        for (int f = 0; f < n_slots; f++) {
            const size_t len = qd_frame_len(n, f % 3);
            if (out_len) out_len[f] = len;
            if (len) {
                memcpy(out + off, h_out + qd_slot_frame_offset(n, frames, f), len);
                off += len;
                nf++;
            }
        }
        if (n_frames) *n_frames = nf;
        return QD_OK;
    }
    QD_CUDA(c, c->in.reserve(n + 16));
    QD_CUDA(c, c->out.reserve(slot_bytes + 16));
    if (int rc = staged_upload(c, c->in.p, dna, n, s)) return rc;
    if (slot_bytes <= SMALL_CALL_BYTES) {
        // small call: the error key lives right behind the slots, so result and key come back in one copy
        unsigned long long* key_dev = reinterpret_cast<unsigned long long*>(c->out.as<uint8_t>() + slot_bytes);  // 16-byte aligned
        QD_CUDA(c, cudaMemsetAsync(key_dev, 0xff, 8, s));
        if (int rc = frames_uniform_dev(c, slot, enc, strict, frames, c->in.as<uint8_t>(), 1, n, c->out.as<uint8_t>(), s, 0, key_dev)) return rc;
        QD_CUDA(c, cudaMemcpyAsync(c->h_stage, c->out.p, slot_bytes + 8, cudaMemcpyDeviceToHost, s));
        QD_CUDA(c, cudaStreamSynchronize(s));
        unsigned long long key;
        memcpy(&key, c->h_stage + slot_bytes, 8);
        if (key != NO_ERROR_KEY) return report_single(enc, strict, dna, key, err);
        size_t off = 0;
        int nf = 0;
        for (int f = 0; f < n_slots; f++) {
            const size_t len = qd_frame_len(n, f % 3);
            if (out_len) out_len[f] = len;
            if (len) {
                memcpy(out + off, c->h_stage + qd_slot_frame_offset(n, frames, f), len);
                off += len;
                nf++;
            }
        }
        if (n_frames) *n_frames = nf;
        return QD_OK;
    }
    if (int rc = reset_err_key(c, s)) return rc;
    if (int rc = frames_uniform_dev(c, slot, enc, strict, frames, c->in.as<uint8_t>(), 1, n, c->out.as<uint8_t>(), s)) return rc;
    // gather the frames tightly: f0|f1|f2[|r0|r1|r2]
    size_t off = 0;
    int nf = 0;
    for (int f = 0; f < n_slots; f++) {
        const size_t len = qd_frame_len(n, f % 3);
        if (out_len) out_len[f] = len;
        if (len) {
            QD_CUDA(c, cudaMemcpyAsync(out + off, c->out.as<uint8_t>() + qd_slot_frame_offset(n, frames, f), len, cudaMemcpyDeviceToHost, s));
            off += len;
            nf++;
        }
    }
    unsigned long long key = 0;
    if (int rc = fetch_err_key(c, s, &key)) return rc;
    if (key != NO_ERROR_KEY) return report_single(enc, strict, dna, key, err);
    if (n_frames) *n_frames = nf;
    return QD_OK;
}

extern "C" int qd_revcomp(qd_ctx* c, int enc, int strict, const uint8_t* dna, size_t n, uint8_t* out, qd_err* err) {
    clear_err(err);
    if (!c || (n && (!dna || !out))) return QD_ERR_INVALID_ARGUMENT;
    Guard g(c);
    if (!g.ok) return QD_ERR_CUDA;
    if (n == 0) return QD_OK;
    cudaStream_t s = c->stream;
    const int which = enc == QD_ENC_ASCII ? (strict ? RC_PAIR_ASCII_STRICT : RC_PAIR_ASCII) : (strict ? RC_PAIR_MASK_STRICT : RC_PAIR_MASK);
    if (long_call(c, n)) {  // chromosome-sized: chunked, copies and kernels overlapped on the pipe streams
        if (int rc = pipes_begin(c)) return rc;
        if (int rc = revcomp_long_pipeline(c, which, dna, n, 0, n, out)) return rc;
        unsigned long long key = 0;
        if (int rc = pipes_end(c, &key)) return rc;
        if (key != NO_ERROR_KEY) return report_single(enc, strict, dna, key, err);
        return QD_OK;
    }
    if (c->tiny_calls && n <= TINY_IN_BYTES) {  // tiny call: input, error key and result in mapped host memory, one launch
        uint8_t* const h_in = c->h_tiny;
        unsigned long long* const h_key = reinterpret_cast<unsigned long long*>(c->h_tiny + TINY_IN_BYTES + 16);
        uint8_t* const h_out = c->h_tiny + TINY_IN_BYTES + 64;
        memcpy(h_in, dna, n);
        *h_key = NO_ERROR_KEY;
        if (int rc = revcomp_dev_table(c, which, h_in, n, h_out, s, h_key)) return rc;
        QD_CUDA(c, cudaStreamSynchronize(s));
        const unsigned long long key = *h_key;
        if (key != NO_ERROR_KEY) return report_single(enc, strict, dna, key, err);
        memcpy(out, h_out, n);
        return QD_OK;
    }
    QD_CUDA(c, c->in.reserve(n + 16));
    QD_CUDA(c, c->out.reserve(n + 16));
    if (int rc = staged_upload(c, c->in.p, dna, n, s)) return rc;
    if (n <= SMALL_CALL_BYTES) {  // small call: [result | error key] in one copy through pinned memory
        const size_t key_off = (n + 15) & ~(size_t)15;
        QD_CUDA(c, c->out.reserve(key_off + 16));
        unsigned long long* key_dev = reinterpret_cast<unsigned long long*>(c->out.as<uint8_t>() + key_off);
        QD_CUDA(c, cudaMemsetAsync(key_dev, 0xff, 8, s));
        if (int rc = revcomp_dev_table(c, which, c->in.as<uint8_t>(), n, c->out.as<uint8_t>(), s, key_dev)) return rc;
        QD_CUDA(c, cudaMemcpyAsync(c->h_stage, c->out.p, key_off + 8, cudaMemcpyDeviceToHost, s));
        QD_CUDA(c, cudaStreamSynchronize(s));
        unsigned long long key;
        memcpy(&key, c->h_stage + key_off, 8);
        if (key != NO_ERROR_KEY) return report_single(enc, strict, dna, key, err);
        memcpy(out, c->h_stage, n);
        return QD_OK;
    }
    if (int rc = reset_err_key(c, s)) return rc;
    if (int rc = revcomp_dev_table(c, which, c->in.as<uint8_t>(), n, c->out.as<uint8_t>(), s)) return rc;
    QD_CUDA(c, cudaMemcpyAsync(out, c->out.p, n, cudaMemcpyDeviceToHost, s));
    unsigned long long key = 0;
    if (int rc = fetch_err_key(c, s, &key)) return rc;
    if (key != NO_ERROR_KEY) return report_single(enc, strict, dna, key, err);
    return QD_OK;
}

// CSR offsets of a host call must not decrease (the lengths are differences of unsigned values)
static int check_host_offsets(qd_ctx* c, const uint64_t* offsets, size_t n_seq, qd_err* err) {
    for (size_t i = 0; i < n_seq; i++) {
        if (offsets[i + 1] < offsets[i]) {
            if (c) c->last_error = "offsets of sequence " + std::to_string(i) + " decrease";
            set_err(err, QD_ERR_INVALID_ARGUMENT, 0, 0, i);
            return QD_ERR_INVALID_ARGUMENT;
        }
    }
    return QD_OK;
}

// chunked H2D -> kernel -> D2H pipeline over whole sequences, N_PIPE streams deep
static int batch_pipeline(qd_ctx* c, int slot, int enc, int strict, int frames, const uint8_t* dna, const uint64_t* offsets,
                          size_t n_seq, size_t uniform_len, uint8_t* out, uint64_t* out_offsets, qd_err* err) {
    const bool uniform = offsets == nullptr;
    const uint64_t total_nt = uniform ? (uint64_t)n_seq * uniform_len : offsets[n_seq] - offsets[0];
    if (n_seq == 0) return QD_OK;
    const uint64_t stride = 16 * (uint64_t)frames;
    if (uniform && c->tiny_calls && total_nt <= TINY_IN_BYTES && n_seq * stride * qd_runs(uniform_len) <= TINY_OUT_BYTES) {
        // a tiny uniform batch: input, error key and slots in mapped host memory -- one launch, no copies (see TINY_IN_BYTES)
        const size_t out_bytes = n_seq * stride * qd_runs(uniform_len);
        uint8_t* const h_in = c->h_tiny;
        unsigned long long* const h_key = reinterpret_cast<unsigned long long*>(c->h_tiny + TINY_IN_BYTES + 16);
        uint8_t* const h_out = c->h_tiny + TINY_IN_BYTES + 64;
        cudaStream_t s = c->stream;
        memcpy(h_in, dna, total_nt);
        *h_key = NO_ERROR_KEY;
        if (int rc = frames_uniform_dev(c, slot, enc, strict, frames, h_in, n_seq, uniform_len, h_out, s, 0, h_key)) return rc;
        QD_CUDA(c, cudaStreamSynchronize(s));
        const unsigned long long key = *h_key;
        if (key != NO_ERROR_KEY) {
            uint8_t payload = 0;
            int kind = classify_byte(enc, strict, dna[key], &payload);
            if (kind == QD_OK) kind = QD_ERR_BAD_NUCLEOTIDE;
            return set_err(err, kind, payload, key % uniform_len, key / uniform_len);
        }
        memcpy(out, h_out, out_bytes);
        return QD_OK;
    }
    // out_offsets (host) for the CSR form: 16*frames*prefix(runs)
    std::vector<uint64_t> own_offsets;
    const uint64_t* oo = out_offsets;
    if (!uniform) {
        if (!out_offsets) {
            own_offsets.resize(n_seq + 1);
            out_offsets = own_offsets.data();
        }
        uint64_t acc = 0;
        for (size_t i = 0; i < n_seq; i++) {
            out_offsets[i] = acc;
            acc += stride * qd_runs((size_t)(offsets[i + 1] - offsets[i]));
        }
        out_offsets[n_seq] = acc;
        oo = out_offsets;
        const size_t offs_at = ((size_t)total_nt + 15) & ~(size_t)15;
        if (c->tiny_calls && offs_at + (n_seq + 1) * 8 <= TINY_IN_BYTES && acc <= TINY_OUT_BYTES) {
            // a tiny variable-length batch: nucleotides, offsets, error key and slots in mapped host memory -- the scan, the
            // map kernels and the frame kernel run on them as they are: no copies, one synchronisation
            uint8_t* const h_in = c->h_tiny;
            uint64_t* const h_offs = reinterpret_cast<uint64_t*>(c->h_tiny + offs_at);
            unsigned long long* const h_key = reinterpret_cast<unsigned long long*>(c->h_tiny + TINY_IN_BYTES + 16);
            uint8_t* const h_out = c->h_tiny + TINY_IN_BYTES + 64;
            cudaStream_t s = c->stream;
            memcpy(h_in, dna, total_nt);
            for (size_t i = 0; i <= n_seq; i++) h_offs[i] = offsets[i] - offsets[0];
            *h_key = NO_ERROR_KEY;
            if (int rc = frames_csr_dev(c, slot, enc, strict, frames, h_in, total_nt, h_offs, n_seq, c->run_prefix, c->tile_first, c->block_sums, h_out, s,
                                        0, 0, h_key))
                return rc;
            QD_CUDA(c, cudaStreamSynchronize(s));
            const unsigned long long key = *h_key;
            if (key != NO_ERROR_KEY) {
                const uint64_t seq = (uint64_t)(std::upper_bound(offsets, offsets + n_seq + 1, (uint64_t)key + offsets[0]) - offsets) - 1;
                uint8_t payload = 0;
                int kind = classify_byte(enc, strict, dna[key], &payload);
                if (kind == QD_OK) kind = QD_ERR_BAD_NUCLEOTIDE;
                return set_err(err, kind, payload, key + offsets[0] - offsets[seq], seq);
            }
            memcpy(out, h_out, acc);
            return QD_OK;
        }
    }
    for (int i = 0; i < N_PIPE; i++) QD_CUDA(c, cudaStreamSynchronize(c->pipe[i].stream));
    if (int rc = reset_err_key(c, c->stream)) return rc;
    QD_CUDA(c, cudaStreamSynchronize(c->stream));

    size_t r0 = 0;
    int turn = 0;
    const uint64_t uni_out = uniform ? stride * qd_runs(uniform_len) : 0;
    Stager st(c, dna, out, total_nt + (uniform ? (uint64_t)n_seq * uni_out : oo[n_seq]));
    const size_t chunk_bytes = st.chunk_bytes();
    while (r0 < n_seq) {
        // pick [r0, r1): whole sequences, about chunk_bytes of input (at least one sequence)
        size_t r1;
        if (uniform) {
            const size_t per = std::max<size_t>(1, uniform_len ? chunk_bytes / uniform_len : n_seq);
            r1 = std::min(n_seq, r0 + per);
        } else {
            const uint64_t target = offsets[r0] + chunk_bytes;
            r1 = (size_t)(std::upper_bound(offsets + r0 + 1, offsets + n_seq + 1, target) - offsets) - 1;
            r1 = std::min(n_seq, std::max(r1, r0 + 1));
        }
        const size_t cnt = r1 - r0;
        const uint64_t in0 = uniform ? (uint64_t)r0 * uniform_len : offsets[r0];
        const uint64_t in_bytes = uniform ? (uint64_t)cnt * uniform_len : offsets[r1] - offsets[r0];
        const uint64_t out0 = uniform ? (uint64_t)r0 * uni_out : oo[r0];
        const uint64_t out_bytes = uniform ? (uint64_t)cnt * uni_out : oo[r1] - oo[r0];
        const int pi = turn % N_PIPE;
        Pipe& P = c->pipe[pi];
        turn++;
        cudaStream_t s = P.stream;
        tls_stream = s;  // this pipe's buffers grow on its own stream
        if (int rc = st.begin(pi)) return rc;
        QD_CUDA(c, P.in.reserve(in_bytes + 16));
        QD_CUDA(c, P.out.reserve(out_bytes + 16));
        if (in_bytes)
            if (int rc = st.h2d(pi, P.in.p, dna + (uniform ? in0 : in0 - offsets[0]), in_bytes)) return rc;
        int rc;
        if (uniform) {
            rc = frames_uniform_dev(c, slot, enc, strict, frames, P.in.as<uint8_t>(), cnt, uniform_len, P.out.as<uint8_t>(), s, in0);
        } else {
            QD_CUDA(c, P.offsets.reserve((cnt + 1) * sizeof(uint64_t)));
            QD_CUDA(c, cudaMemcpyAsync(P.offsets.p, offsets + r0, (cnt + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
            // the chunk's device buffer starts at the sequence whose offset is in0
            rc = frames_csr_dev(c, slot, enc, strict, frames, P.in.as<uint8_t>(), in_bytes, P.offsets.as<uint64_t>(), cnt, P.run_prefix,
                                P.tile_first, P.block_sums, P.out.as<uint8_t>(), s, in0, in0 - offsets[0]);
        }
        if (rc) return rc;
        if (out_bytes) {
            if (int rc2 = st.d2h_reserve(pi, out_bytes)) return rc2;
            if (int rc2 = st.d2h(pi, out + out0, P.out.p, out_bytes)) return rc2;
        }
        if (int rc2 = st.end(pi)) return rc2;
        r0 = r1;
    }
    if (int rc = st.finish(turn)) return rc;
    for (int i = 0; i < N_PIPE; i++) QD_CUDA(c, cudaStreamSynchronize(c->pipe[i].stream));
    tls_stream = c->stream;
    unsigned long long key = 0;
    if (int rc = fetch_err_key(c, c->stream, &key)) return rc;
    if (key != NO_ERROR_KEY) {
        uint64_t seq, pos;
        if (uniform) {
            seq = key / uniform_len;
            pos = key % uniform_len;
        } else {
            seq = (uint64_t)(std::upper_bound(offsets, offsets + n_seq + 1, (uint64_t)key + offsets[0]) - offsets) - 1;
            pos = key + offsets[0] - offsets[seq];
        }
        uint8_t payload = 0;
        int kind = classify_byte(enc, strict, dna[key], &payload);
        if (kind == QD_OK) kind = QD_ERR_BAD_NUCLEOTIDE;
        return set_err(err, kind, payload, pos, seq);
    }
    (void)total_nt;
    return QD_OK;
}

extern "C" int qd_translate_frames_batch(qd_ctx* c, int table, int enc, int strict, int frames, const uint8_t* dna,
                                         const uint64_t* offsets, size_t n_seq, uint8_t* out, uint64_t* out_offsets, qd_err* err) {
    clear_err(err);
    if (!c || !valid_frames(frames) || !offsets) return QD_ERR_INVALID_ARGUMENT;
    int slot;
    if (int rc = check_table(table, err, &slot)) return rc;
    if (int rc = check_host_offsets(c, offsets, n_seq, err)) return rc;
    Guard g(c);
    if (!g.ok) return QD_ERR_CUDA;
    return batch_pipeline(c, slot, enc, strict, frames, dna, offsets, n_seq, 0, out, out_offsets, err);
}

extern "C" int qd_translate_frames_batch_uniform(qd_ctx* c, int table, int enc, int strict, int frames, const uint8_t* dna,
                                                 size_t n_seq, size_t seq_len, uint8_t* out, qd_err* err) {
    clear_err(err);
    if (!c || !valid_frames(frames)) return QD_ERR_INVALID_ARGUMENT;
    int slot;
    if (int rc = check_table(table, err, &slot)) return rc;
    Guard g(c);
    if (!g.ok) return QD_ERR_CUDA;
    if (seq_len == 0) return QD_OK;
    return batch_pipeline(c, slot, enc, strict, frames, dna, nullptr, n_seq, seq_len, out, nullptr, err);
}

extern "C" int qd_canonical_windows(qd_ctx* c, int enc, int forward_only, const uint8_t* dna, size_t n, size_t w, uint8_t* out,
                                    qd_err* err) {
    clear_err(err);
    if (!c || w < 1 || w > CANON_MAX_W) return QD_ERR_INVALID_ARGUMENT;
    Guard g(c);
    if (!g.ok) return QD_ERR_CUDA;
    if (n < w) return QD_OK;
    cudaStream_t s = c->stream;
    const size_t out_bytes = (n - w + 1) * w;
    if (small_result(out_bytes)) {
        unsigned long long key = 0;
        if (int rc = small_call(c, dna, n, out, out_bytes, &key, [&](const uint8_t* d_in, uint8_t* d_out, cudaStream_t st) {
                return canonical_dev(c, enc, forward_only, d_in, 1, n, w, d_out, st);
            }))
            return rc;
        if (key != NO_ERROR_KEY) return report_single(enc, /*strict=*/1, dna, key, err);
        return QD_OK;
    }
    QD_CUDA(c, c->in.reserve(n + 16));
    QD_CUDA(c, c->out.reserve(out_bytes + 16));
    if (int rc = staged_upload(c, c->in.p, dna, n, s)) return rc;
    if (int rc = reset_err_key(c, s)) return rc;
    if (int rc = canonical_dev(c, enc, forward_only, c->in.as<uint8_t>(), 1, n, w, c->out.as<uint8_t>(), s)) return rc;
    if (int rc = staged_download(c, out, c->out.p, out_bytes, s)) return rc;
    unsigned long long key = 0;
    if (int rc = fetch_err_key(c, s, &key)) return rc;
    if (key != NO_ERROR_KEY) return report_single(enc, /*strict=*/1, dna, key, err);
    return QD_OK;
}

extern "C" int qd_canonical_window_keys(qd_ctx* c, int enc, int forward_only, int kind, const uint8_t* dna, size_t n, size_t w, void* keys,
                                        qd_err* err) {
    clear_err(err);
    const size_t key_bytes = qd_canonical_key_bytes(kind, w);
    if (!c || key_bytes == 0) return QD_ERR_INVALID_ARGUMENT;
    Guard g(c);
    if (!g.ok) return QD_ERR_CUDA;
    if (n < w) return QD_OK;
    cudaStream_t s = c->stream;
    const size_t out_bytes = (n - w + 1) * key_bytes;
    if (small_result(out_bytes)) {
        unsigned long long key = 0;
        if (int rc = small_call(c, dna, n, static_cast<uint8_t*>(keys), out_bytes, &key, [&](const uint8_t* d_in, uint8_t* d_out, cudaStream_t st) {
                return canonical_keys_dev(c, enc, forward_only, kind, d_in, 1, n, w, d_out, st);
            }))
            return rc;
        if (key != NO_ERROR_KEY) return report_single(enc, /*strict=*/1, dna, key, err);
        return QD_OK;
    }
    QD_CUDA(c, c->in.reserve(n + 16));
    QD_CUDA(c, c->out.reserve(out_bytes + 16));
    if (int rc = staged_upload(c, c->in.p, dna, n, s)) return rc;
    if (int rc = reset_err_key(c, s)) return rc;
    if (int rc = canonical_keys_dev(c, enc, forward_only, kind, c->in.as<uint8_t>(), 1, n, w, c->out.p, s)) return rc;
    if (int rc = staged_download(c, static_cast<uint8_t*>(keys), c->out.p, out_bytes, s)) return rc;
    unsigned long long key = 0;
    if (int rc = fetch_err_key(c, s, &key)) return rc;
    if (key != NO_ERROR_KEY) return report_single(enc, /*strict=*/1, dna, key, err);
    return QD_OK;
}

static bool expansion_fast_path(size_t w, uint64_t max_expansions) { return w <= 64 && max_expansions < (uint64_t(1) << (EXP_DESC_MAX + 1)); }

static ExpandParams expansion_params(qd_ctx* c, int enc, const uint8_t* windows_dev, size_t n_win, size_t w, uint64_t max_expansions,
                                     uint64_t* counts_dev, uint64_t* emit_dev, uint64_t* out_index_dev, uint64_t* desc_dev) {
    ExpandParams p{};
    p.windows = windows_dev;
    p.n_win = n_win;
    p.w = (uint32_t)w;
    p.max_expansions = max_expansions;
    p.to_mask = tab(c, enc == QD_ENC_ASCII ? TAB_TO_MASK_ASCII : TAB_TO_MASK_MASK);
    p.letters = enc == QD_ENC_ASCII ? ('A' | ('T' << 8) | ('C' << 16) | ((uint32_t)'G' << 24)) : (1u | (2u << 8) | (4u << 16) | (8u << 24));
    p.counts = counts_dev;
    p.emit = emit_dev;
    p.out_index = out_index_dev;
    p.desc = desc_dev;
    p.err_key = c->d_err;
    return p;
}

// size_hint per window + exclusive scan of the rows each window emits
static int expansion_count_dev(qd_ctx* c, const ExpandParams& p, int enc, cudaStream_t s) {
    if (p.desc && p.w <= 64) {  // fast path: warp tiles of 32 windows, four letters per step
        const uint64_t tiles = (p.n_win + 31) / 32;
        const unsigned grid = (unsigned)std::min<uint64_t>((tiles + EXP_WARPS - 1) / EXP_WARPS, (uint64_t)c->sm_count * 8);
        if (enc == QD_ENC_ASCII) {
            if (p.w == 42) expansion_count_tiles_kernel<true, 42><<<grid, 32 * EXP_WARPS, 0, s>>>(p, c->d_parse_pair);  // benches/expansions.rs: unrolled
            else expansion_count_tiles_kernel<true, 0><<<grid, 32 * EXP_WARPS, 0, s>>>(p, c->d_parse_pair);
        } else {
            if (p.w == 42) expansion_count_tiles_kernel<false, 42><<<grid, 32 * EXP_WARPS, 0, s>>>(p, nullptr);
            else expansion_count_tiles_kernel<false, 0><<<grid, 32 * EXP_WARPS, 0, s>>>(p, nullptr);
        }
    } else {
        const unsigned grid = (unsigned)std::min<uint64_t>((p.n_win + 255) / 256, (uint64_t)c->sm_count * 16);
        expansion_count_kernel<<<grid, 256, 0, s>>>(p);
    }
    c->launches++;
    QD_CUDA(c, cudaGetLastError());
    return device_exclusive_scan<1>(c, p.emit, p.n_win, const_cast<uint64_t*>(p.out_index), c->block_sums, s);
}

static int expansion_write_tiles_dev(qd_ctx* c, ExpandParams p, uint8_t* out_dev, uint64_t row_cap, cudaStream_t s) {
    p.out = out_dev;
    p.row_cap = row_cap;
    const uint64_t tiles = (p.n_win + 31) / 32;
    const unsigned grid = (unsigned)std::min<uint64_t>((tiles + EXP_WARPS - 1) / EXP_WARPS, (uint64_t)c->sm_count * 6);
    if (p.w == 42) expansion_write_tiles_kernel<42><<<grid, 32 * EXP_WARPS, 0, s>>>(p);  // the window length of benches/expansions.rs: unrolled
    else expansion_write_tiles_kernel<0><<<grid, 32 * EXP_WARPS, 0, s>>>(p);
    c->launches++;
    QD_CUDA(c, cudaGetLastError());
    return QD_OK;
}

// one pass (device callers with an output buffer): size_hint, row index and rows from one read of the windows;
// `state` = ticket counter + one status word per warp tile of the chained scan, zeroed here
template <bool ASCII, int W>
static int launch_expansion_fused(qd_ctx* c, const ExpandParams& p, uint64_t* state, cudaStream_t s, int kernel_id) {
    auto kern = expansion_fused_tiles_kernel<ASCII, W>;
    const size_t smem = sizeof(ExpPipeSmem<(W ? W : 64)>);
    auto it = c->occ_cache.find(((uint64_t)0xE0 + kernel_id) << 40);
    if (it == c->occ_cache.end()) {
        int occ = 0;
        QD_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        QD_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * EXP_FUSED_WARPS, smem));
        it = c->occ_cache.emplace(((uint64_t)0xE0 + kernel_id) << 40, std::max(1, occ)).first;
    }
    const uint64_t tiles = (p.n_win + 31) / 32;
    const unsigned grid = (unsigned)std::min<uint64_t>((tiles + EXP_FUSED_WARPS - 1) / EXP_FUSED_WARPS, (uint64_t)c->sm_count * it->second);
    QD_CUDA(c, cudaMemsetAsync(state, 0, (tiles + 1) * 8, s));  // ticket counter + one status word per warp tile
    kern<<<grid, 32 * EXP_FUSED_WARPS, smem, s>>>(p, ASCII ? c->d_parse_pair : nullptr, state);
    c->launches++;
    QD_CUDA(c, cudaGetLastError());
    return QD_OK;
}

static int expansion_fused_dev(qd_ctx* c, ExpandParams p, int enc, uint8_t* out_dev, uint64_t row_cap, uint64_t* state, cudaStream_t s) {
    p.out = out_dev;
    p.row_cap = row_cap;
    if (enc == QD_ENC_ASCII) return p.w == 42 ? launch_expansion_fused<true, 42>(c, p, state, s, 0) : launch_expansion_fused<true, 0>(c, p, state, s, 1);
    return p.w == 42 ? launch_expansion_fused<false, 42>(c, p, state, s, 2) : launch_expansion_fused<false, 0>(c, p, state, s, 3);
}

extern "C" int qd_expansion_windows_dev(qd_ctx* c, int enc, const uint8_t* windows_dev, size_t n_win, size_t w, uint64_t max_expansions,
                                        uint64_t* counts_dev, uint64_t* out_index_dev, uint8_t* out_dev, uint64_t out_capacity_rows,
                                        void* stream) {
    if (!c || w < 1 || !expansion_fast_path(w, max_expansions) || !counts_dev || !out_index_dev) return QD_ERR_INVALID_ARGUMENT;
    Guard g(c, pick_stream(c, stream));
    if (!g.ok) return QD_ERR_CUDA;
    if (n_win == 0) return QD_OK;
    cudaStream_t s = pick_stream(c, stream);
    QD_CUDA(c, c->aux1.reserve(((n_win + 31) / 32 + 1) * 8));  // ticket counter + status word per warp tile
    ExpandParams p = expansion_params(c, enc, windows_dev, n_win, w, max_expansions, counts_dev, nullptr, out_index_dev, nullptr);
    return expansion_fused_dev(c, p, enc, out_dev, out_capacity_rows, c->aux1.as<uint64_t>(), s);
}

extern "C" size_t qd_expansion_chunk_bytes(size_t chunk_windows, size_t w, uint64_t max_expansions) {
    return chunk_windows * (size_t)max_expansions * w;
}

extern "C" int qd_expansion_windows_chunked_dev(qd_ctx* c, int enc, const uint8_t* windows_dev, size_t n_win, size_t w, uint64_t max_expansions,
                                                uint8_t* ring_dev, size_t ring_bytes, size_t chunk_windows, uint64_t* counts_dev,
                                                uint64_t* chunk_rows_dev, uint64_t* checksums_dev, qd_chunk_fn consume, void* user,
                                                void* stream) {
    if (!c || w < 1 || !expansion_fast_path(w, max_expansions) || chunk_windows == 0 || !chunk_rows_dev) return QD_ERR_INVALID_ARGUMENT;
    Guard g(c, pick_stream(c, stream));
    if (!g.ok) return QD_ERR_CUDA;
    if (n_win == 0) return QD_OK;
    Ring ring;
    const size_t cw = std::min(chunk_windows, n_win);
    if (!make_ring(ring_dev, ring_bytes, qd_expansion_chunk_bytes(cw, w, max_expansions), &ring)) {
        c->last_error = "ring smaller than one chunk (qd_expansion_chunk_bytes) or not 16-byte aligned";
        return QD_ERR_INVALID_ARGUMENT;
    }
    QD_CUDA(c, c->aux1.reserve(((cw + 31) / 32 + 1) * 8));  // ticket counter + status word per warp tile
    QD_CUDA(c, c->aux2.reserve((cw + 1) * 8));          // row index of the chunk's windows
    if (!counts_dev) QD_CUDA(c, c->aux0.reserve(cw * 8));
    const size_t n_chunks = qd_chunk_count(n_win, chunk_windows);
    if (checksums_dev) QD_CUDA(c, cudaMemsetAsync(checksums_dev, 0, n_chunks * 8, g.s));
    for (size_t k = 0; k < n_chunks; k++) {
        const size_t i0 = k * chunk_windows, cnt = std::min(chunk_windows, n_win - i0);
        uint8_t* out = ring.slot(k);
        ExpandParams p = expansion_params(c, enc, windows_dev + i0 * w, cnt, w, max_expansions, counts_dev ? counts_dev + i0 : c->aux0.as<uint64_t>(),
                                          nullptr, c->aux2.as<uint64_t>(), nullptr);
        if (int rc = expansion_fused_dev(c, p, enc, out, (uint64_t)cnt * max_expansions, c->aux1.as<uint64_t>(), g.s)) return rc;
        QD_CUDA(c, cudaMemcpyAsync(chunk_rows_dev + k, c->aux2.as<uint64_t>() + cnt, 8, cudaMemcpyDeviceToDevice, g.s));
        if (checksums_dev)
            if (int rc = checksum_dev(c, out, (uint64_t)cnt * max_expansions, chunk_rows_dev + k, w, 0, checksums_dev + k, g.s)) return rc;
        if (consume) consume(user, k, i0, cnt, out, (uint64_t)cnt * max_expansions * w, chunk_rows_dev + k, (void*)g.s);
    }
    return QD_OK;
}

extern "C" int qd_expansion_windows(qd_ctx* c, int enc, const uint8_t* windows, size_t n_win, size_t w, uint64_t max_expansions,
                                    uint64_t* counts, uint64_t* out_index, uint8_t* out, qd_err* err) {
    clear_err(err);
    if (!c || w < 1 || w > 0xffffffffull) return QD_ERR_INVALID_ARGUMENT;
    Guard g(c);
    if (!g.ok) return QD_ERR_CUDA;
    if (n_win == 0) {
        if (out_index) out_index[0] = 0;
        return QD_OK;
    }
    cudaStream_t s = c->stream;
    const size_t in_bytes = n_win * w;
    const bool fast = expansion_fast_path(w, max_expansions);
    QD_CUDA(c, c->in.reserve(in_bytes + 16));
    // Small calls (DnaSequence::expansions() of one window, a few hundred windows): the one-pass kernel writes rows, counts,
    // row index and error key into ONE device buffer sized by the cap, which comes back in one copy through pinned memory --
    // no synchronisation between a count and a write pass.
    const size_t row_bound = n_win * (size_t)max_expansions * w;
    const size_t off_c = (row_bound + 15) & ~(size_t)15, off_i = off_c + n_win * 8, off_k = off_i + (n_win + 1) * 8;
    if (fast && out && off_k + 8 <= (size_t(1) << 20)) {
        QD_CUDA(c, c->out.reserve(off_k + 16));
        QD_CUDA(c, c->aux1.reserve(((n_win + 31) / 32 + 1) * 8));  // ticket counter + status word per warp tile
        uint8_t* const d = c->out.as<uint8_t>();
        if (int rc = staged_upload(c, c->in.p, windows, in_bytes, s)) return rc;
        QD_CUDA(c, cudaMemsetAsync(d + off_k, 0xff, 8, s));
        ExpandParams p = expansion_params(c, enc, c->in.as<uint8_t>(), n_win, w, max_expansions, reinterpret_cast<uint64_t*>(d + off_c), nullptr,
                                          reinterpret_cast<uint64_t*>(d + off_i), nullptr);
        p.err_key = reinterpret_cast<unsigned long long*>(d + off_k);
        if (int rc = expansion_fused_dev(c, p, enc, d, n_win * max_expansions, c->aux1.as<uint64_t>(), s)) return rc;
        QD_CUDA(c, cudaMemcpyAsync(c->h_stage, d, off_k + 8, cudaMemcpyDeviceToHost, s));
        QD_CUDA(c, cudaStreamSynchronize(s));
        unsigned long long key;
        memcpy(&key, c->h_stage + off_k, 8);
        if (key != NO_ERROR_KEY) {
            uint8_t payload = 0;
            int kind = classify_byte(enc, 0, windows[key], &payload);
            if (kind == QD_OK) kind = QD_ERR_BAD_NUCLEOTIDE;
            return set_err(err, kind, payload, key % w, key / w);
        }
        const uint64_t* h_counts = reinterpret_cast<const uint64_t*>(c->h_stage + off_c);
        const uint64_t* h_index = reinterpret_cast<const uint64_t*>(c->h_stage + off_i);
        if (counts) memcpy(counts, h_counts, n_win * 8);
        if (out_index) {
            for (size_t i = 0; i < n_win; i++) out_index[i] = h_counts[i] > max_expansions ? ~0ull : h_index[i];  // skipped windows are marked
            out_index[n_win] = h_index[n_win];
        }
        memcpy(out, c->h_stage, (size_t)h_index[n_win] * w);
        return QD_OK;
    }
    QD_CUDA(c, c->aux0.reserve(n_win * 8));        // counts
    QD_CUDA(c, c->aux1.reserve(n_win * 8));        // emit
    QD_CUDA(c, c->aux2.reserve((n_win + 1) * 8));  // out_index
    if (fast) QD_CUDA(c, c->aux3.reserve(n_win * 8));  // descriptors
    if (int rc = staged_upload(c, c->in.p, windows, in_bytes, s)) return rc;
    if (int rc = reset_err_key(c, s)) return rc;
    ExpandParams p = expansion_params(c, enc, c->in.as<uint8_t>(), n_win, w, max_expansions, c->aux0.as<uint64_t>(), c->aux1.as<uint64_t>(),
                                      c->aux2.as<uint64_t>(), fast ? c->aux3.as<uint64_t>() : nullptr);
    if (int rc = expansion_count_dev(c, p, enc, s)) return rc;
    QD_CUDA(c, cudaMemcpyAsync(c->h_scratch, c->aux2.as<uint64_t>() + n_win, 8, cudaMemcpyDeviceToHost, s));
    unsigned long long key = 0;
    if (int rc = fetch_err_key(c, s, &key)) return rc;  // synchronises
    if (key != NO_ERROR_KEY) {
        uint8_t payload = 0;
        int kind = classify_byte(enc, 0, windows[key], &payload);
        if (kind == QD_OK) kind = QD_ERR_BAD_NUCLEOTIDE;
        return set_err(err, kind, payload, key % w, key / w);
    }
    const uint64_t total_rows = c->h_scratch[0];
    std::vector<uint64_t> own_counts;
    if (!counts) {
        own_counts.resize(n_win);
        counts = own_counts.data();
    }
    QD_CUDA(c, cudaMemcpyAsync(counts, c->aux0.p, n_win * 8, cudaMemcpyDeviceToHost, s));
    if (out_index) QD_CUDA(c, cudaMemcpyAsync(out_index, c->aux2.p, (n_win + 1) * 8, cudaMemcpyDeviceToHost, s));
    if (out && total_rows) {
        QD_CUDA(c, c->out.reserve(total_rows * w + 16));
        if (fast) {
            if (int rc = expansion_write_tiles_dev(c, p, c->out.as<uint8_t>(), total_rows, s)) return rc;
        } else {  // any w, any cap: one thread per row
            p.out = c->out.as<uint8_t>();
            const unsigned g2 = (unsigned)std::min<uint64_t>((total_rows + 255) / 256, (uint64_t)c->sm_count * 16);
            expansion_write_kernel<<<g2, 256, 0, s>>>(p, total_rows);
            c->launches++;
            QD_CUDA(c, cudaGetLastError());
        }
        if (int rc = staged_download(c, out, c->out.p, total_rows * w, s)) return rc;
    }
    QD_CUDA(c, cudaStreamSynchronize(s));
    if (out_index) {
        // skipped windows are marked, as the header documents
        // (their device-side index equals the next window's; the mark is host-side only)
        for (size_t i = 0; i < n_win; i++)
            if (counts[i] > max_expansions) out_index[i] = ~0ull;
    }
    return QD_OK;
}

extern "C" int qd_windows(qd_ctx* c, const uint8_t* seq, size_t n, size_t w, uint8_t* out, size_t* n_windows) {
    if (n_windows) *n_windows = 0;
    if (!c || w == 0) return QD_ERR_INVALID_ARGUMENT;
    Guard g(c);
    if (!g.ok) return QD_ERR_CUDA;
    if (n < w) return QD_OK;
    cudaStream_t s = c->stream;
    const size_t count = n - w + 1;
    QD_CUDA(c, c->in.reserve(n + 16));
    QD_CUDA(c, c->out.reserve(count * w + 16));
    if (int rc = staged_upload(c, c->in.p, seq, n, s)) return rc;
    if (int rc = windows_dev(c, c->in.as<uint8_t>(), n, w, c->out.as<uint8_t>(), s)) return rc;
    if (int rc = staged_download(c, out, c->out.p, count * w, s)) return rc;
    if (n_windows) *n_windows = count;
    return QD_OK;
}

extern "C" int qd_windows_all(qd_ctx* c, int enc, const uint8_t* dna, size_t n, size_t w_dna, size_t w_aa, uint8_t* dna_out,
                              uint8_t* aa_out, uint64_t* origin, size_t counts[8], qd_err* err) {
    clear_err(err);
    if (!c || !counts || w_dna == 0 || w_aa == 0) return QD_ERR_INVALID_ARGUMENT;
    for (int i = 0; i < 8; i++) counts[i] = 0;
    Guard g(c);
    if (!g.ok) return QD_ERR_CUDA;
    if (n == 0) return QD_OK;
    cudaStream_t s = c->stream;
    auto nwin = [](size_t len, size_t w) { return len + 1 - std::min(w, len + 1); };  // benches/all_windows.rs:209-211
    const size_t n_dna = nwin(n, w_dna);
    counts[0] = counts[1] = n_dna;
    // frames the reference returns, in list order (`operation` = list index, benches/all_windows.rs:52-72)
    size_t list_len[6];
    int list_frame[6];
    int n_list = 0;
    for (int f = 0; f < 6; f++)
        if (qd_frame_len(n, f % 3)) {
            list_len[n_list] = qd_frame_len(n, f % 3);
            list_frame[n_list] = f;
            n_list++;
        }
    size_t n_aa = 0;
    for (int op = 0; op < n_list; op++) {
        counts[2 + op] = nwin(list_len[op], w_aa);
        n_aa += counts[2 + op];
    }
    const size_t total = 2 * n_dna + n_aa;
    QD_CUDA(c, c->in.reserve(n + 16));
    QD_CUDA(c, c->aux0.reserve(n + 16));                       // forward ASCII
    QD_CUDA(c, c->aux1.reserve(n + 16));                       // reverse-complement ASCII
    QD_CUDA(c, c->aux2.reserve(qd_slots_bytes(n, 6) + 16));    // six frames
    QD_CUDA(c, c->out.reserve(std::max(2 * n_dna * w_dna, n_aa * w_aa) + 16));
    QD_CUDA(c, c->aux3.reserve(total * 8 + 16));
    if (int rc = staged_upload(c, c->in.p, dna, n, s)) return rc;
    if (int rc = reset_err_key(c, s)) return rc;
    // strict alphabet (DnaSequence<Nucleotide>): the strict tables reject everything else
    const int t_rc = enc == QD_ENC_ASCII ? RC_PAIR_ASCII_STRICT : RC_PAIR_MASK_TO_ASCII_STRICT;
    if (int rc = revcomp_dev_table(c, t_rc, c->in.as<uint8_t>(), n, c->aux1.as<uint8_t>(), s)) return rc;
    // forward ASCII (upper case) = reverse complement of the reverse complement
    if (int rc = revcomp_dev_table(c, RC_PAIR_ASCII_STRICT, c->aux1.as<uint8_t>(), n, c->aux0.as<uint8_t>(), s)) return rc;
    if (int rc = frames_uniform_dev(c, 0 /* Ncbi1 */, enc, 1, QD_FRAMES_ALL, c->in.as<uint8_t>(), 1, n, c->aux2.as<uint8_t>(), s)) return rc;
    unsigned long long key = 0;
    if (int rc = fetch_err_key(c, s, &key)) return rc;
    if (key != NO_ERROR_KEY) return report_single(enc, 1, dna, key, err);
    if (dna_out && n_dna) {
        // the two DNA segments are written to separately aligned buffers, then copied out back to back
        if (int rc = windows_dev(c, c->aux0.as<uint8_t>(), n, w_dna, c->out.as<uint8_t>(), s)) return rc;
        QD_CUDA(c, cudaMemcpyAsync(dna_out, c->out.p, n_dna * w_dna, cudaMemcpyDeviceToHost, s));
        QD_CUDA(c, cudaStreamSynchronize(s));
        if (int rc = windows_dev(c, c->aux1.as<uint8_t>(), n, w_dna, c->out.as<uint8_t>(), s)) return rc;
        QD_CUDA(c, cudaMemcpyAsync(dna_out + n_dna * w_dna, c->out.p, n_dna * w_dna, cudaMemcpyDeviceToHost, s));
        QD_CUDA(c, cudaStreamSynchronize(s));
    }
    if (aa_out) {
        size_t written = 0;
        for (int op = 0; op < n_list; op++) {
            if (!counts[2 + op]) continue;
            const uint8_t* frame = c->aux2.as<uint8_t>() + qd_slot_frame_offset(n, 6, list_frame[op]);
            if (int rc = windows_dev(c, frame, list_len[op], w_aa, c->out.as<uint8_t>(), s)) return rc;
            QD_CUDA(c, cudaMemcpyAsync(aa_out + written * w_aa, c->out.p, counts[2 + op] * w_aa, cudaMemcpyDeviceToHost, s));
            QD_CUDA(c, cudaStreamSynchronize(s));
            written += counts[2 + op];
        }
    }
    if (origin && total) {
        uint64_t* o = c->aux3.as<uint64_t>();
        size_t at = 0;
        auto fill = [&](size_t count, int kind, uint64_t f) {
            if (!count) return;
            const unsigned grid = (unsigned)std::min<uint64_t>((count + 255) / 256, (uint64_t)c->sm_count * 8);
            origin_kernel<<<grid, 256, 0, s>>>(o + at, count, kind, f, (uint64_t)n, (uint64_t)w_aa);
            c->launches++;
            at += count;
        };
        fill(n_dna, 0, 0);
        fill(n_dna, 1, 0);
        for (int op = 0; op < n_list; op++) fill(counts[2 + op], op <= 2 ? 2 : 3, (uint64_t)(op % 3));
        QD_CUDA(c, cudaGetLastError());
        QD_CUDA(c, cudaMemcpyAsync(origin, o, total * 8, cudaMemcpyDeviceToHost, s));
        QD_CUDA(c, cudaStreamSynchronize(s));
    }
    return QD_OK;
}

// =================================================================================================
// several GPUs of one box from one call: partition by bytes, one context / host thread / set of pipe
// streams per device, every device copies its results to the right offset of the caller's ONE output buffer.  Nothing is
// exchanged between devices (no collective).
// =================================================================================================
#include <thread>

namespace {
// the error with the lowest (sequence, position) wins -- what a serial pass over the batch would have hit first
void keep_first_error(int rc, const qd_err& e, int* best_rc, qd_err* best) {
    if (rc == QD_OK) return;
    const bool data_err = rc >= QD_ERR_NON_ASCII_BYTE && rc <= QD_ERR_BAD_MASK;
    const bool best_data = *best_rc >= QD_ERR_NON_ASCII_BYTE && *best_rc <= QD_ERR_BAD_MASK;
    if (*best_rc == QD_OK || (!data_err && best_data) ||
        (data_err == best_data && (e.seq_index < best->seq_index || (e.seq_index == best->seq_index && e.pos < best->pos)))) {
        *best_rc = rc;
        *best = e;
    }
}
bool valid_ctxs(qd_ctx* const* ctxs, int n_ctx) {
    if (!ctxs || n_ctx < 1 || n_ctx > 64) return false;
    for (int i = 0; i < n_ctx; i++) {
        if (!ctxs[i]) return false;
        for (int j = 0; j < i; j++)
            if (ctxs[j] == ctxs[i]) return false;
    }
    return true;
}
template <class F>
void run_per_device(int n_ctx, F&& body) {
    std::vector<std::thread> th;
    for (int i = 1; i < n_ctx; i++) th.emplace_back(body, i);
    body(0);
    for (auto& t : th) t.join();
}
}  // namespace

extern "C" int qd_translate_frames_batch_multi(qd_ctx* const* ctxs, int n_ctx, int table, int enc, int strict, int frames, const uint8_t* dna,
                                               const uint64_t* offsets, size_t n_seq, size_t seq_len, uint8_t* out, uint64_t* out_offsets,
                                               qd_err* err) {
    clear_err(err);
    if (!valid_ctxs(ctxs, n_ctx) || !valid_frames(frames)) return QD_ERR_INVALID_ARGUMENT;
    int slot;
    if (int rc = check_table(table, err, &slot)) return rc;
    if (n_seq == 0 || (!offsets && seq_len == 0)) {
        if (offsets && out_offsets) out_offsets[0] = 0;
        return QD_OK;
    }
    if (offsets)
        if (int rc = check_host_offsets(ctxs[0], offsets, n_seq, err)) return rc;
    const uint64_t stride = 16 * (uint64_t)frames;
    // cut points: sequence boundaries, balanced by BYTES
    std::vector<size_t> cut(n_ctx + 1, n_seq);
    cut[0] = 0;
    std::vector<uint64_t> own_oo;
    if (offsets) {
        const uint64_t total = offsets[n_seq] - offsets[0];
        for (int k = 1; k < n_ctx; k++) {
            const uint64_t target = offsets[0] + total / n_ctx * k + total % n_ctx * k / n_ctx;
            size_t j = (size_t)(std::lower_bound(offsets, offsets + n_seq + 1, target) - offsets);  // first boundary >= target
            if (j > 0 && target - offsets[j - 1] < offsets[std::min(j, n_seq)] - target) j--;        // ... or the nearer one before it
            cut[k] = std::min(std::max(j, cut[k - 1]), n_seq);
        }
        if (!out_offsets) {
            own_oo.resize(n_seq + 1);
            out_offsets = own_oo.data();
        }
        uint64_t acc = 0;
        for (size_t i = 0; i < n_seq; i++) {
            out_offsets[i] = acc;
            acc += stride * qd_runs((size_t)(offsets[i + 1] - offsets[i]));
        }
        out_offsets[n_seq] = acc;
    } else {
        for (int k = 1; k < n_ctx; k++) cut[k] = n_seq / n_ctx * k + n_seq % n_ctx * k / n_ctx;
    }
    std::vector<int> rcs(n_ctx, QD_OK);
    std::vector<qd_err> errs(n_ctx);
    const uint64_t uni_out = offsets ? 0 : stride * qd_runs(seq_len);
    run_per_device(n_ctx, [&](int i) {
        clear_err(&errs[i]);
        const size_t r0 = cut[i], cnt = cut[i + 1] - cut[i];
        if (cnt == 0) return;
        qd_ctx* c = ctxs[i];
        Guard g(c);
        if (!g.ok) {
            rcs[i] = QD_ERR_CUDA;
            return;
        }
        if (offsets)
            rcs[i] = batch_pipeline(c, slot, enc, strict, frames, dna + (offsets[r0] - offsets[0]), offsets + r0, cnt, 0, out + out_offsets[r0], nullptr,
                                    &errs[i]);
        else
            rcs[i] = batch_pipeline(c, slot, enc, strict, frames, dna + (uint64_t)r0 * seq_len, nullptr, cnt, seq_len, out + (uint64_t)r0 * uni_out, nullptr,
                                    &errs[i]);
        errs[i].seq_index += r0;
    });
    int best_rc = QD_OK;
    qd_err best{};
    for (int i = 0; i < n_ctx; i++) keep_first_error(rcs[i], errs[i], &best_rc, &best);
    if (best_rc != QD_OK && err) *err = best;
    return best_rc;
}

// the 48-aligned cut points of an n-nt sequence over n_ctx devices
static std::vector<uint64_t> chromosome_cuts(uint64_t n, int n_ctx) {
    const uint64_t runs = (n + RUN_NT - 1) / RUN_NT;
    std::vector<uint64_t> cut(n_ctx + 1);
    for (int k = 0; k <= n_ctx; k++) cut[k] = std::min<uint64_t>(n, (runs / n_ctx * k + runs % n_ctx * k / n_ctx) * RUN_NT);
    cut[n_ctx] = n;
    return cut;
}

extern "C" int qd_translate_frames_multi(qd_ctx* const* ctxs, int n_ctx, int table, int enc, int strict, int frames, const uint8_t* dna, size_t n,
                                         uint8_t* out, size_t out_len[6], int* n_frames, qd_err* err) {
    clear_err(err);
    if (!valid_ctxs(ctxs, n_ctx) || !valid_frames(frames) || (n && (!dna || !out))) return QD_ERR_INVALID_ARGUMENT;
    int slot;
    if (int rc = check_table(table, err, &slot)) return rc;
    size_t frame_at[6];
    const int nf = tight_frames(n, frames, frame_at, out_len);
    if (n_frames) *n_frames = 0;
    if (n == 0) return QD_OK;
    const std::vector<uint64_t> cut = chromosome_cuts(n, n_ctx);
    std::vector<int> rcs(n_ctx, QD_OK);
    std::vector<unsigned long long> keys(n_ctx, NO_ERROR_KEY);
    run_per_device(n_ctx, [&](int i) {
        if (cut[i] >= cut[i + 1]) return;
        qd_ctx* c = ctxs[i];
        Guard g(c);
        if (!g.ok) {
            rcs[i] = QD_ERR_CUDA;
            return;
        }
        if ((rcs[i] = pipes_begin(c))) return;
        if ((rcs[i] = frames_long_pipeline(c, slot, enc, strict, frames, dna, n, cut[i], cut[i + 1], out, frame_at))) return;
        rcs[i] = pipes_end(c, &keys[i]);
    });
    unsigned long long key = NO_ERROR_KEY;
    for (int i = 0; i < n_ctx; i++) {
        if (rcs[i]) return rcs[i];
        key = std::min(key, keys[i]);
    }
    if (key != NO_ERROR_KEY) return report_single(enc, strict, dna, key, err);
    if (n_frames) *n_frames = nf;
    return QD_OK;
}

extern "C" int qd_revcomp_multi(qd_ctx* const* ctxs, int n_ctx, int enc, int strict, const uint8_t* dna, size_t n, uint8_t* out, qd_err* err) {
    clear_err(err);
    if (!valid_ctxs(ctxs, n_ctx) || (n && (!dna || !out))) return QD_ERR_INVALID_ARGUMENT;
    if (n == 0) return QD_OK;
    const int which = enc == QD_ENC_ASCII ? (strict ? RC_PAIR_ASCII_STRICT : RC_PAIR_ASCII) : (strict ? RC_PAIR_MASK_STRICT : RC_PAIR_MASK);
    const std::vector<uint64_t> cut = chromosome_cuts(n, n_ctx);
    std::vector<int> rcs(n_ctx, QD_OK);
    std::vector<unsigned long long> keys(n_ctx, NO_ERROR_KEY);
    run_per_device(n_ctx, [&](int i) {
        if (cut[i] >= cut[i + 1]) return;
        qd_ctx* c = ctxs[i];
        Guard g(c);
        if (!g.ok) {
            rcs[i] = QD_ERR_CUDA;
            return;
        }
        if ((rcs[i] = pipes_begin(c))) return;
        if ((rcs[i] = revcomp_long_pipeline(c, which, dna, n, cut[i], cut[i + 1], out))) return;
        rcs[i] = pipes_end(c, &keys[i]);
    });
    unsigned long long key = NO_ERROR_KEY;
    for (int i = 0; i < n_ctx; i++) {
        if (rcs[i]) return rcs[i];
        key = std::min(key, keys[i]);
    }
    if (key != NO_ERROR_KEY) return report_single(enc, strict, dna, key, err);
    return QD_OK;
}

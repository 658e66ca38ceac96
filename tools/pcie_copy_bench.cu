// Bare host<->device copy ceiling of a box, shaped like the host batch pipeline of qd_translate_frames_batch*
// (per GPU: `inflight` streams, each H2D of one input chunk then D2H of one output chunk = out_ratio x the input),
// so the e2e numbers of bench.py can be read against what the platform itself delivers at 1/2/4/8 GPUs.
//
//   nvcc -O2 -o tools/pcie_copy_bench tools/pcie_copy_bench.cu
//   tools/pcie_copy_bench --gpus 8 --mode threads|procs [--mb-in 2048] [--chunk-mb 64] [--inflight 3]
//                         [--out-ratio 2.016] [--flags default|portable|wc] [--dir both|h2d|d2h] [--reps 3]
// One JSON line per run: aggregate and per-GPU GB/s for each direction.
#include <cuda_runtime.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

static double now_s() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                            \
            exit(2);                                                                            \
        }                                                                                       \
    } while (0)

struct Shared {  // lives in MAP_SHARED memory so forked workers can use it
    std::atomic<int> arrived[8];
    double t_begin[64], t_end[64];
};

struct Opts {
    int gpus = 1, inflight = 3, reps = 3;
    size_t in_bytes = size_t(2048) << 20, chunk = size_t(64) << 20;
    double out_ratio = 2.016;
    std::string mode = "threads", flags = "default", dir = "both";
};

static void barrier(Shared* sh, int which, int n) {
    sh->arrived[which].fetch_add(1);
    while (sh->arrived[which].load() < n) sched_yield();
}

static void worker(const Opts& o, int g, Shared* sh) {
    CK(cudaSetDevice(g));
    const bool h2d = o.dir != "d2h", d2h = o.dir != "h2d";
    const size_t out_chunk = (size_t)(o.chunk * o.out_ratio + 15) & ~(size_t)15;
    const size_t n_chunks = (o.in_bytes + o.chunk - 1) / o.chunk;
    unsigned fl = cudaHostAllocDefault;
    if (o.flags == "portable") fl = cudaHostAllocPortable;
    uint8_t *h_in = nullptr, *h_out = nullptr;
    CK(cudaHostAlloc(&h_in, n_chunks * o.chunk, o.flags == "wc" ? cudaHostAllocWriteCombined : fl));  // WC only helps the H2D source
    CK(cudaHostAlloc(&h_out, n_chunks * out_chunk, fl));
    memset(h_in, 65, n_chunks * o.chunk);
    memset(h_out, 0, n_chunks * out_chunk);
    std::vector<cudaStream_t> st(o.inflight);
    std::vector<uint8_t*> d_in(o.inflight), d_out(o.inflight);
    for (int i = 0; i < o.inflight; i++) {
        CK(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking));
        CK(cudaMalloc(&d_in[i], o.chunk));
        CK(cudaMalloc(&d_out[i], out_chunk));
        CK(cudaMemsetAsync(d_out[i], 66, out_chunk, st[i]));
    }
    CK(cudaDeviceSynchronize());
    for (int rep = 0; rep < o.reps + 1; rep++) {  // rep 0 = warm-up
        barrier(sh, 2 * rep, o.gpus);
        const double t0 = now_s();
        for (size_t k = 0; k < n_chunks; k++) {
            const int i = (int)(k % o.inflight);
            if (h2d) CK(cudaMemcpyAsync(d_in[i], h_in + k * o.chunk, o.chunk, cudaMemcpyHostToDevice, st[i]));
            if (d2h) CK(cudaMemcpyAsync(h_out + k * out_chunk, d_out[i], out_chunk, cudaMemcpyDeviceToHost, st[i]));
        }
        for (int i = 0; i < o.inflight; i++) CK(cudaStreamSynchronize(st[i]));
        const double t1 = now_s();
        barrier(sh, 2 * rep + 1, o.gpus);
        if (rep > 0 && (rep == 1 || t1 - t0 < sh->t_end[g] - sh->t_begin[g])) {  // best rep of this GPU
            sh->t_begin[g] = t0;
            sh->t_end[g] = t1;
        }
    }
    cudaFreeHost(h_in);
    cudaFreeHost(h_out);
}

int main(int argc, char** argv) {
    Opts o;
    for (int i = 1; i + 1 < argc; i += 2) {
        std::string k = argv[i], v = argv[i + 1];
        if (k == "--gpus") o.gpus = atoi(v.c_str());
        else if (k == "--mode") o.mode = v;
        else if (k == "--mb-in") o.in_bytes = (size_t)atol(v.c_str()) << 20;
        else if (k == "--chunk-mb") o.chunk = (size_t)atol(v.c_str()) << 20;
        else if (k == "--inflight") o.inflight = atoi(v.c_str());
        else if (k == "--out-ratio") o.out_ratio = atof(v.c_str());
        else if (k == "--flags") o.flags = v;
        else if (k == "--dir") o.dir = v;
        else if (k == "--reps") o.reps = atoi(v.c_str());
    }
    if (o.gpus < 1 || o.gpus > 8 || o.reps < 1 || o.reps > 3) return 1;
    Shared* sh = (Shared*)mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    new (sh) Shared();
    for (auto& a : sh->arrived) a.store(0);
    if (o.mode == "procs") {
        std::vector<pid_t> kids;
        for (int g = 0; g < o.gpus; g++) {
            pid_t p = fork();
            if (p == 0) {
                worker(o, g, sh);
                _exit(0);
            }
            kids.push_back(p);
        }
        for (pid_t p : kids) {
            int stt = 0;
            waitpid(p, &stt, 0);
            if (!WIFEXITED(stt) || WEXITSTATUS(stt)) return 3;
        }
    } else {
        std::vector<std::thread> th;
        for (int g = 0; g < o.gpus; g++) th.emplace_back(worker, std::cref(o), g, sh);
        for (auto& t : th) t.join();
    }
    const bool h2d = o.dir != "d2h", d2h = o.dir != "h2d";
    const size_t out_chunk = (size_t)(o.chunk * o.out_ratio + 15) & ~(size_t)15;
    const size_t n_chunks = (o.in_bytes + o.chunk - 1) / o.chunk;
    const double in_b = h2d ? (double)n_chunks * o.chunk : 0, out_b = d2h ? (double)n_chunks * out_chunk : 0;
    double slowest = 0;
    printf("{\"tool\": \"pcie_copy_bench\", \"gpus\": %d, \"mode\": \"%s\", \"flags\": \"%s\", \"dir\": \"%s\", \"chunk_mb\": %zu, \"inflight\": %d, "
           "\"h2d_bytes_per_gpu\": %.0f, \"d2h_bytes_per_gpu\": %.0f, \"per_gpu_GBps\": [",
           o.gpus, o.mode.c_str(), o.flags.c_str(), o.dir.c_str(), o.chunk >> 20, o.inflight, in_b, out_b);
    for (int g = 0; g < o.gpus; g++) {
        const double dt = sh->t_end[g] - sh->t_begin[g];
        slowest = dt > slowest ? dt : slowest;
        printf("%s%.2f", g ? ", " : "", (in_b + out_b) / dt / 1e9);
    }
    printf("], \"seconds_slowest_gpu\": %.4f, \"aggregate_h2d_GBps\": %.2f, \"aggregate_d2h_GBps\": %.2f, \"aggregate_GBps\": %.2f}\n", slowest,
           o.gpus * in_b / slowest / 1e9, o.gpus * out_b / slowest / 1e9, o.gpus * (in_b + out_b) / slowest / 1e9);
    return 0;
}

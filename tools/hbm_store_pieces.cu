// HBM microbenchmark 3: the six-frame kernel's traffic shape (1 byte read : 2 bytes written, 1024-thread persistent CTAs,
// input tiles of 48 KB landed by TMA) with the OUTPUT leaving three ways, to decide how translate_frames_kernel<6> should store:
//   mode 0  per-thread st.global.v4 (what the kernel does today)
//   mode 1  per-WARP staging in shared memory + cp.async.bulk shared->global in pieces of PIECE bytes, issued by lane 0
//           (no CTA barrier on the output side; a warp only waits for its own previous pieces to have been read)
//   mode 2  per-CTA staging + ONE bulk store of the whole 96 KB tile (single output buffer: the wait for its read-out is exposed)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/hbm_store_pieces tools/hbm_store_pieces.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma_g2s(void* dst, const void* src, uint32_t n, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(n), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void tma_s2g(void* dst, const void* src, uint32_t n) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(n) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void stg16(void* p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

constexpr int THREADS = 1024, IN_TILE = THREADS * 48, OUT_TILE = THREADS * 96;

template <int MODE, int PIECE, int SPIN>
__global__ void __launch_bounds__(THREADS, 1) mix_kernel(const uint8_t* in, uint8_t* out, size_t n_tiles) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* in_s = smem;                 // 2 x IN_TILE
    uint8_t* out_s = smem + 2 * IN_TILE;  // OUT_TILE (modes 1, 2)
    __shared__ uint64_t full[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int s = 0; s < 2; s++) {
            size_t t = blockIdx.x + (size_t)s * gridDim.x;
            if (t < n_tiles) { mbar_expect_tx(&full[s], IN_TILE); tma_g2s(in_s + s * IN_TILE, in + t * IN_TILE, IN_TILE, &full[s]); }
        }
    }
    __syncthreads();
    uint32_t it = 0;
    for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int s = it & 1;
        mbar_wait(&full[s], (it >> 1) & 1);
        const uint4* src = reinterpret_cast<const uint4*>(in_s + s * IN_TILE + tid * 48);
        uint4 a = src[0], b = src[1], c = src[2];
        __syncthreads();
        if (tid == 0 && t + 2 * (size_t)gridDim.x < n_tiles) {
            mbar_expect_tx(&full[s], IN_TILE);
            tma_g2s(in_s + s * IN_TILE, in + (t + 2 * (size_t)gridDim.x) * IN_TILE, IN_TILE, &full[s]);
        }
        // stand-in for the codon arithmetic: SPIN dependent integer steps per output vector
        uint4 o[6] = {a, b, c, a, b, c};
#pragma unroll
        for (int k = 0; k < 6; k++)
#pragma unroll 1
            for (int j = 0; j < SPIN; j++) o[k].x = o[k].x * 1664525u + o[k].y;
        uint8_t* gout = out + t * (size_t)OUT_TILE;
        if (MODE == 0) {
            // six 16-byte stores per thread, consecutive threads -> consecutive granules of six separate streams
#pragma unroll
            for (int k = 0; k < 6; k++) stg16(gout + (size_t)k * (THREADS * 16) + tid * 16, o[k]);
        } else if (MODE == 1) {
            // warp-private 3 KB region: [k][lane] granules
            uint8_t* mine = out_s + warp * 3072;
            if (lane == 0) bulk_wait_read0();  // my previous pieces have left the stage
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 6; k++) *reinterpret_cast<uint4*>(mine + k * 512 + lane * 16) = o[k];
            fence_async();
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < 3072 / PIECE; q++) {
                    // pieces of at most 512 bytes go where mode 0 puts them (six streams, 512 contiguous bytes per warp and stream:
                    // the shape of the real kernel); larger pieces are contiguous per warp (what a friendlier layout would allow)
                    const int off = q * PIECE;
                    uint8_t* dst = PIECE <= 512 ? gout + (size_t)(off / 512) * (THREADS * 16) + warp * 512 + off % 512 : gout + warp * 3072 + off;
                    tma_s2g(dst, mine + off, PIECE);
                }
                bulk_commit();
            }
        } else {
            if (tid == 0) bulk_wait_read0();
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 6; k++) *reinterpret_cast<uint4*>(out_s + k * (THREADS * 16) + tid * 16) = o[k];
            fence_async();
            __syncthreads();
            if (tid == 0) {
                tma_s2g(gout, out_s, OUT_TILE);
                bulk_commit();
            }
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int MODE, int PIECE, int SPIN>
void run(const uint8_t* in, uint8_t* out, size_t n_tiles, int grid, const char* label) {
    const size_t smem = 2 * IN_TILE + (MODE ? OUT_TILE : 0);
    cudaFuncSetAttribute(mix_kernel<MODE, PIECE, SPIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 2; i++) mix_kernel<MODE, PIECE, SPIN><<<grid, THREADS, smem>>>(in, out, n_tiles);
    cudaEventRecord(e0);
    for (int i = 0; i < 5; i++) mix_kernel<MODE, PIECE, SPIN><<<grid, THREADS, smem>>>(in, out, n_tiles);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 5;
    cudaError_t e = cudaGetLastError();
    printf("%-34s piece %6d spin %3d  %7.3f ms  %7.1f GB/s  %s\n", label, PIECE, SPIN, ms, n_tiles * (double)(IN_TILE + OUT_TILE) / ms / 1e6,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    const size_t n_tiles = 200000;  // 9.8 GB in, 19.7 GB out
    uint8_t *in, *out;
    cudaMalloc(&in, n_tiles * IN_TILE);
    cudaMalloc(&out, n_tiles * (size_t)OUT_TILE);
    cudaMemset(in, 1, n_tiles * IN_TILE);
    cudaMemset(out, 0, n_tiles * (size_t)OUT_TILE);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    run<0, 16, 0>(in, out, n_tiles, sms, "st.global per thread");
    run<0, 16, 8>(in, out, n_tiles, sms, "st.global per thread");
    run<0, 16, 24>(in, out, n_tiles, sms, "st.global per thread");
    run<1, 512, 0>(in, out, n_tiles, sms, "warp stage + bulk pieces");
    run<1, 512, 8>(in, out, n_tiles, sms, "warp stage + bulk pieces");
    run<1, 512, 24>(in, out, n_tiles, sms, "warp stage + bulk pieces");
    run<1, 1024, 8>(in, out, n_tiles, sms, "warp stage + bulk pieces");
    run<1, 1536, 8>(in, out, n_tiles, sms, "warp stage + bulk pieces");
    run<1, 3072, 0>(in, out, n_tiles, sms, "warp stage + bulk pieces");
    run<1, 3072, 8>(in, out, n_tiles, sms, "warp stage + bulk pieces");
    run<1, 3072, 24>(in, out, n_tiles, sms, "warp stage + bulk pieces");
    run<1, 256, 8>(in, out, n_tiles, sms, "warp stage + bulk pieces");
    run<2, 98304, 0>(in, out, n_tiles, sms, "CTA stage + one bulk store");
    run<2, 98304, 8>(in, out, n_tiles, sms, "CTA stage + one bulk store");
    run<2, 98304, 24>(in, out, n_tiles, sms, "CTA stage + one bulk store");
    return 0;
}

// HBM microbenchmark 2: the same read:write mixes as hbm_mix.cu, but moved by TMA bulk copies
// (global -> shared -> global) from persistent CTAs, to see what the async-proxy path reaches.
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
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// one thread per CTA drives everything: chunk c of the input lands in smem, then is stored W times
template <int W, int STAGES>
__global__ void __launch_bounds__(32) tma_mix_kernel(const uint8_t* in, uint8_t* out, size_t n_chunks, uint32_t chunk) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full[STAGES];
    if (threadIdx.x != 0) return;
    for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const size_t total = n_chunks * (size_t)chunk;
    for (int s = 0; s < STAGES; s++) {
        size_t c = blockIdx.x + (size_t)s * gridDim.x;
        if (c < n_chunks) { mbar_expect_tx(&full[s], chunk); tma_g2s(smem + (size_t)s * chunk, in + c * chunk, chunk, &full[s]); }
    }
    uint32_t it = 0;
    for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x, ++it) {
        const int s = it % STAGES;
        mbar_wait(&full[s], (it / STAGES) & 1);
        for (int w = 0; w < W; w++) tma_s2g(out + (size_t)w * total + c * chunk, smem + (size_t)s * chunk, chunk);
        bulk_commit();
        bulk_wait_read<0>();  // smem source may be overwritten
        size_t nc = c + (size_t)STAGES * gridDim.x;
        if (nc < n_chunks) { mbar_expect_tx(&full[s], chunk); tma_g2s(smem + (size_t)s * chunk, in + nc * chunk, chunk, &full[s]); }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int W, int STAGES>
void run(const uint8_t* in, uint8_t* out, size_t bytes, uint32_t chunk, int grid) {
    const size_t n_chunks = bytes / chunk;
    const size_t smem = (size_t)STAGES * chunk;
    cudaFuncSetAttribute(tma_mix_kernel<W, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 2; i++) tma_mix_kernel<W, STAGES><<<grid, 32, smem>>>(in, out, n_chunks, chunk);
    cudaEventRecord(e0);
    for (int i = 0; i < 5; i++) tma_mix_kernel<W, STAGES><<<grid, 32, smem>>>(in, out, n_chunks, chunk);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 5;
    cudaError_t e = cudaGetLastError();
    printf("tma 1:%d  chunk %6u  stages %d  grid %4d  %7.3f ms  %7.1f GB/s  %s\n", W, chunk, STAGES, grid, ms, bytes * (1.0 + W) / ms / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    const size_t bytes = (size_t)4 << 30;
    uint8_t *in, *out;
    cudaMalloc(&in, bytes);
    cudaMalloc(&out, bytes * 2);
    cudaMemset(in, 1, bytes);
    cudaMemset(out, 0, bytes * 2);
    for (int grid : {148, 296, 592}) {
        run<1, 4>(in, out, bytes, 16384, grid);
        run<2, 4>(in, out, bytes, 16384, grid);
        run<1, 4>(in, out, bytes, 32768, grid);
        run<2, 4>(in, out, bytes, 32768, grid);
        run<2, 2>(in, out, bytes, 49152, grid);
        run<2, 8>(in, out, bytes, 8192, grid);
    }
    return 0;
}

// HBM microbenchmark: what does a streaming kernel reach for a given read:write mix on this GPU?
// (calibrates the roofline of the 1:2 six-frame kernel and the 1:1 reverse-complement kernel
// against the driver's copy peak in MEASURED_PEAKS.json).  build: nvcc -arch=sm_100a -O3 -o hbm_mix tools/hbm_mix.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

template <int R, int W>
__global__ void __launch_bounds__(256) mix_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n_units) {
    // unit i: reads R uint4 (in[i*R ..]) and writes W uint4 (out[i*W ..]); grid-stride over units
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_units; i += (size_t)gridDim.x * blockDim.x) {
        uint4 acc = make_uint4(1, 2, 3, 4);
#pragma unroll
        for (int r = 0; r < R; r++) {
            uint4 v;
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(in + r * n_units + i));
            acc.x ^= v.x; acc.y += v.y; acc.z ^= v.z; acc.w += v.w;
        }
#pragma unroll
        for (int w = 0; w < W; w++) {
            uint4 v = make_uint4(acc.x + w, acc.y, acc.z, acc.w);
            asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(out + w * n_units + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        }
    }
}

template <int R, int W>
void run(const char* name, uint4* in, uint4* out, size_t n_units, int grid) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) mix_kernel<R, W><<<grid, 256>>>(in, out, n_units);
    cudaEventRecord(e0);
    const int reps = 10;
    for (int i = 0; i < reps; i++) mix_kernel<R, W><<<grid, 256>>>(in, out, n_units);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    const double bytes = (double)n_units * 16 * (R + W);
    printf("%-12s grid %6d  R:W %d:%d  %8.3f ms  %8.1f GB/s\n", name, grid, R, W, ms, bytes / ms / 1e6);
}

int main() {
    const size_t unit_bytes = 16, n_units = (size_t)1 << 28;  // 4 GiB per stream
    uint4 *in, *out;
    cudaMalloc(&in, n_units * unit_bytes * 2);
    cudaMalloc(&out, n_units * unit_bytes * 2);
    cudaMemset(in, 1, n_units * unit_bytes * 2);
    cudaMemset(out, 0, n_units * unit_bytes * 2);
    int sms = 148;
    for (int occ : {2, 4, 8, 16}) {
        const int grid = sms * occ;
        run<1, 1>("copy", in, out, n_units, grid);
        run<1, 2>("frames-mix", in, out, n_units, grid);
        run<0, 1>("write", in, out, n_units, grid);
        run<1, 0>("read", in, out, n_units, grid);
        run<2, 1>("2r1w", in, out, n_units / 2, grid);
    }
    // cudaMemcpy D2D and memset for reference
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaMemcpy(out, in, n_units * 16, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; i++) cudaMemcpyAsync(out, in, n_units * 16, cudaMemcpyDeviceToDevice);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("cudaMemcpy D2D %8.1f GB/s (read+write)\n", 2.0 * n_units * 16 * 10 / ms / 1e6);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; i++) cudaMemsetAsync(out, 0, n_units * 16);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("cudaMemset     %8.1f GB/s\n", 1.0 * n_units * 16 * 10 / ms / 1e6);
    return 0;
}

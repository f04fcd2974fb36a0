// Micro-benchmark (development aid, not part of the library): issue rate of the legacy warp-level tensor path
// (mma.sync.m16n8k16 bf16 -> HMMA.16816.F32.BF16) on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_bench hmma_bench.cu && ./hmma_bench
// One CTA per SM, W warps per CTA (W/4 per SM sub-partition); every warp runs ITER iterations of CH independent
// accumulator chains.  Prints cycles per HMMA per sub-partition and the implied dense TFLOP/s of the whole GPU.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITER = 4096;

template <int CH>
__global__ void k(float* out, long long* cyc, uint32_t seed) {
    float acc[CH][4];
    for (int c = 0; c < CH; ++c)
        for (int i = 0; i < 4; ++i) acc[c][i] = 0.f;
    uint32_t a0 = 0x3c003c00u ^ seed, a1 = a0 + threadIdx.x, a2 = a0 ^ 0x00010001u, a3 = a1, b0 = 0x3f803f80u, b1 = b0 ^ seed;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int c = 0; c < CH; ++c)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(acc[c][0]), "+f"(acc[c][1]), "+f"(acc[c][2]), "+f"(acc[c][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    const long long t1 = clock64();
    float s = 0.f;
    for (int c = 0; c < CH; ++c)
        for (int i = 0; i < 4; ++i) s += acc[c][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int CH>
static void run(int warps, int n_sm, float* out, long long* cyc, double ghz) {
    k<CH><<<n_sm, warps * 32>>>(out, cyc, 0u);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<CH><<<n_sm, warps * 32>>>(out, cyc, 0u);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    long long c0 = 0;
    cudaMemcpy(&c0, cyc, sizeof c0, cudaMemcpyDeviceToHost);
    const double per_smsp = static_cast<double>(ITER) * CH * warps / 4.0;      // HMMAs per sub-partition
    const double flops = 2.0 * 16 * 8 * 16 * ITER * CH * warps * n_sm;
    printf("warps/SMSP %d  chains %d : %.2f cycles per HMMA per SMSP (chain latency %.1f cycles), %.0f TFLOP/s by events\n",
           warps / 4, CH, c0 / per_smsp, static_cast<double>(c0) / ITER, flops / (ms * 1e-3) / 1e12);
    (void)ghz;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * p.multiProcessorCount * 1024);
    cudaMalloc(&cyc, sizeof(long long) * p.multiProcessorCount);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    for (int w : {4, 8, 16, 32}) {
        run<1>(w, p.multiProcessorCount, out, cyc, 0);
        run<4>(w, p.multiProcessorCount, out, cyc, 0);
        run<8>(w, p.multiProcessorCount, out, cyc, 0);
    }
    return 0;
}

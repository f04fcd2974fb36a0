// Micro-benchmark (development aid, not part of the library): MUFU.EX2 issue rates on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu && ./mufu_bench
// One CTA per SM, W warps per CTA (W/4 per SM sub-partition); every thread runs ITER iterations of 8
// independent chains.  Prints cycles per warp-instruction per sub-partition for
//   f32    ex2.approx.ftz.f32            (1 element / lane / instruction)
//   f16x2  ex2.approx.f16x2              (2 elements)
//   bf16x2 ex2.approx.ftz.bf16x2         (2 elements)
//   soft32 the current softmax inner loop: FADD2, 2 x ex2.f32, cvt.bf16x2, 2 x unpack, FADD2
//   soft16 the candidate: FADD2, cvt.f16x2, ex2.f16x2, 2 x mixed add (f32 += f16)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

constexpr int ITER = 2048;

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2b2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }

template <int MODE>
__global__ void k(float* out, long long* cyc, float seed) {
    float a[8];
    uint32_t u[8];
    for (int i = 0; i < 8; ++i) { a[i] = seed * (threadIdx.x + i) * 1e-3f - 1.0f; u[i] = 0xb800b800u + i; }   // ~ -0.5 halves
    float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
    float s0 = 0.f, s1 = 0.f;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = ex2f(a[i]) - 1.5f;           // FADD keeps the argument in range
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) u[i] = ex2h2(u[i]) ^ 0x80008000u;   // negate both halves (LOP3)
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) u[i] = ex2b2(u[i]) ^ 0x80008000u;
        } else if (MODE == 3) {                                               // current softmax loop
            const float2 nm = make_float2(-seed, -seed);
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                const float2 x = __fadd2_rn(make_float2(a[i], a[i + 1]), nm);
                uint32_t pk;
                const float e0 = ex2f(x.x), e1 = ex2f(x.y);
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(e1), "f"(e0));
                const float2 r = make_float2(__uint_as_float(pk << 16), __uint_as_float(pk & 0xffff0000u));
                if (i & 2) acc0 = __fadd2_rn(acc0, r); else acc1 = __fadd2_rn(acc1, r);
                u[i >> 1] ^= pk;
                a[i] = r.x - 2.f; a[i + 1] = r.y - 2.f;
            }
        } else if (MODE == 4) {                                               // candidate f16x2 loop
            const float2 nm = make_float2(-seed, -seed);
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                const float2 x = __fadd2_rn(make_float2(a[i], a[i + 1]), nm);
                uint32_t xh, pk;
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(xh) : "f"(x.y), "f"(x.x));
                pk = ex2h2(xh);
                __half2 ph = *reinterpret_cast<__half2*>(&pk);
                asm volatile("add.rn.f32.f16 %0, %1, %0;" : "+f"(s0) : "h"(__half_as_ushort(__low2half(ph))));
                asm volatile("add.rn.f32.f16 %0, %1, %0;" : "+f"(s1) : "h"(__half_as_ushort(__high2half(ph))));
                u[i >> 1] ^= pk;
                a[i] = a[i] * 0.5f - 0.25f; a[i + 1] = a[i + 1] * 0.5f - 0.25f;
            }
        }
    }
    if (MODE == 5 || MODE == 6) {      // 5: polynomial exp2 for every pair; 6: alternate MUFU / polynomial pairs
        const float magic = 12582912.f - seed;
        const float2 k2 = make_float2(magic, magic), nk2 = make_float2(-magic, -magic), nm = make_float2(-seed, -seed);
        const float smin = seed - 125.f;
#pragma unroll 1
        for (int it = 0; it < ITER; ++it) {
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                float2 sv = make_float2(a[i], a[i + 1]);
                uint32_t pk;
                if (MODE == 5 || (i & 2)) {
                    sv.x = fmaxf(sv.x, smin); sv.y = fmaxf(sv.y, smin);
                    const float2 tt = __fadd2_rn(sv, k2);
                    const float2 f = __ffma2_rn(__fadd2_rn(tt, nk2), make_float2(-1.f, -1.f), sv);
                    float2 pp = __ffma2_rn(make_float2(0.0551714078f, 0.0551714078f), f, make_float2(0.2426107526f, 0.2426107526f));
                    pp = __ffma2_rn(pp, f, make_float2(0.6932609677f, 0.6932609677f));
                    pp = __ffma2_rn(pp, f, make_float2(0.9999281168f, 0.9999281168f));
                    const float e0 = __int_as_float(__float_as_int(pp.x) + (__float_as_int(tt.x) << 23));
                    const float e1 = __int_as_float(__float_as_int(pp.y) + (__float_as_int(tt.y) << 23));
                    asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(e1), "f"(e0));
                } else {
                    const float2 x = __fadd2_rn(sv, nm);
                    const float e0 = ex2f(x.x), e1 = ex2f(x.y);
                    asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(e1), "f"(e0));
                }
                const float2 r = make_float2(__uint_as_float(pk << 16), __uint_as_float(pk & 0xffff0000u));
                if (i & 2) acc0 = __fadd2_rn(acc0, r); else acc1 = __fadd2_rn(acc1, r);
                u[i >> 1] ^= pk;
                a[i] = r.x - 2.f; a[i + 1] = r.y - 2.f;
            }
        }
    }
    if (MODE == 7 || MODE == 8) {      // FHADD.BF16 row sums (f32 += one bf16 half of the packed word): no unpack, no FADD2
        const float2 nm = make_float2(-seed, -seed);
        float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll 1
        for (int it = 0; it < ITER; ++it) {
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                float2 x;
                if (MODE == 7) x = __fadd2_rn(make_float2(a[i], a[i + 1]), nm);
                else { x.x = a[i] - seed; x.y = a[i + 1] - seed; }
                uint32_t pk;
                const float e0 = ex2f(x.x), e1 = ex2f(x.y);
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(e1), "f"(e0));
                uint16_t lo, hi;
                asm("mov.b32 {%0, %1}, %2;" : "=h"(lo), "=h"(hi) : "r"(pk));
                if (i & 2) {
                    asm("add.rn.f32.bf16 %0, %1, %0;" : "+f"(q0) : "h"(lo));
                    asm("add.rn.f32.bf16 %0, %1, %0;" : "+f"(q1) : "h"(hi));
                } else {
                    asm("add.rn.f32.bf16 %0, %1, %0;" : "+f"(q2) : "h"(lo));
                    asm("add.rn.f32.bf16 %0, %1, %0;" : "+f"(q3) : "h"(hi));
                }
                u[i >> 1] ^= pk;
                a[i] = e0 - 2.f; a[i + 1] = e1 - 2.f;
            }
        }
        s0 += q0 + q1 + q2 + q3;
    }
    const long long t1 = clock64();
    float r = acc0.x + acc0.y + acc1.x + acc1.y + s0 + s1;
    for (int i = 0; i < 8; ++i) r += a[i] + __uint_as_float(u[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_iter_instr) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    for (int warps : {4, 8, 16, 32}) {
        k<MODE><<<148, warps * 32>>>(out, cyc, 1.0f);
        k<MODE><<<148, warps * 32>>>(out, cyc, 1.0f);
        cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        const double instr_per_smsp = double(ITER) * per_iter_instr * (warps / 4);
        printf("%-7s warps/SMSP %d: %9.0f cycles, %.2f cycles per MUFU warp-instruction per SMSP\n", name, warps / 4, avg,
               avg / instr_per_smsp);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("%s: %s\n", name, cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("f32", 8);
    run<1>("f16x2", 16);
    run<2>("bf16x2", 16);
    run<3>("soft32", 8);     // 8 MUFU for 8 elements
    run<4>("soft16", 8);     // ex2.f16x2 = two MUFU.EX2.F16 in SASS
    run<7>("fhadd", 8);      // FADD2 sub, 2 EX2, cvt, 2 FHADD.BF16
    run<8>("fhadd1", 8);     // FADD sub x2, 2 EX2, cvt, 2 FHADD.BF16
    run<5>("poly", 8);       // cycles per ELEMENT (no MUFU)
    run<6>("mix50", 8);      // cycles per element, half MUFU half polynomial
    return 0;
}

// K1 -- instance-norm statistics (HBM-bound).
//
// Replaces the statistics half of nn.InstanceNorm2d(affine=False) in
// MHAdaSTr/network/adaDecoder.py:147-149 (applied at :173, :178, :198).  The normalisation itself
// is never materialised: it is folded into the projection weights (proj_*.cu) and into the
// attention epilogue (attn_*.cu).
//
// Layout: x is token-major [B, N, ld]; a warp reads 32 x 16 B = 512 contiguous bytes of one token
// row per load (fully coalesced, 128-bit), each lane owning VEC consecutive channels.  Sums are
// taken relative to a per-channel pivot (the first token) so that E[(x-p)^2] - E[x-p]^2 does not
// cancel when |mean| >> std.  Pass 1 writes per-split partial sums (deterministic, no atomics);
// pass 2 combines them in double and emits mean and rstd = 1/sqrt(var + 1e-5).
//
// Algorithmic bytes: B*N*C*sizeof(T) read once; output 2*B*C floats.
#include "common.h"
#include "ptx.cuh"

namespace mh {

constexpr int kStatsThreads = 256;
constexpr int kStatsWarps = kStatsThreads / 32;
constexpr float kInEps = 1e-5f;  // nn.InstanceNorm2d default

template <typename T>
struct VecIO;
template <>
struct VecIO<float> {
    static constexpr int VEC = 4;
    __device__ static void load(const float* p, float (&v)[4]) {
        float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
};
template <>
struct VecIO<__nv_bfloat16> {
    static constexpr int VEC = 8;
    __device__ static void load(const __nv_bfloat16* p, float (&v)[8]) {
        uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
        v[0] = bf16_lo(t.x); v[1] = bf16_hi(t.x); v[2] = bf16_lo(t.y); v[3] = bf16_hi(t.y);
        v[4] = bf16_lo(t.z); v[5] = bf16_hi(t.z); v[6] = bf16_lo(t.w); v[7] = bf16_hi(t.w);
    }
};

// grid (channel chunks, splits, B)
template <typename T>
__global__ void __launch_bounds__(kStatsThreads) stats_partial_kernel(const T* __restrict__ x, int N, int C, int ld,
                                                                      int tokens_per_split,
                                                                      float* __restrict__ partial) {
    constexpr int VEC = VecIO<T>::VEC;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = (blockIdx.x * 32 + lane) * VEC;
    const int split = blockIdx.y, splits = gridDim.y, b = blockIdx.z;
    const bool active = c0 < C;
    const T* xb = x + static_cast<size_t>(b) * N * ld + c0;

    float piv[VEC], s1[VEC], s2[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) s1[v] = s2[v] = 0.f, piv[v] = 0.f;
    if (active) VecIO<T>::load(xb, piv);

    const int t_begin = split * tokens_per_split;
    const int t_end = min(N, t_begin + tokens_per_split);
    if (active) {
        int t = t_begin + warp;
        // 4 tokens in flight per lane: 4 independent 16 B loads before any use
        for (; t + 3 * kStatsWarps < t_end; t += 4 * kStatsWarps) {
            float a[4][VEC];
#pragma unroll
            for (int u = 0; u < 4; ++u) VecIO<T>::load(xb + static_cast<size_t>(t + u * kStatsWarps) * ld, a[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    float d = a[u][v] - piv[v];
                    s1[v] += d;
                    s2[v] = fmaf(d, d, s2[v]);
                }
        }
        for (; t < t_end; t += kStatsWarps) {
            float a[VEC];
            VecIO<T>::load(xb + static_cast<size_t>(t) * ld, a);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                float d = a[v] - piv[v];
                s1[v] += d;
                s2[v] = fmaf(d, d, s2[v]);
            }
        }
    }
    // combine the 8 token lanes of this CTA
    __shared__ float red[kStatsWarps][32][2 * VEC + 1];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        red[warp][lane][v] = s1[v];
        red[warp][lane][VEC + v] = s2[v];
    }
    __syncthreads();
    if (warp == 0 && active) {
        float* dst = partial + ((static_cast<size_t>(b) * splits + split) * C + c0) * 2;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            float a = 0.f, q = 0.f;
#pragma unroll
            for (int w = 0; w < kStatsWarps; ++w) {
                a += red[w][lane][v];
                q += red[w][lane][VEC + v];
            }
            dst[2 * v] = a;
            dst[2 * v + 1] = q;
        }
    }
}

template <typename T>
__global__ void stats_final_kernel(const T* __restrict__ x, const float* __restrict__ partial, int B, int N, int C,
                                   int ld, int splits, float* __restrict__ mean, float* __restrict__ rstd) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * C) return;
    int b = i / C, c = i % C;
    double a = 0.0, q = 0.0;
    for (int s = 0; s < splits; ++s) {
        const float* p = partial + ((static_cast<size_t>(b) * splits + s) * C + c) * 2;
        a += p[0];
        q += p[1];
    }
    double piv;
    if constexpr (sizeof(T) == 2)
        piv = __bfloat162float(x[static_cast<size_t>(b) * N * ld + c]);
    else
        piv = x[static_cast<size_t>(b) * N * ld + c];
    double m = a / N;
    double var = q / N - m * m;
    if (var < 0.0) var = 0.0;
    mean[i] = static_cast<float>(piv + m);
    rstd[i] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(kInEps)));
}

// Tokens per split depend on N only (never on B), so the statistics of an image -- and with them the
// whole layer -- are bit-identical whatever batch the image is part of.  At most 128 splits.
static int stats_tokens_per_split(int N) {
    int t = (N + 127) / 128;
    return t < 64 ? 64 : t;
}
static int stats_splits(int /*B*/, int N) {
    int t = stats_tokens_per_split(N);
    return (N + t - 1) / t;
}

size_t stats_workspace(int B, int N, int C) {
    return static_cast<size_t>(B) * stats_splits(B, N) * C * 2 * sizeof(float);
}

int launch_stats(const void* x, int dtype, int B, int N, int C, int ld, float* mean, float* rstd, float* ws,
                 cudaStream_t s) {
    const int splits = stats_splits(B, N);
    const int tps = stats_tokens_per_split(N);
    if (dtype == MHADA_BF16) {
        dim3 grid((C + 32 * 8 - 1) / (32 * 8), splits, B);
        stats_partial_kernel<__nv_bfloat16><<<grid, kStatsThreads, 0, s>>>(static_cast<const __nv_bfloat16*>(x), N, C, ld, tps, ws);
        count_launch();
        stats_final_kernel<__nv_bfloat16><<<(B * C + 255) / 256, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), ws, B, N, C, ld, splits, mean, rstd);
        count_launch();
    } else {
        dim3 grid((C + 32 * 4 - 1) / (32 * 4), splits, B);
        stats_partial_kernel<float><<<grid, kStatsThreads, 0, s>>>(static_cast<const float*>(x), N, C, ld, tps, ws);
        count_launch();
        stats_final_kernel<float><<<(B * C + 255) / 256, 256, 0, s>>>(static_cast<const float*>(x), ws, B, N, C, ld, splits, mean, rstd);
        count_launch();
    }
    return check_cuda(cudaGetLastError(), "stats launch");
}

}  // namespace mh

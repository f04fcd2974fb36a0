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

// Up to three tensors (fc, fs, fcs of one layer) go through ONE launch of each pass.
#define MH_SEL3(arr, i) ((i) == 0 ? (arr)[0] : ((i) == 1 ? (arr)[1] : (arr)[2]))   // no local copy of the params
struct StatsJob {
    const void* x[3];
    float* mean[3];
    float* rstd[3];
    int N[3];
    int tps[3];        // tokens per split
    int splits[3];
    int n;             // tensors in this job
    int max_splits;
};

// grid (channel chunks, max_splits, B * n_tensors)
template <typename T>
__global__ void __launch_bounds__(kStatsThreads) stats_partial_kernel(const StatsJob job, int B, int C, int ld,
                                                                      float* __restrict__ partial) {
    constexpr int VEC = VecIO<T>::VEC;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = (blockIdx.x * 32 + lane) * VEC;
    const int ti = blockIdx.z / B, b = blockIdx.z % B;
    const int split = blockIdx.y;
    if (split >= MH_SEL3(job.splits, ti)) return;
    const int N = MH_SEL3(job.N, ti), tokens_per_split = MH_SEL3(job.tps, ti);
    const bool active = c0 < C;
    const T* xb = static_cast<const T*>(MH_SEL3(job.x, ti)) + static_cast<size_t>(b) * N * ld + c0;

    float piv[VEC], s1[VEC], s2[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) s1[v] = s2[v] = 0.f, piv[v] = 0.f;
    if (active) VecIO<T>::load(xb, piv);

    const int t_begin = split * tokens_per_split;
    const int t_end = min(N, t_begin + tokens_per_split);
    if (active) {
        int t = t_begin + warp;
        // 8 tokens in flight per lane: 8 independent 16 B loads before any use
        for (; t + 7 * kStatsWarps < t_end; t += 8 * kStatsWarps) {
            float a[8][VEC];
#pragma unroll
            for (int u = 0; u < 8; ++u) VecIO<T>::load(xb + static_cast<size_t>(t + u * kStatsWarps) * ld, a[u]);
#pragma unroll
            for (int u = 0; u < 8; ++u)
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
        float* dst = partial + (((static_cast<size_t>(ti) * B + b) * job.max_splits + split) * C + c0) * 2;
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
__global__ void stats_final_kernel(const StatsJob job, const float* __restrict__ partial, int B, int C, int ld) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= job.n * B * C) return;
    const int ti = i / (B * C), r = i % (B * C), b = r / C, c = r % C;
    const int N = MH_SEL3(job.N, ti);
    const int nsplit = MH_SEL3(job.splits, ti);
    double a = 0.0, q = 0.0;
    for (int s = 0; s < nsplit; ++s) {
        const float* p = partial + (((static_cast<size_t>(ti) * B + b) * job.max_splits + s) * C + c) * 2;
        a += p[0];
        q += p[1];
    }
    const T* x = static_cast<const T*>(MH_SEL3(job.x, ti));
    double piv;
    if constexpr (sizeof(T) == 2)
        piv = __bfloat162float(x[static_cast<size_t>(b) * N * ld + c]);
    else
        piv = x[static_cast<size_t>(b) * N * ld + c];
    double m = a / N;
    double var = q / N - m * m;
    if (var < 0.0) var = 0.0;
    MH_SEL3(job.mean, ti)[r] = static_cast<float>(piv + m);
    MH_SEL3(job.rstd, ti)[r] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(kInEps)));
}

// Tokens per split depend on N only (never on B), so the statistics of an image -- and with them the
// whole layer -- are bit-identical whatever batch the image is part of.  At most 128 splits: each CTA
// streams >= 64 tokens with eight 16-byte loads in flight per lane (with 32 splits a single-tensor launch was
// 512 CTAs on cfg2 and 64 on cfg3: 2.9 TB/s and 1.5 TB/s).
constexpr int kStatsMaxSplits = 128;
static int stats_tokens_per_split(int N) {
    int t = (N + kStatsMaxSplits - 1) / kStatsMaxSplits;
    return t < 64 ? 64 : t;
}
static int stats_splits(int N) {
    int t = stats_tokens_per_split(N);
    return (N + t - 1) / t;
}

size_t stats_workspace(int B, int N, int C) {
    // room for three tensors of up to N tokens (one launch serves fc, fs and fcs of a layer)
    return static_cast<size_t>(3) * B * kStatsMaxSplits * C * 2 * sizeof(float);
}

static StatsJob make_job(int n, const void* const* x, const int* N, float* const* mean, float* const* rstd) {
    StatsJob job{};
    job.n = n;
    job.max_splits = 0;
    for (int i = 0; i < n; ++i) {
        job.x[i] = x[i]; job.N[i] = N[i];
        job.mean[i] = mean ? mean[i] : nullptr;
        job.rstd[i] = rstd ? rstd[i] : nullptr;
        job.tps[i] = stats_tokens_per_split(N[i]);
        job.splits[i] = stats_splits(N[i]);
        if (job.splits[i] > job.max_splits) job.max_splits = job.splits[i];
    }
    return job;
}

static void launch_partial(const StatsJob& job, int dtype, int B, int C, int ld, float* ws, cudaStream_t s) {
    const int vec = dtype == MHADA_BF16 ? 8 : 4;
    dim3 grid((C + 32 * vec - 1) / (32 * vec), job.max_splits, B * job.n);
    if (dtype == MHADA_BF16)
        stats_partial_kernel<__nv_bfloat16><<<grid, kStatsThreads, 0, s>>>(job, B, C, ld, ws);
    else
        stats_partial_kernel<float><<<grid, kStatsThreads, 0, s>>>(job, B, C, ld, ws);
    count_launch();
}

int launch_stats_multi(int n, const void* const* x, const int* N, float* const* mean, float* const* rstd, int dtype,
                       int B, int C, int ld, float* ws, cudaStream_t s) {
    const StatsJob job = make_job(n, x, N, mean, rstd);
    launch_partial(job, dtype, B, C, ld, ws, s);
    const int fin_blocks = (n * B * C + 255) / 256;
    if (dtype == MHADA_BF16)
        stats_final_kernel<__nv_bfloat16><<<fin_blocks, 256, 0, s>>>(job, ws, B, C, ld);
    else
        stats_final_kernel<float><<<fin_blocks, 256, 0, s>>>(job, ws, B, C, ld);
    count_launch();
    return check_cuda(cudaGetLastError(), "stats launch");
}

// First pass only: the per-split partial sums [n][B][max_splits][C][2] stay in `ws`; the caller's next kernel
// (fold_stats_kernel, proj_tc.cu) finishes the statistics of the channels it needs.
int launch_stats_partial(int n, const void* const* x, const int* N, int dtype, int B, int C, int ld, float* ws,
                         StatsPartialInfo* info, cudaStream_t s) {
    const StatsJob job = make_job(n, x, N, nullptr, nullptr);
    launch_partial(job, dtype, B, C, ld, ws, s);
    info->max_splits = job.max_splits;
    for (int i = 0; i < 3; ++i) info->splits[i] = i < n ? job.splits[i] : 0;
    return check_cuda(cudaGetLastError(), "stats partial launch");
}

int launch_stats(const void* x, int dtype, int B, int N, int C, int ld, float* mean, float* rstd, float* ws,
                 cudaStream_t s) {
    const void* xs[1] = {x};
    float* ms[1] = {mean};
    float* rs[1] = {rstd};
    int ns[1] = {N};
    return launch_stats_multi(1, xs, ns, ms, rs, dtype, B, C, ld, ws, s);
}

}  // namespace mh

// Backward of one MHAda layer around the attention stage: the HBM-bound helper kernels.  Every
// contraction of the backward other than the attention itself runs on the tcgen05 token GEMM (gemm_tc.cu):
//   d(cat)      = d(out) . Wo                        A = d(out) [M, C],          W = Wo^T
//   dWo         = d(out)^T . cat                     A = d(out)^T [C, M],        W = cat^T [C, M]     (K = tokens)
//   d(IN(fc))   = dQ . blockdiag(Wf)                 A = dQ [M, C],              W = blockdiag(Wf_h)^T
//   dWf (all heads at once) = dQ^T . IN(fc)          A = dQ^T [C, M],            W = IN(fc)^T [C, M]; diagonal blocks kept
// (same for g / K / IN(fs) and h / V / fs).  The 1x1 convolutions are block diagonal over the heads
// (adaDecoder.py:173-183); running them as dense C x C GEMMs costs 8x the FLOPs of the grouped form and is still
// HBM bound (2 M C^2 = 4.3 GFLOP per GEMM at the training size).  This file provides what those GEMMs need around them:
// transposes (with the instance norm applied on the fly), the block-diagonal weight packing, the per-(image, channel)
// token sums of the instance-norm backward, bias gradients, and the instance-norm backward itself
//   dx = rstd (g - mean_n(g) - x^ mean_n(g x^))      (adaDecoder.py:147-149 differentiated).
// All reductions are two-stage and deterministic (no atomics).
#include "common.h"
#include "ptx.cuh"

namespace mh {
namespace {

__device__ __forceinline__ float ldv(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ldv(const __nv_bfloat16* p) { return __bfloat162float(*p); }

__device__ __forceinline__ float2 ldv2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ float2 ldv2(const __nv_bfloat16* p) {
    const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(p));
    return make_float2(bf16_lo(w), bf16_hi(w));
}

// out[c][m] = bf16(norm(in[m][c])) for m < M, 0 for M <= m < Mpad;  norm = (x - mean[b][c]) * rstd[b][c] with b = m / N
// when mean != nullptr.  64 x 64 tiles through shared memory, two elements per access on both sides (C and Mpad even).
template <typename T>
__global__ void __launch_bounds__(256) transpose_norm_kernel(const T* __restrict__ in, int ld, int M, int Mpad, int C, int N,
                                                             const float* __restrict__ mean, const float* __restrict__ rstd,
                                                             __nv_bfloat16* __restrict__ out) {
    __shared__ float tile[64][65];
    const int m0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 pairs x 8
#pragma unroll
    for (int r = ty; r < 64; r += 8) {
        const int m = m0 + r, c = c0 + 2 * tx;
        float2 v = make_float2(0.f, 0.f);
        if (m < M && c < C) {
            v = ldv2(in + static_cast<size_t>(m) * ld + c);
            if (mean) {
                const int b = m / N;
                const float2 mu = __ldg(reinterpret_cast<const float2*>(mean + static_cast<size_t>(b) * C + c));
                const float2 rs = __ldg(reinterpret_cast<const float2*>(rstd + static_cast<size_t>(b) * C + c));
                v = make_float2((v.x - mu.x) * rs.x, (v.y - mu.y) * rs.y);
            }
        }
        tile[r][2 * tx] = v.x;
        tile[r][2 * tx + 1] = v.y;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 64; r += 8) {
        const int c = c0 + r, m = m0 + 2 * tx;
        if (c < C && m < Mpad)
            *reinterpret_cast<uint32_t*>(out + static_cast<size_t>(c) * Mpad + m) = pack_bf16x2(tile[2 * tx][r], tile[2 * tx + 1][r]);
    }
}

// out[r][ig][og] = w[r][h][o][i] when ig = h d + i and og = h d + o lie in the same head, else 0   (bf16, r = f, g, h)
__global__ void __launch_bounds__(256) blockdiag_t_kernel(const float* __restrict__ w, int H, int d, __nv_bfloat16* __restrict__ out) {
    const int C = H * d;
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= static_cast<size_t>(3) * C * C) return;
    const int og = static_cast<int>(i % C), ig = static_cast<int>((i / C) % C), r = static_cast<int>(i / (static_cast<size_t>(C) * C));
    float v = 0.f;
    if (og / d == ig / d) v = __ldg(w + ((static_cast<size_t>(r) * H + og / d) * d + og % d) * d + ig % d);
    out[i] = __float2bfloat16_rn(v);
}

// dw[h][o][i] = full[h d + o][h d + i]
__global__ void __launch_bounds__(256) extract_blockdiag_kernel(const float* __restrict__ full, int H, int d, float* __restrict__ dw) {
    const int C = H * d;
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= static_cast<size_t>(H) * d * d) return;
    const int ii = static_cast<int>(i % d), o = static_cast<int>((i / d) % d), h = static_cast<int>(i / (static_cast<size_t>(d) * d));
    dw[i] = __ldg(full + static_cast<size_t>(h * d + o) * C + h * d + ii);
}

constexpr int TS_ROWS = 32;      // token rows per CTA of the first stage

// partial[b][split][c] = (sum_n g, sum_n g x^) over the TS_ROWS rows of the split; thread = two channels, the row loop
// is unrolled so that eight independent loads are in flight per thread.
template <typename TG, typename TX>
__global__ void __launch_bounds__(1024) token_sums_kernel(const TG* __restrict__ g, const TX* __restrict__ x,
                                                          const float* __restrict__ mean, const float* __restrict__ rstd, int N, int C,
                                                          float2* __restrict__ partial) {
    const int c = threadIdx.x * 2;
    if (c >= C) return;
    const int split = blockIdx.x, b = blockIdx.y, splits = gridDim.x;
    const int n0 = split * TS_ROWS, n1 = min(N, n0 + TS_ROWS);
    float mu0 = 0.f, mu1 = 0.f, r0 = 0.f, r1 = 0.f;
    if (x) {
        mu0 = __ldg(mean + static_cast<size_t>(b) * C + c); mu1 = __ldg(mean + static_cast<size_t>(b) * C + c + 1);
        r0 = __ldg(rstd + static_cast<size_t>(b) * C + c); r1 = __ldg(rstd + static_cast<size_t>(b) * C + c + 1);
    }
    float s0 = 0.f, s1 = 0.f, t0 = 0.f, t1 = 0.f;
    const size_t base = static_cast<size_t>(b) * N * C + c;
#pragma unroll 8
    for (int n = n0; n < n1; ++n) {
        const size_t off = base + static_cast<size_t>(n) * C;
        const float2 gv = ldv2(g + off);
        s0 += gv.x; s1 += gv.y;
        if (x) {
            const float2 xv = ldv2(x + off);
            t0 = fmaf(gv.x, (xv.x - mu0) * r0, t0);
            t1 = fmaf(gv.y, (xv.y - mu1) * r1, t1);
        }
    }
    float2* o = partial + (static_cast<size_t>(b) * splits + split) * C + c;
    o[0] = make_float2(s0, t0);
    o[1] = make_float2(s1, t1);
}

// sums != nullptr: sums[b][c] = sum over the splits of image b = blockIdx.y;  else bias[c] = sum over splits AND images of
// the first component.  Block = 32 channels x 8 lanes over the partials, fixed-order tree in shared memory.
__global__ void __launch_bounds__(256) finish_sums_kernel(const float2* __restrict__ partial, int B, int splits, int C,
                                                          float2* __restrict__ sums, float* __restrict__ bias) {
    __shared__ float2 red[8][33];
    const int cl = threadIdx.x & 31, lane8 = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    float s = 0.f, t = 0.f;
    if (c < C) {
        const int first = sums ? blockIdx.y * splits : 0, count = sums ? splits : B * splits;
        for (int k = lane8; k < count; k += 8) {
            const float2 v = __ldg(partial + static_cast<size_t>(first + k) * C + c);
            s += v.x; t += v.y;
        }
    }
    red[lane8][cl] = make_float2(s, t);
    __syncthreads();
    if (lane8 == 0 && c < C) {
#pragma unroll
        for (int k = 1; k < 8; ++k) { s += red[k][cl].x; t += red[k][cl].y; }
        if (sums) sums[static_cast<size_t>(blockIdx.y) * C + c] = make_float2(s, t);
        else bias[c] = s;
    }
}

// dx[b][n][c] = rstd (g - s/N - x^ t/N) (+ add[b][n][c]);   four channels per thread
template <typename TX>
__global__ void __launch_bounds__(256) in_bwd_apply_kernel(const float* __restrict__ g, const TX* __restrict__ x,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           const float2* __restrict__ sums, const float* __restrict__ add, int B, int N,
                                                           int C, float* __restrict__ dx) {
    const int cv = C / 4;
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= static_cast<size_t>(B) * N * cv) return;
    const int c = static_cast<int>(i % cv) * 4;
    const int b = static_cast<int>(i / (static_cast<size_t>(cv) * N));
    const size_t off = (i / cv) * C + c;
    const float4 gv = __ldg(reinterpret_cast<const float4*>(g + off));
    const float gg[4] = {gv.x, gv.y, gv.z, gv.w};
    float av[4] = {0.f, 0.f, 0.f, 0.f};
    if (add) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(add + off));
        av[0] = a.x; av[1] = a.y; av[2] = a.z; av[3] = a.w;
    }
    const float invn = 1.f / static_cast<float>(N);
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float r = __ldg(rstd + static_cast<size_t>(b) * C + c + e);
        const float xh = (ldv(x + off + e) - __ldg(mean + static_cast<size_t>(b) * C + c + e)) * r;
        const float2 s = __ldg(sums + static_cast<size_t>(b) * C + c + e);
        o[e] = fmaf(r, gg[e] - s.x * invn - xh * s.y * invn, av[e]);
    }
    *reinterpret_cast<float4*>(dx + off) = make_float4(o[0], o[1], o[2], o[3]);
}

int sums_splits(int N) { return (N + TS_ROWS - 1) / TS_ROWS; }

}  // namespace

int launch_transpose_norm(const void* in, int in_dtype, int ld, int M, int Mpad, int C, int N, const float* mean, const float* rstd,
                          void* out, cudaStream_t s) {
    if (C % 2 != 0 || Mpad % 2 != 0 || ld % 2 != 0) {
        set_error("transpose_norm: C, ld and Mpad must be even (C=%d ld=%d Mpad=%d)", C, ld, Mpad);
        return MHADA_ERR_UNSUPPORTED;
    }
    const dim3 grid(static_cast<unsigned>((Mpad + 63) / 64), static_cast<unsigned>((C + 63) / 64));
    if (in_dtype == MHADA_F32)
        transpose_norm_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(in), ld, M, Mpad, C, N, mean, rstd,
                                                          static_cast<__nv_bfloat16*>(out));
    else
        transpose_norm_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(in), ld, M, Mpad, C, N, mean, rstd,
                                                                  static_cast<__nv_bfloat16*>(out));
    count_launch();
    return check_cuda(cudaGetLastError(), "transpose_norm launch");
}

int launch_blockdiag_t(const float* w, int H, int d, void* out, cudaStream_t s) {
    const size_t n = static_cast<size_t>(3) * H * d * H * d;
    blockdiag_t_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(w, H, d, static_cast<__nv_bfloat16*>(out));
    count_launch();
    return check_cuda(cudaGetLastError(), "blockdiag_t launch");
}

int launch_extract_blockdiag(const float* full, int H, int d, float* dw, cudaStream_t s) {
    const size_t n = static_cast<size_t>(H) * d * d;
    extract_blockdiag_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(full, H, d, dw);
    count_launch();
    return check_cuda(cudaGetLastError(), "extract_blockdiag launch");
}

size_t token_sums_workspace(int B, int N, int C) { return static_cast<size_t>(B) * sums_splits(N) * C * sizeof(float2); }

// sums[b][c] = (sum_n g, sum_n g IN(x)) (x may be null: second component 0); bias (optional) [C] = sum over images too.
// g: f32 or bf16 [B, N, C]; x: bf16 [B, N, C].
int launch_token_sums(const void* g, int g_dtype, const void* x, const float* mean, const float* rstd, int B, int N, int C,
                      void* partial, void* sums, float* bias, cudaStream_t s) {
    if (C % 2 != 0 || C > 2048) {
        set_error("token_sums: C must be even and <= 2048, got %d", C);
        return MHADA_ERR_UNSUPPORTED;
    }
    if ((sums != nullptr) == (bias != nullptr)) {
        set_error("token_sums: exactly one of sums / bias");
        return MHADA_ERR_ARG;
    }
    const int splits = sums_splits(N);
    const dim3 grid(static_cast<unsigned>(splits), static_cast<unsigned>(B));
    const int threads = (C / 2 + 31) / 32 * 32;
    float2* part = static_cast<float2*>(partial);
    const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x);
    if (g_dtype == MHADA_F32)
        token_sums_kernel<float, __nv_bfloat16><<<grid, threads, 0, s>>>(static_cast<const float*>(g), xb, mean, rstd, N, C, part);
    else
        token_sums_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, threads, 0, s>>>(static_cast<const __nv_bfloat16*>(g), xb, mean, rstd, N,
                                                                                 C, part);
    count_launch();
    const dim3 fgrid(static_cast<unsigned>((C + 31) / 32), static_cast<unsigned>(sums ? B : 1));
    finish_sums_kernel<<<fgrid, 256, 0, s>>>(part, B, splits, C, static_cast<float2*>(sums), bias);
    count_launch();
    return check_cuda(cudaGetLastError(), "token_sums launch");
}

// dx = IN-backward of g through x (bf16) (+ add), f32 out
int launch_in_bwd_apply(const float* g, const void* x, const float* mean, const float* rstd, const void* sums, const float* add,
                        int B, int N, int C, float* dx, cudaStream_t s) {
    const size_t n = static_cast<size_t>(B) * N * (C / 4);
    in_bwd_apply_kernel<__nv_bfloat16><<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(
        g, static_cast<const __nv_bfloat16*>(x), mean, rstd, static_cast<const float2*>(sums), add, B, N, C, dx);
    count_launch();
    return check_cuda(cudaGetLastError(), "in_bwd_apply launch");
}

}  // namespace mh

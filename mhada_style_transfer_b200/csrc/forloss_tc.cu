// AdaAttnForLoss on the tensor cores: the parameter-free attention of the local-feature loss,
// MHAdaSTr/network/adaDecoder.py:38-81 (called by lossfn.py:26-34 on VGG features): Q = IN(c_1x), K = IN(s_1x), V = s_x,
//     A = softmax(Q K^T);  M = A V;  S = sqrt(clamp(A V^2 - M^2, 1e-6));  out = S * IN(c_x) + M
// with d_qk = 448 / 960 / 1472 (concatenated VGG levels) != d_v = 256 / 512.
//
// The streaming kernel (attn_tc.cu) keeps Q of two query tiles and a K ring in shared memory: at d_qk >= 256 they no
// longer fit (r1 ran these shapes on the fp32 SIMT kernel: 52 ms for relu3_1 at batch 8).  These problems are small in
// tokens (N <= 4096 at the 256 x 256 training resolution) and wide in channels, so here the logits ARE materialised,
// per image, in HBM/L2 -- but every contraction runs on tcgen05 through the token GEMM (gemm_tc.cu):
//   1. statistics of c_1x, s_1x, c_x, s_x (stats.cu);  Qn = log2(e) * IN(c_1x), Kn = IN(s_1x) as bf16, key rows
//      padded to a multiple of 128 with zeros (normalize_rows_kernel)
//   2. V' = [V - mu_v | (V - mu_v)^2]^T as bf16 [2 dv][Ns_pad] (vprime_t_kernel; centring as in the layer path)
//   3. per image: S = Qn Kn^T (f32 [Nc][Ns_pad], GEMM)  ->  P = 2^(S - rowmax) as bf16, row sums of the ROUNDED
//      weights (softmax_rows_kernel)  ->  O = P V'^T (f32 [Nc][2 dv], GEMM)
//   4. out = sqrt(max(E - M^2, 1e-6)) * IN(c_x) + M + mu_v  (forloss_finalize_kernel)
// FLOPs 2 Nc Ns (d_qk + 2 d_v) per image; the materialised S / P of one image (<= 100 MB) are reused image by image.
#include "common.h"
#include "ptx.cuh"

namespace mh {

constexpr float FL_LOG2E = 1.4426950408889634f;

// y[b, n, c] = scale * (x[b, n, c] - mean[b, c]) * rstd[b, c]  as bf16, rows n in [N, Npad) zero.  8 channels per thread.
__global__ void __launch_bounds__(256) normalize_rows_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ mean,
                                                             const float* __restrict__ rstd, float scale, int B, int N, int Npad,
                                                             int C, __nv_bfloat16* __restrict__ y) {
    const int cv = C / 8;
    const size_t total = static_cast<size_t>(B) * Npad * cv;
    const size_t t = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (t >= total) return;
    const int c = static_cast<int>(t % cv) * 8;
    const int n = static_cast<int>((t / cv) % Npad);
    const int b = static_cast<int>(t / (static_cast<size_t>(cv) * Npad));
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (n < N) {
        const uint4 w = __ldg(reinterpret_cast<const uint4*>(x + (static_cast<size_t>(b) * N + n) * C + c));
        const float4 m0 = __ldg(reinterpret_cast<const float4*>(mean + static_cast<size_t>(b) * C + c));
        const float4 m1 = __ldg(reinterpret_cast<const float4*>(mean + static_cast<size_t>(b) * C + c) + 1);
        const float4 r0 = __ldg(reinterpret_cast<const float4*>(rstd + static_cast<size_t>(b) * C + c));
        const float4 r1 = __ldg(reinterpret_cast<const float4*>(rstd + static_cast<size_t>(b) * C + c) + 1);
        o.x = pack_bf16x2((bf16_lo(w.x) - m0.x) * r0.x * scale, (bf16_hi(w.x) - m0.y) * r0.y * scale);
        o.y = pack_bf16x2((bf16_lo(w.y) - m0.z) * r0.z * scale, (bf16_hi(w.y) - m0.w) * r0.w * scale);
        o.z = pack_bf16x2((bf16_lo(w.z) - m1.x) * r1.x * scale, (bf16_hi(w.z) - m1.y) * r1.y * scale);
        o.w = pack_bf16x2((bf16_lo(w.w) - m1.z) * r1.z * scale, (bf16_hi(w.w) - m1.w) * r1.w * scale);
    }
    *reinterpret_cast<uint4*>(y + (static_cast<size_t>(b) * Npad + n) * C + c) = o;
}

// vt[b][c][n] = v[b, n, c] - mu[b, c];  vt[b][dv + c][n] = (that)^2 (squared in f32);  n >= N -> 0.  32 x 32 tiles
// through shared memory: reads are coalesced along channels, writes along tokens.
__global__ void __launch_bounds__(256) vprime_t_kernel(const __nv_bfloat16* __restrict__ v, const float* __restrict__ mu, int N,
                                                       int Npad, int dv, __nv_bfloat16* __restrict__ vt) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int n = n0 + r, c = c0 + tx;
        float val = 0.f;
        if (n < N && c < dv) val = __bfloat162float(v[(static_cast<size_t>(b) * N + n) * dv + c]) - __ldg(mu + static_cast<size_t>(b) * dv + c);
        tile[r][tx] = val;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r, n = n0 + tx;
        if (c < dv && n < Npad) {
            const float val = (n < N) ? tile[tx][r] : 0.f;
            __nv_bfloat16* o = vt + (static_cast<size_t>(b) * 2 * dv + c) * Npad + n;
            o[0] = __float2bfloat16_rn(val);
            o[static_cast<size_t>(dv) * Npad] = __float2bfloat16_rn(val * val);
        }
    }
}

// One CTA per query row: p[j] = 2^(s[j] - max_j s) as bf16 for j < Ns (0 for the padding), lsum = sum of the ROUNDED
// weights (what the second GEMM multiplies; see attn_tc.cu on why the sum must use the rounded values).
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, int Ns, int Npad,
                                                           __nv_bfloat16* __restrict__ p, float* __restrict__ lsum) {
    __shared__ float red[8];
    const float* row = s + static_cast<size_t>(blockIdx.x) * Npad;
    __nv_bfloat16* prow = p + static_cast<size_t>(blockIdx.x) * Npad;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float mx = -INFINITY;
    for (int j = threadIdx.x * 4; j < Ns; j += 1024) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(row + j));
        mx = fmaxf(mx, v.x);
        if (j + 1 < Ns) mx = fmaxf(mx, v.y);
        if (j + 2 < Ns) mx = fmaxf(mx, v.z);
        if (j + 3 < Ns) mx = fmaxf(mx, v.w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float sum = 0.f;
    for (int j = threadIdx.x * 4; j < Npad; j += 1024) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(row + j));
        const float e0 = j < Ns ? ex2_approx(v.x - mx) : 0.f, e1 = j + 1 < Ns ? ex2_approx(v.y - mx) : 0.f;
        const float e2 = j + 2 < Ns ? ex2_approx(v.z - mx) : 0.f, e3 = j + 3 < Ns ? ex2_approx(v.w - mx) : 0.f;
        const uint32_t lo = pack_bf16x2(e0, e1), hi = pack_bf16x2(e2, e3);
        sum += (bf16_lo(lo) + bf16_hi(lo)) + (bf16_lo(hi) + bf16_hi(hi));
        *reinterpret_cast<uint2*>(prow + j) = make_uint2(lo, hi);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w];
        lsum[blockIdx.x] = t;
    }
}

// out[n, c] = sqrt(max(E - M^2, 1e-6)) * (x[n, c] - mean_x[c]) * rstd_x[c] + M + mu_v[c],  M = o[n, c] / l, E = o[n, dv + c] / l
__global__ void __launch_bounds__(256) forloss_finalize_kernel(const float* __restrict__ o, const float* __restrict__ lsum,
                                                               const __nv_bfloat16* __restrict__ x, const float* __restrict__ mean_x,
                                                               const float* __restrict__ rstd_x, const float* __restrict__ mu_v,
                                                               int Nc, int dv, __nv_bfloat16* __restrict__ out) {
    const int cv = dv / 2;
    const size_t t = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (t >= static_cast<size_t>(Nc) * cv) return;
    const int c = static_cast<int>(t % cv) * 2, n = static_cast<int>(t / cv);
    const float inv = 1.f / __ldg(lsum + n);
    const float2 m2 = __ldg(reinterpret_cast<const float2*>(o + static_cast<size_t>(n) * 2 * dv + c));
    const float2 e2 = __ldg(reinterpret_cast<const float2*>(o + static_cast<size_t>(n) * 2 * dv + dv + c));
    const uint32_t xw = __ldg(reinterpret_cast<const uint32_t*>(x + static_cast<size_t>(n) * dv + c));
    float r[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const float m = (e ? m2.y : m2.x) * inv, ex = (e ? e2.y : e2.x) * inv;
        const float sd = sqrtf(fmaxf(fmaf(-m, m, ex), 1e-6f));
        const float xf = e ? bf16_hi(xw) : bf16_lo(xw);
        r[e] = fmaf(sd, (xf - __ldg(mean_x + c + e)) * __ldg(rstd_x + c + e), m + __ldg(mu_v + c + e));
    }
    *reinterpret_cast<uint32_t*>(out + static_cast<size_t>(n) * dv + c) = pack_bf16x2(r[0], r[1]);
}

struct ForlossWs {
    float *mean_q, *rstd_q, *mean_k, *rstd_k, *mean_x, *rstd_x, *mean_v, *rstd_v, *stats_ws, *s, *o, *lsum;
    __nv_bfloat16 *qn, *kn, *vt, *p;
    size_t total;
};
static ForlossWs forloss_carve(int B, int Nc, int Ns, int dqk, int dv, uint8_t* base) {
    ForlossWs w;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void* p = base ? base + off : nullptr;
        off += align_up(bytes, 1024);
        return p;
    };
    const int Npad = (Ns + 127) / 128 * 128;
    const size_t sq = static_cast<size_t>(B) * dqk * 4, sv = static_cast<size_t>(B) * dv * 4;
    w.mean_q = static_cast<float*>(take(sq)); w.rstd_q = static_cast<float*>(take(sq));
    w.mean_k = static_cast<float*>(take(sq)); w.rstd_k = static_cast<float*>(take(sq));
    w.mean_x = static_cast<float*>(take(sv)); w.rstd_x = static_cast<float*>(take(sv));
    w.mean_v = static_cast<float*>(take(sv)); w.rstd_v = static_cast<float*>(take(sv));
    size_t sb = stats_workspace(B, Nc > Ns ? Nc : Ns, dqk > dv ? dqk : dv);
    const size_t sb2 = stats_workspace(B, Nc < Ns ? Nc : Ns, dqk > dv ? dqk : dv);
    if (sb2 > sb) sb = sb2;
    w.stats_ws = static_cast<float*>(take(sb));
    w.qn = static_cast<__nv_bfloat16*>(take(static_cast<size_t>(B) * Nc * dqk * 2));
    w.kn = static_cast<__nv_bfloat16*>(take(static_cast<size_t>(B) * Npad * dqk * 2));
    w.vt = static_cast<__nv_bfloat16*>(take(static_cast<size_t>(B) * 2 * dv * Npad * 2));
    w.s = static_cast<float*>(take(static_cast<size_t>(Nc) * Npad * 4));           // one image at a time
    w.p = static_cast<__nv_bfloat16*>(take(static_cast<size_t>(Nc) * Npad * 2));
    w.o = static_cast<float*>(take(static_cast<size_t>(Nc) * 2 * dv * 4));
    w.lsum = static_cast<float*>(take(static_cast<size_t>(Nc) * 4));
    w.total = off;
    return w;
}
size_t forloss_workspace(int B, int Nc, int Ns, int dqk, int dv) { return forloss_carve(B, Nc, Ns, dqk, dv, nullptr).total; }

int forloss_forward(const mhada_forloss_args& a, cudaStream_t s) {
    const int B = a.B, Nc = a.Nc, Ns = a.Ns, dqk = a.dqk, dv = a.dv;
    const int Npad = (Ns + 127) / 128 * 128;
    ForlossWs w = forloss_carve(B, Nc, Ns, dqk, dv, static_cast<uint8_t*>(a.ws));
    // 1. statistics (adaDecoder.py:55, :60, :81; the mean of V for the centring)
    if (int e = launch_stats(a.c_1x, MHADA_BF16, B, Nc, dqk, dqk, w.mean_q, w.rstd_q, w.stats_ws, s)) return e;
    if (int e = launch_stats(a.s_1x, MHADA_BF16, B, Ns, dqk, dqk, w.mean_k, w.rstd_k, w.stats_ws, s)) return e;
    if (int e = launch_stats(a.c_x, MHADA_BF16, B, Nc, dv, dv, w.mean_x, w.rstd_x, w.stats_ws, s)) return e;
    if (int e = launch_stats(a.s_x, MHADA_BF16, B, Ns, dv, dv, w.mean_v, w.rstd_v, w.stats_ws, s)) return e;
    // 2. normalised operands
    {
        const size_t tq = static_cast<size_t>(B) * Nc * (dqk / 8), tk = static_cast<size_t>(B) * Npad * (dqk / 8);
        normalize_rows_kernel<<<static_cast<unsigned>((tq + 255) / 256), 256, 0, s>>>(
            static_cast<const __nv_bfloat16*>(a.c_1x), w.mean_q, w.rstd_q, FL_LOG2E, B, Nc, Nc, dqk, w.qn);
        count_launch();
        normalize_rows_kernel<<<static_cast<unsigned>((tk + 255) / 256), 256, 0, s>>>(
            static_cast<const __nv_bfloat16*>(a.s_1x), w.mean_k, w.rstd_k, 1.f, B, Ns, Npad, dqk, w.kn);
        count_launch();
        dim3 g(static_cast<unsigned>(Npad / 32), static_cast<unsigned>((dv + 31) / 32), static_cast<unsigned>(B));
        vprime_t_kernel<<<g, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(a.s_x), w.mean_v, Ns, Npad, dv, w.vt);
        count_launch();
        if (int e = check_cuda(cudaGetLastError(), "forloss prepare launch")) return e;
    }
    // 3. per image: logits, weights, moments; 4. epilogue
    for (int b = 0; b < B; ++b) {
        GemmDesc g{};
        g.a = w.qn + static_cast<size_t>(b) * Nc * dqk; g.lda = dqk;
        g.w = w.kn + static_cast<size_t>(b) * Npad * dqk; g.ldw = dqk;
        g.M = Nc; g.N = Npad; g.K = dqk; g.out_f32 = w.s; g.ldf = Npad;
        if (int e = launch_gemm_bf16(g, s)) return e;                                              // :70 (bmm in Softmax)
        softmax_rows_kernel<<<Nc, 256, 0, s>>>(w.s, Ns, Npad, w.p, w.lsum);
        count_launch();
        g = GemmDesc{};
        g.a = w.p; g.lda = Npad;
        g.w = w.vt + static_cast<size_t>(b) * 2 * dv * Npad; g.ldw = Npad;
        g.M = Nc; g.N = 2 * dv; g.K = Npad; g.out_f32 = w.o; g.ldf = 2 * dv;
        if (int e = launch_gemm_bf16(g, s)) return e;                                              // :71, :74
        const size_t tf = static_cast<size_t>(Nc) * (dv / 2);
        forloss_finalize_kernel<<<static_cast<unsigned>((tf + 255) / 256), 256, 0, s>>>(
            w.o, w.lsum, static_cast<const __nv_bfloat16*>(a.c_x) + static_cast<size_t>(b) * Nc * dv, w.mean_x + static_cast<size_t>(b) * dv,
            w.rstd_x + static_cast<size_t>(b) * dv, w.mean_v + static_cast<size_t>(b) * dv, Nc, dv,
            static_cast<__nv_bfloat16*>(a.out) + static_cast<size_t>(b) * Nc * dv);                 // :74-81
        count_launch();
    }
    return check_cuda(cudaGetLastError(), "forloss launch");
}

}  // namespace mh

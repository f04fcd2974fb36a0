// The fp32 path: true-fp32 FFMA kernels (never TF32) that reproduce the reference's arithmetic
// for BASELINE.json configs[0] and serve as the on-device yardstick for the tensor-core path.
//
//   grouped_linear_f32 : per-head 1x1 convs f/g/h with the instance norm applied on load
//                        (adaDecoder.py:173, :178, :182) and out_conv (adaDecoder.py:205)
//   attn_f32           : streaming softmax(Q K^T) [V, V^2] + sqrt(max(E - M^2, 1e-6)) * IN(x) + M
//                        (adaDecoder.py:186-198; AdaAttnForLoss :70-81), any dqk / dv
//
// Both are shared-memory tiled SIMT kernels (64x64 output tile, 4x4 register micro-tile).
#include <math.h>

#include "common.h"
#include "ptx.cuh"

namespace mh {

constexpr int TILE = 64;
constexpr int KC = 32;  // reduction chunk staged in shared memory

// Compensated (Kahan) add of a short fp32 partial sum into a long-running one.  Dot products are
// accumulated in short chunks (8 or 32 terms, small magnitude) and the chunks are folded in with the
// rounding error carried along: the logits / out_conv sums then sit ~10x closer to the float64 result
// than a single 64..512-term fp32 chain, which is what the <= 1e-3 max-abs budget of the fp32 path needs
// (the reference's CPU sgemm gets the same effect from its many SIMD-lane accumulators).
__device__ __forceinline__ void kahan_add(float& sum, float& comp, float term) {
    float y = term - comp;
    float u = sum + y;
    comp = (u - sum) - y;
    sum = u;
}

// ------------------------------------------------------------------------------------------------
// y[m, g*dout + o] = sum_i w[g][o][i] * xin[m, g*din + i] + bias[g][o]
//   mode 0: xin = x   mode 1: xin = (x - mean) * rstd   mode 2: xin = x - mean  (bias ignored -> centred V)
// rows m = b * rows_per_batch + n;  mean / rstd are [B, G*din].
// grid (ceil(M/64), ceil(dout/64), G)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) grouped_linear_f32_kernel(const float* __restrict__ x, int ldx,
                                                                 const float* __restrict__ w,
                                                                 const float* __restrict__ bias,
                                                                 const float* __restrict__ mean,
                                                                 const float* __restrict__ rstd, int mode, int M,
                                                                 int rows_per_batch, int din, int dout,
                                                                 float* __restrict__ y, int ldy) {
    __shared__ float xs[KC][TILE + 1];
    __shared__ float ws[KC][TILE + 1];
    const int g = blockIdx.z;
    const int m0 = blockIdx.x * TILE, o0 = blockIdx.y * TILE;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const float* wg = w + static_cast<size_t>(g) * dout * din;
    const int cin0 = g * din;

    float acc[4][4], cmp[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = cmp[i][j] = 0.f;

    for (int k0 = 0; k0 < din; k0 += KC) {
        // stage x tile [64 rows][32 k] and w tile [64 outs][32 k]; consecutive threads walk k (contiguous)
        for (int e = threadIdx.x; e < TILE * KC; e += 256) {
            int r = e / KC, kk = e % KC;
            int m = m0 + r, k = k0 + kk;
            float xv = 0.f;
            if (m < M && k < din) {
                xv = __ldg(x + static_cast<size_t>(m) * ldx + cin0 + k);
                if (mode != 0) {
                    int b = m / rows_per_batch;
                    float mu = __ldg(mean + static_cast<size_t>(b) * gridDim.z * din + cin0 + k);
                    xv -= mu;
                    if (mode == 1) xv *= __ldg(rstd + static_cast<size_t>(b) * gridDim.z * din + cin0 + k);
                }
            }
            xs[kk][r] = xv;
            int o = o0 + r;
            ws[kk][r] = (o < dout && k < din) ? __ldg(wg + static_cast<size_t>(o) * din + k) : 0.f;
        }
        __syncthreads();
        float part[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) part[i][j] = 0.f;
#pragma unroll 8
        for (int kk = 0; kk < KC; ++kk) {
            float a[4], bb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = xs[kk][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bb[j] = ws[kk][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) part[i][j] = fmaf(a[i], bb[j], part[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) kahan_add(acc[i][j], cmp[i][j], part[i][j]);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty + 16 * i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int o = o0 + tx + 16 * j;
            if (o >= dout) continue;
            float bv = (mode == 2 || bias == nullptr) ? 0.f : __ldg(bias + static_cast<size_t>(g) * dout + o);
            y[static_cast<size_t>(m) * ldy + g * dout + o] = acc[i][j] + bv;
        }
    }
}

// mu_v[b, g*d + o] = sum_i wh[g][o][i] * mean_s[b, g*d + i] + bh[g][o]      (tiny)
__global__ void muv_kernel(const float* __restrict__ wh, const float* __restrict__ bh,
                           const float* __restrict__ mean_s, int B, int H, int d, float* __restrict__ mu_v) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int C = H * d;
    if (i >= B * C) return;
    int b = i / C, c = i % C, g = c / d, o = c % d;
    const float* wr = wh + (static_cast<size_t>(g) * d + o) * d;
    const float* mu = mean_s + static_cast<size_t>(b) * C + g * d;
    float a = 0.f;
    for (int k = 0; k < d; ++k) a = fmaf(wr[k], mu[k], a);
    mu_v[i] = a + bh[g * d + o];
}

int launch_proj_f32(int parts, const float* fc, const float* fs, const float* mean_c, const float* rstd_c,
                    const float* mean_s, const float* rstd_s, const float* w, const float* bias, int B, int Bs, int Nc,
                    int Ns, int H, int d, float* q, float* k, float* v, float* mu_v, cudaStream_t s) {
    const int C = H * d;
    const size_t wsz = static_cast<size_t>(H) * d * d, bsz = static_cast<size_t>(H) * d;
    if (parts & 1) {
        dim3 gq((B * Nc + TILE - 1) / TILE, (d + TILE - 1) / TILE, H);
        grouped_linear_f32_kernel<<<gq, 256, 0, s>>>(fc, C, w, bias, mean_c, rstd_c, 1, B * Nc, Nc, d, d, q, C);
        count_launch();
    }
    if (parts & 2) {
        dim3 gk((Bs * Ns + TILE - 1) / TILE, (d + TILE - 1) / TILE, H);
        grouped_linear_f32_kernel<<<gk, 256, 0, s>>>(fs, C, w + wsz, bias + bsz, mean_s, rstd_s, 1, Bs * Ns, Ns, d, d, k, C);
        count_launch();
        grouped_linear_f32_kernel<<<gk, 256, 0, s>>>(fs, C, w + 2 * wsz, nullptr, mean_s, nullptr, 2, Bs * Ns, Ns, d, d, v, C);
        count_launch();
        muv_kernel<<<(Bs * C + 255) / 256, 256, 0, s>>>(w + 2 * wsz, bias + 2 * bsz, mean_s, Bs, H, d, mu_v);
        count_launch();
    }
    return check_cuda(cudaGetLastError(), "proj_f32 launch");
}

int launch_linear_f32(const float* x, int ldx, const float* w, const float* bias, int M, int Cin, int Cout, float* y,
                      int ldy, cudaStream_t s) {
    dim3 g((M + TILE - 1) / TILE, (Cout + TILE - 1) / TILE, 1);
    grouped_linear_f32_kernel<<<g, 256, 0, s>>>(x, ldx, w, bias, nullptr, nullptr, 0, M, M, Cin, Cout, y, ldy);
    count_launch();
    return check_cuda(cudaGetLastError(), "linear_f32 launch");
}

// ------------------------------------------------------------------------------------------------
// Streaming attention, fp32.  grid (ceil(Nc/64), ceil(dv/64), B*H), 256 threads as 16x16;
// thread (ty, tx) owns rows ty+16i and columns tx+16j of every 64x64 tile.
// ------------------------------------------------------------------------------------------------
struct AttnF32Params {
    const float *q, *k, *v, *x;
    float* out;
    const float *x_mean, *x_rstd, *mu_v, *q_mean, *q_rstd, *k_mean, *k_rstd;
    int H, Nc, Ns, dqk, dv, ldq, ldk, ldv, ldx, ldo;
    int kv_shared;   // 1: one K / V / mu_v set (style) serves every image of the batch
};

// COSINE = false: A = softmax(Q K^T)                                   (Softmax, adaDecoder.py:11-17)
// COSINE = true : A = (cos(q_i, k_j) + 1) / sum_j (cos(q_i, k_j) + 1)   (CosineSimilarity, adaDecoder.py:20-34);
//                 |q_i|^2 and |k_j|^2 are accumulated next to the dot products, no running maximum is needed.
template <bool COSINE>
__global__ void __launch_bounds__(256, 2) attn_f32_kernel(const AttnF32Params p) {
    __shared__ float Ps[TILE][TILE + 1];
    __shared__ union {
        struct {
            float q[KC][TILE + 1];
            float k[KC][TILE + 1];
        } qk;
        float v[TILE][TILE + 1];
    } sm;

    const int bh = blockIdx.z, b = bh / p.H, h = bh % p.H;
    const int q0 = blockIdx.x * TILE, c0 = blockIdx.y * TILE;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const float* qb = p.q + static_cast<size_t>(b) * p.Nc * p.ldq + h * p.dqk;
    const int bkv = p.kv_shared ? 0 : b;
    const float* kb = p.k + static_cast<size_t>(bkv) * p.Ns * p.ldk + h * p.dqk;
    const float* vb = p.v + static_cast<size_t>(bkv) * p.Ns * p.ldv + h * p.dv;
    const float* qmu = p.q_mean ? p.q_mean + static_cast<size_t>(b) * p.H * p.dqk + h * p.dqk : nullptr;
    const float* qrs = p.q_mean ? p.q_rstd + static_cast<size_t>(b) * p.H * p.dqk + h * p.dqk : nullptr;
    const float* kmu = p.k_mean ? p.k_mean + static_cast<size_t>(bkv) * p.H * p.dqk + h * p.dqk : nullptr;
    const float* krs = p.k_mean ? p.k_rstd + static_cast<size_t>(bkv) * p.H * p.dqk + h * p.dqk : nullptr;

    // running (unnormalised) A.V and A.V^2 with their Kahan compensation terms: a plain fp32 chain over
    // Ns = 4096..32400 keys drifts by ~sqrt(Ns) ulp, which alone would eat the 1e-3 max-abs budget
    float om[4][4], oe[4][4], cm[4][4], ce[4][4], mrow[4], lrow[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        mrow[i] = -INFINITY;
        lrow[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) om[i][j] = oe[i][j] = cm[i][j] = ce[i][j] = 0.f;
    }

    for (int k0 = 0; k0 < p.Ns; k0 += TILE) {
        // ---- S = Q K^T for this key tile
        float s[4][4], sc[4][4];
        float qq[4] = {0.f, 0.f, 0.f, 0.f}, kk2[4] = {0.f, 0.f, 0.f, 0.f};      // squared norms (COSINE)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[i][j] = sc[i][j] = 0.f;
        for (int d0 = 0; d0 < p.dqk; d0 += KC) {
            for (int e = threadIdx.x; e < TILE * KC; e += 256) {
                int r = e / KC, dd = e % KC, d = d0 + dd;
                float qv = 0.f, kv = 0.f;
                if (d < p.dqk) {
                    if (q0 + r < p.Nc) {
                        qv = __ldg(qb + static_cast<size_t>(q0 + r) * p.ldq + d);
                        if (qmu) qv = (qv - __ldg(qmu + d)) * __ldg(qrs + d);
                    }
                    if (k0 + r < p.Ns) {
                        kv = __ldg(kb + static_cast<size_t>(k0 + r) * p.ldk + d);
                        if (kmu) kv = (kv - __ldg(kmu + d)) * __ldg(krs + d);
                    }
                }
                sm.qk.q[dd][r] = qv;
                sm.qk.k[dd][r] = kv;
            }
            __syncthreads();
#pragma unroll 1
            for (int d1 = 0; d1 < KC; d1 += 8) {
                float part[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) part[i][j] = 0.f;
#pragma unroll
                for (int dd = d1; dd < d1 + 8; ++dd) {
                    float a[4], bb[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) a[i] = sm.qk.q[dd][ty + 16 * i];
#pragma unroll
                    for (int j = 0; j < 4; ++j) bb[j] = sm.qk.k[dd][tx + 16 * j];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) part[i][j] = fmaf(a[i], bb[j], part[i][j]);
                    if (COSINE) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            qq[i] = fmaf(a[i], a[i], qq[i]);
                            kk2[i] = fmaf(bb[i], bb[i], kk2[i]);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) kahan_add(s[i][j], sc[i][j], part[i][j]);
            }
            __syncthreads();
        }
        if (COSINE) {
            // ---- weights cos + 1 (>= 0): plain running sums, nothing to rescale
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float rs = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float e = k0 + tx + 16 * j < p.Ns ? s[i][j] / (sqrtf(qq[i]) * sqrtf(kk2[j])) + 1.f : 0.f;
                    rs += e;
                    Ps[ty + 16 * i][tx + 16 * j] = e;
                }
#pragma unroll
                for (int o = 8; o >= 1; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
                lrow[i] += rs;
            }
        }
        // ---- online softmax (rows are shared by the 16 threads with equal ty = one half warp)
#pragma unroll
        for (int i = 0; i < 4 && !COSINE; ++i) {
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (k0 + tx + 16 * j >= p.Ns) s[i][j] = -INFINITY;
                mx = fmaxf(mx, s[i][j]);
            }
#pragma unroll
            for (int o = 8; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            float mnew = fmaxf(mrow[i], mx);  // finite: every tile has >= 1 valid key
            float scale = expf(mrow[i] - mnew);
            float rs = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float e = expf(s[i][j] - mnew);
                rs += e;
                Ps[ty + 16 * i][tx + 16 * j] = e;
            }
#pragma unroll
            for (int o = 8; o >= 1; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
            lrow[i] = lrow[i] * scale + rs;   // 64..500 additions of O(1..64) terms: ~1e-6 relative at worst
            mrow[i] = mnew;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                om[i][j] *= scale; cm[i][j] *= scale;
                oe[i][j] *= scale; ce[i][j] *= scale;
            }
        }
        // ---- stage V tile [64 keys][64 value columns]
        for (int e = threadIdx.x; e < TILE * TILE; e += 256) {
            int r = e / TILE, c = e % TILE;
            float vv = 0.f;
            if (k0 + r < p.Ns && c0 + c < p.dv) vv = __ldg(vb + static_cast<size_t>(k0 + r) * p.ldv + c0 + c);
            sm.v[r][c] = vv;
        }
        __syncthreads();
        {
            float pm[4][4], pe[4][4];       // this tile's 64-key partial sums
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) pm[i][j] = pe[i][j] = 0.f;
#pragma unroll 4
            for (int kk = 0; kk < TILE; ++kk) {
                float a[4], vv[4], v2[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = Ps[ty + 16 * i][kk];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    vv[j] = sm.v[kk][tx + 16 * j];
                    v2[j] = vv[j] * vv[j];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        pm[i][j] = fmaf(a[i], vv[j], pm[i][j]);
                        pe[i][j] = fmaf(a[i], v2[j], pe[i][j]);
                    }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    kahan_add(om[i][j], cm[i][j], pm[i][j]);
                    kahan_add(oe[i][j], ce[i][j], pe[i][j]);
                }
        }
        __syncthreads();
    }
    // ---- epilogue: S * IN(x) + M   (adaDecoder.py:190-198)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int n = q0 + ty + 16 * i;
        if (n >= p.Nc) continue;
        float inv = 1.f / lrow[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int c = c0 + tx + 16 * j;
            if (c >= p.dv) continue;
            int ch = h * p.dv + c;
            float m = om[i][j] * inv, e = oe[i][j] * inv;
            float sd = sqrtf(fmaxf(e - m * m, 1e-6f));
            size_t sidx = static_cast<size_t>(b) * p.H * p.dv + ch;
            float xn = (__ldg(p.x + (static_cast<size_t>(b) * p.Nc + n) * p.ldx + ch) - __ldg(p.x_mean + sidx)) *
                       __ldg(p.x_rstd + sidx);
            float mu = p.mu_v ? __ldg(p.mu_v + static_cast<size_t>(bkv) * p.H * p.dv + ch) : 0.f;
            p.out[(static_cast<size_t>(b) * p.Nc + n) * p.ldo + ch] = fmaf(sd, xn, m + mu);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Cosine activation in CLOSED FORM (CosineSimilarity, adaDecoder.py:20-34), head_dim 64:
//     a_ij = (q^_i . k^_j + 1) / sum_j (q^_i . k^_j + 1)        q^ = q / |q|, k^ = k / |k|
//  => A V' = (q^_i . T + sV) / (q^_i . sK + Ns),   T = sum_j k^_j (x) v'_j  [64 x 128],  sK = sum_j k^_j,  sV = sum_j v'_j
// i.e. O(N d^2) instead of O(N^2 d): one pass over the keys builds the moments of an (image, head), one pass over the
// queries applies them (the N x N kernel above stays for other widths and for normalise-on-load).
// ------------------------------------------------------------------------------------------------
constexpr int COS_D = 64, COS_MOM = COS_D * 2 * COS_D + COS_D + 2 * COS_D;      // floats per (image, head)

size_t attn_cosine_scratch_bytes(int B, int H) { return static_cast<size_t>(B) * H * COS_MOM * sizeof(float); }

// grid (H, Bkv), 256 threads = 8 x 32: thread (ty, tx) owns rows ty*8.. of k^ and columns tx*4.. of v' = [v~ | v~^2]
__global__ void __launch_bounds__(256) cosine_moments_kernel(const AttnF32Params p, float* __restrict__ scratch) {
    __shared__ float ks[32][COS_D + 1];
    __shared__ __align__(16) float vs[32][2 * COS_D];
    const int h = blockIdx.x, b = blockIdx.y;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float* kb = p.k + static_cast<size_t>(b) * p.Ns * p.ldk + h * COS_D;
    const float* vb = p.v + static_cast<size_t>(b) * p.Ns * p.ldv + h * COS_D;
    float acc[8][4], sk[8], sv[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        sk[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) sv[j] = 0.f;
    for (int k0 = 0; k0 < p.Ns; k0 += 32) {
        // warp ty stages keys ty*4 .. ty*4+3: lane = two of the 64 channels; k is normalised here
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = ty * 4 + u, n = k0 + r;
            float k0v = 0.f, k1v = 0.f, v0 = 0.f, v1 = 0.f;
            if (n < p.Ns) {
                k0v = __ldg(kb + static_cast<size_t>(n) * p.ldk + tx); k1v = __ldg(kb + static_cast<size_t>(n) * p.ldk + tx + 32);
                v0 = __ldg(vb + static_cast<size_t>(n) * p.ldv + tx); v1 = __ldg(vb + static_cast<size_t>(n) * p.ldv + tx + 32);
            }
            float nn = fmaf(k0v, k0v, k1v * k1v);
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
            const float inv = n < p.Ns ? 1.f / sqrtf(nn) : 0.f;
            ks[r][tx] = k0v * inv; ks[r][tx + 32] = k1v * inv;
            vs[r][tx] = v0; vs[r][tx + 32] = v1;
            vs[r][COS_D + tx] = v0 * v0; vs[r][COS_D + tx + 32] = v1 * v1;
        }
        __syncthreads();
#pragma unroll 4
        for (int kk = 0; kk < 32; ++kk) {
            const float4 bv = *reinterpret_cast<const float4*>(&vs[kk][tx * 4]);
            const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float a = ks[kk][ty * 8 + i];
                if (tx == 0) sk[i] += a;
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a, bb[j], acc[i][j]);
            }
            if (ty == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) sv[j] += bb[j];
            }
        }
        __syncthreads();
    }
    float* out = scratch + (static_cast<size_t>(b) * p.H + h) * COS_MOM;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4*>(out + (ty * 8 + i) * 2 * COS_D + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    if (tx == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) out[COS_D * 2 * COS_D + ty * 8 + i] = sk[i];
    }
    if (ty == 0) *reinterpret_cast<float4*>(out + COS_D * 2 * COS_D + COS_D + tx * 4) = make_float4(sv[0], sv[1], sv[2], sv[3]);
}

// grid (ceil(Nc / 64), H, B), 256 threads: warp w applies the moments to query rows w, w + 8, ...; lane = 4 of the 128 columns
__global__ void __launch_bounds__(256) cosine_apply_kernel(const AttnF32Params p, const float* __restrict__ scratch) {
    __shared__ __align__(16) float T[COS_D][2 * COS_D];
    __shared__ float sK[COS_D];
    __shared__ __align__(16) float sV[2 * COS_D];
    const int h = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * 64;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bkv = p.kv_shared ? 0 : b;
    const float* mom = scratch + (static_cast<size_t>(bkv) * p.H + h) * COS_MOM;
    for (int e = threadIdx.x; e < COS_D * 2 * COS_D / 4; e += 256)
        reinterpret_cast<float4*>(&T[0][0])[e] = __ldg(reinterpret_cast<const float4*>(mom) + e);
    if (threadIdx.x < COS_D) sK[threadIdx.x] = __ldg(mom + COS_D * 2 * COS_D + threadIdx.x);
    if (threadIdx.x < 2 * COS_D) sV[threadIdx.x] = __ldg(mom + COS_D * 2 * COS_D + COS_D + threadIdx.x);
    __syncthreads();
    for (int r = warp; r < 64; r += 8) {
        const int n = q0 + r;
        if (n >= p.Nc) break;
        const float* qrow = p.q + (static_cast<size_t>(b) * p.Nc + n) * p.ldq + h * COS_D;
        float qa = __ldg(qrow + lane), qb = __ldg(qrow + lane + 32);
        float nn = fmaf(qa, qa, qb * qb);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
        const float inv = 1.f / sqrtf(nn);
        qa *= inv; qb *= inv;
        float den = fmaf(qa, sK[lane], qb * sK[lane + 32]);
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
        den += static_cast<float>(p.Ns);
        float acc[4] = {sV[lane * 4], sV[lane * 4 + 1], sV[lane * 4 + 2], sV[lane * 4 + 3]};
#pragma unroll 8
        for (int d = 0; d < COS_D; ++d) {
            const float qd = __shfl_sync(0xffffffffu, d < 32 ? qa : qb, d & 31);
            const float4 t = *reinterpret_cast<const float4*>(&T[d][lane * 4]);
            acc[0] = fmaf(qd, t.x, acc[0]); acc[1] = fmaf(qd, t.y, acc[1]);
            acc[2] = fmaf(qd, t.z, acc[2]); acc[3] = fmaf(qd, t.w, acc[3]);
        }
        const float rden = 1.f / den;
        // lanes 0..15 hold M of channels lane*4.., lanes 16..31 hold E of channels (lane-16)*4..
        float e4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) e4[j] = __shfl_down_sync(0xffffffffu, acc[j], 16);
        if (lane < 16) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ch = h * COS_D + lane * 4 + j;
                const float m = acc[j] * rden, e = e4[j] * rden;
                const float sd = sqrtf(fmaxf(e - m * m, 1e-6f));
                const size_t sidx = static_cast<size_t>(b) * p.H * COS_D + ch;
                const float xn = (__ldg(p.x + (static_cast<size_t>(b) * p.Nc + n) * p.ldx + ch) - __ldg(p.x_mean + sidx)) *
                                 __ldg(p.x_rstd + sidx);
                const float mu = p.mu_v ? __ldg(p.mu_v + static_cast<size_t>(bkv) * p.H * COS_D + ch) : 0.f;
                p.out[(static_cast<size_t>(b) * p.Nc + n) * p.ldo + ch] = fmaf(sd, xn, m + mu);
            }
        }
    }
}

int launch_attn_f32(const mhada_attn_args& a, cudaStream_t s) {
    AttnF32Params p;
    p.q = static_cast<const float*>(a.q); p.k = static_cast<const float*>(a.k);
    p.v = static_cast<const float*>(a.v); p.x = static_cast<const float*>(a.x);
    p.out = static_cast<float*>(a.out);
    p.x_mean = a.x_mean; p.x_rstd = a.x_rstd; p.mu_v = a.mu_v;
    p.q_mean = a.q_mean; p.q_rstd = a.q_rstd; p.k_mean = a.k_mean; p.k_rstd = a.k_rstd;
    p.H = a.H; p.Nc = a.Nc; p.Ns = a.Ns; p.dqk = a.dqk; p.dv = a.dv;
    p.ldq = a.ldq; p.ldk = a.ldk; p.ldv = a.ldv; p.ldx = a.ldx; p.ldo = a.ldo;
    p.kv_shared = (a.kv_batch == 1 && a.B > 1) ? 1 : 0;
    const int Bkv = p.kv_shared ? 1 : a.B;
    // (not below 128 keys: nothing to gain there, and with a handful of keys Var = E - M^2 can vanish exactly -- one key:
    // a = 1 -- which the N x N kernel reproduces bit for bit while the moments carry fp32 rounding noise into sqrt())
    if (a.activation == MHADA_ACT_COSINE && a.Ns >= 128 && a.dqk == COS_D && a.dv == COS_D && !a.q_mean && !a.k_mean && a.scratch &&
        a.scratch_bytes >= attn_cosine_scratch_bytes(Bkv, a.H) && (reinterpret_cast<uintptr_t>(a.scratch) & 15) == 0) {
        // closed form: moments of every (image, head) over the keys, then one pass over the queries
        float* scratch = static_cast<float*>(a.scratch);
        cosine_moments_kernel<<<dim3(a.H, Bkv), 256, 0, s>>>(p, scratch);
        count_launch();
        cosine_apply_kernel<<<dim3((a.Nc + 63) / 64, a.H, a.B), 256, 0, s>>>(p, scratch);
        count_launch();
        return check_cuda(cudaGetLastError(), "attn_cosine launch");
    }
    dim3 grid((a.Nc + TILE - 1) / TILE, (a.dv + TILE - 1) / TILE, a.B * a.H);
    if (a.activation == MHADA_ACT_COSINE)
        attn_f32_kernel<true><<<grid, 256, 0, s>>>(p);
    else
        attn_f32_kernel<false><<<grid, 256, 0, s>>>(p);
    count_launch();
    return check_cuda(cudaGetLastError(), "attn_f32 launch");
}

}  // namespace mh

// Per-head 1x1 projections on the tensor cores (bf16 path, head_dim 64).
//
// Replaces f_list / g_list / h_list applied to IN(fc), IN(fs), fs at
// MHAdaSTr/network/adaDecoder.py:173, :178, :182.  The instance norm is not applied to the data:
// it is folded into per-image weights by fold_kernel,
//     Q = log2e * (Wf diag(r_c)) x + log2e * (bf - Wf (mu_c * r_c))        (log2e: the attention uses exp2)
//     K =         (Wg diag(r_s)) y +         (bg - Wg (mu_s * r_s))
//     V~ = Wh (y - mu_s)  = V - mu_v,   mu_v = Wh mu_s + bh                 (centred values)
// with the biases evaluated from the bf16-ROUNDED weights so the centring cancels exactly.
// proj_tc_kernel then does one [128 tokens x 64] . [64 x 64]^T tcgen05 MMA per (token tile, head)
// and its epilogue writes Q / K (bf16) or V' = [V~ | V~^2] (squared in fp32, then rounded).
//
// HBM-bound: algorithmic bytes per layer = 2*(B*Nc*C + B*Ns*C) read + 2*(B*Nc*C + 3*B*Ns*C) written.
#include "common.h"
#include "ptx.cuh"

namespace mh {

constexpr int PRJ_BM = 128, PRJ_D = 64, PRJ_THREADS = 192;
constexpr float kLog2e = 1.4426950408889634f;

// workspace layout: [3][B][H][64][64] bf16 folded weights, then [3][B][H][64] f32 folded biases
static size_t fold_w_bytes(int B, int H, int d) { return align_up(static_cast<size_t>(3) * B * H * d * d * 2, 256); }
size_t proj_bf16_workspace(int B, int H, int d) {
    return fold_w_bytes(B, H, d) + align_up(static_cast<size_t>(3) * B * H * d * 4, 256);
}

// grid (H, batch of this part, number of projections), 256 threads = 8 warps: warp w folds output rows w, w+8, ...;
// lanes walk the input channels, so every global read / write is a contiguous row segment.
// which = which0 + blockIdx.z: 0 = f (content side, batch B), 1 = g, 2 = h (style side, batch Bs).  B here is the
// batch STRIDE of the folded-weight workspace ([3][B][H][d][d]), the same for both sides.
__global__ void __launch_bounds__(256) fold_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                                   const float* __restrict__ mean_c, const float* __restrict__ rstd_c,
                                                   const float* __restrict__ mean_s, const float* __restrict__ rstd_s,
                                                   int which0, int B, int H, int d, __nv_bfloat16* __restrict__ wf,
                                                   float* __restrict__ bf, float* __restrict__ mu_v) {
    const int h = blockIdx.x, b = blockIdx.y, which = which0 + blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int C = H * d;
    const float* mu = (which == 0 ? mean_c : mean_s) + static_cast<size_t>(b) * C + h * d;
    const float* rs = (which == 0 ? rstd_c : rstd_s) + static_cast<size_t>(b) * C + h * d;
    const float gain = which == 0 ? kLog2e : 1.f;
    for (int o = warp; o < d; o += 8) {
        const float* wr = w + ((static_cast<size_t>(which) * H + h) * d + o) * d;
        __nv_bfloat16* dst = wf + (((static_cast<size_t>(which) * B + b) * H + h) * d + o) * d;
        float acc = 0.f;
        for (int i = lane; i < d; i += 32) {
            float wv = wr[i] * gain;
            if (which != 2) wv *= rs[i];
            const __nv_bfloat16 wb = __float2bfloat16_rn(wv);
            dst[i] = wb;
            acc = fmaf(__bfloat162float(wb), mu[i], acc);     // bias from the ROUNDED weights: centring cancels exactly
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) {
            const float bo = bias[(static_cast<size_t>(which) * H + h) * d + o];
            const size_t bi = ((static_cast<size_t>(which) * B + b) * H + h) * d + o;
            if (which == 2) {
                bf[bi] = -acc;                                             // V~ = Wh y - Wh mu_s
                mu_v[static_cast<size_t>(b) * C + h * d + o] = acc + bo;   // what the epilogue adds back
            } else {
                bf[bi] = bo * gain - acc;
            }
        }
    }
}

// NOUT = 1: Q from fc.   NOUT = 2: K and V' from the same fs tile.
// grid (ceil(N/128), H, B)
template <int NOUT>
__global__ void __launch_bounds__(PRJ_THREADS)
proj_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
               const float* __restrict__ bfold, int which0, int B, int H, int N,
               __nv_bfloat16* __restrict__ out0, __nv_bfloat16* __restrict__ out1) {
    constexpr uint32_t A_BYTES = PRJ_BM * PRJ_D * 2, W_BYTES = PRJ_D * PRJ_D * 2;
    constexpr uint32_t TM_COLS = NOUT * 64;
    __shared__ __align__(1024) uint8_t smem[A_BYTES + NOUT * W_BYTES];
    __shared__ uint64_t full, accum;
    __shared__ uint32_t tmem_slot;
    __shared__ float bias_s[NOUT][PRJ_D];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * PRJ_BM, h = blockIdx.y, b = blockIdx.z;
    const int C = H * PRJ_D;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW);
    }
    if (warp == 1) {
        if (lane == 0) {
            mbar_init(&full, 1);
            mbar_init(&accum, 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(&tmem_slot, TM_COLS);
        tmem_relinquish();
    }
    if (warp >= 2) {
        for (int i = threadIdx.x - 64; i < NOUT * PRJ_D; i += 128) {
            int t = i / PRJ_D, o = i % PRJ_D;
            bias_s[t][o] = bfold[((static_cast<size_t>(which0 + t) * B + b) * H + h) * PRJ_D + o];
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            mbar_arrive_expect_tx(&full, A_BYTES + NOUT * W_BYTES);
            tma_load_3d(smem, &tmX, &full, h * PRJ_D, n0, b);
#pragma unroll
            for (int t = 0; t < NOUT; ++t)
                tma_load_2d(smem + A_BYTES + t * W_BYTES, &tmW, &full, 0,
                            (((which0 + t) * B + b) * H + h) * PRJ_D);
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(PRJ_BM, PRJ_D, 0, 0);
            mbar_wait(&full, 0);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem);
            const uint64_t da = make_smem_desc(a_addr, 16, 1024);
#pragma unroll
            for (int t = 0; t < NOUT; ++t) {
                const uint64_t db = make_smem_desc(a_addr + A_BYTES + t * W_BYTES, 16, 1024);
#pragma unroll
                for (int k = 0; k < PRJ_D / 16; ++k)
                    umma_ss(tmem + t * 64, desc_advance(da, k * 32), desc_advance(db, k * 32), idesc, k != 0);
            }
            umma_commit(&accum);
        }
    } else {
        const int quarter = warp & 3;
        const int n = n0 + quarter * 32 + lane;
        mbar_wait(&accum, 0);
        tc_fence_after();
#pragma unroll
        for (int t = 0; t < NOUT; ++t) {
            const bool is_v = (NOUT == 2 && t == 1);
#pragma unroll
            for (int c = 0; c < PRJ_D; c += 32) {
                uint32_t r[32];
                tmem_ld_x32(tmem_addr(tmem, quarter * 32, t * 64 + c), r);
                tmem_wait_ld();
                if (n < N) {
                    uint32_t lo[16], sq[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float v0 = __uint_as_float(r[2 * i]) + bias_s[t][c + 2 * i];
                        float v1 = __uint_as_float(r[2 * i + 1]) + bias_s[t][c + 2 * i + 1];
                        lo[i] = pack_bf16x2(v0, v1);
                        sq[i] = pack_bf16x2(v0 * v0, v1 * v1);
                    }
                    if (!is_v) {
                        __nv_bfloat16* o = (t == 0 ? out0 : out1) + (static_cast<size_t>(b) * N + n) * C + h * PRJ_D + c;
                        st_global_256(o, lo);
                        st_global_256(o + 16, lo + 8);
                    } else {
                        __nv_bfloat16* o = out1 + (static_cast<size_t>(b) * N + n) * (2 * C) + h * (2 * PRJ_D) + c;
                        st_global_256(o, lo);
                        st_global_256(o + 16, lo + 8);
                        st_global_256(o + PRJ_D, sq);
                        st_global_256(o + PRJ_D + 16, sq + 8);
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, TM_COLS);
}

static int make_x_map(CUtensorMap* tm, const void* x, int B, int N, int C) {
    uint64_t dims[3] = {static_cast<uint64_t>(C), static_cast<uint64_t>(N), static_cast<uint64_t>(B)};
    uint64_t str[2] = {static_cast<uint64_t>(C) * 2, static_cast<uint64_t>(N) * C * 2};
    uint32_t box[3] = {PRJ_D, PRJ_BM, 1};
    return make_tmap_bf16(tm, x, 3, dims, str, box);
}

int launch_proj_bf16(int parts, const void* fc, const void* fs, const float* mean_c, const float* rstd_c,
                     const float* mean_s, const float* rstd_s, const float* w, const float* bias, int B, int Bs, int Nc,
                     int Ns, int H, int d, void* q, void* k, void* v, float* mu_v, void* ws, cudaStream_t s) {
    // parts: 1 = Q from fc (batch B), 2 = K and V' from fs (batch Bs).  Bw = batch stride of the folded weights.
    const int Bw = B > Bs ? B : Bs;
    __nv_bfloat16* wf = static_cast<__nv_bfloat16*>(ws);
    float* bf = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + fold_w_bytes(Bw, H, d));
    const int C = H * d;
    CUtensorMap tmW;
    uint64_t dimsW[2] = {static_cast<uint64_t>(d), static_cast<uint64_t>(3) * Bw * H * d};
    uint64_t strW[1] = {static_cast<uint64_t>(d) * 2};
    uint32_t boxW[2] = {PRJ_D, PRJ_D};
    if (int e = make_tmap_bf16(&tmW, wf, 2, dimsW, strW, boxW)) return e;
    if (parts & 1) {
        fold_kernel<<<dim3(H, B, 1), 256, 0, s>>>(w, bias, mean_c, rstd_c, mean_s, rstd_s, 0, Bw, H, d, wf, bf, mu_v);
        count_launch();
        CUtensorMap tmC;
        if (int e = make_x_map(&tmC, fc, B, Nc, C)) return e;
        proj_tc_kernel<1><<<dim3((Nc + PRJ_BM - 1) / PRJ_BM, H, B), PRJ_THREADS, 0, s>>>(
            tmC, tmW, bf, 0, Bw, H, Nc, static_cast<__nv_bfloat16*>(q), nullptr);
        count_launch();
    }
    if (parts & 2) {
        fold_kernel<<<dim3(H, Bs, 2), 256, 0, s>>>(w, bias, mean_c, rstd_c, mean_s, rstd_s, 1, Bw, H, d, wf, bf, mu_v);
        count_launch();
        CUtensorMap tmS;
        if (int e = make_x_map(&tmS, fs, Bs, Ns, C)) return e;
        proj_tc_kernel<2><<<dim3((Ns + PRJ_BM - 1) / PRJ_BM, H, Bs), PRJ_THREADS, 0, s>>>(
            tmS, tmW, bf, 1, Bw, H, Ns, static_cast<__nv_bfloat16*>(k), static_cast<__nv_bfloat16*>(v));
        count_launch();
    }
    return check_cuda(cudaGetLastError(), "proj_tc launch");
}

}  // namespace mh

// ViT encoder around the token GEMM (gemm_tc.cu): the small HBM-bound kernels and the per-layer launch sequence.
//
// Replaces VisionTransformer.forward, MHAdaSTr/network/vit.py:148-169:
//   patch_im2col_kernel   the gather half of PatchEmbedding.conv_proj (Conv2d k = stride = patch, :105-117): image
//                         pixels -> bf16 rows [B*N, 3*patch*patch] in Conv2d weight order, so the convolution is a GEMM
//   layernorm_kernel      nn.LayerNorm(hidden, eps=1e-6) (:54-55, used :58, :62): f32 residual stream -> bf16 GEMM operand
//   batch_attn_kernel     nn.MultiheadAttention built without batch_first and fed (B, N, D) (:48, :59): the sequence
//                         axis is the BATCH -- for every token position and head, softmax(q k^T / sqrt(hd)) over the B
//                         images.  Reproduced as is: the reference's features of image k depend on the
//                         other images of its batch.
// All three are bandwidth-bound passes between the GEMMs; the residual stream stays f32 (the per-channel DC of the
// patch embedding is ~20x the spatial signal the MHAda instance norm later extracts; bf16 would round it away).
#include "common.h"
#include "ptx.cuh"

namespace mh {

// ------------------------------------------------------------------------------------------------ patch gather
// thread = (b, c, image row, patch column): reads `P` contiguous pixels, writes `P` contiguous bf16
template <typename T, int P>
__global__ void __launch_bounds__(256) patch_im2col_kernel(const T* __restrict__ img, __nv_bfloat16* __restrict__ a0,
                                                           int B, int Himg, int Wimg) {
    const int wp = Wimg / P, hp = Himg / P;
    const size_t total = static_cast<size_t>(B) * 3 * Himg * wp;
    const size_t t = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (t >= total) return;
    const int xp = static_cast<int>(t % wp);
    const int row = static_cast<int>((t / wp) % Himg);
    const int c = static_cast<int>((t / (static_cast<size_t>(wp) * Himg)) % 3);
    const int b = static_cast<int>(t / (static_cast<size_t>(wp) * Himg * 3));
    const T* src = img + ((static_cast<size_t>(b) * 3 + c) * Himg + row) * Wimg + static_cast<size_t>(xp) * P;
    const int y = row / P, dy = row % P;
    const size_t n = static_cast<size_t>(b) * hp * wp + static_cast<size_t>(y) * wp + xp;
    __nv_bfloat16* dst = a0 + n * (3 * P * P) + c * P * P + dy * P;
    float v[P];
    if constexpr (P == 8 && sizeof(T) == 4) {
        const float4 lo = __ldg(reinterpret_cast<const float4*>(src)), hi = __ldg(reinterpret_cast<const float4*>(src) + 1);
        v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
    } else if constexpr (P == 8 && sizeof(T) == 1) {
        const uint2 w = __ldg(reinterpret_cast<const uint2*>(src));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[i] = static_cast<float>((w.x >> (8 * i)) & 0xffu);
            v[4 + i] = static_cast<float>((w.y >> (8 * i)) & 0xffu);
        }
    } else {
#pragma unroll
        for (int i = 0; i < P; ++i) v[i] = static_cast<float>(src[i]);
    }
    if constexpr (P == 8) {
        uint4 o;
        o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(dst) = o;
    } else {
#pragma unroll
        for (int i = 0; i < P; ++i) dst[i] = __float2bfloat16_rn(v[i]);
    }
}

int launch_patch_im2col(int img_dtype, const void* img, int B, int Himg, int Wimg, int patch, void* a0, cudaStream_t s) {
    if (patch != 8 && patch != 16) {
        set_error("patch_im2col: patch sizes 8 and 16 are implemented (3*patch^2 must be a multiple of 64), got %d", patch);
        return MHADA_ERR_UNSUPPORTED;
    }
    const size_t total = static_cast<size_t>(B) * 3 * Himg * (Wimg / patch);
    const unsigned grid = static_cast<unsigned>((total + 255) / 256);
    __nv_bfloat16* out = static_cast<__nv_bfloat16*>(a0);
    if (img_dtype == MHADA_F32) {
        if (patch == 8) patch_im2col_kernel<float, 8><<<grid, 256, 0, s>>>(static_cast<const float*>(img), out, B, Himg, Wimg);
        else patch_im2col_kernel<float, 16><<<grid, 256, 0, s>>>(static_cast<const float*>(img), out, B, Himg, Wimg);
    } else {
        if (patch == 8) patch_im2col_kernel<uint8_t, 8><<<grid, 256, 0, s>>>(static_cast<const uint8_t*>(img), out, B, Himg, Wimg);
        else patch_im2col_kernel<uint8_t, 16><<<grid, 256, 0, s>>>(static_cast<const uint8_t*>(img), out, B, Himg, Wimg);
    }
    count_launch();
    return check_cuda(cudaGetLastError(), "patch_im2col launch");
}

// ------------------------------------------------------------------------------------------------ LayerNorm
// warp = token row; lane owns 4 consecutive channels of every 128-channel group (128-bit loads, 64-bit stores)
constexpr int LN_MAXV = 8;      // C <= 1024
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, int M, int C,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float eps, __nv_bfloat16* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int nv = C >> 7;
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * C);
    float4 v[LN_MAXV];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < LN_MAXV; ++j)
        if (j < nv) {
            v[j] = __ldg(xr + j * 32 + lane);
            sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / static_cast<float>(C);
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < LN_MAXV; ++j)
        if (j < nv) {
            const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
            sq += (a * a + b * b) + (c * c + d * d);
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / static_cast<float>(C) + eps);      // biased variance, eps inside the root
    uint2* yr = reinterpret_cast<uint2*>(y + static_cast<size_t>(row) * C);
#pragma unroll
    for (int j = 0; j < LN_MAXV; ++j)
        if (j < nv) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + j * 32 + lane);
            const float4 be = __ldg(reinterpret_cast<const float4*>(beta) + j * 32 + lane);
            uint2 o;
            o.x = pack_bf16x2(fmaf((v[j].x - mean) * rstd, g.x, be.x), fmaf((v[j].y - mean) * rstd, g.y, be.y));
            o.y = pack_bf16x2(fmaf((v[j].z - mean) * rstd, g.z, be.z), fmaf((v[j].w - mean) * rstd, g.w, be.w));
            yr[j * 32 + lane] = o;
        }
}

int launch_layernorm(const float* x, int M, int C, const float* gamma, const float* beta, float eps, void* y_bf16,
                     cudaStream_t s) {
    if (C % 128 != 0 || C > 128 * LN_MAXV) {
        set_error("layernorm: C must be a multiple of 128 and at most %d, got %d", 128 * LN_MAXV, C);
        return MHADA_ERR_UNSUPPORTED;
    }
    layernorm_kernel<<<(M + 7) / 8, 256, 0, s>>>(x, M, C, gamma, beta, eps, static_cast<__nv_bfloat16*>(y_bf16));
    count_launch();
    return check_cuda(cudaGetLastError(), "layernorm launch");
}

// ------------------------------------------------------------------------------------------------ batch-axis attention
// warp = (token position n, head); the "sequence" is the B images.  q, k, v of the B images are staged in shared
// memory.  Work is spread over all 32 lanes: the B*B logits as (i, j) pairs, the B*64 outputs as (i, 8-channel group)
// items.
constexpr int BA_HD = 64;
constexpr int BA_WARPS = 8;
constexpr int BA_PITCH = BA_HD + 8;                    // bf16 elements per staged row (144 B: 16-byte reads of 8 rows hit 32 banks)
__host__ __device__ constexpr int ba_warp_bytes(int B) { return 3 * B * BA_PITCH * 2 + ((B * (B + 1) * 4 + 15) & ~15); }
__device__ __forceinline__ void ba_unpack8(const uint4& w, float (&f)[8]) {
    f[0] = bf16_lo(w.x); f[1] = bf16_hi(w.x); f[2] = bf16_lo(w.y); f[3] = bf16_hi(w.y);
    f[4] = bf16_lo(w.z); f[5] = bf16_hi(w.z); f[6] = bf16_lo(w.w); f[7] = bf16_hi(w.w);
}
__global__ void __launch_bounds__(BA_WARPS * 32) batch_attn_kernel(const __nv_bfloat16* __restrict__ qkv, int B, int N,
                                                                   int heads, __nv_bfloat16* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t ba_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long item = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + warp;     // n * heads + h
    if (item >= static_cast<long long>(N) * heads) return;
    const int n = static_cast<int>(item / heads), h = static_cast<int>(item % heads);
    const int D = heads * BA_HD;
    // q, k, v stay bf16 in shared memory (the shared-memory pipe, not the FMAs, bounds this kernel: f32 staging cost
    // twice the wavefronts); they are widened in registers
    __nv_bfloat16* q = reinterpret_cast<__nv_bfloat16*>(ba_smem + static_cast<size_t>(warp) * ba_warp_bytes(B));
    __nv_bfloat16* k = q + B * BA_PITCH;
    __nv_bfloat16* v = k + B * BA_PITCH;
    float* sc = reinterpret_cast<float*>(v + B * BA_PITCH);      // [B][B + 1] logits, then probabilities
    // stage: three coalesced 128-byte rows per image, four images in flight
    for (int b0 = 0; b0 < B; b0 += 4) {
        uint32_t w[4][3];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (b0 + u < B) {
                const uint32_t* base = reinterpret_cast<const uint32_t*>(qkv + (static_cast<size_t>(b0 + u) * N + n) * 3 * D + h * BA_HD);
                w[u][0] = __ldg(base + lane); w[u][1] = __ldg(base + D / 2 + lane); w[u][2] = __ldg(base + D + lane);
            }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (b0 + u < B) {
                const int o = (b0 + u) * BA_PITCH + 2 * lane;
                *reinterpret_cast<uint32_t*>(q + o) = w[u][0];
                *reinterpret_cast<uint32_t*>(k + o) = w[u][1];
                *reinterpret_cast<uint32_t*>(v + o) = w[u][2];
            }
    }
    __syncwarp();
    // logits: one (i, j) pair per lane and round
    for (int pr = lane; pr < B * B; pr += 32) {
        const int i = pr / B, j = pr % B;
        const uint4* qi = reinterpret_cast<const uint4*>(q + i * BA_PITCH);
        const uint4* kj = reinterpret_cast<const uint4*>(k + j * BA_PITCH);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int d = 0; d < BA_HD / 8; ++d) {
            float x[8], y[8];
            ba_unpack8(qi[d], x);
            ba_unpack8(kj[d], y);
            a0 = fmaf(x[0], y[0], a0); a1 = fmaf(x[1], y[1], a1); a2 = fmaf(x[2], y[2], a2); a3 = fmaf(x[3], y[3], a3);
            a0 = fmaf(x[4], y[4], a0); a1 = fmaf(x[5], y[5], a1); a2 = fmaf(x[6], y[6], a2); a3 = fmaf(x[7], y[7], a3);
        }
        sc[i * (B + 1) + j] = ((a0 + a1) + (a2 + a3)) * 0.125f;        // 1 / sqrt(64)
    }
    __syncwarp();
    if (lane < B) {                                    // softmax of row `lane` over the B images
        float* row = sc + lane * (B + 1);
        float mx = row[0];
        for (int j = 1; j < B; ++j) mx = fmaxf(mx, row[j]);
        float sum = 0.f;
        for (int j = 0; j < B; ++j) {
            const float e = __expf(row[j] - mx);
            row[j] = e;
            sum += e;
        }
        const float inv = 1.f / sum;
        for (int j = 0; j < B; ++j) row[j] *= inv;
    }
    __syncwarp();
    // outputs: (image i, 8-channel group) items; 8 consecutive lanes write one 128-byte row
    for (int t = lane; t < B * (BA_HD / 8); t += 32) {
        const int i = t / (BA_HD / 8), d8 = t % (BA_HD / 8);
        const float* pi = sc + i * (B + 1);
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < B; ++j) {
            const float pw = pi[j];
            float y[8];
            ba_unpack8(reinterpret_cast<const uint4*>(v + j * BA_PITCH)[d8], y);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = fmaf(pw, y[e], acc[e]);
        }
        uint4 o;
        o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
        o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
        reinterpret_cast<uint4*>(out + (static_cast<size_t>(i) * N + n) * D + h * BA_HD)[d8] = o;
    }
}

// B <= 8 (the batch sizes the reference uses: 8 in train_image.py, 1 per frame in the inference scripts): the whole
// problem of a (token position, head) -- S = Q K^T (8 x 8 x 64), softmax over the 8 images, O = P V (8 x 64 x 8) -- is
// twelve warp-level mma.sync.m16n8k16 instructions on fragments loaded straight from global memory (rows = images,
// lane (g, t) reads row g), no shared memory.  The V tile of every 8-channel group is transposed in registers with
// movmatrix.  (ncu r02 of the SIMT kernel above: issue slots 82 % busy, ~1000 instructions per warp; this one ~110.)
// The legacy warp MMA is the right tool here: one tcgen05 tile (M = 128) would need 16 token positions x 8 images
// gathered into one operand, and the kernel is bandwidth-bound (2.1 MFLOP per 4 KB).
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                               uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t x) {
    uint32_t y;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
__global__ void __launch_bounds__(256) batch_attn_mma_kernel(const __nv_bfloat16* __restrict__ qkv, int B, int N, int heads,
                                                             __nv_bfloat16* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long item = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);     // n * heads + h
    if (item >= static_cast<long long>(N) * heads) return;
    const int n = static_cast<int>(item / heads), h = static_cast<int>(item % heads);
    const int D = heads * BA_HD;
    const int g = lane >> 2, t = lane & 3;
    const bool row_ok = g < B;
    const uint32_t* row = reinterpret_cast<const uint32_t*>(qkv + (static_cast<size_t>(row_ok ? g : 0) * N + n) * 3 * D + h * BA_HD);
    uint32_t qa[8], kb[8], vv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {                      // word i*4 + t of the 32-word (64 x bf16) row of image g
        qa[i] = row_ok ? __ldg(row + i * 4 + t) : 0u;
        kb[i] = row_ok ? __ldg(row + D / 2 + i * 4 + t) : 0u;
        vv[i] = row_ok ? __ldg(row + D + i * 4 + t) : 0u;
    }
    // S[g][j] = sum_d Q[g][d] K[j][d]: A rows 0..7 = images (rows 8..15 zero), B columns = images
    float sc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) mma_bf16_16816(sc, qa[2 * kk], 0u, qa[2 * kk + 1], 0u, kb[2 * kk], kb[2 * kk + 1]);
    // softmax over the images j = 2t, 2t + 1 of the four lanes of a row group
    float s0 = (2 * t < B) ? sc[0] * 0.125f : -INFINITY, s1 = (2 * t + 1 < B) ? sc[1] * 0.125f : -INFINITY;
    float mx = fmaxf(s0, s1);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const float e0 = __expf(s0 - mx), e1 = __expf(s1 - mx);
    float sum = e0 + e1;
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const float inv = 1.f / sum;
    const uint32_t pa = pack_bf16x2(e0 * inv, e1 * inv);            // A fragment of P: row g, k = 2t, 2t + 1 (k >= 8 zero)
    uint32_t* orow = reinterpret_cast<uint32_t*>(out + (static_cast<size_t>(row_ok ? g : 0) * N + n) * D + h * BA_HD);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {                   // 8 value channels per step: O[g][nt*8 + ..] = sum_j P[g][j] V[j][..]
        const uint32_t vb = movmatrix_trans(vv[nt]);   // lane (g, t): V[2t][nt*8 + g], V[2t+1][nt*8 + g]
        float o[4] = {0.f, 0.f, 0.f, 0.f};
        mma_bf16_16816(o, pa, 0u, 0u, 0u, vb, 0u);
        if (row_ok) orow[nt * 4 + t] = pack_bf16x2(o[0], o[1]);
    }
}

// Backward of the batch-axis attention for the training step (train_image.py:103-144 through vit.py:59), B <= 8:
// one warp per (token position, head), lane = two of the 64 channels.  The 8 x 8 logits S = Q K^T / 8 and dP = dO V^T are
// warp-wide dot products (butterfly sums: every lane ends up with all of them), the rest is lane-local:
//     P = softmax_rows(S),  dV = P^T dO,  dS = P (dP - rowsum(P dP)) / 8,  dQ = dS K,  dK = dS^T Q.
// ~2000 instructions per warp, bandwidth-bound like the forward (reads qkv + dO, writes dqkv: 14 bytes per qkv element).
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__global__ void __launch_bounds__(256) batch_attn_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ dout,
                                                             int B, int N, int heads, __nv_bfloat16* __restrict__ dqkv) {
    const int lane = threadIdx.x & 31;
    const long long item = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);     // n * heads + h
    if (item >= static_cast<long long>(N) * heads) return;
    const int n = static_cast<int>(item / heads), h = static_cast<int>(item % heads);
    const int D = heads * BA_HD;
    float q[8][2], k[8][2], v[8][2], g[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        q[i][0] = q[i][1] = k[i][0] = k[i][1] = v[i][0] = v[i][1] = g[i][0] = g[i][1] = 0.f;
        if (i < B) {
            const uint32_t* row = reinterpret_cast<const uint32_t*>(qkv + (static_cast<size_t>(i) * N + n) * 3 * D + h * BA_HD) + lane;
            const uint32_t wq = __ldg(row), wk = __ldg(row + D / 2), wv = __ldg(row + D);
            const uint32_t wg = __ldg(reinterpret_cast<const uint32_t*>(dout + (static_cast<size_t>(i) * N + n) * D + h * BA_HD) + lane);
            q[i][0] = bf16_lo(wq); q[i][1] = bf16_hi(wq); k[i][0] = bf16_lo(wk); k[i][1] = bf16_hi(wk);
            v[i][0] = bf16_lo(wv); v[i][1] = bf16_hi(wv); g[i][0] = bf16_lo(wg); g[i][1] = bf16_hi(wg);
        }
    }
    float p[8][8], dp[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            p[i][j] = dp[i][j] = 0.f;
            if (i < B && j < B) {
                p[i][j] = warp_sum(fmaf(q[i][0], k[j][0], q[i][1] * k[j][1])) * 0.125f;
                dp[i][j] = warp_sum(fmaf(g[i][0], v[j][0], g[i][1] * v[j][1]));
            }
        }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (i >= B) continue;
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < B) mx = fmaxf(mx, p[i][j]);
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            p[i][j] = j < B ? __expf(p[i][j] - mx) : 0.f;
            sum += p[i][j];
        }
        const float inv = 1.f / sum;
        float dsum = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            p[i][j] *= inv;
            dsum = fmaf(p[i][j], dp[i][j], dsum);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) dp[i][j] = p[i][j] * (dp[i][j] - dsum) * 0.125f;      // dS (with the 1 / sqrt(hd) of S)
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        if (r >= B) continue;
        float dq0 = 0.f, dq1 = 0.f, dk0 = 0.f, dk1 = 0.f, dv0 = 0.f, dv1 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            dq0 = fmaf(dp[r][j], k[j][0], dq0); dq1 = fmaf(dp[r][j], k[j][1], dq1);      // dQ[r] = sum_j dS[r][j] K[j]
            dk0 = fmaf(dp[j][r], q[j][0], dk0); dk1 = fmaf(dp[j][r], q[j][1], dk1);      // dK[r] = sum_i dS[i][r] Q[i]
            dv0 = fmaf(p[j][r], g[j][0], dv0); dv1 = fmaf(p[j][r], g[j][1], dv1);        // dV[r] = sum_i P[i][r] dO[i]
        }
        uint32_t* orow = reinterpret_cast<uint32_t*>(dqkv + (static_cast<size_t>(r) * N + n) * 3 * D + h * BA_HD) + lane;
        orow[0] = pack_bf16x2(dq0, dq1);
        orow[D / 2] = pack_bf16x2(dk0, dk1);
        orow[D] = pack_bf16x2(dv0, dv1);
    }
}

int launch_batch_attn_bwd(const void* qkv, const void* dout, int B, int N, int heads, int hd, void* dqkv, cudaStream_t s) {
    if (hd != BA_HD || B < 1 || B > 8) {
        set_error("batch_attn_bwd: head_dim 64 and batch 1..8 are implemented, got head_dim %d, batch %d", hd, B);
        return MHADA_ERR_UNSUPPORTED;
    }
    const long long items = static_cast<long long>(N) * heads;
    batch_attn_bwd_kernel<<<static_cast<unsigned>((items + 7) / 8), 256, 0, s>>>(
        static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(dout), B, N, heads, static_cast<__nv_bfloat16*>(dqkv));
    count_launch();
    return check_cuda(cudaGetLastError(), "batch_attn_bwd launch");
}

int launch_batch_attn(const void* qkv, int B, int N, int heads, int hd, void* out, cudaStream_t s) {
    if (hd != BA_HD || B < 1 || B > 32) {
        set_error("batch_attn: head_dim 64 and batch 1..32 are implemented, got head_dim %d, batch %d", hd, B);
        return MHADA_ERR_UNSUPPORTED;
    }
    if (B <= 8) {
        const long long items = static_cast<long long>(N) * heads;
        batch_attn_mma_kernel<<<static_cast<unsigned>((items + 7) / 8), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(qkv), B, N,
                                                                                    heads, static_cast<__nv_bfloat16*>(out));
        count_launch();
        return check_cuda(cudaGetLastError(), "batch_attn_mma launch");
    }
    constexpr size_t kSmemCap = 200 * 1024;
    const size_t per_warp = static_cast<size_t>(ba_warp_bytes(B));
    int warps = static_cast<int>(kSmemCap / per_warp);
    if (warps > BA_WARPS) warps = BA_WARPS;
    const size_t smem = warps * per_warp;
    static DeviceOnce once;
    if (int e = smem_attr_once(once, reinterpret_cast<const void*>(batch_attn_kernel), kSmemCap, "batch_attn smem attr")) return e;
    const long long items = static_cast<long long>(N) * heads;
    batch_attn_kernel<<<static_cast<unsigned>((items + warps - 1) / warps), warps * 32, smem, s>>>(
        static_cast<const __nv_bfloat16*>(qkv), B, N, heads, static_cast<__nv_bfloat16*>(out));
    count_launch();
    return check_cuda(cudaGetLastError(), "batch_attn launch");
}

// ------------------------------------------------------------------------------------------------ the encoder
struct VitWs {
    void *a0, *y, *qkv, *att, *hbuf;
    float *x0, *x1;
    size_t total;
};
static VitWs vit_carve(int B, int N, int D, int F, int K0, uint8_t* base) {
    VitWs w;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void* p = base ? base + off : nullptr;
        off += align_up(bytes, 1024);
        return p;
    };
    const size_t M = static_cast<size_t>(B) * N;
    w.a0 = take(M * K0 * 2);
    w.y = take(M * D * 2);
    w.qkv = take(M * (B == 1 ? D : 3 * D) * 2);
    w.att = take(M * D * 2);
    w.hbuf = take(M * F * 2);
    w.x0 = static_cast<float*>(take(M * D * 4));
    w.x1 = static_cast<float*>(take(M * D * 4));
    w.total = off;
    return w;
}
size_t vit_workspace(int B, int N, int D, int F, int K0) { return vit_carve(B, N, D, F, K0, nullptr).total; }

int vit_forward(const mhada_vit_args& a, cudaStream_t s) {
    const int P = a.patch, hp = a.Himg / P, wp = a.Wimg / P, N = hp * wp, K0 = 3 * P * P, D = a.D, F = a.F;
    const int M = a.B * N;
    VitWs w = vit_carve(a.B, N, D, F, K0, static_cast<uint8_t*>(a.ws));
    // PatchEmbedding (+ PosEmbedding): vit.py:153-158
    if (int e = launch_patch_im2col(a.img_dtype, a.img, a.B, a.Himg, a.Wimg, P, w.a0, s)) return e;
    GemmDesc g{};
    g.a = w.a0; g.lda = K0; g.w = a.w_patch; g.ldw = K0; g.bias = a.b_patch; g.M = M; g.N = D; g.K = K0;
    g.out_f32 = w.x0; g.ldf = D; g.resid = a.pos; g.ldr = D; g.resid_mod = a.pos ? N : 0;
    if (int e = launch_gemm_bf16(g, s)) return e;
    const float* x = w.x0;
    for (int l = 0; l < a.n_layers; ++l) {                                        // EncoderBlock.forward, vit.py:57-64
        const mhada_vit_layer& L = a.layers[l];
        if (int e = launch_layernorm(x, M, D, L.ln1_g, L.ln1_b, 1e-6f, w.y, s)) return e;               // :58
        if (a.B == 1) {
            // sequence length 1: softmax over a single logit is exactly 1, the attention output IS v (:59)
            g = GemmDesc{};
            g.a = w.y; g.lda = D; g.w = static_cast<const __nv_bfloat16*>(L.w_in) + static_cast<size_t>(2 * D) * D; g.ldw = D;
            g.bias = L.b_in + 2 * D; g.M = M; g.N = D; g.K = D; g.out_bf16 = w.att; g.ldo = D;
            if (int e = launch_gemm_bf16(g, s)) return e;
        } else {
            g = GemmDesc{};
            g.a = w.y; g.lda = D; g.w = L.w_in; g.ldw = D; g.bias = L.b_in; g.M = M; g.N = 3 * D; g.K = D;
            g.out_bf16 = w.qkv; g.ldo = 3 * D;
            if (int e = launch_gemm_bf16(g, s)) return e;
            if (int e = launch_batch_attn(w.qkv, a.B, N, a.heads, D / a.heads, w.att, s)) return e;
        }
        g = GemmDesc{};                                                                                 // out_proj + residual :60
        g.a = w.att; g.lda = D; g.w = L.w_out; g.ldw = D; g.bias = L.b_out; g.M = M; g.N = D; g.K = D;
        g.out_f32 = w.x1; g.ldf = D; g.resid = x; g.ldr = D;
        if (int e = launch_gemm_bf16(g, s)) return e;
        if (int e = launch_layernorm(w.x1, M, D, L.ln2_g, L.ln2_b, 1e-6f, w.y, s)) return e;             // :62
        g = GemmDesc{};                                                                                 // mlp[0] + ReLU :63
        g.a = w.y; g.lda = D; g.w = L.w_fc1; g.ldw = D; g.bias = L.b_fc1; g.M = M; g.N = F; g.K = D;
        g.out_bf16 = w.hbuf; g.ldo = F; g.relu = 1;
        if (int e = launch_gemm_bf16(g, s)) return e;
        g = GemmDesc{};                                                                                 // mlp[2] + residual :64
        g.a = w.hbuf; g.lda = F; g.w = L.w_fc2; g.ldw = F; g.bias = L.b_fc2; g.M = M; g.N = D; g.K = F;
        g.out_f32 = a.feat_f32[l]; g.ldf = D; g.resid = w.x1; g.ldr = D;
        g.out_bf16 = a.feat_bf16[l]; g.ldo = D;
        if (int e = launch_gemm_bf16(g, s)) return e;
        x = a.feat_f32[l];
    }
    return 0;
}

}  // namespace mh

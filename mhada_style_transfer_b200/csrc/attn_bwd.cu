// Backward of the MHAda attention stage: FlashAttention-style backward with the value
// operand V' = [V~ | V~^2], for the training step of MHAdaSTr/train_image.py:105-144 (autograd through
// adaDecoder.py:186-198).  bf16 operands, f32 accumulation, head_dim 64; nothing of size Nc x Ns is ever written.
//
// With O' = A V' = [M~ | E~] (centred values), sigma = sqrt(max(E~ - M~^2, 1e-6)), Y = sigma * IN(fcs) + M~ + mu_v
// and g = dL/dY:
//     d(IN(fcs)) = g sigma,   dVar = [Var >= 1e-6] g IN(fcs) / (2 sigma),   dO' = [g - 2 M~ dVar | dVar],
//     delta_i = dO'_i . O'_i,   dA = dO' V'^T,   dS = A (dA - delta),   dQ = dS K,   dK = dS^T Q,   dV' = A^T dO',
//     dV = dV'[:, :d] + 2 V~ dV'[:, d:]
// (the centring needs no correction terms: every mu_v contribution is constant along a row of A and cancels in
// dS, and dV computed from the centred quantities IS the gradient of the uncentred V; derivation in DESIGN.md 4.7).
//
// Three kernels, all warp-level mma.sync.m16n8k16 (bf16) on 64 x 64 tiles staged in shared memory by cp.async:
//   attn_bwd_prep_kernel : per query tile, streaming pass over the keys (online softmax) -> row log-sum-exp L (log2
//                          units), then the element-wise part above in its epilogue: dO' (two bf16 terms, hi + lo),
//                          delta, d(IN(fcs)).
//   attn_bwd_dq_kernel   : per query tile, loop over key tiles:   dQ += dS K     (dS stays in registers)
//   attn_bwd_dkv_kernel  : per key tile, loop over query tiles:   dV' += P^T dO', dK += dS^T Q   (P, dS through smem)
// No atomics: every output element has one owner, results are deterministic.  At the training resolution (256 x 256
// images = 1024 tokens, batch 8, 8 heads) a layer's attention backward is 2 N^2 (192 + 256 + 448) FLOP x 64 = 120 GFLOP.
// tcgen05 would be the next step for large N; at 1024 tokens the whole layer backward is launch / HBM bound.
#include "common.h"
#include "ptx.cuh"

namespace mh {
namespace {

constexpr int BW_T = 64;       // query rows / keys per tile
constexpr int BW_P64 = 72;     // shared-memory pitch (bf16) of a 64-column tile: 144 B rows, conflict-free ldmatrix
constexpr int BW_P128 = 136;   // pitch of a 128-column tile: 272 B rows
constexpr float BW_LN2 = 0.6931471805599453f;

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ---- fragment addresses (lane -> row address of ldmatrix.x4) ----------------------------------------------------------
// A fragment of rows r0.., k columns k0.. of a row-major tile X[row][k]
__device__ __forceinline__ uint32_t addr_a(uint32_t base, int pitch, int r0, int k0, int lane) {
    return base + static_cast<uint32_t>(((r0 + (lane & 15)) * pitch + k0 + (lane >> 4) * 8) * 2);
}
// B fragments of two n-tiles (n0.., n0 + 8..) x k columns k0.. of a K-major tile Y[n][k]   (r0,r1 | r2,r3)
__device__ __forceinline__ uint32_t addr_b(uint32_t base, int pitch, int n0, int k0, int lane) {
    return base + static_cast<uint32_t>(((n0 + (lane & 7) + ((lane >> 4) << 3)) * pitch + k0 + ((lane >> 3) & 1) * 8) * 2);
}
// B fragments of two n-tiles from a tile stored Z[k][n] (ldmatrix .trans)                     (r0,r1 | r2,r3)
__device__ __forceinline__ uint32_t addr_bt(uint32_t base, int pitch, int k0, int n0, int lane) {
    return base + static_cast<uint32_t>(((k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * pitch + n0 + (lane >> 4) * 8) * 2);
}
// A fragment of rows m0.., k columns k0.. from a tile stored transposed W[k][m] (ldmatrix .trans)
__device__ __forceinline__ uint32_t addr_at(uint32_t base, int pitch, int k0, int m0, int lane) {
    return base + static_cast<uint32_t>(((k0 + (lane & 7) + (lane >> 4) * 8) * pitch + m0 + ((lane >> 3) & 1) * 8) * 2);
}

// 64 rows x COLS bf16 from global (row pitch ld) into shared (row pitch PITCH); rows >= valid are zero-filled.
template <int COLS, int PITCH>
__device__ __forceinline__ void load_tile(__nv_bfloat16* s, const __nv_bfloat16* g, int ld, int valid) {
    constexpr int CH = COLS / 8;
    for (int i = threadIdx.x; i < BW_T * CH; i += 128) {
        const int r = i / CH, c = (i % CH) * 8;
        const uint32_t dst = smem_u32(s + r * PITCH + c);
        const bool ok = r < valid;
        const __nv_bfloat16* src = g + (ok ? static_cast<size_t>(r) * ld + c : 0);
        const int bytes = ok ? 16 : 0;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
    }
}
__device__ __forceinline__ void cp_commit_wait() {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

struct BwdParams {
    int B, H, Nc, Ns, C;
    const __nv_bfloat16 *q, *k, *v;     // Q [B,Nc,C] (x log2 e), K [B,Ns,C], V' [B,Ns,2C] (per head [V~ | V~^2])
    const __nv_bfloat16* x;             // fcs [B,Nc,C]
    const float *x_mean, *x_rstd;       // [B,C]
    const float* g;                     // dL/d(cat) [B,Nc,C]
    __nv_bfloat16* d_o;                 // dO' [B,Nc,4C] (per head [dM~ | dE] hi, then [dM~ | dE] lo: two bf16 terms)
    float *lse, *delta;                 // [B,H,Nc]
    float* d_xhat;                      // d(IN(fcs)) [B,Nc,C]
    __nv_bfloat16 *d_q, *d_k, *d_v;     // [B,Nc,C], [B,Ns,C], [B,Ns,C]
};

// S (16 rows x 64 keys per warp) = Qfrag . K^T, K tile K-major in shared memory
__device__ __forceinline__ void logits_tile(float (&s)[8][4], const uint32_t (&qf)[4][4], uint32_t ks, int lane) {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            uint32_t b[4];
            ldsm_x4(b, addr_b(ks, BW_P64, n2 * 16, kk * 16, lane));
            mma16816(s[2 * n2], qf[kk], b[0], b[1]);
            mma16816(s[2 * n2 + 1], qf[kk], b[2], b[3]);
        }
}

// dA (16 rows x 64 keys per warp) = dO'frag . V'^T, V' tile [key][128] in shared memory
template <bool ZERO>
__device__ __forceinline__ void dattn_tile(float (&da)[8][4], const uint32_t (&dof)[8][4], uint32_t vs, int lane) {
    if (ZERO) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) da[nt][0] = da[nt][1] = da[nt][2] = da[nt][3] = 0.f;
    }
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2)
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            uint32_t b[4];
            ldsm_x4(b, addr_b(vs, BW_P128, n2 * 16, kk * 16, lane));
            mma16816(da[2 * n2], dof[kk], b[0], b[1]);
            mma16816(da[2 * n2 + 1], dof[kk], b[2], b[3]);
        }
}

// ---------------------------------------------------------------------------------------------------------------------
// prep: row log-sum-exp, O' = [M~ | E~], then dO', delta, d(IN(fcs))
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attn_bwd_prep_kernel(const BwdParams p) {
    __shared__ __align__(16) __nv_bfloat16 ks_[BW_T * BW_P64];
    __shared__ __align__(16) __nv_bfloat16 vs_[BW_T * BW_P128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int q0 = blockIdx.x * BW_T, h = blockIdx.y, b = blockIdx.z;
    const int C = p.C, Nc = p.Nc, Ns = p.Ns;
    const uint32_t ks = smem_u32(ks_), vs = smem_u32(vs_);

    // Q tile -> A fragments (staged through the K buffer)
    load_tile<64, BW_P64>(ks_, p.q + (static_cast<size_t>(b) * Nc + q0) * C + h * 64, C, min(BW_T, Nc - q0));
    cp_commit_wait();
    __syncthreads();
    uint32_t qf[4][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) ldsm_x4(qf[kk], addr_a(ks, BW_P64, warp * 16, kk * 16, lane));

    float o[16][4];
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    for (int k0 = 0; k0 < Ns; k0 += BW_T) {
        __syncthreads();                                   // everyone is done with the previous tiles (and with Q)
        const int valid = min(BW_T, Ns - k0);
        load_tile<64, BW_P64>(ks_, p.k + (static_cast<size_t>(b) * Ns + k0) * C + h * 64, C, valid);
        load_tile<128, BW_P128>(vs_, p.v + (static_cast<size_t>(b) * Ns + k0) * 2 * C + h * 128, 2 * C, valid);
        cp_commit_wait();
        __syncthreads();
        float s[8][4];
        logits_tile(s, qf, ks, lane);
        if (valid < BW_T) {
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const int c = nt * 8 + 2 * t;
                if (c >= valid) s[nt][0] = s[nt][2] = -INFINITY;
                if (c + 1 >= valid) s[nt][1] = s[nt][3] = -INFINITY;
            }
        }
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
            mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
        }
        mx0 = fmaxf(m0, quad_max(mx0));
        mx1 = fmaxf(m1, quad_max(mx1));
        const float sc0 = ex2_approx(m0 - mx0), sc1 = ex2_approx(m1 - mx1);     // first tile: 2^(-inf) = 0
        m0 = mx0; m1 = mx1;
        l0 *= sc0; l1 *= sc1;
#pragma unroll
        for (int nt = 0; nt < 16; ++nt) {
            o[nt][0] *= sc0; o[nt][1] *= sc0; o[nt][2] *= sc1; o[nt][3] *= sc1;
        }
        uint32_t pa[8][2];                                 // P as bf16 pairs: [nt][0] = row g, [nt][1] = row g + 8
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            pa[nt][0] = pack_bf16x2(ex2_approx(s[nt][0] - m0), ex2_approx(s[nt][1] - m0));
            pa[nt][1] = pack_bf16x2(ex2_approx(s[nt][2] - m1), ex2_approx(s[nt][3] - m1));
            l0 += bf16_lo(pa[nt][0]) + bf16_hi(pa[nt][0]);                       // sums of the ROUNDED weights
            l1 += bf16_lo(pa[nt][1]) + bf16_hi(pa[nt][1]);
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const uint32_t a[4] = {pa[2 * kk][0], pa[2 * kk][1], pa[2 * kk + 1][0], pa[2 * kk + 1][1]};
#pragma unroll
            for (int n2 = 0; n2 < 8; ++n2) {
                uint32_t bb[4];
                ldsm_x4_t(bb, addr_bt(vs, BW_P128, kk * 16, n2 * 16, lane));
                mma16816(o[2 * n2], a, bb[0], bb[1]);
                mma16816(o[2 * n2 + 1], a, bb[2], bb[3]);
            }
        }
    }
    l0 = quad_sum(l0); l1 = quad_sum(l1);
    const float inv[2] = {1.f / l0, 1.f / l1};
    const float lse[2] = {m0 + log2f(l0), m1 + log2f(l1)};

    // element-wise part (adaDecoder.py:190-198 differentiated)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int n = q0 + warp * 16 + g + r * 8;
        const bool ok = n < Nc;
        float delta = 0.f;
        if (ok) {
            const size_t row = static_cast<size_t>(b) * Nc + n;
            const float* grow = p.g + row * C + h * 64;
            const __nv_bfloat16* xrow = p.x + row * C + h * 64;
            const float* mu = p.x_mean + static_cast<size_t>(b) * C + h * 64;
            const float* rs = p.x_rstd + static_cast<size_t>(b) * C + h * 64;
            float* dxrow = p.d_xhat + row * C + h * 64;
            __nv_bfloat16* dorow = p.d_o + row * 4 * C + h * 256;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = j * 8 + 2 * t;
                const float2 gg = *reinterpret_cast<const float2*>(grow + c);
                const uint32_t xw = *reinterpret_cast<const uint32_t*>(xrow + c);
                const float2 mm = *reinterpret_cast<const float2*>(mu + c), rr = *reinterpret_cast<const float2*>(rs + c);
                float dm[2], de[2], dx[2], mo[2], eo[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float M = o[j][2 * r + e] * inv[r], E = o[8 + j][2 * r + e] * inv[r];
                    const float gv = e ? gg.y : gg.x;
                    const float xh = ((e ? bf16_hi(xw) : bf16_lo(xw)) - (e ? mm.y : mm.x)) * (e ? rr.y : rr.x);
                    const float var = fmaf(-M, M, E);
                    const float sd = sqrtf(fmaxf(var, 1e-6f));
                    dx[e] = gv * sd;
                    const float dvar = var >= 1e-6f ? gv * xh / (2.f * sd) : 0.f;
                    dm[e] = fmaf(-2.f * M, dvar, gv);
                    de[e] = dvar;
                    mo[e] = M; eo[e] = E;
                }
                // dO' as TWO bf16 terms (hi + lo).  dA - delta = sum_c [dM (v - M) + dE (v^2 - E)] cancels between the
                // dM and the dE products when a row is sharp (dVar = g x^ / (2 sigma) grows as sigma -> 0): with one bf16
                // term the cancellation residue carries 2^-9 of the LARGE terms (measured: dV 28 % off at logit std 6.5).
                // delta is taken from the same rounded values, so sum_j P_ij dA_ij = delta_i holds for the operands used.
                const uint32_t wm = pack_bf16x2(dm[0], dm[1]), we = pack_bf16x2(de[0], de[1]);
                const uint32_t lm = pack_bf16x2(dm[0] - bf16_lo(wm), dm[1] - bf16_hi(wm));
                const uint32_t le = pack_bf16x2(de[0] - bf16_lo(we), de[1] - bf16_hi(we));
                delta = fmaf(bf16_lo(wm) + bf16_lo(lm), mo[0], fmaf(bf16_hi(wm) + bf16_hi(lm), mo[1], delta));
                delta = fmaf(bf16_lo(we) + bf16_lo(le), eo[0], fmaf(bf16_hi(we) + bf16_hi(le), eo[1], delta));
                *reinterpret_cast<float2*>(dxrow + c) = make_float2(dx[0], dx[1]);
                *reinterpret_cast<uint32_t*>(dorow + c) = wm;
                *reinterpret_cast<uint32_t*>(dorow + 64 + c) = we;
                *reinterpret_cast<uint32_t*>(dorow + 128 + c) = lm;
                *reinterpret_cast<uint32_t*>(dorow + 192 + c) = le;
            }
        }
        delta = quad_sum(delta);
        if (ok && t == 0) {
            const size_t i = (static_cast<size_t>(b) * p.H + h) * Nc + n;
            p.lse[i] = lse[r];
            p.delta[i] = delta;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// dQ = dS K, dS = P (dA - delta); one CTA per (query tile, head, image), loop over key tiles
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const BwdParams p) {
    __shared__ __align__(16) __nv_bfloat16 ks_[BW_T * BW_P64];
    __shared__ __align__(16) __nv_bfloat16 vs_[BW_T * BW_P128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int q0 = blockIdx.x * BW_T, h = blockIdx.y, b = blockIdx.z;
    const int C = p.C, Nc = p.Nc, Ns = p.Ns;
    const uint32_t ks = smem_u32(ks_), vs = smem_u32(vs_);
    const int qvalid = min(BW_T, Nc - q0);

    const __nv_bfloat16* dorow = p.d_o + (static_cast<size_t>(b) * Nc + q0) * 4 * C + h * 256;
    load_tile<64, BW_P64>(ks_, p.q + (static_cast<size_t>(b) * Nc + q0) * C + h * 64, C, qvalid);
    load_tile<128, BW_P128>(vs_, dorow, 4 * C, qvalid);
    cp_commit_wait();
    __syncthreads();
    uint32_t qf[4][4], dof[8][4], dol[8][4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) ldsm_x4(qf[kk], addr_a(ks, BW_P64, warp * 16, kk * 16, lane));
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) ldsm_x4(dof[kk], addr_a(vs, BW_P128, warp * 16, kk * 16, lane));
    __syncthreads();
    load_tile<128, BW_P128>(vs_, dorow + 128, 4 * C, qvalid);
    cp_commit_wait();
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) ldsm_x4(dol[kk], addr_a(vs, BW_P128, warp * 16, kk * 16, lane));
    float L[2], D[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int n = q0 + warp * 16 + g + r * 8;
        const size_t i = (static_cast<size_t>(b) * p.H + h) * Nc + n;
        L[r] = n < Nc ? p.lse[i] : INFINITY;               // rows past Nc: P = 2^(-inf) = 0
        D[r] = n < Nc ? p.delta[i] : 0.f;
    }
    float dq[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) dq[nt][0] = dq[nt][1] = dq[nt][2] = dq[nt][3] = 0.f;

    for (int k0 = 0; k0 < Ns; k0 += BW_T) {
        __syncthreads();
        const int valid = min(BW_T, Ns - k0);
        load_tile<64, BW_P64>(ks_, p.k + (static_cast<size_t>(b) * Ns + k0) * C + h * 64, C, valid);
        load_tile<128, BW_P128>(vs_, p.v + (static_cast<size_t>(b) * Ns + k0) * 2 * C + h * 128, 2 * C, valid);
        cp_commit_wait();
        __syncthreads();
        float s[8][4], da[8][4];
        logits_tile(s, qf, ks, lane);
        dattn_tile<true>(da, dof, vs, lane);
        dattn_tile<false>(da, dol, vs, lane);
        uint32_t ds[8][2];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int c = nt * 8 + 2 * t;
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int r = e >> 1;
                v[e] = ex2_approx(s[nt][e] - L[r]) * (da[nt][e] - D[r]);
            }
            if (c >= valid) v[0] = v[2] = 0.f;              // keys past Ns (their K rows are zero anyway)
            if (c + 1 >= valid) v[1] = v[3] = 0.f;
            ds[nt][0] = pack_bf16x2(v[0], v[1]);
            ds[nt][1] = pack_bf16x2(v[2], v[3]);
        }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const uint32_t a[4] = {ds[2 * kk][0], ds[2 * kk][1], ds[2 * kk + 1][0], ds[2 * kk + 1][1]};
#pragma unroll
            for (int n2 = 0; n2 < 4; ++n2) {
                uint32_t bb[4];
                ldsm_x4_t(bb, addr_bt(ks, BW_P64, kk * 16, n2 * 16, lane));
                mma16816(dq[2 * n2], a, bb[0], bb[1]);
                mma16816(dq[2 * n2 + 1], a, bb[2], bb[3]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int n = q0 + warp * 16 + g + r * 8;
        if (n >= Nc) continue;
        __nv_bfloat16* row = p.d_q + (static_cast<size_t>(b) * Nc + n) * C + h * 64;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
            *reinterpret_cast<uint32_t*>(row + nt * 8 + 2 * t) = pack_bf16x2(dq[nt][2 * r], dq[nt][2 * r + 1]);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// dV' = P^T dO', dK = dS^T Q; one CTA per (key tile, head, image), loop over query tiles.  Step A: warp = 16 query rows
// (S, dA, P, dS -> shared memory); step B: warp = 16 keys (the contraction runs over the 64 query rows of the tile).
// ---------------------------------------------------------------------------------------------------------------------
constexpr size_t BW_DKV_SMEM = (4 * BW_T * BW_P64 + 3 * BW_T * BW_P128) * 2 + 2 * BW_T * 4;

__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(const BwdParams p) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __nv_bfloat16* ks_ = reinterpret_cast<__nv_bfloat16*>(smem_raw);
    __nv_bfloat16* qs_ = ks_ + BW_T * BW_P64;
    __nv_bfloat16* ps_ = qs_ + BW_T * BW_P64;
    __nv_bfloat16* dss_ = ps_ + BW_T * BW_P64;
    __nv_bfloat16* vs_ = dss_ + BW_T * BW_P64;
    __nv_bfloat16* dos_ = vs_ + BW_T * BW_P128;
    __nv_bfloat16* dol_ = dos_ + BW_T * BW_P128;
    float* ls_ = reinterpret_cast<float*>(dol_ + BW_T * BW_P128);
    float* dl_ = ls_ + BW_T;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int k0 = blockIdx.x * BW_T, h = blockIdx.y, b = blockIdx.z;
    const int C = p.C, Nc = p.Nc, Ns = p.Ns;
    const uint32_t ks = smem_u32(ks_), qs = smem_u32(qs_), ps = smem_u32(ps_), dss = smem_u32(dss_), vs = smem_u32(vs_),
                   dos = smem_u32(dos_), dol = smem_u32(dol_);
    const int kvalid = min(BW_T, Ns - k0);

    load_tile<64, BW_P64>(ks_, p.k + (static_cast<size_t>(b) * Ns + k0) * C + h * 64, C, kvalid);
    load_tile<128, BW_P128>(vs_, p.v + (static_cast<size_t>(b) * Ns + k0) * 2 * C + h * 128, 2 * C, kvalid);

    float dv[16][4], dk[8][4];
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) dv[nt][0] = dv[nt][1] = dv[nt][2] = dv[nt][3] = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) dk[nt][0] = dk[nt][1] = dk[nt][2] = dk[nt][3] = 0.f;

    for (int q0 = 0; q0 < Nc; q0 += BW_T) {
        __syncthreads();                                   // step B of the previous tile is done with Q, dO', P, dS
        const int qvalid = min(BW_T, Nc - q0);
        load_tile<64, BW_P64>(qs_, p.q + (static_cast<size_t>(b) * Nc + q0) * C + h * 64, C, qvalid);
        load_tile<128, BW_P128>(dos_, p.d_o + (static_cast<size_t>(b) * Nc + q0) * 4 * C + h * 256, 4 * C, qvalid);
        load_tile<128, BW_P128>(dol_, p.d_o + (static_cast<size_t>(b) * Nc + q0) * 4 * C + h * 256 + 128, 4 * C, qvalid);
        if (threadIdx.x < BW_T) {
            const int n = q0 + threadIdx.x;
            const size_t i = (static_cast<size_t>(b) * p.H + h) * Nc + n;
            ls_[threadIdx.x] = n < Nc ? p.lse[i] : INFINITY;
            dl_[threadIdx.x] = n < Nc ? p.delta[i] : 0.f;
        }
        cp_commit_wait();
        __syncthreads();
        {   // step A (A fragments straight from shared memory: Q and dO' stay resident for step B)
            float s[8][4], da[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
                da[nt][0] = da[nt][1] = da[nt][2] = da[nt][3] = 0.f;
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                uint32_t a[4];
                ldsm_x4(a, addr_a(qs, BW_P64, warp * 16, kk * 16, lane));
#pragma unroll
                for (int n2 = 0; n2 < 4; ++n2) {
                    uint32_t bb[4];
                    ldsm_x4(bb, addr_b(ks, BW_P64, n2 * 16, kk * 16, lane));
                    mma16816(s[2 * n2], a, bb[0], bb[1]);
                    mma16816(s[2 * n2 + 1], a, bb[2], bb[3]);
                }
            }
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                uint32_t a[4], al[4];
                ldsm_x4(a, addr_a(dos, BW_P128, warp * 16, kk * 16, lane));
                ldsm_x4(al, addr_a(dol, BW_P128, warp * 16, kk * 16, lane));
#pragma unroll
                for (int n2 = 0; n2 < 4; ++n2) {
                    uint32_t bb[4];
                    ldsm_x4(bb, addr_b(vs, BW_P128, n2 * 16, kk * 16, lane));
                    mma16816(da[2 * n2], a, bb[0], bb[1]);
                    mma16816(da[2 * n2 + 1], a, bb[2], bb[3]);
                    mma16816(da[2 * n2], al, bb[0], bb[1]);
                    mma16816(da[2 * n2 + 1], al, bb[2], bb[3]);
                }
            }
            const int r0 = warp * 16 + g;
            const float L0 = ls_[r0], L1 = ls_[r0 + 8], D0 = dl_[r0], D1 = dl_[r0 + 8];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const int c = nt * 8 + 2 * t;
                const float p0 = ex2_approx(s[nt][0] - L0), p1 = ex2_approx(s[nt][1] - L0);
                const float p2 = ex2_approx(s[nt][2] - L1), p3 = ex2_approx(s[nt][3] - L1);
                *reinterpret_cast<uint32_t*>(ps_ + r0 * BW_P64 + c) = pack_bf16x2(p0, p1);
                *reinterpret_cast<uint32_t*>(ps_ + (r0 + 8) * BW_P64 + c) = pack_bf16x2(p2, p3);
                *reinterpret_cast<uint32_t*>(dss_ + r0 * BW_P64 + c) = pack_bf16x2(p0 * (da[nt][0] - D0), p1 * (da[nt][1] - D0));
                *reinterpret_cast<uint32_t*>(dss_ + (r0 + 8) * BW_P64 + c) = pack_bf16x2(p2 * (da[nt][2] - D1), p3 * (da[nt][3] - D1));
            }
        }
        __syncthreads();
        // step B: keys warp*16 .. +15
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {                   // 16 query rows per step
            uint32_t ap[4], ad[4];
            ldsm_x4_t(ap, addr_at(ps, BW_P64, kk * 16, warp * 16, lane));
            ldsm_x4_t(ad, addr_at(dss, BW_P64, kk * 16, warp * 16, lane));
#pragma unroll
            for (int n2 = 0; n2 < 8; ++n2) {
                uint32_t bb[4], bl[4];
                ldsm_x4_t(bb, addr_bt(dos, BW_P128, kk * 16, n2 * 16, lane));
                ldsm_x4_t(bl, addr_bt(dol, BW_P128, kk * 16, n2 * 16, lane));
                mma16816(dv[2 * n2], ap, bb[0], bb[1]);
                mma16816(dv[2 * n2 + 1], ap, bb[2], bb[3]);
                mma16816(dv[2 * n2], ap, bl[0], bl[1]);
                mma16816(dv[2 * n2 + 1], ap, bl[2], bl[3]);
            }
#pragma unroll
            for (int n2 = 0; n2 < 4; ++n2) {
                uint32_t bb[4];
                ldsm_x4_t(bb, addr_bt(qs, BW_P64, kk * 16, n2 * 16, lane));
                mma16816(dk[2 * n2], ad, bb[0], bb[1]);
                mma16816(dk[2 * n2 + 1], ad, bb[2], bb[3]);
            }
        }
    }
    // dK (Q was stored pre-multiplied by log2 e), dV = dV'[:, :64] + 2 V~ dV'[:, 64:]
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int kr = warp * 16 + g + r * 8;
        if (kr >= kvalid) continue;
        const size_t row = static_cast<size_t>(b) * Ns + k0 + kr;
        __nv_bfloat16* dkrow = p.d_k + row * C + h * 64;
        __nv_bfloat16* dvrow = p.d_v + row * C + h * 64;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int c = nt * 8 + 2 * t;
            *reinterpret_cast<uint32_t*>(dkrow + c) = pack_bf16x2(dk[nt][2 * r] * BW_LN2, dk[nt][2 * r + 1] * BW_LN2);
            const uint32_t vw = *reinterpret_cast<const uint32_t*>(vs_ + kr * BW_P128 + c);
            const float d0 = fmaf(2.f * bf16_lo(vw), dv[8 + nt][2 * r], dv[nt][2 * r]);
            const float d1 = fmaf(2.f * bf16_hi(vw), dv[8 + nt][2 * r + 1], dv[nt][2 * r + 1]);
            *reinterpret_cast<uint32_t*>(dvrow + c) = pack_bf16x2(d0, d1);
        }
    }
}

}  // namespace

size_t attn_bwd_scratch_floats(int B, int H, int Nc) { return static_cast<size_t>(B) * H * Nc; }

// q, k, v: what the forward projections wrote (bf16 path); g: dL/d(heads) f32; outputs as in BwdParams
int launch_attn_bwd(int B, int H, int Nc, int Ns, int C, const void* q, const void* k, const void* v, const void* x,
                    const float* x_mean, const float* x_rstd, const float* g, void* d_o, float* lse, float* delta,
                    float* d_xhat, void* d_q, void* d_k, void* d_v, cudaStream_t s) {
    if (C != H * 64) {
        set_error("attn_bwd: head_dim 64 only (C=%d, H=%d)", C, H);
        return MHADA_ERR_UNSUPPORTED;
    }
    BwdParams p{};
    p.B = B; p.H = H; p.Nc = Nc; p.Ns = Ns; p.C = C;
    p.q = static_cast<const __nv_bfloat16*>(q); p.k = static_cast<const __nv_bfloat16*>(k);
    p.v = static_cast<const __nv_bfloat16*>(v); p.x = static_cast<const __nv_bfloat16*>(x);
    p.x_mean = x_mean; p.x_rstd = x_rstd; p.g = g;
    p.d_o = static_cast<__nv_bfloat16*>(d_o); p.lse = lse; p.delta = delta; p.d_xhat = d_xhat;
    p.d_q = static_cast<__nv_bfloat16*>(d_q); p.d_k = static_cast<__nv_bfloat16*>(d_k); p.d_v = static_cast<__nv_bfloat16*>(d_v);
    const dim3 gq(static_cast<unsigned>((Nc + BW_T - 1) / BW_T), static_cast<unsigned>(H), static_cast<unsigned>(B));
    const dim3 gk(static_cast<unsigned>((Ns + BW_T - 1) / BW_T), static_cast<unsigned>(H), static_cast<unsigned>(B));
    attn_bwd_prep_kernel<<<gq, 128, 0, s>>>(p);
    count_launch();
    attn_bwd_dq_kernel<<<gq, 128, 0, s>>>(p);
    count_launch();
    static DeviceOnce once;
    if (int e = smem_attr_once(once, reinterpret_cast<const void*>(attn_bwd_dkv_kernel), BW_DKV_SMEM, "attn_bwd_dkv smem attr")) return e;
    attn_bwd_dkv_kernel<<<gk, 128, BW_DKV_SMEM, s>>>(p);
    count_launch();
    return check_cuda(cudaGetLastError(), "attn_bwd launch");
}

}  // namespace mh

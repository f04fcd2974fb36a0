// Last decoder block: ReflectionPad2d(1) + Conv2d(64, 3, 3) + ReLU in ONE kernel.
//
// Replaces conv3[1] = ConvReLU(64, 3, 3, 1) of MHAdaSTr/network/conv.py:90-93 (pad :26-27, conv :28, ReLU :43).
// With three output channels the block is HBM-bound (B*H*W*64 bf16 read once, 3 planes written): the library
// route pays a full padded copy of the 64-channel 512 x 512 activation (read + write) and an implicit GEMM whose
// N is padded to 64.  Here a CTA stages a (8+2) x (32+2) x 64 input patch in shared memory with the reflection
// resolved in the load indices (no padded tensor exists) and computes the block in two phases:
//   1. tap products: for EVERY patch pixel p (halo included) and every (tap t, output channel co),
//      part[co*9+t][p] = sum_ci x[p][ci] * w[co][ci][t] -- one GEMM [340 pixels] x [64] x [9*Cout <= 32] run as
//      mma.sync.m16n8k16 (bf16 in, fp32 accumulate) over m16 tiles of consecutive patch pixels.  Every input
//      fragment is read from shared memory ONCE (ldmatrix) and serves all nine taps; the first version of this
//      kernel looped over the taps and re-read each fragment nine times, which made it shared-memory bound
//      (~440 KB of ldmatrix traffic per tile, 152 us for batch 8 of 512 x 512).
//   2. shift-add: out[y][x][co] = bias + sum_t part[co*9+t][(y+dy)*34 + x+dx], fp32, then ReLU and bf16.
// The tap products live in shared memory column-major ([8*NT][356] floats: a pitch of 4 mod 16 makes the accumulator
// stores of a warp hit 32 distinct banks; the shift-add reads consecutive pixels).  The warp-level MMA is used on
// purpose: the math is ~2 % of the tcgen05 peak at the HBM-bound rate; what matters is that every input byte is
// read from HBM once and from shared memory once.
//
// x [B, H, W, 64] bf16 channels_last (NOT padded)  ->  y [B, Cout, H, W] bf16 (NCHW planes, like the reference).
// Algorithmic bytes: 2*B*H*W*64 read + 2*B*H*W*Cout written.
#include "common.h"
#include "ptx.cuh"

namespace mh {

constexpr int CS_TH = 8, CS_TW = 32, CS_CIN = 64, CS_PIX = 72;   // pixel pitch 72 bf16 = 144 B: ldmatrix rows hit distinct banks
constexpr int CS_PH = CS_TH + 2, CS_PW = CS_TW + 2;
constexpr int CS_NPIX = CS_PH * CS_PW;                            // 340 patch pixels
constexpr int CS_MT = (CS_NPIX + 15) / 16;                        // 22 m16 tiles of consecutive patch pixels
constexpr int CS_PP = 356;                                        // pitch of one column of tap products (>= 16 * CS_MT)
constexpr int CS_THREADS = 256;
constexpr size_t CS_PATCH_BYTES = static_cast<size_t>(CS_MT) * 16 * CS_PIX * 2;
static_assert(CS_PP >= CS_MT * 16 && (CS_PP % 16 == 4 || CS_PP % 16 == 12),
              "column-major tap products: unpredicated, bank-conflict-free accumulator stores");
static_assert(CS_TH * CS_TW == CS_THREADS, "shift-add phase: one thread per output pixel");
static_assert(CS_PW * (CS_CIN / 8) <= 2 * CS_THREADS, "patch staging: at most two 16-byte chunks per thread and row");

constexpr size_t cs_smem_bytes(int nt) { return CS_PATCH_BYTES + static_cast<size_t>(8) * nt * CS_PP * sizeof(float); }

__device__ __forceinline__ int reflect_clamp(int i, int n) {   // ReflectionPad2d(1) inside the image, clamped beyond it
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return min(max(i, 0), n - 1);
}

// NT = n8 tiles of the tap-product GEMM, column n = co * 9 + t: 4 covers Cout <= 3 (the decoder's case, two CTAs per
// SM), 9 covers Cout <= 8.  One tile per CTA: a persistent form with a three-deep cp.async ring of patches (one CTA
// per SM) measured slower (144 us against 130): with eight warps per SM the phases below are latency-bound, two
// co-resident CTAs overlap them better than prefetching does.
template <int NT>
__global__ void __launch_bounds__(CS_THREADS, NT == 4 ? 2 : 1)
conv3x3_c64_small_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                         __nv_bfloat16* __restrict__ y, int B, int H, int W, int Cout, int relu) {
    extern __shared__ __align__(16) uint8_t cs_smem[];
    __nv_bfloat16* patch = reinterpret_cast<__nv_bfloat16*>(cs_smem);
    float* part = reinterpret_cast<float*>(cs_smem + CS_PATCH_BYTES);
    __shared__ uint32_t wfrag[4 * NT * 2 * 32];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int x0 = blockIdx.x * CS_TW, y0 = blockIdx.y * CS_TH, b = blockIdx.z;
    const __nv_bfloat16* xb = x + static_cast<size_t>(b) * H * W * CS_CIN;

    // ---- stage the input patch (reflection resolved here), 16 bytes per cp.async.  A patch row is 34 pixels x 8
    //      chunks = 272 chunks: thread tid copies chunk tid (and tid + 256 for tid < 16), so the column arithmetic is
    //      done once per thread and the row arithmetic is CTA-uniform.
    {
        const int chunk = tid & 7, px = tid >> 3;
        const int colA = reflect_clamp(x0 + px - 1, W) * CS_CIN + chunk * 8;
        const int colB = reflect_clamp(x0 + px + 31, W) * CS_CIN + chunk * 8;
        const uint32_t dstA = smem_u32(patch + px * CS_PIX + chunk * 8);
        const bool second = tid < (CS_PW - 32) * 8;
#pragma unroll
        for (int py = 0; py < CS_PH; ++py) {
            const __nv_bfloat16* row = xb + static_cast<size_t>(reflect_clamp(y0 + py - 1, H)) * W * CS_CIN;
            const uint32_t d = dstA + py * CS_PW * CS_PIX * 2;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(row + colA) : "memory");
            if (second)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 32 * CS_PIX * 2), "l"(row + colB) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        // the rows that pad the last m16 tile: zeroed (their products are stored but never read)
        if (tid < (CS_MT * 16 - CS_NPIX) * (CS_CIN / 8))
            *reinterpret_cast<uint4*>(patch + (CS_NPIX + px) * CS_PIX + chunk * 8) = make_uint4(0, 0, 0, 0);
    }

    // ---- weights: one coalesced read of w [Cout][64][3][3] into shared memory (the tap-product area is free until
    //      phase 1), then the B fragments in fragment order: wfrag[kc][nt][hh][lane] = the b16x2 word lane `lane` feeds
    //      to the MMA of k-chunk kc and n-tile nt: b0 (hh = 0): k = (lane % 4) * 2 + {0, 1}, b1 (hh = 1): k + 8;
    //      n = nt * 8 + lane / 4.  (Reading the fragments straight from global memory cost 8 scattered sectors per
    //      thread: as much L2 traffic as the activations.)
    for (int i = tid; i < Cout * CS_CIN * 9; i += CS_THREADS) part[i] = __ldg(w + i);
    __syncthreads();
    for (int i = tid; i < 4 * NT * 2 * 32; i += CS_THREADS) {
        const int l = i & 31, hh = (i >> 5) & 1, nt = (i >> 6) % NT, kc = (i >> 6) / NT;
        const int n = nt * 8 + (l >> 2), ci = kc * 16 + hh * 8 + (l & 3) * 2;
        const int co = n / 9, t = n % 9;
        float v0 = 0.f, v1 = 0.f;
        if (co < Cout) {
            v0 = part[(co * CS_CIN + ci) * 9 + t];
            v1 = part[(co * CS_CIN + ci + 1) * 9 + t];
        }
        wfrag[i] = pack_bf16x2(v0, v1);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // ---- phase 1: tap products.  The weight fragments stay in registers across the warp's m16 tiles.
    uint32_t bfr[4][NT][2];
#pragma unroll
    for (int kc = 0; kc < 4; ++kc)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            bfr[kc][nt][0] = wfrag[((kc * NT + nt) * 2 + 0) * 32 + lane];
            bfr[kc][nt][1] = wfrag[((kc * NT + nt) * 2 + 1) * 32 + lane];
        }
    const int lrow = lane & 15, lk = (lane >> 4) * 8;        // ldmatrix.x4: lane -> (matrix row, k half)
    for (int mt = warp; mt < CS_MT; mt += CS_THREADS / 32) {
        float acc[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
        const __nv_bfloat16* arow = patch + (mt * 16 + lrow) * CS_PIX + lk;
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
            uint32_t a0, a1, a2, a3;
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                         : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3)
                         : "r"(smem_u32(arow + kc * 16)));
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
                asm volatile(
                    "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                    : "+f"(acc[nt][0]), "+f"(acc[nt][1]), "+f"(acc[nt][2]), "+f"(acc[nt][3])
                    : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(bfr[kc][nt][0]), "r"(bfr[kc][nt][1]));
        }
        // accumulator (row = lane/4 [+8], col = (lane%4)*2 [+1]) -> part[col][pixel]; every offset below is an immediate
        float* dst = part + (lane & 3) * 2 * CS_PP + mt * 16 + (lane >> 2);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[(nt * 8 + (i & 1)) * CS_PP + (i >> 1) * 8] = acc[nt][i];
    }
    __syncthreads();

    // ---- phase 2: shift-add of the nine tap products, bias, ReLU, bf16; thread = output pixel (row = warp, px = lane)
    const int gy = y0 + warp, gx = x0 + lane;
    if (gy < H && gx < W) {
        const float* src = part + warp * CS_PW + lane;
        __nv_bfloat16* dst = y + (static_cast<size_t>(b) * Cout * H + gy) * W + gx;
        for (int co = 0; co < Cout; ++co, src += 9 * CS_PP, dst += static_cast<size_t>(H) * W) {
            float v = __ldg(bias + co);
#pragma unroll
            for (int t = 0; t < 9; ++t) v += src[t * CS_PP + (t / 3) * CS_PW + (t % 3)];
            if (relu) v = fmaxf(v, 0.f);
            *dst = __float2bfloat16_rn(v);
        }
    }
}

template <int NT>
static int launch_conv_small_nt(const void* x, const float* w, const float* bias, int B, int H, int W, int Cout, int relu,
                                void* y, cudaStream_t s) {
    static DeviceOnce once;
    if (int e = smem_attr_once(once, reinterpret_cast<const void*>(conv3x3_c64_small_kernel<NT>), cs_smem_bytes(NT), "conv_small smem attr"))
        return e;
    dim3 grid((W + CS_TW - 1) / CS_TW, (H + CS_TH - 1) / CS_TH, B);
    conv3x3_c64_small_kernel<NT><<<grid, CS_THREADS, cs_smem_bytes(NT), s>>>(
        static_cast<const __nv_bfloat16*>(x), w, bias, static_cast<__nv_bfloat16*>(y), B, H, W, Cout, relu);
    count_launch();
    return check_cuda(cudaGetLastError(), "conv3x3_small launch");
}

int launch_conv3x3_small(const void* x, const float* w, const float* bias, int B, int H, int W, int Cin, int Cout, int relu,
                         void* y, cudaStream_t s) {
    if (Cin != CS_CIN || Cout < 1 || Cout > 8) {
        set_error("conv3x3_small: implemented for Cin = 64 and Cout <= 8, got %d -> %d", Cin, Cout);
        return MHADA_ERR_UNSUPPORTED;
    }
    return Cout <= 3 ? launch_conv_small_nt<4>(x, w, bias, B, H, W, Cout, relu, y, s)
                     : launch_conv_small_nt<9>(x, w, bias, B, H, W, Cout, relu, y, s);
}

}  // namespace mh

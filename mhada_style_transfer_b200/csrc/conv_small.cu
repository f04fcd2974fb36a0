// Last decoder block: ReflectionPad2d(1) + Conv2d(64, 3, 3) + ReLU in ONE kernel.
//
// Replaces conv3[1] = ConvReLU(64, 3, 3, 1) of MHAdaSTr/network/conv.py:90-93 (pad :26-27, conv :28, ReLU :43).
// With three output channels the block is HBM-bound (B*H*W*64 bf16 read once, 3 planes written): the library
// route pays a full padded copy of the 64-channel 512 x 512 activation (read + write) and an implicit GEMM whose
// N is padded to 64.  Here a CTA stages a (8+2) x (32+2) x 64 input patch in shared memory with the reflection
// resolved in the load indices (no padded tensor exists), and eight warps run the 3x3 taps as
// mma.sync.m16n8k16 (bf16 in, fp32 accumulate; N = 8 holds the <= 8 output channels).  The tensor-core shape
// is the smallest there is on purpose: the math is 2 % of the tcgen05 peak at the HBM-bound rate, so the
// legacy warp-level MMA keeps the kernel simple; what matters is that every input byte is read from HBM once.
//
// x [B, H, W, 64] bf16 channels_last (NOT padded)  ->  y [B, Cout, H, W] bf16 (NCHW planes, like the reference).
// Algorithmic bytes: 2*B*H*W*64 read + 2*B*H*W*Cout written.
#include "common.h"
#include "ptx.cuh"

namespace mh {

constexpr int CS_TH = 8, CS_TW = 32, CS_CIN = 64, CS_PIX = 72;   // pixel pitch 72 bf16 = 144 B: ldmatrix rows hit distinct banks
constexpr int CS_PH = CS_TH + 2, CS_PW = CS_TW + 2;
constexpr int CS_THREADS = 256;
constexpr size_t CS_SMEM = static_cast<size_t>(CS_PH) * CS_PW * CS_PIX * 2;

__device__ __forceinline__ int reflect_clamp(int i, int n) {   // ReflectionPad2d(1) inside the image, clamped beyond it
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return min(max(i, 0), n - 1);
}

__global__ void __launch_bounds__(CS_THREADS)
conv3x3_c64_small_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                         __nv_bfloat16* __restrict__ y, int B, int H, int W, int Cout, int relu) {
    extern __shared__ __align__(16) uint8_t cs_smem[];
    __nv_bfloat16* patch = reinterpret_cast<__nv_bfloat16*>(cs_smem);
    __shared__ __nv_bfloat16 out_s[8][CS_TH][CS_TW];
    __shared__ uint32_t wfrag[9 * 4 * 2 * 32];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int x0 = blockIdx.x * CS_TW, y0 = blockIdx.y * CS_TH, b = blockIdx.z;
    const __nv_bfloat16* xb = x + static_cast<size_t>(b) * H * W * CS_CIN;

    // ---- stage the input patch (reflection resolved here), 16 bytes per cp.async
    for (int i = tid; i < CS_PH * CS_PW * (CS_CIN / 8); i += CS_THREADS) {
        const int chunk = i & 7, pix = i >> 3;
        const int py = pix / CS_PW, px = pix % CS_PW;
        const int gy = reflect_clamp(y0 + py - 1, H), gx = reflect_clamp(x0 + px - 1, W);
        const __nv_bfloat16* src = xb + (static_cast<size_t>(gy) * W + gx) * CS_CIN + chunk * 8;
        const uint32_t dst = smem_u32(patch + (py * CS_PW + px) * CS_PIX + chunk * 8);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");

    // ---- B fragments (weights) in shared memory, fragment-major: wfrag[t][kc][hh][lane] = the b16x2 word lane `lane`
    //      feeds to mma number (t, kc): b0 (hh = 0): k = (lane % 4) * 2 + {0, 1}, b1 (hh = 1): k + 8; n = lane / 4.
    //      Built once per CTA from the fp32 weights (18 loads per thread); a per-MMA global load of the weights made
    //      the first version LSU-bound (220 us for batch 8 at 512 x 512).
    for (int i = tid; i < 9 * 4 * 2 * 32; i += CS_THREADS) {
        const int l = i & 31, hh = (i >> 5) & 1, kc = (i >> 6) & 3, t = i >> 8;
        const int n = l >> 2, ci = kc * 16 + hh * 8 + (l & 3) * 2;
        float v0 = 0.f, v1 = 0.f;
        if (n < Cout) {
            v0 = __ldg(w + (static_cast<size_t>(n) * CS_CIN + ci) * 9 + t);        // w [Cout][64][3][3]
            v1 = __ldg(w + (static_cast<size_t>(n) * CS_CIN + ci + 1) * 9 + t);
        }
        wfrag[i] = pack_bf16x2(v0, v1);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // ---- warp = one output row of the tile, two m16 tiles of 16 pixels
    float acc[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[mt][i] = 0.f;
    const int lrow = lane & 15, lk = (lane >> 4) * 8;        // ldmatrix.x4: lane -> (matrix row, k half)
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const int dy = t / 3, dx = t % 3;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            const __nv_bfloat16* arow = patch + ((warp + dy) * CS_PW + mt * 16 + lrow + dx) * CS_PIX + lk;
#pragma unroll
            for (int kc = 0; kc < 4; ++kc) {
                uint32_t a0, a1, a2, a3;
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                             : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3)
                             : "r"(smem_u32(arow + kc * 16)));
                asm volatile(
                    "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                    : "+f"(acc[mt][0]), "+f"(acc[mt][1]), "+f"(acc[mt][2]), "+f"(acc[mt][3])
                    : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(wfrag[((t * 4 + kc) * 2 + 0) * 32 + lane]),
                      "r"(wfrag[((t * 4 + kc) * 2 + 1) * 32 + lane]));
            }
        }
    }

    // ---- epilogue: bias, ReLU, bf16; accumulator (row = lane/4 [+8], col = (lane%4)*2 [+1]) -> out_s[co][row][px]
    {
        const int co0 = (lane & 3) * 2, r0 = lane >> 2;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int co = co0 + (i & 1), px = mt * 16 + r0 + (i >> 1) * 8;
                if (co < Cout) {
                    float v = acc[mt][i] + __ldg(bias + co);
                    if (relu) v = fmaxf(v, 0.f);
                    out_s[co][warp][px] = __float2bfloat16_rn(v);
                }
            }
    }
    __syncthreads();
    for (int i = tid; i < Cout * CS_TH * CS_TW; i += CS_THREADS) {
        const int px = i % CS_TW, row = (i / CS_TW) % CS_TH, co = i / (CS_TW * CS_TH);
        const int gy = y0 + row, gx = x0 + px;
        if (gy < H && gx < W) y[((static_cast<size_t>(b) * Cout + co) * H + gy) * W + gx] = out_s[co][row][px];
    }
}

int launch_conv3x3_small(const void* x, const float* w, const float* bias, int B, int H, int W, int Cin, int Cout, int relu,
                         void* y, cudaStream_t s) {
    if (Cin != CS_CIN || Cout < 1 || Cout > 8) {
        set_error("conv3x3_small: implemented for Cin = 64 and Cout <= 8, got %d -> %d", Cin, Cout);
        return MHADA_ERR_UNSUPPORTED;
    }
    static bool attr_done = false;
    if (!attr_done) {
        if (int e = check_cuda(cudaFuncSetAttribute(conv3x3_c64_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    static_cast<int>(CS_SMEM)), "conv_small smem attr"))
            return e;
        attr_done = true;
    }
    dim3 grid((W + CS_TW - 1) / CS_TW, (H + CS_TH - 1) / CS_TH, B);
    conv3x3_c64_small_kernel<<<grid, CS_THREADS, CS_SMEM, s>>>(static_cast<const __nv_bfloat16*>(x), w, bias,
                                                             static_cast<__nv_bfloat16*>(y), B, H, W, Cout, relu);
    count_launch();
    return check_cuda(cudaGetLastError(), "conv3x3_small launch");
}

}  // namespace mh

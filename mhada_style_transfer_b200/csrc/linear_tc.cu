// out_conv on the tensor cores: y[M, Cout] = x[M, Cin] . W[Cout, Cin]^T + bias   (bf16 in/out, f32 accumulate)
//
// Replaces torch.cat + nn.Conv2d(C, C, 1) at MHAdaSTr/network/adaDecoder.py:202-205 (the concat is
// free: every head's attention epilogue already wrote its 64-channel slice of the same [M, C] buffer).
//
// Structure (persistent CTAs, two per SM, 6 warps; a work item is one 128 x BN output tile, the BN tiles of the
// same 128 rows are consecutive items so that x is read from HBM once and from L2 by the other column tiles):
//   warp 0      TMA producer : x and W tiles -> 128B-swizzled shared memory, 3-stage mbarrier ring that runs
//               straight across work items (the loads of item n+1 start while item n is still in the MMAs)
//   warp 1      tcgen05.mma issuer (one elected lane); accumulators DOUBLE-BUFFERED in TMEM (2 x BN columns):
//               the MMAs of item n+1 overlap the epilogue of item n
//   warps 2..5  epilogue     : tcgen05.ld (lane = row) -> + bias -> bf16 -> 256-bit global stores
// HBM-bound at C = 512: algorithmic bytes = M*Cin*2 (read) + M*Cout*2 (write) (+ 0.5 MB weights).
#include "common.h"
#include "ptx.cuh"

namespace mh {

constexpr int LIN_BM = 128, LIN_BK = 64, LIN_STAGES = 3, LIN_THREADS = 192;

__global__ void f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
    size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i < n) dst[i] = __float2bfloat16_rn(src[i]);
}

template <int BN>
__global__ void __launch_bounds__(LIN_THREADS, 2)
linear_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const float* __restrict__ bias, __nv_bfloat16* __restrict__ y, int ldy, int M, int ktiles, int ntiles,
                 int items) {
    constexpr uint32_t A_BYTES = LIN_BM * LIN_BK * 2, B_BYTES = BN * LIN_BK * 2, STAGE = A_BYTES + B_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[LIN_STAGES], empty[LIN_STAGES], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < LIN_STAGES; ++s) {
                mbar_init(&full[s], 1);
                mbar_init(&empty[s], 1);
            }
            for (int u = 0; u < 2; ++u) {
                mbar_init(&acc_full[u], 1);
                mbar_init(&acc_empty[u], 4);      // one arrive per epilogue warp
            }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(&tmem_slot, 2 * BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            int g = 0;                                            // running k-tile counter of this CTA
            for (int it = blockIdx.x; it < items; it += gridDim.x) {
                const int m0 = (it / ntiles) * LIN_BM, n0 = (it % ntiles) * BN;
                for (int kt = 0; kt < ktiles; ++kt, ++g) {
                    const int s = g % LIN_STAGES;
                    mbar_wait(&empty[s], ((g / LIN_STAGES) & 1) ^ 1);
                    mbar_arrive_expect_tx(&full[s], STAGE);
                    uint8_t* a = smem + s * STAGE;
                    tma_load_2d(a, &tmA, &full[s], kt * LIN_BK, m0);
                    tma_load_2d(a + A_BYTES, &tmB, &full[s], kt * LIN_BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(LIN_BM, BN, 0, 0);
            int g = 0, n = 0;
            for (int it = blockIdx.x; it < items; it += gridDim.x, ++n) {
                const int u = n & 1;
                mbar_wait(&acc_empty[u], ((n >> 1) & 1) ^ 1);     // the epilogue has drained this accumulator buffer
                tc_fence_after();
                for (int kt = 0; kt < ktiles; ++kt, ++g) {
                    const int s = g % LIN_STAGES;
                    mbar_wait(&full[s], (g / LIN_STAGES) & 1);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + s * STAGE);
                    const uint64_t da = make_smem_desc(a_addr, 16, 1024);
                    const uint64_t db = make_smem_desc(a_addr + A_BYTES, 16, 1024);
#pragma unroll
                    for (int k = 0; k < LIN_BK / 16; ++k)
                        umma_ss(tmem + u * BN, desc_advance(da, k * 32), desc_advance(db, k * 32), idesc, (kt | k) != 0);
                    umma_commit(&empty[s]);
                }
                umma_commit(&acc_full[u]);
            }
        }
    } else {
        // epilogue warps 2..5 -> TMEM lane quarters 2,3,0,1
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        int n = 0;
        for (int it = blockIdx.x; it < items; it += gridDim.x, ++n) {
            const int m = (it / ntiles) * LIN_BM + row, n0 = (it % ntiles) * BN;
            const int u = n & 1;
            mbar_wait(&acc_full[u], (n >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
                uint32_t r[32];
                tmem_ld_x32(tmem_addr(tmem, quarter * 32, u * BN + c), r);
                tmem_wait_ld();
                if (m < M) {
                    uint32_t o[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float v0 = __uint_as_float(r[2 * i]) + __ldg(bias + n0 + c + 2 * i);
                        float v1 = __uint_as_float(r[2 * i + 1]) + __ldg(bias + n0 + c + 2 * i + 1);
                        o[i] = pack_bf16x2(v0, v1);
                    }
                    __nv_bfloat16* dst = y + static_cast<size_t>(m) * ldy + n0 + c;
                    st_global_256(dst, o);
                    st_global_256(dst + 16, o + 8);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[u]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 2 * BN);
}

size_t linear_bf16_workspace(int Cout, int Cin) { return align_up(static_cast<size_t>(Cout) * Cin * 2, 256); }

template <int BN>
static int launch_linear_tc(const CUtensorMap& tmA, const void* wbf, const float* bias, int M, int Cin, int Cout,
                            void* y, int ldy, cudaStream_t s) {
    CUtensorMap tmB;
    uint64_t dimsB[2] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(Cout)};
    uint64_t strB[1] = {static_cast<uint64_t>(Cin) * 2};
    uint32_t boxB[2] = {LIN_BK, BN};
    if (int e = make_tmap_bf16(&tmB, wbf, 2, dimsB, strB, boxB)) return e;
    constexpr size_t smem = LIN_STAGES * (LIN_BM * LIN_BK * 2 + BN * LIN_BK * 2) + 1024;
    static DeviceOnce once;
    if (int e = smem_attr_once(once, reinterpret_cast<const void*>(linear_tc_kernel<BN>), smem, "linear smem attr")) return e;
    const int ntiles = Cout / BN, items = ((M + LIN_BM - 1) / LIN_BM) * ntiles;
    const int n_sm = sm_count();
    const int grid = items < 2 * n_sm ? items : 2 * n_sm;
    linear_tc_kernel<BN><<<grid, LIN_THREADS, smem, s>>>(tmA, tmB, bias, static_cast<__nv_bfloat16*>(y), ldy, M,
                                                         Cin / LIN_BK, ntiles, items);
    count_launch();
    return check_cuda(cudaGetLastError(), "linear_tc launch");
}

int launch_linear_bf16(const void* x, int ldx, const float* w, const float* bias, int M, int Cin, int Cout, void* y,
                       int ldy, void* ws, cudaStream_t s) {
    const size_t n = static_cast<size_t>(Cout) * Cin;
    f32_to_bf16_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(w, static_cast<__nv_bfloat16*>(ws), n);
    count_launch();
#ifndef MHADA_LINEAR_OWN_KERNEL
    // r2: the token GEMM of the ViT (gemm_tc.cu: 128 x 256 tiles, eight epilogue warps, TMA-store epilogue) also runs
    // out_conv; linear_tc_kernel stays for Cout that is not a multiple of 128 (and behind -DMHADA_LINEAR_OWN_KERNEL)
    if (Cout % 128 == 0) {
        GemmDesc g{};
        g.a = x; g.lda = ldx; g.w = ws; g.ldw = Cin; g.bias = bias; g.M = M; g.N = Cout; g.K = Cin;
        g.out_bf16 = y; g.ldo = ldy;
        return launch_gemm_bf16(g, s);
    }
#endif
    CUtensorMap tmA;
    uint64_t dimsA[2] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(M)};
    uint64_t strA[1] = {static_cast<uint64_t>(ldx) * 2};
    uint32_t boxA[2] = {LIN_BK, LIN_BM};
    if (int e = make_tmap_bf16(&tmA, x, 2, dimsA, strA, boxA)) return e;
    if (Cout % 128 == 0) return launch_linear_tc<128>(tmA, ws, bias, M, Cin, Cout, y, ldy, s);
    return launch_linear_tc<64>(tmA, ws, bias, M, Cin, Cout, y, ldy, s);
}

}  // namespace mh

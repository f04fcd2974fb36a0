// Decoder blocks 0..7 on the tensor cores: ReflectionPad2d(1) + Conv2d(Cin, Cout, 3) + bias + ReLU as an
// IMPLICIT GEMM on tcgen05, replacing `Conv` / `ConvReLU` of MHAdaSTr/network/conv.py:23-45 (used by Decoder, :75-100).
// r1 ran these eight convolutions through cuDNN with a separate reflect-pad pass in front of each one.
//
//   out[b, y, x, co] = relu( bias[co] + sum_{ky, kx, ci} W[co, ci, ky, kx] * xp[b, y + ky, x + kx, ci] )
//
// xp is the reflect-PADDED channels_last activation [B, H + 2, W + 2, Cin] (bf16).  The GEMM view: M = output pixels,
// N = Cout, K = 9 * Cin ordered (ky, kx, ci).  No im2col buffer exists: for tap (ky, kx) and a 64-channel chunk the A
// operand of a TW x TH pixel tile is the 4-D TMA box {64 channels, TW, TH, 1} at (c0, x0 + kx, y0 + ky, b) of xp -- it
// lands in shared memory as 128-byte pixel rows in exactly the K-major 128B-swizzled layout the MMA reads.
//
// Work item = MT sub-tiles of 128 pixels (a TW x (MT * 128 / TW) pixel block) x all Cout channels, MT * Cout = 256:
// the weight tile of a k-step is loaded once per 128 * MT pixels, which keeps the L2 -> shared-memory traffic per FLOP
// the same for Cout = 256, 128 and 64 (the L2 slice bandwidth, ~43 B/clk per SM, is what bounds these kernels).
//   warp 0      TMA producer (A box + weight tile per k-step, mbarrier ring that runs across work items)
//   warp 1      tcgen05.mma issuer, MT accumulators of Cout columns, DOUBLE-BUFFERED in TMEM (2 x 256 columns)
//   warps 2..9  epilogue: TMEM -> + bias -> ReLU -> bf16 -> global, two warps per TMEM lane quarter
// The epilogue can write the result ALREADY REFLECT-PADDED for the next block ([B, H + 2, W + 2, Cout]: interior pixel
// plus up to three mirrored copies for pixels next to the border), which removes the pad-only pass between two
// convolutions; blocks followed by the x2 bilinear up-sample (conv.py:61-72) write plain [B, H, W, Cout] for
// pad_reflect_kernel<UP>.
// Tensor-bound: 2 * B*H*W * Cout * 9*Cin FLOP per launch.
#include <stdlib.h>

#include "common.h"
#include "ptx.cuh"

namespace mh {

constexpr int CV_BK = 64, CV_EPI_WARPS = 8, CV_THREADS = 64 + 32 * CV_EPI_WARPS;

struct ConvParams {
    const float* bias;
    __nv_bfloat16* out;
    int B, H, W, Cin;
    int TW, TR;             // tile: TW pixels wide, TR = MT * 128 / TW rows
    int tiles_x, tiles_y, items;
    int cchunks, ktiles;    // Cin / 64, 9 * cchunks
    int relu, out_padded;
};

// CTA-pair variant: the bytes land in THIS CTA's shared memory and are counted on the LEADER's barrier (ptx.cuh)
__device__ __forceinline__ void tma_load_4d_pair(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// Epilogue shared by both kernels (warps 2..9): TMEM lane quarter = warp % 4; the 256 accumulator columns (MT sub-tiles x
// BN channels) are split in two halves between the two warps of a quarter.
// NCTA = 2 (CTA pairs, Cout = 256): work item `it` of pair `worker` = pixel tiles 2 it and 2 it + 1, one per CTA.
template <int BN, int NCTA = 1>
__device__ __forceinline__ void conv_epilogue(const ConvParams& p, uint32_t tmem, uint64_t* acc_full, uint64_t* acc_empty,
                                              int warp, int lane, int per_img, int cta_rank = 0) {
    const int quarter = warp & 3;
    const int chalf = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const int Hp = p.H + 2, Wp = p.W + 2;
    const int worker = static_cast<int>(blockIdx.x) / NCTA, nworkers = static_cast<int>(gridDim.x) / NCTA;
    int n = 0;
    for (int it = worker; it < p.items; it += nworkers, ++n) {
        const int tile = it * NCTA + cta_rank;
        const int b = tile / per_img, t = tile % per_img;       // b >= B: the odd tile out of the last pair (nothing to store)
        const int x0 = (t % p.tiles_x) * p.TW, y0 = (t / p.tiles_x) * p.TR;
        const int u = n & 1;
        mbar_wait(&acc_full[u], (n >> 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int cg = 0; cg < 128; cg += 32) {
            const int col = chalf * 128 + cg;                 // accumulator column 0..255
            const int st = col / BN, c = col % BN;            // sub-tile, channel offset
            uint32_t r[32];
            tmem_ld_x32(tmem_addr(tmem, quarter * 32, u * 256 + col), r);
            tmem_wait_ld();
            if (cg + 32 == 128) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (NCTA == 2) mbar_arrive_leader(&acc_empty[u]);
                    else mbar_arrive(&acc_empty[u]);
                }
            }
            const int pi = st * 128 + row;                    // pixel index inside the TW x TR block
            const int y = y0 + pi / p.TW, x = x0 + pi % p.TW;
            if (b < p.B && y < p.H && x < p.W) {
                uint32_t o[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float v0 = __uint_as_float(r[2 * i]) + __ldg(p.bias + c + 2 * i);
                    float v1 = __uint_as_float(r[2 * i + 1]) + __ldg(p.bias + c + 2 * i + 1);
                    if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
                    o[i] = pack_bf16x2(v0, v1);
                }
                if (!p.out_padded) {
                    __nv_bfloat16* dst = p.out + ((static_cast<size_t>(b) * p.H + y) * p.W + x) * BN + c;
                    st_global_256(dst, o);
                    st_global_256(dst + 16, o + 8);
                } else {
                    // interior position (y + 1, x + 1) of the padded map, plus the mirror images ReflectionPad2d(1)
                    // takes from this pixel: row 1 -> padded row 0, row H-2 -> padded row H+1 (both when H == 3),
                    // same for columns; a pixel next to a corner feeds up to nine positions
                    __nv_bfloat16* img = p.out + static_cast<size_t>(b) * Hp * Wp * BN + c;
                    int ys[3], xs[3], ny = 0, nx = 0;
                    ys[ny++] = y + 1;
                    if (y == 1) ys[ny++] = 0;
                    if (y == p.H - 2) ys[ny++] = p.H + 1;
                    xs[nx++] = x + 1;
                    if (x == 1) xs[nx++] = 0;
                    if (x == p.W - 2) xs[nx++] = p.W + 1;
                    for (int a = 0; a < ny; ++a)
                        for (int e = 0; e < nx; ++e) {
                            __nv_bfloat16* dst = img + (static_cast<size_t>(ys[a]) * Wp + xs[e]) * BN;
                            st_global_256(dst, o);
                            st_global_256(dst + 16, o + 8);
                        }
                }
            }
        }
    }
}

// NCTA = 2 (BN = 256 only): a CTA pair (cluster of two) shares one MMA of M = 256 pixels (tcgen05.mma.cta_group::2): each
// CTA loads the im2col box of ITS 128-pixel tile and HALF of the weight tile (32 KB per k-step instead of 48 KB), which
// also makes room for a seven-stage ring; barriers as in gemm_tc.cu (bytes of both CTAs on the leader's `full`, multicast
// commits, the epilogue warps of both CTAs arrive on the leader's `acc_empty`).
template <int BN, int STAGES, int NCTA>
__global__ void __launch_bounds__(CV_THREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const ConvParams p) {
    constexpr int MT = 256 / BN;
    static_assert(NCTA == 1 || MT == 1, "CTA pairs: one 128-pixel sub-tile per CTA");
    constexpr uint32_t A_BYTES = MT * 128 * CV_BK * 2, B_BYTES = (BN / NCTA) * CV_BK * 2, STAGE = A_BYTES + B_BYTES;
    const int cta_rank = NCTA == 2 ? static_cast<int>(cluster_ctarank()) : 0;
    const int worker = static_cast<int>(blockIdx.x) / NCTA, nworkers = static_cast<int>(gridDim.x) / NCTA;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[STAGES], empty[STAGES], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(&full[s], 1);
                mbar_init(&empty[s], 1);
            }
            for (int u = 0; u < 2; ++u) {
                mbar_init(&acc_full[u], 1);
                mbar_init(&acc_empty[u], CV_EPI_WARPS * NCTA);
            }
            fence_mbar_init();
        }
        __syncwarp();
        if (NCTA == 2) {
            tmem_alloc2(&tmem_slot, 512);
            tmem_relinquish2();
        } else {
            tmem_alloc(&tmem_slot, 512);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    if (NCTA == 2) cluster_sync_all();
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int per_img = p.tiles_x * p.tiles_y;

    if (warp == 0) {
        if (elect_one()) {
            int g = 0;
            for (int it = worker; it < p.items; it += nworkers) {
                const int tile = it * NCTA + cta_rank;
                const int b = tile / per_img, t = tile % per_img;       // b >= B (odd tile out): the box is zero-filled
                const int x0 = (t % p.tiles_x) * p.TW, y0 = (t / p.tiles_x) * p.TR;
                for (int kt = 0; kt < p.ktiles; ++kt, ++g) {
                    const int tap = kt / p.cchunks, cc = kt % p.cchunks;
                    const int s = g % STAGES;
                    mbar_wait(&empty[s], ((g / STAGES) & 1) ^ 1);
                    uint8_t* a = smem + s * STAGE;
                    if (NCTA == 2) {
                        if (cta_rank == 0) mbar_arrive_expect_tx(&full[s], 2 * STAGE);
                        tma_load_4d_pair(a, &tmX, &full[s], cc * CV_BK, x0 + tap % 3, y0 + tap / 3, b);
                        tma_load_2d_pair(a + A_BYTES, &tmW, &full[s], kt * CV_BK, cta_rank * (BN / 2));
                    } else {
                        mbar_arrive_expect_tx(&full[s], STAGE);
                        tma_load_4d(a, &tmX, &full[s], cc * CV_BK, x0 + tap % 3, y0 + tap / 3, b);
                        tma_load_2d(a + A_BYTES, &tmW, &full[s], kt * CV_BK, 0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (cta_rank == 0 && elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(128 * NCTA, BN, 0, 0);
            int g = 0, n = 0;
            for (int it = worker; it < p.items; it += nworkers, ++n) {
                const int u = n & 1;
                mbar_wait(&acc_empty[u], ((n >> 1) & 1) ^ 1);
                tc_fence_after();
                for (int kt = 0; kt < p.ktiles; ++kt, ++g) {
                    const int s = g % STAGES;
                    mbar_wait(&full[s], (g / STAGES) & 1);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + s * STAGE);
                    const uint64_t db = make_smem_desc(a_addr + A_BYTES, 16, 1024);
#pragma unroll
                    for (int t = 0; t < MT; ++t) {
                        const uint64_t da = make_smem_desc(a_addr + t * (128 * CV_BK * 2), 16, 1024);
#pragma unroll
                        for (int k = 0; k < CV_BK / 16; ++k) {
                            if (NCTA == 2) umma_ss_pair(tmem + u * 256 + t * BN, desc_advance(da, k * 32), desc_advance(db, k * 32), idesc, (kt | k) != 0);
                            else umma_ss(tmem + u * 256 + t * BN, desc_advance(da, k * 32), desc_advance(db, k * 32), idesc, (kt | k) != 0);
                        }
                    }
                    if (NCTA == 2) umma_commit_pair(&empty[s], 3);
                    else umma_commit(&empty[s]);
                }
                if (NCTA == 2) umma_commit_pair(&acc_full[u], 3);
                else umma_commit(&acc_full[u]);
            }
        }
    } else {
        conv_epilogue<BN, NCTA>(p, tmem, acc_full, acc_empty, warp, lane, per_img, cta_rank);
    }
    tc_fence_before();
    if (NCTA == 2) {
        cluster_sync_all();
        if (warp == 1) tmem_dealloc2(tmem, 512);
    } else {
        __syncthreads();
        if (warp == 1) tmem_dealloc(tmem, 512);
    }
}

// HALO variant for Cout = 64 / 128 and W > 64 (decoder blocks 4..7).  With few output channels the A operand dominates
// the L2 -> shared-memory traffic and the nine taps re-read it nine times (ncu r02: 38-41 % tensor-pipe activity at
// Cout = 64).  Here a k-step is (ky, 64-channel chunk): ONE box {64 channels, 128 + 2 pixels, MT rows} is loaded and
// serves the three kx taps -- the tap shifts the operand by kx pixel rows (128 B each) inside the swizzled tile: the
// descriptor start address simply moves by kx * 128 B.  MEASURED (r2, tests/test_gpu_stages.py::test_conv3x3_tc): the
// 128B swizzle of a K-major operand is a function of the ABSOLUTE shared-memory address (bits [4,7) ^= bits [7,10)),
// exactly what TMA wrote, so a start that is 128-byte but not 1024-byte aligned needs NO base-offset field; setting the
// descriptor's base offset to (address >> 7) & 7 applies the phase twice and gives wrong results (MHADA_CONV_HALO=2
// keeps that variant for the record).  Sub-tile = one image row of 128 pixels.
constexpr int CVH_TW = 128, CVH_ROW = CVH_TW + 2;
__host__ __device__ constexpr uint32_t cvh_a_bytes(int MT) { return (static_cast<uint32_t>(MT) * CVH_ROW * 128u + 1023u) & ~1023u; }
__device__ __forceinline__ uint64_t make_smem_desc_rowshift(uint32_t smem_addr, uint32_t sbo_bytes, int with_base_offset) {
    uint64_t d = make_smem_desc(smem_addr, 16, sbo_bytes);
    if (with_base_offset) d |= static_cast<uint64_t>((smem_addr >> 7) & 7u) << 49;
    return d;
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(CV_THREADS, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const ConvParams p,
                    const int with_base_offset) {
    constexpr int MT = 256 / BN;
    constexpr uint32_t A_BYTES = cvh_a_bytes(MT), B_BYTES = BN * CV_BK * 2, STAGE = A_BYTES + 3 * B_BYTES;
    constexpr uint32_t A_BOX_BYTES = MT * CVH_ROW * 128;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[STAGES], empty[STAGES], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(&full[s], 1);
                mbar_init(&empty[s], 1);
            }
            for (int u = 0; u < 2; ++u) {
                mbar_init(&acc_full[u], 1);
                mbar_init(&acc_empty[u], CV_EPI_WARPS);
            }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(&tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int per_img = p.tiles_x * p.tiles_y;
    const int ksteps = 3 * p.cchunks;                      // (ky, channel chunk)

    if (warp == 0) {
        if (elect_one()) {
            int g = 0;
            for (int it = blockIdx.x; it < p.items; it += gridDim.x) {
                const int b = it / per_img, t = it % per_img;
                const int x0 = (t % p.tiles_x) * p.TW, y0 = (t / p.tiles_x) * p.TR;
                for (int ks = 0; ks < ksteps; ++ks, ++g) {
                    const int ky = ks / p.cchunks, cc = ks % p.cchunks;
                    const int s = g % STAGES;
                    mbar_wait(&empty[s], ((g / STAGES) & 1) ^ 1);
                    mbar_arrive_expect_tx(&full[s], A_BOX_BYTES + 3 * B_BYTES);
                    uint8_t* a = smem + s * STAGE;
                    tma_load_4d(a, &tmX, &full[s], cc * CV_BK, x0, y0 + ky, b);
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
                        tma_load_2d(a + A_BYTES + kx * B_BYTES, &tmW, &full[s], ((ky * 3 + kx) * p.cchunks + cc) * CV_BK, 0);
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
            int g = 0, n = 0;
            for (int it = blockIdx.x; it < p.items; it += gridDim.x, ++n) {
                const int u = n & 1;
                mbar_wait(&acc_empty[u], ((n >> 1) & 1) ^ 1);
                tc_fence_after();
                for (int ks = 0; ks < ksteps; ++ks, ++g) {
                    const int s = g % STAGES;
                    mbar_wait(&full[s], (g / STAGES) & 1);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + s * STAGE);
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const uint64_t db = make_smem_desc(a_addr + A_BYTES + kx * B_BYTES, 16, 1024);
#pragma unroll
                        for (int t = 0; t < MT; ++t) {
                            // image row t of the block, shifted by kx pixels: 128 consecutive 128-byte rows
                            const uint64_t da = make_smem_desc_rowshift(a_addr + (t * CVH_ROW + kx) * 128, 1024, with_base_offset);
#pragma unroll
                            for (int k = 0; k < CV_BK / 16; ++k)
                                umma_ss(tmem + u * 256 + t * BN, desc_advance(da, k * 32), desc_advance(db, k * 32), idesc,
                                        (ks | kx | k) != 0);
                        }
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(&acc_full[u]);
            }
        }
    } else {
        conv_epilogue<BN>(p, tmem, acc_full, acc_empty, warp, lane, per_img);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

template <int BN, int STAGES>
static int launch_conv_halo(const void* xp, const void* w, const float* bias, int B, int H, int W, int Cin, int relu,
                            int out_padded, void* y, int with_base_offset, cudaStream_t s) {
    constexpr int MT = 256 / BN;
    ConvParams p;
    p.TW = CVH_TW;
    p.TR = MT;
    CUtensorMap tmX, tmW;
    {
        uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(W + 2), static_cast<uint64_t>(H + 2), static_cast<uint64_t>(B)};
        uint64_t str[3] = {static_cast<uint64_t>(Cin) * 2, static_cast<uint64_t>(W + 2) * Cin * 2,
                           static_cast<uint64_t>(H + 2) * (W + 2) * Cin * 2};
        uint32_t box[4] = {CV_BK, CVH_ROW, static_cast<uint32_t>(MT), 1};
        if (int e = make_tmap(&tmX, xp, 2, 4, dims, str, box)) return e;
    }
    {
        uint64_t dims[2] = {static_cast<uint64_t>(9) * Cin, static_cast<uint64_t>(BN)};
        uint64_t str[1] = {static_cast<uint64_t>(9) * Cin * 2};
        uint32_t box[2] = {CV_BK, BN};
        if (int e = make_tmap(&tmW, w, 2, 2, dims, str, box)) return e;
    }
    p.bias = bias; p.out = static_cast<__nv_bfloat16*>(y);
    p.B = B; p.H = H; p.W = W; p.Cin = Cin;
    p.tiles_x = (W + p.TW - 1) / p.TW; p.tiles_y = (H + p.TR - 1) / p.TR;
    p.items = B * p.tiles_x * p.tiles_y;
    p.cchunks = Cin / CV_BK; p.ktiles = 9 * p.cchunks;
    p.relu = relu; p.out_padded = out_padded;
    constexpr size_t smem = STAGES * (cvh_a_bytes(MT) + 3 * BN * CV_BK * 2) + 1024;
    static DeviceOnce once;
    if (int e = smem_attr_once(once, reinterpret_cast<const void*>(conv3x3_halo_kernel<BN, STAGES>), smem, "conv_halo smem attr")) return e;
    const int n_sm = sm_count();
    const int grid = p.items < n_sm ? p.items : n_sm;
    conv3x3_halo_kernel<BN, STAGES><<<grid, CV_THREADS, smem, s>>>(tmX, tmW, p, with_base_offset);
    count_launch();
    return check_cuda(cudaGetLastError(), "conv3x3_halo launch");
}

template <int BN, int STAGES, int NCTA = 1>
static int launch_conv_bn(const void* xp, const void* w, const float* bias, int B, int H, int W, int Cin, int relu,
                          int out_padded, void* y, cudaStream_t s) {
    constexpr int MT = 256 / BN;
    ConvParams p;
    // tile shape: as wide as the image allows (TMA boxes are rows of pixels), MT * 128 pixels in all
    int TW = 128;
    while (TW > 8 && TW / 2 >= W) TW /= 2;
    if (TW > 128) TW = 128;
    p.TW = TW;
    p.TR = MT * 128 / TW;
    if (p.TR > 256) {                        // TMA box limit: 256 per dimension
        set_error("conv3x3_tc: image too narrow for the %d-pixel tile (W = %d)", MT * 128, W);
        return MHADA_ERR_UNSUPPORTED;
    }
    CUtensorMap tmX, tmW;
    {
        uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(W + 2), static_cast<uint64_t>(H + 2), static_cast<uint64_t>(B)};
        uint64_t str[3] = {static_cast<uint64_t>(Cin) * 2, static_cast<uint64_t>(W + 2) * Cin * 2,
                           static_cast<uint64_t>(H + 2) * (W + 2) * Cin * 2};
        uint32_t box[4] = {CV_BK, static_cast<uint32_t>(p.TW), static_cast<uint32_t>(p.TR), 1};
        if (int e = make_tmap(&tmX, xp, 2, 4, dims, str, box)) return e;
    }
    {
        uint64_t dims[2] = {static_cast<uint64_t>(9) * Cin, static_cast<uint64_t>(BN)};
        uint64_t str[1] = {static_cast<uint64_t>(9) * Cin * 2};
        uint32_t box[2] = {CV_BK, BN / NCTA};
        if (int e = make_tmap(&tmW, w, 2, 2, dims, str, box)) return e;
    }
    p.bias = bias; p.out = static_cast<__nv_bfloat16*>(y);
    p.B = B; p.H = H; p.W = W; p.Cin = Cin;
    p.tiles_x = (W + p.TW - 1) / p.TW; p.tiles_y = (H + p.TR - 1) / p.TR;
    p.items = (B * p.tiles_x * p.tiles_y + NCTA - 1) / NCTA;          // pair kernel: two pixel tiles per work item
    p.cchunks = Cin / CV_BK; p.ktiles = 9 * p.cchunks;
    p.relu = relu; p.out_padded = out_padded;
    constexpr size_t smem = STAGES * (MT * 128 * CV_BK * 2 + (BN / NCTA) * CV_BK * 2) + 1024;
    static DeviceOnce once;
    if (int e = smem_attr_once(once, reinterpret_cast<const void*>(conv3x3_tc_kernel<BN, STAGES, NCTA>), smem, "conv_tc smem attr")) return e;
    const int n_workers = sm_count() / NCTA;
    const int grid = (p.items < n_workers ? p.items : n_workers) * NCTA;
    if (NCTA == 1) {
        conv3x3_tc_kernel<BN, STAGES, NCTA><<<grid, CV_THREADS, smem, s>>>(tmX, tmW, p);
    } else {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(static_cast<unsigned>(grid));
        cfg.blockDim = dim3(CV_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = NCTA; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (int e = check_cuda(cudaLaunchKernelEx(&cfg, conv3x3_tc_kernel<BN, STAGES, NCTA>, tmX, tmW, p), "conv3x3_tc pair launch")) return e;
    }
    count_launch();
    return check_cuda(cudaGetLastError(), "conv3x3_tc launch");
}

int launch_conv3x3_tc(const void* xp, const void* w, const float* bias, int B, int H, int W, int Cin, int Cout, int relu,
                      int out_padded, void* y, cudaStream_t s) {
    if (Cin % CV_BK != 0 || (Cout != 64 && Cout != 128 && Cout != 256)) {
        set_error("conv3x3_tc: implemented for Cin %% 64 == 0 and Cout in {64, 128, 256}, got %d -> %d", Cin, Cout);
        return MHADA_ERR_UNSUPPORTED;
    }
    // MHADA_CONV_HALO (development switch): 0 = halo variant off, 1 = on (default), 2 = on WITH the descriptor
    // base-offset field set (measured wrong, see conv3x3_halo_kernel)
    static const int halo_mode = [] {
        const char* e = getenv("MHADA_CONV_HALO");
        return e ? atoi(e) : 1;
    }();
    if (halo_mode && W > 64 && Cout == 128) return launch_conv_halo<128, 2>(xp, w, bias, B, H, W, Cin, relu, out_padded, y, halo_mode == 2, s);
    if (halo_mode && W > 64 && Cout == 64) return launch_conv_halo<64, 2>(xp, w, bias, B, H, W, Cin, relu, out_padded, y, halo_mode == 2, s);
    if (Cout == 256) {
        // MHADA_CONV_PAIR=0 keeps the single-CTA kernel (A/B switch); pairs need at least one pair of 128-pixel tiles
        static const bool pair_ok = [] { const char* e = getenv("MHADA_CONV_PAIR"); return !(e && e[0] == '0'); }();
        if (pair_ok && static_cast<long long>(B) * H * W >= 256) return launch_conv_bn<256, 7, 2>(xp, w, bias, B, H, W, Cin, relu, out_padded, y, s);
        return launch_conv_bn<256, 4>(xp, w, bias, B, H, W, Cin, relu, out_padded, y, s);
    }
    if (Cout == 128) return launch_conv_bn<128, 4>(xp, w, bias, B, H, W, Cin, relu, out_padded, y, s);
    return launch_conv_bn<64, 3>(xp, w, bias, B, H, W, Cin, relu, out_padded, y, s);
}

}  // namespace mh

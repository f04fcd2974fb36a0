// Token GEMM on the tensor cores for the ViT encoder:
//     y[M, N] = act( x[M, K] . W[N, K]^T + bias (+ resid[row(m), N]) )        bf16 operands, f32 accumulate
//
// Replaces, in MHAdaSTr/network/vit.py: PatchEmbedding.conv_proj as a GEMM over 8x8x3 patches (:105-117, + the
// positional table :96-102 as a row-periodic residual), the in_proj / out_proj of nn.MultiheadAttention (:48,59),
// the two nn.Linear of the MLP with the ReLU between them (:49-53,62-63) and both residual additions (:60,64).
//
// Structure: persistent CTAs (one per SM), 10 warps; a work item is one 128 x BN output tile, the BN tiles of the
// same 128 rows are consecutive items (x comes from HBM once and from L2 for the other column tiles):
//   warp 0      TMA producer: x and W tiles -> 128B-swizzled shared memory, STAGES-deep mbarrier ring that runs
//               straight across work items
//   warp 1      tcgen05.mma issuer (one elected lane), M = 128, N = BN, accumulators DOUBLE-BUFFERED in TMEM
//               (2 x BN columns): the MMAs of item n+1 overlap the epilogue of item n
//   warps 2..9  epilogue, TWO warps per TMEM lane quarter, each owning half of the tile's columns (a 128 x 256 tile has
//               to leave TMEM in less than its ~2 us of MMAs): tcgen05.ld (lane = row) -> + bias (+ residual, the next
//               32 columns prefetched while the current ones are processed) (ReLU) ->
//                 bf16 result: packed into a 128B-swizzled staging slab (32 rows x 64 columns per warp, two slabs)
//                              and written with ONE TMA STORE per slab (cp.async.bulk.tensor, rows past M clipped);
//                 f32 result (the residual stream): the residual tile comes IN by TMA (one 32 x 32 tile ahead), the sum
//                              goes OUT by TMA store from the same slab (thread-per-row global accesses made these
//                              epilogues L1-tag-bound).
// BN = 256 keeps the shared-memory operand traffic of the SS-mode MMAs at 96 B/clk (a 128 x 128 tile needs the
// full 128 B/clk of the SM).  Compute-bound for K >= 512: 2*M*N*K FLOP; HBM bytes 2*M*K + 2*N*K + (2|4|6)*M*N.
#include <cstdlib>

#include "common.h"
#include "ptx.cuh"

namespace mh {

constexpr int GM_BM = 128, GM_BK = 64, GM_EPI_WARPS = 8, GM_THREADS = 64 + 32 * GM_EPI_WARPS;
constexpr uint32_t GM_STG_WARP = 2 * 32 * 128;   // two staging slabs (32 rows x 128 B) per epilogue warp

struct GemmParams {
    const float* bias;     // [N] or nullptr
    const float* resid;    // f32 [*, ldr] or nullptr
    float* out_f32;        // f32 [M, ldf] or nullptr
    __nv_bfloat16* out_bf16;   // bf16 [M, ldo] (direct stores when an f32 result is written too)
    int ldo;
    int ldr, resid_mod;    // residual row = m % resid_mod when resid_mod > 0 (positional table), else m
    int ldf;
    int has_bf16;          // bf16 result through the tensor map tmC
    int resid_tma;         // residual tiles come through tmR (row-periodic tables only when resid_mod % 32 == 0)
    int relu;
    int M, ktiles, ntiles, items;
    int tiles_mn;          // output tiles of one K split (= items without split-K)
    int split_rows;        // split-K: partial s is written at row s * split_rows of the f32 output (a multiple of 128)
};

#ifdef MHADA_GEMM_TRACE
// development build only (-DMHADA_GEMM_TRACE, tools/trace_gemm.py): per CTA, cycles the three roles spend waiting
__device__ unsigned long long mh_gemm_trace[512 * 8];
#define GT_BEGIN(v) const long long v = clock64()
#define GT_ADD(slot, v) gt[slot] += clock64() - v
#else
#define GT_BEGIN(v)
#define GT_ADD(slot, v)
#endif

// v[i] = acc[i] + bias[col + i], i < 32: eight 16-byte loads (the same for every lane: L1 broadcast) instead of 32 scalar
// ones; the launcher guarantees a 16-byte aligned bias pointer, col is a multiple of 32
__device__ __forceinline__ void add_bias32(float (&v)[32], const uint32_t (&r)[32], const float* __restrict__ bias, int col) {
    if (bias) {
        const float4* b4 = reinterpret_cast<const float4*>(bias + col);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 b = __ldg(b4 + q);
            v[4 * q] = __uint_as_float(r[4 * q]) + b.x; v[4 * q + 1] = __uint_as_float(r[4 * q + 1]) + b.y;
            v[4 * q + 2] = __uint_as_float(r[4 * q + 2]) + b.z; v[4 * q + 3] = __uint_as_float(r[4 * q + 3]) + b.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
    }
}

// NCTA = 2: a CTA PAIR (cluster of two, the two SMs of a TPC) works on one 256 x BN tile with tcgen05.mma.cta_group::2:
// each CTA loads its 128 rows of x and HALF of the W tile (32 KB per k-step instead of 48 KB: the L2 -> SM operand
// stream was what bounded the single-CTA kernel), the leader (cluster rank 0) issues the M = 256 MMAs for
// both, each CTA's TMEM receives its 128 x BN slice and its own epilogue warps drain it.  TMA bytes of both CTAs are
// counted on the leader's `full` barrier; tcgen05.commit multicasts `empty` / `acc_full` to both CTAs; the epilogue
// warps of both CTAs arrive on the leader's `acc_empty`.
template <int BN, int STAGES, int NCTA>
__global__ void __launch_bounds__(GM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmF,
               const __grid_constant__ CUtensorMap tmR, const GemmParams p) {
    constexpr uint32_t A_BYTES = GM_BM * GM_BK * 2, B_BYTES = (BN / NCTA) * GM_BK * 2, STAGE = A_BYTES + B_BYTES;
    constexpr int TM = GM_BM * NCTA;                       // output rows of a work item
    const int cta_rank = NCTA == 2 ? static_cast<int>(cluster_ctarank()) : 0;
    const int worker = static_cast<int>(blockIdx.x) / NCTA, nworkers = static_cast<int>(gridDim.x) / NCTA;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[STAGES], empty[STAGES], acc_full[2], acc_empty[2];
    __shared__ uint64_t rbar[GM_EPI_WARPS][2];       // residual tile landed in slab s of epilogue warp w
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (p.has_bf16) tma_prefetch_desc(&tmC);
        if (p.out_f32) tma_prefetch_desc(&tmF);
        if (p.resid_tma) tma_prefetch_desc(&tmR);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(&full[s], 1);
                mbar_init(&empty[s], 1);
            }
            for (int u = 0; u < 2; ++u) {
                mbar_init(&acc_full[u], 1);
                mbar_init(&acc_empty[u], GM_EPI_WARPS * NCTA);   // one arrive per epilogue warp (of both CTAs of a pair)
            }
            for (int w = 0; w < GM_EPI_WARPS; ++w) {
                mbar_init(&rbar[w][0], 1);
                mbar_init(&rbar[w][1], 1);
            }
            fence_mbar_init();
        }
        __syncwarp();
        if (NCTA == 2) {
            tmem_alloc2(&tmem_slot, 2 * BN);
            tmem_relinquish2();
        } else {
            tmem_alloc(&tmem_slot, 2 * BN);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    if (NCTA == 2) cluster_sync_all();     // the peer's barriers are initialised before anything arrives on them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
#ifdef MHADA_GEMM_TRACE
            long long gt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
            int g = 0;                                            // running k-tile counter of this CTA
            for (int it = worker; it < p.items; it += nworkers) {
                const int sp = it / p.tiles_mn, r = it - sp * p.tiles_mn;       // K split, tile within the split
                const int m0 = (r / p.ntiles) * TM + cta_rank * GM_BM, n0 = (r % p.ntiles) * BN + cta_rank * (BN / NCTA);
                const int k0 = sp * p.ktiles * GM_BK;
                for (int kt = 0; kt < p.ktiles; ++kt, ++g) {
                    const int s = g % STAGES;
                    GT_BEGIN(t0);
                    mbar_wait(&empty[s], ((g / STAGES) & 1) ^ 1);
                    GT_ADD(3, t0);
                    uint8_t* a = smem + s * STAGE;
                    if (NCTA == 2) {
                        if (cta_rank == 0) mbar_arrive_expect_tx(&full[s], 2 * STAGE);      // the pair's bytes, one barrier
                        tma_load_2d_pair(a, &tmA, &full[s], k0 + kt * GM_BK, m0);
                        tma_load_2d_pair(a + A_BYTES, &tmB, &full[s], k0 + kt * GM_BK, n0);
                    } else {
                        mbar_arrive_expect_tx(&full[s], STAGE);
                        tma_load_2d(a, &tmA, &full[s], k0 + kt * GM_BK, m0);
                        tma_load_2d(a + A_BYTES, &tmB, &full[s], k0 + kt * GM_BK, n0);
                    }
                }
            }
#ifdef MHADA_GEMM_TRACE
            mh_gemm_trace[blockIdx.x * 8 + 3] = gt[3];
#endif
        }
    } else if (warp == 1) {
        if (cta_rank == 0 && elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(TM, BN, 0, 0);
#ifdef MHADA_GEMM_TRACE
            long long gt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            const long long tstart = clock64();
#endif
            int g = 0, n = 0;
            for (int it = worker; it < p.items; it += nworkers, ++n) {
                const int u = n & 1;
                GT_BEGIN(t1);
                mbar_wait(&acc_empty[u], ((n >> 1) & 1) ^ 1);     // the epilogue has drained this accumulator buffer
                GT_ADD(1, t1);
                tc_fence_after();
                for (int kt = 0; kt < p.ktiles; ++kt, ++g) {
                    const int s = g % STAGES;
                    GT_BEGIN(t2);
                    mbar_wait(&full[s], (g / STAGES) & 1);
                    GT_ADD(2, t2);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + s * STAGE);
                    const uint64_t da = make_smem_desc(a_addr, 16, 1024);
                    const uint64_t db = make_smem_desc(a_addr + A_BYTES, 16, 1024);
#pragma unroll
                    for (int k = 0; k < GM_BK / 16; ++k) {
                        if (NCTA == 2) umma_ss_pair(tmem + u * BN, desc_advance(da, k * 32), desc_advance(db, k * 32), idesc, (kt | k) != 0);
                        else umma_ss(tmem + u * BN, desc_advance(da, k * 32), desc_advance(db, k * 32), idesc, (kt | k) != 0);
                    }
                    if (NCTA == 2) umma_commit_pair(&empty[s], 3);
                    else umma_commit(&empty[s]);
                }
                if (NCTA == 2) umma_commit_pair(&acc_full[u], 3);
                else umma_commit(&acc_full[u]);
            }
#ifdef MHADA_GEMM_TRACE
            mh_gemm_trace[blockIdx.x * 8 + 0] = clock64() - tstart;
            mh_gemm_trace[blockIdx.x * 8 + 1] = gt[1];
            mh_gemm_trace[blockIdx.x * 8 + 2] = gt[2];
            mh_gemm_trace[blockIdx.x * 8 + 7] = n;
#endif
        }
    } else {
        // epilogue warps 2..9: TMEM lane quarter = warp % 4 (hardware rule), column half = (warp - 2) / 4
        const int quarter = warp & 3;
        const int chalf = (warp - 2) >> 2;
        constexpr int CW = BN / 2;                                             // columns per epilogue warp
        const int row = quarter * 32 + lane;
        uint8_t* stg = smem + STAGES * STAGE + (warp - 2) * GM_STG_WARP;      // 1024-byte aligned slabs
        const uint32_t swz = static_cast<uint32_t>(lane & 7);
        int nstore = 0;                                                        // TMA stores issued by this warp
        if (p.out_f32) {
            // ---- f32 result (the residual stream): BOTH directions go through the staging slabs with TMA.  A
            // thread-per-row access touches 32 different 128-byte lines per warp instruction: the residual loads and
            // 256-bit stores of a 128 x 256 tile cost ~13 k L1 tag cycles (6.8 us against 2.1 us of MMAs; ncu r02:
            // long-scoreboard + LSU-throttle stalls, 58 us for out_proj against 24 us for the bf16-only out_conv).
            // The residual tile of the NEXT 32-column chunk (or of the next work item's first chunk) is in flight
            // while the current one is processed; the sum is written back into the same slab and stored from there.
            uint64_t* rb = rbar[warp - 2];
            constexpr int NCH = CW / 32;
            const int my_items = (p.items - worker + nworkers - 1) / nworkers;
            const int total = my_items * NCH;                                   // chunks this warp will process
            auto coords = [&](int k, int& col, int& rowbase) {                   // chunk k -> output column / row of its tile
                const int it = worker + (k / NCH) * nworkers;
                col = (it % p.ntiles) * BN + chalf * CW + (k % NCH) * 32;
                rowbase = (it / p.ntiles) * TM + cta_rank * GM_BM + quarter * 32;
            };
            auto issue_resid = [&](int k) {                                      // lane 0 only
                int col, rowbase;
                coords(k, col, rowbase);
                uint8_t* slab = stg + (k & 1) * 4096;
                mbar_arrive_expect_tx(&rb[k & 1], 4096);
                tma_load_2d(slab, &tmR, &rb[k & 1], col, p.resid_mod > 0 ? rowbase % p.resid_mod : rowbase);
            };
            if (p.resid_tma && total > 0 && lane == 0) issue_resid(0);
            int k = 0;
            for (int it = worker; it < p.items; it += nworkers) {
                const int n = k / NCH;
                const int sp = it / p.tiles_mn, rt = it - sp * p.tiles_mn;
                const int m0 = (rt / p.ntiles) * TM + cta_rank * GM_BM, n0 = (rt % p.ntiles) * BN + chalf * CW;
                const int m = m0 + row;
                const bool row_ok = m < p.M;
                const int u = n & 1;
                const float* rrow = nullptr;                                     // direct residual loads (table rows that wrap)
                if (p.resid && !p.resid_tma && row_ok)
                    rrow = p.resid + static_cast<size_t>(p.resid_mod > 0 ? m % p.resid_mod : m) * p.ldr + n0;
                mbar_wait(&acc_full[u], (n >> 1) & 1);
                tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < CW; c += 32, ++k) {
                    uint8_t* slab = stg + (k & 1) * 4096;
                    // the other slab was the source of the previous chunk's store: once that store has read it, the
                    // next chunk's residual may land there
                    if (lane == 0) {
                        tma_store_wait_read<0>();
                        if (p.resid_tma && k + 1 < total) issue_resid(k + 1);
                    }
                    uint32_t r[32];
                    tmem_ld_x32(tmem_addr(tmem, quarter * 32, u * BN + chalf * CW + c), r);
                    tmem_wait_ld();
                    if (c + 32 == CW) {                       // accumulator drained: the next-but-one item may start
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (NCTA == 2) mbar_arrive_leader(&acc_empty[u]);
                            else mbar_arrive(&acc_empty[u]);
                        }
                    }
                    float v[32];
                    add_bias32(v, r, p.bias, n0 + c);
                    if (p.resid_tma) {
                        mbar_wait(&rb[k & 1], (k >> 1) & 1);
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 x = *reinterpret_cast<const float4*>(slab + lane * 128 + ((static_cast<uint32_t>(q) ^ swz) << 4));
                            v[4 * q] += x.x; v[4 * q + 1] += x.y; v[4 * q + 2] += x.z; v[4 * q + 3] += x.w;
                        }
                    } else if (rrow) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 x = __ldg(reinterpret_cast<const float4*>(rrow + c) + q);
                            v[4 * q] += x.x; v[4 * q + 1] += x.y; v[4 * q + 2] += x.z; v[4 * q + 3] += x.w;
                        }
                    }
                    if (p.relu) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                    }
                    __syncwarp();                             // every lane has read its residual row (and lane 0 has waited)
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        *reinterpret_cast<float4*>(slab + lane * 128 + ((static_cast<uint32_t>(q) ^ swz) << 4)) =
                            make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmF, slab, n0 + c, sp * p.split_rows + m0 + quarter * 32);
                        tma_store_commit();
                    }
                    if (p.has_bf16 && row_ok) {               // the rounded copy (fc2 -> feature map): 64 bytes per lane
                        uint32_t o[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) o[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
                        __nv_bfloat16* dst = p.out_bf16 + static_cast<size_t>(m) * p.ldo + n0 + c;
                        st_global_256(dst, o);
                        st_global_256(dst + 16, o + 8);
                    }
                }
            }
        } else {
        int n = 0;
#ifdef MHADA_GEMM_TRACE
        long long gt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const long long tstart = clock64();
#endif
        for (int it = worker; it < p.items; it += nworkers, ++n) {
            const int m0 = (it / p.ntiles) * TM + cta_rank * GM_BM, n0 = (it % p.ntiles) * BN + chalf * CW;
            const int u = n & 1;
            GT_BEGIN(t5);
            mbar_wait(&acc_full[u], (n >> 1) & 1);
            GT_ADD(5, t5);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < CW; c += 32) {
                uint32_t r[32];
                tmem_ld_x32(tmem_addr(tmem, quarter * 32, u * BN + chalf * CW + c), r);
                tmem_wait_ld();
                if (c + 32 == CW) {                           // accumulator drained: the next-but-one item may start
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (NCTA == 2) mbar_arrive_leader(&acc_empty[u]);
                        else mbar_arrive(&acc_empty[u]);
                    }
                }
                float v[32];
                add_bias32(v, r, p.bias, n0 + c);
                if (p.relu) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                }
                {
                    const int half = (c >> 5) & 1;            // which 32-column half of the 64-column slab
                    uint8_t* slab = stg + (nstore & 1) * (32 * 128);
                    if (half == 0) {
                        // the slab was the source of the store issued two stores ago: wait until it has been read
                        GT_BEGIN(t6);
                        if (lane == 0) tma_store_wait_read<1>();
                        __syncwarp();
                        GT_ADD(6, t6);
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 w;
                        w.x = pack_bf16x2(v[8 * q], v[8 * q + 1]);
                        w.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
                        w.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
                        w.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
                        const uint32_t chunk = static_cast<uint32_t>(half * 4 + q) ^ swz;     // 128B swizzle
                        *reinterpret_cast<uint4*>(slab + lane * 128 + chunk * 16) = w;
                    }
                    if (half == 1) {
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&tmC, slab, n0 + c - 32, m0 + quarter * 32);
                            tma_store_commit();
                        }
                        ++nstore;
                    }
                }
            }
        }
#ifdef MHADA_GEMM_TRACE
        if (warp == 2 && lane == 0) {
            mh_gemm_trace[blockIdx.x * 8 + 4] = clock64() - tstart;
            mh_gemm_trace[blockIdx.x * 8 + 5] = gt[5];
            mh_gemm_trace[blockIdx.x * 8 + 6] = gt[6];
        }
#endif
        }
        if (lane == 0) tma_store_wait<0>();
    }
    tc_fence_before();
    if (NCTA == 2) {
        cluster_sync_all();                 // neither CTA leaves (or frees TMEM) while its peer may still use it
        if (warp == 1) tmem_dealloc2(tmem, 2 * BN);
    } else {
        __syncthreads();
        if (warp == 1) tmem_dealloc(tmem, 2 * BN);
    }
}

template <int BN, int STAGES, int NCTA>
static int launch_gemm_bn(const GemmDesc& d, int ksplit, cudaStream_t s) {
    constexpr int TM = GM_BM * NCTA;
    const int Mp = (d.M + TM - 1) / TM * TM;
    CUtensorMap tmA, tmB, tmC;
    {
        uint64_t dims[2] = {static_cast<uint64_t>(d.K), static_cast<uint64_t>(d.M)};
        uint64_t str[1] = {static_cast<uint64_t>(d.lda) * 2};
        uint32_t box[2] = {GM_BK, GM_BM};
        if (int e = make_tmap_bf16(&tmA, d.a, 2, dims, str, box)) return e;
    }
    {
        uint64_t dims[2] = {static_cast<uint64_t>(d.K), static_cast<uint64_t>(d.N)};
        uint64_t str[1] = {static_cast<uint64_t>(d.ldw) * 2};
        uint32_t box[2] = {GM_BK, BN / NCTA};
        if (int e = make_tmap_bf16(&tmB, d.w, 2, dims, str, box)) return e;
    }
    if (d.out_bf16 && !d.out_f32) {
        uint64_t dims[2] = {static_cast<uint64_t>(d.N), static_cast<uint64_t>(d.M)};
        uint64_t str[1] = {static_cast<uint64_t>(d.ldo) * 2};
        uint32_t box[2] = {64, 32};
        if (int e = make_tmap_bf16(&tmC, d.out_bf16, 2, dims, str, box)) return e;
    } else {
        tmC = tmA;      // never dereferenced
    }
    CUtensorMap tmF = tmA, tmR = tmA;
    if (d.out_f32) {
        // split-K: the partial results of split s occupy rows [s Mp, s Mp + M) of the output
        uint64_t dims[2] = {static_cast<uint64_t>(d.N), static_cast<uint64_t>(ksplit > 1 ? (ksplit - 1) * Mp + d.M : d.M)};
        uint64_t str[1] = {static_cast<uint64_t>(d.ldf) * 4};
        uint32_t box[2] = {32, 32};
        if (int e = make_tmap(&tmF, d.out_f32, 4, 2, dims, str, box)) return e;
    }
    // residual tiles by TMA when the 32-row slabs never wrap around a row-periodic table
    const bool resid_tma = d.out_f32 && d.resid && (d.resid_mod == 0 || d.resid_mod % 32 == 0) && d.ldr % 4 == 0;
    if (resid_tma) {
        uint64_t dims[2] = {static_cast<uint64_t>(d.N), static_cast<uint64_t>(d.resid_mod > 0 ? d.resid_mod : d.M)};
        uint64_t str[1] = {static_cast<uint64_t>(d.ldr) * 4};
        uint32_t box[2] = {32, 32};
        if (int e = make_tmap(&tmR, d.resid, 4, 2, dims, str, box)) return e;
    }
    GemmParams p;
    p.out_bf16 = static_cast<__nv_bfloat16*>(d.out_bf16); p.ldo = d.ldo; p.resid_tma = resid_tma ? 1 : 0;
    p.bias = d.bias; p.resid = d.resid; p.out_f32 = d.out_f32;
    p.ldr = d.ldr; p.resid_mod = d.resid_mod; p.ldf = d.ldf;
    p.has_bf16 = d.out_bf16 ? 1 : 0; p.relu = d.relu;
    p.M = d.M; p.ktiles = d.K / GM_BK / ksplit; p.ntiles = d.N / BN;
    p.tiles_mn = ((d.M + TM - 1) / TM) * p.ntiles;
    p.items = p.tiles_mn * ksplit;
    p.split_rows = Mp;
    constexpr size_t smem = STAGES * (GM_BM * GM_BK * 2 + (BN / NCTA) * GM_BK * 2) + GM_EPI_WARPS * GM_STG_WARP + 1024;
    static DeviceOnce once;
    if (int e = smem_attr_once(once, reinterpret_cast<const void*>(gemm_tc_kernel<BN, STAGES, NCTA>), smem, "gemm smem attr")) return e;
    const int n_workers = sm_count() / NCTA;
    const int grid = (p.items < n_workers ? p.items : n_workers) * NCTA;
    if (NCTA == 1) {
        gemm_tc_kernel<BN, STAGES, NCTA><<<grid, GM_THREADS, smem, s>>>(tmA, tmB, tmC, tmF, tmR, p);
    } else {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(static_cast<unsigned>(grid));
        cfg.blockDim = dim3(GM_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = NCTA; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (int e = check_cuda(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, STAGES, NCTA>, tmA, tmB, tmC, tmF, tmR, p), "gemm_tc pair launch"))
            return e;
    }
    count_launch();
    return check_cuda(cudaGetLastError(), "gemm_tc launch");
}

int launch_gemm_bf16(const GemmDesc& d, cudaStream_t s) {
    if (d.M <= 0 || d.N <= 0 || d.K <= 0 || d.K % GM_BK != 0 || d.N % 128 != 0) {
        set_error("gemm_tc: needs K %% 64 == 0 and N %% 128 == 0, got M=%d N=%d K=%d", d.M, d.N, d.K);
        return MHADA_ERR_UNSUPPORTED;
    }
    if (d.bias && (reinterpret_cast<uintptr_t>(d.bias) & 15) != 0) {
        set_error("gemm_tc: bias must be 16-byte aligned");
        return MHADA_ERR_ARG;
    }
    if (d.N % 256 == 0) {
        // CTA pairs (256 x 256 tiles) when there is at least one full round of pair tiles; MHADA_GEMM_PAIR=0 keeps the
        // single-CTA kernel (A/B switch for measurements)
        static const bool pair_ok = [] { const char* e = getenv("MHADA_GEMM_PAIR"); return !(e && e[0] == '0'); }();
        if (pair_ok && d.M >= 256 * (sm_count() / 2) / (d.N / 256)) return launch_gemm_bn<256, 5, 2>(d, 1, s);
        return launch_gemm_bn<256, 3, 1>(d, 1, s);                   // 3 x 48 KB operand ring + 64 KB of staging slabs
    }
    return launch_gemm_bn<128, 4, 1>(d, 1, s);
}

// ---- split-K (the weight-gradient GEMMs of the training path: C x C outputs, K = all tokens of the batch) ----------
// A 512 x 512 x 8192 GEMM is 8 output tiles: 8 of 148 SMs busy.  The K range is cut in `ksplit` slices that run as
// independent work items of the SAME persistent kernel (the k coordinate of the TMA loads is offset by the slice), each
// writing its partial tile to its own rows of a scratch buffer; splitk_reduce_kernel adds them in a fixed order.
static int pick_ksplit(int M, int N, int K) {
    const int bn = N % 256 == 0 ? 256 : 128;
    const int tiles = ((M + GM_BM - 1) / GM_BM) * (N / bn);
    const int ktiles = K / GM_BK;
    int best = 1;
    for (int s = 2; s <= 32; ++s) {
        if (ktiles % s != 0 || ktiles / s < 4) continue;
        if (tiles * s <= sm_count()) best = s;           // one round of work items: more splits only add partial-tile traffic
    }
    return best;
}

__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ part, int splits, size_t split_stride, int M,
                                                            int N, int ldp, float* __restrict__ out, int ldo) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    const int nv = N / 4;
    if (i >= static_cast<size_t>(M) * nv) return;
    const int m = static_cast<int>(i / nv), n = static_cast<int>(i % nv) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < splits; ++s) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(part + s * split_stride + static_cast<size_t>(m) * ldp + n));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(out + static_cast<size_t>(m) * ldo + n) = acc;
}

size_t gemm_splitk_workspace(int M, int N, int K) {
    if (M <= 0 || N <= 0 || K <= 0 || K % GM_BK != 0 || N % 128 != 0) return 0;
    const int ks = pick_ksplit(M, N, K);
    if (ks <= 1) return 0;
    const size_t Mp = static_cast<size_t>(M + GM_BM - 1) / GM_BM * GM_BM;
    return ks * Mp * N * sizeof(float);
}

// f32 result only, no bias / residual / ReLU: out_f32 [M, ldf] = a . w^T with the K range split when that fills the GPU
int launch_gemm_bf16_splitk(const GemmDesc& d, void* ws, size_t ws_bytes, cudaStream_t s) {
    if (d.M <= 0 || d.N <= 0 || d.K <= 0 || d.K % GM_BK != 0 || d.N % 128 != 0 || !d.out_f32 || d.out_bf16 || d.bias || d.resid || d.relu) {
        set_error("gemm_tc split-K: f32 result only, K %% 64 == 0, N %% 128 == 0 (M=%d N=%d K=%d)", d.M, d.N, d.K);
        return MHADA_ERR_UNSUPPORTED;
    }
    const int ks = pick_ksplit(d.M, d.N, d.K);
    if (ks <= 1 || !ws || ws_bytes < gemm_splitk_workspace(d.M, d.N, d.K)) return launch_gemm_bf16(d, s);
    GemmDesc p = d;
    p.out_f32 = static_cast<float*>(ws);
    p.ldf = d.N;
    if (int e = d.N % 256 == 0 ? launch_gemm_bn<256, 3, 1>(p, ks, s) : launch_gemm_bn<128, 4, 1>(p, ks, s)) return e;
    const size_t Mp = static_cast<size_t>(d.M + GM_BM - 1) / GM_BM * GM_BM;
    const size_t n = static_cast<size_t>(d.M) * (d.N / 4);
    splitk_reduce_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(static_cast<const float*>(ws), ks, Mp * d.N, d.M, d.N,
                                                                                d.N, d.out_f32, d.ldf);
    count_launch();
    return check_cuda(cudaGetLastError(), "splitk_reduce launch");
}

}  // namespace mh

#ifdef MHADA_GEMM_TRACE
extern "C" __attribute__((visibility("default"))) int mhada_debug_gemm_trace(unsigned long long* host_out, int n_ctas) {
    return static_cast<int>(cudaMemcpyFromSymbol(host_out, mh::mh_gemm_trace, static_cast<size_t>(n_ctas) * 8 * sizeof(unsigned long long)));
}
#endif

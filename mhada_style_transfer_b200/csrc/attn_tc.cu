// K2 + K3 -- streaming MHAda attention on tcgen05 / TMEM / TMA with the AdaIN-style epilogue fused.
//
// Replaces MHAdaSTr/network/adaDecoder.py:186-198 for one head:
//     A = softmax(Q K^T)   (no 1/sqrt(d); Q arrives pre-multiplied by log2 e -> exp2)
//     M = A V ;  Var = A V^2 - M^2 ;  S = sqrt(clamp(Var, 1e-6)) ;  out = S * IN(fcs) + M
// The Nc x Ns map never leaves the SM: S tiles live in TMEM, P overwrites S in place as bf16 and
// is consumed straight from TMEM by the second MMA, which multiplies by V' = [V~ | V~^2] (128 wide)
// so the mean and the second moment come out of ONE pass (V~ = V - mu_v, added back at the end).
//
// CTA = 2 query tiles of 128 rows x 1 head, key tiles of 64, 16 warps:
//   warp 0       TMA producer: Q (once), K ring (4 x 8 KB), V' ring (4 x 16 KB), 128B swizzle
//   warp 1, 2    tcgen05.mma issuers, one elected lane each: warp 1 for query tile 0, warp 2 for query tile 1
//                (warp 2 also allocates the 512 TMEM columns; per query tile t: S buffers at t*256 + {0, 64},
//                O at t*256 + 128)
//   warp 3       idle after set-up
//   warps 4-7    softmax for query tile 0 (thread = row; TMEM lane = row, no shuffles)
//   warps 8-11   same for query tile 1
//   warps 12-15  epilogue of both query tiles (reads O out of TMEM while the other warps are in the next work item)
// Pipeline (measured motivation in DESIGN.md, attention section): S is DOUBLE-BUFFERED in TMEM, so
// S(j+1), S(j+2) are computed while the softmax warps still work on tile j -- they do not wait for the
// tensor core in steady state, and the exp2 (MUFU) units, which bound this head_dim-64 problem, stay busy.
// MMA issue order per key tile j of query tile t:  [P_t(j) ready]  PV_t(j)  S_t(j+2).
// The softmax reference (row maximum) LAGS: exact for the first key tile of a work item, afterwards it only
// moves when a tile maximum exceeds it by 2^64 (checked before P leaves the registers; then O is rescaled in
// place in TMEM and the tile redone).  The per-tile maximum therefore is off the critical path.
//
// Tensor-bound: algorithmic FLOPs = 6 * B * Nc * Ns * C per layer (2 QK^T + 2 AV + 2 AV^2).
#include <math.h>

#include <type_traits>

#include "common.h"
#include "ptx.cuh"

namespace mh {

constexpr int AT_BM = 128;       // query rows per tile
constexpr int AT_BN = 64;        // keys per tile
constexpr int AT_D = 64;         // head dim (dqk = dv)
constexpr int AT_DV2 = 128;      // [V~ | V~^2]
constexpr int AT_KST = 4, AT_VST = 4;
// Which of every 8 consecutive column PAIRS of an S tile take exp2 on the FMA pipes (Cody-Waite split + degree-3
// polynomial + exponent insertion) instead of MUFU.EX2: bit i set = pair i.  The MUFU issues one warp-wide EX2
// per 8 cycles per sub-partition -- 1024 cycles per key tile for the two query tiles against 768 tensor-core
// cycles -- which is what the FlashAttention-4 trick attacks.  MEASURED ON B200 (tools/microbench/mufu_bench.cu,
// DESIGN.md): the packed FFMA2 / FADD2 cost two issue cycles each, the polynomial path comes to ~12 issue cycles
// per element against 8 MUFU cycles, and with two softmax warps per sub-partition no mask beats 0 (0x88: +3 %
// before the row sums moved to FHADD, 0 % after; 0xAA: -1 %).  Kept (parity-tested) for sm_103, whose MUFU is
// the same width but whose clocks / ratios differ; default off.
#ifndef MHADA_AT_POLY
#define MHADA_AT_POLY 0
#endif
constexpr int AT_TILE_THREADS = 128;             // softmax threads per query tile (thread = row)
constexpr int AT_THREADS = 128 + 2 * AT_TILE_THREADS + 128;   // + one epilogue warpgroup
constexpr unsigned AT_POLY = MHADA_AT_POLY;
#ifndef MHADA_AT_TRACE_QUARTER
#define MHADA_AT_TRACE_QUARTER 0
#endif
constexpr int AT_TRACE_QUARTER = MHADA_AT_TRACE_QUARTER;   // which softmax warp of a tile writes the trace stamps
// 2^f on [-0.5, 0.5], minimax in relative error (7.5e-5 = 2^-13.7, far below the bf16 rounding of P that follows)
constexpr float AT_EX2_C0 = 0.9999281168f, AT_EX2_C1 = 0.6932609677f, AT_EX2_C2 = 0.2426107526f, AT_EX2_C3 = 0.0551714078f;
constexpr uint32_t AT_QC_BYTES = AT_BM * AT_D * 2;         // 16 KB per query tile and 64-channel chunk of the head
constexpr uint32_t AT_KC_BYTES = AT_BN * AT_D * 2;         // 8 KB per key tile and chunk
constexpr uint32_t AT_V_BYTES = AT_BN * AT_DV2 * 2;        // 16 KB (two 64-column boxes)
// KCH = head_dim / 64.  The kernel's "head" index runs over VALUE SLICES of 64 channels (p.H = heads * KCH): slice
// hv belongs to head hv / KCH, whose full-width Q and K (KCH chunks) it contracts for the logits, and owns value
// columns [hv * 64, +64) and their squares.  With KCH = 2 (4 heads at C = 512) each head's logits and weights are
// computed once per slice, i.e. twice: the accumulator [V~ | V~^2] of a whole 128-wide head would need 256 TMEM
// columns per query tile and leave room for one tile per CTA.
template <int KCH> constexpr uint32_t at_smem_data() {
    return 2 * KCH * AT_QC_BYTES + AT_KST * KCH * AT_KC_BYTES + AT_VST * AT_V_BYTES;
}
constexpr float AT_RESCALE_THRESHOLD = 64.0f;              // log2 units: weights of a tile may reach 2^64 before the reference moves
// 0: the two query tiles' softmax warps run free (default since the row maximum left the critical path: with one
//    MMA issuer per tile the streams de-phase on their own; cfg2 0.422 ms / cfg3 0.857 ms);
// 2: half-section hand-off through named barriers every key tile (was the default while each tile had a max phase
//    to hide; now 0.437 / 0.891 ms).
#ifndef MHADA_AT_PINGPONG
#define MHADA_AT_PINGPONG 0
#endif
constexpr int AT_PINGPONG = MHADA_AT_PINGPONG;

// Trace slots (development aid, attn_tc_kernel<true> only): clock64() stamps of CTA (0,0,0).
//   softmax WG t, iteration j: trace[(t*64 + j)*8 + e], e = 0 S ready, 1 S in registers, 2 max/rescale done,
//                              3 P stores issued, 4 P arrived
//   MMA thread:                trace[(2*64 + j)*8 + e], e = 0/2 P_t ready seen (t=0/1), 1/3 PV_t+S_t issued
constexpr int AT_TRACE_WORDS = 4 * 64 * 8;   // role 3 = TMA producer: e0 K(j+2) issued, e1 V(j) slot free, e2 V(j) issued

struct AttnTcParams {
    const __nv_bfloat16* x;   // fcs [B, Nc, ldx]
    __nv_bfloat16* out;       // [B, Nc, ldo]
    const float *x_mean, *x_rstd, *mu_v;   // [B, H*64]
    int B, H, Nc, Ns, ldx, ldo;
    int kv_shared;            // 1: one K / V' / mu_v set (style) serves every image of the batch
    int n_full;               // work items 0 .. n_full-1 are query-tile PAIRS; the items after them are single query
    int n_single;             // tiles (the two halves of the remaining pairs), each CTA's LAST item -- see attn_item()
    long long* trace;         // AT_TRACE_WORDS entries or nullptr
};

// Work item decode.  Items [0, n_full) are query-tile pairs of (image, value slice); query pairs are the fastest index so
// CTAs that run side by side share the K / V' tiles of one (image, head) in L2.  Items [n_full, n_full + n_single) are
// the pairs that do not fill a whole round of the persistent CTAs, cut into their two 128-row tiles: a single-tile item
// keeps one softmax warpgroup and one issuer busy and takes ~60 % of a pair's time (its MUFU is uncontended), so the
// last round of a single 1024 x 1024 image (512 pairs on 148 SMs: 3 rounds + 68 pairs) costs 0.6 instead of 1 round.
// A single item is always the LAST item of its CTA: the idle tile's barriers are simply never touched again.
struct AttnItem {
    int h, b, q0;
    bool single;
};
__device__ __forceinline__ AttnItem attn_item(const AttnTcParams& p, int it, int XT) {
    AttnItem a;
    int pair = it, tile = 0;
    a.single = it >= p.n_full;
    if (a.single) {
        const int u = it - p.n_full;
        pair = p.n_full + (u >> 1);
        tile = u & 1;
    }
    a.h = (pair / XT) % p.H;
    a.b = pair / (XT * p.H);
    a.q0 = (pair % XT) * (2 * AT_BM) + tile * AT_BM;
    return a;
}

struct AttnBars {
    uint64_t q_full, q_empty;
    uint64_t k_full[AT_KST], k_empty[AT_KST];
    uint64_t v_full[AT_VST], v_empty[AT_VST];
    uint64_t s_full[2][2];    // [query tile][S buffer]
    uint64_t p_ready[2][2];
    uint64_t pv_done[2];      // one phase per key tile: PV_t(j) retired (O_t quiescent until P_t(j+1) arrives)
    uint64_t o_full[2];       // one phase per work item: its last PV_t retired
    uint32_t tmem_slot;
    uint64_t epi_ready[2];    // one phase per work item: the softmax warps of tile t are done, lsum[t] is written
    uint64_t o_free[2];       // one phase per work item: the epilogue warps have read O_t out of TMEM
    float cst[3][AT_D];       // x_mean, x_rstd, mu_v of the (image, head) the epilogue warps are working on
    float lsum[2][AT_BM];     // row sums of the finished work item, per query tile
};

template <bool TRACE, int KCH>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const AttnTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    constexpr uint32_t AT_Q_BYTES = KCH * AT_QC_BYTES, AT_K_BYTES = KCH * AT_KC_BYTES;   // per query tile / key tile
    uint8_t* sK = sQ + 2 * AT_Q_BYTES;
    uint8_t* sV = sK + AT_KST * AT_K_BYTES;
    AttnBars* bars = reinterpret_cast<AttnBars*>(sV + AT_VST * AT_V_BYTES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = (p.Ns + AT_BN - 1) / AT_BN;                 // key tiles per work item
    const int XT = (p.Nc + 2 * AT_BM - 1) / (2 * AT_BM);      // query-tile pairs per (image, head)
    const int n_items = p.n_full + p.n_single;
    // PERSISTENT: this CTA handles work items blockIdx.x, blockIdx.x + gridDim.x, ... (query-tile pair fastest,
    // so CTAs that run side by side share K / V' of the same (image, head) in L2).  All pipeline counters run
    // on across items: g = (local item number) * T + j is the global key-tile index of this CTA.
    const bool tracing = TRACE && p.trace && blockIdx.x == 0;
    auto stamp = [&](int role, int j, int e) {
        if (TRACE && tracing && j < 64) p.trace[(role * 64 + j) * 8 + e] = clock64();
    };

    if (TRACE && p.trace && warp == 0 && lane == 0) {   // every CTA: SM id and entry time (globaltimer, ns)
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        p.trace[AT_TRACE_WORDS + blockIdx.x * 3 + 0] = smid;
        p.trace[AT_TRACE_WORDS + blockIdx.x * 3 + 1] = static_cast<long long>(global_timer_ns());
    }
    if (warp == 0 && lane == 0) {
        stamp(3, 0, 4);                               // kernel entry
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(&bars->q_full, 1);
        mbar_init(&bars->q_empty, 2);                   // "empty" barriers: one commit per MMA issuer
        for (int s = 0; s < AT_KST; ++s) { mbar_init(&bars->k_full[s], 1); mbar_init(&bars->k_empty[s], 2); }
        for (int s = 0; s < AT_VST; ++s) { mbar_init(&bars->v_full[s], 1); mbar_init(&bars->v_empty[s], 2); }
        for (int t = 0; t < 2; ++t) {
            for (int u = 0; u < 2; ++u) {
                mbar_init(&bars->s_full[t][u], 1);
                mbar_init(&bars->p_ready[t][u], 4);     // one arrive per softmax warp
            }
            mbar_init(&bars->pv_done[t], 1);
            mbar_init(&bars->o_full[t], 1);
            mbar_init(&bars->epi_ready[t], 4);      // one arrive per softmax warp
            mbar_init(&bars->o_free[t], 4);         // one arrive per epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(&bars->tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_slot;

    // Register hand-over: the setmaxnreg calls sit INSIDE the role branches (no join before the role
    // code) so ptxas allocates each branch against its own budget.
    if (warp < 4) {
      setmaxnreg_dec<56>();
      if (warp == 0) {
        // ===================================================================== TMA producer
        if (elect_one()) {
            int n = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++n) {
                const AttnItem wi = attn_item(p, it, XT);
                const int h = wi.h, b = wi.b, q0 = wi.q0, g0 = n * T;
                const int bkv = p.kv_shared ? 0 : b;
                auto load_k = [&](int j) {
                    const int g = g0 + j, ks = g % AT_KST;
                    mbar_wait(&bars->k_empty[ks], ((g / AT_KST) & 1) ^ 1);
                    mbar_arrive_expect_tx(&bars->k_full[ks], AT_K_BYTES);
#pragma unroll
                    for (int c = 0; c < KCH; ++c)
                        tma_load_3d(sK + ks * AT_K_BYTES + c * AT_KC_BYTES, &tmK, &bars->k_full[ks],
                                    ((h / KCH) * KCH + c) * AT_D, j * AT_BN, bkv);
                };
                mbar_wait(&bars->q_empty, (n & 1) ^ 1);      // all S MMAs of the previous item have read Q
                mbar_arrive_expect_tx(&bars->q_full, (wi.single ? 1 : 2) * AT_Q_BYTES);
#pragma unroll
                for (int c = 0; c < KCH; ++c) {
                    tma_load_3d(sQ + c * AT_QC_BYTES, &tmQ, &bars->q_full, ((h / KCH) * KCH + c) * AT_D, q0, b);
                    if (!wi.single)
                        tma_load_3d(sQ + AT_Q_BYTES + c * AT_QC_BYTES, &tmQ, &bars->q_full, ((h / KCH) * KCH + c) * AT_D,
                                    q0 + AT_BM, b);
                }
                load_k(0);
                if (T > 1) load_k(1);
                for (int j = 0; j < T; ++j) {
                    if (j + 2 < T) load_k(j + 2);          // S runs two key tiles ahead of PV
                    if (n == 0) stamp(3, j, 0);
                    const int g = g0 + j, vs = g % AT_VST;
                    mbar_wait(&bars->v_empty[vs], ((g / AT_VST) & 1) ^ 1);
                    if (n == 0) stamp(3, j, 1);
                    mbar_arrive_expect_tx(&bars->v_full[vs], AT_V_BYTES);
                    uint8_t* v = sV + vs * AT_V_BYTES;
                    tma_load_3d(v, &tmV, &bars->v_full[vs], h * AT_DV2, j * AT_BN, bkv);
                    tma_load_3d(v + AT_V_BYTES / 2, &tmV, &bars->v_full[vs], h * AT_DV2 + 64, j * AT_BN, bkv);
                    if (n == 0) stamp(3, j, 2);
                }
            }
        }
      } else if (warp == 1 || warp == 2) {
        // ===================================================================== MMA issuers
        // One issuing thread PER QUERY TILE (warp 1: tile 0, warp 2: tile 1), on two different SM sub-partitions:
        // an issuer runs ~120 instructions per key tile in the issue slots it shares with two softmax warps
        // (one thread issuing for both tiles cost its sub-partition's softmax warps ~220 cycles per key tile, and
        // the whole CTA waits for the slowest sub-partition), and the two tiles' chains no longer wait for each
        // other's P.  K / V' / Q slots are released when BOTH issuers' MMAs have read them (empty barriers count 2;
        // tcgen05.commit tracks the MMAs of the committing thread).
        if (elect_one()) {
            const int t = warp - 1;
            constexpr uint32_t idesc_s = make_idesc_bf16(AT_BM, AT_BN, 0, 0);     // S = Q K^T, both K-major
            constexpr uint32_t idesc_o = make_idesc_bf16(AT_BM, AT_DV2, 0, 1);    // O += P V', V' MN-major
            // descriptors differ in the 14-bit start-address field only: keep the high words constant
            const uint64_t q_desc = make_smem_desc(smem_u32(sQ) + t * AT_Q_BYTES, 16, 1024);
            const uint64_t k_desc = make_smem_desc(smem_u32(sK), 16, 1024);
            const uint64_t v_desc = make_smem_desc(smem_u32(sV), AT_V_BYTES / 2, 1024);   // LBO = box stride, SBO = 8 key rows
            const uint32_t s_tm0 = tmem + t * 256, o_tm0 = tmem + t * 256 + 128;
            auto issue_s = [&](int g) {                // S_t(g) -> S buffer g&1
                const uint64_t db = desc_advance(k_desc, (g % AT_KST) * AT_K_BYTES);
                const uint32_t d_tm = s_tm0 + (g & 1) * 64;
#pragma unroll
                for (int c = 0; c < KCH; ++c)
#pragma unroll
                    for (int k = 0; k < AT_D / 16; ++k)
                        umma_ss(d_tm, desc_advance(q_desc, c * AT_QC_BYTES + k * 32), desc_advance(db, c * AT_KC_BYTES + k * 32),
                                idesc_s, (c | k) != 0);
                umma_commit(&bars->s_full[t][g & 1]);
            };
            auto issue_pv = [&](int g, bool first, bool last) {
                const uint64_t db = desc_advance(v_desc, (g % AT_VST) * AT_V_BYTES);
                const uint32_t a_tm = s_tm0 + (g & 1) * 64;          // P_t(g) aliases S buffer g&1
#pragma unroll
                for (int k = 0; k < AT_BN / 16; ++k) {
                    umma_ts(o_tm0, a_tm + k * 8, desc_advance(db, k * 2048), idesc_o, (!first || k != 0) ? 1u : 0u);
                }
                umma_commit(&bars->pv_done[t]);
                if (last) umma_commit(&bars->o_full[t]);
            };
            int n = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++n) {
                const int g0 = n * T;
                mbar_wait(&bars->q_full, n & 1);
                if (t == 1 && it >= p.n_full) {
                    // single-tile item: tile 1 is idle, but the K / V' / Q slots are released by TWO arrivals -- keep
                    // pace with the loads and give this issuer's share
                    for (int j = 0; j < 2 && j < T; ++j) {
                        const int g = g0 + j;
                        mbar_wait(&bars->k_full[g % AT_KST], (g / AT_KST) & 1);
                        mbar_arrive(&bars->k_empty[g % AT_KST]);
                    }
                    if (T <= 2) mbar_arrive(&bars->q_empty);
                    for (int j = 0; j < T; ++j) {
                        const int g = g0 + j, vs = g % AT_VST;
                        mbar_wait(&bars->v_full[vs], (g / AT_VST) & 1);
                        mbar_arrive(&bars->v_empty[vs]);
                        if (j + 2 < T) {
                            mbar_wait(&bars->k_full[(g + 2) % AT_KST], ((g + 2) / AT_KST) & 1);
                            mbar_arrive(&bars->k_empty[(g + 2) % AT_KST]);
                            if (j + 3 == T) mbar_arrive(&bars->q_empty);
                        }
                    }
                    continue;
                }
                for (int j = 0; j < 2 && j < T; ++j) {
                    const int g = g0 + j;
                    mbar_wait(&bars->k_full[g % AT_KST], (g / AT_KST) & 1);
                    tc_fence_after();
                    issue_s(g);
                    umma_commit(&bars->k_empty[g % AT_KST]);
                }
                if (T <= 2) umma_commit(&bars->q_empty);
                for (int j = 0; j < T; ++j) {
                    const int g = g0 + j, vs = g % AT_VST;
                    mbar_wait(&bars->v_full[vs], (g / AT_VST) & 1);
                    if (n == 0 && t == 0) stamp(2, j, 4);
                    mbar_wait(&bars->p_ready[t][g & 1], (g >> 1) & 1);
                    if (j == 0 && n > 0) mbar_wait(&bars->o_free[t], (n - 1) & 1);   // O_t of the previous item has been read
                    tc_fence_after();
                    if (n == 0) stamp(2, j, 2 * t);
                    issue_pv(g, j == 0, j == T - 1);
                    umma_commit(&bars->v_empty[vs]);
                    if (j + 2 < T) {
                        mbar_wait(&bars->k_full[(g + 2) % AT_KST], ((g + 2) / AT_KST) & 1);
                        tc_fence_after();
                        issue_s(g + 2);
                        umma_commit(&bars->k_empty[(g + 2) % AT_KST]);
                        if (j + 3 == T) umma_commit(&bars->q_empty);   // that was the last S of this item
                    }
                    if (n == 0) stamp(2, j, 2 * t + 1);
                }
            }
        }
      }
    } else if (warp < 12) {
        setmaxnreg_inc<152>();
        // ===================================================================== softmax
        const int t = (warp - 4) >> 2;           // query tile of this warp
        const int quarter = warp & 3;            // TMEM lane quarter this warp may touch (= its SM sub-partition)
        const int row = quarter * 32 + lane;
        const uint32_t s_tm = tmem_addr(tmem, quarter * 32, t * 256);        // P(j) overwrites the first 32 columns of S(j)
        const uint32_t o_tm = tmem_addr(tmem, quarter * 32, t * 256 + 128);
        const bool tr0 = TRACE && quarter == AT_TRACE_QUARTER && lane == 0;

        // De-phasing of the two query tiles (named barriers 1, 2): the warps of tile t may start the exp2 stream
        // of a key tile once the OTHER tile's warps are half way through theirs, so one group's TMEM loads / row
        // max / P stores run while the other keeps the MUFU busy.
        if (AT_PINGPONG == 2 && t == 1) named_bar_arrive(1, 2 * AT_TILE_THREADS);

        int n = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++n) {
            const int g0 = n * T;
            const bool last_item = it + static_cast<int>(gridDim.x) >= n_items;
            if (t == 1 && it >= p.n_full) continue;       // single-tile item (always the CTA's last): tile 1 idles

            // Online softmax with a LAGGING reference: the weights of key tile j are 2^(s - m) with m the (integer)
            // reference fixed by the tiles before it; the exact maximum is taken up front for the first tile of an
            // item only.  For every later tile the row maximum is computed in the shadow of the exp2 stream -- the
            // MUFU, not the issue port, paces that stream -- and checked BEFORE the weights leave the registers:
            // only if a row's tile maximum exceeds its reference by more than 2^AT_RESCALE_THRESHOLD does the warp
            // take the slow path (rescale its part of O, recompute the tile with the new reference).  bf16 / fp32
            // are floating point, so weights up to 2^64 lose nothing; the threshold only guards the exponent range
            // (2^64 * |V~|^2 * Ns stays far below 2^127).
            float m_used = 0.f, l = 0.f;
            // FIRST is a compile-time tag: the first tile's copy of the code has the maximum in front of the weights,
            // every other tile has no data dependence between the two (a run-time select would create one)
            auto key_tile = [&](const int j, auto first_tag) {
                constexpr bool FIRST = decltype(first_tag)::value;
                const int g = g0 + j;
                mbar_wait(&bars->s_full[t][g & 1], (g >> 1) & 1);
                tc_fence_after();
                const bool tr = tr0 && n == 0;
                if (tr) stamp(t, j, 0);
                uint32_t s[AT_BN];
                const int valid = p.Ns - j * AT_BN;               // keys in this tile (tail tile: < 64)
                auto load_s = [&]() {
                    tmem_ld_x32(s_tm + (g & 1) * 64, s);
                    tmem_ld_x32(s_tm + (g & 1) * 64 + 32, s + 32);
                    tmem_wait_ld();
                    if (valid < AT_BN) {
#pragma unroll
                        for (int i = 0; i < AT_BN; ++i)
                            if (i >= valid) s[i] = 0xff800000u;   // -inf
                    }
                };
                load_s();
                if (tr) stamp(t, j, 1);
                // row maximum of the tile: 8 independent chains (a single deep FMNMX chain costs ~4 cycles per link)
                auto tile_max = [&]() {
                    float mxa[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) mxa[i] = __uint_as_float(s[i]);
#pragma unroll
                    for (int i = 8; i < AT_BN; ++i) mxa[i & 7] = fmaxf(mxa[i & 7], __uint_as_float(s[i]));
                    return fmaxf(fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3])),
                                 fmaxf(fmaxf(mxa[4], mxa[5]), fmaxf(mxa[6], mxa[7])));
                };
                // The reference is kept INTEGER-valued: rescale factors are exact powers of two and
                // frac(s - m) = frac(s), which the polynomial path uses.
                if (FIRST) m_used = ceilf(tile_max());
                if (AT_PINGPONG == 2) named_bar_sync(1 + t, 2 * AT_TILE_THREADS);
                if (tr) stamp(t, j, 2);

                // P = exp2(S - m), packed bf16x2.  Packed f32x2 adds (FADD2) for the subtraction and for the row sum.
                // The row sum uses the ROUNDED weights the MMA sees: with l = sum(p) but M, E built from bf16(p),
                // Var = E - M^2 picks up eps * M^2 (eps ~ 2^-9) and sqrt() of that is percent-level when the
                // attention is peaked.
                uint32_t pk[AT_BN / 2];
                float l_tile;
                float mxl[4];                         // this thread's running maxima of the tile, filled by weights()
                auto weights = [&](bool hand_off) {
                    const float2 neg_m = make_float2(-m_used, -m_used);
                    // polynomial path: t = s + (1.5 * 2^23 - m) rounds s to the nearest integer j (m is an integer, so
                    // the constant is exact) and leaves j - m in the low mantissa bits; f = s - j in [-0.5, 0.5];
                    // 2^(s - m) = poly(f) with (j - m) added to the exponent field.  s is clamped to m - 125 first so
                    // the exponent cannot wrap (also turns the -inf of masked tail columns into ~2^-125).
                    const float magic = 12582912.f - m_used;
                    const float2 k2 = make_float2(magic, magic), nk2 = make_float2(-magic, -magic);
                    const float s_min = m_used - 125.f;
                    float la[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int i = 0; i < 4; ++i) mxl[i] = -INFINITY;
#pragma unroll
                    for (int i = 0; i < AT_BN / 2; ++i) {
                        float2 sv = make_float2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1]));
                        // the tile maximum rides along in the issue slots the MUFU leaves free (FMNMX3)
                        mxl[i & 3] = fmaxf(fmaxf(mxl[i & 3], sv.x), sv.y);
                        if ((AT_POLY >> (i & 7)) & 1u) {
                            sv.x = fmaxf(sv.x, s_min);
                            sv.y = fmaxf(sv.y, s_min);
                            const float2 tt = __fadd2_rn(sv, k2);
                            const float2 f = __ffma2_rn(__fadd2_rn(tt, nk2), make_float2(-1.f, -1.f), sv);
                            float2 pp = __ffma2_rn(make_float2(AT_EX2_C3, AT_EX2_C3), f, make_float2(AT_EX2_C2, AT_EX2_C2));
                            pp = __ffma2_rn(pp, f, make_float2(AT_EX2_C1, AT_EX2_C1));
                            pp = __ffma2_rn(pp, f, make_float2(AT_EX2_C0, AT_EX2_C0));
                            pk[i] = pack_bf16x2(__int_as_float(__float_as_int(pp.x) + (__float_as_int(tt.x) << 23)),
                                                __int_as_float(__float_as_int(pp.y) + (__float_as_int(tt.y) << 23)));
                        } else {
                            const float2 x = __fadd2_rn(sv, neg_m);
                            pk[i] = pack_bf16x2(ex2_approx(x.x), ex2_approx(x.y));
                        }
                        // row sum of the ROUNDED weights: FHADD.BF16 adds one bf16 half of the packed word to an f32
                        la[(2 * i) & 3] = add_f32_bf16_lo(la[(2 * i) & 3], pk[i]);
                        la[(2 * i + 1) & 3] = add_f32_bf16_hi(la[(2 * i + 1) & 3], pk[i]);
                        // half way: let the other tile's warps start their stream (not after the very last tile of the CTA)
                        if (AT_PINGPONG == 2 && hand_off && i == AT_BN / 4 - 1 && !(t == 1 && last_item && j == T - 1))
                            named_bar_arrive(2 - t, 2 * AT_TILE_THREADS);
                    }
                    l_tile = (la[0] + la[1]) + (la[2] + la[3]);
                };
                weights(true);
                if (!FIRST) {
                    const float mx = fmaxf(fmaxf(mxl[0], mxl[1]), fmaxf(mxl[2], mxl[3]));
                    const bool grow = mx > m_used + AT_RESCALE_THRESHOLD;
                    if (__any_sync(0xffffffffu, grow)) {
                        // slow path (rare): new reference for the rows that grew, O and l brought to it, weights redone
                        const float m_new = grow ? ceilf(mx) : m_used;
                        const float sc = ex2_approx(m_used - m_new);    // 1 for rows that keep their reference
                        l *= sc;
                        m_used = m_new;
                        // O_t may only be touched once PV_t(g-1) has retired; PV_t(g) cannot start before our P arrives
                        mbar_wait(&bars->pv_done[t], (g - 1) & 1);
                        tc_fence_after();
#pragma unroll 1
                        for (int c = 0; c < AT_DV2; c += 32) {
                            uint32_t o[32];
                            tmem_ld_x32(o_tm + c, o);
                            tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * sc);
                            tmem_st_x32(o_tm + c, o);
                        }
                        load_s();                       // S is still intact in TMEM: P has not been stored yet
                        weights(false);
                    }
                }
                l += l_tile;
#pragma unroll
                for (int c = 0; c < AT_BN / 2; c += 16) tmem_st_x16(s_tm + (g & 1) * 64 + c, pk + c);
                if (tr) stamp(t, j, 3);
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->p_ready[t][g & 1]);
                if (tr) stamp(t, j, 4);
            };
            key_tile(0, std::true_type{});
            for (int j = 1; j < T; ++j) key_tile(j, std::false_type{});

            // ---- hand the finished tile to the epilogue warps and go straight on to the next work item
            // (lsum[t] / epi_ready[t] may only be reused once the epilogue of the PREVIOUS item has read them: it
            // arrived on o_free[t] after doing so; in steady state that was ~a whole item ago)
            if (n > 0) mbar_wait(&bars->o_free[t], (n - 1) & 1);
            bars->lsum[t][row] = l;
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->epi_ready[t]);
            if (tr0 && n == 0) stamp(t, 63, 5);
        }
    } else {
        setmaxnreg_inc<152>();
        // ===================================================================== epilogue warpgroup (warps 12-15)
        // O/l -> (M~, E~) -> sqrt(max(E~ - M~^2, 1e-6)) * IN(fcs) + M~ + mu_v for both query tiles of a work item,
        // while the softmax warps and the tensor core are already in the next item: per item the softmax warps used
        // to spend ~3400 cycles here plus the wait for their slowest sibling (8 % of a cfg2 item).
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int et = threadIdx.x - (AT_THREADS - 128);          // 0 .. 127
        const bool tr0 = TRACE && quarter == AT_TRACE_QUARTER && lane == 0;
        int n = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++n) {
            const AttnItem wi = attn_item(p, it, XT);
            const int h = wi.h, b = wi.b, q0 = wi.q0;
            {
                const size_t sidx = (static_cast<size_t>(b) * p.H + h) * AT_D;
                named_bar_sync(3, 128);                   // the previous item's epilogue is done with cst
                if (et < AT_D) {
                    bars->cst[0][et] = p.x_mean[sidx + et];
                    bars->cst[1][et] = p.x_rstd[sidx + et];
                } else {
                    const size_t vidx = (static_cast<size_t>(p.kv_shared ? 0 : b) * p.H + h) * AT_D;
                    bars->cst[2][et - AT_D] = p.mu_v ? p.mu_v[vidx + et - AT_D] : 0.f;
                }
                named_bar_sync(3, 128);
            }
#pragma unroll 1
            for (int t = 0; t < (wi.single ? 1 : 2); ++t) {
                const uint32_t o_tm = tmem_addr(tmem, quarter * 32, t * 256 + 128);
                const int nrow = q0 + t * AT_BM + row;
                const bool row_ok = nrow < p.Nc;
                const size_t tok = static_cast<size_t>(b) * p.Nc + (row_ok ? nrow : 0);
                const __nv_bfloat16* xrow = p.x + tok * p.ldx + h * AT_D;
                __nv_bfloat16* orow = p.out + tok * p.ldo + h * AT_D;
                // The fcs row is fetched before waiting for the tile, so its latency hides behind the attention.
                uint4 xv[AT_D / 8];
                if (row_ok) {
#pragma unroll
                    for (int i = 0; i < AT_D / 8; ++i) xv[i] = __ldg(reinterpret_cast<const uint4*>(xrow) + i);
                }
                mbar_wait(&bars->epi_ready[t], n & 1);    // softmax warps done: lsum[t] is valid
                // (pv_done cannot be used here: a parity wait is only meaningful while the barrier is at most one
                // phase ahead -- o_full has one phase per item)
                mbar_wait(&bars->o_full[t], n & 1);       // the last PV of the item has retired
                tc_fence_after();
                const float inv = 1.f / bars->lsum[t][row];
                const uint32_t* xw = reinterpret_cast<const uint32_t*>(xv);
#pragma unroll 1
                for (int c = 0; c < AT_D; c += 32) {
                    uint32_t mm[32], ee[32];
                    tmem_ld_x32(o_tm + c, mm);
                    tmem_ld_x32(o_tm + AT_D + c, ee);
                    tmem_wait_ld();
                    if (c + 32 == AT_D) {
                        // O_t and lsum[t] are in registers: the next item's first PV may overwrite the accumulator
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bars->o_free[t]);
                    }
                    uint32_t ov[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float r[2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int ch = c + 2 * i + e;
                            const float m = __uint_as_float(mm[2 * i + e]) * inv;
                            const float ex = __uint_as_float(ee[2 * i + e]) * inv;
                            float sd;
                            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sd) : "f"(fmaxf(fmaf(-m, m, ex), 1e-6f)));
                            const float xf = e == 0 ? bf16_lo(xw[(c >> 1) + i]) : bf16_hi(xw[(c >> 1) + i]);
                            const float xn = (xf - bars->cst[0][ch]) * bars->cst[1][ch];
                            r[e] = fmaf(sd, xn, m + bars->cst[2][ch]);
                        }
                        ov[i] = pack_bf16x2(r[0], r[1]);
                    }
                    if (row_ok) {
                        st_global_256(orow + c, ov);
                        st_global_256(orow + c + 16, ov + 8);
                    }
                }
                if (tr0 && n == 0) stamp(t, 63, 6);
            }
        }
    }
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem, 512);
    if (TRACE && warp == 2 && lane == 0) stamp(3, 0, 5);   // kernel exit
    if (TRACE && p.trace && warp == 2 && lane == 0)
        p.trace[AT_TRACE_WORDS + blockIdx.x * 3 + 2] = static_cast<long long>(global_timer_ns());
}

int launch_attn_bf16(const mhada_attn_args& a, cudaStream_t s) { return launch_attn_bf16_impl(a, nullptr, s); }

template <int KCH>
static int launch_attn_bf16_kch(const mhada_attn_args& a, long long* trace, cudaStream_t s);

int launch_attn_bf16_impl(const mhada_attn_args& a, long long* trace, cudaStream_t s) {
    if (a.dqk == 2 * AT_D) return launch_attn_bf16_kch<2>(a, trace, s);
    return launch_attn_bf16_kch<1>(a, trace, s);
}

template <int KCH>
static int launch_attn_bf16_kch(const mhada_attn_args& a, long long* trace, cudaStream_t s) {
    const int C = a.H * KCH * AT_D;
    const int kvB = a.kv_batch == 1 ? 1 : a.B;       // style batch: 1 = shared by all images
    CUtensorMap tmQ, tmK, tmV;
    {
        uint64_t dims[3] = {static_cast<uint64_t>(C), static_cast<uint64_t>(a.Nc), static_cast<uint64_t>(a.B)};
        uint64_t str[2] = {static_cast<uint64_t>(a.ldq) * 2, static_cast<uint64_t>(a.Nc) * a.ldq * 2};
        uint32_t box[3] = {AT_D, AT_BM, 1};
        if (int e = make_tmap_bf16(&tmQ, a.q, 3, dims, str, box)) return e;
    }
    {
        uint64_t dims[3] = {static_cast<uint64_t>(C), static_cast<uint64_t>(a.Ns), static_cast<uint64_t>(kvB)};
        uint64_t str[2] = {static_cast<uint64_t>(a.ldk) * 2, static_cast<uint64_t>(a.Ns) * a.ldk * 2};
        uint32_t box[3] = {AT_D, AT_BN, 1};
        if (int e = make_tmap_bf16(&tmK, a.k, 3, dims, str, box)) return e;
    }
    {
        uint64_t dims[3] = {static_cast<uint64_t>(2 * C), static_cast<uint64_t>(a.Ns), static_cast<uint64_t>(kvB)};
        uint64_t str[2] = {static_cast<uint64_t>(a.ldv) * 2, static_cast<uint64_t>(a.Ns) * a.ldv * 2};
        uint32_t box[3] = {64, AT_BN, 1};
        if (int e = make_tmap_bf16(&tmV, a.v, 3, dims, str, box)) return e;
    }
    AttnTcParams p;
    p.x = static_cast<const __nv_bfloat16*>(a.x);
    p.out = static_cast<__nv_bfloat16*>(a.out);
    p.x_mean = a.x_mean; p.x_rstd = a.x_rstd; p.mu_v = a.mu_v;
    p.B = a.B; p.H = a.H * KCH; p.Nc = a.Nc; p.Ns = a.Ns; p.ldx = a.ldx; p.ldo = a.ldo;     // p.H: value slices
    p.kv_shared = (a.kv_batch == 1 && a.B > 1) ? 1 : 0;
    p.trace = trace;
    constexpr size_t smem = at_smem_data<KCH>() + sizeof(AttnBars) + 1024;
    static DeviceOnce once_plain, once_trace;          // per (kernel instantiation, device)
    if (int e = smem_attr_once(once_plain, reinterpret_cast<const void*>(attn_tc_kernel<false, KCH>), smem, "attn smem attr")) return e;
    if (trace)
        if (int e = smem_attr_once(once_trace, reinterpret_cast<const void*>(attn_tc_kernel<true, KCH>), smem, "attn smem attr")) return e;
    const int n_pairs = ((a.Nc + 2 * AT_BM - 1) / (2 * AT_BM)) * p.H * a.B;
    int n_items = n_pairs;
    p.n_full = n_pairs;
    p.n_single = 0;
    const int n_sm = sm_count();
    // Persistent (one CTA per SM, static round-robin over equal-cost items) when the last round is well filled;
    // otherwise one item per CTA so the hardware scheduler hands the tail to whichever SM is free first
    // (measured r1: cfg3 = 512 items -> 0.93 ms dynamic vs 0.99 ms static; cfg2 = 1024 items -> 0.465 vs 0.479).
    const int rounds = (n_items + n_sm - 1) / n_sm;
#ifdef MHADA_AT_FORCE_PERSISTENT
    bool persistent = MHADA_AT_FORCE_PERSISTENT != 0;
#else
    bool persistent = n_items > n_sm && static_cast<double>(n_items) / (static_cast<double>(rounds) * n_sm) >= 0.9;
#endif
#ifndef MHADA_AT_NO_SINGLE_TAIL
    // A remainder of at most n_sm / 2 pairs after whole rounds: run it as single-tile items in the last round
    // (cfg3: 512 pairs = 3 x 148 + 68 -> 444 pairs + 136 single tiles on 148 persistent CTAs).
    const int rem = n_pairs % n_sm;
    if (!trace && n_pairs > n_sm && rem > 0 && 2 * rem <= n_sm) {
        p.n_full = n_pairs - rem;
        p.n_single = 2 * rem;
        n_items = p.n_full + p.n_single;
        persistent = true;
    }
#endif
    dim3 grid(persistent ? n_sm : n_items);
    if (trace)
        attn_tc_kernel<true, KCH><<<grid, AT_THREADS, smem, s>>>(tmQ, tmK, tmV, p);
    else
        attn_tc_kernel<false, KCH><<<grid, AT_THREADS, smem, s>>>(tmQ, tmK, tmV, p);
    count_launch();
    return check_cuda(cudaGetLastError(), "attn_tc launch");
}

}  // namespace mh

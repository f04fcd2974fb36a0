// K2 + K3 -- streaming MHAda attention on tcgen05 / TMEM / TMA with the AdaIN-style epilogue fused.
//
// Replaces MHAdaSTr/network/adaDecoder.py:186-198 for one head:
//     A = softmax(Q K^T)   (no 1/sqrt(d); Q arrives pre-multiplied by log2 e -> exp2)
//     M = A V ;  Var = A V^2 - M^2 ;  S = sqrt(clamp(Var, 1e-6)) ;  out = S * IN(fcs) + M
// The Nc x Ns map never leaves the SM: S tiles live in TMEM, P overwrites S in place as bf16 and
// is consumed straight from TMEM by the second MMA, which multiplies by V' = [V~ | V~^2] (128 wide)
// so the mean and the second moment come out of ONE pass (V~ = V - mu_v, added back at the end).
//
// CTA = 2 query tiles of 128 rows x 1 head, key tiles of 64, 12 warps:
//   warp 0       TMA producer: Q (once), K ring (4 x 8 KB), V' ring (4 x 16 KB), 128B swizzle
//   warp 1       tcgen05.mma issuer (one elected lane)
//   warp 2       TMEM allocator (512 columns; per query tile t: S buffers at t*256 + {0, 64}, O at t*256 + 128)
//   warp 3       stages the epilogue constants (mean/rstd of fcs, mu_v) into shared memory
//   warps 4-7    softmax + epilogue for query tile 0 (thread = row; TMEM lane = row, no shuffles)
//   warps 8-11   same for query tile 1
// Pipeline (measured motivation in DESIGN.md, attention section): S is DOUBLE-BUFFERED in TMEM, so
// S(j+1), S(j+2) are computed while the softmax warps still work on tile j -- they do not wait for the
// tensor core in steady state, and the exp2 (MUFU) units, which bound this head_dim-64 problem, stay busy.
// MMA issue order per key tile j and query tile t:  [P_t(j) ready]  PV_t(j)  S_t(j+2).
// Rescaling of O is lazy (only when a row max grows by more than 2^8) and done in place in TMEM by the
// softmax warps after waiting for the last PV into that accumulator.
//
// Tensor-bound: algorithmic FLOPs = 6 * B * Nc * Ns * C per layer (2 QK^T + 2 AV + 2 AV^2).
#include <math.h>

#include "common.h"
#include "ptx.cuh"

namespace mh {

constexpr int AT_BM = 128;       // query rows per tile
constexpr int AT_BN = 64;        // keys per tile
constexpr int AT_D = 64;         // head dim (dqk = dv)
constexpr int AT_DV2 = 128;      // [V~ | V~^2]
constexpr int AT_KST = 4, AT_VST = 4;
constexpr int AT_THREADS = 384;
constexpr uint32_t AT_Q_BYTES = AT_BM * AT_D * 2;          // 16 KB per query tile
constexpr uint32_t AT_K_BYTES = AT_BN * AT_D * 2;          // 8 KB
constexpr uint32_t AT_V_BYTES = AT_BN * AT_DV2 * 2;        // 16 KB (two 64-column boxes)
constexpr uint32_t AT_SMEM_DATA = 2 * AT_Q_BYTES + AT_KST * AT_K_BYTES + AT_VST * AT_V_BYTES;
constexpr float AT_RESCALE_THRESHOLD = 8.0f;               // log2 units: P <= 2^8
#ifndef MHADA_AT_PINGPONG
#define MHADA_AT_PINGPONG 2
#endif
constexpr int AT_PINGPONG = MHADA_AT_PINGPONG;   // 0 free-running, 1 one-time half-tile offset, 2 half-section hand-off every tile

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
    long long* trace;         // AT_TRACE_WORDS entries or nullptr
};

struct AttnBars {
    uint64_t q_full, q_empty;
    uint64_t k_full[AT_KST], k_empty[AT_KST];
    uint64_t v_full[AT_VST], v_empty[AT_VST];
    uint64_t s_full[2][2];    // [query tile][S buffer]
    uint64_t p_ready[2][2];
    uint64_t pv_done[2];      // one phase per key tile: PV_t(j) retired (O_t quiescent until P_t(j+1) arrives)
    uint64_t o_full[2];       // one phase per work item: its last PV_t retired
    uint32_t tmem_slot;
    float cst[2][3][AT_D];    // per warpgroup: x_mean, x_rstd, mu_v of the current (image, head)
};

template <bool TRACE>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const AttnTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + 2 * AT_Q_BYTES;
    uint8_t* sV = sK + AT_KST * AT_K_BYTES;
    AttnBars* bars = reinterpret_cast<AttnBars*>(sV + AT_VST * AT_V_BYTES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = (p.Ns + AT_BN - 1) / AT_BN;                 // key tiles per work item
    const int XT = (p.Nc + 2 * AT_BM - 1) / (2 * AT_BM);      // query-tile pairs per (image, head)
    const int n_items = XT * p.H * p.B;
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
        mbar_init(&bars->q_empty, 1);
        for (int s = 0; s < AT_KST; ++s) { mbar_init(&bars->k_full[s], 1); mbar_init(&bars->k_empty[s], 1); }
        for (int s = 0; s < AT_VST; ++s) { mbar_init(&bars->v_full[s], 1); mbar_init(&bars->v_empty[s], 1); }
        for (int t = 0; t < 2; ++t) {
            for (int u = 0; u < 2; ++u) {
                mbar_init(&bars->s_full[t][u], 1);
                mbar_init(&bars->p_ready[t][u], 4);     // one arrive per softmax warp
            }
            mbar_init(&bars->pv_done[t], 1);
            mbar_init(&bars->o_full[t], 1);
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
      setmaxnreg_dec<64>();
      if (warp == 0) {
        // ===================================================================== TMA producer
        if (elect_one()) {
            int n = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++n) {
                const int qx = it % XT, h = (it / XT) % p.H, b = it / (XT * p.H);
                const int q0 = qx * (2 * AT_BM), g0 = n * T;
                const int bkv = p.kv_shared ? 0 : b;
                auto load_k = [&](int j) {
                    const int g = g0 + j, ks = g % AT_KST;
                    mbar_wait(&bars->k_empty[ks], ((g / AT_KST) & 1) ^ 1);
                    mbar_arrive_expect_tx(&bars->k_full[ks], AT_K_BYTES);
                    tma_load_3d(sK + ks * AT_K_BYTES, &tmK, &bars->k_full[ks], h * AT_D, j * AT_BN, bkv);
                };
                mbar_wait(&bars->q_empty, (n & 1) ^ 1);      // all S MMAs of the previous item have read Q
                mbar_arrive_expect_tx(&bars->q_full, 2 * AT_Q_BYTES);
                tma_load_3d(sQ, &tmQ, &bars->q_full, h * AT_D, q0, b);
                tma_load_3d(sQ + AT_Q_BYTES, &tmQ, &bars->q_full, h * AT_D, q0 + AT_BM, b);
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
      } else if (warp == 1) {
        // ===================================================================== MMA issuer
        if (elect_one()) {
            constexpr uint32_t idesc_s = make_idesc_bf16(AT_BM, AT_BN, 0, 0);     // S = Q K^T, both K-major
            constexpr uint32_t idesc_o = make_idesc_bf16(AT_BM, AT_DV2, 0, 1);    // O += P V', V' MN-major
            const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV);
            auto issue_s = [&](int t, int g) {         // S_t(g) -> S buffer g&1 of query tile t
                const int ks = g % AT_KST;
                const uint64_t da = make_smem_desc(q_addr + t * AT_Q_BYTES, 16, 1024);
                const uint64_t db = make_smem_desc(k_addr + ks * AT_K_BYTES, 16, 1024);
                const uint32_t d_tm = tmem + t * 256 + (g & 1) * 64;
#pragma unroll
                for (int k = 0; k < AT_D / 16; ++k)
                    umma_ss(d_tm, desc_advance(da, k * 32), desc_advance(db, k * 32), idesc_s, k != 0);
                umma_commit(&bars->s_full[t][g & 1]);
            };
            auto issue_pv = [&](int t, int g, bool first, bool last) {
                // B = V' tile [64 keys][128 cols] as two [64][64] boxes: LBO = box stride, SBO = 8 key rows
                const int vs = g % AT_VST;
                const uint64_t db = make_smem_desc(v_addr + vs * AT_V_BYTES, AT_V_BYTES / 2, 1024);
                const uint32_t a_tm = tmem + t * 256 + (g & 1) * 64;      // P_t(g) aliases S buffer g&1
#pragma unroll
                for (int k = 0; k < AT_BN / 16; ++k)
                    umma_ts(tmem + t * 256 + 128, a_tm + k * 8, desc_advance(db, k * 2048), idesc_o,
                            (!first || k != 0) ? 1u : 0u);
                umma_commit(&bars->pv_done[t]);
                if (last) umma_commit(&bars->o_full[t]);
            };
            int n = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++n) {
                const int g0 = n * T;
                mbar_wait(&bars->q_full, n & 1);
                for (int j = 0; j < 2 && j < T; ++j) {
                    const int g = g0 + j;
                    mbar_wait(&bars->k_full[g % AT_KST], (g / AT_KST) & 1);
                    tc_fence_after();
                    issue_s(0, g);
                    issue_s(1, g);
                    umma_commit(&bars->k_empty[g % AT_KST]);
                }
                if (T <= 2) umma_commit(&bars->q_empty);
                for (int j = 0; j < T; ++j) {
                    const int g = g0 + j, vs = g % AT_VST;
                    const bool more = (j + 2 < T);
                    mbar_wait(&bars->v_full[vs], (g / AT_VST) & 1);
                    if (n == 0) stamp(2, j, 4);
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        mbar_wait(&bars->p_ready[t][g & 1], (g >> 1) & 1);
                        tc_fence_after();
                        if (n == 0) stamp(2, j, 2 * t);
                        issue_pv(t, g, j == 0, j == T - 1);
                        if (t == 1) umma_commit(&bars->v_empty[vs]);
                        if (more) {
                            if (t == 0) {
                                mbar_wait(&bars->k_full[(g + 2) % AT_KST], ((g + 2) / AT_KST) & 1);
                                tc_fence_after();
                            }
                            issue_s(t, g + 2);
                            if (t == 1) {
                                umma_commit(&bars->k_empty[(g + 2) % AT_KST]);
                                if (j + 3 == T) umma_commit(&bars->q_empty);   // that was the last S of this item
                            }
                        }
                        if (n == 0) stamp(2, j, 2 * t + 1);
                    }
                }
            }
        }
      }
    } else {
        setmaxnreg_inc<216>();
        // ===================================================================== softmax + epilogue
        const int t = (warp - 4) >> 2;           // query tile of this warpgroup
        const int quarter = warp & 3;            // TMEM lane quarter this warp may touch
        const int row = quarter * 32 + lane;
        const int wg_tid = threadIdx.x - 128 - t * 128;
        const uint32_t s_tm = tmem_addr(tmem, quarter * 32, t * 256);
        const uint32_t o_tm = tmem_addr(tmem, quarter * 32, t * 256 + 128);
        const bool tr0 = TRACE && quarter == 0 && lane == 0;

        // De-phasing of the two warpgroups (named barriers 1, 2; 256 threads each): warpgroup t may start the
        // exp2 stream of a tile once the OTHER one is half way through its own, so one group's TMEM loads / row
        // max / P stores run while the other keeps the MUFU busy, and the two streams overlap by half (a lone
        // warp per SMSP cannot saturate the MUFU, two can).
        if (AT_PINGPONG == 2 && t == 1) named_bar_arrive(1, 256);

        int n = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++n) {
            const int qx = it % XT, h = (it / XT) % p.H, b = it / (XT * p.H);
            const int q0 = qx * (2 * AT_BM), g0 = n * T;
            const bool last_item = it + static_cast<int>(gridDim.x) >= n_items;
            // epilogue operands that do not depend on the attention: issue their loads now
            {
                const size_t sidx = (static_cast<size_t>(b) * p.H + h) * AT_D;
                named_bar_sync(3 + t, 128);          // previous item's epilogue of this warpgroup is done with cst
                if (wg_tid < AT_D) {
                    bars->cst[t][0][wg_tid] = p.x_mean[sidx + wg_tid];
                    bars->cst[t][1][wg_tid] = p.x_rstd[sidx + wg_tid];
                } else {
                    const size_t vidx = (static_cast<size_t>(p.kv_shared ? 0 : b) * p.H + h) * AT_D;
                    bars->cst[t][2][wg_tid - AT_D] = p.mu_v ? p.mu_v[vidx + wg_tid - AT_D] : 0.f;
                }
            }
            const int nrow = q0 + t * AT_BM + row;
            const bool row_ok = nrow < p.Nc;
            const size_t tok = static_cast<size_t>(b) * p.Nc + (row_ok ? nrow : 0);
            const __nv_bfloat16* xrow = p.x + tok * p.ldx + h * AT_D;
            __nv_bfloat16* orow = p.out + tok * p.ldo + h * AT_D;

            float m_used = -INFINITY, l = 0.f;
            for (int j = 0; j < T; ++j) {
                const int g = g0 + j;
                const uint32_t sb_tm = s_tm + (g & 1) * 64;
                mbar_wait(&bars->s_full[t][g & 1], (g >> 1) & 1);
                tc_fence_after();
                const bool tr = tr0 && n == 0;
                if (tr) stamp(t, j, 0);
                uint32_t s[64];
                tmem_ld_x32(sb_tm, s);
                tmem_ld_x32(sb_tm + 32, s + 32);
                tmem_wait_ld();
                if (tr) stamp(t, j, 1);
                const int valid = p.Ns - j * AT_BN;       // keys in this tile (tail tile: < 64)
                if (valid < AT_BN) {
#pragma unroll
                    for (int i = 0; i < 64; ++i)
                        if (i >= valid) s[i] = 0xff800000u;   // -inf
                }
                // row max: 8 independent chains (a single deep FMNMX chain costs ~4 cycles per link)
                float mxa[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) mxa[i] = __uint_as_float(s[i]);
#pragma unroll
                for (int i = 8; i < 64; ++i) mxa[i & 7] = fmaxf(mxa[i & 7], __uint_as_float(s[i]));
                const float mx = fmaxf(fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3])),
                                       fmaxf(fmaxf(mxa[4], mxa[5]), fmaxf(mxa[6], mxa[7])));
                if (j == 0) {
                    m_used = mx;
                } else {
                    const bool grow = mx > m_used + AT_RESCALE_THRESHOLD;
                    if (__any_sync(0xffffffffu, grow)) {
                        const float m_new = grow ? mx : m_used;
                        const float sc = ex2_approx(m_used - m_new);    // 1 for rows that keep their max
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
                    }
                }
                if (AT_PINGPONG == 2) named_bar_sync(1 + t, 256);
                if (tr) stamp(t, j, 2);
                // P = exp2(S - m), packed bf16x2, written over the first 32 columns of this S buffer.
                // Packed f32x2 adds (FADD2) for the subtraction and for the row sum; four independent sum chains.
                // The row sum uses the ROUNDED weights the MMA sees: with l = sum(p) but M, E built from bf16(p),
                // Var = E - M^2 picks up eps * M^2 (eps ~ 2^-9) and sqrt() of that is percent-level when the
                // attention is peaked.
                const float2 neg_m = make_float2(-m_used, -m_used);
                float2 la = make_float2(0.f, 0.f), lb = make_float2(0.f, 0.f);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float2 x = __fadd2_rn(make_float2(__uint_as_float(s[c * 32 + 2 * i]),
                                                                __uint_as_float(s[c * 32 + 2 * i + 1])), neg_m);
                        pk[i] = pack_bf16x2(ex2_approx(x.x), ex2_approx(x.y));
                        const float2 r = make_float2(bf16_lo(pk[i]), bf16_hi(pk[i]));
                        if (i & 1) la = __fadd2_rn(la, r); else lb = __fadd2_rn(lb, r);
                    }
                    tmem_st_x16(sb_tm + c * 16, pk);
                    // half way: let the other warpgroup start its stream (not after the very last tile of the CTA)
                    if (AT_PINGPONG == 2 && c == 0 && !(t == 1 && last_item && j == T - 1)) named_bar_arrive(2 - t, 256);
                }
                l += (la.x + la.y) + (lb.x + lb.y);
                if (tr) stamp(t, j, 3);
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->p_ready[t][g & 1]);
                if (tr) stamp(t, j, 4);
            }

            // ---- epilogue: O/l -> (M~, E~) -> sqrt(max(E~ - M~^2, 1e-6)) * IN(fcs) + M~ + mu_v
            // The fcs row is fetched before waiting for the last PV, so its latency hides behind the MMA tail.
            uint4 xv[8];
            if (row_ok) {
#pragma unroll
                for (int i = 0; i < 8; ++i) xv[i] = __ldg(reinterpret_cast<const uint4*>(xrow) + i);
            }
            named_bar_sync(3 + t, 128);              // cst[t] of this item is complete
            // (pv_done cannot be used here: a parity wait is only meaningful while the barrier is at most one
            // phase ahead, and a late warp may find both PV(T-2) and PV(T-1) retired -- o_full has one phase per item)
            mbar_wait(&bars->o_full[t], n & 1);
            if (tr0 && n == 0) stamp(t, 63, 5);
            tc_fence_after();
            const float inv = 1.f / l;
            const uint32_t* xw = reinterpret_cast<const uint32_t*>(xv);
            // kept rolled on purpose: unrolling it costs the main loop registers (measured 0.472 -> 0.444 ms on cfg2)
#pragma unroll 1
            for (int c = 0; c < AT_D; c += 32) {
                uint32_t mm[32], ee[32];
                tmem_ld_x32(o_tm + c, mm);
                tmem_ld_x32(o_tm + AT_D + c, ee);
                tmem_wait_ld();
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
                        const float xn = (xf - bars->cst[t][0][ch]) * bars->cst[t][1][ch];
                        r[e] = fmaf(sd, xn, m + bars->cst[t][2][ch]);
                    }
                    ov[i] = pack_bf16x2(r[0], r[1]);
                }
                if (row_ok) {
                    st_global_256(orow + c, ov);
                    st_global_256(orow + c + 16, ov + 8);
                }
            }
            tc_fence_before();
            if (tr0 && n == 0) stamp(t, 63, 6);
        }
    }
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem, 512);
    if (TRACE && warp == 2 && lane == 0) stamp(3, 0, 5);   // kernel exit
    if (TRACE && p.trace && warp == 2 && lane == 0)
        p.trace[AT_TRACE_WORDS + blockIdx.x * 3 + 2] = static_cast<long long>(global_timer_ns());
}

int launch_attn_bf16(const mhada_attn_args& a, cudaStream_t s) { return launch_attn_bf16_impl(a, nullptr, s); }

int launch_attn_bf16_impl(const mhada_attn_args& a, long long* trace, cudaStream_t s) {
    const int C = a.H * AT_D;
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
    p.B = a.B; p.H = a.H; p.Nc = a.Nc; p.Ns = a.Ns; p.ldx = a.ldx; p.ldo = a.ldo;
    p.kv_shared = (a.kv_batch == 1 && a.B > 1) ? 1 : 0;
    p.trace = trace;
    constexpr size_t smem = AT_SMEM_DATA + sizeof(AttnBars) + 1024;
    static bool attr_done = false;
    if (!attr_done) {
        if (int e = check_cuda(cudaFuncSetAttribute(attn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    static_cast<int>(smem)), "attn smem attr"))
            return e;
        if (int e = check_cuda(cudaFuncSetAttribute(attn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    static_cast<int>(smem)), "attn smem attr"))
            return e;
        attr_done = true;
    }
    const int n_items = ((a.Nc + 2 * AT_BM - 1) / (2 * AT_BM)) * a.H * a.B;
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
    }
    // Persistent (one CTA per SM, static round-robin over equal-cost items) when the last round is well filled;
    // otherwise one item per CTA so the hardware scheduler hands the tail to whichever SM is free first
    // (measured r1: cfg3 = 512 items -> 0.93 ms dynamic vs 0.99 ms static; cfg2 = 1024 items -> 0.465 vs 0.479).
    const int rounds = (n_items + n_sm - 1) / n_sm;
#ifdef MHADA_AT_FORCE_PERSISTENT
    const bool persistent = MHADA_AT_FORCE_PERSISTENT != 0;
#else
    const bool persistent = n_items > n_sm && static_cast<double>(n_items) / (static_cast<double>(rounds) * n_sm) >= 0.9;
#endif
    dim3 grid(persistent ? n_sm : n_items);
    if (trace)
        attn_tc_kernel<true><<<grid, AT_THREADS, smem, s>>>(tmQ, tmK, tmV, p);
    else
        attn_tc_kernel<false><<<grid, AT_THREADS, smem, s>>>(tmQ, tmK, tmV, p);
    count_launch();
    return check_cuda(cudaGetLastError(), "attn_tc launch");
}

}  // namespace mh

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
// Folds the statistics (mu, rs: 64 channels of one head, global or shared memory) into the weights of projection
// `which` of image b, head h.  256 threads = 8 warps: warp w folds output rows w, w+8, ...; lanes walk the input
// channels, so every global read / write is a contiguous row segment.
__device__ __forceinline__ void fold_head(const float* __restrict__ w, const float* __restrict__ bias, const float* mu,
                                          const float* rs, int which, int b, int h, int B, int H, int d,
                                          __nv_bfloat16* __restrict__ wf, float* __restrict__ bf,
                                          float* __restrict__ mu_v) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int C = H * d;
    const float gain = which == 0 ? kLog2e : 1.f;
    if (d == PRJ_D) {
        // head_dim 64 (the only one the tensor-core path runs): all 16 weight loads of the warp's 8 rows are issued
        // before the first use -- row by row, each row waited a full L2 round trip (14.6 us for 256 tiny blocks)
        float wv[8][2];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int k = 0; k < 2; ++k)
                wv[r][k] = __ldg(w + ((static_cast<size_t>(which) * H + h) * PRJ_D + warp + 8 * r) * PRJ_D + lane + 32 * k);
        const float mu0 = mu[lane], mu1 = mu[lane + 32];
        const float rs0 = which != 2 ? rs[lane] : 1.f, rs1 = which != 2 ? rs[lane + 32] : 1.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int o = warp + 8 * r;
            __nv_bfloat16* dst = wf + (((static_cast<size_t>(which) * B + b) * H + h) * PRJ_D + o) * PRJ_D;
            float wa = wv[r][0] * gain, wb = wv[r][1] * gain;
            if (which != 2) { wa *= rs0; wb *= rs1; }
            const __nv_bfloat16 ba = __float2bfloat16_rn(wa), bb = __float2bfloat16_rn(wb);
            dst[lane] = ba;
            dst[lane + 32] = bb;
            float acc = fmaf(__bfloat162float(ba), mu0, 0.f);      // same order as the generic loop below
            acc = fmaf(__bfloat162float(bb), mu1, acc);
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
            if (lane == 0) {
                const float bo = bias[(static_cast<size_t>(which) * H + h) * PRJ_D + o];
                const size_t bi = ((static_cast<size_t>(which) * B + b) * H + h) * PRJ_D + o;
                if (which == 2) {
                    bf[bi] = -acc;
                    mu_v[static_cast<size_t>(b) * C + h * PRJ_D + o] = acc + bo;
                } else {
                    bf[bi] = bo * gain - acc;
                }
            }
        }
        return;
    }
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

// grid (H, batch of this part, number of projections).
// which = which0 + blockIdx.z: 0 = f (content side, batch B), 1 = g, 2 = h (style side, batch Bs).  B here is the
// batch STRIDE of the folded-weight workspace ([3][B][H][d][d]), the same for both sides.
__global__ void __launch_bounds__(256) fold_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                                   const float* __restrict__ mean_c, const float* __restrict__ rstd_c,
                                                   const float* __restrict__ mean_s, const float* __restrict__ rstd_s,
                                                   int which0, int B, int H, int d, __nv_bfloat16* __restrict__ wf,
                                                   float* __restrict__ bf, float* __restrict__ mu_v) {
    const int h = blockIdx.x, b = blockIdx.y, which = which0 + blockIdx.z;
    const int C = H * d;
    const float* mu = (which == 0 ? mean_c : mean_s) + static_cast<size_t>(b) * C + h * d;
    const float* rs = (which == 0 ? rstd_c : rstd_s) + static_cast<size_t>(b) * C + h * d;
    fold_head(w, bias, mu, rs, which, b, h, B, H, d, wf, bf, mu_v);
}

// Second pass of the instance-norm statistics FUSED with the fold (the layer path; saves a launch and a round trip
// per layer): block (h, b, role) finishes mean / rstd of its head's 64 channels from the per-split partial sums of
// stats_partial_kernel -- same arithmetic as stats_final_kernel (double, pivot = first token) -- publishes them, and
// folds them into its projection.  Roles g and h both finish the fs statistics (64 channels x <= 32 splits: cheaper
// than a hand-over); role kind 3 only publishes (fcs: consumed by the attention epilogue).  ti < 0: the statistics
// are already in mean / rstd (MHADA_REUSE_FS_STATS).
__global__ void __launch_bounds__(256) fold_stats_kernel(const FoldStatsJob job, const float* __restrict__ partial,
                                                         const float* __restrict__ w, const float* __restrict__ bias, int B,
                                                         int H, int d, __nv_bfloat16* __restrict__ wf, float* __restrict__ bf,
                                                         float* __restrict__ mu_v) {
    // d = head_dim: 64 or 128 (a power of two <= 256 in general); 256 / d groups of d threads share the splits
    __shared__ float mu_s[256], rs_s[256];
    __shared__ double part_a[256], part_q[256];                        // [group][channel], group-major
    const int h = blockIdx.x, b = blockIdx.y, role = blockIdx.z;
    const int C = H * d;
    const int kind = job.kind[role], ti = job.ti[role];
    const int ch = threadIdx.x & (d - 1), grp = threadIdx.x / d, ngrp = 256 / d;
    if (ti >= 0) {
        // split s goes to group s % 4; the per-group sums are then added in group order 0..3.  (The order differs
        // from stats_final_kernel's sequential loop only in double precision: the float results are the same
        // unless a sum lands within 1e-16 relative of a rounding boundary.)
        const int nsplit = job.splits[role];
        const int c = h * d + ch;
        double a = 0.0, q = 0.0;
        const float* p0 = partial + ((static_cast<size_t>(ti) * B + b) * job.max_splits * C + c) * 2;
#pragma unroll 8
        for (int sp = grp; sp < nsplit; sp += ngrp) {
            const float2 v = __ldg(reinterpret_cast<const float2*>(p0 + static_cast<size_t>(sp) * C * 2));
            a += v.x;
            q += v.y;
        }
        part_a[grp * d + ch] = a;
        part_q[grp * d + ch] = q;
    }
    __syncthreads();
    if (threadIdx.x < d) {
        const int c = h * d + threadIdx.x;
        const size_t r = static_cast<size_t>(b) * C + c;
        float m, rs;
        if (ti >= 0) {
            const int N = job.N[role];
            double a = part_a[ch], q = part_q[ch];
            for (int gI = 1; gI < ngrp; ++gI) {
                a += part_a[gI * d + ch];
                q += part_q[gI * d + ch];
            }
            const double piv = __bfloat162float(static_cast<const __nv_bfloat16*>(job.x[role])[static_cast<size_t>(b) * N * C + c]);
            const double mm = a / N;
            double var = q / N - mm * mm;
            if (var < 0.0) var = 0.0;
            m = static_cast<float>(piv + mm);
            rs = static_cast<float>(1.0 / sqrt(var + static_cast<double>(1e-5f)));   // nn.InstanceNorm2d eps, as stats.cu
            if (kind != 2) {                       // roles g and h compute the same numbers: one of them writes
                job.mean[role][r] = m;
                job.rstd[role][r] = rs;
            }
        } else {
            m = job.mean[role][r];
            rs = job.rstd[role][r];
        }
        mu_s[threadIdx.x] = m;
        rs_s[threadIdx.x] = rs;
    }
    __syncthreads();
    if (kind < 3) fold_head(w, bias, mu_s, rs_s, kind, b, h, B, H, d, wf, bf, mu_v);
}

// Persistent projection kernel: one launch covers the content side (Q from fc) and the style side (K and V' from
// the same fs tile).  A work item is one (128-token tile, head); heads are the fastest index, so the CTAs that run
// side by side read and write the eight 128-byte head slices of the same token rows (whole 1 KB rows in DRAM).
// Style items (three output tiles) come first, content items (one) after, dealt round-robin to the CTAs.
//   warp 0      TMA producer: X tile [128 x 64] + one or two folded 64 x 64 weight tiles per item, 3-stage ring
//   warp 1      tcgen05.mma issuer; accumulators double-buffered in TMEM (2 x 128 columns), so the MMAs of item
//               n+1 run while the epilogue warps drain item n
//   warps 2-5   epilogue: TMEM -> registers, + folded bias, bf16, 256-bit stores (V' also stores the squares)
// Two CTAs per SM (96 KB of shared memory, 256 TMEM columns each).
// KCH = head_dim / 64: a head's d x d projection is done as KCH output blocks of 64 channels (one per work item),
// each contracting the KCH input chunks of the token tile with KCH 64 x 64 weight blocks.
constexpr int PRJ_STAGES = 3;
constexpr uint32_t PRJ_A_BYTES = PRJ_BM * PRJ_D * 2, PRJ_W_BYTES = PRJ_D * PRJ_D * 2;
template <int KCH> constexpr uint32_t prj_stage_bytes() { return KCH * (PRJ_A_BYTES + 2 * PRJ_W_BYTES); }

struct ProjParams {
    const float* bfold;              // [3][Bw][H][64] folded biases
    __nv_bfloat16 *q, *k, *v;
    int Bw, H, Bc, Nc, Bs, Ns;       // Bc / Bs = 0 when that side is not projected
    int tiles_c, tiles_s, items_s, items;
};

struct ProjItem {
    int side, b, h, n0;              // side 0 = content (Q), 1 = style (K, V'); h = output block = head * KCH + block
};

__device__ __forceinline__ ProjItem proj_item(const ProjParams& p, int idx) {
    ProjItem it;                     // p.H counts 64-channel output blocks (heads * KCH)
    it.side = idx < p.items_s ? 1 : 0;
    const int r = it.side ? idx : idx - p.items_s;
    const int tiles = it.side ? p.tiles_s : p.tiles_c;
    it.h = r % p.H;
    it.n0 = ((r / p.H) % tiles) * PRJ_BM;
    it.b = r / (p.H * tiles);
    return it;
}

template <int KCH>
__global__ void __launch_bounds__(PRJ_THREADS, KCH == 1 ? 2 : 1)
proj_tc_kernel(const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmS,
               const __grid_constant__ CUtensorMap tmW, const ProjParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[PRJ_STAGES], empty[PRJ_STAGES], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float bias_s[2][2][PRJ_D];        // [accumulator buffer][output][channel]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t PRJ_STAGE_BYTES = prj_stage_bytes<KCH>();
    constexpr uint32_t X_BYTES = KCH * PRJ_A_BYTES, W_BYTES = KCH * PRJ_W_BYTES;   // per stage: X | W(out 0) | W(out 1)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmC);
        tma_prefetch_desc(&tmS);
        tma_prefetch_desc(&tmW);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < PRJ_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
            for (int u = 0; u < 2; ++u) { mbar_init(&acc_full[u], 1); mbar_init(&acc_empty[u], 4); }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(&tmem_slot, 256);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            int n = 0;
            for (int idx = blockIdx.x; idx < p.items; idx += gridDim.x, ++n) {
                const ProjItem it = proj_item(p, idx);
                const int s = n % PRJ_STAGES;
                mbar_wait(&empty[s], ((n / PRJ_STAGES) & 1) ^ 1);
                uint8_t* st = smem + s * PRJ_STAGE_BYTES;
                mbar_arrive_expect_tx(&full[s], X_BYTES + (it.side ? 2 : 1) * W_BYTES);
                const int head = it.h / KCH;
                const int which0 = it.side ? 1 : 0;
#pragma unroll
                for (int c = 0; c < KCH; ++c) {
                    // input chunk c of the head; weight block (output block it.h, input chunk c): rows of W' are output
                    // channels ([3][Bw][heads * KCH * 64] rows in all), columns input channels of the head
                    tma_load_3d(st + c * PRJ_A_BYTES, it.side ? &tmS : &tmC, &full[s], (head * KCH + c) * PRJ_D, it.n0, it.b);
                    tma_load_2d(st + X_BYTES + c * PRJ_W_BYTES, &tmW, &full[s], c * PRJ_D,
                                ((which0 * p.Bw + it.b) * p.H + it.h) * PRJ_D);
                    if (it.side)
                        tma_load_2d(st + X_BYTES + W_BYTES + c * PRJ_W_BYTES, &tmW, &full[s], c * PRJ_D,
                                    ((2 * p.Bw + it.b) * p.H + it.h) * PRJ_D);
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(PRJ_BM, PRJ_D, 0, 0);
            int n = 0;
            for (int idx = blockIdx.x; idx < p.items; idx += gridDim.x, ++n) {
                const int side = idx < p.items_s ? 1 : 0;
                const int s = n % PRJ_STAGES, u = n & 1;
                mbar_wait(&full[s], (n / PRJ_STAGES) & 1);
                mbar_wait(&acc_empty[u], ((n >> 1) & 1) ^ 1);       // epilogue has drained this accumulator buffer
                tc_fence_after();
                const uint32_t a_addr = smem_u32(smem + s * PRJ_STAGE_BYTES);
                const uint64_t da = make_smem_desc(a_addr, 16, 1024);
                for (int t = 0; t <= side; ++t) {
                    const uint64_t db = make_smem_desc(a_addr + X_BYTES + t * W_BYTES, 16, 1024);
#pragma unroll
                    for (int c = 0; c < KCH; ++c)
#pragma unroll
                        for (int k = 0; k < PRJ_D / 16; ++k)
                            umma_ss(tmem + u * 128 + t * 64, desc_advance(da, c * PRJ_A_BYTES + k * 32),
                                    desc_advance(db, c * PRJ_W_BYTES + k * 32), idesc, (c | k) != 0);
                }
                umma_commit(&empty[s]);         // the stage may be refilled once these MMAs have read it
                umma_commit(&acc_full[u]);
            }
        }
    } else {
        const int quarter = warp & 3;
        const int et = threadIdx.x - 64;            // 0 .. 127 among the epilogue threads
        const int C = p.H * PRJ_D;
        int n = 0;
        for (int idx = blockIdx.x; idx < p.items; idx += gridDim.x, ++n) {
            const ProjItem it = proj_item(p, idx);
            const int u = n & 1;
            const int nout = it.side ? 2 : 1;
            const int N = it.side ? p.Ns : p.Nc;
            // bias_s[u] was last read two items ago; every epilogue thread has passed the barrier of item n-1 since
            if (et < nout * PRJ_D) {
                const int t = et / PRJ_D, o = et % PRJ_D;
                bias_s[u][t][o] = p.bfold[((static_cast<size_t>(it.side ? 1 + t : 0) * p.Bw + it.b) * p.H + it.h) * PRJ_D + o];
            }
            named_bar_sync(1, 128);
            mbar_wait(&acc_full[u], (n >> 1) & 1);
            tc_fence_after();
            const int row = it.n0 + quarter * 32 + lane;
            for (int t = 0; t < nout; ++t) {
                const bool is_v = it.side && t == 1;
#pragma unroll
                for (int c = 0; c < PRJ_D; c += 32) {
                    uint32_t r[32];
                    tmem_ld_x32(tmem_addr(tmem, quarter * 32, u * 128 + t * 64 + c), r);
                    tmem_wait_ld();
                    if (row < N) {
                        uint32_t lo[16], sq[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            float v0 = __uint_as_float(r[2 * i]) + bias_s[u][t][c + 2 * i];
                            float v1 = __uint_as_float(r[2 * i + 1]) + bias_s[u][t][c + 2 * i + 1];
                            lo[i] = pack_bf16x2(v0, v1);
                            sq[i] = pack_bf16x2(v0 * v0, v1 * v1);
                        }
                        if (!is_v) {
                            __nv_bfloat16* o = (it.side ? p.k : p.q) + (static_cast<size_t>(it.b) * N + row) * C + it.h * PRJ_D + c;
                            st_global_256(o, lo);
                            st_global_256(o + 16, lo + 8);
                        } else {
                            __nv_bfloat16* o = p.v + (static_cast<size_t>(it.b) * N + row) * (2 * C) + it.h * (2 * PRJ_D) + c;
                            st_global_256(o, lo);
                            st_global_256(o + 16, lo + 8);
                            st_global_256(o + PRJ_D, sq);
                            st_global_256(o + PRJ_D + 16, sq + 8);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[u]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 256);
}

static int make_x_map(CUtensorMap* tm, const void* x, int B, int N, int C) {
    uint64_t dims[3] = {static_cast<uint64_t>(C), static_cast<uint64_t>(N), static_cast<uint64_t>(B)};
    uint64_t str[2] = {static_cast<uint64_t>(C) * 2, static_cast<uint64_t>(N) * C * 2};
    uint32_t box[3] = {PRJ_D, PRJ_BM, 1};
    return make_tmap_bf16(tm, x, 3, dims, str, box);
}

// the persistent projection kernel on weights that are already folded (in `ws`)
int launch_proj_bf16_folded(int parts, const void* fc, const void* fs, int B, int Bs, int Nc, int Ns, int H, int d, void* q,
                            void* k, void* v, void* ws, cudaStream_t s) {
    // parts: 1 = Q from fc (batch B), 2 = K and V' from fs (batch Bs).  Bw = batch stride of the folded weights.
    const int Bw = B > Bs ? B : Bs;
    __nv_bfloat16* wf = static_cast<__nv_bfloat16*>(ws);
    float* bf = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + fold_w_bytes(Bw, H, d));
    const int C = H * d;
    const int kch = d / PRJ_D;
    CUtensorMap tmW;
    uint64_t dimsW[2] = {static_cast<uint64_t>(d), static_cast<uint64_t>(3) * Bw * H * d};
    uint64_t strW[1] = {static_cast<uint64_t>(d) * 2};
    uint32_t boxW[2] = {PRJ_D, PRJ_D};
    if (int e = make_tmap_bf16(&tmW, wf, 2, dimsW, strW, boxW)) return e;
    const bool do_c = parts & 1, do_s = parts & 2;
    CUtensorMap tmC, tmS;
    if (int e = make_x_map(&tmC, do_c ? fc : fs, do_c ? B : Bs, do_c ? Nc : Ns, C)) return e;
    if (int e = make_x_map(&tmS, do_s ? fs : fc, do_s ? Bs : B, do_s ? Ns : Nc, C)) return e;
    ProjParams p;
    p.bfold = bf;
    p.q = static_cast<__nv_bfloat16*>(q); p.k = static_cast<__nv_bfloat16*>(k); p.v = static_cast<__nv_bfloat16*>(v);
    p.Bw = Bw; p.H = H * kch;                      // 64-channel output blocks
    p.Bc = do_c ? B : 0; p.Nc = Nc; p.Bs = do_s ? Bs : 0; p.Ns = Ns;
    p.tiles_c = (Nc + PRJ_BM - 1) / PRJ_BM;
    p.tiles_s = (Ns + PRJ_BM - 1) / PRJ_BM;
    p.items_s = p.Bs * p.H * p.tiles_s;
    p.items = p.items_s + p.Bc * p.H * p.tiles_c;
    const size_t smem = PRJ_STAGES * (kch == 1 ? prj_stage_bytes<1>() : prj_stage_bytes<2>()) + 1024;
    static DeviceOnce once1, once2;
    if (kch == 1) {
        if (int e = smem_attr_once(once1, reinterpret_cast<const void*>(proj_tc_kernel<1>), smem, "proj smem attr")) return e;
    } else {
        if (int e = smem_attr_once(once2, reinterpret_cast<const void*>(proj_tc_kernel<2>), smem, "proj smem attr")) return e;
    }
    const int n_sm = sm_count();
    const int ctas = (kch == 1 ? 2 : 1) * n_sm;
    const int grid = p.items < ctas ? p.items : ctas;
    if (kch == 1)
        proj_tc_kernel<1><<<grid, PRJ_THREADS, smem, s>>>(tmC, tmS, tmW, p);
    else
        proj_tc_kernel<2><<<grid, PRJ_THREADS, smem, s>>>(tmC, tmS, tmW, p);
    count_launch();
    return check_cuda(cudaGetLastError(), "proj_tc launch");
}

int launch_fold_stats(const FoldStatsJob& job, const float* partial, const float* w, const float* bias, int B, int H,
                      int d, float* mu_v, void* proj_ws, cudaStream_t s) {
    __nv_bfloat16* wf = static_cast<__nv_bfloat16*>(proj_ws);
    float* bf = reinterpret_cast<float*>(static_cast<uint8_t*>(proj_ws) + fold_w_bytes(B, H, d));
    fold_stats_kernel<<<dim3(H, B, job.n_roles), 256, 0, s>>>(job, partial, w, bias, B, H, d, wf, bf, mu_v);
    count_launch();
    return check_cuda(cudaGetLastError(), "fold_stats launch");
}

int launch_proj_bf16(int parts, const void* fc, const void* fs, const float* mean_c, const float* rstd_c,
                     const float* mean_s, const float* rstd_s, const float* w, const float* bias, int B, int Bs, int Nc,
                     int Ns, int H, int d, void* q, void* k, void* v, float* mu_v, void* ws, cudaStream_t s) {
    const int Bw = B > Bs ? B : Bs;
    __nv_bfloat16* wf = static_cast<__nv_bfloat16*>(ws);
    float* bf = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + fold_w_bytes(Bw, H, d));
    const bool do_c = parts & 1, do_s = parts & 2;
    // fold the instance norm into the per-image weights: one launch when both sides have the same batch
    if (do_c && do_s && B == Bs) {
        fold_kernel<<<dim3(H, B, 3), 256, 0, s>>>(w, bias, mean_c, rstd_c, mean_s, rstd_s, 0, Bw, H, d, wf, bf, mu_v);
        count_launch();
    } else {
        if (do_c) {
            fold_kernel<<<dim3(H, B, 1), 256, 0, s>>>(w, bias, mean_c, rstd_c, mean_s, rstd_s, 0, Bw, H, d, wf, bf, mu_v);
            count_launch();
        }
        if (do_s) {
            fold_kernel<<<dim3(H, Bs, 2), 256, 0, s>>>(w, bias, mean_c, rstd_c, mean_s, rstd_s, 1, Bw, H, d, wf, bf, mu_v);
            count_launch();
        }
    }
    return launch_proj_bf16_folded(parts, fc, fs, B, Bs, Nc, Ns, H, d, q, k, v, ws, s);
}

}  // namespace mh

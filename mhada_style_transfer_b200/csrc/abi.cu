// extern "C" surface of libmhada_b200.so (declared in include/mhada_b200.h): argument checks,
// workspace carving and the per-layer launch sequence.  No torch types, no allocation, no syncs.
#include <vector>

#include "common.h"

using namespace mh;

namespace {

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; }

#define REQUIRE(cond, code, ...)      \
    do {                              \
        if (!(cond)) {                \
            set_error(__VA_ARGS__);   \
            return code;              \
        }                             \
    } while (0)

size_t esize(int dtype) { return dtype == MHADA_BF16 ? 2 : 4; }

struct LayerWs {
    float *mean_c, *rstd_c, *mean_s, *rstd_s, *mean_x, *rstd_x, *mu_v;
    void *stats_ws, *proj_ws, *q, *k, *v, *heads, *lin_ws, *wide, *cos_ws;
    size_t stats_bytes, proj_bytes, lin_bytes, wide_bytes, cos_bytes, total;
};

LayerWs carve(int dtype, int B, int Nc, int Ns, int C, int H, uint8_t* base) {
    LayerWs w;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void* p = base ? base + off : nullptr;
        off += align_up(bytes, 256);
        return p;
    };
    const size_t sc = static_cast<size_t>(B) * C * sizeof(float);
    w.mean_c = static_cast<float*>(take(sc)); w.rstd_c = static_cast<float*>(take(sc));
    w.mean_s = static_cast<float*>(take(sc)); w.rstd_s = static_cast<float*>(take(sc));
    w.mean_x = static_cast<float*>(take(sc)); w.rstd_x = static_cast<float*>(take(sc));
    w.mu_v = static_cast<float*>(take(sc));
    const int nmax = Nc > Ns ? Nc : Ns;
    w.stats_bytes = stats_workspace(B, nmax, C);
    // the split count depends on N; size for the larger of the two token counts and the smaller one too
    size_t sb2 = stats_workspace(B, Nc < Ns ? Nc : Ns, C);
    if (sb2 > w.stats_bytes) w.stats_bytes = sb2;
    w.stats_ws = take(w.stats_bytes);
    const int d = C / H;
    const bool wide = dtype == MHADA_BF16 && d > 128;      // wide heads: projections + materialised attention (forloss_tc.cu)
    w.proj_bytes = dtype == MHADA_BF16 && !wide ? proj_bf16_workspace(B, H, d) : 0;
    w.proj_ws = take(w.proj_bytes);
    const size_t e = esize(dtype);
    w.q = take(wide ? 0 : static_cast<size_t>(B) * Nc * C * e);
    w.k = take(wide ? 0 : static_cast<size_t>(B) * Ns * C * e);
    w.v = take(wide ? 0 : static_cast<size_t>(B) * Ns * C * e * (dtype == MHADA_BF16 ? 2 : 1));
    w.wide_bytes = wide ? layer_wide_workspace(B, Nc, Ns, C, H) : 0;
    w.wide = take(w.wide_bytes);
    w.cos_bytes = dtype == MHADA_F32 && d == 64 ? attn_cosine_scratch_bytes(B, H) : 0;      // closed-form cosine moments
    w.cos_ws = take(w.cos_bytes);
    w.heads = take(static_cast<size_t>(B) * Nc * C * e);
    w.lin_bytes = dtype == MHADA_BF16 ? linear_bf16_workspace(C, C) : 0;
    w.lin_ws = take(w.lin_bytes);
    w.total = off;
    return w;
}

// ---- optional event bracketing of the stages of a layer (mhada_profile_*): statistics, projections, attention,
//      out_conv, each between a pair of events recorded on the launching stream
struct Profiler {
    bool on = false;
    std::vector<cudaEvent_t> ev;   // pairs: start, stop
    std::vector<int> stage;        // stage of pair i
    size_t used = 0;               // events in use
    float ms[MHADA_STAGE_COUNT] = {};
    int n[MHADA_STAGE_COUNT] = {};
};
thread_local Profiler g_prof;

struct StageTimer {
    cudaEvent_t e1 = nullptr;
    cudaStream_t s;
    StageTimer(int stage, cudaStream_t stream) : s(stream) {
        if (!g_prof.on) return;
        while (g_prof.used + 2 > g_prof.ev.size()) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) return;
            g_prof.ev.push_back(e);
        }
        if (g_prof.stage.size() < g_prof.ev.size() / 2) g_prof.stage.resize(g_prof.ev.size() / 2);
        g_prof.stage[g_prof.used / 2] = stage;
        cudaEventRecord(g_prof.ev[g_prof.used], s);
        e1 = g_prof.ev[g_prof.used + 1];
        g_prof.used += 2;
    }
    ~StageTimer() {
        if (e1) cudaEventRecord(e1, s);
    }
};

int attn_dispatch(const mhada_attn_args& a, cudaStream_t s) {
    StageTimer timer(MHADA_STAGE_ATTN, s);
    return a.dtype == MHADA_BF16 ? launch_attn_bf16(a, s) : launch_attn_f32(a, s);
}

struct StyleCacheView {
    void *k, *v;
    float* mu_v;
    size_t total;
};

StyleCacheView carve_cache(int dtype, int Bs, int Ns, int C, uint8_t* base) {
    StyleCacheView c;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void* p = base ? base + off : nullptr;
        off += align_up(bytes, 256);
        return p;
    };
    const size_t e = esize(dtype);
    c.k = take(static_cast<size_t>(Bs) * Ns * C * e);
    c.v = take(static_cast<size_t>(Bs) * Ns * C * e * (dtype == MHADA_BF16 ? 2 : 1));
    c.mu_v = static_cast<float*>(take(static_cast<size_t>(Bs) * C * sizeof(float)));
    c.total = off;
    return c;
}

int attn_dispatch(const mhada_attn_args& a, cudaStream_t s);

// One MHAda layer.  cache == nullptr: the style side (fs statistics, K, V, mu_v) is computed into the workspace;
// otherwise it is read from `cache` (style batch Bs = 1 or B) and fs is not touched.
int layer_impl(const char* who, int dtype, const void* fc, const void* fs, const void* fcs, const void* cache, int Bs,
               const float* w_fgh, const float* b_fgh, const float* w_out, const float* b_out, int B, int Nc, int Ns,
               int C, int H, int flags, void* out, void* ws, size_t ws_bytes, mhada_stream_t stream) {
    g_launches = 0;
    REQUIRE(fc && fcs && w_fgh && b_fgh && out && ws, MHADA_ERR_ARG, "%s: null pointer", who);
    REQUIRE((w_out == nullptr) == (b_out == nullptr), MHADA_ERR_ARG, "%s: w_out/b_out must be given together", who);
    REQUIRE(B > 0 && Nc > 0 && Ns > 0 && C > 0 && H > 0, MHADA_ERR_ARG, "%s: bad sizes", who);
    REQUIRE(C % H == 0, MHADA_ERR_ARG, "%s: C=%d not divisible by H=%d", who, C, H);
    REQUIRE(dtype == MHADA_F32 || dtype == MHADA_BF16, MHADA_ERR_ARG, "%s: bad dtype %d", who, dtype);
    REQUIRE(out != fc && out != fs && out != fcs, MHADA_ERR_ARG, "%s: out must not alias an input", who);
    const int d = C / H;
    const bool wide = dtype == MHADA_BF16 && d > 128;
    if (dtype == MHADA_BF16)
        REQUIRE(d == 64 || d == 128 || (d % 128 == 0 && !cache), MHADA_ERR_UNSUPPORTED,
                "%s: the bf16 tensor-core path implements head_dim 64, 128 and (without a style cache) multiples of 128 "
                "(C/H = %d); use MHADA_F32", who, d);
    REQUIRE(!(flags & MHADA_WOUT_BF16) || (dtype == MHADA_BF16 && w_out && C % 128 == 0 && aligned16(w_out)), MHADA_ERR_ARG,
            "%s: MHADA_WOUT_BF16 needs the MHADA_BF16 path, a 16-byte aligned bf16 w_out and C %% 128 == 0", who);
    REQUIRE(!(flags & MHADA_LAYER_COSINE) || dtype == MHADA_F32, MHADA_ERR_UNSUPPORTED,
            "%s: the cosine activation (adaDecoder.py:20-34) runs on the MHADA_F32 path only", who);
    REQUIRE(aligned16(fc) && (!fs || aligned16(fs)) && aligned16(fcs) && aligned32(out) && aligned32(ws), MHADA_ERR_ARG,
            "%s: inputs must be 16-byte aligned, out and ws 32-byte aligned", who);
    if (dtype == MHADA_BF16) REQUIRE(C % 16 == 0, MHADA_ERR_ARG, "%s: bf16 path needs C %% 16 == 0", who);
    REQUIRE(C % (dtype == MHADA_BF16 ? 8 : 4) == 0, MHADA_ERR_ARG, "%s: C must be a multiple of %d", who,
            dtype == MHADA_BF16 ? 8 : 4);
    LayerWs w = carve(dtype, B, Nc, Ns, C, H, static_cast<uint8_t*>(ws));
    REQUIRE(ws_bytes >= w.total, MHADA_ERR_WORKSPACE, "%s: workspace %zu < %zu", who, ws_bytes, w.total);
    if (int e = device_check()) return e;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    StyleCacheView cv{};
    if (cache) cv = carve_cache(dtype, Bs, Ns, C, static_cast<uint8_t*>(const_cast<void*>(cache)));

    // (1) statistics of fc, fs and (unless it is the same tensor) fcs, one launch   adaDecoder.py:173,178,198
    //     MHADA_REUSE_FS_STATS: fs (and this workspace) are the ones of the previous call, so mean_s / rstd_s
    //     are still valid -- the two layers of a level share fs (adaDecoder.py:264-265)
    const float *mean_x = w.mean_c, *rstd_x = w.rstd_c;
    if (fcs != fc) {
        mean_x = w.mean_x;
        rstd_x = w.rstd_x;
    }
    const int parts = cache ? MHADA_PROJ_Q : (MHADA_PROJ_Q | MHADA_PROJ_KV);
    const bool fs_stats = !cache && !(flags & MHADA_REUSE_FS_STATS);
    if (wide) {
        // wide heads (head_dim 256 / 512): full statistics, per-head projections on the token GEMM, materialised attention
        {
            const void* xs[3];
            float* ms[3];
            float* rs[3];
            int ns[3];
            int n = 0;
            xs[n] = fc; ms[n] = w.mean_c; rs[n] = w.rstd_c; ns[n] = Nc; ++n;
            if (fs_stats) { xs[n] = fs; ms[n] = w.mean_s; rs[n] = w.rstd_s; ns[n] = Ns; ++n; }
            if (fcs != fc) { xs[n] = fcs; ms[n] = w.mean_x; rs[n] = w.rstd_x; ns[n] = Nc; ++n; }
            StageTimer timer(MHADA_STAGE_STATS, s);
            if (int e = launch_stats_multi(n, xs, ns, ms, rs, dtype, B, C, C, static_cast<float*>(w.stats_ws), s)) return e;
        }
        {
            StageTimer timer(MHADA_STAGE_ATTN, s);
            if (int e = layer_wide_attention(fc, fs, fcs, w.mean_c, w.rstd_c, w.mean_s, w.rstd_s, mean_x, rstd_x, w_fgh, b_fgh, B, Nc,
                                             Ns, C, H, w_out ? w.heads : out, w.wide, s))
                return e;
        }
    } else if (dtype == MHADA_BF16) {
        // bf16 path: first pass of the statistics, then ONE kernel that finishes them and folds them into the
        // projection weights (fold_stats_kernel), then the persistent projection kernel
        const void* xs[3];
        int ns[3];
        int n = 0, ti_s = -1, ti_x = -1;
        xs[n] = fc; ns[n] = Nc; ++n;
        if (fs_stats) { ti_s = n; xs[n] = fs; ns[n] = Ns; ++n; }
        if (fcs != fc) { ti_x = n; xs[n] = fcs; ns[n] = Nc; ++n; }
        StatsPartialInfo info{};
        {
            StageTimer timer(MHADA_STAGE_STATS, s);
            if (int e = launch_stats_partial(n, xs, ns, dtype, B, C, C, static_cast<float*>(w.stats_ws), &info, s)) return e;
        }
        StageTimer timer(MHADA_STAGE_PROJ, s);
        FoldStatsJob job{};
        job.max_splits = info.max_splits;
        auto role = [&](int kind, int ti, const void* x, int N, float* mean, float* rstd) {
            const int r = job.n_roles++;
            job.kind[r] = kind; job.ti[r] = ti; job.x[r] = x; job.N[r] = N;
            job.splits[r] = ti >= 0 ? info.splits[ti] : 0;
            job.mean[r] = mean; job.rstd[r] = rstd;
        };
        role(0, 0, fc, Nc, w.mean_c, w.rstd_c);
        if (!cache) {
            role(1, ti_s, fs, Ns, w.mean_s, w.rstd_s);
            role(2, ti_s, fs, Ns, w.mean_s, w.rstd_s);
        }
        if (fcs != fc) role(3, ti_x, fcs, Nc, w.mean_x, w.rstd_x);
        if (int e = launch_fold_stats(job, static_cast<const float*>(w.stats_ws), w_fgh, b_fgh, B, H, d, w.mu_v, w.proj_ws, s))
            return e;
        if (int e = launch_proj_bf16_folded(parts, fc, fs, B, cache ? 0 : B, Nc, Ns, H, d, w.q, w.k, w.v, w.proj_ws, s))
            return e;
    } else {
        {
            const void* xs[3];
            float* ms[3];
            float* rs[3];
            int ns[3];
            int n = 0;
            xs[n] = fc; ms[n] = w.mean_c; rs[n] = w.rstd_c; ns[n] = Nc; ++n;
            if (fs_stats) { xs[n] = fs; ms[n] = w.mean_s; rs[n] = w.rstd_s; ns[n] = Ns; ++n; }
            if (fcs != fc) { xs[n] = fcs; ms[n] = w.mean_x; rs[n] = w.rstd_x; ns[n] = Nc; ++n; }
            StageTimer timer(MHADA_STAGE_STATS, s);
            if (int e = launch_stats_multi(n, xs, ns, ms, rs, dtype, B, C, C, static_cast<float*>(w.stats_ws), s)) return e;
        }
        StageTimer timer(MHADA_STAGE_PROJ, s);
        if (int e = launch_proj_f32(parts, static_cast<const float*>(fc), static_cast<const float*>(fs), w.mean_c, w.rstd_c,
                                    w.mean_s, w.rstd_s, w_fgh, b_fgh, B, cache ? 0 : B, Nc, Ns, H, d, static_cast<float*>(w.q),
                                    static_cast<float*>(w.k), static_cast<float*>(w.v), w.mu_v, s))
            return e;
    }
    // (3) attention + fused epilogue                                         adaDecoder.py:186-198
    mhada_attn_args a{};
    if (!wide) {
    a.dtype = dtype; a.B = B; a.H = H; a.Nc = Nc; a.Ns = Ns; a.dqk = d; a.dv = d;
    a.q = w.q; a.k = cache ? cv.k : w.k; a.v = cache ? cv.v : w.v; a.x = fcs;
    a.out = w_out ? w.heads : out;
    a.ldq = C; a.ldk = C; a.ldv = dtype == MHADA_BF16 ? 2 * C : C; a.ldx = C; a.ldo = C;
    a.x_mean = mean_x; a.x_rstd = rstd_x; a.mu_v = cache ? cv.mu_v : w.mu_v;
    a.kv_batch = cache ? Bs : B;
    a.activation = (flags & MHADA_LAYER_COSINE) ? MHADA_ACT_COSINE : MHADA_ACT_SOFTMAX;
    a.scratch = w.cos_ws; a.scratch_bytes = w.cos_bytes;
    if (int e = attn_dispatch(a, s)) return e;
    }
    // (4) out_conv                                                           adaDecoder.py:202-205
    if (w_out) {
        StageTimer timer(MHADA_STAGE_LINEAR, s);
        if (dtype == MHADA_BF16 && (flags & MHADA_WOUT_BF16)) {
            // the caller keeps a bf16 copy of out_conv.weight: no per-call conversion launch
            GemmDesc g{};
            g.a = w.heads; g.lda = C; g.w = w_out; g.ldw = C; g.bias = b_out; g.M = B * Nc; g.N = C; g.K = C;
            g.out_bf16 = out; g.ldo = C;
            if (int e = launch_gemm_bf16(g, s)) return e;
        } else if (dtype == MHADA_BF16) {
            if (int e = launch_linear_bf16(w.heads, C, w_out, b_out, B * Nc, C, C, out, C, w.lin_ws, s)) return e;
        } else {
            if (int e = launch_linear_f32(static_cast<const float*>(w.heads), C, w_out, b_out, B * Nc, C, C,
                                          static_cast<float*>(out), C, s))
                return e;
        }
    }
    return 0;
}

// ---- backward of a layer: workspace = the forward workspace (recompute) + the gradient intermediates
struct BwdWs {
    void* fwd;                 // LayerWs region (MHADA_BF16)
    size_t fwd_bytes;
    void *heads, *woT, *wbd, *d_o, *dq, *dk, *dv, *tA, *tB, *partial, *sums, *splitk;
    size_t splitk_bytes;
    float *dcat, *lse, *delta, *dxhat, *gq, *gk, *gv, *dwfull;
    int Mpad_c, Mpad_s;
    size_t total;
};

BwdWs carve_bwd(int B, int Nc, int Ns, int C, int H, uint8_t* base) {
    BwdWs w;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void* p = base ? base + off : nullptr;
        off += align_up(bytes, 256);
        return p;
    };
    w.fwd_bytes = carve(MHADA_BF16, B, Nc, Ns, C, H, nullptr).total;
    w.fwd = take(w.fwd_bytes);
    const size_t Mc = static_cast<size_t>(B) * Nc, Ms = static_cast<size_t>(B) * Ns;
    w.Mpad_c = static_cast<int>((Mc + 63) / 64 * 64);
    w.Mpad_s = static_cast<int>((Ms + 63) / 64 * 64);
    const size_t Mpad = w.Mpad_c > w.Mpad_s ? w.Mpad_c : w.Mpad_s;
    w.heads = take(Mc * C * 2);
    w.woT = take(static_cast<size_t>(C) * C * 2);
    w.wbd = take(static_cast<size_t>(3) * C * C * 2);
    w.dcat = static_cast<float*>(take(Mc * C * 4));
    w.d_o = take(Mc * 4 * C * 2);
    w.lse = static_cast<float*>(take(static_cast<size_t>(B) * H * Nc * 4));
    w.delta = static_cast<float*>(take(static_cast<size_t>(B) * H * Nc * 4));
    w.dxhat = static_cast<float*>(take(Mc * C * 4));
    w.dq = take(Mc * C * 2);
    w.dk = take(Ms * C * 2);
    w.dv = take(Ms * C * 2);
    w.gq = static_cast<float*>(take(Mc * C * 4));
    w.gk = static_cast<float*>(take(Ms * C * 4));
    w.gv = static_cast<float*>(take(Ms * C * 4));
    w.tA = take(static_cast<size_t>(C) * Mpad * 2);
    w.tB = take(static_cast<size_t>(C) * Mpad * 2);
    w.dwfull = static_cast<float*>(take(static_cast<size_t>(C) * C * 4));
    const size_t pb = token_sums_workspace(B, Nc > Ns ? Nc : Ns, C), pb2 = token_sums_workspace(B, Nc < Ns ? Nc : Ns, C);
    w.partial = take(pb > pb2 ? pb : pb2);
    w.sums = take(static_cast<size_t>(B) * C * 8);
    w.splitk_bytes = gemm_splitk_workspace(C, C, static_cast<int>(Mpad));
    const size_t sk2 = gemm_splitk_workspace(C, C, w.Mpad_c < w.Mpad_s ? w.Mpad_c : w.Mpad_s);
    if (sk2 > w.splitk_bytes) w.splitk_bytes = sk2;
    w.splitk = take(w.splitk_bytes);
    w.total = off;
    return w;
}

}  // namespace

extern "C" {

int mhada_profile_begin(void) {
    g_prof.on = true;
    g_prof.used = 0;
    return 0;
}

int mhada_profile_end(float* attn_ms_total, int* attn_launches) {
    g_prof.on = false;
    for (int st = 0; st < MHADA_STAGE_COUNT; ++st) g_prof.ms[st] = 0.f, g_prof.n[st] = 0;
    for (size_t i = 0; i + 1 < g_prof.used; i += 2) {
        if (int e = check_cuda(cudaEventSynchronize(g_prof.ev[i + 1]), "cudaEventSynchronize")) return e;
        float ms = 0.f;
        if (int e = check_cuda(cudaEventElapsedTime(&ms, g_prof.ev[i], g_prof.ev[i + 1]), "cudaEventElapsedTime")) return e;
        const int st = g_prof.stage[i / 2];
        g_prof.ms[st] += ms;
        ++g_prof.n[st];
    }
    g_prof.used = 0;
    if (attn_ms_total) *attn_ms_total = g_prof.ms[MHADA_STAGE_ATTN];
    if (attn_launches) *attn_launches = g_prof.n[MHADA_STAGE_ATTN];
    return 0;
}

int mhada_profile_stage(int stage, float* ms_total, int* brackets) {
    REQUIRE(stage >= 0 && stage < MHADA_STAGE_COUNT, MHADA_ERR_ARG, "mhada_profile_stage: bad stage %d", stage);
    if (ms_total) *ms_total = g_prof.ms[stage];
    if (brackets) *brackets = g_prof.n[stage];
    return 0;
}

int mhada_abi_version(void) { return MHADA_ABI_VERSION; }
size_t mhada_attn_cosine_scratch(int B, int H) { return B > 0 && H > 0 ? attn_cosine_scratch_bytes(B, H) : 0; }
const char* mhada_last_error(void) { return last_error(); }
int mhada_device_check(void) { return device_check(); }
int mhada_last_launch_count(void) { return g_launches; }
long long mhada_total_launch_count(void) { return g_launches_total; }

size_t mhada_in_stats_workspace(int B, int N, int C) {
    if (B <= 0 || N <= 0 || C <= 0) return 0;
    return stats_workspace(B, N, C);
}

int mhada_in_stats(const void* x, int dtype, int B, int N, int C, int ld, float* mean, float* rstd, void* ws,
                   size_t ws_bytes, mhada_stream_t stream) {
    REQUIRE(x && mean && rstd && ws, MHADA_ERR_ARG, "mhada_in_stats: null pointer");
    REQUIRE(B > 0 && N > 0 && C > 0 && ld >= C, MHADA_ERR_ARG, "mhada_in_stats: bad sizes B=%d N=%d C=%d ld=%d", B, N, C, ld);
    REQUIRE(dtype == MHADA_F32 || dtype == MHADA_BF16, MHADA_ERR_ARG, "mhada_in_stats: bad dtype %d", dtype);
    const int vec = dtype == MHADA_BF16 ? 8 : 4;
    REQUIRE(C % vec == 0 && ld % vec == 0 && aligned16(x), MHADA_ERR_ARG,
            "mhada_in_stats: C and ld must be multiples of %d and x 16-byte aligned", vec);
    REQUIRE(ws_bytes >= stats_workspace(B, N, C), MHADA_ERR_WORKSPACE, "mhada_in_stats: workspace %zu < %zu", ws_bytes,
            stats_workspace(B, N, C));
    if (int e = device_check()) return e;
    return launch_stats(x, dtype, B, N, C, ld, mean, rstd, static_cast<float*>(ws), static_cast<cudaStream_t>(stream));
}

size_t mhada_proj_workspace(int B, int H, int d) {
    if (B <= 0 || H <= 0 || d <= 0) return 0;
    return proj_bf16_workspace(B, H, d);
}

int mhada_proj(int dtype, int parts, const void* fc, const void* fs, const float* mean_c, const float* rstd_c,
               const float* mean_s, const float* rstd_s, const float* w, const float* bias, int B, int Bs, int Nc,
               int Ns, int H, int d, void* q, void* k, void* v, float* mu_v, void* ws, size_t ws_bytes,
               mhada_stream_t stream) {
    REQUIRE(parts >= 1 && parts <= 3, MHADA_ERR_ARG, "mhada_proj: parts must be MHADA_PROJ_Q | MHADA_PROJ_KV");
    REQUIRE(w && bias, MHADA_ERR_ARG, "mhada_proj: null weights");
    if (parts & MHADA_PROJ_Q) REQUIRE(fc && mean_c && rstd_c && q && B > 0 && Nc > 0, MHADA_ERR_ARG, "mhada_proj: Q side incomplete");
    if (parts & MHADA_PROJ_KV)
        REQUIRE(fs && mean_s && rstd_s && k && v && mu_v && Bs > 0 && Ns > 0, MHADA_ERR_ARG, "mhada_proj: K/V side incomplete");
    REQUIRE(H > 0 && d > 0, MHADA_ERR_ARG, "mhada_proj: bad sizes");
    if (!(parts & MHADA_PROJ_Q)) B = 0;
    if (!(parts & MHADA_PROJ_KV)) Bs = 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == MHADA_F32) {
        if (int e = device_check()) return e;
        return launch_proj_f32(parts, static_cast<const float*>(fc), static_cast<const float*>(fs), mean_c, rstd_c, mean_s,
                               rstd_s, w, bias, B, Bs, Nc, Ns, H, d, static_cast<float*>(q), static_cast<float*>(k),
                               static_cast<float*>(v), mu_v, s);
    }
    REQUIRE(dtype == MHADA_BF16, MHADA_ERR_ARG, "mhada_proj: bad dtype %d", dtype);
    REQUIRE(d == 64 || d == 128, MHADA_ERR_UNSUPPORTED,
            "mhada_proj: the bf16 tensor-core path implements head_dim 64 and 128, got %d", d);
    REQUIRE(ws && ws_bytes >= proj_bf16_workspace(B > Bs ? B : Bs, H, d), MHADA_ERR_WORKSPACE, "mhada_proj: workspace too small");
    REQUIRE((!(parts & 1) || (aligned16(fc) && aligned32(q))) && (!(parts & 2) || (aligned16(fs) && aligned32(k) && aligned32(v))) &&
                aligned16(ws),
            MHADA_ERR_ARG, "mhada_proj: inputs must be 16-byte aligned, outputs 32-byte aligned");
    if (int e = device_check()) return e;
    return launch_proj_bf16(parts, fc, fs, mean_c, rstd_c, mean_s, rstd_s, w, bias, B, Bs, Nc, Ns, H, d, q, k, v, mu_v, ws, s);
}

int mhada_attn(const mhada_attn_args* a, mhada_stream_t stream) {
    REQUIRE(a, MHADA_ERR_ARG, "mhada_attn: null args");
    REQUIRE(a->q && a->k && a->v && a->x && a->out && a->x_mean && a->x_rstd, MHADA_ERR_ARG, "mhada_attn: null pointer");
    REQUIRE(a->B > 0 && a->H > 0 && a->Nc > 0 && a->Ns > 0 && a->dqk > 0 && a->dv > 0, MHADA_ERR_ARG,
            "mhada_attn: bad sizes");
    REQUIRE(a->activation == MHADA_ACT_SOFTMAX || a->activation == MHADA_ACT_COSINE, MHADA_ERR_ARG,
            "mhada_attn: unknown activation %d", a->activation);
    REQUIRE(a->activation == MHADA_ACT_SOFTMAX || a->dtype == MHADA_F32, MHADA_ERR_UNSUPPORTED,
            "mhada_attn: the cosine activation runs on the MHADA_F32 path only");
    REQUIRE(a->kv_batch == 0 || a->kv_batch == 1 || a->kv_batch == a->B, MHADA_ERR_ARG,
            "mhada_attn: kv_batch must be 0, 1 or B");
    REQUIRE((a->q_mean == nullptr) == (a->q_rstd == nullptr) && (a->k_mean == nullptr) == (a->k_rstd == nullptr),
            MHADA_ERR_ARG, "mhada_attn: mean/rstd must be given together");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (a->dtype == MHADA_F32) {
        REQUIRE(a->ldq >= a->H * a->dqk && a->ldk >= a->H * a->dqk && a->ldv >= a->H * a->dv &&
                    a->ldx >= a->H * a->dv && a->ldo >= a->H * a->dv,
                MHADA_ERR_ARG, "mhada_attn: pitch smaller than H*d");
        if (int e = device_check()) return e;
        return attn_dispatch(*a, s);
    }
    REQUIRE(a->dtype == MHADA_BF16, MHADA_ERR_ARG, "mhada_attn: bad dtype %d", a->dtype);
    REQUIRE(a->dqk == a->dv && (a->dqk == 64 || a->dqk == 128), MHADA_ERR_UNSUPPORTED,
            "mhada_attn: the bf16 tensor-core path implements dqk = dv = 64 or 128, got %d / %d", a->dqk, a->dv);
    REQUIRE(!a->q_mean && !a->k_mean, MHADA_ERR_UNSUPPORTED, "mhada_attn: normalise-on-load is f32-path only");
    const int C = a->H * a->dqk;
    REQUIRE(a->ldq >= C && a->ldk >= C && a->ldv >= 2 * C && a->ldx >= C && a->ldo >= C, MHADA_ERR_ARG,
            "mhada_attn: pitch smaller than the row");
    REQUIRE(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->ldx % 8 == 0 && a->ldo % 16 == 0 &&
                aligned16(a->q) && aligned16(a->k) && aligned16(a->v) && aligned16(a->x) && aligned32(a->out),
            MHADA_ERR_ARG,
            "mhada_attn: bf16 inputs must be 16-byte aligned with pitches multiples of 8; out 32-byte aligned, ldo a multiple of 16");
    if (int e = device_check()) return e;
    return attn_dispatch(*a, s);
}

int mhada_debug_attn_trace(const mhada_attn_args* a, long long* trace, mhada_stream_t stream) {
    REQUIRE(a && trace, MHADA_ERR_ARG, "mhada_debug_attn_trace: null pointer");
    REQUIRE(a->dtype == MHADA_BF16 && a->dqk == 64 && a->dv == 64, MHADA_ERR_UNSUPPORTED,
            "mhada_debug_attn_trace: bf16 / head_dim 64 only");
    if (int e = device_check()) return e;
    return launch_attn_bf16_impl(*a, trace, static_cast<cudaStream_t>(stream));
}

size_t mhada_linear_workspace(int dtype, int Cout, int Cin) {
    if (dtype != MHADA_BF16 || Cout <= 0 || Cin <= 0) return 0;
    return linear_bf16_workspace(Cout, Cin);
}

int mhada_linear(int dtype, const void* x, int ldx, const float* w, const float* bias, int M, int Cin, int Cout,
                 void* y, int ldy, void* ws, size_t ws_bytes, mhada_stream_t stream) {
    REQUIRE(x && w && bias && y, MHADA_ERR_ARG, "mhada_linear: null pointer");
    REQUIRE(M > 0 && Cin > 0 && Cout > 0 && ldx >= Cin && ldy >= Cout, MHADA_ERR_ARG, "mhada_linear: bad sizes");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == MHADA_F32) {
        if (int e = device_check()) return e;
        return launch_linear_f32(static_cast<const float*>(x), ldx, w, bias, M, Cin, Cout, static_cast<float*>(y), ldy, s);
    }
    REQUIRE(dtype == MHADA_BF16, MHADA_ERR_ARG, "mhada_linear: bad dtype %d", dtype);
    REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, MHADA_ERR_UNSUPPORTED,
            "mhada_linear: bf16 path needs Cin and Cout multiples of 64, got %d / %d", Cin, Cout);
    REQUIRE(ldx % 8 == 0 && ldy % 16 == 0 && aligned16(x) && aligned32(y) && aligned16(ws), MHADA_ERR_ARG,
            "mhada_linear: x 16-byte aligned with ldx a multiple of 8; y 32-byte aligned with ldy a multiple of 16");
    REQUIRE(ws && ws_bytes >= linear_bf16_workspace(Cout, Cin), MHADA_ERR_WORKSPACE, "mhada_linear: workspace too small");
    if (int e = device_check()) return e;
    return launch_linear_bf16(x, ldx, w, bias, M, Cin, Cout, y, ldy, ws, s);
}

size_t mhada_vit_workspace(int B, int N, int D, int F, int K0) {
    if (B <= 0 || N <= 0 || D <= 0 || F <= 0 || K0 <= 0) return 0;
    return vit_workspace(B, N, D, F, K0);
}

int mhada_vit_forward(const mhada_vit_args* a, mhada_stream_t stream) {
    g_launches = 0;
    REQUIRE(a, MHADA_ERR_ARG, "mhada_vit_forward: null args");
    REQUIRE(a->img && a->w_patch && a->b_patch && a->ws, MHADA_ERR_ARG, "mhada_vit_forward: null pointer");
    REQUIRE(a->img_dtype == MHADA_F32 || a->img_dtype == MHADA_U8, MHADA_ERR_ARG,
            "mhada_vit_forward: images are MHADA_F32 or MHADA_U8 (0..255), got dtype %d", a->img_dtype);
    REQUIRE(a->B > 0 && a->Himg > 0 && a->Wimg > 0 && a->patch > 0 && a->D > 0 && a->F > 0 && a->heads > 0 && a->n_layers > 0,
            MHADA_ERR_ARG, "mhada_vit_forward: bad sizes");
    REQUIRE(a->n_layers <= MHADA_VIT_MAX_LAYERS, MHADA_ERR_UNSUPPORTED, "mhada_vit_forward: at most %d layers", MHADA_VIT_MAX_LAYERS);
    REQUIRE(a->Himg % a->patch == 0 && a->Wimg % a->patch == 0, MHADA_ERR_ARG,
            "mhada_vit_forward: image %dx%d is not a multiple of the patch size %d", a->Himg, a->Wimg, a->patch);
    REQUIRE(a->patch == 8 || a->patch == 16, MHADA_ERR_UNSUPPORTED, "mhada_vit_forward: patch size 8 or 16, got %d", a->patch);
    REQUIRE(a->D % 128 == 0 && a->D <= 1024 && a->F % 128 == 0 && a->D % a->heads == 0 && a->D / a->heads == 64,
            MHADA_ERR_UNSUPPORTED, "mhada_vit_forward: needs D %% 128 == 0, D <= 1024, F %% 128 == 0 and head_dim 64 (D=%d F=%d heads=%d)",
            a->D, a->F, a->heads);
    REQUIRE(a->B <= 32, MHADA_ERR_UNSUPPORTED,
            "mhada_vit_forward: the batch-axis attention (vit.py:48,59) is implemented for batch <= 32, got %d", a->B);
    const int N = (a->Himg / a->patch) * (a->Wimg / a->patch), K0 = 3 * a->patch * a->patch;
    const int elem = a->img_dtype == MHADA_F32 ? 4 : 1;
    REQUIRE((reinterpret_cast<uintptr_t>(a->img) % (8 * elem)) == 0 && (a->Wimg * elem) % (8 * elem) == 0, MHADA_ERR_ARG,
            "mhada_vit_forward: image rows must be aligned to 8 pixels");
    REQUIRE(aligned16(a->w_patch) && aligned16(a->ws) && (!a->pos || aligned16(a->pos)), MHADA_ERR_ARG,
            "mhada_vit_forward: misaligned pointer");
    for (int l = 0; l < a->n_layers; ++l) {
        const mhada_vit_layer& L = a->layers[l];
        REQUIRE(L.w_in && L.w_out && L.w_fc1 && L.w_fc2 && L.b_in && L.b_out && L.b_fc1 && L.b_fc2 && L.ln1_g && L.ln1_b &&
                    L.ln2_g && L.ln2_b && a->feat_f32[l],
                MHADA_ERR_ARG, "mhada_vit_forward: null pointer in layer %d", l);
        REQUIRE(aligned16(L.w_in) && aligned16(L.w_out) && aligned16(L.w_fc1) && aligned16(L.w_fc2) && aligned16(L.ln1_g) &&
                    aligned16(L.ln1_b) && aligned16(L.ln2_g) && aligned16(L.ln2_b) && aligned32(a->feat_f32[l]) &&
                    (!a->feat_bf16[l] || aligned16(a->feat_bf16[l])),
                MHADA_ERR_ARG, "mhada_vit_forward: misaligned pointer in layer %d", l);
    }
    REQUIRE(a->ws_bytes >= vit_workspace(a->B, N, a->D, a->F, K0), MHADA_ERR_WORKSPACE, "mhada_vit_forward: workspace %zu < %zu",
            a->ws_bytes, vit_workspace(a->B, N, a->D, a->F, K0));
    if (int e = device_check()) return e;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    StageTimer timer(MHADA_STAGE_VIT, s);
    return vit_forward(*a, s);
}

int mhada_patch_im2col(int img_dtype, const void* img, int B, int Himg, int Wimg, int patch, void* a0, mhada_stream_t stream) {
    REQUIRE(img && a0, MHADA_ERR_ARG, "mhada_patch_im2col: null pointer");
    REQUIRE(img_dtype == MHADA_F32 || img_dtype == MHADA_U8, MHADA_ERR_ARG, "mhada_patch_im2col: bad image dtype %d", img_dtype);
    REQUIRE(B > 0 && Himg > 0 && Wimg > 0 && patch > 0 && Himg % patch == 0 && Wimg % patch == 0, MHADA_ERR_ARG,
            "mhada_patch_im2col: bad sizes");
    const int elem = img_dtype == MHADA_F32 ? 4 : 1;
    REQUIRE((reinterpret_cast<uintptr_t>(img) % (8 * elem)) == 0 && aligned16(a0), MHADA_ERR_ARG, "mhada_patch_im2col: misaligned pointer");
    if (int e = device_check()) return e;
    return launch_patch_im2col(img_dtype, img, B, Himg, Wimg, patch, a0, static_cast<cudaStream_t>(stream));
}

int mhada_gemm_bf16(const void* x, int lda, const void* w, int ldw, const float* bias, int M, int N, int K, void* out_bf16,
                    int ldo, float* out_f32, int ldf, const float* resid, int ldr, int resid_mod, int relu,
                    mhada_stream_t stream) {
    REQUIRE(x && w && (out_bf16 || out_f32), MHADA_ERR_ARG, "mhada_gemm_bf16: null pointer");
    REQUIRE(M > 0 && N > 0 && K > 0 && lda >= K && ldw >= K, MHADA_ERR_ARG, "mhada_gemm_bf16: bad sizes");
    REQUIRE(K % 64 == 0 && N % 128 == 0, MHADA_ERR_UNSUPPORTED, "mhada_gemm_bf16: needs K %% 64 == 0 and N %% 128 == 0, got N=%d K=%d", N, K);
    REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && aligned16(x) && aligned16(w), MHADA_ERR_ARG,
            "mhada_gemm_bf16: operands must be 16-byte aligned with pitches multiples of 8");
    REQUIRE(!out_bf16 || (ldo >= N && ldo % 8 == 0 && aligned16(out_bf16)), MHADA_ERR_ARG, "mhada_gemm_bf16: bad bf16 output");
    REQUIRE(!out_f32 || (ldf >= N && ldf % 8 == 0 && aligned32(out_f32)), MHADA_ERR_ARG, "mhada_gemm_bf16: bad f32 output");
    REQUIRE(!resid || (ldr >= N && ldr % 4 == 0 && aligned16(resid) && resid_mod >= 0), MHADA_ERR_ARG, "mhada_gemm_bf16: bad residual");
    if (int e = device_check()) return e;
    GemmDesc g{};
    g.a = x; g.lda = lda; g.w = w; g.ldw = ldw; g.bias = bias; g.M = M; g.N = N; g.K = K;
    g.out_bf16 = out_bf16; g.ldo = ldo; g.out_f32 = out_f32; g.ldf = ldf; g.resid = resid; g.ldr = ldr; g.resid_mod = resid_mod;
    g.relu = relu;
    return launch_gemm_bf16(g, static_cast<cudaStream_t>(stream));
}

int mhada_layernorm(const float* x, int M, int C, const float* gamma, const float* beta, float eps, void* y_bf16,
                    mhada_stream_t stream) {
    REQUIRE(x && gamma && beta && y_bf16, MHADA_ERR_ARG, "mhada_layernorm: null pointer");
    REQUIRE(M > 0 && C > 0, MHADA_ERR_ARG, "mhada_layernorm: bad sizes");
    REQUIRE(aligned16(x) && aligned16(gamma) && aligned16(beta) && aligned16(y_bf16), MHADA_ERR_ARG, "mhada_layernorm: misaligned pointer");
    if (int e = device_check()) return e;
    return launch_layernorm(x, M, C, gamma, beta, eps, y_bf16, static_cast<cudaStream_t>(stream));
}

int mhada_batch_attn(const void* qkv, int B, int N, int heads, int hd, void* out, mhada_stream_t stream) {
    REQUIRE(qkv && out, MHADA_ERR_ARG, "mhada_batch_attn: null pointer");
    REQUIRE(B > 0 && N > 0 && heads > 0 && hd > 0, MHADA_ERR_ARG, "mhada_batch_attn: bad sizes");
    REQUIRE(aligned16(qkv) && aligned16(out), MHADA_ERR_ARG, "mhada_batch_attn: misaligned pointer");
    if (int e = device_check()) return e;
    return launch_batch_attn(qkv, B, N, heads, hd, out, static_cast<cudaStream_t>(stream));
}

int mhada_conv3x3_small(int dtype, const void* x, const float* w, const float* bias, int B, int H, int W, int Cin, int Cout,
                        int relu, void* y, mhada_stream_t stream) {
    g_launches = 0;
    REQUIRE(x && w && bias && y, MHADA_ERR_ARG, "mhada_conv3x3_small: null pointer");
    REQUIRE(dtype == MHADA_BF16, MHADA_ERR_UNSUPPORTED, "mhada_conv3x3_small: MHADA_BF16 only");
    REQUIRE(B > 0 && H >= 2 && W >= 2, MHADA_ERR_ARG, "mhada_conv3x3_small: bad sizes (ReflectionPad2d(1) needs H, W >= 2)");
    REQUIRE(aligned16(x), MHADA_ERR_ARG, "mhada_conv3x3_small: x must be 16-byte aligned");
    if (int e = device_check()) return e;
    return launch_conv3x3_small(x, w, bias, B, H, W, Cin, Cout, relu, y, static_cast<cudaStream_t>(stream));
}

size_t mhada_forloss_workspace(int B, int Nc, int Ns, int dqk, int dv) {
    if (B <= 0 || Nc <= 0 || Ns <= 0 || dqk <= 0 || dv <= 0) return 0;
    return forloss_workspace(B, Nc, Ns, dqk, dv);
}

int mhada_forloss_forward(const mhada_forloss_args* a, mhada_stream_t stream) {
    g_launches = 0;
    REQUIRE(a, MHADA_ERR_ARG, "mhada_forloss_forward: null args");
    REQUIRE(a->c_x && a->s_x && a->c_1x && a->s_1x && a->out && a->ws, MHADA_ERR_ARG, "mhada_forloss_forward: null pointer");
    REQUIRE(a->B > 0 && a->Nc > 0 && a->Ns > 0 && a->dqk > 0 && a->dv > 0, MHADA_ERR_ARG, "mhada_forloss_forward: bad sizes");
    REQUIRE(a->dtype == MHADA_F32 || a->dtype == MHADA_BF16, MHADA_ERR_ARG, "mhada_forloss_forward: bad dtype %d", a->dtype);
    REQUIRE(a->dqk % 64 == 0 && a->dv % 64 == 0, MHADA_ERR_UNSUPPORTED,
            "mhada_forloss_forward: the tensor-core path needs dqk %% 64 == 0 and dv %% 64 == 0, got %d / %d (use mhada_attn, MHADA_F32)",
            a->dqk, a->dv);
    REQUIRE(aligned16(a->c_x) && aligned16(a->s_x) && aligned16(a->c_1x) && aligned16(a->s_1x) && aligned16(a->out) && aligned32(a->ws),
            MHADA_ERR_ARG, "mhada_forloss_forward: misaligned pointer");
    REQUIRE(a->ws_bytes >= forloss_workspace(a->B, a->Nc, a->Ns, a->dqk, a->dv), MHADA_ERR_WORKSPACE,
            "mhada_forloss_forward: workspace %zu < %zu", a->ws_bytes, forloss_workspace(a->B, a->Nc, a->Ns, a->dqk, a->dv));
    if (int e = device_check()) return e;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    StageTimer timer(MHADA_STAGE_ATTN, s);
    return forloss_forward(*a, s);
}

int mhada_conv3x3(int dtype, const void* xp, const void* w, const float* bias, int B, int H, int W, int Cin, int Cout, int relu,
                  int out_padded, void* y, mhada_stream_t stream) {
    g_launches = 0;
    REQUIRE(xp && w && bias && y, MHADA_ERR_ARG, "mhada_conv3x3: null pointer");
    REQUIRE(dtype == MHADA_BF16, MHADA_ERR_UNSUPPORTED, "mhada_conv3x3: MHADA_BF16 only");
    REQUIRE(B > 0 && H >= 2 && W >= 2 && Cin > 0 && Cout > 0, MHADA_ERR_ARG, "mhada_conv3x3: bad sizes (ReflectionPad2d(1) needs H, W >= 2)");
    REQUIRE(Cin % 64 == 0 && (Cout == 64 || Cout == 128 || Cout == 256), MHADA_ERR_UNSUPPORTED,
            "mhada_conv3x3: implemented for Cin %% 64 == 0 and Cout in {64, 128, 256}, got %d -> %d", Cin, Cout);
    REQUIRE(aligned16(xp) && aligned16(w) && aligned32(y), MHADA_ERR_ARG, "mhada_conv3x3: xp, w 16-byte and y 32-byte aligned");
    if (int e = device_check()) return e;
    return launch_conv3x3_tc(xp, w, bias, B, H, W, Cin, Cout, relu, out_padded, y, static_cast<cudaStream_t>(stream));
}

int mhada_pad_reflect(int dtype, const void* x, int B, int H, int W, int C, int upsample, void* y,
                      mhada_stream_t stream) {
    REQUIRE(x && y, MHADA_ERR_ARG, "mhada_pad_reflect: null pointer");
    REQUIRE(dtype == MHADA_F32 || dtype == MHADA_BF16, MHADA_ERR_ARG, "mhada_pad_reflect: bad dtype %d", dtype);
    REQUIRE(B > 0 && H > 1 && W > 1 && C > 0, MHADA_ERR_ARG,
            "mhada_pad_reflect: bad sizes B=%d H=%d W=%d C=%d (reflection needs H, W >= 2)", B, H, W, C);
    REQUIRE(C % (dtype == MHADA_BF16 ? 8 : 4) == 0 && aligned16(x) && aligned16(y), MHADA_ERR_ARG,
            "mhada_pad_reflect: C must be a multiple of %d and pointers 16-byte aligned", dtype == MHADA_BF16 ? 8 : 4);
    if (int e = device_check()) return e;
    return launch_pad_reflect(dtype, x, B, H, W, C, upsample, y, static_cast<cudaStream_t>(stream));
}

int mhada_pad_reflect_bwd(int dtype, const void* dyp, int B, int H, int W, int C, int upsample, void* dx,
                          mhada_stream_t stream) {
    REQUIRE(dyp && dx, MHADA_ERR_ARG, "mhada_pad_reflect_bwd: null pointer");
    REQUIRE(dtype == MHADA_F32 || dtype == MHADA_BF16, MHADA_ERR_ARG, "mhada_pad_reflect_bwd: bad dtype %d", dtype);
    REQUIRE(B > 0 && H > 1 && W > 1 && C > 0, MHADA_ERR_ARG,
            "mhada_pad_reflect_bwd: bad sizes B=%d H=%d W=%d C=%d (reflection needs H, W >= 2)", B, H, W, C);
    REQUIRE(C % (dtype == MHADA_BF16 ? 8 : 4) == 0 && aligned16(dyp) && aligned16(dx), MHADA_ERR_ARG,
            "mhada_pad_reflect_bwd: C must be a multiple of %d and pointers 16-byte aligned", dtype == MHADA_BF16 ? 8 : 4);
    if (int e = device_check()) return e;
    return launch_pad_reflect_bwd(dtype, dyp, B, H, W, C, upsample, dx, static_cast<cudaStream_t>(stream));
}

size_t mhada_layer_workspace(int dtype, int B, int Nc, int Ns, int C, int H) {
    if (B <= 0 || Nc <= 0 || Ns <= 0 || C <= 0 || H <= 0 || C % H != 0) return 0;
    return carve(dtype, B, Nc, Ns, C, H, nullptr).total;
}

size_t mhada_style_cache_bytes(int dtype, int Bs, int Ns, int C, int H) {
    if (Bs <= 0 || Ns <= 0 || C <= 0 || H <= 0 || C % H != 0) return 0;
    return carve_cache(dtype, Bs, Ns, C, nullptr).total;
}

int mhada_layer_forward(int dtype, const void* fc, const void* fs, const void* fcs, const float* w_fgh,
                        const float* b_fgh, const float* w_out, const float* b_out, int B, int Nc, int Ns, int C,
                        int H, int flags, void* out, void* ws, size_t ws_bytes, mhada_stream_t stream) {
    REQUIRE(fs, MHADA_ERR_ARG, "mhada_layer_forward: null pointer");
    return layer_impl("mhada_layer_forward", dtype, fc, fs, fcs, nullptr, B, w_fgh, b_fgh, w_out, b_out, B, Nc, Ns, C, H,
                      flags, out, ws, ws_bytes, stream);
}

int mhada_layer_forward_cached(int dtype, const void* fc, const void* fcs, const void* cache, int Bs,
                               const float* w_fgh, const float* b_fgh, const float* w_out, const float* b_out, int B,
                               int Nc, int Ns, int C, int H, int flags, void* out, void* ws, size_t ws_bytes,
                               mhada_stream_t stream) {
    REQUIRE(cache && aligned32(cache), MHADA_ERR_ARG, "mhada_layer_forward_cached: cache must be a 32-byte aligned pointer");
    REQUIRE(Bs == 1 || Bs == B, MHADA_ERR_ARG, "mhada_layer_forward_cached: style batch %d must be 1 or the content batch %d", Bs, B);
    return layer_impl("mhada_layer_forward_cached", dtype, fc, nullptr, fcs, cache, Bs, w_fgh, b_fgh, w_out, b_out, B, Nc, Ns,
                      C, H, flags & ~MHADA_REUSE_FS_STATS, out, ws, ws_bytes, stream);   /* MHADA_WOUT_BF16 passes through */
}

int mhada_style_precompute(int dtype, const void* fs, const float* w_fgh, const float* b_fgh, int Bs, int Ns, int C,
                           int H, void* cache, size_t cache_bytes, void* ws, size_t ws_bytes, mhada_stream_t stream) {
    g_launches = 0;
    REQUIRE(fs && w_fgh && b_fgh && cache && ws, MHADA_ERR_ARG, "mhada_style_precompute: null pointer");
    REQUIRE(Bs > 0 && Ns > 0 && C > 0 && H > 0 && C % H == 0, MHADA_ERR_ARG, "mhada_style_precompute: bad sizes");
    REQUIRE(dtype == MHADA_F32 || dtype == MHADA_BF16, MHADA_ERR_ARG, "mhada_style_precompute: bad dtype %d", dtype);
    const int d = C / H;
    if (dtype == MHADA_BF16)
        REQUIRE((d == 64 || d == 128) && C % 16 == 0, MHADA_ERR_UNSUPPORTED,
                "mhada_style_precompute: the bf16 tensor-core path implements head_dim 64 and 128 (C/H = %d)", d);
    REQUIRE(aligned16(fs) && aligned32(cache) && aligned32(ws), MHADA_ERR_ARG, "mhada_style_precompute: misaligned pointer");
    StyleCacheView cv = carve_cache(dtype, Bs, Ns, C, static_cast<uint8_t*>(cache));
    REQUIRE(cache_bytes >= cv.total, MHADA_ERR_WORKSPACE, "mhada_style_precompute: cache %zu < %zu", cache_bytes, cv.total);
    LayerWs w = carve(dtype, Bs, Ns, Ns, C, H, static_cast<uint8_t*>(ws));
    REQUIRE(ws_bytes >= w.total, MHADA_ERR_WORKSPACE, "mhada_style_precompute: workspace %zu < %zu", ws_bytes, w.total);
    if (int e = device_check()) return e;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (int e = launch_stats(fs, dtype, Bs, Ns, C, C, w.mean_s, w.rstd_s, static_cast<float*>(w.stats_ws), s)) return e;
    if (dtype == MHADA_BF16)
        return launch_proj_bf16(MHADA_PROJ_KV, nullptr, fs, nullptr, nullptr, w.mean_s, w.rstd_s, w_fgh, b_fgh, 0, Bs, 0, Ns, H,
                                d, nullptr, cv.k, cv.v, cv.mu_v, w.proj_ws, s);
    return launch_proj_f32(MHADA_PROJ_KV, nullptr, static_cast<const float*>(fs), nullptr, nullptr, w.mean_s, w.rstd_s, w_fgh,
                           b_fgh, 0, Bs, 0, Ns, H, d, nullptr, static_cast<float*>(cv.k), static_cast<float*>(cv.v), cv.mu_v, s);
}

size_t mhada_layer_backward_workspace(int B, int Nc, int Ns, int C, int H) {
    if (B <= 0 || Nc <= 0 || Ns <= 0 || C <= 0 || H <= 0 || C % H != 0) return 0;
    return carve_bwd(B, Nc, Ns, C, H, nullptr).total;
}

int mhada_attn_bwd(int B, int H, int Nc, int Ns, const void* q, const void* k, const void* v, const void* x,
                   const float* x_mean, const float* x_rstd, const float* g, void* d_o, float* lse, float* delta,
                   float* d_xhat, void* d_q, void* d_k, void* d_v, mhada_stream_t stream) {
    g_launches = 0;
    REQUIRE(q && k && v && x && x_mean && x_rstd && g && d_o && lse && delta && d_xhat && d_q && d_k && d_v, MHADA_ERR_ARG,
            "mhada_attn_bwd: null pointer");
    REQUIRE(B > 0 && H > 0 && Nc > 0 && Ns > 0, MHADA_ERR_ARG, "mhada_attn_bwd: bad sizes");
    REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(x) && aligned16(g) && aligned16(d_o) && aligned16(d_xhat) &&
                aligned16(d_q) && aligned16(d_k) && aligned16(d_v),
            MHADA_ERR_ARG, "mhada_attn_bwd: pointers must be 16-byte aligned");
    if (int e = device_check()) return e;
    return launch_attn_bwd(B, H, Nc, Ns, H * 64, q, k, v, x, x_mean, x_rstd, g, d_o, lse, delta, d_xhat, d_q, d_k, d_v,
                           static_cast<cudaStream_t>(stream));
}

int mhada_layer_backward(const mhada_layer_bwd_args* a, mhada_stream_t stream) {
    REQUIRE(a, MHADA_ERR_ARG, "mhada_layer_backward: null args");
    REQUIRE(a->fc && a->fs && a->fcs && a->w_fgh && a->b_fgh && a->w_out && a->b_out && a->d_out && a->ws, MHADA_ERR_ARG,
            "mhada_layer_backward: null input pointer");
    REQUIRE(a->d_fc && a->d_fs && a->d_fcs && a->d_w_fgh && a->d_b_fgh && a->d_w_out && a->d_b_out, MHADA_ERR_ARG,
            "mhada_layer_backward: null output pointer");
    const int B = a->B, Nc = a->Nc, Ns = a->Ns, C = a->C, H = a->H;
    REQUIRE(B > 0 && Nc > 0 && Ns > 0 && C > 0 && H > 0 && C % H == 0, MHADA_ERR_ARG, "mhada_layer_backward: bad sizes");
    REQUIRE(C / H == 64 && C % 128 == 0, MHADA_ERR_UNSUPPORTED,
            "mhada_layer_backward: implemented for head_dim 64 and C %% 128 == 0 (C=%d, H=%d)", C, H);
    REQUIRE(aligned16(a->fc) && aligned16(a->fs) && aligned16(a->fcs) && aligned16(a->d_out) && aligned16(a->d_fc) &&
                aligned16(a->d_fs) && aligned16(a->d_fcs) && aligned32(a->ws),
            MHADA_ERR_ARG, "mhada_layer_backward: misaligned pointer");
    BwdWs w = carve_bwd(B, Nc, Ns, C, H, static_cast<uint8_t*>(a->ws));
    REQUIRE(a->ws_bytes >= w.total, MHADA_ERR_WORKSPACE, "mhada_layer_backward: workspace %zu < %zu", a->ws_bytes, w.total);
    if (int e = device_check()) return e;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int d = 64, Mc = B * Nc, Ms = B * Ns;
    // (0) recompute the forward up to the attention output (statistics, Q, K, V', heads)       adaDecoder.py:173-198
    if (int e = layer_impl("mhada_layer_backward", MHADA_BF16, a->fc, a->fs, a->fcs, nullptr, 0, a->w_fgh, a->b_fgh, nullptr,
                           nullptr, B, Nc, Ns, C, H, 0, w.heads, w.fwd, w.fwd_bytes, stream))
        return e;
    int launches = g_launches;
    LayerWs f = carve(MHADA_BF16, B, Nc, Ns, C, H, static_cast<uint8_t*>(w.fwd));
    const float *mean_x = a->fcs != a->fc ? f.mean_x : f.mean_c, *rstd_x = a->fcs != a->fc ? f.rstd_x : f.rstd_c;
    auto gemm = [&](const void* A, int lda, const void* W, int ldw, int M, int N, int K, float* out, int ldf) {
        GemmDesc g{};
        g.a = A; g.lda = lda; g.w = W; g.ldw = ldw; g.M = M; g.N = N; g.K = K; g.out_f32 = out; g.ldf = ldf;
        return launch_gemm_bf16(g, s);
    };
    auto gemm_wgrad = [&](const void* A, const void* W, int Mpad, float* out) {       // C x C result, K = tokens: split-K
        GemmDesc g{};
        g.a = A; g.lda = Mpad; g.w = W; g.ldw = Mpad; g.M = C; g.N = C; g.K = Mpad; g.out_f32 = out; g.ldf = C;
        return launch_gemm_bf16_splitk(g, w.splitk, w.splitk_bytes, s);
    };
    // (1) out_conv backward                                                                    adaDecoder.py:202-205
    if (int e = launch_transpose_norm(a->w_out, MHADA_F32, C, C, C, C, C, nullptr, nullptr, w.woT, s)) return e;
    if (int e = gemm(a->d_out, C, w.woT, C, Mc, C, C, w.dcat, C)) return e;                         // d(cat) = d(out) Wo
    if (int e = launch_transpose_norm(a->d_out, MHADA_BF16, C, Mc, w.Mpad_c, C, Nc, nullptr, nullptr, w.tA, s)) return e;
    if (int e = launch_transpose_norm(w.heads, MHADA_BF16, C, Mc, w.Mpad_c, C, Nc, nullptr, nullptr, w.tB, s)) return e;
    if (int e = gemm_wgrad(w.tA, w.tB, w.Mpad_c, a->d_w_out)) return e;                             // dWo = d(out)^T cat
    if (int e = launch_token_sums(a->d_out, MHADA_BF16, nullptr, nullptr, nullptr, B, Nc, C, w.partial, nullptr, a->d_b_out, s))
        return e;
    // (2) attention backward                                                                   adaDecoder.py:186-198
    if (int e = launch_attn_bwd(B, H, Nc, Ns, C, f.q, f.k, f.v, a->fcs, mean_x, rstd_x, w.dcat, w.d_o, w.lse, w.delta, w.dxhat,
                                w.dq, w.dk, w.dv, s))
        return e;
    // (3) projections backward: inputs                                                         adaDecoder.py:173-183
    if (int e = launch_blockdiag_t(a->w_fgh, H, d, w.wbd, s)) return e;
    const uint16_t* wbd = static_cast<const uint16_t*>(w.wbd);      // bf16 [3][C][C]
    const size_t CC = static_cast<size_t>(C) * C;
    if (int e = gemm(w.dq, C, wbd, C, Mc, C, C, w.gq, C)) return e;                                 // d(IN(fc)) = dQ Wf
    if (int e = gemm(w.dk, C, wbd + CC, C, Ms, C, C, w.gk, C)) return e;                            // d(IN(fs)) = dK Wg
    if (int e = gemm(w.dv, C, wbd + 2 * CC, C, Ms, C, C, w.gv, C)) return e;                        // d(fs) via V = dV Wh
    // (4) projections backward: weights and biases
    struct Role { const void* dy; const void* x; const float *mean, *rstd; int N, M, Mpad; };
    const Role roles[3] = {{w.dq, a->fc, f.mean_c, f.rstd_c, Nc, Mc, w.Mpad_c},
                           {w.dk, a->fs, f.mean_s, f.rstd_s, Ns, Ms, w.Mpad_s},
                           {w.dv, a->fs, nullptr, nullptr, Ns, Ms, w.Mpad_s}};
    for (int r = 0; r < 3; ++r) {
        const Role& R = roles[r];
        if (int e = launch_transpose_norm(R.dy, MHADA_BF16, C, R.M, R.Mpad, C, R.N, nullptr, nullptr, w.tA, s)) return e;
        if (int e = launch_transpose_norm(R.x, MHADA_BF16, C, R.M, R.Mpad, C, R.N, R.mean, R.rstd, w.tB, s)) return e;
        if (int e = gemm_wgrad(w.tA, w.tB, R.Mpad, w.dwfull)) return e;
        if (int e = launch_extract_blockdiag(w.dwfull, H, d, a->d_w_fgh + static_cast<size_t>(r) * H * d * d, s)) return e;
        if (int e = launch_token_sums(R.dy, MHADA_BF16, nullptr, nullptr, nullptr, B, R.N, C, w.partial, nullptr,
                                      a->d_b_fgh + static_cast<size_t>(r) * C, s))
            return e;
    }
    // (5) instance-norm backward                                                               adaDecoder.py:147-149
    if (int e = launch_token_sums(w.gq, MHADA_F32, a->fc, f.mean_c, f.rstd_c, B, Nc, C, w.partial, w.sums, nullptr, s)) return e;
    if (int e = launch_in_bwd_apply(w.gq, a->fc, f.mean_c, f.rstd_c, w.sums, nullptr, B, Nc, C, a->d_fc, s)) return e;
    if (int e = launch_token_sums(w.gk, MHADA_F32, a->fs, f.mean_s, f.rstd_s, B, Ns, C, w.partial, w.sums, nullptr, s)) return e;
    if (int e = launch_in_bwd_apply(w.gk, a->fs, f.mean_s, f.rstd_s, w.sums, w.gv, B, Ns, C, a->d_fs, s)) return e;
    if (int e = launch_token_sums(w.dxhat, MHADA_F32, a->fcs, mean_x, rstd_x, B, Nc, C, w.partial, w.sums, nullptr, s)) return e;
    if (int e = launch_in_bwd_apply(w.dxhat, a->fcs, mean_x, rstd_x, w.sums, nullptr, B, Nc, C, a->d_fcs, s)) return e;
    g_launches += launches;
    return 0;
}

int mhada_transpose_bf16(const void* x, int dtype, int ld, int M, int C, int Mpad, void* out, mhada_stream_t stream) {
    REQUIRE(x && out, MHADA_ERR_ARG, "mhada_transpose_bf16: null pointer");
    REQUIRE(dtype == MHADA_F32 || dtype == MHADA_BF16, MHADA_ERR_ARG, "mhada_transpose_bf16: bad dtype %d", dtype);
    REQUIRE(M > 0 && C > 0 && ld >= C && Mpad >= M, MHADA_ERR_ARG, "mhada_transpose_bf16: bad sizes");
    if (int e = device_check()) return e;
    return launch_transpose_norm(x, dtype, ld, M, Mpad, C, M, nullptr, nullptr, out, static_cast<cudaStream_t>(stream));
}

size_t mhada_colsum_workspace(int M, int C) {
    if (M <= 0 || C <= 0) return 0;
    return token_sums_workspace(1, M, C);
}

int mhada_colsum(const void* x, int dtype, int M, int C, void* ws, size_t ws_bytes, float* out, mhada_stream_t stream) {
    REQUIRE(x && ws && out, MHADA_ERR_ARG, "mhada_colsum: null pointer");
    REQUIRE(dtype == MHADA_F32 || dtype == MHADA_BF16, MHADA_ERR_ARG, "mhada_colsum: bad dtype %d", dtype);
    REQUIRE(M > 0 && C > 0, MHADA_ERR_ARG, "mhada_colsum: bad sizes");
    REQUIRE(ws_bytes >= token_sums_workspace(1, M, C), MHADA_ERR_WORKSPACE, "mhada_colsum: workspace too small");
    if (int e = device_check()) return e;
    return launch_token_sums(x, dtype, nullptr, nullptr, nullptr, 1, M, C, ws, nullptr, out, static_cast<cudaStream_t>(stream));
}

size_t mhada_gemm_splitk_workspace(int M, int N, int K) { return gemm_splitk_workspace(M, N, K); }

int mhada_gemm_bf16_splitk(const void* x, int lda, const void* w, int ldw, int M, int N, int K, float* out_f32, int ldf,
                           void* ws, size_t ws_bytes, mhada_stream_t stream) {
    REQUIRE(x && w && out_f32, MHADA_ERR_ARG, "mhada_gemm_bf16_splitk: null pointer");
    REQUIRE(M > 0 && N > 0 && K > 0 && lda >= K && ldw >= K, MHADA_ERR_ARG, "mhada_gemm_bf16_splitk: bad sizes");
    REQUIRE(K % 64 == 0 && N % 128 == 0, MHADA_ERR_UNSUPPORTED, "mhada_gemm_bf16_splitk: needs K %% 64 == 0 and N %% 128 == 0, got N=%d K=%d", N, K);
    REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && aligned16(x) && aligned16(w), MHADA_ERR_ARG,
            "mhada_gemm_bf16_splitk: operands must be 16-byte aligned with pitches multiples of 8");
    REQUIRE(ldf >= N && ldf % 8 == 0 && aligned32(out_f32) && (!ws || aligned32(ws)), MHADA_ERR_ARG, "mhada_gemm_bf16_splitk: bad output / workspace");
    if (int e = device_check()) return e;
    GemmDesc g{};
    g.a = x; g.lda = lda; g.w = w; g.ldw = ldw; g.M = M; g.N = N; g.K = K; g.out_f32 = out_f32; g.ldf = ldf;
    return launch_gemm_bf16_splitk(g, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}

int mhada_batch_attn_bwd(const void* qkv, const void* d_out, int B, int N, int heads, int hd, void* d_qkv, mhada_stream_t stream) {
    REQUIRE(qkv && d_out && d_qkv, MHADA_ERR_ARG, "mhada_batch_attn_bwd: null pointer");
    REQUIRE(B > 0 && N > 0 && heads > 0 && hd > 0, MHADA_ERR_ARG, "mhada_batch_attn_bwd: bad sizes");
    REQUIRE(aligned16(qkv) && aligned16(d_out) && aligned16(d_qkv), MHADA_ERR_ARG, "mhada_batch_attn_bwd: misaligned pointer");
    if (int e = device_check()) return e;
    return launch_batch_attn_bwd(qkv, d_out, B, N, heads, hd, d_qkv, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

/*
 * mhada_b200.h -- C ABI of libmhada_b200.so: the B200 (sm_100a) MHAda forward hot path.
 *
 * The reference (Maboroshi0327/MHAda-Style-Transfer) is pure Python/PyTorch and has no FFI; the
 * interface each entry point replaces is therefore a span of the reference's nn.Module code,
 * cited per function as MHAdaSTr/network/<file>:<lines>.  INTEGRATION.md shows the ctypes binding
 * a maintainer of the reference would add.
 *
 * Conventions
 *   - Every function returns 0 on success, a negative MHADA_ERR_* for a rejected argument and a
 *     positive cudaError_t for a CUDA failure; mhada_last_error() gives the message (thread local).
 *     Nothing throws across the boundary.
 *   - All buffers are DEVICE pointers owned by the caller (inputs const, outputs and workspace
 *     pre-allocated).  No allocation, no synchronisation, no default-stream use inside: every
 *     launch goes on the caller's `stream` (a cudaStream_t passed as void*), so a sequence of calls
 *     is CUDA-graph capturable.
 *   - Feature maps are TOKEN-MAJOR: element (b, n, c) of a (B, C, h, w) map lives at
 *     ((b * N) + n) * ld + c, n = y * w + x  (what torch calls channels_last; the reference ViT
 *     already emits this memory order, MHAdaSTr/network/vit.py:163-166).  `ld` >= C is the row
 *     pitch in elements.
 *   - dtype: MHADA_F32 = the reference's arithmetic (true fp32 FFMA, never TF32);
 *            MHADA_BF16 = bf16 storage, tcgen05 tensor-core math with fp32 accumulation.
 *   - There is no CPU fallback: on a device that is not compute capability 10.x every compute
 *     entry point returns MHADA_ERR_DEVICE.
 */
#ifndef MHADA_B200_H_
#define MHADA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MHADA_ABI_VERSION 13

#if defined(__GNUC__)
#define MHADA_API __attribute__((visibility("default")))
#else
#define MHADA_API
#endif

typedef void* mhada_stream_t; /* cudaStream_t */

enum mhada_dtype { MHADA_F32 = 0, MHADA_BF16 = 1, MHADA_U8 = 2 /* images only */ };
enum mhada_activation { MHADA_ACT_SOFTMAX = 0, MHADA_ACT_COSINE = 1 };

enum mhada_status {
    MHADA_OK = 0,
    MHADA_ERR_ARG = -1,         /* null pointer, non-positive size, misaligned pointer / pitch   */
    MHADA_ERR_UNSUPPORTED = -2, /* shape outside what the kernels implement (message says which) */
    MHADA_ERR_DEVICE = -3,      /* current device is not sm_100                                  */
    MHADA_ERR_WORKSPACE = -4,   /* workspace too small                                           */
    MHADA_ERR_DRIVER = -5       /* cuTensorMapEncodeTiled unavailable / failed                   */
};

MHADA_API int mhada_abi_version(void);
MHADA_API const char* mhada_last_error(void);
/* 0 when the current CUDA device can run the kernels (compute capability 10.x). */
MHADA_API int mhada_device_check(void);

/* ------------------------------------------------------------------------------------------------
 * (1) Instance-norm statistics -- replaces the statistics half of nn.InstanceNorm2d(affine=False)
 *     at MHAdaSTr/network/adaDecoder.py:147-149 (used :173, :178, :198): per (b, channel)
 *     mean over the N tokens and rstd = 1/sqrt(biased_var + 1e-5).
 *     x: [B, N, ld] (dtype), mean / rstd: float [B, C].
 *     ws: float scratch of mhada_in_stats_workspace(B, N, C) bytes.
 * ---------------------------------------------------------------------------------------------- */
MHADA_API size_t mhada_in_stats_workspace(int B, int N, int C);
MHADA_API int mhada_in_stats(const void* x, int dtype, int B, int N, int C, int ld, float* mean, float* rstd, void* ws,
                   size_t ws_bytes, mhada_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (2) Per-head 1x1 projections -- replaces f_list / g_list / h_list applied to IN(fc), IN(fs), fs
 *     at adaDecoder.py:173, :178, :182 (a grouped 1x1 convolution, groups = H, d = C / H):
 *         Q  = Wf . IN(fc) + bf        K = Wg . IN(fs) + bg        V = Wh . fs + bh
 *     Weights are passed packed: w float [3][H][d][d] (f, g, h; [out][in]), bias float [3][H][d].
 *     Outputs (token-major, pitch C):
 *         q  [B, Nc, C]   (bf16 path: pre-multiplied by log2(e) so the kernel can use exp2)
 *         k  [B, Ns, C]
 *         v  f32 path : [B, Ns, C]     centred values  V - mu_v
 *            bf16 path: [B, Ns, 2C]    per head h the 2d columns [V - mu_v | (V - mu_v)^2]
 *         mu_v float [B, C] = Wh . mean(fs) + bh   (added back by the attention epilogue; centring
 *            is exact for M and for A.V^2 - M^2 and removes the bf16 cancellation)
 *     ws (bf16 path): mhada_proj_workspace(B, H, d) bytes for the folded per-image weights.
 * ---------------------------------------------------------------------------------------------- */
#define MHADA_PROJ_Q 1  /* Q from fc (content batch B)         */
#define MHADA_PROJ_KV 2 /* K, V, mu_v from fs (style batch Bs) */
MHADA_API size_t mhada_proj_workspace(int B, int H, int d);   /* B = max(B, Bs) */
MHADA_API int mhada_proj(int dtype, int parts, const void* fc, const void* fs, const float* mean_c, const float* rstd_c,
               const float* mean_s, const float* rstd_s, const float* w, const float* bias, int B, int Bs, int Nc,
               int Ns, int H, int d, void* q, void* k, void* v, float* mu_v, void* ws, size_t ws_bytes,
               mhada_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (3) Streaming attention + fused AdaIN-style epilogue -- replaces adaDecoder.py:186-198
 *     (and :70-81 of AdaAttnForLoss, :120-131 of AdaAttN):
 *         A = softmax_rows(Q K^T)  (no 1/sqrt(d));  M = A V;  S = sqrt(max(A V^2 - M^2, 1e-6));
 *         out[:, head] = S * IN(x)[:, head] + M (+ mu_v)
 *     The Nc x Ns map A is never written anywhere.
 *     f32 path : any dqk, dv; q/k may be given un-normalised with (q_mean, q_rstd) / (k_mean, k_rstd)
 *                applied on load (AdaAttnForLoss); logits in natural units.
 *     bf16 path: dqk = dv = 64 (the 8-head configuration every reference script uses), q in
 *                log2 units, v = [V~ | V~^2] as written by mhada_proj; pitches fixed by C = H*64.
 * ---------------------------------------------------------------------------------------------- */
typedef struct mhada_attn_args {
    int dtype;
    int B, H, Nc, Ns, dqk, dv;
    const void* q; /* [B, Nc, ldq], head h at column h*dqk */
    const void* k; /* [B, Ns, ldk], head h at column h*dqk */
    const void* v; /* f32: [B, Ns, ldv] head h at column h*dv;  bf16: [B, Ns, ldv] head h at column h*2*dv */
    const void* x; /* [B, Nc, ldx] tensor whose instance norm is modulated (fcs), head h at column h*dv */
    void* out;     /* [B, Nc, ldo] head h at column h*dv */
    int ldq, ldk, ldv, ldx, ldo;
    const float* x_mean; /* [B, H*dv] */
    const float* x_rstd;
    const float* mu_v;   /* [B, H*dv] or NULL */
    const float* q_mean; /* f32 path only, [B, H*dqk] or NULL */
    const float* q_rstd;
    const float* k_mean;
    const float* k_rstd;
    int kv_batch;        /* batch of k / v / mu_v (/ k_mean, k_rstd): 0 or B = one per image; 1 = one style shared by
                            all B images (the reference cannot do this, adaDecoder.py:177-183; infer_video.py uses one
                            style for every frame) */
    int activation;      /* MHADA_ACT_SOFTMAX (Softmax, adaDecoder.py:11-17) or MHADA_ACT_COSINE (CosineSimilarity,
                            adaDecoder.py:20-34: a = (cos(q, k) + 1) / sum_k (cos(q, k) + 1)); cosine is f32-path only */
    void* scratch;       /* optional device scratch, 16-byte aligned: with mhada_attn_cosine_scratch(B or 1, H) bytes the
                            cosine activation at dqk = dv = 64 and Ns >= 128 runs in CLOSED FORM, O(N d^2): A V' = (q^ . T +
                            sum v') / (q^ . sum k^ + Ns) with T = sum_j k^_j (x) v'_j; NULL = the O(N^2 d) kernel */
    size_t scratch_bytes;
} mhada_attn_args;
MHADA_API size_t mhada_attn_cosine_scratch(int B, int H);
MHADA_API int mhada_attn(const mhada_attn_args* args, mhada_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (4) out_conv -- replaces torch.cat + nn.Conv2d(C, C, 1) at adaDecoder.py:202-205:
 *         y[m, o] = sum_i x[m, i] * w[o, i] + bias[o],   x [M, ldx], y [M, ldy], w float [Cout][Cin].
 *     bf16 path needs Cin % 64 == 0 and Cout % 64 == 0; ws = mhada_linear_workspace(Cout, Cin) bytes
 *     (bf16 copy of w).
 * ---------------------------------------------------------------------------------------------- */
MHADA_API size_t mhada_linear_workspace(int dtype, int Cout, int Cin);
MHADA_API int mhada_linear(int dtype, const void* x, int ldx, const float* w, const float* bias, int M, int Cin, int Cout,
                 void* y, int ldy, void* ws, size_t ws_bytes, mhada_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (5) One whole layer -- replaces AdaAttnMultiHead.forward(fc, fs, fcs), adaDecoder.py:162-206.
 *     fc, fcs [B, Nc, C]; fs [B, Ns, C]; out [B, Nc, C]; all token-major with pitch C, in `dtype`.
 *     fcs may alias fc (layer 0 of the transformer, adaDecoder.py:262).  out must not alias inputs.
 *     w_fgh / b_fgh as in (2); w_out float [C][C], b_out float [C]; both may be NULL together to
 *     skip out_conv (the single-head AdaAttN, adaDecoder.py:102-131).
 *     MHADA_BF16 head widths: 64 and 128 run the streaming tcgen05 kernel; multiples of 128 above that (1- and 2-head
 *     layers, AdaAttN) run per-head projections on the token GEMM and the materialised attention of (8).
 *     ws: mhada_layer_workspace(dtype, B, Nc, Ns, C, H) bytes.
 *     flags: MHADA_REUSE_FS_STATS when `fs` and `ws` are the ones passed to the previous call on this stream
 *     (the two layers of a level share fs, adaDecoder.py:264-265): its statistics are not recomputed.
 * ---------------------------------------------------------------------------------------------- */
#define MHADA_REUSE_FS_STATS 1
#define MHADA_LAYER_COSINE 2   /* activation = "cosine" (adaDecoder.py:155-160); MHADA_F32 only */
#define MHADA_WOUT_BF16 4      /* MHADA_BF16 path: w_out points to a bf16 [C][C] copy of out_conv.weight kept by the
                                  caller (saves the per-call f32 -> bf16 conversion launch); C % 128 == 0 */
MHADA_API size_t mhada_layer_workspace(int dtype, int B, int Nc, int Ns, int C, int H);
MHADA_API int mhada_layer_forward(int dtype, const void* fc, const void* fs, const void* fcs, const float* w_fgh,
                        const float* b_fgh, const float* w_out, const float* b_out, int B, int Nc, int Ns, int C,
                        int H, int flags, void* out, void* ws, size_t ws_bytes, mhada_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (5a) Style-side cache (no counterpart in the reference, which recomputes the style side of every
 *     layer for every frame, infer_video.py:91-92).  K, V and mu_v of a layer depend on fs and on the layer's g / h
 *     weights only: mhada_style_precompute() runs the fs statistics and the K / V projections once and stores
 *     [K | V (bf16: V') | mu_v] in `cache` (mhada_style_cache_bytes); mhada_layer_forward_cached() then runs the
 *     content side only (stats of fc / fcs, Q projection, attention, out_conv).  Bs = 1 serves any content batch B.
 *     Results are bit-identical to mhada_layer_forward on the same fs.
 *     ws for precompute: mhada_layer_workspace(dtype, Bs, Ns, Ns, C, H) bytes.
 * ---------------------------------------------------------------------------------------------- */
MHADA_API size_t mhada_style_cache_bytes(int dtype, int Bs, int Ns, int C, int H);
MHADA_API int mhada_style_precompute(int dtype, const void* fs, const float* w_fgh, const float* b_fgh, int Bs, int Ns,
                                     int C, int H, void* cache, size_t cache_bytes, void* ws, size_t ws_bytes,
                                     mhada_stream_t stream);
MHADA_API int mhada_layer_forward_cached(int dtype, const void* fc, const void* fcs, const void* cache, int Bs,
                                         const float* w_fgh, const float* b_fgh, const float* w_out,
                                         const float* b_out, int B, int Nc, int Ns, int C, int H, int flags, void* out,
                                         void* ws, size_t ws_bytes, mhada_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (5b) Decoder glue -- replaces nn.ReflectionPad2d(1) (MHAdaSTr/network/conv.py:26-27) and, when
 *     `upsample` != 0, the F.interpolate(scale_factor=2, mode="bilinear", align_corners=False) that
 *     precedes it (conv.py:71), in one pass over channels_last activations:
 *         x [B, H, W, C]  ->  y [B, Ho + 2, Wo + 2, C],   Ho, Wo = H, W or 2H, 2W.
 *     C must be a multiple of 8 (bf16) / 4 (f32); pointers 16-byte aligned.
 * ---------------------------------------------------------------------------------------------- */
/*     Last decoder block in one kernel -- replaces ReflectionPad2d(1) + Conv2d(64, Cout <= 8, 3) (+ ReLU) of
 *     conv3[1] = ConvReLU(64, 3, 3, 1), conv.py:90-93: x [B, H, W, 64] bf16 channels_last, NOT padded (the
 *     reflection is resolved in the kernel's load indices); w float [Cout][64][3][3]; y [B, Cout, H, W] bf16.
 *     MHADA_BF16 only (the fp32 path keeps the library convolution). */
MHADA_API int mhada_conv3x3_small(int dtype, const void* x, const float* w, const float* bias, int B, int H, int W, int Cin,
                                  int Cout, int relu, void* y, mhada_stream_t stream);
/*     Decoder blocks 0..7 -- replaces ReflectionPad2d(1) + Conv2d(Cin, Cout, 3) + ReLU, conv.py:23-45, as an implicit
 *     GEMM on tcgen05 (r1 called cuDNN here).  xp [B, H + 2, W + 2, Cin] bf16: the reflect-PADDED channels_last input
 *     (from mhada_pad_reflect, or from a previous mhada_conv3x3 with out_padded = 1);
 *     w bf16 [Cout][3][3][Cin] (conv weight permuted to (o, ky, kx, i)); bias float [Cout];
 *     y bf16: out_padded = 0 -> [B, H, W, Cout]; out_padded = 1 -> [B, H + 2, W + 2, Cout] with the reflection ring
 *     written by the epilogue, i.e. directly the next block's xp.  Cin % 64 == 0, Cout in {64, 128, 256}, H, W >= 2.
 *     MHADA_BF16 only. */
MHADA_API int mhada_conv3x3(int dtype, const void* xp, const void* w, const float* bias, int B, int H, int W, int Cin,
                            int Cout, int relu, int out_padded, void* y, mhada_stream_t stream);
MHADA_API int mhada_pad_reflect(int dtype, const void* x, int B, int H, int W, int C, int upsample, void* y,
                                mhada_stream_t stream);
/*     backward of mhada_pad_reflect for the training step (train_image.py:139 through conv.py:26-27, :71):
 *     dyp [B, Ho + 2, Wo + 2, C] (gradient of the padded, optionally up-sampled map) -> dx [B, H, W, C]; gather form,
 *     deterministic. */
MHADA_API int mhada_pad_reflect_bwd(int dtype, const void* dyp, int B, int H, int W, int C, int upsample, void* dx,
                                    mhada_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (7) ViT encoder -- replaces VisionTransformer.forward, MHAdaSTr/network/vit.py:148-169
 *     (PatchEmbedding :105-117, PosEmbedding :67-102, EncoderBlock :45-64), so that the host boundary of the
 *     pipeline is the IMAGE (infer_image.py:83-85: fc = vit_c(c); fs = vit_s(s); adaFormer(fc, fs)) and the
 *     feature maps never leave HBM.  bf16 tensor-core path (tcgen05 GEMMs, f32 residual stream).
 *
 *     img: [B, 3, Himg, Wimg] planar (NCHW), MHADA_F32 or MHADA_U8 values 0..255 (utilities.py:11-16);
 *          Himg, Wimg multiples of `patch`;  N = (Himg/patch) * (Wimg/patch) tokens, K0 = 3*patch*patch.
 *     w_patch bf16 [D][K0] (conv_proj.weight flattened), b_patch f32 [D];
 *     pos f32 [N][D] token-major positional table already resized to this N (vit.py:96-102), or NULL (vit_s).
 *     Per layer l: LayerNorm(eps 1e-6) -> in_proj -> attention -> out_proj (+x) -> LayerNorm -> fc1 -> ReLU ->
 *          fc2 (+x).  nn.MultiheadAttention is built WITHOUT batch_first (vit.py:48) and fed (B, N, D): it attends
 *          ACROSS THE BATCH for every token position (sequence length = B) -- reproduced exactly:
 *          for B = 1 the softmax is over one logit and the block reduces to out_proj(v_proj(LN(x))).
 *     feat_f32[l] f32 [B, N, D]: the residual stream after layer l (required, they double as working buffers);
 *     feat_bf16[l] bf16 [B, N, D] or NULL: the same values rounded, token-major = exactly the layout
 *          mhada_layer_forward takes (the reference permutes to (B, D, h, w) views, vit.py:163-166).
 *     ws: mhada_vit_workspace(B, N, D, F, K0) bytes.   D % 128 == 0, F % 128 == 0, K0 % 64 == 0, D / heads == 64,
 *          B <= 32.
 * ---------------------------------------------------------------------------------------------- */
#define MHADA_VIT_MAX_LAYERS 8
typedef struct mhada_vit_layer {
    const void *w_in, *w_out, *w_fc1, *w_fc2;        /* bf16 [3D][D], [D][D], [F][D], [D][F] */
    const float *b_in, *b_out, *b_fc1, *b_fc2;       /* f32  [3D],    [D],    [F],    [D]    */
    const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;      /* f32  [D] */
} mhada_vit_layer;
typedef struct mhada_vit_args {
    int img_dtype;
    const void* img;
    int B, Himg, Wimg, patch;
    int D, F, heads, n_layers;
    const void* w_patch;
    const float* b_patch;
    const float* pos;
    mhada_vit_layer layers[MHADA_VIT_MAX_LAYERS];
    float* feat_f32[MHADA_VIT_MAX_LAYERS];
    void* feat_bf16[MHADA_VIT_MAX_LAYERS];
    void* ws;
    size_t ws_bytes;
} mhada_vit_args;
MHADA_API size_t mhada_vit_workspace(int B, int N, int D, int F, int K0);
MHADA_API int mhada_vit_forward(const mhada_vit_args* args, mhada_stream_t stream);
/*     Stages of (7), exposed for the parity tests:
 *     mhada_patch_im2col: img -> a0 bf16 [B*N][K0], column = c*patch*patch + dy*patch + dx (Conv2d weight order);
 *     mhada_gemm_bf16:    y = act(x . W^T + bias (+ resid[m % resid_mod])), x bf16 [M, lda], W bf16 [N, ldw],
 *                         results bf16 [M, ldo] and / or f32 [M, ldf]; K % 64 == 0, N % 128 == 0;
 *     mhada_layernorm:    nn.LayerNorm(C, eps) over the last axis of f32 [M, C] -> bf16 [M, C]  (vit.py:54-55);
 *     mhada_batch_attn:   the batch_first=False attention: qkv bf16 [B, N, 3*heads*hd] -> bf16 [B, N, heads*hd],
 *                         softmax over the B images of a token position, scale 1/sqrt(hd). */
MHADA_API int mhada_patch_im2col(int img_dtype, const void* img, int B, int Himg, int Wimg, int patch, void* a0,
                                 mhada_stream_t stream);
MHADA_API int mhada_gemm_bf16(const void* x, int lda, const void* w, int ldw, const float* bias, int M, int N, int K,
                              void* out_bf16, int ldo, float* out_f32, int ldf, const float* resid, int ldr,
                              int resid_mod, int relu, mhada_stream_t stream);
MHADA_API int mhada_layernorm(const float* x, int M, int C, const float* gamma, const float* beta, float eps,
                              void* y_bf16, mhada_stream_t stream);
MHADA_API int mhada_batch_attn(const void* qkv, int B, int N, int heads, int hd, void* out, mhada_stream_t stream);
/*     backward of mhada_batch_attn for the training step (B <= 8): d_out bf16 [B, N, heads*hd] -> d_qkv bf16 [B, N, 3*heads*hd] */
MHADA_API int mhada_batch_attn_bwd(const void* qkv, const void* d_out, int B, int N, int heads, int hd, void* d_qkv,
                                   mhada_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (8) AdaAttnForLoss on the tensor cores -- replaces AdaAttnForLoss.forward, MHAdaSTr/network/adaDecoder.py:53-81
 *     (lossfn.py:26-34 calls it on concatenated VGG features: d_qk = 448 / 960 / 1472, d_v = 256 / 512):
 *         Q = IN(c_1x), K = IN(s_1x), V = s_x;  A = softmax(Q K^T);  out = sqrt(max(A V^2 - (A V)^2, 1e-6)) * IN(c_x) + A V
 *     All tensors in `dtype` (f32 or bf16), token-major with pitch = their channel count: c_x, out [B, Nc, dv];
 *     s_x [B, Ns, dv]; c_1x [B, Nc, dqk]; s_1x [B, Ns, dqk].  dqk % 64 == 0, dv % 64 == 0.
 *     ws: mhada_forloss_workspace(...) bytes.  The contractions run on bf16 tcgen05 MMAs with f32 accumulation; the
 *     normalised Q / K are split in two bf16 terms each (three products, one GEMM) and V^2 in an exact hi + lo pair,
 *     because these logits are 448..1472 terms wide and the attention is sharp (csrc/forloss_tc.cu, "Numerics").
 *     (mhada_attn with MHADA_F32 remains the reference-arithmetic path for any dqk / dv.)
 * ---------------------------------------------------------------------------------------------- */
typedef struct mhada_forloss_args {
    int dtype;                 /* MHADA_F32 or MHADA_BF16: storage type of the four inputs and of out */
    int B, Nc, Ns, dqk, dv;
    const void *c_x, *s_x, *c_1x, *s_1x;
    void* out;
    void* ws;
    size_t ws_bytes;
} mhada_forloss_args;
MHADA_API size_t mhada_forloss_workspace(int B, int Nc, int Ns, int dqk, int dv);
MHADA_API int mhada_forloss_forward(const mhada_forloss_args* args, mhada_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (9) Backward of one layer -- replaces what torch.autograd does for AdaAttnMultiHead.forward,
 *     MHAdaSTr/network/adaDecoder.py:162-206, inside the training step train_image.py:105-144 (loss.backward()).
 *     MHADA_BF16 path, head_dim 64, softmax activation, out_conv present.  The forward intermediates are RECOMPUTED
 *     (statistics, folded projections, attention) -- nothing but the layer inputs has to be kept from the forward.
 *     fc, fs, fcs, d_out: bf16 token-major as in (5) (fcs may alias fc);  weights as in (5), w_out f32 [C][C].
 *     Gradients: d_fc, d_fcs f32 [B, Nc, C], d_fs f32 [B, Ns, C] (when fcs aliases fc the caller adds d_fc + d_fcs);
 *     d_w_fgh f32 [3][H][d][d], d_b_fgh f32 [3][H][d], d_w_out f32 [C][C], d_b_out f32 [C].
 *     The attention part is a FlashAttention-style backward with value operand V' = [V~ | V~^2] (csrc/attn_bwd.cu);
 *     every other contraction runs on the tcgen05 token GEMM.  Deterministic (no atomics).
 *     ws: mhada_layer_backward_workspace(B, Nc, Ns, C, H) bytes.
 * ---------------------------------------------------------------------------------------------- */
typedef struct mhada_layer_bwd_args {
    int B, Nc, Ns, C, H;
    const void *fc, *fs, *fcs;
    const float *w_fgh, *b_fgh, *w_out, *b_out;
    const void* d_out;
    float *d_fc, *d_fs, *d_fcs;
    float *d_w_fgh, *d_b_fgh, *d_w_out, *d_b_out;
    void* ws;
    size_t ws_bytes;
} mhada_layer_bwd_args;
MHADA_API size_t mhada_layer_backward_workspace(int B, int Nc, int Ns, int C, int H);
MHADA_API int mhada_layer_backward(const mhada_layer_bwd_args* args, mhada_stream_t stream);
/*     The attention stage of (9) alone (stage tests): q (pre-multiplied by log2 e), k bf16 [B, N, C], v bf16 [B, Ns, 2C]
 *     as mhada_proj writes them, x = fcs bf16 with its statistics, g = dL/d(heads) f32 [B, Nc, C].  Outputs: d_o bf16
 *     [B, Nc, 4C] (per head [dM~ | dE] as two bf16 terms: hi 128 columns, lo 128 columns), lse (log2 units) and delta f32 [B, H, Nc], d_xhat = dL/d(IN(fcs)) f32 [B, Nc, C],
 *     d_q, d_k, d_v bf16 (natural units; d_v is the gradient of V, the squares already folded in). */
MHADA_API int mhada_attn_bwd(int B, int H, int Nc, int Ns, const void* q, const void* k, const void* v, const void* x,
                             const float* x_mean, const float* x_rstd, const float* g, void* d_o, float* lse, float* delta,
                             float* d_xhat, void* d_q, void* d_k, void* d_v, mhada_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (10) Helpers of the training path (backward of y = x W^T + b on the token GEMM): dW = dy^T x needs both operands with
 *     the TOKEN axis contiguous, db = column sums of dy.
 *     mhada_transpose_bf16: x [M, ld] (MHADA_F32 or MHADA_BF16) -> out bf16 [C][Mpad], out[c][m] = x[m][c], columns
 *                           M <= m < Mpad zero (Mpad % 64 == 0 so that it can be the K axis of mhada_gemm_bf16);
 *     mhada_colsum:         out f32 [C] = sum over the M rows of x [M, C] (two-stage, deterministic);
 *                           C even, <= 2048; ws: mhada_colsum_workspace(M, C) bytes.
 * ---------------------------------------------------------------------------------------------- */
/*     mhada_gemm_bf16_splitk: out_f32 [M, ldf] = x . w^T like mhada_gemm_bf16 (f32 result only, no bias / residual), with the
 *                           K range split over extra work items of the same kernel when M x N has too few tiles to fill
 *                           the GPU (dW = dy^T x: a 512 x 512 result over K = 8192 tokens is 8 tiles); partial tiles go
 *                           to ws (mhada_gemm_splitk_workspace bytes; 0 or NULL = no split) and are added in a fixed
 *                           order. */
MHADA_API size_t mhada_gemm_splitk_workspace(int M, int N, int K);
MHADA_API int mhada_gemm_bf16_splitk(const void* x, int lda, const void* w, int ldw, int M, int N, int K, float* out_f32,
                                     int ldf, void* ws, size_t ws_bytes, mhada_stream_t stream);
MHADA_API int mhada_transpose_bf16(const void* x, int dtype, int ld, int M, int C, int Mpad, void* out, mhada_stream_t stream);
MHADA_API size_t mhada_colsum_workspace(int M, int C);
MHADA_API int mhada_colsum(const void* x, int dtype, int M, int C, void* ws, size_t ws_bytes, float* out, mhada_stream_t stream);

/* Number of kernel launches the last mhada_layer_forward on this thread issued (bench bookkeeping). */
MHADA_API int mhada_last_launch_count(void);
/* Kernel launches issued by this library from this thread since it was loaded (monotonic). */
MHADA_API long long mhada_total_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * (6) Kernel timing for the roofline line of bench.py (no reference counterpart).
 *     Between mhada_profile_begin() and mhada_profile_end() every attention launch issued from this
 *     thread (by mhada_attn or mhada_layer_forward) is bracketed by cudaEventRecord on the caller's
 *     stream.  mhada_profile_end() waits for those events and returns the summed device time (ms) and
 *     the number of launches.  Not graph-capturable while enabled.
 * ---------------------------------------------------------------------------------------------- */
/* Development aid: runs the bf16 attention kernel with clock64() stamps of CTA (0,0,0) written to
 * `trace` (device, 4*64*8 + 3*4096 int64; layout in csrc/attn_tc.cu).  Same results as mhada_attn. */
MHADA_API int mhada_debug_attn_trace(const mhada_attn_args* args, long long* trace, mhada_stream_t stream);
MHADA_API int mhada_profile_begin(void);
MHADA_API int mhada_profile_end(float* attn_ms_total, int* attn_launches);
/*     Per-stage totals of the last begin/end bracket (valid after mhada_profile_end): device time between the
 *     events that surround the launches of one stage of mhada_layer_forward(_cached), and how many such brackets. */
enum mhada_stage { MHADA_STAGE_STATS = 0, MHADA_STAGE_PROJ = 1, MHADA_STAGE_ATTN = 2, MHADA_STAGE_LINEAR = 3,
                   MHADA_STAGE_VIT = 4 /* one bracket per mhada_vit_forward */, MHADA_STAGE_COUNT = 5 };
MHADA_API int mhada_profile_stage(int stage, float* ms_total, int* brackets);

#ifdef __cplusplus
}
#endif
#endif /* MHADA_B200_H_ */

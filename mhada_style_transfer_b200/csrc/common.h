// Host-side plumbing shared by the kernel translation units: error reporting, device check,
// TMA tensor-map construction (through the driver entry point, so libcuda is not a link-time
// dependency and the library loads on a box without a GPU), launch counting.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/mhada_b200.h"

namespace mh {

void set_error(const char* fmt, ...);
const char* last_error();
int check_cuda(cudaError_t e, const char* what);   // 0 or the error (message recorded)
int device_check();

// Per-device one-time set-up (SURVEY.md 8b "keyed by device"): cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a
// PER-DEVICE property of a kernel, so a process that drives several GPUs must set it once on each; the SM count that
// sizes persistent grids is per device too.  DeviceOnce holds one done-bit per device ordinal (thread safe; a race
// only repeats the idempotent call).
struct DeviceOnce {
    unsigned long long done[4] = {0, 0, 0, 0};   // 256 device ordinals
};
int smem_attr_once(DeviceOnce& once, const void* kernel, size_t smem_bytes, const char* what);
int sm_count();   // multiprocessors of the CURRENT device (cached per device)

extern thread_local int g_launches;
extern thread_local long long g_launches_total;
inline void count_launch() { ++g_launches; ++g_launches_total; }

// bf16 tensor map, up to 3 dims, dims[0] innermost (contiguous).  strides_bytes[i] = byte pitch of
// dim i+1.  box[i] = tile extent.  128B swizzle when box[0]*2 == 128, else none.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box);

// same for 2-byte (bf16) or 4-byte (f32) elements
int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box);

// ---- launchers (one per kernel family); all return 0 / error code ------------------------------
size_t stats_workspace(int B, int N, int C);
int launch_stats(const void* x, int dtype, int B, int N, int C, int ld, float* mean, float* rstd, float* ws,
                 cudaStream_t s);
// up to 3 tensors with the same B, C, ld (fc, fs, fcs of a layer) in one launch of each pass
int launch_stats_multi(int n, const void* const* x, const int* N, float* const* mean, float* const* rstd, int dtype,
                       int B, int C, int ld, float* ws, cudaStream_t s);

// first pass of the statistics only (bf16 layer path: fold_stats_kernel finishes them)
struct StatsPartialInfo {
    int max_splits;
    int splits[3];
};
int launch_stats_partial(int n, const void* const* x, const int* N, int dtype, int B, int C, int ld, float* ws,
                         StatsPartialInfo* info, cudaStream_t s);

// f32 SIMT path
// parts: 1 = Q (content batch B), 2 = K, V (style batch Bs)
int launch_proj_f32(int parts, const float* fc, const float* fs, const float* mean_c, const float* rstd_c,
                    const float* mean_s, const float* rstd_s, const float* w, const float* bias, int B, int Bs, int Nc,
                    int Ns, int H, int d, float* q, float* k, float* v, float* mu_v, cudaStream_t s);
int launch_attn_f32(const mhada_attn_args& a, cudaStream_t s);
size_t attn_cosine_scratch_bytes(int B, int H);   // closed-form cosine moments (head_dim 64), see simt_f32.cu
int launch_linear_f32(const float* x, int ldx, const float* w, const float* bias, int M, int Cin, int Cout, float* y,
                      int ldy, cudaStream_t s);

// bf16 tcgen05 path
size_t proj_bf16_workspace(int B, int H, int d);
int launch_proj_bf16(int parts, const void* fc, const void* fs, const float* mean_c, const float* rstd_c,
                     const float* mean_s, const float* rstd_s, const float* w, const float* bias, int B, int Bs, int Nc,
                     int Ns, int H, int d, void* q, void* k, void* v, float* mu_v, void* ws, cudaStream_t s);
// Roles of fold_stats_kernel: finish the statistics of one tensor from the partial sums (ti >= 0) or take them from
// mean / rstd (ti < 0), write them, and fold them into one projection's weights (kind 0 f, 1 g, 2 h; 3 = none).
struct FoldStatsJob {
    int n_roles;
    int kind[4], ti[4], N[4], splits[4];
    const void* x[4];
    float *mean[4], *rstd[4];
    int max_splits;
};
int launch_fold_stats(const FoldStatsJob& job, const float* partial, const float* w, const float* bias, int B, int H,
                      int d, float* mu_v, void* proj_ws, cudaStream_t s);
int launch_proj_bf16_folded(int parts, const void* fc, const void* fs, int B, int Bs, int Nc, int Ns, int H, int d, void* q,
                            void* k, void* v, void* ws, cudaStream_t s);
int launch_attn_bf16(const mhada_attn_args& a, cudaStream_t s);
int launch_attn_bf16_impl(const mhada_attn_args& a, long long* trace, cudaStream_t s);   // trace: 3*64*8 slots or null
size_t linear_bf16_workspace(int Cout, int Cin);
int launch_linear_bf16(const void* x, int ldx, const float* w, const float* bias, int M, int Cin, int Cout, void* y,
                       int ldy, void* ws, cudaStream_t s);

// ViT encoder: generic token GEMM with fused epilogue (gemm_tc.cu) and the small kernels (vit.cu)
struct GemmDesc {
    const void* a; int lda;          // bf16 [M, lda]
    const void* w; int ldw;          // bf16 [N, ldw]  (nn.Linear / Conv2d weight layout: [out][in])
    const float* bias;               // [N] or nullptr
    int M, N, K;
    void* out_bf16; int ldo;         // bf16 result or nullptr
    float* out_f32; int ldf;         // f32 result or nullptr
    const float* resid; int ldr;     // f32 residual or nullptr, row = m % resid_mod when resid_mod > 0
    int resid_mod;
    int relu;
};
int launch_gemm_bf16(const GemmDesc& d, cudaStream_t s);
// same, f32 result only, K split over extra work items when the output has too few tiles to fill the GPU (weight gradients)
size_t gemm_splitk_workspace(int M, int N, int K);
int launch_gemm_bf16_splitk(const GemmDesc& d, void* ws, size_t ws_bytes, cudaStream_t s);
int launch_patch_im2col(int img_dtype, const void* img, int B, int Himg, int Wimg, int patch, void* a0, cudaStream_t s);
int launch_layernorm(const float* x, int M, int C, const float* gamma, const float* beta, float eps, void* y_bf16,
                     cudaStream_t s);
int launch_batch_attn(const void* qkv, int B, int N, int heads, int hd, void* out, cudaStream_t s);
int launch_batch_attn_bwd(const void* qkv, const void* dout, int B, int N, int heads, int hd, void* dqkv, cudaStream_t s);
size_t vit_workspace(int B, int N, int D, int F, int K0);
int vit_forward(const mhada_vit_args& a, cudaStream_t s);

// AdaAttnForLoss on the tensor cores (forloss_tc.cu)
size_t forloss_workspace(int B, int Nc, int Ns, int dqk, int dv);
int forloss_forward(const mhada_forloss_args& a, cudaStream_t s);
// layers with wide heads (head_dim a multiple of 128 above 128) on the same materialised core
size_t layer_wide_workspace(int B, int Nc, int Ns, int C, int H);
int layer_wide_attention(const void* fc, const void* fs, const void* fcs, const float* mean_c, const float* rstd_c,
                         const float* mean_s, const float* rstd_s, const float* mean_x, const float* rstd_x, const float* w_fgh,
                         const float* b_fgh, int B, int Nc, int Ns, int C, int H, void* heads, void* ws, cudaStream_t s);

// backward of a layer: attention backward kernels (attn_bwd.cu) and the helpers around the GEMMs (layer_bwd.cu)
int launch_attn_bwd(int B, int H, int Nc, int Ns, int C, const void* q, const void* k, const void* v, const void* x,
                    const float* x_mean, const float* x_rstd, const float* g, void* d_o, float* lse, float* delta,
                    float* d_xhat, void* d_q, void* d_k, void* d_v, cudaStream_t s);
int launch_transpose_norm(const void* in, int in_dtype, int ld, int M, int Mpad, int C, int N, const float* mean, const float* rstd,
                          void* out, cudaStream_t s);
int launch_blockdiag_t(const float* w, int H, int d, void* out, cudaStream_t s);
int launch_extract_blockdiag(const float* full, int H, int d, float* dw, cudaStream_t s);
size_t token_sums_workspace(int B, int N, int C);
int launch_token_sums(const void* g, int g_dtype, const void* x, const float* mean, const float* rstd, int B, int N, int C,
                      void* partial, void* sums, float* bias, cudaStream_t s);
int launch_in_bwd_apply(const float* g, const void* x, const float* mean, const float* rstd, const void* sums, const float* add,
                        int B, int N, int C, float* dx, cudaStream_t s);

// decoder blocks 0..7: reflect-padded NHWC input -> conv3x3 + bias + ReLU on tcgen05 (conv_tc.cu)
int launch_conv3x3_tc(const void* xp, const void* w, const float* bias, int B, int H, int W, int Cin, int Cout, int relu,
                      int out_padded, void* y, cudaStream_t s);
// decoder: last block (reflect pad + 3x3 conv to <= 8 channels + ReLU) in one kernel
int launch_conv3x3_small(const void* x, const float* w, const float* bias, int B, int H, int W, int Cin, int Cout, int relu,
                         void* y, cudaStream_t s);
// decoder glue
int launch_pad_reflect(int dtype, const void* x, int B, int H, int W, int C, int upsample, void* y, cudaStream_t s);
int launch_pad_reflect_bwd(int dtype, const void* gp, int B, int H, int W, int C, int upsample, void* dx, cudaStream_t s);

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace mh

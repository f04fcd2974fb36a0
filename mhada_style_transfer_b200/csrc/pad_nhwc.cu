// Decoder glue (first step): ReflectionPad2d(1) and the x2 bilinear up-sample that
// precede the decoder's 3x3 convolutions, fused into ONE pass over channels_last activations.
//
// Replaces nn.ReflectionPad2d(1) (MHAdaSTr/network/conv.py:26-27, run before every conv) and
// F.interpolate(scale_factor=2, mode="bilinear", align_corners=False) (conv.py:71, after three of the
// nine convs).  In the reference (and in stock PyTorch on the GPU) these are separate memory-bound
// kernels with layout round trips; on the first B200 profile they were 2/3 of the whole step
// (profiles/r01_launches_summary.md).  Here each activation is read once and written once, already
// padded (and up-sampled), in the layout cuDNN's NHWC implicit-GEMM kernels consume with padding 0.
//
// x: [B, H, W, C] -> y: [B, Ho + 2, Wo + 2, C],  Ho = H or 2H.  One thread per 16-byte channel vector
// of one output pixel; blends are done in fp32 and rounded once (like ATen's upsample for bf16).
// HBM-bound: algorithmic bytes = |x| + |y|.
#include "common.h"
#include "ptx.cuh"

namespace mh {

template <typename T> struct PadVec;
template <> struct PadVec<float> {
    static constexpr int VEC = 4;
    __device__ static void load(const float* p, float (&v)[4]) {
        float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ static void store(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct PadVec<__nv_bfloat16> {
    static constexpr int VEC = 8;
    __device__ static void load(const __nv_bfloat16* p, float (&v)[8]) {
        uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
        v[0] = bf16_lo(t.x); v[1] = bf16_hi(t.x); v[2] = bf16_lo(t.y); v[3] = bf16_hi(t.y);
        v[4] = bf16_lo(t.z); v[5] = bf16_hi(t.z); v[6] = bf16_lo(t.w); v[7] = bf16_hi(t.w);
    }
    __device__ static void store(__nv_bfloat16* p, const float (&v)[8]) {
        *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                                  pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
};

__device__ __forceinline__ int reflect1(int i, int n) {   // ReflectionPad2d(1): -1 -> 1, n -> n-2
    if (i < 0) return -i;
    if (i >= n) return 2 * n - 2 - i;
    return i;
}

template <typename T, bool UP>
__global__ void __launch_bounds__(256) pad_reflect_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int H,
                                                          int W, int C) {
    constexpr int VEC = PadVec<T>::VEC;
    const int Ho = UP ? 2 * H : H, Wo = UP ? 2 * W : W;
    const int cv = C / VEC;
    const size_t total = static_cast<size_t>(B) * (Ho + 2) * (Wo + 2) * cv;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += stride) {
        const int c = static_cast<int>(i % cv) * VEC;
        size_t pix = i / cv;
        const int xp = static_cast<int>(pix % (Wo + 2));
        pix /= (Wo + 2);
        const int yp = static_cast<int>(pix % (Ho + 2));
        const int b = static_cast<int>(pix / (Ho + 2));
        const int Y = reflect1(yp - 1, Ho), X = reflect1(xp - 1, Wo);
        const T* xb = x + static_cast<size_t>(b) * H * W * C + c;
        float out[VEC];
        if (!UP) {
            PadVec<T>::load(xb + (static_cast<size_t>(Y) * W + X) * C, out);
        } else {
            // align_corners=False: src = (dst + 0.5) / 2 - 0.5, clamped at 0 (conv.py:71)
            const float sy = fmaxf((Y + 0.5f) * 0.5f - 0.5f, 0.f), sx = fmaxf((X + 0.5f) * 0.5f - 0.5f, 0.f);
            const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
            const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
            const float ly = sy - y0, lx = sx - x0;
            float v00[VEC], v01[VEC], v10[VEC], v11[VEC];
            PadVec<T>::load(xb + (static_cast<size_t>(y0) * W + x0) * C, v00);
            PadVec<T>::load(xb + (static_cast<size_t>(y0) * W + x1) * C, v01);
            PadVec<T>::load(xb + (static_cast<size_t>(y1) * W + x0) * C, v10);
            PadVec<T>::load(xb + (static_cast<size_t>(y1) * W + x1) * C, v11);
#pragma unroll
            for (int k = 0; k < VEC; ++k)
                out[k] = (1.f - ly) * ((1.f - lx) * v00[k] + lx * v01[k]) + ly * ((1.f - lx) * v10[k] + lx * v11[k]);
        }
        PadVec<T>::store(y + ((static_cast<size_t>(b) * (Ho + 2) + yp) * (Wo + 2) + xp) * C + c, out);
    }
}

int launch_pad_reflect(int dtype, const void* x, int B, int H, int W, int C, int upsample, void* y, cudaStream_t s) {
    const int Ho = upsample ? 2 * H : H, Wo = upsample ? 2 * W : W;
    const int vec = dtype == MHADA_BF16 ? 8 : 4;
    const size_t total = static_cast<size_t>(B) * (Ho + 2) * (Wo + 2) * (C / vec);
    size_t blocks = (total + 255) / 256;
    const size_t cap = 148 * 16;           // grid-stride beyond 16 CTAs per SM
    if (blocks > cap) blocks = cap;
    const unsigned g = static_cast<unsigned>(blocks);
    if (dtype == MHADA_BF16) {
        auto xi = static_cast<const __nv_bfloat16*>(x);
        auto yo = static_cast<__nv_bfloat16*>(y);
        if (upsample) pad_reflect_kernel<__nv_bfloat16, true><<<g, 256, 0, s>>>(xi, yo, B, H, W, C);
        else pad_reflect_kernel<__nv_bfloat16, false><<<g, 256, 0, s>>>(xi, yo, B, H, W, C);
    } else {
        auto xi = static_cast<const float*>(x);
        auto yo = static_cast<float*>(y);
        if (upsample) pad_reflect_kernel<float, true><<<g, 256, 0, s>>>(xi, yo, B, H, W, C);
        else pad_reflect_kernel<float, false><<<g, 256, 0, s>>>(xi, yo, B, H, W, C);
    }
    count_launch();
    return check_cuda(cudaGetLastError(), "pad_reflect launch");
}

}  // namespace mh

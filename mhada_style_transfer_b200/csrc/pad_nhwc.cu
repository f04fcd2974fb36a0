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
// x: [B, H, W, C] -> y: [B, Ho + 2, Wo + 2, C],  Ho = H or 2H.  One CTA per padded output row, 16-byte channel
// vectors; blends are done in fp32 and rounded once (like ATen's upsample for bf16).
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

// grid (Ho + 2, B, xsplit): one CTA per padded output row (or per slice of it).  Everything that depends on the row (reflected source row, the
// two source rows and the vertical blend weight of the up-sample) is computed once per CTA; a thread then walks the
// row's 16-byte vectors with four independent vectors in flight (the first version spent its time in per-vector
// div / mod chains and ran at 1.3 - 2.7 TB/s).
template <typename T, bool UP>
__global__ void __launch_bounds__(256) pad_reflect_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int H,
                                                          int W, int C, int cv_shift) {
    constexpr int VEC = PadVec<T>::VEC;
    const int Ho = UP ? 2 * H : H, Wo = UP ? 2 * W : W;
    const int cv = C / VEC;
    const int yp = blockIdx.x, b = blockIdx.y;
    const int Y = reflect1(yp - 1, Ho);
    const T* xb = x + static_cast<size_t>(b) * H * W * C;
    T* yrow = y + (static_cast<size_t>(b) * (Ho + 2) + yp) * (Wo + 2) * C;
    const int nvec = (Wo + 2) * cv;
    // vertical part of the bilinear blend: align_corners=False, src = (dst + 0.5) / 2 - 0.5 clamped at 0 (conv.py:71)
    const float sy = UP ? fmaxf((Y + 0.5f) * 0.5f - 0.5f, 0.f) : 0.f;
    const int y0 = UP ? static_cast<int>(sy) : Y;
    const int y1 = UP ? min(y0 + 1, H - 1) : Y;
    const float ly = sy - y0;
    const T* r0 = xb + static_cast<size_t>(y0) * W * C;
    const T* r1 = xb + static_cast<size_t>(y1) * W * C;
    auto one = [&](int v) {
        int xp, c;
        if (cv_shift >= 0) { xp = v >> cv_shift; c = (v & (cv - 1)) * VEC; }
        else { xp = v / cv; c = (v % cv) * VEC; }
        const int X = reflect1(xp - 1, Wo);
        float out[VEC];
        if (!UP) {
            PadVec<T>::load(r0 + static_cast<size_t>(X) * C + c, out);
        } else {
            const float sx = fmaxf((X + 0.5f) * 0.5f - 0.5f, 0.f);
            const int x0 = static_cast<int>(sx);
            const int x1 = min(x0 + 1, W - 1);
            const float lx = sx - x0;
            float v00[VEC], v01[VEC], v10[VEC], v11[VEC];
            PadVec<T>::load(r0 + static_cast<size_t>(x0) * C + c, v00);
            PadVec<T>::load(r0 + static_cast<size_t>(x1) * C + c, v01);
            PadVec<T>::load(r1 + static_cast<size_t>(x0) * C + c, v10);
            PadVec<T>::load(r1 + static_cast<size_t>(x1) * C + c, v11);
#pragma unroll
            for (int k = 0; k < VEC; ++k)
                out[k] = (1.f - ly) * ((1.f - lx) * v00[k] + lx * v01[k]) + ly * ((1.f - lx) * v10[k] + lx * v11[k]);
        }
        PadVec<T>::store(yrow + static_cast<size_t>(v) * VEC, out);
    };
    // gridDim.z CTAs share a row when the image is small (keeps >= ~2000 CTAs in flight)
    const int per = (nvec + gridDim.z - 1) / gridDim.z;
    const int vend = min(nvec, (static_cast<int>(blockIdx.z) + 1) * per);
    int v = blockIdx.z * per + threadIdx.x;
    for (; v + 3 * 256 < vend; v += 4 * 256) {
#pragma unroll
        for (int u = 0; u < 4; ++u) one(v + u * 256);
    }
    for (; v < vend; v += 256) one(v);
}

int launch_pad_reflect(int dtype, const void* x, int B, int H, int W, int C, int upsample, void* y, cudaStream_t s) {
    const int Ho = upsample ? 2 * H : H;
    const int vec = dtype == MHADA_BF16 ? 8 : 4;
    const int cv = C / vec;
    int cv_shift = -1;
    if (cv > 0 && (cv & (cv - 1)) == 0) {
        cv_shift = 0;
        while ((1 << cv_shift) < cv) ++cv_shift;
    }
    int xsplit = (2048 + (Ho + 2) * B - 1) / ((Ho + 2) * B);
    if (xsplit < 1) xsplit = 1;
    if (xsplit > 8) xsplit = 8;
    dim3 g(static_cast<unsigned>(Ho + 2), static_cast<unsigned>(B), static_cast<unsigned>(xsplit));
    if (dtype == MHADA_BF16) {
        auto xi = static_cast<const __nv_bfloat16*>(x);
        auto yo = static_cast<__nv_bfloat16*>(y);
        if (upsample) pad_reflect_kernel<__nv_bfloat16, true><<<g, 256, 0, s>>>(xi, yo, B, H, W, C, cv_shift);
        else pad_reflect_kernel<__nv_bfloat16, false><<<g, 256, 0, s>>>(xi, yo, B, H, W, C, cv_shift);
    } else {
        auto xi = static_cast<const float*>(x);
        auto yo = static_cast<float*>(y);
        if (upsample) pad_reflect_kernel<float, true><<<g, 256, 0, s>>>(xi, yo, B, H, W, C, cv_shift);
        else pad_reflect_kernel<float, false><<<g, 256, 0, s>>>(xi, yo, B, H, W, C, cv_shift);
    }
    count_launch();
    return check_cuda(cudaGetLastError(), "pad_reflect launch");
}

}  // namespace mh

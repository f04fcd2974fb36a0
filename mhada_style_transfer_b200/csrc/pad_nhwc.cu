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

// ---- backward of the same op (training path of the decoder, train_image.py:139): dx[B, H, W, C] from the gradient of the
// padded (and up-sampled) map dyp[B, Ho + 2, Wo + 2, C].  Gather form, one thread per (pixel, 16-byte channel vector), no
// atomics: the gradient of pixel (u, v) of the UNPADDED up-sampled image is the interior value (u + 1, v + 1) plus the
// ring positions ReflectionPad2d(1) filled from it (row 1 -> padded row 0, row Ho - 2 -> padded row Ho + 1, same for
// columns); a source pixel i receives 0.75 / 0.25 of the up-sampled rows 2i-1 .. 2i+2 (align_corners = False, scale 2:
// row 2k blends 0.25 (k-1) + 0.75 k, row 2k+1 blends 0.75 k + 0.25 min(k+1, H-1), row 0 copies row 0).
template <typename T, bool UP>
__global__ void __launch_bounds__(256) pad_reflect_bwd_kernel(const T* __restrict__ gp, T* __restrict__ dx, int B, int H, int W,
                                                              int C) {
    constexpr int VEC = PadVec<T>::VEC;
    const int Ho = UP ? 2 * H : H, Wo = UP ? 2 * W : W;
    const int cv = C / VEC;
    const size_t t = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (t >= static_cast<size_t>(B) * H * W * cv) return;
    const int c = static_cast<int>(t % cv) * VEC;
    const int j = static_cast<int>((t / cv) % W), i = static_cast<int>((t / (static_cast<size_t>(cv) * W)) % H);
    const int b = static_cast<int>(t / (static_cast<size_t>(cv) * W * H));
    const T* gb = gp + static_cast<size_t>(b) * (Ho + 2) * (Wo + 2) * C + c;
    float acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
    auto folded = [&](int u, int v, float w) {          // acc += w * d(unpadded)[u][v]
        int ys[3], xs[3], ny = 0, nx = 0;
        ys[ny++] = u + 1;
        if (u == 1) ys[ny++] = 0;
        if (u == Ho - 2) ys[ny++] = Ho + 1;
        xs[nx++] = v + 1;
        if (v == 1) xs[nx++] = 0;
        if (v == Wo - 2) xs[nx++] = Wo + 1;
        for (int a = 0; a < ny; ++a)
            for (int e = 0; e < nx; ++e) {
                float val[VEC];
                PadVec<T>::load(gb + (static_cast<size_t>(ys[a]) * (Wo + 2) + xs[e]) * C, val);
#pragma unroll
                for (int k = 0; k < VEC; ++k) acc[k] = fmaf(w, val[k], acc[k]);
            }
    };
    if (!UP) {
        folded(i, j, 1.f);
    } else {
        int us[4], vs[4], nu = 0, nv = 0;
        float wu[4], wv[4];
        us[nu] = 2 * i; wu[nu++] = i == 0 ? 1.f : 0.75f;
        us[nu] = 2 * i + 1; wu[nu++] = i == H - 1 ? 1.f : 0.75f;
        if (i >= 1) { us[nu] = 2 * i - 1; wu[nu++] = 0.25f; }
        if (i + 1 <= H - 1) { us[nu] = 2 * i + 2; wu[nu++] = 0.25f; }
        vs[nv] = 2 * j; wv[nv++] = j == 0 ? 1.f : 0.75f;
        vs[nv] = 2 * j + 1; wv[nv++] = j == W - 1 ? 1.f : 0.75f;
        if (j >= 1) { vs[nv] = 2 * j - 1; wv[nv++] = 0.25f; }
        if (j + 1 <= W - 1) { vs[nv] = 2 * j + 2; wv[nv++] = 0.25f; }
        for (int a = 0; a < nu; ++a)
            for (int e = 0; e < nv; ++e) folded(us[a], vs[e], wu[a] * wv[e]);
    }
    PadVec<T>::store(dx + ((static_cast<size_t>(b) * H + i) * W + j) * C + c, acc);
}

int launch_pad_reflect_bwd(int dtype, const void* gp, int B, int H, int W, int C, int upsample, void* dx, cudaStream_t s) {
    const int vec = dtype == MHADA_BF16 ? 8 : 4;
    const size_t n = static_cast<size_t>(B) * H * W * (C / vec);
    const unsigned grid = static_cast<unsigned>((n + 255) / 256);
    if (dtype == MHADA_BF16) {
        auto g = static_cast<const __nv_bfloat16*>(gp);
        auto o = static_cast<__nv_bfloat16*>(dx);
        if (upsample) pad_reflect_bwd_kernel<__nv_bfloat16, true><<<grid, 256, 0, s>>>(g, o, B, H, W, C);
        else pad_reflect_bwd_kernel<__nv_bfloat16, false><<<grid, 256, 0, s>>>(g, o, B, H, W, C);
    } else {
        auto g = static_cast<const float*>(gp);
        auto o = static_cast<float*>(dx);
        if (upsample) pad_reflect_bwd_kernel<float, true><<<grid, 256, 0, s>>>(g, o, B, H, W, C);
        else pad_reflect_bwd_kernel<float, false><<<grid, 256, 0, s>>>(g, o, B, H, W, C);
    }
    count_launch();
    return check_cuda(cudaGetLastError(), "pad_reflect_bwd launch");
}

}  // namespace mh

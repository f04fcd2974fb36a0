// Error reporting, device check and TMA tensor-map construction.
#include <stdarg.h>
#include <stdio.h>
#include <mutex>

#include "common.h"

namespace mh {

static thread_local char g_err[512] = "";
thread_local int g_launches = 0;
thread_local long long g_launches_total = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    set_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return static_cast<int>(e);
}

int device_check() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_error("no CUDA device: %s", cudaGetErrorString(e));
        (void)cudaGetLastError();
        return MHADA_ERR_DEVICE;
    }
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    int minor = -1;
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    // built with -gencode arch=compute_100a,code=sm_100a only: arch-specific ("a") code is not forward compatible,
    // so sm_103 (B300) would pass a major-only check and then fail with "no kernel image" at the first launch
    if (e != cudaSuccess || major != 10 || minor != 0) {
        set_error("libmhada_b200 is built for sm_100a (B200, compute capability 10.0) only; found %d.%d", major, minor);
        (void)cudaGetLastError();
        return MHADA_ERR_DEVICE;
    }
    return 0;
}

int smem_attr_once(DeviceOnce& once, const void* kernel, size_t smem_bytes, const char* what) {
    int dev = 0;
    if (int e = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return e;
    const bool tracked = dev >= 0 && dev < 256;
    if (tracked && (__atomic_load_n(&once.done[dev >> 6], __ATOMIC_ACQUIRE) >> (dev & 63)) & 1ull) return 0;
    if (int e = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes)),
                           what))
        return e;
    if (tracked) __atomic_fetch_or(&once.done[dev >> 6], 1ull << (dev & 63), __ATOMIC_RELEASE);
    return 0;
}

int sm_count() {
    static int cache[256];   // 0 = unknown
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    const bool tracked = dev >= 0 && dev < 256;
    if (tracked) {
        const int c = __atomic_load_n(&cache[dev], __ATOMIC_RELAXED);
        if (c > 0) return c;
    }
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    if (tracked) __atomic_store_n(&cache[dev], n, __ATOMIC_RELAXED);
    return n;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
        (void)cudaGetLastError();
    });
    return fn;
}

int make_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return MHADA_ERR_DRIVER;
    }
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t bx[5], es[5];
    if (rank < 1 || rank > 5) {
        set_error("make_tmap: rank %d", rank);
        return MHADA_ERR_ARG;
    }
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
    }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUtensorMapSwizzle sw = (box[0] * elem_bytes == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
    const CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUresult r = enc(out, dt, static_cast<cuuint32_t>(rank), const_cast<void*>(base),
                     gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu,%llu,%llu box %u,%u,%u)",
                  static_cast<int>(r), rank, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
                  (unsigned long long)(rank > 2 ? gdim[2] : 0), bx[0], rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0);
        return MHADA_ERR_DRIVER;
    }
    return 0;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
    return make_tmap(out, base, 2, rank, dims, strides_bytes, box);
}

}  // namespace mh

// Host-side plumbing shared by every C-ABI entry point: last-error string, device query, TMA descriptor
// encoding through the driver entry point (fetched at run time so that the shared object has no link-time
// dependency on libcuda and still dlopen()s on a CPU-only box).
#include "common.cuh"
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <string.h>
#include <mutex>

namespace xcp {

static thread_local char g_err[512] = "";

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    set_last_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return (int)e;
}

int num_sms() {
    static int sms[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (sms[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        sms[dev] = v;
    }
    return sms[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

static CUtensorMapSwizzle swz(int bytes) {
    switch (bytes) {
        case 128: return CU_TENSOR_MAP_SWIZZLE_128B;
        case 64: return CU_TENSOR_MAP_SWIZZLE_64B;
        case 32: return CU_TENSOR_MAP_SWIZZLE_32B;
        default: return CU_TENSOR_MAP_SWIZZLE_NONE;
    }
}

int make_tmap_2d(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint64_t row_pitch_bytes,
                 uint32_t box_cols, uint32_t box_rows, int swizzle_bytes) {
    EncodeTiledFn enc = get_encode();
    XCP_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled driver entry point unavailable");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {row_pitch_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz(swizzle_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    XCP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(2d) failed: CUresult %d (cols %llu rows %llu pitch %llu box %ux%u)",
                (int)r, (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)row_pitch_bytes, box_cols,
                box_rows);
    return 0;
}

int make_tmap_4d(CUtensorMap* map, const void* base, const uint64_t dims_[4], const uint64_t strides_bytes[3],
                 const uint32_t box_[4], int swizzle_bytes) {
    return make_tmap_4d_l2(map, base, dims_, strides_bytes, box_, swizzle_bytes, 256);
}

int make_tmap_4d_l2(CUtensorMap* map, const void* base, const uint64_t dims_[4], const uint64_t strides_bytes[3],
                    const uint32_t box_[4], int swizzle_bytes, int l2_promotion_bytes) {
    EncodeTiledFn enc = get_encode();
    XCP_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled driver entry point unavailable");
    cuuint64_t dims[4] = {dims_[0], dims_[1], dims_[2], dims_[3]};
    cuuint64_t strides[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
    cuuint32_t box[4] = {box_[0], box_[1], box_[2], box_[3]};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz(swizzle_bytes),
                     l2_promotion_bytes >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                     : l2_promotion_bytes >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                     : l2_promotion_bytes >= 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    XCP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(4d) failed: CUresult %d", (int)r);
    return 0;
}

}  // namespace xcp

extern "C" const char* xcp_last_error_string(void) { return xcp::g_err; }

extern "C" int xcp_version(void) { return 100; }

// 0 when the current device is an sm_100 part; the library refuses to run anywhere else (no fallback).
extern "C" int xcp_check_device(int device) {
    cudaDeviceProp prop;
    XCP_CUDA(cudaGetDeviceProperties(&prop, device));
    XCP_REQUIRE(prop.major == 10, "device %d is sm_%d%d; this library only contains sm_100a code", device, prop.major,
                prop.minor);
    return 0;
}

#include "mmx_host.cuh"

#include <cstring>
#include <mutex>
#include <utility>
#include <vector>

#include "../../include/mmx.h"

namespace mmx {

char* error_buffer() {
    static thread_local char buf[1024] = {0};
    return buf;
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_buffer(), 1024, fmt, ap);
    va_end(ap);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess) {
            fn = reinterpret_cast<EncodeTiledFn>(p);
        }
    });
    return fn;
}

int make_tmap_bf16_sw128(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                         uint32_t box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    MMX_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    MMX_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
    MMX_REQUIRE((ld * 2) % 16 == 0, "TMA row pitch must be a multiple of 16 bytes (ld=%llu)",
                (unsigned long long)ld);
    MMX_REQUIRE(box_rows >= 1 && box_rows <= 256, "TMA box rows out of range");
    cuuint64_t gdim[2] = {cols, rows ? rows : 1};
    cuuint64_t gstride[1] = {ld * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MMX_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

bool first_use_on_device(const void* tag) {
    static std::mutex mu;
    static std::vector<std::pair<const void*, int>> seen;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    for (const auto& e : seen)
        if (e.first == tag && e.second == dev) return false;
    seen.emplace_back(tag, dev);
    return true;
}

int sm_count() {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
}

}  // namespace mmx

extern "C" const char* mmx_last_error(void) { return mmx::error_buffer(); }
extern "C" int mmx_abi_version(void) { return MMX_ABI_VERSION; }

extern "C" int mmx_device_check(void) {
    int dev = 0;
    MMX_CHECK_CUDA(cudaGetDevice(&dev));
    int major = 0, minor = 0;
    MMX_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    MMX_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    MMX_REQUIRE(major == 10, "libmmx is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
    return 0;
}

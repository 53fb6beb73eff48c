// Host-side helpers shared by every translation unit of libmmx.so: error string, CUDA call checking,
// and TMA tensor-map construction (the driver entry point is resolved at run time so the library links
// against cudart only and can be built on a box without libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

namespace mmx {

// thread-local message returned by mmx_last_error()
char* error_buffer();
void set_error(const char* fmt, ...);

#define MMX_CHECK_CUDA(expr)                                                                        \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            ::mmx::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return -1;                                                                              \
        }                                                                                           \
    } while (0)

#define MMX_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            ::mmx::set_error(__VA_ARGS__); \
            return -2;                    \
        }                                 \
    } while (0)

// Row-major bf16 matrix [rows, cols] with leading dimension ld (elements); box = box_rows x 64 columns,
// 128-byte swizzle (what umma_desc_kmajor_sw128 expects). Out-of-range rows/columns read as zero.
int make_tmap_bf16_sw128(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                         uint32_t box_rows);

int sm_count();

// true the first time it is called for `tag` on the CURRENT device (kernel attributes such as the dynamic shared-memory
// limit are per device: a process that drives several GPUs has to set them on each one)
bool first_use_on_device(const void* tag);

}  // namespace mmx

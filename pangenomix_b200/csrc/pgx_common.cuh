// Shared helpers for libpgx_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <algorithm>
#include <atomic>

#include "pgx.h"

namespace pgx {

extern thread_local char g_error[512];
extern std::atomic<long long> g_launches;

int fail(int code, const char *fmt, ...);

#define PGX_CUDA(expr)                                                                   \
    do {                                                                                 \
        cudaError_t err__ = (expr);                                                      \
        if (err__ != cudaSuccess)                                                        \
            return ::pgx::fail(PGX_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,             \
                               cudaGetErrorString(err__), __FILE__, __LINE__);           \
    } while (0)

#define PGX_LAUNCH_CHECK(name)                                                           \
    do {                                                                                 \
        cudaError_t err__ = cudaGetLastError();                                          \
        if (err__ != cudaSuccess)                                                        \
            return ::pgx::fail(PGX_ERR_CUDA, "launch of %s failed: %s", name,            \
                               cudaGetErrorString(err__));                               \
        ::pgx::g_launches.fetch_add(1, std::memory_order_relaxed);                       \
    } while (0)

constexpr unsigned FULL_MASK = 0xffffffffu;

// Device-side bounds checks for debug builds (python -m pangenomix_b200.build --debug): compute-sanitizer
// is not available on every pool, so index invariants of the kernels can be asserted instead.
#ifdef PGX_DEBUG_BOUNDS
#define PGX_DEVICE_CHECK(cond)                                                                  \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            printf("PGX_DEVICE_CHECK failed: %s (%s:%d) block %d thread %d\n", #cond, __FILE__, __LINE__, \
                   static_cast<int>(blockIdx.x), static_cast<int>(threadIdx.x));                \
            __trap();                                                                           \
        }                                                                                       \
    } while (0)
#else
#define PGX_DEVICE_CHECK(cond) do { } while (0)
#endif

// 128-bit streaming load: the folded index chunks are read once per CTA pass and must
// not displace the rank table's neighbours in L1.
__device__ __forceinline__ uint4 ldg_stream(const uint4 *ptr)
{
    uint4 r;
    asm("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(ptr));
    return r;
}

}  // namespace pgx

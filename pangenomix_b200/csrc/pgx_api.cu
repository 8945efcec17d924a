// Library-wide pieces of libpgx_b200: error reporting, device query, launch counter.
// (The host-side numpy-legacy permutation stream lives in pgx_rng.cpp.)
#include <stdarg.h>
#include <string.h>

#include "pgx_common.cuh"

namespace pgx {

thread_local char g_error[512] = "";
std::atomic<long long> g_launches{0};

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

}  // namespace pgx

extern "C" {

int pgx_version(void) { return PGX_VERSION; }

const char *pgx_last_error(void) { return pgx::g_error; }

int64_t pgx_launch_count(void) { return pgx::g_launches.load(std::memory_order_relaxed); }

int pgx_device_info(int32_t *sm_count, int32_t *smem_optin_bytes, int64_t *l2_bytes)
{
    int dev = 0, v = 0;
    PGX_CUDA(cudaGetDevice(&dev));
    if (sm_count) {
        PGX_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
        *sm_count = v;
    }
    if (smem_optin_bytes) {
        PGX_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        *smem_optin_bytes = v;
    }
    if (l2_bytes) {
        PGX_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, dev));
        *l2_bytes = v;
    }
    return PGX_OK;
}

}  // extern "C"

// Library-wide pieces of libpgx_b200: error reporting, device query, launch counter and the
// host-side numpy-legacy permutation stream.
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

namespace {

// MT19937 (Matsumoto & Nishimura) in block form: refill all 624 state words, then temper
// the whole block at once (both loops vectorise).  numpy's legacy RandomState is this
// generator; ``pos`` counts the words of the current block already consumed (624 = block
// exhausted), exactly as np.random.get_state() reports it.
struct Mt19937 {
    uint32_t key[624];
    uint32_t out[624];
    int pos;

    void temper_block()
    {
        for (int k = 0; k < 624; ++k) {
            uint32_t y = key[k];
            y ^= y >> 11;
            y ^= (y << 7) & 0x9d2c5680u;
            y ^= (y << 15) & 0xefc60000u;
            y ^= y >> 18;
            out[k] = y;
        }
    }

    void refill()
    {
        constexpr uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MAGIC = 0x9908b0dfu;
        int k = 0;
        for (; k < 624 - 397; ++k) {
            const uint32_t y = (key[k] & UPPER) | (key[k + 1] & LOWER);
            key[k] = key[k + 397] ^ (y >> 1) ^ (-(y & 1u) & MAGIC);
        }
        for (; k < 623; ++k) {
            const uint32_t y = (key[k] & UPPER) | (key[k + 1] & LOWER);
            key[k] = key[k + (397 - 624)] ^ (y >> 1) ^ (-(y & 1u) & MAGIC);
        }
        const uint32_t y = (key[623] & UPPER) | (key[0] & LOWER);
        key[623] = key[396] ^ (y >> 1) ^ (-(y & 1u) & MAGIC);
        temper_block();
        pos = 0;
    }
};

}  // namespace

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

// np.arange(n) + np.random.shuffle (pangenome_analysis.py:84-85) for the legacy
// RandomState: Fisher-Yates from the top, j in [0, i] by masked rejection on successive
// 32-bit outputs (mask = smallest 2^b - 1 >= i).
int pgx_legacy_shuffles(uint32_t *mt_key, int32_t *mt_pos, int64_t n, int64_t count,
                        uint16_t *h_perms)
{
    if (!mt_key || !mt_pos || (!h_perms && n * count > 0))
        return pgx::fail(PGX_ERR_INVALID, "null pointer passed to pgx_legacy_shuffles");
    if (n < 0 || n > 65535) return pgx::fail(PGX_ERR_UNSUPPORTED, "n = %lld outside 0..65535", (long long)n);
    if (count < 0 || *mt_pos < 0 || *mt_pos > 624) return pgx::fail(PGX_ERR_INVALID, "bad MT19937 position / count");
    pgx::Mt19937 mt;
    memcpy(mt.key, mt_key, sizeof(mt.key));
    mt.pos = *mt_pos;
    mt.temper_block();
    for (int64_t t = 0; t < count; ++t) {
        uint16_t *a = h_perms + t * n;
        for (int64_t i = 0; i < n; ++i) a[i] = static_cast<uint16_t>(i);
        // Branch-free rejection: a rejected draw swaps a[i] with itself and leaves i alone.
        uint32_t i = n > 0 ? static_cast<uint32_t>(n - 1) : 0;
        while (i > 0) {
            if (mt.pos >= 624) mt.refill();
            const uint32_t mask = 0xffffffffu >> __builtin_clz(i);   // smallest 2^b - 1 >= i
            const uint32_t v = mt.out[mt.pos++] & mask;
            const uint32_t ok = v <= i;
            const uint32_t j = ok ? v : i;
            const uint16_t ai = a[i], aj = a[j];
            a[i] = aj;
            a[j] = ai;
            i -= ok;
        }
    }
    memcpy(mt_key, mt.key, sizeof(mt.key));
    *mt_pos = mt.pos;
    return PGX_OK;
}

}  // extern "C"

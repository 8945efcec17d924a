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

// ---- device plans owned by the library --------------------------------------------------------------
namespace {
struct DevicePlan {
    pgx_plan plan;                 // first member: the pointer handed to the caller
    uint64_t magic;
    int device;
    void *buffers[8];
};
constexpr uint64_t PLAN_MAGIC = 0x70677870'6c616e21ull;

template <typename T>
int upload(const T *src, int64_t count, void **slot, const T **dst)
{
    // empty arrays still get a valid, 16-byte aligned pointer (check_plan insists on alignment, not on size)
    const size_t bytes = sizeof(T) * static_cast<size_t>(std::max<int64_t>(count, 4));
    PGX_CUDA(cudaMalloc(slot, bytes));
    if (count > 0) PGX_CUDA(cudaMemcpy(*slot, src, sizeof(T) * static_cast<size_t>(count), cudaMemcpyHostToDevice));
    else PGX_CUDA(cudaMemset(*slot, 0, bytes));
    *dst = static_cast<const T *>(*slot);
    return PGX_OK;
}
}  // namespace

extern "C" {

int pgx_version(void) { return PGX_VERSION; }

const char *pgx_last_error(void) { return pgx::g_error; }

int64_t pgx_launch_count(void) { return pgx::g_launches.load(std::memory_order_relaxed); }

int pgx_plan_destroy(pgx_plan *plan)
{
    if (!plan) return PGX_OK;
    DevicePlan *dp = reinterpret_cast<DevicePlan *>(plan);
    if (dp->magic != PLAN_MAGIC) return pgx::fail(PGX_ERR_INVALID, "pgx_plan_destroy: not a plan made by pgx_plan_create / pgx_plan_upload");
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(dp->device);
    for (void *b : dp->buffers)
        if (b) cudaFree(b);
    cudaSetDevice(prev);
    dp->magic = 0;
    delete dp;
    return PGX_OK;
}

int pgx_plan_upload(const pgx_host_plan *h, pgx_plan **out)
{
    if (!h || !out) return pgx::fail(PGX_ERR_INVALID, "null pointer passed to pgx_plan_upload");
    *out = nullptr;
    DevicePlan *dp = new DevicePlan();
    memset(dp, 0, sizeof(*dp));
    dp->magic = PLAN_MAGIC;
    int rc = PGX_OK;
    if (cudaGetDevice(&dp->device) != cudaSuccess) rc = pgx::fail(PGX_ERR_CUDA, "no CUDA device");
    pgx_plan &p = dp->plan;
    if (!rc) rc = upload(h->chunks, h->n_chunks * 8, &dp->buffers[0], &p.d_chunks);
    if (!rc) rc = upload(h->tasks, static_cast<int64_t>(h->n_tasks) * 4, &dp->buffers[1], &p.d_tasks);
    if (!rc) rc = upload(h->sorted_idx, h->n_sorted, &dp->buffers[2], &p.d_sorted_idx);
    if (!rc) rc = upload(h->sorted_ptr, static_cast<int64_t>(h->n_rows) + 1, &dp->buffers[3], &p.d_sorted_ptr);
    if (!rc) rc = upload(h->bits, h->n_bits_words, &dp->buffers[4], &p.d_bits);
    if (!rc) rc = upload(h->colsum, h->n_genomes, &dp->buffers[5], &p.d_colsum);
    if (!rc) rc = upload(h->w_present, h->n_genomes, &dp->buffers[6], &p.d_w_present);
    if (!rc) rc = upload(h->w_absent, h->n_genomes, &dp->buffers[7], &p.d_w_absent);
    if (rc) {
        char text[512];
        snprintf(text, sizeof(text), "%s", pgx_last_error());
        pgx_plan_destroy(&dp->plan);
        return pgx::fail(rc, "%s", text);
    }
    p.n_chunks = h->n_chunks;
    p.n_genomes = h->n_genomes;
    p.n_genes = h->n_genes;
    p.n_rows = h->n_rows;
    p.n_tasks = h->n_tasks;
    p.n_long = h->n_long;
    p.n_superblocks = h->n_superblocks;
    p.perms_per_cta = h->perms_per_cta;
    p.slice_words = h->slice_words;
    p.max_colsum = h->max_colsum;
    *out = &dp->plan;
    return PGX_OK;
}

int pgx_plan_create(const int32_t *row, const int32_t *col, int64_t nnz, int32_t n_genes, int32_t n_genomes,
                    int32_t long_threshold, pgx_plan **out)
{
    if (!out) return pgx::fail(PGX_ERR_INVALID, "out is null");
    *out = nullptr;
    pgx_host_plan *host = nullptr;
    if (int rc = pgx_host_plan_create(row, col, nnz, n_genes, n_genomes, long_threshold, 0, 0, &host)) return rc;
    const int rc = pgx_plan_upload(host, out);
    pgx_host_plan_destroy(host);
    return rc;
}

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

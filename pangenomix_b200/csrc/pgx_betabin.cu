// Beta-binomial core estimate on sm_100a: the two data-parallel pieces of compute_beta_binomial_core_genome
// (/root/reference/pangenomix/pangenome_analysis.py:295-400).
//
// 1. Marginals of the presence/absence table (:352-355 ``gene_mat.sum(axis=1)`` -> collections.Counter;
//    sparse_utils.py:284-292 LightSparseDataFrame.sum; core_genome.py:127-155 count_gene_occurence):
//    genes per genome, genomes per gene, and the gene-frequency spectrum {frequency: number of genes} together
//    with the first gene of every frequency (a Counter is ordered by first appearance, and :364 slices that order).
//    8 bytes per COO entry, counted with warp-aggregated reductions; the host-buffer call is bound by PCIe.
//
// 2. The Monte-Carlo Kolmogorov-Smirnov statistics of ks_montecarlo_bbn (:457-482): ``iterations`` simulated samples
//    of ``n_samples`` draws each from the fitted beta-binomial (draw_bbn :484-492 = numpy legacy
//    RandomState.choice(p=probs)), their eCDFs (ecdf_from_counts :494-499) and ks_sim[i] = max |eCDF_i - model CDF|.
//    The reference spends one np.unique + a Python loop per iteration; here an iteration is one CTA (or a few):
//    a draw is two raw MT19937 words -> u = ((a >> 5) * 2^26 + (b >> 6)) / 2^53 -> upper bound in the choice CDF
//    (shared memory) -> shared-memory histogram; the eCDF is an integer prefix sum divided once per bin, so every
//    ks_sim[i] equals the reference's float64 bit for bit.  The raw words come from the host's bit-exact MT19937
//    stream (pgx_rng.cpp; the stream is serial), 8 bytes per draw: HBM-bound on the device, bound by the generator
//    (and PCIe) in the host-buffer call.
#include <limits.h>
#include <string.h>

#include <mutex>
#include <thread>
#include <vector>

#include "pgx_common.cuh"

extern "C" int pgx_legacy_random_raw(uint32_t *mt_key, int32_t *mt_pos, int64_t count, uint32_t *h_out);

namespace pgx {

namespace {

constexpr int KS_THREADS = 256;
constexpr int KS_UNROLL = 4;
constexpr int COUNT_THREADS = 256;

// ---------------------------------------------------------------------------------------------------------
// marginals
// ---------------------------------------------------------------------------------------------------------
// One COO entry per lane and step (a warp reads 128 contiguous bytes of ``row`` and of ``col``); lanes holding the
// same gene (or genome) elect a leader that adds their number with ONE reduction: tables arrive sorted by gene
// or by genome, so an unaggregated warp would send 32 reductions to one address.  (A version in which every thread
// run-length-counts a contiguous stretch of 32 entries was slower on config C4, 1.02 against 0.66 ms: the stretches of
// a CTA's threads do not fit L1 together, and the scattered axis pays one reduction per entry either way.)
__device__ __forceinline__ void aggregated_add(int32_t *bins, int key, bool valid)
{
    const unsigned active = __ballot_sync(FULL_MASK, valid);
    if (!valid) return;
    const unsigned same = __match_any_sync(active, key);
    if ((__ffs(same) - 1) == static_cast<int>(threadIdx.x & 31)) atomicAdd(bins + key, __popc(same));
}

__global__ void __launch_bounds__(COUNT_THREADS)
coo_count_kernel(const int32_t *__restrict__ row, const int32_t *__restrict__ col, long long nnz, int n_genes,
                 int n_genomes, int32_t *__restrict__ row_sum, int32_t *__restrict__ col_sum, int32_t *__restrict__ bad)
{
    const long long stride = static_cast<long long>(gridDim.x) * COUNT_THREADS;
    const long long rounds = (nnz + stride - 1) / stride;                      // same trip count for the whole warp
    long long i = static_cast<long long>(blockIdx.x) * COUNT_THREADS + threadIdx.x;
    int n_bad = 0;
    for (long long r = 0; r < rounds; ++r, i += stride) {
        int g = -1, c = -1;
        if (i < nnz) {
            g = __ldg(row + i);
            c = __ldg(col + i);
            if (g < 0 || g >= n_genes || c < 0 || c >= n_genomes) {
                g = c = -1;
                ++n_bad;
            }
        }
        aggregated_add(row_sum, g, g >= 0);
        aggregated_add(col_sum, c, c >= 0);
    }
    if (n_bad) atomicAdd(bad, n_bad);
}

__global__ void __launch_bounds__(COUNT_THREADS)
spectrum_kernel(const int32_t *__restrict__ row_sum, long long n_genes, int n_genomes,
                unsigned long long *__restrict__ spectrum, int32_t *__restrict__ first_gene)
{
    const long long stride = static_cast<long long>(gridDim.x) * COUNT_THREADS;
    const long long rounds = (n_genes + stride - 1) / stride;
    long long g = static_cast<long long>(blockIdx.x) * COUNT_THREADS + threadIdx.x;
    for (long long r = 0; r < rounds; ++r, g += stride) {
        const bool valid = g < n_genes;
        int m = valid ? __ldg(row_sum + g) : 0;
        m = min(max(m, 0), n_genomes);                                         // duplicates could exceed N: clamp
        const unsigned active = __ballot_sync(FULL_MASK, valid);
        if (!valid) continue;
        const unsigned same = __match_any_sync(active, m);
        if ((__ffs(same) - 1) == static_cast<int>(threadIdx.x & 31)) {        // the lowest lane holds the lowest gene
            atomicAdd(spectrum + m, static_cast<unsigned long long>(__popc(same)));
            atomicMin(first_gene + m, static_cast<int>(g));
        }
    }
}

__global__ void fill_i32_kernel(int32_t *__restrict__ p, long long n, int32_t value)
{
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) p[i] = value;
}

// ---------------------------------------------------------------------------------------------------------
// Monte-Carlo KS
// ---------------------------------------------------------------------------------------------------------
// numpy: cdf.searchsorted(u, side='right') = number of entries <= u
__device__ __forceinline__ int upper_bound(const double *cdf, int len, double u)
{
    int lo = 0, n = len;
    while (n > 0) {
        const int half = n >> 1;
        if (cdf[lo + half] <= u) {
            lo += half + 1;
            n -= half + 1;
        } else {
            n = half;
        }
    }
    return lo;
}

// grid = (splits, iterations).  CTA (s, it) draws samples [s * per_split, ...) of iteration ``it``.  With
// ``in_smem`` the choice CDF and the histogram live in shared memory (sim_limit * 12 bytes); otherwise the CDF is
// read through L1/L2 and the histogram is the iteration's row of ``g_hist`` (zeroed by the caller).  When an
// iteration is split, every CTA adds its bins to ``g_hist`` and the last one to arrive (ticket) finishes it.
template <bool IN_SMEM>
__global__ void __launch_bounds__(KS_THREADS)
ks_kernel(const uint2 *__restrict__ raw, long long n_samples, long long per_split, const double *__restrict__ choice_cdf,
          const double *__restrict__ model_cdf, int sim_limit, int32_t *__restrict__ g_hist,
          int32_t *__restrict__ tickets, double *__restrict__ ks_sim)
{
    constexpr bool in_smem = IN_SMEM;          // a template parameter: the shared-memory instance gets LDS / ATOMS, not generic accesses
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ long long seg_sum[KS_THREADS];
    __shared__ double warp_max[KS_THREADS / 32];
    __shared__ int is_last;

    const int it = blockIdx.y, splits = gridDim.x, tid = threadIdx.x, lane = tid & 31;
    double *s_cdf = reinterpret_cast<double *>(smem_raw);
    int32_t *s_hist = reinterpret_cast<int32_t *>(s_cdf + sim_limit);
    int32_t *row_hist = g_hist ? g_hist + static_cast<long long>(it) * sim_limit : nullptr;
    if constexpr (in_smem) {
        for (int k = tid; k < sim_limit; k += KS_THREADS) {
            s_cdf[k] = __ldg(choice_cdf + k);
            s_hist[k] = 0;
        }
        __syncthreads();
    }

    // ---- draws ------------------------------------------------------------------------------------------
    const long long s0 = static_cast<long long>(blockIdx.x) * per_split;
    const long long s1 = min(n_samples, s0 + per_split);
    const uint2 *src = raw + static_cast<long long>(it) * n_samples;
    // KS_UNROLL draws per thread and round: their loads are issued together (one 8-byte load in flight per thread left
    // the kernel waiting on the long scoreboard, ncu capture r02q)
    const long long rounds = (max(s1 - s0, 0ll) + KS_THREADS * KS_UNROLL - 1) / (KS_THREADS * KS_UNROLL);
    long long s = s0 + tid;
    for (long long r = 0; r < rounds; ++r, s += KS_THREADS * KS_UNROLL) {
        uint2 w[KS_UNROLL];
#pragma unroll
        for (int j = 0; j < KS_UNROLL; ++j)
            w[j] = s + j * KS_THREADS < s1 ? __ldcs(src + s + j * KS_THREADS) : make_uint2(0u, 0u);      // read once: streaming
#pragma unroll
        for (int j = 0; j < KS_UNROLL; ++j) {
            const bool valid = s + j * KS_THREADS < s1;
            int idx = 0;
            if (valid) {
                const double u = (static_cast<double>(w[j].x >> 5) * 67108864.0 + static_cast<double>(w[j].y >> 6)) *
                                 (1.0 / 9007199254740992.0);                   // exact: a power of two
                if constexpr (in_smem) idx = min(upper_bound(s_cdf, sim_limit, u), sim_limit - 1);
                else idx = min(upper_bound(choice_cdf, sim_limit, u), sim_limit - 1);
            }
            // most draws fall into the first few bins (a core gene is rarely missing): aggregate equal bins per warp
            const unsigned active = __ballot_sync(FULL_MASK, valid);
            if (valid) {
                const unsigned same = __match_any_sync(active, idx);
                if ((__ffs(same) - 1) == lane) {
                    if constexpr (in_smem) atomicAdd(s_hist + idx, __popc(same));
                    else atomicAdd(row_hist + idx, __popc(same));
                }
            }
        }
    }
    __syncthreads();

    // ---- hand-over when the iteration is split ------------------------------------------------------------
    if (splits > 1) {
        if constexpr (in_smem)
            for (int k = tid; k < sim_limit; k += KS_THREADS)
                if (s_hist[k]) atomicAdd(row_hist + k, s_hist[k]);
        __threadfence();
        __syncthreads();
        if (tid == 0) is_last = atomicAdd(tickets + it, 1) == splits - 1;
        __syncthreads();
        if (!is_last) return;
        __threadfence();
        if constexpr (in_smem) {
            for (int k = tid; k < sim_limit; k += KS_THREADS) s_hist[k] = __ldcg(row_hist + k);
            __syncthreads();
        }
    }
    auto bin = [&](int k) -> long long {
        if constexpr (in_smem) return s_hist[k];
        else return __ldcg(row_hist + k);
    };

    // ---- eCDF and KS statistic (:475-479): cumsum(pmf) / pmf.sum(), max |. - model_cdf| ------------------------
    const int seg = (sim_limit + KS_THREADS - 1) / KS_THREADS;
    const int k0 = min(tid * seg, sim_limit), k1 = min(k0 + seg, sim_limit);
    long long local = 0;
    for (int k = k0; k < k1; ++k) local += bin(k);
    seg_sum[tid] = local;
    __syncthreads();
    long long before = 0;
    for (int t = 0; t < tid; ++t) before += seg_sum[t];                        // 256 adds: not worth a tree
    const double total = static_cast<double>(n_samples);
    double worst = 0.0;
    bool nan_seen = false;
    for (int k = k0; k < k1; ++k) {
        before += bin(k);
        const double d = fabs(static_cast<double>(before) / total - __ldg(model_cdf + k));
        nan_seen |= d != d;
        worst = fmax(worst, d);
    }
    if (nan_seen) worst = __longlong_as_double(0x7ff8000000000000ll);          // np.max propagates NaN
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double other = __shfl_xor_sync(FULL_MASK, worst, off);
        worst = (worst != worst || other != other) ? __longlong_as_double(0x7ff8000000000000ll) : fmax(worst, other);
    }
    if (lane == 0) warp_max[tid >> 5] = worst;
    __syncthreads();
    if (tid == 0) {
        double m = warp_max[0];
        for (int w = 1; w < KS_THREADS / 32; ++w)
            m = (m != m || warp_max[w] != warp_max[w]) ? __longlong_as_double(0x7ff8000000000000ll) : fmax(m, warp_max[w]);
        ks_sim[it] = m;
    }
}

struct KsShape {
    int splits;
    long long per_split;
    int in_smem;
    size_t smem_bytes;
};

int ks_shape(int64_t iterations, int64_t n_samples, int32_t sim_limit, KsShape *out)
{
    int dev = 0, sms = 0, smem_optin = 0;
    PGX_CUDA(cudaGetDevice(&dev));
    PGX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    PGX_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const size_t want = static_cast<size_t>(sim_limit) * 12;
    out->in_smem = want + 4096 <= static_cast<size_t>(smem_optin);
    out->smem_bytes = out->in_smem ? want : 0;
    // enough CTAs for four per SM, but no CTA with fewer than 8,192 draws
    const long long by_grid = (4ll * sms + iterations - 1) / iterations;
    const long long by_work = std::max<long long>(1, n_samples / 8192);
    out->splits = static_cast<int>(std::max<long long>(1, std::min<long long>(std::min(by_grid, by_work), 65535)));
    out->per_split = (n_samples + out->splits - 1) / out->splits;
    return PGX_OK;
}

int ks_launch(const uint32_t *d_raw, int64_t iterations, int64_t n_samples, const double *d_choice_cdf,
              const double *d_model_cdf, int32_t sim_limit, double *d_ks_sim, void *d_scratch, cudaStream_t st)
{
    KsShape shape;
    if (int rc = ks_shape(iterations, n_samples, sim_limit, &shape)) return rc;
    int32_t *g_hist = nullptr, *tickets = nullptr;
    if (shape.splits > 1 || !shape.in_smem) {
        if (!d_scratch) return fail(PGX_ERR_INVALID, "pgx_ks_montecarlo needs its scratch buffer for this shape");
        tickets = static_cast<int32_t *>(d_scratch);
        g_hist = tickets + iterations;
        PGX_CUDA(cudaMemsetAsync(d_scratch, 0, sizeof(int32_t) * static_cast<size_t>(iterations) * (1 + sim_limit), st));
    }
    if (shape.smem_bytes > 48 * 1024) {
        static std::mutex mu;
        std::lock_guard<std::mutex> lock(mu);
        PGX_CUDA(cudaFuncSetAttribute(ks_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(shape.smem_bytes)));
    }
    const dim3 grid(static_cast<unsigned>(shape.splits), static_cast<unsigned>(iterations));
    if (shape.in_smem)
        ks_kernel<true><<<grid, KS_THREADS, shape.smem_bytes, st>>>(reinterpret_cast<const uint2 *>(d_raw), n_samples, shape.per_split,
                                                                    d_choice_cdf, d_model_cdf, sim_limit, g_hist, tickets, d_ks_sim);
    else
        ks_kernel<false><<<grid, KS_THREADS, 0, st>>>(reinterpret_cast<const uint2 *>(d_raw), n_samples, shape.per_split,
                                                      d_choice_cdf, d_model_cdf, sim_limit, g_hist, tickets, d_ks_sim);
    PGX_LAUNCH_CHECK("ks_kernel");
    return PGX_OK;
}

int check_ks_args(int64_t iterations, int64_t n_samples, int32_t sim_limit)
{
    if (iterations < 0 || n_samples < 1 || sim_limit < 1)
        return fail(PGX_ERR_INVALID, "pgx_ks_montecarlo: iterations >= 0, n_samples >= 1 and sim_limit >= 1 are required");
    if (iterations > 65535) return fail(PGX_ERR_UNSUPPORTED, "pgx_ks_montecarlo: more than 65,535 iterations per call");
    if (n_samples > (1ll << 31) - 1) return fail(PGX_ERR_UNSUPPORTED, "pgx_ks_montecarlo: more than 2^31 - 1 draws per iteration");
    return PGX_OK;
}

// staging of the host-buffer call, kept for the life of the library (a fit calls it once per num_points)
struct KsStage {
    int device = -1;
    size_t raw_bytes = 0, cdf_len = 0, sim_len = 0, scratch_bytes = 0;
    uint32_t *h_raw[2] = {nullptr, nullptr};
    uint32_t *d_raw[2] = {nullptr, nullptr};
    double *d_cdf = nullptr, *d_ks = nullptr;
    void *d_scratch = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t done[2] = {nullptr, nullptr};
};

void ks_release(KsStage &s)
{
    for (int i = 0; i < 2; ++i) {
        if (s.h_raw[i]) cudaFreeHost(s.h_raw[i]);
        if (s.d_raw[i]) cudaFree(s.d_raw[i]);
        if (s.done[i]) cudaEventDestroy(s.done[i]);
    }
    if (s.d_cdf) cudaFree(s.d_cdf);
    if (s.d_ks) cudaFree(s.d_ks);
    if (s.d_scratch) cudaFree(s.d_scratch);
    if (s.stream) cudaStreamDestroy(s.stream);
    s = KsStage();
}

int ks_acquire(KsStage &s, int device, size_t raw_bytes, size_t cdf_len, size_t sim_len, size_t scratch_bytes)
{
    if (s.device == device && s.raw_bytes >= raw_bytes && s.cdf_len >= cdf_len && s.sim_len >= sim_len &&
        s.scratch_bytes >= scratch_bytes)
        return PGX_OK;
    // generous floors: a fit calls this with a few iterations first and with thousands later, and re-pinning the
    // staging costs more than the whole simulation (16 ms against 2 ms, measured)
    raw_bytes = std::max<size_t>(std::max(raw_bytes, s.device == device ? s.raw_bytes : 0), 32u << 20);
    cdf_len = std::max<size_t>(std::max(cdf_len, s.device == device ? s.cdf_len : 0), 4096);
    sim_len = std::max<size_t>(std::max(sim_len, s.device == device ? s.sim_len : 0), 65536);
    scratch_bytes = std::max<size_t>(std::max(scratch_bytes, s.device == device ? s.scratch_bytes : 0), 8u << 20);
    ks_release(s);
    PGX_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        PGX_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&s.h_raw[i]), raw_bytes, cudaHostAllocDefault));
        PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&s.d_raw[i]), raw_bytes));
        PGX_CUDA(cudaEventCreateWithFlags(&s.done[i], cudaEventDisableTiming));
    }
    PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&s.d_cdf), sizeof(double) * 2 * cdf_len));
    PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&s.d_ks), sizeof(double) * sim_len));
    PGX_CUDA(cudaMalloc(&s.d_scratch, std::max<size_t>(scratch_bytes, 16)));
    s.device = device;
    s.raw_bytes = raw_bytes;
    s.cdf_len = cdf_len;
    s.sim_len = sim_len;
    s.scratch_bytes = scratch_bytes;
    return PGX_OK;
}

// staging of pgx_coo_marginals_host, kept for the life of the library (64 MB pinned + 64 MB on the device in all)
constexpr int MARGINAL_LANES = 4;
constexpr int64_t MARGINAL_CHUNK = 1ll << 20;                  // entries per chunk: 2 x 4 MB

struct MarginalLane {
    int device = -1;
    cudaStream_t stream = nullptr;
    int32_t *h_row[2] = {nullptr, nullptr}, *h_col[2] = {nullptr, nullptr};
    int32_t *d_row[2] = {nullptr, nullptr}, *d_col[2] = {nullptr, nullptr};
    cudaEvent_t sent[2] = {nullptr, nullptr};
};
MarginalLane g_marginal_lanes[MARGINAL_LANES];

int marginal_lanes_acquire(int device, int n_lanes)
{
    for (int i = 0; i < n_lanes; ++i) {
        MarginalLane &ln = g_marginal_lanes[i];
        if (ln.device == device) continue;
        if (ln.device >= 0) {                                  // another device before: start over
            for (int s = 0; s < 2; ++s) {
                cudaFreeHost(ln.h_row[s]);
                cudaFreeHost(ln.h_col[s]);
                cudaFree(ln.d_row[s]);
                cudaFree(ln.d_col[s]);
                cudaEventDestroy(ln.sent[s]);
            }
            cudaStreamDestroy(ln.stream);
            ln = MarginalLane();
        }
        PGX_CUDA(cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking));
        const size_t bytes = sizeof(int32_t) * static_cast<size_t>(MARGINAL_CHUNK);
        for (int s = 0; s < 2; ++s) {
            PGX_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&ln.h_row[s]), bytes, cudaHostAllocDefault));
            PGX_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&ln.h_col[s]), bytes, cudaHostAllocDefault));
            PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&ln.d_row[s]), bytes));
            PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&ln.d_col[s]), bytes));
            PGX_CUDA(cudaEventCreateWithFlags(&ln.sent[s], cudaEventDisableTiming));
        }
        ln.device = device;
    }
    return PGX_OK;
}

}  // namespace

}  // namespace pgx

extern "C" {

int pgx_coo_marginals(const int32_t *d_row, const int32_t *d_col, int64_t nnz, int32_t n_genes, int32_t n_genomes,
                      int32_t *d_row_sum, int32_t *d_col_sum, int32_t *d_bad, int32_t accumulate, void *stream)
{
    if (nnz < 0 || n_genes < 0 || n_genomes < 0) return pgx::fail(PGX_ERR_INVALID, "bad shape passed to pgx_coo_marginals");
    if (!d_row_sum || !d_col_sum || !d_bad || (nnz > 0 && (!d_row || !d_col)))
        return pgx::fail(PGX_ERR_INVALID, "null pointer passed to pgx_coo_marginals");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!accumulate) {
        PGX_CUDA(cudaMemsetAsync(d_row_sum, 0, sizeof(int32_t) * static_cast<size_t>(n_genes), st));
        PGX_CUDA(cudaMemsetAsync(d_col_sum, 0, sizeof(int32_t) * static_cast<size_t>(n_genomes), st));
        PGX_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int32_t), st));
    }
    if (nnz == 0) return PGX_OK;
    int dev = 0, sms = 0;
    PGX_CUDA(cudaGetDevice(&dev));
    PGX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long want = (nnz + pgx::COUNT_THREADS - 1) / pgx::COUNT_THREADS;
    const unsigned grid = static_cast<unsigned>(std::min<long long>(want, 8ll * sms));
    pgx::coo_count_kernel<<<grid, pgx::COUNT_THREADS, 0, st>>>(d_row, d_col, nnz, n_genes, n_genomes, d_row_sum, d_col_sum, d_bad);
    PGX_LAUNCH_CHECK("coo_count_kernel");
    return PGX_OK;
}

int pgx_frequency_spectrum(const int32_t *d_row_sum, int64_t n_genes, int32_t n_genomes, int64_t *d_spectrum,
                           int32_t *d_first_gene, void *stream)
{
    if (n_genes < 0 || n_genes > INT_MAX || n_genomes < 0) return pgx::fail(PGX_ERR_INVALID, "bad shape passed to pgx_frequency_spectrum");
    if (!d_spectrum || !d_first_gene || (n_genes > 0 && !d_row_sum))
        return pgx::fail(PGX_ERR_INVALID, "null pointer passed to pgx_frequency_spectrum");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long bins = static_cast<long long>(n_genomes) + 1;
    PGX_CUDA(cudaMemsetAsync(d_spectrum, 0, sizeof(int64_t) * static_cast<size_t>(bins), st));
    pgx::fill_i32_kernel<<<static_cast<unsigned>((bins + 255) / 256), 256, 0, st>>>(d_first_gene, bins, INT_MAX);
    PGX_LAUNCH_CHECK("fill_i32_kernel");
    if (n_genes == 0) return PGX_OK;
    int dev = 0, sms = 0;
    PGX_CUDA(cudaGetDevice(&dev));
    PGX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long want = (n_genes + pgx::COUNT_THREADS - 1) / pgx::COUNT_THREADS;
    const unsigned grid = static_cast<unsigned>(std::min<long long>(want, 8ll * sms));
    pgx::spectrum_kernel<<<grid, pgx::COUNT_THREADS, 0, st>>>(d_row_sum, n_genes, n_genomes,
                                                             reinterpret_cast<unsigned long long *>(d_spectrum), d_first_gene);
    PGX_LAUNCH_CHECK("spectrum_kernel");
    return PGX_OK;
}

int pgx_coo_marginals_host(const int32_t *h_row, const int32_t *h_col, int64_t nnz, int32_t n_genes, int32_t n_genomes,
                           int32_t *h_row_sum, int32_t *h_col_sum, int64_t *h_spectrum, int32_t *h_first_gene)
{
    if (nnz < 0 || n_genes < 0 || n_genomes < 0) return pgx::fail(PGX_ERR_INVALID, "bad shape passed to pgx_coo_marginals_host");
    if (!h_row_sum || !h_col_sum || (nnz > 0 && (!h_row || !h_col)))
        return pgx::fail(PGX_ERR_INVALID, "null pointer passed to pgx_coo_marginals_host");
    static std::mutex mu;                                      // one call at a time: the lanes' staging is shared
    std::lock_guard<std::mutex> lock(mu);
    int dev = 0;
    PGX_CUDA(cudaGetDevice(&dev));
    const size_t bins = static_cast<size_t>(n_genomes) + 1;
    int32_t *d_row_sum = nullptr, *d_col_sum = nullptr, *d_bad = nullptr, *d_first = nullptr;
    int64_t *d_spectrum = nullptr;
    cudaStream_t st = nullptr;
    int bad = 0;
    auto body = [&]() -> int {
        PGX_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&d_row_sum), sizeof(int32_t) * std::max<size_t>(n_genes, 1)));
        PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&d_col_sum), sizeof(int32_t) * std::max<size_t>(n_genomes, 1)));
        PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&d_bad), sizeof(int32_t)));
        PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&d_spectrum), sizeof(int64_t) * bins));
        PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&d_first), sizeof(int32_t) * bins));
        if (int r = pgx_coo_marginals(nullptr, nullptr, 0, n_genes, n_genomes, d_row_sum, d_col_sum, d_bad, 0, st)) return r;
        PGX_CUDA(cudaStreamSynchronize(st));                   // the lanes add into zeroed counters
        // The table lives in pageable memory.  A few lanes -- host thread + stream + two pinned / device chunk pairs
        // each -- copy their chunks into pinned staging, send them over and count them; the counters are shared
        // (reductions), so the lanes need no order among themselves.
        const int n_lanes = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(pgx::MARGINAL_LANES, (nnz + pgx::MARGINAL_CHUNK - 1) / pgx::MARGINAL_CHUNK)));
        if (int r = pgx::marginal_lanes_acquire(dev, n_lanes)) return r;
        std::atomic<int64_t> next_chunk{0};
        std::atomic<int> failed{0};
        char lane_error[pgx::MARGINAL_LANES][512];
        auto lane_body = [&](int lane) {
            pgx::MarginalLane &ln = pgx::g_marginal_lanes[lane];
            auto run = [&]() -> int {
                PGX_CUDA(cudaSetDevice(dev));
                int64_t turn = 0;
                while (!failed.load(std::memory_order_relaxed)) {
                    const int64_t p0 = next_chunk.fetch_add(1) * pgx::MARGINAL_CHUNK;
                    if (p0 >= nnz) break;
                    const int64_t cnt = std::min<int64_t>(pgx::MARGINAL_CHUNK, nnz - p0);
                    const int slot = static_cast<int>(turn++ & 1);
                    if (turn > 2) PGX_CUDA(cudaEventSynchronize(ln.sent[slot]));          // staging free again
                    memcpy(ln.h_row[slot], h_row + p0, sizeof(int32_t) * static_cast<size_t>(cnt));
                    memcpy(ln.h_col[slot], h_col + p0, sizeof(int32_t) * static_cast<size_t>(cnt));
                    PGX_CUDA(cudaMemcpyAsync(ln.d_row[slot], ln.h_row[slot], sizeof(int32_t) * static_cast<size_t>(cnt), cudaMemcpyHostToDevice, ln.stream));
                    PGX_CUDA(cudaMemcpyAsync(ln.d_col[slot], ln.h_col[slot], sizeof(int32_t) * static_cast<size_t>(cnt), cudaMemcpyHostToDevice, ln.stream));
                    PGX_CUDA(cudaEventRecord(ln.sent[slot], ln.stream));
                    if (int r = pgx_coo_marginals(ln.d_row[slot], ln.d_col[slot], cnt, n_genes, n_genomes, d_row_sum, d_col_sum, d_bad, 1, ln.stream))
                        return r;
                }
                PGX_CUDA(cudaStreamSynchronize(ln.stream));
                return PGX_OK;
            };
            if (run() != PGX_OK) {
                snprintf(lane_error[lane], sizeof(lane_error[lane]), "%s", pgx_last_error());
                failed.store(lane + 1);
            }
        };
        std::vector<std::thread> threads;
        for (int lane = 1; lane < n_lanes; ++lane) threads.emplace_back(lane_body, lane);
        lane_body(0);
        for (auto &t : threads) t.join();
        if (int f = failed.load()) return pgx::fail(PGX_ERR_CUDA, "%s", lane_error[f - 1]);
        if (h_spectrum || h_first_gene)
            if (int r = pgx_frequency_spectrum(d_row_sum, n_genes, n_genomes, d_spectrum, d_first, st)) return r;
        PGX_CUDA(cudaMemcpyAsync(h_row_sum, d_row_sum, sizeof(int32_t) * static_cast<size_t>(n_genes), cudaMemcpyDeviceToHost, st));
        PGX_CUDA(cudaMemcpyAsync(h_col_sum, d_col_sum, sizeof(int32_t) * static_cast<size_t>(n_genomes), cudaMemcpyDeviceToHost, st));
        PGX_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        if (h_spectrum) PGX_CUDA(cudaMemcpyAsync(h_spectrum, d_spectrum, sizeof(int64_t) * bins, cudaMemcpyDeviceToHost, st));
        if (h_first_gene) PGX_CUDA(cudaMemcpyAsync(h_first_gene, d_first, sizeof(int32_t) * bins, cudaMemcpyDeviceToHost, st));
        PGX_CUDA(cudaStreamSynchronize(st));
        return PGX_OK;
    };
    const int rc = body();
    char text[512];
    snprintf(text, sizeof(text), "%s", pgx_last_error());
    if (st) cudaStreamSynchronize(st);
    if (d_row_sum) cudaFree(d_row_sum);
    if (d_col_sum) cudaFree(d_col_sum);
    if (d_bad) cudaFree(d_bad);
    if (d_spectrum) cudaFree(d_spectrum);
    if (d_first) cudaFree(d_first);
    if (st) cudaStreamDestroy(st);
    if (rc) return pgx::fail(rc, "%s", text);
    if (bad) return pgx::fail(PGX_ERR_INVALID, "%d COO entries lie outside the %d x %d table", bad, n_genes, n_genomes);
    return PGX_OK;
}

size_t pgx_ks_scratch_bytes(int64_t iterations, int32_t sim_limit)
{
    if (iterations < 0 || sim_limit < 0) return 0;
    return sizeof(int32_t) * static_cast<size_t>(iterations) * (static_cast<size_t>(sim_limit) + 1);
}

int pgx_ks_montecarlo(const uint32_t *d_raw, int64_t iterations, int64_t n_samples, const double *d_choice_cdf,
                      const double *d_model_cdf, int32_t sim_limit, double *d_ks_sim, void *d_scratch, void *stream)
{
    if (int rc = pgx::check_ks_args(iterations, n_samples, sim_limit)) return rc;
    if (iterations == 0) return PGX_OK;
    if (!d_raw || !d_choice_cdf || !d_model_cdf || !d_ks_sim) return pgx::fail(PGX_ERR_INVALID, "null pointer passed to pgx_ks_montecarlo");
    return pgx::ks_launch(d_raw, iterations, n_samples, d_choice_cdf, d_model_cdf, sim_limit, d_ks_sim, d_scratch,
                          static_cast<cudaStream_t>(stream));
}

int pgx_ks_montecarlo_host(uint32_t *mt_key, int32_t *mt_pos, int64_t iterations, int64_t n_samples,
                           const double *h_choice_cdf, const double *h_model_cdf, int32_t sim_limit, double *h_ks_sim)
{
    if (int rc = pgx::check_ks_args(iterations, n_samples, sim_limit)) return rc;
    if (!mt_key || !mt_pos || !h_choice_cdf || !h_model_cdf || (iterations > 0 && !h_ks_sim))
        return pgx::fail(PGX_ERR_INVALID, "null pointer passed to pgx_ks_montecarlo_host");
    if (iterations == 0) return PGX_OK;
    for (int k = 0; k < sim_limit; ++k)
        if (!(h_choice_cdf[k] >= 0.0) || (k > 0 && h_choice_cdf[k] < h_choice_cdf[k - 1]))
            return pgx::fail(PGX_ERR_INVALID, "pgx_ks_montecarlo_host: the choice CDF must be finite, non-negative and non-decreasing");
    if (!(h_choice_cdf[sim_limit - 1] >= 1.0))
        return pgx::fail(PGX_ERR_INVALID, "pgx_ks_montecarlo_host: the choice CDF must end at 1 (numpy divides it by its last entry)");
    const size_t per_iteration = sizeof(uint32_t) * 2 * static_cast<size_t>(n_samples);
    if (per_iteration > (1ull << 30)) return pgx::fail(PGX_ERR_UNSUPPORTED, "pgx_ks_montecarlo_host: more than 2^27 draws per iteration");
    // blocks of whole iterations, about 32 MB of raw words each: the generator fills one while the other is in flight
    const int64_t per_block = std::max<int64_t>(1, std::min<int64_t>(iterations, static_cast<int64_t>((32ull << 20) / per_iteration)));
    static pgx::KsStage stage;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    int dev = 0;
    PGX_CUDA(cudaGetDevice(&dev));
    if (int rc = pgx::ks_acquire(stage, dev, per_iteration * static_cast<size_t>(per_block), static_cast<size_t>(sim_limit),
                                 static_cast<size_t>(iterations), pgx_ks_scratch_bytes(per_block, sim_limit)))
        return rc;
    cudaStream_t st = stage.stream;
    PGX_CUDA(cudaMemcpyAsync(stage.d_cdf, h_choice_cdf, sizeof(double) * sim_limit, cudaMemcpyHostToDevice, st));
    PGX_CUDA(cudaMemcpyAsync(stage.d_cdf + stage.cdf_len, h_model_cdf, sizeof(double) * sim_limit, cudaMemcpyHostToDevice, st));
    PGX_CUDA(cudaStreamSynchronize(st));                                       // the cdf arrays are pageable: copied now
    int rc = PGX_OK;
    int64_t block = 0;
    for (int64_t i0 = 0; i0 < iterations && !rc; i0 += per_block, ++block) {
        const int slot = static_cast<int>(block & 1);
        const int64_t cnt = std::min(per_block, iterations - i0);
        if (block >= 2) PGX_CUDA(cudaEventSynchronize(stage.done[slot]));
        rc = pgx_legacy_random_raw(mt_key, mt_pos, 2 * n_samples * cnt, stage.h_raw[slot]);
        if (rc) break;
        PGX_CUDA(cudaMemcpyAsync(stage.d_raw[slot], stage.h_raw[slot], per_iteration * static_cast<size_t>(cnt), cudaMemcpyHostToDevice, st));
        rc = pgx::ks_launch(stage.d_raw[slot], cnt, n_samples, stage.d_cdf, stage.d_cdf + stage.cdf_len, sim_limit,
                            stage.d_ks + i0, stage.d_scratch, st);
        if (!rc) PGX_CUDA(cudaEventRecord(stage.done[slot], st));
    }
    if (rc) {
        char text[512];
        snprintf(text, sizeof(text), "%s", pgx_last_error());
        cudaStreamSynchronize(st);
        return pgx::fail(rc, "%s", text);
    }
    PGX_CUDA(cudaMemcpyAsync(h_ks_sim, stage.d_ks, sizeof(double) * static_cast<size_t>(iterations), cudaMemcpyDeviceToHost, st));
    PGX_CUDA(cudaStreamSynchronize(st));
    return PGX_OK;
}

}  // extern "C"

// Pan/core rarefaction on sm_100a.
//
// Replaces the double loop of estimate_pan_core_size
// (/root/reference/pangenomix/pangenome_analysis.py:81-90).  For one genome order pi with
// rank = pi^-1, a gene g enters the pan-genome at step fp_g = min rank of its present
// genomes and leaves the core genome at step fa_g = min rank of its absent genomes, so
//   pan[k]  = #{g : fp_g <= k} = cumsum(hist(fp))[k]
//   core[k] = #{g : fa_g >  k} = G - cumsum(hist(fa))[k]
// (bit-exact with :89-:90; the identity itself is under test in tests/test_oracle_golden.py).
//
// Kernel 1 (minrank_kernel<B>): a CTA stages the rank tables of B permutations in shared
//   memory as T[genome][B] uint16 -- one 16-byte line per genome for B = 8 -- streams the
//   folded index chunks of its share of the rows with 128-bit no-allocate loads, and for
//   every index does ONE shared-memory gather that serves all B permutations, folding it
//   into packed uint16x2 running minima (VIMNMX.U16x2).  Rows are served by 1..32 lanes
//   according to their length class; the lanes of a row combine with a shuffle butterfly.
//   The streamed list yields one statistic directly (its min rank); the other one is the
//   mex of the streamed ranks, which is 0 unless the min is 0 -- then it is found by
//   probing genomes in rank order with a binary search of the sorted list (rare).
//   Results go straight into the output rows, used as per-permutation histograms.
// Kernel 2 (scan_kernel): adds the closed-form gene classes (empty / universal /
//   single-genome / single-absence genes are functions of perm[0] and perm[k] only) and
//   turns each histogram into its curve with a block-wide prefix scan, in place.
#include <mutex>
#include <vector>

#include "pgx_common.cuh"

namespace pgx {

namespace {

// Optional CUDA-event brackets around the two kernels of every call (bench.py's roofline).
struct ProfileEvents { cudaEvent_t begin, mid, end; };
bool g_profile_on = false;
std::mutex g_profile_mu;
std::vector<ProfileEvents> g_profile_events;

struct Tuning {
    int perms_per_cta = 0;
    int row_splits = 0;
    int threads = 0;
};
Tuning g_tuning;

template <int B>
struct Packed {
    static constexpr int REGS = (B + 1) / 2;
};

template <int B>
__device__ __forceinline__ uint32_t pmin(uint32_t x, uint32_t y)
{
    if constexpr (B == 1) return min(x, y);
    return __vminu2(x, y);
}

// One gather of the B ranks of genome c, folded into the running minima.
template <int B>
__device__ __forceinline__ void gather_min(const uint16_t *table, uint32_t c,
                                           uint32_t (&acc)[Packed<B>::REGS])
{
    if constexpr (B == 8) {
        const uint4 v = *reinterpret_cast<const uint4 *>(table + c * 8u);
        acc[0] = __vminu2(acc[0], v.x);
        acc[1] = __vminu2(acc[1], v.y);
        acc[2] = __vminu2(acc[2], v.z);
        acc[3] = __vminu2(acc[3], v.w);
    } else if constexpr (B == 4) {
        const uint2 v = *reinterpret_cast<const uint2 *>(table + c * 4u);
        acc[0] = __vminu2(acc[0], v.x);
        acc[1] = __vminu2(acc[1], v.y);
    } else if constexpr (B == 2) {
        const uint32_t v = *reinterpret_cast<const uint32_t *>(table + c * 2u);
        acc[0] = __vminu2(acc[0], v);
    } else {
        acc[0] = min(acc[0], static_cast<uint32_t>(table[c]));
    }
}

// mex of the ranks of a sorted genome list, given that rank 0 is in it: walk the genome
// order from rank 1 and stop at the first genome that is not in the list.
__device__ __noinline__ int mex_probe(const uint16_t *__restrict__ perm,
                                      const uint16_t *__restrict__ list, int len, int n)
{
    int k = 1;
    while (k < n) {
        const uint32_t c = perm[k];
        int lo = 0, hi = len;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (list[mid] < c) lo = mid + 1; else hi = mid;
        }
        if (lo < len && list[lo] == c) ++k; else break;
    }
    return k;
}

template <int B>
__global__ void __launch_bounds__(1024)
minrank_kernel(const pgx_plan plan, const uint16_t *__restrict__ perms, const long long n_perm,
               int32_t *__restrict__ hist)
{
    extern __shared__ __align__(16) uint16_t table[];   // [(N + 1)][B]
    __shared__ int s_next_task;

    constexpr int REGS = Packed<B>::REGS;
    const int n = plan.n_genomes;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const long long p0 = static_cast<long long>(blockIdx.y) * B;
    const int n_valid = static_cast<int>(min(static_cast<long long>(B), n_perm - p0));

    // ---- stage the inverse permutations: T[perm[k]][q] = k ----
    for (int q = 0; q < B; ++q) {
        if (q < n_valid) {
            const uint16_t *perm = perms + (p0 + q) * n;
            for (int k = tid; k < n; k += blockDim.x) table[static_cast<uint32_t>(perm[k]) * B + q] = static_cast<uint16_t>(k);
        } else {
            for (int k = tid; k < n; k += blockDim.x) table[k * B + q] = 0xffffu;
        }
    }
    if (tid < B) table[n * B + tid] = 0xffffu;   // padding index N never wins a min
    if (tid == 0) s_next_task = 0;
    __syncthreads();

    const uint4 *__restrict__ chunks = reinterpret_cast<const uint4 *>(plan.d_chunks);
    const int2 *__restrict__ tasks = reinterpret_cast<const int2 *>(plan.d_tasks);
    const long long row_stride = 2ll * n;

    uint32_t zero_pan[B], zero_core[B];   // warp-uniform counts of "other statistic == 0"
#pragma unroll
    for (int q = 0; q < B; ++q) { zero_pan[q] = 0; zero_core[q] = 0; }

    for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(&s_next_task, 1);
        t = __shfl_sync(FULL_MASK, t, 0);
        const long long task = static_cast<long long>(t) * gridDim.x + blockIdx.x;
        if (task >= plan.n_tasks) break;

        const int2 td = tasks[task];
        const int log_w = (td.y >> 1) & 7;
        const int absent_list = td.y & 1;
        const int n_rows = td.y >> 8;
        const int w = 1 << log_w;
        const int group = lane >> log_w;
        const int lig = lane & (w - 1);
        const bool valid = group < n_rows;
        const int row = td.x + (valid ? group : 0);
        const int c0 = plan.d_row_ptr[row];
        const int c1 = valid ? plan.d_row_ptr[row + 1] : c0;

        uint32_t acc[REGS];
#pragma unroll
        for (int i = 0; i < REGS; ++i) acc[i] = 0xffffffffu;

#pragma unroll 2
        for (int ch = c0 + lig; ch < c1; ch += w) {
            const uint4 v = ldg_stream(chunks + ch);
            gather_min<B>(table, v.x & 0xffffu, acc);
            gather_min<B>(table, v.x >> 16, acc);
            gather_min<B>(table, v.y & 0xffffu, acc);
            gather_min<B>(table, v.y >> 16, acc);
            gather_min<B>(table, v.z & 0xffffu, acc);
            gather_min<B>(table, v.z >> 16, acc);
            gather_min<B>(table, v.w & 0xffffu, acc);
            gather_min<B>(table, v.w >> 16, acc);
        }
        for (int off = w >> 1; off > 0; off >>= 1) {
#pragma unroll
            for (int i = 0; i < REGS; ++i)
                acc[i] = pmin<B>(acc[i], __shfl_xor_sync(FULL_MASK, acc[i], off));
        }

        // present list: min -> pan histogram, mex -> core histogram; absent list: swapped.
        int32_t *list_hist = hist + (absent_list ? n : 0);
        int32_t *other_hist = hist + (absent_list ? 0 : n);
#pragma unroll
        for (int q = 0; q < B; ++q) {
            if (q < n_valid) {
                uint32_t mn;
                if constexpr (B == 1) mn = acc[0] & 0xffffu;
                else mn = (acc[q >> 1] >> ((q & 1) * 16)) & 0xffffu;
                const bool mine = valid && (lig == (q & (w - 1)));
                if (mine) {
                    const long long base = (p0 + q) * row_stride;
                    atomicAdd(list_hist + base + mn, 1);
                    if (mn == 0) {
                        const int k = mex_probe(perms + (p0 + q) * n, plan.d_chunks + 8ll * c0,
                                                (c1 - c0) * 8, n);
                        if (k < n) atomicAdd(other_hist + base + k, 1);
                    }
                }
                const uint32_t nz = __popc(__ballot_sync(FULL_MASK, mine && mn != 0));
                if (absent_list) zero_pan[q] += nz; else zero_core[q] += nz;
            }
        }
    }

    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < B; ++q) {
            if (q < n_valid) {
                const long long base = (p0 + q) * row_stride;
                if (zero_pan[q]) atomicAdd(hist + base, static_cast<int>(zero_pan[q]));
                if (zero_core[q]) atomicAdd(hist + base + n, static_cast<int>(zero_core[q]));
            }
        }
    }
}

// Histogram -> curve, in place when OutT == int32_t and out == hist.
template <typename OutT>
__global__ void __launch_bounds__(256)
scan_kernel(const pgx_plan plan, const uint16_t *__restrict__ perms, const int32_t *hist, OutT *out)
{
    constexpr int ITEMS = 4;
    constexpr int THREADS = 256;
    __shared__ int warp_tot[THREADS / 32];

    const int n = plan.n_genomes;
    const long long p = blockIdx.x;
    const int side = blockIdx.y;   // 0 = pan, 1 = core
    const int32_t *h = hist + p * 2ll * n + static_cast<long long>(side) * n;
    OutT *o = out + p * 2ll * n + static_cast<long long>(side) * n;
    const uint16_t *perm = perms + p * n;
    // genes living in / missing from a single genome c: the list-side statistic is rank[c],
    // the other one is 1 if c comes first and 0 otherwise.
    const int32_t *w_list = side == 0 ? plan.d_w_present : plan.d_w_absent;
    const int32_t *w_other = side == 0 ? plan.d_w_absent : plan.d_w_present;
    const int first = perm[0];
    const int other_first = w_other[first];
    const int add0 = side == 0 ? plan.n_full + (plan.sum_w_absent - other_first)
                               : plan.n_empty + (plan.sum_w_present - other_first);
    const int add1 = other_first;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int carry = 0;
    for (int base = 0; base < n; base += THREADS * ITEMS) {
        int v[ITEMS];
        int run = 0;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const int k = base + tid * ITEMS + i;
            int x = 0;
            if (k < n) {
                x = h[k] + w_list[perm[k]];
                if (k == 0) x += add0;
                if (k == 1) x += add1;
            }
            run += x;
            v[i] = run;
        }
        int incl = run;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int y = __shfl_up_sync(FULL_MASK, incl, off);
            if (lane >= off) incl += y;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int i = 0; i < THREADS / 32; ++i) {
            const int wt = warp_tot[i];
            if (i < warp) before += wt;
            total += wt;
        }
        const int excl = carry + before + incl - run;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const int k = base + tid * ITEMS + i;
            if (k < n) {
                const int c = excl + v[i];
                o[k] = static_cast<OutT>(side == 0 ? c : plan.n_genes - c);
            }
        }
        carry += total;
        __syncthreads();
    }
}

int check_plan(const pgx_plan *plan)
{
    if (!plan) return fail(PGX_ERR_INVALID, "plan is null");
    if (plan->n_genomes < 1 || plan->n_genomes > 65535)
        return fail(PGX_ERR_UNSUPPORTED, "n_genomes = %d outside 1..65535 (uint16 genome indices)",
                    plan->n_genomes);
    if (plan->n_genes < 0 || plan->n_rows < 0 || plan->n_tasks < 0 || plan->n_chunks < 0)
        return fail(PGX_ERR_INVALID, "negative size in plan");
    if (!plan->d_w_present || !plan->d_w_absent)
        return fail(PGX_ERR_INVALID, "plan weight vectors are null");
    if (plan->n_tasks > 0 && (!plan->d_chunks || !plan->d_row_ptr || !plan->d_tasks))
        return fail(PGX_ERR_INVALID, "plan has tasks but null row arrays");
    if (reinterpret_cast<uintptr_t>(plan->d_chunks) & 15)
        return fail(PGX_ERR_INVALID, "d_chunks must be 16-byte aligned");
    return PGX_OK;
}

struct DeviceLimits {
    int device = -1;
    int sm_count = 0;
    int smem_optin = 0;
};

int device_limits(DeviceLimits *out)
{
    static thread_local DeviceLimits cached;
    int dev = 0;
    PGX_CUDA(cudaGetDevice(&dev));
    if (cached.device != dev) {
        PGX_CUDA(cudaDeviceGetAttribute(&cached.sm_count, cudaDevAttrMultiProcessorCount, dev));
        PGX_CUDA(cudaDeviceGetAttribute(&cached.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        cached.device = dev;
    }
    *out = cached;
    return PGX_OK;
}

template <int B>
int launch_minrank(const pgx_plan &plan, const uint16_t *d_perms, long long n_perm, int32_t *d_hist,
                   const DeviceLimits &lim, cudaStream_t stream)
{
    const size_t smem = static_cast<size_t>(plan.n_genomes + 1) * B * sizeof(uint16_t);
    PGX_CUDA(cudaFuncSetAttribute(minrank_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
    int threads = g_tuning.threads;
    if (threads <= 0) threads = smem > 100 * 1024 ? 1024 : (smem > 40 * 1024 ? 512 : 256);
    threads = max(32, min(1024, (threads / 32) * 32));
    const long long batches = (n_perm + B - 1) / B;
    // Row splits: enough CTAs for ~16 waves so the tail is small, but every CTA keeps at
    // least a few tasks per warp (the table build is amortised over them).
    int splits = g_tuning.row_splits;
    if (splits <= 0) {
        const long long resident = static_cast<long long>(lim.sm_count) *
                                   max(1, min(8, static_cast<int>((200 * 1024) / (smem + 1024))));
        const long long want = (16 * resident + batches - 1) / batches;
        const long long most = max(1ll, static_cast<long long>(plan.n_tasks) / ((threads / 32) * 4));
        splits = static_cast<int>(max(1ll, min(want, most)));
    }
    splits = max(1, min(splits, 65535));
    if (batches > 65535ll * 32768ll) return fail(PGX_ERR_UNSUPPORTED, "too many permutations in one call");
    // blockIdx.y is limited to 65535: fold the batch index when needed.
    for (long long b0 = 0; b0 < batches; b0 += 65535) {
        const long long nb = min(65535ll, batches - b0);
        dim3 grid(static_cast<unsigned>(splits), static_cast<unsigned>(nb));
        minrank_kernel<B><<<grid, threads, smem, stream>>>(plan, d_perms + b0 * B * plan.n_genomes,
                                                            n_perm - b0 * B,
                                                            d_hist + b0 * B * 2ll * plan.n_genomes);
        PGX_LAUNCH_CHECK("minrank_kernel");
    }
    return PGX_OK;
}

template <typename OutT>
int run_curves(const pgx_plan *plan, const uint16_t *d_perms, long long n_perm, int32_t *d_hist,
               OutT *d_out, cudaStream_t stream)
{
    if (int rc = check_plan(plan)) return rc;
    if (n_perm < 0) return fail(PGX_ERR_INVALID, "n_perm < 0");
    if (n_perm == 0) return PGX_OK;
    if (!d_perms || !d_hist || !d_out) return fail(PGX_ERR_INVALID, "null permutation / output pointer");
    DeviceLimits lim;
    if (int rc = device_limits(&lim)) return rc;
    const int n = plan->n_genomes;
    PGX_CUDA(cudaMemsetAsync(d_hist, 0, sizeof(int32_t) * 2ull * n * n_perm, stream));
    ProfileEvents ev{};
    const bool profile = g_profile_on;
    if (profile) {
        PGX_CUDA(cudaEventCreate(&ev.begin));
        PGX_CUDA(cudaEventCreate(&ev.mid));
        PGX_CUDA(cudaEventCreate(&ev.end));
        PGX_CUDA(cudaEventRecord(ev.begin, stream));
    }
    if (plan->n_tasks > 0) {
        const size_t per_perm = static_cast<size_t>(n + 1) * sizeof(uint16_t);
        const size_t budget = static_cast<size_t>(lim.smem_optin) - 64;
        int b = g_tuning.perms_per_cta;
        if (b != 1 && b != 2 && b != 4 && b != 8) b = 8;
        while (b > 1 && per_perm * b > budget) b >>= 1;
        if (per_perm * b > budget) return fail(PGX_ERR_UNSUPPORTED, "rank table does not fit shared memory");
        int rc;
        switch (b) {
            case 8: rc = launch_minrank<8>(*plan, d_perms, n_perm, d_hist, lim, stream); break;
            case 4: rc = launch_minrank<4>(*plan, d_perms, n_perm, d_hist, lim, stream); break;
            case 2: rc = launch_minrank<2>(*plan, d_perms, n_perm, d_hist, lim, stream); break;
            default: rc = launch_minrank<1>(*plan, d_perms, n_perm, d_hist, lim, stream); break;
        }
        if (rc) return rc;
    }
    if (profile) PGX_CUDA(cudaEventRecord(ev.mid, stream));
    for (long long p0 = 0; p0 < n_perm; p0 += 2147483647ll) {
        const long long np = min(2147483647ll, n_perm - p0);
        dim3 grid(static_cast<unsigned>(np), 2);
        scan_kernel<OutT><<<grid, 256, 0, stream>>>(*plan, d_perms + p0 * n, d_hist + p0 * 2ll * n,
                                                    d_out + p0 * 2ll * n);
        PGX_LAUNCH_CHECK("scan_kernel");
    }
    if (profile) {
        PGX_CUDA(cudaEventRecord(ev.end, stream));
        std::lock_guard<std::mutex> lock(g_profile_mu);
        g_profile_events.push_back(ev);
    }
    return PGX_OK;
}

}  // namespace

}  // namespace pgx

extern "C" {

int pgx_pan_core_curves(const pgx_plan *plan, const uint16_t *d_perms, int64_t n_perm,
                        int32_t *d_curves, void *stream)
{
    return pgx::run_curves<int32_t>(plan, d_perms, n_perm, d_curves, d_curves,
                                    static_cast<cudaStream_t>(stream));
}

int pgx_pan_core_curves_f64(const pgx_plan *plan, const uint16_t *d_perms, int64_t n_perm,
                            int32_t *d_hist, double *d_curves, void *stream)
{
    return pgx::run_curves<double>(plan, d_perms, n_perm, d_hist, d_curves,
                                   static_cast<cudaStream_t>(stream));
}

int pgx_pan_core_curves_host(const pgx_plan *plan, const uint16_t *h_perms, int64_t n_perm,
                             void *h_curves, int32_t out_f64, int64_t perms_per_block)
{
    if (int rc = pgx::check_plan(plan)) return rc;
    if (n_perm < 0) return pgx::fail(PGX_ERR_INVALID, "n_perm < 0");
    if (n_perm == 0) return PGX_OK;
    if (!h_perms || !h_curves) return pgx::fail(PGX_ERR_INVALID, "null host pointer");
    const long long n = plan->n_genomes;
    long long block = perms_per_block;
    if (block <= 0) {
        // ~64 MB of curves per block keeps both PCIe directions and the SMs busy at once.
        block = (64ll << 20) / (2 * n * (out_f64 ? 8 : 4));
        block = std::max(64ll, std::min(block, 1ll << 16));
        block = (block + 7) / 8 * 8;
    }
    block = std::min<long long>(block, n_perm);
    const size_t perm_bytes = sizeof(uint16_t) * n * block;
    const size_t hist_bytes = sizeof(int32_t) * 2 * n * block;
    const size_t out_bytes = out_f64 ? sizeof(double) * 2 * n * block : 0;

    cudaStream_t streams[2] = {nullptr, nullptr};
    uint16_t *d_perms[2] = {nullptr, nullptr};
    int32_t *d_hist[2] = {nullptr, nullptr};
    double *d_out[2] = {nullptr, nullptr};
    int rc = PGX_OK;
    // Scratch comes from the device's default memory pool so that repeated calls reuse it.
    {
        int dev = 0;
        cudaMemPool_t pool;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    auto cleanup = [&]() {
        for (int s = 0; s < 2; ++s) {
            if (!streams[s]) continue;
            if (d_perms[s]) cudaFreeAsync(d_perms[s], streams[s]);
            if (d_hist[s]) cudaFreeAsync(d_hist[s], streams[s]);
            if (d_out[s]) cudaFreeAsync(d_out[s], streams[s]);
            cudaStreamSynchronize(streams[s]);
            cudaStreamDestroy(streams[s]);
        }
    };
#define PGX_TRY(expr)                                                                        \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            rc = pgx::fail(PGX_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__));   \
            cleanup();                                                                       \
            return rc;                                                                       \
        }                                                                                    \
    } while (0)
    const int n_streams = n_perm > block ? 2 : 1;
    for (int s = 0; s < n_streams; ++s) {
        PGX_TRY(cudaStreamCreateWithFlags(&streams[s], cudaStreamNonBlocking));
        PGX_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_perms[s]), perm_bytes, streams[s]));
        PGX_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_hist[s]), hist_bytes, streams[s]));
        if (out_f64) PGX_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_out[s]), out_bytes, streams[s]));
    }
    int slot = 0;
    for (long long p0 = 0; p0 < n_perm; p0 += block, slot ^= (n_streams - 1)) {
        const long long np = std::min<long long>(block, n_perm - p0);
        cudaStream_t st = streams[slot];
        PGX_TRY(cudaMemcpyAsync(d_perms[slot], h_perms + p0 * n, sizeof(uint16_t) * n * np,
                                cudaMemcpyHostToDevice, st));
        if (out_f64) {
            rc = pgx::run_curves<double>(plan, d_perms[slot], np, d_hist[slot], d_out[slot], st);
            if (rc) { cleanup(); return rc; }
            PGX_TRY(cudaMemcpyAsync(static_cast<double *>(h_curves) + p0 * 2 * n, d_out[slot],
                                    sizeof(double) * 2 * n * np, cudaMemcpyDeviceToHost, st));
        } else {
            rc = pgx::run_curves<int32_t>(plan, d_perms[slot], np, d_hist[slot], d_hist[slot], st);
            if (rc) { cleanup(); return rc; }
            PGX_TRY(cudaMemcpyAsync(static_cast<int32_t *>(h_curves) + p0 * 2 * n, d_hist[slot],
                                    sizeof(int32_t) * 2 * n * np, cudaMemcpyDeviceToHost, st));
        }
    }
    for (int s = 0; s < n_streams; ++s) PGX_TRY(cudaStreamSynchronize(streams[s]));
#undef PGX_TRY
    cleanup();
    return PGX_OK;
}

int pgx_profile_enable(int32_t on)
{
    pgx::g_profile_on = on != 0;
    return PGX_OK;
}

int pgx_profile_read(double *minrank_ms, double *scan_ms, int64_t *calls)
{
    std::lock_guard<std::mutex> lock(pgx::g_profile_mu);
    double a = 0.0, b = 0.0;
    for (auto &ev : pgx::g_profile_events) {
        float t0 = 0.f, t1 = 0.f;
        PGX_CUDA(cudaEventSynchronize(ev.end));
        PGX_CUDA(cudaEventElapsedTime(&t0, ev.begin, ev.mid));
        PGX_CUDA(cudaEventElapsedTime(&t1, ev.mid, ev.end));
        a += t0;
        b += t1;
        cudaEventDestroy(ev.begin);
        cudaEventDestroy(ev.mid);
        cudaEventDestroy(ev.end);
    }
    if (minrank_ms) *minrank_ms = a;
    if (scan_ms) *scan_ms = b;
    if (calls) *calls = static_cast<int64_t>(pgx::g_profile_events.size());
    pgx::g_profile_events.clear();
    return PGX_OK;
}

int pgx_set_tuning(int32_t perms_per_cta, int32_t row_splits, int32_t threads_per_cta)
{
    pgx::g_tuning.perms_per_cta = perms_per_cta;
    pgx::g_tuning.row_splits = row_splits;
    pgx::g_tuning.threads = threads_per_cta;
    return PGX_OK;
}

}  // extern "C"

// Batched Heaps-law fits on sm_100a:  y = kappa * x^alpha,  x = 1 .. n_points.
//
// GPU counterpart of fit_heaps_by_iteration / __fit_heaps_single__
// (/root/reference/pangenomix/pangenome_analysis.py:24-48), which calls
// scipy.optimize.curve_fit (unbounded -> MINPACK Levenberg-Marquardt) once per row of the
// pan/core table with the start point alpha = 0.5, kappa = min(y).  Here every curve gets one
// CTA that runs Levenberg-Marquardt with Marquardt scaling and the analytic Jacobian
//   d/d alpha = kappa * x^alpha * ln x,   d/d kappa = x^alpha
// to full fp64 convergence; all reductions use a fixed tree, so results are reproducible bit
// for bit.  scipy stops at ftol = xtol = 1.49e-8 and lands within ~1e-6 relative of the least
// squares optimum; the two agree to better than 5e-6 relative (tests/test_gpu_parity.py).
// The drop-in fit_heaps_by_iteration itself keeps calling scipy on the host; this entry point
// is for fitting every iteration of a large table.
#include "pgx_common.cuh"

namespace pgx {

namespace {

constexpr int HEAPS_THREADS = 128;
constexpr int HEAPS_MAX_TRIALS = 400;

__global__ void log_table_kernel(double *__restrict__ lnx, long long n_points)
{
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n_points) lnx[i] = log(static_cast<double>(i + 1));
}

struct Sums {
    double aa, ak, kk, ga, gk, cost;
};

__device__ __forceinline__ double block_sum(double v, double *scratch)
{
    // fixed shuffle tree inside the warp, fixed order across the warps
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL_MASK, v, off);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < HEAPS_THREADS / 32; ++w) s += scratch[w];
    return s;
}

template <typename T>
__device__ Sums evaluate(const T *__restrict__ y, const double *__restrict__ lnx, long long n_points,
                         double alpha, double kappa, double *scratch)
{
    Sums s{0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (long long i = threadIdx.x; i < n_points; i += HEAPS_THREADS) {
        const double lx = lnx[i];
        const double xa = exp(alpha * lx);
        const double f = kappa * xa;
        const double r = f - static_cast<double>(y[i]);
        const double ja = f * lx;
        s.aa += ja * ja;
        s.ak += ja * xa;
        s.kk += xa * xa;
        s.ga += ja * r;
        s.gk += xa * r;
        s.cost += r * r;
    }
    s.aa = block_sum(s.aa, scratch);
    s.ak = block_sum(s.ak, scratch);
    s.kk = block_sum(s.kk, scratch);
    s.ga = block_sum(s.ga, scratch);
    s.gk = block_sum(s.gk, scratch);
    s.cost = block_sum(s.cost, scratch);
    return s;
}

template <typename T>
__global__ void __launch_bounds__(HEAPS_THREADS, 8)
heaps_kernel(const T *__restrict__ curves, long long n_points, long long stride,
             const double *__restrict__ lnx, double *__restrict__ fit, int32_t *__restrict__ info)
{
    __shared__ double scratch[HEAPS_THREADS / 32];
    const T *y = curves + static_cast<long long>(blockIdx.x) * stride;

    // start point of the reference: alpha = 0.5, kappa = min(y)  (:45)
    double mn = 1.0e300;
    for (long long i = threadIdx.x; i < n_points; i += HEAPS_THREADS) mn = fmin(mn, static_cast<double>(y[i]));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mn = fmin(mn, __shfl_xor_sync(FULL_MASK, mn, off));
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = mn;
    __syncthreads();
    mn = scratch[0];
#pragma unroll
    for (int w = 1; w < HEAPS_THREADS / 32; ++w) mn = fmin(mn, scratch[w]);

    double alpha = 0.5, kappa = mn;
    Sums cur = evaluate(y, lnx, n_points, alpha, kappa, scratch);
    double lambda = 1.0e-3;
    int trials = 0;
    bool converged = false;
    while (trials < HEAPS_MAX_TRIALS) {
        ++trials;
        // (J^T J + lambda diag(J^T J)) delta = -J^T r     (every thread solves the same 2 x 2 system)
        const double a11 = cur.aa * (1.0 + lambda), a22 = cur.kk * (1.0 + lambda), a12 = cur.ak;
        const double det = a11 * a22 - a12 * a12;
        if (!(fabs(det) > 0.0) || !isfinite(det)) break;
        const double da = (-cur.ga * a22 + cur.gk * a12) / det;
        const double dk = (-cur.gk * a11 + cur.ga * a12) / det;
        const double alpha_t = alpha + da, kappa_t = kappa + dk;
        const Sums trial = evaluate(y, lnx, n_points, alpha_t, kappa_t, scratch);
        const bool tiny_step = fabs(da) <= 1.0e-15 * (fabs(alpha) + 1.0e-300) && fabs(dk) <= 1.0e-15 * (fabs(kappa) + 1.0e-300);
        if (isfinite(trial.cost) && trial.cost <= cur.cost) {
            const double gain = cur.cost - trial.cost;
            alpha = alpha_t;
            kappa = kappa_t;
            const double before = cur.cost;
            cur = trial;
            lambda = fmax(lambda * 0.1, 1.0e-12);
            if (tiny_step || gain <= 1.0e-16 * before) {
                converged = true;
                break;
            }
        } else {
            if (tiny_step || lambda > 1.0e16) {
                converged = tiny_step;
                break;
            }
            lambda *= 10.0;
        }
    }
    if (threadIdx.x == 0) {
        fit[2ll * blockIdx.x] = alpha;
        fit[2ll * blockIdx.x + 1] = kappa;
        if (info) info[blockIdx.x] = converged ? trials : -trials;
    }
}

}  // namespace

}  // namespace pgx

extern "C" {

size_t pgx_heaps_scratch_bytes(int64_t n_points)
{
    return n_points > 0 ? sizeof(double) * static_cast<size_t>(n_points) : 0;
}

int pgx_heaps_fit(const void *d_curves, int32_t is_f64, int64_t n_curves, int64_t n_points, int64_t stride,
                  double *d_fit, int32_t *d_info, void *d_scratch, void *stream)
{
    if (n_curves < 0 || n_points < 1 || stride < n_points)
        return pgx::fail(PGX_ERR_INVALID, "bad shape passed to pgx_heaps_fit");
    if (n_curves == 0) return PGX_OK;
    if (!d_curves || !d_fit || !d_scratch) return pgx::fail(PGX_ERR_INVALID, "null pointer passed to pgx_heaps_fit");
    if (n_curves > 2147483647ll) return pgx::fail(PGX_ERR_UNSUPPORTED, "too many curves in one call");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double *lnx = static_cast<double *>(d_scratch);
    pgx::log_table_kernel<<<static_cast<unsigned>((n_points + 255) / 256), 256, 0, st>>>(lnx, n_points);
    PGX_LAUNCH_CHECK("heaps log_table_kernel");
    if (is_f64)
        pgx::heaps_kernel<double><<<static_cast<unsigned>(n_curves), pgx::HEAPS_THREADS, 0, st>>>(
            static_cast<const double *>(d_curves), n_points, stride, lnx, d_fit, d_info);
    else
        pgx::heaps_kernel<int32_t><<<static_cast<unsigned>(n_curves), pgx::HEAPS_THREADS, 0, st>>>(
            static_cast<const int32_t *>(d_curves), n_points, stride, lnx, d_fit, d_info);
    PGX_LAUNCH_CHECK("heaps_kernel");
    return PGX_OK;
}

}  // extern "C"

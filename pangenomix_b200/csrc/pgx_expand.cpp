// Host half of the compact curve transfer of libpgx_b200 (plain C++, no CUDA).
//
// A pan/core table row (pangenome_analysis.py:89-90, one permutation) is two monotone curves whose steps are
// bounded by the gene count of one genome: pan[k] - pan[k-1] = genes first seen in genome perm[k] <= colsum,
// core[k-1] - core[k] = genes first missed in genome perm[k] <= core[k-1] <= colsum[perm[0]].  Whenever every
// genome of the table holds at most 65,535 genes the device therefore ships a row as 2N uint16 STEPS
//   d[0] = pan[0],  d[k] = pan[k] - pan[k-1];   d[N] = core[0],  d[N+k] = core[k-1] - core[k]
// -- a quarter of the bytes of the float64 row the reference builds (:76-77, :97) and half of an int32 row, over
// the PCIe link that bounds the host-buffer calls -- and the host threads that have to touch every element of the
// result anyway (it is fresh pageable memory) rebuild the curves with one prefix sum per half row.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#define PGX_X86 1
#else
#define PGX_X86 0
#endif

#include "pgx.h"

namespace pgx {
int fail(int code, const char *fmt, ...);

namespace {

template <typename OutT>
void expand_row_scalar(const uint16_t *d, long long n, OutT *out)
{
    int32_t run = 0;
    for (long long k = 0; k < n; ++k) {
        run += d[k];
        out[k] = static_cast<OutT>(run);
    }
    run = d[n];
    out[n] = static_cast<OutT>(run);
    for (long long k = 1; k < n; ++k) {
        run -= d[n + k];
        out[n + k] = static_cast<OutT>(run);
    }
}

// The SPLIT format of a step row (pgx_rarefy.cu, split_steps_kernel): large steps only occur while the first genomes
// are added, so the first ``head`` steps of each curve travel as uint16 and the rest as uint8:
//   [pan head: head x u16][core head: head x u16][pan tail: (N - head) x u8][core tail: (N - head) x u8]
// (2N + 2 head bytes per row instead of 4N).  1 <= head <= N.
template <typename InT, typename OutT>
int32_t prefix_scalar(const InT *d, long long count, int32_t run, bool subtract, OutT *out)
{
    for (long long k = 0; k < count; ++k) {
        run = subtract ? run - static_cast<int32_t>(d[k]) : run + static_cast<int32_t>(d[k]);
        out[k] = static_cast<OutT>(run);
    }
    return run;
}

template <typename OutT>
void expand_split_row_scalar(const uint8_t *row, long long n, long long head, OutT *out)
{
    const uint16_t *pan_head = reinterpret_cast<const uint16_t *>(row), *core_head = pan_head + head;
    const uint8_t *pan_tail = row + 4 * head, *core_tail = pan_tail + (n - head);
    int32_t run = prefix_scalar(pan_head, head, 0, false, out);
    prefix_scalar(pan_tail, n - head, run, false, out + head);
    run = core_head[0];
    out[n] = static_cast<OutT>(run);
    run = prefix_scalar(core_head + 1, head - 1, run, true, out + n + 1);
    prefix_scalar(core_tail, n - head, run, true, out + n + head);
}

#if PGX_X86
// inclusive prefix sum of 8 int32 lanes
__attribute__((target("avx2"))) inline __m256i scan8(__m256i v)
{
    v = _mm256_add_epi32(v, _mm256_slli_si256(v, 4));
    v = _mm256_add_epi32(v, _mm256_slli_si256(v, 8));                       // inclusive inside each 128-bit half
    const __m256i low_total = _mm256_permutevar8x32_epi32(v, _mm256_set1_epi32(3));
    return _mm256_add_epi32(v, _mm256_blend_epi32(_mm256_setzero_si256(), low_total, 0xf0));
}

// The result is written once and not read back by these threads: non-temporal stores skip the read-for-ownership of
// every output line, which is half of the memory traffic of this loop (the host-buffer calls are bound by the host's
// memory bandwidth once the PCIe bytes are halved).  ``out`` is 32-byte aligned here.
__attribute__((target("avx2"))) inline void store8(int32_t *out, __m256i v)
{
    _mm256_stream_si256(reinterpret_cast<__m256i *>(out), v);
}

__attribute__((target("avx2"))) inline void store8(double *out, __m256i v)
{
    _mm256_stream_pd(out, _mm256_cvtepi32_pd(_mm256_castsi256_si128(v)));
    _mm256_stream_pd(out + 4, _mm256_cvtepi32_pd(_mm256_extracti128_si256(v, 1)));
}

__attribute__((target("avx2"))) inline __m256i load8(const uint16_t *d)
{
    return _mm256_cvtepu16_epi32(_mm_loadu_si128(reinterpret_cast<const __m128i *>(d)));
}

__attribute__((target("avx2"))) inline __m256i load8(const uint8_t *d)
{
    return _mm256_cvtepu8_epi32(_mm_loadl_epi64(reinterpret_cast<const __m128i *>(d)));
}

// out[k] = base + sign * (d[0] + ... + d[k]) for k < count; returns the last value (base when count == 0)
template <typename OutT, typename InT>
__attribute__((target("avx2"))) int32_t prefix_avx2(const InT *d, long long count, int32_t base, bool subtract, OutT *out)
{
    int32_t run = base;
    long long k = 0;
    // scalar head up to the first 32-byte boundary of the output
    const long long head = std::min<long long>(count, static_cast<long long>(((32 - (reinterpret_cast<uintptr_t>(out) & 31)) & 31) / sizeof(OutT)));
    for (; k < head; ++k) {
        run = subtract ? run - d[k] : run + d[k];
        out[k] = static_cast<OutT>(run);
    }
    __m256i carry = _mm256_set1_epi32(run);
    for (; k + 8 <= count; k += 8) {
        __m256i v = load8(d + k);
        v = scan8(v);
        v = subtract ? _mm256_sub_epi32(carry, v) : _mm256_add_epi32(carry, v);
        store8(out + k, v);
        carry = _mm256_permutevar8x32_epi32(v, _mm256_set1_epi32(7));
    }
    run = _mm256_extract_epi32(carry, 0);
    for (; k < count; ++k) {
        run = subtract ? run - d[k] : run + d[k];
        out[k] = static_cast<OutT>(run);
    }
    return run;
}

template <typename OutT>
__attribute__((target("avx2"))) void expand_split_row_avx2(const uint8_t *row, long long n, long long head, OutT *out)
{
    const uint16_t *pan_head = reinterpret_cast<const uint16_t *>(row), *core_head = pan_head + head;
    const uint8_t *pan_tail = row + 4 * head, *core_tail = pan_tail + (n - head);
    int32_t run = prefix_avx2<OutT>(pan_head, head, 0, false, out);
    prefix_avx2<OutT>(pan_tail, n - head, run, false, out + head);
    run = core_head[0];
    out[n] = static_cast<OutT>(run);
    run = prefix_avx2<OutT>(core_head + 1, head - 1, run, true, out + n + 1);
    prefix_avx2<OutT>(core_tail, n - head, run, true, out + n + head);
}

template <typename OutT>
__attribute__((target("avx2"))) void expand_row_avx2(const uint16_t *d, long long n, OutT *out)
{
    prefix_avx2<OutT>(d, n, 0, false, out);
    out[n] = static_cast<OutT>(d[n]);
    if (n > 1) prefix_avx2<OutT>(d + n + 1, n - 1, d[n], true, out + n + 1);
}
#endif

template <typename OutT>
void expand_rows(const uint16_t *deltas, long long r0, long long r1, long long n, OutT *out)
{
#if PGX_X86
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2 && (reinterpret_cast<uintptr_t>(out) % sizeof(OutT)) == 0) {
        for (long long r = r0; r < r1; ++r) expand_row_avx2<OutT>(deltas + r * 2 * n, n, out + r * 2 * n);
        _mm_sfence();                          // the streaming stores are visible before the caller is told so
        return;
    }
#endif
    for (long long r = r0; r < r1; ++r) expand_row_scalar<OutT>(deltas + r * 2 * n, n, out + r * 2 * n);
}

template <typename OutT>
void expand_split(const uint8_t *rows, long long r0, long long r1, long long n, long long head, OutT *out)
{
    const long long stride = 2 * n + 2 * head;
#if PGX_X86
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2 && (reinterpret_cast<uintptr_t>(out) % sizeof(OutT)) == 0) {
        for (long long r = r0; r < r1; ++r) expand_split_row_avx2<OutT>(rows + r * stride, n, head, out + r * 2 * n);
        _mm_sfence();
        return;
    }
#endif
    for (long long r = r0; r < r1; ++r) expand_split_row_scalar<OutT>(rows + r * stride, n, head, out + r * 2 * n);
}

}  // namespace

// Rows [r0, r1) of a block in the split format -> curves, on the calling thread.
void expand_split_rows(const uint8_t *rows, long long r0, long long r1, long long n, long long head, void *out, bool out_f64)
{
    if (out_f64) expand_split<double>(rows, r0, r1, n, head, static_cast<double *>(out));
    else expand_split<int32_t>(rows, r0, r1, n, head, static_cast<int32_t *>(out));
}

// Rows [r0, r1) of a block of step rows -> curves (int32 or float64), on the calling thread.
void expand_delta_rows(const uint16_t *deltas, long long r0, long long r1, long long n, void *out, bool out_f64)
{
    if (out_f64) expand_rows<double>(deltas, r0, r1, n, static_cast<double *>(out));
    else expand_rows<int32_t>(deltas, r0, r1, n, static_cast<int32_t *>(out));
}

}  // namespace pgx

extern "C" int pgx_expand_deltas(const uint16_t *h_deltas, int64_t n_rows, int32_t n_genomes, void *h_curves,
                                 int32_t out_f64, int32_t n_threads)
{
    if (n_rows < 0 || n_genomes < 1) return pgx::fail(PGX_ERR_INVALID, "bad shape passed to pgx_expand_deltas");
    if (n_rows == 0) return PGX_OK;
    if (!h_deltas || !h_curves) return pgx::fail(PGX_ERR_INVALID, "null pointer passed to pgx_expand_deltas");
    int threads = n_threads > 0 ? n_threads : static_cast<int>(std::min(8u, std::max(1u, std::thread::hardware_concurrency())));
    threads = static_cast<int>(std::max<long long>(1, std::min<long long>(threads, n_rows * n_genomes / 65536)));
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t)
        pool.emplace_back([=]() {
            pgx::expand_delta_rows(h_deltas, n_rows * t / threads, n_rows * (t + 1) / threads, n_genomes, h_curves, out_f64 != 0);
        });
    pgx::expand_delta_rows(h_deltas, 0, n_rows / threads, n_genomes, h_curves, out_f64 != 0);
    for (auto &th : pool) th.join();
    return PGX_OK;
}

extern "C" int pgx_expand_split(const uint8_t *h_rows, int64_t n_rows, int32_t n_genomes, int32_t head, void *h_curves,
                                int32_t out_f64, int32_t n_threads)
{
    if (n_rows < 0 || n_genomes < 1 || head < 1 || head > n_genomes)
        return pgx::fail(PGX_ERR_INVALID, "bad shape passed to pgx_expand_split (1 <= head <= n_genomes)");
    if (n_rows == 0) return PGX_OK;
    if (!h_rows || !h_curves) return pgx::fail(PGX_ERR_INVALID, "null pointer passed to pgx_expand_split");
    int threads = n_threads > 0 ? n_threads : static_cast<int>(std::min(8u, std::max(1u, std::thread::hardware_concurrency())));
    threads = static_cast<int>(std::max<long long>(1, std::min<long long>(threads, n_rows * n_genomes / 65536)));
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t)
        pool.emplace_back([=]() {
            pgx::expand_split_rows(h_rows, n_rows * t / threads, n_rows * (t + 1) / threads, n_genomes, head, h_curves, out_f64 != 0);
        });
    pgx::expand_split_rows(h_rows, 0, n_rows / threads, n_genomes, head, h_curves, out_f64 != 0);
    for (auto &th : pool) th.join();
    return PGX_OK;
}

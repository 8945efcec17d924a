// Pan/core rarefaction on sm_100a.
//
// Replaces the double loop of estimate_pan_core_size
// (/root/reference/pangenomix/pangenome_analysis.py:81-90).  For one genome order pi with
// rank = pi^-1, a gene g enters the pan-genome at step fp_g = min rank of its present
// genomes and leaves the core genome at step fa_g = min rank of its absent genomes, so
//   pan[k]  = #{g : fp_g <= k} = cumsum(hist(fp))[k]
//   core[k] = #{g : fa_g >  k} = G - cumsum(hist(fa))[k]
// (bit-exact with :89-:90; the identity itself is under test in tests/test_oracle_golden.py).
// Exactly one of fp_g, fa_g is 0 -- the rank-0 genome either has the gene or not -- so bin 0
// of both histograms is a column sum of the table (closed form); the row kernels only ever
// produce the ONE non-zero statistic of every (gene, permutation).
//
// Kernel 1 (list_kernel<B>): genes with a short list.  Persistent CTAs; per item a CTA stages the
//   rank tables of B permutations in shared memory as T[genome][B] uint16 -- one 16-byte line
//   per genome for B = 8 -- and its warps stream runs of sub-blocks of 32 rows, one lane per row,
//   4 chunk loads in flight.  Every index costs ONE shared-memory gather that serves all B
//   permutations, folded into packed uint16x2 running minima (VIMNMX.U16x2).  The host ordered
//   the indices of a row so that the lanes of a wavefront hit distinct banks.  If the min is 0
//   the wanted statistic is the mex of the list's ranks instead: such (row, permutation) events
//   are queued per warp and resolved 32 at a time against the sorted copy of the lists.
// Kernel 2 (probe_kernel<W>): genes with long lists, stored as a genome-major, bit-sliced
//   bitmap (32 genes per word).  A warp walks one genome order for 1,024 W genes at once, one
//   coalesced line of 128 W bytes per step; the first genome whose bit differs from the rank-0
//   genome's bit is the gene's statistic.  O(N / m) steps instead of O(m) gathers.  It runs
//   beside kernel 1 on a second stream (shared-memory-pipe bound vs issue bound).
// Kernel 3 (scan_kernel): adds the closed-form gene classes and turns each histogram into
//   its curve with a block-wide prefix scan, in place.
// Host entry points at the end of the file: the device-pointer calls, the host-buffer pipeline
// (pgx_pan_core_curves_host) and the reference's whole loop in one call (pgx_estimate_pan_core).
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "pgx_common.cuh"

namespace pgx {

namespace {

constexpr int SENTINELS = 32;          // rank-table rows N .. N+31 hold 0xffff (plan.py)
// Optional CUDA-event brackets around the kernels of every call (bench.py's roofline).
struct ProfileEvents { cudaEvent_t begin, list_done, probe_done, end; };
bool g_profile_on = false;
std::mutex g_profile_mu;
std::vector<ProfileEvents> g_profile_events;

struct Tuning {
    int perms_per_cta = 0;
    int row_splits = 0;
    int threads = 0;
};
Tuning g_tuning;
const bool g_no_overlap = getenv("PGX_NO_OVERLAP") != nullptr;

template <int B>
struct Packed {
    static constexpr int REGS = (B + 1) / 2;
};

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

template <int B>
__device__ __forceinline__ void gather_chunk(const uint16_t *table, const uint4 v,
                                             uint32_t (&acc)[Packed<B>::REGS])
{
    gather_min<B>(table, v.x & 0xffffu, acc);
    gather_min<B>(table, v.x >> 16, acc);
    gather_min<B>(table, v.y & 0xffffu, acc);
    gather_min<B>(table, v.y >> 16, acc);
    gather_min<B>(table, v.z & 0xffffu, acc);
    gather_min<B>(table, v.z >> 16, acc);
    gather_min<B>(table, v.w & 0xffffu, acc);
    gather_min<B>(table, v.w >> 16, acc);
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

// Deferred mex events of a warp: (list row << 4 | absent_list << 3 | permutation slot).
constexpr int EVENT_QUEUE = 64;
constexpr int LIST_DEPTH = 4;           // chunk loads in flight per lane

// 48 registers x 1,024 threads leave a quarter of the register file to the probe kernel's CTAs,
// which run beside this one on a second stream (the list kernel is bound by the shared-memory
// pipe, the probe kernel by instruction issue).
template <int B>
__global__ void __maxnreg__(48)
list_kernel(const pgx_plan plan, const uint16_t *__restrict__ perms, const long long n_perm,
            int32_t *__restrict__ hist, const int splits, const long long n_items)
{
    extern __shared__ __align__(16) uint16_t table[];   // [(N + 32)][B], then the warps' event queues
    __shared__ int s_next_task;

    constexpr int REGS = Packed<B>::REGS;
    const int n = plan.n_genomes;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    uint32_t *queue = reinterpret_cast<uint32_t *>(table + ((static_cast<size_t>(n + SENTINELS) * B + 7) & ~size_t(7))) +
                      (tid >> 5) * EVENT_QUEUE;
    const uint4 *__restrict__ chunks = reinterpret_cast<const uint4 *>(plan.d_chunks);
    const int4 *__restrict__ tasks = reinterpret_cast<const int4 *>(plan.d_tasks);
    const long long row_stride = 2ll * n;

    // Persistent CTAs: item = (batch of B permutations, share ``split`` of the tasks).  The whole
    // grid is resident at once, so CTAs of the probe kernel can fill the rest of every SM.
    for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
    // split-major: CTAs that run at the same time stream the SAME share of the chunks for different
    // batches, so a table whose chunks exceed L2 (C5: 128 MB) is read from HBM once per launch, not per batch
    const long long batches = n_items / splits;
    const long long p0 = (item % batches) * B;
    const int split = static_cast<int>(item / batches);
    const int n_valid = static_cast<int>(min(static_cast<long long>(B), n_perm - p0));

    // ---- stage the inverse permutations: T[perm[k]][q] = k (B independent loads in flight) ----
    for (int k = tid; k < n; k += blockDim.x) {
        uint32_t g[B];
#pragma unroll
        for (int q = 0; q < B; ++q) g[q] = q < n_valid ? perms[(p0 + q) * n + k] : static_cast<uint32_t>(k);
#pragma unroll
        for (int q = 0; q < B; ++q) table[g[q] * B + q] = q < n_valid ? static_cast<uint16_t>(k) : static_cast<uint16_t>(0xffffu);
    }
    for (int k = tid; k < SENTINELS * B; k += blockDim.x) table[n * B + k] = 0xffffu;   // never win a min
    if (tid == 0) s_next_task = 0;
    __syncthreads();

    int queued = 0;                                       // warp-uniform

    auto next_task = [&]() -> int4 {
        int t = 0;
        if (lane == 0) t = atomicAdd(&s_next_task, 1);
        t = __shfl_sync(FULL_MASK, t, 0);
        const long long task = static_cast<long long>(t) * splits + split;
        return task < plan.n_tasks ? __ldg(tasks + task) : make_int4(0, 0, 0, 0);
    };
    // mex of one queued (row, permutation): all lanes of the warp work on different events
    auto resolve = [&](uint32_t ev) {
        const int q = ev & 7, row = ev >> 4;
        PGX_DEVICE_CHECK(row < plan.n_rows && q < n_valid);
        const int s0 = plan.d_sorted_ptr[row];
        const int k = mex_probe(perms + (p0 + q) * n, plan.d_sorted_idx + s0, plan.d_sorted_ptr[row + 1] - s0, n);
        // present list: min -> pan histogram, mex -> core histogram; absent list: swapped.
        if (k < n) atomicAdd(hist + (p0 + q) * row_stride + ((ev & 8) ? 0 : n) + k, 1);
    };

    // A task is a run of consecutive sub-blocks (32 rows x nch chunks each) laid out back to
    // back, so the warp streams chunk iterations 0 .. n_sub * nch - 1 with LIST_DEPTH loads in
    // flight and closes a sub-block every nch iterations.  The next task's descriptor is
    // fetched while the current one streams.
    int4 td = next_task();
    while (td.y & 0xffff) {
        const int4 td_next = next_task();
        const int nch = td.y & 0xffff;
        const uint32_t absent_list = (td.y >> 24) & 1;
        const int n_rows = td.w;
        const int total = ((n_rows + 31) >> 5) * nch;
        const uint4 *cp = chunks + td.x + lane;
        PGX_DEVICE_CHECK(td.x >= 0 && static_cast<long long>(td.x) + static_cast<long long>(total) * 32 <= plan.n_chunks);
        PGX_DEVICE_CHECK(td.z >= 0 && td.z + n_rows <= plan.n_rows);
        int32_t *list_hist = hist + (absent_list ? n : 0);

        uint4 buf[LIST_DEPTH];
#pragma unroll
        for (int d = 0; d < LIST_DEPTH; ++d) buf[d] = d < total ? ldg_stream(cp + d * 32) : make_uint4(0, 0, 0, 0);

        uint32_t acc[REGS];
#pragma unroll
        for (int i = 0; i < REGS; ++i) acc[i] = 0xffffffffu;
        int left = nch, row = td.z + lane, rows_left = n_rows - lane;
        for (int s = 0; s < total; ++s) {
            const uint4 cur = buf[0];
#pragma unroll
            for (int d = 0; d + 1 < LIST_DEPTH; ++d) buf[d] = buf[d + 1];
            if (s + LIST_DEPTH < total) buf[LIST_DEPTH - 1] = ldg_stream(cp + (s + LIST_DEPTH) * 32);
            PGX_DEVICE_CHECK((cur.x & 0xffffu) < static_cast<uint32_t>(n + SENTINELS) && (cur.x >> 16) < static_cast<uint32_t>(n + SENTINELS) &&
                             (cur.y & 0xffffu) < static_cast<uint32_t>(n + SENTINELS) && (cur.y >> 16) < static_cast<uint32_t>(n + SENTINELS) &&
                             (cur.z & 0xffffu) < static_cast<uint32_t>(n + SENTINELS) && (cur.z >> 16) < static_cast<uint32_t>(n + SENTINELS) &&
                             (cur.w & 0xffffu) < static_cast<uint32_t>(n + SENTINELS) && (cur.w >> 16) < static_cast<uint32_t>(n + SENTINELS));
            gather_chunk<B>(table, cur, acc);
            if (--left) continue;

            // ---- a sub-block is complete: one statistic per (row, permutation) ----
            const bool valid = rows_left > 0;
#pragma unroll
            for (int q = 0; q < B; ++q) {
                if (q < n_valid) {
                    uint32_t mn;
                    if constexpr (B == 1) mn = acc[0] & 0xffffu;
                    else mn = (acc[q >> 1] >> ((q & 1) * 16)) & 0xffffu;
                    PGX_DEVICE_CHECK(!valid || mn < static_cast<uint32_t>(n));          // a real row always holds a genome
                    if (valid && mn != 0) atomicAdd(list_hist + (p0 + q) * row_stride + mn, 1);
                    // min == 0: the wanted statistic is the mex; queue it, resolve 32 at a time
                    const bool ev = valid && mn == 0;
                    const uint32_t m = __ballot_sync(FULL_MASK, ev);
                    if (m) {
                        PGX_DEVICE_CHECK(queued + __popc(m) <= EVENT_QUEUE);
                    if (ev) queue[queued + __popc(m & ((1u << lane) - 1u))] =
                            (static_cast<uint32_t>(row) << 4) | (absent_list << 3) | q;
                        queued += __popc(m);
                        __syncwarp();
                        if (queued >= 32) {
                            queued -= 32;
                            const uint32_t e = queue[queued + lane];
                            __syncwarp();
                            resolve(e);
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < REGS; ++i) acc[i] = 0xffffffffu;
            left = nch;
            row += 32;
            rows_left -= 32;
        }
        td = td_next;
    }
    if (lane < queued) resolve(queue[lane]);
    __syncthreads();                                      // every warp is done with this table
    }
}

// Bitmap rows, bit-sliced: 32 genes per word, genome-major.  A warp owns one superblock of
// 1,024 W bitmap rows (W = 1, 2 or 4 words per lane) and one permutation; lane l holds rows
// 32 W l .. 32 W l + 32 W - 1 of the superblock as W words.  Walking the genome order, step k
// loads ONE coalesced line of 128 W bytes -- the presence bits of all 1,024 W genes in genome
// perm[k] -- and
//     flipped = (line ^ line_of_rank_0) & pending
// marks the genes whose bit differs from the rank-0 genome's for the first time: k is their
// statistic (first presence if the rank-0 bit was 0, first absence if it was 1).  The walk
// stops when no gene of the superblock is pending: O(N / m) steps of O(1) work per 32 genes,
// the reference's own genome-by-genome accumulation (:87-:90) restricted to unresolved genes.
constexpr int SLICE_WARPS = 4;

template <int W>
struct SliceVec;
template <>
struct SliceVec<1> { using type = uint32_t; };
template <>
struct SliceVec<2> { using type = uint2; };
template <>
struct SliceVec<4> { using type = uint4; };

template <int W>
__device__ __forceinline__ void load_line(const uint32_t *p, uint32_t (&w)[W])
{
    const typename SliceVec<W>::type v = __ldg(reinterpret_cast<const typename SliceVec<W>::type *>(p));
    if constexpr (W == 1) {
        w[0] = v;
    } else if constexpr (W == 2) {
        w[0] = v.x;
        w[1] = v.y;
    } else {
        w[0] = v.x;
        w[1] = v.y;
        w[2] = v.z;
        w[3] = v.w;
    }
}

template <int W>
__global__ void __launch_bounds__(SLICE_WARPS * 32, 8)
probe_kernel(const pgx_plan plan, const uint16_t *__restrict__ perms, const long long n_perm,
             int32_t *__restrict__ hist)
{
    constexpr int DEPTH = W == 4 ? 4 : 8;          // lines in flight per warp
    const int n = plan.n_genomes;
    const int lane = threadIdx.x & 31;
    const long long unit = static_cast<long long>(blockIdx.x) * SLICE_WARPS + (threadIdx.x >> 5);
    if (unit >= n_perm * plan.n_superblocks) return;           // whole warps leave together
    const long long sb = unit / n_perm, q = unit - sb * n_perm;   // costly superblocks first

    const uint32_t *__restrict__ lines = plan.d_bits + (static_cast<size_t>(sb) * n) * (32 * W) + lane * W;
    const uint16_t *__restrict__ perm = perms + q * n;
    int32_t *out = hist + q * 2ll * n + (lane == 1 ? n : 0);      // lane 0 adds pan counts, lane 1 core counts

    uint32_t pending[W], b0[W];
#pragma unroll
    for (int j = 0; j < W; ++j) {
        const long long left = static_cast<long long>(plan.n_long) - (sb * (1024 * W) + (lane * W + j) * 32);
        pending[j] = left >= 32 ? 0xffffffffu : (left <= 0 ? 0u : ((1u << left) - 1u));
    }
    const uint32_t genome0 = perm[0];
    load_line<W>(lines + static_cast<size_t>(genome0) * (32 * W), b0);

    for (int k0 = 0; k0 < n; k0 += 32) {
        // genomes of ranks k0 .. k0 + 31; ranks past the end repeat the rank-0 genome (never a flip)
        const uint32_t chunk = k0 + lane < n ? perm[k0 + lane] : genome0;
#pragma unroll 1
        for (int j0 = 0; j0 < 32; j0 += DEPTH) {
            uint32_t d[DEPTH][W];
            uint32_t any = 0;
#pragma unroll
            for (int j = 0; j < DEPTH; ++j) {
                const uint32_t c = __shfl_sync(FULL_MASK, chunk, j0 + j);
                PGX_DEVICE_CHECK(c < static_cast<uint32_t>(n));
                load_line<W>(lines + static_cast<size_t>(c) * (32 * W), d[j]);
#pragma unroll
                for (int x = 0; x < W; ++x) {
                    d[j][x] ^= b0[x];
                    any |= d[j][x] & pending[x];
                }
            }
            // late in the walk few genes are pending and most groups of genomes flip none
            if (!__any_sync(FULL_MASK, any)) continue;
#pragma unroll
            for (int j = 0; j < DEPTH; ++j) {
                uint32_t flipped[W], some = 0;
#pragma unroll
                for (int x = 0; x < W; ++x) {
                    flipped[x] = d[j][x] & pending[x];
                    some |= flipped[x];
                }
                if (!__any_sync(FULL_MASK, some)) continue;        // mid / late walk: most single steps flip nothing
                // low half: genes first seen at k (pan side); high half: genes first missed at k (core side)
                uint32_t packed = 0;
#pragma unroll
                for (int x = 0; x < W; ++x) {
                    pending[x] &= ~flipped[x];
                    packed += __popc(flipped[x] & ~b0[x]) | (__popc(flipped[x] & b0[x]) << 16);
                }
                const uint32_t total = __reduce_add_sync(FULL_MASK, packed);     // <= 4,096 per half
                const uint32_t mine = lane == 1 ? total >> 16 : total & 0xffffu;
                PGX_DEVICE_CHECK(k0 + j0 + j > 0 && k0 + j0 + j < n);
                if (lane < 2 && mine) atomicAdd(out + (k0 + j0 + j), static_cast<int>(mine));
            }
            uint32_t left = 0;
#pragma unroll
            for (int x = 0; x < W; ++x) left |= pending[x];
            if (!__any_sync(FULL_MASK, left)) return;
        }
    }
    PGX_DEVICE_CHECK(false && "a bitmap row never flipped: it is empty or universal and must not be in the bitmap");
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
    // Closed forms.  Bin 0: genes present in (pan) / absent from (core) the first genome.
    // Genes living in / missing from a single genome c: the list-side statistic is rank[c];
    // the other one is 1 if c comes first and 0 otherwise.
    const int32_t *w_list = side == 0 ? plan.d_w_present : plan.d_w_absent;
    const int32_t *w_other = side == 0 ? plan.d_w_absent : plan.d_w_present;
    const int first = perm[0];
    const int col_first = plan.d_colsum[first];
    const int bin0 = side == 0 ? col_first : plan.n_genes - col_first;
    const int add1 = w_other[first];

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
                if (k == 0) x = bin0;
                else x = h[k] + w_list[perm[k]] + (k == 1 ? add1 : 0);
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

// EXPERIMENT, off unless PGX_SCAN_V8=1: the same scan with 16-byte loads and stores, 8 bins per thread, for
// tables whose genome count is a multiple of 8.  Measured on C4: 0.523 -> 0.422 ms per 10,000 permutations; six
// parity tests passed with it, the whole -m gpu suite has not run under it yet (scripts/r02_first.sh does that).
template <typename OutT>
__global__ void __launch_bounds__(256)
scan_kernel_v8(const pgx_plan plan, const uint16_t *__restrict__ perms, const int32_t *hist, OutT *out)
{
    constexpr int ITEMS = 8;
    constexpr int THREADS = 256;
    __shared__ int warp_tot[THREADS / 32];

    const int n = plan.n_genomes;                    // n % 8 == 0 (checked by the launcher)
    const long long p = blockIdx.x;
    const int side = blockIdx.y;
    const int32_t *h = hist + p * 2ll * n + static_cast<long long>(side) * n;
    OutT *o = out + p * 2ll * n + static_cast<long long>(side) * n;
    const uint16_t *perm = perms + p * n;
    const int32_t *w_list = side == 0 ? plan.d_w_present : plan.d_w_absent;
    const int32_t *w_other = side == 0 ? plan.d_w_absent : plan.d_w_present;
    const int first = perm[0];
    const int col_first = plan.d_colsum[first];
    const int bin0 = side == 0 ? col_first : plan.n_genes - col_first;
    const int add1 = w_other[first];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int carry = 0;
    for (int base = 0; base < n; base += THREADS * ITEMS) {
        const int k0 = base + tid * ITEMS;
        int v[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) v[i] = 0;
        if (k0 < n) {
            const int4 a = *reinterpret_cast<const int4 *>(h + k0);
            const int4 b = *reinterpret_cast<const int4 *>(h + k0 + 4);
            const uint4 pr = *reinterpret_cast<const uint4 *>(perm + k0);
            v[0] = a.x + w_list[pr.x & 0xffffu];
            v[1] = a.y + w_list[pr.x >> 16];
            v[2] = a.z + w_list[pr.y & 0xffffu];
            v[3] = a.w + w_list[pr.y >> 16];
            v[4] = b.x + w_list[pr.z & 0xffffu];
            v[5] = b.y + w_list[pr.z >> 16];
            v[6] = b.z + w_list[pr.w & 0xffffu];
            v[7] = b.w + w_list[pr.w >> 16];
            if (k0 == 0) {
                v[0] = bin0;
                v[1] += add1;
            }
        }
        int run = 0;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            run += v[i];
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
        if (k0 < n) {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const int c = excl + v[i];
                v[i] = side == 0 ? c : plan.n_genes - c;
            }
            if constexpr (sizeof(OutT) == 4) {
                *reinterpret_cast<int4 *>(o + k0) = make_int4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<int4 *>(o + k0 + 4) = make_int4(v[4], v[5], v[6], v[7]);
            } else {
#pragma unroll
                for (int i = 0; i < ITEMS; i += 2)
                    *reinterpret_cast<double2 *>(o + k0 + i) = make_double2(static_cast<double>(v[i]), static_cast<double>(v[i + 1]));
            }
        }
        carry += total;
        __syncthreads();
    }
}

int check_plan(const pgx_plan *plan)
{
    if (!plan) return fail(PGX_ERR_INVALID, "plan is null");
    if (plan->n_genomes < 1 || plan->n_genomes > 65503)
        return fail(PGX_ERR_UNSUPPORTED, "n_genomes = %d outside 1..65503 (uint16 genome indices)",
                    plan->n_genomes);
    if (plan->n_genes < 0 || plan->n_rows < 0 || plan->n_tasks < 0 || plan->n_chunks < 0 || plan->n_long < 0)
        return fail(PGX_ERR_INVALID, "negative size in plan");
    if (!plan->d_w_present || !plan->d_w_absent || !plan->d_colsum)
        return fail(PGX_ERR_INVALID, "plan closed-form vectors are null");
    if (plan->n_tasks > 0 && (!plan->d_chunks || !plan->d_tasks || !plan->d_sorted_idx || !plan->d_sorted_ptr))
        return fail(PGX_ERR_INVALID, "plan has list tasks but null list arrays");
    if (plan->n_long > 0 && !plan->d_bits)
        return fail(PGX_ERR_INVALID, "plan has bitmap rows but a null bitmap");
    if (plan->n_long > 0 && plan->slice_words != 1 && plan->slice_words != 2 && plan->slice_words != 4)
        return fail(PGX_ERR_INVALID, "slice_words must be 1, 2 or 4");
    if (plan->n_long > 0 && plan->n_superblocks != (plan->n_long + 1024 * plan->slice_words - 1) / (1024 * plan->slice_words))
        return fail(PGX_ERR_INVALID, "n_superblocks must be ceil(n_long / (1024 * slice_words))");
    if ((reinterpret_cast<uintptr_t>(plan->d_chunks) | reinterpret_cast<uintptr_t>(plan->d_tasks) |
         reinterpret_cast<uintptr_t>(plan->d_bits)) & 15)
        return fail(PGX_ERR_INVALID, "d_chunks, d_tasks and d_bits must be 16-byte aligned");
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
int launch_list(const pgx_plan &plan, const uint16_t *d_perms, long long n_perm, int32_t *d_hist,
                const DeviceLimits &lim, cudaStream_t stream)
{
    int threads = g_tuning.threads;
    if (threads <= 0 && getenv("PGX_LIST_THREADS")) threads = atoi(getenv("PGX_LIST_THREADS"));
    const size_t table_bytes = ((static_cast<size_t>(plan.n_genomes + SENTINELS) * B + 7) & ~size_t(7)) * sizeof(uint16_t);
    if (threads <= 0) threads = table_bytes > 100 * 1024 ? 1024 : (table_bytes > 40 * 1024 ? 512 : 256);
    threads = max(32, min(1024, (threads / 32) * 32));
    const size_t smem = table_bytes + static_cast<size_t>(threads / 32) * EVENT_QUEUE * sizeof(uint32_t);
    PGX_CUDA(cudaFuncSetAttribute(list_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
    const long long batches = (n_perm + B - 1) / B;
    // Row splits: enough CTAs for ~16 waves so the tail is small, but every CTA keeps at
    // least a few tasks per warp (the table build is amortised over them).
    int splits = g_tuning.row_splits;
    if (splits <= 0) {
        const long long resident = static_cast<long long>(lim.sm_count) *
                                   max(1, min(8, static_cast<int>((200 * 1024) / (smem + 1024))));
        const long long want = (16 * resident + batches - 1) / batches;
        const long long most = max(1ll, static_cast<long long>(plan.n_tasks) / ((threads / 32) * 2));
        splits = static_cast<int>(max(1ll, min(want, most)));
    }
    splits = max(1, min(splits, 65535));
    const long long n_items = batches * splits;
    int per_sm = 1;
    PGX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, list_kernel<B>, threads, smem));
    const long long grid = max(1ll, min(n_items, static_cast<long long>(lim.sm_count) * max(1, per_sm)));
    list_kernel<B><<<static_cast<unsigned>(grid), threads, smem, stream>>>(plan, d_perms, n_perm, d_hist, splits, n_items);
    PGX_LAUNCH_CHECK("list_kernel");
    return PGX_OK;
}

int launch_probe(const pgx_plan &plan, const uint16_t *d_perms, long long n_perm, int32_t *d_hist,
                 cudaStream_t stream)
{
    // one warp per (superblock, permutation); 2^31 blocks of 4 warps cover any realistic call
    const long long units = n_perm * plan.n_superblocks;
    const long long blocks = (units + SLICE_WARPS - 1) / SLICE_WARPS;
    if (blocks > 2147483647ll) return fail(PGX_ERR_UNSUPPORTED, "too many (superblock, permutation) units in one call");
    switch (plan.slice_words) {
        case 4: probe_kernel<4><<<static_cast<unsigned>(blocks), SLICE_WARPS * 32, 0, stream>>>(plan, d_perms, n_perm, d_hist); break;
        case 2: probe_kernel<2><<<static_cast<unsigned>(blocks), SLICE_WARPS * 32, 0, stream>>>(plan, d_perms, n_perm, d_hist); break;
        default: probe_kernel<1><<<static_cast<unsigned>(blocks), SLICE_WARPS * 32, 0, stream>>>(plan, d_perms, n_perm, d_hist); break;
    }
    PGX_LAUNCH_CHECK("probe_kernel");
    return PGX_OK;
}

// Second stream + fork/join events of the calling thread's current device.
struct Aux {
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};

int aux_for_device(Aux **out)
{
    static thread_local Aux aux[64];
    int dev = 0;
    PGX_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(PGX_ERR_UNSUPPORTED, "device ordinal %d out of range", dev);
    Aux &a = aux[dev];
    if (a.device != dev) {
        PGX_CUDA(cudaStreamCreateWithFlags(&a.stream, cudaStreamNonBlocking));
        PGX_CUDA(cudaEventCreateWithFlags(&a.fork, cudaEventDisableTiming));
        PGX_CUDA(cudaEventCreateWithFlags(&a.join, cudaEventDisableTiming));
        a.device = dev;
    }
    *out = &a;
    return PGX_OK;
}

template <typename OutT>
int run_curves(const pgx_plan *plan, const uint16_t *d_perms, long long n_perm, int32_t *d_hist,
               OutT *d_out, cudaStream_t stream, Aux *own_aux = nullptr)
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
        PGX_CUDA(cudaEventCreate(&ev.list_done));
        PGX_CUDA(cudaEventCreate(&ev.probe_done));
        PGX_CUDA(cudaEventCreate(&ev.end));
        PGX_CUDA(cudaEventRecord(ev.begin, stream));
    }
    // The two row kernels only add into the histogram, in any order: unless per-kernel timing is
    // on (or PGX_NO_OVERLAP is set), the probe kernel runs beside the list kernel on a second stream.
    Aux *aux = own_aux;
    const bool overlap = !profile && !g_no_overlap && plan->n_tasks > 0 && plan->n_long > 0;
    if (overlap) {
        if (!aux) {
            if (int rc = aux_for_device(&aux)) return rc;
        }
        PGX_CUDA(cudaEventRecord(aux->fork, stream));
        PGX_CUDA(cudaStreamWaitEvent(aux->stream, aux->fork, 0));
    }
    if (plan->n_tasks > 0) {
        const size_t per_perm = static_cast<size_t>(n + SENTINELS) * sizeof(uint16_t);
        const size_t budget = static_cast<size_t>(lim.smem_optin) - 64 - 32 * EVENT_QUEUE * sizeof(uint32_t);
        int b = g_tuning.perms_per_cta;
        if (b != 1 && b != 2 && b != 4 && b != 8) b = plan->perms_per_cta;
        if (b != 1 && b != 2 && b != 4 && b != 8) b = 8;
        while (b > 1 && per_perm * b > budget) b >>= 1;
        if (per_perm * b > budget) return fail(PGX_ERR_UNSUPPORTED, "rank table does not fit shared memory");
        int rc;
        switch (b) {
            case 8: rc = launch_list<8>(*plan, d_perms, n_perm, d_hist, lim, stream); break;
            case 4: rc = launch_list<4>(*plan, d_perms, n_perm, d_hist, lim, stream); break;
            case 2: rc = launch_list<2>(*plan, d_perms, n_perm, d_hist, lim, stream); break;
            default: rc = launch_list<1>(*plan, d_perms, n_perm, d_hist, lim, stream); break;
        }
        if (rc) return rc;
    }
    if (profile) PGX_CUDA(cudaEventRecord(ev.list_done, stream));
    if (overlap) {
        // launched AFTER the list kernel: one list CTA per SM first, probe CTAs fill what is left
        if (int rc = launch_probe(*plan, d_perms, n_perm, d_hist, aux->stream)) return rc;
        PGX_CUDA(cudaEventRecord(aux->join, aux->stream));
        PGX_CUDA(cudaStreamWaitEvent(stream, aux->join, 0));
    } else if (plan->n_long > 0) {
        if (int rc = launch_probe(*plan, d_perms, n_perm, d_hist, stream)) return rc;
    }
    if (profile) PGX_CUDA(cudaEventRecord(ev.probe_done, stream));
    for (long long p0 = 0; p0 < n_perm; p0 += 2147483647ll) {
        const long long np = min(2147483647ll, n_perm - p0);
        dim3 grid(static_cast<unsigned>(np), 2);
        static const bool scan_v8 = getenv("PGX_SCAN_V8") != nullptr;       // experiment, see scan_kernel_v8
        if (scan_v8 && n % 8 == 0)
            scan_kernel_v8<OutT><<<grid, 256, 0, stream>>>(*plan, d_perms + p0 * n, d_hist + p0 * 2ll * n,
                                                           d_out + p0 * 2ll * n);
        else
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

    // Three slots: while one block's curves travel to the host, the next block's kernels run and the one
    // after that uploads its permutations, so both PCIe directions and the SMs stay busy.
    constexpr int SLOTS = 3;
    cudaStream_t streams[SLOTS] = {nullptr, nullptr, nullptr};
    uint16_t *d_perms[SLOTS] = {nullptr, nullptr, nullptr};
    int32_t *d_hist[SLOTS] = {nullptr, nullptr, nullptr};
    double *d_out[SLOTS] = {nullptr, nullptr, nullptr};
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
        for (int s = 0; s < SLOTS; ++s) {
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
    const int n_streams = static_cast<int>(std::min<long long>(SLOTS, (n_perm + block - 1) / block));
    for (int s = 0; s < n_streams; ++s) {
        PGX_TRY(cudaStreamCreateWithFlags(&streams[s], cudaStreamNonBlocking));
        PGX_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_perms[s]), perm_bytes, streams[s]));
        PGX_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_hist[s]), hist_bytes, streams[s]));
        if (out_f64) PGX_TRY(cudaMallocAsync(reinterpret_cast<void **>(&d_out[s]), out_bytes, streams[s]));
    }
    int slot = 0;
    for (long long p0 = 0; p0 < n_perm; p0 += block, slot = (slot + 1) % n_streams) {
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

// estimate_pan_core_size in one call (pangenome_analysis.py:76-90): RNG stream, upload, kernels,
// download.  Three slots of pinned staging + device buffers (cached per device for the life of the
// library); a producer thread draws the shuffles of a block and enqueues its H2D copy, kernels and
// D2H copy, the calling thread waits for finished blocks in order and moves them into the caller's
// (ordinary, pageable) result.
namespace pgx {
namespace {

struct EstimateSlot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    Aux aux;                          // second stream + fork/join events of this slot's kernels
    uint16_t *h_perm = nullptr, *d_perm = nullptr;
    int32_t *d_hist = nullptr;
    double *d_out = nullptr, *h_out = nullptr;
};

struct EstimateBuffers {
    int device = -1;
    long long block = 0, n = 0;
    EstimateSlot slot[3];
};

void release(EstimateBuffers &b)
{
    for (auto &s : b.slot) {
        if (s.h_perm) cudaFreeHost(s.h_perm);
        if (s.h_out) cudaFreeHost(s.h_out);
        if (s.d_perm) cudaFree(s.d_perm);
        if (s.d_hist) cudaFree(s.d_hist);
        if (s.d_out) cudaFree(s.d_out);
        if (s.done) cudaEventDestroy(s.done);
        if (s.stream) cudaStreamDestroy(s.stream);
        if (s.aux.stream) cudaStreamDestroy(s.aux.stream);
        if (s.aux.fork) cudaEventDestroy(s.aux.fork);
        if (s.aux.join) cudaEventDestroy(s.aux.join);
        s = EstimateSlot{};
    }
    b.device = -1;
    b.block = b.n = 0;
}

int acquire(EstimateBuffers &b, int device, long long block, long long n)
{
    if (b.device == device && b.n == n && b.block >= block) return PGX_OK;
    release(b);
    for (auto &s : b.slot) {
        PGX_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        PGX_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        PGX_CUDA(cudaStreamCreateWithFlags(&s.aux.stream, cudaStreamNonBlocking));
        PGX_CUDA(cudaEventCreateWithFlags(&s.aux.fork, cudaEventDisableTiming));
        PGX_CUDA(cudaEventCreateWithFlags(&s.aux.join, cudaEventDisableTiming));
        s.aux.device = device;
        PGX_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&s.h_perm), sizeof(uint16_t) * block * n, cudaHostAllocDefault));
        PGX_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&s.h_out), sizeof(double) * block * 2 * n, cudaHostAllocDefault));
        PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&s.d_perm), sizeof(uint16_t) * block * n));
        PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&s.d_hist), sizeof(int32_t) * block * 2 * n));
        PGX_CUDA(cudaMalloc(reinterpret_cast<void **>(&s.d_out), sizeof(double) * block * 2 * n));
    }
    b.device = device;
    b.block = block;
    b.n = n;
    return PGX_OK;
}

std::mutex g_estimate_mu;          // one estimate call at a time per process (it owns the staging)
EstimateBuffers g_estimate_buffers;

}  // namespace
}  // namespace pgx

int pgx_estimate_pan_core(const pgx_plan *plan, uint32_t *mt_key, int32_t *mt_pos, int64_t n_iter,
                          double *h_curves, int64_t perms_per_block)
{
    if (int rc = pgx::check_plan(plan)) return rc;
    if (n_iter < 0) return pgx::fail(PGX_ERR_INVALID, "n_iter < 0");
    if (!mt_key || !mt_pos || (!h_curves && n_iter > 0)) return pgx::fail(PGX_ERR_INVALID, "null pointer");
    if (n_iter == 0) return PGX_OK;
    const long long n = plan->n_genomes;
    long long block = perms_per_block;
    if (block <= 0) block = std::max(32ll, std::min(4096ll, (32ll << 20) / (16 * n)));
    std::lock_guard<std::mutex> lock(pgx::g_estimate_mu);
    int dev = 0;
    PGX_CUDA(cudaGetDevice(&dev));
    pgx::EstimateBuffers &buf = pgx::g_estimate_buffers;
    if (int rc = pgx::acquire(buf, dev, block, n)) return rc;
    block = std::min<long long>(block, n_iter);
    const long long n_blocks = (n_iter + block - 1) / block;
    {
        // The result is usually fresh memory: ask for huge pages so that filling it costs one page
        // fault per 2 MB instead of one per 4 KB (a hint; ignored where THP is off).
        const uintptr_t page = 2u << 20;
        const uintptr_t lo = (reinterpret_cast<uintptr_t>(h_curves) + page - 1) & ~(page - 1);
        const uintptr_t hi = (reinterpret_cast<uintptr_t>(h_curves) + sizeof(double) * 2ull * n * n_iter) & ~(page - 1);
        if (hi > lo) madvise(reinterpret_cast<void *>(lo), hi - lo, MADV_HUGEPAGE);
    }

    // PGX_ESTIMATE_TRACE=1: per-block timeline of the pipeline on stderr (development aid)
    const bool trace = getenv("PGX_ESTIMATE_TRACE") != nullptr;
    const auto trace_t0 = std::chrono::steady_clock::now();
    auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - trace_t0).count(); };

    // producer -> consumer hand-off: ``issued`` blocks have their GPU work enqueued, ``retired`` blocks
    // have been copied out; slot of block k is k % 3, reusable once block k - 3 retired
    std::mutex mu;
    std::condition_variable cv;
    long long issued = 0, retired = 0;
    // EXPERIMENT, off unless PGX_ESTIMATE_SPIN=1 (never run on a GPU yet): the two sides poll atomic copies of the
    // counters instead of sleeping on the condition variable -- a notify that finds a sleeper costs the notifier
    // 0.2-0.5 ms on the virtualised hosts this runs on.
    const bool spin = getenv("PGX_ESTIMATE_SPIN") != nullptr;
    std::atomic<long long> a_issued{0}, a_retired{0};
    auto publish = [&](bool wake) {                    // call with ``mu`` released
        if (!spin || wake) cv.notify_all();
    };
    int producer_rc = PGX_OK;
    char producer_err[512] = "";
    std::thread producer([&]() {
        cudaSetDevice(dev);
        for (long long k = 0; k < n_blocks; ++k) {
            if (spin) {
                while (a_retired.load(std::memory_order_acquire) + 3 <= k) std::this_thread::yield();
            } else {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return retired + 3 > k; });
            }
            pgx::EstimateSlot &s = buf.slot[k % 3];
            const long long p0 = k * block, cnt = std::min<long long>(block, n_iter - p0);
            const double t_wait = since();
            int rc = pgx_legacy_shuffles(mt_key, mt_pos, n, cnt, s.h_perm);
            const double t_rng = since();
            if (!rc && cudaMemcpyAsync(s.d_perm, s.h_perm, sizeof(uint16_t) * cnt * n, cudaMemcpyHostToDevice, s.stream) != cudaSuccess)
                rc = pgx::fail(PGX_ERR_CUDA, "H2D copy of the permutations failed");
            if (!rc) rc = pgx::run_curves<double>(plan, s.d_perm, cnt, s.d_hist, s.d_out, s.stream, &s.aux);
            if (!rc && cudaMemcpyAsync(s.h_out, s.d_out, sizeof(double) * cnt * 2 * n, cudaMemcpyDeviceToHost, s.stream) != cudaSuccess)
                rc = pgx::fail(PGX_ERR_CUDA, "D2H copy of the curves failed");
            if (!rc && cudaEventRecord(s.done, s.stream) != cudaSuccess) rc = pgx::fail(PGX_ERR_CUDA, "event record failed");
            if (trace) fprintf(stderr, "[pgx trace] block %lld: rng %.2f -> %.2f ms, enqueued %.2f ms\n", k, t_wait, t_rng, since());
            {
                std::lock_guard<std::mutex> lk(mu);
                if (rc) {
                    producer_rc = rc;
                    snprintf(producer_err, sizeof(producer_err), "%s", pgx_last_error());   // thread-local text
                }
                issued = rc ? n_blocks : k + 1;
                a_issued.store(issued, std::memory_order_release);
            }
            publish(false);
            if (rc) return;
        }
    });
    int rc = PGX_OK;
    int copy_threads = std::max(1, std::min(6, static_cast<int>(std::thread::hardware_concurrency()) / 2));
    if (const char *env = getenv("PGX_COPY_THREADS")) copy_threads = std::max(1, std::min(16, atoi(env)));
    for (long long k = 0; k < n_blocks; ++k) {
        if (spin) {
            while (a_issued.load(std::memory_order_acquire) <= k) std::this_thread::yield();
        }
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return issued > k; });         // (already true when spinning)
            if (producer_rc) break;
        }
        pgx::EstimateSlot &s = buf.slot[k % 3];
        const long long p0 = k * block, cnt = std::min<long long>(block, n_iter - p0);
        if (cudaEventSynchronize(s.done) != cudaSuccess) {
            rc = pgx::fail(PGX_ERR_CUDA, "a block of curves failed on the device: %s", cudaGetErrorString(cudaGetLastError()));
            std::lock_guard<std::mutex> lk(mu);
            retired = n_blocks + 3;                     // let the producer run out
            a_retired.store(retired, std::memory_order_release);
            cv.notify_all();
            break;
        }
        const double t_ready = since();
        {
            // staging -> result with a few threads: one core fills fresh pages at only ~5 GB/s (2,000 permutations of
            // C4 on the box: 69 / 39 / 28 / 27 ms per call with 1 / 2 / 4 / 6 threads)
            const size_t bytes = sizeof(double) * cnt * 2 * n;
            char *dst = reinterpret_cast<char *>(h_curves + p0 * 2 * n);
            const char *src = reinterpret_cast<const char *>(s.h_out);
            const int parts = static_cast<int>(std::max<size_t>(1, std::min<size_t>(copy_threads, bytes >> 20)));   // >= 1 MB each
            std::vector<std::thread> movers;
            for (int t = 1; t < parts; ++t) {
                const size_t lo = bytes / parts * t, hi = t + 1 == parts ? bytes : bytes / parts * (t + 1);
                movers.emplace_back([=]() { memcpy(dst + lo, src + lo, hi - lo); });
            }
            memcpy(dst, src, parts > 1 ? bytes / parts : bytes);
            for (auto &m : movers) m.join();
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            retired = k + 1;
            a_retired.store(retired, std::memory_order_release);
        }
        publish(false);
        if (trace) fprintf(stderr, "[pgx trace] block %lld: on the host %.2f ms, copied out %.2f ms\n", k, t_ready, since());
    }
    {
        std::lock_guard<std::mutex> lk(mu);
        retired = n_blocks + 3;
        a_retired.store(retired, std::memory_order_release);
    }
    cv.notify_all();
    producer.join();
    if (producer_rc) return pgx::fail(producer_rc, "%s", producer_err);
    return rc;
}

int pgx_profile_enable(int32_t on)
{
    pgx::g_profile_on = on != 0;
    return PGX_OK;
}

int pgx_profile_read(double *list_ms, double *probe_ms, double *scan_ms, int64_t *calls)
{
    std::lock_guard<std::mutex> lock(pgx::g_profile_mu);
    double a = 0.0, b = 0.0, c = 0.0;
    for (auto &ev : pgx::g_profile_events) {
        float t0 = 0.f, t1 = 0.f, t2 = 0.f;
        PGX_CUDA(cudaEventSynchronize(ev.end));
        PGX_CUDA(cudaEventElapsedTime(&t0, ev.begin, ev.list_done));
        PGX_CUDA(cudaEventElapsedTime(&t1, ev.list_done, ev.probe_done));
        PGX_CUDA(cudaEventElapsedTime(&t2, ev.probe_done, ev.end));
        a += t0;
        b += t1;
        c += t2;
        cudaEventDestroy(ev.begin);
        cudaEventDestroy(ev.list_done);
        cudaEventDestroy(ev.probe_done);
        cudaEventDestroy(ev.end);
    }
    if (list_ms) *list_ms = a;
    if (probe_ms) *probe_ms = b;
    if (scan_ms) *scan_ms = c;
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

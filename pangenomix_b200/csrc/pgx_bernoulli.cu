// Bernoulli-grid log-likelihood and gradient in fp64 on sm_100a.
//
// Replaces __bernoulli_grid_loglikelihood__ (/root/reference/pangenomix/pangenome_analysis.py:244-249)
// and __bernoulli_grid_loglikelihood_gradient__ (:257-266) as evaluated by the L-BFGS-B
// lambdas at :156-160.  With m_i = sum_j X_ij and n_j = sum_i X_ij,
//   LL      = sum_i m_i log p_i + sum_j n_j log q_j + sum_{X_ij = 0} log(1 - p_i q_j)
//   dL/dp_i = m_i / p_i - sum_{j : X_ij = 0} q_j / (1 - p_i q_j)
//   dL/dq_j = n_j / q_j - sum_{i : X_ij = 0} p_i / (1 - p_i q_j)
// which is the reference's dense expression with the X_ij = 1 cells summed in closed form.
// X is bit-packed (one bit per cell), so the kernel is fp64-ALU bound (one log and one
// division per absent cell); the table stays L2-resident across optimiser iterations.
//
// grid_kernel: CTA (row tile, column tile of 512 genomes); the warps of the CTA walk the rows
//   of the tile.  For every row the warp compacts the absent columns of its 16 bitmap words
//   into a shared-memory list (popcount prefix over the words) and then every lane evaluates
//   one absent cell per step; dL/dq partial sums live in a per-warp shared-memory row (a
//   column occurs once per row, so there are no collisions and the order is fixed).  Row sums
//   use a fixed shuffle tree, column sums are combined across warps in warp order, so the
//   result is bit-reproducible run to run.
// finish_kernel: fixed-order reduction of the per-tile partials plus the closed-form terms.
#include "pgx_common.cuh"

namespace pgx {

namespace {

constexpr int BG_THREADS = 256;
constexpr int BG_WARPS = BG_THREADS / 32;
constexpr int BG_COLS_PER_LANE = 16;
constexpr int BG_TILE_COLS = 32 * BG_COLS_PER_LANE;   // 512 genomes = 16 bitmap words

struct BernoulliShape {
    long long n_genes, n_genomes, words_per_row;
    int row_tiles, col_tiles, rows_per_tile;
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL_MASK, v, off);
    return v;
}

__global__ void __launch_bounds__(BG_THREADS)
grid_kernel(const uint32_t *__restrict__ xbits, const BernoulliShape shape,
            const int32_t *__restrict__ row_count,
            const double *__restrict__ p, const double *__restrict__ q,
            double *__restrict__ gp_part,   // [col_tiles][G]
            double *__restrict__ gq_part,   // [row_tiles][N]
            double *__restrict__ ll_part)   // [row_tiles * col_tiles]
{
    __shared__ double s_gq[BG_WARPS][BG_TILE_COLS];
    __shared__ double s_q[BG_TILE_COLS];
    __shared__ uint16_t s_cols[BG_WARPS][BG_TILE_COLS];     // per warp: the absent columns of its current row
    __shared__ double s_ll[BG_WARPS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row_tile = blockIdx.x, col_tile = blockIdx.y;
    const long long col0 = static_cast<long long>(col_tile) * BG_TILE_COLS;
    const long long word0 = col0 >> 5;
    const long long row_begin = static_cast<long long>(row_tile) * shape.rows_per_tile;
    const long long row_end = min(shape.n_genes, row_begin + shape.rows_per_tile);

    for (int c = tid; c < BG_TILE_COLS; c += BG_THREADS) {
        s_q[c] = col0 + c < shape.n_genomes ? q[col0 + c] : 0.0;
#pragma unroll
        for (int w = 0; w < BG_WARPS; ++w) s_gq[w][c] = 0.0;
    }
    __syncthreads();
    const int n_words = static_cast<int>(min(static_cast<long long>(BG_COLS_PER_LANE),
                                             shape.words_per_row - word0));
    // columns of word ``lane`` that exist (the last word of a row may be ragged)
    uint32_t valid = 0;
    if (lane < n_words) {
        const long long left = shape.n_genomes - (col0 + 32ll * lane);
        valid = left >= 32 ? 0xffffffffu : (left <= 0 ? 0u : ((1u << left) - 1u));
    }
    uint16_t *cols = s_cols[warp];
    double *gq = s_gq[warp];

    double ll = 0.0;
    for (long long i = row_begin + warp; i < row_end; i += BG_WARPS) {
        const double pi = p[i];
        // Only absent cells cost arithmetic (:248, :261-265 with the X = 1 cells in closed form) and
        // they are few and scattered: compact them so that every lane works.  Lane w < 16 owns
        // bitmap word w of this row's column tile and lists its absent columns at its offset.
        uint32_t absent = 0;
        if (lane < n_words) absent = ~xbits[i * shape.words_per_row + word0 + lane] & valid;
        const int mine = __popc(absent);
        int before = mine;
#pragma unroll
        for (int off = 1; off < 16; off <<= 1) {
            const int v = __shfl_up_sync(FULL_MASK, before, off);
            if (lane >= off) before += v;
        }
        const int total = __shfl_sync(FULL_MASK, before, 15);
        int at = before - mine;
        for (uint32_t m = absent; m; m &= m - 1u) cols[at++] = static_cast<uint16_t>(32 * lane + __ffs(m) - 1);
        __syncwarp();
        double gp = 0.0;
        for (int t = lane; t < total; t += 32) {
            const int c = cols[t];
            const double qj = s_q[c];
            const double u = 1.0 - pi * qj;            // :261  nprobs = 1 - outer(P, Q)
            ll += log(u);                              // :248  (1 - X) * log(1 - probs)
            const double r = 1.0 / u;
            gp += qj * r;                              // :263
            gq[c] += pi * r;                           // :265  (one lane per column within a row)
        }
        __syncwarp();
        gp = warp_sum(gp);
        if (lane == 0) gp_part[static_cast<long long>(col_tile) * shape.n_genes + i] = gp;
        // closed-form term of the present cells of this gene, m_i log p_i, counted once (column tile 0)
        if (lane == 0 && col_tile == 0) ll += static_cast<double>(row_count[i]) * log(pi);
    }

    // combine the warps' column sums in warp order
    ll = warp_sum(ll);
    if (lane == 0) s_ll[warp] = ll;
    __syncthreads();
    for (int c = tid; c < BG_TILE_COLS; c += BG_THREADS) {
        const long long j = col0 + c;
        if (j < shape.n_genomes) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < BG_WARPS; ++w) s += s_gq[w][c];
            gq_part[static_cast<long long>(row_tile) * shape.n_genomes + j] = s;
        }
    }
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < BG_WARPS; ++w) s += s_ll[w];
        ll_part[static_cast<long long>(col_tile) * shape.row_tiles + row_tile] = s;
    }
}

// grad[i] = m_i / p_i - sum_ct gp_part[ct][i];  grad[G + j] = n_j / q_j - sum_rt gq_part[rt][j];
// block 0 additionally reduces the log-likelihood in a fixed order.
__global__ void __launch_bounds__(256)
finish_kernel(const BernoulliShape shape, const int32_t *__restrict__ row_count,
              const int32_t *__restrict__ col_count, const double *__restrict__ p,
              const double *__restrict__ q, const double *__restrict__ gp_part,
              const double *__restrict__ gq_part, const double *__restrict__ ll_part,
              double *__restrict__ ll_out, double *__restrict__ grad)
{
    const long long total = shape.n_genes + shape.n_genomes;
    const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx < shape.n_genes) {
        double s = 0.0;
        for (int ct = 0; ct < shape.col_tiles; ++ct) s += gp_part[ct * shape.n_genes + idx];
        grad[idx] = static_cast<double>(row_count[idx]) / p[idx] - s;
    } else if (idx < total) {
        const long long j = idx - shape.n_genes;
        double s = 0.0;
        for (int rt = 0; rt < shape.row_tiles; ++rt) s += gq_part[rt * shape.n_genomes + j];
        grad[idx] = static_cast<double>(col_count[j]) / q[j] - s;
    }
    if (blockIdx.x == 0) {
        __shared__ double s_part[256];
        double s = 0.0;
        for (long long j = threadIdx.x; j < shape.n_genomes; j += 256)
            s += static_cast<double>(col_count[j]) * log(q[j]);
        const long long n_part = static_cast<long long>(shape.row_tiles) * shape.col_tiles;
        for (long long k = threadIdx.x; k < n_part; k += 256) s += ll_part[k];
        s_part[threadIdx.x] = s;
        __syncthreads();
        for (int off = 128; off > 0; off >>= 1) {
            if (threadIdx.x < off) s_part[threadIdx.x] += s_part[threadIdx.x + off];
            __syncthreads();
        }
        if (threadIdx.x == 0) *ll_out = s_part[0];
    }
}

int make_shape(long long n_genes, long long n_genomes, long long words_per_row, BernoulliShape *s)
{
    if (n_genes < 1 || n_genomes < 1) return fail(PGX_ERR_INVALID, "empty Bernoulli grid");
    if (words_per_row * 32 < n_genomes) return fail(PGX_ERR_INVALID, "words_per_row too small");
    if (n_genes > 2000000000ll || n_genomes > 16000000ll)
        return fail(PGX_ERR_UNSUPPORTED, "Bernoulli grid too large");
    s->n_genes = n_genes;
    s->n_genomes = n_genomes;
    s->words_per_row = words_per_row;
    s->col_tiles = static_cast<int>((n_genomes + BG_TILE_COLS - 1) / BG_TILE_COLS);
    // about four CTAs per SM slot overall; at least 8 rows per warp pass
    long long want_tiles = std::max<long long>(1, 592 / s->col_tiles);
    long long rows = (n_genes + want_tiles - 1) / want_tiles;
    rows = std::max<long long>(rows, 4 * BG_WARPS);
    rows = (rows + BG_WARPS - 1) / BG_WARPS * BG_WARPS;
    s->rows_per_tile = static_cast<int>(std::min<long long>(rows, 1 << 30));
    s->row_tiles = static_cast<int>((n_genes + s->rows_per_tile - 1) / s->rows_per_tile);
    return PGX_OK;
}

}  // namespace

}  // namespace pgx

extern "C" {

size_t pgx_bernoulli_scratch_bytes(int64_t n_genes, int64_t n_genomes)
{
    pgx::BernoulliShape s;
    if (pgx::make_shape(n_genes, n_genomes, (n_genomes + 31) / 32, &s)) return 0;
    const size_t doubles = static_cast<size_t>(s.col_tiles) * n_genes +
                           static_cast<size_t>(s.row_tiles) * n_genomes +
                           static_cast<size_t>(s.row_tiles) * s.col_tiles;
    return doubles * sizeof(double);
}

int pgx_bernoulli_ll_grad(const uint32_t *d_xbits, int64_t words_per_row, int64_t n_genes,
                          int64_t n_genomes, const int32_t *d_row_count,
                          const int32_t *d_col_count, const double *d_p, const double *d_q,
                          double *d_ll, double *d_grad, void *d_scratch, void *stream)
{
    pgx::BernoulliShape s;
    if (int rc = pgx::make_shape(n_genes, n_genomes, words_per_row, &s)) return rc;
    if (!d_xbits || !d_row_count || !d_col_count || !d_p || !d_q || !d_ll || !d_grad || !d_scratch)
        return pgx::fail(PGX_ERR_INVALID, "null pointer passed to pgx_bernoulli_ll_grad");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double *gp_part = static_cast<double *>(d_scratch);
    double *gq_part = gp_part + static_cast<size_t>(s.col_tiles) * n_genes;
    double *ll_part = gq_part + static_cast<size_t>(s.row_tiles) * n_genomes;
    dim3 grid(s.row_tiles, s.col_tiles);
    pgx::grid_kernel<<<grid, pgx::BG_THREADS, 0, st>>>(d_xbits, s, d_row_count, d_p, d_q, gp_part, gq_part, ll_part);
    PGX_LAUNCH_CHECK("bernoulli grid_kernel");
    const long long total = n_genes + n_genomes;
    pgx::finish_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(
        s, d_row_count, d_col_count, d_p, d_q, gp_part, gq_part, ll_part, d_ll, d_grad);
    PGX_LAUNCH_CHECK("bernoulli finish_kernel");
    return PGX_OK;
}

}  // extern "C"

/* pgx.h -- C ABI of libpgx_b200.so: the pan/core rarefaction + Bernoulli-grid hot path of
 * pangenomix on NVIDIA B200 (sm_100a).
 *
 * The reference (AnnaLew/pangenomix) is pure Python and has no FFI of its own; the
 * functions declared here are what a ctypes binding inside the reference's
 * pangenome_analysis.py would call instead of its Python loops:
 *
 *   pgx_pan_core_curves[_host]   replaces the double loop of estimate_pan_core_size,
 *                                /root/reference/pangenomix/pangenome_analysis.py:81-90
 *   pgx_bernoulli_ll_grad        replaces __bernoulli_grid_loglikelihood__ (:244-249) and
 *                                __bernoulli_grid_loglikelihood_gradient__ (:257-266) as
 *                                called from the L-BFGS-B lambdas at :156-160
 *   pgx_legacy_shuffles          replaces the np.arange + np.random.shuffle pair at :84-85
 *                                (numpy legacy MT19937 stream, bit-exact)
 *   pgx_heaps_fit                batched counterpart of __fit_heaps_single__ (:39-48)
 *   pgx_ks_montecarlo[_host]     replaces the simulation loop of ks_montecarlo_bbn (:471-480) incl. the
 *                                np.random.choice draws of draw_bbn (:484-492), bit-exact
 *   pgx_coo_marginals[_host],    replace ``gene_mat.sum(axis=1)`` + collections.Counter at :352-355
 *   pgx_frequency_spectrum       (LightSparseDataFrame.sum, sparse_utils.py:284-292;
 *                                count_gene_occurence, core_genome.py:127-155)
 *
 * Conventions: every function returns 0 on success and a non-zero code otherwise;
 * pgx_last_error() then returns a thread-local message.  No C++ types, exceptions or
 * torch types cross this boundary.  Pointers prefixed d_ are DEVICE pointers owned by
 * the caller (the library never frees caller memory); h_ are HOST pointers.  Functions
 * taking a ``stream`` (a cudaStream_t passed as void*) are asynchronous on that stream;
 * the *_host variants are synchronous and move the data themselves.
 */
#ifndef PGX_H_
#define PGX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGX_VERSION 320

enum {
    PGX_OK = 0,
    PGX_ERR_INVALID = 1,   /* bad argument (null pointer, size out of range, ...) */
    PGX_ERR_CUDA = 2,      /* a CUDA runtime call or kernel launch failed */
    PGX_ERR_UNSUPPORTED = 3
};

/* Device-resident plan of a BINARY gene x genome presence/absence table, built once per
 * matrix by pgx_plan_create (below; pangenomix_b200/plan.py is its numpy specification) from
 * ``df_genes.data`` -- the same object estimate_pan_core_size reads at pangenome_analysis.py:74.
 * Genes fall into three groups:
 *
 *  closed forms  Empty, universal, single-genome and single-absence genes never reach a row
 *                kernel: they are functions of perm[0] and of one rank and live in the
 *                per-genome vectors d_colsum / d_w_present / d_w_absent.
 *  list rows     Every other gene whose SHORTER list (present genomes or absent genomes) has
 *                fewer than the plan's long threshold entries stores that list as uint16
 *                genome indices in 16-byte chunks.  A sub-block is up to 32 rows with the
 *                same chunk count and list kind, one lane per row; lane l reads chunk
 *                first_chunk + it * 32 + l at iteration it; a warp task streams a run of
 *                consecutive sub-blocks.  Unused slots hold sentinel
 *                indices in [N, N + 32) whose rank-table entry is 0xffff.  Entries are
 *                ordered inside a row so that neighbouring lanes gather from different
 *                shared-memory banks.  d_sorted_* hold the same lists plainly sorted.
 *  bitmap rows   The remaining (long) genes, as a genome-major bit-sliced bitmap: rows are
 *                grouped in superblocks of 1,024 W (W = slice_words = 1, 2 or 4);
 *                d_bits[(sb * N + c) * 32 W + w] holds, in bit b, the presence of row
 *                sb * 1024 W + 32 w + b in genome c (one line of 128 W bytes per (superblock,
 *                genome); lane l of the probing warp owns words l W .. l W + W - 1).  Rows of
 *                similar density share a superblock.
 */
typedef struct pgx_plan {
    const uint16_t *d_chunks;      /* [n_chunks * 8] list rows, 16-byte aligned */
    const int32_t *d_tasks;        /* [n_tasks * 4] {first_chunk, chunks_per_row | absent_list << 24, first_row,
                                      rows}: a run of ceil(rows / 32) consecutive sub-blocks; 16-byte aligned,
                                      costly tasks first */
    const uint16_t *d_sorted_idx;  /* sorted copy of the list rows */
    const int32_t *d_sorted_ptr;   /* [n_rows + 1] offsets into d_sorted_idx */
    const uint32_t *d_bits;        /* [n_superblocks * n_genomes * 32 W] bit-sliced bitmap rows, 16-byte aligned */
    const void *reserved_ptr;
    const int32_t *d_colsum;       /* [n_genomes] #genes present in genome c (all genes of the table) */
    const int32_t *d_w_present;    /* [n_genomes] #genes present ONLY in genome c */
    const int32_t *d_w_absent;     /* [n_genomes] #genes absent ONLY from genome c */
    int64_t n_chunks;
    int32_t n_genomes;             /* N  (1 <= N <= 65503) */
    int32_t n_genes;               /* G  (all genes of the table, incl. closed-form ones) */
    int32_t n_rows;                /* list rows */
    int32_t n_tasks;
    int32_t n_long;                /* bitmap rows */
    int32_t n_superblocks;         /* ceil(n_long / (1024 * slice_words)) */
    int32_t perms_per_cta;         /* 8, 4, 2 or 1: rank tables that share a CTA's shared memory */
    int32_t slice_words;           /* W = 1, 2 or 4: words per lane of a (superblock, genome) line */
    int32_t max_colsum;            /* max over genomes of d_colsum; 1 .. 65535 lets the kernels count in uint16 bins and
                                      the host-buffer calls ship uint16 curve steps (0 = unknown: int32 everywhere) */
    int32_t reserved_i32;
} pgx_plan;

/* Host image of a plan: what pgx_host_plan_create makes of the COO table and pgx_plan_upload copies to the device.
 * All arrays are owned by the image (pgx_host_plan_destroy frees them); sizes as in struct pgx_plan.  The row_* /
 * long_gene arrays say which gene became which list / bitmap row (diagnostics and tests). */
typedef struct pgx_host_plan {
    void *owner;
    const uint16_t *chunks;        /* [n_chunks * 8] */
    const int32_t *tasks;          /* [n_tasks * 4] */
    const uint16_t *sorted_idx;    /* [n_sorted] */
    const int32_t *sorted_ptr;     /* [n_rows + 1] */
    const uint32_t *bits;          /* [n_bits_words] */
    const int32_t *colsum, *w_present, *w_absent;   /* [n_genomes] each */
    const int64_t *row_gene;       /* [n_rows] gene of every list row */
    const int32_t *row_len;        /* [n_rows] folded list length */
    const uint8_t *row_absent;     /* [n_rows] 1 when the list holds the ABSENT genomes */
    const int64_t *long_gene;      /* [n_long] gene of every bitmap row */
    int64_t nnz, nnz_list, nnz_long, n_chunks, n_bits_words, n_sorted;
    int32_t n_genomes, n_genes, n_rows, n_tasks, n_long, n_superblocks, perms_per_cta, slice_words;
    int32_t long_threshold, max_colsum, n_empty, n_full;
} pgx_host_plan;

/* The planner (host only, threaded; no GPU work): COO entries of a BINARY gene x genome table -- row = gene,
 * col = genome, every entry a presence, as every producer of the reference writes ``df_genes.data``
 * (pangenome.py:631-650; read back by read_lsdf, sparse_utils.py:35) -- to the host image of a plan.
 *   long_threshold : folded list length from which a gene is served from the bitmap (< 0: the default,
 *                    about 1.28 sqrt(N); 0: list rows only)
 *   perms_per_cta  : 8, 4, 2, 1 or 0 = as many rank tables as fit one CTA's shared memory
 *   slice_words    : 1, 2, 4 or 0 = choose by the number of bitmap rows
 * Duplicate (gene, genome) entries -- which scipy's tocsr() at pangenome_analysis.py:75 would sum to a non-binary
 * table -- are rejected with PGX_ERR_INVALID. */
int pgx_host_plan_create(const int32_t *row, const int32_t *col, int64_t nnz, int32_t n_genes, int32_t n_genomes,
                         int32_t long_threshold, int32_t perms_per_cta, int32_t slice_words, pgx_host_plan **out);
void pgx_host_plan_destroy(pgx_host_plan *plan);

/* Device plans owned by the library, on the calling thread's current device.  pgx_plan_create = pgx_host_plan_create
 * with the defaults + pgx_plan_upload; the returned plan is what every pgx_pan_core_* call takes and lives until
 * pgx_plan_destroy.  (A caller may instead fill a struct pgx_plan with device pointers of its own, as the Python
 * engine can with torch tensors: the calls never look behind the struct.) */
int pgx_plan_create(const int32_t *row, const int32_t *col, int64_t nnz, int32_t n_genes, int32_t n_genomes,
                    int32_t long_threshold, pgx_plan **out);
int pgx_plan_upload(const pgx_host_plan *host, pgx_plan **out);
int pgx_plan_destroy(pgx_plan *plan);

int pgx_version(void);
const char *pgx_last_error(void);

/* SM count, opt-in shared memory per block and L2 size of the current device. */
int pgx_device_info(int32_t *sm_count, int32_t *smem_optin_bytes, int64_t *l2_bytes);

/* Pan/core curves for n_perm genome orders (pangenome_analysis.py:81-90).
 *   d_perms  : [n_perm][N] uint16, row i = shuffle_indices of iteration i (:84-85).  TRUSTED input: every row
 *              must be a permutation of 0 .. N-1 (entries >= N are ignored, a genome that does not occur counts
 *              as never sampled; the host-buffer calls below report such rows as PGX_ERR_INVALID)
 *   d_curves : [n_perm][2N] int32; columns 0..N-1 = pan_genomes[i,:], N..2N-1 =
 *              core_genomes[i,:] (the np.hstack layout of :97).  Overwritten; also the kernels' workspace
 *              (rank rows and histogram rows live in it until the final scan), 16-byte aligned.
 * Asynchronous on ``stream``. */
int pgx_pan_core_curves(const pgx_plan *plan, const uint16_t *d_perms, int64_t n_perm,
                        int32_t *d_curves, void *stream);

/* Same, float64 output exactly as the reference's DataFrame values (:76-77, :97).
 * d_hist is caller scratch of n_perm * 2N int32. */
int pgx_pan_core_curves_f64(const pgx_plan *plan, const uint16_t *d_perms, int64_t n_perm,
                            int32_t *d_hist, double *d_curves, void *stream);

/* Host-buffer entry point: h_perms [n_perm][N] uint16 and h_curves [n_perm][2N]
 * (int32 when out_f64 == 0, float64 otherwise) live in host memory (pinned h_perms make
 * the uploads asynchronous).  The plan stays device-resident.  The call pipelines
 * H2D -> kernels -> D2H -> result over three internal slots in blocks of ``perms_per_block``
 * permutations (0 = choose) and returns when h_curves is complete.  When plan->max_colsum <= 65535 the
 * device ships the curves' STEPS as uint16 (half the bytes of int32 curves, a quarter of float64) -- on tables of
 * 2,048 genomes or more as uint16 heads + uint8 tails (pgx_expand_split; PGX_SPLIT_HEAD=0 switches that off) -- into pinned
 * staging owned by the library and host threads rebuild the curves straight into h_curves (PGX_COPY_THREADS).
 * Rows of h_perms that are not permutations of 0 .. N-1 make the call fail with PGX_ERR_INVALID.
 * One host-buffer call (this one or pgx_estimate_pan_core) at a time per process. */
int pgx_pan_core_curves_host(const pgx_plan *plan, const uint16_t *h_perms, int64_t n_perm,
                             void *h_curves, int32_t out_f64, int64_t perms_per_block);

/* estimate_pan_core_size in one call (pangenome_analysis.py:76-90): draws n_iter consecutive
 * ``np.arange(N); np.random.shuffle`` permutations from the numpy-legacy MT19937 state
 * (mt_key[624] / mt_pos as np.random.get_state() reports them; advanced in place exactly as the
 * reference's loop advances the global stream), rarefies them on the current device and writes
 * the float64 [n_iter][2N] table the reference builds with np.hstack (:97) into h_curves
 * (ordinary host memory).  RNG, H2D, kernels, D2H and the final copy are pipelined over three
 * internal slots; staging is cached for the life of the library.  One call at a time per process. */
int pgx_estimate_pan_core(const pgx_plan *plan, uint32_t *mt_key, int32_t *mt_pos, int64_t n_iter,
                          double *h_curves, int64_t perms_per_block);

/* ``head`` of the split row format the host-buffer calls currently use for the table size they last served
 * (pgx_expand_split; 0 = plain uint16 rows): 2N + 2 head bytes per permutation cross PCIe instead of 4N. */
int pgx_split_head(void);

/* Launch-shape overrides for experiments (0 = heuristic): permutations per CTA of the list
 * kernel (1, 2, 4 or 8), row splits per permutation batch and threads per CTA. */
int pgx_set_tuning(int32_t perms_per_cta, int32_t row_splits, int32_t threads_per_cta);

/* Per-kernel timing for roofline reports: while enabled, every pgx_pan_core_curves[_f64]
 * call brackets its kernels with CUDA events on the caller's stream and runs the two row kernels
 * back to back.  pgx_profile_read waits for them, returns the summed durations (ms) of the list
 * kernel, the bitmap-probe kernel and everything else (prep + scan kernels) and the number of
 * calls since the last read, and releases the events. */
int pgx_profile_enable(int32_t on);
int pgx_profile_read(double *list_ms, double *probe_ms, double *scan_ms, int64_t *calls);

/* Evidence of how the two row kernels share the SMs: while a device buffer of 1 + 3 * capacity_records uint64 is set
 * (zero-initialised by the caller; null switches the trace off), every CTA of the list kernel and every warp of the
 * probe kernel appends {kind (1 list, 2 probe) | smid << 8, start, end} in %globaltimer nanoseconds; word 0 counts
 * the records.  scripts/overlap_trace.py turns it into profiles/r02/overlap_timeline_<name>.json. */
int pgx_set_trace(void *d_trace, int64_t capacity_records);

/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
int64_t pgx_launch_count(void);

/* Bernoulli grid: log-likelihood and gradient at (P, Q) (pangenome_analysis.py:244-266).
 *   d_xbits     : [G][words_per_row] uint32, bit (j & 31) of word (j >> 5) = X[i][j]
 *   d_row_count : [G] int32  = X.sum(axis=1);   d_col_count : [N] int32 = X.sum(axis=0)
 *   d_p [G], d_q [N] float64;  d_ll [1], d_grad [G + N] float64 (dL/dp then dL/dq)
 *   d_scratch   : pgx_bernoulli_scratch_bytes(G, N) bytes
 * Deterministic (fixed reduction order).  Asynchronous on ``stream``. */
size_t pgx_bernoulli_scratch_bytes(int64_t n_genes, int64_t n_genomes);
int pgx_bernoulli_ll_grad(const uint32_t *d_xbits, int64_t words_per_row, int64_t n_genes,
                          int64_t n_genomes, const int32_t *d_row_count,
                          const int32_t *d_col_count, const double *d_p, const double *d_q,
                          double *d_ll, double *d_grad, void *d_scratch, void *stream);

/* Batched Heaps-law fits y = kappa * x^alpha, x = 1 .. n_points, one per curve: the GPU counterpart of
 * fit_heaps_by_iteration / __fit_heaps_single__ (pangenome_analysis.py:24-48; scipy curve_fit, start point
 * alpha = 0.5, kappa = min(y)).  Levenberg-Marquardt with the analytic Jacobian, run to fp64 convergence.
 *   d_curves : n_curves rows of ``stride`` elements, int32 (is_f64 == 0) or float64; the first n_points
 *              of every row are fitted (the Pan half of a pgx_pan_core_curves row: stride = 2N)
 *   d_fit    : [n_curves][2] float64 {alpha, kappa}
 *   d_info   : optional [n_curves] int32, LM trials used (negative: stopped without converging)
 *   d_scratch: pgx_heaps_scratch_bytes(n_points) bytes
 * Deterministic.  Asynchronous on ``stream``. */
size_t pgx_heaps_scratch_bytes(int64_t n_points);
int pgx_heaps_fit(const void *d_curves, int32_t is_f64, int64_t n_curves, int64_t n_points, int64_t stride,
                  double *d_fit, int32_t *d_info, void *d_scratch, void *stream);

/* Host-side helper of the planner (no GPU work): orders the entries of the list rows so that the lanes
 * of a shared-memory wavefront of the list kernel hit distinct banks (pangenomix_b200/plan.py,
 * _bank_ordered_chunks; the layout is the one struct pgx_plan documents).
 *   flat / ptr        : sorted folded lists of the n_rows list rows, concatenated (int32) / offsets (int64)
 *   block_*           : per sub-block of 32 rows: first chunk, chunks per row, first row, rows (int64 [n_blocks])
 *   modulus           : lanes per 128-byte wavefront: 8, 16 or 32 (= 64 / perms_per_cta, capped)
 *   colour_max_chunks : rows of more chunks use the cheap positional order instead of the edge colouring
 *   chunks            : out, uint16 [sum(block_nch) * 32 * 8]
 *   n_threads         : host threads (0 = choose) */
int pgx_plan_bank_order(const int32_t *flat, const int64_t *ptr, int64_t n_rows,
                        const int64_t *block_first, const int64_t *block_nch,
                        const int64_t *block_first_row, const int64_t *block_rows, int64_t n_blocks,
                        int32_t n_genomes, int32_t modulus, int32_t colour_max_chunks,
                        uint16_t *chunks, int32_t n_threads);

/* Host-side helper of the planner: the bit-sliced bitmap d_bits of struct pgx_plan for the long rows
 * ``long_gene`` (gene ids, in superblock order) of a gene-major CSR (indptr int64, indices int32).
 * ``bits`` (uint32 [n_superblocks * n_genomes * 32 * slice_words]) must be zero-initialised. */
int pgx_plan_build_bitmap(const int64_t *indptr, const int32_t *indices, const int64_t *long_gene,
                          int64_t n_long, int32_t n_genomes, int32_t slice_words, uint32_t *bits,
                          int32_t n_threads);

/* Host-side helpers of the planner for the table ingest (no GPU work; threaded; n_threads 0 = choose).
 * They do what pangenome_analysis.py:74-75 asks of scipy (``df_genes.data`` COO -> compressed rows), gene-major.
 *
 * pgx_plan_coo_to_csr: COO entries (row = gene, col = genome; every stored value 1 -- the caller checks) ->
 *   canonical CSR: indptr int64 [n_genes + 1], indices int32 [nnz] ascending inside a row, colsum int32
 *   [n_genomes].  Duplicate (gene, genome) pairs are kept and counted in *n_duplicates (scipy's tocsr would
 *   sum them to 2, a non-binary table: the caller rejects it).
 * pgx_plan_folded_lists: for list row r (gene genes[r]) the ascending list of its present genomes
 *   (use_abs[r] == 0) or absent genomes (use_abs[r] != 0) into flat[ptr[r] .. ptr[r + 1]).
 * pgx_plan_missing_genome: missing[r] = the one genome gene genes[r] (present in N - 1 genomes) lacks. */
int pgx_plan_coo_to_csr(const int32_t *row, const int32_t *col, int64_t nnz, int32_t n_genes, int32_t n_genomes,
                        int64_t *indptr, int32_t *indices, int32_t *colsum, int64_t *n_duplicates,
                        int32_t n_threads);
int pgx_plan_folded_lists(const int64_t *indptr, const int32_t *indices, const int64_t *genes,
                          const uint8_t *use_abs, const int64_t *ptr, int64_t n_rows, int32_t n_genomes,
                          int32_t *flat, int32_t n_threads);
int pgx_plan_missing_genome(const int64_t *indptr, const int32_t *indices, const int64_t *genes, int64_t n_rows,
                            int32_t n_genomes, int32_t *missing, int32_t n_threads);
/* Host-side helper of the planner: the order of the list rows inside their classes (rows of one chunk count and
 * list kind, consecutive in genes / use_abs / class_key) that balances, per group of ``modulus`` consecutive rows,
 * how many entries fall on each shared-memory bank residue (genome index modulo ``modulus``): such a group needs at
 * least max-over-residues(entries of the residue) gather steps however pgx_plan_bank_order arranges them.
 * order[i] (int64 [n_rows]) = current position of the row that takes place i.  Deterministic; ``window`` = candidates
 * looked at per choice (pangenomix_b200/plan.py, _balanced_row_order, is the numpy specification). */
int pgx_plan_balance_rows(const int64_t *indptr, const int32_t *indices, const int64_t *genes,
                          const uint8_t *use_abs, const int64_t *class_key, int64_t n_rows, int32_t n_genomes,
                          int32_t modulus, int32_t window, int64_t *order, int32_t n_threads);
/* Returns 1 when every 64-bit word of words[0 .. n) equals ``value``, else 0: the planner's "every stored value
 * is 1" test for the int64 (value 1) and float64 (value = bits of 1.0) ``data`` arrays of a table. */
int pgx_plan_all_equal_u64(const uint64_t *words, int64_t n, uint64_t value, int32_t n_threads);

/* LSDF ingest (host): raw DEFLATE (RFC 1951) of one zip member straight into the caller's buffer.
 * Replaces the zlib inflate inside scipy.sparse.load_npz, which read_lsdf calls
 * (/root/reference/pangenomix/sparse_utils.py:35; the archive is written deflated at :314).
 * dst_len must be the exact inflated size (the zip entry's file_size).  Returns PGX_ERR_INVALID for
 * anything that is not a complete, well-formed stream of that size; the caller verifies the
 * entry's CRC-32 and falls back to zlib on any error.  Thread-safe (no shared state). */
int pgx_inflate_raw(const uint8_t *src, int64_t src_len, uint8_t *dst, int64_t dst_len);

/* Host half of the compact transfer (no GPU work): rows of 2N uint16 curve steps
 *   d[0] = pan[0], d[k] = pan[k] - pan[k-1];  d[N] = core[0], d[N+k] = core[k-1] - core[k]
 * -> curves [n_rows][2N], int32 (out_f64 == 0) or float64.  n_threads 0 = choose. */
int pgx_expand_deltas(const uint16_t *h_deltas, int64_t n_rows, int32_t n_genomes, void *h_curves,
                      int32_t out_f64, int32_t n_threads);

/* The same for the SPLIT row format the host-buffer calls use on tables of many genomes: a step above 255 only
 * occurs while the first genomes are added, so the first ``head`` steps of each curve travel as uint16, the rest as
 * uint8 -- per row [pan head: head x u16][core head: head x u16][pan tail: (N - head) x u8][core tail: (N - head) x u8],
 * 2N + 2 head bytes instead of 4N (a block with a larger step in a tail is sent again as uint16 and ``head`` doubled
 * for the blocks after it).  1 <= head <= N. */
int pgx_expand_split(const uint8_t *h_rows, int64_t n_rows, int32_t n_genomes, int32_t head, void *h_curves,
                     int32_t out_f64, int32_t n_threads);

/* numpy legacy RandomState stream (host): ``count`` consecutive
 * ``a = np.arange(n); np.random.shuffle(a)`` results as uint16 rows, continuing from the
 * MT19937 state in ``mt_key`` (624 words) / ``mt_pos`` exactly as
 * np.random.get_state() reports them; the advanced state is written back so the caller
 * can np.random.set_state() it.  Requires n <= 65535. */
int pgx_legacy_shuffles(uint32_t *mt_key, int32_t *mt_pos, int64_t n, int64_t count,
                        uint16_t *h_perms);

/* Raw 32-bit outputs of the same stream (host): ``count`` words continuing the MT19937 state, written back
 * advanced.  numpy's legacy random_sample() consumes two of them per double, (a >> 5, b >> 6) -> (a * 2^26 + b) / 2^53;
 * RandomState.choice(p=...) -- draw_bbn, pangenome_analysis.py:484-492 -- one such double per draw. */
int pgx_legacy_random_raw(uint32_t *mt_key, int32_t *mt_pos, int64_t count, uint32_t *h_out);

/* Marginals of a COO presence/absence table on the device: d_row_sum[g] = entries of gene g (what
 * ``gene_mat.sum(axis=1)`` gives for a binary table without duplicates, pangenome_analysis.py:354-355; also
 * LightSparseDataFrame.sum, sparse_utils.py:284-292, and count_gene_occurence, core_genome.py:127-155),
 * d_col_sum[c] = entries of genome c.  Entries outside the table are skipped and counted in *d_bad.
 *   accumulate == 0 : the three outputs are zeroed first;  != 0 : counts are added (tables sent in chunks)
 * Asynchronous on ``stream``. */
int pgx_coo_marginals(const int32_t *d_row, const int32_t *d_col, int64_t nnz, int32_t n_genes, int32_t n_genomes,
                      int32_t *d_row_sum, int32_t *d_col_sum, int32_t *d_bad, int32_t accumulate, void *stream);

/* Gene-frequency spectrum of the row sums: d_spectrum[m] (int64 [N + 1]) = genes present in exactly m genomes,
 * d_first_gene[m] (int32 [N + 1]) = the lowest gene index among them (INT32_MAX when there is none): the order of
 * first appearance is the order of the collections.Counter the reference slices at pangenome_analysis.py:355, :364.
 * Row sums above N (duplicate entries) are counted at N.  Asynchronous on ``stream``. */
int pgx_frequency_spectrum(const int32_t *d_row_sum, int64_t n_genes, int32_t n_genomes, int64_t *d_spectrum,
                           int32_t *d_first_gene, void *stream);

/* Both of the above for a table in host memory (row / col as read_lsdf returns them in ``.data``); synchronous.
 * h_spectrum / h_first_gene may be null.  Entries outside the table make the call fail with PGX_ERR_INVALID. */
int pgx_coo_marginals_host(const int32_t *h_row, const int32_t *h_col, int64_t nnz, int32_t n_genes, int32_t n_genomes,
                           int32_t *h_row_sum, int32_t *h_col_sum, int64_t *h_spectrum, int32_t *h_first_gene);

/* Monte-Carlo Kolmogorov-Smirnov statistics of ks_montecarlo_bbn (pangenome_analysis.py:457-482): for iteration i,
 * n_samples draws of ``np.random.choice(np.arange(sim_limit), p=probs)`` (draw_bbn, :484-492), their eCDF
 * (ecdf_from_counts, :494-499) and d_ks_sim[i] = max |eCDF - model_cdf| (:479), float64, bit-exact.
 *   d_raw        : [iterations][n_samples][2] uint32, the raw MT19937 words of the draws in stream order
 *                  (pgx_legacy_random_raw), 8-byte aligned
 *   d_choice_cdf : [sim_limit] float64, ``cdf = probs.cumsum(); cdf /= cdf[-1]`` as numpy's choice forms it
 *   d_model_cdf  : [sim_limit] float64, ``np.cumsum(model_pmf)`` of :464-465
 *   d_scratch    : pgx_ks_scratch_bytes(iterations, sim_limit) bytes
 * iterations <= 65,535 per call, n_samples < 2^31.  Deterministic.  Asynchronous on ``stream``. */
size_t pgx_ks_scratch_bytes(int64_t iterations, int32_t sim_limit);
int pgx_ks_montecarlo(const uint32_t *d_raw, int64_t iterations, int64_t n_samples, const double *d_choice_cdf,
                      const double *d_model_cdf, int32_t sim_limit, double *d_ks_sim, void *d_scratch, void *stream);

/* The same with host buffers, drawing from the numpy-legacy MT19937 state (mt_key[624] / mt_pos as
 * np.random.get_state() reports them; advanced in place by the 2 * iterations * n_samples words the reference's
 * np.random.choice call consumes).  The generator (host, serial) fills one block of iterations while the previous one
 * is uploaded and reduced on the current device.  Synchronous; one call at a time per process. */
int pgx_ks_montecarlo_host(uint32_t *mt_key, int32_t *mt_pos, int64_t iterations, int64_t n_samples,
                           const double *h_choice_cdf, const double *h_model_cdf, int32_t sim_limit, double *h_ks_sim);

#ifdef __cplusplus
}
#endif
#endif /* PGX_H_ */

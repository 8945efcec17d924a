/* A host in plain C that links ONLY libpgx_b200.so (no Python, no torch): plans a small table with
 * pgx_host_plan_create, and -- when a GPU is present -- uploads it, rarefies a few genome orders through the
 * host-buffer call and checks the curves against the definition (pangenome_analysis.py:86-90) computed here.
 *   gcc -O2 -I include tests/c/abi_host.c -L pangenomix_b200 -lpgx_b200 -Wl,-rpath,$PWD/pangenomix_b200 -o abi_host
 *   ./abi_host [--gpu]
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pgx.h"

#define G 3000
#define N 257
#define PERMS 21

static uint32_t rng_state = 12345u;
static uint32_t next_u32(void) { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

int main(int argc, char **argv)
{
    const int want_gpu = argc > 1 && !strcmp(argv[1], "--gpu");
    static uint8_t x[G][N];
    int32_t *row = malloc(sizeof(int32_t) * G * N), *col = malloc(sizeof(int32_t) * G * N);
    int64_t nnz = 0;
    for (int g = 0; g < G; ++g) {
        const uint32_t dens = (uint32_t)(g % 11) * 1600000u + (g % 7 == 0 ? 16000000u : 0u);   /* 0 .. ~1 of 2^24 */
        for (int c = 0; c < N; ++c) {
            x[g][c] = next_u32() < dens;
            if (g % 97 == 0) x[g][c] = 1;                 /* universal genes */
            if (g % 89 == 0) x[g][c] = c == g % N;        /* single-genome genes */
            if (x[g][c]) { row[nnz] = g; col[nnz] = c; ++nnz; }
        }
    }
    pgx_host_plan *host = NULL;
    if (pgx_host_plan_create(row, col, nnz, G, N, -1, 0, 0, &host)) { fprintf(stderr, "plan: %s\n", pgx_last_error()); return 1; }
    printf("planned %d genes x %d genomes, nnz %lld: %d list rows in %d tasks, %d bitmap rows, threshold %d, max colsum %d\n",
           host->n_genes, host->n_genomes, (long long)host->nnz, host->n_rows, host->n_tasks, host->n_long,
           host->long_threshold, host->max_colsum);
    if (host->n_rows <= 0 || host->n_long <= 0 || host->n_full <= 0) { fprintf(stderr, "unexpected plan\n"); return 1; }
    /* duplicates are rejected */
    row[nnz] = row[0]; col[nnz] = col[0];
    pgx_host_plan *bad = NULL;
    if (pgx_host_plan_create(row, col, nnz + 1, G, N, -1, 0, 0, &bad) != PGX_ERR_INVALID || bad) { fprintf(stderr, "duplicate accepted\n"); return 1; }
    if (!want_gpu) { pgx_host_plan_destroy(host); printf("host ok\n"); return 0; }

    pgx_plan *plan = NULL;
    if (pgx_plan_upload(host, &plan)) { fprintf(stderr, "upload: %s\n", pgx_last_error()); return 1; }
    pgx_host_plan_destroy(host);
    static uint16_t perms[PERMS][N];
    static int32_t curves[PERMS][2 * N];
    for (int p = 0; p < PERMS; ++p) {
        for (int k = 0; k < N; ++k) perms[p][k] = (uint16_t)k;
        for (int k = N - 1; k > 0; --k) { const int j = (int)(next_u32() % (uint32_t)(k + 1)); const uint16_t t = perms[p][k]; perms[p][k] = perms[p][j]; perms[p][j] = t; }
    }
    if (pgx_pan_core_curves_host(plan, &perms[0][0], PERMS, curves, 0, 8)) { fprintf(stderr, "curves: %s\n", pgx_last_error()); return 1; }
    for (int p = 0; p < PERMS; ++p) {
        static int incidence[G];
        memset(incidence, 0, sizeof(incidence));
        for (int k = 0; k < N; ++k) {
            int pan = 0, core = 0;
            for (int g = 0; g < G; ++g) { incidence[g] += x[g][perms[p][k]]; pan += incidence[g] > 0; core += incidence[g] == k + 1; }
            if (curves[p][k] != pan || curves[p][N + k] != core) {
                fprintf(stderr, "permutation %d step %d: got %d / %d, want %d / %d\n", p, k, curves[p][k], curves[p][N + k], pan, core);
                return 1;
            }
        }
    }
    /* marginals of the table (pangenome_analysis.py:354-355) against the dense copy */
    {
        static int32_t row_sum[G], col_sum[N], first_gene[N + 1];
        static int64_t spectrum[N + 1];
        if (pgx_coo_marginals_host(row, col, nnz, G, N, row_sum, col_sum, spectrum, first_gene)) { fprintf(stderr, "marginals: %s\n", pgx_last_error()); return 1; }
        int64_t want_spectrum[N + 1] = {0};
        for (int g = 0; g < G; ++g) {
            int m = 0;
            for (int c = 0; c < N; ++c) m += x[g][c];
            if (row_sum[g] != m) { fprintf(stderr, "row sum of gene %d: got %d, want %d\n", g, row_sum[g], m); return 1; }
            ++want_spectrum[m];
        }
        for (int c = 0; c < N; ++c) {
            int m = 0;
            for (int g = 0; g < G; ++g) m += x[g][c];
            if (col_sum[c] != m) { fprintf(stderr, "column sum of genome %d: got %d, want %d\n", c, col_sum[c], m); return 1; }
        }
        for (int m = 0; m <= N; ++m) {
            if (spectrum[m] != want_spectrum[m]) { fprintf(stderr, "spectrum[%d]: got %lld, want %lld\n", m, (long long)spectrum[m], (long long)want_spectrum[m]); return 1; }
            if (spectrum[m] && row_sum[first_gene[m]] != m) { fprintf(stderr, "first gene of frequency %d is wrong\n", m); return 1; }
        }
    }
    /* Monte-Carlo KS statistics (pangenome_analysis.py:471-480) against the definition, from the same raw words */
    {
        enum { BINS = 9, SAMPLES = 777, ITER = 13 };
        double cdf[BINS], ks[ITER];
        for (int k = 0; k < BINS; ++k) cdf[k] = 1.0 - 1.0 / (double)(1 << (k + 1));
        cdf[BINS - 1] = 1.0;
        static uint32_t key[624], key2[624], raw[2 * SAMPLES * ITER];
        for (int k = 0; k < 624; ++k) key[k] = key2[k] = next_u32();
        int32_t pos = 17, pos2 = 17;
        if (pgx_ks_montecarlo_host(key, &pos, ITER, SAMPLES, cdf, cdf, BINS, ks)) { fprintf(stderr, "ks: %s\n", pgx_last_error()); return 1; }
        if (pgx_legacy_random_raw(key2, &pos2, 2 * SAMPLES * ITER, raw)) { fprintf(stderr, "raw: %s\n", pgx_last_error()); return 1; }
        if (pos != pos2 || memcmp(key, key2, sizeof(key))) { fprintf(stderr, "ks: the stream did not advance by 2 words per draw\n"); return 1; }
        for (int it = 0; it < ITER; ++it) {
            int hist[BINS] = {0};
            for (int s = 0; s < SAMPLES; ++s) {
                const uint32_t a = raw[2 * (it * SAMPLES + s)] >> 5, b = raw[2 * (it * SAMPLES + s) + 1] >> 6;
                const double u = (a * 67108864.0 + b) / 9007199254740992.0;
                int idx = 0;
                while (idx < BINS && cdf[idx] <= u) ++idx;
                ++hist[idx];
            }
            double worst = 0.0;
            long long cum = 0;
            for (int k = 0; k < BINS; ++k) {
                cum += hist[k];
                const double d = (double)cum / (double)SAMPLES - cdf[k];
                if ((d < 0 ? -d : d) > worst) worst = d < 0 ? -d : d;
            }
            if (ks[it] != worst) { fprintf(stderr, "ks statistic %d: got %.17g, want %.17g\n", it, ks[it], worst); return 1; }
        }
    }
    /* the second entry point: plan and upload in one call */
    pgx_plan *again = NULL;
    if (pgx_plan_create(row, col, nnz, G, N, -1, &again) || pgx_plan_destroy(again)) { fprintf(stderr, "create: %s\n", pgx_last_error()); return 1; }
    if (pgx_plan_destroy(plan)) return 1;
    printf("gpu ok: %d curves, the marginals and %d KS statistics bit-exact\n", PERMS, 13);
    return 0;
}

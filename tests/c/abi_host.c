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
    /* the second entry point: plan and upload in one call */
    pgx_plan *again = NULL;
    if (pgx_plan_create(row, col, nnz, G, N, -1, &again) || pgx_plan_destroy(again)) { fprintf(stderr, "create: %s\n", pgx_last_error()); return 1; }
    if (pgx_plan_destroy(plan)) return 1;
    printf("gpu ok: %d curves bit-exact\n", PERMS);
    return 0;
}

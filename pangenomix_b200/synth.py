"""Deterministic synthetic gene x genome presence/absence matrices.

The reference ships no data, so every parity test and every bench line runs on
Bernoulli-sampled matrices with the U-shaped gene-frequency spectrum of a bacterial
pangenome: ``n_core`` near-universal genes plus a Beta(0.3, b) accessory tail whose
mean is tuned so that each genome carries about ``genes_per_genome`` genes.

The matrices honour the producer invariants of the reference's table builder
(/root/reference/pangenomix/pangenome.py:631-650, :672-675): binary int64 data, no
duplicate COO entries, every gene row has at least one presence.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse

# name -> (n_genes, n_genomes, genes_per_genome, seed, default permutations)
CONFIGS = {
    "kat": (6, 5, 3, 20239, 3),
    "c1": (8_000, 50, 4_500, 20241, 100),
    "c2": (40_000, 400, 4_500, 20242, 1_000),
    "c4": (200_000, 10_000, 4_500, 20244, 10_000),
    "c5": (2_000_000, 50_000, 4_000, 20245, 1_000),
}


def gene_frequencies(n_genes, genes_per_genome, rng):
    """Per-gene Bernoulli rates: core block U(0.98, 1) then a Beta(0.3, b) accessory tail."""
    n_core = min(int(round(0.45 * genes_per_genome)), n_genes)
    f = np.empty(n_genes, dtype=np.float64)
    f[:n_core] = rng.uniform(0.98, 1.0, size=n_core)
    n_acc = n_genes - n_core
    if n_acc > 0:
        mu = (genes_per_genome - 0.99 * n_core) / float(n_acc)
        mu = min(max(mu, 1e-6), 0.999)
        b = 0.3 * (1.0 - mu) / mu
        f[n_core:] = rng.beta(0.3, b, size=n_acc)
    return f


def bernoulli_matrix(n_genes, n_genomes, genes_per_genome=4_500, seed=0, method="auto"):
    """Returns a scipy COO matrix (int64 ones) of shape (n_genes, n_genomes).

    method="columns": one uniform draw per cell, genome column by genome column (the
        section-8d recipe; O(G*N) draws).
    method="binomial": per gene draw m ~ Binomial(N, f) and then m distinct genomes;
        same distribution, O(nnz) work -- the only practical way to build C5.
    """
    rng = np.random.RandomState(seed)
    f = gene_frequencies(n_genes, genes_per_genome, rng)
    if method == "auto":
        method = "columns" if n_genes * n_genomes <= 2_500_000_000 else "binomial"
    if method == "columns":
        rows, cols = [], []
        seen = np.zeros(n_genes, dtype=bool)
        for j in range(n_genomes):
            hit = np.flatnonzero(rng.random_sample(n_genes) < f)
            seen[hit] = True
            rows.append(hit.astype(np.int32))
            cols.append(np.full(hit.size, j, dtype=np.int32))
        empty = np.flatnonzero(~seen)
        if empty.size:
            rows.append(empty.astype(np.int32))
            cols.append(rng.randint(n_genomes, size=empty.size).astype(np.int32))
        row = np.concatenate(rows)
        col = np.concatenate(cols)
    elif method == "binomial":
        m = rng.binomial(n_genomes, f)
        m[m == 0] = 1
        row = np.repeat(np.arange(n_genes, dtype=np.int32), m)
        col = np.empty(row.size, dtype=np.int32)
        indptr = np.concatenate(([0], np.cumsum(m)))
        # sparse rows: rejection-free sampling through a random key sort in blocks of
        # equal m would be faster, but a per-gene choice keeps the stream simple.
        dense_cut = n_genomes // 8
        for g in np.flatnonzero(m > dense_cut):
            col[indptr[g]:indptr[g + 1]] = rng.permutation(n_genomes)[:m[g]]
        sparse_genes = np.flatnonzero(m <= dense_cut)
        # draw with replacement, then repair collisions gene by gene (rare for m << N)
        draws = rng.randint(n_genomes, size=int(m[sparse_genes].sum())).astype(np.int32)
        pos = 0
        for g in sparse_genes:
            k = m[g]
            seg = draws[pos:pos + k]
            pos += k
            if k > 1 and np.unique(seg).size != k:
                seg = rng.choice(n_genomes, size=k, replace=False).astype(np.int32)
            col[indptr[g]:indptr[g + 1]] = seg
    else:
        raise ValueError("unknown method %r" % (method,))
    data = np.ones(row.size, dtype=np.int64)
    return scipy.sparse.coo_matrix((data, (row, col)), shape=(n_genes, n_genomes))


def config_matrix(name, scale=1.0):
    """COO matrix for a named BASELINE.json config ('c1', 'c2', 'c4', 'c5', 'kat')."""
    n_genes, n_genomes, per_genome, seed, _ = CONFIGS[name]
    if scale != 1.0:
        n_genes = max(4, int(n_genes * scale))
        n_genomes = max(4, int(n_genomes * scale))
        per_genome = max(2, min(per_genome, n_genes // 2))
    return bernoulli_matrix(n_genes, n_genomes, per_genome, seed)


def labels_for(n_genes, n_genomes):
    """Row / column labels in the style the reference's builder emits (T_C<g>, genome<j>)."""
    index = ["T_C%d" % g for g in range(n_genes)]
    columns = ["genome%d" % j for j in range(n_genomes)]
    return index, columns


def bernoulli_grid_matrix(n_genes, n_genomes, seed=3):
    """Dense 0/1 float64 table for the Bernoulli-grid config C3.

    X_ij ~ Bernoulli(p_i q_j); p_i = 1 w.p. 0.6 else U(0.8, 1); q_j = 1 - Beta(1, 60).
    Returns (X, p, q).
    """
    rng = np.random.RandomState(seed)
    p = np.where(rng.random_sample(n_genes) < 0.6, 1.0, rng.uniform(0.8, 1.0, size=n_genes))
    q = 1.0 - rng.beta(1.0, 60.0, size=n_genomes)
    x = (rng.random_sample((n_genes, n_genomes)) < np.outer(p, q)).astype(np.float64)
    return x, p, q

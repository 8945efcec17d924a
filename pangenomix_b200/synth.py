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
    method="subset": per gene draw m ~ Binomial(N, f) and then a uniformly random m-subset of the
        genomes (``_subset_rows``, vectorised); same distribution as "columns", O(nnz log nnz) work --
        the practical way to build C5 ("auto" picks it above 2.5e9 cells).
    method="binomial": the per-gene Python loop of round 1 for the same distribution (minutes at C5;
        kept so that round-1 tables can be regenerated).
    """
    rng = np.random.RandomState(seed)
    f = gene_frequencies(n_genes, genes_per_genome, rng)
    if method == "auto":
        method = "columns" if n_genes * n_genomes <= 2_500_000_000 else "subset"
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
    elif method == "subset":
        row, col = _subset_rows(rng, rng.binomial(n_genomes, f), n_genomes)
    else:
        raise ValueError("unknown method %r" % (method,))
    data = np.ones(row.size, dtype=np.int64)
    return scipy.sparse.coo_matrix((data, (row, col)), shape=(n_genes, n_genomes))


def _subset_rows(rng, m, n_genomes):
    """Gene g present in a uniformly random m[g]-subset of the genomes, vectorised (numpy only, O(nnz log nnz)).

    Per gene the SHORTER of the present and the absent list (k = min(m, N - m) genomes) is drawn with
    replacement for all genes at once; rows with a repeated genome keep the first occurrence of every genome
    and redraw the repeats until none is left -- a procedure symmetric under relabelling of the genomes, so
    every k-subset is equally likely.  Rows drawn as absent lists are complemented at the end.  C5
    (2,000,000 x 50,000, nnz 2.0e8) takes well under a minute instead of the per-gene Python loop's minutes.
    """
    n = int(n_genomes)
    m = np.asarray(m, dtype=np.int64).copy()
    m[m == 0] = 1                                    # producer invariant: every row has a presence
    n_genes = m.shape[0]
    absent = m > n - m
    k = np.where(absent, n - m, m)
    ptr = np.concatenate(([0], np.cumsum(k)))
    total = int(ptr[-1])
    gene_of = np.repeat(np.arange(n_genes, dtype=np.int64), k)
    key = gene_of * n + rng.randint(n, size=total)          # sorting the keys sorts every row's genomes
    key.sort()
    while True:
        dup = np.flatnonzero(key[1:] == key[:-1]) + 1
        if dup.size == 0:
            break
        # the rows with a repeat are re-sorted after their repeats were redrawn; all other rows are final
        bad = np.zeros(n_genes, dtype=bool)
        bad[gene_of[dup]] = True
        seg = np.flatnonzero(bad[gene_of])                   # ascending, covers whole rows
        sub = key[seg]
        sub_dup = np.flatnonzero(sub[1:] == sub[:-1]) + 1
        sub[sub_dup] = (sub[sub_dup] // n) * n + rng.randint(n, size=sub_dup.size)
        sub.sort()
        key[seg] = sub
    col = (key % n).astype(np.int32)
    del key
    if not absent.any():
        return gene_of.astype(np.int32), col
    # complement the rows that were drawn as absent lists (few: the near-universal genes), in place in gene order
    new_ptr = np.concatenate(([0], np.cumsum(m)))
    out = np.empty(int(new_ptr[-1]), dtype=np.int32)
    keep = np.flatnonzero(~absent[gene_of])
    out[keep + (new_ptr[:-1] - ptr[:-1])[gene_of[keep]]] = col[keep]
    mask = np.empty(n, dtype=bool)
    for g in np.flatnonzero(absent):
        mask[:] = True
        mask[col[ptr[g]:ptr[g + 1]]] = False
        out[new_ptr[g]:new_ptr[g + 1]] = np.flatnonzero(mask)
    return np.repeat(np.arange(n_genes, dtype=np.int32), m), out


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


def core_miss_spectrum(n_genomes=300, n_core=6000, a=0.6, b=90.0, n_accessory=1500, seed=11):
    """Gene-frequency spectrum {frequency: number of genes} for the beta-binomial core estimate
    (pangenome_analysis.py:295-400): ``n_core`` core genes whose miss counts follow BetaBinomial(n_genomes, a, b) --
    the model the reference fits -- plus ``n_accessory`` accessory genes of uniform frequency.  Returned as the
    arrays (frequencies ascending, counts > 0) a ``df_counts`` Series is made of."""
    rng = np.random.RandomState(seed)
    misses = rng.binomial(n_genomes, rng.beta(a, b, size=n_core))
    freq = np.concatenate((n_genomes - misses, rng.randint(1, n_genomes + 1, size=n_accessory)))
    freq = freq[freq >= 1]
    counts = np.bincount(freq, minlength=n_genomes + 1)
    keep = np.flatnonzero(counts)
    return keep.astype(np.int64), counts[keep].astype(np.int64)


def table_with_frequencies(freq, n_genomes, seed=0):
    """Binary gene x genome COO table whose gene g is present in exactly ``freq[g]`` genomes chosen at random (rows
    in the order given)."""
    import scipy.sparse
    rng = np.random.RandomState(seed)
    freq = np.asarray(freq, dtype=np.int64)
    rows, cols = [], []
    for g, m in enumerate(freq):
        cols.append(np.sort(rng.permutation(n_genomes)[:m]))
        rows.append(np.full(m, g, dtype=np.int64))
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    return scipy.sparse.coo_matrix((np.ones(rows.shape[0], dtype=np.int64), (rows, cols)), shape=(freq.shape[0], n_genomes))

"""TEST INFRASTRUCTURE -- numpy restatement of the reference hot path (see oracle/__init__.py).

Every function cites the lines of /root/reference/pangenomix it follows.  Two
restatements of the rarefaction are kept on purpose:

* ``*_direct``  : the reference's own algorithm (running per-gene incidence vector,
  two dense compare+count passes per added genome), O(iter * N * G).  This is the one
  that is pinned against the golden fixtures and timed as the CPU baseline.
* ``*_minrank`` : the O(nnz) identity the CUDA path is built on (first-presence /
  first-absence rank -> histogram -> prefix sum).  Checked against ``*_direct`` and
  the fixtures so that the identity itself is under test, not assumed.
"""
from __future__ import annotations

import numpy as np
import pandas as pd
import scipy.optimize
import scipy.sparse


# --------------------------------------------------------------------------------------
# numpy legacy shuffle (third party; call site pangenome_analysis.py:84-85)
# --------------------------------------------------------------------------------------
def legacy_shuffle(n, next_uint32):
    """``np.random.shuffle(np.arange(n))`` of the legacy RandomState, restated.

    Fisher-Yates from the top; index j in [0, i] by masked rejection on successive
    32-bit MT19937 outputs (numpy/random/_legacy: ``random_interval``).  ``next_uint32``
    is a zero-argument callable yielding raw 32-bit outputs.
    """
    a = np.arange(n, dtype=np.int64)
    for i in range(n - 1, 0, -1):
        mask = i
        mask |= mask >> 1
        mask |= mask >> 2
        mask |= mask >> 4
        mask |= mask >> 8
        mask |= mask >> 16
        while True:
            v = next_uint32() & mask
            if v <= i:
                break
        a[i], a[v] = a[v], a[i]
    return a


def legacy_shuffle_stream(seed, n, count):
    """``count`` consecutive shuffles of arange(n) as ``np.random.seed(seed)`` would give.

    Drives ``legacy_shuffle`` from numpy's MT19937 bit generator seeded the legacy way
    so the emulation can be compared with ``np.random.RandomState(seed).shuffle``.
    """
    legacy = np.random.RandomState(seed)
    bitgen = np.random.MT19937()
    bitgen.state = legacy.get_state(legacy=False)
    buf = {"vals": np.empty(0, dtype=np.uint64), "pos": 0}

    def next_uint32():
        if buf["pos"] >= buf["vals"].size:
            buf["vals"] = bitgen.random_raw(4096)
            buf["pos"] = 0
        v = int(buf["vals"][buf["pos"]])
        buf["pos"] += 1
        return v

    return np.stack([legacy_shuffle(n, next_uint32) for _ in range(count)])


# --------------------------------------------------------------------------------------
# rarefaction (pangenome_analysis.py:51-98)
# --------------------------------------------------------------------------------------
def _genome_major_csr(data):
    """pangenome_analysis.py:74-75 -- ``df_genes.data.T.tocsr()`` (duplicates are summed)."""
    return scipy.sparse.coo_matrix(data).T.tocsr()


def pan_core_curves_direct(data, perms):
    """pangenome_analysis.py:81-90 for a given table of genome orders.

    data  : scipy sparse (or dense array) gene x genome table
    perms : (n_iter, N) integer array, row i = ``shuffle_indices`` of iteration i
    Returns (pan, core) float64 arrays of shape (n_iter, N), exactly as :76-77/:89-90.
    """
    gene_data = _genome_major_csr(data)
    num_strains, num_genes = gene_data.shape
    perms = np.asarray(perms)
    num_iter = perms.shape[0]
    pan = np.zeros((num_iter, num_strains))
    core = np.zeros((num_iter, num_strains))
    indptr, indices, values = gene_data.indptr, gene_data.indices, gene_data.data
    for i in range(num_iter):
        gene_incidence = np.zeros(num_genes, dtype="int")
        for j, shuffle_col in enumerate(perms[i]):
            lo, hi = indptr[shuffle_col], indptr[shuffle_col + 1]
            # :88  gene_incidence += gene_data[shuffle_col,:]
            np.add.at(gene_incidence, indices[lo:hi], values[lo:hi])
            pan[i, j] = (gene_incidence > 0).sum()          # :89
            core[i, j] = (gene_incidence == j + 1).sum()    # :90
    return pan, core


def pan_core_curves_minrank(data, perms):
    """The O(nnz) identity of SURVEY.md appendix A.1, for a BINARY table.

    rank = inverse permutation; fp_g = min rank over present genomes (N if none);
    fa_g = min rank over absent genomes (N if none) = mex of the present ranks;
    pan[k] = #{g : fp_g <= k}; core[k] = #{g : fa_g > k}.
    """
    csr = scipy.sparse.coo_matrix(data).tocsr()
    csr.sum_duplicates()
    if csr.nnz and not np.all(csr.data == 1):
        raise ValueError("min-rank identity needs a binary table without duplicates")
    n_genes, n = csr.shape
    perms = np.asarray(perms)
    num_iter = perms.shape[0]
    counts = np.diff(csr.indptr)
    row_of = np.repeat(np.arange(n_genes), counts)
    pos_in_row = np.arange(csr.nnz) - np.repeat(csr.indptr[:-1], counts)
    nonempty = counts > 0
    starts = csr.indptr[:-1][nonempty]
    pan = np.zeros((num_iter, n))
    core = np.zeros((num_iter, n))
    for i in range(num_iter):
        rank = np.empty(n, dtype=np.int64)
        rank[perms[i]] = np.arange(n)
        r = rank[csr.indices]
        fp = np.full(n_genes, n, dtype=np.int64)
        fa = np.zeros(n_genes, dtype=np.int64)
        if csr.nnz:
            fp[nonempty] = np.minimum.reduceat(r, starts)
            order = np.lexsort((r, row_of))              # ranks ascending inside each row
            mismatch = np.where(r[order] != pos_in_row, pos_in_row, n + 1)
            first_gap = np.minimum.reduceat(mismatch, starts)
            fa[nonempty] = np.where(first_gap > n, counts[nonempty], first_gap)
        pan[i] = np.cumsum(np.bincount(fp, minlength=n + 1)[:n])
        core[i] = n_genes - np.cumsum(np.bincount(fa, minlength=n + 1)[:n])
    return pan, core


def _draw_perms(num_strains, num_iter):
    """pangenome_analysis.py:84-85 -- one arange + global np.random.shuffle per iteration."""
    perms = np.empty((num_iter, num_strains), dtype=np.int64)
    for i in range(num_iter):
        shuffle_indices = np.arange(num_strains)
        np.random.shuffle(shuffle_indices)
        perms[i] = shuffle_indices
    return perms


def _curve_frame(pan, core):
    """pangenome_analysis.py:93-97 -- labels and hstack."""
    num_iter, num_strains = pan.shape
    iter_index = ["Iter" + str(x) for x in range(1, num_iter + 1)]
    pan_cols = ["Pan" + str(x) for x in range(1, num_strains + 1)]
    core_cols = ["Core" + str(x) for x in range(1, num_strains + 1)]
    return pd.DataFrame(index=iter_index, columns=pan_cols + core_cols,
                        data=np.hstack([pan, core]))


def estimate_pan_core_size_direct(df_genes, num_iter):
    """pangenome_analysis.py:51-98 end to end (global RNG stream, labels, float64)."""
    _, num_strains = df_genes.shape
    perms = _draw_perms(num_strains, num_iter)
    return _curve_frame(*pan_core_curves_direct(df_genes.data, perms))


def estimate_pan_core_size_minrank(df_genes, num_iter):
    _, num_strains = df_genes.shape
    perms = _draw_perms(num_strains, num_iter)
    return _curve_frame(*pan_core_curves_minrank(df_genes.data, perms))


# --------------------------------------------------------------------------------------
# mean curve + Heaps fit (plot.py:8-11, pangenome_analysis.py:24-48)
# --------------------------------------------------------------------------------------
def calculate_mean(df_pan_core):
    """plot.py:8-11 without the matplotlib part (:21-41)."""
    mean_values = df_pan_core.mean()
    return pd.DataFrame([mean_values], columns=df_pan_core.columns)


def fit_heaps_by_iteration(df_pan_core):
    """pangenome_analysis.py:24-48: kappa * x**alpha on the Pan half, p0=[0.5, min(y)]."""
    df = df_pan_core.iloc[:, :int(df_pan_core.shape[1] / 2)].T
    fits = {}
    for i, label in enumerate(df.columns):
        y = df.iloc[:, i].values
        popt, _ = scipy.optimize.curve_fit(
            lambda x, alpha, kappa: kappa * np.power(x, alpha),
            np.arange(1, y.shape[0] + 1), y, p0=[0.5, float(min(y))])
        fits[label] = {"alpha": popt[0], "kappa": popt[1]}
    return pd.DataFrame.from_dict(fits, orient="index").reindex(df.columns)


# --------------------------------------------------------------------------------------
# Bernoulli grid (pangenome_analysis.py:244-266)
# --------------------------------------------------------------------------------------
def bernoulli_ll(x, p, q):
    """pangenome_analysis.py:244-249."""
    probs = np.outer(p, q)
    lls = np.multiply(x, np.log(probs)) + np.multiply(1.0 - x, np.log(1.0 - probs))
    return lls.sum()


def bernoulli_grad(x, p, q):
    """pangenome_analysis.py:257-266."""
    nprobs = 1.0 - np.outer(p, q)
    dldp = x.sum(axis=1) / p - ((1.0 - x) * q[None, :] / nprobs).sum(axis=1)
    dldq = x.sum(axis=0) / q - ((1.0 - x) * p[:, None] / nprobs).sum(axis=0)
    return np.concatenate((dldp, dldq))

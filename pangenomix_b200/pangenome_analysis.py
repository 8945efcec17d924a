"""Drop-in for the analysis hot path of /root/reference/pangenomix/pangenome_analysis.py.

Same function names, arguments, printed lines, return types, labels and dtypes as the
reference; the work is done by libpgx_b200 on a B200:

* ``estimate_pan_core_size``  (:51-98)  -> PanCoreEngine (min-rank kernels, bit-exact)
* ``fit_heaps_by_iteration``  (:24-48)  -> unchanged host scipy ``curve_fit`` on the curves
* ``compute_bernoulli_grid_core_genome`` (:101-166) -> scipy L-BFGS-B on the host driving
  the fp64 likelihood/gradient kernel (BernoulliGrid)

Not provided (out of scope): the coordinate-descent variant the
reference marks "DON'T USE THIS", the beta-binomial core estimate and the mlst wrapper.
"""
from __future__ import print_function

import weakref

import numpy as np
import pandas as pd
import scipy.optimize

from .engine import BernoulliGrid, PanCoreEngine

_ENGINE_CACHE = weakref.WeakKeyDictionary()


def _fingerprint(data):
    """Cheap identity of a COO table: the arrays it is made of (address, length) plus a strided sample of their
    content (at most 4,096 entries each), so that an in-place edit of ``data.row / .col / .data`` is noticed with
    high probability without reading the whole table on every call (the reference re-reads it every time,
    pangenome_analysis.py:74-75).  Tables should still be treated as immutable once they have been rarefied."""
    parts = [tuple(data.shape), int(getattr(data, "nnz", 0)), getattr(data, "format", None)]
    for name in ("row", "col", "data", "indices", "indptr"):
        arr = getattr(data, name, None)
        if isinstance(arr, np.ndarray):
            step = max(1, arr.shape[0] // 4096)
            sample = np.ascontiguousarray(arr[::step])
            parts.append((name, arr.__array_interface__["data"][0], arr.shape[0], str(arr.dtype), hash(sample.tobytes())))
    return tuple(parts)


def _engine_for(df_genes, device=None):
    """One uploaded matrix per LSDF object ("uploaded once"); re-planned if ``.data`` is replaced or (as far as the
    sampled fingerprint sees) edited in place."""
    try:
        cached = _ENGINE_CACHE.get(df_genes)
    except TypeError:
        cached = None
    mark = _fingerprint(df_genes.data)
    if cached is not None and cached[0] is df_genes.data and cached[2] == mark and \
            (device is None or str(cached[1].device) == str(device)):
        return cached[1]
    engine = PanCoreEngine(df_genes.data, device=device)
    try:
        _ENGINE_CACHE[df_genes] = (df_genes.data, engine, mark)
    except TypeError:
        pass
    return engine


_LABEL_CACHE = {}


def _labels(prefixes, count):
    """pd.Index of prefix1..prefixN for every prefix in turn, exactly what pandas makes of the reference's label
    lists (pangenome_analysis.py:93-97).  The label data is immutable, so it is shared between calls: building
    the 20,000 column labels of a 10,000-genome table costs more than rarefying 500 permutations."""
    key = (prefixes, int(count))
    index = _LABEL_CACHE.get(key)
    if index is None:
        if len(_LABEL_CACHE) >= 16:
            _LABEL_CACHE.pop(next(iter(_LABEL_CACHE)))
        index = pd.Index([prefix + str(x) for prefix in prefixes for x in range(1, int(count) + 1)])
        _LABEL_CACHE[key] = index
    return index.copy(deep=False)          # a new Index object over the shared labels (its .name stays private)


def fit_heaps_by_iteration(df_pan_core):
    '''
    Fits Heaps Law (PG size = kappa * genomes^alpha) to the Pan half of every row of a
    table produced by estimate_pan_core_size() (or of its mean row) and returns a
    DataFrame with columns alpha, kappa indexed like the input rows.
    '''
    pan = df_pan_core.iloc[:, :int(df_pan_core.shape[1] / 2)].T      # genomes x rows
    fits = {}
    for pos, label in enumerate(pan.columns):
        alpha, kappa = __fit_heaps_single__(pan.iloc[:, pos])
        fits[label] = {'alpha': alpha, 'kappa': kappa}
    return pd.DataFrame.from_dict(fits, orient='index').reindex(pan.columns)


def fit_heaps_by_iteration_gpu(df_pan_core, device=None):
    '''
    Same table as fit_heaps_by_iteration() -- one Heaps Law fit (alpha, kappa) per row of a pan/core
    table -- computed for all rows at once on the GPU (libpgx pgx_heaps_fit: Levenberg-Marquardt with
    the reference's start point, run to fp64 convergence).  Agrees with the scipy fits of
    fit_heaps_by_iteration() to better than 5e-6 relative (scipy's own stopping tolerance); meant for fitting every iteration of a
    large table, where thousands of scipy.optimize.curve_fit calls take minutes.
    '''
    from .engine import _require_cuda, _torch, fit_heaps_device
    torch = _torch()
    dev = _require_cuda(device)
    n_points = int(df_pan_core.shape[1] / 2)
    pan = np.ascontiguousarray(df_pan_core.values[:, :n_points], dtype=np.float64)
    fit, info = fit_heaps_device(torch.from_numpy(pan).to(dev), n_points=n_points)
    fit, info = fit.cpu().numpy(), info.cpu().numpy()
    if np.any(info < 0):
        raise RuntimeError("Heaps fit did not converge for rows %s" % np.flatnonzero(info < 0)[:8].tolist())
    return pd.DataFrame(fit, index=df_pan_core.index, columns=['alpha', 'kappa'])


def __fit_heaps_single__(df_freqs):
    ''' One Heaps Law fit; start point alpha = 0.5, kappa = min(y) as in the reference (:45). '''
    y = df_freqs.values
    x = np.arange(1, y.shape[0] + 1)
    popt, _ = scipy.optimize.curve_fit(
        lambda x, alpha, kappa: kappa * np.power(x, alpha), x, y, p0=[0.5, float(min(y))])
    return popt


def estimate_pan_core_size(df_genes, num_iter, log_batch=-1, device=None):
    '''
    Computes pan/core genome size curves for many randomizations of genome order.

    Parameters
    ----------
    df_genes : LightSparseDataFrame
        Sparse binary gene x genome table (anything with .shape and a scipy .data)
    num_iter : int
        Number of randomizations; each consumes one np.arange + np.random.shuffle from
        the global numpy RNG, exactly like the reference (:84-85).
    log_batch : int
        Prints progress every log_batch runs, or silent if negative (default -1)
    device : optional CUDA device (extension; default: current device)

    Returns
    -------
    df_pan_core : pd.DataFrame
        float64, iterations Iter1.. as index, columns Pan1..PanN, Core1..CoreN.
    '''
    num_genes, num_strains = df_genes.shape
    print('Converting DataFrame to matrix...')
    engine = _engine_for(df_genes, device)
    print('Generating pan/core curves from shuffled strains')
    curves = engine.estimate(num_iter, log_batch=log_batch)

    # labels of :93-95; the curves are a fresh array nobody else holds: wrap it instead of copying it
    # (pandas 3 copies by default)
    return pd.DataFrame(curves, index=_labels(('Iter',), num_iter),
                        columns=_labels(('Pan', 'Core'), num_strains), copy=False)


def compute_bernoulli_grid_core_genome(df_genes_dense,
    prob_bounds=(0.8,0.99999999), init_capture_prob=0.9999,
    init_gene_freqs=None, device=None):
    '''
    Bernoulli-grid maximum likelihood: gene i has true frequency p_i, genome j captures
    genes at rate q_j, X_ij ~ Bernoulli(p_i q_j).  P and Q are fitted with L-BFGS-B.

    Parameters and returns are those of the reference (:101-166): a dense binary
    gene x genome DataFrame in, ``(df_opt, res)`` out, where df_opt has index
    ['Loglikelihood', 'p_<gene>'..., 'q_<genome>'...] and columns ['initial', 'optimum'],
    and res is scipy's OptimizeResult.
    '''
    n_genes, n_genomes = df_genes_dense.shape
    X = df_genes_dense.values
    grid = BernoulliGrid(X, device=device)
    if init_gene_freqs is None:
        P_guess = grid.row_count / float(n_genomes)
    else:
        P_guess = np.array(init_gene_freqs)
    Q_guess = init_capture_prob * np.ones(n_genomes)
    PQ_guess = np.clip(np.concatenate((P_guess, Q_guess)), prob_bounds[0], prob_bounds[1])
    init_ll = grid.ll_grad(PQ_guess)[0]
    print('Initial loglikelihood:', init_ll)

    labels = ['Loglikelihood'] + ['p_' + x for x in df_genes_dense.index] \
        + ['q_' + x for x in df_genes_dense.columns]
    df_init = pd.Series(data=[init_ll] + PQ_guess.tolist(), index=labels)
    df_init.name = 'initial'

    neg_ll = lambda PQ: -grid.ll_grad(PQ)[0]
    neg_ll_grad = lambda PQ: -grid.ll_grad(PQ)[1]
    bounds = [prob_bounds] * len(PQ_guess)
    res = scipy.optimize.minimize(neg_ll, PQ_guess, method='L-BFGS-B', jac=neg_ll_grad,
                                  bounds=bounds, options={'disp': True})
    print('Final loglikelihood:', -res.fun)
    df_opt = pd.Series(data=[-res.fun] + res.x.tolist(), index=labels)
    df_opt.name = 'optimum'
    return pd.concat([df_init, df_opt], axis=1), res


def __bernoulli_grid_loglikelihood__(X, P, Q):
    ''' LL of the observed table X for gene frequencies P and capture rates Q (GPU, fp64). '''
    return BernoulliGrid(X).loglikelihood(np.asarray(P, dtype=np.float64), np.asarray(Q, dtype=np.float64))


def __bernoulli_grid_loglikelihood_gradient__(X, P, Q):
    ''' Gradient of the LL with respect to concat(P, Q) (GPU, fp64). '''
    return BernoulliGrid(X).gradient(np.asarray(P, dtype=np.float64), np.asarray(Q, dtype=np.float64))

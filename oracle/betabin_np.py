"""TEST INFRASTRUCTURE -- numpy restatement of the beta-binomial core estimate and its Monte-Carlo KS test
(see oracle/__init__.py for who may import this).

Every function cites the lines of /root/reference/pangenomix it follows.  Pinned by tests/test_oracle_golden.py
against the ``betabin_*`` / ``gene_occurence_*`` fixtures that tests/golden/make_golden.py --betabin generated from
the live reference.

Third-party arithmetic restated here: numpy's legacy ``RandomState.choice(a, size, p=p)`` with replacement (call
site pangenome_analysis.py:492; numpy/random/mtrand.pyx): ``cdf = p.cumsum(); cdf /= cdf[-1]``, one
``random_sample()`` double per draw -- two successive 32-bit MT19937 outputs, ``(a >> 5, b >> 6)``,
``(a * 2**26 + b) / 2**53`` -- and ``cdf.searchsorted(u, side='right')``.  ``legacy_choice_from_raw`` is checked
against numpy itself.  scipy's Nelder-Mead (:369) and ``scipy.stats.shapiro`` (:379) stay host scipy on both sides.
"""
from __future__ import annotations

import collections

import numpy as np
import pandas as pd
import scipy.optimize
import scipy.stats
from scipy.special import betaln


def betabin_logpmf(x, n, a, b):
    """pangenome_analysis.py:501-508."""
    k = np.floor(x)
    combiln = -np.log(n + 1) - betaln(n - k + 1, k + 1)
    return combiln + betaln(k + a, n - k + b) - betaln(a, b)


def ecdf_from_counts(vals, counts, limit):
    """pangenome_analysis.py:494-499 (``pmf[vals[i]] += counts[i]`` in a Python loop: an index >= limit raises
    IndexError there and here)."""
    pmf = np.zeros(limit)
    vals = np.asarray(vals)
    if vals.size and (vals.max() >= limit or vals.min() < -limit):
        raise IndexError("index %d is out of bounds for axis 0 with size %d" % (int(vals.max()), limit))
    np.add.at(pmf, vals, np.asarray(counts, dtype=np.float64))
    return np.cumsum(pmf) / pmf.sum()


def raw_words(state, count):
    """``count`` raw 32-bit outputs continuing the legacy MT19937 state ``np.random.get_state()`` reports, and the
    state afterwards (same tuple layout)."""
    bg = np.random.MT19937()
    s = bg.state
    s["state"]["key"] = np.asarray(state[1], dtype=np.uint32)
    s["state"]["pos"] = int(state[2])
    bg.state = s
    raw = bg.random_raw(int(count)).astype(np.uint32)
    after = bg.state["state"]
    return raw, (state[0], after["key"].copy(), int(after["pos"]), state[3], state[4])


def uniforms_from_raw(raw):
    """numpy legacy ``random_sample``: one double per PAIR of raw words."""
    raw = np.asarray(raw, dtype=np.uint64)
    a = (raw[0::2] >> np.uint64(5)).astype(np.float64)
    b = (raw[1::2] >> np.uint64(6)).astype(np.float64)
    return (a * 67108864.0 + b) / 9007199254740992.0


def legacy_choice_from_raw(probs, raw):
    """``np.random.choice(np.arange(len(probs)), size=len(raw) // 2, p=probs)`` of the legacy RandomState."""
    cdf = np.asarray(probs, dtype=np.float64).cumsum()
    cdf /= cdf[-1]
    return cdf.searchsorted(uniforms_from_raw(raw), side="right").astype(np.int64)


def ks_statistics_from_raw(raw, iterations, n_samples, choice_cdf, model_cdf):
    """ks_sim of pangenome_analysis.py:471-480 from the raw words of the ``iterations * n_samples`` draws: what
    the CUDA kernel computes (the unit of work of pgx_ks_montecarlo)."""
    limit = model_cdf.shape[0]
    draws = np.asarray(choice_cdf).searchsorted(uniforms_from_raw(raw), side="right").reshape(iterations, n_samples)
    ks_sim = np.zeros(iterations)
    for i in range(iterations):
        pmf = np.bincount(draws[i], minlength=limit).astype(np.float64)
        ks_sim[i] = np.max(np.abs(np.cumsum(pmf) / pmf.sum() - model_cdf))
    return ks_sim


def ks_montecarlo_bbn(ycounts, n, a, b, iterations=100, sim_limit=1000, state=None):
    """pangenome_analysis.py:457-482 on an explicit legacy RNG state (default: the global one, which is advanced
    exactly as the reference advances it).  Returns (pvalue, ks_stat, ks_sim)."""
    use_global = state is None
    if use_global:
        state = np.random.get_state()
    xrange_ = np.arange(sim_limit)
    model_pmf = np.exp(betabin_logpmf(xrange_, n, a, b))                      # :464
    model_cdf = np.cumsum(model_pmf)
    ecdf = ecdf_from_counts(ycounts.index, ycounts.values, sim_limit)         # :468
    ks_stat = np.max(np.abs(ecdf - model_cdf))
    n_samples = int(ycounts.sum())
    probs = np.exp(betabin_logpmf(xrange_, n, a, b))                          # draw_bbn, :489-492
    probs /= probs.sum()
    raw, after = raw_words(state, 2 * n_samples * iterations)
    cdf = probs.cumsum()
    cdf /= cdf[-1]
    ks_sim = ks_statistics_from_raw(raw, iterations, n_samples, cdf, model_cdf)
    if use_global:
        np.random.set_state(after)
    pvalue = (ks_stat < ks_sim).sum() / float(iterations)
    return pvalue, ks_stat, ks_sim


def counter_spectrum(row_sums):
    """pangenome_analysis.py:355: ``pd.Series(collections.Counter(row sums))`` -- ordered by FIRST APPEARANCE."""
    return pd.Series(collections.Counter(np.asarray(row_sums)))


def durbin_watson(residuals):
    """statsmodels.stats.stattools.durbin_watson (call site :380), by definition."""
    residuals = np.asarray(residuals, dtype=np.float64)
    return float(np.sum(np.diff(residuals) ** 2) / np.sum(residuals ** 2))


def compute_beta_binomial_core_genome(row_sums, n_genomes, frac_recovered=0.999, df_counts=None, num_points=100,
                                      ks_iter=1000):
    """pangenome_analysis.py:295-400 with the table given by its row sums (the only thing :352-355 take from it)."""
    if df_counts is None:
        df_counts = counter_spectrum(row_sums)
    else:
        n_genomes = max(df_counts.index)
    results = {}
    fit_points = num_points if type(num_points) != int else [num_points]      # noqa: E721 (as the reference)
    for n_points in fit_points:
        df = df_counts.iloc[-n_points:]                                       # :364 (overrides :363)
        df = pd.Series(df.values, index=n_genomes - df.index)                # :365
        df = df.reindex(list(reversed(df.index)))                            # :366
        x = np.asarray(df.index)
        y = df.values
        nll = lambda ab: -np.dot(y, betabin_logpmf(x, n_genomes, ab[0], ab[1]))         # noqa: E731
        res = scipy.optimize.minimize(nll, x0=(1, 100), method="Nelder-Mead")
        a, b = res.x
        cutoff = 0
        cdf = np.exp(betabin_logpmf(cutoff, n_genomes, a, b))
        while cdf < frac_recovered:
            cutoff += 1
            cdf += np.exp(betabin_logpmf(cutoff, n_genomes, a, b))
        yhat = y.sum() * np.exp(betabin_logpmf(x, n_genomes, a, b))
        residuals = np.asarray(y - yhat)
        mae = np.abs(residuals).mean()
        _, sw_pvalue = scipy.stats.shapiro(residuals)
        dwstat = durbin_watson(residuals)
        model_cdf = np.cumsum(np.exp(betabin_logpmf(np.arange(n_genomes), n_genomes, a, b)))
        sim_limit = np.where(1 - model_cdf < 1e-8)[0][0]
        if sim_limit > 0:
            ks_pvalue = ks_montecarlo_bbn(df, n_genomes, a, b, iterations=ks_iter, sim_limit=sim_limit)[0]
        else:
            ks_pvalue = np.nan
        results[n_points] = pd.Series({"alpha": a, "beta": b, "cutoff": cutoff, "mae": mae,
                                       "kolmogorov_smirnov_pvalue": ks_pvalue, "shapiro_wilk_pvalue": sw_pvalue,
                                       "durbin_watson_stat": dwstat})
    table = pd.DataFrame.from_dict(results, orient="index")
    if table.shape[0] == 1:
        return table.iloc[0, :]
    return table


def count_gene_occurence(row, n_genes=None):
    """core_genome.py:127-155: (gene_index, count) of every gene that occurs, ascending gene index."""
    row = np.asarray(row)
    counts = np.bincount(row, minlength=0 if n_genes is None else n_genes)
    genes = np.flatnonzero(counts)
    return pd.DataFrame({"gene_index": genes.astype(row.dtype), "count": counts[genes].astype(np.int64)})

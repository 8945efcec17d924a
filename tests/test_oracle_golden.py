"""The oracle against the fixtures produced by the live reference (tests/golden/make_golden.py)."""
import numpy as np
import pandas as pd
import pytest
import scipy.sparse

import oracle
from conftest import BIG_CURVE_CASES, CURVE_CASES, draw_perms, golden_matrix, load_golden


@pytest.mark.parametrize("name", CURVE_CASES)
def test_minrank_matches_reference(name):
    g = load_golden(name)
    coo = golden_matrix(name, g)
    perms = draw_perms(int(g["seed"]), coo.shape[1], int(g["num_iter"]))
    pan, core = oracle.pan_core_curves_minrank(coo, perms)
    assert np.array_equal(np.hstack([pan, core]), g["curves"].astype(np.float64))


@pytest.mark.parametrize("name", [c for c in CURVE_CASES if c != "c1_8000x50"])
def test_direct_matches_reference(name):
    g = load_golden(name)
    coo = golden_matrix(name, g)
    num_iter = min(int(g["num_iter"]), 4)
    perms = draw_perms(int(g["seed"]), coo.shape[1], num_iter)
    pan, core = oracle.pan_core_curves_direct(coo, perms)
    assert np.array_equal(np.hstack([pan, core]), g["curves"][:num_iter].astype(np.float64))


@pytest.mark.parametrize("name", CURVE_CASES + BIG_CURVE_CASES)
def test_c_port_matches_reference(name):
    """oracle/pancore_ref.c -- the checker of the full-size GPU tests and bench.py's CPU arm -- against every
    curve fixture the live reference produced, directly (not through the numpy oracle)."""
    import os
    from oracle import cport
    g = load_golden(name)
    coo = golden_matrix(name, g)
    num_iter = int(g["num_iter"])
    perms = draw_perms(int(g["seed"]), coo.shape[1], num_iter)
    pan, core = cport.curves_direct(coo, perms.astype(np.int32), n_threads=min(num_iter, os.cpu_count() or 1))
    assert pan.dtype == np.float64 and np.array_equal(np.hstack([pan, core]), g["curves"].astype(np.float64))
    assert np.array_equal(cport.legacy_shuffles(int(g["seed"]), coo.shape[1], min(num_iter, 3)), perms[:3])


@pytest.mark.parametrize("name", BIG_CURVE_CASES)
def test_minrank_matches_reference_at_full_size(name):
    g = load_golden(name)
    coo = golden_matrix(name, g)
    num_iter = min(int(g["num_iter"]), 4)
    perms = draw_perms(int(g["seed"]), coo.shape[1], num_iter)
    pan, core = oracle.pan_core_curves_minrank(coo, perms)
    assert np.array_equal(np.hstack([pan, core]), g["curves"][:num_iter].astype(np.float64))


def test_c_port_sums_duplicates_like_the_reference():
    from oracle import cport
    g = load_golden("dup_3x4")
    coo = scipy.sparse.coo_matrix((g["data"], (g["row"], g["col"])), shape=tuple(g["shape"]))
    pan, core = cport.curves_direct(coo, g["perms"].astype(np.int32))
    assert np.array_equal(np.hstack([pan, core]), g["curves"].astype(np.float64))


def test_stored_perms_are_the_numpy_stream():
    g = load_golden("kat_6x5")
    assert np.array_equal(draw_perms(12345, 5, 3), g["perms"])
    assert g["perms"].tolist() == [[0, 4, 3, 1, 2], [3, 0, 2, 1, 4], [4, 0, 3, 1, 2]]


def test_survey_kat_values():
    g = load_golden("kat_6x5")
    assert g["curves"][:, :5].tolist() == [[3, 5, 5, 5, 6], [2, 3, 5, 5, 6], [4, 5, 5, 5, 6]]
    assert g["curves"][:, 5:].tolist() == [[3, 2, 2, 1, 1], [2, 2, 2, 1, 1], [4, 2, 2, 1, 1]]
    np.testing.assert_allclose(g["heaps_mean"], [0.385038, 3.155842], rtol=1e-5)


def test_duplicates_are_summed_like_the_reference():
    g = load_golden("dup_3x4")
    coo = scipy.sparse.coo_matrix((g["data"], (g["row"], g["col"])), shape=tuple(g["shape"]))
    pan, core = oracle.pan_core_curves_direct(coo, g["perms"])
    assert np.array_equal(np.hstack([pan, core]), g["curves"].astype(np.float64))
    with pytest.raises(ValueError):
        oracle.pan_core_curves_minrank(coo, g["perms"])


def test_end_to_end_frame_and_rng_consumption():
    g = load_golden("synth_800x50_s0")
    coo = golden_matrix("synth_800x50_s0", g)

    class Holder:
        shape = coo.shape
        data = coo

    np.random.seed(0)
    df = oracle.estimate_pan_core_size_minrank(Holder, 10)
    after = np.random.random_sample()
    assert df.values.dtype == np.float64
    assert np.array_equal(df.values, g["curves"].astype(np.float64))
    assert df.index[0] == "Iter1" and df.columns[0] == "Pan1" and df.columns[-1] == "Core50"
    np.random.seed(0)
    draw_perms_state = [np.random.shuffle(np.arange(50)) for _ in range(10)]
    assert after == np.random.random_sample()
    del draw_perms_state


@pytest.mark.parametrize("name", ["kat_6x5", "synth_800x50_s0", "c1_8000x50", "c2slice_4000x400"])
def test_mean_and_heaps(name):
    g = load_golden(name)
    n = int(g["shape"][1])
    cols = ["Pan%d" % (i + 1) for i in range(n)] + ["Core%d" % (i + 1) for i in range(n)]
    df = pd.DataFrame(g["curves"].astype(np.float64), columns=cols,
                      index=["Iter%d" % (i + 1) for i in range(g["curves"].shape[0])])
    mean = oracle.calculate_mean(df)
    assert np.array_equal(mean.values[0], g["mean"])
    fit = oracle.fit_heaps_by_iteration(mean)
    assert list(fit.columns) == ["alpha", "kappa"]
    np.testing.assert_allclose(fit.values[0], g["heaps_mean"], rtol=1e-12)


@pytest.mark.parametrize("n", [2, 7, 50, 400])
def test_legacy_shuffle_emulation(n):
    got = oracle.legacy_shuffle_stream(12345, n, 3)
    assert np.array_equal(got, draw_perms(12345, n, 3))


def test_bernoulli_ll_grad():
    g = load_golden("bernoulli_300x40")
    x = g["x"].astype(np.float64)
    for tag in ("0", "1"):
        p, q = g["p" + tag], g["q" + tag]
        np.testing.assert_allclose(oracle.bernoulli_ll(x, p, q), g["ll" + tag], rtol=1e-13)
        np.testing.assert_allclose(oracle.bernoulli_grad(x, p, q), g["grad" + tag], rtol=1e-12)
    np.testing.assert_allclose(oracle.bernoulli_ll(g["kat_x"], g["kat_p"], g["kat_q"]),
                               -2.502512292672613, rtol=1e-13)
    np.testing.assert_allclose(oracle.bernoulli_grad(g["kat_x"], g["kat_p"], g["kat_q"]),
                               g["kat_grad"], rtol=1e-13)


def test_bernoulli_ll_grad_at_c3_size():
    """The oracle at config C3's candidate-core size (4,000 x 400) against the live reference's LL and gradient
    at the reference's own start point and optimum."""
    import hashlib
    from pangenomix_b200 import synth
    g = load_golden("bernoulli_c3_4000x400")
    x, _, _ = synth.bernoulli_grid_matrix(4000, 400, seed=3)
    assert hashlib.sha256(x.astype(np.uint8).tobytes()).hexdigest() == str(g["x_digest"])
    for tag, pq in (("init", g["fit_initial"][1:]), ("opt", g["fit_optimum"][1:])):
        np.testing.assert_allclose(oracle.bernoulli_ll(x, pq[:4000], pq[4000:]), g["ll_" + tag], rtol=1e-13)
        np.testing.assert_allclose(oracle.bernoulli_grad(x, pq[:4000], pq[4000:]), g["grad_" + tag], rtol=1e-12)
    assert float(g["ll_init"]) == float(g["fit_initial"][0])
    np.testing.assert_allclose(g["ll_opt"], -float(g["fit_fun"]), rtol=1e-15)


@pytest.mark.skipif(__import__("os").environ.get("PGX_TEST_C5") != "1",
                    reason="config C5 at full size: 3 GB of table and a minute of C port per permutation (PGX_TEST_C5=1)")
def test_c_port_matches_reference_on_c5():
    """The live reference's own curve of ONE permutation of config C5 at full size (2,000,000 x 50,000; the reference
    needs tens of minutes for it, tests/golden/make_golden.py --c5) against the oracle's C port -- the checker of the
    full-size GPU test, which also compares the GPU curve with this fixture."""
    import os
    from oracle import cport
    g = load_golden("c5_2000000x50000")
    coo = golden_matrix("c5_2000000x50000", g)
    perms = draw_perms(int(g["seed"]), coo.shape[1], 1)
    pan, core = cport.curves_direct(coo, perms.astype(np.int32), n_threads=1)
    assert np.array_equal(np.hstack([pan, core]).astype(np.int32), g["curves"])


def test_bernoulli_ll_grad_on_the_whole_table():
    """The oracle at the whole-table size of config C3 (40,000 x 400) against the live reference's numbers at its start
    point and at its optimum (make_golden.py --c3-whole)."""
    import hashlib
    from pangenomix_b200 import synth
    g = load_golden("bernoulli_c3_40000x400")
    x, _, _ = synth.bernoulli_grid_matrix(40000, 400, seed=3)
    assert hashlib.sha256(x.astype(np.uint8).tobytes()).hexdigest() == str(g["x_digest"])
    init = np.clip(np.concatenate((x.sum(axis=1) / 400.0, 0.9999 * np.ones(400))), 0.8, 0.99999999)
    for tag, pq in (("init", init), ("opt", g["fit_x"])):
        np.testing.assert_allclose(oracle.bernoulli_ll(x, pq[:40000], pq[40000:]), g["ll_" + tag], rtol=1e-13)
        np.testing.assert_allclose(oracle.bernoulli_grad(x, pq[:40000], pq[40000:]), g["grad_" + tag], rtol=1e-12,
                                   atol=1e-12 * np.abs(g["grad_" + tag]).max())
    assert float(g["fit_ll_initial"]) == float(g["ll_init"])


def test_config_c2_mean_and_heaps_fit_of_the_reference():
    """Config C2 in full (1,000 permutations + the Heaps fit on their mean, BASELINE.json): the host half of the
    drop-in (plot.calculate_mean, fit_heaps_by_iteration) applied to the live reference's curves gives the live
    reference's mean row and fit."""
    from pangenomix_b200 import pangenome_analysis as pa, plot
    g = load_golden("c2_40000x400")
    n = int(g["shape"][1])
    assert g["curves"].shape == (1000, 2 * n)
    df = pd.DataFrame(g["curves"].astype(np.float64),
                      columns=["Pan%d" % (i + 1) for i in range(n)] + ["Core%d" % (i + 1) for i in range(n)])
    mean = plot.calculate_mean(df)
    assert np.array_equal(mean.values[0], g["mean"])
    np.testing.assert_allclose(pa.fit_heaps_by_iteration(mean).values[0], g["heaps_mean"], rtol=1e-12)
    np.testing.assert_allclose(pa.fit_heaps_by_iteration(df.iloc[:8]).values, g["heaps_iter"], rtol=1e-12)

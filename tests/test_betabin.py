"""Beta-binomial core estimate, its Monte-Carlo KS test and the table marginals (SURVEY.md section 8f, rank 4;
pangenome_analysis.py:295-400, :457-508; core_genome.py:127-155).

CPU tests: the numpy oracle against the live-reference fixtures (tests/golden/make_golden.py --betabin), the raw
MT19937 export against numpy, and the host logic of the drop-in with the device calls replaced by the oracle.
GPU tests (-m gpu): the CUDA path through the C ABI against the same fixtures and against the oracle, bit for bit.
"""
import collections
import ctypes
import os
import warnings

import numpy as np
import pandas as pd
import pytest
import scipy.sparse

from conftest import config_matrix_cached, load_golden
from oracle import betabin_np as ob
from pangenomix_b200 import _native, engine, synth

BETABIN_CASES = ["betabin_spectrum300", "betabin_spectrum2000", "betabin_spectrum2000_single", "betabin_c1",
                 "betabin_table_sorted_3600x60", "betabin_table_unsorted_800x50"]


def _case_inputs(g):
    points = g["num_points"].tolist()
    num_points = points[0] if bool(g["single"]) else points
    if "counts_index" in g:
        return None, pd.Series(g["counts_values"], index=g["counts_index"]), num_points
    shape = tuple(int(v) for v in g["shape"])
    coo = scipy.sparse.coo_matrix((np.ones(g["row"].shape[0], dtype=np.int64), (g["row"], g["col"])), shape=shape)
    return coo, None, num_points


def _as_table(out):
    return out.to_frame().T if isinstance(out, pd.Series) else out


def _check_against_fixture(out, g):
    table = _as_table(out)
    assert list(table.columns) == [str(c) for c in g["columns"]]
    assert list(table.index) == g["num_points"].tolist()
    assert isinstance(out, pd.Series) == bool(g["single"])
    assert np.array_equal(table.values.astype(np.float64), g["result"], equal_nan=True)
    state = np.random.get_state()
    assert np.array_equal(state[1], g["rng_key_after"]) and state[2] == int(g["rng_pos_after"])


# ---------------------------------------------------------------------------------------------------------
# CPU: oracle, raw stream, host logic
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", BETABIN_CASES)
def test_oracle_matches_reference(name):
    g = load_golden(name)
    coo, counts, num_points = _case_inputs(g)
    np.random.seed(int(g["seed"]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if counts is not None:
            out = ob.compute_beta_binomial_core_genome(None, None, df_counts=counts, num_points=num_points,
                                                       ks_iter=int(g["ks_iter"]))
        else:
            row_sum = np.bincount(coo.row, minlength=coo.shape[0])
            out = ob.compute_beta_binomial_core_genome(row_sum, coo.shape[1], num_points=num_points, ks_iter=int(g["ks_iter"]))
    _check_against_fixture(out, g)


@pytest.mark.parametrize("name", ["betabin_spectrum300", "betabin_table_sorted_3600x60"])
def test_oracle_ks_calls_match_reference(name):
    """Every ks_montecarlo_bbn call the reference made for the fixture, replayed in stream order."""
    g = load_golden(name)
    np.random.seed(int(g["seed"]))
    for i in range(int(g["n_ks_calls"])):
        y = pd.Series(g["ks%d_y" % i], index=g["ks%d_x" % i])
        p, stat, sim = ob.ks_montecarlo_bbn(y, int(g["ks%d_n" % i]), float(g["ks%d_a" % i]), float(g["ks%d_b" % i]),
                                            iterations=int(g["ks%d_iterations" % i]), sim_limit=int(g["ks%d_sim_limit" % i]))
        assert p == float(g["ks%d_pvalue" % i]) and stat == float(g["ks%d_ks_stat" % i])
        assert np.array_equal(sim, g["ks%d_ks_sim" % i])


@pytest.mark.parametrize("length", [1, 3, 17, 300])
def test_legacy_choice_restatement_is_numpys(length):
    probs = np.random.RandomState(length).dirichlet(np.ones(length) * 0.3)
    np.random.seed(5)
    state = np.random.get_state()
    want = np.random.choice(np.arange(length), size=10007, p=probs)
    after = np.random.get_state()
    raw, state_after = ob.raw_words(state, 2 * 10007)
    assert np.array_equal(ob.legacy_choice_from_raw(probs, raw), want)
    assert np.array_equal(state_after[1], after[1]) and state_after[2] == after[2]


def test_gene_occurence_oracle_matches_reference():
    g = load_golden("gene_occurence_800x50")
    got = ob.count_gene_occurence(g["row"], int(g["shape"][0]))
    assert list(got.columns) == [str(c) for c in g["columns"]]
    assert np.array_equal(got["gene_index"].values, g["gene_index"]) and got["gene_index"].dtype == g["gene_index"].dtype
    assert np.array_equal(got["count"].values, g["count"]) and got["count"].dtype == g["count"].dtype


@pytest.mark.parametrize("pos", [0, 1, 300, 623, 624])
@pytest.mark.parametrize("count", [0, 1, 2, 623, 624, 625, 5000])
def test_raw_words_are_numpys(pos, count):
    """pgx_legacy_random_raw continues the legacy stream from any block position, also across and onto block ends."""
    lib = _native.load()
    rs = np.random.RandomState(77)
    rs.random_sample(5)
    state = list(rs.get_state())
    state[2] = pos
    want, after = ob.raw_words(tuple(state), count)
    key = np.ascontiguousarray(state[1], dtype=np.uint32).copy()
    c_pos = ctypes.c_int32(pos)
    got = np.zeros(count + 3, dtype=np.uint32)
    assert lib.pgx_legacy_random_raw(key.ctypes.data, ctypes.byref(c_pos), count, got.ctypes.data) == 0
    assert np.array_equal(got[:count], want) and not got[count:].any()
    # numpy refills lazily: equal states or (624 of the old key) == (0 of the next)
    rs_a, rs_b = np.random.RandomState(0), np.random.RandomState(0)
    rs_a.set_state((state[0], key, int(c_pos.value), state[3], state[4]))
    rs_b.set_state(after)
    assert np.array_equal(rs_a.random_sample(700), rs_b.random_sample(700))
    if count > 0:
        assert int(c_pos.value) == after[2] and np.array_equal(key, after[1])


def test_global_stream_raw_words():
    np.random.seed(8)
    want = np.random.random_sample(1001)
    np.random.seed(8)
    got = ob.uniforms_from_raw(engine.legacy_random_raw(2002))
    assert np.array_equal(got, want)
    np.random.seed(8)
    np.random.random_sample(1001)
    follow = np.random.random_sample(4)
    np.random.seed(8)
    engine.legacy_random_raw(2002)
    assert np.array_equal(np.random.random_sample(4), follow)


def test_new_entry_points_reject_bad_arguments():
    lib = _native.load()
    assert lib.pgx_legacy_random_raw(None, None, 4, None) == 1
    assert lib.pgx_ks_montecarlo(None, 3, 0, None, None, 5, None, None, None) == 1            # n_samples < 1
    assert lib.pgx_ks_montecarlo(None, 70000, 10, None, None, 5, None, None, None) == 3       # too many iterations
    assert lib.pgx_ks_montecarlo(None, 3, 10, None, None, 5, None, None, None) == 1           # null pointers
    assert b"null pointer" in lib.pgx_last_error()
    assert lib.pgx_ks_scratch_bytes(10, 100) == 4 * 10 * 101
    assert lib.pgx_coo_marginals(None, None, -1, 3, 3, None, None, None, 0, None) == 1
    assert lib.pgx_frequency_spectrum(None, 5, 3, None, None, None) == 1
    key = np.zeros(624, dtype=np.uint32)
    pos = ctypes.c_int32(0)
    cdf = np.array([0.5, 0.4], dtype=np.float64)                                              # decreasing
    out = np.zeros(2)
    rc = lib.pgx_ks_montecarlo_host(key.ctypes.data, ctypes.byref(pos), 2, 10, cdf.ctypes.data, cdf.ctypes.data, 2,
                                    out.ctypes.data)
    assert rc == 1 and b"non-decreasing" in lib.pgx_last_error()


@pytest.fixture
def oracle_device_calls(monkeypatch):
    """The two device calls of the drop-in replaced by the oracle, to run its host logic without a GPU (the same
    arrangement as tests/test_distributed_cpu.py)."""
    def marginals(data, device=None, spectrum=True):
        coo = data.tocoo()
        row_sum = np.bincount(coo.row, minlength=coo.shape[0]).astype(np.int64)
        col_sum = np.bincount(coo.col, minlength=coo.shape[1]).astype(np.int64)
        spec = np.bincount(row_sum, minlength=coo.shape[1] + 1).astype(np.int64)
        first = np.full(coo.shape[1] + 1, np.iinfo(np.int32).max, dtype=np.int32)
        for gene in range(coo.shape[0] - 1, -1, -1):
            first[row_sum[gene]] = gene
        return row_sum, col_sum, spec, first

    def ks(choice_cdf, model_cdf, n_samples, iterations, device=None):
        raw, after = ob.raw_words(np.random.get_state(), 2 * int(n_samples) * int(iterations))
        np.random.set_state(after)
        return ob.ks_statistics_from_raw(raw, int(iterations), int(n_samples), choice_cdf, model_cdf)

    monkeypatch.setattr(engine, "table_marginals", marginals)
    monkeypatch.setattr(engine, "ks_montecarlo_statistics", ks)


def _run_drop_in(g):
    from pangenomix_b200 import pangenome_analysis as pa
    from pangenomix_b200.sparse_utils import LightSparseDataFrame
    coo, counts, num_points = _case_inputs(g)
    df_genes = None
    if coo is not None:
        index, columns = synth.labels_for(*coo.shape)
        df_genes = LightSparseDataFrame(np.array(index), np.array(columns), coo)
    np.random.seed(int(g["seed"]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return pa.compute_beta_binomial_core_genome(df_genes, df_counts=counts, num_points=num_points,
                                                    ks_iter=int(g["ks_iter"]))


@pytest.mark.parametrize("name", BETABIN_CASES)
def test_drop_in_host_logic_matches_reference(oracle_device_calls, name):
    g = load_golden(name)
    _check_against_fixture(_run_drop_in(g), g)


def test_count_gene_occurence_host_logic_matches_reference(oracle_device_calls, tmp_path, capsys):
    """core_genome.count_gene_occurence on the archive to_npz writes, the device counts replaced by numpy's."""
    from pangenomix_b200 import core_genome
    from pangenomix_b200.sparse_utils import LightSparseDataFrame
    g = load_golden("gene_occurence_800x50")
    shape = tuple(int(v) for v in g["shape"])
    coo = scipy.sparse.coo_matrix((np.ones(g["row"].shape[0], dtype=np.int64), (g["row"], g["col"])), shape=shape)
    index, columns = synth.labels_for(*shape)
    path = os.path.join(str(tmp_path), "t_strain_by_gene.npz")
    LightSparseDataFrame(np.array(index), np.array(columns), coo).to_npz(path)
    got = core_genome.count_gene_occurence(path)
    assert capsys.readouterr().out == "\nCounted gene occurence\n"
    assert list(got.columns) == [str(c) for c in g["columns"]] and list(got.index) == list(range(len(got)))
    assert np.array_equal(got["gene_index"].values, g["gene_index"]) and got["gene_index"].dtype == g["gene_index"].dtype
    assert np.array_equal(got["count"].values, g["count"]) and got["count"].dtype == g["count"].dtype
    # an archive that is not a to_npz table (no shape / format members): read as the reference reads it
    plain = os.path.join(str(tmp_path), "plain.npz")
    np.savez(plain, row=g["row"], col=g["col"])
    again = core_genome.count_gene_occurence(plain)
    assert np.array_equal(again.values, got.values)


def _check_sum_gpu_equals_sum(real_device_call):
    from pangenomix_b200.sparse_utils import LightSparseDataFrame
    coo = scipy.sparse.coo_matrix(synth.bernoulli_matrix(700, 60, 40, seed=2))
    index, columns = synth.labels_for(*coo.shape)
    lsdf = LightSparseDataFrame(np.array(index), np.array(columns), coo)
    for axis in ("index", 0, "columns", 1):
        got, want = lsdf.sum_gpu(axis), lsdf.sum(axis)
        assert got.dtype == np.int64 and np.array_equal(got, want)
    assert lsdf.sum_gpu("rows") is None and lsdf.sum("rows") is None          # the reference's silent None for other axes
    if real_device_call:                                   # entries are counted: other values are refused
        weighted = LightSparseDataFrame(np.array(index), np.array(columns), coo * 2)
        with pytest.raises(ValueError):
            weighted.sum_gpu()


def test_lsdf_sum_gpu_host_logic(oracle_device_calls):
    _check_sum_gpu_equals_sum(False)


@pytest.mark.gpu
def test_gpu_lsdf_sum_gpu(cuda):
    """LightSparseDataFrame.sum_gpu: the marginals of sparse_utils.py:284-292 counted on the device."""
    _check_sum_gpu_equals_sum(True)


def test_drop_in_helpers_match_oracle():
    from pangenomix_b200 import pangenome_analysis as pa
    x = np.arange(40)
    assert np.array_equal(pa.betabin_logpmf(x, 300, 0.7, 55.0), ob.betabin_logpmf(x, 300, 0.7, 55.0))
    vals, counts = np.array([0, 3, 4, 9]), np.array([5, 1, 7, 2])
    assert np.array_equal(pa.ecdf_from_counts(vals, counts, 12), ob.ecdf_from_counts(vals, counts, 12))
    with pytest.raises(IndexError):
        pa.ecdf_from_counts(vals, counts, 9)
    np.random.seed(4)
    got = pa.draw_bbn(300, 0.7, 55.0, 500, sim_limit=60)
    np.random.seed(4)
    raw, _ = ob.raw_words(np.random.get_state(), 1000)
    probs = np.exp(ob.betabin_logpmf(np.arange(60), 300, 0.7, 55.0))
    assert np.array_equal(got, ob.legacy_choice_from_raw(probs / probs.sum(), raw))
    with pytest.raises(ValueError):
        pa.draw_bbn(300, -2.0, 55.0, 5, sim_limit=60)


# ---------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device (run with -m gpu on the B200 box)")
    _native.load()
    return torch


@pytest.mark.gpu
@pytest.mark.parametrize("name", BETABIN_CASES)
def test_gpu_drop_in_matches_reference(cuda, name):
    """compute_beta_binomial_core_genome through the C ABI: every entry of the reference's table, NaNs included, and
    the global RNG state afterwards, bit for bit."""
    g = load_golden(name)
    _check_against_fixture(_run_drop_in(g), g)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["betabin_spectrum300", "betabin_spectrum2000", "betabin_table_sorted_3600x60"])
def test_gpu_ks_calls_match_reference(cuda, name):
    from pangenomix_b200 import pangenome_analysis as pa
    g = load_golden(name)
    np.random.seed(int(g["seed"]))
    assert int(g["n_ks_calls"]) > 0
    for i in range(int(g["n_ks_calls"])):
        y = pd.Series(g["ks%d_y" % i], index=g["ks%d_x" % i])
        p, stat, sim = pa.ks_montecarlo_bbn(y, int(g["ks%d_n" % i]), float(g["ks%d_a" % i]), float(g["ks%d_b" % i]),
                                            iterations=int(g["ks%d_iterations" % i]), sim_limit=int(g["ks%d_sim_limit" % i]))
        assert p == float(g["ks%d_pvalue" % i]) and stat == float(g["ks%d_ks_stat" % i])
        assert sim.dtype == np.float64 and np.array_equal(sim, g["ks%d_ks_sim" % i])


@pytest.mark.gpu
@pytest.mark.parametrize("n, a, b, sim_limit, n_samples, iterations", [
    (50, 0.5, 9.0, 1, 100, 7),                   # a single bin
    (50, 0.5, 9.0, 2, 31, 300),                  # fewer draws than a warp
    (300, 0.6, 90.0, 66, 6141, 1000),            # the reference's default iteration count
    (2000, 0.4, 300.0, 500, 200_000, 5),         # iterations split over CTAs, shared-memory histograms
    (30000, 2.0, 3.0, 25000, 50_000, 3),         # sim_limit beyond shared memory: global histograms, split
    (30000, 2.0, 3.0, 25000, 3000, 40),          # ... one CTA per iteration
    (400, 0.8, 40.0, 120, 257, 70000),           # more iterations than one C call takes
])
def test_gpu_ks_statistics_match_oracle(cuda, n, a, b, sim_limit, n_samples, iterations):
    support = np.arange(sim_limit)
    model_cdf = np.cumsum(np.exp(ob.betabin_logpmf(support, n, a, b)))
    probs = np.exp(ob.betabin_logpmf(support, n, a, b))
    probs /= probs.sum()
    cdf = probs.cumsum()
    cdf /= cdf[-1]
    np.random.seed(n_samples)
    state = np.random.get_state()
    got = engine.ks_montecarlo_statistics(cdf, model_cdf, n_samples, iterations)
    after = np.random.get_state()
    raw, want_after = ob.raw_words(state, 2 * n_samples * iterations)
    want = ob.ks_statistics_from_raw(raw, iterations, n_samples, cdf, model_cdf)
    assert np.array_equal(got, want)
    rs_a, rs_b = np.random.RandomState(0), np.random.RandomState(0)
    rs_a.set_state(after)
    rs_b.set_state(want_after)
    assert np.array_equal(rs_a.random_sample(700), rs_b.random_sample(700))


@pytest.mark.gpu
def test_gpu_ks_device_pointer_call(cuda):
    """pgx_ks_montecarlo on device buffers (raw words uploaded by the caller), with and without the split path."""
    torch = cuda
    lib = _native.load()
    for sim_limit, n_samples, iterations in ((40, 5000, 64), (40, 300_000, 4)):
        model_cdf = np.cumsum(np.exp(ob.betabin_logpmf(np.arange(sim_limit), 300, 0.6, 90.0)))
        cdf = model_cdf / model_cdf[-1]
        raw = np.random.RandomState(3).randint(0, 2 ** 32, size=2 * n_samples * iterations, dtype=np.uint64).astype(np.uint32)
        d_raw = torch.from_numpy(raw.view(np.int32)).cuda()
        d_cdf, d_model = torch.from_numpy(cdf).cuda(), torch.from_numpy(model_cdf).cuda()
        d_out = torch.empty(iterations, dtype=torch.float64, device="cuda")
        d_scratch = torch.empty(int(lib.pgx_ks_scratch_bytes(iterations, sim_limit)) // 4 + 1, dtype=torch.int32, device="cuda")
        _native.check(lib.pgx_ks_montecarlo(d_raw.data_ptr(), iterations, n_samples, d_cdf.data_ptr(), d_model.data_ptr(),
                                            sim_limit, d_out.data_ptr(), d_scratch.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream))
        want = ob.ks_statistics_from_raw(raw, iterations, n_samples, cdf, model_cdf)
        assert np.array_equal(d_out.cpu().numpy(), want)


def _counter_order(row_sum):
    return list(collections.Counter(row_sum.tolist()).keys())


@pytest.mark.gpu
@pytest.mark.parametrize("table", ["edge", "c1", "c2", "by_genome", "empty", "one_gene"])
def test_gpu_marginals_match_numpy(cuda, table):
    if table == "edge":
        g = load_golden("edge_500x37")
        coo = scipy.sparse.coo_matrix((g["data"], (g["row"], g["col"])), shape=tuple(g["shape"]))
        coo.sum_duplicates()
        coo = scipy.sparse.coo_matrix((np.ones(coo.nnz, dtype=np.int64), (coo.row, coo.col)), shape=coo.shape)
    elif table in ("c1", "c2"):
        coo = config_matrix_cached(table)
    elif table == "by_genome":                    # entries sorted by genome, then shuffled: no order is assumed
        coo = synth.bernoulli_matrix(3000, 130, 450, seed=4).tocoo()
        order = np.random.RandomState(1).permutation(coo.nnz)
        coo = scipy.sparse.coo_matrix((coo.data[order], (coo.row[order], coo.col[order])), shape=coo.shape)
    elif table == "empty":
        coo = scipy.sparse.coo_matrix((7, 5), dtype=np.int64)
    else:
        coo = scipy.sparse.coo_matrix((np.ones(3, dtype=np.int64), ([0, 0, 0], [4, 0, 2])), shape=(1, 6))
    row_sum, col_sum, spectrum, first = engine.table_marginals(coo)
    want_row = np.bincount(coo.row, minlength=coo.shape[0])
    assert row_sum.dtype == np.int64 and np.array_equal(row_sum, want_row)
    assert np.array_equal(col_sum, np.bincount(coo.col, minlength=coo.shape[1]))
    assert np.array_equal(spectrum, np.bincount(want_row, minlength=coo.shape[1] + 1))
    # the same numbers scipy gives the reference (sparse_utils.py:284-292)
    assert np.array_equal(row_sum, np.asarray(coo.sum(axis=1)).ravel())
    assert np.array_equal(col_sum, np.asarray(coo.sum(axis=0)).ravel())
    present = np.flatnonzero(spectrum)
    assert [int(m) for m in present[np.argsort(first[present], kind="stable")]] == _counter_order(want_row)
    assert np.all(first[spectrum == 0] == np.iinfo(np.int32).max)


@pytest.mark.gpu
def test_gpu_marginals_device_pointer_calls(cuda):
    """pgx_coo_marginals in two chunks (accumulate) + pgx_frequency_spectrum on caller-owned device buffers; entries
    outside the table are skipped and counted."""
    torch = cuda
    lib = _native.load()
    coo = synth.bernoulli_matrix(2000, 90, 450, seed=8).tocoo()
    row = np.ascontiguousarray(coo.row, dtype=np.int32).copy()
    col = np.ascontiguousarray(coo.col, dtype=np.int32).copy()
    row[5], col[11] = 2000, -1                                  # two entries outside the table
    good = np.ones(row.shape[0], dtype=bool)
    good[[5, 11]] = False
    d_row, d_col = torch.from_numpy(row).cuda(), torch.from_numpy(col).cuda()
    d_rs = torch.empty(2000, dtype=torch.int32, device="cuda")
    d_cs = torch.empty(90, dtype=torch.int32, device="cuda")
    d_bad = torch.empty(1, dtype=torch.int32, device="cuda")
    d_spec = torch.empty(91, dtype=torch.int64, device="cuda")
    d_first = torch.empty(91, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    half = 4 * (row.shape[0] // 8) + 1                          # the second chunk starts on a 4-byte boundary only
    _native.check(lib.pgx_coo_marginals(d_row.data_ptr(), d_col.data_ptr(), half, 2000, 90, d_rs.data_ptr(), d_cs.data_ptr(),
                                        d_bad.data_ptr(), 0, stream))
    _native.check(lib.pgx_coo_marginals(d_row.data_ptr() + 4 * half, d_col.data_ptr() + 4 * half, row.shape[0] - half, 2000, 90,
                                        d_rs.data_ptr(), d_cs.data_ptr(), d_bad.data_ptr(), 1, stream))
    _native.check(lib.pgx_frequency_spectrum(d_rs.data_ptr(), 2000, 90, d_spec.data_ptr(), d_first.data_ptr(), stream))
    want_row = np.bincount(row[good], minlength=2000)
    assert int(d_bad.item()) == 2
    assert np.array_equal(d_rs.cpu().numpy(), want_row)
    assert np.array_equal(d_cs.cpu().numpy(), np.bincount(col[good], minlength=90))
    spectrum, first = d_spec.cpu().numpy(), d_first.cpu().numpy()
    assert np.array_equal(spectrum, np.bincount(want_row, minlength=91))
    present = np.flatnonzero(spectrum)
    assert [int(m) for m in present[np.argsort(first[present], kind="stable")]] == _counter_order(want_row)


@pytest.mark.gpu
def test_gpu_marginals_reject_bad_tables(cuda):
    bad = scipy.sparse.coo_matrix((np.ones(2, dtype=np.int64), ([0, 1], [0, 1])), shape=(2, 2))
    bad.row = bad.row.copy()
    bad.row[1] = 5                                # outside the table
    with pytest.raises(_native.PgxError):
        engine.table_marginals(bad)
    two = scipy.sparse.coo_matrix((np.array([1, 2]), ([0, 1], [0, 1])), shape=(2, 2))
    with pytest.raises(ValueError):
        engine.table_marginals(two)


@pytest.mark.gpu
def test_gpu_count_gene_occurence_matches_reference(cuda, tmp_path, capsys):
    from pangenomix_b200 import core_genome
    from pangenomix_b200.sparse_utils import LightSparseDataFrame
    g = load_golden("gene_occurence_800x50")
    shape = tuple(int(v) for v in g["shape"])
    coo = scipy.sparse.coo_matrix((np.ones(g["row"].shape[0], dtype=np.int64), (g["row"], g["col"])), shape=shape)
    index, columns = synth.labels_for(*shape)
    path = os.path.join(str(tmp_path), "t_strain_by_gene.npz")
    LightSparseDataFrame(np.array(index), np.array(columns), coo).to_npz(path)
    got = core_genome.count_gene_occurence(path)
    assert capsys.readouterr().out == "\nCounted gene occurence\n"
    assert list(got.columns) == [str(c) for c in g["columns"]] and list(got.index) == list(range(len(got)))
    assert np.array_equal(got["gene_index"].values, g["gene_index"]) and got["gene_index"].dtype == g["gene_index"].dtype
    assert np.array_equal(got["count"].values, g["count"]) and got["count"].dtype == g["count"].dtype


@pytest.mark.gpu
def test_gpu_spectrum_of_c4_feeds_the_fit(cuda):
    """Config C4 at full size (45 M entries): marginals against numpy, and the estimate from its LSDF equals the
    estimate from the Counter-ordered spectrum computed on the host."""
    from pangenomix_b200 import pangenome_analysis as pa
    from pangenomix_b200.sparse_utils import LightSparseDataFrame
    coo = config_matrix_cached("c4")
    row_sum, col_sum, spectrum, _ = engine.table_marginals(coo)
    want_row = np.bincount(coo.row, minlength=coo.shape[0])
    assert np.array_equal(row_sum, want_row) and np.array_equal(col_sum, np.bincount(coo.col, minlength=coo.shape[1]))
    assert np.array_equal(spectrum, np.bincount(want_row, minlength=coo.shape[1] + 1))
    index, columns = synth.labels_for(*coo.shape)
    lsdf = LightSparseDataFrame(np.array(index), np.array(columns), coo)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        np.random.seed(2)
        got = pa.compute_beta_binomial_core_genome(lsdf, num_points=10, ks_iter=20)
        np.random.seed(2)
        want = ob.compute_beta_binomial_core_genome(want_row, coo.shape[1], num_points=10, ks_iter=20)
    assert np.array_equal(got.values.astype(np.float64), want.values.astype(np.float64), equal_nan=True)

"""Timing of the beta-binomial pieces on the box: the Monte-Carlo KS simulation
(pgx_ks_montecarlo_host and the device-pointer kernel alone) beside the reference's own loop restated in numpy
(np.random.choice + np.unique + ecdf_from_counts per iteration, pangenome_analysis.py:471-480), and the table
marginals beside scipy / numpy.

    python scripts/probe_betabin.py [c4|c2] > profiles/r02/probe_betabin.log
"""
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import torch  # noqa: E402

from oracle import betabin_np as ob  # noqa: E402
from pangenomix_b200 import _native, engine  # noqa: E402


def reference_loop(n, a, b, n_samples, iterations, sim_limit):
    """The reference's simulation loop (:471-480), as written there."""
    xs = np.arange(sim_limit)
    model_cdf = np.cumsum(np.exp(ob.betabin_logpmf(xs, n, a, b)))
    probs = np.exp(ob.betabin_logpmf(xs, n, a, b))
    probs /= probs.sum()
    draws = np.random.choice(xs, size=n_samples * iterations, p=probs).reshape(iterations, n_samples)
    ks_sim = np.zeros(iterations)
    for i in np.arange(iterations):
        vals, counts = np.unique(draws[i, :], return_counts=True)
        pmf = np.zeros(sim_limit)
        for j in np.arange(len(vals)):
            pmf[vals[j]] += counts[j]
        ks_sim[i] = np.max(np.abs(np.cumsum(pmf) / pmf.sum() - model_cdf))
    return ks_sim


def main():
    lib = _native.load()
    print("device:", torch.cuda.get_device_name(0))
    for n, a, b, sim_limit, n_samples, iterations in ((300, 0.56, 77.0, 66, 6141, 1000),
                                                      (2000, 0.42, 324.0, 104, 40039, 1000),
                                                      (10000, 0.4, 900.0, 300, 200_000, 1000),
                                                      (50000, 0.4, 3000.0, 600, 2_000_000, 100)):
        xs = np.arange(sim_limit)
        model_cdf = np.cumsum(np.exp(ob.betabin_logpmf(xs, n, a, b)))
        probs = np.exp(ob.betabin_logpmf(xs, n, a, b))
        cdf = (probs / probs.sum()).cumsum()
        cdf /= cdf[-1]
        draws = n_samples * iterations
        np.random.seed(1)
        engine.ks_montecarlo_statistics(cdf, model_cdf, n_samples, min(iterations, 8))           # staging, context
        best = None
        for _ in range(3):
            np.random.seed(1)
            t = time.perf_counter()
            got = engine.ks_montecarlo_statistics(cdf, model_cdf, n_samples, iterations)
            dt = time.perf_counter() - t
            best = dt if best is None else min(best, dt)
        # the raw stream alone (host) and the kernel alone (device-resident words)
        t = time.perf_counter()
        np.random.seed(1)
        raw = engine.legacy_random_raw(2 * min(draws, 50_000_000))
        rng_s = (time.perf_counter() - t) * draws / (raw.shape[0] / 2)
        it_dev = max(1, min(iterations, raw.shape[0] // (2 * n_samples)))
        d_raw = torch.from_numpy(raw[:2 * n_samples * it_dev].view(np.int32)).cuda()
        d_cdf, d_model = torch.from_numpy(cdf).cuda(), torch.from_numpy(model_cdf).cuda()
        d_out = torch.empty(it_dev, dtype=torch.float64, device="cuda")
        d_scratch = torch.empty(int(lib.pgx_ks_scratch_bytes(it_dev, sim_limit)) // 4 + 4, dtype=torch.int32, device="cuda")
        stream = torch.cuda.current_stream().cuda_stream
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rep in range(3):
            if rep == 1:
                ev0.record()
            _native.check(lib.pgx_ks_montecarlo(d_raw.data_ptr(), it_dev, n_samples, d_cdf.data_ptr(), d_model.data_ptr(),
                                                sim_limit, d_out.data_ptr(), d_scratch.data_ptr(), stream))
        ev1.record()
        torch.cuda.synchronize()
        kernel_s = ev0.elapsed_time(ev1) / 2 * 1e-3
        kernel_draws = n_samples * it_dev
        # the reference's loop on a bounded sample of the iterations
        it_ref = max(2, min(iterations, int(4e6 // n_samples)))
        np.random.seed(1)
        t = time.perf_counter()
        want = reference_loop(n, a, b, n_samples, it_ref, sim_limit)
        ref_s = time.perf_counter() - t
        assert np.array_equal(got[:it_ref], want)
        print("n=%d sim_limit=%d n_samples=%d iterations=%d: host-buffer call %.1f ms = %.2f G draws/s (MT19937 stream alone "
              "%.1f ms); kernel alone %.3f ms per %d draws = %.1f G draws/s = %.0f GB/s of raw words; reference loop "
              "%.2f s per %d iterations = %.1f M draws/s -> x%.0f; first %d statistics identical" % (
                  n, sim_limit, n_samples, iterations, best * 1e3, draws / best / 1e9, rng_s * 1e3, kernel_s * 1e3,
                  kernel_draws, kernel_draws / kernel_s / 1e9, 8 * kernel_draws / kernel_s / 1e9, ref_s, it_ref,
                  n_samples * it_ref / ref_s / 1e6, (ref_s / it_ref) / (best / iterations), it_ref), flush=True)

    from conftest import config_matrix_cached
    for name in sys.argv[1:] or ["c2"]:
        coo = config_matrix_cached(name)
        engine.table_marginals(coo)
        t = time.perf_counter()
        row_sum, col_sum, spectrum, first = engine.table_marginals(coo)
        gpu_s = time.perf_counter() - t
        t = time.perf_counter()
        want_row = np.asarray(coo.sum(axis=1)).ravel()
        want_col = np.asarray(coo.sum(axis=0)).ravel()
        scipy_s = time.perf_counter() - t
        assert np.array_equal(row_sum, want_row) and np.array_equal(col_sum, want_col)
        d_row, d_col = torch.from_numpy(np.ascontiguousarray(coo.row, dtype=np.int32)).cuda(), \
            torch.from_numpy(np.ascontiguousarray(coo.col, dtype=np.int32)).cuda()
        d_rs = torch.empty(coo.shape[0], dtype=torch.int32, device="cuda")
        d_cs = torch.empty(coo.shape[1], dtype=torch.int32, device="cuda")
        d_bad = torch.empty(1, dtype=torch.int32, device="cuda")
        stream = torch.cuda.current_stream().cuda_stream
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rep in range(4):
            if rep == 1:
                ev0.record()
            _native.check(lib.pgx_coo_marginals(d_row.data_ptr(), d_col.data_ptr(), coo.nnz, coo.shape[0], coo.shape[1],
                                                d_rs.data_ptr(), d_cs.data_ptr(), d_bad.data_ptr(), 0, stream))
        ev1.record()
        torch.cuda.synchronize()
        k_ms = ev0.elapsed_time(ev1) / 3
        print("%s marginals (nnz %d): host-buffer call %.1f ms (pageable upload of %.0f MB inside); count kernel alone %.3f ms = "
              "%.0f GB/s of COO entries; scipy .sum(axis=1) + .sum(axis=0) %.1f ms" % (
                  name, coo.nnz, gpu_s * 1e3, 8 * coo.nnz / 1e6, k_ms, 8 * coo.nnz / k_ms / 1e6, scipy_s * 1e3), flush=True)


if __name__ == "__main__":
    main()

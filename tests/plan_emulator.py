"""TEST HELPER -- numpy walk of a HostPlan that mirrors libpgx's two kernels step by step
(pangenomix_b200/csrc/pgx_rarefy.cu).  It lets the CPU suite check the folded layout,
the task table, the closed-form gene classes and the mex probe against the golden
fixtures without a GPU.  It is not a fallback: nothing in the package imports it."""
import numpy as np


def _mex_probe(perm, lst, n):
    k = 1
    while k < n:
        c = perm[k]
        lo = int(np.searchsorted(lst, c, side="left"))
        if lo < lst.shape[0] and lst[lo] == c:
            k += 1
        else:
            break
    return k


def curves_from_plan(hp, perms):
    n, g = hp.n_genomes, hp.n_genes
    perms = np.asarray(perms)
    n_perm = perms.shape[0]
    out = np.zeros((n_perm, 2 * n), dtype=np.int64)
    chunks = hp.chunks.reshape(-1, 8)
    covered = np.zeros(hp.n_rows, dtype=bool)
    sum_wp, sum_wa = int(hp.w_present.sum()), int(hp.w_absent.sum())
    for p in range(n_perm):
        perm = perms[p]
        table = np.full(n + 1, 0xFFFF, dtype=np.int64)
        table[perm] = np.arange(n)
        hist = out[p]
        for row0, meta in hp.tasks:
            n_rows, lw, flag = int(meta) >> 8, (int(meta) >> 1) & 7, int(meta) & 1
            assert 1 <= n_rows <= (32 >> lw)
            for row in range(row0, row0 + n_rows):
                if p == 0:
                    assert not covered[row]
                    covered[row] = True
                    assert bool(hp.row_absent[row]) == bool(flag)
                c0, c1 = hp.row_ptr[row], hp.row_ptr[row + 1]
                if lw < 5:
                    assert c1 - c0 <= (1 << lw)
                lst = chunks[c0:c1].reshape(-1).astype(np.int64)
                assert np.all(np.diff(lst) >= 0) and lst[hp.row_len[row] - 1] < n
                assert np.all(lst[hp.row_len[row]:] == n)
                mn = int(table[lst].min())
                list_off, other_off = (n, 0) if flag else (0, n)
                hist[list_off + mn] += 1
                if mn == 0:
                    k = _mex_probe(perm, lst, n)
                    if k < n:
                        hist[other_off + k] += 1
                else:
                    hist[other_off] += 1
        # scan kernel: closed-form classes, then prefix sums
        first = perm[0]
        pan = hist[:n].copy()
        core = hist[n:].copy()
        pan += hp.w_present[perm]
        core += hp.w_absent[perm]
        pan[0] += hp.n_full + (sum_wa - hp.w_absent[first])
        core[0] += hp.n_empty + (sum_wp - hp.w_present[first])
        if n > 1:
            pan[1] += hp.w_absent[first]
            core[1] += hp.w_present[first]
        out[p, :n] = np.cumsum(pan)
        out[p, n:] = g - np.cumsum(core)
    if n_perm:
        assert covered.all()
    return out

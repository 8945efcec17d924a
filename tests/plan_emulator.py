"""TEST HELPER -- numpy walk of a HostPlan that mirrors libpgx's rarefaction kernels step by
step (pangenomix_b200/csrc/pgx_rarefy.cu): the lane-per-row list kernel with its mex
probe, the bitmap probe kernel and the scan kernel's closed forms.  It lets the CPU suite
check the device layout against the golden fixtures without a GPU.  It is not a fallback:
nothing in the package imports it."""
import numpy as np

from pangenomix_b200.plan import residue_modulus_for


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


def list_rows(hp):
    """Yields (row, absent_flag, entries-as-stored) by walking the task table like the kernel."""
    chunks = hp.chunks.reshape(-1, 8)
    for first, meta, first_row, n_rows in hp.tasks:
        meta = int(meta)
        nch, flag = meta & 0xFFFF, (meta >> 24) & 1
        assert n_rows >= 1 and nch >= 1 and first % 32 == 0
        for sub in range((int(n_rows) + 31) // 32):
            base = int(first) + sub * nch * 32
            for lane in range(min(32, int(n_rows) - 32 * sub)):
                stored = np.concatenate([chunks[base + it * 32 + lane] for it in range(nch)])
                yield int(first_row) + 32 * sub + lane, flag, stored.astype(np.int64)


def check_layout(hp):
    """Structural invariants of the list layout; returns the shared-memory wavefronts per
    gather step (1.0 = conflict-free) the bank ordering achieves."""
    n = hp.n_genomes
    modulus = residue_modulus_for(hp.perms_per_cta)
    covered = np.zeros(hp.n_rows, dtype=bool)
    for row, flag, stored in list_rows(hp):
        assert not covered[row]
        covered[row] = True
        assert bool(hp.row_absent[row]) == bool(flag)
        want = hp.sorted_idx[hp.sorted_ptr[row]:hp.sorted_ptr[row + 1]].astype(np.int64)
        assert want.shape[0] == hp.row_len[row] and np.all(np.diff(want) > 0) and want[-1] < n
        real = stored[stored < n]
        assert np.array_equal(np.sort(real), want)
        assert np.all(stored[stored >= n] < n + 32)
    assert covered.all()
    chunks = hp.chunks.reshape(-1, 8).astype(np.int64)
    wavefronts, steps = 0, 0
    for first, meta, _, n_rows in hp.tasks:
        nch = int(meta) & 0xFFFF
        for sub in range((int(n_rows) + 31) // 32):
            base = int(first) + sub * nch * 32
            block = chunks[base:base + nch * 32].reshape(nch, 32, 8)      # [it][lane][j]
            res = block % modulus
            for g0 in range(0, 32, modulus):
                grp = res[:, g0:g0 + modulus, :]                            # [it][lane in group][j]
                counts = np.zeros((nch, 8, modulus), dtype=np.int64)
                for lane in range(grp.shape[1]):
                    np.add.at(counts, (np.arange(nch)[:, None], np.arange(8)[None, :], grp[:, lane, :]), 1)
                wavefronts += counts.max(axis=2).sum()
                steps += nch * 8
    return wavefronts / max(1, steps)


def curves_from_plan(hp, perms, hist_bins=None):
    n, g = hp.n_genomes, hp.n_genes
    perms = np.asarray(perms)
    n_perm = perms.shape[0]
    out = np.zeros((n_perm, 2 * n), dtype=np.int64)
    rows = list(list_rows(hp))
    words = 32 * hp.slice_words
    bits = hp.bits.reshape(hp.n_superblocks, n, words) if hp.n_long else None
    for p in range(n_perm):
        perm = perms[p]
        table = np.full(n + 32, 0xFFFF, dtype=np.int64)
        table[perm] = np.arange(n)
        hist = out[p]
        # list kernel: min rank of the stored list; bin 0 is never written (closed form)
        for row, flag, stored in rows:
            mn = int(table[stored].min())
            list_off, other_off = (n, 0) if flag else (0, n)
            if mn != 0:
                hist[list_off + mn] += 1
            else:
                lst = hp.sorted_idx[hp.sorted_ptr[row]:hp.sorted_ptr[row + 1]].astype(np.int64)
                k = _mex_probe(perm, lst, n)
                assert k < n
                hist[other_off + k] += 1
        # probe kernel: walk the genome order for a superblock of 1,024 W rows at once; a row's
        # statistic is the first rank whose bit differs from the rank-0 genome's bit
        for sb in range(hp.n_superblocks):
            rows_here = min(32 * words, hp.n_long - sb * 32 * words)
            pending = np.zeros(words, dtype=np.uint32)
            for w in range(words):
                left = rows_here - 32 * w
                pending[w] = 0xFFFFFFFF if left >= 32 else ((1 << left) - 1 if left > 0 else 0)
            b0 = bits[sb, perm[0]]
            for k in range(n):
                flipped = (bits[sb, perm[k]] ^ b0) & pending
                pending &= ~flipped
                hist[k] += sum(bin(int(x)).count("1") for x in flipped & ~b0)
                hist[n + k] += sum(bin(int(x)).count("1") for x in flipped & b0)
                if not pending.any():
                    break
            assert not pending.any()
        # scan kernel: closed forms, then prefix sums
        first = perm[0]
        pan = hist[:n].copy()
        core = hist[n:].copy()
        assert pan[0] == 0 and core[0] == 0
        pan[1:] += hp.w_present[perm[1:]]
        core[1:] += hp.w_absent[perm[1:]]
        pan[0] = hp.colsum[first]
        core[0] = g - hp.colsum[first]
        if n > 1:
            pan[1] += hp.w_absent[first]
            core[1] += hp.w_present[first]
        out[p, :n] = np.cumsum(pan)
        out[p, n:] = g - np.cumsum(core)
    return out

"""SURVEY.md section 8(f) rank 1, "then CUDA": how parallel is the accept / reject scan of numpy's legacy shuffle?

The raw MT19937 words can be generated anywhere; what is serial is deciding which of them are accepted (word t of a
mask regime is accepted iff ``v_t & mask <= i`` where i has dropped by one for every accept before it) and hence where
the next regime and the next shuffle start.  This script restates the scan the way a GPU would have to run it -- a
window of W words at a time, every word bracketing the number of accepts before it from below (sure accepts) and from
above (sure + doubtful), two prefix sums per round, rounds repeated until no word is in doubt, the window cut where the
regime's last accept falls -- checks the result against numpy's own shuffle, and counts the DEPENDENT block-wide
rounds per shuffle.  No GPU is used: the count times the latency of a block-wide scan round is the lower bound of a
single-CTA implementation, to be compared with the host stream's 7.3 us per 10,000-genome shuffle.

    python scripts/probe_parallel_acceptance.py > profiles/r02/probe_parallel_acceptance.log
"""
import sys

import numpy as np


def parallel_scan_shuffle(raw, at, n, window):
    """Accepted swap targets of one shuffle of n elements from raw[at:], window by window.  Returns (js, new position,
    dependent rounds, windows)."""
    js = []
    rounds = windows = 0
    i = n - 1
    while i > 0:
        bits = int(i).bit_length()
        mask = (1 << bits) - 1
        lo = 1 << (bits - 1)                       # the mask holds while i >= lo
        while i >= lo:
            need = i - lo + 1                      # accepts left in this regime
            v = (raw[at:at + window] & mask).astype(np.int64)
            w = v.shape[0]
            sure = np.zeros(w, dtype=bool)         # accepted whatever the doubtful words before do
            maybe = np.ones(w, dtype=bool)         # not yet rejected
            while True:
                rounds += 1
                before_lo = np.concatenate(([0], np.cumsum(sure)[:-1]))      # accepts before t, at least
                before_hi = np.concatenate(([0], np.cumsum(maybe)[:-1]))     # ... at most
                new_sure = v <= i - before_hi
                new_maybe = v <= i - before_lo
                if np.array_equal(new_sure, sure) and np.array_equal(new_maybe, maybe):
                    break
                sure, maybe = new_sure, new_maybe
                if np.array_equal(sure, maybe):
                    rounds += 0
            assert np.array_equal(sure, maybe)
            windows += 1
            acc = np.flatnonzero(sure)
            if acc.shape[0] >= need:               # the regime ends inside the window: cut behind its last accept
                acc = acc[:need]
                used = int(acc[-1]) + 1
            else:
                used = w
            js.extend(v[acc].tolist())
            i -= acc.shape[0]
            at += used
    return js, at, rounds, windows


def main():
    for n, window in ((10000, 1024), (10000, 4096), (10000, 16384), (50000, 4096), (50000, 16384)):
        shuffles = 24 if n <= 10000 else 8
        rs = np.random.RandomState(12345)
        state = rs.get_state()
        bg = np.random.MT19937()
        st = bg.state
        st["state"]["key"], st["state"]["pos"] = state[1], state[2]
        bg.state = st
        raw = bg.random_raw(int(shuffles * n * 1.6) + 65536).astype(np.uint32).astype(np.int64)
        at = 0
        total_rounds = total_windows = 0
        for _ in range(shuffles):
            want = np.arange(n)
            rs.shuffle(want)
            js, at, rounds, windows = parallel_scan_shuffle(raw, at, n, window)
            a = np.arange(n)
            for t, j in enumerate(js):             # Fisher-Yates from the top with the accepted targets
                k = n - 1 - t
                a[k], a[j] = a[j], a[k]
            assert np.array_equal(a, want)
            total_rounds += rounds
            total_windows += windows
        words = at / shuffles
        print("n = %6d, windows of %5d words: %.0f words per shuffle, %.1f windows and %.1f dependent block-wide rounds "
              "(two prefix sums + a compare each) per shuffle = %.2f rounds per window; at 1.5 us per round of a "
              "1,024-thread CTA that is >= %.0f us per shuffle" % (
                  n, window, words, total_windows / shuffles, total_rounds / shuffles, total_rounds / total_windows,
                  1.5 * total_rounds / shuffles), flush=True)


if __name__ == "__main__":
    sys.exit(main())

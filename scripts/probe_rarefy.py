"""Device-timing probe of the row kernels (development aid; bench.py is the contract).

    python scripts/probe_rarefy.py c4 2000 --thresholds 0,192,256,320,448 [--sweep]
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pangenomix_b200 import _native, engine, synth
from pangenomix_b200.plan import build_host_plan


def timeit(eng, d_perms, out, reps=3):
    eng.curves_device(d_perms, out=out)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.curves_device(d_perms, out=out)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="?", default="c2")
    ap.add_argument("n_perm", nargs="?", type=int, default=1000)
    ap.add_argument("--thresholds", default="")
    ap.add_argument("--sweep", action="store_true", help="also sweep (perms per CTA, row splits, threads)")
    args = ap.parse_args()
    print(_native.device_info())
    t = time.time()
    import bench
    coo = bench.load_matrix(args.which, 0, lambda: None)       # cached in /tmp per box
    print("matrix", coo.shape, coo.nnz, "gen %.1fs" % (time.time() - t), flush=True)
    n = coo.shape[1]
    np.random.seed(12345)
    perms = engine.draw_legacy_permutations(n, args.n_perm)
    d_perms = torch.from_numpy(perms.view(np.int16)).cuda()
    out = torch.empty((args.n_perm, 2 * n), dtype=torch.int32, device="cuda")
    thresholds = [int(x) for x in args.thresholds.split(",") if x] or [None]
    ref = None
    _native.profile_enable(True)
    for thr in thresholds:
        t = time.time()
        hp = build_host_plan(coo, long_threshold=thr)
        eng = engine.PanCoreEngine(coo, host_plan=hp)
        print("threshold %s: plan %.1fs, %d list rows (%d tasks, %d folded, %d slots), %d bitmap rows (%.1f MB)" % (
            hp.long_threshold, time.time() - t, hp.n_rows, hp.n_tasks, hp.folded_nnz, hp.chunks.size, hp.n_long,
            hp.bits.nbytes / 1e6), flush=True)
        a_perm = hp.algorithmic_bytes_per_perm
        configs = [(0, 0, 0)]
        if args.sweep:
            for splits in (1, 2, 4, 8, 16):
                for threads in (256, 512, 1024):
                    configs.append((0, splits, threads))
        for cfg in configs:
            _native.set_tuning(*cfg)
            try:
                ms = timeit(eng, d_perms, out)
            except Exception as e:
                print(cfg, "ERR", e)
                continue
            a, pr, b_, calls = _native.profile_read()
            print("  cfg %-14s %8.3f ms  %9.0f perms/s  %7.1f GB/s alg  (list %.3f probe %.3f scan %.3f ms per call)" % (
                cfg, ms, args.n_perm / ms * 1e3, a_perm * args.n_perm / ms / 1e6, a / calls, pr / calls, b_ / calls),
                flush=True)
        _native.set_tuning(0, 0, 0)
        # the production shape of the step: list and probe kernels side by side (per-kernel timing off)
        _native.profile_enable(False)
        ms = timeit(eng, d_perms, out, reps=5)
        print("  overlapped step %8.3f ms  %9.0f perms/s" % (ms, args.n_perm / ms * 1e3), flush=True)
        _native.profile_read()
        _native.profile_enable(True)
        res = out[:32].cpu().numpy()
        if ref is None:
            ref = res
        assert np.array_equal(ref, res), "thresholds disagree"
        del eng


if __name__ == "__main__":
    main()

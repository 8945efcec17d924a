"""Quick device-timing probe of the row kernel (development aid; bench.py is the contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pangenomix_b200 import _native, engine, synth


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
    which = sys.argv[1] if len(sys.argv) > 1 else "c2"
    n_perm = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    print(_native.device_info())
    t = time.time()
    coo = synth.config_matrix(which)
    print("matrix", coo.shape, coo.nnz, "gen %.1fs" % (time.time() - t))
    t = time.time()
    eng = engine.PanCoreEngine(coo)
    hp = eng.host_plan
    print("plan %.1fs rows %d tasks %d folded %d chunks %d" % (time.time() - t, hp.n_rows, hp.n_tasks, hp.folded_nnz, hp.n_chunks))
    n = hp.n_genomes
    np.random.seed(12345)
    t = time.time()
    perms = engine.draw_legacy_permutations(n, n_perm)
    print("draw %.2fs" % (time.time() - t))
    d_perms = torch.from_numpy(perms.view(np.int16)).cuda()
    out = torch.empty((n_perm, 2 * n), dtype=torch.int32, device="cuda")
    a_perm = hp.algorithmic_bytes_per_perm
    configs = [(0, 0, 0)]
    for b in (8, 4, 2):
        for splits in (1, 2, 4, 8, 16):
            for threads in (256, 512, 1024):
                configs.append((b, splits, threads))
    _native.profile_enable(True)
    for cfg in configs:
        _native.set_tuning(*cfg)
        try:
            ms = timeit(eng, d_perms, out)
        except Exception as e:
            print(cfg, "ERR", e)
            continue
        a, b_, calls = _native.profile_read()
        print("cfg %-16s %8.3f ms  %9.0f perms/s  %7.1f GB/s alg  (row %.3f ms scan %.3f ms per call)" % (
            cfg, ms, n_perm / ms * 1e3, a_perm * n_perm / ms / 1e6, a / calls, b_ / calls), flush=True)
    _native.set_tuning(0, 0, 0)


if __name__ == "__main__":
    main()

"""Scale probe (development aid): a C5-shaped table on one GPU -- plan time, memory, perms/s, spot parity.

    python scripts/probe_scale.py --genes 2000000 --genomes 50000 --per-genome 4000 --perms 256
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pangenomix_b200 import _native, engine, synth
from pangenomix_b200.plan import build_host_plan


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genes", type=int, default=200000)
    ap.add_argument("--genomes", type=int, default=50000)
    ap.add_argument("--per-genome", type=int, default=1000)
    ap.add_argument("--perms", type=int, default=256)
    ap.add_argument("--thresholds", default="")
    args = ap.parse_args()
    t = time.time()
    coo = synth.bernoulli_matrix(args.genes, args.genomes, args.per_genome, seed=20245, method="binomial")
    print("matrix %s nnz %d gen %.1fs" % (coo.shape, coo.nnz, time.time() - t), flush=True)
    n = args.genomes
    np.random.seed(12345)
    perms = engine.draw_legacy_permutations(n, args.perms)
    d_perms = torch.from_numpy(perms.view(np.int16)).cuda()
    out = torch.empty((args.perms, 2 * n), dtype=torch.int32, device="cuda")
    _native.profile_enable(True)
    for thr in [int(x) for x in args.thresholds.split(",") if x] or [None]:
        t = time.time()
        hp = build_host_plan(coo, long_threshold=thr)
        t_plan = time.time() - t
        eng = engine.PanCoreEngine(coo, host_plan=hp)
        print("threshold %d: plan %.1fs; %d list rows (%d tasks, %d folded, %d slots), %d bitmap rows (%.0f MB), B=%d" % (
            hp.long_threshold, t_plan, hp.n_rows, hp.n_tasks, hp.folded_nnz, hp.chunks.size, hp.n_long,
            hp.bits.nbytes / 1e6, hp.perms_per_cta), flush=True)
        for rep in range(3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.curves_device(d_perms, out=out)
            e1.record()
            torch.cuda.synchronize()
            a, pr, sc, calls = _native.profile_read()
            ms = e0.elapsed_time(e1)
        print("  %d perms: %.2f ms (list %.2f probe %.2f scan %.2f) -> %.0f perms/s, %.3g cells/s" % (
            args.perms, ms, a / calls, pr / calls, sc / calls, args.perms / ms * 1e3,
            args.perms / ms * 1e3 * args.genes * n), flush=True)
        # parity proper lives in tests/ (the oracle is test infrastructure); here only the invariants
        curves = out.cpu().numpy()
        assert np.all(np.diff(curves[:, :n], axis=1) >= 0) and np.all(np.diff(curves[:, n:], axis=1) <= 0)
        del eng


if __name__ == "__main__":
    main()

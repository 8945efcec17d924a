"""Sharding + gather of permutation blocks over world_size 2 and 3 with the gloo backend.
The per-rank compute is injected (the oracle, as the checker's arithmetic): what is under
test is the host-side partitioning, the broadcast of rank 0's RNG table and the gather."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import REPO, load_golden


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, num_iter, dst, result_dir):
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import numpy as np
    import scipy.sparse
    import torch.distributed as dist
    import oracle
    from pangenomix_b200 import distributed as pd_, sparse_utils as su
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = dict(np.load(os.path.join(REPO, "tests", "golden", "synth_800x50_s0.npz")))
    coo = scipy.sparse.coo_matrix((g["data"], (g["row"], g["col"])), shape=tuple(g["shape"]))
    lsdf = su.LightSparseDataFrame(["T_C%d" % i for i in range(800)], ["genome%d" % i for i in range(50)], coo)
    calls = []

    def compute(perms):
        calls.append(perms.shape[0])
        pan, core = oracle.pan_core_curves_minrank(coo, perms.astype(np.int64))
        return np.hstack([pan, core]).astype(np.int32)

    np.random.seed(0 if rank == 0 else 777)      # only rank 0's stream may matter
    df = pd_.estimate_pan_core_size_sharded(lsdf, num_iter, dst=dst, compute=compute)
    lo, hi = pd_.shard_bounds(num_iter, world, rank)
    assert sum(calls) == hi - lo
    if df is not None:
        np.save(os.path.join(result_dir, "r%d.npy" % rank), df.values)
        assert df.index[0] == "Iter1" and df.columns[50] == "Core1"
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,num_iter,dst", [(2, 10, None), (2, 7, 0), (3, 10, 0), (2, 1, None)])
def test_sharded_curves_equal_single_stream(tmp_path, world, num_iter, dst):
    g = load_golden("synth_800x50_s0")
    mp.spawn(_worker, args=(world, _free_port(), num_iter, dst, str(tmp_path)), nprocs=world, join=True)
    receivers = range(world) if dst is None else [dst]
    for r in receivers:
        got = np.load(str(tmp_path / ("r%d.npy" % r)))
        assert got.dtype == np.float64
        assert np.array_equal(got, g["curves"][:num_iter].astype(np.float64))
    if dst is not None:
        assert not os.path.exists(str(tmp_path / "r1.npy"))


def test_shard_bounds_cover_everything():
    from pangenomix_b200.distributed import shard_bounds
    for n in (0, 1, 7, 8, 10000):
        for world in (1, 2, 3, 8):
            blocks = [shard_bounds(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def _gather_worker(rank, world, port, result_dir):
    sys.path.insert(0, REPO)
    import numpy as np
    import torch
    import torch.distributed as dist
    from pangenomix_b200.distributed import CurveGather
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rows, width = 5, 12
    gather = CurveGather(rows, width, "cpu", dst=0, n_buffers=2)
    assert gather.mode == "nccl-gather"            # no peer memory on CPU: the collective path
    last = {}
    for step in range(5):                          # more steps than buffers: reuse after before_overwrite
        b = step % 2
        gather.before_overwrite(b)
        block = torch.full((rows, width), 1000 * step + rank, dtype=torch.int32)
        gather.send(b, block)
        last[b] = step
    gather.drain()
    dist.barrier()
    if rank == 0:
        for b, step in last.items():
            got = gather.gathered(b)
            assert len(got) == world and tuple(got[0].shape) == (rows, width)
            for r in range(world):
                assert np.all(got[r].numpy() == 1000 * step + r)
        np.save(os.path.join(result_dir, "ok.npy"), np.ones(1))
    else:
        assert gather.gathered(0) is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_curve_gather_collective_path(tmp_path, world):
    mp.spawn(_gather_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert os.path.exists(str(tmp_path / "ok.npy"))

"""The multi-GPU API on real GPUs ("1/2/4/8 ranks give identical tables"): one process per GPU,
NCCL, estimate_pan_core_size_sharded against the single-GPU DataFrame bit for bit.  Skipped on boxes with one GPU
(run it with ``gpurun --gpus 2`` or more)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import REPO

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, num_iter, dst, result_dir):
    sys.path.insert(0, REPO)
    import contextlib
    import io
    import numpy as np
    import torch
    import torch.distributed as dist
    from pangenomix_b200 import distributed as pd_, pangenome_analysis as pa, sparse_utils as su, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    coo = synth.bernoulli_matrix(20000, 1000, 2000, seed=11)
    index, columns = synth.labels_for(*coo.shape)
    lsdf = su.LightSparseDataFrame(index, columns, coo)
    np.random.seed(2024 if rank == 0 else 99 + rank)          # only rank 0's stream may matter
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        df = pd_.estimate_pan_core_size_sharded(lsdf, num_iter, log_batch=max(1, num_iter // 2), dst=dst)
    tail = np.random.random_sample(2)
    if rank == 0:
        assert out.getvalue().startswith("Converting DataFrame to matrix...\nGenerating pan/core curves from shuffled strains\n")
        # the single-GPU call on the same stream: the table every receiving rank must hold
        np.random.seed(2024)
        with contextlib.redirect_stdout(io.StringIO()):
            want = pa.estimate_pan_core_size(lsdf, num_iter)
        assert np.array_equal(tail, np.random.random_sample(2))       # the same number of shuffles was consumed
        np.save(os.path.join(result_dir, "want.npy"), want.values)
        assert list(df.index) == list(want.index) and list(df.columns) == list(want.columns)
    else:
        assert out.getvalue() == ""
    if df is not None:
        assert df.values.dtype == np.float64
        np.save(os.path.join(result_dir, "r%d.npy" % rank), df.values)
    # a second call reuses the uploaded table (engine cache per LSDF object)
    if rank == 0:
        np.random.seed(5)
    with contextlib.redirect_stdout(io.StringIO()):
        again = pd_.estimate_pan_core_size_sharded(lsdf, 3, dst=0)
    assert (again is not None) == (rank == 0)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("num_iter,dst", [(37, None), (64, 0), (1, None)])
def test_sharded_estimate_equals_single_gpu(tmp_path, num_iter, dst):
    import torch
    import torch.multiprocessing as mp
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    world = min(torch.cuda.device_count(), 8)
    mp.spawn(_worker, args=(world, _free_port(), num_iter, dst, str(tmp_path)), nprocs=world, join=True)
    want = np.load(str(tmp_path / "want.npy"))
    assert want.shape == (num_iter, 2000)
    for r in (range(world) if dst is None else [dst]):
        assert np.array_equal(np.load(str(tmp_path / ("r%d.npy" % r))), want)
    if dst is not None:
        assert not os.path.exists(str(tmp_path / "r1.npy"))

"""World-size-2/3 CPU runs of the sharded MU step logic over gloo (no GPU needed).

The step logic under test is cmf_jl_b200.sharded.ShardedMultFit / ShardPlan -- the very code the
multi-GPU bench drives with NCCL -- with the CUDA engine swapped for a NumPy engine
(tests/np_shard_engine.py).  Results must equal the single-process literal oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cnmf_oracle as po


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, X, W0, H0, L, iters, reg, out):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from cmf_jl_b200.sharded import ShardedMultFit, ShardPlan
    from tests.np_shard_engine import NumpyShard

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = ShardPlan(X.shape[1], world, L)
        t0, t1 = plan.ranges[rank]
        shard = NumpyShard(X, W0, H0, t0, t1, L)
        fitter = ShardedMultFit(shard, rank, world, dist)
        fitter.setup_data_norm()
        hist = fitter.fit(max_itr=iters, check_convergence=False, **reg)
        W, H = shard.get_factors()
        out[rank] = (hist, W, H, (t0, t1))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,reg", [(2, {}), (3, dict(l1W=0.1, l2W=0.5, l1H=0.1, l2H=0.2))])
def test_sharded_mu_matches_single_process_oracle(world, reg):
    N, T, K, L, iters = 9, 61, 3, 5, 4
    rng = np.random.default_rng(11)
    X, W0, H0 = rng.random((N, T)), rng.random((K, N, L)), rng.random((K, T))
    ref = po.fit(po.MultUpdate(X, W0, H0), X, W0, H0, iters, check_convergence=False, **reg)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), X, W0, H0, L, iters, reg, out), nprocs=world, join=True)
    H = np.zeros((K, T))
    for r in range(world):
        hist, W, Hr, (t0, t1) = out[r]
        assert np.allclose(hist, ref.loss_hist, rtol=1e-11), (r, hist, ref.loss_hist)
        assert np.allclose(W, ref.W, rtol=1e-10)
        H[:, t0:t1] = Hr
    assert np.allclose(H, ref.H, rtol=1e-10)


def test_shard_plan():
    from cmf_jl_b200.sharded import ShardPlan

    p = ShardPlan(10, 3, 3)
    assert p.ranges == [(0, 4), (4, 7), (7, 10)]
    assert p.owner(6) == 1
    with pytest.raises(ValueError):
        ShardPlan(10, 4, 5)  # shards of 2-3 columns < L-1
    assert ShardPlan(5, 1, 5).ranges == [(0, 5)]

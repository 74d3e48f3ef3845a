"""Launched by tests/test_gpu_multi.py under torch.distributed.run: a fit with one process per GPU, every collective
inside libcmf_sm100 (cmf_create_rank); rank 0 checks the histories and the gathered factors against the C oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import cmf_jl_b200 as cmf
    from oracle import c_oracle as co
    from oracle import cnmf_oracle as po

    alg = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("gloo")          # only carries the 128-byte NCCL id and the final gather of H
    N, T, K, L = 48, 1000, 5, 8
    X, _, _ = po.synthetic_sequences(K=3, N=N, L=L, T=T, rng=np.random.default_rng(21))
    W0, H0 = po.init_rand(X, L, K, np.random.default_rng(22))
    reg = dict(l1W=0.1, l2W=0.5, l1H=0.1, l2H=0.2)
    plan = cmf.ShardPlan(T, world, L)
    t0, t1 = plan.ranges[rank]
    for dtype, tol in (("f64", 1e-9), ("f32", 1e-4)):
        box = [cmf.DeviceShard.unique_id() if rank == 0 else None]      # an NCCL id builds ONE communicator
        dist.broadcast_object_list(box, src=0)
        sh = cmf.DeviceShard(N, T, t0, t1, K, L, dtype=dtype, device=int(os.environ["LOCAL_RANK"]), alg=alg,
                             comm=(box[0], rank, world), use_torch_stream=False)
        sh.set_data(X, 0)
        sh.set_factors(W0, H0, 0)
        fit = cmf.LibraryFit(sh)
        hist = fit.fit(max_itr=8, check_convergence=False, **reg)
        W, Hl = sh.get_factors()
        parts = [None] * world
        dist.all_gather_object(parts, np.asarray(Hl, dtype=np.float64))
        if rank == 0:
            rule = co.MultUpdate if alg == "mult" else co.HALSUpdate
            ref = co.fit(rule, X, W0, H0, 8, check_convergence=False, **reg)
            H = np.concatenate(parts, axis=1)
            rel = np.max(np.abs(np.asarray(hist) - np.asarray(ref.loss_hist)) / np.asarray(ref.loss_hist))
            assert rel < tol, (dtype, rel)
            ftol = 1e-8 if dtype == "f64" else 2e-3
            assert np.allclose(W, ref.W, rtol=ftol, atol=ftol * 1e-2) and np.allclose(H, ref.H, rtol=ftol, atol=ftol * 1e-2), dtype
        sh.close()
        dist.barrier()
    if rank == 0:
        print("RANK_FIT_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

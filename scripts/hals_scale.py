"""Times HALS iterations on one GPU at a BASELINE-shaped size (default: config 5 shape with a chosen T).
    python scripts/hals_scale.py --N 512 --T 1048576 --K 128 --L 32 --iters 3
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=512)
ap.add_argument("--T", type=int, default=1 << 20)
ap.add_argument("--K", type=int, default=128)
ap.add_argument("--L", type=int, default=32)
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
ge.build()
import cmf_jl_b200 as cmf  # noqa: E402

s = cmf.DeviceShard(a.N, a.T, 0, a.T, a.K, a.L, dtype="f32", device=0, alg="hals")
f = cmf.ShardedMultFit(s)
s.synth_data(1234, a.K, a.L, 0.05, 0.1)
f.setup_data_norm()
s.init_rand(0)
f.rescale_init()
print("engine", s.get_engine(), "initial loss", f.loss(), flush=True)
for it in range(a.iters):
    torch.cuda.synchronize()
    t0 = time.time()
    s.update_motifs(0.1, 0.5)
    torch.cuda.synchronize()
    t1 = time.time()
    loss = s.update_feature_maps(0.1, 0.2)
    torch.cuda.synchronize()
    t2 = time.time()
    print(f"iter {it}: W step {1e3 * (t1 - t0):.1f} ms, H step {1e3 * (t2 - t1):.1f} ms, loss {loss:.6f}", flush=True)

"""A/B timing of the numW correlation kernel under different CMF_CORR_ORDER values, interleaved in one process
so both variants see the same thermal / power state."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402

ge.build()
import cmf_jl_b200 as cmf  # noqa: E402

N, T, K, L = 4096, int(os.environ.get("AB_T", 1 << 20)), 64, 100
s = cmf.DeviceShard(N, T, 0, T, K, L, dtype="f32", device=0)
f = cmf.ShardedMultFit(s)
s.synth_data(1234, K, L, 0.05, 0.1)
f.setup_data_norm()
s.init_rand(0)
f.rescale_init()
for _ in range(2):
    s.w_partials()
torch.cuda.synchronize()
variants = os.environ.get("AB_VARIANTS", "0,1").split(",")
env_name = os.environ.get("AB_ENV", "CMF_CORR_ORDER")
for rep in range(4):
    for v in variants:
        os.environ[env_name] = v
        s.profile(True)
        for _ in range(3):
            s.w_partials()
        p = s.profile_read()
        s.profile(False)
        print(f"rep {rep} {env_name}={v}: corr {p['corr'][0] / p['corr'][1]:.2f} ms", flush=True)

"""Times BASELINE.json configs[0] and configs[1] (N=500, T=2000, K=5, L=10, Float64, 100 iterations, no early stop)
on the GPU through the public API and on the host CPU with the oracle (NumPy/OpenBLAS MU, plain-C HALS), and checks
parity of the two loss histories.  Prints one JSON object."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

ge.build()
import cmf_jl_b200 as cmf  # noqa: E402
from oracle import c_oracle as co  # noqa: E402
from oracle import cnmf_oracle as po  # noqa: E402

X, _, _ = po.synthetic_sequences(K=3, N=500, L=20, T=2000, rng=np.random.default_rng(1234))
W0, H0 = po.init_rand(X, 10, 5, np.random.default_rng(0))
reg = dict(l1W=0.1, l2W=0.5, l1H=0.1, l2H=0.2)
out = {}
for name, alg, kw, cpu in (("config1_mu", "mult", {}, lambda: po.fit(po.MultUpdate(X, W0, H0), X, W0, H0, 100, check_convergence=False)),
                           ("config2_hals_reg", "hals", reg, lambda: co.fit(co.HALSUpdate, X, W0, H0, 100, check_convergence=False, **reg))):
    cmf.fit_cnmf(X, L=10, K=5, alg=alg, max_itr=3, W_init=W0, H_init=H0, check_convergence=False, **kw)   # warm-up
    res = {}
    for dtype in ("f64", "f32"):
        t0 = time.perf_counter()
        r = cmf.fit_cnmf(X, L=10, K=5, alg=alg, max_itr=100, W_init=W0, H_init=H0, check_convergence=False, dtype=dtype, **kw)
        res[dtype] = (time.perf_counter() - t0, r)
    t0 = time.perf_counter()
    ref = cpu()
    tc = time.perf_counter() - t0
    rel = {d: float(np.max(np.abs(np.asarray(res[d][1].loss_hist) - np.asarray(ref.loss_hist)) / np.asarray(ref.loss_hist))) for d in res}
    out[name] = {"gpu_f64_s_per_100it": res["f64"][0], "gpu_f32_s_per_100it": res["f32"][0], "cpu_oracle_s_per_100it": tc,
                 "gpu_f64_it_per_s": 100 / res["f64"][0], "cpu_it_per_s": 100 / tc, "final_loss": ref.loss_hist[-1],
                 "max_rel_loss_err_f64": rel["f64"], "max_rel_loss_err_f32": rel["f32"], "cpu_threads": os.cpu_count()}
print(json.dumps(out, indent=1))

"""Accuracy of the tensor-core engine with 3 vs 2 bf16 products per logical product, against the fp64 oracle."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

ge.build()
import cmf_jl_b200 as cmf  # noqa: E402
from oracle import c_oracle as co  # noqa: E402
from oracle import cnmf_oracle as po  # noqa: E402

for (N, T, K, L, noise, p_h) in ((256, 4096, 8, 10, 1.0, 0.5), (256, 4096, 8, 10, 0.05, 0.1), (512, 16384, 16, 20, 0.3, 0.2)):
    X, _, _ = po.synthetic_sequences(K=4, N=N, L=L, T=T, noise_scale=noise, p_h=p_h, rng=np.random.default_rng(1234))
    W0, H0 = po.init_rand(X, L, K, np.random.default_rng(0))
    ref = co.fit(co.MultUpdate, X, W0, H0, 60, check_convergence=False)
    for prod in ("3", "2"):
        os.environ["CMF_TC_PRODUCTS"] = prod
        r = cmf.fit_cnmf(X, L=L, K=K, alg="mult", max_itr=60, W_init=W0, H_init=H0, check_convergence=False,
                         dtype="f32", engine=1, layout="KNL")
        rel = np.abs(np.asarray(r.loss_hist) - np.asarray(ref.loss_hist)) / np.asarray(ref.loss_hist)
        dW = np.linalg.norm(r.W - ref.W) / np.linalg.norm(ref.W)
        dH = np.linalg.norm(r.H - ref.H) / np.linalg.norm(ref.H)
        print(f"N={N} T={T} K={K} L={L} noise={noise}: products={prod} max rel loss err {rel.max():.2e} last {rel[-1]:.2e} "
              f"dW {dW:.2e} dH {dH:.2e} (final loss {ref.loss_hist[-1]:.4f})", flush=True)

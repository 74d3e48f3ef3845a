"""Generates the committed golden fixtures from the float64 oracle (numpy restatement).

The reference cannot run here (no Julia) and holds no golden vectors of its own, so these
fixtures pin the ORACLE (and through it the CUDA path) against regressions; they are not
reference outputs.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import cnmf_oracle as po  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def case(name, alg, N, T, K, L, max_itr, reg, Ktrue=3, Ltrue=None):
    rng = np.random.default_rng(1234)
    X, _, _ = po.synthetic_sequences(K=Ktrue, N=N, L=Ltrue or L, T=T, rng=rng)
    W0, H0 = po.init_rand(X, L, K, np.random.default_rng(0))
    r = po.fit_cnmf(X, L=L, K=K, alg=alg, max_itr=max_itr, W_init=W0, H_init=H0,
                    check_convergence=False, **reg)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), X=X, W0=W0, H0=H0, W=r.W, H=r.H,
                        loss_hist=np.asarray(r.loss_hist), alg=alg, max_itr=max_itr,
                        **{k: reg.get(k, 0.0) for k in ("l1W", "l2W", "l1H", "l2H")})
    print(name, r.loss_hist[0], r.loss_hist[-1])


if __name__ == "__main__":
    reg = dict(l1W=0.1, l2W=0.5, l1H=0.1, l2H=0.2)  # README.md:52 / BASELINE.json configs[1]
    case("mu_small", "mult", 40, 160, 4, 6, 30, {})
    case("mu_reg_small", "mult", 40, 160, 4, 6, 30, reg)
    case("hals_small", "hals", 30, 120, 3, 5, 10, {})
    case("hals_reg_small", "hals", 30, 120, 3, 5, 10, reg)
    rng = np.random.default_rng(7)
    W, H, X = rng.random((5, 37, 11)), rng.random((5, 211)), rng.random((37, 211))
    np.savez_compressed(os.path.join(HERE, "prims.npz"), W=W, H=H, X=X, conv=po.tensor_conv(W, H),
                        transconv=po.tensor_transconv(W, X), corr=po.corr_w(H, X, 11))

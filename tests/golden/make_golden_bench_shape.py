"""100-iteration MU fixtures at the (K, L) of the benchmark configurations, from the float64 oracle.

BASELINE.json's fp32 bar is "1e-4 relative on the loss after 100 iterations"; the headline number is quoted on K=64, L=100
(config 4) and config 3 is K=20, L=50.  The reference cannot run here (no Julia), so the histories come from the literal
NumPy restatement of src/algs/mult.jl (oracle/cnmf_oracle.py; its first iterations are cross-checked against the plain-C
restatement, oracle/cnmf_oracle.c, and the agreement is stored).  N and T are reduced so that 100 literal iterations take
minutes; the per-frequency products of the frequency-domain engine see the same K, L, block length and frequency count as
at full size.  Inputs are NOT stored: the GPU test regenerates them from the seeds below with the same oracle functions.

    python tests/golden/make_golden_bench_shape.py [case ...]       (about 25 minutes on 8 cores for all four)
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import c_oracle as co  # noqa: E402
from oracle import cnmf_oracle as po  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name -> (N, T, K, L, K_true, p_h, noise_scale)
CASES = {
    "c4shape_dense": (256, 8192, 64, 100, 8, 0.5, 3.0),        # loss stays near 0.26: the expansion loss serves (almost) every iteration
    "c4shape_sparse": (256, 8192, 64, 100, 8, 0.05, 0.05),     # sparse activations, low noise: the loss crosses the 25 % guard mid-run
    "c3shape_dense": (256, 16384, 20, 50, 8, 0.5, 3.0),
    "c3shape_sparse": (256, 16384, 20, 50, 8, 0.05, 0.05),
}
SEED_DATA, SEED_INIT, ITERS = 1234, 0, 100


def inputs(name):
    N, T, K, L, Kt, p_h, noise = CASES[name]
    X, _, _ = po.synthetic_sequences(K=Kt, N=N, L=L, T=T, p_h=p_h, noise_scale=noise, rng=np.random.default_rng(SEED_DATA))
    W0, H0 = po.init_rand(X, L, K, np.random.default_rng(SEED_INIT))
    return X, W0, H0


def make(name):
    N, T, K, L, Kt, p_h, noise = CASES[name]
    X, W0, H0 = inputs(name)
    t0 = time.time()
    r = po.fit_cnmf(X, L=L, K=K, alg="mult", max_itr=ITERS, W_init=W0, H_init=H0, check_convergence=False)
    # cross-check of the first two iterations against the plain-C restatement
    c = co.fit(co.MultUpdate, X, W0, H0, 2, check_convergence=False)
    agree = float(np.max(np.abs(np.asarray(c.loss_hist) - np.asarray(r.loss_hist[:3])) / np.asarray(c.loss_hist)))
    np.savez_compressed(
        os.path.join(HERE, f"mu_bench_{name}.npz"), loss_hist=np.asarray(r.loss_hist), dims=np.asarray([N, T, K, L]),
        K_true=Kt, p_h=p_h, noise_scale=noise, seed_data=SEED_DATA, seed_init=SEED_INIT, iters=ITERS,
        W_norm=float(np.linalg.norm(r.W)), H_norm=float(np.linalg.norm(r.H)), X_norm=float(np.linalg.norm(X)),
        H_sample=r.H[:, :: max(1, T // 64)][:, :64], numpy_vs_c_first_iterations=agree)
    print(name, "loss", r.loss_hist[0], "->", r.loss_hist[-1], "numpy vs C (3 entries):", agree, f"{time.time() - t0:.0f} s", flush=True)


if __name__ == "__main__":
    for nm in (sys.argv[1:] or list(CASES)):
        make(nm)

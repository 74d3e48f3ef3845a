"""Golden fixture from EXACT rational arithmetic (oracle/exact_pin.py): one MultUpdate iteration and one HALSUpdate
iteration on the reference's toy data (datasets/toy.jl:5-48, N=7, T=250, K=2, L=5) from dyadic starting factors, all four
regularisers on.  The stored W, H are the exact rationals rounded once to Float64; the losses take one square root.
Unlike the other fixtures this one is not produced by floating-point code, so it pins the oracles AND the CUDA path.

    python tests/golden/make_exact_pin.py        (about a minute; a second exact HALS iteration is out of reach: the rationals grow without bound)
"""
import os
import sys
from fractions import Fraction as Fr

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import exact_pin as ex  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REG = dict(l1W=Fr(1, 10), l2W=Fr(1, 2), l1H=Fr(1, 10), l2H=Fr(1, 5))

if __name__ == "__main__":
    X, Wt, Ht = ex.toy_data_exact(5)
    N, T, K, L = ex.dims(Wt, Ht)
    W0, H0 = ex.rational_init(K, N, L, T)
    Wm, Hm, lm = ex.mu_iteration(X, W0, H0, **REG)
    W1, H1, R1, l1 = ex.hals_iteration(X, W0, H0, None, **REG)
    f = lambda a: np.asarray(ex.to_float(a), dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "exact_pin_toy.npz"), X=f(X), W0=f(W0), H0=f(H0), mu_W=f(Wm), mu_H=f(Hm), mu_loss=lm,
                        hals_W=f(W1), hals_H=f(H1), hals_R=f(R1), hals_loss=l1,
                        reg=np.asarray([0.1, 0.5, 0.1, 0.2]))
    print("mu loss", lm, "hals loss", l1)

"""TEST INFRASTRUCTURE (not product code): an exact-arithmetic pin of the oracles.

The reference holds no golden vectors and Julia is not installed, so the floating-point oracles
(oracle/cnmf_oracle.py, oracle/cnmf_oracle.c) cannot be compared with reference outputs.  This module restates one
MultUpdate iteration and one HALSUpdate iteration in EXACT rational arithmetic (`fractions.Fraction`), written directly
from the Julia sources and sharing no code with the other oracles:

    tensor_conv!            src/common.jl:24-34   (s_dot! :108-118: est[:, lag+1:T] += W[:,:,lag+1]' * H[:, 1:T-lag])
    tensor_transconv!       src/common.jl:71-81   (out[:, 1:T-lag] += W[:,:,lag+1] * X[:, 1+lag:T])
    MultUpdate iteration    src/algs/mult.jl:23-58
    HALSUpdate iteration    src/algs/hals.jl:18-154 (persistent residual, k outer / l inner W sweep, k outer / t inner H sweep)
    EPSILON = eps(Float64)  src/CMF.jl:20  = 2^-52 exactly

Every quantity up to and including the updated W and H is an exact rational; only the final norm of the loss takes a
square root (of a sum formed with `math.fsum` over correctly rounded terms).  A floating-point implementation that
follows the same formulas must agree with these values to rounding error -- a check in which no floating-point code
validates itself.  Indices below are 0-based; W is K x N x L (current-src layout), data `datasets/toy.jl:5-48`.
"""
from fractions import Fraction as Fr
import math

EPS = Fr(1, 2 ** 52)   # eps(Float64), src/CMF.jl:20


def toy_data_exact(tiles=5):
    """datasets/toy.jl:5-48: N=7, T=50*tiles, K=2, L=5; returns (X, W, H) as nested lists of Fractions.
    toy.jl writes W as L x N x K; W[k][n][l] here is its W[l, n, k]."""
    N, T0, K, L = 7, 50, 2, 5
    rows1 = [[0, 0, 0, 0, 0], [0, 0, 0, 0, 0], [0, 4, 0, 4, 0], [0, 0, 0, 0, 0], [3, 0, 0, 0, 3], [0, 2, 1, 2, 0], [0, 0, 0, 0, 0]]
    rows2 = [[0, 1, 0, 0, 0], [0, 0, 2, 0, 0], [0, 0, 0, 3, 0], [0, 5, 0, 0, 0], [0, 0, 5, 0, 0], [0, 0, 0, 5, 0], [0, 0, 0, 0, 0]]
    # toy.jl: W[:, :, k] = rows'  =>  W[l, n, k] = rows[n][l]
    W = [[[Fr(r[n][l]) for l in range(L)] for n in range(N)] for r in (rows1, rows2)]
    H0 = [[Fr(0)] * T0 for _ in range(K)]
    for (k, t1, v) in [(0, 2, Fr(1)), (0, 20, Fr(2)), (0, 32, Fr(1)), (0, 48, Fr(1)), (1, 12, Fr(1)), (1, 30, Fr(1)), (1, 38, Fr(1, 2))]:
        H0[k][t1 - 1] = v                                    # Julia's 1-based column t1
    H = [row * tiles for row in H0]
    return conv(W, H), W, H


def dims(W, H):
    return len(W[0]), len(H[0]), len(W), len(W[0][0])       # N, T, K, L


def conv(W, H):
    """est[n][t] = sum_{l=0}^{min(L-1,t)} sum_k W[k][n][l] * H[k][t-l]   (common.jl:24-34 with s_dot! :108-118)"""
    N, T, K, L = dims(W, H)
    est = [[Fr(0)] * T for _ in range(N)]
    for lag in range(L):
        for n in range(N):
            row = est[n]
            for k in range(K):
                w = W[k][n][lag]
                if w == 0:
                    continue
                Hk = H[k]
                for t in range(lag, T):                     # est[:, lag+1:T] += W_lag' * H[:, 1:T-lag]
                    row[t] += w * Hk[t - lag]
    return est


def transconv(W, X):
    """out[k][t] = sum_{l=0}^{min(L-1,T-1-t)} sum_n W[k][n][l] * X[n][t+l]   (common.jl:71-81)"""
    N, T, K, L = len(X), len(X[0]), len(W), len(W[0][0])
    out = [[Fr(0)] * T for _ in range(K)]
    for lag in range(L):
        for k in range(K):
            row = out[k]
            for n in range(N):
                w = W[k][n][lag]
                if w == 0:
                    continue
                Xn = X[n]
                for t in range(0, T - lag):                 # out[:, 1:T-lag] += W_lag * X[:, 1+lag:T]
                    row[t] += w * Xn[t + lag]
    return out


def rel_loss(est, X):
    """norm(est - X) / norm(X) (common.jl:54-55): exact residuals, each rounded once, summed with fsum."""
    r2 = math.fsum(float((e - x) * (e - x)) for er, xr in zip(est, X) for e, x in zip(er, xr))
    x2 = math.fsum(float(x * x) for xr in X for x in xr)
    return math.sqrt(r2) / math.sqrt(x2)


def mu_iteration(X, W, H, l1W=Fr(0), l2W=Fr(0), l1H=Fr(0), l2H=Fr(0)):
    """One update_motifs! + update_feature_maps! of MultUpdate (mult.jl:23-58).  Returns (W', H', loss)."""
    N, T, K, L = dims(W, H)
    est = conv(W, H)                                                        # :28
    Wn = [[[None] * L for _ in range(N)] for _ in range(K)]
    for lag in range(L):                                                    # :31-34
        for k in range(K):
            for n in range(N):
                num = sum((H[k][t] * X[n][t + lag] for t in range(T - lag)), Fr(0))
                den = sum((H[k][t] * est[n][t + lag] for t in range(T - lag)), Fr(0))
                w = W[k][n][lag]
                w = w * num / (den + l1W + 2 * l2W * w + EPS)               # :37
                Wn[k][n][lag] = max(EPS, w)                                 # :38
    est = conv(Wn, H)                                                       # :44
    numH = transconv(Wn, X)                                                 # :47
    denH = transconv(Wn, est)                                               # :48
    Hn = [[max(EPS, H[k][t] * numH[k][t] / (denH[k][t] + l1H + 2 * l2H * H[k][t] + EPS)) for t in range(T)] for k in range(K)]  # :51-52
    return Wn, Hn, rel_loss(conv(Wn, Hn), X)                                # :55-57


def hals_iteration(X, W, H, R=None, l1W=Fr(0), l2W=Fr(0), l1H=Fr(0), l2H=Fr(0)):
    """One update_motifs! + update_feature_maps! of HALSUpdate on the persistent residual R (hals.jl:18-154).
    R = None builds it as the constructor does (:22).  Returns (W', H', R', loss); W, H are not modified."""
    N, T, K, L = dims(W, H)
    W = [[list(r) for r in Wk] for Wk in W]
    H = [list(r) for r in H]
    if R is None:
        est = conv(W, H)
        R = [[est[n][t] - X[n][t] for t in range(T)] for n in range(N)]     # :22
    else:
        R = [list(r) for r in R]
    # ---- W sweep: H_unfold row ind = l*K + k holds H[k] shifted right by l (common.jl:133-142); norms squared exactly (:56-60)
    for k in range(K):                                                      # :92-96  k outer, l inner
        for l in range(L):
            hs = [Fr(0)] * l + H[k][: T - l]                                # H_unfold[ind, :]
            nrm2 = sum((h * h for h in hs), Fr(0))                          # H_norms[ind]^2
            wcol = [W[k][n][l] for n in range(N)]
            for n in range(N):                                              # :104  resids .-= w * h'
                wn = wcol[n]
                if wn != 0:
                    Rn = R[n]
                    for t in range(l, T):
                        Rn[t] -= wn * hs[t]
            new = [max((-sum((R[n][t] * hs[t] for t in range(l, T)), Fr(0)) - l1W) / (nrm2 + EPS + l2W), Fr(0)) for n in range(N)]  # :105,:111
            for n in range(N):                                              # :106
                W[k][n][l] = new[n]
                if new[n] != 0:
                    Rn = R[n]
                    for t in range(l, T):
                        Rn[t] += new[n] * hs[t]
    # ---- H sweep (:124-128 k outer, t inner; W_norms / Wk_list frozen :64-80)
    Wn2 = [[sum((W[k][n][l] ** 2 for n in range(N)), Fr(0)) for l in range(L)] for k in range(K)]   # W_norms[k,l]^2
    for k in range(K):
        for t in range(T):
            w = min(T - t, L)                                               # min(T-t+1, L) in 1-based t (:137)
            nrm2 = sum(Wn2[k][:w], Fr(0))                                   # norm(W_norms[k, 1:w])^2
            h = H[k][t]
            trace = Fr(0)
            for l in range(w):
                for n in range(N):
                    rem = R[n][t + l] - h * W[k][n][l]                      # :140-141
                    R[n][t + l] = rem
                    trace -= W[k][n][l] * rem                               # dot(Wkt, -remainder) :152
            hn = max((trace - l1H) / (nrm2 + EPS + l2H), Fr(0))             # :153
            H[k][t] = hn                                                    # :144
            if hn != 0:
                for l in range(w):
                    for n in range(N):
                        R[n][t + l] += hn * W[k][n][l]                      # :147
    r2 = math.fsum(float(r * r) for Rn in R for r in Rn)
    x2 = math.fsum(float(x * x) for xr in X for x in xr)
    return W, H, R, math.sqrt(r2) / math.sqrt(x2)                          # :41


def to_float(a):
    """Nested lists of Fractions -> nested lists of correctly rounded floats."""
    if isinstance(a, list):
        return [to_float(x) for x in a]
    return float(a)


def rational_init(K, N, L, T):
    """Deterministic strictly positive rational starting factors (dyadic, so they are exact in Float64 too)."""
    W = [[[Fr(1 + ((3 * k + 5 * n + 7 * l) % 8), 8) for l in range(L)] for n in range(N)] for k in range(K)]
    H = [[Fr(1 + ((5 * k + 3 * t) % 16), 16) for t in range(T)] for k in range(K)]
    return W, H

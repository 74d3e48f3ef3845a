"""CPU oracle for the CNMF fit hot path of degleris1/CMF.jl  --  TEST INFRASTRUCTURE ONLY.

This file is a NumPy float64 restatement of the reference algorithm.  It is the
checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product path (``cmf.jl_b200``) never routes through it and has no CPU fallback.

PARITY UNPINNED.  The reference ships no tests, golden vectors or fixtures for this
path (no ``@test`` anywhere) and Julia is not installed in this image,
so the oracle cannot be checked against reference outputs.  It is pinned instead by
(i) the element-wise definition in ``notebooks/benchmarks.ipynb`` (``tconv4``),
(ii) adjointness of the three contractions, (iii) the ``tconv3`` identity
``conv = W_unf * shift_and_stack(H)``, (iv) agreement with the independent plain-C
restatement in ``oracle/cnmf_oracle.c`` and (v) MU monotonicity -- see
``tests/test_oracle.py``.

Conventions follow the *current* ``src/`` generation of the reference:
``data`` is N x T, ``W`` is K x N x L, ``H`` is K x T (Julia column-major at the C-ABI;
here plain NumPy index order ``W[k, n, l]``, ``H[k, t]``, ``X[n, t]``, 0-based lags).

Each function cites the reference file:line (relative to /root/reference) it follows.
"""
from __future__ import annotations

import time as _time

import numpy as np

# src/CMF.jl:20   const EPSILON = eps()   (Float64 machine epsilon)
EPSILON = float(np.finfo(np.float64).eps)


# --------------------------------------------------------------------------------------
# Tensor primitives  (src/common.jl)
# --------------------------------------------------------------------------------------
def tensor_conv(W, H):
    """src/common.jl:17-34 (tensor_conv, tensor_conv!, s_dot! :108-118).

    est[:, l:T] += W[:, :, l]' * H[:, 0:T-l]   for each lag l  (one dgemm per lag).
    """
    K, N, L = W.shape
    T = H.shape[1]
    est = np.zeros((N, T), dtype=np.result_type(W, H))
    for lag in range(L):
        if lag >= T:
            break
        est[:, lag:] += W[:, :, lag].T @ H[:, : T - lag]
    return est


def tensor_transconv(W, X):
    """src/common.jl:62-81 (tensor_transconv!, shift_cols :121-130).

    out[:, 0:T-l] += W[:, :, l] * X[:, l:T]   for each lag l.
    """
    K, N, L = W.shape
    T = X.shape[1]
    out = np.zeros((K, T), dtype=np.result_type(W, X))
    for lag in range(L):
        if lag >= T:
            break
        out[:, : T - lag] += W[:, :, lag] @ X[:, lag:]
    return out


def corr_w(H, X, L):
    """src/algs/mult.jl:31-34: numW[:, :, l] = H[:, 0:T-l] * X[:, l:T]'  (K x N per lag)."""
    K, T = H.shape
    N = X.shape[0]
    out = np.zeros((K, N, L), dtype=np.result_type(H, X))
    for lag in range(L):
        if lag >= T:
            break
        out[:, :, lag] = H[:, : T - lag] @ X[:, lag:].T
    return out


def shift_and_stack(H, L):
    """src/common.jl:133-142: H_stacked[l*K + k, t] = H[k, t - l] (zero for t < l)."""
    K, T = H.shape
    Hs = np.zeros((L * K, T), dtype=H.dtype)
    for lag in range(L):
        if lag >= T:
            break
        Hs[K * lag : K * (lag + 1), lag:] = H[:, : T - lag]
    return Hs


def compute_resids(data, W, H):
    """src/common.jl:58-59."""
    return tensor_conv(W, H) - data


def compute_loss(data, W, H):
    """src/common.jl:54-55: ||conv(W,H) - data||_F / ||data||_F."""
    return float(np.linalg.norm(compute_resids(data, W, H)) / np.linalg.norm(data))


# --- naive element-wise definitions: the index-convention anchors --------------------
def naive_conv(W, H):
    """notebooks/benchmarks.ipynb `tconv4` (element-wise quadruple loop), K x N x L layout."""
    K, N, L = W.shape
    T = H.shape[1]
    X = np.zeros((N, T))
    for t in range(T):
        for k in range(K):
            for n in range(N):
                for l in range(min(L, t + 1)):
                    X[n, t] += W[k, n, l] * H[k, t - l]
    return X


def naive_transconv(W, X):
    """Element-wise form of src/common.jl:71-81."""
    K, N, L = W.shape
    T = X.shape[1]
    out = np.zeros((K, T))
    for t in range(T):
        for k in range(K):
            for l in range(min(L, T - t)):
                for n in range(N):
                    out[k, t] += W[k, n, l] * X[n, t + l]
    return out


def naive_corr_w(H, X, L):
    """Element-wise form of src/algs/mult.jl:31-34."""
    K, T = H.shape
    N = X.shape[0]
    out = np.zeros((K, N, L))
    for l in range(L):
        for k in range(K):
            for n in range(N):
                for t in range(T - l):
                    out[k, n, l] += H[k, t] * X[n, t + l]
    return out


# --------------------------------------------------------------------------------------
# Initialisation and convergence  (src/model.jl)
# --------------------------------------------------------------------------------------
def init_rand(data, L, K, rng):
    """src/model.jl:113-125.  W ~ U[0,1)^(K,N,L) drawn first, then H ~ U[0,1)^(K,T);
    both rescaled by sqrt(|alpha|), alpha = <data, est> / ||est||^2.

    Julia's RNG stream cannot be reproduced here; ``rng`` is a numpy Generator.  Parity
    tests always pass explicit W_init / H_init to both implementations.
    """
    N, T = data.shape
    W = rng.random((K, N, L))
    H = rng.random((K, T))
    return rescale_init(data, W, H)


def rescale_init(data, W, H):
    """src/model.jl:119-122 (the alpha rescale, split out so callers can bring their own draws)."""
    est = tensor_conv(W, H)
    alpha = float(np.vdot(data, est) / np.linalg.norm(est) ** 2)
    s = np.sqrt(abs(alpha))
    return W * s, H * s


def converged(loss_hist, patience, tol):
    """src/model.jl:91-107."""
    if len(loss_hist) <= patience:
        return False
    d_loss = np.diff(np.asarray(loss_hist[-(patience + 1) :]))
    return bool(np.all(np.abs(d_loss) < tol))


# --------------------------------------------------------------------------------------
# Multiplicative updates  (src/algs/mult.jl)
# --------------------------------------------------------------------------------------
class MultUpdate:
    """src/algs/mult.jl:1-20."""

    def __init__(self, data, W, H):
        self.resids = compute_resids(data, W, H)
        self.data_norm = float(np.linalg.norm(data))
        self.est = np.zeros_like(data)

    def update_motifs(self, data, W, H, l1W=0.0, l2W=0.0, **_):
        """src/algs/mult.jl:23-39 (mutates W in place)."""
        K, N, L = W.shape
        self.est = tensor_conv(W, H)
        numW = corr_w(H, data, L)
        denomW = corr_w(H, self.est, L)
        W *= numW / (denomW + l1W + 2 * l2W * W + EPSILON)
        np.maximum(W, EPSILON, out=W)

    def update_feature_maps(self, data, W, H, l1H=0.0, l2H=0.0, **_):
        """src/algs/mult.jl:42-58 (mutates H in place, returns the relative loss)."""
        self.est = tensor_conv(W, H)
        numH = tensor_transconv(W, data)
        denomH = tensor_transconv(W, self.est)
        H *= numH / (denomH + l1H + 2 * l2H * H + EPSILON)
        np.maximum(H, EPSILON, out=H)
        self.est = tensor_conv(W, H)
        self.resids = self.est - data
        return float(np.linalg.norm(self.resids) / self.data_norm)


# --------------------------------------------------------------------------------------
# HALS  (src/algs/hals.jl)  -- literal sweeps; Python loops, so small cases only.
# The plain-C twin in oracle/cnmf_oracle.c runs the same sweeps fast.
# --------------------------------------------------------------------------------------
class HALSUpdate:
    """src/algs/hals.jl:6-28: persistent residual, maintained incrementally."""

    def __init__(self, data, W, H):
        self.resids = tensor_conv(W, H) - data
        self.data_norm = float(np.linalg.norm(data))

    def update_motifs(self, data, W, H, l1W=0.0, l2W=0.0, **_):
        """src/algs/hals.jl:31-34,53-61,90-112."""
        K, N, L = W.shape
        H_unfold = shift_and_stack(H, L)
        H_norms = np.linalg.norm(H_unfold, axis=1)
        R = self.resids
        for k in range(K):  # k outer, lag inner: hals.jl:92-96
            for l in range(L):
                ind = l * K + k  # hals.jl:102 (0-based)
                hkl = H_unfold[ind]
                R -= np.outer(W[k, :, l], hkl)  # :104
                W[k, :, l] = np.maximum(
                    (-R @ hkl - l1W) / (H_norms[ind] ** 2 + EPSILON + l2W), 0.0
                )  # :105,:111
                R += np.outer(W[k, :, l], hkl)  # :106

    def update_feature_maps(self, data, W, H, l1H=0.0, l2H=0.0, **_):
        """src/algs/hals.jl:37-42,64-80,121-154."""
        K, N, L = W.shape
        T = H.shape[1]
        W_norms = np.linalg.norm(W, axis=1)  # (K, L): ||W[k, :, l]||   hals.jl:68-73
        R = self.resids
        for k in range(K):  # k outer, t inner: hals.jl:124-128
            Wk = W[k]  # N x L   hals.jl:76-79
            for t in range(T):
                w = min(T - t, L)
                norm_Wkt = np.linalg.norm(W_norms[k, :w])  # :137
                rem = R[:, t : t + w]
                rem += -H[k, t] * Wk[:, :w]  # :141 (in place on the view)
                trace = float(np.vdot(Wk[:, :w], -rem))  # :152
                H[k, t] = max((trace - l1H) / (norm_Wkt ** 2 + EPSILON + l2H), 0.0)  # :153
                R[:, t : t + w] = rem + H[k, t] * Wk[:, :w]  # :147
        return float(np.linalg.norm(R) / self.data_norm)  # :41


# --------------------------------------------------------------------------------------
# Projected gradient descent  (src/algs/pgd.jl)  -- SquareLoss, Square/Absolute penalties,
# NonnegConstraint: the configuration the reference's own callers use (test/test.jl:28,
# figures/thesis/*.jl) minus the masked / absolute losses.
# --------------------------------------------------------------------------------------
class PGDUpdate:
    """src/algs/pgd.jl:112-155.  Penalties are given as weights: ``l2W`` = SquarePenalty weight on W
    (reference default ``penaltiesW=[SquarePenalty(1)]``, pgd.jl:161), ``l1W`` = AbsolutePenalty weight,
    likewise ``l1H``/``l2H`` (reference default ``penaltiesH=[]``, pgd.jl:185).  ``loss_func`` is "square"
    (SquareLoss, pgd.jl:28-35) or "absolute" (AbsoluteLoss, :38-45); ``mask`` (N x T) wraps it in a MaskedLoss
    (:59-70: the gradient is multiplied by the mask, the loss is evaluated on mask.*data and mask.*est).  ``constrW`` /
    ``constrH`` are "nonneg" (NonnegConstraint, :91-95, the default) or "unitnorm" (UnitNormConstraint, :98-110: every slice
    along the first dimension -- one component -- whose 2-norm exceeds 1 is divided by it; no non-negativity)."""

    def __init__(self, data, W, H, loss_func="square", mask=None, constrW="nonneg", constrH="nonneg"):
        self.datanorm = float(np.linalg.norm(data))
        self.est = tensor_conv(W, H)
        self.stepW = 5.0          # pgd.jl:147-148
        self.stepH = 5.0
        self.cur_loss = self.datanorm   # pgd.jl:149 (sic: the norm, not its square)
        self.step_incr, self.step_decr = 1.05, 0.70
        self.loss_func, self.mask = loss_func, mask
        self.constrW, self.constrH = constrW, constrH

    @staticmethod
    def _project(x, constr):
        if constr == "nonneg":
            np.maximum(x, EPSILON, out=x)                 # NonnegConstraint :93-95
        else:
            for m in range(x.shape[0]):                   # UnitNormConstraint :101-110
                mag = float(np.linalg.norm(x[m]))
                if mag > 1:
                    x[m] /= mag

    def _loss_grad(self, data):
        g = 2.0 * (self.est - data) if self.loss_func == "square" else np.sign(self.est - data)   # :30-32 / :40-42
        return g if self.mask is None else g * self.mask                                           # :63-66

    def _loss_eval(self, data):
        b, e = (data, self.est) if self.mask is None else (self.mask * data, self.mask * self.est)  # :67-69
        return float(np.linalg.norm(b - e) ** 2) if self.loss_func == "square" else float(np.sum(np.abs(b - e)))

    def _pgd(self, x, grad_fn, step, data, W, H, l1, l2, constr="nonneg"):
        """pgd.jl:224-255."""
        g = grad_fn(self._loss_grad(data))
        g = g + 2.0 * l2 * x + l1 * np.sign(x)            # SquarePenalty :77-79, AbsolutePenalty :86-88
        alpha = step / (np.linalg.norm(g) + EPSILON)      # :236
        x -= alpha * g                                    # :239
        self._project(x, constr)                          # :241
        self.est = tensor_conv(W, H)                      # :244
        loss = self._loss_eval(data)                      # :245
        step *= self.step_incr if loss < self.cur_loss else self.step_decr   # :247-251
        self.cur_loss = loss
        return step

    def update_motifs(self, data, W, H, l1W=0.0, l2W=1.0, **_):
        """pgd.jl:158-178; gradient pgd.jl:206-214."""
        L = W.shape[2]
        self.stepW = self._pgd(W, lambda ge: corr_w(H, ge, L), self.stepW, data, W, H, l1W, l2W, self.constrW)

    def update_feature_maps(self, data, W, H, l1H=0.0, l2H=0.0, **_):
        """pgd.jl:181-203; gradient pgd.jl:218-221."""
        self.stepH = self._pgd(H, lambda ge: tensor_transconv(W, ge), self.stepH, data, W, H, l1H, l2H, self.constrH)
        return float(np.sqrt(self.cur_loss / self.datanorm ** 2))


# --------------------------------------------------------------------------------------
# Alternating driver and public entry  (src/algs/alternating.jl, src/model.jl)
# --------------------------------------------------------------------------------------
class CNMFResults:
    """src/model.jl:11-17."""

    def __init__(self, data, W, H, time_hist, loss_hist):
        self.data, self.W, self.H = data, W, H
        self.time_hist, self.loss_hist = time_hist, loss_hist


def fit(rule, data, W_init, H_init, max_itr=100, max_time=np.inf, *, verbose=False,
        printer=print, **kwargs):
    """src/algs/alternating.jl:16-71."""
    check_convergence = kwargs.get("check_convergence", True)
    patience = kwargs.get("patience", 3)
    eval_mode = kwargs.get("eval_mode", False)
    assert patience >= 1
    tol = kwargs.get("tol", 1e-4)

    W = np.array(W_init, dtype=np.float64, copy=True)
    H = np.array(H_init, dtype=np.float64, copy=True)
    loss_hist = [compute_loss(data, W, H)]
    time_hist = [0.0]

    itr = 1
    while itr <= max_itr and time_hist[-1] <= max_time:
        itr += 1
        t0 = _time.time()
        if not eval_mode:
            rule.update_motifs(data, W, H, **kwargs)
        loss = rule.update_feature_maps(data, W, H, **kwargs)
        time_hist.append(time_hist[-1] + (_time.time() - t0))
        loss_hist.append(loss)
        if check_convergence and converged(loss_hist, patience, tol):
            printer("Converged early.")  # alternating.jl:64 (unconditional)
            break
    return CNMFResults(data, W, H, time_hist, loss_hist)


_ALGS = {"mult": MultUpdate, "hals": HALSUpdate, "pgd": PGDUpdate}


def fit_cnmf(data, L=10, K=5, alg="mult", max_itr=100, max_time=np.inf, seed=None,
             W_init=None, H_init=None, **kwargs):
    """src/model.jl:58-85.  ``alg`` is "mult"/"hals" (README symbols) or a rule class."""
    data = np.asarray(data, dtype=np.float64)
    rng = np.random.default_rng(seed)
    W0, H0 = init_rand(data, L, K, rng)
    if W_init is not None:
        W0 = np.asarray(W_init, dtype=np.float64)
    if H_init is not None:
        H0 = np.asarray(H_init, dtype=np.float64)
    rule_cls = _ALGS[alg] if isinstance(alg, str) else alg
    rule = rule_cls(data, W0, H0)
    return fit(rule, data, W0, H0, max_itr, max_time, **kwargs)


# --------------------------------------------------------------------------------------
# Data model  (datasets/synthetic.jl, datasets/toy.jl)
# --------------------------------------------------------------------------------------
def synthetic_sequences(K=3, N=100, L=20, T=500, alpha=0.1, p_h=0.5, sigma=0.2,
                        noise_scale=1.0, rng=None):
    """datasets/synthetic.jl:29-61 restated with a numpy Generator (draw order: Dirichlet
    weights, centres, H magnitudes, H mask, noise).  Returns (data, W[K,N,L], H[K,T])."""
    rng = np.random.default_rng(1234) if rng is None else rng
    mW = rng.dirichlet(np.full(K, alpha), size=N)  # N x K
    W = np.repeat(mW.T[:, :, None], L, axis=2)  # K x N x L
    lag_axis = np.linspace(-1.0, 1.0, L)
    cent = rng.uniform(-1.0, 1.0, size=(K, N))
    pdf = np.exp(-0.5 * ((lag_axis[None, None, :] - cent[:, :, None]) / sigma) ** 2) / (
        sigma * np.sqrt(2 * np.pi)
    )
    W = W * pdf
    H = rng.exponential(1.0, size=(K, T)) * (rng.random((K, T)) < p_h)
    noise = rng.normal(0.0, noise_scale, size=(N, T))
    data = np.maximum(0.0, tensor_conv(W, H) + noise)
    return data, W, H


def toy_data():
    """datasets/toy.jl:5-48 (N=7, T=250, K=2, L=5; small-integer W, sparse H) in K x N x L."""
    N, T, K, L = 7, 50, 2, 5
    W = np.zeros((K, N, L))
    W[0] = np.array([[0, 0, 0, 0, 0], [0, 0, 0, 0, 0], [0, 4, 0, 4, 0], [0, 0, 0, 0, 0],
                     [3, 0, 0, 0, 3], [0, 2, 1, 2, 0], [0, 0, 0, 0, 0]], dtype=float)
    W[1] = np.array([[0, 1, 0, 0, 0], [0, 0, 2, 0, 0], [0, 0, 0, 3, 0], [0, 5, 0, 0, 0],
                     [0, 0, 5, 0, 0], [0, 0, 0, 5, 0], [0, 0, 0, 0, 0]], dtype=float)
    H = np.zeros((K, T))
    H[0, 1], H[0, 19], H[0, 31], H[0, 47] = 1, 2, 1, 1
    H[1, 11], H[1, 29], H[1, 37] = 1, 1, 0.5
    H = np.tile(H, (1, 5))
    return tensor_conv(W, H), W, H


# --------------------------------------------------------------------------------------
# Layout helpers for the C ABI (Julia column-major)
# --------------------------------------------------------------------------------------
def to_julia(a):
    """NumPy array -> Fortran-ordered copy (what a Julia Array looks like in memory)."""
    return np.asfortranarray(a)

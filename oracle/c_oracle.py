"""ctypes loader for oracle/cnmf_oracle.c  --  TEST INFRASTRUCTURE ONLY.

Same contract as oracle/cnmf_oracle.py: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this.  Arrays cross in Julia column-major
order (Fortran-ordered float64)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_D = ctypes.POINTER(ctypes.c_double)
_I = ctypes.c_int64


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    src = os.path.join(_HERE, "cnmf_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liborc.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.orc_tensor_conv.argtypes = [_D, _D, _D, _I, _I, _I, _I]
        L.orc_tensor_transconv.argtypes = [_D, _D, _D, _I, _I, _I, _I]
        L.orc_corr_w.argtypes = [_D, _D, _D, _I, _I, _I, _I]
        L.orc_norm.argtypes = [_D, _I]
        L.orc_norm.restype = ctypes.c_double
        L.orc_compute_loss.argtypes = [_D, _D, _D, _D, _I, _I, _I, _I]
        L.orc_compute_loss.restype = ctypes.c_double
        L.orc_mu_update_motifs.argtypes = [_D, _D, _D, _D, _D, _D, _I, _I, _I, _I,
                                           ctypes.c_double, ctypes.c_double]
        L.orc_mu_update_feature_maps.argtypes = [_D, _D, _D, _D, _D, _D, _I, _I, _I, _I,
                                                 ctypes.c_double, ctypes.c_double, ctypes.c_double]
        L.orc_mu_update_feature_maps.restype = ctypes.c_double
        L.orc_hals_init.argtypes = [_D, _D, _D, _D, _I, _I, _I, _I]
        L.orc_hals_update_motifs.argtypes = [_D, _D, _D, _I, _I, _I, _I,
                                             ctypes.c_double, ctypes.c_double]
        L.orc_hals_update_feature_maps.argtypes = [_D, _D, _D, _I, _I, _I, _I, ctypes.c_double,
                                                   ctypes.c_double, ctypes.c_double]
        L.orc_hals_update_feature_maps.restype = ctypes.c_double
        _LIB = L
    return _LIB


def _p(a):
    assert a.dtype == np.float64 and a.flags.f_contiguous
    return a.ctypes.data_as(_D)


def _f(a):
    return np.asfortranarray(a, dtype=np.float64)


def tensor_conv(W, H):
    K, N, L = W.shape
    T = H.shape[1]
    W, H = _f(W), _f(H)
    est = np.zeros((N, T), order="F")
    lib().orc_tensor_conv(_p(W), _p(H), _p(est), N, T, K, L)
    return est


def tensor_transconv(W, X):
    K, N, L = W.shape
    T = X.shape[1]
    W, X = _f(W), _f(X)
    out = np.zeros((K, T), order="F")
    lib().orc_tensor_transconv(_p(W), _p(X), _p(out), N, T, K, L)
    return out


def corr_w(H, X, L):
    K, T = H.shape
    N = X.shape[0]
    H, X = _f(H), _f(X)
    out = np.zeros((K, N, L), order="F")
    lib().orc_corr_w(_p(H), _p(X), _p(out), N, T, K, L)
    return out


class _Rule:
    def __init__(self, data, W, H):
        self.data = _f(data)
        self.N, self.T = data.shape
        self.K, _, self.L = W.shape
        self.data_norm = float(lib().orc_norm(_p(self.data), self.N * self.T))


class MultUpdate(_Rule):
    """oracle/cnmf_oracle.c orc_mu_* (src/algs/mult.jl).  W, H must be Fortran-ordered float64
    arrays and are updated in place."""

    def __init__(self, data, W, H):
        super().__init__(data, W, H)
        self.est = np.zeros((self.N, self.T), order="F")
        self.numW = np.zeros(W.shape, order="F")
        self.denW = np.zeros(W.shape, order="F")
        self.numH = np.zeros(H.shape, order="F")
        self.denH = np.zeros(H.shape, order="F")

    def update_motifs(self, data, W, H, l1W=0.0, l2W=0.0, **_):
        lib().orc_mu_update_motifs(_p(self.data), _p(W), _p(H), _p(self.est), _p(self.numW),
                                   _p(self.denW), self.N, self.T, self.K, self.L, l1W, l2W)

    def update_feature_maps(self, data, W, H, l1H=0.0, l2H=0.0, **_):
        return float(lib().orc_mu_update_feature_maps(
            _p(self.data), _p(W), _p(H), _p(self.est), _p(self.numH), _p(self.denH),
            self.N, self.T, self.K, self.L, l1H, l2H, self.data_norm))


class HALSUpdate(_Rule):
    """oracle/cnmf_oracle.c orc_hals_* (src/algs/hals.jl), literal sweeps on a persistent residual."""

    def __init__(self, data, W, H):
        super().__init__(data, W, H)
        self.resids = np.zeros((self.N, self.T), order="F")
        lib().orc_hals_init(_p(self.data), _p(_f(W)), _p(_f(H)), _p(self.resids),
                            self.N, self.T, self.K, self.L)

    def update_motifs(self, data, W, H, l1W=0.0, l2W=0.0, **_):
        lib().orc_hals_update_motifs(_p(self.resids), _p(W), _p(H), self.N, self.T, self.K,
                                     self.L, l1W, l2W)

    def update_feature_maps(self, data, W, H, l1H=0.0, l2H=0.0, **_):
        return float(lib().orc_hals_update_feature_maps(
            _p(self.resids), _p(W), _p(H), self.N, self.T, self.K, self.L, l1H, l2H,
            self.data_norm))


def fit(rule_cls, data, W_init, H_init, max_itr=100, **kwargs):
    """Drive the C rules with the oracle's alternating loop (src/algs/alternating.jl:16-71)."""
    from . import cnmf_oracle as po

    class _Wrap:
        def __init__(self, r):
            self.r = r

        def update_motifs(self, data, W, H, **kw):
            self.r.update_motifs(data, W, H, **kw)

        def update_feature_maps(self, data, W, H, **kw):
            return self.r.update_feature_maps(data, W, H, **kw)

    data = _f(data)
    W0, H0 = _f(W_init), _f(H_init)
    rule = rule_cls(data, W0, H0)
    # po.fit copies W/H with np.array(copy=True), which keeps Fortran order for F-inputs
    return po.fit(_Wrap(rule), data, W0, H0, max_itr, **kwargs)

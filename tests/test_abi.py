"""CPU-side checks of the C-ABI boundary: the library builds, loads and exports exactly the
symbols include/cmf_sm100.h declares; argument validation works without a device; and the
compute entry points fail loudly (no CPU fallback) when no GPU is visible."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge

    ge.build()
    from cmf_jl_b200 import _lib

    return _lib


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "cmf_sm100.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cmf_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    so = ctypes.CDLL(lib.SO_PATH)
    syms = _header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(so, s), f"{s} declared in include/cmf_sm100.h but not exported"
    bound = set(lib.SIGNATURES) | {"cmf_last_error"}
    assert set(syms) == bound, set(syms) ^ bound


def test_no_oracle_import_in_product():
    # the product package must never route through the CPU oracle
    pkg = os.path.join(ROOT, "cmf.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "liborc" not in src, f


def test_compute_fails_loudly_without_gpu(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import cmf_jl_b200 as cmf

    with pytest.raises(cmf.CMFError) as ei:
        cmf.tensor_conv(np.ones((2, 3, 2)), np.ones((2, 5)))
    assert ei.value.code == 2 and "no CPU fallback" in str(ei.value)
    with pytest.raises(cmf.CMFError):
        cmf.fit_cnmf(np.ones((4, 20)), L=3, K=2, max_itr=2)


def test_argument_validation(lib):
    L = lib.load()
    h = ctypes.c_void_p()
    assert L.cmf_create(ctypes.byref(h), 4, 3, 2, 5, 0, 0, 0) == 1   # L > T
    assert b"L <= T" in L.cmf_last_error()
    assert L.cmf_create(ctypes.byref(h), 0, 10, 2, 5, 0, 0, 0) == 1
    assert L.cmf_create(ctypes.byref(h), 4, 10, 2, 5, 7, 0, 0) == 1  # bad dtype
    assert L.cmf_create_shard(ctypes.byref(h), 4, 100, 0, 2, 2, 5, 0, 0, 0) == 1   # shard < L-1
    assert L.cmf_create_shard(ctypes.byref(h), 4, 100, 0, 50, 2, 5, 0, 2, 0) == 3  # sharded PGD
    assert L.cmf_update_motifs(None, 0.0, 0.0) == 1
    assert L.cmf_destroy(None) == 0


def test_kwarg_normalisation_and_alg_resolution():
    import warnings

    from cmf_jl_b200 import model

    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        kw = model._normalise_kwargs(dict(l1_W=1.0, l2_H=2.0, initW="w", bogus=3))
    assert kw == {"l1W": 1.0, "l2H": 2.0, "W_init": "w"}
    assert any("bogus" in str(x.message) for x in w)
    assert model._resolve_alg(":mult") is model.MultUpdate
    assert model._resolve_alg("hals") is model.HALSUpdate
    assert model._resolve_alg(model.HALSUpdate) is model.HALSUpdate
    with pytest.raises(ValueError):
        model._resolve_alg("anls")
    assert model.converged([1, 1, 1, 1], 3, 1e-4) and not model.converged([1, 1, 1], 3, 1e-4)


def test_save_load_round_trip(tmp_path):
    # src/model.jl:149-181 with the real struct fields
    import cmf_jl_b200 as cmf

    rng = np.random.default_rng(0)
    r = cmf.CNMF_results(rng.random((4, 9)), rng.random((3, 4, 2)), rng.random((2, 9)), [0.0, 0.1], [0.9, 0.5], "LNK")
    p = tmp_path / "model.npz"
    cmf.save_model(r, p)
    q = cmf.load_model(p)
    assert np.array_equal(q.W, r.W) and np.array_equal(q.H, r.H) and np.array_equal(q.data, r.data)
    assert q.loss_hist == r.loss_hist and q.time_hist == r.time_hist and q.layout == "LNK"
    assert cmf.num_lags(q) == 3 and cmf.num_units(q) == 4 and cmf.num_components(q) == 2 and cmf.num_iter(q) == 2

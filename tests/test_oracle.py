"""Pins the CPU oracle (oracle/) with everything the reference offers for this path.

The reference has no tests or golden vectors, so the pins are: the
element-wise definition from notebooks/benchmarks.ipynb (tconv4), adjointness, the tconv3
shift-and-stack identity, agreement between the NumPy and plain-C restatements, the algebraic
Gram/recurrence forms the CUDA kernels use, MU monotonicity, driver/convergence semantics
(src/algs/alternating.jl, src/model.jl:91-107) and the committed golden fixtures.
"""
import os

import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import cnmf_oracle as po
from oracle import restructured as rs

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _rand(N, T, K, L, seed=0):
    rng = np.random.default_rng(seed)
    return rng.random((K, N, L)), rng.random((K, T)), rng.random((N, T))


@pytest.mark.parametrize("dims", [(13, 57, 3, 5), (4, 6, 2, 6), (5, 3, 2, 7), (1, 1, 1, 1)])
def test_definitions_match_naive(dims):
    N, T, K, L = dims
    W, H, X = _rand(*dims)
    assert np.allclose(po.tensor_conv(W, H), po.naive_conv(W, H), atol=1e-13)
    assert np.allclose(po.tensor_transconv(W, X), po.naive_transconv(W, X), atol=1e-13)
    assert np.allclose(po.corr_w(H, X, L), po.naive_corr_w(H, X, L), atol=1e-13)


def test_c_twin_matches_numpy():
    W, H, X = _rand(31, 203, 4, 9, seed=3)
    assert np.allclose(co.tensor_conv(W, H), po.tensor_conv(W, H), atol=1e-12)
    assert np.allclose(co.tensor_transconv(W, X), po.tensor_transconv(W, X), atol=1e-12)
    assert np.allclose(co.corr_w(H, X, 9), po.corr_w(H, X, 9), atol=1e-11)


def test_adjoint_identities():
    W, H, X = _rand(17, 91, 3, 8, seed=1)
    a = np.vdot(po.tensor_conv(W, H), X)
    b = np.vdot(H, po.tensor_transconv(W, X))
    c = np.vdot(W, po.corr_w(H, X, 8))
    assert abs(a - b) < 1e-10 * abs(a) and abs(a - c) < 1e-10 * abs(a)


def test_shift_and_stack_identity():
    # notebooks/benchmarks.ipynb tconv3: conv = W_unf * shift_and_stack(H)
    W, H, _ = _rand(11, 64, 3, 6, seed=2)
    Wu = rs.unfold_W(W)  # KL x N
    assert np.allclose(Wu.T @ po.shift_and_stack(H, 6), po.tensor_conv(W, H), atol=1e-13)


def test_toy_data_exact_integers():
    # datasets/toy.jl: small-integer W, sparse H -> conv is exact in floating point
    X, W, H = po.toy_data()
    assert X.shape == (7, 250)
    assert np.array_equal(X, po.naive_conv(W, H))
    assert np.array_equal(co.tensor_conv(W, H), X)
    assert np.array_equal(X * 2, np.round(X * 2))


@pytest.mark.parametrize("dims", [(11, 64, 3, 6), (7, 9, 2, 5), (6, 5, 2, 5)])
def test_gram_forms_equal_literal(dims):
    N, T, K, L = dims
    W, H, X = _rand(*dims, seed=5)
    est = po.tensor_conv(W, H)
    assert np.allclose(rs.denomW_gram(W, H), po.corr_w(H, est, L), rtol=1e-12, atol=1e-12)
    assert np.allclose(rs.denomH_gram(W, H), po.tensor_transconv(W, est), rtol=1e-12, atol=1e-12)
    assert np.allclose(rs.build_G(H, L), po.shift_and_stack(H, L) @ po.shift_and_stack(H, L).T,
                       atol=1e-12)
    assert abs(rs.loss_expansion(X, W, H) - po.compute_loss(X, W, H)) < 1e-12


def test_mu_gram_iteration_equals_literal():
    N, T, K, L = 23, 120, 3, 7
    W, H, X = _rand(N, T, K, L, seed=7)
    reg = dict(l1W=0.1, l2W=0.5, l1H=0.1, l2H=0.2)
    Wl, Hl = W.copy(), H.copy()
    rule = po.MultUpdate(X, Wl, Hl)
    Wg, Hg = W.copy(), H.copy()
    for _ in range(5):
        rule.update_motifs(X, Wl, Hl, **reg)
        loss_l = rule.update_feature_maps(X, Wl, Hl, **reg)
        Wg, Hg, loss_g = rs.mu_iteration_gram(X, Wg, Hg, **reg)
        assert abs(loss_l - loss_g) < 1e-12
    assert np.allclose(Wl, Wg, rtol=1e-11) and np.allclose(Hl, Hg, rtol=1e-11)


def test_mu_c_twin_equals_numpy():
    N, T, K, L = 19, 150, 4, 6
    W, H, X = _rand(N, T, K, L, seed=8)
    reg = dict(l1W=0.05, l2W=0.1, l1H=0.02, l2H=0.3)
    rp = po.fit(po.MultUpdate(X, W, H), X, W, H, 12, check_convergence=False, **reg)
    rc = co.fit(co.MultUpdate, X, W, H, 12, check_convergence=False, **reg)
    assert np.allclose(rp.loss_hist, rc.loss_hist, rtol=1e-12)
    assert np.allclose(rp.W, rc.W, rtol=1e-10) and np.allclose(rp.H, rc.H, rtol=1e-10)


def test_hals_c_twin_and_gram_equal_literal():
    N, T, K, L = 11, 64, 3, 6
    W, H, X = _rand(N, T, K, L, seed=9)
    reg = dict(l1W=0.1, l2W=0.5, l1H=0.1, l2H=0.2)
    rp = po.fit(po.HALSUpdate(X, W, H), X, W, H, 6, check_convergence=False, **reg)
    rc = co.fit(co.HALSUpdate, X, W, H, 6, check_convergence=False, **reg)
    assert np.allclose(rp.loss_hist, rc.loss_hist, rtol=1e-12)
    assert np.allclose(rp.W, rc.W, atol=1e-12) and np.allclose(rp.H, rc.H, atol=1e-12)
    Wg, Hg = W.copy(), H.copy()
    for i in range(6):
        Wg, Hg, loss = rs.hals_iteration_gram(X, Wg, Hg, **reg)
        assert abs(loss - rp.loss_hist[i + 1]) < 1e-12
    assert np.allclose(Wg, rp.W, atol=1e-12) and np.allclose(Hg, rp.H, atol=1e-12)


def test_hals_short_T_tail():
    # T < 2L: every column of H is in the truncated tail (hals.jl:137 w = min(T-t+1, L))
    N, T, K, L = 5, 7, 2, 5
    W, H, X = _rand(N, T, K, L, seed=10)
    rp = po.fit(po.HALSUpdate(X, W, H), X, W, H, 3, check_convergence=False)
    Wg, Hg = W.copy(), H.copy()
    for i in range(3):
        Wg, Hg, loss = rs.hals_iteration_gram(X, Wg, Hg)
        assert abs(loss - rp.loss_hist[i + 1]) < 1e-12


def test_mu_monotone_on_synthetic():
    data, _, _ = po.synthetic_sequences(K=3, N=60, L=8, T=300, rng=np.random.default_rng(1234))
    r = po.fit_cnmf(data, L=8, K=3, alg="mult", max_itr=40, seed=0, check_convergence=False)
    lh = np.asarray(r.loss_hist)
    assert len(lh) == 41 and np.all(np.diff(lh) <= 1e-12)


def test_converged_semantics():
    # src/model.jl:91-107
    assert not po.converged([1.0, 1.0, 1.0], 3, 1e-4)  # length <= patience
    assert po.converged([1.0, 1.0, 1.0, 1.0], 3, 1e-4)
    assert not po.converged([1.0, 0.9, 0.9, 0.9], 3, 1e-4)
    assert po.converged([5.0, 0.9, 0.90001, 0.90002, 0.90001], 3, 1e-4)


def test_driver_history_and_early_stop():
    data, _, _ = po.synthetic_sequences(K=2, N=20, L=4, T=80, rng=np.random.default_rng(1))
    msgs = []
    r = po.fit_cnmf(data, L=4, K=2, alg="mult", max_itr=500, seed=0, tol=1e-3, printer=msgs.append)
    assert len(r.loss_hist) == len(r.time_hist) < 501 and r.time_hist[0] == 0.0
    assert msgs == ["Converged early."]
    r2 = po.fit_cnmf(data, L=4, K=2, alg="mult", max_itr=0, seed=0)
    assert len(r2.loss_hist) == 1
    # eval_mode skips the motif update (alternating.jl:51-53)
    W0 = np.random.default_rng(3).random((2, 20, 4))
    r3 = po.fit_cnmf(data, L=4, K=2, max_itr=3, W_init=W0, eval_mode=True, check_convergence=False)
    assert np.array_equal(r3.W, W0)


def test_init_rand_alpha_is_least_squares_scale():
    data, _, _ = po.synthetic_sequences(K=2, N=15, L=4, T=60, rng=np.random.default_rng(2))
    W, H = po.init_rand(data, 4, 2, np.random.default_rng(0))
    est = po.tensor_conv(W, H)
    # after the rescale, <data, est> == ||est||^2  (alpha == 1)
    assert abs(np.vdot(data, est) / np.vdot(est, est) - 1.0) < 1e-12


@pytest.mark.parametrize("name", ["mu_small", "mu_reg_small", "hals_small", "hals_reg_small", "prims"])
def test_golden_fixtures(name):
    path = os.path.join(GOLD, name + ".npz")
    g = np.load(path)
    if name == "prims":
        assert np.allclose(po.tensor_conv(g["W"], g["H"]), g["conv"], rtol=1e-13, atol=1e-13)
        assert np.allclose(po.tensor_transconv(g["W"], g["X"]), g["transconv"], rtol=1e-13, atol=1e-13)
        assert np.allclose(po.corr_w(g["H"], g["X"], g["W"].shape[2]), g["corr"], rtol=1e-13, atol=1e-12)
        return
    reg = {k: float(g[k]) for k in ("l1W", "l2W", "l1H", "l2H")}
    alg = str(g["alg"])
    rule = {"mult": co.MultUpdate, "hals": co.HALSUpdate}[alg]
    r = co.fit(rule, g["X"], g["W0"], g["H0"], int(g["max_itr"]), check_convergence=False, **reg)
    assert np.allclose(r.loss_hist, g["loss_hist"], rtol=1e-11)
    assert np.allclose(r.W, g["W"], rtol=1e-9, atol=1e-12)
    assert np.allclose(r.H, g["H"], rtol=1e-9, atol=1e-12)


def test_pgd_gradients_are_the_derivatives_of_the_square_loss():
    # pgd.jl:206-221: dW = corr(H, 2(est - X)), dH = transconv(W, 2(est - X)); checked by central differences
    N, T, K, L = 6, 17, 2, 4
    W, H, X = _rand(N, T, K, L, seed=12)
    f = lambda W_, H_: float(np.sum((po.tensor_conv(W_, H_) - X) ** 2))
    ge = 2.0 * (po.tensor_conv(W, H) - X)
    gW, gH = po.corr_w(H, ge, L), po.tensor_transconv(W, ge)
    rng = np.random.default_rng(0)
    for _ in range(5):
        k, n, l, t = rng.integers(K), rng.integers(N), rng.integers(L), rng.integers(T)
        e = 1e-6
        Wp, Wm = W.copy(), W.copy()
        Wp[k, n, l] += e
        Wm[k, n, l] -= e
        assert abs((f(Wp, H) - f(Wm, H)) / (2 * e) - gW[k, n, l]) < 1e-6 * max(1.0, abs(gW[k, n, l]))
        Hp, Hm = H.copy(), H.copy()
        Hp[k, t] += e
        Hm[k, t] -= e
        assert abs((f(W, Hp) - f(W, Hm)) / (2 * e) - gH[k, t]) < 1e-6 * max(1.0, abs(gH[k, t]))


def test_pgd_driver_semantics():
    data, _, _ = po.synthetic_sequences(K=2, N=20, L=4, T=80, rng=np.random.default_rng(1))
    W0, H0 = po.init_rand(data, 4, 2, np.random.default_rng(0))
    rule = po.PGDUpdate(data, W0, H0)
    assert rule.stepW == 5.0 and rule.cur_loss == np.linalg.norm(data)      # pgd.jl:147-149
    r = po.fit(rule, data, W0, H0, 25, check_convergence=False)
    assert len(r.loss_hist) == 26 and r.loss_hist[-1] < r.loss_hist[0]
    assert np.all(r.W >= po.EPSILON) and np.all(r.H >= po.EPSILON)          # NonnegConstraint floor


# ---------------------------------------------------------------------------------------------
# frequency-domain (overlap-save) forms of the device engine == the literal restatement
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dims", [(13, 300, 3, 5), (8, 157, 4, 16), (6, 90, 2, 1), (5, 700, 3, 30)])
def test_overlap_save_forms_equal_literal(dims):
    from oracle import restructured as rs

    N, T, K, L = dims
    rng = np.random.default_rng(sum(dims))
    W, H, X = rng.random((K, N, L)), rng.random((K, T)), rng.random((N, T))
    for B in (None, 2 * rs.fd_block_length(L)):
        assert np.allclose(rs.numH_overlap_save(W, X, B), po.tensor_transconv(W, X), rtol=1e-10, atol=1e-10)
        assert np.allclose(rs.numW_overlap_save(H, X, L, B), po.corr_w(H, X, L), rtol=1e-10, atol=1e-10)
        assert np.allclose(rs.gram_overlap_save(H, L, B), rs.gram_R(H, L), rtol=1e-10, atol=1e-10)
        lit = rs.denomH_gram(W, H)
        fd = rs.denomH_interior_overlap_save(W, H, B)
        assert np.allclose(fd[:, : T - (L - 1)], lit[:, : T - (L - 1)], rtol=1e-10, atol=1e-10)
        assert np.allclose(rs.conv_overlap_save(W, H, B), po.tensor_conv(W, H), rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("reg", [{}, dict(l1W=0.1, l2W=0.5, l1H=0.1, l2H=0.2)])
def test_mu_iteration_in_the_frequency_domain_equals_literal(reg):
    # the whole iteration of the device's frequency-domain engine (all five overlap-save pieces) == src/algs/mult.jl
    from oracle import restructured as rs

    N, T, K, L = 11, 260, 3, 7
    X, _, _ = po.synthetic_sequences(K=2, N=N, L=L, T=T, rng=np.random.default_rng(5))
    W, H = po.init_rand(X, L, K, np.random.default_rng(1))
    rule = po.MultUpdate(X, W, H)
    Wl, Hl = W.copy(), H.copy()
    Wf, Hf = W.copy(), H.copy()
    for _ in range(4):
        rule.update_motifs(X, Wl, Hl, **{k: v for k, v in reg.items() if k.endswith("W")})
        loss_l = rule.update_feature_maps(X, Wl, Hl, **{k: v for k, v in reg.items() if k.endswith("H")})
        Wf, Hf, loss_f = rs.mu_iteration_overlap_save(X, Wf, Hf, **reg)
        assert abs(loss_l - loss_f) < 1e-11
    assert np.allclose(Wl, Wf, rtol=1e-9, atol=1e-13) and np.allclose(Hl, Hf, rtol=1e-9, atol=1e-13)


# ---- exact-arithmetic pin (oracle/exact_pin.py): rational restatement written from the Julia sources ----------------------
def _exact_case(tiles):
    from fractions import Fraction as Fr

    from oracle import exact_pin as ex

    X, Wt, Ht = ex.toy_data_exact(tiles)
    N, T, K, L = ex.dims(Wt, Ht)
    W0, H0 = ex.rational_init(K, N, L, T)
    reg_x = dict(l1W=Fr(1, 10), l2W=Fr(1, 2), l1H=Fr(1, 10), l2H=Fr(1, 5))
    f = lambda a: np.asfortranarray(np.asarray(ex.to_float(a), dtype=np.float64))
    return ex, X, W0, H0, reg_x, f


def test_exact_pin_toy_data_matches_oracle_toy():
    from oracle import exact_pin as ex

    X, W, H = ex.toy_data_exact(5)
    Xo, Wo, Ho = po.toy_data()
    assert np.array_equal(np.asarray(ex.to_float(X)), Xo) and np.array_equal(np.asarray(ex.to_float(W)), Wo)
    assert np.array_equal(np.asarray(ex.to_float(H)), Ho)


@pytest.mark.parametrize("oracle_name", ["numpy", "c"])
def test_exact_pin_mu_iteration(oracle_name):
    # one MultUpdate iteration (mult.jl:23-58) in exact rationals vs the floating-point oracles: rounding error only
    ex, X, W0, H0, reg_x, f = _exact_case(2)
    We, He, le = ex.mu_iteration(X, W0, H0, **reg_x)
    o = po if oracle_name == "numpy" else co
    Xf, W, H = f(X), f(W0), f(H0)
    rule = o.MultUpdate(Xf, W, H)
    rule.update_motifs(Xf, W, H, l1W=0.1, l2W=0.5)
    loss = rule.update_feature_maps(Xf, W, H, l1H=0.1, l2H=0.2)
    assert abs(loss - le) < 1e-13 * le
    assert np.max(np.abs(W - f(We)) / f(We)) < 1e-13 and np.max(np.abs(H - f(He)) / f(He)) < 1e-13


@pytest.mark.parametrize("oracle_name", ["numpy", "c"])
def test_exact_pin_hals_iteration(oracle_name):
    # one HALSUpdate iteration (hals.jl:18-154, persistent residual, both sweeps) in exact rationals vs the oracles
    ex, X, W0, H0, reg_x, f = _exact_case(5)
    We, He, Re, le = ex.hals_iteration(X, W0, H0, None, **reg_x)
    o = po if oracle_name == "numpy" else co
    Xf, W, H = f(X), f(W0), f(H0)
    rule = o.HALSUpdate(Xf, W, H)
    rule.update_motifs(Xf, W, H, l1W=0.1, l2W=0.5)
    loss = rule.update_feature_maps(Xf, W, H, l1H=0.1, l2H=0.2)
    assert abs(loss - le) < 1e-13 * le
    assert np.max(np.abs(W - f(We))) < 1e-13 and np.max(np.abs(H - f(He))) < 1e-12
    assert (f(He) == 0).sum() > 0            # the clamp at zero (hals.jl:153) is exercised: exact zeros stay exact zeros
    assert np.array_equal(H == 0, f(He) == 0)


def test_exact_pin_fixture_is_current():
    # tests/golden/exact_pin_toy.npz (what the GPU tests compare with) is what oracle/exact_pin.py produces
    ex, X, W0, H0, reg_x, f = _exact_case(5)
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "exact_pin_toy.npz"))
    assert np.array_equal(g["X"], f(X)) and np.array_equal(g["W0"], f(W0)) and np.array_equal(g["H0"], f(H0))
    We, He, Re, le = ex.hals_iteration(X, W0, H0, None, **reg_x)
    assert np.array_equal(g["hals_W"], f(We)) and np.array_equal(g["hals_H"], f(He)) and float(g["hals_loss"]) == le


# ---- round schedule of the second-generation HALS H sweep (cmf.jl_b200/csrc/kernels_hals.cuh) replayed on the CPU ------------
@pytest.mark.parametrize("dims", [(6, 150, 11, 4, 16, 4), (5, 97, 9, 5, 8, 2), (4, 64, 3, 1, 16, 8), (5, 200, 20, 6, 32, 8),
                                  (4, 40, 5, 7, 8, 4), (4, 300, 17, 9, 8, 8)])
def test_hals2_round_schedule_equals_sequential_sweep(dims):
    # every read of the replay is checked against the round its data was written in (bulk-synchronous rounds: strictly
    # earlier), ring slots against the chunk they hold; the result must be the k-outer / t-inner sweep of hals.jl:121-154
    from oracle import hals2_schedule as h2

    N, T, K, L, CW, GS = dims
    rng = np.random.default_rng(N + T + K)
    W = rng.random((K, N, L))
    H = rng.random((K, T)) * (rng.random((K, T)) < 0.6)
    X = rng.random((N, T))
    R = po.tensor_conv(W, H) - X
    ref = rs.hals_H_sweep_gram(R, W, H, 0.05, 0.1)
    got, n_rounds = h2.sweep(po.tensor_transconv(W, R), H, W, 0.05, 0.1, CW=CW, GS=GS, STAG=4)
    assert n_rounds == -(-T // CW) + 4 * (K - 1) + 3
    assert np.max(np.abs(got - ref)) < 1e-12
    assert np.array_equal(got == 0, ref == 0)

"""GPU parity tests: the CUDA path (through the C ABI, via the Python mirror of the reference
interface) against the CPU oracle on the same seeded inputs, against the committed golden
fixtures, and through size-independent properties at larger sizes.

Tolerances (BASELINE.json north_star): fp64 kernels within 1e-9 relative of the Float64
reference arithmetic on W, H and loss_hist; fp32 within 1e-4 relative on the loss after 100
iterations."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
F64_RTOL = 1e-9
F32_LOSS_RTOL = 1e-4


@pytest.fixture(scope="module")
def cmf():
    import __graft_entry__ as ge

    ge.build()
    import cmf_jl_b200

    return cmf_jl_b200


@pytest.fixture(scope="module")
def orc():
    from oracle import c_oracle, cnmf_oracle, restructured

    class O:
        po, co, rs = cnmf_oracle, c_oracle, restructured

    return O


def _rand(N, T, K, L, seed=0):
    rng = np.random.default_rng(seed)
    return rng.random((K, N, L)), rng.random((K, T)), rng.random((N, T))


def _relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


PRIM_DIMS = [(13, 57, 3, 5), (1, 1, 1, 1), (5, 7, 2, 7), (130, 300, 20, 9), (500, 2000, 5, 10),
             (64, 257, 64, 33), (33, 140, 7, 1), (150, 1000, 15, 100), (17, 90, 130, 3)]


@pytest.mark.parametrize("dims", PRIM_DIMS)
@pytest.mark.parametrize("dtype,tol", [("f64", 1e-12), ("f32", 2e-5)])
def test_resids_and_shift_and_stack_match_oracle(cmf, orc, dims, dtype, tol):
    # compute_resids (src/common.jl:58-59) and shift_and_stack (:133-142) as callable primitives
    N, T, K, L = dims
    W, H, X = _rand(*dims, seed=sum(dims) + 1)
    ref = orc.po.compute_resids(X, W, H)
    got = cmf.compute_resids(X, W, H, dtype)
    assert np.max(np.abs(got - ref)) < tol * max(np.max(np.abs(orc.po.tensor_conv(W, H))), 1.0)
    Hs = cmf.shift_and_stack(H, L, dtype)
    assert Hs.shape == (K * L, T)
    assert np.array_equal(Hs, orc.po.shift_and_stack(H, L).astype(Hs.dtype))     # pure data movement: bit-exact
    # the tconv3 identity of notebooks/benchmarks.ipynb:96 on the device primitives: conv(W,H) = W_unf * Htilde
    W_unf = np.concatenate([W[:, :, l].T for l in range(L)], axis=1)             # N x (L*K), column l*K + k
    assert _relerr(W_unf @ Hs.astype(np.float64), orc.po.tensor_conv(W, H)) < 10 * tol


@pytest.mark.parametrize("dims", PRIM_DIMS)
@pytest.mark.parametrize("dtype,tol", [("f64", 1e-12), ("f32", 2e-5)])
def test_primitives_match_oracle(cmf, orc, dims, dtype, tol):
    N, T, K, L = dims
    W, H, X = _rand(*dims, seed=sum(dims))
    assert _relerr(cmf.tensor_conv(W, H, dtype), orc.co.tensor_conv(W, H)) < tol
    assert _relerr(cmf.tensor_transconv(W, X, dtype), orc.co.tensor_transconv(W, X)) < tol
    assert _relerr(cmf.corr_w(H, X, L, dtype), orc.co.corr_w(H, X, L)) < tol


def test_toy_data_is_bit_exact(cmf, orc):
    # datasets/toy.jl: small integers and halves -> every partial sum is exact in fp32 and fp64
    X, W, H = orc.po.toy_data()
    assert np.array_equal(cmf.tensor_conv(W, H, "f64"), X)
    assert np.array_equal(cmf.tensor_conv(W, H, "f32").astype(np.float64), X)


def test_golden_primitives(cmf):
    g = np.load(os.path.join(GOLD, "prims.npz"))
    assert _relerr(cmf.tensor_conv(g["W"], g["H"]), g["conv"]) < 1e-12
    assert _relerr(cmf.tensor_transconv(g["W"], g["X"]), g["transconv"]) < 1e-12
    assert _relerr(cmf.corr_w(g["H"], g["X"], g["W"].shape[2]), g["corr"]) < 1e-12


@pytest.mark.parametrize("name", ["mu_small", "mu_reg_small", "hals_small", "hals_reg_small"])
def test_golden_fits_fp64(cmf, name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    reg = {k: float(g[k]) for k in ("l1W", "l2W", "l1H", "l2H")}
    K, N, L = g["W0"].shape
    r = cmf.fit_cnmf(g["X"], L=L, K=K, alg=str(g["alg"]), max_itr=int(g["max_itr"]), W_init=g["W0"],
                     H_init=g["H0"], check_convergence=False, layout="KNL", **reg)
    assert len(r.loss_hist) == int(g["max_itr"]) + 1
    assert np.allclose(r.loss_hist, g["loss_hist"], rtol=F64_RTOL)
    assert np.allclose(r.W, g["W"], rtol=F64_RTOL, atol=1e-13)
    assert np.allclose(r.H, g["H"], rtol=F64_RTOL, atol=1e-13)


def _config1(orc, N=500, T=2000):
    # BASELINE.json configs[0]: synthetic N=500, T=2000, fit K=5 L=10 (datasets/synthetic.jl data model)
    X, _, _ = orc.po.synthetic_sequences(K=3, N=N, L=20, T=T, rng=np.random.default_rng(1234))
    W0, H0 = orc.po.init_rand(X, 10, 5, np.random.default_rng(0))
    return X, W0, H0


def test_config1_mu_fp64_100_iterations(cmf, orc):
    X, W0, H0 = _config1(orc)
    ref = orc.co.fit(orc.co.MultUpdate, X, W0, H0, 100, check_convergence=False)
    r = cmf.fit_cnmf(X, L=10, K=5, alg="mult", max_itr=100, W_init=W0, H_init=H0, check_convergence=False,
                     layout="KNL")
    assert len(r.loss_hist) == 101 and r.time_hist[0] == 0.0
    assert np.allclose(r.loss_hist, ref.loss_hist, rtol=F64_RTOL)
    assert np.allclose(r.W, ref.W, rtol=F64_RTOL, atol=1e-13)
    assert np.allclose(r.H, ref.H, rtol=F64_RTOL, atol=1e-13)
    assert np.all(np.diff(r.loss_hist) <= 1e-12)      # MU is monotone


def test_config1_mu_fp32_loss_within_1e4(cmf, orc):
    X, W0, H0 = _config1(orc)
    ref = orc.co.fit(orc.co.MultUpdate, X, W0, H0, 100, check_convergence=False)
    r = cmf.fit_cnmf(X, L=10, K=5, alg="mult", max_itr=100, W_init=W0, H_init=H0, check_convergence=False,
                     dtype="f32", layout="KNL")
    rel = np.abs(np.asarray(r.loss_hist) - np.asarray(ref.loss_hist)) / np.asarray(ref.loss_hist)
    assert rel[-1] < F32_LOSS_RTOL and rel.max() < F32_LOSS_RTOL, rel.max()


def test_config2_hals_regularised_fp64(cmf, orc):
    # BASELINE.json configs[1]: same data, HALS with l1_H=0.1, l2_H=0.2, l1_W=0.1, l2_W=0.5 (README.md:52)
    X, W0, H0 = _config1(orc)
    reg = dict(l1W=0.1, l2W=0.5, l1H=0.1, l2H=0.2)
    ref = orc.co.fit(orc.co.HALSUpdate, X, W0, H0, 25, check_convergence=False, **reg)
    r = cmf.fit_cnmf(X, L=10, K=5, alg=":hals", max_itr=25, W_init=W0, H_init=H0, check_convergence=False,
                     layout="KNL", l1_W=0.1, l2_W=0.5, l1_H=0.1, l2_H=0.2)      # README spellings
    assert np.allclose(r.loss_hist, ref.loss_hist, rtol=F64_RTOL)
    assert np.allclose(r.W, ref.W, rtol=1e-8, atol=1e-11)
    assert np.allclose(r.H, ref.H, rtol=1e-8, atol=1e-11)


def test_config2_hals_fp32_loss_within_1e4(cmf, orc):
    X, W0, H0 = _config1(orc, N=200, T=800)
    reg = dict(l1W=0.1, l2W=0.5, l1H=0.1, l2H=0.2)
    ref = orc.co.fit(orc.co.HALSUpdate, X, W0, H0, 30, check_convergence=False, **reg)
    r = cmf.fit_cnmf(X, L=10, K=5, alg="hals", max_itr=30, W_init=W0, H_init=H0, check_convergence=False,
                     dtype="f32", layout="KNL", **reg)
    rel = np.abs(np.asarray(r.loss_hist) - np.asarray(ref.loss_hist)) / np.asarray(ref.loss_hist)
    assert rel[-1] < F32_LOSS_RTOL, rel


def test_hals_short_T_all_tail(cmf, orc):
    N, T, K, L = 5, 7, 2, 5
    W, H, X = _rand(N, T, K, L, seed=10)
    ref = orc.co.fit(orc.co.HALSUpdate, X, W, H, 3, check_convergence=False)
    r = cmf.fit_cnmf(X, L=L, K=K, alg="hals", max_itr=3, W_init=W, H_init=H, check_convergence=False, layout="KNL")
    assert np.allclose(r.loss_hist, ref.loss_hist, rtol=F64_RTOL)
    ref = orc.co.fit(orc.co.MultUpdate, X, W, H, 3, check_convergence=False)
    r = cmf.fit_cnmf(X, L=L, K=K, alg="mult", max_itr=3, W_init=W, H_init=H, check_convergence=False, layout="KNL")
    assert np.allclose(r.loss_hist, ref.loss_hist, rtol=F64_RTOL)


def test_rule_interface_in_place_semantics(cmf, orc):
    # the plugin boundary: Rule(data,W,H); update_motifs!(...) mutates W; update_feature_maps!(...) -> loss
    N, T, K, L = 23, 120, 3, 7
    W, H, X = _rand(N, T, K, L, seed=7)
    Wo, Ho = np.asfortranarray(W.copy()), np.asfortranarray(H.copy())
    ro = orc.co.MultUpdate(X, Wo, Ho)
    Wg, Hg = W.copy(), H.copy()
    rg = cmf.MultUpdate(X, Wg, Hg)
    for _ in range(3):
        ro.update_motifs(X, Wo, Ho, l1W=0.1, l2W=0.2)
        rg.update_motifs(X, Wg, Hg, l1W=0.1, l2W=0.2)
        assert np.allclose(Wg, Wo, rtol=F64_RTOL)
        lo = ro.update_feature_maps(X, Wo, Ho, l1H=0.3)
        lg = rg.update_feature_maps(X, Wg, Hg, l1H=0.3)
        assert abs(lo - lg) < F64_RTOL * lo and np.allclose(Hg, Ho, rtol=F64_RTOL)
    assert rg.launch_count() > 0
    rg.close()


def test_host_stepped_fit_equals_device_loop(cmf, orc):
    X, W0, H0 = _config1(orc, N=60, T=300)
    a = cmf.fit_cnmf(X, L=10, K=5, max_itr=8, W_init=W0, H_init=H0, check_convergence=False, layout="KNL")
    rule = cmf.MultUpdate(X, W0, H0)
    b = cmf.fit(cmf.AlternatingOptimizer(rule, 8), X, 10, 5, W0, H0, check_convergence=False, layout="KNL")
    assert np.array_equal(a.loss_hist, b.loss_hist) and np.array_equal(a.W, b.W) and np.array_equal(a.H, b.H)


def test_driver_semantics(cmf, orc):
    X, W0, H0 = _config1(orc, N=40, T=200)
    msgs = []
    r = cmf.fit_cnmf(X, L=10, K=5, max_itr=400, W_init=W0, H_init=H0, tol=1e-3, printer=msgs.append)
    ref = orc.co.fit(orc.co.MultUpdate, X, W0, H0, 400, tol=1e-3, printer=lambda s: None)
    assert msgs == ["Converged early."] and len(r.loss_hist) == len(ref.loss_hist) < 401
    assert r.W.shape == (10, 40, 5)                                   # default layout L x N x K
    r0 = cmf.fit_cnmf(X, L=10, K=5, max_itr=0, W_init=W0, H_init=H0)
    assert len(r0.loss_hist) == 1 and abs(r0.loss_hist[0] - orc.po.compute_loss(X, W0, H0)) < 1e-12
    re = cmf.fit_cnmf(X, L=10, K=5, max_itr=3, W_init=W0, H_init=H0, eval_mode=True, check_convergence=False,
                      layout="KNL")
    assert np.allclose(re.W, W0, rtol=1e-15)                          # eval_mode skips update_motifs!
    rs_ = cmf.fit_cnmf(X, L=4, K=2, max_itr=2, seed=3)                # seeded random init path
    assert len(rs_.loss_hist) == 3 and rs_.loss_hist[0] < 1.0
    with pytest.raises(ValueError):
        cmf.fit_cnmf(X, L=4, K=2, alg="anls")


def test_init_rand_rescale(cmf, orc):
    X, _, _ = _config1(orc, N=30, T=150)
    W, H = cmf.init_rand(X, 6, 3, seed=5)
    est = orc.po.tensor_conv(W, H)
    assert abs(np.vdot(X, est) / np.vdot(est, est) - 1.0) < 1e-10


def _lockstep(shards, iters, reg):
    """Drives several DeviceShard handles on ONE GPU in lockstep, doing the collectives by hand
    (sum of the exchange tensors, halo copies) -- emulates the ranks without multi-process spins."""
    import math

    ss = sum(s.data_sumsq() for s in shards)
    for s in shards:
        s.set_data_norm(math.sqrt(ss))
    hist = []

    def loss():
        return math.sqrt(sum(s.loss_partial() for s in shards)) / math.sqrt(ss)

    hist.append(loss())
    for _ in range(iters):
        for s in shards:
            s.w_partials()
        for which in (0, 1):
            tot = sum(s.exchange[which].clone() for s in shards)
            for s in shards:
                s.exchange[which].copy_(tot)
        for s in shards:
            s.w_apply(reg.get("l1W", 0.0), reg.get("l2W", 0.0))
        for s in shards:
            s.h_update(reg.get("l1H", 0.0), reg.get("l2H", 0.0))
        for a, b in zip(shards[:-1], shards[1:]):
            b.recv_left.copy_(a.send_right)
            a.recv_right.copy_(b.send_left)
        hist.append(loss())
    return hist


@pytest.mark.parametrize("dtype,tol", [("f64", F64_RTOL), ("f32", 5e-5)])
def test_sharded_device_path_matches_oracle(cmf, orc, dtype, tol):
    import torch

    N, T, K, L, iters = 37, 301, 4, 9, 6
    W0, H0, X = _rand(N, T, K, L, seed=21)
    reg = dict(l1W=0.1, l2W=0.5, l1H=0.1, l2H=0.2)
    ref = orc.co.fit(orc.co.MultUpdate, X, W0, H0, iters, check_convergence=False, **reg)
    plan = cmf.ShardPlan(T, 3, L)
    shards = []
    for (t0, t1) in plan.ranges:
        s = cmf.DeviceShard(N, T, t0, t1, K, L, dtype=dtype, device=0)
        s.set_data(X, 0)
        s.set_factors(W0, H0, 0)
        shards.append(s)
    hist = _lockstep(shards, iters, reg)
    torch.cuda.synchronize()
    assert np.allclose(hist, ref.loss_hist, rtol=tol)
    H = np.concatenate([s.get_factors()[1] for s in shards], axis=1)
    for s in shards:
        assert np.allclose(s.get_factors()[0], ref.W, rtol=max(tol, 1e-9) * 50, atol=1e-12)
    assert np.allclose(H, ref.H, rtol=max(tol, 1e-9) * 50, atol=1e-12)
    for s in shards:
        s.close()


def test_synth_data_and_init_are_sharding_invariant(cmf):
    import ctypes

    N, T, K, L = 48, 500, 4, 8

    def losses(world):
        plan = cmf.ShardPlan(T, world, L)
        shards = []
        for (t0, t1) in plan.ranges:
            s = cmf.DeviceShard(N, T, t0, t1, K, L, dtype="f64", device=0)
            s.synth_data(1234, K, L, 0.05, 0.1)
            s.init_rand(0)
            shards.append(s)
        dot = sum(s.init_scale_partials()[0] for s in shards)
        nrm = sum(s.init_scale_partials()[1] for s in shards)
        for s in shards:
            s.scale_factors(np.sqrt(abs(dot / nrm)))
        h = _lockstep(shards, 3, {})
        for s in shards:
            s.close()
        return h

    a, b = losses(1), losses(2)
    assert a[0] < 1.0 and a[-1] < a[0]
    assert np.allclose(a, b, rtol=1e-10)


def test_adjointness_at_medium_size_fp32(cmf):
    # size-independent property: <conv(W,H), X> == <H, transconv(W,X)> == <W, corr(H,X)>
    N, T, K, L = 256, 20000, 16, 24
    W, H, X = _rand(N, T, K, L, seed=99)
    a = np.vdot(cmf.tensor_conv(W, H, "f32").astype(np.float64), X)
    b = np.vdot(H, cmf.tensor_transconv(W, X, "f32").astype(np.float64))
    c = np.vdot(W, cmf.corr_w(H, X, L, "f32").astype(np.float64))
    assert abs(a - b) < 1e-5 * abs(a) and abs(a - c) < 1e-5 * abs(a)


@pytest.mark.parametrize("reg", [{}, dict(l1W=0.05, l2W=0.3, l1H=0.02, l2H=0.1)])
def test_pgd_fp64_matches_oracle(cmf, orc, reg):
    # PGDUpdate (src/algs/pgd.jl) behind the same contractions
    X, W0, H0 = _config1(orc, N=120, T=600)
    ref = orc.po.fit(orc.po.PGDUpdate(X, W0, H0), X, W0, H0, 40, check_convergence=False, **reg)
    r = cmf.fit_cnmf(X, L=10, K=5, alg="pgd", max_itr=40, W_init=W0, H_init=H0, check_convergence=False,
                     layout="KNL", **reg)
    assert np.allclose(r.loss_hist, ref.loss_hist, rtol=F64_RTOL)
    assert np.allclose(r.W, ref.W, rtol=1e-8, atol=1e-12) and np.allclose(r.H, ref.H, rtol=1e-8, atol=1e-12)
    rule = cmf.PGDUpdate(X, W0.copy(), H0.copy())                     # rule-level interface, reference defaults
    ro = orc.po.PGDUpdate(X, W0, H0)
    Wg, Hg, Wo, Ho = W0.copy(), H0.copy(), W0.copy(), H0.copy()
    for _ in range(3):
        rule.update_motifs(X, Wg, Hg)
        ro.update_motifs(X, Wo, Ho)
        assert np.allclose(Wg, Wo, rtol=F64_RTOL)
        assert abs(rule.update_feature_maps(X, Wg, Hg) - ro.update_feature_maps(X, Wo, Ho)) < 1e-10
    rule.close()


def test_pgd_fp32_engines_loss_within_1e4(cmf, orc):
    N, T, K, L = 256, 4096, 8, 10
    X, _, _ = orc.po.synthetic_sequences(K=4, N=N, L=L, T=T, rng=np.random.default_rng(1234))
    W0, H0 = orc.po.init_rand(X, L, K, np.random.default_rng(0))
    ref = orc.po.fit(orc.po.PGDUpdate(X, W0, H0), X, W0, H0, 25, check_convergence=False)
    for engine in (0, 1):
        r = cmf.fit_cnmf(X, L=L, K=K, alg="pgd", max_itr=25, W_init=W0, H_init=H0, check_convergence=False,
                         dtype="f32", engine=engine, layout="KNL")
        rel = np.abs(np.asarray(r.loss_hist) - np.asarray(ref.loss_hist)) / np.asarray(ref.loss_hist)
        assert rel.max() < F32_LOSS_RTOL, (engine, rel.max())


@pytest.mark.parametrize("loss_func,masked", [("square", True), ("absolute", False), ("absolute", True)])
def test_pgd_masked_and_absolute_losses(cmf, orc, loss_func, masked):
    # the pluggable losses of PGD (src/algs/pgd.jl:28-70) on the device -- MaskedLoss (the configuration
    # of the reference's own test script, test/test.jl:28-44) and AbsoluteLoss -- against the oracle's PGDUpdate(loss_func=, mask=)
    X, W0, H0 = _config1(orc, N=60, T=400)
    mask = (np.random.default_rng(3).random(X.shape) < 0.8).astype(np.float64) if masked else None
    reg = dict(l1W=0.05, l2W=1.0, l1H=0.02, l2H=0.1)
    ref = orc.po.fit(orc.po.PGDUpdate(X, W0, H0, loss_func=loss_func, mask=mask), X, W0, H0, 30, check_convergence=False, **reg)
    r = cmf.fit_cnmf(X, L=10, K=5, alg="pgd", max_itr=30, W_init=W0, H_init=H0, check_convergence=False, layout="KNL",
                     loss_func=loss_func, mask=mask, **reg)
    assert np.allclose(r.loss_hist, ref.loss_hist, rtol=F64_RTOL), (r.loss_hist[-3:], ref.loss_hist[-3:])
    assert np.allclose(r.W, ref.W, rtol=1e-7, atol=1e-11) and np.allclose(r.H, ref.H, rtol=1e-7, atol=1e-11)
    if loss_func == "square":       # fp32 (the sign() of AbsoluteLoss flips on rounding-level residuals, so only the smooth loss is held to 1e-4)
        r32 = cmf.fit_cnmf(X, L=10, K=5, alg="pgd", max_itr=30, W_init=W0, H_init=H0, check_convergence=False, layout="KNL",
                           dtype="f32", loss_func=loss_func, mask=mask, **reg)
        rel = np.abs(np.asarray(r32.loss_hist) - np.asarray(ref.loss_hist)) / np.asarray(ref.loss_hist)
        assert rel.max() < F32_LOSS_RTOL, rel.max()
    # rule-level interface
    rule = cmf.PGDUpdate(X, W0.copy(), H0.copy(), loss_func=loss_func, mask=mask)
    ro = orc.po.PGDUpdate(X, W0, H0, loss_func=loss_func, mask=mask)
    Wg, Hg, Wo, Ho = W0.copy(), H0.copy(), W0.copy(), H0.copy()
    for _ in range(2):
        rule.update_motifs(X, Wg, Hg)
        ro.update_motifs(X, Wo, Ho)
        assert abs(rule.update_feature_maps(X, Wg, Hg) - ro.update_feature_maps(X, Wo, Ho)) < 1e-10
    rule.close()


@pytest.mark.parametrize("constrW,constrH", [("unitnorm", "nonneg"), ("nonneg", "unitnorm"), ("unitnorm", "unitnorm")])
def test_pgd_unit_norm_constraint(cmf, orc, constrW, constrH):
    # UnitNormConstraint (src/algs/pgd.jl:98-110) in place of NonnegConstraint on either factor, vs the oracle's restatement
    X, W0, H0 = _config1(orc, N=60, T=400)
    reg = dict(l1W=0.0, l2W=1.0, l1H=0.0, l2H=0.05)
    ref = orc.po.fit(orc.po.PGDUpdate(X, W0, H0, constrW=constrW, constrH=constrH), X, W0, H0, 25, check_convergence=False, **reg)
    r = cmf.fit_cnmf(X, L=10, K=5, alg="pgd", max_itr=25, W_init=W0, H_init=H0, check_convergence=False, layout="KNL",
                     constrW=constrW, constrH=constrH, **reg)
    assert np.allclose(r.loss_hist, ref.loss_hist, rtol=F64_RTOL), (r.loss_hist[-3:], ref.loss_hist[-3:])
    assert np.allclose(r.W, ref.W, rtol=1e-7, atol=1e-11) and np.allclose(r.H, ref.H, rtol=1e-7, atol=1e-11)
    if constrW == "unitnorm":
        assert np.all(np.linalg.norm(r.W.reshape(5, -1), axis=1) <= 1 + 1e-12)
    if constrH == "unitnorm":
        assert np.all(np.linalg.norm(r.H, axis=1) <= 1 + 1e-12)


def test_gen_synthetic_and_parameter_sweep(cmf):
    # README.md:14-23: data = CMF.gen_synthetic(N=500, T=2000); fit_cnmf(data; L=10, K=5, alg=:hals)
    data = cmf.gen_synthetic(N=60, T=300, K=3, L=8, seed=7)
    assert data.shape == (60, 300) and data.min() >= 0.0 and data.std() > 0
    assert np.array_equal(data, cmf.gen_synthetic(N=60, T=300, K=3, L=8, seed=7))
    assert not np.array_equal(data, cmf.gen_synthetic(N=60, T=300, K=3, L=8, seed=8))
    sweep = cmf.parameter_sweep(data, L_vals=[4, 6], K_vals=[2], alg_vals=["mult", ":hals"], max_itr=5, seed=0)
    assert set(sweep) == {(4, 2, "mult"), (6, 2, "mult"), (4, 2, ":hals"), (6, 2, ":hals")}
    for r in sweep.values():
        assert r.loss_hist[-1] < r.loss_hist[0]


def test_gen_synthetic_follows_the_reference_data_model(cmf):
    """datasets/synthetic.jl:29-61: Dirichlet(0.1) unit weights x Gaussian lag bumps (sigma 0.2 on linspace(-1,1,L), centres
    U(-1,1)), Exp(1)*Bernoulli(p_h) activations, N(0, noise^2) noise, clipped at 0.  The device generator uses its own
    counter-based RNG, so the comparison with the oracle's restatement of that file is distributional: pooled over seeds,
    the first two moments, the clipped fraction and upper quantiles of the data must agree."""
    from oracle import cnmf_oracle as po

    N, T, K, L, p_h, noise = 160, 3000, 3, 20, 0.3, 0.5
    dev = np.concatenate([cmf.gen_synthetic(N=N, T=T, K=K, L=L, p_h=p_h, noise_scale=noise, seed=s).ravel() for s in (1, 2, 3, 4)])
    ref = np.concatenate([po.synthetic_sequences(K=K, N=N, L=L, T=T, p_h=p_h, noise_scale=noise,
                                                 rng=np.random.default_rng(100 + s))[0].ravel() for s in (1, 2, 3, 4)])
    assert dev.min() >= 0.0
    assert abs(dev.mean() - ref.mean()) < 0.06 * ref.mean(), (dev.mean(), ref.mean())
    assert abs(dev.std() - ref.std()) < 0.08 * ref.std(), (dev.std(), ref.std())
    assert abs((dev == 0).mean() - (ref == 0).mean()) < 0.02, ((dev == 0).mean(), (ref == 0).mean())
    for q in (0.5, 0.9, 0.99):
        a, b = np.quantile(dev, q), np.quantile(ref, q)
        assert abs(a - b) < 0.08 * b + 0.02, (q, a, b)
    # noise-free data: every unit's weights sum to one over the components, so sum_t X[n, t] / sum_k,t H-mass is the bump mass
    clean = cmf.gen_synthetic(N=N, T=T, K=K, L=L, p_h=p_h, noise_scale=0.0, seed=9)
    clean_ref = po.synthetic_sequences(K=K, N=N, L=L, T=T, p_h=p_h, noise_scale=0.0, rng=np.random.default_rng(9))[0]
    assert abs(clean.mean() - clean_ref.mean()) < 0.08 * clean_ref.mean(), (clean.mean(), clean_ref.mean())
    assert abs((clean == 0).mean() - (clean_ref == 0).mean()) < 0.03


def test_cuda_fp64_matches_exact_rational_pin(cmf):
    # The pin that is not floating-point code checking floating-point code: one MU iteration and one HALS iteration on the
    # reference's toy data (datasets/toy.jl:5-48) computed in exact rational arithmetic (oracle/exact_pin.py,
    # tests/golden/make_exact_pin.py) vs the fp64 CUDA path through the plugin boundary.
    g = np.load(os.path.join(GOLD, "exact_pin_toy.npz"))
    X, W0, H0 = g["X"], g["W0"], g["H0"]
    l1W, l2W, l1H, l2H = (float(v) for v in g["reg"])
    for rule_cls, key in ((cmf.MultUpdate, "mu"), (cmf.HALSUpdate, "hals")):
        W, H = W0.copy(), H0.copy()
        rule = rule_cls(X, W, H, dtype="f64")
        rule.update_motifs(X, W, H, l1W=l1W, l2W=l2W)
        loss = rule.update_feature_maps(X, W, H, l1H=l1H, l2H=l2H)
        rule.close()
        We, He, le = g[key + "_W"], g[key + "_H"], float(g[key + "_loss"])
        assert abs(loss - le) < 1e-11 * le, (key, loss, le)
        assert np.max(np.abs(W - We)) < 1e-11 * np.max(We) and np.max(np.abs(H - He)) < 1e-11 * np.max(He), key


def test_cuda_fp32_close_to_exact_rational_pin(cmf):
    g = np.load(os.path.join(GOLD, "exact_pin_toy.npz"))
    X, W0, H0 = g["X"], g["W0"], g["H0"]
    l1W, l2W, l1H, l2H = (float(v) for v in g["reg"])
    for rule_cls, key in ((cmf.MultUpdate, "mu"), (cmf.HALSUpdate, "hals")):
        W, H = W0.copy(), H0.copy()
        rule = rule_cls(X, W, H, dtype="f32")
        rule.update_motifs(X, W, H, l1W=l1W, l2W=l2W)
        loss = rule.update_feature_maps(X, W, H, l1H=l1H, l2H=l2H)
        rule.close()
        assert abs(loss - float(g[key + "_loss"])) < 1e-5 * float(g[key + "_loss"]), key


# ---- config-5 shape of the components (K = 128, L = 32): HALS against the plain-C oracle ---------------------------------------
@pytest.fixture(scope="module")
def c5shape(orc):
    N, T, K, L = 32, 3000, 128, 32
    X, _, _ = orc.po.synthetic_sequences(K=6, N=N, L=L, T=T, p_h=0.3, noise_scale=0.5, rng=np.random.default_rng(5))
    W0, H0 = orc.po.init_rand(X, L, K, np.random.default_rng(6))
    reg = dict(l1W=0.05, l2W=0.1, l1H=0.05, l2H=0.1)
    ref = orc.co.fit(orc.co.HALSUpdate, X, W0, H0, 10, check_convergence=False, **reg)
    return X, W0, H0, reg, ref


def test_hals_c5_component_shape_fp64(cmf, c5shape):
    X, W0, H0, reg, ref = c5shape
    r = cmf.fit_cnmf(X, L=32, K=128, alg="hals", max_itr=10, W_init=W0, H_init=H0, check_convergence=False, layout="KNL", **reg)
    assert np.allclose(r.loss_hist, ref.loss_hist, rtol=1e-9, atol=0)
    assert np.allclose(r.W, ref.W, rtol=1e-7, atol=1e-10) and np.allclose(r.H, ref.H, rtol=1e-7, atol=1e-10)


@pytest.mark.parametrize("sweep", ["2", "1"])
def test_hals_c5_component_shape_fp32(cmf, c5shape, monkeypatch, sweep):
    # sweep 2 = round-based kernel (kernels_hals.cuh: lane = component recurrences, 8 x 8 pull blocks; 140 CTAs at K = 128),
    # sweep 1 = first-generation wavefront kernel; both against the oracle at the north-star's fp32 bar
    X, W0, H0, reg, ref = c5shape
    monkeypatch.setenv("CMF_HALS_SWEEP", sweep)
    r = cmf.fit_cnmf(X, L=32, K=128, alg="hals", max_itr=10, W_init=W0, H_init=H0, check_convergence=False, layout="KNL",
                     dtype="f32", **reg)
    rel = np.abs(np.asarray(r.loss_hist) - np.asarray(ref.loss_hist)) / np.asarray(ref.loss_hist)
    assert rel.max() < 1e-4, rel


@pytest.mark.parametrize("dims", [(24, 5000, 9, 6), (16, 2600, 40, 32), (16, 1500, 128, 17), (8, 900, 3, 1), (12, 1024, 20, 8)])
def test_hals_round_kernel_equals_wavefront_kernel(cmf, dims, monkeypatch):
    # one HALS iteration from the same factors with the two H-sweep kernels: same sequential order, different order of the
    # floating-point sums in the pull (fixed in both) -> agreement to fp32 rounding, identical zero pattern up to rounding
    N, T, K, L = dims
    W, H, X = _rand(N, T, K, L, seed=sum(dims))
    out = {}
    for sweep in ("1", "2"):
        monkeypatch.setenv("CMF_HALS_SWEEP", sweep)
        Wc, Hc = W.copy(), H.copy()
        rule = cmf.HALSUpdate(X, Wc, Hc, dtype="f32", engine=0)
        rule.update_motifs(X, Wc, Hc, l2W=0.1)
        loss = rule.update_feature_maps(X, Wc, Hc, l1H=0.05, l2H=0.1)
        rule.close()
        out[sweep] = (Hc, loss)
    assert abs(out["1"][1] - out["2"][1]) < 1e-5 * out["1"][1]
    assert np.max(np.abs(out["1"][0] - out["2"][0])) < 2e-4 * np.max(np.abs(out["1"][0]))

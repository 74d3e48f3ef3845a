"""GPU parity tests of the tcgen05 tensor-core engine (fp32 data, split-bf16 operands) against the
Float64 CPU oracle: each of the three contraction kernels in isolation, then whole MU fits.

Tolerance: the engine carries ~16-17 mantissa bits per operand, so single contractions are
checked to 3e-5 of the output scale and the loss trajectory to the north-star's 1e-4."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cmf():
    import __graft_entry__ as ge

    ge.build()
    import cmf_jl_b200

    return cmf_jl_b200


@pytest.fixture(scope="module")
def orc():
    from oracle import c_oracle, cnmf_oracle

    class O:
        po, co = cnmf_oracle, c_oracle

    return O


def _rand(N, T, K, L, seed=0):
    rng = np.random.default_rng(seed)
    return rng.random((K, N, L)), rng.random((K, T)), rng.random((N, T))


def _scale_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


# (N, T, K, L): Kp/G variants (16/8, 32/4, 64/2, 128/1), ragged N and T, multiple splits / flushes
TC_DIMS = [(256, 5000, 16, 12), (264, 3001, 20, 7), (384, 9000, 64, 20), (128, 20000, 128, 3),
           (512, 40000, 8, 33)]


@pytest.mark.parametrize("dims", TC_DIMS)
def test_tc_contractions_match_oracle(cmf, orc, dims):
    import torch

    N, T, K, L = dims
    W, H, X = _rand(*dims, seed=sum(dims))
    s = cmf.DeviceShard(N, T, 0, T, K, L, dtype="f32", device=0)
    s.set_engine(1)
    assert s.get_engine() == 1
    s.set_data(X, 0)
    s.set_factors(W, H, 0)
    s.set_data_norm(1.0)
    # TC_CONV: sum of squared residuals
    ref_ss = float(np.sum((orc.co.tensor_conv(W, H) - X) ** 2))
    got_ss = s.loss_partial()
    assert abs(got_ss - ref_ss) < 2e-5 * ref_ss, (got_ss, ref_ss)
    # TC_CORR: numW in the internal [(l,k)][n] layout
    s.w_partials()
    torch.cuda.synchronize()
    numW = s.exchange[0].cpu().numpy().reshape(L, K, N).transpose(1, 2, 0)
    assert _scale_err(numW, orc.co.corr_w(H, X, L)) < 3e-5
    # Gram partial Rg[d][k][k'] (tensor-core correlation of H with itself)
    from oracle import restructured as rs

    Rg = s.exchange[1].cpu().numpy()[: L * K * K].reshape(L, K, K).transpose(1, 2, 0)
    assert _scale_err(Rg, rs.gram_R(H, L)) < 3e-5
    # TC_TRANS: numH [t][K]
    s.h_update(0.0, 0.0)
    torch.cuda.synchronize()
    numH = s.exchange[2].cpu().numpy().reshape(T, K).T
    assert _scale_err(numH, orc.co.tensor_transconv(W, X)) < 3e-5
    # denomH = C (*) H on tensor cores + truncated tail
    denH = s.exchange[3].cpu().numpy().reshape(T, K).T
    assert _scale_err(denH, orc.co.tensor_transconv(W, orc.co.tensor_conv(W, H))) < 3e-5
    s.close()


def test_tc_fit_matches_oracle_loss(cmf, orc):
    N, T, K, L = 256, 4096, 8, 10
    X, _, _ = orc.po.synthetic_sequences(K=4, N=N, L=L, T=T, rng=np.random.default_rng(1234))
    W0, H0 = orc.po.init_rand(X, L, K, np.random.default_rng(0))
    reg = dict(l1W=0.1, l2W=0.5, l1H=0.1, l2H=0.2)
    ref = orc.co.fit(orc.co.MultUpdate, X, W0, H0, 40, check_convergence=False, **reg)
    r = cmf.fit_cnmf(X, L=L, K=K, alg="mult", max_itr=40, W_init=W0, H_init=H0, check_convergence=False,
                     dtype="f32", engine=1, layout="KNL", **reg)
    rel = np.abs(np.asarray(r.loss_hist) - np.asarray(ref.loss_hist)) / np.asarray(ref.loss_hist)
    assert rel.max() < 1e-4, rel.max()
    assert np.linalg.norm(r.W - ref.W) / np.linalg.norm(ref.W) < 1e-3
    assert np.linalg.norm(r.H - ref.H) / np.linalg.norm(ref.H) < 1e-3


@pytest.mark.parametrize("noise,p_h", [(1.0, 0.5), (0.05, 0.1)])
def test_tc_expansion_loss_matches_oracle(cmf, orc, noise, p_h):
    # loss evaluated by the algebraic expansion on the resident numH / C table (cmf_set_loss_mode 1)
    N, T, K, L = 256, 4096, 8, 10
    X, _, _ = orc.po.synthetic_sequences(K=4, N=N, L=L, T=T, noise_scale=noise, p_h=p_h,
                                         rng=np.random.default_rng(1234))
    W0, H0 = orc.po.init_rand(X, L, K, np.random.default_rng(0))
    ref = orc.co.fit(orc.co.MultUpdate, X, W0, H0, 40, check_convergence=False)
    r = cmf.fit_cnmf(X, L=L, K=K, alg="mult", max_itr=40, W_init=W0, H_init=H0, check_convergence=False,
                     dtype="f32", engine=1, loss_mode=1, layout="KNL")
    rel = np.abs(np.asarray(r.loss_hist) - np.asarray(ref.loss_hist)) / np.asarray(ref.loss_hist)
    print("expansion loss: max rel err", rel.max(), "final loss", ref.loss_hist[-1])
    assert rel.max() < 1e-4, (rel.max(), ref.loss_hist[-1])


def test_tc_sharded_matches_single(cmf, orc):
    import math

    import torch

    N, T, K, L, iters = 256, 12000, 16, 9, 4
    W0, H0, X = _rand(N, T, K, L, seed=5)

    def run(world, loss_mode=0):
        plan = cmf.ShardPlan(T, world, L)
        shards = []
        for (t0, t1) in plan.ranges:
            s = cmf.DeviceShard(N, T, t0, t1, K, L, dtype="f32", device=0)
            s.set_engine(1)
            s.set_loss_mode(loss_mode)
            s.set_data(X, 0)
            s.set_factors(W0, H0, 0)
            shards.append(s)
        ss = sum(s.data_sumsq() for s in shards)
        for s in shards:
            s.set_data_norm(math.sqrt(ss))
        hist = []
        for _ in range(iters):
            for s in shards:
                s.w_partials()
            for which in (0, 1):
                tot = sum(s.exchange[which].clone() for s in shards)
                for s in shards:
                    s.exchange[which].copy_(tot)
            for s in shards:
                s.w_apply(0.0, 0.0)
                s.h_update(0.0, 0.0)
            for a, b in zip(shards[:-1], shards[1:]):
                b.recv_left.copy_(a.send_right)
                a.recv_right.copy_(b.send_left)
            hist.append(math.sqrt(sum(s.loss_partial() for s in shards) / ss))
        torch.cuda.synchronize()
        for s in shards:
            s.close()
        return np.asarray(hist)

    a, b, c = run(1), run(3), run(3, loss_mode=1)
    ref = orc.co.fit(orc.co.MultUpdate, X, W0, H0, iters, check_convergence=False)
    assert np.allclose(a, ref.loss_hist[1:], rtol=1e-4)
    assert np.allclose(a, b, rtol=2e-6)
    assert np.allclose(c, ref.loss_hist[1:], rtol=1e-4)      # expansion loss summed over shards


def test_tc_hals_matches_oracle(cmf, orc):
    # HALS on the tensor-core engine: gradients P = denomW - numW, Q = denomH - numH from the tcgen05 contractions,
    # per-unit W sweep and the cooperative wavefront H sweep
    N, T, K, L = 256, 4096, 8, 10
    X, _, _ = orc.po.synthetic_sequences(K=4, N=N, L=L, T=T, rng=np.random.default_rng(1234))
    W0, H0 = orc.po.init_rand(X, L, K, np.random.default_rng(0))
    reg = dict(l1W=0.1, l2W=0.5, l1H=0.1, l2H=0.2)
    ref = orc.co.fit(orc.co.HALSUpdate, X, W0, H0, 12, check_convergence=False, **reg)
    r = cmf.fit_cnmf(X, L=L, K=K, alg="hals", max_itr=12, W_init=W0, H_init=H0, check_convergence=False,
                     dtype="f32", engine=1, layout="KNL", **reg)
    rel = np.abs(np.asarray(r.loss_hist) - np.asarray(ref.loss_hist)) / np.asarray(ref.loss_hist)
    assert rel.max() < 1e-4, rel


def test_hals_wavefront_many_components_fp64(cmf, orc):
    # more components than fit one wave of the pipeline order, several time chunks, truncated tail
    N, T, K, L = 24, 700, 9, 6
    W, H, X = _rand(N, T, K, L, seed=3)
    ref = orc.co.fit(orc.co.HALSUpdate, X, W, H, 4, check_convergence=False, l1H=0.05, l2W=0.1)
    r = cmf.fit_cnmf(X, L=L, K=K, alg="hals", max_itr=4, W_init=W, H_init=H, check_convergence=False,
                     layout="KNL", l1H=0.05, l2W=0.1)
    assert np.allclose(r.loss_hist, ref.loss_hist, rtol=1e-9)
    assert np.allclose(r.H, ref.H, rtol=1e-8, atol=1e-11) and np.allclose(r.W, ref.W, rtol=1e-8, atol=1e-11)


@pytest.mark.parametrize("dims", [(64, 5000, 9, 6), (128, 9000, 40, 32), (32, 3000, 5, 40)])
def test_hals_fp32_sweep_shapes(cmf, orc, dims):
    # fp32 H sweep over several chunks: few / many components, L = 32 (register window) and L = 40 (shared ring)
    N, T, K, L = dims
    W, H, X = _rand(N, T, K, L, seed=sum(dims))
    r = cmf.fit_cnmf(X, L=L, K=K, alg="hals", max_itr=3, W_init=W, H_init=H, check_convergence=False,
                     dtype="f32", engine=0, layout="KNL", l1H=0.05, l2W=0.1)
    ref = orc.co.fit(orc.co.HALSUpdate, X, W, H, 3, check_convergence=False, l1H=0.05, l2W=0.1)
    assert np.allclose(r.loss_hist, ref.loss_hist, rtol=1e-4)

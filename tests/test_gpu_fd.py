"""GPU parity tests of the frequency-domain engine (cmf_set_engine 2: overlap-save blocks, SIMT FFT kernels, the two
per-frequency complex products on tcgen05 with split-bf16 operands) against the Float64 CPU oracle: numW and numH in
isolation on ragged shapes, then whole fits (MU both loss modes, sharded MU, HALS, PGD).

Tolerance: like the time-domain tensor-core engine, single contractions to 3e-5 of the output scale, the loss
trajectory to the north-star's 1e-4."""
import math

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


# (N, T, K, L): block lengths 64..512, odd K, ragged N (not a multiple of 32/64/256) and T, L = 1, small T,
# one block only, many column tiles of blocks (T / V > 256)
FD_DIMS = [(256, 5000, 16, 12), (264, 3001, 21, 7), (384, 9000, 64, 20), (512, 40000, 8, 33), (128, 3000, 5, 1),
           (256, 6000, 10, 100), (136, 400, 3, 30), (64, 20000, 4, 5),
           # more than 64 components: two 128-row tiles per frequency (Kq = 128)
           (128, 20000, 128, 3), (256, 6000, 100, 12), (136, 3000, 65, 9)]


@pytest.mark.parametrize("dims", FD_DIMS)
def test_fd_contractions_match_oracle(cmf, orc, dims):
    import torch

    N, T, K, L = dims
    W, H, X = _rand(*dims, seed=sum(dims))
    s = cmf.DeviceShard(N, T, 0, T, K, L, dtype="f32", device=0)
    s.set_engine(2)
    assert s.get_engine() == 2
    s.set_data(X, 0)
    s.set_factors(W, H, 0)
    s.set_data_norm(1.0)
    # direct loss pass through the frequency domain (TC_FQX + inverse FFT + residual)
    s.set_loss_mode(0)
    ref_ss = float(np.sum((orc.co.tensor_conv(W, H) - X) ** 2))
    got_ss = s.loss_partial()
    assert abs(got_ss - ref_ss) < 2e-5 * ref_ss, (got_ss, ref_ss)
    s.w_partials()
    torch.cuda.synchronize()
    numW = s.exchange[0].cpu().numpy().reshape(L, K, N).transpose(1, 2, 0)
    assert _scale_err(numW, orc.co.corr_w(H, X, L)) < 3e-5
    # Gram partial Rg[d][k][k'] through the spectrum of H
    from oracle import restructured as rs

    Rg = s.exchange[1].cpu().numpy()[: L * K * K].reshape(L, K, K).transpose(1, 2, 0)
    assert _scale_err(Rg, rs.gram_R(H, L)) < 3e-5
    s.h_update(0.0, 0.0)
    torch.cuda.synchronize()
    numH = s.exchange[2].cpu().numpy().reshape(T, K).T
    assert _scale_err(numH, orc.co.tensor_transconv(W, X)) < 3e-5
    # denomH = C (*) H through the spectrum of H + truncated tail
    denH = s.exchange[3].cpu().numpy().reshape(T, K).T
    assert _scale_err(denH, orc.co.tensor_transconv(W, orc.co.tensor_conv(W, H))) < 3e-5
    # new data on the same handle: the spectrum of X is rebuilt
    X2 = np.random.default_rng(7).random((N, T))
    s.set_data(X2, 0)
    s.set_factors(W, H, 0)
    s.w_partials()
    torch.cuda.synchronize()
    numW = s.exchange[0].cpu().numpy().reshape(L, K, N).transpose(1, 2, 0)
    assert _scale_err(numW, orc.co.corr_w(H, X2, L)) < 3e-5
    s.close()


@pytest.mark.parametrize("loss_mode", [0, 1])
def test_fd_fit_matches_oracle_loss(cmf, orc, loss_mode):
    N, T, K, L = 256, 4096, 8, 10
    X, _, _ = orc.po.synthetic_sequences(K=4, N=N, L=L, T=T, rng=np.random.default_rng(1234))
    W0, H0 = orc.po.init_rand(X, L, K, np.random.default_rng(0))
    reg = dict(l1W=0.1, l2W=0.5, l1H=0.1, l2H=0.2) if loss_mode == 0 else {}
    ref = orc.co.fit(orc.co.MultUpdate, X, W0, H0, 100, check_convergence=False, **reg)
    r = cmf.fit_cnmf(X, L=L, K=K, alg="mult", max_itr=100, W_init=W0, H_init=H0, check_convergence=False,
                     dtype="f32", engine=2, loss_mode=loss_mode, layout="KNL", **reg)
    rel = np.abs(np.asarray(r.loss_hist) - np.asarray(ref.loss_hist)) / np.asarray(ref.loss_hist)
    print("fd fit: max rel loss err", rel.max(), "last", rel[-1])
    assert rel.max() < 1e-4, rel.max()
    assert np.linalg.norm(r.W - ref.W) / np.linalg.norm(ref.W) < 1e-3
    assert np.linalg.norm(r.H - ref.H) / np.linalg.norm(ref.H) < 1e-3


def test_fd_sparse_low_noise_fit(cmf, orc):
    # the hard case of SURVEY Appendix E: sparse activations, low noise, small final loss
    N, T, K, L = 256, 4096, 8, 10
    X, _, _ = orc.po.synthetic_sequences(K=4, N=N, L=L, T=T, noise_scale=0.05, p_h=0.1, rng=np.random.default_rng(1234))
    W0, H0 = orc.po.init_rand(X, L, K, np.random.default_rng(0))
    ref = orc.co.fit(orc.co.MultUpdate, X, W0, H0, 100, check_convergence=False)
    r = cmf.fit_cnmf(X, L=L, K=K, alg="mult", max_itr=100, W_init=W0, H_init=H0, check_convergence=False,
                     dtype="f32", engine=2, layout="KNL")
    rel = np.abs(np.asarray(r.loss_hist) - np.asarray(ref.loss_hist)) / np.asarray(ref.loss_hist)
    print("fd sparse fit: max rel loss err", rel.max(), "last", rel[-1], "final loss", ref.loss_hist[-1])
    assert rel[-1] < 1e-4 and rel.max() < 1e-4, (rel.max(), rel[-1])


def test_fd_sharded_matches_single(cmf, orc):
    import torch

    N, T, K, L, iters = 256, 12000, 16, 9, 4
    W0, H0, X = _rand(N, T, K, L, seed=5)

    def run(world):
        plan = cmf.ShardPlan(T, world, L)
        shards = []
        for (t0, t1) in plan.ranges:
            s = cmf.DeviceShard(N, T, t0, t1, K, L, dtype="f32", device=0)
            s.set_engine(2)
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

    a, b = run(1), run(3)
    ref = orc.co.fit(orc.co.MultUpdate, X, W0, H0, iters, check_convergence=False)
    assert np.allclose(a, ref.loss_hist[1:], rtol=1e-4)
    assert np.allclose(a, b, rtol=2e-5)


@pytest.mark.parametrize("alg,iters", [("hals", 12), ("pgd", 30)])
def test_fd_other_rules_match_oracle(cmf, orc, alg, iters):
    # HALS and PGD take numW / numH from the same two contractions
    N, T, K, L = 256, 4096, 8, 10
    X, _, _ = orc.po.synthetic_sequences(K=4, N=N, L=L, T=T, rng=np.random.default_rng(1234))
    W0, H0 = orc.po.init_rand(X, L, K, np.random.default_rng(0))
    if alg == "hals":
        reg = dict(l1W=0.1, l2W=0.5, l1H=0.1, l2H=0.2)
        ref = orc.co.fit(orc.co.HALSUpdate, X, W0, H0, iters, check_convergence=False, **reg)
    else:
        reg = {}
        ref = orc.po.fit(orc.po.PGDUpdate(X, W0, H0), X, W0, H0, iters, check_convergence=False)
    for loss_mode in (0, 1):      # 1: every loss (incl. the initial one and those after W-only steps) by the expansion
        r = cmf.fit_cnmf(X, L=L, K=K, alg=alg, max_itr=iters, W_init=W0, H_init=H0, check_convergence=False,
                         dtype="f32", engine=2, loss_mode=loss_mode, layout="KNL", **reg)
        rel = np.abs(np.asarray(r.loss_hist) - np.asarray(ref.loss_hist)) / np.asarray(ref.loss_hist)
        print(alg, "loss_mode", loss_mode, "max rel loss err", rel.max())
        assert rel.max() < 1e-4, (loss_mode, rel)


def test_fd_fit_many_components(cmf, orc):
    # K > 64: every product runs on two 128-row tiles per frequency; MU with both loss modes and HALS
    N, T, K, L = 128, 3000, 80, 6
    X, _, _ = orc.po.synthetic_sequences(K=6, N=N, L=L, T=T, rng=np.random.default_rng(1234))
    W0, H0 = orc.po.init_rand(X, L, K, np.random.default_rng(0))
    ref = orc.co.fit(orc.co.MultUpdate, X, W0, H0, 30, check_convergence=False)
    for loss_mode in (0, 1):
        r = cmf.fit_cnmf(X, L=L, K=K, alg="mult", max_itr=30, W_init=W0, H_init=H0, check_convergence=False,
                         dtype="f32", engine=2, loss_mode=loss_mode, layout="KNL")
        rel = np.abs(np.asarray(r.loss_hist) - np.asarray(ref.loss_hist)) / np.asarray(ref.loss_hist)
        print("K=80 mult loss_mode", loss_mode, "max rel loss err", rel.max())
        assert rel.max() < 1e-4, (loss_mode, rel.max())
    refh = orc.co.fit(orc.co.HALSUpdate, X, W0, H0, 6, check_convergence=False, l1H=0.05, l2W=0.1)
    rh = cmf.fit_cnmf(X, L=L, K=K, alg="hals", max_itr=6, W_init=W0, H_init=H0, check_convergence=False,
                      dtype="f32", engine=2, layout="KNL", l1H=0.05, l2W=0.1)
    rel = np.abs(np.asarray(rh.loss_hist) - np.asarray(refh.loss_hist)) / np.asarray(refh.loss_hist)
    print("K=80 hals max rel loss err", rel.max())
    assert rel.max() < 1e-4, rel


def test_fd_loss_pass_in_chunks(cmf, orc, monkeypatch):
    # the direct loss pass works through the blocks in chunks (Yf holds the spectrum of Xhat for one chunk): force 2 chunks
    monkeypatch.setenv("CMF_FD_NBC", "256")
    N, T, K, L = 64, 20000, 4, 5                       # block length 64, hop 60 -> 334 blocks
    W, H, X = _rand(N, T, K, L, seed=11)
    s = cmf.DeviceShard(N, T, 0, T, K, L, dtype="f32", device=0)
    s.set_engine(2)
    s.set_loss_mode(0)
    s.set_data(X, 0)
    s.set_factors(W, H, 0)
    ref_ss = float(np.sum((orc.co.tensor_conv(W, H) - X) ** 2))
    got_ss = s.loss_partial()
    assert abs(got_ss - ref_ss) < 2e-5 * ref_ss, (got_ss, ref_ss)
    s.close()


def test_fd_unsupported_shapes_fail_loudly(cmf):
    s = cmf.DeviceShard(128, 20000, 0, 20000, 130, 3, dtype="f32", device=0)   # K > 128
    with pytest.raises(Exception):
        s.set_engine(2)
    s.close()

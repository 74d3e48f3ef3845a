"""Multi-GPU through the reference-facing boundary: the same fit_cnmf / update-rule calls with the time axis sharded
INSIDE libcmf_sm100 (NCCL bound by the library).  Needs >= 2 visible GPUs (`gpurun --gpus 2`); skipped otherwise.

  * cmf_create_multi  (one process, one calling thread)      -- fit_cnmf(..., ngpu=2), MultUpdate(..., ngpu=2)
  * cmf_create_rank   (one process per GPU, torchrun style)  -- tests/mp_rank_fit.py launched with torch.distributed.run

The sharded run reorders floating-point sums only (per-shard partials, then the all-reduce), so fp64 results stay within
1e-9 of the single-process oracle and fp32 losses within 1e-4 (the bars of BASELINE.json north_star)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")


@pytest.fixture(scope="module")
def cmf():
    import __graft_entry__ as ge

    ge.build()
    import cmf_jl_b200

    return cmf_jl_b200


@pytest.fixture(scope="module")
def co():
    from oracle import c_oracle

    return c_oracle


def _problem(N, T, K, L, seed):
    from oracle import cnmf_oracle as po

    X, _, _ = po.synthetic_sequences(K=max(2, K // 2), N=N, L=L, T=T, rng=np.random.default_rng(seed))
    W0, H0 = po.init_rand(X, L, K, np.random.default_rng(seed + 1))
    return X, W0, H0


REG = dict(l1W=0.1, l2W=0.5, l1H=0.1, l2H=0.2)


@needs2
@pytest.mark.parametrize("alg", ["mult", "hals"])
@pytest.mark.parametrize("reg", [{}, REG])
def test_group_handle_fp64_matches_oracle(cmf, co, alg, reg):
    X, W0, H0 = _problem(40, 700, 4, 7, 11)
    rule = co.MultUpdate if alg == "mult" else co.HALSUpdate
    ref = co.fit(rule, X, W0, H0, 12, check_convergence=False, **reg)
    for ngpu in (2, min(_ngpu(), 4)):
        r = cmf.fit_cnmf(X, L=7, K=4, alg=alg, max_itr=12, W_init=W0, H_init=H0, check_convergence=False, layout="KNL",
                         ngpu=ngpu, **reg)
        assert np.allclose(r.loss_hist, ref.loss_hist, rtol=1e-9, atol=0), (ngpu, r.loss_hist, ref.loss_hist)
        assert np.allclose(r.W, ref.W, rtol=1e-8, atol=1e-11) and np.allclose(r.H, ref.H, rtol=1e-8, atol=1e-11)


@needs2
@pytest.mark.parametrize("alg,engine,loss_mode", [("mult", 0, 0), ("mult", 1, 0), ("mult", 2, 0), ("mult", 2, 1), ("hals", 2, 0)])
def test_group_handle_fp32_engines(cmf, co, alg, engine, loss_mode):
    X, W0, H0 = _problem(256, 9000, 8, 10, 5)
    rule = co.MultUpdate if alg == "mult" else co.HALSUpdate
    ref = co.fit(rule, X, W0, H0, 10, check_convergence=False, **REG)
    r = cmf.fit_cnmf(X, L=10, K=8, alg=alg, max_itr=10, W_init=W0, H_init=H0, check_convergence=False, layout="KNL",
                     dtype="f32", engine=engine, loss_mode=loss_mode, ngpu=2, **REG)
    rel = np.abs(np.asarray(r.loss_hist) - np.asarray(ref.loss_hist)) / np.asarray(ref.loss_hist)
    assert rel.max() < 1e-4, rel


@needs2
def test_group_handle_rule_interface_and_random_init(cmf, co):
    # the plugin boundary itself (Rule(data,W,H); update_motifs!; update_feature_maps!) on a 2-GPU group handle, and the
    # seeded init + alpha rescale path of fit_cnmf (no W_init/H_init)
    X, W0, H0 = _problem(32, 500, 3, 6, 3)
    Wr, Hr = np.asfortranarray(W0.copy()), np.asfortranarray(H0.copy())
    ref = co.MultUpdate(X, Wr, Hr)
    rule = cmf.MultUpdate(X, W0, H0, ngpu=2)
    W, H = W0.copy(), H0.copy()
    for _ in range(3):
        ref.update_motifs(X, Wr, Hr)
        lr = ref.update_feature_maps(X, Wr, Hr)
        rule.update_motifs(X, W, H)
        lg = rule.update_feature_maps(X, W, H)
        assert abs(lg - lr) < 1e-9 * lr
        assert np.allclose(W, Wr, rtol=1e-8, atol=1e-12) and np.allclose(H, Hr, rtol=1e-8, atol=1e-12)
    rule.close()
    a = cmf.fit_cnmf(X, L=6, K=3, max_itr=5, seed=7, check_convergence=False, ngpu=2)
    b = cmf.fit_cnmf(X, L=6, K=3, max_itr=5, seed=7, check_convergence=False, ngpu=1)
    assert np.allclose(a.loss_hist, b.loss_hist, rtol=1e-9)


@needs2
@pytest.mark.parametrize("alg", ["mult", "hals"])
def test_rank_per_process_matches_oracle(alg):
    # one process per GPU (torchrun style): every rank calls the same library entry points, NCCL inside the library
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "mp_rank_fit.py"), alg]
    p = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "RANK_FIT_OK" in p.stdout

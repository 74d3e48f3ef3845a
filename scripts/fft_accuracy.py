"""CPU emulation of the frequency-domain contraction path: does it hold the fp32 parity bar?

MU on the fp64 oracle's data and inits, with numW (mult.jl:32) and numH (mult.jl:47) computed the way the device does:
overlap-save blocks of length B (hop V = B-L+1), fp32 FFT, spectra split into bf16 hi/lo planes, three products
(hi*hi + hi*lo + lo*hi) accumulated in fp32, fp32 inverse FFT.  Everything else (Gram-form denominators, update, loss)
stays fp64 so the difference to the oracle is the error of the frequency-domain path alone.  TEST TOOLING ONLY.
"""
import os
import sys

import numpy as np
import scipy.fft as sfft

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import c_oracle as co  # noqa: E402
from oracle import cnmf_oracle as po  # noqa: E402
from oracle import restructured as rs  # noqa: E402

EPS = po.EPSILON


def bf16(x):
    """Round-to-nearest-even fp32 -> bf16, returned as fp32."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + 0x7FFF
    return ((u + r) & 0xFFFF0000).view(np.float32)


def split(x):
    hi = bf16(x)
    lo = bf16(x.astype(np.float32) - hi)
    return hi, lo


def mm3(A, B, products=3):
    """A @ B over the last/first axes with split-bf16 operands, fp32 accumulation (batched over axis 0)."""
    Ah, Al = split(A)
    Bh, Bl = split(B)
    out = np.matmul(Ah, Bh) + np.matmul(Ah, Bl)
    if products == 3:
        out = out + np.matmul(Al, Bh)
    return out.astype(np.float32)


def cplx_stack_A(Z):
    """conj(Z) as the real 2M x 2R matrix [[Zr, Zi], [-Zi, Zr]] (batched over axis 0)."""
    return np.concatenate([np.concatenate([Z.real, Z.imag], axis=2), np.concatenate([-Z.imag, Z.real], axis=2)], axis=1)


class FFTPath:
    def __init__(self, X, L, B, products=3):
        N, T = X.shape
        self.N, self.T, self.L, self.B, self.V = N, T, L, B, B - L + 1
        self.nblk = -(-T // self.V)
        self.products = products
        Xp = np.zeros((N, self.nblk * self.V + B), np.float32)
        Xp[:, :T] = X
        idx = (np.arange(self.nblk) * self.V)[:, None] + np.arange(B)[None, :]
        blocks = Xp[:, idx]                                    # N x nblk x B
        Xf = sfft.rfft(blocks, axis=2).astype(np.complex64)    # N x nblk x F
        self.Xf = np.ascontiguousarray(Xf.transpose(2, 0, 1))  # F x N x nblk

    def numH(self, W):
        """transconv(W, X)[k,t] = sum_l sum_n W[k,n,l] X[n,t+l]."""
        K, N, L = W.shape
        Wf = sfft.rfft(W.astype(np.float32), n=self.B, axis=2).astype(np.complex64).transpose(2, 0, 1)   # F x K x N
        A = cplx_stack_A(Wf)                                                      # F x 2K x 2N
        Bm = np.concatenate([self.Xf.real, self.Xf.imag], axis=1)                 # F x 2N x nblk
        C = mm3(A, Bm, self.products)                                             # F x 2K x nblk
        Cf = (C[:, :K] + 1j * C[:, K:]).transpose(1, 2, 0)                        # K x nblk x F
        out = sfft.irfft(Cf.astype(np.complex64), n=self.B, axis=2)[:, :, : self.V]
        return out.reshape(K, -1)[:, : self.T].astype(np.float64)

    def numW(self, H):
        """corr_w(H, X)[k,n,l] = sum_t H[k,t] X[n,t+l]."""
        K, T = H.shape
        Hp = np.zeros((K, self.nblk * self.V), np.float32)
        Hp[:, :T] = H
        seg = Hp.reshape(K, self.nblk, self.V)
        Hf = sfft.rfft(seg, n=self.B, axis=2).astype(np.complex64).transpose(2, 0, 1)   # F x K x nblk
        A = cplx_stack_A(Hf)                                                            # F x 2K x 2nblk
        Bm = np.concatenate([self.Xf.real, self.Xf.imag], axis=2).transpose(0, 2, 1)    # F x 2nblk x N
        C = mm3(A, Bm, self.products)                                                   # F x 2K x N
        Cf = (C[:, :K] + 1j * C[:, K:]).transpose(1, 2, 0)                              # K x N x F
        return sfft.irfft(Cf.astype(np.complex64), n=self.B, axis=2)[:, :, : self.L].astype(np.float64)


def run(N, T, K, L, B, noise, p_h, iters=100, products=3):
    X, _, _ = po.synthetic_sequences(K=4, N=N, L=L, T=T, noise_scale=noise, p_h=p_h, rng=np.random.default_rng(1234))
    W0, H0 = po.init_rand(X, L, K, np.random.default_rng(0))
    ref = co.fit(co.MultUpdate, X, W0, H0, iters, check_convergence=False)
    fp = FFTPath(X, L, B, products)
    # one-shot operator error
    nH = fp.numH(W0)
    nW = fp.numW(H0)
    eH = np.linalg.norm(nH - po.tensor_transconv(W0, X)) / np.linalg.norm(nH)
    eW = np.linalg.norm(nW - po.corr_w(H0, X, L)) / np.linalg.norm(nW)
    W, H = W0.copy(), H0.copy()
    xn = np.linalg.norm(X)
    hist = [np.linalg.norm(po.tensor_conv(W, H) - X) / xn]
    for _ in range(iters):
        numW = fp.numW(H)
        denW = rs.denomW_gram(W, H)
        W = np.maximum(W * numW / (denW + EPS), EPS)
        numH = fp.numH(W)
        denH = rs.denomH_gram(W, H)
        H = np.maximum(H * numH / (denH + EPS), EPS)
        hist.append(np.linalg.norm(po.tensor_conv(W, H) - X) / xn)
    rel = np.abs(np.asarray(hist) - np.asarray(ref.loss_hist)) / np.asarray(ref.loss_hist)
    dW = np.linalg.norm(W - ref.W) / np.linalg.norm(ref.W)
    dH = np.linalg.norm(H - ref.H) / np.linalg.norm(ref.H)
    print(f"N={N} T={T} K={K} L={L} B={B} noise={noise} p_h={p_h} products={products}: op err numH {eH:.1e} numW {eW:.1e}; "
          f"loss err max {rel.max():.2e} last {rel[-1]:.2e}; dW {dW:.2e} dH {dH:.2e} (final loss {ref.loss_hist[-1]:.4f})",
          flush=True)


if __name__ == "__main__":
    for products in (3, 2):
        run(128, 2048, 4, 10, 64, 1.0, 0.5, products=products)
        run(128, 2048, 4, 10, 64, 0.05, 0.1, products=products)
        run(256, 4096, 8, 16, 64, 0.3, 0.2, iters=60, products=products)

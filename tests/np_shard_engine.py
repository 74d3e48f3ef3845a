"""NumPy float64 engine with the DeviceShard step interface -- TEST INFRASTRUCTURE ONLY.

Lets tests drive cmf_jl_b200.sharded.ShardedMultFit over gloo on the CPU, so the partition /
halo / all-reduce logic is checked against the single-process literal oracle without a GPU.
The per-shard algebra is the one the CUDA library implements (oracle/restructured.py)."""
import numpy as np
import torch

from oracle import cnmf_oracle as po
from oracle import restructured as rs

EPS = po.EPSILON


class NumpyShard:
    def __init__(self, X, W, H, t0, t1, L):
        N, T = X.shape
        K = W.shape[0]
        self.N, self.T, self.K, self.L, self.t0, self.t1 = N, T, K, L, t0, t1
        self.Tl = Tl = t1 - t0
        h = L - 1
        self.is_last = t1 == T
        self.Xe = np.zeros((N, Tl + h))
        hi = min(t1 + h, T)
        self.Xe[:, : hi - t0] = X[:, t0:hi]
        # H stored [t][K] (t-major) so halo regions are contiguous, like the device layout
        self.Hb = np.zeros((Tl + 2 * h, K))
        lo, hi = max(t0 - h, 0), min(t1 + h, T)
        self.Hb[lo - t0 + h : hi - t0 + h] = H[:, lo:hi].T
        self.W = W.copy()
        self.numW = np.zeros(K * N * L)
        self.ex1 = np.zeros(L * K * K + h * K)
        self.exchange = [torch.from_numpy(self.numW), torch.from_numpy(self.ex1)]
        hb = torch.from_numpy(self.Hb)
        self.recv_left, self.send_left = hb[:h].reshape(-1), hb[h : 2 * h].reshape(-1)
        self.send_right, self.recv_right = hb[Tl : Tl + h].reshape(-1), hb[Tl + h :].reshape(-1)
        self.data_norm = None

    def scalar_tensor(self, values):
        return torch.tensor(values, dtype=torch.float64)

    # views
    def Hown(self):
        h = self.L - 1
        return self.Hb[h : h + self.Tl].T  # K x Tl

    def data_sumsq(self):
        return float(np.sum(self.Xe[:, : self.Tl] ** 2))

    def set_data_norm(self, v):
        self.data_norm = v

    def init_scale_partials(self):
        est = self._conv_owned()
        X = self.Xe[:, : self.Tl]
        return [float(np.vdot(X, est)), float(np.vdot(est, est))]

    def scale_factors(self, s):
        self.W *= s
        self.Hb *= s

    def _conv_owned(self):
        K, N, L, Tl, h = self.K, self.N, self.L, self.Tl, self.L - 1
        Hl = self.Hb[: h + Tl].T  # K x (h+Tl): left halo + owned
        return po.tensor_conv(self.W, Hl)[:, h:]

    def w_partials(self):
        K, N, L, Tl, h = self.K, self.N, self.L, self.Tl, self.L - 1
        Ho = self.Hown()
        num = np.zeros((K, N, L))
        for l in range(L):
            num[:, :, l] = Ho @ self.Xe[:, l : l + Tl].T
        self.numW[:] = rs.unfold_W(num).reshape(-1)  # [(l,k)][n]
        Hr = self.Hb[h:].T  # owned + right halo, K x (Tl+h)
        Rg = np.zeros((L, K, K))
        for d in range(L):
            Rg[d] = Ho @ Hr[:, d : d + Tl].T
        self.ex1[: L * K * K] = Rg.reshape(-1)
        tail = self.Hb[Tl + h - h : Tl + h] if self.is_last else np.zeros((h, K))  # last L-1 owned cols
        self.ex1[L * K * K :] = tail.reshape(-1)

    def w_apply(self, l1W, l2W):
        K, N, L, h = self.K, self.N, self.L, self.L - 1
        Rg = self.ex1[: L * K * K].reshape(L, K, K)
        Ht = self.ex1[L * K * K :].reshape(h, K)
        G = np.zeros((L * K, L * K))
        for l in range(L):
            for lp in range(L):
                base = Rg[l - lp] if l >= lp else Rg[lp - l].T
                tail = np.zeros((K, K))
                for i in range(min(l, lp)):
                    tail += np.outer(Ht[h - l + i], Ht[h - lp + i])
                G[l * K : (l + 1) * K, lp * K : (lp + 1) * K] = base - tail
        Wu = rs.unfold_W(self.W)
        den = G @ Wu
        num = self.numW.reshape(L * K, N)
        Wu = Wu * num / (den + l1W + 2 * l2W * Wu + EPS)
        self.W = rs.fold_W(np.maximum(Wu, EPS), K, L)

    def h_update(self, l1H, l2H):
        K, N, L, Tl, h = self.K, self.N, self.L, self.Tl, self.L - 1
        numH = np.zeros((K, Tl))
        for l in range(L):
            numH += self.W[:, :, l] @ self.Xe[:, l : l + Tl]
        C = rs.Cw_tables(self.W)
        denH = np.zeros((K, Tl))
        Hb = self.Hb.T  # K x (Tl+2h); column c <-> local t = c - h
        for t in range(Tl):
            w = min(L, self.T - (self.t0 + t))
            denH[:, t] = np.einsum("kjd,jd->k", C[w - 1], Hb[:, t : t + 2 * L - 1])
        Ho = self.Hown()
        Hn = np.maximum(Ho * numH / (denH + l1H + 2 * l2H * Ho + EPS), EPS)
        self.Hb[h : h + Tl] = Hn.T

    def loss_partial(self):
        r = self._conv_owned() - self.Xe[:, : self.Tl]
        return float(np.sum(r * r))

    def get_factors(self):
        return self.W.copy(), self.Hown().copy()

"""TEST INFRASTRUCTURE (not product code): CPU replay of the round schedule of the second-generation HALS H sweep
(cmf.jl_b200/csrc/kernels_hals.cuh).

The CUDA kernel is bulk-synchronous: in round s every CTA reads only what earlier rounds wrote, and one grid barrier
separates the rounds.  This module replays the same roles, rounds and index arithmetic in NumPy (float64) with a write
stamp on every array element, asserts on every read that the value was written in an EARLIER round (and, for the ring of
block partials, that it still belongs to the chunk being read), and returns the new H -- which must equal the sequential
k-outer / t-inner sweep of src/algs/hals.jl:121-154 (oracle.restructured.hals_H_sweep_gram).  Chunk width, group size and
stagger are parameters so that small problems exercise many chunks, groups and ring wrap-arounds.
"""
import numpy as np

from .restructured import Cw_tables

EPS = np.finfo(np.float64).eps


class Stamped:
    """Array with a per-element write round (and an optional tag, e.g. the chunk a ring slot holds)."""

    def __init__(self, shape, fill=0.0):
        self.v = np.full(shape, fill, dtype=np.float64)
        self.r = np.full(shape, -1, dtype=np.int64)       # -1: written before the kernel starts (prepare kernel)
        self.tag = np.full(shape, -1, dtype=np.int64)

    def write(self, idx, val, rnd, tag=-1):
        self.v[idx] = val
        self.r[idx] = rnd
        self.tag[idx] = tag

    def read(self, idx, rnd, tag=None):
        assert np.all(self.r[idx] < rnd), f"read in round {rnd} of data written in round {np.max(self.r[idx])}"
        if tag is not None:
            assert np.all(self.tag[idx] == tag), "ring slot holds another chunk"
        return self.v[idx]


def sweep(Q, H, W, l1=0.0, l2=0.0, CW=16, GS=4, STAG=4):
    """Q: K x T gradient at the start of the sweep, H: K x T, W: K x N x L.  Returns (new H, number of rounds)."""
    K, T = H.shape
    L = W.shape[2]
    Lm = L - 1
    assert Lm <= CW, "a window reaches one chunk to either side"
    C = Cw_tables(W)                                       # C[w-1, k_src, k_tgt, dd + L-1]
    Cint = C[L - 1]
    nC = -(-T // CW)
    Tp = nC * CW
    Tint = T - Lm
    G = -(-K // GS)
    RING = STAG * GS
    n_rounds = nC + STAG * (K - 1) + 3
    AD = Stamped((K, Tp))
    Hcm = Stamped((K, Tp))
    AD.v[:, :T] = Q
    Hcm.v[:, :T] = H                                       # prepare kernel
    part = Stamped((G, K, RING, CW))

    def window(k_src, c, rnd):
        """Delta H of one source over columns [c*CW - Lm, (c+1)*CW + Lm), zero outside [0, Tint)."""
        t = np.arange(c * CW - Lm, (c + 1) * CW + Lm)
        ok = (t >= 0) & (t < Tint)
        out = np.zeros(t.shape)
        if ok.any():
            out[ok] = AD.read((k_src, t[ok]), rnd)
        return out

    def pull(win, k_src, k_tgt):
        """sum_dd D[k_src, t'-dd] * Cint[k_src, k_tgt, dd] for the CW target columns of the window's chunk."""
        out = np.zeros(CW)
        for j in range(2 * L - 1):                         # dd = j - Lm, source column t = t' + Lm - j -> window index col + 2 Lm - j
            out += win[2 * Lm - j: 2 * Lm - j + CW] * Cint[k_src, k_tgt, j]
        return out

    pend = np.zeros((K, L))                                # pending window of every lane: pend[k][s] for column (pos + s)
    c0 = np.array([Cint[k, k, Lm] for k in range(K)])
    inv = 1.0 / (c0 + EPS + l2)

    for s in range(n_rounds):
        # ---- block items (g, g'), g' < g
        for g in range(1, G):
            c = s + 1 - STAG * GS * g
            if not (0 <= c < nC):
                continue
            for gp in range(g):
                srcs = [k for k in range(gp * GS, min(K, gp * GS + GS))]
                wins = {k: window(k, c, s) for k in srcs}
                for kt in range(g * GS, min(K, g * GS + GS)):
                    acc = np.zeros(CW)
                    for ks in srcs:
                        acc += pull(wins[ks], ks, kt)
                    part.write((gp, kt, c % RING), acc, s, tag=c)
        # ---- diagonal items
        for g in range(G):
            for i in range(GS):
                k = g * GS + i
                c = s - STAG * k
                if k >= K or not (0 <= c < nC):
                    continue
                t0 = c * CW
                qsum = np.zeros(CW)
                for ks in range(g * GS, k):
                    qsum += pull(window(ks, c, s), ks, k)
                if k > 0 and Lm > 0 and t0 + CW + Lm > Tint:       # sources in the truncated tail, all earlier components
                    for tp in range(max(t0, Tint - Lm), min(t0 + CW, T)):
                        for t in range(max(tp - Lm, Tint, 0), min(tp + Lm, T - 1) + 1):
                            w = T - t
                            for kp in range(k):
                                qsum[tp - t0] += AD.read((kp, t), s) * C[w - 1, kp, k, (tp - t) + Lm]
                cols = np.arange(CW)
                ok = t0 + cols < T
                q = AD.read((k, t0 + cols[ok]), s).copy()
                assert np.all(AD.r[k, t0 + cols[ok]] == -1), "Q of this cell must still be the prepared value"
                for gp in range(g):
                    q += part.read((gp, k, c % RING), s, tag=c)[ok]
                q += qsum[ok]
                h = Hcm.read((k, t0 + cols[ok]), s)
                interior = (t0 + cols[ok]) < Tint
                AD.write((k, t0 + cols[ok]), np.where(interior, (h * c0[k] - q - l1) * inv[k], q), s)
        # ---- recurrence lanes
        for k in range(K):
            c = s - 1 - STAG * k
            if 0 <= c < nC:
                t0 = c * CW
                nv = int(np.clip(Tint - t0, 0, CW))
                if nv > 0:
                    a = AD.read((k, np.arange(t0, t0 + nv)), s).copy()
                    h = Hcm.read((k, np.arange(t0, t0 + nv)), s).copy()
                    assert np.all(AD.r[k, t0:t0 + nv] == s - 1), "hand-over must come from the diagonal item of the previous round"
                    for u in range(nv):
                        vn = max(a[u] - pend[k][0], 0.0)
                        d = vn - h[u]
                        h[u] = vn
                        a[u] = d
                        pend[k] = np.append(pend[k][1:], 0.0)       # slot 0 now stands for the next column
                        for j in range(1, L):
                            pend[k][j - 1] += d * Cint[k, k, j + Lm] * inv[k]
                    Hcm.write((k, np.arange(t0, t0 + nv)), h, s)
                    AD.write((k, np.arange(t0, t0 + nv)), a, s)
            # ---- tail job: the last L-1 columns with the truncated tables
            if Lm > 0 and s == nC + STAG * k + 1:
                for t in range(max(Tint, 0), T):
                    w = T - t
                    c0w = C[w - 1, k, k, Lm]
                    pe = 0.0
                    for sft in range(1, Lm + 1):
                        ts = t - sft
                        if ts < 0:
                            break
                        ws = min(L, T - ts)
                        d = AD.v[k, ts] if AD.r[k, ts] == s else AD.read((k, ts), s)      # own writes of this round are visible
                        pe += d * C[ws - 1, k, k, sft + Lm]
                    h = Hcm.read((k, t), s)
                    q = AD.read((k, t), s) + pe
                    vn = max((h * c0w - q - l1) / (c0w + EPS + l2), 0.0)
                    Hcm.v[k, t] = vn                                 # same thread: later reads of this job see it
                    AD.v[k, t] = vn - h
                    Hcm.r[k, t] = s
                    AD.r[k, t] = s
    assert np.all(Hcm.r[:, :T] >= 0), "every column of every component must have been processed"
    return Hcm.v[:, :T].copy(), n_rounds

"""Gram / recurrence restatements of the MU and HALS half-steps  --  TEST INFRASTRUCTURE ONLY.

Same access contract as oracle/cnmf_oracle.py.  These NumPy forms are the *algebra* the CUDA
kernels implement; tests/test_oracle.py proves each one equal to the literal
restatement of the reference to ~1e-13, so that a kernel checked against either is checked
against the reference's arithmetic (src/algs/mult.jl:23-58, src/algs/hals.jl:31-154).

Index conventions (0-based): j = l*K + k is the unfolded row (src/algs/hals.jl:102,
src/common.jl:133-142).
"""
from __future__ import annotations

import numpy as np

from .cnmf_oracle import EPSILON, corr_w, tensor_conv, tensor_transconv


# ---------------------------------------------------------------------------- W side
def gram_R(H, L):
    """R[k, k', d] = sum_{u=0}^{T-1-d} H[k,u] H[k',u+d],  d = 0..L-1 (cross-correlation of H rows)."""
    K, T = H.shape
    R = np.zeros((K, K, L))
    for d in range(min(L, T)):
        R[:, :, d] = H[:, : T - d] @ H[:, d:].T
    return R


def build_G(H, L):
    """G = Htilde Htilde' (KL x KL) from the Toeplitz part R and the end-of-T tail:
    G[(l,k),(l',k')] = R[k,k',l-l'] - sum_{i<min(l,l')} H[k,T-l+i] H[k',T-l'+i]   (l >= l'),
    symmetric for l < l'."""
    K, T = H.shape
    R = gram_R(H, L)
    G = np.zeros((L * K, L * K))
    for l in range(L):
        for lp in range(L):
            if l >= lp:
                base = R[:, :, l - lp]
            else:
                base = R[:, :, lp - l].T
            m = min(l, lp)
            tail = np.zeros((K, K))
            for i in range(m):
                tail += np.outer(H[:, T - l + i], H[:, T - lp + i])
            G[l * K : (l + 1) * K, lp * K : (lp + 1) * K] = base - tail
    return G


def unfold_W(W):
    """W[k,n,l] -> Wu[(l*K+k), n]."""
    K, N, L = W.shape
    return W.transpose(2, 0, 1).reshape(L * K, N)


def fold_W(Wu, K, L):
    return Wu.reshape(L, K, -1).transpose(1, 2, 0)


def denomW_gram(W, H):
    """corr_w(H, conv(W,H)) == fold(G * Wu)   (replaces mult.jl:28+33)."""
    K, N, L = W.shape
    return fold_W(build_G(H, L) @ unfold_W(W), K, L)


# ---------------------------------------------------------------------------- H side
def lag_gram_S(W):
    """S[k, k', l, l'] = sum_n W[k,n,l] W[k',n,l']."""
    return np.einsum("knl,jnm->kjlm", W, W)


def Cw_tables(W):
    """C[w-1, k, k', d + L-1] = sum_{l < w, 0 <= l-d < L} S[k,k',l,l-d],  w = 1..L, d in (-L, L).
    C[L-1] is the interior table; C[w-1] serves columns t with T - t = w < L (truncated tail)."""
    K, N, L = W.shape
    S = lag_gram_S(W)
    C = np.zeros((L, K, K, 2 * L - 1))
    for w in range(1, L + 1):
        for l in range(w):
            for lp in range(L):
                C[w - 1, :, :, (l - lp) + L - 1] += S[:, :, l, lp]
    return C


def denomH_gram(W, H):
    """transconv(W, conv(W,H)) == sum_{k',d} C_{w(t)}[k,k',d] Hpad[k', t+d]  (replaces mult.jl:44+48)."""
    K, N, L = W.shape
    T = H.shape[1]
    C = Cw_tables(W)
    Hp = np.zeros((K, T + 2 * (L - 1)))
    Hp[:, L - 1 : L - 1 + T] = H
    out = np.zeros((K, T))
    for t in range(T):
        w = min(L, T - t)
        win = Hp[:, t : t + 2 * L - 1]  # Hpad[k', t + d], d = -(L-1)..L-1
        out[:, t] = np.einsum("kjd,jd->k", C[w - 1], win)
    return out


def mu_iteration_gram(data, W, H, l1W=0.0, l2W=0.0, l1H=0.0, l2H=0.0):
    """One MU iteration in the form the CUDA path executes: numW by direct correlation,
    denomW = G*W, numH by direct transposed conv, denomH = C (*) H, loss by direct conv."""
    K, N, L = W.shape
    numW = corr_w(H, data, L)
    denW = denomW_gram(W, H)
    W = W * numW / (denW + l1W + 2 * l2W * W + EPSILON)
    W = np.maximum(W, EPSILON)
    numH = tensor_transconv(W, data)
    denH = denomH_gram(W, H)
    H = H * numH / (denH + l1H + 2 * l2H * H + EPSILON)
    H = np.maximum(H, EPSILON)
    loss = np.linalg.norm(tensor_conv(W, H) - data) / np.linalg.norm(data)
    return W, H, float(loss)


def loss_expansion(data, W, H):
    """||conv(W,H) - X||^2 = ||X||^2 - 2 <transconv(W,X), H> + <H, C (*) H>."""
    numH = tensor_transconv(W, data)
    denH = denomH_gram(W, H)
    v = np.vdot(data, data) - 2 * np.vdot(numH, H) + np.vdot(H, denH)
    return float(np.sqrt(max(v, 0.0)) / np.linalg.norm(data))


# ---------------------------------------------------------------------------- HALS
def hals_W_sweep_gram(R, W, H, l1W=0.0, l2W=0.0):
    """hals.jl:90-112 restated on P = R Htilde' and G: rows n are independent;
    w_new = max((w_old*g - P[n,j] - l1)/(g + eps + l2), 0), then P[n,:] += (w_new-w_old) G[j,:].
    Returns the new W (R is NOT updated here; the caller refreshes it with one conv)."""
    K, N, L = W.shape
    G = build_G(H, L)
    P = unfold_W(corr_w(H, R, L)).T.copy()  # N x KL : P[n, j] = sum_t R[n,t] Htilde[j,t]
    Wu = unfold_W(W).T.copy()  # N x KL
    for k in range(K):
        for l in range(L):
            j = l * K + k
            g = G[j, j]
            new = np.maximum((Wu[:, j] * g - P[:, j] - l1W) / (g + EPSILON + l2W), 0.0)
            delta = new - Wu[:, j]
            P += np.outer(delta, G[j])
            Wu[:, j] = new
    return fold_W(Wu.T, K, L)


def hals_H_sweep_gram(R, W, H, l1H=0.0, l2H=0.0):
    """hals.jl:121-154 restated on Q = transconv(W, R) and the C_w tables:
    h_new = max((h_old*c0 - Q[k,t] - l1)/(c0 + eps + l2), 0), c0 = C_w[k,k,0];
    Q[k',t'] += (h_new-h_old) * C_{w(t)}[k,k',t'-t].  Returns the new H."""
    K, N, L = W.shape
    T = H.shape[1]
    C = Cw_tables(W)
    Q = tensor_transconv(W, R)
    H = H.copy()
    for k in range(K):
        for t in range(T):
            w = min(L, T - t)
            c0 = C[w - 1, k, k, L - 1]
            new = max((H[k, t] * c0 - Q[k, t] - l1H) / (c0 + EPSILON + l2H), 0.0)
            delta = new - H[k, t]
            H[k, t] = new
            if delta != 0.0:
                lo, hi = max(0, t - (L - 1)), min(T - 1, t + (L - 1))
                # Q[k', t'] += delta * C_w[k, k', d = t' - t]
                Q[:, lo : hi + 1] += delta * C[w - 1, k, :, (lo - t) + L - 1 : (hi - t) + L]
    return H


def hals_iteration_gram(data, W, H, l1W=0.0, l2W=0.0, l1H=0.0, l2H=0.0):
    """One HALS iteration in the form the CUDA path executes (residual recomputed by conv)."""
    R = tensor_conv(W, H) - data
    W = hals_W_sweep_gram(R, W, H, l1W, l2W)
    R = tensor_conv(W, H) - data
    H = hals_H_sweep_gram(R, W, H, l1H, l2H)
    R = tensor_conv(W, H) - data
    return W, H, float(np.linalg.norm(R) / np.linalg.norm(data))


# ---------------------------------------------------------------------------- frequency-domain (overlap-save) forms
# The algebra of the device's frequency-domain engine (kernels_fd.cuh) in float64 NumPy: blocks
# of length B with hop V = B-L+1 (numH, numW, Gram) or V2 = B-2L+2 (denomH), per-frequency products with the conjugate
# spectrum of the small operand, inverse transform, keep the valid samples.  tests/test_oracle.py proves each one equal
# to the literal restatement, which pins the block/mask/offset conventions the CUDA kernels follow.
def _blocks(A, B, hop, nblk, start=0):
    """A[:, start + b*hop + i] for b < nblk, i < B, zero outside A's columns -> rows x nblk x B."""
    rows, T = A.shape
    idx = start + (np.arange(nblk) * hop)[:, None] + np.arange(B)[None, :]
    ok = (idx >= 0) & (idx < T)
    return np.where(ok[None], A[:, np.clip(idx, 0, T - 1)], 0.0)


def fd_block_length(L):
    B = 64
    while B < 4 * L:
        B *= 2
    return B


def numH_overlap_save(W, X, B=None):
    """transconv(W, X) (common.jl:71-81): numH^[k,b,f] = sum_n conj(W^[k,n,f]) X^[n,b,f], first V samples of each block."""
    K, N, L = W.shape
    T = X.shape[1]
    B = B or fd_block_length(L)
    V = B - L + 1
    nblk = -(-T // V)
    Xf = np.fft.rfft(_blocks(X, B, V, nblk), axis=2)                 # N x nblk x F
    Wf = np.fft.rfft(W, n=B, axis=2)                                  # K x N x F
    out = np.fft.irfft(np.einsum("knf,nbf->kbf", Wf.conj(), Xf), n=B, axis=2)[:, :, :V]
    return out.reshape(K, -1)[:, :T]


def numW_overlap_save(H, X, L, B=None):
    """corr_w(H, X) (mult.jl:31-34): numW^[k,n,f] = sum_b conj(Hz^[k,b,f]) X^[n,b,f] with Hz the V owned columns of block b
    zero padded to B; first L lags."""
    K, T = H.shape
    B = B or fd_block_length(L)
    V = B - L + 1
    nblk = -(-T // V)
    Xf = np.fft.rfft(_blocks(X, B, V, nblk), axis=2)
    Hz = _blocks(H, B, V, nblk)
    Hz[:, :, V:] = 0.0
    Hf = np.fft.rfft(Hz, axis=2)                                      # K x nblk x F
    return np.fft.irfft(np.einsum("kbf,nbf->knf", Hf.conj(), Xf), n=B, axis=2)[:, :, :L]


def gram_overlap_save(H, L, B=None):
    """gram_R(H, L): the numW form with the full blocks of H itself in place of X."""
    return numW_overlap_save(H, H, L, B)                               # R[k, k', d]


def denomH_interior_overlap_save(W, H, B=None):
    """C (*) H with the interior lag table (all columns; the last L-1 are only valid where the conv is not truncated):
    2L-1 lags d' = d + L-1 over blocks of hop V2 = B-2L+2 that start L-1 columns early."""
    K, N, L = W.shape
    T = H.shape[1]
    B = B or fd_block_length(L)
    V2 = B - 2 * L + 2
    nblk = -(-T // V2)
    C = Cw_tables(W)[L - 1]                                            # K x K x (2L-1): C[k, k', d + L-1]
    Cf = np.fft.rfft(C, n=B, axis=2)
    Hf = np.fft.rfft(_blocks(H, B, V2, nblk, start=-(L - 1)), axis=2)  # K x nblk x F
    out = np.fft.irfft(np.einsum("kjf,jbf->kbf", Cf.conj(), Hf), n=B, axis=2)[:, :, :V2]
    return out.reshape(K, -1)[:, :T]


def conv_overlap_save(W, H, B=None):
    """tensor_conv(W, H) (common.jl:24-34), the way the device's direct loss pass forms it: Xhat^[n,b,f] = sum_k W^[k,n,f]
    H^[k,b,f] over full blocks of H that start L-1 columns early (hop V = B-L+1, zero history before column 0); the last V
    samples of each inverse transform are exact."""
    K, N, L = W.shape
    T = H.shape[1]
    B = B or fd_block_length(L)
    V = B - L + 1
    nblk = -(-T // V)
    Hf = np.fft.rfft(_blocks(H, B, V, nblk, start=-(L - 1)), axis=2)   # K x nblk x F
    Wf = np.fft.rfft(W, n=B, axis=2)                                    # K x N x F
    y = np.fft.irfft(np.einsum("knf,kbf->nbf", Wf, Hf), n=B, axis=2)[:, :, L - 1:]
    return y.reshape(N, -1)[:, :T]


def build_G_from_R(R, H, L):
    """build_G with the Toeplitz part given (e.g. from gram_overlap_save)."""
    K, T = H.shape
    G = np.zeros((L * K, L * K))
    for l in range(L):
        for lp in range(L):
            base = R[:, :, l - lp] if l >= lp else R[:, :, lp - l].T
            tail = np.zeros((K, K))
            for i in range(min(l, lp)):
                tail += np.outer(H[:, T - l + i], H[:, T - lp + i])
            G[l * K : (l + 1) * K, lp * K : (lp + 1) * K] = base - tail
    return G


def mu_iteration_overlap_save(data, W, H, l1W=0.0, l2W=0.0, l1H=0.0, l2H=0.0, B=None):
    """One MU iteration exactly as the frequency-domain engine executes it (float64): numW, the Gram of H, numH, denomH and the
    direct loss all through overlap-save blocks; G*W, the truncated tail of denomH and the ratio updates in the time domain."""
    K, N, L = W.shape
    T = H.shape[1]
    numW = numW_overlap_save(H, data, L, B)
    G = build_G_from_R(gram_overlap_save(H, L, B), H, L)
    denW = fold_W(G @ unfold_W(W), K, L)
    W = np.maximum(W * numW / (denW + l1W + 2 * l2W * W + EPSILON), EPSILON)
    numH = numH_overlap_save(W, data, B)
    denH = denomH_interior_overlap_save(W, H, B)
    if L > 1:
        denH[:, T - (L - 1):] = denomH_gram(W, H)[:, T - (L - 1):]       # truncated tail (denomH_tail_prefix_kernel)
    H = np.maximum(H * numH / (denH + l1H + 2 * l2H * H + EPSILON), EPSILON)
    loss = np.linalg.norm(conv_overlap_save(W, H, B) - data) / np.linalg.norm(data)
    return W, H, float(loss)

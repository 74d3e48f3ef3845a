// HALS H sweep (src/algs/hals.jl:121-154), second generation: one cooperative launch organised in ROUNDS with a grid-wide
// barrier between them (fp32, L <= 32).  Same arithmetic order as the reference's k-outer / t-inner sweep per component;
// the corrections that earlier components send to later ones are summed in a fixed order (deterministic).
//
// Recurrence form (oracle/restructured.py): with Q = transconv(W, conv(W,H) - X) at the start of the
// sweep, C[k',k,dd] the lag table and D = Delta H,
//     qeff[k,t]  = Q[k,t] + sum_{k'<k} sum_{|dd|<L} D[k',t-dd] C_w[k',k,dd]          ("pull", needs all earlier components)
//     h'[k,t]    = max((h c0 - qeff - pend(t) - l1) / (c0 + eps + l2), 0),   pend(t+s) += D[k,t] C[k,k,s], s = 1..L-1
//
// Why rounds.  The first-generation kernel gave one CTA per component and let each CTA stage, pull and run its recurrence
// one after the other; a cell took as long as the LAST component's pull (k-1 predecessors) and its one-lane recurrence.
// Here the three kinds of work are different CTAs that all run in every round:
//   * recurrence warps: one warp per 32 components, LANE = COMPONENT.  Every lane runs the sequential recurrence of its own
//     component over one chunk of CW columns per round; the pending window and the lag table of the component live in
//     registers, so a column costs ~L issue slots for 32 components at once (the old kernel spent the same slots on one).
//   * block items (g, g'), g' < g: the pull of the 8 targets of group g from the 8 sources of group g' for one chunk: an
//     8 x 8 block of component pairs on one staged window of Delta H, register tile 8 columns x 4 targets per thread.
//   * diagonal items (g): for every target of group g, the pull from the earlier components of its own group, the sum of
//     the block partials in the fixed order g' = 0, 1, ..., and the hand-over to the recurrence: a = (h c0 - qeff - l1) inv.
// Component k works on chunk c in round c + 4k + 1 (stagger of 4 chunks): the block items of (group g, chunk c) run in round
// c + 32g - 1 (their last source, component 8g-1, finished chunk c+1 in round c + 32g - 2), the diagonal item of (k, c) in
// round c + 4k (component k-1 finished chunk c+1 in round c + 4k - 2), the recurrence of (k, c) in round c + 4k + 1, and the
// truncated last L-1 columns of component k (tables C_w, w < L) in round nC + 4k + 1 by one thread.  Every read in a round is of
// data written in an EARLIER round, so a round is a bulk-synchronous step and one barrier per round is all the ordering
// there is; tests/test_oracle.py replays this schedule on the CPU and checks every read against the write rounds.
// For K = 128 that is 120 + 16 + 4 = 140 CTAs on 148 SMs, each with the same work in every round.
//
// Working layout: component-major (H_cm[k][t], AD_cm[k][t]) so that a lane / a staged source reads contiguous memory;
// the [t][K] arrays of the rest of the library are transposed on the way in (diagonal items) and out (hals2_finish_kernel).
#pragma once

#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cmf {
namespace hals2 {

namespace cg = cooperative_groups;

constexpr int CW = 1024;        // columns per chunk
constexpr int NT = 256;         // threads of an item CTA
constexpr int GS = 8;           // components per group
constexpr int STAG = 4;         // chunks between consecutive components
constexpr int RING = 32;        // chunks a block partial stays alive: written in round c+32g-1, read no later than c+32g+28
constexpr int LMAX = 32;
constexpr int WIN = CW + 2 * (LMAX - 1);      // staged window of Delta H (columns t0-(L-1) .. t0+CW+(L-1)-1), sized for L = 32
constexpr int QP = ((WIN + 7) >> 3) | 1;      // slots per plane of the 8-plane layout (odd: conflict-free)
constexpr int WINQ = 8 * QP;                  // words per staged source

struct Args {
    const float *Cf;      // [(dd+L-1)][k'][k]  interior lag table
    const float *Ct;      // [w-1][dd+L-1][k][k'] truncated tables (w = 1..L-1), nullptr when L == 1
    const float *S2;      // W W' (truncated self tables of the tail job)
    int64_t Ks, ld;       // addressing of S2
    float *H_cm;          // [K][Tp]  H, component-major (old value until the recurrence overwrites it with the new one)
    float *AD_cm;         // [K][Tp]  Q (from hals2_prepare_kernel), then a (hand-over to the recurrence), then Delta H
    float *part;          // [G][K][RING][CW] block partials, indexed by source group
    int64_t K, L, T, Tp, nC;
    float l1, l2;
    int G;                // groups of 8 components
    int n_block_items, n_diag_items, n_rec;
};

__device__ __forceinline__ float ldcg(const float *p) { return __ldcg(p); }

// Stages Delta H of `nsrc` sources (components k0 .. k0+nsrc-1) over the window of chunk c into the 8-plane layout.
// Columns outside [0, Tint) read as zero (sources in the truncated tail go through the tail tables instead).
__device__ __forceinline__ void stage_window(float *Dwin, const Args &a, int64_t k0, int nsrc, int64_t c, int64_t Tint) {
    const int Lm = (int)a.L - 1, WW = CW + 2 * Lm;
    const int64_t tbase = c * CW - Lm;
    for (int idx = threadIdx.x; idx < nsrc * WW; idx += NT) {
        const int s = idx / WW, i = idx - s * WW;
        const int64_t t = tbase + i;
        float v = 0.f;
        if (t >= 0 && t < Tint) v = ldcg(a.AD_cm + (k0 + s) * a.Tp + t);
        Dwin[s * WINQ + (i & 7) * QP + (i >> 3)] = v;
    }
}

// acc[r][q] += sum over the staged sources and lags of D[src][t'+(L-1)-j] * tab[src][j][q],  t' = 8*cgp + r, q = 0..NTG-1.
// tab: [nsrc][NLP][8] floats (lag rows padded with zeros to NLP = a multiple of 8), tq0 = first target column of the table.
template <int NTG>
__device__ __forceinline__ void pull_tile(float (&acc)[8][NTG], const float *Dwin, const float *tab, int nsrc, int L, int cgp, int tq0) {
    const int nl = 2 * L - 1, NLP = (nl + 7) & ~7;
    const int ib = 8 * cgp + 2 * (L - 1);
    for (int s = 0; s < nsrc; ++s) {
        const float *dw = Dwin + s * WINQ;
        const float *tb = tab + (size_t)s * NLP * 8 + tq0;
        float w[8];
        w[0] = 0.f;
#pragma unroll
        for (int o = 1; o < 8; ++o) { const int i = ib + o; w[o] = dw[(i & 7) * QP + (i >> 3)]; }
        for (int j0 = 0; j0 < nl; j0 += 8) {
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int j = j0 + jj;
                const int in = max(ib - j, 0);              // window index entering at this step (column r = 0); padded lags multiply by zero
                w[(8 - jj) & 7] = dw[(in & 7) * QP + (in >> 3)];
                float cq[NTG];
#pragma unroll
                for (int q = 0; q < NTG; ++q) cq[q] = tb[j * 8 + q];
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int q = 0; q < NTG; ++q) acc[r][q] = fmaf(w[(r - jj + 8) & 7], cq[q], acc[r][q]);
            }
        }
    }
}

// table slice tab[s][j][q] = C[k0s+s, k0t+q, j-(L-1)] (zero where the pair is not a (source < target) pair or out of range)
__device__ __forceinline__ void load_table(float *tab, const Args &a, int64_t k0s, int64_t k0t, bool diagonal) {
    const int L = (int)a.L, nl = 2 * L - 1, NLP = (nl + 7) & ~7;
    for (int idx = threadIdx.x; idx < GS * NLP * 8; idx += NT) {
        const int q = idx & 7, j = (idx >> 3) % NLP, s = idx / (8 * NLP);
        const int64_t ks = k0s + s, kt = k0t + q;
        float v = 0.f;
        if (j < nl && ks < a.K && kt < a.K && (!diagonal || ks < kt)) v = a.Cf[((int64_t)j * a.K + ks) * a.K + kt];
        tab[idx] = v;
    }
}

// 16 columns of the lane's recurrence (HH = which half of the 32-slot pending window they occupy: static register indices).
// an / hn hold a and h of these columns on entry and of the next 16 columns on exit (prefetch).
template <int HH>
__device__ __forceinline__ void rec_half(float (&p)[32], const float (&cs)[32], float4 (&an)[4], float4 (&hn)[4], float *adp, float *hp,
                                         int64_t t0, int i0, int nv, bool on, bool more) {
    float av[16], hv[16];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        av[4 * v] = an[v].x; av[4 * v + 1] = an[v].y; av[4 * v + 2] = an[v].z; av[4 * v + 3] = an[v].w;
        hv[4 * v] = hn[v].x; hv[4 * v + 1] = hn[v].y; hv[4 * v + 2] = hn[v].z; hv[4 * v + 3] = hn[v].w;
    }
    if (more && on) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            an[v] = __ldcg(reinterpret_cast<const float4 *>(adp + t0 + i0 + 16) + v);
            hn[v] = __ldcg(reinterpret_cast<const float4 *>(hp + t0 + i0 + 16) + v);
        }
    }
#pragma unroll
    for (int uu = 0; uu < 16; ++uu) {
        constexpr int base = HH * 16;
        const int u = base + uu;
        const bool ok = i0 + uu < nv;
        float vn = av[uu] - p[u];
        vn = vn > 0.f ? vn : 0.f;
        const float d = ok ? vn - hv[uu] : 0.f;
        hv[uu] = ok ? vn : hv[uu];
        av[uu] = ok ? d : av[uu];                       // tail columns keep the raw qeff for the tail job
        p[u] = 0.f;                                     // this slot now stands for column t + 32
#pragma unroll
        for (int j = 1; j < 32; ++j) p[(u + j) & 31] = fmaf(d, cs[j], p[(u + j) & 31]);
    }
    if (on) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            reinterpret_cast<float4 *>(hp + t0 + i0)[v] = make_float4(hv[4 * v], hv[4 * v + 1], hv[4 * v + 2], hv[4 * v + 3]);
            reinterpret_cast<float4 *>(adp + t0 + i0)[v] = make_float4(av[4 * v], av[4 * v + 1], av[4 * v + 2], av[4 * v + 3]);
        }
    }
}

__global__ void __launch_bounds__(NT, 1) hals2_sweep_kernel(Args a) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int L = (int)a.L, Lm = L - 1;
    const int64_t K = a.K, T = a.T, Tp = a.Tp, nC = a.nC;
    const int64_t Tint = T - Lm;                               // columns t < Tint use the full lag window
    const int64_t n_rounds = nC + (int64_t)STAG * (K - 1) + 3;
    const int b = blockIdx.x, tid = threadIdx.x;
    const float eps = 2.220446049250313e-16f;

    if (b < a.n_block_items) {
        // ============================================================ block item (g, g'), g' < g
        int g = 1, rem = b;
        while (rem >= g) { rem -= g; ++g; }                     // items enumerated as (1,0), (2,0), (2,1), (3,0), ...
        const int gp = rem;
        float *Dwin = reinterpret_cast<float *>(smem_raw);      // [8][WINQ]
        float *tab = Dwin + GS * WINQ;                          // [8][NLP][8]
        load_table(tab, a, (int64_t)gp * GS, (int64_t)g * GS, false);
        __syncthreads();
        const int cgp = tid & 127, th = tid >> 7;               // column group (8 columns), half of the targets
        const int64_t nleft = K - (int64_t)gp * GS;
        const int nsrc = (int)(nleft < GS ? nleft : GS);
        for (int64_t s = 0; s < n_rounds; ++s) {
            const int64_t c = s + 1 - (int64_t)STAG * GS * g;
            if (c >= 0 && c < nC) {
                stage_window(Dwin, a, (int64_t)gp * GS, nsrc, c, Tint);
                __syncthreads();
                float acc[8][4];
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[r][q] = 0.f;
                pull_tile<4>(acc, Dwin, tab, nsrc, L, cgp, th * 4);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int64_t kt = (int64_t)g * GS + th * 4 + q;
                    if (kt < K) {
                        float *o = a.part + (((int64_t)gp * K + kt) * RING + (c % RING)) * CW + 8 * cgp;
                        *reinterpret_cast<float4 *>(o) = make_float4(acc[0][q], acc[1][q], acc[2][q], acc[3][q]);
                        *reinterpret_cast<float4 *>(o + 4) = make_float4(acc[4][q], acc[5][q], acc[6][q], acc[7][q]);
                    }
                }
                __syncthreads();
            }
            grid.sync();
        }
    } else if (b < a.n_block_items + a.n_diag_items) {
        // ============================================================ diagonal item (g): finalises qeff of its 8 targets
        const int g = b - a.n_block_items;
        float *Dwin = reinterpret_cast<float *>(smem_raw);      // [8][WINQ]
        float *tab = Dwin + GS * WINQ;                          // [8][NLP][8]
        float *qsum = tab + GS * (((2 * LMAX - 1) + 7) & ~7) * 8;   // [CW]
        load_table(tab, a, (int64_t)g * GS, (int64_t)g * GS, true);
        __syncthreads();
        const int cgp = tid & 127, th = tid >> 7;
        for (int64_t s = 0; s < n_rounds; ++s) {
            for (int i = 0; i < GS; ++i) {
                const int64_t k = (int64_t)g * GS + i;
                const int64_t c = s - (int64_t)STAG * k;
                if (k >= K || c < 0 || c >= nC) continue;      // uniform over the CTA
                const int64_t t0 = c * CW;
                // ---- pull from the earlier components of the own group (sources g*8 .. k-1): the two thread halves split the sources
                float acc[8][1];
#pragma unroll
                for (int r = 0; r < 8; ++r) acc[r][0] = 0.f;
                if (i > 0) {
                    stage_window(Dwin, a, (int64_t)g * GS, i, c, Tint);
                    __syncthreads();
                    const int s_lo = th == 0 ? 0 : (i + 1) / 2, s_hi = th == 0 ? (i + 1) / 2 : i;
                    if (s_hi > s_lo) pull_tile<1>(acc, Dwin + s_lo * WINQ, tab + (size_t)s_lo * (((2 * L - 1) + 7) & ~7) * 8, s_hi - s_lo, L, cgp, i);
                }
                if (th == 1) {
#pragma unroll
                    for (int r = 0; r < 8; ++r) qsum[8 * cgp + r] = acc[r][0];
                }
                __syncthreads();
                if (th == 0) {
#pragma unroll
                    for (int r = 0; r < 8; ++r) qsum[8 * cgp + r] += acc[r][0];
                }
                __syncthreads();
                // ---- truncated tail: sources t >= Tint of ALL earlier components with the tables C_w (w = T - t); work item =
                //      (target column, block of earlier components), block sums added in a fixed order (deterministic)
                if (k > 0 && Lm > 0 && t0 + CW + Lm > Tint) {
                    const int64_t c_lo = (t0 > Tint - Lm) ? t0 : Tint - Lm;
                    const int64_t c_hi = (t0 + CW < T) ? t0 + CW : T;
                    const int ncol = (int)(c_hi - c_lo);          // <= 2 (L-1)
                    const int CH = (int)((k + 15) / 16 > 8 ? (k + 15) / 16 : 8);
                    const int nch = (int)((k + CH - 1) / CH);     // <= 16
                    float *tpart = Dwin;                          // [nch][64]
                    if (ncol > 0) {
                        for (int it = tid; it < ncol * nch; it += NT) {
                            const int ci = it % ncol, ch = it / ncol;
                            const int64_t tp = c_lo + ci;
                            int64_t ta = (tp - Lm > Tint) ? tp - Lm : Tint;
                            if (ta < 0) ta = 0;
                            const int64_t tb = tp + Lm < T - 1 ? tp + Lm : T - 1;
                            const int64_t kp_lo = (int64_t)ch * CH, kp_hi = (kp_lo + CH < k) ? kp_lo + CH : k;
                            double accd = 0.0;
                            for (int64_t t = ta; t <= tb; ++t) {
                                const int64_t dd = tp - t, w = T - t;
                                const float *ct = a.Ct + ((((w - 1) * (2 * a.L - 1) + (dd + Lm)) * K + k) * K);
                                for (int64_t kp = kp_lo; kp < kp_hi; ++kp) accd += (double)ldcg(a.AD_cm + kp * Tp + t) * (double)ct[kp];
                            }
                            tpart[ch * 64 + ci] = (float)accd;
                        }
                        __syncthreads();
                        for (int ci = tid; ci < ncol; ci += NT) {
                            double sacc = 0.0;
                            for (int ch = 0; ch < nch; ++ch) sacc += (double)tpart[ch * 64 + ci];
                            qsum[(int)(c_lo - t0) + ci] += (float)sacc;
                        }
                    }
                    __syncthreads();
                }
                // ---- Q + the block partials in the fixed order g' = 0 .. g-1 + the own-group pull, and the hand-over
                const float c0 = a.Cf[((int64_t)Lm * K + k) * K + k];
                const float inv = 1.f / (c0 + eps + a.l2);
                for (int col = tid; col < CW; col += NT) {
                    const int64_t t = t0 + col;
                    if (t >= T) continue;
                    float q = ldcg(a.AD_cm + k * Tp + t);                                    // Q[k][t]
                    for (int gp = 0; gp < g; ++gp) q += ldcg(a.part + (((int64_t)gp * K + k) * RING + (c % RING)) * CW + col);
                    q += qsum[col];
                    const float h = ldcg(a.H_cm + k * Tp + t);
                    a.AD_cm[k * Tp + t] = (t < Tint) ? (h * c0 - q - a.l1) * inv : q;        // tail columns keep the raw qeff
                }
                __syncthreads();
            }
            grid.sync();
        }
    } else {
        // ============================================================ recurrence warp: lane = component
        const int rb = b - a.n_block_items - a.n_diag_items;
        if (tid >= 32) {                                         // only warp 0 works; the others just keep the barrier count
            for (int64_t s = 0; s < n_rounds; ++s) grid.sync();
            return;
        }
        const int lane = tid;
        const int64_t k = (int64_t)rb * 32 + lane;
        const bool live = k < K;
        float cs[32], p[32];
        float c0 = 1.f, inv = 1.f;
        if (live) {
            c0 = a.Cf[((int64_t)Lm * K + k) * K + k];
            inv = 1.f / (c0 + eps + a.l2);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            cs[j] = (live && j >= 1 && j < L) ? a.Cf[((int64_t)(j + Lm) * K + k) * K + k] * inv : 0.f;
            p[j] = 0.f;
        }
        float *adp = a.AD_cm + (live ? k : 0) * Tp;
        float *hp = a.H_cm + (live ? k : 0) * Tp;
        for (int64_t s = 0; s < n_rounds; ++s) {
            const int64_t c = s - 1 - (int64_t)STAG * k;
            const bool on = live && c >= 0 && c < nC;
            if (__any_sync(0xffffffffu, on)) {
                const int64_t t0 = on ? c * CW : 0;
                // interior columns of this chunk (the truncated tail is left to the tail job)
                int nv = 0;
                if (on) { const int64_t r = Tint - t0; nv = (int)(r < 0 ? 0 : (r > CW ? CW : r)); }
                const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                float4 an[4], hn[4];                             // the next 16 columns (prefetched one half-block ahead: L2 hits)
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    an[v] = on ? __ldcg(reinterpret_cast<const float4 *>(adp + t0) + v) : z4;
                    hn[v] = on ? __ldcg(reinterpret_cast<const float4 *>(hp + t0) + v) : z4;
                }
                for (int i0 = 0; i0 < CW; i0 += 32) {
                    rec_half<0>(p, cs, an, hn, adp, hp, t0, i0, nv, on, i0 + 16 < CW);
                    rec_half<1>(p, cs, an, hn, adp, hp, t0, i0 + 16, nv, on, i0 + 32 < CW);
                }
            }
            // ---- tail job of component k: the last L-1 columns with the truncated tables, one round after its last chunk
            if (live && Lm > 0 && s == nC + (int64_t)STAG * k + 1) {
                for (int64_t t = (Tint > 0 ? Tint : 0); t < T; ++t) {
                    const int64_t w = T - t;                     // 1 .. L-1 lags left
                    // C_w[k,k,sft] = sum_{l<w, l-sft>=0} S2[(l,k)][(l-sft,k)]
                    double c0w = 0.0;
                    for (int64_t l = 0; l < w; ++l) c0w += (double)a.S2[(l * a.Ks + k) * a.ld + l * a.Ks + k];
                    // pending corrections from the earlier columns of this component (interior sources: full table)
                    double pend = 0.0;
                    for (int64_t sft = 1; sft <= Lm && t - sft >= 0; ++sft) {
                        const int64_t ts = t - sft, ws = T - ts;
                        const float d = ldcg(a.AD_cm + k * Tp + ts);
                        if (d == 0.f) continue;
                        double cw;
                        if (ws >= a.L) cw = (double)a.Cf[((int64_t)(sft + Lm) * K + k) * K + k];
                        else { cw = 0.0; for (int64_t l = sft; l < ws; ++l) cw += (double)a.S2[(l * a.Ks + k) * a.ld + (l - sft) * a.Ks + k]; }
                        pend += (double)d * cw;
                    }
                    const float h = ldcg(a.H_cm + k * Tp + t);
                    const float q = ldcg(a.AD_cm + k * Tp + t) + (float)pend;
                    const float c0f = (float)c0w;
                    float vn = (h * c0f - q - a.l1) / (c0f + eps + a.l2);
                    vn = vn > 0.f ? vn : 0.f;
                    a.H_cm[k * Tp + t] = vn;
                    a.AD_cm[k * Tp + t] = vn - h;
                }
            }
            grid.sync();
        }
    }
}

// Q and H into the component-major working arrays (AD_cm <- Q', H_cm <- H')
__global__ void hals2_prepare_kernel(const float *__restrict__ Q, const float *__restrict__ H, float *__restrict__ AD_cm,
                                     float *__restrict__ H_cm, int64_t K, int64_t T, int64_t Tp) {
    __shared__ float tq[32][33], thh[32][33];
    const int64_t t0 = (int64_t)blockIdx.x * 32, k0 = (int64_t)blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int64_t t = t0 + r, k = k0 + threadIdx.x;
        const bool ok = t < T && k < K;
        tq[r][threadIdx.x] = ok ? Q[t * K + k] : 0.f;
        thh[r][threadIdx.x] = ok ? H[t * K + k] : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int64_t k = k0 + r, t = t0 + threadIdx.x;
        if (k < K && t < Tp) { AD_cm[k * Tp + t] = tq[threadIdx.x][r]; H_cm[k * Tp + t] = thh[threadIdx.x][r]; }
    }
}

// new H back to the [t][K] layout
__global__ void hals2_finish_kernel(const float *__restrict__ H_cm, float *__restrict__ H, int64_t K, int64_t T, int64_t Tp) {
    __shared__ float tile[32][33];
    const int64_t t0 = (int64_t)blockIdx.x * 32, k0 = (int64_t)blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int64_t k = k0 + r, t = t0 + threadIdx.x;
        tile[r][threadIdx.x] = (k < K && t < T) ? H_cm[k * Tp + t] : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int64_t t = t0 + r, k = k0 + threadIdx.x;
        if (t < T && k < K) H[t * K + k] = tile[threadIdx.x][r];
    }
}

inline size_t smem_bytes() {
    return (size_t)(GS * WINQ + GS * (((2 * LMAX - 1) + 7) & ~7) * 8 + CW) * sizeof(float);
}

}  // namespace hals2
}  // namespace cmf

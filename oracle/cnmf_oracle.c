/* CPU oracle, plain-C twin  --  TEST INFRASTRUCTURE ONLY (see oracle/cnmf_oracle.py header).
 *
 * PARITY UNPINNED: the reference holds no tests / golden vectors for this path and Julia
 * is not installed here, so this restatement is pinned by the element-wise definitions,
 * adjointness and agreement with the independent NumPy restatement (tests/test_oracle.py).
 *
 * Literal float64 restatement of the reference's loops, same loop orders, same constants.
 * All arrays are Julia column-major:
 *   X[n + N*t]           data / est / resids,  N x T
 *   H[k + K*t]           feature maps,         K x T
 *   W[k + K*(n + N*l)]   motifs,               K x N x L
 * File:line citations are relative to /root/reference.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_EPS 2.220446049250313e-16 /* src/CMF.jl:20  EPSILON = eps() */

typedef int64_t i64;

static inline i64 imin(i64 a, i64 b) { return a < b ? a : b; }

/* src/common.jl:24-34 tensor_conv!: est = 0; for lag: est[:, lag+1:T] += W[:,:,lag+1]' * H[:, 1:T-lag] */
void orc_tensor_conv(const double *W, const double *H, double *est, i64 N, i64 T, i64 K, i64 L) {
    memset(est, 0, sizeof(double) * (size_t)(N * T));
    for (i64 lag = 0; lag < L && lag < T; ++lag) {
        const double *Wl = W + K * N * lag;
#pragma omp parallel for schedule(static)
        for (i64 t = lag; t < T; ++t) {
            const double *h = H + K * (t - lag);
            double *e = est + N * t;
            for (i64 n = 0; n < N; ++n) {
                const double *w = Wl + K * n;
                double s = 0.0;
                for (i64 k = 0; k < K; ++k) s += w[k] * h[k];
                e[n] += s;
            }
        }
    }
}

/* src/common.jl:71-81 tensor_transconv!: out = 0; for lag: out[:, 1:T-lag] += W[:,:,lag+1] * X[:, 1+lag:T] */
void orc_tensor_transconv(const double *W, const double *X, double *out, i64 N, i64 T, i64 K, i64 L) {
    memset(out, 0, sizeof(double) * (size_t)(K * T));
    for (i64 lag = 0; lag < L && lag < T; ++lag) {
        const double *Wl = W + K * N * lag;
#pragma omp parallel for schedule(static)
        for (i64 t = 0; t < T - lag; ++t) {
            const double *x = X + N * (t + lag);
            double *o = out + K * t;
            for (i64 n = 0; n < N; ++n) {
                const double xv = x[n];
                const double *w = Wl + K * n;
                for (i64 k = 0; k < K; ++k) o[k] += w[k] * xv;
            }
        }
    }
}

/* src/algs/mult.jl:31-34: num[:, :, lag+1] = H[:, 1:T-lag] * X[:, 1+lag:T]' */
void orc_corr_w(const double *H, const double *X, double *out, i64 N, i64 T, i64 K, i64 L) {
    memset(out, 0, sizeof(double) * (size_t)(K * N * L));
#pragma omp parallel for schedule(dynamic)
    for (i64 lag = 0; lag < L; ++lag) {
        double *o = out + K * N * lag;
        for (i64 t = 0; t < T - lag; ++t) {
            const double *h = H + K * t;
            const double *x = X + N * (t + lag);
            for (i64 n = 0; n < N; ++n) {
                const double xv = x[n];
                double *on = o + K * n;
                for (i64 k = 0; k < K; ++k) on[k] += h[k] * xv;
            }
        }
    }
}

static double frob(const double *a, i64 n) {
    double s = 0.0;
#pragma omp parallel for reduction(+ : s)
    for (i64 i = 0; i < n; ++i) s += a[i] * a[i];
    return sqrt(s);
}

double orc_norm(const double *a, i64 n) { return frob(a, n); }

/* src/common.jl:54-59 */
double orc_compute_loss(const double *X, const double *W, const double *H, double *scratch_NT,
                        i64 N, i64 T, i64 K, i64 L) {
    orc_tensor_conv(W, H, scratch_NT, N, T, K, L);
    for (i64 i = 0; i < N * T; ++i) scratch_NT[i] -= X[i];
    return frob(scratch_NT, N * T) / frob(X, N * T);
}

/* src/algs/mult.jl:23-39.  est: N*T scratch; num, den: K*N*L scratch. */
void orc_mu_update_motifs(const double *X, double *W, const double *H, double *est, double *num,
                          double *den, i64 N, i64 T, i64 K, i64 L, double l1W, double l2W) {
    orc_tensor_conv(W, H, est, N, T, K, L);
    orc_corr_w(H, X, num, N, T, K, L);
    orc_corr_w(H, est, den, N, T, K, L);
    for (i64 i = 0; i < K * N * L; ++i) {
        double w = W[i];
        w *= num[i] / (den[i] + l1W + 2 * l2W * w + ORC_EPS); /* :37 */
        W[i] = w > ORC_EPS ? w : ORC_EPS;                     /* :38 */
    }
}

/* src/algs/mult.jl:42-58.  est: N*T scratch; num, den: K*T scratch.  Returns relative loss. */
double orc_mu_update_feature_maps(const double *X, const double *W, double *H, double *est,
                                  double *num, double *den, i64 N, i64 T, i64 K, i64 L,
                                  double l1H, double l2H, double data_norm) {
    orc_tensor_conv(W, H, est, N, T, K, L);
    orc_tensor_transconv(W, X, num, N, T, K, L);
    orc_tensor_transconv(W, est, den, N, T, K, L);
    for (i64 i = 0; i < K * T; ++i) {
        double h = H[i];
        h *= num[i] / (den[i] + l1H + 2 * l2H * h + ORC_EPS); /* :51 */
        H[i] = h > ORC_EPS ? h : ORC_EPS;                     /* :52 */
    }
    orc_tensor_conv(W, H, est, N, T, K, L); /* :55 */
    for (i64 i = 0; i < N * T; ++i) est[i] -= X[i];
    return frob(est, N * T) / data_norm; /* :57 */
}

/* src/algs/hals.jl:18-28: resids = tensor_conv(W, H) - data */
void orc_hals_init(const double *X, const double *W, const double *H, double *R, i64 N, i64 T,
                   i64 K, i64 L) {
    orc_tensor_conv(W, H, R, N, T, K, L);
    for (i64 i = 0; i < N * T; ++i) R[i] -= X[i];
}

/* src/algs/hals.jl:31-34,53-61,90-112.  The unfolded row ind = l*K + k of H-tilde is
 * H[k, t-l] for t >= l, 0 before; it is addressed in place instead of being materialised
 * (same values).  H_norms[ind] = || Htilde[ind, :] ||. */
void orc_hals_update_motifs(double *R, double *W, const double *H, i64 N, i64 T, i64 K, i64 L,
                            double l1W, double l2W) {
    double *proj = (double *)malloc(sizeof(double) * (size_t)N);
    for (i64 k = 0; k < K; ++k) {
        for (i64 l = 0; l < L; ++l) {
            double *w = W + k + K * N * l; /* stride K over n */
            double hn2 = 0.0;
            for (i64 t = l; t < T; ++t) hn2 += H[k + K * (t - l)] * H[k + K * (t - l)];
            const double hnorm = sqrt(hn2); /* :59 norm(), squared again at :111 */
            /* :104 resids .-= W[k,:,l+1] * Htilde[ind,:]' ; :111 -resids * Hkl */
            for (i64 n = 0; n < N; ++n) proj[n] = 0.0;
#pragma omp parallel for schedule(static)
            for (i64 n = 0; n < N; ++n) {
                const double wn = w[K * n];
                double s = 0.0;
                for (i64 t = l; t < T; ++t) {
                    const double h = H[k + K * (t - l)];
                    double r = R[n + N * t] - wn * h;
                    R[n + N * t] = r;
                    s += -r * h;
                }
                proj[n] = s;
            }
            const double den = hnorm * hnorm + ORC_EPS + l2W;
#pragma omp parallel for schedule(static)
            for (i64 n = 0; n < N; ++n) {
                double v = (proj[n] - l1W) / den;
                v = v > 0.0 ? v : 0.0;
                w[K * n] = v;
                for (i64 t = l; t < T; ++t) R[n + N * t] += v * H[k + K * (t - l)]; /* :106 */
            }
        }
    }
    free(proj);
}

/* src/algs/hals.jl:37-42,64-80,121-154.  Returns ||R|| / data_norm. */
double orc_hals_update_feature_maps(double *R, const double *W, double *H, i64 N, i64 T, i64 K,
                                    i64 L, double l1H, double l2H, double data_norm) {
    double *Wn = (double *)malloc(sizeof(double) * (size_t)(K * L)); /* W_norms[k,l] :68-73 */
    for (i64 k = 0; k < K; ++k)
        for (i64 l = 0; l < L; ++l) {
            double s = 0.0;
            for (i64 n = 0; n < N; ++n) {
                const double v = W[k + K * (n + N * l)];
                s += v * v;
            }
            Wn[k + K * l] = sqrt(s);
        }
    for (i64 k = 0; k < K; ++k) {
        for (i64 t = 0; t < T; ++t) {
            const i64 w = imin(T - t, L);
            double n2 = 0.0; /* :137 norm(W_norms[k, 1:w]) then squared at :153 */
            for (i64 l = 0; l < w; ++l) n2 += Wn[k + K * l] * Wn[k + K * l];
            const double nrm = sqrt(n2);
            const double hkt = H[k + K * t];
            double trace = 0.0;
            for (i64 l = 0; l < w; ++l) {
                double *r = R + N * (t + l);
                const double *wk = W + k + K * N * l;
                for (i64 n = 0; n < N; ++n) {
                    const double wv = wk[K * n];
                    const double rem = r[n] - hkt * wv; /* :141 */
                    r[n] = rem;
                    trace += wv * (-rem); /* :152 */
                }
            }
            double hv = (trace - l1H) / (nrm * nrm + ORC_EPS + l2H); /* :153 */
            hv = hv > 0.0 ? hv : 0.0;
            H[k + K * t] = hv;
            for (i64 l = 0; l < w; ++l) { /* :147 */
                double *r = R + N * (t + l);
                const double *wk = W + k + K * N * l;
                for (i64 n = 0; n < N; ++n) r[n] += hv * wk[K * n];
            }
        }
    }
    free(Wn);
    return frob(R, N * T) / data_norm; /* :41 */
}

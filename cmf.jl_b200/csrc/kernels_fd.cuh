// Frequency-domain engine: the two data-sized contractions of the fit path,
//   numH[k,t]   = sum_l sum_n W[k,n,l] X[n,t+l]      (src/common.jl:71-81, called at src/algs/mult.jl:47)
//   numW[k,n,l] = sum_t H[k,t] X[n,t+l]              (src/algs/mult.jl:31-34)
// are correlations along t, so in overlap-save blocks of length B (hop V = B-L+1) they become, per frequency f, one
// small complex matrix product over the spectrum of X:
//   numH^[k,b,f] = sum_n conj(W^[k,n,f]) X^[n,b,f]          numW^[k,n,f] = sum_b conj(Hz^[k,b,f]) X^[n,b,f]
// (Hz = the V owned columns of block b, zero padded to B).  That cuts 2NKLT flops to ~8NKT(B/V) and makes the path
// HBM-bound on the spectrum of X, which is constant over the fit and computed once (the circular-convolution idea of
// src/common.jl:36-50, made exact by overlap-save).  The complex products run as real GEMMs on the tcgen05 kernel
// (kernels_tc.cuh, modes TC_FQT / TC_FQC) with split-bf16 operands; this file holds the SIMT FFT kernels around them.
//
// Layouts (bf16 hi/lo planes unless noted; c = 0 real part, 1 imaginary part; rows m = k real, m = Kq + k imaginary, Kq = 64 or 128):
//   Xf [f][b][c][n]          B operand of both products (K-major for TC_FQT, MN-major for TC_FQC)
//   Aw [f][m][c][n]          conj(W^) as the real 128 x 2N matrix [[Wr, Wi], [-Wi, Wr]]          (A of TC_FQT)
//   Ah [f][b][c][m]          conj(Hz^) as the real 2nblk x 128 matrix, rows (b,c): [Hr | -Hi], [Hi | Hr]  (A of TC_FQC)
//   Of [f][b][m]   fp32      numH^ (output of TC_FQT)
//   Df [f][m][n]   fp32      numW^ (output of TC_FQC)
// The Gram cross-correlation of H (Rg[d][k][k'] = sum_u H[k,u] H[k',u+d], the Toeplitz part of Htilde Htilde', mult.jl:28,33)
// is the same product with the spectrum of H itself as the B operand:
//   Hf [f][b][c][k']         full blocks of H (owned columns + right halo), 64 columns per row    (B of TC_FQC)
//   Gf [f][m][k']  fp32      Rg^ (output of TC_FQC)
// and denomH = C (*) H (mult.jl:44,48 in Gram form: 2L-1 lags of the K x K table C over H with both halos) is the numH
// product with N -> K: blocks of hop V2 = B-2L+2 starting L-1 columns early,
//   Ac [f][m][c][k']         conj(C^) as the real 128 x 128 matrix                                 (A of TC_FQT)
//   Hf [f][b][c][k']         full blocks of H (same layout as above, other blocking)               (B of TC_FQT)
// The direct loss ||conv(W,H) - X||^2 (mult.jl:55-57) goes the other way: Xhat^[n,b,f] = sum_k W^[k,n,f] H^[k,b,f] over full
// blocks of H that start L-1 columns early (hop V), inverse transform, the last V samples of a block are exact:
//   Awm[f][co][c][k][n]      W^ as the real matrices [Wr | -Wi] (co = re) and [Wi | Wr] (co = im), n contiguous (A of TC_FQX)
//   Yf [f][b][co][n] fp32    Xhat^ for a chunk of blocks (output of TC_FQX) -> ifft_resid_kernel subtracts X and sums squares
//
// FFT: in-place radix-2 decimation-in-frequency (two stages fused per pass) in shared memory over a tile d[B][C] of C
// independent complex columns (column index fastest: conflict-free), output in bit-reversed order.  Two real sequences ride in one complex transform
// (z = a + i b; A[f] = (Z[f] + conj(Z[B-f]))/2, B[f] = (Z[f] - conj(Z[B-f]))/(2i)), forward and inverse.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cmf {
namespace fd {

constexpr int NT = 512;      // threads per CTA (A/B at c4: 256 -> 49, 512 -> 44, 1024 -> 50 ms per iteration)
constexpr int KQ_MAX = 128;  // components are padded to Kq = 64 or 128 rows; the A operands / outputs have MROWS = 2 Kq rows per
                             // frequency (m = k real part, m = Kq + k imaginary part), i.e. one or two 128-row tensor-core tiles

__device__ __forceinline__ void make_twiddles(float2 *tw, int B) {
    for (int m = threadIdx.x; m < B / 2; m += NT) {
        float s, c;
        sincospif(-2.0f * (float)m / (float)B, &s, &c);   // exp(-2 pi i m / B)
        tw[m] = make_float2(c, s);
    }
}

__device__ __forceinline__ int rev(int x, int logB) { return (int)(__brev((unsigned)x) >> (32 - logB)); }

// log2(B) radix-2 decimation-in-frequency stages over d[B][C], two stages fused per pass (four rows in registers, the
// second twiddle of the first stage is the first times -+i): half the shared-memory traffic and barriers of plain
// radix-2, same bit-reversed output order (X[f] ends up in row rev(f)).  INV uses the conjugate twiddles (no scaling).
__device__ __forceinline__ float2 cmul(float2 a, float2 w) { return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }

template <bool INV>
__device__ __forceinline__ void fft_passes(float2 *d, const float2 *tw, int B, int logB, int C) {
    const int c = threadIdx.x % C, j0 = threadIdx.x / C, jstep = NT / C;
    int s = 0;
    for (; s + 1 < logB; s += 2) {
        const int n = B >> s, q = n >> 2, lq = logB - s - 2;      // block size of the first stage, its quarter
        for (int j = j0; j < B / 4; j += jstep) {
            const int pos = j & (q - 1), g = j >> lq;
            float2 *r0 = d + (size_t)(g * n + pos) * C + c, *r1 = r0 + (size_t)q * C, *r2 = r1 + (size_t)q * C, *r3 = r2 + (size_t)q * C;
            const float2 a0 = *r0, a1 = *r1, a2 = *r2, a3 = *r3;
            float2 w1 = tw[pos << s], w2 = tw[pos << (s + 1)];
            if (INV) { w1.y = -w1.y; w2.y = -w2.y; }
            const float2 b0 = make_float2(a0.x + a2.x, a0.y + a2.y), b2 = cmul(make_float2(a0.x - a2.x, a0.y - a2.y), w1);
            const float2 b1 = make_float2(a1.x + a3.x, a1.y + a3.y), uw = cmul(make_float2(a1.x - a3.x, a1.y - a3.y), w1);
            const float2 b3 = INV ? make_float2(-uw.y, uw.x) : make_float2(uw.y, -uw.x);
            *r0 = make_float2(b0.x + b1.x, b0.y + b1.y);
            *r1 = cmul(make_float2(b0.x - b1.x, b0.y - b1.y), w2);
            *r2 = make_float2(b2.x + b3.x, b2.y + b3.y);
            *r3 = cmul(make_float2(b2.x - b3.x, b2.y - b3.y), w2);
        }
        __syncthreads();
    }
    if (s < logB) {                                               // odd log2(B): last stage, distance 1, twiddle 1
        for (int j = j0; j < B / 2; j += jstep) {
            float2 *r0 = d + (size_t)(2 * j) * C + c, *r1 = r0 + C;
            const float2 a = *r0, b = *r1;
            *r0 = make_float2(a.x + b.x, a.y + b.y);
            *r1 = make_float2(a.x - b.x, a.y - b.y);
        }
        __syncthreads();
    }
}

// spectra of the two real sequences packed in one complex transform: A = (ar, ai), B = (br, bi) at frequency f
__device__ __forceinline__ void unpack_pair(const float2 *d, int f, int B, int logB, int C, int p, float &ar, float &ai, float &br,
                                            float &bi) {
    const float2 z = d[rev(f, logB) * C + p], y = d[rev((B - f) & (B - 1), logB) * C + p];
    ar = 0.5f * (z.x + y.x); ai = 0.5f * (z.y - y.y);
    br = 0.5f * (z.y + y.y); bi = -0.5f * (z.x - y.x);
}

// the inverse of unpack_pair: rows f and B-f of Z = A + iB from the half spectra A = (re.x, im.x), B = (re.y, im.y)
__device__ __forceinline__ void pack_pair(float2 *d, int f, int B, int C, int p, float2 re, float2 im) {
    if (f == 0 || f == B / 2) im = make_float2(0.f, 0.f);          // real sequences: these bins are real
    d[f * C + p] = make_float2(re.x - im.y, im.x + re.y);
    if (f > 0 && f < B / 2) d[(B - f) * C + p] = make_float2(re.x + im.y, re.y - im.x);
}

__device__ __forceinline__ void store_split2(__nv_bfloat16 *hi, __nv_bfloat16 *lo, int64_t idx, float a, float b) {
    const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
    const __nv_bfloat16 al = __float2bfloat16_rn(a - __bfloat162float(ah)), bl = __float2bfloat16_rn(b - __bfloat162float(bh));
    __nv_bfloat162 h, l;
    h.x = ah; h.y = bh; l.x = al; l.y = bl;
    *reinterpret_cast<__nv_bfloat162 *>(hi + idx) = h;
    *reinterpret_cast<__nv_bfloat162 *>(lo + idx) = l;
}

// X[t][N] fp32 (xcols rows) -> Xf.  grid (nblkp, ceil(N/32)); 16 complex columns = 32 units per CTA.
__global__ void __launch_bounds__(NT)
fft_x_kernel(const float *__restrict__ X, __nv_bfloat16 *__restrict__ hi, __nv_bfloat16 *__restrict__ lo, int64_t N, int64_t xcols,
             int B, int logB, int V, int64_t nblkp) {
    extern __shared__ float2 fd_smem[];
    constexpr int C = 16;
    float2 *d = fd_smem, *tw = fd_smem + (size_t)B * C;
    const int64_t b = blockIdx.x;
    const int p = threadIdx.x % C;
    const int64_t n = (int64_t)blockIdx.y * 32 + 2 * p;
    make_twiddles(tw, B);
    for (int i = threadIdx.x / C; i < B; i += NT / C) {
        const int64_t t = b * V + i;
        float2 v = make_float2(0.f, 0.f);
        if (t < xcols && n < N) v = *reinterpret_cast<const float2 *>(X + t * N + n);
        d[i * C + p] = v;
    }
    __syncthreads();
    fft_passes<false>(d, tw, B, logB, C);
    if (n >= N) return;
    for (int f = threadIdx.x / C; f <= B / 2; f += NT / C) {
        float ar, ai, br, bi;
        unpack_pair(d, f, B, logB, C, p, ar, ai, br, bi);
        const int64_t r = ((int64_t)f * nblkp + b) * 2;
        store_split2(hi, lo, r * N + n, ar, br);
        store_split2(hi, lo, (r + 1) * N + n, ai, bi);
    }
}

// H[t][K] fp32 (owned column 0 first; block b starts at column b*V + t_off, t_off <= 0 reaches into the left halo).
// full == 0: only the V owned columns of each block, zero padded -> Ah;
// full != 0: whole blocks over the columns present, t < hcols = Tl + L-1 (owned + right halo) -> Hf.
// grid (nblkp, 32 / C); C complex columns = 2C components per CTA.
__global__ void __launch_bounds__(NT)
fft_h_kernel(const float *__restrict__ H, __nv_bfloat16 *__restrict__ hi, __nv_bfloat16 *__restrict__ lo, int64_t K, int64_t Tl,
             int64_t hcols, int B, int logB, int V, int64_t nblkp, int C, int full, int64_t t_off, int kq) {
    const int64_t KQ = kq, MROWS = 2 * kq;
    extern __shared__ float2 fd_smem[];
    float2 *d = fd_smem, *tw = fd_smem + (size_t)B * C;
    const int64_t b = blockIdx.x;
    const int p = threadIdx.x % C;
    const int k = 2 * ((int)blockIdx.y * C + p);
    make_twiddles(tw, B);
    for (int i = threadIdx.x / C; i < B; i += NT / C) {
        const int64_t t = b * V + t_off + i;
        float2 v = make_float2(0.f, 0.f);
        if (full ? (t < hcols) : (i < V && t < Tl)) {
            if (k < K) v.x = H[t * K + k];
            if (k + 1 < K) v.y = H[t * K + k + 1];
        }
        d[i * C + p] = v;
    }
    __syncthreads();
    fft_passes<false>(d, tw, B, logB, C);
    for (int f = threadIdx.x / C; f <= B / 2; f += NT / C) {
        float ar, ai, br, bi;
        unpack_pair(d, f, B, logB, C, p, ar, ai, br, bi);
        if (full) {
            const int64_t r = (((int64_t)f * nblkp + b) * 2) * KQ;
            store_split2(hi, lo, r + k, ar, br);
            store_split2(hi, lo, r + KQ + k, ai, bi);
            continue;
        }
        const int64_t r = (((int64_t)f * nblkp + b) * 2) * MROWS;
        store_split2(hi, lo, r + k, ar, br);                       // row (b, re): [ Hr | -Hi ]
        store_split2(hi, lo, r + KQ + k, -ai, -bi);
        store_split2(hi, lo, r + MROWS + k, ai, bi);               // row (b, im): [ Hi |  Hr ]
        store_split2(hi, lo, r + MROWS + KQ + k, ar, br);
    }
}

// Wi[(l*K+k)][N] fp32 -> Aw (rows of ldw elements, imaginary block at column coff: 2N / N for W; 128 / 64 for the
// K x K lag table C with N = K and L = 2L-1 lags -> Ac).  grid (ceil(N/32), K).
__global__ void __launch_bounds__(NT)
fft_w_kernel(const float *__restrict__ Wi, __nv_bfloat16 *__restrict__ hi, __nv_bfloat16 *__restrict__ lo, int64_t N, int64_t K,
             int64_t L, int B, int logB, int64_t ldw, int64_t coff, int xmode, int kq) {
    const int64_t KQ = kq, MROWS = 2 * kq;
    extern __shared__ float2 fd_smem[];
    constexpr int C = 16;
    float2 *d = fd_smem, *tw = fd_smem + (size_t)B * C;
    const int64_t k = blockIdx.y;
    const int p = threadIdx.x % C;
    const int64_t n = (int64_t)blockIdx.x * 32 + 2 * p;
    make_twiddles(tw, B);
    for (int i = threadIdx.x / C; i < B; i += NT / C) {
        float2 v = make_float2(0.f, 0.f);
        if (i < L && n < N) {
            const float *src = Wi + ((int64_t)i * K + k) * N + n;
            if ((N & 1) == 0) v = *reinterpret_cast<const float2 *>(src);
            else { v.x = src[0]; if (n + 1 < N) v.y = src[1]; }
        }
        d[i * C + p] = v;
    }
    __syncthreads();
    fft_passes<false>(d, tw, B, logB, C);
    if (n >= N) return;
    for (int f = threadIdx.x / C; f <= B / 2; f += NT / C) {
        float ar, ai, br, bi;
        unpack_pair(d, f, B, logB, C, p, ar, ai, br, bi);
        if (xmode) {                                               // Awm[f][co][c][k][n]: re = [Wr | -Wi], im = [Wi | Wr]
            const int64_t r0 = (((int64_t)f * 2 + 0) * MROWS + k) * N + n, r1 = (((int64_t)f * 2 + 1) * MROWS + k) * N + n;
            store_split2(hi, lo, r0, ar, br);
            store_split2(hi, lo, r0 + KQ * N, -ai, -bi);
            store_split2(hi, lo, r1, ai, bi);
            store_split2(hi, lo, r1 + KQ * N, ar, br);
            continue;
        }
        const int64_t re = ((int64_t)f * MROWS + k) * ldw, im = ((int64_t)f * MROWS + KQ + k) * ldw;
        store_split2(hi, lo, re + n, ar, br);                      // row k:      [  Wr | Wi ]
        store_split2(hi, lo, re + coff + n, ai, bi);
        store_split2(hi, lo, im + n, -ai, -bi);                    // row 64 + k: [ -Wi | Wr ]
        store_split2(hi, lo, im + coff + n, ar, br);
    }
}

// Of[f][b][m] fp32 -> numH[t][K] (owned columns; V valid outputs per block).  grid (nblk, 32 / C).
__global__ void __launch_bounds__(NT)
ifft_numH_kernel(const float *__restrict__ Of, float *__restrict__ numH, int64_t K, int64_t Tl, int B, int logB, int V,
                 int64_t nblkp, int C, int kq) {
    const int64_t KQ = kq, MROWS = 2 * kq;
    extern __shared__ float2 fd_smem[];
    float2 *d = fd_smem, *tw = fd_smem + (size_t)B * C;
    const int64_t b = blockIdx.x;
    const int p = threadIdx.x % C;
    const int k = 2 * ((int)blockIdx.y * C + p);
    make_twiddles(tw, B);
    for (int f = threadIdx.x / C; f <= B / 2; f += NT / C) {
        const float *o = Of + ((int64_t)f * nblkp + b) * MROWS;
        pack_pair(d, f, B, C, p, *reinterpret_cast<const float2 *>(o + k), *reinterpret_cast<const float2 *>(o + KQ + k));
    }
    __syncthreads();
    fft_passes<true>(d, tw, B, logB, C);
    const float sc = 1.0f / (float)B;
    for (int i = threadIdx.x / C; i < V; i += NT / C) {
        const int64_t t = b * V + i;
        if (t >= Tl) break;
        const float2 z = d[rev(i, logB) * C + p];
        if (k < K) numH[t * K + k] = z.x * sc;
        if (k + 1 < K) numH[t * K + k + 1] = z.y * sc;
    }
}

// Yf[f][bc][co][n] fp32 (a chunk of nbc blocks starting at block b0) -> partial[bc * gridDim.y + blockIdx.y] =
// sum over the V exact samples of the block and the CTA's 32 units of (Xhat - X)^2.  Block b covers the columns
// [b*V - (L-1), b*V - (L-1) + B) of H, so sample i >= L-1 of the inverse transform is Xhat at t = b*V + i - (L-1).
// grid (nbc, ceil(N/32)).
__global__ void __launch_bounds__(NT)
ifft_resid_kernel(const float *__restrict__ Yf, const float *__restrict__ X, double *__restrict__ partial, int64_t N, int64_t Tl,
                  int64_t L, int B, int logB, int V, int64_t nbc, int64_t b0) {
    extern __shared__ float2 fd_smem[];
    __shared__ double red[NT / 32];
    constexpr int C = 16;
    float2 *d = fd_smem, *tw = fd_smem + (size_t)B * C;
    const int64_t bc = blockIdx.x;
    const int p = threadIdx.x % C;
    const int64_t n = (int64_t)blockIdx.y * 32 + 2 * p;
    make_twiddles(tw, B);
    for (int f = threadIdx.x / C; f <= B / 2; f += NT / C) {
        float2 re = make_float2(0.f, 0.f), im = re;
        if (n < N) {
            const float *y = Yf + (((int64_t)f * nbc + bc) * 2) * N + n;
            re = *reinterpret_cast<const float2 *>(y);
            im = *reinterpret_cast<const float2 *>(y + N);
        }
        pack_pair(d, f, B, C, p, re, im);
    }
    __syncthreads();
    fft_passes<true>(d, tw, B, logB, C);
    const float sc = 1.0f / (float)B;
    float acc = 0.f;
    double accd = 0.0;
    if (n < N) {
        int cnt = 0;
        for (int i = (int)(L - 1) + threadIdx.x / C; i < B; i += NT / C) {
            const int64_t t = (b0 + bc) * V + (i - (L - 1));
            if (t >= Tl) break;
            const float2 z = d[rev(i, logB) * C + p];
            const float2 x = *reinterpret_cast<const float2 *>(X + t * N + n);
            const float r0 = z.x * sc - x.x, r1 = z.y * sc - x.y;
            acc = fmaf(r0, r0, acc);
            acc = fmaf(r1, r1, acc);
            if (++cnt == 8) { accd += (double)acc; acc = 0.f; cnt = 0; }
        }
    }
    accd += (double)acc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) accd += __shfl_xor_sync(0xffffffffu, accd, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = accd;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < NT / 32; ++w) s += red[w];
        partial[bc * gridDim.y + blockIdx.y] = s;
    }
}

// Df[f][m][n] fp32 (row stride ldi) -> out[(l*K+k)*N + n], l < L.  grid (ceil(N/32), K).
// OutT = float: numW (N even, paired stores);  OutT = double: the Gram partial Rg[d][k][k'] with N = K.
template <typename OutT>
__global__ void __launch_bounds__(NT)
ifft_numW_kernel(const float *__restrict__ Df, OutT *__restrict__ out, int64_t N, int64_t ldi, int64_t K, int64_t L, int B, int logB,
                 int kq) {
    const int64_t KQ = kq, MROWS = 2 * kq;
    extern __shared__ float2 fd_smem[];
    constexpr int C = 16;
    float2 *d = fd_smem, *tw = fd_smem + (size_t)B * C;
    const int64_t k = blockIdx.y;
    const int p = threadIdx.x % C;
    const int64_t n = (int64_t)blockIdx.x * 32 + 2 * p;
    make_twiddles(tw, B);
    for (int f = threadIdx.x / C; f <= B / 2; f += NT / C) {
        float2 re = make_float2(0.f, 0.f), im = re;
        if (n < N) {
            re = *reinterpret_cast<const float2 *>(Df + ((int64_t)f * MROWS + k) * ldi + n);
            im = *reinterpret_cast<const float2 *>(Df + ((int64_t)f * MROWS + KQ + k) * ldi + n);
        }
        pack_pair(d, f, B, C, p, re, im);
    }
    __syncthreads();
    fft_passes<true>(d, tw, B, logB, C);
    if (n >= N) return;
    const float sc = 1.0f / (float)B;
    for (int l = threadIdx.x / C; l < L; l += NT / C) {
        const float2 z = d[rev(l, logB) * C + p];
        OutT *o = out + ((int64_t)l * K + k) * N + n;
        if (sizeof(OutT) == 4) {
            *reinterpret_cast<float2 *>(o) = make_float2(z.x * sc, z.y * sc);
        } else {
            o[0] = (OutT)(z.x * sc);
            if (n + 1 < N) o[1] = (OutT)(z.y * sc);
        }
    }
}

}  // namespace fd
}  // namespace cmf

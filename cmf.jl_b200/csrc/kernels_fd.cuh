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

#ifndef CMF_FD_MINB
#define CMF_FD_MINB 3        // CTAs per SM the transforms are compiled for (3 x 512 threads: 40 registers per thread)
#endif
constexpr int LD_U = 8;      // global loads a thread keeps in flight while it fills / drains the shared-memory tile
constexpr int NT = 512;      // threads per CTA (A/B at c4: 256 -> 49, 512 -> 44, 1024 -> 50 ms per iteration)
constexpr int KQ_MAX = 128;  // components are padded to Kq = 64 or 128 rows; the A operands / outputs have MROWS = 2 Kq rows per
                             // frequency (m = k real part, m = Kq + k imaginary part), i.e. one or two 128-row tensor-core tiles

// full-circle table tw[m] = exp(-2 pi i m / B), m < B (the radix-8 passes index it up to 7B/8)
__device__ __forceinline__ void make_twiddles(float2 *tw, int B) {
    for (int m = threadIdx.x; m < B; m += NT) {
        float s, c;
        sincospif(-2.0f * (float)m / (float)B, &s, &c);
        tw[m] = make_float2(c, s);
    }
}

__device__ __forceinline__ int rev(int x, int logB) { return (int)(__brev((unsigned)x) >> (32 - logB)); }

__device__ __forceinline__ float2 cmul(float2 a, float2 w) { return make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiplication by -i (forward) / +i (inverse), and by the eighth roots w8 = exp(-+ i pi / 4), w8^3
template <bool INV> __device__ __forceinline__ float2 mul_mi(float2 a) { return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }
template <bool INV> __device__ __forceinline__ float2 mul_w8(float2 a) {
    const float r = 0.70710678118654752f;
    return INV ? make_float2((a.x - a.y) * r, (a.x + a.y) * r) : make_float2((a.x + a.y) * r, (a.y - a.x) * r);
}
template <bool INV> __device__ __forceinline__ float2 mul_w83(float2 a) {
    const float r = 0.70710678118654752f;
    return INV ? make_float2((-a.x - a.y) * r, (a.x - a.y) * r) : make_float2((a.y - a.x) * r, (-a.x - a.y) * r);
}

// log2(B) radix-2 decimation-in-frequency stages over d[B][C], in place, output in bit-reversed order (X[f] ends up in row
// rev(f)), THREE stages fused per pass: a thread holds the 8 rows {pos + j n/8} of one block of size n in registers, runs the
// radix-8 butterfly with the constant eighth roots, and applies the position twiddles on the way out (row m of the butterfly
// carries w_n^(pos * bitrev3(m)): 7 table reads per 8 points).  B = 512 is three such passes; what is left of log2(B) after
// the radix-8 passes (blocks of 4 or 2 rows, all twiddles 1) is one radix-4 or radix-2 pass.  INV conjugates every root.
template <bool INV>
__device__ __forceinline__ void fft_passes(float2 *d, const float2 *tw, int B, int logB, int C) {
    const int c = threadIdx.x % C, j0 = threadIdx.x / C, jstep = NT / C;
    int s = 0;
    for (; s + 2 < logB; s += 3) {
        const int lq = logB - s - 3, q = 1 << lq, qC = q * C;             // eighth of the block size n = B >> s
        for (int j = j0; j < (B >> 3); j += jstep) {
            const int pos = j & (q - 1), g = j >> lq;
            float2 *r = d + ((g << (lq + 3)) + pos) * C + c;
            float2 a0 = r[0], a1 = r[qC], a2 = r[2 * qC], a3 = r[3 * qC], a4 = r[4 * qC], a5 = r[5 * qC], a6 = r[6 * qC], a7 = r[7 * qC];
            // stage 1: rows j and j+4, constant root w8^j on the difference
            float2 t;
            t = csub(a0, a4); a0 = cadd(a0, a4); a4 = t;
            t = csub(a1, a5); a1 = cadd(a1, a5); a5 = mul_w8<INV>(t);
            t = csub(a2, a6); a2 = cadd(a2, a6); a6 = mul_mi<INV>(t);
            t = csub(a3, a7); a3 = cadd(a3, a7); a7 = mul_w83<INV>(t);
            // stage 2: rows j and j+2 inside each half, constant root w4^j
            t = csub(a0, a2); a0 = cadd(a0, a2); a2 = t;
            t = csub(a1, a3); a1 = cadd(a1, a3); a3 = mul_mi<INV>(t);
            t = csub(a4, a6); a4 = cadd(a4, a6); a6 = t;
            t = csub(a5, a7); a5 = cadd(a5, a7); a7 = mul_mi<INV>(t);
            // stage 3: neighbouring rows
            t = csub(a0, a1); a0 = cadd(a0, a1); a1 = t;
            t = csub(a2, a3); a2 = cadd(a2, a3); a3 = t;
            t = csub(a4, a5); a4 = cadd(a4, a5); a5 = t;
            t = csub(a6, a7); a6 = cadd(a6, a7); a7 = t;
            if (lq > 0) {                                                  // position twiddles (all 1 in the last radix-8 pass of 8^m)
                const int e = pos << s;
                float2 w;
#define CMF_TW(m_, a_) w = tw[(m_) * e]; if (INV) w.y = -w.y; a_ = cmul(a_, w);
                CMF_TW(4, a1) CMF_TW(2, a2) CMF_TW(6, a3) CMF_TW(1, a4) CMF_TW(5, a5) CMF_TW(3, a6) CMF_TW(7, a7)
#undef CMF_TW
            }
            r[0] = a0; r[qC] = a1; r[2 * qC] = a2; r[3 * qC] = a3; r[4 * qC] = a4; r[5 * qC] = a5; r[6 * qC] = a6; r[7 * qC] = a7;
        }
        __syncthreads();
    }
    if (logB - s == 2) {                                                  // blocks of 4 rows: pos = 0, twiddles 1
        for (int j = j0; j < (B >> 2); j += jstep) {
            float2 *r = d + (4 * j) * C + c;
            float2 a0 = r[0], a1 = r[C], a2 = r[2 * C], a3 = r[3 * C], t;
            t = csub(a0, a2); a0 = cadd(a0, a2); a2 = t;
            t = csub(a1, a3); a1 = cadd(a1, a3); a3 = mul_mi<INV>(t);
            r[0] = cadd(a0, a1); r[C] = csub(a0, a1); r[2 * C] = cadd(a2, a3); r[3 * C] = csub(a2, a3);
        }
        __syncthreads();
    } else if (logB - s == 1) {                                           // blocks of 2 rows
        for (int j = j0; j < (B >> 1); j += jstep) {
            float2 *r0 = d + (2 * j) * C + c, *r1 = r0 + C;
            const float2 a = *r0, b = *r1;
            *r0 = cadd(a, b);
            *r1 = csub(a, b);
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

// (a, b) -> packed bf16 pairs hi = bf16(x), lo = bf16(x - hi) (round to nearest, as tc::split_bf16)
__device__ __forceinline__ void split2(float a, float b, uint32_t &h, uint32_t &l) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(a, b);
    const float2 hf = __bfloat1622float2(hh);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(a - hf.x, b - hf.y);
    h = *reinterpret_cast<const uint32_t *>(&hh);
    l = *reinterpret_cast<const uint32_t *>(&ll);
}
__device__ __forceinline__ void store_split2(__nv_bfloat16 *hi, __nv_bfloat16 *lo, int64_t idx, float a, float b) {
    uint32_t h, l;
    split2(a, b, h, l);
    *reinterpret_cast<uint32_t *>(hi + idx) = h;
    *reinterpret_cast<uint32_t *>(lo + idx) = l;
}

// X[t][N] fp32 (xcols rows) -> Xf.  grid nblkp * ceil(N/32) (1-D); 16 complex columns = 32 units per CTA.
__global__ void __launch_bounds__(NT, CMF_FD_MINB)
fft_x_kernel(const float *__restrict__ X, __nv_bfloat16 *__restrict__ hi, __nv_bfloat16 *__restrict__ lo, int64_t N, int64_t xcols,
             int B, int logB, int V, int64_t nblkp, int64_t nbx) {
    extern __shared__ float2 fd_smem[];
    constexpr int C = 16;
    float2 *d = fd_smem, *tw = fd_smem + (size_t)B * C;
    // 1-D grid over (block, unit tile); nbx = number of blocks when the block index runs fastest, 0 = the tile index runs fastest
    const int ny = (int)((N + 31) / 32);
    const int64_t b = nbx ? blockIdx.x % nbx : blockIdx.x / ny;
    const int p = threadIdx.x % C;
    const int64_t n = (int64_t)(nbx ? blockIdx.x / nbx : blockIdx.x % ny) * 32 + 2 * p;
    make_twiddles(tw, B);
    {
        // LD_U rows per thread are requested before any is stored: the transforms are bound by the latency of these loads, not by
        // bandwidth or issue slots (profiles/r2_ncu_fft_kernels.md), so memory-level parallelism is what counts
        const int i0 = threadIdx.x / C, istep = NT / C;
        const int64_t lim64 = xcols - b * V;
        const int lim = (n < N) ? (int)(lim64 < 0 ? 0 : (lim64 > B ? B : lim64)) : 0;
        const float *src = X + (b * V + i0) * N + n;
        for (int i = i0; i < B; i += LD_U * istep, src += (int64_t)LD_U * istep * N) {
            float2 v[LD_U];
#pragma unroll
            for (int u = 0; u < LD_U; ++u) {
                v[u] = make_float2(0.f, 0.f);
                if (i + u * istep < lim) v[u] = *reinterpret_cast<const float2 *>(src + (int64_t)u * istep * N);
            }
#pragma unroll
            for (int u = 0; u < LD_U; ++u)
                if (i + u * istep < B) d[(i + u * istep) * C + p] = v[u];
        }
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
// grid nblkp * (kq / 2 / C) (1-D); C complex columns = 2C components per CTA.
__global__ void __launch_bounds__(NT, CMF_FD_MINB)
fft_h_kernel(const float *__restrict__ H, __nv_bfloat16 *__restrict__ hi, __nv_bfloat16 *__restrict__ lo, int64_t K, int64_t Tl,
             int64_t hcols, int B, int logB, int V, int64_t nblkp, int C, int full, int64_t t_off, int kq, int64_t nbx) {
    const int64_t KQ = kq, MROWS = 2 * kq;
    extern __shared__ float2 fd_smem[];
    float2 *d = fd_smem, *tw = fd_smem + (size_t)B * C;
    const int ny = (kq / 2) / C;                      // 1-D grid over (block, component tile); nbx as in fft_x_kernel
    const int64_t b = nbx ? blockIdx.x % nbx : blockIdx.x / ny;
    const int p = threadIdx.x % C;
    const int k = 2 * ((int)(nbx ? blockIdx.x / nbx : blockIdx.x % ny) * C + p);
    make_twiddles(tw, B);
    {
        const int i0 = threadIdx.x / C, istep = NT / C;
        const int64_t tb = b * V + t_off;
        // rows of the block that hold data: [0, lim) (owned segment, or whatever of the block lies before hcols)
        const int64_t lim64 = full ? (hcols - tb) : ((Tl - tb < V) ? Tl - tb : V);
        const int lim = (int)(lim64 < 0 ? 0 : (lim64 > B ? B : lim64));
        const float *src = H + (tb + i0) * K + k;
        const bool pair = ((K & 1) == 0) && (k + 1 < K);
        for (int i = i0; i < B; i += LD_U * istep, src += (int64_t)LD_U * istep * K) {
            float2 v[LD_U];
#pragma unroll
            for (int u = 0; u < LD_U; ++u) {
                v[u] = make_float2(0.f, 0.f);
                if (i + u * istep < lim) {
                    const float *q = src + (int64_t)u * istep * K;
                    if (pair) v[u] = *reinterpret_cast<const float2 *>(q);
                    else { if (k < K) v[u].x = q[0]; if (k + 1 < K) v[u].y = q[1]; }
                }
            }
#pragma unroll
            for (int u = 0; u < LD_U; ++u)
                if (i + u * istep < B) d[(i + u * istep) * C + p] = v[u];
        }
    }
    __syncthreads();
    fft_passes<false>(d, tw, B, logB, C);
    for (int f = threadIdx.x / C; f <= B / 2; f += NT / C) {
        float ar, ai, br, bi;
        unpack_pair(d, f, B, logB, C, p, ar, ai, br, bi);
        uint32_t hr, lr, hi_, li_;
        split2(ar, br, hr, lr);
        split2(ai, bi, hi_, li_);
        if (full) {
            const int64_t r = (((int64_t)f * nblkp + b) * 2) * KQ + k;
            *reinterpret_cast<uint32_t *>(hi + r) = hr; *reinterpret_cast<uint32_t *>(lo + r) = lr;
            *reinterpret_cast<uint32_t *>(hi + r + KQ) = hi_; *reinterpret_cast<uint32_t *>(lo + r + KQ) = li_;
            continue;
        }
        const int64_t r = (((int64_t)f * nblkp + b) * 2) * MROWS + k;
        const uint32_t sg = 0x80008000u;                           // -x of a packed bf16 pair (hi and lo planes alike)
        *reinterpret_cast<uint32_t *>(hi + r) = hr; *reinterpret_cast<uint32_t *>(lo + r) = lr;                                 // row (b, re): [ Hr | -Hi ]
        *reinterpret_cast<uint32_t *>(hi + r + KQ) = hi_ ^ sg; *reinterpret_cast<uint32_t *>(lo + r + KQ) = li_ ^ sg;
        *reinterpret_cast<uint32_t *>(hi + r + MROWS) = hi_; *reinterpret_cast<uint32_t *>(lo + r + MROWS) = li_;               // row (b, im): [ Hi |  Hr ]
        *reinterpret_cast<uint32_t *>(hi + r + MROWS + KQ) = hr; *reinterpret_cast<uint32_t *>(lo + r + MROWS + KQ) = lr;
    }
}

// Wi[(l*K+k)][N] fp32 -> Aw (rows of ldw elements, imaginary block at column coff: 2N / N for W; 128 / 64 for the
// K x K lag table C with N = K and L = 2L-1 lags -> Ac).  grid (ceil(N / 2C), K): C complex columns = 2C units per CTA.
__global__ void __launch_bounds__(NT, CMF_FD_MINB)
fft_w_kernel(const float *__restrict__ Wi, __nv_bfloat16 *__restrict__ hi, __nv_bfloat16 *__restrict__ lo, int64_t N, int64_t K,
             int64_t L, int B, int logB, int64_t ldw, int64_t coff, int xmode, int kq, int C) {
    const int64_t KQ = kq, MROWS = 2 * kq;
    extern __shared__ float2 fd_smem[];
    float2 *d = fd_smem, *tw = fd_smem + (size_t)B * C;
    const int64_t k = blockIdx.y;
    const int p = threadIdx.x % C;
    const int64_t n = (int64_t)blockIdx.x * (2 * C) + 2 * p;
    make_twiddles(tw, B);
    {
        const int i0 = threadIdx.x / C, istep = NT / C;
        for (int i = i0; i < B; i += LD_U * istep) {
            float2 v[LD_U];
#pragma unroll
            for (int u = 0; u < LD_U; ++u) {
                const int ii = i + u * istep;
                v[u] = make_float2(0.f, 0.f);
                if (ii < L && n < N) {
                    const float *src = Wi + ((int64_t)ii * K + k) * N + n;
                    if ((N & 1) == 0) v[u] = *reinterpret_cast<const float2 *>(src);
                    else { v[u].x = src[0]; if (n + 1 < N) v[u].y = src[1]; }
                }
            }
#pragma unroll
            for (int u = 0; u < LD_U; ++u)
                if (i + u * istep < B) d[(i + u * istep) * C + p] = v[u];
        }
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

// Of[f][b][m] fp32 -> numH[t][K] (owned columns; V valid outputs per block).  grid nblk * (kq / 2 / C) (1-D).
__global__ void __launch_bounds__(NT, CMF_FD_MINB)
ifft_numH_kernel(const float *__restrict__ Of, float *__restrict__ numH, int64_t K, int64_t Tl, int B, int logB, int V,
                 int64_t nblkp, int C, int kq, int64_t nbx) {
    const int64_t KQ = kq, MROWS = 2 * kq;
    extern __shared__ float2 fd_smem[];
    float2 *d = fd_smem, *tw = fd_smem + (size_t)B * C;
    const int ny = (kq / 2) / C;                      // 1-D grid over (block, component tile); nbx as in fft_x_kernel
    const int64_t b = nbx ? blockIdx.x % nbx : blockIdx.x / ny;
    const int p = threadIdx.x % C;
    const int k = 2 * ((int)(nbx ? blockIdx.x / nbx : blockIdx.x % ny) * C + p);
    make_twiddles(tw, B);
    {
        const int f0 = threadIdx.x / C, fstep = NT / C;
        const float *o = Of + ((int64_t)f0 * nblkp + b) * MROWS + k;
        for (int f = f0; f <= B / 2; f += 4 * fstep, o += (int64_t)4 * fstep * nblkp * MROWS) {
            float2 re[4], im[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (f + u * fstep <= B / 2) {
                    const float *q = o + (int64_t)u * fstep * nblkp * MROWS;
                    re[u] = *reinterpret_cast<const float2 *>(q); im[u] = *reinterpret_cast<const float2 *>(q + KQ);
                }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (f + u * fstep <= B / 2) pack_pair(d, f + u * fstep, B, C, p, re[u], im[u]);
        }
    }
    __syncthreads();
    fft_passes<true>(d, tw, B, logB, C);
    const float sc = 1.0f / (float)B;
    const bool pair = ((K & 1) == 0) && (k + 1 < K);
    for (int i = threadIdx.x / C; i < V; i += NT / C) {
        const int64_t t = b * V + i;
        if (t >= Tl) break;
        const float2 z = d[rev(i, logB) * C + p];
        if (pair) *reinterpret_cast<float2 *>(numH + t * K + k) = make_float2(z.x * sc, z.y * sc);
        else { if (k < K) numH[t * K + k] = z.x * sc; if (k + 1 < K) numH[t * K + k + 1] = z.y * sc; }
    }
}

// Yf[f][bc][co][n] fp32 (a chunk of nbc blocks starting at block b0) -> partial[bc * ceil(N/32) + unit tile] =
// sum over the V exact samples of the block and the CTA's 32 units of (Xhat - X)^2.  Block b covers the columns
// [b*V - (L-1), b*V - (L-1) + B) of H, so sample i >= L-1 of the inverse transform is Xhat at t = b*V + i - (L-1).
// grid nbc * ceil(N/32) (1-D).
__global__ void __launch_bounds__(NT, CMF_FD_MINB)
ifft_resid_kernel(const float *__restrict__ Yf, const float *__restrict__ X, double *__restrict__ partial, int64_t N, int64_t Tl,
                  int64_t L, int B, int logB, int V, int64_t nbc, int64_t b0, int64_t nbx) {
    extern __shared__ float2 fd_smem[];
    __shared__ double red[NT / 32];
    constexpr int C = 16;
    float2 *d = fd_smem, *tw = fd_smem + (size_t)B * C;
    const int ny = (int)((N + 31) / 32);              // 1-D grid over (block, unit tile); nbx as in fft_x_kernel
    const int64_t bc = nbx ? blockIdx.x % nbx : blockIdx.x / ny;
    const int by = (int)(nbx ? blockIdx.x / nbx : blockIdx.x % ny);
    const int p = threadIdx.x % C;
    const int64_t n = (int64_t)by * 32 + 2 * p;
    make_twiddles(tw, B);
    {
        const int f0 = threadIdx.x / C, fstep = NT / C;
        const float *y = Yf + (((int64_t)f0 * nbc + bc) * 2) * N + n;
        for (int f = f0; f <= B / 2; f += 4 * fstep, y += (int64_t)4 * fstep * nbc * 2 * N) {
            float2 re[4], im[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                re[u] = make_float2(0.f, 0.f); im[u] = re[u];
                if (f + u * fstep <= B / 2 && n < N) {
                    const float *q = y + (int64_t)u * fstep * nbc * 2 * N;
                    re[u] = *reinterpret_cast<const float2 *>(q); im[u] = *reinterpret_cast<const float2 *>(q + N);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (f + u * fstep <= B / 2) pack_pair(d, f + u * fstep, B, C, p, re[u], im[u]);
        }
    }
    __syncthreads();
    fft_passes<true>(d, tw, B, logB, C);
    const float sc = 1.0f / (float)B;
    double accd = 0.0;
    if (n < N) {
        // rows of this thread: i = L-1 + i0 + m*istep; the X loads of 8 rows go out together, their squares are summed in fp32
        // and every group of 8 rows is added to the fp64 accumulator (same grouping as before)
        const int istep = NT / C;
        const int64_t tb = (b0 + bc) * V - (L - 1);
        for (int i = (int)(L - 1) + threadIdx.x / C; i < B; i += LD_U * istep) {
            float2 x[LD_U];
#pragma unroll
            for (int u = 0; u < LD_U; ++u) {
                const int ii = i + u * istep;
                x[u] = make_float2(0.f, 0.f);
                if (ii < B && tb + ii < Tl) x[u] = *reinterpret_cast<const float2 *>(X + (tb + ii) * N + n);
            }
            float acc = 0.f;
#pragma unroll
            for (int u = 0; u < LD_U; ++u) {
                const int ii = i + u * istep;
                if (ii < B && tb + ii < Tl) {
                    const float2 z = d[rev(ii, logB) * C + p];
                    const float r0 = z.x * sc - x[u].x, r1 = z.y * sc - x[u].y;
                    acc = fmaf(r0, r0, acc);
                    acc = fmaf(r1, r1, acc);
                }
            }
            accd += (double)acc;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) accd += __shfl_xor_sync(0xffffffffu, accd, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = accd;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < NT / 32; ++w) s += red[w];
        partial[bc * ny + by] = s;
    }
}

// Df[f][m][n] fp32 (row stride ldi) -> out[(l*K+k)*N + n], l < L.  grid (ceil(N / 2C), K).
// OutT = float: numW (N even, paired stores);  OutT = double: the Gram partial Rg[d][k][k'] with N = K.
template <typename OutT>
__global__ void __launch_bounds__(NT, CMF_FD_MINB)
ifft_numW_kernel(const float *__restrict__ Df, OutT *__restrict__ out, int64_t N, int64_t ldi, int64_t K, int64_t L, int B, int logB,
                 int kq, int C) {
    const int64_t KQ = kq, MROWS = 2 * kq;
    extern __shared__ float2 fd_smem[];
    float2 *d = fd_smem, *tw = fd_smem + (size_t)B * C;
    const int64_t k = blockIdx.y;
    const int p = threadIdx.x % C;
    const int64_t n = (int64_t)blockIdx.x * (2 * C) + 2 * p;
    make_twiddles(tw, B);
    {
        const int f0 = threadIdx.x / C, fstep = NT / C;
        for (int f = f0; f <= B / 2; f += 4 * fstep) {
            float2 re[4], im[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int ff = f + u * fstep;
                re[u] = make_float2(0.f, 0.f); im[u] = re[u];
                if (ff <= B / 2 && n < N) {
                    re[u] = *reinterpret_cast<const float2 *>(Df + ((int64_t)ff * MROWS + k) * ldi + n);
                    im[u] = *reinterpret_cast<const float2 *>(Df + ((int64_t)ff * MROWS + KQ + k) * ldi + n);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (f + u * fstep <= B / 2) pack_pair(d, f + u * fstep, B, C, p, re[u], im[u]);
        }
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

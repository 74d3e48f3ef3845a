// libcmf_sm100: C ABI + host orchestration of the B200 CNMF fit path (see include/cmf_sm100.h).
//
// One handle = one time-shard on one GPU.  All device memory is owned here; host arrays are
// copied during the call.  Errors never escape as C++ exceptions.
#include "../../include/cmf_sm100.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <tuple>
#include <stdexcept>
#include <string>
#include <vector>

#include <type_traits>

#include "kernels_simt.cuh"
#include "kernels_tc.cuh"
#include "kernels_fd.cuh"
#include "kernels_hals.cuh"
#include "comm.h"

namespace {

thread_local std::string g_err;

struct CmfError : std::runtime_error {
    int code;
    CmfError(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            throw CmfError(CMF_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));     \
    } while (0)

#define REQUIRE(cond, msg)                                                                         \
    do {                                                                                           \
        if (!(cond)) throw CmfError(CMF_ERR_ARG, std::string(msg));                                \
    } while (0)

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    void alloc(size_t count) {
        free();
        n = count;
        if (count) {
            CK(cudaMalloc(&p, count * sizeof(T)));
            // the zero-fill runs on the legacy default stream, which the handle's non-blocking stream is not ordered
            // against: wait for it here (allocation is never on the hot path)
            CK(cudaMemsetAsync(p, 0, count * sizeof(T), cudaStreamLegacy));
            CK(cudaStreamSynchronize(cudaStreamLegacy));
        }
    }
    void free() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { free(); }
};

}  // namespace

// ------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------
struct cmf_ctx {
    int64_t N = 0, T = 0, t0 = 0, t1 = 0, Tl = 0, K = 0, L = 0;
    int dtype = 0, alg = 0, device = 0, engine = 0;
    bool is_first = true, is_last = true;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int64_t launches = 0;
    double data_norm = 0.0;
    double data_sumsq_local = 0.0;   // ||X_owned||^2 of this shard
    double data_sumsq_global = 0.0;  // ||X||^2 over all shards (== local without a communicator)
    int loss_mode = 0;               // 0 = direct residual pass, 1 = algebraic expansion when its inputs are resident
    bool loss_mode_explicit = false; // set by cmf_set_loss_mode: engine changes then leave the mode alone
    // Calibrated expansion (loss_mode 1 at or below `loss_guard`): the expansion's error is a slowly varying bias (the
    // truncating fp32 adder of the tensor core shrinks numH and the Grams by ~1e-6), so it is measured against the direct
    // residual pass every `calib_interval` evaluations and subtracted in between; the interval doubles while the
    // prediction made with the previous bias agrees with the direct pass and halves when it does not.
    double loss_guard = 0.25;        // relative loss at or below which the expansion needs the calibration
    int calib_max_interval = 16;
    bool calib_have = false;
    double calib_bias = 0.0, calib_last_err = 0.0;
    int calib_left = 0, calib_interval = 1, calib_fail = 0;
    int64_t n_loss_direct = 0, n_loss_expansion = 0;
    void calib_reset() { calib_have = false; calib_left = 0; calib_interval = 1; calib_fail = 0; numH_dot_valid = false; }   // called wherever the factors or the data are replaced
    double pgd_stepW = 5.0, pgd_stepH = 5.0, pgd_cur_loss = 0.0;   // PGDUpdate state (pgd.jl:147-151)
    bool numH_valid = false;         // numH buffer == transconv(current W, X) and GS == W W' of the current W
    bool numH_dot_valid = false;     // scal[4] == <numH, H> of the current numH and H (left there by the MU update of H)
    bool gram_valid = false;         // exchange buffer 1 holds the local Gram/tail partial of the current H
    bool have_data = false, have_factors = false;
    // optional per-kernel-class event timing (bench.py's roofline): class -> list of event pairs
    bool profiling = false;
    struct ProfEv { int which; cudaEvent_t a, b; };
    std::vector<ProfEv> prof_events;
    int prof_open = -1;
    void prof_begin(int which) {
        if (!profiling) return;
        ProfEv e; e.which = which;
        cudaEventCreate(&e.a); cudaEventCreate(&e.b);
        cudaEventRecord(e.a, stream);
        prof_events.push_back(e);
        prof_open = (int)prof_events.size() - 1;
    }
    void prof_end() {
        if (!profiling || prof_open < 0) return;
        cudaEventRecord(prof_events[prof_open].b, stream);
        prof_open = -1;
    }
    void prof_clear() {
        for (auto &e : prof_events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
        prof_events.clear();
    }
    virtual ~cmf_ctx() {
        prof_clear();
        if (ev_a) cudaEventDestroy(ev_a);
        if (ev_b) cudaEventDestroy(ev_b);
        if (comm_stream) cudaStreamDestroy(comm_stream);
    }
    // ---- multi-GPU (the reference is single-process: no counterpart)
    cmf::Comm comm;                  // NCCL communicator of this rank (world == 1: collectives are no-ops)
    cudaStream_t comm_stream = nullptr;  // side stream for collectives that overlap with kernels of the main stream
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    struct cmf_multi *multi = nullptr;   // set on the single-process multi-GPU handle only (cmf_create_multi)
    int world() const { return comm.world; }
    int rank() const { return comm.rank; }
    [[noreturn]] static void no_multi() { throw std::runtime_error("this call is not available on a multi-GPU group handle"); }
    virtual void get_data(void *, int) { no_multi(); }
    virtual bool tc_available() { no_multi(); }
    virtual bool fd_available() { no_multi(); }
    virtual void fd_release() { no_multi(); }
    virtual void set_data(const void *, int64_t) { no_multi(); }
    virtual void synth_data(uint64_t, int64_t, int64_t, double, double) { no_multi(); }
    virtual double data_sumsq() { no_multi(); }
    virtual void set_factors(const void *, const void *, int64_t) { no_multi(); }
    virtual void init_rand(uint64_t) { no_multi(); }
    virtual void init_scale_partials(double *) { no_multi(); }
    virtual void scale_factors(double) { no_multi(); }
    virtual void get_factors(void *, void *) { no_multi(); }
    virtual void w_partials() { no_multi(); }
    virtual void w_apply(double, double) { no_multi(); }
    virtual void w_denom_rows(int64_t, int64_t) { no_multi(); }
    virtual void w_update_rows(double, double, int64_t, int64_t) { no_multi(); }
    virtual void w_rows_buffers(void **, void **, int64_t *) { no_multi(); }
    virtual void fd_denomH_prefetch() {}
    virtual void h_update(double, double) { no_multi(); }
    virtual double loss_partial() { no_multi(); }
    virtual int loss_partial_dev() { no_multi(); }                 // leaves 1 (direct) or 2 (expansion) doubles in scalars()
    virtual double *scalars() { no_multi(); }                      // device buffer of 8 doubles
    virtual void hals_update_motifs(double, double) { no_multi(); }
    virtual double hals_update_feature_maps(double, double) { no_multi(); }
    virtual void hals_h_prepare() { no_multi(); }                  // Q = denomH - numH on the owned columns
    virtual void hals_h_local(void **, void **, int64_t *) { no_multi(); }     // device pointers of the local Q and H (owned columns), element count
    virtual void hals_h_full(void **, void **) { no_multi(); }     // rank 0 of a sharded fit: full-T Q and H buffers (allocated on first use)
    virtual void hals_h_sweep(int full, double, double) { no_multi(); }
    virtual void h_changed() { no_multi(); }                       // invalidates what depends on H (after a halo exchange / scatter)
    virtual void fd_layout(int *, int *, int64_t *) { no_multi(); }
    virtual void set_pgd_loss(int, const void *) { no_multi(); }
    virtual void set_pgd_constraints(int, int) { no_multi(); }
    virtual void pgd_update_motifs(double, double) { no_multi(); }
    virtual double pgd_update_feature_maps(double, double) { no_multi(); }
    virtual void exchange_buffer(int, void **, int64_t *, int *) { no_multi(); }
    virtual void halo_buffers(void **, void **, void **, void **, int64_t *) { no_multi(); }
    virtual void prim_conv(void *) { no_multi(); }
    virtual void prim_transconv(const void *, void *) { no_multi(); }
    virtual void prim_corr(const void *, void *) { no_multi(); }
    virtual void prim_resids(void *) { no_multi(); }
    virtual void prim_shift_and_stack(void *) { no_multi(); }
};

enum { PROF_CONV = 0, PROF_TRANSCONV = 1, PROF_CORR = 2, PROF_SWEEP = 3, PROF_NCLASS = 4 };

namespace {

using namespace cmf;


// ------------------------------------------------------------------------------------------
// tcgen05 engine state (fp32 handles only): bf16 hi/lo operand copies and their TMA tensor maps
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !ptr) throw CmfError(CMF_ERR_CUDA, "cuTensorMapEncodeTiled not available");
        fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

// 2-D bf16 tensor map: dim0 (contiguous) x dim1 with row stride `stride1` bytes (rows may overlap)
CUtensorMap make_map_2d(void *base, uint64_t dim0, uint64_t dim1, uint64_t stride1, uint32_t box0, uint32_t box1,
                        CUtensorMapSwizzle swz) {
    CUtensorMap m;
    cuuint64_t gdim[2] = {dim0, dim1};
    cuuint64_t gstr[1] = {stride1};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_tiled_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstr, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw CmfError(CMF_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return m;
}

// 3-D bf16 map for MN-major operands: (64 contiguous elements, rows with stride `row_stride` bytes,
// groups of 64 elements 128 bytes apart); box = (64, box_rows, box_groups), SWIZZLE_128B.
CUtensorMap make_map_mn(void *base, uint64_t rows, uint64_t row_stride, uint64_t groups, uint32_t box_rows, uint32_t box_groups) {
    CUtensorMap m;
    cuuint64_t gdim[3] = {64, rows, groups};
    cuuint64_t gstr[2] = {row_stride, 128};
    cuuint32_t box[3] = {64, box_rows, box_groups};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode_tiled_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, gdim, gstr, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw CmfError(CMF_ERR_CUDA, "cuTensorMapEncodeTiled (3-D) failed with code " + std::to_string((int)r));
    return m;
}

struct TcState {
    bool ok = false, tried = false;
    int Kp = 0, G = 0, num_sms = 148;
    int64_t KLp = 0, rows_u = 0, groups = 0, hrows = 0;
    DevBuf<__nv_bfloat16> X_hi, X_lo, Hw_hi, Hw_lo, Hm_hi, Hm_lo, Wc_hi, Wc_lo, Wu_hi, Wu_lo, Cc_hi, Cc_lo;
    CUtensorMap mWc[2], mHw[2], mHm[2], mWu[2], mXk[2], mXmn[2];
    CUtensorMap mHmn[2];   // H (owned + right halo) as the MN-major "X" operand of the Gram correlation
    CUtensorMap mCu[2];    // C table as the A operand of denomH
    CUtensorMap mHk[2];    // H (both halos) as the K-major "X" operand of denomH
    CUtensorMap mGc[2];    // G = Htilde Htilde' as the A operand of G*W
    CUtensorMap mWuB[2], mWcB[2];   // Wu / Wc with 256-row boxes (B operands of the plain GEMMs)
    DevBuf<__nv_bfloat16> Gc_hi, Gc_lo;
    int nsplit = 1, nsplit_g = 1;
    int64_t split_len = 0, split_len_g = 0, tiles_m = 0, tiles_n = 0, groups_c = 0, rows_c = 0;
    bool x_dirty = true, w_dirty = true;
};

// frequency-domain engine state (kernels_fd.cuh): spectrum of X (built once per data set) and the per-iteration spectra
struct FdState {
    bool ok = false, tried = false;
    int B = 0, logB = 0, V = 0, F = 0;
    int64_t nblk = 0, nblkp = 0;          // overlap-save blocks covering the owned columns; padded to a multiple of 16
    int Kq = 64, MR = 128;                // components padded to Kq = 64 or 128; MR = 2 Kq rows (real | imaginary) per frequency
    int V2 = 0;                           // hop of the denomH blocking (2L-1 lags): B - 2L + 2
    int64_t nblk2 = 0;
    DevBuf<__nv_bfloat16> Xf_hi, Xf_lo, Ah_hi, Ah_lo, Aw_hi, Aw_lo, Hf_hi, Hf_lo, Ac_hi, Ac_lo;
    DevBuf<float> Of, Df, Gf, Yf;
    DevBuf<__nv_bfloat16> Awm_hi, Awm_lo; // direct loss pass: spectrum of W as the A operand of TC_FQX (allocated on first use)
    int64_t nbc = 0;                      // blocks per chunk of the direct loss pass (Yf holds F x nbc x 2 x N floats)
    bool wx_dirty = true;
    CUtensorMap mXfK[2], mXfMN[2], mAw[2], mAh[2], mHfMN[2], mAc[2], mHf2K[2], mAwm[2], mHcK[2];
    bool x_dirty = true, w_dirty = true, h_dirty = true;   // h_dirty: Ah does not hold the spectrum of the current H
    bool hf2_valid = false;               // Hf holds the denomH blocking (hop V2) of the current H (prefetched while W is all-gathered)
    std::string why;                      // reason the engine is unavailable
    void release() {
        Xf_hi.free(); Xf_lo.free(); Ah_hi.free(); Ah_lo.free(); Aw_hi.free(); Aw_lo.free(); Of.free(); Df.free();
        Hf_hi.free(); Hf_lo.free(); Gf.free(); Ac_hi.free(); Ac_lo.free();
        Yf.free(); Awm_hi.free(); Awm_lo.free(); nbc = 0; wx_dirty = true;
        ok = tried = false;
        x_dirty = w_dirty = h_dirty = true; hf2_valid = false;
    }
};

template <typename S>
struct Ctx : cmf_ctx {
    TcState tcs;
    FdState fds;
    DevBuf<S> X, Hbuf, Wi, Wtmp, numW, denW, GS, Cf, numH, denH, tailC;
    DevBuf<int> progress, lockstep;
    DevBuf<S> tailCt;                  // HALS H sweep: truncated lag tables of all component pairs
    int hals_grid = 0;
    DevBuf<double> exch1, corr_part, loss_part, scal;
    S *H = nullptr;  // owned column 0 inside Hbuf
    int nsplit_w = 1, nsplit_g = 1;
    int64_t split_w = 0, split_g = 0;
    int conv_blocks_max = 0;

    // conv tiling per dtype
    static constexpr int TN = sizeof(S) == 4 ? 8 : 4;
    static constexpr int TT = 8, TY = 16;
    static constexpr int BN = 16 * TN, BT = TY * TT;

    int64_t KL() const { return K * L; }

    void hals_setup() {
        // wavefront H sweep: one cooperative launch, CTA b owns components b, b+grid, ...
        REQUIRE(L - 1 <= HW_TC, "HALS: L-1 must not exceed the sweep chunk (1024 columns)");
        const size_t smem = hals_wave_smem_elems<S>(L) * sizeof(S);
        int per_sm = 0, sms = 0, coop = 0;
        CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
        REQUIRE(coop != 0, "HALS needs cooperative launch support");
        CK(cudaFuncSetAttribute(hals_h_wave_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hals_h_wave_kernel<S>, HW_NT, smem));
        hals_grid = (int)std::min<int64_t>(K, (int64_t)per_sm * sms);
        REQUIRE(hals_grid >= 1, "HALS: H sweep kernel does not fit on the device");
        tailC.alloc((size_t)hals_grid * (size_t)(L * L));
        progress.alloc((size_t)K);
    }

    void init() {
        CK(cudaSetDevice(device));
        CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        own_stream = true;
        const int64_t hal = L - 1;
        X.alloc((size_t)((Tl + hal) * N));
        Hbuf.alloc((size_t)((Tl + 2 * hal) * K));
        H = Hbuf.p + hal * K;
        Wi.alloc((size_t)(KL() * N));
        Wtmp.alloc((size_t)(KL() * N));
        numW.alloc((size_t)(KL() * N));
        denW.alloc((size_t)(KL() * N));
        GS.alloc((size_t)(KL() * KL()));
        Cf.alloc((size_t)((2 * L - 1) * K * K));
        numH.alloc((size_t)(Tl * K));
        denH.alloc((size_t)(std::max<int64_t>(Tl * K, Tl)));
        exch1.alloc((size_t)(L * K * K + hal * K + 1));
        scal.alloc(8);
        if (alg == CMF_HALS) hals_setup();
        // corr splits: numW (Nin = N) and Gram (Nin = K)
        plan_split(N, Tl + hal, nsplit_w, split_w);
        plan_split(K, Tl + hal, nsplit_g, split_g);
        corr_part.alloc((size_t)std::max<int64_t>((int64_t)nsplit_w * KL() * N, (int64_t)nsplit_g * KL() * K));
        conv_blocks_max = (int)(cdiv(N, BN) * cdiv(Tl + hal, BT));
        loss_part.alloc((size_t)std::max(2 * conv_blocks_max, 4096));
        // the tensor-core engine is set up right away only where it is the default (big contractions); small problems
        // build it lazily on cmf_set_engine(h, 1)
        if (sizeof(S) == 4 && N >= 256 && Tl >= 4096 && K * L >= 256) {
            tc_setup();
            // frequency-domain engine by default where it wins: enough lags that 2*L direct flops per element exceed the
            // ~8.5 of the per-frequency products, and room for the spectrum of X (CMF_ENGINE_DEFAULT=1 keeps engine 1)
            const char *e = getenv("CMF_ENGINE_DEFAULT");
            if (tcs.ok && engine == 1 && L >= 8 && K <= fd::KQ_MAX && !(e && atoi(e) == 1)) {
                fd_setup();
                // with this engine the expansion loss is the default too: the direct pass would cost 10x the iteration
                if (fds.ok) { engine = 2; loss_mode = 1; }
            }
        }
        CK(cudaStreamSynchronize(stream));
    }

    // ---------------------------------------------------------------- tcgen05 engine (fp32 only)
    bool tc_active() const { return engine >= 1 && tcs.ok; }
    bool fd_active() const { return engine == 2 && tcs.ok && fds.ok; }
    bool fd_available() override {
        if (!tc_available()) return false;
        if (!fds.ok && !fds.tried) fd_setup();
        if (fds.ok) { tcs.X_hi.free(); tcs.X_lo.free(); tcs.x_dirty = true; }   // the two engines never hold both copies of X
        return fds.ok;
    }
    void fd_release() override { fds.release(); }
    void fd_layout(int *B, int *V, int64_t *nblk) override {
        *B = fd_active() ? fds.B : 0; *V = fd_active() ? fds.V : 0; *nblk = fd_active() ? fds.nblkp : 0;
    }
    void mark_w_dirty() { tcs.w_dirty = true; fds.w_dirty = true; fds.wx_dirty = true; }

    // ---------------------------------------------------------------- frequency-domain engine (fp32, K <= 64, L <= 256)
    void fd_setup() {
        if constexpr (!std::is_same<S, float>::value) { return; } else {
            FdState &f = fds;
            f.tried = true;
            if (!tcs.ok) { f.why = "needs the tcgen05 engine"; return; }
            if (K > fd::KQ_MAX || L > 256 || N < 16) { f.why = "needs K <= 128, L <= 256, N >= 16"; return; }
            f.Kq = (K <= 64) ? 64 : 128;
            f.MR = 2 * f.Kq;
            // block length: the smallest power of two >= 8L, at most 1024 (>= 4L always: L <= 256).  A/B on B200 with the round-2
            // transforms, same box: c4 (L = 100) 512 -> 45.9 ms, 1024 -> 43.2 ms per iteration; c3 (L = 50) 256 -> 3.86, 512 -> 3.48,
            // 1024 -> 3.92 ms: the products move and multiply (B / V)(F / (B/2)) ~ 1.24 / 1.11 / 1.05 times the size of X at
            // B = 4L.. / 8L.. / 16L.., the transforms get longer
            int B0 = 64;
            while (B0 < 8 * L && B0 < 1024) B0 *= 2;
            while (B0 < 4 * L) B0 *= 2;
            // ... unless this shard is so short that the longer block only adds padding: the numH product works on tiles of 256
            // blocks, so its work goes like F * 256 * ceil(blocks / 256), the numW product's like F * blocks.  A rank of an
            // 8-GPU c4 fit (524288 columns) has 1270 blocks of 512 (5 tiles, 1 % padding) or 567 blocks of 1024 (3 tiles, 26 %
            // padding): the shorter block is kept when the longer one does not win at least 3 % on that count.
            if (B0 / 2 >= 4 * L && B0 / 2 >= 64) {
                auto work = [&](int B) {
                    const int64_t nb = cdiv(Tl, (int64_t)(B - (int)L + 1));
                    return (double)(B / 2 + 1) * (double)(cdiv(nb, tc::BN) * tc::BN + cdiv(nb, 16) * 16);
                };
                if (work(B0) > 0.97 * work(B0 / 2)) B0 /= 2;
            }
            int forced = 0;
            if (const char *e = getenv("CMF_FD_B")) {
                const int want = atoi(e);
                if (want >= 4 * L && want >= 64 && want <= 1024 && (want & (want - 1)) == 0) forced = want;
            }
            tcs.X_hi.free(); tcs.X_lo.free(); tcs.x_dirty = true;
            size_t free_b = 0, total_b = 0;
            CK(cudaMemGetInfo(&free_b, &total_b));
            // what this handle allocates later, beside the spectra: the HALS round kernel's component-major copies of H and
            // Delta H and its ring of block partials; the direct loss pass's spectrum of W and a minimal chunk buffer
            size_t later = (size_t)1 << 30;
            if (alg == CMF_HALS) later += (size_t)2 * (size_t)(Tl + 2048) * (size_t)K * 4 + (size_t)hals2::RING * (size_t)cdiv(K, hals2::GS) * (size_t)K * hals2::CW * 4;
            size_t xf = 0, ah = 0, aw = 0, of = 0, df = 0, hf = 0, gf = 0, ac = 0;
            bool fits = false;
            // the next power of two moves fewer bytes still (hop / length grows), so it is the fallback when the first choice does
            // not fit beside the data
            for (int B = forced ? forced : B0; B <= (forced ? forced : std::min(2 * B0, 1024)); B *= 2) {
                int logB = 0;
                while ((1 << logB) < B) ++logB;
                f.B = B; f.logB = logB; f.V = B - (int)L + 1; f.F = B / 2 + 1;
                f.nblk = cdiv(Tl, f.V);
                f.nblkp = cdiv(f.nblk, 16) * 16;
                xf = (size_t)f.F * (size_t)f.nblkp * 2 * (size_t)N + 64;
                ah = (size_t)f.F * (size_t)f.nblkp * 2 * f.MR + 64;
                aw = (size_t)f.F * f.MR * 2 * (size_t)N + 64;
                f.V2 = B - 2 * (int)L + 2;
                f.nblk2 = cdiv(Tl, f.V2);
                of = (size_t)f.F * (size_t)std::max(f.nblkp, f.nblk2) * f.MR; df = (size_t)f.F * f.MR * (size_t)N;
                // Hf serves the Gram partial (nblkp blocks) and, later on the same stream, denomH (nblk2 blocks)
                hf = (size_t)f.F * (size_t)std::max(f.nblkp, f.nblk2) * 2 * f.Kq + 256; gf = (size_t)f.F * f.MR * f.Kq;
                ac = (size_t)f.F * f.MR * f.MR;
                const size_t need = 4 * (xf + ah + aw + hf + ac) + 4 * (of + df + gf);
                const size_t loss_pass = 4 * aw + (size_t)256 * (size_t)f.F * 2 * (size_t)N * sizeof(float);
                if (need + later + loss_pass <= free_b) { fits = true; break; }
            }
            if (!fits) { f.why = "not enough device memory for the spectrum of X"; return; }
            f.Xf_hi.alloc(xf); f.Xf_lo.alloc(xf);
            f.Ah_hi.alloc(ah); f.Ah_lo.alloc(ah);
            f.Aw_hi.alloc(aw); f.Aw_lo.alloc(aw);
            f.Of.alloc(of); f.Df.alloc(df);
            f.Hf_hi.alloc(hf); f.Hf_lo.alloc(hf); f.Gf.alloc(gf);
            f.Ac_hi.alloc(ac); f.Ac_lo.alloc(ac);
            __nv_bfloat16 *xs[2] = {f.Xf_hi.p, f.Xf_lo.p}, *as[2] = {f.Ah_hi.p, f.Ah_lo.p}, *ws[2] = {f.Aw_hi.p, f.Aw_lo.p};
            const uint64_t rows = (uint64_t)f.F * (uint64_t)f.nblkp;
            for (int i = 0; i < 2; ++i) {
                f.mXfK[i] = make_map_2d(xs[i], (uint64_t)(2 * N), rows, (uint64_t)N * 4, tc::BK, tc::BN, CU_TENSOR_MAP_SWIZZLE_64B);
                f.mXfMN[i] = make_map_mn(xs[i], rows * 2, (uint64_t)N * 2, (uint64_t)cdiv(N, 64), tc::BK, 4);
                f.mAw[i] = make_map_2d(ws[i], (uint64_t)(2 * N), (uint64_t)f.F * f.MR, (uint64_t)N * 4, tc::BK, tc::BM, CU_TENSOR_MAP_SWIZZLE_64B);
                f.mAh[i] = make_map_mn(as[i], rows * 2, (uint64_t)f.MR * 2, (uint64_t)(f.MR / 64), tc::BK, 2);
                f.mHfMN[i] = make_map_mn(i == 0 ? f.Hf_hi.p : f.Hf_lo.p, rows * 2, (uint64_t)f.Kq * 2, (uint64_t)(f.Kq / 64), tc::BK, 4);
                f.mAc[i] = make_map_2d(i == 0 ? f.Ac_hi.p : f.Ac_lo.p, f.MR, (uint64_t)f.F * f.MR, (uint64_t)f.MR * 2, tc::BK, tc::BM, CU_TENSOR_MAP_SWIZZLE_64B);
                f.mHf2K[i] = make_map_2d(i == 0 ? f.Hf_hi.p : f.Hf_lo.p, f.MR, (uint64_t)f.F * (uint64_t)f.nblk2, (uint64_t)f.MR * 2, tc::BK, tc::BN, CU_TENSOR_MAP_SWIZZLE_64B);
            }
            CK(cudaFuncSetAttribute(tc::tc_kernel<tc::TC_FQT>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
            CK(cudaFuncSetAttribute(tc::tc_kernel<tc::TC_FQC>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
            CK(cudaFuncSetAttribute(tc::tc_kernel<tc::TC_FQX>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));

            const int big = (int)fd_smem(fd_cols_h()), small_ = (int)fd_smem(16);
            CK(cudaFuncSetAttribute(fd::fft_x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, small_));
            CK(cudaFuncSetAttribute(fd::ifft_resid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, small_));
            CK(cudaFuncSetAttribute(fd::fft_w_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, small_));
            CK(cudaFuncSetAttribute(fd::ifft_numW_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, small_));
            CK(cudaFuncSetAttribute(fd::ifft_numW_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, small_));
            CK(cudaFuncSetAttribute(fd::fft_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
            CK(cudaFuncSetAttribute(fd::ifft_numH_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
            f.x_dirty = f.w_dirty = f.h_dirty = true;
            f.ok = true;
        }
    }
    // component pairs per CTA of the H-side transforms: 32 (all 64 rows) while the tile fits in shared memory
    int fd_cols_h() const {
        if (const char *e = getenv("CMF_FD_COLS")) {
            const int c = atoi(e);
            if ((c == 8 || c == 16 || c == 32) && (size_t)fds.B * (size_t)c * sizeof(float2) <= 160 * 1024) return c;
        }
        return fds.B >= 1024 ? 8 : 16;          // tiles of <= 64 KB + twiddles: three CTAs per SM (c4 at B = 1024: 44.96 -> 43.23 ms)
    }
    // scheduling order of the 1-D grids of the transforms: 0 = tile index fastest, block count = block index fastest (CMF_FD_ORDER=1)
    int64_t fd_order(int64_t nblocks) const { static const int o = getenv("CMF_FD_ORDER") ? atoi(getenv("CMF_FD_ORDER")) : 0; return o ? nblocks : 0; }
    // complex columns per CTA of the W-side transforms (fft_w / ifft_numW): as for the H side, 8 at B = 1024 keeps three CTAs per SM
    int fd_cols_w() const { return fds.B >= 1024 ? 8 : 16; }
    size_t fd_smem(int C) const { return ((size_t)fds.B * (size_t)C + (size_t)fds.B) * sizeof(float2); }   // tile + full-circle twiddle table

    void fd_build_X() {
        if constexpr (std::is_same<S, float>::value) {
            FdState &f = fds;
            if (!f.x_dirty) return;
            dim3 grid((unsigned)(f.nblkp * cdiv(N, 32)));
            fd::fft_x_kernel<<<grid, fd::NT, fd_smem(16), stream>>>(X.p, f.Xf_hi.p, f.Xf_lo.p, N, Tl + (L - 1), f.B, f.logB, f.V, f.nblkp, fd_order(f.nblkp));
            post_launch();
            f.x_dirty = false;
        }
    }
    // numW through the spectrum of X (mult.jl:32)
    void fd_corr() {
        if constexpr (std::is_same<S, float>::value) {
            FdState &f = fds;
            fd_build_X();
            fd_spectrum_H();
            tc::Params q = tc_base_params();
            q.nprod = 3;
            q.tiles_n = cdiv(N, tc::BN);
            q.nkb = f.nblkp * 2 / tc::BK;
            q.MR = f.MR; q.mtiles = f.MR / tc::BM; q.units = (int64_t)f.F * q.tiles_n * q.mtiles;
            q.fq_rows = f.nblkp * 2;
            q.out = f.Df.p; q.Mrows = f.MR; q.Ncols = N; q.ldo = N;
            const unsigned grid = (unsigned)std::min<int64_t>(q.units, tcs.num_sms);
            prof_begin(PROF_CORR);
            tc::tc_kernel<tc::TC_FQC><<<grid, tc::THREADS, tc::SMEM_BYTES, stream>>>(f.mAh[0], f.mAh[1], f.mXfMN[0], f.mXfMN[1], q);
            prof_end();
            post_launch();
            fd::ifft_numW_kernel<float><<<dim3((unsigned)cdiv(N, 2 * fd_cols_w()), (unsigned)K), fd::NT, fd_smem(fd_cols_w()), stream>>>(f.Df.p, numW.p, N, N, K, L, f.B, f.logB, f.Kq, fd_cols_w());
            post_launch();
        }
    }
    // denomH = C (*) H through the spectrum of H (mult.jl:44,48): the numH product with the lag table in place of W and
    // H (both halos) in place of X; needs lag_tables() done.  The last L-1 columns are overwritten by the truncated tail.
    void fd_denomH() {
        if constexpr (std::is_same<S, float>::value) {
            FdState &f = fds;
            fd::fft_w_kernel<<<dim3((unsigned)cdiv(K, 2 * fd_cols_w()), (unsigned)K), fd::NT, fd_smem(fd_cols_w()), stream>>>(
                Cf.p, f.Ac_hi.p, f.Ac_lo.p, K, K, 2 * L - 1, f.B, f.logB, f.MR, f.Kq, 0, f.Kq, fd_cols_w());
            post_launch();
            const int C = fd_cols_h();
            fd_denomH_prefetch();
            f.hf2_valid = false;                                   // consumed below; Hf is shared with the Gram / loss passes
            tc::Params q = tc_base_params();
            q.nprod = 3;
            q.tiles_n = cdiv(f.nblk2, tc::BN);
            q.nkb = f.MR / tc::BK;
            q.MR = f.MR; q.mtiles = f.MR / tc::BM; q.units = (int64_t)f.F * q.tiles_n * q.mtiles;
            q.fq_rows = f.nblk2;
            q.out = f.Of.p;
            const unsigned grid = (unsigned)std::min<int64_t>(q.units, tcs.num_sms);
            tc::tc_kernel<tc::TC_FQT><<<grid, tc::THREADS, tc::SMEM_BYTES, stream>>>(f.mAc[0], f.mAc[1], f.mHf2K[0], f.mHf2K[1], q);
            post_launch();
            fd::ifft_numH_kernel<<<(unsigned)(f.nblk2 * (f.Kq / 2 / C)), fd::NT, fd_smem(C), stream>>>(
                f.Of.p, denH.p, K, Tl, f.B, f.logB, f.V2, f.nblk2, C, f.Kq, fd_order(f.nblk2));
            post_launch();
        }
    }
    // the H-only part of fd_denomH: spectrum of the blocks of H (hop V2, both halos).  Depends on H alone, so a sharded fit runs it
    // while the all-gather of the new W is in flight (r_update_motifs).
    void fd_denomH_prefetch() override {
        if constexpr (std::is_same<S, float>::value) {
            FdState &f = fds;
            if (!fd_active() || f.hf2_valid) return;
            const int C = fd_cols_h();
            fd::fft_h_kernel<<<(unsigned)(f.nblk2 * (f.Kq / 2 / C)), fd::NT, fd_smem(C), stream>>>(
                H, f.Hf_hi.p, f.Hf_lo.p, K, Tl, Tl + (L - 1), f.B, f.logB, f.V2, f.nblk2, C, 1, -(L - 1), f.Kq, fd_order(f.nblk2));
            post_launch();
            f.hf2_valid = true;
        }
    }
    // sum of squared residuals (mult.jl:55-57) through the frequency domain: Xhat^ = W^ H^ per frequency for a chunk of
    // blocks (TC_FQX), inverse transform, subtract X, sum of squares (ifft_resid_kernel); returns the number of partials.
    // ~5x cheaper than the time-domain TC_CONV pass at c4, and free of the cancellation of the expansion.
    int64_t fd_conv_loss() {
        if constexpr (std::is_same<S, float>::value) {
            FdState &f = fds;
            f.hf2_valid = false;
            const int64_t ntile32 = cdiv(N, 32);
            if (f.Awm_hi.n == 0) {
                const size_t aw = (size_t)f.F * 2 * f.MR * (size_t)N + 64;
                size_t free_b = 0, total_b = 0;
                CK(cudaMemGetInfo(&free_b, &total_b));
                const size_t per_block = (size_t)f.F * 2 * (size_t)N * sizeof(float);          // Yf bytes per block
                REQUIRE(free_b > 4 * aw + 256 * per_block + ((size_t)1 << 29), "not enough device memory for the frequency-domain loss pass");
                size_t budget = std::min<size_t>((free_b - 4 * aw - ((size_t)1 << 29)) / 2, (size_t)12 << 30);
                int64_t nbc = (int64_t)(budget / per_block) / tc::BN * tc::BN;
                nbc = std::max<int64_t>(tc::BN, std::min<int64_t>(nbc, cdiv(f.nblk, tc::BN) * tc::BN));
                if (const char *e = getenv("CMF_FD_NBC")) { const int64_t v = atoll(e); if (v >= tc::BN && v % tc::BN == 0 && v < nbc) nbc = v; }   // tests: force several chunks
                f.nbc = nbc;
                f.Awm_hi.alloc(aw); f.Awm_lo.alloc(aw);
                f.Yf.alloc((size_t)f.F * (size_t)nbc * 2 * (size_t)N);
                for (int i = 0; i < 2; ++i) {
                    f.mAwm[i] = make_map_mn(i == 0 ? f.Awm_hi.p : f.Awm_lo.p, (uint64_t)f.F * 2 * f.MR, (uint64_t)N * 2, (uint64_t)cdiv(N, 64), tc::BK, 2);
                    f.mHcK[i] = make_map_2d(i == 0 ? f.Hf_hi.p : f.Hf_lo.p, f.MR, (uint64_t)f.F * (uint64_t)f.nblk, (uint64_t)f.MR * 2, tc::BK, tc::BN, CU_TENSOR_MAP_SWIZZLE_64B);
                }
                f.wx_dirty = true;
            }
            const size_t need_lp = (size_t)(f.nblk * ntile32);
            if (loss_part.n < need_lp) loss_part.alloc(std::max<size_t>(need_lp, 4096));
            if (f.wx_dirty) {
                fd::fft_w_kernel<<<dim3((unsigned)cdiv(N, 2 * fd_cols_w()), (unsigned)K), fd::NT, fd_smem(fd_cols_w()), stream>>>(Wi.p, f.Awm_hi.p, f.Awm_lo.p, N, K, L, f.B, f.logB, 2 * N, N, 1, f.Kq, fd_cols_w());
                post_launch();
                f.wx_dirty = false;
            }
            const int C = fd_cols_h();
            fd::fft_h_kernel<<<(unsigned)(f.nblk * (f.Kq / 2 / C)), fd::NT, fd_smem(C), stream>>>(
                H, f.Hf_hi.p, f.Hf_lo.p, K, Tl, Tl + (L - 1), f.B, f.logB, f.V, f.nblk, C, 1, -(L - 1), f.Kq, fd_order(f.nblk));
            post_launch();
            prof_begin(PROF_CONV);
            for (int64_t b0 = 0; b0 < f.nblk; b0 += f.nbc) {
                const int64_t cur = std::min(f.nbc, f.nblk - b0);
                tc::Params q = tc_base_params();
                q.nprod = 3;
                q.tiles_n = cdiv(cur, tc::BN);
                q.tiles_m = cdiv(N, tc::BM);
                q.nkb = f.MR / tc::BK;
                q.MR = f.MR; q.mtiles = 1; q.units = (int64_t)f.F * 2 * q.tiles_m * q.tiles_n;
                q.fq_rows = f.nblk; q.b_off = b0; q.nbc = cur;
                q.out = f.Yf.p;
                const unsigned grid = (unsigned)std::min<int64_t>(q.units, tcs.num_sms);
                tc::tc_kernel<tc::TC_FQX><<<grid, tc::THREADS, tc::SMEM_BYTES, stream>>>(f.mAwm[0], f.mAwm[1], f.mHcK[0], f.mHcK[1], q);
                post_launch();
                fd::ifft_resid_kernel<<<(unsigned)(cur * ntile32), fd::NT, fd_smem(16), stream>>>(
                    f.Yf.p, X.p, loss_part.p + b0 * ntile32, N, Tl, L, f.B, f.logB, f.V, cur, b0, fd_order(cur));
                post_launch();
            }
            prof_end();
            return f.nblk * ntile32;
        }
        return 0;
    }
    // Ah = spectrum of the owned columns of the current H, block by block (shared by numW and the Gram partial)
    void fd_spectrum_H() {
        if constexpr (std::is_same<S, float>::value) {
            FdState &f = fds;
            if (!f.h_dirty) return;
            const int C = fd_cols_h();
            fd::fft_h_kernel<<<(unsigned)(f.nblkp * (f.Kq / 2 / C)), fd::NT, fd_smem(C), stream>>>(
                H, f.Ah_hi.p, f.Ah_lo.p, K, Tl, Tl + (L - 1), f.B, f.logB, f.V, f.nblkp, C, 0, 0, f.Kq, fd_order(f.nblkp));
            post_launch();
            f.h_dirty = false;
        }
    }
    // Gram partial Rg[d][k][k'] = sum_u H[u][k] H[u+d][k'] through the spectrum of H (u owned, u+d into the right halo)
    void fd_gram() {
        if constexpr (std::is_same<S, float>::value) {
            FdState &f = fds;
            f.hf2_valid = false;
            fd_spectrum_H();
            const int C = fd_cols_h();
            fd::fft_h_kernel<<<(unsigned)(f.nblkp * (f.Kq / 2 / C)), fd::NT, fd_smem(C), stream>>>(
                H, f.Hf_hi.p, f.Hf_lo.p, K, Tl, Tl + (L - 1), f.B, f.logB, f.V, f.nblkp, C, 1, 0, f.Kq, fd_order(f.nblkp));
            post_launch();
            tc::Params q = tc_base_params();
            q.nprod = 3;
            q.tiles_n = 1;
            q.nkb = f.nblkp * 2 / tc::BK;
            q.MR = f.MR; q.mtiles = f.MR / tc::BM; q.units = (int64_t)f.F * q.mtiles;
            q.fq_rows = f.nblkp * 2;
            q.out = f.Gf.p; q.Mrows = f.MR; q.Ncols = K; q.ldo = f.Kq;
            const unsigned grid = (unsigned)std::min<int64_t>(q.units, tcs.num_sms);
            tc::tc_kernel<tc::TC_FQC><<<grid, tc::THREADS, tc::SMEM_BYTES, stream>>>(f.mAh[0], f.mAh[1], f.mHfMN[0], f.mHfMN[1], q);
            post_launch();
            fd::ifft_numW_kernel<double><<<dim3((unsigned)cdiv(K, 2 * fd_cols_w()), (unsigned)K), fd::NT, fd_smem(fd_cols_w()), stream>>>(f.Gf.p, exch1.p, K, f.Kq, K, L, f.B, f.logB, f.Kq, fd_cols_w());
            post_launch();
        }
    }
    // numH through the spectrum of X (mult.jl:47)
    void fd_transconv() {
        if constexpr (std::is_same<S, float>::value) {
            FdState &f = fds;
            fd_build_X();
            if (f.w_dirty) {
                fd::fft_w_kernel<<<dim3((unsigned)cdiv(N, 2 * fd_cols_w()), (unsigned)K), fd::NT, fd_smem(fd_cols_w()), stream>>>(Wi.p, f.Aw_hi.p, f.Aw_lo.p, N, K, L, f.B, f.logB, 2 * N, N, 0, f.Kq, fd_cols_w());
                post_launch();
                f.w_dirty = false;
            }
            tc::Params q = tc_base_params();
            q.nprod = 3;
            q.tiles_n = cdiv(f.nblkp, tc::BN);
            q.nkb = cdiv(2 * N, tc::BK);
            q.MR = f.MR; q.mtiles = f.MR / tc::BM; q.units = (int64_t)f.F * q.tiles_n * q.mtiles;
            q.fq_rows = f.nblkp;
            q.out = f.Of.p;
            const unsigned grid = (unsigned)std::min<int64_t>(q.units, tcs.num_sms);
            prof_begin(PROF_TRANSCONV);
            tc::tc_kernel<tc::TC_FQT><<<grid, tc::THREADS, tc::SMEM_BYTES, stream>>>(f.mAw[0], f.mAw[1], f.mXfK[0], f.mXfK[1], q);
            prof_end();
            post_launch();
            const int C = fd_cols_h();
            fd::ifft_numH_kernel<<<(unsigned)(f.nblk * (f.Kq / 2 / C)), fd::NT, fd_smem(C), stream>>>(
                f.Of.p, numH.p, K, Tl, f.B, f.logB, f.V, f.nblkp, C, f.Kq, fd_order(f.nblk));
            post_launch();
        }
    }
    bool tc_available() override {
        if (!tcs.ok && !tcs.tried) tc_setup();
        return tcs.ok;
    }

    void tc_setup() {
        if constexpr (!std::is_same<S, float>::value) { return; } else {
            TcState &t = tcs;
            t.tried = true;
            if (K > 128 || N % 8 != 0) return;
            t.Kp = (int)(cdiv(K, 8) * 8);                          // TMA row strides must be multiples of 16 bytes
            t.G = 128 / t.Kp;                                      // lags per 128-row tile of the transposed conv
            t.KLp = cdiv(L * t.Kp, 64) * 64;
            t.groups = cdiv(L, t.G);
            t.rows_u = cdiv(L * t.Kp + 128, 128) * 128;             // dense rows l*Kp + k, padded so every 128-row box is in bounds
            const int64_t hal = L - 1;
            t.hrows = Tl + hal;                                   // window rows addressed (X columns incl. right halo)
            cudaDeviceProp prop;
            CK(cudaGetDeviceProperties(&prop, device));
            if (prop.major != 10) return;                         // tcgen05 needs sm_100
            t.num_sms = prop.multiProcessorCount;
            // the bf16 planes of H (4 x the size of H in bf16) are allocated on first use (tc_h_planes): a handle that runs on
            // the frequency-domain engine never touches them
            t.Wc_hi.alloc((size_t)(N * t.KLp)); t.Wc_lo.alloc((size_t)(N * t.KLp));
            t.Wu_hi.alloc((size_t)(t.rows_u * N)); t.Wu_lo.alloc((size_t)(t.rows_u * N));
            t.groups_c = cdiv(2 * L - 1, t.G);
            t.rows_c = cdiv((2 * L - 1) * t.Kp + 128, 128) * 128;
            t.Cc_hi.alloc((size_t)(t.rows_c * t.Kp)); t.Cc_lo.alloc((size_t)(t.rows_c * t.Kp));
            t.Gc_hi.alloc((size_t)(KL() * t.KLp)); t.Gc_lo.alloc((size_t)(KL() * t.KLp));
            if (GS.n < (size_t)(t.rows_u * t.rows_u)) GS.alloc((size_t)(t.rows_u * t.rows_u));
            // tensor maps (index 0 = hi plane, 1 = lo plane)
            __nv_bfloat16 *wc[2] = {t.Wc_hi.p, t.Wc_lo.p};
            __nv_bfloat16 *wu[2] = {t.Wu_hi.p, t.Wu_lo.p};
            for (int i = 0; i < 2; ++i) {
                t.mWc[i] = make_map_2d(wc[i], (uint64_t)t.KLp, (uint64_t)N, (uint64_t)t.KLp * 2, tc::BK, tc::BM, CU_TENSOR_MAP_SWIZZLE_64B);
                t.mWu[i] = make_map_2d(wu[i], (uint64_t)N, (uint64_t)t.rows_u, (uint64_t)N * 2, tc::BK, tc::BM, CU_TENSOR_MAP_SWIZZLE_64B);
                // denomH: A = C table rows (d'*Kp + k) x Kp
                __nv_bfloat16 *cc[2] = {t.Cc_hi.p, t.Cc_lo.p};
                t.mCu[i] = make_map_2d(cc[i], (uint64_t)t.Kp, (uint64_t)t.rows_c, (uint64_t)t.Kp * 2, tc::BK, tc::BM, CU_TENSOR_MAP_SWIZZLE_64B);
                __nv_bfloat16 *gc[2] = {t.Gc_hi.p, t.Gc_lo.p};
                t.mWuB[i] = make_map_2d(wu[i], (uint64_t)N, (uint64_t)t.rows_u, (uint64_t)N * 2, tc::BK, tc::BN, CU_TENSOR_MAP_SWIZZLE_64B);
                t.mWcB[i] = make_map_2d(wc[i], (uint64_t)t.KLp, (uint64_t)N, (uint64_t)t.KLp * 2, tc::BK, tc::BN, CU_TENSOR_MAP_SWIZZLE_64B);
                t.mGc[i] = make_map_2d(gc[i], (uint64_t)t.KLp, (uint64_t)KL(), (uint64_t)t.KLp * 2, tc::BK, tc::BM, CU_TENSOR_MAP_SWIZZLE_64B);
            }
            // correlation work split: balance persistent CTAs, keep splits >= FLUSH_T columns
            t.tiles_m = cdiv(t.KLp, tc::BM);
            t.tiles_n = cdiv(N, tc::BN);
            const int64_t tau = Tl + hal, base = t.tiles_m * t.tiles_n;
            double best = -1.0;
            for (int ns = 1; ns <= 16; ++ns) {
                const int64_t sl = cdiv(cdiv(tau, ns), tc::BK) * tc::BK;
                if (ns > 1 && sl < tc::FLUSH_T) break;
                const int64_t units = base * cdiv(tau, sl);
                const double eff = (double)units / (double)(cdiv(units, t.num_sms) * t.num_sms);
                if (eff > best + 1e-9) { best = eff; t.nsplit = (int)cdiv(tau, sl); t.split_len = sl; }
            }
            // Gram correlation (one n tile): enough splits to fill the machine
            {
                const int64_t want = std::max<int64_t>(1, cdiv(2 * t.num_sms, t.tiles_m));
                const int64_t maxs = std::max<int64_t>(1, tau / tc::FLUSH_T);
                const int64_t ns = std::min(want, maxs);
                t.split_len_g = cdiv(cdiv(tau, ns), tc::BK) * tc::BK;
                t.nsplit_g = (int)cdiv(tau, t.split_len_g);
            }
            const size_t need = std::max((size_t)t.nsplit * (size_t)(KL() * N), (size_t)t.nsplit_g * (size_t)(KL() * K));
            if (corr_part.n < need) corr_part.alloc(need);
            const size_t need_lp = (size_t)(tc::EPI_WARPS * cdiv(Tl, tc::BN) * cdiv(N, tc::BM));
            if (loss_part.n < need_lp) loss_part.alloc(std::max<size_t>(need_lp, 4096));
            CK(cudaFuncSetAttribute(tc::tc_kernel<tc::TC_CONV>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
            CK(cudaFuncSetAttribute(tc::tc_kernel<tc::TC_TRANS>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
            CK(cudaFuncSetAttribute(tc::tc_kernel<tc::TC_CORR>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
            CK(cudaFuncSetAttribute(tc::tc_kernel<tc::TC_PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
            t.ok = true;
            t.x_dirty = t.w_dirty = true;
            // default engine: tensor cores when the contraction is big enough to fill 128x256 tiles
            if (N >= 256 && Tl >= 4096 && K * L >= 256) engine = 1;
        }
    }

    tc::Params tc_base_params() {
        tc::Params q;
        memset(&q, 0, sizeof(q));
        q.N = N; q.K = K; q.L = L; q.Tl = Tl; q.G = tcs.G; q.Kp = tcs.Kp;
        { const char *e = getenv("CMF_PROMO"); q.promo = e ? std::max(1, atoi(e)) : tc::PROMO; }
        { const char *e = getenv("CMF_TC_PRODUCTS"); q.nprod = (e && atoi(e) == 2) ? 2 : 3; }   // experiment knob, default 3
        return q;
    }

    // the time-domain bf16 planes of X are allocated on first use: the frequency-domain engine never needs them
    void tc_split_X() {
        if constexpr (std::is_same<S, float>::value) {
            if (!tcs.ok || !tcs.x_dirty) return;
            TcState &t = tcs;
            if (t.X_hi.n == 0) {
                const int64_t hal = L - 1;
                t.X_hi.alloc(X.n + 64); t.X_lo.alloc(X.n + 64);      // +64: the 3-D maps may touch one atom past the last row
                __nv_bfloat16 *xs[2] = {t.X_hi.p, t.X_lo.p};
                for (int i = 0; i < 2; ++i) {
                    t.mXk[i] = make_map_2d(xs[i], (uint64_t)N, (uint64_t)(Tl + hal), (uint64_t)N * 2, tc::BK, tc::BN, CU_TENSOR_MAP_SWIZZLE_64B);
                    t.mXmn[i] = make_map_mn(xs[i], (uint64_t)(Tl + hal), (uint64_t)N * 2, (uint64_t)cdiv(N, 64), tc::BK, 4);
                }
            }
            tc::split_plain_kernel<<<(unsigned)cdiv((int64_t)X.n, 256), 256, 0, stream>>>(X.p, tcs.X_hi.p, tcs.X_lo.p, (int64_t)X.n);
            post_launch();
            tcs.x_dirty = false;
        }
    }
    void tc_split_W() {
        if constexpr (std::is_same<S, float>::value) {
            if (!tcs.w_dirty) return;
            tc::split_W_kernel<<<dim3((unsigned)cdiv(N, 32), (unsigned)cdiv(tcs.Kp, 32), (unsigned)L), dim3(32, 8), 0, stream>>>(
                Wi.p, tcs.Wc_hi.p, tcs.Wc_lo.p, tcs.Wu_hi.p, tcs.Wu_lo.p, N, K, L, tcs.Kp, tcs.KLp);
            post_launch();
            tcs.w_dirty = false;
        }
    }
    // bf16 hi/lo planes of H for the time-domain tensor-core kernels and the maps over them, on first use
    void tc_h_planes() {
        if constexpr (std::is_same<S, float>::value) {
            TcState &t = tcs;
            if (t.Hw_hi.n != 0) return;
            const int64_t hal = L - 1;
            const size_t hw_elems = (size_t)((Tl + 2 * hal) * t.Kp + t.KLp + 64);
            t.Hw_hi.alloc(hw_elems); t.Hw_lo.alloc(hw_elems);
            t.Hm_hi.alloc(hw_elems); t.Hm_lo.alloc(hw_elems);
            __nv_bfloat16 *hw[2] = {t.Hw_hi.p, t.Hw_lo.p}, *hm[2] = {t.Hm_hi.p, t.Hm_lo.p};
            for (int i = 0; i < 2; ++i) {
                // overlapping-row "window" map: row t starts at element t*Kp and is KLp long
                t.mHw[i] = make_map_2d(hw[i], (uint64_t)t.KLp, (uint64_t)t.hrows, (uint64_t)t.Kp * 2, tc::BK, tc::BN, CU_TENSOR_MAP_SWIZZLE_64B);
                t.mHm[i] = make_map_mn(hm[i], (uint64_t)t.hrows, (uint64_t)t.Kp * 2, (uint64_t)(t.KLp / 64), tc::BK, 2);
                // Gram: "X" = H rows from owned column 0 (owned + right halo), [t][Kp] MN-major
                t.mHmn[i] = make_map_mn(hw[i] + hal * t.Kp, (uint64_t)(Tl + hal), (uint64_t)t.Kp * 2, (uint64_t)cdiv(t.Kp, 64), tc::BK, 4);
                // denomH: "X" = H rows from the left halo on, K-major
                t.mHk[i] = make_map_2d(hw[i], (uint64_t)t.Kp, (uint64_t)(Tl + 2 * hal), (uint64_t)t.Kp * 2, tc::BK, tc::BN, CU_TENSOR_MAP_SWIZZLE_64B);
            }
        }
    }
    void tc_split_H(bool masked) {
        if constexpr (std::is_same<S, float>::value) {
            tc_h_planes();
            const int64_t rows = Tl + 2 * (L - 1);
            tc::split_H_kernel<<<(unsigned)cdiv(rows * tcs.Kp, 256), 256, 0, stream>>>(
                Hbuf.p, masked ? tcs.Hm_hi.p : tcs.Hw_hi.p, masked ? tcs.Hm_lo.p : tcs.Hw_lo.p, rows, K, tcs.Kp,
                L - 1, L - 1 + Tl, masked ? 1 : 0);
            post_launch();
        }
    }

    // numW via tensor cores (mult.jl:32)
    void tc_corr() {
        if constexpr (std::is_same<S, float>::value) {
            if (fd_active()) { fd_corr(); return; }
            tc_split_X();
            tc_split_H(true);
            tc::Params q = tc_base_params();
            q.tiles_m = tcs.tiles_m; q.tiles_n = tcs.tiles_n; q.split_len = tcs.split_len; q.tau_hi = Tl + (L - 1);
            q.units = tcs.tiles_m * tcs.tiles_n * tcs.nsplit;
            q.part = corr_part.p;
            q.corr_order = 0;   // j tiles fastest: the CTAs that run together share the X tile (A/B: 124 vs 148 ms at T=1M)
            {
                const char *e = getenv("CMF_LOCKSTEP");            // k-block window; 0 disables (diagnostics)
                const int w = e ? atoi(e) : 256;
                if (w > 0) {
                    if (lockstep.n == 0) lockstep.alloc(1024);
                    CK(cudaMemsetAsync(lockstep.p, 0, lockstep.n * sizeof(int), stream));
                    q.lockstep = lockstep.p; q.lockstep_window = w;
                }
            }
            const unsigned grid = (unsigned)std::min<int64_t>(q.units, tcs.num_sms);
            prof_begin(PROF_CORR);
            if (q.lockstep != nullptr) {
                // CTAs wait on one another inside this launch: co-residency must be guaranteed
                void *args[] = {&tcs.mHm[0], &tcs.mHm[1], &tcs.mXmn[0], &tcs.mXmn[1], &q};
                CK(cudaLaunchCooperativeKernel((void *)tc::tc_kernel<tc::TC_CORR>, dim3(grid), dim3(tc::THREADS), args, tc::SMEM_BYTES, stream));
            } else {
                tc::tc_kernel<tc::TC_CORR><<<grid, tc::THREADS, tc::SMEM_BYTES, stream>>>(tcs.mHm[0], tcs.mHm[1], tcs.mXmn[0], tcs.mXmn[1], q);
            }
            prof_end();
            post_launch();
            const int64_t n = KL() * N;
            reduce_partials_kernel<S><<<(unsigned)cdiv(n, 256), 256, 0, stream>>>(corr_part.p, tcs.nsplit, n, numW.p, nullptr);
            post_launch();
        }
    }
    // plain GEMM out[m*ldo + n] = sum_k A[m][k] B[n][k] on tensor cores (operands given as hi/lo tensor maps)
    void tc_plain(const CUtensorMap *mA, const CUtensorMap *mB, S *out, int64_t Mrows, int64_t Ncols, int64_t Kdim, int64_t ldo, bool sym = false) {
        if constexpr (std::is_same<S, float>::value) {
            tc::Params q = tc_base_params();
            q.sym = (sym && !getenv("CMF_S2_FULL")) ? 1 : 0;
            q.tiles_n = cdiv(Ncols, tc::BN);
            q.nkb = cdiv(Kdim, tc::BK);
            q.units = q.tiles_n * cdiv(Mrows, tc::BM);
            q.out = out; q.Mrows = Mrows; q.Ncols = Ncols; q.ldo = ldo;
            const unsigned grid = (unsigned)std::min<int64_t>(q.units, tcs.num_sms);
            tc::tc_kernel<tc::TC_PLAIN><<<grid, tc::THREADS, tc::SMEM_BYTES, stream>>>(mA[0], mA[1], mB[0], mB[1], q);
            post_launch();
            if (q.sym) {
                tc::mirror_upper_kernel<<<dim3((unsigned)cdiv(Mrows, 256), (unsigned)Mrows), 256, 0, stream>>>(out, Mrows, ldo);
                post_launch();
            }
        }
    }
    // denomW = G * Wi (mult.jl:28,33): G is built straight into bf16 planes in the K-dim order of Wc
    void tc_denomW(int64_t j0 = 0, int64_t j1 = -1) {
        if (j1 < 0) j1 = KL();
        if constexpr (std::is_same<S, float>::value) {
            if (getenv("CMF_G_DIRECT"))          // element-wise form (kept for A/B)
                tc::build_G_split_kernel<<<(unsigned)cdiv(KL() * tcs.KLp, 256), 256, 0, stream>>>(
                    exch1.p, exch1.p + L * K * K, tcs.Gc_hi.p, tcs.Gc_lo.p, K, L, tcs.Kp, tcs.KLp);
            else
                tc::build_G_split_diag_kernel<<<(unsigned)cdiv((2 * L - 1) * K * K, 256), 256, 0, stream>>>(
                    exch1.p, exch1.p + L * K * K, tcs.Gc_hi.p, tcs.Gc_lo.p, K, L, tcs.Kp, tcs.KLp);
            post_launch();
            tc_split_W();
            if (j0 == 0 && j1 == KL()) {
                tc_plain(tcs.mGc, tcs.mWcB, denW.p, KL(), N, tcs.KLp, N);
            } else {
                // rows [j0, j1) of G only (the W update is sharded by rows over the ranks): A maps over that row block
                CUtensorMap mg[2];
                __nv_bfloat16 *gc[2] = {tcs.Gc_hi.p, tcs.Gc_lo.p};
                for (int i = 0; i < 2; ++i)
                    mg[i] = make_map_2d(gc[i] + j0 * tcs.KLp, (uint64_t)tcs.KLp, (uint64_t)(j1 - j0), (uint64_t)tcs.KLp * 2, tc::BK, tc::BM, CU_TENSOR_MAP_SWIZZLE_64B);
                tc_plain(mg, tcs.mWcB, denW.p + j0 * N, j1 - j0, N, tcs.KLp, N);
            }
        }
    }
    // Gram partial Rg[d][k][k'] = sum_u H[u][k] H[u+d][k'] via tensor cores (needs tc_split_H(true) and (false) done)
    void tc_gram() {
        if constexpr (std::is_same<S, float>::value) {
            tc::Params q = tc_base_params();
            q.N = K;                                   // output row stride / column bound: Rg is [L][K][K]
            q.tiles_m = tcs.tiles_m; q.tiles_n = 1; q.split_len = tcs.split_len_g; q.tau_hi = Tl + (L - 1);
            q.units = tcs.tiles_m * tcs.nsplit_g;
            q.part = corr_part.p;
            const unsigned grid = (unsigned)std::min<int64_t>(q.units, tcs.num_sms);
            tc::tc_kernel<tc::TC_CORR><<<grid, tc::THREADS, tc::SMEM_BYTES, stream>>>(tcs.mHm[0], tcs.mHm[1], tcs.mHmn[0], tcs.mHmn[1], q);
            post_launch();
            const int64_t n = L * K * K;
            reduce_partials_kernel<S><<<(unsigned)cdiv(n, 256), 256, 0, stream>>>(corr_part.p, tcs.nsplit_g, n, nullptr, exch1.p);
            post_launch();
        }
    }
    // denomH = C (*) H via tensor cores (mult.jl:44,48); needs lag_tables() and tc_split_H(false) done
    void tc_denomH(bool resplit_C = true) {
        if constexpr (std::is_same<S, float>::value) {
            if (resplit_C) {
                tc::split_C_kernel<<<(unsigned)cdiv(tcs.rows_c * tcs.Kp, 256), 256, 0, stream>>>(Cf.p, tcs.Cc_hi.p, tcs.Cc_lo.p, K, 2 * L - 1, tcs.Kp, tcs.rows_c);
                post_launch();
            }
            tc::Params q = tc_base_params();
            q.groups = tcs.groups_c; q.nblocks = cdiv(tcs.Kp, tc::BK); q.own = tc::BN - (tcs.G - 1);
            q.units = cdiv(Tl, q.own);
            q.out = denH.p;
            const unsigned grid = (unsigned)std::min<int64_t>(q.units, tcs.num_sms);
            tc::tc_kernel<tc::TC_TRANS><<<grid, tc::THREADS, tc::SMEM_BYTES, stream>>>(tcs.mCu[0], tcs.mCu[1], tcs.mHk[0], tcs.mHk[1], q);
            post_launch();
        }
    }
    // numH via tensor cores (mult.jl:47)
    void tc_transconv() {
        if constexpr (std::is_same<S, float>::value) {
            if (fd_active()) { fd_transconv(); return; }
            tc_split_X();
            tc_split_W();
            tc::Params q = tc_base_params();
            q.groups = tcs.groups; q.nblocks = cdiv(N, tc::BK); q.own = tc::BN - (tcs.G - 1);
            q.units = cdiv(Tl, q.own);
            q.out = numH.p;
            const unsigned grid = (unsigned)std::min<int64_t>(q.units, tcs.num_sms);
            {
                const char *e = getenv("CMF_LOCKSTEP_TRANS");      // k-block window; 0 (default) disables
                const int w = e ? atoi(e) : 0;
                if (w > 0 && q.units >= (int64_t)grid) {
                    if (lockstep.n == 0) lockstep.alloc(1024);
                    CK(cudaMemsetAsync(lockstep.p, 0, lockstep.n * sizeof(int), stream));
                    q.lockstep = lockstep.p; q.lockstep_window = w;
                }
            }
            prof_begin(PROF_TRANSCONV);
            if (q.lockstep != nullptr) {
                void *args[] = {&tcs.mWu[0], &tcs.mWu[1], &tcs.mXk[0], &tcs.mXk[1], &q};
                CK(cudaLaunchCooperativeKernel((void *)tc::tc_kernel<tc::TC_TRANS>, dim3(grid), dim3(tc::THREADS), args, tc::SMEM_BYTES, stream));
            } else {
                tc::tc_kernel<tc::TC_TRANS><<<grid, tc::THREADS, tc::SMEM_BYTES, stream>>>(tcs.mWu[0], tcs.mWu[1], tcs.mXk[0], tcs.mXk[1], q);
            }
            prof_end();
            post_launch();
        }
    }
    // sum of squared residuals via tensor cores (mult.jl:55-57); returns the number of partials written
    int64_t tc_conv_loss() {
        if constexpr (std::is_same<S, float>::value) {
            tc_split_W();
            tc_split_H(false);
            tc::Params q = tc_base_params();
            q.tiles_n = cdiv(N, tc::BM);
            q.nkb = tcs.KLp / tc::BK;
            q.units = q.tiles_n * cdiv(Tl, tc::BN);
            q.X = X.p; q.partial = loss_part.p;
            const unsigned grid = (unsigned)std::min<int64_t>(q.units, tcs.num_sms);
            prof_begin(PROF_CONV);
            tc::tc_kernel<tc::TC_CONV><<<grid, tc::THREADS, tc::SMEM_BYTES, stream>>>(tcs.mWc[0], tcs.mWc[1], tcs.mHw[0], tcs.mHw[1], q);
            prof_end();
            post_launch();
            return tc::EPI_WARPS * q.units;
        }
        return 0;
    }

    ~Ctx() override {
        if (tracing) trace_dump();
        if (own_stream && stream) cudaStreamDestroy(stream);
    }

    void plan_split(int64_t Nin, int64_t tau_hi, int &nsplit, int64_t &split_len) {
        const int64_t ngrp = cdiv(L, 8);
        const int64_t bxy = cdiv(Nin, 128) * cdiv(K * ngrp, CORR_PB);
        int64_t ns = std::min<int64_t>(std::max<int64_t>(cdiv(592, bxy), 1), 64);
        split_len = cdiv(cdiv(tau_hi, ns), CORR_BTAU) * CORR_BTAU;
        if (split_len < CORR_BTAU) split_len = CORR_BTAU;
        nsplit = (int)cdiv(tau_hi, split_len);
    }

    // CMF_TRACE=1: an event after every launch; the destructor prints, per call site, the live stream time between consecutive
    // events (= the kernel plus whatever gap preceded it, at the clocks of the real run -- ncu's per-kernel times are not)
    struct TraceEv { cudaEvent_t e; const char *fn; int line; };
    std::vector<TraceEv> trace;
    const bool tracing = getenv("CMF_TRACE") && atoi(getenv("CMF_TRACE")) > 0;
    void post_launch(const char *fn = __builtin_FUNCTION(), int line = __builtin_LINE()) {
        ++launches;
        CK(cudaGetLastError());
        if (tracing) {
            TraceEv t; t.fn = fn; t.line = line;
            cudaEventCreate(&t.e);
            cudaEventRecord(t.e, stream);
            trace.push_back(t);
        }
    }
    void trace_dump() {
        if (trace.size() < 2) return;
        cudaStreamSynchronize(stream);
        struct Agg { double ms = 0.0; int64_t n = 0; };
        std::map<std::pair<std::string, int>, Agg> agg;
        const size_t skip = getenv("CMF_TRACE_SKIP") ? (size_t)atoll(getenv("CMF_TRACE_SKIP")) : 0;
        double total = 0.0;
        for (size_t i = std::max<size_t>(1, skip); i < trace.size(); ++i) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, trace[i - 1].e, trace[i].e) != cudaSuccess) continue;
            Agg &g = agg[{trace[i].fn, trace[i].line}];
            g.ms += ms; ++g.n; total += ms;
        }
        fprintf(stderr, "CMF_TRACE: %zu launches, %.3f ms between the first and last traced event (launches before #%zu skipped)\n", trace.size(), total, skip);
        for (auto &kv : agg)
            fprintf(stderr, "CMF_TRACE %-28s:%-5d n=%-6lld total %10.3f ms  mean %9.4f ms  %5.1f%%\n", kv.first.first.c_str(), kv.first.second,
                    (long long)kv.second.n, kv.second.ms, kv.second.ms / (double)kv.second.n, 100.0 * kv.second.ms / total);
        for (auto &t : trace) cudaEventDestroy(t.e);
        trace.clear();
    }

    // ---------------------------------------------------------------- kernel launch helpers
    // est over local columns [t_lo, t_hi) from (Wsrc, Hsrc) with dims (Kc, Lc).
    void launch_conv(const S *Wsrc, const S *Hsrc, int64_t Kc, int64_t Lc, int64_t h_lo, int64_t h_hi,
                     int64_t t_lo, int64_t t_hi, int flags, S *out, double *partial, uint64_t seed = 0,
                     double noise = 0.0) {
        if (t_hi <= t_lo) return;
        ConvArgs<S> a;
        a.Wi = Wsrc; a.H = Hsrc; a.X = X.p; a.out = out; a.partial = partial;
        a.N = N; a.K = Kc; a.L = Lc; a.t_lo = t_lo; a.t_hi = t_hi; a.h_lo = h_lo; a.h_hi = h_hi;
        a.flags = flags; a.seed = seed; a.t_global0 = t0; a.noise = noise;
        a.mask = pgd_mask.n ? pgd_mask.p : nullptr; a.lossf = pgd_loss;
        const int Lpad = (int)(cdiv(Lc, 8) * 8);
        const int HW = BT + Lpad - 1;
        const size_t ws_bytes = (size_t)CONV_LC * BN * sizeof(S);
        const size_t budget = 160 * 1024 - ws_bytes;
        int KC = (int)std::min<int64_t>(Kc, (int64_t)(budget / ((size_t)HW * sizeof(S))));
        REQUIRE(KC >= 1, "conv: L too large for the shared-memory H window");
        a.KC = KC;
        const size_t smem = (size_t)KC * HW * sizeof(S) + ws_bytes;
        auto kern = conv_kernel<S, TN, TT, TY>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((unsigned)cdiv(t_hi - t_lo, BT), (unsigned)cdiv(N, BN));
        prof_begin(PROF_CONV);
        kern<<<grid, dim3(16, TY), smem, stream>>>(a);
        prof_end();
        post_launch();
    }
    int conv_nblocks(int64_t t_lo, int64_t t_hi) const { return (int)(cdiv(N, BN) * cdiv(t_hi - t_lo, BT)); }

    void reduce_scalar(const double *part, int64_t n, double *dst) {
        reduce_sum_kernel<<<1, 1024, 0, stream>>>(part, n, dst);
        post_launch();
    }
    double fetch_scalar(const double *dst) {
        double v;
        CK(cudaMemcpyAsync(&v, dst, sizeof(double), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        return v;
    }

    // out[t][k] for t in [0, t_out)
    void launch_transconv(const S *Wg, const S *Xin, S *out, int64_t Nin, int64_t Kout, int64_t Lin,
                          int64_t ldx, int64_t t_out, int64_t x_cols) {
        if (t_out <= 0) return;
        // choose TK / kgroups minimising padding
        int best_tk = 8, best_kg = 1;
        double best_waste = 1e30;
        const int tks[3] = {8, 5, 4};
        for (int tk : tks) {
            int kg = 1;
            while (kg < 16 && kg * tk < Kout) kg *= 2;
            const int64_t kp = (int64_t)kg * tk;
            const double waste = (double)(cdiv(Kout, kp) * kp - Kout) / (double)Kout;
            if (waste < best_waste - 1e-9) { best_waste = waste; best_tk = tk; best_kg = kg; }
        }
        int tg = std::max(128 / best_kg, 8);
        while (tg * best_kg > 32 && cdiv(t_out, (int64_t)tg * 8) < 296) tg /= 2;
        TransArgs<S> a;
        a.Wg = Wg; a.Xin = Xin; a.out = out; a.Nin = Nin; a.Kout = Kout; a.Lin = Lin; a.ldx = ldx;
        a.t_out = t_out; a.x_cols = x_cols; a.kgroups = best_kg; a.tgroups = tg;
        const int BTt = tg * 8, Lpad = (int)(cdiv(Lin, 8) * 8);
        const int XW = BTt + Lpad, XWP = XW + XW / 8 + 1, KP = best_kg * best_tk;
        const size_t smem = ((size_t)TR_NC * XWP + (size_t)8 * TR_NC * KP) * sizeof(S);
        REQUIRE(smem <= 200 * 1024, "transconv: L too large for the shared-memory window");
        dim3 grid((unsigned)cdiv(t_out, BTt), (unsigned)cdiv(Kout, KP));
        const int nthr = best_kg * tg;
        // small problems (configs 1-2: T = 2000 gives 8 CTAs) do not fill the device over (t, k) tiles alone: split the sum over
        // units across blockIdx.z into partial outputs and add them in a fixed order (deterministic)
        int nsplit = 1;
        a.n_chunk = cdiv(Nin, TR_NC) * TR_NC; a.part_stride = 0;
        const int64_t ctas = (int64_t)grid.x * grid.y;
        if (ctas < 296 && Nin >= 64) {
            const int64_t want = std::min<int64_t>(std::min<int64_t>(cdiv(Nin, 32), cdiv(296, ctas)), 32);
            a.n_chunk = cdiv(cdiv(Nin, want), TR_NC) * TR_NC;
            nsplit = (int)cdiv(Nin, a.n_chunk);
            if (nsplit > 1) {
                a.part_stride = t_out * Kout;
                const size_t need = (size_t)nsplit * (size_t)a.part_stride;
                if (tr_part.n < need) tr_part.alloc(need);
                a.out = tr_part.p;
                grid.z = (unsigned)nsplit;
            }
        }
#define LAUNCH_TR(TKV)                                                                             \
    {                                                                                              \
        auto kern = transconv_kernel<S, TKV, 8>;                                                   \
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
        kern<<<grid, nthr, smem, stream>>>(a);                                                     \
    }
        if (Nin == N) prof_begin(PROF_TRANSCONV);
        if (best_tk == 8) LAUNCH_TR(8) else if (best_tk == 5) LAUNCH_TR(5) else LAUNCH_TR(4)
#undef LAUNCH_TR
        prof_end();
        post_launch();
        if (nsplit > 1) {
            sum_parts_kernel<S><<<(unsigned)cdiv(a.part_stride, 256), 256, 0, stream>>>(tr_part.p, out, a.part_stride, a.part_stride, nsplit);
            post_launch();
        }
    }
    DevBuf<S> tr_part;     // partial outputs of the N-split transposed convolution

    // out_S / out_D [(l*K+k)*Nin + n] = sum_tau hm(tau-l)[k] Xin[tau][n]
    void launch_corr(const S *Xin, int64_t Nin, int64_t ldx, int64_t tau_hi, int nsplit, int64_t split_len,
                     S *out_s, double *out_d) {
        CorrArgs<S> a;
        a.H = H; a.Xin = Xin; a.part = corr_part.p; a.Nin = Nin; a.K = K; a.L = L; a.ldx = ldx;
        a.u_hi = Tl; a.tau_hi = tau_hi; a.split_len = split_len;
        const int64_t ngrp = cdiv(L, 8);
        dim3 grid((unsigned)cdiv(Nin, 128), (unsigned)cdiv(K * ngrp, CORR_PB), (unsigned)nsplit);
        if (Nin == N) prof_begin(PROF_CORR);
        corr_kernel<S><<<grid, dim3(16, 16), 0, stream>>>(a);
        prof_end();
        post_launch();
        const int64_t n = KL() * Nin;
        reduce_partials_kernel<S><<<(unsigned)cdiv(n, 256), 256, 0, stream>>>(corr_part.p, nsplit, n, out_s, out_d);
        post_launch();
    }

    template <bool TB>
    void launch_gemm(const S *A, const S *B, S *C, int64_t M, int64_t Nn, int64_t Kg, int64_t lda, int64_t ldb,
                     int64_t ldc) {
        dim3 grid((unsigned)cdiv(Nn, 64), (unsigned)cdiv(M, 64));
        gemm_kernel<S, TB><<<grid, 256, 0, stream>>>(A, B, C, M, Nn, Kg, lda, ldb, ldc);
        post_launch();
    }

    // mult.jl:37-38 / 51-52.  dot_dst != nullptr additionally leaves <num, x'> (new x) there: the expansion loss's <numH, H'>
    // comes out of the H update for free.  Returns whether the inner product was produced.
    DevBuf<double> mu_part;
    bool launch_mu(S *x, const S *num, const S *den, double l1, double l2, int64_t n, double *dot_dst = nullptr) {
        if constexpr (std::is_same<S, float>::value) {
            const bool aligned = (((uintptr_t)x | (uintptr_t)num | (uintptr_t)den) & 15) == 0;
            if (aligned && n % 4 == 0 && n >= 4096) {
                const unsigned grid = (unsigned)std::min<int64_t>(cdiv(n / 4, 256), 148 * 8);
                if (dot_dst && mu_part.n == 0) mu_part.alloc(148 * 8);
                mu_update_vec4_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<float4 *>(x), reinterpret_cast<const float4 *>(num),
                                                                reinterpret_cast<const float4 *>(den), (float)l1, (float)l2, n / 4,
                                                                dot_dst ? mu_part.p : nullptr);
                post_launch();
                if (dot_dst) reduce_scalar(mu_part.p, grid, dot_dst);
                return dot_dst != nullptr;
            }
        }
        mu_update_kernel<S><<<(unsigned)cdiv(n, 256), 256, 0, stream>>>(x, num, den, (S)l1, (S)l2, n);
        post_launch();
        return false;
    }

    // ---------------------------------------------------------------- data
    void set_data(const void *Xh, int64_t first_col) override {
        REQUIRE(Xh != nullptr, "set_data: null pointer");
        REQUIRE(first_col <= t0, "set_data: host array starts after the shard's first column");
        const int64_t hi = std::min(t1 + (L - 1), T);
        CK(cudaMemsetAsync(X.p, 0, X.n * sizeof(S), stream));
        const S *src = static_cast<const S *>(Xh) + (t0 - first_col) * N;
        CK(cudaMemcpyAsync(X.p, src, (size_t)((hi - t0) * N) * sizeof(S), cudaMemcpyHostToDevice, stream));
        finish_data();
    }
    void get_data(void *X_out, int with_halo) override {
        const int64_t cols = with_halo ? std::min(t1 + (L - 1), T) - t0 : Tl;
        CK(cudaMemcpyAsync(X_out, X.p, (size_t)(cols * N) * sizeof(S), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
    }
    void finish_data() {
        tcs.x_dirty = true;
        fds.x_dirty = true;
        numH_valid = false; numH_dot_valid = false;
        calib_reset();
        data_sumsq_local = data_sumsq();
        data_norm = std::sqrt(data_sumsq_local);
        pgd_cur_loss = data_norm;
        have_data = true;
    }
    double data_sumsq() override {
        dot_partial_kernel<S><<<1024, 256, 0, stream>>>(X.p, X.p, Tl * N, loss_part.p);
        post_launch();
        reduce_scalar(loss_part.p, 1024, scal.p);
        return fetch_scalar(scal.p);
    }

    void synth_data(uint64_t seed, int64_t Kt, int64_t Lt, double p_h, double noise) override {
        REQUIRE(Kt >= 1 && Lt >= 1, "synth_data: bad ground-truth dims");
        DevBuf<S> Wt, Ht;
        Wt.alloc((size_t)(Kt * Lt * N));
        const int64_t cols_out = std::min(Tl + (L - 1), T - t0);   // local columns to produce
        const int64_t hcols = cols_out + (Lt - 1);
        Ht.alloc((size_t)(hcols * Kt));
        synth_W_kernel<S><<<(unsigned)cdiv(N, 128), 128, 0, stream>>>(Wt.p, Kt, N, Lt, seed, 0.1, 0.2);
        post_launch();
        synth_H_kernel<S><<<(unsigned)cdiv(hcols * Kt, 256), 256, 0, stream>>>(Ht.p, Kt, t0 - (Lt - 1), hcols, T, seed, p_h);
        post_launch();
        CK(cudaMemsetAsync(X.p, 0, X.n * sizeof(S), stream));
        launch_conv(Wt.p, Ht.p + (Lt - 1) * Kt, Kt, Lt, -(Lt - 1), cols_out, 0, cols_out, 8, X.p, nullptr, seed, noise);
        CK(cudaStreamSynchronize(stream));
        finish_data();
    }

    // ---------------------------------------------------------------- factors
    void set_factors(const void *Wh, const void *Hh, int64_t first_col) override {
        REQUIRE(Wh && Hh, "set_factors: null pointer");
        const int64_t lo = std::max<int64_t>(t0 - (L - 1), 0), hi = std::min(t1 + (L - 1), T);
        REQUIRE(first_col <= lo, "set_factors: host H starts after the shard's left halo");
        CK(cudaMemcpyAsync(Wtmp.p, Wh, Wi.n * sizeof(S), cudaMemcpyHostToDevice, stream));
        w_julia_to_internal<S><<<(unsigned)cdiv(KL() * N, 256), 256, 0, stream>>>(Wtmp.p, Wi.p, K, N, L);
        post_launch();
        CK(cudaMemsetAsync(Hbuf.p, 0, Hbuf.n * sizeof(S), stream));
        const S *src = static_cast<const S *>(Hh) + (lo - first_col) * K;
        CK(cudaMemcpyAsync(H + (lo - t0) * K, src, (size_t)((hi - lo) * K) * sizeof(S), cudaMemcpyHostToDevice, stream));
        CK(cudaStreamSynchronize(stream));
        have_factors = true;
        mark_w_dirty();
        numH_valid = false; numH_dot_valid = false;
        gram_valid = false; fds.h_dirty = true; fds.hf2_valid = false;
        pgd_stepW = pgd_stepH = 5.0;            // a new rule instance (pgd.jl:147-149)
        pgd_cur_loss = data_norm;
        calib_reset();
    }

    void init_rand(uint64_t seed) override {
        // W in Julia order (k + K*(n + N*l)) so the draw for element (k,n,l) does not depend on layout
        uniform_kernel<S><<<(unsigned)cdiv(KL() * N, 256), 256, 0, stream>>>(Wtmp.p, KL() * N, seed, 100, 0);
        post_launch();
        w_julia_to_internal<S><<<(unsigned)cdiv(KL() * N, 256), 256, 0, stream>>>(Wtmp.p, Wi.p, K, N, L);
        post_launch();
        const int64_t lo = std::max<int64_t>(t0 - (L - 1), 0), hi = std::min(t1 + (L - 1), T);
        CK(cudaMemsetAsync(Hbuf.p, 0, Hbuf.n * sizeof(S), stream));
        uniform_kernel<S><<<(unsigned)cdiv((hi - lo) * K, 256), 256, 0, stream>>>(H + (lo - t0) * K, (hi - lo) * K, seed, 101, lo * K);
        post_launch();
        CK(cudaStreamSynchronize(stream));
        have_factors = true;
        mark_w_dirty();
        numH_valid = false; numH_dot_valid = false;
        gram_valid = false; fds.h_dirty = true; fds.hf2_valid = false;
        calib_reset();
    }

    void init_scale_partials(double out[2]) override {
        REQUIRE(have_data && have_factors, "init_scale_partials: data and factors must be set");
        const int nb = conv_nblocks(0, Tl);
        launch_conv(Wi.p, H, K, L, -(L - 1), Tl + (L - 1), 0, Tl, 16, nullptr, loss_part.p);
        // partial layout: [2*b] = <est, X>, [2*b+1] = ||est||^2  -> de-interleave by strided sums
        std::vector<double> hp((size_t)2 * nb);
        CK(cudaMemcpyAsync(hp.data(), loss_part.p, hp.size() * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        double a = 0.0, b = 0.0;
        for (int i = 0; i < nb; ++i) { a += hp[2 * i]; b += hp[2 * i + 1]; }
        out[0] = a; out[1] = b;
    }

    void scale_factors(double s) override {
        scale_kernel<S><<<(unsigned)cdiv(KL() * N, 256), 256, 0, stream>>>(Wi.p, (S)s, KL() * N);
        post_launch();
        scale_kernel<S><<<(unsigned)cdiv((int64_t)Hbuf.n, 256), 256, 0, stream>>>(Hbuf.p, (S)s, (int64_t)Hbuf.n);
        post_launch();
        mark_w_dirty();
        numH_valid = false; numH_dot_valid = false;
        gram_valid = false; fds.h_dirty = true; fds.hf2_valid = false;
        calib_reset();
    }

    void get_factors(void *Wo, void *Ho) override {
        if (Wo) {
            w_internal_to_julia<S><<<(unsigned)cdiv(KL() * N, 256), 256, 0, stream>>>(Wi.p, Wtmp.p, K, N, L);
            post_launch();
            CK(cudaMemcpyAsync(Wo, Wtmp.p, Wi.n * sizeof(S), cudaMemcpyDeviceToHost, stream));
        }
        if (Ho) CK(cudaMemcpyAsync(Ho, H, (size_t)(Tl * K) * sizeof(S), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
    }

    // ---------------------------------------------------------------- MU (src/algs/mult.jl)
    void w_partials() override {
        REQUIRE(have_data && have_factors, "update: data and factors must be set first");
        // numW partial over the owned u (mult.jl:32): Xin = X incl. right halo
        if (tc_active()) tc_corr();
        else launch_corr(X.p, N, N, Tl + (L - 1), nsplit_w, split_w, numW.p, nullptr);
        // Gram partial Rg[d][k][k'] = sum_u H[u][k] H[u+d][k'] (owned u, right halo for u+d); already resident
        // when the expansion loss of the previous iteration computed it for the same H
        if (!gram_valid) gram_partial();
        gram_valid = false;     // the host all-reduces the buffer in place: it is consumed by this step
    }

    void gram_partial() {
        if (fd_active()) fd_gram();
        else if (tc_active()) { tc_split_H(true); tc_split_H(false); tc_gram(); }
        else launch_corr(H, K, K, Tl + (L - 1), nsplit_g, split_g, nullptr, exch1.p);
        if (L > 1) {
            h_tail_kernel<S><<<(unsigned)cdiv((L - 1) * K, 256), 256, 0, stream>>>(H, exch1.p + L * K * K, K, L, Tl, is_last ? 1 : 0);
            post_launch();
        }
    }

    void build_G() {
        build_G_kernel<S><<<(unsigned)cdiv(KL() * KL(), 256), 256, 0, stream>>>(exch1.p, exch1.p + L * K * K, GS.p, K, L);
        post_launch();
    }

    void w_apply(double l1W, double l2W) override {
        if (alg == CMF_HALS) { hals_w_apply(l1W, l2W); return; }
        REQUIRE(alg == CMF_MULT, "the split-phase W update serves MultUpdate and HALSUpdate");
        w_denom_rows(0, KL());
        w_update_rows(l1W, l2W, 0, KL());
    }
    // denomW = G * Wi (mult.jl:28,33) for the unfolded rows [j0, j1)
    void w_denom_rows(int64_t j0, int64_t j1) override {
        if (tc_active()) {
            tc_denomW(j0, j1);                                                    // on tensor cores
        } else {
            build_G();
            launch_gemm<false>(GS.p + j0 * KL(), Wi.p, denW.p + j0 * N, j1 - j0, N, KL(), KL(), N, N);
        }
    }
    // mult.jl:37-38 on the rows [j0, j1) (the whole update when the range is everything)
    void w_update_rows(double l1W, double l2W, int64_t j0, int64_t j1) override {
        launch_mu(Wi.p + j0 * N, numW.p + j0 * N, denW.p + j0 * N, l1W, l2W, (j1 - j0) * N);
        mark_w_dirty();
        numH_valid = false; numH_dot_valid = false;
    }
    void w_rows_buffers(void **numw, void **w, int64_t *row_elems) override { *numw = numW.p; *w = Wi.p; *row_elems = N; }

    // layout of the W W' product held in GS: S2[(l*s2_ks + k)*s2_ld + l'*s2_ks + k']
    int64_t s2_ks = 0, s2_ld = 0;

    void lag_tables() {
        if (tc_active()) {
            tc_split_W();
            tc_plain(tcs.mWu, tcs.mWuB, GS.p, tcs.rows_u, tcs.rows_u, N, tcs.rows_u, true);   // S2 = Wu Wu' on tensor cores (lower tiles + mirror)
            s2_ks = tcs.Kp; s2_ld = tcs.rows_u;
        } else {
            launch_gemm<true>(Wi.p, Wi.p, GS.p, KL(), KL(), N, N, N, KL());              // S2 = Wi Wi'
            s2_ks = K; s2_ld = KL();
        }
        lag_table_kernel<S><<<(unsigned)cdiv((2 * L - 1) * K * K, 256), 256, 0, stream>>>(GS.p, Cf.p, K, L, s2_ks, s2_ld);
        post_launch();
    }

    // truncated tail of denomH (the last L-1 columns, mult.jl:44,48 with the conv cut at T) from S2 = W W' in GS
    DevBuf<double> tail_part;
    void launch_denomH_tail() {
        const int nb = (int)cdiv((2 * L - 1) * K, 256);
        const size_t smem = (size_t)8 * (size_t)(L - 1) * sizeof(double);
        if (getenv("CMF_TAIL_DIRECT") || smem > 48 * 1024) {          // literal form (kept for A/B and very long lags)
            dim3 grid((unsigned)(L - 1), (unsigned)K);
            denomH_tail_kernel<S><<<grid, 256, 0, stream>>>(GS.p, H, denH.p, K, L, Tl, -(L - 1), s2_ks, s2_ld);
            post_launch();
            return;
        }
        const size_t need = (size_t)nb * (size_t)(L - 1) * (size_t)K;
        if (tail_part.n < need) tail_part.alloc(need);
        denomH_tail_prefix_kernel<S><<<dim3((unsigned)nb, (unsigned)K), 256, smem, stream>>>(GS.p, H, tail_part.p, K, L, Tl, -(L - 1), s2_ks, s2_ld);
        post_launch();
        denomH_tail_reduce_kernel<S><<<(unsigned)cdiv((L - 1) * K, 256), 256, 0, stream>>>(tail_part.p, denH.p, K, L, Tl, nb);
        post_launch();
    }

    void h_update(double l1H, double l2H) override {
        REQUIRE(have_data && have_factors, "update: data and factors must be set first");
        if (tc_active()) tc_transconv();
        else launch_transconv(Wi.p, X.p, numH.p, N, K, L, N, Tl, Tl + (L - 1));    // numH (mult.jl:47)
        lag_tables();
        // denomH = C (*) H on all owned columns (mult.jl:44,48), then the truncated tail
        if (fd_active()) fd_denomH();
        else if (tc_active()) { tc_split_H(false); tc_denomH(); }
        else launch_transconv(Cf.p, Hbuf.p, denH.p, K, K, 2 * L - 1, K, Tl, Tl + 2 * (L - 1));
        if (is_last && L > 1) {
            launch_denomH_tail();
        }
        // mult.jl:51-52; with the expansion loss in force the update also leaves <numH, H'> in scal[4]
        numH_dot_valid = launch_mu(H, numH.p, denH.p, l1H, l2H, Tl * K, (loss_mode == 1 && alg == CMF_MULT) ? scal.p + 4 : nullptr);
        numH_valid = (alg == CMF_MULT);
        gram_valid = false; fds.h_dirty = true; fds.hf2_valid = false;
    }

    double *scalars() override { return scal.p; }

    // Leaves the local loss scalars in scal.p without a host synchronisation and returns how many there are:
    //   1 (direct pass):  scal[0] = sum of squared residuals over the owned columns (mult.jl:55-57)
    //   2 (expansion):    scal[0] = <numH, H>, scal[1] = <W W', Htilde Htilde'> over the owned columns
    // scal[1] is zeroed in the direct case so that a fixed-size all-reduce of two doubles serves both.
    int loss_partial_dev() override {
        REQUIRE(have_data && have_factors, "loss: data and factors must be set first");
        if (loss_mode == 1 && tc_active()) {
            // frequency-domain engine: numH and W W' are cheap, so the expansion also serves calls that find them stale
            // (the loss at the initial factors, after a W-only step, after the HALS sweep overwrote numH)
            if (!numH_valid && fd_active()) { tc_transconv(); lag_tables(); numH_valid = true; numH_dot_valid = false; }
            if (numH_valid) { loss_partial_expansion_dev(); return 2; }
        }
        int64_t nb = conv_nblocks(0, Tl);
        if (fd_active() && !getenv("CMF_FD_LOSS_TC")) nb = fd_conv_loss();
        else if (tc_active()) nb = tc_conv_loss();
        else launch_conv(Wi.p, H, K, L, -(L - 1), Tl + (L - 1), 0, Tl, 4, nullptr, loss_part.p);  // mult.jl:55-57
        reduce_scalar(loss_part.p, nb, scal.p);
        CK(cudaMemsetAsync(scal.p + 1, 0, sizeof(double), stream));
        return 1;
    }

    // local sum of squared residuals (for the expansion: this shard's share of the global identity)
    double loss_partial() override {
        const int n = loss_partial_dev();
        double bc[2];
        CK(cudaMemcpyAsync(bc, scal.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        return n == 2 ? data_sumsq_local - 2.0 * bc[0] + bc[1] : bc[0];
    }

    // ||conv(W,H) - X||^2 = ||X||^2 - 2 <transconv(W,X), H> + <W W', Htilde Htilde'>, summed over the owned columns
    // (the global sum over shards is exact; a single shard's value is not its own residual).  numH and W W' are
    // resident from the H update with the current W; the Gram of the new H is computed here and reused by the
    // next cmf_w_partials (the halos must not change in between).
    void loss_partial_expansion_dev() {
        gram_partial();
        gram_valid = true;
        if (getenv("CMF_G_DIRECT"))
            s2_dot_G_kernel<S><<<1024, 256, 0, stream>>>(GS.p, s2_ks, s2_ld, exch1.p, exch1.p + L * K * K, K, L, loss_part.p + 1024);
        else
            s2_dot_G_diag_kernel<S><<<1024, 256, 0, stream>>>(GS.p, s2_ks, s2_ld, exch1.p, exch1.p + L * K * K, K, L, loss_part.p + 1024);
        post_launch();
        reduce_scalar(loss_part.p + 1024, 1024, scal.p + 1);
        if (numH_dot_valid) {             // produced by the H update itself (consumed once: anything may happen before the next call)
            CK(cudaMemcpyAsync(scal.p, scal.p + 4, sizeof(double), cudaMemcpyDeviceToDevice, stream));
            numH_dot_valid = false;
            return;
        }
        dot_partial_kernel<S><<<1024, 256, 0, stream>>>(numH.p, H, Tl * K, loss_part.p);
        post_launch();
        reduce_scalar(loss_part.p, 1024, scal.p);
    }

    // ---------------------------------------------------------------- HALS (src/algs/hals.jl)
    // The HALS gradients come from the same four quantities as MU (no residual matrix is kept):
    //   P = R Htilde' = denomW - numW   (hals.jl:111 projections for every column at once)
    //   Q = transconv(W, R) = denomH - numH
    // with R = conv(W,H) - X (hals.jl:22).  w_partials() / the Gram all-reduce are shared with MU.
    void hals_w_apply(double l1W, double l2W) {
        if (tc_active()) tc_denomW();
        build_G();                                                            // G as a matrix for the sweep (GS)
        if (!tc_active()) launch_gemm<false>(GS.p, Wi.p, denW.p, KL(), N, KL(), KL(), N, N);
        sub_kernel<S><<<(unsigned)cdiv(KL() * N, 256), 256, 0, stream>>>(numW.p, denW.p, numW.p, KL() * N);   // P
        post_launch();
        const size_t smem = (size_t)KL() * sizeof(S);
        REQUIRE(smem <= 200 * 1024, "HALS W sweep: K*L too large for shared memory");
        auto kern = hals_w_sweep_kernel<S>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)N, 256, smem, stream>>>(GS.p, numW.p, Wi.p, K, L, N, (S)l1W, (S)l2W);   // hals.jl:90-112
        post_launch();
        mark_w_dirty();
        numH_valid = false; numH_dot_valid = false;
    }

    void hals_update_motifs(double l1W, double l2W) override {
        w_partials();
        hals_w_apply(l1W, l2W);
    }

    // Q = transconv(W, R) = denomH - numH on the owned columns (left in numH)
    void hals_h_prepare() override {
        REQUIRE(have_data && have_factors, "update: data and factors must be set first");
        if (tc_active()) tc_transconv();
        else launch_transconv(Wi.p, X.p, numH.p, N, K, L, N, Tl, Tl + (L - 1));
        lag_tables();
        if (fd_active()) fd_denomH();
        else if (tc_active()) { tc_split_H(false); tc_denomH(); }
        else launch_transconv(Cf.p, Hbuf.p, denH.p, K, K, 2 * L - 1, K, Tl, Tl + 2 * (L - 1));
        if (is_last && L > 1) {
            launch_denomH_tail();
        }
        sub_kernel<S><<<(unsigned)cdiv(Tl * K, 256), 256, 0, stream>>>(numH.p, denH.p, numH.p, Tl * K);        // Q
        post_launch();
    }
    void hals_h_local(void **q, void **h, int64_t *count) override { *q = numH.p; *h = H; *count = Tl * K; }
    // rank 0 of a T-sharded HALS fit holds Q, H and Delta H of ALL columns during the sweep: the sweep is a chain of T
    // dependent steps per component (hals.jl:124-128) that more GPUs cannot shorten, so the shards' Q are gathered here,
    // swept once in the reference's order, and the new H is scattered back
    DevBuf<S> Qfull, Hfull, Dfull;
    void hals_h_full(void **q, void **h) override {
        const size_t n = (size_t)(T * K);
        if (Qfull.n < n) { Qfull.alloc(n); Hfull.alloc(n); Dfull.alloc(n); }
        *q = Qfull.p; *h = Hfull.p;
    }

    // second-generation sweep (kernels_hals.cuh): rounds with a grid barrier, lane = component recurrences, 8 x 8 pull blocks.
    // Returns false when the handle / shape is outside its envelope (the wavefront kernel below then runs).
    DevBuf<float> h2_Hcm, h2_AD, h2_part, h2_coef;
    DevBuf<unsigned> h2_bar;
    int h2_max_ctas = -1;
    bool hals2_sweep(const S *Qp, S *Hp, int64_t Tt, const S *ct, double l1H, double l2H) {
        if constexpr (!std::is_same<S, float>::value) { return false; } else {
            if (L > hals2::LMAX || K > 128 || (L > 1 && ct == nullptr)) return false;
            if (const char *e = getenv("CMF_HALS_SWEEP")) { if (atoi(e) == 1) return false; }      // 1: first-generation wavefront kernel
            const int G = (int)cdiv(K, hals2::GS);
            const int n_block = G * (G - 1) / 2, n_diag = G, n_rec = (int)cdiv(K, 32);
            const int grid = n_block + n_diag + n_rec;
            const size_t smem = hals2::smem_bytes();
            if (h2_max_ctas < 0) {
                int per_sm = 0, sms = 0;
                CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
                cudaError_t e1 = cudaFuncSetAttribute(hals2::hals2_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                cudaError_t e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hals2::hals2_sweep_kernel, hals2::NT, smem);
                h2_max_ctas = (e1 == cudaSuccess && e2 == cudaSuccess) ? per_sm * sms : 0;
                if (e1 != cudaSuccess || e2 != cudaSuccess) cudaGetLastError();
            }
            if (grid > h2_max_ctas) return false;
            const int64_t nC = cdiv(Tt, hals2::CW), Tp = nC * hals2::CW;
            const size_t ncm = (size_t)hals2::cells_elems(K, nC), npart = (size_t)G * (size_t)K * hals2::RING * hals2::CW;
            if (h2_Hcm.n < ncm) { h2_Hcm.alloc(ncm); h2_AD.alloc(ncm); }
            if (h2_part.n < npart) h2_part.alloc(npart);
            dim3 tb(32, 8), tg((unsigned)cdiv(Tp, 32), (unsigned)cdiv(K, 32));
            hals2::hals2_prepare_kernel<<<tg, tb, 0, stream>>>(Qp, Hp, h2_AD.p, h2_Hcm.p, K, Tt, Tp);
            post_launch();
            if (h2_coef.n < (size_t)(K * 64)) h2_coef.alloc((size_t)(K * 64));
            hals2::hals2_coef_kernel<<<(unsigned)cdiv(K * 64, 256), 256, 0, stream>>>(Cf.p, h2_coef.p, K, L, (float)l2H);
            post_launch();
            if (h2_bar.n == 0) h2_bar.alloc(32);
            CK(cudaMemsetAsync(h2_bar.p, 0, sizeof(unsigned), stream));
            hals2::Args a;
            a.barrier = h2_bar.p;
            a.coef = h2_coef.p;
            a.Cf = Cf.p; a.Ct = ct; a.S2 = GS.p; a.Ks = s2_ks; a.ld = s2_ld;
            a.H_cm = h2_Hcm.p; a.AD_cm = h2_AD.p; a.part = h2_part.p;
            a.K = K; a.L = L; a.T = Tt; a.nC = nC;
            a.l1 = (float)l1H; a.l2 = (float)l2H;
            a.G = G; a.n_block_items = n_block; a.n_diag_items = n_diag; a.n_rec = n_rec;
            DevBuf<long long> dbgbuf;
            const bool dbg2 = getenv("CMF_HALS_DEBUG") && atoi(getenv("CMF_HALS_DEBUG")) == 1;
            if (dbg2) dbgbuf.alloc((size_t)12 * grid);
            a.dbg = dbg2 ? dbgbuf.p : nullptr;
            a.dbg_mode = getenv("CMF_HALS_DBGMODE") ? atoi(getenv("CMF_HALS_DBGMODE")) : 0;
            void *args[] = {&a};
            prof_begin(PROF_SWEEP);
            CK(cudaLaunchCooperativeKernel((void *)hals2::hals2_sweep_kernel, dim3((unsigned)grid), dim3(hals2::NT), args, smem, stream));
            prof_end();
            post_launch();
            hals2::hals2_finish_kernel<<<dim3((unsigned)cdiv(Tt, 32), (unsigned)cdiv(K, 32)), tb, 0, stream>>>(h2_Hcm.p, Hp, K, Tt, Tp);
            post_launch();
            if (dbg2) {      // kcycles of work per active round (barrier waits excluded): worst and mean CTA of every role
                std::vector<long long> hv((size_t)12 * grid);
                CK(cudaMemcpyAsync(hv.data(), dbgbuf.p, hv.size() * sizeof(long long), cudaMemcpyDeviceToHost, stream));
                CK(cudaStreamSynchronize(stream));
                fprintf(stderr, "hals2: SM clock over the launch %lld MHz (clock64 / globaltimer, CTA 0)\n", hv[7]);
                const char *names[3] = {"block items", "diagonal items", "recurrence warps"};
                const int lo[4] = {0, n_block, n_block + n_diag, grid};
                for (int r = 0; r < 3; ++r) {
                    double worst = 0.0, mean = 0.0; int cnt = 0, wi = -1;
                    double phw[5] = {0, 0, 0, 0, 0};
                    long long mx = 0, slow = 0, rds = 0;
                    for (int i = lo[r]; i < lo[r + 1]; ++i) {
                        if (hv[12 * i + 1] == 0) continue;
                        const double v = (double)hv[12 * i] / (double)hv[12 * i + 1] / 1e3;
                        if (v > worst) { worst = v; wi = i - lo[r]; for (int q = 0; q < 5; ++q) phw[q] = (double)hv[12 * i + 2 + q] / (double)hv[12 * i + 1] / 1e3; }
                        mean += v; ++cnt;
                        mx = std::max(mx, hv[12 * i + 8]); slow += hv[12 * i + 9]; rds += hv[12 * i + 1];
                    }
                    fprintf(stderr, "hals2 %s: slowest single round %.1f kcycles, %.2f%% of the (CTA, round) pairs above 90 kcycles\n", names[r], mx / 1e3, rds ? 100.0 * slow / rds : 0.0);
                    fprintf(stderr, "hals2 %s: %d CTAs, kcycles of work per active round: mean %.1f, worst %.1f (item %d; phases stage %.1f pull %.1f tail/store %.1f final %.1f, barrier wait %.1f); rounds %lld, chunks %lld\n",
                            names[r], lo[r + 1] - lo[r], cnt ? mean / cnt : 0.0, worst, wi, phw[0], phw[1], phw[2], phw[3], phw[4], (long long)(nC + 4 * (K - 1) + 3), (long long)nC);
                }
            }
            return true;
        }
    }

    // wavefront sweep (hals.jl:121-154) over the local shard (single-shard handles) or over the gathered full-T buffers
    void hals_h_sweep(int full, double l1H, double l2H) override {
        REQUIRE(full || (is_first && is_last), "the local HALS H sweep needs a single-shard handle");
        if (progress.n == 0) hals_setup();
        CK(cudaMemsetAsync(progress.p, 0, progress.n * sizeof(int), stream));
        const size_t smem = hals_wave_smem_elems<S>(L) * sizeof(S);
        const S *cf = Cf.p, *s2 = GS.p, *q = full ? Qfull.p : numH.p;
        S *hh = full ? Hfull.p : H, *dd = full ? Dfull.p : denH.p, *tc_ = tailC.p;
        int *pr = progress.p;
        int64_t Kk = K, Ll = L, Tt = full ? T : Tl, ks = s2_ks, ldv = s2_ld;
        S a1 = (S)l1H, a2 = (S)l2H;
        int dbg = getenv("CMF_HALS_DEBUG") ? atoi(getenv("CMF_HALS_DEBUG")) : 0;   // 1: phase clocks of the last component; 2 / 4: timing runs without the pull / recurrence
        // truncated lag tables of all component pairs for the pull over the last L-1 columns (skipped when they would not fit: the
        // kernel then forms them on the fly)
        const S *ct = nullptr;
        {
            const size_t need = (size_t)(L - 1) * (size_t)(2 * L - 1) * (size_t)(K * K);
            if (L > 1 && need * sizeof(S) <= ((size_t)1 << 30) && !getenv("CMF_HALS_TAIL_DIRECT")) {
                if (tailCt.n < need) tailCt.alloc(need);
                hals_tail_table_kernel<S><<<(unsigned)cdiv((2 * L - 1) * K * K, 256), 256, 0, stream>>>(GS.p, tailCt.p, K, L, s2_ks, s2_ld);
                post_launch();
                ct = tailCt.p;
            }
        }
        if (hals2_sweep(q, hh, Tt, ct, l1H, l2H)) return;     // round-based kernel (fp32, L <= 32, K <= 128)
        void *args[] = {&cf, &s2, &q, &hh, &dd, &tc_, &pr, &Kk, &Ll, &Tt, &ks, &ldv, &a1, &a2, &dbg, &ct};
        prof_begin(PROF_SWEEP);
        CK(cudaLaunchCooperativeKernel((void *)hals_h_wave_kernel<S>, dim3((unsigned)hals_grid), dim3(HW_NT), args, smem, stream));
        prof_end();
        post_launch();
    }
    void h_changed() override {
        gram_valid = false; fds.h_dirty = true; fds.hf2_valid = false;
        numH_valid = false; numH_dot_valid = false;
    }

    double hals_update_feature_maps(double l1H, double l2H) override {
        REQUIRE(is_first && is_last, "single-shard entry; sharded fits go through the communicator path");
        hals_h_prepare();
        hals_h_sweep(0, l1H, l2H);
        h_changed();
        return loss_partial();                                                // hals.jl:41
    }

    // ---------------------------------------------------------------- PGD (src/algs/pgd.jl, SquareLoss)
    // The gradients are the MU quantities again: dW = 2 (denomW - numW), dH = 2 (denomH - numH) (pgd.jl:206-221).
    int pgd_constrW = 0, pgd_constrH = 0;     // 0 NonnegConstraint (pgd.jl:91-95), 1 UnitNormConstraint (:98-110)
    DevBuf<double> un_part;
    // `inner` = elements between consecutive components in memory (N for Wi rows (l, k), 1 for H[t][K])
    void pgd_step(S *x, S *g, int64_t n, double &step, int constr, int64_t inner) {
        dot_partial_kernel<S><<<1024, 256, 0, stream>>>(g, g, n, loss_part.p);
        post_launch();
        reduce_scalar(loss_part.p, 1024, scal.p + 2);                                   // ||grad||^2 stays on the device
        pgd_step_kernel<S><<<(unsigned)cdiv(n, 256), 256, 0, stream>>>(x, g, step, scal.p + 2, n, constr == 0 ? 1 : 0);   // pgd.jl:236-241
        post_launch();
        if (constr == 1) {
            if (un_part.n == 0) un_part.alloc((size_t)K * UN_CHUNKS);
            unit_norm_partial_kernel<S><<<dim3(UN_CHUNKS, (unsigned)K), 256, 0, stream>>>(x, un_part.p, K, inner, n / K);
            post_launch();
            unit_norm_apply_kernel<S><<<(unsigned)cdiv(n, 256), 256, 0, stream>>>(x, un_part.p, K, inner, n);
            post_launch();
        }
    }
    void pgd_adapt(double loss, double &step) {
        step *= (loss < pgd_cur_loss) ? 1.05 : 0.70;                                    // pgd.jl:247-251
        pgd_cur_loss = loss;
    }

    // ---- PGD with the pluggable losses of pgd.jl:28-70 (AbsoluteLoss, MaskedLoss): the loss gradient d D / d est is not
    // linear in (X, est) any more (sign) or carries a mask, so the Gram forms do not apply and the path follows pgd.jl:224-255
    // literally: conv + loss-gradient epilogue into an N x T scratch, then the W-side correlation (pgd.jl:206-214) resp. the
    // transposed conv (:218-221) of that scratch, on the SIMT kernels in the handle's type.
    int pgd_loss = 0;                 // 0 SquareLoss, 1 AbsoluteLoss
    DevBuf<S> pgd_mask, pgd_ge;       // mask [t][N] (empty = none), loss-gradient scratch [t][N] incl. a zero right halo
    bool pgd_general() const { return pgd_loss != 0 || pgd_mask.n != 0; }
    void set_pgd_constraints(int cw, int ch) override {
        REQUIRE(alg == CMF_PGD, "cmf_set_pgd_constraints: PGD handles only");
        REQUIRE((cw == 0 || cw == 1) && (ch == 0 || ch == 1), "constraint must be 0 (NonnegConstraint) or 1 (UnitNormConstraint)");
        pgd_constrW = cw; pgd_constrH = ch;
    }
    void set_pgd_loss(int lossf, const void *mask_host) override {
        REQUIRE(alg == CMF_PGD, "the pluggable loss belongs to PGDUpdate handles");
        REQUIRE(lossf == 0 || lossf == 1, "loss_func must be 0 (SquareLoss) or 1 (AbsoluteLoss)");
        pgd_loss = lossf;
        if (mask_host) {
            pgd_mask.alloc((size_t)(N * Tl));
            CK(cudaMemcpyAsync(pgd_mask.p, mask_host, pgd_mask.n * sizeof(S), cudaMemcpyHostToDevice, stream));
            CK(cudaStreamSynchronize(stream));
        } else pgd_mask.free();
        if (pgd_general() && pgd_ge.n == 0) pgd_ge.alloc((size_t)((Tl + (L - 1)) * N));
    }
    void pgd_loss_gradient() {       // pgd.jl:231: grad!(loss_func, est, est, data)
        launch_conv(Wi.p, H, K, L, -(L - 1), Tl + (L - 1), 0, Tl, 32, pgd_ge.p, nullptr);
    }
    double pgd_loss_eval() {         // pgd.jl:244-245: tensor_conv!(est, W, H); eval(loss_func, data, est)
        const int nb = conv_nblocks(0, Tl);
        launch_conv(Wi.p, H, K, L, -(L - 1), Tl + (L - 1), 0, Tl, 64, nullptr, loss_part.p);
        reduce_scalar(loss_part.p, nb, scal.p);
        return fetch_scalar(scal.p);
    }

    void pgd_update_motifs(double l1W, double l2W) override {
        REQUIRE(is_first && is_last, "PGD is single-shard");
        if (pgd_general()) {
            pgd_loss_gradient();
            launch_corr(pgd_ge.p, N, N, Tl + (L - 1), nsplit_w, split_w, denW.p, nullptr);          // gradW (pgd.jl:206-214)
            pgd_penalty_kernel<S><<<(unsigned)cdiv(KL() * N, 256), 256, 0, stream>>>(denW.p, Wi.p, (S)l1W, (S)l2W, KL() * N);
            post_launch();
            pgd_step(Wi.p, denW.p, KL() * N, pgd_stepW, pgd_constrW, N);
            mark_w_dirty();
            numH_valid = false; numH_dot_valid = false;
            pgd_adapt(pgd_loss_eval(), pgd_stepW);
            return;
        }
        w_partials();
        if (tc_active()) tc_denomW();
        else { build_G(); launch_gemm<false>(GS.p, Wi.p, denW.p, KL(), N, KL(), KL(), N, N); }
        pgd_grad_kernel<S><<<(unsigned)cdiv(KL() * N, 256), 256, 0, stream>>>(denW.p, denW.p, numW.p, Wi.p, (S)l1W, (S)l2W, KL() * N);
        post_launch();
        pgd_step(Wi.p, denW.p, KL() * N, pgd_stepW, pgd_constrW, N);
        mark_w_dirty();
        numH_valid = false; numH_dot_valid = false;
        pgd_adapt(loss_partial(), pgd_stepW);                                           // pgd.jl:244-252
    }

    double pgd_update_feature_maps(double l1H, double l2H) override {
        REQUIRE(is_first && is_last, "PGD is single-shard");
        if (pgd_general()) {
            pgd_loss_gradient();
            launch_transconv(Wi.p, pgd_ge.p, denH.p, N, K, L, N, Tl, Tl + (L - 1));                   // gradH (pgd.jl:218-221)
            pgd_penalty_kernel<S><<<(unsigned)cdiv(Tl * K, 256), 256, 0, stream>>>(denH.p, H, (S)l1H, (S)l2H, Tl * K);
            post_launch();
            pgd_step(H, denH.p, Tl * K, pgd_stepH, pgd_constrH, 1);
            gram_valid = false; fds.h_dirty = true; fds.hf2_valid = false;
            numH_valid = false; numH_dot_valid = false;
            pgd_adapt(pgd_loss_eval(), pgd_stepH);
            return pgd_cur_loss;
        }
        if (tc_active()) tc_transconv();
        else launch_transconv(Wi.p, X.p, numH.p, N, K, L, N, Tl, Tl + (L - 1));
        lag_tables();
        if (fd_active()) fd_denomH();
        else if (tc_active()) { tc_split_H(false); tc_denomH(); }
        else launch_transconv(Cf.p, Hbuf.p, denH.p, K, K, 2 * L - 1, K, Tl, Tl + 2 * (L - 1));
        if (L > 1) {
            launch_denomH_tail();
        }
        pgd_grad_kernel<S><<<(unsigned)cdiv(Tl * K, 256), 256, 0, stream>>>(denH.p, denH.p, numH.p, H, (S)l1H, (S)l2H, Tl * K);
        post_launch();
        pgd_step(H, denH.p, Tl * K, pgd_stepH, pgd_constrH, 1);
        gram_valid = false; fds.h_dirty = true; fds.hf2_valid = false;
        numH_valid = true;                                                              // W unchanged: the expansion loss may reuse numH / W W'
        pgd_adapt(loss_partial(), pgd_stepH);
        return pgd_cur_loss;                                                            // caller: sqrt(cur_loss / datanorm^2), pgd.jl:202
    }

    // ---------------------------------------------------------------- exchange
    void exchange_buffer(int which, void **p, int64_t *count, int *dt) override {
        if (which == 0) { *p = numW.p; *count = KL() * N; *dt = dtype; }
        else if (which == 1) { *p = exch1.p; *count = L * K * K + (L - 1) * K; *dt = CMF_F64; }
        else if (which == 2) { *p = numH.p; *count = Tl * K; *dt = dtype; }   // diagnostics: numH [t][K]
        else if (which == 3) { *p = denH.p; *count = Tl * K; *dt = dtype; }   // diagnostics: denomH [t][K]
        else throw CmfError(CMF_ERR_ARG, "exchange_buffer: which must be 0..3");
    }
    void halo_buffers(void **sl, void **sr, void **rl, void **rr, int64_t *count) override {
        *sl = H; *sr = H + (Tl - (L - 1)) * K; *rl = Hbuf.p; *rr = H + Tl * K; *count = (L - 1) * K;
    }

    // ---------------------------------------------------------------- primitives
    void prim_conv(void *out_host) override {
        DevBuf<S> out;
        out.alloc((size_t)(N * Tl));
        launch_conv(Wi.p, H, K, L, -(L - 1), Tl + (L - 1), 0, Tl, 1, out.p, nullptr);
        CK(cudaMemcpyAsync(out_host, out.p, out.n * sizeof(S), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
    }
    void prim_transconv(const void *X_host, void *out_host) override {
        set_data(X_host, 0);
        launch_transconv(Wi.p, X.p, numH.p, N, K, L, N, Tl, Tl + (L - 1));
        CK(cudaMemcpyAsync(out_host, numH.p, numH.n * sizeof(S), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
    }
    void prim_corr(const void *X_host, void *out_host) override {
        set_data(X_host, 0);
        launch_corr(X.p, N, N, Tl + (L - 1), nsplit_w, split_w, numW.p, nullptr);
        w_internal_to_julia<S><<<(unsigned)cdiv(KL() * N, 256), 256, 0, stream>>>(numW.p, Wtmp.p, K, N, L);
        post_launch();
        CK(cudaMemcpyAsync(out_host, Wtmp.p, Wtmp.n * sizeof(S), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
    }
    // compute_resids (src/common.jl:58-59): conv(W,H) - X through the conv kernel's residual epilogue
    void prim_resids(void *out_host) override {
        REQUIRE(have_data && have_factors, "compute_resids: data and factors must be set");
        DevBuf<S> out;
        out.alloc((size_t)(N * Tl));
        launch_conv(Wi.p, H, K, L, -(L - 1), Tl + (L - 1), 0, Tl, 2, out.p, nullptr);
        CK(cudaMemcpyAsync(out_host, out.p, out.n * sizeof(S), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
    }
    // shift_and_stack (src/common.jl:133-142): Htilde[(l*K+k), t] = H[k, t-l], zero for t < l (materialised only here, for tests
    // and small problems: the fit itself addresses the same rows through overlapping windows of H)
    void prim_shift_and_stack(void *out_host) override {
        REQUIRE(have_factors, "shift_and_stack: factors must be set");
        DevBuf<S> out;
        out.alloc((size_t)(KL() * Tl));
        shift_stack_kernel<S><<<(unsigned)cdiv(KL() * Tl, 256), 256, 0, stream>>>(H, out.p, K, L, Tl);
        post_launch();
        CK(cudaMemcpyAsync(out_host, out.p, out.n * sizeof(S), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
    }
};

cmf_ctx *make_ctx(int64_t N, int64_t T, int64_t t0, int64_t t1, int64_t K, int64_t L, int dtype, int alg, int device) {
    REQUIRE(N >= 1 && K >= 1 && L >= 1 && T >= 1, "dimensions must be positive");
    REQUIRE(L <= T, "need L <= T (src/common.jl:28-31)");
    REQUIRE(0 <= t0 && t0 < t1 && t1 <= T, "bad shard range");
    REQUIRE(dtype == CMF_F64 || dtype == CMF_F32, "dtype must be 0 (f64) or 1 (f32)");
    REQUIRE(alg == CMF_MULT || alg == CMF_HALS || alg == CMF_PGD, "alg must be 0 (mult), 1 (hals) or 2 (pgd)");
    const bool sharded = !(t0 == 0 && t1 == T);
    if (sharded) {
        REQUIRE(t1 - t0 >= L - 1, "each shard needs at least L-1 columns");
        if (alg == CMF_PGD) throw CmfError(CMF_ERR_UNSUPPORTED, "PGD is single-shard only in this version");
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        throw CmfError(CMF_ERR_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(e) + "); libcmf_sm100 has no CPU fallback");
    REQUIRE(device >= 0 && device < ndev, "bad device ordinal");
    cmf_ctx *c = (dtype == CMF_F64) ? static_cast<cmf_ctx *>(new Ctx<double>()) : static_cast<cmf_ctx *>(new Ctx<float>());
    c->N = N; c->T = T; c->t0 = t0; c->t1 = t1; c->Tl = t1 - t0; c->K = K; c->L = L;
    c->dtype = dtype; c->alg = alg; c->device = device;
    c->is_first = (t0 == 0); c->is_last = (t1 == T);
    try {
        if (dtype == CMF_F64) static_cast<Ctx<double> *>(c)->init();
        else static_cast<Ctx<float> *>(c)->init();
    } catch (...) {
        delete c;
        throw;
    }
    return c;
}

template <typename F>
int guarded(F &&f) {
    try {
        f();
        return CMF_OK;
    } catch (const CmfError &e) {
        g_err = e.what();
        return e.code;
    } catch (const std::exception &e) {
        g_err = e.what();
        return CMF_ERR_ARG;
    } catch (...) {
        g_err = "unknown error";
        return CMF_ERR_ARG;
    }
}

// Every ABI call runs on the handle's device and leaves the caller's current device as it found it.
struct DevGuard {
    int prev = -1;
    explicit DevGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        CK(cudaSetDevice(dev));
    }
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DevGuard(const DevGuard &) = delete;
    DevGuard &operator=(const DevGuard &) = delete;
};
// only restores the caller's current device on exit (constructors: the argument checks come before any CUDA call)
struct DevRestore {
    int prev = -1;
    DevRestore() { if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); } }
    ~DevRestore() { if (prev >= 0) cudaSetDevice(prev); }
};

// src/model.jl:91-107
bool converged(const double *loss_hist, int64_t n, int patience, double tol) {
    if (n <= patience) return false;
    for (int64_t i = n - patience; i < n; ++i)
        if (!(std::fabs(loss_hist[i] - loss_hist[i - 1]) < tol)) return false;
    return true;
}

// ------------------------------------------------------------------------------------------
// multi-GPU: T-sharding with NCCL inside the library (the reference is single-process)
// ------------------------------------------------------------------------------------------
#define NK(call)                                                                                   \
    do {                                                                                           \
        ncclResult_t r__ = (call);                                                                 \
        if (r__ != ncclSuccess)                                                                    \
            throw CmfError(CMF_ERR_NCCL, std::string(#call) + ": " + NcclApi::get().GetErrorString(r__)); \
    } while (0)

NcclApi &nccl() {
    NcclApi &a = NcclApi::get();
    if (!a.ok()) throw CmfError(CMF_ERR_NCCL, "NCCL is not available: " + a.why);
    return a;
}

// balanced contiguous partition of the time axis (the same rule as cmf.jl_b200/sharded.py ShardPlan)
void shard_range(int64_t T, int world, int rank, int64_t *t0, int64_t *t1) {
    const int64_t base = T / world, rem = T % world;
    *t0 = rank * base + std::min<int64_t>(rank, rem);
    *t1 = *t0 + base + (rank < rem ? 1 : 0);
}

ncclDataType_t nccl_dt(const cmf_ctx *h) { return h->dtype == CMF_F64 ? ncclFloat64 : ncclFloat32; }
size_t elem_size(const cmf_ctx *h) { return h->dtype == CMF_F64 ? 8 : 4; }

void c_allreduce(cmf_ctx *h, void *buf, size_t count, ncclDataType_t dt, ncclRedOp_t op = ncclSum) {
    if (h->world() == 1) return;
    NK(nccl().AllReduce(buf, buf, count, dt, op, h->comm.comm, h->stream));
}

// all-reduce of a few host doubles through the handle's scalar buffer (slots 4..7); synchronises the stream
void c_allreduce_host(cmf_ctx *h, double *vals, int n, ncclRedOp_t op = ncclSum) {
    if (h->world() == 1) return;
    REQUIRE(n <= 4, "at most 4 scalars");
    double *d = h->scalars() + 4;
    CK(cudaMemcpyAsync(d, vals, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    c_allreduce(h, d, (size_t)n, ncclFloat64, op);
    CK(cudaMemcpyAsync(vals, d, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
}

// L-1 columns of H each way (left neighbour's right halo <- my first L-1 columns; right neighbour's left halo <- my last)
void c_halo_exchange(cmf_ctx *h) {
    if (h->world() == 1 || h->L <= 1) return;
    void *sl, *sr, *rl, *rr;
    int64_t cnt;
    h->halo_buffers(&sl, &sr, &rl, &rr, &cnt);
    NcclApi &a = nccl();
    const ncclDataType_t dt = nccl_dt(h);
    const int r = h->rank(), w = h->world();
    NK(a.GroupStart());
    if (r + 1 < w) {
        NK(a.Send(sr, (size_t)cnt, dt, r + 1, h->comm.comm, h->stream));
        NK(a.Recv(rr, (size_t)cnt, dt, r + 1, h->comm.comm, h->stream));
    }
    if (r > 0) {
        NK(a.Send(sl, (size_t)cnt, dt, r - 1, h->comm.comm, h->stream));
        NK(a.Recv(rl, (size_t)cnt, dt, r - 1, h->comm.comm, h->stream));
    }
    NK(a.GroupEnd());
}

// ||X||_F over all shards (mult.jl:13, hals.jl:23)
void r_data_norm(cmf_ctx *h) {
    double ss = h->data_sumsq_local;
    c_allreduce_host(h, &ss, 1);
    h->data_sumsq_global = ss;
    h->data_norm = std::sqrt(ss);
    h->pgd_cur_loss = h->data_norm;
}

// <X, est> and ||est||^2 of the alpha rescale (src/model.jl:119-120), summed over shards
void r_init_scale_partials(cmf_ctx *h, double out[2]) {
    h->init_scale_partials(out);
    c_allreduce_host(h, out, 2);
}

// sum of squared residuals over ALL shards: one fixed-size all-reduce of the two loss scalars, one read-back
double r_loss_sumsq(cmf_ctx *h, bool force_direct = false, bool *expansion_used = nullptr) {
    const int saved = h->loss_mode;
    if (force_direct) h->loss_mode = 0;
    int n;
    try { n = h->loss_partial_dev(); } catch (...) { h->loss_mode = saved; throw; }
    h->loss_mode = saved;
    double *d = h->scalars();
    c_allreduce(h, d, 2, ncclFloat64);
    double v[2];
    CK(cudaMemcpyAsync(v, d, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (expansion_used) *expansion_used = (n == 2);
    ++(n == 2 ? h->n_loss_expansion : h->n_loss_direct);
    return n == 2 ? h->data_sumsq_global - 2.0 * v[0] + v[1] : v[0];
}

// Loss of the current factors, mult.jl:55-57.  With loss_mode 1 the algebraic expansion serves while the relative loss is
// above `loss_guard` (25 %): its error is ~2e-6/loss^2 relative.  At or below the guard the expansion is CALIBRATED: its
// error is a bias that moves slowly with the factors, so the direct residual pass is run every `calib_interval`
// evaluations, the difference is kept and subtracted in between, and at each direct pass the value the previous bias would
// have predicted is checked against it -- agreement to 5e-6 (relative, on the loss) doubles the interval (up to 16), more
// than 2e-5 halves it, more than 1e-4 three times in a row returns the handle to the direct pass for good.  Values returned
// at calibration points are the direct pass itself.  Every rank sees the same all-reduced numbers and takes the same branch.
double r_guarded_loss(cmf_ctx *h) {
    if (h->loss_mode != 1) return std::sqrt(std::max(r_loss_sumsq(h), 0.0)) / h->data_norm;
    bool exp_used = false;
    const double se = r_loss_sumsq(h, false, &exp_used);
    if (!exp_used) return std::sqrt(std::max(se, 0.0)) / h->data_norm;
    const double g2 = h->loss_guard * h->loss_guard * h->data_sumsq_global;
    if (!h->calib_have && se > g2) return std::sqrt(se) / h->data_norm;
    if (h->calib_have && h->calib_left > 0 && se - h->calib_bias > 0.0) {
        --h->calib_left;
        return std::sqrt(se - h->calib_bias) / h->data_norm;
    }
    const double sd = r_loss_sumsq(h, true);
    if (h->calib_have && sd > 0.0) {
        const double err = std::fabs((se - h->calib_bias) - sd) / (2.0 * sd);
        h->calib_last_err = err;
        if (err > 1e-4) { h->calib_interval = 1; ++h->calib_fail; }
        else {
            h->calib_fail = 0;
            if (err > 2e-5) h->calib_interval = std::max(1, h->calib_interval / 2);
            else if (err < 5e-6) h->calib_interval = std::min(h->calib_max_interval, h->calib_interval * 2);
        }
        if (h->calib_fail >= 3) { h->loss_mode = 0; h->calib_reset(); }      // the expansion is not reliable on this problem
    }
    h->calib_bias = se - sd;
    h->calib_have = true;
    h->calib_left = h->calib_interval - 1;
    return std::sqrt(std::max(sd, 0.0)) / h->data_norm;
}

// update_motifs! (mult.jl:23-39 / hals.jl:31-34 / pgd.jl:158-178) on this rank's shard
void r_update_motifs(cmf_ctx *h, double l1W, double l2W) {
    if (h->alg == CMF_PGD) { h->pgd_update_motifs(l1W, l2W); return; }
    h->w_partials();
    const int64_t KL = h->K * h->L;
    if (h->world() > 1 && h->alg == CMF_MULT && KL % h->world() == 0 && !getenv("CMF_W_REPLICATED")) {
        // MultUpdate: the W update is sharded by unfolded rows j = l*K + k (contiguous row blocks of numW / W):
        // reduce-scatter of numW on the side stream while this rank forms its rows of denomW = G W, update of those
        // rows, all-gather of W.  Same bytes on the wire as an all-reduce; G W, the ratio update and the eps floor cost 1/world.
        NcclApi &a = nccl();
        void *p; int64_t cnt; int dt;
        h->exchange_buffer(1, &p, &cnt, &dt);
        c_allreduce(h, p, (size_t)cnt, ncclFloat64);                                   // Gram + H tail (small, needed by G)
        void *nw, *w; int64_t row;
        h->w_rows_buffers(&nw, &w, &row);
        const int64_t rows = KL / h->world(), j0 = rows * h->rank(), chunk = rows * row;
        const size_t es = elem_size(h);
        CK(cudaEventRecord(h->ev_a, h->stream));
        CK(cudaStreamWaitEvent(h->comm_stream, h->ev_a, 0));
        NK(a.ReduceScatter(nw, (char *)nw + (size_t)(j0 * row) * es, (size_t)chunk, nccl_dt(h), ncclSum, h->comm.comm, h->comm_stream));
        CK(cudaEventRecord(h->ev_b, h->comm_stream));
        h->w_denom_rows(j0, j0 + rows);
        CK(cudaStreamWaitEvent(h->stream, h->ev_b, 0));
        h->w_update_rows(l1W, l2W, j0, j0 + rows);
        // all-gather of W on the side stream; the one piece of the H half-step that does not need W (the spectrum of H for
        // denomH) runs meanwhile
        CK(cudaEventRecord(h->ev_a, h->stream));
        CK(cudaStreamWaitEvent(h->comm_stream, h->ev_a, 0));
        NK(a.AllGather((char *)w + (size_t)(j0 * row) * es, w, (size_t)chunk, nccl_dt(h), h->comm.comm, h->comm_stream));
        CK(cudaEventRecord(h->ev_b, h->comm_stream));
        if (h->alg == CMF_MULT) h->fd_denomH_prefetch();
        CK(cudaStreamWaitEvent(h->stream, h->ev_b, 0));
        return;
    }
    if (h->world() > 1) {
        void *p; int64_t cnt; int dt;
        h->exchange_buffer(0, &p, &cnt, &dt);
        c_allreduce(h, p, (size_t)cnt, dt == CMF_F64 ? ncclFloat64 : ncclFloat32);     // numW            (K*N*L)
        h->exchange_buffer(1, &p, &cnt, &dt);
        c_allreduce(h, p, (size_t)cnt, ncclFloat64);                                   // Gram + H tail   (K*K*L + (L-1)*K doubles)
    }
    h->w_apply(l1W, l2W);                                                              // identical update on every rank
}

// HALS H sweep of a T-sharded fit: gather Q and H on rank 0, sweep once in the reference's order, scatter H
void r_hals_sweep_sharded(cmf_ctx *h, double l1H, double l2H) {
    NcclApi &a = nccl();
    const ncclDataType_t dt = nccl_dt(h);
    const size_t es = elem_size(h);
    const int w = h->world(), r = h->rank();
    void *ql, *hl, *qf = nullptr, *hf = nullptr;
    int64_t cnt;
    h->hals_h_local(&ql, &hl, &cnt);
    if (r == 0) h->hals_h_full(&qf, &hf);
    for (int pass = 0; pass < 2; ++pass) {          // 0: gather Q and H, 1: scatter H
        if (pass == 1 && r == 0) h->hals_h_sweep(1, l1H, l2H);
        NK(a.GroupStart());
        if (r == 0) {
            for (int p = 1; p < w; ++p) {
                int64_t a0, a1;
                shard_range(h->T, w, p, &a0, &a1);
                const size_t off = (size_t)(a0 * h->K) * es, c = (size_t)((a1 - a0) * h->K);
                if (pass == 0) {
                    NK(a.Recv((char *)qf + off, c, dt, p, h->comm.comm, h->stream));
                    NK(a.Recv((char *)hf + off, c, dt, p, h->comm.comm, h->stream));
                } else {
                    NK(a.Send((char *)hf + off, c, dt, p, h->comm.comm, h->stream));
                }
            }
        } else if (pass == 0) {
            NK(a.Send(ql, (size_t)cnt, dt, 0, h->comm.comm, h->stream));
            NK(a.Send(hl, (size_t)cnt, dt, 0, h->comm.comm, h->stream));
        } else {
            NK(a.Recv(hl, (size_t)cnt, dt, 0, h->comm.comm, h->stream));
        }
        NK(a.GroupEnd());
        if (r == 0) {
            if (pass == 0) {
                CK(cudaMemcpyAsync(qf, ql, (size_t)cnt * es, cudaMemcpyDeviceToDevice, h->stream));
                CK(cudaMemcpyAsync(hf, hl, (size_t)cnt * es, cudaMemcpyDeviceToDevice, h->stream));
            } else {
                CK(cudaMemcpyAsync(hl, hf, (size_t)cnt * es, cudaMemcpyDeviceToDevice, h->stream));
            }
        }
    }
}

// loss = update_feature_maps! (mult.jl:42-58 / hals.jl:37-42 / pgd.jl:181-203)
double r_update_feature_maps(cmf_ctx *h, double l1H, double l2H) {
    if (h->alg == CMF_PGD) {
        const double loss = std::sqrt(h->pgd_update_feature_maps(l1H, l2H)) / h->data_norm;
        if (h->loss_mode == 1 && !(loss > 0.25)) h->loss_mode = 0;   // PGD adapts its step on this value: no re-evaluation
        return loss;
    }
    if (h->alg == CMF_HALS) {
        h->hals_h_prepare();
        if (h->world() == 1) h->hals_h_sweep(0, l1H, l2H);
        else r_hals_sweep_sharded(h, l1H, l2H);
        h->h_changed();
    } else {
        h->h_update(l1H, l2H);
    }
    c_halo_exchange(h);
    return r_guarded_loss(h);
}

// src/algs/alternating.jl:16-71 on one rank (every rank runs the same loop: the loss is all-reduced, and the elapsed
// time that decides `max_time` is rank 0's, so all ranks take the same branches)
void r_fit(cmf_ctx *h, int64_t max_itr, double max_time, int eval_mode, int check_convergence, int patience, double tol,
           double l1W, double l2W, double l1H, double l2H, std::vector<double> &loss_hist, std::vector<double> &time_hist,
           int64_t cap, int *converged_early) {
    *converged_early = 0;
    loss_hist.clear(); time_hist.clear();
    loss_hist.push_back(r_guarded_loss(h));   // alternating.jl:37
    time_hist.push_back(0.0);
    int64_t itr = 1;
    const bool timed = max_time < 1e300;
    while ((max_itr < 0 || itr <= max_itr) && time_hist.back() <= max_time) {   // alternating.jl:45
        ++itr;
        REQUIRE((int64_t)loss_hist.size() < cap, "history capacity exhausted");
        auto t_start = std::chrono::steady_clock::now();
        if (!eval_mode) r_update_motifs(h, l1W, l2W);
        const double loss = r_update_feature_maps(h, l1H, l2H);
        double dur = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
        if (timed && h->world() > 1) {          // rank 0's clock decides for everyone
            if (h->rank() != 0) dur = 0.0;
            c_allreduce_host(h, &dur, 1);
        }
        time_hist.push_back(time_hist.back() + dur);
        loss_hist.push_back(loss);
        if (check_convergence && converged(loss_hist.data(), (int64_t)loss_hist.size(), patience, tol)) {   // alternating.jl:63-66
            *converged_early = 1;
            break;
        }
    }
}

void r_set_engine(cmf_ctx *h, int engine) {
    REQUIRE(engine >= 0 && engine <= 2, "engine must be 0 (SIMT), 1 (tcgen05, time domain) or 2 (tcgen05, frequency domain)");
    if (engine == 2 && !h->fd_available())
        throw CmfError(CMF_ERR_UNSUPPORTED, "the frequency-domain engine needs an fp32 handle with K <= 128, L <= 256, N % 8 == 0 on sm_100 and room for the spectrum of X");
    if (engine < 2) h->fd_release();               // the spectrum of X and the time-domain planes never coexist
    if (engine >= 1 && !h->tc_available())
        throw CmfError(CMF_ERR_UNSUPPORTED, "tcgen05 engine needs an fp32 MultUpdate handle with K <= 128 and N % 8 == 0 on sm_100");
    h->engine = engine;
    if (!h->loss_mode_explicit) h->loss_mode = (engine == 2) ? 1 : 0;   // the expansion is the default of engine 2 only
}

// every rank must run the same engine (the collectives are matched by program order): take the minimum of what the
// ranks selected by themselves (shard lengths and free memory may differ by a little)
void r_agree_engine(cmf_ctx *h) {
    if (h->world() == 1) return;
    double e = (double)h->engine;
    c_allreduce_host(h, &e, 1, ncclMin);
    if ((int)e != h->engine) r_set_engine(h, (int)e);
}

void attach_comm(cmf_ctx *h, ncclComm_t comm, int rank, int world, bool owned) {
    int64_t a0, a1;
    shard_range(h->T, world, rank, &a0, &a1);
    REQUIRE(a0 == h->t0 && a1 == h->t1, "the handle's column range is not the balanced shard of this rank (use cmf_shard_range)");
    h->comm.comm = comm; h->comm.rank = rank; h->comm.world = world; h->comm.owned = owned;
    if (world > 1 && !h->comm_stream) {
        CK(cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&h->ev_a, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_b, cudaEventDisableTiming));
    }
    r_agree_engine(h);
}

}  // namespace

// single-process multi-GPU group: one rank context per device, one worker thread per rank
struct cmf_multi {
    std::vector<cmf_ctx *> ranks;
    std::vector<ncclComm_t> comms;
    std::unique_ptr<cmf::Workers> workers;
};

namespace {

struct MultiCtx : cmf_ctx {
    cmf_multi m;
    ~MultiCtx() override {
        if (m.workers) {
            try {
                m.workers->run_all([&](int i) {
                    cmf_ctx *r = m.ranks[(size_t)i];
                    if (!r) return;
                    cudaSetDevice(r->device);
                    delete r;
                    if ((size_t)i < m.comms.size() && m.comms[(size_t)i]) NcclApi::get().CommDestroy(m.comms[(size_t)i]);
                });
            } catch (...) {}
        }
    }
};

// runs f(rank context) on the handle itself or, for a group handle, on every rank concurrently (worker threads)
template <typename F>
void on_ranks(cmf_handle h, F &&f) {
    REQUIRE(h != nullptr, "null handle");
    if (h->multi) {
        cmf_multi *m = h->multi;
        m->workers->run_all([&](int i) {
            cmf_ctx *r = m->ranks[(size_t)i];
            DevGuard g(r->device);
            f(r);
        });
    } else {
        DevGuard g(h->device);
        f(h);
    }
}
cmf_ctx *rank0(cmf_handle h) { return h->multi ? h->multi->ranks[0] : h; }

}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

const char *cmf_last_error(void) { return g_err.c_str(); }

int cmf_create(cmf_handle *out, int64_t N, int64_t T, int64_t K, int64_t L, int dtype, int alg, int device) {
    return guarded([&] {
        REQUIRE(out != nullptr, "null output pointer");
        DevRestore g;
        *out = make_ctx(N, T, 0, T, K, L, dtype, alg, device);
    });
}

int cmf_create_shard(cmf_handle *out, int64_t N, int64_t T_global, int64_t t_begin, int64_t t_end, int64_t K,
                     int64_t L, int dtype, int alg, int device) {
    return guarded([&] {
        REQUIRE(out != nullptr, "null output pointer");
        DevRestore g;
        *out = make_ctx(N, T_global, t_begin, t_end, K, L, dtype, alg, device);
    });
}

int cmf_shard_range(int64_t T, int world, int rank, int64_t *t_begin, int64_t *t_end) {
    return guarded([&] {
        REQUIRE(t_begin && t_end, "null output");
        REQUIRE(world >= 1 && rank >= 0 && rank < world && T >= world, "need 0 <= rank < world <= T");
        shard_range(T, world, rank, t_begin, t_end);
    });
}

int cmf_comm_unique_id(void *id_out) {
    return guarded([&] {
        REQUIRE(id_out, "null output");
        ncclUniqueId id;
        NK(nccl().GetUniqueId(&id));
        memcpy(id_out, &id, sizeof(id));
    });
}

// One communicator per (device, rank, world) of this process is kept alive and reused by later handles created with
// unique_id == NULL (building one costs ~0.5 s; a fit does not need a new one).  Communicators in the cache live until exit.
struct CommKey { int device, rank, world; bool operator<(const CommKey &o) const { return std::tie(device, rank, world) < std::tie(o.device, o.rank, o.world); } };
static std::map<CommKey, ncclComm_t> g_comm_cache;
static std::mutex g_comm_mu;

int cmf_create_rank(cmf_handle *out, int64_t N, int64_t T, int64_t K, int64_t L, int dtype, int alg, int device,
                    const void *unique_id, int rank, int world) {
    return guarded([&] {
        REQUIRE(out != nullptr, "null pointer");
        REQUIRE(world >= 1 && rank >= 0 && rank < world, "need 0 <= rank < world");
        DevRestore g;
        int64_t a0, a1;
        shard_range(T, world, rank, &a0, &a1);
        cmf_ctx *c = make_ctx(N, T, a0, a1, K, L, dtype, alg, device);
        try {
            ncclComm_t comm = nullptr;
            if (world > 1) {
                std::lock_guard<std::mutex> lk(g_comm_mu);
                const CommKey key{device, rank, world};
                if (unique_id == nullptr) {
                    auto it = g_comm_cache.find(key);
                    REQUIRE(it != g_comm_cache.end(), "unique_id == NULL reuses this process's communicator for (device, rank, world), but none exists yet");
                    comm = it->second;
                } else {
                    ncclUniqueId id;
                    memcpy(&id, unique_id, sizeof(id));
                    CK(cudaSetDevice(device));
                    NK(nccl().CommInitRank(&comm, world, id, rank));
                    g_comm_cache[key] = comm;            // replaces (and leaks until exit) an older one: it may still serve a live handle
                }
            }
            attach_comm(c, comm, rank, world, false);
        } catch (...) { delete c; throw; }
        *out = c;
    });
}

int cmf_create_multi(cmf_handle *out, int64_t N, int64_t T, int64_t K, int64_t L, int dtype, int alg, int ngpu,
                     const int *devices) {
    return guarded([&] {
        REQUIRE(out != nullptr, "null output pointer");
        REQUIRE(ngpu >= 1 && ngpu <= 64, "ngpu must be in 1..64");
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            throw CmfError(CMF_ERR_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(e) + "); libcmf_sm100 has no CPU fallback");
        std::vector<int> devs((size_t)ngpu);
        for (int i = 0; i < ngpu; ++i) {
            devs[(size_t)i] = devices ? devices[i] : i;
            REQUIRE(devs[(size_t)i] >= 0 && devs[(size_t)i] < ndev, "bad device ordinal in the device list");
        }
        if (ngpu == 1) {
            DevRestore g;
            *out = make_ctx(N, T, 0, T, K, L, dtype, alg, devs[0]);
            return;
        }
        int prev = 0;
        cudaGetDevice(&prev);
        auto *mc = new MultiCtx();
        mc->N = N; mc->T = T; mc->t0 = 0; mc->t1 = T; mc->Tl = T; mc->K = K; mc->L = L;
        mc->dtype = dtype; mc->alg = alg; mc->device = devs[0];
        mc->multi = &mc->m;
        try {
            mc->m.ranks.assign((size_t)ngpu, nullptr);
            mc->m.comms.assign((size_t)ngpu, nullptr);
            mc->m.workers.reset(new Workers(ngpu));
            NK(nccl().CommInitAll(mc->m.comms.data(), ngpu, devs.data()));
            cudaSetDevice(prev);
            mc->m.workers->run_all([&](int i) {
                DevGuard g(devs[(size_t)i]);
                int64_t a0, a1;
                shard_range(T, ngpu, i, &a0, &a1);
                mc->m.ranks[(size_t)i] = make_ctx(N, T, a0, a1, K, L, dtype, alg, devs[(size_t)i]);
            });
            mc->m.workers->run_all([&](int i) {
                cmf_ctx *r = mc->m.ranks[(size_t)i];
                DevGuard g(r->device);
                attach_comm(r, mc->m.comms[(size_t)i], i, ngpu, false);
            });
        } catch (...) {
            cudaSetDevice(prev);
            delete mc;
            throw;
        }
        *out = mc;
    });
}

int cmf_comm_info(cmf_handle h, int *rank_out, int *world_out, int64_t *t_begin, int64_t *t_end) {
    return guarded([&] {
        REQUIRE(h, "null handle");
        if (rank_out) *rank_out = h->multi ? 0 : h->rank();
        if (world_out) *world_out = h->multi ? (int)h->multi->ranks.size() : h->world();
        if (t_begin) *t_begin = h->t0;
        if (t_end) *t_end = h->t1;
    });
}

int cmf_destroy(cmf_handle h) {
    return guarded([&] {
        if (!h) return;
        if (h->multi) { delete h; return; }
        DevGuard g(h->device);
        ncclComm_t comm = (h->comm.owned ? h->comm.comm : nullptr);
        delete h;
        if (comm) NcclApi::get().CommDestroy(comm);
    });
}

int cmf_set_data(cmf_handle h, const void *X, int64_t first_col) {
    return guarded([&] { on_ranks(h, [&](cmf_ctx *r) { r->set_data(X, first_col); r_data_norm(r); }); });
}
int cmf_synth_data(cmf_handle h, uint64_t seed, int64_t K_true, int64_t L_true, double p_h, double noise) {
    return guarded([&] { on_ranks(h, [&](cmf_ctx *r) { r->synth_data(seed, K_true, L_true, p_h, noise); r_data_norm(r); }); });
}
int cmf_data_sumsq(cmf_handle h, double *out) {
    return guarded([&] {
        REQUIRE(out, "null output");
        cmf_ctx *r = rank0(h);
        REQUIRE(r->have_data, "no data");
        *out = (h->multi || r->world() > 1) ? r->data_sumsq_global : r->data_sumsq_local;
    });
}
int cmf_set_data_norm(cmf_handle h, double norm) {
    return guarded([&] { on_ranks(h, [&](cmf_ctx *r) { r->data_norm = norm; r->data_sumsq_global = norm * norm; }); });
}
int cmf_set_factors(cmf_handle h, const void *W, const void *H, int64_t first_col) {
    return guarded([&] { on_ranks(h, [&](cmf_ctx *r) { r->set_factors(W, H, first_col); }); });
}
int cmf_exchange_halos(cmf_handle h) {
    return guarded([&] { on_ranks(h, [&](cmf_ctx *r) { c_halo_exchange(r); r->h_changed(); }); });
}
int cmf_init_rand(cmf_handle h, uint64_t seed) {
    return guarded([&] { on_ranks(h, [&](cmf_ctx *r) { r->init_rand(seed); }); });
}
int cmf_init_scale_partials(cmf_handle h, double out[2]) {
    return guarded([&] {
        REQUIRE(out, "null output");
        on_ranks(h, [&](cmf_ctx *r) {
            double v[2];
            r_init_scale_partials(r, v);
            if (!h->multi || r->rank() == 0) { out[0] = v[0]; out[1] = v[1]; }
        });
    });
}
int cmf_scale_factors(cmf_handle h, double s) {
    return guarded([&] { on_ranks(h, [&](cmf_ctx *r) { r->scale_factors(s); }); });
}
int cmf_get_factors(cmf_handle h, void *W_out, void *H_out) {
    return guarded([&] {
        const bool grp = h && h->multi;
        on_ranks(h, [&](cmf_ctx *r) {
            REQUIRE(r->have_factors, "no factors");
            void *Hp = H_out;
            if (grp && H_out) Hp = (char *)H_out + (size_t)(r->t0 * r->K) * elem_size(r);     // group handle: H_out is K x T
            r->get_factors((!grp || r->rank() == 0) ? W_out : nullptr, Hp);
        });
    });
}

int cmf_update_motifs(cmf_handle h, double l1W, double l2W) {
    return guarded([&] { on_ranks(h, [&](cmf_ctx *r) { r_update_motifs(r, l1W, l2W); }); });
}

int cmf_update_feature_maps(cmf_handle h, double l1H, double l2H, double *loss_out) {
    return guarded([&] {
        on_ranks(h, [&](cmf_ctx *r) {
            const double loss = r_update_feature_maps(r, l1H, l2H);
            if (loss_out && (r->rank() == 0 || !h->multi)) *loss_out = loss;
        });
    });
}

int cmf_loss(cmf_handle h, double *loss_out) {
    return guarded([&] {
        REQUIRE(loss_out, "null output");
        on_ranks(h, [&](cmf_ctx *r) {
            REQUIRE(r->world() > 1 || (r->is_first && r->is_last), "shards without a communicator use cmf_loss_partial");
            const double loss = r_guarded_loss(r);
            if (r->rank() == 0 || !h->multi) *loss_out = loss;
        });
    });
}

int cmf_fit(cmf_handle h, int64_t max_itr, double max_time, int eval_mode, int check_convergence, int patience,
            double tol, double l1W, double l2W, double l1H, double l2H, double *loss_hist, double *time_hist,
            int64_t cap, int64_t *n_hist, int *converged_early) {
    return guarded([&] {
        REQUIRE(loss_hist && time_hist && n_hist, "null history pointers");
        REQUIRE(patience >= 1, "patience must be >= 1 (src/algs/alternating.jl:30)");
        REQUIRE(cap >= 1, "history capacity must be >= 1");
        if (converged_early) *converged_early = 0;
        on_ranks(h, [&](cmf_ctx *r) {
            REQUIRE(r->world() > 1 || (r->is_first && r->is_last), "shards without a communicator use the split-phase calls");
            std::vector<double> lh, th;
            int early = 0;
            r_fit(r, max_itr, max_time, eval_mode, check_convergence, patience, tol, l1W, l2W, l1H, l2H, lh, th, cap, &early);
            if (r->rank() == 0 || !h->multi) {
                std::copy(lh.begin(), lh.end(), loss_hist);
                std::copy(th.begin(), th.end(), time_hist);
                *n_hist = (int64_t)lh.size();
                if (converged_early) *converged_early = early;
            }
        });
    });
}

int cmf_w_partials(cmf_handle h) {
    return guarded([&] { REQUIRE(h && !h->multi, "single-rank call"); DevGuard g(h->device); h->w_partials(); });
}
int cmf_w_apply(cmf_handle h, double l1W, double l2W) {
    return guarded([&] { REQUIRE(h && !h->multi, "single-rank call"); DevGuard g(h->device); h->w_apply(l1W, l2W); });
}
int cmf_h_update(cmf_handle h, double l1H, double l2H) {
    return guarded([&] {
        REQUIRE(h && !h->multi, "single-rank call");
        DevGuard g(h->device);
        REQUIRE(h->alg == CMF_MULT, "the split-phase H update is MultUpdate only");
        h->h_update(l1H, l2H);
    });
}
int cmf_loss_partial(cmf_handle h, double *sumsq_out) {
    return guarded([&] { REQUIRE(h && !h->multi, "single-rank call"); REQUIRE(sumsq_out, "null output"); DevGuard g(h->device); *sumsq_out = h->loss_partial(); });
}
int cmf_exchange_buffer(cmf_handle h, int which, void **dev_ptr, int64_t *count, int *dtype) {
    return guarded([&] { REQUIRE(h && !h->multi, "single-rank call"); REQUIRE(dev_ptr && count && dtype, "null output"); h->exchange_buffer(which, dev_ptr, count, dtype); });
}
int cmf_halo_buffers(cmf_handle h, void **sl, void **sr, void **rl, void **rr, int64_t *count) {
    return guarded([&] { REQUIRE(h && !h->multi, "single-rank call"); REQUIRE(sl && sr && rl && rr && count, "null output"); h->halo_buffers(sl, sr, rl, rr, count); });
}
int cmf_sync(cmf_handle h) {
    return guarded([&] { on_ranks(h, [&](cmf_ctx *r) { CK(cudaStreamSynchronize(r->stream)); }); });
}
int cmf_launch_count(cmf_handle h, int64_t *out) {
    return guarded([&] {
        REQUIRE(h && out, "null argument");
        if (!h->multi) { *out = h->launches; return; }
        int64_t n = 0;
        for (cmf_ctx *r : h->multi->ranks) n += r->launches;
        *out = n;
    });
}
int cmf_stream(cmf_handle h, void **stream_out) {
    return guarded([&] { REQUIRE(h && stream_out, "null argument"); *stream_out = (void *)rank0(h)->stream; });
}
int cmf_get_data(cmf_handle h, void *X_out, int with_halo) {
    return guarded([&] {
        REQUIRE(X_out, "null output");
        const bool grp = h && h->multi;
        on_ranks(h, [&](cmf_ctx *r) {
            REQUIRE(r->have_data, "no data");
            r->get_data(grp ? (char *)X_out + (size_t)(r->t0 * r->N) * elem_size(r) : X_out, grp ? 0 : with_halo);
        });
    });
}
int cmf_profile(cmf_handle h, int enable) {
    return guarded([&] {
        on_ranks(h, [&](cmf_ctx *r) {
            CK(cudaStreamSynchronize(r->stream));
            r->prof_clear();
            r->profiling = enable != 0;
        });
    });
}
int cmf_profile_read(cmf_handle h, int which, double *ms_total, int64_t *count) {
    return guarded([&] {
        REQUIRE(h && ms_total && count, "null output");
        REQUIRE(which >= 0 && which < PROF_NCLASS, "which must be 0 (conv), 1 (transconv), 2 (corr) or 3 (HALS H sweep)");
        cmf_ctx *r = rank0(h);
        DevGuard g(r->device);
        CK(cudaStreamSynchronize(r->stream));
        double tot = 0.0;
        int64_t n = 0;
        for (auto &e : r->prof_events) {
            if (e.which != which) continue;
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, e.a, e.b));
            tot += ms;
            ++n;
        }
        *ms_total = tot;
        *count = n;
    });
}
int cmf_set_stream(cmf_handle h, void *stream) {
    return guarded([&] {
        REQUIRE(h && !h->multi, "single-rank call");
        DevGuard g(h->device);
        CK(cudaStreamSynchronize(h->stream));
        if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
        h->own_stream = false;
        h->stream = (cudaStream_t)stream;
    });
}
int cmf_set_engine(cmf_handle h, int engine) {
    return guarded([&] { on_ranks(h, [&](cmf_ctx *r) { r_set_engine(r, engine); }); });
}

int cmf_set_loss_mode(cmf_handle h, int mode) {
    return guarded([&] {
        REQUIRE(mode == 0 || mode == 1, "loss mode must be 0 (direct) or 1 (expansion)");
        on_ranks(h, [&](cmf_ctx *r) { r->loss_mode = mode; r->loss_mode_explicit = true; r->calib_reset(); });
    });
}
int cmf_set_loss_guard(cmf_handle h, double guard, int max_interval) {
    return guarded([&] {
        REQUIRE(guard >= 0.0, "loss guard must be >= 0");
        REQUIRE(max_interval >= 1 && max_interval <= 1024, "calibration interval must be in [1, 1024]");
        on_ranks(h, [&](cmf_ctx *r) { r->loss_guard = guard; r->calib_max_interval = max_interval; r->calib_reset(); });
    });
}
int cmf_get_loss_stats(cmf_handle h, int64_t *n_direct, int64_t *n_expansion, int *interval, double *last_err) {
    return guarded([&] {
        REQUIRE(h, "null handle");
        cmf_ctx *r = rank0(h);
        if (n_direct) *n_direct = r->n_loss_direct;
        if (n_expansion) *n_expansion = r->n_loss_expansion;
        if (interval) *interval = r->calib_have ? r->calib_interval : 0;
        if (last_err) *last_err = r->calib_last_err;
    });
}
int cmf_get_loss_mode(cmf_handle h, int *mode_out) {
    return guarded([&] { REQUIRE(h && mode_out, "null argument"); *mode_out = rank0(h)->loss_mode; });
}
int cmf_get_fd_layout(cmf_handle h, int *block_len, int *hop, int64_t *nblocks) {
    return guarded([&] {
        REQUIRE(h && block_len && hop && nblocks, "null argument");
        cmf_ctx *r = rank0(h);
        r->fd_layout(block_len, hop, nblocks);
    });
}
int cmf_get_engine(cmf_handle h, int *engine_out) {
    return guarded([&] { REQUIRE(h && engine_out, "null argument"); *engine_out = rank0(h)->engine; });
}

static int current_device() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) d = 0;
    return d;
}

int cmf_set_pgd_constraints(cmf_handle h, int constrW, int constrH) {
    return guarded([&] { REQUIRE(h && !h->multi, "single-rank call"); h->set_pgd_constraints(constrW, constrH); });
}
int cmf_set_pgd_loss(cmf_handle h, int loss_func, const void *mask) {
    return guarded([&] { REQUIRE(h && !h->multi, "single-rank call"); DevGuard g(h->device); h->set_pgd_loss(loss_func, mask); });
}

int cmf_tensor_conv(int64_t N, int64_t T, int64_t K, int64_t L, int dtype, const void *W, const void *H, void *out) {
    return guarded([&] {
        REQUIRE(W && H && out, "null pointer");
        cmf_ctx *c = make_ctx(N, T, 0, T, K, L, dtype, CMF_MULT, current_device());
        try { c->set_factors(W, H, 0); c->prim_conv(out); } catch (...) { delete c; throw; }
        delete c;
    });
}
int cmf_tensor_transconv(int64_t N, int64_t T, int64_t K, int64_t L, int dtype, const void *W, const void *X, void *out) {
    return guarded([&] {
        REQUIRE(W && X && out, "null pointer");
        cmf_ctx *c = make_ctx(N, T, 0, T, K, L, dtype, CMF_MULT, current_device());
        try {
            std::vector<char> Hz((size_t)(K * T) * (dtype == CMF_F64 ? 8 : 4), 0);
            c->set_factors(W, Hz.data(), 0);
            c->prim_transconv(X, out);
        } catch (...) { delete c; throw; }
        delete c;
    });
}
int cmf_corr_w(int64_t N, int64_t T, int64_t K, int64_t L, int dtype, const void *H, const void *X, void *out) {
    return guarded([&] {
        REQUIRE(H && X && out, "null pointer");
        cmf_ctx *c = make_ctx(N, T, 0, T, K, L, dtype, CMF_MULT, current_device());
        try {
            std::vector<char> Wz((size_t)(K * N * L) * (dtype == CMF_F64 ? 8 : 4), 0);
            c->set_factors(Wz.data(), H, 0);
            c->prim_corr(X, out);
        } catch (...) { delete c; throw; }
        delete c;
    });
}
int cmf_compute_resids(int64_t N, int64_t T, int64_t K, int64_t L, int dtype, const void *X, const void *W, const void *H, void *out) {
    return guarded([&] {
        REQUIRE(X && W && H && out, "null pointer");
        cmf_ctx *c = make_ctx(N, T, 0, T, K, L, dtype, CMF_MULT, current_device());
        try { c->set_data(X, 0); c->set_factors(W, H, 0); c->prim_resids(out); } catch (...) { delete c; throw; }
        delete c;
    });
}
int cmf_shift_and_stack(int64_t K, int64_t T, int64_t L, int dtype, const void *H, void *out) {
    return guarded([&] {
        REQUIRE(H && out, "null pointer");
        cmf_ctx *c = make_ctx(8, T, 0, T, K, L, dtype, CMF_MULT, current_device());
        try {
            std::vector<char> Wz((size_t)(K * 8 * L) * (dtype == CMF_F64 ? 8 : 4), 0);
            c->set_factors(Wz.data(), H, 0);
            c->prim_shift_and_stack(out);
        } catch (...) { delete c; throw; }
        delete c;
    });
}

}  // extern "C"

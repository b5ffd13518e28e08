// dcnv3_b200 — C-ABI entry points (include/dcnv3_b200.h) and kernel dispatch.
//
// Replaces the reference's ATen host wrappers
// (/root/reference/models/ops_dcnv3/src/cuda/dcnv3_cuda.cu:21-85, :87-173) and
// launchers (dcnv3_im2col_cuda.cuh:841-868, :870-1045).  No torch headers: the
// caller owns every buffer and passes the stream.
#include "dcnv3_b200.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <atomic>
#include <mutex>
#include <initializer_list>

#include "dcnv3_kernels.cuh"
#include "dcnv3_bwd_tile.cuh"
#include "dcnv3_imat.cuh"
#include "dcnv3_win.cuh"

using namespace dcnv3;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char *what) {
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return (int)e;
}

size_t dtype_size(int dtype) {
    switch (dtype) {
        case DCNV3_B200_F32: return 4;
        case DCNV3_B200_F16: return 2;
        case DCNV3_B200_BF16: return 2;
        case DCNV3_B200_F64: return 8;
        default: return 0;
    }
}

// Validates the geometry (the reference's AT_ASSERTMs, dcnv3_cuda.cu:48-53, plus what it
// never checked) and derives Ho/Wo (dcnv3_cuda.cu:40-45).
int make_geo(const dcnv3_b200_geometry *g, Geo &q) {
    if (!g) return fail(DCNV3_B200_ENULL, "geometry is null");
    if (g->N < 0 || g->H <= 0 || g->W <= 0)
        return fail(DCNV3_B200_EINVAL, "bad input extent N=%d H=%d W=%d", g->N, g->H, g->W);
    if (g->group <= 0 || g->group_channels <= 0)
        return fail(DCNV3_B200_EINVAL, "group (%d) and group_channels (%d) must be positive",
                    g->group, g->group_channels);
    if (g->kernel_h <= 0 || g->kernel_w <= 0 || g->stride_h <= 0 || g->stride_w <= 0 ||
        g->pad_h < 0 || g->pad_w < 0 || g->dilation_h <= 0 || g->dilation_w <= 0)
        return fail(DCNV3_B200_EINVAL,
                    "bad kernel geometry k=(%d,%d) stride=(%d,%d) pad=(%d,%d) dil=(%d,%d)",
                    g->kernel_h, g->kernel_w, g->stride_h, g->stride_w, g->pad_h, g->pad_w,
                    g->dilation_h, g->dilation_w);
    q.N = g->N; q.H = g->H; q.W = g->W; q.G = g->group; q.gc = g->group_channels;
    q.kh = g->kernel_h; q.kw = g->kernel_w; q.sh = g->stride_h; q.sw = g->stride_w;
    q.ph = g->pad_h; q.pw = g->pad_w; q.dh = g->dilation_h; q.dw = g->dilation_w;
    q.scale = g->offset_scale;
    const long long C = (long long)q.G * q.gc;
    if (C > (1 << 24)) return fail(DCNV3_B200_ERANGE, "too many channels (%lld)", C);
    q.C = (int)C;
    q.Ho = (q.H + 2 * q.ph - (q.dh * (q.kh - 1) + 1)) / q.sh + 1;
    q.Wo = (q.W + 2 * q.pw - (q.dw * (q.kw - 1) + 1)) / q.sw + 1;
    if (q.H + 2 * q.ph < q.dh * (q.kh - 1) + 1 || q.W + 2 * q.pw < q.dw * (q.kw - 1) + 1 ||
        q.Ho <= 0 || q.Wo <= 0)
        return fail(DCNV3_B200_EINVAL, "kernel window larger than the padded input (Ho=%d Wo=%d)",
                    q.Ho, q.Wo);
    q.P = q.kh * q.kw;
    q.opitch = q.G * q.P * 2;  // separate, dense offset / mask tensors (dcnv3_b200_*_packed overrides these)
    q.mpitch = q.G * q.P;
    q.half_h = (q.dh * (q.kh - 1)) >> 1;
    q.half_w = (q.dw * (q.kw - 1)) >> 1;
    // per-image element counts are addressed with 32-bit ints inside the kernels
    if ((long long)q.H * q.W * q.C >= (1LL << 31) ||
        (long long)q.Ho * q.Wo * q.G * q.P * 2 >= (1LL << 31))
        return fail(DCNV3_B200_ERANGE, "one image exceeds 2^31 elements");
    return 0;
}

int check_device() {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(DCNV3_B200_EDEVICE, "no CUDA device: %s (this library has no CPU path)",
                    cudaGetErrorString(e));
    }
    static int ok_mask[64] = {0};  // benign race: idempotent
    if (dev < 64 && ok_mask[dev]) return 0;
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute");
    if (major != 10)
        return fail(DCNV3_B200_EDEVICE,
                    "device %d has compute capability %d.x; dcnv3_b200 is built for sm_100a only",
                    dev, major);
    if (dev < 64) ok_mask[dev] = 1;
    return 0;
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

struct Plan {
    bool vec;
    int bpl;  // bytes of channels per lane: 16 (LDG.128) or 32 (LDG.256)
    int vec_per_pix, lanes_per_group;
    unsigned total_vec;
};

bool aligned_to(const void *p, int a) { return (reinterpret_cast<uintptr_t>(p) & (uintptr_t)(a - 1)) == 0; }

// Tuning knobs (environment, all optional).  Read ONCE per process — the hot path must not call getenv — and
// cached; dcnv3_b200_reload_knobs() (test-only entry point) reads them again.
//   DCNV3_B200_BPL=8|16|32          bytes of channels per lane of the vector kernels
//   DCNV3_B200_PDL=0                no programmatic dependent launch
//   DCNV3_B200_FWD=vec|imat|pts|win forward kernel family
//   DCNV3_B200_BWD=vec|tile|imat|win backward kernel family (ACC_OPMATH: vec|tile|imat; ACC_TILE: win, or vec = ACC_STORAGE path)
//   DCNV3_B200_TILE="R,warps"       halo / warps of the experimental privatised backward
//   DCNV3_B200_GUARD_PER_CTA=n      logical blocks per CTA of the selector-guarded vector backward
struct Knobs {
    int bpl = 0;
    bool pdl = true;
    int fwd = 0, bwd = 0;  // 0 default, 1 vec, 2 tile, 3 imat, 4 pts, 5 win
    int tile_R = 2, tile_warps = 4;
    unsigned guard_per_cta = 4;
};
Knobs g_knobs;
std::atomic<bool> g_knobs_ready{false};
std::mutex g_knobs_mu;

int family_of(const char *e) {
    if (!e) return 0;
    if (!strcmp(e, "vec")) return 1;
    if (!strcmp(e, "tile")) return 2;
    if (!strcmp(e, "imat")) return 3;
    if (!strcmp(e, "pts")) return 4;
    if (!strcmp(e, "win")) return 5;
    return 0;
}
void read_knobs() {
    Knobs k;
    if (const char *e = getenv("DCNV3_B200_BPL")) { const int v = atoi(e); k.bpl = (v == 8 || v == 16 || v == 32) ? v : 0; }
    if (const char *e = getenv("DCNV3_B200_PDL")) k.pdl = e[0] != '0';
    k.fwd = family_of(getenv("DCNV3_B200_FWD"));
    k.bwd = family_of(getenv("DCNV3_B200_BWD"));
    if (const char *e = getenv("DCNV3_B200_TILE")) {
        int c, d;
        if (sscanf(e, "%d,%d", &c, &d) == 2 && c >= 0 && c <= 8 && d >= 1 && d <= 4) { k.tile_R = c; k.tile_warps = d; }
    }
    if (const char *e = getenv("DCNV3_B200_GUARD_PER_CTA")) { const int v = atoi(e); k.guard_per_cta = v >= 1 && v <= 64 ? (unsigned)v : 4u; }
    std::lock_guard<std::mutex> lk(g_knobs_mu);
    g_knobs = k;
    g_knobs_ready.store(true, std::memory_order_release);
}
inline const Knobs &knobs() {
    if (!g_knobs_ready.load(std::memory_order_acquire)) read_knobs();
    return g_knobs;
}
int forced_bpl() { return knobs().bpl; }

// Can this call take the vector kernels, and with how many bytes per lane?
// `ptrs` are the channel-vector tensors (input, output / grad_output): they must be aligned to
// the lane width.  `off` only needs the alignment of one (x, y) pair.
template <typename T>
Plan plan_vec(const Geo &q, size_t n_pix, bool logits, std::initializer_list<const void *> ptrs,
              const void *off, size_t acc_elem, bool allow8 = false) {
    Plan pl{false, 0, 0, 0, 0};
    if (sizeof(T) > 4) return pl;  // f64 always generic
    if (logits && !(q.kh == 3 && q.kw == 3)) return pl;
    if (reinterpret_cast<uintptr_t>(off) & (2 * sizeof(T) - 1)) return pl;
    // byte offsets inside one image are 32-bit ints in the kernels
    if ((unsigned long long)q.H * q.W * q.C * (acc_elem > sizeof(T) ? acc_elem : sizeof(T)) >= (1ull << 31)) return pl;
    // Candidate lane widths in order of preference.  8-byte lanes exist for the 16-bit backward with
    // fp32 accumulation, where they are the default: the 4 lanes of a group then cover the whole
    // 64-byte fp32 slab in ONE reduction instruction (one line-request on the SM->crossbar port
    // instead of two).  Everything else defaults to 16; DCNV3_B200_BPL forces one.
    const bool can8 = allow8 && sizeof(T) == 2;
    int want = forced_bpl();
    if (want == 8 && !can8) want = 0;
    int cand[2] = {16, 0};
    if (want) cand[0] = want;
    else if (can8) { cand[0] = 8; cand[1] = 16; }
    for (int bpl : cand) {
        if (!bpl) continue;
        const int ch = bpl / (int)sizeof(T);
        if (q.gc % ch) continue;
        const int L = q.gc / ch;
        if (!is_pow2(L) || L > 32) continue;
        bool ok = true;
        for (const void *p : ptrs) ok = ok && aligned_to(p, bpl);
        if (!ok) continue;
        const unsigned long long tv = (unsigned long long)n_pix * (q.C / ch);
        if (tv >= (1ull << 31)) continue;
        pl.vec = true;
        pl.bpl = bpl;
        pl.vec_per_pix = q.C / ch;
        pl.lanes_per_group = L;
        pl.total_vec = (unsigned)tv;
        return pl;
    }
    return pl;
}

unsigned blocks_for(size_t threads) { return (unsigned)((threads + kThreads - 1) / kThreads); }

// Launch with programmatic stream serialization (the kernel may be scheduled while its predecessor on
// the stream drains; it calls pdl_enter() before touching memory, dcnv3_common.cuh).  DCNV3_B200_PDL=0
// turns the attribute off (plain stream order).  Errors surface through cudaGetLastError in finish().
bool pdl_on() { return knobs().pdl; }
template <typename... P, typename... A>
void launch(void (*kernel)(P...), dim3 grid, unsigned block, size_t smem, cudaStream_t st, A... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at{};
    at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &at;
    cfg.numAttrs = pdl_on() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}

#ifndef DCNV3_NO_TMA
// the input [N, H, W, C] (16-bit) as a TMA tensor map: dims (C, W, H, N), box (64 channels, box_w columns, box_h rows, 1 image);
// coordinates outside the map are legal and read as zeros
// a 4-D tiled tensor map over 16-bit elements: dims / box innermost first, rows of dims[0] elements `pitch0` elements apart
int make_tmap4(const void *base, const cuuint64_t (&dims)[4], cuuint64_t pitch0, const cuuint32_t (&box)[4], int dtype, int swizzle,
               CUtensorMap *tm) {
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn enc = nullptr;  // benign race: idempotent
    if (!enc) {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) != cudaSuccess || !f) {
            cudaGetLastError();
            return fail(DCNV3_B200_EDEVICE, "cuTensorMapEncodeTiled is not available from this driver");
        }
        enc = (encode_fn)f;
    }
    const cuuint64_t strides[3] = {pitch0 * 2, dims[1] * pitch0 * 2, dims[2] * dims[1] * pitch0 * 2};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUtensorMapDataType dt = dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : dtype == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                                                            : CU_TENSOR_MAP_DATA_TYPE_UINT16;
    const CUresult r = enc(tm, dt, 4, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           swizzle == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : swizzle == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(DCNV3_B200_EINVAL, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}
int make_window_tmap(const void *in, const Geo &q, unsigned box_w, unsigned box_h, CUtensorMap *tm,
                     unsigned box_c = 64, int dtype = 0 /* 0: 16-bit words, 1: fp16, 2: bf16 (reductions) */, int swizzle = 0 /* 0, 32, 128 */) {
    const cuuint64_t dims[4] = {(cuuint64_t)q.C, (cuuint64_t)q.W, (cuuint64_t)q.H, (cuuint64_t)q.N};
    const cuuint32_t box[4] = {box_c, box_w, box_h, 1};
    return make_tmap4(in, dims, (cuuint64_t)q.C, box, dtype, swizzle, tm);
}
#endif

// ---------------------------------------------- interpolation-matrix family
// Knobs: DCNV3_B200_FWD / DCNV3_B200_BWD = vec | imat force a family.  Defaults: the backward takes the
// interpolation-matrix kernel when eligible (16-bit, gc = 16, 3x3 s1 d1, fp32 accumulation), the forward the
// vector kernel (its imat variant is correct but slower: profiles/r01_imat.md).
template <typename T>
bool imat_eligible(const Geo &q, std::initializer_list<const void *> vec_ptrs, const void *off) {
    if (sizeof(T) != 2) return false;
    if (q.gc != 16 || q.G % imat::kWarps) return false;
    if (q.kh != 3 || q.kw != 3 || q.sh != 1 || q.sw != 1 || q.dh != 1 || q.dw != 1) return false;
    for (const void *p : vec_ptrs) if (!aligned16(p)) return false;
    if (reinterpret_cast<uintptr_t>(off) & 3u) return false;
    const unsigned long long blocks = (unsigned long long)q.N * ((q.Ho + 7) / 8) * ((q.Wo + 7) / 8) * (q.G / imat::kWarps);
    return blocks > 0 && blocks < (1ull << 31);
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device), not once per launch
template <typename K> int set_smem(K kernel, int bytes, const char *what) {
    static std::mutex mu;
    static struct { const void *fn; unsigned long long devs; } seen[64];
    static int n_seen = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    const void *fn = reinterpret_cast<const void *>(kernel);
    {
        std::lock_guard<std::mutex> lk(mu);
        for (int i = 0; i < n_seen; ++i)
            if (seen[i].fn == fn && dev < 64 && ((seen[i].devs >> dev) & 1ull)) return 0;
    }
    const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return cuda_fail(e, what);
    if (dev < 64) {
        std::lock_guard<std::mutex> lk(mu);
        int i = 0;
        for (; i < n_seen; ++i) if (seen[i].fn == fn) break;
        if (i == n_seen && n_seen < 64) { seen[n_seen].fn = fn; seen[n_seen].devs = 0; ++n_seen; }
        if (i < 64) seen[i].devs |= 1ull << dev;
    }
    return 0;
}

template <typename T>
int launch_fwd_imat(const T *in, const T *off, const T *mask, T *out, const Geo &q, bool logits, cudaStream_t st) {
    const int tiles_y = (q.Ho + 7) / 8, tiles_x = (q.Wo + 7) / 8, GQ = q.G / imat::kWarps;
    const unsigned grid = (unsigned)((size_t)q.N * tiles_y * tiles_x * GQ);
    int rc;
    if (logits) {
        if ((rc = set_smem(imat::fwd_imat_kernel<T, true>, imat::kSmemFwd, "cudaFuncSetAttribute(fwd_imat_kernel)"))) return rc;
        launch(imat::fwd_imat_kernel<T, true>, grid, 32 * imat::kWarps, imat::kSmemFwd, st, in, off, mask, out, q, tiles_x, tiles_y, GQ);
    } else {
        if ((rc = set_smem(imat::fwd_imat_kernel<T, false>, imat::kSmemFwd, "cudaFuncSetAttribute(fwd_imat_kernel)"))) return rc;
        launch(imat::fwd_imat_kernel<T, false>, grid, 32 * imat::kWarps, imat::kSmemFwd, st, in, off, mask, out, q, tiles_x, tiles_y, GQ);
    }
    return 0;
}

template <typename T>
int launch_bwd_imat(const T *in, const T *off, const T *mask, const T *gout, float *acc, T *goff, T *gmask,
                    const Geo &q, bool logits, const int *sel, cudaStream_t st) {
    const int tiles_y = (q.Ho + 7) / 8, tiles_x = (q.Wo + 7) / 8, GQ = q.G / imat::kWarps;
    const unsigned grid = (unsigned)((size_t)q.N * tiles_y * tiles_x * GQ);
    int rc;
    if (logits) {
        if ((rc = set_smem(imat::bwd_imat_kernel<T, true>, imat::kSmemBwd, "cudaFuncSetAttribute(bwd_imat_kernel)"))) return rc;
        launch(imat::bwd_imat_kernel<T, true>, grid, 32 * imat::kWarps, imat::kSmemBwd, st, in, off, mask, gout, acc, goff, gmask, q, tiles_x, tiles_y, GQ, sel);
    } else {
        if ((rc = set_smem(imat::bwd_imat_kernel<T, false>, imat::kSmemBwd, "cudaFuncSetAttribute(bwd_imat_kernel)"))) return rc;
        launch(imat::bwd_imat_kernel<T, false>, grid, 32 * imat::kWarps, imat::kSmemBwd, st, in, off, mask, gout, acc, goff, gmask, q, tiles_x, tiles_y, GQ, sel);
    }
    return 0;
}

// ------------------------------------------------ single-kernel window backward (ACC_TILE)
bool win_geometry(const Geo &q) {
    return q.gc == 16 && q.G % imat::kWarps == 0 && q.kh == 3 && q.kw == 3 && q.sh == 1 && q.sw == 1 && q.dh == 1 && q.dw == 1;
}
template <typename T>
bool win_eligible(const Geo &q, const void *in, const void *off, const void *mask, const void *gout, const void *gin,
                  const void *goff, const void *gmask) {
    if (sizeof(T) != 2 || !win_geometry(q)) return false;
    if (!aligned16(in) || !aligned16(gout) || !aligned16(gin)) return false;
    // offsets / masks move as 16- / 8-byte chunks (4 groups of a pixel: 144 / 72 contiguous bytes)
    if (!aligned16(off) || !aligned16(goff) || (reinterpret_cast<uintptr_t>(mask) & 7u) ||
        (reinterpret_cast<uintptr_t>(gmask) & 7u)) return false;
    if ((q.opitch * 2) % 16 || (q.mpitch * 2) % 8) return false;  // packed heads: every pixel's run starts on a chunk boundary
    if (q.N > 65535 || (q.Ho + 3) / 4 > 65535) return false;      // grid.z / grid.y
    const unsigned long long blocks = (unsigned long long)q.N * ((q.Ho + 3) / 4) * ((q.Wo + 7) / 8) * (q.G / imat::kWarps);
    if (blocks == 0 || blocks >= (1ull << 30)) return false;
    // the far-band fallback indexes (pixel, 8-channel vector) lanes with 32 bits
    return (unsigned long long)q.N * q.Ho * q.Wo * (q.C / 8) < (1ull << 31);
}

// one launch, one CTA per unit (image, 4-row band, 8-column tile, group quad): grad_offset / grad_mask, then grad_input
// over the same shared memory (dcnv3_win.cuh)
template <typename T>
int launch_bwd_win(const T *in, const T *off, const T *mask, const T *gout, T *gin, T *goff, T *gmask,
                   const Geo &q, bool logits, cudaStream_t st) {
    int tiles_x = (q.Wo + 7) / 8;
    const int bands_y = (q.Ho + 3) / 4, GQ = q.G / imat::kWarps;
    // 1-4 columns beyond the last whole tile column: a strip of transposed tiles (8 rows x 4 columns) instead of a third,
    // half-empty tile column (dcnv3_win.cuh, STRIP).  They take the grid rows behind the bands.
    int strip_tiles = 0, strip_rows = 0;
#if defined(DCNV3_WIN_TMA) && !defined(DCNV3_WIN_GRID1D) && !defined(DCNV3_WIN_NO_STRIP)
    if (q.Wo >= 8 && q.Wo % 8 >= 1 && q.Wo % 8 <= 4) {
        const int tx_n = q.Wo / 8, nt = (q.Ho + 7) / 8, rows = (nt + tx_n - 1) / tx_n;
        if (bands_y + rows <= 65535) { tiles_x = tx_n; strip_tiles = nt; strip_rows = rows; }
    }
#endif
#ifdef DCNV3_WIN_GRID1D
    const dim3 grid((unsigned)((size_t)q.N * bands_y * tiles_x * GQ));
#else
    const dim3 grid((unsigned)(tiles_x * GQ), (unsigned)(bands_y + strip_rows), (unsigned)q.N);
#endif
    int rc, stage_tma = 0;
#ifdef DCNV3_WIN_TMA
    alignas(64) CUtensorMap tm, tms;
    if ((rc = make_window_tmap(in, q, 16, 12, &tm))) return rc;
    if (strip_tiles) { if ((rc = make_window_tmap(in, q, 12, 16, &tms))) return rc; }
    else tms = tm;
    // offsets / masks of a tile as two more boxes (80 / 40 elements = the staging area's pixel pitch x 8 columns x 4 rows) when
    // their rows are 16-byte multiples (G a multiple of 8 with unpacked heads); otherwise the warps' own cp.async chunks
    alignas(64) CUtensorMap tmo, tmm;
#if DCNV3_WIN_TMA_STAGE
    if ((q.opitch * 2) % 16 == 0 && (q.mpitch * 2) % 16 == 0 && aligned16(mask)) {
        const cuuint64_t od[4] = {(cuuint64_t)q.opitch, (cuuint64_t)q.Wo, (cuuint64_t)q.Ho, (cuuint64_t)q.N};
        const cuuint64_t md[4] = {(cuuint64_t)q.mpitch, (cuuint64_t)q.Wo, (cuuint64_t)q.Ho, (cuuint64_t)q.N};
        const cuuint32_t ob[4] = {win::kStOffPx / 2, 8, 4, 1}, mb[4] = {win::kStMaskPx / 2, 8, 4, 1};
        // (a shape the encoder refuses simply keeps the warps' own chunks)
        stage_tma = !make_tmap4(off, od, (cuuint64_t)q.opitch, ob, 0, 0, &tmo) && !make_tmap4(mask, md, (cuuint64_t)q.mpitch, mb, 0, 0, &tmm);
    }
#endif
    if (!stage_tma) { tmo = tm; tmm = tm; }
#if DCNV3_WIN_TMA_FLUSH
    alignas(64) CUtensorMap tmr;
    if ((rc = make_window_tmap(gin, q, 16, 6, &tmr, DCNV3_WIN_TMA_FLUSH == 2 ? 64 : 16, std::is_same<T, __half>::value ? 1 : 2,
                               DCNV3_WIN_TMA_FLUSH == 2 ? 128 : 32))) return rc;
#define WIN_EXTRA , tm, tms, tmo, tmm, tmr
#else
#define WIN_EXTRA , tm, tms, tmo, tmm
#endif
#else
#define WIN_EXTRA
#endif
#define WIN_LAUNCH(LG, SP)                                                                                              \
    do {                                                                                                                \
        if ((rc = set_smem(win::bwd_win_kernel<T, LG, SP>, win::kSmemB, "cudaFuncSetAttribute(bwd_win_kernel)"))) return rc; \
        launch(win::bwd_win_kernel<T, LG, SP>, grid, win::kThreadsW, win::kSmemB, st, in, off, mask, gout, gin, goff, gmask, q, \
               GQ, tiles_x, bands_y WIN_EXTRA, strip_tiles, stage_tma);                                                            \
    } while (0)
#ifdef DCNV3_WIN_TMA
    if (strip_tiles) {
        if (logits) WIN_LAUNCH(true, true); else WIN_LAUNCH(false, true);
        return 0;
    }
#endif
    if (logits) WIN_LAUNCH(true, false); else WIN_LAUNCH(false, false);
#undef WIN_LAUNCH
#undef WIN_EXTRA
    return 0;
}

// ------------------------------------------------------------------ forward
template <typename T>
int forward_t(const void *in_, const void *off_, const void *mask_, void *out_, const Geo &q,
              bool logits, cudaStream_t st, bool packed = false) {
    const T *in = (const T *)in_, *off = (const T *)off_, *mask = (const T *)mask_;
    T *out = (T *)out_;
    const size_t n_pix = (size_t)q.N * q.Ho * q.Wo;
    if (n_pix == 0) return 0;
    if constexpr (sizeof(T) == 2) {
        const int fam = knobs().fwd;
        if (fam == 3 && !packed && imat_eligible<T>(q, {in_, out_}, off_))
            return launch_fwd_imat<T>(in, off, mask, out, q, logits, st);
        // Default for the C3-DCN shapes (16-bit, group_channels = 16, 3x3 s1 d1, G % 4 = 0): the forward from a
        // staged window (imat::fwd_tile_kernel), unless the map wastes more than 40 % of its 8x8 tiles.
        // DCNV3_B200_FWD=vec forces the vector kernel, =win this one.
        const long long tiles64 = 64ll * ((q.Ho + 7) / 8) * ((q.Wo + 7) / 8);
        if ((packed || fam == 5 || (fam == 0 && 10ll * q.Ho * q.Wo >= 6ll * tiles64)) && imat_eligible<T>(q, {in_, out_}, off_) &&
            !(reinterpret_cast<uintptr_t>(mask_) & 1u) && q.N <= 65535 && (q.Ho + 7) / 8 <= 65535) {
            const int tiles_y = (q.Ho + 7) / 8, tiles_x = (q.Wo + 7) / 8, GQ = q.G / imat::kWarps;
            const dim3 grid((unsigned)(tiles_x * GQ), (unsigned)tiles_y, (unsigned)q.N);
            int rc;
#ifdef DCNV3_FWD_TMA
            alignas(64) CUtensorMap tm;
            if ((rc = make_window_tmap(in, q, imat::kFwin, imat::kFwin, &tm))) return rc;
#if DCNV3_FWD_TMA_STAGE
            alignas(64) CUtensorMap tmo = tm, tmm = tm;
            int stage_tma = 0;
            if ((q.opitch * 2) % 16 == 0 && (q.mpitch * 2) % 16 == 0 && aligned16(off_) && aligned16(mask_)) {
                const cuuint64_t od[4] = {(cuuint64_t)q.opitch, (cuuint64_t)q.Wo, (cuuint64_t)q.Ho, (cuuint64_t)q.N};
                const cuuint64_t md[4] = {(cuuint64_t)q.mpitch, (cuuint64_t)q.Wo, (cuuint64_t)q.Ho, (cuuint64_t)q.N};
                const cuuint32_t ob[4] = {imat::kFstOffPx / 2, 8, 8, 1}, mb[4] = {imat::kFstMaskPx / 2, 8, 8, 1};
                // (a shape the encoder refuses simply keeps the lanes' own loads)
                stage_tma = !make_tmap4(off, od, (cuuint64_t)q.opitch, ob, 0, 0, &tmo) && !make_tmap4(mask, md, (cuuint64_t)q.mpitch, mb, 0, 0, &tmm);
                if (!stage_tma) { tmo = tm; tmm = tm; }
            }
#define FWD_EXTRA , tm, tmo, tmm, stage_tma
#else
#define FWD_EXTRA , tm
#endif
#else
#define FWD_EXTRA
#endif
#define FWD_LAUNCH(LG, SG)                                                                                                    \
    do {                                                                                                                      \
        if ((rc = set_smem(imat::fwd_tile_kernel<T, LG, SG>, imat::kFwdSmemB, "cudaFuncSetAttribute(fwd_tile_kernel)"))) return rc; \
        launch(imat::fwd_tile_kernel<T, LG, SG>, grid, imat::kFwdTileThreads, imat::kFwdSmemB, st, in, off, mask, out, q, GQ FWD_EXTRA); \
    } while (0)
#if defined(DCNV3_FWD_TMA) && DCNV3_FWD_TMA_STAGE
            if (stage_tma) { if (logits) FWD_LAUNCH(true, true); else FWD_LAUNCH(false, true); }
            else
#endif
            { if (logits) FWD_LAUNCH(true, false); else FWD_LAUNCH(false, false); }
#undef FWD_LAUNCH
#undef FWD_EXTRA
            return 0;
        }
        if (packed) return fail(DCNV3_B200_ENOTSUP, "packed heads: this shape / alignment does not take the staged-window forward");
        // opt-in (DCNV3_B200_FWD=pts): the point-split kernel — 25 % fewer instructions, measured SLOWER
        // (P3 98.7 vs 82.3 us): a lane's 32-byte LDG.256 costs L1 two passes where the channel-split
        // pair of 16-byte lanes shares one; the forward is bound by L1 wavefronts, not by issue slots
        const unsigned long long lanes = (unsigned long long)n_pix * q.G * 2ull;
        if (fam == 4 && q.gc == 16 && q.kh == 3 && q.kw == 3 && aligned_to(in_, 32) && aligned16(out_) &&
            !(reinterpret_cast<uintptr_t>(off_) & 3u) && lanes < (1ull << 31) &&
            (unsigned long long)q.H * q.W * q.C * 2ull < (1ull << 31)) {
            const unsigned total = (unsigned)lanes;
            if (logits) launch(fwd_pts_kernel<T, true>, blocks_for(total), kThreads, 0, st, in, off, mask, out, q, total);
            else launch(fwd_pts_kernel<T, false>, blocks_for(total), kThreads, 0, st, in, off, mask, out, q, total);
            return 0;
        }
    }
    if (packed) return fail(DCNV3_B200_ENOTSUP, "packed heads need 16-bit storage");
    const Plan pl = plan_vec<T>(q, n_pix, logits, {in_, out_}, off_, sizeof(T));
    if constexpr (sizeof(T) <= 4) {
        if (pl.vec) {
            const unsigned grid = blocks_for(pl.total_vec);
            const bool k9 = (q.kh == 3 && q.kw == 3);
#define LAUNCH_FWD(BPL, KP, LG)                                                               \
    launch(fwd_vec_kernel<T, BPL, KP, LG>, grid, kThreads, 0, st, in, off, mask, out, q,      \
           pl.vec_per_pix, pl.lanes_per_group, pl.total_vec)
#define LAUNCH_FWD_B(BPL)                      \
    if (k9 && logits) LAUNCH_FWD(BPL, 9, true); \
    else if (k9) LAUNCH_FWD(BPL, 9, false);     \
    else LAUNCH_FWD(BPL, 0, false)
            if (pl.bpl == 32) { LAUNCH_FWD_B(32); } else { LAUNCH_FWD_B(16); }
#undef LAUNCH_FWD_B
#undef LAUNCH_FWD
            return 0;
        }
    }
    const size_t total = n_pix * q.C;
    if (blocks_for(total) == 0 || total / kThreads >= (1ull << 31))
        return fail(DCNV3_B200_ERANGE, "output too large for one launch");
    if (logits) fwd_any_kernel<T, true><<<blocks_for(total), kThreads, 0, st>>>(in, off, mask, out, q, total);
    else fwd_any_kernel<T, false><<<blocks_for(total), kThreads, 0, st>>>(in, off, mask, out, q, total);
    return 0;
}

// ------------------------------------------------ privatised backward (tile)
// Knobs: DCNV3_B200_BWD=tile selects this family, DCNV3_B200_TILE="R,warps" its halo / warps per CTA.
struct TileKnobs { int force; int R, warps; };
TileKnobs tile_knobs() {
    const Knobs &k = knobs();
    return TileKnobs{k.bwd == 1 ? 1 : (k.bwd == 2 ? 2 : 0), k.tile_R, k.tile_warps};
}

template <typename T>
bool plan_tile(const Geo &q, bool logits, const void *in, const void *gout, const void *off,
               const void *acc, TileCfg &tc) {
    // Experimental (round 1): only runs when asked for with DCNV3_B200_BWD=tile.
    const TileKnobs kn = tile_knobs();
    if (kn.force != 2) return false;
    if (sizeof(T) > 4 || q.gc != kTileGC) return false;
    if (logits && q.P > kMaxSoftmaxP) return false;
    if (!aligned_to(in, 8 * (int)sizeof(T)) || !aligned_to(gout, 8 * (int)sizeof(T)) || !aligned16(acc)) return false;
    if (reinterpret_cast<uintptr_t>(off) & (2 * sizeof(T) - 1)) return false;
    if ((unsigned long long)q.H * q.W * q.C * 4ull >= (1ull << 31)) return false;
    const float s = fabsf(q.scale);
    if (!(s < 64.f)) return false;
    tc.R = kn.R;
    // footprint of the un-offset taps around (ho*stride - pad): [half - s*half, half + s*(dil*(k-1) - half)]
    const int lo_h = (int)ceilf(s * q.half_h), hi_h = (int)ceilf(s * (q.dh * (q.kh - 1) - q.half_h));
    const int lo_w = (int)ceilf(s * q.half_w), hi_w = (int)ceilf(s * (q.dw * (q.kw - 1) - q.half_w));
    tc.oy = q.half_h - q.ph - lo_h - tc.R;
    tc.ox = q.half_w - q.pw - lo_w - tc.R;
    tc.WH = (kTileT - 1) * q.sh + lo_h + hi_h + 2 + 2 * tc.R;
    tc.WW = (kTileT - 1) * q.sw + lo_w + hi_w + 2 + 2 * tc.R;
    tc.tiles_y = (q.Ho + kTileT - 1) / kTileT;
    tc.tiles_x = (q.Wo + kTileT - 1) / kTileT;
    tc.warps = q.G < kn.warps ? q.G : kn.warps;
    size_t per = (size_t)tc.WH * tc.WW * (64 + 4);
    per = (per + 127) & ~(size_t)127;
    if (per > 56 * 1024 || per * tc.warps > 220 * 1024) return false;
    tc.smem_per_warp = (int)per;
    const unsigned long long blocks = (unsigned long long)q.N * tc.tiles_y * tc.tiles_x * ((q.G + tc.warps - 1) / tc.warps);
    if (blocks >= (1ull << 31)) return false;
    return true;
}

template <typename T>
int launch_tile(const T *in, const T *off, const T *mask, const T *gout, float *acc, T *goff, T *gmask,
                const Geo &q, bool logits, const TileCfg &tc, cudaStream_t st) {
    const unsigned grid = (unsigned)((size_t)q.N * tc.tiles_y * tc.tiles_x * ((q.G + tc.warps - 1) / tc.warps));
    const size_t smem = (size_t)tc.smem_per_warp * tc.warps;
    cudaError_t e;
#define LAUNCH_TILE(KP, LG)                                                                              \
    do {                                                                                                 \
        if ((e = cudaFuncSetAttribute(bwd_tile_kernel<T, KP, LG>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      (int)smem)) != cudaSuccess)                                        \
            return cuda_fail(e, "cudaFuncSetAttribute(bwd_tile_kernel)");                                \
        bwd_tile_kernel<T, KP, LG><<<grid, 32 * tc.warps, smem, st>>>(in, off, mask, gout, acc, goff, gmask, q, tc); \
    } while (0)
    const bool k9 = (q.kh == 3 && q.kw == 3);
    if (k9 && logits) LAUNCH_TILE(9, true);
    else if (k9) LAUNCH_TILE(9, false);
    else if (logits) LAUNCH_TILE(0, true);
    else LAUNCH_TILE(0, false);
#undef LAUNCH_TILE
    return 0;
}

// ----------------------------------------------------------------- backward
template <typename T, typename A>
int backward_launch(const T *in, const T *off, const T *mask, const T *gout, A *acc, T *goff,
                    T *gmask, const Geo &q, bool logits, const Plan &pl, size_t n_pix,
                    cudaStream_t st, const int *sel = nullptr) {
    if constexpr (sizeof(T) <= 4) {
        if (pl.vec) {
            const unsigned n_blocks = blocks_for(pl.total_vec);
            // Selector-guarded launch: `per_cta` consecutive logical blocks per CTA, so that a launch the
            // selector turns into a no-op is a few thousand empty CTAs, not 12 800.  Measured on one box
            // (step of the common path / P3 backward when the selector picks THIS kernel, us):
            //   per_cta 1: 575.0 / 432   2: 563.4 / 440   4: 557.6 / 456   8: 556 / 464   32: 553.1 / 514
            // (fewer, longer CTAs quantise the last wave).  4 is the default; DCNV3_B200_GUARD_PER_CTA overrides.
            unsigned per_cta = 1u;
            if (sel) per_cta = knobs().guard_per_cta;
            const unsigned grid = (n_blocks + per_cta - 1) / per_cta;
            const bool k9 = (q.kh == 3 && q.kw == 3);
#define LAUNCH_BWD(BPL, KP, LG)                                                                  \
    launch(bwd_vec_kernel<T, A, BPL, KP, LG>, grid, kThreads, 0, st, in, off, mask, gout, acc, goff, \
           gmask, q, pl.vec_per_pix, pl.lanes_per_group, pl.total_vec, sel, per_cta, n_blocks)
#define LAUNCH_BWD_B(BPL)                        \
    if (k9 && logits) { LAUNCH_BWD(BPL, 9, true); } \
    else if (k9) { LAUNCH_BWD(BPL, 9, false); }     \
    else { LAUNCH_BWD(BPL, 0, false); }
            if (pl.bpl == 32) { LAUNCH_BWD_B(32); }
            else if (pl.bpl == 8) {
                if constexpr (sizeof(T) == 2 && sizeof(A) == 4) { LAUNCH_BWD_B(8); }
                else return fail(DCNV3_B200_EINVAL, "8-byte lanes need 16-bit storage with fp32 accumulation");
            } else { LAUNCH_BWD_B(16); }
#undef LAUNCH_BWD_B
#undef LAUNCH_BWD
            return 0;
        }
    }
    const size_t n_units = n_pix * q.G;
    int lpu = 1;  // lanes per (pixel, group): next power of two >= min(group_channels, 32)
    while (lpu < q.gc && lpu < 32) lpu <<= 1;
    const size_t threads = n_units * lpu;
    if (threads / kThreads >= (1ull << 31))
        return fail(DCNV3_B200_ERANGE, "too many (pixel, group) units for one launch");
    if (logits) {
        if (q.P > kMaxSoftmaxP)
            return fail(DCNV3_B200_EINVAL, "fused softmax supports at most %d sampling points (got %d)",
                        kMaxSoftmaxP, q.P);
        bwd_any_kernel<T, A, true><<<blocks_for(threads), kThreads, 0, st>>>(in, off, mask, gout, acc, goff, gmask, q, n_units, lpu);
    } else {
        bwd_any_kernel<T, A, false><<<blocks_for(threads), kThreads, 0, st>>>(in, off, mask, gout, acc, goff, gmask, q, n_units, lpu);
    }
    return 0;
}

template <typename T>
int backward_t(const void *in_, const void *off_, const void *mask_, const void *gout_,
               void *gin_, void *goff_, void *gmask_, void *ws, size_t ws_bytes, const Geo &q,
               bool logits, int grad_accum, cudaStream_t st, bool packed = false) {
    using M = typename OpMath<T>::type;
    const T *in = (const T *)in_, *off = (const T *)off_, *mask = (const T *)mask_;
    const T *gout = (const T *)gout_;
    T *gin = (T *)gin_, *goff = (T *)goff_, *gmask = (T *)gmask_;
    const size_t n_in = (size_t)q.N * q.H * q.W * q.C;
    const size_t n_pix = (size_t)q.N * q.Ho * q.Wo;
    if (n_in == 0) return 0;
    cudaError_t e;
    constexpr bool lowp = sizeof(T) == 2;
    if constexpr (lowp) {
        // ACC_TILE (default of the Python front): one kernel, fp32 sums per tile window, 16-bit reductions across
        // tiles, no workspace.  Shapes it does not take fall through to ACC_OPMATH (which needs the workspace).
        if (grad_accum == DCNV3_B200_ACC_TILE) {
            const int fam = knobs().bwd;
            if ((packed || fam == 0 || fam == 5) && win_eligible<T>(q, in_, off_, mask_, gout_, gin_, goff_, gmask_)) {
                launch(win::zero_fill_kernel, 148 * 4, 256, 0, st, reinterpret_cast<uint4 *>(gin), n_in * sizeof(T) / 16);
                if (n_pix == 0) return 0;
                return launch_bwd_win<T>(in, off, mask, gout, gin, goff, gmask, q, logits, st);
            }
            if (packed) return fail(DCNV3_B200_ENOTSUP, "packed heads: this shape / alignment does not take the window backward");
            if (win_geometry(q) && (!ws || ws_bytes == 0))
                return fail(DCNV3_B200_EALIGN, "ACC_TILE needs 16-byte aligned input / grad_output / grad_input / offset / "
                                               "grad_offset (8-byte mask / grad_mask); pass an ACC_OPMATH workspace to fall back");
            grad_accum = DCNV3_B200_ACC_OPMATH;
        }
    }
    if (packed) return fail(DCNV3_B200_ENOTSUP, "packed heads need 16-bit storage and grad_accum = ACC_TILE");
    if (lowp && grad_accum == DCNV3_B200_ACC_OPMATH) {
        // reference semantics (dcnv3_cuda.cu:126-133,168-170): fp32 accumulation, one rounding
        const size_t acc_bytes = (n_in * sizeof(float) + 255) & ~(size_t)255;
        const size_t need = acc_bytes + 256;  // + the family selector word (imat::select_kernel)
        if (!ws || ws_bytes < need)
            return fail(DCNV3_B200_EWORKSPACE, "workspace of %zu bytes required, got %zu", need, ws ? ws_bytes : 0);
        if (!aligned16(ws)) return fail(DCNV3_B200_EALIGN, "workspace must be 16-byte aligned");
        float *acc = (float *)ws;
        bool zeroed = false;
        if (n_pix) {
            TileCfg tc;
            int rc = 0;
            bool tiled = false;
            if constexpr (lowp) {
                if (plan_tile<T>(q, logits, in_, gout_, off_, acc, tc)) {
                    if ((e = cudaMemsetAsync(acc, 0, need, st)) != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(workspace)");
                    zeroed = true;
                    rc = launch_tile<T>(in, off, mask, gout, acc, goff, gmask, q, logits, tc, st);
                    tiled = true;
                }
            }
            if constexpr (lowp) {
                // Default for eligible shapes: a 2 K-sample look at the offsets picks the family on
                // the device (imat::select_kernel); both kernels are launched, one returns at once.
                // DCNV3_B200_BWD=vec|imat forces one (no selector).
                const int fam = knobs().bwd;
                if (!tiled && (fam == 0 || fam == 3) && imat_eligible<T>(q, {in_, gout_, ws}, off_) &&
                    !(reinterpret_cast<uintptr_t>(goff_) & 3u)) {
                    const Plan pl = plan_vec<T>(q, n_pix, logits, {in_, gout_}, off_, sizeof(float), true);
                    int *sel = nullptr;
                    if (fam == 0 && pl.vec) sel = reinterpret_cast<int *>(static_cast<char *>(ws) + acc_bytes);
                    // zero fill of the accumulators (the memset) with the selector riding along in block 0
                    launch(imat::zero_select_kernel<T>, 148 * 4, imat::kSelThreads, 0, st,
                           reinterpret_cast<uint4 *>(acc), acc_bytes / 16, off, (unsigned long long)n_pix * q.G * q.P, q.scale, sel);
                    zeroed = true;
                    rc = launch_bwd_imat<T>(in, off, mask, gout, acc, goff, gmask, q, logits, sel, st);
                    if (!rc && sel) rc = backward_launch<T, float>(in, off, mask, gout, acc, goff, gmask, q, logits, pl, n_pix, st, sel);
                    tiled = true;
                }
            }
            if (!tiled) {
                if ((e = cudaMemsetAsync(acc, 0, need, st)) != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(workspace)");
                zeroed = true;
                const Plan pl = plan_vec<T>(q, n_pix, logits, {in_, gout_}, off_, sizeof(float), true);
                rc = backward_launch<T, float>(in, off, mask, gout, acc, goff, gmask, q, logits, pl, n_pix, st);
            }
            if (rc) return rc;
        }
        if (!zeroed && (e = cudaMemsetAsync(acc, 0, need, st)) != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(workspace)");
        if constexpr (lowp) {
            const bool v8 = aligned16(gin_);
            const size_t n8 = v8 ? n_in / 8 : 0;
            launch(cast_ws_kernel<T>, blocks_for(n8 + 1), kThreads, 0, st, (const float *)acc, gin, n8, n_in);
        }
        return 0;
    }
    // accumulate straight into grad_input (f32/f64 storage, or 16-bit ACC_STORAGE)
    if ((e = cudaMemsetAsync(gin, 0, n_in * sizeof(T), st)) != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(grad_input)");
    if (n_pix == 0) return 0;
    const Plan pl = aligned16(gin_) ? plan_vec<T>(q, n_pix, logits, {in_, gout_}, off_, sizeof(T)) : Plan{false, 0, 0, 0, 0};
    if constexpr (lowp) {
        return backward_launch<T, T>(in, off, mask, gout, gin, goff, gmask, q, logits, pl, n_pix, st);
    } else {
        if constexpr (sizeof(T) == 4) {
            TileCfg tc;
            if (plan_tile<T>(q, logits, in_, gout_, off_, gin_, tc))
                return launch_tile<T>(in, off, mask, gout, (float *)gin, goff, gmask, q, logits, tc, st);
        }
        return backward_launch<T, M>(in, off, mask, gout, (M *)gin, goff, gmask, q, logits, pl, n_pix, st);
    }
}

int finish(cudaStream_t st, const char *what) {
    (void)st;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, what);
    return 0;
}

}  // namespace

extern "C" {

int dcnv3_b200_version(void) { return DCNV3_B200_VERSION; }

void dcnv3_b200_reload_knobs(void) { read_knobs(); }

const char *dcnv3_b200_last_error(void) { return g_err; }

int dcnv3_b200_output_size(const dcnv3_b200_geometry *geo, int *Ho, int *Wo) {
    Geo q;
    int rc = make_geo(geo, q);
    if (rc) return rc;
    if (Ho) *Ho = q.Ho;
    if (Wo) *Wo = q.Wo;
    return 0;
}

int dcnv3_b200_forward(const void *input, const void *offset, const void *mask, void *output,
                       int dtype, const dcnv3_b200_geometry *geo, int mask_is_logits,
                       void *cuda_stream) {
    g_err[0] = 0;
    Geo q;
    int rc = make_geo(geo, q);
    if (rc) return rc;
    if (!dtype_size(dtype)) return fail(DCNV3_B200_EINVAL, "unknown dtype %d", dtype);
    if (mask_is_logits != 0 && mask_is_logits != 1) return fail(DCNV3_B200_EINVAL, "mask_is_logits must be 0 or 1");
    // same bound as the backward (backward_launch): a fused-softmax forward must not succeed where autograd's
    // backward would then fail
    if (mask_is_logits && q.P > kMaxSoftmaxP)
        return fail(DCNV3_B200_EINVAL, "fused softmax supports at most %d sampling points (got %d)", kMaxSoftmaxP, q.P);
    if (q.N == 0) return 0;
    if (!input || !offset || !mask || !output) return fail(DCNV3_B200_ENULL, "null tensor pointer");
    if ((rc = check_device())) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const bool lg = mask_is_logits != 0;
    switch (dtype) {
        case DCNV3_B200_F32: rc = forward_t<float>(input, offset, mask, output, q, lg, st); break;
        case DCNV3_B200_F16: rc = forward_t<__half>(input, offset, mask, output, q, lg, st); break;
        case DCNV3_B200_BF16: rc = forward_t<__nv_bfloat16>(input, offset, mask, output, q, lg, st); break;
        default: rc = forward_t<double>(input, offset, mask, output, q, lg, st); break;
    }
    if (rc) return rc;
    return finish(st, "dcnv3_b200_forward launch");
}

// offsets and mask (logits) of a pixel side by side in ONE tensor [N, Ho, Wo, 3*G*P] — the output of a single Linear
int dcnv3_b200_forward_packed(const void *input, const void *heads, void *output, int dtype,
                              const dcnv3_b200_geometry *geo, int mask_is_logits, void *cuda_stream) {
    g_err[0] = 0;
    Geo q;
    int rc = make_geo(geo, q);
    if (rc) return rc;
    if (dtype != DCNV3_B200_F16 && dtype != DCNV3_B200_BF16) return fail(DCNV3_B200_ENOTSUP, "packed heads need 16-bit storage");
    if (mask_is_logits != 0 && mask_is_logits != 1) return fail(DCNV3_B200_EINVAL, "mask_is_logits must be 0 or 1");
    if (q.N == 0) return 0;
    if (!input || !heads || !output) return fail(DCNV3_B200_ENULL, "null tensor pointer");
    if ((rc = check_device())) return rc;
    q.opitch = q.mpitch = 3 * q.G * q.P;
    const void *mask = static_cast<const char *>(heads) + (size_t)2 * q.G * q.P * 2;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    rc = dtype == DCNV3_B200_F16 ? forward_t<__half>(input, heads, mask, output, q, mask_is_logits != 0, st, true)
                                 : forward_t<__nv_bfloat16>(input, heads, mask, output, q, mask_is_logits != 0, st, true);
    if (rc) return rc;
    return finish(st, "dcnv3_b200_forward_packed launch");
}

int dcnv3_b200_backward_packed(const void *input, const void *heads, const void *grad_output, void *grad_input,
                               void *grad_heads, int dtype, const dcnv3_b200_geometry *geo, int mask_is_logits,
                               void *cuda_stream) {
    g_err[0] = 0;
    Geo q;
    int rc = make_geo(geo, q);
    if (rc) return rc;
    if (dtype != DCNV3_B200_F16 && dtype != DCNV3_B200_BF16) return fail(DCNV3_B200_ENOTSUP, "packed heads need 16-bit storage");
    if (mask_is_logits != 0 && mask_is_logits != 1) return fail(DCNV3_B200_EINVAL, "mask_is_logits must be 0 or 1");
    if (q.N == 0) return 0;
    if (!input || !heads || !grad_output || !grad_input || !grad_heads) return fail(DCNV3_B200_ENULL, "null tensor pointer");
    if ((rc = check_device())) return rc;
    q.opitch = q.mpitch = 3 * q.G * q.P;
    const size_t mo = (size_t)2 * q.G * q.P * 2;
    const void *mask = static_cast<const char *>(heads) + mo;
    void *gmask = static_cast<char *>(grad_heads) + mo;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const bool lg = mask_is_logits != 0;
    rc = dtype == DCNV3_B200_F16
             ? backward_t<__half>(input, heads, mask, grad_output, grad_input, grad_heads, gmask, nullptr, 0, q, lg, DCNV3_B200_ACC_TILE, st, true)
             : backward_t<__nv_bfloat16>(input, heads, mask, grad_output, grad_input, grad_heads, gmask, nullptr, 0, q, lg, DCNV3_B200_ACC_TILE, st, true);
    if (rc) return rc;
    return finish(st, "dcnv3_b200_backward_packed launch");
}

size_t dcnv3_b200_backward_workspace_bytes(int dtype, const dcnv3_b200_geometry *geo, int grad_accum) {
    Geo q;
    if (make_geo(geo, q)) return 0;
    const bool tile_takes_it = grad_accum == DCNV3_B200_ACC_TILE && win_geometry(q) && knobs().bwd != 1 &&
                               knobs().bwd != 2 && knobs().bwd != 3;
    if (dtype_size(dtype) == 2 && !tile_takes_it &&
        (grad_accum == DCNV3_B200_ACC_OPMATH || grad_accum == DCNV3_B200_ACC_TILE))
        return (((size_t)q.N * q.H * q.W * q.C * sizeof(float) + 255) & ~(size_t)255) + 256;
    return 0;
}

int dcnv3_b200_backward(const void *input, const void *offset, const void *mask,
                        const void *grad_output, void *grad_input, void *grad_offset,
                        void *grad_mask, void *workspace, size_t workspace_bytes, int dtype,
                        const dcnv3_b200_geometry *geo, int mask_is_logits, int grad_accum,
                        void *cuda_stream) {
    g_err[0] = 0;
    Geo q;
    int rc = make_geo(geo, q);
    if (rc) return rc;
    if (!dtype_size(dtype)) return fail(DCNV3_B200_EINVAL, "unknown dtype %d", dtype);
    if (mask_is_logits != 0 && mask_is_logits != 1) return fail(DCNV3_B200_EINVAL, "mask_is_logits must be 0 or 1");
    if (grad_accum != DCNV3_B200_ACC_OPMATH && grad_accum != DCNV3_B200_ACC_STORAGE && grad_accum != DCNV3_B200_ACC_TILE)
        return fail(DCNV3_B200_EINVAL, "unknown grad_accum %d", grad_accum);
    if (q.N == 0) return 0;
    if (!input || !offset || !mask || !grad_output || !grad_input || !grad_offset || !grad_mask)
        return fail(DCNV3_B200_ENULL, "null tensor pointer");
    if ((rc = check_device())) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const bool lg = mask_is_logits != 0;
#define BWD(T) backward_t<T>(input, offset, mask, grad_output, grad_input, grad_offset, grad_mask, \
                             workspace, workspace_bytes, q, lg, grad_accum, st)
    switch (dtype) {
        case DCNV3_B200_F32: rc = BWD(float); break;
        case DCNV3_B200_F16: rc = BWD(__half); break;
        case DCNV3_B200_BF16: rc = BWD(__nv_bfloat16); break;
        default: rc = BWD(double); break;
    }
#undef BWD
    if (rc) return rc;
    return finish(st, "dcnv3_b200_backward launch");
}

int dcnv3_b200_debug_indices(const void *offset, int32_t *hw_low, uint8_t *bounds, int dtype,
                             const dcnv3_b200_geometry *geo, void *cuda_stream) {
    g_err[0] = 0;
    Geo q;
    int rc = make_geo(geo, q);
    if (rc) return rc;
    if (!dtype_size(dtype)) return fail(DCNV3_B200_EINVAL, "unknown dtype %d", dtype);
    if (q.N == 0) return 0;
    if (!offset || !hw_low || !bounds) return fail(DCNV3_B200_ENULL, "null tensor pointer");
    if ((rc = check_device())) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t total = (size_t)q.N * q.Ho * q.Wo * q.G * q.P;
    const unsigned grid = blocks_for(total);
    switch (dtype) {
        case DCNV3_B200_F32: indices_kernel<float><<<grid, kThreads, 0, st>>>((const float *)offset, hw_low, bounds, q, total); break;
        case DCNV3_B200_F16: indices_kernel<__half><<<grid, kThreads, 0, st>>>((const __half *)offset, hw_low, bounds, q, total); break;
        case DCNV3_B200_BF16: indices_kernel<__nv_bfloat16><<<grid, kThreads, 0, st>>>((const __nv_bfloat16 *)offset, hw_low, bounds, q, total); break;
        default: indices_kernel<double><<<grid, kThreads, 0, st>>>((const double *)offset, hw_low, bounds, q, total); break;
    }
    return finish(st, "dcnv3_b200_debug_indices launch");
}

}  // extern "C"

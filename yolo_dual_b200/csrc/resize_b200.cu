// NHWC nearest / bilinear resize, forward and gather backward (include/resize_b200.h).  HBM streams on the large
// side of the resize; the small side is re-read from L1 / L2.  One thread per 16-byte channel vector of one pixel.
#include "resize_b200.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace {

constexpr int THREADS = 256;
thread_local char g_err[256] = "";

int fail(int rc, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return rc;
}

template <typename T> struct Vec;
template <> struct Vec<float> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void unpack(const uint4& v, float (&f)[4]) {
        f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y); f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
    }
    static __device__ __forceinline__ uint4 pack(const float (&f)[4]) {
        return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
    }
};
template <> struct Vec<__nv_bfloat16> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void unpack(const uint4& v, float (&f)[8]) {
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            f[2 * i] = __uint_as_float(w[i] << 16);
            f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    static __device__ __forceinline__ uint4 pack(const float (&f)[8]) {
        unsigned w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 p = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
            w[i] = *reinterpret_cast<unsigned*>(&p);
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
};
template <> struct Vec<__half> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void unpack(const uint4& v, float (&f)[8]) {
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 p = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
            f[2 * i] = p.x; f[2 * i + 1] = p.y;
        }
    }
    static __device__ __forceinline__ uint4 pack(const float (&f)[8]) {
        unsigned w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __half2 p = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
            w[i] = *reinterpret_cast<unsigned*>(&p);
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
};

// ATen's area_pixel_compute_source_index (align_corners = false) and the corner / weight choice of
// upsample_bilinear2d: never contracted, so forward and backward agree on every (i0, i1, l1)
struct Tap { int i0, i1; float l0, l1; };
__device__ __forceinline__ Tap tap(int dst, float scale, int in) {
    float src = __fadd_rn(__fmul_rn(scale, (float)dst + 0.5f), -0.5f);
    src = src < 0.f ? 0.f : src;
    int i0 = (int)src;
    i0 = i0 < in - 1 ? i0 : in - 1;
    Tap t;
    t.i0 = i0;
    t.i1 = i0 < in - 1 ? i0 + 1 : i0;
    t.l1 = src - (float)i0;
    t.l0 = 1.f - t.l1;
    return t;
}

template <typename T, int MODE>
__global__ void __launch_bounds__(THREADS)
fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int64_t total, int H, int W, int CV, int Ho, int Wo,
           float sy, float sx, int fh, int fw) {
    constexpr int N = Vec<T>::N;
    for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * THREADS) {
        const int cv = (int)(i % CV);
        int64_t p = i / CV;
        const int ox = (int)(p % Wo); p /= Wo;
        const int oy = (int)(p % Ho);
        const int64_t n = p / Ho;
        const uint4* img = x + n * H * W * CV + cv;
        if (MODE == 0) {
            y[i] = __ldg(img + ((int64_t)(oy / fh) * W + ox / fw) * CV);
        } else {
            const Tap ty = tap(oy, sy, H), tx = tap(ox, sx, W);
            float a[N], b[N], c[N], d[N];
            Vec<T>::unpack(__ldg(img + ((int64_t)ty.i0 * W + tx.i0) * CV), a);
            Vec<T>::unpack(__ldg(img + ((int64_t)ty.i0 * W + tx.i1) * CV), b);
            Vec<T>::unpack(__ldg(img + ((int64_t)ty.i1 * W + tx.i0) * CV), c);
            Vec<T>::unpack(__ldg(img + ((int64_t)ty.i1 * W + tx.i1) * CV), d);
#pragma unroll
            for (int k = 0; k < N; ++k)
                a[k] = ty.l0 * (tx.l0 * a[k] + tx.l1 * b[k]) + ty.l1 * (tx.l0 * c[k] + tx.l1 * d[k]);
            y[i] = Vec<T>::pack(a);
        }
    }
}

// outputs whose taps can touch input index i: a superset, the exact weights decide
__device__ __forceinline__ void reach(int i, float inv_scale, int out, int& lo, int& hi) {
    const float a = ((float)i - 0.5f) * inv_scale - 0.5f, b = ((float)i + 1.5f) * inv_scale - 0.5f;
    lo = (int)floorf(a) - 1;
    hi = (int)ceilf(b) + 1;
    lo = lo < 0 ? 0 : lo;
    hi = hi > out - 1 ? out - 1 : hi;
}

template <typename T, int MODE>
__global__ void __launch_bounds__(THREADS)
bwd_kernel(const uint4* __restrict__ gy, uint4* __restrict__ gx, int64_t total, int H, int W, int CV, int Ho, int Wo,
           float sy, float sx, int fh, int fw) {
    constexpr int N = Vec<T>::N;
    for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * THREADS) {
        const int cv = (int)(i % CV);
        int64_t p = i / CV;
        const int ix = (int)(p % W); p /= W;
        const int iy = (int)(p % H);
        const int64_t n = p / H;
        const uint4* img = gy + n * Ho * Wo * CV + cv;
        float acc[N];
#pragma unroll
        for (int k = 0; k < N; ++k) acc[k] = 0.f;
        if (MODE == 0) {
            for (int oy = iy * fh; oy < (iy + 1) * fh; ++oy)
                for (int ox = ix * fw; ox < (ix + 1) * fw; ++ox) {
                    float g[N];
                    Vec<T>::unpack(__ldg(img + ((int64_t)oy * Wo + ox) * CV), g);
#pragma unroll
                    for (int k = 0; k < N; ++k) acc[k] += g[k];
                }
        } else {
            int ylo, yhi, xlo, xhi;
            reach(iy, 1.f / sy, Ho, ylo, yhi);
            reach(ix, 1.f / sx, Wo, xlo, xhi);
            for (int oy = ylo; oy <= yhi; ++oy) {
                const Tap ty = tap(oy, sy, H);
                const float wy = (ty.i0 == iy ? ty.l0 : 0.f) + (ty.i1 == iy ? ty.l1 : 0.f);
                if (ty.i0 != iy && ty.i1 != iy) continue;
                for (int ox = xlo; ox <= xhi; ++ox) {
                    const Tap tx = tap(ox, sx, W);
                    if (tx.i0 != ix && tx.i1 != ix) continue;
                    const float wgt = wy * ((tx.i0 == ix ? tx.l0 : 0.f) + (tx.i1 == ix ? tx.l1 : 0.f));
                    float g[N];
                    Vec<T>::unpack(__ldg(img + ((int64_t)oy * Wo + ox) * CV), g);
#pragma unroll
                    for (int k = 0; k < N; ++k) acc[k] = fmaf(wgt, g[k], acc[k]);
                }
            }
        }
        gx[i] = Vec<T>::pack(acc);
    }
}

int vec_of(int dtype) { return dtype == 0 ? 4 : 8; }

bool supported(int dtype, int C, int H, int W, int Ho, int Wo, int mode) {
    if (dtype < 0 || dtype > 2 || C <= 0 || C % vec_of(dtype)) return false;
    if (H <= 0 || W <= 0 || Ho <= 0 || Wo <= 0) return false;
    if (mode == 0) return Ho % H == 0 && Wo % W == 0;
    return mode == 1;
}

int check(const void* a, const void* b, int dtype, int N, int H, int W, int C, int Ho, int Wo, int mode) {
    if (!a || !b) return fail(-2, "null pointer");
    if (N <= 0 || !supported(dtype, C, H, W, Ho, Wo, mode)) return fail(-1, "unsupported resize (resize_b200_supported)");
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess || major != 10) {
        cudaGetLastError();
        return fail(-3, "no sm_100 CUDA device: resize_b200 has no CPU path");
    }
    return 0;
}

template <typename T, bool BWD>
int launch(const void* src, void* dst, int N, int H, int W, int C, int Ho, int Wo, int mode, cudaStream_t st) {
    const int CV = C / Vec<T>::N;
    const int64_t total = (int64_t)N * (BWD ? (int64_t)H * W : (int64_t)Ho * Wo) * CV;
    int64_t blocks = (total + THREADS - 1) / THREADS;
    if (blocks > 148 * 64) blocks = 148 * 64;
    // ATen: scale = in / out in float when a size is given (area_pixel_compute_scale)
    const float sy = (float)H / (float)Ho, sx = (float)W / (float)Wo;
    const int fh = Ho / H > 0 ? Ho / H : 1, fw = Wo / W > 0 ? Wo / W : 1;
    auto s = (const uint4*)src;
    auto d = (uint4*)dst;
    if (BWD) {
        if (mode == 0) bwd_kernel<T, 0><<<(unsigned)blocks, THREADS, 0, st>>>(s, d, total, H, W, CV, Ho, Wo, sy, sx, fh, fw);
        else bwd_kernel<T, 1><<<(unsigned)blocks, THREADS, 0, st>>>(s, d, total, H, W, CV, Ho, Wo, sy, sx, fh, fw);
    } else {
        if (mode == 0) fwd_kernel<T, 0><<<(unsigned)blocks, THREADS, 0, st>>>(s, d, total, H, W, CV, Ho, Wo, sy, sx, fh, fw);
        else fwd_kernel<T, 1><<<(unsigned)blocks, THREADS, 0, st>>>(s, d, total, H, W, CV, Ho, Wo, sy, sx, fh, fw);
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail((int)e, cudaGetErrorString(e));
}

template <bool BWD>
int dispatch(const void* src, void* dst, int dtype, int N, int H, int W, int C, int Ho, int Wo, int mode, void* stream) {
    if (int rc = check(src, dst, dtype, N, H, W, C, Ho, Wo, mode)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case 0: return launch<float, BWD>(src, dst, N, H, W, C, Ho, Wo, mode, st);
        case 1: return launch<__half, BWD>(src, dst, N, H, W, C, Ho, Wo, mode, st);
        default: return launch<__nv_bfloat16, BWD>(src, dst, N, H, W, C, Ho, Wo, mode, st);
    }
}

}  // namespace

extern "C" {

int resize_b200_version(void) { return RESIZE_B200_VERSION; }
const char* resize_b200_last_error(void) { return g_err; }
int resize_b200_supported(int dtype, int C, int H, int W, int Ho, int Wo, int mode) {
    return supported(dtype, C, H, W, Ho, Wo, mode) ? 1 : 0;
}
int resize_b200_forward(const void* x, void* y, int dtype, int N, int H, int W, int C, int Ho, int Wo, int mode,
                        void* cuda_stream) {
    return dispatch<false>(x, y, dtype, N, H, W, C, Ho, Wo, mode, cuda_stream);
}
int resize_b200_backward(const void* gy, void* gx, int dtype, int N, int H, int W, int C, int Ho, int Wo, int mode,
                         void* cuda_stream) {
    return dispatch<true>(gy, gx, dtype, N, H, W, C, Ho, Wo, mode, cuda_stream);
}

}  // extern "C"

// Fused CE + weighted-Dice segmentation loss (include/segloss_b200.h).  HBM-bound streaming kernels: the forward
// reads pred and target once and leaves a 4.6 KB table of per-(image, class) sums; the backward reads them once
// more and writes grad_pred.  One thread per pred pixel, NCHW so a warp's loads of one channel are contiguous.
#include "segloss_b200.h"

#include <cuda_runtime.h>
#include <stdio.h>

namespace {

constexpr int MAXC = SEGLOSS_B200_MAX_CLASSES;
constexpr int THREADS = 256;
thread_local char g_err[256] = "";

int fail(int rc, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return rc;
}

// softmax over the C channels of one pixel; q[c] = probabilities, returns log(sum exp(x - max)) + max
template <int C>
__device__ __forceinline__ float pixel_softmax(const float* __restrict__ p, size_t plane, float (&q)[C]) {
    float x[C], m = -INFINITY;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        x[c] = __ldg(p + c * plane);
        m = fmaxf(m, x[c]);
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        q[c] = __expf(x[c] - m);
        s += q[c];
    }
    const float inv = 1.f / s;
#pragma unroll
    for (int c = 0; c < C; ++c) q[c] *= inv;
    return m + __logf(s);
}

// how many of the scale x scale labels under pred pixel (y, x) equal each class
template <int C>
__device__ __forceinline__ void block_counts(const int64_t* __restrict__ t, int y, int x, int w, int scale,
                                             float (&cnt)[C]) {
#pragma unroll
    for (int c = 0; c < C; ++c) cnt[c] = 0.f;
    const size_t W = (size_t)w * scale;
    const int64_t* row = t + (size_t)y * scale * W + (size_t)x * scale;
    for (int j = 0; j < scale; ++j, row += W)
        for (int i = 0; i < scale; ++i) {
            const int64_t l = __ldg(row + i);
#pragma unroll
            for (int c = 0; c < C; ++c) cnt[c] += (l == c) ? 1.f : 0.f;
        }
}

template <int C>
__global__ void __launch_bounds__(THREADS)
segloss_fwd_kernel(const float* __restrict__ pred, const int64_t* __restrict__ target, const float* __restrict__ cw,
                   double* __restrict__ stats, int N, int h, int w, int scale, int per_block) {
    constexpr int NS = 3 * C + 2;
    const int n = blockIdx.y;
    const size_t plane = (size_t)h * w;
    const float* p = pred + (size_t)n * C * plane;
    const int64_t* t = target + (size_t)n * plane * scale * scale;
    float wgt[C], acc[NS];
#pragma unroll
    for (int c = 0; c < C; ++c) wgt[c] = __ldg(cw + c);
#pragma unroll
    for (int k = 0; k < NS; ++k) acc[k] = 0.f;
    const float area = (float)(scale * scale);
    const size_t first = (size_t)blockIdx.x * per_block;
    for (size_t pix = first + threadIdx.x; pix < first + per_block && pix < plane; pix += THREADS) {
        float q[C], cnt[C];
        const float lse = pixel_softmax<C>(p + pix, plane, q);
        block_counts<C>(t, (int)(pix / w), (int)(pix % w), w, scale, cnt);
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float wq = wgt[c] * q[c];
            acc[c] += wq * cnt[c];                                   // I
            acc[C + c] += wq * area;                                 // P
            acc[2 * C + c] += cnt[c];                                // O
            acc[3 * C] += wgt[c] * cnt[c] * (lse - __ldg(p + pix + c * plane));   // CE numerator: -w log q_t
            acc[3 * C + 1] += wgt[c] * cnt[c];                       // CE denominator
        }
    }
    __shared__ float red[THREADS / 32][NS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        float v = acc[k];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < NS) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < THREADS / 32; ++k) v += (double)red[k][threadIdx.x];
        double* dst = threadIdx.x < 3 * C ? stats + (size_t)n * 3 * C + threadIdx.x
                                          : stats + (size_t)N * 3 * C + (threadIdx.x - 3 * C);
        atomicAdd(dst, v);
    }
}

template <int C>
__global__ void __launch_bounds__(THREADS)
segloss_bwd_kernel(const float* __restrict__ pred, const int64_t* __restrict__ target, const float* __restrict__ cw,
                   const float* __restrict__ coef, float* __restrict__ grad, int N, int h, int w, int scale) {
    const int n = blockIdx.y;
    const size_t plane = (size_t)h * w;
    const size_t pix = (size_t)blockIdx.x * THREADS + threadIdx.x;
    if (pix >= plane) return;
    const float* p = pred + (size_t)n * C * plane + pix;
    const int64_t* t = target + (size_t)n * plane * scale * scale;
    const float inv_den = __ldg(coef + (size_t)N * 2 * C);
    const float area = (float)(scale * scale);
    float q[C], cnt[C];
    pixel_softmax<C>(p, plane, q);
    block_counts<C>(t, (int)(pix / w), (int)(pix % w), w, scale, cnt);
    float G[C], wsum = 0.f, dot = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float wc = __ldg(cw + c);
        wsum += wc * cnt[c];
        G[c] = __ldg(coef + ((size_t)n * 2) * C + c) * cnt[c] + __ldg(coef + ((size_t)n * 2 + 1) * C + c) * area;
        dot += q[c] * G[c];
    }
    float* g = grad + (size_t)n * C * plane + pix;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const float wc = __ldg(cw + c);
        g[c * plane] = inv_den * (q[c] * wsum - wc * cnt[c]) + q[c] * (G[c] - dot);
    }
}

int check(const void* a, const void* b, const void* c, const void* d, int N, int C, int h, int w, int scale) {
    if (!a || !b || !c || !d) return fail(-2, "null pointer");
    if (N <= 0 || h <= 0 || w <= 0 || scale <= 0) return fail(-1, "N, h, w and scale must be positive");
    if (C <= 0 || C > MAXC) return fail(-1, "number of classes must be in [1, SEGLOSS_B200_MAX_CLASSES]");
    if (N > 65535) return fail(-1, "N > 65535");
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess
        || major != 10) {
        cudaGetLastError();
        return fail(-3, "no sm_100 CUDA device: segloss_b200 has no CPU path");
    }
    return 0;
}

template <int C>
int fwd(const float* pred, const int64_t* target, const float* cw, double* stats, int N, int h, int w, int scale,
        cudaStream_t st) {
    const size_t plane = (size_t)h * w;
    // enough blocks for ~4 per SM, each reducing at least one and at most 16 strips of THREADS pixels
    size_t strips = (plane + THREADS - 1) / THREADS;
    size_t want = (592 + N - 1) / N;
    size_t per = strips / (want ? want : 1);
    per = per < 1 ? 1 : (per > 16 ? 16 : per);
    const int per_block = (int)per * THREADS;
    dim3 grid((unsigned)((plane + per_block - 1) / per_block), N);
    cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(double) * ((size_t)N * 3 * C + 2), st);
    if (e != cudaSuccess) return fail((int)e, cudaGetErrorString(e));
    segloss_fwd_kernel<C><<<grid, THREADS, 0, st>>>(pred, target, cw, stats, N, h, w, scale, per_block);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail((int)e, cudaGetErrorString(e));
}

template <int C>
int bwd(const float* pred, const int64_t* target, const float* cw, const float* coef, float* grad, int N, int h, int w,
        int scale, cudaStream_t st) {
    const size_t plane = (size_t)h * w;
    dim3 grid((unsigned)((plane + THREADS - 1) / THREADS), N);
    segloss_bwd_kernel<C><<<grid, THREADS, 0, st>>>(pred, target, cw, coef, grad, N, h, w, scale);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail((int)e, cudaGetErrorString(e));
}

#define SEGLOSS_DISPATCH(C, CALL)                                                                       \
    switch (C) {                                                                                        \
        case 1: { constexpr int K = 1; return CALL; }   case 2: { constexpr int K = 2; return CALL; }   \
        case 3: { constexpr int K = 3; return CALL; }   case 4: { constexpr int K = 4; return CALL; }   \
        case 5: { constexpr int K = 5; return CALL; }   case 6: { constexpr int K = 6; return CALL; }   \
        case 7: { constexpr int K = 7; return CALL; }   case 8: { constexpr int K = 8; return CALL; }   \
        case 9: { constexpr int K = 9; return CALL; }   case 10: { constexpr int K = 10; return CALL; } \
        case 11: { constexpr int K = 11; return CALL; } case 12: { constexpr int K = 12; return CALL; } \
        case 13: { constexpr int K = 13; return CALL; } case 14: { constexpr int K = 14; return CALL; } \
        case 15: { constexpr int K = 15; return CALL; } default: { constexpr int K = 16; return CALL; } \
    }

}  // namespace

extern "C" {

int segloss_b200_version(void) { return SEGLOSS_B200_VERSION; }
const char* segloss_b200_last_error(void) { return g_err; }

int segloss_b200_forward(const float* pred, const int64_t* target, const float* class_weights, double* stats, int N,
                         int C, int h, int w, int scale, void* cuda_stream) {
    if (int rc = check(pred, target, class_weights, stats, N, C, h, w, scale)) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    SEGLOSS_DISPATCH(C, (fwd<K>(pred, target, class_weights, stats, N, h, w, scale, st)))
}

int segloss_b200_backward(const float* pred, const int64_t* target, const float* class_weights, const float* coef,
                          float* grad_pred, int N, int C, int h, int w, int scale, void* cuda_stream) {
    if (int rc = check(pred, target, class_weights, coef, N, C, h, w, scale)) return rc;
    if (!grad_pred) return fail(-2, "null pointer");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    SEGLOSS_DISPATCH(C, (bwd<K>(pred, target, class_weights, coef, grad_pred, N, h, w, scale, st)))
}

}  // extern "C"

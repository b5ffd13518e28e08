// Fused training-mode BatchNorm2d + SiLU over channels-last activations (include/bnact_b200.h).
// All four big kernels are HBM streams: a thread owns one 16-byte vector column (the launch's thread count is a
// multiple of the number of vector columns, so a grid-stride walk never changes column and the per-channel
// constants and accumulators live in registers); consecutive threads read consecutive 16-byte vectors.
#include "bnact_b200.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>

namespace {

constexpr int THREADS = 256;
constexpr int UNROLL = 4;        // 16-byte loads in flight per thread in the apply kernel
constexpr int UNROLL_STATS = 8;  // ... in the statistics kernel (no stores, few registers)
constexpr int MAX_BLOCKS = 148 * 4;
constexpr int FIN_CH = 32, FIN_LANES = 32;   // finalize: a block sums the partials of 32 channels with 32 lanes each
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

// 16-byte read-only streaming load (ld.global.nc), kept in program order by asm volatile
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

template <int N>
__device__ __forceinline__ void load_consts(const float* __restrict__ p, int c0, float (&f)[N]) {
#pragma unroll
    for (int i = 0; i < N; i += 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p + c0 + i));
        f[i] = v.x; f[i + 1] = v.y; f[i + 2] = v.z; f[i + 3] = v.w;
    }
}

// two sums per channel, reduced over the threads of a block that share a vector column, to partial[block][2][C]
template <int N>
__device__ __forceinline__ void block_partials(float (&a)[N], float (&b)[N], int CV, int C, float* __restrict__ partial) {
    __shared__ float red[THREADS][2 * N + 1];
    const int t = threadIdx.x;
#pragma unroll
    for (int i = 0; i < N; ++i) { red[t][i] = a[i]; red[t][N + i] = b[i]; }
    __syncthreads();
    // CV * 2N sums of THREADS / CV terms each, spread over the block
    for (int item = t; item < CV * 2 * N; item += THREADS) {
        const int cv = item / (2 * N), k = item % (2 * N);
        float s = 0.f;
        for (int r = cv; r < THREADS; r += CV) s += red[r][k];
        const int ch = cv * N + (k % N);
        partial[((size_t)blockIdx.x * 2 + (k / N)) * C + ch] = s;
    }
}

__device__ __forceinline__ float silu_f(float y) { return __fdividef(y, 1.f + __expf(-y)); }
// d silu / dy = s (1 + y (1 - s))
__device__ __forceinline__ float dsilu_f(float y) {
    const float s = __fdividef(1.f, 1.f + __expf(-y));
    return s * (1.f + y * (1.f - s));
}

template <typename T>
__global__ void __launch_bounds__(THREADS)
stats_kernel(const uint4* __restrict__ x, float* __restrict__ partial, int64_t M, int C) {
    constexpr int N = Vec<T>::N;
    const int CV = C / N;
    const int64_t g = (int64_t)blockIdx.x * THREADS + threadIdx.x;
    const int cv = (int)(g % CV);
    const int64_t rs = (int64_t)gridDim.x * THREADS / CV;
    float piv[N], s[N], q[N];
    Vec<T>::unpack(__ldg(x + cv), piv);                       // row 0 of this column: the pivot
#pragma unroll
    for (int i = 0; i < N; ++i) s[i] = q[i] = 0.f;
    int64_t r = g / CV;
    for (; r + (UNROLL_STATS - 1) * rs < M; r += UNROLL_STATS * rs) {
        uint4 v[UNROLL_STATS];
#pragma unroll
        for (int u = 0; u < UNROLL_STATS; ++u) v[u] = ldg_stream(x + (r + u * rs) * CV + cv);
#pragma unroll
        for (int u = 0; u < UNROLL_STATS; ++u) {
            float f[N];
            Vec<T>::unpack(v[u], f);
#pragma unroll
            for (int i = 0; i < N; ++i) { const float d = f[i] - piv[i]; s[i] += d; q[i] = fmaf(d, d, q[i]); }
        }
    }
    for (; r < M; r += rs) {
        float f[N];
        Vec<T>::unpack(__ldg(x + r * CV + cv), f);
#pragma unroll
        for (int i = 0; i < N; ++i) { const float d = f[i] - piv[i]; s[i] += d; q[i] = fmaf(d, d, q[i]); }
    }
    block_partials<N>(s, q, CV, C, partial);
}

// sums over the blocks' partials for channel c = blockIdx.x * FIN_CH + (threadIdx.x % FIN_CH); valid in lane 0
__device__ __forceinline__ void sum_partials(const float* __restrict__ partial, int nblk, int C, int c, double& a, double& b) {
    __shared__ double red[2][FIN_LANES][FIN_CH + 1];
    const int ch = threadIdx.x % FIN_CH, lane = threadIdx.x / FIN_CH;
    a = b = 0.0;
    if (c < C)
        for (int k = lane; k < nblk; k += FIN_LANES) {
            a += (double)__ldg(partial + ((size_t)k * 2) * C + c);
            b += (double)__ldg(partial + ((size_t)k * 2 + 1) * C + c);
        }
    red[0][lane][ch] = a;
    red[1][lane][ch] = b;
    __syncthreads();
    if (lane == 0) {
        a = b = 0.0;
#pragma unroll 8
        for (int k = 0; k < FIN_LANES; ++k) { a += red[0][k][ch]; b += red[1][k][ch]; }
    }
}

template <typename T>
__global__ void __launch_bounds__(FIN_CH * FIN_LANES) stats_finalize_kernel(const T* __restrict__ x, const float* __restrict__ partial, int nblk,
                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                      float* __restrict__ running_mean, float* __restrict__ running_var,
                                      float* __restrict__ save, int64_t M, int C, float eps, float momentum) {
    const int c = blockIdx.x * FIN_CH + threadIdx.x % FIN_CH;
    double s, q;
    sum_partials(partial, nblk, C, c, s, q);
    if (c >= C || threadIdx.x >= FIN_CH) return;
    const double ms = s / (double)M;
    double var = q / (double)M - ms * ms;
    var = var < 0.0 ? 0.0 : var;
    const double mean = (double)(float)x[c] + ms;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float scale = gamma[c] * invstd;
    save[c] = (float)mean;
    save[C + c] = invstd;
    save[2 * C + c] = scale;
    save[3 * C + c] = beta[c] - (float)mean * scale;
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    if (running_var) {
        const double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

template <typename T, bool SILU>
__global__ void __launch_bounds__(THREADS)
apply_kernel(const uint4* __restrict__ x, uint4* __restrict__ z, const float* __restrict__ save, int64_t M, int C, int64_t ZV) {
    constexpr int N = Vec<T>::N;
    const int CV = C / N;
    const int64_t g = (int64_t)blockIdx.x * THREADS + threadIdx.x;
    const int cv = (int)(g % CV);
    const int64_t rs = (int64_t)gridDim.x * THREADS / CV;
    float sc[N], sh[N];
    load_consts<N>(save + 2 * C, cv * N, sc);
    load_consts<N>(save + 3 * C, cv * N, sh);
    int64_t r = g / CV;
    for (; r + (UNROLL - 1) * rs < M; r += UNROLL * rs) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = ldg_stream(x + (r + u * rs) * CV + cv);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            float f[N];
            Vec<T>::unpack(v[u], f);
#pragma unroll
            for (int i = 0; i < N; ++i) { const float y = fmaf(f[i], sc[i], sh[i]); f[i] = SILU ? silu_f(y) : y; }
            z[(r + u * rs) * ZV + cv] = Vec<T>::pack(f);
        }
    }
    for (; r < M; r += rs) {
        float f[N];
        Vec<T>::unpack(__ldg(x + r * CV + cv), f);
#pragma unroll
        for (int i = 0; i < N; ++i) { const float y = fmaf(f[i], sc[i], sh[i]); f[i] = SILU ? silu_f(y) : y; }
        z[r * ZV + cv] = Vec<T>::pack(f);
    }
}

// Inference (running statistics): z = act((x - mean) * gamma / sqrt(var + eps) + beta) in ONE pass.  The per-channel
// parameters come in float32 or in the activation's own 16-bit dtype (a model cast with .half() has half BatchNorm
// parameters and buffers); scale / shift are formed in float per thread (a thread owns one 16-byte channel vector).
template <typename T, typename P, int N>
__device__ __forceinline__ void load_params(const P* __restrict__ p, int c0, float (&f)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) f[i] = (float)__ldg(p + c0 + i);
}
template <typename T, typename P, bool SILU>
__global__ void __launch_bounds__(THREADS)
eval_kernel(const uint4* __restrict__ x, uint4* __restrict__ z, const P* __restrict__ gamma, const P* __restrict__ beta,
            const P* __restrict__ mean, const P* __restrict__ var, float eps, int64_t M, int C, int64_t ZV) {
    constexpr int N = Vec<T>::N;
    const int CV = C / N;
    const int64_t g = (int64_t)blockIdx.x * THREADS + threadIdx.x;
    const int cv = (int)(g % CV);
    const int64_t rs = (int64_t)gridDim.x * THREADS / CV;
    float sc[N], sh[N];
    {
        float ga[N], be[N], mu[N], va[N];
        load_params<T, P, N>(gamma, cv * N, ga);
        load_params<T, P, N>(beta, cv * N, be);
        load_params<T, P, N>(mean, cv * N, mu);
        load_params<T, P, N>(var, cv * N, va);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            sc[i] = ga[i] / sqrtf(va[i] + eps);
            sh[i] = fmaf(-mu[i], sc[i], be[i]);
        }
    }
    int64_t r = g / CV;
    for (; r + (UNROLL - 1) * rs < M; r += UNROLL * rs) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = ldg_stream(x + (r + u * rs) * CV + cv);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            float f[N];
            Vec<T>::unpack(v[u], f);
#pragma unroll
            for (int i = 0; i < N; ++i) { const float y = fmaf(f[i], sc[i], sh[i]); f[i] = SILU ? silu_f(y) : y; }
            z[(r + u * rs) * ZV + cv] = Vec<T>::pack(f);
        }
    }
    for (; r < M; r += rs) {
        float f[N];
        Vec<T>::unpack(__ldg(x + r * CV + cv), f);
#pragma unroll
        for (int i = 0; i < N; ++i) { const float y = fmaf(f[i], sc[i], sh[i]); f[i] = SILU ? silu_f(y) : y; }
        z[r * ZV + cv] = Vec<T>::pack(f);
    }
}

// sums over rows of gy and gy * (x - mean), gy = gz * act'(y)
template <typename T, bool SILU>
__global__ void __launch_bounds__(THREADS)
bwd_reduce_kernel(const uint4* __restrict__ x, const uint4* __restrict__ gz, const float* __restrict__ save,
                  const float* __restrict__ beta, float* __restrict__ partial, int64_t M, int C, int64_t GV) {
    constexpr int N = Vec<T>::N;
    const int CV = C / N;
    const int64_t g = (int64_t)blockIdx.x * THREADS + threadIdx.x;
    const int cv = (int)(g % CV);
    const int64_t rs = (int64_t)gridDim.x * THREADS / CV;
    float mean[N], sc[N], be[N], a[N], b[N];
    load_consts<N>(save, cv * N, mean);
    load_consts<N>(save + 2 * C, cv * N, sc);
    load_consts<N>(beta, cv * N, be);
#pragma unroll
    for (int i = 0; i < N; ++i) a[i] = b[i] = 0.f;
    constexpr int U = 4;
    int64_t r = g / CV;
    uint4 cx[U], cg[U], nx[U], ng[U];                         // software pipeline as in stats_kernel
    const uint4 zero = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const bool in = r + u * rs < M;
        cx[u] = in ? ldg_stream(x + (r + u * rs) * CV + cv) : zero;
        cg[u] = in ? ldg_stream(gz + (r + u * rs) * GV + cv) : zero;
    }
    while (r < M) {
        const int64_t rn = r + U * rs;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool in = rn + u * rs < M;
            nx[u] = in ? ldg_stream(x + (rn + u * rs) * CV + cv) : zero;
            ng[u] = in ? ldg_stream(gz + (rn + u * rs) * GV + cv) : zero;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (r + u * rs < M) {
                float f[N], gzz[N];
                Vec<T>::unpack(cx[u], f);
                Vec<T>::unpack(cg[u], gzz);
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const float d = f[i] - mean[i];
                    const float gy = SILU ? gzz[i] * dsilu_f(fmaf(d, sc[i], be[i])) : gzz[i];
                    a[i] += gy;
                    b[i] = fmaf(gy, d, b[i]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) { cx[u] = nx[u]; cg[u] = ng[u]; }
        r = rn;
    }
    block_partials<N>(a, b, CV, C, partial);
}

__global__ void __launch_bounds__(FIN_CH * FIN_LANES) bwd_finalize_kernel(const float* __restrict__ partial, int nblk, const float* __restrict__ save,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ coef,
                                    int64_t M, int C) {
    const int c = blockIdx.x * FIN_CH + threadIdx.x % FIN_CH;
    double a, b;
    sum_partials(partial, nblk, C, c, a, b);
    if (c >= C || threadIdx.x >= FIN_CH) return;
    const double invstd = save[C + c], scale = save[2 * C + c];
    dbeta[c] = (float)a;
    dgamma[c] = (float)(b * invstd);
    coef[c] = (float)(scale * a / (double)M);                       // k1
    coef[C + c] = (float)(scale * invstd * invstd * b / (double)M);  // k2:  dx = scale*gy - k1 - (x - mean)*k2
}

template <typename T, bool SILU>
__global__ void __launch_bounds__(THREADS)
bwd_apply_kernel(const uint4* __restrict__ x, const uint4* __restrict__ gz, uint4* __restrict__ dx,
                 const float* __restrict__ save, const float* __restrict__ beta, const float* __restrict__ coef,
                 int64_t M, int C, int64_t GV) {
    constexpr int N = Vec<T>::N;
    const int CV = C / N;
    const int64_t g = (int64_t)blockIdx.x * THREADS + threadIdx.x;
    const int cv = (int)(g % CV);
    const int64_t rs = (int64_t)gridDim.x * THREADS / CV;
    float mean[N], sc[N], be[N], k1[N], k2[N];
    load_consts<N>(save, cv * N, mean);
    load_consts<N>(save + 2 * C, cv * N, sc);
    load_consts<N>(beta, cv * N, be);
    load_consts<N>(coef, cv * N, k1);
    load_consts<N>(coef + C, cv * N, k2);
    constexpr int U = 4;
    int64_t r = g / CV;
    for (; r < M; r += U * rs) {
        uint4 vx[U], vg[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (r + u * rs < M) { vx[u] = __ldg(x + (r + u * rs) * CV + cv); vg[u] = __ldg(gz + (r + u * rs) * GV + cv); }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (r + u * rs < M) {
                float f[N], gzz[N];
                Vec<T>::unpack(vx[u], f);
                Vec<T>::unpack(vg[u], gzz);
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const float d = f[i] - mean[i];
                    const float gy = SILU ? gzz[i] * dsilu_f(fmaf(d, sc[i], be[i])) : gzz[i];
                    f[i] = fmaf(sc[i], gy, -k1[i]) - d * k2[i];
                }
                dx[(r + u * rs) * CV + cv] = Vec<T>::pack(f);
            }
    }
}

int vec_of(int dtype) { return dtype == 0 ? 4 : 8; }

bool supported(int dtype, int C) {
    if (dtype < 0 || dtype > 2 || C <= 0) return false;
    const int n = vec_of(dtype);
    if (C % n) return false;
    const int cv = C / n;
    return cv <= THREADS && (cv & (cv - 1)) == 0;
}

int blocks_for(int dtype, int64_t M, int C) {
    const int64_t vecs = M * (C / vec_of(dtype));
    int64_t b = (vecs + (int64_t)THREADS * UNROLL - 1) / ((int64_t)THREADS * UNROLL);
    return (int)(b < 1 ? 1 : (b > MAX_BLOCKS ? MAX_BLOCKS : b));
}

int check_common(int dtype, int64_t M, int C, int act) {
    if (!supported(dtype, C)) return fail(-1, "unsupported dtype / channel count (bnact_b200_supported)");
    if (M <= 0) return fail(-1, "M must be positive");
    if (act != 0 && act != 1) return fail(-1, "act must be 0 (identity) or 1 (SiLU)");
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess || major != 10) {
        cudaGetLastError();
        return fail(-3, "no sm_100 CUDA device: bnact_b200 has no CPU path");
    }
    return 0;
}

template <typename T>
int forward_t(const void* x, void* z, const float* gamma, const float* beta, float* rm, float* rv, float* save,
              float* partial, int dtype, int64_t M, int C, float eps, float momentum, int act, cudaStream_t st,
              int64_t z_pitch) {
    const int nblk = blocks_for(dtype, M, C);
    const int64_t ZV = z_pitch / vec_of(dtype);
    stats_kernel<T><<<nblk, THREADS, 0, st>>>((const uint4*)x, partial, M, C);
    stats_finalize_kernel<T><<<(C + FIN_CH - 1) / FIN_CH, FIN_CH * FIN_LANES, 0, st>>>((const T*)x, partial, nblk, gamma, beta, rm, rv, save, M, C,
                                                             eps, momentum);
    if (act) apply_kernel<T, true><<<nblk, THREADS, 0, st>>>((const uint4*)x, (uint4*)z, save, M, C, ZV);
    else apply_kernel<T, false><<<nblk, THREADS, 0, st>>>((const uint4*)x, (uint4*)z, save, M, C, ZV);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail((int)e, cudaGetErrorString(e));
}

template <typename T>
int backward_t(const void* x, const void* gz, void* dx, const float* beta, const float* save, float* dgamma,
               float* dbeta, float* coef, float* partial, int dtype, int64_t M, int C, int act, cudaStream_t st,
               int64_t gz_pitch) {
    const int nblk = blocks_for(dtype, M, C);
    const int64_t GV = gz_pitch / vec_of(dtype);  // row pitch of gz in 16-byte vectors
    if (act) bwd_reduce_kernel<T, true><<<nblk, THREADS, 0, st>>>((const uint4*)x, (const uint4*)gz, save, beta, partial, M, C, GV);
    else bwd_reduce_kernel<T, false><<<nblk, THREADS, 0, st>>>((const uint4*)x, (const uint4*)gz, save, beta, partial, M, C, GV);
    bwd_finalize_kernel<<<(C + FIN_CH - 1) / FIN_CH, FIN_CH * FIN_LANES, 0, st>>>(partial, nblk, save, dgamma, dbeta, coef, M, C);
    if (act) bwd_apply_kernel<T, true><<<nblk, THREADS, 0, st>>>((const uint4*)x, (const uint4*)gz, (uint4*)dx, save, beta, coef, M, C, GV);
    else bwd_apply_kernel<T, false><<<nblk, THREADS, 0, st>>>((const uint4*)x, (const uint4*)gz, (uint4*)dx, save, beta, coef, M, C, GV);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail((int)e, cudaGetErrorString(e));
}

template <typename T, typename P>
int eval_t(const void* x, void* z, const void* gamma, const void* beta, const void* mean, const void* var, int dtype,
           int64_t M, int C, float eps, int act, cudaStream_t st, int64_t z_pitch) {
    const int nblk = blocks_for(dtype, M, C);
    const int64_t ZV = z_pitch / vec_of(dtype);
    if (act) eval_kernel<T, P, true><<<nblk, THREADS, 0, st>>>((const uint4*)x, (uint4*)z, (const P*)gamma, (const P*)beta, (const P*)mean, (const P*)var, eps, M, C, ZV);
    else eval_kernel<T, P, false><<<nblk, THREADS, 0, st>>>((const uint4*)x, (uint4*)z, (const P*)gamma, (const P*)beta, (const P*)mean, (const P*)var, eps, M, C, ZV);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : fail((int)e, cudaGetErrorString(e));
}

}  // namespace

extern "C" {

int bnact_b200_eval_pitched(const void* x, void* z, const void* gamma, const void* beta, const void* running_mean,
                            const void* running_var, int dtype, int params_in_dtype, int64_t M, int C, float eps, int act,
                            int64_t z_pitch, void* cuda_stream) {
    if (!x || !z || !gamma || !beta || !running_mean || !running_var) return fail(-2, "null pointer");
    if (int rc = check_common(dtype, M, C, act)) return rc;
    if (z_pitch < C || z_pitch % vec_of(dtype) || (reinterpret_cast<uintptr_t>(z) & 15u))
        return fail(-1, "z_pitch must be >= C and a whole number of 16-byte vectors, z 16-byte aligned");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const bool pt = params_in_dtype != 0 && dtype != 0;
    switch (dtype) {
        case 0: return eval_t<float, float>(x, z, gamma, beta, running_mean, running_var, dtype, M, C, eps, act, st, z_pitch);
        case 1: return pt ? eval_t<__half, __half>(x, z, gamma, beta, running_mean, running_var, dtype, M, C, eps, act, st, z_pitch)
                          : eval_t<__half, float>(x, z, gamma, beta, running_mean, running_var, dtype, M, C, eps, act, st, z_pitch);
        default: return pt ? eval_t<__nv_bfloat16, __nv_bfloat16>(x, z, gamma, beta, running_mean, running_var, dtype, M, C, eps, act, st, z_pitch)
                           : eval_t<__nv_bfloat16, float>(x, z, gamma, beta, running_mean, running_var, dtype, M, C, eps, act, st, z_pitch);
    }
}

int bnact_b200_eval(const void* x, void* z, const void* gamma, const void* beta, const void* running_mean,
                    const void* running_var, int dtype, int params_in_dtype, int64_t M, int C, float eps, int act,
                    void* cuda_stream) {
    return bnact_b200_eval_pitched(x, z, gamma, beta, running_mean, running_var, dtype, params_in_dtype, M, C, eps, act, C,
                                   cuda_stream);
}

int bnact_b200_version(void) { return BNACT_B200_VERSION; }
const char* bnact_b200_last_error(void) { return g_err; }
int bnact_b200_supported(int dtype, int C) { return supported(dtype, C) ? 1 : 0; }

size_t bnact_b200_partial_floats(int dtype, int64_t M, int C) {
    if (!supported(dtype, C) || M <= 0) return 0;
    return (size_t)blocks_for(dtype, M, C) * 2 * (size_t)C;
}

int bnact_b200_forward_pitched(const void* x, void* z, const float* gamma, const float* beta, float* running_mean,
                               float* running_var, float* save, float* partial, int dtype, int64_t M, int C, float eps,
                               float momentum, int act, int64_t z_pitch, void* cuda_stream) {
    if (!x || !z || !gamma || !beta || !save || !partial) return fail(-2, "null pointer");
    if (int rc = check_common(dtype, M, C, act)) return rc;
    if (z_pitch < C || z_pitch % vec_of(dtype) || (reinterpret_cast<uintptr_t>(z) & 15u))
        return fail(-1, "z_pitch must be >= C and a whole number of 16-byte vectors, z 16-byte aligned");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    switch (dtype) {
        case 0: return forward_t<float>(x, z, gamma, beta, running_mean, running_var, save, partial, dtype, M, C, eps, momentum, act, st, z_pitch);
        case 1: return forward_t<__half>(x, z, gamma, beta, running_mean, running_var, save, partial, dtype, M, C, eps, momentum, act, st, z_pitch);
        default: return forward_t<__nv_bfloat16>(x, z, gamma, beta, running_mean, running_var, save, partial, dtype, M, C, eps, momentum, act, st, z_pitch);
    }
}

int bnact_b200_forward(const void* x, void* z, const float* gamma, const float* beta, float* running_mean,
                       float* running_var, float* save, float* partial, int dtype, int64_t M, int C, float eps,
                       float momentum, int act, void* cuda_stream) {
    return bnact_b200_forward_pitched(x, z, gamma, beta, running_mean, running_var, save, partial, dtype, M, C, eps,
                                      momentum, act, C, cuda_stream);
}

int bnact_b200_backward_pitched(const void* x, const void* gz, void* dx, const float* gamma, const float* beta,
                                const float* save, float* dgamma, float* dbeta, float* coef, float* partial, int dtype,
                                int64_t M, int C, int act, int64_t gz_pitch, void* cuda_stream) {
    (void)gamma;
    if (!x || !gz || !dx || !beta || !save || !dgamma || !dbeta || !coef || !partial) return fail(-2, "null pointer");
    if (int rc = check_common(dtype, M, C, act)) return rc;
    if (gz_pitch < C || gz_pitch % vec_of(dtype) || (reinterpret_cast<uintptr_t>(gz) & 15u))
        return fail(-1, "gz_pitch must be >= C and a whole number of 16-byte vectors, gz 16-byte aligned");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    switch (dtype) {
        case 0: return backward_t<float>(x, gz, dx, beta, save, dgamma, dbeta, coef, partial, dtype, M, C, act, st, gz_pitch);
        case 1: return backward_t<__half>(x, gz, dx, beta, save, dgamma, dbeta, coef, partial, dtype, M, C, act, st, gz_pitch);
        default: return backward_t<__nv_bfloat16>(x, gz, dx, beta, save, dgamma, dbeta, coef, partial, dtype, M, C, act, st, gz_pitch);
    }
}

int bnact_b200_backward(const void* x, const void* gz, void* dx, const float* gamma, const float* beta,
                        const float* save, float* dgamma, float* dbeta, float* coef, float* partial, int dtype,
                        int64_t M, int C, int act, void* cuda_stream) {
    return bnact_b200_backward_pitched(x, gz, dx, gamma, beta, save, dgamma, dbeta, coef, partial, dtype, M, C, act, C,
                                       cuda_stream);
}

}  // extern "C"

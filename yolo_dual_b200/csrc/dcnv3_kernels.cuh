// dcnv3_b200 — kernels.
//
// Two families:
//   *_vec_*    one thread owns a 16-byte channel vector (4 x f32 / 8 x f16|bf16) of one
//              (n, ho, wo, g); lanes run fastest over the vectors of a pixel, so every corner
//              fetch is an LDG.128 and the lanes of one group are an aligned power-of-two
//              segment of a warp (grad_offset / grad_mask reduce with __shfl_xor, no shared
//              memory, no barriers, no zero-init, one plain store per value).
//   *_any_*    generic fallback for shapes the vector path cannot take (group_channels not a
//              multiple of the vector width, f64, odd alignment, runtime-sized fused softmax).
//              Still CUDA, still this library: there is no CPU path.
//
// Reference semantics reproduced (models/ops_dcnv3/src/cuda/dcnv3_im2col_cuda.cuh):
//   forward  dcnv3_im2col_gpu_kernel :216-275 + dcnv3_im2col_bilinear :32-80
//   backward dcnv3_col2im_gpu_kernel_* :278-839 + dcnv3_col2im_bilinear :82-147
#pragma once

#include "dcnv3_common.cuh"

namespace dcnv3 {

constexpr int kThreads = 256;
constexpr int kMaxSoftmaxP = 49;  // runtime-sized fused softmax keeps P values in registers/local

// Decode a flat vector index into (pixel, vector-in-pixel, group, n, ho, wo).
struct VecCoord {
    unsigned pix;
    int v, g, n, ho, wo;
};
__device__ __forceinline__ VecCoord decode_vec(unsigned idx, const Geo &q, int vec_per_pix,
                                               int lanes_per_group) {
    VecCoord c;
    c.pix = idx / (unsigned)vec_per_pix;
    c.v = (int)(idx - c.pix * (unsigned)vec_per_pix);
    c.g = c.v / lanes_per_group;
    const unsigned row = c.pix / (unsigned)q.Wo;
    c.wo = (int)(c.pix - row * (unsigned)q.Wo);
    c.n = (int)(row / (unsigned)q.Ho);
    c.ho = (int)(row - (unsigned)c.n * (unsigned)q.Ho);
    return c;
}

// softmax statistics of the P logits of one (pixel, g): max and 1/sum(exp(l - max))
template <typename T, int KP>
__device__ __forceinline__ void softmax_stats(const T *pm, int P, float &mx, float &inv) {
    const int n = KP ? KP : P;
    mx = -INFINITY;
#pragma unroll
    for (int p = 0; p < n; ++p) mx = fmaxf(mx, to_math(pm[p]));
    float sum = 0.f;
#pragma unroll
    for (int p = 0; p < n; ++p) sum += expf(to_math(pm[p]) - mx);
    inv = 1.f / sum;
}

// ===========================================================================
// forward, vector path
// ===========================================================================
template <typename T, int KP, bool LOGITS>
__global__ void __launch_bounds__(kThreads)
fwd_vec_kernel(const T *__restrict__ in, const T *__restrict__ off, const T *__restrict__ mask,
               T *__restrict__ out, const Geo q, const int vec_per_pix,
               const int lanes_per_group, const unsigned total) {
    constexpr int VEC = Vec<T>::N;
    const unsigned idx = blockIdx.x * (unsigned)kThreads + threadIdx.x;
    if (idx >= total) return;
    const VecCoord c = decode_vec(idx, q, vec_per_pix, lanes_per_group);

    float p0h_, p0w_;
    window_origin<float>(q, c.ho, c.wo, p0h_, p0w_);
    const T *im = in + (size_t)c.n * q.H * q.W * q.C + c.v * VEC;
    const size_t pg = (size_t)c.pix * q.G + c.g;
    const T *po = off + pg * q.P * 2;
    const T *pm = mask + pg * q.P;

    float mx = 0.f, inv = 1.f;
    if (LOGITS) softmax_stats<T, KP>(pm, q.P, mx, inv);

    float acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = 0.f;

    const int kw = KP ? 3 : q.kw, kh = KP ? 3 : q.kh;
    int p = 0;
#pragma unroll
    for (int i = 0; i < kw; ++i) {
#pragma unroll
        for (int j = 0; j < kh; ++j, ++p) {
            const float2 o = load_offset_pair(po + 2 * p);
            Point<float> t;
            locate<float>(q, p0h_, p0w_, i, j, o.x, o.y, t);
            float m = to_math(pm[p]);
            if (LOGITS) m = expf(m - mx) * inv;
            if (t.bits) {
                const float w1 = t.hh * t.hw * m, w2 = t.hh * t.lw * m;
                const float w3 = t.lh * t.hw * m, w4 = t.lh * t.lw * m;
                const T *r1 = im + (t.h_low * q.W + t.w_low) * q.C;
                const T *r3 = r1 + q.W * q.C;
                float v[VEC];
                if (t.bits & B_C1) {
                    unpack(ldg128(r1), v, (const T *)nullptr);
#pragma unroll
                    for (int k = 0; k < VEC; ++k) acc[k] = fmaf(w1, v[k], acc[k]);
                }
                if (t.bits & B_C2) {
                    unpack(ldg128(r1 + q.C), v, (const T *)nullptr);
#pragma unroll
                    for (int k = 0; k < VEC; ++k) acc[k] = fmaf(w2, v[k], acc[k]);
                }
                if (t.bits & B_C3) {
                    unpack(ldg128(r3), v, (const T *)nullptr);
#pragma unroll
                    for (int k = 0; k < VEC; ++k) acc[k] = fmaf(w3, v[k], acc[k]);
                }
                if (t.bits & B_C4) {
                    unpack(ldg128(r3 + q.C), v, (const T *)nullptr);
#pragma unroll
                    for (int k = 0; k < VEC; ++k) acc[k] = fmaf(w4, v[k], acc[k]);
                }
            }
        }
    }
    *reinterpret_cast<uint4 *>(out + (size_t)c.pix * q.C + c.v * VEC) = pack(acc, (const T *)nullptr);
}

// ===========================================================================
// backward, vector path.  A = accumulation type of grad_input (float: gin itself
// for f32 storage or the fp32 workspace for 16-bit storage; T: packed 16-bit reds).
// ===========================================================================
template <typename T> __device__ __forceinline__ void store_pair(T *p, float x, float y);
template <> __device__ __forceinline__ void store_pair<float>(float *p, float x, float y) {
    *reinterpret_cast<float2 *>(p) = make_float2(x, y);
}
template <> __device__ __forceinline__ void store_pair<__half>(__half *p, float x, float y) {
    *reinterpret_cast<__half2 *>(p) = __floats2half2_rn(x, y);
}
template <> __device__ __forceinline__ void store_pair<__nv_bfloat16>(__nv_bfloat16 *p, float x, float y) {
    *reinterpret_cast<__nv_bfloat162 *>(p) = __floats2bfloat162_rn(x, y);
}

template <typename T, typename A, int KP, bool LOGITS>
__global__ void __launch_bounds__(kThreads)
bwd_vec_kernel(const T *__restrict__ in, const T *__restrict__ off, const T *__restrict__ mask,
               const T *__restrict__ gout, A *__restrict__ gin, T *__restrict__ goff,
               T *__restrict__ gmask, const Geo q, const int vec_per_pix,
               const int lanes_per_group, const unsigned total) {
    static_assert(!LOGITS || KP > 0, "fused softmax in the vector path needs a compile-time P");
    constexpr int VEC = Vec<T>::N;
    unsigned idx = blockIdx.x * (unsigned)kThreads + threadIdx.x;
    const bool active = idx < total;  // tail lanes stay for the shuffles
    if (!active) idx = total - 1;
    const VecCoord c = decode_vec(idx, q, vec_per_pix, lanes_per_group);

    float p0h_, p0w_;
    window_origin<float>(q, c.ho, c.wo, p0h_, p0w_);
    const size_t img = (size_t)c.n * q.H * q.W * q.C + c.v * VEC;
    const T *im = in + img;
    A *gim = gin + img;
    const size_t pg = (size_t)c.pix * q.G + c.g;
    const T *po = off + pg * q.P * 2;
    const T *pm = mask + pg * q.P;
    T *d_o = goff + pg * q.P * 2;
    T *d_m = gmask + pg * q.P;
    const bool writer = active && (c.v % lanes_per_group) == 0;

    float go[VEC];
    unpack(ldg128(gout + (size_t)c.pix * q.C + c.v * VEC), go, (const T *)nullptr);

    float mx = 0.f, inv = 1.f;
    if (LOGITS) softmax_stats<T, KP>(pm, q.P, mx, inv);
    float prob[KP ? KP : 1], gm[KP ? KP : 1];  // only live when LOGITS

    const int kw = KP ? 3 : q.kw, kh = KP ? 3 : q.kh;
    int p = 0;
#pragma unroll
    for (int i = 0; i < kw; ++i) {
#pragma unroll
        for (int j = 0; j < kh; ++j, ++p) {
            const float2 o = load_offset_pair(po + 2 * p);
            Point<float> t;
            locate<float>(q, p0h_, p0w_, i, j, o.x, o.y, t);
            float m = to_math(pm[p]);
            if (LOGITS) m = expf(m - mx) * inv;

            float s_m = 0.f, s_w = 0.f, s_h = 0.f;
            if (t.bits) {
                const int base = (t.h_low * q.W + t.w_low) * q.C;
                const int row = q.W * q.C;
                float v1[VEC], v2[VEC], v3[VEC], v4[VEC];
#pragma unroll
                for (int k = 0; k < VEC; ++k) v1[k] = v2[k] = v3[k] = v4[k] = 0.f;
                if (t.bits & B_C1) unpack(ldg128(im + base), v1, (const T *)nullptr);
                if (t.bits & B_C2) unpack(ldg128(im + base + q.C), v2, (const T *)nullptr);
                if (t.bits & B_C3) unpack(ldg128(im + base + row), v3, (const T *)nullptr);
                if (t.bits & B_C4) unpack(ldg128(im + base + row + q.C), v4, (const T *)nullptr);
                const float w1 = t.hh * t.hw, w2 = t.hh * t.lw, w3 = t.lh * t.hw, w4 = t.lh * t.lw;
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    // cuh:114-139: grad_w_weight = hh*(v2-v1) + lh*(v4-v3),
                    //              grad_h_weight = hw*(v3-v1) + lw*(v4-v2)
                    const float val = w1 * v1[k] + w2 * v2[k] + w3 * v3[k] + w4 * v4[k];
                    const float gw = t.hh * (v2[k] - v1[k]) + t.lh * (v4[k] - v3[k]);
                    const float gh = t.hw * (v3[k] - v1[k]) + t.lw * (v4[k] - v2[k]);
                    s_m = fmaf(go[k], val, s_m);
                    s_w = fmaf(go[k], gw, s_w);
                    s_h = fmaf(go[k], gh, s_h);
                }
                if (active) {  // cuh:116,124,132,140: grad_im[corner] += w_k * top_grad * mask
                    if (t.bits & B_C1) RedAdd<VEC>::run(gim + base, go, w1 * m);
                    if (t.bits & B_C2) RedAdd<VEC>::run(gim + base + q.C, go, w2 * m);
                    if (t.bits & B_C3) RedAdd<VEC>::run(gim + base + row, go, w3 * m);
                    if (t.bits & B_C4) RedAdd<VEC>::run(gim + base + row + q.C, go, w4 * m);
                }
            }
            // sum over the channels of the group: lanes of one group are an aligned segment
            for (int d = lanes_per_group >> 1; d > 0; d >>= 1) {
                s_m += shfl_xor(s_m, d);
                s_w += shfl_xor(s_w, d);
                s_h += shfl_xor(s_h, d);
            }
            const float sm = q.scale * m;  // cuh:145-146
            if (writer) store_pair<T>(d_o + 2 * p, sm * s_w, sm * s_h);
            if (LOGITS) {
                prob[KP ? p : 0] = m;
                gm[KP ? p : 0] = s_m;
            } else if (writer) {
                d_m[p] = from_math<T>(s_m);  // cuh:144
            }
        }
    }
    if (LOGITS && writer) {  // softmax Jacobian folded in: dl_p = m_p * (gm_p - sum_q m_q gm_q)
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < (KP ? KP : 1); ++k) dot = fmaf(prob[k], gm[k], dot);
#pragma unroll
        for (int k = 0; k < (KP ? KP : 1); ++k) d_m[k] = from_math<T>(prob[k] * (gm[k] - dot));
    }
}

// fp32 workspace -> 16-bit grad_input (ACC_OPMATH), 8 elements per thread
template <typename T>
__global__ void __launch_bounds__(kThreads)
cast_ws_kernel(const float *__restrict__ ws, T *__restrict__ dst, const size_t n_vec8,
               const size_t n_total) {
    const size_t i = blockIdx.x * (size_t)kThreads + threadIdx.x;
    if (i < n_vec8) {
        const float4 a = *reinterpret_cast<const float4 *>(ws + i * 8);
        const float4 b = *reinterpret_cast<const float4 *>(ws + i * 8 + 4);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        *reinterpret_cast<uint4 *>(dst + i * 8) = pack(v, (const T *)nullptr);
    } else if (i == n_vec8) {  // scalar tail (generic path only)
        for (size_t k = n_vec8 * 8; k < n_total; ++k) dst[k] = from_math<T>(ws[k]);
    }
}

// ===========================================================================
// generic fallback kernels (any group_channels, any dtype incl. f64)
// ===========================================================================
template <typename T, bool LOGITS>
__global__ void __launch_bounds__(kThreads)
fwd_any_kernel(const T *__restrict__ in, const T *__restrict__ off, const T *__restrict__ mask,
               T *__restrict__ out, const Geo q, const size_t total) {
    using M = typename OpMath<T>::type;
    const size_t idx = blockIdx.x * (size_t)kThreads + threadIdx.x;
    if (idx >= total) return;
    const size_t pix = idx / q.C;
    const int ch = (int)(idx - pix * q.C);
    const int g = ch / q.gc;
    const size_t row = pix / q.Wo;
    const int wo = (int)(pix - row * q.Wo);
    const int n = (int)(row / q.Ho);
    const int ho = (int)(row - (size_t)n * q.Ho);

    M p0h_, p0w_;
    window_origin<M>(q, ho, wo, p0h_, p0w_);
    const T *im = in + (size_t)n * q.H * q.W * q.C + ch;
    const size_t pg = pix * q.G + g;
    const T *po = off + pg * q.P * 2;
    const T *pm = mask + pg * q.P;

    M mx = 0, inv = 1;
    if (LOGITS) {
        mx = -INFINITY;
        for (int p = 0; p < q.P; ++p) mx = max(mx, (M)to_math(pm[p]));
        M sum = 0;
        for (int p = 0; p < q.P; ++p) sum += exp((M)to_math(pm[p]) - mx);
        inv = (M)1 / sum;
    }
    M acc = 0;
    int p = 0;
    for (int i = 0; i < q.kw; ++i)
        for (int j = 0; j < q.kh; ++j, ++p) {
            Point<M> t;
            locate<M>(q, p0h_, p0w_, i, j, (M)to_math(po[2 * p]), (M)to_math(po[2 * p + 1]), t);
            if (!t.bits) continue;
            M m = to_math(pm[p]);
            if (LOGITS) m = exp(m - mx) * inv;
            const size_t base = ((size_t)t.h_low * q.W + t.w_low) * q.C;
            const size_t rw = (size_t)q.W * q.C;
            const M v1 = (t.bits & B_C1) ? (M)to_math(im[base]) : (M)0;
            const M v2 = (t.bits & B_C2) ? (M)to_math(im[base + q.C]) : (M)0;
            const M v3 = (t.bits & B_C3) ? (M)to_math(im[base + rw]) : (M)0;
            const M v4 = (t.bits & B_C4) ? (M)to_math(im[base + rw + q.C]) : (M)0;
            acc += (t.hh * t.hw * v1 + t.hh * t.lw * v2 + t.lh * t.hw * v3 + t.lh * t.lw * v4) * m;
        }
    out[idx] = from_math<T>(acc);
}

// one warp per (pixel, g); lanes stride over the group's channels
template <typename T, typename A, bool LOGITS>
__global__ void __launch_bounds__(kThreads)
bwd_any_kernel(const T *__restrict__ in, const T *__restrict__ off, const T *__restrict__ mask,
               const T *__restrict__ gout, A *__restrict__ gin, T *__restrict__ goff,
               T *__restrict__ gmask, const Geo q, const size_t n_units) {
    using M = typename OpMath<T>::type;
    const size_t unit = (blockIdx.x * (size_t)kThreads + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (unit >= n_units) return;  // whole warp leaves together
    const size_t pix = unit / q.G;
    const int g = (int)(unit - pix * q.G);
    const size_t row = pix / q.Wo;
    const int wo = (int)(pix - row * q.Wo);
    const int n = (int)(row / q.Ho);
    const int ho = (int)(row - (size_t)n * q.Ho);

    M p0h_, p0w_;
    window_origin<M>(q, ho, wo, p0h_, p0w_);
    const size_t img = (size_t)n * q.H * q.W * q.C + (size_t)g * q.gc;
    const T *im = in + img;
    A *gim = gin + img;
    const T *po = off + unit * q.P * 2;
    const T *pm = mask + unit * q.P;
    const T *go = gout + pix * q.C + (size_t)g * q.gc;
    T *d_o = goff + unit * q.P * 2;
    T *d_m = gmask + unit * q.P;

    M mx = 0, inv = 1;
    M prob[LOGITS ? kMaxSoftmaxP : 1], gmv[LOGITS ? kMaxSoftmaxP : 1];
    if (LOGITS) {
        mx = -INFINITY;
        for (int p = 0; p < q.P; ++p) mx = max(mx, (M)to_math(pm[p]));
        M sum = 0;
        for (int p = 0; p < q.P; ++p) sum += exp((M)to_math(pm[p]) - mx);
        inv = (M)1 / sum;
    }
    int p = 0;
    for (int i = 0; i < q.kw; ++i)
        for (int j = 0; j < q.kh; ++j, ++p) {
            Point<M> t;
            locate<M>(q, p0h_, p0w_, i, j, (M)to_math(po[2 * p]), (M)to_math(po[2 * p + 1]), t);
            M m = to_math(pm[p]);
            if (LOGITS) m = exp(m - mx) * inv;
            M s_m = 0, s_w = 0, s_h = 0;
            if (t.bits) {  // warp-uniform: every lane sees the same point
                const size_t base = ((size_t)t.h_low * q.W + t.w_low) * q.C;
                const size_t rw = (size_t)q.W * q.C;
                const M w1 = t.hh * t.hw, w2 = t.hh * t.lw, w3 = t.lh * t.hw, w4 = t.lh * t.lw;
                for (int ch = lane; ch < q.gc; ch += 32) {
                    const M top = to_math(go[ch]);
                    const M tg = top * m;
                    M v1 = 0, v2 = 0, v3 = 0, v4 = 0;
                    if (t.bits & B_C1) { v1 = to_math(im[base + ch]); atomic_add(gim + base + ch, w1 * tg); }
                    if (t.bits & B_C2) { v2 = to_math(im[base + q.C + ch]); atomic_add(gim + base + q.C + ch, w2 * tg); }
                    if (t.bits & B_C3) { v3 = to_math(im[base + rw + ch]); atomic_add(gim + base + rw + ch, w3 * tg); }
                    if (t.bits & B_C4) { v4 = to_math(im[base + rw + q.C + ch]); atomic_add(gim + base + rw + q.C + ch, w4 * tg); }
                    const M val = w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4;
                    s_m += top * val;
                    s_w += top * (t.hh * (v2 - v1) + t.lh * (v4 - v3));
                    s_h += top * (t.hw * (v3 - v1) + t.lw * (v4 - v2));
                }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                s_m += shfl_xor(s_m, d);
                s_w += shfl_xor(s_w, d);
                s_h += shfl_xor(s_h, d);
            }
            const M sm = (M)q.scale * m;
            if (lane == 0) {
                d_o[2 * p] = from_math<T>(sm * s_w);
                d_o[2 * p + 1] = from_math<T>(sm * s_h);
                if (!LOGITS) d_m[p] = from_math<T>(s_m);
            }
            if (LOGITS) { prob[p] = m; gmv[p] = s_m; }
        }
    if (LOGITS && lane == 0) {
        M dot = 0;
        for (int k = 0; k < q.P; ++k) dot += prob[k] * gmv[k];
        for (int k = 0; k < q.P; ++k) d_m[k] = from_math<T>(prob[k] * (gmv[k] - dot));
    }
}

// ===========================================================================
// integer contract (include/dcnv3_b200.h: dcnv3_b200_debug_indices)
// ===========================================================================
template <typename T>
__global__ void __launch_bounds__(kThreads)
indices_kernel(const T *__restrict__ off, int32_t *__restrict__ hw_low,
               uint8_t *__restrict__ bounds, const Geo q, const size_t total) {
    using M = typename OpMath<T>::type;
    const size_t idx = blockIdx.x * (size_t)kThreads + threadIdx.x;  // (pix, g, p)
    if (idx >= total) return;
    const int p = (int)(idx % q.P);
    const size_t pix = idx / ((size_t)q.P * q.G);
    const size_t row = pix / q.Wo;
    const int wo = (int)(pix - row * q.Wo);
    const int ho = (int)(row % q.Ho);
    const int i = p / q.kh, j = p - i * q.kh;
    M p0h_, p0w_;
    window_origin<M>(q, ho, wo, p0h_, p0w_);
    Point<M> t;
    locate<M>(q, p0h_, p0w_, i, j, (M)to_math(off[2 * idx]), (M)to_math(off[2 * idx + 1]), t);
    hw_low[2 * idx] = t.h_low;
    hw_low[2 * idx + 1] = t.w_low;
    bounds[idx] = (uint8_t)t.bits;
}

}  // namespace dcnv3

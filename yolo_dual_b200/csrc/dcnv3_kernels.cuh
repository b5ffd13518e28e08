// dcnv3_b200 — kernels.
//
// Two families:
//   *_vec_*    one lane owns BPL bytes of channels (8, 16 or 32: LDG.64 / .128 / .256) of one
//              (n, ho, wo, g); lanes run fastest over the channel vectors of a pixel, so the lanes
//              of one group are an aligned power-of-two segment of a warp: grad_offset / grad_mask
//              reduce with __shfl_xor (no shared memory, no barriers, no zero-init, one plain store
//              per value) and grad_input goes out as vector reductions that fill whole sectors.
//   *_any_*    generic fallback for shapes the vector path cannot take (group_channels not a
//              multiple of the vector width, f64, odd alignment, runtime-sized fused softmax).
//              Still CUDA, still this library: there is no CPU path.
// (dcnv3_bwd_tile.cuh holds a third, experimental family: the shared-memory privatised backward.)
//
// Reference semantics reproduced (models/ops_dcnv3/src/cuda/dcnv3_im2col_cuda.cuh):
//   forward  dcnv3_im2col_gpu_kernel :216-275 + dcnv3_im2col_bilinear :32-80
//   backward dcnv3_col2im_gpu_kernel_* :278-839 + dcnv3_col2im_bilinear :82-147
#pragma once

#include "dcnv3_common.cuh"

namespace dcnv3 {

constexpr int kThreads = 256;
constexpr int kMaxSoftmaxP = 49;  // runtime-sized fused softmax keeps P values in registers/local

// ---------------------------------------------------------------------------
// lane chunks: BPL bytes of channels per lane (16 -> LDG.128, 32 -> LDG.256)
// ---------------------------------------------------------------------------
template <int BPL> struct Words { unsigned w[BPL / 4]; };

// predicated read-only load: zeros when !pred (an invalid corner contributes exactly 0,
// and its address is never dereferenced — same as the reference's `if (valid)` reads)
template <int BPL> __device__ __forceinline__ Words<BPL> ldg_pred(const void *p, bool pred);
template <> __device__ __forceinline__ Words<16> ldg_pred<16>(const void *p, bool pred) {
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
    if (pred) r = __ldg(reinterpret_cast<const uint4 *>(p));
    Words<16> o;
    o.w[0] = r.x; o.w[1] = r.y; o.w[2] = r.z; o.w[3] = r.w;
    return o;
}
template <> __device__ __forceinline__ Words<8> ldg_pred<8>(const void *p, bool pred) {
    uint2 r = make_uint2(0u, 0u);
    if (pred) r = __ldg(reinterpret_cast<const uint2 *>(p));
    Words<8> o;
    o.w[0] = r.x; o.w[1] = r.y;
    return o;
}
template <> __device__ __forceinline__ Words<32> ldg_pred<32>(const void *p, bool pred) {
    Words<32> o;
    asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %9, 0;\n\t"
        "mov.b32 %0, 0; mov.b32 %1, 0; mov.b32 %2, 0; mov.b32 %3, 0;\n\t"
        "mov.b32 %4, 0; mov.b32 %5, 0; mov.b32 %6, 0; mov.b32 %7, 0;\n\t"
        "@q ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\t}"
        : "=r"(o.w[0]), "=r"(o.w[1]), "=r"(o.w[2]), "=r"(o.w[3]), "=r"(o.w[4]), "=r"(o.w[5]),
          "=r"(o.w[6]), "=r"(o.w[7])
        : "l"(p), "r"((int)pred));
    return o;
}

template <typename T, int BPL> struct Lane {
    static constexpr int CH = BPL / (int)sizeof(T);  // channels per lane
    static constexpr int NP = CH / 2;                // float2 pairs per lane
};

// storage words -> op-math float2 pairs (channel 2k, 2k+1)
template <int BPL>
__device__ __forceinline__ void to_pairs(const Words<BPL> &r, float2 (&v)[BPL / 8], const float *) {
#pragma unroll
    for (int k = 0; k < BPL / 8; ++k)
        v[k] = make_float2(__uint_as_float(r.w[2 * k]), __uint_as_float(r.w[2 * k + 1]));
}
template <int BPL>
__device__ __forceinline__ void to_pairs(const Words<BPL> &r, float2 (&v)[BPL / 4], const __half *) {
#pragma unroll
    for (int k = 0; k < BPL / 4; ++k) v[k] = __half22float2(*reinterpret_cast<const __half2 *>(&r.w[k]));
}
template <int BPL>
__device__ __forceinline__ void to_pairs(const Words<BPL> &r, float2 (&v)[BPL / 4], const __nv_bfloat16 *) {
#pragma unroll
    for (int k = 0; k < BPL / 4; ++k)  // bf16 -> f32 is a 16-bit shift / mask
        v[k] = make_float2(__uint_as_float(r.w[k] << 16), __uint_as_float(r.w[k] & 0xffff0000u));
}

// op-math pairs -> storage words (round to nearest even)
template <int BPL>
__device__ __forceinline__ void from_pairs(const float2 (&v)[BPL / 8], Words<BPL> &r, const float *) {
#pragma unroll
    for (int k = 0; k < BPL / 8; ++k) { r.w[2 * k] = __float_as_uint(v[k].x); r.w[2 * k + 1] = __float_as_uint(v[k].y); }
}
template <int BPL>
__device__ __forceinline__ void from_pairs(const float2 (&v)[BPL / 4], Words<BPL> &r, const __half *) {
#pragma unroll
    for (int k = 0; k < BPL / 4; ++k) { const __half2 h = __float22half2_rn(v[k]); r.w[k] = *reinterpret_cast<const unsigned *>(&h); }
}
template <int BPL>
__device__ __forceinline__ void from_pairs(const float2 (&v)[BPL / 4], Words<BPL> &r, const __nv_bfloat16 *) {
#pragma unroll
    for (int k = 0; k < BPL / 4; ++k) { const __nv_bfloat162 h = __float22bfloat162_rn(v[k]); r.w[k] = *reinterpret_cast<const unsigned *>(&h); }
}

template <int BPL> __device__ __forceinline__ void st_words(void *p, const Words<BPL> &r) {
    if constexpr (BPL == 8) {
        *reinterpret_cast<uint2 *>(p) = make_uint2(r.w[0], r.w[1]);
    } else {
#pragma unroll
        for (int k = 0; k < BPL / 16; ++k)
            reinterpret_cast<uint4 *>(p)[k] = make_uint4(r.w[4 * k], r.w[4 * k + 1], r.w[4 * k + 2], r.w[4 * k + 3]);
    }
}

// Decode a flat lane index into (pixel, lane-in-pixel, group, n, ho, wo).
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
// forward, vector path: one lane owns BPL bytes of channels of one (n, ho, wo, g)
// ===========================================================================
// one lane's work: BPL bytes of channels (vector c.v of pixel c.pix, group c.g), all points
template <typename T, int BPL, int KP, bool LOGITS>
__device__ __forceinline__ void fwd_vec_body(const VecCoord &c, const T *__restrict__ in, const T *__restrict__ off,
                                             const T *__restrict__ mask, T *__restrict__ out, const Geo &q) {
    constexpr int CH = Lane<T, BPL>::CH, NP = Lane<T, BPL>::NP;

    float p0h_, p0w_;
    window_origin<float>(q, c.ho, c.wo, p0h_, p0w_);
    const char *im = reinterpret_cast<const char *>(in + (size_t)c.n * q.H * q.W * q.C + c.v * CH);
    const int sC = q.C * (int)sizeof(T);  // byte strides of one pixel / one row
    const int sW = q.W * sC;
    const T *po = off + (size_t)c.pix * q.opitch + c.g * q.P * 2;
    const T *pm = mask + (size_t)c.pix * q.mpitch + c.g * q.P;

    float mx = 0.f, inv = 1.f;
    if (LOGITS) softmax_stats<T, KP>(pm, q.P, mx, inv);

    float2 acc[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) acc[k] = make_float2(0.f, 0.f);

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
            {
                const float hm = t.hh * m, lm = t.lh * m;
                const float w1 = hm * t.hw, w2 = hm * t.lw, w3 = lm * t.hw, w4 = lm * t.lw;
                const char *r1 = im + (t.h_low * q.W + t.w_low) * sC;
                const Words<BPL> c1 = ldg_pred<BPL>(r1, t.ok1);
                const Words<BPL> c2 = ldg_pred<BPL>(r1 + sC, t.ok2);
                const Words<BPL> c3 = ldg_pred<BPL>(r1 + sW, t.ok3);
                const Words<BPL> c4 = ldg_pred<BPL>(r1 + sW + sC, t.ok4);
                float2 v[NP];
                to_pairs<BPL>(c1, v, (const T *)nullptr);
#pragma unroll
                for (int k = 0; k < NP; ++k) acc[k] = __ffma2_rn(v[k], make_float2(w1, w1), acc[k]);
                to_pairs<BPL>(c2, v, (const T *)nullptr);
#pragma unroll
                for (int k = 0; k < NP; ++k) acc[k] = __ffma2_rn(v[k], make_float2(w2, w2), acc[k]);
                to_pairs<BPL>(c3, v, (const T *)nullptr);
#pragma unroll
                for (int k = 0; k < NP; ++k) acc[k] = __ffma2_rn(v[k], make_float2(w3, w3), acc[k]);
                to_pairs<BPL>(c4, v, (const T *)nullptr);
#pragma unroll
                for (int k = 0; k < NP; ++k) acc[k] = __ffma2_rn(v[k], make_float2(w4, w4), acc[k]);
            }
        }
    }
    Words<BPL> r;
    from_pairs<BPL>(acc, r, (const T *)nullptr);
    st_words<BPL>(out + (size_t)c.pix * q.C + c.v * CH, r);
}

template <typename T, int BPL, int KP, bool LOGITS>
__global__ void __launch_bounds__(kThreads)
fwd_vec_kernel(const T *__restrict__ in, const T *__restrict__ off, const T *__restrict__ mask,
               T *__restrict__ out, const Geo q, const int vec_per_pix,
               const int lanes_per_group, const unsigned total) {
    pdl_enter();
    const unsigned idx = blockIdx.x * (unsigned)kThreads + threadIdx.x;
    if (idx >= total) return;
    const VecCoord c = decode_vec(idx, q, vec_per_pix, lanes_per_group);
    fwd_vec_body<T, BPL, KP, LOGITS>(c, in, off, mask, out, q);
}

// ===========================================================================
// forward, point-split path (16-bit storage, group_channels = 16, 3x3): the two lanes of a
// (n, ho, wo, g) split the nine sampling POINTS instead of the sixteen channels.
//
// fwd_vec_kernel gives each of the two lanes 8 channels and lets both of them locate all nine points:
// per point ~70 of its ~135 instructions are location / weight / address arithmetic that the partner
// lane repeats (ncu: issue slots 77 % busy, the binding resource).  Here lane h takes points 4h..4h+3
// with all 16 channels (one 32-byte LDG.256 per corner = the group's whole slab = one sector) and the
// lanes share point 8 by channel halves; the partial sums meet in one shuffle exchange at the end.
// 4.5 locates per lane instead of 9, the same number of sectors requested from L1.
// Semantics and arithmetic per point are fwd_vec_kernel's (shared locate(), fp32 accumulation).
// ===========================================================================
template <typename T, bool LOGITS>
__global__ void __launch_bounds__(kThreads)
fwd_pts_kernel(const T *__restrict__ in, const T *__restrict__ off, const T *__restrict__ mask,
               T *__restrict__ out, const Geo q, const unsigned total) {
    static_assert(sizeof(T) == 2, "point-split forward: 16-bit storage");
    pdl_enter();
    unsigned idx = blockIdx.x * (unsigned)kThreads + threadIdx.x;
    const bool active = idx < total;  // tail lanes stay for the shuffle (total is even: pairs are whole)
    if (!active) idx = total - 2u + (idx & 1u);
    const unsigned unit = idx >> 1;  // (pixel, group)
    const int h = (int)(idx & 1u);
    const unsigned pix = unit / (unsigned)q.G;
    const int g = (int)(unit - pix * (unsigned)q.G);
    const unsigned row = pix / (unsigned)q.Wo;
    const int wo = (int)(pix - row * (unsigned)q.Wo);
    const int n = (int)(row / (unsigned)q.Ho);
    const int ho = (int)(row - (unsigned)n * (unsigned)q.Ho);

    float p0h_, p0w_;
    window_origin<float>(q, ho, wo, p0h_, p0w_);
    const char *im = reinterpret_cast<const char *>(in + (size_t)n * q.H * q.W * q.C + g * 16);
    const int sC = q.C * 2, sW = q.W * sC;
    const T *po = off + (size_t)unit * 18;
    const T *pm = mask + (size_t)unit * 9;

    float mx = 0.f, inv = 1.f;
    if (LOGITS) softmax_stats<T, 9>(pm, 9, mx, inv);

    float2 acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = make_float2(0.f, 0.f);

    // ---- four whole points: p = 4h + k  (i = p / 3 indexes kernel_w, j = p % 3 kernel_h, cuh:253-254)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int p = 4 * h + k;
        const int i = h ? (4 + k) / 3 : k / 3, j = h ? (4 + k) % 3 : k % 3;
        const float2 o = load_offset_pair(po + 2 * p);
        Point<float> t;
        locate<float>(q, p0h_, p0w_, i, j, o.x, o.y, t);
        float m = to_math(pm[p]);
        if (LOGITS) m = expf(m - mx) * inv;
        const float hm = t.hh * m, lm = t.lh * m;
        const float w1 = hm * t.hw, w2 = hm * t.lw, w3 = lm * t.hw, w4 = lm * t.lw;
        const char *r1 = im + (t.h_low * q.W + t.w_low) * sC;
        const Words<32> c1 = ldg_pred<32>(r1, t.ok1);
        const Words<32> c2 = ldg_pred<32>(r1 + sC, t.ok2);
        const Words<32> c3 = ldg_pred<32>(r1 + sW, t.ok3);
        const Words<32> c4 = ldg_pred<32>(r1 + sW + sC, t.ok4);
        float2 v[8];
        to_pairs<32>(c1, v, (const T *)nullptr);
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = __ffma2_rn(v[c], make_float2(w1, w1), acc[c]);
        to_pairs<32>(c2, v, (const T *)nullptr);
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = __ffma2_rn(v[c], make_float2(w2, w2), acc[c]);
        to_pairs<32>(c3, v, (const T *)nullptr);
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = __ffma2_rn(v[c], make_float2(w3, w3), acc[c]);
        to_pairs<32>(c4, v, (const T *)nullptr);
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = __ffma2_rn(v[c], make_float2(w4, w4), acc[c]);
    }

    // ---- exchange: lane h keeps channels 8h..8h+7 and gets the partner's partial sums for them
    float2 mine[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float2 keep = h ? acc[4 + c] : acc[c];
        const float2 send = h ? acc[c] : acc[4 + c];
        mine[c].x = keep.x + __shfl_xor_sync(0xffffffffu, send.x, 1);
        mine[c].y = keep.y + __shfl_xor_sync(0xffffffffu, send.y, 1);
    }

    // ---- point 8 (i = 2, j = 2): both lanes, 8 channels each
    {
        const float2 o = load_offset_pair(po + 16);
        Point<float> t;
        locate<float>(q, p0h_, p0w_, 2, 2, o.x, o.y, t);
        float m = to_math(pm[8]);
        if (LOGITS) m = expf(m - mx) * inv;
        const float hm = t.hh * m, lm = t.lh * m;
        const float w1 = hm * t.hw, w2 = hm * t.lw, w3 = lm * t.hw, w4 = lm * t.lw;
        const char *r1 = im + (t.h_low * q.W + t.w_low) * sC + 16 * h;
        const Words<16> c1 = ldg_pred<16>(r1, t.ok1);
        const Words<16> c2 = ldg_pred<16>(r1 + sC, t.ok2);
        const Words<16> c3 = ldg_pred<16>(r1 + sW, t.ok3);
        const Words<16> c4 = ldg_pred<16>(r1 + sW + sC, t.ok4);
        float2 v[4];
        to_pairs<16>(c1, v, (const T *)nullptr);
#pragma unroll
        for (int c = 0; c < 4; ++c) mine[c] = __ffma2_rn(v[c], make_float2(w1, w1), mine[c]);
        to_pairs<16>(c2, v, (const T *)nullptr);
#pragma unroll
        for (int c = 0; c < 4; ++c) mine[c] = __ffma2_rn(v[c], make_float2(w2, w2), mine[c]);
        to_pairs<16>(c3, v, (const T *)nullptr);
#pragma unroll
        for (int c = 0; c < 4; ++c) mine[c] = __ffma2_rn(v[c], make_float2(w3, w3), mine[c]);
        to_pairs<16>(c4, v, (const T *)nullptr);
#pragma unroll
        for (int c = 0; c < 4; ++c) mine[c] = __ffma2_rn(v[c], make_float2(w4, w4), mine[c]);
    }
    Words<16> r;
    from_pairs<16>(mine, r, (const T *)nullptr);
    if (active) st_words<16>(out + (size_t)pix * q.C + g * 16 + 8 * h, r);
}

// ===========================================================================
// backward, vector path.  A = accumulation type of grad_input (float: gin itself for f32
// storage or the fp32 workspace for 16-bit storage; T: packed 16-bit reductions).
//
// Per sampling point only the four corner dot products d_k = sum_c go[c] * v_k[c] are needed:
//   grad_mask          = w1 d1 + w2 d2 + w3 d3 + w4 d4                      (cuh:144)
//   sum_c go*grad_w_w  = hh (d2 - d1) + lh (d4 - d3)                        (cuh:114-139,145)
//   sum_c go*grad_h_w  = hw (d3 - d1) + lw (d4 - d2)                        (cuh:114-139,146)
// and the lanes of a group reduce the three sums with __shfl_xor.
//
// grad_input reductions are issued so that the L lanes of a group cover one contiguous
// L*16-byte range per instruction (whole 32-byte sectors): instruction j of lane l carries
// 16-byte chunk j*L + l of the group's slab, with the matching grad_output values re-read
// from global memory once per thread (L1 hits).
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

// one 16-byte reduction chunk: CPQ channels of accumulation type A
template <typename A> struct RedChunk;
template <> struct RedChunk<float> {
    static constexpr int CPQ = 4;
    __device__ static __forceinline__ void run(char *dst, const float (&g)[4], float w, bool pred) {
        const float2 a = __fmul2_rn(make_float2(g[0], g[1]), make_float2(w, w));
        const float2 b = __fmul2_rn(make_float2(g[2], g[3]), make_float2(w, w));
        red_add_v4_f32(reinterpret_cast<float *>(dst), a.x, a.y, b.x, b.y, pred);
    }
};
template <> struct RedChunk<__half> {
    static constexpr int CPQ = 8;
    __device__ static __forceinline__ void run(char *dst, const float (&g)[8], float w, bool pred) {
        float t[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) t[k] = w * g[k];
        const uint4 v = pack(t, (const __half *)nullptr);
        red_add_v4_f16x2(reinterpret_cast<__half *>(dst), v, pred);
    }
};
template <> struct RedChunk<__nv_bfloat16> {
    static constexpr int CPQ = 8;
    __device__ static __forceinline__ void run(char *dst, const float (&g)[8], float w, bool pred) {
        float t[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) t[k] = w * g[k];
        const uint4 v = pack(t, (const __nv_bfloat16 *)nullptr);
        red_add_v4_bf16x2(reinterpret_cast<__nv_bfloat16 *>(dst), v, pred);
    }
};

// CPQ consecutive channels of grad_output as op-math floats
template <typename T, int CPQ>
__device__ __forceinline__ void load_go_chunk(const T *p, float (&g)[CPQ]) {
    if constexpr (sizeof(T) == 4) {
        static_assert(CPQ == 4, "");
        const float4 r = __ldg(reinterpret_cast<const float4 *>(p));
        g[0] = r.x; g[1] = r.y; g[2] = r.z; g[3] = r.w;
    } else if constexpr (CPQ == 4) {
        const uint2 r = __ldg(reinterpret_cast<const uint2 *>(p));
        Words<8> w; w.w[0] = r.x; w.w[1] = r.y;
        float2 v[2];
        to_pairs<8>(w, v, (const T *)nullptr);
        g[0] = v[0].x; g[1] = v[0].y; g[2] = v[1].x; g[3] = v[1].y;
    } else {
        static_assert(CPQ == 8, "");
        const uint4 r = __ldg(reinterpret_cast<const uint4 *>(p));
        Words<16> w; w.w[0] = r.x; w.w[1] = r.y; w.w[2] = r.z; w.w[3] = r.w;
        float2 v[4];
        to_pairs<16>(w, v, (const T *)nullptr);
#pragma unroll
        for (int k = 0; k < 4; ++k) { g[2 * k] = v[k].x; g[2 * k + 1] = v[k].y; }
    }
}

// One lane's work: BPL bytes of channels (vector c.v of pixel c.pix, group c.g), all points.  `active` false:
// the lane only takes part in the shuffles (its coordinates must still be valid ones).
template <typename T, typename A, int BPL, int KP, bool LOGITS>
__device__ __forceinline__ void
bwd_vec_lane(const VecCoord &c, const int lane_in_group, const bool active, const T *__restrict__ in,
             const T *__restrict__ off, const T *__restrict__ mask, const T *__restrict__ gout, A *__restrict__ gin,
             T *__restrict__ goff, T *__restrict__ gmask, const Geo &q, const int lanes_per_group) {
    static_assert(!LOGITS || KP > 0, "fused softmax in the vector path needs a compile-time P");
    constexpr int CH = Lane<T, BPL>::CH, NP = Lane<T, BPL>::NP;
    constexpr int CPQ = RedChunk<A>::CPQ;                 // channels per 16-byte reduction
    constexpr int R = CH * (int)sizeof(A) / 16;           // reductions per corner per lane
    static_assert(R >= 1, "a lane must own at least one 16-byte reduction chunk");

    float p0h_, p0w_;
    window_origin<float>(q, c.ho, c.wo, p0h_, p0w_);
    const size_t img = (size_t)c.n * q.H * q.W * q.C;
    const char *im = reinterpret_cast<const char *>(in + img + c.v * CH);
    char *gim = reinterpret_cast<char *>(gin + img + c.g * q.gc);  // group slab start, A units
    const int sC = q.C * (int)sizeof(T), sW = q.W * sC;
    constexpr int ARATIO = (int)sizeof(A) / (int)sizeof(T) > 0 ? (int)sizeof(A) / (int)sizeof(T) : 1;
    static_assert(sizeof(A) >= sizeof(T), "");
    const size_t o_el = (size_t)c.pix * q.opitch + c.g * q.P * 2, m_el = (size_t)c.pix * q.mpitch + c.g * q.P;
    const T *po = off + o_el;
    const T *pm = mask + m_el;
    T *d_o = goff + o_el;
    T *d_m = gmask + m_el;
    const bool writer = active && lane_in_group == 0;

    // grad_output: this lane's channels (for the dots) ...
    const T *gop = gout + (size_t)c.pix * q.C;
    float2 gp[NP];
    {
        const Words<BPL> gw = ldg_pred<BPL>(gop + c.v * CH, true);
        to_pairs<BPL>(gw, gp, (const T *)nullptr);
    }
    // ... and the channels of the reduction chunks this lane issues (chunk j*L + l)
    float gr[R][CPQ];
    int red_off[R];  // byte offset of the chunk inside the group's slab (A units)
#pragma unroll
    for (int j = 0; j < R; ++j) {
        const int chunk = j * lanes_per_group + lane_in_group;
        load_go_chunk<T, CPQ>(gop + c.g * q.gc + chunk * CPQ, gr[j]);
        red_off[j] = chunk * 16;
    }

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

            float s_m, s_w, s_h;
            {
                const int e0 = (t.h_low * q.W + t.w_low) * sC;  // byte offset of corner 1 (T units)
                const char *r1 = im + e0;
                const Words<BPL> c1 = ldg_pred<BPL>(r1, t.ok1);
                const Words<BPL> c2 = ldg_pred<BPL>(r1 + sC, t.ok2);
                const Words<BPL> c3 = ldg_pred<BPL>(r1 + sW, t.ok3);
                const Words<BPL> c4 = ldg_pred<BPL>(r1 + sW + sC, t.ok4);
                float d[4];
                {
                    float2 v[NP];
                    float2 a;
#define DCNV3_DOT(CW, K)                                                          \
    to_pairs<BPL>(CW, v, (const T *)nullptr);                                     \
    a = make_float2(0.f, 0.f);                                                    \
    _Pragma("unroll") for (int k = 0; k < NP; ++k) a = __ffma2_rn(gp[k], v[k], a); \
    d[K] = a.x + a.y;
                    DCNV3_DOT(c1, 0)
                    DCNV3_DOT(c2, 1)
                    DCNV3_DOT(c3, 2)
                    DCNV3_DOT(c4, 3)
#undef DCNV3_DOT
                }
                const float w1 = t.hh * t.hw, w2 = t.hh * t.lw, w3 = t.lh * t.hw, w4 = t.lh * t.lw;
                s_m = w1 * d[0] + w2 * d[1] + w3 * d[2] + w4 * d[3];
                s_w = t.hh * (d[1] - d[0]) + t.lh * (d[3] - d[2]);
                s_h = t.hw * (d[2] - d[0]) + t.lw * (d[3] - d[1]);
                {  // cuh:116,124,132,140: grad_im[corner] += w_k * top_grad * mask
                    char *g1 = gim + (size_t)e0 * ARATIO;
                    const int aC = sC * ARATIO, aW = sW * ARATIO;
#pragma unroll
                    for (int jj = 0; jj < R; ++jj) {
                        RedChunk<A>::run(g1 + red_off[jj], gr[jj], w1 * m, active && t.ok1);
                        RedChunk<A>::run(g1 + aC + red_off[jj], gr[jj], w2 * m, active && t.ok2);
                        RedChunk<A>::run(g1 + aW + red_off[jj], gr[jj], w3 * m, active && t.ok3);
                        RedChunk<A>::run(g1 + aW + aC + red_off[jj], gr[jj], w4 * m, active && t.ok4);
                    }
                }
            }
            // sum over the channels of the group: its lanes are an aligned warp segment
            for (int dlt = lanes_per_group >> 1; dlt > 0; dlt >>= 1) {
                s_m += shfl_xor(s_m, dlt);
                s_w += shfl_xor(s_w, dlt);
                s_h += shfl_xor(s_h, dlt);
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

template <typename T, typename A, int BPL, int KP, bool LOGITS>
__device__ __forceinline__ void
bwd_vec_body(const unsigned block, const T *__restrict__ in, const T *__restrict__ off, const T *__restrict__ mask,
             const T *__restrict__ gout, A *__restrict__ gin, T *__restrict__ goff,
             T *__restrict__ gmask, const Geo &q, const int vec_per_pix,
             const int lanes_per_group, const unsigned total) {
    unsigned idx = block * (unsigned)kThreads + threadIdx.x;
    const bool active = idx < total;  // tail lanes stay for the shuffles
    if (!active) idx = total - 1;
    const VecCoord c = decode_vec(idx, q, vec_per_pix, lanes_per_group);
    bwd_vec_lane<T, A, BPL, KP, LOGITS>(c, c.v - c.g * lanes_per_group, active, in, off, mask, gout, gin, goff, gmask,
                                        q, lanes_per_group);
}

// `per_cta` consecutive logical blocks per CTA (1 except for the selector-guarded launch, where a
// kernel that returns at once should not cost 12800 CTA launches).  `sel`: family selector written by
// imat::zero_select_kernel on the same stream (this kernel runs when it reads 1).
template <typename T, typename A, int BPL, int KP, bool LOGITS>
__global__ void __launch_bounds__(kThreads)
bwd_vec_kernel(const T *__restrict__ in, const T *__restrict__ off, const T *__restrict__ mask,
               const T *__restrict__ gout, A *__restrict__ gin, T *__restrict__ goff,
               T *__restrict__ gmask, const Geo q, const int vec_per_pix,
               const int lanes_per_group, const unsigned total, const int *sel = nullptr,
               const unsigned per_cta = 1, const unsigned n_blocks = 0) {
    // Wait first, like every kernel of the chain.  (Reading `sel` ahead of the wait and letting unneeded CTAs
    // leave at once is legal — zero_select_kernel has completed by then — but it lets the NEXT kernel's
    // thousands of CTAs become resident and sit in their own wait while bwd_imat_kernel is still in its
    // tail, which slowed the back-to-back step by 20-170 us.  Waiting CTAs are not free.)
    pdl_enter();
    if (sel != nullptr && __ldcg(sel) != 1) return;  // written by the previous kernel of the PDL chain: coherent load, never .nc
    if (per_cta == 1) {
        bwd_vec_body<T, A, BPL, KP, LOGITS>(blockIdx.x, in, off, mask, gout, gin, goff, gmask, q, vec_per_pix, lanes_per_group, total);
        return;
    }
    for (unsigned b = blockIdx.x * per_cta, e = min(b + per_cta, n_blocks); b < e; ++b)
        bwd_vec_body<T, A, BPL, KP, LOGITS>(b, in, off, mask, gout, gin, goff, gmask, q, vec_per_pix, lanes_per_group, total);
}

// fp32 workspace -> 16-bit grad_input (ACC_OPMATH), 8 elements per thread
template <typename T>
__global__ void __launch_bounds__(kThreads)
cast_ws_kernel(const float *ws, T *__restrict__ dst, const size_t n_vec8,
               const size_t n_total) {
    pdl_enter();
    // `ws` was written by the previous kernels of the PDL chain while this one was already resident: it is neither
    // const __restrict__ nor read through the non-coherent path (ld.global.nc is only defined for data that is
    // read-only for the kernel's whole lifetime) — __ldcg = ld.global.cg, served by L2 where the reductions landed.
    const size_t i = blockIdx.x * (size_t)kThreads + threadIdx.x;
    if (i < n_vec8) {
        const float4 a = __ldcg(reinterpret_cast<const float4 *>(ws + i * 8));
        const float4 b = __ldcg(reinterpret_cast<const float4 *>(ws + i * 8 + 4));
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        *reinterpret_cast<uint4 *>(dst + i * 8) = pack(v, (const T *)nullptr);
    } else if (i == n_vec8) {  // scalar tail (generic path only)
        for (size_t k = n_vec8 * 8; k < n_total; ++k) dst[k] = from_math<T>(__ldcg(ws + k));
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
            if (!t.inside) continue;
            M m = to_math(pm[p]);
            if (LOGITS) m = exp(m - mx) * inv;
            const size_t base = ((size_t)t.h_low * q.W + t.w_low) * q.C;
            const size_t rw = (size_t)q.W * q.C;
            const M v1 = t.ok1 ? (M)to_math(im[base]) : (M)0;
            const M v2 = t.ok2 ? (M)to_math(im[base + q.C]) : (M)0;
            const M v3 = t.ok3 ? (M)to_math(im[base + rw]) : (M)0;
            const M v4 = t.ok4 ? (M)to_math(im[base + rw + q.C]) : (M)0;
            acc += (t.hh * t.hw * v1 + t.hh * t.lw * v2 + t.lh * t.hw * v3 + t.lh * t.lw * v4) * m;
        }
    out[idx] = from_math<T>(acc);
}

// `lpu` lanes per (pixel, g) unit (a power of two <= 32, >= min(group_channels, 32) rounded up):
// a warp holds 32/lpu units; the lanes of a unit stride over its channels and reduce with shuffles
template <typename T, typename A, bool LOGITS>
__global__ void __launch_bounds__(kThreads)
bwd_any_kernel(const T *__restrict__ in, const T *__restrict__ off, const T *__restrict__ mask,
               const T *__restrict__ gout, A *__restrict__ gin, T *__restrict__ goff,
               T *__restrict__ gmask, const Geo q, const size_t n_units, const int lpu) {
    using M = typename OpMath<T>::type;
    const size_t tid = blockIdx.x * (size_t)kThreads + threadIdx.x;
    size_t unit = tid / lpu;
    const int cl = (int)(tid - unit * lpu);  // lane inside the unit
    const bool active = unit < n_units;      // tail lanes stay for the shuffles
    if (!active) unit = n_units - 1;
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
    const bool writer = active && cl == 0;

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
            if (t.inside) {  // uniform inside a unit
                const size_t base = ((size_t)t.h_low * q.W + t.w_low) * q.C;
                const size_t rw = (size_t)q.W * q.C;
                const M w1 = t.hh * t.hw, w2 = t.hh * t.lw, w3 = t.lh * t.hw, w4 = t.lh * t.lw;
                for (int ch = cl; ch < q.gc; ch += lpu) {
                    const M top = to_math(go[ch]);
                    const M tg = top * m;
                    M v1 = 0, v2 = 0, v3 = 0, v4 = 0;
                    if (t.ok1) { v1 = to_math(im[base + ch]); if (active) atomic_add(gim + base + ch, w1 * tg); }
                    if (t.ok2) { v2 = to_math(im[base + q.C + ch]); if (active) atomic_add(gim + base + q.C + ch, w2 * tg); }
                    if (t.ok3) { v3 = to_math(im[base + rw + ch]); if (active) atomic_add(gim + base + rw + ch, w3 * tg); }
                    if (t.ok4) { v4 = to_math(im[base + rw + q.C + ch]); if (active) atomic_add(gim + base + rw + q.C + ch, w4 * tg); }
                    const M val = w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4;
                    s_m += top * val;
                    s_w += top * (t.hh * (v2 - v1) + t.lh * (v4 - v3));
                    s_h += top * (t.hw * (v3 - v1) + t.lw * (v4 - v2));
                }
            }
            for (int d = lpu >> 1; d > 0; d >>= 1) {
                s_m += shfl_xor(s_m, d);
                s_w += shfl_xor(s_w, d);
                s_h += shfl_xor(s_h, d);
            }
            const M sm = (M)q.scale * m;
            if (writer) {
                d_o[2 * p] = from_math<T>(sm * s_w);
                d_o[2 * p + 1] = from_math<T>(sm * s_h);
                if (!LOGITS) d_m[p] = from_math<T>(s_m);
            }
            if (LOGITS) { prob[p] = m; gmv[p] = s_m; }
        }
    if (LOGITS && writer) {
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
    bounds[idx] = (uint8_t)t.bits();
}

}  // namespace dcnv3

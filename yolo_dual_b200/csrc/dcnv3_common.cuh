// dcnv3_b200 — device-side building blocks shared by every kernel.
//
// Semantics follow the reference CUDA kernels
// (/root/reference/models/ops_dcnv3/src/cuda/dcnv3_im2col_cuda.cuh); each block
// cites the lines it reproduces.  Nothing here is derived from that file's
// structure: the reference is one-thread-per-scalar SIMT, this is a
// channel-vector-per-lane design (8 / 16 / 32 bytes) for sm_100a.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dcnv3 {

// ---------------------------------------------------------------------------
// geometry passed by value to every kernel
// ---------------------------------------------------------------------------
struct Geo {
    int N, H, W, G, gc, C;
    int kh, kw, sh, sw, ph, pw, dh, dw;
    int Ho, Wo, P;
    int half_h, half_w;  // (dilation*(kernel-1)) >> 1, cuh:232,235
    float scale;         // offset_scale
    // elements between one output pixel's offsets (masks) and the next pixel's: G*P*2 (G*P) for the reference's separate
    // tensors, 3*G*P for the packed heads tensor [N,Ho,Wo, offsets | masks] (dcnv3_b200_*_packed).  Honoured by the
    // kernels those entry points can reach: fwd_tile_kernel, bwd_win_kernel and the vector lane bodies inside them.
    int opitch, mpitch;
};

// storage type -> op-math type (at::opmath_type, cuh:30): float for 16/32-bit, double for f64
template <typename T> struct OpMath { using type = float; };
template <> struct OpMath<double> { using type = double; };

__device__ __forceinline__ float to_math(float v) { return v; }
__device__ __forceinline__ float to_math(__half v) { return __half2float(v); }
__device__ __forceinline__ float to_math(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ double to_math(double v) { return v; }

template <typename T> __device__ __forceinline__ T from_math(float v);
template <> __device__ __forceinline__ float from_math<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_math<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_math<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <typename T> __device__ __forceinline__ T from_math(double v);
template <> __device__ __forceinline__ double from_math<double>(double v) { return v; }

// Separately rounded IEEE ops: the location arithmetic below must never be
// contracted into FMAs, or floor() flips on cell borders and the corner
// indices stop being bit-exact against the CPU oracle (SURVEY §7 hard part 1).
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ int floor_to_int(float v) { return __float2int_rd(v); }
__device__ __forceinline__ int floor_to_int(double v) { return __double2int_rd(v); }

// ---------------------------------------------------------------------------
// One sampling point: location (cuh:249-260), gate (cuh:262-263), corners and
// validity (cuh:39-42,57,62,67,72), fractional weights (cuh:44-46).
// ---------------------------------------------------------------------------
enum : unsigned { B_INSIDE = 1u, B_C1 = 2u, B_C2 = 4u, B_C3 = 8u, B_C4 = 16u };

template <typename M> struct Point {
    int h_low, w_low;        // (0, 0) when the gate is closed
    M lh, lw, hh, hw;        // all 0 when the gate is closed: every corner weight is then 0
    bool inside;             // gate, cuh:262-263
    bool ok1, ok2, ok3, ok4; // corner validity (false when the gate is closed)
    __device__ __forceinline__ unsigned bits() const {  // the bounds byte of the integer contract
        return (inside ? B_INSIDE : 0u) | (ok1 ? B_C1 : 0u) | (ok2 ? B_C2 : 0u) |
               (ok3 ? B_C3 : 0u) | (ok4 ? B_C4 : 0u);
    }
};

// Top-left of the kernel window in op-math, cuh:232-236 and :249-252.
template <typename M>
__device__ __forceinline__ void window_origin(const Geo &q, int ho, int wo, M &p0h_, M &p0w_) {
    const M s = (M)q.scale;
    const int p0_w = q.half_w - q.pw + wo * q.sw;
    const int p0_h = q.half_h - q.ph + ho * q.sh;
    p0w_ = sub_rn((M)p0_w, mul_rn((M)q.half_w, s));
    p0h_ = sub_rn((M)p0_h, mul_rn((M)q.half_h, s));
}

// i indexes kernel_w (outer loop), j kernel_h (inner): p = i*kh + j, cuh:253-254.
// Branch-free on purpose: with no divergent region per point the compiler can hoist the
// corner loads of several points ahead of their use (memory-level parallelism per thread).
// A closed gate yields zero weights and false validity, so nothing is read and 0 is added.
template <typename M>
__device__ __forceinline__ void locate(const Geo &q, M p0h_, M p0w_, int i, int j, M off_w,
                                       M off_h, Point<M> &t) {
    const M s = (M)q.scale;
    const M loc_w = add_rn(p0w_, mul_rn(add_rn((M)(i * q.dw), off_w), s));
    const M loc_h = add_rn(p0h_, mul_rn(add_rn((M)(j * q.dh), off_h), s));
    const bool inside = loc_h > (M)-1 && loc_w > (M)-1 && loc_h < (M)q.H && loc_w < (M)q.W;
    // float -> int saturates for out-of-range / NaN inputs; the results are discarded then
    const int h_low = floor_to_int(loc_h);
    const int w_low = floor_to_int(loc_w);
    const M lh = sub_rn(loc_h, (M)h_low);
    const M lw = sub_rn(loc_w, (M)w_low);
    t.inside = inside;
    t.h_low = inside ? h_low : 0;
    t.w_low = inside ? w_low : 0;
    t.lh = inside ? lh : (M)0;
    t.lw = inside ? lw : (M)0;
    t.hh = inside ? sub_rn((M)1, lh) : (M)0;
    t.hw = inside ? sub_rn((M)1, lw) : (M)0;
    const bool h0 = h_low >= 0, w0 = w_low >= 0;
    const bool h1 = h_low + 1 <= q.H - 1, w1 = w_low + 1 <= q.W - 1;
    t.ok1 = inside && h0 && w0;
    t.ok2 = inside && h0 && w1;
    t.ok3 = inside && h1 && w0;
    t.ok4 = inside && h1 && w1;
}

// ---------------------------------------------------------------------------
// packing helpers
// ---------------------------------------------------------------------------
// pack op-math floats back to 16 bytes of storage (round to nearest even)
__device__ __forceinline__ uint4 pack(const float (&v)[8], const __half *) {
    unsigned w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const __half2 h = __floats2half2_rn(v[2 * k], v[2 * k + 1]);
        w[k] = *reinterpret_cast<const unsigned *>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ uint4 pack(const float (&v)[8], const __nv_bfloat16 *) {
    unsigned w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
        w[k] = *reinterpret_cast<const unsigned *>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// (offset_x, offset_y) of one point: one 8-byte (f32) or 4-byte (16-bit) load
__device__ __forceinline__ float2 load_offset_pair(const float *p) {
    return __ldg(reinterpret_cast<const float2 *>(p));
}
__device__ __forceinline__ float2 load_offset_pair(const __half *p) {
    const unsigned u = __ldg(reinterpret_cast<const unsigned *>(p));
    return __half22float2(*reinterpret_cast<const __half2 *>(&u));
}
__device__ __forceinline__ float2 load_offset_pair(const __nv_bfloat16 *p) {
    const unsigned u = __ldg(reinterpret_cast<const unsigned *>(p));
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

// ---------------------------------------------------------------------------
// vector reductions into global memory (sm_90+ PTX; REDG.E.ADD.{F32x4,F16x8,BF16x8})
// ---------------------------------------------------------------------------
// The predicate lives inside the asm so each reduction is ONE predicated REDG (inline asm
// under a C++ `if` becomes a branch + reconvergence region per reduction).
__device__ __forceinline__ void red_add_v4_f32(float *p, float a, float b, float c, float d,
                                               bool pred = true) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                 "@q red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n\t}" ::"l"(p), "f"(a), "f"(b),
                 "f"(c), "f"(d), "r"((int)pred)
                 : "memory");
}
__device__ __forceinline__ void red_add_v2_f32(float *p, float a, float b, bool pred = true) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\t"
                 "@q red.global.add.v2.f32 [%0], {%1, %2};\n\t}" ::"l"(p), "f"(a), "f"(b), "r"((int)pred)
                 : "memory");
}
__device__ __forceinline__ void red_add_v4_f16x2(__half *p, const uint4 &v, bool pred = true) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                 "@q red.global.add.noftz.v4.f16x2 [%0], {%1, %2, %3, %4};\n\t}" ::"l"(p), "r"(v.x),
                 "r"(v.y), "r"(v.z), "r"(v.w), "r"((int)pred)
                 : "memory");
}
__device__ __forceinline__ void red_add_v4_bf16x2(__nv_bfloat16 *p, const uint4 &v, bool pred = true) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t"
                 "@q red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};\n\t}" ::"l"(p), "r"(v.x),
                 "r"(v.y), "r"(v.z), "r"(v.w), "r"((int)pred)
                 : "memory");
}

// scalar atomics for the generic path
__device__ __forceinline__ void atomic_add(float *p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void atomic_add(double *p, double v) { atomicAdd(p, v); }
__device__ __forceinline__ void atomic_add(__half *p, float v) { atomicAdd(p, __float2half_rn(v)); }
__device__ __forceinline__ void atomic_add(__nv_bfloat16 *p, float v) { atomicAdd(p, __float2bfloat16_rn(v)); }

// ---------------------------------------------------------------------------
// Programmatic dependent launch (sm_90+).  Every kernel of the vector / imat families starts with
// pdl_enter(): wait until the previous kernel of the stream has completed and its writes are visible,
// then let the NEXT kernel of the stream be scheduled early (its CTAs sit in their own wait).  Because
// every kernel waits before it touches memory the chain is transitive and behaves exactly like plain
// stream order; what is saved is the launch latency between the 3-4 short kernels of one backward.
// Both instructions are no-ops when the kernel was launched without the attribute.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_wait(); pdl_release(); }

template <typename M> __device__ __forceinline__ M shfl_xor(M v, int m) {
    return __shfl_xor_sync(0xffffffffu, v, m);
}

}  // namespace dcnv3

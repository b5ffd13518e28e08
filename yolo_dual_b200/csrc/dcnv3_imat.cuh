// dcnv3_b200 — "interpolation-matrix" kernels (16-bit storage, group_channels = 16, 3x3 s1 d1).
//
// The vector kernels (dcnv3_kernels.cuh) spend one lane-instruction stream per 16-byte channel
// vector and point: the forward is instruction-issue bound and the backward is bound by the
// SM -> L2 reduction port.  This family removes the per-channel work from the
// SIMT lanes altogether:
//
//   * a CTA owns an 8x8 tile of output pixels of one image and 4 groups (64 channels); the
//     16x16-cell input window the tile can reach with |offset| < 3 px is staged ONCE in shared
//     memory (cp.async, zero-filled outside the image: that IS the reference's per-corner
//     validity, dcnv3_im2col_cuda.cuh:57-75);
//   * one warp per group; one LANE per (pixel, group).  The lane turns its 9 sampling points into
//     a row of a sparse interpolation matrix  Wm[pixel][cell] += w_corner * mask  (4 scalar
//     shared-memory updates per point, private to the lane: no atomics);
//   * out[pixel][ch] = sum_cell Wm[pixel][cell] * X[cell][ch] is then a dense [16 x 144] x
//     [144 x 16] product per 4x4 sub-tile on the tensor cores (mma.sync m16n8k8, TF32 weights,
//     the 16-bit activations are exact in TF32, fp32 accumulation).
//
// Location arithmetic is the shared locate() (bit-exact integer contract).  A point whose corners
// leave the sub-tile's 12x12 sub-window (|offset*scale| >= 3 px) takes a per-lane slow path that
// gathers from global memory exactly like the generic kernel, so any offset is handled.
//
// Reference semantics: dcnv3_im2col_gpu_kernel :216-275 + dcnv3_im2col_bilinear :32-80.
#pragma once

#include <type_traits>

#include "dcnv3_kernels.cuh"

// The staged windows of the default 16-bit kernels are ONE TMA box load each (cp.async.bulk.tensor.4d over the input as a
// (C, W, H, N) tensor, out-of-map cells zero-filled by the hardware).  A/B on one box, round 2: backward P3 168.5 ->
// 162.4 us, step 0.4093 -> 0.3978 ms against the per-thread cp.async fill (-DDCNV3_NO_TMA keeps that path).
#ifndef DCNV3_NO_TMA
#define DCNV3_WIN_TMA 1
#define DCNV3_FWD_TMA 1
#include <cuda.h>
#endif

namespace dcnv3 {
namespace imat {

constexpr int kTile = 8;                 // output pixels per tile edge
constexpr int kWin = 16;                 // window cells per edge (tile + 2*4)
constexpr int kSub = 12;                 // sub-window cells per edge (4x4 sub-tile + 2*4)
constexpr int kCells = kSub * kSub;      // 144 = K of the product
constexpr int kKSteps = kCells / 16;     // 9
constexpr int kRow = 152;                // words per Wm row (144 + pad; 152 % 32 == 24: LDS.64 conflict-free)
constexpr int kWarps = 4;                // groups per CTA
constexpr int kWinBytes = kWin * kWin * 128;       // 64 channels x 2 B per cell
constexpr int kWmWords = 32 * kRow;                // one Wm buffer: 32 pixels
constexpr int kSmemFwd = kWinBytes + kWarps * kWmWords * 4;
// backward, pass B / mma #2: the interpolation matrix in the storage dtype inside the same per-warp buffer
constexpr int kW16Stride = 400;                    // bytes per pixel row: 12 x 16 elements + 16 B skew
constexpr int kW16Bytes = 32 * kW16Stride;         // 12800
constexpr int kGoStride = 48;                      // grad_output rows behind it: 32 B + 16 B skew
static_assert(kW16Bytes + 32 * kGoStride <= kWmWords * 4, "Wm16 + grad_output must fit the Wm buffer");
static_assert(kW16Bytes % 512 == 0, "zero fill: whole 512-byte warp stores");

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
#ifndef DCNV3_NO_TMA
// thread-0 side of a TMA box load into shared memory: mbarrier (1 arrival + `bytes` of transaction), then the copy
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *tm, uint32_t bar_s, uint32_t bytes, int c0, int c1,
                                            int c2, int c3) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar_s), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// phase 0 of the mbarrier completes when all bytes have landed (bounded: a descriptor the hardware rejects must not hang the GPU)
__device__ __forceinline__ void tma_wait(uint32_t bar_s) {
    uint32_t done = 0u;
    for (int spin = 0; !done && spin < (1 << 22); ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar_s) : "memory");
}
#endif
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, int bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
// D += A(16x8, tf32) * B(8x8, tf32), fp32 accumulate
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// D += A(16x16) * B(16x8) in the storage dtype (bf16 / fp16 operands are used as they are), fp32 accumulate
template <typename T>
__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                          uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma_16816<__nv_bfloat16>(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2,
                                                         uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma_16816<__half>(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                  uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// D = A * B (zero accumulator input: no registers to clear per product)
template <typename T>
__device__ __forceinline__ void mma_16816_z(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                            uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma_16816_z<__nv_bfloat16>(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2,
                                                           uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%10, %10, %10, %10};"
                 : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f));
}
template <>
__device__ __forceinline__ void mma_16816_z<__half>(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                    uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%10, %10, %10, %10};"
                 : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f));
}

// two packed 16-bit storage values -> fp32 bit patterns (exact; both are valid TF32 operands)
template <typename T> __device__ __forceinline__ void unpack2(uint32_t w, uint32_t &lo, uint32_t &hi);
template <> __device__ __forceinline__ void unpack2<__nv_bfloat16>(uint32_t w, uint32_t &lo, uint32_t &hi) {
    lo = w << 16; hi = w & 0xffff0000u;
}
template <> __device__ __forceinline__ void unpack2<__half>(uint32_t w, uint32_t &lo, uint32_t &hi) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&w));
    lo = __float_as_uint(f.x); hi = __float_as_uint(f.y);
}
template <typename T> __device__ __forceinline__ float2 unpack2f(uint32_t w) {
    uint32_t lo, hi;
    unpack2<T>(w, lo, hi);
    return make_float2(__uint_as_float(lo), __uint_as_float(hi));
}
template <typename T> __device__ __forceinline__ float half_to_float(unsigned short h);
template <> __device__ __forceinline__ float half_to_float<__nv_bfloat16>(unsigned short h) {
    return __uint_as_float((uint32_t)h << 16);
}
template <> __device__ __forceinline__ float half_to_float<__half>(unsigned short h) {
    return __half2float(__ushort_as_half(h));
}
// bit pattern of a positive value in the storage dtype (rounded down: a threshold)
template <typename T> __device__ __forceinline__ uint32_t storage_bits(float v);
template <> __device__ __forceinline__ uint32_t storage_bits<__nv_bfloat16>(float v) { return __float_as_uint(v) >> 16; }
template <> __device__ __forceinline__ uint32_t storage_bits<__half>(float v) {
    return (uint32_t)__half_as_ushort(__float2half_rz(fminf(v, 60000.f)));
}
template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
}

// packed add of two storage-dtype pairs, round to nearest (HADD2 / HADD2.BF16)
template <typename T> __device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b);
template <> __device__ __forceinline__ uint32_t add2<__nv_bfloat16>(uint32_t a, uint32_t b) {
    const __nv_bfloat162 r = __hadd2(*reinterpret_cast<const __nv_bfloat162 *>(&a), *reinterpret_cast<const __nv_bfloat162 *>(&b));
    return *reinterpret_cast<const uint32_t *>(&r);
}
template <> __device__ __forceinline__ uint32_t add2<__half>(uint32_t a, uint32_t b) {
    const __half2 r = __hadd2(*reinterpret_cast<const __half2 *>(&a), *reinterpret_cast<const __half2 *>(&b));
    return *reinterpret_cast<const uint32_t *>(&r);
}

// ---------------------------------------------------------------------------------------------------------
// mixed-precision FMA (PTX ISA 8.6, sm_100+): d = a * b + c with 16-bit a, b and fp32 c, d.  The product of two
// 16-bit values is exact in fp32, so a dot of storage values is an fp32-accumulated exact dot.
// ---------------------------------------------------------------------------------------------------------
template <typename T> struct Mix;
template <> struct Mix<__nv_bfloat16> {
    // s0 += x.lo * g.lo ; s1 += x.hi * g.hi
    static __device__ __forceinline__ void dot2(float &s0, float &s1, uint32_t x, uint32_t g) {
        asm("{\n\t.reg .b16 xl, xh, gl, gh;\n\tmov.b32 {xl, xh}, %2;\n\tmov.b32 {gl, gh}, %3;\n\t"
            "fma.rn.f32.bf16 %0, xl, gl, %0;\n\tfma.rn.f32.bf16 %1, xh, gh, %1;\n\t}"
            : "+f"(s0), "+f"(s1) : "r"(x), "r"(g));
    }
    // a0 += x.lo * w ; a1 += x.hi * w   (w: 16-bit weight in the low half of a register)
    static __device__ __forceinline__ void axpy2(float &a0, float &a1, uint32_t x, uint32_t w) {
        asm("{\n\t.reg .b16 xl, xh, wl, wh;\n\tmov.b32 {xl, xh}, %2;\n\tmov.b32 {wl, wh}, %3;\n\t"
            "fma.rn.f32.bf16 %0, xl, wl, %0;\n\tfma.rn.f32.bf16 %1, xh, wl, %1;\n\t}"
            : "+f"(a0), "+f"(a1) : "r"(x), "r"(w));
    }
    static __device__ __forceinline__ uint32_t weight(float w) { return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(w)); }
};
template <> struct Mix<__half> {
    static __device__ __forceinline__ void dot2(float &s0, float &s1, uint32_t x, uint32_t g) {
        asm("{\n\t.reg .b16 xl, xh, gl, gh;\n\tmov.b32 {xl, xh}, %2;\n\tmov.b32 {gl, gh}, %3;\n\t"
            "fma.rn.f32.f16 %0, xl, gl, %0;\n\tfma.rn.f32.f16 %1, xh, gh, %1;\n\t}"
            : "+f"(s0), "+f"(s1) : "r"(x), "r"(g));
    }
    static __device__ __forceinline__ void axpy2(float &a0, float &a1, uint32_t x, uint32_t w) {
        asm("{\n\t.reg .b16 xl, xh, wl, wh;\n\tmov.b32 {xl, xh}, %2;\n\tmov.b32 {wl, wh}, %3;\n\t"
            "fma.rn.f32.f16 %0, xl, wl, %0;\n\tfma.rn.f32.f16 %1, xh, wl, %1;\n\t}"
            : "+f"(a0), "+f"(a1) : "r"(x), "r"(w));
    }
    static __device__ __forceinline__ uint32_t weight(float w) { return (uint32_t)__half_as_ushort(__float2half_rn(w)); }
};

// 16-byte chunk swizzle of the window: chunk' = chunk ^ key(cell); 8 consecutive cells of a
// sub-window row (also across its 12-cell wrap into the next row) get 8 distinct keys.
__device__ __forceinline__ int win_key(int cell) { return ((cell & 15) + 4 * (cell >> 4)) & 7; }

struct TileCoord { int n, ty, tx, gq; };
__device__ __forceinline__ TileCoord decode_tile(unsigned b, int tiles_x, int tiles_y, int GQ) {
    TileCoord c;
    c.gq = (int)(b % (unsigned)GQ); b /= (unsigned)GQ;
    c.tx = (int)(b % (unsigned)tiles_x); b /= (unsigned)tiles_x;
    c.ty = (int)(b % (unsigned)tiles_y);
    c.n = (int)(b / (unsigned)tiles_y);
    return c;
}
// Longest-first order for the backward: CTAs are dealt to SMs in blockIdx order and one CTA runs ~20 us, so a
// map whose extent is not a multiple of 8 (P5: 20 x 20 -> 4 whole, 4 half and 1 quarter tile per image)
// should start its whole tiles first and leave the cheap partial ones to fill the tail.  Classes: whole
// tiles of all images, then right-edge column, bottom-edge row, corner.
__device__ __forceinline__ TileCoord decode_tile_lpt(unsigned b, int Ho, int Wo, int N, int GQ) {
    const unsigned fy = (unsigned)Ho / kTile, fx = (unsigned)Wo / kTile;
    const unsigned ry = (Ho % kTile) ? 1u : 0u, rx = (Wo % kTile) ? 1u : 0u;
    const unsigned nA = fy * fx * (unsigned)N, nB = fy * rx * (unsigned)N, nC = fx * ry * (unsigned)N;
    TileCoord c;
    c.gq = (int)(b % (unsigned)GQ); b /= (unsigned)GQ;
    unsigned cls, pc;  // class and its tiles per image (non-zero in the class b falls into)
    if (b < nA) { cls = 0; pc = fy * fx; }
    else if ((b -= nA) < nB) { cls = 1; pc = fy; }
    else if ((b -= nB) < nC) { cls = 2; pc = fx; }
    else { b -= nC; cls = 3; pc = 1; }
    const unsigned t = b % pc;
    c.n = (int)(b / pc);
    if (cls == 0) { c.ty = (int)(t / fx); c.tx = (int)(t % fx); }
    else if (cls == 1) { c.ty = (int)t; c.tx = (int)fx; }
    else if (cls == 2) { c.ty = (int)fy; c.tx = (int)t; }
    else { c.ty = (int)fy; c.tx = (int)fx; }
    return c;
}

// Stage the 16x16-cell x 64-channel input window of (tile, group quad); zero outside the image.
// 128 threads: thread = (column, 16-byte chunk), one window row per iteration.
template <typename T>
__device__ __forceinline__ void fill_window(unsigned char *win, const T *in, const T *img, const Geo &q,
                                            int wy0, int wx0, int tid) {
    const uint32_t ws = smem_u32(win);
    const int ch = tid & 7, col = tid >> 3;
    const int ix = wx0 + col;
    const bool col_ok = (unsigned)ix < (unsigned)q.W;
    const long long row_stride = (long long)q.W * q.C;
    const T *p = img + ((long long)wy0 * q.W + ix) * q.C + ch * 8;
#pragma unroll
    for (int i = 0; i < kWin; ++i) {
        const bool ok = col_ok && (unsigned)(wy0 + i) < (unsigned)q.H;
        const T *src = ok ? p + i * row_stride : in;
        cp_async16(ws + (i * kWin + col) * 128 + ((ch ^ ((col + 4 * i) & 7)) << 4), src, ok ? 16 : 0);
    }
}

// Pixel of the build phase: lane -> (sub-tile column sx, position in the 4x4 sub-tile).
struct PixCoord { int oy, ox; bool valid; };
__device__ __forceinline__ PixCoord pix_of(int px, int ty, int tx, int pass, const Geo &q) {
    const int m = px & 15;
    PixCoord c;
    c.oy = ty * kTile + pass * 4 + (m >> 2);
    c.ox = tx * kTile + (px >> 4) * 4 + (m & 3);
    c.valid = c.oy < q.Ho && c.ox < q.Wo;
    return c;
}

// Byte offset (inside the window) of the ldmatrix row this lane addresses in k-step ks, for
// sub-tile column 0: matrix jm = lane >> 3 covers cells 16ks + 8(jm >> 1) + (lane & 7) and
// 16-byte chunk 2*group + (jm & 1).  Sub-tile column s: (off ^ (s << 6)) + s * 512.
__device__ __forceinline__ int b_row_offset(int ks, int lane, int group_in_quad) {
    const int jm = lane >> 3, jr = lane & 7;
    const int kk = 16 * ks + 8 * (jm >> 1) + jr;
    const int rr = kk / kSub, cc = kk - rr * kSub;
    const int chunk = 2 * group_in_quad + (jm & 1);
    return (rr * kWin + cc) * 128 + ((chunk ^ ((cc + 4 * rr) & 7)) << 4);
}

// One out-of-window sampling point of the forward: gather from global memory (generic semantics).
template <typename T>
__device__ __forceinline__ void slow_point_fwd(const T *img_g, const Geo &q, const Point<float> &t, float m,
                                               float (&v)[16]) {
    const float hm = t.hh * m, lm = t.lh * m;
    const float w[4] = {hm * t.hw, hm * t.lw, lm * t.hw, lm * t.lw};
    const bool ok[4] = {t.ok1, t.ok2, t.ok3, t.ok4};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (!ok[k]) continue;
        const T *p = img_g + ((size_t)(t.h_low + (k >> 1)) * q.W + (t.w_low + (k & 1))) * q.C;
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(p));
        const uint4 b = __ldg(reinterpret_cast<const uint4 *>(p) + 1);
        const uint32_t wd[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float2 f = unpack2f<T>(wd[e]);
            v[2 * e] = fmaf(w[k], f.x, v[2 * e]);
            v[2 * e + 1] = fmaf(w[k], f.y, v[2 * e + 1]);
        }
    }
}

// Offsets / mask of one pass (32 pixels x 9 points of group g), coalesced: word wi = i*32 + lane of
// the pass belongs to pixel wi / 9, point wi % 9 (pixel order = lane order of the build phase).
// `off32` / `msk16` point at the image's first pixel; indices inside one image fit 32 bits.
struct PassIO {
    int idx[9];  // element index (pixel * G*9 + g*9 + point) or -1 outside the map
};
__device__ __forceinline__ void pass_indices(PassIO &io, int lane, int ty, int tx, int pass, int g, const Geo &q) {
    const int GP = q.G * 9;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const int wi = i * 32 + lane;
        const int px = wi / 9, w = wi - px * 9;
        const PixCoord c = pix_of(px, ty, tx, pass, q);
        io.idx[i] = c.valid ? (c.oy * q.Wo + c.ox) * GP + g * 9 + w : -1;
    }
}

// Words 0..287 of the staging area live in Wm rows 0 and 1 (144 words each, the 8-word row pads are
// left alone: the backward keeps grad_output there), the 288 16-bit values in row 2.
__device__ __forceinline__ int stage_word(int wi) { return wi < kCells ? wi : wi + (kRow - kCells); }
constexpr int kStage16 = 2 * kRow * 2;  // index of the first 16-bit slot (row 2), in 16-bit units

template <typename T, bool LOGITS>
__device__ __forceinline__ void softmax9(float (&m)[9]) {
    if (!LOGITS) return;
    float mx = m[0];
#pragma unroll
    for (int p = 1; p < 9; ++p) mx = fmaxf(mx, m[p]);
    float sum = 0.f;
#pragma unroll
    for (int p = 0; p < 9; ++p) { m[p] = expf(m[p] - mx); sum += m[p]; }
    const float inv = 1.f / sum;
#pragma unroll
    for (int p = 0; p < 9; ++p) m[p] *= inv;
}

// ===========================================================================
// forward
// ===========================================================================
template <typename T, bool LOGITS>
__global__ void __launch_bounds__(32 * kWarps, 2)
fwd_imat_kernel(const T *__restrict__ in, const T *__restrict__ off, const T *__restrict__ mask,
                T *__restrict__ out, const Geo q, const int tiles_x, const int tiles_y, const int GQ) {
    extern __shared__ __align__(128) unsigned char smem[];
    pdl_enter();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TileCoord tc = decode_tile(blockIdx.x, tiles_x, tiles_y, GQ);
    const int g = tc.gq * kWarps + warp;
    const int wy0 = tc.ty * kTile + (q.half_h - q.ph) - 4;  // input row of window cell (0, 0)
    const int wx0 = tc.tx * kTile + (q.half_w - q.pw) - 4;
    const T *img = in + (size_t)tc.n * q.H * q.W * q.C + tc.gq * 64;
    const T *img_g = img + warp * 16;

    fill_window<T>(smem, in, img, q, wy0, wx0, tid);

    const size_t pix0 = (size_t)tc.n * q.Ho * q.Wo * q.G * 9;
    const uint32_t *off32 = reinterpret_cast<const uint32_t *>(off) + pix0;
    const unsigned short *msk16 = reinterpret_cast<const unsigned short *>(mask) + pix0;
    uint32_t roff[2][9];
    unsigned short rmsk[2][9];
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        PassIO io;
        pass_indices(io, lane, tc.ty, tc.tx, pass, g, q);
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            roff[pass][i] = io.idx[i] >= 0 ? __ldg(off32 + io.idx[i]) : 0u;
            rmsk[pass][i] = io.idx[i] >= 0 ? __ldg(msk16 + io.idx[i]) : (unsigned short)0;
        }
    }
    cp_async_commit_wait();
    __syncthreads();

    float *Wm = reinterpret_cast<float *>(smem + kWinBytes) + warp * kWmWords;
    const uint32_t win_s = smem_u32(smem);
    const int gID = lane >> 2, tq = lane & 3;
    int boff[kKSteps];
#pragma unroll
    for (int ks = 0; ks < kKSteps; ++ks) boff[ks] = b_row_offset(ks, lane, warp);

#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        // ---- this lane's 9 offset pairs and mask values (through the Wm buffer)
        uint32_t *st32 = reinterpret_cast<uint32_t *>(Wm);
        unsigned short *st16 = reinterpret_cast<unsigned short *>(Wm) + kStage16;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            st32[stage_word(i * 32 + lane)] = roff[pass][i];
            st16[i * 32 + lane] = rmsk[pass][i];
        }
        __syncwarp();
        uint32_t myoff[9];
        float mym[9];
#pragma unroll
        for (int p = 0; p < 9; ++p) {
            myoff[p] = st32[stage_word(lane * 9 + p)];
            mym[p] = half_to_float<T>(st16[lane * 9 + p]);
        }
        __syncwarp();
        {
            float4 *W4 = reinterpret_cast<float4 *>(Wm);
#pragma unroll 2
            for (int i = 0; i < kWmWords / 4 / 32; ++i) W4[i * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncwarp();

        // ---- build: lane = pixel; Wm[lane][cell] += corner weight * mask
        const int sx = lane >> 4;
        const PixCoord pc = pix_of(lane, tc.ty, tc.tx, pass, q);
        const int sy0 = wy0 + 4 * pass, sx0 = wx0 + 4 * sx;  // input coords of sub-window cell (0, 0)
        float p0h_, p0w_;
        window_origin<float>(q, pc.oy, pc.ox, p0h_, p0w_);
        softmax9<T, LOGITS>(mym);
        unsigned slow = 0u;
        float *Wrow = Wm + lane * kRow;
        if (pc.valid) {
#pragma unroll
            for (int p = 0; p < 9; ++p) {
                const float2 o = unpack2f<T>(myoff[p]);
                Point<float> t;
                locate<float>(q, p0h_, p0w_, p / 3, p % 3, o.x, o.y, t);
                if (t.inside) {
                    const unsigned u = (unsigned)(t.w_low - sx0), v = (unsigned)(t.h_low - sy0);
                    if (u <= (unsigned)(kSub - 2) && v <= (unsigned)(kSub - 2)) {
                        const float hm = t.hh * mym[p], lm = t.lh * mym[p];
                        float *c = Wrow + v * kSub + u;
                        c[0] += hm * t.hw;
                        c[1] += hm * t.lw;
                        c[kSub] += lm * t.hw;
                        c[kSub + 1] += lm * t.lw;
                    } else {
                        slow |= 1u << p;
                    }
                }
            }
        }
        __syncwarp();

        // ---- product: two 4x4 sub-tiles (Wm rows 0-15 / 16-31), 9 k-steps of 16 cells
        float acc[2][2][4];
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[s][nt][e] = 0.f;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const float *A0 = Wm + (16 * s + gID) * kRow + 2 * tq;
            const float *A1 = A0 + 8 * kRow;
            const uint32_t wbase = win_s + pass * (4 * kWin * 128) + s * 512;
#pragma unroll
            for (int ks = 0; ks < kKSteps; ++ks) {
                uint32_t r0, r1, r2, r3;
                ldmatrix_x4_trans(r0, r1, r2, r3, wbase + (boff[ks] ^ (s << 6)));
                const float2 a00 = *reinterpret_cast<const float2 *>(A0 + 16 * ks);
                const float2 a10 = *reinterpret_cast<const float2 *>(A1 + 16 * ks);
                const float2 a01 = *reinterpret_cast<const float2 *>(A0 + 16 * ks + 8);
                const float2 a11 = *reinterpret_cast<const float2 *>(A1 + 16 * ks + 8);
                uint32_t b0, b1;
                unpack2<T>(r0, b0, b1);
                mma_tf32(acc[s][0], __float_as_uint(a00.x), __float_as_uint(a10.x), __float_as_uint(a00.y), __float_as_uint(a10.y), b0, b1);
                unpack2<T>(r1, b0, b1);
                mma_tf32(acc[s][1], __float_as_uint(a00.x), __float_as_uint(a10.x), __float_as_uint(a00.y), __float_as_uint(a10.y), b0, b1);
                unpack2<T>(r2, b0, b1);
                mma_tf32(acc[s][0], __float_as_uint(a01.x), __float_as_uint(a11.x), __float_as_uint(a01.y), __float_as_uint(a11.y), b0, b1);
                unpack2<T>(r3, b0, b1);
                mma_tf32(acc[s][1], __float_as_uint(a01.x), __float_as_uint(a11.x), __float_as_uint(a01.y), __float_as_uint(a11.y), b0, b1);
            }
        }
        __syncwarp();

        // ---- epilogue: fragments -> per-pixel rows (stride 20 words), slow points, 32-byte store
        float *S = Wm;
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                *reinterpret_cast<float2 *>(S + (16 * s + gID) * 20 + nt * 8 + 2 * tq) = make_float2(acc[s][nt][0], acc[s][nt][1]);
                *reinterpret_cast<float2 *>(S + (16 * s + gID + 8) * 20 + nt * 8 + 2 * tq) = make_float2(acc[s][nt][2], acc[s][nt][3]);
            }
        __syncwarp();
        if (pc.valid) {
            float v[16];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float4 f = *reinterpret_cast<const float4 *>(S + lane * 20 + 4 * e);
                v[4 * e] = f.x; v[4 * e + 1] = f.y; v[4 * e + 2] = f.z; v[4 * e + 3] = f.w;
            }
            if (slow) {
#pragma unroll 1
                for (int p = 0; p < 9; ++p) {
                    if (!(slow & (1u << p))) continue;
                    uint32_t ow = myoff[0];
                    float mm = mym[0];
#pragma unroll
                    for (int e = 1; e < 9; ++e) if (e == p) { ow = myoff[e]; mm = mym[e]; }
                    const float2 o = unpack2f<T>(ow);
                    Point<float> t;
                    locate<float>(q, p0h_, p0w_, p / 3, p % 3, o.x, o.y, t);
                    slow_point_fwd<T>(img_g, q, t, mm, v);
                }
            }
            uint4 lo, hi;
            lo.x = pack2<T>(v[0], v[1]); lo.y = pack2<T>(v[2], v[3]); lo.z = pack2<T>(v[4], v[5]); lo.w = pack2<T>(v[6], v[7]);
            hi.x = pack2<T>(v[8], v[9]); hi.y = pack2<T>(v[10], v[11]); hi.z = pack2<T>(v[12], v[13]); hi.w = pack2<T>(v[14], v[15]);
            uint4 *dst = reinterpret_cast<uint4 *>(out + (((size_t)tc.n * q.Ho + pc.oy) * q.Wo + pc.ox) * q.C + g * 16);
            dst[0] = lo;
            dst[1] = hi;
        }
        __syncwarp();
    }
}

// ===========================================================================
// backward (16-bit storage, fp32 accumulation of grad_input into `gacc` [N,H,W,C], pre-zeroed)
//
// Per pass (32 pixels of one group = two 4x4 sub-tiles), all in the warp's private Wm buffer:
//   1. D[pixel][cell] = sum_ch go[pixel][ch] * X[cell][ch]          tensor cores  (mma #1, m16n8k16 in the storage dtype)
//   2. lane = pixel: for its 9 points read the four corner dots d_k from D ->
//        grad_mask   = sum_k w_k d_k                                   (cuh:144)
//        grad_offset = scale*m*(hh(d2-d1)+lh(d4-d3), hw(d3-d1)+lw(d4-d2))   (cuh:114-139,145-146)
//      staged and written coalesced;
//   3. Wm[pixel][cell] += w_k * m  (the forward's interpolation matrix)
//   4. GW[cell][ch] += sum_pixel Wm[pixel][cell] * go[pixel][ch]      tensor cores  (mma #2),
//      accumulated in registers over the four sub-tiles of the tile (16 window rows x 16 ch);
// then the CTA's GW window (256 cells x 64 channels fp32) goes to `gacc` as 256-byte-contiguous
// vector reductions: ~4 reduction bytes per grad_input byte instead of the vector kernel's 36.
// grad_output of a pixel (32 bytes) lives in the 8-word pad of its Wm row.
// ===========================================================================
// ---------------------------------------------------------------------------------------------
// Family selector.  The imat backward is ~1.7x faster than the vector kernel while the sampling
// points stay inside the staged window, and slower once more than ~1 in 7 leave it (each one is a
// divergent trip through global memory).  This kernel looks at 2 K offset pairs spread over the
// tensor and writes which family runs; both kernels are launched and the other one returns at once.
// A point is out of window when |(i - 1 + off) * scale| >= 4 (x: i = p / 3; y: j = p % 3).
// ---------------------------------------------------------------------------------------------
constexpr int kSelVec = 1, kSelImat = 3;
// 2 K samples: block 0 is ONE SM, and its load/store unit takes the scattered 4-byte loads one lane per
// cycle, so 16 K samples cost ~10 us of latency in front of a 3-10 us zero fill; 2 K decide the same
// question (6 % +- 0.5 %).
constexpr int kSelThreads = 1024, kSelPerThread = 2, kSelMaxSlowPct = 6;
// Zero fill of the fp32 workspace (replaces cudaMemsetAsync) with the selector riding along in block
// 0: independent samples per thread, one round trip to memory, no cross-block traffic.
template <typename T>
__global__ void __launch_bounds__(kSelThreads)
zero_select_kernel(uint4 *__restrict__ ws, const size_t n16, const T *__restrict__ off,
                   const unsigned long long n_points, const float scale, int *__restrict__ sel) {
    pdl_enter();
    if (blockIdx.x == 0 && sel != nullptr) {
        __shared__ int cnt;
        if (threadIdx.x == 0) cnt = 0;
        __syncthreads();
        // 32-bit sample ids: the stride is capped so that id stays below 2^32 (the samples then
        // cover the first 2^32 points, far beyond any real tensor)
        const unsigned long long capped = n_points < 0xffff0000ull ? n_points : 0xffff0000ull;
        const unsigned np = (unsigned)capped, stride = np / (kSelThreads * kSelPerThread) + 1;
        float2 o[kSelPerThread];
        bool live[kSelPerThread];
#pragma unroll
        for (int k = 0; k < kSelPerThread; ++k) {
            const unsigned id = (unsigned)(k * kSelThreads + threadIdx.x) * stride;
            live[k] = id < np;
            o[k] = live[k] ? load_offset_pair(off + 2 * (size_t)id) : make_float2(0.f, 0.f);
        }
        int v = 0;
#pragma unroll
        for (int k = 0; k < kSelPerThread; ++k) {
            const unsigned id = (unsigned)(k * kSelThreads + threadIdx.x) * stride;
            const int p = (int)(id % 9u);
            const float dx = ((float)(p / 3 - 1) + o[k].x) * scale, dy = ((float)(p % 3 - 1) + o[k].y) * scale;
            if (live[k]) v += ((fabsf(dx) < 4.f && fabsf(dy) < 4.f) ? 0 : (1 << 16)) | 1;
        }
        v = __reduce_add_sync(0xffffffffu, v);
        if ((threadIdx.x & 31) == 0) atomicAdd(&cnt, v);
        __syncthreads();
        if (threadIdx.x == 0) {
            const int s = cnt >> 16, n = cnt & 0xffff;
            sel[0] = (s * 100 > kSelMaxSlowPct * n) ? kSelVec : kSelImat;
        }
    }
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    const size_t step = (size_t)gridDim.x * kSelThreads;
    size_t i = (size_t)blockIdx.x * kSelThreads + threadIdx.x;
    for (; i + 3 * step < n16; i += 4 * step) {  // four independent 16-byte stores in flight per thread
        ws[i] = z; ws[i + step] = z; ws[i + 2 * step] = z; ws[i + 3 * step] = z;
    }
    for (; i < n16; i += step) ws[i] = z;
}

constexpr int kSmemBwd = kSmemFwd + 16;  // + one 16-byte zero word (operand of masked mma #2 lanes)

// The point loops of the backward are calls into these two non-inlined functions: the kernel is
// otherwise ~190 KB of straight-line SASS and stalls on instruction fetch (ncu: no_instruction 2.0
// per issue with 8 warps per SM).  dilation is 1 in this family, so locate() only needs these.
struct PtGeo { int H, W; float scale; };
__device__ __forceinline__ Geo geo_of(const PtGeo &g) {
    Geo q{};
    q.H = g.H; q.W = g.W; q.scale = g.scale; q.dh = 1; q.dw = 1;
    return q;
}
struct PtResA { uint32_t off; float m; uint32_t kq; uint32_t frac; };  // kq: cell | 256 when in-window; frac: (lh, lw) as 16-bit fixed point

// Location of one point with the arithmetic of locate() (same operations in the same order, so the
// same h_low / w_low / fractions bit for bit) but without the per-corner validity flags: inside the
// staged window the zero fill plays that role.  dilation is 1: (float)(i * dw) == fi.
struct LeanPoint { int h_low, w_low; float lh, lw; bool inside; };
__device__ __forceinline__ LeanPoint locate_lean(const PtGeo &g, float p0h_, float p0w_, float fi, float fj,
                                                 float off_w, float off_h) {
    LeanPoint t;
    const float loc_w = add_rn(p0w_, mul_rn(add_rn(fi, off_w), g.scale));
    const float loc_h = add_rn(p0h_, mul_rn(add_rn(fj, off_h), g.scale));
    t.inside = loc_h > -1.f && loc_w > -1.f && loc_h < (float)g.H && loc_w < (float)g.W;
    t.h_low = floor_to_int(loc_h);
    t.w_low = floor_to_int(loc_w);
    t.lh = sub_rn(loc_h, (float)t.h_low);
    t.lw = sub_rn(loc_w, (float)t.w_low);
    return t;
}

// Out-of-window point of pass A: corner dots from global memory; its grad_input contributions go out
// right here as 64-byte vector reductions (the vector kernel's path).  Rare: |offset*scale| >= 3 px.
template <typename T>
__device__ __noinline__ PtResA bwd_point_slow(const PtGeo pg, const float p0h_, const float p0w_, const int p,
                                              const uint32_t offw, const float m, const float *Drow,
                                              const T *img_g, float *gacc_g, const int C) {
    PtResA r{0u, 0.f, 0u, 0u};
    const Geo q = geo_of(pg);
    const int i = p / 3, j = p - 3 * i;
    const float2 o = unpack2f<T>(offw);
    Point<float> t;
    locate<float>(q, p0h_, p0w_, i, j, o.x, o.y, t);
    float go[16];
    const uint4 ga = *reinterpret_cast<const uint4 *>(Drow + kCells);
    const uint4 gb = *reinterpret_cast<const uint4 *>(Drow + kCells + 4);
    const uint32_t gwd[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float2 f = unpack2f<T>(gwd[c]);
        go[2 * c] = f.x; go[2 * c + 1] = f.y;
    }
    const float w[4] = {t.hh * t.hw, t.hh * t.lw, t.lh * t.hw, t.lh * t.lw};
    const bool ok[4] = {t.ok1, t.ok2, t.ok3, t.ok4};
    float d[4];
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
        d[k] = 0.f;
        if (!ok[k]) continue;
        const size_t e = ((size_t)(t.h_low + (k >> 1)) * q.W + (t.w_low + (k & 1))) * C;
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(img_g + e));
        const uint4 b = __ldg(reinterpret_cast<const uint4 *>(img_g + e) + 1);
        const uint32_t wd[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float2 f = unpack2f<T>(wd[c]);
            acc = fmaf(go[2 * c], f.x, acc);
            acc = fmaf(go[2 * c + 1], f.y, acc);
        }
        d[k] = acc;
        const float wm = w[k] * m;
        float *dst = gacc_g + e;
#pragma unroll
        for (int c = 0; c < 4; ++c)
            red_add_v4_f32(dst + 4 * c, wm * go[4 * c], wm * go[4 * c + 1], wm * go[4 * c + 2], wm * go[4 * c + 3]);
    }
    const float s_m = w[0] * d[0] + w[1] * d[1] + w[2] * d[2] + w[3] * d[3];
    const float s_w = t.hh * (d[1] - d[0]) + t.lh * (d[3] - d[2]);
    const float s_h = t.hw * (d[2] - d[0]) + t.lw * (d[3] - d[1]);
    const float sm = q.scale * m;
    r.off = pack2<T>(sm * s_w, sm * s_h);
    r.m = s_m;
    return r;
}

// pass A of one kernel column (three points p = 3i + j, j = 0..2): corner dots from D (or the slow
// path), grad_offset / grad_mask terms, and the cell + fractions pass B reuses.  Three points per call
// so that their location chains and the twelve D look-ups overlap (one point at a time was a serial
// chain of locate -> 4 shared loads -> math per point: 20 % of the kernel's stall samples); the
// look-ups are unconditional at a clamped cell so that no branch separates them.
struct PtRes3 { uint32_t off[3]; float m[3]; uint32_t frac[3]; uint32_t kq; };  // kq: 3 x (Wm16 element | 256 when in-window), 10 bits each
template <typename T>
__device__ __noinline__ PtRes3 bwd_points3_a(const PtGeo pg, const float p0h_, const float p0w_, const int i,
                                             const uint32_t offw0, const uint32_t offw1, const uint32_t offw2,
                                             const float m0, const float m1, const float m2, const float *Drow,
                                             const int sy0, const int sx0, const int cshift, const T *img_g,
                                             float *gacc_g, const int C) {
    const uint32_t offw[3] = {offw0, offw1, offw2};
    const float m[3] = {m0, m1, m2};
    const float fi = (float)i;
    LeanPoint t[3];
    unsigned k[3], e16[3];
    bool inwin[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float2 o = unpack2f<T>(offw[j]);
        t[j] = locate_lean(pg, p0h_, p0w_, fi, (float)j, o.x, o.y);
        const unsigned u = (unsigned)(t[j].w_low - sx0), v = (unsigned)(t[j].h_low - sy0);
        inwin[j] = t[j].inside && u <= (unsigned)(kSub - 2) && v <= (unsigned)(kSub - 2);
        k[j] = inwin[j] ? v * kSub + u : 0u;
        e16[j] = v * kWin + u + cshift;  // element of the pixel's Wm16 row: sub-window row x window column
    }
    float d[3][4];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float *c = Drow + k[j];
        d[j][0] = c[0]; d[j][1] = c[1]; d[j][2] = c[kSub]; d[j][3] = c[kSub + 1];
    }
    PtRes3 r;
    r.kq = 0u;
    unsigned slow = 0u;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        slow |= (t[j].inside && !inwin[j] ? 1u : 0u) << j;
        const float lh = t[j].lh, lw = t[j].lw;
        const float hh = sub_rn(1.f, lh), hw = sub_rn(1.f, lw);
        const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
        const float s_m = w1 * d[j][0] + w2 * d[j][1] + w3 * d[j][2] + w4 * d[j][3];
        const float s_w = hh * (d[j][1] - d[j][0]) + lh * (d[j][3] - d[j][2]);
        const float s_h = hw * (d[j][2] - d[j][0]) + lw * (d[j][3] - d[j][1]);
        const float sm = pg.scale * m[j];
        r.off[j] = inwin[j] ? pack2<T>(sm * s_w, sm * s_h) : 0u;
        r.m[j] = inwin[j] ? s_m : 0.f;
        // fractions for pass B, truncated to 2^-16 (the weights there feed TF32 operands anyway)
        r.frac[j] = (__float2uint_rz(lh * 65536.f) << 16) | __float2uint_rz(lw * 65536.f);
        r.kq |= (inwin[j] ? (256u | e16[j]) : 0u) << (10 * j);
    }
    if (slow)
#pragma unroll 1
    for (int j = 0; j < 3; ++j) {
        if ((slow >> j) & 1u) {  // rare: |offset * scale| >= 3 px
            uint32_t ow = offw0; float mm = m0;
            if (j == 1) { ow = offw1; mm = m1; }
            if (j == 2) { ow = offw2; mm = m2; }
            const PtResA s = bwd_point_slow<T>(pg, p0h_, p0w_, 3 * i + j, ow, mm, Drow, img_g, gacc_g, C);
            if (j == 0) { r.off[0] = s.off; r.m[0] = s.m; }
            if (j == 1) { r.off[1] = s.off; r.m[1] = s.m; }
            if (j == 2) { r.off[2] = s.off; r.m[2] = s.m; }
        }
    }
    return r;
}

template <typename T, bool LOGITS>
__global__ void __launch_bounds__(32 * kWarps, 2)
bwd_imat_kernel(const T *__restrict__ in, const T *__restrict__ off, const T *__restrict__ mask,
                const T *__restrict__ gout, float *__restrict__ gacc, T *__restrict__ goff,
                T *__restrict__ gmask, const Geo q, const int tiles_x, const int tiles_y, const int GQ,
                const int *sel) {
    extern __shared__ __align__(128) unsigned char smem[];
    pdl_enter();
    if (sel != nullptr && __ldcg(sel) != kSelImat) return;  // select_kernel chose the vector family
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TileCoord tc = decode_tile_lpt(blockIdx.x, q.Ho, q.Wo, q.N, GQ);
    const int g = tc.gq * kWarps + warp;
    const int wy0 = tc.ty * kTile + (q.half_h - q.ph) - 4;
    const int wx0 = tc.tx * kTile + (q.half_w - q.pw) - 4;
    const size_t img_off = (size_t)tc.n * q.H * q.W * q.C + tc.gq * 64;
    const T *img = in + img_off;
    const T *img_g = img + warp * 16;
    float *gacc_g = gacc + img_off + warp * 16;
    const PtGeo pg{q.H, q.W, q.scale};

    fill_window<T>(smem, in, img, q, wy0, wx0, tid);
    asm volatile("cp.async.commit_group;" ::: "memory");

    const size_t pix0 = (size_t)tc.n * q.Ho * q.Wo * q.G * 9;
    const uint32_t *off32 = reinterpret_cast<const uint32_t *>(off) + pix0;
    const unsigned short *msk16 = reinterpret_cast<const unsigned short *>(mask) + pix0;
    uint32_t *goff32 = reinterpret_cast<uint32_t *>(goff) + pix0;
    unsigned short *gmsk16 = reinterpret_cast<unsigned short *>(gmask) + pix0;

    float *Wm = reinterpret_cast<float *>(smem + kWinBytes) + warp * kWmWords;
    const uint32_t wm_s = smem_u32(Wm);
    const uint32_t win_s = smem_u32(smem);
    const int gID = lane >> 2, tq = lane & 3;
    const int jm = lane >> 3, jr = lane & 7;

    // grad_input window of this (tile, group): 16 rows x (16 cells x 16 channels) as mma accumulators
    constexpr bool kScaled = sizeof(T) == 2 && !std::is_same<T, __half>::value;  // bf16: see pass B
    int e_ref = -1;
    float gw[kWin][2][4];
#pragma unroll
    for (int r = 0; r < kWin; ++r)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) gw[r][nt][e] = 0.f;

    // element indices of this lane's 9 staging words in pass 0; pass 1 is 4 rows further down
    PassIO io;
    unsigned io1 = 0u;  // validity of the same words in pass 1
    const int pass_stride = 4 * q.Wo * q.G * 9;
    {
        const int GP = q.G * 9;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const int wi = i * 32 + lane;
            const int px = wi / 9, w = wi - px * 9;
            const PixCoord c = pix_of(px, tc.ty, tc.tx, 0, q);
            io.idx[i] = c.valid ? (c.oy * q.Wo + c.ox) * GP + g * 9 + w : -1;
            io1 |= (c.oy + 4 < q.Ho && c.ox < q.Wo ? 1u : 0u) << i;
        }
    }

    bool window_ready = false;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        if (tc.ty * kTile + 4 * pass >= q.Ho) break;  // the tile's lower half lies below the map (warp-uniform)
        const PixCoord pc = pix_of(lane, tc.ty, tc.tx, pass, q);
        const int n_sub = (tc.tx * kTile + 4 < q.Wo) ? 2 : 1;  // right sub-tile column outside the map?
        if (pass == 1) {
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                // (a word valid in pass 1 is valid in pass 0: same column, higher row)
                io.idx[i] = ((io1 >> i) & 1u) ? io.idx[i] + pass_stride : -1;
            }
        }
        // ---- loads of the pass: offsets / mask (coalesced, staged) and this pixel's grad_output
        uint32_t *st32 = reinterpret_cast<uint32_t *>(Wm);
        unsigned short *st16 = reinterpret_cast<unsigned short *>(Wm) + kStage16;
        {
            uint4 go_lo = make_uint4(0u, 0u, 0u, 0u), go_hi = go_lo;
            uint32_t ro[9];
            unsigned short rm[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                ro[i] = io.idx[i] >= 0 ? __ldg(off32 + io.idx[i]) : 0u;
                rm[i] = io.idx[i] >= 0 ? __ldg(msk16 + io.idx[i]) : (unsigned short)0;
            }
            if (pc.valid) {
                const uint4 *gp = reinterpret_cast<const uint4 *>(gout + (((size_t)tc.n * q.Ho + pc.oy) * q.Wo + pc.ox) * q.C + g * 16);
                go_lo = __ldg(gp);
                go_hi = __ldg(gp + 1);
                if (pass == 0 && pc.oy + 4 < q.Ho)  // next pass's grad_output row: 4 map rows further down
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(gp) + (size_t)4 * q.Wo * q.C * sizeof(T)));
            }
            if (pass == 0) {
#pragma unroll
                for (int i = 0; i < 9; ++i)
                    if ((io1 >> i) & 1u) {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(off32 + io.idx[i] + pass_stride));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(msk16 + io.idx[i] + pass_stride));
                    }
            }
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                st32[stage_word(i * 32 + lane)] = ro[i];
                st16[i * 32 + lane] = rm[i];
            }
            // grad_output slab of this pixel into the pad of its row (outside the staging words)
            *reinterpret_cast<uint4 *>(Wm + lane * kRow + kCells) = go_lo;
            *reinterpret_cast<uint4 *>(Wm + lane * kRow + kCells + 4) = go_hi;
        }
        __syncwarp();
        uint32_t myoff[9];
        float mym[9];
#pragma unroll
        for (int p = 0; p < 9; ++p) {
            myoff[p] = st32[stage_word(lane * 9 + p)];
            mym[p] = half_to_float<T>(st16[lane * 9 + p]);
        }
        softmax9<T, LOGITS>(mym);
        if (!window_ready) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            window_ready = true;
        } else {
            __syncwarp();
        }

        // ---- mma #1: D[pixel][cell] for both sub-tiles (rolled: 2 x 9 iterations)
#pragma unroll 1
        for (int s = 0; s < n_sub; ++s) {
            // both operands are exact in the storage dtype: one m16n8k16 per 8 cells, no conversions.
            // A = go [16 px x 16 ch]: (px 0-7, ch 0-7) (px 8-15, ch 0-7) (px 0-7, ch 8-15) (px 8-15, ch 8-15)
            uint32_t a0, a1, a2, a3;
            ldmatrix_x4(a0, a1, a2, a3, wm_s + ((16 * s + 8 * (jm & 1) + jr) * kRow + kCells) * 4 + (jm >> 1) * 16);
            const uint32_t wbase = win_s + pass * (4 * kWin * 128) + s * 512;
            float *D0 = Wm + (16 * s + gID) * kRow + 2 * tq;
            // B rows: cell 16ks + 8(jm>>1) + jr of the sub-window.  Three k-steps are exactly four
            // sub-window rows, so the offsets of ks = 3m + r are those of ks = r plus m*4 window rows
            // (and the swizzle key does not change).
            const int chunk = (2 * warp + (jm & 1)) ^ (4 * s);
            uint32_t bo[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const int kk = 16 * r + 8 * (jm >> 1) + jr;
                const int rr = kk / kSub, cc = kk - rr * kSub;
                bo[r] = wbase + (rr * kWin + cc) * 128 + ((chunk ^ ((cc + 4 * rr) & 7)) << 4);
            }
            // software pipeline over the three groups of three k-steps: the ldmatrix of group m + 1 are in
            // flight while group m multiplies and stores (every asm here is volatile, so program order is
            // issue order: without this each mma waited out its own ldmatrix)
            uint32_t fr[2][3][4];  // (cells 0-7, ch 0-7) (cells 0-7, ch 8-15) (cells 8-15, ch 0-7) (cells 8-15, ch 8-15)
#pragma unroll
            for (int r = 0; r < 3; ++r) ldmatrix_x4(fr[0][r][0], fr[0][r][1], fr[0][r][2], fr[0][r][3], bo[r]);
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                if (m < 2) {
#pragma unroll
                    for (int r = 0; r < 3; ++r)
                        ldmatrix_x4(fr[(m + 1) & 1][r][0], fr[(m + 1) & 1][r][1], fr[(m + 1) & 1][r][2], fr[(m + 1) & 1][r][3],
                                    bo[r] + (m + 1) * (4 * kWin * 128));
                }
                float d[3][2][4];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    mma_16816_z<T>(d[r][0], a0, a1, a2, a3, fr[m & 1][r][0], fr[m & 1][r][1]);
                    mma_16816_z<T>(d[r][1], a0, a1, a2, a3, fr[m & 1][r][2], fr[m & 1][r][3]);
                }
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    float *Dk = D0 + 48 * m + 16 * r;
                    *reinterpret_cast<float2 *>(Dk) = make_float2(d[r][0][0], d[r][0][1]);
                    *reinterpret_cast<float2 *>(Dk + 8 * kRow) = make_float2(d[r][0][2], d[r][0][3]);
                    *reinterpret_cast<float2 *>(Dk + 8) = make_float2(d[r][1][0], d[r][1][1]);
                    *reinterpret_cast<float2 *>(Dk + 8 * kRow + 8) = make_float2(d[r][1][2], d[r][1][3]);
                }
            }
        }
        __syncwarp();

        // ---- pass A: lane = pixel; grad_offset / grad_mask from the corner dots
        const int sy0 = wy0 + 4 * pass, sx0 = wx0 + 4 * (lane >> 4);
        float p0h_, p0w_;
        window_origin<float>(q, pc.oy, pc.ox, p0h_, p0w_);
        float *Wrow = Wm + lane * kRow;
        uint32_t res_off[9], frac[9];
        float res_m[9];
        uint32_t kq[3] = {0u, 0u, 0u};  // 9 x (cell | in-window bit), 10 bits each
#pragma unroll
        for (int p = 0; p < 9; ++p) {
            res_off[p] = 0u;
            res_m[p] = 0.f;
            frac[p] = 0u;
        }
        if (pc.valid) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const PtRes3 r = bwd_points3_a<T>(pg, p0h_, p0w_, i, myoff[3 * i], myoff[3 * i + 1], myoff[3 * i + 2],
                                                  mym[3 * i], mym[3 * i + 1], mym[3 * i + 2], Wrow, sy0, sx0, 4 * (lane >> 4), img_g, gacc_g, q.C);
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    res_off[3 * i + j] = r.off[j];
                    res_m[3 * i + j] = r.m[j];
                    frac[3 * i + j] = r.frac[j];
                }
                kq[i] = r.kq;
            }
        }
        if (LOGITS) {  // softmax Jacobian: dl_p = m_p (gm_p - sum_q m_q gm_q)
            float dot = 0.f;
#pragma unroll
            for (int p = 0; p < 9; ++p) dot = fmaf(mym[p], res_m[p], dot);
#pragma unroll
            for (int p = 0; p < 9; ++p) res_m[p] = mym[p] * (res_m[p] - dot);
        }
        __syncwarp();  // every lane is done with D
#pragma unroll
        for (int p = 0; p < 9; ++p) {
            st32[stage_word(lane * 9 + p)] = res_off[p];
            st16[lane * 9 + p] = (unsigned short)(pack2<T>(res_m[p], 0.f) & 0xffffu);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            if (io.idx[i] >= 0) {
                goff32[io.idx[i]] = st32[stage_word(i * 32 + lane)];
                gmsk16[io.idx[i]] = st16[i * 32 + lane];
            }
        }
        __syncwarp();

        // ---- pass B: the interpolation matrix in 16 bits.  Wm16[pixel][sub-window row][window
        // column] (12 x 16 elements = 384 B of a 400-byte row: the 16-byte skew makes the eight rows of
        // every ldmatrix hit distinct banks) replaces the fp32 [pixel][144] matrix: 12.5 KB to zero instead
        // of 18 KB, and mma #2 reads it with ONE ldmatrix.trans per window row (16 cells x 16 pixels)
        // instead of 8 scalar loads, with K = 16 pixels per m16n8k16.  The matrix is fp16 for both storage
        // dtypes (11-bit weights, what TF32 gave); accumulation stays fp32.
        // The window column is absolute (0..15), so sub-tile 0 never touches columns 12-15 and sub-tile
        // 1 never columns 0-3: they stay zero and no operand masking is needed.
        unsigned char *W16 = reinterpret_cast<unsigned char *>(Wm);
        {
            // this pixel's grad_output moves out of the D-layout pad (the zero fill runs over it)
            uint4 g0 = *reinterpret_cast<const uint4 *>(Wrow + kCells);
            uint4 g1 = *reinterpret_cast<const uint4 *>(Wrow + kCells + 4);
            if constexpr (kScaled) {
                // bf16 storage: mma #2 runs in fp16 (weights keep 11 bits, as TF32 did), so grad_output is
                // brought into fp16 range by a power of two per (tile, group): go' = go * 2^(127 - e_ref)
                // with e_ref the biased exponent of the largest |go| seen so far.  Exact for every value
                // within 2^-14 of that maximum (smaller ones lose low bits: an absolute error below
                // 2^-24 of the tile's largest gradient).  A later pass that would overflow fp16 rescales
                // the accumulators instead (online max, warp-uniform, rare).
                uint32_t w[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                uint32_t mx = 0u;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint32_t a = w[c] & 0x7fff7fffu;  // |bf16| compares like an integer
                    mx = max(mx, max(a >> 16, a & 0xffffu));
                }
                mx = __reduce_max_sync(0xffffffffu, mx);
                const int e = min(max((int)(mx >> 7), 1), 253);
                if (e_ref < 0) {
                    e_ref = e;
                } else if (e > e_ref + 14) {
                    const int d = e_ref - e;  // < -14
                    const float f = d >= -126 ? __uint_as_float((uint32_t)(127 + d) << 23) : 0.f;
#pragma unroll
                    for (int r = 0; r < kWin; ++r)
#pragma unroll
                        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                            for (int c = 0; c < 4; ++c) gw[r][nt][c] *= f;
                    e_ref = e;
                }
                const float sc = __uint_as_float((uint32_t)(254 - e_ref) << 23);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float2 f = unpack2f<T>(w[c]);
                    w[c] = pack2<__half>(f.x * sc, f.y * sc);
                }
                g0 = make_uint4(w[0], w[1], w[2], w[3]);
                g1 = make_uint4(w[4], w[5], w[6], w[7]);
            }
            __syncwarp();
            uint4 *Z = reinterpret_cast<uint4 *>(W16);
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int r = 0; r < kW16Bytes / 512; ++r) Z[r * 32 + lane] = z;
            uint4 *G = reinterpret_cast<uint4 *>(W16 + kW16Bytes + lane * kGoStride);
            G[0] = g0;
            G[1] = g1;
        }
        __syncwarp();
        {
            uint32_t *R32 = reinterpret_cast<uint32_t *>(W16 + lane * kW16Stride);
#pragma unroll
            for (int p = 0; p < 9; ++p) {
                const uint32_t e = kq[p / 3] >> (10 * (p % 3));
                if (e & 256u) {
                    const float lh = (float)(frac[p] >> 16) * (1.f / 65536.f), lw = (float)(frac[p] & 0xffffu) * (1.f / 65536.f);
                    const float hm = (1.f - lh) * mym[p], lm = lh * mym[p], hw = 1.f - lw;
                    // the two corners of a row are adjacent elements: one packed add when the left one is
                    // even, else the pair straddles two words (the funnel shift splits it; adding 0 is free;
                    // predicating the second add on the parity was measured slower)
                    const uint32_t top = pack2<__half>(hm * hw, hm * lw), bot = pack2<__half>(lm * hw, lm * lw);
                    const uint32_t sh = (e & 1u) << 4;
                    uint32_t *w = R32 + ((e & 255u) >> 1);
                    w[0] = add2<__half>(w[0], top << sh);
                    w[1] = add2<__half>(w[1], __funnelshift_l(top, 0u, sh));
                    w[kWin / 2] = add2<__half>(w[kWin / 2], bot << sh);
                    w[kWin / 2 + 1] = add2<__half>(w[kWin / 2 + 1], __funnelshift_l(bot, 0u, sh));
                }
            }
        }
        __syncwarp();

        // ---- mma #2: gw[window row][ch] += Wm16^T * go, per sub-tile; m-tile = one window row (16 cells),
        // K = the 16 pixels of the sub-tile, both operands in the storage dtype (m16n8k16)
        {
            const uint32_t w16_s = smem_u32(W16);
#pragma unroll 1
            for (int s = 0; s < n_sub; ++s) {
                // B = go [16 px x 16 ch]: matrix jm = (px 8(jm & 1).., ch 8(jm >> 1)..)
                uint32_t b00, b01, b10, b11;  // n-tile 0: (k 0-7, k 8-15); n-tile 1: (k 0-7, k 8-15)
                ldmatrix_x4_trans(b00, b01, b10, b11,
                                  w16_s + kW16Bytes + (16 * s + 8 * (jm & 1) + jr) * kGoStride + (jm >> 1) * 16);
                // A = Wm16^T [16 cells x 16 px]: matrix jm = (cells 8(jm & 1).., px 8(jm >> 1)..)
                const uint32_t abase = w16_s + (16 * s + 8 * (jm >> 1) + jr) * kW16Stride + (jm & 1) * 16;
#pragma unroll
                for (int rr = 0; rr < kSub; ++rr) {
                    uint32_t a0, a1, a2, a3;
                    ldmatrix_x4_trans(a0, a1, a2, a3, abase + rr * (kWin * 2));
                    mma_16816<__half>(gw[4 * pass + rr][0], a0, a1, a2, a3, b00, b01);
                    mma_16816<__half>(gw[4 * pass + rr][1], a0, a1, a2, a3, b10, b11);
                }
            }
        }
        __syncwarp();
    }

    // ---- flush, per warp (no CTA barrier)
#ifdef DCNV3_IMAT_FLUSH_DIRECT
    // A/B variant (-DDCNV3_IMAT_FLUSH_DIRECT), measured SLOWER: straight from the mma fragments — (c0, c1) of a
    // fragment are two consecutive channels of one cell, the four lanes of a quad cover one 32-byte sector, each
    // REDG.F32x2 instruction carries 8 sectors.  It saves the detour through shared memory (11 % of the kernel's
    // shared wavefronts, 8 % of its instructions) and loses more on the reduction path: twice the reduction
    // instructions at 32 instead of 64 contiguous bytes (2.46 vs 1.83 port clocks per sector,
    // profiles/r01_red_egress_microbench.md): P3 backward 251.4 vs 233.1 us on one box.
    {
        const float unscale = __uint_as_float((uint32_t)max(e_ref, 1) << 23);  // 2^(e_ref - 127), bf16 storage only
        const int ix0 = wx0 + gID, ix1 = ix0 + 8;
        const bool ok0 = (unsigned)ix0 < (unsigned)q.W, ok1 = (unsigned)ix1 < (unsigned)q.W;
        const long long row_stride = (long long)q.W * q.C;
        float *d0 = gacc_g + ((long long)wy0 * q.W + ix0) * q.C + 2 * tq;
#pragma unroll
        for (int r = 0; r < kWin; ++r) {
            const bool row_ok = (unsigned)(wy0 + r) < (unsigned)q.H;
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                float v0 = gw[r][nt][0], v1 = gw[r][nt][1], v2 = gw[r][nt][2], v3 = gw[r][nt][3];
                if constexpr (kScaled) { v0 *= unscale; v1 *= unscale; v2 *= unscale; v3 *= unscale; }
                const bool nz0 = ((__float_as_uint(v0) | __float_as_uint(v1)) << 1) != 0u;
                const bool nz1 = ((__float_as_uint(v2) | __float_as_uint(v3)) << 1) != 0u;
                red_add_v2_f32(row_ok && ok0 ? d0 + 8 * nt : gacc, v0, v1, row_ok && ok0 && nz0);
                red_add_v2_f32(row_ok && ok1 ? d0 + 8 * nt + 8 * q.C : gacc, v2, v3, row_ok && ok1 && nz1);
            }
            d0 += row_stride;
        }
    }
#else
    // through the warp's own Wm buffer as [256 cells][16 ch] fp32 (16-byte chunks swizzled by 2*((cell >> 1) & 1):
    // conflict-free both ways), then out as 64-byte-contiguous vector reductions; cells that received nothing are skipped
    {
        float *GW = Wm;
#pragma unroll
        for (int r = 0; r < kWin; ++r)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
                const int c0 = r * kWin + gID;  // and c0 + 8: same swizzle key
                const int sw = (((2 * nt + (tq >> 1)) ^ (((gID >> 1) & 1) << 1)) << 2) + 2 * (tq & 1);
                *reinterpret_cast<float2 *>(GW + c0 * 16 + sw) = make_float2(gw[r][nt][0], gw[r][nt][1]);
                *reinterpret_cast<float2 *>(GW + (c0 + 8) * 16 + sw) = make_float2(gw[r][nt][2], gw[r][nt][3]);
            }
        __syncwarp();
        const int j = lane & 3, c8 = lane >> 2;  // 16-byte chunk of the cell, cell within 8
        const int ix0 = wx0 + c8, ix1 = ix0 + 8;
        const bool ok0 = (unsigned)ix0 < (unsigned)q.W, ok1 = (unsigned)ix1 < (unsigned)q.W;
        const long long row_stride = (long long)q.W * q.C;
        float *dst0 = gacc_g + ((long long)wy0 * q.W + ix0) * q.C + 4 * j;
        const float *src = GW + c8 * 16 + ((j ^ (((c8 >> 1) & 1) << 1)) << 2);
        const float unscale = __uint_as_float((uint32_t)max(e_ref, 1) << 23);  // 2^(e_ref - 127), bf16 storage only
#pragma unroll 4
        for (int r = 0; r < kWin; ++r) {
            const bool row_ok = (unsigned)(wy0 + r) < (unsigned)q.H;
            float4 v0 = *reinterpret_cast<const float4 *>(src + r * (kWin * 16));
            float4 v1 = *reinterpret_cast<const float4 *>(src + r * (kWin * 16) + 8 * 16);
            if constexpr (kScaled) {
                v0.x *= unscale; v0.y *= unscale; v0.z *= unscale; v0.w *= unscale;
                v1.x *= unscale; v1.y *= unscale; v1.z *= unscale; v1.w *= unscale;
            }
            const bool nz0 = ((__float_as_uint(v0.x) | __float_as_uint(v0.y) | __float_as_uint(v0.z) | __float_as_uint(v0.w)) << 1) != 0u;
            const bool nz1 = ((__float_as_uint(v1.x) | __float_as_uint(v1.y) | __float_as_uint(v1.z) | __float_as_uint(v1.w)) << 1) != 0u;
            float *d = dst0 + r * row_stride;
            red_add_v4_f32(row_ok && ok0 ? d : gacc, v0.x, v0.y, v0.z, v0.w, row_ok && ok0 && nz0);
            red_add_v4_f32(row_ok && ok1 ? d + 8 * q.C : gacc, v1.x, v1.y, v1.z, v1.w, row_ok && ok1 && nz1);
        }
    }
#endif
}

// ===========================================================================
// forward from a staged window (16-bit storage, group_channels = 16, 3x3 s1 d1): SIMT gathers out of
// shared memory instead of L1.
//
// fwd_vec_kernel is co-limited by instruction issue (77 %) and by L1 wavefronts (74 %): every corner of every
// point is a 32-byte sector of its own cache line, ~1.9 sectors per L1 wavefront.  Here a CTA stages the
// 16x16-cell x 64-channel window of an (8x8 tile, 4 groups) once, UNswizzled: a cell is one 128-byte row and
// the 16-byte chunk of lane (g, h) inside it is always chunk 2g + h.  The eight lanes of a quarter-warp are the
// four groups x two channel halves of ONE pixel, so whatever cells the four groups sample, their eight chunks
// fall into eight different bank groups: every LDS.128 is conflict-free (4 wavefronts for 16 slabs instead of
// ~8.5).  The zero fill outside the map is the reference's per-corner validity, so the lanes carry no
// validity predicates, no predicated loads and 32-bit addresses: ~40 % fewer instructions per point.
// Points that leave the window (|offset * scale| >= ~4 px) gather from global memory as the vector kernel does.
// ===========================================================================
constexpr int kFwdTileThreads = 256;
constexpr int kFwin = 20, kFhalo = 6;            // window of the staged forward: 8x8 tile + 6 cells each side
constexpr int kFwinCells = kFwin * kFwin;
constexpr int kFzeroCells = kFwin + 2;            // zero cells behind the window: where a closed point's four corners read
constexpr int kFwinBytes = (kFwinCells + kFzeroCells) * 128;  // 54 016 B: three CTAs per SM
// staging area of the tile's offsets / masks (two TMA boxes, as in win::bwd_win_kernel): 64 pixels x 160 B / 80 B
constexpr int kFstOffPx = 160, kFstMaskPx = 80;
constexpr int kFstOffB = 64 * kFstOffPx, kFstMaskB = 64 * kFstMaskPx;
#ifndef DCNV3_FWD_TMA_STAGE
#define DCNV3_FWD_TMA_STAGE 1
#endif
#if defined(DCNV3_FWD_TMA) && DCNV3_FWD_TMA_STAGE
constexpr int kFwdSmemB = kFwinBytes + kFstOffB + kFstMaskB;   // 69 376 B: still three CTAs per SM
#else
constexpr int kFwdSmemB = kFwinBytes;
#endif
static_assert(kFwinBytes % 128 == 0 && kFstOffB % 128 == 0, "TMA destinations are 128-byte aligned");

template <typename T>
__device__ __forceinline__ void fill_window_plain(unsigned char *win, const T *in, const T *img, const Geo &q,
                                                  int wy0, int wx0, int tid) {
    // thread = (column slot, 16-byte chunk): slots 0..19 copy one column of the 20 window rows each (warps 5-7
    // have no column and go straight on); pointer and shared address advance by constants
    const int ch = tid & 7, col = tid >> 3;
    if (col >= kFwin) return;
    const int ix = wx0 + col;
    const bool col_ok = (unsigned)ix < (unsigned)q.W;
    const size_t step = (size_t)q.W * q.C;
    const T *p = img + ((long long)wy0 * q.W + ix) * q.C + ch * 8;
    uint32_t dst = smem_u32(win) + col * 128 + (ch << 4);
    int iy = wy0;
#pragma unroll 4
    for (int i = 0; i < kFwin; ++i) {
        const bool ok = col_ok && (unsigned)iy < (unsigned)q.H;
        cp_async16(dst, ok ? p : in, ok ? 16 : 0);
        p += step; dst += kFwin * 128; ++iy;
    }
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// acc += w * (8 storage values): mixed-precision FMAs (FHFMA) straight on the halves of the loaded registers, fp32
// accumulation; `w` is the corner weight rounded to the storage dtype (Mix<T>::weight)
template <typename T>
__device__ __forceinline__ void fma8(float2 (&acc)[4], const uint4 &c, uint32_t w) {
    Mix<T>::axpy2(acc[0].x, acc[0].y, c.x, w); Mix<T>::axpy2(acc[1].x, acc[1].y, c.y, w);
    Mix<T>::axpy2(acc[2].x, acc[2].y, c.z, w); Mix<T>::axpy2(acc[3].x, acc[3].y, c.w, w);
}
// fp32-weight variant (unpack + FFMA2) for the rare gathers from global memory
template <typename T>
__device__ __forceinline__ void fma8f(float2 (&acc)[4], const uint4 &c, float w) {
    const uint32_t wd[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] = __ffma2_rn(unpack2f<T>(wd[k]), make_float2(w, w), acc[k]);
}

// one corner of a window-resident point: both 16-byte chunks of the group's slab (own chunk first)
template <typename T>
__device__ __forceinline__ void corner16(float2 (&acc)[8], uint32_t a0, uint32_t a1, uint32_t w) {
    const uint4 x0 = lds128(a0), x1 = lds128(a1);
    Mix<T>::axpy2(acc[0].x, acc[0].y, x0.x, w); Mix<T>::axpy2(acc[1].x, acc[1].y, x0.y, w);
    Mix<T>::axpy2(acc[2].x, acc[2].y, x0.z, w); Mix<T>::axpy2(acc[3].x, acc[3].y, x0.w, w);
    Mix<T>::axpy2(acc[4].x, acc[4].y, x1.x, w); Mix<T>::axpy2(acc[5].x, acc[5].y, x1.y, w);
    Mix<T>::axpy2(acc[6].x, acc[6].y, x1.z, w); Mix<T>::axpy2(acc[7].x, acc[7].y, x1.w, w);
}

// The two lanes (h = 0, 1) of a (pixel, group) split the nine POINTS: lane h takes points 4h..4h+3 with all 16
// channels and both take point 8 with their own 8.  A lane reads its chunk 2g + h first and the partner's
// chunk 2g + 1 - h second, so in either LDS.128 the eight lanes of a quarter-warp still touch eight different
// bank groups; one shuffle exchange per pixel merges the halves.  4.5 locates per lane instead of 9.
// (The same split over L1 — fwd_pts_kernel — lost: there a 32-byte request costs L1 two passes.)
__device__ __forceinline__ uint32_t lds32s(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds16s(uint32_t addr) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}
#ifndef DCNV3_FWD_ROLL_IT
#define DCNV3_FWD_ROLL_IT 0  // 1: roll the loop over the lane's two pixels (5 432 -> 3 392 SASS instructions); measured slower, 55.7 vs 53.7 us at P3
#endif
#ifndef DCNV3_FWD_MIN_CTAS
#define DCNV3_FWD_MIN_CTAS 3  // A/B on one box (round 2): 4 CTAs per SM at 64 registers — see profiles/r02_fwd_occupancy.md
#endif
template <typename T, bool LOGITS, bool STAGE = false>  // STAGE: offsets / masks as TMA boxes (its own instance: the other
__global__ void __launch_bounds__(kFwdTileThreads, DCNV3_FWD_MIN_CTAS)  // path's code would only cost instruction-cache space)
fwd_tile_kernel(const T *__restrict__ in, const T *__restrict__ off, const T *__restrict__ mask,
                T *__restrict__ out, const Geo q, const int GQ
#ifdef DCNV3_FWD_TMA
                , const __grid_constant__ CUtensorMap tmap   // input as a 4-D tensor (C, W, H, N), box (64, 20, 20, 1)
#if DCNV3_FWD_TMA_STAGE
                , const __grid_constant__ CUtensorMap tmap_o // offsets (opitch, Wo, Ho, N), box (80, 8, 8, 1)
                , const __grid_constant__ CUtensorMap tmap_m // masks (mpitch, Wo, Ho, N), box (40, 8, 8, 1)
                , const int stage_tma
#endif
#endif
                ) {
    extern __shared__ __align__(128) unsigned char smem[];
#ifdef DCNV3_FWD_TMA
    __shared__ __align__(8) unsigned long long win_bar, st_bar;
#endif
    pdl_enter();
    const int tid = threadIdx.x;
    // grid = (tiles_x * GQ, tiles_y, N): no divisions by run-time extents except this one
    TileCoord tc;
    tc.tx = (int)(blockIdx.x / (unsigned)GQ); tc.gq = (int)(blockIdx.x - (unsigned)tc.tx * (unsigned)GQ);
    tc.ty = (int)blockIdx.y; tc.n = (int)blockIdx.z;
    const int wy0 = tc.ty * kTile + (q.half_h - q.ph) - kFhalo;  // input row of window cell (0, 0)
    const int wx0 = tc.tx * kTile + (q.half_w - q.pw) - kFhalo;
    const T *img = in + (size_t)tc.n * q.H * q.W * q.C + tc.gq * 64;
#ifdef DCNV3_FWD_TMA
#if DCNV3_FWD_TMA_STAGE
    // the tile's offsets / masks as two TMA boxes on their own mbarrier (80 / 40 elements per pixel: the group quad's 72 / 36
    // first; a box starts on a 16-byte boundary of its row, so the masks of an odd quad sit 8 bytes into theirs), when their
    // rows are 16-byte multiples; otherwise the lanes' own loads below
    constexpr bool tstage = STAGE;
    (void)stage_tma;
    const uint32_t stg_s = smem_u32(smem) + kFwinBytes;
    const uint32_t mshift = tstage ? (uint32_t)((tc.gq * kWarps * 9) & 7) * 2u : 0u;
    if (tstage && tid == 0) {
        const uint32_t bar_s = smem_u32(&st_bar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(kFstOffB + kFstMaskB) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(stg_s), "l"(reinterpret_cast<uint64_t>(&tmap_o)), "r"(bar_s), "r"(tc.gq * kWarps * 18), "r"(tc.tx * kTile), "r"(tc.ty * kTile), "r"(tc.n) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(stg_s + kFstOffB), "l"(reinterpret_cast<uint64_t>(&tmap_m)), "r"(bar_s), "r"((tc.gq * kWarps * 9) & ~7), "r"(tc.tx * kTile), "r"(tc.ty * kTile), "r"(tc.n) : "memory");
    }
#endif
    if (tid == 0) tma_load_4d(smem_u32(smem), &tmap, smem_u32(&win_bar), kFwinCells * 128, tc.gq * 64, wx0, wy0, tc.n);
#else
    fill_window_plain<T>(smem, in, img, q, wy0, wx0, tid);
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
    if (tid < kFzeroCells * 8) reinterpret_cast<uint4 *>(smem + kFwinCells * 128)[tid] = make_uint4(0u, 0u, 0u, 0u);

    const PtGeo pg{q.H, q.W, q.scale};
    const uint32_t win_s = smem_u32(smem);
    const int sub = tid & 7, h = sub & 1;          // chunk of the cell row: 2 * (group in quad) + channel half
    const int g = tc.gq * kWarps + (sub >> 1);
    const uint32_t own16 = (uint32_t)sub << 4, other16 = (uint32_t)(sub ^ 1) << 4;
    // kernel-grid coordinates of this lane's four whole points p = 4h + k: i = p / 3 (kernel_w), j = p % 3
    float fi[4], fj[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        fi[k] = h ? (float)((4 + k) / 3) : (float)(k / 3);
        fj[k] = h ? (float)((4 + k) % 3) : (float)(k % 3);
    }

    // two pixels per lane (tile rows 0-3 and 4-7); offsets of points 4h..4h+3 and 8, all nine masks
    uint32_t roff[2][5];
    float rm[2][9];
    int oy[2], ox[2];
    bool valid[2];
#if defined(DCNV3_FWD_TMA) && DCNV3_FWD_TMA_STAGE
    if (tstage) {  // CTA-uniform: the two boxes have landed (the window may still be in flight)
        if (tid == 0) tma_wait(smem_u32(&st_bar));
        __syncthreads();
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int px = it * 32 + (tid >> 3);
            oy[it] = tc.ty * kTile + (px >> 3);
            ox[it] = tc.tx * kTile + (px & 7);
            valid[it] = oy[it] < q.Ho && ox[it] < q.Wo;
            // (a pixel outside the map was zero-filled by the copy engine: zero offsets, zero masks)
            const uint32_t so = stg_s + px * kFstOffPx + (sub >> 1) * 36, sm = stg_s + kFstOffB + px * kFstMaskPx + (sub >> 1) * 18 + mshift;
#pragma unroll
            for (int k = 0; k < 4; ++k) roff[it][k] = lds32s(so + (4 * h + k) * 4);
            roff[it][4] = lds32s(so + 32);
            if (LOGITS) {
#pragma unroll
                for (int p = 0; p < 9; ++p) rm[it][p] = half_to_float<T>((unsigned short)lds16s(sm + p * 2));
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) rm[it][k] = half_to_float<T>((unsigned short)lds16s(sm + (4 * h + k) * 2));
                rm[it][8] = half_to_float<T>((unsigned short)lds16s(sm + 16));
            }
        }
    } else
#endif
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int px = it * 32 + (tid >> 3);
        oy[it] = tc.ty * kTile + (px >> 3);
        ox[it] = tc.tx * kTile + (px & 7);
        valid[it] = oy[it] < q.Ho && ox[it] < q.Wo;
        const size_t pixi = ((size_t)tc.n * q.Ho + oy[it]) * q.Wo + ox[it];
        const uint32_t *po = reinterpret_cast<const uint32_t *>(off) + pixi * (q.opitch >> 1) + g * 9;
        const unsigned short *pm = reinterpret_cast<const unsigned short *>(mask) + pixi * q.mpitch + g * 9;
#pragma unroll
        for (int k = 0; k < 4; ++k) roff[it][k] = valid[it] ? __ldg(po + 4 * h + k) : 0u;
        roff[it][4] = valid[it] ? __ldg(po + 8) : 0u;
        if (LOGITS) {
#pragma unroll
            for (int p = 0; p < 9; ++p) rm[it][p] = valid[it] ? half_to_float<T>(__ldg(pm + p)) : 0.f;
        } else {  // slots 0-3: this lane's points, slot 8: point 8
#pragma unroll
            for (int k = 0; k < 4; ++k) rm[it][k] = valid[it] ? half_to_float<T>(__ldg(pm + 4 * h + k)) : 0.f;
            rm[it][8] = valid[it] ? half_to_float<T>(__ldg(pm + 8)) : 0.f;
        }
    }
    // Mode of the CTA.  A point that leaves the window costs a divergent trip through global memory, so a
    // tile whose offsets are large (more than 1 in 16 of its lanes sees an offset coordinate of 5 px or more) runs the vector kernel's body instead: same result, ~the vector
    // kernel's speed, no dependence on the offset distribution.  Decided per CTA, on the device, no state.
    // (an offset coordinate with |off * scale| >= 5 px: the window reaches 5 px beyond the 3x3 grid; tested on the packed
    // storage bits — positive 16-bit floats order like integers)
    const uint32_t thr = storage_bits<T>(5.f / fabsf(q.scale));
    uint32_t far = 0u;
#pragma unroll
    for (int it = 0; it < 2; ++it)
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const uint32_t a = roff[it][k] & 0x7fff7fffu;
            far |= (a >= (thr << 16) || (a & 0xffffu) >= thr) ? 1u : 0u;
        }
#ifdef DCNV3_FWD_TMA
    if (tid == 0) tma_wait(smem_u32(&win_bar));
#else
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
    const bool cta_far = __syncthreads_count(far != 0u) * 16 > kFwdTileThreads;  // (also the barrier for the window)
    if (cta_far) {
#pragma unroll 1
        for (int it = 0; it < 2; ++it) {
            if (!valid[it]) continue;
            VecCoord c;
            c.pix = (unsigned)(((size_t)tc.n * q.Ho + oy[it]) * q.Wo + ox[it]);
            c.v = tc.gq * 8 + sub; c.g = g; c.n = tc.n; c.ho = oy[it]; c.wo = ox[it];
            fwd_vec_body<T, 16, 9, LOGITS>(c, in, off, mask, out, q);
        }
        return;
    }

    // The loop over the lane's two pixels can be rolled (-DDCNV3_FWD_ROLL_IT=1: the pixel's registers are picked with
    // selects, the kernel shrinks from 84 to 53 KB of SASS); the instruction-fetch stalls it removes (ncu no_instruction 4 %)
    // are worth less than the overlap between the two pixels' points it loses: 55.7 vs 53.7 us at P3.
#if DCNV3_FWD_ROLL_IT
#pragma unroll 1
#else
#pragma unroll
#endif
    for (int it = 0; it < 2; ++it) {
        uint32_t roff_c[5];
        float rm_c[9];
#pragma unroll
        for (int k = 0; k < 5; ++k) roff_c[k] = it ? roff[1][k] : roff[0][k];
#pragma unroll
        for (int k = 0; k < 9; ++k) rm_c[k] = it ? rm[1][k] : rm[0][k];
        const int oy_c = it ? oy[1] : oy[0], ox_c = it ? ox[1] : ox[0];
        const bool valid_c = it ? valid[1] : valid[0];
        // (a pair shares its pixel, so both lanes take the same branch and the shuffle below is safe)
        if (__ballot_sync(0xffffffffu, valid_c) == 0u) continue;
        float p0h_, p0w_;
        window_origin<float>(q, oy_c, ox_c, p0h_, p0w_);
        float mk[5];  // masks of points 4h..4h+3 and 8
        if (LOGITS) {
            softmax9<T, LOGITS>(rm_c);
#pragma unroll
            for (int k = 0; k < 4; ++k) mk[k] = h ? rm_c[4 + k] : rm_c[k];
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) mk[k] = rm_c[k];
        }
        mk[4] = rm_c[8];
        float2 acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = make_float2(0.f, 0.f);
        const T *img_g = img + (sub >> 1) * 16;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            if (k == 4) {  // the four whole points are done: merge the halves, then point 8 on 8 channels
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    acc[c].x += __shfl_xor_sync(0xffffffffu, acc[4 + c].x, 1);
                    acc[c].y += __shfl_xor_sync(0xffffffffu, acc[4 + c].y, 1);
                }
            }
            const float2 o = unpack2f<T>(roff_c[k]);
            const LeanPoint t = locate_lean(pg, p0h_, p0w_, k == 4 ? 2.f : fi[k], k == 4 ? 2.f : fj[k], o.x, o.y);
            const unsigned u = (unsigned)(t.w_low - wx0), v = (unsigned)(t.h_low - wy0);
            const bool inwin = u <= (unsigned)(kFwin - 2) && v <= (unsigned)(kFwin - 2);
            const bool fast = t.inside && inwin && valid_c;
            const float m = fast ? mk[k] : 0.f;
            const uint32_t a = win_s + (fast ? v * kFwin + u : (unsigned)kFwinCells) * 128u;  // closed point: the zero cells (0 * Inf would be NaN)
            const float hh = sub_rn(1.f, t.lh), hw = sub_rn(1.f, t.lw);
            const float hm = hh * m, lm = t.lh * m;
            // corner weights rounded to the storage dtype (2^-9 / 2^-11 relative): operands of the mixed-precision FMAs
            const uint32_t w1 = Mix<T>::weight(hm * hw), w2 = Mix<T>::weight(hm * t.lw), w3 = Mix<T>::weight(lm * hw), w4 = Mix<T>::weight(lm * t.lw);
            if (k < 4) {
                corner16<T>(acc, a + own16, a + other16, w1);
                corner16<T>(acc, a + own16 + 128, a + other16 + 128, w2);
                corner16<T>(acc, a + own16 + kFwin * 128, a + other16 + kFwin * 128, w3);
                corner16<T>(acc, a + own16 + kFwin * 128 + 128, a + other16 + kFwin * 128 + 128, w4);
            } else {
                float2 (&lo)[4] = reinterpret_cast<float2 (&)[4]>(acc);
                fma8<T>(lo, lds128(a + own16), w1);
                fma8<T>(lo, lds128(a + own16 + 128), w2);
                fma8<T>(lo, lds128(a + own16 + kFwin * 128), w3);
                fma8<T>(lo, lds128(a + own16 + kFwin * 128 + 128), w4);
            }
            if (t.inside && !inwin && valid_c) {
                // rare: the point left the staged window; gather from global memory with the reference's
                // validity (whole points before the merge: both chunks; point 8: the own chunk).  Inline on
                // purpose: as a non-inlined function the accumulators live in local memory (72 -> 97 us).
                Point<float> tt;
                locate<float>(q, p0h_, p0w_, k == 4 ? 2 : (h ? (4 + k) / 3 : k / 3), k == 4 ? 2 : (h ? (4 + k) % 3 : k % 3), o.x, o.y, tt);
                const float sm_h = tt.hh * mk[k], sm_l = tt.lh * mk[k];
                const float w[4] = {sm_h * tt.hw, sm_h * tt.lw, sm_l * tt.hw, sm_l * tt.lw};
                const bool ok[4] = {tt.ok1, tt.ok2, tt.ok3, tt.ok4};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (!ok[c]) continue;
                    const T *src = img_g + ((size_t)(tt.h_low + (c >> 1)) * q.W + (tt.w_low + (c & 1))) * q.C;
                    float2 (&lo)[4] = reinterpret_cast<float2 (&)[4]>(acc);
                    fma8f<T>(lo, __ldg(reinterpret_cast<const uint4 *>(src + 8 * h)), w[c]);
                    if (k < 4) {
                        float2 (&hi)[4] = reinterpret_cast<float2 (&)[4]>(acc[4]);
                        fma8f<T>(hi, __ldg(reinterpret_cast<const uint4 *>(src + 8 * (h ^ 1))), w[c]);
                    }
                }
            }
        }
        if (valid_c) {
            const size_t pix = ((size_t)tc.n * q.Ho + oy_c) * q.Wo + ox_c;
            uint4 r;
            r.x = pack2<T>(acc[0].x, acc[0].y); r.y = pack2<T>(acc[1].x, acc[1].y);
            r.z = pack2<T>(acc[2].x, acc[2].y); r.w = pack2<T>(acc[3].x, acc[3].y);
            *reinterpret_cast<uint4 *>(out + pix * q.C + g * 16 + h * 8) = r;
        }
    }
}

}  // namespace imat
}  // namespace dcnv3

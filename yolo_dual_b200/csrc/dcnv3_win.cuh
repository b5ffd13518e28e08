// dcnv3_b200 — round-2 default backward for 16-bit storage, group_channels = 16, 3x3 s1 d1 (ACC_TILE): ONE kernel
// (behind a zero fill of grad_input) instead of the four-launch chain zero_select -> bwd_imat -> guarded bwd_vec -> cast:
// no fp32 workspace, no selector kernel, no cast.  profiles/r02_bwd_kernel_history.md has every variant that was built
// and measured on the way (fused with everything resident at two CTAs per SM, split in two kernels, two roles in one
// launch, persistent with prefetch, phases aliased over one buffer, ...).
//
// win::bwd_win_kernel.  One CTA per unit = (image, 4-row x 8-column band of output pixels, 4 groups), 256 threads = 32
// pixels x 4 groups x 2 point halves.  The CTA stages the band's 12x16-cell x 64-channel window once (cp.async, zero fill
// outside the map = the reference's per-corner validity, dcnv3_im2col_cuda.cuh:57-75), unswizzled, so that the eight lanes
// of a pixel (4 groups x 2 halves) hit eight different bank groups whatever cells they sample (conflict-free LDS.128); each
// warp stages the offsets / masks of its own four pixels as whole 16- / 8-byte chunks.  Then, per sampling point, ONE
// location (locate_lean) feeds
//   * grad_offset / grad_mask: the four corner dots d_k = sum_c go[c] * x_k[c] as EXACT mixed-precision FMAs (PTX
//     fma.rn.f32.bf16 / .f16 -> FHFMA: 16-bit operands straight from the halves of the loaded registers, fp32 accumulation)
//         grad_mask   = sum_k w_k d_k                                                            (cuh:144)
//         grad_offset = scale * m * (hh (d2 - d1) + lh (d4 - d3), hw (d3 - d1) + lw (d4 - d2))   (cuh:114-139,145-146)
//   * grad_input: the lane adds w_k * m as packed fp16 pairs into its pixel's private row of the interpolation matrix
//     Wm[pixel][9 band rows relative to the pixel's own][16 window columns] (the two lanes of a pixel own the rows of even
//     / odd parity, one shuffle exchange, no atomics).
// After one CTA barrier the tensor cores expand the matrix,
//         GW[cell][ch] = sum_pixel Wm[pixel][cell] * go[pixel][ch]       mma.sync m16n8k16, fp32 accumulators
// and the band's 12x16-cell window leaves the SM once as packed 16-bit vector reductions (red.global.add.noftz.v4.bf16x2 /
// .f16x2) into the zero-filled grad_input: every partial is an fp32 sum of all the band's contributions to that cell, and a
// cell sees at most six of them.
// A band whose offsets are large (more than 1 lane in 4 with an offset coordinate of 3 px or more) runs the vector
// family's lane body instead (bwd_vec_lane: all three gradients, packed 16-bit reductions per contribution); a single
// point that leaves the window / the pixel's reach takes a per-point slow path (global gathers for the dots, reductions for
// that point only).
//
// Location arithmetic is the shared locate() sequence (locate_lean: same operations in the same order), so the
// integer contract is unchanged.  Reference semantics: dcnv3_col2im_gpu_kernel_* :278-839 + dcnv3_col2im_bilinear :82-147.
#pragma once

#include "dcnv3_imat.cuh"

namespace dcnv3 {
namespace win {

using imat::kTile;
using imat::kWarps;
using imat::LeanPoint;
using imat::PtGeo;
using imat::TileCoord;

constexpr int kThreadsW = 256;
constexpr int kWinW = 16;                                // window columns the band's interpolation matrix spans: 8 + 2 * 4
constexpr int kBandRows = 12;                            // window rows a 4-row band of pixels reaches
constexpr int kRelRows = 9;                              // band rows a pixel can reach: its own band row r .. r + 8 (|offset * scale| < 3 px)
constexpr int kRowB = kRelRows * kWinW * 2 + 16;         // 304 B per (pixel, group): 9 RELATIVE band rows x 16 window columns x 2 B + 16 B
                                                         // skew (76 words = 12 mod 32: the 8 rows of an ldmatrix hit distinct banks)
#ifndef DCNV3_WIN_GRP_SKEW
#define DCNV3_WIN_GRP_SKEW 16
#endif
// +16 B per group so the groups' rows start on different banks.  (A/B on one box, round 2: skews of 16 / 32 / 64 bytes
// all gave the same time at P3: the read-modify-write bank conflicts of the interpolation matrix are not what bounds the kernel.)
constexpr int kGrpB = 32 * kRowB + DCNV3_WIN_GRP_SKEW;   // 9 744 B per group
constexpr int kWmB = kWarps * kGrpB;                     // 38 976 B: [group][32 pixels][9 relative rows][16 columns]
constexpr int kGoGrpB = 32 * 32;                         // grad_output (B operand) of a group: 32 pixels x 32 B, 16-byte halves
constexpr int kGoB = kWarps * kGoGrpB;                   // XOR-swizzled by (pixel >> 2) & 1: 4 096 B
constexpr int kDwinB = kBandRows * kWinW * 128;          // the band's 12x16-cell x 64-channel window, 24 576 B (later: flush staging)
constexpr int kFlushWarpB = kDwinB / 8;                  // 3 072 B per warp = 6 band rows x 16 cells x 32 B
// Staging area for the band's offsets / masks (in) and grad_offset / grad_mask (out): the four groups of a pixel are one
// contiguous 144-byte (offsets) resp. 72-byte (masks) run in global memory, so they move as whole 16- / 8-byte
// chunks (cp.async in, vector stores out) instead of one 4- / 2-byte access per lane and point (ncu, round 2: those
// scalar accesses were 9.5 M of the launch's 38 M L1 wavefronts and moved 5x their bytes in L2 sectors).  Pixel pitch
// 160 B: lane (pixel i, group g, half h) reads word 40 i + 9 g + 4 h + k — all 32 banks distinct.
constexpr int kStOffPx = 160;
constexpr int kStMaskPx = 80;
constexpr int kStOffB = 32 * kStOffPx;                   // 5 120 B
constexpr int kStMaskB = 32 * kStMaskPx;                 // 2 560 B
constexpr int kStageB = kStOffB + kStMaskB;
// shared memory of a CTA: [window | staging | Wm | Gos | one zero row]; nothing aliases, so a warp walks from the dots of
// its lanes straight into their scatter without waiting for the other warps
constexpr int kStageOff = kDwinB;                        // (behind the window: TMA destinations are 128-byte aligned)
constexpr int kWmOff = kStageOff + kStageB;
constexpr int kGosOff = kWmOff + kWmB;
constexpr int kZeroOff = kGosOff + kGoB;                 // 32 zero bytes: ldmatrix rows of pixels that cannot reach a band row
static_assert(kStageOff % 128 == 0 && kStOffB % 128 == 0, "TMA destinations (offsets, masks) are 128-byte aligned");
#ifndef DCNV3_WIN_EXTRA_SMEM
#define DCNV3_WIN_EXTRA_SMEM 0  // occupancy experiments: pad the CTA's shared memory (profiles/r02_bwd_kernel_history.md)
#endif
constexpr int kSmemB = kZeroOff + 32 + DCNV3_WIN_EXTRA_SMEM;  // 75 360 B: three CTAs per SM (limit 76 800 - static)
static_assert(kSmemB - DCNV3_WIN_EXTRA_SMEM <= 76000, "three CTAs per SM");
constexpr int kFarLanes = kThreadsW / 4;                 // a band is "far" when more lanes than this see a >= 3 px offset
static_assert(kWmB % 16 == 0 && kGrpB % 16 == 0 && kRowB % 16 == 0, "ldmatrix rows are 16-byte aligned");
static_assert(kFlushWarpB == 6 * 16 * 32, "flush staging: 6 band rows x 16 cells x 32 B per warp");

using imat::Mix;

template <typename T> __device__ __forceinline__ void red_add_v4(T *p, const uint4 &v, bool pred);
template <> __device__ __forceinline__ void red_add_v4<__half>(__half *p, const uint4 &v, bool pred) { red_add_v4_f16x2(p, v, pred); }
template <> __device__ __forceinline__ void red_add_v4<__nv_bfloat16>(__nv_bfloat16 *p, const uint4 &v, bool pred) { red_add_v4_bf16x2(p, v, pred); }

__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v, bool pred) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q st.shared.u32 [%0], %1;\n\t}" ::"r"(addr), "r"(v), "r"((int)pred) : "memory");
}
__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src, int bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4 &v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Window gathers are plain (non-volatile) asm so that ptxas may hoist the next point's loads above the current
// point's arithmetic; what keeps them below the barrier that publishes the window is a data dependency: their
// base address passes through an `asm volatile` after that barrier (bwd_win_kernel).
__device__ __forceinline__ uint4 lds128_free(uint32_t addr) {
    uint4 v;
#ifdef DCNV3_WIN_LDS_VOLATILE
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
#else
    asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
#endif
    return v;
}

// dot of the pixel's 16 grad_output channels with one window cell's slab: own chunk first, partner's second
template <typename T>
__device__ __forceinline__ float corner_dot16(uint32_t a_own, uint32_t a_oth, const uint4 &g_own, const uint4 &g_oth) {
    const uint4 x0 = lds128_free(a_own), x1 = lds128_free(a_oth);
    float s0 = 0.f, s1 = 0.f;
    Mix<T>::dot2(s0, s1, x0.x, g_own.x); Mix<T>::dot2(s0, s1, x0.y, g_own.y);
    Mix<T>::dot2(s0, s1, x0.z, g_own.z); Mix<T>::dot2(s0, s1, x0.w, g_own.w);
    Mix<T>::dot2(s0, s1, x1.x, g_oth.x); Mix<T>::dot2(s0, s1, x1.y, g_oth.y);
    Mix<T>::dot2(s0, s1, x1.z, g_oth.z); Mix<T>::dot2(s0, s1, x1.w, g_oth.w);
    return s0 + s1;
}
template <typename T>
__device__ __forceinline__ float corner_dot8(uint32_t a_own, const uint4 &g_own) {
    const uint4 x0 = lds128_free(a_own);
    float s0 = 0.f, s1 = 0.f;
    Mix<T>::dot2(s0, s1, x0.x, g_own.x); Mix<T>::dot2(s0, s1, x0.y, g_own.y);
    Mix<T>::dot2(s0, s1, x0.z, g_own.z); Mix<T>::dot2(s0, s1, x0.w, g_own.w);
    return s0 + s1;
}

// packed 16-bit pair (two adjacent window columns of one band row) += into the pixel's Wm row.  Element e = band
// row * 16 + column (column <= 14).  An odd column straddles two words; the second store is predicated so that a
// lane never writes a word of a row it does not own (rows of the other parity belong to the partner lane).
__device__ __forceinline__ void wm_add(uint32_t row_s, uint32_t e, uint32_t pair, bool on) {
    const uint32_t wa = row_s + ((e >> 1) << 2);
    const bool odd = (e & 1u) != 0u;
#if DCNV3_WIN_PRED_W1
    const uint32_t w0 = lds32(wa);
    uint32_t w1 = 0u;  // the second word only where the pair straddles it: fewer active lanes, fewer bank conflicts
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q ld.shared.u32 %0, [%1];\n\t}" : "+r"(w1) : "r"(wa + 4), "r"((int)(on && odd)));
#else
    const uint32_t w0 = lds32(wa), w1 = lds32(wa + 4);
#endif
    sts32(wa, imat::add2<__half>(w0, odd ? (pair << 16) : pair), on);
    sts32(wa + 4, imat::add2<__half>(w1, pair >> 16), on && odd);
}

// w * (8 packed 16-bit grad_output values) -> 8 packed 16-bit values
template <typename T>
__device__ __forceinline__ uint4 scale_chunk(const uint4 &g, float w) {
    const uint32_t wd[4] = {g.x, g.y, g.z, g.w};
    uint32_t r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 f = imat::unpack2f<T>(wd[k]);
        r[k] = imat::pack2<T>(w * f.x, w * f.y);
    }
    return make_uint4(r[0], r[1], r[2], r[3]);
}
template <typename T>
__device__ __forceinline__ float dot_chunk(const uint4 &x, const uint4 &g) {
    float s0 = 0.f, s1 = 0.f;
    Mix<T>::dot2(s0, s1, x.x, g.x); Mix<T>::dot2(s0, s1, x.y, g.y);
    Mix<T>::dot2(s0, s1, x.z, g.z); Mix<T>::dot2(s0, s1, x.w, g.w);
    return s0 + s1;
}

#ifndef DCNV3_WIN_PRED_W1
#define DCNV3_WIN_PRED_W1 0
#endif
#ifndef DCNV3_WIN_TMA_ZERO
#define DCNV3_WIN_TMA_ZERO 1   // the interpolation matrix is zeroed by a bulk copy of zeros (TMA engine, on the window's mbarrier) instead of 10 STS.128 per thread: 148.9 -> 145.9 us at P3
#endif
#ifndef DCNV3_WIN_TMA_FLUSH
#define DCNV3_WIN_TMA_FLUSH 1  // (2: one box per warp half, 64 channels wide) a warp's 6 band rows x 16 cells x 16 channels leave as ONE TMA reduce-add (cp.reduce.async.bulk.tensor -> UTMAREDG.4D.ADD): 145.9 -> 137.7 us at P3
#endif
#ifndef DCNV3_WIN_TMA_STAGE
#define DCNV3_WIN_TMA_STAGE 1  // the tile's offsets / masks arrive as two TMA boxes (when their rows are 16-byte multiples) instead of per-warp cp.async chunks
#endif
#ifndef DCNV3_WIN_TMA  // -DDCNV3_NO_TMA: cp.async window fill, the threads' own zero fill and reductions
#undef DCNV3_WIN_TMA_ZERO
#define DCNV3_WIN_TMA_ZERO 0
#undef DCNV3_WIN_TMA_FLUSH
#define DCNV3_WIN_TMA_FLUSH 0
#undef DCNV3_WIN_TMA_STAGE
#define DCNV3_WIN_TMA_STAGE 0
#endif
#if DCNV3_WIN_TMA_FLUSH  // band rows of warp half q: contiguous (6q .. 6q + 5: one TMA box) or interleaved (q, q + 2, ..)
#define DCNV3_WIN_ROW(q, i) (6 * (q) + (i))
#else
#define DCNV3_WIN_ROW(q, i) ((q) + 2 * (i))
#endif
#ifndef DCNV3_WIN_SKIP_ZERO_MMA
#define DCNV3_WIN_SKIP_ZERO_MMA 1  // skip the two (k-step, band row) pairs per warp whose A operand is out of every pixel's reach
#endif
#ifndef DCNV3_WIN_STBULK
#define DCNV3_WIN_STBULK 0  // 1: one thread zeroes the interpolation matrix with st.bulk (UMEMSETS.64); measured +2 us at P3, profiles/r02_bwd_kernel_history.md #14
#endif
#ifndef DCNV3_WIN_COOP_SLOW
#define DCNV3_WIN_COOP_SLOW 1  // 0: every lane runs its own rare points (dots_point_slow / scatter_point_slow), round 2's first version
#endif
// ---- slow paths (rare: |offset * scale| >= 3 px resp. 5 px; non-inlined) ---------------------------------------
// Corner dots of a point outside the staged window: global gathers with the reference's per-corner validity.  `img_c`
// points at the first channel of chunk a of the group's slab, chunk b is `delta_b` elements away.
template <typename T>
__device__ __noinline__ float4 dots_point_slow(const PtGeo pg, const float p0h_, const float p0w_, const int i, const int j,
                                               const uint32_t offw, const uint4 ga, const uint4 gb, const T *img_c,
                                               const int delta_b, const int C) {
    const Geo q = imat::geo_of(pg);
    const float2 o = imat::unpack2f<T>(offw);
    Point<float> t;
    locate<float>(q, p0h_, p0w_, i, j, o.x, o.y, t);
    const bool ok[4] = {t.ok1, t.ok2, t.ok3, t.ok4};
    float d[4];
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
        d[k] = 0.f;
        if (!ok[k]) continue;
        const size_t e = ((size_t)(t.h_low + (k >> 1)) * q.W + (t.w_low + (k & 1))) * C;
        const uint4 xa = __ldg(reinterpret_cast<const uint4 *>(img_c + e));
        const uint4 xb = __ldg(reinterpret_cast<const uint4 *>(img_c + e + delta_b));
        d[k] = dot_chunk<T>(xa, ga) + dot_chunk<T>(xb, gb);
    }
    return make_float4(d[0], d[1], d[2], d[3]);
}

// grad_input contributions of a point outside the band's window: packed 16-bit vector reductions, chunk a of the
// group's slab (nred = 1: point 8, whose two lanes reduce one half each) or both chunks (nred = 2: a whole point).
template <typename T>
__device__ __noinline__ void scatter_point_slow(const PtGeo pg, const float p0h_, const float p0w_, const int i, const int j,
                                                const uint32_t offw, const float m, const uint4 ga, const T *go_b,
                                                const int nred, T *gin_c, const int delta_b, const int C) {
    const Geo q = imat::geo_of(pg);
    const float2 o = imat::unpack2f<T>(offw);
    Point<float> t;
    locate<float>(q, p0h_, p0w_, i, j, o.x, o.y, t);
    const float w[4] = {t.hh * t.hw, t.hh * t.lw, t.lh * t.hw, t.lh * t.lw};
    const bool ok[4] = {t.ok1, t.ok2, t.ok3, t.ok4};
    uint4 gb = make_uint4(0u, 0u, 0u, 0u);
    if (nred == 2) gb = __ldg(reinterpret_cast<const uint4 *>(go_b));
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
        if (!ok[k]) continue;
        const size_t e = ((size_t)(t.h_low + (k >> 1)) * q.W + (t.w_low + (k & 1))) * C;
        const float wm = w[k] * m;
        red_add_v4<T>(gin_c + e, scale_chunk<T>(ga, wm), true);
        if (nred == 2) red_add_v4<T>(gin_c + e + delta_b, scale_chunk<T>(gb, wm), true);
    }
}

// The vector family's lane body for one (pixel, 8 channels): all three gradients of a band whose offsets are large.
template <typename T, bool LOGITS>
__device__ __noinline__ void bwd_far_lane(const VecCoord c, const int h, const bool active, const T *in, const T *off,
                                          const T *mask, const T *gout, T *gin, T *goff, T *gmask, const Geo q) {
    bwd_vec_lane<T, T, 16, 9, LOGITS>(c, h, active, in, off, mask, gout, gin, goff, gmask, q, 2);
}

// Does this lane see an offset coordinate of 3 px or more among its five points?  (packed storage bits: positive
// 16-bit floats order like integers.)  Both kernels count these lanes per band with the same code on the same data.
template <typename T>
__device__ __forceinline__ bool lane_far(const uint32_t (&roff)[5], float scale) {
    const uint32_t thr = imat::storage_bits<T>(3.f / fabsf(scale));
    uint32_t far = 0u;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const uint32_t a = roff[k] & 0x7fff7fffu;
        far |= (a >= (thr << 16) || (a & 0xffffu) >= thr) ? 1u : 0u;
    }
    return far != 0u;
}

__device__ __forceinline__ uint32_t sel4(const uint32_t (&a)[4], int idx) {
    const uint32_t lo = (idx & 1) ? a[1] : a[0], hi = (idx & 1) ? a[3] : a[2];
    return (idx & 2) ? hi : lo;
}

// kernel-grid coordinates of a lane's point slot k: points 4h + k (k < 4) and 8 (k = 4); i = p / 3 (kernel_w), j = p % 3
__device__ __forceinline__ int slot_i(int k, int h) { return k == 4 ? 2 : (h ? (4 + k) / 3 : k / 3); }
__device__ __forceinline__ int slot_j(int k, int h) { return k == 4 ? 2 : (h ? (4 + k) % 3 : k % 3); }

// ===========================================================================================================
// Shared memory of a CTA (75 360 B, three CTAs per SM): [window 24 576 | staging 7 680 | Wm 38 976 | Gos 4 096 | zero row 32].
// Nothing aliases: the interpolation matrix is zeroed in the shadow of the first loads and a warp goes from a point's dots
// straight to the same point's matrix updates.  The one CTA barrier after the far-band vote is in front of the tensor
// cores; behind it the window is dead and becomes the flush staging area (3 KB per warp).
// The occupancy experiments (profiles/r02_bwd_kernel_history.md) are why a CTA carries a whole unit: the kernel's time is
// T = 136 us + 109 us / (CTAs per SM) at P3, and only co-resident work hides the per-CTA latency.
// ===========================================================================================================
__device__ __forceinline__ void stmatrix_x4(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
    asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}

// Zero fill of grad_input in front of bwd_win_kernel: a kernel of this library instead of cudaMemsetAsync, so that it is
// part of the programmatic launch chain (the backward kernel's CTAs are resident when it drains) and its DRAM
// traffic shows up in the ncu launch list next to the backward kernel's (profiles/r02_ncu_traffic.json).
__global__ void __launch_bounds__(256) zero_fill_kernel(uint4 *__restrict__ p, const size_t n16) {
    pdl_enter();
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n16; i += (size_t)gridDim.x * 256) p[i] = z;
}

#if DCNV3_WIN_TMA_ZERO
__device__ __align__(128) uint4 g_wm_zeros[kWmB / 16];  // zero-initialised: the source of the bulk copy that clears Wm
#endif
#ifndef DCNV3_WIN_MIN_CTAS
#define DCNV3_WIN_MIN_CTAS 3
#endif
// STRIP (maps whose width leaves 1-4 columns beyond the last whole 8-column tile, e.g. 20 x 20): those columns are not
// a third, half-empty tile column but a strip of TRANSPOSED tiles, 8 rows x 4 columns each (grid rows bands_y and up).
// A strip tile's window is 16 rows x 12 columns (the same 192 cells, row pitch 12); in the interpolation matrix, the
// tensor-core expansion and the flush the roles of x and y are swapped (A = the 16-cell axis, B = the 12-cell axis), so
// everything behind the window gathers is the same code.  20 x 20: 13 CTAs per (image, group quad) instead of 15.
template <typename T, bool LOGITS, bool STRIP>
__global__ void __launch_bounds__(kThreadsW, DCNV3_WIN_MIN_CTAS)
bwd_win_kernel(const T *__restrict__ in, const T *__restrict__ off, const T *__restrict__ mask,
               const T *__restrict__ gout, T *__restrict__ gin, T *__restrict__ goff, T *__restrict__ gmask,
               const Geo q, const int GQ, const int tiles_x, const int bands_y
#ifdef DCNV3_WIN_TMA
               , const __grid_constant__ CUtensorMap tmap   // input as a 4-D tensor (C, W, H, N), box (64, 16, 12, 1)
               , const __grid_constant__ CUtensorMap tmap_s // the same tensor, box (64, 12, 16, 1): strip tiles
               , const __grid_constant__ CUtensorMap tmap_o // offsets (opitch, Wo, Ho, N), box (80, 8, 4, 1)
               , const __grid_constant__ CUtensorMap tmap_m // masks (mpitch, Wo, Ho, N), box (40, 8, 4, 1)
#if DCNV3_WIN_TMA_FLUSH
               , const __grid_constant__ CUtensorMap tmap_r // grad_input (C, W, H, N) in the storage dtype, box (16, 16, 6, 1), 32-byte swizzle
#endif
#endif
               , const int strip_tiles, const int stage_tma) {
    extern __shared__ __align__(DCNV3_WIN_TMA_FLUSH == 2 ? 1024 : 256) unsigned char smem[];  // the period of the TMA 128- / 32-byte swizzle (flush staging)
    constexpr bool kScaled = std::is_same<T, __nv_bfloat16>::value;
    __shared__ __align__(16) uint32_t smax[8];
#ifdef DCNV3_WIN_TMA
    __shared__ __align__(8) unsigned long long win_bar, st_bar;
#endif
    pdl_enter();  // (waiting only in front of the first grad_input access instead measured nothing: 190.4 vs 191.0 us at P3)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    TileCoord tc;  // (image, band row, tile column, group quad)
#ifdef DCNV3_WIN_GRID1D
    {
        unsigned b = blockIdx.x;
        tc.gq = (int)(b % (unsigned)GQ); b /= (unsigned)GQ;
        tc.tx = (int)(b % (unsigned)tiles_x); b /= (unsigned)tiles_x;
        tc.ty = (int)(b % (unsigned)bands_y);
        tc.n = (int)(b / (unsigned)bands_y);
    }
#else
    // grid = (tiles_x * GQ, bands_y, N): one division by a run-time extent instead of three; same CTA order as a flat grid
    tc.tx = (int)(blockIdx.x / (unsigned)GQ); tc.gq = (int)(blockIdx.x - (unsigned)tc.tx * (unsigned)GQ);
    tc.ty = (int)blockIdx.y; tc.n = (int)blockIdx.z;
#endif
    int py0 = tc.ty * 4, px0 = tc.tx * kTile;               // output row / column of the tile's pixel 0
    bool tm = false;                                         // CTA-uniform: a strip tile (8 rows x 4 columns, A = y, B = x)
    if (STRIP && tc.ty >= bands_y) {
        const int t = (tc.ty - bands_y) * tiles_x + tc.tx;
        if (t >= strip_tiles) return;
        tm = true; py0 = t * 8; px0 = tiles_x * kTile;
    }
    (void)tiles_x; (void)bands_y; (void)strip_tiles;
    const int by0 = py0 + (q.half_h - q.ph) - 4;            // input row / column of window cell (0, 0)
    const int wx0 = px0 + (q.half_w - q.pw) - 4;
    // pixel p of the tile: (row, column) = (p >> 3, p & 7), transposed in a strip tile
    auto pix_y = [&](int p) { return py0 + (tm ? (p & 7) : (p >> 3)); };
    auto pix_x = [&](int p) { return px0 + (tm ? (p >> 3) : (p & 7)); };
    const size_t img_off = (size_t)tc.n * q.H * q.W * q.C + tc.gq * 64;
    const PtGeo pg{q.H, q.W, q.scale};
    const uint32_t smem_s = imat::smem_u32(smem);

    // the band's offsets and masks -> staging area, whole chunks (a pixel outside the map: zero fill).  A warp's lanes
    // are 4 pixels x 4 groups x 2 halves, so every warp stages (and later writes back) exactly the 4 x 9 chunks of its
    // own four pixels: no CTA barrier on either side, __syncwarp() is enough.
    const uint32_t stage_s = smem_s + kStageOff;
    const size_t gq_unit = (size_t)tc.gq * kWarps;
#if DCNV3_WIN_TMA_STAGE
    // a whole tile's offsets / masks as two TMA boxes: 80 / 40 elements per pixel (the staging area's pixel pitch; the group
    // quad's 72 / 36 come first), 8 columns x 4 rows; pixels outside the map are zero-filled.  Not for strip tiles (their
    // pixel order is transposed).
    const bool tstage = stage_tma != 0 && !(STRIP && tm);
    // a box starts on a 16-byte boundary of its row (else: illegal instruction): the masks of an odd group quad begin 8 bytes
    // into their box (36 elements per quad; the box is 40 wide), and so do the results that go back out of the staging area
    const uint32_t mshift = tstage ? (uint32_t)(((int)gq_unit * 9) & 7) * 2u : 0u;
    if (tstage) {
        if (tid == 0) {
            const uint32_t bar_s = imat::smem_u32(&st_bar);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(kStOffB + kStMaskB) : "memory");
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                         ::"r"(stage_s), "l"(reinterpret_cast<uint64_t>(&tmap_o)), "r"(bar_s), "r"((int)gq_unit * 18), "r"(px0), "r"(py0), "r"(tc.n) : "memory");
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                         ::"r"(stage_s + kStOffB), "l"(reinterpret_cast<uint64_t>(&tmap_m)), "r"(bar_s), "r"(((int)gq_unit * 9) & ~7), "r"(px0), "r"(py0), "r"(tc.n) : "memory");
        }
    } else
#else
    (void)stage_tma;
    constexpr bool tstage = false;
    constexpr uint32_t mshift = 0u;
#endif
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int t = it * 32 + lane;           // chunk 0..35 of the warp
        if (t < 36) {
            const int ppx = warp * 4 + t / 9, chk = t % 9;
            const int poy = pix_y(ppx), pox = pix_x(ppx);
            const bool ok = poy < q.Ho && pox < q.Wo;
            const size_t pp = ((size_t)tc.n * q.Ho + min(poy, q.Ho - 1)) * q.Wo + min(pox, q.Wo - 1);
            imat::cp_async16(stage_s + ppx * kStOffPx + chk * 16,
                             reinterpret_cast<const char *>(off) + (pp * q.opitch + gq_unit * 18) * 2 + chk * 16, ok ? 16 : 0);
            cp_async8(stage_s + kStOffB + ppx * kStMaskPx + chk * 8,
                      reinterpret_cast<const char *>(mask) + (pp * q.mpitch + gq_unit * 9) * 2 + chk * 8, ok ? 8 : 0);
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");  // group 0: this warp's offsets / masks
#ifdef DCNV3_WIN_TMA
    // The window as ONE TMA box load (cp.async.bulk.tensor.4d): coordinates may lie outside the map, the hardware fills
    // those cells with zeros (= the reference's per-corner validity); completion is signalled on an mbarrier that
    // thread 0 waits for in front of barrier A.
#if DCNV3_WIN_TMA_ZERO
    if (tid == 0) {  // the window box and the interpolation matrix's zeros on one mbarrier
        const uint32_t bar_s = imat::smem_u32(&win_bar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(kDwinB + kWmB) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(smem_s), "l"(reinterpret_cast<uint64_t>(STRIP && tm ? &tmap_s : &tmap)), "r"(bar_s), "r"(tc.gq * 64), "r"(wx0), "r"(by0), "r"(tc.n) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_s + kWmOff), "l"(reinterpret_cast<uint64_t>(g_wm_zeros)), "r"(kWmB), "r"(bar_s) : "memory");
    }
#else
    if (tid == 0) imat::tma_load_4d(smem_s, STRIP && tm ? &tmap_s : &tmap, imat::smem_u32(&win_bar), kDwinB, tc.gq * 64, wx0, by0, tc.n);
#endif
#else
    static_assert(!STRIP, "strip tiles are loaded by TMA");
    {  // stage the 12x16-cell x 64-channel window, unswizzled; zero outside the map
        const int ch = tid & 7, col = (tid >> 3) & 15, r0 = tid >> 7;
        const int ix = wx0 + col;
        const bool col_ok = (unsigned)ix < (unsigned)q.W;
        const size_t step = (size_t)2 * q.W * q.C;
        const T *p = in + img_off + ((long long)(by0 + r0) * q.W + ix) * q.C + ch * 8;
        uint32_t dst = smem_s + (r0 * kWinW + col) * 128 + (ch << 4);
        int iy = by0 + r0;
#pragma unroll
        for (int i = 0; i < kBandRows / 2; ++i) {
            const bool ok = col_ok && (unsigned)iy < (unsigned)q.H;
            imat::cp_async16(dst, ok ? p : in, ok ? 16 : 0);
            p += step; dst += 2 * kWinW * 128; iy += 2;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");  // group 1: the window (needed behind barrier A)
    }
#endif
#if DCNV3_WIN_STBULK
    // zero the interpolation matrix with ONE instruction of one thread (st.bulk -> UMEMSETS.64, sm_100); barrier A publishes
    // it.  Measured slower than the 10 x 16-byte stores per thread it replaces (154.0 vs 152.0 us at P3): kept for A/B only.
    if (tid == 32) asm volatile("st.bulk.weak.shared::cta [%0], %1, 0;" ::"r"(smem_s + kWmOff), "l"((unsigned long long)kWmB) : "memory");
    if (tid < 2) sts128(smem_s + kZeroOff + tid * 16, make_uint4(0u, 0u, 0u, 0u));
#elif DCNV3_WIN_TMA_ZERO
    if (tid < 2) sts128(smem_s + kZeroOff + tid * 16, make_uint4(0u, 0u, 0u, 0u));
#else
    {   // zero the interpolation matrix and the zero row (in the shadow of the loads above)
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int i = 0; i < (kWmB / 16 + kThreadsW - 1) / kThreadsW; ++i) {
            const int id = i * kThreadsW + tid;
            if (id < kWmB / 16) sts128(smem_s + kWmOff + id * 16, z);
        }
        if (tid < 2) sts128(smem_s + kZeroOff + tid * 16, z);
    }

#endif
    // lanes: 8 per pixel = 4 groups x 2 point halves; the band's 32 pixels
    const int px = tid >> 3, sub = tid & 7, gl = sub >> 1, h = sub & 1;
    const int g = tc.gq * kWarps + gl;
    const int oy = pix_y(px), ox = pix_x(px);
    const bool valid = oy < q.Ho && ox < q.Wo;
    const int cy = min(oy, q.Ho - 1), cx = min(ox, q.Wo - 1);  // clamped: addresses of an idle lane stay legal
    const size_t pix = ((size_t)tc.n * q.Ho + cy) * q.Wo + cx;
    const size_t unit = pix * q.G + g;
    const uint4 *gp = reinterpret_cast<const uint4 *>(gout + pix * q.C + g * 16);
    uint4 g_own = make_uint4(0u, 0u, 0u, 0u), g_oth = g_own;
    if (valid) {  // in flight while the staging copies land
        g_own = __ldg(gp + h);
        g_oth = __ldg(gp + (h ^ 1));
    }
#ifdef DCNV3_WIN_TMA
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#else
    asm volatile("cp.async.wait_group 1;" ::: "memory");
#endif
#if DCNV3_WIN_TMA_STAGE
    if (tstage) {  // CTA-uniform: the two boxes have landed (the window may still be in flight)
        if (tid == 0) imat::tma_wait(imat::smem_u32(&st_bar));
        __syncthreads();
    }
#endif
    __syncwarp();  // this warp's offsets / masks are staged (the window may still be in flight)

    // ---- this lane's offsets of points 4h..4h+3 and 8, their masks (all nine for the softmax), grad_output
    const uint32_t so_l = stage_s + px * kStOffPx + gl * 36, sm_l = stage_s + kStOffB + px * kStMaskPx + gl * 18 + mshift;
    uint32_t roff[5];
    float rm[LOGITS ? 9 : 5];
#pragma unroll
    for (int k = 0; k < 4; ++k) roff[k] = lds32(so_l + (4 * h + k) * 4);
    roff[4] = lds32(so_l + 32);
    if (LOGITS) {
#pragma unroll
        for (int p = 0; p < 9; ++p) rm[p] = valid ? imat::half_to_float<T>((unsigned short)lds16(sm_l + p * 2)) : 0.f;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) rm[k] = imat::half_to_float<T>((unsigned short)lds16(sm_l + (4 * h + k) * 2));
        rm[4] = imat::half_to_float<T>((unsigned short)lds16(sm_l + 16));
    }
    if constexpr (kScaled) {  // scatter role: largest |grad_output| of the band (bf16 magnitudes compare like integers)
        const uint32_t w[4] = {g_own.x, g_own.y, g_own.z, g_own.w};
        uint32_t mx = 0u;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t a = w[c] & 0x7fff7fffu;
            mx = max(mx, max(a >> 16, a & 0xffffu));
        }
        mx = __reduce_max_sync(0xffffffffu, mx);
        if (lane == 0) smax[warp] = mx;
    }
#ifdef DCNV3_WIN_TMA
    if (tid == 0) imat::tma_wait(imat::smem_u32(&win_bar));
#else
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
    // barrier A: the window is staged / Wm is zero and the per-warp maxima are visible.  A far band is skipped by the
    // dots CTA and handled as a whole by the scatter CTA (same count in both: same lanes, same offsets).
    const bool far_band = __syncthreads_count(lane_far<T>(roff, q.scale)) > kFarLanes;
    if (far_band) {
        VecCoord c;  // the vector family's lane body computes all three gradients of the band
        c.pix = (unsigned)pix; c.v = tc.gq * 8 + sub; c.g = g; c.n = tc.n; c.ho = cy; c.wo = cx;
        bwd_far_lane<T, LOGITS>(c, h, valid, in, off, mask, gout, gin, goff, gmask, q);
        return;
    }

    float p0h_, p0w_;
    window_origin<float>(q, oy, ox, p0h_, p0w_);
    float mk[5];
    if (LOGITS) {
        imat::softmax9<T, LOGITS>(reinterpret_cast<float (&)[9]>(rm));
#pragma unroll
        for (int k = 0; k < 4; ++k) mk[k] = h ? rm[(4 + k) % (LOGITS ? 9 : 5)] : rm[k];
        mk[4] = rm[LOGITS ? 8 : 4];
    } else {
#pragma unroll
        for (int k = 0; k < 5; ++k) mk[k] = rm[k];
    }

    // Per point, ONE location: the corner dots out of the window (grad_offset / grad_mask) and the point's packed fp16
    // contributions into the pixel's interpolation-matrix row (grad_input) — the matrix has its own memory, zeroed before
    // barrier A, and a lane only touches its own rows of it until barrier B, so nothing separates the two.
    const uint32_t wm_s = smem_s + kWmOff, gos_base = smem_s + kGosOff;
    const uint32_t row_s = wm_s + gl * kGrpB + px * kRowB;
    const int rband = px >> 3;  // band row of this lane's pixel: its Wm row holds window rows rband .. rband + 8
    uint32_t slowmask2 = 0u;    // points whose grad_input contributions leave the pixel's reach
    {
        // =================================================================== grad_offset / grad_mask (+ Wm build)
        const uint32_t own16 = (uint32_t)sub << 4, oth16 = (uint32_t)(sub ^ 1) << 4;
        const T *img_g = in + img_off + gl * 16;
        uint32_t win_s = smem_s;
        asm volatile("" : "+r"(win_s));  // the window gathers below depend on this: they stay behind barrier A
        const uint32_t pitch = STRIP && tm ? 12u : (uint32_t)kWinW;  // cells per window row
        const uint32_t rowb = pitch * 128u;
        uint32_t res_off[5];
        float res_m[5];
        uint32_t slowmask = 0u;
        // one straight-line block: a point that is not window-resident reads cell 0 and its results are dropped
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const float2 o = imat::unpack2f<T>(roff[k]);
            const LeanPoint t = imat::locate_lean(pg, p0h_, p0w_, (float)slot_i(k, h), (float)slot_j(k, h), o.x, o.y);
            const unsigned wxu = (unsigned)(t.w_low - wx0), wyu = (unsigned)(t.h_low - by0);
            const unsigned u = STRIP && tm ? wyu : wxu, vb = STRIP && tm ? wxu : wyu;  // index along A (16 cells) / B (12 cells)
            const bool inband = u <= (unsigned)(kWinW - 2) && vb <= (unsigned)(kBandRows - 2);
            const bool fast = t.inside && inband && valid;
            slowmask |= (t.inside && !inband && valid ? 1u : 0u) << k;  // same for both lanes of the pair
            const uint32_t a = win_s + (fast ? wyu * pitch + wxu : 0u) * 128u;
            float d0, d1, d2, d3;
            if (k < 4) {
                d0 = corner_dot16<T>(a + own16, a + oth16, g_own, g_oth);
                d1 = corner_dot16<T>(a + own16 + 128, a + oth16 + 128, g_own, g_oth);
                d2 = corner_dot16<T>(a + own16 + rowb, a + oth16 + rowb, g_own, g_oth);
                d3 = corner_dot16<T>(a + own16 + rowb + 128, a + oth16 + rowb + 128, g_own, g_oth);
            } else {  // point 8: the two lanes take 8 channels each and meet
                d0 = corner_dot8<T>(a + own16, g_own);
                d1 = corner_dot8<T>(a + own16 + 128, g_own);
                d2 = corner_dot8<T>(a + own16 + rowb, g_own);
                d3 = corner_dot8<T>(a + own16 + rowb + 128, g_own);
                d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
                d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
                d2 += __shfl_xor_sync(0xffffffffu, d2, 1);
                d3 += __shfl_xor_sync(0xffffffffu, d3, 1);
            }
            const float lh = t.lh, lw = t.lw;
            const float hh = sub_rn(1.f, lh), hw = sub_rn(1.f, lw);
            const float s_m = (hh * hw) * d0 + (hh * lw) * d1 + (lh * hw) * d2 + (lh * lw) * d3;
            const float s_w = hh * (d1 - d0) + lh * (d3 - d2);
            const float s_h = hw * (d2 - d0) + lw * (d3 - d1);
            const float sm = q.scale * mk[k];
            res_off[k] = fast ? imat::pack2<T>(sm * s_w, sm * s_h) : 0u;
            res_m[k] = fast ? s_m : 0.f;
            {   // interpolation-matrix pairs (top: band row vb, bottom: vb + 1), stored at the row RELATIVE to the pixel's
                // band row.  The lane adds the pair whose row has parity h and hands the other one to its partner, so
                // neither ever writes a row of the other's.  (Point 8 is known to both lanes: each adds the row of its parity.)
                const unsigned rel = vb - (unsigned)rband;  // 0 .. 7 when in reach
                const bool fast2 = t.inside && valid && u <= (unsigned)(kWinW - 2) && rel <= (unsigned)(kRelRows - 2);
                slowmask2 |= (t.inside && valid && !fast2 ? 1u : 0u) << k;  // same for both lanes of the pair
                const uint32_t e = fast2 ? rel * kWinW + u : 0u;
                const float hm = hh * mk[k], lm = lh * mk[k];
                // pairs run along A: (x, x + 1) of rows y / y + 1 — in a strip tile (y, y + 1) of columns x / x + 1
                const float w01 = hm * lw, w10 = lm * hw;
                const uint32_t top = imat::pack2<__half>(hm * hw, STRIP && tm ? w10 : w01);
                const uint32_t bot = imat::pack2<__half>(STRIP && tm ? w01 : w10, lm * lw);
                const bool keep_top = ((vb & 1u) == (unsigned)h);
                wm_add(row_s, keep_top ? e : e + kWinW, keep_top ? top : bot, fast2);
                if (k < 4) {
                    const uint32_t se = (keep_top ? e + kWinW : e) | (fast2 ? 256u : 0u);
                    const uint32_t re = __shfl_xor_sync(0xffffffffu, se, 1);
                    const uint32_t rw = __shfl_xor_sync(0xffffffffu, keep_top ? bot : top, 1);
                    wm_add(row_s, re & 255u, rw, (re & 256u) != 0u);
                }
            }
        }
#if DCNV3_WIN_COOP_SLOW
        // Rare points (|offset * scale| >= 3 px: inside the map but outside the band's window -> dots, or out of the
        // pixel's reach -> grad_input) are handled by the WARP, four at a time: a lane that runs one alone costs the
        // warp ~400 issue slots per point (measured: 16 us of 164 at P3 with offsets ~ N(0, 1) px,
        // profiles/r02_slow_path.md).  Eight lanes per point = 4 corners x 2 channel chunks: one 16-byte gather, 8 FMAs
        // and one packed vector reduction each; the point's owner gets its dots back by shuffle.
        {
            uint32_t need = slowmask | (slowmask2 << 8);          // bits 0-4: dots, bits 8-12: grad_input
            if (h) need &= ~((1u << 4) | (1u << 12));             // point 8 is known to both lanes: the even one reports it
            const int grp = lane >> 3, crn = (lane >> 1) & 3, hf = lane & 1;
            while (true) {
                const uint32_t bal = __ballot_sync(0xffffffffu, need != 0u);
                if (bal == 0u) break;
                int src = -1;                                     // lane that owns this group's point
                {
                    uint32_t bb = bal;
#pragma unroll
                    for (int gi = 0; gi < 4; ++gi) {
                        if (gi == grp) src = __ffs(bb) - 1;
                        bb &= bb - 1u;
                    }
                }
                const bool act = src >= 0;
                const int srcl = act ? src : 0;
                // owner side: its lowest pending slot
                const int ks = __ffs((need | (need >> 8)) & 31u) - 1;   // -1: nothing pending
                uint32_t s_off = 0u, s_fl = 0u;
                float s_mk = 0.f;
#pragma unroll
                for (int k = 0; k < 5; ++k)
                    if (k == ks) {
                        s_off = roff[k]; s_mk = mk[k];
                        s_fl = (uint32_t)k | (((need >> k) & 1u) << 3) | (((need >> (8 + k)) & 1u) << 4);
                    }
                const uint32_t offw = __shfl_sync(0xffffffffu, s_off, srcl);
                const float pm = __shfl_sync(0xffffffffu, s_mk, srcl);
                const uint32_t fl = __shfl_sync(0xffffffffu, s_fl, srcl);
                const float sp0h = __shfl_sync(0xffffffffu, p0h_, srcl), sp0w = __shfl_sync(0xffffffffu, p0w_, srcl);
                const int spx = warp * 4 + (srcl >> 3), sgl = (srcl >> 1) & 3, sh = srcl & 1;   // the owner's pixel / group / half
                const int sk = (int)(fl & 7u);
                const bool want_d = act && ((fl >> 3) & 1u), want_s = act && ((fl >> 4) & 1u);
                const int pnt = sk == 4 ? 8 : 4 * sh + sk, pi = pnt / 3, pj = pnt - 3 * pi;
                const Geo gq_ = imat::geo_of(pg);
                const float2 o = imat::unpack2f<T>(offw);
                Point<float> t;
                locate<float>(gq_, sp0h, sp0w, pi, pj, o.x, o.y, t);
                const bool okc = crn == 0 ? t.ok1 : crn == 1 ? t.ok2 : crn == 2 ? t.ok3 : t.ok4;
                const float wc = (crn & 2 ? t.lh : t.hh) * (crn & 1 ? t.lw : t.hw);
                const size_t spix = ((size_t)tc.n * q.Ho + pix_y(spx)) * q.Wo + pix_x(spx);
                const size_t ce = img_off + (okc ? ((size_t)(t.h_low + (crn >> 1)) * q.W + (t.w_low + (crn & 1))) * q.C : 0)
                                  + sgl * 16 + hf * 8;
                uint4 gch = make_uint4(0u, 0u, 0u, 0u);
                if (act) gch = __ldg(reinterpret_cast<const uint4 *>(gout + spix * q.C + (tc.gq * kWarps + sgl) * 16 + hf * 8));
                float d = 0.f;
                if (want_d && okc) d = dot_chunk<T>(__ldg(reinterpret_cast<const uint4 *>(in + ce)), gch);
                d += __shfl_xor_sync(0xffffffffu, d, 1);
                const float d0 = __shfl_sync(0xffffffffu, d, grp * 8), d1 = __shfl_sync(0xffffffffu, d, grp * 8 + 2);
                const float d2 = __shfl_sync(0xffffffffu, d, grp * 8 + 4), d3 = __shfl_sync(0xffffffffu, d, grp * 8 + 6);
                const float c_m = (t.hh * t.hw) * d0 + (t.hh * t.lw) * d1 + (t.lh * t.hw) * d2 + (t.lh * t.lw) * d3;
                const float c_w = t.hh * (d1 - d0) + t.lh * (d3 - d2);
                const float c_h = t.hw * (d2 - d0) + t.lw * (d3 - d1);
                const float sm = q.scale * pm;
                const uint32_t c_off = imat::pack2<T>(sm * c_w, sm * c_h);
                red_add_v4<T>(want_s && okc ? gin + ce : gin, scale_chunk<T>(gch, wc * pm), want_s && okc);
                // back to the owners: the r-th pending lane was served by group r
                const int rnk = __popc(bal & ((1u << lane) - 1u));
                const uint32_t got_off = __shfl_sync(0xffffffffu, c_off, (rnk & 3) * 8);
                const float got_m = __shfl_sync(0xffffffffu, c_m, (rnk & 3) * 8);
                if (need != 0u && rnk < 4) {
#pragma unroll
                    for (int k = 0; k < 5; ++k)
                        if (k == ks && ((need >> k) & 1u)) { res_off[k] = got_off; res_m[k] = got_m; }
                    need &= ~((1u | (1u << 8)) << ks);
                }
            }
        }
#else
        if (slowmask) {  // rare: points inside the map but outside the band's window (|offset * scale| >= 3 px)
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                if (!((slowmask >> k) & 1u)) continue;
                const float4 d = dots_point_slow<T>(pg, p0h_, p0w_, slot_i(k, h), slot_j(k, h), roff[k], g_own, g_oth,
                                                    img_g + 8 * h, h ? -8 : 8, q.C);
                const float2 o = imat::unpack2f<T>(roff[k]);
                const LeanPoint t = imat::locate_lean(pg, p0h_, p0w_, (float)slot_i(k, h), (float)slot_j(k, h), o.x, o.y);
                const float lh = t.lh, lw = t.lw;
                const float hh = sub_rn(1.f, lh), hw = sub_rn(1.f, lw);
                const float s_m = (hh * hw) * d.x + (hh * lw) * d.y + (lh * hw) * d.z + (lh * lw) * d.w;
                const float s_w = hh * (d.y - d.x) + lh * (d.w - d.z);
                const float s_h = hw * (d.z - d.x) + lw * (d.w - d.y);
                const float sm = q.scale * mk[k];
                res_off[k] = imat::pack2<T>(sm * s_w, sm * s_h);
                res_m[k] = s_m;
            }
        }
#endif
        // grad_offset / grad_mask (cuh:144-146); fused softmax: dl_p = m_p (gm_p - sum_q m_q gm_q)
        if (LOGITS) {
            float dot = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) dot = fmaf(mk[k], res_m[k], dot);
            if (h == 0) dot = fmaf(mk[4], res_m[4], dot);
            dot += __shfl_xor_sync(0xffffffffu, dot, 1);
#pragma unroll
            for (int k = 0; k < 5; ++k) res_m[k] = mk[k] * (res_m[k] - dot);
        }
        // results -> the warp's own part of the staging area -> whole chunks to global memory
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            sts32(so_l + (4 * h + k) * 4, res_off[k], true);
            sts16(sm_l + (4 * h + k) * 2, imat::pack2<T>(res_m[k], 0.f) & 0xffffu);
        }
        if (h == 0) {
            sts32(so_l + 32, res_off[4], true);
            sts16(sm_l + 16, imat::pack2<T>(res_m[4], 0.f) & 0xffffu);
        }
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int t = it * 32 + lane;
            if (t < 36) {
                const int ppx = warp * 4 + t / 9, chk = t % 9;
                const int poy = pix_y(ppx), pox = pix_x(ppx);
                if (poy < q.Ho && pox < q.Wo) {
                    const size_t pp = ((size_t)tc.n * q.Ho + poy) * q.Wo + pox;
                    *reinterpret_cast<uint4 *>(reinterpret_cast<char *>(goff) + (pp * q.opitch + gq_unit * 18) * 2 + chk * 16) =
                        imat::lds128(stage_s + ppx * kStOffPx + chk * 16);
                    *reinterpret_cast<uint2 *>(reinterpret_cast<char *>(gmask) + (pp * q.mpitch + gq_unit * 9) * 2 + chk * 8) =
                        lds64(stage_s + kStOffB + ppx * kStMaskPx + chk * 8 + mshift);
                }
            }
        }
    }

    // ======================================================================= grad_input: expansion + flush
    T *gin_g = gin + img_off + gl * 16;
    // B operand of the mma (an idle lane stores zeros).  The interpolation matrix is fp16 for both storage dtypes
    // (11-bit weights; bf16 weights, 8 bits, miss the atol 2e-3 bar on ~1e-5 of the elements), so bf16 grad_output enters
    // the product as fp16 after a power-of-two scaling per band: go' = go * 2^(127 - e_ref), e_ref = biased exponent of
    // the band's largest |go|.  Exact for every value within 2^-14 of that maximum, smaller ones lose low bits
    // (absolute error below 2^-24 of the band's largest gradient).
    int e_ref = 127;
    {
        const uint32_t gos_s = gos_base + gl * kGoGrpB + px * 32 + ((uint32_t)(h ^ ((px >> 2) & 1)) << 4);
        if constexpr (kScaled) {
            const uint4 s0 = *reinterpret_cast<const uint4 *>(smax), s1 = *reinterpret_cast<const uint4 *>(smax + 4);
            const uint32_t mx = max(max(max(s0.x, s0.y), max(s0.z, s0.w)), max(max(s1.x, s1.y), max(s1.z, s1.w)));
            e_ref = min(max((int)(mx >> 7), 1), 253);
            const float sc = __uint_as_float((uint32_t)(254 - e_ref) << 23);
            const uint32_t w[4] = {g_own.x, g_own.y, g_own.z, g_own.w};
            uint32_t r[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float2 f = imat::unpack2f<T>(w[c]);
                r[c] = imat::pack2<__half>(f.x * sc, f.y * sc);
            }
            sts128(gos_s, make_uint4(r[0], r[1], r[2], r[3]));
        } else {
            sts128(gos_s, g_own);
        }
    }
#if !DCNV3_WIN_COOP_SLOW
    if (slowmask2) {  // rare: points inside the map but out of the pixel's reach (|offset * scale| >= 3 px)
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            if (!((slowmask2 >> k) & 1u)) continue;
            scatter_point_slow<T>(pg, p0h_, p0w_, slot_i(k, h), slot_j(k, h), roff[k], mk[k], g_own,
                                  reinterpret_cast<const T *>(gp + (h ^ 1)), k < 4 ? 2 : 1, gin_g + 8 * h, h ? -8 : 8, q.C);
        }
    }
#endif
    __syncthreads();  // barrier B: Wm and Gos of the band are complete; every warp has left the window

    // ---- tensor cores: GW[band row][cell][ch] = Wm^T * go.  Warp = (group, parity of its band rows); m-tile = one band
    // row (16 cells), K = the band's 32 pixels (two k-steps), N = the group's 16 channels (two n-tiles).  A pixel of band
    // row r stores band row rr at relative row rr - r; out of its reach (rr - r outside 0 .. 8) it reads the zero row.
    const int mg = warp & 3, qpar = warp >> 2;
    const uint32_t wm_g = wm_s + mg * kGrpB;
    float gw[kBandRows / 2][2][4];
#pragma unroll
    for (int i = 0; i < kBandRows / 2; ++i)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int c = 0; c < 4; ++c) gw[i][nt][c] = 0.f;
    {
        const int jm = lane >> 3, jr = lane & 7;
        const uint32_t go_g = gos_base + mg * kGoGrpB;
        const uint32_t zero_a = smem_s + kZeroOff + (jm & 1) * 16;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            // B = go [16 px x 16 ch]: matrix jm = (px 8(jm & 1).., ch 8(jm >> 1)..); the 16-byte half is swizzled by (px >> 2) & 1
            uint32_t b00, b01, b10, b11;
            {
                const int bpx = 16 * s + 8 * (jm & 1) + jr;
                imat::ldmatrix_x4_trans(b00, b01, b10, b11, go_g + bpx * 32 + ((uint32_t)((jm >> 1) ^ ((bpx >> 2) & 1)) << 4));
            }
            // A = Wm^T [16 cells x 16 px]: matrix jm = (cells 8(jm & 1).., px 8(jm >> 1)..): all 8 pixels of a matrix share a band row
            const int apx = 16 * s + 8 * (jm >> 1) + jr, ar = 2 * s + (jm >> 1);
            const uint32_t abase = wm_g + apx * kRowB + (jm & 1) * 16;
#pragma unroll
            for (int i = 0; i < kBandRows / 2; ++i) {
                const int rr = DCNV3_WIN_ROW(qpar, i);  // band row
#if DCNV3_WIN_SKIP_ZERO_MMA
                // band rows 10 / 11 are out of reach of the pixels of band rows 0 / 1 (k-step 0), band rows 0 / 1 lie above
                // the pixels of band rows 2 / 3 (k-step 1): A = 0 whatever the data  (warp-uniform)
                if (s == 0 ? rr >= 10 : rr <= 1) continue;
#endif
                const int rel = rr - ar;
                uint32_t a0, a1, a2, a3;
                imat::ldmatrix_x4_trans(a0, a1, a2, a3, (unsigned)rel < (unsigned)kRelRows ? abase + rel * (kWinW * 2) : zero_a);
                imat::mma_16816<__half>(gw[i][0], a0, a1, a2, a3, b00, b01);
                imat::mma_16816<__half>(gw[i][1], a0, a1, a2, a3, b10, b11);
            }
        }
    }

    // ---- flush: the warp's 6 band rows x 16 cells x 16 channels leave as packed 16-bit vector reductions.  The mma
    // fragments (c0, c1) / (c2, c3) of n-tile nt = channels 8nt + 2tq, +1 of cells gID / gID + 8, packed to the storage
    // dtype, are exactly stmatrix fragments: one stmatrix.x4 lays a band row out as [16 cells][16 channels] — in this
    // warp's 3 KB of the window, which nobody reads any more (barrier B) — and every lane reads back the 8 channels of
    // ONE (cell, half).  A cell is 32 B; its halves swap places for cells 4-7 / 12-15 so that the 8 rows of an stmatrix
    // matrix (and the 8 lanes of an LDS.128 phase) fall into 8 different 16-byte bank groups.
    {
        const int jm = lane >> 3, jr = lane & 7;
        const int scell = jr + 8 * (jm & 1);
#if DCNV3_WIN_TMA_FLUSH == 2
        // staging of a warp half (4 groups): [6 band rows][16 cells][64 channels = 128 B], the cell's eight 16-byte chunks
        // XOR-ed with cell & 7 (= CU_TENSOR_MAP_SWIZZLE_128B; also what keeps the stmatrix rows on distinct bank groups)
        constexpr uint32_t kStRow = kWinW * 128;
        const uint32_t fb = smem_s + qpar * (6 * kStRow);
        const uint32_t st_addr = fb + scell * 128 + ((uint32_t)((2 * mg + (jm >> 1)) ^ (scell & 7)) << 4);
#else
        constexpr uint32_t kStRow = 512;
        const uint32_t fb = smem_s + warp * kFlushWarpB;
        const uint32_t st_addr = fb + scell * 32 + ((uint32_t)((jm >> 1) ^ ((scell >> 2) & 1)) << 4);
#endif
        const float unscale = __uint_as_float((uint32_t)e_ref << 23);  // 2^(e_ref - 127), bf16 storage only
#pragma unroll
        for (int i = 0; i < kBandRows / 2; ++i) {
            if constexpr (kScaled) {
#pragma unroll
                for (int n2 = 0; n2 < 2; ++n2)
#pragma unroll
                    for (int c = 0; c < 4; ++c) gw[i][n2][c] *= unscale;
            }
            stmatrix_x4(st_addr + i * kStRow, imat::pack2<T>(gw[i][0][0], gw[i][0][1]), imat::pack2<T>(gw[i][0][2], gw[i][0][3]),
                        imat::pack2<T>(gw[i][1][0], gw[i][1][1]), imat::pack2<T>(gw[i][1][2], gw[i][1][3]));
        }
#if DCNV3_WIN_TMA_FLUSH
        // (a box that STARTS at a negative coordinate is an illegal instruction for the reduce — unlike the loads, and unlike
        // a box that overhangs the far edges, which is clipped: tools/tma_reduce_probe.cu — so the warps of the left tile
        // column and of the top band's upper half keep their own reductions)
        if (!(STRIP && tm) && wx0 >= 0 && by0 + 6 * qpar >= 0) {
            // the staging area is exactly a TMA box in shared memory ([6 rows][16 cells][32 B], the two 16-byte halves of a
            // cell swapped where address bit 7 is set = CU_TENSOR_MAP_SWIZZLE_32B): the copy engine adds it into grad_input,
            // clips what lies outside the map, and the warp issues no loads, predicates or reductions of its own
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#if DCNV3_WIN_TMA_FLUSH == 2
            // one box per warp half: 64 channels (128-byte rows for the copy engine instead of four times as many 32-byte
            // ones) x 16 cells x 6 band rows; the half's four warps meet at a named barrier, one lane issues
            if (qpar) asm volatile("bar.sync 2, 128;" ::: "memory"); else asm volatile("bar.sync 1, 128;" ::: "memory");
            if (mg == 0 && lane == 0) {
                asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                             ::"l"(reinterpret_cast<uint64_t>(&tmap_r)), "r"(fb), "r"(tc.gq * 64), "r"(wx0),
                               "r"(by0 + 6 * qpar), "r"(tc.n) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the engine has read the staging area
            }
#else
            __syncwarp();
            if (lane == 0) {
                asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                             ::"l"(reinterpret_cast<uint64_t>(&tmap_r)), "r"(fb), "r"(tc.gq * 64 + mg * 16), "r"(wx0),
                               "r"(by0 + 6 * qpar), "r"(tc.n) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the engine has read the staging area
            }
#endif
            return;
        }
#endif
        __syncwarp();
        const int cell = lane >> 1, half = lane & 1;
        // the lane's cell runs along A (x; y in a strip tile), the warp's band rows along B
        const bool sw = STRIP && tm;
        const int ca = (sw ? by0 : wx0) + cell, cb0 = (sw ? wx0 : by0);
        const unsigned ext_a = (unsigned)(sw ? q.H : q.W), ext_b = (unsigned)(sw ? q.W : q.H);
        const size_t str_a = sw ? (size_t)q.W * q.C : (size_t)q.C, str_b = sw ? (size_t)q.C : (size_t)q.W * q.C;
        const bool a_ok = (unsigned)ca < ext_a;
#if DCNV3_WIN_TMA_FLUSH == 2
        const uint32_t ld_addr = fb + cell * 128 + ((uint32_t)((2 * mg + half) ^ (cell & 7)) << 4);
#else
        const uint32_t ld_addr = fb + cell * 32 + ((uint32_t)(half ^ ((cell >> 2) & 1)) << 4);
#endif
        T *dst0 = gin + img_off + mg * 16 + half * 8 + (size_t)ca * str_a;
#pragma unroll
        for (int i = 0; i < kBandRows / 2; ++i) {
            const int cb = cb0 + DCNV3_WIN_ROW(qpar, i);
            const uint4 o = imat::lds128(ld_addr + i * kStRow);
            const bool nz = ((o.x | o.y | o.z | o.w) & 0x7fff7fffu) != 0u;
            const bool ok = a_ok && (unsigned)cb < ext_b && nz;
            red_add_v4<T>(ok ? dst0 + (size_t)cb * str_b : gin, o, ok);
        }
    }
}

}  // namespace win
}  // namespace dcnv3

// dcnv3_b200 — privatised backward for group_channels == 16.
//
// Why: with global reductions the backward is bound by the SM -> crossbar request port
// (l1tex__m_l1tex2xbar_req_cycles_active 86 %, ~13.6 B/clk/SM of reduction payload;
// profiles/r01_v2_ncu_summary.md): every grad_input byte leaves the SM ~36 times.  Here a warp
// owns one (output tile, group) and accumulates that group's grad_input over the tile's input
// window (tile + halo) in shared memory in fp32, then flushes the window once.
//
// Work split inside a warp.  An *entry* is one sampling point of one output pixel of the tile.
//   param phase   lane <-> entry (32 per round): offsets/mask loads, locate() (the same device
//                 function as every other kernel: same integer contract), corner weights, window
//                 offset; parameters go to shared memory.
//   iterations    8 lanes per entry = 4 corners x 2 channel halves (8 channels each); 4 entries per
//                 iteration.  The 4 entries are the same point of 4 pixels that sit TH/2, TW/2
//                 apart, so their 2x2 corner patches almost never overlap; when they do (detected
//                 in the param phase) the read-modify-writes of that iteration are serialised.
//                 Per lane: one LDG.128 gather, an 8-channel partial dot (grad_offset / grad_mask),
//                 and a 32-byte fp32 read-modify-write of the window.  The two 16-byte halves are
//                 visited in opposite order by the upper and lower corner rows, which makes the
//                 8 lanes of an entry hit 8 distinct 16-byte bank groups (conflict-free).
//   combine       lane <-> entry again: corner dots -> grad_mask, grad_offset (plain stores).
//   flush         window -> grad_input accumulator with REDG.F32x4, skipping untouched chunks.
// Corners that fall outside the window (offsets larger than the halo) use global reductions directly.
//
// Reference semantics: dcnv3_col2im_gpu_kernel_* + dcnv3_col2im_bilinear,
// models/ops_dcnv3/src/cuda/dcnv3_im2col_cuda.cuh:82-147, :278-370.
#pragma once

#include "dcnv3_kernels.cuh"

namespace dcnv3 {

struct TileCfg {
    int TH, TW;        // output tile (powers of two)
    int lth, ltw;      // log2(TH/2), log2(TW/2)
    int R;             // halo in input pixels beyond the un-offset sampling footprint
    int oy, ox;        // window origin relative to (tile_h0*stride_h, tile_w0*stride_w)
    int WH, WW;        // window extent in input pixels
    int WWi;           // row pitch (pixels) of the staged input window (padded against bank conflicts)
    int tiles_y, tiles_x;
    int warps;         // warps (= groups) per CTA
    int smem_per_warp; // bytes
};

constexpr int kTileGC = 16;

// 8 channels of storage -> floats, from global (predicated) or shared memory
template <typename T> struct TileIO;
template <> struct TileIO<float> {
    static constexpr int BYTES8 = 32;
    __device__ static __forceinline__ void unpack8(const Words<32> &w, float (&v)[8]) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __uint_as_float(w.w[k]);
    }
    __device__ static __forceinline__ void load8(const char *p, bool pred, float (&v)[8]) {
        unpack8(ldg_pred<32>(p, pred), v);
    }
    // `first` = 0/16: which 16-byte half is read first (bank-conflict-free order, see kernel).
    // The result is in *lane order* (first-read half, then the other): pair it with `gs`.
    static constexpr bool LANE_ORDER = true;
    __device__ static __forceinline__ void lds8(const char *p, int first, float (&v)[8]) {
        const float4 a = *reinterpret_cast<const float4 *>(p + first);
        const float4 b = *reinterpret_cast<const float4 *>(p + (first ^ 16));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
};
template <> struct TileIO<__half> {
    static constexpr int BYTES8 = 16;
    __device__ static __forceinline__ void unpack8(const Words<16> &w, float (&v)[8]) {
        float2 f[4];
        to_pairs<16>(w, f, (const __half *)nullptr);
#pragma unroll
        for (int k = 0; k < 4; ++k) { v[2 * k] = f[k].x; v[2 * k + 1] = f[k].y; }
    }
    __device__ static __forceinline__ void load8(const char *p, bool pred, float (&v)[8]) {
        unpack8(ldg_pred<16>(p, pred), v);
    }
    static constexpr bool LANE_ORDER = false;
    __device__ static __forceinline__ void lds8(const char *p, int, float (&v)[8]) {
        const uint4 r = *reinterpret_cast<const uint4 *>(p);
        Words<16> w; w.w[0] = r.x; w.w[1] = r.y; w.w[2] = r.z; w.w[3] = r.w;
        unpack8(w, v);
    }
};
template <> struct TileIO<__nv_bfloat16> {
    static constexpr int BYTES8 = 16;
    __device__ static __forceinline__ void unpack8(const Words<16> &w, float (&v)[8]) {
        float2 f[4];
        to_pairs<16>(w, f, (const __nv_bfloat16 *)nullptr);
#pragma unroll
        for (int k = 0; k < 4; ++k) { v[2 * k] = f[k].x; v[2 * k + 1] = f[k].y; }
    }
    __device__ static __forceinline__ void load8(const char *p, bool pred, float (&v)[8]) {
        unpack8(ldg_pred<16>(p, pred), v);
    }
    static constexpr bool LANE_ORDER = false;
    __device__ static __forceinline__ void lds8(const char *p, int, float (&v)[8]) {
        const uint4 r = *reinterpret_cast<const uint4 *>(p);
        Words<16> w; w.w[0] = r.x; w.w[1] = r.y; w.w[2] = r.z; w.w[3] = r.w;
        unpack8(w, v);
    }
};

// flag bits stored in the low 6 bits of the window offset (a multiple of 64 bytes)
enum : int { TF_OK1 = 1, TF_OK2 = 2, TF_OK3 = 4, TF_OK4 = 8, TF_INWIN = 16 };

template <typename T, int KP, bool LOGITS>
__global__ void __launch_bounds__(128)
bwd_tile_kernel(const T *__restrict__ in, const T *__restrict__ off, const T *__restrict__ mask,
                const T *__restrict__ gout, float *__restrict__ acc, T *__restrict__ goff,
                T *__restrict__ gmask, const Geo q, const TileCfg tc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int SLAB = kTileGC * (int)sizeof(T);  // bytes of one staged input slab
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // ---- task decode: blockIdx.x -> (n, tile_y, tile_x, group block); warp -> group
    const int gblocks = (q.G + tc.warps - 1) / tc.warps;
    int b = blockIdx.x;
    const int gb = b % gblocks; b /= gblocks;
    const int tx = b % tc.tiles_x; b /= tc.tiles_x;
    const int ty = b % tc.tiles_y;
    const int n = b / tc.tiles_y;
    const int g = gb * tc.warps + warp;
    if (g >= q.G) return;  // whole warp; no block-level barrier is used anywhere

    const int P = KP ? KP : q.P;
    const int KH = KP ? 3 : q.kh;
    const int TP = tc.TH * tc.TW;
    const int h0 = ty * tc.TH, w0 = tx * tc.TW;          // output tile origin
    const int wh0 = h0 * q.sh + tc.oy;                    // window origin in input coordinates
    const int ww0 = w0 * q.sw + tc.ox;
    const int WW = tc.WW, WH = tc.WH, WWi = tc.WWi;

    // ---- per-warp shared memory carve-up
    unsigned char *sm = smem_raw + (size_t)warp * tc.smem_per_warp;
    float *win = reinterpret_cast<float *>(sm);                           // [WH*WW][16] fp32 grad window
    char *inw = reinterpret_cast<char *>(sm) + (size_t)WH * WW * 64;      // [WH][WWi] staged input slabs
    int2 *parA = reinterpret_cast<int2 *>(inw + (size_t)WH * WWi * SLAB); // [32] {gofs | iofs, wofs|flags}
    float4 *parW = reinterpret_cast<float4 *>(parA + 32);                 // [32] corner weights * mask
    float *dts = reinterpret_cast<float *>(parW + 32);                    // [32][8] partial dots
    float *sm_prob = dts + 32 * 8;                                        // [TP][P] (LOGITS only)
    float *sm_gm = sm_prob + (LOGITS ? TP * P : 0);                       // [TP][P] (LOGITS only)

    const size_t img = (size_t)n * q.H * q.W * q.C;
    const char *im_g = reinterpret_cast<const char *>(in + img + g * kTileGC);
    float *acc_g = acc + img + g * kTileGC;
    const int sC = q.C * (int)sizeof(T), sW = q.W * sC;  // global strides (bytes)

    // ---- stage the input window (zeros outside the image: an invalid corner then contributes
    //      exactly 0, cuh:57-75) and zero the grad window
    {
        constexpr int CPS = SLAB / 16;  // 16-byte chunks per slab
        const int nchunk = WH * WW * CPS;
        for (int c = lane; c < nchunk; c += 32) {
            const int wpix = c / CPS, part = c - wpix * CPS;
            const int yy = wpix / WW, xx = wpix - yy * WW;
            const int hy = wh0 + yy, hx = ww0 + xx;
            const bool okp = hy >= 0 && hy < q.H && hx >= 0 && hx < q.W;
            const Words<16> w = ldg_pred<16>(im_g + (hy * q.W + hx) * sC + part * 16, okp);
            *reinterpret_cast<uint4 *>(inw + (yy * WWi + xx) * SLAB + part * 16) =
                make_uint4(w.w[0], w.w[1], w.w[2], w.w[3]);
        }
        float4 *w4 = reinterpret_cast<float4 *>(win);
        const int nz = WH * WW * 4;
        for (int c = lane; c < nz; c += 32) w4[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    // lane roles inside an iteration
    const int j = lane >> 3;         // entry of the iteration = pixel of the quad
    const int k = (lane >> 1) & 3;   // corner
    const int hf = lane & 1;         // channel half
    const int gdelta = ((k & 1) ? sC : 0) + ((k & 2) ? sW : 0) + hf * TileIO<T>::BYTES8;        // global
    const int idelta = ((k & 1) ? SLAB : 0) + ((k & 2) ? WWi * SLAB : 0) + hf * TileIO<T>::BYTES8;  // staged
    const int wdelta = ((k & 1) ? 64 : 0) + ((k & 2) ? WW * 64 : 0) + hf * 32;                  // grad window
    const int c_first = (k >> 1) ? 16 : 0;  // upper corner row visits bytes 0..15 first, lower 16..31
    const int adelta = ((k & 1) ? q.C : 0) + ((k & 2) ? q.W * q.C : 0) + hf * 8;  // fallback (floats)

    const int half_th = tc.TH >> 1, half_tw = tc.TW >> 1;
    const int nquad = half_th * half_tw;
    const int total_entries = nquad * P * 4;
    const int rounds = (total_entries + 31) >> 5;

    // grad_output of (pixel j of a quad, channel half hf): current quad and the next one (prefetched)
    // `gs` is `go` in lane order: the 16-byte half this lane touches first, then the other half.
    float go[8], gs[8], go_nx[8];
    auto load_go = [&](int iq, float (&dst)[8]) {
        const int qy2 = iq >> tc.ltw, qx2 = iq & (half_tw - 1);
        const int ho2 = h0 + qy2 + ((j & 2) ? half_th : 0), wo2 = w0 + qx2 + ((j & 1) ? half_tw : 0);
        const bool v2 = iq < nquad && ho2 < q.Ho && wo2 < q.Wo;
        const T *gp = gout + (((size_t)n * q.Ho + (v2 ? ho2 : 0)) * q.Wo + (v2 ? wo2 : 0)) * q.C +
                      g * kTileGC + hf * 8;
        TileIO<T>::load8(reinterpret_cast<const char *>(gp), v2, dst);
    };
    load_go(0, go_nx);

    // entry E of a round -> (quad, point, pixel of quad); prefetch of its offset pair and mask value
    struct Ent { int p, py, px, ho, wo; bool valid; size_t pg; };
    auto decode = [&](int E) {
        Ent e;
        const int qd = E / (4 * P);
        const int rem = E - qd * 4 * P;
        e.p = rem >> 2;
        const int ej = rem & 3;
        e.py = (qd >> tc.ltw) + ((ej & 2) ? half_th : 0);
        e.px = (qd & (half_tw - 1)) + ((ej & 1) ? half_tw : 0);
        e.ho = h0 + e.py; e.wo = w0 + e.px;
        e.valid = E < total_entries && e.ho < q.Ho && e.wo < q.Wo;
        e.pg = (((size_t)n * q.Ho + (e.valid ? e.ho : 0)) * q.Wo + (e.valid ? e.wo : 0)) * q.G + g;
        return e;
    };
    Ent en = decode(lane);
    float2 o_nx = make_float2(0.f, 0.f);
    float m_nx = 0.f;
    if (en.valid) {
        o_nx = load_offset_pair(off + en.pg * P * 2 + 2 * en.p);
        m_nx = to_math(mask[en.pg * P + en.p]);
    }

    int ip = 0, iq = 0;  // (point, quad) of the next iteration
    __syncwarp();

    for (int r = 0; r < rounds; ++r) {
        // ================= param phase: lane <-> entry E =================
        const Ent e0 = en;
        const float2 o = o_nx;
        float m = m_nx;
        {   // prefetch the next round's global loads; they land while this round iterates
            en = decode(((r + 1) << 5) + lane);
            if (en.valid) {
                o_nx = load_offset_pair(off + en.pg * P * 2 + 2 * en.p);
                m_nx = to_math(mask[en.pg * P + en.p]);
            }
        }
        float hh = 0.f, lh = 0.f, hw = 0.f, lw = 0.f;
        float w1 = 0.f, w2 = 0.f, w3 = 0.f, w4 = 0.f;
        int flags = 0, wy = -100000, wx = -100000, gofs = 0, wofs = 0;
        if (e0.valid) {
            if (LOGITS) {
                float mx, inv;
                softmax_stats<T, KP>(mask + e0.pg * P, P, mx, inv);
                m = expf(m - mx) * inv;
            }
            float p0h_, p0w_;
            window_origin<float>(q, e0.ho, e0.wo, p0h_, p0w_);
            const int pi = e0.p / KH, pj = e0.p - pi * KH;  // p = i*kh + j
            Point<float> t;
            locate<float>(q, p0h_, p0w_, pi, pj, o.x, o.y, t);
            hh = t.hh; lh = t.lh; hw = t.hw; lw = t.lw;
            w1 = hh * hw; w2 = hh * lw; w3 = lh * hw; w4 = lh * lw;
            const int yy = t.h_low - wh0, xx = t.w_low - ww0;
            const bool inwin = t.inside && yy >= 0 && yy + 1 < WH && xx >= 0 && xx + 1 < WW;
            if (inwin) {
                wy = yy; wx = xx;
                wofs = (yy * WW + xx) * 64;
                gofs = (yy * WWi + xx) * SLAB;          // offset into the staged input window
            } else {
                gofs = (t.h_low * q.W + t.w_low) * sC;  // offset into the image (fallback path)
            }
            flags = (t.ok1 ? TF_OK1 : 0) | (t.ok2 ? TF_OK2 : 0) | (t.ok3 ? TF_OK3 : 0) |
                    (t.ok4 ? TF_OK4 : 0) | (inwin ? TF_INWIN : 0);
        } else {
            m = 0.f;
        }
        parA[lane] = make_int2(gofs, wofs | flags);
        parW[lane] = make_float4(w1 * m, w2 * m, w3 * m, w4 * m);

        // conflicts among the 4 entries of one iteration (aligned lane quads): overlapping 2x2 patches
        bool conf = false;
#pragma unroll
        for (int d = 1; d < 4; ++d) {
            const int oy = __shfl_xor_sync(0xffffffffu, wy, d);
            const int ox = __shfl_xor_sync(0xffffffffu, wx, d);
            conf = conf || (abs(oy - wy) <= 1 && abs(ox - wx) <= 1);
        }
        const unsigned conf_lanes = __ballot_sync(0xffffffffu, conf && (flags & TF_INWIN));
        __syncwarp();

        // ================= iterations: 4 entries each =================
        const int n_it = min(8, (total_entries - (r << 5) + 3) >> 2);
#pragma unroll 1
        for (int it = 0; it < n_it; ++it) {
            if (ip == 0) {  // new quad (warp-uniform): rotate in the prefetched grad_output, fetch the next
#pragma unroll
                for (int c = 0; c < 8; ++c) go[c] = go_nx[c];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    gs[c] = c_first ? go[4 + c] : go[c];
                    gs[4 + c] = c_first ? go[c] : go[4 + c];
                }
                load_go(iq + 1, go_nx);
            }
            if (++ip == P) { ip = 0; ++iq; }

            const int e = (it << 2) + j;
            const int2 pa = parA[e];
            const float wk = reinterpret_cast<const float *>(parW)[(e << 2) + k];
            const bool ok = (pa.y >> k) & 1;
            const bool inwin = (pa.y & TF_INWIN) != 0;

            // this corner's 8 channels: staged window (zeros outside the image) or global fallback
            float v[8];
            float2 a = make_float2(0.f, 0.f);
            if (inwin) {
                TileIO<T>::lds8(inw + pa.x + idelta, c_first, v);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float2 gg = TileIO<T>::LANE_ORDER ? make_float2(gs[2 * c], gs[2 * c + 1])
                                                            : make_float2(go[2 * c], go[2 * c + 1]);
                    a = __ffma2_rn(gg, make_float2(v[2 * c], v[2 * c + 1]), a);
                }
            } else {
                TileIO<T>::load8(im_g + pa.x + gdelta, ok, v);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    a = __ffma2_rn(make_float2(go[2 * c], go[2 * c + 1]), make_float2(v[2 * c], v[2 * c + 1]), a);
            }
            dts[(e << 3) + (k << 1) + hf] = a.x + a.y;

            // grad_input: window read-modify-write (or direct reduction outside the window)
            const bool conflict = (conf_lanes >> (it << 2)) & 0xfu;  // warp-uniform
            char *wp = reinterpret_cast<char *>(win) + (pa.y & ~63) + wdelta;
            if (ok && !inwin) {
                float *ap = acc_g + pa.x / (int)sizeof(T) + adelta;
                red_add_v4_f32(ap, wk * go[0], wk * go[1], wk * go[2], wk * go[3]);
                red_add_v4_f32(ap + 4, wk * go[4], wk * go[5], wk * go[6], wk * go[7]);
            }
            if (!conflict) {
                if (inwin) {
                    float4 *c0 = reinterpret_cast<float4 *>(wp + c_first);
                    float4 *c1 = reinterpret_cast<float4 *>(wp + (c_first ^ 16));
                    float4 x = *c0, y = *c1;
                    x.x = fmaf(wk, gs[0], x.x); x.y = fmaf(wk, gs[1], x.y);
                    x.z = fmaf(wk, gs[2], x.z); x.w = fmaf(wk, gs[3], x.w);
                    y.x = fmaf(wk, gs[4], y.x); y.y = fmaf(wk, gs[5], y.y);
                    y.z = fmaf(wk, gs[6], y.z); y.w = fmaf(wk, gs[7], y.w);
                    *c0 = x; *c1 = y;
                }
            } else {
                for (int jj = 0; jj < 4; ++jj) {  // rare: one entry at a time
                    if (inwin && j == jj) {
                        float4 *c0 = reinterpret_cast<float4 *>(wp);
                        float4 x = c0[0], y = c0[1];
                        x.x = fmaf(wk, go[0], x.x); x.y = fmaf(wk, go[1], x.y);
                        x.z = fmaf(wk, go[2], x.z); x.w = fmaf(wk, go[3], x.w);
                        y.x = fmaf(wk, go[4], y.x); y.y = fmaf(wk, go[5], y.y);
                        y.z = fmaf(wk, go[6], y.z); y.w = fmaf(wk, go[7], y.w);
                        c0[0] = x; c0[1] = y;
                    }
                    __syncwarp();
                }
            }
            __syncwarp();  // order this iteration's window stores before the next one's loads
        }

        // ================= combine: lane <-> entry E again =================
        if (e0.valid) {
            const float4 da = reinterpret_cast<const float4 *>(dts)[lane * 2];
            const float4 db = reinterpret_cast<const float4 *>(dts)[lane * 2 + 1];
            const float d1 = da.x + da.y, d2 = da.z + da.w, d3 = db.x + db.y, d4 = db.z + db.w;
            const float s_m = w1 * d1 + w2 * d2 + w3 * d3 + w4 * d4;          // cuh:144
            const float s_w = hh * (d2 - d1) + lh * (d4 - d3);                // cuh:114-139,145
            const float s_h = hw * (d3 - d1) + lw * (d4 - d2);                // cuh:114-139,146
            const float sm_ = q.scale * m;
            store_pair<T>(goff + e0.pg * P * 2 + 2 * e0.p, sm_ * s_w, sm_ * s_h);
            if (LOGITS) {
                sm_prob[(e0.py * tc.TW + e0.px) * P + e0.p] = m;
                sm_gm[(e0.py * tc.TW + e0.px) * P + e0.p] = s_m;
            } else {
                gmask[e0.pg * P + e0.p] = from_math<T>(s_m);
            }
        }
        __syncwarp();  // params / dots are rewritten by the next round
    }

    if (LOGITS) {  // softmax Jacobian per (pixel, g): dl_p = m_p (gm_p - sum_q m_q gm_q)
        for (int t = lane; t < TP; t += 32) {
            const int py = t / tc.TW, px = t - py * tc.TW;
            const int ho = h0 + py, wo = w0 + px;
            if (ho >= q.Ho || wo >= q.Wo) continue;
            const size_t pg = (((size_t)n * q.Ho + ho) * q.Wo + wo) * q.G + g;
            float dot = 0.f;
            for (int pp = 0; pp < P; ++pp) dot = fmaf(sm_prob[t * P + pp], sm_gm[t * P + pp], dot);
            for (int pp = 0; pp < P; ++pp)
                gmask[pg * P + pp] = from_math<T>(sm_prob[t * P + pp] * (sm_gm[t * P + pp] - dot));
        }
    }

    // ================= flush the window =================
    {
        const float4 *w4 = reinterpret_cast<const float4 *>(win);
        const int nchunk = WH * WW * 4;
        for (int c = lane; c < nchunk; c += 32) {
            const int wpix = c >> 2, part = c & 3;
            const int yy = wpix / WW, xx = wpix - yy * WW;
            const int hy = wh0 + yy, hx = ww0 + xx;
            if (hy < 0 || hy >= q.H || hx < 0 || hx >= q.W) continue;
            const float4 v = w4[c];
            if (v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f) continue;
            red_add_v4_f32(acc_g + ((size_t)hy * q.W + hx) * q.C + part * 4, v.x, v.y, v.z, v.w);
        }
    }
}

}  // namespace dcnv3

// dcnv3_b200 — privatised backward for group_channels == 16 (experimental, DCNV3_B200_BWD=tile).
//
// Why: with global reductions the backward is bound by the SM -> crossbar request port
// (l1tex__m_l1tex2xbar_req_cycles_active 86 %, ~13.6 B/clk/SM of reduction payload;
// profiles/r01_v2_ncu_summary.md): every grad_input byte leaves the SM ~36 times.  Here a warp owns
// one (8x8 output tile, group) and accumulates that group's grad_input over the tile's input
// window (tile + halo) in shared memory in fp32, then flushes the window once.
//
// Mapping (same as bwd_vec_kernel, so the instruction stream stays lean): 2 lanes per *entry*
// (= one sampling point of one output pixel), 8 channels each; a step handles point p of 16 pixels
// of the tile (one of four interleaved 4x4 sub-lattices), 36 steps per tile.  Each lane gathers its
// four corners with LDG.128 (L1: the windows leave it ~40 KB), forms the four corner dots
// (grad_mask / grad_offset, reduced over the lane pair with one shuffle), and adds w_k*m*go into
// the four corner slabs of the window, one corner after the other.  Because corners are visited
// in lockstep, two entries can only collide when they sit in the SAME cell (h_low, w_low); that is
// detected with a one-word claim per entry in shared memory and the losers retry in a second pass.
// Corners outside the window (offsets beyond the halo) use global vector reductions.
//
// (A first version with 8 lanes per entry — 4 corners x 2 halves, parameters staged through shared
// memory — was correct but needed 37 warp-instructions per entry against 16 here)
//
// Reference semantics: dcnv3_col2im_gpu_kernel_* + dcnv3_col2im_bilinear,
// models/ops_dcnv3/src/cuda/dcnv3_im2col_cuda.cuh:82-147, :278-370.
#pragma once

#include "dcnv3_kernels.cuh"

namespace dcnv3 {

struct TileCfg {
    int R;             // halo in input pixels beyond the un-offset sampling footprint
    int oy, ox;        // window origin relative to (tile_h0*stride_h, tile_w0*stride_w)
    int WH, WW;        // window extent in input pixels
    int tiles_y, tiles_x;
    int warps;         // warps (= groups) per CTA
    int smem_per_warp; // bytes
};

constexpr int kTileGC = 16;
constexpr int kTileT = 8;  // output tile edge

// window slab read-modify-write: this lane's 8 channels (32 bytes) of one corner.
// `first` (0 / 16) swaps the order of the two 16-byte halves for odd entries so that the lanes of
// a quarter-warp spread over all eight 16-byte bank groups; `gs` is grad_output in that order.
__device__ __forceinline__ void window_add(char *slab, int first, const float (&gs)[8], float w) {
    float4 *c0 = reinterpret_cast<float4 *>(slab + first);
    float4 *c1 = reinterpret_cast<float4 *>(slab + (first ^ 16));
    float4 x = *c0, y = *c1;
    x.x = fmaf(w, gs[0], x.x); x.y = fmaf(w, gs[1], x.y); x.z = fmaf(w, gs[2], x.z); x.w = fmaf(w, gs[3], x.w);
    y.x = fmaf(w, gs[4], y.x); y.y = fmaf(w, gs[5], y.y); y.z = fmaf(w, gs[6], y.z); y.w = fmaf(w, gs[7], y.w);
    *c0 = x; *c1 = y;
}

template <typename T, int KP, bool LOGITS>
__global__ void __launch_bounds__(128)
bwd_tile_kernel(const T *__restrict__ in, const T *__restrict__ off, const T *__restrict__ mask,
                const T *__restrict__ gout, float *__restrict__ acc, T *__restrict__ goff,
                T *__restrict__ gmask, const Geo q, const TileCfg tc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int BPL = 8 * (int)sizeof(T);  // bytes of this lane's 8 channels in storage
    constexpr int NP = 4;                    // float2 pairs
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
    const int kw = KP ? 3 : q.kw, kh = KP ? 3 : q.kh;
    const int h0 = ty * kTileT, w0 = tx * kTileT;  // output tile origin
    const int wh0 = h0 * q.sh + tc.oy;              // window origin in input coordinates
    const int ww0 = w0 * q.sw + tc.ox;
    const int WW = tc.WW, WH = tc.WH;

    // ---- per-warp shared memory: fp32 window [WH*WW][16] and the claim word per cell [WH*WW]
    unsigned char *sm = smem_raw + (size_t)warp * tc.smem_per_warp;
    char *win = reinterpret_cast<char *>(sm);
    int *claim = reinterpret_cast<int *>(sm + (size_t)WH * WW * 64);
    {
        float4 *w4 = reinterpret_cast<float4 *>(win);
        const int nz = WH * WW * 4;
        for (int c = lane; c < nz; c += 32) w4[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    const size_t img = (size_t)n * q.H * q.W * q.C;
    const int e = lane >> 1, hf = lane & 1;  // entry of the step, channel half
    const char *im = reinterpret_cast<const char *>(in + img + g * kTileGC + hf * 8);
    float *acc_g = acc + img + g * kTileGC;
    const int sC = q.C * (int)sizeof(T), sW = q.W * sC;  // gather strides (bytes)
    const int first = (e & 1) ? 16 : 0;
    int stepid = lane >> 1;  // unique claim id per (step, entry): advanced by 16 every step
    __syncwarp();

    for (int sub = 0; sub < 4; ++sub) {
        // pixel of this entry in sub-lattice `sub`: (2i + a, 2j + b)
        const int py = ((e >> 2) << 1) + (sub >> 1), px = ((e & 3) << 1) + (sub & 1);
        const int ho = h0 + py, wo = w0 + px;
        const bool valid = ho < q.Ho && wo < q.Wo;
        const size_t pix = ((size_t)n * q.Ho + (valid ? ho : 0)) * q.Wo + (valid ? wo : 0);
        const size_t pg = pix * q.G + g;
        const T *po = off + pg * P * 2;
        const T *pm = mask + pg * P;
        T *d_o = goff + pg * P * 2;
        T *d_m = gmask + pg * P;
        const bool writer = valid && hf == 0;

        float2 gp[NP];  // grad_output, this lane's 8 channels, natural order (for the dots)
        float gs[8];    // the same in window-access order
        float gr[2][4]; // channels of the two fallback reduction chunks (sector-filling layout of
                        // bwd_vec_kernel: instruction jj of lane hf carries 16-byte chunk 2jj + hf)
        {
            const T *gop = gout + pix * q.C + g * kTileGC;
            const Words<BPL> gw = ldg_pred<BPL>(gop + hf * 8, valid);
            to_pairs<BPL>(gw, gp, (const T *)nullptr);
            const float gn[8] = {gp[0].x, gp[0].y, gp[1].x, gp[1].y, gp[2].x, gp[2].y, gp[3].x, gp[3].y};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                gs[c] = first ? gn[4 + c] : gn[c];
                gs[4 + c] = first ? gn[c] : gn[4 + c];
            }
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                if (valid) load_go_chunk<T, 4>(gop + (2 * jj + hf) * 4, gr[jj]);
                else gr[jj][0] = gr[jj][1] = gr[jj][2] = gr[jj][3] = 0.f;
            }
        }
        float mx = 0.f, inv = 1.f;
        if (LOGITS && valid) softmax_stats<T, KP>(pm, P, mx, inv);
        float prob[KP ? KP : kMaxSoftmaxP], gm[KP ? KP : kMaxSoftmaxP];  // only live when LOGITS

        float p0h_, p0w_;
        window_origin<float>(q, ho, wo, p0h_, p0w_);

        int p = 0;
#pragma unroll
        for (int i = 0; i < kw; ++i) {
#pragma unroll
            for (int j = 0; j < kh; ++j, ++p) {
                float2 o = make_float2(0.f, 0.f);
                float m = 0.f;
                if (valid) {
                    o = load_offset_pair(po + 2 * p);
                    m = to_math(pm[p]);
                    if (LOGITS) m = expf(m - mx) * inv;
                }
                Point<float> t;
                locate<float>(q, p0h_, p0w_, i, j, o.x, o.y, t);
                const bool live = valid && t.inside;
                const bool k1 = live && t.ok1, k2 = live && t.ok2, k3 = live && t.ok3, k4 = live && t.ok4;

                // ---- gather + corner dots (as bwd_vec_kernel)
                const int e0 = (t.h_low * q.W + t.w_low) * sC;
                const char *r1 = im + e0;
                const Words<BPL> c1 = ldg_pred<BPL>(r1, k1);
                const Words<BPL> c2 = ldg_pred<BPL>(r1 + sC, k2);
                const Words<BPL> c3 = ldg_pred<BPL>(r1 + sW, k3);
                const Words<BPL> c4 = ldg_pred<BPL>(r1 + sW + sC, k4);
                float d[4];
                {
                    float2 v[NP];
                    float2 a;
#define DCNV3_DOT(CW, K)                                                          \
    to_pairs<BPL>(CW, v, (const T *)nullptr);                                     \
    a = make_float2(0.f, 0.f);                                                    \
    _Pragma("unroll") for (int c = 0; c < NP; ++c) a = __ffma2_rn(gp[c], v[c], a); \
    d[K] = a.x + a.y;
                    DCNV3_DOT(c1, 0)
                    DCNV3_DOT(c2, 1)
                    DCNV3_DOT(c3, 2)
                    DCNV3_DOT(c4, 3)
#undef DCNV3_DOT
                }
                const float w1 = t.hh * t.hw, w2 = t.hh * t.lw, w3 = t.lh * t.hw, w4 = t.lh * t.lw;
                float s_m = w1 * d[0] + w2 * d[1] + w3 * d[2] + w4 * d[3];
                float s_w = t.hh * (d[1] - d[0]) + t.lh * (d[3] - d[2]);
                float s_h = t.hw * (d[2] - d[0]) + t.lw * (d[3] - d[1]);
                s_m += shfl_xor(s_m, 1);
                s_w += shfl_xor(s_w, 1);
                s_h += shfl_xor(s_h, 1);
                const float sm_ = q.scale * m;  // cuh:145-146
                if (writer) store_pair<T>(d_o + 2 * p, sm_ * s_w, sm_ * s_h);
                if (LOGITS) {
                    prob[p] = m;
                    gm[p] = s_m;
                } else if (writer) {
                    d_m[p] = from_math<T>(s_m);  // cuh:144
                }

                // ---- grad_input: window (private, conflict-checked) or global reduction
                const float a1 = w1 * m, a2 = w2 * m, a3 = w3 * m, a4 = w4 * m;
                const int yy = t.h_low - wh0, xx = t.w_low - ww0;
                const bool inwin = live && yy >= 0 && yy + 1 < WH && xx >= 0 && xx + 1 < WW;
                const int cell = inwin ? yy * WW + xx : 0;
                if (live && !inwin) {  // rare: beyond the halo
                    float *g1 = acc_g + (t.h_low * q.W + t.w_low) * q.C;
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) {
                        float *gq = g1 + (2 * jj + hf) * 4;
                        red_add_v4_f32(gq, a1 * gr[jj][0], a1 * gr[jj][1], a1 * gr[jj][2], a1 * gr[jj][3], k1);
                        red_add_v4_f32(gq + q.C, a2 * gr[jj][0], a2 * gr[jj][1], a2 * gr[jj][2], a2 * gr[jj][3], k2);
                        red_add_v4_f32(gq + q.W * q.C, a3 * gr[jj][0], a3 * gr[jj][1], a3 * gr[jj][2], a3 * gr[jj][3], k3);
                        red_add_v4_f32(gq + q.W * q.C + q.C, a4 * gr[jj][0], a4 * gr[jj][1], a4 * gr[jj][2], a4 * gr[jj][3], k4);
                    }
                }
                char *s1 = win + cell * 64 + hf * 32;
                const int rowb = WW * 64;
                bool pending = inwin;
                stepid += 16;
                do {  // one pass unless two entries of this step share a cell
                    if (pending) claim[cell] = stepid;
                    __syncwarp();
                    const bool mine = pending && claim[cell] == stepid;
                    // corners in lockstep: different corners of overlapping patches never meet in one
                    // instruction; the barriers order each corner's stores before the next one's loads
                    if (mine && k1) window_add(s1, first, gs, a1);
                    __syncwarp();
                    if (mine && k2) window_add(s1 + 64, first, gs, a2);
                    __syncwarp();
                    if (mine && k3) window_add(s1 + rowb, first, gs, a3);
                    __syncwarp();
                    if (mine && k4) window_add(s1 + rowb + 64, first, gs, a4);
                    __syncwarp();
                    pending = pending && !mine;
                } while (__any_sync(0xffffffffu, pending));
            }
        }
        if (LOGITS && writer) {  // softmax Jacobian: dl_p = m_p (gm_p - sum_q m_q gm_q)
            float dot = 0.f;
            for (int kk = 0; kk < P; ++kk) dot = fmaf(prob[kk], gm[kk], dot);
            for (int kk = 0; kk < P; ++kk) d_m[kk] = from_math<T>(prob[kk] * (gm[kk] - dot));
        }
    }

    // ================= flush the window =================
    __syncwarp();
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

/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement, in plain C, of the DCNv3 deformable-sampling core as the
 * reference's CUDA kernels compute it (pixel-space location arithmetic).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference leg may load this library, and only as the checker.  The product
 * path (yolo_dual_b200/csrc) never links or calls it.
 *
 * Reference lines followed (all under /root/reference/models/ops_dcnv3/):
 *   output size ............ src/cuda/dcnv3_cuda.cu:40-45
 *   index decode, p0 ....... src/cuda/dcnv3_im2col_cuda.cuh:226-238
 *   top-left, location ..... src/cuda/dcnv3_im2col_cuda.cuh:249-260
 *   inside gate ............ src/cuda/dcnv3_im2col_cuda.cuh:262-263
 *   corners / validity ..... src/cuda/dcnv3_im2col_cuda.cuh:39-42,56-75
 *   bilinear value ......... src/cuda/dcnv3_im2col_cuda.cuh:76-79
 *   backward per point ..... src/cuda/dcnv3_im2col_cuda.cuh:106-146
 *   channel reduction ...... src/cuda/dcnv3_im2col_cuda.cuh:345-360 (serial, c = 0..gc-1)
 *   point order p = i_w*Kh + j_h, offset (x,y) interleave: :253-261
 *
 * Parity pin: tests/test_oracle_golden.py checks this file against golden
 * vectors produced by the reference's own dcnv3_core_pytorch
 * (tests/golden/make_golden.py, run in the build container where
 * /root/reference is mounted).
 *
 * Arithmetic contract for the integer outputs (h_low, w_low, bounds bits):
 * every location operation is a separately rounded IEEE operation in the
 * op-math type (float for f32/f16/bf16 storage, double for f64), in the
 * reference's source order.  Build with -ffp-contract=off (see Makefile) so
 * gcc never fuses them; the CUDA product uses __fmul_rn/__fadd_rn for the same
 * reason.
 *
 * Build: make -C oracle   ->  oracle/libdcnv3_oracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int N, H, W, G, gc;
    int kh, kw, sh, sw, ph, pw, dh, dw;
    int Ho, Wo, P;
} geo_t;

static int make_geo(geo_t *q, int N, int H, int W, int G, int gc, int kh, int kw,
                    int sh, int sw, int ph, int pw, int dh, int dw) {
    if (N < 0 || H <= 0 || W <= 0 || G <= 0 || gc <= 0 || kh <= 0 || kw <= 0 ||
        sh <= 0 || sw <= 0 || ph < 0 || pw < 0 || dh <= 0 || dw <= 0)
        return -1;
    q->N = N; q->H = H; q->W = W; q->G = G; q->gc = gc;
    q->kh = kh; q->kw = kw; q->sh = sh; q->sw = sw;
    q->ph = ph; q->pw = pw; q->dh = dh; q->dw = dw;
    /* dcnv3_cuda.cu:40-45 */
    q->Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) / sh + 1;
    q->Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) / sw + 1;
    q->P = kh * kw;
    if (q->Ho <= 0 || q->Wo <= 0) return -2;
    return 0;
}

int dcnv3_oracle_output_hw(int H, int W, int kh, int kw, int sh, int sw, int ph,
                           int pw, int dh, int dw, int *Ho, int *Wo) {
    geo_t q;
    int rc = make_geo(&q, 1, H, W, 1, 1, kh, kw, sh, sw, ph, pw, dh, dw);
    if (rc) return rc;
    *Ho = q.Ho; *Wo = q.Wo;
    return 0;
}

int dcnv3_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void dcnv3_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* bounds byte layout (shared with the CUDA debug kernel, include/dcnv3_b200.h):
 *   bit0 inside gate, bit1 corner(h_low,w_low) valid, bit2 (h_low,w_high),
 *   bit3 (h_high,w_low), bit4 (h_high,w_high).  Corner bits are reported only
 *   when the gate is open (a closed gate never evaluates them in the reference). */

#define REAL float
#define FLOOR floorf
#define SFX(name) name##_f32
#include "dcnv3_oracle_body.inc"
#undef REAL
#undef FLOOR
#undef SFX

#define REAL double
#define FLOOR floor
#define SFX(name) name##_f64
#include "dcnv3_oracle_body.inc"
#undef REAL
#undef FLOOR
#undef SFX

/*
 * segloss_b200 — C-ABI of the fused segmentation loss of the C3-DCN seg trainers.
 *
 * Replaces, on CUDA tensors, the chain of ~25 elementwise / reduction launches behind the reference's
 *   SegmentationLoss.forward      unet-lite/yolo5-seg/seg_diceloss_yolov5.py:712-735
 *   SegmentationLoss._one_hot_encode / _dice_loss                         :737-750
 * (same class in unet-lite/yolo8-seg/seg_diceloss_yolov8.py):
 *   total = CE(pred, target; class weights) + 0.5 * (1 - mean_{n,c} dice[n,c]),
 *   dice[n,c] = (2 I + eps) / (P + O + eps),  q = softmax_c(pred),
 *   I = sum_pix w_c q_c [t = c],  P = sum_pix w_c q_c,  O = sum_pix [t = c],  eps = 1e-6,
 *   CE = sum_pix w_t (-log q_t) / sum_pix w_t         (label_smoothing = 0 only).
 *
 * `pred` is [N, C, h, w] float32, NCHW-contiguous; `target` is [N, h*scale, w*scale] int64.  With scale > 1 every
 * pred pixel stands for the scale x scale block of full-resolution pixels a nearest-neighbour Upsample would
 * replicate it into (the seg models end in Upsample -> 1x1 Conv -> Softmax, all pointwise after the Upsample),
 * so the replicated map is never materialised; grad_pred is the block sum of the full-resolution gradients.
 * Labels outside [0, C) contribute nothing.  C <= SEGLOSS_B200_MAX_CLASSES.
 *
 * The caller allocates everything and passes the CUDA stream; no entry point synchronises.  Return: 0 ok,
 * < 0 argument error, > 0 cudaError_t; segloss_b200_last_error() has the message.
 */
#ifndef SEGLOSS_B200_H_
#define SEGLOSS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEGLOSS_B200_VERSION 100
#define SEGLOSS_B200_MAX_CLASSES 16

int segloss_b200_version(void);
const char* segloss_b200_last_error(void);

/* stats (double, zeroed by this call): [N][3][C] = I, P, O per (image, class), then [2] = CE numerator, CE
 * denominator.  The scalar loss is a handful of operations on this 4.6 KB table (done by the caller). */
int segloss_b200_forward(const float* pred, const int64_t* target, const float* class_weights, double* stats,
                         int N, int C, int h, int w, int scale, void* cuda_stream);

/* coef (float): [N][2][C] = A, B with dTotal/dq_c(pixel) = A[n][c] [t = c] + B[n][c] (the Dice part, the upstream
 * gradient and the factor 0.5 folded in), then [1] = upstream gradient / CE denominator.  grad_pred like pred. */
int segloss_b200_backward(const float* pred, const int64_t* target, const float* class_weights, const float* coef,
                          float* grad_pred, int N, int C, int h, int w, int scale, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif

/*
 * bnact_b200 — C-ABI of the fused training-mode BatchNorm2d + SiLU of the reference's `Conv` block.
 *
 * The block `act(bn(conv(x)))` appears 46 times in the C3-DCN seg model and in front of DCNv3's offset / mask heads
 * (models/ops_dcnv3/.../modules/dcnv3.py:26-40 `Conv`, :88 `dw_conv`; models/common.py Conv).  PyTorch runs it as
 * collect-statistics + transform + SiLU forward and SiLU-backward + reduce + elementwise backward: 5 + 8 passes over the
 * activation.  Here: 3 + 5 passes (statistics, apply; reduce, apply), SiLU recomputed instead of stored.
 *
 * Tensors: x, z, gz, dx are [M, C] row-major = NHWC-contiguous ("channels_last") activations with M = N*H*W, all of
 * one dtype (0 float32, 1 float16, 2 bfloat16); C must be 2^k vectors of 16 bytes with at most 256 vectors
 * (bnact_b200_supported).  gamma, beta, running_*, save, dgamma, dbeta are float32.  Statistics are accumulated in
 * float32 around a per-channel pivot (row 0) and combined in double; biased variance normalises, the unbiased one
 * updates running_var (as torch.nn.BatchNorm2d).  act: 0 identity, 1 SiLU.
 *
 * The caller allocates everything (`partial` = bnact_b200_partial_floats() floats of scratch) and passes the CUDA
 * stream; no entry point synchronises.  Return: 0 ok, < 0 argument error, > 0 cudaError_t.
 */
#ifndef BNACT_B200_H_
#define BNACT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BNACT_B200_VERSION 100

int bnact_b200_version(void);
const char* bnact_b200_last_error(void);
int bnact_b200_supported(int dtype, int C);
size_t bnact_b200_partial_floats(int dtype, int64_t M, int C);

/* save: [4][C] = mean, invstd, scale (= gamma*invstd), shift (= beta - mean*scale); running_* may be NULL */
int bnact_b200_forward(const void* x, void* z, const float* gamma, const float* beta, float* running_mean,
                       float* running_var, float* save, float* partial, int dtype, int64_t M, int C, float eps,
                       float momentum, int act, void* cuda_stream);

/* bnact_b200_forward writing z at a row pitch of z_pitch elements (a channel slice of a wider NHWC tensor). */
int bnact_b200_forward_pitched(const void* x, void* z, const float* gamma, const float* beta, float* running_mean,
                               float* running_var, float* save, float* partial, int dtype, int64_t M, int C, float eps,
                               float momentum, int act, int64_t z_pitch, void* cuda_stream);

/* Inference: z = act((x - running_mean) * gamma / sqrt(running_var + eps) + beta) in one pass (PyTorch: transform +
 * SiLU = two passes and a tiny invstd kernel).  The four per-channel vectors are float32 (params_in_dtype = 0) or in the
 * activation's own 16-bit dtype (params_in_dtype = 1: a model cast with .half() / .bfloat16()).  No autograd side. */
int bnact_b200_eval(const void* x, void* z, const void* gamma, const void* beta, const void* running_mean,
                    const void* running_var, int dtype, int params_in_dtype, int64_t M, int C, float eps, int act,
                    void* cuda_stream);

/* Same, writing z at a row pitch of z_pitch elements: straight into a channel slice of a wider NHWC tensor (the buffer a
 * torch.cat would otherwise fill with a copy). */
int bnact_b200_eval_pitched(const void* x, void* z, const void* gamma, const void* beta, const void* running_mean,
                            const void* running_var, int dtype, int params_in_dtype, int64_t M, int C, float eps, int act,
                            int64_t z_pitch, void* cuda_stream);

/* coef scratch: [2][C]; dgamma, dbeta: [C] */
int bnact_b200_backward(const void* x, const void* gz, void* dx, const float* gamma, const float* beta,
                        const float* save, float* dgamma, float* dbeta, float* coef, float* partial, int dtype,
                        int64_t M, int C, int act, void* cuda_stream);

/* Same, with gz read at a row pitch of gz_pitch elements (>= C, a whole number of 16-byte vectors): the gradient of a
 * channel slice of a wider NHWC tensor — what torch.cat's backward hands to each of its inputs — without a copy. */
int bnact_b200_backward_pitched(const void* x, const void* gz, void* dx, const float* gamma, const float* beta,
                                const float* save, float* dgamma, float* dbeta, float* coef, float* partial, int dtype,
                                int64_t M, int C, int act, int64_t gz_pitch, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif

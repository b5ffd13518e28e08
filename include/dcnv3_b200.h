/*
 * dcnv3_b200 — C-ABI of the B200-native DCNv3 deformable-sampling core.
 *
 * This is the drop-in boundary for the one native component of the reference
 * (Z1HaoC/YOLO-Dual, models/ops_dcnv3): the pybind module `DCNv3` with
 *   dcnv3_forward (...) -> Tensor          src/cuda/dcnv3_cuda.h:15-21   (impl dcnv3_cuda.cu:21-85)
 *   dcnv3_backward(...) -> [Tensor x3]     src/cuda/dcnv3_cuda.h:23-31   (impl dcnv3_cuda.cu:87-173)
 * called from functions/dcnv3_func.py:39-43 and :54-58.  The pybind entry file
 * (src/vision.cpp) is missing from the reference tree; the signatures above are
 * what it exported.
 *
 * Differences from the reference boundary, on purpose:
 *   - plain pointers and sizes, no torch/ATen types: the library links only
 *     the CUDA runtime and is loaded with ctypes (or cgo/JNI/anything);
 *   - the CALLER allocates every output (reference: at::zeros / zeros_like
 *     inside, dcnv3_cuda.cu:55-57,131-133) and passes the CUDA stream
 *     (reference: at::cuda::getCurrentCUDAStream(), dcnv3_cuda.cu:72,150);
 *   - errors are returned, never printf'd and swallowed (reference:
 *     dcnv3_im2col_cuda.cuh:864-867,1041-1044);
 *   - im2col_step is not a parameter: the whole batch is one launch (the
 *     Python wrapper keeps the positional argument for API parity);
 *   - bf16 is accepted (the reference dispatches double/float/half only,
 *     dcnv3_cuda.cu:69,147);
 *   - optional fused softmax over the P sampling points (`mask_is_logits`);
 *     the reference does F.softmax in Python (modules/dcnv3.py:122-123).
 *
 * Tensors are contiguous, channel-last, all of one dtype:
 *   input        [N, H,  W,  G*gc]
 *   offset       [N, Ho, Wo, G*P*2]   (x, y) interleaved per (g, p), p = i_w*kh + j_h
 *   mask         [N, Ho, Wo, G*P]     probabilities, or logits if mask_is_logits
 *   output       [N, Ho, Wo, G*gc]
 * with P = kh*kw, Ho = (H + 2*ph - (dh*(kh-1)+1))/sh + 1, Wo likewise
 * (dcnv3_cuda.cu:40-45).
 *
 * There is no CPU implementation (as in the reference, src/cpu/dcnv3_cpu.cpp:25,36)
 * and no fallback of any kind: every entry point needs a CUDA device of
 * compute capability 10.0 (sm_100a).
 *
 * Return convention: 0 ok; < 0 argument error (DCNV3_B200_E*); > 0 a
 * cudaError_t from the launch.  dcnv3_b200_last_error() gives the message for
 * the calling thread.  No entry point synchronises the host with the stream.
 * The library keeps no state besides the thread-local error string: it is
 * re-entrant, safe under one-process-per-GPU data parallelism and capturable
 * in a CUDA graph.
 */
#ifndef DCNV3_B200_H_
#define DCNV3_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCNV3_B200_VERSION 200 /* 0.2.0 */

/* storage dtypes (op-math is float for F32/F16/BF16, double for F64 — as
 * at::opmath_type in dcnv3_im2col_cuda.cuh:30) */
enum {
    DCNV3_B200_F32 = 0,
    DCNV3_B200_F16 = 1,
    DCNV3_B200_BF16 = 2,
    DCNV3_B200_F64 = 3
};

/* grad_input accumulation for 16-bit storage (ignored for F32/F64):
 *   ACC_OPMATH  accumulate in an fp32 workspace, round once at the end
 *               (what the reference does, dcnv3_cuda.cu:126-133,168-170);
 *   ACC_STORAGE packed 16-bit vector reductions straight into grad_input
 *               (no workspace, less traffic, one rounding per contribution);
 *   ACC_TILE    one kernel, no workspace: the contributions of a 4x8 band of
 *               output pixels are summed in fp32 on the SM and the band's window
 *               is added into grad_input in the storage dtype (TMA reduce-add /
 *               packed 16-bit vector reductions), so a cell of grad_input sees at
 *               most six roundings (one per band window that reaches it)
 *               instead of one (ACC_OPMATH) or ~36 (ACC_STORAGE).  Sampling points
 *               further than 3 px from their kernel-grid position leave the window
 *               and are reduced one by one like ACC_STORAGE.  Shapes the tile kernel
 *               does not take (group_channels != 16, kernel != 3x3 s1 d1, group % 4)
 *               run as ACC_OPMATH and need its workspace. */
enum {
    DCNV3_B200_ACC_OPMATH = 0,
    DCNV3_B200_ACC_STORAGE = 1,
    DCNV3_B200_ACC_TILE = 2
};

/* error codes */
enum {
    DCNV3_B200_OK = 0,
    DCNV3_B200_EINVAL = -1,    /* bad geometry / dtype / flag */
    DCNV3_B200_ENULL = -2,     /* null pointer for a required buffer */
    DCNV3_B200_EALIGN = -3,    /* buffer not aligned for the vector path (16 B) */
    DCNV3_B200_EWORKSPACE = -4,/* workspace missing or too small */
    DCNV3_B200_ERANGE = -5,    /* a tensor has >= 2^31 * 16 B addressable units */
    DCNV3_B200_EDEVICE = -6,   /* no sm_100 device / wrong device */
    DCNV3_B200_ENOTSUP = -7    /* the *_packed entry points only: shape / dtype / alignment outside what the staged-window
                                  kernels take; nothing was launched, call the unpacked entry points instead */
};

/* geometry shared by every call (reference argument order kept:
 * kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w,
 * group, group_channels, offset_scale — dcnv3_cuda.h:16-21) */
typedef struct dcnv3_b200_geometry {
    int N, H, W;          /* input [N, H, W, group*group_channels] */
    int kernel_h, kernel_w;
    int stride_h, stride_w;
    int pad_h, pad_w;
    int dilation_h, dilation_w;
    int group, group_channels;
    float offset_scale;
} dcnv3_b200_geometry;

int dcnv3_b200_version(void);

/* The library reads its DCNV3_B200_* tuning knobs from the environment once per process; this
 * re-reads them (for tests that switch kernel families in-process).  Not for production use. */
void dcnv3_b200_reload_knobs(void);
const char *dcnv3_b200_last_error(void);

/* Ho/Wo of dcnv3_cuda.cu:40-45.  Returns EINVAL if the geometry is not valid. */
int dcnv3_b200_output_size(const dcnv3_b200_geometry *geo, int *Ho, int *Wo);

/* Replaces DCNv3.dcnv3_forward (dcnv3_cuda.cu:21-85).  `output` is fully
 * overwritten (no zero-fill needed). */
int dcnv3_b200_forward(const void *input, const void *offset, const void *mask,
                       void *output, int dtype, const dcnv3_b200_geometry *geo,
                       int mask_is_logits, void *cuda_stream);

/* Bytes of scratch dcnv3_b200_backward needs for this call (0 when none): for 16-bit storage
 * with ACC_OPMATH the fp32 accumulators of grad_input plus 256 bytes in which the library keeps
 * the device-side choice between its two backward kernel families for that call. */
size_t dcnv3_b200_backward_workspace_bytes(int dtype, const dcnv3_b200_geometry *geo,
                                           int grad_accum);

/* Replaces DCNv3.dcnv3_backward (dcnv3_cuda.cu:87-173).  grad_input,
 * grad_offset and grad_mask are fully overwritten, in the storage dtype; the
 * zero-fill grad_input needs is done inside, on `cuda_stream`.  With
 * mask_is_logits, grad_mask is the gradient w.r.t. the logits. */
int dcnv3_b200_backward(const void *input, const void *offset, const void *mask,
                        const void *grad_output, void *grad_input, void *grad_offset,
                        void *grad_mask, void *workspace, size_t workspace_bytes,
                        int dtype, const dcnv3_b200_geometry *geo, int mask_is_logits,
                        int grad_accum, void *cuda_stream);

/* Packed sampling heads (no reference counterpart; LIB/modules/dcnv3.py:121-123 computes offset and mask with two
 * Linear layers).  `heads` is ONE tensor [N, Ho, Wo, 3*G*P] in the storage dtype: per pixel the G*P*2 offsets in the
 * reference's order followed by the G*P masks (or mask logits) — the output of a single Linear(C, 3*G*P) whose weight is
 * the two heads' weights stacked.  The kernels read it with a pixel pitch, so the split costs no copy, and the backward
 * writes `grad_heads` in the same layout, so the Linear's backward sees one tensor.  16-bit storage, group_channels = 16,
 * 3x3 s1 d1, group % 8 == 0 (16-byte pixel pitch), 16-byte aligned buffers; grad_input accumulates as ACC_TILE (no
 * workspace).  Anything else: ENOTSUP, nothing launched — split the tensor and call the functions above. */
int dcnv3_b200_forward_packed(const void *input, const void *heads, void *output, int dtype,
                              const dcnv3_b200_geometry *geo, int mask_is_logits, void *cuda_stream);
int dcnv3_b200_backward_packed(const void *input, const void *heads, const void *grad_output,
                               void *grad_input, void *grad_heads, int dtype,
                               const dcnv3_b200_geometry *geo, int mask_is_logits, void *cuda_stream);

/* The integer contract, exposed for parity tests: for every (n, ho, wo, g, p)
 *   hw_low [N,Ho,Wo,G,P,2] int32 = (h_low, w_low)   (dcnv3_im2col_cuda.cuh:39-40)
 *   bounds [N,Ho,Wo,G,P]   uint8: bit0 inside gate (:262-263), bit1..4 validity
 *          of corners (h_low,w_low) (h_low,w_high) (h_high,w_low) (h_high,w_high)
 *          (:57,62,67,72); all zero, and hw_low = (0,0), when the gate is closed.
 * Computed by the same device function the forward/backward kernels use. */
int dcnv3_b200_debug_indices(const void *offset, int32_t *hw_low, uint8_t *bounds,
                             int dtype, const dcnv3_b200_geometry *geo, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* DCNV3_B200_H_ */

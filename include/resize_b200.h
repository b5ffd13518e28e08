/*
 * resize_b200 — C-ABI of the NHWC resize kernels of the C3-DCN seg models' heads.
 *
 * Replaces, for NHWC-contiguous ("channels_last") CUDA activations, the ATen kernels behind
 *   nn.Upsample(scale_factor=s, mode='nearest')                       unet-lite/yolo5-seg/yolov5_seg.yaml head
 *   F.interpolate(x, size=..., mode='bilinear', align_corners=False)  Concat.forward, seg_diceloss_yolov5.py:484-507
 * with the same index arithmetic: nearest (integer factors only) src = dst / s; bilinear src = max((dst + 0.5) *
 * (in / out) - 0.5, 0), i0 = min(floor(src), in - 1), i1 = min(i0 + 1, in - 1), weights 1 - frac / frac in float.
 * The backward is a gather (every input pixel sums the output pixels that read it): no atomics, deterministic.
 *
 * x / gx: [N, H, W, C]; y / gy: [N, Ho, Wo, C]; one dtype (0 float32, 1 float16, 2 bfloat16); C a multiple of the
 * 16-byte vector (4 / 8 elements).  mode: 0 nearest (Ho = sh*H, Wo = sw*W), 1 bilinear (any sizes).
 * Caller allocates; stream passed; nothing synchronises.  Return: 0 ok, < 0 argument error, > 0 cudaError_t.
 */
#ifndef RESIZE_B200_H_
#define RESIZE_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define RESIZE_B200_VERSION 100

int resize_b200_version(void);
const char* resize_b200_last_error(void);
int resize_b200_supported(int dtype, int C, int H, int W, int Ho, int Wo, int mode);
int resize_b200_forward(const void* x, void* y, int dtype, int N, int H, int W, int C, int Ho, int Wo, int mode,
                        void* cuda_stream);
int resize_b200_backward(const void* gy, void* gx, int dtype, int N, int H, int W, int C, int Ho, int Wo, int mode,
                         void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif

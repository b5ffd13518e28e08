// Issue-rate microbenchmark for DESIGN.md §7 item 3: bf16 x weight + fp32 accumulate over a 16-byte slab (8 values),
//   (a) today's forward: 8 integer unpacks (SHL / LOP3) + 4 FFMA2 with an fp32 weight,
//   (b) 8 FHFMA.BF16 (PTX fma.rn.f32.bf16: 16-bit operands taken from register halves, weight rounded to bf16).
// Registers only, 8 independent accumulator chains per thread.   nvcc -arch=sm_100a -O3 -o tools/fhfma.bin tools/fhfma_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 unpack_bf16x2(unsigned w) {
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}

template <int MODE>
__global__ void __launch_bounds__(256) chain(const uint4* __restrict__ in, float* __restrict__ out, int iters, float wf) {
    uint4 v = in[blockIdx.x * 256 + threadIdx.x];
    float2 acc[4] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
    unsigned short wb = (unsigned short)(__float_as_uint(wf) >> 16);
    for (int i = 0; i < iters; ++i) {
        unsigned wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (MODE == 0) {
                acc[k] = __ffma2_rn(unpack_bf16x2(wd[k]), make_float2(wf, wf), acc[k]);
            } else {
                unsigned short lo, hi;
                asm("mov.b32 {%0, %1}, %2;" : "=h"(lo), "=h"(hi) : "r"(wd[k]));
                asm volatile("fma.rn.f32.bf16 %0, %1, %2, %0;" : "+f"(acc[k].x) : "h"(lo), "h"(wb));
                asm volatile("fma.rn.f32.bf16 %0, %1, %2, %0;" : "+f"(acc[k].y) : "h"(hi), "h"(wb));
            }
        }
        v.x += 0x00010001u; v.y += 0x00010001u; v.z += 0x00010001u; v.w += 0x00010001u;   // keep the unpack in the loop
    }
    out[blockIdx.x * 256 + threadIdx.x] = acc[0].x + acc[0].y + acc[1].x + acc[1].y + acc[2].x + acc[2].y + acc[3].x + acc[3].y;
}

int main() {
    const int blocks = 148 * 8, iters = 8192;
    uint4* in; float* out;
    cudaMalloc(&in, blocks * 256 * sizeof(uint4));
    cudaMalloc(&out, blocks * 256 * sizeof(float));
    cudaMemset(in, 0x3c, blocks * 256 * sizeof(uint4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) {
        float best = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) chain<0><<<blocks, 256>>>(in, out, iters, 0.37f);
            else chain<1><<<blocks, 256>>>(in, out, iters, 0.37f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) best = ms;
        }
        const double vals = (double)blocks * 256 * iters * 8;
        printf("%s: %.3f ms, %.2f Tvalues/s\n", mode == 0 ? "unpack + FFMA2 (fp32 weight)" : "FHFMA.BF16 (bf16 weight)  ", best,
               vals / (best * 1e-3) / 1e12);
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

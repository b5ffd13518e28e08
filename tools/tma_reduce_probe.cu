// Probe: does cp.reduce.async.bulk.tensor.4d (.add) work on this GPU for a 16-bit tensor map, with / without the 32-byte swizzle?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tma_reduce_probe.bin tools/tma_reduce_probe.cu && tools/tma_reduce_probe.bin
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>
#include <vector>
__global__ void k(const __grid_constant__ CUtensorMap tm, int x0, int y0, int mode) {
    extern __shared__ __align__(256) unsigned char smem[];
    __nv_bfloat16 *s = reinterpret_cast<__nv_bfloat16 *>(smem);
    for (int i = threadIdx.x; i < 6 * 16 * 16; i += 32) s[i] = __float2bfloat16(1.0f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (threadIdx.x == 0) {
        const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem);
        if (mode == 0)
            asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(&tm)), "r"(a), "r"(0), "r"(x0), "r"(y0), "r"(0) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(&tm)), "r"(a), "r"(0), "r"(x0), "r"(y0), "r"(0) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}
int main() {
    typedef CUresult (*enc_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                              const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *f = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr);
    enc_t enc = (enc_t)f;
    const int C = 64, W = 32, H = 24, N = 1;
    __nv_bfloat16 *g; cudaMalloc(&g, (size_t)C * W * H * N * 2);
    const CUtensorMapDataType dts[3] = {CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, CU_TENSOR_MAP_DATA_TYPE_UINT16};
    const char *dn[3] = {"bf16", "fp16", "u16"};
    for (int mode = 0; mode < 2; ++mode) for (int d = 0; d < 2; ++d) for (int sw = 0; sw < 2; ++sw) for (int oob = 0; oob < 2; ++oob) {  // oob = 2 (a negative start coordinate) raises "illegal instruction" for the reduce: gpurun_out/tma_reduce_probe2.log
        cudaMemset(g, 0, (size_t)C * W * H * 2);
        const cuuint64_t dims[4] = {C, W, H, N}, str[3] = {C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
        const cuuint32_t box[4] = {16, 16, 6, 1}, es[4] = {1, 1, 1, 1};
        alignas(64) CUtensorMap tm;
        CUresult r = enc(&tm, dts[d], 4, g, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        k<<<1, 32, 4096>>>(tm, oob == 1 ? W - 8 : oob == 2 ? -4 : 8, oob == 1 ? H - 3 : oob == 2 ? 6 : 6, mode);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<__nv_bfloat16> h((size_t)C * W * H);
        double sum = 0;
        if (e == cudaSuccess) { cudaMemcpy(h.data(), g, h.size() * 2, cudaMemcpyDeviceToHost); for (auto v : h) sum += __bfloat162float(v); }
        printf("%s %s swizzle32=%d oob=%d: encode=%d run=%s sum=%.0f (expect %d in bounds)\n", mode ? "store " : "reduce", dn[d], sw, oob, (int)r,
               cudaGetErrorString(e), sum, oob == 1 ? 16 * 8 * 3 : oob == 2 ? 16 * 12 * 6 : 16 * 16 * 6);
        if (e != cudaSuccess) { printf("context lost after this case\n"); return 0; }
    }
    return 0;
}

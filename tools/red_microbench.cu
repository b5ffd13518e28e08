// Microbenchmark: cost of global vector reductions (REDG.E.ADD.F32x4) on the SM -> L2 path as a
// function of how many contiguous bytes of one 128-byte line a warp instruction covers.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/red_microbench.bin tools/red_microbench.cu
// Each warp instruction writes 32 lanes x 16 B = 512 B, split into 512/span groups; every group goes to
// a pseudo-random, span-aligned (or deliberately line-straddling) place in a 64 MB fp32 buffer.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned hash32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

template <int SPAN, bool STRADDLE>
__global__ void red_kernel(float *buf, unsigned n_slots /* number of 128-byte lines */, int iters) {
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31;
    constexpr int LPG = SPAN / 16;  // lanes per contiguous group
    const unsigned grp = lane / LPG, lig = lane % LPG;
    for (int it = 0; it < iters; ++it) {
        const unsigned h = hash32((tid / 32) * 131071u + it * 2654435761u + grp * 97u);
        unsigned line = h % (n_slots - 2);
        // byte offset inside the line: span-aligned, or shifted by half a span across the line end
        unsigned sub = (SPAN >= 128) ? 0u : ((h >> 20) % (128 / SPAN)) * SPAN;
        unsigned off = line * 128u + sub + (STRADDLE ? 128u - SPAN / 2 - sub : 0u) + lig * 16u;
        float *p = reinterpret_cast<float *>(reinterpret_cast<char *>(buf) + off);
        asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(p), "f"(1.0f) : "memory");
    }
}

template <int SPAN, bool STRADDLE> void run(float *buf, unsigned n_lines, const char *name) {
    const int iters = 64, threads = 256, blocks = 148 * 16;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    red_kernel<SPAN, STRADDLE><<<blocks, threads>>>(buf, n_lines, 8);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        red_kernel<SPAN, STRADDLE><<<blocks, threads>>>(buf, n_lines, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    const double bytes = (double)blocks * threads * iters * 16.0;
    const double cyc_per_sm = best * 1e-3 * 1.965e9;
    const double sectors_per_sm = bytes / 32.0 / 148.0;
    printf("| %-34s | %7.3f ms | %7.1f GB/s | %5.2f B/clk/SM | %4.2f clk/sector |\n", name, best,
           bytes / best / 1e6, bytes / 148.0 / cyc_per_sm, cyc_per_sm / sectors_per_sm);
}

int main() {
    const size_t bytes = 64u << 20;
    float *buf;
    cudaMalloc(&buf, bytes);
    cudaMemset(buf, 0, bytes);
    const unsigned n_lines = (unsigned)(bytes / 128);
    printf("| contiguous bytes per group of lanes    | time       | payload      | per SM          | port cost       |\n");
    printf("|---|---|---|---|---|\n");
    run<16, false>(buf, n_lines, "16 B (one lane, half a sector)");
    run<32, false>(buf, n_lines, "32 B (2 lanes, one sector)");
    run<64, false>(buf, n_lines, "64 B (4 lanes, 2 sectors, one line)");
    run<64, true>(buf, n_lines, "64 B straddling two lines");
    run<128, false>(buf, n_lines, "128 B (8 lanes, a whole line)");
    run<128, true>(buf, n_lines, "128 B straddling two lines");
    run<256, false>(buf, n_lines, "256 B (16 lanes, two lines)");
    run<512, false>(buf, n_lines, "512 B (whole warp, four lines)");
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}

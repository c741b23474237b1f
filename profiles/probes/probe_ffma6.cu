// probe_ffma6.cu -- the kernel's own gemm_tile (run-time strides, packed FMAs) on its own in a
// 512-thread CTA per SM, 200 tokens x 80 frames per unit, random operands.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../glow-tts-train_b200/csrc/mas_logp_tile.cuh"

using namespace mas;

template <bool kMeanOnly>
__global__ void __launch_bounds__(512, 1) tile(float *out, int iters, int D, int tile_rows, int F, int CG, int RG) {
    extern __shared__ __align__(16) float sm[];
    float *sa = sm, *sb = sm + D * tile_rows, *sz = sm + 2 * D * tile_rows;
    uint32_t s = threadIdx.x * 2654435761u + blockIdx.x;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 32768.0f - 1.0f; };
    for (int i = threadIdx.x; i < D * tile_rows; i += blockDim.x) { sa[i] = -0.5f * (1.0f + 0.3f * rnd()); sb[i] = rnd(); }
    for (int i = threadIdx.x; i < D * F; i += blockDim.x) sz[i] = 2.f * rnd();
    __syncthreads();
    const int rg = threadIdx.x / CG, cg = threadIdx.x - rg * CG;
    GemmAcc acc;
    float t = 0.f;
    for (int it = 0; it < iters; ++it) {
        asm volatile("bar.sync 1, 512;");
        if (rg < RG) {
            gemm_tile<true, kMeanOnly>(sa, sb, sz, D, tile_rows, F, rg, cg, acc);
            float c[4];
            acc.quad(it & 3, it & 1, c);
            t += c[0] + c[3];
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

int main() {
    float *out; cudaMalloc(&out, 148 * 512 * 4);
    const int D = 80, rows = 200, F = 80, CG = 10, RG = 50, iters = 40;
    const int smem = (2 * D * rows + D * F) * 4;
    cudaFuncSetAttribute(tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) {
        float best = 1e9;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) tile<false><<<148, 512, smem>>>(out, iters, D, rows, F, CG, RG);
            else tile<true><<<148, 512, smem>>>(out, iters, D, rows, F, CG, RG);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("gemm_tile %-10s %7.2f us per 200x80 unit   %s\n", mode ? "mean_only" : "general", best * 1e3 / iters, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}

// probe_ffma.cu -- FP32 FFMA issue rate on B200 for the log-likelihood contraction: 8x8 register
// tile, operands from shared memory (6 LDS.128 + 8 FMUL per 128 FFMA), different FFMA orders.
// Register-bank conflicts (two fresh source registers of the same parity) make an FFMA take two
// issue cycles; only the order of the FFMAs in the source is under our control with nvcc.
#include <cstdio>
#include <cuda_runtime.h>

template <int ORDER>
__global__ void __launch_bounds__(256, 1) ffma_tile(float *out, int iters, long long *cyc) {
    extern __shared__ __align__(16) float sm[];
    float *sa = sm, *sb = sm + 80 * 208, *sz = sm + 2 * 80 * 208;
    for (int i = threadIdx.x; i < 80 * 208; i += blockDim.x) { sa[i] = 1.0f + 1e-3f * (i % 13); sb[i] = 0.5f; }
    for (int i = threadIdx.x; i < 80 * 64; i += blockDim.x) sz[i] = 1e-3f * (i % 17);
    __syncthreads();
    const int rg = threadIdx.x >> 3, cg = threadIdx.x & 7;
    float acc[8][8];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    long long t0 = clock64();
    if (rg * 8 < 208)
    for (int it = 0; it < iters; ++it) {
        const float *pa = sa + rg * 8, *pb = sb + rg * 8, *pz = sz + cg * 4;
#pragma unroll 2
        for (int d = 0; d < 80; ++d) {
            const float4 a0 = *reinterpret_cast<const float4 *>(pa), a1 = *reinterpret_cast<const float4 *>(pa + 4);
            const float4 b0 = *reinterpret_cast<const float4 *>(pb), b1 = *reinterpret_cast<const float4 *>(pb + 4);
            const float4 z0 = *reinterpret_cast<const float4 *>(pz), z1 = *reinterpret_cast<const float4 *>(pz + 32);
            pa += 208; pb += 208; pz += 64;
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            const float zv[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
            float qv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) qv[j] = -0.5f * (zv[j] * zv[j]);
            if (ORDER == 0) {          // cell by cell, both terms
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) { acc[i][j] = fmaf(av[i], qv[j], acc[i][j]); acc[i][j] = fmaf(bv[i], zv[j], acc[i][j]); }
            } else if (ORDER == 1) {   // two sweeps, token outer
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], qv[j], acc[i][j]);
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(bv[i], zv[j], acc[i][j]);
            } else if (ORDER == 2) {   // two sweeps, frame outer
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i][j] = fmaf(av[i], qv[j], acc[i][j]);
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i][j] = fmaf(bv[i], zv[j], acc[i][j]);
            } else if (ORDER == 3) {   // two sweeps, boustrophedon (each FFMA shares one operand with its predecessor)
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) { const int j = (i & 1) ? 7 - jj : jj; acc[i][j] = fmaf(av[i], qv[j], acc[i][j]); }
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) { const int j = (i & 1) ? 7 - jj : jj; acc[i][j] = fmaf(bv[i], zv[j], acc[i][j]); }
            } else if (ORDER == 4) {   // single term only (mean_only shape): token outer
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(bv[i], zv[j], acc[i][j]);
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int ORDER>
void run(float *out, long long *cyc, const char *name) {
    const int smem = (2 * 80 * 208 + 80 * 64) * 4, iters = 20;
    cudaFuncSetAttribute(ffma_tile<ORDER>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long h;
    for (int rep = 0; rep < 2; ++rep) ffma_tile<ORDER><<<148, 224, smem>>>(out, iters, cyc);
    cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double ffma = (ORDER == 4 ? 64.0 : 128.0) * 80 * iters * 208;
    printf("%-44s %.1f FFMA/cycle/SM (of 128)   err=%s\n", name, ffma / h, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 8);
    run<0>(out, cyc, "cell by cell (both terms)");
    run<1>(out, cyc, "two sweeps, token outer");
    run<2>(out, cyc, "two sweeps, frame outer");
    run<3>(out, cyc, "two sweeps, boustrophedon");
    run<4>(out, cyc, "one term only, token outer");
    return 0;
}

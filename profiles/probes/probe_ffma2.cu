// probe_ffma2.cu -- FFMA rate of the log-likelihood contraction per SM for the CTA shapes the fused
// launch can use: token tile 208 or 104, thread tile 8x8 or 4x8, one or two CTAs per SM.
#include <cstdio>
#include <cuda_runtime.h>

template <int ROWS, int TM>     // TM tokens x 8 frames per thread
__global__ void __launch_bounds__(448) tile(float *out, int iters, long long *cyc) {
    extern __shared__ __align__(16) float sm[];
    float *sa = sm, *sb = sm + 80 * ROWS, *sz = sm + 2 * 80 * ROWS;
    for (int i = threadIdx.x; i < 80 * ROWS; i += blockDim.x) { sa[i] = 1.0f + 1e-3f * (i % 13); sb[i] = 0.5f; }
    for (int i = threadIdx.x; i < 80 * 64; i += blockDim.x) sz[i] = 1e-3f * (i % 17);
    __syncthreads();
    const int rg = threadIdx.x >> 3, cg = threadIdx.x & 7;
    float acc[TM][8];
    for (int i = 0; i < TM; ++i) for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    long long t0 = clock64();
    if (rg * TM < ROWS)
    for (int it = 0; it < iters; ++it) {
        const float *pa = sa + rg * TM, *pb = sb + rg * TM, *pz = sz + cg * 4;
#pragma unroll 2
        for (int d = 0; d < 80; ++d) {
            float av[TM], bv[TM];
#pragma unroll
            for (int q = 0; q < TM / 4; ++q) {
                const float4 a = *reinterpret_cast<const float4 *>(pa + 4 * q), b = *reinterpret_cast<const float4 *>(pb + 4 * q);
                av[4 * q] = a.x; av[4 * q + 1] = a.y; av[4 * q + 2] = a.z; av[4 * q + 3] = a.w;
                bv[4 * q] = b.x; bv[4 * q + 1] = b.y; bv[4 * q + 2] = b.z; bv[4 * q + 3] = b.w;
            }
            const float4 z0 = *reinterpret_cast<const float4 *>(pz), z1 = *reinterpret_cast<const float4 *>(pz + 32);
            pa += ROWS; pb += ROWS; pz += 64;
            const float zv[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
            float qv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) qv[j] = -0.5f * (zv[j] * zv[j]);
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], qv[j], acc[i][j]);
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(bv[i], zv[j], acc[i][j]);
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < TM; ++i) for (int j = 0; j < 8; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int ROWS, int TM>
void run(float *out, long long *cyc, int ctas_per_sm, const char *name) {
    const int smem = (2 * 80 * ROWS + 80 * 64) * 4, iters = 20;
    const int threads = ((ROWS / TM * 8) + 31) / 32 * 32;
    cudaFuncSetAttribute(tile<ROWS, TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long h;
    for (int rep = 0; rep < 2; ++rep) tile<ROWS, TM><<<148 * ctas_per_sm, threads, smem>>>(out, iters, cyc);
    cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double ffma = 2.0 * 8 * 80 * iters * ROWS * 8 / 8.0 * ctas_per_sm;   // per SM: ROWS x 64 cells x 160
    printf("%-52s %3d thr x %d CTA/SM: %.1f FFMA/cycle/SM  err=%s\n", name, threads, ctas_per_sm, ffma * 8 / h, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float *out; long long *cyc;
    cudaMalloc(&out, 296 * 256 * 4); cudaMalloc(&cyc, 8);
    run<208, 8>(out, cyc, 1, "208 tokens, 8x8 tiles");
    run<104, 8>(out, cyc, 1, "104 tokens, 8x8 tiles");
    run<104, 8>(out, cyc, 2, "104 tokens, 8x8 tiles");
    run<104, 4>(out, cyc, 1, "104 tokens, 4x8 tiles");
    run<104, 4>(out, cyc, 2, "104 tokens, 4x8 tiles");
    run<208, 4>(out, cyc, 1, "208 tokens, 4x8 tiles (416 thr)");
    return 0;
}

// probe_ffma3.cu -- FFMA rate of the log-likelihood contraction per SM, whole-kernel timing (events),
// for CTA geometries (threads, row groups x column groups) and operand prefetch depth.
//   thread tile: 4 tokens x 8 frames ({4 cg..4 cg+3} and {4 CG + 4 cg ..}); chunk = 8 CG frames
#include <cstdio>
#include <cuda_runtime.h>

template <int ROWS, int CG, int THREADS, bool PREFETCH, int UNROLL>
__global__ void __launch_bounds__(THREADS, 2) tile(float *out, int iters) {
    extern __shared__ __align__(16) float sm[];
    constexpr int F = 8 * CG;
    float *sa = sm, *sb = sm + 80 * ROWS, *sz = sm + 2 * 80 * ROWS;
    for (int i = threadIdx.x; i < 80 * ROWS; i += blockDim.x) { sa[i] = 1.0f + 1e-3f * (i % 13); sb[i] = 0.5f; }
    for (int i = threadIdx.x; i < 80 * F; i += blockDim.x) sz[i] = 1e-3f * (i % 17);
    __syncthreads();
    const int rg = threadIdx.x / CG, cg = threadIdx.x % CG;
    float acc[4][8];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    if (rg * 4 < ROWS)
    for (int it = 0; it < iters; ++it) {
        const float *pa = sa + rg * 4, *pb = sb + rg * 4, *pz = sz + cg * 4;
        if (PREFETCH) {
            float4 a = *reinterpret_cast<const float4 *>(pa), b = *reinterpret_cast<const float4 *>(pb);
            float4 z0 = *reinterpret_cast<const float4 *>(pz), z1 = *reinterpret_cast<const float4 *>(pz + 4 * CG);
#pragma unroll UNROLL
            for (int d = 0; d < 80; ++d) {
                const int dn = d + 1 < 80 ? d + 1 : 79;
                const float4 na = *reinterpret_cast<const float4 *>(pa + dn * ROWS), nb = *reinterpret_cast<const float4 *>(pb + dn * ROWS);
                const float4 nz0 = *reinterpret_cast<const float4 *>(pz + dn * F), nz1 = *reinterpret_cast<const float4 *>(pz + dn * F + 4 * CG);
                const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
                const float zv[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
                float qv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) qv[j] = zv[j] * zv[j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], qv[j], acc[i][j]);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(bv[i], zv[j], acc[i][j]);
                a = na; b = nb; z0 = nz0; z1 = nz1;
            }
        } else {
#pragma unroll UNROLL
            for (int d = 0; d < 80; ++d) {
                const float4 a = *reinterpret_cast<const float4 *>(pa), b = *reinterpret_cast<const float4 *>(pb);
                const float4 z0 = *reinterpret_cast<const float4 *>(pz), z1 = *reinterpret_cast<const float4 *>(pz + 4 * CG);
                pa += ROWS; pb += ROWS; pz += F;
                const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
                const float zv[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
                float qv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) qv[j] = zv[j] * zv[j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], qv[j], acc[i][j]);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(bv[i], zv[j], acc[i][j]);
            }
        }
    }
    float s = 0;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 8; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ROWS, int CG, int THREADS, bool PREFETCH, int UNROLL>
void run(float *out, const char *name) {
    constexpr int F = 8 * CG;
    const int smem = (2 * 80 * ROWS + 80 * F) * 4, iters = 40;
    cudaFuncSetAttribute(tile<ROWS, CG, THREADS, PREFETCH, UNROLL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tile<ROWS, CG, THREADS, PREFETCH, UNROLL>, THREADS, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        tile<ROWS, CG, THREADS, PREFETCH, UNROLL><<<148 * 2, THREADS, smem>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double useful = 2.0 * (double)ROWS * F * 160 * iters;      // FFMA per SM (2 CTAs), useful cells only
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-40s thr %3d occ %d smem %6d : %7.1f us  %6.1f useful FFMA/cycle/SM at %.2f GHz   %s\n", name, THREADS, occ, smem, best * 1e3,
           useful / (best * 1e-3 * clk * 1e3), clk * 1e-6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float *out; cudaMalloc(&out, 296 * 512 * 4);
    run<104, 8, 224, false, 2>(out, "104x64  26rg x 8cg   (current)");
    run<112, 8, 224, false, 2>(out, "112x64  28rg x 8cg");
    run<104, 8, 224, true, 2>(out, "104x64  26rg x 8cg   prefetch");
    run<100, 10, 256, false, 2>(out, "100x80  25rg x 10cg");
    run<100, 10, 256, true, 2>(out, "100x80  25rg x 10cg  prefetch");
    run<128, 8, 256, false, 2>(out, "128x64  32rg x 8cg");
    run<128, 8, 256, true, 2>(out, "128x64  32rg x 8cg   prefetch");
    run<128, 8, 256, true, 4>(out, "128x64  32rg x 8cg   prefetch unroll 4");
    run<64, 16, 256, false, 2>(out, "64x128  16rg x 16cg");
    run<64, 16, 256, true, 2>(out, "64x128  16rg x 16cg  prefetch");
    return 0;
}

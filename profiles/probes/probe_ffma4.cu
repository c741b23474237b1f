// probe_ffma4.cu -- inner-loop variants of the contraction at the kernel's geometry (512 threads,
// 200 tokens x 80 frames, one CTA per SM): FFMA issue order and packed fma.rn.f32x2.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ROWS = 200, CG = 10, F = 80, THREADS = 512;

__device__ __forceinline__ uint64_t pack(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack(uint64_t v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

template <int MODE>   // 0: i outer (kernel today)  1: j outer  2: f32x2
__global__ void __launch_bounds__(THREADS, 1) tile(float *out, int iters) {
    extern __shared__ __align__(16) float sm[];
    float *sa = sm, *sb = sm + 80 * ROWS, *sz = sm + 2 * 80 * ROWS;
    for (int i = threadIdx.x; i < 80 * ROWS; i += blockDim.x) { sa[i] = 1.0f + 1e-3f * (i % 13); sb[i] = 0.5f; }
    for (int i = threadIdx.x; i < 80 * F; i += blockDim.x) sz[i] = 1e-3f * (i % 17);
    __syncthreads();
    const int rg = threadIdx.x / CG, cg = threadIdx.x % CG;
    float s = 0;
    if (rg * 4 < ROWS) {
        if (MODE < 2) {
            float acc[4][8];
            for (int i = 0; i < 4; ++i) for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
            for (int it = 0; it < iters; ++it) {
                const float *pa = sa + rg * 4, *pb = sb + rg * 4, *pz = sz + cg * 4;
#pragma unroll 2
                for (int d = 0; d < 80; ++d) {
                    const float4 a = *reinterpret_cast<const float4 *>(pa), b = *reinterpret_cast<const float4 *>(pb);
                    const float4 z0 = *reinterpret_cast<const float4 *>(pz), z1 = *reinterpret_cast<const float4 *>(pz + 4 * CG);
                    pa += ROWS; pb += ROWS; pz += F;
                    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
                    const float zv[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
                    float qv[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) qv[j] = zv[j] * zv[j];
                    if (MODE == 0) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], qv[j], acc[i][j]);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(bv[i], zv[j], acc[i][j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
#pragma unroll
                            for (int i = 0; i < 4; ++i) acc[i][j] = fmaf(av[i], qv[j], acc[i][j]);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
#pragma unroll
                            for (int i = 0; i < 4; ++i) acc[i][j] = fmaf(bv[i], zv[j], acc[i][j]);
                    }
                }
            }
            for (int i = 0; i < 4; ++i) for (int j = 0; j < 8; ++j) s += acc[i][j];
        } else {
            uint64_t acc[4][4];
            for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = 0ull;
            for (int it = 0; it < iters; ++it) {
                const float *pa = sa + rg * 4, *pb = sb + rg * 4, *pz = sz + cg * 4;
#pragma unroll 2
                for (int d = 0; d < 80; ++d) {
                    const float4 a = *reinterpret_cast<const float4 *>(pa), b = *reinterpret_cast<const float4 *>(pb);
                    const ulonglong2 z0 = *reinterpret_cast<const ulonglong2 *>(pz), z1 = *reinterpret_cast<const ulonglong2 *>(pz + 4 * CG);
                    pa += ROWS; pb += ROWS; pz += F;
                    const uint64_t av[4] = {pack(a.x, a.x), pack(a.y, a.y), pack(a.z, a.z), pack(a.w, a.w)};
                    const uint64_t bv[4] = {pack(b.x, b.x), pack(b.y, b.y), pack(b.z, b.z), pack(b.w, b.w)};
                    const uint64_t zv[4] = {z0.x, z0.y, z1.x, z1.y};
                    uint64_t qv[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) qv[j] = mul2(zv[j], zv[j]);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = fma2(av[i], qv[j], acc[i][j]);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = fma2(bv[i], zv[j], acc[i][j]);
                }
            }
            for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { float lo, hi; unpack(acc[i][j], lo, hi); s += lo + hi; }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(float *out, const char *name) {
    const int smem = (2 * 80 * ROWS + 80 * F) * 4, iters = 40;
    cudaFuncSetAttribute(tile<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        tile<MODE><<<148, THREADS, smem>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("%-28s %7.2f us per 200x80 unit   %s\n", name, best * 1e3 / iters, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float *out; cudaMalloc(&out, 148 * 512 * 4);
    run<0>(out, "i outer (kernel today)");
    run<1>(out, "j outer");
    run<2>(out, "fma.rn.f32x2");
    return 0;
}

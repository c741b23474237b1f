// probe_ffma7.cu -- what bounds the contraction loop: FMA pipe or shared-memory return bandwidth?
// The kernel's loop (4 tokens x 8 frames per thread, 512 threads) loads 16 floats per 64 FMAs per
// channel: 1 B per FMA = 128 B per clock per SM at the FMA peak, which is all the shared-memory pipe
// returns.  Variants: TM tokens x 8 frames per thread with 512 / TM*64 threads; the two terms as two
// FMAs + a square ("sep") or factored through z ("horner": t = fma(a, z, b); c = fma(t, z, c));
// mean_only (one FMA per cell and channel).  200 tokens x 80 frames x 80 channels per unit.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../glow-tts-train_b200/csrc/mas_logp_tile.cuh"

using namespace mas;

template <int TM, int MODE>   // MODE 0: sep, 1: horner, 2: mean_only
__device__ __forceinline__ void contract(const float *sa, const float *sb, const float *sz, int D, int tile_rows, int F, int rg, int cg,
                                         f32x2 (&acc)[TM][4]) {
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0ull;
    const float *pa = sa + rg * TM, *pb = sb + rg * TM, *pz = sz + cg * 4;
    const int half = F >> 1;
#pragma unroll 8
    for (int d = 0; d < D; ++d) {
        float a[TM], b[TM];
#pragma unroll
        for (int q = 0; q < TM / 4; ++q) {
            const float4 v = *reinterpret_cast<const float4 *>(pb + 4 * q);
            b[4 * q] = v.x, b[4 * q + 1] = v.y, b[4 * q + 2] = v.z, b[4 * q + 3] = v.w;
            if (MODE != 2) {
                const float4 w = *reinterpret_cast<const float4 *>(pa + 4 * q);
                a[4 * q] = w.x, a[4 * q + 1] = w.y, a[4 * q + 2] = w.z, a[4 * q + 3] = w.w;
            }
        }
        const ulonglong2 z0 = *reinterpret_cast<const ulonglong2 *>(pz), z1 = *reinterpret_cast<const ulonglong2 *>(pz + half);
        pa += tile_rows, pb += tile_rows, pz += F;
        const f32x2 zv[4] = {z0.x, z0.y, z1.x, z1.y};
        if (MODE == 0) {
            f32x2 qv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) qv[j] = f32x2_mul(zv[j], zv[j]);
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                const f32x2 av = f32x2_pack(a[i], a[i]);
#pragma unroll
                for (int j = 0; j < 4; ++j) f32x2_fma_acc(acc[i][j], av, qv[j]);
            }
        }
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const f32x2 bv = f32x2_pack(b[i], b[i]);
            if (MODE == 1) {
                const f32x2 av = f32x2_pack(a[i], a[i]);
                f32x2 tv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) tv[j] = f32x2_fma(av, zv[j], bv);
#pragma unroll
                for (int j = 0; j < 4; ++j) f32x2_fma_acc(acc[i][j], tv[j], zv[j]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) f32x2_fma_acc(acc[i][j], bv, zv[j]);
            }
        }
    }
}

template <int TM, int MODE, int NT>
__global__ void __launch_bounds__(NT, 1) tile(float *out, int iters, int D, int tile_rows, int F, int CG, int RG) {
    extern __shared__ __align__(16) float sm[];
    float *sa = sm, *sb = sm + D * tile_rows, *sz = sm + 2 * D * tile_rows;
    uint32_t s = threadIdx.x * 2654435761u + blockIdx.x;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 32768.0f - 1.0f; };
    for (int i = threadIdx.x; i < D * tile_rows; i += blockDim.x) { sa[i] = -0.5f * (1.0f + 0.3f * rnd()); sb[i] = rnd(); }
    for (int i = threadIdx.x; i < D * F; i += blockDim.x) sz[i] = 2.f * rnd();
    __syncthreads();
    const int rg = threadIdx.x / CG, cg = threadIdx.x - rg * CG;
    f32x2 acc[TM][4];
    float t = 0.f;
    for (int it = 0; it < iters; ++it) {
        __syncthreads();
        if (rg < RG) {
            contract<TM, MODE>(sa, sb, sz, D, tile_rows, F, rg, cg, acc);
            float lo, hi;
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) { f32x2_unpack(acc[i][j], lo, hi); t += lo + hi; }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int TM, int MODE, int NT>
static void run(float *out, const char *name) {
    const int D = 80, rows = 200, F = 80, CG = 10, RG = rows / TM, iters = 40;
    const int smem = (2 * D * rows + D * F) * 4;
    cudaFuncSetAttribute(tile<TM, MODE, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        tile<TM, MODE, NT><<<148, NT, smem>>>(out, iters, D, rows, F, CG, RG);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, tile<TM, MODE, NT>);
    printf("%-28s %3d threads %3d regs  %7.2f us per 200x80 unit   %s\n", name, NT, fa.numRegs, best * 1e3 / iters, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float *out; cudaMalloc(&out, 148 * 512 * 4);
    run<4, 0, 512>(out, "4x8 sep (the kernel's)");
    run<4, 1, 512>(out, "4x8 horner");
    run<4, 2, 512>(out, "4x8 mean_only");
    run<8, 0, 256>(out, "8x8 sep");
    run<8, 1, 256>(out, "8x8 horner");
    run<8, 2, 256>(out, "8x8 mean_only");
    return 0;
}

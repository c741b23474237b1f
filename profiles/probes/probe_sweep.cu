// probe_sweep.cu -- cycles per 32-frame block of the real sweep_block<R> (one warp, synthetic tile
// in shared memory, no TMA, no inter-warp waits).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../include -I../../glow-tts-train_b200/csrc -o probe_sweep probe_sweep.cu
#include <cstdio>
#include "../../glow-tts-train_b200/csrc/mas_path_systolic.cu"

namespace mas { thread_local int g_last_cuda_error = 0; long long *g_dbg_cycles = nullptr; }
using namespace mas::systolic;

template <int R, bool kOut>
__global__ void probe(float *out, long long *cycles, int slot, int nblocks) {
    extern __shared__ __align__(1024) float sm[];
    float *tile = sm;                                  // [32R][32]
    float *bin = sm + 32 * R * 32;                     // [32]
    float *bout = bin + 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 32 * R * 32 + 64; i += blockDim.x) sm[i] = -1.0f - 0.001f * (i % 977);
    __syncthreads();
    float v[R];
    uint32_t acc[R];
    for (int i = 0; i < R; ++i) { v[i] = -1e9f; acc[i] = 0; }
    float carry = 0.f;
    long long t0 = clock64();
    for (int k = 0; k < nblocks; ++k)
        sweep_block<R, false>(tile, v, acc, carry, reinterpret_cast<const float4 *>(bin),
                              kOut ? reinterpret_cast<float4 *>(bout) : nullptr, lane, lane * R, 4096 + k * 32);
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < R; ++i) s += v[i] + __uint_as_float(acc[i]);
    out[threadIdx.x] = s + carry;
    if (threadIdx.x == 0) cycles[slot] = (t1 - t0) / nblocks;
}

int main() {
    float *out;
    long long *cyc, h[16] = {0};
    cudaMalloc(&out, 4096);
    cudaMalloc(&cyc, 16 * 8);
    const int nb = 256;
    auto smem = [](int R) { return (32 * R * 32 + 64) * 4; };
    cudaFuncSetAttribute(probe<5, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem(5));
    for (int rep = 0; rep < 2; ++rep) {
        probe<1, false><<<1, 32, smem(1)>>>(out, cyc, 0, nb);
        probe<1, true><<<1, 32, smem(1)>>>(out, cyc, 1, nb);
        probe<3, false><<<1, 32, smem(3)>>>(out, cyc, 2, nb);
        probe<3, true><<<1, 32, smem(3)>>>(out, cyc, 3, nb);
        probe<5, true><<<1, 32, smem(5)>>>(out, cyc, 4, nb);
        probe<3, true><<<1, 128, smem(3)>>>(out, cyc, 5, nb);   // 4 warps, one per scheduler, same tile
        probe<3, true><<<1, 256, smem(3)>>>(out, cyc, 6, nb);
        probe<3, true><<<1, 512, smem(3)>>>(out, cyc, 7, nb);
        probe<1, true><<<1, 256, smem(1)>>>(out, cyc, 8, nb);
        probe<1, true><<<1, 512, smem(1)>>>(out, cyc, 9, nb);
    }
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    const char *names[] = {"R=1", "R=1 +bnd_out", "R=3", "R=3 +bnd_out", "R=5 +bnd_out", "R=3 +bnd_out x4 warps", "R=3 x8 warps", "R=3 x16 warps", "R=1 x8 warps", "R=1 x16 warps"};
    for (int i = 0; i < 10; ++i) printf("%-24s %6lld cycles/block  %.1f cycles/frame\n", names[i], h[i], h[i] / 32.0);
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

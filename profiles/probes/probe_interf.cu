// probe_interf.cu -- does a warp that spin-polls shared memory (ld.acquire) or floods global stores
// on ANOTHER scheduler slow down the sweep warp?  (cycles per 32-frame block of warp 0, R = 3)
#include <cstdio>
#include "../../glow-tts-train_b200/csrc/mas_path_systolic.cu"
namespace mas { thread_local int g_last_cuda_error = 0; long long *g_dbg_cycles = nullptr; }
using namespace mas::systolic;

template <int R, int MODE>   // MODE bit0: warps 1,2 poll; bit1: warp 3 floods STG; bit2: pollers use plain volatile LDS + nanosleep
__global__ void probe(float *out, float4 *sink, long long *cycles, int slot, int nblocks) {
    extern __shared__ __align__(1024) float sm[];
    __shared__ int flag;
    float *tile = sm, *bin = sm + 32 * R * 32, *bout = bin + 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 32 * R * 32 + 64; i += blockDim.x) sm[i] = -1.0f - 0.001f * (i % 977);
    if (threadIdx.x == 0) flag = 0;
    __syncthreads();
    if (warp == 0) {
        float v[R]; uint32_t acc[R];
        for (int i = 0; i < R; ++i) { v[i] = -1e9f; acc[i] = 0; }
        float carry = 0.f;
        long long t0 = clock64();
        for (int k = 0; k < nblocks; ++k)
            sweep_block<R>(tile, v, acc, carry, reinterpret_cast<const float4 *>(bin), reinterpret_cast<float4 *>(bout), lane);
        long long t1 = clock64();
        float s = 0;
        for (int i = 0; i < R; ++i) s += v[i] + __uint_as_float(acc[i]);
        out[lane] = s + carry;
        if (lane == 0) { cycles[slot] = (t1 - t0) / nblocks; mas::ptx::st_release_shared(&flag, 1); }
    } else if (warp < 3) {
        if (MODE & 1) {
            if (MODE & 4) { while (*(volatile int *)&flag == 0) __nanosleep(200); }
            else { while (mas::ptx::ld_acquire_shared(&flag) == 0) {} }
        }
    } else {
        if (MODE & 2) {
            const float4 z = make_float4(0, 0, 0, 0);
            for (int it = 0; it < 4000 && *(volatile int *)&flag == 0; ++it)
                for (int j = 0; j < 4; ++j) mas::ptx::st_global_cs_v4(sink + ((it * 4 + j) * 32 + lane) % (1 << 20), z);
        }
    }
}

int main() {
    float *out; float4 *sink; long long *cyc, h[16] = {0};
    cudaMalloc(&out, 4096); cudaMalloc(&sink, (size_t)(1 << 20) * 16); cudaMalloc(&cyc, 16 * 8);
    const int nb = 64, smem = (32 * 3 * 32 + 64) * 4;
    for (int rep = 0; rep < 2; ++rep) {
        probe<3, 0><<<1, 128, smem>>>(out, sink, cyc, 0, nb);
        probe<3, 1><<<1, 128, smem>>>(out, sink, cyc, 1, nb);
        probe<3, 2><<<1, 128, smem>>>(out, sink, cyc, 2, nb);
        probe<3, 3><<<1, 128, smem>>>(out, sink, cyc, 3, nb);
        probe<3, 5><<<1, 128, smem>>>(out, sink, cyc, 4, nb);
    }
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    const char *names[] = {"alone", "+2 warps polling ld.acquire", "+1 warp flooding STG", "+both", "+2 warps polling w/ nanosleep"};
    for (int i = 0; i < 5; ++i) printf("%-32s %6lld cycles/block\n", names[i], h[i]);
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
}

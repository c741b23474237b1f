// probe_skew.cu -- why a lone sweep warp runs at ~55 cycles per frame, and what lane skew would buy.
// One warp per SM (a sweep warp is alone on its scheduler).  Per frame and token: the kernel's cell
// (diff, max.NaN, add, funnel shift).  Scores come from registers (no memory in the loop).
//   lockstep : the kernel's scheme -- frame j needs the neighbour lane's value of frame j-1: one
//              SHFL.UP per frame whose result is consumed ~10 instructions later; the warp issues in
//              order, so every frame eats the shuffle's latency.
//   skewed   : lane L works four frames behind lane L-1: the four boundary values a lane needs in a
//              step were shuffled during the PREVIOUS step, nothing waits.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float fmax_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void cell(float &v, float adv, float l, uint32_t &acc) {
    const float diff = v - adv;
    v = fmax_nan(adv, v) + l;
    acc = __funnelshift_l(__float_as_uint(diff), acc, 1);
}

template <int R, bool SKEW, bool SEL = true>
__global__ void __launch_bounds__(32) sweep(float *out, int quads, long long *cyc, float l0) {
    const int lane = threadIdx.x;
    float v[R];
    uint32_t acc[R];
    for (int i = 0; i < R; ++i) { v[i] = -1e9f + lane; acc[i] = 0; }
    const float ll = l0 + 1e-3f * lane;                 // lane-dependent: nothing here is warp-uniform
    float l[4] = {ll, ll * 1.5f, ll * 0.5f, ll * 2.f};
    float up4[4] = {-1e9f, -1e9f, -1e9f, -1e9f};       // skewed: boundary values for this step
    const long long t0 = clock64();
    for (int q = 0; q < quads; ++q) {
        if (SKEW) {
            float nxt[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int i = R - 1; i >= 0; --i) cell(v[i], i == 0 ? up4[j] : v[i - 1], l[j] + i, acc[i]);
                nxt[j] = __shfl_up_sync(0xffffffffu, v[R - 1], 1);      // consumed in the next step
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) up4[j] = lane == 0 ? -1e9f : nxt[j];
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float up = __shfl_up_sync(0xffffffffu, v[R - 1], 1);
                if (SEL && lane == 0) up = -1e9f;
#pragma unroll
                for (int i = R - 1; i >= 0; --i) cell(v[i], i == 0 ? up : v[i - 1], l[j] + i, acc[i]);
            }
        }
    }
    const long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < R; ++i) s += v[i] + acc[i];
    out[blockIdx.x * 32 + lane] = s;
    if (lane == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int R, bool SKEW, bool SEL = true>
void run(float *out, long long *cyc, const char *name) {
    const int quads = 4096;
    long long h = 0;
    for (int rep = 0; rep < 2; ++rep) sweep<R, SKEW, SEL><<<148, 32>>>(out, quads, cyc, -3.f);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("R=%d %-28s %6.1f cycles per frame  (%s)\n", R, name, (double)h / (4.0 * quads), cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 32 * 4); cudaMalloc(&cyc, 8);
    run<1, false>(out, cyc, "lockstep"); run<1, true>(out, cyc, "skewed");
    run<2, false>(out, cyc, "lockstep"); run<2, true>(out, cyc, "skewed");
    run<3, false>(out, cyc, "lockstep"); run<3, true>(out, cyc, "skewed");
    run<4, false>(out, cyc, "lockstep"); run<4, true>(out, cyc, "skewed");
    run<2, false, false>(out, cyc, "lockstep, no lane-0 select"); run<3, false, false>(out, cyc, "lockstep, no lane-0 select");
    return 0;
}

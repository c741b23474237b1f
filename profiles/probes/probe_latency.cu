// probe_latency.cu -- dependent-chain latencies on sm_100a that bound the MAS sweep per mel frame.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_latency probe_latency.cu && ./probe_latency
#include <cstdio>
#include <cuda_runtime.h>

constexpr int N = 4096;

// variant 0: select-then-add  (v = (a > v ? a : v) + l)        FSETP -> FSEL -> FADD
// variant 1: max-then-add     (v = fmaxf(a, v) + l)             FMNMX -> FADD
// variant 2: add-then-select  (v = a > v ? a + l : v + l)       (FSETP | FADD,FADD) -> FSEL
// variant 3: shuffle only     (v = shfl_up(v))
// variant 4: variant 0 with a = shfl_up(v, 1) every step (R = 1 lockstep systolic step)
// variant 5: variant 2 with shuffle
template <int V>
__global__ void chain(float *out, const float *in, long long *cycles) {
    float v = in[threadIdx.x], a = in[32 + threadIdx.x], l = in[64 + threadIdx.x];
    unsigned bits = 0;
    long long t0 = clock64();
#pragma unroll 32
    for (int i = 0; i < N; ++i) {
        if (V == 0) { bool p = a > v; v = (p ? a : v) + l; bits |= p ? (1u << (i & 31)) : 0u; }
        if (V == 1) { v = fmaxf(a, v) + l; }
        if (V == 2) { bool p = a > v; float x = a + l, y = v + l; v = p ? x : y; bits |= p ? (1u << (i & 31)) : 0u; }
        if (V == 3) { v = __shfl_up_sync(0xffffffffu, v, 1); }
        if (V == 4) { float u = __shfl_up_sync(0xffffffffu, v, 1); bool p = u > v; v = (p ? u : v) + l; bits |= p ? (1u << (i & 31)) : 0u; }
        if (V == 5) { float u = __shfl_up_sync(0xffffffffu, v, 1); bool p = u > v; float x = u + l, y = v + l; v = p ? x : y; bits |= p ? (1u << (i & 31)) : 0u; }
    }
    long long t1 = clock64();
    out[threadIdx.x] = v + __uint_as_float(bits);
    if (threadIdx.x == 0) cycles[V] = t1 - t0;
}

// R rows per lane, lockstep (the kernel's structure), select-then-add
template <int R, int V>
__global__ void systolic(float *out, const float *in, long long *cycles, int slot) {
    float v[R], l[R];
    unsigned bits[R];
    for (int i = 0; i < R; ++i) { v[i] = in[threadIdx.x + i]; l[i] = in[64 + threadIdx.x + i]; bits[i] = 0; }
    long long t0 = clock64();
#pragma unroll 8
    for (int s = 0; s < N; ++s) {
        float up = __shfl_up_sync(0xffffffffu, v[R - 1], 1);
#pragma unroll
        for (int i = R - 1; i >= 0; --i) {
            float a = i ? v[i - 1] : up;
            bool p = a > v[i];
            if (V == 0) v[i] = (p ? a : v[i]) + l[i];
            else { float x = a + l[i], y = v[i] + l[i]; v[i] = p ? x : y; }
            bits[i] |= p ? (1u << (s & 31)) : 0u;
        }
    }
    long long t1 = clock64();
    float acc = 0;
    for (int i = 0; i < R; ++i) acc += v[i] + __uint_as_float(bits[i]);
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[slot] = t1 - t0;
}

int main() {
    float *in, *out;
    long long *cyc, h[32] = {0};
    cudaMalloc(&in, 1024 * 4);
    cudaMalloc(&out, 1024 * 4);
    cudaMalloc(&cyc, 32 * 8);
    float hin[1024];
    for (int i = 0; i < 1024; ++i) hin[i] = -1.0f - 0.37f * (i % 17);
    cudaMemcpy(in, hin, sizeof hin, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 2; ++rep) {
        chain<0><<<1, 32>>>(out, in, cyc);
        chain<1><<<1, 32>>>(out, in, cyc);
        chain<2><<<1, 32>>>(out, in, cyc);
        chain<3><<<1, 32>>>(out, in, cyc);
        chain<4><<<1, 32>>>(out, in, cyc);
        chain<5><<<1, 32>>>(out, in, cyc);
        systolic<1, 0><<<1, 32>>>(out, in, cyc, 8);
        systolic<3, 0><<<1, 32>>>(out, in, cyc, 9);
        systolic<5, 0><<<1, 32>>>(out, in, cyc, 10);
        systolic<7, 0><<<1, 32>>>(out, in, cyc, 11);
        systolic<1, 1><<<1, 32>>>(out, in, cyc, 12);
        systolic<3, 1><<<1, 32>>>(out, in, cyc, 13);
        systolic<5, 1><<<1, 32>>>(out, in, cyc, 14);
        systolic<7, 1><<<1, 32>>>(out, in, cyc, 15);
    }
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    const char *names[] = {"select->add", "fmax->add", "add->select", "shfl only", "shfl+select->add", "shfl+add->select"};
    for (int i = 0; i < 6; ++i) printf("%-20s %.2f cycles/step\n", names[i], (double)h[i] / N);
    for (int v = 0; v < 2; ++v)
        for (int r = 0; r < 4; ++r) printf("systolic R=%d %s: %.2f cycles/step\n", 2 * r + 1, v ? "add->select" : "select->add", (double)h[8 + 4 * v + r] / N);
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

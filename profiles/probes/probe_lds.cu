// probe_lds.cu -- cost of LDS.128 / LDS.64 / LDS.32 per warp instruction on B200 for the address
// patterns of a register-tiled FFMA contraction (broadcast groups), 4 and 8 warps per SM.
#include <cstdio>
#include <cuda_runtime.h>

template <int VEC>   // floats per lane per load: 1, 2, 4
__global__ void lds(float *out, int iters, long long *cyc, int groups, int stride_floats) {
    __shared__ __align__(16) float sm[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
    __syncthreads();
    // lane -> group id (lanes of a group read the same address)
    const int lane = threadIdx.x & 31;
    const int gid = lane % groups;
    const float *p = sm + gid * stride_floats + (threadIdx.x >> 5) * 512;
    float s = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float *q = p + ((k * VEC * 32) & 255);
            if (VEC == 4) { float4 v = *reinterpret_cast<const float4 *>(q); s += v.x + v.y + v.z + v.w; }
            if (VEC == 2) { float2 v = *reinterpret_cast<const float2 *>(q); s += v.x + v.y; }
            if (VEC == 1) { s += *q; }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    float *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    struct { int groups, stride; const char *name; } pat[] = {
        {1, 0, "all lanes same address"}, {4, 4, "4 groups, contiguous 16B"}, {4, 64, "4 groups, 256B apart"},
        {8, 4, "8 groups, contiguous 16B (128B)"}, {8, 8, "8 groups, 32B apart"}, {32, 4, "32 distinct contiguous"}};
    for (int warps : {4, 8}) {
        for (auto &pt : pat) {
            lds<4><<<148, warps * 32>>>(out, iters, cyc, pt.groups, pt.stride);
            cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            double c4 = (double)h / (iters * 16.0 * warps);
            lds<2><<<148, warps * 32>>>(out, iters, cyc, pt.groups, pt.stride / 2 ? pt.stride / 2 : 0);
            cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            double c2 = (double)h / (iters * 16.0 * warps);
            lds<1><<<148, warps * 32>>>(out, iters, cyc, pt.groups, pt.stride / 4 ? pt.stride / 4 : 0);
            cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            double c1 = (double)h / (iters * 16.0 * warps);
            printf("%d warps/SM  %-34s  SM cycles per warp-instr: LDS.128 %.2f  LDS.64 %.2f  LDS.32 %.2f\n", warps, pt.name, c4, c2, c1);
        }
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
}

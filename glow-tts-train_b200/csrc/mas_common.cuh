// mas_common.cuh -- shared helpers for the sm_100a alignment kernels (no torch, no host sync).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "mas_b200.h"

namespace mas {

// Last failing CUDA call on this host thread (reported through mas_b200_last_cuda_error()).
extern thread_local int g_last_cuda_error;

inline int cuda_fail(cudaError_t e) {
    g_last_cuda_error = static_cast<int>(e);
    return MAS_ERR_CUDA;
}

#define MAS_CUDA_TRY(expr)                                   \
    do {                                                     \
        cudaError_t mas_e_ = (expr);                         \
        if (mas_e_ != cudaSuccess) return ::mas::cuda_fail(mas_e_); \
    } while (0)

constexpr int kWarp = 32;

__host__ __device__ constexpr int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ constexpr size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Per-utterance valid sizes as the kernels use them, after clamping into the defined domain:
//   tx in [0, T_x], ty in [0, T_y], and tx <= ty (tokens beyond the frame count can never be
//   visited; the reference has undefined behaviour there).
struct Lengths {
    int tx;
    int ty;
};

__device__ __forceinline__ Lengths clamp_lengths(int tx, int ty, int T_x, int T_y) {
    tx = max(0, min(tx, T_x));
    ty = max(0, min(ty, T_y));
    if (tx == 0 || ty == 0) {
        tx = 0;
        ty = 0;
    }
    if (tx > ty) tx = ty;
    return Lengths{tx, ty};
}

// Exclusive prefix sum of `n` ints (an utterance's durations: the first frame of every token) into
// shared memory by the whole CTA: every thread sums a contiguous segment, the segment totals are
// scanned by warp shuffles.  `s_warp`: >= 32 ints of shared scratch.  Ends with a CTA barrier.
__device__ __forceinline__ void block_exclusive_scan(const int32_t *__restrict__ in, int n, int *s_out, int *s_warp) {
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int per = (n + nthr - 1) / nthr, lo = min(n, tid * per), hi = min(n, lo + per);
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += in[i];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < (nthr + 31) / 32 ? s_warp[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += v;
        }
        s_warp[lane] = w;                                    // inclusive over the warps
    }
    __syncthreads();
    int run = incl - sum + (warp > 0 ? s_warp[warp - 1] : 0);
    for (int i = lo; i < hi; ++i) {
        s_out[i] = run;
        run += in[i];
    }
    __syncthreads();
}

}  // namespace mas

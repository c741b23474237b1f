// mas_train.cu -- SURVEY.md 8(f) ranks 2 and 3, the step around the path:
//   * duration_loss(logw, logw_, lengths) (glow_tts_train/utils.py:26-28, train.py:125) straight from
//     the integer durations the alignment kernels emit: logw_ = log(1e-8 + durations) * x_mask
//     (models.py:393) is formed on the fly, forward and backward;
//   * clip_grad_value_(parameters, clip) (utils.py:118-132, train.py:141/145) without its host
//     synchronisation per parameter tensor: the reference calls `.item()` on every gradient's norm,
//     i.e. one device sync per tensor and step (~300 tensors); here ONE launch walks a table of
//     gradient chunks, clamps in place and leaves the total norm on the device.
// Element-wise, HBM-bound, deterministic (fixed summation order).
#include "mas_kernels.cuh"

namespace mas {
namespace train {

constexpr int kThreads = 256;

// block-wide sum in a fixed order (warp shuffles, then warp 0 over the warps' sums); result in thread 0
__device__ __forceinline__ double block_sum(double v, double *s_warp) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = threadIdx.x < (blockDim.x >> 5) ? s_warp[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    }
    return v;
}

// ---- duration loss: one CTA (B x T_x is a few thousand elements) ----
// out2[0] = sum((logw - logw_)^2) / sum(x_len), out2[1] = 2 / sum(x_len)
__global__ void __launch_bounds__(1024) duration_loss_kernel(const float *__restrict__ logw, const int32_t *__restrict__ dur,
                                                             const int32_t *__restrict__ x_len, float *__restrict__ out2, int B,
                                                             int T_x) {
    __shared__ double s_warp[32];
    __shared__ double s_num;
    double acc = 0.0;
    for (int i = threadIdx.x; i < B * T_x; i += blockDim.x) {
        const int b = i / T_x, x = i - b * T_x;
        const float target = (x < x_len[b]) ? logf(1e-8f + (float)dur[i]) : 0.f;     // models.py:393
        const float d = logw[i] - target;
        acc += (double)(d * d);                                                       // utils.py:27
    }
    const double num = block_sum(acc, s_warp);
    if (threadIdx.x == 0) s_num = num;
    __syncthreads();
    double len = 0.0;
    for (int b = threadIdx.x; b < B; b += blockDim.x) len += (double)x_len[b];
    len = block_sum(len, s_warp);
    if (threadIdx.x == 0) {
        out2[0] = (float)(s_num / len);
        out2[1] = (float)(2.0 / len);
    }
}

// dlogw = scale * (logw - logw_), scale = upstream gradient * 2 / sum(x_len) (a device scalar)
__global__ void __launch_bounds__(kThreads) duration_loss_backward_kernel(const float *__restrict__ logw, const int32_t *__restrict__ dur,
                                                                          const int32_t *__restrict__ x_len,
                                                                          const float *__restrict__ scale, float *__restrict__ dlogw,
                                                                          int T_x, int total) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= total) return;
    const int b = i / T_x, x = i - b * T_x;
    const float target = (x < x_len[b]) ? logf(1e-8f + (float)dur[i]) : 0.f;
    dlogw[i] = __ldg(scale) * (logw[i] - target);
}

// ---- gradient clipping by value over a table of chunks ----
// chunk c: `count[c]` floats at `ptr[c]` (a slice of one gradient tensor).  Clamps in place exactly
// like torch.clamp_ (NaN stays NaN) and leaves the chunk's sum of squares (of the UNclamped values,
// utils.py:127-130) in partial[c].
__global__ void __launch_bounds__(kThreads) clip_chunks_kernel(float *const *__restrict__ ptr, const int32_t *__restrict__ count,
                                                               float clip, double *__restrict__ partial) {
    __shared__ double s_warp[32];
    float *p = ptr[blockIdx.x];
    const int n = count[blockIdx.x];
    double acc = 0.0;
    const bool vec = (reinterpret_cast<uintptr_t>(p) & 15) == 0;
    const int n4 = vec ? n >> 2 : 0;
    float4 *p4 = reinterpret_cast<float4 *>(p);
    for (int i = threadIdx.x; i < n4; i += kThreads) {
        float4 g = p4[i];
        float s = g.x * g.x;
        s = fmaf(g.y, g.y, s), s = fmaf(g.z, g.z, s), s = fmaf(g.w, g.w, s);
        acc += (double)s;
        g.x = g.x < -clip ? -clip : (g.x > clip ? clip : g.x);
        g.y = g.y < -clip ? -clip : (g.y > clip ? clip : g.y);
        g.z = g.z < -clip ? -clip : (g.z > clip ? clip : g.z);
        g.w = g.w < -clip ? -clip : (g.w > clip ? clip : g.w);
        p4[i] = g;
    }
    for (int i = 4 * n4 + threadIdx.x; i < n; i += kThreads) {
        const float g = p[i];
        acc += (double)(g * g);
        p[i] = g < -clip ? -clip : (g > clip ? clip : g);
    }
    acc = block_sum(acc, s_warp);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// total norm = sqrt(sum over chunks, in order); one CTA
__global__ void __launch_bounds__(1024) clip_finish_kernel(const double *__restrict__ partial, int nchunks, float *__restrict__ total_norm) {
    __shared__ double s_warp[32];
    // thread t sums chunks t, t + 1024, ... (fixed order), then the block sum (fixed order)
    double acc = 0.0;
    for (int c = threadIdx.x; c < nchunks; c += blockDim.x) acc += partial[c];
    acc = block_sum(acc, s_warp);
    if (threadIdx.x == 0) total_norm[0] = (float)sqrt(acc);
}

}  // namespace train

int launch_duration_loss(const float *logw, const int32_t *durations, const int32_t *x_len, float *out2, int B, int T_x,
                         cudaStream_t stream) {
    train::duration_loss_kernel<<<1, 1024, 0, stream>>>(logw, durations, x_len, out2, B, T_x);
    MAS_CUDA_TRY(cudaGetLastError());
    return MAS_OK;
}

int launch_duration_loss_backward(const float *logw, const int32_t *durations, const int32_t *x_len, const float *scale, float *dlogw,
                                  int B, int T_x, cudaStream_t stream) {
    const int total = B * T_x;
    train::duration_loss_backward_kernel<<<ceil_div(total, train::kThreads), train::kThreads, 0, stream>>>(logw, durations, x_len, scale,
                                                                                                             dlogw, T_x, total);
    MAS_CUDA_TRY(cudaGetLastError());
    return MAS_OK;
}

int launch_clip_grad_value(float *const *chunk_ptr, const int32_t *chunk_count, int nchunks, float clip, double *partial,
                           float *total_norm, cudaStream_t stream) {
    train::clip_chunks_kernel<<<nchunks, train::kThreads, 0, stream>>>(chunk_ptr, chunk_count, clip, partial);
    MAS_CUDA_TRY(cudaGetLastError());
    train::clip_finish_kernel<<<1, 1024, 0, stream>>>(partial, nchunks, total_norm);
    MAS_CUDA_TRY(cudaGetLastError());
    return MAS_OK;
}

}  // namespace mas

// mas_loss.cu -- SURVEY.md 8(f) rank 2: the maximum-likelihood loss of the aligned prior
// (glow_tts_train/utils.py:14-23, called at train.py:124) computed straight from the token-level
// prior and the frame->token map, without materialising z_m / z_logs (models.py:383-392):
//
//   loss = ( sum_{b,d,y} [ logs + 0.5 exp(-2 logs) (z - m)^2 ]  -  sum_b logdet[b] ) / (D sum_b y_len[b])
//          + 0.5 log(2 pi)
//   with m = x_m[b,d,tok[b,y]], logs = x_logs[b,d,tok[b,y]]  (both 0 where tok < 0: the zero columns of
//   attn give z_m = z_logs = 0 there, and the reference sums over padded frames too).
//
// Backward (one upstream scalar g, s = g / (D sum y_len)):
//   dz[b,d,y]      =  s exp(-2 logs) (z - m)
//   dx_m[b,d,x]    = -s sum_{y in run(x)} exp(-2 logs) (z - m)
//   dx_logs[b,d,x] =  s sum_{y in run(x)} (1 - exp(-2 logs) (z - m)^2)
//   dlogdet[b]     = -s
// HBM-bound element-wise / segmented kernels.  Sums are deterministic: fixed per-thread order,
// tree per block, fp64 over the blocks.
#include "mas_kernels.cuh"

namespace mas {
namespace loss {

constexpr int kFrames = 128;        // frames per CTA of the forward kernel

// grid: (ceil(T_y/128), B); thread = one frame, all channels.  partial[b * gridDim.x + blockIdx.x]
__global__ void __launch_bounds__(kFrames) mle_partial_kernel(const float *__restrict__ z, const float *__restrict__ x_m,
                                                              const float *__restrict__ x_logs, const int32_t *__restrict__ tok,
                                                              double *__restrict__ partial, int D, int T_x, int T_y) {
    const int b = blockIdx.y, y = blockIdx.x * kFrames + threadIdx.x;
    float acc = 0.f;
    if (y < T_y) {
        const int t = __ldg(tok + (int64_t)b * T_y + y);
        const bool on = t >= 0 && t < T_x;
        const float *zc = z + (int64_t)b * D * T_y + y;
        const float *mc = x_m + (int64_t)b * D * T_x + (on ? t : 0);
        const float *lc = x_logs ? x_logs + (int64_t)b * D * T_x + (on ? t : 0) : nullptr;
#pragma unroll 4
        for (int d = 0; d < D; ++d) {
            const float zv = __ldg(zc + (int64_t)d * T_y);
            const float m = on ? __ldg(mc + (int64_t)d * T_x) : 0.f;
            const float ls = (on && lc) ? __ldg(lc + (int64_t)d * T_x) : 0.f;
            const float e = zv - m;
            acc += ls + 0.5f * (expf(-2.0f * ls) * (e * e));          // utils.py:15-17
        }
    }
    __shared__ float s_warp[kFrames / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < kFrames / 32; ++w) s += (double)s_warp[w];
        partial[(int64_t)b * gridDim.x + blockIdx.x] = s;
    }
}

// one CTA: fixed-order fp64 sum of the partials, then utils.py:18-22.  out[0] = loss, out[1] = 1 / denominator
__global__ void __launch_bounds__(256) mle_finish_kernel(const double *__restrict__ partial, int n, const float *__restrict__ logdet,
                                                         const int32_t *__restrict__ y_len, int B, int D, float *__restrict__ out) {
    __shared__ double s_sum[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += partial[i];
    s_sum[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double total = s_sum[0], frames = 0.0;
        for (int b = 0; b < B; ++b) {
            if (logdet) total -= (double)logdet[b];                       // utils.py:18
            frames += (double)y_len[b];
        }
        const double denom = frames * (double)D;                            // utils.py:19-21: sum(ones_like(z) * mask)
        out[0] = (float)(total / denom + 0.91893853320467274178);          // utils.py:22
        out[1] = (float)(1.0 / denom);
    }
}

// grid: (ceil(T_y/256), D, B)
__global__ void __launch_bounds__(256) mle_grad_z_kernel(const float *__restrict__ z, const float *__restrict__ x_m,
                                                         const float *__restrict__ x_logs, const int32_t *__restrict__ tok,
                                                         const float *__restrict__ scale, float *__restrict__ dz, int D, int T_x,
                                                         int T_y) {
    const int b = blockIdx.z, d = blockIdx.y, y = blockIdx.x * 256 + threadIdx.x;
    if (y >= T_y) return;
    const int t = __ldg(tok + (int64_t)b * T_y + y);
    const bool on = t >= 0 && t < T_x;
    const int64_t row = ((int64_t)b * D + d);
    const float m = on ? __ldg(x_m + row * T_x + t) : 0.f;
    const float ls = (on && x_logs) ? __ldg(x_logs + row * T_x + t) : 0.f;
    dz[row * T_y + y] = __ldg(scale) * (expf(-2.0f * ls) * (__ldg(z + row * T_y + y) - m));
}

// grid: (ceil(T_x/128), D, B); one thread per (channel, token): its run of frames, ascending
__global__ void __launch_bounds__(128) mle_grad_tokens_kernel(const float *__restrict__ z, const float *__restrict__ x_m,
                                                              const float *__restrict__ x_logs, const int32_t *__restrict__ dur,
                                                              const float *__restrict__ scale, float *__restrict__ dx_m,
                                                              float *__restrict__ dx_logs, int D, int T_x, int T_y) {
    extern __shared__ int s_start[];                      // exclusive prefix sum of this utterance's durations
    const int b = blockIdx.z, d = blockIdx.y;
    const int32_t *du = dur + (int64_t)b * T_x;
    __shared__ int s_warp[32];
    block_exclusive_scan(du, T_x, s_start, s_warp);
    const int x = blockIdx.x * 128 + threadIdx.x;
    if (x >= T_x) return;
    const int64_t row = (int64_t)b * D + d;
    const int n = du[x], y0 = s_start[x];
    const float m = __ldg(x_m + row * T_x + x);
    const float ls = x_logs ? __ldg(x_logs + row * T_x + x) : 0.f;
    const float r = expf(-2.0f * ls);
    const float *zr = z + row * T_y;
    float a = 0.f, q = 0.f;
    for (int k = 0; k < n && y0 + k < T_y; ++k) {
        const float e = __ldg(zr + y0 + k) - m;
        a += r * e;
        q += r * (e * e);
    }
    const float s = __ldg(scale);
    dx_m[row * T_x + x] = -s * a;
    if (dx_logs) dx_logs[row * T_x + x] = s * ((float)n - q);
}

}  // namespace loss

size_t mle_loss_workspace_bytes(int B, int T_y) { return align_up((size_t)B * ceil_div(T_y, loss::kFrames) * sizeof(double), 256); }

int launch_mle_loss(const float *z, const float *x_m, const float *x_logs, const int32_t *frame_token, const float *logdet,
                    const int32_t *y_len, float *out2, void *workspace, int B, int D, int T_x, int T_y, cudaStream_t stream) {
    const int nb = ceil_div(T_y, loss::kFrames);
    double *partial = static_cast<double *>(workspace);
    loss::mle_partial_kernel<<<dim3(nb, B), loss::kFrames, 0, stream>>>(z, x_m, x_logs, frame_token, partial, D, T_x, T_y);
    MAS_CUDA_TRY(cudaGetLastError());
    loss::mle_finish_kernel<<<1, 256, 0, stream>>>(partial, nb * B, logdet, y_len, B, D, out2);
    MAS_CUDA_TRY(cudaGetLastError());
    return MAS_OK;
}

int launch_mle_loss_backward(const float *z, const float *x_m, const float *x_logs, const int32_t *frame_token,
                             const int32_t *durations, const float *scale, float *dz, float *dx_m, float *dx_logs, int B, int D,
                             int T_x, int T_y, cudaStream_t stream) {
    if (dz) {
        loss::mle_grad_z_kernel<<<dim3(ceil_div(T_y, 256), D, B), 256, 0, stream>>>(z, x_m, x_logs, frame_token, scale, dz, D, T_x, T_y);
        MAS_CUDA_TRY(cudaGetLastError());
    }
    if (dx_m) {
        loss::mle_grad_tokens_kernel<<<dim3(ceil_div(T_x, 128), D, B), 128, (size_t)T_x * sizeof(int), stream>>>(
            z, x_m, x_logs, durations, scale, dx_m, dx_logs, D, T_x, T_y);
        MAS_CUDA_TRY(cudaGetLastError());
    }
    return MAS_OK;
}

}  // namespace mas

// mas_expand.cu -- the path's consumers, SURVEY.md 8(f) rank 1: expanding the token-level prior to
// frame level and the target durations (glow_tts_train/models.py:383-393).
//
// The reference multiplies the dense 0/1 path with x_m / x_logs (two [B,T_y,T_x] x [B,T_x,D] matmuls,
// 1.02 GFLOP each at B=32, 200x1000, D=80) -- a GATHER by token index written as a GEMM -- and takes
// log(1e-8 + row sums) for the durations.  With the frame->token map and the integer durations the
// alignment kernels already emit, the same results are a gather (forward), a segmented sum over each
// token's run of frames (backward w.r.t. x_m / x_logs) and a log:
//   z_m[b,d,y]   = x_m[b,d,tok[b,y]]      (0 where tok < 0)           models.py:383-387 (exact: 1*x + 0*...)
//   z_logs       likewise                                              models.py:388-392
//   logw_[b,x]   = log(1e-8 + dur[b,x]) * (x < x_len[b])              models.py:393
//   dx[b,d,x]    = sum_{y : tok[b,y] == x} dz[b,d,y]                  (autograd of the matmul)
// HBM-bound element-wise kernels: coalesced over frames / tokens, grid sized in multiples of the SMs.
#include "mas_kernels.cuh"

namespace mas {
namespace expand {

// grid: (ceil(T_y/256), D, B)
__global__ void __launch_bounds__(256) gather_kernel(const float *__restrict__ x, const int32_t *__restrict__ tok,
                                                     float *__restrict__ z, int D, int T_x, int T_y) {
    const int b = blockIdx.z, d = blockIdx.y, y = blockIdx.x * 256 + threadIdx.x;
    if (y >= T_y) return;
    const int t = __ldg(tok + (int64_t)b * T_y + y);
    const float *row = x + ((int64_t)b * D + d) * T_x;
    z[((int64_t)b * D + d) * T_y + y] = (t >= 0 && t < T_x) ? __ldg(row + t) : 0.f;
}

// Segmented sum over the run of frames of each token: frames of token x are [start[x], start[x] + dur[x]).
// grid: (ceil(T_x/128), D, B); one thread per (d, token): runs average T_y/T_x frames.
__global__ void __launch_bounds__(128) scatter_kernel(const float *__restrict__ dz, const int32_t *__restrict__ dur,
                                                      float *__restrict__ dx, int D, int T_x, int T_y) {
    extern __shared__ int s_start[];                      // exclusive prefix sum of the durations of this utterance
    const int b = blockIdx.z, d = blockIdx.y;
    const int32_t *du = dur + (int64_t)b * T_x;
    // every CTA recomputes the prefix (T_x <= 2048 ints), in parallel: cheap next to the frame traffic
    __shared__ int s_warp[32];
    block_exclusive_scan(du, T_x, s_start, s_warp);
    const int x = blockIdx.x * 128 + threadIdx.x;
    if (x >= T_x) return;
    const int n = du[x], y0 = s_start[x];
    const float *row = dz + ((int64_t)b * D + d) * T_y;
    float acc = 0.f;
    for (int k = 0; k < n && y0 + k < T_y; ++k) acc += __ldg(row + y0 + k);   // ascending frames: deterministic
    dx[((int64_t)b * D + d) * T_x + x] = acc;
}

__global__ void __launch_bounds__(256) logw_kernel(const int32_t *__restrict__ dur, const int32_t *__restrict__ x_len,
                                                   float *__restrict__ logw, int T_x, int total) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const int b = i / T_x, x = i - b * T_x;
    logw[i] = (x < x_len[b]) ? logf(1e-8f + (float)dur[i]) : 0.f;
}

}  // namespace expand

int launch_expand_gather(const float *x, const int32_t *frame_token, float *z, int B, int D, int T_x, int T_y, cudaStream_t stream) {
    if (B == 0 || D == 0 || T_y == 0) return MAS_OK;
    dim3 grid(ceil_div(T_y, 256), D, B);
    expand::gather_kernel<<<grid, 256, 0, stream>>>(x, frame_token, z, D, T_x, T_y);
    MAS_CUDA_TRY(cudaGetLastError());
    return MAS_OK;
}

int launch_expand_scatter(const float *dz, const int32_t *durations, float *dx, int B, int D, int T_x, int T_y, cudaStream_t stream) {
    if (B == 0 || D == 0 || T_x == 0) return MAS_OK;
    dim3 grid(ceil_div(T_x, 128), D, B);
    expand::scatter_kernel<<<grid, 128, (size_t)T_x * sizeof(int), stream>>>(dz, durations, dx, D, T_x, T_y);
    MAS_CUDA_TRY(cudaGetLastError());
    return MAS_OK;
}

int launch_logw(const int32_t *durations, const int32_t *x_len, float *logw, int B, int T_x, cudaStream_t stream) {
    if (B == 0 || T_x == 0) return MAS_OK;
    const int total = B * T_x;
    expand::logw_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(durations, x_len, logw, T_x, total);
    MAS_CUDA_TRY(cudaGetLastError());
    return MAS_OK;
}

}  // namespace mas

// ---------------------------------------------------------------------------------------------
// SURVEY.md 8(f) rank 4, the inference-side analogue: durations -> dense path
// (glow_tts_train/utils.py:99-115 `generate_path`, called at models.py:340):
//   cum[x] = duration[b,0] + ... + duration[b,x]       (fp32, ascending; exact for the integer-valued
//                                                        ceil(w) the model passes, < 2^24)
//   path[b,x,y] = ((y < cum[x]) - (y < cum[x-1])) * mask[b,x,y]          (cum[-1] = 0)
// One CTA per (utterance, 32 tokens): prefix sums in shared memory, then every warp writes token
// rows with 16-byte stores.  HBM-bound: 4 B/cell written (+ 4 B/cell of mask read, as the reference does).
// ---------------------------------------------------------------------------------------------
namespace mas {
namespace expand {

constexpr int kGenRows = 32;

__global__ void __launch_bounds__(256) generate_path_kernel(const float *__restrict__ duration, const float *__restrict__ mask,
                                                            int64_t ms_b, int64_t ms_x, int64_t ms_y, float *__restrict__ path,
                                                            int T_x, int T_y) {
    extern __shared__ float s_cum[];                      // [T_x + 1], s_cum[0] = 0
    const int b = blockIdx.y, x0 = blockIdx.x * kGenRows;
    const float *du = duration + (int64_t)b * T_x;
    if (threadIdx.x == 0) {
        float run = 0.f;
        s_cum[0] = 0.f;
        const int upto = min(T_x, x0 + kGenRows);
        for (int x = 0; x < upto; ++x) {                  // torch.cumsum order; only the prefix this CTA needs
            run += du[x];
            s_cum[x + 1] = run;
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int xr = warp; xr < kGenRows; xr += 8) {
        const int x = x0 + xr;
        if (x >= T_x) break;
        const float lo = s_cum[x], hi = s_cum[x + 1];
        float *row = path + ((int64_t)b * T_x + x) * T_y;
        const float *mrow = mask + (int64_t)b * ms_b + (int64_t)x * ms_x;
        for (int y = lane; y < T_y; y += 32) {
            const float fy = (float)y;                    // sequence_mask compares arange in the length's dtype (utils.py:52-56)
            const float v = (float)(fy < hi) - (float)(fy < lo);
            row[y] = v * mrow[(int64_t)y * ms_y];
        }
    }
}

}  // namespace expand

int launch_generate_path(const float *duration, const float *mask, int64_t ms_b, int64_t ms_x, int64_t ms_y, float *path, int B,
                         int T_x, int T_y, cudaStream_t stream) {
    if (B == 0 || T_x == 0 || T_y == 0) return MAS_OK;
    dim3 grid(ceil_div(T_x, expand::kGenRows), B);
    expand::generate_path_kernel<<<grid, 256, (size_t)(T_x + 1) * sizeof(float), stream>>>(duration, mask, ms_b, ms_x, ms_y, path,
                                                                                             T_x, T_y);
    MAS_CUDA_TRY(cudaGetLastError());
    return MAS_OK;
}

}  // namespace mas

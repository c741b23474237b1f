// mas_api.cu -- the extern "C" surface declared in include/mas_b200.h: argument validation,
// workspace carving and kernel dispatch.  No torch, no exceptions, no host synchronisation
// (except mas_b200_maximum_path_host_i32, which is synchronous by contract).
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "mas_kernels.cuh"

namespace mas {
thread_local int g_last_cuda_error = 0;
std::atomic<long long *> g_dbg_cycles{nullptr};
}
static std::atomic<int> g_force_unfused{0};   // testing hook: 0 = by estimate, 1 = the two kernels, 2 = the single launch

using namespace mas;

namespace mas {
int get_device_info(int dev, DeviceInfo &out) {
    static std::mutex mu;
    static DeviceInfo cache[64];
    static bool have[64] = {false};
    if (dev < 0 || dev >= 64) return MAS_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(mu);
    if (!have[dev]) {
        DeviceInfo d{};
        MAS_CUDA_TRY(cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        MAS_CUDA_TRY(cudaDeviceGetAttribute(&d.num_sms, cudaDevAttrMultiProcessorCount, dev));
        cache[dev] = d;
        have[dev] = true;
    }
    out = cache[dev];
    return MAS_OK;
}
}  // namespace mas

namespace {

size_t path_workspace_bytes(int B, int T_x, int T_y) {
    const size_t a = path_simple_workspace_bytes(B, T_x, T_y), b = path_systolic_workspace_bytes(B, T_x, T_y);
    return (a > b ? a : b) + mask_flag_bytes(B);       // + the per-utterance "mask is not all-ones" flags (mas_mask.cu)
}

struct FusedWorkspace {
    size_t logp_bytes, path_ws_bytes;
};

FusedWorkspace fused_ws(int B, int T_x, int T_y) {
    FusedWorkspace w;
    w.logp_bytes = align_up((size_t)B * T_x * T_y * sizeof(float), 256);
    w.path_ws_bytes = path_workspace_bytes(B, T_x, T_y);
    return w;
}

bool shape_ok(int B, int T_x, int T_y) {
    return B >= 0 && T_x >= 0 && T_y >= 0 && T_x <= MAS_B200_MAX_TOKENS && T_y <= MAS_B200_MAX_FRAMES;
}

}  // namespace

extern "C" {

int mas_b200_abi_version(void) { return MAS_B200_ABI_VERSION; }

const char *mas_b200_status_string(int status) {
    switch (status) {
        case MAS_OK: return "ok";
        case MAS_ERR_INVALID_ARGUMENT: return "invalid argument";
        case MAS_ERR_UNSUPPORTED_SHAPE: return "unsupported shape";
        case MAS_ERR_WORKSPACE_TOO_SMALL: return "workspace too small";
        case MAS_ERR_NO_DEVICE: return "no sm_100 CUDA device";
        case MAS_ERR_CUDA: return "CUDA error";
        case MAS_ERR_BAD_LENGTHS: return "bad utterance lengths";
        default: return "unknown status";
    }
}

int mas_b200_last_cuda_error(void) { return g_last_cuda_error; }

void mas_b200_debug_set_cycle_buffer(void *device_buffer) { g_dbg_cycles.store(static_cast<long long *>(device_buffer)); }

void mas_b200_debug_force_cluster(int ctas_per_utterance) { path_systolic_force_cluster(ctas_per_utterance); }

void mas_b200_debug_force_unfused(int mode) { g_force_unfused.store(mode); }

int mas_b200_device_ok(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return MAS_ERR_NO_DEVICE;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
        return MAS_ERR_NO_DEVICE;
    return major == 10 ? MAS_OK : MAS_ERR_NO_DEVICE;
}

size_t mas_b200_workspace_bytes(int B, int T_x, int T_y) {
    if (!shape_ok(B, T_x, T_y)) return 0;
    return path_workspace_bytes(B, T_x, T_y);
}

size_t mas_b200_fused_workspace_bytes(int B, int D, int T_x, int T_y) {
    (void)D;
    if (!shape_ok(B, T_x, T_y)) return 0;
    FusedWorkspace w = fused_ws(B, T_x, T_y);
    const size_t two_kernels = w.logp_bytes + w.path_ws_bytes, one_launch = fused_workspace_bytes(B, D, T_x, T_y);
    return two_kernels > one_launch ? two_kernels : one_launch;
}

int mas_b200_maximum_path_f32(const float *value, int64_t value_stride_b, int64_t value_stride_x,
                              const int32_t *t_x, const int32_t *t_y,
                              const float *mask, int64_t mask_stride_b, int64_t mask_stride_x,
                              int64_t mask_stride_y,
                              float *path, int32_t *durations, int32_t *frame_token,
                              void *workspace, size_t workspace_bytes,
                              int B, int T_x, int T_y, float max_neg_val, mas_stream_t stream) {
    if (!shape_ok(B, T_x, T_y)) return (B < 0 || T_x < 0 || T_y < 0) ? MAS_ERR_INVALID_ARGUMENT : MAS_ERR_UNSUPPORTED_SHAPE;
    if (B == 0 || T_x == 0 || T_y == 0) return MAS_OK;
    if (value == nullptr || path == nullptr) return MAS_ERR_INVALID_ARGUMENT;
    if ((t_x == nullptr) != (t_y == nullptr)) return MAS_ERR_INVALID_ARGUMENT;
    if (t_x == nullptr && mask == nullptr) return MAS_ERR_INVALID_ARGUMENT;
    if (value_stride_x < T_y || value_stride_b < 0) return MAS_ERR_INVALID_ARGUMENT;
    PathParams p{};
    p.value = value;
    p.value_stride_b = value_stride_b;
    p.value_stride_x = value_stride_x;
    p.t_x = t_x;
    p.t_y = t_y;
    p.mask = mask;
    p.mask_stride_b = mask_stride_b;
    p.mask_stride_x = mask_stride_x;
    p.mask_stride_y = mask_stride_y;
    p.path = path;
    p.durations = durations;
    p.frame_token = frame_token;
    p.B = B;
    p.T_x = T_x;
    p.T_y = T_y;
    p.max_neg_val = max_neg_val;
    p.dbg_cycles = g_dbg_cycles.load();
    if (t_x == nullptr) {
        // lengths AND scores come through the mask (__init__.py:11,18-19): prove on the device that
        // value * mask == value on the valid rectangle; utterances where it is not are flagged and
        // computed from value * mask literally
        const size_t fb = mask_flag_bytes(B);
        if (workspace == nullptr || workspace_bytes < fb) return MAS_ERR_WORKSPACE_TOO_SMALL;
        int *flags = reinterpret_cast<int *>(static_cast<unsigned char *>(workspace) + workspace_bytes - fb);
        flags = reinterpret_cast<int *>(reinterpret_cast<uintptr_t>(flags) & ~(uintptr_t)3);
        int rcm = launch_mask_check(p, flags, static_cast<cudaStream_t>(stream));
        if (rcm != MAS_OK) return rcm;
        p.exact_flag = flags;
        workspace_bytes -= fb;
    }
    // the TMA-staged systolic kernel when shape/alignment allow, else the generic kernel
    int rc = launch_path_systolic(p, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
    if (rc != MAS_ERR_UNSUPPORTED_SHAPE) return rc;
    return launch_path_simple(p, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int mas_b200_logp_f32(const float *x_m, const float *x_logs, const float *z, float *logp,
                      int B, int D, int T_x, int T_y, mas_stream_t stream) {
    if (B < 0 || D < 0 || T_x < 0 || T_y < 0) return MAS_ERR_INVALID_ARGUMENT;
    if (!shape_ok(B, T_x, T_y) || D > MAS_B200_MAX_CHANNELS) return MAS_ERR_UNSUPPORTED_SHAPE;
    if (B == 0 || T_x == 0 || T_y == 0) return MAS_OK;
    if (x_m == nullptr || z == nullptr || logp == nullptr) return MAS_ERR_INVALID_ARGUMENT;
    LogpParams p{x_m, x_logs, z, logp, B, D, T_x, T_y};
    return launch_logp(p, static_cast<cudaStream_t>(stream));
}

int mas_b200_fused_maximum_path_f32(const float *x_m, const float *x_logs, const float *z,
                                    const int32_t *x_len, const int32_t *y_len,
                                    float *path, int32_t *durations, int32_t *frame_token,
                                    void *workspace, size_t workspace_bytes,
                                    int B, int D, int T_x, int T_y, float max_neg_val,
                                    mas_stream_t stream) {
    if (B < 0 || D < 0 || T_x < 0 || T_y < 0) return MAS_ERR_INVALID_ARGUMENT;
    if (!shape_ok(B, T_x, T_y) || D > MAS_B200_MAX_CHANNELS) return MAS_ERR_UNSUPPORTED_SHAPE;
    if (B == 0 || T_x == 0 || T_y == 0) return MAS_OK;
    if (x_m == nullptr || z == nullptr || path == nullptr || x_len == nullptr || y_len == nullptr)
        return MAS_ERR_INVALID_ARGUMENT;
    FusedWorkspace w = fused_ws(B, T_x, T_y);
    if (workspace == nullptr || workspace_bytes < mas_b200_fused_workspace_bytes(B, D, T_x, T_y)) return MAS_ERR_WORKSPACE_TOO_SMALL;
    // One launch, a cluster of CTAs per utterance, scores produced and consumed in shared memory
    // (mas_fused.cu) -- when its cost estimate for the shape beats the two kernels'.
    const int mode = g_force_unfused.load();
    if (mode != 1) {
        LogpParams lp{x_m, x_logs, z, nullptr, B, D, T_x, T_y};
        const int rc1 = launch_fused(lp, x_len, y_len, path, durations, frame_token, workspace, workspace_bytes, max_neg_val,
                                     mode == 2, static_cast<cudaStream_t>(stream));
        if (rc1 != MAS_ERR_UNSUPPORTED_SHAPE) return rc1;
    }
    // Shapes the single launch does not take (frame count not a multiple of 4, > 80 channels, > 1024
    // tokens): the same two programs back to back, scores staged in the workspace.
    float *logp = static_cast<float *>(workspace);
    int rc = mas_b200_logp_f32(x_m, x_logs, z, logp, B, D, T_x, T_y, stream);
    if (rc != MAS_OK) return rc;
    return mas_b200_maximum_path_f32(logp, (int64_t)T_x * T_y, T_y, x_len, y_len, nullptr, 0, 0, 0, path,
                                     durations, frame_token, static_cast<unsigned char *>(workspace) + w.logp_bytes,
                                     workspace_bytes - w.logp_bytes, B, T_x, T_y, max_neg_val, stream);
}

int mas_b200_expand_prior_f32(const float *x, const int32_t *frame_token, float *z, int B, int D, int T_x, int T_y,
                              mas_stream_t stream) {
    if (B < 0 || D < 0 || T_x < 0 || T_y < 0) return MAS_ERR_INVALID_ARGUMENT;
    if (!shape_ok(B, T_x, T_y)) return MAS_ERR_UNSUPPORTED_SHAPE;
    if (B == 0 || D == 0 || T_y == 0) return MAS_OK;
    if (!x || !frame_token || !z) return MAS_ERR_INVALID_ARGUMENT;
    return launch_expand_gather(x, frame_token, z, B, D, T_x, T_y, static_cast<cudaStream_t>(stream));
}

int mas_b200_expand_prior_backward_f32(const float *dz, const int32_t *durations, float *dx, int B, int D, int T_x, int T_y,
                                       mas_stream_t stream) {
    if (B < 0 || D < 0 || T_x < 0 || T_y < 0) return MAS_ERR_INVALID_ARGUMENT;
    if (!shape_ok(B, T_x, T_y)) return MAS_ERR_UNSUPPORTED_SHAPE;
    if (B == 0 || D == 0 || T_x == 0) return MAS_OK;
    if (!dz || !durations || !dx) return MAS_ERR_INVALID_ARGUMENT;
    return launch_expand_scatter(dz, durations, dx, B, D, T_x, T_y, static_cast<cudaStream_t>(stream));
}

int mas_b200_log_durations_f32(const int32_t *durations, const int32_t *x_len, float *logw, int B, int T_x, mas_stream_t stream) {
    if (B < 0 || T_x < 0) return MAS_ERR_INVALID_ARGUMENT;
    if (B == 0 || T_x == 0) return MAS_OK;
    if (!durations || !x_len || !logw) return MAS_ERR_INVALID_ARGUMENT;
    return launch_logw(durations, x_len, logw, B, T_x, static_cast<cudaStream_t>(stream));
}

int mas_b200_generate_path_f32(const float *duration, const float *mask, int64_t mask_stride_b, int64_t mask_stride_x,
                               int64_t mask_stride_y, float *path, int B, int T_x, int T_y, mas_stream_t stream) {
    if (B < 0 || T_x < 0 || T_y < 0) return MAS_ERR_INVALID_ARGUMENT;
    if (!shape_ok(B, T_x, T_y)) return MAS_ERR_UNSUPPORTED_SHAPE;
    if (B == 0 || T_x == 0 || T_y == 0) return MAS_OK;
    if (!duration || !mask || !path) return MAS_ERR_INVALID_ARGUMENT;
    return launch_generate_path(duration, mask, mask_stride_b, mask_stride_x, mask_stride_y, path, B, T_x, T_y,
                                static_cast<cudaStream_t>(stream));
}

int mas_b200_debug_path_plan(int B, int T_x, int T_y, int max_smem, int num_sms, int32_t *out8) {
    if (B <= 0 || T_x <= 0 || T_y <= 0 || max_smem <= 0 || num_sms <= 0 || !out8) return MAS_ERR_INVALID_ARGUMENT;
    return debug_path_plan(B, T_x, T_y, max_smem, num_sms, out8) ? MAS_OK : MAS_ERR_UNSUPPORTED_SHAPE;
}

int mas_b200_debug_fused_geom(int B, int D, int T_x, int T_y, int max_smem, int num_sms, int32_t *out12) {
    if (B <= 0 || D <= 0 || T_x <= 0 || T_y <= 0 || max_smem <= 0 || num_sms <= 0 || !out12) return MAS_ERR_INVALID_ARGUMENT;
    return debug_fused_geom(B, D, T_x, T_y, max_smem, num_sms, out12) ? MAS_OK : MAS_ERR_UNSUPPORTED_SHAPE;
}

int mas_b200_debug_deal(int P, int BT, int nchunks, int32_t *owner, int32_t *order) { return debug_deal(P, BT, nchunks, owner, order); }

int mas_b200_debug_tile_shape(int T_x, int T_y, int32_t *out6) {
    if (T_x <= 0 || T_y <= 0 || !out6) return MAS_ERR_INVALID_ARGUMENT;
    debug_tile_shape(T_x, T_y, out6);
    return MAS_OK;
}

size_t mas_b200_mle_loss_workspace_bytes(int B, int T_y) {
    if (B <= 0 || T_y <= 0) return 0;
    return mle_loss_workspace_bytes(B, T_y);
}

int mas_b200_mle_loss_f32(const float *z, const float *x_m, const float *x_logs, const int32_t *frame_token, const float *logdet,
                          const int32_t *y_len, float *loss_and_inv_denom, void *workspace, size_t workspace_bytes, int B, int D,
                          int T_x, int T_y, mas_stream_t stream) {
    if (B <= 0 || D <= 0 || T_x <= 0 || T_y <= 0) return MAS_ERR_INVALID_ARGUMENT;   // an empty batch has no mean
    if (!shape_ok(B, T_x, T_y) || D > MAS_B200_MAX_CHANNELS) return MAS_ERR_UNSUPPORTED_SHAPE;
    if (!z || !x_m || !frame_token || !y_len || !loss_and_inv_denom) return MAS_ERR_INVALID_ARGUMENT;
    if (!workspace || workspace_bytes < mle_loss_workspace_bytes(B, T_y)) return MAS_ERR_WORKSPACE_TOO_SMALL;
    return launch_mle_loss(z, x_m, x_logs, frame_token, logdet, y_len, loss_and_inv_denom, workspace, B, D, T_x, T_y,
                           static_cast<cudaStream_t>(stream));
}

int mas_b200_mle_loss_backward_f32(const float *z, const float *x_m, const float *x_logs, const int32_t *frame_token,
                                   const int32_t *durations, const float *scale, float *dz, float *dx_m, float *dx_logs, int B,
                                   int D, int T_x, int T_y, mas_stream_t stream) {
    if (B <= 0 || D <= 0 || T_x <= 0 || T_y <= 0) return MAS_ERR_INVALID_ARGUMENT;
    if (!shape_ok(B, T_x, T_y) || D > MAS_B200_MAX_CHANNELS) return MAS_ERR_UNSUPPORTED_SHAPE;
    if (!z || !x_m || !frame_token || !durations || !scale) return MAS_ERR_INVALID_ARGUMENT;
    if (dx_logs && (!x_logs || !dx_m)) return MAS_ERR_INVALID_ARGUMENT;
    return launch_mle_loss_backward(z, x_m, x_logs, frame_token, durations, scale, dz, dx_m, dx_logs, B, D, T_x, T_y,
                                    static_cast<cudaStream_t>(stream));
}

int mas_b200_duration_loss_f32(const float *logw, const int32_t *durations, const int32_t *x_len, float *loss_and_scale, int B,
                               int T_x, mas_stream_t stream) {
    if (B <= 0 || T_x <= 0) return MAS_ERR_INVALID_ARGUMENT;
    if ((int64_t)B * T_x > (1 << 30)) return MAS_ERR_UNSUPPORTED_SHAPE;
    if (!logw || !durations || !x_len || !loss_and_scale) return MAS_ERR_INVALID_ARGUMENT;
    return launch_duration_loss(logw, durations, x_len, loss_and_scale, B, T_x, static_cast<cudaStream_t>(stream));
}

int mas_b200_duration_loss_backward_f32(const float *logw, const int32_t *durations, const int32_t *x_len, const float *scale,
                                        float *dlogw, int B, int T_x, mas_stream_t stream) {
    if (B <= 0 || T_x <= 0) return MAS_ERR_INVALID_ARGUMENT;
    if ((int64_t)B * T_x > (1 << 30)) return MAS_ERR_UNSUPPORTED_SHAPE;
    if (!logw || !durations || !x_len || !scale || !dlogw) return MAS_ERR_INVALID_ARGUMENT;
    return launch_duration_loss_backward(logw, durations, x_len, scale, dlogw, B, T_x, static_cast<cudaStream_t>(stream));
}

size_t mas_b200_clip_grad_workspace_bytes(int nchunks) { return nchunks > 0 ? align_up((size_t)nchunks * sizeof(double), 256) : 0; }

int mas_b200_clip_grad_value_f32(float *const *chunk_ptr, const int32_t *chunk_count, int nchunks, float clip_value, void *workspace,
                                 size_t workspace_bytes, float *total_norm, mas_stream_t stream) {
    if (nchunks < 0 || !(clip_value >= 0.f)) return MAS_ERR_INVALID_ARGUMENT;
    if (!total_norm) return MAS_ERR_INVALID_ARGUMENT;
    if (nchunks == 0) return cudaMemsetAsync(total_norm, 0, sizeof(float), static_cast<cudaStream_t>(stream)) == cudaSuccess ? MAS_OK : MAS_ERR_CUDA;
    if (!chunk_ptr || !chunk_count) return MAS_ERR_INVALID_ARGUMENT;
    if (!workspace || workspace_bytes < mas_b200_clip_grad_workspace_bytes(nchunks)) return MAS_ERR_WORKSPACE_TOO_SMALL;
    return launch_clip_grad_value(chunk_ptr, chunk_count, nchunks, clip_value, static_cast<double *>(workspace), total_norm,
                                  static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------
// Host-buffer entry: staging buffers are cached per device and grown on demand.
// ---------------------------------------------------------------------------------------------
namespace {
struct HostStage {
    void *d_value = nullptr, *d_path = nullptr, *d_len = nullptr, *d_ws = nullptr;
    size_t value_cap = 0, path_cap = 0, len_cap = 0, ws_cap = 0;
    cudaStream_t stream = nullptr;
};
std::mutex g_stage_mutex;
HostStage g_stage[64];

int grow(void **ptr, size_t *cap, size_t need) {
    if (need <= *cap) return MAS_OK;
    if (*ptr) MAS_CUDA_TRY(cudaFree(*ptr));
    *ptr = nullptr;
    *cap = 0;
    MAS_CUDA_TRY(cudaMalloc(ptr, need));
    *cap = need;
    return MAS_OK;
}

__global__ void path_f32_to_i32_kernel(const float *__restrict__ in, int32_t *__restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = (int32_t)in[i];
}
}  // namespace

int mas_b200_maximum_path_host_i32(int32_t *paths, const float *values, const int32_t *t_xs,
                                   const int32_t *t_ys, int B, int T_x, int T_y,
                                   float max_neg_val, int device) {
    if (!shape_ok(B, T_x, T_y)) return MAS_ERR_UNSUPPORTED_SHAPE;
    if (B == 0 || T_x == 0 || T_y == 0) return MAS_OK;
    if (!paths || !values || !t_xs || !t_ys || device < 0 || device >= 64) return MAS_ERR_INVALID_ARGUMENT;
    for (int b = 0; b < B; ++b)
        if (t_xs[b] < 0 || t_ys[b] < 0 || t_xs[b] > T_x || t_ys[b] > T_y || t_xs[b] > t_ys[b]) return MAS_ERR_BAD_LENGTHS;
    std::lock_guard<std::mutex> lock(g_stage_mutex);
    int prev_dev = 0;
    MAS_CUDA_TRY(cudaGetDevice(&prev_dev));
    MAS_CUDA_TRY(cudaSetDevice(device));
    HostStage &s = g_stage[device];
    int rc = MAS_OK;
    const size_t cells = (size_t)B * T_x * T_y;
    do {
        if (!s.stream && cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) { rc = cuda_fail(cudaGetLastError()); break; }
        // the fp32 path and the int32 result share one buffer pair: value | path(f32) -> path(i32) reuses value
        if ((rc = grow(&s.d_value, &s.value_cap, cells * 4)) != MAS_OK) break;
        if ((rc = grow(&s.d_path, &s.path_cap, cells * 4)) != MAS_OK) break;
        if ((rc = grow(&s.d_len, &s.len_cap, (size_t)B * 8)) != MAS_OK) break;
        if ((rc = grow(&s.d_ws, &s.ws_cap, mas_b200_workspace_bytes(B, T_x, T_y) + 256)) != MAS_OK) break;
        int32_t *d_tx = static_cast<int32_t *>(s.d_len), *d_ty = d_tx + B;
        cudaError_t e;
        if ((e = cudaMemcpyAsync(s.d_value, values, cells * 4, cudaMemcpyHostToDevice, s.stream)) != cudaSuccess ||
            (e = cudaMemcpyAsync(d_tx, t_xs, (size_t)B * 4, cudaMemcpyHostToDevice, s.stream)) != cudaSuccess ||
            (e = cudaMemcpyAsync(d_ty, t_ys, (size_t)B * 4, cudaMemcpyHostToDevice, s.stream)) != cudaSuccess) { rc = cuda_fail(e); break; }
        rc = mas_b200_maximum_path_f32(static_cast<float *>(s.d_value), (int64_t)T_x * T_y, T_y, d_tx, d_ty, nullptr, 0, 0, 0,
                                       static_cast<float *>(s.d_path), nullptr, nullptr, s.d_ws, s.ws_cap, B, T_x, T_y,
                                       max_neg_val, s.stream);
        if (rc != MAS_OK) break;
        // int32 result written over the (no longer needed) staged scores
        int num_sms = 0;
        if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || num_sms <= 0) num_sms = 128;
        path_f32_to_i32_kernel<<<num_sms * 4, 256, 0, s.stream>>>(static_cast<float *>(s.d_path), static_cast<int32_t *>(s.d_value), cells);
        if ((e = cudaGetLastError()) != cudaSuccess) { rc = cuda_fail(e); break; }
        if ((e = cudaMemcpyAsync(paths, s.d_value, cells * 4, cudaMemcpyDeviceToHost, s.stream)) != cudaSuccess) { rc = cuda_fail(e); break; }
        if ((e = cudaStreamSynchronize(s.stream)) != cudaSuccess) { rc = cuda_fail(e); break; }
    } while (false);
    cudaSetDevice(prev_dev);
    return rc;
}

// Frees the staging buffers and the stream the host entry keeps per device.  The library holds no
// other resources; safe to call at any time no host entry is running, and more than once.
void mas_b200_shutdown(void) {
    std::lock_guard<std::mutex> lock(g_stage_mutex);
    int prev_dev = 0;
    if (cudaGetDevice(&prev_dev) != cudaSuccess) return;
    for (int d = 0; d < 64; ++d) {
        HostStage &s = g_stage[d];
        if (!s.d_value && !s.d_path && !s.d_len && !s.d_ws && !s.stream) continue;
        if (cudaSetDevice(d) != cudaSuccess) continue;
        if (s.stream) cudaStreamSynchronize(s.stream);
        cudaFree(s.d_value);
        cudaFree(s.d_path);
        cudaFree(s.d_len);
        cudaFree(s.d_ws);
        if (s.stream) cudaStreamDestroy(s.stream);
        s = HostStage{};
    }
    cudaSetDevice(prev_dev);
}

}  // extern "C"

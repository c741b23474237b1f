// mas_kernels.cuh -- parameter blocks and launchers shared by the C-ABI layer (mas_api.cu).
#pragma once

#include <atomic>

#include "mas_common.cuh"

namespace mas {

// Kernel (1) arguments (see mas_b200_maximum_path_f32 in include/mas_b200.h).
struct PathParams {
    const float *value;
    int64_t value_stride_b, value_stride_x;
    const int32_t *t_x, *t_y;  // both null -> lengths from the mask
    const float *mask;
    int64_t mask_stride_b, mask_stride_x, mask_stride_y;
    float *path;
    int32_t *durations;    // nullable
    int32_t *frame_token;  // nullable
    uint32_t *ws_bits;     // workspace: packed direction bits [B][ceil(T_y/32)][T_x]
    int32_t *ws_tok;       // workspace: frame -> token [B][T_y] when frame_token is null
    const int *exact_flag; // nullable: [B], non-zero = compute this utterance from value * mask literally (mas_mask.cu)
    int B, T_x, T_y;
    float max_neg_val;
    long long *dbg_cycles;  // profiling hook (mas_b200_debug_set_cycle_buffer): [B][16 warps][16] clock64 stamps, or null
};

// Log-likelihood matrix arguments (mas_b200_logp_f32).
struct LogpParams {
    const float *x_m, *x_logs, *z;  // x_logs nullable (mean_only)
    float *logp;
    int B, D, T_x, T_y;
    // single launch only: the utterances' lengths.  Scores outside the reference's band
    // (core.pyx:18) are never read by the sweep, so the producers do not contract them.
    const int32_t *x_len = nullptr, *y_len = nullptr;
};

// Per-device facts the launchers need, queried once per device under a lock (mas_api.cu).
struct DeviceInfo {
    int max_smem_optin;   // cudaDevAttrMaxSharedMemoryPerBlockOptin
    int num_sms;          // cudaDevAttrMultiProcessorCount
};
int get_device_info(int dev, DeviceInfo &out);   // MAS_OK or an error status; dev in [0, 64)

// "the opt-in shared-memory attribute of this kernel has been raised to N bytes on device d": the
// attribute is sticky, so it is only raised; concurrent callers may both raise it, which is harmless.
struct SmemOptIn {
    std::atomic<int> bytes[64];
    template <typename Kernel>
    int ensure(Kernel kern, int dev, int need) {
        if (need > bytes[dev & 63].load(std::memory_order_acquire)) {
            MAS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, need));
            int seen = bytes[dev & 63].load();
            while (seen < need && !bytes[dev & 63].compare_exchange_weak(seen, need)) {}
        }
        return MAS_OK;
    }
};

size_t mask_flag_bytes(int B);
int launch_mask_check(const PathParams &p, int *flags, cudaStream_t stream);

size_t path_simple_workspace_bytes(int B, int T_x, int T_y);
int launch_path_simple(PathParams p, void *workspace, size_t workspace_bytes, cudaStream_t stream);

size_t path_systolic_workspace_bytes(int B, int T_x, int T_y);
// MAS_ERR_UNSUPPORTED_SHAPE = "not for the TMA path": the caller falls back to launch_path_simple.
int launch_path_systolic(PathParams p, void *workspace, size_t workspace_bytes, cudaStream_t stream);
void path_systolic_force_cluster(int k);
bool debug_path_plan(int B, int T_x, int T_y, int max_smem, int num_sms, int32_t *out8);   // host only   // testing hook: CTAs per utterance (0 = heuristic)

// path consumers (mas_expand.cu)
int launch_expand_gather(const float *x, const int32_t *frame_token, float *z, int B, int D, int T_x, int T_y, cudaStream_t stream);
int launch_expand_scatter(const float *dz, const int32_t *durations, float *dx, int B, int D, int T_x, int T_y, cudaStream_t stream);
int launch_logw(const int32_t *durations, const int32_t *x_len, float *logw, int B, int T_x, cudaStream_t stream);
int launch_generate_path(const float *duration, const float *mask, int64_t ms_b, int64_t ms_x, int64_t ms_y, float *path, int B,
                         int T_x, int T_y, cudaStream_t stream);
int debug_deal(int P, int BT, int nchunks, int32_t *owner, int32_t *order);   // host only (mas_logp.cu)
void debug_tile_shape(int T_x, int T_y, int32_t *out6);                        // host only
size_t mle_loss_workspace_bytes(int B, int T_y);
int launch_mle_loss(const float *z, const float *x_m, const float *x_logs, const int32_t *frame_token, const float *logdet,
                    const int32_t *y_len, float *out2, void *workspace, int B, int D, int T_x, int T_y, cudaStream_t stream);
int launch_mle_loss_backward(const float *z, const float *x_m, const float *x_logs, const int32_t *frame_token,
                             const int32_t *durations, const float *scale, float *dz, float *dx_m, float *dx_logs, int B, int D,
                             int T_x, int T_y, cudaStream_t stream);

// the step around the path (mas_train.cu)
int launch_duration_loss(const float *logw, const int32_t *durations, const int32_t *x_len, float *out2, int B, int T_x,
                         cudaStream_t stream);
int launch_duration_loss_backward(const float *logw, const int32_t *durations, const int32_t *x_len, const float *scale, float *dlogw,
                                  int B, int T_x, cudaStream_t stream);
int launch_clip_grad_value(float *const *chunk_ptr, const int32_t *chunk_count, int nchunks, float clip, double *partial,
                           float *total_norm, cudaStream_t stream);

// developer profiling hook: when non-null, kernels stamp clock64() phase times into it
extern std::atomic<long long *> g_dbg_cycles;

int launch_logp(const LogpParams &p, cudaStream_t stream);

// single-launch logp + alignment search (mas_fused.cu); MAS_ERR_UNSUPPORTED_SHAPE -> run the two kernels
size_t fused_workspace_bytes(int B, int D, int T_x, int T_y);
bool debug_fused_geom(int B, int D, int T_x, int T_y, int max_smem, int num_sms, int32_t *out12);   // host only
int launch_fused(const LogpParams &lp, const int32_t *x_len, const int32_t *y_len, float *path, int32_t *durations,
                 int32_t *frame_token, void *workspace, size_t workspace_bytes, float max_neg_val, bool force, cudaStream_t stream);

}  // namespace mas

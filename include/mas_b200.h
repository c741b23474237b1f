/*
 * mas_b200.h -- C ABI of the B200-native (sm_100a) alignment hot path of Glow-TTS training.
 *
 * This is the drop-in boundary.  Every entry point is `extern "C"`, takes plain device/host
 * pointers, sizes and a CUDA stream handle (no torch types), never throws, never synchronises the
 * host (except the *_host convenience entry), and returns an `int` status (MAS_OK == 0).
 * The library is libmas_b200.so, built by nvcc for sm_100a from glow-tts-train_b200/csrc/.
 *
 * Reference interfaces replaced (paths relative to the reference tree, rhasspy/glow-tts-train 0.3.0):
 *   glow_tts_train/monotonic_align/core.pyx:40-45     maximum_path_c(paths, values, t_xs, t_ys, max_neg_val)
 *   glow_tts_train/monotonic_align/core.pyx:9-35      maximum_path_each (forward DP + backtrack)
 *   glow_tts_train/monotonic_align/__init__.py:18-19  t_x / t_y derived from the mask
 *   glow_tts_train/models.py:362-376                  the [B,T_x,T_y] log-likelihood matrix
 *   glow_tts_train/models.py:378-382                  logp -> maximum_path -> attn
 *   glow_tts_train/models.py:393                      token durations (row sums of the path)
 *
 * Vocabulary: an *utterance* b has t_x[b] text tokens (rows, index x) and t_y[b] mel frames
 * (columns, index y).  A *cell* is one (token, frame) pair; cells = B*T_x*T_y is the unit of the
 * throughput metric.  All tensors are row-major with the frame index contiguous.
 */
#ifndef MAS_B200_H_
#define MAS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAS_B200_ABI_VERSION 2

/* ---- status codes ------------------------------------------------------------------------- */
#define MAS_OK 0
#define MAS_ERR_INVALID_ARGUMENT 1   /* null pointer, negative size, bad stride ...               */
#define MAS_ERR_UNSUPPORTED_SHAPE 2  /* T_x / T_y / D beyond what the kernels are built for       */
#define MAS_ERR_WORKSPACE_TOO_SMALL 3
#define MAS_ERR_NO_DEVICE 4          /* no CUDA device, or device is not sm_100 (B200)            */
#define MAS_ERR_CUDA 5               /* a CUDA runtime/driver call failed: mas_b200_last_cuda_error() */
#define MAS_ERR_BAD_LENGTHS 6        /* (host entry only) some t_x > t_y or t_x > T_x ...          */

/* limits of this build */
#define MAS_B200_MAX_TOKENS 2048     /* T_x  */
#define MAS_B200_MAX_FRAMES 65536    /* T_y  */
#define MAS_B200_MAX_CHANNELS 256    /* D    */

/* The reference's "minus infinity" (core.pyx:40 default argument). */
#define MAS_B200_MAX_NEG_VAL (-1e9f)

/* opaque CUDA stream handle (cudaStream_t / CUstream); NULL = legacy default stream */
typedef void *mas_stream_t;

int mas_b200_abi_version(void);
const char *mas_b200_status_string(int status);
/* cudaError_t of the last failing CUDA call made by this library on the calling thread (0 if none) */
int mas_b200_last_cuda_error(void);
/* Developer profiling hook: when `device_buffer` (int64 [B][16][16], device memory) is non-NULL the
 * kernels stamp clock64() phase boundaries per warp into it; NULL (the default) switches it off. */
void mas_b200_debug_set_cycle_buffer(void *device_buffer);
/* Testing hook: force the number of CTAs (thread-block cluster size: 1, 2, 4, 8) that share one
 * utterance in kernel (1); 0 restores the heuristic. */
void mas_b200_debug_force_cluster(int ctas_per_utterance);
/* Testing hook for mas_b200_fused_maximum_path_f32: 0 = the entry chooses between its single launch
 * (a cluster of CTAs per utterance, scores never leave shared memory) and the two kernels back to
 * back by its cost estimate for the shape; 1 = always the two kernels; 2 = the single launch
 * whenever the shape allows it. */
void mas_b200_debug_force_unfused(int mode);
/* Testing hooks, HOST only (no device needed): how the contraction's work is laid out.
 *   debug_tile_shape: out6 = {row_tiles, tile_rows, row groups, column groups, frames per chunk, chunks}
 *   debug_deal: the deal of (row, chunk) units to P persistent CTAs (rows = utterances x row_tiles):
 *     owner[r * nchunks + c] = the CTA that contracts chunk c of row r (-1: nobody), order[...] = its
 *     position in that CTA's sequence; returns how many units were dealt twice (0), < 0 on bad arguments. */
int mas_b200_debug_tile_shape(int T_x, int T_y, int32_t *out6);
/*   debug_path_plan: kernel (1)'s geometry on a device with `max_smem` bytes of opt-in shared memory per
 *     CTA and `num_sms` SMs: out8 = {tokens per lane, sweep warps, TMA ring depth, CTAs per utterance,
 *     tokens per CTA, 32-frame blocks, direction bits in shared memory?, shared memory bytes};
 *     MAS_ERR_UNSUPPORTED_SHAPE when the TMA path cannot take the shape. */
int mas_b200_debug_path_plan(int B, int T_x, int T_y, int max_smem, int num_sms, int32_t *out8);
int mas_b200_debug_deal(int P, int BT, int nchunks, int32_t *owner, int32_t *order);
/*   debug_fused_geom: kernel (2)'s geometry on such a device: out12 = {CTAs per utterance (cluster size),
 *     tokens per sweep lane, tokens per CTA (bound), FFMA teams, warps per team, column groups, frames per
 *     chunk, score-ring depth in 32-frame boxes, direction bits in shared memory?, shared memory bytes,
 *     FFMA warps, rows per ring box}; MAS_ERR_UNSUPPORTED_SHAPE when the single launch cannot take the shape. */
int mas_b200_debug_fused_geom(int B, int D, int T_x, int T_y, int max_smem, int num_sms, int32_t *out12);
/* MAS_OK iff the current CUDA device can run the kernels (compute capability 10.x). */
int mas_b200_device_ok(void);

/*
 * Scratch needed by the entry points below for a batch of shape (B, T_x, T_y): packed backtrack
 * direction bits that do not fit in shared memory, per-utterance lengths, ... .  Caller-owned,
 * device memory, 256-byte aligned, contents undefined before and after a call.  May return 0.
 */
size_t mas_b200_workspace_bytes(int B, int T_x, int T_y);
/* Same for mas_b200_fused_maximum_path_f32 (never smaller than the above). */
size_t mas_b200_fused_workspace_bytes(int B, int D, int T_x, int T_y);

/*
 * Kernel (1): monotonic alignment search on a materialised score matrix.
 * Replaces maximum_path_c (core.pyx:40-45) together with the mask handling of
 * monotonic_align/__init__.py:11,18-19.
 *
 *   value      [B][T_x][T_y] fp32 device, frame index contiguous; utterance/token strides in
 *              ELEMENTS (value_stride_b, value_stride_x).  NOT modified (the reference clobbers
 *              its private copy; there is no copy here).
 *   t_x, t_y   int32 [B] device, per-utterance valid sizes; or both NULL, in which case
 *   mask       [B][T_x][T_y] fp32 device (any strides, in elements) supplies them the reference's
 *              way: t_x[b] = sum_x mask[b,x,0], t_y[b] = sum_y mask[b,0,y] (__init__.py:18-19), and
 *              the scores are value*mask (__init__.py:11).  For the prefix masks models.py:334-337
 *              builds the product equals value on every cell the algorithm touches; that is
 *              VERIFIED on the device (one pass over the mask's valid rectangle, mas_mask.cu), and
 *              utterances whose mask is anything else (interior zeros, fractional entries) are
 *              computed from value*mask literally -- same result as the reference for any mask,
 *              no host synchronisation.
 *   path       [B][T_x][T_y] fp32 device, contiguous: written completely (zeros and ones).
 *   durations  int32 [B][T_x] device or NULL: frames per token (row sums of path; models.py:393).
 *   frame_token int32 [B][T_y] device or NULL: token index of every frame, -1 for y >= t_y[b].
 *   max_neg_val the reference's -1e9.
 *
 * Semantics are bit-exact with the reference for 1 <= t_x <= t_y (tie -> stay on the token,
 * strict '>' / '<' compares, -1e9 boundary, fp32 round-to-nearest adds).  Defined behaviour where
 * the reference has none: t_x == 0 or t_y == 0 -> all-zero path for that utterance; t_x > t_y ->
 * treated as t_x = t_y (first t_y tokens on the diagonal).  Asynchronous on `stream`.
 */
int mas_b200_maximum_path_f32(const float *value, int64_t value_stride_b, int64_t value_stride_x,
                              const int32_t *t_x, const int32_t *t_y,
                              const float *mask, int64_t mask_stride_b, int64_t mask_stride_x,
                              int64_t mask_stride_y,
                              float *path, int32_t *durations, int32_t *frame_token,
                              void *workspace, size_t workspace_bytes,
                              int B, int T_x, int T_y, float max_neg_val, mas_stream_t stream);

/*
 * The log-likelihood matrix of models.py:362-376, materialised (used for tolerance tests and as
 * the unfused pipeline stage):
 *   logp[b,x,y] = sum_d(-0.5*log(2pi) - logs[b,d,x]) + sum_d exp(-2 logs[b,d,x]) * (-0.5 z[b,d,y]^2)
 *               + sum_d (m[b,d,x] exp(-2 logs[b,d,x])) * z[b,d,y] + sum_d -0.5 m[b,d,x]^2 exp(-2 logs[b,d,x])
 * in fp32: (l1 + c) + l4 with c = one accumulator over channels ascending (the l2 and l3 terms
 * interleaved); ((l1+l2)+l3)+l4 when x_logs == NULL.  Within 1e-5 relative of the fp64 formula.
 *   x_m, x_logs [B][D][T_x] fp32 device contiguous; x_logs NULL == zeros (mean_only, config.py:52)
 *   z           [B][D][T_y] fp32 device contiguous
 *   logp        [B][T_x][T_y] fp32 device contiguous
 */
int mas_b200_logp_f32(const float *x_m, const float *x_logs, const float *z, float *logp,
                      int B, int D, int T_x, int T_y, mas_stream_t stream);

/*
 * Kernel (2): log-likelihood + alignment search in one call (models.py:362-382, + :393 through
 * `durations`).  The single launch gives every utterance a thread-block cluster of CTAs that slice its
 * tokens; each CTA contracts its slice's scores over the channels into a ring in SHARED memory and
 * sweeps them there (neighbouring CTAs exchange one boundary token's scores over distributed shared
 * memory), so the [B,T_x,T_y] score matrix is never materialised and the workspace holds only the
 * scratch of the rare literal redo (non-finite scores).  Any batch size runs in that one launch; where
 * the entry's cost estimate says the two kernels back to back (scores of a group of utterances staged
 * in the workspace, at most 256 MB) are faster -- several rounds of utterances per cluster, frame
 * counts that are not a multiple of 4, > 80 channels -- it runs those instead: same results.
 *   x_len, y_len int32 [B] device: valid tokens / frames (what the prefix masks encode).
 * Other arguments as above.  logp tiles are accurate to 1e-5 relative against the fp64 formula;
 * the path equals kernel (1) run on mas_b200_logp_f32's output bit for bit.
 */
int mas_b200_fused_maximum_path_f32(const float *x_m, const float *x_logs, const float *z,
                                    const int32_t *x_len, const int32_t *y_len,
                                    float *path, int32_t *durations, int32_t *frame_token,
                                    void *workspace, size_t workspace_bytes,
                                    int B, int D, int T_x, int T_y, float max_neg_val,
                                    mas_stream_t stream);

/*
 * The path's consumers (models.py:383-393) without the dense path:
 *   expand_prior           z[b,d,y] = x[b,d,frame_token[b,y]] (0 where frame_token < 0): what
 *                          (attn^T @ x^T)^T computes for x = x_m / x_logs (models.py:383-392), exactly
 *   expand_prior_backward  dx[b,d,x] = sum of dz[b,d,y] over token x's run of frames (durations)
 *   log_durations          logw_[b,x] = log(1e-8 + durations[b,x]) for x < x_len[b], else 0 (models.py:393)
 * x, z: [B][D][T_x] / [B][D][T_y] fp32 contiguous; frame_token int32 [B][T_y]; durations int32 [B][T_x].
 */
int mas_b200_expand_prior_f32(const float *x, const int32_t *frame_token, float *z, int B, int D, int T_x, int T_y,
                              mas_stream_t stream);
int mas_b200_expand_prior_backward_f32(const float *dz, const int32_t *durations, float *dx, int B, int D, int T_x, int T_y,
                                       mas_stream_t stream);
int mas_b200_log_durations_f32(const int32_t *durations, const int32_t *x_len, float *logw, int B, int T_x, mas_stream_t stream);

/*
 * SURVEY.md 8f rank 4, the inference-side analogue of the path: `generate_path(duration, mask)`
 * (glow_tts_train/utils.py:99-115, called at models.py:340): cum = cumsum(duration) per utterance,
 *   path[b,x,y] = ((y < cum[b,x]) - (y < cum[b,x-1])) * mask[b,x,y]
 * duration fp32 [B][T_x] contiguous (the model passes ceil(w): integer-valued, for which the fp32
 * running sum is exact), mask fp32 [B,T_x,T_y] with ELEMENT strides (the reference passes the view
 * attn_mask.squeeze(1)), path fp32 [B][T_x][T_y] contiguous, fully written.
 */
int mas_b200_generate_path_f32(const float *duration, const float *mask, int64_t mask_stride_b, int64_t mask_stride_x,
                               int64_t mask_stride_y, float *path, int B, int T_x, int T_y, mas_stream_t stream);

/*
 * the maximum-likelihood loss of the aligned prior, `mle_loss(z, z_m, z_logs,
 * logdet, z_mask)` (glow_tts_train/utils.py:14-23, called at train.py:124), from the TOKEN-level prior
 * and the frame->token map -- z_m / z_logs (models.py:383-392) are never materialised:
 *   loss = (sum_{b,d,y} [logs + 0.5 exp(-2 logs) (z - m)^2] - sum_b logdet[b]) / (D sum_b y_len[b]) + 0.5 log(2 pi),
 *   m = x_m[b,d,frame_token[b,y]], logs = x_logs[b,d,frame_token[b,y]] (0 where frame_token < 0, and
 *   everywhere when x_logs == NULL).  logdet may be NULL.
 * loss_and_inv_denom: 2 floats on the device: the loss and 1 / (D sum y_len) (what the backward scales by).
 * backward: `scale` = DEVICE pointer to (upstream gradient) x (1 / denominator);
 *   dz[b,d,y] = scale exp(-2 logs)(z - m);  dx_m[b,d,x] = -scale sum over x's frames of exp(-2 logs)(z - m);
 *   dx_logs[b,d,x] = scale sum over x's frames of (1 - exp(-2 logs)(z - m)^2);  (d logdet[b] = -scale, left to the caller).
 *   Any of dz / dx_m / dx_logs may be NULL (dx_logs needs x_logs and dx_m).
 * Deterministic: fixed summation order.  z [B][D][T_y], x_m / x_logs [B][D][T_x] fp32 contiguous.
 */
size_t mas_b200_mle_loss_workspace_bytes(int B, int T_y);
int mas_b200_mle_loss_f32(const float *z, const float *x_m, const float *x_logs, const int32_t *frame_token, const float *logdet,
                          const int32_t *y_len, float *loss_and_inv_denom, void *workspace, size_t workspace_bytes, int B, int D,
                          int T_x, int T_y, mas_stream_t stream);
int mas_b200_mle_loss_backward_f32(const float *z, const float *x_m, const float *x_logs, const int32_t *frame_token,
                                   const int32_t *durations, const float *scale, float *dz, float *dx_m, float *dx_logs, int B,
                                   int D, int T_x, int T_y, mas_stream_t stream);

/*
 * SURVEY.md 8f rank 2, second half: `duration_loss(logw, logw_, lengths)` (glow_tts_train/utils.py:26-28,
 * called at train.py:125) from the integer durations: logw_ = log(1e-8 + durations) for x < x_len, else 0
 * (models.py:393) is formed on the fly,
 *   loss = sum_{b,x} (logw[b,x] - logw_[b,x])^2 / sum_b x_len[b].
 * logw fp32 [B][T_x] (the duration predictor's output, already masked as models.py:354 leaves it).
 * loss_and_scale: 2 floats on the device: the loss and 2 / sum x_len (what the backward scales by).
 * backward: `scale` = DEVICE pointer to (upstream gradient) x (2 / sum x_len); dlogw = scale (logw - logw_).
 */
int mas_b200_duration_loss_f32(const float *logw, const int32_t *durations, const int32_t *x_len, float *loss_and_scale, int B,
                               int T_x, mas_stream_t stream);
int mas_b200_duration_loss_backward_f32(const float *logw, const int32_t *durations, const int32_t *x_len, const float *scale,
                                        float *dlogw, int B, int T_x, mas_stream_t stream);

/*
 * `clip_grad_value_(parameters, clip_value)` (glow_tts_train/utils.py:118-132, called
 * at train.py:141/145) without a host synchronisation: the reference reads every gradient's norm back
 * with `.item()` (one sync per parameter tensor and step).  Here the caller passes a DEVICE table of
 * gradient chunks -- chunk_ptr[c] = first float of the chunk, chunk_count[c] = its length (a chunk is a
 * slice of one gradient tensor; any split works, 64 K floats per chunk keep every SM busy) -- and one
 * launch clamps every value to [-clip_value, clip_value] in place (NaN stays NaN, like torch.clamp_) and
 * leaves total_norm[0] = sqrt(sum of squares of the UNclamped gradients) on the device, as the
 * reference's return value.  Deterministic.  workspace: mas_b200_clip_grad_workspace_bytes(nchunks).
 */
size_t mas_b200_clip_grad_workspace_bytes(int nchunks);
int mas_b200_clip_grad_value_f32(float *const *chunk_ptr, const int32_t *chunk_count, int nchunks, float clip_value, void *workspace,
                                 size_t workspace_bytes, float *total_norm, mas_stream_t stream);

/*
 * Host-buffer convenience used for end-to-end timing and by non-torch callers: takes HOST
 * pointers with the layout of maximum_path_c (core.pyx:40) -- values fp32 [B][T_x][T_y]
 * C-contiguous (NOT clobbered), t_xs / t_ys int32 [B], paths int32 [B][T_x][T_y] (fully written)
 * -- stages them through device memory owned by the library on `device`, runs kernel (1) and
 * returns after the result is in `paths`.  Synchronous.
 */
int mas_b200_maximum_path_host_i32(int32_t *paths, const float *values, const int32_t *t_xs,
                                   const int32_t *t_ys, int B, int T_x, int T_y,
                                   float max_neg_val, int device);
/* Releases the device staging buffers and the stream the host entry above keeps per device (the
 * only state the library holds).  Call when no host entry is running; may be called repeatedly. */
void mas_b200_shutdown(void);

#ifdef __cplusplus
}
#endif
#endif /* MAS_B200_H_ */

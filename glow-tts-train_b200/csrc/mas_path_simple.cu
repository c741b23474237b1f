// mas_path_simple.cu -- kernel (1), generic-shape variant: one CTA per utterance, score column in
// shared memory, a CTA barrier per mel frame.  Correct for every shape the ABI admits (any T_y
// alignment, strided value); it is the path taken when the TMA-staged systolic kernel
// (mas_path_systolic.cu) cannot be used, and the on-device cross-check for it.
//
// Algorithm = glow_tts_train/monotonic_align/core.pyx:9-35 restated for a rolling column:
//   forward (core.pyx:17-30):  V[x] <- ((V[x-1] > V[x]) ? V[x-1] : V[x]) + L[x,y], all x in parallel,
//                              one direction bit d[x,y] = (V[x-1] > V[x]) recorded per cell;
//   backtrack (core.pyx:32-35): idx -= (idx != 0 && (idx == y || d[idx,y])).
// Cells with x > y are kept at exactly max_neg_val by adding 0 instead of L (they are outside the
// reference's band, core.pyx:18), which makes `cur` at x == y equal max_neg_val as core.pyx:19-20
// demands without a per-cell special case.
#include "mas_kernels.cuh"

namespace mas {
namespace simple {

constexpr int kThreads = 256;
constexpr int kMaxRowsPerThread = MAS_B200_MAX_TOKENS / kThreads;  // 8

struct SmemLayout {
    int col_off, tile_off, bits_off, dur_off, total;
};

__host__ __device__ inline SmemLayout smem_layout(int T_x, int T_y, int tile_w, bool bits_in_smem) {
    SmemLayout s;
    int off = 0;
    s.col_off = off;
    off += 2 * T_x * 4;
    s.tile_off = off;
    off += T_x * (tile_w + 1) * 4;
    s.dur_off = off;
    off += T_x * 4;
    s.bits_off = off;
    if (bits_in_smem) off += ceil_div(T_y, 32) * T_x * 4;
    s.total = off;
    return s;
}

__global__ void __launch_bounds__(kThreads)
mas_path_simple_kernel(PathParams p, int tile_w, int bits_in_smem) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float s_len[2];
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int T_x = p.T_x, T_y = p.T_y;

    // ---- lengths (monotonic_align/__init__.py:18-19 when they come from the mask) ----
    int tx_raw, ty_raw;
    if (p.t_x != nullptr) {
        tx_raw = p.t_x[b];
        ty_raw = p.t_y[b];
    } else {
        if (tid < 2) s_len[tid] = 0.f;
        __syncthreads();
        float sx = 0.f, sy = 0.f;
        const float *m = p.mask + (int64_t)b * p.mask_stride_b;
        for (int x = tid; x < T_x; x += kThreads) sx += m[(int64_t)x * p.mask_stride_x];
        for (int y = tid; y < T_y; y += kThreads) sy += m[(int64_t)y * p.mask_stride_y];
        atomicAdd(&s_len[0], sx);
        atomicAdd(&s_len[1], sy);
        __syncthreads();
        tx_raw = (int)s_len[0];
        ty_raw = (int)s_len[1];
    }
    const Lengths len = clamp_lengths(tx_raw, ty_raw, T_x, T_y);
    const int tx = len.tx, ty = len.ty;

    const SmemLayout lay = smem_layout(T_x, T_y, tile_w, bits_in_smem != 0);
    float *col = reinterpret_cast<float *>(smem_raw + lay.col_off);    // [2][T_x]
    float *tile = reinterpret_cast<float *>(smem_raw + lay.tile_off);  // [T_x][tile_w+1]
    int *dur = reinterpret_cast<int *>(smem_raw + lay.dur_off);        // [T_x]
    uint32_t *bits = bits_in_smem ? reinterpret_cast<uint32_t *>(smem_raw + lay.bits_off)
                                  : p.ws_bits + (size_t)b * ceil_div(T_y, 32) * T_x;  // [T_y/32][T_x]
    int *tok = p.frame_token ? p.frame_token + (int64_t)b * T_y : p.ws_tok + (int64_t)b * T_y;

    const float neg = p.max_neg_val;
    const float *val = p.value + (int64_t)b * p.value_stride_b;
    // utterances whose mask is not all-ones on the valid rectangle (mas_mask.cu): value * mask, literally
    const float *msk = (p.exact_flag != nullptr && p.exact_flag[b] != 0) ? p.mask + (int64_t)b * p.mask_stride_b : nullptr;
    const int pitch = tile_w + 1;

    for (int x = tid; x < T_x; x += kThreads) {
        col[x] = neg;
        dur[x] = 0;
    }
    uint32_t acc[kMaxRowsPerThread];
#pragma unroll
    for (int k = 0; k < kMaxRowsPerThread; ++k) acc[k] = 0u;

    // ---- forward sweep ----
    int cur_buf = 0;
    for (int y0 = 0; y0 < ty; y0 += tile_w) {
        __syncthreads();  // previous tile fully consumed (and the init above visible)
        const int w = min(tile_w, ty - y0);
        for (int i = tid; i < tx * tile_w; i += kThreads) {
            const int x = i / tile_w, j = i - x * tile_w;
            float v = (j < w) ? val[(int64_t)x * p.value_stride_x + y0 + j] : 0.f;
            if (msk != nullptr && j < w) v *= msk[(int64_t)x * p.mask_stride_x + (int64_t)(y0 + j) * p.mask_stride_y];
            tile[x * pitch + j] = v;
        }
        __syncthreads();
        for (int j = 0; j < w; ++j) {
            const int y = y0 + j;
            const float *vin = col + cur_buf * T_x;
            float *vout = col + (cur_buf ^ 1) * T_x;
#pragma unroll
            for (int k = 0; k < kMaxRowsPerThread; ++k) {
                const int x = tid + k * kThreads;
                if (x < tx) {
                    const float stay = vin[x];
                    const float adv = (x == 0) ? ((y == 0) ? 0.f : neg) : vin[x - 1];
                    const float l = (x > y) ? 0.f : tile[x * pitch + j];
                    const bool take = adv > stay;
                    vout[x] = (take ? adv : stay) + l;
                    acc[k] |= (take ? 1u : 0u) << (y & 31);
                    if ((y & 31) == 31 || y == ty - 1) {
                        bits[(size_t)(y >> 5) * T_x + x] = acc[k];
                        acc[k] = 0u;
                    }
                }
            }
            cur_buf ^= 1;
            __syncthreads();
        }
    }
    if (!bits_in_smem) __threadfence_block();
    __syncthreads();

    // ---- backtrack (core.pyx:32-35), one thread; then the dense path ----
    if (tid == 0) {
        int idx = tx - 1;
        for (int y = ty - 1; y >= 0; --y) {
            tok[y] = idx;
            const uint32_t wbits = bits[(size_t)(y >> 5) * T_x + idx];
            if (idx != 0 && (idx == y || ((wbits >> (y & 31)) & 1u))) --idx;
        }
        __threadfence_block();
    }
    float *out = p.path + (int64_t)b * T_x * T_y;
    const int64_t cells = (int64_t)T_x * T_y;
    if ((cells & 3) == 0 && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
        float4 *o4 = reinterpret_cast<float4 *>(out);
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int64_t i = tid; i < cells / 4; i += kThreads) o4[i] = z4;
    } else {
        for (int64_t i = tid; i < cells; i += kThreads) out[i] = 0.f;
    }
    __syncthreads();
    for (int y = tid; y < T_y; y += kThreads) {
        if (y < ty) {
            const int t = tok[y];
            out[(int64_t)t * T_y + y] = 1.f;
            atomicAdd(&dur[t], 1);
        } else if (p.frame_token) {
            tok[y] = -1;
        }
    }
    __syncthreads();
    if (p.durations) {
        for (int x = tid; x < T_x; x += kThreads) p.durations[(int64_t)b * T_x + x] = dur[x];
    }
}

}  // namespace simple

size_t path_simple_workspace_bytes(int B, int T_x, int T_y) {
    // direction bits (only used when they do not fit in shared memory) + frame->token scratch
    size_t bits = align_up((size_t)B * ceil_div(T_y, 32) * T_x * 4, 256);
    size_t tok = align_up((size_t)B * T_y * 4, 256);
    return bits + tok;
}

int launch_path_simple(PathParams p, void *workspace, size_t workspace_bytes, cudaStream_t stream) {
    using namespace simple;
    if (p.B == 0) return MAS_OK;
    if (workspace_bytes < path_simple_workspace_bytes(p.B, p.T_x, p.T_y) || workspace == nullptr)
        return MAS_ERR_WORKSPACE_TOO_SMALL;
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    p.ws_bits = reinterpret_cast<uint32_t *>(ws);
    p.ws_tok = reinterpret_cast<int *>(ws + align_up((size_t)p.B * ceil_div(p.T_y, 32) * p.T_x * 4, 256));

    int dev = 0, max_smem = 0;
    MAS_CUDA_TRY(cudaGetDevice(&dev));
    MAS_CUDA_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    max_smem -= 1024;  // static shared + reserve

    int tile_w = 32;
    while (tile_w > 4 && smem_layout(p.T_x, p.T_y, tile_w, false).total > max_smem) tile_w >>= 1;
    if (smem_layout(p.T_x, p.T_y, tile_w, false).total > max_smem) return MAS_ERR_UNSUPPORTED_SHAPE;
    const int bits_in_smem = smem_layout(p.T_x, p.T_y, tile_w, true).total <= max_smem ? 1 : 0;
    const int smem = smem_layout(p.T_x, p.T_y, tile_w, bits_in_smem != 0).total;
    MAS_CUDA_TRY(cudaFuncSetAttribute(mas_path_simple_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mas_path_simple_kernel<<<p.B, kThreads, smem, stream>>>(p, tile_w, bits_in_smem);
    MAS_CUDA_TRY(cudaGetLastError());
    return MAS_OK;
}

}  // namespace mas

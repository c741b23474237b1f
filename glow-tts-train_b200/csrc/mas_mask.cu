// mas_mask.cu -- is `value * mask` (monotonic_align/__init__.py:11) an identity for this utterance?
//
// The reference multiplies the scores by the mask before the search and derives the valid sizes
// from the mask's first column / first row (__init__.py:18-19).  For the prefix masks that
// models.py:334-337 builds the product changes no cell the algorithm reads, so the kernels read
// `value` directly.  That shortcut is only taken when it is PROVEN on the device: this kernel
// scans the mask over the valid rectangle [0,t_x) x [0,t_y) and raises a per-utterance flag when
// any entry differs from 1.0f.  Flagged utterances are recomputed from value * mask literally by
// the exact compare/select sweep (mas_dp_cta.cuh: exact_sweep_cta0; mas_path_simple.cu), so the
// result equals the reference's for ANY mask, with no host synchronisation.
#include "mas_kernels.cuh"

namespace mas {
namespace {

constexpr int kThreads = 256;
constexpr int kSlabRows = 16;      // tokens per CTA

__global__ void __launch_bounds__(kThreads) mas_mask_check_kernel(PathParams p, int *flags) {
    __shared__ float s_len[2];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int T_x = p.T_x, T_y = p.T_y;
    const float *m = p.mask + (int64_t)b * p.mask_stride_b;
    if (tid < 2) s_len[tid] = 0.f;
    __syncthreads();
    float sx = 0.f, sy = 0.f;
    for (int x = tid; x < T_x; x += kThreads) sx += m[(int64_t)x * p.mask_stride_x];
    for (int y = tid; y < T_y; y += kThreads) sy += m[(int64_t)y * p.mask_stride_y];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sy += __shfl_xor_sync(0xffffffffu, sy, o);
    }
    if ((tid & 31) == 0) {
        atomicAdd(&s_len[0], sx);
        atomicAdd(&s_len[1], sy);
    }
    __syncthreads();
    // the sums are integers for a 0/1 mask, so their order does not matter; for any other mask the
    // rectangle test below fails and the exact path recomputes the lengths the same way
    const Lengths len = clamp_lengths((int)s_len[0], (int)s_len[1], T_x, T_y);
    // a sum that is not the rectangle's extent (fractional entries, entries > 1) must flag as well
    int bad = (s_len[0] != (float)(int)s_len[0]) || (s_len[1] != (float)(int)s_len[1]) ||
              (int)s_len[0] > T_x || (int)s_len[1] > T_y;
    const int x0 = blockIdx.y * kSlabRows, x1 = min(x0 + kSlabRows, len.tx);
    const bool vec = p.mask_stride_y == 1 && (p.mask_stride_x & 3) == 0 && (p.mask_stride_b & 3) == 0 &&
                     (reinterpret_cast<uintptr_t>(p.mask) & 15) == 0;
    for (int x = x0; x < x1; ++x) {
        const float *row = m + (int64_t)x * p.mask_stride_x;
        if (vec) {
            const int n4 = len.ty >> 2;
            for (int i = tid; i < n4; i += kThreads) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(row) + i);
                bad |= (v.x != 1.f) | (v.y != 1.f) | (v.z != 1.f) | (v.w != 1.f);
            }
            for (int y = (n4 << 2) + tid; y < len.ty; y += kThreads) bad |= row[y] != 1.f;
        } else {
            for (int y = tid; y < len.ty; y += kThreads) bad |= row[(int64_t)y * p.mask_stride_y] != 1.f;
        }
    }
    if (__syncthreads_or(bad) && tid == 0) atomicOr(flags + b, 1);
}

}  // namespace

size_t mask_flag_bytes(int B) { return align_up((size_t)B * 4, 256); }

// flags: int [B] device, zeroed here; 1 = the utterance's mask is not all-ones on its valid rectangle
int launch_mask_check(const PathParams &p, int *flags, cudaStream_t stream) {
    MAS_CUDA_TRY(cudaMemsetAsync(flags, 0, (size_t)p.B * 4, stream));
    dim3 grid((unsigned)p.B, (unsigned)ceil_div(p.T_x, kSlabRows));
    mas_mask_check_kernel<<<grid, kThreads, 0, stream>>>(p, flags);
    MAS_CUDA_TRY(cudaGetLastError());
    return MAS_OK;
}

}  // namespace mas

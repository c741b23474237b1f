// mas_logp_tile.cuh -- the per-cell arithmetic of the log-likelihood matrix (models.py:362-376),
// shared by the materialising kernel (mas_logp.cu) and the fused kernel so that both give
// bit-identical scores: same operands, same FFMA order (ascending channel), same final adds.
#pragma once

#include "mas_common.cuh"

namespace mas {

// -0.5 * log(2*pi) rounded to fp32 (models.py:364: -0.5 * math.log(2 * math.pi), a Python double
// that torch turns into the fp32 scalar of the tensor expression).
constexpr float kNegHalfLog2Pi = -0.91893853320467274178f;

// Token-side operands of channel d for token x: inv_var = exp(-2 logs) (models.py:363) and
// m * inv_var (models.py:371).  Out-of-range tokens give zeros.
__device__ __forceinline__ void token_operands(const float *__restrict__ xm, const float *__restrict__ xl,
                                               int T_x, int d, int x, float &inv_var, float &mean_inv_var) {
    if (x < T_x) {
        const float m = __ldg(xm + (int64_t)d * T_x + x);
        const float r = xl ? expf(-2.0f * __ldg(xl + (int64_t)d * T_x + x)) : 1.0f;
        inv_var = r;
        mean_inv_var = m * r;
    } else {
        inv_var = 0.f;
        mean_inv_var = 0.f;
    }
}

// Frame-side operands: z and -0.5 z^2 (models.py:368).
__device__ __forceinline__ void frame_operands(const float *__restrict__ z, int T_y, int d, int y,
                                               float &zv, float &neg_half_zsq) {
    zv = (y < T_y) ? __ldg(z + (int64_t)d * T_y + y) : 0.f;
    neg_half_zsq = -0.5f * (zv * zv);
}

// l1 and l4 of token x (models.py:364-366, 373-375), channels summed in ascending order.
__device__ __forceinline__ void row_constants(const float *__restrict__ xm, const float *__restrict__ xl,
                                              int D, int T_x, int x, float &l1, float &l4) {
    l1 = 0.f;
    l4 = 0.f;
    if (x >= T_x) return;
    for (int d = 0; d < D; ++d) {
        const float m = __ldg(xm + (int64_t)d * T_x + x);
        const float ls = xl ? __ldg(xl + (int64_t)d * T_x + x) : 0.f;
        const float r = xl ? expf(-2.0f * ls) : 1.0f;
        l1 += kNegHalfLog2Pi - ls;
        l4 = fmaf(-0.5f * (m * m), r, l4);
    }
}

// One channel of one cell: l2 += inv_var * (-0.5 z^2), l3 += (m inv_var) * z.
__device__ __forceinline__ void logp_cell_fma(float &l2, float &l3, float inv_var, float mean_inv_var,
                                              float neg_half_zsq, float zv) {
    l2 = fmaf(inv_var, neg_half_zsq, l2);
    l3 = fmaf(mean_inv_var, zv, l3);
}

// Final adds in the reference's order (models.py:376).
__device__ __forceinline__ float logp_cell_finish(float l1, float l2, float l3, float l4) {
    return ((l1 + l2) + l3) + l4;
}

}  // namespace mas

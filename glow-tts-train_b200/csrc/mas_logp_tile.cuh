// mas_logp_tile.cuh -- the arithmetic of the log-likelihood matrix (models.py:362-376), shared by
// the materialising kernel (mas_logp.cu) and the fused kernel so that both give bit-identical
// scores: same operands, same FFMA order, same final adds.
//
//   logp[x,y] = (l1[x] + c[x,y]) + l4[x]
//   c[x,y]    = sum over channels d, ascending, ONE accumulator:
//                   c = fma(-0.5 inv_var[d,x], z[d,y]^2, c)        (models.py:367-369, "logp2"; the
//                       -0.5 rides on the token-side operand: a power of two, so the product is
//                       the one the reference forms, and the inner loop has one FMUL per z less)
//                   c = fma(m[d,x] inv_var[d,x], z[d,y], c)        (models.py:370-372, "logp3")
//   l1[x]     = sum_d (-0.5 log(2 pi) - logs[d,x])                 (models.py:364-366)
//   l4[x]     = sum_d -0.5 m[d,x]^2 inv_var[d,x]                   (models.py:373-375)
//
// The reference adds ((l1 + l2) + l3) + l4 with l2, l3 from two cuBLAS/MKL matmuls whose K order is
// unspecified; interleaving the two contractions into one accumulator halves the register tile and
// stays inside the 1e-5 relative tolerance the north star states (measured ~1e-6).
#pragma once

#include "mas_common.cuh"

namespace mas {

// -0.5 * log(2*pi) rounded to fp32 (models.py:364: -0.5 * math.log(2 * math.pi), a Python double
// that torch turns into the fp32 scalar of the tensor expression).
constexpr float kNegHalfLog2Pi = -0.91893853320467274178f;

// Final adds (models.py:376).
__device__ __forceinline__ float logp_cell_finish(float l1, float c, float l4) { return (l1 + c) + l4; }
// mean_only: l2 is the per-frame sum of -0.5 z^2 and the reference's own order applies, ((l1+l2)+l3)+l4
__device__ __forceinline__ float logp_cell_finish_mean_only(float l1, float l2, float l3, float l4) {
    return ((l1 + l2) + l3) + l4;
}

// ---------------------------------------------------------------------------------------------
// Register-tiled contraction: every thread owns a TM (tokens) x 8 (frames) block of cells.
// Shared-memory operands:
//   sInv, sMiv : [D][tile_rows]   token-side (-0.5 inv_var, m inv_var), token index contiguous
//   sZ         : [D][64]          frame-side, a 64-frame chunk of z; z^2 is formed in registers
// Thread (rg, cg): tokens TM rg .. TM rg + TM-1; frames {4 cg .. 4 cg + 3} and {32 + 4 cg .. 32 + 4 cg + 3},
// so that the 8 column groups of a warp read one contiguous 128-byte line per LDS.128.
// Measured on B200 (profiles/probes/probe_ffma2.cu, FFMA per cycle per SM of 128):
//   8x8 tiles, 208 tokens, 224 threads, 1 CTA/SM : 68      4x8 tiles, 104 tokens, 224 threads, 2 CTA/SM : 98
// The smaller tile loads more operand floats per FFMA (0.31 vs 0.25) but leaves room for 14 warps per
// SM, which hides the shared-memory latency and the register-bank conflicts nvcc leaves behind.
// ---------------------------------------------------------------------------------------------
constexpr int kGemmFrames = 64;   // frames per chunk
constexpr int kGemmTM = 4;        // tokens per thread

// kD: compile-time channel count (80 mel channels, the case that matters) or 0 = run-time D
// kMeanOnly: logs == 0 (config.py:52 `mean_only`, the reference default): inv_var == 1, so the
// inv_var term is a per-FRAME sum handled by the caller and only the mean term is contracted here.
template <int TM, bool kInit, int kD, bool kMeanOnly>
__device__ __forceinline__ void gemm_tile_d(const float *__restrict__ sInv, const float *__restrict__ sMiv,
                                            const float *__restrict__ sZ, int D_rt, int tile_rows, int rg, int cg,
                                            float (&acc)[TM][8]) {
    const int D = kD ? kD : D_rt;
    static_assert(TM % 4 == 0, "token tile is loaded with 16-byte reads");
    if (kInit) {
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    }
    const float *pa = sInv + rg * TM, *pb = sMiv + rg * TM;
    const float *pz = sZ + cg * 4;
#pragma unroll 2
    for (int d = 0; d < D; ++d) {
        float av[TM], bv[TM];
#pragma unroll
        for (int q = 0; q < TM / 4; ++q) {
            const float4 b = *reinterpret_cast<const float4 *>(pb + 4 * q);
            bv[4 * q] = b.x, bv[4 * q + 1] = b.y, bv[4 * q + 2] = b.z, bv[4 * q + 3] = b.w;
            if (!kMeanOnly) {
                const float4 a = *reinterpret_cast<const float4 *>(pa + 4 * q);
                av[4 * q] = a.x, av[4 * q + 1] = a.y, av[4 * q + 2] = a.z, av[4 * q + 3] = a.w;
            }
        }
        const float4 z0 = *reinterpret_cast<const float4 *>(pz), z1 = *reinterpret_cast<const float4 *>(pz + 32);
        pa += tile_rows;
        pb += tile_rows;
        pz += kGemmFrames;
        const float zv[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
        // Same per-cell order as logp_cell_fma (first the inv_var term, then the mean term), issued
        // as two sweeps over the register tile so that consecutive FFMAs share an operand.
        if (!kMeanOnly) {
            float qv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) qv[j] = zv[j] * zv[j];                // models.py:368, -0.5 is in av
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], qv[j], acc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(bv[i], zv[j], acc[i][j]);
    }
}

template <int TM, bool kInit, bool kMeanOnly = false>
__device__ __forceinline__ void gemm_tile(const float *__restrict__ sInv, const float *__restrict__ sMiv,
                                          const float *__restrict__ sZ, int D, int tile_rows, int rg, int cg,
                                          float (&acc)[TM][8]) {
    if (D == 80)
        gemm_tile_d<TM, kInit, 80, kMeanOnly>(sInv, sMiv, sZ, D, tile_rows, rg, cg, acc);
    else
        gemm_tile_d<TM, kInit, 0, kMeanOnly>(sInv, sMiv, sZ, D, tile_rows, rg, cg, acc);
}

}  // namespace mas

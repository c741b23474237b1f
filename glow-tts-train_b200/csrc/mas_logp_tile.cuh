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
// Register-tiled contraction: every thread owns a 4 (tokens) x 8 (frames) block of cells.
// Shared-memory operands:
//   sInv, sMiv : [D][tile_rows]   token-side (-0.5 inv_var, m inv_var), token index contiguous
//   sZ         : [D][F]           frame-side, an F-frame chunk of z (F = 8 CG); z^2 is formed in registers
// Thread (rg, cg), rg < RG = tile_rows / 4, cg < CG: tokens 4 rg .. 4 rg + 3; frames {4 cg .. 4 cg + 3}
// and {4 CG + 4 cg .. 4 CG + 4 cg + 3}: neighbouring lanes read neighbouring 16-byte pieces.
//
// Geometry (profiles/probes/probe_ffma3.cu, whole-kernel timing, useful FFMA per cycle per SM of 128;
// the loop also issues one FMUL per eight FFMA, so 114 would be a saturated FMA pipe):
//   7 warps x 2 CTAs per SM (104 x 64 cells per CTA)   72      8 warps x 2 CTAs (128 x 64 / 100 x 80)   87-89
// A warp lives on scheduler (warp id mod 4): with 7 warps per CTA one scheduler in four carries half
// the load of the others and the busiest ones set the time.  Hence CTAs of 16 warps, one per SM, four
// warps per scheduler, and a tile of RG x CG <= 512 threads shaped to the utterance (200 tokens ->
// 50 x 10: 500 threads, 80-frame chunks).  Operand prefetch by hand was slower than nvcc's schedule.
// ---------------------------------------------------------------------------------------------
constexpr int kGemmThreads = 512;     // 16 warps, one CTA per SM
constexpr int kGemmTM = 4;            // tokens per thread
constexpr int kGemmMinCG = 8, kGemmMaxCG = 16;   // column groups per chunk -> 64 .. 128 frames
constexpr int kGemmMaxTileRows = 256; // 64 token groups x 8 column groups = 512 threads

// Packed FP32 pairs (Blackwell `fma.rn.f32x2` / `mul.rn.f32x2`, SASS FFMA2 / FMUL2): one issue slot
// for two IEEE fp32 FMAs, each lane rounded exactly like a scalar FFMA.  The FMA pipe does the same
// work either way; what it frees are issue slots -- the scalar loop was issue-bound (164 slots per two
// channels of which 144 FMA-pipe), the packed one is pipe-bound (86 slots).
// profiles/probes/probe_ffma4.cu: 14.85 -> 14.16 us per 200 x 80 unit.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f32x2_pack(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f32x2_unpack(f32x2 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 f32x2_fma(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
// acc += a * b in place: the same register pair in and out, so nvcc has nothing to rotate at the
// loop's back edge (with separate operands it renamed the accumulators across the unrolled channels
// and paid 16 moves per trip, most of them IMAD.MOV on the FMA pipe)
__device__ __forceinline__ void f32x2_fma_acc(f32x2 &acc, f32x2 a, f32x2 b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ f32x2 f32x2_mul(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// The register tile: v[i][j] = cells (token 4 rg + i, frames 2j and 2j+1 of the thread's eight).
struct GemmAcc {
    f32x2 v[kGemmTM][4];
    // frames 4h .. 4h+3 of token i as four floats
    __device__ __forceinline__ void quad(int i, int h, float (&c)[4]) const {
        f32x2_unpack(v[i][2 * h], c[0], c[1]);
        f32x2_unpack(v[i][2 * h + 1], c[2], c[3]);
    }
};

// kD: compile-time channel count (80 mel channels, the case that matters) or 0 = run-time D
// kMeanOnly: logs == 0 (config.py:52 `mean_only`, the reference default): inv_var == 1, so the
// inv_var term is a per-FRAME sum handled by the caller and only the mean term is contracted here.
// F: frames per chunk (row stride of sZ); the thread's second frame group starts at F / 2.
template <bool kInit, int kD, bool kMeanOnly>
__device__ __forceinline__ void gemm_tile_d(const float *__restrict__ sInv, const float *__restrict__ sMiv,
                                            const float *__restrict__ sZ, int D_rt, int tile_rows, int F, int rg, int cg,
                                            GemmAcc &acc) {
    const int D = kD ? kD : D_rt;
    constexpr int TM = kGemmTM;
    if (kInit) {
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc.v[i][j] = 0ull;
    }
    const float *pa = sInv + rg * TM, *pb = sMiv + rg * TM;
    const float *pz = sZ + cg * 4;
    const int half = F >> 1;
#pragma unroll 16
    for (int d = 0; d < D; ++d) {
        f32x2 av[TM], bv[TM];
        const float4 b = *reinterpret_cast<const float4 *>(pb);
        bv[0] = f32x2_pack(b.x, b.x), bv[1] = f32x2_pack(b.y, b.y), bv[2] = f32x2_pack(b.z, b.z), bv[3] = f32x2_pack(b.w, b.w);
        if (!kMeanOnly) {
            const float4 a = *reinterpret_cast<const float4 *>(pa);
            av[0] = f32x2_pack(a.x, a.x), av[1] = f32x2_pack(a.y, a.y), av[2] = f32x2_pack(a.z, a.z), av[3] = f32x2_pack(a.w, a.w);
        }
        const ulonglong2 z0 = *reinterpret_cast<const ulonglong2 *>(pz), z1 = *reinterpret_cast<const ulonglong2 *>(pz + half);
        pa += tile_rows;
        pb += tile_rows;
        pz += F;
        const f32x2 zv[4] = {z0.x, z0.y, z1.x, z1.y};
        // Same per-cell order everywhere (first the inv_var term, then the mean term), issued as two
        // sweeps over the register tile so that consecutive FMAs share an operand.
        if (!kMeanOnly) {
            f32x2 qv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) qv[j] = f32x2_mul(zv[j], zv[j]);      // models.py:368, -0.5 is in av
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) f32x2_fma_acc(acc.v[i][j], av[i], qv[j]);
        }
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) f32x2_fma_acc(acc.v[i][j], bv[i], zv[j]);
    }
}

template <bool kInit, bool kMeanOnly = false>
__device__ __forceinline__ void gemm_tile(const float *__restrict__ sInv, const float *__restrict__ sMiv,
                                          const float *__restrict__ sZ, int D, int tile_rows, int F, int rg, int cg,
                                          GemmAcc &acc) {
    if (D == 80)
        gemm_tile_d<kInit, 80, kMeanOnly>(sInv, sMiv, sZ, D, tile_rows, F, rg, cg, acc);
    else
        gemm_tile_d<kInit, 0, kMeanOnly>(sInv, sMiv, sZ, D, tile_rows, F, rg, cg, acc);
}

// Shape of one CTA's tile for T_x tokens / T_y frames: row tiles of at most 256 tokens, as many
// column groups as fit in 512 threads.
struct TileShape {
    int row_tiles, tile_rows, RG, CG, F, nchunks;
};
__host__ __device__ inline TileShape make_tile_shape(int T_x, int T_y) {
    TileShape t;
    t.row_tiles = ceil_div(T_x, kGemmMaxTileRows);
    t.tile_rows = ceil_div(ceil_div(T_x, t.row_tiles), kGemmTM) * kGemmTM;
    t.RG = t.tile_rows / kGemmTM;
    int cg = kGemmThreads / t.RG;
    cg = cg > kGemmMaxCG ? kGemmMaxCG : cg;
    const int need = ceil_div(T_y, 8);                  // no wider than the utterance
    cg = cg > need ? need : cg;
    cg = cg < kGemmMinCG ? kGemmMinCG : cg;
    // even out the chunks: the narrowest CG that keeps the chunk count
    const int n = ceil_div(T_y, 8 * cg);
    while (cg > kGemmMinCG && ceil_div(T_y, 8 * (cg - 1)) == n) --cg;
    t.CG = cg;
    t.F = 8 * cg;
    t.nchunks = ceil_div(T_y, t.F);
    return t;
}

}  // namespace mas

// mas_logp_cta.cuh -- the per-CTA program of the log-likelihood contraction (models.py:362-376):
// one tile of tokens of one utterance, a list of F-frame chunks.  Shared by the materialising
// kernel (mas_logp.cu) and the fused launch (mas_fused.cu), where it also raises a ready flag per
// chunk for the sweep CTAs.  Needs T_y % 4 == 0, 16-byte aligned z / logp rows, D <= 80.
//
// Also here: how the (utterance, token tile, chunk) units are dealt to P persistent CTAs
// (struct Deal), the same for both kernels.
#pragma once

#include "mas_kernels.cuh"
#include "mas_logp_tile.cuh"
#include "mas_ptx.cuh"

namespace mas {
namespace logp {

constexpr int kPanel = 80;          // channels resident in shared memory at a time

// floats of shared memory the program needs
__host__ __device__ inline int cta_smem_floats(int D, const TileShape &t) {
    return 2 * D * t.tile_rows + 2 * D * t.F + 2 * t.tile_rows + t.F + 8 * t.tile_rows;
}

// ---------------------------------------------------------------------------------------------
// The three steps of one (utterance, token tile, chunk) unit, shared by every kernel that contracts
// scores: the materialising kernel walks them over a static list of chunks (logp_cta below), the
// single launch (mas_fused.cu) over units it claims at run time.  Same operands, same order of
// operations everywhere -> bit-identical scores.
// ---------------------------------------------------------------------------------------------
struct CtaSmem {
    float *sInv, *sMiv, *sZ, *sL1, *sL4, *sL2, *sPart;
};
__device__ __forceinline__ CtaSmem carve_smem(float *sm, int D, const TileShape &t) {
    CtaSmem s;
    s.sInv = sm;                                    // [D][tile_rows]
    s.sMiv = s.sInv + D * t.tile_rows;              // [D][tile_rows]
    s.sZ = s.sMiv + D * t.tile_rows;                // [2][D][F]  (double-buffered chunk of z)
    s.sL1 = s.sZ + 2 * D * t.F;                     // [tile_rows]
    s.sL4 = s.sL1 + t.tile_rows;                    // [tile_rows]
    s.sL2 = s.sL4 + t.tile_rows;                    // [F] mean_only: per-frame sum of -0.5 z^2
    s.sPart = s.sL2 + t.F;                          // [nsh <= 4][2][tile_rows] partial row constants
    return s;
}

// frames [ch F, ch F + F) of z, all channels, into buffer `buf` (cp.async, 16-byte pieces; frames
// beyond T_y are zero-filled).  Commits one cp.async group.
__device__ __forceinline__ void stage_frames_async(const LogpParams &p, const CtaSmem &s, const TileShape &t, int b, int ch,
                                                   int buf) {
    const int D = p.D, T_y = p.T_y, F = t.F, f4 = F >> 2;
    const float *zg = p.z + (int64_t)b * D * T_y;
    const int y0 = ch * F;
    float *dst = s.sZ + buf * D * F;
    for (int i = threadIdx.x; i < D * f4; i += blockDim.x) {
        const int d = i / f4, y = y0 + ((i - d * f4) << 2);
        ptx::cp_async_16(dst + (i << 2), zg + (int64_t)d * T_y + (y < T_y ? y : 0), y < T_y);
    }
    ptx::cp_async_commit();
}

// token side of tile [x0, x0 + tile_rows) of utterance b: thread (x, h) stages token x0+x for one
// contiguous share of the channels (coalesced over x) and sums its share of the row constants,
// channels ascending; the shares are then added in order.  nsh = how many threads serve a token
// (2 at 512 / 200).  The loop is kept SMALL: this is cold code that every warp runs once, and
// unrolled by 20 it was bound by instruction fetch (stall_no_inst, profiles/r1_ncu_logp.txt), not
// by the loads.  Ends with a CTA barrier; the caller guarantees nobody still reads the old tile.
__device__ __forceinline__ void stage_tokens(const LogpParams &p, const CtaSmem &s, const TileShape &t, int b, int x0) {
    const int D = p.D, T_x = p.T_x, tile_rows = t.tile_rows;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const float *xm = p.x_m + (int64_t)b * D * T_x;
    const float *xl = p.x_logs ? p.x_logs + (int64_t)b * D * T_x : nullptr;
    const int nsh = max(1, min(4, nthr / tile_rows));
    const int dsh = ceil_div(D, nsh);
    if (tid < nsh * tile_rows) {
        const int h = tid / tile_rows, x = tid - h * tile_rows, xg = x0 + x;
        const int d0 = h * dsh, d1 = min(D, d0 + dsh);
        float l1 = 0.f, l4 = 0.f;
        if (xg < T_x) {
#pragma unroll 4
            for (int d = d0; d < d1; ++d) {
                const float m = __ldg(xm + (int64_t)d * T_x + xg);
                const float ls = xl ? __ldg(xl + (int64_t)d * T_x + xg) : 0.f;
                const float r = xl ? expf(-2.0f * ls) : 1.0f;         // models.py:363
                s.sInv[d * tile_rows + x] = -0.5f * r;                  // models.py:368
                s.sMiv[d * tile_rows + x] = m * r;                      // models.py:371
                l1 += kNegHalfLog2Pi - ls;                              // models.py:364-366
                l4 = fmaf(-0.5f * (m * m), r, l4);                      // models.py:373-375
            }
        } else {
            for (int d = d0; d < d1; ++d) s.sInv[d * tile_rows + x] = s.sMiv[d * tile_rows + x] = 0.f;
        }
        s.sPart[(2 * h) * tile_rows + x] = l1;
        s.sPart[(2 * h + 1) * tile_rows + x] = l4;
    }
    __syncthreads();
    if (tid < tile_rows) {
        float l1 = s.sPart[tid], l4 = s.sPart[tile_rows + tid];
        for (int h = 1; h < nsh; ++h) {
            l1 += s.sPart[(2 * h) * tile_rows + tid];
            l4 += s.sPart[(2 * h + 1) * tile_rows + tid];
        }
        s.sL1[tid] = l1;
        s.sL4[tid] = l4;
    }
    __syncthreads();
}

// mean_only: inv_var == 1, the inv_var term does not depend on the token (models.py:367-369 with
// x_logs == 0): one sum per frame, channels ascending.  Ends with a CTA barrier.
__device__ __forceinline__ void frame_sums_mean_only(const CtaSmem &s, int D, int F, int buf) {
    if ((int)threadIdx.x < F) {
        const float *zc = s.sZ + buf * D * F + threadIdx.x;
        float l2 = 0.f;
        for (int d = 0; d < D; ++d) {
            const float zv = zc[d * F];
            l2 = fmaf(-0.5f * zv, zv, l2);
        }
        s.sL2[threadIdx.x] = l2;
    }
    __syncthreads();
}

// Where a chunk's scores go: row x of the tile's utterance starts at out + x * row_stride; frame
// y0 + j of the chunk lands in column col0 + j.  `keep_frames`: frames >= T_y of the last chunk are
// skipped (the materialised matrix has no such columns) or stored (a ring row has room for the
// whole chunk, and the sweep's boxes read it: they must be finite).
struct ChunkOut {
    float *out;
    int64_t row_stride;
    int col0;
    bool all_frames;
};

// Contract chunk `ch` (already staged in buffer `buf`) of tile x0 and store the scores.
// mode 2: contract, 1: store zeros (cells the sweep covers but never uses), 0: nothing -- per THREAD
// (its four tokens x the chunk), decided by the caller's band test.
__device__ __forceinline__ void contract_chunk(const LogpParams &p, const CtaSmem &s, const TileShape &t, int x0, int ch, int buf,
                                               int mode, const ChunkOut &o) {
    const int D = p.D, T_x = p.T_x, T_y = p.T_y, tile_rows = t.tile_rows, F = t.F, CG = t.CG;
    const bool mean_only = p.x_logs == nullptr;
    const int tid = threadIdx.x;
    const int rg = tid / CG, cg = tid - rg * CG;        // 4 tokens x {4+4} frames per thread
    if (rg >= t.RG || mode == 0) return;
    GemmAcc acc;
    if (mode == 1) {
#pragma unroll
        for (int i = 0; i < kGemmTM; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc.v[i][j] = 0ull;
    } else if (mean_only) {
        gemm_tile<true, true>(s.sInv, s.sMiv, s.sZ + buf * D * F, D, tile_rows, F, rg, cg, acc);
    } else {
        gemm_tile<true, false>(s.sInv, s.sMiv, s.sZ + buf * D * F, D, tile_rows, F, rg, cg, acc);
    }
    const int y0 = ch * F;
#pragma unroll
    for (int i = 0; i < kGemmTM; ++i) {
        const int xr = rg * kGemmTM + i, x = x0 + xr;
        if (x >= T_x) break;
        const float l1 = s.sL1[xr], l4 = s.sL4[xr];
        float *row = o.out + (int64_t)x * o.row_stride + o.col0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int yl = (F >> 1) * h + 4 * cg;
            if (o.all_frames || y0 + yl < T_y) {            // T_y % 4 == 0: whole float4 or nothing
                float c[4];
                acc.quad(i, h, c);
                float4 r;
                if (mode == 1) {
                    r = make_float4(0.f, 0.f, 0.f, 0.f);
                } else if (mean_only) {
                    const float4 l2 = *reinterpret_cast<const float4 *>(s.sL2 + yl);
                    r.x = logp_cell_finish_mean_only(l1, l2.x, c[0], l4);
                    r.y = logp_cell_finish_mean_only(l1, l2.y, c[1], l4);
                    r.z = logp_cell_finish_mean_only(l1, l2.z, c[2], l4);
                    r.w = logp_cell_finish_mean_only(l1, l2.w, c[3], l4);
                } else {
                    r.x = logp_cell_finish(l1, c[0], l4);
                    r.y = logp_cell_finish(l1, c[1], l4);
                    r.z = logp_cell_finish(l1, c[2], l4);
                    r.w = logp_cell_finish(l1, c[3], l4);
                }
                *reinterpret_cast<float4 *>(row + yl) = r;
            }
        }
    }
}

// Single launch: only the sweep reads the scores, and only inside the reference's band
// (core.pyx:18: max(0, t_x + y - t_y) <= x < min(t_x, y + 1)).  A thread whose four tokens are
// outside the band for every frame of the chunk skips the contraction and stores zeros (the
// sweep's boxes still cover those cells: they must be finite, their value is irrelevant); tokens
// beyond t_x are not touched at all.  With full lengths (200 x 1000) that is 15 % of the cells,
// with ragged batches whatever the padding is.
__device__ __forceinline__ int band_mode(const TileShape &t, int x0, int ch, int band_tx, int band_ty) {
    const int rg = threadIdx.x / t.CG;
    const int y0 = ch * t.F, t0 = x0 + rg * kGemmTM;
    const int lo = max(0, band_tx + y0 - band_ty), hi = min(band_tx, min(y0 + t.F, band_ty));
    return (t0 >= band_tx) ? 0 : (t0 + kGemmTM > lo && t0 < hi) ? 2 : 1;
}

// chunks chunk_first, chunk_first + chunk_stride, ... (chunk_count of them) of token tile x0 of
// utterance b; indices from skip_from up are shifted by skip_by (the chunks the spare CTAs take).
// The materialising kernel's program: scores go to p.logp [B][T_x][T_y].
__device__ __forceinline__ void logp_cta(const LogpParams &p, float *sm, const TileShape &t, int b, int x0, int chunk_first,
                                         int chunk_stride, int chunk_count, int skip_from, int skip_by) {
    if (chunk_count <= 0) return;
    const int D = p.D, F = t.F;
    const CtaSmem s = carve_smem(sm, D, t);
    const bool mean_only = p.x_logs == nullptr;         // config.py:52, the reference default
    ChunkOut o{p.logp + (int64_t)b * p.T_x * p.T_y, p.T_y, 0, false};

    // the sequence is kept in un-shifted numbering; `ch` is the real chunk
    int seq = chunk_first;
    auto next_chunk = [&]() {
        seq += chunk_stride;
        return seq >= skip_from ? seq + skip_by : seq;
    };
    __syncthreads();                                    // a previous tile of this CTA is done with the shared memory
    int ch = seq >= skip_from ? seq + skip_by : seq;
    stage_frames_async(p, s, t, b, ch, 0);              // in flight while the token side is prepared
    stage_tokens(p, s, t, b, x0);
    // One CTA barrier per chunk: it makes chunk k's frames visible and says that every warp is done
    // with chunk k-1 (its frame buffer may be refilled).
    for (int k = 0; k < chunk_count; ++k) {
        const int buf = k & 1;
        const int ch_next = next_chunk();
        ptx::cp_async_wait<0>();
        __syncthreads();
        if (k + 1 < chunk_count) stage_frames_async(p, s, t, b, ch_next, buf ^ 1);   // lands while this chunk is contracted
        if (mean_only) frame_sums_mean_only(s, D, F, buf);
        o.col0 = ch * F;
        contract_chunk(p, s, t, x0, ch, buf, 2, o);
        ch = ch_next;
    }
}

// ---------------------------------------------------------------------------------------------
// Dealing the work to P persistent CTAs.  A "row" is one token tile of one utterance (BT = B x
// row_tiles of them), with nchunks chunks each; staging a row's token side costs about a third of a
// chunk, so a CTA should stay on its row.
//   * whole passes while rows outnumber CTAs: CTA p takes rows p, p + P, ... with all their chunks;
//   * the remaining `rem` rows (rem <= P) get d = P / rem dedicated CTAs each, CTA j of a row taking
//     chunks j, j + d, ... below `cover` -- so the scores of EARLY frames of every utterance exist
//     first, which is what the sweep CTAs of the fused launch wait for;
//   * when that would need one round more than the work justifies, `nchunks - cover` chunks per row
//     are dealt one at a time to the P - d rem spare CTAs instead: chunks [spare_first, spare_first +
//     nchunks - cover), taken from just before the dedicated CTAs' last round when the spares are
//     done by then (the last round then ends with the final, usually narrower, chunk and the sweep
//     has less left to do after it), else from the end.
//     (200 x 1000, 32 utterances, 116 producers: 3 dedicated CTAs x 4 rounds take chunks 0-8 and
//     10-12, 20 spares take the 32 chunks number 9.)
// ---------------------------------------------------------------------------------------------
struct Deal {
    int P, BT, nchunks;
    int passes;        // whole passes over P rows
    int rem;           // rows of the last, partial pass
    int d;             // dedicated CTAs per remaining row
    int cover;         // chunks the dedicated CTAs take per row; nchunks - cover go to the spares
    int spare_first;   // first chunk of the spares' range
    int spares;
};

inline Deal make_deal(int P, int BT, int nchunks) {
    Deal q{};
    q.P = P;
    q.BT = BT;
    q.nchunks = nchunks;
    q.passes = BT / P;
    q.rem = BT - q.passes * P;
    q.d = 0;
    q.cover = nchunks;
    q.spare_first = nchunks;
    q.spares = 0;
    if (q.rem == 0) return q;
    q.d = P / q.rem;
    if (q.d > nchunks) q.d = nchunks;
    q.spares = P - q.d * q.rem;
    // rounds if the dedicated CTAs do everything, against a cover that leaves the tail to the spares
    const int all = ceil_div(nchunks, q.d);
    if (q.spares > 0 && all > 1) {
        const int cover = q.d * (all - 1);                                 // one round less
        const int left = (nchunks - cover) * q.rem;                        // units for the spares
        const double spare_rounds = ceil_div(left, q.spares) * 1.35;       // each restages its token side
        if (spare_rounds <= all - 1) {
            q.cover = cover;
            q.spare_first = (spare_rounds <= all - 2) ? cover - q.d : cover;
        }
    }
    return q;
}

// The `it`-th piece of CTA `pidx`'s share: row `r`, chunks first, first + stride, ... (`count` of them,
// numbers from `skip_from` up shifted by the spares' range, see logp_cta).  False when the CTA is done.
// Host and device: the kernels walk it, the debug entry of the C ABI enumerates it for the CPU tests.
struct DealPiece {
    int r, first, stride, count, skip_from;
};
__host__ __device__ inline bool deal_piece(const Deal &q, int pidx, int it, DealPiece &o) {
    const int base = q.passes * q.P, dedicated = q.d * q.rem, left = (q.nchunks - q.cover) * q.rem;
    o.stride = 1;
    o.count = 1;
    o.skip_from = 0x7fffffff;
    if (it < q.passes) {                                // a whole row
        o.r = it * q.P + pidx;
        o.first = 0;
        o.count = q.nchunks;
        return true;
    }
    if (q.rem == 0) return false;
    if (pidx < dedicated) {                             // dedicated CTA j of a remaining row
        if (it > q.passes) return false;
        const int rr = pidx / q.d, j = pidx - rr * q.d;
        o.r = base + rr;
        o.first = j;
        o.stride = q.d;
        o.count = (q.cover - j + q.d - 1) / q.d;
        o.skip_from = q.spare_first;                    // (== cover or beyond when the spares take the end)
        return o.count > 0;
    }
    const int u = (pidx - dedicated) + (it - q.passes) * q.spares;   // spare: one left-over chunk at a time
    if (q.spares <= 0 || u >= left) return false;
    const int l = u / q.rem;
    o.r = base + (u - l * q.rem);
    o.first = q.spare_first + l;
    return true;
}

// Runs CTA `pidx`'s share (the materialising kernel).  One call site of logp_cta: the program is a
// few thousand instructions and instruction fetch is not free.
__device__ __forceinline__ void run_deal(const LogpParams &p, float *sm, const TileShape &t, const Deal &q, int pidx) {
    DealPiece o;
    for (int it = 0; deal_piece(q, pidx, it, o); ++it) {
        const int b = o.r / t.row_tiles, rt = o.r - b * t.row_tiles;
        logp_cta(p, sm, t, b, rt * t.tile_rows, o.first, o.stride, o.count, o.skip_from, q.nchunks - q.cover);
    }
}

}  // namespace logp
}  // namespace mas

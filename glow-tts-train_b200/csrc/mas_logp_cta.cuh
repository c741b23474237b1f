// mas_logp_cta.cuh -- the per-CTA program of the log-likelihood contraction (models.py:362-376):
// one tile of tokens of one utterance, a list of 64-frame chunks.  Shared by the materialising
// kernel (mas_logp.cu) and the fused launch (mas_fused.cu), where it also raises a ready flag per
// chunk for the sweep CTAs.  Needs T_y % 4 == 0, 16-byte aligned z / logp rows, D <= 80.
#pragma once

#include "mas_kernels.cuh"
#include "mas_logp_tile.cuh"
#include "mas_ptx.cuh"

namespace mas {
namespace logp {

constexpr int kPanel = 80;          // channels resident in shared memory at a time

// floats of shared memory the program needs
__host__ __device__ inline int cta_smem_floats(int D, int tile_rows) {
    return 2 * D * tile_rows + 2 * D * kGemmFrames + 2 * tile_rows + kGemmFrames;
}

// chunks chunk_first, chunk_first + chunk_stride, ... (chunk_count of them, < nchunks)
template <bool kSignal>
__device__ __forceinline__ void logp_cta(const LogpParams &p, float *sm, int tile_rows, int b, int x0, int chunk_first,
                                         int chunk_stride, int chunk_count, int *ready, long long *dbg_ns = nullptr,
                                         int *queue = nullptr, int nchunks = 0) {
    const int D = p.D, T_x = p.T_x, T_y = p.T_y;
    float *sInv = sm;                                   // [D][tile_rows]
    float *sMiv = sInv + D * tile_rows;                 // [D][tile_rows]
    float *sZ = sMiv + D * tile_rows;                   // [2][D][64]  (double-buffered chunk of z)
    float *sL1 = sZ + 2 * D * kGemmFrames;              // [tile_rows]
    float *sL4 = sL1 + tile_rows;                       // [tile_rows]
    float *sL2 = sL4 + tile_rows;                       // [64] mean_only: per-frame sum of -0.5 z^2
    const bool mean_only = p.x_logs == nullptr;         // config.py:52, the reference default

    const int tid = threadIdx.x, nthr = blockDim.x;
    const float *xm = p.x_m + (int64_t)b * D * T_x;
    const float *xl = p.x_logs ? p.x_logs + (int64_t)b * D * T_x : nullptr;
    const float *zg = p.z + (int64_t)b * D * T_y;
    float *out = p.logp + (int64_t)b * T_x * T_y;
    const int rg = tid >> 3, cg = tid & 7;              // kGemmTM tokens x {4+4} frames per thread
    const bool worker = rg * kGemmTM < tile_rows;

    auto stage_frames_async = [&](int ch, int buf) {
        const int y0 = ch * kGemmFrames;
        float *dst = sZ + buf * D * kGemmFrames;
        for (int i = tid; i < D * (kGemmFrames / 4); i += nthr) {
            const int d = i >> 4, y = y0 + ((i & 15) << 2);
            ptx::cp_async_16(dst + (i << 2), zg + (int64_t)d * T_y + (y < T_y ? y : 0), y < T_y);
        }
        ptx::cp_async_commit();
    };

    // Chunk order: a fixed arithmetic sequence, or (queue != null) whatever this tile's shared counter
    // hands out next -- CTAs that share their SM with a sweep CTA then simply take fewer chunks.
    __shared__ int s_next;
    auto take = [&]() -> int {                          // every thread gets the same answer
        __syncthreads();
        if (tid == 0) s_next = atomicAdd(queue, 1);
        __syncthreads();
        return s_next;
    };
    int ch = chunk_first, ch_next = 0;
    if (queue != nullptr) {
        ch = take();
        if (ch >= nchunks) return;
        chunk_count = nchunks;                          // upper bound; the loop ends when the queue is dry
    } else if (chunk_count <= 0) {
        return;
    }
    if (dbg_ns && tid == 0) dbg_ns[0] = ptx::globaltimer_ns();
    stage_frames_async(ch, 0);                          // in flight while the token side is prepared
    // token side: thread (x, h) stages token x0+x for one contiguous share of the channels (coalesced
    // over x, eight loads in flight) and sums its share of the row constants, channels ascending;
    // the shares are then added in order.  nsh = how many threads serve a token (2 at 224 / 112).
    const int nsh = max(1, min(4, nthr / tile_rows));
    const int dsh = ceil_div(D, nsh);
    float *sPart = sZ + D * kGemmFrames;                // buffer 1 of sZ is still free: [nsh][2][tile_rows]
    if (tid < nsh * tile_rows) {
        const int h = tid / tile_rows, x = tid - h * tile_rows, xg = x0 + x;
        const int d0 = h * dsh, d1 = min(D, d0 + dsh);
        float l1 = 0.f, l4 = 0.f;
        if (xg < T_x) {
#pragma unroll 8
            for (int d = d0; d < d1; ++d) {
                const float m = __ldg(xm + (int64_t)d * T_x + xg);
                const float ls = xl ? __ldg(xl + (int64_t)d * T_x + xg) : 0.f;
                const float r = xl ? expf(-2.0f * ls) : 1.0f;         // models.py:363
                sInv[d * tile_rows + x] = -0.5f * r;                    // models.py:368
                sMiv[d * tile_rows + x] = m * r;                        // models.py:371
                l1 += kNegHalfLog2Pi - ls;                              // models.py:364-366
                l4 = fmaf(-0.5f * (m * m), r, l4);                      // models.py:373-375
            }
        } else {
            for (int d = d0; d < d1; ++d) sInv[d * tile_rows + x] = sMiv[d * tile_rows + x] = 0.f;
        }
        sPart[(2 * h) * tile_rows + x] = l1;
        sPart[(2 * h + 1) * tile_rows + x] = l4;
    }
    __syncthreads();
    if (tid < tile_rows) {
        float l1 = sPart[tid], l4 = sPart[tile_rows + tid];
        for (int h = 1; h < nsh; ++h) {
            l1 += sPart[(2 * h) * tile_rows + tid];
            l4 += sPart[(2 * h + 1) * tile_rows + tid];
        }
        sL1[tid] = l1;
        sL4[tid] = l4;
    }
    __syncthreads();                                    // sPart is read before chunk 1 lands in that buffer
    float acc[kGemmTM][8];
    for (int k = 0; k < chunk_count; ++k) {
        const int buf = k & 1;
        bool more;
        if (queue != nullptr) {
            ch_next = take();
            more = ch_next < nchunks;
        } else {
            ch_next = ch + chunk_stride;
            more = k + 1 < chunk_count;
        }
        if (more) {
            stage_frames_async(ch_next, buf ^ 1);       // buffer buf^1 was released by the barrier below
            ptx::cp_async_wait<1>();
        } else {
            ptx::cp_async_wait<0>();
        }
        __syncthreads();                                // chunk ch (and the token side) visible to everyone
        if (mean_only) {
            // inv_var == 1: the inv_var term does not depend on the token (models.py:367-369 with
            // x_logs == 0): one sum per frame, channels ascending
            if (tid < kGemmFrames) {
                const float *zc = sZ + buf * D * kGemmFrames + tid;
                float l2 = 0.f;
                for (int d = 0; d < D; ++d) {
                    const float zv = zc[d * kGemmFrames];
                    l2 = fmaf(-0.5f * zv, zv, l2);
                }
                sL2[tid] = l2;
            }
            __syncthreads();
        }
        if (worker) {
            if (mean_only)
                gemm_tile<kGemmTM, true, true>(sInv, sMiv, sZ + buf * D * kGemmFrames, D, tile_rows, rg, cg, acc);
            else
                gemm_tile<kGemmTM, true, false>(sInv, sMiv, sZ + buf * D * kGemmFrames, D, tile_rows, rg, cg, acc);
            const int y0 = ch * kGemmFrames;
#pragma unroll
            for (int i = 0; i < kGemmTM; ++i) {
                const int xr = rg * kGemmTM + i, x = x0 + xr;
                if (x >= T_x) break;
                const float l1 = sL1[xr], l4 = sL4[xr];
                float *row = out + (int64_t)x * T_y;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int y = y0 + 32 * h + 4 * cg;
                    if (y < T_y) {                      // T_y % 4 == 0: whole float4 or nothing
                        float4 r;
                        if (mean_only) {
                            const float4 l2 = *reinterpret_cast<const float4 *>(sL2 + 32 * h + 4 * cg);
                            r.x = logp_cell_finish_mean_only(l1, l2.x, acc[i][4 * h + 0], l4);
                            r.y = logp_cell_finish_mean_only(l1, l2.y, acc[i][4 * h + 1], l4);
                            r.z = logp_cell_finish_mean_only(l1, l2.z, acc[i][4 * h + 2], l4);
                            r.w = logp_cell_finish_mean_only(l1, l2.w, acc[i][4 * h + 3], l4);
                        } else {
                            r.x = logp_cell_finish(l1, acc[i][4 * h + 0], l4);
                            r.y = logp_cell_finish(l1, acc[i][4 * h + 1], l4);
                            r.z = logp_cell_finish(l1, acc[i][4 * h + 2], l4);
                            r.w = logp_cell_finish(l1, acc[i][4 * h + 3], l4);
                        }
                        *reinterpret_cast<float4 *>(row + y) = r;
                    }
                }
            }
            if (kSignal) __threadfence();               // this thread's scores before the flag below
        }
        __syncthreads();                                // everyone is done with buffer buf (and has stored)
        if (kSignal && tid == 0) {
            ptx::red_release_gpu_add(ready + ch, 1);
            if (dbg_ns) dbg_ns[k < 15 ? k + 1 : 15] = ptx::globaltimer_ns();
        }
        if (!more) break;
        ch = ch_next;
    }
}

}  // namespace logp
}  // namespace mas

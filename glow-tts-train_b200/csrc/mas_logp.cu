// mas_logp.cu -- the [B,T_x,T_y] Gaussian log-likelihood matrix of FlowGenerator.forward
// (glow_tts_train/models.py:362-376), materialised.  FP32 FFMA contraction over the mel channels on
// the CUDA cores (K = 80 is too thin for a tensor-core pipeline to pay off; BASELINE.json north_star).
//
// One CTA owns a tile of tokens of one utterance (up to 208, all of a 200-token text) and walks a
// range of 64-frame chunks: the token-side operands (exp(-2 logs), m exp(-2 logs), 80 channels) are
// computed once and stay in shared memory, the frame-side operands (z, -0.5 z^2) are staged per
// chunk, and every thread contracts an 8 x 8 block of cells in registers (mas_logp_tile.cuh).
// The arithmetic (operands, FFMA order, final adds) is the one the fused kernel uses, so both
// produce bit-identical scores.
#include "mas_logp_cta.cuh"

namespace mas {
namespace logp {

constexpr int kMaxTileRows = 112;   // 28 token groups of 4 x 8 frame groups -> 224 threads; two CTAs share an SM

struct Geometry {
    int tile_rows;      // multiple of 8
    int row_tiles;      // ceil(T_x / tile_rows)
    int nchunks;        // ceil(T_y / 64)
    int splits;         // CTAs along the frame axis
    int chunks_per_cta;
    int threads;
    int panel;          // min(D, kPanel)
    int smem_bytes;
};

static Geometry make_geometry(int B, int D, int T_x, int T_y, int num_sms) {
    Geometry g;
    g.row_tiles = ceil_div(T_x, kMaxTileRows);
    g.tile_rows = ceil_div(ceil_div(T_x, g.row_tiles), 8) * 8;   // multiple of 8: 32-byte aligned operand rows
    g.nchunks = ceil_div(T_y, kGemmFrames);
    // CTAs along the frame axis: fewest (waves x chunks per CTA), counting ~0.7 chunk of token-side
    // staging per CTA
    const int64_t base = (int64_t)B * g.row_tiles;
    double best = 1e30;
    g.chunks_per_cta = g.nchunks;
    for (int cpc = 1; cpc <= g.nchunks; ++cpc) {
        const int64_t ctas = base * ceil_div(g.nchunks, cpc);
        const double cost = (double)((ctas + 2 * num_sms - 1) / (2 * num_sms)) * (cpc + 0.7);   // two CTAs per SM
        if (cost < best - 1e-9) {
            best = cost;
            g.chunks_per_cta = cpc;
        }
    }
    g.splits = ceil_div(g.nchunks, g.chunks_per_cta);
    g.threads = max(64, ceil_div(g.tile_rows / kGemmTM * 8, 32) * 32);
    g.panel = D < kPanel ? D : kPanel;
    g.smem_bytes = cta_smem_floats(g.panel, g.tile_rows) * 4;
    return g;
}

// grid: (splits, row_tiles, B)
__global__ void __launch_bounds__(224, 2) mas_logp_kernel(LogpParams p, Geometry g) {
    extern __shared__ __align__(16) float sm[];
    const int tile_rows = g.tile_rows, panel = g.panel;
    float *sInv = sm;                                   // [panel][tile_rows]
    float *sMiv = sInv + panel * tile_rows;             // [panel][tile_rows]
    float *sZ = sMiv + panel * tile_rows;               // [2][panel][64]  (double-buffered chunk of z)
    float *sL1 = sZ + 2 * panel * kGemmFrames;          // [tile_rows]
    float *sL4 = sL1 + tile_rows;                       // [tile_rows]

    const int b = blockIdx.z, x0 = blockIdx.y * tile_rows;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int D = p.D, T_x = p.T_x, T_y = p.T_y;
    const float *xm = p.x_m + (int64_t)b * D * T_x;
    const float *xl = p.x_logs ? p.x_logs + (int64_t)b * D * T_x : nullptr;
    const float *zg = p.z + (int64_t)b * D * T_y;
    float *out = p.logp + (int64_t)b * T_x * T_y;

    const int rg = tid >> 3, cg = tid & 7;              // kGemmTM tokens x {4+4} frames per thread
    const bool worker = rg * kGemmTM < tile_rows;
    const bool vec_ok = ((T_y & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    const int npanels = ceil_div(D, panel);
    // frames staged by cp.async (16 bytes, no registers, overlapped with the previous chunk's FFMAs)
    const bool async_z = npanels == 1 && ((T_y & 3) == 0) && ((reinterpret_cast<uintptr_t>(zg) & 15) == 0);

    // token side of one channel panel: thread x stages token x0+x for every channel (coalesced over
    // x, eight loads in flight) and sums its row constants on the way, channels ascending
    auto stage_tokens = [&](int pn, int d0, int dn, bool keep_consts) {
        for (int x = tid; x < tile_rows; x += nthr) {
            const int xg = x0 + x;
            float l1 = (pn == 0) ? 0.f : sL1[x], l4 = (pn == 0) ? 0.f : sL4[x];
            if (xg < T_x) {
#pragma unroll 8
                for (int d = 0; d < dn; ++d) {
                    const float m = __ldg(xm + (int64_t)(d0 + d) * T_x + xg);
                    const float ls = xl ? __ldg(xl + (int64_t)(d0 + d) * T_x + xg) : 0.f;
                    const float r = xl ? expf(-2.0f * ls) : 1.0f;         // models.py:363
                    sInv[d * tile_rows + x] = -0.5f * r;                    // models.py:368
                    sMiv[d * tile_rows + x] = m * r;                        // models.py:371
                    l1 += kNegHalfLog2Pi - ls;                              // models.py:364-366
                    l4 = fmaf(-0.5f * (m * m), r, l4);                      // models.py:373-375
                }
            } else {
                for (int d = 0; d < dn; ++d) sInv[d * tile_rows + x] = sMiv[d * tile_rows + x] = 0.f;
            }
            if (keep_consts) {
                sL1[x] = l1;
                sL4[x] = l4;
            }
        }
    };
    auto store_tile = [&](int y0, float (&acc)[kGemmTM][8]) {
#pragma unroll
        for (int i = 0; i < kGemmTM; ++i) {
            const int xr = rg * kGemmTM + i, x = x0 + xr;
            if (x >= T_x) break;
            const float l1 = sL1[xr], l4 = sL4[xr];
            float *row = out + (int64_t)x * T_y;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int y = y0 + 32 * h + 4 * cg;
                float4 r;
                r.x = logp_cell_finish(l1, acc[i][4 * h + 0], l4);
                r.y = logp_cell_finish(l1, acc[i][4 * h + 1], l4);
                r.z = logp_cell_finish(l1, acc[i][4 * h + 2], l4);
                r.w = logp_cell_finish(l1, acc[i][4 * h + 3], l4);
                if (vec_ok && y + 3 < T_y) {
                    *reinterpret_cast<float4 *>(row + y) = r;
                } else {
                    if (y < T_y) row[y] = r.x;
                    if (y + 1 < T_y) row[y + 1] = r.y;
                    if (y + 2 < T_y) row[y + 2] = r.z;
                    if (y + 3 < T_y) row[y + 3] = r.w;
                }
            }
        }
    };

    const int chunk0 = blockIdx.x * g.chunks_per_cta;
    const int chunk1 = min(chunk0 + g.chunks_per_cta, g.nchunks);
    if (async_z && ((reinterpret_cast<uintptr_t>(out) & 15) == 0)) {
        logp_cta<false>(p, sm, tile_rows, b, x0, chunk0, 1, chunk1 - chunk0, nullptr);
        return;
    }
    float acc[kGemmTM][8];
    // generic path: any alignment, any channel count (panels of 80)
    for (int ch = chunk0; ch < chunk1; ++ch) {
        const int y0 = ch * kGemmFrames;
        for (int pn = 0; pn < npanels; ++pn) {
            const int d0 = pn * panel, dn = min(panel, D - d0);
            __syncthreads();                            // previous contraction done with the staged operands
            if (npanels > 1 || ch == chunk0) stage_tokens(pn, d0, dn, ch == chunk0);
#pragma unroll 4
            for (int i = tid; i < dn * kGemmFrames; i += nthr) {    // frame side, coalesced over y
                const int d = i >> 6, yg = y0 + (i & 63);
                sZ[i] = (yg < T_y) ? __ldg(zg + (int64_t)(d0 + d) * T_y + yg) : 0.f;
            }
            __syncthreads();
            if (worker) {
                if (pn == 0)
                    gemm_tile<kGemmTM, true>(sInv, sMiv, sZ, dn, tile_rows, rg, cg, acc);
                else
                    gemm_tile<kGemmTM, false>(sInv, sMiv, sZ, dn, tile_rows, rg, cg, acc);
            }
        }
        if (worker) store_tile(y0, acc);
    }
}

}  // namespace logp

int launch_logp(const LogpParams &p, cudaStream_t stream) {
    using namespace logp;
    if (p.B == 0 || p.T_x == 0 || p.T_y == 0) return MAS_OK;
    static int num_sms_cached[64] = {0}, configured[64] = {0};
    int dev = 0;
    MAS_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return MAS_ERR_INVALID_ARGUMENT;
    if (num_sms_cached[dev] == 0)
        MAS_CUDA_TRY(cudaDeviceGetAttribute(&num_sms_cached[dev], cudaDevAttrMultiProcessorCount, dev));
    const Geometry g = make_geometry(p.B, p.D, p.T_x, p.T_y, num_sms_cached[dev]);
    if (g.smem_bytes > configured[dev]) {
        MAS_CUDA_TRY(cudaFuncSetAttribute(mas_logp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem_bytes));
        configured[dev] = g.smem_bytes;
    }
    dim3 grid(g.splits, g.row_tiles, p.B);
    mas_logp_kernel<<<grid, g.threads, g.smem_bytes, stream>>>(p, g);
    MAS_CUDA_TRY(cudaGetLastError());
    return MAS_OK;
}

}  // namespace mas

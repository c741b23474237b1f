// mas_logp.cu -- the [B,T_x,T_y] Gaussian log-likelihood matrix of FlowGenerator.forward
// (glow_tts_train/models.py:362-376), materialised.  FP32 FFMA contraction over the mel channels on
// the CUDA cores (K = 80 is too thin for a tensor-core pipeline to pay off; BASELINE.json north_star).
//
// One persistent CTA per SM (512 threads).  A CTA owns a tile of tokens of one utterance (up to 256,
// all of a 200-token text) and walks a list of chunks of 64..128 frames (mas_logp_cta.cuh): the
// token-side operands (-0.5 exp(-2 logs), m exp(-2 logs), 80 channels) are computed once and stay in
// shared memory, the frame-side operand z is staged per chunk with cp.async, and every thread
// contracts a 4 x 8 block of cells in registers (mas_logp_tile.cuh).
// The arithmetic (operands, FFMA order, final adds) is the one the fused kernel uses, so both
// produce bit-identical scores.
#include "mas_logp_cta.cuh"

namespace mas {
namespace logp {

struct Geometry {
    TileShape t;
    Deal deal;          // async path: persistent CTAs
    int panel;          // min(D, kPanel)
    int smem_bytes;
    int generic;        // unaligned rows or more than 80 channels: the plain path below
};

// grid: P persistent CTAs of 512 threads, one per SM
__global__ void __launch_bounds__(kGemmThreads, 1) mas_logp_kernel(LogpParams p, Geometry g) {
    extern __shared__ __align__(16) float sm[];
    if (!g.generic) {
        run_deal(p, sm, g.t, g.deal, blockIdx.x);
        return;
    }
    // ---- generic path: any alignment, any channel count (panels of 80); one (row, chunk) unit at a time ----
    const TileShape &t = g.t;
    const int tile_rows = t.tile_rows, panel = g.panel, F = t.F;
    float *sInv = sm;                                   // [panel][tile_rows]
    float *sMiv = sInv + panel * tile_rows;             // [panel][tile_rows]
    float *sZ = sMiv + panel * tile_rows;               // [panel][F] (the second buffer stays unused)
    float *sL1 = sZ + 2 * panel * F;                    // [tile_rows]
    float *sL4 = sL1 + tile_rows;                       // [tile_rows]
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int D = p.D, T_x = p.T_x, T_y = p.T_y;
    const int rg = tid / t.CG, cg = tid - rg * t.CG;
    const bool worker = rg < t.RG;
    const int npanels = ceil_div(D, panel);
    const int units = g.deal.BT * t.nchunks;
    GemmAcc acc;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int r = u / t.nchunks, ch = u - r * t.nchunks;
        const int b = r / t.row_tiles, x0 = (r - b * t.row_tiles) * tile_rows, y0 = ch * F;
        const float *xm = p.x_m + (int64_t)b * D * T_x;
        const float *xl = p.x_logs ? p.x_logs + (int64_t)b * D * T_x : nullptr;
        const float *zg = p.z + (int64_t)b * D * T_y;
        float *out = p.logp + (int64_t)b * T_x * T_y;
        for (int pn = 0; pn < npanels; ++pn) {
            const int d0 = pn * panel, dn = min(panel, D - d0);
            __syncthreads();                            // previous contraction done with the staged operands
            // token side of this channel panel: thread x stages token x0+x for every channel (coalesced
            // over x) and sums its row constants on the way, channels ascending
            for (int x = tid; x < tile_rows; x += nthr) {
                const int xg = x0 + x;
                float l1 = (pn == 0) ? 0.f : sL1[x], l4 = (pn == 0) ? 0.f : sL4[x];
                if (xg < T_x) {
#pragma unroll 8
                    for (int d = 0; d < dn; ++d) {
                        const float m = __ldg(xm + (int64_t)(d0 + d) * T_x + xg);
                        const float ls = xl ? __ldg(xl + (int64_t)(d0 + d) * T_x + xg) : 0.f;
                        const float rr = xl ? expf(-2.0f * ls) : 1.0f;        // models.py:363
                        sInv[d * tile_rows + x] = -0.5f * rr;                   // models.py:368
                        sMiv[d * tile_rows + x] = m * rr;                       // models.py:371
                        l1 += kNegHalfLog2Pi - ls;                              // models.py:364-366
                        l4 = fmaf(-0.5f * (m * m), rr, l4);                     // models.py:373-375
                    }
                } else {
                    for (int d = 0; d < dn; ++d) sInv[d * tile_rows + x] = sMiv[d * tile_rows + x] = 0.f;
                }
                sL1[x] = l1;
                sL4[x] = l4;
            }
            for (int i = tid; i < dn * F; i += nthr) {  // frame side, coalesced over y
                const int d = i / F, yg = y0 + (i - d * F);
                sZ[i] = (yg < T_y) ? __ldg(zg + (int64_t)(d0 + d) * T_y + yg) : 0.f;
            }
            __syncthreads();
            if (worker) {
                if (pn == 0)
                    gemm_tile<true>(sInv, sMiv, sZ, dn, tile_rows, F, rg, cg, acc);
                else
                    gemm_tile<false>(sInv, sMiv, sZ, dn, tile_rows, F, rg, cg, acc);
            }
        }
        if (worker) {
#pragma unroll
            for (int i = 0; i < kGemmTM; ++i) {
                const int xr = rg * kGemmTM + i, x = x0 + xr;
                if (x >= T_x) break;
                const float l1 = sL1[xr], l4 = sL4[xr];
                float *row = out + (int64_t)x * T_y;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float c[4];
                    acc.quad(i, h, c);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int y = y0 + (F >> 1) * h + 4 * cg + j;
                        if (y < T_y) row[y] = logp_cell_finish(l1, c[j], l4);
                    }
                }
            }
        }
    }
}

}  // namespace logp

int launch_logp(const LogpParams &p, cudaStream_t stream) {
    using namespace logp;
    if (p.B == 0 || p.T_x == 0 || p.T_y == 0) return MAS_OK;
    static SmemOptIn optin;
    int dev = 0;
    MAS_CUDA_TRY(cudaGetDevice(&dev));
    DeviceInfo di{};
    if (int rc = get_device_info(dev, di)) return rc;
    const int num_sms = di.num_sms;
    Geometry g{};
    g.t = make_tile_shape(p.T_x, p.T_y);
    g.panel = p.D < kPanel ? p.D : kPanel;
    g.generic = (p.D > kPanel) || (p.T_y & 3) || (reinterpret_cast<uintptr_t>(p.z) & 15) || (reinterpret_cast<uintptr_t>(p.logp) & 15);
    const int BT = p.B * g.t.row_tiles;
    const int64_t units = (int64_t)BT * g.t.nchunks;
    const int P = (int)(units < num_sms ? units : num_sms);
    g.deal = make_deal(P, BT, g.t.nchunks);
    g.smem_bytes = cta_smem_floats(g.panel, g.t) * 4;
    if (int rc = optin.ensure(mas_logp_kernel, dev, g.smem_bytes)) return rc;
    mas_logp_kernel<<<P, kGemmThreads, g.smem_bytes, stream>>>(p, g);
    MAS_CUDA_TRY(cudaGetLastError());
    return MAS_OK;
}

// Host-side enumeration of the deal (no device needed): which CTA takes chunk c of row r, and as
// the how-manieth chunk of its share.  Returns the number of (row, chunk) units assigned more than
// once (0 for a correct deal); units nobody takes keep owner -1.
int debug_deal(int P, int BT, int nchunks, int32_t *owner, int32_t *order) {
    using namespace logp;
    if (P <= 0 || BT <= 0 || nchunks <= 0 || !owner || !order) return -1;
    for (int64_t i = 0; i < (int64_t)BT * nchunks; ++i) owner[i] = order[i] = -1;
    const Deal q = make_deal(P, BT, nchunks);
    int doubles = 0;
    for (int pidx = 0; pidx < P; ++pidx) {
        int pos = 0;
        DealPiece o;
        for (int it = 0; deal_piece(q, pidx, it, o); ++it)
            for (int k = 0; k < o.count; ++k) {
                int c = o.first + k * o.stride;
                if (c >= o.skip_from) c += q.nchunks - q.cover;
                if (o.r < 0 || o.r >= BT || c < 0 || c >= nchunks) return -2;
                int32_t &w = owner[(int64_t)o.r * nchunks + c];
                if (w >= 0) ++doubles;
                w = pidx;
                order[(int64_t)o.r * nchunks + c] = pos++;
            }
    }
    return doubles;
}

void debug_tile_shape(int T_x, int T_y, int32_t *out6) {
    const TileShape t = make_tile_shape(T_x, T_y);
    out6[0] = t.row_tiles, out6[1] = t.tile_rows, out6[2] = t.RG, out6[3] = t.CG, out6[4] = t.F, out6[5] = t.nchunks;
}

}  // namespace mas

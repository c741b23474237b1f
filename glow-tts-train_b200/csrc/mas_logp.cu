// mas_logp.cu -- the [B,T_x,T_y] Gaussian log-likelihood matrix of FlowGenerator.forward
// (glow_tts_train/models.py:362-376), materialised.  FP32 FFMA contraction over the mel channels on
// the CUDA cores (K = 80 is too thin for a tensor-core pipeline to pay off; BASELINE.json north_star).
//
//   logp[b,x,y] = ((l1[x] + l2[x,y]) + l3[x,y]) + l4[x]                      models.py:376
//   l1[x]   = sum_d (-0.5 log(2 pi) - logs[d,x])                             models.py:364-366
//   l2[x,y] = sum_d inv_var[d,x] * (-0.5 z[d,y]^2),  inv_var = exp(-2 logs)  models.py:363,367-369
//   l3[x,y] = sum_d (m[d,x] inv_var[d,x]) * z[d,y]                           models.py:370-372
//   l4[x]   = sum_d -0.5 m[d,x]^2 inv_var[d,x]                               models.py:373-375
//
// Every cell is contracted by logp_cell_fma() in ascending channel order; the fused kernel uses the
// same routine so that both produce bit-identical scores.
#include "mas_kernels.cuh"
#include "mas_logp_tile.cuh"

namespace mas {
namespace logp {

constexpr int kTileX = 64, kTileY = 64, kThreads = 256, kChunkD = 40;

// grid: (ceil(T_y/64), ceil(T_x/64), B); block 256 = 16 (token groups of 4) x 16 (frame groups of 4)
__global__ void __launch_bounds__(kThreads) mas_logp_kernel(LogpParams p) {
    __shared__ __align__(16) float s_inv[kChunkD][kTileX];   // inv_var[d][x]
    __shared__ __align__(16) float s_miv[kChunkD][kTileX];   // m * inv_var
    __shared__ __align__(16) float s_z[kChunkD][kTileY];     // z[d][y]
    __shared__ __align__(16) float s_zz[kChunkD][kTileY];    // -0.5 z^2
    __shared__ float s_l1[kTileX], s_l4[kTileX];

    const int b = blockIdx.z, x0 = blockIdx.y * kTileX, y0 = blockIdx.x * kTileY;
    const int tid = threadIdx.x;
    const int D = p.D, T_x = p.T_x, T_y = p.T_y;
    const float *xm = p.x_m + (int64_t)b * D * T_x;
    const float *xl = p.x_logs ? p.x_logs + (int64_t)b * D * T_x : nullptr;
    const float *zz = p.z + (int64_t)b * D * T_y;

    if (tid < kTileX) {
        float l1, l4;
        row_constants(xm, xl, D, T_x, x0 + tid, l1, l4);
        s_l1[tid] = l1;
        s_l4[tid] = l4;
    }

    const int rx = (tid >> 4) * 4, cy = (tid & 15) * 4;
    float acc2[4][4], acc3[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc2[i][j] = acc3[i][j] = 0.f;

    for (int d0 = 0; d0 < D; d0 += kChunkD) {
        const int dn = min(kChunkD, D - d0);
        __syncthreads();
        for (int i = tid; i < dn * kTileX; i += kThreads) {
            const int d = i / kTileX, x = i % kTileX;
            float inv, miv;
            token_operands(xm, xl, T_x, d0 + d, x0 + x, inv, miv);
            s_inv[d][x] = inv;
            s_miv[d][x] = miv;
        }
        for (int i = tid; i < dn * kTileY; i += kThreads) {
            const int d = i / kTileY, y = i % kTileY;
            float zv, zsq;
            frame_operands(zz, T_y, d0 + d, y0 + y, zv, zsq);
            s_z[d][y] = zv;
            s_zz[d][y] = zsq;
        }
        __syncthreads();
        for (int d = 0; d < dn; ++d) {
            const float4 inv = *reinterpret_cast<const float4 *>(&s_inv[d][rx]);
            const float4 miv = *reinterpret_cast<const float4 *>(&s_miv[d][rx]);
            const float4 zv = *reinterpret_cast<const float4 *>(&s_z[d][cy]);
            const float4 zq = *reinterpret_cast<const float4 *>(&s_zz[d][cy]);
            const float a[4] = {inv.x, inv.y, inv.z, inv.w}, m[4] = {miv.x, miv.y, miv.z, miv.w};
            const float zc[4] = {zv.x, zv.y, zv.z, zv.w}, qc[4] = {zq.x, zq.y, zq.z, zq.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) logp_cell_fma(acc2[i][j], acc3[i][j], a[i], m[i], qc[j], zc[j]);
        }
    }
    __syncthreads();
    float *out = p.logp + (int64_t)b * T_x * T_y;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int x = x0 + rx + i;
        if (x >= T_x) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int y = y0 + cy + j;
            if (y < T_y) out[(int64_t)x * T_y + y] = logp_cell_finish(s_l1[rx + i], acc2[i][j], acc3[i][j], s_l4[rx + i]);
        }
    }
}

}  // namespace logp

int launch_logp(const LogpParams &p, cudaStream_t stream) {
    using namespace logp;
    if (p.B == 0 || p.T_x == 0 || p.T_y == 0) return MAS_OK;
    dim3 grid(ceil_div(p.T_y, kTileY), ceil_div(p.T_x, kTileX), p.B);
    mas_logp_kernel<<<grid, kThreads, 0, stream>>>(p);
    MAS_CUDA_TRY(cudaGetLastError());
    return MAS_OK;
}

}  // namespace mas

// mas_fused.cu -- kernel (2): log-likelihood + alignment search in ONE launch (models.py:362-382).
//
// One CTA per SM, two kinds sharing the grid and running concurrently:
//   * producers (blockIdx < P): the FFMA contraction of mas_logp_cta.cuh, 16 warps, a token tile of
//     one utterance resident in shared memory, walking chunks in an interleaved order (Deal:
//     producer j of the d on a tile takes chunks j, j+d, ...) so that the scores of EARLY frames of
//     every utterance exist first; after each chunk is stored they raise that chunk's ready flag
//     (fence + red.release.gpu);
//   * one sweep CTA per utterance: kernel (1)'s program (mas_dp_cta.cuh) whose loader warp waits for
//     the flags of the chunks a TMA box reads (ld.acquire.gpu + fence.proxy.async), so the systolic
//     sweep trails the producers by a few blocks instead of starting after the last FFMA.
// A sweep warp runs alone on its scheduler at ~0.45 IPC; sharing the SM with FFMA warps halved its
// speed and made it the tail of the launch (profiles/r1_fused_timeline.txt, first layout), so a sweep
// CTA now has its SM to itself: the shared-memory request is sized so that only one CTA fits.
// Producers never wait for anybody and carry the lower block indices, so the launch cannot
// deadlock whatever the residency.  With more utterances than a third of the SMs the two kernels
// run back to back instead (the caller's fallback).
//
// The scores travel through a [B,T_x,T_y] fp32 scratch in the caller's workspace, written once and
// read once while still in the 126 MB L2 (25.6 MB for B=32, 200x1000).
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstdio>
#include <cstdlib>

#include "mas_dp_cta.cuh"
#include "mas_logp_cta.cuh"

namespace mas {
namespace fused {

constexpr int kThreads = kGemmThreads;          // producers: 16 warps | sweep CTAs use W + 1 of them

struct Geometry {
    int P;             // producer CTAs
    TileShape t;
    logp::Deal deal;
    int *ready;        // [B][nchunks] chunk ready counters
    int teams;         // utterances per sweep CTA: 2 (a half CTA each) when two plans fit one SM, else 1
    int team_stride;   // bytes of shared memory per team
    int head;          // chunks 0 .. head-1 of an utterance are produced by its own sweep CTA first
};

template <int R, bool kDbg>
__global__ void __launch_bounds__(kThreads, 1)
mas_fused_kernel(const __grid_constant__ CUtensorMap tmap, PathParams pp, systolic::Plan plan, LogpParams lp, Geometry g) {
    extern __shared__ __align__(1024) unsigned char smem[];
    if ((int)blockIdx.x < g.P) {
        // profiling: producers stamp globaltimer per chunk behind the sweep CTAs' [B][16][16] block
        long long *dbg_ns = (kDbg && pp.dbg_cycles) ? pp.dbg_cycles + ((size_t)pp.B * 16 + blockIdx.x) * 16 : nullptr;
        logp::run_deal<true>(lp, reinterpret_cast<float *>(smem), g.t, g.deal, blockIdx.x, g.ready, g.head, dbg_ns);
    } else if (g.teams == 1) {
        const int b = blockIdx.x - g.P;
        // experiment hook (launch_fused): head chunks of the own utterance first
        if (g.head > 0) {
            for (int rt = 0; rt < g.t.row_tiles; ++rt)
                logp::logp_cta<true>(lp, reinterpret_cast<float *>(smem), g.t, b, rt * g.t.tile_rows, 0, 1, g.head, 0x7fffffff, 0, 0,
                                     g.ready + (size_t)b * g.t.nchunks, nullptr);
            __syncthreads();                            // the shared memory changes hands
            ptx::fence_proxy_async();                   // generic-proxy writes before the TMA boxes land there
        }
        systolic::dp_cta<R, kDbg, false, true>(tmap, pp, plan, smem, b, b, g.ready + (size_t)b * g.t.nchunks, g.t.row_tiles,
                                               g.t.F, g.t.nchunks);
    } else {
        // two utterances per sweep CTA, half the threads and half the shared memory each, a hardware
        // barrier of their own: the sweeps then hold B/2 SMs instead of B and the producers get the rest
        constexpr int kTeam = kThreads / 2;
        const int team = threadIdx.x / kTeam;
        const int b = 2 * ((int)blockIdx.x - g.P) + team;
        if (b >= pp.B) return;
        systolic::dp_cta<R, kDbg, false, true>(tmap, pp, plan, smem + (size_t)team * g.team_stride, b, b,
                                               g.ready + (size_t)b * g.t.nchunks, g.t.row_tiles, g.t.F, g.t.nchunks,
                                               systolic::Team{(int)threadIdx.x - team * kTeam, kTeam, 1 + team});
    }
}

template <int R>
static int launch_r(const CUtensorMap &tmap, const PathParams &pp, const systolic::Plan &plan, const LogpParams &lp,
                    const Geometry &g, int smem_bytes, cudaStream_t stream) {
    static int configured[64] = {0};
    int dev = 0;
    MAS_CUDA_TRY(cudaGetDevice(&dev));
    if (pp.dbg_cycles != nullptr) {
        MAS_CUDA_TRY(cudaFuncSetAttribute(mas_fused_kernel<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        mas_fused_kernel<R, true><<<g.P + (pp.B + g.teams - 1) / g.teams, kThreads, smem_bytes, stream>>>(tmap, pp, plan, lp, g);
        MAS_CUDA_TRY(cudaGetLastError());
        return MAS_OK;
    }
    if (smem_bytes > configured[dev & 63]) {
        MAS_CUDA_TRY(cudaFuncSetAttribute(mas_fused_kernel<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        configured[dev & 63] = smem_bytes;
    }
    if (getenv("MAS_B200_DEBUG") != nullptr) {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, mas_fused_kernel<R, false>, kThreads, smem_bytes);
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, mas_fused_kernel<R, false>);
        fprintf(stderr, "[mas_b200] fused kernel: %d CTAs/SM resident, %d regs, %zu B static smem\n", nb, fa.numRegs, fa.sharedSizeBytes);
    }
    mas_fused_kernel<R, false><<<g.P + (pp.B + g.teams - 1) / g.teams, kThreads, smem_bytes, stream>>>(tmap, pp, plan, lp, g);
    MAS_CUDA_TRY(cudaGetLastError());
    return MAS_OK;
}

}  // namespace fused

static size_t fused_flag_bytes(int B, int T_y) {      // ready counters, one per chunk (chunks are >= 64 frames)
    return align_up((size_t)B * (ceil_div(T_y, 8 * kGemmMinCG) + 1) * 4, 256);
}

size_t fused_workspace_bytes(int B, int D, int T_x, int T_y) {
    (void)D;
    return align_up((size_t)B * T_x * T_y * 4, 256) + fused_flag_bytes(B, T_y) + path_systolic_workspace_bytes(B, T_x, T_y);
}

// MAS_OK: launched.  MAS_ERR_UNSUPPORTED_SHAPE: not for the single-launch path (the caller runs the two
// kernels back to back instead).
int launch_fused(const LogpParams &lp_in, const int32_t *x_len, const int32_t *y_len, float *path, int32_t *durations,
                 int32_t *frame_token, void *workspace, size_t workspace_bytes, float max_neg_val, cudaStream_t stream) {
    using namespace fused;
    static const bool debug = getenv("MAS_B200_DEBUG") != nullptr;
#define MAS_FUSED_NO(why) do { if (debug) fprintf(stderr, "[mas_b200] single launch not taken: %s\n", why); return MAS_ERR_UNSUPPORTED_SHAPE; } while (0)
    const int B = lp_in.B, D = lp_in.D, T_x = lp_in.T_x, T_y = lp_in.T_y;
    if (B == 0) return MAS_OK;
    if ((T_y & 3) || T_y < systolic::kBlk || D > logp::kPanel || D < 1 ||
        (reinterpret_cast<uintptr_t>(lp_in.z) & 15) || (reinterpret_cast<uintptr_t>(path) & 15))
        MAS_FUSED_NO("alignment / frame count / channel count");
    int R, W;
    if (!systolic::choose_shape(T_x, R, W) || (W + 1) * 32 > kThreads) MAS_FUSED_NO("too many sweep warps");
    PFN_cuTensorMapEncodeTiled_v12000 encode = systolic::get_encode_fn();
    if (encode == nullptr) MAS_FUSED_NO("no cuTensorMapEncodeTiled");
    if (workspace == nullptr || workspace_bytes < fused_workspace_bytes(B, D, T_x, T_y)) return MAS_ERR_WORKSPACE_TOO_SMALL;

    static int max_smem_cached[64] = {0}, num_sms_cached[64] = {0};
    int dev = 0;
    MAS_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return MAS_ERR_INVALID_ARGUMENT;
    if (max_smem_cached[dev] == 0) {
        MAS_CUDA_TRY(cudaDeviceGetAttribute(&max_smem_cached[dev], cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        MAS_CUDA_TRY(cudaDeviceGetAttribute(&num_sms_cached[dev], cudaDevAttrMultiProcessorCount, dev));
    }
    const int max_smem = max_smem_cached[dev] - 2048;
    const int num_sms = num_sms_cached[dev];

    Geometry g{};
    g.t = make_tile_shape(T_x, T_y);
    const int BT = B * g.t.row_tiles;
    const int gemm_smem = logp::cta_smem_floats(D, g.t) * 4;
    if (gemm_smem > max_smem) MAS_FUSED_NO("shared memory (producers)");

    // sweep CTAs (K = 1): one utterance per CTA with the deepest ring that fits.
    // Experiment hook MAS_B200_FUSED_TWO_TEAMS: two utterances per sweep CTA (a half CTA -- 8 warps --
    // and half the shared memory each), so that the sweeps hold B/2 SMs and the producers get 16
    // more.  Measured at C2: the producers then finish at 58 us instead of 69, but the sweeps -- two
    // per scheduler, ring only two boxes deep -- take 55 us instead of 42 and become the critical
    // path: 110 us per step against 94.  Off by default.
    static const bool two_teams = getenv("MAS_B200_FUSED_TWO_TEAMS") != nullptr;
    systolic::Plan plan{};
    bool ok = false;
    g.teams = 1;
    if (two_teams && B > 1) {
        const int groups = ceil_div(T_x, systolic::kBlk);
        for (int r : {R, 2, 3, 5, 4, 6, 8}) {
            const int w = ceil_div(groups, r);
            if (w + 1 > kThreads / 64) continue;                      // 8 warps per team
            for (int S = 4; S >= 2 && !ok; --S) {
                systolic::Plan pl = systolic::make_plan(r, w, S, 1, T_y, true, 8192);
                const int stride = (int)align_up((size_t)pl.total, 1024);
                if (2 * stride <= max_smem) {
                    plan = pl;
                    g.teams = 2;
                    g.team_stride = stride;
                    ok = true;
                }
            }
            if (ok) break;
        }
    }
    for (int bits_smem = 1; bits_smem >= 0 && !ok; --bits_smem)
        for (int S = 4; S >= 2 && !ok; --S) {
            plan = systolic::make_plan(R, W, S, 1, T_y, bits_smem != 0, 8192);
            ok = plan.total <= max_smem;
        }
    if (!ok) MAS_FUSED_NO("shared memory (sweep)");
    const int sweep_smem = g.teams == 2 ? 2 * g.team_stride : plan.total;
    int smem_bytes = gemm_smem > sweep_smem ? gemm_smem : sweep_smem;
    const int solo = (max_smem_cached[dev] + 1024) / 2;                 // more than half an SM: one CTA per SM
    if (smem_bytes < solo) smem_bytes = solo;

    // the SMs the sweeps do not hold produce; with fewer than two producers per token tile they
    // would be the long pole and two full-width launches are faster
    g.P = num_sms - ceil_div(B, g.teams);
    if (g.P < 2 * BT) MAS_FUSED_NO("too many utterances for one wave");
    // Experiment hook MAS_B200_FUSED_HEAD=n: the sweep CTA of an utterance contracts its first n
    // chunks itself (its SM idles until the first scores exist) and the producers' deal covers the
    // rest.  Measured at C2 (graph of 10 steps): n = 0 / 1 / 2 / 3 -> 99 / 109 / 120 / 137 us per step.
    // A head chunk costs the sweep CTA a full unit although most of its cells are below the diagonal
    // (17-20 us: the sweep starts at 27 / 44 us instead of 24), the sweep itself needs 55-60 us after
    // its start, and 12 or 11 chunks over three producers are still four rounds.  Off by default.
    g.head = 0;
    if (g.teams == 1) {
        static const char *head_env = getenv("MAS_B200_FUSED_HEAD");
        g.head = head_env ? atoi(head_env) : 0;
        const int cap = g.t.nchunks / 4;                // the sweep still has most of the utterance to wait for
        if (g.head > cap) g.head = cap;
        if (g.head < 0) g.head = 0;
    }
    g.deal = logp::make_deal(g.P, BT, g.t.nchunks - g.head);

    unsigned char *ws = static_cast<unsigned char *>(workspace);
    float *scores = reinterpret_cast<float *>(ws);
    ws += align_up((size_t)B * T_x * T_y * 4, 256);
    g.ready = reinterpret_cast<int *>(ws);
    ws += fused_flag_bytes(B, T_y);
    MAS_CUDA_TRY(cudaMemsetAsync(g.ready, 0, (size_t)B * g.t.nchunks * 4, stream));

    LogpParams lp = lp_in;
    lp.logp = scores;
    lp.x_len = x_len;
    lp.y_len = y_len;
    PathParams pp{};
    pp.value = scores;
    pp.value_stride_b = (int64_t)T_x * T_y;
    pp.value_stride_x = T_y;
    pp.t_x = x_len;
    pp.t_y = y_len;
    pp.path = path;
    pp.durations = durations;
    pp.frame_token = frame_token;
    pp.ws_bits = plan.bits_in_smem ? nullptr : reinterpret_cast<uint32_t *>(ws);
    pp.B = B;
    pp.T_x = T_x;
    pp.T_y = T_y;
    pp.max_neg_val = max_neg_val;
    pp.dbg_cycles = g_dbg_cycles.load();

    CUtensorMap tmap;
    const cuuint64_t gdim[3] = {(cuuint64_t)T_y, (cuuint64_t)T_x, (cuuint64_t)B};
    const cuuint64_t gstride[2] = {(cuuint64_t)T_y * 4, (cuuint64_t)T_x * T_y * 4};
    const cuuint32_t box[3] = {(cuuint32_t)systolic::kBlk, (cuuint32_t)(systolic::kBlk * plan.R), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, scores, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) MAS_FUSED_NO("tensor map encode failed");
    if (debug)
        fprintf(stderr, "[mas_b200] single launch: %d producers (%d per tile + %d spares, cover %d of %d chunks of %d frames after %d head chunks, tiles of %d tokens) + %d sweep CTAs of %d utterance(s), %d B smem, R=%d W=%d S=%d\n",
                g.P, g.deal.d, g.deal.spares, g.deal.cover, g.deal.nchunks, g.t.F, g.head, g.t.tile_rows, ceil_div(B, g.teams), g.teams, smem_bytes,
                plan.R, plan.W, plan.S);

    switch (plan.R) {
        case 1: return launch_r<1>(tmap, pp, plan, lp, g, smem_bytes, stream);
        case 2: return launch_r<2>(tmap, pp, plan, lp, g, smem_bytes, stream);
        case 3: return launch_r<3>(tmap, pp, plan, lp, g, smem_bytes, stream);
        case 4: return launch_r<4>(tmap, pp, plan, lp, g, smem_bytes, stream);
        case 5: return launch_r<5>(tmap, pp, plan, lp, g, smem_bytes, stream);
        case 6: return launch_r<6>(tmap, pp, plan, lp, g, smem_bytes, stream);
        case 8: return launch_r<8>(tmap, pp, plan, lp, g, smem_bytes, stream);
        default: return MAS_ERR_UNSUPPORTED_SHAPE;
    }
}

}  // namespace mas

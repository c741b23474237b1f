// mas_path_systolic.cu -- kernel (1), the B200 fast path.
//
// One CLUSTER of K CTAs per utterance (K = 1, 2, 4 or 8), the utterance's tokens sharded over the
// CTAs in contiguous ranges; inside a CTA every sweep warp owns 32R consecutive tokens, R per lane.
// The forward sweep of core.pyx:17-30 is a systolic array over TOKENS: the running scores live in
// registers, lane l+1 gets lane l's last score through one warp shuffle per mel frame, warp w+1
// gets warp w's last score through a small shared-memory ring published once per 32-frame block,
// and the first warp of CTA c+1 gets the last warp of CTA c's through the same ring written over
// distributed shared memory.  Nobody meets at a barrier per frame: warps (and CTAs) are skewed in
// time, and the skew is free because token x is outside the reference's band before frame x
// anyway (core.pyx:18).
//
// Scores are streamed from HBM by TMA: each warp keeps its own S-deep ring of [32R tokens x 32
// frames] boxes (128-byte rows, SWIZZLE_128B so that the per-thread 16-byte row reads are
// bank-conflict free), refilled by the warp's lane 0 as soon as a box has been consumed.
//
// Per cell the sweep issues four instructions: FMNMX.NAN (best predecessor), FADD (new score),
// FADD (stay - advance, whose SIGN is the backtrack direction: advance > stay <=> stay - advance < 0,
// exact in IEEE arithmetic without flush-to-zero) and one funnel shift that pushes that sign bit
// into the token's direction word.  This equals the reference's compare/select bit for bit as long
// as every score is finite; a non-finite score is sticky under max.NaN, so it is detected once per
// token at the end of the sweep and the (rare) utterance is redone by an exact compare/select
// sweep inside the same kernel.
//
// Directions are packed 1 bit per cell (32 frames of one token per word) into the owning CTA's
// shared memory, or into the caller's workspace when they do not fit.  The backtrack of
// core.pyx:32-35 walks TOKENS, not frames: for token x it finds, with one count-leading-zeros, the
// frame at which the path stepped onto x; the walk is handed from CTA to CTA down the token
// ranges.  A dedicated warp per CTA zero-fills its slice of the dense output with bulk async
// copies while the sweep runs; the ones are written last.
#include <cuda.h>
#include <cstdlib>
#include <cudaTypedefs.h>

#include "mas_dp_cta.cuh"

namespace mas {
namespace systolic {

// kThreads: launch bound.  Up to 4 sweep warps (+ the filler) run as 160 threads so that the
// compiler may keep loop-invariants in registers.
template <int R, int kThreads, bool kDbg, bool kCluster>
__global__ void __launch_bounds__(kThreads, 1)
mas_path_systolic_kernel(const __grid_constant__ CUtensorMap tmap, PathParams p, Plan plan) {
    extern __shared__ __align__(1024) unsigned char smem[];
    // programmatic dependent launch (mas_fused.cu): back-to-back instances overlap the launch itself
    ptx::grid_launch_dependents();
    ptx::grid_dependency_wait();
    const int K = kCluster ? plan.K : 1;
    dp_cta<R, kDbg, kCluster>(tmap, p, plan, smem, blockIdx.x / K, blockIdx.x);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static std::atomic<int> g_force_cluster{0};   // testing hook (mas_b200_debug_force_cluster): 0 = heuristic

// The whole launch geometry: CTAs per utterance, warp shape, ring depth, where the direction bits go.
static bool choose_plan(int B, int T_x, int T_y, int max_smem, int num_sms, Plan &best) {
    auto try_k = [&](int K, Plan &out) -> int {       // 0: does not fit; 1: bits in workspace; 2: bits in smem
        int R, W;
        if (!choose_shape(ceil_div(T_x, K), R, W)) return 0;
        for (int bits_smem = 1; bits_smem >= 0; --bits_smem)
            for (int S = 4; S >= 2; --S) {
                Plan pl = make_plan(R, W, S, K, T_y, bits_smem != 0);
                if (pl.total <= max_smem && pl.total >= 2 * 4 * K * pl.rows) {
                    out = pl;
                    return bits_smem ? 2 : 1;
                }
            }
        return 0;
    };
    const int forced = g_force_cluster.load();
    if (forced > 0) return try_k(forced, best) != 0;
    // Measured on B200 (profiles/sweep_k.py): one CTA per utterance wins whenever it fits with the
    // direction bits in shared memory (a DSMEM hop costs more than it buys); more CTAs per
    // utterance only pay for CAPACITY -- long texts whose boxes or bits do not fit one SM -- and
    // then only while the whole grid stays resident.
    Plan cand[4];
    int q[4];
    for (int i = 0, K = 1; i < 4; ++i, K *= 2) q[i] = (K == 1 || ceil_div(T_x, K) >= kBlk) ? try_k(K, cand[i]) : 0;
    int pick = -1;
    for (int i = 0; i < 4 && pick < 0; ++i)          // smallest resident K with the bits in shared memory
        if (q[i] == 2 && (int64_t)B * (1 << i) <= num_sms) pick = i;
    if (pick < 0 && q[0] != 0) pick = 0;                // one CTA per utterance, bits in the workspace
    for (int i = 3; i >= 0 && pick < 0; --i)          // does not fit one SM: largest resident K ...
        if (q[i] != 0 && (int64_t)B * (1 << i) <= num_sms) pick = i;
    for (int i = 0; i < 4 && pick < 0; ++i)           // ... else the smallest K that fits at all
        if (q[i] != 0) pick = i;
    if (pick < 0) return false;
    best = cand[pick];
    return true;
}

template <int R, int kThreads>
static int launch_rt(const CUtensorMap &tmap, const PathParams &p, const Plan &plan, cudaStream_t stream) {
    static SmemOptIn optin[4];                        // opt-in attribute is sticky per device: raise it only when needed
    int dev = 0;
    MAS_CUDA_TRY(cudaGetDevice(&dev));
    const bool dbgk = p.dbg_cycles != nullptr;        // profiling build of the same kernel (clock64 stamps)
    const bool clus = plan.K > 1;
    auto kern = clus ? (dbgk ? mas_path_systolic_kernel<R, kThreads, true, true> : mas_path_systolic_kernel<R, kThreads, false, true>)
                     : (dbgk ? mas_path_systolic_kernel<R, kThreads, true, false> : mas_path_systolic_kernel<R, kThreads, false, false>);
    if (int rc = optin[dbgk * 2 + clus].ensure(kern, dev, plan.total)) return rc;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(p.B * plan.K));
    cfg.blockDim = dim3((unsigned)((plan.W + 1) * 32));
    cfg.dynamicSmemBytes = (size_t)plan.total;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)plan.K;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    static const char *pdl_env = getenv("MAS_B200_PDL");       // experiment hook (shared with the single launch)
    cfg.attrs = attr;
    cfg.numAttrs = (pdl_env && atoi(pdl_env) == 0) ? 1 : 2;
    MAS_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tmap, p, plan));
    return MAS_OK;
}

template <int R>
static int launch_r(const CUtensorMap &tmap, const PathParams &p, const Plan &plan, cudaStream_t stream) {
    if (plan.W <= 4) return launch_rt<R, 160>(tmap, p, plan, stream);
    return launch_rt<R, (kMaxDpWarps + 1) * 32>(tmap, p, plan, stream);
}

}  // namespace systolic

void path_systolic_force_cluster(int k) { systolic::g_force_cluster.store(k); }

// Host-only: the plan the launcher would pick on a device with `max_smem` bytes of opt-in shared
// memory per CTA and `num_sms` SMs.  out8 = {R, W, S, K, rows, nblk, bits_in_smem, total bytes}.
bool debug_path_plan(int B, int T_x, int T_y, int max_smem, int num_sms, int32_t *out8) {
    systolic::Plan pl{};
    if (!systolic::choose_plan(B, T_x, T_y, max_smem, num_sms, pl)) return false;
    out8[0] = pl.R, out8[1] = pl.W, out8[2] = pl.S, out8[3] = pl.K, out8[4] = pl.rows, out8[5] = pl.nblk;
    out8[6] = pl.bits_in_smem, out8[7] = pl.total;
    return true;
}

size_t path_systolic_workspace_bytes(int B, int T_x, int T_y) {
    // upper bound over every plan the heuristic may pick: tokens rounded up per CTA, K <= 8
    size_t worst = 0;
    for (int K = 1; K <= 8; K *= 2) {
        int R, W;
        if (!systolic::choose_shape(ceil_div(T_x, K), R, W)) continue;
        const size_t need = (size_t)B * K * ceil_div(T_y, 32) * (W * 32 * R) * 4;
        if (need > worst) worst = need;
    }
    return align_up(worst, 256);
}

// MAS_OK: launched.  MAS_ERR_UNSUPPORTED_SHAPE: this shape/alignment is not for the TMA path (the
// caller falls back to the generic kernel).  Anything else is an error.
int launch_path_systolic(PathParams p, void *workspace, size_t workspace_bytes, cudaStream_t stream) {
    using namespace systolic;
    if (p.B == 0) return MAS_OK;
    // TMA needs 16-byte aligned rows: base and both strides
    if ((p.T_y & 3) || (p.value_stride_x & 3) || (p.value_stride_b & 3) || p.value_stride_b <= 0 ||
        (reinterpret_cast<uintptr_t>(p.value) & 15) || (reinterpret_cast<uintptr_t>(p.path) & 15) || p.T_y < kBlk)
        return MAS_ERR_UNSUPPORTED_SHAPE;
    PFN_cuTensorMapEncodeTiled_v12000 encode = get_encode_fn();
    if (encode == nullptr) return MAS_ERR_UNSUPPORTED_SHAPE;

    int dev = 0;
    MAS_CUDA_TRY(cudaGetDevice(&dev));
    DeviceInfo di{};
    if (int rc = get_device_info(dev, di)) return rc;
    const int max_smem = di.max_smem_optin - 2048;   // static shared + alignment slack
    Plan plan{};
    if (!choose_plan(p.B, p.T_x, p.T_y, max_smem, di.num_sms, plan)) return MAS_ERR_UNSUPPORTED_SHAPE;
    if (!plan.bits_in_smem) {
        const size_t need = (size_t)p.B * plan.K * plan.nblk * plan.rows * 4;
        if (workspace == nullptr || workspace_bytes < need) return MAS_ERR_WORKSPACE_TOO_SMALL;
        p.ws_bits = static_cast<uint32_t *>(workspace);
    }

    CUtensorMap tmap;
    const cuuint64_t gdim[3] = {(cuuint64_t)p.T_y, (cuuint64_t)p.T_x, (cuuint64_t)p.B};
    const cuuint64_t gstride[2] = {(cuuint64_t)p.value_stride_x * 4, (cuuint64_t)p.value_stride_b * 4};
    const cuuint32_t box[3] = {(cuuint32_t)kBlk, (cuuint32_t)(kBlk * plan.R), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(p.value), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return MAS_ERR_UNSUPPORTED_SHAPE;

    switch (plan.R) {
        case 1: return launch_r<1>(tmap, p, plan, stream);
        case 2: return launch_r<2>(tmap, p, plan, stream);
        case 3: return launch_r<3>(tmap, p, plan, stream);
        case 4: return launch_r<4>(tmap, p, plan, stream);
        case 5: return launch_r<5>(tmap, p, plan, stream);
        case 6: return launch_r<6>(tmap, p, plan, stream);
        case 8: return launch_r<8>(tmap, p, plan, stream);
        default: return MAS_ERR_UNSUPPORTED_SHAPE;
    }
}

}  // namespace mas

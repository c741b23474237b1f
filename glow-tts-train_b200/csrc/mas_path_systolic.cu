// mas_path_systolic.cu -- kernel (1), the B200 fast path.
//
// One CTA per utterance.  The forward sweep of core.pyx:17-30 is laid out as a systolic array over
// TOKENS: every thread owns R consecutive tokens (rows) whose running scores live in registers,
// lane l+1 gets lane l's last score through one warp shuffle per mel frame, and warp w+1 gets warp
// w's last score through a small shared-memory ring published once per 32-frame block.  Warps are
// therefore skewed in time instead of meeting at a CTA barrier every frame; the skew is free
// because token x is outside the reference's band before frame x anyway (core.pyx:18).
//
// Scores are streamed from HBM by TMA: each warp keeps its own S-deep ring of [32R tokens x 32
// frames] boxes (128-byte rows, SWIZZLE_128B so that the per-thread 16-byte row reads are
// bank-conflict free), refilled by the warp's lane 0 as soon as a box has been consumed.
//
// Backtrack directions are packed 1 bit per cell (32 frames of one token per word) into shared
// memory, or into the caller's workspace when they do not fit.  The backtrack of core.pyx:32-35
// then walks TOKENS, not frames: for token x it finds, with one count-leading-zeros, the frame at
// which the path entered x.  A dedicated warp zero-fills the dense output while the sweep runs;
// the ones are written last.
#include <cuda.h>
#include <cudaTypedefs.h>

#include "mas_kernels.cuh"
#include "mas_ptx.cuh"

namespace mas {
namespace systolic {

constexpr int kBlk = 32;            // frames per box / per direction word
constexpr int kMaxDpWarps = 15;     // + 1 filler warp = 512 threads
constexpr int kBndBlocks = 4;       // depth of the warp-to-warp boundary ring, in 32-frame blocks
constexpr int kDoneAll = 0x3fffffff;
constexpr uint32_t kSpinLimit = 1u << 27;   // watchdog: a wedged wait traps instead of hanging the GPU

struct Plan {
    int R, W, S;
    int rows;            // W * 32 * R
    int nblk;            // ceil(T_y / 32)
    int bits_in_smem;
    // byte offsets into dynamic shared memory (base is 1024-aligned)
    int off_ring, off_bits, off_bnd, off_bar, off_done, off_run, total;
};

__host__ __device__ inline int stage_bytes(int R) { return kBlk * R * kBlk * 4; }   // 32R rows x 128 B

__host__ __device__ inline Plan make_plan(int R, int W, int S, int T_y, bool bits_in_smem) {
    Plan p;
    p.R = R;
    p.W = W;
    p.S = S;
    p.rows = W * kBlk * R;
    p.nblk = ceil_div(T_y, kBlk);
    p.bits_in_smem = bits_in_smem ? 1 : 0;
    int off = 0;
    p.off_ring = off;
    off += W * S * stage_bytes(R);
    p.off_run = p.off_ring;                       // run table aliases the ring (free after the sweep)
    p.off_bits = off;
    if (bits_in_smem) off += p.nblk * p.rows * 4;
    p.off_bnd = off;
    off += W * kBndBlocks * kBlk * 4;
    p.off_bar = off;
    off += W * S * 8;
    p.off_done = off;
    off += (W + 1) * 4;
    p.total = (int)align_up((size_t)off, 16);
    return p;
}

__device__ __forceinline__ void spin_fail() {
    __trap();
}

// One (token, frame) cell: core.pyx:19-30.  `adv` = score of token-1 at frame-1, `stay` = score of
// this token at frame-1.  Strict '>' so that a tie (or a NaN) keeps `stay`, like core.c:2697-2708.
__device__ __forceinline__ void cell(float &stay, float adv, float l, uint32_t &acc, int bit) {
    const bool take = adv > stay;
    const float best = take ? adv : stay;
    stay = best + l;                       // plain fp32 round-to-nearest add (core.pyx:30)
    acc |= take ? (1u << bit) : 0u;
}

// 32 frames of R tokens per lane.  tile: this warp's [32R][32] fp32 box (128B-swizzled).
// The loop over 4-frame groups is deliberately NOT unrolled: a DP warp runs alone on its scheduler,
// so nothing hides instruction fetch and the body has to stay inside the L0 instruction cache.
template <int R, bool kGuard>
__device__ __forceinline__ void sweep_block(const float *__restrict__ tile, float (&v)[R], uint32_t (&acc)[R],
                                            float &carry, const float4 *__restrict__ bnd_in, float4 *bnd_out,
                                            bool first_block_of_warp0, int lane, int row0, int col0, float neg) {
    int swz[R];                                     // per-row XOR term of the 128B swizzle
    const float *rowp[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
        const int q = lane * R + i;
        rowp[i] = tile + q * kBlk;
        swz[i] = q & 7;
    }
#pragma unroll 1
    for (int g = 0; g < 8; ++g) {
        float4 L[R];
#pragma unroll
        for (int i = 0; i < R; ++i) L[i] = *reinterpret_cast<const float4 *>(rowp[i] + ((g ^ swz[i]) << 2));
        // scores of the previous warp's last token after frames col0+4g-1 .. col0+4g+2
        float up4[4];
        if (bnd_in != nullptr) {
            const float4 b = bnd_in[g];
            up4[0] = carry;
            up4[1] = b.x;
            up4[2] = b.y;
            up4[3] = b.z;
            carry = b.w;
        } else {
            // token 0: "advance" comes from outside the lattice: 0 at frame 0, -1e9 after (core.pyx:23-27)
            up4[0] = (first_block_of_warp0 && g == 0) ? 0.f : neg;
            up4[1] = up4[2] = up4[3] = neg;
        }
        float out4[4];
        uint32_t acc4[R];
#pragma unroll
        for (int i = 0; i < R; ++i) acc4[i] = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float up = __shfl_up_sync(0xffffffffu, v[R - 1], 1);
            if (lane == 0) up = up4[j];
#pragma unroll
            for (int i = R - 1; i >= 0; --i) {
                float l = (j == 0) ? L[i].x : (j == 1) ? L[i].y : (j == 2) ? L[i].z : L[i].w;
                if (kGuard) {
                    // below the diagonal (token > frame) the reference never computes the cell and
                    // reads -1e9 instead (core.pyx:19-20): adding 0 keeps the score at exactly -1e9
                    if (row0 + i > col0 + 4 * g + j) l = 0.f;
                }
                cell(v[i], (i == 0) ? up : v[i - 1], l, acc4[i], j);
            }
            out4[j] = v[R - 1];
        }
#pragma unroll
        for (int i = 0; i < R; ++i) acc[i] |= acc4[i] << (4 * g);
        if (bnd_out != nullptr && lane == 31) bnd_out[g] = make_float4(out4[0], out4[1], out4[2], out4[3]);
    }
}

template <int R>
__global__ void __launch_bounds__((kMaxDpWarps + 1) * 32, 1)
mas_path_systolic_kernel(const __grid_constant__ CUtensorMap tmap, PathParams p, Plan plan) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ float s_len[2];

    const int b = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int W = plan.W, S = plan.S;
    const int T_x = p.T_x, T_y = p.T_y;

    float *ring = reinterpret_cast<float *>(smem + plan.off_ring);
    // packed directions [nblk][rows]: shared memory when they fit, else the caller's workspace
    uint32_t *bits_s = reinterpret_cast<uint32_t *>(smem + plan.off_bits);
    uint32_t *bits_g = plan.bits_in_smem ? nullptr : p.ws_bits + (size_t)b * plan.nblk * plan.rows;
    float *bnd = reinterpret_cast<float *>(smem + plan.off_bnd);                          // [W][kBndBlocks*32]
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + plan.off_bar);                   // [W][S]
    int *done = reinterpret_cast<int *>(smem + plan.off_done);                            // [W+1]
    int2 *run = reinterpret_cast<int2 *>(smem + plan.off_run);                            // [T_x] after the sweep

    // ---- lengths (monotonic_align/__init__.py:18-19 when they come from the mask) ----
    int tx_raw, ty_raw;
    if (p.t_x != nullptr) {
        tx_raw = p.t_x[b];
        ty_raw = p.t_y[b];
    } else {
        if (tid < 2) s_len[tid] = 0.f;
        __syncthreads();
        float sx = 0.f, sy = 0.f;
        const float *m = p.mask + (int64_t)b * p.mask_stride_b;
        for (int x = tid; x < T_x; x += blockDim.x) sx += m[(int64_t)x * p.mask_stride_x];
        for (int y = tid; y < T_y; y += blockDim.x) sy += m[(int64_t)y * p.mask_stride_y];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sx += __shfl_xor_sync(0xffffffffu, sx, o);
            sy += __shfl_xor_sync(0xffffffffu, sy, o);
        }
        if (lane == 0) {
            atomicAdd(&s_len[0], sx);
            atomicAdd(&s_len[1], sy);
        }
        __syncthreads();
        tx_raw = (int)s_len[0];
        ty_raw = (int)s_len[1];
    }
    const Lengths len = clamp_lengths(tx_raw, ty_raw, T_x, T_y);
    const int tx = len.tx, ty = len.ty;

    // ---- per-warp geometry: tokens [x0, x0+32R), 32-frame blocks [cb0, cbend] ----
    const int rows_per_warp = kBlk * R;
    const int x0 = warp * rows_per_warp;
    const bool dp_warp = warp < W;
    const bool active = dp_warp && x0 < tx;
    int cb0 = 0, cbend = -1;
    if (active) {
        const int x1 = min(x0 + rows_per_warp, tx) - 1;
        cb0 = x0 >> 5;                                       // token x0 enters the band at frame x0
        cbend = min(ty - 1, x1 + (ty - tx)) >> 5;            // ... and token x1 leaves it here (core.pyx:18)
    }
    if (tid == 0) {
        for (int i = 0; i < W * S; ++i) ptx::mbar_init(&full[i], 1);
        ptx::fence_barrier_init();
        ptx::fence_proxy_async();
        done[W] = kDoneAll;                                  // nobody consumes the last warp's boundary
    }
    // done[w] = number of 32-frame blocks warp w has finished.  It starts one short of the warp's
    // first block: the warp still needs the LAST frame of block cb0-1 from its predecessor (the
    // diagonal cell of token x0-1), so that ring slot must not be recycled yet.
    if (dp_warp && lane == 0) done[warp] = active ? cb0 - 1 : kDoneAll;
    __syncthreads();

    long long *dbg = p.dbg_cycles ? p.dbg_cycles + ((size_t)b * 16 + warp) * 16 : nullptr;
    long long t_wait_prev = 0, t_wait_next = 0, t_wait_tma = 0, t_sweep = 0, t_tail = 0;
    if (dbg && lane == 0) dbg[0] = clock64();
    if (dp_warp) {
        if (active) {
            float *my_ring = ring + (size_t)warp * S * (rows_per_warp * kBlk);
            uint64_t *my_full = full + warp * S;
            const uint32_t box_bytes = stage_bytes(R);
            if (lane == 0) {
                ptx::prefetch_tensormap(&tmap);
                for (int k = 0; k < S && cb0 + k <= cbend; ++k) {
                    ptx::mbar_arrive_expect_tx(&my_full[k], box_bytes);
                    ptx::tma_load_3d(my_ring + (size_t)k * rows_per_warp * kBlk, &tmap, &my_full[k], (cb0 + k) * kBlk, x0, b);
                }
            }
            float v[R];
            uint32_t acc[R];
#pragma unroll
            for (int i = 0; i < R; ++i) v[i] = p.max_neg_val;
            float carry = p.max_neg_val;
            const float *bnd_in_base = (warp > 0) ? bnd + (size_t)(warp - 1) * kBndBlocks * kBlk : nullptr;
            float *bnd_out_base = (warp + 1 < W) ? bnd + (size_t)warp * kBndBlocks * kBlk : nullptr;
            const int row0 = x0 + lane * R;

            for (int cb = cb0; cb <= cbend; ++cb) {
                const int k = cb - cb0, slot = k % S;
                const uint32_t parity = (k / S) & 1;
                uint32_t spins = 0;
                long long t0 = dbg ? clock64() : 0;
                if (warp > 0)                                   // previous warp has published block cb
                    while (ptx::ld_acquire_shared(&done[warp - 1]) <= cb)
                        if (++spins > kSpinLimit) spin_fail();
                long long t1 = dbg ? clock64() : 0;
                if (bnd_out_base != nullptr)                    // next warp has consumed block cb - ring depth
                    while (ptx::ld_acquire_shared(&done[warp + 1]) + kBndBlocks <= cb)
                        if (++spins > kSpinLimit) spin_fail();
                long long t2 = dbg ? clock64() : 0;
                while (!ptx::mbar_try_wait(&my_full[slot], parity))
                    if (++spins > kSpinLimit) spin_fail();
                if (dbg) {
                    const long long t3 = clock64();
                    t_wait_prev += t1 - t0;
                    t_wait_next += t2 - t1;
                    t_wait_tma += t3 - t2;
                }
                if (cb == cb0 && warp > 0)                      // score of token x0-1 on the diagonal frame x0-1
                    carry = bnd_in_base[((cb0 - 1) % kBndBlocks) * kBlk + (kBlk - 1)];

                const long long t4 = dbg ? clock64() : 0;
#pragma unroll
                for (int i = 0; i < R; ++i) acc[i] = 0u;
                const float *tile = my_ring + (size_t)slot * rows_per_warp * kBlk;
                const float4 *bin = bnd_in_base ? reinterpret_cast<const float4 *>(bnd_in_base + (cb % kBndBlocks) * kBlk) : nullptr;
                float4 *bout = bnd_out_base ? reinterpret_cast<float4 *>(bnd_out_base + (cb % kBndBlocks) * kBlk) : nullptr;
                if (cb < cb0 + R)
                    sweep_block<R, true>(tile, v, acc, carry, bin, bout, warp == 0 && cb == 0, lane, row0, cb * kBlk, p.max_neg_val);
                else
                    sweep_block<R, false>(tile, v, acc, carry, bin, bout, false, lane, row0, cb * kBlk, p.max_neg_val);
                const long long t5 = dbg ? clock64() : 0;
                if (plan.bits_in_smem) {
#pragma unroll
                    for (int i = 0; i < R; ++i) bits_s[cb * plan.rows + row0 + i] = acc[i];
                } else {
#pragma unroll
                    for (int i = 0; i < R; ++i) bits_g[(size_t)cb * plan.rows + row0 + i] = acc[i];
                }
                __syncwarp();
                if (lane == 0) {
                    if (cb + S <= cbend) {
                        ptx::mbar_arrive_expect_tx(&my_full[slot], box_bytes);
                        ptx::tma_load_3d(my_ring + (size_t)slot * rows_per_warp * kBlk, &tmap, &my_full[slot], (cb + S) * kBlk, x0, b);
                    }
                    ptx::st_release_shared(&done[warp], cb + 1);
                }
                if (dbg) {
                    t_sweep += t5 - t4;
                    t_tail += clock64() - t5;
                }
            }
            __syncwarp();
            if (lane == 0) ptx::st_release_shared(&done[warp], kDoneAll);
        }
        if (dbg && lane == 0) {
            dbg[1] = clock64();
            dbg[2] = t_wait_prev;
            dbg[3] = t_wait_next;
            dbg[4] = t_wait_tma;
            dbg[8] = t_sweep;
            dbg[9] = t_tail;
            dbg[10] = cbend - cb0 + 1;
        }
    } else {
        // ---- filler warp: zero the dense output while the sweep runs ----
        float4 *o4 = reinterpret_cast<float4 *>(p.path + (int64_t)b * T_x * T_y);
        const int64_t n4 = ((int64_t)T_x * T_y) >> 2;            // T_y % 4 == 0 on this path
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        int64_t i = lane;
        for (; i + 96 < n4; i += 128) {
            ptx::st_global_cs_v4(o4 + i, z4);
            ptx::st_global_cs_v4(o4 + i + 32, z4);
            ptx::st_global_cs_v4(o4 + i + 64, z4);
            ptx::st_global_cs_v4(o4 + i + 96, z4);
        }
        for (; i < n4; i += 32) ptx::st_global_cs_v4(o4 + i, z4);
        if (dbg && lane == 0) dbg[1] = clock64();
    }
    if (!plan.bits_in_smem) __threadfence_block();
    __syncthreads();

    // ---- backtrack (core.pyx:32-35) by TOKENS, warp 0 in lockstep ----
    // State: the path sits on token x for frames (.., y_hi]; ys <= y_hi is how far left that run has
    // been scanned without finding the frame where the path stepped onto x.  Lane k holds the
    // direction word of token x-k for the current 32-frame block, so a block is fetched with one
    // load and every further token costs a shuffle + count-leading-zeros instead of a memory trip.
    // Stepping is forced on the diagonal (frame == token, core.pyx:34 `index == y`).
    if (dbg && tid == 0) dbg[5] = clock64();
    if (warp == 0 && tx > 0) {
        int x = tx - 1, y_hi = ty - 1, ys = ty - 1;
        while (x > 0) {
            const int cb = ys >> 5;
            const int r = x - lane;
            uint32_t wd = 0u;
            if (r > 0) wd = plan.bits_in_smem ? bits_s[cb * plan.rows + r] : __ldcg(bits_g + (size_t)cb * plan.rows + r);
#pragma unroll 4
            for (int k = 0; k < 32; ++k) {
                const uint32_t w = __shfl_sync(0xffffffffu, wd, k);
                const uint32_t m = w & (0xffffffffu >> (31 - (ys & 31)));
                int ylo = (m != 0u) ? (cb << 5) + 31 - __clz(m) : -1;
                ylo = max(ylo, x);
                if (ylo < (cb << 5)) {          // stepped onto x in an earlier block: same token, next block
                    ys = (cb << 5) - 1;
                    break;
                }
                if (lane == 0) run[x] = make_int2(ylo, y_hi);
                y_hi = ys = ylo - 1;
                --x;
                if (x == 0 || (ys >> 5) != cb) break;
            }
        }
        if (lane == 0) run[0] = make_int2(0, y_hi);
    }
    if (dbg && tid == 0) dbg[6] = clock64();
    __syncthreads();

    // ---- dense path: ones, durations, frame -> token ----
    float *out = p.path + (int64_t)b * T_x * T_y;
    for (int x = tid; x < T_x; x += blockDim.x) {
        int d = 0;
        if (x < tx) {
            const int2 r = run[x];
            d = r.y - r.x + 1;
            float *row = out + (int64_t)x * T_y;
            for (int y = r.x; y <= r.y; ++y) row[y] = 1.f;
            if (p.frame_token)
                for (int y = r.x; y <= r.y; ++y) p.frame_token[(int64_t)b * T_y + y] = x;
        }
        if (p.durations) p.durations[(int64_t)b * T_x + x] = d;
    }
    if (p.frame_token)
        for (int y = ty + tid; y < T_y; y += blockDim.x) p.frame_token[(int64_t)b * T_y + y] = -1;
    if (dbg && lane == 0) dbg[7] = clock64();
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
    }
    return fn;
}

static bool choose_shape(int T_x, int &R, int &W) {
    const int groups = ceil_div(T_x, kBlk);            // 32-token groups
    if (groups <= 1) {
        R = 1;
        W = 1;
        return true;
    }
    for (int r : {3, 5}) {
        const int w = ceil_div(groups, r);
        if (w <= kMaxDpWarps) {
            R = r;
            W = w;
            return true;
        }
    }
    return false;
}

template <int R>
static int launch_r(const CUtensorMap &tmap, const PathParams &p, const Plan &plan, cudaStream_t stream) {
    MAS_CUDA_TRY(cudaFuncSetAttribute(mas_path_systolic_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.total));
    mas_path_systolic_kernel<R><<<p.B, (plan.W + 1) * 32, plan.total, stream>>>(tmap, p, plan);
    MAS_CUDA_TRY(cudaGetLastError());
    return MAS_OK;
}

}  // namespace systolic

size_t path_systolic_workspace_bytes(int B, int T_x, int T_y) {
    int R, W;
    if (!systolic::choose_shape(T_x, R, W)) return 0;
    return align_up((size_t)B * ceil_div(T_y, 32) * (W * 32 * R) * 4, 256);
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
    int R, W;
    if (!choose_shape(p.T_x, R, W)) return MAS_ERR_UNSUPPORTED_SHAPE;
    PFN_cuTensorMapEncodeTiled_v12000 encode = get_encode_fn();
    if (encode == nullptr) return MAS_ERR_UNSUPPORTED_SHAPE;

    int dev = 0, max_smem = 0;
    MAS_CUDA_TRY(cudaGetDevice(&dev));
    MAS_CUDA_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    max_smem -= 2048;   // static shared + alignment slack
    // deepest ring that fits, bits in shared memory if possible
    Plan plan{};
    bool ok = false;
    for (int bits_smem = 1; bits_smem >= 0 && !ok; --bits_smem)
        for (int S = 4; S >= 2; --S) {
            plan = make_plan(R, W, S, p.T_y, bits_smem != 0);
            if (plan.total <= max_smem && (bits_smem || S >= 3 || true)) {
                ok = true;
                break;
            }
        }
    if (!ok) return MAS_ERR_UNSUPPORTED_SHAPE;
    if (!plan.bits_in_smem) {
        const size_t need = path_systolic_workspace_bytes(p.B, p.T_x, p.T_y);
        if (workspace == nullptr || workspace_bytes < need) return MAS_ERR_WORKSPACE_TOO_SMALL;
        p.ws_bits = static_cast<uint32_t *>(workspace);
    }

    CUtensorMap tmap;
    const cuuint64_t gdim[3] = {(cuuint64_t)p.T_y, (cuuint64_t)p.T_x, (cuuint64_t)p.B};
    const cuuint64_t gstride[2] = {(cuuint64_t)p.value_stride_x * 4, (cuuint64_t)p.value_stride_b * 4};
    const cuuint32_t box[3] = {(cuuint32_t)kBlk, (cuuint32_t)(kBlk * R), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(p.value), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return MAS_ERR_UNSUPPORTED_SHAPE;

    switch (R) {
        case 1: return launch_r<1>(tmap, p, plan, stream);
        case 3: return launch_r<3>(tmap, p, plan, stream);
        case 5: return launch_r<5>(tmap, p, plan, stream);
        default: return MAS_ERR_UNSUPPORTED_SHAPE;
    }
}

}  // namespace mas

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
#include <cudaTypedefs.h>

#include "mas_kernels.cuh"
#include "mas_ptx.cuh"

namespace mas {
namespace systolic {

constexpr int kBlk = 32;            // frames per box / per direction word
constexpr int kMaxDpWarps = 15;     // + 1 filler warp = 512 threads
constexpr int kBndBlocks = 4;       // depth of the warp-to-warp boundary ring, in 32-frame blocks
constexpr int kDoneAll = 0x3fffffff;
constexpr int kZeroBytes = 16384;  // shared zero page the filler streams to the dense output with bulk copies
constexpr uint32_t kSpinLimit = 1u << 27;   // watchdog: a wedged wait traps instead of hanging the GPU

struct Plan {
    int R, W, S, K;      // tokens per lane, sweep warps per CTA, TMA ring depth, CTAs per utterance
    int rows;            // tokens per CTA: W * 32 * R
    int nblk;            // ceil(T_y / 32)
    int bits_in_smem;
    // byte offsets into dynamic shared memory (base is 1024-aligned)
    int off_ring, off_bits, off_bnd, off_bar, off_done, off_misc, off_run, off_zero, total;
};

__host__ __device__ inline int stage_bytes(int R) { return kBlk * R * kBlk * 4; }   // 32R rows x 128 B

__host__ __device__ inline Plan make_plan(int R, int W, int S, int K, int T_y, bool bits_in_smem) {
    Plan p;
    p.R = R;
    p.W = W;
    p.S = S;
    p.K = K;
    p.rows = W * kBlk * R;
    p.nblk = ceil_div(T_y, kBlk);
    p.bits_in_smem = bits_in_smem ? 1 : 0;
    int off = 0;
    p.off_ring = off;
    off += W * S * stage_bytes(R);
    p.off_run = p.off_ring;                       // run table / exact-sweep columns alias the ring (free after the sweep)
    p.off_bits = off;
    if (bits_in_smem) off += p.nblk * p.rows * 4;
    p.off_zero = off;
    off += kZeroBytes;
    p.off_bnd = off;
    off += (W + 1) * kBndBlocks * kBlk * 4;       // ring w = boundary INTO warp w; ring 0 is constant -1e9
    p.off_bar = off;
    off += W * S * 8;
    p.off_done = off;
    off += (W + 2) * 4;                           // [prev CTA's last warp | own warps | next CTA's first warp]
    p.off_misc = off;
    off += 8 * 4;                                 // backtrack hand-over (flag, token, frame), redo flag
    p.total = (int)align_up((size_t)off, 16);
    return p;
}

__device__ __forceinline__ void spin_fail() { __trap(); }

__device__ __forceinline__ float fmax_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

// One (token, frame) cell of core.pyx:19-30.  `adv` = score of token-1 at frame-1, `stay` = score of
// this token at frame-1.  The direction bit (advance strictly better, core.c:2697-2708) is the sign
// of stay - adv; it is shifted into `acc` from the right, so after 32 frames bit 31 is the block's
// first frame (the caller bit-reverses the word).
__device__ __forceinline__ void cell_fast(float &stay, float adv, float l, uint32_t &acc) {
    const float diff = stay - adv;
    const float best = fmax_nan(adv, stay);
    stay = best + l;                                       // plain fp32 round-to-nearest add (core.pyx:30)
    acc = __funnelshift_l(__float_as_uint(diff), acc, 1);
}

// Four frames (one 16-byte group) of R tokens per lane.
template <int R, bool kCluster>
__device__ __forceinline__ void sweep_group(const float4 (&L)[R], const float4 &b, float (&v)[R], uint32_t (&acc)[R],
                                            float &carry, uint32_t bnd_out, bool publisher, int g) {
    // scores of the previous warp's last token after frames col0+4g-1 .. col0+4g+2
    const float up4[4] = {carry, b.x, b.y, b.z};
    carry = b.w;
    float out4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float up = __shfl_up_sync(0xffffffffu, v[R - 1], 1);
        if ((threadIdx.x & 31) == 0) up = up4[j];
#pragma unroll
        for (int i = R - 1; i >= 0; --i) {
            const float l = (j == 0) ? L[i].x : (j == 1) ? L[i].y : (j == 2) ? L[i].z : L[i].w;
            cell_fast(v[i], (i == 0) ? up : v[i - 1], l, acc[i]);
        }
        out4[j] = v[R - 1];
    }
    // lane 31 hands its last token's four scores to the next warp (possibly in the next CTA: the
    // address is a shared::cluster one)
    if (kCluster)
        ptx::st_cluster_v4_if(publisher, bnd_out + g * 16, make_float4(out4[0], out4[1], out4[2], out4[3]));
    else
        ptx::st_shared_v4_if(publisher, bnd_out + g * 16, make_float4(out4[0], out4[1], out4[2], out4[3]));
}

// Below the diagonal (token > frame) the reference never computes a cell and reads -1e9 in its place
// (core.pyx:18-20).  Zeroing those scores in the staged box makes the sweep reproduce that without
// a per-cell test: max(-1e9, -1e9) + 0 stays exactly -1e9.  Only the first R blocks of a warp touch
// the diagonal.  Each lane edits its own rows (in place, swizzled 16-byte chunks).
template <int R>
__device__ __forceinline__ void zero_below_diagonal(float *tile, int lane, int row0, int col0) {
#pragma unroll
    for (int i = 0; i < R; ++i) {
        const int q = lane * R + i;
        const int d = row0 + i - col0;              // frames [0, d) of this block are below the diagonal
        if (d <= 0) continue;
        float *rowp = tile + q * kBlk;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (4 * c >= d) break;
            float4 *p4 = reinterpret_cast<float4 *>(rowp + ((c ^ (q & 7)) << 2));
            float4 x = *p4;
            x.x = 0.f;
            if (4 * c + 1 < d) x.y = 0.f;
            if (4 * c + 2 < d) x.z = 0.f;
            if (4 * c + 3 < d) x.w = 0.f;
            *p4 = x;
        }
    }
}

// 32 frames of R tokens per lane.  tile: this warp's [32R][32] fp32 box (128B-swizzled); bnd_in: the
// previous warp's last-token scores for these 32 frames (for warp 0 a constant -1e9 ring: token 0
// can only be "advanced into" from outside the lattice, core.pyx:23-27).
// The loop is software-pipelined by hand (group g+1 is fetched from shared memory while group g is
// swept) and only unrolled twice: a DP warp runs alone on its scheduler, so nothing else hides a
// shared-memory round trip or an instruction-cache miss.
template <int R, bool kCluster>
__device__ __forceinline__ void sweep_block(const float *__restrict__ tile, float (&v)[R], uint32_t (&acc)[R],
                                            float &carry, const float4 *__restrict__ bnd_in, uint32_t bnd_out,
                                            bool publisher, int lane) {
    int swz[R];                                     // per-row XOR term of the 128B swizzle
    const float *rowp[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
        const int q = lane * R + i;
        rowp[i] = tile + q * kBlk;
        swz[i] = q & 7;
    }
    float4 LA[R], LB[R], bA, bB;
#pragma unroll
    for (int i = 0; i < R; ++i) LA[i] = *reinterpret_cast<const float4 *>(rowp[i] + (swz[i] << 2));
    bA = bnd_in[0];
#pragma unroll 1
    for (int g = 0; g < 8; g += 2) {
#pragma unroll
        for (int i = 0; i < R; ++i) LB[i] = *reinterpret_cast<const float4 *>(rowp[i] + (((g + 1) ^ swz[i]) << 2));
        bB = bnd_in[g + 1];
        sweep_group<R, kCluster>(LA, bA, v, acc, carry, bnd_out, publisher, g);
        const int gn = (g + 2) & 7;                 // the last prefetch wraps to group 0 and is discarded
#pragma unroll
        for (int i = 0; i < R; ++i) LA[i] = *reinterpret_cast<const float4 *>(rowp[i] + ((gn ^ swz[i]) << 2));
        bA = bnd_in[gn];
        sweep_group<R, kCluster>(LB, bB, v, acc, carry, bnd_out, publisher, g + 1);
    }
}

// Exact compare/select sweep for utterances whose scores are not all finite (NaN / +-inf): CTA 0
// of the cluster walks the frames for ALL tokens with a barrier per frame, score column in shared
// memory, scores read straight from global memory.  Slow, rare, and literal: core.pyx:17-30 as
// written.  col: [2][tx] floats.  The direction words go where the fast sweep would have put
// them: word (cb, x) belongs to CTA x / rows, in its shared memory (bits_smem_addr, written over
// DSMEM) or in the workspace (bits_g, [K][nblk][rows]).
__device__ void exact_sweep_cta0(const float *__restrict__ val, int64_t stride_x, float *col, uint32_t bits_smem_addr,
                                 uint32_t *bits_g, int rows, int nblk, int tx, int ty, float neg) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int x = tid; x < tx; x += nthr) col[x] = neg;
    __syncthreads();
    int buf = 0;
    for (int y = 0; y < ty; ++y) {
        const float *vin = col + buf * tx;
        float *vout = col + (buf ^ 1) * tx;
        for (int x = tid; x < tx; x += nthr) {
            const float stay = vin[x];                                      // == -1e9 while x > y-1, core.pyx:19-20
            const float adv = (x == 0) ? ((y == 0) ? 0.f : neg) : vin[x - 1];   // core.pyx:23-29
            const float l = (x > y) ? 0.f : __ldg(val + (int64_t)x * stride_x + y);
            const bool take = adv > stay;
            vout[x] = (take ? adv : stay) + l;
            const uint32_t bit = ((take || (x == y && x > 0)) ? 1u : 0u) << (y & 31);
            const int owner = x / rows, xl = x - owner * rows;
            const size_t word = (size_t)(y >> 5) * rows + xl;
            if (bits_g == nullptr) {
                const uint32_t addr = ptx::mapa(bits_smem_addr + (uint32_t)word * 4u, owner);
                ptx::st_cluster_u32(addr, ((y & 31) ? ptx::ld_cluster_u32(addr) : 0u) | bit);
            } else {
                uint32_t *w = bits_g + (size_t)owner * nblk * rows + word;
                __stcg(w, ((y & 31) ? __ldcg(w) : 0u) | bit);
            }
        }
        buf ^= 1;
        __syncthreads();
    }
}

// Backtrack of core.pyx:32-35, walking TOKENS instead of frames.  The path sits on token x for
// frames (.., y_hi]; the frame where it stepped onto x is the highest set direction bit at or
// below the scan position -- one count-leading-zeros per token.  The forced step on the diagonal
// (frame == token, core.pyx:34 `index == y`) was OR-ed into the words by the sweep.
// bits: [nblk][rows] words, bit j of word (cb, x) = direction of cell (x, 32 cb + j).
// This CTA's part: tokens x .. x_min (global numbering; `bits` and `run` are indexed by the local
// token x - xc).  Returns the last frame of token x_min - 1 (where the next CTA down continues).
template <bool kSmem>
__device__ __forceinline__ int backtrack_tokens(const uint32_t *bits, int rows, int xc, int x, int y_hi, int x_min,
                                                int2 *run) {
    int base = y_hi & ~31;
    uint32_t elig = 0xffffffffu >> (31 - (y_hi & 31));     // bits at or below the scan position
    const uint32_t *p = bits + (size_t)(y_hi >> 5) * rows + (x - xc);
    int2 *r = run + (x - xc);
    // Branches are what a lone thread pays for (~25 cycles each): one per token, two per block.
    while (x >= x_min) {
        uint32_t m = (kSmem ? *p : __ldcg(p)) & elig;
        while (m != 0u) {                                   // the path stepped onto x inside this block
            const uint32_t wn = (x > x_min) ? (kSmem ? p[-1] : __ldcg(p - 1)) : 0u;   // next token, same block
            const int lo = 31 - __clz(m);
            *r = make_int2(base + lo, y_hi);
            y_hi = base + lo - 1;
            --x;
            --p;
            --r;
            elig = (1u << lo) - 1u;                         // lo == 0: nothing left here, leave the block
            m = wn & elig;                                  // x < x_min: wn == 0 ends both loops
        }
        base -= 32;                                         // same token, previous block
        p -= rows;
        elig = 0xffffffffu;
    }
    return y_hi;
}

// kThreads: launch bound.  Up to 4 sweep warps (+ the filler) run as 160 threads so that the
// compiler may keep loop-invariants in registers.
// kCluster: compiled with the distributed-shared-memory paths (K > 1); the K == 1 build carries none.
template <int R, int kThreads, bool kDbg, bool kCluster>
__global__ void __launch_bounds__(kThreads, 1)
mas_path_systolic_kernel(const __grid_constant__ CUtensorMap tmap, PathParams p, Plan plan) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ float s_len[2];

    const int K = kCluster ? plan.K : 1;
    const int b = blockIdx.x / K;
    const int c = kCluster ? (int)ptx::cluster_ctarank() : 0;    // which slice of the utterance's tokens
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int W = plan.W, S = plan.S, rows = plan.rows;
    const int T_x = p.T_x, T_y = p.T_y;
    const bool bits_smem = plan.bits_in_smem != 0;
    const int xc = c * rows;                                     // first token of this CTA

    float *ring = reinterpret_cast<float *>(smem + plan.off_ring);
    // packed directions [nblk][rows] of this CTA's tokens: shared memory when they fit, else workspace
    uint32_t *bits_s = reinterpret_cast<uint32_t *>(smem + plan.off_bits);
    uint32_t *bits_g = bits_smem ? nullptr : p.ws_bits + ((size_t)b * K + c) * plan.nblk * rows;
    float *bnd = reinterpret_cast<float *>(smem + plan.off_bnd);       // [W+1][kBndBlocks*32]; ring w = INTO warp w
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + plan.off_bar);   // [W][S]
    // progress counters: [0] = last warp of the previous CTA (written over DSMEM), [1+w] = own warp w,
    // [W+1] = first warp of the next CTA (written over DSMEM)
    int *done = reinterpret_cast<int *>(smem + plan.off_done);
    volatile int *misc = reinterpret_cast<volatile int *>(smem + plan.off_misc);   // [0] hand-over flag [1] token [2] frame [3] redo
    int2 *run = reinterpret_cast<int2 *>(smem + plan.off_run);         // [rows] after the sweep

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
    constexpr int rows_per_warp = kBlk * R;
    const int x0 = xc + warp * rows_per_warp;
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
        done[0] = (c == 0) ? kDoneAll : -1;                  // CTA 0 has no predecessor (its ring 0 is constant)
        if (c == K - 1) done[W + 1] = kDoneAll;              // nobody consumes the last CTA's last boundary
        misc[0] = 0;
        misc[3] = 0;
        if (kCluster && c > 0) {
            // tell the previous CTA where its consumer (my warp 0) starts
            const int first = (xc < tx) ? (xc >> 5) - 1 : kDoneAll;
            ptx::st_cluster_u32(ptx::mapa(ptx::smem_u32(&done[W + 1]), (uint32_t)(c - 1)), (uint32_t)first);
        }
    }
    // what "advances" into token 0 after frame 0 (core.pyx:26-27)
    if (c == 0)
        for (int i = tid; i < kBndBlocks * kBlk; i += blockDim.x) bnd[i] = p.max_neg_val;
    // done[1+w] = number of 32-frame blocks warp w has finished.  It starts one short of the warp's
    // first block: the warp still needs the LAST frame of block cb0-1 from its predecessor (the
    // diagonal cell of token x0-1), so that ring slot must not be recycled yet.
    if (dp_warp && lane == 0) done[1 + warp] = active ? cb0 - 1 : kDoneAll;
    if (kCluster)
        ptx::cluster_sync();
    else
        __syncthreads();

    long long *dbg = (kDbg && p.dbg_cycles) ? p.dbg_cycles + ((size_t)blockIdx.x * 16 + warp) * 16 : nullptr;
    long long t_wait_prev = 0, t_wait_tma = 0, t_sweep = 0;
    if (kDbg && dbg && lane == 0) dbg[0] = clock64();
    int nonfinite = 0;
    if (dp_warp) {
        if (active) {
            float *my_ring = ring + (size_t)warp * S * (rows_per_warp * kBlk);
            uint64_t *my_full = full + warp * S;
            constexpr uint32_t box_bytes = kBlk * R * kBlk * 4;
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
            float carry = (x0 == 0) ? 0.f : p.max_neg_val;     // frame 0 of token 0 starts from 0 (core.pyx:24-25)
            const float *bnd_in_base = bnd + (size_t)warp * kBndBlocks * kBlk;
            // where this warp's last-token scores go: the next warp's ring, or ring 0 of the next CTA
            const bool last_warp = warp == W - 1;
            const bool has_next = !last_warp || c + 1 < K;
            uint32_t bnd_out_base = 0;
            if (has_next) {
                if (kCluster)
                    bnd_out_base = last_warp ? ptx::mapa(ptx::smem_u32(bnd), c + 1)
                                             : ptx::mapa(ptx::smem_u32(bnd + (size_t)(warp + 1) * kBndBlocks * kBlk), c);
                else
                    bnd_out_base = ptx::smem_u32(bnd + (size_t)(warp + 1) * kBndBlocks * kBlk);
            }
            const bool publisher = has_next && lane == 31;
            // progress is mirrored into the neighbour CTA when the neighbour warp lives there
            const uint32_t mirror_prev = (kCluster && warp == 0 && c > 0) ? ptx::mapa(ptx::smem_u32(&done[W + 1]), (uint32_t)(c - 1)) : 0u;
            const uint32_t mirror_next = (kCluster && last_warp && c + 1 < K) ? ptx::mapa(ptx::smem_u32(&done[0]), c + 1) : 0u;
            const bool prev_remote = kCluster && warp == 0 && c > 0, next_remote = kCluster && last_warp && c + 1 < K;
            const int row0 = x0 + lane * R;
            int slot = 0;
            uint32_t parity = 0;
            // cached progress of the neighbours: shared memory is only polled when the cached value
            // does not already answer the question
            int seen_prev = (x0 > 0) ? -1 : kDoneAll;
            int seen_next = has_next ? -1 : kDoneAll;
            const bool lane0 = lane == 0, lane31 = lane == 31;

            for (int cb = cb0; cb <= cbend; ++cb) {
                uint32_t spins = 0;
                const long long t0 = kDbg ? clock64() : 0;
                while (seen_prev <= cb) {                       // previous warp has published block cb
                    seen_prev = prev_remote ? ptx::ld_acquire_cluster_shared(&done[warp]) : ptx::ld_acquire_shared(&done[warp]);
                    if (++spins > kSpinLimit) spin_fail();
                }
                while (seen_next + kBndBlocks <= cb) {          // next warp has consumed block cb - ring depth
                    seen_next = next_remote ? ptx::ld_acquire_cluster_shared(&done[warp + 2]) : ptx::ld_acquire_shared(&done[warp + 2]);
                    if (++spins > kSpinLimit) spin_fail();
                }
                const long long t2 = kDbg ? clock64() : 0;
                while (!ptx::mbar_try_wait(&my_full[slot], parity))
                    if (++spins > kSpinLimit) spin_fail();
                const long long t3 = kDbg ? clock64() : 0;
                if (cb == cb0 && x0 > 0)                        // score of token x0-1 on the diagonal frame x0-1
                    carry = bnd_in_base[((cb0 - 1) & (kBndBlocks - 1)) * kBlk + (kBlk - 1)];

#pragma unroll
                for (int i = 0; i < R; ++i) acc[i] = 0u;
                float *tile = my_ring + (size_t)slot * rows_per_warp * kBlk;
                const float4 *bin = reinterpret_cast<const float4 *>(bnd_in_base + (cb & (kBndBlocks - 1)) * kBlk);
                const uint32_t bout = bnd_out_base + (cb & (kBndBlocks - 1)) * kBlk * 4;
                const int col0 = cb * kBlk;
                const bool on_diagonal = cb < cb0 + R;        // warp-uniform
                if (on_diagonal) {
                    zero_below_diagonal<R>(tile, lane, row0, col0);
                    ptx::fence_proxy_async();               // these generic writes precede the TMA refill of the slot
                    __syncwarp();
                }
                sweep_block<R, kCluster>(tile, v, acc, carry, bin, bout, publisher, lane);
#pragma unroll
                for (int i = 0; i < R; ++i) acc[i] = __brev(acc[i]);   // first frame came in first: bit 31 -> bit 0
                if (on_diagonal) {
#pragma unroll
                    for (int i = 0; i < R; ++i) {
                        // the forced step on the diagonal (frame == token, core.pyx:34), tokens > 0 only
                        const int d = row0 + i - col0;
                        if (d >= 0 && d < kBlk && row0 + i > 0) acc[i] |= 1u << d;
                    }
                }
                if (bits_smem) {
#pragma unroll
                    for (int i = 0; i < R; ++i) bits_s[cb * rows + (row0 - xc) + i] = acc[i];
                } else {
#pragma unroll
                    for (int i = 0; i < R; ++i) bits_g[(size_t)cb * rows + (row0 - xc) + i] = acc[i];
                }
                // lane 31 wrote the boundary scores, so lane 31 publishes the progress (program order +
                // release); remote mirrors first, the local counter last
                if (kCluster) {
                    ptx::st_release_cluster_if(lane31 && mirror_prev != 0u, mirror_prev, cb + 1);
                    ptx::st_release_cluster_if(lane31 && mirror_next != 0u, mirror_next, cb + 1);
                }
                ptx::st_release_shared_if(lane31, &done[1 + warp], cb + 1);
                __syncwarp();                                  // every lane has read the box: refill the slot
                ptx::tma_load_3d_if(lane0 && cb + S <= cbend, tile, &tmap, &my_full[slot], box_bytes, (cb + S) * kBlk, x0, b);
                if (++slot == S) {
                    slot = 0;
                    parity ^= 1u;
                }
                if (kDbg) {
                    t_wait_prev += t2 - t0;
                    t_wait_tma += t3 - t2;
                    t_sweep += clock64() - t3;
                }
            }
            if (kCluster) {
                ptx::st_release_cluster_if(lane31 && mirror_prev != 0u, mirror_prev, kDoneAll);
                ptx::st_release_cluster_if(lane31 && mirror_next != 0u, mirror_next, kDoneAll);
            }
            ptx::st_release_shared_if(lane31, &done[1 + warp], kDoneAll);
            // a NaN or an infinity anywhere in this token's history is still in its score now
#pragma unroll
            for (int i = 0; i < R; ++i)
                if (row0 + i < tx && !(fabsf(v[i]) <= 3.402823466e38f)) nonfinite = 1;
        }
        if (kDbg && dbg && lane == 0) {
            dbg[1] = clock64();
            dbg[2] = t_wait_prev;
            dbg[4] = t_wait_tma;
            dbg[8] = t_sweep;
            dbg[10] = cbend - cb0 + 1;
        }
    } else {
        // ---- filler warp: zero this CTA's slice of the dense output while the sweep runs ----
        // One lane streams a shared zero page to global memory with bulk async copies (UBLKCP):
        // a few dozen instructions for the whole slab instead of a flood of vector stores that
        // would compete with the sweep warps for the load/store pipe.
        float4 *zero4 = reinterpret_cast<float4 *>(smem + plan.off_zero);
        for (int i = lane; i < kZeroBytes / 16; i += 32) zero4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        ptx::fence_proxy_async();                              // generic writes -> visible to the async proxy
        __syncwarp();
        const int my_rows = max(0, min(rows, T_x - xc));
        if (lane == 0 && my_rows > 0) {
            char *dst = reinterpret_cast<char *>(p.path + ((int64_t)b * T_x + xc) * T_y);
            const int64_t total = (int64_t)my_rows * T_y * 4;    // multiple of 16: T_y % 4 == 0 on this path
            for (int64_t off = 0; off < total; off += kZeroBytes) {
                const int64_t n = total - off;
                ptx::bulk_store_s2g(dst + off, zero4, (uint32_t)(n < kZeroBytes ? n : kZeroBytes));
            }
            ptx::bulk_commit_group();
            ptx::bulk_wait_all();                              // the ones are written after the next barrier
        }
        __syncwarp();
        if (kDbg && dbg && lane == 0) dbg[1] = clock64();
    }
    if (!bits_smem) __threadfence();
    // ---- were all scores finite?  (cluster-wide) ----
    const int any_bad = __syncthreads_or(nonfinite);
    int redo = any_bad;
    if (kCluster) {
        if (any_bad && tid == 0)
            for (int r = 0; r < K; ++r) ptx::st_cluster_u32(ptx::mapa(ptx::smem_u32(const_cast<int *>(&misc[3])), r), 1u);
        ptx::cluster_sync();
        redo = misc[3];
    }
    if (redo) {
        // non-finite scores: the sign trick is not the reference's compare there -- redo literally
        if (c == 0) {
            float *col = reinterpret_cast<float *>(smem + plan.off_ring);
            const float *val = p.value + (int64_t)b * p.value_stride_b;
            exact_sweep_cta0(val, p.value_stride_x, col, ptx::smem_u32(bits_s),
                             bits_smem ? nullptr : p.ws_bits + (size_t)b * K * plan.nblk * rows, rows, plan.nblk, tx, ty,
                             p.max_neg_val);
            if (!bits_smem) __threadfence();
        }
        if (kCluster)
            ptx::cluster_sync();
        else
            __syncthreads();
    }

    // ---- backtrack (core.pyx:32-35) by TOKENS, handed down from CTA to CTA ----
    if (kDbg && dbg && tid == 0) dbg[5] = clock64();
    const int c_last = (tx > 0) ? (tx - 1) / rows : -1;          // CTA that owns the last token
    if (tid == 0 && c <= c_last) {
        int x, y_hi;
        if (c == c_last) {
            x = tx - 1;
            y_hi = ty - 1;
        } else {                                             // (kCluster only: c < c_last)
            uint32_t spins = 0;
            while (ptx::ld_acquire_cluster_shared(const_cast<int *>(&misc[0])) == 0)
                if (++spins > kSpinLimit) spin_fail();
            x = misc[1];
            y_hi = misc[2];
        }
        const int x_min = max(xc, 1);
        if (x >= x_min)
            y_hi = bits_smem ? backtrack_tokens<true>(bits_s, rows, xc, x, y_hi, x_min, run)
                             : backtrack_tokens<false>(bits_g, rows, xc, x, y_hi, x_min, run);
        if (!kCluster || c == 0) {
            run[0] = make_int2(0, y_hi);
        } else {
            const uint32_t peer = ptx::mapa(ptx::smem_u32(const_cast<int *>(&misc[0])), (uint32_t)(c - 1));
            ptx::st_cluster_u32(peer + 4, (uint32_t)(xc - 1));
            ptx::st_cluster_u32(peer + 8, (uint32_t)y_hi);
            ptx::st_release_cluster_if(true, peer, 1);
        }
    }
    if (kDbg && dbg && tid == 0) dbg[6] = clock64();
    __syncthreads();

    // ---- dense path: ones, durations, frame -> token (this CTA's tokens) ----
    float *out = p.path + (int64_t)b * T_x * T_y;
    for (int xl = tid; xl < rows; xl += blockDim.x) {
        const int x = xc + xl;
        if (x >= T_x) break;
        int d = 0;
        if (x < tx) {
            const int2 r = run[xl];
            d = r.y - r.x + 1;
            float *row = out + (int64_t)x * T_y;
            for (int y = r.x; y <= r.y; ++y) row[y] = 1.f;
            if (p.frame_token)
                for (int y = r.x; y <= r.y; ++y) p.frame_token[(int64_t)b * T_y + y] = x;
        }
        if (p.durations) p.durations[(int64_t)b * T_x + x] = d;
    }
    if (p.frame_token && c == 0)
        for (int y = ty + tid; y < T_y; y += blockDim.x) p.frame_token[(int64_t)b * T_y + y] = -1;
    if (kDbg && dbg && lane == 0) dbg[7] = clock64();
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

// Tokens per lane (R) and sweep warps (W) for `tokens` tokens in one CTA.  A lone warp issues about
// one instruction every two cycles, so the per-frame cost is ~2 x (4R + 5) cycles, and never below
// the shuffle round trip amortised over R frames (~(29 + 10R)/R): R = 2..4 with at most one sweep
// warp per scheduler is the sweet spot; longer texts take more tokens per lane first, more warps
// second.
static bool choose_shape(int tokens, int &R, int &W) {
    const int groups = ceil_div(tokens, kBlk);         // 32-token groups
    if (groups <= 1) {
        R = 1;
        W = 1;
        return true;
    }
    for (int max_w : {4, 8, kMaxDpWarps})
        for (int r : {2, 3, 4, 5, 6, 8}) {
            const int w = ceil_div(groups, r);
            if (w <= max_w) {
                R = r;
                W = w;
                return true;
            }
        }
    return false;
}

static int g_force_cluster = 0;   // testing hook (mas_b200_debug_force_cluster): 0 = heuristic

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
    if (g_force_cluster > 0) return try_k(g_force_cluster, best) != 0;
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
    static int configured_smem[64][4] = {{0}};        // opt-in attribute is sticky per device: raise it only when needed
    int dev = 0;
    MAS_CUDA_TRY(cudaGetDevice(&dev));
    const bool dbgk = p.dbg_cycles != nullptr;        // profiling build of the same kernel (clock64 stamps)
    const bool clus = plan.K > 1;
    auto kern = clus ? (dbgk ? mas_path_systolic_kernel<R, kThreads, true, true> : mas_path_systolic_kernel<R, kThreads, false, true>)
                     : (dbgk ? mas_path_systolic_kernel<R, kThreads, true, false> : mas_path_systolic_kernel<R, kThreads, false, false>);
    if (plan.total > configured_smem[dev & 63][dbgk * 2 + clus]) {
        MAS_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, plan.total));
        configured_smem[dev & 63][dbgk * 2 + clus] = plan.total;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(p.B * plan.K));
    cfg.blockDim = dim3((unsigned)((plan.W + 1) * 32));
    cfg.dynamicSmemBytes = (size_t)plan.total;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)plan.K;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    MAS_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tmap, p, plan));
    return MAS_OK;
}

template <int R>
static int launch_r(const CUtensorMap &tmap, const PathParams &p, const Plan &plan, cudaStream_t stream) {
    if (plan.W <= 4) return launch_rt<R, 160>(tmap, p, plan, stream);
    return launch_rt<R, (kMaxDpWarps + 1) * 32>(tmap, p, plan, stream);
}

}  // namespace systolic

void path_systolic_force_cluster(int k) { systolic::g_force_cluster = k; }

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

    static int max_smem_cached[64] = {0}, num_sms_cached[64] = {0};
    int dev = 0;
    MAS_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return MAS_ERR_INVALID_ARGUMENT;
    if (max_smem_cached[dev] == 0) {
        MAS_CUDA_TRY(cudaDeviceGetAttribute(&max_smem_cached[dev], cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        MAS_CUDA_TRY(cudaDeviceGetAttribute(&num_sms_cached[dev], cudaDevAttrMultiProcessorCount, dev));
    }
    const int max_smem = max_smem_cached[dev] - 2048;   // static shared + alignment slack
    Plan plan{};
    if (!choose_plan(p.B, p.T_x, p.T_y, max_smem, num_sms_cached[dev], plan)) return MAS_ERR_UNSUPPORTED_SHAPE;
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

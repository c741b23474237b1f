// mas_dp_cta.cuh -- the per-CTA program of kernel (1) (sweep, exact redo, backtrack, dense output),
// shared by the stand-alone kernel (mas_path_systolic.cu) and the fused launch (mas_fused.cu).
// See mas_path_systolic.cu for the design notes.
#pragma once

#include <cstdlib>

#include <cuda.h>
#include <cudaTypedefs.h>

#include "mas_kernels.cuh"
#include "mas_ptx.cuh"

namespace mas {
namespace systolic {


constexpr int kBlk = 32;            // frames per box / per direction word
constexpr int kPubBlocks = 4;      // blocks between two publications of a sweep warp's progress
constexpr int kMaxDpWarps = 15;     // + 1 filler warp = 512 threads
constexpr int kBndBlocks = 16;      // depth of the warp-to-warp boundary ring, in 32-frame blocks
constexpr int kDoneAll = 0x3fffffff;
constexpr uint32_t kSpinLimit = 1u << 27;   // watchdog: a wedged wait traps instead of hanging the GPU

struct Plan {
    int R, W, S, K;      // tokens per lane, sweep warps per CTA, TMA ring depth, CTAs per utterance
    int rows;            // tokens per CTA: W * 32 * R
    int nblk;            // ceil(T_y / 32)
    int bits_in_smem;
    // byte offsets into dynamic shared memory (base is 1024-aligned)
    int zero_bytes;      // shared zero page the filler streams to the dense output with bulk copies
    int off_ring, off_bits, off_bnd, off_bar, off_done, off_misc, off_run, off_zero, total;
};

__host__ __device__ inline int stage_bytes(int R) { return kBlk * R * kBlk * 4; }   // 32R rows x 128 B

__host__ __device__ inline Plan make_plan(int R, int W, int S, int K, int T_y, bool bits_in_smem, int zero_bytes = 16384) {
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
    p.zero_bytes = zero_bytes;
    p.off_zero = off;
    off += zero_bytes;
    p.off_bnd = off;
    off += (W + 1) * kBndBlocks * kBlk * 4;       // ring w = boundary INTO warp w; ring 0 is constant -1e9
    p.off_bar = off;
    off += 2 * W * S * 8;                         // full[W][S] then empty[W][S]
    p.off_done = off;
    off += (W + 2) * 4;                           // [prev CTA's last warp | own warps | next CTA's first warp]
    p.off_misc = off;
    off += 8 * 4;                                 // backtrack hand-over (flag, token, frame), redo flag, mask sums
    p.total = (int)align_up((size_t)off, 16);
    return p;
}

__device__ __forceinline__ void spin_fail() { __trap(); }

// The threads that run one utterance's program: a whole CTA (hardware barrier 0), or a part of one
// with a named barrier of its own (round 1 measured two utterances per CTA that way: slower, not kept;
// the exact-sweep helpers still take the abstraction).
struct Team {
    int tid, nthr, bar;
    __device__ __forceinline__ void sync() const {
        if (bar == 0)
            __syncthreads();
        else
            asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nthr) : "memory");
    }
    __device__ __forceinline__ int sync_or(int pred) const {
        if (bar == 0) return __syncthreads_or(pred);
        uint32_t r;
        asm volatile(
            "{\n"
            ".reg .pred p, q;\n"
            "setp.ne.u32 q, %1, 0;\n"
            "bar.red.or.pred p, %2, %3, q;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(r)
            : "r"(pred), "r"(bar), "r"(nthr)
            : "memory");
        return (int)r;
    }
};
__device__ __forceinline__ Team whole_cta() { return Team{(int)threadIdx.x, (int)blockDim.x, 0}; }

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
// kOut: where lane `publisher` hands its last token's four scores: 0 = this CTA's shared memory,
// 1 = a shared::cluster address (plain remote store; the caller publishes progress with a release),
// 2 = a shared::cluster address by st.async, completing 16 bytes on the remote mbarrier `bar_out`
// (the consumer waits on that barrier: no fence on either side).
template <int R, int kOut>
__device__ __forceinline__ void sweep_group(const float4 (&L)[R], const float4 &b, float (&v)[R], uint32_t (&acc)[R],
                                            float &carry, uint32_t bnd_out, bool publisher, int g, uint32_t bar_out) {
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
    if (kOut == 2)
        ptx::st_async_v4_if(publisher, bnd_out + g * 16, make_float4(out4[0], out4[1], out4[2], out4[3]), bar_out);
    else if (kOut == 1)
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
template <int R, int kOut>
__device__ __forceinline__ void sweep_block(uint32_t tile, const uint32_t (&lane_c)[R], float (&v)[R], uint32_t (&acc)[R],
                                            float &carry, uint32_t bin, uint32_t bnd_out, bool publisher, uint32_t bar_out = 0u) {
    // tile: shared address of the staged box; lane_c[i] = this lane's row i inside it with its
    // swizzle term folded in (rows are 128-byte aligned, so byte offset | swizzle): the 16-byte group g
    // of row i sits at q[i] ^ (g << 4) -- one LOP3 with an immediate per load (the generic-pointer form
    // took three instructions, and a lone warp pays ~2.3 cycles for each).  bin: the previous warp's
    // boundary scores of this block.
    uint32_t q[R];
#pragma unroll
    for (int i = 0; i < R; ++i) q[i] = tile + lane_c[i];
    float4 LA[R], LB[R], bA, bB;
#pragma unroll
    for (int i = 0; i < R; ++i) LA[i] = ptx::ld_shared_v4(q[i]);
    bA = ptx::ld_shared_v4(bin);
#pragma unroll
    for (int g = 0; g < 8; g += 2) {
#pragma unroll
        for (int i = 0; i < R; ++i) LB[i] = ptx::ld_shared_v4(q[i] ^ (uint32_t)((g + 1) << 4));
        bB = ptx::ld_shared_v4(bin + (g + 1) * 16);
        sweep_group<R, kOut>(LA, bA, v, acc, carry, bnd_out, publisher, g, bar_out);
        if (g + 2 < 8) {
#pragma unroll
            for (int i = 0; i < R; ++i) LA[i] = ptx::ld_shared_v4(q[i] ^ (uint32_t)((g + 2) << 4));
            bA = ptx::ld_shared_v4(bin + (g + 2) * 16);
        }
        sweep_group<R, kOut>(LB, bB, v, acc, carry, bnd_out, publisher, g + 1, bar_out);
    }
}

// The same block with the shared-memory loads issued kAhead - 1 groups before their use: for a sweep
// warp that shares its SM with warps saturating the shared-memory pipe (the fused launch), where a
// load takes several times its idle latency.
template <int R, int kOut, int kAhead>
__device__ __forceinline__ void sweep_block_ahead(uint32_t tile, const uint32_t (&lane_c)[R], float (&v)[R], uint32_t (&acc)[R],
                                                  float &carry, uint32_t bin, uint32_t bnd_out, bool publisher, uint32_t bar_out) {
    uint32_t q[R];
#pragma unroll
    for (int i = 0; i < R; ++i) q[i] = tile + lane_c[i];
    float4 L[kAhead][R], bq[kAhead];
#pragma unroll
    for (int a = 0; a < kAhead - 1; ++a) {
#pragma unroll
        for (int i = 0; i < R; ++i) L[a][i] = ptx::ld_shared_v4(q[i] ^ (uint32_t)(a << 4));
        bq[a] = ptx::ld_shared_v4(bin + a * 16);
    }
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        if (g + kAhead - 1 < 8) {
#pragma unroll
            for (int i = 0; i < R; ++i) L[(g + kAhead - 1) % kAhead][i] = ptx::ld_shared_v4(q[i] ^ (uint32_t)((g + kAhead - 1) << 4));
            bq[(g + kAhead - 1) % kAhead] = ptx::ld_shared_v4(bin + (g + kAhead - 1) * 16);
        }
        sweep_group<R, kOut>(L[g % kAhead], bq[g % kAhead], v, acc, carry, bnd_out, publisher, g, bar_out);
    }
}

// Exact compare/select sweep for utterances whose scores are not all finite (NaN / +-inf) or whose
// mask is not all-ones (mas_mask.cu): CTA 0 of the cluster walks the frames for ALL tokens with a
// barrier per frame, score column in shared memory, scores read straight from global memory.
// Slow, rare, and literal: core.pyx:17-30 as written.  col: [2][tx] floats.  The direction words go
// where the fast sweep would have put them: word (cb, x) belongs to CTA x / rows, in its shared
// memory (bits_smem_addr, written over DSMEM) or in global memory (bits_g, [K][nblk][rows]).
// msk: null, or the utterance's mask (element strides ms_x, ms_y): the scores are then value * mask
// as monotonic_align/__init__.py:11 forms them.
// Frames [y_begin, y_end) of the sweep; `buf` = which half of `col` holds the scores after frame
// y_begin - 1 (exact_sweep_init sets half 0).  kCoherent: `val` was written by this CTA during this
// launch (the single launch's redo scratch): plain loads instead of the read-only path.
struct ExactBits {
    uint32_t smem_addr;      // shared::cta address of this CTA's bit table (used when g == nullptr)
    uint32_t *g;             // global bit table [K][nblk][rows], or nullptr
    int rows, nblk;
};
__device__ __forceinline__ void exact_sweep_init(const Team team, float *col, int tx, float neg) {
    for (int x = team.tid; x < tx; x += team.nthr) col[x] = neg;
    team.sync();
}
template <bool kCoherent>
static __device__ __noinline__ void exact_sweep_frames(const Team team, const float *val, int64_t stride_x, float *col, int &buf,
                                                       const ExactBits eb, int tx, int y_begin, int y_end, float neg,
                                                       const float *__restrict__ msk = nullptr, int64_t ms_x = 0,
                                                       int64_t ms_y = 0) {
    const int tid = team.tid, nthr = team.nthr;
    for (int y = y_begin; y < y_end; ++y) {
        const float *vin = col + buf * tx;
        float *vout = col + (buf ^ 1) * tx;
        for (int x = tid; x < tx; x += nthr) {
            const float stay = vin[x];                                      // == -1e9 while x > y-1, core.pyx:19-20
            const float adv = (x == 0) ? ((y == 0) ? 0.f : neg) : vin[x - 1];   // core.pyx:23-29
            const float *vp = val + (int64_t)x * stride_x + y;
            float l = (x > y) ? 0.f : (kCoherent ? *reinterpret_cast<const volatile float *>(vp) : __ldg(vp));
            if (msk != nullptr && x <= y) l *= __ldg(msk + (int64_t)x * ms_x + (int64_t)y * ms_y);
            const bool take = adv > stay;
            vout[x] = (take ? adv : stay) + l;
            const uint32_t bit = ((take || (x == y && x > 0)) ? 1u : 0u) << (y & 31);
            const int owner = x / eb.rows, xl = x - owner * eb.rows;
            const size_t word = (size_t)(y >> 5) * eb.rows + xl;
            if (eb.g == nullptr) {
                const uint32_t addr = ptx::mapa(eb.smem_addr + (uint32_t)word * 4u, owner);
                ptx::st_cluster_u32(addr, ((y & 31) ? ptx::ld_cluster_u32(addr) : 0u) | bit);
            } else {
                uint32_t *w = eb.g + (size_t)owner * eb.nblk * eb.rows + word;
                __stcg(w, ((y & 31) ? __ldcg(w) : 0u) | bit);
            }
        }
        buf ^= 1;
        team.sync();
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
    // One step per (token, block) word the path touches: either the path stepped onto x inside this
    // block (-> next token, same block) or it did not (-> same token, previous block).  A lone thread
    // issues an instruction every ~2.3 cycles and pays ~25 per branch, so: both successor words are
    // fetched before the decision, the decision is a handful of selects, four steps per loop trip.
    // In shared memory the successor loads are unconditional (a word before / a row above the table
    // is still this CTA's shared memory -- the plan puts the sweep ring there -- and is never used).
    int base = y_hi & ~31;
    uint32_t elig = 0xffffffffu >> (31 - (y_hi & 31));     // bits at or below the scan position
    const uint32_t *p = bits + (size_t)(y_hi >> 5) * rows + (x - xc);
    run -= xc;                                              // indexed by the global token number
    uint32_t w = kSmem ? *p : __ldcg(p);
    while (x >= x_min && base >= 0) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            uint32_t w_left, w_up;
            if (kSmem) {
                w_left = p[-1];                             // next token, same block
                w_up = *(p - rows);                         // same token, previous block
            } else {
                w_left = (x > x_min) ? __ldcg(p - 1) : 0u;
                w_up = (x >= x_min && base > 0) ? __ldcg(p - rows) : 0u;
            }
            const bool alive = x >= x_min && base >= 0;
            const uint32_t m = alive ? (w & elig) : 0u;
            const bool step = m != 0u;                      // the path stepped onto x inside this block
            const int lo = 31 - __clz(m);
            const int y_lo = base + lo;
            if (step) run[x] = make_int2(y_lo, y_hi);
            y_hi = step ? y_lo - 1 : y_hi;
            // dead steps (after the last token): stay put in global memory; in shared memory they may
            // wander a few rows up -- nothing there is dereferenced for use, and it saves a select
            p = step ? p - 1 : ((kSmem || alive) ? p - rows : p);
            elig = step ? (1u << lo) - 1u : 0xffffffffu;    // lo == 0: nothing left here, the next step goes up
            base -= (step || !alive) ? 0 : 32;
            x -= step ? 1 : 0;
            w = step ? w_left : w_up;
        }
    }
    return y_hi;
}

// The same backtrack by a whole WARP, for direction bits in the L2-resident WORKSPACE (long
// utterances), where a lone thread pays an L2 round trip (~700 cycles) per step and the walk was 28 % of
// a CTA's time (400 x 2000): the direction words of the 32 tokens below the current one come in with ONE
// coalesced load per 32-frame block (lane i: token x - i) and the steps inside the block run on shuffles;
// only moving to the block above costs a load again (404 -> 379 us at B=256, 400 x 2000).  Every lane
// passes the same arguments and gets the same result; lane 0 writes the run table.  For bits in SHARED
// memory the tuned single-thread walk above is faster (62 against ~100 cycles per step: measured,
// profiles/r2_sweep_k.txt), so that case keeps it.
template <bool kSmem>
__device__ __forceinline__ int backtrack_tokens_warp(const uint32_t *bits, int rows, int xc, int x, int y_hi, int x_min, int2 *run,
                                                     int lane) {
    run -= xc;                                              // indexed by the global token number
    int scan = y_hi;                                        // frame the search for the step onto x starts from
    while (x >= x_min && scan >= 0) {
        const int cb = scan >> 5, col0 = cb << 5;
        const uint32_t *row = bits + (size_t)cb * rows - xc;
        const int t = x - lane;
        const uint32_t mine = (t >= x_min) ? (kSmem ? row[t] : __ldcg(row + t)) : 0u;
        int y = scan & 31;
        bool up = false;                                    // continue in the block above?
#pragma unroll 1
        for (int i = 0; i < 32 && x >= x_min; ++i) {
            const uint32_t w = __shfl_sync(0xffffffffu, mine, i) & (0xffffffffu >> (31 - y));
            if (w == 0u) {                                  // on this token since before the block
                up = true;
                break;
            }
            const int lo = 31 - __clz(w);                   // stepped onto x at frame col0 + lo
            if (lane == 0) run[x] = make_int2(col0 + lo, y_hi);
            y_hi = col0 + lo - 1;
            --x;
            if (lo == 0) {                                  // the next token ends at the block above's last frame
                up = true;
                break;
            }
            y = lo - 1;
        }
        scan = up ? col0 - 1 : col0 + y;                    // (else: 32 tokens inside one block: reload from the new x)
    }
    return y_hi;
}

// The whole per-CTA program of kernel (1): lengths, sweep, (exact redo), backtrack, dense output.
// kCluster: compiled with the distributed-shared-memory paths (K > 1); the K == 1 build carries none.
// `b`: utterance; `cta_tag`: index for the profiling buffer; warps beyond plan.W + 1 idle.
template <int R, bool kDbg, bool kCluster>
__device__ __forceinline__ void dp_cta(const CUtensorMap &tmap, const PathParams &p, const Plan &plan, unsigned char *smem,
                                       int b, int cta_tag, const Team team = whole_cta()) {
    float *s_len = reinterpret_cast<float *>(smem + plan.off_misc + 16);

    const int K = kCluster ? plan.K : 1;
    const int c = kCluster ? (int)ptx::cluster_ctarank() : 0;    // which slice of the utterance's tokens
    const int tid = team.tid, lane = tid & 31;
    // broadcast from lane 0 so that the compiler KNOWS the warp index is warp-uniform: the role
    // branches and block loops below are then uniform control flow and the per-frame SHFL.UP of the
    // sweep is a plain shuffle (with `tid >> 5` every one of them was wrapped in WARPSYNC.COLLECTIVE /
    // ENDCOLLECTIVE plus three register moves, five extra issue slots per frame of a lone warp)
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int W = plan.W, S = plan.S, rows = plan.rows;
    const int T_x = p.T_x, T_y = p.T_y;
    const bool bits_smem = plan.bits_in_smem != 0;
    const int xc = c * rows;                                     // first token of this CTA

    float *ring = reinterpret_cast<float *>(smem + plan.off_ring);
    // packed directions [nblk][rows] of this CTA's tokens: shared memory when they fit, else workspace
    uint32_t *bits_s = reinterpret_cast<uint32_t *>(smem + plan.off_bits);
    uint32_t *bits_g = bits_smem ? nullptr : p.ws_bits + ((size_t)b * K + c) * plan.nblk * rows;
    float *bnd = reinterpret_cast<float *>(smem + plan.off_bnd);       // [W+1][kBndBlocks*32]; ring w = INTO warp w
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + plan.off_bar);   // [W][S] box landed (TMA complete_tx)
    uint64_t *empty = full + W * S;                                        // [W][S] box consumed (lane 0 of the sweep warp)
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
        team.sync();
        float sx = 0.f, sy = 0.f;
        const float *m = p.mask + (int64_t)b * p.mask_stride_b;
        for (int x = tid; x < T_x; x += team.nthr) sx += m[(int64_t)x * p.mask_stride_x];
        for (int y = tid; y < T_y; y += team.nthr) sy += m[(int64_t)y * p.mask_stride_y];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sx += __shfl_xor_sync(0xffffffffu, sx, o);
            sy += __shfl_xor_sync(0xffffffffu, sy, o);
        }
        if (lane == 0) {
            atomicAdd(&s_len[0], sx);
            atomicAdd(&s_len[1], sy);
        }
        team.sync();
        tx_raw = (int)s_len[0];
        ty_raw = (int)s_len[1];
    }
    const Lengths len = clamp_lengths(tx_raw, ty_raw, T_x, T_y);
    const int tx = len.tx, ty = len.ty;

    // ---- per-warp geometry: tokens [x0, x0+32R), 32-frame blocks [cb0, cbend] ----
    constexpr int rows_per_warp = kBlk * R;
    const int x0 = xc + warp * rows_per_warp;
    const bool dp_warp = warp < W;
    const bool filler_warp = warp == W;                         // warps beyond W + 1 (fused launch) only meet the barriers
    const bool active = dp_warp && x0 < tx;
    int cb0 = 0, cbend = -1;
    if (active) {
        const int x1 = min(x0 + rows_per_warp, tx) - 1;
        cb0 = x0 >> 5;                                       // token x0 enters the band at frame x0
        cbend = min(ty - 1, x1 + (ty - tx)) >> 5;            // ... and token x1 leaves it here (core.pyx:18)
    }
    if (tid == 0) {
        for (int i = 0; i < 2 * W * S; ++i) ptx::mbar_init(&full[i], 1);
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
        for (int i = tid; i < kBndBlocks * kBlk; i += team.nthr) bnd[i] = p.max_neg_val;
    // done[1+w] = number of 32-frame blocks warp w has finished.  It starts one short of the warp's
    // first block: the warp still needs the LAST frame of block cb0-1 from its predecessor (the
    // diagonal cell of token x0-1), so that ring slot must not be recycled yet.
    if (dp_warp && lane == 0) done[1 + warp] = active ? cb0 - 1 : kDoneAll;
    if (kCluster)
        ptx::cluster_sync();
    else
        team.sync();

    long long *dbg = (kDbg && p.dbg_cycles && warp < 16) ? p.dbg_cycles + ((size_t)cta_tag * 16 + warp) * 16 : nullptr;
    long long t_wait_prev = 0, t_wait_tma = 0, t_sweep = 0, t_core = 0;
    if (kDbg && dbg && lane == 0) dbg[0] = clock64();
    if (kDbg && dbg && tid == 0) dbg[11] = ptx::globaltimer_ns();
    int nonfinite = 0;
    if (dp_warp) {
        if (active) {
            float *my_ring = ring + (size_t)warp * S * (rows_per_warp * kBlk);
            uint64_t *my_full = full + warp * S;
            uint64_t *my_empty = empty + warp * S;
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
            // every shared address the loop touches, as 32-bit shared-window addresses in registers
            constexpr uint32_t kStage = rows_per_warp * kBlk * 4;
            const uint32_t ring_a = ptx::smem_u32(my_ring), full_a = ptx::smem_u32(my_full), empty_a = ptx::smem_u32(my_empty);
            const uint32_t bnd_in_a = ptx::smem_u32(bnd_in_base);
            const uint32_t done_prev_a = ptx::smem_u32(&done[warp]), done_self_a = ptx::smem_u32(&done[1 + warp]),
                           done_next_a = ptx::smem_u32(&done[warp + 2]);
            uint32_t bits_a = ptx::smem_u32(bits_s + (size_t)cb0 * rows + (row0 - xc));   // advances one block row per block
            uint32_t *bits_gp = bits_smem ? nullptr : bits_g + (size_t)cb0 * rows + (row0 - xc);
            uint32_t lane_c[R];
#pragma unroll
            for (int i = 0; i < R; ++i) {
                const int row = lane * R + i;
                lane_c[i] = (uint32_t)(row * kBlk * 4) | (uint32_t)((row & 7) << 4);
            }

            // Progress is exchanged with the neighbour warps once per kPub blocks: the polls, the release
            // store (a MEMBAR) and the mirrors cost a lone warp a few hundred cycles, as much as a third
            // of a block's sweep.  The price is one more block of skew per warp, and the boundary ring
            // must be deep enough for producer and consumer to overlap (with 4 blocks and steps of 2 they
            // alternated: 41 -> 58 us at 200 x 1000).  No deadlock while
            // 2 kPubBlocks <= kBndBlocks: a producer stuck before the step ending at block h_p has published
            // h_p - kPubBlocks + 1 and needs its consumer past h_p - kBndBlocks; the consumer stuck before
            // its step ending at h_c needs the producer past h_c -- both at once would need
            // h_c >= h_p - kPubBlocks + 1 and h_c <= h_p - kBndBlocks + kPubBlocks - 1.
            static_assert(2 * kPubBlocks <= kBndBlocks, "boundary ring too shallow for the publication step");
            // Measured (profiles/r1_sweep_k.txt): with the neighbour in another CTA (polls and mirrors
            // over DSMEM) four blocks per publication take 22-27 % off the 1024 x 8192 cluster shapes;
            // inside one CTA the extra skew costs more than the saved polls (41.1 -> 42.8 us at 200 x 1000),
            // so K = 1 publishes every block.
            constexpr int kPub = kCluster ? kPubBlocks : 1;
            for (int cbs = cb0; cbs <= cbend; cbs += kPub) {
                const int cb_hi = min(cbs + kPub - 1, cbend);
                uint32_t spins = 0;
                const long long t0 = kDbg ? clock64() : 0;
                while (seen_prev <= cb_hi) {                    // previous warp has published blocks .. cb_hi
                    seen_prev = prev_remote ? ptx::ld_acquire_cluster_shared_a(done_prev_a) : ptx::ld_acquire_shared_a(done_prev_a);
                    if (++spins > kSpinLimit) spin_fail();
                }
                while (seen_next + kBndBlocks <= cb_hi) {       // next warp has consumed block cb_hi - ring depth
                    seen_next = next_remote ? ptx::ld_acquire_cluster_shared_a(done_next_a) : ptx::ld_acquire_shared_a(done_next_a);
                    if (++spins > kSpinLimit) spin_fail();
                }
                if (kDbg) t_wait_prev += clock64() - t0;
              for (int cb = cbs; cb <= cb_hi; ++cb) {
                const long long t2 = kDbg ? clock64() : 0;
                const uint32_t slot8 = (uint32_t)slot * 8u;
                while (!ptx::mbar_try_wait_a(full_a + slot8, parity))
                    if (++spins > kSpinLimit) spin_fail();
                const long long t3 = kDbg ? clock64() : 0;
                if (cb == cb0 && x0 > 0)                        // score of token x0-1 on the diagonal frame x0-1
                    carry = ptx::ld_shared_f32_a(bnd_in_a + (uint32_t)((((cb0 - 1) & (kBndBlocks - 1)) * kBlk + (kBlk - 1)) * 4));

#pragma unroll
                for (int i = 0; i < R; ++i) acc[i] = 0u;
                const uint32_t tile_a = ring_a + (uint32_t)slot * kStage;
                const uint32_t ring_slot = (uint32_t)(cb & (kBndBlocks - 1)) * (kBlk * 4);
                const int col0 = cb * kBlk;
                const bool on_diagonal = cb < cb0 + R;        // warp-uniform
                if (on_diagonal) {
                    zero_below_diagonal<R>(my_ring + (size_t)slot * rows_per_warp * kBlk, lane, row0, col0);
                    ptx::fence_proxy_async();               // these generic writes precede the TMA refill of the slot
                    __syncwarp();
                }
                const long long t4 = kDbg ? clock64() : 0;
                sweep_block<R, kCluster ? 1 : 0>(tile_a, lane_c, v, acc, carry, bnd_in_a + ring_slot, bnd_out_base + ring_slot, publisher);
                if (kDbg) t_core += clock64() - t4;
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
                    for (int i = 0; i < R; ++i) ptx::st_shared_u32_a(bits_a + 4u * i, acc[i]);
                    bits_a += (uint32_t)rows * 4u;
                } else {
#pragma unroll
                    for (int i = 0; i < R; ++i) bits_gp[i] = acc[i];
                    bits_gp += rows;
                }
                __syncwarp();                                  // every lane has read the box: hand the slot back
                ptx::mbar_arrive_if_a(lane0, empty_a + slot8);  // the loader warp refills it
                if (++slot == S) {
                    slot = 0;
                    parity ^= 1u;
                }
                if (kDbg) {
                    t_wait_tma += t3 - t2;
                    t_sweep += clock64() - t3;
                }
              }
                // lane 31 wrote the boundary scores, so lane 31 publishes the progress (program order +
                // release); remote mirrors first, the local counter last
                if (kCluster) {
                    ptx::st_release_cluster_if(lane31 && mirror_prev != 0u, mirror_prev, cb_hi + 1);
                    ptx::st_release_cluster_if(lane31 && mirror_next != 0u, mirror_next, cb_hi + 1);
                }
                ptx::st_release_shared_if_a(lane31, done_self_a, cb_hi + 1);
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
            dbg[9] = t_core;
            dbg[10] = cbend - cb0 + 1;
        }
    } else if (filler_warp) {
        // ---- loader / filler warp ----
        // (1) zero this CTA's slice of the dense output with bulk async copies of a shared zero page
        //     (UBLKCP: a few dozen instructions for the whole slab instead of a flood of vector
        //     stores that would compete with the sweep warps for the load/store pipe);
        // (2) feed every sweep warp's ring: wait for a slot to be handed back (`empty`), arm `full`,
        //     issue the TMA box.
        float4 *zero4 = reinterpret_cast<float4 *>(smem + plan.off_zero);
        for (int i = lane; i < plan.zero_bytes / 16; i += 32) zero4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        ptx::fence_proxy_async();                              // generic writes -> visible to the async proxy
        __syncwarp();
        {
            // lane w feeds sweep warp w (state in registers, the W lanes poll in parallel); lane 0
            // also drips one zero-fill bulk copy per pass
            constexpr uint32_t box_bytes = kBlk * R * kBlk * 4;
            if (lane == 0) ptx::prefetch_tensormap(&tmap);
            const int wx0 = xc + lane * rows_per_warp;
            const bool feeds = lane < W && wx0 < tx;
            const int w_cb0 = feeds ? wx0 >> 5 : 0;
            const int w_cbend = feeds ? min(ty - 1, min(wx0 + rows_per_warp, tx) - 1 + (ty - tx)) >> 5 : -1;
            int k = 0;                                          // boxes issued so far
            const int my_rows = max(0, min(rows, T_x - xc));
            char *zdst = reinterpret_cast<char *>(p.path + ((int64_t)b * T_x + xc) * T_y);
            const int64_t ztotal = (int64_t)my_rows * T_y * 4;  // multiple of 16: T_y % 4 == 0 on this path
            int64_t zoff = lane == 0 ? 0 : ztotal;
            uint32_t spins = 0;
            for (;;) {
                const bool want = w_cb0 + k <= w_cbend;
                if (!__any_sync(0xffffffffu, want || zoff < ztotal)) break;
                bool go = false;
                if (want) {
                    const int sl = k % S;
                    // a slot's n-th reuse waits for its n-th hand-back; the first S boxes need none
                    bool ok = k < S || ptx::mbar_test_wait(&empty[lane * S + sl], (uint32_t)(((k / S) & 1) ^ 1));
                    const int f0 = (w_cb0 + k) * kBlk;
                    if (ok) {
                        float *dst = ring + ((size_t)lane * S + sl) * (rows_per_warp * kBlk);
                        ptx::mbar_arrive_expect_tx(&full[lane * S + sl], box_bytes);
                        ptx::tma_load_3d(dst, &tmap, &full[lane * S + sl], f0, wx0, b);
                        ++k;
                        go = true;
                    }
                }
                if (zoff < ztotal) {
                    const int64_t n = ztotal - zoff;
                    ptx::bulk_store_s2g(zdst + zoff, zero4, (uint32_t)(n < plan.zero_bytes ? n : plan.zero_bytes));
                    zoff += plan.zero_bytes;
                    go = true;
                }
                if (!__any_sync(0xffffffffu, go)) {
                    __nanosleep(32);
                    if (++spins > kSpinLimit) spin_fail();
                } 
            }
            if (lane == 0 && ztotal > 0) {
                ptx::bulk_commit_group();
                ptx::bulk_wait_all();                          // the ones are written after the next barrier
            }
        }
        __syncwarp();
        if (kDbg && dbg && lane == 0) dbg[1] = clock64();
    }
    if (!bits_smem) __threadfence();
    // ---- were all scores finite?  (cluster-wide) ----
    const bool masked = p.exact_flag != nullptr && p.exact_flag[b] != 0;   // value * mask is not value here
    const int any_bad = team.sync_or(nonfinite | (masked ? 1 : 0));
    int redo = any_bad;
    if (kCluster) {
        if (any_bad && tid == 0)
            for (int r = 0; r < K; ++r) ptx::st_cluster_u32(ptx::mapa(ptx::smem_u32(const_cast<int *>(&misc[3])), r), 1u);
        ptx::cluster_sync();
        redo = misc[3];
    }
    bool bits_from_smem = bits_smem;
    if (redo) {
        // non-finite scores (or a mask that is not all-ones): the sign trick is not the reference's
        // compare there -- redo literally
        if (c == 0) {
            float *col = reinterpret_cast<float *>(smem + plan.off_ring);
            const float *val = p.value + (int64_t)b * p.value_stride_b;
            const ExactBits eb{ptx::smem_u32(bits_s), bits_smem ? nullptr : p.ws_bits + (size_t)b * K * plan.nblk * rows, rows,
                               plan.nblk};
            int buf = 0;
            exact_sweep_init(team, col, tx, p.max_neg_val);
            exact_sweep_frames<false>(team, val, p.value_stride_x, col, buf, eb, tx, 0, ty, p.max_neg_val,
                                      masked ? p.mask + (int64_t)b * p.mask_stride_b : nullptr, p.mask_stride_x,
                                      p.mask_stride_y);
            if (!bits_smem) __threadfence();
        }
        if (kCluster)
            ptx::cluster_sync();
        else
            team.sync();
    }

    // ---- backtrack (core.pyx:32-35) by TOKENS, handed down from CTA to CTA ----
    if (kDbg && dbg && tid == 0) {
        dbg[5] = clock64();
        dbg[12] = ptx::globaltimer_ns();
    }
    const int c_last = (tx > 0) ? (tx - 1) / rows : -1;          // CTA that owns the last token
    if (warp == 0 && c <= c_last) {                              // (the whole first warp, uniformly)
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
        if (x >= x_min) {
            if (!bits_from_smem)
                y_hi = backtrack_tokens_warp<false>(bits_g, rows, xc, x, y_hi, x_min, run, lane);
            else if (lane == 0)
                y_hi = backtrack_tokens<true>(bits_s, rows, xc, x, y_hi, x_min, run);
            y_hi = __shfl_sync(0xffffffffu, y_hi, 0);
        }
        __syncwarp();
        if (lane == 0) {
            if (!kCluster || c == 0) {
                run[0] = make_int2(0, y_hi);
            } else {
                const uint32_t peer = ptx::mapa(ptx::smem_u32(const_cast<int *>(&misc[0])), (uint32_t)(c - 1));
                ptx::st_cluster_u32(peer + 4, (uint32_t)(xc - 1));
                ptx::st_cluster_u32(peer + 8, (uint32_t)y_hi);
                ptx::st_release_cluster_if(true, peer, 1);
            }
        }
    }
    if (kDbg && dbg && tid == 0) dbg[6] = clock64();
    team.sync();

    // ---- dense path: ones, durations, frame -> token (this CTA's tokens) ----
    float *out = p.path + (int64_t)b * T_x * T_y;
    for (int xl = tid; xl < rows; xl += team.nthr) {
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
        for (int y = ty + tid; y < T_y; y += team.nthr) p.frame_token[(int64_t)b * T_y + y] = -1;
    if (kDbg && dbg && lane == 0) dbg[7] = clock64();
    if (kDbg && dbg && tid == 0) dbg[13] = ptx::globaltimer_ns();
}

// ---------------------------------------------------------------------------------------------
// host helpers shared by the launchers
// ---------------------------------------------------------------------------------------------
inline PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    // resolved once (C++11 magic static: thread-safe)
    static const PFN_cuTensorMapEncodeTiled_v12000 fn = [] {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
        return static_cast<PFN_cuTensorMapEncodeTiled_v12000>(nullptr);
    }();
    return fn;
}

// Tokens per lane (R) and sweep warps (W) for `tokens` tokens in one CTA.  The sweep is a chain of
// (blocks + W - 1) steps; what a step costs was measured per R on B200 (32-frame blocks, cycles;
// `python profiles/measure_step_costs.py` re-measures the table on the device at hand): the in-order warp eats one shuffle latency per frame
// whatever R is, so two or three tokens per lane cost the same per step and three need fewer warps
// (less skew); even R cost more than their instruction count says -- lane l reads rows l R .. l R + R - 1 of
// the TMA box, the hardware's 128-byte swizzle keys a row by (row & 7), so with an even R the eight lanes
// of a quarter warp share keys and every LDS.128 is a 2- (R = 2, 6) to 8-way (R = 8) bank conflict; R = 4
// (4-way) is only taken when nothing else fits.  (Kernel (2) writes its score ring itself and keys the
// rows by their LANE instead: mas_fused.cu, row_key.)
inline bool choose_shape(int tokens, int &R, int &W) {
    const int groups = ceil_div(tokens, kBlk);         // 32-token groups
    if (groups <= 1) {
        R = 1;
        W = 1;
        return true;
    }
    static const char *force_r = getenv("MAS_B200_FORCE_R");   // experiment hook
    if (force_r != nullptr) {
        R = atoi(force_r);
        W = ceil_div(groups, R);
        return R >= 1 && R <= 8 && R != 7 && W <= kMaxDpWarps;
    }
    static const int kR[] = {2, 3, 5, 6, 4, 8};
    static const int kStep[] = {1950, 1955, 2130, 2500, 3100, 3400};
    long best = -1;
    for (int i = 0; i < 6; ++i) {
        const int w = ceil_div(groups, kR[i]);
        if (w > kMaxDpWarps) continue;
        const long cost = (long)(40 + w - 1) * kStep[i];                 // ~40 blocks: a 1 300-frame utterance
        if (best < 0 || cost < best) {
            best = cost;
            R = kR[i];
            W = w;
        }
    }
    return best >= 0;
}

}  // namespace systolic
}  // namespace mas

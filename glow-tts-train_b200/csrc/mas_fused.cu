// mas_fused.cu -- kernel (2): log-likelihood + alignment search in ONE launch (models.py:362-382),
// any batch size, the [B,T_x,T_y] score matrix never materialised -- not in HBM, not in L2: every
// score is produced and consumed inside one SM's shared memory.
//
// One thread-block CLUSTER of K CTAs per utterance (K = 1, 2, 4 or 8; persistent: cluster j takes
// utterances j, 2 NC - 1 - j, 2 NC + j, ... of the batch).  The utterance's TOKENS are sliced over
// the K CTAs (slice = ceil(t_x / K), by the utterance's actual length), and a CTA does everything
// for its slice:
//   * FFMA warps (15 of the 16) contract the slice's scores over the mel channels, chunk of F frames
//     by chunk, into a SCORE RING in shared memory: boxes of [slice tokens x 32 frames] fp32 in the
//     128-byte-swizzled layout kernel (1) stages with TMA.  The warps work in independent TEAMS with a
//     named barrier each (chunk j belongs to team j mod nteams), so that finished scores appear every
//     F frames instead of every nteams x F.  The token-side operands (-0.5 exp(-2 logs), m exp(-2 logs))
//     stay in shared memory for the whole utterance; z arrives in panels of 16 channels by TMA
//     (one elected thread per team issues [16 channels x F frames] boxes into a 3-stage ring with
//     full/empty mbarriers: no thread spends an instruction on staging, no barrier per panel).
//     Arithmetic, operand order and the row constants' summation order are those of the
//     materialising kernel (mas_logp_cta.cuh): bit-identical scores.
//   * ONE sweep warp (warp 0) runs kernel (1)'s recurrence (mas_dp_cta.cuh: sweep_block) over the
//     ring, 32 frames per step: waits for the teams' chunk counters (shared memory, acquire), takes the
//     score of the slice's predecessor token from the previous CTA's sweep through distributed shared
//     memory and hands its own last token's to the next CTA (one 128-byte bulk copy per block into the
//     neighbour's 16-block boundary ring, completing its bytes on the neighbour's mbarrier: data and
//     signal in one message, no fence in the loop; credits come back as a plain remote store), packs
//     the direction bits, frees
//     the box (`consumed`, the producers' back-pressure), and drips bulk copies of a zero page into
//     the dense output on the way.
//   * backtrack, CTA K-1 -> 0 over DSMEM.  Short slices (<= 64 tokens): while the sweep runs, the warps
//     whose contraction is done tabulate per 32-frame block the BLOCK MAP (token at the block's last
//     frame -> token at the previous block's last frame) and the EXIT TABLE (the same composed all the
//     way down: the frame where the path leaves the slice); a hop through a CTA is then the entry
//     block's walk and one look-up, the (token, frame) hand-over an 8-byte st.async onto the next
//     CTA's mbarrier, and the per-block walks that record the path run one warp per block.  Long
//     slices walk tokens CTA after CTA like kernel (1).  Ones, durations, frame -> token by all threads.
// Only the cells the reference's band touches are ever contracted (core.pyx:18): a slice starts at
// the 32-frame block of its first token and ends where its last token leaves the band.
//
// Utterances with a non-finite score (NaN / inf inputs) are detected by the sweep as in kernel (1)
// and recomputed literally by CTA 0 of the cluster: scores chunk by chunk (the materialising
// kernel's program) into a one-chunk scratch, compare/select sweep, direction bits in the workspace.
//
// Shapes the launch does not take (frame count not a multiple of 4, more than 80 channels, slices
// that do not fit an SM) run as the two kernels back to back over groups of utterances (mas_api.cu).
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "mas_dp_cta.cuh"
#include "mas_logp_cta.cuh"

namespace mas {
namespace fused {

using systolic::kBlk;
using systolic::kBndBlocks;
using systolic::kSpinLimit;

constexpr int kThreads = 512;
constexpr int kChan = 16;              // channels per staged panel of z
constexpr int kStages = 3;             // z panels in flight per team
constexpr int kZeroPage = 8192;        // bytes of zeros the dense output is filled from
constexpr int kMaxTeams = 8;
constexpr int kBarFfma = 15;           // named barrier of all FFMA threads; teams use 1 .. nteams
constexpr int kBarAll = 14;            // ... and the sweep warp (operand hand-over, row constants ready)
constexpr int kMaxClusters = 192;      // workspace bound (no device query in the size function)

// control words in shared memory (ints)
// kTeamDone[t]: warps of team t x chunks they have stored; kConsumed: boxes swept; kCredit: blocks of
// this CTA's boundary scores the next CTA has consumed (written remotely)
// kBtToken / kBtFrame: the backtrack's hand-over from the CTA above (one 8-byte st.async)
enum Ctl { kTeamDone = 0, kConsumed = 8, kCredit = 11, kBtToken = 12, kBtFrame = 13, kRedo = 15, kCtlInts = 16 };

struct Geom {
    int K, NC;                         // CTAs per cluster, clusters in the grid
    int R;                             // tokens per sweep lane (the kernel's template parameter)
    int ffma_all;                      // 15 FFMA warps (all but the sweep warp) or 12 (the sweep warp's scheduler stays free)
    int nteams, team_warps;
    int max_slice, tr_max, ring_rows;  // tokens per CTA (bound), the same rounded up to 4 / to 8
    int CG_cap, F_cap;                 // column groups / frames per chunk (bound)
    int NB;                            // score ring depth in 32-frame boxes
    int nblk, bits_in_smem;
    int ops_tmp3;                      // the ring has room for stage_ops' third scratch array (exp(-2 logs))
    int use_maps;                      // build the backtrack's block maps while the sweep runs (short slices only)
    int use_exit;                      // ... and with them the exit table (build_block_maps)
    int passes;                        // contraction passes of a full-length utterance (host estimate)
    int nsh, dsh;                      // channel shares of the row constants, as the materialising kernel sums them
    int off_zero, off_bnd, off_run, off_xend, off_ctl, off_bar, off_big, off_ops, off_l14, off_part, off_z, off_l2, off_ring, off_bits,
        off_maps, off_exit, off_erow, total;
    uint32_t *ws_bits;                 // [NC][K][nblk][ring_rows] when the bits do not fit shared memory
    unsigned char *ws_maps;            // ... and the backtrack's block maps [NC][K][nblk][ring_rows] bytes with them
    float *redo_scratch;               // [NC][T_x][t_ref.F]
    uint32_t *redo_bits;               // [NC][nblk][T_x + 64]
    TileShape t_ref;                   // the materialising kernel's tile (row-constant shares, redo)
};

// One utterance as one CTA of its cluster sees it.
struct Utt {
    int b, tx, ty;
    int c;                             // rank in the cluster
    int n_c;                           // tokens per slice (multiple of R)
    int x0, n_real;                    // first token / real tokens of this CTA's slice
    int TR, RG, CG, F;                 // contraction tile: TR = 4 RG rows, chunks of F = 8 CG frames
    int cb0, cbend;                    // 32-frame blocks the slice is in the band for (cbend < cb0: none)
    int f_lo, f_hi;                    // frames [f_lo, f_hi) are contracted: the band of the slice, to multiples of 8
    int nch;                           // chunks to contract
};

// Swizzle of a score-ring box: the 16-byte group g of row q sits at group (g ^ key(q)).  kernel (1)'s TMA
// boxes use the hardware's key (q & 7); the sweep lane l reads rows l R .. l R + R - 1, so with an even R
// the eight lanes of a quarter warp share keys and every LDS.128 of the block is a 2- (R = 2) to 8-way
// (R = 8) bank conflict.  This ring is written by the FFMA warps, not by TMA, so the key is free:
// the LANE that owns the row -- eight consecutive lanes, eight different keys, for every R.
__device__ __forceinline__ uint32_t row_key(int q, int R) { return (uint32_t)((q / R) & 7); }

__device__ __forceinline__ void named_sync(int bar, int nthr) { asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nthr) : "memory"); }
__device__ __forceinline__ int ld_acquire_shared(const int *p) { return ptx::ld_acquire_shared_a(ptx::smem_u32(p)); }
__device__ __forceinline__ void st_release_shared(int *p, int v) { ptx::st_release_shared_if_a(true, ptx::smem_u32(p), v); }

// ---------------------------------------------------------------------------------------------
// FFMA side
// ---------------------------------------------------------------------------------------------

// Token-side operands of the slice, by all FFMA threads (fidx of nffma).  This is the cold start of
// the utterance and nothing else hides the first touch of x_m / x_logs, so the raw values come in
// with ONE round of 4-byte cp.async (a slice starts at any token: no 16-byte alignment) into the
// still empty score ring, and everything else reads shared memory: the element-wise operands by
// all threads, the row constants summed the way mas_logp_cta.cuh::stage_tokens does (nsh shares of
// dsh channels, ascending inside a share, shares added in order) from the same expressions, so
// that they are bit-identical.
// kIssue: only the copies are issued (before the utterance's cluster barrier: nothing of the previous
// utterance is left in the ring after its last barrier, and the copies fly while this one is passed).
template <bool kIssue>
__device__ __forceinline__ void stage_ops(const Geom &g, const Utt &u, const LogpParams &p, unsigned char *smem, int fidx, int nffma,
                                          long long *dbg) {
    const int D = p.D, T_x = p.T_x, TR = u.TR;
    float *sInv = reinterpret_cast<float *>(smem + g.off_ops), *sMiv = sInv + D * TR;
    float *sL1 = reinterpret_cast<float *>(smem + g.off_l14), *sL4 = sL1 + g.tr_max;
    float *sPart = reinterpret_cast<float *>(smem + g.off_part);
    // raw m (then -0.5 m^2), logs, exp(-2 logs): [D][TR] each
    float *tM = reinterpret_cast<float *>(smem + g.off_ring), *tL = tM + D * TR, *tR = tL + D * TR;
    const float *xm = p.x_m + (int64_t)u.b * D * T_x + u.x0;
    const float *xl = p.x_logs ? p.x_logs + (int64_t)u.b * D * T_x + u.x0 : nullptr;
    const int total = D * TR;
    // element e = fidx + k nffma is (channel d, token x): stepped without a division per element
    const int d_first = fidx / TR, x_first = fidx - d_first * TR, d_step = nffma / TR, x_step = nffma - d_step * TR;
    if (kIssue) {
        int d = d_first, x = x_first;
        for (int e = fidx; e < total; e += nffma) {
            if (x < u.n_real) {
                ptx::cp_async_4(tM + e, xm + (int64_t)d * T_x + x);
                if (xl) ptx::cp_async_4(tL + e, xl + (int64_t)d * T_x + x);
            }
            d += d_step, x += x_step;
            if (x >= TR) x -= TR, ++d;
        }
        ptx::cp_async_commit();
        return;
    }
    ptx::cp_async_wait<0>();
    named_sync(kBarFfma, nffma);
    if (dbg && fidx == 0) dbg[26] = ptx::globaltimer_ns();
    {
        int x = x_first;
        for (int e = fidx; e < total; e += nffma) {
            float inv = 0.f, miv = 0.f;
            if (x < u.n_real) {
                const float m = tM[e];
                const float r = xl ? expf(-2.0f * tL[e]) : 1.0f;      // models.py:363
                inv = -0.5f * r;                                      // models.py:368
                miv = m * r;                                          // models.py:371
                tM[e] = -0.5f * (m * m);
                if (g.ops_tmp3) tR[e] = r;
            }
            sInv[e] = inv;
            sMiv[e] = miv;
            x += x_step;
            if (x >= TR) x -= TR;
        }
    }
    if (dbg && fidx == 0) dbg[27] = ptx::globaltimer_ns();
    named_sync(kBarAll, nffma + 32);                       // the raw values are parked: the sweep warp sums the row constants
}

// The row constants l1, l4 of the slice's tokens, summed the way mas_logp_cta.cuh::stage_tokens does (nsh
// shares of dsh channels, ascending inside a share, shares added in order: bit-identical) from what
// stage_ops left in the ring.  By the SWEEP warp: it has nothing to sweep before the first chunk is
// stored a whole contraction pass later, the FFMA warps start contracting at once, and both meet
// at the barrier before the first store (team_contract).
__device__ __forceinline__ void sum_row_constants(const Geom &g, const Utt &u, const LogpParams &p, unsigned char *smem, int lane) {
    const int D = p.D, TR = u.TR;
    float *sL1 = reinterpret_cast<float *>(smem + g.off_l14), *sL4 = sL1 + g.tr_max;
    float *sPart = reinterpret_cast<float *>(smem + g.off_part);
    const float *tM = reinterpret_cast<const float *>(smem + g.off_ring), *tL = tM + D * TR, *tR = tL + D * TR;   // -0.5 m^2, logs, exp(-2 logs)
    const bool has_logs = p.x_logs != nullptr;
    for (int pair = lane; pair < g.nsh * TR; pair += 32) {
        const int h = pair / TR, x = pair - h * TR;
        const int d0 = h * g.dsh, d1 = min(D, d0 + g.dsh);
        float l1 = 0.f, l4 = 0.f;
        if (x < u.n_real) {
#pragma unroll 8
            for (int d = d0; d < d1; ++d) {
                const float ls = has_logs ? tL[d * TR + x] : 0.f;
                l1 += kNegHalfLog2Pi - ls;                          // models.py:364-366
                const float r = !has_logs ? 1.0f : g.ops_tmp3 ? tR[d * TR + x] : expf(-2.0f * ls);   // (no room: the same expression again)
                l4 = fmaf(tM[d * TR + x], r, l4);                   // models.py:373-375: -0.5 m^2 exp(-2 logs)
            }
        }
        sPart[(2 * h) * TR + x] = l1;
        sPart[(2 * h + 1) * TR + x] = l4;
    }
    __syncwarp();
    for (int x = lane; x < TR; x += 32) {
        float l1 = sPart[x], l4 = sPart[TR + x];
        for (int h = 1; h < g.nsh; ++h) {
            l1 += sPart[(2 * h) * TR + x];
            l4 += sPart[(2 * h + 1) * TR + x];
        }
        sL1[x] = l1;
        sL4[x] = l4;
    }
}

// The z panels of one team: a kStages-deep ring of [kChan channels][F frames] boxes, loaded by TMA.
// Panel n of the team's sequence (its chunks in order, npan panels each) lives in stage n % kStages.
struct ZPipe {
    uint32_t full_a, empty_a;          // shared addresses of full[kStages], empty[kStages]
    float *buf;                        // [kStages][kChan][F]
    int stage_floats;
    uint32_t box_bytes;
};

// One team's share of the slice's chunks: chunk j = team, team + nteams, ...  ttid of tn threads.
// `seq0`: panels this team has already consumed in this launch (the mbarrier phases run on across
// utterances).  Returns the new count.
template <bool kMeanOnly>
__device__ __forceinline__ int team_contract(const CUtensorMap &tmap_z, const Geom &g, const Utt &u, const LogpParams &p,
                                             unsigned char *smem, int *ctl, int team, int ttid, int tn, int seq0, bool prologue) {
    if (team >= u.nch) {                                    // more teams than chunks
        if (!prologue) named_sync(kBarAll, g.nteams * tn + 32);   // (the barrier the others pass before their first store)
        return seq0;
    }
    const int D = p.D, F = u.F, CG = u.CG, TR = u.TR, NB = g.NB;
    const float *sInv = reinterpret_cast<const float *>(smem + g.off_ops), *sMiv = sInv + D * TR;
    const float *sL1 = reinterpret_cast<const float *>(smem + g.off_l14), *sL4 = sL1 + g.tr_max;
    // per-frame sums of the mean_only mode, two copies used alternately: with few channels a fast warp is
    // through the next chunk's panels (all of them already landed) before a slow one has finished the
    // stores that read this chunk's sums; the per-chunk barrier keeps them less than two chunks apart
    float *sL2_base = reinterpret_cast<float *>(smem + g.off_l2) + team * 2 * g.F_cap;
    const uint32_t ring_a = ptx::smem_u32(smem + g.off_ring), box_bytes = (uint32_t)g.ring_rows * 128u;
    ZPipe zp;
    zp.stage_floats = kChan * g.F_cap;
    zp.buf = reinterpret_cast<float *>(smem + g.off_z) + (size_t)team * kStages * zp.stage_floats;
    zp.full_a = ptx::smem_u32(smem + g.off_bar) + (uint32_t)(team * 2 * kStages) * 8u;
    zp.empty_a = zp.full_a + kStages * 8u;
    zp.box_bytes = (uint32_t)zp.stage_floats * 4u;
    const int bar = 1 + team;
    const int lane = ttid & 31;
    const int rg = ttid / CG, cg = ttid - rg * CG;
    const bool worker = rg < u.RG;
    const int npan = ceil_div(D, kChan);
    const int my_chunks = (u.nch - team + g.nteams - 1) / g.nteams;
    const int total_n = my_chunks * npan;
    const volatile int *consumed = ctl + kConsumed;
    uint32_t spins = 0;
    // panel k of this utterance (sequence number seq0 + k): wait until every warp of the team has
    // handed the stage back, then arm `full` and issue the box
    auto issue = [&](int k) {
        if (k >= total_n) return;
        const int n = seq0 + k, st = n % kStages;
        if (n >= kStages) {
            const uint32_t parity = (uint32_t)((n / kStages - 1) & 1);
            while (!ptx::mbar_try_wait_a(zp.empty_a + st * 8u, parity))
                if (++spins > kSpinLimit) systolic::spin_fail();
        }
        const int jj = team + (k / npan) * g.nteams, pd = k - (k / npan) * npan;
        ptx::mbar_arrive_expect_tx_a(zp.full_a + st * 8u, zp.box_bytes);
        ptx::tma_load_3d_a(ptx::smem_u32(zp.buf + st * zp.stage_floats), &tmap_z, zp.full_a + st * 8u, u.f_lo + jj * F, pd * kChan, u.b);
    };
    if (prologue) {
        // the first panels, issued before the token side is staged (they land meanwhile)
        if (ttid == 0)
            for (int k = 0; k < kStages - 1; ++k) issue(k);
        return seq0;
    }
    int k = 0, count = 0;
    for (int j = team; j < u.nch; j += g.nteams) {
        GemmAcc acc;
#pragma unroll
        for (int i = 0; i < kGemmTM; ++i)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc.v[i][q] = 0ull;
        float l2 = 0.f;
        for (int pd = 0; pd < npan; ++pd, ++k) {
            if (ttid == 0) issue(k + kStages - 1);          // (its stage was panel k-1's: this very warp handed it back last)
            const int n = seq0 + k, st = n % kStages;
            while (!ptx::mbar_try_wait_a(zp.full_a + st * 8u, (uint32_t)((n / kStages) & 1)))
                if (++spins > kSpinLimit) systolic::spin_fail();
            const float *zb = zp.buf + st * zp.stage_floats;
            const int d0 = pd * kChan, cnt = min(kChan, D - d0);
            if (worker) {
                if (cnt == kChan)
                    gemm_tile_d<false, kChan, kMeanOnly>(sInv + d0 * TR, sMiv + d0 * TR, zb, cnt, TR, F, rg, cg, acc);
                else
                    gemm_tile_d<false, 0, kMeanOnly>(sInv + d0 * TR, sMiv + d0 * TR, zb, cnt, TR, F, rg, cg, acc);
            }
            if (kMeanOnly && ttid < F) {                    // models.py:367-369 with logs == 0: one sum per frame
                for (int d = 0; d < cnt; ++d) {
                    const float zv = zb[d * F + ttid];
                    l2 = fmaf(-0.5f * zv, zv, l2);
                }
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_a(zp.empty_a + st * 8u);
        }
        // the sweep warp has summed the row constants (the ring no longer holds the raw token-side values).
        // BEFORE the back-pressure wait: with a shallow ring the first round of chunks can already span more
        // boxes than the ring has, and the sweep -- which frees them -- passes this barrier only with every team
        if (count == 0) named_sync(kBarAll, g.nteams * tn + 32);
        if (ttid == 0) {
            // back-pressure: the boxes this chunk is stored into must have been swept
            const int bx_last = ((min(u.f_hi, u.f_lo + j * F + F) - 1) >> 5) - u.cb0;
            while (*consumed + NB <= bx_last) {
                __nanosleep(64);
                if (++spins > kSpinLimit) systolic::spin_fail();
            }
        }
        float *sL2 = sL2_base + (count & 1) * g.F_cap;
        if (kMeanOnly && ttid < F) sL2[ttid] = l2;
        named_sync(bar, tn);                                // the ring has room (and the frame sums are there)
        if (worker) {
            const int fb = u.f_lo + j * F;                  // the chunk's first frame
#pragma unroll
            for (int i = 0; i < kGemmTM; ++i) {
                const int xr = rg * kGemmTM + i;
                const float l1 = sL1[xr], l4 = sL4[xr];
                const uint32_t row_a = ring_a + (uint32_t)xr * 128u;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int yl = (F >> 1) * h + 4 * cg, f = fb + yl;
                    if (f >= u.f_hi) continue;              // beyond the band
                    float cq[4];
                    acc.quad(i, h, cq);
                    float4 r;
                    if (kMeanOnly) {
                        const float4 q2 = *reinterpret_cast<const float4 *>(sL2 + yl);
                        r.x = logp_cell_finish_mean_only(l1, q2.x, cq[0], l4);
                        r.y = logp_cell_finish_mean_only(l1, q2.y, cq[1], l4);
                        r.z = logp_cell_finish_mean_only(l1, q2.z, cq[2], l4);
                        r.w = logp_cell_finish_mean_only(l1, q2.w, cq[3], l4);
                    } else {
                        r.x = logp_cell_finish(l1, cq[0], l4);
                        r.y = logp_cell_finish(l1, cq[1], l4);
                        r.z = logp_cell_finish(l1, cq[2], l4);
                        r.w = logp_cell_finish(l1, cq[3], l4);
                    }
                    const int bx = (f >> 5) - u.cb0, slot = bx % NB, gq = (f & 31) >> 2;
                    ptx::st_shared_v4_if(true, row_a + (uint32_t)slot * box_bytes + (((uint32_t)gq ^ row_key(xr, g.R)) << 4), r);
                }
            }
        }
        ++count;
        // this warp's part of the chunk is stored: count it (release: the sweep reads the counter with acquire)
        __syncwarp();
        if (lane == 0) ptx::red_release_shared_add(ctl + kTeamDone + team, 1);
    }
    return seq0 + total_n;
}

// ---------------------------------------------------------------------------------------------
// sweep side (one warp)
// ---------------------------------------------------------------------------------------------

// Cells below the diagonal (token > frame) of a staged box -> 0 (see systolic::zero_below_diagonal);
// rows beyond the slice's tile are not touched.
template <int R>
__device__ __forceinline__ void zero_below_diagonal_rows(uint32_t tile_a, int lane, int row0, int col0, int nrows) {
#pragma unroll
    for (int i = 0; i < R; ++i) {
        const int q = lane * R + i;
        const int d = row0 + i - col0;              // frames [0, d) of this block are below the diagonal
        if (d <= 0 || q >= nrows) continue;
        const uint32_t row_a = tile_a + (uint32_t)q * 128u;
#pragma unroll
        for (int cidx = 0; cidx < 8; ++cidx) {
            if (4 * cidx >= d) break;
            const uint32_t a = row_a + (((uint32_t)cidx ^ row_key(q, R)) << 4);
            float4 x = ptx::ld_shared_v4(a);
            x.x = 0.f;
            if (4 * cidx + 1 < d) x.y = 0.f;
            if (4 * cidx + 2 < d) x.z = 0.f;
            if (4 * cidx + 3 < d) x.w = 0.f;
            ptx::st_shared_v4_if(true, a, x);
        }
    }
}

// The dense output's zeros: this CTA's share of the utterance's rows, dripped by lane 0 of the sweep
// warp as bulk copies of the shared zero page.
struct ZeroFill {
    char *dst;
    int64_t total, off;
    int per_block;        // copies per sweep block
    __device__ __forceinline__ void drip(const void *zero_page, int n) {
        for (int i = 0; i < n && off < total; ++i) {
            const int64_t left = total - off;
            ptx::bulk_store_s2g(dst + off, zero_page, (uint32_t)(left < kZeroPage ? left : kZeroPage));
            off += kZeroPage;
        }
    }
};

// Returns non-zero when a real token of the slice ended with a non-finite score.
//
// Boundary protocol between the sweeps of neighbouring CTAs (token x0-1 of the previous CTA feeds
// token x0 of this one): the previous CTA's publisher lane parks the 32 scores of a block in a local
// slot and sends them with ONE bulk copy into THIS CTA's boundary ring (kBndBlocks slots of 32
// frames), completing 128 bytes on the slot's mbarrier here; this warp arms a slot with
// arrive.expect_tx(128), waits for its phase, sweeps, re-arms it and returns a credit (a plain
// remote store of the number of consumed blocks -- it only guards the slot's reuse).  The previous
// CTA publishes blocks [max(its first block, my first block - 1), its last block]; beyond its last
// block the boundary token has left the band and whatever finite values the slot holds are never used.
template <int R, bool kDbg>
__device__ __forceinline__ int sweep_slice(const Geom &g, const Utt &u, unsigned char *smem, int *ctl, float neg, ZeroFill &zf,
                                           uint32_t *bits_g, long long *dbg) {
    const int lane = threadIdx.x & 31;
    int nonfinite = 0;
    long long t_chunks = 0, t_prev = 0, t_credit = 0, t_core = 0, t_bits = 0, t_done = 0, t_fill = 0, t_begin = 0, t_core_tail = 0, t_all_tail = 0;
    if (kDbg) t_begin = clock64();
    if (u.cbend >= u.cb0) {
        float v[R];
        uint32_t acc[R];
#pragma unroll
        for (int i = 0; i < R; ++i) v[i] = neg;
        float carry = (u.x0 == 0) ? 0.f : neg;              // frame 0 of token 0 starts from 0 (core.pyx:24-25)
        const uint32_t bnd_a = ptx::smem_u32(smem + g.off_bnd);
        const uint32_t stage_a = bnd_a + kBndBlocks * kBlk * 4;           // outgoing boundary scores, one slot per block of the ring
        const uint32_t bnd_bar_a = ptx::smem_u32(smem + g.off_bar) + (uint32_t)(g.nteams * 2 * kStages) * 8u;   // [kBndBlocks]
        // ---- as the producer of the next CTA's boundary ----
        const bool has_next = u.x0 + u.n_c < u.tx;          // the next CTA has real tokens
        const int pub_first = max(u.cb0, ((u.x0 + u.n_c) >> 5) - 1);
        const bool pub_lane = lane == u.n_c / R - 1;        // owns the slice's last token (n_c % R == 0)
        const uint32_t bnd_out_base = has_next ? ptx::mapa(bnd_a, (uint32_t)(u.c + 1)) : 0u;
        const uint32_t bar_out_base = has_next ? ptx::mapa(bnd_bar_a, (uint32_t)(u.c + 1)) : 0u;
        const volatile int *credit = ctl + kCredit;         // blocks the next CTA has consumed (set to pub_first before the sweep)
        int seen_credit = pub_first;
        // ---- as the consumer of the previous CTA's ----
        const int prev_x0 = u.x0 - u.n_c;
        const int prev_first = (u.c > 0) ? max(prev_x0 >> 5, u.cb0 - 1) : 0x3fffffff;
        const int prev_last = (u.c > 0) ? min(u.ty - 1, u.x0 - 1 + (u.ty - u.tx)) >> 5 : -1;
        const uint32_t credit_out = (u.c > 0) ? ptx::mapa(ptx::smem_u32(ctl + kCredit), (uint32_t)(u.c - 1)) : 0u;
        uint32_t phase_bits = 0u;                           // parity of every boundary slot's next phase
        const int row0 = u.x0 + lane * R;
        const int x_last = u.x0 + u.n_real - 1;
        uint32_t lane_c[R];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int row = min(lane * R + i, u.TR - 1);   // lanes beyond the tile re-read its last row (their tokens are inert)
            lane_c[i] = (uint32_t)(row * 128) | (row_key(row, R) << 4);
        }
        const uint32_t ring_a = ptx::smem_u32(smem + g.off_ring), box_bytes = (uint32_t)g.ring_rows * 128u;
        uint32_t *bits_s = reinterpret_cast<uint32_t *>(smem + g.off_bits);
        uint32_t *bits_p = (g.bits_in_smem ? bits_s : bits_g) + (size_t)u.cb0 * g.ring_rows + lane * R;
        int jn_team = 0, jn_need = g.team_warps, jn_end = u.f_lo + u.F;    // the chunk the sweep is waiting for
        int ready_end = u.f_lo;                                              // frames below are known to be stored
        int slot = 0;
        uint32_t spins = 0;
        // consume boundary block cb: wait for its bytes; afterwards (done) re-arm the slot and return the credit
        auto bnd_wait = [&](int cb) {
            const uint32_t s = (uint32_t)cb & (kBndBlocks - 1);
            while (!ptx::mbar_try_wait_a(bnd_bar_a + s * 8u, (phase_bits >> s) & 1u))
                if (++spins > kSpinLimit) systolic::spin_fail();
            phase_bits ^= 1u << s;
        };
        auto bnd_done = [&](int cb) {
            const uint32_t s = (uint32_t)cb & (kBndBlocks - 1);
            if (lane == 0) {
                ptx::mbar_arrive_expect_tx_a(bnd_bar_a + s * 8u, kBlk * 4);
                ptx::st_cluster_u32(credit_out, (uint32_t)(cb + 1));
            }
        };
        if (prev_first == u.cb0 - 1 && prev_first >= 0) {
            // score of token x0-1 at the last frame before my first block (when the previous CTA swept that
            // block; if its slice starts in my first block, that cell is below the diagonal: -1e9 already)
            bnd_wait(prev_first);
            carry = ptx::ld_shared_f32_a(bnd_a + (uint32_t)(((prev_first & (kBndBlocks - 1)) * kBlk + (kBlk - 1)) * 4));
            __syncwarp();
            bnd_done(prev_first);
        }
        for (int cb = u.cb0; cb <= u.cbend; ++cb) {
            const long long t0 = kDbg ? clock64() : 0;
            if (cb >= prev_first && cb <= prev_last) bnd_wait(cb);
            const long long t1 = kDbg ? clock64() : 0;
            const bool publishes = has_next && cb >= pub_first;
            if (publishes) {
                while (seen_credit + kBndBlocks <= cb) {    // the next CTA has consumed block cb - ring depth
                    seen_credit = *credit;
                    if (++spins > kSpinLimit) systolic::spin_fail();
                }
            }
            const long long t2 = kDbg ? clock64() : 0;
            const int last = min(cb * kBlk + kBlk, u.f_hi) - 1;
            if (last >= ready_end) {
                // this CTA's teams have stored the chunks up to the box's last frame (chunk jn ends at jn_end;
                // chunk j is stored when every warp of team j % nteams has counted j / nteams + 1 chunks)
                for (;;) {
                    while (ld_acquire_shared(ctl + kTeamDone + jn_team) < jn_need) {
                        __nanosleep(32);
                        if (++spins > kSpinLimit) systolic::spin_fail();
                    }
                    ready_end = jn_end;
                    if (last < jn_end) break;
                    jn_end += u.F;
                    if (++jn_team == g.nteams) {
                        jn_team = 0;
                        jn_need += g.team_warps;
                    }
                }
            }
            const long long t3 = kDbg ? clock64() : 0;
#pragma unroll
            for (int i = 0; i < R; ++i) acc[i] = 0u;
            const uint32_t tile_a = ring_a + (uint32_t)slot * box_bytes;
            const uint32_t ring_slot = (uint32_t)(cb & (kBndBlocks - 1)) * (kBlk * 4);
            const int col0 = cb * kBlk;
            const bool on_diagonal = col0 <= x_last;        // warp-uniform: some token of the slice is below the diagonal here
            if (on_diagonal) {
                zero_below_diagonal_rows<R>(tile_a, lane, row0, col0, u.TR);
                __syncwarp();
            }
            if (col0 + kBlk > u.f_hi) {
                // the band ends inside this box: nobody contracted the frames behind it, and whatever the
                // ring holds there must at least be finite (f_hi is a multiple of 8: whole 16-byte groups)
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    const int q = lane * R + i;
                    if (q >= u.TR) continue;
                    for (int cidx = (u.f_hi - col0) >> 2; cidx < 8; ++cidx)
                        ptx::st_shared_v4_if(true, tile_a + (uint32_t)q * 128u + (((uint32_t)cidx ^ row_key(q, R)) << 4), make_float4(0.f, 0.f, 0.f, 0.f));
                }
                __syncwarp();
            }
            // the publisher lane parks its last token's 32 scores in a local slot (predicated stores, nothing
            // that splits the unrolled block) and sends them to the next CTA as ONE bulk copy that
            // completes its bytes on the neighbour's mbarrier
            systolic::sweep_block_ahead<R, 0, (R <= 3 ? 3 : 2)>(tile_a, lane_c, v, acc, carry, bnd_a + ring_slot, stage_a + ring_slot,
                                                               publishes && pub_lane, 0u);
            if (publishes && pub_lane) {
                ptx::fence_proxy_async();
                ptx::bulk_copy_s2peer(bnd_out_base + ring_slot, stage_a + ring_slot, kBlk * 4,
                                      bar_out_base + (uint32_t)(cb & (kBndBlocks - 1)) * 8u);
            }
            const long long t4 = kDbg ? clock64() : 0;
#pragma unroll
            for (int i = 0; i < R; ++i) acc[i] = __brev(acc[i]);
            if (on_diagonal) {
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    // the forced step on the diagonal (frame == token, core.pyx:34), tokens > 0 only
                    const int d = row0 + i - col0;
                    if (d >= 0 && d < kBlk && row0 + i > 0) acc[i] |= 1u << d;
                }
            }
#pragma unroll
            for (int i = 0; i < R; ++i)
                if (lane * R + i < g.ring_rows) bits_p[i] = acc[i];
            bits_p += g.ring_rows;
            __syncwarp();                                   // every lane has read the box and the boundary slot
            const long long t5 = kDbg ? clock64() : 0;
            if (cb >= prev_first && cb <= prev_last) bnd_done(cb);
            const long long t6 = kDbg ? clock64() : 0;
            if (lane == 0) st_release_shared(ctl + kConsumed, cb - u.cb0 + 1);   // frees the box; the map builders read the bits after it
            if (++slot == g.NB) slot = 0;
            if (kDbg) {
                t_prev += t1 - t0;
                t_credit += t2 - t1;
                t_chunks += t3 - t2;
                t_core += t4 - t3;
                t_bits += t5 - t4;
                t_done += t6 - t5;
                t_fill += clock64() - t6;
                if (cb > u.cbend - 8) {                     // the last eight blocks: the SM's FFMA warps are done by then
                    t_core_tail += t4 - t3;
                    t_all_tail += clock64() - t0;
                }
            }
        }
        // a NaN or an infinity anywhere in a token's history is still in its score now
#pragma unroll
        for (int i = 0; i < R; ++i)
            if (lane * R + i < u.n_real && !(fabsf(v[i]) <= 3.402823466e38f)) nonfinite = 1;
    }
    if (kDbg && dbg && lane == 0) {
        dbg[4] = ptx::globaltimer_ns();
        dbg[16] = t_chunks, dbg[17] = t_prev, dbg[18] = t_credit, dbg[19] = t_core;
        dbg[20] = clock64() - t_begin, dbg[21] = u.cbend - u.cb0 + 1;
        dbg[22] = t_bits, dbg[23] = t_done, dbg[24] = t_fill;
        dbg[28] = t_core_tail, dbg[29] = t_all_tail;
    }
    if (lane == 0) {
        if (zf.total > 0) {
            ptx::bulk_commit_group();
            ptx::bulk_wait_all();                          // the ones are written after the next barriers
        }
        if (kDbg && dbg) dbg[14] = ptx::globaltimer_ns();
    }
    __syncwarp();
    return nonfinite;
}

// ---------------------------------------------------------------------------------------------
// backtrack (core.pyx:32-35), in parallel over 32-frame blocks
//
// The path sits on token x for frames (.., y]; the frame where it stepped onto x is the highest set
// direction bit of x at or below y (mas_dp_cta.cuh: backtrack_tokens).  Inside one block that is a
// short walk (the path descends a handful of tokens per 32 frames); what makes the whole backtrack
// serial is only that a block's walk must know on which token the path LEAVES the block above.  So:
//   * while the sweep is still running, the other warps tabulate for every finished block and every
//     token of the slice "at the block's last frame on token x -> at the previous block's last
//     frame on token x - m" (BLOCK MAP, one byte per token and block; 255 = the path leaves the slice
//     inside the block);
//   * after the sweep one thread per CTA composes the maps from its entry point downwards -- one
//     dependent shared-memory load per block instead of a walk -- until the path leaves the slice,
//     walks that one block to find the exact frame, and hands (token, frame) to the CTA below;
//   * then one lane per block walks its block from the now known token and records for every token
//     the frame where the path stepped onto it (`ylo`); a token's run ends one frame before the
//     next token's begins.
// ---------------------------------------------------------------------------------------------
constexpr int kLeft = 255;

// The path is on local token xl at local frame y (0..31) of the block whose direction words are
// `row` ([tokens of the slice]); col0 = the block's first frame, x0 = the slice's first token.
// Returns the local token at the PREVIOUS block's last frame, -1 when the path leaves the slice
// (then ylo[0] is the frame where it stepped onto the slice's first token).
template <bool kSmem, bool kRecord>
__device__ __forceinline__ int walk_block(const uint32_t *row, int xl, int y, int x0, int col0, int *ylo) {
    while (xl >= 0 && x0 + xl > 0) {                        // token 0 is never left (core.pyx:34 `index != 0`)
        const uint32_t w = (kSmem ? row[xl] : __ldcg(row + xl)) & (0xffffffffu >> (31 - y));
        if (w == 0u) break;                                 // on this token since before the block
        const int lo = 31 - __clz(w);                       // stepped onto it here
        if (kRecord) ylo[xl] = col0 + lo;
        --xl;
        if (lo == 0) break;                                 // ... from the previous block's last frame
        y = lo - 1;
    }
    return xl;
}

// The same walk by a whole warp (every lane passes the same arguments and gets the same result): the
// direction words of the 32 tokens below the start come in with ONE load, lane i holding token
// xl - i, and the chain runs on shuffles -- ~30 cycles per step instead of a dependent shared-memory
// round trip for a lone thread (~120).  For the walks ON the backtrack's critical path.
template <bool kSmem, bool kRecord>
__device__ __forceinline__ int walk_block_warp(const uint32_t *row, int xl, int y, int x0, int col0, int *ylo, int lane) {
    int base = xl, i = 0;
    uint32_t mine = (base - lane >= 0) ? (kSmem ? row[base - lane] : __ldcg(row + base - lane)) : 0u;
    while (xl >= 0 && x0 + xl > 0) {                        // token 0 is never left (core.pyx:34 `index != 0`)
        if (i == 32) {
            base = xl, i = 0;
            mine = (base - lane >= 0) ? (kSmem ? row[base - lane] : __ldcg(row + base - lane)) : 0u;
        }
        const uint32_t w = __shfl_sync(0xffffffffu, mine, i) & (0xffffffffu >> (31 - y));
        if (w == 0u) break;                                 // on this token since before the block
        const int lo = 31 - __clz(w);                       // stepped onto it here
        if (kRecord && lane == 0) ylo[xl] = col0 + lo;
        --xl, ++i;
        if (lo == 0) break;                                 // ... from the previous block's last frame
        y = lo - 1;
    }
    return xl;
}

// walk_block that also reports where the path stepped onto the slice's FIRST token (y0), if it did.
template <bool kSmem>
__device__ __forceinline__ int walk_block_exit(const uint32_t *row, int xl, int y, int x0, int col0, int &y0) {
    while (xl >= 0 && x0 + xl > 0) {
        const uint32_t w = (kSmem ? row[xl] : __ldcg(row + xl)) & (0xffffffffu >> (31 - y));
        if (w == 0u) break;
        const int lo = 31 - __clz(w);
        if (xl == 0) y0 = col0 + lo;
        --xl;
        if (lo == 0) break;
        y = lo - 1;
    }
    return xl;
}

// Block maps of this CTA's slice as the sweep finishes the blocks.  A unit is (block, 32 tokens); the
// builder warps take units first, first + step, ...: the groups of one block go to different warps, so
// that the maps of the slice's last blocks -- the only ones on the critical path, between the end of the
// sweep and the backtrack -- are there one walk after the sweep, not one walk per group.
//
// EXIT TABLE (`exits` != nullptr: CTAs 1 .. K-1): exits[block][token] = the frame where the path steps onto
// the slice's first token when it is on `token` at the block's last frame -- the composition of all the
// maps below, tabulated: exits[b][x] = exits[b-1][x - map[b][x]], or the frame the walk itself found when
// the path leaves the slice inside block b.  A row needs the row below complete (`erow` counts a row's
// finished groups; units are taken in ascending order, so the wait is always for a unit that is done or
// in progress).  The backtrack's hop through this CTA is then the entry block's walk and ONE look-up
// before the hand-over to the CTA below, instead of a dependent load per block in between.
template <bool kSmem>
__device__ __forceinline__ void build_block_maps(const Geom &g, const Utt &u, const uint32_t *bits, unsigned char *maps,
                                                 unsigned short *exits, int *erow, const int *ctl, int first, int step, int lane) {
    const int nbox = u.cbend - u.cb0;                       // (not the last block: a path only ever ENTERS a slice there)
    const int ngrp = ceil_div(u.n_real, 32);
    uint32_t spins = 0;
    for (int unit = first; unit < nbox * ngrp; unit += step) {
        const int cbl = unit / ngrp, xl = (unit - cbl * ngrp) * 32 + lane;
        while (ld_acquire_shared(ctl + kConsumed) <= cbl) {
            __nanosleep(cbl + 3 >= nbox ? 50 : 400);
            if (++spins > kSpinLimit) systolic::spin_fail();
        }
        int r = -1, y0 = 0;
        if (xl < u.n_real) {
            const uint32_t *row = bits + (size_t)(u.cb0 + cbl) * g.ring_rows;
            r = walk_block_exit<kSmem>(row, xl, 31, u.x0, (u.cb0 + cbl) * kBlk, y0);
            maps[(size_t)cbl * g.ring_rows + xl] = (unsigned char)(r < 0 ? kLeft : xl - r);
        }
        if (exits != nullptr) {
            if (cbl > 0)
                while (ld_acquire_shared(erow + cbl - 1) < ngrp) {
                    __nanosleep(20);
                    if (++spins > kSpinLimit) systolic::spin_fail();
                }
            // (in the slice's first block every walk leaves the slice: the path is on a token <= its frame)
            if (xl < u.n_real) exits[cbl * g.ring_rows + xl] = (unsigned short)(r < 0 ? y0 : exits[(cbl - 1) * g.ring_rows + r]);
            __syncwarp();
            if (lane == 0) ptx::red_release_shared_add(erow + cbl, 1);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// the literal redo of an utterance with non-finite scores (CTA 0 of the cluster, all its threads)
// ---------------------------------------------------------------------------------------------
static __device__ __forceinline__ void redo_utterance(const Geom &g, const Utt &u, const LogpParams &p, unsigned char *smem,
                                                   float *scratch, uint32_t *bits, float neg) {
    using namespace logp;
    const TileShape &t = g.t_ref;
    const int D = p.D, F = t.F, tx = u.tx, ty = u.ty;
    // the contraction's operands as mas_logp_cta.cuh lays them out, with ONE frame buffer (nothing is
    // prefetched here), then the exact sweep's two score columns
    float *sm = reinterpret_cast<float *>(smem + g.off_big);
    CtaSmem s;
    s.sInv = sm;
    s.sMiv = s.sInv + D * t.tile_rows;
    s.sZ = s.sMiv + D * t.tile_rows;
    s.sL1 = s.sZ + D * F;
    s.sL4 = s.sL1 + t.tile_rows;
    s.sL2 = s.sL4 + t.tile_rows;
    s.sPart = s.sL2 + F;
    float *col = s.sPart + 8 * t.tile_rows;                 // [2][tx]
    const systolic::Team team = systolic::whole_cta();
    const systolic::ExactBits eb{0u, bits, u.n_c, g.nblk};
    const int rts = ceil_div(tx, t.tile_rows);
    const ChunkOut o{scratch, F, 0, true};
    int buf = 0;
    __syncthreads();                                        // the fast path's shared memory changes hands
    systolic::exact_sweep_init(team, col, tx, neg);
    for (int ch = 0; ch * F < ty; ++ch) {
        for (int rt = 0; rt < rts; ++rt) {
            if (rts > 1 || ch == 0) {
                __syncthreads();
                stage_tokens(p, s, t, u.b, rt * t.tile_rows);
            }
            if (rt == 0) {
                stage_frames_async(p, s, t, u.b, ch, 0);
                ptx::cp_async_wait<0>();
                __syncthreads();
                if (p.x_logs == nullptr) frame_sums_mean_only(s, D, F, 0);
            }
            contract_chunk(p, s, t, rt * t.tile_rows, ch, 0, 2, o);
        }
        __syncthreads();                                    // the chunk's scores are in the scratch (same CTA: visible)
        systolic::exact_sweep_frames<true>(team, scratch - ch * F, F, col, buf, eb, tx, ch * F, min(ty, ch * F + F), neg);
    }
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <int R, bool kDbg>
__global__ void __launch_bounds__(kThreads, 1)
mas_fused_kernel(const __grid_constant__ CUtensorMap tmap_z, PathParams pp, LogpParams lp, Geom g) {
    extern __shared__ __align__(1024) unsigned char smem[];
    ptx::grid_launch_dependents();      // (a following instance may be scheduled from now on; it waits before it touches global memory)
    const int tid = threadIdx.x, lane = tid & 31;
    // broadcast so that the compiler knows the warp index is warp-uniform (see dp_cta)
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int K = g.K, c = (int)ptx::cluster_ctarank();
    const int cluster_id = (int)blockIdx.x / K;
    const int B = pp.B, T_x = pp.T_x, T_y = pp.T_y;
    int *ctl = reinterpret_cast<int *>(smem + g.off_ctl);
    volatile int *vctl = ctl;
    float *bnd = reinterpret_cast<float *>(smem + g.off_bnd);
    int2 *run = reinterpret_cast<int2 *>(smem + g.off_run);       // redo path: [first frame, last frame] per token
    int *ylo = reinterpret_cast<int *>(smem + g.off_run);         // frame where the path steps onto each token (+ one past the top)
    int *xend = reinterpret_cast<int *>(smem + g.off_xend);       // token at the last frame of each block of the slice (-1: not composed)
    uint32_t *bits_s = reinterpret_cast<uint32_t *>(smem + g.off_bits);
    unsigned char *maps_s = smem + g.off_maps;
    unsigned short *exits_s = reinterpret_cast<unsigned short *>(smem + g.off_exit);
    int *erow = reinterpret_cast<int *>(smem + g.off_erow);
    const float neg = pp.max_neg_val;

    // FFMA role: every warp but the sweep warp, or only those on the other three schedulers
    // (a warp lives on scheduler warp % 4)
    const bool is_ffma = warp != 0 && (g.ffma_all || (warp & 3) != 0);
    const int fw = g.ffma_all ? warp - 1 : (warp - 1) - (warp >> 2);   // rank among the FFMA warps
    const int fidx = fw * 32 + lane;
    const int team = fw / g.team_warps, tn = g.team_warps * 32, ttid = fidx - team * tn;
    const bool in_team = is_ffma && team < g.nteams;

    uint64_t *zbars = reinterpret_cast<uint64_t *>(smem + g.off_bar);          // per team: full[kStages], empty[kStages]
    uint64_t *bnd_bars = zbars + g.nteams * 2 * kStages;                       // [kBndBlocks]
    {   // once per CTA: the zero page; what "advances" into token 0 after frame 0 (core.pyx:26-27) for
        // CTA 0, finite values elsewhere (a slot may be read after the previous CTA's boundary token has
        // left the band); the z pipelines' barriers
        float4 *zero4 = reinterpret_cast<float4 *>(smem + g.off_zero);
        for (int i = tid; i < kZeroPage / 16; i += kThreads) zero4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = tid; i < kBndBlocks * kBlk; i += kThreads) bnd[i] = neg;
        if (tid == 0) {
            for (int t = 0; t < g.nteams; ++t)
                for (int st = 0; st < kStages; ++st) {
                    ptx::mbar_init(zbars + t * 2 * kStages + st, 1);
                    ptx::mbar_init(zbars + t * 2 * kStages + kStages + st, (uint32_t)g.team_warps);
                }
            ptx::fence_barrier_init();
            ptx::prefetch_tensormap(&tmap_z);
        }
        ptx::fence_proxy_async();
        __syncthreads();
    }
    // Programmatic dependent launch: when the kernel BEFORE this one in the stream is another instance of
    // this launch (back-to-back steps), it lets this grid's CTAs take the SMs its own CTAs leave -- they get
    // as far as here (shared memory only) and then wait for that grid to complete and flush.  After any
    // other kernel the wait returns at once (the launch was serialised the ordinary way).
    ptx::grid_dependency_wait();
    long long *dbg = (kDbg && pp.dbg_cycles) ? pp.dbg_cycles + (size_t)blockIdx.x * 32 : nullptr;
    int zseq = 0;                       // z panels this thread's team has consumed so far (mbarrier phases run on)

    for (int it = 0;; ++it) {
        // static, serpentine: batches come sorted by length (dataset.py:79-81), so consecutive rounds
        // in alternating directions even out the clusters' totals
        const int b = it * g.NC + ((it & 1) ? g.NC - 1 - cluster_id : cluster_id);
        if (b >= B) break;
        Utt u;
        u.b = b;
        u.c = c;
        {
            const Lengths len = clamp_lengths(lp.x_len[b], lp.y_len[b], T_x, T_y);
            u.tx = len.tx;
            u.ty = len.ty;
        }
        u.n_c = max(R, ceil_div(ceil_div(u.tx, K), R) * R);
        u.x0 = c * u.n_c;
        u.n_real = max(0, min(u.n_c, u.tx - u.x0));
        u.TR = ceil_div(u.n_c, kGemmTM) * kGemmTM;
        u.RG = u.TR / kGemmTM;
        u.CG = max(1, min(tn / u.RG, g.CG_cap));
        u.F = 8 * u.CG;
        u.cb0 = u.x0 >> 5;
        {
            // the slice is in the band from frame x0 (core.pyx:18: x <= y) to where its last token leaves it
            const int y_last = min(u.ty - 1, u.x0 + u.n_real - 1 + (u.ty - u.tx));
            u.cbend = u.n_real > 0 ? y_last >> 5 : u.cb0 - 1;
            u.f_lo = u.x0 & ~7;
            u.f_hi = u.n_real > 0 ? min((y_last + 8) & ~7, (u.cbend + 1) * kBlk) : u.f_lo;
            u.nch = ceil_div(u.f_hi - u.f_lo, u.F);
        }
        const bool active = u.n_real > 0;

        if (tid == 0) {
            for (int i = 0; i < kMaxTeams; ++i) ctl[kTeamDone + i] = 0;
            ctl[kConsumed] = 0;
            ctl[kCredit] = max(u.cb0, ((u.x0 + u.n_c) >> 5) - 1);     // = the first block this sweep publishes
            ctl[kRedo] = 0;
            // the boundary ring's barriers start every utterance afresh; a consuming sweep arms them all
            for (int i = 0; i < kBndBlocks; ++i) {
                if (it > 0) ptx::mbar_inval(bnd_bars + i);
                ptx::mbar_init(bnd_bars + i, 1);
            }
            ptx::fence_barrier_init();
            if (c > 0 && active)
                for (int i = 0; i < kBndBlocks; ++i) ptx::mbar_arrive_expect_tx(bnd_bars + i, kBlk * 4);
            // the backtrack's hand-over from the CTA above: 8 bytes (token, frame)
            if (it > 0) ptx::mbar_inval(bnd_bars + kBndBlocks);
            ptx::mbar_init(bnd_bars + kBndBlocks, 1);
            ptx::fence_barrier_init();
            ptx::mbar_arrive_expect_tx(bnd_bars + kBndBlocks, 8);
        }
        for (int i = tid; i <= u.cbend - u.cb0; i += kThreads) {
            xend[i] = -1;
            if (g.use_exit) erow[i] = 0;
        }
        if (kDbg && dbg && tid == 0) dbg[0] = ptx::globaltimer_ns();
        if (in_team && active) {
            // the first z panels and the raw token-side values: issued before the barrier, they land meanwhile
            team_contract<false>(tmap_z, g, u, lp, smem, ctl, team, ttid, tn, zseq, true);
            stage_ops<true>(g, u, lp, smem, fidx, g.nteams * tn, nullptr);
        }
        ptx::cluster_sync();

        int nonfinite = 0;
        if (warp == 0) {
            ZeroFill zf;
            const int rs = ceil_div(T_x, K), r0 = min(T_x, c * rs), r1 = min(T_x, r0 + rs);
            zf.dst = reinterpret_cast<char *>(pp.path + ((int64_t)b * T_x + r0) * T_y);
            zf.total = (int64_t)(r1 - r0) * T_y * 4;        // multiple of 16: T_y % 4 == 0 on this path
            zf.off = 0;
            zf.per_block = ceil_div((int)((zf.total + kZeroPage - 1) / kZeroPage), max(1, u.cbend - u.cb0 + 1));
            uint32_t *bits_g = g.bits_in_smem ? nullptr : g.ws_bits + ((size_t)cluster_id * K + c) * g.nblk * g.ring_rows;
            // the dense output's zeros: all bulk copies now, while the sweep has nothing to do yet (the first
            // chunk of scores is a whole contraction pass away); they drain in the background
            if (lane == 0) zf.drip(smem + g.off_zero, 0x7fffffff);
            if (active) {
                named_sync(kBarAll, g.nteams * tn + 32);    // stage_ops has parked the raw token-side values
                sum_row_constants(g, u, lp, smem, lane);
                if (kDbg && dbg && lane == 0) dbg[1] = ptx::globaltimer_ns();
                named_sync(kBarAll, g.nteams * tn + 32);    // (the teams wait here before their first store)
            }
            nonfinite = sweep_slice<R, kDbg>(g, u, smem, ctl, neg, zf, bits_g, dbg);
        } else if (in_team && active) {
            stage_ops<false>(g, u, lp, smem, fidx, g.nteams * tn, kDbg ? dbg : nullptr);
            if (lp.x_logs == nullptr)
                zseq = team_contract<true>(tmap_z, g, u, lp, smem, ctl, team, ttid, tn, zseq, false);
            else
                zseq = team_contract<false>(tmap_z, g, u, lp, smem, ctl, team, ttid, tn, zseq, false);
            if (kDbg && dbg && ttid == 0) dbg[8 + team] = ptx::globaltimer_ns();
        }
        uint32_t *const bits_w = g.bits_in_smem ? nullptr : g.ws_bits + ((size_t)cluster_id * K + c) * g.nblk * g.ring_rows;
        unsigned char *const maps = g.bits_in_smem ? maps_s : g.ws_maps + ((size_t)cluster_id * K + c) * g.nblk * g.ring_rows;
        if (g.use_maps && (warp & 3) != 0 && active) {
            // the warps on the other three schedulers, once their contraction is done: the backtrack's block maps
            const int bw = (warp - 1) - (warp >> 2);
            if (g.bits_in_smem)
                build_block_maps<true>(g, u, bits_s, maps, (g.use_exit && c > 0) ? exits_s : nullptr, erow, ctl, bw, 12, lane);
            else
                build_block_maps<false>(g, u, bits_w, maps, nullptr, erow, ctl, bw, 12, lane);
            if (kDbg && dbg && warp == 1 && lane == 0) dbg[3] = ptx::globaltimer_ns();
            if (kDbg && dbg && lane == 0) atomicMax(reinterpret_cast<unsigned long long *>(dbg + 15), (unsigned long long)ptx::globaltimer_ns());
        }
        if (!g.bits_in_smem) __threadfence();

        // ---- were all scores finite?  (cluster-wide) ----
        const int any_bad = __syncthreads_or(nonfinite);
        if (kDbg && dbg && tid == 0) dbg[2] = ptx::globaltimer_ns();
        // The only thing this barrier carries across the cluster is the redo flag, so only a CTA that raises
        // it pays for a release (a fence at cluster scope is a MEMBAR.ALL.GPU: ~1 us on the critical path
        // between the last sweep and the backtrack); everybody else arrives relaxed.
        if (any_bad && tid == 0) {
            for (int r = 0; r < K; ++r) ptx::st_cluster_u32(ptx::mapa(ptx::smem_u32(ctl + kRedo), (uint32_t)r), 1u);
            ptx::fence_acq_rel_cluster();
        }
        ptx::cluster_sync_arrive_relaxed();
        const bool redo = vctl[kRedo] != 0;
        const uint32_t *bits_gl = g.bits_in_smem ? nullptr : g.ws_bits + ((size_t)cluster_id * K + c) * g.nblk * g.ring_rows;
        int bits_rows = g.ring_rows;
        if (redo) {
            uint32_t *rb = g.redo_bits + (size_t)cluster_id * g.nblk * (T_x + 64);
            if (c == 0) {
                redo_utterance(g, u, lp, smem, g.redo_scratch + (size_t)cluster_id * T_x * g.t_ref.F, rb, neg);
                __threadfence();
            }
            ptx::cluster_sync();
            bits_gl = rb + (size_t)c * g.nblk * u.n_c;
            bits_rows = u.n_c;
        }
        if (kDbg && dbg && tid == 0) dbg[5] = ptx::globaltimer_ns();

        // ---- backtrack (core.pyx:32-35), handed down from CTA to CTA ----
        const int c_last = (u.tx > 0) ? (u.tx - 1) / u.n_c : -1;   // CTA that owns the last token
        const uint32_t bt_bar_a = ptx::smem_u32(bnd_bars + kBndBlocks);
        float *out = pp.path + (int64_t)b * T_x * T_y;
        if (redo || !g.use_maps) {
            // the token walk of kernel (1), CTA after CTA: for long slices (tabulating the block maps costs
            // the builders more than the maps save: 39 us of maps for an 11 us backtrack at 400 x 2000), and
            // after a redo (rare; its direction words are in global memory, tokens strided by the slice)
            if (warp == 0 && c <= c_last) {
                int x, y_hi;
                if (c == c_last) {
                    x = u.tx - 1;
                    y_hi = u.ty - 1;
                } else {
                    uint32_t spins = 0;
                    while (!ptx::mbar_try_wait_a(bt_bar_a, 0u))
                        if (++spins > kSpinLimit) systolic::spin_fail();
                    x = vctl[kBtToken];
                    y_hi = vctl[kBtFrame];
                }
                const int x_min = max(u.x0, 1);
                if (x >= x_min) {
                    // (shared memory: the lone thread's tuned walk; workspace: one coalesced load per block, mas_dp_cta.cuh)
                    if (bits_gl != nullptr)
                        y_hi = systolic::backtrack_tokens_warp<false>(bits_gl, bits_rows, u.x0, x, y_hi, x_min, run, lane);
                    else if (lane == 0)
                        y_hi = systolic::backtrack_tokens<true>(bits_s, bits_rows, u.x0, x, y_hi, x_min, run);
                    y_hi = __shfl_sync(0xffffffffu, y_hi, 0);
                }
                __syncwarp();
                if (lane == 0) {
                    if (c == 0) {
                        run[0] = make_int2(0, y_hi);
                    } else {
                        const uint32_t peer = ptx::mapa(ptx::smem_u32(ctl + kBtToken), (uint32_t)(c - 1));   // (8-byte aligned)
                        ptx::st_async_b64(peer, (uint64_t)(uint32_t)(u.x0 - 1) | ((uint64_t)(uint32_t)y_hi << 32),
                                          ptx::mapa(bt_bar_a, (uint32_t)(c - 1)));
                    }
                }
            }
            __syncthreads();
            for (int xl = tid; xl < u.n_real; xl += kThreads) {
                const int x = u.x0 + xl;
                const int2 r = run[xl];
                float *row = out + (int64_t)x * T_y;
                for (int y = r.x; y <= r.y; ++y) row[y] = 1.f;
                if (pp.frame_token)
                    for (int y = r.x; y <= r.y; ++y) pp.frame_token[(int64_t)b * T_y + y] = x;
                if (pp.durations) pp.durations[(int64_t)b * T_x + x] = r.y - r.x + 1;
            }
        } else {
            // (one instantiation per place the direction words and maps live in: with the pointers selected at
            // run time every load of the composition and of the walks was a generic LD -- ~260 cycles per
            // composed block on the backtrack's critical path instead of a shared-memory load's ~50)
            auto backtrack_by_maps = [&](auto in_smem) {
                constexpr bool kS = decltype(in_smem)::value;
                const uint32_t *bits_b = kS ? bits_s : bits_w;
                const unsigned char *maps_b = kS ? maps_s : maps;
                const uint32_t maps_a = kS ? ptx::smem_u32(maps_s) : 0u;
                if (warp == 0 && c <= c_last) {
                    // (the whole sweep warp, uniformly) compose the block maps from the entry point down; walk only
                    // the entry block and the block where the path leaves the slice
                    int xl, y;
                    if (c == c_last) {
                        xl = u.tx - 1 - u.x0;
                        y = u.ty - 1;
                    } else {
                        uint32_t spins = 0;
                        while (!ptx::mbar_try_wait_a(bt_bar_a, 0u))
                            if (++spins > kSpinLimit) systolic::spin_fail();
                        xl = vctl[kBtToken] - u.x0;
                        y = vctl[kBtFrame];
                    }
                    const long long tb0 = kDbg ? clock64() : 0;
                    if (lane == 0) ylo[xl + 1] = y + 1;         // where the token above begins
                    int cb = y >> 5;
                    xl = walk_block_warp<kS, true>(bits_b + (size_t)cb * g.ring_rows, xl, y & 31, u.x0, cb * kBlk, ylo, lane);
                    const long long tb1 = kDbg ? clock64() : 0;
                    // the hand-over to the CTA below, as early as it is known: (token, frame) just before the path
                    // steps onto this slice's first token
                    const uint32_t peer_tok = c > 0 ? ptx::mapa(ptx::smem_u32(ctl + kBtToken), (uint32_t)(c - 1)) : 0u;   // (8-byte aligned)
                    const uint32_t peer_bar = c > 0 ? ptx::mapa(bt_bar_a, (uint32_t)(c - 1)) : 0u;
                    bool sent = false;
                    if (c > 0 && lane == 0) {
                        int y0 = -1;
                        if (xl < 0)
                            y0 = ylo[0];                        // it left the slice inside the entry block
                        else if (kS && g.use_exit && cb - 1 >= u.cb0)
                            y0 = exits_s[(cb - 1 - u.cb0) * g.ring_rows + xl];
                        if (y0 >= 0) {
                            ptx::st_async_b64(peer_tok, (uint64_t)(uint32_t)(u.x0 - 1) | ((uint64_t)(uint32_t)(y0 - 1) << 32), peer_bar);
                            sent = true;
                        }
                    }
                    if (kDbg && dbg && lane == 0) dbg[13] = ptx::globaltimer_ns();
                    int nsteps = 0;
                    int row_off = (cb - u.cb0) * g.ring_rows;
                    while (xl >= 0 && --cb >= u.cb0) {
                        ++nsteps;
                        row_off -= g.ring_rows;
                        // (every map below a slice's last block is complete: its builders passed the CTA barrier above;
                        // the last block is only ever an entry block, walked directly, and nobody builds its map)
                        const int m = kS ? (int)ptx::ld_shared_u8_a(maps_a + (uint32_t)(row_off + xl)) : (int)maps_b[row_off + xl];
                        if (m == kLeft) {                       // the path leaves the slice here
                            xl = walk_block_warp<kS, true>(bits_b + (size_t)cb * g.ring_rows, xl, 31, u.x0, cb * kBlk, ylo, lane);
                            continue;
                        }
                        if (lane == 0) xend[cb - u.cb0] = xl;
                        xl -= m;
                    }
                    __syncwarp();
                    if (kDbg && dbg && lane == 0) dbg[30] = tb1 - tb0, dbg[31] = clock64() - tb1, dbg[25] = nsteps;
                    if (lane == 0) {
                        if (c == 0)
                            ylo[0] = 0;
                        else if (!sent)
                            ptx::st_async_b64(peer_tok, (uint64_t)(uint32_t)(u.x0 - 1) | ((uint64_t)(uint32_t)(ylo[0] - 1) << 32), peer_bar);
                    }
                }
                __syncthreads();
                if (kDbg && dbg && tid == 0) dbg[6] = ptx::globaltimer_ns();
                // one warp per composed block: the frames where the path steps onto the tokens it visits there
                if (c <= c_last) {
                    for (int cbl = warp; cbl <= u.cbend - u.cb0; cbl += kThreads / 32) {
                        const int xe = xend[cbl];
                        if (xe < 0) continue;
                        walk_block_warp<kS, true>(bits_b + (size_t)(u.cb0 + cbl) * g.ring_rows, xe, 31, u.x0, (u.cb0 + cbl) * kBlk, ylo, lane);
                    }
                }
            };
            if (g.bits_in_smem)
                backtrack_by_maps(std::true_type{});
            else
                backtrack_by_maps(std::false_type{});
            __syncthreads();
            // ---- dense path: ones, durations, frame -> token ----
            for (int xl = tid; xl < u.n_real; xl += kThreads) {
                const int x = u.x0 + xl;
                const int y0 = ylo[xl], y1 = ylo[xl + 1];
                float *row = out + (int64_t)x * T_y;
                for (int y = y0; y < y1; ++y) row[y] = 1.f;
                if (pp.frame_token)
                    for (int y = y0; y < y1; ++y) pp.frame_token[(int64_t)b * T_y + y] = x;
                if (pp.durations) pp.durations[(int64_t)b * T_x + x] = y1 - y0;
            }
        }
        if (pp.durations)
            for (int x = u.tx + c * kThreads + tid; x < T_x; x += K * kThreads) pp.durations[(int64_t)b * T_x + x] = 0;
        if (pp.frame_token && c == 0)
            for (int y = u.ty + tid; y < T_y; y += kThreads) pp.frame_token[(int64_t)b * T_y + y] = -1;
        if (kDbg && dbg && tid == 0) dbg[7] = ptx::globaltimer_ns();
        // nobody may reset its control words (next utterance) or leave while a neighbour can still write them
        ptx::cluster_sync();
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int tokens_per_lane(int T_x, int K) {
    int R = ceil_div(ceil_div(T_x, K), kBlk);
    if (R < 1) R = 1;
    if (R == 7) R = 8;
    return R;
}

// Shared-memory layout and team shape for K CTAs per utterance; false when it does not fit.
// (the contraction's frames per utterance, for the estimates: the band of a full-length slice)
static int slice_frames(int T_x, int T_y, int max_slice) { return (T_y > T_x ? T_y - T_x : 0) + max_slice + 8; }

static bool layout_geom(int D, int T_x, int max_smem, Geom &g) {
    int off = 0;
    g.off_zero = off, off += kZeroPage;
    g.off_bnd = off, off += 2 * kBndBlocks * kBlk * 4;      // incoming ring + outgoing staging slots
    g.off_run = off, off += (g.ring_rows + 2) * 8;
    g.off_xend = off, off += g.nblk * 4;
    off = (int)align_up((size_t)off, 16);
    g.off_ctl = off, off += kCtlInts * 4;
    g.off_bar = off, off += (g.nteams * 2 * kStages + kBndBlocks + 1) * 8;
    off = (int)align_up((size_t)off, 1024);
    g.off_big = off;
    g.off_ops = off, off += 2 * D * g.tr_max * 4;
    g.off_l14 = off, off += 2 * g.tr_max * 4;
    g.off_part = off, off += 8 * g.tr_max * 4;
    off = (int)align_up((size_t)off, 128);
    g.off_z = off, off += g.nteams * kStages * kChan * g.F_cap * 4;   // TMA destinations: 128-byte aligned (F_cap is even)
    g.off_l2 = off, off += g.nteams * 2 * g.F_cap * 4;
    off = (int)align_up((size_t)off, 1024);
    g.off_ring = off;
    const int box = g.ring_rows * 128;
    // the ring must hold the chunk being stored and the box being swept; with every team's chunk in
    // flight it never stalls them; stage_ops parks the raw token-side values in it (2 arrays at least)
    int nb_floor = ceil_div(g.F_cap + 2 * kBlk, kBlk);
    if (nb_floor * box < 2 * D * g.tr_max * 4) nb_floor = ceil_div(2 * D * g.tr_max * 4, box);
    static const char *maps_env = getenv("MAS_B200_FUSED_MAPS");           // experiment hooks
    static const char *exit_env = getenv("MAS_B200_FUSED_EXIT");
    const bool want_maps = maps_env ? atoi(maps_env) != 0 : g.max_slice <= 64;
    const bool want_exit = want_maps && g.nblk * kBlk < 65536 && (exit_env ? atoi(exit_env) != 0 : true);
    // direction words + one map byte per token and block (+ the exit table: two bytes each, and a counter per block)
    const int bits_bytes = g.nblk * g.ring_rows * 5 + (want_exit ? g.nblk * g.ring_rows * 2 + g.nblk * 4 : 0);
    const int redo_need = g.off_big + (logp::cta_smem_floats(D, g.t_ref) - D * g.t_ref.F) * 4 + 2 * T_x * 4 + 16;   // (redo_utterance)
    if (redo_need > max_smem) return false;
    for (int bits_smem = 1; bits_smem >= 0; --bits_smem) {
        const int left = max_smem - g.off_ring - (bits_smem ? bits_bytes : 0);
        int nb = left / box;
        if (nb > 32) nb = 32;
        // the direction bits and block maps stay in shared memory as long as the ring still holds two
        // chunks and the box being swept: in the workspace every step of the map builders and of the
        // backtrack is an L2 round trip (400 x 2000: 170 us per utterance against 20)
        const int nb_bits = ceil_div(2 * g.F_cap + 2 * kBlk, kBlk) > nb_floor ? ceil_div(2 * g.F_cap + 2 * kBlk, kBlk) : nb_floor;
        if (nb >= (bits_smem ? nb_bits : nb_floor)) {
            g.NB = nb;
            g.bits_in_smem = bits_smem;
            g.ops_tmp3 = (int64_t)nb * box >= (int64_t)3 * D * g.tr_max * 4;
            g.use_maps = want_maps && bits_smem;
            g.use_exit = want_exit && bits_smem;
            g.off_bits = g.off_ring + nb * box;
            g.off_maps = g.off_bits + g.nblk * g.ring_rows * 4;
            g.off_exit = g.off_maps + g.nblk * g.ring_rows;        // (ring_rows is a multiple of 8: 2-byte aligned, then 4-byte aligned)
            g.off_erow = g.off_exit + g.nblk * g.ring_rows * 2;
            g.total = g.off_bits + (bits_smem ? bits_bytes : 0);
            if (g.total < redo_need) g.total = redo_need;
            return true;
        }
    }
    return false;
}

// Shared-memory layout and team shape for K CTAs per utterance; false when nothing fits.
// Teams: a pass (every thread contracts its 4 x 8 cells over all channels) takes the same time
// whatever the team shape, so the shape that needs the fewest passes for a full-length utterance
// wins; among those, more teams (narrower chunks: the sweep starts earlier and trails less).
static bool make_geom(int D, int T_x, int T_y, int K, int ffma_all, int teams_forced, int max_smem, Geom &best) {
    Geom g0{};
    const int R = tokens_per_lane(T_x, K);
    if (R > 8) return false;
    g0.K = K;
    g0.R = R;
    g0.ffma_all = ffma_all;
    g0.max_slice = ceil_div(ceil_div(T_x, K), R) * R;
    g0.tr_max = ceil_div(g0.max_slice, kGemmTM) * kGemmTM;
    g0.ring_rows = ceil_div(g0.tr_max, 8) * 8;
    g0.nblk = ceil_div(T_y, kBlk);
    g0.t_ref = make_tile_shape(T_x, T_y);
    g0.nsh = kGemmThreads / g0.t_ref.tile_rows;
    g0.nsh = g0.nsh < 1 ? 1 : (g0.nsh > 4 ? 4 : g0.nsh);
    g0.dsh = ceil_div(D, g0.nsh);
    const int RG = g0.tr_max / kGemmTM, warps = ffma_all ? 15 : 12;
    const int frames = slice_frames(T_x, T_y, g0.max_slice);
    int best_passes = -1;
    for (int nt = 1; nt <= 5; ++nt) {
        if (warps % nt) continue;
        if (teams_forced > 0 && nt != teams_forced) continue;
        const int tw = warps / nt, tn = tw * 32;
        int cg = tn / RG;
        cg = cg > 32 ? 32 : cg;
        cg = cg > tn / 8 ? tn / 8 : cg;
        if (cg < 4) continue;
        Geom g = g0;
        g.nteams = nt;
        g.team_warps = tw;
        g.CG_cap = cg;
        g.F_cap = 8 * cg;
        if (!layout_geom(D, T_x, max_smem, g)) continue;
        g.passes = ceil_div(frames, nt * g.F_cap);
        if (best_passes < 0 || g.passes <= best_passes) {
            best_passes = g.passes;
            best = g;
        }
    }
    return best_passes >= 0;
}

static size_t redo_scratch_bytes(int slots, int T_x, int T_y) {
    const TileShape t = make_tile_shape(T_x, T_y);
    return align_up((size_t)slots * T_x * t.F * 4, 256);
}
static size_t redo_bits_bytes(int slots, int T_x, int T_y) { return align_up((size_t)slots * ceil_div(T_y, kBlk) * (T_x + 64) * 4, 256); }
static size_t ws_bits_bytes(int slots, int T_x, int T_y) {
    // [clusters][K][nblk][ring_rows]: K x ring_rows <= T_x + K x (R + 4 + 8) rounding
    return align_up((size_t)slots * ceil_div(T_y, kBlk) * (T_x + 160) * 4, 256);
}
static size_t ws_maps_bytes(int slots, int T_x, int T_y) { return align_up((size_t)slots * ceil_div(T_y, kBlk) * (T_x + 160), 256); }

template <int R>
static int launch_r(const CUtensorMap &tmap_z, const PathParams &pp, const LogpParams &lp, Geom &g, int B, int dev, cudaStream_t stream) {
    static SmemOptIn optin[2];
    const bool dbgk = pp.dbg_cycles != nullptr;
    auto kern = dbgk ? mas_fused_kernel<R, true> : mas_fused_kernel<R, false>;
    if (int rc = optin[dbgk].ensure(kern, dev, g.total)) return rc;
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = (size_t)g.total;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)g.K;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    static const char *pdl_env = getenv("MAS_B200_PDL");       // experiment hook
    cfg.attrs = attr;
    cfg.numAttrs = (pdl_env && atoi(pdl_env) == 0) ? 1 : 2;
    // how many clusters the device holds at once (per device and K: the query is not free)
    static std::atomic<int> cache[2][64][9];
    int nc = cache[dbgk][dev & 63][g.K].load();
    if (nc == 0) {
        cfg.gridDim = dim3((unsigned)(g.K * kMaxClusters));
        MAS_CUDA_TRY(cudaOccupancyMaxActiveClusters(&nc, kern, &cfg));
        if (nc < 1) return MAS_ERR_UNSUPPORTED_SHAPE;
        cache[dbgk][dev & 63][g.K].store(nc);
    }
    if (nc > kMaxClusters) nc = kMaxClusters;
    g.NC = B < nc ? B : nc;
    cfg.gridDim = dim3((unsigned)(g.NC * g.K));
    MAS_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tmap_z, pp, lp, g));
    return MAS_OK;
}

// What the launch costs, from the timelines measured on B200 (profiles/r2_fused_timeline.txt): per
// round of resident clusters ~18.4 us per contraction pass, ~25 us of start, sweep tail and output,
// ~4 us per CTA the sweep's chain and the backtrack hop through; direction bits and block maps in
// the workspace (long slices) cost ~3 us per 32-frame block of L2 round trips.
static double estimate_us(const Geom &g, int B, int num_sms) {
    const int nc = num_sms / g.K < 1 ? 1 : num_sms / g.K;
    const int rounds = ceil_div(B, nc);
    const int slice_blocks = ceil_div(g.passes * g.nteams * g.F_cap, kBlk);
    // (the serial token walk of long slices: ~0.05 us per token and CTA hop; from the workspace ~0.4 us per block more)
    const double walk = g.use_maps ? 0.0 : 0.05 * g.max_slice * g.K + (g.bits_in_smem ? 0.0 : 0.4 * slice_blocks * g.K);
    return rounds * (g.passes * 18.4 + 25.0 + 4.0 * g.K + walk);
}
// ... and the two kernels back to back: the materialising kernel's rounds of (token tile, chunk)
// units (16.5 us each on every SM, mas_logp.cu) + kernel (1) (latency floor 41 ns per frame, else
// ~500 Gcells/s).
static double estimate_two_kernels_us(int B, int T_x, int T_y, int num_sms) {
    const TileShape t = make_tile_shape(T_x, T_y);
    const double units = (double)B * t.row_tiles * t.nchunks;
    const double logp_us = ceil(units / (double)num_sms) * 16.5 + 6.0;
    const double cells = (double)B * T_x * T_y;
    // kernel (1): ~500 Gcells/s once every SM streams, else the dependent chain: 41 ns per frame with up to
    // three sweep warps (<= 288 tokens), 53 ns with more (measured: profiles/r2_sweep_k.txt)
    const double per_frame = T_x <= 288 ? 0.041 : 0.0535;
    const double k1 = cells / 0.5e6 > per_frame * T_y ? cells / 0.5e6 : per_frame * T_y;
    return logp_us + k1 + 5.0;
}

// The geometry for `num_sms` SMs with `max_smem` bytes of opt-in shared memory per CTA.
static bool choose_geom(int B, int D, int T_x, int T_y, int max_smem, int num_sms, Geom &best) {
    static const char *k_env = getenv("MAS_B200_FUSED_K");          // experiment hooks
    static const char *ffma_env = getenv("MAS_B200_FUSED_FFMA");
    static const char *teams_env = getenv("MAS_B200_FUSED_TEAMS");
    const int ffma_all = ffma_env ? (atoi(ffma_env) >= 15) : 1;
    const int teams_forced = teams_env ? atoi(teams_env) : 0;
    double best_cost = -1.0;
    for (int K = 1; K <= 8; K *= 2) {
        if (k_env && atoi(k_env) != K) continue;
        if (K > 1 && ceil_div(T_x, K) < 8) continue;                // slices of a handful of tokens are not worth a CTA
        Geom g;
        if (!make_geom(D, T_x, T_y, K, ffma_all, teams_forced, max_smem, g)) continue;
        const int nc = num_sms / K;
        if (nc < 1) continue;
        const double cost = estimate_us(g, B, num_sms);
        if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            best = g;
        }
    }
    return best_cost >= 0;
}

}  // namespace fused

// Scratch of the single launch: per resident cluster (at most 192) the redo scratch and direction
// bits; no score matrix, no rings.  Independent of the batch size beyond 192 utterances.
size_t fused_workspace_bytes(int B, int D, int T_x, int T_y) {
    (void)D;
    using namespace fused;
    const int slots = B < kMaxClusters ? B : kMaxClusters;
    return redo_scratch_bytes(slots, T_x, T_y) + redo_bits_bytes(slots, T_x, T_y) + ws_bits_bytes(slots, T_x, T_y) +
           ws_maps_bytes(slots, T_x, T_y);
}

// Host-only: the geometry the launcher picks.
// out12 = {K, R, max_slice, nteams, team_warps, CG, F, NB, bits_in_smem, total, ffma warps, ring_rows}.
bool debug_fused_geom(int B, int D, int T_x, int T_y, int max_smem, int num_sms, int32_t *out12) {
    fused::Geom g;
    if (!fused::choose_geom(B, D, T_x, T_y, max_smem, num_sms, g)) return false;
    out12[0] = g.K, out12[1] = fused::tokens_per_lane(T_x, g.K), out12[2] = g.max_slice, out12[3] = g.nteams;
    out12[4] = g.team_warps, out12[5] = g.CG_cap, out12[6] = g.F_cap, out12[7] = g.NB, out12[8] = g.bits_in_smem;
    out12[9] = g.total, out12[10] = g.ffma_all ? 15 : 12, out12[11] = g.ring_rows;
    return true;
}

// MAS_OK: launched.  MAS_ERR_UNSUPPORTED_SHAPE: not for the single launch (the caller runs the two
// kernels back to back instead).
int launch_fused(const LogpParams &lp_in, const int32_t *x_len, const int32_t *y_len, float *path, int32_t *durations,
                 int32_t *frame_token, void *workspace, size_t workspace_bytes, float max_neg_val, bool force, cudaStream_t stream) {
    using namespace fused;
    static const bool debug = getenv("MAS_B200_DEBUG") != nullptr;
#define MAS_FUSED_NO(why) do { if (debug) fprintf(stderr, "[mas_b200] single launch not taken: %s\n", why); return MAS_ERR_UNSUPPORTED_SHAPE; } while (0)
    const int B = lp_in.B, D = lp_in.D, T_x = lp_in.T_x, T_y = lp_in.T_y;
    if (B == 0) return MAS_OK;
    if ((T_y & 3) || D > logp::kPanel || D < 1 || (reinterpret_cast<uintptr_t>(lp_in.z) & 15) ||
        (reinterpret_cast<uintptr_t>(path) & 15))
        MAS_FUSED_NO("alignment / frame count / channel count");
    if (workspace == nullptr || workspace_bytes < fused_workspace_bytes(B, D, T_x, T_y)) return MAS_ERR_WORKSPACE_TOO_SMALL;

    int dev = 0;
    MAS_CUDA_TRY(cudaGetDevice(&dev));
    DeviceInfo di{};
    if (int rc = get_device_info(dev, di)) return rc;
    Geom g;
    if (!choose_geom(B, D, T_x, T_y, di.max_smem_optin - 1024, di.num_sms, g)) MAS_FUSED_NO("no slice geometry fits shared memory");
    static const char *mode_env = getenv("MAS_B200_FUSED_MODE");    // "cluster": always the single launch (experiments)
    if (!force && !(mode_env && mode_env[0] == 'c')) {
        const double one = estimate_us(g, B, di.num_sms), two = estimate_two_kernels_us(B, T_x, T_y, di.num_sms);
        if (debug) fprintf(stderr, "[mas_b200] estimates: single launch %.0f us, two kernels %.0f us\n", one, two);
        if (two < one) MAS_FUSED_NO("the two kernels are estimated faster for this shape");
    }

    const int slots = B < kMaxClusters ? B : kMaxClusters;
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    g.redo_scratch = reinterpret_cast<float *>(ws);
    ws += redo_scratch_bytes(slots, T_x, T_y);
    g.redo_bits = reinterpret_cast<uint32_t *>(ws);
    ws += redo_bits_bytes(slots, T_x, T_y);
    g.ws_bits = reinterpret_cast<uint32_t *>(ws);
    ws += ws_bits_bytes(slots, T_x, T_y);
    g.ws_maps = ws;

    LogpParams lp = lp_in;
    lp.logp = nullptr;
    lp.x_len = x_len;
    lp.y_len = y_len;
    PathParams pp{};
    pp.t_x = x_len;
    pp.t_y = y_len;
    pp.path = path;
    pp.durations = durations;
    pp.frame_token = frame_token;
    pp.B = B;
    pp.T_x = T_x;
    pp.T_y = T_y;
    pp.max_neg_val = max_neg_val;
    pp.dbg_cycles = g_dbg_cycles.load();

    // z [B][D][T_y] for the teams' panel loads: boxes of [kChan channels x F frames]; channels and frames
    // out of range arrive as zeros
    PFN_cuTensorMapEncodeTiled_v12000 encode = systolic::get_encode_fn();
    if (encode == nullptr) MAS_FUSED_NO("no cuTensorMapEncodeTiled");
    CUtensorMap tmap_z;
    {
        const cuuint64_t gdim[3] = {(cuuint64_t)T_y, (cuuint64_t)D, (cuuint64_t)B};
        const cuuint64_t gstride[2] = {(cuuint64_t)T_y * 4, (cuuint64_t)D * T_y * 4};
        const cuuint32_t box[3] = {(cuuint32_t)g.F_cap, (cuuint32_t)kChan, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult cr = encode(&tmap_z, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(lp_in.z), gdim, gstride, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) MAS_FUSED_NO("tensor map encode failed");
    }
    const int R = tokens_per_lane(T_x, g.K);
    if (debug)
        fprintf(stderr, "[mas_b200] single launch: clusters of %d CTAs, slices of <= %d tokens (R=%d), %d FFMA warps in %d teams, chunks of <= %d frames, ring of %d boxes x %d rows, bits %s, %d B smem\n",
                g.K, g.max_slice, R, g.ffma_all ? 15 : 12, g.nteams, g.F_cap, g.NB, g.ring_rows, g.bits_in_smem ? "in smem" : "in workspace", g.total);
    switch (R) {
        case 1: return launch_r<1>(tmap_z, pp, lp, g, B, dev, stream);
        case 2: return launch_r<2>(tmap_z, pp, lp, g, B, dev, stream);
        case 3: return launch_r<3>(tmap_z, pp, lp, g, B, dev, stream);
        case 4: return launch_r<4>(tmap_z, pp, lp, g, B, dev, stream);
        case 5: return launch_r<5>(tmap_z, pp, lp, g, B, dev, stream);
        case 6: return launch_r<6>(tmap_z, pp, lp, g, B, dev, stream);
        case 8: return launch_r<8>(tmap_z, pp, lp, g, B, dev, stream);
        default: return MAS_ERR_UNSUPPORTED_SHAPE;
    }
}

}  // namespace mas

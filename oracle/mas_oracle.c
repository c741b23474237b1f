/*
 * mas_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C CPU restatement of the reference's alignment hot path, used only as the
 * checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs.  Nothing under glow-tts-train_b200/ may import, link or execute it.
 *
 * Parity status: PINNED.  oracle/build_ref.py compiles the reference's own Cython kernel
 * (glow_tts_train/monotonic_align/core.pyx) from where it lies under /root/reference into
 * oracle/_ref/, and tests/test_oracle.py + tests/golden/make_golden.py check this restatement
 * against it (and against the committed golden vectors it produced) bit for bit.
 * The logp restatement has no reference test pinning it at the cuBLAS/MKL boundary
 * (SURVEY.md 8c); it is pinned against the reference's models.py run on CPU here
 * (tests/golden/model_logp_*.npz) to 1e-5 relative.
 *
 * What is restated (reference file:line):
 *   mas_oracle_each        <- glow_tts_train/monotonic_align/core.pyx:9-35   (maximum_path_each)
 *   mas_oracle_batch       <- glow_tts_train/monotonic_align/core.pyx:40-45  (maximum_path_c)
 *   mas_oracle_lengths     <- glow_tts_train/monotonic_align/__init__.py:18-19 (t_x / t_y from the mask)
 *   mas_oracle_logp_f64    <- glow_tts_train/models.py:363-376 (the four log-likelihood terms), in fp64
 *   mas_oracle_logp_f32    <- same, fp32 arithmetic in the reference's term order
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared [-fopenmp] -o oracle/_build/libmas_oracle.so oracle/mas_oracle.c -lm
 * (no -ffast-math, no FMA contraction: every fp32 add/mul has to be a plain round-to-nearest op.)
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* One utterance.  value is [t_x_stride rows][ld] fp32, row-major (mel frame index contiguous),
 * mutated in place into cumulative scores exactly like the reference; path is int32, same shape,
 * pre-zeroed by the caller (core.pyx:9-35). */
void mas_oracle_each(int32_t *path, float *value, int64_t ld, int t_x, int t_y, float max_neg_val)
{
    /* forward sweep: frame by frame, token by token inside the reachable band (core.pyx:17-30) */
    for (int frame = 0; frame < t_y; ++frame) {
        int tok_lo = t_x + frame - t_y;
        if (tok_lo < 0) tok_lo = 0;
        int tok_hi = (frame + 1 < t_x) ? frame + 1 : t_x;
        for (int tok = tok_lo; tok < tok_hi; ++tok) {
            float stay, advance;
            if (tok == frame)
                stay = max_neg_val;                                  /* core.pyx:19-20 */
            else
                stay = value[(int64_t)tok * ld + (frame - 1)];      /* core.pyx:22 */
            if (tok == 0)
                advance = (frame == 0) ? 0.0f : max_neg_val;         /* core.pyx:23-27 */
            else
                advance = value[(int64_t)(tok - 1) * ld + (frame - 1)]; /* core.pyx:29 */
            /* Cython's max(v_cur, v_prev) expands to (v_prev > v_cur) ? v_prev : v_cur
             * (core.c:2697-2708): a tie, or any NaN, keeps v_cur. */
            float best = (advance > stay) ? advance : stay;
            value[(int64_t)tok * ld + frame] = best + value[(int64_t)tok * ld + frame]; /* plain RN fp32 add, core.pyx:30 */
        }
    }
    /* backtrack from the last token at the last frame (core.pyx:32-35) */
    int tok = t_x - 1;
    for (int frame = t_y - 1; frame >= 0; --frame) {
        path[(int64_t)tok * ld + frame] = 1;
        if (tok != 0 &&
            (tok == frame ||
             value[(int64_t)tok * ld + (frame - 1)] < value[(int64_t)(tok - 1) * ld + (frame - 1)]))
            tok -= 1;
    }
}

/* Batch driver (core.pyx:40-45).  paths/values are [B][T_x][T_y] C-contiguous.
 * threads <= 1 -> serial (what the reference's own setup.py builds);
 * threads  > 1 -> OpenMP static schedule over utterances (the prange) when compiled -fopenmp. */
void mas_oracle_batch(int32_t *paths, float *values, const int32_t *t_xs, const int32_t *t_ys,
                      int B, int T_x, int T_y, float max_neg_val, int threads)
{
    const int64_t slab = (int64_t)T_x * T_y;
#ifdef _OPENMP
    if (threads > 1) {
#pragma omp parallel for schedule(static) num_threads(threads)
        for (int b = 0; b < B; ++b)
            mas_oracle_each(paths + b * slab, values + b * slab, T_y, t_xs[b], t_ys[b], max_neg_val);
        return;
    }
#else
    (void)threads;
#endif
    for (int b = 0; b < B; ++b)
        mas_oracle_each(paths + b * slab, values + b * slab, T_y, t_xs[b], t_ys[b], max_neg_val);
}

int mas_oracle_has_openmp(void)
{
#ifdef _OPENMP
    return 1;
#else
    return 0;
#endif
}

/* Lengths the way the reference wrapper derives them: column 0 / row 0 sums of the mask, cast to
 * int32 (__init__.py:18-19: mask.sum(1)[:, 0] and mask.sum(2)[:, 0]). */
void mas_oracle_lengths(const float *mask, int B, int T_x, int T_y, int32_t *t_xs, int32_t *t_ys)
{
    const int64_t slab = (int64_t)T_x * T_y;
    for (int b = 0; b < B; ++b) {
        float sx = 0.0f, sy = 0.0f;
        for (int x = 0; x < T_x; ++x) sx += mask[b * slab + (int64_t)x * T_y];
        for (int y = 0; y < T_y; ++y) sy += mask[b * slab + y];
        t_xs[b] = (int32_t)sx;
        t_ys[b] = (int32_t)sy;
    }
}

/* Log-likelihood matrix, fp64 evaluation of models.py:363-376.
 *   x_m, x_logs : [B][D][T_x]   (x_logs may be NULL == all zeros, the mean_only case)
 *   z           : [B][D][T_y]
 *   logp        : [B][T_x][T_y] (written as double)
 * D <= MAS_ORACLE_MAX_D.                                                                      */
#define MAS_ORACLE_MAX_D 1024
int mas_oracle_logp_f64(const float *x_m, const float *x_logs, const float *z, double *logp,
                        int B, int D, int T_x, int T_y)
{
    if (D > MAS_ORACLE_MAX_D) return -1;
    const double half_log_2pi = 0.5 * log(2.0 * M_PI);
#ifdef _OPENMP
#pragma omp parallel for collapse(2) schedule(static)
#endif
    for (int b = 0; b < B; ++b) {
        for (int x = 0; x < T_x; ++x) {
            const float *m = x_m + (int64_t)b * D * T_x;
            const float *s = x_logs ? x_logs + (int64_t)b * D * T_x : NULL;
            const float *zz = z + (int64_t)b * D * T_y;
            double *out = logp + (int64_t)b * T_x * T_y;
            double inv_var[MAS_ORACLE_MAX_D], mean_over_var[MAS_ORACLE_MAX_D];
            double l1 = 0.0, l4 = 0.0;
            for (int d = 0; d < D; ++d) {
                double ls = s ? (double)s[(int64_t)d * T_x + x] : 0.0;
                double r = exp(-2.0 * ls);                         /* models.py:363 */
                double mu = (double)m[(int64_t)d * T_x + x];
                inv_var[d] = r;
                mean_over_var[d] = mu * r;
                l1 += -half_log_2pi - ls;                          /* models.py:364-366 */
                l4 += -0.5 * mu * mu * r;                          /* models.py:373-375 */
            }
            for (int y = 0; y < T_y; ++y) {
                double l2 = 0.0, l3 = 0.0;
                for (int d = 0; d < D; ++d) {
                    double zv = (double)zz[(int64_t)d * T_y + y];
                    l2 += inv_var[d] * (-0.5 * zv * zv);           /* models.py:367-369 */
                    l3 += mean_over_var[d] * zv;                   /* models.py:370-372 */
                }
                out[(int64_t)x * T_y + y] = ((l1 + l2) + l3) + l4; /* models.py:376 */
            }
        }
    }
    return 0;
}

/* Same formula in fp32 arithmetic, reference term order ((l1+l2)+l3)+l4, contraction summed in
 * ascending channel order.  The reference's matmul K-order is unspecified (cuBLAS/MKL), so this is
 * ONE admissible fp32 result, not THE result; tests use it to bound fp32 noise, and the f64
 * version as the pin. */
int mas_oracle_logp_f32(const float *x_m, const float *x_logs, const float *z, float *logp,
                        int B, int D, int T_x, int T_y)
{
    if (D > MAS_ORACLE_MAX_D) return -1;
    const float c = (float)(-0.5 * log(2.0 * M_PI));
#ifdef _OPENMP
#pragma omp parallel for collapse(2) schedule(static)
#endif
    for (int b = 0; b < B; ++b) {
        for (int x = 0; x < T_x; ++x) {
            const float *m = x_m + (int64_t)b * D * T_x;
            const float *s = x_logs ? x_logs + (int64_t)b * D * T_x : NULL;
            const float *zz = z + (int64_t)b * D * T_y;
            float *out = logp + (int64_t)b * T_x * T_y;
            float inv_var[MAS_ORACLE_MAX_D], mean_over_var[MAS_ORACLE_MAX_D];
            float l1 = 0.0f, l4 = 0.0f;
            for (int d = 0; d < D; ++d) {
                float ls = s ? s[(int64_t)d * T_x + x] : 0.0f;
                float r = expf(-2.0f * ls);
                float mu = m[(int64_t)d * T_x + x];
                inv_var[d] = r;
                mean_over_var[d] = mu * r;
                l1 = l1 + (c - ls);
                l4 = l4 + (-0.5f * (mu * mu)) * r;
            }
            for (int y = 0; y < T_y; ++y) {
                float l2 = 0.0f, l3 = 0.0f;
                for (int d = 0; d < D; ++d) {
                    float zv = zz[(int64_t)d * T_y + y];
                    l2 = l2 + inv_var[d] * (-0.5f * (zv * zv));
                    l3 = l3 + mean_over_var[d] * zv;
                }
                out[(int64_t)x * T_y + y] = ((l1 + l2) + l3) + l4;
            }
        }
    }
    return 0;
}

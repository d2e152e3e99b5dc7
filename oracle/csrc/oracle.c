/* Plain-C oracle routines for the Flood-complex hot path.  TEST INFRASTRUCTURE ONLY
 * (see oracle/__init__.py): the checker for the CUDA kernels, never the product.
 *
 * Build: oracle/Makefile  ->  oracle/_build/liboracle.so   (gcc -O2 -ffp-contract=off, so
 * that every float operation below rounds exactly as written; fused operations are
 * requested explicitly with fmaf()).
 *
 * Reference lines restated (paths relative to /root/reference):
 *   oracle_fps_f32          exact farthest-point sampling, what flooder/core.py:337-342
 *                           obtains from fpsample 0.3.3 (bucket-FPS == exact FPS)
 *   oracle_ball_counts_f32  predicate of flooder/triton_kernels.py:137-148
 *   oracle_min_dist_f32     quantity of flooder/triton_kernels.py:36-45 (min over the ball
 *                           candidates of the direct-difference distance), brute force
 */
#include <math.h>
/* fmaf() is a libm call on baseline x86-64; clone the hot routines for CPUs with FMA3 so
 * the same correctly-rounded fused operation runs as one instruction (result identical). */
#if defined(__x86_64__) && defined(__GNUC__)
#define ORACLE_CLONES __attribute__((target_clones("avx2,fma", "default")))
#else
#define ORACLE_CLONES
#endif
#include <stdint.h>
#include <stdlib.h>

/* idx[0] = start; idx[k+1] = argmax_i min_{j<=k} |p_i - p_idx[j]|^2, first maximum wins.
 * Squared distance: ((dx*dx + dy*dy) + dz*dz) ... each product and sum rounded to float. */
int oracle_fps_f32(const float *pts, int64_t n, int d, int64_t n_samples, int64_t start,
                   int64_t *out_idx)
{
    if (n <= 0 || d <= 0 || n_samples <= 0 || start < 0 || start >= n) return -1;
    if (n_samples > n) n_samples = n;
    float *mind = (float *)malloc(sizeof(float) * (size_t)n);
    if (!mind) return -2;
    for (int64_t i = 0; i < n; ++i) mind[i] = INFINITY;
    int64_t cur = start;
    for (int64_t k = 0; k < n_samples; ++k) {
        out_idx[k] = cur;
        const float *q = pts + cur * d;
        float best = -1.0f;
        int64_t best_i = 0;
        for (int64_t i = 0; i < n; ++i) {
            const float *p = pts + i * d;
            float t = p[0] - q[0];
            float s = t * t;
            for (int j = 1; j < d; ++j) {
                t = p[j] - q[j];
                float sq = t * t;
                s = s + sq;
            }
            float m = mind[i];
            if (s < m) { m = s; mind[i] = m; }
            if (m > best) { best = m; best_i = i; }
        }
        cur = best_i;
    }
    free(mind);
    return 0;
}

/* counts[b] = #{ i : fma-chain sum_j (p_ij - c_bj)^2 <= r_b * r_b } */
ORACLE_CLONES
int oracle_ball_counts_f32(const float *pts, int64_t n, int d, const float *centers,
                           const float *radii, int64_t n_balls, int64_t *counts)
{
    for (int64_t b = 0; b < n_balls; ++b) {
        const float *c = centers + b * d;
        const float r2 = radii[b] * radii[b];
        int64_t cnt = 0;
        for (int64_t i = 0; i < n; ++i) {
            const float *p = pts + i * d;
            float s = 0.0f;
            for (int j = 0; j < d; ++j) {
                float t = p[j] - c[j];
                s = fmaf(t, t, s);
            }
            cnt += (s <= r2);
        }
        counts[b] = cnt;
    }
    return 0;
}

/* out[s*R + r] = sqrt( min_{i in ball(s)} fma-chain |x_sr - p_i|^2 ), +inf when the ball
 * is empty.  use_ball == 0 takes the minimum over the whole cloud. */
ORACLE_CLONES
int oracle_min_dist_f32(const float *pts, int64_t n, int d, const float *samples, int64_t S,
                        int64_t R, const float *centers, const float *radii, int use_ball,
                        float *out)
{
    unsigned char *in = (unsigned char *)malloc((size_t)n);
    if (!in) return -2;
    for (int64_t s = 0; s < S; ++s) {
        if (use_ball) {
            const float *c = centers + s * d;
            const float r2 = radii[s] * radii[s];
            for (int64_t i = 0; i < n; ++i) {
                const float *p = pts + i * d;
                float a = 0.0f;
                for (int j = 0; j < d; ++j) {
                    float t = p[j] - c[j];
                    a = fmaf(t, t, a);
                }
                in[i] = (a <= r2);
            }
        }
        for (int64_t r = 0; r < R; ++r) {
            const float *x = samples + (s * R + r) * d;
            float best = INFINITY;
            for (int64_t i = 0; i < n; ++i) {
                if (use_ball && !in[i]) continue;
                const float *p = pts + i * d;
                float t = x[0] - p[0];
                float a = t * t;
                for (int j = 1; j < d; ++j) {
                    t = x[j] - p[j];
                    a = fmaf(t, t, a);
                }
                if (a < best) best = a;
            }
            out[s * R + r] = sqrtf(best);
        }
    }
    free(in);
    return 0;
}

/* x[s,r,:] = fma-chain over k of w[r,k] * v[s,k,:]  (k ascending, accumulator starts at the
 * k = 0 product), the arithmetic the CUDA kernel uses for on-the-fly sample points. */
ORACLE_CLONES
int oracle_sample_points_f32(const float *weights, int64_t R, int K, const float *verts,
                             int64_t S, int d, float *out)
{
    for (int64_t s = 0; s < S; ++s)
        for (int64_t r = 0; r < R; ++r)
            for (int j = 0; j < d; ++j) {
                float a = weights[r * K] * verts[(s * K) * d + j];
                for (int k = 1; k < K; ++k)
                    a = fmaf(weights[r * K + k], verts[(s * K + k) * d + j], a);
                out[(s * R + r) * d + j] = a;
            }
    return 0;
}

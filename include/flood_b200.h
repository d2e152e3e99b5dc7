/* flood_b200.h -- C ABI of libflood_b200.so: the B200-native (sm_100a) Flood-complex hot path.
 *
 * The reference (plus-rkwitt/flooder) has no FFI layer; its seams for this path are Python
 * call sites.  Each entry point below names the reference interface it replaces
 * (paths relative to the reference root):
 *
 *   flood_fps_f32                 fpsample.bucket_fps_kdline_sampling(...)      flooder/core.py:337-342
 *   flood_cloud_build_f32         sort of the cloud + slab search column         flooder/core.py:140-144, 201-208
 *   flood_bounding_balls_f32      simplex centres / radii                        flooder/core.py:156-172
 *   flood_covering_radius_f32     compute_mask + torch.nonzero + compute_filtration
 *                                 (+ the weights @ vertices product)             flooder/core.py:188, 210-226
 *                                                                                flooder/triton_kernels.py:48-96, 161-223
 *   flood_face_max_f32            per-face / per-simplex maxima                  flooder/core.py:251-257, 270
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless it is marked "host";
 *     the library never allocates or frees device memory: scratch space is passed in as a
 *     workspace whose size the matching *_workspace_bytes() function reports;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = default stream)
 *     and the calls return without synchronising, except where stated;
 *   - return value 0 = success, negative = error (FLOOD_E_*); a human-readable message for
 *     the calling thread's last error is returned by flood_last_error();
 *   - no exceptions cross the boundary; the compute entry points are re-entrant per (device,
 *     stream) as long as the workspaces are distinct.  flood_set_option, flood_kernel_ms and
 *     flood_launch_count are process-wide diagnostics (one option table, one set of timers) and
 *     are not covered by that guarantee: set options before concurrent use, not during it.
 */
#ifndef FLOOD_B200_H
#define FLOOD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLOOD_ABI_VERSION 1

#define FLOOD_OK 0
#define FLOOD_E_INVALID (-1)   /* bad argument (null pointer, unsupported dimension, ...) */
#define FLOOD_E_WORKSPACE (-2) /* workspace too small */
#define FLOOD_E_CUDA (-3)      /* a CUDA runtime call or launch failed */
#define FLOOD_E_UNSUPPORTED (-4)

#define FLOOD_MAX_DIM 8          /* ambient dimension D: 1..8 */
#define FLOOD_MAX_SIMPLEX_VERTS 9 /* K = d+1 <= D+1 */

int flood_abi_version(void);
const char *flood_last_error(void);

/* Number of SMs / clock of the current device (used by the benchmark's roofline). */
int flood_device_info(int *sm_count, int *sm_clock_khz);

/* ---------------------------------------------------------------------------------------
 * Farthest-point sampling (exact).  idx[0] = start_idx,
 *   idx[k+1] = argmax_i min_{j<=k} |p_i - p_idx[j]|^2      (first maximum wins)
 * float32, squared distance summed in coordinate order WITHOUT fused multiply-add, i.e. the
 * arithmetic of a scalar CPU implementation, so that indices are bit-exact.
 * pts: [n, d] row-major.  out_idx: [n_lms] int64.  Persistent cooperative kernel.
 * ------------------------------------------------------------------------------------- */
size_t flood_fps_workspace_bytes(int64_t n, int d, int64_t n_lms);
int flood_fps_f32(const float *pts, int64_t n, int d, int64_t n_lms, int64_t start_idx,
                  int64_t *out_idx, void *workspace, size_t workspace_bytes, void *stream);

/* Same result (bit-exact indices) from the bucketed variant: it walks the cell grid of a prepared
 * cloud (flood_cloud_build_f32 for the same pts / n / d, d >= 2) and only touches the cells whose
 * running maxima can change -- the pruning idea of the reference's bucket-FPS on the grid the
 * covering kernel needs anyway.  Workspace size: flood_fps_workspace_bytes().  Reads the grid
 * header back from the device, i.e. synchronises `stream` once before launching. */
int flood_fps_grid_f32(const void *cloud_workspace, const float *pts, int64_t n, int d, int64_t n_lms,
                       int64_t start_idx, int64_t *out_idx, void *workspace, size_t workspace_bytes,
                       void *stream);

/* ---------------------------------------------------------------------------------------
 * Cloud preparation: bins the cloud into a uniform cell grid over its first min(d,5) axes and
 * stores it cell-sorted as padded records (float2 / float4 / 2 x float4), so that a ball maps to
 * a few contiguous runs.  The prepared cloud lives entirely inside `workspace`; the same (pointer, n, d) triple
 * is handed to flood_covering_radius_f32.  points_per_cell <= 0 selects the default.
 * ------------------------------------------------------------------------------------- */
size_t flood_cloud_workspace_bytes(int64_t n, int d);
int flood_cloud_build_f32(const float *pts, int64_t n, int d, int points_per_cell,
                          void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------
 * Bounding balls of simplices, the reference's candidate rule:
 *   centre = midpoint of the longest edge (first maximum of the flattened KxK distance matrix),
 *   radius = max_k |v_k - centre| * (K > 2 ? 1.42 : 1.01) + 1e-3.
 * verts: [S, K, d].  centers: [S, d].  radii: [S].
 * ------------------------------------------------------------------------------------- */
int flood_bounding_balls_f32(const float *verts, int64_t S, int K, int d, float *centers,
                             float *radii, void *stream);

/* ---------------------------------------------------------------------------------------
 * Covering radius samples.  For every simplex s and sample r
 *     out_min_dist2[s, r] = min_{p in cloud, sum_i (p_i - c_s,i)^2 <= r_s^2} |x_sr - p|^2
 * with x_sr = sum_k weights[r, k] * verts[s, k, :] (fused multiply-add chain over k
 * ascending; bit-identical to the reference's float32 matmul) or, when `samples` is not
 * NULL, x_sr = samples[s, r, :].  Distances are direct differences in float32
 * ((x-p)^2 accumulated with FMA), +inf when the ball holds no cloud point.
 *
 *   cloud_workspace   result of flood_cloud_build_f32 for (n, d)
 *   verts             [S, K, d]       weights  [R, K]        samples  [S, R, d] or NULL
 *   centers, radii    [S, d], [S]     (flood_bounding_balls_f32 or caller-supplied)
 *   out_min_dist2     [S, R] float32  squared distances
 *   out_cand_count    [S] int64 or NULL: number of cloud points inside ball s
 *   out_evals         host-invisible device counter (1 x uint64) or NULL: sum_s R * cand_count[s],
 *                     the algorithmic work count E of the call
 *
 * By default (option "prune" = 1) the sweep is pruned exactly: the samples are handled in bricks
 * of up to 256; a candidate at least as far from the bounding box of a brick as the brick's largest
 * running minimum is skipped for that brick, and (option "level2", sample sets of more than two
 * bricks) the survivors are tested in the same way against the brick's pairs of 32-sample groups
 * before they are evaluated; a seed pass over every 32nd record of the candidate stream (option
 * "seed_stride") gives every sample a finite bound first.  The result is bit-identical to the
 * exhaustive sweep ("prune" = 0); fewer evaluations are executed, E still counts the reference's
 * ball rule.  The number of evaluations actually executed is left as a
 * uint64 at byte FLOOD_COVER_WS_EXECUTED_OFFSET of `workspace` (device memory).
 * ------------------------------------------------------------------------------------- */
#define FLOOD_COVER_WS_EXECUTED_OFFSET 16
size_t flood_covering_workspace_bytes(int64_t S, int64_t R, int d);
int flood_covering_radius_f32(const void *cloud_workspace, int64_t n, int d, const float *verts,
                              int64_t S, int K, const float *weights, int64_t R,
                              const float *samples, const float *centers, const float *radii,
                              float *out_min_dist2, int64_t *out_cand_count,
                              unsigned long long *out_evals, void *workspace,
                              size_t workspace_bytes, void *stream);

/* Sample layout of the evaluation kernel (host-side query, no device work).  The kernel keeps the
 * samples of a simplex in registers, "brick" by brick: brick i = out_groups[i] consecutive groups of
 * 32 samples held by one warp; *bricks_per_block consecutive bricks belong to one CTA.  A caller
 * that is free to order its samples (the weights of flood_complex are shared by all simplices,
 * flooder/core.py:182-188) should make every brick spatially compact: the exact pruning then skips
 * more candidates.  The result does not depend on the order.  Returns the number of bricks
 * (out_groups may be NULL to query it) or a negative error code.  Depends on the options in
 * effect ("small_max_bricks", "small_shape", "warps"). */
int flood_covering_bricks(int64_t R, int d, int32_t *out_groups, int capacity, int *bricks_per_block);

/* Cost estimate: out_tested[s] = number of cloud points in the cell rows touched by ball s (an
 * upper bound of cand_count[s], obtained from the cell table alone).  New with the multi-GPU
 * sharding (the reference is single-GPU): the host balances simplices over ranks with it. */
int flood_covering_plan_f32(const void *cloud_workspace, int64_t n, int d, const float *centers,
                            const float *radii, int64_t S, int32_t *out_tested, void *stream);

/* ---------------------------------------------------------------------------------------
 * Face maxima.  support: [R] int32 bit masks (bit k set <=> weights[r, k] != 0) or NULL.
 *   support != NULL (grid mode):  out[s, m-1] = sqrt(max_{r : support[r] subset of m} min_dist2[s, r])
 *                                 for every non-empty vertex subset m in 1 .. 2^K - 1
 *   support == NULL (random mode): out[s] = sqrt(max_r min_dist2[s, r])
 * ------------------------------------------------------------------------------------- */
int flood_face_max_f32(const float *min_dist2, int64_t S, int64_t R, const int32_t *support,
                       int K, float *out, void *stream);

/* ---------------------------------------------------------------------------------------
 * float64 variants (the reference runs its kernels in float64 for float64 inputs,
 * flooder/triton_kernels.py:226-229, flooder/core.py:116-123).  Same meaning as the float32 entry
 * points; the ball predicate, the sample points and the distances are evaluated in float64 on
 * `pts` (the original coordinates, [n, d] row-major).  `cloud_workspace` is the prepared cloud of the
 * SAME points rounded to float32 (flood_cloud_build_f32): it only enumerates candidates.  Plain
 * kernels (no pruning): B200's FP64 rate is a fraction of its FP32 rate.
 * ------------------------------------------------------------------------------------- */
int flood_bounding_balls_f64(const double *verts, int64_t S, int K, int d, double *centers,
                             double *radii, void *stream);
size_t flood_covering_workspace_bytes_f64(int64_t S, int d);
int flood_covering_radius_f64(const void *cloud_workspace, const double *pts, int64_t n, int d,
                              const double *verts, int64_t S, int K, const double *weights, int64_t R,
                              const double *centers, const double *radii, double *out_min_dist2,
                              int64_t *out_cand_count, unsigned long long *out_evals, void *workspace,
                              size_t workspace_bytes, void *stream);
int flood_face_max_f64(const double *min_dist2, int64_t S, int64_t R, const int32_t *support, int K,
                       double *out, void *stream);

/* Tuning knobs for experiments (process-wide; returns the previous value, -1 if it was unset). */
int flood_set_option(const char *name, int value);

/* With option "time_kernels" = 1 the library brackets its dominant kernels with CUDA events on the
 * launching stream.  flood_kernel_ms() synchronises on the last recorded launch of `name`
 * ("cover_eval", "fps") and reports the accumulated device time and launch count. */
int flood_kernel_ms(const char *name, double *total_ms, long *launches, int reset);

/* Number of kernels the library has launched in this process (every launch site counts itself);
 * reset != 0 returns the count and clears it.  bench.py reports it as gpu_launches. */
long long flood_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* FLOOD_B200_H */

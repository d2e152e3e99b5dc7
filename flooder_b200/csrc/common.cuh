// Shared declarations for libflood_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "flood_b200.h"

#ifndef __CUDA_ARCH__
#define FLOOD_HOST_ONLY 1
#endif

namespace flood {

// --------------------------------------------------------------------------------------
// error reporting (thread-local message, see abi.cu)
// --------------------------------------------------------------------------------------
int set_error(int code, const char *fmt, ...);

#define FLOOD_CUDA_CHECK(expr)                                                              \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            return ::flood::set_error(FLOOD_E_CUDA, "%s failed: %s (%s:%d)", #expr,         \
                                      cudaGetErrorString(_e), __FILE__, __LINE__);          \
    } while (0)

#define FLOOD_LAUNCH_CHECK(name)                                                            \
    do {                                                                                    \
        cudaError_t _e = cudaGetLastError();                                                \
        if (_e != cudaSuccess)                                                              \
            return ::flood::set_error(FLOOD_E_CUDA, "launch of %s failed: %s", name,        \
                                      cudaGetErrorString(_e));                              \
    } while (0)

int get_option(const char *name, int fallback);
// optional CUDA-event timing of individual kernels (option "time_kernels"); see flood_kernel_ms
void kernel_timer_start(const char *name, cudaStream_t st);
void kernel_timer_stop(const char *name, cudaStream_t st);
int device_sm_count();
// every kernel the library launches is counted (flood_launch_count; bench.py's gpu_launches)
void count_launches(int n);

// --------------------------------------------------------------------------------------
// prepared cloud: layout of the workspace filled by flood_cloud_build_f32
// --------------------------------------------------------------------------------------
// Cell grid over the first min(d, kMaxGridAxes) axes (cubic cells; cell index = mixed radix with
// axis 0 fastest).  Lives in DEVICE memory (head of the cloud workspace) so that no host
// synchronisation is needed between building and using it.
#ifndef FLOOD_MAX_GRID_AXES
#define FLOOD_MAX_GRID_AXES 5
#endif
constexpr int kMaxGridAxes = FLOOD_MAX_GRID_AXES;
__host__ __device__ constexpr int grid_axes(int d) { return d < kMaxGridAxes ? d : kMaxGridAxes; }

struct GridParams {
    float origin[kMaxGridAxes];   // lower corner of the cloud's bounding box
    float inv_h;                  // 1 / cell edge
    float h;
    int n[kMaxGridAxes];          // cells per axis (1 for unused axes)
    int ncells;                   // product of n[]
    int g;                        // binned axes
    int64_t npts;
    int d;
    int pad_;
};

constexpr int kMaxCellsLog2 = 21;
constexpr int64_t kMaxCells = int64_t(1) << kMaxCellsLog2;

__host__ __device__ inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// floats per stored point record: 2 for d<=2, 4 for d<=4, 8 for d<=8
__host__ __device__ constexpr int record_floats(int d) { return d <= 2 ? 2 : (d <= 4 ? 4 : 8); }

struct CloudLayout {
    int64_t off_grid;        // GridParams
    int64_t off_bbox;        // 2 x kMaxGridAxes x uint32 (ordered-int encoded min/max)
    int64_t off_cell_start;  // (max_cells + 1) x int32
    int64_t off_cell_fill;   // max_cells x int32 (scratch)
    int64_t off_cell_id;     // n x int32 (scratch)
    int64_t off_perm;        // n x int32: original index of the i-th cell-sorted point
    int64_t off_points;      // n x record_floats(d) x float, cell-sorted
    int64_t total;
    int64_t max_cells;
};

__host__ __device__ inline int64_t cloud_max_cells(int64_t n) {
    int64_t c = n / 4;
    if (c < 64) c = 64;
    if (c > kMaxCells) c = kMaxCells;
    return c;
}

__host__ __device__ inline CloudLayout cloud_layout(int64_t n, int d) {
    CloudLayout L;
    L.max_cells = cloud_max_cells(n);
    int64_t o = 0;
    L.off_grid = o;        o = align_up(o + (int64_t)sizeof(GridParams), 256);
    L.off_bbox = o;        o = align_up(o + 2 * kMaxGridAxes * 4, 256);
    L.off_cell_start = o;  o = align_up(o + (L.max_cells + 1) * 4, 256);
    L.off_cell_fill = o;   o = align_up(o + L.max_cells * 4, 256);
    L.off_cell_id = o;     o = align_up(o + n * 4, 256);
    L.off_perm = o;        o = align_up(o + n * 4, 256);
    L.off_points = o;      o = align_up(o + n * record_floats(d) * 4, 256);
    L.total = o;
    return L;
}

// cell coordinate of x along one axis (monotone in x); shared by the builder and the queries
__device__ __forceinline__ float cell_coord(float x, float origin, float inv_h) {
    return (x - origin) * inv_h;
}
__device__ __forceinline__ int cell_clamp(float g, int n) {
    int i = (int)floorf(g);
    return i < 0 ? 0 : (i >= n ? n - 1 : i);
}

// --------------------------------------------------------------------------------------
// cell-sorted point records: float2 (d<=2), float4 (d<=4), 2 x float4 (d<=8)
// --------------------------------------------------------------------------------------
template <int D> struct Rec;
template <> struct Rec<1> { using type = float2; };
template <> struct Rec<2> { using type = float2; };
template <> struct Rec<3> { using type = float4; };
template <> struct Rec<4> { using type = float4; };
template <int D> struct Rec { struct alignas(16) type { float4 a, b; }; };  // D = 5..8

template <int D> __device__ __forceinline__ void rec_unpack(const typename Rec<D>::type &r, float (&p)[D]);
template <> __device__ __forceinline__ void rec_unpack<1>(const float2 &r, float (&p)[1]) { p[0] = r.x; }
template <> __device__ __forceinline__ void rec_unpack<2>(const float2 &r, float (&p)[2]) { p[0] = r.x; p[1] = r.y; }
template <> __device__ __forceinline__ void rec_unpack<3>(const float4 &r, float (&p)[3]) { p[0] = r.x; p[1] = r.y; p[2] = r.z; }
template <> __device__ __forceinline__ void rec_unpack<4>(const float4 &r, float (&p)[4]) { p[0] = r.x; p[1] = r.y; p[2] = r.z; p[3] = r.w; }
template <int D> __device__ __forceinline__ void rec_unpack(const typename Rec<D>::type &r, float (&p)[D]) {
    const float q[8] = {r.a.x, r.a.y, r.a.z, r.a.w, r.b.x, r.b.y, r.b.z, r.b.w};
#pragma unroll
    for (int a = 0; a < D; ++a) p[a] = q[a];
}

template <int D> __device__ __forceinline__ typename Rec<D>::type rec_sentinel();
template <> __device__ __forceinline__ float2 rec_sentinel<1>() { return make_float2(INFINITY, INFINITY); }
template <> __device__ __forceinline__ float2 rec_sentinel<2>() { return make_float2(INFINITY, INFINITY); }
template <> __device__ __forceinline__ float4 rec_sentinel<3>() { return make_float4(INFINITY, INFINITY, INFINITY, INFINITY); }
template <> __device__ __forceinline__ float4 rec_sentinel<4>() { return make_float4(INFINITY, INFINITY, INFINITY, INFINITY); }
template <int D> __device__ __forceinline__ typename Rec<D>::type rec_sentinel() {
    typename Rec<D>::type r;
    r.a = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
    r.b = r.a;
    return r;
}

// read-only (non-coherent) load of one record
template <int D> __device__ __forceinline__ typename Rec<D>::type rec_ldg(const typename Rec<D>::type *p) {
    if constexpr (D <= 4) {
        return __ldg(p);
    } else {
        typename Rec<D>::type r;
        r.a = __ldg(&p->a);
        r.b = __ldg(&p->b);
        return r;
    }
}

// asynchronous global -> shared copy of one record (cp.async, LDGSTS in SASS): no register staging,
// completion is awaited per thread with cp_async_wait<N>() (N = groups still allowed in flight)
template <int D> __device__ __forceinline__ void rec_cp_async(typename Rec<D>::type *smem_dst,
                                                              const typename Rec<D>::type *gmem_src) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    if constexpr (sizeof(typename Rec<D>::type) == 8) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(gmem_src));
    } else if constexpr (sizeof(typename Rec<D>::type) == 16) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(gmem_src));
    } else {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(gmem_src));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst + 16),
                     "l"(reinterpret_cast<const char *>(gmem_src) + 16));
    }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// --------------------------------------------------------------------------------------
// internal entry points (one per .cu file), wrapped by abi.cu
// --------------------------------------------------------------------------------------
int cloud_build(const float *pts, int64_t n, int d, int points_per_cell, void *ws, size_t ws_bytes,
                cudaStream_t st);
int bounding_balls(const float *verts, int64_t S, int K, int d, float *centers, float *radii,
                   cudaStream_t st);
size_t covering_workspace_bytes(int64_t S, int64_t R, int d);
int covering_radius(const void *cloud_ws, int64_t n, int d, const float *verts, int64_t S, int K,
                    const float *weights, int64_t R, const float *samples, const float *centers,
                    const float *radii, float *out_min_dist2, int64_t *out_cand_count,
                    unsigned long long *out_evals, void *ws, size_t ws_bytes, cudaStream_t st);
int covering_bricks(int64_t R, int d, int32_t *out_groups, int capacity, int *bricks_per_block);
int covering_plan(const void *cloud_ws, int64_t n, int d, const float *centers, const float *radii,
                  int64_t S, int32_t *out_tested, cudaStream_t st);
int bounding_balls_f64(const double *verts, int64_t S, int K, int d, double *centers, double *radii,
                       cudaStream_t st);
size_t covering_workspace_bytes_f64(int64_t S, int d);
int covering_radius_f64(const void *cloud_ws, const double *pts, int64_t n, int d, const double *verts, int64_t S,
                        int K, const double *weights, int64_t R, const double *centers, const double *radii,
                        double *out_min_dist2, int64_t *out_cand_count, unsigned long long *out_evals, void *ws,
                        size_t ws_bytes, cudaStream_t st);
int face_max_f64(const double *min_dist2, int64_t S, int64_t R, const int32_t *support, int K, double *out,
                 cudaStream_t st);
int face_max(const float *min_dist2, int64_t S, int64_t R, const int32_t *support, int K, float *out,
             cudaStream_t st);
size_t fps_workspace_bytes(int64_t n, int d, int64_t n_lms);
int fps_grid(const void *cloud_ws, const float *pts, int64_t n, int d, int64_t n_lms, int64_t start_idx,
             int64_t *out_idx, void *ws, size_t ws_bytes, cudaStream_t st);
int fps(const float *pts, int64_t n, int d, int64_t n_lms, int64_t start_idx, int64_t *out_idx,
        void *ws, size_t ws_bytes, cudaStream_t st);

}  // namespace flood

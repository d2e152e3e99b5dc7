// Exact farthest-point sampling as one persistent cooperative kernel.
//
// Replaces the reference's fpsample.bucket_fps_kdline_sampling call (flooder/core.py:337-342:
// device->host copy of the whole cloud, single-socket KD-bucket FPS in Rust, host->device copy of
// the indices).  Bucket-FPS is an exact acceleration of the FPS recurrence, so this kernel
// evaluates the recurrence itself:
//
//     idx[0] = start;   idx[k+1] = argmax_i min_{j<=k} |p_i - p_idx[j]|^2     (first maximum wins)
//
// Every iteration is a grid-wide fused min-update + argmax: each thread updates the running
// minima of the points it owns (registers when the cloud fits the chip's register files,
// otherwise a global scratch array), the (distance, index) argmax is reduced with warp shuffles,
// then across the CTA through shared memory, then across the grid with one 64-bit atomicMax on
// the key  float_bits(min_d2) << 32 | ~index  (larger distance wins, ties go to the smaller
// index), followed by a cooperative-groups grid sync.  Squared distances are summed in
// coordinate order WITHOUT FMA contraction so the chosen indices equal a scalar CPU
// implementation bit for bit.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace flood {
namespace {

struct FpsParams {
    const float *pts;
    long long n, n_lms, start;
    long long *out_idx;
    unsigned long long *best;  // [n_lms] zero-initialised argmax slots
    float *mind;               // [n] scratch (streaming mode only)
    unsigned *arrive;          // [n_lms] arrival counters (counter barrier only)
    int custom_barrier;        // 0 (default): cooperative-groups grid.sync() (4.1 us / iteration at 148
                               // CTAs); 1: one arrival counter per iteration polled by one thread per
                               // CTA (3.6 us, experimental)
};

template <int D>
__device__ __forceinline__ float sqdist_unfused(const float (&p)[D], const float (&q)[D]) {
    float t = __fsub_rn(p[0], q[0]);
    float s = __fmul_rn(t, t);
#pragma unroll
    for (int a = 1; a < D; ++a) {
        t = __fsub_rn(p[a], q[a]);
        s = __fadd_rn(s, __fmul_rn(t, t));
    }
    return s;
}

__device__ __forceinline__ unsigned long long make_key(float d2, long long idx) {
    return ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned long long)(0xffffffffu - (unsigned)idx);
}

// grid-wide argmax of `key`; returns the winner to every thread
__device__ __forceinline__ unsigned long long grid_argmax(unsigned long long key, const FpsParams &P,
                                                          long long k, unsigned long long *smem,
                                                          cg::grid_group &grid) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other > key ? other : key;
    }
    if (lane == 0) smem[warp] = key;
    __syncthreads();
    if (warp == 0) {
        unsigned long long v = lane < nw ? smem[lane] : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
            v = other > v ? other : v;
        }
        if (lane == 0) {
            atomicMax(P.best + k, v);
            if (P.custom_barrier) {
                __threadfence();
                atomicAdd(P.arrive + k, 1u);
                volatile unsigned *flag = P.arrive + k;
                while (*flag < gridDim.x) { }
                __threadfence();
            }
        }
    }
    if (P.custom_barrier) __syncthreads();
    else grid.sync();
    return __ldcg(P.best + k);
}

// PPT > 0: the thread's points live in registers.  PPT == 0: streaming mode.
template <int D, int PPT>
__global__ void __launch_bounds__(1024, 1) fps_kernel(const FpsParams P) {
    __shared__ unsigned long long smem[32];
    cg::grid_group grid = cg::this_grid();
    const long long gsize = (long long)gridDim.x * blockDim.x;
    const long long gtid = blockIdx.x * (long long)blockDim.x + threadIdx.x;

    float x[PPT > 0 ? PPT : 1][D], md[PPT > 0 ? PPT : 1];
    if (PPT > 0) {
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            const long long i = gtid + j * gsize;
            md[j] = INFINITY;
#pragma unroll
            for (int a = 0; a < D; ++a) x[j][a] = i < P.n ? P.pts[i * D + a] : 0.f;
        }
    } else {
        for (long long i = gtid; i < P.n; i += gsize) P.mind[i] = INFINITY;
    }

    long long cur = P.start;
    for (long long k = 0; k < P.n_lms; ++k) {
        if (gtid == 0) P.out_idx[k] = cur;
        if (k + 1 == P.n_lms) break;
        float q[D];
#pragma unroll
        for (int a = 0; a < D; ++a) q[a] = __ldg(P.pts + cur * D + a);
        unsigned long long key = 0ull;
        if (PPT > 0) {
#pragma unroll
            for (int j = 0; j < PPT; ++j) {
                const long long i = gtid + j * gsize;
                md[j] = fminf(md[j], sqdist_unfused<D>(x[j], q));
                const unsigned long long kj = make_key(md[j], i);
                if (i < P.n && kj > key) key = kj;
            }
        } else {
            for (long long i = gtid; i < P.n; i += gsize) {
                float p[D];
#pragma unroll
                for (int a = 0; a < D; ++a) p[a] = P.pts[i * D + a];
                const float m = fminf(P.mind[i], sqdist_unfused<D>(p, q));
                P.mind[i] = m;
                const unsigned long long ki = make_key(m, i);
                if (ki > key) key = ki;
            }
        }
        const unsigned long long win = grid_argmax(key, P, k, smem, grid);
        cur = (long long)(0xffffffffu - (unsigned)(win & 0xffffffffull));
    }
}

// ---------------------------------------------------------------------------------------------
// Bucketed exact FPS on the cell grid of a prepared cloud (cloud.cu).
//
// Same recurrence and the same arithmetic as fps_kernel, but an iteration only touches the cells
// that can change: a point's running minimum can drop only if the new landmark q is closer than
// the cell's current maximum, so a cell whose box is at least sqrt(cell_max) away from q is
// skipped.  This is the pruning idea of the reference's bucket-FPS (fpsample's KD-buckets,
// flooder/core.py:337-342) on the uniform grid the covering kernel already needs.  After k
// landmarks only O(1/k) of the cells are live, so the total work drops from N * n_lms point
// updates to a few tens of N, and the iteration is bound by the grid-wide argmax + sync.
//
// Every warp owns a fixed, strided set of cells (neighbouring cells go to different warps, so the
// live region around q is spread over the whole chip).  The per-cell argmax keys
// (float_bits(max min_d2) << 32 | ~original index) stay in registers for the whole run:
// lane l of a warp holds the cells of slots j * 32 + l.
// ---------------------------------------------------------------------------------------------
struct FpsGridParams {
    const GridParams *gp;
    const int *cell_start;
    const int *perm;
    const void *points;   // cell-sorted records
    const float *pts;     // original order (landmark coordinates are read from here)
    float *mind;          // [n] running minima, cell-sorted order
    long long n, n_lms, start;
    long long *out_idx;
    unsigned long long *best;
    unsigned *arrive;
    int custom_barrier;
};

template <int D, int KPL>
__global__ void __launch_bounds__(1024, 1) fps_grid_kernel(const FpsGridParams G) {
    using RecT = typename Rec<D>::type;
    __shared__ unsigned long long smem[32];
    cg::grid_group grid = cg::this_grid();
    const GridParams gp = *G.gp;
    const RecT *__restrict__ points = reinterpret_cast<const RecT *>(G.points);
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long gw = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long gsize = (long long)gridDim.x * blockDim.x;
    const long long gtid = blockIdx.x * (long long)blockDim.x + threadIdx.x;

    for (long long i = gtid; i < G.n; i += gsize) G.mind[i] = INFINITY;

    // this lane's cells: packed cell coordinates (-1 = no cell / empty cell) and argmax keys
    int cell_xyz[KPL];
    unsigned long long key[KPL];
#pragma unroll
    for (int j = 0; j < KPL; ++j) {
        const long long c = gw + (long long)(j * 32 + lane) * nwarps;
        cell_xyz[j] = -1;
        key[j] = 0ull;
        if (c < gp.ncells && G.cell_start[c + 1] > G.cell_start[c]) {
            const int ix = (int)(c % gp.n[0]), iy = (int)((c / gp.n[0]) % gp.n[1]);
            // axes beyond the third are ignored here: the distance to the (x, y, z) footprint of a
            // cell is still a lower bound of the distance to the cell
            const int iz = (int)((c / ((long long)gp.n[0] * gp.n[1])) % gp.n[2]);
            cell_xyz[j] = ix | (iy << 10) | (iz << 20);
            key[j] = 0x7f800000ull << 32;   // +inf: every non-empty cell is live in iteration 0
        }
    }
    grid.sync();

    FpsParams P;   // argmax plumbing shared with the brute-force kernel
    P.best = G.best;
    P.arrive = G.arrive;
    P.custom_barrier = G.custom_barrier;
    const float h2 = gp.h * gp.h;

    long long cur = G.start;
    for (long long k = 0; k < G.n_lms; ++k) {
        if (gtid == 0) G.out_idx[k] = cur;
        if (k + 1 == G.n_lms) break;
        float q[D];
#pragma unroll
        for (int a = 0; a < D; ++a) q[a] = __ldg(G.pts + cur * D + a);
        // landmark in cell coordinates (unused axes sit inside their single cell)
        const float gq0 = cell_coord(q[0], gp.origin[0], gp.inv_h);
        const float gq1 = D > 1 ? cell_coord(q[D > 1 ? 1 : 0], gp.origin[1], gp.inv_h) : 0.5f;
        const float gq2 = D > 2 ? cell_coord(q[D > 2 ? 2 : 0], gp.origin[2], gp.inv_h) : 0.5f;
#pragma unroll
        for (int j = 0; j < KPL; ++j) {
            bool live = false;
            if (cell_xyz[j] >= 0) {
                const float fx = (float)(cell_xyz[j] & 1023), fy = (float)((cell_xyz[j] >> 10) & 1023);
                const float fz = (float)(cell_xyz[j] >> 20);
                // distance from q to the cell box, in cells; 1e-3 cells of slack cover the
                // float32 rounding of the binning, the factor below the rounding of this test
                const float dx = fmaxf(0.f, fmaxf(fx - gq0, gq0 - (fx + 1.f)) - 1e-3f);
                const float dy = fmaxf(0.f, fmaxf(fy - gq1, gq1 - (fy + 1.f)) - 1e-3f);
                const float dz = fmaxf(0.f, fmaxf(fz - gq2, gq2 - (fz + 1.f)) - 1e-3f);
                const float box2 = (dx * dx + dy * dy + dz * dz) * h2;
                live = box2 * 0.9999f < __uint_as_float((unsigned)(key[j] >> 32));
            }
            unsigned todo = __ballot_sync(0xffffffffu, live);
            while (todo) {
                const int b = __ffs(todo) - 1;
                todo &= todo - 1;
                const long long c = gw + (long long)(j * 32 + b) * nwarps;
                const int cs = __ldg(G.cell_start + c), ce = __ldg(G.cell_start + c + 1);
                unsigned long long best = 0ull;
                for (int i = cs + lane; i < ce; i += 32) {
                    float p[D];
                    rec_unpack<D>(points[i], p);
                    const float old = G.mind[i];
                    const float m = fminf(old, sqdist_unfused<D>(p, q));
                    if (m < old) G.mind[i] = m;
                    const unsigned long long ki = make_key(m, (long long)__ldg(G.perm + i));
                    if (ki > best) best = ki;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
                    best = other > best ? other : best;
                }
                if (lane == b) key[j] = best;
            }
        }
        unsigned long long mine = 0ull;
#pragma unroll
        for (int j = 0; j < KPL; ++j) mine = key[j] > mine ? key[j] : mine;
        const unsigned long long win = grid_argmax(mine, P, k, smem, grid);
        cur = (long long)(0xffffffffu - (unsigned)(win & 0xffffffffull));
    }
}

struct FpsLayout {
    int64_t off_best, off_arrive, off_mind, total;
};

FpsLayout fps_layout(int64_t n, int64_t n_lms) {
    FpsLayout L;
    int64_t o = 0;
    L.off_best = o;    o = align_up(o + n_lms * 8, 256);
    L.off_arrive = o;  o = align_up(o + n_lms * 4, 256);
    L.off_mind = o;    o = align_up(o + n * 4, 256);
    L.total = o;
    return L;
}

template <int D, int PPT>
int launch_fps(FpsParams &P, cudaStream_t st, bool probe_only, long long *capacity) {
    auto kern = fps_kernel<D, PPT>;
    const int threads = 1024;
    int per_sm = 0;
    FLOOD_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0));
    if (per_sm < 1) return set_error(FLOOD_E_CUDA, "fps_kernel<%d,%d> does not fit on an SM", D, PPT);
    const int grid = device_sm_count() * per_sm;
    if (capacity) *capacity = (long long)grid * threads * (PPT > 0 ? PPT : 0);
    if (probe_only) return FLOOD_OK;
    void *args[] = {(void *)&P};
    const bool timed = get_option("time_kernels", 0) != 0;
    if (timed) kernel_timer_start("fps", st);
    FLOOD_CUDA_CHECK(cudaLaunchCooperativeKernel((void *)kern, dim3(grid), dim3(threads), args, 0, st));
    count_launches(1);
    if (timed) kernel_timer_stop("fps", st);
    return FLOOD_OK;
}

template <int D>
int dispatch_fps(FpsParams &P, cudaStream_t st) {
    long long cap = 0;
    const int force_stream = get_option("fps_stream", 0);
    if (!force_stream) {
        int rc;
        if ((rc = launch_fps<D, 1>(P, st, true, &cap)) != FLOOD_OK) return rc;
        if (P.n <= cap) return launch_fps<D, 1>(P, st, false, nullptr);
        if ((rc = launch_fps<D, 2>(P, st, true, &cap)) != FLOOD_OK) return rc;
        if (P.n <= cap) return launch_fps<D, 2>(P, st, false, nullptr);
        if ((rc = launch_fps<D, 4>(P, st, true, &cap)) != FLOOD_OK) return rc;
        if (P.n <= cap) return launch_fps<D, 4>(P, st, false, nullptr);
        if constexpr (D <= 4) {
            if ((rc = launch_fps<D, 8>(P, st, true, &cap)) != FLOOD_OK) return rc;
            if (P.n <= cap) return launch_fps<D, 8>(P, st, false, nullptr);
        }
    }
    return launch_fps<D, 0>(P, st, false, nullptr);
}

template <int D, int KPL>
int launch_fps_grid(FpsGridParams &G, long long ncells_bound, cudaStream_t st, bool *launched) {
    auto kern = fps_grid_kernel<D, KPL>;
    const int threads = 1024;
    int per_sm = 0;
    FLOOD_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0));
    if (per_sm < 1) return set_error(FLOOD_E_CUDA, "fps_grid_kernel<%d,%d> does not fit on an SM", D, KPL);
    const int grid = device_sm_count();   // one CTA per SM keeps the barrier small
    const long long capacity = (long long)grid * (threads / 32) * 32 * KPL;
    *launched = false;
    if (ncells_bound > capacity) return FLOOD_OK;
    void *args[] = {(void *)&G};
    const bool timed = get_option("time_kernels", 0) != 0;
    if (timed) kernel_timer_start("fps", st);
    FLOOD_CUDA_CHECK(cudaLaunchCooperativeKernel((void *)kern, dim3(grid), dim3(threads), args, 0, st));
    count_launches(1);
    if (timed) kernel_timer_stop("fps", st);
    *launched = true;
    return FLOOD_OK;
}

template <int D>
int dispatch_fps_grid(FpsGridParams &G, long long ncells_bound, cudaStream_t st) {
    bool ok = false;
    int rc;
    if ((rc = launch_fps_grid<D, 1>(G, ncells_bound, st, &ok)) != FLOOD_OK || ok) return rc;
    if ((rc = launch_fps_grid<D, 2>(G, ncells_bound, st, &ok)) != FLOOD_OK || ok) return rc;
    if ((rc = launch_fps_grid<D, 4>(G, ncells_bound, st, &ok)) != FLOOD_OK || ok) return rc;
    if ((rc = launch_fps_grid<D, 8>(G, ncells_bound, st, &ok)) != FLOOD_OK || ok) return rc;
    if ((rc = launch_fps_grid<D, 16>(G, ncells_bound, st, &ok)) != FLOOD_OK || ok) return rc;
    return set_error(FLOOD_E_UNSUPPORTED, "fps_grid: %lld cells exceed the per-lane key capacity", ncells_bound);
}

}  // namespace

int fps_grid(const void *cloud_ws, const float *pts, int64_t n, int d, int64_t n_lms, int64_t start_idx,
             int64_t *out_idx, void *ws, size_t ws_bytes, cudaStream_t st) {
    if (!cloud_ws || !pts || !out_idx || !ws || n < 1 || n_lms < 1 || n_lms > n || start_idx < 0 ||
        start_idx >= n || d < 2 || d > FLOOD_MAX_DIM)
        return set_error(FLOOD_E_INVALID, "fps_grid: bad arguments (n=%lld d=%d n_lms=%lld start=%lld)",
                         (long long)n, d, (long long)n_lms, (long long)start_idx);
    const FpsLayout L = fps_layout(n, n_lms);
    if ((int64_t)ws_bytes < L.total)
        return set_error(FLOOD_E_WORKSPACE, "fps_grid: workspace %zu < %lld bytes", ws_bytes, (long long)L.total);
    const CloudLayout C = cloud_layout(n, d);
    const char *cbase = static_cast<const char *>(cloud_ws);
    char *base = static_cast<char *>(ws);
    FpsGridParams G;
    G.gp = reinterpret_cast<const GridParams *>(cbase + C.off_grid);
    G.cell_start = reinterpret_cast<const int *>(cbase + C.off_cell_start);
    G.perm = reinterpret_cast<const int *>(cbase + C.off_perm);
    G.points = cbase + C.off_points;
    G.pts = pts;
    G.mind = reinterpret_cast<float *>(base + L.off_mind);
    G.n = n;
    G.n_lms = n_lms;
    G.start = start_idx;
    G.out_idx = reinterpret_cast<long long *>(out_idx);
    G.best = reinterpret_cast<unsigned long long *>(base + L.off_best);
    G.arrive = reinterpret_cast<unsigned *>(base + L.off_arrive);
    G.custom_barrier = get_option("fps_barrier", 0);
    FLOOD_CUDA_CHECK(cudaMemsetAsync(base, 0, (size_t)L.off_mind, st));
    // the number of cells lives in device memory: fetch it (one small synchronising copy)
    GridParams host_gp;
    FLOOD_CUDA_CHECK(cudaMemcpyAsync(&host_gp, G.gp, sizeof(GridParams), cudaMemcpyDeviceToHost, st));
    FLOOD_CUDA_CHECK(cudaStreamSynchronize(st));
    if (host_gp.npts != n || host_gp.d != d)
        return set_error(FLOOD_E_INVALID, "fps_grid: cloud workspace was built for (%lld, %d), not (%lld, %d)",
                         (long long)host_gp.npts, host_gp.d, (long long)n, d);
    const long long bound = host_gp.ncells;
    switch (d) {
        case 2: return dispatch_fps_grid<2>(G, bound, st);
        case 3: return dispatch_fps_grid<3>(G, bound, st);
        case 4: return dispatch_fps_grid<4>(G, bound, st);
        case 5: return dispatch_fps_grid<5>(G, bound, st);
        case 6: return dispatch_fps_grid<6>(G, bound, st);
        case 7: return dispatch_fps_grid<7>(G, bound, st);
        case 8: return dispatch_fps_grid<8>(G, bound, st);
    }
    return set_error(FLOOD_E_UNSUPPORTED, "fps_grid: d=%d", d);
}

size_t fps_workspace_bytes(int64_t n, int d, int64_t n_lms) {
    (void)d;
    return (size_t)fps_layout(n < 1 ? 1 : n, n_lms < 1 ? 1 : n_lms).total;
}

int fps(const float *pts, int64_t n, int d, int64_t n_lms, int64_t start_idx, int64_t *out_idx,
        void *ws, size_t ws_bytes, cudaStream_t st) {
    if (!pts || !out_idx || !ws || n < 1 || n_lms < 1 || n_lms > n || start_idx < 0 || start_idx >= n ||
        d < 1 || d > FLOOD_MAX_DIM)
        return set_error(FLOOD_E_INVALID, "fps: bad arguments (n=%lld d=%d n_lms=%lld start=%lld)",
                         (long long)n, d, (long long)n_lms, (long long)start_idx);
    if (n >= (int64_t(1) << 32) - 1) return set_error(FLOOD_E_UNSUPPORTED, "fps: n too large");
    const FpsLayout L = fps_layout(n, n_lms);
    if ((int64_t)ws_bytes < L.total)
        return set_error(FLOOD_E_WORKSPACE, "fps: workspace %zu < %lld bytes", ws_bytes, (long long)L.total);
    char *base = static_cast<char *>(ws);
    FpsParams P;
    P.pts = pts;
    P.n = n;
    P.n_lms = n_lms;
    P.start = start_idx;
    P.out_idx = reinterpret_cast<long long *>(out_idx);
    P.best = reinterpret_cast<unsigned long long *>(base + L.off_best);
    P.arrive = reinterpret_cast<unsigned *>(base + L.off_arrive);
    P.mind = reinterpret_cast<float *>(base + L.off_mind);
    P.custom_barrier = get_option("fps_barrier", 0);
    FLOOD_CUDA_CHECK(cudaMemsetAsync(base, 0, (size_t)L.off_mind, st));
    switch (d) {
        case 1: return dispatch_fps<1>(P, st);
        case 2: return dispatch_fps<2>(P, st);
        case 3: return dispatch_fps<3>(P, st);
        case 4: return dispatch_fps<4>(P, st);
        case 5: return dispatch_fps<5>(P, st);
        case 6: return dispatch_fps<6>(P, st);
        case 7: return dispatch_fps<7>(P, st);
        case 8: return dispatch_fps<8>(P, st);
    }
    return set_error(FLOOD_E_UNSUPPORTED, "fps: d=%d", d);
}

}  // namespace flood

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
    unsigned *arrive;          // [n_lms] arrival counters (custom barrier only)
    int custom_barrier;
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

}  // namespace

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

// float64 variants of the covering path (the reference runs its Triton kernels in float64 when the
// inputs are float64: flooder/triton_kernels.py:226-229, flooder/core.py:116-123; tested by
// tests/test_flooder.py:214-246).
//
// B200's FP64 rate is a small fraction of its FP32 rate and float64 clouds are the exception, so
// this path is deliberately plain: same work decomposition as the float32 kernel (cell rows of the
// ball -> runs of the cell-sorted cloud, (simplex, chunk) items from an atomic queue, tile of
// in-ball candidates in shared memory, atomicMin merge), no pruning, no packed arithmetic.  The cell
// grid is the one built from the float32-rounded cloud (flood_cloud_build_f32): it only enumerates
// candidates -- the row selection is inflated by far more than a float32 rounding -- while the ball
// predicate, the sample points and the distances are evaluated in float64 on the original
// coordinates (gathered through the grid's permutation).
#include "covering_kernels.cuh"

namespace flood {
namespace {

// ---- bounding balls (flooder/core.py:156-172), float64 --------------------------------------------
__global__ void bounding_balls_f64_kernel(const double *__restrict__ verts, int64_t S, int K, int d,
                                          double *__restrict__ centers, double *__restrict__ radii) {
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= S) return;
    const double *v = verts + s * K * d;
    double best = -1.0;
    int b0 = 0, b1 = 0;
    for (int i = 0; i < K; ++i)
        for (int j = 0; j < K; ++j) {
            double acc = 0.0;
            for (int a = 0; a < d; ++a) {
                const double t = __dsub_rn(v[i * d + a], v[j * d + a]);
                acc = __dadd_rn(acc, __dmul_rn(t, t));
            }
            const double dist = __dsqrt_rn(acc);
            if (dist > best) { best = dist; b0 = i; b1 = j; }
        }
    double c[FLOOD_MAX_DIM];
    for (int a = 0; a < d; ++a) {
        c[a] = __dmul_rn(__dadd_rn(v[b0 * d + a], v[b1 * d + a]), 0.5);
        centers[s * d + a] = c[a];
    }
    double far = 0.0;
    for (int k = 0; k < K; ++k) {
        double acc = 0.0;
        for (int a = 0; a < d; ++a) {
            const double t = __dsub_rn(v[k * d + a], c[a]);
            acc = __dadd_rn(acc, __dmul_rn(t, t));
        }
        far = fmax(far, __dsqrt_rn(acc));
    }
    const double factor = (K - 1) > 1 ? 1.42 : 1.01;
    radii[s] = __dadd_rn(__dmul_rn(far, factor), 1e-3);
}

__global__ void narrow_balls_kernel(const double *__restrict__ c64, const double *__restrict__ r64, int64_t S,
                                    int d, float *__restrict__ c32, float *__restrict__ r32) {
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= S) return;
    for (int a = 0; a < d; ++a) c32[s * d + a] = (float)c64[s * d + a];
    // rounded up: the float32 ball only selects cell rows and must contain the float64 ball
    r32[s] = __double2float_ru(r64[s] * (1.0 + 1e-6));
}

__global__ void fill_inf_f64_kernel(double *p, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        p[i] = INFINITY;
}

struct F64Params {
    CoverParams base;        // grid, cell_start, tested, item_base, queue, S, R, K, chunk (float32 balls inside)
    const int *perm;         // cell-sorted position -> original index
    const double *pts;       // [n,D] original order
    const double *verts;     // [S,K,D]
    const double *weights;   // [R,K]
    const double *centers;   // [S,D]
    const double *radii;     // [S]
    double *out;             // [S,R]
};

constexpr int kF64Threads = 256;
constexpr int kF64Tile = 512;      // candidates per shared-memory tile
constexpr int kF64Spt = 4;         // samples per thread and pass

template <int D>
__global__ void __launch_bounds__(kF64Threads) cover_f64_kernel(const F64Params Q) {
    constexpr int G = grid_axes(D);
    const CoverParams &P = Q.base;
    __shared__ double tile[kF64Tile][D];
    __shared__ int run_start[kF64Threads], run_pos[kF64Threads + 1];
    __shared__ int warp_sums[32];
    __shared__ int s_fill;
    __shared__ long long s_item[2];
    __shared__ GridParams s_gp;
    const int tid = threadIdx.x, lane = tid & 31;
    const int NT = kF64Threads;
    const long long total_items = P.item_base[P.S];
    if (tid == 0) { s_gp = *P.gp; s_fill = 0; }

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            const unsigned long long g = atomicAdd(P.queue, 1ull);
            if (g >= (unsigned long long)total_items) {
                s_item[0] = -1;
            } else {
                long long lo = 0, hi = P.S;
                while (hi - lo > 1) {
                    const long long mid = (lo + hi) >> 1;
                    if (P.item_base[mid] <= (long long)g) lo = mid; else hi = mid;
                }
                s_item[0] = lo;
                s_item[1] = (long long)g - P.item_base[lo];
            }
        }
        __syncthreads();
        const long long s = s_item[0];
        if (s < 0) break;
        const long long chunk_j = s_item[1];
        float c32[G];
#pragma unroll
        for (int a = 0; a < G; ++a) c32[a] = __ldg(P.centers + s * D + a);
        const BallCells<G> bc = ball_cells<G>(c32, __ldg(P.radii + s), s_gp);
        double c[D];
#pragma unroll
        for (int a = 0; a < D; ++a) c[a] = Q.centers[s * D + a];
        const double rad = Q.radii[s], r2 = rad * rad;
        const long long tested = P.tested[s];
        const long long nch = P.item_base[s + 1] - P.item_base[s];
        const int win_lo = (int)(chunk_j * tested / nch);
        const int win_hi = (int)((chunk_j + 1) * tested / nch);

        // sweep of the current tile: every thread takes kF64Spt samples per pass
        auto sweep = [&](int n) {
            for (long long r0 = 0; r0 < P.R; r0 += (long long)NT * kF64Spt) {
                double x[kF64Spt][D], m[kF64Spt];
#pragma unroll
                for (int q = 0; q < kF64Spt; ++q) {
                    const long long r = r0 + (long long)q * NT + tid;
                    m[q] = INFINITY;
#pragma unroll
                    for (int a = 0; a < D; ++a) x[q][a] = c[a];
                    if (r < P.R) {
                        // x = sum_k w[r,k] * v[s,k,:] (core.py:188), FMA chain over k ascending
                        const double *w = Q.weights + r * P.K;
                        const double *v = Q.verts + s * P.K * D;
#pragma unroll
                        for (int a = 0; a < D; ++a) x[q][a] = __dmul_rn(w[0], v[a]);
                        for (int k = 1; k < P.K; ++k) {
#pragma unroll
                            for (int a = 0; a < D; ++a) x[q][a] = fma(w[k], v[k * D + a], x[q][a]);
                        }
                    }
                }
                for (int j = 0; j < n; ++j) {
                    double p[D];
#pragma unroll
                    for (int a = 0; a < D; ++a) p[a] = tile[j][a];
#pragma unroll
                    for (int q = 0; q < kF64Spt; ++q) {
                        double t = x[q][0] - p[0];
                        double acc = t * t;
#pragma unroll
                        for (int a = 1; a < D; ++a) {
                            t = x[q][a] - p[a];
                            acc = fma(t, t, acc);
                        }
                        m[q] = fmin(m[q], acc);
                    }
                }
#pragma unroll
                for (int q = 0; q < kF64Spt; ++q) {
                    const long long r = r0 + (long long)q * NT + tid;
                    // non-negative doubles order like their bit patterns
                    if (r < P.R && m[q] < INFINITY)
                        atomicMin(reinterpret_cast<unsigned long long *>(Q.out + s * P.R + r),
                                  (unsigned long long)__double_as_longlong(m[q]));
                }
            }
        };

        int fill = 0;
        long long accepted = 0;
        int carry = 0;
        for (int rb = 0; rb < bc.nrows; rb += NT) {
            if (carry >= win_hi) break;
            const int row = rb + tid;
            int a0 = 0, len = 0;
            if (row < bc.nrows) row_run<G>(bc, row, s_gp, P.cell_start, a0, len);
            int batch_total;
            const int off = carry + block_exclusive_scan(len, warp_sums, batch_total);
            carry += batch_total;
            if (carry <= win_lo) continue;
            const int s0 = max(off, win_lo), s1 = min(off + len, win_hi);
            int total2;
            const int pos2 = block_exclusive_scan(max(0, s1 - s0), warp_sums, total2);
            run_start[tid] = a0 + (s0 - off);
            run_pos[tid] = pos2;
            if (tid == 0) run_pos[NT] = total2;
            __syncthreads();
            for (int base = 0; base < total2; base += NT) {
                if (fill + NT > kF64Tile) {
                    accepted += fill;
                    sweep(fill);
                    __syncthreads();
                    if (tid == 0) s_fill = 0;
                    __syncthreads();
                    fill = 0;
                }
                const int q = base + tid;
                bool pass = false;
                double p[D];
                if (q < total2) {
                    int lo = 0, hi = NT;   // last run with run_pos <= q
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        if (run_pos[mid] <= q) lo = mid; else hi = mid;
                    }
                    const long long idx = __ldg(Q.perm + run_start[lo] + (q - run_pos[lo]));
                    // the reference predicate in float64 (triton_kernels.py:137-148)
                    double acc = 0.0;
#pragma unroll
                    for (int a = 0; a < D; ++a) {
                        p[a] = Q.pts[idx * D + a];
                        const double t = p[a] - c[a];
                        acc = a == 0 ? t * t : fma(t, t, acc);
                    }
                    pass = acc <= r2;
                }
                const unsigned ballot = __ballot_sync(0xffffffffu, pass);
                int wbase = 0;
                if (lane == 0 && ballot) wbase = atomicAdd(&s_fill, __popc(ballot));
                wbase = __shfl_sync(0xffffffffu, wbase, 0);
                if (pass) {
                    const int slot = wbase + __popc(ballot & ((1u << lane) - 1u));
#pragma unroll
                    for (int a = 0; a < D; ++a) tile[slot][a] = p[a];
                }
                __syncthreads();
                fill = s_fill;
            }
        }
        if (fill > 0) {
            accepted += fill;
            sweep(fill);
            __syncthreads();
            if (tid == 0) s_fill = 0;
        }
        if (tid == 0 && accepted > 0) {
            if (P.cand_count) atomicAdd(reinterpret_cast<unsigned long long *>(P.cand_count + s),
                                        (unsigned long long)accepted);
            if (P.evals) atomicAdd(P.evals, (unsigned long long)accepted * (unsigned long long)P.R);
        }
    }
}

// per-face maxima, float64 (see face_max_kernel in balls.cu)
__global__ void face_max_f64_kernel(const double *__restrict__ min_dist2, int64_t R,
                                    const int32_t *__restrict__ support, int K, double *__restrict__ out) {
    extern __shared__ unsigned long long bins64[];
    const int64_t s = blockIdx.x;
    const int nb = support ? (1 << K) : 1;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) bins64[i] = 0ull;
    __syncthreads();
    const double *row = min_dist2 + s * R;
    for (int64_t r = threadIdx.x; r < R; r += blockDim.x)
        atomicMax(&bins64[support ? (support[r] & (nb - 1)) : 0], (unsigned long long)__double_as_longlong(row[r]));
    __syncthreads();
    if (support) {
        for (int m = 1 + threadIdx.x; m < nb; m += blockDim.x) {
            unsigned long long best = 0ull;
            for (int sub = m; sub; sub = (sub - 1) & m) best = max(best, bins64[sub]);
            out[s * (nb - 1) + (m - 1)] = sqrt(__longlong_as_double((long long)best));
        }
    } else if (threadIdx.x == 0) {
        out[s] = sqrt(__longlong_as_double((long long)bins64[0]));
    }
}

template <int D>
void launch_f64(const F64Params &Q, cudaStream_t st) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cover_f64_kernel<D>, kF64Threads, 0);
    if (per_sm < 1) per_sm = 1;
    cover_f64_kernel<D><<<device_sm_count() * per_sm, kF64Threads, 0, st>>>(Q);
}

}  // namespace

// cover_plan_kernel / cover_scan_kernel live in covering.cu
int covering_plan_items(CoverParams &P, int d, int64_t S, void *ws, cudaStream_t st);

int bounding_balls_f64(const double *verts, int64_t S, int K, int d, double *centers, double *radii,
                       cudaStream_t st) {
    if (S == 0) return FLOOD_OK;
    if (!verts || !centers || !radii || S < 0 || K < 1 || K > FLOOD_MAX_SIMPLEX_VERTS || d < 1 || d > FLOOD_MAX_DIM)
        return set_error(FLOOD_E_INVALID, "bounding_balls_f64: bad arguments (S=%lld, K=%d, d=%d)", (long long)S, K, d);
    const int threads = 128;
    bounding_balls_f64_kernel<<<(unsigned)((S + threads - 1) / threads), threads, 0, st>>>(verts, S, K, d, centers, radii);
    count_launches(1);
    FLOOD_LAUNCH_CHECK("bounding_balls_f64_kernel");
    return FLOOD_OK;
}

size_t covering_workspace_bytes_f64(int64_t S, int d) {
    if (S < 1) S = 1;
    // float32 copies of the balls (row selection) behind the float32 workspace
    return covering_workspace_bytes(S, 1, d) + (size_t)align_up(S * d * 4, 256) + (size_t)align_up(S * 4, 256);
}

int covering_radius_f64(const void *cloud_ws, const double *pts, int64_t n, int d, const double *verts, int64_t S,
                        int K, const double *weights, int64_t R, const double *centers, const double *radii,
                        double *out_min_dist2, int64_t *out_cand_count, unsigned long long *out_evals, void *ws,
                        size_t ws_bytes, cudaStream_t st) {
    if (S == 0) return FLOOD_OK;
    if (!cloud_ws || !pts || !verts || !weights || !centers || !radii || !out_min_dist2 || !ws || S < 0 || R < 1 ||
        n < 1 || d < 1 || d > FLOOD_MAX_DIM || K < 1 || K > FLOOD_MAX_SIMPLEX_VERTS)
        return set_error(FLOOD_E_INVALID, "covering_radius_f64: bad arguments (S=%lld R=%lld n=%lld d=%d K=%d)",
                         (long long)S, (long long)R, (long long)n, d, K);
    if (ws_bytes < covering_workspace_bytes_f64(S, d))
        return set_error(FLOOD_E_WORKSPACE, "covering_radius_f64: workspace %zu < %zu bytes", ws_bytes,
                         covering_workspace_bytes_f64(S, d));
    const CloudLayout C = cloud_layout(n, d);
    const char *cbase = static_cast<const char *>(cloud_ws);
    char *wbase = static_cast<char *>(ws);
    const size_t w32 = covering_workspace_bytes(S, 1, d);
    float *c32 = reinterpret_cast<float *>(wbase + w32);
    float *r32 = reinterpret_cast<float *>(wbase + w32 + align_up(S * d * 4, 256));

    F64Params Q = {};
    CoverParams &P = Q.base;
    P.gp = reinterpret_cast<const GridParams *>(cbase + C.off_grid);
    P.cell_start = reinterpret_cast<const int *>(cbase + C.off_cell_start);
    P.centers = c32;
    P.radii = r32;
    P.cand_count = reinterpret_cast<long long *>(out_cand_count);
    P.evals = out_evals;
    P.S = S;
    P.R = R;
    P.K = K;
    Q.perm = reinterpret_cast<const int *>(cbase + C.off_perm);
    Q.pts = pts;
    Q.verts = verts;
    Q.weights = weights;
    Q.centers = centers;
    Q.radii = radii;
    Q.out = out_min_dist2;

    const int threads = 128;
    narrow_balls_kernel<<<(unsigned)((S + threads - 1) / threads), threads, 0, st>>>(centers, radii, S, d, c32, r32);
    if (out_cand_count) FLOOD_CUDA_CHECK(cudaMemsetAsync(out_cand_count, 0, (size_t)S * 8, st));
    {
        const long long total = S * R;
        int blocks = (int)((total + 1023) / 1024);
        const int cap = device_sm_count() * 8;
        if (blocks > cap) blocks = cap;
        fill_inf_f64_kernel<<<blocks, 256, 0, st>>>(out_min_dist2, total);
    }
    const int rc = covering_plan_items(P, d, S, ws, st);   // tested[], item_base[] (scanned), queue reset
    if (rc != FLOOD_OK) return rc;
    switch (d) {
        case 1: launch_f64<1>(Q, st); break;
        case 2: launch_f64<2>(Q, st); break;
        case 3: launch_f64<3>(Q, st); break;
        case 4: launch_f64<4>(Q, st); break;
        case 5: launch_f64<5>(Q, st); break;
        case 6: launch_f64<6>(Q, st); break;
        case 7: launch_f64<7>(Q, st); break;
        case 8: launch_f64<8>(Q, st); break;
    }
    count_launches(3);
    FLOOD_LAUNCH_CHECK("cover_f64_kernel");
    return FLOOD_OK;
}

int face_max_f64(const double *min_dist2, int64_t S, int64_t R, const int32_t *support, int K, double *out,
                 cudaStream_t st) {
    if (S == 0) return FLOOD_OK;
    if (!min_dist2 || !out || S < 0 || R < 1 || K < 1 || K > FLOOD_MAX_SIMPLEX_VERTS || S > 2147483647LL)
        return set_error(FLOOD_E_INVALID, "face_max_f64: bad arguments (S=%lld, R=%lld, K=%d)", (long long)S,
                         (long long)R, K);
    const size_t smem = sizeof(unsigned long long) * (support ? (size_t(1) << K) : 1);
    face_max_f64_kernel<<<(unsigned)S, 256, smem, st>>>(min_dist2, R, support, K, out);
    count_launches(1);
    FLOOD_LAUNCH_CHECK("face_max_f64_kernel");
    return FLOOD_OK;
}

}  // namespace flood

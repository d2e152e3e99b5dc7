// Covering-radius kernel family: for every simplex s and sample point x_sr on it,
//     min_dist2[s, r] = min over the cloud points inside the simplex's bounding ball of |x_sr - p|^2.
//
// One fused persistent kernel replaces the reference's per-batch pipeline
//     compute_mask (dense bool mask)  ->  torch.nonzero  ->  compute_filtration (gather + atomic_min)
// (flooder/core.py:193-226, flooder/triton_kernels.py:12-45 and :99-158) and its
// `weights @ vertices` matmul (core.py:188):
//
//   * work item  = (simplex, chunk of its candidate stream, block of its samples), pulled from a
//     global atomic queue by persistent CTAs;
//   * the candidate stream of a simplex is the concatenation of the cell-row runs its ball
//     touches in the cell-sorted cloud (cloud.cu); a CTA enumerates the rows, prefix-sums the run
//     lengths and gathers its window of the stream with a binary search per point;
//   * every gathered point is tested against the ball (the reference predicate) and the
//     survivors are compacted with warp ballots into a shared-memory tile of float4 records;
//   * when the tile fills up, all warps sweep it: each thread keeps up to 8 sample points and their
//     running minima in registers, tile records are broadcast LDS.128 reads, the distance is the
//     direct difference form in FP32 (D sub, 1 mul, D-1 fma, 1 min per pair), issued as packed
//     FP32x2 instructions and 3-input minima (sweep_tile);
//   * by default the sweep is pruned exactly: a warp skips the tile records that are at least as
//     far from the box of its samples as its largest running minimum (sweep_tile_pruned), after
//     a seed pass over a sub-sampled stream has given every sample a finite bound;
//   * the minima are merged into min_dist2 with an unsigned atomicMin (non-negative floats order
//     like their bit patterns), because a simplex may be split over several chunks.
//
// The dense mask, the index lists and the (S, R, D) sample tensor never exist.
#include "common.cuh"

namespace flood {
namespace {

constexpr int kUnroll = 4;       // candidates per inner-loop trip

struct CoverParams {
    const GridParams *gp;
    const int *cell_start;
    const void *points;
    const float *verts;      // [S,K,D]
    const float *weights;    // [R,K]
    const float *samples;    // [S,R,D] or null
    const float *centers;    // [S,D]
    const float *radii;      // [S]
    float *out;              // [S,R]
    long long *cand_count;   // [S] or null
    unsigned long long *evals;  // or null
    int *tested;             // [S]     plan output
    long long *item_base;    // [S+1]   exclusive prefix of chunks per simplex
    unsigned long long *queue;
    unsigned long long *executed;   // evaluations actually performed (pruned sweeps skip some)
    int stream_stride;       // 1 = every stream position; k > 1 = every k-th (seed pass of the pruned mode)
    int count_work;          // add to cand_count / evals (exactly one pass per call does)
    long long S, R;
    int K;
    int nsb;                 // sample blocks per simplex
    int groups;              // ceil(R / 32) sample groups per simplex
    int groups_per_block;    // groups handled by one CTA pass (sample block)
    int tile_cap;            // candidate records per shared-memory tile
    int chunk;               // target tested points per chunk
    int rows_per_chunk_factor;  // chunk >= factor * (cell rows of the simplex)
};

// ---------------------------------------------------------------------------------------------
// geometry of a ball in cell coordinates
// ---------------------------------------------------------------------------------------------
struct BallCells {
    float gx, gy, gz, gr2;
    int iy0, iz0, nyb, nrows;
};

__device__ __forceinline__ int clamp_cell(float v, int n) {
    v = fminf(fmaxf(floorf(v), -1.0f), (float)n);
    return (int)v;
}

__device__ __forceinline__ BallCells ball_cells(const float *c, float r, int d, const GridParams &gp) {
    BallCells b;
    b.gx = cell_coord(c[0], gp.origin[0], gp.inv_h);
    b.gy = d > 1 ? cell_coord(c[1], gp.origin[1], gp.inv_h) : 0.5f;
    b.gz = d > 2 ? cell_coord(c[2], gp.origin[2], gp.inv_h) : 0.5f;
    // inflate: the cell mapping and the ball predicate are evaluated in float32
    const float gr = r * gp.inv_h * (1.0f + 1e-5f) + 2e-3f;
    b.gr2 = gr * gr;
    int iy0 = max(0, clamp_cell(b.gy - gr, gp.n[1]));
    int iy1 = min(gp.n[1] - 1, clamp_cell(b.gy + gr, gp.n[1]));
    int iz0 = max(0, clamp_cell(b.gz - gr, gp.n[2]));
    int iz1 = min(gp.n[2] - 1, clamp_cell(b.gz + gr, gp.n[2]));
    b.iy0 = iy0;
    b.iz0 = iz0;
    b.nyb = max(0, iy1 - iy0 + 1);
    b.nrows = b.nyb * max(0, iz1 - iz0 + 1);
    return b;
}

// run [a, a+len) of cell-sorted points covered by the ball in cell row `row` (rows are numbered
// in memory order: y fastest, then z)
__device__ __forceinline__ void row_run(const BallCells &b, int row, const GridParams &gp,
                                        const int *__restrict__ cell_start, int &a, int &len) {
    const int iy = b.iy0 + row % b.nyb;
    const int iz = b.iz0 + row / b.nyb;
    const float dy = fmaxf(0.f, fmaxf((float)iy - b.gy, b.gy - (float)(iy + 1)));
    const float dz = fmaxf(0.f, fmaxf((float)iz - b.gz, b.gz - (float)(iz + 1)));
    const float rem = b.gr2 - dy * dy - dz * dz;
    a = 0;
    len = 0;
    if (rem < 0.f) return;
    const float half = sqrtf(rem);
    const int ix0 = max(0, clamp_cell(b.gx - half, gp.n[0]));
    const int ix1 = min(gp.n[0] - 1, clamp_cell(b.gx + half, gp.n[0]));
    if (ix0 > ix1) return;
    const int base = (iz * gp.n[1] + iy) * gp.n[0];
    a = __ldg(cell_start + base + ix0);
    len = __ldg(cell_start + base + ix1 + 1) - a;
}

// ---------------------------------------------------------------------------------------------
// plan: tested[s] = length of the candidate stream of simplex s (one warp per simplex)
// ---------------------------------------------------------------------------------------------
__global__ void cover_plan_kernel(CoverParams P, int d) {
    const GridParams gp = *P.gp;
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (warp >= P.S) return;
    float c[FLOOD_MAX_DIM];
    for (int a = 0; a < d; ++a) c[a] = P.centers[warp * d + a];
    const BallCells b = ball_cells(c, P.radii[warp], d, gp);
    long long total = 0;
    for (int row = lane; row < b.nrows; row += 32) {
        int a, len;
        row_run(b, row, gp, P.cell_start, a, len);
        total += len;
    }
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    if (lane == 0) {
        P.tested[warp] = (int)total;
        // Number of chunks, stored in place and scanned by cover_scan_kernel.  Every chunk walks the
        // simplex's row list again, so simplices with many rows get proportionally longer chunks
        // (the walk stays a few per cent of the chunk's sweep).
        const long long chunk = max((long long)P.chunk, (long long)P.rows_per_chunk_factor * b.nrows);
        if (P.item_base) P.item_base[warp] = (total + chunk - 1) / chunk;
    }
}

// single-CTA exclusive scan (int64) of item_base[0..S) in place; item_base[S] = total
__global__ void cover_scan_kernel(long long *v, long long S) {
    __shared__ long long warp_sums[32];
    __shared__ long long carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (long long base = 0; base < S; base += blockDim.x) {
        const long long i = base + threadIdx.x;
        const long long val = i < S ? v[i] : 0;
        long long x = val;
        for (int o = 1; o < 32; o <<= 1) {
            long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            long long w = lane < nw ? warp_sums[lane] : 0;
            for (int o = 1; o < 32; o <<= 1) {
                long long y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const long long carry = carry_s;
        if (i < S) v[i] = carry + (warp > 0 ? warp_sums[warp - 1] : 0) + x - val;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry_s = carry + warp_sums[nw - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) v[S] = carry_s;
}

// 3-input minimum (FMNMX3 on sm_100a)
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// squared distance, direct difference form: (x0-p0)^2 rounded, then FMA accumulation
template <int D>
__device__ __forceinline__ float dist2(const float (&x)[D], const float (&p)[D]) {
    float t = x[0] - p[0];
    float acc = t * t;
#pragma unroll
    for (int a = 1; a < D; ++a) {
        t = x[a] - p[a];
        acc = fmaf(t, t, acc);
    }
    return acc;
}

// ---------------------------------------------------------------------------------------------
// block-wide exclusive scan of one int per thread; returns the exclusive prefix, total in `total`
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int block_exclusive_scan(int v, int *warp_sums, int &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    __syncthreads();  // protect warp_sums from the previous use
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    int prefix = 0, tot = 0;
    // nw <= 32: every thread folds the warp totals it needs (broadcast LDS, no third barrier)
    for (int w = 0; w < nw; ++w) {
        int s = warp_sums[w];
        if (w < warp) prefix += s;
        tot += s;
    }
    total = tot;
    return prefix + x - v;
}

// ---------------------------------------------------------------------------------------------
// the persistent evaluation kernel
// ---------------------------------------------------------------------------------------------

// Sweep of a padded tile by one warp that holds NT_ sample groups in registers: the hot loop.
// Two samples share one packed FP32x2 instruction (FADD2 / FMUL2 / FFMA2 take the candidate
// coordinate as a broadcast scalar operand), two candidates share one 3-input FMNMX3: per pair of
// samples and pair of candidates that is 2 x (D FADD2 + FMUL2 + (D-1) FFMA2) + 2 FMNMX3 issue slots
// for four evaluations.  An odd group is handled with the scalar form.  Each lane result is the
// same IEEE operation as the scalar form (x - p, round; *, round; fma, round), so the minima are
// bit-identical to a scalar evaluation.
template <int D, int NT_, int MAXT>
__device__ __forceinline__ void sweep_tile(const typename Rec<D>::type *__restrict__ tile, int npad,
                                           const float (&x)[MAXT][D], float (&m)[MAXT]) {
#pragma unroll 1
    for (int j = 0; j < npad; j += kUnroll) {
        float p[kUnroll][D];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) rec_unpack<D>(tile[j + u], p[u]);
#pragma unroll
        for (int u = 0; u < kUnroll; u += 2) {
#pragma unroll
            for (int t = 0; t + 1 < NT_; t += 2) {
                float2 acc[2];
#pragma unroll
                for (int v = 0; v < 2; ++v) {
                    float2 df = __fadd2_rn(make_float2(x[t][0], x[t + 1][0]),
                                           make_float2(-p[u + v][0], -p[u + v][0]));
                    acc[v] = __fmul2_rn(df, df);
#pragma unroll
                    for (int a = 1; a < D; ++a) {
                        df = __fadd2_rn(make_float2(x[t][a], x[t + 1][a]),
                                        make_float2(-p[u + v][a], -p[u + v][a]));
                        acc[v] = __ffma2_rn(df, df, acc[v]);
                    }
                }
                m[t] = fmin3(m[t], acc[0].x, acc[1].x);
                m[t + 1] = fmin3(m[t + 1], acc[0].y, acc[1].y);
            }
            if (NT_ & 1) {
                constexpr int t = NT_ - 1;
                m[t] = fmin3(m[t], dist2<D>(x[t], p[u]), dist2<D>(x[t], p[u + 1]));
            }
        }
    }
}

// Pruned sweep (exact).  The warp keeps the axis-aligned box of its sample points and the largest
// of its running minima u.  A candidate whose distance to that box is at least sqrt(u) cannot lower
// any of the warp's minima, so it is skipped: lanes test 32 tile records at a time against the
// box, the survivors (ballot mask) go through the same 4-candidate packed body as sweep_tile.
// u only shrinks, so a skip stays justified; the minima that come out are bit-identical to the
// exhaustive sweep (tests/test_gpu_kernels.py::test_pruning_is_exact).  This is the "tighter
// candidate rule" of SURVEY.md section 8(f2): the unit of work E is still counted by the
// reference's ball rule, fewer evaluations are executed.
template <int D, int NT_, int MAXT>
__device__ __forceinline__ void sweep_tile_pruned(const typename Rec<D>::type *__restrict__ tile, int n,
                                                  int sentinel_idx, const float (&x)[MAXT][D],
                                                  float (&m)[MAXT], const float (&blo)[D],
                                                  const float (&bhi)[D], int lane,
                                                  unsigned long long &executed) {
    auto bound = [&]() {
        float u = m[0];
#pragma unroll
        for (int t = 1; t < NT_; ++t) u = fmaxf(u, m[t]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) u = fmaxf(u, __shfl_xor_sync(0xffffffffu, u, o));
        return u;
    };
    float u = bound();
    int pend0 = 0, pend1 = 0, pend2 = 0, npend = 0;
#pragma unroll 1
    for (int base = 0;; base += 32) {
        const bool last = base >= n;   // one extra trip flushes the carried survivors
        unsigned mask = 0u;
        if (!last) {
            const int idx = base + lane;
            float own[D];
            rec_unpack<D>(tile[idx < n ? idx : sentinel_idx], own);
            float box2 = 0.f;
#pragma unroll
            for (int a = 0; a < D; ++a) {
                const float e = fmaxf(fmaxf(blo[a] - own[a], own[a] - bhi[a]), 0.f);
                box2 = fmaf(e, e, box2);
            }
            // 0.9999: the box distance and the pair distances are rounded differently
            mask = __ballot_sync(0xffffffffu, idx < n && box2 * 0.9999f <= u);
            if (mask == 0u) continue;
            executed += (unsigned)__popc(mask);
        }
        // survivors are swept four at a time; fewer than four are carried over to the next block
        // (pend0..2, warp-uniform) so that the packed body runs on full groups
        bool swept = false;
        while (npend + __popc(mask) >= kUnroll || (last && npend > 0)) {
            float p[kUnroll][D];
#pragma unroll
            for (int v = 0; v < kUnroll; ++v) {
                int j = sentinel_idx;
                if (v < npend) {
                    j = v == 0 ? pend0 : (v == 1 ? pend1 : pend2);
                } else if (mask) {
                    j = base + __ffs(mask) - 1;
                    mask &= mask - 1;
                }
                rec_unpack<D>(tile[j], p[v]);
            }
            npend = 0;
            swept = true;
#pragma unroll
            for (int v0 = 0; v0 < kUnroll; v0 += 2) {
#pragma unroll
                for (int t = 0; t + 1 < NT_; t += 2) {
                    float2 acc[2];
#pragma unroll
                    for (int v = 0; v < 2; ++v) {
                        float2 df = __fadd2_rn(make_float2(x[t][0], x[t + 1][0]),
                                               make_float2(-p[v0 + v][0], -p[v0 + v][0]));
                        acc[v] = __fmul2_rn(df, df);
#pragma unroll
                        for (int a = 1; a < D; ++a) {
                            df = __fadd2_rn(make_float2(x[t][a], x[t + 1][a]),
                                            make_float2(-p[v0 + v][a], -p[v0 + v][a]));
                            acc[v] = __ffma2_rn(df, df, acc[v]);
                        }
                    }
                    m[t] = fmin3(m[t], acc[0].x, acc[1].x);
                    m[t + 1] = fmin3(m[t + 1], acc[0].y, acc[1].y);
                }
                if (NT_ & 1) {
                    constexpr int t = NT_ - 1;
                    m[t] = fmin3(m[t], dist2<D>(x[t], p[v0]), dist2<D>(x[t], p[v0 + 1]));
                }
            }
        }
        while (mask) {
            const int jn = base + __ffs(mask) - 1;
            mask &= mask - 1;
            if (npend == 0) pend0 = jn; else if (npend == 1) pend1 = jn; else pend2 = jn;
            ++npend;
        }
        if (swept) u = bound();
        if (last) break;
    }
}

// MAXT = sample groups per warp held in registers, MAXW = warps per CTA, MINB = CTAs per SM the
// register budget is sized for: (8, 20, 1) is the wide shape; (4, 16, 2) trades registers for
// twice the warps per SM, which the latency-bound pruned sweep needs.
template <int D, bool PRUNE, int MAXT, int MAXW, int MINB>
__global__ void __launch_bounds__(MAXW * 32, MINB) cover_eval_kernel(const CoverParams P) {
    constexpr int kMaxT = MAXT;
    using RecT = typename Rec<D>::type;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NT = blockDim.x;
    const int tile_cap = P.tile_cap;
    RecT *tile = reinterpret_cast<RecT *>(smem_raw);
    int *run_start = reinterpret_cast<int *>(smem_raw + (size_t)(tile_cap + kUnroll) * sizeof(RecT));
    int *run_pos = run_start + NT;
    __shared__ int warp_sums[32];
    __shared__ int s_fill;
    __shared__ long long s_item[3];  // simplex, chunk, sample block (-1 = queue drained)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = NT >> 5;
    const GridParams gp = *P.gp;
    const RecT *__restrict__ points = reinterpret_cast<const RecT *>(P.points);
    const long long total_chunks = P.item_base[P.S];
    const unsigned long long total_items = (unsigned long long)total_chunks * (unsigned)P.nsb;

    if (tid == 0) s_fill = 0;
    if (tid < kUnroll) tile[tile_cap + tid] = rec_sentinel<D>();   // never overwritten (pruned sweeps pad with them)
    unsigned long long executed = 0;         // candidates this warp swept in the current item (pruned mode)
    unsigned long long executed_evals = 0;   // evaluations this warp performed in the whole launch

    for (;;) {
        // ---- fetch a work item ---------------------------------------------------------------
        __syncthreads();  // previous item fully retired (s_item, tile, run arrays reusable)
        if (tid == 0) {
            const unsigned long long g = atomicAdd(P.queue, 1ull);
            if (g >= total_items) {
                s_item[0] = -1;
            } else {
                const long long gi = (long long)(g / (unsigned)P.nsb);
                long long lo = 0, hi = P.S;  // last s with item_base[s] <= gi
                while (hi - lo > 1) {
                    const long long mid = (lo + hi) >> 1;
                    if (P.item_base[mid] <= gi) lo = mid; else hi = mid;
                }
                s_item[0] = lo;
                s_item[1] = gi - P.item_base[lo];
                s_item[2] = (long long)(g % (unsigned)P.nsb);
            }
        }
        __syncthreads();
        const long long s = s_item[0];
        if (s < 0) break;
        const long long chunk_j = s_item[1];
        const int sb = (int)s_item[2];

        // ---- simplex constants ---------------------------------------------------------------
        float c[D];
#pragma unroll
        for (int a = 0; a < D; ++a) c[a] = __ldg(P.centers + s * D + a);
        const float rad = __ldg(P.radii + s);
        const float r2 = rad * rad;
        const BallCells bc = ball_cells(c, rad, D, gp);
        const long long tested = P.tested[s];
        const long long nch = P.item_base[s + 1] - P.item_base[s];
        const int win_lo = (int)(chunk_j * tested / nch);
        const int win_hi = (int)((chunk_j + 1) * tested / nch);

        // ---- this warp's sample groups ---------------------------------------------------------
        // The sample block's groups (32 consecutive samples each) are dealt to the warps as evenly
        // as possible; consecutive warps sit on different SM sub-partitions, so the sub-partition
        // loads differ by at most one group.
        const int blk_g0 = sb * P.groups_per_block;
        const int blk_groups = min(P.groups_per_block, P.groups - blk_g0);
        const int g_base = blk_groups / W, g_rem = blk_groups % W;
        const int nt = g_base + (warp < g_rem ? 1 : 0);
        const int g0 = blk_g0 + warp * g_base + min(warp, g_rem);
        float x[kMaxT][D], m[kMaxT];
#pragma unroll
        for (int t = 0; t < kMaxT; ++t) {
            const long long r = (long long)(g0 + t) * 32 + lane;
            // pruned mode starts from what other chunks / the seed pass already found (an upper
            // bound of the minimum); unused slots carry 0 so that they never loosen the warp bound
            m[t] = PRUNE ? 0.f : INFINITY;
            if (t < nt && r < P.R) {
                if (PRUNE) m[t] = __ldcg(P.out + s * P.R + r);
                if (P.samples) {
#pragma unroll
                    for (int a = 0; a < D; ++a) x[t][a] = __ldg(P.samples + (s * P.R + r) * D + a);
                } else {
                    // x = sum_k w[r,k] * v[s,k,:], FMA chain over k ascending (== the reference's
                    // float32 matmul, core.py:188)
                    const float *w = P.weights + r * P.K;
                    const float *v = P.verts + s * P.K * D;
                    const float w0 = __ldg(w);
#pragma unroll
                    for (int a = 0; a < D; ++a) x[t][a] = __fmul_rn(w0, __ldg(v + a));
                    for (int k = 1; k < P.K; ++k) {
                        const float wk = __ldg(w + k);
#pragma unroll
                        for (int a = 0; a < D; ++a) x[t][a] = fmaf(wk, __ldg(v + k * D + a), x[t][a]);
                    }
                }
            } else {
#pragma unroll
                for (int a = 0; a < D; ++a) x[t][a] = c[a];
            }
        }

        // box of this warp's sample points
        float blo[D], bhi[D];
        if (PRUNE) {
#pragma unroll
            for (int a = 0; a < D; ++a) {
                float lo = INFINITY, hi = -INFINITY;
#pragma unroll
                for (int t = 0; t < kMaxT; ++t) {
                    const long long r = (long long)(g0 + t) * 32 + lane;
                    if (t < nt && r < P.R) { lo = fminf(lo, x[t][a]); hi = fmaxf(hi, x[t][a]); }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
                    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
                }
                blo[a] = lo;
                bhi[a] = hi;
            }
        }

        auto sweep = [&](int n) __attribute__((always_inline)) {
            if (PRUNE) {
                __syncthreads();
                switch (nt) {  // warp-uniform
                    case 1: if constexpr (1 <= MAXT) sweep_tile_pruned<D, 1, MAXT>(tile, n, tile_cap, x, m, blo, bhi, lane, executed); break;
                    case 2: if constexpr (2 <= MAXT) sweep_tile_pruned<D, 2, MAXT>(tile, n, tile_cap, x, m, blo, bhi, lane, executed); break;
                    case 3: if constexpr (3 <= MAXT) sweep_tile_pruned<D, 3, MAXT>(tile, n, tile_cap, x, m, blo, bhi, lane, executed); break;
                    case 4: if constexpr (4 <= MAXT) sweep_tile_pruned<D, 4, MAXT>(tile, n, tile_cap, x, m, blo, bhi, lane, executed); break;
                    case 5: if constexpr (5 <= MAXT) sweep_tile_pruned<D, 5, MAXT>(tile, n, tile_cap, x, m, blo, bhi, lane, executed); break;
                    case 6: if constexpr (6 <= MAXT) sweep_tile_pruned<D, 6, MAXT>(tile, n, tile_cap, x, m, blo, bhi, lane, executed); break;
                    case 7: if constexpr (7 <= MAXT) sweep_tile_pruned<D, 7, MAXT>(tile, n, tile_cap, x, m, blo, bhi, lane, executed); break;
                    case 8: if constexpr (8 <= MAXT) sweep_tile_pruned<D, 8, MAXT>(tile, n, tile_cap, x, m, blo, bhi, lane, executed); break;
                    default: break;
                }
                __syncthreads();
                if (tid == 0) s_fill = 0;
                return;
            }
            const int npad = (n + kUnroll - 1) / kUnroll * kUnroll;
            if (tid < npad - n) tile[n + tid] = rec_sentinel<D>();
            __syncthreads();
            switch (nt) {  // warp-uniform
                case 1: if constexpr (1 <= MAXT) sweep_tile<D, 1, MAXT>(tile, npad, x, m); break;
                case 2: if constexpr (2 <= MAXT) sweep_tile<D, 2, MAXT>(tile, npad, x, m); break;
                case 3: if constexpr (3 <= MAXT) sweep_tile<D, 3, MAXT>(tile, npad, x, m); break;
                case 4: if constexpr (4 <= MAXT) sweep_tile<D, 4, MAXT>(tile, npad, x, m); break;
                case 5: if constexpr (5 <= MAXT) sweep_tile<D, 5, MAXT>(tile, npad, x, m); break;
                case 6: if constexpr (6 <= MAXT) sweep_tile<D, 6, MAXT>(tile, npad, x, m); break;
                case 7: if constexpr (7 <= MAXT) sweep_tile<D, 7, MAXT>(tile, npad, x, m); break;
                case 8: if constexpr (8 <= MAXT) sweep_tile<D, 8, MAXT>(tile, npad, x, m); break;
                default: break;
            }
            __syncthreads();
            if (tid == 0) s_fill = 0;
        };

        // ---- stream the candidate window -----------------------------------------------------
        int fill = 0;
        long long accepted = 0;
        int carry = 0;  // stream offset of the current row batch
        for (int rb = 0; rb < bc.nrows; rb += NT) {
            if (carry >= win_hi) break;
            const int row = rb + tid;
            int a = 0, len = 0;
            if (row < bc.nrows) row_run(bc, row, gp, P.cell_start, a, len);
            int batch_total;
            const int off = carry + block_exclusive_scan(len, warp_sums, batch_total);
            carry += batch_total;
            if (carry <= win_lo) continue;
            // clip the run to this chunk's window of the stream
            const int s0 = max(off, win_lo), s1 = min(off + len, win_hi);
            const int len2 = max(0, s1 - s0);
            int total2;
            const int pos2 = block_exclusive_scan(len2, warp_sums, total2);
            run_start[tid] = a + (s0 - off);
            run_pos[tid] = pos2;
            if (tid == 0) run_pos[NT] = total2;
            __syncthreads();

            const int stride = P.stream_stride;
            for (int base = 0; base * stride < total2; base += NT) {
                if (fill + NT > tile_cap) {
                    accepted += fill;
                    sweep(fill);
                    __syncthreads();
                    fill = 0;
                }
                const int q = (base + tid) * stride;
                bool pass = false;
                RecT rec;
                if (q < total2) {
                    // last run with run_pos <= q
                    int lo = 0, hi = NT;
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        if (run_pos[mid] <= q) lo = mid; else hi = mid;
                    }
                    rec = points[run_start[lo] + (q - run_pos[lo])];
                    float p[D];
                    rec_unpack<D>(rec, p);
                    // the reference predicate (triton_kernels.py:137-148): sum (p-c)^2 <= r^2
                    float t = p[0] - c[0];
                    float acc = t * t;
#pragma unroll
                    for (int a2 = 1; a2 < D; ++a2) {
                        t = p[a2] - c[a2];
                        acc = fmaf(t, t, acc);
                    }
                    pass = acc <= r2;
                }
                const unsigned ballot = __ballot_sync(0xffffffffu, pass);
                int wbase = 0;
                if (lane == 0 && ballot) wbase = atomicAdd(&s_fill, __popc(ballot));
                wbase = __shfl_sync(0xffffffffu, wbase, 0);
                if (pass) tile[wbase + __popc(ballot & ((1u << lane) - 1u))] = rec;
                __syncthreads();
                fill = s_fill;
            }
        }
        if (fill > 0) {
            accepted += fill;
            sweep(fill);
        }

        // ---- merge -----------------------------------------------------------------------------
#pragma unroll
        for (int t = 0; t < kMaxT; ++t) {
            const long long r = (long long)(g0 + t) * 32 + lane;
            if (t < nt && r < P.R && m[t] < INFINITY)
                atomicMin(reinterpret_cast<unsigned *>(P.out + s * P.R + r), __float_as_uint(m[t]));
        }
        if (tid == 0 && sb == 0 && accepted > 0 && P.count_work) {
            if (P.cand_count) atomicAdd(reinterpret_cast<unsigned long long *>(P.cand_count + s),
                                        (unsigned long long)accepted);
            if (P.evals) atomicAdd(P.evals, (unsigned long long)accepted * (unsigned long long)P.R);
        }
        if (PRUNE) {
            executed_evals += executed * (unsigned long long)(nt * 32);
            executed = 0;
        } else {
            executed_evals += (unsigned long long)accepted * (unsigned long long)(nt * 32);
        }
    }
    if (lane == 0 && executed_evals) atomicAdd(P.executed, executed_evals);
}

__global__ void fill_inf_kernel(float *p, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        p[i] = INFINITY;
}

struct CoverLayout {
    int64_t off_tested, off_item_base, off_queue, total;
};

CoverLayout cover_layout(int64_t S) {
    CoverLayout L;
    int64_t o = 0;
    L.off_queue = o;      o = align_up(o + 64, 256);
    L.off_tested = o;     o = align_up(o + S * 4, 256);
    L.off_item_base = o;  o = align_up(o + (S + 1) * 8, 256);
    L.total = o;
    return L;
}

// launch with a given kernel shape: MAXT groups per warp, at most MAXW warps per CTA
template <int D, bool PRUNE, int MAXT, int MAXW, int MINB>
int launch_eval_shape(CoverParams &P, int64_t R, cudaStream_t st) {
    using RecT = typename Rec<D>::type;
    auto kern = cover_eval_kernel<D, PRUNE, MAXT, MAXW, MINB>;
    // Shape of a CTA pass: G sample groups over W warps (a multiple of 4, one set per SM
    // sub-partition), at most MAXT groups per warp; more than MAXW * MAXT groups are split into
    // equal sample blocks.
    const int G = (int)((R + 31) / 32);
    auto warps_for = [](int groups) {
        if (groups < 4) return groups < 1 ? 1 : groups;
        int w = (groups + MAXT - 1) / MAXT;
        w = (w + 3) / 4 * 4;
        return w > MAXW ? MAXW : w;
    };
    int forced = get_option("warps", 0);
    if (forced < 0 || forced > MAXW) forced = 0;
    int W = forced ? forced : warps_for(G);
    P.nsb = (G + W * MAXT - 1) / (W * MAXT);
    P.groups = G;
    P.groups_per_block = (G + P.nsb - 1) / P.nsb;
    if (!forced) W = warps_for(P.groups_per_block);
    const int NT = W * 32;
    int cap = NT >= 512 ? 2048 : (4 * NT < 512 ? 512 : 4 * NT);
    const int forced_cap = get_option("tile_cap", 0);
    if (forced_cap >= NT + kUnroll) cap = forced_cap / kUnroll * kUnroll;
    P.tile_cap = cap;
    const size_t smem = (size_t)(cap + kUnroll) * sizeof(RecT) + (size_t)(2 * NT + 1) * sizeof(int);
    FLOOD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    FLOOD_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem));
    if (per_sm < 1) return set_error(FLOOD_E_CUDA, "cover_eval_kernel does not fit on an SM");
    const int limit = get_option("ctas_per_sm", 0);
    if (limit > 0 && per_sm > limit) per_sm = limit;
    const int grid = device_sm_count() * per_sm;
    const bool timed = get_option("time_kernels", 0) != 0;
    if (timed) kernel_timer_start("cover_eval", st);
    unsigned long long *queue0 = P.queue;
    const int seed_stride = get_option("seed_stride", 16);
    if (PRUNE && seed_stride > 1) {
        // seed pass: every seed_stride-th stream position gives every sample an upper bound of its
        // minimum, so the full pass prunes from its first tile on
        P.stream_stride = seed_stride;
        P.count_work = 0;
        kern<<<grid, NT, smem, st>>>(P);
        P.queue = queue0 + 1;
    }
    P.stream_stride = 1;
    P.count_work = 1;
    kern<<<grid, NT, smem, st>>>(P);
    P.queue = queue0;
    if (timed) kernel_timer_stop("cover_eval", st);
    FLOOD_LAUNCH_CHECK("cover_eval_kernel");
    return FLOOD_OK;
}

template <int D>
int dispatch_eval(CoverParams &P, int64_t R, cudaStream_t st) {
    const bool prune = get_option("prune", 1) != 0;
    // One shape for both modes: 8 sample groups per warp, up to 20 warps, one CTA per SM for wide
    // sample sets.  (A thin shape -- 4 groups per warp, 16 warps, 2 CTAs per SM -- prunes more,
    // 26 % instead of 37 % of E executed on the torus, but its per-record overhead makes it slower.)
    if (!prune) return launch_eval_shape<D, false, 8, 20, 1>(P, R, st);
    return launch_eval_shape<D, true, 8, 20, 1>(P, R, st);
}

}  // namespace

// cost estimate per simplex: length of its candidate stream (points in the cell rows its ball
// touches); used by the host to balance simplices over GPUs
int covering_plan(const void *cloud_ws, int64_t n, int d, const float *centers, const float *radii,
                  int64_t S, int32_t *out_tested, cudaStream_t st) {
    if (S == 0) return FLOOD_OK;
    if (!cloud_ws || !centers || !radii || !out_tested || S < 0 || n < 1 || d < 2 || d > FLOOD_MAX_DIM)
        return set_error(FLOOD_E_INVALID, "covering_plan: bad arguments (S=%lld n=%lld d=%d)", (long long)S,
                         (long long)n, d);
    const CloudLayout C = cloud_layout(n, d);
    const char *cbase = static_cast<const char *>(cloud_ws);
    CoverParams P = {};
    P.gp = reinterpret_cast<const GridParams *>(cbase + C.off_grid);
    P.cell_start = reinterpret_cast<const int *>(cbase + C.off_cell_start);
    P.centers = centers;
    P.radii = radii;
    P.tested = out_tested;
    P.item_base = nullptr;
    P.S = S;
    P.chunk = 1;
    P.rows_per_chunk_factor = 0;
    const int threads = 128;
    const long long blocks = (S * 32 + threads - 1) / threads;
    cover_plan_kernel<<<(unsigned)blocks, threads, 0, st>>>(P, d);
    FLOOD_LAUNCH_CHECK("cover_plan_kernel");
    return FLOOD_OK;
}

size_t covering_workspace_bytes(int64_t S, int64_t R, int d) {
    (void)R; (void)d;
    return (size_t)cover_layout(S < 1 ? 1 : S).total;
}

int covering_radius(const void *cloud_ws, int64_t n, int d, const float *verts, int64_t S, int K,
                    const float *weights, int64_t R, const float *samples, const float *centers,
                    const float *radii, float *out_min_dist2, int64_t *out_cand_count,
                    unsigned long long *out_evals, void *ws, size_t ws_bytes, cudaStream_t st) {
    if (S == 0) return FLOOD_OK;
    if (!cloud_ws || !centers || !radii || !out_min_dist2 || !ws || S < 0 || R < 1 || n < 1 ||
        d < 2 || d > FLOOD_MAX_DIM || K < 1 || K > FLOOD_MAX_SIMPLEX_VERTS)
        return set_error(FLOOD_E_INVALID, "covering_radius: bad arguments (S=%lld R=%lld n=%lld d=%d K=%d)",
                         (long long)S, (long long)R, (long long)n, d, K);
    if (!samples && (!verts || !weights))
        return set_error(FLOOD_E_INVALID, "covering_radius: need either samples or (verts, weights)");
    const CoverLayout L = cover_layout(S);
    if ((int64_t)ws_bytes < L.total)
        return set_error(FLOOD_E_WORKSPACE, "covering_radius: workspace %zu < %lld bytes", ws_bytes,
                         (long long)L.total);
    const CloudLayout C = cloud_layout(n, d);
    const char *cbase = static_cast<const char *>(cloud_ws);
    char *wbase = static_cast<char *>(ws);

    CoverParams P;
    P.gp = reinterpret_cast<const GridParams *>(cbase + C.off_grid);
    P.cell_start = reinterpret_cast<const int *>(cbase + C.off_cell_start);
    P.points = cbase + C.off_points;
    P.verts = verts;
    P.weights = weights;
    P.samples = samples;
    P.centers = centers;
    P.radii = radii;
    P.out = out_min_dist2;
    P.cand_count = reinterpret_cast<long long *>(out_cand_count);
    P.evals = out_evals;
    P.tested = reinterpret_cast<int *>(wbase + L.off_tested);
    P.item_base = reinterpret_cast<long long *>(wbase + L.off_item_base);
    P.queue = reinterpret_cast<unsigned long long *>(wbase + L.off_queue);
    P.executed = P.queue + 2;
    P.S = S;
    P.R = R;
    P.K = K;
    P.nsb = 1;
    P.chunk = get_option("chunk", 8192);
    if (P.chunk < 256) P.chunk = 256;
    P.rows_per_chunk_factor = get_option("rows_per_chunk_factor", 32);

    FLOOD_CUDA_CHECK(cudaMemsetAsync(P.queue, 0, 64, st));
    if (out_cand_count) FLOOD_CUDA_CHECK(cudaMemsetAsync(out_cand_count, 0, (size_t)S * 8, st));
    {
        const long long total = S * R;
        int blocks = (int)((total + 1023) / 1024);
        const int cap = device_sm_count() * 8;
        if (blocks > cap) blocks = cap;
        fill_inf_kernel<<<blocks, 256, 0, st>>>(out_min_dist2, total);
    }
    {
        const int threads = 128;  // 4 simplices per CTA
        const long long blocks = (S * 32 + threads - 1) / threads;
        cover_plan_kernel<<<(unsigned)blocks, threads, 0, st>>>(P, d);
        cover_scan_kernel<<<1, 1024, 0, st>>>(P.item_base, S);
    }
    FLOOD_LAUNCH_CHECK("cover plan kernels");
    switch (d) {
        case 2: return dispatch_eval<2>(P, R, st);
        case 3: return dispatch_eval<3>(P, R, st);
        case 4: return dispatch_eval<4>(P, R, st);
        case 5: return dispatch_eval<5>(P, R, st);
        case 6: return dispatch_eval<6>(P, R, st);
        case 7: return dispatch_eval<7>(P, R, st);
        case 8: return dispatch_eval<8>(P, R, st);
    }
    return set_error(FLOOD_E_UNSUPPORTED, "covering_radius: d=%d", d);
}

}  // namespace flood

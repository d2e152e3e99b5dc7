// Covering-radius evaluation kernel (templates).  Included by covering.cu (host entry points) and
// by covering_d<N>.cu, which instantiate dispatch_eval<N> one ambient dimension per translation
// unit so that the dimensions compile in parallel.  See covering.cu for the description.
#pragma once
#include "common.cuh"

namespace flood {


constexpr int kUnroll = 4;       // candidates per inner-loop trip
constexpr int kAsyncLPL = 2;       // cp.async gather: records per lane and staging buffer
constexpr int kBoundRefresh = 16;  // pruned sweep: candidates swept between refreshes of the warp bound

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
    long long *item_base_seed;  // [S+1] the same for the seed pass (longer chunks), or null
    int chunk_seed;          // target tested points per chunk of the seed pass
    int async_gather;        // 1: records are staged through shared memory with cp.async (double-buffered)
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

// Shape of a CTA pass: G sample groups over W warps (a multiple of 4, one set per SM
// sub-partition), at most maxt groups per warp; more than maxw * maxt groups are split into nsb
// sample blocks of (almost) equal size.
struct EvalShape {
    int W, nsb, groups, groups_per_block;
};

inline EvalShape eval_shape(int64_t R, int maxt, int maxw) {
    const int G = (int)((R + 31) / 32);
    auto warps_for = [&](int groups) {
        if (groups < 4) return groups < 1 ? 1 : groups;
        int w = (groups + maxt - 1) / maxt;
        w = (w + 3) / 4 * 4;
        return w > maxw ? maxw : w;
    };
    int forced = get_option("warps", 0);
    if (forced < 0 || forced > maxw) forced = 0;
    EvalShape sh;
    sh.W = forced ? forced : warps_for(G);
    sh.nsb = (G + sh.W * maxt - 1) / (sh.W * maxt);
    sh.groups = G;
    sh.groups_per_block = (G + sh.nsb - 1) / sh.nsb;
    if (!forced) sh.W = warps_for(sh.groups_per_block);
    return sh;
}

// Kernel shapes (template parameters MAXT, MAXW, MINB), selected with option "shape":
//   0  (8, 20, 1)  wide: one CTA per SM holds all samples of a simplex (R <= 5120)
//   1  (8,  4, 5)  narrow: sample blocks of 4 warps, five independent CTAs per SM
//   2  (8,  8, 2)
struct ShapeDesc {
    int maxt, maxw, minb;
};
constexpr ShapeDesc kShapes[3] = {{8, 20, 1}, {8, 4, 5}, {8, 8, 2}};

inline int pick_shape(bool prune) {
    int sh = get_option("shape", prune ? 1 : 0);
    return (sh < 0 || sh > 2) ? 0 : sh;
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
__device__ __forceinline__ float sweep_tile_pruned(const typename Rec<D>::type *__restrict__ tile, int n,
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
    int stale = 0;   // candidates swept since u was last refreshed (a stale u is larger, i.e. still valid)
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
            stale += kUnroll;
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
        if (stale >= kBoundRefresh) {
            u = bound();
            stale = 0;
        }
        if (last) break;
    }
    return stale ? bound() : u;
}

// last i in [0, n) with pos[i] <= p, for a non-decreasing pos[] with pos[0] <= p.  Warp-uniform
// two-level ballot search (two shared-memory rounds instead of a log2(n) dependent chain).
__device__ __forceinline__ int find_run(const int *pos, int n, int p, int lane) {
    const int step = (n + 31) >> 5;  // n <= 1024
    const int ia = lane * step;
    const int c = __popc(__ballot_sync(0xffffffffu, ia < n && pos[ia] <= p)) - 1;
    const int ib = c * step + lane;
    return c * step + __popc(__ballot_sync(0xffffffffu, lane < step && ib < n && pos[ib] <= p)) - 1;
}

// MAXT = sample groups per warp held in registers, MAXW = warps per CTA, MINB = CTAs per SM the
// register budget is sized for.  (8, 20, 1) is the wide shape: one CTA per SM holds every sample
// of a simplex.  (8, 4, 5) splits the samples of a simplex over sample blocks of 4 warps: five
// independent CTAs per SM (one warp per sub-partition each), so a warp that waits at a CTA
// barrier for a slower one leaves the issue slots to the other CTAs.
template <int D, bool PRUNE, int MAXT, int MAXW, int MINB>
__global__ void __launch_bounds__(MAXW * 32, MINB) cover_eval_kernel(const CoverParams P) {
    constexpr int kMaxT = MAXT;
    constexpr int LPL = D <= 4 ? 4 : 2;  // records in flight per lane while gathering
    constexpr int UNIT = 32 * LPL;       // stream positions per warp work unit
    constexpr int ALPL = kAsyncLPL;      // the same for the cp.async path (staged in shared memory)
    constexpr int AUNIT = 32 * ALPL;
    using RecT = typename Rec<D>::type;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NT = blockDim.x;
    const int tile_cap = P.tile_cap;
    RecT *tile = reinterpret_cast<RecT *>(smem_raw);
    // per-warp staging ring of the cp.async gather: 2 buffers x AUNIT records
    RecT *stage = tile + (tile_cap + kUnroll) + (size_t)(threadIdx.x >> 5) * (P.async_gather ? 2 * AUNIT : 0);
    int *run_start = reinterpret_cast<int *>(smem_raw + (size_t)(tile_cap + kUnroll) * sizeof(RecT) +
                                             (P.async_gather ? (size_t)(NT >> 5) * 2 * AUNIT * sizeof(RecT) : 0));
    int *run_pos = run_start + NT;
    __shared__ int warp_sums[32];
    __shared__ int s_fill;
    __shared__ long long s_item[3];          // simplex, chunk, sample block (-1 = queue drained)
    __shared__ float s_wbox[MAXW][2 * D];    // per-warp sample boxes (pruned mode)
    __shared__ float s_wu[MAXW];             // per-warp largest running minimum
    __shared__ float s_cbox[2 * D + 1];      // box of all samples of the CTA, and the CTA-wide bound

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = NT >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const GridParams gp = *P.gp;
    const RecT *__restrict__ points = reinterpret_cast<const RecT *>(P.points);
    const long long total_chunks = P.item_base[P.S];
    const unsigned long long total_items = (unsigned long long)total_chunks * (unsigned)P.nsb;
    const int stride = P.stream_stride;

    if (tid == 0) s_fill = 0;
    if (tid < kUnroll) tile[tile_cap + tid] = rec_sentinel<D>();   // never overwritten (pruned sweeps pad with them)
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
        // loads differ by at most one group.  (flood_covering_bricks reports this layout, so the
        // host can order the samples such that every warp holds a compact brick.)
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
            if (PRUNE && t < nt && r < P.R) m[t] = __ldcg(P.out + s * P.R + r);
#pragma unroll
            for (int a = 0; a < D; ++a) x[t][a] = c[a];
        }
        if (P.samples) {
#pragma unroll
            for (int t = 0; t < kMaxT; ++t) {
                const long long r = (long long)(g0 + t) * 32 + lane;
                if (t < nt && r < P.R) {
#pragma unroll
                    for (int a = 0; a < D; ++a) x[t][a] = __ldg(P.samples + (s * P.R + r) * D + a);
                }
            }
        } else {
            // x = sum_k w[r,k] * v[s,k,:], FMA chain over k ascending (== the reference's float32
            // matmul, core.py:188); k is the outer loop so that a vertex is loaded once per warp
#pragma unroll 1
            for (int k = 0; k < P.K; ++k) {
                float v[D];
#pragma unroll
                for (int a = 0; a < D; ++a) v[a] = __ldg(P.verts + (s * P.K + k) * D + a);
#pragma unroll
                for (int t = 0; t < kMaxT; ++t) {
                    const long long r = (long long)(g0 + t) * 32 + lane;
                    if (t < nt && r < P.R) {
                        const float wk = __ldg(P.weights + r * P.K + k);
#pragma unroll
                        for (int a = 0; a < D; ++a)
                            x[t][a] = k == 0 ? __fmul_rn(wk, v[a]) : fmaf(wk, v[a], x[t][a]);
                    }
                }
            }
        }

        // boxes and bounds (pruned mode): per warp in registers, per CTA in shared memory
        float blo[D], bhi[D];
        float U = INFINITY;   // CTA-wide: the largest running minimum of any sample of the CTA
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
                if (lane == 0) { s_wbox[warp][a] = lo; s_wbox[warp][D + a] = hi; }
            }
            float u = m[0];
#pragma unroll
            for (int t = 1; t < kMaxT; ++t) u = fmaxf(u, m[t]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) u = fmaxf(u, __shfl_xor_sync(0xffffffffu, u, o));
            if (lane == 0) s_wu[warp] = u;
            __syncthreads();
            if (tid <= 2 * D) {
                float v;
                if (tid < D) { v = INFINITY; for (int w = 0; w < W; ++w) v = fminf(v, s_wbox[w][tid]); }
                else if (tid < 2 * D) { v = -INFINITY; for (int w = 0; w < W; ++w) v = fmaxf(v, s_wbox[w][tid]); }
                else { v = 0.f; for (int w = 0; w < W; ++w) v = fmaxf(v, s_wu[w]); }
                s_cbox[tid] = v;
            }
            __syncthreads();
            U = s_cbox[2 * D];
        }

        unsigned long long executed = 0;   // candidate records this warp swept in the current item
        auto sweep = [&](int n) __attribute__((always_inline)) {
            // on entry: the tile holds n records and every thread is past the barrier that
            // completed it; on exit: the tile is empty and reusable
            if (PRUNE) {
                float u = 0.f;
                switch (nt) {  // warp-uniform
                    case 1: if constexpr (1 <= MAXT) u = sweep_tile_pruned<D, 1, MAXT>(tile, n, tile_cap, x, m, blo, bhi, lane, executed); break;
                    case 2: if constexpr (2 <= MAXT) u = sweep_tile_pruned<D, 2, MAXT>(tile, n, tile_cap, x, m, blo, bhi, lane, executed); break;
                    case 3: if constexpr (3 <= MAXT) u = sweep_tile_pruned<D, 3, MAXT>(tile, n, tile_cap, x, m, blo, bhi, lane, executed); break;
                    case 4: if constexpr (4 <= MAXT) u = sweep_tile_pruned<D, 4, MAXT>(tile, n, tile_cap, x, m, blo, bhi, lane, executed); break;
                    case 5: if constexpr (5 <= MAXT) u = sweep_tile_pruned<D, 5, MAXT>(tile, n, tile_cap, x, m, blo, bhi, lane, executed); break;
                    case 6: if constexpr (6 <= MAXT) u = sweep_tile_pruned<D, 6, MAXT>(tile, n, tile_cap, x, m, blo, bhi, lane, executed); break;
                    case 7: if constexpr (7 <= MAXT) u = sweep_tile_pruned<D, 7, MAXT>(tile, n, tile_cap, x, m, blo, bhi, lane, executed); break;
                    case 8: if constexpr (8 <= MAXT) u = sweep_tile_pruned<D, 8, MAXT>(tile, n, tile_cap, x, m, blo, bhi, lane, executed); break;
                    default: break;
                }
                if (lane == 0) s_wu[warp] = u;
                __syncthreads();
                if (tid == 0) s_fill = 0;
                float cu = 0.f;
                for (int w = 0; w < W; ++w) cu = fmaxf(cu, s_wu[w]);
                U = cu;
                __syncthreads();
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
            executed += (unsigned)n;
            __syncthreads();
            if (tid == 0) s_fill = 0;
            __syncthreads();
        };

        // ball test (+ CTA-level cull) of up to N records per lane, warp-ballot compaction into the tile
        unsigned inball = 0;   // records inside the ball seen by this warp (lane-uniform)
        auto test_and_compact = [&](auto &rec, int nrec) __attribute__((always_inline)) {
            constexpr int N = sizeof(rec) / sizeof(rec[0]);
            unsigned keep[N];
            int nkeep = 0;
#pragma unroll
            for (int v = 0; v < N; ++v) {
                bool pass = false, near = false;
                if (lane + 32 * v < nrec) {
                    float q[D];
                    rec_unpack<D>(rec[v], q);
                    // the reference predicate (triton_kernels.py:137-148): sum (p-c)^2 <= r^2
                    float t = q[0] - c[0];
                    float acc = t * t;
#pragma unroll
                    for (int a2 = 1; a2 < D; ++a2) {
                        t = q[a2] - c[a2];
                        acc = fmaf(t, t, acc);
                    }
                    pass = acc <= r2;
                    if (PRUNE && pass) {
                        // CTA-level cull: a record at least sqrt(U) away from the box of the CTA's
                        // samples cannot lower any of their minima
                        float box2 = 0.f;
#pragma unroll
                        for (int a2 = 0; a2 < D; ++a2) {
                            const float e2 = fmaxf(fmaxf(s_cbox[a2] - q[a2], q[a2] - s_cbox[D + a2]), 0.f);
                            box2 = fmaf(e2, e2, box2);
                        }
                        near = box2 * 0.9999f <= U;
                    }
                }
                const unsigned bm = __ballot_sync(0xffffffffu, pass);
                inball += (unsigned)__popc(bm);
                keep[v] = PRUNE ? __ballot_sync(0xffffffffu, near) : bm;
                nkeep += __popc(keep[v]);
            }
            int wbase = 0;
            if (lane == 0 && nkeep) wbase = atomicAdd(&s_fill, nkeep);
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
#pragma unroll
            for (int v = 0; v < N; ++v) {
                if ((keep[v] >> lane) & 1u) tile[wbase + __popc(keep[v] & lt_mask)] = rec[v];
                wbase += __popc(keep[v]);
            }
        };

        // ---- stream the candidate window -----------------------------------------------------
        // Rows of the ball -> runs of the cell-sorted cloud, NT rows at a time; the runs of a
        // batch (clipped to this item's window of the stream) are gathered by the warps in units
        // of UNIT consecutive stream positions: coalesced record loads, ball test, warp-ballot
        // compaction into the tile.  The tile is swept whenever the next round might overflow it.
        int fill = 0;
        int carry = 0;         // stream offset of the next row batch
        int rb = 0;            // first row of the next row batch
        int q0 = 0, total2 = 0;  // progress inside the current row batch (stream positions)
        bool stream_done = false;
        // One sweep call site (the sweep is instantiated for every group count): alternate between
        // "gather until the tile cannot take the next round" and "sweep".
        for (;;) {
            while (!stream_done) {
                if (q0 >= total2) {
                    // next batch of NT rows -> runs
                    if (rb >= bc.nrows || carry >= win_hi) {
                        stream_done = true;
                        break;
                    }
                    const int row = rb + tid;
                    rb += NT;
                    int a = 0, len = 0;
                    if (row < bc.nrows) row_run(bc, row, gp, P.cell_start, a, len);
                    int batch_total;
                    const int off = carry + block_exclusive_scan(len, warp_sums, batch_total);
                    carry += batch_total;
                    q0 = total2 = 0;
                    if (carry <= win_lo) continue;
                    // clip the run to this chunk's window of the stream; a seed pass (stride > 1)
                    // takes every stride-th record of each clipped run
                    const int s0 = max(off, win_lo), s1 = min(off + len, win_hi);
                    const int len2 = max(0, s1 - s0);
                    const int pos2 = block_exclusive_scan((len2 + stride - 1) / stride, warp_sums, total2);
                    run_start[tid] = a + (s0 - off);
                    run_pos[tid] = pos2;
                    if (tid == 0) run_pos[NT] = total2;
                    __syncthreads();
                    continue;
                }
                if (fill > 0 && tile_cap - fill < min(total2 - q0, UNIT * W)) break;   // tile full: sweep first
            const int take = min(total2 - q0, tile_cap - fill);
            if (P.async_gather) {
                // cp.async (LDGSTS) double buffering: the records of the warp's next unit are in
                // flight to its staging buffer while the current unit is tested and compacted
                const int nunits = (take + AUNIT - 1) / AUNIT;
                auto issue = [&](int unit, int buf) {
                    int p = q0 + unit * AUNIT;
                    const int p1 = min(p + AUNIT, q0 + take);
                    const int n_u = p1 - p;
                    int i = find_run(run_pos, NT, p, lane);
                    RecT *dst = stage + buf * AUNIT;
                    while (p < p1) {
                        const int e = run_pos[i + 1];
                        if (e <= p) { ++i; continue; }
                        const int nrec = min(e, p1) - p;
                        const RecT *src = points + run_start[i] + (long long)(p - run_pos[i]) * stride;
                        for (int k = lane; k < nrec; k += 32) rec_cp_async<D>(dst + k, src + (long long)k * stride);
                        dst += nrec;
                        p += nrec;
                    }
                    cp_async_commit();
                    return n_u;
                };
                int unit = warp, buf = 0, n_cur = 0;
                if (unit < nunits) n_cur = issue(unit, 0);
                while (unit < nunits) {
                    const int next = unit + W;
                    int n_next = 0;
                    if (next < nunits) {
                        n_next = issue(next, buf ^ 1);
                        cp_async_wait<1>();
                    } else {
                        cp_async_wait<0>();
                    }
                    __syncwarp();
                    RecT rec[ALPL];
#pragma unroll
                    for (int v = 0; v < ALPL; ++v)
                        if (lane + 32 * v < n_cur) rec[v] = stage[buf * AUNIT + lane + 32 * v];
                    test_and_compact(rec, n_cur);
                    __syncwarp();   // every lane has read the buffer before it is refilled
                    unit = next;
                    buf ^= 1;
                    n_cur = n_next;
                }
            } else {
                const int nunits = (take + UNIT - 1) / UNIT;
                for (int unit = warp; unit < nunits; unit += W) {
                    int p = q0 + unit * UNIT;
                    const int p1 = min(p + UNIT, q0 + take);
                    int i = find_run(run_pos, NT, p, lane);
                    while (p < p1) {
                        const int e = run_pos[i + 1];
                        if (e <= p) { ++i; continue; }
                        const int nrec = min(e, p1) - p;
                        const RecT *src = points + run_start[i] + (long long)(p - run_pos[i]) * stride;
                        RecT rec[LPL];
#pragma unroll
                        for (int v = 0; v < LPL; ++v) {
                            const int k = lane + 32 * v;
                            if (k < nrec) rec[v] = rec_ldg<D>(src + (long long)k * stride);
                        }
                        test_and_compact(rec, nrec);
                        p += nrec;
                    }
                }
            }
                __syncthreads();
                fill = s_fill;
                q0 += take;
            }
            if (fill > 0) {
                sweep(fill);
                fill = 0;
            }
            if (stream_done) break;
        }

        // ---- merge -----------------------------------------------------------------------------
#pragma unroll
        for (int t = 0; t < kMaxT; ++t) {
            const long long r = (long long)(g0 + t) * 32 + lane;
            if (t < nt && r < P.R && m[t] < INFINITY)
                atomicMin(reinterpret_cast<unsigned *>(P.out + s * P.R + r), __float_as_uint(m[t]));
        }
        if (lane == 0 && sb == 0 && inball > 0 && P.count_work) {
            if (P.cand_count) atomicAdd(reinterpret_cast<unsigned long long *>(P.cand_count + s),
                                        (unsigned long long)inball);
            if (P.evals) atomicAdd(P.evals, (unsigned long long)inball * (unsigned long long)P.R);
        }
        executed_evals += executed * (unsigned long long)(nt * 32);
    }
    if (lane == 0 && executed_evals) atomicAdd(P.executed, executed_evals);
}

// launch with a given kernel shape
template <int D, bool PRUNE, int MAXT, int MAXW, int MINB>
int launch_eval_shape(CoverParams &P, int64_t R, cudaStream_t st) {
    using RecT = typename Rec<D>::type;
    auto kern = cover_eval_kernel<D, PRUNE, MAXT, MAXW, MINB>;
    const EvalShape sh = eval_shape(R, MAXT, MAXW);
    P.nsb = sh.nsb;
    P.groups = sh.groups;
    P.groups_per_block = sh.groups_per_block;
    const int NT = sh.W * 32;
    // tile capacity: what the shared memory of an SM allows for MINB resident CTAs (narrow CTAs
    // of small sample sets pack more per SM and get proportionally smaller tiles)
    P.async_gather = get_option("async_gather", 1) != 0;
    const size_t staging = P.async_gather ? (size_t)sh.W * 2 * 32 * kAsyncLPL * sizeof(RecT) : 0;
    const size_t fixed = (size_t)kUnroll * sizeof(RecT) + (size_t)(2 * NT + 1) * sizeof(int) + staging + 2048;
    int per_sm_target = (MAXW * MINB * 32) / NT;
    if (per_sm_target < 1) per_sm_target = 1;
    if (per_sm_target > 20) per_sm_target = 20;
    long long cap = ((long long)(227 * 1024) / per_sm_target - (long long)fixed) / (long long)sizeof(RecT);
    if (cap > 4096) cap = 4096;
    const int forced_cap = get_option("tile_cap", 0);
    if (forced_cap > 0) cap = forced_cap;
    if (cap < 2 * NT) cap = 2 * NT;
    cap = cap / kUnroll * kUnroll;
    P.tile_cap = (int)cap;
    const size_t smem = (size_t)(cap + kUnroll) * sizeof(RecT) + staging + (size_t)(2 * NT + 1) * sizeof(int);
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
    const int seed_stride = get_option("seed_stride", 32);
    if (PRUNE && seed_stride > 1) {
        // seed pass: every seed_stride-th record of the stream gives every sample an upper bound
        // of its minimum, so the full pass prunes from its first tile on
        P.stream_stride = seed_stride;
        P.count_work = 0;
        long long *main_base = P.item_base;
        if (P.item_base_seed) P.item_base = P.item_base_seed;   // chunks seed_stride times longer
        if (timed) kernel_timer_start("cover_seed", st);
        kern<<<grid, NT, smem, st>>>(P);
        count_launches(1);
        if (timed) kernel_timer_stop("cover_seed", st);
        P.item_base = main_base;
        P.queue = queue0 + 1;
    }
    P.stream_stride = 1;
    P.count_work = 1;
    kern<<<grid, NT, smem, st>>>(P);
    count_launches(1);
    P.queue = queue0;
    if (timed) kernel_timer_stop("cover_eval", st);
    FLOOD_LAUNCH_CHECK("cover_eval_kernel");
    return FLOOD_OK;
}

template <int D>
int dispatch_eval(CoverParams &P, int64_t R, cudaStream_t st) {
    const bool prune = get_option("prune", 1) != 0;
    const int sh = pick_shape(prune);
    if (!prune) {
        if (sh == 1) return launch_eval_shape<D, false, 8, 4, 5>(P, R, st);
        if (sh == 2) return launch_eval_shape<D, false, 8, 8, 2>(P, R, st);
        return launch_eval_shape<D, false, 8, 20, 1>(P, R, st);
    }
    if (sh == 1) return launch_eval_shape<D, true, 8, 4, 5>(P, R, st);
    if (sh == 2) return launch_eval_shape<D, true, 8, 8, 2>(P, R, st);
    return launch_eval_shape<D, true, 8, 20, 1>(P, R, st);
}

}  // namespace flood

// Covering-radius evaluation kernel (templates).  Included by covering.cu (host entry points) and
// by covering_d<N>.cu, which instantiate dispatch_eval<N> one ambient dimension per translation
// unit so that the dimensions compile in parallel.  See covering.cu for the description.
#pragma once
#include "common.cuh"

namespace flood {


constexpr int kUnroll = 4;       // candidates per inner-loop trip
// cp.async gather: records per lane and staging buffer (one for the 32-byte records of D >= 5)
__host__ __device__ constexpr int async_lpl(int d) { return d <= 4 ? 2 : 1; }
// pruned sweep: records each lane box-tests per trip (two independent chains for the 16-byte
// records; the 32-byte records of D >= 5 cost more registers and shared memory than the latency
// hiding returns: measured)
__host__ __device__ constexpr int box_test_ilp(int d) { return d <= 4 ? 2 : 1; }
constexpr int kSurvivorFlush = 32;   // pruned sweep: survivors collected per warp before they are swept (default)

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
    int nb;                  // bricks per sample block (shared-memory resident)
    int seg;                 // tile records per sweep task (work-stealing granularity)
    int flush;               // pruned sweep: survivors collected per warp before they are swept (multiple of 4;
                             // the per-warp buffer holds flush + 32 * box_test_ilp(D) records)
    int tile_cap;            // candidate records per shared-memory tile
    int level2;              // pruned sweep: survivors of the brick test are re-tested per pair of groups
    int slab_cull;           // pruned sweep: whole slabs of tile records are skipped by their bounding box
    int l2_bypass;           // second level: units whose pair-survivor share is >= l2_bypass/8 take the whole-brick sweep
    int off_slab;            // shared-memory offset of the slab boxes
    int off_pairbox;         // shared-memory offset of the pair boxes (16-byte aligned)
    int wbuf_stride;         // records of per-warp scratch (survivor buffer + second-level buffer)
    // dynamic shared memory layout (byte offsets; tile at 0)
    int off_stage, off_wbuf, off_bricks, off_misc, off_runs;
    int chunk;               // target tested points per chunk
    int rows_per_chunk_factor;  // chunk >= factor * (cell rows of the simplex)
};

// ---------------------------------------------------------------------------------------------
// geometry of a ball in cell coordinates (G = number of binned axes, grid_axes(D))
// ---------------------------------------------------------------------------------------------
// A "row" is a line of cells along axis 0; the rows a ball touches are the cells of its footprint on
// the other G-1 axes, numbered in memory order (axis 1 fastest).
template <int G>
struct BallCells {
    float g[G];                           // centre in cell coordinates
    float gr2;                            // (inflated) squared radius in cells
    int i0[G > 1 ? G : 2];                // first cell of the footprint on axes 1..G-1 (index 0 unused)
    int nb[G > 1 ? G : 2];                // footprint extent on axes 1..G-1
    int nrows;
};

__device__ __forceinline__ int clamp_cell(float v, int n) {
    v = fminf(fmaxf(floorf(v), -1.0f), (float)n);
    return (int)v;
}

template <int G>
__device__ __forceinline__ BallCells<G> ball_cells(const float *c, float r, const GridParams &gp) {
    BallCells<G> b;
    // inflate: the cell mapping and the ball predicate are evaluated in float32
    const float gr = r * gp.inv_h * (1.0f + 1e-5f) + 2e-3f;
    b.gr2 = gr * gr;
    b.nrows = 1;
#pragma unroll
    for (int a = 0; a < G; ++a) {
        // an axis the grid does not bin (gp.g < G: option "grid_axes") is one cell holding the centre
        b.g[a] = a < gp.g ? cell_coord(c[a], gp.origin[a], gp.inv_h) : 0.5f;
        if (a > 0) {
            const int lo = a < gp.g ? max(0, clamp_cell(b.g[a] - gr, gp.n[a])) : 0;
            const int hi = a < gp.g ? min(gp.n[a] - 1, clamp_cell(b.g[a] + gr, gp.n[a])) : 0;
            b.i0[a] = lo;
            b.nb[a] = max(0, hi - lo + 1);
            b.nrows *= b.nb[a];
        }
    }
    return b;
}

// run [a, a+len) of cell-sorted points covered by the ball in cell row `row`
template <int G>
__device__ __forceinline__ void row_run(const BallCells<G> &b, int row, const GridParams &gp,
                                        const int *__restrict__ cell_start, int &a, int &len) {
    float rem = b.gr2;
    int idx[G > 1 ? G : 2];
#pragma unroll
    for (int ax = 1; ax < G; ++ax) {
        const int i = b.i0[ax] + row % b.nb[ax];
        row /= b.nb[ax];
        idx[ax] = i;
        const float dd = fmaxf(0.f, fmaxf((float)i - b.g[ax], b.g[ax] - (float)(i + 1)));
        rem -= dd * dd;
    }
    a = 0;
    len = 0;
    if (rem < 0.f) return;
    const float half = sqrtf(rem);
    const int ix0 = max(0, clamp_cell(b.g[0] - half, gp.n[0]));
    const int ix1 = min(gp.n[0] - 1, clamp_cell(b.g[0] + half, gp.n[0]));
    if (ix0 > ix1) return;
    int base = 0;
#pragma unroll
    for (int ax = G - 1; ax >= 1; --ax) base = base * gp.n[ax] + idx[ax];
    base *= gp.n[0];
    a = __ldg(cell_start + base + ix0);
    len = __ldg(cell_start + base + ix1 + 1) - a;
}

// ---------------------------------------------------------------------------------------------
// launch shapes
// ---------------------------------------------------------------------------------------------
// The samples of a simplex are cut into bricks of at most kMaxT groups of 32 samples; the bricks
// of one sample block live in the shared memory of one CTA (nb bricks), more bricks than fit are
// split into nsb sample blocks.  Two CTA shapes:
//   wide    20 warps (exhaustive) or 16 warps (pruned), 1 CTA per SM   (many bricks: every warp
//           starts on its own brick)
//   medium   8 warps, 2 CTAs per SM  } few bricks: independent CTAs overlap each other's gather and
//   narrow   4 warps, 5 CTAs per SM  } sweep phases; the warps share the bricks segment by segment
constexpr int kMaxT = 8;            // sample groups per brick (register-resident while swept)
constexpr int kWideWarps = 20;
constexpr int kWidePrunedWarps = 16;
constexpr int kMediumWarps = 8;
constexpr int kMediumCtas = 2;
constexpr int kNarrowWarps = 4;
constexpr int kNarrowCtas = 5;
constexpr int kMaxBricks = 20;      // bricks per CTA (<= 32: one lane per brick when stealing)
// CTAs per SM of the shapes whose pruned kernel carries the second-level code (the wide shape; a
// 4 x 4-warp narrow shape with 128 registers was measured with it and dropped: r2_sweeps.md r2p-6)
__host__ __device__ constexpr bool eval_shape_has_level2(int minb) { return minb == 1; }

struct EvalShape {
    int W;        // warps per CTA
    int minb;     // CTAs per SM the kernel is compiled for (1 or kNarrowCtas)
    int nb;       // bricks per sample block
    int nsb;      // sample blocks per simplex
    int groups;   // ceil(R / 32)
    int groups_per_block;
};

// shared-memory bytes of one brick: kMaxT groups x (D coordinates + running minimum) x 32 lanes
__host__ __device__ constexpr int brick_bytes(int d) { return kMaxT * (d + 1) * 32 * 4; }

inline EvalShape eval_shape(int64_t R, int d) {
    EvalShape sh;
    sh.groups = (int)((R + 31) / 32);
    const int bricks = (sh.groups + kMaxT - 1) / kMaxT;
    const int small_max = get_option("small_max_bricks", 2);
    if (bricks <= small_max) {
        const bool narrow = get_option("small_shape", 1) == 1;   // measured: narrow beats medium in 5-D
        sh.W = narrow ? kNarrowWarps : kMediumWarps;
        sh.minb = narrow ? kNarrowCtas : kMediumCtas;
        sh.nb = bricks;
        sh.nsb = 1;
    } else {
        sh.W = kWideWarps;
        sh.minb = 1;
        int cap = (96 * 1024) / brick_bytes(d);          // at most 96 KB of samples per CTA
        if (cap > kMaxBricks) cap = kMaxBricks;
        if (cap < 1) cap = 1;
        sh.nsb = (bricks + cap - 1) / cap;
        sh.nb = 0;
    }
    sh.groups_per_block = (sh.groups + sh.nsb - 1) / sh.nsb;
    if (sh.nb == 0) sh.nb = (sh.groups_per_block + kMaxT - 1) / kMaxT;
    const int forced = get_option("warps", 0);
    if (forced > 0 && forced <= (sh.minb == 1 ? kWideWarps : (sh.minb == kMediumCtas ? kMediumWarps : kNarrowWarps)))
        sh.W = forced;
    return sh;
}

// groups of brick b of a sample block with blk_groups groups dealt over nb bricks
__host__ __device__ inline void brick_span(int blk_groups, int nb, int b, int &g_first, int &g_count) {
    const int g_base = blk_groups / nb, g_rem = blk_groups % nb;
    g_count = g_base + (b < g_rem ? 1 : 0);
    g_first = b * g_base + (b < g_rem ? b : g_rem);
}

// 3-input minimum (FMNMX3 on sm_100a)
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// order-preserving map float -> unsigned (and back): lets REDUX (integer warp reduction) take
// minima / maxima of signed floats
__device__ __forceinline__ unsigned float_key(float f) {
    const unsigned b = __float_as_uint(f);
    return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned k) {
    return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu));
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
template <int D, int NT_, int MAXT, int T0 = 0, bool PREFETCH = false>
__device__ __forceinline__ void sweep_tile(const typename Rec<D>::type *__restrict__ tile, int npad,
                                           const float (&x)[MAXT][D], float (&m)[MAXT]) {
    // sweeps groups T0 .. T0+NT_-1 of the brick held in registers.  PREFETCH (the two-group sweeps
    // of the pruned path: 28 arithmetic instructions per trip do not cover the LDS latency) loads
    // the next trip's records before the current trip's arithmetic.
    float pn[kUnroll][D];
    if (PREFETCH) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) rec_unpack<D>(tile[u], pn[u]);
    }
#pragma unroll 1
    for (int j = 0; j < npad; j += kUnroll) {
        float p[kUnroll][D];
        if (PREFETCH) {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)
#pragma unroll
                for (int a = 0; a < D; ++a) p[u][a] = pn[u][a];
            const int jn = min(j + kUnroll, npad - kUnroll);   // the last trip re-reads itself
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) rec_unpack<D>(tile[jn + u], pn[u]);
        } else {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) rec_unpack<D>(tile[j + u], p[u]);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; u += 2) {
#pragma unroll
            for (int t = T0; t + 1 < T0 + NT_; t += 2) {
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
                constexpr int t = T0 + NT_ - 1;
                m[t] = fmin3(m[t], dist2<D>(x[t], p[u]), dist2<D>(x[t], p[u + 1]));
            }
        }
    }
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

// ---------------------------------------------------------------------------------------------
// candidate stream of one work item -> shared-memory tile
// ---------------------------------------------------------------------------------------------
// Rows of the ball -> runs of the cell-sorted cloud, NT rows at a time; the runs of a batch (clipped
// to the item's window of the stream) are gathered by the warps in units of consecutive stream
// positions: coalesced record loads (cp.async double-buffered through a per-warp staging ring, or
// plain loads), the reference's ball test, an optional CTA-level cull, warp-ballot compaction
// into the tile.
template <int D>
struct Stream {
    using RecT = typename Rec<D>::type;
    static constexpr int LPL = D <= 4 ? 4 : 2;   // plain loads: records in flight per lane
    static constexpr int UNIT = 32 * LPL;
    static constexpr int ALPL = async_lpl(D);    // cp.async: records per lane and staging buffer
    static constexpr int AUNIT = 32 * ALPL;

    // per CTA (shared memory / constants)
    const RecT *points;
    RecT *tile, *stage;          // stage: this warp's ring of 2 x AUNIT records
    int *run_start, *run_pos, *warp_sums, *s_fill;
    const float *s_cbox;         // [2D] box of the CTA's samples (cull)
    const int *cell_start;
    const GridParams *gp;
    int tile_cap, stride, async_gather;
    // per item
    BallCells<grid_axes(D)> bc;
    float c[D], r2;
    int win_lo, win_hi;
    int rb, carry, q0, total2;
    bool done;
    unsigned inball;             // records inside the ball seen by this warp (lane-uniform)

    __device__ __forceinline__ void begin(const BallCells<grid_axes(D)> &b, const float (&centre)[D], float radius2, int lo, int hi) {
        bc = b;
#pragma unroll
        for (int a = 0; a < D; ++a) c[a] = centre[a];
        r2 = radius2;
        win_lo = lo;
        win_hi = hi;
        rb = carry = q0 = total2 = 0;
        done = false;
        inball = 0;
    }

    // ball test (+ cull against the CTA's sample box) of up to N records per lane, compaction
    template <bool CULL, typename RecArray>
    __device__ __forceinline__ void test_and_compact(RecArray &rec, int nrec, float U, int lane) {
        constexpr int N = sizeof(rec) / sizeof(rec[0]);
        const unsigned lt_mask = (1u << lane) - 1u;
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
                if (CULL && pass) {
                    // a record at least sqrt(U) away from the box of the CTA's samples cannot
                    // lower any of their minima
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
            keep[v] = CULL ? __ballot_sync(0xffffffffu, near) : bm;
            nkeep += __popc(keep[v]);
        }
        int wbase = 0;
        if (lane == 0 && nkeep) wbase = atomicAdd(s_fill, nkeep);
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
#pragma unroll
        for (int v = 0; v < N; ++v) {
            if ((keep[v] >> lane) & 1u) tile[wbase + __popc(keep[v] & lt_mask)] = rec[v];
            wbase += __popc(keep[v]);
        }
    }

    // Gather until the tile cannot take another round or the stream is exhausted (done).  Every
    // thread of the CTA calls it; on return all threads are past the barrier that completed the
    // tile.  Returns the new fill.
    template <bool CULL>
    __device__ __forceinline__ int fill_tile(int fill, float U) {
        const int NT = blockDim.x, W = NT >> 5;
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
        while (!done) {
            if (q0 >= total2) {
                // next batch of NT rows -> runs
                if (rb >= bc.nrows || carry >= win_hi) {
                    done = true;
                    break;
                }
                const int row = rb + tid;
                rb += NT;
                int a = 0, len = 0;
                if (row < bc.nrows) row_run(bc, row, *gp, cell_start, a, len);
                int batch_total;
                const int off = carry + block_exclusive_scan(len, warp_sums, batch_total);
                carry += batch_total;
                q0 = total2 = 0;
                if (carry <= win_lo) continue;
                // clip the run to this item's window of the stream; a seed pass (stride > 1) takes
                // every stride-th record of each clipped run
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
            if (async_gather) {
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
                    test_and_compact<CULL>(rec, n_cur, U, lane);
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
                        test_and_compact<CULL>(rec, nrec, U, lane);
                        p += nrec;
                    }
                }
            }
            __syncthreads();
            fill = *s_fill;
            q0 += take;
        }
        return fill;
    }
};

// sweep of n (a multiple of kUnroll) records by a warp holding nt sample groups (warp-uniform)
template <int D>
__device__ __forceinline__ void sweep_records(const typename Rec<D>::type *recs, int n, int nt,
                                              const float (&x)[kMaxT][D], float (&m)[kMaxT]) {
    switch (nt) {
        case 1: sweep_tile<D, 1, kMaxT>(recs, n, x, m); break;
        case 2: sweep_tile<D, 2, kMaxT>(recs, n, x, m); break;
        case 3: sweep_tile<D, 3, kMaxT>(recs, n, x, m); break;
        case 4: sweep_tile<D, 4, kMaxT>(recs, n, x, m); break;
        case 5: sweep_tile<D, 5, kMaxT>(recs, n, x, m); break;
        case 6: sweep_tile<D, 6, kMaxT>(recs, n, x, m); break;
        case 7: sweep_tile<D, 7, kMaxT>(recs, n, x, m); break;
        case 8: sweep_tile<D, 8, kMaxT>(recs, n, x, m); break;
        default: break;
    }
}

// ---------------------------------------------------------------------------------------------
// the persistent evaluation kernel
// ---------------------------------------------------------------------------------------------
// Work item = (simplex, chunk of its candidate stream, sample block), pulled from a global atomic
// queue.  Per item the CTA alternates between two phases:
//
//   gather  all warps stream the item's window of the candidate stream into the tile (Stream);
//   sweep   the tile is cut into segments of P.seg records; (brick, segment) pairs are the tasks.
//           A warp starts on "its" brick (warp % nb), claims segments from the brick's cursor
//           (shared-memory atomic) and, when the brick has none left, steals from the brick with
//           the most unclaimed segments.  While it works on a brick the warp holds the brick's
//           sample coordinates and running minima in registers (loaded from / merged back into
//           shared memory with atomicMin), so the inner loop is sweep_tile: broadcast LDS.128 of
//           the records, packed FP32x2 arithmetic, FMNMX3.
//
// PRUNE (default) skips work exactly: per task the warp tests the segment's records, a slab of
// 32 * box_test_ilp(D) at a time, against the bounding box of the brick's samples -- a record at
// least as far from the box as the brick's largest running minimum u cannot lower any minimum --
// and collects the survivors in a per-warp buffer (slabs whose own bounding box is out of reach are
// skipped unread: slab_cull).  The buffer is swept P.flush records at a time; in the wide shape
// (level2) through a second test against the boxes and bounds of the brick's pairs of groups
// (sweep_unit): per pair the survivors are compacted once more and swept with the two-group loop.
// u and the pair bounds only shrink, so a skip stays justified; the minima are bit-identical to
// the exhaustive sweep (tests/test_gpu_kernels.py::test_pruning_is_exact, test_covering_options).
// This is the "tighter candidate rule" of SURVEY.md section 8(f2): the unit of work E is still
// counted by the reference's ball rule, fewer evaluations are executed.  The same bound, taken over
// all bricks of the CTA, culls records before they enter the tile.
template <int D, bool PRUNE, int MAXW, int MINB>
__global__ void __launch_bounds__(MAXW * 32, MINB) cover_eval_kernel(const CoverParams P) {
    using RecT = typename Rec<D>::type;
    using StreamT = Stream<D>;
    constexpr int XS = (D + 1) * 32;           // floats per (brick, group): D coordinate rows + the minima row
    // the second level of the pruned sweep and the slab culling are compiled for the shapes that use
    // them (the 96-register narrow / medium kernels spill heavily with that code in them)
    constexpr bool L2 = PRUNE && eval_shape_has_level2(MINB);
    const int level2 = L2 ? P.level2 : 0;
    const bool slab_cull = L2 && P.slab_cull;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int NT = blockDim.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = NT >> 5;
    const int nb = P.nb;
    const int tile_cap = P.tile_cap;
    RecT *tile = reinterpret_cast<RecT *>(smem_raw);
    RecT *wbuf = reinterpret_cast<RecT *>(smem_raw + P.off_wbuf) + (size_t)warp * P.wbuf_stride;
    RecT *wbuf2 = wbuf + P.flush + 32 * box_test_ilp(D);   // second-level survivors (32 * box_test_ilp(D) + kUnroll records)
    float *slabbox = reinterpret_cast<float *>(smem_raw + P.off_slab);        // [tile slabs][2D] boxes of the tile's slabs
    float *bricks = reinterpret_cast<float *>(smem_raw + P.off_bricks);       // [nb][kMaxT][D+1][32]
    float *sbox = reinterpret_cast<float *>(smem_raw + P.off_misc);           // [nb][2D]
    unsigned *ub = reinterpret_cast<unsigned *>(sbox + nb * 2 * D);           // [nb] largest minimum per brick
    int *cursor = reinterpret_cast<int *>(ub + nb);                           // [nb] next unclaimed segment
    constexpr int DP = (D + 3) / 4 * 4;        // padded row of a pair box (16-byte loads)
    float *sbox2 = reinterpret_cast<float *>(smem_raw + P.off_pairbox);       // [nb][kMaxT/2][2][DP] boxes of the pairs of groups
    int *run_start = reinterpret_cast<int *>(smem_raw + P.off_runs);
    __shared__ int warp_sums[32];
    __shared__ int s_fill;
    __shared__ long long s_item[3];          // simplex, chunk, sample block (-1 = queue drained)
    __shared__ float s_cbox[2 * D + 1];      // box of all samples of the CTA, and the CTA-wide bound
    __shared__ GridParams s_gp;

    const unsigned lt_mask = (1u << lane) - 1u;
    const long long total_chunks = P.item_base[P.S];
    const unsigned long long total_items = (unsigned long long)total_chunks * (unsigned)P.nsb;

    StreamT st;
    st.points = reinterpret_cast<const RecT *>(P.points);
    st.tile = tile;
    st.stage = reinterpret_cast<RecT *>(smem_raw + P.off_stage) + (size_t)warp * 2 * StreamT::AUNIT;
    st.run_start = run_start;
    st.run_pos = run_start + NT;
    st.warp_sums = warp_sums;
    st.s_fill = &s_fill;
    st.s_cbox = s_cbox;
    st.cell_start = P.cell_start;
    st.gp = &s_gp;
    st.tile_cap = tile_cap;
    st.stride = P.stream_stride;
    st.async_gather = P.async_gather;

    if (tid == 0) {
        s_fill = 0;
        s_gp = *P.gp;
    }
    if (tid < kUnroll) tile[tile_cap + tid] = rec_sentinel<D>();   // never overwritten
    unsigned long long executed_evals = 0;   // evaluations this warp performed in the whole launch

    for (;;) {
        // ---- fetch a work item ---------------------------------------------------------------
        __syncthreads();  // previous item fully retired (s_item, tile, bricks, run arrays reusable)
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
        {
            const long long tested = P.tested[s];
            const long long nch = P.item_base[s + 1] - P.item_base[s];
            st.begin(ball_cells<grid_axes(D)>(c, rad, s_gp), c, rad * rad, (int)(chunk_j * tested / nch),
                     (int)((chunk_j + 1) * tested / nch));
        }

        // ---- the sample block's bricks -> shared memory ----------------------------------------
        // Groups of 32 consecutive samples are dealt to the bricks as evenly as possible
        // (flood_covering_bricks reports this layout, so the host can order the samples such that
        // every brick is spatially compact).  Row a < D of a group holds coordinate a of its 32
        // samples, row D their running minima (pruned mode starts from what other chunks / the
        // seed pass already found; unused slots carry 0 so that they never loosen a bound).
        const int blk_g0 = sb * P.groups_per_block;
        const int blk_groups = min(P.groups_per_block, P.groups - blk_g0);
        for (int b = warp; b < nb; b += W) {
            int gf, gc;
            brick_span(blk_groups, nb, b, gf, gc);
            float lo[D], hi[D], plo[D], phi[D], u = 0.f;
#pragma unroll
            for (int a = 0; a < D; ++a) { lo[a] = plo[a] = INFINITY; hi[a] = phi[a] = -INFINITY; }
            for (int t = 0; t < kMaxT; ++t) {
                const long long r = (long long)(blk_g0 + gf + t) * 32 + lane;
                const bool valid = t < gc && r < P.R;
                float x[D], mval = PRUNE ? 0.f : INFINITY;
#pragma unroll
                for (int a = 0; a < D; ++a) x[a] = c[a];
                if (valid) {
                    if (PRUNE) mval = __ldcg(P.out + s * P.R + r);
                    if (P.samples) {
#pragma unroll
                        for (int a = 0; a < D; ++a) x[a] = __ldg(P.samples + (s * P.R + r) * D + a);
                    } else {
                        // x = sum_k w[r,k] * v[s,k,:], FMA chain over k ascending (== the reference's
                        // float32 matmul, core.py:188)
                        const float *w = P.weights + r * P.K;
                        const float *v = P.verts + s * P.K * D;
                        const float w0 = __ldg(w);
#pragma unroll
                        for (int a = 0; a < D; ++a) x[a] = __fmul_rn(w0, __ldg(v + a));
                        for (int k = 1; k < P.K; ++k) {
                            const float wk = __ldg(w + k);
#pragma unroll
                            for (int a = 0; a < D; ++a) x[a] = fmaf(wk, __ldg(v + k * D + a), x[a]);
                        }
                    }
#pragma unroll
                    for (int a = 0; a < D; ++a) { plo[a] = fminf(plo[a], x[a]); phi[a] = fmaxf(phi[a], x[a]); }
                }
                float *g = bricks + ((size_t)b * kMaxT + t) * XS;
#pragma unroll
                for (int a = 0; a < D; ++a) g[a * 32 + lane] = x[a];
                g[D * 32 + lane] = mval;
                u = fmaxf(u, mval);
                if (PRUNE && (t & 1)) {
                    // box of the pair of groups (t-1, t): the second-level test of the pruned sweep
                    // (an empty pair keeps the inverted box, which no record is near)
#pragma unroll
                    for (int a = 0; a < D; ++a) {
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            plo[a] = fminf(plo[a], __shfl_xor_sync(0xffffffffu, plo[a], o));
                            phi[a] = fmaxf(phi[a], __shfl_xor_sync(0xffffffffu, phi[a], o));
                        }
                        if (L2 && lane == 0) {
                            sbox2[((size_t)b * (kMaxT / 2) + (t >> 1)) * 2 * DP + a] = plo[a];
                            sbox2[((size_t)b * (kMaxT / 2) + (t >> 1)) * 2 * DP + DP + a] = phi[a];
                        }
                        lo[a] = fminf(lo[a], plo[a]);
                        hi[a] = fmaxf(hi[a], phi[a]);
                        plo[a] = INFINITY;
                        phi[a] = -INFINITY;
                    }
                }
            }
            if (PRUNE) {
                if (lane == 0) {
#pragma unroll
                    for (int a = 0; a < D; ++a) { sbox[b * 2 * D + a] = lo[a]; sbox[b * 2 * D + D + a] = hi[a]; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) u = fmaxf(u, __shfl_xor_sync(0xffffffffu, u, o));
                if (lane == 0) ub[b] = __float_as_uint(u);
            }
        }
        if (tid < nb) cursor[tid] = 0;
        __syncthreads();
        float U = INFINITY;   // CTA-wide: the largest running minimum of any sample of the CTA
        if (PRUNE) {
            if (tid <= 2 * D) {
                float v;
                if (tid < D) { v = INFINITY; for (int b = 0; b < nb; ++b) v = fminf(v, sbox[b * 2 * D + tid]); }
                else if (tid < 2 * D) { v = -INFINITY; for (int b = 0; b < nb; ++b) v = fmaxf(v, sbox[b * 2 * D + tid]); }
                else { v = 0.f; for (int b = 0; b < nb; ++b) v = fmaxf(v, __uint_as_float(ub[b])); }
                s_cbox[tid] = v;
            }
            __syncthreads();
            U = s_cbox[2 * D];
        }

        // ---- gather / sweep ---------------------------------------------------------------------
        int fill = 0;
        for (;;) {
            fill = st.template fill_tile<PRUNE>(fill, U);
            if (fill > 0) {
                const int n = fill;
                if (!PRUNE) {
                    const int npad = (n + kUnroll - 1) / kUnroll * kUnroll;
                    if (tid < npad - n) tile[n + tid] = rec_sentinel<D>();
                    __syncthreads();
                }
                constexpr int SL = 32 * box_test_ilp(D);   // records per slab = per trip of the brick test
                if (slab_cull) {
                    // bounding boxes of the tile's slabs (the stream is in cell order, a slab is a
                    // few neighbouring cells): a task skips the slabs whose box is out of its brick's reach
                    const int nslab = (n + SL - 1) / SL;
                    for (int sl = warp; sl < nslab; sl += W) {
                        float lo[D], hi[D];
#pragma unroll
                        for (int a = 0; a < D; ++a) { lo[a] = INFINITY; hi[a] = -INFINITY; }
#pragma unroll
                        for (int v = 0; v < SL / 32; ++v) {
                            const int idx = sl * SL + 32 * v + lane;
                            if (idx < n) {
                                float q[D];
                                rec_unpack<D>(tile[idx], q);
#pragma unroll
                                for (int a = 0; a < D; ++a) { lo[a] = fminf(lo[a], q[a]); hi[a] = fmaxf(hi[a], q[a]); }
                            }
                        }
#pragma unroll
                        for (int a = 0; a < D; ++a) {
                            // order-preserving float -> unsigned key, one REDUX per bound
                            const unsigned kl = __reduce_min_sync(0xffffffffu, float_key(lo[a]));
                            const unsigned kh = __reduce_max_sync(0xffffffffu, float_key(hi[a]));
                            if (lane == 0) {
                                slabbox[(size_t)sl * 2 * D + a] = key_float(kl);
                                slabbox[(size_t)sl * 2 * D + D + a] = key_float(kh);
                            }
                        }
                    }
                    __syncthreads();
                }
                const int seg_len = P.seg;
                const int nseg = (n + seg_len - 1) / seg_len;
                int b = warp % nb;
                bool have = false;     // registers hold brick b
                int cnt = 0, nt = 0;   // survivors waiting in wbuf; groups of brick b
                unsigned long long swept = 0;   // records swept for brick b in this stint
                unsigned swept_groups = 0;      // second level: (record, group) pairs swept in this stint
                float x[kMaxT][D], m[kMaxT], blo[D], bhi[D], u = 0.f;
#pragma unroll
                for (int t = 0; t < kMaxT; ++t) m[t] = 0.f;
                auto bound = [&]() {
                    float v = m[0];
#pragma unroll
                    for (int t = 1; t < kMaxT; ++t) v = fmaxf(v, m[t]);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
                    return v;
                };
                // Second level of the pruned sweep: the survivors of the brick test (up to 32 records
                // at src) are tested against the boxes of the brick's pairs of groups (64 samples,
                // bound = the pair's largest running minimum), compacted per pair and swept with the
                // two-group loop.  Finer boxes and bounds skip about half of the evaluations the
                // brick-level rule alone would execute; the brick test keeps the cost of the fine
                // tests proportional to the survivors.
                constexpr int NP = kMaxT / 2;
                float uj[NP];
#pragma unroll
                for (int j = 0; j < NP; ++j) uj[j] = 0.f;
                auto pair_bound = [&](int j) {
                    // non-negative floats order like their bit patterns: one REDUX
                    return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(m[2 * j], m[2 * j + 1]))));
                };
                constexpr int TI2 = box_test_ilp(D);       // records per lane in a unit
                constexpr int UNIT2 = 32 * TI2;
                const float4 *pbox = nullptr;   // boxes of the pairs of the brick in registers (set with the brick)
                auto sweep_unit = [&](const RecT *src, int n_u) {
                    RecT rec[TI2];
                    float q[TI2][D];
                    bool inr[TI2];
                    unsigned mk[NP][TI2];
#pragma unroll
                    for (int v = 0; v < TI2; ++v) {
                        inr[v] = lane + 32 * v < n_u;
                        rec[v] = src[inr[v] ? lane + 32 * v : 0];
                        rec_unpack<D>(rec[v], q[v]);
                    }
                    // every pair unconditionally: an empty pair (brick with fewer groups) keeps an
                    // inverted box, which nothing is near, and a bound of 0
#pragma unroll
                    for (int j = 0; j < NP; ++j) {
                        float lo[DP], hi[DP];
#pragma unroll
                        for (int a4 = 0; a4 < DP / 4; ++a4) {
                            const float4 l4 = pbox[j * (2 * DP / 4) + a4];
                            const float4 h4 = pbox[j * (2 * DP / 4) + DP / 4 + a4];
                            lo[4 * a4] = l4.x; lo[4 * a4 + 1] = l4.y; lo[4 * a4 + 2] = l4.z; lo[4 * a4 + 3] = l4.w;
                            hi[4 * a4] = h4.x; hi[4 * a4 + 1] = h4.y; hi[4 * a4 + 2] = h4.z; hi[4 * a4 + 3] = h4.w;
                        }
#pragma unroll
                        for (int v = 0; v < TI2; ++v) {
                            float box2 = 0.f;
#pragma unroll
                            for (int a = 0; a < D; ++a) {
                                const float e = fmaxf(fmaxf(lo[a] - q[v][a], q[v][a] - hi[a]), 0.f);
                                box2 = fmaf(e, e, box2);
                            }
                            mk[j][v] = __ballot_sync(0xffffffffu, inr[v] && box2 * 0.9999f <= uj[j]);
                        }
                    }
                    {
                        // nearly every record is near nearly every pair (first tiles of a seed pass,
                        // bricks in sparse regions): one sweep with the whole brick, no compaction
                        int tot = 0;
#pragma unroll
                        for (int j = 0; j < NP; ++j)
#pragma unroll
                            for (int v = 0; v < TI2; ++v) tot += __popc(mk[j][v]);
                        if (tot * 8 >= P.l2_bypass * n_u * ((nt + 1) >> 1)) {
                            const int npad = (n_u + kUnroll - 1) / kUnroll * kUnroll;
                            // (a partial unit is the last of the buffer: the slots behind it are free)
                            if (lane < npad - n_u) const_cast<RecT *>(src)[n_u + lane] = rec_sentinel<D>();
                            __syncwarp();
                            sweep_records<D>(src, npad, nt, x, m);
                            swept += (unsigned)n_u;
#pragma unroll
                            for (int j = 0; j < NP; ++j) uj[j] = pair_bound(j);
                            float v = uj[0];
#pragma unroll
                            for (int j = 1; j < NP; ++j) v = fmaxf(v, uj[j]);
                            u = v;
                            return;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < NP; ++j) {
                        int c2 = 0;
#pragma unroll
                        for (int v = 0; v < TI2; ++v) c2 += __popc(mk[j][v]);
                        if (c2 == 0) continue;
                        const int pad = (-c2) & (kUnroll - 1);
                        __syncwarp();   // the previous pair's sweep has read wbuf2
                        int off = 0;
#pragma unroll
                        for (int v = 0; v < TI2; ++v) {
                            if ((mk[j][v] >> lane) & 1u) wbuf2[off + __popc(mk[j][v] & lt_mask)] = rec[v];
                            off += __popc(mk[j][v]);
                        }
                        if (lane < pad) wbuf2[c2 + lane] = rec_sentinel<D>();
                        __syncwarp();
                        if (level2 & 2) {
                            switch (j) {
                                case 0: sweep_tile<D, 2, kMaxT, 0, true>(wbuf2, c2 + pad, x, m); break;
                                case 1: sweep_tile<D, 2, kMaxT, 2, true>(wbuf2, c2 + pad, x, m); break;
                                case 2: sweep_tile<D, 2, kMaxT, 4, true>(wbuf2, c2 + pad, x, m); break;
                                default: sweep_tile<D, 2, kMaxT, 6, true>(wbuf2, c2 + pad, x, m); break;
                            }
                        } else {
                            switch (j) {
                                case 0: sweep_tile<D, 2, kMaxT, 0>(wbuf2, c2 + pad, x, m); break;
                                case 1: sweep_tile<D, 2, kMaxT, 2>(wbuf2, c2 + pad, x, m); break;
                                case 2: sweep_tile<D, 2, kMaxT, 4>(wbuf2, c2 + pad, x, m); break;
                                default: sweep_tile<D, 2, kMaxT, 6>(wbuf2, c2 + pad, x, m); break;
                            }
                        }
                        swept_groups += (unsigned)(c2 * min(2, nt - 2 * j));
                        uj[j] = pair_bound(j);
                    }
                    float v = uj[0];
#pragma unroll
                    for (int j = 1; j < NP; ++j) v = fmaxf(v, uj[j]);
                    u = v;
                };
                // Exhaustive sweep: segments are claimed one ahead (lane 0 holds the claim; it is
                // broadcast when it is needed), so the atomic's latency overlaps the current segment.
                // Pruned sweep: segments cost anything between nothing and a full sweep, a hoarded
                // segment unbalances the end of the tile (measured: +7 %), so it claims on demand.
                // (The same holds when a few bricks are shared by all warps: no claiming ahead.)
                const bool ahead = !PRUNE && nb >= W;
                int claim = 0;
                if (ahead && lane == 0) claim = atomicAdd(&cursor[b], 1);
                for (;;) {
                    if (!ahead && lane == 0) claim = atomicAdd(&cursor[b], 1);
                    const int seg = __shfl_sync(0xffffffffu, claim, 0);
                    if (seg >= nseg) {
                        if (have) {
                            // end of the stint on brick b: flush the survivors, merge the minima
                            if (PRUNE && cnt > 0) {
                                __syncwarp();
                                if (level2) {
                                    for (int k0 = 0; k0 < cnt; k0 += UNIT2) sweep_unit(wbuf + k0, min(UNIT2, cnt - k0));
                                } else {
                                    const int npad = (cnt + kUnroll - 1) / kUnroll * kUnroll;
                                    if (lane < npad - cnt) wbuf[cnt + lane] = rec_sentinel<D>();
                                    __syncwarp();
                                    sweep_records<D>(wbuf, npad, nt, x, m);
                                    swept += (unsigned)cnt;
                                }
                                __syncwarp();
                                cnt = 0;
                            }
#pragma unroll
                            for (int t = 0; t < kMaxT; ++t)
                                if (t < nt)
                                    atomicMin(reinterpret_cast<unsigned *>(bricks + ((size_t)b * kMaxT + t) * XS + D * 32 + lane),
                                              __float_as_uint(m[t]));
                            if (PRUNE) {
                                const float v = bound();
                                if (lane == 0) atomicMin(&ub[b], __float_as_uint(v));
                            }
                            executed_evals += swept * (unsigned long long)(nt * 32) + 32ull * swept_groups;
                            swept = 0;
                            swept_groups = 0;
                            have = false;
                        }
                        // steal from the brick with the most unclaimed segments
                        int best = lane < nb ? nseg - cursor[lane] : 0, who = lane;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            const int ob = __shfl_xor_sync(0xffffffffu, best, o);
                            const int ow = __shfl_xor_sync(0xffffffffu, who, o);
                            if (ob > best || (ob == best && ow < who)) { best = ob; who = ow; }
                        }
                        if (best <= 0) break;
                        b = who;
                        if (ahead && lane == 0) claim = atomicAdd(&cursor[b], 1);
                        continue;
                    }
                    if (ahead && lane == 0) claim = atomicAdd(&cursor[b], 1);   // the next segment of this brick
                    if (!have) {
                        int gf;
                        brick_span(blk_groups, nb, b, gf, nt);
#pragma unroll
                        for (int t = 0; t < kMaxT; ++t) {
                            const float *g = bricks + ((size_t)b * kMaxT + t) * XS;
#pragma unroll
                            for (int a = 0; a < D; ++a) x[t][a] = g[a * 32 + lane];
                            m[t] = g[D * 32 + lane];
                        }
                        if (PRUNE) {
#pragma unroll
                            for (int a = 0; a < D; ++a) { blo[a] = sbox[b * 2 * D + a]; bhi[a] = sbox[b * 2 * D + D + a]; }
                            u = bound();
#pragma unroll
                            for (int j = 0; j < NP; ++j) uj[j] = pair_bound(j);
                            pbox = reinterpret_cast<const float4 *>(sbox2 + (size_t)b * NP * 2 * DP);
                        }
                        have = true;
                    }
                    const int lo = seg * seg_len, hi = min(n, lo + seg_len);
                    if (!PRUNE) {
                        sweep_records<D>(tile + lo, (hi - lo + kUnroll - 1) / kUnroll * kUnroll, nt, x, m);
                        swept += (unsigned)(hi - lo);
                    }
                    // box test, two records per lane and trip (independent chains hide each other's
                    // latency: with 5 warps per sub-partition the test loop is latency-bound)
                    constexpr int TI = box_test_ilp(D);
                    unsigned live = 0xffffffffu;   // slabs of the segment (bit k = records lo + k*SL ...) the brick can reach
                    if (slab_cull) {
                        const int ns = (hi - lo + SL - 1) / SL;        // <= 32 (P.seg <= 32 * SL)
                        bool reach = lane < ns;
                        if (reach) {
                            const float *sb = slabbox + (size_t)(lo / SL + lane) * 2 * D;
                            float box2 = 0.f;
#pragma unroll
                            for (int a = 0; a < D; ++a) {
                                const float e = fmaxf(fmaxf(blo[a] - sb[D + a], sb[a] - bhi[a]), 0.f);
                                box2 = fmaf(e, e, box2);
                            }
                            reach = box2 * 0.9999f <= u;   // every record of the slab is at least that far
                        }
                        live = __ballot_sync(0xffffffffu, reach);
                    }
                    for (int base = lo, k = 0; PRUNE && base < hi; base += SL, ++k) {
                        if (!((live >> k) & 1u)) continue;
                        RecT rec[TI];
                        bool keep[TI];
#pragma unroll
                        for (int v = 0; v < TI; ++v) {
                            const int idx = base + 32 * v + lane;
                            rec[v] = tile[idx < hi ? idx : tile_cap];
                            float own[D];
                            rec_unpack<D>(rec[v], own);
                            float box2 = 0.f;
#pragma unroll
                            for (int a = 0; a < D; ++a) {
                                const float e = fmaxf(fmaxf(blo[a] - own[a], own[a] - bhi[a]), 0.f);
                                box2 = fmaf(e, e, box2);
                            }
                            // 0.9999: the box distance and the pair distances are rounded differently
                            keep[v] = idx < hi && box2 * 0.9999f <= u;
                        }
                        unsigned mask[TI], any = 0u;
#pragma unroll
                        for (int v = 0; v < TI; ++v) {
                            mask[v] = __ballot_sync(0xffffffffu, keep[v]);
                            any |= mask[v];
                        }
                        if (any == 0u) continue;
#pragma unroll
                        for (int v = 0; v < TI; ++v) {
                            if (keep[v]) wbuf[cnt + __popc(mask[v] & lt_mask)] = rec[v];
                            cnt += __popc(mask[v]);
                        }
                        if (cnt >= P.flush) {
                            int nsweep = P.flush;          // whole multiples of P.flush (at most a few: no division)
                            while (nsweep + P.flush <= cnt) nsweep += P.flush;
                            __syncwarp();
                            if (level2) {
                                for (int k0 = 0; k0 < nsweep; k0 += UNIT2) sweep_unit(wbuf + k0, UNIT2);
                                __syncwarp();
                            } else {
                                sweep_records<D>(wbuf, nsweep, nt, x, m);
                                swept += (unsigned)nsweep;
                            }
                            cnt -= nsweep;
                            // carry the rest (< flush) to the front, 32 records at a time (a chunk's
                            // destination ends where its source begins or earlier: nsweep >= 32)
                            for (int k0 = 0; k0 < cnt; k0 += 32) {
                                RecT carry_rec;
                                if (k0 + lane < cnt) carry_rec = wbuf[nsweep + k0 + lane];
                                __syncwarp();
                                if (k0 + lane < cnt) wbuf[k0 + lane] = carry_rec;
                                __syncwarp();
                            }
                            if (!level2) u = bound();
                        }
                    }
                }
                // ---- closing: every stint merged, tile free ------------------------------------
                __syncthreads();
                if (tid < nb) cursor[tid] = 0;
                if (tid == 0) s_fill = 0;
                if (PRUNE && warp == 0) {
                    float v = lane < nb ? __uint_as_float(ub[lane]) : 0.f;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
                    if (lane == 0) s_cbox[2 * D] = v;
                }
                __syncthreads();
                if (PRUNE) U = s_cbox[2 * D];
                fill = 0;
            }
            if (st.done) break;
        }

        // ---- merge -----------------------------------------------------------------------------
        for (int b = warp; b < nb; b += W) {
            int gf, gc;
            brick_span(blk_groups, nb, b, gf, gc);
            for (int t = 0; t < gc; ++t) {
                const long long r = (long long)(blk_g0 + gf + t) * 32 + lane;
                const float v = bricks[((size_t)b * kMaxT + t) * XS + D * 32 + lane];
                if (r < P.R && v < INFINITY)
                    atomicMin(reinterpret_cast<unsigned *>(P.out + s * P.R + r), __float_as_uint(v));
            }
        }
        if (lane == 0 && sb == 0 && st.inball > 0 && P.count_work) {
            if (P.cand_count) atomicAdd(reinterpret_cast<unsigned long long *>(P.cand_count + s),
                                        (unsigned long long)st.inball);
            if (P.evals) atomicAdd(P.evals, (unsigned long long)st.inball * (unsigned long long)P.R);
        }
    }
    if (lane == 0 && executed_evals) atomicAdd(P.executed, executed_evals);
}

// launch (seed pass + full pass in pruned mode)
template <int D, bool PRUNE, int MAXW, int MINB>
int launch_eval_shape(CoverParams &P, const EvalShape &sh, cudaStream_t st) {
    using RecT = typename Rec<D>::type;
    auto kern = cover_eval_kernel<D, PRUNE, MAXW, MINB>;
    P.nsb = sh.nsb;
    P.nb = sh.nb;
    P.groups = sh.groups;
    P.groups_per_block = sh.groups_per_block;
    const int NT = sh.W * 32;
    // cp.async staging pays (marginally) for the 16-byte records of D <= 4; the 32-byte records of
    // D >= 5 leave room for one record per lane and buffer only, where plain loads are faster
    P.async_gather = get_option("async_gather", D <= 4 ? 1 : 0) != 0;
    // sweep task = one brick x P.seg tile records; the exhaustive sweep has uniform tasks and
    // prefers long ones (fewer loop prologues), the pruned sweep short ones (balance)
    P.seg = get_option("seg", PRUNE ? 256 : 1024);
    // dynamic shared memory: tile | staging rings / survivor buffers | bricks | boxes, bounds, cursors | runs
    const size_t staging = P.async_gather ? (size_t)sh.W * 2 * 32 * async_lpl(D) * sizeof(RecT) : 0;
    P.flush = get_option("flush", kSurvivorFlush);
    if (P.flush < 32) P.flush = 32;
    if (P.flush > 512) P.flush = 512;
    P.flush = P.flush / kUnroll * kUnroll;
    // second level of the pruned sweep (per pair of groups); it takes the survivors 32 at a time
    // level2: 0 off, 1 on, 3 on with prefetching two-group sweeps.  Default: on for the wide shape
    // (measured in 5-D with the narrow shape: 2.5 x fewer evaluations executed, 10 % slower)
    P.level2 = PRUNE && eval_shape_has_level2(MINB) ? get_option("level2", sh.minb == 1 ? 3 : 0) : 0;
    P.l2_bypass = get_option("l2_bypass", 7);
    P.slab_cull = PRUNE && eval_shape_has_level2(MINB) && get_option("slab_cull", sh.minb == 1 ? 1 : 0) != 0;   // (narrow shape in 5-D: 5 % slower with it)
    constexpr int SL = 32 * box_test_ilp(D);     // records per unit of the second level = per slab of the tile
    if (P.level2) {
        if (get_option("flush", 0) <= 0) P.flush = SL;
        P.flush = (P.flush + SL - 1) / SL * SL;
    }
    P.wbuf_stride = P.flush + SL + (P.level2 ? SL + kUnroll : 0);
    const size_t wbuf = PRUNE ? (size_t)sh.W * P.wbuf_stride * sizeof(RecT) : 0;
    const size_t bricks = (size_t)sh.nb * brick_bytes(D);
    constexpr int DPh = (D + 3) / 4 * 4;
    const size_t misc = ((size_t)sh.nb * (2 * D + 2) * 4 + 15) / 16 * 16;
    const size_t pairbox = (size_t)sh.nb * (kMaxT / 2) * 2 * DPh * 4;
    const size_t runs = ((size_t)(2 * NT + 1) * sizeof(int) + 15) / 16 * 16;
    // the staging rings are live only while gathering, the survivor buffers only while sweeping:
    // they share one region
    const size_t scratch = staging > wbuf ? staging : wbuf;
    const size_t fixed = (size_t)kUnroll * sizeof(RecT) + scratch + bricks + misc + pairbox + runs + (P.slab_cull ? 16 + 2 * D * 4 : 0);
    const long long budget = (long long)(227 * 1024) / sh.minb - 2048;   // static shared memory + per-CTA reserve
    // a tile record costs its bytes plus its share of a slab box
    long long cap = (budget - (long long)fixed) * SL / ((long long)sizeof(RecT) * SL + (P.slab_cull ? 2 * D * 4 : 0));
    const int cap_max = get_option("tile_cap_max", 8192);
    if (cap > cap_max) cap = cap_max;
    const int forced_cap = get_option("tile_cap", 0);   // experiments: a smaller tile (at least 2 records per thread)
    if (forced_cap > 0 && forced_cap < cap) cap = forced_cap > 2 * NT ? forced_cap : (cap < 2 * NT ? cap : 2 * NT);
    cap = cap / kUnroll * kUnroll;
    if (cap < 2 * NT)
        return set_error(FLOOD_E_UNSUPPORTED, "cover_eval_kernel: shared memory budget exceeded (d=%d, %d bricks)", D, sh.nb);
    P.tile_cap = (int)cap;
    {
        // at least ~4 tasks per warp and tile, also when a few bricks are shared by all warps
        const long long balanced = (cap * sh.nb / (4 * sh.W) + 31) / 32 * 32;
        if (get_option("seg", 0) <= 0 && P.seg > balanced) P.seg = (int)balanced;
        if (P.seg < 32) P.seg = 32;
        P.seg = P.seg / kUnroll * kUnroll;
        if (PRUNE) {
            // pruned sweep: a segment is a whole number of slabs (at most 32: one lane per slab)
            P.seg = (P.seg + SL - 1) / SL * SL;
            if (P.seg > 32 * SL) P.seg = 32 * SL;
        }
    }
    size_t o = (size_t)(cap + kUnroll) * sizeof(RecT);
    P.off_stage = (int)o;
    P.off_wbuf = (int)o;    o += scratch;
    P.off_bricks = (int)o;  o += bricks;
    P.off_misc = (int)o;    o += misc;
    P.off_pairbox = (int)o; o += pairbox;
    P.off_runs = (int)o;    o += runs;
    P.off_slab = (int)o;
    if (P.slab_cull) o += ((size_t)(cap / SL + 1) * 2 * D * 4 + 15) / 16 * 16;
    const size_t smem = o;
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
    const EvalShape sh = eval_shape(R, D);
    if (sh.minb == 1) {
        // The exhaustive sweep is bound by the FMA pipe and wants every warp it can get (20 x 96
        // registers); the pruned sweep is a chain of short dependent phases (box tests, compaction,
        // two-group sweeps) that the compiler schedules much better with 128 registers: 16 warps
        // (measured: 20 warps 52 ms, 16 warps 42.6 ms, 12 warps 43.5 ms on torus 1 M / 1 k).
        // (20, 18, 14 and 12 warps for the pruned sweep and 16 for the exhaustive one were measured too
        // and are not compiled in: profiles/r2_sweeps.md, r2p-3, r2p-4 and r2p-7.)
        if (prune) {
            EvalShape sh2 = sh;
            if (sh2.W > kWidePrunedWarps) sh2.W = kWidePrunedWarps;
            return launch_eval_shape<D, true, kWidePrunedWarps, 1>(P, sh2, st);
        }
        return launch_eval_shape<D, false, kWideWarps, 1>(P, sh, st);
    }
    if (sh.minb == kMediumCtas) {
        if (prune) return launch_eval_shape<D, true, kMediumWarps, kMediumCtas>(P, sh, st);
        return launch_eval_shape<D, false, kMediumWarps, kMediumCtas>(P, sh, st);
    }
    if (prune) return launch_eval_shape<D, true, kNarrowWarps, kNarrowCtas>(P, sh, st);
    return launch_eval_shape<D, false, kNarrowWarps, kNarrowCtas>(P, sh, st);
}

}  // namespace flood

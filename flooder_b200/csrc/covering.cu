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
//     lengths and its warps stream their window of it in units of consecutive positions
//     (coalesced loads, cp.async double-buffered for the 16-byte records of D <= 4);
//   * every gathered point is tested against the ball (the reference predicate) and the
//     survivors are compacted with warp ballots into a shared-memory tile of records;
//   * the samples of the block live in shared memory as bricks of up to 8 groups of 32; when the
//     tile is full, (brick, tile segment) pairs are tasks that warps claim and steal; a warp
//     holds the brick it works on in registers, tile records are broadcast LDS.128 reads, the
//     distance is the direct difference form in FP32 (D sub, 1 mul, D-1 fma, 1 min per pair),
//     issued as packed FP32x2 instructions and 3-input minima (sweep_tile);
//   * by default the sweep is pruned exactly, on two levels: per task the records that are at
//     least as far from the box of the brick's samples as the brick's largest running minimum are
//     skipped (whole slabs of the tile by their bounding box first), the survivors are tested in
//     the same way against the brick's pairs of groups and swept pair by pair with a prefetching
//     two-group loop; a seed pass over a sub-sampled stream gives every sample a finite bound first;
//   * the minima are merged into min_dist2 with an unsigned atomicMin (non-negative floats order
//     like their bit patterns), because a simplex may be split over several chunks.
//
// This file holds the host entry points and the small planning kernels; the evaluation kernel
// is in covering_kernels.cuh and is instantiated per ambient dimension in covering_d<N>.cu.
//
// The dense mask, the index lists and the (S, R, D) sample tensor never exist.
#include "covering_kernels.cuh"

namespace flood {

extern template int dispatch_eval<1>(CoverParams &, int64_t, cudaStream_t);
extern template int dispatch_eval<2>(CoverParams &, int64_t, cudaStream_t);
extern template int dispatch_eval<3>(CoverParams &, int64_t, cudaStream_t);
extern template int dispatch_eval<4>(CoverParams &, int64_t, cudaStream_t);
extern template int dispatch_eval<5>(CoverParams &, int64_t, cudaStream_t);
extern template int dispatch_eval<6>(CoverParams &, int64_t, cudaStream_t);
extern template int dispatch_eval<7>(CoverParams &, int64_t, cudaStream_t);
extern template int dispatch_eval<8>(CoverParams &, int64_t, cudaStream_t);

namespace {

// ---------------------------------------------------------------------------------------------
// plan: tested[s] = length of the candidate stream of simplex s (one warp per simplex)
// ---------------------------------------------------------------------------------------------
template <int G>
__global__ void cover_plan_kernel(CoverParams P, int d) {
    const GridParams gp = *P.gp;
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (warp >= P.S) return;
    float c[G];
#pragma unroll
    for (int a = 0; a < G; ++a) c[a] = P.centers[warp * d + a];
    const BallCells<G> b = ball_cells<G>(c, P.radii[warp], gp);
    long long total = 0;
    for (int row = lane; row < b.nrows; row += 32) {
        int a, len;
        row_run<G>(b, row, gp, P.cell_start, a, len);
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
        if (P.item_base_seed) {
            const long long chunk2 = max((long long)P.chunk_seed, (long long)P.rows_per_chunk_factor * b.nrows);
            P.item_base_seed[warp] = (total + chunk2 - 1) / chunk2;
        }
    }
}

void launch_plan(const CoverParams &P, int d, int64_t S, cudaStream_t st) {
    const int threads = 128;  // 4 simplices per CTA
    const unsigned blocks = (unsigned)((S * 32 + threads - 1) / threads);
    switch (grid_axes(d)) {
        case 1: cover_plan_kernel<1><<<blocks, threads, 0, st>>>(P, d); break;
        case 2: cover_plan_kernel<2><<<blocks, threads, 0, st>>>(P, d); break;
        case 3: cover_plan_kernel<3><<<blocks, threads, 0, st>>>(P, d); break;
        case 4: cover_plan_kernel<4><<<blocks, threads, 0, st>>>(P, d); break;
        default: cover_plan_kernel<kMaxGridAxes><<<blocks, threads, 0, st>>>(P, d); break;
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

__global__ void fill_inf_kernel(float *p, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        p[i] = INFINITY;
}

struct CoverLayout {
    int64_t off_tested, off_item_base, off_item_base_seed, off_queue, total;
};

CoverLayout cover_layout(int64_t S) {
    CoverLayout L;
    int64_t o = 0;
    L.off_queue = o;      o = align_up(o + 64, 256);
    L.off_tested = o;     o = align_up(o + S * 4, 256);
    L.off_item_base = o;  o = align_up(o + (S + 1) * 8, 256);
    L.off_item_base_seed = o;  o = align_up(o + (S + 1) * 8, 256);
    L.total = o;
    return L;
}

}  // namespace

// cost estimate per simplex: length of its candidate stream (points in the cell rows its ball
// touches); used by the host to balance simplices over GPUs
int covering_plan(const void *cloud_ws, int64_t n, int d, const float *centers, const float *radii,
                  int64_t S, int32_t *out_tested, cudaStream_t st) {
    if (S == 0) return FLOOD_OK;
    if (!cloud_ws || !centers || !radii || !out_tested || S < 0 || n < 1 || d < 1 || d > FLOOD_MAX_DIM)
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
    launch_plan(P, d, S, st);
    count_launches(1);
    FLOOD_LAUNCH_CHECK("cover_plan_kernel");
    return FLOOD_OK;
}

// Sample layout of the evaluation kernel for R samples: brick i (the samples one warp keeps in
// registers) holds out_groups[i] consecutive groups of 32 samples; bricks_per_block consecutive
// bricks form the sample block of one CTA.  Returns the number of bricks.
int covering_bricks(int64_t R, int d, int32_t *out_groups, int capacity, int *bricks_per_block) {
    if (R < 1 || d < 1 || d > FLOOD_MAX_DIM)
        return set_error(FLOOD_E_INVALID, "covering_bricks: bad arguments (R=%lld d=%d)", (long long)R, d);
    const EvalShape sh = eval_shape(R, d);
    const int total = sh.nsb * sh.nb;
    if (bricks_per_block) *bricks_per_block = sh.nb;
    if (!out_groups) return total;
    if (capacity < total)
        return set_error(FLOOD_E_INVALID, "covering_bricks: capacity %d < %d bricks", capacity, total);
    for (int sb = 0; sb < sh.nsb; ++sb) {
        const int blk_g0 = sb * sh.groups_per_block;
        int blk_groups = sh.groups - blk_g0;
        if (blk_groups > sh.groups_per_block) blk_groups = sh.groups_per_block;
        if (blk_groups < 0) blk_groups = 0;
        for (int b = 0; b < sh.nb; ++b) {
            int gf, gc;
            brick_span(blk_groups, sh.nb, b, gf, gc);
            out_groups[sb * sh.nb + b] = gc;
        }
    }
    return total;
}

// work items of a call: tested[s], the scanned chunk counts item_base[], a cleared queue (shared by
// the float32 kernel's caller below and the float64 path in f64.cu)
int covering_plan_items(CoverParams &P, int d, int64_t S, void *ws, cudaStream_t st) {
    const CoverLayout L = cover_layout(S);
    char *wbase = static_cast<char *>(ws);
    P.tested = reinterpret_cast<int *>(wbase + L.off_tested);
    P.item_base = reinterpret_cast<long long *>(wbase + L.off_item_base);
    P.item_base_seed = nullptr;
    P.queue = reinterpret_cast<unsigned long long *>(wbase + L.off_queue);
    P.executed = P.queue + 2;
    P.chunk = get_option("chunk", 16384);
    if (P.chunk < 256) P.chunk = 256;
    P.rows_per_chunk_factor = get_option("rows_per_chunk_factor", 32);
    FLOOD_CUDA_CHECK(cudaMemsetAsync(P.queue, 0, 64, st));
    launch_plan(P, d, S, st);
    cover_scan_kernel<<<1, 1024, 0, st>>>(P.item_base, S);
    count_launches(2);
    FLOOD_LAUNCH_CHECK("cover plan kernels");
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
        d < 1 || d > FLOOD_MAX_DIM || K < 1 || K > FLOOD_MAX_SIMPLEX_VERTS)
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

    CoverParams P = {};
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
    // target tested points per work item: the pruned sweep executes a fraction of a chunk's
    // evaluations and amortises its per-item work (brick load, bounds, merges) over longer chunks
    const bool prune = get_option("prune", 1) != 0;
    const int chunk_default = prune ? 32768 : 16384;
    {
        // the seed pass (pruned mode) takes every seed_stride-th record: its chunks are that much
        // longer, so that a seed item carries as many records as an item of the full pass
        const long long seed_stride = get_option("seed_stride", 32);
        const long long mult = get_option("seed_chunk_mult", 0) > 0 ? get_option("seed_chunk_mult", 0) : seed_stride;
        if (prune && seed_stride > 1 && mult > 1) {
            P.item_base_seed = reinterpret_cast<long long *>(wbase + L.off_item_base_seed);
            long long cs = (long long)(get_option("chunk", chunk_default) < 256 ? 256 : get_option("chunk", chunk_default)) * mult;
            P.chunk_seed = (int)(cs > (1ll << 30) ? (1ll << 30) : cs);
        }
    }
    P.queue = reinterpret_cast<unsigned long long *>(wbase + L.off_queue);
    P.executed = P.queue + 2;
    P.S = S;
    P.R = R;
    P.K = K;
    P.nsb = 1;
    P.chunk = get_option("chunk", chunk_default);
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
        launch_plan(P, d, S, st);
        cover_scan_kernel<<<1, 1024, 0, st>>>(P.item_base, S);
        if (P.item_base_seed) cover_scan_kernel<<<1, 1024, 0, st>>>(P.item_base_seed, S);
        count_launches(P.item_base_seed ? 4 : 3);   // fill, plan, scan (+ scan)
    }
    FLOOD_LAUNCH_CHECK("cover plan kernels");
    switch (d) {
        case 1: return dispatch_eval<1>(P, R, st);
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

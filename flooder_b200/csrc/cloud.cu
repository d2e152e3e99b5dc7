// Cloud preparation: uniform cell grid + cell-sorted padded point records.
//
// Replaces the reference's 1-D ordering of the cloud (flooder/core.py:140-144: argsort along
// the widest axis, then a per-batch slab via searchsorted, :201-208).  A 3-D cell order makes
// the candidates of a bounding ball a handful of contiguous runs (one per (y,z) cell row), so
// the covering-radius kernel streams them without the dense mask / nonzero / gather pipeline.
#include "common.cuh"

namespace flood {
namespace {

__device__ __forceinline__ unsigned ordered_bits(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

constexpr int GA = kMaxGridAxes;

__global__ void bbox_init_kernel(unsigned *bbox) {
    int t = threadIdx.x;
    if (t < GA) bbox[t] = 0xffffffffu;          // running minima
    else if (t < 2 * GA) bbox[t] = 0u;          // running maxima
}

// bounding box over the binned coordinates
__global__ void bbox_kernel(const float *__restrict__ pts, int64_t n, int d, unsigned *bbox) {
    const int g = grid_axes(d);
    float lo[GA], hi[GA];
    for (int a = 0; a < GA; ++a) { lo[a] = INFINITY; hi[a] = -INFINITY; }
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        for (int a = 0; a < g; ++a) {
            float v = pts[i * d + a];
            lo[a] = fminf(lo[a], v);
            hi[a] = fmaxf(hi[a], v);
        }
    }
    for (int a = 0; a < g; ++a) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if ((threadIdx.x & 31) == 0 && lo[a] <= hi[a]) {
            atomicMin(&bbox[a], ordered_bits(lo[a]));
            atomicMax(&bbox[GA + a], ordered_bits(hi[a]));
        }
    }
}

// Choose a cubic cell edge h so that the box holds ~ n / points_per_cell cells.
__global__ void grid_params_kernel(const unsigned *bbox, int64_t n, int d, int points_per_cell,
                                   int64_t max_cells, int axes_limit, GridParams *gp) {
    int g = grid_axes(d);
    if (axes_limit > 0 && axes_limit < g) g = axes_limit;
    float lo[GA], ext[GA];
    double vol = 1.0;
    int live = 0;
    for (int a = 0; a < GA; ++a) { lo[a] = 0.f; ext[a] = 0.f; }
    for (int a = 0; a < g; ++a) {
        lo[a] = from_ordered_bits(bbox[a]);
        ext[a] = from_ordered_bits(bbox[GA + a]) - lo[a];
        if (ext[a] > 0.f) { vol *= (double)ext[a]; ++live; }
    }
    double target = (double)n / (double)(points_per_cell > 0 ? points_per_cell : 32);
    if (target < 1.0) target = 1.0;
    if (target > (double)max_cells) target = (double)max_cells;
    float h = 1.0f;
    int nc[GA];
    for (int a = 0; a < GA; ++a) nc[a] = 1;
    if (live > 0) {
        h = (float)pow(vol / target, 1.0 / (double)live);
        if (!(h > 0.f)) h = 1.0f;
        // grow h until every axis fits the 10-bit cell coordinate and the total fits the tables
        for (int iter = 0; iter < 400; ++iter) {
            int64_t total = 1;
            bool fits = true;
            for (int a = 0; a < GA; ++a) {
                int c = 1;
                if (a < g && ext[a] > 0.f) {
                    const float q = ext[a] / h;
                    if (q >= 1023.f) { fits = false; c = 1024; } else c = (int)q + 1;
                }
                nc[a] = c;
                total *= c;
            }
            if (fits && total <= max_cells) break;
            h *= 1.15f;
        }
    }
    int ncells = 1;
    for (int a = 0; a < GA; ++a) {
        gp->origin[a] = lo[a];
        gp->n[a] = nc[a];
        ncells *= nc[a];
    }
    gp->h = h;
    gp->inv_h = 1.0f / h;
    gp->ncells = ncells;
    gp->g = g;
    gp->npts = n;
    gp->d = d;
}

__device__ __forceinline__ int cell_of_point(const float *p, int d, const GridParams &gp) {
    int c = 0;
    for (int a = gp.g - 1; a >= 0; --a)
        c = c * gp.n[a] + cell_clamp(cell_coord(p[a], gp.origin[a], gp.inv_h), gp.n[a]);
    return c;
}

__global__ void cell_count_kernel(const float *__restrict__ pts, int64_t n, int d,
                                  const GridParams *__restrict__ gpp, int *cell_id, int *cell_fill) {
    const GridParams gp = *gpp;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        int c = cell_of_point(pts + i * d, d, gp);
        cell_id[i] = c;
        atomicAdd(&cell_fill[c], 1);
    }
}

// single-CTA exclusive scan of the cell counts; clears the counts for the scatter pass
__global__ void cell_scan_kernel(const GridParams *__restrict__ gpp, int *cell_fill, int *cell_start) {
    __shared__ int warp_sums[32];
    __shared__ int carry_s;
    const int ncells = gpp->ncells;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < ncells; base += blockDim.x) {
        int i = base + threadIdx.x;
        int v = i < ncells ? cell_fill[i] : 0;
        int x = v;
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = lane < nw ? warp_sums[lane] : 0;
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            warp_sums[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        int carry = carry_s;
        int excl = carry + (warp > 0 ? warp_sums[warp - 1] : 0) + x - v;
        if (i < ncells) { cell_start[i] = excl; cell_fill[i] = 0; }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry_s = carry + warp_sums[nw - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) cell_start[ncells] = carry_s;
}

template <int PD>
__global__ void scatter_kernel(const float *__restrict__ pts, int64_t n, int d,
                               const int *__restrict__ cell_id, const int *__restrict__ cell_start,
                               int *cell_fill, int *__restrict__ perm, float *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        int c = cell_id[i];
        int64_t pos = cell_start[c] + atomicAdd(&cell_fill[c], 1);
        perm[pos] = (int)i;
        float rec[PD];
#pragma unroll
        for (int a = 0; a < PD; ++a) rec[a] = a < d ? pts[i * d + a] : 0.f;
        if (PD == 2) {
            reinterpret_cast<float2 *>(out)[pos] = make_float2(rec[0], rec[1]);
        } else {
#pragma unroll
            for (int q = 0; q < PD / 4; ++q)
                reinterpret_cast<float4 *>(out)[pos * (PD / 4) + q] =
                    make_float4(rec[4 * q], rec[4 * q + 1], rec[4 * q + 2], rec[4 * q + 3]);
        }
    }
}

}  // namespace

int cloud_build(const float *pts, int64_t n, int d, int points_per_cell, void *ws, size_t ws_bytes,
                cudaStream_t st) {
    if (!pts || !ws || n <= 0 || d < 1 || d > FLOOD_MAX_DIM)
        return set_error(FLOOD_E_INVALID, "cloud_build: bad arguments (n=%lld, d=%d)", (long long)n, d);
    if (n >= (int64_t(1) << 31))
        return set_error(FLOOD_E_UNSUPPORTED, "cloud_build: n=%lld exceeds 2^31-1", (long long)n);
    const CloudLayout L = cloud_layout(n, d);
    if ((int64_t)ws_bytes < L.total)
        return set_error(FLOOD_E_WORKSPACE, "cloud_build: workspace %zu < %lld bytes", ws_bytes,
                         (long long)L.total);
    if (points_per_cell <= 0) points_per_cell = get_option("points_per_cell", 32);
    char *base = static_cast<char *>(ws);
    GridParams *gp = reinterpret_cast<GridParams *>(base + L.off_grid);
    unsigned *bbox = reinterpret_cast<unsigned *>(base + L.off_bbox);
    int *cell_start = reinterpret_cast<int *>(base + L.off_cell_start);
    int *cell_fill = reinterpret_cast<int *>(base + L.off_cell_fill);
    int *cell_id = reinterpret_cast<int *>(base + L.off_cell_id);
    int *perm = reinterpret_cast<int *>(base + L.off_perm);
    float *out = reinterpret_cast<float *>(base + L.off_points);

    const int threads = 256;
    const int sms = device_sm_count();
    int blocks = (int)((n + threads - 1) / threads);
    if (blocks > sms * 8) blocks = sms * 8;

    bbox_init_kernel<<<1, 32, 0, st>>>(bbox);
    bbox_kernel<<<blocks, threads, 0, st>>>(pts, n, d, bbox);
    grid_params_kernel<<<1, 1, 0, st>>>(bbox, n, d, points_per_cell, L.max_cells, get_option("grid_axes", 0), gp);
    FLOOD_CUDA_CHECK(cudaMemsetAsync(cell_fill, 0, (size_t)L.max_cells * 4, st));
    cell_count_kernel<<<blocks, threads, 0, st>>>(pts, n, d, gp, cell_id, cell_fill);
    cell_scan_kernel<<<1, 1024, 0, st>>>(gp, cell_fill, cell_start);
    const int pd = record_floats(d);
    if (pd == 2) scatter_kernel<2><<<blocks, threads, 0, st>>>(pts, n, d, cell_id, cell_start, cell_fill, perm, out);
    else if (pd == 4) scatter_kernel<4><<<blocks, threads, 0, st>>>(pts, n, d, cell_id, cell_start, cell_fill, perm, out);
    else scatter_kernel<8><<<blocks, threads, 0, st>>>(pts, n, d, cell_id, cell_start, cell_fill, perm, out);
    count_launches(6);
    FLOOD_LAUNCH_CHECK("cloud_build kernels");
    return FLOOD_OK;
}

}  // namespace flood

// Bounding balls of simplices and per-face maxima: the two small kernels on either side of the
// covering-radius kernel.
#include "common.cuh"

namespace flood {
namespace {

// Reference rule, flooder/core.py:156-172.  One thread per simplex (S is at most a few
// million; the kernel is bandwidth-trivial).  Arithmetic is float32 with explicit rounding and
// NO fused multiply-add: squared differences are rounded, then added in coordinate order, then
// the square root is taken -- the same operation sequence as the oracle
// (oracle/flood_oracle.py::bounding_balls), so that near-ties between edges resolve identically.
// The centre is the midpoint of the first maximum of the flattened K x K distance matrix in
// row-major (i0, i1) order (core.py:157-161); the radius is scaled and padded with two
// separately rounded operations like the torch expression `amax * factor + 1e-3`.
__global__ void bounding_balls_kernel(const float *__restrict__ verts, int64_t S, int K, int d,
                                      float *__restrict__ centers, float *__restrict__ radii) {
    int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= S) return;
    const float *v = verts + s * K * d;
    float best = -1.f;
    int b0 = 0, b1 = 0;
    for (int i = 0; i < K; ++i)
        for (int j = 0; j < K; ++j) {
            float acc = 0.f;
            for (int a = 0; a < d; ++a) {
                const float t = __fsub_rn(v[i * d + a], v[j * d + a]);
                acc = __fadd_rn(acc, __fmul_rn(t, t));
            }
            const float dist = __fsqrt_rn(acc);
            if (dist > best) { best = dist; b0 = i; b1 = j; }
        }
    float c[FLOOD_MAX_DIM];
    for (int a = 0; a < d; ++a) {
        c[a] = __fmul_rn(__fadd_rn(v[b0 * d + a], v[b1 * d + a]), 0.5f);
        centers[s * d + a] = c[a];
    }
    float far = 0.f;
    for (int k = 0; k < K; ++k) {
        float acc = 0.f;
        for (int a = 0; a < d; ++a) {
            const float t = __fsub_rn(v[k * d + a], c[a]);
            acc = __fadd_rn(acc, __fmul_rn(t, t));
        }
        far = fmaxf(far, __fsqrt_rn(acc));
    }
    const float factor = (K - 1) > 1 ? 1.42f : 1.01f;
    radii[s] = __fadd_rn(__fmul_rn(far, factor), 1e-3f);
}

// Per-face maxima (flooder/core.py:251-257 in grid mode, :270 in random mode).
// One CTA per simplex.  Grid mode: samples are binned by the support of their barycentric
// weights (bit k <=> weight k non-zero); the value of face m is the maximum over all bins
// whose support is contained in m -- exactly the reference's `distances[:, face_idx].amax`,
// because face_idx lists the samples whose weights vanish outside the face.
__global__ void face_max_kernel(const float *__restrict__ min_dist2, int64_t R,
                                const int32_t *__restrict__ support, int K, float *__restrict__ out) {
    extern __shared__ unsigned bins[];  // 2^K entries (index 0 unused) or 1 entry
    const int64_t s = blockIdx.x;
    const int nb = support ? (1 << K) : 1;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) bins[i] = 0u;
    __syncthreads();
    const float *row = min_dist2 + s * R;
    if (support) {
        for (int64_t r = threadIdx.x; r < R; r += blockDim.x) {
            // non-negative floats order like their bit patterns (+inf included)
            atomicMax(&bins[support[r] & (nb - 1)], __float_as_uint(row[r]));
        }
    } else {
        unsigned m = 0u;
        for (int64_t r = threadIdx.x; r < R; r += blockDim.x) m = max(m, __float_as_uint(row[r]));
        m = __reduce_max_sync(0xffffffffu, m);
        if ((threadIdx.x & 31) == 0) atomicMax(&bins[0], m);
    }
    __syncthreads();
    if (support) {
        for (int m = 1 + threadIdx.x; m < nb; m += blockDim.x) {
            unsigned best = 0u;
            // enumerate the non-empty subsets of m
            for (int sub = m; sub; sub = (sub - 1) & m) best = max(best, bins[sub]);
            out[s * (nb - 1) + (m - 1)] = sqrtf(__uint_as_float(best));
        }
    } else if (threadIdx.x == 0) {
        out[s] = sqrtf(__uint_as_float(bins[0]));
    }
}

}  // namespace

int bounding_balls(const float *verts, int64_t S, int K, int d, float *centers, float *radii,
                   cudaStream_t st) {
    if (S == 0) return FLOOD_OK;
    if (!verts || !centers || !radii || S < 0 || K < 1 || K > FLOOD_MAX_SIMPLEX_VERTS || d < 1 ||
        d > FLOOD_MAX_DIM)
        return set_error(FLOOD_E_INVALID, "bounding_balls: bad arguments (S=%lld, K=%d, d=%d)",
                         (long long)S, K, d);
    const int threads = 128;
    bounding_balls_kernel<<<(unsigned)((S + threads - 1) / threads), threads, 0, st>>>(verts, S, K, d,
                                                                                      centers, radii);
    count_launches(1);
    FLOOD_LAUNCH_CHECK("bounding_balls_kernel");
    return FLOOD_OK;
}

int face_max(const float *min_dist2, int64_t S, int64_t R, const int32_t *support, int K, float *out,
             cudaStream_t st) {
    if (S == 0) return FLOOD_OK;
    if (!min_dist2 || !out || S < 0 || R < 1 || K < 1 || K > FLOOD_MAX_SIMPLEX_VERTS)
        return set_error(FLOOD_E_INVALID, "face_max: bad arguments (S=%lld, R=%lld, K=%d)",
                         (long long)S, (long long)R, K);
    if (S > 2147483647LL) return set_error(FLOOD_E_UNSUPPORTED, "face_max: S too large");
    const size_t smem = sizeof(unsigned) * (support ? (size_t(1) << K) : 1);
    face_max_kernel<<<(unsigned)S, 256, smem, st>>>(min_dist2, R, support, K, out);
    count_launches(1);
    FLOOD_LAUNCH_CHECK("face_max_kernel");
    return FLOOD_OK;
}

}  // namespace flood

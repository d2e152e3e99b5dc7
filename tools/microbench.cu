// Instruction-mix ceilings for the covering-radius inner loop on sm_100a.
// Each variant runs the same per-pair arithmetic from registers/shared memory only, so the
// measured pairs/s is the practical ceiling for that instruction selection.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define T 8
#define TILE 2048

__device__ __forceinline__ float min3(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

template <int VARIANT>
__global__ void __launch_bounds__(640, 1) mix_kernel(const float4 *__restrict__ cand, float *out, int iters) {
    __shared__ float4 tile[TILE];
    for (int i = threadIdx.x; i < TILE; i += blockDim.x) tile[i] = cand[i];
    __syncthreads();
    float x[T], y[T], z[T], m[T];
#pragma unroll
    for (int t = 0; t < T; ++t) {
        x[t] = threadIdx.x * 0.001f + t;
        y[t] = threadIdx.x * 0.002f - t;
        z[t] = blockIdx.x * 0.003f + t;
        m[t] = 1e30f;
    }
    if (VARIANT >= 5) {
        // software-pipelined: the next trip's candidates are fetched before the current math
        for (int it = 0; it < iters; ++it) {
            float4 q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) q[u] = tile[u];
#pragma unroll 1
            for (int j = 0; j < TILE; j += 4) {
                float4 p[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) p[u] = q[u];
                const int jn = (j + 4) & (TILE - 1);
#pragma unroll
                for (int u = 0; u < 4; ++u) q[u] = tile[jn + u];
                if (VARIANT == 5) {
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int t = 0; t < T; ++t) {
                            float dx = x[t] - p[u].x, dy = y[t] - p[u].y, dz = z[t] - p[u].z;
                            m[t] = fminf(m[t], fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
                        }
                } else {
#pragma unroll
                    for (int u = 0; u < 4; u += 2)
#pragma unroll
                        for (int t = 0; t < T; t += 2) {
                            float2 s[2];
#pragma unroll
                            for (int v = 0; v < 2; ++v) {
                                const float4 c = p[u + v];
                                float2 dx = __fadd2_rn(make_float2(x[t], x[t + 1]), make_float2(-c.x, -c.x));
                                float2 dy = __fadd2_rn(make_float2(y[t], y[t + 1]), make_float2(-c.y, -c.y));
                                float2 dz = __fadd2_rn(make_float2(z[t], z[t + 1]), make_float2(-c.z, -c.z));
                                float2 a = __fmul2_rn(dx, dx);
                                a = __ffma2_rn(dy, dy, a);
                                s[v] = __ffma2_rn(dz, dz, a);
                            }
                            m[t] = min3(m[t], s[0].x, s[1].x);
                            m[t + 1] = min3(m[t + 1], s[0].y, s[1].y);
                        }
                }
            }
        }
    } else
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int j = 0; j < TILE; j += 4) {
            float4 p[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) p[u] = tile[j + u];
            if (VARIANT == 0) {  // scalar: 3 FADD, FMUL, 2 FFMA, FMNMX
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        float dx = x[t] - p[u].x, dy = y[t] - p[u].y, dz = z[t] - p[u].z;
                        m[t] = fminf(m[t], fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
                    }
            } else if (VARIANT == 1) {  // scalar + 3-input min
#pragma unroll
                for (int u = 0; u < 4; u += 2)
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        float dx = x[t] - p[u].x, dy = y[t] - p[u].y, dz = z[t] - p[u].z;
                        float s0 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                        dx = x[t] - p[u + 1].x; dy = y[t] - p[u + 1].y; dz = z[t] - p[u + 1].z;
                        float s1 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                        m[t] = min3(m[t], s0, s1);
                    }
            } else if (VARIANT == 2) {  // packed f32x2 over sample pairs + 3-input min
#pragma unroll
                for (int u = 0; u < 4; u += 2)
#pragma unroll
                    for (int t = 0; t < T; t += 2) {
                        float2 s[2];
#pragma unroll
                        for (int v = 0; v < 2; ++v) {
                            const float4 q = p[u + v];
                            float2 dx = __fadd2_rn(make_float2(x[t], x[t + 1]), make_float2(-q.x, -q.x));
                            float2 dy = __fadd2_rn(make_float2(y[t], y[t + 1]), make_float2(-q.y, -q.y));
                            float2 dz = __fadd2_rn(make_float2(z[t], z[t + 1]), make_float2(-q.z, -q.z));
                            float2 a = __fmul2_rn(dx, dx);
                            a = __ffma2_rn(dy, dy, a);
                            s[v] = __ffma2_rn(dz, dz, a);
                        }
                        m[t] = min3(m[t], s[0].x, s[1].x);
                        m[t + 1] = min3(m[t + 1], s[0].y, s[1].y);
                    }
            } else if (VARIANT == 3) {  // norm expansion: 3 FFMA + FMNMX (for reference only)
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int t = 0; t < T; ++t)
                        m[t] = fminf(m[t], fmaf(x[t], p[u].x, fmaf(y[t], p[u].y, fmaf(z[t], p[u].z, p[u].w))));
            } else if (VARIANT == 4) {  // pure FFMA issue-rate probe (7 per pair)
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int t = 0; t < T; ++t) {
                        float a = fmaf(x[t], p[u].x, m[t]);
                        a = fmaf(y[t], p[u].y, a); a = fmaf(z[t], p[u].z, a); a = fmaf(x[t], p[u].w, a);
                        a = fmaf(y[t], p[u].x, a); a = fmaf(z[t], p[u].y, a);
                        m[t] = fmaf(a, p[u].z, a);
                    }
            }
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < T; ++t) acc += m[t];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int V>
void run(const char *name, const float4 *cand, float *out, int sms, double slots_per_pair, int threads = 640) {
    const int iters = 40;
    mix_kernel<V><<<sms, threads>>>(cand, out, 2);
    cudaDeviceSynchronize();
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a);
        mix_kernel<V><<<sms, threads>>>(cand, out, iters);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    double pairs = (double)sms * threads * T * (double)TILE * iters;
    double rate = pairs / (best * 1e-3);
    printf("[%4d thr] %-28s %8.3f ms  %.3e pairs/s  (= %.1f%% of the 7-slot roofline at 1.965 GHz; %.2f slots/pair nominal)\n",
           threads, name, best, rate, 100.0 * rate / (sms * 128.0 * 1.965e9 / 7.0), slots_per_pair);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float4 *cand; float *out;
    cudaMalloc(&cand, TILE * sizeof(float4));
    cudaMemset(cand, 0, TILE * sizeof(float4));
    cudaMalloc(&out, (size_t)sms * 640 * sizeof(float));
    printf("SMs: %d\n", sms);
    run<0>("scalar (7 slots)", cand, out, sms, 7);
    run<1>("scalar + min3 (6.5)", cand, out, sms, 6.5);
    run<2>("packed f32x2 + min3 (3.5)", cand, out, sms, 3.5);
    run<3>("norm expansion (4)", cand, out, sms, 4);
    run<4>("7 x FFMA probe", cand, out, sms, 7);
    run<5>("scalar, sw-pipelined", cand, out, sms, 7);
    run<6>("packed+min3, sw-pipelined", cand, out, sms, 3.5);
    const int sweep[] = {128, 256, 384, 512};
    for (int th : sweep) {
        run<0>("scalar (7 slots)", cand, out, sms, 7, th);
        run<5>("scalar, sw-pipelined", cand, out, sms, 7, th);
        run<2>("packed f32x2 + min3", cand, out, sms, 3.5, th);
        run<6>("packed+min3, sw-pipelined", cand, out, sms, 3.5, th);
        run<4>("7 x FFMA probe", cand, out, sms, 7, th);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}

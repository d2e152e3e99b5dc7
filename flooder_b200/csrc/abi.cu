// extern "C" surface of libflood_b200.so (declared in include/flood_b200.h).
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>

#include "common.cuh"

namespace flood {

static thread_local char g_error[512] = "";

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

static std::mutex g_opt_mutex;
static std::map<std::string, int> g_options;

int get_option(const char *name, int fallback) {
    std::lock_guard<std::mutex> lock(g_opt_mutex);
    auto it = g_options.find(name);
    return (it == g_options.end() || it->second < 0) ? fallback : it->second;   // negative = unset
}

struct KernelTimer {
    cudaEvent_t a = nullptr, b = nullptr;
    double total_ms = 0.0;
    long launches = 0;
    bool pending = false;
};
static std::map<std::string, KernelTimer> g_timers;

static void timer_collect(KernelTimer &t) {
    if (!t.pending) return;
    float ms = 0.f;
    if (cudaEventSynchronize(t.b) == cudaSuccess && cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) {
        t.total_ms += ms;
        t.launches += 1;
    }
    t.pending = false;
}

void kernel_timer_start(const char *name, cudaStream_t st) {
    std::lock_guard<std::mutex> lock(g_opt_mutex);
    KernelTimer &t = g_timers[name];
    if (!t.a) { cudaEventCreate(&t.a); cudaEventCreate(&t.b); }
    timer_collect(t);  // one launch in flight per timer: fold the previous one in first
    cudaEventRecord(t.a, st);
}

void kernel_timer_stop(const char *name, cudaStream_t st) {
    std::lock_guard<std::mutex> lock(g_opt_mutex);
    KernelTimer &t = g_timers[name];
    cudaEventRecord(t.b, st);
    t.pending = true;
}

static std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int device_sm_count() {
    static thread_local int cached_dev = -1, cached_sms = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1)
            sms = 148;
        cached_dev = dev;
        cached_sms = sms;
    }
    return cached_sms;
}

}  // namespace flood

using namespace flood;

extern "C" {

int flood_abi_version(void) { return FLOOD_ABI_VERSION; }

const char *flood_last_error(void) { return g_error; }

int flood_device_info(int *sm_count, int *sm_clock_khz) {
    int dev = 0;
    FLOOD_CUDA_CHECK(cudaGetDevice(&dev));
    if (sm_count) FLOOD_CUDA_CHECK(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
    if (sm_clock_khz) FLOOD_CUDA_CHECK(cudaDeviceGetAttribute(sm_clock_khz, cudaDevAttrClockRate, dev));
    return FLOOD_OK;
}

int flood_bounding_balls_f64(const double *verts, int64_t S, int K, int d, double *centers, double *radii,
                             void *stream) {
    return bounding_balls_f64(verts, S, K, d, centers, radii, (cudaStream_t)stream);
}

size_t flood_covering_workspace_bytes_f64(int64_t S, int d) { return covering_workspace_bytes_f64(S, d); }

int flood_covering_radius_f64(const void *cloud_workspace, const double *pts, int64_t n, int d, const double *verts,
                              int64_t S, int K, const double *weights, int64_t R, const double *centers,
                              const double *radii, double *out_min_dist2, int64_t *out_cand_count,
                              unsigned long long *out_evals, void *workspace, size_t workspace_bytes, void *stream) {
    return covering_radius_f64(cloud_workspace, pts, n, d, verts, S, K, weights, R, centers, radii, out_min_dist2,
                               out_cand_count, out_evals, workspace, workspace_bytes, (cudaStream_t)stream);
}

int flood_face_max_f64(const double *min_dist2, int64_t S, int64_t R, const int32_t *support, int K, double *out,
                       void *stream) {
    return face_max_f64(min_dist2, S, R, support, K, out, (cudaStream_t)stream);
}

int flood_set_option(const char *name, int value) {
    if (!name) return 0;
    std::lock_guard<std::mutex> lock(g_opt_mutex);
    int prev = -1;   // -1: the option was unset (library default in effect)
    auto it = g_options.find(name);
    if (it != g_options.end()) prev = it->second;
    g_options[name] = value;
    return prev;
}

int flood_kernel_ms(const char *name, double *total_ms, long *launches, int reset) {
    if (!name) return FLOOD_E_INVALID;
    std::lock_guard<std::mutex> lock(g_opt_mutex);
    auto it = g_timers.find(name);
    if (it == g_timers.end()) {
        if (total_ms) *total_ms = 0.0;
        if (launches) *launches = 0;
        return FLOOD_OK;
    }
    timer_collect(it->second);
    if (total_ms) *total_ms = it->second.total_ms;
    if (launches) *launches = it->second.launches;
    if (reset) { it->second.total_ms = 0.0; it->second.launches = 0; }
    return FLOOD_OK;
}

long long flood_launch_count(int reset) {
    return reset ? g_launches.exchange(0) : g_launches.load();
}

size_t flood_fps_workspace_bytes(int64_t n, int d, int64_t n_lms) { return fps_workspace_bytes(n, d, n_lms); }

int flood_fps_f32(const float *pts, int64_t n, int d, int64_t n_lms, int64_t start_idx, int64_t *out_idx,
                  void *workspace, size_t workspace_bytes, void *stream) {
    return fps(pts, n, d, n_lms, start_idx, out_idx, workspace, workspace_bytes, (cudaStream_t)stream);
}

int flood_fps_grid_f32(const void *cloud_workspace, const float *pts, int64_t n, int d, int64_t n_lms,
                       int64_t start_idx, int64_t *out_idx, void *workspace, size_t workspace_bytes,
                       void *stream) {
    return fps_grid(cloud_workspace, pts, n, d, n_lms, start_idx, out_idx, workspace, workspace_bytes,
                    (cudaStream_t)stream);
}

size_t flood_cloud_workspace_bytes(int64_t n, int d) {
    if (n < 1 || d < 1 || d > FLOOD_MAX_DIM) return 0;
    return (size_t)cloud_layout(n, d).total;
}

int flood_cloud_build_f32(const float *pts, int64_t n, int d, int points_per_cell, void *workspace,
                          size_t workspace_bytes, void *stream) {
    return cloud_build(pts, n, d, points_per_cell, workspace, workspace_bytes, (cudaStream_t)stream);
}

int flood_bounding_balls_f32(const float *verts, int64_t S, int K, int d, float *centers, float *radii,
                             void *stream) {
    return bounding_balls(verts, S, K, d, centers, radii, (cudaStream_t)stream);
}

size_t flood_covering_workspace_bytes(int64_t S, int64_t R, int d) { return covering_workspace_bytes(S, R, d); }

int flood_covering_radius_f32(const void *cloud_workspace, int64_t n, int d, const float *verts, int64_t S,
                              int K, const float *weights, int64_t R, const float *samples,
                              const float *centers, const float *radii, float *out_min_dist2,
                              int64_t *out_cand_count, unsigned long long *out_evals, void *workspace,
                              size_t workspace_bytes, void *stream) {
    return covering_radius(cloud_workspace, n, d, verts, S, K, weights, R, samples, centers, radii,
                           out_min_dist2, out_cand_count, out_evals, workspace, workspace_bytes,
                           (cudaStream_t)stream);
}

int flood_covering_bricks(int64_t R, int d, int32_t *out_groups, int capacity, int *bricks_per_block) {
    return covering_bricks(R, d, out_groups, capacity, bricks_per_block);
}

int flood_covering_plan_f32(const void *cloud_workspace, int64_t n, int d, const float *centers,
                            const float *radii, int64_t S, int32_t *out_tested, void *stream) {
    return covering_plan(cloud_workspace, n, d, centers, radii, S, out_tested, (cudaStream_t)stream);
}

int flood_face_max_f32(const float *min_dist2, int64_t S, int64_t R, const int32_t *support, int K,
                       float *out, void *stream) {
    return face_max(min_dist2, S, R, support, K, out, (cudaStream_t)stream);
}

}  // extern "C"

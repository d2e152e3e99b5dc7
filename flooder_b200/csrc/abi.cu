// extern "C" surface of libflood_b200.so (declared in include/flood_b200.h).
#include <stdarg.h>
#include <string.h>

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
    return it == g_options.end() ? fallback : it->second;
}

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

int flood_set_option(const char *name, int value) {
    if (!name) return 0;
    std::lock_guard<std::mutex> lock(g_opt_mutex);
    int prev = 0;
    auto it = g_options.find(name);
    if (it != g_options.end()) prev = it->second;
    g_options[name] = value;
    return prev;
}

size_t flood_fps_workspace_bytes(int64_t n, int d, int64_t n_lms) { return fps_workspace_bytes(n, d, n_lms); }

int flood_fps_f32(const float *pts, int64_t n, int d, int64_t n_lms, int64_t start_idx, int64_t *out_idx,
                  void *workspace, size_t workspace_bytes, void *stream) {
    return fps(pts, n, d, n_lms, start_idx, out_idx, workspace, workspace_bytes, (cudaStream_t)stream);
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

int flood_face_max_f32(const float *min_dist2, int64_t S, int64_t R, const int32_t *support, int K,
                       float *out, void *stream) {
    return face_max(min_dist2, S, R, support, K, out, (cudaStream_t)stream);
}

}  // extern "C"

// PyTorch C++ extension: the thin loader between torch tensors and the C ABI of
// libflood_b200.so (include/flood_b200.h).  It only validates tensors, allocates outputs and
// workspaces with torch's caching allocator, passes raw device pointers plus torch's current
// CUDA stream across the C boundary, and turns error codes into Python exceptions.
#include <torch/extension.h>

#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include "flood_b200.h"

namespace {

void check(int rc, const char *what) {
    TORCH_CHECK(rc == FLOOD_OK, what, " failed (", rc, "): ", flood_last_error());
}

void *current_stream(const torch::Tensor &t) {
    return (void *)c10::cuda::getCurrentCUDAStream(t.device().index()).stream();
}

void need_cuda_f32(const torch::Tensor &t, const char *name) {
    TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor (flooder_b200 has no CPU path)");
    TORCH_CHECK(t.scalar_type() == torch::kFloat32, name, " must be float32");
    TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
}

torch::Tensor bytes_like(const torch::Tensor &ref, size_t n) {
    return torch::empty({(int64_t)n}, ref.options().dtype(torch::kUInt8));
}

torch::Tensor fps(const torch::Tensor &pts, int64_t n_lms, int64_t start_idx) {
    need_cuda_f32(pts, "points");
    TORCH_CHECK(pts.dim() == 2, "points must be (N, D)");
    const c10::cuda::CUDAGuard guard(pts.device());
    const int64_t n = pts.size(0);
    const int d = (int)pts.size(1);
    auto out = torch::empty({n_lms}, pts.options().dtype(torch::kInt64));
    const size_t wsb = flood_fps_workspace_bytes(n, d, n_lms);
    auto ws = bytes_like(pts, wsb);
    check(flood_fps_f32(pts.data_ptr<float>(), n, d, n_lms, start_idx, out.data_ptr<int64_t>(),
                        ws.data_ptr(), wsb, current_stream(pts)),
          "flood_fps_f32");
    return out;
}

torch::Tensor fps_grid(const torch::Tensor &cloud_ws, const torch::Tensor &pts, int64_t n_lms, int64_t start_idx) {
    need_cuda_f32(pts, "points");
    TORCH_CHECK(pts.dim() == 2, "points must be (N, D)");
    TORCH_CHECK(cloud_ws.is_cuda() && cloud_ws.scalar_type() == torch::kUInt8, "cloud workspace must be CUDA uint8");
    const c10::cuda::CUDAGuard guard(pts.device());
    const int64_t n = pts.size(0);
    const int d = (int)pts.size(1);
    auto out = torch::empty({n_lms}, pts.options().dtype(torch::kInt64));
    const size_t wsb = flood_fps_workspace_bytes(n, d, n_lms);
    auto ws = bytes_like(pts, wsb);
    check(flood_fps_grid_f32(cloud_ws.data_ptr(), pts.data_ptr<float>(), n, d, n_lms, start_idx,
                             out.data_ptr<int64_t>(), ws.data_ptr(), wsb, current_stream(pts)),
          "flood_fps_grid_f32");
    return out;
}

torch::Tensor cloud_build(const torch::Tensor &pts, int64_t points_per_cell) {
    need_cuda_f32(pts, "points");
    TORCH_CHECK(pts.dim() == 2, "points must be (N, D)");
    const c10::cuda::CUDAGuard guard(pts.device());
    const int64_t n = pts.size(0);
    const int d = (int)pts.size(1);
    const size_t wsb = flood_cloud_workspace_bytes(n, d);
    TORCH_CHECK(wsb > 0, "unsupported cloud shape (", n, ", ", d, ")");
    auto ws = bytes_like(pts, wsb);
    check(flood_cloud_build_f32(pts.data_ptr<float>(), n, d, (int)points_per_cell, ws.data_ptr(), wsb,
                                current_stream(pts)),
          "flood_cloud_build_f32");
    return ws;
}

std::tuple<torch::Tensor, torch::Tensor> bounding_balls(const torch::Tensor &verts) {
    need_cuda_f32(verts, "simplex_vertices");
    TORCH_CHECK(verts.dim() == 3, "simplex_vertices must be (S, K, D)");
    const c10::cuda::CUDAGuard guard(verts.device());
    const int64_t S = verts.size(0);
    auto centers = torch::empty({S, verts.size(2)}, verts.options());
    auto radii = torch::empty({S}, verts.options());
    check(flood_bounding_balls_f32(verts.data_ptr<float>(), S, (int)verts.size(1), (int)verts.size(2),
                                   centers.data_ptr<float>(), radii.data_ptr<float>(), current_stream(verts)),
          "flood_bounding_balls_f32");
    return {centers, radii};
}

// returns (min_dist2 [S,R], cand_count [S] int64, evals [1] int64, executed [1] int64)
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor> covering_radius(
    const torch::Tensor &cloud_ws, int64_t n, int64_t d, const torch::Tensor &verts,
    const torch::Tensor &weights, const c10::optional<torch::Tensor> &samples,
    const torch::Tensor &centers, const torch::Tensor &radii) {
    need_cuda_f32(verts, "simplex_vertices");
    need_cuda_f32(weights, "weights");
    need_cuda_f32(centers, "centers");
    need_cuda_f32(radii, "radii");
    TORCH_CHECK(cloud_ws.is_cuda() && cloud_ws.scalar_type() == torch::kUInt8, "cloud workspace must be CUDA uint8");
    TORCH_CHECK(verts.dim() == 3 && verts.size(2) == d, "simplex_vertices must be (S, K, D)");
    TORCH_CHECK(weights.dim() == 2 && weights.size(1) == verts.size(1), "weights must be (R, K)");
    const int64_t S = verts.size(0), K = verts.size(1), R = weights.size(0);
    TORCH_CHECK(centers.dim() == 2 && centers.size(0) == S && centers.size(1) == d, "centers must be (S, D)");
    TORCH_CHECK(radii.dim() == 1 && radii.size(0) == S, "radii must be (S,)");
    const float *samples_ptr = nullptr;
    if (samples.has_value()) {
        need_cuda_f32(*samples, "samples");
        TORCH_CHECK(samples->dim() == 3 && samples->size(0) == S && samples->size(1) == R && samples->size(2) == d,
                    "samples must be (S, R, D)");
        samples_ptr = samples->data_ptr<float>();
    }
    const c10::cuda::CUDAGuard guard(verts.device());
    auto out = torch::empty({S, R}, verts.options());
    auto counts = torch::empty({S}, verts.options().dtype(torch::kInt64));
    auto evals = torch::zeros({1}, verts.options().dtype(torch::kInt64));
    const size_t wsb = flood_covering_workspace_bytes(S, R, (int)d);
    auto ws = bytes_like(verts, wsb);
    check(flood_covering_radius_f32(cloud_ws.data_ptr(), n, (int)d, verts.data_ptr<float>(), S, (int)K,
                                    weights.data_ptr<float>(), R, samples_ptr, centers.data_ptr<float>(),
                                    radii.data_ptr<float>(), out.data_ptr<float>(), counts.data_ptr<int64_t>(),
                                    reinterpret_cast<unsigned long long *>(evals.data_ptr<int64_t>()),
                                    ws.data_ptr(), wsb, current_stream(verts)),
          "flood_covering_radius_f32");
    // evaluations actually executed: uint64 at byte FLOOD_COVER_WS_EXECUTED_OFFSET of the workspace
    auto executed = ws.slice(0, FLOOD_COVER_WS_EXECUTED_OFFSET, FLOOD_COVER_WS_EXECUTED_OFFSET + 8)
                        .view(torch::kInt64)
                        .clone();
    return {out, counts, evals, executed};
}

// (groups per brick, bricks per CTA) -- host-side layout query
std::tuple<std::vector<int64_t>, int64_t> covering_bricks(int64_t R, int64_t d) {
    int per_block = 0;
    const int n = flood_covering_bricks(R, (int)d, nullptr, 0, &per_block);
    check(n < 0 ? n : FLOOD_OK, "flood_covering_bricks");
    std::vector<int32_t> g((size_t)n);
    check(flood_covering_bricks(R, (int)d, g.data(), n, &per_block) < 0 ? FLOOD_E_INVALID : FLOOD_OK,
          "flood_covering_bricks");
    return {std::vector<int64_t>(g.begin(), g.end()), (int64_t)per_block};
}

torch::Tensor covering_plan(const torch::Tensor &cloud_ws, int64_t n, int64_t d, const torch::Tensor &centers,
                            const torch::Tensor &radii) {
    need_cuda_f32(centers, "centers");
    need_cuda_f32(radii, "radii");
    TORCH_CHECK(centers.dim() == 2 && centers.size(1) == d && radii.numel() == centers.size(0), "bad ball shapes");
    const c10::cuda::CUDAGuard guard(centers.device());
    auto out = torch::empty({centers.size(0)}, centers.options().dtype(torch::kInt32));
    check(flood_covering_plan_f32(cloud_ws.data_ptr(), n, (int)d, centers.data_ptr<float>(), radii.data_ptr<float>(),
                                  centers.size(0), out.data_ptr<int32_t>(), current_stream(centers)),
          "flood_covering_plan_f32");
    return out;
}

torch::Tensor face_max(const torch::Tensor &min_dist2, const c10::optional<torch::Tensor> &support, int64_t K) {
    need_cuda_f32(min_dist2, "min_dist2");
    TORCH_CHECK(min_dist2.dim() == 2, "min_dist2 must be (S, R)");
    const int64_t S = min_dist2.size(0), R = min_dist2.size(1);
    const int32_t *sup = nullptr;
    int64_t width = 1;
    if (support.has_value()) {
        TORCH_CHECK(support->is_cuda() && support->scalar_type() == torch::kInt32 && support->is_contiguous() &&
                        support->numel() == R,
                    "support must be a contiguous CUDA int32 tensor of R masks");
        sup = support->data_ptr<int32_t>();
        width = (int64_t(1) << K) - 1;
    }
    const c10::cuda::CUDAGuard guard(min_dist2.device());
    auto out = torch::empty({S, width}, min_dist2.options());
    check(flood_face_max_f32(min_dist2.data_ptr<float>(), S, R, sup, (int)K, out.data_ptr<float>(),
                             current_stream(min_dist2)),
          "flood_face_max_f32");
    return out;
}

// ---- float64 path -------------------------------------------------------------------------
void need_cuda_f64(const torch::Tensor &t, const char *name) {
    TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor (flooder_b200 has no CPU path)");
    TORCH_CHECK(t.scalar_type() == torch::kFloat64, name, " must be float64");
    TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
}

std::tuple<torch::Tensor, torch::Tensor> bounding_balls_f64(const torch::Tensor &verts) {
    need_cuda_f64(verts, "simplex_vertices");
    TORCH_CHECK(verts.dim() == 3, "simplex_vertices must be (S, K, D)");
    const c10::cuda::CUDAGuard guard(verts.device());
    const int64_t S = verts.size(0);
    auto centers = torch::empty({S, verts.size(2)}, verts.options());
    auto radii = torch::empty({S}, verts.options());
    check(flood_bounding_balls_f64(verts.data_ptr<double>(), S, (int)verts.size(1), (int)verts.size(2),
                                   centers.data_ptr<double>(), radii.data_ptr<double>(), current_stream(verts)),
          "flood_bounding_balls_f64");
    return {centers, radii};
}

// returns (min_dist2 [S,R] float64, cand_count [S] int64, evals [1] int64)
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor> covering_radius_f64(
    const torch::Tensor &cloud_ws, const torch::Tensor &pts, const torch::Tensor &verts, const torch::Tensor &weights,
    const torch::Tensor &centers, const torch::Tensor &radii) {
    need_cuda_f64(pts, "points");
    need_cuda_f64(verts, "simplex_vertices");
    need_cuda_f64(weights, "weights");
    need_cuda_f64(centers, "centers");
    need_cuda_f64(radii, "radii");
    TORCH_CHECK(cloud_ws.is_cuda() && cloud_ws.scalar_type() == torch::kUInt8, "cloud workspace must be CUDA uint8");
    TORCH_CHECK(pts.dim() == 2 && verts.dim() == 3 && verts.size(2) == pts.size(1), "bad point / vertex shapes");
    TORCH_CHECK(weights.dim() == 2 && weights.size(1) == verts.size(1), "weights must be (R, K)");
    const int64_t n = pts.size(0), d = pts.size(1), S = verts.size(0), K = verts.size(1), R = weights.size(0);
    TORCH_CHECK(centers.dim() == 2 && centers.size(0) == S && centers.size(1) == d && radii.numel() == S, "bad ball shapes");
    const c10::cuda::CUDAGuard guard(verts.device());
    auto out = torch::empty({S, R}, verts.options());
    auto counts = torch::empty({S}, verts.options().dtype(torch::kInt64));
    auto evals = torch::zeros({1}, verts.options().dtype(torch::kInt64));
    const size_t wsb = flood_covering_workspace_bytes_f64(S, (int)d);
    auto ws = bytes_like(verts, wsb);
    check(flood_covering_radius_f64(cloud_ws.data_ptr(), pts.data_ptr<double>(), n, (int)d, verts.data_ptr<double>(), S,
                                    (int)K, weights.data_ptr<double>(), R, centers.data_ptr<double>(),
                                    radii.data_ptr<double>(), out.data_ptr<double>(), counts.data_ptr<int64_t>(),
                                    reinterpret_cast<unsigned long long *>(evals.data_ptr<int64_t>()), ws.data_ptr(), wsb,
                                    current_stream(verts)),
          "flood_covering_radius_f64");
    return {out, counts, evals};
}

torch::Tensor face_max_f64(const torch::Tensor &min_dist2, const c10::optional<torch::Tensor> &support, int64_t K) {
    need_cuda_f64(min_dist2, "min_dist2");
    TORCH_CHECK(min_dist2.dim() == 2, "min_dist2 must be (S, R)");
    const int64_t S = min_dist2.size(0), R = min_dist2.size(1);
    const int32_t *sup = nullptr;
    int64_t width = 1;
    if (support.has_value()) {
        TORCH_CHECK(support->is_cuda() && support->scalar_type() == torch::kInt32 && support->is_contiguous() &&
                        support->numel() == R,
                    "support must be a contiguous CUDA int32 tensor of R masks");
        sup = support->data_ptr<int32_t>();
        width = (int64_t(1) << K) - 1;
    }
    const c10::cuda::CUDAGuard guard(min_dist2.device());
    auto out = torch::empty({S, width}, min_dist2.options());
    check(flood_face_max_f64(min_dist2.data_ptr<double>(), S, R, sup, (int)K, out.data_ptr<double>(),
                             current_stream(min_dist2)),
          "flood_face_max_f64");
    return out;
}

std::tuple<double, int64_t> kernel_ms(const std::string &name, bool reset) {
    double ms = 0.0;
    long launches = 0;
    check(flood_kernel_ms(name.c_str(), &ms, &launches, reset ? 1 : 0), "flood_kernel_ms");
    return {ms, (int64_t)launches};
}

int64_t set_option(const std::string &name, int64_t value) { return flood_set_option(name.c_str(), (int)value); }

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "torch loader for the C ABI of libflood_b200.so";
    m.def("abi_version", &flood_abi_version);
    m.def("fps", &fps);
    m.def("fps_grid", &fps_grid);
    m.def("cloud_build", &cloud_build);
    m.def("bounding_balls", &bounding_balls);
    m.def("covering_radius", &covering_radius);
    m.def("face_max", &face_max);
    m.def("covering_plan", &covering_plan);
    m.def("covering_bricks", &covering_bricks);
    m.def("bounding_balls_f64", &bounding_balls_f64);
    m.def("covering_radius_f64", &covering_radius_f64);
    m.def("face_max_f64", &face_max_f64);
    m.def("set_option", &set_option);
    // error plumbing self-test: turns a library error code into the Python exception a failing call raises
    m.def("_raise_for_code", [](int64_t rc) { check((int)rc, "self-test"); });
    m.def("launch_count", [](bool reset) { return (int64_t)flood_launch_count(reset ? 1 : 0); });
    m.def("kernel_ms", &kernel_ms);
}

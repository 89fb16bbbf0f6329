// ORACLE tooling: pybind shim exposing two of the reference's CUDA entry points
// (declared in torkit3d/csrc/include/{sample_farthest_points,chamfer_distance}.h).
#include <torch/extension.h>
#include <vector>

at::Tensor sample_farthest_points_cuda(const at::Tensor points, const int64_t num_samples);
std::vector<at::Tensor> chamfer_distance_forward_cuda(const at::Tensor xyz1, const at::Tensor xyz2);

PYBIND11_MODULE(torkit3d_ref_ops, m) {
    m.def("sample_farthest_points_cuda", &sample_farthest_points_cuda);
    m.def("chamfer_distance_forward_cuda", &chamfer_distance_forward_cuda);
}

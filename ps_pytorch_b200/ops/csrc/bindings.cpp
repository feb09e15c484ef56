#include <torch/extension.h>

namespace psb {
// optim_kernels.cu
void sgd_flat(at::Tensor, at::Tensor, at::Tensor, double, double, double, double, bool, bool, double, c10::optional<at::Tensor>);
void adam_flat(at::Tensor, at::Tensor, at::Tensor, at::Tensor, c10::optional<at::Tensor>, double, double, double, double, double, double, c10::optional<at::Tensor>);
// comm_kernels.cu
void push(int64_t, bool, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, bool);
void reduce_update(std::vector<int64_t>, std::vector<int64_t>, int64_t, double, int64_t, int64_t, int64_t, std::vector<int64_t>, std::vector<int64_t>,
                   at::Tensor, c10::optional<at::Tensor>, c10::optional<at::Tensor>, c10::optional<at::Tensor>, int64_t, std::vector<int64_t>, int64_t,
                   std::vector<int64_t>, int64_t, int64_t, double, bool, int64_t, bool, std::vector<int64_t>, double);
void bcast(int64_t, std::vector<int64_t>, int64_t, int64_t, std::vector<int64_t>, int64_t, int64_t, int64_t, int64_t);
void wait_flag(int64_t, int64_t, int64_t, int64_t, double);
void ru_set_mc_src(int64_t);
void add_step(int64_t, int64_t);
void set_word(int64_t, int64_t);
at::Tensor make_hp(double, double, double, double, bool, double, double, double, bool, int64_t, int64_t);
int64_t hp_steps(at::Tensor);
// bn_kernels.cu
std::vector<at::Tensor> bn_forward(at::Tensor, c10::optional<at::Tensor>, c10::optional<at::Tensor>, c10::optional<at::Tensor>, c10::optional<at::Tensor>,
                                   c10::optional<at::Tensor>, bool, double, double, bool, c10::optional<at::Tensor>, c10::optional<at::Tensor>);
std::vector<at::Tensor> bn_backward(at::Tensor, c10::optional<at::Tensor>, at::Tensor, c10::optional<at::Tensor>, at::Tensor, at::Tensor, bool, bool, c10::optional<at::Tensor>,
                                    c10::optional<at::Tensor>, c10::optional<at::Tensor>);
void bn_set_finalize_wide(int);
// programmatic dependent launch of the step's kernels (pdl.cuh).  OFF by default: measured 6.87 vs 6.24 ms/step (profiles/run30_pdl.md) — grids that
// sit resident behind griddepcontrol.wait take the thread / shared-memory slots the side-stream weight gradients would have filled.  PSB200_PDL=1 enables.
static int g_pdl = 0;
int pdl_enabled() { return g_pdl; }
void pdl_set(int on) { g_pdl = on ? 1 : 0; }
// elementwise.cu
at::Tensor relu_forward(at::Tensor);
at::Tensor relu_backward(at::Tensor, at::Tensor);
at::Tensor maxpool2x2_forward(at::Tensor, bool);
at::Tensor maxpool2x2_backward(at::Tensor, at::Tensor, bool);
at::Tensor maxpool3x3s2_forward(at::Tensor);
at::Tensor maxpool3x3s2_backward(at::Tensor, at::Tensor, at::Tensor);
std::vector<at::Tensor> dropout_forward(at::Tensor, double, at::Tensor, int64_t);
at::Tensor dropout_backward(at::Tensor, double, at::Tensor);
void bump_word(at::Tensor);
at::Tensor gap_forward(at::Tensor);
at::Tensor gap_backward(at::Tensor, int64_t, int64_t);
std::vector<at::Tensor> softmax_ce(at::Tensor, at::Tensor);
// arena.cpp
std::vector<int64_t> vmm_alloc(int64_t, int64_t, int64_t);
std::vector<int64_t> vmm_import(int64_t, int64_t, int64_t);
void vmm_unmap(int64_t, int64_t, int64_t);
bool multicast_supported(int64_t);
int64_t mc_granularity(int64_t, int64_t);
std::vector<int64_t> mc_create(int64_t, int64_t);
int64_t mc_import(int64_t);
void mc_add_device(int64_t, int64_t);
void mc_bind(int64_t, int64_t, int64_t);
int64_t mc_map(int64_t, int64_t, int64_t, int64_t);
void mc_unmap(int64_t, int64_t, int64_t, int64_t);
void close_fd(int64_t);
at::Tensor tensor_from_ptr(int64_t, int64_t, int64_t);
}  // namespace psb

void bind_gemm(pybind11::module_& m);

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    bind_gemm(m);
    using namespace psb;
    m.def("sgd_flat", &sgd_flat); m.def("adam_flat", &adam_flat);
    m.def("push", &push); m.def("reduce_update", &reduce_update); m.def("bcast", &bcast); m.def("wait_flag", &wait_flag);
    m.def("add_step", &add_step); m.def("set_word", &set_word); m.def("make_hp", &make_hp); m.def("hp_steps", &hp_steps);
    m.def("ru_set_mc_src", &ru_set_mc_src); m.def("bn_forward", &bn_forward, pybind11::arg("x"), pybind11::arg("gamma"), pybind11::arg("beta"), pybind11::arg("running_mean"), pybind11::arg("running_var"),
          pybind11::arg("residual"), pybind11::arg("training"), pybind11::arg("momentum"), pybind11::arg("eps"), pybind11::arg("relu"), pybind11::arg("ws"),
          pybind11::arg("conv_partial") = c10::optional<at::Tensor>()); m.def("bn_backward", &bn_backward); m.def("bn_set_finalize_wide", &bn_set_finalize_wide); m.def("set_pdl", &pdl_set); m.def("pdl_enabled", &pdl_enabled);
    m.def("relu_forward", &relu_forward); m.def("relu_backward", &relu_backward);
    m.def("maxpool2x2_forward", &maxpool2x2_forward); m.def("maxpool2x2_backward", &maxpool2x2_backward);
    m.def("maxpool3x3s2_forward", &maxpool3x3s2_forward); m.def("maxpool3x3s2_backward", &maxpool3x3s2_backward);
    m.def("dropout_forward", &dropout_forward); m.def("dropout_backward", &dropout_backward); m.def("bump_word", &bump_word);
    m.def("gap_forward", &gap_forward); m.def("gap_backward", &gap_backward); m.def("softmax_ce", &softmax_ce);
    m.def("vmm_alloc", &vmm_alloc); m.def("vmm_import", &vmm_import); m.def("vmm_unmap", &vmm_unmap);
    m.def("multicast_supported", &multicast_supported); m.def("mc_granularity", &mc_granularity); m.def("mc_create", &mc_create);
    m.def("mc_import", &mc_import); m.def("mc_add_device", &mc_add_device); m.def("mc_bind", &mc_bind); m.def("mc_map", &mc_map); m.def("mc_unmap", &mc_unmap);
    m.def("close_fd", &close_fd); m.def("tensor_from_ptr", &tensor_from_ptr);
}

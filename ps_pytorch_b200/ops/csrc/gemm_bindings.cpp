// torch-facing wrappers of the tcgen05 GEMM / implicit-GEMM convolution launchers (gemm_tcgen05.cu is torch-free so it
// compiles in seconds).
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>

namespace psb {
const char* tc_gemm_tn(const void*, const void*, void*, const void*, int, int, int, int, cudaStream_t);
const char* tc_conv_fprop(const void*, const void*, void*, const void*, const void*, int, int, int, int, int, int, int, int, int, int, float*, cudaStream_t);
const char* tc_conv2(const void*, const void*, void*, const void*, const void*, int, int, int, int, int, int, int, int, int, int, int, int, int, float*, cudaStream_t);
int num_sms();
const char* weight_transpose(const void*, void*, int, int, int, int, cudaStream_t);
const char* weight_transpose_batched(const void*, int, int, const void*, int, void*, void*, double, cudaStream_t);
void tc_set_conv2_bn(int);
void tc_set_conv_algo(int);
void tc_set_halo_desc_mode(int);
void tc_set_halo_cluster(int);
void tc_set_halo_pair(int);
void tc_set_halo_pitch(int);
void tc_set_wgrad_sms(int);
void tc_set_halo_resident(int);
void tc_set_gemm_bn(int);
int tc_check_error();
const char* stem_fprop(const void*, const void*, void*, const void*, int, int, int, int, int, int, float*, cudaStream_t);
const char* stem_wgrad_tc(const void*, const void*, void*, float*, int, int, int, int, int, cudaStream_t);
bool stem_tc_eligible(int, int, int, int, int, int, int, int);
const char* stem_wgrad_launch(const void*, const void*, void*, float*, int, int, int, int, int, cudaStream_t);
const char* tc_conv_wgrad(const void*, const void*, void*, float*, int, int, int, int, int, int, int, int, int, cudaStream_t);
const char* tc_conv_wgrad_halo(const void*, const void*, void*, float*, int, int, int, int, int, cudaStream_t);
bool tc_wgrad_halo_eligible(int, int, int, int, int, int, int, int);
const char* tc_conv_wgrad2(const void*, const void*, void*, float*, int, int, int, int, int, int, int, int, int, cudaStream_t);
bool tc_wgrad2_eligible(int, int);
static int g_wgrad_algo = -1;          // -1 auto, 0 per-tap, 1 halo (tap sharing), 2 CTA pair (256 x 256 tiles)
void tc_set_wgrad_algo(int a) { g_wgrad_algo = a; }

static const void* opt_ptr(const c10::optional<at::Tensor>& t, int64_t n) {
    if (!t.has_value()) return nullptr;
    TORCH_CHECK(t->scalar_type() == at::kBFloat16 && t->is_contiguous() && t->numel() == n, "bias must be contiguous bf16 of size N");
    return t->data_ptr();
}

// optional BatchNorm partial-statistics buffer for a convolution epilogue: fp32, zero on entry, at least [4 * #SMs][2][N]
static float* stats_ptr(const c10::optional<at::Tensor>& t, int64_t N) {
    if (!t.has_value()) return nullptr;
    TORCH_CHECK(t->is_cuda() && t->scalar_type() == at::kFloat && t->is_contiguous() && t->numel() >= (int64_t)4 * num_sms() * 2 * N,
                "stats must be a contiguous fp32 buffer of at least 4 * #SMs * 2 * N floats");
    return t->data_ptr<float>();
}

// y[M,N] = x[M,K] @ w[N,K]^T (+bias)(+relu); all bf16
at::Tensor gemm_tn(at::Tensor x, at::Tensor w, c10::optional<at::Tensor> bias, bool relu) {
    TORCH_CHECK(x.is_cuda() && w.is_cuda() && x.scalar_type() == at::kBFloat16 && w.scalar_type() == at::kBFloat16);
    TORCH_CHECK(x.dim() == 2 && w.dim() == 2 && x.is_contiguous() && w.is_contiguous() && x.size(1) == w.size(1));
    const int64_t M = x.size(0), K = x.size(1), N = w.size(0);
    TORCH_CHECK((reinterpret_cast<uintptr_t>(x.data_ptr()) % 16) == 0 && (reinterpret_cast<uintptr_t>(w.data_ptr()) % 16) == 0, "operands must be 16-byte aligned");
    at::Tensor y = at::empty({M, N}, x.options());
    const char* e = tc_gemm_tn(x.data_ptr(), w.data_ptr(), y.data_ptr(), opt_ptr(bias, N), (int)M, (int)N, (int)K, relu, at::cuda::getCurrentCUDAStream());
    TORCH_CHECK(e == nullptr, "tc_gemm_tn: ", e);
    return y;
}

// x: [NB,Cin,H,W] channels_last; w: [Cout,Cin,R,S] channels_last (physically [Cout][R][S][Cin]); returns channels_last y
at::Tensor conv_fprop(at::Tensor x, at::Tensor w, c10::optional<at::Tensor> bias, int64_t stride, int64_t pad, bool relu,
                      c10::optional<at::Tensor> residual, c10::optional<at::Tensor> stats) {
    TORCH_CHECK(x.is_cuda() && x.scalar_type() == at::kBFloat16 && w.scalar_type() == at::kBFloat16 && x.dim() == 4 && w.dim() == 4);
    TORCH_CHECK(x.is_contiguous(at::MemoryFormat::ChannelsLast), "x must be NHWC");
    TORCH_CHECK(w.is_contiguous(at::MemoryFormat::ChannelsLast), "w must be [Cout][R][S][Cin]");
    const int NB = x.size(0), Cin = x.size(1), H = x.size(2), W = x.size(3), Cout = w.size(0), R = w.size(2), S = w.size(3);
    TORCH_CHECK(w.size(1) == Cin);
    const int Ho = (H + 2 * pad - R) / stride + 1, Wo = (W + 2 * pad - S) / stride + 1;
    at::Tensor y = at::empty({NB, Cout, Ho, Wo}, x.options().memory_format(at::MemoryFormat::ChannelsLast));
    const void* rp = nullptr;
    if (residual.has_value()) {
        TORCH_CHECK(residual->scalar_type() == at::kBFloat16 && residual->sizes() == y.sizes() && residual->is_contiguous(at::MemoryFormat::ChannelsLast),
                    "residual must be an NHWC bf16 tensor of the output shape");
        rp = residual->data_ptr();
    }
    const char* e = tc_conv_fprop(x.data_ptr(), w.data_ptr(), y.data_ptr(), opt_ptr(bias, Cout), rp, NB, H, W, Cin, Cout, R, S, (int)stride, (int)pad, relu,
                                  stats_ptr(stats, Cout), at::cuda::getCurrentCUDAStream());
    TORCH_CHECK(e == nullptr, "tc_conv_fprop: ", e);
    return y;
}

// CTA-pair tap-table kernel (conv2_tcgen05.cu).
//   dgrad = false: a = x [NB,Cin,H,W] NHWC, w = the parameter [Cout,Cin,R,S] channels_last; returns y.
//   dgrad = true:  a = dy [NB,Cout,Ho,Wo] NHWC, w = the TRANSPOSED parameter (weight_transpose: logically [Cin,Cout,R,S] channels_last, physically
//                  [Cin][R][S][Cout]); in_h/in_w = spatial size of the forward input; returns dx [NB,Cin,in_h,in_w].  residual (same shape as the
//                  result) is added in the epilogue.
at::Tensor conv2(at::Tensor a, at::Tensor w, c10::optional<at::Tensor> bias, int64_t stride, int64_t pad, bool relu, c10::optional<at::Tensor> residual,
                 bool dgrad, int64_t in_h, int64_t in_w, c10::optional<at::Tensor> stats) {
    TORCH_CHECK(a.is_cuda() && a.scalar_type() == at::kBFloat16 && w.scalar_type() == at::kBFloat16 && a.dim() == 4 && w.dim() == 4);
    TORCH_CHECK(a.is_contiguous(at::MemoryFormat::ChannelsLast), "activations must be NHWC");
    TORCH_CHECK(w.is_contiguous(at::MemoryFormat::ChannelsLast), "weights must be [N][R][S][K]");
    const int NB = a.size(0), Cred = a.size(1), Ha = a.size(2), Wa = a.size(3), N = w.size(0), R = w.size(2), S = w.size(3);
    TORCH_CHECK(w.size(1) == Cred, "weight reduction dimension does not match the activations");
    int Ho, Wo;
    if (dgrad) { Ho = (int)in_h; Wo = (int)in_w; }
    else { Ho = (Ha + 2 * pad - R) / stride + 1; Wo = (Wa + 2 * pad - S) / stride + 1; }
    at::Tensor y = at::empty({NB, N, Ho, Wo}, a.options().memory_format(at::MemoryFormat::ChannelsLast));
    const void* rp = nullptr;
    if (residual.has_value()) {
        TORCH_CHECK(residual->scalar_type() == at::kBFloat16 && residual->sizes() == y.sizes() && residual->is_contiguous(at::MemoryFormat::ChannelsLast),
                    "residual must be an NHWC bf16 tensor of the output shape");
        rp = residual->data_ptr();
    }
    const char* e = tc_conv2(a.data_ptr(), w.data_ptr(), y.data_ptr(), opt_ptr(bias, N), rp, NB, Ha, Wa, Cred, N, R, S, (int)stride, (int)pad, dgrad ? 1 : 0, Ho, Wo,
                             relu, stats_ptr(stats, N), at::cuda::getCurrentCUDAStream());
    TORCH_CHECK(e == nullptr, "tc_conv2: ", e);
    return y;
}

// parameter [Cout,Cin,R,S] channels_last -> logically [Cin,Cout,R,S] channels_last (physically [Cin][R][S][Cout]); flip rotates the taps by 180 degrees
at::Tensor weight_transpose_t(at::Tensor w, bool flip) {
    TORCH_CHECK(w.is_cuda() && w.scalar_type() == at::kBFloat16 && w.dim() == 4 && w.is_contiguous(at::MemoryFormat::ChannelsLast));
    const int Cout = w.size(0), Cin = w.size(1), R = w.size(2), S = w.size(3);
    at::Tensor out = at::empty({Cin, Cout, R, S}, w.options().memory_format(at::MemoryFormat::ChannelsLast));
    const char* e = weight_transpose(w.data_ptr(), out.data_ptr(), Cout, R * S, Cin, flip ? 1 : 0, at::cuda::getCurrentCUDAStream());
    TORCH_CHECK(e == nullptr, e);
    return out;
}

// Batched form: `srcs[i]` ([Cout,Cin,R,S] channels_last) -> `dsts[i]` ([Cin,Cout,R,S] channels_last), flips[i] rotates the taps.  Returns the device
// table (uint8) to be passed to weight_transpose_run every step; tensors must keep their addresses (arena views).
std::vector<at::Tensor> weight_transpose_plan(std::vector<at::Tensor> srcs, std::vector<at::Tensor> dsts, std::vector<int64_t> flips) {
    struct E { const void* src; void* dst; int Cout, T, Cin, flip, first_block, bx, by, pad; };
    static_assert(sizeof(E) == 48, "table entry layout");
    TORCH_CHECK(srcs.size() == dsts.size() && srcs.size() == flips.size() && !srcs.empty() && srcs.size() <= 1024);
    std::vector<E> tab(srcs.size());
    int blocks = 0;
    for (size_t i = 0; i < srcs.size(); ++i) {
        const at::Tensor &w = srcs[i], &o = dsts[i];
        TORCH_CHECK(w.is_cuda() && w.scalar_type() == at::kBFloat16 && w.dim() == 4 && w.is_contiguous(at::MemoryFormat::ChannelsLast));
        TORCH_CHECK(o.is_cuda() && o.scalar_type() == at::kBFloat16 && o.dim() == 4 && o.is_contiguous(at::MemoryFormat::ChannelsLast));
        const int Cout = w.size(0), Cin = w.size(1), T = w.size(2) * w.size(3);
        TORCH_CHECK(o.size(0) == Cin && o.size(1) == Cout && o.size(2) == w.size(2) && o.size(3) == w.size(3));
        E e{w.data_ptr(), o.data_ptr(), Cout, T, Cin, (int)flips[i], blocks, (Cin + 31) / 32, (Cout + 31) / 32, 0};
        blocks += e.bx * e.by * T;
        tab[i] = e;
    }
    at::Tensor host = at::empty({(int64_t)(tab.size() * sizeof(E))}, at::TensorOptions().dtype(at::kByte));
    memcpy(host.data_ptr(), tab.data(), tab.size() * sizeof(E));
    at::Tensor dev = host.to(srcs[0].device());
    at::Tensor meta = at::empty({2}, at::TensorOptions().dtype(at::kLong));
    meta.data_ptr<int64_t>()[0] = (int64_t)tab.size(); meta.data_ptr<int64_t>()[1] = blocks;
    return {dev, meta};
}
// flags != 0: the kernel first acquires `nflags` "weights published" words >= *step_ptr (the consumer-side wait of the PS protocol, see conv2_tcgen05.cu)
void weight_transpose_run(at::Tensor table, int64_t n, int64_t blocks, int64_t flags, int64_t nflags, int64_t step_ptr, int64_t err, double timeout_s) {
    const char* e = weight_transpose_batched(table.data_ptr(), (int)n, (int)blocks, (const void*)flags, (int)nflags, (void*)step_ptr, (void*)err, timeout_s,
                                             at::cuda::getCurrentCUDAStream());
    TORCH_CHECK(e == nullptr, e);
}

// dW [Cout,Cin,R,S] channels_last (physically [Cout][R][S][Cin]) from dy (NHWC) and x (NHWC).  `scratch` is a persistent fp32 buffer of
// Cout*R*S*Cin floats that must be zero on entry and is left zeroed.
// `out`: write dw there instead of a fresh tensor (a channels_last view of the gradient arena: the push kernel's source, no gather copy)
static at::Tensor wgrad_out(const c10::optional<at::Tensor>& out, const at::Tensor& x, int64_t Cout, int64_t Cin, int64_t R, int64_t S) {
    if (!out.has_value()) return at::empty({Cout, Cin, R, S}, x.options().memory_format(at::MemoryFormat::ChannelsLast));
    const at::Tensor& o = *out;
    TORCH_CHECK(o.is_cuda() && o.scalar_type() == at::kBFloat16 && o.dim() == 4 && o.size(0) == Cout && o.size(1) == Cin && o.size(2) == R && o.size(3) == S &&
                o.is_contiguous(at::MemoryFormat::ChannelsLast), "wgrad out: need a bf16 channels_last [Cout,Cin,R,S] tensor");
    return o;
}
at::Tensor conv_wgrad(at::Tensor dy, at::Tensor x, at::Tensor scratch, int64_t R, int64_t S, int64_t stride, int64_t pad, c10::optional<at::Tensor> out) {
    TORCH_CHECK(dy.is_cuda() && dy.scalar_type() == at::kBFloat16 && x.scalar_type() == at::kBFloat16 && dy.dim() == 4 && x.dim() == 4);
    TORCH_CHECK(dy.is_contiguous(at::MemoryFormat::ChannelsLast) && x.is_contiguous(at::MemoryFormat::ChannelsLast), "dy and x must be NHWC");
    const int NB = x.size(0), Cin = x.size(1), H = x.size(2), W = x.size(3), Cout = dy.size(1);
    TORCH_CHECK(scratch.scalar_type() == at::kFloat && scratch.is_contiguous() && scratch.numel() >= (int64_t)Cout * R * S * Cin);
    at::Tensor dw = wgrad_out(out, x, Cout, Cin, R, S);
    const char* e;
    // measured (profiles/run28_conv_per_layer.md, after the MMA-issue fix): the CTA-pair kernel wins on 256-wide tiles (81 vs 105 us on 256->256 @8x8),
    // the tap-sharing kernel everywhere else it applies (110 vs 223 us on 64->64 @32x32, 81 vs 94 on 128->128 @16x16)
    const bool halo_auto = !tc_wgrad2_eligible(Cin, Cout);
    if ((g_wgrad_algo == 1 || (g_wgrad_algo == -1 && halo_auto)) && tc_wgrad_halo_eligible(H, W, Cin, Cout, (int)R, (int)S, (int)stride, (int)pad))
        e = tc_conv_wgrad_halo(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), scratch.data_ptr<float>(), NB, H, W, Cin, Cout, at::cuda::getCurrentCUDAStream());
    else if ((g_wgrad_algo == 2 || g_wgrad_algo == -1) && tc_wgrad2_eligible(Cin, Cout))
        e = tc_conv_wgrad2(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), scratch.data_ptr<float>(), NB, H, W, Cin, Cout, (int)R, (int)S, (int)stride, (int)pad,
                           at::cuda::getCurrentCUDAStream());
    else
        e = tc_conv_wgrad(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), scratch.data_ptr<float>(), NB, H, W, Cin, Cout, (int)R, (int)S, (int)stride, (int)pad,
                          at::cuda::getCurrentCUDAStream());
    TORCH_CHECK(e == nullptr, "tc_conv_wgrad: ", e);
    return dw;
}

// network stem on the tensor cores (stem_tcgen05.cu): y = conv3x3/s1/p1(x [NB,Cin<=3,H,W] NHWC, w [64,Cin,3,3] channels_last) (+bias)(+relu)
at::Tensor stem_conv(at::Tensor x, at::Tensor w, c10::optional<at::Tensor> bias, bool relu, c10::optional<at::Tensor> stats) {
    TORCH_CHECK(x.is_cuda() && x.scalar_type() == at::kBFloat16 && w.scalar_type() == at::kBFloat16 && x.dim() == 4 && w.dim() == 4);
    TORCH_CHECK(x.is_contiguous(at::MemoryFormat::ChannelsLast) && w.is_contiguous(at::MemoryFormat::ChannelsLast), "x and w must be channels_last");
    const int NB = x.size(0), Cin = x.size(1), H = x.size(2), W = x.size(3), Cout = w.size(0);
    TORCH_CHECK(w.size(1) == Cin && stem_tc_eligible(H, W, Cin, Cout, w.size(2), w.size(3), 1, 1), "shape not eligible for the stem kernel");
    at::Tensor y = at::empty({NB, Cout, H, W}, x.options().memory_format(at::MemoryFormat::ChannelsLast));
    const char* e = stem_fprop(x.data_ptr(), w.data_ptr(), y.data_ptr(), opt_ptr(bias, Cout), NB, H, W, Cin, Cout, relu, stats_ptr(stats, Cout),
                               at::cuda::getCurrentCUDAStream());
    TORCH_CHECK(e == nullptr, e);
    return y;
}
at::Tensor stem_wgrad_tensorcore(at::Tensor dy, at::Tensor x, at::Tensor scratch, c10::optional<at::Tensor> out) {
    TORCH_CHECK(dy.is_cuda() && dy.scalar_type() == at::kBFloat16 && x.scalar_type() == at::kBFloat16 && dy.dim() == 4 && x.dim() == 4);
    TORCH_CHECK(dy.is_contiguous(at::MemoryFormat::ChannelsLast) && x.is_contiguous(at::MemoryFormat::ChannelsLast), "dy and x must be NHWC");
    const int NB = x.size(0), Cin = x.size(1), H = x.size(2), W = x.size(3), Cout = dy.size(1);
    TORCH_CHECK(dy.size(2) == H && dy.size(3) == W && scratch.scalar_type() == at::kFloat && scratch.numel() >= (int64_t)Cout * 9 * Cin);
    at::Tensor dw = wgrad_out(out, x, Cout, Cin, 3, 3);
    const char* e = stem_wgrad_tc(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), scratch.data_ptr<float>(), NB, H, W, Cin, Cout, at::cuda::getCurrentCUDAStream());
    TORCH_CHECK(e == nullptr, e);
    return dw;
}
bool stem_eligible(int64_t H, int64_t W, int64_t Cin, int64_t Cout, int64_t R, int64_t S, int64_t stride, int64_t pad) {
    return stem_tc_eligible((int)H, (int)W, (int)Cin, (int)Cout, (int)R, (int)S, (int)stride, (int)pad);
}

// weight gradient of a 3x3/s1/p1 convolution with Cin <= 3 (network stem); dy, x NHWC bf16; scratch: zeroed fp32 [Cout*9*Cin], left zeroed
at::Tensor stem_wgrad(at::Tensor dy, at::Tensor x, at::Tensor scratch) {
    TORCH_CHECK(dy.is_cuda() && dy.scalar_type() == at::kBFloat16 && x.scalar_type() == at::kBFloat16 && dy.dim() == 4 && x.dim() == 4);
    TORCH_CHECK(dy.is_contiguous(at::MemoryFormat::ChannelsLast) && x.is_contiguous(at::MemoryFormat::ChannelsLast), "dy and x must be NHWC");
    const int N = x.size(0), Cin = x.size(1), H = x.size(2), W = x.size(3), Cout = dy.size(1);
    TORCH_CHECK(dy.size(2) == H && dy.size(3) == W && scratch.scalar_type() == at::kFloat && scratch.numel() >= (int64_t)Cout * 9 * Cin);
    at::Tensor dw = at::empty({Cout, Cin, 3, 3}, x.options().memory_format(at::MemoryFormat::ChannelsLast));
    const char* e = stem_wgrad_launch(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), scratch.data_ptr<float>(), N, H, W, Cin, Cout, at::cuda::getCurrentCUDAStream());
    TORCH_CHECK(e == nullptr, e);
    return dw;
}

std::string conv_unsupported_reason(std::vector<int64_t> xs, std::vector<int64_t> ws, std::vector<int64_t> stride, std::vector<int64_t> pad,
                                    std::vector<int64_t> dil, int64_t groups, bool bf16) {
    if (!bf16) return "needs bf16 operands";
    if (groups != 1) return "grouped convolution";
    if (dil[0] != 1 || dil[1] != 1) return "dilated convolution";
    if (stride[0] != stride[1] || pad[0] != pad[1]) return "anisotropic stride/padding";
    if (xs[1] % 64) return "Cin is not a multiple of 64";
    if (ws[0] % 8) return "Cout is not a multiple of 8";
    return "";
}

int64_t tc_error() { return tc_check_error(); }
}  // namespace psb

void bind_gemm(pybind11::module_& m) {
    m.def("gemm_tn", &psb::gemm_tn);
    m.def("conv_fprop", &psb::conv_fprop, pybind11::arg("x"), pybind11::arg("w"), pybind11::arg("bias"), pybind11::arg("stride"), pybind11::arg("pad"),
          pybind11::arg("relu"), pybind11::arg("residual") = c10::optional<at::Tensor>(), pybind11::arg("stats") = c10::optional<at::Tensor>());
    m.def("conv2", &psb::conv2, pybind11::arg("a"), pybind11::arg("w"), pybind11::arg("bias"), pybind11::arg("stride"), pybind11::arg("pad"), pybind11::arg("relu"),
          pybind11::arg("residual") = c10::optional<at::Tensor>(), pybind11::arg("dgrad") = false, pybind11::arg("in_h") = 0, pybind11::arg("in_w") = 0,
          pybind11::arg("stats") = c10::optional<at::Tensor>());
    m.def("num_sms", &psb::num_sms);
    m.def("weight_transpose", &psb::weight_transpose_t);
    m.def("weight_transpose_plan", &psb::weight_transpose_plan);
    m.def("weight_transpose_run", &psb::weight_transpose_run, pybind11::arg("table"), pybind11::arg("n"), pybind11::arg("blocks"), pybind11::arg("flags") = 0,
          pybind11::arg("nflags") = 0, pybind11::arg("step_ptr") = 0, pybind11::arg("err") = 0, pybind11::arg("timeout_s") = 0.0);
    m.def("tc_set_conv2_bn", &psb::tc_set_conv2_bn);
    m.def("tc_set_conv_algo", &psb::tc_set_conv_algo);
    m.def("tc_set_halo_desc_mode", &psb::tc_set_halo_desc_mode);
    m.def("tc_set_halo_cluster", &psb::tc_set_halo_cluster);
    m.def("tc_set_halo_pair", &psb::tc_set_halo_pair);
    m.def("tc_set_halo_pitch", &psb::tc_set_halo_pitch);
    m.def("tc_set_wgrad_sms", &psb::tc_set_wgrad_sms);
    m.def("tc_set_halo_resident", &psb::tc_set_halo_resident);
    m.def("tc_set_gemm_bn", &psb::tc_set_gemm_bn);
    m.def("conv_unsupported_reason", &psb::conv_unsupported_reason);
    m.def("conv_wgrad", &psb::conv_wgrad, pybind11::arg("dy"), pybind11::arg("x"), pybind11::arg("scratch"), pybind11::arg("R"), pybind11::arg("S"),
          pybind11::arg("stride"), pybind11::arg("pad"), pybind11::arg("out") = c10::optional<at::Tensor>());
    m.def("tc_set_wgrad_algo", &psb::tc_set_wgrad_algo);
    m.def("stem_wgrad", &psb::stem_wgrad);
    m.def("stem_conv", &psb::stem_conv);
    m.def("stem_wgrad_tc", &psb::stem_wgrad_tensorcore, pybind11::arg("dy"), pybind11::arg("x"), pybind11::arg("scratch"), pybind11::arg("out") = c10::optional<at::Tensor>());
    m.def("stem_eligible", &psb::stem_eligible);
    m.def("tc_error", &psb::tc_error);
}

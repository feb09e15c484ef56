// Fused flat-arena optimizers: one launch updates every parameter of the model (fp32 master + state) from an
// fp32 or bf16 gradient arena and optionally emits the bf16 compute copy.  Memory-bound streaming kernels:
// 16-byte vector accesses, grid = 148 SMs x 8 CTAs.  Reference semantics: src/optim/sgd.py:59-92, adam.py:44-94.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include "optim_math.cuh"

namespace psb {

template <typename G> struct GradLoad;
template <> struct GradLoad<float> {
    static __device__ __forceinline__ void load8(const float* g, long long i, float* f) {
        uint4 a = ld_nc_v4(g + i), b = ld_nc_v4(g + i + 4);
        f[0] = __uint_as_float(a.x); f[1] = __uint_as_float(a.y); f[2] = __uint_as_float(a.z); f[3] = __uint_as_float(a.w);
        f[4] = __uint_as_float(b.x); f[5] = __uint_as_float(b.y); f[6] = __uint_as_float(b.z); f[7] = __uint_as_float(b.w);
    }
};
template <> struct GradLoad<__nv_bfloat16> {
    static __device__ __forceinline__ void load8(const __nv_bfloat16* g, long long i, float* f) {
        unpack8(ld_nc_v4(g + i), f);
    }
};

__device__ __forceinline__ void ld8f(const float* p, float* f) {
    uint4 a = ld_v4(p), b = ld_v4(p + 4);
    f[0] = __uint_as_float(a.x); f[1] = __uint_as_float(a.y); f[2] = __uint_as_float(a.z); f[3] = __uint_as_float(a.w);
    f[4] = __uint_as_float(b.x); f[5] = __uint_as_float(b.y); f[6] = __uint_as_float(b.z); f[7] = __uint_as_float(b.w);
}
__device__ __forceinline__ void st8f(float* p, const float* f) {
    st_v4(p, make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3])));
    st_v4(p + 4, make_uint4(__float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]), __float_as_uint(f[7])));
}

template <typename G>
__global__ void __launch_bounds__(256) sgd_flat_kernel(float* __restrict__ w, const G* __restrict__ g, float* __restrict__ buf,
                                                       __nv_bfloat16* __restrict__ out_lowp, long long n, OptHP hp, int first, float scale) {
    const long long stride = (long long)gridDim.x * blockDim.x * 8;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
        float gf[8], wf[8], bf[8];
        GradLoad<G>::load8(g, i, gf);
        ld8f(w + i, wf);
        if (hp.momentum != 0.f && !first) ld8f(buf + i, bf);
#pragma unroll
        for (int k = 0; k < 8; ++k) wf[k] = sgd_update(wf[k], gf[k] * scale, bf[k], hp, first != 0);
        st8f(w + i, wf);
        if (hp.momentum != 0.f) st8f(buf + i, bf);
        if (out_lowp) st_v4(out_lowp + i, pack8(wf));
    }
}

template <typename G>
__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ w, const G* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, float* __restrict__ vmax,
                                                        __nv_bfloat16* __restrict__ out_lowp, long long n, OptHP hp,
                                                        float step_size, float scale) {
    const long long stride = (long long)gridDim.x * blockDim.x * 8;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
        float gf[8], wf[8], mf[8], vf[8], xf[8];
        GradLoad<G>::load8(g, i, gf);
        ld8f(w + i, wf); ld8f(m + i, mf); ld8f(v + i, vf);
        if (vmax) ld8f(vmax + i, xf);
#pragma unroll
        for (int k = 0; k < 8; ++k) wf[k] = adam_update(wf[k], gf[k] * scale, mf[k], vf[k], vmax ? &xf[k] : nullptr, hp, step_size);
        st8f(w + i, wf); st8f(m + i, mf); st8f(v + i, vf);
        if (vmax) st8f(vmax + i, xf);
        if (out_lowp) st_v4(out_lowp + i, pack8(wf));
    }
}

static int grid_for(long long n) {
    long long blocks = (n / 8 + 255) / 256;
    long long cap = (long long)kNumSMs * 8;
    return (int)std::max<long long>(1, std::min(blocks, cap));
}

void sgd_flat(at::Tensor w, at::Tensor g, at::Tensor buf, double lr, double momentum, double dampening, double wd,
              bool nesterov, bool first, double scale, c10::optional<at::Tensor> out_lowp) {
    TORCH_CHECK(w.is_cuda() && w.scalar_type() == at::kFloat && w.is_contiguous() && g.is_contiguous() && buf.is_contiguous());
    const long long n = w.numel();
    TORCH_CHECK(n % 8 == 0 && g.numel() == n && buf.numel() == n, "flat arenas must be padded to a multiple of 8");
    OptHP hp{};
    hp.lr = (float)lr; hp.momentum = (float)momentum; hp.dampening = (float)dampening; hp.weight_decay = (float)wd;
    hp.nesterov = nesterov;
    __nv_bfloat16* lp = nullptr;
    if (out_lowp.has_value()) {
        TORCH_CHECK(out_lowp->scalar_type() == at::kBFloat16 && out_lowp->numel() == n && out_lowp->is_contiguous());
        lp = reinterpret_cast<__nv_bfloat16*>(out_lowp->data_ptr());
    }
    auto stream = at::cuda::getCurrentCUDAStream();
    const int grid = grid_for(n);
    if (g.scalar_type() == at::kFloat)
        sgd_flat_kernel<float><<<grid, 256, 0, stream>>>(w.data_ptr<float>(), g.data_ptr<float>(), buf.data_ptr<float>(), lp, n, hp, first, (float)scale);
    else if (g.scalar_type() == at::kBFloat16)
        sgd_flat_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(w.data_ptr<float>(), reinterpret_cast<const __nv_bfloat16*>(g.data_ptr()),
                                                                 buf.data_ptr<float>(), lp, n, hp, first, (float)scale);
    else
        TORCH_CHECK(false, "gradient arena must be fp32 or bf16");
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void adam_flat(at::Tensor w, at::Tensor g, at::Tensor m, at::Tensor v, c10::optional<at::Tensor> vmax, double step_size,
               double beta1, double beta2, double eps, double wd, double scale, c10::optional<at::Tensor> out_lowp) {
    TORCH_CHECK(w.is_cuda() && w.scalar_type() == at::kFloat && w.is_contiguous());
    const long long n = w.numel();
    TORCH_CHECK(n % 8 == 0 && g.numel() == n && m.numel() == n && v.numel() == n);
    OptHP hp{};
    hp.beta1 = (float)beta1; hp.beta2 = (float)beta2; hp.eps = (float)eps; hp.weight_decay = (float)wd;
    float* vm = vmax.has_value() ? vmax->data_ptr<float>() : nullptr;
    __nv_bfloat16* lp = nullptr;
    if (out_lowp.has_value()) {
        TORCH_CHECK(out_lowp->scalar_type() == at::kBFloat16 && out_lowp->numel() == n);
        lp = reinterpret_cast<__nv_bfloat16*>(out_lowp->data_ptr());
    }
    auto stream = at::cuda::getCurrentCUDAStream();
    const int grid = grid_for(n);
    if (g.scalar_type() == at::kFloat)
        adam_flat_kernel<float><<<grid, 256, 0, stream>>>(w.data_ptr<float>(), g.data_ptr<float>(), m.data_ptr<float>(), v.data_ptr<float>(), vm, lp, n, hp, (float)step_size, (float)scale);
    else if (g.scalar_type() == at::kBFloat16)
        adam_flat_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(w.data_ptr<float>(), reinterpret_cast<const __nv_bfloat16*>(g.data_ptr()),
                                                                  m.data_ptr<float>(), v.data_ptr<float>(), vm, lp, n, hp, (float)step_size, (float)scale);
    else
        TORCH_CHECK(false, "gradient arena must be fp32 or bf16");
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

}  // namespace psb

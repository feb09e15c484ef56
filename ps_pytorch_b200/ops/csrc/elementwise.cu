// Small memory-bound layers of the reference models as sm_100a streaming kernels (NHWC bf16, 16-byte vectors):
// ReLU, MaxPool2d(2,2) (+ fused ReLU, LeNet's pool->relu), global average pool, and a fused softmax-cross-entropy
// that produces the loss and dL/dlogits in one pass (reference ops: /root/reference/src/model_ops/lenet.py:26-31,
// vgg.py:51, resnet.py:94, distributed_worker.py:98).
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include "common.cuh"

namespace psb {

static int ew_grid(long long items, int per_block = 256) {
    long long blocks = (items + per_block - 1) / per_block;
    return (int)std::max<long long>(1, std::min<long long>(blocks, (long long)kNumSMs * 8));
}

// ---- ReLU --------------------------------------------------------------------------------------------------------------
__global__ void relu_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, long long nvec) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        float f[8]; unpack8(ld_nc_v4(x + i), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
        st_v4(y + i, pack8(f));
    }
}
__global__ void relu_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ y, uint4* __restrict__ dx, long long nvec) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        float d[8], v[8]; unpack8(ld_nc_v4(dy + i), d); unpack8(ld_nc_v4(y + i), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = v[j] > 0.f ? d[j] : 0.f;
        st_v4(dx + i, pack8(d));
    }
}
at::Tensor relu_forward(at::Tensor x) {
    TORCH_CHECK(x.is_cuda() && x.scalar_type() == at::kBFloat16 && x.numel() % 8 == 0);
    at::Tensor xc = x;   // caller guarantees a dense layout (contiguous or channels_last)
    at::Tensor y = at::empty_like(xc);
    relu_fwd_kernel<<<ew_grid(xc.numel() / 8), 256, 0, at::cuda::getCurrentCUDAStream()>>>((const uint4*)xc.data_ptr(), (uint4*)y.data_ptr(), xc.numel() / 8);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    return y;
}
at::Tensor relu_backward(at::Tensor dy, at::Tensor y) {
    TORCH_CHECK(dy.scalar_type() == at::kBFloat16 && dy.strides() == y.strides() && y.is_non_overlapping_and_dense());
    at::Tensor dx = at::empty_like(y);
    relu_bwd_kernel<<<ew_grid(y.numel() / 8), 256, 0, at::cuda::getCurrentCUDAStream()>>>((const uint4*)dy.data_ptr(), (const uint4*)y.data_ptr(), (uint4*)dx.data_ptr(), y.numel() / 8);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    return dx;
}

// ---- MaxPool 2x2/2, NHWC --------------------------------------------------------------------------------------------------
template <bool RELU>
__global__ void maxpool_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int N, int H, int W, int LC) {
    const int Ho = H / 2, Wo = W / 2;
    const long long total = (long long)N * Ho * Wo * LC;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int l = (int)(i % LC); long long p = i / LC;
        const int wo = (int)(p % Wo); p /= Wo;
        const int ho = (int)(p % Ho); const long long n = p / Ho;
        const long long base = ((n * H + 2 * ho) * W + 2 * wo) * LC + l;
        float a[8], b[8], c[8], d[8];
        unpack8(ld_nc_v4(x + base), a); unpack8(ld_nc_v4(x + base + LC), b);
        unpack8(ld_nc_v4(x + base + (long long)W * LC), c); unpack8(ld_nc_v4(x + base + (long long)W * LC + LC), d);
#pragma unroll
        for (int j = 0; j < 8; ++j) { float m = fmaxf(fmaxf(a[j], b[j]), fmaxf(c[j], d[j])); a[j] = RELU ? fmaxf(m, 0.f) : m; }
        st_v4(y + i, pack8(a));
    }
}
// gradient goes to the first maximal element in window order (ATen's tie rule); with RELU only where the pooled value > 0
template <bool RELU>
__global__ void maxpool_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x, uint4* __restrict__ dx, int N, int H, int W, int LC) {
    const int Ho = H / 2, Wo = W / 2;
    const long long total = (long long)N * Ho * Wo * LC;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int l = (int)(i % LC); long long p = i / LC;
        const int wo = (int)(p % Wo); p /= Wo;
        const int ho = (int)(p % Ho); const long long n = p / Ho;
        const long long base = ((n * H + 2 * ho) * W + 2 * wo) * LC + l;
        const long long o1 = LC, o2 = (long long)W * LC, o3 = o2 + LC;
        float a[8], b[8], c[8], d[8], g[8], ga[8], gb[8], gc[8], gd[8];
        unpack8(ld_nc_v4(x + base), a); unpack8(ld_nc_v4(x + base + o1), b);
        unpack8(ld_nc_v4(x + base + o2), c); unpack8(ld_nc_v4(x + base + o3), d);
        unpack8(ld_nc_v4(dy + i), g);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int arg = 0; float m = a[j];
            if (b[j] > m) { m = b[j]; arg = 1; }
            if (c[j] > m) { m = c[j]; arg = 2; }
            if (d[j] > m) { m = d[j]; arg = 3; }
            const float gg = (RELU && !(m > 0.f)) ? 0.f : g[j];
            ga[j] = arg == 0 ? gg : 0.f; gb[j] = arg == 1 ? gg : 0.f; gc[j] = arg == 2 ? gg : 0.f; gd[j] = arg == 3 ? gg : 0.f;
        }
        st_v4(dx + base, pack8(ga)); st_v4(dx + base + o1, pack8(gb)); st_v4(dx + base + o2, pack8(gc)); st_v4(dx + base + o3, pack8(gd));
    }
}
static void check_nhwc(const at::Tensor& t) {
    TORCH_CHECK(t.is_cuda() && t.scalar_type() == at::kBFloat16 && t.dim() == 4 && t.is_contiguous(at::MemoryFormat::ChannelsLast) && t.size(1) % 8 == 0,
                "NHWC bf16 tensor with C % 8 == 0 expected");
}
at::Tensor maxpool2x2_forward(at::Tensor x, bool relu) {
    check_nhwc(x);
    const int N = x.size(0), C = x.size(1), H = x.size(2), W = x.size(3);
    TORCH_CHECK(H % 2 == 0 && W % 2 == 0);
    at::Tensor y = at::empty({N, C, H / 2, W / 2}, x.options().memory_format(at::MemoryFormat::ChannelsLast));
    const long long total = (long long)N * (H / 2) * (W / 2) * (C / 8);
    auto s = at::cuda::getCurrentCUDAStream();
    if (relu) maxpool_fwd_kernel<true><<<ew_grid(total), 256, 0, s>>>((const uint4*)x.data_ptr(), (uint4*)y.data_ptr(), N, H, W, C / 8);
    else maxpool_fwd_kernel<false><<<ew_grid(total), 256, 0, s>>>((const uint4*)x.data_ptr(), (uint4*)y.data_ptr(), N, H, W, C / 8);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    return y;
}
at::Tensor maxpool2x2_backward(at::Tensor dy, at::Tensor x, bool relu) {
    check_nhwc(x); check_nhwc(dy);
    const int N = x.size(0), C = x.size(1), H = x.size(2), W = x.size(3);
    at::Tensor dx = at::empty_like(x);
    const long long total = (long long)N * (H / 2) * (W / 2) * (C / 8);
    auto s = at::cuda::getCurrentCUDAStream();
    if (relu) maxpool_bwd_kernel<true><<<ew_grid(total), 256, 0, s>>>((const uint4*)dy.data_ptr(), (const uint4*)x.data_ptr(), (uint4*)dx.data_ptr(), N, H, W, C / 8);
    else maxpool_bwd_kernel<false><<<ew_grid(total), 256, 0, s>>>((const uint4*)dy.data_ptr(), (const uint4*)x.data_ptr(), (uint4*)dx.data_ptr(), N, H, W, C / 8);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    return dx;
}

// ---- global average pool: [N,H,W,C] -> [N,1,1,C] ---------------------------------------------------------------------------
__global__ void gap_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, long long N, int HW, int LC) {
    const float inv = 1.f / (float)HW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N * LC; i += (long long)gridDim.x * blockDim.x) {
        const long long n = i / LC; const int l = (int)(i % LC);
        float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, f[8];
        for (int p = 0; p < HW; ++p) { unpack8(ld_nc_v4(x + (n * HW + p) * LC + l), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) s[j] += f[j]; }
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] *= inv;
        st_v4(y + i, pack8(s));
    }
}
__global__ void gap_bwd_kernel(const uint4* __restrict__ dy, uint4* __restrict__ dx, long long N, int HW, int LC) {
    const float inv = 1.f / (float)HW;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N * HW * LC; i += (long long)gridDim.x * blockDim.x) {
        const int l = (int)(i % LC); const long long n = i / ((long long)HW * LC);
        float f[8]; unpack8(ld_nc_v4(dy + n * LC + l), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] *= inv;
        st_v4(dx + i, pack8(f));
    }
}
at::Tensor gap_forward(at::Tensor x) {
    check_nhwc(x);
    const long long N = x.size(0); const int C = x.size(1), HW = x.size(2) * x.size(3);
    at::Tensor y = at::empty({N, C, 1, 1}, x.options().memory_format(at::MemoryFormat::ChannelsLast));
    gap_fwd_kernel<<<ew_grid(N * (C / 8)), 256, 0, at::cuda::getCurrentCUDAStream()>>>((const uint4*)x.data_ptr(), (uint4*)y.data_ptr(), N, HW, C / 8);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    return y;
}
at::Tensor gap_backward(at::Tensor dy, int64_t H, int64_t W) {
    TORCH_CHECK(dy.is_cuda() && dy.scalar_type() == at::kBFloat16 && dy.size(1) % 8 == 0);
    at::Tensor dyc = dy.contiguous(at::MemoryFormat::ChannelsLast);
    const long long N = dy.size(0); const int C = dy.size(1);
    at::Tensor dx = at::empty({N, C, H, W}, dy.options().memory_format(at::MemoryFormat::ChannelsLast));
    gap_bwd_kernel<<<ew_grid(N * H * W * (C / 8)), 256, 0, at::cuda::getCurrentCUDAStream()>>>((const uint4*)dyc.data_ptr(), (uint4*)dx.data_ptr(), N, (int)(H * W), C / 8);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    return dx;
}

// ---- softmax cross-entropy (mean), forward + gradient in one pass --------------------------------------------------------------
// one warp per row; row_loss[b] = logsumexp - logit[target]; dlogits = (softmax - onehot) / B
template <typename T>
__global__ void softmax_ce_kernel(const T* __restrict__ logits, const long long* __restrict__ target, float* __restrict__ row_loss,
                                  T* __restrict__ dlogits, int B, int NC) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= B) return;
    const T* row = logits + (long long)warp * NC;
    float mx = -INFINITY;
    for (int c = lane; c < NC; c += 32) mx = fmaxf(mx, to_f(row[c]));
    mx = warp_max(mx);
    float se = 0.f;
    for (int c = lane; c < NC; c += 32) se += __expf(to_f(row[c]) - mx);
    se = warp_sum(se);
    const int t = (int)target[warp];
    const float lse = mx + __logf(se);
    if (lane == 0) row_loss[warp] = lse - to_f(row[t]);
    const float invB = 1.f / (float)B, inv_se = 1.f / se;
    T* drow = dlogits + (long long)warp * NC;
    for (int c = lane; c < NC; c += 32) {
        const float p = __expf(to_f(row[c]) - mx) * inv_se;
        drow[c] = from_f<T>((p - (c == t ? 1.f : 0.f)) * invB);
    }
}
__global__ void mean_kernel(const float* __restrict__ v, int n, float* __restrict__ out) {
    __shared__ float sm[32];
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += v[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.f;
        s = warp_sum(s);
        if (threadIdx.x == 0) *out = s / (float)n;
    }
}
std::vector<at::Tensor> softmax_ce(at::Tensor logits, at::Tensor target) {
    TORCH_CHECK(logits.is_cuda() && logits.dim() == 2 && logits.is_contiguous() && target.scalar_type() == at::kLong && target.is_contiguous());
    const int B = logits.size(0), NC = logits.size(1);
    auto s = at::cuda::getCurrentCUDAStream();
    at::Tensor row_loss = at::empty({B}, logits.options().dtype(at::kFloat));
    at::Tensor loss = at::empty({}, logits.options().dtype(at::kFloat));
    at::Tensor dl = at::empty_like(logits);
    const int blocks = (B * 32 + 255) / 256;
    if (logits.scalar_type() == at::kBFloat16)
        softmax_ce_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>((const __nv_bfloat16*)logits.data_ptr(), target.data_ptr<int64_t>() ? (const long long*)target.data_ptr<int64_t>() : nullptr,
                                                               row_loss.data_ptr<float>(), (__nv_bfloat16*)dl.data_ptr(), B, NC);
    else if (logits.scalar_type() == at::kFloat)
        softmax_ce_kernel<float><<<blocks, 256, 0, s>>>(logits.data_ptr<float>(), (const long long*)target.data_ptr<int64_t>(), row_loss.data_ptr<float>(), dl.data_ptr<float>(), B, NC);
    else TORCH_CHECK(false, "softmax_ce: bf16 or fp32 logits");
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    mean_kernel<<<1, 256, 0, s>>>(row_loss.data_ptr<float>(), B, loss.data_ptr<float>());
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    return {loss, dl};
}

}  // namespace psb

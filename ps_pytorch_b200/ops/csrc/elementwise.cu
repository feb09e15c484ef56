// Small memory-bound layers of the reference models as sm_100a streaming kernels (NHWC bf16, 16-byte vectors):
// ReLU, MaxPool2d(2,2) (+ fused ReLU, LeNet's pool->relu), MaxPool2d(3,2,1) (ImageNet stem), Dropout, global average pool, and a fused softmax-cross-entropy
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

// ---- MaxPool 3x3 / stride 2 / pad 1, NHWC (the ImageNet ResNet stem) ---------------------------------------------------------
// forward: max over the (clipped) window.  backward is a GATHER over dx pixels — every input pixel belongs to at most 2 x 2 windows; it
// receives dy of a window iff it is that window's first maximum in window order (ATen's tie rule) — so no atomics and dx is written once.
__global__ void maxpool3s2_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int N, int H, int W, int Ho, int Wo, int LC) {
    const long long total = (long long)N * Ho * Wo * LC;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int l = (int)(i % LC); long long p = i / LC;
        const int wo = (int)(p % Wo); p /= Wo;
        const int ho = (int)(p % Ho); const long long n = p / Ho;
        float m[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
        for (int r = 0; r < 3; ++r) {
            const int h = 2 * ho - 1 + r;
            if (h < 0 || h >= H) continue;
            for (int c = 0; c < 3; ++c) {
                const int w = 2 * wo - 1 + c;
                if (w < 0 || w >= W) continue;
                float a[8];
                unpack8(ld_nc_v4(x + ((n * H + h) * W + w) * LC + l), a);
#pragma unroll
                for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], a[j]);
            }
        }
        st_v4(y + i, pack8(m));
    }
}
__global__ void maxpool3s2_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x, const uint4* __restrict__ y, uint4* __restrict__ dx,
                                      int N, int H, int W, int Ho, int Wo, int LC) {
    const long long total = (long long)N * H * W * LC;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int l = (int)(i % LC); long long p = i / LC;
        const int w = (int)(p % W); p /= W;
        const int h = (int)(p % H); const long long n = p / H;
        float xv[8], g[8];
        unpack8(ld_nc_v4(x + i), xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = 0.f;
        // windows (ho, wo) that contain (h, w): 2*ho - 1 <= h <= 2*ho + 1
        const int ho0 = h / 2, ho1 = (h + 1) / 2, wo0 = w / 2, wo1 = (w + 1) / 2;
        for (int ho = ho0; ho <= ho1; ++ho) {
            if (ho >= Ho) continue;
            for (int wo = wo0; wo <= wo1; ++wo) {
                if (wo >= Wo) continue;
                const long long o = ((n * Ho + ho) * Wo + wo) * LC + l;
                float yv[8], dv[8];
                unpack8(ld_nc_v4(y + o), yv); unpack8(ld_nc_v4(dy + o), dv);
                // is (h, w) the FIRST element of this window (row-major) equal to the window maximum?
                bool first[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) first[j] = xv[j] == yv[j];
                for (int r = 0; r < 3; ++r) {
                    const int hh = 2 * ho - 1 + r;
                    if (hh < 0 || hh >= H) continue;
                    for (int c = 0; c < 3; ++c) {
                        const int ww = 2 * wo - 1 + c;
                        if (ww < 0 || ww >= W) continue;
                        if (hh > h || (hh == h && ww >= w)) continue;            // only elements BEFORE (h, w) in window order
                        float a[8];
                        unpack8(ld_nc_v4(x + ((n * H + hh) * W + ww) * LC + l), a);
#pragma unroll
                        for (int j = 0; j < 8; ++j) if (a[j] == yv[j]) first[j] = false;
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) if (first[j]) g[j] += dv[j];
            }
        }
        st_v4(dx + i, pack8(g));
    }
}
at::Tensor maxpool3x3s2_forward(at::Tensor x) {
    check_nhwc(x);
    const int N = x.size(0), C = x.size(1), H = x.size(2), W = x.size(3);
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    at::Tensor y = at::empty({N, C, Ho, Wo}, x.options().memory_format(at::MemoryFormat::ChannelsLast));
    const long long total = (long long)N * Ho * Wo * (C / 8);
    maxpool3s2_fwd_kernel<<<ew_grid(total), 256, 0, at::cuda::getCurrentCUDAStream()>>>((const uint4*)x.data_ptr(), (uint4*)y.data_ptr(), N, H, W, Ho, Wo, C / 8);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    return y;
}
at::Tensor maxpool3x3s2_backward(at::Tensor dy, at::Tensor x, at::Tensor y) {
    check_nhwc(x); check_nhwc(dy); check_nhwc(y);
    const int N = x.size(0), C = x.size(1), H = x.size(2), W = x.size(3), Ho = y.size(2), Wo = y.size(3);
    at::Tensor dx = at::empty_like(x);
    const long long total = (long long)N * H * W * (C / 8);
    maxpool3s2_bwd_kernel<<<ew_grid(total), 256, 0, at::cuda::getCurrentCUDAStream()>>>((const uint4*)dy.data_ptr(), (const uint4*)x.data_ptr(), (const uint4*)y.data_ptr(),
                                                                                         (uint4*)dx.data_ptr(), N, H, W, Ho, Wo, C / 8);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    return dx;
}

// ---- Dropout (VGG classifier): counter-based mask, nothing stored ---------------------------------------------------------------------
// keep(i) = hash(seed, i) >= p * 2^32; y = keep ? x / (1 - p) : 0.  The backward pass recomputes the same mask from the same (seed, i).
// `seed_ptr` is a device word (bumped once per training step by the owner of the step), so a CUDA graph replays with fresh masks; the
// value used is written to `used_seed` for the backward pass.
__device__ __forceinline__ uint32_t mix32(uint64_t v) {
    v ^= v >> 33; v *= 0xff51afd7ed558ccdULL; v ^= v >> 33; v *= 0xc4ceb9fe1a85ec53ULL; v ^= v >> 33;
    return (uint32_t)v;
}
__global__ void dropout_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, long long nvec, float p, const long long* __restrict__ seed_ptr,
                               long long salt, long long* __restrict__ used_seed) {
    const long long seed = seed_ptr[0] * 0x9E3779B97F4A7C15LL + salt;
    if (used_seed != nullptr && blockIdx.x == 0 && threadIdx.x == 0) used_seed[0] = seed;
    const uint32_t thr = (uint32_t)fminf(p * 4294967296.f, 4294967295.f);
    const float scale = 1.f / (1.f - p);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        float a[8];
        unpack8(ld_nc_v4(x + i), a);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = mix32((uint64_t)seed ^ (uint64_t)(i * 8 + j) * 0xD6E8FEB86659FD93ULL) >= thr ? a[j] * scale : 0.f;
        st_v4(y + i, pack8(a));
    }
}
__global__ void dropout_bwd_kernel(const uint4* __restrict__ dy, uint4* __restrict__ dx, long long nvec, float p, const long long* __restrict__ used_seed) {
    const long long seed = used_seed[0];
    const uint32_t thr = (uint32_t)fminf(p * 4294967296.f, 4294967295.f);
    const float scale = 1.f / (1.f - p);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        float a[8];
        unpack8(ld_nc_v4(dy + i), a);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = mix32((uint64_t)seed ^ (uint64_t)(i * 8 + j) * 0xD6E8FEB86659FD93ULL) >= thr ? a[j] * scale : 0.f;
        st_v4(dx + i, pack8(a));
    }
}
std::vector<at::Tensor> dropout_forward(at::Tensor x, double p, at::Tensor seed_state, int64_t salt) {
    TORCH_CHECK(x.is_cuda() && x.scalar_type() == at::kBFloat16 && x.is_contiguous() && x.numel() % 8 == 0, "dropout: contiguous bf16 tensor with numel % 8 == 0");
    TORCH_CHECK(seed_state.is_cuda() && seed_state.scalar_type() == at::kLong && seed_state.numel() >= 1);
    at::Tensor y = at::empty_like(x), used = at::empty({1}, seed_state.options());
    dropout_kernel<<<ew_grid(x.numel() / 8), 256, 0, at::cuda::getCurrentCUDAStream()>>>((const uint4*)x.data_ptr(), (uint4*)y.data_ptr(), x.numel() / 8, (float)p,
                                                                                       (const long long*)seed_state.data_ptr(), salt, (long long*)used.data_ptr());
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    return {y, used};
}
at::Tensor dropout_backward(at::Tensor dy, double p, at::Tensor used_seed) {
    TORCH_CHECK(dy.is_cuda() && dy.scalar_type() == at::kBFloat16 && dy.is_contiguous() && dy.numel() % 8 == 0);
    at::Tensor dx = at::empty_like(dy);
    dropout_bwd_kernel<<<ew_grid(dy.numel() / 8), 256, 0, at::cuda::getCurrentCUDAStream()>>>((const uint4*)dy.data_ptr(), (uint4*)dx.data_ptr(), dy.numel() / 8, (float)p,
                                                                                           (const long long*)used_seed.data_ptr());
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    return dx;
}
__global__ void bump_word_kernel(long long* p) { p[0] += 1; }
void bump_word(at::Tensor t) {
    TORCH_CHECK(t.is_cuda() && t.scalar_type() == at::kLong && t.numel() >= 1);
    bump_word_kernel<<<1, 1, 0, at::cuda::getCurrentCUDAStream()>>>((long long*)t.data_ptr());
    C10_CUDA_KERNEL_LAUNCH_CHECK();
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

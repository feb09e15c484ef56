// BatchNorm2d (training + inference) for NHWC bf16 activations, fused with the residual add and ReLU that follow it
// in every ResNet/VGG block of the reference models (/root/reference/src/model_ops/resnet.py:32-35,59-64, vgg.py:55).
//
// On B200 these layers are pure HBM streaming: stock ATen spends 7.6 ms of a 13.4 ms ResNet-18/b1024 step in its BN
// kernels (profiles/probe0_stock_torch.md) at ~0.5 TB/s.  Here every pass is a 16-byte-vector grid-stride stream:
//   forward : stats (1 read)  -> finalize (tiny)  -> apply+residual+ReLU (1-2 reads, 1 write)
//   backward: reduce (dz, z, x reads) -> finalize (tiny) -> dx (+ d_residual) (3 reads, 1-2 writes)
// Per-channel statistics use a shifted sum (shift = first row) so E[x^2]-E[x]^2 does not cancel; partial sums are
// written per CTA and merged by the finalize kernel in a fixed order (deterministic, no atomics).
//
// Thread mapping: a row of C channels is LC = C/8 16-byte lanes; LC is a power of two <= 256, so with 256-thread
// CTAs and strides that are multiples of 256 every thread keeps ONE channel lane for its whole life and holds that
// lane's 8 per-channel coefficients in registers.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include "common.cuh"
#include "pdl.cuh"

namespace psb {

constexpr int kBnThreads = 256;


// ---- fused "last CTA finalizes" variants -----------------------------------------------------------------------------
// The stand-alone finalize kernels above cost one more launch + drain (~3-5 us inside a CUDA graph) per BN pass; with 40
// passes per ResNet-18 step that is ~0.15 ms.  In the fused variants every CTA adds its per-channel partial sums to a
// 2*C-float accumulator with fp32 atomics (RED), takes a ticket, and the last CTA to arrive finalizes from the accumulator and
// leaves accumulator + ticket zeroed for the next call (self-cleaning persistent workspace owned by the layer).
// Summation order across CTAs is not fixed, so results can differ in the last bit between runs;
// PSB200_BN_DETERMINISTIC=1 selects the two-kernel, fixed-order path.
struct BnFwdFin {
    const __nv_bfloat16* x_row0; const void* gamma; const void* beta; float* running_mean; float* running_var;
    float momentum, eps; float *mean, *invstd, *coef_a, *coef_b; int p_bf16;
};
struct BnBwdFin {
    const void* gamma; const float* invstd; void* dgamma; void* dbeta; float *c1, *c2, *c3; int p_bf16;
};

__device__ __forceinline__ bool last_cta_ticket(float* ws, int C) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(reinterpret_cast<unsigned*>(ws + 2 * C), 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) __threadfence();
    return s_last != 0;
}

__device__ __forceinline__ void fwd_finalize_from_acc(float* ws, int C, long long M, const BnFwdFin& f) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float s = __ldcg(ws + c), q = __ldcg(ws + C + c);
        ws[c] = 0.f; ws[C + c] = 0.f;
        const float inv_m = 1.f / (float)M;
        const float dm = s * inv_m;
        const float mu = __bfloat162float(f.x_row0[c]) + dm;
        const float var = fmaxf(q * inv_m - dm * dm, 0.f);
        const float is = rsqrtf(var + f.eps);
        f.mean[c] = mu; f.invstd[c] = is;
        float ga = 1.f, be = 0.f;
        if (f.gamma) ga = f.p_bf16 ? __bfloat162float(((const __nv_bfloat16*)f.gamma)[c]) : ((const float*)f.gamma)[c];
        if (f.beta) be = f.p_bf16 ? __bfloat162float(((const __nv_bfloat16*)f.beta)[c]) : ((const float*)f.beta)[c];
        f.coef_a[c] = ga * is;
        f.coef_b[c] = be - mu * ga * is;
        if (f.running_mean) {
            const float unbiased = M > 1 ? var * (float)M / (float)(M - 1) : var;
            f.running_mean[c] = (1.f - f.momentum) * f.running_mean[c] + f.momentum * mu;
            f.running_var[c] = (1.f - f.momentum) * f.running_var[c] + f.momentum * unbiased;
        }
    }
    if (threadIdx.x == 0) *reinterpret_cast<unsigned*>(ws + 2 * C) = 0u;
}

__device__ __forceinline__ void bwd_finalize_from_acc(float* ws, int C, long long M, const BnBwdFin& f) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float s = __ldcg(ws + c), q = __ldcg(ws + C + c);
        ws[c] = 0.f; ws[C + c] = 0.f;
        const float is = f.invstd[c];
        float ga = 1.f;
        if (f.gamma) ga = f.p_bf16 ? __bfloat162float(((const __nv_bfloat16*)f.gamma)[c]) : ((const float*)f.gamma)[c];
        if (f.dgamma) { if (f.p_bf16) ((__nv_bfloat16*)f.dgamma)[c] = __float2bfloat16_rn(q * is); else ((float*)f.dgamma)[c] = q * is; }
        if (f.dbeta) { if (f.p_bf16) ((__nv_bfloat16*)f.dbeta)[c] = __float2bfloat16_rn(s); else ((float*)f.dbeta)[c] = s; }
        const float inv_m = 1.f / (float)M;
        f.c1[c] = ga * is; f.c2[c] = s * inv_m; f.c3[c] = q * is * is * inv_m;
    }
    if (threadIdx.x == 0) *reinterpret_cast<unsigned*>(ws + 2 * C) = 0u;
}

// ---- forward: partial statistics --------------------------------------------------------------------------------------
// partial[blk][0][C] = sum(x - K), partial[blk][1][C] = sum((x-K)^2), K = x[row 0]
__global__ void __launch_bounds__(kBnThreads) bn_stats_kernel(const uint4* __restrict__ x, long long M, int LC, float* __restrict__ partial,
                                                              float* ws, BnFwdFin fin) {
    extern __shared__ float sm[];                       // [kBnThreads][16]
    const int lane = threadIdx.x % LC;
    const int ty = threadIdx.x / LC;
    const int RP = kBnThreads / LC;                     // rows per CTA pass
    float K[8], s[8], q[8];
    unpack8(ld_nc_v4(x + lane), K);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
    const long long rstride = (long long)gridDim.x * RP;
    long long r = (long long)blockIdx.x * RP + ty;
    // 4 independent 16-byte loads in flight per thread
    for (; r + 3 * rstride < M; r += 4 * rstride) {
        uint4 v0 = ld_nc_v4(x + r * LC + lane), v1 = ld_nc_v4(x + (r + rstride) * LC + lane);
        uint4 v2 = ld_nc_v4(x + (r + 2 * rstride) * LC + lane), v3 = ld_nc_v4(x + (r + 3 * rstride) * LC + lane);
        float f[8];
#define ACC(v) unpack8(v, f); _Pragma("unroll") for (int j = 0; j < 8; ++j) { float d = f[j] - K[j]; s[j] += d; q[j] = fmaf(d, d, q[j]); }
        ACC(v0) ACC(v1) ACC(v2) ACC(v3)
    }
    for (; r < M; r += rstride) { uint4 v = ld_nc_v4(x + r * LC + lane); float f[8]; ACC(v) }
#undef ACC
    float* my = sm + threadIdx.x * 16;
#pragma unroll
    for (int j = 0; j < 8; ++j) { my[j] = s[j]; my[8 + j] = q[j]; }
    __syncthreads();
    if (ty == 0) {
        for (int t = 1; t < RP; ++t) {
            const float* o = sm + (t * LC + lane) * 16;
#pragma unroll
            for (int j = 0; j < 8; ++j) { s[j] += o[j]; q[j] += o[8 + j]; }
        }
        const int C = LC * 8;
        if (ws) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { atomicAdd(ws + lane * 8 + j, s[j]); atomicAdd(ws + C + lane * 8 + j, q[j]); }
        } else {
            float* ps = partial + (long long)blockIdx.x * 2 * C + lane * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) { ps[j] = s[j]; ps[C + j] = q[j]; }
        }
    }
    if (ws && last_cta_ticket(ws, LC * 8)) fwd_finalize_from_acc(ws, LC * 8, M, fin);
}

// one CTA per 32 channels, 8 partial-groups x 32 channels
template <typename PT>
__global__ void __launch_bounds__(256) bn_finalize_kernel(const float* __restrict__ partial, int nblk, int C, long long M,
                                                          const __nv_bfloat16* __restrict__ x_row0, const PT* __restrict__ gamma,
                                                          const PT* __restrict__ beta, float* running_mean, float* running_var,
                                                          float momentum, float eps, float* __restrict__ mean, float* __restrict__ invstd,
                                                          float* __restrict__ coef_a, float* __restrict__ coef_b) {
    __shared__ float ss[8][32], sq[8][32];
    const int cl = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    float s = 0.f, q = 0.f;
    if (c < C) {
        // the partial rows are independent: keep 8 loads in flight instead of one dependent load per iteration
        // (a serial loop here cost ~32 us per call: 74 dependent L2/DRAM round trips)
        float s8[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        int b = grp;
        for (; b + 56 < nblk; b += 64) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                s8[u] += __ldg(partial + (long long)(b + 8 * u) * 2 * C + c);
                q8[u] += __ldg(partial + (long long)(b + 8 * u) * 2 * C + C + c);
            }
        }
        for (; b < nblk; b += 8) { s8[0] += __ldg(partial + (long long)b * 2 * C + c); q8[0] += __ldg(partial + (long long)b * 2 * C + C + c); }
#pragma unroll
        for (int u = 0; u < 8; ++u) { s += s8[u]; q += q8[u]; }
    }
    ss[grp][cl] = s; sq[grp][cl] = q;
    __syncthreads();
    if (grp == 0 && c < C) {
        for (int g = 1; g < 8; ++g) { s += ss[g][cl]; q += sq[g][cl]; }
        const float inv_m = 1.f / (float)M;
        const float dm = s * inv_m;
        const float K = __bfloat162float(x_row0[c]);
        const float mu = K + dm;
        const float var = fmaxf(q * inv_m - dm * dm, 0.f);
        const float is = rsqrtf(var + eps);
        mean[c] = mu; invstd[c] = is;
        const float ga = gamma ? to_f(gamma[c]) : 1.f, be = beta ? to_f(beta[c]) : 0.f;
        coef_a[c] = ga * is;
        coef_b[c] = be - mu * ga * is;
        if (running_mean) {
            const float unbiased = M > 1 ? var * (float)M / (float)(M - 1) : var;
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mu;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
        }
    }
}

// inference: coefficients from running statistics
template <typename PT>
__global__ void bn_eval_coef_kernel(int C, const PT* gamma, const PT* beta, const float* rm, const float* rv, float eps, float* a, float* b) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float is = rsqrtf(rv[c] + eps);
    const float ga = gamma ? to_f(gamma[c]) : 1.f, be = beta ? to_f(beta[c]) : 0.f;
    a[c] = ga * is; b[c] = be - rm[c] * ga * is;
}

// ---- forward: z = relu(x*a + b + res) -----------------------------------------------------------------------------------
template <bool RELU, bool RES>
__global__ void __launch_bounds__(kBnThreads) bn_apply_kernel(const uint4* __restrict__ x, const uint4* __restrict__ res, uint4* __restrict__ z,
                                                              const float* __restrict__ coef_a, const float* __restrict__ coef_b, long long nvec, int LC) {
    pdl_trigger(); pdl_wait();      // pdl.cuh: the successor may become resident now; nothing of a predecessor is touched before this line
    const int lane = threadIdx.x % LC;
    float a[8], b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { a[j] = coef_a[lane * 8 + j]; b[j] = coef_b[lane * 8 + j]; }
    const long long stride = (long long)gridDim.x * kBnThreads;
    for (long long i = (long long)blockIdx.x * kBnThreads + threadIdx.x; i < nvec; i += 2 * stride) {
        const bool two = i + stride < nvec;
        uint4 v0 = ld_nc_v4(x + i), v1 = two ? ld_nc_v4(x + i + stride) : v0;
        uint4 r0, r1;
        if (RES) { r0 = ld_nc_v4(res + i); r1 = two ? ld_nc_v4(res + i + stride) : r0; }
        float f[8], g[8];
        unpack8(v0, f);
        if (RES) unpack8(r0, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) { float y = fmaf(f[j], a[j], b[j]); if (RES) y += g[j]; f[j] = RELU ? fmaxf(y, 0.f) : y; }
        st_v4(z + i, pack8(f));
        if (two) {
            unpack8(v1, f);
            if (RES) unpack8(r1, g);
#pragma unroll
            for (int j = 0; j < 8; ++j) { float y = fmaf(f[j], a[j], b[j]); if (RES) y += g[j]; f[j] = RELU ? fmaxf(y, 0.f) : y; }
            st_v4(z + i + stride, pack8(f));
        }
    }
}

// ---- backward: partial reductions ------------------------------------------------------------------------------------------
// dy = dz * (z > 0) if RELU else dz;  partial[blk][0][C] = sum(dy), partial[blk][1][C] = sum(dy * (x - mean))
// MASKX: the ReLU mask is recomputed from x (z > 0  <=>  a*x + b > 0 when there is no residual input), so the output tensor z
// is not read at all: one of the three streamed tensors disappears from both backward passes of every BN+ReLU layer.
template <bool RELU, bool MASKX>
__global__ void __launch_bounds__(kBnThreads) bn_bwd_reduce_kernel(const uint4* __restrict__ dz, const uint4* __restrict__ z, const uint4* __restrict__ x,
                                                                   const float* __restrict__ mean, long long M, int LC, float* __restrict__ partial,
                                                                   float* ws, BnBwdFin fin, const float* __restrict__ coef_a, const float* __restrict__ coef_b) {
    pdl_trigger(); pdl_wait();      // pdl.cuh: the successor may become resident now; nothing of a predecessor is touched before this line
    extern __shared__ float sm[];
    const int lane = threadIdx.x % LC, ty = threadIdx.x / LC, RP = kBnThreads / LC;
    float mu[8], s[8], q[8], ca[8], cb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { mu[j] = mean[lane * 8 + j]; s[j] = 0.f; q[j] = 0.f; ca[j] = MASKX ? coef_a[lane * 8 + j] : 0.f; cb[j] = MASKX ? coef_b[lane * 8 + j] : 0.f; }
    const long long rstride = (long long)gridDim.x * RP;
    long long r = (long long)blockIdx.x * RP + ty;
#define ACCB(idx) { float d[8], xx[8], zz[8]; unpack8(vd##idx, d); unpack8(vx##idx, xx); if (RELU && !MASKX) unpack8(vz##idx, zz); \
        _Pragma("unroll") for (int j = 0; j < 8; ++j) { const float pre = MASKX ? fmaf(xx[j], ca[j], cb[j]) : zz[j]; \
            float dy = (RELU && !(pre > 0.f)) ? 0.f : d[j]; s[j] += dy; q[j] = fmaf(dy, xx[j] - mu[j], q[j]); } }
    for (; r + rstride < M; r += 2 * rstride) {
        const long long i0 = r * LC + lane, i1 = (r + rstride) * LC + lane;
        uint4 vd0 = ld_nc_v4(dz + i0), vx0 = ld_nc_v4(x + i0), vd1 = ld_nc_v4(dz + i1), vx1 = ld_nc_v4(x + i1), vz0, vz1;
        if (RELU && !MASKX) { vz0 = ld_nc_v4(z + i0); vz1 = ld_nc_v4(z + i1); }
        ACCB(0) ACCB(1)
    }
    for (; r < M; r += rstride) {
        const long long i0 = r * LC + lane;
        uint4 vd0 = ld_nc_v4(dz + i0), vx0 = ld_nc_v4(x + i0), vz0;
        if (RELU && !MASKX) vz0 = ld_nc_v4(z + i0);
        ACCB(0)
    }
#undef ACCB
    float* my = sm + threadIdx.x * 16;
#pragma unroll
    for (int j = 0; j < 8; ++j) { my[j] = s[j]; my[8 + j] = q[j]; }
    __syncthreads();
    if (ty == 0) {
        for (int t = 1; t < RP; ++t) {
            const float* o = sm + (t * LC + lane) * 16;
#pragma unroll
            for (int j = 0; j < 8; ++j) { s[j] += o[j]; q[j] += o[8 + j]; }
        }
        const int C = LC * 8;
        if (ws) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { atomicAdd(ws + lane * 8 + j, s[j]); atomicAdd(ws + C + lane * 8 + j, q[j]); }
        } else {
            float* ps = partial + (long long)blockIdx.x * 2 * C + lane * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) { ps[j] = s[j]; ps[C + j] = q[j]; }
        }
    }
    if (ws && last_cta_ticket(ws, LC * 8)) bwd_finalize_from_acc(ws, LC * 8, M, fin);
}

// dgamma = sum(dy*xhat) = invstd*sum(dy*(x-mean)); dbeta = sum(dy);
// dx = c1*(dy - c2 - (x-mean)*c3), c1 = gamma*invstd, c2 = sum_dy/M, c3 = invstd^2*sum(dy*(x-mean))/M
template <typename PT>
__global__ void __launch_bounds__(256) bn_bwd_finalize_kernel(const float* __restrict__ partial, int nblk, int C, long long M,
                                                              const PT* __restrict__ gamma, const float* __restrict__ invstd,
                                                              PT* __restrict__ dgamma, PT* __restrict__ dbeta, float* __restrict__ c1,
                                                              float* __restrict__ c2, float* __restrict__ c3) {
    __shared__ float ss[8][32], sq[8][32];
    const int cl = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    float s = 0.f, q = 0.f;
    if (c < C) {
        // the partial rows are independent: keep 8 loads in flight instead of one dependent load per iteration
        // (a serial loop here cost ~32 us per call: 74 dependent L2/DRAM round trips)
        float s8[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        int b = grp;
        for (; b + 56 < nblk; b += 64) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                s8[u] += __ldg(partial + (long long)(b + 8 * u) * 2 * C + c);
                q8[u] += __ldg(partial + (long long)(b + 8 * u) * 2 * C + C + c);
            }
        }
        for (; b < nblk; b += 8) { s8[0] += __ldg(partial + (long long)b * 2 * C + c); q8[0] += __ldg(partial + (long long)b * 2 * C + C + c); }
#pragma unroll
        for (int u = 0; u < 8; ++u) { s += s8[u]; q += q8[u]; }
    }
    ss[grp][cl] = s; sq[grp][cl] = q;
    __syncthreads();
    if (grp == 0 && c < C) {
        for (int g = 1; g < 8; ++g) { s += ss[g][cl]; q += sq[g][cl]; }
        const float is = invstd[c];
        const float ga = gamma ? to_f(gamma[c]) : 1.f;
        if (dgamma) dgamma[c] = from_f<PT>(q * is);
        if (dbeta) dbeta[c] = from_f<PT>(s);
        const float inv_m = 1.f / (float)M;
        c1[c] = ga * is; c2[c] = s * inv_m; c3[c] = q * is * is * inv_m;
    }
}

// ---- wide finalize (default; bn_set_finalize_wide(0) / PSB200_BN_FINALIZE=narrow selects the old kernels) -------------------------------------------
// profiles/run23_launch_list_final.md: the finalize kernels above cost 26-35 us each under ncu (40 launches per ResNet-18 step) because 2-16 CTAs
// walk ~74 partial rows per thread in dependent batches.  Here one CTA owns 8 channels (one 32-byte sector of a partial row): thread
// (g = tid / 8, cc = tid % 8) sums rows g, g+32, g+64, ... with eight predicated loads in flight per batch (3 batches for 592 rows),
// the 32 row groups are folded through shared memory and threads 0..7 finish their channel with the same arithmetic as above.
// Sums s (first half of a partial row) and q (second half) for channel c over all nblk rows; valid in threads with g == 0 after the call.
__device__ __forceinline__ void wide_partial_sum(const float* partial, int nblk, int C, int c, float& s, float& q, bool clear = false) {
    __shared__ float ss[32][8], sq[32][8];
    const int cc = threadIdx.x & 7, g = threadIdx.x >> 3;
    float s8[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (c < C) {
        for (int b = g; b < nblk; b += 8 * 32) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int bb = b + 32 * u;
                if (bb < nblk) {
                    float* ps = const_cast<float*>(partial) + (long long)bb * 2 * C + c;
                    if (clear) {                 // rows written by a convolution epilogue: read once, leave zeroed for the next step
                        s8[u] += __ldcg(ps); q8[u] += __ldcg(ps + C);
                        ps[0] = 0.f; ps[C] = 0.f;
                    } else {
                        s8[u] += __ldg(ps); q8[u] += __ldg(ps + C);
                    }
                }
            }
        }
    }
    s = 0.f; q = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) { s += s8[u]; q += q8[u]; }
    ss[g][cc] = s; sq[g][cc] = q;
    __syncthreads();
    if (g == 0) {
        for (int t = 1; t < 32; ++t) { s += ss[t][cc]; q += sq[t][cc]; }
    }
}

template <typename PT>
__global__ void __launch_bounds__(256) bn_finalize_wide_kernel(const float* partial, int nblk, int C, long long M,
                                                               const __nv_bfloat16* __restrict__ x_row0, const PT* __restrict__ gamma,
                                                               const PT* __restrict__ beta, float* running_mean, float* running_var,
                                                               float momentum, float eps, float* __restrict__ mean, float* __restrict__ invstd,
                                                               float* __restrict__ coef_a, float* __restrict__ coef_b, int clear) {
    pdl_trigger(); pdl_wait();      // pdl.cuh: the successor may become resident now; nothing of a predecessor is touched before this line
    const int c = blockIdx.x * 8 + (threadIdx.x & 7);
    float s, q;
    wide_partial_sum(partial, nblk, C, c, s, q, clear != 0);
    if ((threadIdx.x >> 3) == 0 && c < C) {
        const float inv_m = 1.f / (float)M;
        const float dm = s * inv_m;
        const float K = x_row0 ? __bfloat162float(x_row0[c]) : 0.f;      // conv-epilogue statistics are unshifted sums
        const float mu = K + dm;
        const float var = fmaxf(q * inv_m - dm * dm, 0.f);
        const float is = rsqrtf(var + eps);
        mean[c] = mu; invstd[c] = is;
        const float ga = gamma ? to_f(gamma[c]) : 1.f, be = beta ? to_f(beta[c]) : 0.f;
        coef_a[c] = ga * is;
        coef_b[c] = be - mu * ga * is;
        if (running_mean) {
            const float unbiased = M > 1 ? var * (float)M / (float)(M - 1) : var;
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mu;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
        }
    }
}

template <typename PT>
__global__ void __launch_bounds__(256) bn_bwd_finalize_wide_kernel(const float* __restrict__ partial, int nblk, int C, long long M,
                                                                   const PT* __restrict__ gamma, const float* __restrict__ invstd,
                                                                   PT* __restrict__ dgamma, PT* __restrict__ dbeta, float* __restrict__ c1,
                                                                   float* __restrict__ c2, float* __restrict__ c3) {
    pdl_trigger(); pdl_wait();      // pdl.cuh: the successor may become resident now; nothing of a predecessor is touched before this line
    const int c = blockIdx.x * 8 + (threadIdx.x & 7);
    float s, q;
    wide_partial_sum(partial, nblk, C, c, s, q);
    if ((threadIdx.x >> 3) == 0 && c < C) {
        const float is = invstd[c];
        const float ga = gamma ? to_f(gamma[c]) : 1.f;
        if (dgamma) dgamma[c] = from_f<PT>(q * is);
        if (dbeta) dbeta[c] = from_f<PT>(s);
        const float inv_m = 1.f / (float)M;
        c1[c] = ga * is; c2[c] = s * inv_m; c3[c] = q * is * is * inv_m;
    }
}

static int g_bn_finalize_wide = 1;      // default since round 2 (validated on B200: 6.96 vs 7.23 ms/step, profiles/run24_*.md)
void bn_set_finalize_wide(int on) { g_bn_finalize_wide = on ? 1 : 0; }

template <bool RELU, bool RES, bool MASKX>
__global__ void __launch_bounds__(kBnThreads) bn_bwd_apply_kernel(const uint4* __restrict__ dz, const uint4* __restrict__ z, const uint4* __restrict__ x,
                                                                  const float* __restrict__ mean, const float* __restrict__ c1, const float* __restrict__ c2,
                                                                  const float* __restrict__ c3, uint4* __restrict__ dx, uint4* __restrict__ dres,
                                                                  long long nvec, int LC, const float* __restrict__ coef_a, const float* __restrict__ coef_b) {
    pdl_trigger(); pdl_wait();      // pdl.cuh: the successor may become resident now; nothing of a predecessor is touched before this line
    const int lane = threadIdx.x % LC;
    float mu[8], k1[8], k2[8], k3[8], ca[8], cb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { mu[j] = mean[lane * 8 + j]; k1[j] = c1[lane * 8 + j]; k2[j] = c2[lane * 8 + j]; k3[j] = c3[lane * 8 + j];
                                  ca[j] = MASKX ? coef_a[lane * 8 + j] : 0.f; cb[j] = MASKX ? coef_b[lane * 8 + j] : 0.f; }
    const long long stride = (long long)gridDim.x * kBnThreads;
    for (long long i = (long long)blockIdx.x * kBnThreads + threadIdx.x; i < nvec; i += 2 * stride) {
        const bool two = i + stride < nvec;
        const long long i1 = two ? i + stride : i;
        uint4 vd0 = ld_nc_v4(dz + i), vx0 = ld_nc_v4(x + i), vd1 = ld_nc_v4(dz + i1), vx1 = ld_nc_v4(x + i1), vz0, vz1;
        if (RELU && !MASKX) { vz0 = ld_nc_v4(z + i); vz1 = ld_nc_v4(z + i1); }
#define APP(idx, off) { float d[8], xx[8], zz[8], o[8]; unpack8(vd##idx, d); unpack8(vx##idx, xx); if (RELU && !MASKX) unpack8(vz##idx, zz); \
        _Pragma("unroll") for (int j = 0; j < 8; ++j) { const float pre = MASKX ? fmaf(xx[j], ca[j], cb[j]) : zz[j]; \
            float dy = (RELU && !(pre > 0.f)) ? 0.f : d[j]; d[j] = dy; \
            o[j] = k1[j] * (dy - k2[j] - (xx[j] - mu[j]) * k3[j]); } \
        st_v4(dx + off, pack8(o)); if (RES && RELU) st_v4(dres + off, pack8(d)); }
        APP(0, i)
        if (two) APP(1, i1)
#undef APP
    }
}

// ---- host ---------------------------------------------------------------------------------------------------------------------
static int stream_grid(long long work_items, int per_block, int ctas_per_sm = 4) {
    long long blocks = (work_items + per_block - 1) / per_block;
    return (int)std::max<long long>(1, std::min<long long>(blocks, (long long)kNumSMs * ctas_per_sm));
}

static void check_act(const at::Tensor& t, const char* name) {
    TORCH_CHECK(t.is_cuda() && t.scalar_type() == at::kBFloat16, name, ": bf16 CUDA tensor expected");
    TORCH_CHECK(t.dim() == 4 && t.is_contiguous(at::MemoryFormat::ChannelsLast), name, ": NHWC (channels_last) layout expected");
}

// returns {z, mean, invstd}
std::vector<at::Tensor> bn_forward(at::Tensor x, c10::optional<at::Tensor> gamma, c10::optional<at::Tensor> beta,
                                   c10::optional<at::Tensor> running_mean, c10::optional<at::Tensor> running_var,
                                   c10::optional<at::Tensor> residual, bool training, double momentum, double eps, bool relu,
                                   c10::optional<at::Tensor> ws, c10::optional<at::Tensor> conv_partial) {
    check_act(x, "x");
    const int C = (int)x.size(1);
    const long long M = x.numel() / C;
    const int LC = C / 8;
    TORCH_CHECK(C % 8 == 0 && LC >= 1 && LC <= 256 && (LC & (LC - 1)) == 0, "bn: C/8 must be a power of two <= 256");
    if (residual.has_value()) { check_act(*residual, "residual"); TORCH_CHECK(residual->sizes() == x.sizes()); }
    auto stream = at::cuda::getCurrentCUDAStream();
    auto fopt = x.options().dtype(at::kFloat);
    at::Tensor z = at::empty_like(x);
    at::Tensor mean = at::empty({C}, fopt), invstd = at::empty({C}, fopt), ca = at::empty({C}, fopt), cb = at::empty({C}, fopt);
    const bool p_bf16 = gamma.has_value() && gamma->scalar_type() == at::kBFloat16;
    if (gamma.has_value()) TORCH_CHECK(p_bf16 || gamma->scalar_type() == at::kFloat);
    const void* gp = gamma.has_value() ? gamma->data_ptr() : nullptr;
    const void* bp = beta.has_value() ? beta->data_ptr() : nullptr;
    float* rm = running_mean.has_value() ? running_mean->data_ptr<float>() : nullptr;
    float* rv = running_var.has_value() ? running_var->data_ptr<float>() : nullptr;
    const uint4* xp = reinterpret_cast<const uint4*>(x.data_ptr());
    if (training && conv_partial.has_value()) {
        // statistics were accumulated by the producing convolution's epilogue (rows [r][2][C] of plain sums, zero where unused)
        TORCH_CHECK(conv_partial->is_cuda() && conv_partial->scalar_type() == at::kFloat && conv_partial->is_contiguous() && conv_partial->numel() % (2 * C) == 0,
                    "conv_partial must be a contiguous fp32 [rows, 2, C] tensor");
        const int rows = (int)(conv_partial->numel() / (2 * C));
        if (p_bf16)
            launch_pdl(bn_finalize_wide_kernel<__nv_bfloat16>, dim3((C + 7) / 8), dim3(256), 0, stream, conv_partial->data_ptr<float>(), rows, C, M, nullptr,
                (const __nv_bfloat16*)gp, (const __nv_bfloat16*)bp, rm, rv, (float)momentum, (float)eps, mean.data_ptr<float>(), invstd.data_ptr<float>(),
                ca.data_ptr<float>(), cb.data_ptr<float>(), 1);
        else
            launch_pdl(bn_finalize_wide_kernel<float>, dim3((C + 7) / 8), dim3(256), 0, stream, conv_partial->data_ptr<float>(), rows, C, M, nullptr,
                (const float*)gp, (const float*)bp, rm, rv, (float)momentum, (float)eps, mean.data_ptr<float>(), invstd.data_ptr<float>(),
                ca.data_ptr<float>(), cb.data_ptr<float>(), 1);
        C10_CUDA_KERNEL_LAUNCH_CHECK();
    } else if (training) {
        const int RP = kBnThreads / LC;
        const int nblk = stream_grid(M, RP * 8);
        if (ws.has_value()) {
            TORCH_CHECK(ws->scalar_type() == at::kFloat && ws->numel() >= 2 * C + 1 && ws->is_cuda(), "bn workspace must hold 2*C+1 floats");
            BnFwdFin fin{(const __nv_bfloat16*)x.data_ptr(), gp, bp, rm, rv, (float)momentum, (float)eps, mean.data_ptr<float>(), invstd.data_ptr<float>(),
                         ca.data_ptr<float>(), cb.data_ptr<float>(), p_bf16 ? 1 : 0};
            bn_stats_kernel<<<nblk, kBnThreads, kBnThreads * 16 * sizeof(float), stream>>>(xp, M, LC, nullptr, ws->data_ptr<float>(), fin);
            C10_CUDA_KERNEL_LAUNCH_CHECK();
        } else {
        at::Tensor partial = at::empty({nblk, 2, C}, fopt);
        bn_stats_kernel<<<nblk, kBnThreads, kBnThreads * 16 * sizeof(float), stream>>>(xp, M, LC, partial.data_ptr<float>(), nullptr, BnFwdFin{});
        C10_CUDA_KERNEL_LAUNCH_CHECK();
        const int fb = (C + 31) / 32;
        if (g_bn_finalize_wide && p_bf16)
            launch_pdl(bn_finalize_wide_kernel<__nv_bfloat16>, dim3((C + 7) / 8), dim3(256), 0, stream, partial.data_ptr<float>(), nblk, C, M, (const __nv_bfloat16*)x.data_ptr(),
                (const __nv_bfloat16*)gp, (const __nv_bfloat16*)bp, rm, rv, (float)momentum, (float)eps, mean.data_ptr<float>(), invstd.data_ptr<float>(),
                ca.data_ptr<float>(), cb.data_ptr<float>(), 0);
        else if (g_bn_finalize_wide)
            launch_pdl(bn_finalize_wide_kernel<float>, dim3((C + 7) / 8), dim3(256), 0, stream, partial.data_ptr<float>(), nblk, C, M, (const __nv_bfloat16*)x.data_ptr(),
                (const float*)gp, (const float*)bp, rm, rv, (float)momentum, (float)eps, mean.data_ptr<float>(), invstd.data_ptr<float>(),
                ca.data_ptr<float>(), cb.data_ptr<float>(), 0);
        else if (p_bf16)
            bn_finalize_kernel<__nv_bfloat16><<<fb, 256, 0, stream>>>(partial.data_ptr<float>(), nblk, C, M, (const __nv_bfloat16*)x.data_ptr(),
                (const __nv_bfloat16*)gp, (const __nv_bfloat16*)bp, rm, rv, (float)momentum, (float)eps, mean.data_ptr<float>(), invstd.data_ptr<float>(),
                ca.data_ptr<float>(), cb.data_ptr<float>());
        else
            bn_finalize_kernel<float><<<fb, 256, 0, stream>>>(partial.data_ptr<float>(), nblk, C, M, (const __nv_bfloat16*)x.data_ptr(),
                (const float*)gp, (const float*)bp, rm, rv, (float)momentum, (float)eps, mean.data_ptr<float>(), invstd.data_ptr<float>(),
                ca.data_ptr<float>(), cb.data_ptr<float>());
        C10_CUDA_KERNEL_LAUNCH_CHECK();
        }
    } else {
        TORCH_CHECK(rm && rv, "bn eval needs running statistics");
        if (p_bf16) bn_eval_coef_kernel<__nv_bfloat16><<<(C + 127) / 128, 128, 0, stream>>>(C, (const __nv_bfloat16*)gp, (const __nv_bfloat16*)bp, rm, rv, (float)eps, ca.data_ptr<float>(), cb.data_ptr<float>());
        else bn_eval_coef_kernel<float><<<(C + 127) / 128, 128, 0, stream>>>(C, (const float*)gp, (const float*)bp, rm, rv, (float)eps, ca.data_ptr<float>(), cb.data_ptr<float>());
        C10_CUDA_KERNEL_LAUNCH_CHECK();
    }
    const long long nvec = x.numel() / 8;
    const int grid = stream_grid(nvec, kBnThreads * 2);
    const uint4* rp = residual.has_value() ? reinterpret_cast<const uint4*>(residual->data_ptr()) : nullptr;
    uint4* zp = reinterpret_cast<uint4*>(z.data_ptr());
    const float *pa = ca.data_ptr<float>(), *pb = cb.data_ptr<float>();
    if (relu && rp) launch_pdl(bn_apply_kernel<true, true>, dim3(grid), dim3(kBnThreads), 0, stream, xp, rp, zp, pa, pb, nvec, LC);
    else if (relu) launch_pdl(bn_apply_kernel<true, false>, dim3(grid), dim3(kBnThreads), 0, stream, xp, rp, zp, pa, pb, nvec, LC);
    else if (rp) launch_pdl(bn_apply_kernel<false, true>, dim3(grid), dim3(kBnThreads), 0, stream, xp, rp, zp, pa, pb, nvec, LC);
    else launch_pdl(bn_apply_kernel<false, false>, dim3(grid), dim3(kBnThreads), 0, stream, xp, rp, zp, pa, pb, nvec, LC);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    return {z, mean, invstd, ca, cb};
}

// returns {dx, dres (or undefined), dgamma, dbeta}
std::vector<at::Tensor> bn_backward(at::Tensor dz, c10::optional<at::Tensor> z_opt, at::Tensor x, c10::optional<at::Tensor> gamma, at::Tensor mean,
                                    at::Tensor invstd, bool relu, bool has_res, c10::optional<at::Tensor> ws,
                                    c10::optional<at::Tensor> coef_a, c10::optional<at::Tensor> coef_b) {
    check_act(dz, "dz"); check_act(x, "x");
    const bool maskx = relu && !has_res && coef_a.has_value() && coef_b.has_value();
    at::Tensor z = z_opt.has_value() ? *z_opt : x;
    if (relu && !maskx) { TORCH_CHECK(z_opt.has_value(), "bn backward with ReLU + residual needs the saved output"); check_act(z, "z"); }
    const float* pca = maskx ? coef_a->data_ptr<float>() : nullptr;
    const float* pcb = maskx ? coef_b->data_ptr<float>() : nullptr;
    const int C = (int)x.size(1);
    const long long M = x.numel() / C;
    const int LC = C / 8;
    auto stream = at::cuda::getCurrentCUDAStream();
    auto fopt = x.options().dtype(at::kFloat);
    const int RP = kBnThreads / LC;
    const int nblk = stream_grid(M, RP * 8);
    const uint4 *dzp = (const uint4*)dz.data_ptr(), *zp = (relu && !maskx) ? (const uint4*)z.data_ptr() : nullptr, *xp = (const uint4*)x.data_ptr();
    at::Tensor c1 = at::empty({C}, fopt), c2 = at::empty({C}, fopt), c3 = at::empty({C}, fopt);
    at::Tensor dgamma, dbeta;
    const bool p_bf16 = gamma.has_value() && gamma->scalar_type() == at::kBFloat16;
    if (p_bf16) { dgamma = at::empty({C}, x.options()); dbeta = at::empty({C}, x.options()); }
    else { dgamma = at::empty({C}, fopt); dbeta = at::empty({C}, fopt); }
    const size_t smem = kBnThreads * 16 * sizeof(float);
    if (ws.has_value()) {
        TORCH_CHECK(ws->scalar_type() == at::kFloat && ws->numel() >= 2 * C + 1 && ws->is_cuda(), "bn workspace must hold 2*C+1 floats");
        BnBwdFin fin{gamma.has_value() ? gamma->data_ptr() : nullptr, invstd.data_ptr<float>(), dgamma.data_ptr(), dbeta.data_ptr(),
                     c1.data_ptr<float>(), c2.data_ptr<float>(), c3.data_ptr<float>(), p_bf16 ? 1 : 0};
        if (maskx) launch_pdl(bn_bwd_reduce_kernel<true, true>, dim3(nblk), dim3(kBnThreads), smem, stream, dzp, zp, xp, mean.data_ptr<float>(), M, LC, nullptr, ws->data_ptr<float>(), fin, pca, pcb);
        else if (relu) launch_pdl(bn_bwd_reduce_kernel<true, false>, dim3(nblk), dim3(kBnThreads), smem, stream, dzp, zp, xp, mean.data_ptr<float>(), M, LC, nullptr, ws->data_ptr<float>(), fin, pca, pcb);
        else launch_pdl(bn_bwd_reduce_kernel<false, false>, dim3(nblk), dim3(kBnThreads), smem, stream, dzp, zp, xp, mean.data_ptr<float>(), M, LC, nullptr, ws->data_ptr<float>(), fin, pca, pcb);
        C10_CUDA_KERNEL_LAUNCH_CHECK();
    } else {
        at::Tensor partial = at::empty({nblk, 2, C}, fopt);
        if (maskx) launch_pdl(bn_bwd_reduce_kernel<true, true>, dim3(nblk), dim3(kBnThreads), smem, stream, dzp, zp, xp, mean.data_ptr<float>(), M, LC, partial.data_ptr<float>(), nullptr, BnBwdFin{}, pca, pcb);
        else if (relu) launch_pdl(bn_bwd_reduce_kernel<true, false>, dim3(nblk), dim3(kBnThreads), smem, stream, dzp, zp, xp, mean.data_ptr<float>(), M, LC, partial.data_ptr<float>(), nullptr, BnBwdFin{}, pca, pcb);
        else launch_pdl(bn_bwd_reduce_kernel<false, false>, dim3(nblk), dim3(kBnThreads), smem, stream, dzp, zp, xp, mean.data_ptr<float>(), M, LC, partial.data_ptr<float>(), nullptr, BnBwdFin{}, pca, pcb);
        C10_CUDA_KERNEL_LAUNCH_CHECK();
        const int fb = (C + 31) / 32;
        if (g_bn_finalize_wide && p_bf16)
            launch_pdl(bn_bwd_finalize_wide_kernel<__nv_bfloat16>, dim3((C + 7) / 8), dim3(256), 0, stream, partial.data_ptr<float>(), nblk, C, M, (const __nv_bfloat16*)gamma->data_ptr(), invstd.data_ptr<float>(),
                (__nv_bfloat16*)dgamma.data_ptr(), (__nv_bfloat16*)dbeta.data_ptr(), c1.data_ptr<float>(), c2.data_ptr<float>(), c3.data_ptr<float>());
        else if (g_bn_finalize_wide)
            launch_pdl(bn_bwd_finalize_wide_kernel<float>, dim3((C + 7) / 8), dim3(256), 0, stream, partial.data_ptr<float>(), nblk, C, M, gamma.has_value() ? gamma->data_ptr<float>() : nullptr, invstd.data_ptr<float>(),
                dgamma.data_ptr<float>(), dbeta.data_ptr<float>(), c1.data_ptr<float>(), c2.data_ptr<float>(), c3.data_ptr<float>());
        else if (p_bf16)
            bn_bwd_finalize_kernel<__nv_bfloat16><<<fb, 256, 0, stream>>>(partial.data_ptr<float>(), nblk, C, M, (const __nv_bfloat16*)gamma->data_ptr(), invstd.data_ptr<float>(),
                (__nv_bfloat16*)dgamma.data_ptr(), (__nv_bfloat16*)dbeta.data_ptr(), c1.data_ptr<float>(), c2.data_ptr<float>(), c3.data_ptr<float>());
        else
            bn_bwd_finalize_kernel<float><<<fb, 256, 0, stream>>>(partial.data_ptr<float>(), nblk, C, M, gamma.has_value() ? gamma->data_ptr<float>() : nullptr, invstd.data_ptr<float>(),
                dgamma.data_ptr<float>(), dbeta.data_ptr<float>(), c1.data_ptr<float>(), c2.data_ptr<float>(), c3.data_ptr<float>());
        C10_CUDA_KERNEL_LAUNCH_CHECK();
    }
    at::Tensor dx = at::empty_like(x);
    at::Tensor dres;
    if (has_res) dres = relu ? at::empty_like(x) : dz;       // without ReLU the residual gradient IS dz: no copy
    const long long nvec = x.numel() / 8;
    const int grid = stream_grid(nvec, kBnThreads * 2);
    uint4* dxp = (uint4*)dx.data_ptr();
    uint4* drp = (has_res && relu) ? (uint4*)dres.data_ptr() : nullptr;
    const float *mp = mean.data_ptr<float>(), *p1 = c1.data_ptr<float>(), *p2 = c2.data_ptr<float>(), *p3 = c3.data_ptr<float>();
    if (relu && has_res) launch_pdl(bn_bwd_apply_kernel<true, true, false>, dim3(grid), dim3(kBnThreads), 0, stream, dzp, zp, xp, mp, p1, p2, p3, dxp, drp, nvec, LC, pca, pcb);
    else if (maskx) launch_pdl(bn_bwd_apply_kernel<true, false, true>, dim3(grid), dim3(kBnThreads), 0, stream, dzp, zp, xp, mp, p1, p2, p3, dxp, drp, nvec, LC, pca, pcb);
    else if (relu) launch_pdl(bn_bwd_apply_kernel<true, false, false>, dim3(grid), dim3(kBnThreads), 0, stream, dzp, zp, xp, mp, p1, p2, p3, dxp, drp, nvec, LC, pca, pcb);
    else launch_pdl(bn_bwd_apply_kernel<false, false, false>, dim3(grid), dim3(kBnThreads), 0, stream, dzp, zp, xp, mp, p1, p2, p3, dxp, drp, nvec, LC, pca, pcb);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
    return {dx, dres, dgamma, dbeta};
}

}  // namespace psb

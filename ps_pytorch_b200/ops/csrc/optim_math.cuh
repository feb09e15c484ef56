// Per-element optimizer math shared by the stand-alone fused optimizer kernels (optim_kernels.cu) and the PS
// reduce+update+broadcast kernel (comm_kernels.cu).  Semantics: /root/reference/src/optim/sgd.py:59-92, adam.py:44-94.
#pragma once
#include "common.cuh"

namespace psb {

// Hyper-parameters live in device memory so a captured CUDA graph keeps working when the host changes them.
struct OptHP {
    float lr, momentum, dampening, weight_decay;   // SGD
    float beta1, beta2, eps, _pad;                  // Adam
    int nesterov, amsgrad, kind, _pad2;             // kind: 0 = SGD, 1 = Adam
    long long steps;                                // optimizer steps taken so far (device-side counter)
};

__device__ __forceinline__ float sgd_update(float w, float g, float& buf, const OptHP& hp, bool first) {
    float d = g;
    if (hp.weight_decay != 0.f) d = fmaf(hp.weight_decay, w, d);
    if (hp.momentum != 0.f) {
        buf = first ? d : fmaf(hp.momentum, buf, (1.f - hp.dampening) * d);
        d = hp.nesterov ? fmaf(hp.momentum, buf, d) : buf;
    }
    return fmaf(-hp.lr, d, w);
}

// step_size = lr * sqrt(1 - b2^t) / (1 - b1^t), computed once per thread
__device__ __forceinline__ float adam_step_size(const OptHP& hp, long long t) {
    float bc1 = 1.f - powf(hp.beta1, (float)t);
    float bc2 = 1.f - powf(hp.beta2, (float)t);
    return hp.lr * sqrtf(bc2) / bc1;
}

__device__ __forceinline__ float adam_update(float w, float g, float& m, float& v, float* vmax, const OptHP& hp, float step_size) {
    if (hp.weight_decay != 0.f) g = fmaf(hp.weight_decay, w, g);
    m = fmaf(hp.beta1, m, (1.f - hp.beta1) * g);
    v = fmaf(hp.beta2, v, (1.f - hp.beta2) * g * g);
    float vv = v;
    if (vmax) { *vmax = fmaxf(*vmax, v); vv = *vmax; }
    return w - step_size * (m / (sqrtf(vv) + hp.eps));
}

}  // namespace psb

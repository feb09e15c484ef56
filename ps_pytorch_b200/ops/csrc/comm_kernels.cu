// The two communication-bound hot paths of the parameter server as fused kernels over peer-mapped memory
// (NVLink 5 / NVSwitch), with no NCCL/MPI call on the path:
//
//   push_kernel           worker side.  gradient arena -> cast (bf16) or block-scaled fp8 quantise (32-element blocks,
//                         e8m0 scale) -> 16-byte stores straight into the PS's slot for this worker (a peer pointer)
//                         -> system-scope release of the arrival flag (value = step id).  Also zeroes the local
//                         gradient arena for the next backward.  Replaces D2H + numpy + Blosc + pickle + MPI isend
//                         (/root/reference/src/distributed_worker.py:254-272, compression.py:18-23).
//   reduce_update_kernel  PS side.  per bucket: acquire-poll the P arrival flags (first K of P win; stale step ids are
//                         ignored), load the K slots (local HBM in push mode, or peers' gradient arenas by P2P loads in
//                         pull mode), dequantise, fp32 sum, x 1/K, SGD-momentum/Nesterov/WD or Adam/AMSGrad on the fp32
//                         master + state, and store the fresh bf16 weights into every worker's replica (peer stores, or
//                         ONE multimem.st through the NVLS multicast mapping) followed by a release of the per-bucket
//                         "weights ready" flag (value = next step id, which doubles as the step message).  Replaces
//                         Waitany + decompress + numpy sum + /K + optimizer.step + per-layer Blosc bcast
//                         (/root/reference/src/sync_replicas_master_nn.py:157-225, optim/*.py).
//   wait_flag_kernel      worker side: acquire-spin until the local "weights ready" flag reaches this worker's step.
//
// All step counters live in device memory, so a whole worker step (wait -> fwd -> bwd -> push) and a whole PS step are
// CUDA-graph capturable.  Every spin has a wall-clock timeout (globaltimer) that raises an error word instead of
// hanging the GPU.
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include "optim_math.cuh"

namespace psb {

constexpr int kMaxPeers = 8;
constexpr int kMaxBuckets = 16;
enum WireMode { WIRE_BF16 = 0, WIRE_FP8 = 1, WIRE_FP32 = 2 };

// ---- block-scaled fp8 (e4m3 payload, one e8m0 scale per 32 elements) ---------------------------------------------
__device__ __forceinline__ int e8m0_exponent(float amax) {
    // smallest e with amax / 2^e <= 448 (e4m3 max finite)
    float s = amax * (1.0f / 448.0f);
    if (!(s > 0.f)) return -127;
    uint32_t u = __float_as_uint(s);
    int e = (int)((u >> 23) & 0xff) - 127 + ((u & 0x7fffffu) ? 1 : 0);
    return max(-126, min(127, e));
}
__device__ __forceinline__ float exp2i(int e) { return __uint_as_float((uint32_t)(e + 127) << 23); }   // e in [-126,127]

__device__ __forceinline__ uint32_t quant4(float a, float b, float c, float d, float inv) {
    __nv_fp8x2_storage_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a * inv, b * inv), __NV_SATFINITE, __NV_E4M3);
    __nv_fp8x2_storage_t hi = __nv_cvt_float2_to_fp8x2(make_float2(c * inv, d * inv), __NV_SATFINITE, __NV_E4M3);
    return (uint32_t)lo | ((uint32_t)hi << 16);
}
__device__ __forceinline__ void dequant4(uint32_t q, float sc, float* f) {
    __half2_raw lo = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(q & 0xffffu), __NV_E4M3);
    __half2_raw hi = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(q >> 16), __NV_E4M3);
    float2 a = __half22float2(*reinterpret_cast<__half2*>(&lo));
    float2 b = __half22float2(*reinterpret_cast<__half2*>(&hi));
    f[0] = a.x * sc; f[1] = a.y * sc; f[2] = b.x * sc; f[3] = b.y * sc;
}

// ================================================================================================================
struct PushArgs {
    const void* src;          // local gradient arena (bf16 or fp32), element-indexed
    void* dst;                // slot base in the PS's memory (peer or local), element-indexed in the wire type
    uint8_t* dst_scale;       // fp8 mode: scale bytes, one per 32 elements
    long long lo, hi;         // element range, multiples of 128
    long long* flag;          // arrival flag in the PS's memory
    const long long* step_ptr;
    int* counter;             // local, zero-initialised; used to detect the last CTA
    int mode, src_fp32, zero_src;
};

// Work decomposition: one warp-iteration covers 1024 consecutive elements as 4 rows of 32 x 8-element vectors, lane l taking
// vector (row*32 + l).  Every load/store instruction of a warp therefore touches ONE contiguous 512-byte (bf16), 256-byte
// (fp8) or 1-KB (fp32) span: NVLink packets are full-sized.  (The first version gave each thread 32 consecutive elements;
// its peer stores were 16-byte pieces 64 bytes apart and the 8-GPU incast reached only 183 GB/s, profiles/run_n8.md.)
__global__ void __launch_bounds__(256) push_kernel(PushArgs a) {
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint4 zero = make_uint4(0, 0, 0, 0);
    for (long long base = a.lo + warp0 * 1024; base < a.hi; base += nwarps * 1024) {
        float f[4][8];
        bool ok[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {                       // 4 independent 16/32-byte loads in flight per lane
            const long long i = base + (long long)(k * 32 + lane) * 8;
            ok[k] = i < a.hi;
            if (ok[k]) {
                if (a.src_fp32) {
                    const float* sp = reinterpret_cast<const float*>(a.src) + i;
                    const uint4 x = ld_nc_v4(sp), y = ld_nc_v4(sp + 4);
                    f[k][0] = __uint_as_float(x.x); f[k][1] = __uint_as_float(x.y); f[k][2] = __uint_as_float(x.z); f[k][3] = __uint_as_float(x.w);
                    f[k][4] = __uint_as_float(y.x); f[k][5] = __uint_as_float(y.y); f[k][6] = __uint_as_float(y.z); f[k][7] = __uint_as_float(y.w);
                } else {
                    unpack8(ld_nc_v4(reinterpret_cast<const __nv_bfloat16*>(a.src) + i), f[k]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[k][j] = 0.f;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long i = base + (long long)(k * 32 + lane) * 8;
            if (a.zero_src && ok[k]) {   // next backward accumulates into a clean arena: saves a separate memset pass
                if (a.src_fp32) { float* sp = (float*)a.src + i; st_v4(sp, zero); st_v4(sp + 4, zero); }
                else st_v4((__nv_bfloat16*)a.src + i, zero);
            }
            if (a.mode == WIRE_BF16) {
                if (ok[k]) st_na_v4(reinterpret_cast<__nv_bfloat16*>(a.dst) + i, pack8(f[k]));
            } else if (a.mode == WIRE_FP32) {
                if (ok[k]) {
                    float* d = reinterpret_cast<float*>(a.dst) + i;
                    st_na_v4(d, make_uint4(__float_as_uint(f[k][0]), __float_as_uint(f[k][1]), __float_as_uint(f[k][2]), __float_as_uint(f[k][3])));
                    st_na_v4(d + 4, make_uint4(__float_as_uint(f[k][4]), __float_as_uint(f[k][5]), __float_as_uint(f[k][6]), __float_as_uint(f[k][7])));
                }
            } else {
                // one 32-element quantisation block = 4 neighbouring lanes of this row
                float amax = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) amax = fmaxf(amax, fabsf(f[k][j]));
                amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
                amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
                const int e = e8m0_exponent(amax);
                const float inv = exp2i(-e);
                if (ok[k]) {
                    const uint32_t q0 = quant4(f[k][0], f[k][1], f[k][2], f[k][3], inv), q1 = quant4(f[k][4], f[k][5], f[k][6], f[k][7], inv);
                    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(reinterpret_cast<uint8_t*>(a.dst) + i), "r"(q0), "r"(q1) : "memory");
                    if ((lane & 3) == 0) a.dst_scale[i / 32] = (uint8_t)(e + 127);
                }
            }
        }
    }
    // publish: every CTA fences its peer stores; the last CTA to arrive releases the flag with this worker's step id
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int old = atomicAdd(a.counter, 1);
        if (old == (int)gridDim.x - 1) {
            *a.counter = 0;
            __threadfence_system();
            st_release_sys(a.flag, *a.step_ptr);
        }
    }
}

// ================================================================================================================
struct RUArgs {
    const void* src[kMaxPeers];           // gradient sources, element-indexed (slots in push mode, peer arenas in pull mode)
    const uint8_t* src_scale[kMaxPeers];
    int P, take, mode;                    // P sources; sum the first `take` arrivals
    float inv_div;                        // 1 / divisor
    const long long* g_ready;             // [P][kMaxBuckets] arrival flags (local memory)
    long long* step_ptr;                  // PS step counter (device)
    int NB;
    long long lo[kMaxBuckets], hi[kMaxBuckets];
    float *w, *s0, *s1, *s2;              // fp32 master; SGD: s0 = momentum buf; Adam: s0 = m, s1 = v, s2 = vmax|null
    OptHP* hp;
    void* dst[kMaxPeers + 1];             // bf16 weight replicas (element-indexed); D of them
    int D;
    void* mc_dst;                         // NVLS multicast mapping of the replicas (or null)
    long long* w_ready[kMaxPeers + 1];    // per destination: [kMaxBuckets] "weights ready" flags
    int* counters;                        // [kMaxBuckets]
    long long* err;
    unsigned long long timeout_ns;
    int advance_step;                     // last CTA bumps *step_ptr and hp->steps
    const void* mc_src;                   // NVLS variant only: multicast mapping of the (symmetric) gradient arenas
    unsigned long long* sel;              // [kMaxBuckets] published arrival decision per bucket: (step << 16) | (count << 8) | mask
    int partial_ok;                       // --mode timeout: after partial_ns proceed with the >= 1 gradients that did arrive
    unsigned long long partial_ns;        //   (nothing arrived yet: keep waiting for the first one, up to the hard timeout_ns)
    long long* abort_w[kMaxPeers];        // --mode kill: abort word of source k's worker (set to the step id for the laggards), or null
};

// raw wire words of 8 elements from source k (kept packed in registers so that all K loads can be in flight at once)
struct Raw8 { uint4 lo; uint4 hi; };        // bf16: lo only; fp32: lo+hi; fp8: lo.x,lo.y = payload, lo.z = scale byte
template <int MODE>
__device__ __forceinline__ void load_raw8(const RUArgs& a, int k, long long i, Raw8& r) {
    if (MODE == WIRE_BF16) {
        r.lo = ld_relaxed_sys_v4(reinterpret_cast<const __nv_bfloat16*>(a.src[k]) + i);
    } else if (MODE == WIRE_FP32) {
        const float* s = reinterpret_cast<const float*>(a.src[k]) + i;
        r.lo = ld_relaxed_sys_v4(s); r.hi = ld_relaxed_sys_v4(s + 4);
    } else {
        const uint8_t* q = reinterpret_cast<const uint8_t*>(a.src[k]) + i;
        asm volatile("ld.relaxed.sys.global.v2.u32 {%0,%1}, [%2];" : "=r"(r.lo.x), "=r"(r.lo.y) : "l"(q) : "memory");
        asm volatile("ld.relaxed.sys.global.u8 %0, [%1];" : "=r"(r.lo.z) : "l"(a.src_scale[k] + i / 32) : "memory");
    }
}
template <int MODE>
__device__ __forceinline__ void accum_raw8(const Raw8& r, float* g) {
    float f[8];
    if (MODE == WIRE_BF16) {
        unpack8(r.lo, f);
    } else if (MODE == WIRE_FP32) {
        f[0] = __uint_as_float(r.lo.x); f[1] = __uint_as_float(r.lo.y); f[2] = __uint_as_float(r.lo.z); f[3] = __uint_as_float(r.lo.w);
        f[4] = __uint_as_float(r.hi.x); f[5] = __uint_as_float(r.hi.y); f[6] = __uint_as_float(r.hi.z); f[7] = __uint_as_float(r.hi.w);
    } else {
        const float sc = exp2i(max(-126, (int)r.lo.z - 127));
        dequant4(r.lo.x, sc, f);
        dequant4(r.lo.y, sc, f + 4);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] += f[j];
}

// NVLS = true (EXPERIMENTAL, not yet run on hardware; bf16 pull mode with take == P == every device of the multicast group): the P peer loads per element
// are replaced by one multimem.ld_reduce through the switch, so the server receives one reduced stream instead of P.
template <int MODE, bool NVLS = false>
__global__ void __launch_bounds__(256, 2) reduce_update_kernel(RUArgs a) {
    __shared__ unsigned s_mask, s_cnt;
    const long long step = *a.step_ptr;
    const OptHP hp = *a.hp;
    const bool first = hp.steps == 0;
    const float adam_ss = hp.kind == 1 ? adam_step_size(hp, hp.steps + 1) : 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x * 8;

    for (int b = a.NB - 1; b >= 0; --b) {          // gradients arrive last-layer-first
        // ONE arrival decision per (step, bucket) for the whole grid: every CTA polls, the first one that sees `take` arrivals (or, in
        // timeout mode, the deadline with >= 1 arrival) publishes (step, count, mask) with a CAS and everybody uses the published set —
        // with K < P different CTAs would otherwise average different subsets of workers in different tiles of the same update.
        if (threadIdx.x == 0) {
            const unsigned long long tag = (unsigned long long)step << 16;
            const unsigned long long t0 = globaltimer_ns();
            unsigned long long v;
            bool winner = false;
            while (true) {
                v = *reinterpret_cast<volatile unsigned long long*>(a.sel + b);
                if ((v >> 16) == (unsigned long long)step) break;                   // already decided
                unsigned mask = 0;
                int cnt = 0;
                for (int k = 0; k < a.P && cnt < a.take; ++k)
                    if (ld_acquire_sys(a.g_ready + k * kMaxBuckets + b) == step) { mask |= 1u << k; ++cnt; }
                const unsigned long long waited = globaltimer_ns() - t0;
                const bool timed = waited > a.timeout_ns;
                if (cnt >= a.take || (a.partial_ok && waited > a.partial_ns && cnt >= 1)) {
                    const unsigned long long nv = tag | ((unsigned long long)cnt << 8) | mask;
                    if (atomicCAS(a.sel + b, v, nv) == v) { v = nv; winner = true; break; }
                    continue;                                                       // lost the race: read the winner's decision
                }
                if (timed) {                                                        // nobody (or, without partial_ok, too few) arrived
                    atomicExch((unsigned long long*)a.err, 1ull + (unsigned long long)b);
                    atomicCAS(a.sel + b, v, tag);                                   // publish "failed" (mask 0) so the grid leaves together
                    v = tag;
                    break;
                }
                __nanosleep(200);
            }
            s_mask = (unsigned)(v & 0xffu);
            s_cnt = (unsigned)((v >> 8) & 0xffu);
            if (winner) {
                __threadfence();
                for (int k = 0; k < a.P; ++k)                                       // straggler kill: tell the laggards to abandon this step
                    if (!((v >> k) & 1ull) && a.abort_w[k] != nullptr) st_release_sys(a.abort_w[k], step);
            }
        }
        __syncthreads();
        const unsigned mask = s_mask;
        if (mask == 0) return;                     // timed out: error word is set, host raises
        const float inv_div = ((int)s_cnt < a.take) ? 1.f / (float)max(1u, s_cnt) : a.inv_div;     // partial set: mean of what arrived
        for (long long i = a.lo[b] + ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < a.hi[b]; i += stride) {
            float g[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            if (NVLS) {
                unpack8(multimem_ld_reduce_bf16x8(reinterpret_cast<const __nv_bfloat16*>(a.mc_src) + i), g);
            } else {
                // issue every selected source's load before the first add: up to 8 independent P2P/HBM loads in flight per
                // lane, held as packed wire words (4-8 registers each)
                constexpr int kBatch = MODE == WIRE_FP32 ? 4 : kMaxPeers;
#pragma unroll
                for (int k0 = 0; k0 < kMaxPeers; k0 += kBatch) {
                    Raw8 raw[kBatch];
#pragma unroll
                    for (int k = 0; k < kBatch; ++k)
                        if (k0 + k < a.P && (mask >> (k0 + k) & 1)) load_raw8<MODE>(a, k0 + k, i, raw[k]);
#pragma unroll
                    for (int k = 0; k < kBatch; ++k)
                        if (k0 + k < a.P && (mask >> (k0 + k) & 1)) accum_raw8<MODE>(raw[k], g);
                }
            }
            float wf[8], x0[8], x1[8], x2[8];
            if (hp.kind != 2) {
                uint4 p = ld_v4(a.w + i), q = ld_v4(a.w + i + 4);
                wf[0] = __uint_as_float(p.x); wf[1] = __uint_as_float(p.y); wf[2] = __uint_as_float(p.z); wf[3] = __uint_as_float(p.w);
                wf[4] = __uint_as_float(q.x); wf[5] = __uint_as_float(q.y); wf[6] = __uint_as_float(q.z); wf[7] = __uint_as_float(q.w);
            }
            auto ldf = [&](const float* ptr, float* o) {
                uint4 p = ld_v4(ptr + i), q = ld_v4(ptr + i + 4);
                o[0] = __uint_as_float(p.x); o[1] = __uint_as_float(p.y); o[2] = __uint_as_float(p.z); o[3] = __uint_as_float(p.w);
                o[4] = __uint_as_float(q.x); o[5] = __uint_as_float(q.y); o[6] = __uint_as_float(q.z); o[7] = __uint_as_float(q.w);
            };
            auto stf = [&](float* ptr, const float* o) {
                st_v4(ptr + i, make_uint4(__float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(o[3])));
                st_v4(ptr + i + 4, make_uint4(__float_as_uint(o[4]), __float_as_uint(o[5]), __float_as_uint(o[6]), __float_as_uint(o[7])));
            };
            if (hp.kind == 2) {
                // all-reduce mode (DDP): no optimizer, the averaged gradient itself is stored back to every peer
#pragma unroll
                for (int j = 0; j < 8; ++j) wf[j] = g[j] * inv_div;
            } else if (hp.kind == 0) {
                if (hp.momentum != 0.f && !first) ldf(a.s0, x0);
#pragma unroll
                for (int j = 0; j < 8; ++j) wf[j] = sgd_update(wf[j], g[j] * inv_div, x0[j], hp, first);
                if (hp.momentum != 0.f) stf(a.s0, x0);
            } else {
                ldf(a.s0, x0); ldf(a.s1, x1);
                if (a.s2) ldf(a.s2, x2);
#pragma unroll
                for (int j = 0; j < 8; ++j) wf[j] = adam_update(wf[j], g[j] * inv_div, x0[j], x1[j], a.s2 ? &x2[j] : nullptr, hp, adam_ss);
                stf(a.s0, x0); stf(a.s1, x1);
                if (a.s2) stf(a.s2, x2);
            }
            if (hp.kind != 2) stf(a.w, wf);
            // broadcast the fresh compute copy: one multicast store through the switch, or D peer stores
            const uint4 lp = pack8(wf);
            if (a.mc_dst) {
                multimem_st_v4(reinterpret_cast<__nv_bfloat16*>(a.mc_dst) + i, lp);
            } else {
#pragma unroll 1
                for (int d = 0; d < a.D; ++d)
                    if (d == 0 || a.dst[d] != a.dst[0]) st_na_v4(reinterpret_cast<__nv_bfloat16*>(a.dst[d]) + i, lp);   // pull mode lists the PS replica D times
            }
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const int old = atomicAdd(a.counters + b, 1);
            if (old == (int)gridDim.x - 1) {
                a.counters[b] = 0;
                __threadfence_system();
                for (int d = 0; d < a.D; ++d) st_release_sys(a.w_ready[d] + b, step + 1);
                if (b == 0 && a.advance_step) { a.hp->steps = hp.steps + 1; *a.step_ptr = step + 1; }
            }
        }
        __syncthreads();
    }
}

// Initial publish / re-publish of the bf16 replica (step 1, resume) and the worker-pull path (--comm-type Async):
// copy `src` (local or peer) to D destinations, then release their flags with value `*step_ptr + step_add`.
struct BcastArgs {
    const void* src; void* dst[kMaxPeers + 1]; int D; void* mc_dst;
    long long n_vec;            // 16-byte vectors
    long long* flags[kMaxPeers + 1]; int nflags;      // per destination: set flags[0..nflags)
    const long long* step_ptr; long long step_add; int* counter;
};
__global__ void __launch_bounds__(256) bcast_kernel(BcastArgs a) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n_vec; i += stride) {
        const uint4 v = ld_relaxed_sys_v4(reinterpret_cast<const uint4*>(a.src) + i);
        if (a.mc_dst) multimem_st_v4(reinterpret_cast<uint4*>(a.mc_dst) + i, v);
        else for (int d = 0; d < a.D; ++d) st_na_v4(reinterpret_cast<uint4*>(a.dst[d]) + i, v);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int old = atomicAdd(a.counter, 1);
        if (old == (int)gridDim.x - 1) {
            *a.counter = 0;
            __threadfence_system();
            const long long v = *a.step_ptr + a.step_add;
            for (int d = 0; d < a.D; ++d)
                for (int f = 0; f < a.nflags; ++f) st_release_sys(a.flags[d] + f, v);
        }
    }
}

// acquire-spin until flags[0..n) >= *step_ptr (weights for this worker's step have landed), with timeout.  A worker that fell behind
// (its gradient missed a K-of-P decision, or it abandoned a step) finds a flag value AHEAD of its own counter: it adopts the server's step
// (the reference's workers learn the step id from the PS, distributed_worker.py:193-199) — otherwise every later push of this worker
// would carry a stale step id and be ignored for the rest of the job.
__global__ void wait_flag_kernel(const long long* flags, int n, long long* step_ptr, long long* err, unsigned long long timeout_ns) {
    const long long want = *step_ptr;
    __syncthreads();
    if (threadIdx.x >= n) return;
    const unsigned long long t0 = globaltimer_ns();
    long long f;
    while ((f = ld_acquire_sys(flags + threadIdx.x)) < want) {
        if (globaltimer_ns() - t0 > timeout_ns) { atomicExch((unsigned long long*)err, 100ull + threadIdx.x); return; }
        __nanosleep(100);
    }
    if (threadIdx.x == 0 && f > want) *step_ptr = f;
}

__global__ void add_step_kernel(long long* p, long long v) { *p += v; }
__global__ void set_word_kernel(long long* p, long long v) { __threadfence_system(); st_release_sys(p, v); }

// ---- host wrappers (raw pointers as int64: they may be peer mappings that no torch tensor owns) ---------------------
static inline void* P(int64_t x) { return reinterpret_cast<void*>(x); }

static int grid_for(long long elems, int per_thread) {
    long long blocks = (elems / per_thread + 255) / 256;
    return (int)std::max<long long>(1, std::min<long long>(blocks, (long long)kNumSMs * 4));
}

void push(int64_t src, bool src_fp32, int64_t dst, int64_t dst_scale, int64_t lo, int64_t hi, int64_t flag, int64_t step_ptr,
          int64_t counter, int64_t mode, bool zero_src) {
    TORCH_CHECK(lo % 128 == 0 && hi % 128 == 0 && hi > lo, "push range must be 128-element aligned");
    PushArgs a{P(src), P(dst), (uint8_t*)P(dst_scale), lo, hi, (long long*)P(flag), (const long long*)P(step_ptr), (int*)P(counter),
               (int)mode, src_fp32 ? 1 : 0, zero_src ? 1 : 0};
    push_kernel<<<grid_for(hi - lo, 32), 256, 0, at::cuda::getCurrentCUDAStream()>>>(a);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// one-shot selector of the NVLS variant for the NEXT reduce_update calls (0 = off): multicast address of the gradient arenas
static int64_t g_ru_mc_src = 0;
void ru_set_mc_src(int64_t p) { g_ru_mc_src = p; }

void reduce_update(std::vector<int64_t> src, std::vector<int64_t> src_scale, int64_t take, double inv_div, int64_t mode,
                   int64_t g_ready, int64_t step_ptr, std::vector<int64_t> lo, std::vector<int64_t> hi, at::Tensor w,
                   c10::optional<at::Tensor> s0, c10::optional<at::Tensor> s1, c10::optional<at::Tensor> s2, int64_t hp,
                   std::vector<int64_t> dst, int64_t mc_dst, std::vector<int64_t> w_ready, int64_t counters, int64_t err,
                   double timeout_s, bool advance_step, int64_t sel, bool partial_ok, std::vector<int64_t> abort_words, double partial_s) {
    RUArgs a{};
    TORCH_CHECK(sel != 0, "reduce_update needs the per-bucket decision words");
    a.sel = (unsigned long long*)P(sel); a.partial_ok = partial_ok ? 1 : 0; a.partial_ns = (unsigned long long)(partial_s * 1e9);
    for (size_t k = 0; k < kMaxPeers; ++k) a.abort_w[k] = k < abort_words.size() ? (long long*)P(abort_words[k]) : nullptr;
    TORCH_CHECK(src.size() <= kMaxPeers && dst.size() <= kMaxPeers + 1 && lo.size() <= kMaxBuckets && lo.size() == hi.size());
    a.P = (int)src.size(); a.take = (int)take; a.mode = (int)mode; a.inv_div = (float)inv_div;
    for (size_t k = 0; k < src.size(); ++k) { a.src[k] = P(src[k]); a.src_scale[k] = (const uint8_t*)P(k < src_scale.size() ? src_scale[k] : 0); }
    a.g_ready = (const long long*)P(g_ready); a.step_ptr = (long long*)P(step_ptr);
    a.NB = (int)lo.size();
    for (size_t b = 0; b < lo.size(); ++b) { TORCH_CHECK(lo[b] % 32 == 0 && hi[b] % 32 == 0); a.lo[b] = lo[b]; a.hi[b] = hi[b]; }
    TORCH_CHECK(w.scalar_type() == at::kFloat && w.is_cuda());
    a.w = w.data_ptr<float>();
    a.s0 = s0.has_value() ? s0->data_ptr<float>() : nullptr;
    a.s1 = s1.has_value() ? s1->data_ptr<float>() : nullptr;
    a.s2 = s2.has_value() ? s2->data_ptr<float>() : nullptr;
    a.hp = (OptHP*)P(hp);
    a.D = (int)dst.size();
    for (size_t d = 0; d < dst.size(); ++d) { a.dst[d] = P(dst[d]); a.w_ready[d] = (long long*)P(w_ready[d]); }
    a.mc_dst = P(mc_dst);
    a.counters = (int*)P(counters); a.err = (long long*)P(err);
    a.timeout_ns = (unsigned long long)(timeout_s * 1e9);
    a.advance_step = advance_step;
    long long tot = 0;
    for (size_t b = 0; b < lo.size(); ++b) tot = std::max<long long>(tot, hi[b] - lo[b]);
    const int grid = grid_for(tot, 8);
    auto stream = at::cuda::getCurrentCUDAStream();
    a.mc_src = nullptr;
    if (g_ru_mc_src && mode == WIRE_BF16 && a.take == a.P) {
        a.mc_src = P(g_ru_mc_src);
        reduce_update_kernel<WIRE_BF16, true><<<grid, 256, 0, stream>>>(a);
    } else if (mode == WIRE_BF16) reduce_update_kernel<WIRE_BF16><<<grid, 256, 0, stream>>>(a);
    else if (mode == WIRE_FP8) reduce_update_kernel<WIRE_FP8><<<grid, 256, 0, stream>>>(a);
    else reduce_update_kernel<WIRE_FP32><<<grid, 256, 0, stream>>>(a);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void bcast(int64_t src, std::vector<int64_t> dst, int64_t mc_dst, int64_t nbytes, std::vector<int64_t> flags, int64_t nflags,
           int64_t step_ptr, int64_t step_add, int64_t counter) {
    BcastArgs a{};
    TORCH_CHECK(nbytes % 16 == 0 && dst.size() <= kMaxPeers + 1 && flags.size() == dst.size());
    a.src = P(src); a.D = (int)dst.size(); a.mc_dst = P(mc_dst); a.n_vec = nbytes / 16;
    for (size_t d = 0; d < dst.size(); ++d) { a.dst[d] = P(dst[d]); a.flags[d] = (long long*)P(flags[d]); }
    a.nflags = (int)nflags; a.step_ptr = (const long long*)P(step_ptr); a.step_add = step_add; a.counter = (int*)P(counter);
    bcast_kernel<<<grid_for(a.n_vec, 1), 256, 0, at::cuda::getCurrentCUDAStream()>>>(a);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void wait_flag(int64_t flags, int64_t n, int64_t step_ptr, int64_t err, double timeout_s) {
    TORCH_CHECK(n >= 1 && n <= kMaxBuckets);
    wait_flag_kernel<<<1, 32, 0, at::cuda::getCurrentCUDAStream()>>>((const long long*)P(flags), (int)n, (long long*)P(step_ptr),
                                                                      (long long*)P(err), (unsigned long long)(timeout_s * 1e9));
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void add_step(int64_t p, int64_t v) {
    add_step_kernel<<<1, 1, 0, at::cuda::getCurrentCUDAStream()>>>((long long*)P(p), v);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}
void set_word(int64_t p, int64_t v) {
    set_word_kernel<<<1, 1, 0, at::cuda::getCurrentCUDAStream()>>>((long long*)P(p), v);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// OptHP blob <-> host
at::Tensor make_hp(double lr, double momentum, double dampening, double wd, bool nesterov, double beta1, double beta2, double eps,
                   bool amsgrad, int64_t kind, int64_t steps) {
    OptHP h{};
    h.lr = (float)lr; h.momentum = (float)momentum; h.dampening = (float)dampening; h.weight_decay = (float)wd;
    h.beta1 = (float)beta1; h.beta2 = (float)beta2; h.eps = (float)eps; h.nesterov = nesterov; h.amsgrad = amsgrad;
    h.kind = (int)kind; h.steps = steps;
    auto t = at::empty({(long long)sizeof(OptHP)}, at::TensorOptions().dtype(at::kByte));
    memcpy(t.data_ptr(), &h, sizeof(OptHP));
    return t;
}
int64_t hp_steps(at::Tensor host_blob) {
    OptHP h;
    memcpy(&h, host_blob.data_ptr(), sizeof(OptHP));
    return h.steps;
}

}  // namespace psb

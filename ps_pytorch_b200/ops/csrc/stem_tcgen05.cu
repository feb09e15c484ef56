// Network stem on the tensor cores: 3x3 / stride 1 / pad 1 convolution with Cin <= 3 (RGB -> 64 channels), forward and weight gradient.
//
// K = 9 * Cin <= 27 is far too short for a TMA-fed implicit GEMM (a 64-channel block would be 95 % padding), and until now both passes went
// to cuDNN (forward) / a CUDA-core kernel (weight gradient, 199 us for 134 MB of dY).  Both are really memory-bound streams — 134 MB of
// output to write, 134 MB of dY to read — so the tensor core only has to keep up with HBM:
//
//   im2col rows are built BY THREADS in shared memory: a tile is 128 output pixels (128 / W full image rows); the (rows + 2) x (W + 2) x Cin
//   input patch is staged zero-padded, and thread p packs the 27 values of pixel p (three runs of 3 * Cin contiguous elements) into one
//   128-byte row in the SWIZZLE_128B pattern (chunk c of row p at 16 * (c ^ (p & 7))); K is padded to 32, so only the first 64 bytes of a
//   row carry data.  The same bytes serve as
//     forward:          the K-major A operand  [128 pixels][32]   x  weights [64 co][32]            -> D[128 px][64 co]   (2 MMAs of K = 16)
//     weight gradient:  the MN-major B operand [128 pixels][64]   x  dY tile [128 pixels][64 co] (TMA) -> D[64 co][64]    (8 MMAs, K = pixels)
//   generic-proxy stores are made visible to the tensor core with fence.proxy.async before the hand-over barrier.
//
// Reference op replaced: the first nn.Conv2d of every model (/root/reference/src/model_ops/resnet.py:72, vgg.py:53) -> cuDNN.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "tc_common.cuh"

namespace psb {

struct StemParams {
    // epilogue_tile contract
    int N;                               // Cout (= 64)
    const __nv_bfloat16* bias;
    const __nv_bfloat16* residual;
    int relu;
    __nv_bfloat16* C;
    float* stats;
    // geometry
    const __nv_bfloat16* x;              // [NB,H,W,Cin]
    const __nv_bfloat16* w;              // [64][3][3][Cin]
    int NB, H, W, Cin, TR, n_tiles, lw;  // TR = 128 / W image rows per tile, lw = log2(W)
    float* dW32;                         // wgrad: fp32 [64][9*Cin] accumulator (zero on entry)
    int* err;
};

constexpr int kStemRowBytes = 128, kStemTile = 128 * kStemRowBytes;      // one operand tile: 128 rows of 128 bytes
constexpr int kStemXinMax = 6 * 34 * 3;                                  // staged patch elements (W = 32, Cin = 3)

// ---- patch staging: one 16-byte global load per thread and tile, issued one tile AHEAD (register prefetch) ------------------------------
// A patch row is W * CIN contiguous bf16 in global memory (a multiple of 16 bytes for W >= 8); vector t of a tile is (row = t / nvr, v = t % nvr).
// xin[(TR+2)][(W+2) * CIN]: CIN zero elements left and right of every row (written once), rows outside the image arrive as zero vectors.
__device__ __forceinline__ uint4 stem_load_vec(const StemParams& p, int n, int h0, int t, int nvr) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (t < (p.TR + 2) * nvr) {
        const int row = t / nvr, vi = t - row * nvr, h = h0 + row - 1;
        if (h >= 0 && h < p.H) v = __ldg(reinterpret_cast<const uint4*>(p.x + ((long long)n * p.H + h) * p.W * p.Cin) + vi);
    }
    return v;
}
__device__ __forceinline__ void stem_store_vec(const StemParams& p, uint16_t* xin, int t, int nvr, const uint4& v) {
    if (t < (p.TR + 2) * nvr) {
        const int row = t / nvr, vi = t - row * nvr;
        uint16_t* d = xin + row * (p.W + 2) * p.Cin + p.Cin + vi * 8;          // 2-byte aligned only (CIN may be odd)
        d[0] = (uint16_t)(v.x & 0xffffu); d[1] = (uint16_t)(v.x >> 16); d[2] = (uint16_t)(v.y & 0xffffu); d[3] = (uint16_t)(v.y >> 16);
        d[4] = (uint16_t)(v.z & 0xffffu); d[5] = (uint16_t)(v.z >> 16); d[6] = (uint16_t)(v.w & 0xffffu); d[7] = (uint16_t)(v.w >> 16);
    }
}

// thread `pix` writes the im2col row of its pixel: bytes [0, 64) = 32 bf16 (k = r * 3 * CIN + s * CIN + c, zero beyond 9 * CIN) as swizzled
// 16-byte chunks; zero_upper also clears bytes [64, 128) (needed when the tile is read as an N = 64 operand).  CIN is a template parameter so the
// 9 * CIN shared-memory loads have static register destinations (no division, no local memory).
template <int CIN>
__device__ __forceinline__ void stem_build_row(const StemParams& p, const uint16_t* xs, uint8_t* tile, int pix, bool zero_upper) {
    constexpr int RUN = 3 * CIN;
    const int tr = pix >> p.lw, wq = pix & (p.W - 1);
    const int rowlen = (p.W + 2) * CIN;
    uint32_t v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = 0u;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const uint16_t* src = xs + (tr + r) * rowlen + wq * CIN;
#pragma unroll
        for (int j = 0; j < RUN; ++j) v[r * RUN + j] = src[j];
    }
    uint8_t* row = tile + pix * kStemRowBytes;
    const int sw = pix & 7;
#pragma unroll
    for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(row + ((c ^ sw) << 4)) = make_uint4(v[8 * c] | (v[8 * c + 1] << 16), v[8 * c + 2] | (v[8 * c + 3] << 16),
                                                                      v[8 * c + 4] | (v[8 * c + 5] << 16), v[8 * c + 6] | (v[8 * c + 7] << 16));
    if (zero_upper) {
#pragma unroll
        for (int c = 4; c < 8; ++c) *reinterpret_cast<uint4*>(row + ((c ^ sw) << 4)) = make_uint4(0, 0, 0, 0);
    }
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// =================================================================================================================== forward
template <int CIN>
__global__ void __launch_bounds__(kThreads, 1) stem_fprop_kernel(const StemParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sa = smem;                                   // 2 x 16 KB im2col tiles
    uint8_t* sb = smem + 2 * kStemTile;                   // weights [64][128 B]
    __nv_bfloat16* xin = reinterpret_cast<__nv_bfloat16*>(sb + 64 * kStemRowBytes);      // 2 staged patches
    uint64_t* a_empty = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xin) + 2 * ((kStemXinMax * 2 + 127) & ~127));
    uint64_t* tmem_full = a_empty + 2;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(a_empty + s, 1); mbar_init(tmem_full + s, 1); mbar_init(tmem_empty + s, 4); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_base_smem, 128);
    // weights: row co = the 9 * Cin filter values of output channel co (already in k order), padded to 32, same swizzle
    if (threadIdx.x < 64) {
        const int co = threadIdx.x, kk = 9 * p.Cin;
        uint32_t packed[16];
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) {
            const uint32_t lo = 2 * k2 < kk ? reinterpret_cast<const uint16_t*>(p.w)[co * kk + 2 * k2] : 0u;
            const uint32_t hi = 2 * k2 + 1 < kk ? reinterpret_cast<const uint16_t*>(p.w)[co * kk + 2 * k2 + 1] : 0u;
            packed[k2] = lo | (hi << 16);
        }
        uint8_t* row = sb + co * kStemRowBytes;
#pragma unroll
        for (int c = 0; c < 4; ++c)
            *reinterpret_cast<uint4*>(row + ((c ^ (co & 7)) << 4)) = make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
    }
    const int xin_stride = ((kStemXinMax * 2 + 127) & ~127) / 2;
    for (int e = threadIdx.x; e < 2 * xin_stride; e += kThreads) reinterpret_cast<uint16_t*>(xin)[e] = 0;      // the zero borders are written once
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    const int tiles_per_img = p.H / p.TR;
    const int nvr = p.W * CIN / 8;

    if (warp < 4) {
        // ===================================================== builders (128 threads) + MMA issue by warp 0
        const int tid = threadIdx.x;
        constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
        int it = 0;
        uint4 pre = make_uint4(0, 0, 0, 0);
        if ((int)blockIdx.x < p.n_tiles) { const int n = blockIdx.x / tiles_per_img; pre = stem_load_vec(p, n, ((int)blockIdx.x - n * tiles_per_img) * p.TR, tid, nvr); }
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
            const int s = it & 1;
            const uint32_t ph = (uint32_t)(it >> 1) & 1u;
            const int n = tile / tiles_per_img, h0 = (tile - n * tiles_per_img) * p.TR;
            const uint4 cur = pre;
            const int nxt = tile + gridDim.x;
            if (nxt < p.n_tiles) { const int nn = nxt / tiles_per_img; pre = stem_load_vec(p, nn, (nxt - nn * tiles_per_img) * p.TR, tid, nvr); }   // in flight during this tile
            if (!mbar_wait(a_empty + s, ph ^ 1)) *p.err = 61;                      // the MMAs that read this A buffer have retired (a timeout is
                                                                                   // reported and the loop goes on: nobody may skip the named barriers)
            uint16_t* xs = reinterpret_cast<uint16_t*>(xin) + s * xin_stride;
            stem_store_vec(p, xs, tid, nvr, cur);
            named_bar_sync(1, 128);
            stem_build_row<CIN>(p, xs, sa + s * kStemTile, tid, false);
            fence_proxy_async();
            named_bar_sync(1, 128);
            if (warp == 0) {
                if (!mbar_wait(tmem_empty + s, ph ^ 1)) *p.err = 62;
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a0 = smem_u32(sa + s * kStemTile), b0 = smem_u32(sb);
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        tc_mma_bf16(tmem_base + s * 64, umma_desc_k_sw128(a0 + k * 32), umma_desc_k_sw128(b0 + k * 32), idesc, k != 0);
                    tc_commit(a_empty + s);
                    tc_commit(tmem_full + s);
                }
                __syncwarp();
            }
        }
    } else {
        // ===================================================== epilogue (4 warps)
        const int q = warp & 3;
        float* const srow = p.stats ? p.stats + (size_t)(blockIdx.x * 4 + q) * 2 * p.N : nullptr;
        const bool use_racc = srow != nullptr;
        float racc[128];
#pragma unroll
        for (int j = 0; j < 128; ++j) racc[j] = 0.f;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
            const int s = it & 1;
            const uint32_t ph = (uint32_t)(it >> 1) & 1u;
            const int n = tile / tiles_per_img, h0 = (tile - n * tiles_per_img) * p.TR;
            const int m = q * 32 + lane;
            const long long out_row = ((long long)n * p.H + h0) * p.W + m;          // the tile's 128 pixels are consecutive in NHWC order
            if (!mbar_wait(tmem_full + s, ph)) { if (lane == 0) *p.err = 63; }
            tc_fence_after();
            epilogue_tile<64>(p, tmem_base, s, q, lane, out_row, 0, false, srow, use_racc ? racc : nullptr);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_relaxed(tmem_empty + s);
        }
        if (use_racc) stats_flush<64>(racc, srow, p.N, 0, lane);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 2) tmem_dealloc(tmem_base, 128);
}

// =================================================================================================================== weight gradient
constexpr int kStemWgStages = 4;

template <int CIN>
__global__ void __launch_bounds__(kThreads, 1) stem_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_dy, const StemParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sdy = smem;                                        // kStemWgStages x 16 KB dY tiles (TMA)
    uint8_t* sxi = smem + kStemWgStages * kStemTile;            // 2 x 16 KB im2col tiles (threads)
    __nv_bfloat16* xin = reinterpret_cast<__nv_bfloat16*>(sxi + 2 * kStemTile);
    const int xin_stride = ((kStemXinMax * 2 + 127) & ~127) / 2;
    uint64_t* dy_full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xin) + 2 * xin_stride * 2);
    uint64_t* dy_empty = dy_full + kStemWgStages;
    uint64_t* xi_full = dy_empty + kStemWgStages;
    uint64_t* xi_empty = xi_full + 2;
    uint64_t* tmem_full = xi_empty + 2;
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) prefetch_tmap(&tmap_dy);
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStemWgStages; ++s) { mbar_init(dy_full + s, 1); mbar_init(dy_empty + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(xi_full + s, 128); mbar_init(xi_empty + s, 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_base_smem, 64);
    for (int e = threadIdx.x; e < 2 * xin_stride; e += kThreads) reinterpret_cast<uint16_t*>(xin)[e] = 0;      // zero borders, written once
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    const int tiles_per_img = p.H / p.TR;
    const int nvr = p.W * CIN / 8;
    const int my_tiles = blockIdx.x < p.n_tiles ? (p.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (warp == 0) {
        // ===================================================== TMA producer: dY tile = [TR rows][W pixels][64 co] = 128 rows of 128 B
        int stage = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            const int n = tile / tiles_per_img, h0 = (tile - n * tiles_per_img) * p.TR;
            if (!mbar_wait(dy_empty + stage, phase ^ 1)) { *p.err = 71; return; }
            if (elect_one()) {
                mbar_arrive_expect_tx(dy_full + stage, kStemTile);
                tma_load_4d(sdy + stage * kStemTile, &tmap_dy, dy_full + stage, 0, 0, h0, n);
            }
            __syncwarp();
            if (++stage == kStemWgStages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer: D[64 co][64] += dY^T (MN-major A) x im2col (MN-major B)
        constexpr uint32_t idesc = umma_idesc_bf16(64, 64, 1, 1);
        int stage = 0; uint32_t phase = 0;
        for (int it = 0; it < my_tiles; ++it) {
            const int b = it & 1;
            if (!mbar_wait(dy_full + stage, phase)) { *p.err = 72; return; }
            if (!mbar_wait(xi_full + b, (uint32_t)(it >> 1) & 1u)) { *p.err = 73; return; }
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a0 = smem_u32(sdy + stage * kStemTile), b0 = smem_u32(sxi + b * kStemTile);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    tc_mma_bf16(tmem_base, umma_desc_mn_sw128(a0 + j * 2048, kStemTile), umma_desc_mn_sw128(b0 + j * 2048, kStemTile), idesc, (it | j) != 0);
                tc_commit(dy_empty + stage);
                tc_commit(xi_empty + b);
            }
            __syncwarp();
            if (++stage == kStemWgStages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) tc_commit(tmem_full);
        __syncwarp();
    } else if (warp >= kEpiWarp0) {
        // ===================================================== im2col builders (128 threads), then the final drain of D
        const int tid = threadIdx.x - kEpiWarp0 * 32, q = warp & 3;
        int it = 0;
        uint4 pre = make_uint4(0, 0, 0, 0);
        if ((int)blockIdx.x < p.n_tiles) { const int n = blockIdx.x / tiles_per_img; pre = stem_load_vec(p, n, ((int)blockIdx.x - n * tiles_per_img) * p.TR, tid, nvr); }
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
            const int b = it & 1;
            const uint4 cur = pre;
            const int nxt = tile + gridDim.x;
            if (nxt < p.n_tiles) { const int nn = nxt / tiles_per_img; pre = stem_load_vec(p, nn, (nxt - nn * tiles_per_img) * p.TR, tid, nvr); }
            if (!mbar_wait(xi_empty + b, ((uint32_t)(it >> 1) & 1u) ^ 1)) *p.err = 74;
            uint16_t* xs = reinterpret_cast<uint16_t*>(xin) + b * xin_stride;
            stem_store_vec(p, xs, tid, nvr, cur);
            named_bar_sync(1, 128);
            stem_build_row<CIN>(p, xs, sxi + b * kStemTile, tid, true);
            fence_proxy_async();
            mbar_arrive(xi_full + b);                     // release: the row stores above are ordered before the arrival
        }
        if (my_tiles > 0) {
            if (!mbar_wait(tmem_full, 0)) { if (lane == 0) *p.err = 75; return; }
            tc_fence_after();
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16), v);        // M = 64: row m lives in lane (m % 16) + 32 * (m / 16)
            tmem_ld_wait();
            const int kk = 9 * p.Cin;
            if (lane < 16) {
                float* d = p.dW32 + (long long)(q * 16 + lane) * kk;
                for (int k = 0; k < kk; ++k) atomicAdd(d + k, __uint_as_float(v[k]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 2) tmem_dealloc(tmem_base, 64);
}

__global__ void stem_finish_kernel(float* __restrict__ acc, __nv_bfloat16* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { out[i] = __float2bfloat16_rn(acc[i]); acc[i] = 0.f; }
}

// ------------------------------------------------------------------------------------------------------------------ host
bool stem_tc_eligible(int H, int W, int Cin, int Cout, int R, int S, int stride, int pad) {
    return R == 3 && S == 3 && stride == 1 && pad == 1 && Cin >= 1 && Cin <= 3 && Cout == 64 && (W == 8 || W == 16 || W == 32) && H % (128 / W) == 0;
}

static int stem_smem_bytes(bool wgrad) {
    const int xin = 2 * ((kStemXinMax * 2 + 127) & ~127);
    return (wgrad ? (kStemWgStages + 2) * kStemTile : 2 * kStemTile + 64 * kStemRowBytes) + xin + 256 + 1024;
}

const char* stem_fprop(const void* X, const void* Wt, void* Y, const void* bias, int NB, int H, int W, int Cin, int Cout, int relu, float* stats,
                       cudaStream_t stream) {
    if (!stem_tc_eligible(H, W, Cin, Cout, 3, 3, 1, 1)) return "stem kernel: needs 3x3/s1/p1, Cin <= 3, Cout == 64, W in {8,16,32}, H % (128/W) == 0";
    StemParams p{};
    p.N = Cout; p.bias = (const __nv_bfloat16*)bias; p.relu = relu; p.C = (__nv_bfloat16*)Y; p.stats = stats;
    p.x = (const __nv_bfloat16*)X; p.w = (const __nv_bfloat16*)Wt; p.NB = NB; p.H = H; p.W = W; p.Cin = Cin; p.TR = 128 / W;
    p.n_tiles = NB * (H / p.TR); p.err = err_word();
    p.lw = W == 32 ? 5 : (W == 16 ? 4 : 3);
    const int smem = stem_smem_bytes(false);
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(stem_fprop_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
            cudaFuncSetAttribute(stem_fprop_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
            cudaFuncSetAttribute(stem_fprop_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return "cannot raise dynamic smem";
        attr = true;
    }
    const int grid = p.n_tiles < num_sms() ? p.n_tiles : num_sms();
    if (Cin == 1) stem_fprop_kernel<1><<<grid, kThreads, smem, stream>>>(p);
    else if (Cin == 2) stem_fprop_kernel<2><<<grid, kThreads, smem, stream>>>(p);
    else stem_fprop_kernel<3><<<grid, kThreads, smem, stream>>>(p);
    return cudaGetLastError() == cudaSuccess ? nullptr : "stem_fprop launch failed";
}

// dW (bf16 [64][3][3][Cin]) from dY [NB,H,W,64] and X [NB,H,W,Cin]; scratch32: zeroed fp32 [64 * 9 * Cin], left zeroed
const char* stem_wgrad_tc(const void* dY, const void* X, void* dW, float* scratch32, int NB, int H, int W, int Cin, int Cout, cudaStream_t stream) {
    if (!stem_tc_eligible(H, W, Cin, Cout, 3, 3, 1, 1)) return "stem kernel: shape not eligible";
    StemParams p{};
    p.N = Cout; p.x = (const __nv_bfloat16*)X; p.NB = NB; p.H = H; p.W = W; p.Cin = Cin; p.TR = 128 / W;
    p.n_tiles = NB * (H / p.TR); p.dW32 = scratch32; p.err = err_word();
    p.lw = W == 32 ? 5 : (W == 16 ? 4 : 3);
    CUtensorMap ta;
    uint64_t da[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)NB};
    uint64_t sa[3] = {(uint64_t)Cout * 2, (uint64_t)W * Cout * 2, (uint64_t)H * W * Cout * 2};
    uint32_t boxa[4] = {64, (uint32_t)W, (uint32_t)p.TR, 1}, es[4] = {1, 1, 1, 1};
    if (const char* e = make_tmap(&ta, dY, 4, da, sa, boxa, es)) return e;
    const int smem = stem_smem_bytes(true);
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(stem_wgrad_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
            cudaFuncSetAttribute(stem_wgrad_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
            cudaFuncSetAttribute(stem_wgrad_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return "cannot raise dynamic smem";
        attr = true;
    }
    const int grid = p.n_tiles < num_sms() ? p.n_tiles : num_sms();
    if (Cin == 1) stem_wgrad_tc_kernel<1><<<grid, kThreads, smem, stream>>>(ta, p);
    else if (Cin == 2) stem_wgrad_tc_kernel<2><<<grid, kThreads, smem, stream>>>(ta, p);
    else stem_wgrad_tc_kernel<3><<<grid, kThreads, smem, stream>>>(ta, p);
    const int n = Cout * 9 * Cin;
    stem_finish_kernel<<<(n + 255) / 256, 256, 0, stream>>>(scratch32, (__nv_bfloat16*)dW, n);
    return cudaGetLastError() == cudaSuccess ? nullptr : "stem_wgrad launch failed";
}

}  // namespace psb

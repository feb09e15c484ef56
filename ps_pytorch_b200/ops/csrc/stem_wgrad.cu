// Weight gradient of the network stem (3x3 / stride 1 / pad 1 convolution with Cin <= 4, e.g. RGB -> 64 channels).
//
// dW[co][r][s][c] = sum over 1 M pixels of dY[p,co] * X[p+(r,s)-1, c]: 64x27 outputs, a one-million-long reduction.  cuDNN runs this
// as a generic implicit GEMM (358 us on B200 for batch 1024, profiles/run3_conv.md) although it is ~4 GFLOP over 134 MB of dY.
// (With cudnn.benchmark cuDNN finds a 189 us algorithm; this kernel measures 199 us, so it is opt-in: PSB200_STEM_WGRAD=1.)
// Here: CUDA-core kernel, register-blocked.  A CTA walks 256-pixel tiles (8 image rows); the dY tile (32 KB) and the X halo
// (10 x 34 x float4) sit in smem; a thread owns 4 output channels x all 27 (tap, channel) pairs = 108 fp32 accumulators and
// visits every 16th pixel of the tile, so one pixel costs 1 LDS.64 + 9 LDS.128 for 108 FMAs.  Accumulators live across tiles; at the
// end they are folded in smem and added to an fp32 scratch with one atomic per output per CTA, then converted to bf16.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace psb {

constexpr int kSwThreads = 256;
constexpr int kSwRows = 8;              // image rows per tile

struct StemWgParams {
    const __nv_bfloat16* dy;            // [N,H,W,Cout]
    const __nv_bfloat16* x;             // [N,H,W,Cin]
    float* acc;                         // [Cout][9][Cin] fp32, zero on entry
    int N, H, W, Cin, Cout, tiles;      // tiles = N * H/8
};

__global__ void __launch_bounds__(kSwThreads, 1) stem_wgrad_kernel(StemWgParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int W = p.W, PIX = kSwRows * W;                        // pixels per tile (<= 256)
    __nv_bfloat16* dys = reinterpret_cast<__nv_bfloat16*>(smem);                 // [PIX][64]
    float4* xs = reinterpret_cast<float4*>(smem + PIX * 64 * 2);                  // [(kSwRows+2)][W+2]
    float* red = reinterpret_cast<float*>(xs + (kSwRows + 2) * (W + 2));          // [64*27]
    const int cg = threadIdx.x & 15, pg = threadIdx.x >> 4;      // 16 channel groups of 4, 16 pixel groups
    const int co_blk = blockIdx.y * 64;
    float a[4][27];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 27; ++k) a[i][k] = 0.f;
    const int tiles_per_img = p.H / kSwRows;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        const int n = tile / tiles_per_img, h0 = (tile % tiles_per_img) * kSwRows;
        __syncthreads();
        // dY tile: PIX x 64 channels of this CTA's channel block, 16-byte loads
        for (int v = threadIdx.x; v < PIX * 8; v += kSwThreads) {
            const int pix = v >> 3, ch8 = (v & 7) * 8;
            const long long g = (((long long)n * p.H + h0) * W + pix) * p.Cout + co_blk + ch8;
            *reinterpret_cast<uint4*>(dys + pix * 64 + ch8) = *reinterpret_cast<const uint4*>(p.dy + g);
        }
        // X halo as float4 per pixel (channels beyond Cin and pixels outside the image are zero)
        for (int v = threadIdx.x; v < (kSwRows + 2) * (W + 2); v += kSwThreads) {
            const int hh = v / (W + 2) + h0 - 1, ww = v % (W + 2) - 1;
            float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
            if (hh >= 0 && hh < p.H && ww >= 0 && ww < W) {
                const __nv_bfloat16* px = p.x + (((long long)n * p.H + hh) * W + ww) * p.Cin;
                f.x = __bfloat162float(px[0]);
                if (p.Cin > 1) f.y = __bfloat162float(px[1]);
                if (p.Cin > 2) f.z = __bfloat162float(px[2]);
                if (p.Cin > 3) f.w = __bfloat162float(px[3]);
            }
            xs[v] = f;
        }
        __syncthreads();
        for (int pix = pg; pix < PIX; pix += 16) {
            const int h = pix / W, w = pix - h * W;
            const uint2 d = *reinterpret_cast<const uint2*>(dys + pix * 64 + cg * 4);
            const float d0 = __uint_as_float(d.x << 16), d1 = __uint_as_float(d.x & 0xffff0000u);
            const float d2 = __uint_as_float(d.y << 16), d3 = __uint_as_float(d.y & 0xffff0000u);
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int s = 0; s < 3; ++s) {
                    const float4 xv = xs[(h + r) * (W + 2) + (w + s)];
                    const int k = (r * 3 + s) * 3;
                    a[0][k] = fmaf(d0, xv.x, a[0][k]); a[0][k + 1] = fmaf(d0, xv.y, a[0][k + 1]); a[0][k + 2] = fmaf(d0, xv.z, a[0][k + 2]);
                    a[1][k] = fmaf(d1, xv.x, a[1][k]); a[1][k + 1] = fmaf(d1, xv.y, a[1][k + 1]); a[1][k + 2] = fmaf(d1, xv.z, a[1][k + 2]);
                    a[2][k] = fmaf(d2, xv.x, a[2][k]); a[2][k + 1] = fmaf(d2, xv.y, a[2][k + 1]); a[2][k + 2] = fmaf(d2, xv.z, a[2][k + 2]);
                    a[3][k] = fmaf(d3, xv.x, a[3][k]); a[3][k + 1] = fmaf(d3, xv.y, a[3][k + 1]); a[3][k + 2] = fmaf(d3, xv.z, a[3][k + 2]);
                }
        }
    }
    // fold the 16 pixel groups in smem, then one global atomic per output per CTA
    __syncthreads();
    for (int v = threadIdx.x; v < 64 * 27; v += kSwThreads) red[v] = 0.f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 27; ++k) atomicAdd(red + (cg * 4 + i) * 27 + k, a[i][k]);
    __syncthreads();
    for (int v = threadIdx.x; v < 64 * 27; v += kSwThreads) {
        const int co = v / 27, k = v % 27, tap = k / 3, c = k % 3;
        if (c < p.Cin) atomicAdd(p.acc + ((long long)(co_blk + co) * 9 + tap) * p.Cin + c, red[v]);
    }
}

__global__ void stem_wgrad_finish_kernel(float* __restrict__ acc, __nv_bfloat16* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { out[i] = __float2bfloat16_rn(acc[i]); acc[i] = 0.f; }
}

const char* stem_wgrad_launch(const void* dy, const void* x, void* dw, float* scratch, int N, int H, int W, int Cin, int Cout, cudaStream_t stream) {
    if (Cin > 3 || Cout % 64 || H % kSwRows || W > 32 || W < 1) return "stem wgrad: needs Cin <= 3, Cout % 64 == 0, H % 8 == 0, W <= 32";
    StemWgParams p{(const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, scratch, N, H, W, Cin, Cout, N * (H / kSwRows)};
    const int smem = kSwRows * W * 64 * 2 + (kSwRows + 2) * (W + 2) * 16 + 64 * 27 * 4;
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(stem_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024); attr = true; }
    int sms = 148; int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    dim3 grid(p.tiles < 2 * sms ? p.tiles : 2 * sms, Cout / 64);
    stem_wgrad_kernel<<<grid, kSwThreads, smem, stream>>>(p);
    const int n = Cout * 9 * Cin;
    stem_wgrad_finish_kernel<<<(n + 255) / 256, 256, 0, stream>>>(scratch, (__nv_bfloat16*)dw, n);
    return cudaGetLastError() == cudaSuccess ? nullptr : "stem wgrad launch failed";
}

}  // namespace psb

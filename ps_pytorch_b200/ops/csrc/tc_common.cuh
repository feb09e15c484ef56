// Shared pieces of the tcgen05 kernels (gemm_tcgen05.cu, conv2_tcgen05.cu): tile constants, the parameter block of the GEMM / per-tap /
// halo kernels, the TMEM -> global epilogue, and the host helpers (tensor maps, error word, SM count).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "tc_ptx.cuh"

namespace psb {

using namespace tc;

constexpr int BM = 128;
constexpr int BK = 64;                 // 64 bf16 = one 128-byte swizzle row
constexpr int kThreads = 256;
constexpr int kEpiWarp0 = 4;

struct TcParams {
    int M, N, num_kb;
    int mode;                          // 0 GEMM, 1 implicit-GEMM conv
    int n_tiles_m, n_tiles_n;
    // conv geometry
    int Hout, Wout, NB, stride, pad, S, cblocks;
    int BW, BH, BNI, tiles_w, tiles_h;
    const __nv_bfloat16* bias;
    const __nv_bfloat16* residual;      // optional [M,N] tensor added in the epilogue (dgrad + skip-gradient fusion)
    int relu;
    int R, halo_w, a_stage_bytes;       // halo kernel: filter rows, patch pitch in pixels, bytes reserved per A stage
    int a_tx_bytes;                     // halo kernel: bytes one patch load actually delivers
    int b_resident, a_stages, b_stages; // weight-stationary mode: all filter rows of B stay in smem for the whole kernel (halo kernels)
    int desc_mode;
    __nv_bfloat16* C;
    int* err;                          // set to 1 if a barrier wait timed out (protocol bug), never hang
    float* stats;                      // optional BatchNorm partial statistics [gridDim.x * 4 warps][2][N] (see epilogue_tile)
};

// TMEM accumulator quarter -> (+bias) (+residual) (ReLU) -> bf16 -> global.  Thread `lane` of warp quarter q owns one output row.
// P is any parameter block with N, C, bias, residual, relu.  zero_acc: the tile had no MMAs (a dgrad parity class without taps): the
// accumulator is taken as 0 instead of being read.
// stats_row (optional): this warp's private row [2][N] of per-channel partial sums for the BatchNorm that follows the convolution:
// the 32 output rows a warp holds are reduced per column with a transposing butterfly (31 shuffles per quantity: lane i ends up with the
// total of column i) and added to the row with one coalesced RED per quantity — the separate bn_stats pass (one full read of the
// activation per BN layer) disappears.  Rows are private per warp, so the REDs never contend; bn_finalize sums the rows and re-zeroes them.
__device__ __forceinline__ void butterfly_colsum(float (&a)[32], int lane) {
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int j = 0; j < half; ++j) {
            const float keep = up ? a[j + half] : a[j];
            const float send = up ? a[j] : a[j + half];
            a[j] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
}

// For BN = 64 (the 64-channel layers, where a tile's MMAs take ~1200 cycles and the epilogue is close to the critical path) the butterfly
// per tile costs more than the statistics pass it replaces (measured: +0.06 ms/step).  There the per-column sums are instead carried in
// REGISTERS across all tiles of the persistent CTA (racc[0..BN) sums, racc[BN..2BN) sums of squares of this thread's rows: 2 FP32 ops per
// element) and reduced + written once per kernel by stats_flush.
template <int BN> struct StatsCfg { static constexpr bool kInRegs = BN <= 64; };

template <int BN>
__device__ __forceinline__ void stats_flush(float (&racc)[2 * BN], float* stats_row, int N, int col_base, int lane) {
#pragma unroll
    for (int c0 = 0; c0 < BN; c0 += 32) {
        float a[32], b[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) { a[j] = racc[c0 + j]; b[j] = racc[BN + c0 + j]; }
        butterfly_colsum(a, lane);
        butterfly_colsum(b, lane);
        if (col_base + c0 + lane < N) {
            atomicAdd(stats_row + col_base + c0 + lane, a[0]);
            atomicAdd(stats_row + N + col_base + c0 + lane, b[0]);
        }
    }
}

template <int BN, class P>
__device__ __forceinline__ void epilogue_tile(const P& p, uint32_t tmem_base, int acc, int q, int lane, long long out_row, int nt, bool zero_acc = false,
                                              float* stats_row = nullptr, float* racc = nullptr) {
    const bool vec_ok = (p.N % 8) == 0;
    __nv_bfloat16* crow = p.C + out_row * (long long)p.N + (long long)nt * BN;
    const __nv_bfloat16* rrow = p.residual ? p.residual + out_row * (long long)p.N + (long long)nt * BN : nullptr;
#pragma unroll (StatsCfg<BN>::kInRegs ? BN / 32 : 1)
    for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        if (!zero_acc) {
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), v);
            tmem_ld_wait();
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
        const int col0 = nt * BN + c0;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = 0.f;
        if (out_row >= 0 && col0 < p.N) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
            if (p.bias) {
#pragma unroll
                for (int j = 0; j < 32; ++j) if (col0 + j < p.N) f[j] += __bfloat162float(p.bias[col0 + j]);
            }
            const bool full = vec_ok && col0 + 32 <= p.N;
            if (rrow) {
                if (full) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        const uint4 rv = *reinterpret_cast<const uint4*>(rrow + c0 + j);
                        const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) { f[j + 2 * k] += __uint_as_float(w[k] << 16); f[j + 2 * k + 1] += __uint_as_float(w[k] & 0xffff0000u); }
                    }
                } else {
                    for (int j = 0; j < 32; ++j) if (col0 + j < p.N) f[j] += __bfloat162float(rrow[c0 + j]);
                }
            }
            if (p.relu) {
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            if (full && (p.N % 16) == 0) {
                // 256-bit stores (STG.E.ENL2.256): every thread writes whole 32-byte sectors of its output row
#pragma unroll
                for (int j = 0; j < 32; j += 16) {
                    uint32_t o[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        __nv_bfloat162 h = __floats2bfloat162_rn(f[j + 2 * k], f[j + 2 * k + 1]);
                        o[k] = *reinterpret_cast<uint32_t*>(&h);
                    }
                    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(crow + c0 + j), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]),
                                 "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
                }
            } else if (full) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    uint4 o;
                    __nv_bfloat162 h0 = __floats2bfloat162_rn(f[j], f[j + 1]), h1 = __floats2bfloat162_rn(f[j + 2], f[j + 3]);
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(f[j + 4], f[j + 5]), h3 = __floats2bfloat162_rn(f[j + 6], f[j + 7]);
                    o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
                    o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
                    *reinterpret_cast<uint4*>(crow + c0 + j) = o;
                }
            } else {
                for (int j = 0; j < 32; ++j) if (col0 + j < p.N) crow[c0 + j] = __float2bfloat16_rn(f[j]);
            }
        }
        if (StatsCfg<BN>::kInRegs && racc != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { racc[c0 + j] += f[j]; racc[BN + c0 + j] = fmaf(f[j], f[j], racc[BN + c0 + j]); }
        } else if (stats_row != nullptr && col0 < p.N) {   // warp-uniform; rows without an output pixel contribute zeros
            float sq[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) sq[j] = f[j] * f[j];
            butterfly_colsum(f, lane);
            butterfly_colsum(sq, lane);
            if (col0 + lane < p.N) {
                atomicAdd(stats_row + col0 + lane, f[0]);
                atomicAdd(stats_row + p.N + col0 + lane, sq[0]);
            }
        }
    }
}


// ---- host helpers (defined in gemm_tcgen05.cu)
// rank-`rank` bf16 tensor map with 128-byte swizzle; dims/strides innermost first (strides in BYTES, for dims 1..rank-1)
const char* make_tmap(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                      const uint32_t* estrides);
int* err_word();                        // one device word shared by all launches; checked by tc_check_error()
int num_sms();
int wgrad_sms();                        // SMs the weight-gradient kernels may use (tc_set_wgrad_sms)

}  // namespace psb

// Weight gradient of 3x3 / stride-1 / pad-1 convolutions with tap sharing ("halo" form), tcgen05 + TMA + TMEM.
//
//   dW[co][r][s][ci] = sum over output pixels p of dY[p, co] * X[p + (r-1, s-1), ci]
//
// The per-tap kernel (gemm_tcgen05.cu: tc_wgrad_kernel) fetches a [128 pixels x 64 ch] dY tile AND a tap-shifted X tile for every
// (tap, K block): 9x redundant traffic, and with 64-channel operands an MMA step needs 64 B/cycle/SM of shared-memory fill while L2->SM
// delivers ~40 — measured 237 us for the 64->64 @32x32 layer of ResNet-18 (cuDNN: 134 us), L2->SM bound (profiles/run25_conv_ncu.md).
//
// Here a K block is a rectangle of 128 output pixels (BNI images x BH rows x 8 columns, BH * BNI = 16).  Per K block ONE dY tile and
// ONE X patch WITH ITS HALO ((BH+2) rows x 16-pixel pitch per image, the same layout as the forward halo kernel) are loaded, and every
// filter tap reads the patch through a shifted MN-major window:
//   * K = 16 pixels per MMA = two 8-pixel runs in consecutive patch rows (SBO = 2048 = one patch row);
//   * the three taps s = 0,1,2 of a filter row are ONE instruction with N = 192: the 64-channel N groups are 128 bytes apart
//     (LBO = 128 = one pixel), i.e. the groups are the same patch shifted by one pixel each.
// D (fp32, TMEM) = [MM out-channels] x [192 columns per filter row].  Nine taps x 64 ci = 576 columns exceed the 512 TMEM columns, so an
// item covers a group of filter rows: group 0 = rows 0,1 (384 columns), group 1 = row 2 (192 columns); both groups stream the same
// pixels (2x instead of 9x traffic).  Items = (row group, co tile, ci tile, K split); partial tiles are added into an fp32
// [Cout][9][Cin] scratch with vector reductions and converted by wgrad_finish_kernel (shared with the per-tap kernel).
//
// Reference op replaced: the weight-gradient half of nn.Conv2d.backward -> cuDNN (/root/reference/src/model_ops/resnet.py:19-27).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "tc_common.cuh"
#include "pdl.cuh"

namespace psb {

struct WgHaloParams {
    int Cin, Cout, NB, H, W;
    int BH, BNI;                         // pixel rectangle: BNI images x BH rows x 8 columns (BH * BNI = 16)
    int tiles_w, tiles_h, tiles_n;       // rectangles per row / column / image groups
    int kb_total;                        // K blocks (pixel rectangles) in the whole activation
    int co_tiles, ci_tiles, ksplit, kb_per_split, n_items;
    int a_bytes, b_bytes, stages;        // bytes of the dY tile (all co groups) and of the X patch per stage
    float* dW32;
    int* err;
};

constexpr int kWhMaxStages = 4;

// MN-major SWIZZLE_128B operand with explicit strides: `lbo_bytes` between 64-element MN groups, `sbo_bytes` between 8-row K atoms
__device__ __forceinline__ uint64_t umma_desc_mn_sw128_ex(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

template <int MM>          // out-channel tile: 64 or 128
__global__ void __launch_bounds__(kThreads, 1)
tc_wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x, const WgHaloParams p) {
    constexpr int kGroupsA = MM / 64;
    constexpr int kABox = 128 * 128;                        // one [128 pixels][64 co] box
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_bytes = p.a_bytes + p.b_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
    uint64_t* empty_bar = full_bar + kWhMaxStages;
    uint64_t* tmem_full = empty_bar + kWhMaxStages;
    uint64_t* tmem_empty = tmem_full + 1;
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_dy); prefetch_tmap(&tmap_x); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
        mbar_init(tmem_full, 1); mbar_init(tmem_empty, 4);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_base_smem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_trigger(); pdl_wait();                  // pdl.cuh: prologue done; from here on the kernel touches global memory
    const uint32_t tmem_base = *tmem_base_smem;

    // item -> (row group fastest, then ci tile, co tile: the CTAs of one wave stream the same pixels; K split slowest)
    // (the row group alternates from round to round — item & 1 alone would give every even CTA only the heavy two-row items)
    auto decode = [&](int item, int& rg, int& nt, int& mt, int& ks) {
        rg = (item & 1) ^ ((item / (int)gridDim.x) & 1); int r = item >> 1;
        nt = r % p.ci_tiles; r /= p.ci_tiles;
        mt = r % p.co_tiles; ks = r / p.co_tiles;
    };

    if (warp == 0) {
        int stage = 0; uint32_t phase = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            int rg, nt, mt, ks; decode(item, rg, nt, mt, ks);
            const int kb0 = ks * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.kb_total);
            for (int kb = kb0; kb < kb1; ++kb) {
                const int tw = kb % p.tiles_w, th = (kb / p.tiles_w) % p.tiles_h, tn = kb / (p.tiles_w * p.tiles_h);
                const int n0 = tn * p.BNI, h0 = th * p.BH, w0 = tw * 8;
                if (!mbar_wait(empty_bar + stage, phase ^ 1)) { *p.err = 41; return; }
                uint8_t* sa = smem + stage * stage_bytes;
                uint8_t* sb = sa + p.a_bytes;
                if (elect_one()) {
                    mbar_arrive_expect_tx(full_bar + stage, stage_bytes);
#pragma unroll
                    for (int g = 0; g < kGroupsA; ++g)
                        tma_load_4d(sa + g * kABox, &tmap_dy, full_bar + stage, mt * MM + g * 64, w0, h0, n0);
                    tma_load_4d(sb, &tmap_x, full_bar + stage, nt * 64, w0 - 1, h0 - 1, n0);
                }
                __syncwarp();
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc_bf16(MM, 192, 1, 1);          // both operands MN-major, N = 3 taps x 64 ci
        const uint64_t a_tmpl = umma_desc_mn_sw128(0, kABox), b_tmpl = umma_desc_mn_sw128_ex(0, 128, 2048);
        int stage = 0; uint32_t phase = 0; uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            int rg, nt, mt, ks; decode(item, rg, nt, mt, ks);
            const int kb0 = ks * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.kb_total);
            const int r0 = rg == 0 ? 0 : 2, nrows = rg == 0 ? 2 : 1;
            if (!mbar_wait(tmem_empty, acc_phase ^ 1)) { *p.err = 42; return; }     // single accumulator: the previous item is drained
            tc_fence_after();
            for (int kb = kb0; kb < kb1; ++kb) {
                if (!mbar_wait(full_bar + stage, phase)) { *p.err = 43; return; }
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * stage_bytes);
                // descriptors advance by adding (byte offset >> 4) to the address field: one patch row = 2048 B = +128
                const uint64_t da = a_tmpl + (uint64_t)((sa & 0x3FFFFu) >> 4);
                const uint64_t db = b_tmpl + (uint64_t)(((sa + (uint32_t)p.a_bytes) & 0x3FFFFu) >> 4) + (uint64_t)(r0 * 128);
                const uint32_t acc0 = kb > kb0 ? 1u : 0u;
                if (elect_one()) {
                    if (p.BH == 16) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {                    // 16 pixels = rows 2j, 2j+1 of the 16-row rectangle (one image)
                            tc_mma_bf16(tmem_base, da + (uint64_t)(j * 128), db + (uint64_t)(2 * j * 128), idesc, j > 0 ? 1u : acc0);
                            if (nrows == 2) tc_mma_bf16(tmem_base + 192u, da + (uint64_t)(j * 128), db + (uint64_t)((2 * j + 1) * 128), idesc, j > 0 ? 1u : acc0);
                        }
                    } else {                                             // BH = 8: rows 8..15 belong to the second image, whose patch starts BH + 2 rows later
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int prow = 2 * j + (j >= 4 ? 2 : 0);
                            tc_mma_bf16(tmem_base, da + (uint64_t)(j * 128), db + (uint64_t)(prow * 128), idesc, j > 0 ? 1u : acc0);
                            if (nrows == 2) tc_mma_bf16(tmem_base + 192u, da + (uint64_t)(j * 128), db + (uint64_t)((prow + 1) * 128), idesc, j > 0 ? 1u : acc0);
                        }
                    }
                    tc_commit(empty_bar + stage);
                }
                __syncwarp();
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) tc_commit(tmem_full);
            __syncwarp();
            acc_phase ^= 1;
        }
    } else if (warp >= kEpiWarp0) {
        const int q = warp & 3;
        uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            int rg, nt, mt, ks; decode(item, rg, nt, mt, ks);
            const int r0 = rg == 0 ? 0 : 2, ncols = (rg == 0 ? 2 : 1) * 192;
            // TMEM rows: M=128 -> lane = row; M=64 -> row m lives in lane (m%16) + 32*(m/16)
            const int row = MM == 128 ? q * 32 + lane : q * 16 + lane;
            const bool row_ok = (MM == 128 || lane < 16) && (mt * MM + row) < p.Cout;
            const bool has_work = ks * p.kb_per_split < p.kb_total;
            if (!mbar_wait(tmem_full, acc_phase)) { if (lane == 0) *p.err = 44; return; }
            tc_fence_after();
            float* wrow = p.dW32 + (long long)(mt * MM + row) * 9 * p.Cin + (long long)nt * 64;
#pragma unroll 1
            for (int c0 = 0; c0 < ncols; c0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                tmem_ld_wait();
                if (row_ok && has_work) {
                    const int tap = (r0 + c0 / 192) * 3 + (c0 % 192) / 64, ci = c0 % 64;
                    float* d = wrow + (long long)tap * p.Cin + ci;
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(d + j), "f"(__uint_as_float(v[j])), "f"(__uint_as_float(v[j + 1])),
                                     "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3])) : "memory");
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_relaxed(tmem_empty);
            acc_phase ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

void wgrad_finish_launch(float* acc, void* out, long long n, cudaStream_t stream);      // gemm_tcgen05.cu

template <int MM>
static const char* launch_wgh(const CUtensorMap& ta, const CUtensorMap& tb, const WgHaloParams& p, int smem_bytes, cudaStream_t stream) {
    static int attr = 0;
    if (attr < smem_bytes) {
        if (cudaFuncSetAttribute(tc_wgrad_halo_kernel<MM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess) return "cannot raise dynamic smem";
        attr = smem_bytes;
    }
    int grid = p.n_items < wgrad_sms() ? p.n_items : wgrad_sms();
    grid &= ~1;                                       // even: the two row groups of a pixel range stay in the same round (decode)
    if (grid < 2) grid = 2;
    launch_pdl(tc_wgrad_halo_kernel<MM>, dim3(grid), dim3(kThreads), smem_bytes, stream, ta, tb, p);
    return cudaGetLastError() == cudaSuccess ? nullptr : "kernel launch failed";
}

bool tc_wgrad_halo_eligible(int Hin, int Win, int Cin, int Cout, int R, int S, int stride, int pad) {
    return R == 3 && S == 3 && stride == 1 && pad == 1 && Cin % 64 == 0 && Cout % 64 == 0 && Win >= 8 && Hin >= 8 && (Hin >= 16 || Hin == 8);
}

// dW (bf16 [Cout][3][3][Cin]) from dY [NB,H,W,Cout] and X [NB,H,W,Cin]; scratch32: zeroed fp32 [Cout*9*Cin], left zeroed
const char* tc_conv_wgrad_halo(const void* dY, const void* X, void* dW, float* scratch32, int NB, int Hin, int Win, int Cin, int Cout, cudaStream_t stream) {
    if (!tc_wgrad_halo_eligible(Hin, Win, Cin, Cout, 3, 3, 1, 1)) return "shape not eligible for the halo weight-gradient kernel";
    const int BH = Hin >= 16 ? 16 : 8, BNI = 16 / BH;
    const int MM = Cout % 128 == 0 ? 128 : 64;
    CUtensorMap ta, tb;
    uint64_t da[4] = {(uint64_t)Cout, (uint64_t)Win, (uint64_t)Hin, (uint64_t)NB};
    uint64_t sa[3] = {(uint64_t)Cout * 2, (uint64_t)Win * Cout * 2, (uint64_t)Hin * Win * Cout * 2};
    uint32_t boxa[4] = {64, 8, (uint32_t)BH, (uint32_t)BNI}, es[4] = {1, 1, 1, 1};
    uint64_t db[4] = {(uint64_t)Cin, (uint64_t)Win, (uint64_t)Hin, (uint64_t)NB};
    uint64_t sb[3] = {(uint64_t)Cin * 2, (uint64_t)Win * Cin * 2, (uint64_t)Hin * Win * Cin * 2};
    uint32_t boxb[4] = {64, 16, (uint32_t)(BH + 2), (uint32_t)BNI};
    if (const char* e = make_tmap(&ta, dY, 4, da, sa, boxa, es)) return e;
    if (const char* e = make_tmap(&tb, X, 4, db, sb, boxb, es)) return e;
    WgHaloParams p{};
    p.Cin = Cin; p.Cout = Cout; p.NB = NB; p.H = Hin; p.W = Win; p.BH = BH; p.BNI = BNI;
    p.tiles_w = (Win + 7) / 8; p.tiles_h = (Hin + BH - 1) / BH; p.tiles_n = (NB + BNI - 1) / BNI;
    p.kb_total = p.tiles_w * p.tiles_h * p.tiles_n;
    p.co_tiles = Cout / MM; p.ci_tiles = Cin / 64;
    const int base_items = 2 * p.co_tiles * p.ci_tiles;
    int ksplit = (2 * wgrad_sms() + base_items - 1) / base_items;
    if (ksplit < 1) ksplit = 1;
    if (ksplit > p.kb_total) ksplit = p.kb_total;
    p.kb_per_split = (p.kb_total + ksplit - 1) / ksplit;
    p.ksplit = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
    p.n_items = base_items * p.ksplit;
    p.a_bytes = (MM / 64) * 128 * 128;
    p.b_bytes = BNI * (BH + 2) * 16 * 128;
    p.stages = (220 * 1024) / (p.a_bytes + p.b_bytes);
    if (p.stages > kWhMaxStages) p.stages = kWhMaxStages;
    if (p.stages < 2) return "stage does not fit in shared memory";
    p.dW32 = scratch32; p.err = err_word();
    const int smem_bytes = p.stages * (p.a_bytes + p.b_bytes) + 1024 + 256;
    const char* e = MM == 128 ? launch_wgh<128>(ta, tb, p, smem_bytes, stream) : launch_wgh<64>(ta, tb, p, smem_bytes, stream);
    if (e) return e;
    wgrad_finish_launch(scratch32, dW, (long long)Cout * 9 * Cin, stream);
    return cudaGetLastError() == cudaSuccess ? nullptr : "finish kernel launch failed";
}

}  // namespace psb

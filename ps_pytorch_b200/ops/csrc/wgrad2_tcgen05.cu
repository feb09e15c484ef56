// CTA-pair weight gradient for wide layers (Cout % 256 == 0, Cin % 128 == 0): tcgen05.mma.cta_group::2, one 256 (co) x BN2 (ci) tile
// of one filter tap per pair of SMs, both operands MN-major (the reduction index is the output PIXEL; channels are the contiguous
// dimension of NHWC tensors).
//
// Why: with MN-major operands the tensor core fetches ~40-45 B/cycle/SM from shared memory (measured on the 1-CTA kernel: a
// 128 x 256 x 16 MMA reads 4 KB of dY + 8 KB of X and takes ~270 cycles for a 128-cycle floor, profiles/run25_conv_ncu.md: tensor pipe
// 48 % busy on the 512-channel layer).  In a pair each SM reads its own 128 out-channels of dY and only HALF of the X tile:
// 8 KB instead of 12 KB per instruction, and L2 -> SM traffic per FLOP drops by the same third.
//   rank r loads dY columns [256 mt + 128 r, +128) and X columns [BN2 nt + (BN2/2) r, +BN2/2) of every K block (KP pixels = a rectangle
//   of the output map; stride-2 / tap shifts / padding through the 4-D tensor maps exactly as in tc_wgrad_kernel); every TMA signals the
//   leader's barrier; the leader issues the MMAs; commits are multicast; each CTA's epilogue warps add their 128 rows of the partial
//   tile into the fp32 scratch with vector reductions (split-K).
//
// Reference op replaced: the weight-gradient half of nn.Conv2d.backward -> cuDNN (/root/reference/src/model_ops/resnet.py:19-27).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "tc_common.cuh"
#include "pdl.cuh"

namespace psb {

struct Wg2Params {
    int taps, S, stride, pad, Cin, Cout, NB, Ho, Wo;
    int BW, BH, BNI, tiles_w, tiles_h, kb_total;
    int co_tiles, ci_tiles, ksplit, kb_per_split, n_items;
    float* dW32;
    int* err;
};

constexpr int kWg2KP = 64;                         // pixels per K block
constexpr int kWg2Box = kWg2KP * 128;              // one [64 pixels][64 channels] box

template <int BN2> struct Wg2Cfg {
    static constexpr int kGroupsA = 2;                       // this CTA's 128 out-channels
    static constexpr int kGroupsB = BN2 / 128;               // this CTA's BN2/2 in-channels
    static constexpr int kStageBytes = (kGroupsA + kGroupsB) * kWg2Box;
    static constexpr int kStages = (200 * 1024) / kStageBytes > 8 ? 8 : (200 * 1024) / kStageBytes;
    static constexpr int kTmemCols = 2 * BN2;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

template <int BN2>
__global__ void __launch_bounds__(kThreads, 1)
tc_wgrad2_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x, const Wg2Params p) {
    using Cfg = Wg2Cfg<BN2>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
    uint64_t* empty_bar = full_bar + Cfg::kStages;
    uint64_t* tmem_full = empty_bar + Cfg::kStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_dy); prefetch_tmap(&tmap_x); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tmem_full + s, 1); mbar_init(tmem_empty + s, 8); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc2(tmem_base_smem, Cfg::kTmemCols);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    pdl_trigger(); pdl_wait();                  // pdl.cuh: prologue done; from here on the kernel touches global memory
    const uint32_t tmem_base = *tmem_base_smem;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    // item -> (tap fastest: the pairs of one wave read the same dY / X region, L2-friendly), then K split, ci tile, co tile
    auto decode = [&](int item, int& tap, int& ks, int& nt, int& mt) {
        tap = item % p.taps; int r = item / p.taps;
        ks = r % p.ksplit; r /= p.ksplit;
        nt = r % p.ci_tiles; mt = r / p.ci_tiles;
    };

    if (warp == 0) {
        int stage = 0; uint32_t phase = 0;
        for (int item = cluster_id; item < p.n_items; item += n_clusters) {
            int tap, ks, nt, mt; decode(item, tap, ks, nt, mt);
            const int fr = tap / p.S, fs = tap - fr * p.S;
            const int kb0 = ks * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.kb_total);
            for (int kb = kb0; kb < kb1; ++kb) {
                const int tw = kb % p.tiles_w, th = (kb / p.tiles_w) % p.tiles_h, tn = kb / (p.tiles_w * p.tiles_h);
                const int n0 = tn * p.BNI, h0 = th * p.BH, w0 = tw * p.BW;
                if (!mbar_wait(empty_bar + stage, phase ^ 1)) { *p.err = 51; goto teardown; }
                uint8_t* sa = smem + stage * Cfg::kStageBytes;
                uint8_t* sb = sa + Cfg::kGroupsA * kWg2Box;
                if (elect_one()) {
                    const uint32_t lead_full = mapa_u32(smem_u32(full_bar + stage), 0);
                    if (leader) mbar_arrive_expect_tx(full_bar + stage, 2 * Cfg::kStageBytes);
#pragma unroll
                    for (int g = 0; g < Cfg::kGroupsA; ++g)
                        tma_load_4d_2sm(sa + g * kWg2Box, &tmap_dy, lead_full, mt * 256 + (int)rank * 128 + g * 64, w0, h0, n0);
#pragma unroll
                    for (int g = 0; g < Cfg::kGroupsB; ++g)
                        tma_load_4d_2sm(sb + g * kWg2Box, &tmap_x, lead_full, nt * BN2 + (int)rank * (BN2 / 2) + g * 64, w0 * p.stride + fs - p.pad,
                                        h0 * p.stride + fr - p.pad, n0);
                }
                __syncwarp();
                if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && leader) {
        constexpr uint32_t idesc = umma_idesc_bf16(256, BN2, 1, 1);          // both operands MN-major
        const uint64_t tmpl = umma_desc_mn_sw128(0, kWg2Box);
        int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
        for (int item = cluster_id; item < p.n_items; item += n_clusters) {
            int tap, ks, nt, mt; decode(item, tap, ks, nt, mt);
            const int kb0 = ks * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.kb_total);
            if (!mbar_wait(tmem_empty + acc, acc_phase ^ 1)) { *p.err = 52; goto teardown; }
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * BN2;
            for (int kb = kb0; kb < kb1; ++kb) {
                if (!mbar_wait(full_bar + stage, phase)) { *p.err = 53; goto teardown; }
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
                const uint64_t da = tmpl + (uint64_t)((sa & 0x3FFFFu) >> 4), db = da + (uint64_t)((Cfg::kGroupsA * kWg2Box) >> 4);
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < kWg2KP / 16; ++j)            // 16 pixels per MMA = two 8-row atoms = 2048 bytes (address field += 128)
                        tc_mma2_bf16(tmem_d, da + (uint64_t)(j * 128), db + (uint64_t)(j * 128), idesc, j > 0 ? 1u : (uint32_t)(kb > kb0));
                    tc_commit2_mcast(empty_bar + stage, 0b11);
                }
                __syncwarp();
                if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) tc_commit2_mcast(tmem_full + acc, 0b11);
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= kEpiWarp0) {
        const int q = warp & 3;
        int acc = 0; uint32_t acc_phase = 0;
        for (int item = cluster_id; item < p.n_items; item += n_clusters) {
            int tap, ks, nt, mt; decode(item, tap, ks, nt, mt);
            const int row = mt * 256 + (int)rank * 128 + q * 32 + lane;          // this thread's out-channel
            const bool has_work = ks * p.kb_per_split < p.kb_total;
            if (!mbar_wait(tmem_full + acc, acc_phase)) { if (lane == 0) *p.err = 54; goto teardown; }
            tc_fence_after();
            float* drow = p.dW32 + ((long long)row * p.taps + tap) * p.Cin + (long long)nt * BN2;
#pragma unroll 1
            for (int c0 = 0; c0 < BN2; c0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN2 + c0), v);
                tmem_ld_wait();
                if (row < p.Cout && has_work) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(drow + c0 + j), "f"(__uint_as_float(v[j])), "f"(__uint_as_float(v[j + 1])),
                                     "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3])) : "memory");
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(tmem_empty + acc), 0));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
teardown:
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    if (warp == 2) tmem_dealloc2(tmem_base, Cfg::kTmemCols);
}

void wgrad_finish_launch(float* acc, void* out, long long n, cudaStream_t stream);      // gemm_tcgen05.cu

template <int BN2>
static const char* launch_wg2(const CUtensorMap& ta, const CUtensorMap& tb, const Wg2Params& p, cudaStream_t stream) {
    using Cfg = Wg2Cfg<BN2>;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(tc_wgrad2_kernel<BN2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess) return "cannot raise dynamic smem";
        attr = true;
    }
    int pairs = wgrad_sms() / 2;
    if (p.n_items < pairs) pairs = p.n_items;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = Cfg::kSmemBytes; cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_attr(at, 1);
    return cudaLaunchKernelEx(&cfg, tc_wgrad2_kernel<BN2>, ta, tb, p) == cudaSuccess ? nullptr : "kernel launch failed";
}

// Cin % 256: with a 128-wide N tile the pair kernel measured slower than the 1-CTA kernel (72.7 vs 66.8 us on 128->256/s2, profiles/run25_conv_per_layer.md)
bool tc_wgrad2_eligible(int Cin, int Cout) { return Cout % 256 == 0 && Cin % 256 == 0; }

const char* tc_conv_wgrad2(const void* dY, const void* X, void* dW, float* scratch32, int NB, int Hin, int Win, int Cin, int Cout, int R, int S,
                           int stride, int pad, cudaStream_t stream) {
    if (!tc_wgrad2_eligible(Cin, Cout)) return "shape not eligible for the CTA-pair weight-gradient kernel";
    const int Ho = (Hin + 2 * pad - R) / stride + 1, Wo = (Win + 2 * pad - S) / stride + 1;
    const int BN2 = Cin % 256 == 0 ? 256 : 128;
    int BW = 1; while (BW * 2 <= Wo && BW * 2 <= 32) BW *= 2;
    int BH = 1; while (BH * 2 <= Ho && BW * BH * 2 <= kWg2KP) BH *= 2;
    const int BNI = kWg2KP / (BW * BH);
    if ((BW - 1) * stride + 1 > 256 || (BH - 1) * stride + 1 > 256 || BNI > 256) return "pixel rectangle exceeds the TMA box limit";
    CUtensorMap ta, tb;
    uint64_t da[4] = {(uint64_t)Cout, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)NB};
    uint64_t sa[3] = {(uint64_t)Cout * 2, (uint64_t)Wo * Cout * 2, (uint64_t)Ho * Wo * Cout * 2};
    uint32_t boxa[4] = {64, (uint32_t)BW, (uint32_t)BH, (uint32_t)BNI}, esa[4] = {1, 1, 1, 1};
    uint64_t db[4] = {(uint64_t)Cin, (uint64_t)Win, (uint64_t)Hin, (uint64_t)NB};
    uint64_t sb[3] = {(uint64_t)Cin * 2, (uint64_t)Win * Cin * 2, (uint64_t)Hin * Win * Cin * 2};
    uint32_t boxb[4] = {64, (uint32_t)((BW - 1) * stride + 1), (uint32_t)((BH - 1) * stride + 1), (uint32_t)BNI};
    uint32_t esb[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
    if (const char* e = make_tmap(&ta, dY, 4, da, sa, boxa, esa)) return e;
    if (const char* e = make_tmap(&tb, X, 4, db, sb, boxb, esb)) return e;
    Wg2Params p{};
    p.taps = R * S; p.S = S; p.stride = stride; p.pad = pad; p.Cin = Cin; p.Cout = Cout; p.NB = NB; p.Ho = Ho; p.Wo = Wo;
    p.BW = BW; p.BH = BH; p.BNI = BNI;
    p.tiles_w = (Wo + BW - 1) / BW; p.tiles_h = (Ho + BH - 1) / BH;
    p.kb_total = p.tiles_w * p.tiles_h * ((NB + BNI - 1) / BNI);
    p.co_tiles = Cout / 256; p.ci_tiles = Cin / BN2;
    const int base_items = p.taps * p.co_tiles * p.ci_tiles;
    const int pairs = wgrad_sms() / 2;
    int ksplit = (2 * pairs + base_items - 1) / base_items;
    if (ksplit < 1) ksplit = 1;
    if (ksplit > p.kb_total) ksplit = p.kb_total;
    p.kb_per_split = (p.kb_total + ksplit - 1) / ksplit;
    p.ksplit = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
    p.n_items = base_items * p.ksplit;
    p.dW32 = scratch32; p.err = err_word();
    const char* e = BN2 == 256 ? launch_wg2<256>(ta, tb, p, stream) : launch_wg2<128>(ta, tb, p, stream);
    if (e) return e;
    wgrad_finish_launch(scratch32, dW, (long long)Cout * R * S * Cin, stream);
    return cudaGetLastError() == cudaSuccess ? nullptr : "finish kernel launch failed";
}

}  // namespace psb

// CTA-pair implicit-GEMM convolution (tcgen05.mma.cta_group::2, M = 2 x 128 output pixels, N = BN channels) driven by a tap table.
//
// One kernel covers every pass that is "a sum over filter taps of (a shifted pixel rectangle of A) x (a K-slice of B)":
//   fprop  (any stride)   Y[n,oh,ow,co]  = sum_{r,s,ci} X[n, oh*st + r - pad, ow*st + s - pad, ci] * W[co][r][s][ci]
//   dgrad  (stride 1)     dX[n,h,w,ci]   = sum_{r,s,co} dY[n, h + pad - r, w + pad - s, co]        * Wt[ci][r][s][co]
//   dgrad  (stride st>1)  the output pixels split into st*st PARITY CLASSES (h % st, w % st); class (ph,pw) only sees the taps with
//                         (ph + pad - r) % st == 0 and (pw + pad - s) % st == 0 and is a dense stride-1 problem on the dY grid:
//                         dX[n, st*a + ph, st*b + pw, ci] = sum_{taps of the class, co} dY[n, a + dh(r), b + dw(s), co] * Wt[ci][r][s][co]
//                         (dh = (ph + pad - r)/st).  No zero-insertion, no wasted MMAs: the strided dgrad costs exactly the FLOPs of the
//                         forward pass (cuDNN's strided-dgrad kernels are the most expensive launches of the ResNet-18 step,
//                         profiles/run23_launch_list_final.md).  A class without taps (1x1/s2: three of four) just writes zeros (+residual).
//
// Mechanics (validated by tc_gemm2_kernel / tc_conv_halo2_kernel): the two CTAs of a cluster take neighbouring 128-pixel M tiles of the
// same N tile; per (tap, 64-channel block) each loads its own A rectangle (one 4-D TMA box; padding = TMA zero fill; fprop strides =
// tensor-map element strides) and HALF of the B tile (BN/2 rows); every TMA signals the leader's barrier; the leader issues
// tcgen05.mma.cta_group::2 (256 x BN x 16); commits are multicast to both CTAs; each CTA's four epilogue warps drain their own 128 TMEM
// lanes.  Per SM and MMA the tensor core reads 4 KB of A and BN/2 x 32 B of B from shared memory and L2 delivers (128 + BN/2) x 128 B per
// K block: half the per-FLOP smem and L2 traffic of the 1-CTA 128 x BN kernel, which is what bounded the deep layers.
//
// Reference ops replaced: nn.Conv2d forward/backward -> cuDNN (/root/reference/src/model_ops/resnet.py:19-27,72, vgg.py:53).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "tc_common.cuh"
#include "pdl.cuh"
#include "common.cuh"

namespace psb {

constexpr int kMaxCls = 4, kMaxTaps = 9;

struct ConvTab {
    int ncls;
    int ntaps[kMaxCls];
    int oh0[kMaxCls], ow0[kMaxCls];                     // output pixel = grid pixel * out_mul + (oh0, ow0)
    int dh[kMaxCls][kMaxTaps], dw[kMaxCls][kMaxTaps];   // A rectangle origin = grid origin * in_mul + (dh, dw)
    int kidx[kMaxCls][kMaxTaps];                        // B column block = kidx * (64 * cblocks) + 64 * cb
};

struct Conv2Params {
    int N;                       // output channels of this pass (GEMM N)
    int cblocks;                 // reduction channels / 64
    int n_tiles_n;
    int tiles_per_cls;           // 128-pixel M tiles of one class
    int BW, BH, BNI, tiles_w, tiles_h;
    int Hg, Wg, NB;              // the M grid (fprop: output map; dgrad: dY-sized grid of one parity class)
    int in_mul, out_mul;
    int Hout, Wout;              // output tensor
    ConvTab tab;
    const __nv_bfloat16* bias;
    const __nv_bfloat16* residual;
    int relu;
    __nv_bfloat16* C;
    int* err;
    float* stats;                // optional BatchNorm partial statistics [gridDim.x * 4][2][N] (tc_common.cuh: epilogue_tile)
};

template <int BN> struct Conv2Cfg {
    static constexpr int kABytes = BM * BK * 2;                   // this CTA's 128 pixels x 64 channels
    static constexpr int kBBytes = (BN / 2) * BK * 2;             // this CTA's half of the B tile
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = BN == 256 ? 6 : 8;
    static constexpr int kTmemCols = 2 * BN;                      // two accumulator stages
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
tc_conv2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ Conv2Params p) {
    using Cfg = Conv2Cfg<BN>;
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
    if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_a); prefetch_tmap(&tmap_b); }
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

    // pair tile t -> (class, pair of M tiles, N tile); N tile fastest so that the pairs of one wave share their A rectangles through L2
    const int pairs_per_cls = (p.tiles_per_cls + 1) >> 1;
    const int n_ptiles = p.tab.ncls * pairs_per_cls * p.n_tiles_n;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int kred = p.cblocks * BK;

    if (warp == 0) {
        // ===================================================== TMA producer (both CTAs)
        int stage = 0; uint32_t phase = 0;
        for (int t = cluster_id; t < n_ptiles; t += n_clusters) {
            const int nt = t % p.n_tiles_n; const int rest = t / p.n_tiles_n;
            const int cls = rest / pairs_per_cls, mt = 2 * (rest % pairs_per_cls) + (int)rank;
            const bool ghost = mt >= p.tiles_per_cls;                     // odd tile count: the second half of the last pair
            const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tn = mt / (p.tiles_w * p.tiles_h);
            const int n0 = ghost ? p.NB : tn * p.BNI, h0 = th * p.BH * p.in_mul, w0 = tw * p.BW * p.in_mul;
            const int ntaps = p.tab.ntaps[cls];
            for (int tap = 0; tap < ntaps; ++tap) {
                const int ah = h0 + p.tab.dh[cls][tap], aw = w0 + p.tab.dw[cls][tap], kb0 = p.tab.kidx[cls][tap] * kred;
                for (int cb = 0; cb < p.cblocks; ++cb) {
                    if (!mbar_wait(empty_bar + stage, phase ^ 1)) { *p.err = 31; goto teardown; }
                    uint8_t* sa = smem + stage * Cfg::kStageBytes;
                    uint8_t* sb = sa + Cfg::kABytes;
                    if (elect_one()) {
                        const uint32_t lead_full = mapa_u32(smem_u32(full_bar + stage), 0);
                        if (leader) mbar_arrive_expect_tx(full_bar + stage, 2 * Cfg::kStageBytes);      // both CTAs' bytes land on this barrier
                        tma_load_4d_2sm(sa, &tmap_a, lead_full, cb * BK, aw, ah, n0);
                        tma_load_2d_2sm(sb, &tmap_b, lead_full, kb0 + cb * BK, nt * BN + (int)rank * (BN / 2));
                    }
                    __syncwarp();
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1 && leader) {
        // ===================================================== MMA issuer (leader CTA only)
        constexpr uint32_t idesc = umma_idesc_bf16(256, BN, 0, 0);
        const uint64_t tmpl = umma_desc_k_sw128(0);          // descriptors are advanced by adding (byte offset >> 4) to the address field
        int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
        for (int t = cluster_id; t < n_ptiles; t += n_clusters) {
            const int cls = (t / p.n_tiles_n) / pairs_per_cls;
            const int num_kb = p.tab.ntaps[cls] * p.cblocks;
            if (!mbar_wait(tmem_empty + acc, acc_phase ^ 1)) { *p.err = 32; goto teardown; }
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * BN;
            for (int kb = 0; kb < num_kb; ++kb) {
                if (!mbar_wait(full_bar + stage, phase)) { *p.err = 33; goto teardown; }
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
                const uint64_t da = tmpl + (uint64_t)((sa & 0x3FFFFu) >> 4), db = da + (uint64_t)(Cfg::kABytes >> 4);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        tc_mma2_bf16(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, k != 0 ? 1u : (uint32_t)(kb != 0));
                    tc_commit2_mcast(empty_bar + stage, 0b11);       // the stage is free in BOTH CTAs once these MMAs retire
                }
                __syncwarp();
                if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) tc_commit2_mcast(tmem_full + acc, 0b11);      // (a class without taps: nothing pending, arrives at once)
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= kEpiWarp0) {
        // ===================================================== epilogue (both CTAs; each drains its own 128 TMEM lanes)
        const int q = warp & 3;
        int acc = 0; uint32_t acc_phase = 0;
        float* const srow = p.stats ? p.stats + (size_t)(blockIdx.x * 4 + q) * 2 * p.N : nullptr;
        constexpr bool kRegStats = StatsCfg<BN>::kInRegs;
        const bool use_racc = kRegStats && srow != nullptr && p.n_tiles_n == 1;      // one N tile: a thread's columns never change
        float racc[kRegStats ? 2 * BN : 2];
#pragma unroll
        for (int j = 0; j < (kRegStats ? 2 * BN : 2); ++j) racc[j] = 0.f;
        for (int t = cluster_id; t < n_ptiles; t += n_clusters) {
            const int nt = t % p.n_tiles_n; const int rest = t / p.n_tiles_n;
            const int cls = rest / pairs_per_cls, mt = 2 * (rest % pairs_per_cls) + (int)rank;
            const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tn = mt / (p.tiles_w * p.tiles_h);
            const int row = q * 32 + lane;
            const int bw = row % p.BW, bh = (row / p.BW) % p.BH, bn = row / (p.BW * p.BH);
            const int n = tn * p.BNI + bn, hg = th * p.BH + bh, wg = tw * p.BW + bw;
            long long out_row = -1;
            if (mt < p.tiles_per_cls && n < p.NB && hg < p.Hg && wg < p.Wg) {
                const int oh = hg * p.out_mul + p.tab.oh0[cls], ow = wg * p.out_mul + p.tab.ow0[cls];
                if (oh < p.Hout && ow < p.Wout) out_row = ((long long)n * p.Hout + oh) * p.Wout + ow;
            }
            if (!mbar_wait(tmem_full + acc, acc_phase)) { if (lane == 0) *p.err = 34; goto teardown; }
            tc_fence_after();
            epilogue_tile<BN>(p, tmem_base, acc, q, lane, out_row, nt, p.tab.ntaps[cls] == 0, srow, use_racc ? racc : nullptr);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(tmem_empty + acc), 0));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if constexpr (kRegStats) { if (use_racc) stats_flush<BN>(racc, srow, p.N, 0, lane); }
    }
teardown:
    // every thread of both CTAs gets here (also after a timed-out wait): nobody may still be signalling the peer or using TMEM
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    if (warp == 2) tmem_dealloc2(tmem_base, Cfg::kTmemCols);
}

template <int BN>
static const char* launch_conv2(const CUtensorMap& ta, const CUtensorMap& tb, const Conv2Params& p, cudaStream_t stream) {
    using Cfg = Conv2Cfg<BN>;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(tc_conv2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess) return "cannot raise dynamic smem";
        attr = true;
    }
    const int ptiles = p.tab.ncls * ((p.tiles_per_cls + 1) / 2) * p.n_tiles_n;
    int pairs = num_sms() / 2;
    if (ptiles < pairs) pairs = ptiles;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = Cfg::kSmemBytes; cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_attr(at, 1);
    return cudaLaunchKernelEx(&cfg, tc_conv2_kernel<BN>, ta, tb, p) == cudaSuccess ? nullptr : "kernel launch failed";
}

static int g_conv2_bn = -1;          // -1 auto; 64 / 128 / 256 force the N tile (benchmarks)
void tc_set_conv2_bn(int b) { g_conv2_bn = b; }

// A: NHWC activations [NB, Ha, Wa, Cred] (fprop: X; dgrad: dY).  Wt: [N][R*S][Cred] K-major (fprop: the parameter itself, N = Cout,
// Cred = Cin; dgrad: the transposed parameter [Cin][R][S][Cout], N = Cin, Cred = Cout).  Y: NHWC [NB, Hout, Wout, N].
// dgrad = 0: Hout/Wout are derived.  dgrad = 1: Hout/Wout are the INPUT map of the forward convolution and must be given.
const char* tc_conv2(const void* A, const void* Wt, void* Y, const void* bias, const void* residual, int NB, int Ha, int Wa, int Cred, int N,
                     int R, int S, int stride, int pad, int dgrad, int Hout, int Wout, int relu, float* stats, cudaStream_t stream) {
    if (Cred % 64) return "reduction channels must be a multiple of 64";
    if (N % 64) return "output channels must be a multiple of 64";
    if (stride < 1 || stride > 2) return "stride must be 1 or 2";
    if (R * S > kMaxTaps) return "filter larger than 9 taps";
    Conv2Params p{};
    ConvTab& tb_ = p.tab;
    if (!dgrad) {
        Hout = (Ha + 2 * pad - R) / stride + 1; Wout = (Wa + 2 * pad - S) / stride + 1;
        tb_.ncls = 1; tb_.ntaps[0] = R * S; tb_.oh0[0] = tb_.ow0[0] = 0;
        for (int r = 0; r < R; ++r) for (int s = 0; s < S; ++s) { const int t = r * S + s; tb_.dh[0][t] = r - pad; tb_.dw[0][t] = s - pad; tb_.kidx[0][t] = t; }
        p.Hg = Hout; p.Wg = Wout; p.in_mul = stride; p.out_mul = 1;
    } else {
        if (Hout % stride || Wout % stride) return "strided dgrad needs an input map divisible by the stride";
        if ((Hout + 2 * pad - R) / stride + 1 != Ha || (Wout + 2 * pad - S) / stride + 1 != Wa) return "dY shape does not match the convolution geometry";
        // parity classes, heaviest first (the static round-robin schedule then ends on the cheap tiles)
        int order[kMaxCls], cnt[kMaxCls], nc = 0;
        for (int ph = 0; ph < stride; ++ph) for (int pw = 0; pw < stride; ++pw) {
            int c = 0;
            for (int r = 0; r < R; ++r) for (int s = 0; s < S; ++s)
                if (((ph + pad - r) % stride + stride) % stride == 0 && ((pw + pad - s) % stride + stride) % stride == 0) ++c;
            order[nc] = ph * stride + pw; cnt[nc] = c; ++nc;
        }
        for (int i = 0; i < nc; ++i) for (int j = i + 1; j < nc; ++j) if (cnt[j] > cnt[i]) { int t = cnt[i]; cnt[i] = cnt[j]; cnt[j] = t; t = order[i]; order[i] = order[j]; order[j] = t; }
        tb_.ncls = nc;
        for (int c = 0; c < nc; ++c) {
            const int ph = order[c] / stride, pw = order[c] % stride;
            int t = 0;
            for (int r = 0; r < R; ++r) for (int s = 0; s < S; ++s) {
                const int a = ph + pad - r, b = pw + pad - s;
                if (((a % stride) + stride) % stride || ((b % stride) + stride) % stride) continue;
                tb_.dh[c][t] = a >= 0 ? a / stride : -((-a) / stride); tb_.dw[c][t] = b >= 0 ? b / stride : -((-b) / stride); tb_.kidx[c][t] = r * S + s; ++t;
            }
            tb_.ntaps[c] = t; tb_.oh0[c] = ph; tb_.ow0[c] = pw;
        }
        p.Hg = Hout / stride; p.Wg = Wout / stride; p.in_mul = 1; p.out_mul = stride;
    }
    // tile rectangle in the M grid: BW = largest power of two <= min(Wg, 32), then rows, then images; product 128
    int BW = 1; while (BW * 2 <= p.Wg && BW * 2 <= 32) BW *= 2;
    int BH = 1; while (BH * 2 <= p.Hg && BW * BH * 2 <= 128) BH *= 2;
    const int BNI = 128 / (BW * BH);
    const int es = dgrad ? 1 : stride;
    if ((BW - 1) * es + 1 > 256 || (BH - 1) * es + 1 > 256 || BNI > 256) return "tile rectangle exceeds the TMA box limit";
    int BN = N % 256 == 0 ? 256 : (N % 128 == 0 ? 128 : 64);
    if (g_conv2_bn == 64 || (g_conv2_bn == 128 && N % 128 == 0) || (g_conv2_bn == 256 && N % 256 == 0)) BN = g_conv2_bn;
    CUtensorMap ta, tb;
    uint64_t da[4] = {(uint64_t)Cred, (uint64_t)Wa, (uint64_t)Ha, (uint64_t)NB};
    uint64_t sa[3] = {(uint64_t)Cred * 2, (uint64_t)Wa * Cred * 2, (uint64_t)Ha * Wa * Cred * 2};
    uint32_t boxa[4] = {BK, (uint32_t)((BW - 1) * es + 1), (uint32_t)((BH - 1) * es + 1), (uint32_t)BNI};
    uint32_t esa[4] = {1, (uint32_t)es, (uint32_t)es, 1};
    const int Ktot = R * S * Cred;
    uint64_t db[2] = {(uint64_t)Ktot, (uint64_t)N}, sb[1] = {(uint64_t)Ktot * 2};
    uint32_t boxb[2] = {BK, (uint32_t)(BN / 2)}, esb[2] = {1, 1};
    if (const char* e = make_tmap(&ta, A, 4, da, sa, boxa, esa)) return e;
    if (const char* e = make_tmap(&tb, Wt, 2, db, sb, boxb, esb)) return e;
    p.N = N; p.cblocks = Cred / BK; p.n_tiles_n = N / BN;
    p.BW = BW; p.BH = BH; p.BNI = BNI;
    p.tiles_w = (p.Wg + BW - 1) / BW; p.tiles_h = (p.Hg + BH - 1) / BH;
    p.tiles_per_cls = p.tiles_w * p.tiles_h * ((NB + BNI - 1) / BNI);
    p.NB = NB; p.Hout = Hout; p.Wout = Wout;
    p.bias = (const __nv_bfloat16*)bias; p.residual = (const __nv_bfloat16*)residual; p.relu = relu; p.C = (__nv_bfloat16*)Y; p.err = err_word(); p.stats = stats;
    return BN == 64 ? launch_conv2<64>(ta, tb, p, stream) : (BN == 128 ? launch_conv2<128>(ta, tb, p, stream) : launch_conv2<256>(ta, tb, p, stream));
}

// ---- weight layout helper:  [Cout][T][Cin] (the parameter, channels-last) -> [Cin][T][Cout] (the K-major B operand of dgrad), optional tap flip
// (flip = 1 gives the operand of "dgrad as a forward convolution of dY with the 180-degree rotated filter", used by the halo kernels).
__global__ void weight_transpose_kernel(const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cout, int T, int Cin, int flip) {
    __shared__ __nv_bfloat16 tile[32][33];
    const int t = blockIdx.z, to = flip ? T - 1 - t : t;
    const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int co = co0 + i, ci = ci0 + threadIdx.x;
        if (co < Cout && ci < Cin) tile[i][threadIdx.x] = w[((long long)co * T + t) * Cin + ci];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int ci = ci0 + i, co = co0 + threadIdx.x;
        if (co < Cout && ci < Cin) out[((long long)ci * T + to) * Cout + co] = tile[threadIdx.x][i];
    }
}
// All convolution weights of a model in ONE launch (per step: the dgrad operands of every layer, ~20 launches -> 1).
// table[i] = {src, dst, Cout, T, Cin, flip, first_block, blocks_x, blocks_y}; block -> tensor by linear search over n <= 256 entries.
struct WtEntry { const __nv_bfloat16* src; __nv_bfloat16* dst; int Cout, T, Cin, flip, first_block, bx, by, pad; };
// The launch is also the step's FIRST reader of the weights, so it can carry the protocol's consumer-side wait: with `wait.flags` set, one warp of
// every CTA acquires the PS's "weights of step s published" words (system scope, peer-written) before the CTA touches a weight — the separate
// 1-warp wait_flag_kernel disappears from the step and everything later in the stream is ordered behind this kernel as before.
struct ReadyWait { const long long* flags; int n; long long* step_ptr; long long* err; unsigned long long timeout_ns; };
__global__ void weight_transpose_batched_kernel(const WtEntry* __restrict__ table, int n, ReadyWait wait) {
    __shared__ __nv_bfloat16 tile[32][33];
    if (wait.flags != nullptr) {
        if (threadIdx.y == 0 && (int)threadIdx.x < wait.n) {
            const long long want = *(volatile long long*)wait.step_ptr;
            const unsigned long long t0 = globaltimer_ns();
            long long f;
            while ((f = ld_acquire_sys(wait.flags + threadIdx.x)) < want) {
                if (globaltimer_ns() - t0 > wait.timeout_ns) { atomicExch((unsigned long long*)wait.err, 100ull + threadIdx.x); break; }
                __nanosleep(100);
            }
            // a worker that (re)joins a running job adopts the server's step (same rule as wait_flag_kernel); flags only grow, so a CTA that
            // reads the adopted value simply waits for the same publication
            if (blockIdx.x == 0 && threadIdx.x == 0 && f > want) *(volatile long long*)wait.step_ptr = f;
        }
        __syncthreads();
    }
    int i = 0;
    while (i + 1 < n && (int)blockIdx.x >= table[i + 1].first_block) ++i;
    const WtEntry e = table[i];
    int rel = blockIdx.x - e.first_block;
    const int bx = rel % e.bx; rel /= e.bx;
    const int by = rel % e.by; const int t = rel / e.by;
    const int to = e.flip ? e.T - 1 - t : t;
    const int ci0 = bx * 32, co0 = by * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int co = co0 + r, ci = ci0 + threadIdx.x;
        if (co < e.Cout && ci < e.Cin) tile[r][threadIdx.x] = e.src[((long long)co * e.T + t) * e.Cin + ci];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int ci = ci0 + r, co = co0 + threadIdx.x;
        if (co < e.Cout && ci < e.Cin) e.dst[((long long)ci * e.T + to) * e.Cout + co] = tile[threadIdx.x][r];
    }
}
const char* weight_transpose_batched(const void* table, int n, int total_blocks, const void* flags, int nflags, void* step_ptr, void* err, double timeout_s,
                                     cudaStream_t stream) {
    if (nflags > 32) return "weight_transpose_batched: at most 32 ready flags";
    ReadyWait w{(const long long*)flags, nflags, (long long*)step_ptr, (long long*)err, (unsigned long long)(timeout_s * 1e9)};
    weight_transpose_batched_kernel<<<total_blocks, dim3(32, 8), 0, stream>>>((const WtEntry*)table, n, w);
    return cudaGetLastError() == cudaSuccess ? nullptr : "weight_transpose_batched launch failed";
}

const char* weight_transpose(const void* w, void* out, int Cout, int T, int Cin, int flip, cudaStream_t stream) {
    dim3 grid((Cin + 31) / 32, (Cout + 31) / 32, T), block(32, 8);
    weight_transpose_kernel<<<grid, block, 0, stream>>>((const __nv_bfloat16*)w, (__nv_bfloat16*)out, Cout, T, Cin, flip);
    return cudaGetLastError() == cudaSuccess ? nullptr : "weight_transpose launch failed";
}

}  // namespace psb

// tcgen05 GEMM and implicit-GEMM convolution for sm_100a.
//
//   C[M,N] (bf16) = A[M,K] x B[N,K]^T (+bias) (+ReLU)        A, B bf16 K-major ("TN"), fp32 accumulation in TMEM
//
// One persistent, warp-specialised kernel (grid = #SMs, static tile scheduler):
//   warp 0   TMA producer      cp.async.bulk.tensor -> 128B-swizzled smem ring (kStages deep), mbarrier complete_tx
//   warp 1   MMA issuer        one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (128 x BN x 16), accumulators
//                              double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1
//   warp 2   TMEM allocator
//   warps 4-7 epilogue         tcgen05.ld (32 lanes x 32 columns per warp) -> bias/ReLU -> bf16 -> global
//
// Convolution (fprop; dgrad of stride-1 convs is the same kernel on flipped/transposed weights) is the same GEMM with
// M = output pixels, N = Cout, K = R*S*Cin and NO im2col buffer: for filter tap (r,s) and channel block cb the A tile of
// 128 output pixels (a BNI x BH x BW rectangle of the NHWC activation) is ONE 4-D TMA box load at coordinates
// (cb*64, w0*stride+s-pad, h0*stride+r-pad, n0); out-of-bounds rows/columns (the padding halo) are zero-filled by
// the TMA unit, stride-2 convolutions use the tensor map's element strides.  Weights are [Cout][R][S][Cin] (the
// physical layout of channels-last parameters in the flat arena), i.e. already the K-major B operand.
//
// Reference ops replaced: nn.Conv2d / nn.Linear -> cuDNN / cuBLAS (/root/reference/src/model_ops/*.py).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <mutex>
#include "tc_common.cuh"
#include "pdl.cuh"

namespace psb {

template <int BN> struct TcCfg {
    static constexpr int kStages = BN == 64 ? 8 : (BN == 128 ? 6 : 4);
    static constexpr int kABytes = BM * BK * 2;
    static constexpr int kBBytes = BN * BK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kTmemCols = 2 * BN;                      // two accumulator stages (power of two: 128 / 256)
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const TcParams p) {
    using Cfg = TcCfg<BN>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
    uint64_t* empty_bar = full_bar + Cfg::kStages;
    uint64_t* tmem_full = empty_bar + Cfg::kStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_a); prefetch_tmap(&tmap_b); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tmem_full + s, 1); mbar_init(tmem_empty + s, 4); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_base_smem, Cfg::kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    const int n_tiles = p.n_tiles_m * p.n_tiles_n;

    if (warp == 0) {
        // ===================================================== TMA producer (whole warp runs the loop; one elected lane issues)
        int stage = 0; uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int mt = tile / p.n_tiles_n, nt = tile % p.n_tiles_n;
            int n0 = 0, h0 = 0, w0 = 0;
            if (p.mode == 1) {
                const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tn = mt / (p.tiles_w * p.tiles_h);
                n0 = tn * p.BNI; h0 = th * p.BH; w0 = tw * p.BW;
            }
            int tap = 0, cb = 0;
            for (int kb = 0; kb < p.num_kb; ++kb) {
                if (!mbar_wait(empty_bar + stage, phase ^ 1)) { *p.err = 1; return; }
                uint8_t* sa = smem + stage * Cfg::kStageBytes;
                uint8_t* sb = sa + Cfg::kABytes;
                if (elect_one()) {
                    mbar_arrive_expect_tx(full_bar + stage, Cfg::kStageBytes);
                    if (p.mode == 0) {
                        tma_load_2d(sa, &tmap_a, full_bar + stage, kb * BK, mt * BM);
                    } else {
                        const int r = tap / p.S, s = tap - r * p.S;
                        tma_load_4d(sa, &tmap_a, full_bar + stage, cb * BK, w0 * p.stride + s - p.pad, h0 * p.stride + r - p.pad, n0);
                    }
                    tma_load_2d(sb, &tmap_b, full_bar + stage, kb * BK, nt * BN);
                }
                __syncwarp();
                if (++cb == p.cblocks) { cb = 0; ++tap; }
                if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer (whole warp waits; one elected lane issues)
        constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0, 0);
        int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            if (!mbar_wait(tmem_empty + acc, acc_phase ^ 1)) { *p.err = 2; return; }
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * BN;
            for (int kb = 0; kb < p.num_kb; ++kb) {
                if (!mbar_wait(full_bar + stage, phase)) { *p.err = 3; return; }
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
                const uint32_t sb = sa + Cfg::kABytes;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        tc_mma_bf16(tmem_d, umma_desc_k_sw128(sa + k * 32), umma_desc_k_sw128(sb + k * 32), idesc, (kb | k) != 0);
                    tc_commit(empty_bar + stage);                   // smem slot reusable once these MMAs retire
                }
                __syncwarp();
                if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) tc_commit(tmem_full + acc);             // accumulator complete -> epilogue
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= kEpiWarp0) {
        // ===================================================== epilogue (4 warps, one TMEM lane quarter each)
        const int q = warp & 3;
        int acc = 0; uint32_t acc_phase = 0;
        float* const srow = p.stats ? p.stats + (size_t)(blockIdx.x * 4 + q) * 2 * p.N : nullptr;
        constexpr bool kRegStats = StatsCfg<BN>::kInRegs;
        const bool use_racc = kRegStats && srow != nullptr && p.n_tiles_n == 1;      // one N tile: a thread's columns never change
        float racc[kRegStats ? 2 * BN : 2];
#pragma unroll
        for (int j = 0; j < (kRegStats ? 2 * BN : 2); ++j) racc[j] = 0.f;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int mt = tile / p.n_tiles_n, nt = tile % p.n_tiles_n;
            const int row_in_tile = q * 32 + lane;
            long long out_row = -1;
            if (p.mode == 0) {
                const int m = mt * BM + row_in_tile;
                if (m < p.M) out_row = m;
            } else {
                const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tn = mt / (p.tiles_w * p.tiles_h);
                const int bw = row_in_tile % p.BW, bh = (row_in_tile / p.BW) % p.BH, bn = row_in_tile / (p.BW * p.BH);
                const int n = tn * p.BNI + bn, h = th * p.BH + bh, w = tw * p.BW + bw;
                if (n < p.NB && h < p.Hout && w < p.Wout) out_row = ((long long)n * p.Hout + h) * p.Wout + w;
            }
            if (!mbar_wait(tmem_full + acc, acc_phase)) { if (lane == 0) *p.err = 4; return; }
            tc_fence_after();
            epilogue_tile<BN>(p, tmem_base, acc, q, lane, out_row, nt, false, srow, use_racc ? racc : nullptr);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_relaxed(tmem_empty + acc);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if constexpr (kRegStats) { if (use_racc) stats_flush<BN>(racc, srow, p.N, 0, lane); }
    }
    // ---- teardown
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ===================================================================================================================
// CTA-pair GEMM: one 256 x 256 output tile per pair of SMs (cluster of 2), tcgen05.mma.cta_group::2.
//
// Why: an SS-mode MMA is paced by its shared-memory operand reads, ~ (A bytes + B bytes) / 64 cycles per instruction
// (DESIGN.md section 5).  With one CTA, 128 x 256 x 16 reads 4 KB of A + 8 KB of B = 192 cycles for 128 cycles of math.  In a
// pair each SM reads its own 128 rows of A and only HALF of B (128 of the 256 columns; the tensor cores exchange operands
// across the pair), 8 KB = 128 cycles: the math rate.
//   rank r of the pair loads A rows [256 mt + 128 r, +128) and B rows [256 nt + 128 r, +128); every TMA signals the
//   LEADER's full barrier (cta_group::2 form, barrier address mapped to rank 0); only the leader issues MMAs; tcgen05.commit
//   multicasts the "stage free" / "accumulator ready" arrivals to both CTAs; each CTA's epilogue drains its own 128 TMEM
//   lanes and arrives on the leader's tmem_empty barrier (8 arrivals: 4 warps x 2 CTAs).
struct Tc2Cfg {
    static constexpr int kStages = 6;
    static constexpr int kABytes = BM * BK * 2;                 // 128 rows of A
    static constexpr int kBBytes = 128 * BK * 2;                // 128 of the tile's 256 B rows
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kTmemCols = 512;                       // two 256-column accumulator stages
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

__global__ void __launch_bounds__(kThreads, 1)
tc_gemm2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const TcParams p) {
    using Cfg = Tc2Cfg;
    constexpr int BN2 = 256;
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
    cluster_sync_all();                         // peer barriers initialised before any remote arrive / TMA signal / commit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    const int n_tiles = p.n_tiles_m * p.n_tiles_n;
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

    if (warp == 0) {
        // ===================================================== TMA producer (both CTAs)
        int stage = 0; uint32_t phase = 0;
        for (int tile = pair; tile < n_tiles; tile += n_pairs) {
            const int mt = tile / p.n_tiles_n, nt = tile % p.n_tiles_n;
            for (int kb = 0; kb < p.num_kb; ++kb) {
                if (!mbar_wait(empty_bar + stage, phase ^ 1)) { *p.err = 1; goto teardown; }
                uint8_t* sa = smem + stage * Cfg::kStageBytes;
                uint8_t* sb = sa + Cfg::kABytes;
                if (elect_one()) {
                    const uint32_t lead_full = mapa_u32(smem_u32(full_bar + stage), 0);
                    if (leader) mbar_arrive_expect_tx(full_bar + stage, 2 * Cfg::kStageBytes);      // both CTAs' bytes land on this barrier
                    tma_load_2d_2sm(sa, &tmap_a, lead_full, kb * BK, mt * 256 + (int)rank * 128);
                    tma_load_2d_2sm(sb, &tmap_b, lead_full, kb * BK, nt * BN2 + (int)rank * 128);
                }
                __syncwarp();
                if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1 && leader) {
        // ===================================================== MMA issuer (leader CTA only)
        constexpr uint32_t idesc = umma_idesc_bf16(256, BN2, 0, 0);
        int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
        for (int tile = pair; tile < n_tiles; tile += n_pairs) {
            if (!mbar_wait(tmem_empty + acc, acc_phase ^ 1)) { *p.err = 2; goto teardown; }
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * BN2;
            for (int kb = 0; kb < p.num_kb; ++kb) {
                if (!mbar_wait(full_bar + stage, phase)) { *p.err = 3; goto teardown; }
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
                const uint32_t sb = sa + Cfg::kABytes;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        tc_mma2_bf16(tmem_d, umma_desc_k_sw128(sa + k * 32), umma_desc_k_sw128(sb + k * 32), idesc, (kb | k) != 0);
                    tc_commit2_mcast(empty_bar + stage, 0b11);       // the stage is free in BOTH CTAs once these MMAs retire
                }
                __syncwarp();
                if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) tc_commit2_mcast(tmem_full + acc, 0b11);
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= kEpiWarp0) {
        // ===================================================== epilogue (both CTAs; each drains its own 128 TMEM lanes)
        const int q = warp & 3;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = pair; tile < n_tiles; tile += n_pairs) {
            const int mt = tile / p.n_tiles_n, nt = tile % p.n_tiles_n;
            const int m = mt * 256 + (int)rank * 128 + q * 32 + lane;
            const long long out_row = m < p.M ? m : -1;
            if (!mbar_wait(tmem_full + acc, acc_phase)) { if (lane == 0) *p.err = 4; goto teardown; }
            tc_fence_after();
            epilogue_tile<BN2>(p, tmem_base, acc, q, lane, out_row, nt);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(tmem_empty + acc), 0));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
teardown:
    // every thread of both CTAs gets here (also after a timed-out wait): nobody may still be signalling the peer or using TMEM
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    if (warp == 2) tmem_dealloc2(tmem_base, Cfg::kTmemCols);
}

// ===================================================================================================================
// Halo variant for stride-1 RxS convolutions on feature maps with W >= 16, H >= 8.
//
// The per-tap kernel above re-fetches the A tile from L2 once per filter tap (9x for 3x3); measured on B200 it is
// bound by L2->SM bandwidth (1.2 GB of smem fills per ResNet layer at ~12 TB/s ~ 100 us, profiles/run3_conv.md).
// Here a tile is 16 rows x 8 columns of ONE image and, per 64-channel block, the input patch with its halo is loaded
// ONCE as a single TMA box of (64 ch, 16 columns, 16+R-1 rows) (the 8+S-1 needed columns padded to a 16-pixel pitch), so
// the smem buffer is [row][column][64 ch]: 128-byte pixels, 16 pixels (2048 B) per image row, and every box row is a
// contiguous 2 KB run in global memory (a first version used H as the fast dimension: 288 scattered 128-byte pieces per
// box throttled the TMA unit to ~10 B/cycle/SM).  With the GEMM rows ordered m = bh*8 + bw,
// the A operand of filter tap (r,s) is the *shifted window* that starts at byte ((r*16 + s) * 128) of that buffer:
// 8 consecutive pixels form one UMMA core group, groups are 2048 B apart (SBO), and because the window starts off the
// 1024-byte swizzle pattern boundary the descriptor carries base_offset = r.  A traffic drops from R*S*16 KB to
// (16+S-1)*2 KB per (tile, channel block): 4x less for 3x3.  B (weights) streams per (tap, channel block) as before.
// Weights stream one filter ROW per TMA: a 3-D box (64 k, BN out-channels, S taps) = S consecutive K-major B tiles, so the
// producer issues R+1 bulk copies per channel block instead of R*S+1 and the MMA warp waits on R barriers (the first
// version issued one 8-16 KB copy per tap and the MMA warp spent a third of its time waiting on them, profiles/run7_*.md).
constexpr int kMaxAStages = 4, kMaxBStages = 8;
template <int BN> struct HaloCfg {
    static constexpr int kAStages = 2;
    static constexpr int kBStages = BN == 64 ? 4 : 3;
    static constexpr int kBTapBytes = BN * BK * 2;
    static constexpr int kTmemCols = 2 * BN;
};

// CL = cluster size (1 or 2).  With CL = 2 the two CTAs of a cluster work on neighbouring M tiles of the same N tile and
// share the weight stream: each filter-row stage is fetched from L2 by ONE of them (alternating) and delivered to both
// through TMA multicast; a stage is recycled when both CTAs' MMAs have retired (multicast tcgen05.commit on both
// "empty" barriers, count 2).  Halves the dominant L2->SM term of the 64-channel layers.
template <int BN, int CL>
__global__ void __launch_bounds__(kThreads, 1)
tc_conv_halo_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const TcParams p) {
    using Cfg = HaloCfg<BN>;
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0;
    constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + p.a_stages * p.a_stage_bytes;
    const int b_stage_bytes = p.S * Cfg::kBTapBytes;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_b + Cfg::kBStages * b_stage_bytes);
    uint64_t* a_empty = a_full + kMaxAStages;
    uint64_t* b_full = a_empty + kMaxAStages;
    uint64_t* b_empty = b_full + Cfg::kBStages;
    uint64_t* tmem_full = b_empty + Cfg::kBStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_a); prefetch_tmap(&tmap_b); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.a_stages; ++s) { mbar_init(a_full + s, 1); mbar_init(a_empty + s, 1); }
        for (int s = 0; s < Cfg::kBStages; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, CL); }
        for (int s = 0; s < 2; ++s) { mbar_init(tmem_full + s, 1); mbar_init(tmem_empty + s, 4); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_base_smem, Cfg::kTmemCols);
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();          // peers' barriers must exist before any multicast copy / commit targets them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    // tile order: n-tile slowest, m-tiles fastest, so the CTAs of a cluster (consecutive blockIdx) get neighbouring M tiles of
    // the same N tile.  Every CTA runs the same number of iterations (ghost tiles past the end fetch zeros and store nothing).
    const int n_tiles = p.n_tiles_m * p.n_tiles_n;
    const int n_iters = (n_tiles + gridDim.x - 1) / gridDim.x;

    if (warp == 0) {
        int as = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0; uint32_t bcount = 0;
        for (int it = 0; it < n_iters; ++it) {
            const int tile = it * gridDim.x + blockIdx.x;
            const bool ghost = tile >= n_tiles;
            const int nt = ghost ? (n_tiles - 1) / p.n_tiles_m : tile / p.n_tiles_m, mt = ghost ? 0 : tile % p.n_tiles_m;
            const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, n = ghost ? p.NB : mt / (p.tiles_w * p.tiles_h);
            for (int cb = 0; cb < p.cblocks; ++cb) {
                if (!mbar_wait(a_empty + as, aph ^ 1)) { *p.err = 11; return; }
                if (elect_one()) {
                    mbar_arrive_expect_tx(a_full + as, p.a_tx_bytes);
                    tma_load_4d(smem_a + as * p.a_stage_bytes, &tmap_a, a_full + as, cb * BK, tw * 8 - p.pad, th * 16 - p.pad, n);
                }
                __syncwarp();
                if (++as == p.a_stages) { as = 0; aph ^= 1; }
                if (p.b_resident && it > 0) continue;          // weights were loaded once (stage r = filter row r) and stay
                for (int r = 0; r < p.R; ++r, ++bcount) {
                    if (!mbar_wait(b_empty + bs, bph ^ 1)) { *p.err = 12; return; }
                    if (elect_one()) {
                        mbar_arrive_expect_tx(b_full + bs, b_stage_bytes);
                        if (CL == 1) tma_load_3d(smem_b + bs * b_stage_bytes, &tmap_b, b_full + bs, cb * BK, nt * BN, r * p.S);
                        else if ((bcount % CL) == crank) tma_load_3d_mcast(smem_b + bs * b_stage_bytes, &tmap_b, b_full + bs, cb * BK, nt * BN, r * p.S, kMask);
                    }
                    __syncwarp();
                    if (++bs == Cfg::kBStages) { bs = 0; bph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0, 0);
        int as = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0; int acc = 0; uint32_t acc_phase = 0;
        for (int it = 0; it < n_iters; ++it) {
            if (!mbar_wait(tmem_empty + acc, acc_phase ^ 1)) { *p.err = 13; return; }
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * BN;
            for (int cb = 0; cb < p.cblocks; ++cb) {
                if (!mbar_wait(a_full + as, aph)) { *p.err = 14; return; }
                const uint32_t sa0 = smem_u32(smem_a + as * p.a_stage_bytes);
                for (int r = 0; r < p.R; ++r) {
                    if (p.b_resident) { bs = r; bph = 0; }      // stage r holds filter row r for the whole kernel (phase 0 stays complete)
                    if (!mbar_wait(b_full + bs, bph)) { *p.err = 15; return; }
                    tc_fence_after();
                    const uint32_t sb0 = smem_u32(smem_b + bs * b_stage_bytes);
                    if (elect_one()) {
                        for (int s = 0; s < p.S; ++s) {
                            // shifted window: the swizzle XOR is a function of the absolute smem address, so base_offset stays 0
                            const uint32_t sa = sa0 + (uint32_t)((r * p.halo_w + s) * 128);
                            const uint32_t sb = sb0 + (uint32_t)(s * Cfg::kBTapBytes);
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k)
                                tc_mma_bf16(tmem_d, umma_desc_k_sw128_sbo(sa + k * 32, (uint32_t)p.halo_w * 128u), umma_desc_k_sw128(sb + k * 32), idesc, (cb | r | s | k) != 0);
                        }
                        if (!p.b_resident) { if (CL == 1) tc_commit(b_empty + bs); else tc_commit_mcast(b_empty + bs, kMask); }
                    }
                    __syncwarp();
                    if (!p.b_resident && ++bs == Cfg::kBStages) { bs = 0; bph ^= 1; }
                }
                if (elect_one()) tc_commit(a_empty + as);
                __syncwarp();
                if (++as == p.a_stages) { as = 0; aph ^= 1; }
            }
            if (elect_one()) tc_commit(tmem_full + acc);
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= kEpiWarp0) {
        const int q = warp & 3;
        int acc = 0; uint32_t acc_phase = 0;
        float* const srow = p.stats ? p.stats + (size_t)(blockIdx.x * 4 + q) * 2 * p.N : nullptr;
        constexpr bool kRegStats = StatsCfg<BN>::kInRegs;
        const bool use_racc = kRegStats && srow != nullptr && p.n_tiles_n == 1;      // one N tile: a thread's columns never change
        float racc[kRegStats ? 2 * BN : 2];
#pragma unroll
        for (int j = 0; j < (kRegStats ? 2 * BN : 2); ++j) racc[j] = 0.f;
        for (int it = 0; it < n_iters; ++it) {
            const int tile = it * gridDim.x + blockIdx.x;
            const bool ghost = tile >= n_tiles;
            const int nt = ghost ? 0 : tile / p.n_tiles_m, mt = ghost ? 0 : tile % p.n_tiles_m;
            const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, n = mt / (p.tiles_w * p.tiles_h);
            const int m = q * 32 + lane;                       // GEMM row: m = bh*8 + bw
            const int h = th * 16 + (m >> 3), w = tw * 8 + (m & 7);
            const long long out_row = (!ghost && h < p.Hout && w < p.Wout) ? ((long long)n * p.Hout + h) * p.Wout + w : -1;
            if (!mbar_wait(tmem_full + acc, acc_phase)) { if (lane == 0) *p.err = 16; return; }
            tc_fence_after();
            epilogue_tile<BN>(p, tmem_base, acc, q, lane, out_row, nt, false, srow, use_racc ? racc : nullptr);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_relaxed(tmem_empty + acc);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if constexpr (kRegStats) { if (use_racc) stats_flush<BN>(racc, srow, p.N, 0, lane); }
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();          // no CTA may exit while a peer can still multicast into it
    tc_fence_after();
    if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ===================================================================================================================
// CTA-pair halo convolution (default for halo-eligible shapes since it was validated on hardware in round 2).
//
// Same tile geometry and shifted-window A descriptors as tc_conv_halo_kernel, with the pair mechanics proven by tc_gemm2_kernel:
// the two CTAs of a cluster take neighbouring M tiles (2 x 128 output pixels) of the same N tile; each loads its own halo patch
// and HALF of the weight rows (BN/2 output channels of every tap); every TMA signals the leader's barriers; the leader issues
// tcgen05.mma.cta_group::2 with M = 256, N = BN; commits are multicast to both CTAs.  Per MMA each SM then reads 4 KB of A and
// BN/2 x 32 B of B instead of BN x 32 B (128 -> 96 cycles per instruction for BN = 128, 96 -> 80 for BN = 64).
template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
tc_conv_halo2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const TcParams p) {
    using Cfg = HaloCfg<BN>;
    constexpr int kHalfTapBytes = (BN / 2) * BK * 2;          // this CTA's share of one tap's B tile
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + p.a_stages * p.a_stage_bytes;
    const int b_stage_bytes = p.S * kHalfTapBytes;
    // B stages: a ring of p.b_stages filter-row stages, or (p.b_resident) ALL cblocks * R filter rows of this CTA's half of the weights, loaded
    // once and kept for the whole kernel (stage = cb * R + r): the weights of a 64- or 128-channel layer are 36-144 KB per CTA, while
    // re-streaming them for every tile made L2->SM delivery (~40 B/cycle/SM) the bound of these layers (profiles/run25_conv_ncu.md).
    const int nbs = p.b_stages;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_b + nbs * b_stage_bytes);
    uint64_t* a_empty = a_full + kMaxAStages;
    uint64_t* b_full = a_empty + kMaxAStages;
    uint64_t* b_empty = b_full + kMaxBStages;
    uint64_t* tmem_full = b_empty + kMaxBStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_a); prefetch_tmap(&tmap_b); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.a_stages; ++s) { mbar_init(a_full + s, 1); mbar_init(a_empty + s, 1); }
        for (int s = 0; s < nbs; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tmem_full + s, 1); mbar_init(tmem_empty + s, 8); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc2(tmem_base_smem, Cfg::kTmemCols);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    pdl_trigger(); pdl_wait();                  // pdl.cuh: prologue done; from here on the kernel touches global memory
    const uint32_t tmem_base = *tmem_base_smem;
    // pair tiles: nt slowest, pairs of neighbouring M tiles fastest; rank r takes M tile 2*pm + r (a ghost when n_tiles_m is odd)
    const int n_pairs_m = (p.n_tiles_m + 1) / 2;
    const int n_ptiles = n_pairs_m * p.n_tiles_n;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

    if (warp == 0) {
        int as = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0;
        bool first = true;
        for (int pt = cluster_id; pt < n_ptiles; pt += n_clusters) {
            const int nt = pt / n_pairs_m, mt = 2 * (pt % n_pairs_m) + (int)rank;
            const bool ghost = mt >= p.n_tiles_m;
            const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, n = ghost ? p.NB : mt / (p.tiles_w * p.tiles_h);
            for (int cb = 0; cb < p.cblocks; ++cb) {
                if (!mbar_wait(a_empty + as, aph ^ 1)) { *p.err = 21; goto teardown; }
                if (elect_one()) {
                    if (leader) mbar_arrive_expect_tx(a_full + as, 2 * p.a_tx_bytes);
                    tma_load_4d_2sm(smem_a + as * p.a_stage_bytes, &tmap_a, mapa_u32(smem_u32(a_full + as), 0), cb * BK, tw * 8 - p.pad, th * 16 - p.pad, n);
                }
                __syncwarp();
                if (++as == p.a_stages) { as = 0; aph ^= 1; }
                if (p.b_resident && !first) continue;              // the weights are already resident (stage cb * R + r)
                for (int r = 0; r < p.R; ++r) {
                    if (!mbar_wait(b_empty + bs, bph ^ 1)) { *p.err = 22; goto teardown; }
                    if (elect_one()) {
                        if (leader) mbar_arrive_expect_tx(b_full + bs, 2 * b_stage_bytes);
                        tma_load_3d_2sm(smem_b + bs * b_stage_bytes, &tmap_b, mapa_u32(smem_u32(b_full + bs), 0), cb * BK, nt * BN + (int)rank * (BN / 2), r * p.S);
                    }
                    __syncwarp();
                    if (++bs == nbs) { bs = 0; bph ^= 1; }
                }
            }
            first = false;
        }
    } else if (warp == 1 && leader) {
        constexpr uint32_t idesc = umma_idesc_bf16(256, BN, 0, 0);
        const uint64_t a_tmpl = umma_desc_k_sw128_sbo(0, (uint32_t)p.halo_w * 128u), b_tmpl = umma_desc_k_sw128(0);
        int as = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0; int acc = 0; uint32_t acc_phase = 0;
        for (int pt = cluster_id; pt < n_ptiles; pt += n_clusters) {
            if (!mbar_wait(tmem_empty + acc, acc_phase ^ 1)) { *p.err = 23; goto teardown; }
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * BN;
            for (int cb = 0; cb < p.cblocks; ++cb) {
                if (!mbar_wait(a_full + as, aph)) { *p.err = 24; goto teardown; }
                // Descriptors are advanced by ADDING to the 14-bit address field (byte offset >> 4; shared memory is < 256 KB, no carry out of the
                // field): with N = 64 an MMA retires every 32 cycles, and rebuilding two 64-bit descriptors per instruction (~20 dependent
                // integer ops in one thread) made the ISSUE loop the bound — 49 cycles per MMA measured (profiles/run28_mma_issue.md).
                const uint64_t a0 = a_tmpl + (uint64_t)((smem_u32(smem_a + as * p.a_stage_bytes) & 0x3FFFFu) >> 4);
                for (int r = 0; r < p.R; ++r) {
                    if (p.b_resident) { bs = cb * p.R + r; bph = 0; }        // phase 0 of a resident stage stays complete
                    if (!mbar_wait(b_full + bs, bph)) { *p.err = 25; goto teardown; }
                    tc_fence_after();
                    const uint64_t ar = a0 + (uint64_t)(r * p.halo_w * 8);
                    const uint64_t b0 = b_tmpl + (uint64_t)((smem_u32(smem_b + bs * b_stage_bytes) & 0x3FFFFu) >> 4);
                    if (elect_one()) {
                        if (p.S == 3) {
#pragma unroll
                            for (int s = 0; s < 3; ++s)
#pragma unroll
                                for (int k = 0; k < BK / 16; ++k)
                                    tc_mma2_bf16(tmem_d, ar + (uint64_t)(s * 8 + k * 2), b0 + (uint64_t)(s * (kHalfTapBytes >> 4) + k * 2), idesc,
                                                 (s | k) != 0 ? 1u : (uint32_t)((cb | r) != 0));
                        } else {
                            for (int s = 0; s < p.S; ++s)
#pragma unroll
                                for (int k = 0; k < BK / 16; ++k)
                                    tc_mma2_bf16(tmem_d, ar + (uint64_t)(s * 8 + k * 2), b0 + (uint64_t)(s * (kHalfTapBytes >> 4) + k * 2), idesc, (cb | r | s | k) != 0);
                        }
                        if (!p.b_resident) tc_commit2_mcast(b_empty + bs, 0b11);
                    }
                    __syncwarp();
                    if (!p.b_resident && ++bs == nbs) { bs = 0; bph ^= 1; }
                }
                if (elect_one()) tc_commit2_mcast(a_empty + as, 0b11);
                __syncwarp();
                if (++as == p.a_stages) { as = 0; aph ^= 1; }
            }
            if (elect_one()) tc_commit2_mcast(tmem_full + acc, 0b11);
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= kEpiWarp0) {
        const int q = warp & 3;
        int acc = 0; uint32_t acc_phase = 0;
        float* const srow = p.stats ? p.stats + (size_t)(blockIdx.x * 4 + q) * 2 * p.N : nullptr;
        constexpr bool kRegStats = StatsCfg<BN>::kInRegs;
        const bool use_racc = kRegStats && srow != nullptr && p.n_tiles_n == 1;      // one N tile: a thread's columns never change
        float racc[kRegStats ? 2 * BN : 2];
#pragma unroll
        for (int j = 0; j < (kRegStats ? 2 * BN : 2); ++j) racc[j] = 0.f;
        for (int pt = cluster_id; pt < n_ptiles; pt += n_clusters) {
            const int nt = pt / n_pairs_m, mt = 2 * (pt % n_pairs_m) + (int)rank;
            const bool ghost = mt >= p.n_tiles_m;
            const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, n = mt / (p.tiles_w * p.tiles_h);
            const int m = q * 32 + lane;
            const int h = th * 16 + (m >> 3), w = tw * 8 + (m & 7);
            const long long out_row = (!ghost && h < p.Hout && w < p.Wout) ? ((long long)n * p.Hout + h) * p.Wout + w : -1;
            if (!mbar_wait(tmem_full + acc, acc_phase)) { if (lane == 0) *p.err = 26; goto teardown; }
            tc_fence_after();
            epilogue_tile<BN>(p, tmem_base, acc, q, lane, out_row, nt, false, srow, use_racc ? racc : nullptr);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(tmem_empty + acc), 0));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if constexpr (kRegStats) { if (use_racc) stats_flush<BN>(racc, srow, p.N, 0, lane); }
    }
teardown:
    // every thread of both CTAs reaches this point, also after a timed-out wait (error word set): a thread that returned early would
    // leave its peers blocked in the cluster barrier
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    if (warp == 2) tmem_dealloc2(tmem_base, Cfg::kTmemCols);
}

// ===================================================================================================================
// Weight gradient:  dW[co][r][s][ci] = sum over output pixels p of dY[p, co] * X[p*stride + (r,s) - pad, ci]
//
// A GEMM whose reduction dimension is the PIXEL index, so both operands are "MN-major" for the tensor core (the channel
// dimension — M = co for dY, N = ci for X — is the contiguous one in NHWC memory).  A K-block is a rectangle of 128 output
// pixels; its dY tile and its tap-shifted X tile are the same kind of 4-D TMA box loads the forward kernel uses (padding and
// stride-2 come from TMA zero-fill / element strides), landing in smem as [pixel][64 channels] 128-byte rows = the canonical
// MN-major SWIZZLE_128B operand (8 pixel rows per 1024-byte atom = SBO, 64-channel groups one box apart = LBO).
// Work item = (filter tap, co tile, ci tile, K split); every item accumulates 128 K-blocks worth of 128xN products in TMEM and
// adds its partial tile into an fp32 [Cout][R*S][Cin] scratch with vector reductions (REDG.E.ADD.F32x4); a small kernel then
// converts the scratch to the bf16 gradient and leaves it zeroed.
struct WgParams {
    int taps, S, stride, pad, Cin, Cout, NB, Ho, Wo;
    int BW, BH, BNI, tiles_w, tiles_h, kb_total;
    int co_tiles, ci_tiles, ksplit, kb_per_split, n_items;
    float* dW32;
    int* err;
};
// KP = pixels per K-block (128, or 64 for the 256-column configuration so that four stages still fit in shared memory);
// one [KP pixels][64 channels] box = KP*128 bytes
template <int BN, int MM, int KP> struct WgCfg {
    static constexpr int kBoxBytes = KP * 128;
    static constexpr int kGroupsA = MM / 64, kGroupsB = BN / 64;
    static constexpr int kStageBytes = (kGroupsA + kGroupsB) * kBoxBytes;
    static constexpr int kStages = (200 * 1024) / kStageBytes;
    static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

template <int BN, int MM, int KP>
__global__ void __launch_bounds__(kThreads, 1)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x, const WgParams p) {
    using Cfg = WgCfg<BN, MM, KP>;
    constexpr int kWgBoxBytes = Cfg::kBoxBytes;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
    uint64_t* empty_bar = full_bar + Cfg::kStages;
    uint64_t* tmem_full = empty_bar + Cfg::kStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_dy); prefetch_tmap(&tmap_x); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tmem_full + s, 1); mbar_init(tmem_empty + s, 4); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_base_smem, Cfg::kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_trigger(); pdl_wait();                  // pdl.cuh: prologue done; from here on the kernel touches global memory
    const uint32_t tmem_base = *tmem_base_smem;

    // item -> (tap fastest: CTAs of one wave read the same dY / X region, L2-friendly), then K split, ci tile, co tile
    auto decode = [&](int item, int& tap, int& ks, int& nt, int& mt) {
        tap = item % p.taps; int r = item / p.taps;
        ks = r % p.ksplit; r /= p.ksplit;
        nt = r % p.ci_tiles; mt = r / p.ci_tiles;
    };

    if (warp == 0) {
        int stage = 0; uint32_t phase = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            int tap, ks, nt, mt; decode(item, tap, ks, nt, mt);
            const int fr = tap / p.S, fs = tap - fr * p.S;
            const int kb0 = ks * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.kb_total);
            for (int kb = kb0; kb < kb1; ++kb) {
                const int tw = kb % p.tiles_w, th = (kb / p.tiles_w) % p.tiles_h, tn = kb / (p.tiles_w * p.tiles_h);
                const int n0 = tn * p.BNI, h0 = th * p.BH, w0 = tw * p.BW;
                if (!mbar_wait(empty_bar + stage, phase ^ 1)) { *p.err = 21; return; }
                uint8_t* sa = smem + stage * Cfg::kStageBytes;
                uint8_t* sb = sa + Cfg::kGroupsA * kWgBoxBytes;
                if (elect_one()) {
                    mbar_arrive_expect_tx(full_bar + stage, Cfg::kStageBytes);
#pragma unroll
                    for (int g = 0; g < Cfg::kGroupsA; ++g)
                        tma_load_4d(sa + g * kWgBoxBytes, &tmap_dy, full_bar + stage, mt * MM + g * 64, w0, h0, n0);
#pragma unroll
                    for (int g = 0; g < Cfg::kGroupsB; ++g)
                        tma_load_4d(sb + g * kWgBoxBytes, &tmap_x, full_bar + stage, nt * BN + g * 64, w0 * p.stride + fs - p.pad, h0 * p.stride + fr - p.pad, n0);
                }
                __syncwarp();
                if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc_bf16(MM, BN, 1, 1);          // both operands MN-major
        const uint64_t wg_tmpl = umma_desc_mn_sw128(0, kWgBoxBytes);
        int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            int tap, ks, nt, mt; decode(item, tap, ks, nt, mt);
            const int kb0 = ks * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, p.kb_total);
            if (!mbar_wait(tmem_empty + acc, acc_phase ^ 1)) { *p.err = 22; return; }
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * BN;
            for (int kb = kb0; kb < kb1; ++kb) {
                if (!mbar_wait(full_bar + stage, phase)) { *p.err = 23; return; }
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
                const uint64_t da = wg_tmpl + (uint64_t)((sa & 0x3FFFFu) >> 4), db = da + (uint64_t)((Cfg::kGroupsA * kWgBoxBytes) >> 4);
                if (elect_one()) {
#pragma unroll
                    for (int j = 0; j < KP / 16; ++j)            // 16 pixels per MMA = two 8-row atoms = 2048 bytes (descriptor address field += 128)
                        tc_mma_bf16(tmem_d, da + (uint64_t)(j * 128), db + (uint64_t)(j * 128), idesc, j > 0 ? 1u : (uint32_t)(kb > kb0));
                    tc_commit(empty_bar + stage);
                }
                __syncwarp();
                if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) tc_commit(tmem_full + acc);
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else if (warp >= kEpiWarp0) {
        const int q = warp & 3;
        int acc = 0; uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            int tap, ks, nt, mt; decode(item, tap, ks, nt, mt);
            // TMEM rows: M=128 -> lane = row; M=64 -> "half sub-partition" layout: row m lives in lane (m%16) + 32*(m/16)
            const int row = MM == 128 ? q * 32 + lane : q * 16 + lane;
            const bool row_ok = (MM == 128 || lane < 16) && (mt * MM + row) < p.Cout;
            const int kb0 = ks * p.kb_per_split;
            const bool has_work = kb0 < p.kb_total;
            if (!mbar_wait(tmem_full + acc, acc_phase)) { if (lane == 0) *p.err = 24; return; }
            tc_fence_after();
            float* drow = p.dW32 + ((long long)(mt * MM + row) * p.taps + tap) * p.Cin + (long long)nt * BN;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), v);
                tmem_ld_wait();
                if (row_ok && has_work) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        if (nt * BN + c0 + j < p.Cin)
                            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(drow + c0 + j), "f"(__uint_as_float(v[j])), "f"(__uint_as_float(v[j + 1])),
                                         "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3])) : "memory");
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_relaxed(tmem_empty + acc);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// scratch fp32 -> bf16 gradient, scratch left zeroed (self-cleaning)
__global__ void wgrad_finish_kernel(float* __restrict__ acc, __nv_bfloat16* __restrict__ out, long long n) {
    pdl_trigger(); pdl_wait();
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += (long long)gridDim.x * blockDim.x * 4) {
        float4 v = *reinterpret_cast<float4*>(acc + i);
        *reinterpret_cast<float4*>(acc + i) = make_float4(0.f, 0.f, 0.f, 0.f);
        __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 o; o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(out + i) = o;
    }
}

void wgrad_finish_launch(float* acc, void* out, long long n, cudaStream_t stream) {
    launch_pdl(wgrad_finish_kernel, dim3((int)((n / 4 + 255) / 256 < 1184 ? (n / 4 + 255) / 256 : 1184)), dim3(256), 0, stream, acc, (__nv_bfloat16*)out, n);
}

// ------------------------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// rank-`rank` bf16 tensor map with 128-byte swizzle; dims/strides innermost first (strides in BYTES, for dims 1..rank-1)
const char* make_tmap(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                             const uint32_t* box, const uint32_t* estrides) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return "cuTensorMapEncodeTiled unavailable";
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), (const cuuint64_t*)dims,
                     (const cuuint64_t*)strides_bytes, (const cuuint32_t*)box, (const cuuint32_t*)estrides, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { static char buf[64]; snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d)", (int)r); return buf; }
    return nullptr;
}

int* err_word() {                // one device word shared by all launches; checked by tc_check_error()
    static int* p = nullptr;
    if (!p) { cudaMalloc(&p, sizeof(int)); cudaMemset(p, 0, sizeof(int)); }
    return p;
}
int tc_check_error() {
    int h = 0;
    cudaMemcpy(&h, err_word(), sizeof(int), cudaMemcpyDeviceToHost);
    if (h) cudaMemset(err_word(), 0, sizeof(int));
    return h;
}

static int g_wgrad_sms = 0;          // 0 = all; otherwise cap the grid of the weight-gradient kernels (they run on a side stream next to the critical path)
void tc_set_wgrad_sms(int n) { g_wgrad_sms = n > 0 ? n : 0; }
int wgrad_sms() { const int n = num_sms(); return g_wgrad_sms > 0 && g_wgrad_sms < n ? g_wgrad_sms : n; }
int num_sms() {
    static int n = 0;
    if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); }
    return n;
}

template <int BN>
static const char* launch(const CUtensorMap& ta, const CUtensorMap& tb, const TcParams& p, cudaStream_t stream) {
    using Cfg = TcCfg<BN>;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(tc_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess) return "cannot raise dynamic smem";
        attr = true;
    }
    const int tiles = p.n_tiles_m * p.n_tiles_n;
    const int grid = tiles < num_sms() ? tiles : num_sms();
    tc_gemm_kernel<BN><<<grid, kThreads, Cfg::kSmemBytes, stream>>>(ta, tb, p);
    return cudaGetLastError() == cudaSuccess ? nullptr : "kernel launch failed";
}

static int g_halo_resident = 1;
void tc_set_halo_resident(int r) { g_halo_resident = r; }
static int g_halo_cluster = 2;
void tc_set_halo_cluster(int c) { g_halo_cluster = c == 2 ? 2 : 1; }

template <int BN, int CL>
static const char* launch_halo(const CUtensorMap& ta, const CUtensorMap& tb, const TcParams& p, cudaStream_t stream) {
    using Cfg = HaloCfg<BN>;
    const int smem_bytes = p.a_stages * p.a_stage_bytes + Cfg::kBStages * p.S * Cfg::kBTapBytes + 1024 + 512;
    if (smem_bytes > 227 * 1024) return "halo tile does not fit in shared memory";
    static int attr = 0;
    if (attr < smem_bytes) {
        if (cudaFuncSetAttribute(tc_conv_halo_kernel<BN, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess) return "cannot raise dynamic smem";
        attr = smem_bytes;
    }
    const int tiles = p.n_tiles_m * p.n_tiles_n;
    int grid = tiles < num_sms() ? tiles : num_sms();
    grid = (grid + CL - 1) / CL * CL;
    if (grid > num_sms()) grid -= CL;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, tc_conv_halo_kernel<BN, CL>, ta, tb, p) == cudaSuccess ? nullptr : "kernel launch failed";
}

static int g_halo_pitch = 16;         // 16, or 0 = tight (8 + S - 1 pixels); tc_set_halo_pitch
void tc_set_halo_pitch(int v) { g_halo_pitch = v == 16 ? 16 : 0; }
static int g_halo_pair = 1;           // CTA-pair halo kernel (validated on B200 in round 2); tc_set_halo_pair(0) selects the 1-CTA variants
void tc_set_halo_pair(int v) { g_halo_pair = v ? 1 : 0; }

template <int BN>
static const char* launch_halo_pair(const CUtensorMap& ta, const CUtensorMap& tb, const TcParams& p, cudaStream_t stream) {
    using Cfg = HaloCfg<BN>;
    const int smem_bytes = p.a_stages * p.a_stage_bytes + p.b_stages * p.S * (BN / 2) * BK * 2 + 1024 + 512;
    if (smem_bytes > 227 * 1024) return "halo tile does not fit in shared memory";
    static int attr = 0;
    if (attr < smem_bytes) {
        if (cudaFuncSetAttribute(tc_conv_halo2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess) return "cannot raise dynamic smem";
        attr = smem_bytes;
    }
    const int ptiles = ((p.n_tiles_m + 1) / 2) * p.n_tiles_n;
    int pairs = num_sms() / 2;
    if (ptiles < pairs) pairs = ptiles;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = pdl_attr(at, 1);
    return cudaLaunchKernelEx(&cfg, tc_conv_halo2_kernel<BN>, ta, tb, p) == cudaSuccess ? nullptr : "kernel launch failed";
}

static int g_conv_algo = -1;     // -1 auto (= halo where eligible), 0 per-tap, 1 halo
void tc_set_halo_desc_mode(int) {}   // kept for scripts/conv_one.py; the base_offset experiment is settled (always 0)
void tc_set_conv_algo(int a) { g_conv_algo = a; }

static int g_gemm_bn = -1;       // -1 auto, 128 / 256 force the N tile of the plain GEMM, 2 = CTA-pair kernel (256 x 256 tiles)
void tc_set_gemm_bn(int b) { g_gemm_bn = b; }

static const char* launch_pair(const CUtensorMap& ta, const CUtensorMap& tb, const TcParams& p, cudaStream_t stream) {
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(tc_gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Tc2Cfg::kSmemBytes) != cudaSuccess) return "cannot raise dynamic smem";
        attr = true;
    }
    const int tiles = p.n_tiles_m * p.n_tiles_n;
    int pairs = num_sms() / 2;
    if (tiles < pairs) pairs = tiles;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = Tc2Cfg::kSmemBytes; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, tc_gemm2_kernel, ta, tb, p) == cudaSuccess ? nullptr : "kernel launch failed";
}

// C[M,N] = A[M,K] B[N,K]^T ; lda/ldb = K (dense rows)
const char* tc_gemm_tn(const void* A, const void* B, void* C, const void* bias, int M, int N, int K, int relu, cudaStream_t stream) {
    if (K % 8) return "K must be a multiple of 8";
    int BN = N <= 64 ? 64 : 128;
    // CTA-pair kernel (256 x 256 tiles): forced by tc_set_gemm_bn(2); automatic when the problem fills every SM pair
    // (measured 1.28-1.56 PF/s vs 1.12-1.53 for the 1-CTA kernel, profiles/run22_gemm_pair.md)
    const bool pair = N >= 256 && M >= 256 && (g_gemm_bn == 2 || (g_gemm_bn == -1 && ((M + 255) / 256) * ((N + 255) / 256) >= num_sms() / 2));
    if (pair || (g_gemm_bn == 256 && N >= 256)) BN = 256;
    else if (g_gemm_bn == -1 && N >= 512 && ((M + BM - 1) / BM) * ((N + 255) / 256) >= num_sms()) BN = 256;   // enough tiles to fill the chip
    CUtensorMap ta, tb;
    uint64_t da[2] = {(uint64_t)K, (uint64_t)M}, sa[1] = {(uint64_t)K * 2};
    uint64_t db[2] = {(uint64_t)K, (uint64_t)N}, sb[1] = {(uint64_t)K * 2};
    uint32_t boxa[2] = {BK, BM}, boxb[2] = {BK, (uint32_t)(pair ? 128 : BN)}, es[2] = {1, 1};
    if (const char* e = make_tmap(&ta, A, 2, da, sa, boxa, es)) return e;
    if (const char* e = make_tmap(&tb, B, 2, db, sb, boxb, es)) return e;
    TcParams p{};
    p.M = M; p.N = N; p.num_kb = (K + BK - 1) / BK; p.mode = 0;
    p.n_tiles_m = (M + (pair ? 256 : BM) - 1) / (pair ? 256 : BM); p.n_tiles_n = (N + BN - 1) / BN;
    p.bias = (const __nv_bfloat16*)bias; p.relu = relu; p.C = (__nv_bfloat16*)C; p.err = err_word();
    if (pair) return launch_pair(ta, tb, p, stream);
    return BN == 64 ? launch<64>(ta, tb, p, stream) : (BN == 128 ? launch<128>(ta, tb, p, stream) : launch<256>(ta, tb, p, stream));
}

// Y[NB,Hout,Wout,Cout] = conv(X[NB,Hin,Win,Cin], Wt[Cout][R][S][Cin]), NHWC bf16
const char* tc_conv_fprop(const void* X, const void* Wt, void* Y, const void* bias, const void* residual, int NB, int Hin, int Win, int Cin, int Cout,
                          int R, int S, int stride, int pad, int relu, float* stats, cudaStream_t stream) {
    if (Cin % 64) return "Cin must be a multiple of 64";
    if (Cout % 8) return "Cout must be a multiple of 8";
    const int Hout = (Hin + 2 * pad - R) / stride + 1, Wout = (Win + 2 * pad - S) / stride + 1;
    const bool halo_ok = stride == 1 && R * S > 1 && S <= 8 && Wout >= 8 && Hout >= 16 && (16 + R - 1) <= 256;
    if (g_conv_algo != 0 && halo_ok) {
        const int BN = Cout <= 64 ? 64 : 128;
        const bool pair = g_halo_pair && Cout % BN == 0;        // each CTA of the pair loads BN/2 weight rows per tap
        CUtensorMap ta, tb;
        uint64_t da[4] = {(uint64_t)Cin, (uint64_t)Win, (uint64_t)Hin, (uint64_t)NB};
        uint64_t sa[3] = {(uint64_t)Cin * 2, (uint64_t)Win * Cin * 2, (uint64_t)Hin * Win * Cin * 2};
        // patch pitch: 16 pixels (2 KB rows), or exactly the 8 + S - 1 columns a tile needs (g_halo_pitch = 0: 10 pixels for 3x3 -> 37 % less
        // L2->SM traffic per tile; the 8-pixel row groups are then 1280 bytes apart, which the address-based swizzle tolerates)
        const int pitch = g_halo_pitch == 16 ? 16 : 8 + S - 1;
        uint32_t boxa[4] = {BK, (uint32_t)pitch, (uint32_t)(16 + R - 1), 1}, esa[4] = {1, 1, 1, 1};
        const int Ktot = R * S * Cin;
        // weights [Cout][R*S][Cin] as a 3-D tensor (Cin, Cout, taps): one box = S taps of one filter row
        uint64_t db[3] = {(uint64_t)Cin, (uint64_t)Cout, (uint64_t)(R * S)}, sb[2] = {(uint64_t)Ktot * 2, (uint64_t)Cin * 2};
        uint32_t boxb[3] = {BK, (uint32_t)(pair ? BN / 2 : BN), (uint32_t)S}, esb[3] = {1, 1, 1};
        if (const char* e = make_tmap(&ta, X, 4, da, sa, boxa, esa)) return e;
        if (const char* e = make_tmap(&tb, Wt, 3, db, sb, boxb, esb)) return e;
        TcParams p{};
        p.M = NB * Hout * Wout; p.N = Cout; p.num_kb = Ktot / BK; p.mode = 2;
        p.Hout = Hout; p.Wout = Wout; p.NB = NB; p.stride = 1; p.pad = pad; p.S = S; p.R = R; p.cblocks = Cin / BK;
        p.tiles_w = (Wout + 7) / 8; p.tiles_h = (Hout + 15) / 16;
        p.n_tiles_m = p.tiles_w * p.tiles_h * NB; p.n_tiles_n = (Cout + BN - 1) / BN;
        p.halo_w = pitch; p.a_tx_bytes = (16 + R - 1) * pitch * 128; p.a_stage_bytes = (p.a_tx_bytes + 1023) / 1024 * 1024;      // stages stay 1024-byte aligned
        // weight-stationary: with one channel block and one N tile the R filter-row stages hold ALL the weights this CTA ever needs
        p.b_resident = (g_halo_resident && p.cblocks == 1 && p.n_tiles_n == 1 && R <= HaloCfg<64>::kBStages && BN == 64) ? 1 : 0;
        p.a_stages = 2;
        if (p.b_resident && 3 * p.a_stage_bytes + HaloCfg<64>::kBStages * S * HaloCfg<64>::kBTapBytes + 2048 <= 227 * 1024) p.a_stages = 3;
        p.bias = (const __nv_bfloat16*)bias; p.residual = (const __nv_bfloat16*)residual; p.relu = relu; p.C = (__nv_bfloat16*)Y; p.err = err_word(); p.stats = stats;
        if (pair) {
            // keep this CTA's half of ALL the weights in shared memory when they fit next to the activation stages
            const int b_stage = S * (BN / 2) * BK * 2, all_b = p.cblocks * R * b_stage;
            p.b_resident = 0; p.a_stages = 2; p.b_stages = BN == 64 ? HaloCfg<64>::kBStages : HaloCfg<128>::kBStages;
            if (g_halo_resident && p.n_tiles_n == 1 && p.cblocks * R <= kMaxBStages && 2 * p.a_stage_bytes + all_b + 2048 <= 227 * 1024) {
                p.b_resident = 1; p.b_stages = p.cblocks * R;
                while (p.a_stages < kMaxAStages && (p.a_stages + 1) * p.a_stage_bytes + all_b + 2048 <= 227 * 1024) ++p.a_stages;
            }
            return BN == 64 ? launch_halo_pair<64>(ta, tb, p, stream) : launch_halo_pair<128>(ta, tb, p, stream);
        }
        if (g_halo_cluster == 2 && (p.n_tiles_m % 2) == 0 && !p.b_resident) return BN == 64 ? launch_halo<64, 2>(ta, tb, p, stream) : launch_halo<128, 2>(ta, tb, p, stream);
        return BN == 64 ? launch_halo<64, 1>(ta, tb, p, stream) : launch_halo<128, 1>(ta, tb, p, stream);
    }
    // tile rectangle: BW = largest power of two <= min(Wout,32), then rows, then images, product 128
    int BW = 1; while (BW * 2 <= Wout && BW * 2 <= 32) BW *= 2;
    int BH = 1; while (BH * 2 <= Hout && BW * BH * 2 <= 128) BH *= 2;
    int BNI = 128 / (BW * BH);
    if (BW * stride > 256 || BH * stride > 256 || BNI > 256) return "tile rectangle exceeds the TMA box limit";
    // 256 accumulator columns per MMA when the layer is wide enough (+30 % tensor throughput, profiles/run20_gemm.md)
    const int BN = Cout <= 64 ? 64 : ((Cout % 256 == 0 && g_gemm_bn != 128) ? 256 : 128);
    CUtensorMap ta, tb;
    uint64_t da[4] = {(uint64_t)Cin, (uint64_t)Win, (uint64_t)Hin, (uint64_t)NB};
    uint64_t sa[3] = {(uint64_t)Cin * 2, (uint64_t)Win * Cin * 2, (uint64_t)Hin * Win * Cin * 2};
    // with element strides the box extent is given in source elements: ceil(box/stride) elements are fetched
    uint32_t boxa[4] = {BK, (uint32_t)((BW - 1) * stride + 1), (uint32_t)((BH - 1) * stride + 1), (uint32_t)BNI};
    uint32_t esa[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
    const int Ktot = R * S * Cin;
    uint64_t db[2] = {(uint64_t)Ktot, (uint64_t)Cout}, sb[1] = {(uint64_t)Ktot * 2};
    uint32_t boxb[2] = {BK, (uint32_t)BN}, esb[2] = {1, 1};
    if (const char* e = make_tmap(&ta, X, 4, da, sa, boxa, esa)) return e;
    if (const char* e = make_tmap(&tb, Wt, 2, db, sb, boxb, esb)) return e;
    TcParams p{};
    p.M = NB * Hout * Wout; p.N = Cout; p.num_kb = Ktot / BK; p.mode = 1;
    p.Hout = Hout; p.Wout = Wout; p.NB = NB; p.stride = stride; p.pad = pad; p.S = S; p.cblocks = Cin / BK;
    p.BW = BW; p.BH = BH; p.BNI = BNI;
    p.tiles_w = (Wout + BW - 1) / BW; p.tiles_h = (Hout + BH - 1) / BH;
    const int tiles_n_img = (NB + BNI - 1) / BNI;
    p.n_tiles_m = p.tiles_w * p.tiles_h * tiles_n_img; p.n_tiles_n = (Cout + BN - 1) / BN;
    p.bias = (const __nv_bfloat16*)bias; p.residual = (const __nv_bfloat16*)residual; p.relu = relu; p.C = (__nv_bfloat16*)Y; p.err = err_word(); p.stats = stats;
    return BN == 64 ? launch<64>(ta, tb, p, stream) : (BN == 128 ? launch<128>(ta, tb, p, stream) : launch<256>(ta, tb, p, stream));
}

template <int BN, int MM, int KP>
static const char* launch_wgrad(const CUtensorMap& ta, const CUtensorMap& tb, const WgParams& p, cudaStream_t stream) {
    using Cfg = WgCfg<BN, MM, KP>;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(tc_wgrad_kernel<BN, MM, KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess) return "cannot raise dynamic smem";
        attr = true;
    }
    const int grid = p.n_items < wgrad_sms() ? p.n_items : wgrad_sms();
    launch_pdl(tc_wgrad_kernel<BN, MM, KP>, dim3(grid), dim3(kThreads), Cfg::kSmemBytes, stream, ta, tb, p);
    return cudaGetLastError() == cudaSuccess ? nullptr : "kernel launch failed";
}

// dW (bf16, [Cout][R][S][Cin]) from dY [NB,Ho,Wo,Cout] and X [NB,H,W,Cin]; scratch32 = zeroed fp32 buffer of Cout*R*S*Cin floats
const char* tc_conv_wgrad(const void* dY, const void* X, void* dW, float* scratch32, int NB, int Hin, int Win, int Cin, int Cout, int R, int S,
                          int stride, int pad, cudaStream_t stream) {
    if (Cin % 64 || Cout % 64) return "Cin and Cout must be multiples of 64";
    const int Ho = (Hin + 2 * pad - R) / stride + 1, Wo = (Win + 2 * pad - S) / stride + 1;
    const int MM = Cout % 128 == 0 ? 128 : 64;
    const int BN = (Cin % 256 == 0 && MM == 128) ? 256 : (Cin % 128 == 0 ? 128 : 64);     // 256-column MMAs for wide layers
    const int kWgKPix = BN == 256 ? 64 : 128;
    int BW = 1; while (BW * 2 <= Wo && BW * 2 <= 32) BW *= 2;
    int BH = 1; while (BH * 2 <= Ho && BW * BH * 2 <= kWgKPix) BH *= 2;
    const int BNI = kWgKPix / (BW * BH);
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
    WgParams p{};
    p.taps = R * S; p.S = S; p.stride = stride; p.pad = pad; p.Cin = Cin; p.Cout = Cout; p.NB = NB; p.Ho = Ho; p.Wo = Wo;
    p.BW = BW; p.BH = BH; p.BNI = BNI;
    p.tiles_w = (Wo + BW - 1) / BW; p.tiles_h = (Ho + BH - 1) / BH;
    p.kb_total = p.tiles_w * p.tiles_h * ((NB + BNI - 1) / BNI);
    p.co_tiles = Cout / MM; p.ci_tiles = Cin / BN;
    const int base_items = p.taps * p.co_tiles * p.ci_tiles;
    int ksplit = (2 * wgrad_sms() + base_items - 1) / base_items;
    if (ksplit < 1) ksplit = 1;
    if (ksplit > p.kb_total) ksplit = p.kb_total;
    p.kb_per_split = (p.kb_total + ksplit - 1) / ksplit;
    p.ksplit = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
    p.n_items = base_items * p.ksplit;
    p.dW32 = scratch32; p.err = err_word();
    const char* e;
    if (BN == 256) e = launch_wgrad<256, 128, 64>(ta, tb, p, stream);
    else if (MM == 128 && BN == 128) e = launch_wgrad<128, 128, 128>(ta, tb, p, stream);
    else if (MM == 128) e = launch_wgrad<64, 128, 128>(ta, tb, p, stream);
    else if (BN == 128) e = launch_wgrad<128, 64, 128>(ta, tb, p, stream);
    else e = launch_wgrad<64, 64, 128>(ta, tb, p, stream);
    if (e) return e;
    const long long n = (long long)Cout * R * S * Cin;
    launch_pdl(wgrad_finish_kernel, dim3((int)((n / 4 + 255) / 256 < 1184 ? (n / 4 + 255) / 256 : 1184)), dim3(256), 0, stream, scratch32, (__nv_bfloat16*)dW, n);
    return cudaGetLastError() == cudaSuccess ? nullptr : "finish kernel launch failed";
}

}  // namespace psb

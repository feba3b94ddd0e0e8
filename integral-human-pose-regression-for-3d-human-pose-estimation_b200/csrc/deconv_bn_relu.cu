// deconv_bn_relu.cu -- K9: a deconv block of HeadNet.deconv_layers at inference, ConvTranspose2d(C_in -> 256, k4 s2 p1, no bias) +
// BatchNorm2d (running statistics) + ReLU (/root/reference/main/model.py:22-38; the second and the third block of the head for the
// reference's 256 x 256 input: 16 x 16 -> 32 x 32 -> 64 x 64), as ONE tensor-core kernel; the last block writes the bf16 NHWC activation
// K3 (head_fused_fwd.cu) reads as its operand -- SURVEY section 8 row N1.
//
// A stride-2 4x4 transposed convolution is four independent stride-1 2x2 convolutions, one per output phase (py, px) = (oy & 1, ox & 1):
//     out[b, 2 y0 + py, 2 x0 + px, co] = sum_{ty, tx in {0,1}} sum_ci  x[b, y0 + dy(py, ty), x0 + dx(px, tx), ci] * w[ci, co, ky(py, ty), kx(px, tx)]
//     py = 0: (ky, dy) = (1, 0), (3, -1)        py = 1: (ky, dy) = (0, +1), (2, 0)            (same table for x)
// i.e. per phase a GEMM  [pixels x (4 taps * C_in)] . [(4 taps * C_in) x C_out]  whose A operand is the input shifted by (dy, dx) -- a
// plain 4-D TMA box of the NHWC input with out-of-bounds rows / columns zero-filled by the TMA unit (no im2col buffer, no halo code).
//
// One work item = (sample, phase, 8 input rows of a 32-wide map / 16 of a 16-wide one): TWO accumulators of 128 pixels x 256 output channels fill the 512 TMEM
// columns, so every 32 KiB weight k-block feeds 8 MMAs (the weights are the operand every item re-reads: 512 KiB per accumulator pair).
// Per k-step (one tap, 64 input channels): A0, A1 = 2 x [128 px x 64] (16 KiB each), B = [256 co x 64] (32 KiB); 16 k-steps per item,
// 3-stage ring.  Epilogue (16 warps): y = max(0, acc * scale[co] + shift[co]) in fp32 (BatchNorm folded to scale / shift by the prep
// kernel), bf16; every warp stages its own 32 pixels x 32 channels and stores them by 5-D TMA into the strided phase positions of the
// NHWC output.
// Clusters (CS = 2 or 4 CTAs, opt-in): half of the operand stream (64 KiB per k-step) is the weight k-block every CTA of the same
// phase reads.  The CTAs of a cluster take the same phase and adjacent 8-row groups of one sample, each loads 1 / CS of every weight
// k-block and TMA-multicasts it into all CS shared memories; a stage is refilled only when the MMA issuers of ALL CS CTAs have released
// it (tcgen05.commit multicast onto every CTA's `empty` barrier).  Parity-green; on B200 the kernel turned out NOT to be bound by that
// stream (knock-out without any loads: -4 %), so clusters of 2 only tie and clusters of 4 lose to cluster scheduling -- the default is
// CS = 1.  What bounds it: the epilogue cannot overlap the main loop (both accumulators fill tensor memory) and costs ~5.5 k clk per
// item next to 16.4 k clk of MMAs (profiles/r02_k9_deconv_bench.txt).
// The last, partial wave of work items is cut into HALF items (one accumulator) when that shortens the schedule: 512 items
// on 148 SMs are 3 waves of full items + one wave of 136 halves instead of 4 waves; a batch of 4 is one wave of 128 halves.
//
// Training (the training side) reuses the same kernel through the MODE template parameter:
//   kTrain  the forward of the block under BatchNorm's batch statistics: the epilogue stores the RAW convolution output (bf16) and every
//           thread keeps running sums (sum y, sum y^2 of the bf16-rounded values it just staged, read back transposed from its warp's
//           staging piece) for its channel pair; at the end each warp writes one row of per-channel partials.  bn_train.cu reduces the
//           rows in fixed order (deterministic), forms mean / rstd / scale / shift and applies normalise + ReLU in one streaming pass.
//   kDgrad  d loss / d input of the transposed convolution = a stride-2 4x4 convolution of the output gradient: ONE GEMM per 128 input
//           pixels, [pixels x (16 taps * C_out)] . [(16 taps * C_out) x C_in]; tap (ky, kx) reads the gradient at (2 iy + ky - 1,
//           2 ix + kx - 1) = phase plane ((ky - 1) & 1, (kx - 1) & 1) shifted by ((ky - 1) >> 1, (kx - 1) >> 1) -- again a zero-filled TMA
//           box, now of the 5-D phase view {2 C (px, c), W, 2 (py), H, B} of the NHWC gradient.  64 k-steps per item, plain bf16 epilogue.
#include "head_tc.cuh"

namespace ihpr {
namespace k9 {

enum Mode { kInfer = 0, kTrain = 1, kDgrad = 2 };

using namespace tc;

constexpr int BM = 128;                 // pixels per accumulator: 128 / Win input rows x Win columns
constexpr int BN = 256;                 // output channels (one UMMA N)
constexpr int BK = 64;
// input width Win = 32 or 16 (the x extent of the TMA boxes); an accumulator is rows = 128 / Win input rows (4 or 8)
constexpr int NACC = 2;
constexpr int STAGES = 3;
constexpr int A_BYTES = BM * BK * 2;    // 16 KiB
constexpr int B_BYTES = BN * BK * 2;    // 32 KiB
constexpr int STAGE_BYTES = NACC * A_BYTES + B_BYTES;   // 64 KiB
constexpr int STG_BLK_BYTES = BM * 128; // 2 x 16 KiB of staging = 16 warps x 2 KiB
constexpr int EPI_WARPS = 16;
constexpr uint32_t TMEM_COLS = 512;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 2 * STG_BLK_BYTES + 256;
constexpr uint32_t kIdesc = make_idesc(BM, BN);

struct Params {
    int B, Hin, KB;             // KB = C_in / 64
    int rows;                   // input rows per accumulator: 128 / Win
    int wrows;                  // input rows per epilogue warp (32 pixels): 32 / Win
    int items;                  // cluster items: B * 4 phases * (Hin / (2 * rows)) / CS
    int full_items;             // cluster work units [0, full_items) are whole items; unit full_items + h is half (h & 1) of item full_items + h / 2
    int units;                  // full_items + 2 * (items - full_items)
    const float* scale;         // (256): gamma / sqrt(var + eps)
    const float* shift;         // (256): beta - mean * scale
    float* stat_part;           // kTrain: [gridDim.x * 4 rows][2 (sum, sum of squares)][256] partial batch statistics, one row per (CTA, TMEM lane quarter)
    int dbg;                    // -DIHPR_TIMING_EXPERIMENTS builds only (IHPR_K9_DEBUG): 1 = no epilogue work, 2 = no operand loads, 4 = no MMAs, 8 = no output stores -- WRONG results
};

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
                 : "memory");
}
// the box lands at the same offset of EVERY CTA in `mask`, and each of them gets the complete_tx on its own barrier at `bar`'s offset
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
// arrive on the barrier at `bar`'s offset in every CTA of `mask` when all MMAs this thread has issued are complete
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
                 "r"(c4), "r"(smem_u32(src))
                 : "memory");
}

// item -> (sample, phase, first input row); phase fastest: neighbouring CTAs read the same input rows
struct Item {
    int b, py, px, y0, nacc;
};
// `unit` counts CLUSTER work units: one unit = CS items (same sample and phase, 8-row groups CS * g + rank); units [0, full) are whole
// items, unit full + h is half (h & 1) of cluster item full + h / 2
template <int PH>        // PH = 4 output phases per input position (forward), 1 (kDgrad: the item is just (sample, rows))
__device__ __forceinline__ Item decode(int unit, int full, int ygroups, int cs, int rank, int rows) {
    Item it;
    int item = unit, half = 0;
    it.nacc = NACC;
    if (unit >= full) {
        const int h = unit - full;
        item = full + (h >> 1);
        half = h & 1;
        it.nacc = 1;
    }
    const int ph = PH == 4 ? item & 3 : 0;
    const int r = PH == 4 ? item >> 2 : item;
    const int ygc = ygroups / cs;               // cluster 8-row groups per sample
    it.py = ph >> 1;
    it.px = ph & 1;
    it.b = r / ygc;
    it.y0 = ((r - it.b * ygc) * cs + rank) * (NACC * rows) + half * rows;
    return it;
}

//   map_x: input, 4-D {C_in, Win, Hin, B} bf16 NHWC, box {64, Win, 128 / Win, 1}, zero fill out of bounds
//   map_w: re-laid weights, 2-D {C_in, 16 * 256}: row (phase * 4 + tap) * 256 + co, box {64, 256 / CS}
//   map_y: output, 5-D {256, 2 (px), Win (x0), 2 (py), Hin * B (y0 of every sample)} bf16 NHWC, box {32, 1, Win, 1, 32 / Win}, SWIZZLE_64B
//   kDgrad: map_x = the output gradient, 5-D {2 * 256 (px, c), W, 2 (py), H, B}, box {64, W, 1, 128 / W, 1}, zero fill; map_w = weights re-laid
//   [tap][ci][co] as 2-D {C_out, 16 * C_in}, box {64, 256}; map_y = the input gradient, 5-D {C_in, 1, W, 1, H * B}, box {32, 1, W, 1, 32 / W}
template <int CS, int MODE>
__global__ void __launch_bounds__(32 * (4 + EPI_WARPS), 1)
deconv_bn_relu_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_y,
                      const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sRing = smem;                                  // [STAGES][A0 | A1 | B]
    uint8_t* sS = sRing + STAGES * STAGE_BYTES;             // [16 warps][32 px x 32 co] bf16 staging for the output stores
    uint64_t* bars = reinterpret_cast<uint64_t*>(sS + 2 * STG_BLK_BYTES);
    uint64_t* full = bars;                      // [STAGES] TMA -> MMA
    uint64_t* empty = full + STAGES;            // [STAGES] MMA -> TMA
    uint64_t* acc_full = empty + STAGES;        // [1] MMA -> epilogue
    uint64_t* acc_empty = acc_full + 1;         // [2] epilogue -> MMA, one per accumulator (EPI_WARPS arrivals): accumulator 0 is drained first
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + NACC);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ygroups = p.Hin / (NACC * p.rows);
    constexpr int PH = MODE == kDgrad ? 1 : 4;
    constexpr int TAPS = MODE == kDgrad ? 16 : 4;
    const int ksteps = TAPS * p.KB;
    const int rank = CS > 1 ? (int)cluster_ctarank() : 0;
    const int cluster = (int)blockIdx.x / CS, nclusters = (int)gridDim.x / CS;
    constexpr uint16_t kAll = (uint16_t)((1u << CS) - 1);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, CS); }
        mbar_init(acc_full, 1);
        for (int a = 0; a < NACC; ++a) mbar_init(acc_empty + a, EPI_WARPS);
        mbar_fence_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (CS > 1) cluster_sync_all();     // every CTA's barriers exist before a peer multicasts into this CTA or signals it
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int unit = cluster; unit < p.units; unit += nclusters) {
                const Item w = decode<PH>(unit, p.full_items, ygroups, CS, rank, p.rows);
                for (int tap = 0; tap < TAPS; ++tap) {
                    // forward: tap (ty, tx) of this phase reads the input shifted by (dy, dx); kDgrad: tap (ky, kx) reads phase plane (qy, qx) of
                    // the output gradient shifted by (dy, dx)
                    int dy, dx, qy = 0, qx = 0, wrow;
                    if (MODE == kDgrad) {
                        const int ky = tap >> 2, kx = tap & 3;
                        dy = (ky - 1) >> 1; qy = (ky - 1) & 1;
                        dx = (kx - 1) >> 1; qx = (kx - 1) & 1;
                        wrow = tap * BN;
                    } else {
                        const int ty = tap >> 1, tx = tap & 1;
                        dy = w.py ? 1 - ty : -ty;               // py = 0: 0, -1;  py = 1: +1, 0
                        dx = w.px ? 1 - tx : -tx;
                        wrow = ((w.py * 2 + w.px) * 4 + tap) * BN;
                    }
                    for (int kb = 0; kb < p.KB; ++kb, ++it) {
                        const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                        mbar_wait(empty + s, ph ^ 1);
#ifdef IHPR_TIMING_EXPERIMENTS
                        if (CS == 1 && (p.dbg & 2)) { mbar_arrive(full + s); continue; }
#endif
                        mbar_expect_tx(full + s, (uint32_t)(w.nacc * A_BYTES + B_BYTES));
                        uint8_t* st = sRing + s * STAGE_BYTES;
                        if (MODE == kDgrad) {
                            tma_load_5d(st, &map_x, qx * (p.KB * BK) + kb * BK, dx, qy, w.y0 + dy, w.b, full + s);
                            if (w.nacc == NACC) tma_load_5d(st + A_BYTES, &map_x, qx * (p.KB * BK) + kb * BK, dx, qy, w.y0 + p.rows + dy, w.b, full + s);
                        } else {
                            tma_load_4d(st, &map_x, kb * BK, dx, w.y0 + dy, w.b, full + s);
                            if (w.nacc == NACC) tma_load_4d(st + A_BYTES, &map_x, kb * BK, dx, w.y0 + p.rows + dy, w.b, full + s);
                        }
                        // this CTA's 1 / CS of the weight k-block goes to every CTA of the cluster (their stage s is free: `empty` counts all CS issuers)
                        if (CS > 1) tma_load_2d_mc(st + NACC * A_BYTES + rank * (B_BYTES / CS), &map_w, kb * BK, wrow + rank * (BN / CS), full + s, kAll);
                        else tma_load_2d(st + NACC * A_BYTES, &map_w, kb * BK, wrow, full + s);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            // The epilogue of the previous item drains accumulator 0 first and hands it back ~2.7 k clk before accumulator 1, so an item
            // STARTS SKEWED: the accumulator-0 MMAs of the first STAGES k-steps are issued as soon as accumulator 0 is free (their stages are
            // already loaded), the accumulator-1 MMAs of those k-steps follow when accumulator 1 is, then both run interleaved per k-step.
            uint32_t it = 0;
            uint32_t n_acc[NACC] = {0, 0};             // items that have used each accumulator (phase of its acc_empty barrier)
            auto mma_ks = [&](uint32_t i, int ks, int acc) {       // the four MMAs of k-step ks (ring position i) into one accumulator
                const uint32_t st = smem_u32(sRing + (i % STAGES) * STAGE_BYTES);
                const uint64_t ad = umma_desc(st + acc * A_BYTES), bd = umma_desc(st + NACC * A_BYTES);
#ifdef IHPR_TIMING_EXPERIMENTS
                if (p.dbg & 4) return;
#endif
#pragma unroll
                for (int k16 = 0; k16 < BK / 16; ++k16) umma(tmem_base + acc * BN, ad + 2 * k16, bd + 2 * k16, kIdesc, (uint32_t)((ks | k16) != 0));
            };
            auto release = [&](uint32_t i) {
                if (CS > 1) tc_commit_mc(empty + i % STAGES, kAll);
                else tc_commit(empty + i % STAGES);
            };
            for (int unit = cluster; unit < p.units; unit += nclusters) {
                const bool both = unit < p.full_items;
                const int skew = both ? (ksteps < STAGES ? ksteps : STAGES) : 0;
                mbar_wait(acc_empty, (n_acc[0]++ & 1) ^ 1);         // the epilogue has pulled the previous item out of accumulator 0
                tc_fence_after();
                for (int ks = 0; ks < skew; ++ks) {
                    mbar_wait(full + (it + ks) % STAGES, ((it + ks) / STAGES) & 1);
                    tc_fence_after();
                    mma_ks(it + ks, ks, 0);
                }
                if (both) {
                    mbar_wait(acc_empty + 1, (n_acc[1]++ & 1) ^ 1);
                    tc_fence_after();
                }
                for (int ks = 0; ks < skew; ++ks) {
                    mma_ks(it + ks, ks, 1);
                    release(it + ks);
                }
                it += skew;
                for (int ks = skew; ks < ksteps; ++ks, ++it) {
                    mbar_wait(full + it % STAGES, (it / STAGES) & 1);
                    tc_fence_after();
                    mma_ks(it, ks, 0);
                    if (both) mma_ks(it, ks, 1);
                    release(it);
                }
                tc_commit(acc_full);
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue: BatchNorm scale / shift, ReLU, bf16, TMA store into the phase positions =================
        // Every warp works alone: its TMEM lane quarter is ONE input row (32 pixels) of the accumulator, its column group 64 output
        // channels; it stages 32 pixels x 32 channels (2 KiB of its own) and stores them with its own TMA box -- no CTA-wide barrier
        // anywhere in the epilogue (a first version with 512-thread named barriers around a shared staging block spent 1650 clk per
        // 64-channel round, 13 k clk per item against 16 k clk of MMA work: profiles/r02_k9_deconv_bench.txt).
        const int e = warp - 4;
        const int qd = warp & 3;                    // TMEM lane quarter (hardware: warp id % 4) = input row within the accumulator
        const int cg = e >> 2;                      // 64 of the 256 output channels
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        uint8_t* stg = sS + e * 2048;               // [32 pixels][64 bytes], SWIZZLE_64B: 16-byte chunk ^= (pixel >> 1) & 3
        const int swz = (lane >> 1) & 3;
        uint32_t n = 0;
        float st_acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};      // kTrain: [channel half][sum c, sum c^2, sum c+1, sum (c+1)^2] of this lane's channel pair
        for (int unit = cluster; unit < p.units; unit += nclusters, ++n) {
            const Item w = decode<PH>(unit, p.full_items, ygroups, CS, rank, p.rows);
            mbar_wait(acc_full, n & 1);
            tc_fence_after();
#ifdef IHPR_TIMING_EXPERIMENTS
            if (p.dbg & 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive(acc_empty); if (w.nacc == NACC) mbar_arrive(acc_empty + 1); }
                continue;
            }
#endif
            // (Pulling the whole item out of tensor memory first -- 64 packed registers per thread -- so that the stores run underneath
            // the next item's main loop was tried: at 96 registers per thread it spills ~1 KiB per thread.)
            // round = (accumulator a, channel half hs of this warp's 64 output channels)
            auto round = [&](const int a, const int hs) {
                {
                    const int co0 = cg * 64 + hs * 32;
                    float v[32];
                    tmem_ld32(tmem_base + lane_off + (uint32_t)(a * BN + co0), v);
                    if (hs) {                                   // this accumulator is in registers / stored: the next item's MMAs into it may start
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(acc_empty + a);
                    }
                    uint32_t o[16];
                    if (MODE == kInfer) {
#pragma unroll
                        for (int i4 = 0; i4 < 8; ++i4) {
                            const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale + co0) + i4);
                            const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift + co0) + i4);
                            const float y0 = fmaxf(fmaf(v[4 * i4], sc.x, sh.x), 0.f), y1 = fmaxf(fmaf(v[4 * i4 + 1], sc.y, sh.y), 0.f);
                            const float y2 = fmaxf(fmaf(v[4 * i4 + 2], sc.z, sh.z), 0.f), y3 = fmaxf(fmaf(v[4 * i4 + 3], sc.w, sh.w), 0.f);
                            o[2 * i4] = Elem<__nv_bfloat16>::pk(y0, y1);
                            o[2 * i4 + 1] = Elem<__nv_bfloat16>::pk(y2, y3);
                        }
                    } else {
#pragma unroll
                        for (int i2 = 0; i2 < 16; ++i2) o[i2] = Elem<__nv_bfloat16>::pk(v[2 * i2], v[2 * i2 + 1]);
                    }
                    if (lane == 0) tma_store_wait_read();       // this warp's previous store has read the staging piece
                    __syncwarp();
                    uint8_t* srow = stg + lane * 64;
#pragma unroll
                    for (int j = 0; j < 4; ++j) sts16(srow + ((j ^ swz) << 4), make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]));
                    fence_async_smem();
                    __syncwarp();
#ifdef IHPR_TIMING_EXPERIMENTS
                    if (p.dbg & 8) return;
#endif
                    if (lane == 0) {
                        tma_store_5d(&map_y, stg, co0, w.px, 0, w.py, w.b * p.Hin + w.y0 + a * p.rows + qd * p.wrows);
                        tma_store_commit();
                    }
                    if (MODE == kTrain) {
                        // batch statistics of what was just staged (the bf16 values BatchNorm will normalise), read back transposed: lane =
                        // (channel pair cp, pixel half hh); the two half-warps walk rows of opposite parity, i.e. disjoint bank halves
                        const int cp = lane & 15, hh = lane >> 4;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int pix = hh * 16 + (i ^ hh);
                            uint32_t wd;
                            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(wd) : "r"(smem_u32(stg + pix * 64 + ((((cp >> 2) ^ (pix >> 1)) & 3) << 4) + (cp & 3) * 4)));
                            const float y0 = __uint_as_float(wd << 16), y1 = __uint_as_float(wd & 0xffff0000u);
                            st_acc[hs][0] += y0;
                            st_acc[hs][1] = fmaf(y0, y0, st_acc[hs][1]);
                            st_acc[hs][2] += y1;
                            st_acc[hs][3] = fmaf(y1, y1, st_acc[hs][3]);
                        }
                    }
                }
            };
            if (MODE == kTrain) {           // hs must be a compile-time constant here: it indexes the statistics registers
#pragma unroll 1
                for (int a = 0; a < w.nacc; ++a) { round(a, 0); round(a, 1); }
            } else {
                const int rounds = 2 * w.nacc;
#pragma unroll 1
                for (int r = 0; r < rounds; ++r) round(r >> 1, r & 1);
            }
        }
        if (MODE == kTrain) {
            // one row of partials per (CTA, lane quarter): [sum | sum of squares][256 channels]; the half-warps are added in a fixed order
            const int cp = lane & 15;
            float* row = p.stat_part + ((size_t)blockIdx.x * 4 + qd) * 512;
#pragma unroll
            for (int hs = 0; hs < 2; ++hs) {
                float t[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) t[k] = st_acc[hs][k] + __shfl_xor_sync(0xffffffffu, st_acc[hs][k], 16);
                if (lane < 16) {
                    const int ch = cg * 64 + hs * 32 + cp * 2;
                    *reinterpret_cast<float2*>(row + ch) = make_float2(t[0], t[2]);
                    *reinterpret_cast<float2*>(row + 256 + ch) = make_float2(t[1], t[3]);
                }
            }
        }
        if (lane == 0) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (CS > 1) cluster_sync_all();     // no CTA leaves while a peer may still signal its barriers
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
}

// weights (C_in, C_out, 4, 4) bf16 -> [phase][tap][co][ci] bf16 (K-major B operand per (phase, tap)); BatchNorm -> scale / shift.
// One thread per (co, ci): reads its 16 taps as one 32-byte sector, writes each of the 16 (phase, tap) planes with ci fastest.
__global__ void deconv_prep_kernel(const __nv_bfloat16* __restrict__ w, int Cin, int Cout, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   const float* __restrict__ mean, const float* __restrict__ var, float eps, __nv_bfloat16* __restrict__ wp,
                                   float* __restrict__ scale, float* __restrict__ shift, __nv_bfloat16* __restrict__ wp_dgrad) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (wp && idx < Cout * Cin) {
        const int ci = idx % Cin, co = idx / Cin;
        const uint4* src = reinterpret_cast<const uint4*>(w + ((size_t)ci * Cout + co) * 16);
        const uint4 lo = __ldg(src), hi = __ldg(src + 1);
        const uint32_t words[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
        for (int pt = 0; pt < 16; ++pt) {
            const int tap = pt & 3, phase = pt >> 2;
            const int py = phase >> 1, px = phase & 1, ty = tap >> 1, tx = tap & 1;
            const int ky = py ? 2 * ty : 1 + 2 * ty;        // py = 0: 1, 3;  py = 1: 0, 2
            const int kx = px ? 2 * tx : 1 + 2 * tx;
            const int e = ky * 4 + kx;
            const uint16_t bits = (uint16_t)(words[e >> 1] >> ((e & 1) * 16));
            reinterpret_cast<uint16_t*>(wp)[((size_t)pt * Cout + co) * Cin + ci] = bits;
        }
    }
    // training: the same weights as the B operand of the input-gradient GEMM, [tap = ky * 4 + kx][ci][co] (co = its K, contiguous); one thread per
    // (ci, co) with co fastest, so that the 32-byte reads AND the 2-byte writes of a warp are contiguous
    if (wp_dgrad && idx < Cout * Cin) {
        const int co = idx % Cout, ci = idx / Cout;
        const uint4* src = reinterpret_cast<const uint4*>(w + ((size_t)ci * Cout + co) * 16);
        const uint4 lo = __ldg(src), hi = __ldg(src + 1);
        const uint32_t words[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
        for (int e = 0; e < 16; ++e)
            reinterpret_cast<uint16_t*>(wp_dgrad)[((size_t)e * Cin + ci) * Cout + co] = (uint16_t)(words[e >> 1] >> ((e & 1) * 16));
    }
    if (gamma && idx < Cout) {
        const float sc = __ldg(gamma + idx) * rsqrtf(__ldg(var + idx) + eps);
        scale[idx] = sc;
        shift[idx] = __ldg(beta + idx) - __ldg(mean + idx) * sc;
    }
}

}  // namespace k9

// ---- host side --------------------------------------------------------------------------------------------------
size_t deconv_workspace_bytes(int Cin, int Cout) { return (size_t)16 * Cin * Cout * 2 + 2 * (size_t)Cout * sizeof(float) + 512; }

static bool encode(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
                   CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    tc::EncodeTiledFn enc = tc::encode_tiled();
    if (!enc) return false;
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static void carve(void* workspace, int Cin, int Cout, __nv_bfloat16** wp, float** scale, float** shift) {
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    *wp = reinterpret_cast<__nv_bfloat16*>(ws);
    *scale = reinterpret_cast<float*>(ws + (((size_t)16 * Cin * Cout * 2 + 255) / 256) * 256);
    *shift = *scale + Cout;
}

// once per set of parameters: the workspace then holds everything launch_deconv_bn_relu needs besides the input
void launch_deconv_prepare(const void* weight, const float* gamma, const float* beta, const float* mean, const float* var, float eps, int Cin, int Cout,
                           void* workspace, int* launches, cudaStream_t s) {
    __nv_bfloat16* wp;
    float *scale, *shift;
    carve(workspace, Cin, Cout, &wp, &scale, &shift);
    const int n = Cin * Cout, th = 256;
    k9::deconv_prep_kernel<<<(n + th - 1) / th, th, 0, s>>>(static_cast<const __nv_bfloat16*>(weight), Cin, Cout, gamma, beta, mean, var, eps, wp, scale, shift,
                                                            nullptr);
    ++*launches;
}

// training: the weights change every step -- re-lay them for the forward GEMM (wp_fwd, [phase][tap][co][ci]) and / or the input-gradient
// GEMM (wp_dgrad, [tap][ci][co]); no BatchNorm fold (batch statistics do not exist yet)
void launch_deconv_relayout(const void* weight, int Cin, int Cout, void* wp_fwd, void* wp_dgrad, int* launches, cudaStream_t s) {
    const int n = Cin * Cout, th = 256;
    k9::deconv_prep_kernel<<<(n + th - 1) / th, th, 0, s>>>(static_cast<const __nv_bfloat16*>(weight), Cin, Cout, nullptr, nullptr, nullptr, nullptr, 0.f,
                                                            static_cast<__nv_bfloat16*>(wp_fwd), nullptr, nullptr, static_cast<__nv_bfloat16*>(wp_dgrad));
    ++*launches;
}

template <int CS, int MODE>
static const char* launch_k9(const CUtensorMap& mx, const CUtensorMap& mw, const CUtensorMap& my, k9::Params p, int T, int num_sms, cudaStream_t s, int* grid_out) {
    using namespace k9;
    auto kern = deconv_bn_relu_kernel<CS, MODE>;
    const size_t smem = SMEM_BYTES + 1024;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return "cudaFuncSetAttribute failed (deconv_bn_relu_kernel)";
    // whole waves of full items, then the rest as half items if that ends sooner (time in half-item waves); G = clusters that run at once
    const int G = num_sms / CS;
    p.items = T / CS;
    const int whole = (p.items / G) * G, rest = p.items - whole;
    const int t_full = 2 * ((p.items + G - 1) / G), t_half = 2 * (p.items / G) + (2 * rest + G - 1) / G;
    p.full_items = t_half < t_full ? whole : p.items;
    p.units = p.full_items + 2 * (p.items - p.full_items);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(CS * (p.units < G ? p.units : G)));
    cfg.blockDim = dim3(32 * (4 + EPI_WARPS));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (grid_out) *grid_out = (int)cfg.gridDim.x;
    if (cudaLaunchKernelEx(&cfg, kern, mx, mw, my, p) != cudaSuccess) return "deconv_bn_relu_kernel launch failed";
    return nullptr;
}

static bool forward_maps(const void* x_nhwc, const void* wp, void* y_nhwc, int B, int Cin, int Cout, int Hin, int Win, int cs, CUtensorMap* map_x,
                         CUtensorMap* map_w, CUtensorMap* map_y, const char** err) {
    using namespace k9;
    const int rows = BM / Win;
    {
        const cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)B};
        const cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)Win * Cin * 2, (cuuint64_t)Hin * Win * Cin * 2};
        const cuuint32_t box[4] = {BK, (cuuint32_t)Win, (cuuint32_t)rows, 1};
        if (!encode(map_x, x_nhwc, 4, dims, strides, box)) { *err = "cuTensorMapEncodeTiled failed for the deconv input"; return false; }
    }
    if (!tc::make_map(map_w, wp, (uint64_t)16 * Cout, (uint64_t)Cin, BN / cs)) { *err = "cuTensorMapEncodeTiled failed for the deconv weights"; return false; }
    {
        const cuuint64_t dims[5] = {(cuuint64_t)Cout, 2, (cuuint64_t)Win, 2, (cuuint64_t)Hin * B};
        const cuuint64_t strides[4] = {(cuuint64_t)Cout * 2, (cuuint64_t)2 * Cout * 2, (cuuint64_t)2 * Win * Cout * 2, (cuuint64_t)2 * 2 * Win * Cout * 2};
        const cuuint32_t box[5] = {32, 1, (cuuint32_t)Win, 1, (cuuint32_t)(32 / Win)};
        if (!encode(map_y, y_nhwc, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B)) { *err = "cuTensorMapEncodeTiled failed for the deconv output"; return false; }
    }
    return true;
}

// cluster: 0 = default (one CTA per SM, no cluster), else 1 / 2 / 4 CTAs per cluster (halved until it divides the 8-row groups of a sample)
const char* launch_deconv_bn_relu(const void* x_nhwc, const void* prepared, int B, int Cin, int Cout, int Hin, int Win, void* y_nhwc, int num_sms, int cluster,
                                  int* launches, cudaStream_t s) {
    using namespace k9;
    __nv_bfloat16* wp;
    float *scale, *shift;
    carve(const_cast<void*>(prepared), Cin, Cout, &wp, &scale, &shift);
    const int rows = BM / Win, ygroups = Hin / (NACC * rows);
    int cs = cluster;
    if (cs != 1 && cs != 2 && cs != 4) cs = 1;      // measured on B200: clusters of 2 tie (62.0 vs 59.9 us at B = 32), clusters of 4 lose (121 us)
    while (cs > 1 && (ygroups % cs != 0 || num_sms < cs)) cs >>= 1;
    CUtensorMap map_x, map_w, map_y;
    const char* err = nullptr;
    if (!forward_maps(x_nhwc, wp, y_nhwc, B, Cin, Cout, Hin, Win, cs, &map_x, &map_w, &map_y, &err)) return err;
    Params p = {};
    p.B = B; p.Hin = Hin; p.KB = Cin / BK;
    p.rows = rows; p.wrows = 32 / Win;
    p.scale = scale; p.shift = shift;
    p.dbg = 0;
#ifdef IHPR_TIMING_EXPERIMENTS
    if (const char* e = getenv("IHPR_K9_DEBUG")) p.dbg = atoi(e);
#endif
    const int T = B * 4 * ygroups;
    err = cs == 4 ? launch_k9<4, kInfer>(map_x, map_w, map_y, p, T, num_sms, s, nullptr) : cs == 2 ? launch_k9<2, kInfer>(map_x, map_w, map_y, p, T, num_sms, s, nullptr)
                                                                                                   : launch_k9<1, kInfer>(map_x, map_w, map_y, p, T, num_sms, s, nullptr);
    if (err) return err;
    ++*launches;
    return cudaGetLastError() == cudaSuccess ? nullptr : "deconv_bn_relu_kernel launch failed";
}

// training forward: y_raw = conv_transpose2d(x, w) (bf16 NHWC) and stat_part[*stat_rows][2][256] = partial per-channel sums of y_raw, y_raw^2
const char* launch_deconv_train_fwd(const void* x_nhwc, const void* wp_fwd, int B, int Cin, int Cout, int Hin, int Win, void* y_raw_nhwc, float* stat_part,
                                    int* stat_rows, int num_sms, int* launches, cudaStream_t s) {
    using namespace k9;
    const int rows = BM / Win, ygroups = Hin / (NACC * rows);
    CUtensorMap map_x, map_w, map_y;
    const char* err = nullptr;
    if (!forward_maps(x_nhwc, wp_fwd, y_raw_nhwc, B, Cin, Cout, Hin, Win, 1, &map_x, &map_w, &map_y, &err)) return err;
    Params p = {};
    p.B = B; p.Hin = Hin; p.KB = Cin / BK;
    p.rows = rows; p.wrows = 32 / Win;
    p.stat_part = stat_part;
    int grid = 0;
    err = launch_k9<1, kTrain>(map_x, map_w, map_y, p, B * 4 * ygroups, num_sms, s, &grid);
    if (err) return err;
    *stat_rows = grid * 4;
    ++*launches;
    return cudaGetLastError() == cudaSuccess ? nullptr : "deconv_bn_relu_kernel (training forward) launch failed";
}

// d loss / d input of the transposed convolution: dy (B, 2 H, 2 W, C_out) bf16 NHWC -> dx (B, H, W, C_in) bf16 NHWC; C_in = C_out = 256
const char* launch_deconv_dgrad(const void* dy_nhwc, const void* wp_dgrad, int B, int Cin, int Cout, int Hin, int Win, void* dx_nhwc, int num_sms, int* launches,
                                cudaStream_t s) {
    using namespace k9;
    const int rows = BM / Win, ygroups = Hin / (NACC * rows);
    CUtensorMap map_x, map_w, map_y;
    {
        // phase view of the output gradient: element (c, px, x0, py, y0, b) at ((b * 2H + 2 y0 + py) * 2W + 2 x0 + px) * C + c; (px, c) merged
        const cuuint64_t C2 = (cuuint64_t)Cout * 2;       // bytes per output pixel
        const cuuint64_t dims[5] = {(cuuint64_t)2 * Cout, (cuuint64_t)Win, 2, (cuuint64_t)Hin, (cuuint64_t)B};
        const cuuint64_t strides[4] = {2 * C2, (cuuint64_t)2 * Win * C2, (cuuint64_t)2 * 2 * Win * C2, (cuuint64_t)2 * Hin * 2 * Win * C2};
        const cuuint32_t box[5] = {BK, (cuuint32_t)Win, 1, (cuuint32_t)rows, 1};
        if (!encode(&map_x, dy_nhwc, 5, dims, strides, box)) return "cuTensorMapEncodeTiled failed for the deconv output gradient";
    }
    if (!tc::make_map(&map_w, wp_dgrad, (uint64_t)16 * Cin, (uint64_t)Cout, BN)) return "cuTensorMapEncodeTiled failed for the deconv weights (dgrad)";
    {
        const cuuint64_t dims[5] = {(cuuint64_t)Cin, 1, (cuuint64_t)Win, 1, (cuuint64_t)Hin * B};
        const cuuint64_t strides[4] = {(cuuint64_t)Cin * 2, (cuuint64_t)Cin * 2, (cuuint64_t)Win * Cin * 2, (cuuint64_t)Win * Cin * 2};
        const cuuint32_t box[5] = {32, 1, (cuuint32_t)Win, 1, (cuuint32_t)(32 / Win)};
        if (!encode(&map_y, dx_nhwc, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_64B)) return "cuTensorMapEncodeTiled failed for the deconv input gradient";
    }
    Params p = {};
    p.B = B; p.Hin = Hin; p.KB = Cout / BK;
    p.rows = rows; p.wrows = 32 / Win;
    const char* err = launch_k9<1, kDgrad>(map_x, map_w, map_y, p, B * ygroups, num_sms, s, nullptr);
    if (err) return err;
    ++*launches;
    return cudaGetLastError() == cudaSuccess ? nullptr : "deconv_bn_relu_kernel (input gradient) launch failed";
}

}  // namespace ihpr

// deconv_wgrad.cu -- K11: the weight gradient of a deconv block of HeadNet, ConvTranspose2d(256 -> 256, k4 s2 p1) (/root/reference/main/model.py:22-38
// under main/train.py:71), on the tensor cores.  SURVEY section 8 row N1, the last GEMM of the block's backward.
//
// With the sub-pixel decomposition of deconv_bn_relu.cu (output phase (py, px), tap (ty, tx) -> input shift (dy, dx)) the forward is
//     out[b, 2 y0 + py, 2 x0 + px, co] = sum_{tap} sum_ci  x[b, y0 + dy, x0 + dx, ci] * wp[phase][tap][co][ci]
// so   dwp[phase][tap][co][ci] = sum_{b, y0, x0}  dout[b, 2 y0 + py, 2 x0 + px, co] * x[b, y0 + dy, x0 + dx, ci] :
// 16 GEMMs  D[co, ci] = A^T B  whose contraction runs over the PIXELS.  Both operands are pixel-major tiles [64 pixels x 256 channels] exactly as TMA
// delivers them from the NHWC tensors (A = a box of the 5-D phase view of the gradient, B = the shifted, zero-filled 4-D box of the input the
// forward uses), i.e. MN-major UMMA operands: 64 channels contiguous (one SWIZZLE_128B row), 8-pixel groups 1 KiB apart, 64-channel blocks 8 KiB apart.
//
// One CTA = one (phase, tap) and one contiguous range of 64-pixel tiles (the batch is split over gridDim.x / 16 CTAs per (phase, tap)); its
// accumulator is the whole [256 co x 256 ci] fp32 tile = two 128-lane halves x 256 columns = all 512 TMEM columns.  Per tile: 64 KiB of operands
// (3-stage ring), 8 MMAs (M = 128, N = 256, K = 16).  Epilogue: the fp32 tile goes to a partial buffer [split][phase * 4 + tap][ci][co] (co fastest:
// a warp's TMEM lanes are consecutive co); wgrad_reduce_kernel adds the splits in index order (deterministic) and writes dW (C_in, C_out, 4, 4) fp32.
// Arithmetic intensity: a [256 x 256] output tile is the largest tensor memory holds, so every pixel costs 1 KiB of L2 -> SM traffic for 131 kFLOP
// (128 FLOP / B; 537 MB per call at B = 32): the kernel sits on the L2 -> SM path, not on the tensor pipe.
#include "head_tc.cuh"

namespace ihpr {
namespace k11 {

using namespace tc;

constexpr int C = 256;                      // C_in = C_out
// BKP = pixels per stage (contraction), 64 or 32: one [BKP pixels x 64 channels] box is SUB bytes, an operand [BKP x 256] four of them, a stage
// A | B; the ring always holds 192 KiB (3 stages of 64 KiB or 6 of 32 KiB)
constexpr int RING_BYTES = 192 * 1024;
constexpr int EPI_WARPS = 16;
constexpr uint32_t TMEM_COLS = 512;
constexpr size_t SMEM_BYTES = (size_t)RING_BYTES + 256;
// D = f32, A = B = bf16, both MN-major (bits 15, 16), M = 128, N = 256
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

struct Params {
    int B, Hin;
    int rows;               // input rows per tile: BKP / Win
    int tiles;              // B * Hin / rows: tiles per (phase, tap)
    int splits;             // CTAs per (phase, tap) = gridDim.x / 16
    float* part;            // [splits][16][C ci][C co] fp32
};

// MN-major SWIZZLE_128B descriptor (see head_fused_bwd.cu): LBO = distance between 64-channel blocks, SBO = distance between 8-pixel groups
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((1024 >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar))
                 : "memory");
}

// the box lands at the same offset of EVERY CTA in `mask`, and each of them gets the complete_tx on its own barrier at `bar`'s offset
__device__ __forceinline__ void tma_load_5d_mc(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4, uint64_t* bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3, %4, %5, %6}], [%7], %8;" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar)), "h"(mask)
        : "memory");
}
// arrive on the barrier at `bar`'s offset in every CTA of `mask` when all MMAs this thread has issued are complete
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}

// Clusters (CS = 2 or 4 CTAs): the CTAs of a cluster are consecutive taps of ONE phase on the same pixel range, so their A operand (the gradient
// tile, half of the operand stream) is the same: each CTA loads 1 / CS of its four 64-channel boxes and TMA-multicasts them into all CS shared
// memories; a stage is refilled only when the MMA issuers of all CS CTAs have released it (tcgen05.commit multicast onto every `empty` barrier).
// L2 -> SM traffic per call: 537 MB (CS = 1), 403 MB (CS = 2), 336 MB (CS = 4) at B = 32.  Measured on B200 (profiles/r02_k11_wgrad_bench.txt): parity-green and
// NOT faster (50.4 us alone, 52.1 us in pairs, ~84 us in fours; 32-pixel stages in a 6-deep ring 58.8 us) -- the kernel runs at 80 % of the MMA rate and is
// not bound by that stream, so the default is CS = 1 with 64-pixel stages; the cluster / fine-stage forms stay as opt-in variants.
//   map_x : input, 4-D {C, Win, Hin, B} bf16 NHWC, box {64, Win, 64 / Win, 1}, zero fill out of bounds
//   map_dy: output gradient, 5-D phase view {2 C (px, c), Win, 2 (py), Hin, B}, box {64, Win, 1, 64 / Win, 1}
template <int CS, int BKP>
__global__ void __launch_bounds__(32 * (4 + EPI_WARPS), 1)
deconv_wgrad_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy, const Params p) {
    constexpr int SUB_BYTES = BKP * 64 * 2, OP_BYTES = (C / 64) * SUB_BYTES, STAGE_BYTES = 2 * OP_BYTES, STAGES = RING_BYTES / STAGE_BYTES;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sRing = smem;                                  // [STAGES][A: 4 boxes | B: 4 boxes]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sRing + STAGES * STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = full + STAGES;
    uint64_t* acc_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pt = (int)blockIdx.x & 15, split = (int)blockIdx.x >> 4;      // (phase, tap) fastest: neighbouring CTAs read the same pixels
    const int phase = pt >> 2, tap = pt & 3;
    const int py = phase >> 1, px = phase & 1, ty = tap >> 1, tx = tap & 1;
    const int dy = py ? 1 - ty : -ty, dx = px ? 1 - tx : -tx;               // the forward's shift table (deconv_bn_relu.cu)
    const int t_lo = (int)((long long)p.tiles * split / p.splits), t_hi = (int)((long long)p.tiles * (split + 1) / p.splits);
    const int ygroups = p.Hin / p.rows;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, CS); }
        mbar_init(acc_full, 1);
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
    const int rank = CS > 1 ? (int)cluster_ctarank() : 0;          // = tap % CS (1-D cluster of consecutive blocks)
    constexpr uint16_t kAll = (uint16_t)((1u << CS) - 1);

    if (warp == 0) {
        // ================= TMA producer =================
        // A stage is eight 8 KiB boxes (SWIZZLE_128B limits a box to 64 channels): lanes 0 ... 7 issue one box each, so the stage's loads go
        // out together instead of one thread paying eight issue latencies (the single-thread form was the bound: 50.4 us at B = 32).
        uint32_t it = 0;
        const int cb = lane & 3;                    // the 64-channel block this lane loads
        for (int t = t_lo; t < t_hi; ++t, ++it) {
            const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
            const int b = t / ygroups, y0 = (t - b * ygroups) * p.rows;
            mbar_wait(empty + s, ph ^ 1);
            if (lane == 0) mbar_expect_tx(full + s, (uint32_t)STAGE_BYTES);
            __syncwarp();
            uint8_t* st = sRing + s * STAGE_BYTES;
            if (lane < 4) {
                if (CS > 1) {
                    // this CTA's 1 / CS of the gradient tile goes to every CTA of the cluster
                    if (cb / (C / 64 / CS) == rank) tma_load_5d_mc(st + cb * SUB_BYTES, &map_dy, px * C + cb * 64, 0, py, y0, b, full + s, kAll);
                } else {
                    tma_load_5d(st + cb * SUB_BYTES, &map_dy, px * C + cb * 64, 0, py, y0, b, full + s);
                }
            } else if (lane < 8) {
                tma_load_4d(st + OP_BYTES + cb * SUB_BYTES, &map_x, cb * 64, dx, y0 + dy, b, full + s);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = t_lo; t < t_hi; ++t, ++it) {
                const uint32_t s = it % STAGES;
                mbar_wait(full + s, (it / STAGES) & 1);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(sRing + s * STAGE_BYTES), b_addr = a_addr + OP_BYTES;
#pragma unroll
                for (int k16 = 0; k16 < BKP / 16; ++k16) {
                    // 16 contraction rows (pixels) = 2 groups of 8 = +2 KiB in both operands; the second co half starts two 64-channel blocks later
                    const uint64_t bd = desc_mn(b_addr + k16 * 2048, SUB_BYTES);
                    const uint32_t acc = (uint32_t)((it | k16) != 0);
                    umma(tmem_base, desc_mn(a_addr + k16 * 2048, SUB_BYTES), bd, kIdesc, acc);
                    umma(tmem_base + 256, desc_mn(a_addr + 2 * SUB_BYTES + k16 * 2048, SUB_BYTES), bd, kIdesc, acc);
                }
                if (CS > 1) tc_commit_mc(empty + s, kAll);
                else tc_commit(empty + s);
            }
            tc_commit(acc_full);
        }
    } else if (warp >= 4) {
        // ================= epilogue: the fp32 tile -> this CTA's partial, [ci][co] with co fastest =================
        const int e = warp - 4;
        const int qd = warp & 3;                    // TMEM lane quarter (hardware: warp id % 4)
        const int cg = e >> 2;                      // 64 of the 256 ci columns
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        float* part = p.part + ((size_t)split * 16 + pt) * C * C;
        mbar_wait(acc_full, 0);
        tc_fence_after();
#pragma unroll 1
        for (int r = 0; r < 4; ++r) {
            const int h = r >> 1, ci0 = cg * 64 + (r & 1) * 32;
            float v[32];
            tmem_ld32(tmem_base + lane_off + (uint32_t)(h * 256 + ci0), v);
            float* dst = part + (size_t)ci0 * C + h * 128 + qd * 32 + lane;
#pragma unroll
            for (int i = 0; i < 32; ++i) dst[(size_t)i * C] = v[i];
        }
    }

    tc_fence_before();
    __syncthreads();
    if (CS > 1) cluster_sync_all();     // no CTA leaves while a peer may still multicast into it or signal its barriers
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
}

// dW[ci][co][ky][kx] = sum over the splits (index order) of part[split][phase * 4 + tap][ci][co]; one thread per (ci, co), co fastest
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ part, int splits, float* __restrict__ dw) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;          // ci * C + co
    if (idx >= C * C) return;
    float acc[16];
#pragma unroll
    for (int pt = 0; pt < 16; ++pt) acc[pt] = 0.f;
    for (int sp = 0; sp < splits; ++sp) {           // 16 independent loads in flight per thread; the splits are added in index order
#pragma unroll
        for (int pt = 0; pt < 16; ++pt) acc[pt] += __ldg(part + ((size_t)sp * 16 + pt) * C * C + idx);
    }
    float res[16];
#pragma unroll
    for (int pt = 0; pt < 16; ++pt) {
        const int tap = pt & 3, phase = pt >> 2;
        const int py = phase >> 1, px = phase & 1, ty = tap >> 1, tx = tap & 1;
        const int ky = py ? 2 * ty : 1 + 2 * ty;        // the forward's tap table (deconv_prep_kernel)
        const int kx = px ? 2 * tx : 1 + 2 * tx;
        res[ky * 4 + kx] = acc[pt];
    }
    float4* out = reinterpret_cast<float4*>(dw + (size_t)idx * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) out[q] = make_float4(res[4 * q], res[4 * q + 1], res[4 * q + 2], res[4 * q + 3]);
}

}  // namespace k11

// ---- host side --------------------------------------------------------------------------------------------------
constexpr int kWgradMaxSplits = 16;
size_t deconv_wgrad_workspace_bytes() { return (size_t)kWgradMaxSplits * 16 * k11::C * k11::C * sizeof(float); }

static bool encode_bf16(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box) {
    tc::EncodeTiledFn enc = tc::encode_tiled();
    if (!enc) return false;
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int CS, int BKP>
static const char* launch_k11(const CUtensorMap& mx, const CUtensorMap& mdy, const k11::Params& p, cudaStream_t s) {
    using namespace k11;
    auto kern = deconv_wgrad_kernel<CS, BKP>;
    const size_t smem = SMEM_BYTES + 1024;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return "cudaFuncSetAttribute failed (deconv_wgrad_kernel)";
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(16 * p.splits));
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
    if (cudaLaunchKernelEx(&cfg, kern, mx, mdy, p) != cudaSuccess) return "deconv_wgrad_kernel launch failed";
    return nullptr;
}

// cluster: CTAs per cluster (1, 2 or 4; anything else = the default)
// x (B, Hin, Win, 256) bf16 NHWC, dy (B, 2 Hin, 2 Win, 256) bf16 NHWC -> dw (256, 256, 4, 4) fp32; workspace: deconv_wgrad_workspace_bytes()
const char* launch_deconv_wgrad(const void* x_nhwc, const void* dy_nhwc, int B, int Hin, int Win, float* dw, void* workspace, int num_sms, int cluster,
                                int* launches, cudaStream_t s) {
    // cluster: 1 / 2 / 4 = CTAs per cluster with 64-pixel stages (3-deep ring), 11 / 12 / 14 = the same with 32-pixel stages (6-deep ring); else default
    const bool fine = cluster > 10;
    int cs = fine ? cluster - 10 : cluster;
    if (cs != 1 && cs != 2 && cs != 4) cs = 1;
    const int BKP = fine ? 32 : 64;
    using namespace k11;
    const int rows = BKP / Win;
    if (rows < 1 || Hin % rows != 0) return "deconv_wgrad: the map does not divide into whole tiles";
    CUtensorMap map_x, map_dy;
    {
        const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)B};
        const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)Win * C * 2, (cuuint64_t)Hin * Win * C * 2};
        const cuuint32_t box[4] = {64, (cuuint32_t)Win, (cuuint32_t)rows, 1};
        if (!encode_bf16(&map_x, x_nhwc, 4, dims, strides, box)) return "cuTensorMapEncodeTiled failed for the deconv input (wgrad)";
    }
    {
        const cuuint64_t C2 = (cuuint64_t)C * 2;          // bytes per output pixel
        const cuuint64_t dims[5] = {(cuuint64_t)2 * C, (cuuint64_t)Win, 2, (cuuint64_t)Hin, (cuuint64_t)B};
        const cuuint64_t strides[4] = {2 * C2, (cuuint64_t)2 * Win * C2, (cuuint64_t)2 * 2 * Win * C2, (cuuint64_t)2 * Hin * 2 * Win * C2};
        const cuuint32_t box[5] = {64, (cuuint32_t)Win, 1, (cuuint32_t)rows, 1};
        if (!encode_bf16(&map_dy, dy_nhwc, 5, dims, strides, box)) return "cuTensorMapEncodeTiled failed for the deconv output gradient (wgrad)";
    }
    Params p;
    p.B = B; p.Hin = Hin; p.rows = rows;
    p.tiles = B * (Hin / rows);
    int splits = num_sms / 16;
    if (splits < 1) splits = 1;
    if (splits > kWgradMaxSplits) splits = kWgradMaxSplits;
    if (splits > p.tiles) splits = p.tiles;
    p.splits = splits;
    p.part = static_cast<float*>(workspace);
    const char* err = fine ? (cs == 4 ? launch_k11<4, 32>(map_x, map_dy, p, s) : cs == 2 ? launch_k11<2, 32>(map_x, map_dy, p, s) : launch_k11<1, 32>(map_x, map_dy, p, s))
                           : (cs == 4 ? launch_k11<4, 64>(map_x, map_dy, p, s) : cs == 2 ? launch_k11<2, 64>(map_x, map_dy, p, s) : launch_k11<1, 64>(map_x, map_dy, p, s));
    if (err) return err;
    wgrad_reduce_kernel<<<C * C / 256, 256, 0, s>>>(p.part, splits, dw);
    *launches += 2;
    return cudaGetLastError() == cudaSuccess ? nullptr : "wgrad_reduce_kernel launch failed";
}

}  // namespace ihpr

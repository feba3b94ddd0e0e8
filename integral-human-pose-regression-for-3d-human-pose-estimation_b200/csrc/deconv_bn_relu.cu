// deconv_bn_relu.cu -- K9: the last stage of HeadNet.deconv_layers at inference, ConvTranspose2d(256 -> 256, k4 s2 p1, no bias) +
// BatchNorm2d (running statistics) + ReLU (/root/reference/main/model.py:22-38 with num_layers = 3: the third block, 32 x 32 -> 64 x 64),
// as ONE tensor-core kernel that writes the bf16 NHWC activation K3 (head_fused_fwd.cu) reads as its operand -- SURVEY section 8 row N1.
//
// A stride-2 4x4 transposed convolution is four independent stride-1 2x2 convolutions, one per output phase (py, px) = (oy & 1, ox & 1):
//     out[b, 2 y0 + py, 2 x0 + px, co] = sum_{ty, tx in {0,1}} sum_ci  x[b, y0 + dy(py, ty), x0 + dx(px, tx), ci] * w[ci, co, ky(py, ty), kx(px, tx)]
//     py = 0: (ky, dy) = (1, 0), (3, -1)        py = 1: (ky, dy) = (0, +1), (2, 0)            (same table for x)
// i.e. per phase a GEMM  [pixels x (4 taps * C_in)] . [(4 taps * C_in) x C_out]  whose A operand is the input shifted by (dy, dx) -- a
// plain 4-D TMA box of the NHWC input with out-of-bounds rows / columns zero-filled by the TMA unit (no im2col buffer, no halo code).
//
// One work item = (sample, phase, 8 input rows): TWO accumulators of 128 pixels (4 rows x 32) x 256 output channels fill the 512 TMEM
// columns, so every 32 KiB weight k-block feeds 8 MMAs (the weights are the operand every item re-reads: 512 KiB per accumulator pair).
// Per k-step (one tap, 64 input channels): A0, A1 = 2 x [128 px x 64] (16 KiB each), B = [256 co x 64] (32 KiB); 16 k-steps per item,
// 3-stage ring.  Epilogue (16 warps): y = max(0, acc * scale[co] + shift[co]) in fp32 (BatchNorm folded to scale / shift by the prep
// kernel), bf16, staged 128 px x 64 co at a time in two alternating 16 KiB blocks and stored by 5-D TMA into the strided phase
// positions of the NHWC output (the store of round r is only waited for in round r + 2).
// The last, partial wave of work items is cut into HALF items (one accumulator, 4 input rows) when that shortens the schedule: 512 items
// on 148 SMs are 3 waves of full items + one wave of 136 halves instead of 4 waves; a batch of 4 is one wave of 128 halves.
#include "head_tc.cuh"

namespace ihpr {
namespace k9 {

using namespace tc;

constexpr int BM = 128;                 // pixels per accumulator: 4 input rows x 32 columns
constexpr int BN = 256;                 // output channels (one UMMA N)
constexpr int BK = 64;
constexpr int WIN = 32;                 // input width (= the x extent of the TMA box)
constexpr int ROWS = 4;                 // input rows per accumulator
constexpr int NACC = 2;
constexpr int STAGES = 3;
constexpr int A_BYTES = BM * BK * 2;    // 16 KiB
constexpr int B_BYTES = BN * BK * 2;    // 32 KiB
constexpr int STAGE_BYTES = NACC * A_BYTES + B_BYTES;   // 64 KiB
constexpr int STG_BLK_BYTES = BM * 128; // 16 KiB: 128 pixels x 64 output channels of bf16
constexpr int EPI_WARPS = 16;
constexpr uint32_t TMEM_COLS = 512;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 2 * STG_BLK_BYTES + 256;
constexpr uint32_t kIdesc = make_idesc(BM, BN);

struct Params {
    int B, Hin, KB;             // KB = C_in / 64
    int items;                  // B * 4 phases * (Hin / 8)
    int full_items;             // work units [0, full_items) are whole items; unit full_items + h is half (h & 1) of item full_items + h / 2
    int units;                  // full_items + 2 * (items - full_items)
    const float* scale;         // (256): gamma / sqrt(var + eps)
    const float* shift;         // (256): beta - mean * scale
};

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
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
__device__ __forceinline__ Item decode(int unit, int full_items, int ygroups) {
    Item it;
    int item = unit, half = 0;
    it.nacc = NACC;
    if (unit >= full_items) {
        const int h = unit - full_items;
        item = full_items + (h >> 1);
        half = h & 1;
        it.nacc = 1;
    }
    const int ph = item & 3;
    const int r = item >> 2;
    it.py = ph >> 1;
    it.px = ph & 1;
    it.b = r / ygroups;
    it.y0 = (r - it.b * ygroups) * (NACC * ROWS) + half * ROWS;
    return it;
}

//   map_x: input, 4-D {C_in, 32, Hin, B} bf16 NHWC, box {64, 32, 4, 1}, zero fill out of bounds
//   map_w: re-laid weights, 2-D {C_in, 16 * 256}: row (phase * 4 + tap) * 256 + co, box {64, 256}
//   map_y: output, 5-D {256, 2 (px), 32 (x0), 2 (py), Hin * B (y0 of every sample)} bf16 NHWC, box {64, 1, 32, 1, 4}
__global__ void __launch_bounds__(32 * (4 + EPI_WARPS), 1)
deconv_bn_relu_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_y,
                      const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sRing = smem;                                  // [STAGES][A0 | A1 | B]
    uint8_t* sS = sRing + STAGES * STAGE_BYTES;             // [2][128 px x 64 co] bf16 staging for the output store
    uint64_t* bars = reinterpret_cast<uint64_t*>(sS + 2 * STG_BLK_BYTES);
    uint64_t* full = bars;                      // [STAGES] TMA -> MMA
    uint64_t* empty = full + STAGES;            // [STAGES] MMA -> TMA
    uint64_t* acc_full = empty + STAGES;        // [1] MMA -> epilogue
    uint64_t* acc_empty = acc_full + 1;         // [1] epilogue -> MMA (EPI_WARPS arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ygroups = p.Hin / (NACC * ROWS);
    const int ksteps = 4 * p.KB;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, EPI_WARPS);
        mbar_fence_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            uint32_t it = 0;
            for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
                const Item w = decode(unit, p.full_items, ygroups);
                for (int tap = 0; tap < 4; ++tap) {
                    const int ty = tap >> 1, tx = tap & 1;
                    const int dy = w.py ? 1 - ty : -ty;         // py = 0: 0, -1;  py = 1: +1, 0
                    const int dx = w.px ? 1 - tx : -tx;
                    const int wrow = ((w.py * 2 + w.px) * 4 + tap) * BN;
                    for (int kb = 0; kb < p.KB; ++kb, ++it) {
                        const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                        mbar_wait(empty + s, ph ^ 1);
                        mbar_expect_tx(full + s, (uint32_t)(w.nacc * A_BYTES + B_BYTES));
                        uint8_t* st = sRing + s * STAGE_BYTES;
                        tma_load_4d(st, &map_x, kb * BK, dx, w.y0 + dy, w.b, full + s);
                        if (w.nacc == NACC) tma_load_4d(st + A_BYTES, &map_x, kb * BK, dx, w.y0 + ROWS + dy, w.b, full + s);
                        tma_load_2d(st + NACC * A_BYTES, &map_w, kb * BK, wrow, full + s);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            uint32_t it = 0, n = 0;
            for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x, ++n) {
                const bool both = unit < p.full_items;
                mbar_wait(acc_empty, (n & 1) ^ 1);          // the epilogue has pulled the previous item out of tensor memory
                tc_fence_after();
                for (int ks = 0; ks < ksteps; ++ks, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(full + s, ph);
                    tc_fence_after();
                    const uint32_t st = smem_u32(sRing + s * STAGE_BYTES);
                    const uint64_t a0 = umma_desc(st), a1 = umma_desc(st + A_BYTES), bd = umma_desc(st + NACC * A_BYTES);
#pragma unroll
                    for (int k16 = 0; k16 < BK / 16; ++k16) {
                        umma(tmem_base, a0 + 2 * k16, bd + 2 * k16, kIdesc, (uint32_t)((ks | k16) != 0));
                        if (both) umma(tmem_base + BN, a1 + 2 * k16, bd + 2 * k16, kIdesc, (uint32_t)((ks | k16) != 0));
                    }
                    tc_commit(empty + s);
                }
                tc_commit(acc_full);
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue: BatchNorm scale / shift, ReLU, bf16, TMA store into the phase positions =================
        const int e = warp - 4;
        const int qd = warp & 3;                    // TMEM lane quarter
        const int cg = e >> 2;                      // 16 of the 64 output channels of a round
        const int row = qd * 32 + lane;             // pixel of the accumulator = TMEM lane = staging row
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        const int sw = row & 7;
        const bool elected = e == 0 && lane == 0;
        uint32_t n = 0, rr = 0;                     // rr: running round count = which staging block
        for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x, ++n) {
            const Item w = decode(unit, p.full_items, ygroups);
            const int rounds = 4 * w.nacc;          // round = (accumulator, 64 output channels)
            mbar_wait(acc_full, n & 1);
            tc_fence_after();
#pragma unroll 1
            for (int r = 0; r < rounds; ++r, ++rr) {
                const int a = r >> 2;
                const int co0 = (r & 3) * 64 + cg * 16;
                uint32_t v[16];
                tmem_ld16_issue(tmem_base + lane_off + (uint32_t)(a * BN + co0), v);
                float4 sc[4], sh[4];
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                    sc[i4] = __ldg(reinterpret_cast<const float4*>(p.scale + co0) + i4);
                    sh[i4] = __ldg(reinterpret_cast<const float4*>(p.shift + co0) + i4);
                }
                tmem_ld16_wait(v);
                if (r == rounds - 1) {                      // the accumulators are in registers / stored: the next item's MMAs may start
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty);
                }
                uint32_t o[8];
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                    const float y0 = fmaxf(fmaf(__uint_as_float(v[4 * i4]), sc[i4].x, sh[i4].x), 0.f);
                    const float y1 = fmaxf(fmaf(__uint_as_float(v[4 * i4 + 1]), sc[i4].y, sh[i4].y), 0.f);
                    const float y2 = fmaxf(fmaf(__uint_as_float(v[4 * i4 + 2]), sc[i4].z, sh[i4].z), 0.f);
                    const float y3 = fmaxf(fmaf(__uint_as_float(v[4 * i4 + 3]), sc[i4].w, sh[i4].w), 0.f);
                    o[2 * i4] = Elem<__nv_bfloat16>::pk(y0, y1);
                    o[2 * i4 + 1] = Elem<__nv_bfloat16>::pk(y2, y3);
                }
                // staging block rr & 1 was last read by the store of round rr - 2: at most the store of round rr - 1 may still be in flight
                if (elected) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                named_bar_sync(1, EPI_WARPS * 32);
                uint8_t* blk = sS + (rr & 1) * STG_BLK_BYTES;
                uint8_t* srow = blk + row * 128;
                sts16(srow + (((2 * cg) ^ sw) << 4), make_uint4(o[0], o[1], o[2], o[3]));
                sts16(srow + (((2 * cg + 1) ^ sw) << 4), make_uint4(o[4], o[5], o[6], o[7]));
                fence_async_smem();
                named_bar_sync(2, EPI_WARPS * 32);
                if (elected) {
                    tma_store_5d(&map_y, blk, (r & 3) * 64, w.px, 0, w.py, w.b * p.Hin + w.y0 + a * ROWS);
                    tma_store_commit();
                }
            }
        }
        if (elected) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
}

// weights (C_in, C_out, 4, 4) bf16 -> [phase][tap][co][ci] bf16 (K-major B operand per (phase, tap)); BatchNorm -> scale / shift.
// One thread per (co, ci): reads its 16 taps as one 32-byte sector, writes each of the 16 (phase, tap) planes with ci fastest.
__global__ void deconv_prep_kernel(const __nv_bfloat16* __restrict__ w, int Cin, int Cout, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   const float* __restrict__ mean, const float* __restrict__ var, float eps, __nv_bfloat16* __restrict__ wp,
                                   float* __restrict__ scale, float* __restrict__ shift) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < Cout * Cin) {
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
    if (idx < Cout) {
        const float sc = __ldg(gamma + idx) * rsqrtf(__ldg(var + idx) + eps);
        scale[idx] = sc;
        shift[idx] = __ldg(beta + idx) - __ldg(mean + idx) * sc;
    }
}

}  // namespace k9

// ---- host side --------------------------------------------------------------------------------------------------
size_t deconv_workspace_bytes(int Cin, int Cout) { return (size_t)16 * Cin * Cout * 2 + 2 * (size_t)Cout * sizeof(float) + 512; }

static bool encode(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box) {
    tc::EncodeTiledFn enc = tc::encode_tiled();
    if (!enc) return false;
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
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
    k9::deconv_prep_kernel<<<(n + th - 1) / th, th, 0, s>>>(static_cast<const __nv_bfloat16*>(weight), Cin, Cout, gamma, beta, mean, var, eps, wp, scale, shift);
    ++*launches;
}

const char* launch_deconv_bn_relu(const void* x_nhwc, const void* prepared, int B, int Cin, int Cout, int Hin, int Win, void* y_nhwc, int num_sms, int* launches,
                                  cudaStream_t s) {
    using namespace k9;
    (void)Win;
    __nv_bfloat16* wp;
    float *scale, *shift;
    carve(const_cast<void*>(prepared), Cin, Cout, &wp, &scale, &shift);
    CUtensorMap map_x, map_w, map_y;
    {
        const cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)WIN, (cuuint64_t)Hin, (cuuint64_t)B};
        const cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)WIN * Cin * 2, (cuuint64_t)Hin * WIN * Cin * 2};
        const cuuint32_t box[4] = {BK, WIN, ROWS, 1};
        if (!encode(&map_x, x_nhwc, 4, dims, strides, box)) return "cuTensorMapEncodeTiled failed for the deconv input";
    }
    if (!tc::make_map(&map_w, wp, (uint64_t)16 * Cout, (uint64_t)Cin, BN)) return "cuTensorMapEncodeTiled failed for the deconv weights";
    {
        const cuuint64_t dims[5] = {(cuuint64_t)Cout, 2, (cuuint64_t)WIN, 2, (cuuint64_t)Hin * B};
        const cuuint64_t strides[4] = {(cuuint64_t)Cout * 2, (cuuint64_t)2 * Cout * 2, (cuuint64_t)2 * WIN * Cout * 2, (cuuint64_t)2 * 2 * WIN * Cout * 2};
        const cuuint32_t box[5] = {BK, 1, WIN, 1, ROWS};
        if (!encode(&map_y, y_nhwc, 5, dims, strides, box)) return "cuTensorMapEncodeTiled failed for the deconv output";
    }
    Params p;
    p.B = B; p.Hin = Hin; p.KB = Cin / BK;
    p.items = B * 4 * (Hin / (NACC * ROWS));
    // whole waves of full items, then the rest as half items if that ends sooner (time in half-item waves)
    {
        const int G = num_sms, T = p.items;
        const int whole = (T / G) * G, rest = T - whole;
        const int t_full = 2 * ((T + G - 1) / G), t_half = 2 * (T / G) + (2 * rest + G - 1) / G;
        p.full_items = t_half < t_full ? whole : T;
        p.units = p.full_items + 2 * (T - p.full_items);
    }
    p.scale = scale; p.shift = shift;
    const size_t smem = SMEM_BYTES + 1024;
    if (cudaFuncSetAttribute(deconv_bn_relu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return "cudaFuncSetAttribute failed (deconv_bn_relu_kernel)";
    deconv_bn_relu_kernel<<<p.units < num_sms ? p.units : num_sms, 32 * (4 + EPI_WARPS), smem, s>>>(map_x, map_w, map_y, p);
    ++*launches;
    return cudaGetLastError() == cudaSuccess ? nullptr : "deconv_bn_relu_kernel launch failed";
}

}  // namespace ihpr

// head_fused_bwd.cu -- K4w / K4x: the WHOLE backward of HeadNet.final_layer + soft_argmax + JointLocationLoss
// (/root/reference/main/model.py:14-20,42 under main/train.py:67-71) on the tensor cores, with neither the heat-map nor its
// gradient ever stored: d loss / d weight, d loss / d bias and d loss / d x leave the kernels, nothing (B, J*D, H, W)-sized exists.
//
// Both kernels are the same two-GEMM chain (flash-attention-backward shaped):
//
//   GEMM1   D1[128 x 256] = A1 . B1^T          the heat-map tile, recomputed (K = C_in <= 256), accumulator in TMEM
//   epilogue dH = p * sum_c g_c (c(i) - coord_c)  from D1 and the (m, l, coords) the forward saved; bf16, written into a
//           shared-memory staging buffer in the K-major SWIZZLE_128B layout of an A operand
//   GEMM2   D2[128 x C_in] += dH . B2          accumulated in a second TMEM region over all the tiles of the work item
//
//   K4w (dW, dbias): rows = 128 output channels (one channel tile), D1 columns = 256 pixels;  A1 = W tile (stationary), B1 = X
//        k-blocks (NHWC, K-major), B2 = the SAME X pixels read as an MN-major operand (N = C_in, contraction = pixels);
//        D2 = dW[128 x C_in] summed over all pixels of the sample; per-(sample, channel tile) fp32 partials are reduced over the
//        batch in fixed order by head_bwd_reduce_kernel (deterministic, no atomics).  Row sums of dH give dbias.
//   K4x (dX): rows = 128 pixels, D1 columns = 256 output channels (H^T);  A1 = X tile (stationary), B1 = W k-blocks (K-major),
//        B2 = W rows read MN-major (N = C_in, contraction = channels);  D2 = dX[128 px x C_in] summed over ALL channel tiles,
//        converted to bf16 and stored NHWC by TMA.
//
// Why two kernels (5 GEMM passes with the forward's) instead of one: dW sums over pixels (channel-stationary work), dX sums over
// channels (pixel-stationary work); one traversal keeps only one of them in TMEM (128 KB each next to the 128 KB heat-map tile),
// and spilling the other as fp32 partials costs 128-256 KB per 50 MFLOP tile step -- more shared-memory and L2 traffic than the
// MMA operands themselves.  Recomputing the K = 256 heat-map GEMM once more is cheaper.
//
// Pipeline per CTA (one per SM, persistent): warp 0 = TMA producer, warp 1 = MMA issuer (one lane), warp 2 = TMEM allocator,
// warps 4..19 = epilogue.  TMEM: D1 = columns [0, 256), D2 = [256, 512): no room to double-buffer D1, so the two GEMMs take
// turns on the tensor pipe instead.  The epilogue works through a tile in four 64-column steps (one staging block = one GEMM2
// k-block each, all 16 warps on 16 columns apiece).  It pulls the WHOLE of D1 into registers during its first step (three
// 16-column buffers per thread) and releases D1 at once, so GEMM1 of the next tile runs underneath the rest of the epilogue;
// every staged block is handed to the issuer as soon as it is written.  The issuer's order per tile t is
//     GEMM1(t+1), GEMM2(blocks 0..3 of t)
// i.e. GEMM2 lags one tile behind: the tensor pipe works on GEMM1(t+1) while the epilogue of tile t fills the four staging blocks
// (64 KiB: the whole bf16 gradient tile), then on GEMM2(t) while the epilogue of tile t+1 runs -- the per-tile period approaches
// max(epilogue, GEMM1 + GEMM2) instead of their sum.  (A clock64 trace of the first versions, profiles/r02_k4_trace.txt: 6600-6800 clk
// per tile against 4096 clk of MMA work -- with two staging blocks the epilogue waited on GEMM2 and GEMM2 on the epilogue in turn.)
// The stationary operand A1 is handed over per k-block: its slot is reloaded for the next item (by warp 3) as soon as the last
// tile's GEMM1 has read that k-block, not at the end of the item.
#include "head_tc.cuh"

namespace ihpr {
namespace k4 {

using namespace tc;

constexpr int BM = 128;                 // rows of both accumulators (TMEM lanes)
constexpr int BN = 256;                 // columns of the heat-map tile
constexpr int BK = 64;                  // one SWIZZLE_128B row of bf16
constexpr int MAXKB = 4;                // C_in <= 256
constexpr int STAGES = 3;               // operand ring: 32 KiB stages shared by B1 k-blocks and B2 blocks, strictly FIFO in the issuer's order
constexpr int STAGES_PAIR = 6;          // SM-pair form: each CTA holds HALF of every stage (16 KiB), so the same 96 KiB are six stages deep
constexpr int NSTG = 4;                 // staging blocks: the whole bf16 gradient tile (64 KiB) -- GEMM2 of tile t runs AFTER GEMM1 of tile t+1
constexpr int A_KB_BYTES = BM * BK * 2;         // 16 KiB
constexpr int STAGE_BYTES = BN * BK * 2;        // 32 KiB
constexpr int B2_SUB_BYTES = BK * BK * 2;       // 8 KiB: one [64 contraction rows x 64 n] box of an MN-major B2 block
constexpr int STG_BLK_BYTES = BM * BK * 2;      // 16 KiB: staging block = 128 rows x 64 columns of dH (one A-operand k-block)
constexpr int NBLK = BN / BK;                   // 4 staging blocks per tile
constexpr int EPI_WARPS = 16;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t D2_COL = 256;

constexpr size_t SMEM_BYTES = (size_t)MAXKB * A_KB_BYTES + (size_t)STAGES * STAGE_BYTES + (size_t)NSTG * STG_BLK_BYTES + 512;

struct Params {
    int B, K, J, D, H, W;
    int M;                  // J * D output channels
    int Mpad;               // M rounded up to 256: row length of k0tab
    int Jpad;               // Mpad / D: row length of jtab
    int KB;                 // K / 64
    int MT;                 // ceil(M / 128)     (K4w)
    int NT;                 // H*W / 256         (K4w: pixel tiles per item)
    int PT;                 // H*W / 128         (K4x: items per sample)
    int CB;                 // ceil(M / 256)     (K4x: channel blocks per item)
    const float* k0tab;     // (B, Mpad): bias[c] * log2e - m[b, joint(c)] * log2e; -inf for c >= M (weight 0)
    const float4* jtab;     // (B, Jpad): {gx, gy, gz, -(gx cx + gy cy + gz cz)} with g pre-divided by l; zeros for joints >= J
    float* dw_part;         // K4w out: (B * MT, K, 128) fp32 partial d loss / d weight per (sample, channel tile), channel fastest
    float* db_part;         // K4w out: (B, 4, Mpad) fp32 partial d loss / d bias
    int dbg;                // -DIHPR_TIMING_EXPERIMENTS builds only (IHPR_K4_DEBUG): 1 = no GEMM2 operand loads, 2 = no epilogue math,
                            // 4 = no GEMM1 operand loads, 8 = no GEMM2 MMAs -- WRONG results, timing experiments only
    long long* trace;       // same builds (IHPR_K4_TRACE=1): clock64 stamps of CTA 0, 16 slots per tile: 0-4 issuer, 8-14 epilogue warp 0
};

#ifdef IHPR_TIMING_EXPERIMENTS
#define K4_STAMP(tile, slot) do { if (p.trace && blockIdx.x == 0 && (tile) < 64) p.trace[(tile) * 16 + (slot)] = clock64(); } while (0)
#else
#define K4_STAMP(tile, slot) do { } while (0)
#endif

// MN-major SWIZZLE_128B shared-memory matrix descriptor: 64 consecutive N (or M) elements are contiguous (128 B); LBO = byte
// distance between consecutive 64-element blocks along N; SBO = byte distance between groups of 8 contraction rows
// (cute::UMMA canonical layout  Swizzle<3,4,3> o ((8,n),(8,k)):((1,LBO),(8,SBO))  in 16-byte units)
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor with a runtime N and an MN-major B operand (bit 16)
__device__ __forceinline__ uint32_t idesc_bmn(int N) { return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(BM >> 4) << 24); }
constexpr uint32_t kIdesc1 = make_idesc(BM, BN);
constexpr uint32_t kIdesc1Pair = make_idesc(2 * BM, BN);        // cta_group::2: M = 256 over the two CTAs

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0), "r"(c1), "r"(smem_u32(src))
                 : "memory");
}
__device__ __forceinline__ uint4 ldg_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint64_t pk2u(uint32_t lo, uint32_t hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}

// DX = false: K4w (rows = channels).  DX = true: K4x (rows = pixels).
//   map_a : stationary A1, box 64 x 128 rows        (K4w: W,  K4x: X)
//   map_b1: GEMM1's streamed operand, box 64 x 256  (K4w: X,  K4x: W)
//   map_b2: GEMM2's streamed operand, box 64 x 64   (K4w: X,  K4x: W)  -- the same tensor as B1, other box
//   map_dx: K4x only: d loss / d x, (B*H*W, K) bf16, box 64 x 32 rows (one epilogue warp's slice)
// PAIR = true: the two CTAs of a cluster run every UMMA together (tcgen05.mma.cta_group::2, M = 256): each CTA owns its own 128 rows
// (its A1, its D1 / D2 in its own TMEM, its own epilogue and staging buffer) and loads only HALF of every streamed operand stage --
// GEMM1's B1 is split by its 256 N rows, GEMM2's MN-major B2 by its C_in columns (n-blocks) -- so the L2 -> SM traffic per SM, which is
// what bounds the single-CTA form (256 KiB per tile against ~42 B/clk of ingress: ~6200 clk per tile for 4096 clk of MMA work), halves.
// Rank 0 issues all MMAs; its barriers collect both CTAs' TMA bytes and epilogue arrivals; tcgen05.commit multicasts to both.
template <bool DX, bool PAIR>
__global__ void __launch_bounds__(32 * (4 + EPI_WARPS), 1)
head_bwd_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b1, const __grid_constant__ CUtensorMap map_b2,
                const __grid_constant__ CUtensorMap map_dx, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                                     // [KB][128 x 64] bf16, stationary per item
    constexpr int NST = PAIR ? STAGES_PAIR : STAGES;        // ring depth
    constexpr int STB = PAIR ? STAGE_BYTES / 2 : STAGE_BYTES;   // bytes of a stage in THIS CTA
    constexpr uint32_t NCTA = PAIR ? 2 : 1;
    uint8_t* sB = sA + MAXKB * A_KB_BYTES;                  // [NST][STB]
    uint8_t* sS = sB + NST * STB;                // [4][128 x 64] bf16: dH of one tile = GEMM2's A operand, one k-block per staging block
    uint64_t* bars = reinterpret_cast<uint64_t*>(sS + NSTG * STG_BLK_BYTES);
    uint64_t* b_full = bars;                    // [NST] TMA -> MMA
    uint64_t* b_empty = b_full + NST;           // [NST] MMA -> TMA
    uint64_t* a_full = b_empty + NST;           // [MAXKB] A1 k-block landed
    uint64_t* a_empty = a_full + MAXKB;         // [MAXKB] the item's last GEMM1 has read A1 k-block kb
    uint64_t* d1_full = a_empty + MAXKB;        // [1]  MMA -> epilogue: heat-map tile complete
    uint64_t* d1_empty = d1_full + 1;           // [1]  epilogue -> MMA: D1 is in registers (EPI_WARPS arrivals)
    uint64_t* s_full = d1_empty + 1;            // [4]  epilogue -> MMA: staging block written (EPI_WARPS arrivals); one phase per tile
    uint64_t* s_empty = s_full + NSTG;          // [4]  MMA -> epilogue: GEMM2 has read the staging block
    uint64_t* d2_full = s_empty + NSTG;         // [1]  MMA -> epilogue: the item's D2 is complete
    uint64_t* d2_empty = d2_full + 1;           // [1]  epilogue -> MMA: D2 drained (EPI_WARPS arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d2_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? pair_rank() : 0;           // 0 = leader (issues the MMAs)
    const int unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, nunits = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    // a work item of a pair = two adjacent 128-row items of the same sample (K4w: channel tiles 2 i, 2 i + 1; K4x: pixel tiles); an odd
    // number of channel tiles leaves a dead CTA in the last pair of every sample (its W rows are zero-filled, its tables say weight 0)
    const int per_sample = DX ? p.PT : p.MT;
    const int ips = PAIR ? (per_sample + 1) / 2 : per_sample;           // items per sample
    const int items = p.B * ips;
    const int tiles = DX ? p.CB : p.NT;                 // GEMM1 tiles per item
    const int HW = p.H * p.W;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, 1); }
        for (int kb = 0; kb < MAXKB; ++kb) { mbar_init(a_full + kb, 1); mbar_init(a_empty + kb, 1); }
        mbar_init(d1_full, 1); mbar_init(d1_empty, NCTA * EPI_WARPS);
        for (int u = 0; u < NSTG; ++u) { mbar_init(s_full + u, NCTA * EPI_WARPS); mbar_init(s_empty + u, 1); }
        mbar_init(d2_full, 1); mbar_init(d2_empty, NCTA * EPI_WARPS);
        mbar_fence_init();
    }
    if (warp == 2) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) pair_sync();      // the peer's barriers exist and its TMEM is allocated before anything is signalled across the pair
    tc_fence_after();
    // item -> sample and this CTA's 128-row index within the sample (K4w: channel tile, may be the dead one; K4x: pixel tile)
    auto item_b = [&](int item) { return item / ips; };
    auto item_row = [&](int item) { const int i = item - (item / ips) * ips; return PAIR ? 2 * i + (int)rank : i; };
    // barrier of the leader CTA as seen from this one (this CTA's own when there is no pair)
    auto lead = [&](uint64_t* bar) -> uint32_t { return PAIR ? pair_addr(bar, 0) : smem_u32(bar); };
    auto arrive_lead = [&](uint64_t* bar) {        // epilogue -> issuer
        if (PAIR) mbar_arrive_cluster(pair_addr(bar, 0));
        else mbar_arrive(bar);
    };
    auto load_2d = [&](void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {      // bytes are counted on the leader's barrier
        if (PAIR) tma_load_2d_pair(dst, map, c0, c1, pair_addr(bar, 0));
        else tma_load_2d(dst, map, c0, c1, bar);
    };
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    // live 64-column blocks of tile t (K4x: the last channel block of J*D may be partial; its missing W rows are zero-filled by TMA)
    auto live_blocks = [&](int t) -> int {
        if (!DX) return NBLK;
        const int left = p.M - t * BN;
        return left >= BN ? NBLK : (left + BK - 1) / BK;
    };

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            uint32_t it = 0, n_item = 0;
            const int nsub = PAIR ? p.KB / 2 : p.KB;        // n-blocks of a GEMM2 operand block this CTA loads (the pair splits C_in)
            auto load_b2 = [&](int row0) {       // one GEMM2 block: boxes of [64 contraction rows x 64 n], n-block i at +i * 8 KiB
                const uint32_t s = it % NST, ph = (it / NST) & 1;
                mbar_wait(b_empty + s, ph ^ 1);
#ifdef IHPR_TIMING_EXPERIMENTS
                if (!PAIR && (p.dbg & 1)) { mbar_arrive(b_full + s); ++it; return; }
#endif
                if (rank == 0) mbar_expect_tx(b_full + s, (uint32_t)(p.KB * B2_SUB_BYTES));
                for (int i = 0; i < nsub; ++i) load_2d(sB + s * STB + i * B2_SUB_BYTES, &map_b2, ((int)rank * nsub + i) * BK, row0, b_full + s);
                ++it;
            };
            auto load_b1 = [&](int row0) {       // GEMM1 operand of one tile: KB stages of [256 rows x 64 k]
                for (int kb = 0; kb < p.KB; ++kb, ++it) {
                    const uint32_t s = it % NST, ph = (it / NST) & 1;
                    mbar_wait(b_empty + s, ph ^ 1);
#ifdef IHPR_TIMING_EXPERIMENTS
                    if (!PAIR && (p.dbg & 4)) { mbar_arrive(b_full + s); continue; }
#endif
                    if (rank == 0) mbar_expect_tx(b_full + s, (uint32_t)STAGE_BYTES);
                    load_2d(sB + s * STB, &map_b1, kb * BK, row0 + (PAIR ? (int)rank * (BN / 2) : 0), b_full + s);     // the pair splits the 256 N rows
                }
            };
            for (int item = unit; item < items; item += nunits, ++n_item) {
                // K4x: W rows of channel block t are t * 256; K4w: X rows of pixel tile t are b * HW + t * 256
                const int row0 = DX ? 0 : item_b(item) * HW;
                // the issuer's order: GEMM1(0); per tile t: GEMM1(t+1), then GEMM2 blocks 0..3 of t
                load_b1(row0);
                for (int t = 0; t < tiles; ++t) {
                    if (t + 1 < tiles) load_b1(row0 + (t + 1) * BN);
                    const int nb = live_blocks(t);
                    for (int j = 0; j < nb; ++j) load_b2(row0 + t * BN + j * BK);
                }
            }
        }
    } else if (warp == 3) {
        // ================= A1 producer: the stationary operand, one k-block at a time =================
        if (lane == 0) {
            uint32_t n_item = 0;
            for (int item = unit; item < items; item += nunits, ++n_item) {
                // K4x: this CTA's 128 pixels of X; K4w: its 128 channels of W (all out of bounds -> zero-filled for the dead CTA of a pair)
                const int a_row = DX ? item_b(item) * HW + item_row(item) * BM : item_row(item) * BM;
                for (int kb = 0; kb < p.KB; ++kb) {
                    mbar_wait(a_empty + kb, (n_item & 1) ^ 1);
                    if (rank == 0) mbar_expect_tx(a_full + kb, NCTA * (uint32_t)A_KB_BYTES);
                    load_2d(sA + kb * A_KB_BYTES, &map_a, kb * BK, a_row, a_full + kb);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0 && rank == 0) {
            uint32_t it = 0, n_item = 0, tile_it = 0;
            const uint32_t idesc1 = PAIR ? kIdesc1Pair : kIdesc1;
            const uint32_t idesc2 = idesc_bmn(p.K) + (PAIR ? ((uint32_t)(BM >> 4) << 24) : 0u);        // M = 256 for the pair
            auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
                if (PAIR) umma_pair(d, a, b, idesc, acc);
                else umma(d, a, b, idesc, acc);
            };
            auto commit = [&](uint64_t* bar) {      // arrives on `bar` (in both CTAs of a pair) when every MMA issued so far has completed
                if (PAIR) tc_commit_pair(bar);
                else tc_commit(bar);
            };
            const uint32_t tmem_d1 = tmem_base, tmem_d2 = tmem_base + D2_COL;
            uint32_t acc2 = 0;
            // GEMM2 on staging blocks [j0, j1) of a tile: D2 += dH[:, 64 j .. 64 j + 63] . B2 block
            auto gemm2 = [&](int j0, int j1) {
                for (int j = j0; j < j1; ++j, ++it) {
                    const uint32_t s = it % NST, ph = (it / NST) & 1;
                    mbar_wait(b_full + s, ph);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(sS + j * STG_BLK_BYTES), b_addr = smem_u32(sB + s * STB);
#ifdef IHPR_TIMING_EXPERIMENTS
                    if (!PAIR && (p.dbg & 8)) { tc_commit(b_empty + s); continue; }
#endif
#pragma unroll
                    for (int k16 = 0; k16 < BK / 16; ++k16) {
                        // A: K-major, +32 B per 16 contraction columns; B: MN-major, 16 contraction rows = 2 groups of 8 = +2 KiB
                        mma(tmem_d2, umma_desc(a_addr) + 2 * k16, umma_desc_mn(b_addr + k16 * 2048, B2_SUB_BYTES, 1024), idesc2, acc2);
                        acc2 = 1;
                    }
                    commit(b_empty + s);
                }
            };
            // GEMM1 of one tile; first = the item's first tile (A1 k-blocks may still be landing), last = its last (hand A1 back per k-block)
            auto gemm1 = [&](bool first, bool last, uint32_t item_parity) {
                K4_STAMP(tile_it, 0);
                mbar_wait(d1_empty, (tile_it & 1) ^ 1);         // the epilogue has pulled the previous tile out of D1
                tc_fence_after();
                K4_STAMP(tile_it, 1);
                for (int kb = 0; kb < p.KB; ++kb, ++it) {
                    const uint32_t s = it % NST, ph = (it / NST) & 1;
                    if (first) mbar_wait(a_full + kb, item_parity);
                    mbar_wait(b_full + s, ph);
                    tc_fence_after();
                    const uint64_t ad = umma_desc(smem_u32(sA + kb * A_KB_BYTES)), bd = umma_desc(smem_u32(sB + s * STB));
#pragma unroll
                    for (int k16 = 0; k16 < BK / 16; ++k16) mma(tmem_d1, ad + 2 * k16, bd + 2 * k16, idesc1, (uint32_t)((kb | k16) != 0));
                    commit(b_empty + s);
                    if (last) commit(a_empty + kb);
                }
                commit(d1_full);
                K4_STAMP(tile_it, 2);
                ++tile_it;
            };
            // GEMM2 block q of the tile whose epilogue phase parity is tp, nb live blocks
            auto gemm2_block = [&](int q, int nb, uint32_t tp) {
                mbar_wait(s_full + q, tp);
                tc_fence_after();
                gemm2(q, min(q + 1, nb));
                commit(s_empty + q);
            };
            for (int item = unit; item < items; item += nunits, ++n_item) {
                acc2 = 0;
                gemm1(true, tiles == 1, n_item & 1);
                for (int t = 0; t < tiles; ++t) {
                    const uint32_t tp = (tile_it - 1) & 1;                 // tile t is GEMM1 number tile_it - 1 of this CTA
                    // GEMM1 of the NEXT tile first: D1 is free as soon as the epilogue has pulled tile t into registers (a few hundred
                    // clocks), while the four staging blocks of tile t only fill up over the whole of its epilogue
                    if (t + 1 < tiles) gemm1(false, t + 2 == tiles, n_item & 1);
                    if (t == 0) {
                        mbar_wait(d2_empty, (n_item & 1) ^ 1);         // the previous item's D2 has been drained
                        tc_fence_after();
                    }
                    K4_STAMP(tile_it - (t + 1 < tiles ? 2 : 1), 3);
                    const int nb = live_blocks(t);
                    for (int q = 0; q < NBLK; ++q) gemm2_block(q, nb, tp);
                    K4_STAMP(tile_it - (t + 1 < tiles ? 2 : 1), 4);
                }
                commit(d2_full);                                // D2 complete: every MMA of the item has finished
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue =================
        const int e = warp - 4;
        const int qd = warp & 3;                    // TMEM lane quarter this warp may read
        const int cg = e >> 2;                      // 32-column group within a 128-column half
        const int row = qd * 32 + lane;             // row of the tile = TMEM lane
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        const FastDiv divW = make_fastdiv((uint32_t)p.W);
        const int sw = lane & 7;                    // SWIZZLE_128B: 16-byte chunk index ^= row & 7
        const int dshift = 31 - __clz(p.D);         // D is a power of two (32 / 64 / 128)
        const uint64_t l2e2 = pk2(kLog2e, kLog2e);
        uint32_t tile_it = 0, n_item = 0;
        for (int item = unit; item < items; item += nunits, ++n_item) {
            const int b = item_b(item), irow = item_row(item);
            // per-row constants
            float k0_row = 0.f, gx_row = 0.f, gy_row = 0.f, tz_row = 0.f;       // K4w: this channel
            float xf = 0.f, yf = 0.f;                                           // K4x: this pixel
            int c_row = 0;
            if (DX) {
                const uint32_t pix = (uint32_t)(irow * BM + row);
                const uint32_t y = fdiv(pix, divW);
                yf = u2f(y);
                xf = u2f(pix - y * divW.d);
            } else {
                c_row = irow * BM + row;                            // < Mpad: the tables are padded (dead CTA of a pair included)
                k0_row = __ldg(p.k0tab + (size_t)b * p.Mpad + c_row);
                const float4 jc = __ldg(p.jtab + (size_t)b * p.Jpad + (c_row >> dshift));
                gx_row = jc.x; gy_row = jc.y;
                tz_row = fmaf(jc.z, (float)(c_row & (p.D - 1)), jc.w);
            }
            float dsum = 0.f;
            for (int t = 0; t < tiles; ++t, ++tile_it) {
                const uint32_t tp = tile_it & 1;
                if (e == 0 && lane == 0) K4_STAMP(tile_it, 8);
                mbar_wait(d1_full, tp);
                tc_fence_after();
                if (e == 0 && lane == 0) K4_STAMP(tile_it, 9);
                // four steps of 64 columns (= staging block q = one GEMM2 k-block); this warp's 16 columns of step q are
                // q * 64 + cg * 16 ..  Three register buffers: steps 0, 1, 2 are fetched up front, step 3 reuses buffer 0 as soon
                // as step 0 has been computed -- D1 is released to the next GEMM1 before step 1 starts.
                const uint32_t tbase = tmem_base + lane_off + (uint32_t)(cg * 16);
                uint32_t r0[16], r1[16], r2[16];
                tmem_ld16_issue(tbase, r0);
                tmem_ld16_wait(r0);
                if (e == 0 && lane == 0) K4_STAMP(tile_it, 10);
                tmem_ld16_issue(tbase + 64, r1);
                tmem_ld16_issue(tbase + 128, r2);
#pragma unroll
                for (int q = 0; q < NBLK; ++q) {
                    uint32_t(&cur)[16] = (q == 1) ? r1 : ((q == 2) ? r2 : r0);
                    const int col0 = q * 64 + cg * 16;              // first of this warp's 16 columns in this step
                    uint32_t o[8];
                    uint64_t b01, g22;          // (base, base + g) and (2 g, 2 g): the weight of column i is base + i * g
                    uint64_t kk = 0;            // K4w: the row's exponent offset for every column
                    const uint8_t* kp = nullptr;    // K4x: per-column exponent offsets of this run, 4 columns per 16-byte load
                    if (DX) {
                        const int c0 = t * BN + col0;                       // first channel of the run (one joint: D % 32 == 0)
                        const float4 jc = __ldg(p.jtab + (size_t)b * p.Jpad + (c0 >> dshift));
                        const float qq = fmaf(jc.x, xf, fmaf(jc.y, yf, jc.w));
                        const float z0 = (float)(c0 & (p.D - 1));
                        b01 = pk2(fmaf(jc.z, z0, qq), fmaf(jc.z, z0 + 1.f, qq));
                        g22 = pk2(2.f * jc.z, 2.f * jc.z);
                        kp = reinterpret_cast<const uint8_t*>(p.k0tab + (size_t)b * p.Mpad + c0);
                    } else {
                        const uint32_t pix = (uint32_t)(t * BN + col0);     // first pixel of the run (one image row: W % 32 == 0)
                        const uint32_t y = fdiv(pix, divW);
                        const float base = fmaf(gy_row, u2f(y), fmaf(gx_row, u2f(pix - y * divW.d), tz_row));
                        b01 = pk2(base, base + gx_row);
                        g22 = pk2(2.f * gx_row, 2.f * gx_row);
                        kk = pk2(k0_row, k0_row);
                    }
                    uint64_t ds2 = pk2(0.f, 0.f);
#ifdef IHPR_TIMING_EXPERIMENTS
                    if (p.dbg & 2) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) o[i] = cur[2 * i] ^ cur[2 * i + 1];
                    } else
#endif
#pragma unroll
                    for (int i2 = 0; i2 < 4; ++i2) {
                        uint64_t ka = kk, kb2 = kk;
                        if (DX) {
                            const uint4 u = ldg_u4(kp + i2 * 16);
                            ka = pk2u(u.x, u.y);
                            kb2 = pk2u(u.z, u.w);
                        }
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            const int i = 2 * i2 + half;
                            float t0, t1, d0, d1;
                            up2(ffma2(pk2(__uint_as_float(cur[2 * i]), __uint_as_float(cur[2 * i + 1])), l2e2, half ? kb2 : ka), t0, t1);
                            const uint64_t dd = fmul2(pk2(ex2(t0), ex2(t1)), ffma2(pk2((float)i, (float)i), g22, b01));
                            if (!DX) ds2 = fadd2(ds2, dd);
                            up2(dd, d0, d1);
                            o[i] = Elem<__nv_bfloat16>::pk(d0, d1);
                        }
                    }
                    if (!DX) {
                        float da, db;
                        up2(ds2, da, db);
                        dsum += da + db;
                    }
                    if (q == 0) {
                        tmem_ld16_issue(tbase + 192, r0);       // step 3's columns into the buffer step 0 has just finished with
                        if (e == 0 && lane == 0) K4_STAMP(tile_it, 11);
                        // the whole of D1 is in registers now: release it, GEMM1 of the next tile runs underneath the rest of this epilogue
                        tmem_ld16_wait3(r1, r2, r0);
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) arrive_lead(d1_empty);
                        if (e == 0 && lane == 0) K4_STAMP(tile_it, 13);
                    }
                    // GEMM2 of the tile before has read this staging block
                    if (lane == 0) mbar_wait(s_empty + q, tp ^ 1);
                    __syncwarp();
                    // 16 columns = 32 B of this row: chunks 2 cg, 2 cg + 1 of the 128-byte row of staging block q
                    uint8_t* srow = sS + q * STG_BLK_BYTES + row * 128;
                    sts16(srow + (((2 * cg) ^ sw) << 4), make_uint4(o[0], o[1], o[2], o[3]));
                    sts16(srow + (((2 * cg + 1) ^ sw) << 4), make_uint4(o[4], o[5], o[6], o[7]));
                    fence_async_smem();             // generic-proxy stores -> visible to the tensor core (async proxy)
                    __syncwarp();
                    if (lane == 0) arrive_lead(s_full + q);
                    if (q == 0 && e == 0 && lane == 0) K4_STAMP(tile_it, 12);
                    if (q == 3 && e == 0 && lane == 0) K4_STAMP(tile_it, 14);

                }
            }
            // ---- the item's second accumulator
            mbar_wait(d2_full, n_item & 1);
            tc_fence_after();
            if (DX) {
                // dX tile [128 px x K] -> bf16 in registers (D2 is free again at once) -> staging -> TMA store, NHWC.  Every GEMM2 of the
                // item has completed, so the staging blocks are free.
                uint32_t o[32];
                if (cg < p.KB) {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        float v[32];
                        tmem_ld32(tmem_base + lane_off + D2_COL + (uint32_t)(cg * 64 + q * 32), v);
#pragma unroll
                        for (int i = 0; i < 16; ++i) o[q * 16 + i] = Elem<__nv_bfloat16>::pk(v[2 * i], v[2 * i + 1]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive_lead(d2_empty);
                if (cg < p.KB) {
                    uint8_t* srow = sS + cg * STG_BLK_BYTES + row * 128;
#pragma unroll
                    for (int i = 0; i < 8; ++i) sts16(srow + ((i ^ sw) << 4), make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]));
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        const int px_row = b * HW + irow * BM + qd * 32;
                        tma_store_2d(&map_dx, sS + cg * STG_BLK_BYTES + qd * 32 * 128, cg * BK, px_row);
                        tma_store_commit();
                        tma_store_wait_read();      // the staging rows may be rewritten once the store has read them
                    }
                }
                // the next item's epilogue writes other rows / blocks of the staging buffer than this warp just stored from
                named_bar_sync(1, EPI_WARPS * 32);
            } else {
                // dW partial [128 channels x K] fp32 -> workspace, TRANSPOSED ([K][128 channels]): a thread owns a channel (TMEM lane), so the 32
                // lanes of a warp write 128 contiguous bytes per store.  (Row-major float4 stores put 16 bytes into each of 32 different
                // lines per instruction: 8192 line transactions per drain, ~10 k clk -- profiles/r02_k4_ts_trace.txt.)
                const bool warp_live = irow * BM + qd * 32 < p.M;     // 32 consecutive channels: live or dead together (M % 32 == 0)
                if (cg < p.KB && warp_live) {
                    float* dst = p.dw_part + (((size_t)b * p.MT + irow) * p.K + cg * 64) * BM + row;
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        float v[32];
                        tmem_ld32(tmem_base + lane_off + D2_COL + (uint32_t)(cg * 64 + q * 32), v);
#pragma unroll
                        for (int i = 0; i < 32; ++i) dst[(size_t)(q * 32 + i) * BM] = v[i];
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive_lead(d2_empty);
                if (p.db_part && c_row < p.M) p.db_part[((size_t)b * 4 + cg) * p.Mpad + c_row] = dsum;
            }
        }
        if (DX && lane == 0) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (PAIR) pair_sync();      // neither CTA leaves (or frees its TMEM) while the other may still signal it or read its shared memory
    if (warp == 2) {
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// per-(sample, joint) and per-(sample, channel) constants of the backward epilogues: one tiny launch
__global__ void head_bwd_prep_kernel(const float* __restrict__ bias, const float* __restrict__ coords, const float* __restrict__ stats,
                                     const float* __restrict__ gt, const float* __restrict__ vis, const float* __restrict__ have_depth,
                                     const float* __restrict__ grad_out, float loss_scale, int B, int J, int D, int M, int Mpad, int Jpad,
                                     float* __restrict__ k0tab, float4* __restrict__ jtab) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < B * Mpad) {
        const int b = idx / Mpad, c = idx - b * Mpad;
        float k0 = -INFINITY;
        if (c < M) k0 = __ldg(bias + c) * kLog2e - safe_c(__ldg(stats + 2 * ((size_t)b * J + c / D)));
        k0tab[idx] = k0;
    }
    if (idx < B * Jpad) {
        const int b = idx / Jpad, j = idx - b * Jpad;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < J) {
            const size_t r = (size_t)b * J + j;
            const float il = 1.0f / __ldg(stats + 2 * r + 1);
            const float cx = __ldg(coords + 3 * r), cy = __ldg(coords + 3 * r + 1), cz = __ldg(coords + 3 * r + 2);
            const float sc = __ldg(grad_out) * __ldg(vis + r) * loss_scale * il;
            o.x = sc * sgn(cx - __ldg(gt + 3 * r));
            o.y = sc * sgn(cy - __ldg(gt + 3 * r + 1));
            o.z = sc * sgn(cz - __ldg(gt + 3 * r + 2)) * __ldg(have_depth + b);
            o.w = -(o.x * cx + o.y * cy + o.z * cz);
        }
        jtab[idx] = o;
    }
}

// dW[c, k] = sum_b part[b, tile(c), k, c % 128]  and  dbias[c] = sum_b sum_g db_part[b, g, c]   -- fixed order: bit-reproducible.
// The thread index runs over (tile, k, channel) like the partials, so the reads are coalesced; 4 independent partial sums keep
// enough loads in flight (combined in a fixed order).
__global__ void head_bwd_reduce_kernel(const float* __restrict__ dw_part, const float* __restrict__ db_part, int B, int M, int Mpad, int K, int MT,
                                       float* __restrict__ dw, float* __restrict__ dbias) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (dw && idx < MT * BM * K) {
        const int r = idx % BM, k = (idx / BM) % K, mt = idx / (BM * K);
        const int c = mt * BM + r;
        if (c < M) {
            const size_t stride = (size_t)MT * K * BM;
            const float* src = dw_part + ((size_t)mt * K + k) * BM + r;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            int b = 0;
            for (; b + 4 <= B; b += 4) {
                s0 += __ldg(src + (size_t)b * stride);
                s1 += __ldg(src + (size_t)(b + 1) * stride);
                s2 += __ldg(src + (size_t)(b + 2) * stride);
                s3 += __ldg(src + (size_t)(b + 3) * stride);
            }
            for (; b < B; ++b) s0 += __ldg(src + (size_t)b * stride);
            dw[(size_t)c * K + k] = (s0 + s1) + (s2 + s3);
        }
    }
    if (dbias && idx < M) {
        float s = 0.f;
        for (int b = 0; b < B; ++b)
            for (int g = 0; g < 4; ++g) s += __ldg(db_part + ((size_t)b * 4 + g) * Mpad + idx);
        dbias[idx] = s;
    }
}

}  // namespace k4

// ---- host side --------------------------------------------------------------------------------------------------
size_t head_bwd_workspace_bytes(int B, int K, int J, int D, int H, int W) {
    (void)H; (void)W;
    const size_t M = (size_t)J * D, Mpad = (M + 255) / 256 * 256, Jpad = Mpad / D, MT = (M + 127) / 128;
    size_t n = 0;
    n += (size_t)B * Mpad * sizeof(float);              // k0tab
    n = (n + 255) / 256 * 256;
    n += (size_t)B * Jpad * sizeof(float4);             // jtab
    n = (n + 255) / 256 * 256;
    n += (size_t)B * 4 * Mpad * sizeof(float);          // db_part
    n = (n + 255) / 256 * 256;
    n += (size_t)B * MT * k4::BM * K * sizeof(float);   // dw_part
    return n + 256;
}

// one CTA per SM (persistent), or one CTA PAIR per two SMs: a cluster of 2 running tcgen05.mma.cta_group::2
template <bool DX, bool PAIR>
static const char* launch_k4(const CUtensorMap& ma, const CUtensorMap& mb1, const CUtensorMap& mb2, const CUtensorMap& mdx, const k4::Params& p, int items,
                             int num_sms, cudaStream_t s) {
    using namespace k4;
    auto kern = head_bwd_kernel<DX, PAIR>;
    const size_t smem = SMEM_BYTES + 1024;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return "cudaFuncSetAttribute failed (head_bwd_kernel)";
    cudaLaunchConfig_t cfg = {};
    const int units = PAIR ? num_sms / 2 : num_sms;
    const int g = items < units ? items : units;
    cfg.gridDim = dim3((unsigned)(PAIR ? 2 * g : g));
    cfg.blockDim = dim3(32 * (4 + EPI_WARPS));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = PAIR ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, kern, ma, mb1, mb2, mdx, p) != cudaSuccess) return "head_bwd_kernel launch failed";
    return nullptr;
}

// pairs: 0 = never, 1 = the dX kernel on SM pairs (default when C_in is 128 or 256), 3 = the dW kernel too (experiment: 5 pair-items per
// sample of 9 channel tiles quantise badly over 74 SM pairs)
const char* launch_head_bwd_params(const void* x_nhwc, const void* w, const float* bias, int B, int K, int J, int D, int H, int W, const float* coords,
                                   const float* stats, const float* gt, const float* vis, const float* have_depth, const float* grad_out,
                                   void* dx_nhwc, float* dweight, float* dbias, void* workspace, int num_sms, int pairs, int* launches, cudaStream_t s) {
    using namespace k4;
    if ((K / BK) % 2 != 0 || num_sms < 2) pairs = 0;         // the pair splits C_in in two halves of whole 64-channel blocks
    Params p;
    p.B = B; p.K = K; p.J = J; p.D = D; p.H = H; p.W = W;
    p.M = J * D;
    p.Mpad = (p.M + 255) / 256 * 256;
    p.Jpad = p.Mpad / D;
    p.KB = K / BK;
    p.MT = (p.M + BM - 1) / BM;
    p.NT = H * W / BN;
    p.PT = H * W / BM;
    p.CB = (p.M + BN - 1) / BN;
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    size_t off = 0;
    float* k0tab = reinterpret_cast<float*>(ws + off);
    off = (off + (size_t)B * p.Mpad * sizeof(float) + 255) / 256 * 256;
    float4* jtab = reinterpret_cast<float4*>(ws + off);
    off = (off + (size_t)B * p.Jpad * sizeof(float4) + 255) / 256 * 256;
    float* db_part = reinterpret_cast<float*>(ws + off);
    off = (off + (size_t)B * 4 * p.Mpad * sizeof(float) + 255) / 256 * 256;
    float* dw_part = reinterpret_cast<float*>(ws + off);
    p.k0tab = k0tab; p.jtab = jtab; p.dw_part = dw_part; p.db_part = db_part;
    p.dbg = 0;
    p.trace = nullptr;
#ifdef IHPR_TIMING_EXPERIMENTS
    if (const char* e = getenv("IHPR_K4_DEBUG")) p.dbg = atoi(e);
    static long long* d_trace = nullptr;
    if (getenv("IHPR_K4_TRACE")) {
        if (!d_trace) cudaMalloc(&d_trace, 64 * 16 * sizeof(long long));
        cudaMemsetAsync(d_trace, 0, 64 * 16 * sizeof(long long), s);
        p.trace = d_trace;
    }
#endif

    const float loss_scale = 1.0f / (3.0f * (float)B * (float)J);
    {
        const int n = B * p.Mpad, th = 256;
        head_bwd_prep_kernel<<<(n + th - 1) / th, th, 0, s>>>(bias, coords, stats, gt, vis, have_depth, grad_out, loss_scale, B, J, D, p.M, p.Mpad, p.Jpad,
                                                              k0tab, jtab);
        ++*launches;
    }
    CUtensorMap map_w128, map_w256, map_w64, map_x128, map_x256, map_x64, map_dx;
    memset(&map_dx, 0, sizeof(map_dx));
    const uint64_t xrows = (uint64_t)B * H * W;
    if (!tc::make_map(&map_w128, w, (uint64_t)p.M, (uint64_t)K, 128) || !tc::make_map(&map_w256, w, (uint64_t)p.M, (uint64_t)K, 256) ||
        !tc::make_map(&map_w64, w, (uint64_t)p.M, (uint64_t)K, 64))
        return "cuTensorMapEncodeTiled failed for the weight";
    if (!tc::make_map(&map_x128, x_nhwc, xrows, (uint64_t)K, 128) || !tc::make_map(&map_x256, x_nhwc, xrows, (uint64_t)K, 256) ||
        !tc::make_map(&map_x64, x_nhwc, xrows, (uint64_t)K, 64))
        return "cuTensorMapEncodeTiled failed for the activations";
    if (dx_nhwc && !tc::make_map(&map_dx, dx_nhwc, xrows, (uint64_t)K, 32)) return "cuTensorMapEncodeTiled failed for d loss / d x";
    if (dweight || dbias) {
        const char* err = (pairs & 2) ? launch_k4<false, true>(map_w128, map_x128, map_x64, map_dx, p, B * ((p.MT + 1) / 2), num_sms, s)
                                      : launch_k4<false, false>(map_w128, map_x256, map_x64, map_dx, p, B * p.MT, num_sms, s);
        if (err) return err;
        ++*launches;
#ifdef IHPR_TIMING_EXPERIMENTS
        if (p.trace && getenv("IHPR_K4_TRACE")[0] == '2') {        // dump once: stamps relative to the first, per tile
            static bool dumped = false;
            if (!dumped) {
                dumped = true;
                long long h[64 * 16];
                cudaStreamSynchronize(s);
                cudaMemcpy(h, p.trace, sizeof(h), cudaMemcpyDeviceToHost);
                const long long t0 = h[0];
                fprintf(stderr, "K4w trace of CTA 0 (clk since first stamp): tile | issuer: wait_d1e got_d1e g1_issued wait_s0 got_s0 | epilogue: wait_d1f got_d1f ld0 step0_math staged0 d1_released step3_staged\n");
                for (int t = 0; t < 34; ++t) {
                    fprintf(stderr, "%2d |", t);
                    for (int k = 0; k < 5; ++k) fprintf(stderr, " %7lld", h[t * 16 + k] ? h[t * 16 + k] - t0 : -1);
                    fprintf(stderr, " |");
                    for (int k = 8; k < 15; ++k) fprintf(stderr, " %7lld", h[t * 16 + k] ? h[t * 16 + k] - t0 : -1);
                    fprintf(stderr, "\n");
                }
            }
        }
#endif
        const int n = p.MT * BM * K, th = 256;
        head_bwd_reduce_kernel<<<(n + th - 1) / th, th, 0, s>>>(dw_part, db_part, B, p.M, p.Mpad, K, p.MT, dweight, dbias);
        ++*launches;
    }
    if (dx_nhwc) {
        const char* err = (pairs & 1) ? launch_k4<true, true>(map_x128, map_w128, map_w64, map_dx, p, B * (p.PT / 2), num_sms, s)
                                      : launch_k4<true, false>(map_x128, map_w256, map_w64, map_dx, p, B * p.PT, num_sms, s);
        if (err) return err;
        ++*launches;
    }
    return nullptr;
}

}  // namespace ihpr

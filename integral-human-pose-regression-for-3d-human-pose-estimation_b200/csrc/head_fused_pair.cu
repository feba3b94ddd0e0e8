// head_fused_pair.cu -- K3 / K4 on SM pairs (tcgen05 cta_group::2): the fused 1x1 conv + soft-argmax forward and the heat-map
// gradient of head_fused_fwd.cu with HALF the activation traffic per SM.
//
// Same math, same epilogues, same outputs as head_fused_fwd.cu (HeadNet.final_layer, /root/reference/main/model.py:14-20,42, fused with
// soft_argmax / JointLocationLoss and its backward, /root/reference/common/nets/loss.py:13-52).  What changes is who feeds the
// tensor cores: profiles/r01_ncu_K4.txt shows both kernels pay for re-reading the activations X[b] from L2 once per 128-channel
// tile.  Here the two CTAs of a cluster own two ADJACENT channel tiles (mt = 2*mtp + rank) of the same sample and run ONE
// 256 x 256 x 16 UMMA per k-step: each CTA loads its own Wt tile (A, 128 rows) and only HALF of the X tile (B: 128 of the 256
// pixels, 16 KiB per k-block instead of 32), the tensor cores read both halves across the pair, and each CTA ends up with its
// own 128 channels x 256 pixels accumulator in its own TMEM -- so the epilogues are untouched.  The halved stages buy a deeper
// ring (8 / 6 k-blocks in flight instead of 4 / 3).
//   leader (rank 0): issues every MMA; its b_full / a_full barriers count the bytes of BOTH CTAs' TMA loads
//                    (cp.async.bulk.tensor ... .cta_group::2 may complete on the peer's barrier); its t_empty barriers collect
//                    the epilogue warps of both CTAs (the peer's arrive remotely)
//   both           : tcgen05.commit ... .multicast::cluster frees the X stage / Wt tile and publishes the accumulator stage in
//                    both CTAs at once
// With an odd number of channel tiles (J*D = 1152 -> 9) the last pair has a dead CTA whose Wt tile is all zero-fill.
#include "head_tc.cuh"

namespace ihpr {

namespace k3p {
using namespace tc;

constexpr int BM = 128;             // channels per CTA = its half of the UMMA's M = 256
constexpr int BN = 256;             // pixels per tile = UMMA N
constexpr int BNH = BN / 2;         // pixels of the X tile this CTA loads
constexpr int BK = 64;
constexpr int STAGES_FWD = 8;
constexpr int STAGES_BWD = 6;
constexpr int STG_ROW = 128;
constexpr int STG_WARP = 32 * STG_ROW;
constexpr int MAXKB = 4;
constexpr int A_KB_BYTES = BM * BK * 2;     // 16 KiB
constexpr int B_KB_BYTES = BNH * BK * 2;    // 16 KiB: this CTA's half of a k-block
constexpr int EPI_FWD = 16;
constexpr int EPI_BWD = 16;
constexpr int EPI_MAX = 16;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t kIdesc = make_idesc(2 * BM, BN);

struct Params {
    int B, K, J, D, H, W;
    int MT;                 // channel tiles = ceil(J*D / 128)
    int MTP;                // channel-tile pairs = ceil(MT / 2)
    int NT;                 // pixel tiles = H*W / 256
    int KB;                 // k-blocks = K / 64
    const float* bias;
    float* coords;
    float* stats;
    const float* gt;
    const float* vis;
    const float* have_depth;
    const float* grad_out;
    float loss_scale;
    __nv_bfloat16* grad_heat;
    float* dbias_part;
};

template <bool BWD>
__global__ void __launch_bounds__(32 * (4 + (BWD ? EPI_BWD : EPI_FWD)), 1)
head_softargmax_pair_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x,
                       const __grid_constant__ CUtensorMap map_g, const Params p) {
    constexpr int STAGES = BWD ? STAGES_BWD : STAGES_FWD;
    constexpr int EPI_WARPS = BWD ? EPI_BWD : EPI_FWD;
    constexpr int CS = EPI_WARPS / 4;           // column splits of the 256-column accumulator stage
    constexpr int CW = BN / CS;                 // columns per epilogue warp
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                                     // [KB][128 x 64] bf16
    uint8_t* sB = smem + MAXKB * A_KB_BYTES;                // [STAGES][256 x 64] bf16
    uint8_t* stg = sB + STAGES * B_KB_BYTES;                // K4: [EPI_WARPS][32 rows x 128 B] gradient staging, 1024-byte aligned (SWIZZLE_128B)
    uint64_t* bars = reinterpret_cast<uint64_t*>(stg + (BWD ? EPI_WARPS * STG_WARP : 0));
    uint64_t* b_full = bars;                // [STAGES] TMA -> MMA
    uint64_t* b_empty = bars + STAGES;      // [STAGES] MMA -> TMA (tcgen05.commit)
    uint64_t* a_full = bars + 2 * STAGES;   // [1]
    uint64_t* a_empty = a_full + 1;         // [1]      MMA -> TMA: the item's last MMA has read Wt
    uint64_t* t_full = a_empty + 1;         // [2]      MMA -> epilogue: accumulator stage complete
    uint64_t* t_empty = t_full + 2;         // [2]      epilogue -> MMA: accumulator stage drained (EPI_WARPS arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
    float(*red)[8] = reinterpret_cast<float(*)[8]>(tmem_slot + 2);      // [2][EPI_MAX][8]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = pair_rank();           // 0 = leader (issues the MMAs), 1 = peer
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int items = p.B * p.MTP;               // (sample, channel-tile pair)

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, 1); }
        mbar_init(a_full, 1); mbar_init(a_empty, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(t_full + s, 1); mbar_init(t_empty + s, 2 * EPI_WARPS); }
        mbar_fence_init();
    }
    if (warp == 2) {        // TMEM allocation: one warp in each CTA of the pair, same columns in both
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    pair_sync();            // the peer's barriers exist and its TMEM is allocated before anything is signalled across the pair
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    if (warp == 0) {
        // ================= TMA producer (both CTAs: own Wt tile, own half of every X k-block; bytes counted by the leader) =================
        if (lane == 0) {
            const uint32_t a_full_l = pair_addr(a_full, 0);
            uint32_t it = 0, n_item = 0;
            for (int item = pair; item < items; item += npairs, ++n_item) {
                const int b = item / p.MTP, mt = 2 * (item - b * p.MTP) + (int)rank;
                mbar_wait(a_empty, (n_item & 1) ^ 1);
                if (rank == 0) mbar_expect_tx(a_full, (uint32_t)(2 * p.KB * A_KB_BYTES));
                for (int kb = 0; kb < p.KB; ++kb)       // a dead tile (mt == MT) is all out of bounds: zero-filled
                    tma_load_2d_pair(sA + kb * A_KB_BYTES, &map_w, kb * BK, mt * BM, a_full_l);
                for (int nt = 0; nt < p.NT; ++nt)
                    for (int kb = 0; kb < p.KB; ++kb, ++it) {
                        const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                        mbar_wait(b_empty + s, ph ^ 1);
                        if (rank == 0) mbar_expect_tx(b_full + s, (uint32_t)(2 * B_KB_BYTES));
                        tma_load_2d_pair(sB + s * B_KB_BYTES, &map_x, kb * BK, (b * p.NT + nt) * BN + (int)rank * BNH, pair_addr(b_full + s, 0));
                    }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader only) =================
        if (lane == 0 && rank == 0) {
            uint32_t it = 0, n_item = 0, acc_it = 0;
            for (int item = pair; item < items; item += npairs, ++n_item) {
                mbar_wait(a_full, n_item & 1);
                for (int nt = 0; nt < p.NT; ++nt, ++acc_it) {
                    const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
                    mbar_wait(t_empty + as, aph ^ 1);       // the epilogue warps of BOTH CTAs have drained this stage
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + as * BN;
                    for (int kb = 0; kb < p.KB; ++kb, ++it) {
                        const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                        mbar_wait(b_full + s, ph);
                        tc_fence_after();
                        const uint64_t ad = umma_desc(smem_u32(sA + kb * A_KB_BYTES)), bd = umma_desc(smem_u32(sB + s * B_KB_BYTES));
#pragma unroll
                        for (int k16 = 0; k16 < BK / 16; ++k16)
                            umma_pair(tmem_d, ad + 2 * k16, bd + 2 * k16, kIdesc, (uint32_t)((kb | k16) != 0));
                        tc_commit_pair(b_empty + s);         // frees the X stage in both CTAs
                    }
                    tc_commit_pair(t_full + as);             // accumulator stage complete in both CTAs
                }
                // Wt tiles may be overwritten; not after the last item: nobody waits for it and the peer may already be gone
                if (item + npairs < items) tc_commit_pair(a_empty);
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue =================
        const int e = warp - 4;
        const int qd = warp & 3;                 // TMEM lane quarter this warp may read
        const int ch = e >> 2;                   // column split
        const int lane_c = qd * 32 + lane;       // channel within the tile = TMEM lane
        const FastDiv divW = make_fastdiv((uint32_t)p.W);
        uint32_t acc_it = 0, n_item = 0;
        const uint32_t t_empty_l0 = pair_addr(t_empty, 0), t_empty_l1 = pair_addr(t_empty + 1, 0);
        auto release_stage = [&](uint32_t as) {     // this warp is done with accumulator stage `as`: tell the leader's MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(as ? t_empty_l1 : t_empty_l0);
        };
        for (int item = pair; item < items; item += npairs, ++n_item) {
            const int b = item / p.MTP, mt = 2 * (item - b * p.MTP) + (int)rank;
            const int c = mt * BM + lane_c;                      // output channel
            const bool valid = c < p.J * p.D;
            const float bias_f = valid ? __ldg(p.bias + c) : 0.f;
            const float bias2 = bias_f * kLog2e;
            const float zf = (float)(c % p.D);
            if constexpr (BWD) {
                // K4: d loss / d heat = p * sum_c g_c (c(i) - coord_c), p recomputed from the TMEM tile + stats (m, l)
                const int joint = min(c / p.D, p.J - 1);
                const size_t r = (size_t)b * p.J + joint;
                const float m = __ldg(p.stats + 2 * r), il = 1.0f / __ldg(p.stats + 2 * r + 1);
                const float cx = __ldg(p.coords + 3 * r), cy = __ldg(p.coords + 3 * r + 1), cz = __ldg(p.coords + 3 * r + 2);
                const float sc = __ldg(p.grad_out) * __ldg(p.vis + r) * p.loss_scale * il;
                const float gx = sc * sgn(cx - __ldg(p.gt + 3 * r)), gy = sc * sgn(cy - __ldg(p.gt + 3 * r + 1));
                const float gz = sc * sgn(cz - __ldg(p.gt + 3 * r + 2)) * __ldg(p.have_depth + b);
                const float k0 = bias2 - safe_c(m);
                const float tz = gz * (zf - cz);
                static_assert(!BWD || CW == 64, "K4: one epilogue warp owns 64 pixels = one 128-byte staging row per channel");
                uint8_t* wstg = stg + e * STG_WARP;
                uint8_t* srow = wstg + lane * STG_ROW;
                const int sw = lane & 7;                     // SWIZZLE_128B: 16-byte chunk index ^= row & 7
                float dsum = 0.f;
                for (int nt = 0; nt < p.NT; ++nt, ++acc_it) {
                    const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
                    mbar_wait(t_full + as, aph);
                    tc_fence_after();
                    if (mt >= p.MT) {               // dead CTA of the last pair: nothing to compute or store
                        release_stage(as);
                        continue;
                    }
                    // the TMA store of the previous tile has read this warp's staging buffer
                    if (lane == 0) tma_store_wait_read();
                    __syncwarp();
                    const uint32_t tbase = tmem_base + ((uint32_t)(qd * 32) << 16) + as * BN + ch * CW;
#pragma unroll 1
                    for (int j = 0; j < CW / 32; ++j) {
                        float v[32];
                        tmem_ld32(tbase + j * 32, v);
                        if (j == CW / 32 - 1) release_stage(as);     // the accumulator stage is in registers: hand it back to the MMA warp
                        const uint32_t pix = (uint32_t)(nt * BN + ch * CW + j * 32);
                        const uint32_t y = fdiv(pix, divW);
                        const float base = fmaf(gy, u2f(y) - cy, fmaf(gx, u2f(pix - y * divW.d) - cx, tz));
                        uint32_t o[16];
                        const uint64_t l2e2 = pk2(kLog2e, kLog2e), k02 = pk2(k0, k0);
                        const uint64_t b01 = pk2(base, base + gx), gx22 = pk2(2.f * gx, 2.f * gx);
                        uint64_t ds2 = pk2(0.f, 0.f);
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            float t0, t1, d0, d1;
                            up2(ffma2(pk2(v[2 * i], v[2 * i + 1]), l2e2, k02), t0, t1);
                            // (base + 2i gx, base + (2i+1) gx) = (base, base + gx) + i (2gx, 2gx)
                            const uint64_t dd = fmul2(pk2(ex2(t0), ex2(t1)), ffma2(pk2((float)i, (float)i), gx22, b01));
                            ds2 = fadd2(ds2, dd);
                            up2(dd, d0, d1);
                            o[i] = Elem<__nv_bfloat16>::pk(d0, d1);
                        }
                        {
                            float da, db;
                            up2(ds2, da, db);
                            dsum += da + db;
                        }
                        // this thread's 32 pixels (64 B) of its channel row: four 16-byte chunks of the row's swizzled 128 bytes
                        // (conflict-free: the 8 lanes of a store phase hit 8 different chunk positions)
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            sts16(srow + (((j * 4 + i) ^ sw) << 4), make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]));
                    }
                    // 32 channels x 64 pixels leave as ONE TMA tile store: full 128-byte lines per channel row, rows beyond J*D
                    // clipped by the tensor map (the first version stored 16-byte pieces at an 8 KiB stride from registers: 154 us
                    // at B = 32; staging + 64-byte STG segments: 103 us)
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_3d(&map_g, wstg, nt * BN + ch * CW, mt * BM + qd * 32, b);
                        tma_store_commit();
                    }
                }
                if (p.dbias_part && valid) p.dbias_part[((size_t)b * CS + ch) * (p.J * p.D) + c] = dsum;
                continue;
            }
            // weights are 2^(acc*log2e + bias2 - cref); (m, c) of Acc hold the reference point of (acc + bias)
            Acc a;
            a.reset();
            for (int nt = 0; nt < p.NT; ++nt, ++acc_it) {
                const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
                mbar_wait(t_full + as, aph);
                tc_fence_after();
                const uint32_t tbase = tmem_base + ((uint32_t)(qd * 32) << 16) + as * BN + ch * CW;
                // 16 columns at a time, software-pipelined: the tcgen05.ld of chunk q+1 is in flight while chunk q is reduced, so the
                // TMEM load latency and the max chain of one chunk hide behind the exponentials of another (the 4 epilogue warps of
                // an SM sub-partition start every tile in lockstep; with 32-column loads the MUFU idled during every load)
                uint32_t ra[16], rb[16];
                tmem_ld16_issue(tbase, ra);
                tmem_ld16_wait(ra);
#pragma unroll
                for (int q = 0; q < CW / 16; ++q) {
                    uint32_t(&cur)[16] = (q & 1) ? rb : ra;
                    uint32_t(&nxt)[16] = (q & 1) ? ra : rb;
                    if (q + 1 < CW / 16) tmem_ld16_issue(tbase + (q + 1) * 16, nxt);
                    const uint32_t pix = (uint32_t)(nt * BN + ch * CW + q * 16);     // first pixel of this 16-column chunk (W % 32 == 0: one image row)
                    const uint32_t y = fdiv(pix, divW);
                    const float yf = u2f(y), x0f = u2f(pix - y * divW.d);
                    float cmax = __uint_as_float(cur[0]);
#pragma unroll
                    for (int i = 1; i < 16; ++i) cmax = fmaxf(cmax, __uint_as_float(cur[i]));
                    // reference point in "h" units: h = acc + bias
                    const float hmax = cmax + bias_f;
                    a.mx = fmaxf(a.mx, hmax);
                    if (hmax > a.lim) acc_raise(a, hmax);
                    const float k0 = bias2 - a.c;
                    // (even, odd) column pairs: s2 = (sum p_even, sum p_odd), w2 = (sum i p_2i, sum i p_2i+1)
                    const uint64_t l2e2 = pk2(kLog2e, kLog2e), k02 = pk2(k0, k0);
                    uint64_t s2a = pk2(0.f, 0.f), s2b = s2a, w2a = s2a, w2b = s2a;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float t0, t1;
                        up2(ffma2(pk2(__uint_as_float(cur[2 * i]), __uint_as_float(cur[2 * i + 1])), l2e2, k02), t0, t1);
                        const uint64_t pp = pk2(ex2(t0), ex2(t1));
                        if (i & 1) { s2b = fadd2(s2b, pp); w2b = ffma2(pp, pk2((float)i, (float)i), w2b); }
                        else { s2a = fadd2(s2a, pp); w2a = ffma2(pp, pk2((float)i, (float)i), w2a); }
                    }
                    float se, so, we, wo;
                    up2(fadd2(s2a, s2b), se, so);
                    up2(fadd2(w2a, w2b), we, wo);
                    const float s = se + so;
                    const float w = fmaf(2.f, we + wo, so);          // sum_j j p_j = 2 (sum i p_2i + sum i p_2i+1) + sum p_2i+1
                    a.l += s;
                    a.sx += fmaf(x0f, s, w);
                    a.sy = fmaf(yf, s, a.sy);
                    if (q + 1 < CW / 16) {
                        tmem_ld16_wait(nxt);
                        if (q + 2 == CW / 16) {         // the whole accumulator stage is in registers: hand it back to the MMA warp
                            release_stage(as);
                        }
                    }
                }
            }
            // ---- merge the lanes (z) and the two column halves of each joint, write coords / stats
            if (!valid) a.reset();
            a.sz = zf * a.l;
            a = acc_warp_merge(a);              // 32 lanes of one warp always belong to one joint (D % 32 == 0)
            float(*rb)[8] = red + (n_item & 1) * EPI_MAX;
            if (lane == 0) partial_to_smem(rb[e], a);
            named_bar_sync(1, EPI_WARPS * 32);
            const int jpt = BM / p.D;           // joints per channel tile
            if (e == 0 && lane < jpt) {
                const int jt = lane;
                Acc t;
                t.reset();
                const int q0 = jt * p.D / 32, q1 = (jt + 1) * p.D / 32;       // lane quarters of this joint
                for (int qq = q0; qq < q1; ++qq)
                    for (int hh = 0; hh < CS; ++hh) t = acc_merge(t, partial_from_smem(rb[hh * 4 + qq]));
                const int joint = mt * jpt + jt;
                if (joint < p.J) {
                    const size_t r = (size_t)b * p.J + joint;
                    const float inv = 1.0f / t.l;
                    p.coords[3 * r + 0] = t.sx * inv;
                    p.coords[3 * r + 1] = t.sy * inv;
                    p.coords[3 * r + 2] = t.sz * inv;
                    if (p.stats) {
                        const float f = (t.m == -INFINITY) ? 0.f : ex2(t.c - safe_c(t.mx));
                        p.stats[2 * r] = t.mx;
                        p.stats[2 * r + 1] = t.l * f;
                    }
                }
            }
            // red[] is double-buffered by item parity; the barrier of the NEXT item orders its reuse
        }
    }

    if (BWD && warp >= 4 && lane == 0) tma_store_wait_all();       // gradient tiles are out before the CTA (and its shared memory) goes away
    tc_fence_before();
    __syncthreads();
    pair_sync();            // both CTAs are done with the pair's MMAs, barriers and TMEM
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
}

}  // namespace k3p

const char* launch_head_fused_pair(const void* x_nhwc, const void* w, const float* bias, int B, int K, int J, int D, int H, int W, float* coords,
                                   float* stats, const float* gt, const float* vis, const float* have_depth, const float* grad_out, void* grad_heat,
                                   float* dbias_part, int num_sms, cudaStream_t s) {
    k3p::Params p;
    p.B = B; p.K = K; p.J = J; p.D = D; p.H = H; p.W = W;
    p.MT = (J * D + k3p::BM - 1) / k3p::BM;
    p.MTP = (p.MT + 1) / 2;
    p.NT = H * W / k3p::BN;
    p.KB = K / k3p::BK;
    p.bias = bias; p.coords = coords; p.stats = stats;
    p.gt = gt; p.vis = vis; p.have_depth = have_depth; p.grad_out = grad_out;
    p.loss_scale = 1.0f / (3.0f * (float)B * (float)J);
    p.grad_heat = static_cast<__nv_bfloat16*>(grad_heat);
    p.dbias_part = dbias_part;
    const bool bwd = grad_heat != nullptr;
    CUtensorMap map_w, map_x, map_g;
    memset(&map_g, 0, sizeof(map_g));
    if (!tc::make_map(&map_w, w, (uint64_t)J * D, (uint64_t)K, k3p::BM)) return "cuTensorMapEncodeTiled failed for the weight";
    if (!tc::make_map(&map_x, x_nhwc, (uint64_t)B * H * W, (uint64_t)K, k3p::BNH)) return "cuTensorMapEncodeTiled failed for the activations";
    if (bwd && !tc::make_map_grad(&map_g, grad_heat, (uint64_t)B, (uint64_t)J * D, (uint64_t)H * W, 64, CU_TENSOR_MAP_SWIZZLE_128B))
        return "cuTensorMapEncodeTiled failed for the gradient";
    const size_t smem = 1024 + k3p::MAXKB * k3p::A_KB_BYTES + (bwd ? k3p::STAGES_BWD : k3p::STAGES_FWD) * k3p::B_KB_BYTES + 32 * sizeof(uint64_t) +
                        2 * k3p::EPI_MAX * 8 * sizeof(float) + (bwd ? k3p::EPI_BWD * k3p::STG_WARP : 0);
    auto kern = bwd ? k3p::head_softargmax_pair_kernel<true> : k3p::head_softargmax_pair_kernel<false>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return "cudaFuncSetAttribute failed";
    int npairs = B * p.MTP;
    if (npairs > num_sms / 2) npairs = num_sms / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * npairs));
    cfg.blockDim = dim3(32 * (4 + (bwd ? k3p::EPI_BWD : k3p::EPI_FWD)));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, kern, map_w, map_x, map_g, p) != cudaSuccess) return "cluster launch of the SM-pair fused head failed";
    return nullptr;
}

}  // namespace ihpr

// tools/experiments/head_bwd_ts.cuh -- NOT BUILT, kept as the record of a measured experiment (round 2).
//
// "TS form" of K4w / K4x: 128-column heat-map tiles, D1 double-buffered in tensor memory, dH written back into tensor memory and used as
// GEMM2's A operand from there (tcgen05.mma with A in TMEM), ONE load of the streamed operand per tile (the K-major image GEMM1 reads is
// the MN-major image GEMM2 reads), two issuing warps, a ring of 16 KiB k-block slots.  It was parity-green on B200 at the first run in all
// three versions tried (the whole test matrix of tests/test_gpu_parity.py::test_fused_head_*), and SLOWER than the shipped form:
//     shipped (256-column tiles, smem staging, every tile loaded twice)   K4w 146.0 us   K4x 143.8 us      (B = 32, tools/head_bench.py)
//     TS, one polling issuer, two 64 KiB tile buffers                      K4w 174.7 us   K4x 176.4 us
//     TS, two issuing warps                                                K4w ~152 us    K4x ~160 us      (timing build: 164 / 171)
//     TS, slot ring + N = 64 GEMM2 MMAs (this file)                        K4w 157 us     K4x 194 us       (ncu, cold)
// Why (profiles/r02_k4_ts_trace.txt, clock64 traces and knock-out builds):
//   * inside this kernel a tcgen05.mma with M = 128, K = 16 cost ~115-137 clk whatever its N: N = 128 (GEMM1 here) 117 clk, N = 64 from a
//     TMEM A operand 115 clk, N = 256 128-137 clk.  (tools/mma_issue_bench.cu later showed why: one thread needs ~58 clk to ISSUE an MMA,
//     more next to busy epilogue warps, so MMAs narrower than N = 256 are issue-bound, not pipe-bound.)  Halving the tile width therefore
//     does not halve GEMM1.  MMA time per 128 columns: 1870 + 1100 clk here against 1024 + 1024 in the shipped form.
//   * only two tiles fit in shared memory next to the stationary operand, and the per-tile chain  slot free -> load (~1700 clk to the
//     first k-block) -> GEMM1 -> epilogue (1600 clk) -> GEMM2 -> slot free  is ~6900 clk for two tiles in flight = 3450 clk per tile.
//   * a lone polling issuer thread needs ~500 clk per poll round next to four busy epilogue warps on its scheduler.
// What carried over into the shipped kernels: the coalesced (transposed) dW partial drain and the matching reduce kernel.
// The helpers it needs beyond csrc/head_tc.cuh (umma_ts, tmem_st16, tmem_st_wait, mbar_test, make_map_box) are at the end of this file.
//
// =====================================================================================================================
// TS form (the default since round 2's last third): the same two-GEMM chain, re-cut around what the traces of the first form showed
// (6200 clk per 256-column tile for 4096 clk of MMA work; the operand ring, which loaded every tile TWICE, delivered ~730 clk per
// 32 KiB stage = 5800 clk per tile):
//   * ONE load of the streamed operand per tile.  A [128 rows x 64] k-block that GEMM1 reads K-major (B1: N = rows) is the very
//     shared-memory image GEMM2 reads MN-major (B2: contraction = rows, N = those 64 input channels), so a k-block stays in its
//     slot from GEMM1 until GEMM2 has read it and the L2 -> SM traffic per tile halves.
//   * 128-column tiles with D1 DOUBLE-buffered in tensor memory (2 x 128 columns next to D2's 256): GEMM1(u+1) never waits for the
//     epilogue of tile u.
//   * dH goes back INTO tensor memory (bf16, over the first 64 columns of the D1 buffer it was computed from) and is GEMM2's A operand
//     from there (tcgen05.mma with A in TMEM): no staging buffer, no generic-proxy shared-memory stores, no proxy fence.
//   * the operand slots are a RING of 16 KiB k-blocks (10 for K4w, 9 + one staging slot for K4x): 2.25-2.5 tiles in flight, every slot
//     released by GEMM2 as soon as ITS 64 input channels are done (GEMM2 runs as N = 64 MMAs per k-block), so the loads of tile u+2
//     are under way before GEMM2(u) has finished and the ~1700 clk load latency hides behind it.
//   * TWO issuing warps, one per GEMM: per tile the dependency chain is  slot free -> load -> GEMM1 -> epilogue -> GEMM2 -> slot free,
//     so any fixed order of the two GEMMs in one thread leaves the tensor pipe idle for the load latency of every tile.  (A single
//     issuer that POLLED both sets of barriers was tried first: a lone thread competing with four busy epilogue warps for issue
//     slots needs ~500 clk per poll round -- profiles/r02_k4_ts_trace.txt.)
//   * coalesced drains: the first form's per-item D2 drain wrote 16-byte pieces at a 512 / 1024-byte stride (32 lines per store
//     instruction, 8-14 k clk per item); K4w now writes its fp32 partial TRANSPOSED (lanes = consecutive addresses), K4x stages
//     32-byte row pieces per warp and stores them by TMA.
constexpr int TN = 128;                                 // columns of a heat-map tile (K4w: pixels, K4x: channels) = rows of a k-block slot
constexpr int TS_KB_BYTES = TN * BK * 2;                // 16 KiB: one slot, [128 rows x 64] bf16, SWIZZLE_128B
constexpr int TS_MAX_SLOTS = 10;
constexpr uint32_t kIdescTs1 = make_idesc(BM, TN);
constexpr size_t TS_SMEM_BYTES = (size_t)MAXKB * A_KB_BYTES + (size_t)TS_MAX_SLOTS * TS_KB_BYTES + 512;

//   map_a: stationary A1, box 64 x 128 rows (K4w: W, K4x: X);  map_b: the streamed tile, box 64 x 128 rows (K4w: X, K4x: W)
//   map_dx: K4x only: d loss / d x, (B*H*W, K) bf16, box 16 channels x 32 rows, no swizzle (one warp's staging piece)
template <bool DX>
__global__ void __launch_bounds__(32 * (4 + EPI_WARPS), 1)
head_bwd_ts_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_dx,
                   const Params p) {
    constexpr int NSLOT = DX ? TS_MAX_SLOTS - 1 : TS_MAX_SLOTS;      // K4x gives the tenth slot to the dX staging
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                                     // [KB][128 x 64] bf16, stationary per item
    uint8_t* sB = sA + MAXKB * A_KB_BYTES;                  // [NSLOT][128 x 64] bf16: the ring
    uint8_t* sS = sB + (TS_MAX_SLOTS - 1) * TS_KB_BYTES;    // K4x: the tenth slot is 16 x 1 KiB of per-warp staging for the dX store
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + TS_MAX_SLOTS * TS_KB_BYTES);
    uint64_t* s_full = bars;                    // [NSLOT] TMA -> MMA: slot landed
    uint64_t* s_empty = s_full + TS_MAX_SLOTS;  // [NSLOT] MMA -> TMA: GEMM2 has read the slot
    uint64_t* a_full = s_empty + TS_MAX_SLOTS;  // [MAXKB]
    uint64_t* a_empty = a_full + MAXKB;         // [MAXKB]
    uint64_t* d1_full = a_empty + MAXKB;        // [2] MMA -> epilogue: heat-map tile complete in D1 buffer
    uint64_t* dh_full = d1_full + 2;            // [2] epilogue -> MMA: dH is in tensor memory (EPI_WARPS arrivals)
    uint64_t* d1_free = dh_full + 2;            // [2] GEMM2 -> GEMM1: the dH that aliased this D1 buffer has been consumed
    uint64_t* d2_full = d1_free + 2;            // [1]
    uint64_t* d2_empty = d2_full + 1;           // [1] (EPI_WARPS arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d2_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per_sample = DX ? p.PT : p.MT;            // 128-row items per sample (K4x: pixel tiles, K4w: channel tiles)
    const int items = p.B * per_sample;
    const int tiles = DX ? p.MT : p.PT;                 // 128-column tiles per item (K4x: channel tiles, K4w: pixel tiles)
    const int HW = p.H * p.W;
    const int n_my = ((int)blockIdx.x < items) ? (items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NSLOT; ++i) { mbar_init(s_full + i, 1); mbar_init(s_empty + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(d1_full + i, 1); mbar_init(dh_full + i, EPI_WARPS); mbar_init(d1_free + i, 1); }
        for (int kb = 0; kb < MAXKB; ++kb) { mbar_init(a_full + kb, 1); mbar_init(a_empty + kb, 1); }
        mbar_init(d2_full, 1); mbar_init(d2_empty, EPI_WARPS);
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

    // the ring position of the producer and of both issuers: slot index and phase, advanced once per k-block in the same order by all three
    struct Ring {
        uint32_t s = 0, ph = 0;
        __device__ __forceinline__ void next() { if (++s == (uint32_t)NSLOT) { s = 0; ph ^= 1; } }
    };

    if (warp == 0) {
        // ================= TMA producer: one tile = KB slots of [128 rows x 64], loaded once, read by both GEMMs =================
        if (lane == 0) {
            Ring r;
            uint32_t u = 0;
            for (int i = 0; i < n_my; ++i) {
                const int item = (int)blockIdx.x + i * (int)gridDim.x;
                const int row0 = DX ? 0 : (item / per_sample) * HW;     // K4x: W rows of channel tile t; K4w: X rows of pixel tile t of the sample
                for (int t = 0; t < tiles; ++t, ++u) {
                    for (int kb = 0; kb < p.KB; ++kb, r.next()) {
                        mbar_wait(s_empty + r.s, r.ph ^ 1);
                        if (kb == 0) K4_STAMP(u, 0);
#ifdef IHPR_TIMING_EXPERIMENTS
                        if (p.dbg & 4) { mbar_arrive(s_full + r.s); continue; }
#endif
                        mbar_expect_tx(s_full + r.s, (uint32_t)TS_KB_BYTES);
                        tma_load_2d(sB + r.s * TS_KB_BYTES, &map_b, kb * BK, row0 + t * TN, s_full + r.s);
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ================= A1 producer: the stationary operand, one k-block at a time =================
        if (lane == 0) {
            for (int i = 0; i < n_my; ++i) {
                const int item = (int)blockIdx.x + i * (int)gridDim.x;
                const int bb = item / per_sample, irow = item - bb * per_sample;
                const int a_row = DX ? bb * HW + irow * BM : irow * BM;
                for (int kb = 0; kb < p.KB; ++kb) {
                    mbar_wait(a_empty + kb, (i & 1) ^ 1);
                    mbar_expect_tx(a_full + kb, (uint32_t)A_KB_BYTES);
                    tma_load_2d(sA + kb * A_KB_BYTES, &map_a, kb * BK, a_row, a_full + kb);
                }
            }
        }
    } else if (warp == 1) {
        // ================= GEMM1 issuer: D1[buf] = A1 . tile^T =================
        if (lane == 0) {
            Ring r;
            uint32_t u = 0;
            for (int i = 0; i < n_my; ++i) {
                for (int t = 0; t < tiles; ++t, ++u) {
                    const uint32_t buf = u & 1;
                    mbar_wait(d1_free + buf, ((u >> 1) & 1) ^ 1);      // GEMM2(u - 2) has consumed the dH that lived in this D1 buffer
                    tc_fence_after();
                    for (int kb = 0; kb < p.KB; ++kb, r.next()) {
                        if (t == 0) mbar_wait(a_full + kb, (uint32_t)i & 1);
                        mbar_wait(s_full + r.s, r.ph);
                        tc_fence_after();
                        const uint64_t ad = umma_desc(smem_u32(sA + kb * A_KB_BYTES));
                        const uint64_t bd = umma_desc(smem_u32(sB + r.s * TS_KB_BYTES));
#ifdef IHPR_TIMING_EXPERIMENTS
                        if (!(p.dbg & 16))
#endif
#pragma unroll
                        for (int k16 = 0; k16 < BK / 16; ++k16) umma(tmem_base + buf * TN, ad + 2 * k16, bd + 2 * k16, kIdescTs1, (uint32_t)((kb | k16) != 0));
                        if (t == tiles - 1) tc_commit(a_empty + kb);        // the item's last GEMM1 has read this A1 k-block
                        if (kb == 0) K4_STAMP(u, 1);
                    }
                    tc_commit(d1_full + buf);
                    K4_STAMP(u, 2);
                }
            }
        }
    } else if (warp == 2) {
        // ================= GEMM2 issuer (its own warp: the two GEMMs wait on different things and must not wait for each other) =================
        // D2[:, 64 kb ..] += dH(u) [TMEM, 128 x 128 bf16] . slot kb read MN-major [128 contraction rows x 64 input channels]
        if (lane == 0) {
            const uint32_t idesc2 = idesc_bmn(BK);
            const uint32_t tmem_d2 = tmem_base + D2_COL;
            Ring r;
            uint32_t u = 0;
            for (int i = 0; i < n_my; ++i) {
                for (int t = 0; t < tiles; ++t, ++u) {
                    const uint32_t buf = u & 1;
                    if (t == 0) mbar_wait(d2_empty, ((uint32_t)i & 1) ^ 1);         // the previous item's D2 has been drained
                    mbar_wait(dh_full + buf, (u >> 1) & 1);
                    tc_fence_after();
                    const uint32_t ta = tmem_base + buf * TN;
                    for (int kb = 0; kb < p.KB; ++kb, r.next()) {
                        mbar_wait(s_full + r.s, r.ph);          // landed long ago (GEMM1 read it): this thread's own acquire of the slot
                        const uint32_t sb = smem_u32(sB + r.s * TS_KB_BYTES);
#ifdef IHPR_TIMING_EXPERIMENTS
                        if (!(p.dbg & 8))
#endif
#pragma unroll
                        for (int k16 = 0; k16 < TN / 16; ++k16)     // A: 8 packed columns per 16 contraction elements; B: 16 rows = +2 KiB
                            umma_ts(tmem_d2 + kb * BK, ta + k16 * 8, umma_desc_mn(sb + k16 * 2048, TS_KB_BYTES, 1024), idesc2, (uint32_t)((t | k16) != 0));
                        tc_commit(s_empty + r.s);               // the slot goes back to the producer as soon as its 64 input channels are done
                    }
                    tc_commit(d1_free + buf);
                    K4_STAMP(u, 3);
                }
                tc_commit(d2_full);                     // D2 complete: every GEMM2 of the item has finished
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue =================
        const int e = warp - 4;
        const int qd = warp & 3;                    // TMEM lane quarter this warp may touch
        const int cg = e >> 2;                      // this warp's 32 columns of the tile: cg * 32 ..
        const int row = qd * 32 + lane;
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        const FastDiv divW = make_fastdiv((uint32_t)p.W);
        const int dshift = 31 - __clz(p.D);
        const uint64_t l2e2 = pk2(kLog2e, kLog2e);
        uint32_t u = 0;
        for (int it = 0; it < n_my; ++it) {
            const int item = (int)blockIdx.x + it * (int)gridDim.x;
            const int b = item / per_sample, irow = item - b * per_sample;
            float k0_row = 0.f, gx_row = 0.f, gy_row = 0.f, tz_row = 0.f;       // K4w: this channel
            float xf = 0.f, yf = 0.f;                                           // K4x: this pixel
            int c_row = 0;
            if (DX) {
                const uint32_t pix = (uint32_t)(irow * BM + row);
                const uint32_t y = fdiv(pix, divW);
                yf = u2f(y);
                xf = u2f(pix - y * divW.d);
            } else {
                c_row = irow * BM + row;
                k0_row = __ldg(p.k0tab + (size_t)b * p.Mpad + c_row);
                const float4 jc = __ldg(p.jtab + (size_t)b * p.Jpad + (c_row >> dshift));
                gx_row = jc.x; gy_row = jc.y;
                tz_row = fmaf(jc.z, (float)(c_row & (p.D - 1)), jc.w);
            }
            float dsum = 0.f;
            for (int t = 0; t < tiles; ++t, ++u) {
                const uint32_t buf = u & 1;
                if (e == 0 && lane == 0) K4_STAMP(u, 8);
                mbar_wait(d1_full + buf, (u >> 1) & 1);
                tc_fence_after();
                if (e == 0 && lane == 0) K4_STAMP(u, 9);
                const uint32_t tb = tmem_base + lane_off + buf * TN;
                uint32_t r0[16], r1[16];
                tmem_ld16_issue(tb + (uint32_t)(cg * 32), r0);
                tmem_ld16_issue(tb + (uint32_t)(cg * 32 + 16), r1);
                tmem_ld16_wait(r0);
                tmem_ld16_wait(r1);
                if (e == 0 && lane == 0) K4_STAMP(u, 10);
                // dH overwrites the first 64 columns of this D1 buffer, which the OTHER warps of this lane quarter read: all four have their
                // columns in registers before any of them stores
                tc_fence_before();
                named_bar_sync(1 + qd, 128);
                tc_fence_after();
                if (e == 0 && lane == 0) K4_STAMP(u, 11);
                uint32_t o[16];
#pragma unroll
                for (int run = 0; run < 2; ++run) {
                    uint32_t(&cur)[16] = run ? r1 : r0;
                    const int col0 = cg * 32 + run * 16;            // first of the 16 columns of this run
                    uint64_t b01, g22;          // (base, base + g) and (2 g, 2 g): the weight of column i is base + i * g
                    uint64_t kk = 0;            // K4w: the row's exponent offset for every column
                    const uint8_t* kp = nullptr;    // K4x: per-column exponent offsets of this run, 4 columns per 16-byte load
                    if (DX) {
                        const int c0 = t * TN + col0;                       // first channel of the run (one joint: D % 16 == 0)
                        const float4 jc = __ldg(p.jtab + (size_t)b * p.Jpad + (c0 >> dshift));
                        const float qq = fmaf(jc.x, xf, fmaf(jc.y, yf, jc.w));
                        const float z0 = (float)(c0 & (p.D - 1));
                        b01 = pk2(fmaf(jc.z, z0, qq), fmaf(jc.z, z0 + 1.f, qq));
                        g22 = pk2(2.f * jc.z, 2.f * jc.z);
                        kp = reinterpret_cast<const uint8_t*>(p.k0tab + (size_t)b * p.Mpad + c0);
                    } else {
                        const uint32_t pix = (uint32_t)(t * TN + col0);     // first pixel of the run (one image row: W % 16 == 0)
                        const uint32_t y = fdiv(pix, divW);
                        const float base = fmaf(gy_row, u2f(y), fmaf(gx_row, u2f(pix - y * divW.d), tz_row));
                        b01 = pk2(base, base + gx_row);
                        g22 = pk2(2.f * gx_row, 2.f * gx_row);
                        kk = pk2(k0_row, k0_row);
                    }
                    uint64_t ds2 = pk2(0.f, 0.f);
#pragma unroll
                    for (int i2 = 0; i2 < 4; ++i2) {
                        uint64_t ka = kk, kb2 = kk;
                        if (DX) {
                            const uint4 uu = ldg_u4(kp + i2 * 16);
                            ka = pk2u(uu.x, uu.y);
                            kb2 = pk2u(uu.z, uu.w);
                        }
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            const int i = 2 * i2 + half;
                            float t0, t1, d0, d1;
                            up2(ffma2(pk2(__uint_as_float(cur[2 * i]), __uint_as_float(cur[2 * i + 1])), l2e2, half ? kb2 : ka), t0, t1);
                            const uint64_t dd = fmul2(pk2(ex2(t0), ex2(t1)), ffma2(pk2((float)i, (float)i), g22, b01));
                            if (!DX) ds2 = fadd2(ds2, dd);
                            up2(dd, d0, d1);
                            o[run * 8 + i] = Elem<__nv_bfloat16>::pk(d0, d1);
                        }
                    }
                    if (!DX) {
                        float da, db;
                        up2(ds2, da, db);
                        dsum += da + db;
                    }
                }
                if (e == 0 && lane == 0) K4_STAMP(u, 12);
                // 32 columns of dH = 16 packed columns of GEMM2's A operand (contraction elements 2 j, 2 j + 1 in packed column j)
                tmem_st16(tb + (uint32_t)(cg * 16), o);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(dh_full + buf);
                if (e == 0 && lane == 0) K4_STAMP(u, 13);
            }
            // ---- the item's second accumulator
            mbar_wait(d2_full, (uint32_t)it & 1);
            tc_fence_after();
            if (e == 0 && lane == 0) K4_STAMP(u - 1, 14);
            if (DX) {
                // dX tile [128 px x K] -> bf16 in registers (D2 is free again at once) -> per-warp staging, 32 rows x 32 bytes at a time -> TMA
                // store into the NHWC rows (a direct store would put 16 bytes into each of 32 different lines per instruction)
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
                if (lane == 0) mbar_arrive(d2_empty);
                if (cg < p.KB) {
                    uint8_t* stg = sS + e * 1024;           // [32 rows][32 bytes], this warp's own
                    const int px_row = b * HW + irow * BM + qd * 32;
#pragma unroll
                    for (int r4 = 0; r4 < 4; ++r4) {
                        sts16(stg + lane * 32, make_uint4(o[r4 * 8], o[r4 * 8 + 1], o[r4 * 8 + 2], o[r4 * 8 + 3]));
                        sts16(stg + lane * 32 + 16, make_uint4(o[r4 * 8 + 4], o[r4 * 8 + 5], o[r4 * 8 + 6], o[r4 * 8 + 7]));
                        fence_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&map_dx, stg, cg * BK + r4 * 16, px_row);
                            tma_store_commit();
                            tma_store_wait_read();      // the staging piece may be rewritten once the store has read it
                        }
                        __syncwarp();
                    }
                }
            } else {
                // dW partial [128 channels x K] fp32 -> workspace, TRANSPOSED ([K][128 channels]): the 32 lanes of a warp write 128 contiguous bytes
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
                if (lane == 0) mbar_arrive(d2_empty);
                if (p.db_part && c_row < p.M) p.db_part[((size_t)b * 4 + cg) * p.Mpad + c_row] = dsum;
            }
            if (e == 0 && lane == 0) K4_STAMP(u - 1, 6);
        }
        if (DX && lane == 0) tma_store_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
}


// ---- helpers (were in csrc/head_tc.cuh while this form was built) ----------------------------------------------------
// TS form: the A operand (M = 128 rows x 16 bf16, two per 32-bit column, K-major) is read from tensor memory instead of shared memory
// (cute::SM100_MMA_F16BF16_TS); B stays a shared-memory descriptor
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// registers -> tensor memory: 16 consecutive 32-bit columns of this thread's TMEM lane (the mirror image of tcgen05.ld 32x32b.x16)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]),
                 "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// non-blocking probe of an mbarrier phase
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}

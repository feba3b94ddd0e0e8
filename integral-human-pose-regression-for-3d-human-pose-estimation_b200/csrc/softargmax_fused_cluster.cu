// softargmax_fused_cluster.cu -- K5c: integral L1 loss forward AND heat-map gradient in ONE launch, the
// joint-volume resident in the shared memory of a thread-block cluster between its two passes (sm_100a).
//
// Replaces the same reference steps as K5 (/root/reference/main/train.py:67-71: JointLocationLoss forward,
// common/nets/loss.py:13-52, then autograd backward).  K5 (softargmax_fused.cu) re-reads every unit from L2 and
// trades partials through global memory between co-resident CTAs of a cooperative launch; here
//   * a cluster of CS CTAs owns one joint-volume at a time; CTA q of the cluster streams chunks [q*C, (q+1)*C) of it
//     ONCE from HBM into its TMA ring (CS * C * chunk = N*s, i.e. 16 CTAs x 4 x 16 KiB for a 64^3 fp32 volume),
//   * pass 1 runs over the chunks as they land and leaves them in the ring,
//   * the CS partials are traded through distributed shared memory (st.shared::cluster + a remote mbarrier arrive:
//     a few hundred cycles, no global memory, no polling, no cooperative launch),
//   * pass 2 reads the SAME ring stages again, writes the gradient and only then hands the stages back to the producer.
// DRAM traffic is exactly read V + write V whatever the L2 does.  The CTA's chunks form one stream; pass 2 trails pass 1 by L >= C chunks
// (C = chunks per slice), so L ring stages hold data between the passes, the remaining STAGES - L are loading, and the L - C chunks of
// slack hide the exchange.  Opt-in (variant 7): on B200 it loses to K5 because clusters of 8 / 16 of these 226 KiB CTAs only fill
// 120 / 112 of the 148 SMs and the exchange needs ~3 us of slack (profiles/r01_k5c_cluster_resident.txt).
#include <atomic>
#include <cstdlib>
#include <type_traits>

#include "ihpr_device.cuh"

namespace ihpr {

namespace {

constexpr int kMaxCluster = 16;

// Roles: warp 0 = TMA producer, warps 1..NCW = consumers, warp NCW+1 = exchanger.
// The CTA's chunks form one stream t = u*C + j (u-th volume of the cluster, j-th chunk of this CTA's slice).  Consumers
// alternate P1(t), P2(t - L): pass 2 trails pass 1 by L >= C chunks, so L stages hold data between the passes and the other
// STAGES - L are loading; the L - C chunks of slack hide the exchange.
template <typename T, int CHUNK_BYTES, int STAGES, int NCW>
__global__ void __launch_bounds__((NCW + 2) * 32, 1) fused_cluster_kernel(const FusedParams p) {
    const uint32_t C = (uint32_t)p.xc_chunks, L = (uint32_t)p.xc_lag;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* ring = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * CHUNK_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* part_full = empty + STAGES;       // [2] consumers -> exchanger: NCW warp partials are in pbuf[b]
    uint64_t* x_full = part_full + 2;           // [2] every CTA of the cluster -> exchanger: xbuf[b] holds all CS partials
    uint64_t* rk_full = x_full + 2;             // [2] exchanger -> consumers: rowk[b] is valid
    uint64_t* rk_empty = rk_full + 2;           // [2] consumers -> exchanger: rowk[b] has been read by all
    float(*pbuf)[8] = reinterpret_cast<float(*)[8]>(rk_empty + 2);              // [2][NCW][8]
    float(*xbuf)[8] = pbuf + 2 * NCW;                                           // [2][kMaxCluster][8], written by the peers
    volatile float* rowk = reinterpret_cast<volatile float*>(xbuf + 2 * kMaxCluster);   // [2][8]

    const Geometry& g = p.f.g;
    const uint32_t CS = cluster_nctarank(), q = cluster_ctarank();
    const uint32_t ncl = gridDim.x / CS, cid = blockIdx.x / CS;
    const uint32_t nunits = cid < (uint32_t)g.R ? ((uint32_t)g.R - cid + ncl - 1) / ncl : 0;   // volumes of this cluster
    const uint32_t kq = q * C;                                                                  // first chunk of this CTA's slice
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, NCW); }
        for (int b = 0; b < 2; ++b) {
            mbar_init(part_full + b, NCW); mbar_init(x_full + b, CS); mbar_init(rk_full + b, 1); mbar_init(rk_empty + b, NCW);
        }
        mbar_fence_init();
    }
    __syncthreads();
    cluster_sync_all();         // peers' barriers exist before anybody arrives on them

    if (warp == 0) {
        // ================= producer: every chunk of the slice is fetched once =================
        if (lane == 0) {
            const uint64_t pol = l2_policy_evict_first();
            const uint8_t* src = reinterpret_cast<const uint8_t*>(p.f.heat);
            uint32_t it = 0;
            for (uint32_t u = 0; u < nunits; ++u) {
                const uint32_t r = cid + u * ncl;
                const uint8_t* base = src + ((size_t)r * g.N + (size_t)kq * g.CE) * sizeof(T);
#pragma unroll 1
                for (uint32_t j = 0; j < C; ++j, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(empty + s, ph ^ 1, 1);
                    mbar_expect_tx(full + s, CHUNK_BYTES);
                    bulk_g2s(ring + (size_t)s * CHUNK_BYTES, base + (size_t)j * CHUNK_BYTES, CHUNK_BYTES, full + s, pol);
                }
            }
        }
        __syncwarp();
    } else if (warp == NCW + 1) {
        // ================= exchanger =================
        for (uint32_t u = 0; u < nunits; ++u) {
            const uint32_t r = cid + u * ncl;
            const uint32_t b = u & 1, ph = (u >> 1) & 1;
            // targets of this volume: fetched before the wait so their latency hides behind pass 1
            const float v = __ldg(p.f.vis + r), hd = __ldg(p.f.have_depth + r / g.J);
            const float gtx = __ldg(p.f.gt + 3 * (size_t)r), gty = __ldg(p.f.gt + 3 * (size_t)r + 1), gtz = __ldg(p.f.gt + 3 * (size_t)r + 2);
            if (lane == 0) mbar_wait(part_full + b, ph, 2);
            __syncwarp();
            Acc t;
            t.reset();
            if (lane < NCW) t = partial_from_smem(pbuf[b * NCW + lane]);
            t = acc_warp_merge(t);
            if (CS > 1) {
                // lane k hands this CTA's partial to CTA k of the cluster (slot q of its xbuf[b]) and arrives on its barrier;
                // xbuf[b] / x_full[b] are free again: every peer finished reading unit u-2 before it published unit u-1,
                // and nobody reaches unit u's publish without having seen all of unit u-1's.
                if (lane < (int)CS) {
                    const uint32_t dst = map_to_cta(xbuf[b * kMaxCluster + q], (uint32_t)lane);
                    st_cluster_f4(dst, t.m, t.l, t.sx, t.sy);
                    st_cluster_f4(dst + 16, t.sz, t.mx, 0.f, 0.f);
                    mbar_arrive_remote(map_to_cta(x_full + b, (uint32_t)lane));
                }
                mbar_wait_cluster(x_full + b, ph);
                __syncwarp();
                t.reset();
                if (lane < (int)CS) t = partial_from_smem(xbuf[b * kMaxCluster + lane]);
                t = acc_warp_merge(t);          // same values, same instruction sequence in all CS CTAs: identical bits
            }
            const float inv = 1.0f / t.l;
            const float cx = t.sx * inv, cy = t.sy * inv, cz = t.sz * inv;
            // consumers are done with the constants of unit u-2 (one lane waits, the warp re-converges before rk_full is signalled)
            if (lane == 0) mbar_wait(rk_empty + b, ph ^ 1, 3);
            __syncwarp();
            if (lane == 0) {
                const float sc = v * p.loss_scale * inv;           // upstream gradient 1, pre-divided by l
                volatile float* rk = rowk + b * 8;
                rk[0] = t.c;
                rk[1] = sc * sgn(cx - gtx);
                rk[2] = sc * sgn(cy - gty);
                rk[3] = sc * sgn(cz - gtz) * hd;
                rk[4] = cx; rk[5] = cy; rk[6] = cz;
                mbar_arrive(rk_full + b);
            }
            // ---- off the critical path: outputs of this joint-volume (plain stores; published by the final ticket)
            if (q == 0 && lane == 0) {
                p.f.coords[3 * (size_t)r + 0] = cx;
                p.f.coords[3 * (size_t)r + 1] = cy;
                p.f.coords[3 * (size_t)r + 2] = cz;
                if (p.f.stats) {
                    const float f = (t.m == -INFINITY) ? 0.f : ex2(t.c - safe_c(t.mx));
                    p.f.stats[2 * (size_t)r] = t.mx;
                    p.f.stats[2 * (size_t)r + 1] = t.l * f;
                }
                __stcg(p.f.row_loss + r, (fabsf(cx - gtx) * v + fabsf(cy - gty) * v + fabsf(cz - gtz) * v * hd) / 3.f);
            }
        }
        // ---- one ticket per CTA: the last one of the grid reduces the loss terms in index order (loss.py:52)
        int t2 = 0;
        if (lane == 0) {
            __threadfence();
            t2 = atomicAdd(p.f.done_rows, 1);
        }
        t2 = __shfl_sync(0xffffffffu, t2, 0);
        if (t2 == (int)gridDim.x - 1) {
            __threadfence();
            float s = 0.f;
            for (int i = lane; i < g.R; i += 32) s += __ldcg(p.f.row_loss + i);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) { p.f.loss[0] = s / (float)g.R; *p.f.done_rows = 0; }
        }
        __syncwarp();
    } else {
        // ================= consumers =================
        constexpr int NC = NCW * 32;
        constexpr int VPC = CHUNK_BYTES / 16;
        constexpr int U = (VPC / NC) < 1 ? 1 : ((VPC / NC) > 4 ? 4 : (VPC / NC));
        constexpr int QPV = Elem<T>::QPV;
        const int tid = threadIdx.x - 32, wid = warp - 1;
        const bool fast = fast_ok<NC, VPC>(g);
        const uint32_t Fv = fast ? g.divFv.d : 1;
        const float x0f = u2f((uint32_t)(tid % Fv) * (4 * QPV)), rsf = u2f(NC / Fv), hf = u2f((uint32_t)g.H);
        uint8_t* out = reinterpret_cast<uint8_t*>(p.grad_heat);
        const uint32_t total = nunits * C;
        Acc a;
        a.reset();
        RowK rk = {};
        float tx[4 * QPV] = {};
        uint32_t u1 = 0, j1 = 0;        // pass-1 position (volume of this cluster, chunk of the slice)
        uint32_t u2 = 0, j2 = 0;        // pass-2 position
        uint32_t s1 = 0, ph1 = 0, s2 = 0;   // ring stage / parity of t, ring stage of t - L
#pragma unroll 1
        for (uint32_t t = 0; t < total + L; ++t) {
            if (t < total) {
                // ---- pass 1 of chunk t: online softmax + coordinate moments; the stage stays full
                mbar_wait(full + s1, ph1, 4);
                const uint8_t* st = ring + (size_t)s1 * CHUNK_BYTES;
                const uint32_t k = kq + j1;
                auto load = [&](uint32_t iv) { return lds16(st + (size_t)iv * 16); };
                if (fast) consume_chunk_fast<T, U, NC, VPC, true>(a, g, VPC, k * VPC, tid, rsf, hf, load);
                else consume_chunk<T, U, NC>(a, g, VPC, (k * g.CE) >> 2, tid, load);
                if (++j1 == C) {            // this warp's partial of volume u1 -> exchanger
                    if (fast) a.sx = fmaf(x0f, a.l, a.sx);
                    a = acc_warp_merge(a);
                    if (lane == 0) {
                        partial_to_smem(pbuf[(u1 & 1) * NCW + wid], a);
                        mbar_arrive(part_full + (u1 & 1));
                    }
                    a.reset();
                    j1 = 0; ++u1;
                }
                if (++s1 == STAGES) { s1 = 0; ph1 ^= 1; }
            }
            if (t >= L) {
                // ---- pass 2 of chunk t - L: the same stage again, gradient out, stage back to the producer
                if (j2 == 0) {
                    const uint32_t b = u2 & 1;
                    mbar_wait(rk_full + b, (u2 >> 1) & 1, 5);
                    const volatile float* c = rowk + b * 8;
                    rk.c = c[0]; rk.gx = c[1]; rk.gy = c[2]; rk.gz = c[3]; rk.cx = c[4]; rk.cy = c[5]; rk.cz = c[6];
                    __syncwarp();
                    if (lane == 0) mbar_arrive(rk_empty + b);
                    make_tx<4 * QPV>(rk, x0f, tx);
                }
                const uint32_t r = cid + u2 * ncl;
                const uint32_t k = kq + j2;
                const uint8_t* st = ring + (size_t)s2 * CHUNK_BYTES;
                uint8_t* dst = out + ((size_t)r * g.N + (size_t)k * g.CE) * sizeof(T);
                auto load = [&](uint32_t iv) { return lds16(st + (size_t)iv * 16); };
                if (fast) bwd_chunk_fast<T, U, NC, VPC, true>(rk, tx, g, VPC, k * VPC, tid, rsf, hf, dst, load);
                else bwd_chunk<T, U, NC>(rk, g, VPC, (k * g.CE) >> 2, tid, dst, load);
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + s2);
                if (++j2 == C) { j2 = 0; ++u2; }
                if (++s2 == STAGES) s2 = 0;
            }
        }
        __syncwarp();
    }
    // nobody leaves while a peer may still write into its shared memory or arrive on its barriers
    cluster_sync_all();
}

constexpr int kNCW = 16;
constexpr size_t kRingBytes = 7 * 32768;       // 7 x 32 KiB or 14 x 16 KiB stages
constexpr size_t smem_bytes(int stages) {
    return kRingBytes + (2 * (size_t)stages + 8) * sizeof(uint64_t) + (size_t)(2 * kNCW + 2 * kMaxCluster + 2) * 8 * sizeof(float);
}

template <typename T, int CB>
cudaError_t prepare(int CS, cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr) {
    constexpr int ST = (int)(kRingBytes / CB);
    auto kern = fused_cluster_kernel<T, CB, ST, kNCW>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(ST));
    if (e != cudaSuccess) return e;
    if (CS > 8) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    *cfg = cudaLaunchConfig_t{};
    cfg->gridDim = dim3((unsigned)CS);
    cfg->blockDim = dim3((kNCW + 2) * 32);
    cfg->dynamicSmemBytes = smem_bytes(ST);
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg->attrs = attr;
    cfg->numAttrs = 1;
    return cudaSuccess;
}

template <typename T>
int max_clusters_t(int CS) {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    if (prepare<T, 32768>(CS, &cfg, attr) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    cfg.gridDim = dim3((unsigned)CS * 64);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, fused_cluster_kernel<T, 32768, 7, kNCW>, &cfg) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return n;
}

int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

}  // namespace

// Plan for this geometry: cluster size CS (CTAs per joint-volume), chunk bytes, chunks per slice and pass-2 lag; CS = 0
// when K5c does not apply.  The joint-volume must split into CS equal slices of whole chunks that fit the ring with at
// least two stages left for loading.  Tuning overrides: IHPR_K5C_CS, IHPR_K5C_CHUNK (32768 | 16384), IHPR_K5C_LAG.
FusedClusterPlan fused_cluster_plan(const Geometry& g, int dtype) {
    FusedClusterPlan pl = {0, 0, 0, 0};
    const uint64_t es = dtype == 0 ? 4 : 2;
    const uint64_t row_bytes = (uint64_t)g.N * es;
    if ((uint64_t)g.CE * es != 32768) return pl;            // geometry of the default (32 KiB chunk) configuration expected
    const int chunk = env_int("IHPR_K5C_CHUNK", 16384);
    if (chunk != 32768 && chunk != 16384) return pl;
    const int stages = (int)(kRingBytes / (size_t)chunk);
    int cs = env_int("IHPR_K5C_CS", 0);
    if (cs == 0) {
        // smallest power-of-two cluster whose slice leaves at least 5/8 of the ring for the pass-2 lag and loads in flight
        for (cs = 1; cs <= kMaxCluster; cs *= 2)
            if (row_bytes % ((uint64_t)cs * chunk) == 0 && row_bytes / cs <= kRingBytes * 3 / 8) break;
    }
    if (cs < 1 || cs > kMaxCluster || (cs & (cs - 1)) != 0 || row_bytes % ((uint64_t)cs * chunk) != 0) return pl;
    const int C = (int)(row_bytes / cs / chunk);
    // pass 2 must trail pass 1 by well over a slice: the exchange (two warp merges on a busy SM plus the slowest partner's
    // jitter) takes ~3 us, and a consumer that reaches pass 2 before its constants stalls the whole cluster (measured:
    // lag C + 2 -> 880 us, C + 4 -> 284 us, C + 6 -> 265 us for 64^3 fp32 at B = 32)
    int lag = env_int("IHPR_K5C_LAG", C + 6 < stages - 4 ? C + 6 : stages - 4);
    if (lag < C) lag = C;
    if (C < 1 || lag > stages - 2) return pl;
    pl.cluster = cs; pl.chunk_bytes = chunk; pl.chunks = C; pl.lag = lag;
    return pl;
}

// how many clusters of CS CTAs the current device can hold at once (cached per device / dtype / CS; 0 = cannot launch)
int fused_cluster_capacity(int dtype, int CS) {
    static std::atomic<int> cache[16][2][5];         // value + 1; 0 = not asked yet
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return 0;
    int lg = 0;
    while ((1 << lg) < CS) ++lg;
    if (lg > 4) return 0;
    std::atomic<int>& slot = cache[dev][dtype ? 1 : 0][lg];
    int v = slot.load(std::memory_order_relaxed);
    if (v == 0) {
        v = 1 + (dtype == 0 ? max_clusters_t<float>(CS) : max_clusters_t<__nv_bfloat16>(CS));
        slot.store(v, std::memory_order_relaxed);
    }
    int n = v - 1;
    const int cap = env_int("IHPR_K5C_CLUSTERS", 0);
    if (cap > 0 && n > cap) n = cap;
    if (getenv("IHPR_K5C_VERBOSE")) fprintf(stderr, "[ihpr] K5c: cluster size %d -> %d active clusters (occupancy API %d)\n", CS, n, v - 1);
    return n;
}

template <typename T, int CB>
static cudaError_t launch_fused_cluster_t(FusedParams p, const FusedClusterPlan& pl, int nclusters, cudaStream_t s) {
    constexpr int ST = (int)(kRingBytes / CB);
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    cudaError_t e = prepare<T, CB>(pl.cluster, &cfg, attr);
    if (e != cudaSuccess) return e;
    cfg.gridDim = dim3((unsigned)(nclusters * pl.cluster));
    cfg.stream = s;
    p.xc_chunks = pl.chunks;
    p.xc_lag = pl.lag;
    p.f.g.CE = (uint32_t)(CB / sizeof(T));
    p.f.g.nch = (p.f.g.N + p.f.g.CE - 1) / p.f.g.CE;
    p.f.g.Gt = (uint64_t)p.f.g.R * p.f.g.nch;
    return cudaLaunchKernelEx(&cfg, fused_cluster_kernel<T, CB, ST, kNCW>, p);
}

cudaError_t launch_fused_cluster(const FusedParams& p, int dtype, const FusedClusterPlan& pl, int nclusters, cudaStream_t s) {
    if (pl.chunk_bytes == 32768)
        return dtype == 0 ? launch_fused_cluster_t<float, 32768>(p, pl, nclusters, s) : launch_fused_cluster_t<__nv_bfloat16, 32768>(p, pl, nclusters, s);
    return dtype == 0 ? launch_fused_cluster_t<float, 16384>(p, pl, nclusters, s) : launch_fused_cluster_t<__nv_bfloat16, 16384>(p, pl, nclusters, s);
}

}  // namespace ihpr

// softargmax_fused.cu -- K5: integral L1 loss forward AND heat-map gradient in ONE launch (sm_100a).
//
// The reference's training step runs JointLocationLoss forward, then autograd backward
// (/root/reference/main/train.py:67-71): 11-15 V of DRAM traffic.  K1 + K2 bring that to the algorithmic
// 3 V (read, re-read, write).  K5 removes the re-read: a joint-volume is split over S persistent CTAs
// (S*unit = N*s bytes, unit <= 256 KiB so that the units all CTAs have in flight, 148 * 256 KiB = 37 MiB,
// stay resident in the 126 MB L2 -- tools/l2reuse.cu measures where that stops working); each CTA
//   pass 1  streams its unit from HBM through the TMA ring (L2 evict_last) and accumulates (m, l, sx, sy, sz),
//   merge   publishes the partial, waits for its S-1 partner CTAs (all co-resident: cooperative launch),
//           merges the S slots in slot order, forms coords, the loss term and g = dLoss/dcoords (loss.py:49-52
//           with upstream gradient 1),
//   pass 2  streams the same unit again -- now L2 hits (evict_first) -- and writes grad_heat.
// DRAM traffic: read V + write V.  The producer warp never stops: pass-2 chunks of this unit and pass-1 chunks
// of the next one are already in the ring while the consumers exchange partials.
#include <cooperative_groups.h>

#include "ihpr_device.cuh"

namespace ihpr {

__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Roles: warp 0 = TMA producer, warp 1 = exchanger (merges warp partials, trades CTA partials with the partner
// CTAs, forms the per-volume backward constants and writes coords / stats / loss), warps 2.. = NCW consumers.
// The stream is software-pipelined by one unit: consumers run  pass1(u0), pass1(u1), pass2(u0), pass1(u2),
// pass2(u1), ...  so the exchange of unit i happens while they are busy with pass 1 of unit i+1 and never
// stalls them; at most two units per CTA are live in L2.
template <typename T, int CHUNK_BYTES, int STAGES, int NCW>
__global__ void __launch_bounds__(NCW * 32 + 64, 1) fused_ring_kernel(const FusedParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* ring = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * CHUNK_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* part_full = empty + STAGES;       // [2] consumers -> exchanger: NCW warp partials are in pbuf[par]
    uint64_t* rk_full = part_full + 2;          // [2] exchanger -> consumers: rowk[par] is valid
    uint64_t* rk_empty = rk_full + 2;           // [2] consumers -> exchanger: rowk[par] has been read by all
    float(*pbuf)[8] = reinterpret_cast<float(*)[8]>(rk_empty + 2);      // [2][NCW][8]
    volatile float* rowk = reinterpret_cast<volatile float*>(pbuf + 2 * NCW);   // [2][8]

    const Geometry& g = p.f.g;
    const int S = p.S;
    const uint32_t cta = blockIdx.x;
    const uint32_t ngroups = gridDim.x / S, group = cta / S, q = cta % S;
    const uint32_t k0 = (uint32_t)((uint64_t)g.nch * q / S), k1 = (uint32_t)((uint64_t)g.nch * (q + 1) / S);
    const uint32_t nunits = group < (uint32_t)g.R ? ((uint32_t)g.R - group + ngroups - 1) / ngroups : 0;   // volumes of this group
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, NCW); }
        for (int b = 0; b < 2; ++b) { mbar_init(part_full + b, NCW); mbar_init(rk_full + b, 1); mbar_init(rk_empty + b, NCW); }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == 0) {
        // ================= producer =================
        if (lane == 0) {
            const uint64_t pol_keep = l2_policy_evict_last(), pol_drop = l2_policy_evict_first();
            const uint8_t* src = reinterpret_cast<const uint8_t*>(p.f.heat);
            uint32_t it = 0;
            auto issue = [&](uint32_t u, int pass) {
                const uint32_t r = group + u * ngroups;
                for (uint32_t k = k0; k < k1; ++k, ++it) {
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(empty + s, ph ^ 1, 1);
                    const uint32_t e0 = k * g.CE;
                    const uint32_t bytes = min(g.CE, g.N - e0) * (uint32_t)sizeof(T);
                    mbar_expect_tx(full + s, bytes);
                    bulk_g2s(ring + (size_t)s * CHUNK_BYTES, src + ((size_t)r * g.N + e0) * sizeof(T), bytes, full + s, pass ? pol_drop : pol_keep);
                }
            };
            if (nunits) issue(0, 0);
            for (uint32_t u = 0; u < nunits; ++u) {
                if (u + 1 < nunits) issue(u + 1, 0);
                issue(u, 1);
            }
        }
        return;
    }

    if (warp == 1) {
        // ================= exchanger =================
        for (uint32_t u = 0; u < nunits; ++u) {
            const uint32_t r = group + u * ngroups;
            const uint32_t par = u & 1, ph = (u >> 1) & 1;
            // targets of this volume: fetched before the wait so their latency hides behind pass 1
            const float v = __ldg(p.f.vis + r), hd = __ldg(p.f.have_depth + r / g.J);
            const float gtx = __ldg(p.f.gt + 3 * (size_t)r), gty = __ldg(p.f.gt + 3 * (size_t)r + 1), gtz = __ldg(p.f.gt + 3 * (size_t)r + 2);
            if (lane == 0) mbar_wait(part_full + par, ph, 2);
            __syncwarp();
            Acc t;
            t.reset();
            if (lane < NCW) t = partial_from_smem(pbuf[par * NCW + lane]);
            t = acc_warp_merge(t);
            if (S > 1) {
                float* slots = p.f.partials + (size_t)r * p.f.maxslots * 8;
                if (lane == 0) {
                    partial_to_global(slots + q * 8, t);
                    __threadfence();
                    atomicAdd(p.f.row_count + r, 1);
                    #ifdef IHPR_DEBUG_HANG
                    for (unsigned long long spins = 0; ld_acquire(p.f.row_count + r) < S; ++spins) {
                        if (spins > (1ull << 20)) { printf("HANG cta %d exchange volume %u count %d of %d\n", blockIdx.x, r, ld_acquire(p.f.row_count + r), S); __trap(); }
                        __nanosleep(32);
                    }
#else
                    while (ld_acquire(p.f.row_count + r) < S) __nanosleep(32);
#endif
                }
                __syncwarp();
                __threadfence();
                t.reset();
                if (lane < S) t = partial_from_global(slots + lane * 8);
                t = acc_warp_merge(t);          // same instruction sequence in all S CTAs: identical bits
            }
            const float inv = 1.0f / t.l;
            const float cx = t.sx * inv, cy = t.sy * inv, cz = t.sz * inv;
            // consumers are done with the constants of unit u-2.  One lane waits and the warp re-converges before
            // rk_full is signalled: a lane that polled rk_empty only after the consumers (released by rk_full) had
            // already completed its NEXT phase would wait on the wrong parity forever.
            if (lane == 0) mbar_wait(rk_empty + par, ph ^ 1, 3);
            __syncwarp();
            if (lane == 0) {
                const float sc = v * p.loss_scale * inv;           // upstream gradient 1, pre-divided by l
                volatile float* rk = rowk + par * 8;
                rk[0] = t.c;
                rk[1] = sc * sgn(cx - gtx);
                rk[2] = sc * sgn(cy - gty);
                rk[3] = sc * sgn(cz - gtz) * hd;
                rk[4] = cx; rk[5] = cy; rk[6] = cz;
                mbar_arrive(rk_full + par);
            }
            // ---- off the critical path: re-arm the ticket, outputs, loss
            if (S > 1 && lane == 0) {
                if (atomicAdd(p.row_pass + r, 1) == S - 1) { p.f.row_count[r] = 0; p.row_pass[r] = 0; }
            }
            if (q == 0) {
                int t2 = 0;
                if (lane == 0) {
                    p.f.coords[3 * (size_t)r + 0] = cx;
                    p.f.coords[3 * (size_t)r + 1] = cy;
                    p.f.coords[3 * (size_t)r + 2] = cz;
                    if (p.f.stats) {
                        const float f = (t.m == -INFINITY) ? 0.f : ex2(t.c - safe_c(t.mx));
                        p.f.stats[2 * (size_t)r] = t.mx;
                        p.f.stats[2 * (size_t)r + 1] = t.l * f;
                    }
                    __stcg(p.f.row_loss + r, (fabsf(cx - gtx) * v + fabsf(cy - gty) * v + fabsf(cz - gtz) * v * hd) / 3.f);
                    __threadfence();
                    t2 = atomicAdd(p.f.done_rows, 1);
                }
                t2 = __shfl_sync(0xffffffffu, t2, 0);
                if (t2 == g.R - 1) {
                    __threadfence();
                    float s = 0.f;
                    for (int i = lane; i < g.R; i += 32) s += __ldcg(p.f.row_loss + i);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                    if (lane == 0) { p.f.loss[0] = s / (float)g.R; *p.f.done_rows = 0; }
                }
            }
        }
        return;
    }

    // ================= consumers =================
    constexpr int NC = NCW * 32;
    constexpr int VPC = CHUNK_BYTES / 16;
    constexpr int U = (VPC / NC) < 1 ? 1 : ((VPC / NC) > 4 ? 4 : (VPC / NC));
    constexpr int QPV = Elem<T>::QPV;
    const int tid = threadIdx.x - 64, wid = warp - 2;
    const bool fast = fast_ok<NC, VPC>(g);
    const uint32_t Fv = fast ? g.divFv.d : 1;
    const float x0f = u2f((uint32_t)(tid % Fv) * (4 * QPV)), rsf = u2f(NC / Fv), hf = u2f((uint32_t)g.H);
    uint8_t* out = reinterpret_cast<uint8_t*>(p.grad_heat);
    uint32_t it = 0;

    auto pass1 = [&](uint32_t u) {          // online softmax + coordinate moments of this CTA's unit of volume u
        Acc a;
        a.reset();
        for (uint32_t k = k0; k < k1; ++k, ++it) {
            const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
            const uint32_t e0 = k * g.CE;
            const uint32_t n_vec = min(g.CE, g.N - e0) / (4 * QPV);
            mbar_wait(full + s, ph, 4);
            const uint8_t* st = ring + (size_t)s * CHUNK_BYTES;
            auto load = [&](uint32_t iv) { return lds16(st + (size_t)iv * 16); };
            if (fast) {
                if (n_vec == VPC) consume_chunk_fast<T, U, NC, VPC, true>(a, g, n_vec, k * VPC, tid, rsf, hf, load);
                else consume_chunk_fast<T, U, NC, VPC, false>(a, g, n_vec, k * VPC, tid, rsf, hf, load);
            } else {
                consume_chunk<T, U, NC>(a, g, n_vec, e0 >> 2, tid, load);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
        }
        if (fast) a.sx = fmaf(x0f, a.l, a.sx);
        a = acc_warp_merge(a);
        if (lane == 0) {
            partial_to_smem(pbuf[(u & 1) * NCW + wid], a);
            mbar_arrive(part_full + (u & 1));
        }
    };
    auto pass2 = [&](uint32_t u) {          // the same unit again (L2 hits), gradient out
        const uint32_t r = group + u * ngroups;
        const uint32_t par = u & 1;
        mbar_wait(rk_full + par, (u >> 1) & 1, 5);
        RowK rk;
        const volatile float* c = rowk + par * 8;
        rk.c = c[0]; rk.gx = c[1]; rk.gy = c[2]; rk.gz = c[3]; rk.cx = c[4]; rk.cy = c[5]; rk.cz = c[6];
        __syncwarp();
        if (lane == 0) mbar_arrive(rk_empty + par);
        float tx[4 * QPV];
        make_tx<4 * QPV>(rk, x0f, tx);
        for (uint32_t k = k0; k < k1; ++k, ++it) {
            const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
            const uint32_t e0 = k * g.CE;
            const uint32_t n_vec = min(g.CE, g.N - e0) / (4 * QPV);
            mbar_wait(full + s, ph, 6);
            const uint8_t* st = ring + (size_t)s * CHUNK_BYTES;
            uint8_t* dst = out + ((size_t)r * g.N + e0) * sizeof(T);
            auto load = [&](uint32_t iv) { return lds16(st + (size_t)iv * 16); };
            if (fast) {
                if (n_vec == VPC) bwd_chunk_fast<T, U, NC, VPC, true>(rk, tx, g, n_vec, k * VPC, tid, rsf, hf, dst, load);
                else bwd_chunk_fast<T, U, NC, VPC, false>(rk, tx, g, n_vec, k * VPC, tid, rsf, hf, dst, load);
            } else {
                bwd_chunk<T, U, NC>(rk, g, n_vec, e0 >> 2, tid, dst, load);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
        }
    };

    if (nunits) pass1(0);
    for (uint32_t u = 0; u < nunits; ++u) {
        if (u + 1 < nunits) pass1(u + 1);
        pass2(u);
    }
}

// grad *= grad_out, skipped entirely when grad_out == 1 (the loss.backward() case)
template <typename T>
__global__ void scale_kernel(T* __restrict__ gh, size_t n_vec, size_t n_tail_start, size_t n, const float* __restrict__ grad_out) {
    const float s = __ldg(grad_out);
    if (s == 1.0f) return;
    constexpr int QPV = Elem<T>::QPV;
    uint4* v = reinterpret_cast<uint4*>(gh);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) {
        float f[QPV][4];
        Elem<T>::unpack(v[i], f);
#pragma unroll
        for (int a = 0; a < QPV; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) f[a][b] *= s;
        v[i] = Elem<T>::pack(f);
    }
    for (size_t i = n_tail_start + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        Elem<T>::store1(gh + i, Elem<T>::load1(gh + i) * s);
}

int fused_split(const Geometry& g, int dtype) {
    const uint64_t row_bytes = (uint64_t)g.N * (dtype == 0 ? 4 : 2);
    int S = 1;
    // two units per CTA are live (pass 2 of one, pass 1 of the next): 148 * 2 * 128 KiB = 37 MiB stays in L2
    while (row_bytes / S > (128u << 10) && S < 16 && (uint32_t)(2 * S) <= g.nch) S *= 2;
    return S;
}

template <typename T>
static cudaError_t launch_fused_t(const FusedParams& p, int num_sms, cudaStream_t s) {
    constexpr int CB = 32768, ST = 6, NCW = 16;
    auto kern = fused_ring_kernel<T, CB, ST, NCW>;
    const size_t smem = (size_t)ST * CB + (2 * ST + 6) * sizeof(uint64_t) + (2 * NCW + 2) * 8 * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int S = p.S;
    unsigned G = (unsigned)(num_sms / S) * S;
    const unsigned need = (unsigned)p.f.g.R * S;
    if (G > need) G = need;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G);
    cfg.blockDim = dim3(NCW * 32 + 64);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;      // partner CTAs wait for each other: they must be co-resident
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, p);
}

cudaError_t launch_fused(const FusedParams& p, int dtype, int num_sms, cudaStream_t s) {
    return dtype == 0 ? launch_fused_t<float>(p, num_sms, s) : launch_fused_t<__nv_bfloat16>(p, num_sms, s);
}

void launch_scale(void* grad, size_t n, int dtype, bool aligned, const float* grad_out, int num_sms, cudaStream_t s) {
    const size_t epv = dtype == 0 ? 4 : 8;
    const size_t n_vec = aligned ? n / epv : 0;
    if (dtype == 0) scale_kernel<float><<<num_sms * 4, 512, 0, s>>>(static_cast<float*>(grad), n_vec, n_vec * epv, n, grad_out);
    else scale_kernel<__nv_bfloat16><<<num_sms * 4, 512, 0, s>>>(static_cast<__nv_bfloat16*>(grad), n_vec, n_vec * epv, n, grad_out);
}

}  // namespace ihpr

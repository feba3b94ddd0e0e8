// softargmax_fused.cu -- K5: integral L1 loss forward AND heat-map gradient in ONE launch (sm_100a).
//
// The reference's training step runs JointLocationLoss forward, then autograd backward
// (/root/reference/main/train.py:67-71): 11-15 V of DRAM traffic.  K1 + K2 bring that to the algorithmic
// 3 V (read, re-read, write).  K5 removes the re-read: a joint-volume is split over S persistent CTAs
// (S*unit = N*s bytes; the units all CTAs hold between their two passes, 148 * 256 KiB = 37 MiB, stay resident in the
// 126 MB L2 -- tools/l2reuse.cu measures where that stops working); each CTA
//   pass 1  streams its unit from HBM through the TMA ring (L2 evict_last) and accumulates (m, l, sx, sy, sz),
//   merge   publishes the partial, waits for its S-1 partner CTAs (all co-resident: cooperative launch),
//           merges the S slots in slot order, forms coords, the loss term and g = dLoss/dcoords (loss.py:49-52
//           with upstream gradient 1),
//   pass 2  streams the same unit again -- now L2 hits (evict_first) -- and writes grad_heat.
// DRAM traffic: read V + write V.  The producer warp never stops: pass-2 chunks of this unit and pass-1 chunks
// of the next one are already in the ring while the consumers exchange partials.
#include <cstdlib>
#include <type_traits>

#include "ihpr_device.cuh"

namespace ihpr {

__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// Roles: warp 0 = TMA producer, warps 1..NCW = consumers, the last DEPTH+1 warps = exchangers (merge warp partials, trades CTA partials with the partner
// CTAs, form the per-volume backward constants and write coords / stats / loss term).
// The stream is software-pipelined by DEPTH units: consumers run pass1(u0..u_{DEPTH-1}), then alternate
// pass1(u+DEPTH), pass2(u), so the exchange of unit u has DEPTH * (pass1 + pass2) - pass2 of time before its result
// is needed; DEPTH + 1 units per CTA are live in L2.
//
// Exchange protocol (no fences, no atomics, one L2 round trip after the slowest partner): every CTA of the group
// stores its 6-word partial as six 8-byte {value, tag} pairs into its slot of the joint-volume's exchange row; every
// exchanger polls all S slots until each pair carries this launch's tag (one 64-bit scalar store / load per pair: single-copy atomic, so a
// matching tag proves the value next to it).  tag = launch epoch + 1; the epoch lives in the workspace and is bumped
// by the last CTA to finish.
// one 64-bit scalar access each way: a .v2 access is two independent 32-bit accesses under the PTX memory model, a .b64 one is
// single-copy atomic, so a matching tag really does vouch for the value stored with it
__device__ __forceinline__ void st_pair(uint2* p, float v, uint32_t tag) {
    const uint64_t w = ((uint64_t)tag << 32) | (uint64_t)__float_as_uint(v);
    asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ uint2 ld_pair(const uint2* p) {
    uint64_t w;
    asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    return make_uint2((uint32_t)w, (uint32_t)(w >> 32));
}

template <typename T, int CHUNK_BYTES, int STAGES, int NCW, int DEPTH>
__global__ void __launch_bounds__((NCW + DEPTH + 2) * 32, 1) fused_ring_kernel(const FusedParams p) {
    constexpr int NB = DEPTH + 1;               // units in flight between pass 1 and pass 2 = exchanger warps
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* ring = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * CHUNK_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* part_full = empty + STAGES;       // [NB] consumers -> exchanger: NCW warp partials are in pbuf[b]
    uint64_t* rk_full = part_full + NB;         // [NB] exchanger -> consumers: rowk[b] is valid
    uint64_t* rk_empty = rk_full + NB;          // [NB] consumers -> exchanger: rowk[b] has been read by all
    float(*pbuf)[8] = reinterpret_cast<float(*)[8]>(rk_empty + NB);             // [NB][NCW][8]
    volatile float* rowk = reinterpret_cast<volatile float*>(pbuf + NB * NCW);  // [NB][8]

    const Geometry& g = p.f.g;
    const int S = p.S;
    const uint32_t cta = blockIdx.x;
    const uint32_t ngroups = gridDim.x / S, group = cta / S, q = cta % S;
    const uint32_t k0 = (uint32_t)((uint64_t)g.nch * q / S), k1 = (uint32_t)((uint64_t)g.nch * (q + 1) / S);
    const uint32_t nunits = group < (uint32_t)g.R ? ((uint32_t)g.R - group + ngroups - 1) / ngroups : 0;   // volumes of this group
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, NCW); }
        for (int b = 0; b < NB; ++b) { mbar_init(part_full + b, NCW); mbar_init(rk_full + b, 1); mbar_init(rk_empty + b, NCW); }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == 0) {
        // ================= producer =================
        if (lane == 0) {
            uint64_t pol_keep = l2_policy_evict_last(), pol_drop = l2_policy_evict_first();
#ifdef IHPR_TIMING_EXPERIMENTS
            if (p.debug_no_exchange & 6) {
                uint64_t pol_normal;
                asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_normal));
                if (p.debug_no_exchange & 2) pol_keep = pol_normal;
                if (p.debug_no_exchange & 4) pol_drop = pol_normal;
            }
#endif
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
            for (uint32_t u = 0; u < (uint32_t)DEPTH && u < nunits; ++u) issue(u, 0);
            for (uint32_t u = 0; u < nunits; ++u) {
                if (u + DEPTH < nunits) issue(u + DEPTH, 0);
                issue(u, 1);
            }
        }
        return;
    }

    if (warp > NCW) {
        // ================= exchangers: warp NCW + 1 + b serves the units u = b (mod NB) =================
        // (highest warp ids: the warp arbiter favours them, which shortens the exchange)
        // (one exchange costs a few loaded-L2 round trips, ~5 us, more than a unit takes to stream: NB of them run
        //  concurrently so that the exchange THROUGHPUT keeps up; the DEPTH-unit look-ahead hides the latency)
        const uint32_t tag = (uint32_t)__ldcg(p.epoch) + 1u;
        const uint32_t b = warp - (NCW + 1);
        for (uint32_t u = b; u < nunits; u += NB) {
            const uint32_t r = group + u * ngroups;
            const uint32_t ph = (u / NB) & 1;
            // targets of this volume: fetched before the wait so their latency hides behind pass 1
            const float v = __ldg(p.f.vis + r), hd = __ldg(p.f.have_depth + r / g.J);
            const float gtx = __ldg(p.f.gt + 3 * (size_t)r), gty = __ldg(p.f.gt + 3 * (size_t)r + 1), gtz = __ldg(p.f.gt + 3 * (size_t)r + 2);
            if (lane == 0) mbar_wait(part_full + b, ph, 2);
            __syncwarp();
            Acc t;
            t.reset();
            if (lane < NCW) t = partial_from_smem(pbuf[b * NCW + lane]);
            t = acc_warp_merge(t);
            if (S > 1 && !(p.debug_no_exchange & 1)) {
                uint2* row = p.xslots + (size_t)r * (kMaxSplit * 8);
                if (lane < 6) {
                    const float w = lane == 0 ? t.m : lane == 1 ? t.l : lane == 2 ? t.sx : lane == 3 ? t.sy : lane == 4 ? t.sz : t.mx;
                    st_pair(row + q * 8 + lane, w, tag);
                }
                t.reset();
                if (lane < S) {
                    // all six pairs of the partner's slot (one 64-byte line) are polled together: every poll is a trip through memory
                    // queues that the TMA rings keep ~4 us deep (144 CTAs x 192 KiB in flight at 6.5 TB/s), so "poll one pair, then fetch
                    // the other five" cost a second such trip per exchange; stores of different lanes are not ordered, hence a tag per pair
                    const uint2* slot = row + lane * 8;
                    uint2 w[6];
                    for (;;) {
#pragma unroll
                        for (int i = 0; i < 6; ++i) w[i] = ld_pair(slot + i);
                        bool ok = true;
#pragma unroll
                        for (int i = 0; i < 6; ++i) ok = ok && (w[i].y == tag);
                        if (ok) break;
                        __nanosleep(64);
                    }
                    t.m = __uint_as_float(w[0].x); t.l = __uint_as_float(w[1].x); t.sx = __uint_as_float(w[2].x);
                    t.sy = __uint_as_float(w[3].x); t.sz = __uint_as_float(w[4].x); t.mx = __uint_as_float(w[5].x);
                    t.c = safe_c(t.m); t.lim = t.m + kRebaseSlack;
                }
                __syncwarp();
                t = acc_warp_merge(t);          // same instruction sequence in all S CTAs: identical bits
            }
            const float inv = 1.0f / t.l;
            const float cx = t.sx * inv, cy = t.sy * inv, cz = t.sz * inv;
            // consumers are done with the constants of unit u-NB.  One lane waits and the warp re-converges before
            // rk_full is signalled: a lane that polled rk_empty only after the consumers (released by rk_full) had
            // already completed its NEXT phase would wait on the wrong parity forever.
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
        // ---- one ticket per exchanger warp: the last one of the grid reduces the loss terms (index order), bumps the epoch
        int t2 = 0;
        if (lane == 0) {
            __threadfence();
            t2 = atomicAdd(p.f.done_rows, 1);
        }
        t2 = __shfl_sync(0xffffffffu, t2, 0);
        if (t2 == (int)gridDim.x * NB - 1) {
            __threadfence();
            float s = 0.f;
            for (int i = lane; i < g.R; i += 32) s += __ldcg(p.f.row_loss + i);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) { p.f.loss[0] = s / (float)g.R; *p.f.done_rows = 0; *p.epoch = (int)tag; }
        }
        return;
    }

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
            partial_to_smem(pbuf[(u % NB) * NCW + wid], a);
            mbar_arrive(part_full + (u % NB));
        }
    };
    auto pass2 = [&](uint32_t u) {          // the same unit again (L2 hits), gradient out
        const uint32_t r = group + u * ngroups;
        const uint32_t b = u % NB;
        mbar_wait(rk_full + b, (u / NB) & 1, 5);
        RowK rk;
        const volatile float* c = rowk + b * 8;
        rk.c = c[0]; rk.gx = c[1]; rk.gy = c[2]; rk.gz = c[3]; rk.cx = c[4]; rk.cy = c[5]; rk.cz = c[6];
        __syncwarp();
        if (lane == 0) mbar_arrive(rk_empty + b);
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

    for (uint32_t u = 0; u < (uint32_t)DEPTH && u < nunits; ++u) pass1(u);
    for (uint32_t u = 0; u < nunits; ++u) {
        if (u + DEPTH < nunits) pass1(u + DEPTH);
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

// Pipeline depth (units between pass 1 and pass 2 of the same unit) and unit size.  DEPTH + 1 units per CTA are live
// in L2 between their two passes (tools/l2reuse.cu puts the cliff near 148 x 512 KiB); every unit costs one exchange
// (a few loaded-L2 round trips), so smaller units trade more often and deeper look-ahead spills out of L2.  The best
// point was found by sweeping on B200 (profiles/r01_fused_sweep.txt): fp32 -> ~85 KiB units (S = 12 for 64^3, 6 for
// 32x64x64, 24 for 128x64x64) with DEPTH 3; bf16 (twice the voxels, i.e. twice the math, per byte) -> 128 KiB units with
// DEPTH 2.  IHPR_FUSED_DEPTH / IHPR_FUSED_SPLIT override the choice for tuning experiments.
static int fused_depth(int dtype) {
    const char* e = getenv("IHPR_FUSED_DEPTH");
    const int d = e ? atoi(e) : (dtype == 0 ? 3 : 2);
    return d < 1 ? 1 : (d > 3 ? 3 : d);
}

int fused_split(const Geometry& g, int dtype) {
    const uint64_t row_bytes = (uint64_t)g.N * (dtype == 0 ? 4 : 2);
    int S;
    if (const char* e = getenv("IHPR_FUSED_SPLIT")) {
        S = atoi(e);
    } else {
        const uint64_t target = (dtype == 0 ? 85u : 128u) << 10;
        S = (int)((row_bytes + target / 2) / target);
    }
    if (S > kMaxSplit) S = kMaxSplit;
    if ((uint32_t)S > g.nch) S = (int)g.nch;
    return S < 1 ? 1 : S;
}

template <typename T, int DEPTH, int ST, int CB = 32768>
static cudaError_t launch_fused_ds(const FusedParams& p, int num_sms, cudaStream_t s) {
    constexpr int NCW = 16, NB = DEPTH + 1;
    auto kern = fused_ring_kernel<T, CB, ST, NCW, DEPTH>;
    const size_t smem = (size_t)ST * CB + (2 * ST + 3 * NB) * sizeof(uint64_t) + (size_t)(NB * NCW + NB) * 8 * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int S = p.S;
    unsigned G = (unsigned)(num_sms / S) * S;
    const unsigned need = (unsigned)p.f.g.R * S;
    if (G > need) G = need;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G);
    cfg.blockDim = dim3((NCW + DEPTH + 2) * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;      // partner CTAs wait for each other: they must be co-resident
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, p);
}

// ring depth: IHPR_FUSED_STAGES = 3 / 4 / 6 (tuning experiments; default 6)
static int fused_stages() {
    const char* e = getenv("IHPR_FUSED_STAGES");
    const int st = e ? atoi(e) : 6;
    return st == 3 || st == 4 ? st : 6;
}
template <typename T, int DEPTH>
static cudaError_t launch_fused_d(const FusedParams& p, int num_sms, cudaStream_t s) {
    if (p.f.g.CE * sizeof(T) == 65536) return launch_fused_ds<T, DEPTH, 3, 65536>(p, num_sms, s);       // IHPR_FUSED_CHUNK=64 (tuning experiments)
    switch (fused_stages()) {
        case 3: return launch_fused_ds<T, DEPTH, 3>(p, num_sms, s);
        case 4: return launch_fused_ds<T, DEPTH, 4>(p, num_sms, s);
        default: return launch_fused_ds<T, DEPTH, 6>(p, num_sms, s);
    }
}

template <typename T>
static cudaError_t launch_fused_t(const FusedParams& p, int num_sms, cudaStream_t s) {
    switch (fused_depth(std::is_same<T, float>::value ? 0 : 1)) {
        case 1: return launch_fused_d<T, 1>(p, num_sms, s);
        case 2: return launch_fused_d<T, 2>(p, num_sms, s);
        default: return launch_fused_d<T, 3>(p, num_sms, s);
    }
}

cudaError_t launch_fused(const FusedParams& p, int dtype, int num_sms, cudaStream_t s) {
    return dtype == 0 ? launch_fused_t<float>(p, num_sms, s) : launch_fused_t<__nv_bfloat16>(p, num_sms, s);
}

void launch_scale(void* grad, size_t n, int dtype, bool aligned, const float* grad_out, int num_sms, cudaStream_t s) {
    const size_t epv = dtype == 0 ? 4 : 8;
    const size_t n_vec = aligned ? n / epv : 0;
    if (dtype == 0) scale_kernel<float><<<num_sms * 2, 256, 0, s>>>(static_cast<float*>(grad), n_vec, n_vec * epv, n, grad_out);
    else scale_kernel<__nv_bfloat16><<<num_sms * 2, 256, 0, s>>>(static_cast<__nv_bfloat16*>(grad), n_vec, n_vec * epv, n, grad_out);
}

}  // namespace ihpr

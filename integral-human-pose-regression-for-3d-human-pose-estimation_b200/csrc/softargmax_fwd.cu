// softargmax_fwd.cu -- K1: single-pass online-softmax + expected-coordinate forward (sm_100a).
//
// Replaces /root/reference/common/nets/loss.py:13-34 (soft_argmax) and, when targets are given,
// loss.py:49-52 (JointLocationLoss) in ONE launch: every joint-volume is read exactly once
// (algorithmic traffic N*s bytes per volume), the softmax and the three marginals are never written.
//
// Work split: the R*nch chunks of all joint-volumes form one stream that is cut into G equal
// contiguous ranges, one per persistent CTA (G = #SMs x CTAs/SM), so every SM streams the same number
// of bytes no matter how B*J relates to 148.  A CTA that leaves a joint-volume publishes its partial
// (m, l, sx, sy, sz) to a fixed workspace slot; the last contributor (atomic ticket) merges the slots
// in slot order -- bit-reproducible -- and writes coords / stats / the loss term; the last finished
// volume reduces the loss terms in index order.
//
//   ring kernel   : warp 0 = TMA producer (cp.async.bulk into a STAGES-deep shared-memory ring,
//                   mbarrier full/empty pairs), NCW consumer warps read the ring with 128-bit LDS.
//   direct kernel : every thread issues U 128-bit streaming global loads per step (also the scalar
//                   fallback for shapes with W % 4 != 0 or unaligned bases).
#include "ihpr_common.cuh"

namespace ihpr {

// ---------------------------------------------------------------------------------------------
// row finalisation shared by all forward kernels.  Called by all `nthreads` consumer threads.
template <int NW>
__device__ __forceinline__ void flush_row(const FwdParams& p, int r, Acc a, float (*red)[8], int bar_id, int wid, int lane,
                                          uint32_t cta, uint32_t G) {
    a = acc_warp_merge(a);
    if (lane == 0) {
        red[wid][0] = a.m; red[wid][1] = a.l; red[wid][2] = a.sx; red[wid][3] = a.sy; red[wid][4] = a.sz;
    }
    named_bar_sync(bar_id, NW * 32);
    if (wid == 0) {
        Acc b;
        b.reset();
        if (lane < NW) {
            b.m = red[lane][0]; b.l = red[lane][1]; b.sx = red[lane][2]; b.sy = red[lane][3]; b.sz = red[lane][4];
            b.c = safe_c(b.m);
        }
        b = acc_warp_merge(b);

        const Geometry& g = p.g;
        const uint64_t g0 = (uint64_t)r * g.nch;
        const uint32_t c_first = owner_of(g0, g.Gt, G);
        const uint32_t c_last = owner_of(g0 + g.nch - 1, g.Gt, G);
        const int ncontrib = (int)(c_last - c_first) + 1;
        bool last = true;
        if (ncontrib > 1) {
            float* slot = p.partials + ((size_t)r * p.maxslots + (cta - c_first)) * 8;
            int ticket = 0;
            if (lane == 0) {
                __stcg(slot + 0, b.m); __stcg(slot + 1, b.l); __stcg(slot + 2, b.sx);
                __stcg(slot + 3, b.sy); __stcg(slot + 4, b.sz);
                __threadfence();
                ticket = atomicAdd(p.row_count + r, 1);
            }
            ticket = __shfl_sync(0xffffffffu, ticket, 0);
            last = (ticket == ncontrib - 1);
            if (last) {
                __threadfence();
                b.reset();
                const float* base = p.partials + (size_t)r * p.maxslots * 8;
                for (int s = lane; s < ncontrib; s += 32) {
                    Acc t;
                    t.m = __ldcg(base + s * 8 + 0); t.l = __ldcg(base + s * 8 + 1); t.sx = __ldcg(base + s * 8 + 2);
                    t.sy = __ldcg(base + s * 8 + 3); t.sz = __ldcg(base + s * 8 + 4);
                    t.c = safe_c(t.m);
                    b = acc_merge(b, t);
                }
                b = acc_warp_merge(b);
                if (lane == 0) p.row_count[r] = 0;      // leave the workspace zeroed for the next launch
            }
        }
        if (last) {
            const float inv = 1.0f / b.l;
            const float cx = b.sx * inv, cy = b.sy * inv, cz = b.sz * inv;
            int t2 = 0;
            if (lane == 0) {
                p.coords[3 * (size_t)r + 0] = cx;
                p.coords[3 * (size_t)r + 1] = cy;
                p.coords[3 * (size_t)r + 2] = cz;
                if (p.stats) { p.stats[2 * (size_t)r] = b.m; p.stats[2 * (size_t)r + 1] = b.l; }
                if (p.gt) {
                    // loss.py:49-50: (|dx| + |dy| + |dz| * have_depth) * vis / 3
                    const float v = p.vis[r], hd = p.have_depth[r / g.J];
                    const float lx = fabsf(cx - p.gt[3 * (size_t)r]) * v;
                    const float ly = fabsf(cy - p.gt[3 * (size_t)r + 1]) * v;
                    const float lz = fabsf(cz - p.gt[3 * (size_t)r + 2]) * v;
                    __stcg(p.row_loss + r, (lx + ly + lz * hd) / 3.f);
                    __threadfence();
                    t2 = atomicAdd(p.done_rows, 1);
                }
            }
            if (p.gt) {
                t2 = __shfl_sync(0xffffffffu, t2, 0);
                if (t2 == g.R - 1) {            // every joint-volume is final: loss.py:52 mean()
                    __threadfence();
                    float s = 0.f;
                    for (int i = lane; i < g.R; i += 32) s += __ldcg(p.row_loss + i);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                    if (lane == 0) { p.loss[0] = s / (float)g.R; *p.done_rows = 0; }
                }
            }
        }
    }
    named_bar_sync(bar_id, NW * 32);
}

// ---------------------------------------------------------------------------------------------
// consume `n_vec` 16-byte vectors of one chunk; LOADER(i) returns vector i of the chunk
template <typename T, int U, int NC, typename Loader>
__device__ __forceinline__ void consume_chunk(Acc& a, const Geometry& g, uint32_t n_vec, uint32_t qbase, int tid, Loader load) {
    constexpr int QPV = Elem<T>::QPV;
    const uint32_t F = g.divF.d;
    for (uint32_t base = 0; base < n_vec; base += NC * U) {
        float v[U][QPV][4];
        uint4 raw[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t iv = base + u * NC + tid;
            raw[u] = (iv < n_vec) ? load(iv) : (sizeof(T) == 4 ? make_uint4(0xff800000u, 0xff800000u, 0xff800000u, 0xff800000u)
                                                               : make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u));
        }
        float cmax = -INFINITY;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            Elem<T>::unpack(raw[u], v[u]);
#pragma unroll
            for (int q = 0; q < QPV; ++q)
                cmax = fmaxf(cmax, fmaxf(fmaxf(v[u][q][0], v[u][q][1]), fmaxf(v[u][q][2], v[u][q][3])));
        }
        if (cmax > a.m) acc_raise(a, cmax);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t iv = base + u * NC + tid;
#pragma unroll
            for (int q = 0; q < QPV; ++q) {
                const uint32_t qi = qbase + iv * QPV + q;
                const uint32_t zy = fdiv(qi, g.divF);
                const uint32_t x4 = qi - zy * F;
                const uint32_t z = fdiv(zy, g.divH);
                const uint32_t y = zy - z * g.divH.d;
                acc_quad(a, v[u][q], u2f(x4 << 2), u2f(y), u2f(z));
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// ring kernel
template <typename T, int CHUNK_BYTES, int STAGES, int NCW, int MINB>
__global__ void __launch_bounds__(NCW * 32 + 32, MINB) fwd_ring_kernel(const FwdParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* ring = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * CHUNK_BYTES);
    uint64_t* empty = full + STAGES;
    float(*red)[8] = reinterpret_cast<float(*)[8]>(empty + STAGES);

    const Geometry& g = p.g;
    const uint32_t G = gridDim.x, cta = blockIdx.x;
    const uint64_t g_lo = range_lo(g.Gt, G, cta), g_hi = range_lo(g.Gt, G, cta + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, NCW); }
        mbar_fence_init();
    }
    __syncthreads();

    uint32_t r = (uint32_t)(g_lo / g.nch);
    uint32_t k = (uint32_t)(g_lo - (uint64_t)r * g.nch);

    if (warp == 0) {
        if (lane == 0) {
            const uint64_t pol = l2_policy_evict_first();
            const uint8_t* src = reinterpret_cast<const uint8_t*>(p.heat);
            uint32_t it = 0;
            for (uint64_t gi = g_lo; gi < g_hi; ++gi, ++it) {
                const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(empty + s, ph ^ 1);
                const uint32_t e0 = k * g.CE;
                const uint32_t n_el = min(g.CE, g.N - e0);
                const uint32_t bytes = n_el * (uint32_t)sizeof(T);
                mbar_expect_tx(full + s, bytes);
                bulk_g2s(ring + (size_t)s * CHUNK_BYTES, src + ((size_t)r * g.N + e0) * sizeof(T), bytes, full + s, pol);
                if (++k == g.nch) { k = 0; ++r; }
            }
        }
        return;
    }

    // consumers
    constexpr int NC = NCW * 32;
    constexpr int VPC = CHUNK_BYTES / 16;                   // vectors per full chunk
    constexpr int U = (VPC / NC) < 1 ? 1 : ((VPC / NC) > 4 ? 4 : (VPC / NC));
    constexpr int QPV = Elem<T>::QPV;
    const int tid = threadIdx.x - 32, wid = warp - 1;
    Acc a;
    a.reset();
    uint32_t it = 0;
    for (uint64_t gi = g_lo; gi < g_hi; ++gi, ++it) {
        const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
        const uint32_t e0 = k * g.CE;
        const uint32_t n_vec = min(g.CE, g.N - e0) / (4 * QPV);
        mbar_wait(full + s, ph);
        const uint8_t* st = ring + (size_t)s * CHUNK_BYTES;
        consume_chunk<T, U, NC>(a, g, n_vec, e0 >> 2, tid, [&](uint32_t iv) { return lds16(st + (size_t)iv * 16); });
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
        if (++k == g.nch) {
            flush_row<NCW>(p, (int)r, a, red, 1, wid, lane, cta, G);
            a.reset();
            k = 0; ++r;
        }
    }
    if (k != 0) flush_row<NCW>(p, (int)r, a, red, 1, wid, lane, cta, G);
}

// ---------------------------------------------------------------------------------------------
// direct kernel (vector): NT threads, each U 16-byte streaming loads per step
template <typename T, int NT, int U, int MINB>
__global__ void __launch_bounds__(NT, MINB) fwd_direct_kernel(const FwdParams p) {
    __shared__ float red[NT / 32][8];
    const Geometry& g = p.g;
    const uint32_t G = gridDim.x, cta = blockIdx.x;
    const uint64_t g_lo = range_lo(g.Gt, G, cta), g_hi = range_lo(g.Gt, G, cta + 1);
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    constexpr int QPV = Elem<T>::QPV;
    uint32_t r = (uint32_t)(g_lo / g.nch);
    uint32_t k = (uint32_t)(g_lo - (uint64_t)r * g.nch);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(p.heat);
    Acc a;
    a.reset();
    for (uint64_t gi = g_lo; gi < g_hi; ++gi) {
        const uint32_t e0 = k * g.CE;
        const uint32_t n_vec = min(g.CE, g.N - e0) / (4 * QPV);
        const uint8_t* cp = src + ((size_t)r * g.N + e0) * sizeof(T);
        consume_chunk<T, U, NT>(a, g, n_vec, e0 >> 2, tid, [&](uint32_t iv) { return ld_stream16(cp + (size_t)iv * 16); });
        if (++k == g.nch) {
            flush_row<NT / 32>(p, (int)r, a, red, 1, wid, lane, cta, G);
            a.reset();
            k = 0; ++r;
        }
    }
    if (k != 0) flush_row<NT / 32>(p, (int)r, a, red, 1, wid, lane, cta, G);
}

// scalar fallback: any shape / alignment (W % 4 != 0, odd N, unaligned base pointer)
template <typename T, int NT>
__global__ void __launch_bounds__(NT) fwd_scalar_kernel(const FwdParams p) {
    __shared__ float red[NT / 32][8];
    const Geometry& g = p.g;
    const uint32_t G = gridDim.x, cta = blockIdx.x;
    const uint64_t g_lo = range_lo(g.Gt, G, cta), g_hi = range_lo(g.Gt, G, cta + 1);
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t r = (uint32_t)(g_lo / g.nch);
    uint32_t k = (uint32_t)(g_lo - (uint64_t)r * g.nch);
    const T* src = reinterpret_cast<const T*>(p.heat);
    Acc a;
    a.reset();
    for (uint64_t gi = g_lo; gi < g_hi; ++gi) {
        const uint32_t e0 = k * g.CE;
        const uint32_t n_el = min(g.CE, g.N - e0);
        const T* cp = src + (size_t)r * g.N + e0;
        for (uint32_t i = threadIdx.x; i < n_el; i += NT) {
            const float h = Elem<T>::load1(cp + i);
            if (h > a.m) acc_raise(a, h);
            const uint32_t e = e0 + i;
            const uint32_t zy = fdiv(e, g.divW);
            const uint32_t x = e - zy * g.divW.d;
            const uint32_t z = fdiv(zy, g.divH);
            const uint32_t y = zy - z * g.divH.d;
            const float pw = ex2(fmaf(h, kLog2e, -a.c));
            a.l += pw;
            a.sx = fmaf(pw, u2f(x), a.sx);
            a.sy = fmaf(pw, u2f(y), a.sy);
            a.sz = fmaf(pw, u2f(z), a.sz);
        }
        if (++k == g.nch) {
            flush_row<NT / 32>(p, (int)r, a, red, 1, wid, lane, cta, G);
            a.reset();
            k = 0; ++r;
        }
    }
    if (k != 0) flush_row<NT / 32>(p, (int)r, a, red, 1, wid, lane, cta, G);
}

// ---------------------------------------------------------------------------------------------
template <typename T, int CHUNK_BYTES, int STAGES, int NCW, int MINB>
static void launch_ring(const FwdParams& p, int num_sms, cudaStream_t s) {
    auto kern = fwd_ring_kernel<T, CHUNK_BYTES, STAGES, NCW, MINB>;
    const size_t smem = (size_t)STAGES * CHUNK_BYTES + 2 * STAGES * sizeof(uint64_t) + NCW * 8 * sizeof(float);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    uint64_t G = (uint64_t)num_sms * MINB;
    if (G > p.g.Gt) G = p.g.Gt;
    if (G > kGridCap) G = kGridCap;
    kern<<<(unsigned)G, NCW * 32 + 32, smem, s>>>(p);
}
template <typename T, int NT, int U, int MINB>
static void launch_direct(const FwdParams& p, int num_sms, cudaStream_t s) {
    uint64_t G = (uint64_t)num_sms * MINB;
    if (G > p.g.Gt) G = p.g.Gt;
    if (G > kGridCap) G = kGridCap;
    fwd_direct_kernel<T, NT, U, MINB><<<(unsigned)G, NT, 0, s>>>(p);
}
template <typename T>
static void launch_scalar(const FwdParams& p, int num_sms, cudaStream_t s) {
    uint64_t G = (uint64_t)num_sms * 4;
    if (G > p.g.Gt) G = p.g.Gt;
    if (G > kGridCap) G = kGridCap;
    fwd_scalar_kernel<T, 256><<<(unsigned)G, 256, 0, s>>>(p);
}

// chunk size (voxels) used by (dtype, vec_ok, variant); must match the launch table below
static uint32_t chunk_bytes_of(bool vec_ok, int variant) {
    if (!vec_ok) return 0;
    switch (variant) {
        case 2: return 32768;       // direct: 512 threads x 4 x 16 B
        case 11: return 32768;
        case 21: return 16384;      // direct: 256 threads x 4 x 16 B
        default: return 16384;
    }
}

Geometry make_geometry(int B, int J, int D, int H, int W, int dtype, bool vec_ok, int variant) {
    Geometry g;
    g.R = B * J; g.J = J; g.D = D; g.H = H; g.W = W;
    g.N = (uint32_t)D * H * W;
    const uint32_t es = dtype == 0 ? 4 : 2;
    const uint32_t cb = chunk_bytes_of(vec_ok, variant);
    g.CE = cb ? cb / es : (uint32_t)kMinChunkElems;
    g.nch = (g.N + g.CE - 1) / g.CE;
    g.Gt = (uint64_t)g.R * g.nch;
    g.divF = make_fastdiv(vec_ok ? (uint32_t)W / 4 : 1);
    g.divW = make_fastdiv((uint32_t)W);
    g.divH = make_fastdiv((uint32_t)H);
    return g;
}

template <typename T>
static void launch_fwd_t(const FwdParams& p, bool vec_ok, int variant, int num_sms, cudaStream_t s) {
    if (!vec_ok) return launch_scalar<T>(p, num_sms, s);
    switch (variant) {
        case 2: return launch_direct<T, 512, 4, 2>(p, num_sms, s);
        case 21: return launch_direct<T, 256, 4, 4>(p, num_sms, s);
        case 11: return launch_ring<T, 32768, 6, 16, 1>(p, num_sms, s);
        case 12: return launch_ring<T, 16384, 12, 16, 1>(p, num_sms, s);
        case 13: return launch_ring<T, 16384, 6, 8, 2>(p, num_sms, s);
        default: return launch_ring<T, 16384, 12, 8, 1>(p, num_sms, s);
    }
}

void launch_fwd(const FwdParams& p, int dtype, bool vec_ok, int variant, int num_sms, cudaStream_t s) {
    if (dtype == 0) launch_fwd_t<float>(p, vec_ok, variant, num_sms, s);
    else launch_fwd_t<__nv_bfloat16>(p, vec_ok, variant, num_sms, s);
}

}  // namespace ihpr

// ihpr_device.cuh -- device-side building blocks shared by K1 (forward), K2 (backward) and K5 (fused
// forward+backward): joint-volume finalisation, chunk consumers (generic and fast path), per-volume
// backward constants.  See softargmax_fwd.cu / softargmax_bwd.cu / softargmax_fused.cu for the kernels.
#pragma once
#include "ihpr_common.cuh"

namespace ihpr {

// The merged partial `b` of joint-volume r is complete (held by every lane of ONE warp): coords / stats / loss term, and the
// loss.py:52 mean once the last joint-volume of the launch has been finalised.
__device__ __forceinline__ void finalize_row(const FwdParams& p, int r, const Acc& b, int lane) {
    const Geometry& g = p.g;
    const float inv = 1.0f / b.l;
    const float cx = b.sx * inv, cy = b.sy * inv, cz = b.sz * inv;
    int t2 = 0;
    if (lane == 0) {
        p.coords[3 * (size_t)r + 0] = cx;
        p.coords[3 * (size_t)r + 1] = cy;
        p.coords[3 * (size_t)r + 2] = cz;
        if (p.stats) {              // re-base (reference point, l) to the true maximum
            const float f = (b.m == -INFINITY) ? 0.f : ex2(b.c - safe_c(b.mx));
            p.stats[2 * (size_t)r] = b.mx;
            p.stats[2 * (size_t)r + 1] = b.l * f;
        }
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
        if (t2 == g.R - 1) {            // every joint-volume is final: loss.py:52 mean(), index order
            __threadfence();
            float s = 0.f;
            for (int i = lane; i < g.R; i += 32) s += __ldcg(p.row_loss + i);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) { p.loss[0] = s / (float)g.R; *p.done_rows = 0; }
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Joint-volume finalisation.
//
// publish_row: called by ONE warp that holds this CTA's partial `b` of joint-volume r.  If other CTAs
// also hold pieces of r, the partial goes to this CTA's fixed workspace slot and the last arriver
// (atomic ticket) merges all slots in slot order; the winner writes coords / stats / loss term.
__device__ __forceinline__ void publish_row(const FwdParams& p, int r, Acc b, int lane, uint32_t cta, uint32_t G) {
    const Geometry& g = p.g;
    const uint64_t g0 = (uint64_t)r * g.nch;
    const uint32_t c_first = owner_of(g0, g.Gt, G);
    const uint32_t c_last = owner_of(g0 + g.nch - 1, g.Gt, G);
    const int ncontrib = (int)(c_last - c_first) + 1;
    bool last = true;
    if (ncontrib > 1) {
        int ticket = 0;
        if (lane == 0) {
            partial_to_global(p.partials + ((size_t)r * p.maxslots + (cta - c_first)) * 8, b);
            __threadfence();
            ticket = atomicAdd(p.row_count + r, 1);
        }
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        last = (ticket == ncontrib - 1);
        if (last) {
            __threadfence();
            b.reset();
            const float* base = p.partials + (size_t)r * p.maxslots * 8;
            for (int s = lane; s < ncontrib; s += 32) b = acc_merge(b, partial_from_global(base + s * 8));
            b = acc_warp_merge(b);
            if (lane == 0) p.row_count[r] = 0;      // leave the workspace zeroed for the next launch
        }
    }
    if (!last) return;
    finalize_row(p, r, b, lane);
}

// Barrier flavour (direct / scalar kernels): all NW warps meet, warp 0 merges and publishes.
template <int NW>
__device__ __forceinline__ void flush_row(const FwdParams& p, int r, Acc a, float (*red)[8], int bar_id, int wid, int lane,
                                          uint32_t cta, uint32_t G) {
    a = acc_warp_merge(a);
    if (lane == 0) partial_to_smem(red[wid], a);
    named_bar_sync(bar_id, NW * 32);
    if (wid == 0) {
        Acc b;
        b.reset();
        if (lane < NW) b = partial_from_smem(red[lane]);
        b = acc_warp_merge(b);
        publish_row(p, r, b, lane, cta, G);
    }
    named_bar_sync(bar_id, NW * 32);
}

// Barrier-free flavour (ring kernel): a warp drops its partial into the shared-memory buffer of this
// joint-volume and moves on; whichever warp arrives last merges the NW partials and publishes.  Warps of
// one CTA are at most STAGES chunks apart (the ring holds them together), hence at most STAGES
// joint-volumes apart: NBUF = STAGES + 1 buffers indexed by the CTA-local volume sequence number suffice.
template <int NW, int NBUF>
__device__ __forceinline__ void flush_row_async(const FwdParams& p, int r, Acc a, float (*pbuf)[8], int* pcnt, uint32_t rowseq, int wid,
                                                int lane, uint32_t cta, uint32_t G) {
    a = acc_warp_merge(a);
    const uint32_t b = rowseq % NBUF;
    int old = 0;
    if (lane == 0) {
        partial_to_smem(pbuf[b * NW + wid], a);
        __threadfence_block();
        old = atomicAdd(pcnt + b, 1);
    }
    old = __shfl_sync(0xffffffffu, old, 0);
    if (old == NW - 1) {
        __threadfence_block();
        Acc t;
        t.reset();
        if (lane < NW) t = partial_from_smem(pbuf[b * NW + lane]);
        t = acc_warp_merge(t);
        if (lane == 0) pcnt[b] = 0;
        publish_row(p, r, t, lane, cta, G);
    }
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
        acc_see_max(a, cmax);
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

// Fast path.  When Fv = 16-byte vectors per x-row divides both the consumer count NC and the chunk's vector
// count, a thread's x position never changes and its (y, z) advance by RS = NC / Fv rows per step (RS <= H, one
// wrap at most): no division and no int->float conversion in the loop.  a.sx only collects the within-vector
// x moment; the thread-constant x0 * l is added once per joint-volume (fast_fix) before the merge.
template <int NC, int VPC>
__device__ __forceinline__ bool fast_ok(const Geometry& g) {
    const uint32_t Fv = g.divFv.d;
    return Fv != 0 && (NC % Fv) == 0 && (VPC % Fv) == 0 && (uint32_t)NC / Fv <= (uint32_t)g.H;
}

template <typename T, int U, int NC, int VPC, bool FULL, typename Loader>
__device__ __forceinline__ void consume_chunk_fast(Acc& a, const Geometry& g, uint32_t n_vec, uint32_t vbase, int tid, float rsf, float hf,
                                                   Loader load) {
    constexpr int QPV = Elem<T>::QPV;
    const uint32_t zy0 = fdiv(vbase + tid, g.divFv);
    const uint32_t z0 = fdiv(zy0, g.divH);
    float yf = u2f(zy0 - z0 * g.divH.d), zf = u2f(z0);
    const uint32_t nv = FULL ? (uint32_t)VPC : n_vec;
#pragma unroll
    for (uint32_t base = 0; base < nv; base += NC * U) {
        float v[U][QPV][4];
        uint4 raw[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t iv = base + u * NC + tid;
            if (FULL || iv < nv) raw[u] = load(iv);
            else raw[u] = (sizeof(T) == 4 ? make_uint4(0xff800000u, 0xff800000u, 0xff800000u, 0xff800000u)
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
        acc_see_max(a, cmax);
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int q = 0; q < QPV; ++q) {
                const float p0 = ex2(fmaf(v[u][q][0], kLog2e, -a.c));
                const float p1 = ex2(fmaf(v[u][q][1], kLog2e, -a.c));
                const float p2 = ex2(fmaf(v[u][q][2], kLog2e, -a.c));
                const float p3 = ex2(fmaf(v[u][q][3], kLog2e, -a.c));
                const float s = (p0 + p1) + (p2 + p3);
                const float w = fmaf(p3, 3.f, fmaf(p2, 2.f, p1));
                a.l += s;
                a.sx = (q == 0) ? (a.sx + w) : fmaf(4.f * q, s, a.sx + w);
                a.sy = fmaf(s, yf, a.sy);
                a.sz = fmaf(s, zf, a.sz);
            }
            yf += rsf;
            if (yf >= hf) { yf -= hf; zf += 1.f; }
        }
    }
}

// Packed form of the fast path (K1): the issue slots are the limiter once the input is bf16 (ncu, profiles/r02_ncu_K1_bf16.txt: 67 % issue
// utilisation at 9.3 instructions per voxel, MUFU 56 %), so the per-voxel work runs on Blackwell's packed fp32 pairs (FFMA2 / FADD2)
// and the bookkeeping moves from the quad to the 16-byte vector: per vector one sum S of its weights (-> l, y, z moments) and one FFMA2
// per pair into a packed accumulator of the within-vector x moment, which lives OUTSIDE `Acc` (w2: even-k moments in the low half,
// odd-k in the high half) and is folded into a.sx before a re-base and when the joint-volume is flushed.
__device__ __forceinline__ uint64_t pk2f(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ uint64_t pk2w(uint32_t a, uint32_t b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ void up2f(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2p(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t fadd2p(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint32_t bmax2(uint32_t a, uint32_t b) {      // per-half maximum of two packed bf16 pairs
    uint32_t r;
    asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ void fold_w2(Acc& a, uint64_t& w2) {
    float lo, hi;
    up2f(w2, lo, hi);
    a.sx += lo + hi;
    w2 = 0;             // two +0.0f
}

// the arithmetic of consume_chunk_fast_pk on U vectors that are already in registers (yf / zf: the row / slice of the thread's first vector,
// advanced here by rsf rows per vector)
template <typename T, int U>
__device__ __forceinline__ void consume_vectors_pk(Acc& a, uint64_t& w2, const uint4 (&raw)[U], float& yf, float& zf, float rsf, float hf) {
    constexpr int QPV = Elem<T>::QPV;
    constexpr int NP = 2 * QPV;             // fp32 pairs per 16-byte vector
    const uint64_t l2e2 = pk2f(kLog2e, kLog2e);
    float cmax;
    if (sizeof(T) == 2) {               // maximum on the packed bf16 words: 4 instructions per vector instead of 8
        uint32_t m = bmax2(bmax2(raw[0].x, raw[0].y), bmax2(raw[0].z, raw[0].w));
#pragma unroll
        for (int u = 1; u < U; ++u) m = bmax2(m, bmax2(bmax2(raw[u].x, raw[u].y), bmax2(raw[u].z, raw[u].w)));
        cmax = fmaxf(__uint_as_float(m << 16), __uint_as_float(m & 0xffff0000u));
    } else {
        cmax = -INFINITY;
#pragma unroll
        for (int u = 0; u < U; ++u)
            cmax = fmaxf(cmax, fmaxf(fmaxf(__uint_as_float(raw[u].x), __uint_as_float(raw[u].y)),
                                     fmaxf(__uint_as_float(raw[u].z), __uint_as_float(raw[u].w))));
    }
    a.mx = fmaxf(a.mx, cmax);
    if (cmax > a.lim) {
        fold_w2(a, w2);
        acc_raise(a, cmax);
    }
    const uint64_t nc2 = pk2f(-a.c, -a.c);
#pragma unroll
    for (int u = 0; u < U; ++u) {
        uint64_t pp[NP];
        if (sizeof(T) == 2) {           // bf16 -> fp32 is a shift / a mask, written straight into the halves of a pair
            const uint32_t w[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) pp[i % NP] = pk2w(w[i] << 16, w[i] & 0xffff0000u);
        } else {
            pp[0] = pk2w(raw[u].x, raw[u].y);
            pp[1] = pk2w(raw[u].z, raw[u].w);
        }
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            float t0, t1;
            up2f(ffma2p(pp[i], l2e2, nc2), t0, t1);
            pp[i] = pk2f(ex2(t0), ex2(t1));
            w2 = ffma2p(pp[i], pk2f((float)(2 * i), (float)(2 * i + 1)), w2);      // within-vector x moment, even / odd halves
        }
        uint64_t s2 = fadd2p(pp[0], pp[1]);
        if (NP == 4) s2 = fadd2p(s2, fadd2p(pp[2], pp[3]));
        float s_lo, s_hi;
        up2f(s2, s_lo, s_hi);
        const float sv = s_lo + s_hi;
        a.l += sv;
        a.sy = fmaf(sv, yf, a.sy);
        a.sz = fmaf(sv, zf, a.sz);
        yf += rsf;
        if (yf >= hf) { yf -= hf; zf += 1.f; }
    }
}

template <typename T, int U, int NC, int VPC, bool FULL, typename Loader>
__device__ __forceinline__ void consume_chunk_fast_pk(Acc& a, uint64_t& w2, const Geometry& g, uint32_t n_vec, uint32_t vbase, int tid, float rsf,
                                                      float hf, Loader load) {
    const uint32_t zy0 = fdiv(vbase + tid, g.divFv);
    const uint32_t z0 = fdiv(zy0, g.divH);
    float yf = u2f(zy0 - z0 * g.divH.d), zf = u2f(z0);
    const uint32_t nv = FULL ? (uint32_t)VPC : n_vec;
#pragma unroll
    for (uint32_t base = 0; base < nv; base += NC * U) {
        uint4 raw[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t iv = base + u * NC + tid;
            if (FULL || iv < nv) raw[u] = load(iv);
            else raw[u] = (sizeof(T) == 4 ? make_uint4(0xff800000u, 0xff800000u, 0xff800000u, 0xff800000u)
                                          : make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u));
        }
        consume_vectors_pk<T, U>(a, w2, raw, yf, zf, rsf, hf);
    }
}

// =============================================================================================
// backward pieces
struct RowK {
    float c;            // safe m*log2e
    float gx, gy, gz;   // upstream gradient, pre-divided by l
    float cx, cy, cz;   // expected coordinate
};

__device__ __forceinline__ float sgn(float d) { return (float)((d > 0.f) - (d < 0.f)); }

__device__ __forceinline__ RowK load_row(const BwdParams& p, uint32_t r) {
    RowK k;
    const float m = __ldg(p.stats + 2 * (size_t)r), l = __ldg(p.stats + 2 * (size_t)r + 1);
    k.c = safe_c(m);
    k.cx = __ldg(p.coords + 3 * (size_t)r);
    k.cy = __ldg(p.coords + 3 * (size_t)r + 1);
    k.cz = __ldg(p.coords + 3 * (size_t)r + 2);
    float gx, gy, gz;
    if (p.grad_coords) {
        gx = __ldg(p.grad_coords + 3 * (size_t)r);
        gy = __ldg(p.grad_coords + 3 * (size_t)r + 1);
        gz = __ldg(p.grad_coords + 3 * (size_t)r + 2);
    } else {
        const float s = (p.grad_out ? __ldg(p.grad_out) : p.grad_out_const) * __ldg(p.vis + r) * p.loss_scale;
        gx = s * sgn(k.cx - __ldg(p.gt + 3 * (size_t)r));
        gy = s * sgn(k.cy - __ldg(p.gt + 3 * (size_t)r + 1));
        gz = s * sgn(k.cz - __ldg(p.gt + 3 * (size_t)r + 2)) * __ldg(p.have_depth + r / p.g.J);
    }
    const float il = 1.0f / l;
    k.gx = gx * il; k.gy = gy * il; k.gz = gz * il;
    return k;
}

__device__ __forceinline__ void bwd_quad(const RowK& k, const float (&v)[4], float (&o)[4], float xf, float yf, float zf) {
    const float base = fmaf(k.gz, zf - k.cz, fmaf(k.gy, yf - k.cy, k.gx * (xf - k.cx)));
    o[0] = ex2(fmaf(v[0], kLog2e, -k.c)) * base;
    o[1] = ex2(fmaf(v[1], kLog2e, -k.c)) * (base + k.gx);
    o[2] = ex2(fmaf(v[2], kLog2e, -k.c)) * fmaf(2.f, k.gx, base);
    o[3] = ex2(fmaf(v[3], kLog2e, -k.c)) * fmaf(3.f, k.gx, base);
}

template <typename T, int U, int NC, typename Loader>
__device__ __forceinline__ void bwd_chunk(const RowK& k, const Geometry& g, uint32_t n_vec, uint32_t qbase, int tid, uint8_t* dst,
                                          Loader load) {
    constexpr int QPV = Elem<T>::QPV;
    const uint32_t F = g.divF.d;
    for (uint32_t base = 0; base < n_vec; base += NC * U) {
        uint4 raw[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t iv = base + u * NC + tid;
            if (iv < n_vec) raw[u] = load(iv);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t iv = base + u * NC + tid;
            if (iv < n_vec) {
                float v[QPV][4], o[QPV][4];
                Elem<T>::unpack(raw[u], v);
#pragma unroll
                for (int q = 0; q < QPV; ++q) {
                    const uint32_t qi = qbase + iv * QPV + q;
                    const uint32_t zy = fdiv(qi, g.divF);
                    const uint32_t x4 = qi - zy * F;
                    const uint32_t z = fdiv(zy, g.divH);
                    const uint32_t y = zy - z * g.divH.d;
                    bwd_quad(k, v[q], o[q], u2f(x4 << 2), u2f(y), u2f(z));
                }
                st_stream16(dst + (size_t)iv * 16, Elem<T>::pack(o));
            }
        }
    }
}

// Fast path (same preconditions as the forward's): x is a thread constant, so gx * (x - cx) for the vector's
// voxels is hoisted into tx[] per joint-volume; (y, z) advance by RS rows per step without division.
template <int E>
__device__ __forceinline__ void make_tx(const RowK& k, float x0f, float (&tx)[E]) {
#pragma unroll
    for (int e = 0; e < E; ++e) tx[e] = k.gx * ((x0f + (float)e) - k.cx);
}

template <typename T, int U, int NC, int VPC, bool FULL, typename Loader>
__device__ __forceinline__ void bwd_chunk_fast(const RowK& k, const float (&tx)[4 * Elem<T>::QPV], const Geometry& g, uint32_t n_vec,
                                               uint32_t vbase, int tid, float rsf, float hf, uint8_t* dst, Loader load) {
    constexpr int QPV = Elem<T>::QPV;
    const uint32_t zy0 = fdiv(vbase + tid, g.divFv);
    const uint32_t z0 = fdiv(zy0, g.divH);
    float yf = u2f(zy0 - z0 * g.divH.d), zf = u2f(z0);
    const uint32_t nv = FULL ? (uint32_t)VPC : n_vec;
#pragma unroll
    for (uint32_t base = 0; base < nv; base += NC * U) {
        uint4 raw[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t iv = base + u * NC + tid;
            if (FULL || iv < nv) raw[u] = load(iv);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t iv = base + u * NC + tid;
            float v[QPV][4], o[QPV][4];
            Elem<T>::unpack(raw[u], v);
            const float t = fmaf(k.gy, yf - k.cy, k.gz * (zf - k.cz));
#pragma unroll
            for (int q = 0; q < QPV; ++q)
#pragma unroll
                for (int e = 0; e < 4; ++e) o[q][e] = ex2(fmaf(v[q][e], kLog2e, -k.c)) * (t + tx[4 * q + e]);
            if (FULL || iv < nv) st_stream16(dst + (size_t)iv * 16, Elem<T>::pack(o));
            yf += rsf;
            if (yf >= hf) { yf -= hf; zf += 1.f; }
        }
    }
}

}  // namespace ihpr

// softargmax_bwd.cu -- K2: recompute-in-backward soft-argmax gradient (sm_100a).
//
// The reference gets this from autograd (loss.backward(), /root/reference/main/train.py:71) through
// softmax-backward + three expand-adds on a saved, V-sized softmax output.  Here nothing V-sized is
// saved: p_i = 2^(h_i*log2e - c) / l is recomputed from heat and the two per-volume scalars the
// forward kept, and
//     grad_heat_i = p_i * sum_c g_c * (c(i) - coord_c)
// streams out: read V, write V (algorithmic 2*N*s bytes per joint-volume).  With the loss fused,
//     g_c = grad_out * sign(coord_c - gt_c) * vis * w_c / (3*B*J)      (loss.py:49-52)
// is formed per joint-volume in the kernel prologue instead of by ~10 tiny launches.
// Same persistent equal-bytes split as the forward; no cross-CTA communication at all.
#include "ihpr_common.cuh"

namespace ihpr {

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
        const float s = __ldg(p.grad_out) * __ldg(p.vis + r) * p.loss_scale;
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

template <typename T, int CHUNK_BYTES, int STAGES, int NCW, int MINB>
__global__ void __launch_bounds__(NCW * 32 + 32, MINB) bwd_ring_kernel(const BwdParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* ring = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * CHUNK_BYTES);
    uint64_t* empty = full + STAGES;

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
                const uint32_t bytes = min(g.CE, g.N - e0) * (uint32_t)sizeof(T);
                mbar_expect_tx(full + s, bytes);
                bulk_g2s(ring + (size_t)s * CHUNK_BYTES, src + ((size_t)r * g.N + e0) * sizeof(T), bytes, full + s, pol);
                if (++k == g.nch) { k = 0; ++r; }
            }
        }
        return;
    }

    constexpr int NC = NCW * 32;
    constexpr int VPC = CHUNK_BYTES / 16;
    constexpr int U = (VPC / NC) < 1 ? 1 : ((VPC / NC) > 4 ? 4 : (VPC / NC));
    constexpr int QPV = Elem<T>::QPV;
    const int tid = threadIdx.x - 32;
    uint8_t* out = reinterpret_cast<uint8_t*>(p.grad_heat);
    RowK rk = load_row(p, r);
    uint32_t it = 0;
    for (uint64_t gi = g_lo; gi < g_hi; ++gi, ++it) {
        const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
        const uint32_t e0 = k * g.CE;
        const uint32_t n_vec = min(g.CE, g.N - e0) / (4 * QPV);
        mbar_wait(full + s, ph);
        const uint8_t* st = ring + (size_t)s * CHUNK_BYTES;
        bwd_chunk<T, U, NC>(rk, g, n_vec, e0 >> 2, tid, out + ((size_t)r * g.N + e0) * sizeof(T),
                            [&](uint32_t iv) { return lds16(st + (size_t)iv * 16); });
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
        if (++k == g.nch) {
            k = 0; ++r;
            if (gi + 1 < g_hi) rk = load_row(p, r);
        }
    }
}

template <typename T, int NT, int U, int MINB>
__global__ void __launch_bounds__(NT, MINB) bwd_direct_kernel(const BwdParams p) {
    const Geometry& g = p.g;
    const uint32_t G = gridDim.x, cta = blockIdx.x;
    const uint64_t g_lo = range_lo(g.Gt, G, cta), g_hi = range_lo(g.Gt, G, cta + 1);
    constexpr int QPV = Elem<T>::QPV;
    uint32_t r = (uint32_t)(g_lo / g.nch);
    uint32_t k = (uint32_t)(g_lo - (uint64_t)r * g.nch);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(p.heat);
    uint8_t* out = reinterpret_cast<uint8_t*>(p.grad_heat);
    RowK rk = load_row(p, r);
    for (uint64_t gi = g_lo; gi < g_hi; ++gi) {
        const uint32_t e0 = k * g.CE;
        const uint32_t n_vec = min(g.CE, g.N - e0) / (4 * QPV);
        const size_t off = ((size_t)r * g.N + e0) * sizeof(T);
        const uint8_t* cp = src + off;
        bwd_chunk<T, U, NT>(rk, g, n_vec, e0 >> 2, threadIdx.x, out + off, [&](uint32_t iv) { return ld_stream16(cp + (size_t)iv * 16); });
        if (++k == g.nch) {
            k = 0; ++r;
            if (gi + 1 < g_hi) rk = load_row(p, r);
        }
    }
}

template <typename T, int NT>
__global__ void __launch_bounds__(NT) bwd_scalar_kernel(const BwdParams p) {
    const Geometry& g = p.g;
    const uint32_t G = gridDim.x, cta = blockIdx.x;
    const uint64_t g_lo = range_lo(g.Gt, G, cta), g_hi = range_lo(g.Gt, G, cta + 1);
    uint32_t r = (uint32_t)(g_lo / g.nch);
    uint32_t k = (uint32_t)(g_lo - (uint64_t)r * g.nch);
    const T* src = reinterpret_cast<const T*>(p.heat);
    T* out = reinterpret_cast<T*>(p.grad_heat);
    RowK rk = load_row(p, r);
    for (uint64_t gi = g_lo; gi < g_hi; ++gi) {
        const uint32_t e0 = k * g.CE;
        const uint32_t n_el = min(g.CE, g.N - e0);
        const size_t off = (size_t)r * g.N + e0;
        for (uint32_t i = threadIdx.x; i < n_el; i += NT) {
            const float h = Elem<T>::load1(src + off + i);
            const uint32_t e = e0 + i;
            const uint32_t zy = fdiv(e, g.divW);
            const uint32_t x = e - zy * g.divW.d;
            const uint32_t z = fdiv(zy, g.divH);
            const uint32_t y = zy - z * g.divH.d;
            const float t = fmaf(rk.gz, u2f(z) - rk.cz, fmaf(rk.gy, u2f(y) - rk.cy, rk.gx * (u2f(x) - rk.cx)));
            Elem<T>::store1(out + off + i, ex2(fmaf(h, kLog2e, -rk.c)) * t);
        }
        if (++k == g.nch) {
            k = 0; ++r;
            if (gi + 1 < g_hi) rk = load_row(p, r);
        }
    }
}

template <typename T, int CHUNK_BYTES, int STAGES, int NCW, int MINB>
static void launch_ring(const BwdParams& p, int num_sms, cudaStream_t s) {
    auto kern = bwd_ring_kernel<T, CHUNK_BYTES, STAGES, NCW, MINB>;
    const size_t smem = (size_t)STAGES * CHUNK_BYTES + 2 * STAGES * sizeof(uint64_t);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    uint64_t G = (uint64_t)num_sms * MINB;
    if (G > p.g.Gt) G = p.g.Gt;
    if (G > kGridCap) G = kGridCap;
    kern<<<(unsigned)G, NCW * 32 + 32, smem, s>>>(p);
}
template <typename T, int NT, int U, int MINB>
static void launch_direct(const BwdParams& p, int num_sms, cudaStream_t s) {
    uint64_t G = (uint64_t)num_sms * MINB;
    if (G > p.g.Gt) G = p.g.Gt;
    if (G > kGridCap) G = kGridCap;
    bwd_direct_kernel<T, NT, U, MINB><<<(unsigned)G, NT, 0, s>>>(p);
}
template <typename T>
static void launch_scalar(const BwdParams& p, int num_sms, cudaStream_t s) {
    uint64_t G = (uint64_t)num_sms * 4;
    if (G > p.g.Gt) G = p.g.Gt;
    if (G > kGridCap) G = kGridCap;
    bwd_scalar_kernel<T, 256><<<(unsigned)G, 256, 0, s>>>(p);
}

template <typename T>
static void launch_bwd_t(const BwdParams& p, bool vec_ok, int variant, int num_sms, cudaStream_t s) {
    if (!vec_ok) return launch_scalar<T>(p, num_sms, s);
    switch (variant) {      // chunk sizes must match chunk_bytes_of() in softargmax_fwd.cu
        case 2: return launch_direct<T, 512, 4, 2>(p, num_sms, s);
        case 21: return launch_direct<T, 256, 4, 4>(p, num_sms, s);
        case 11: return launch_ring<T, 32768, 6, 16, 1>(p, num_sms, s);
        case 12: return launch_ring<T, 16384, 12, 16, 1>(p, num_sms, s);
        case 13: return launch_ring<T, 16384, 6, 8, 2>(p, num_sms, s);
        default: return launch_ring<T, 16384, 12, 8, 1>(p, num_sms, s);
    }
}

void launch_bwd(const BwdParams& p, int dtype, bool vec_ok, int variant, int num_sms, cudaStream_t s) {
    if (dtype == 0) launch_bwd_t<float>(p, vec_ok, variant, num_sms, s);
    else launch_bwd_t<__nv_bfloat16>(p, vec_ok, variant, num_sms, s);
}

}  // namespace ihpr

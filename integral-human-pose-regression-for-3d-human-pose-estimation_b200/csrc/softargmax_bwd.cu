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
#include "ihpr_device.cuh"

namespace ihpr {

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
    const bool fast = fast_ok<NC, VPC>(g);
    const uint32_t Fv = fast ? g.divFv.d : 1;
    const float x0f = u2f((uint32_t)(tid % Fv) * (4 * QPV)), rsf = u2f(NC / Fv), hf = u2f((uint32_t)g.H);
    RowK rk = load_row(p, r);
    float tx[4 * QPV];
    make_tx<4 * QPV>(rk, x0f, tx);
    uint32_t it = 0;
    for (uint64_t gi = g_lo; gi < g_hi; ++gi, ++it) {
        const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
        const uint32_t e0 = k * g.CE;
        const uint32_t n_vec = min(g.CE, g.N - e0) / (4 * QPV);
        mbar_wait(full + s, ph);
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
        if (++k == g.nch) {
            k = 0; ++r;
            if (gi + 1 < g_hi) { rk = load_row(p, r); make_tx<4 * QPV>(rk, x0f, tx); }
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
    constexpr int VPC = NT * U;
    const int tid = threadIdx.x;
    const bool fast = fast_ok<NT, VPC>(g);
    const uint32_t Fv = fast ? g.divFv.d : 1;
    const float x0f = u2f((uint32_t)(tid % Fv) * (4 * QPV)), rsf = u2f(NT / Fv), hf = u2f((uint32_t)g.H);
    RowK rk = load_row(p, r);
    float tx[4 * QPV];
    make_tx<4 * QPV>(rk, x0f, tx);
    for (uint64_t gi = g_lo; gi < g_hi; ++gi) {
        const uint32_t e0 = k * g.CE;
        const uint32_t n_vec = min(g.CE, g.N - e0) / (4 * QPV);
        const size_t off = ((size_t)r * g.N + e0) * sizeof(T);
        const uint8_t* cp = src + off;
        auto load = [&](uint32_t iv) { return ld_stream16(cp + (size_t)iv * 16); };
        if (fast) {
            if (n_vec == VPC) bwd_chunk_fast<T, U, NT, VPC, true>(rk, tx, g, n_vec, k * VPC, tid, rsf, hf, out + off, load);
            else bwd_chunk_fast<T, U, NT, VPC, false>(rk, tx, g, n_vec, k * VPC, tid, rsf, hf, out + off, load);
        } else {
            bwd_chunk<T, U, NT>(rk, g, n_vec, e0 >> 2, tid, out + off, load);
        }
        if (++k == g.nch) {
            k = 0; ++r;
            if (gi + 1 < g_hi) { rk = load_row(p, r); make_tx<4 * QPV>(rk, x0f, tx); }
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
        case 14: return launch_ring<T, 32768, 3, 16, 2>(p, num_sms, s);
        case 15: return launch_ring<T, 65536, 3, 16, 1>(p, num_sms, s);
        case 16: return launch_ring<T, 98304, 2, 16, 1>(p, num_sms, s);
        case 1: return launch_ring<T, 16384, 12, 8, 1>(p, num_sms, s);
        default:
            // auto: once the volumes no longer fit L2 (>= 128 MiB) the fp32 backward streams faster with direct 128-bit loads, two
            // 512-thread CTAs per SM, than through the TMA ring (B=32 64^3: 194 vs 203 us = 6.2 vs 5.9 TB/s; D=128: 383 vs 399 us);
            // below that, and in bf16, the ring wins (profiles/r01_kbench.txt).  Both stream 32 KiB chunks (same Geometry).
            if (sizeof(T) == 4 && (uint64_t)p.g.R * p.g.N * sizeof(T) >= (128ull << 20)) return launch_direct<T, 512, 4, 2>(p, num_sms, s);
            return launch_ring<T, 32768, 6, 16, 1>(p, num_sms, s);
    }
}

void launch_bwd(const BwdParams& p, int dtype, bool vec_ok, int variant, int num_sms, cudaStream_t s) {
    if (dtype == 0) launch_bwd_t<float>(p, vec_ok, variant, num_sms, s);
    else launch_bwd_t<__nv_bfloat16>(p, vec_ok, variant, num_sms, s);
}

}  // namespace ihpr

// bn_train.cu -- K10: the BatchNorm2d (batch statistics) + ReLU of a deconv block of HeadNet in TRAINING, forward and backward
// (/root/reference/main/model.py:22-38 under main/train.py:64-71), around the tensor-core GEMMs of deconv_bn_relu.cu.  SURVEY section 8 row N1.
//
// Layout: every activation is bf16 NHWC with C = 256 channels (512 bytes per pixel); a thread owns ONE channel octet (16 bytes) for the whole
// kernel, so its per-channel constants live in registers and every access is a 128-bit load / store, 32 threads per pixel.
//
//   forward   stat_finalize   rows of (sum y, sum y^2) partials written by the GEMM's epilogue -> mean, biased var (fp64 reduction in row order:
//                             deterministic), rstd, scale = gamma * rstd, shift = beta - mean * scale, running statistics (momentum, unbiased var)
//             bn_relu_apply   out = max(0, y * scale + shift)                                           (read 2 B, write 2 B per element)
//   backward  bwd_reduce      dz = dout * [y * scale + shift > 0];  partial sums of dz and dz * y per channel, one row per CTA
//             bwd_finalize    dbeta = sum dz, dgamma = sum dz * xhat = rstd * (sum dz * y - mean * sum dz) (fp64, row order); the constants of the last pass
//             bwd_apply       dy = scale * (dz - dbeta / n - xhat * dgamma / n)                         (read 4 B, write 2 B per element)
// The ReLU mask is recomputed from the raw convolution output (the only tensor the forward keeps besides its result).
#include "ihpr_common.cuh"

namespace ihpr {
namespace k10 {

constexpr int C = 256;
constexpr int OCT = C / 8;          // channel octets per pixel = threads per pixel

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 v;
    v.x = Elem<__nv_bfloat16>::pk(f[0], f[1]);
    v.y = Elem<__nv_bfloat16>::pk(f[2], f[3]);
    v.z = Elem<__nv_bfloat16>::pk(f[4], f[5]);
    v.w = Elem<__nv_bfloat16>::pk(f[6], f[7]);
    return v;
}
__device__ __forceinline__ void load8f(const float* p, float (&f)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// 16 CTAs x 1024 threads: thread = (channel c of the CTA's 16, row slice sl of 64); rows are summed in fp64 in a fixed order
constexpr int FSL = 64, FCH = 16;
template <bool BWD>
__global__ void __launch_bounds__(FCH * FSL) finalize_kernel(const float* __restrict__ part, int rows, double inv_n, double unbias, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float eps, float momentum, float* __restrict__ running_mean,
                                                       float* __restrict__ running_var, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                       float* __restrict__ scale, float* __restrict__ shift,
                                                       // BWD: inputs mean / rstd / scale(= gamma * rstd), outputs dgamma, dbeta and the constants P, Q of bwd_apply
                                                       float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ cP, float* __restrict__ cQ) {
    __shared__ double sh[2][FSL][FCH];
    const int cl = threadIdx.x % FCH, sl = threadIdx.x / FCH;
    const int c = blockIdx.x * FCH + cl;
    double s0 = 0.0, s1 = 0.0;
#pragma unroll 8
    for (int r = sl; r < rows; r += FSL) {
        s0 += (double)__ldg(part + (size_t)r * 2 * C + c);
        s1 += (double)__ldg(part + (size_t)r * 2 * C + C + c);
    }
    sh[0][sl][cl] = s0;
    sh[1][sl][cl] = s1;
    __syncthreads();
    if (sl != 0) return;
    s0 = 0.0; s1 = 0.0;
#pragma unroll
    for (int i = 0; i < FSL; ++i) { s0 += sh[0][i][cl]; s1 += sh[1][i][cl]; }
    if (!BWD) {
        const double mean = s0 * inv_n;
        double var = s1 * inv_n - mean * mean;
        if (var < 0.0) var = 0.0;
        const float rstd = (float)(1.0 / sqrt(var + (double)eps));
        const float sc = __ldg(gamma + c) * rstd;
        mean_out[c] = (float)mean;
        rstd_out[c] = rstd;
        scale[c] = sc;
        shift[c] = __ldg(beta + c) - (float)mean * sc;
        if (running_mean) {         // torch.nn.BatchNorm2d: running = (1 - momentum) * running + momentum * batch (unbiased variance)
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)(var * unbias);
        }
    } else {
        // s0 = sum dz, s1 = sum dz * y  ->  sum dz * xhat = rstd * (s1 - mean * s0)
        const float mean = __ldg(mean_out + c), rstd = __ldg(rstd_out + c), A = __ldg(scale + c);
        s1 = (double)rstd * (s1 - (double)mean * s0);
        dbeta[c] = (float)s0;
        dgamma[c] = (float)s1;
        const float c1 = (float)(s0 * inv_n), c2 = (float)(s1 * inv_n);
        // dy = A * (dz - c1 - xhat * c2),  xhat = (y - mean) * rstd   ->   dy = A * dz + y * P + Q
        const float P = -A * c2 * rstd;
        cP[c] = P;
        cQ[c] = -A * c1 - P * mean;
    }
}

// out = max(0, y * scale + shift); n_pix pixels of 256 channels.  gridDim.x * (blockDim.x / 32) pixels per sweep.
__global__ void __launch_bounds__(256) bn_relu_apply_kernel(const uint4* __restrict__ y, uint4* __restrict__ out, size_t n_pix, const float* __restrict__ scale,
                                                            const float* __restrict__ shift) {
    const int oct = threadIdx.x & 31;
    float sc[8], sh[8];
    load8f(scale + oct * 8, sc);
    load8f(shift + oct * 8, sh);
    const size_t step = (size_t)gridDim.x * (blockDim.x >> 5);
    size_t pix = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    for (; pix + 3 * step < n_pix; pix += 4 * step) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ld_stream16(y + (pix + u * step) * OCT + oct);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float f[8];
            unpack8(v[u], f);
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = fmaxf(fmaf(f[i], sc[i], sh[i]), 0.f);
            out[(pix + u * step) * OCT + oct] = pack8(f);
        }
    }
    for (; pix < n_pix; pix += step) {
        float f[8];
        unpack8(ld_stream16(y + pix * OCT + oct), f);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = fmaxf(fmaf(f[i], sc[i], sh[i]), 0.f);
        out[pix * OCT + oct] = pack8(f);
    }
}

// partial sums of dz = dout * [y * scale + shift > 0] and dz * y; one row [2][256] per CTA, pixel lanes added in lane order
__global__ void __launch_bounds__(256, 4) bn_relu_bwd_reduce_kernel(const uint4* __restrict__ dout, const uint4* __restrict__ y, size_t n_pix,
                                                                 const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ part) {
    __shared__ float sm[8][2 * C];
    const int oct = threadIdx.x & 31, pl = threadIdx.x >> 5;
    float sc[8], sh[8];
    load8f(scale + oct * 8, sc);
    load8f(shift + oct * 8, sh);
    float a0[8], a1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a0[i] = a1[i] = 0.f;
    auto acc = [&](const uint4& gv, const uint4& yv) {
        float g[8], f[8];
        unpack8(gv, g);
        unpack8(yv, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float dz = fmaf(f[i], sc[i], sh[i]) > 0.f ? g[i] : 0.f;
            a0[i] += dz;
            a1[i] = fmaf(dz, f[i], a1[i]);
        }
    };
    const size_t step = (size_t)gridDim.x * 8;
    size_t pix = (size_t)blockIdx.x * 8 + pl;
    for (; pix + 3 * step < n_pix; pix += 4 * step) {
        uint4 gv[4], yv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            gv[u] = ld_stream16(dout + (pix + u * step) * OCT + oct);
            yv[u] = ld_stream16(y + (pix + u * step) * OCT + oct);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc(gv[u], yv[u]);
    }
    for (; pix < n_pix; pix += step) acc(ld_stream16(dout + pix * OCT + oct), ld_stream16(y + pix * OCT + oct));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        sm[pl][oct * 8 + i] = a0[i];
        sm[pl][C + oct * 8 + i] = a1[i];
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 2 * C; k += 256) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += sm[j][k];
        part[(size_t)blockIdx.x * 2 * C + k] = s;
    }
}

// dy = A * dz + y * P + Q (see finalize_kernel<true>)
__global__ void __launch_bounds__(256) bn_relu_bwd_apply_kernel(const uint4* __restrict__ dout, const uint4* __restrict__ y, uint4* __restrict__ dy, size_t n_pix,
                                                                const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ cP,
                                                                const float* __restrict__ cQ) {
    const int oct = threadIdx.x & 31;
    float sc[8], sh[8], P[8], Q[8];
    load8f(scale + oct * 8, sc);
    load8f(shift + oct * 8, sh);
    load8f(cP + oct * 8, P);
    load8f(cQ + oct * 8, Q);
    auto one = [&](const uint4& gv, const uint4& yv) {
        float g[8], f[8];
        unpack8(gv, g);
        unpack8(yv, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float dz = fmaf(f[i], sc[i], sh[i]) > 0.f ? g[i] : 0.f;
            f[i] = fmaf(sc[i], dz, fmaf(f[i], P[i], Q[i]));
        }
        return pack8(f);
    };
    const size_t step = (size_t)gridDim.x * (blockDim.x >> 5);
    size_t pix = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    for (; pix + step < n_pix; pix += 2 * step) {
        const uint4 g0 = ld_stream16(dout + pix * OCT + oct), y0 = ld_stream16(y + pix * OCT + oct);
        const uint4 g1 = ld_stream16(dout + (pix + step) * OCT + oct), y1 = ld_stream16(y + (pix + step) * OCT + oct);
        dy[pix * OCT + oct] = one(g0, y0);
        dy[(pix + step) * OCT + oct] = one(g1, y1);
    }
    for (; pix < n_pix; pix += step) dy[pix * OCT + oct] = one(ld_stream16(dout + pix * OCT + oct), ld_stream16(y + pix * OCT + oct));
}

}  // namespace k10

// ---- host side --------------------------------------------------------------------------------------------------
// The streaming kernels are grid-stride loops with equal shares per CTA: the grid is exactly the number of CTAs that are resident at once
// (a partial second wave would run at a fraction of the memory parallelism and roughly double the time -- measured with the first version).
template <typename K>
static int resident_ctas(K kern, int num_sms) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    return per_sm * num_sms;
}
// rows of partials the backward reduction writes (one per CTA): 4 CTAs of 256 threads per SM (__launch_bounds__(256, 4))
int bn_bwd_rows(int num_sms) { return 4 * num_sms; }

void launch_bn_stat_finalize(const float* part, int rows, size_t n_per_channel, const float* gamma, const float* beta, float eps, float momentum,
                             float* running_mean, float* running_var, float* mean, float* rstd, float* scale, float* shift, int* launches, cudaStream_t s) {
    const double n = (double)n_per_channel;
    k10::finalize_kernel<false><<<k10::C / k10::FCH, k10::FCH * k10::FSL, 0, s>>>(part, rows, 1.0 / n, n > 1.0 ? n / (n - 1.0) : 1.0, gamma, beta, eps, momentum, running_mean, running_var,
                                                           mean, rstd, scale, shift, nullptr, nullptr, nullptr, nullptr);
    ++*launches;
}

void launch_bn_relu_apply(const void* y, void* out, size_t n_pix, const float* scale, const float* shift, int num_sms, int* launches, cudaStream_t s) {
    static const int grid = resident_ctas(k10::bn_relu_apply_kernel, num_sms);
    k10::bn_relu_apply_kernel<<<grid, 256, 0, s>>>(static_cast<const uint4*>(y), static_cast<uint4*>(out), n_pix, scale, shift);
    ++*launches;
}

// dy, dgamma, dbeta from dout and the saved raw output; part: bn_bwd_rows(num_sms) * 512 floats; cP / cQ: 256 floats each
void launch_bn_relu_bwd(const void* dout, const void* y, void* dy, size_t n_pix, const float* scale, const float* shift, const float* mean, const float* rstd,
                        float* dgamma, float* dbeta, float* part, float* cP, float* cQ, int num_sms, int* launches, cudaStream_t s) {
    const int rows = bn_bwd_rows(num_sms);
    k10::bn_relu_bwd_reduce_kernel<<<rows, 256, 0, s>>>(static_cast<const uint4*>(dout), static_cast<const uint4*>(y), n_pix, scale, shift, part);
    k10::finalize_kernel<true><<<k10::C / k10::FCH, k10::FCH * k10::FSL, 0, s>>>(part, rows, 1.0 / (double)n_pix, 1.0, nullptr, nullptr, 0.f, 0.f, nullptr, nullptr,
                                                          const_cast<float*>(mean), const_cast<float*>(rstd), const_cast<float*>(scale), nullptr, dgamma, dbeta, cP, cQ);
    static const int grid = resident_ctas(k10::bn_relu_bwd_apply_kernel, num_sms);
    k10::bn_relu_bwd_apply_kernel<<<grid, 256, 0, s>>>(static_cast<const uint4*>(dout), static_cast<const uint4*>(y), static_cast<uint4*>(dy), n_pix, scale,
                                                              shift, cP, cQ);
    *launches += 3;
}

}  // namespace ihpr

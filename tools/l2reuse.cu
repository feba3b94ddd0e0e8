// tools/l2reuse.cu -- can a persistent CTA re-read the S bytes it just streamed from HBM out of L2 while also
// writing S bytes of output?  (design question for the fused forward+backward kernel: traffic 2V instead of 3V)
// For unit sizes S per CTA: pass 1 reads the unit, pass 2 re-reads it and writes an equally sized output unit.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int HINT>
__device__ __forceinline__ uint4 ld(const uint4* p, uint64_t pol) {
    uint4 r;
    if (HINT == 0) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    else asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
    return r;
}

// mode 0: plain hints; mode 1: pass-1 evict_last, pass-2 evict_first
template <int MODE>
__global__ void __launch_bounds__(512, 1) reuse_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t unit_vecs, size_t n_units, uint32_t* sink) {
    uint64_t pol_last, pol_first;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
    uint32_t acc = 0;
    for (size_t u = blockIdx.x; u < n_units; u += gridDim.x) {
        const uint4* src = in + u * unit_vecs;
        uint4* dst = out + u * unit_vecs;
        for (size_t i = threadIdx.x; i + 3 * 512 < unit_vecs; i += 4 * 512) {
            uint4 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = ld<MODE>(src + i + k * 512, pol_last);
#pragma unroll
            for (int k = 0; k < 4; ++k) acc ^= v[k].x ^ v[k].w;
        }
        __syncthreads();
        for (size_t i = threadIdx.x; i + 3 * 512 < unit_vecs; i += 4 * 512) {
            uint4 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = ld<MODE>(src + i + k * 512, pol_first);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                v[k].x ^= acc;
                asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst + i + k * 512), "r"(v[k].x), "r"(v[k].y), "r"(v[k].z), "r"(v[k].w) : "memory");
            }
        }
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

int main() {
    const size_t bytes = 576ull << 20;
    uint4 *a, *b; uint32_t* sink;
    cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&sink, 4);
    cudaMemset(a, 1, bytes); cudaMemset(b, 2, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode)
        for (size_t unit : {128ull << 10, 256ull << 10, 512ull << 10, 1024ull << 10, 2048ull << 10}) {
            const size_t uv = unit / 16, nu = bytes / unit;
            float best = 1e30f;
            for (int it = 0; it < 6; ++it) {
                cudaEventRecord(e0);
                if (mode == 0) reuse_kernel<0><<<148, 512>>>(a, b, uv, nu, sink); else reuse_kernel<1><<<148, 512>>>(a, b, uv, nu, sink);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (it > 0 && ms < best) best = ms;
            }
            printf("mode %d unit %5zu KiB (in flight %4zu MiB): %7.1f us  -> %6.0f GB/s counted as 3V, %6.0f GB/s as 2V\n", mode, unit >> 10, (148 * unit) >> 20,
                   best * 1e3, 3.0 * bytes / best / 1e6, 2.0 * bytes / best / 1e6);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

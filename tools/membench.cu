// tools/membench.cu -- what HBM bandwidth can a kernel of each access shape reach on this B200?
// (context for the roofline fractions: MEASURED_PEAKS.json is a read+write COPY; K1 is read-only, K2 is a copy)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/membench tools/membench.cu && /tmp/membench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint4 ldnc(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

template <int U>
__global__ void read_kernel(const uint4* __restrict__ in, size_t n, uint32_t* out) {
    uint32_t acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < n; i += U * stride) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ldnc(in + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// contiguous per-CTA ranges (like the product kernels) instead of grid-stride
template <int U>
__global__ void read_ranges_kernel(const uint4* __restrict__ in, size_t n, uint32_t* out) {
    uint32_t acc = 0;
    const size_t lo = n * blockIdx.x / gridDim.x, hi = n * (blockIdx.x + 1) / gridDim.x;
    for (size_t i = lo + threadIdx.x; i + (U - 1) * blockDim.x < hi; i += U * blockDim.x) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ldnc(in + i + u * blockDim.x);
#pragma unroll
        for (int u = 0; u < U; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

template <int U>
__global__ void copy_kernel(const uint4* __restrict__ in, uint4* __restrict__ outp, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < n; i += U * stride) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ldnc(in + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) outp[i + u * stride] = v[u];
    }
}

__global__ void write_kernel(uint4* __restrict__ outp, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) outp[i] = make_uint4(1, 2, 3, 4);
}

// TMA bulk ring, no compute: one producer thread, consumers just release the stage
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int CHUNK, int STAGES>
__global__ void __launch_bounds__(64, 1) bulk_read_kernel(const uint8_t* __restrict__ in, size_t nchunks, uint32_t* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * CHUNK);
    uint64_t* empty = full + STAGES;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(full + s)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(empty + s)));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t lo = nchunks * blockIdx.x / gridDim.x, hi = nchunks * (blockIdx.x + 1) / gridDim.x;
    auto wait = [](uint64_t* bar, uint32_t par) {
        asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(
                         s32(bar)), "r"(par) : "memory");
    };
    if (threadIdx.x == 0) {
        uint32_t it = 0;
        for (size_t c = lo; c < hi; ++c, ++it) {
            const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
            wait(empty + s, ph ^ 1);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(full + s)), "r"(CHUNK) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(smem + (size_t)s * CHUNK)),
                         "l"(in + c * CHUNK), "r"(CHUNK), "r"(s32(full + s)) : "memory");
        }
    } else if (threadIdx.x == 32) {
        uint32_t it = 0, acc = 0;
        for (size_t c = lo; c < hi; ++c, ++it) {
            const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
            wait(full + s, ph);
            acc ^= *reinterpret_cast<volatile uint32_t*>(smem + (size_t)s * CHUNK);
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(empty + s)) : "memory");
        }
        if (acc == 0x12345678u) out[0] = acc;
    }
}

template <typename F>
float time_it(F f, int iters) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int i = 0; i < iters; ++i) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    const size_t bytes = 576ull << 20;          // one B=32,J=18,64^3 fp32 heatmap batch
    const size_t n = bytes / 16;
    uint4 *a, *b; uint32_t* out;
    CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMalloc(&out, 4));
    CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 2, bytes));
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs %d, buffer %zu MiB\n", sms, bytes >> 20);
    for (int bps : {2, 4, 8}) {
        for (int threads : {256, 512, 1024}) {
            if (bps * threads > 2048) continue;
            float ms = time_it([&] { read_kernel<4><<<sms * bps, threads>>>(a, n, out); }, 10);
            printf("read   grid-stride U=4  %4d thr x %d/SM : %7.1f GB/s\n", threads, bps, bytes / ms / 1e6);
            ms = time_it([&] { read_kernel<8><<<sms * bps, threads>>>(a, n, out); }, 10);
            printf("read   grid-stride U=8  %4d thr x %d/SM : %7.1f GB/s\n", threads, bps, bytes / ms / 1e6);
            ms = time_it([&] { read_ranges_kernel<4><<<sms * bps, threads>>>(a, n, out); }, 10);
            printf("read   cta-ranges  U=4  %4d thr x %d/SM : %7.1f GB/s\n", threads, bps, bytes / ms / 1e6);
            ms = time_it([&] { copy_kernel<4><<<sms * bps, threads>>>(a, b, n); }, 10);
            printf("copy   grid-stride U=4  %4d thr x %d/SM : %7.1f GB/s (r+w)\n", threads, bps, 2.0 * bytes / ms / 1e6);
        }
    }
    float ms = time_it([&] { write_kernel<<<sms * 4, 512>>>(b, n); }, 10);
    printf("write  grid-stride      512 thr x 4/SM : %7.1f GB/s\n", bytes / ms / 1e6);
    ms = time_it([&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); }, 10);
    printf("cudaMemcpy D2D                          : %7.1f GB/s (r+w)\n", 2.0 * bytes / ms / 1e6);
    {
        constexpr int CH = 32768, ST = 6;
        auto k = bulk_read_kernel<CH, ST>;
        const int smem = CH * ST + 2 * ST * 8;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        ms = time_it([&] { k<<<sms, 64, smem>>>(reinterpret_cast<const uint8_t*>(a), bytes / CH, out); }, 10);
        printf("read   TMA bulk ring 32K x 6, 1 CTA/SM  : %7.1f GB/s\n", bytes / ms / 1e6);
    }
    {
        constexpr int CH = 16384, ST = 6;
        auto k = bulk_read_kernel<CH, ST>;
        const int smem = CH * ST + 2 * ST * 8;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        ms = time_it([&] { k<<<sms * 2, 64, smem>>>(reinterpret_cast<const uint8_t*>(a), bytes / CH, out); }, 10);
        printf("read   TMA bulk ring 16K x 6, 2 CTA/SM  : %7.1f GB/s\n", bytes / ms / 1e6);
    }
    CK(cudaGetLastError());
    return 0;
}

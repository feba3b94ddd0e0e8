// tools/mma_issue_bench.cu -- how deep is the tcgen05.mma issue queue?  One thread issues 32 MMAs (M = 128, N = 256, K = 16, bf16, operands =
// whatever is in shared memory) back to back and stamps clock64 after every issue, then commits and waits for completion.  If the issue were
// fully asynchronous the stamps would be a few clocks apart and the tail (last issue -> completion) would be 32 x 128 clk; if the queue is
// shallow the stamps settle at the MMA rate (128 clk) after the first few.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mma_issue_bench tools/mma_issue_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {      // K-major, SWIZZLE_128B, SBO = 1024 B
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

template <int N>
__global__ void __launch_bounds__(128, 1) issue_kernel(long long* out, int nmma) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;     // bf16 ~0.0078
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&slot);
    if (threadIdx.x == 0) {
        const uint64_t ad = umma_desc(smem_u32(smem)), bd = umma_desc(smem_u32(smem + 16384));
        constexpr uint32_t idesc = make_idesc(128, N);
        long long t[66];
        t[0] = clock64();
        for (int i = 0; i < nmma; ++i) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                         "l"(ad + 2 * (i & 3)), "l"(bd + 2 * (i & 3)), "r"(idesc), "r"((uint32_t)(i != 0))
                         : "memory");
            t[i + 1] = clock64();
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        const long long tc = clock64();
        asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
        const long long te = clock64();
        if (blockIdx.x == 0) {
            for (int i = 0; i <= nmma; ++i) out[i] = t[i] - t[0];
            out[nmma + 1] = tc - t[0];
            out[nmma + 2] = te - t[0];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int N>
static void run(int nmma) {
    long long* d;
    cudaMalloc(&d, 80 * sizeof(long long));
    cudaFuncSetAttribute(issue_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 50 * 1024);
    for (int rep = 0; rep < 2; ++rep) issue_kernel<N><<<148, 128, 50 * 1024>>>(d, nmma);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[80];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("N = %3d, %d MMAs of 128 x %d x 16: clk after each issue (relative):", N, nmma, N);
    for (int i = 1; i <= nmma; ++i) printf(" %lld", h[i]);
    printf("\n   commit issued at %lld, all complete at %lld  -> %.1f clk per MMA overall   %s\n", h[nmma + 1], h[nmma + 2], (double)h[nmma + 2] / nmma,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    run<256>(32);
    run<128>(32);
    run<64>(32);
    run<256>(8);
    return 0;
}

// tools/tmem_bench.cu -- how fast can the epilogue warps pull an accumulator tile out of tensor memory?
// One CTA per SM, 512 TMEM columns allocated, W warps (multiple of 4) each reading its own lane quarter with
// tcgen05.ld 32x32b.xN; per (warps, N, outstanding loads) the kernel reports clocks per 128 x 256 fp32 tile (128 KiB) and B/clk/SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tmem_bench tools/tmem_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int N> struct Ld;
template <> struct Ld<16> {
    static __device__ __forceinline__ void issue(uint32_t taddr, uint32_t (&r)[16]) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                       "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr));
    }
};
template <> struct Ld<32> {
    static __device__ __forceinline__ void issue(uint32_t taddr, uint32_t (&r)[32]) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
              "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
              "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
              "=r"(r[31])
            : "r"(taddr));
    }
};
template <int N>
__device__ __forceinline__ void ld_wait(uint32_t (&r)[N]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < N; ++i) asm volatile("" : "+r"(r[i]));
}

// every warp reads `cols_per_warp` columns of its lane quarter per tile, N columns per load, DEPTH loads in flight before a wait
template <int N, int DEPTH>
__global__ void __launch_bounds__(640, 1) tmem_read_kernel(int warps, int tiles, unsigned long long* out, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = *reinterpret_cast<volatile uint32_t*>(&slot);
    uint32_t acc = 0;
    unsigned long long t0 = 0, t1 = 0;
    if (warp < warps) {
        const int qd = warp & 3, cg = warp >> 2, ngroups = warps / 4;
        const int cols_per_warp = 256 / ngroups;            // the 256 columns of a tile are split over the warps of a lane quarter
        const uint32_t tb = base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(cg * cols_per_warp);
        __syncwarp();
        t0 = clock64();
        for (int t = 0; t < tiles; ++t) {
            for (int c = 0; c < cols_per_warp; c += N * DEPTH) {
                uint32_t r[DEPTH][N];
#pragma unroll
                for (int d = 0; d < DEPTH; ++d) Ld<N>::issue(tb + (uint32_t)(c + d * N), r[d]);
#pragma unroll
                for (int d = 0; d < DEPTH; ++d) ld_wait<N>(r[d]);
#pragma unroll
                for (int d = 0; d < DEPTH; ++d)
#pragma unroll
                    for (int i = 0; i < N; ++i) acc ^= r[d][i];
            }
        }
        t1 = clock64();
    }
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}

template <int N, int DEPTH>
static void run(int warps, int grid) {
    unsigned long long* out;
    uint32_t* sink;
    cudaMalloc(&out, 8);
    cudaMalloc(&sink, 4);
    const int tiles = 2000;
    tmem_read_kernel<N, DEPTH><<<grid, 640>>>(warps, 10, out, sink);
    tmem_read_kernel<N, DEPTH><<<grid, 640>>>(warps, tiles, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long clk = 0;
    cudaMemcpy(&clk, out, 8, cudaMemcpyDeviceToHost);
    const double per_tile = (double)clk / tiles;
    printf("warps %2d  x%-2d  depth %d : %8.1f clk per 128x256 fp32 tile  = %6.1f B/clk/SM   %s\n", warps, N, DEPTH, per_tile, 131072.0 / per_tile,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(out);
    cudaFree(sink);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int warps : {4, 8, 16}) {
        run<16, 1>(warps, sms);
        run<16, 2>(warps, sms);
        run<16, 4>(warps, sms);
        run<32, 1>(warps, sms);
        run<32, 2>(warps, sms);
    }
    return 0;
}

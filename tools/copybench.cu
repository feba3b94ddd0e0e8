// tools/copybench.cu -- which store path should a streaming read-modify-write kernel (K2, K5's backward half) use on B200?
// All variants read a 576 MiB buffer through a TMA bulk ring (32 KiB x STAGES, one CTA per SM) and write 576 MiB:
//   mode 0  pure DMA: the landed stage is bulk-stored straight back to global (upper bound, no SM data path at all)
//   mode 1  consumers (16 warps) read the stage from shared memory, touch it, write it back IN PLACE, one thread bulk-stores it
//   mode 2  consumers read the stage from shared memory and store to global with 128-bit STG (what K2 does today)
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/copybench tools/copybench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t par) {
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(s32(bar)),
                 "r"(par) : "memory");
}

constexpr int NCW = 16;      // consumer warps

template <int CHUNK, int STAGES, int MODE, int ILV>
__global__ void __launch_bounds__(32 * (NCW + 2), 1) copy_ring_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ outp, size_t nchunks) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * CHUNK);
    uint64_t* empty = full + STAGES;
    uint64_t* done = empty + STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, MODE == 2 ? NCW : 1);
            mbar_init(done + s, NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // ILV 0: each CTA streams one contiguous range; ILV 1: chunks dealt round-robin (chunk = first + it * step)
    const size_t lo = ILV ? blockIdx.x : nchunks * blockIdx.x / gridDim.x, hi = ILV ? nchunks : nchunks * (blockIdx.x + 1) / gridDim.x;
    const size_t step = ILV ? gridDim.x : 1;
    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (size_t c = lo; c < hi; c += step, ++it) {
                const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(empty + s, ph ^ 1);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(full + s)), "r"(CHUNK) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(smem + (size_t)s * CHUNK)),
                             "l"(in + c * CHUNK), "r"(CHUNK), "r"(s32(full + s)) : "memory");
            }
        }
    } else if (warp == 1) {
        if (MODE != 2 && lane == 0) {      // the store thread
            uint32_t it = 0;
            for (size_t c = lo; c < hi; c += step, ++it) {
                const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(MODE == 0 ? full + s : done + s, ph);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(outp + c * CHUNK), "r"(s32(smem + (size_t)s * CHUNK)), "r"(CHUNK)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                if (it >= 1) {             // all but the newest store have finished READING shared memory: release the previous stage
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    mbar_arrive(empty + (it - 1) % STAGES);
                }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else if (MODE != 0) {
        const int cw = warp - 2;
        uint32_t it = 0;
        for (size_t c = lo; c < hi; c += step, ++it) {
            const uint32_t s = it % STAGES, ph = (it / STAGES) & 1;
            mbar_wait(full + s, ph);
            float4* st = reinterpret_cast<float4*>(smem + (size_t)s * CHUNK);
            float4* g = reinterpret_cast<float4*>(outp + c * CHUNK);
#pragma unroll
            for (int k = 0; k < CHUNK / 16 / (NCW * 32); ++k) {
                const int i = k * NCW * 32 + cw * 32 + lane;
                float4 v = st[i];
                v.x *= 1.0001f, v.y *= 1.0001f, v.z *= 1.0001f, v.w *= 1.0001f;
                if (MODE == 1) st[i] = v; else g[i] = v;
            }
            if (MODE == 1) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the bulk store
                __syncwarp();
                if (lane == 0) mbar_arrive(done + s);
            } else {
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + s);
            }
        }
    }
}

template <int MODE, int STAGES, int ILV>
int run(const uint8_t* a, uint8_t* b, size_t bytes, int sms, const char* name) {
    constexpr int CH = 32768;
    auto k = copy_ring_kernel<CH, STAGES, MODE, ILV>;
    const int smem = STAGES * CH + 3 * STAGES * 8;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) k<<<sms, 32 * (NCW + 2), smem>>>(a, b, bytes / CH);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < 10; ++i) {
        cudaEventRecord(e0); k<<<sms, 32 * (NCW + 2), smem>>>(a, b, bytes / CH); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    printf("%-58s %s %d stages: %7.1f us  %7.1f GB/s (r+w)\n", name, ILV ? "round-robin chunks" : "contiguous ranges ", STAGES, best * 1e3, 2.0 * bytes / best / 1e6);
    return 0;
}

int main() {
    const size_t bytes = 576ull << 20;
    uint8_t *a, *b;
    CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes));
    CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 2, bytes));
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    if (run<0, 6, 0>(a, b, bytes, sms, "mode 0  TMA load -> TMA store (pure DMA)")) return 1;
    if (run<1, 6, 0>(a, b, bytes, sms, "mode 1  TMA load -> LDS/STS in place -> TMA store")) return 1;
    if (run<2, 6, 0>(a, b, bytes, sms, "mode 2  TMA load -> LDS -> STG.128")) return 1;
    if (run<0, 6, 1>(a, b, bytes, sms, "mode 0  TMA load -> TMA store (pure DMA)")) return 1;
    if (run<1, 6, 1>(a, b, bytes, sms, "mode 1  TMA load -> LDS/STS in place -> TMA store")) return 1;
    if (run<2, 6, 1>(a, b, bytes, sms, "mode 2  TMA load -> LDS -> STG.128")) return 1;
    if (run<0, 4, 0>(a, b, bytes, sms, "mode 0  TMA load -> TMA store (pure DMA)")) return 1;
    if (run<2, 4, 1>(a, b, bytes, sms, "mode 2  TMA load -> LDS -> STG.128")) return 1;
    // check mode 1 really transformed the data
    float h[4]; CK(cudaMemcpy(h, b, 16, cudaMemcpyDeviceToHost));
    printf("sample out %g\n", h[0]);
    return 0;
}

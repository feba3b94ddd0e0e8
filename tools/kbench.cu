// tools/kbench.cu -- Python-free timing of the C-ABI entry points (links lib/libihpr_b200.so).
//   usage: kbench [variant] [B] [dtype 0|1] [iters] [D]
// Prints back-to-back launch time (events around `iters` launches) for fwd, bwd, and fwd+bwd alternating.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cuda_runtime.h>
#include <time.h>
#include <unistd.h>
#include "../include/ihpr_b200.h"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
#define IK(x) do { int rc = (x); if (rc) { printf("%s -> %d: %s\n", #x, rc, ihpr_last_error()); return 1; } } while (0)

__global__ void fill(float* p, size_t n, uint32_t seed) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t x = (uint32_t)i * 2654435761u + seed; x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13;
        p[i] = ((x & 0xffff) / 65536.0f - 0.5f) * 6.0f;
    }
}
__global__ void fill_bf16(unsigned short* p, size_t n, uint32_t seed) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t x = (uint32_t)i * 2654435761u + seed; x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13;
        float f = ((x & 0xffff) / 65536.0f - 0.5f) * 6.0f;
        p[i] = (unsigned short)(__float_as_uint(f) >> 16);
    }
}

int main(int argc, char** argv) {
    int variant = argc > 1 ? atoi(argv[1]) : 0;
    int B = argc > 2 ? atoi(argv[2]) : 32;
    int dtype = argc > 3 ? atoi(argv[3]) : 0;
    int iters = argc > 4 ? atoi(argv[4]) : 20;
    const int D = argc > 5 ? atoi(argv[5]) : 64;
    const int J = 18, H = 64, W = 64;
    const size_t R = (size_t)B * J, N = (size_t)D * H * W, es = dtype ? 2 : 4;
    void *heat, *grad, *ws; float *gt, *vis, *hd, *loss, *coords, *stats, *go;
    CK(cudaMalloc(&heat, R * N * es)); CK(cudaMalloc(&grad, R * N * es));
    size_t wsb = ihpr_workspace_bytes(B, J, D, H, W);
    CK(cudaMalloc(&ws, wsb)); CK(cudaMemset(ws, 0, wsb));
    CK(cudaMalloc(&gt, R * 3 * 4)); CK(cudaMalloc(&vis, R * 4)); CK(cudaMalloc(&hd, B * 4)); CK(cudaMalloc(&loss, 4));
    CK(cudaMalloc(&coords, R * 3 * 4)); CK(cudaMalloc(&stats, R * 2 * 4)); CK(cudaMalloc(&go, 4));
    if (dtype) fill_bf16<<<1024, 256>>>((unsigned short*)heat, R * N, 1); else fill<<<1024, 256>>>((float*)heat, R * N, 1);
    std::vector<float> ones(R * 3, 1.0f);
    CK(cudaMemcpy(vis, ones.data(), R * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(hd, ones.data(), B * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(go, ones.data(), 4, cudaMemcpyHostToDevice));
    for (auto& v : ones) v = 31.0f;
    CK(cudaMemcpy(gt, ones.data(), R * 3 * 4, cudaMemcpyHostToDevice));
    ihpr_set_variant(variant);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto fwd = [&] { return ihpr_integral_l1_fwd(heat, dtype, B, J, D, H, W, gt, vis, hd, loss, coords, stats, ws, wsb, nullptr); };
    auto bwd = [&] { return ihpr_integral_l1_bwd(heat, dtype, B, J, D, H, W, coords, stats, gt, vis, hd, go, grad, nullptr); };
    const bool verbose = getenv("KB_VERBOSE") != nullptr;
    const bool only_fused = getenv("KB_ONLY_FUSED") != nullptr;
    auto fused = [&] { return ihpr_integral_l1_fwd_bwd(heat, dtype, B, J, D, H, W, gt, vis, hd, loss, coords, stats, grad, ws, wsb, nullptr); };
    if (only_fused) {
        for (int i = 0; i < iters; ++i) {
            IK(fused());
            if (verbose) { printf("fused launch %d enqueued (%d kernels)\n", i, ihpr_last_launch_count()); fflush(stdout); }
            CK(cudaDeviceSynchronize());
            if (verbose) { printf("fused launch %d done\n", i); fflush(stdout); }
        }
        return 0;
    }
    for (int i = 0; i < 3; ++i) { IK(fwd()); IK(bwd()); }
    CK(cudaDeviceSynchronize());
    if (verbose) { printf("K1/K2 warm-up done\n"); fflush(stdout); }
    float ms;
    cudaEventRecord(e0); for (int i = 0; i < iters; ++i) IK(fwd()); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms, e0, e1); const double tf = ms * 1e3 / iters;
    cudaEventRecord(e0); for (int i = 0; i < iters; ++i) IK(bwd()); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms, e0, e1); const double tb = ms * 1e3 / iters;
    // inference call (main/test.py:53-65): coordinates only, no statistics, no loss
    auto inf = [&] { return ihpr_softargmax3d_fwd(heat, dtype, B, J, D, H, W, coords, nullptr, ws, wsb, nullptr); };
    for (int i = 0; i < 3; ++i) IK(inf());
    cudaEventRecord(e0); for (int i = 0; i < iters; ++i) IK(inf()); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms, e0, e1); const double ti = ms * 1e3 / iters;
    IK(fwd());      // restore coords / stats / loss of the training forward for what follows
    cudaEventRecord(e0); for (int i = 0; i < iters; ++i) { IK(fwd()); IK(bwd()); } cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms, e0, e1); const double tfb = ms * 1e3 / iters;
    if (verbose) { printf("K1/K2 timing done\n"); fflush(stdout); }
    // keep K1 + K2's answer (coords, head and tail of the gradient) to cross-check the one-launch path against
    const size_t probe = std::min<size_t>(R * N * es, 8u << 20);
    std::vector<unsigned char> g_ref(2 * probe), g_fus(2 * probe);
    std::vector<float> c_ref(R * 3), c_fus(R * 3);
    CK(cudaMemcpy(g_ref.data(), grad, probe, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(g_ref.data() + probe, (char*)grad + R * N * es - probe, probe, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(c_ref.data(), coords, R * 3 * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemset(grad, 0xff, R * N * es));
    for (int i = 0; i < 3; ++i) IK(fused());
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); for (int i = 0; i < iters; ++i) IK(fused()); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms, e0, e1); const double tfu = ms * 1e3 / iters;
    float hl; CK(cudaMemcpy(&hl, loss, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(g_fus.data(), grad, probe, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(g_fus.data() + probe, (char*)grad + R * N * es - probe, probe, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(c_fus.data(), coords, R * 3 * 4, cudaMemcpyDeviceToHost));
    {
        auto val = [&](const std::vector<unsigned char>& b, size_t i) -> double {
            if (dtype == 0) return ((const float*)b.data())[i];
            unsigned int u = ((unsigned int)((const unsigned short*)b.data())[i]) << 16; float f; memcpy(&f, &u, 4); return f;
        };
        double dmax = 0, amax = 0, cmax = 0;
        for (size_t i = 0; i < 2 * probe / es; ++i) { double a = val(g_ref, i), b = val(g_fus, i); dmax = std::max(dmax, std::fabs(a - b)); amax = std::max(amax, std::fabs(a)); }
        for (size_t i = 0; i < R * 3; ++i) cmax = std::max(cmax, (double)std::fabs(c_ref[i] - c_fus[i]));
        printf("variant %d B %d dtype %d: one-launch vs K1+K2: max |dgrad| %.3e (max |grad| %.3e, rel %.2e), max |dcoords| %.2e\n", variant, B, dtype, dmax, amax,
               amax > 0 ? dmax / amax : 0.0, cmax);
    }
    const double V = (double)R * N * es;
    printf("variant %d B %d dtype %d: FUSED one-launch fwd+bwd %.1f us (%d launch; %.0f GB/s as 3V, %.0f GB/s as 2V, %.0f vol/s)\n", variant, B, dtype, tfu,
           ihpr_last_launch_count(), 3 * (double)R * N * es / tfu / 1e3, 2 * (double)R * N * es / tfu / 1e3, R / (tfu * 1e-6));
    printf("variant %d B %d dtype %d: fwd %.1f us (%.0f GB/s)  inference fwd %.1f us  bwd %.1f us (%.0f GB/s)  fwd+bwd %.1f us (%.0f GB/s, %.0f vol/s)  loss %.5f\n", variant, B, dtype,
           tf, V / tf / 1e3, ti, tb, 2 * V / tb / 1e3, tfb, 3 * V / tfb / 1e3, R / (tfb * 1e-6), hl);
    return 0;
}

// ihpr_common.cuh -- device helpers shared by the soft-argmax forward / backward kernels (sm_100a).
//
// Reference semantics being reproduced: /root/reference/common/nets/loss.py:13-52.
// Vocabulary: a "joint-volume" (row) is the N = D*H*W run of one (sample, joint); a "quad" is four
// consecutive voxels along x (W % 4 == 0 on the vector paths, so a quad never crosses an x-row);
// a "chunk" is the unit the persistent CTAs stream: CE consecutive voxels of one joint-volume.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace ihpr {

constexpr float kLog2e = 1.4426950408889634f;
constexpr int kGridCap = 2048;          // upper bound on persistent CTAs (workspace sizing)
constexpr int kMaxSplit = 32;           // K5: at most this many CTAs share one joint-volume (one exchanger lane per partner)
constexpr int kMinChunkElems = 2048;    // every kernel config streams chunks of >= this many voxels

// ---------------------------------------------------------------------------------------------
// exact n / d for n < 2^31 (Granlund-Montgomery round-up multiplier): 3 integer instructions
struct FastDiv {
    uint32_t d, mul, shr;
};
__host__ __device__ inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.d = d;
    uint32_t s = 0;
    while ((1ull << s) < d) ++s;
    f.shr = s;
    f.mul = (uint32_t)((((1ull << s) - d) << 32) / d + 1);
    return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) { return (__umulhi(n, f.mul) + n) >> f.shr; }

// small non-negative int -> float without the conversion pipe (n < 2^23)
__device__ __forceinline__ float u2f(uint32_t n) { return __uint_as_float(0x4B000000u | n) - 8388608.0f; }

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---------------------------------------------------------------------------------------------
// Online-softmax partial state of one joint-volume: l = sum 2^(h*log2e - c), s* = sum of the same
// weights times the voxel coordinate.  c = fl(m * log2e) is a pure function of the reference point m,
// so partials from different threads / CTAs re-base exactly.  The reference point is LAZY: it is only
// raised when a batch maximum exceeds it by more than kRebaseSlack (weights then stay below e^30, far
// from fp32 overflow even summed over 2^30 voxels), so the re-base branch runs once per joint-volume on
// ordinary data instead of once per batch; the true maximum is tracked separately in mx and the final
// (m, l) pair is re-based to it.  m == -inf means "nothing seen yet" (c = 0 keeps the arithmetic
// finite); +inf / NaN inputs turn the row into NaN like torch's softmax.
constexpr float kRebaseSlack = 30.f;
struct Acc {
    float m, c, lim, mx, l, sx, sy, sz;
    __device__ __forceinline__ void reset() {
        m = lim = mx = -INFINITY;
        c = 0.f;
        l = sx = sy = sz = 0.f;
    }
};
__device__ __forceinline__ float safe_c(float m) { return (m == -INFINITY) ? 0.f : m * kLog2e; }

// move the reference point up to cmax (> a.m) and re-base the sums
__device__ __forceinline__ void acc_raise(Acc& a, float cmax) {
    const float cn = safe_c(cmax);
    const float s = (a.m == -INFINITY) ? 0.f : ex2(a.c - cn);
    a.l *= s; a.sx *= s; a.sy *= s; a.sz *= s;
    a.m = cmax; a.c = cn; a.lim = cmax + kRebaseSlack;
}
// bookkeeping for a batch whose maximum is cmax, before its weights are accumulated
__device__ __forceinline__ void acc_see_max(Acc& a, float cmax) {
    a.mx = fmaxf(a.mx, cmax);
    if (cmax > a.lim) acc_raise(a, cmax);
}

__device__ __forceinline__ Acc acc_merge(const Acc& a, const Acc& b) {
    Acc o;
    o.m = fmaxf(a.m, b.m);
    o.c = safe_c(o.m);
    o.mx = fmaxf(a.mx, b.mx);
    o.lim = o.m + kRebaseSlack;
    const float fa = (a.m == -INFINITY) ? 0.f : ex2(a.c - o.c);
    const float fb = (b.m == -INFINITY) ? 0.f : ex2(b.c - o.c);
    o.l = a.l * fa + b.l * fb;
    o.sx = a.sx * fa + b.sx * fb;
    o.sy = a.sy * fa + b.sy * fb;
    o.sz = a.sz * fa + b.sz * fb;
    return o;
}

__device__ __forceinline__ Acc acc_warp_merge(Acc a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Acc b;
        b.m = __shfl_xor_sync(0xffffffffu, a.m, o);
        b.mx = __shfl_xor_sync(0xffffffffu, a.mx, o);
        b.l = __shfl_xor_sync(0xffffffffu, a.l, o);
        b.sx = __shfl_xor_sync(0xffffffffu, a.sx, o);
        b.sy = __shfl_xor_sync(0xffffffffu, a.sy, o);
        b.sz = __shfl_xor_sync(0xffffffffu, a.sz, o);
        b.c = safe_c(b.m);
        a = acc_merge(a, b);
    }
    return a;
}

// partial <-> 8-float slot {m, l, sx, sy, sz, mx, -, -}
__device__ __forceinline__ void partial_to_smem(volatile float* d, const Acc& a) {
    d[0] = a.m; d[1] = a.l; d[2] = a.sx; d[3] = a.sy; d[4] = a.sz; d[5] = a.mx;
}
__device__ __forceinline__ Acc partial_from_smem(const volatile float* s) {
    Acc a;
    a.m = s[0]; a.l = s[1]; a.sx = s[2]; a.sy = s[3]; a.sz = s[4]; a.mx = s[5];
    a.c = safe_c(a.m); a.lim = a.m + kRebaseSlack;
    return a;
}
__device__ __forceinline__ void partial_to_global(float* d, const Acc& a) {
    __stcg(d + 0, a.m); __stcg(d + 1, a.l); __stcg(d + 2, a.sx); __stcg(d + 3, a.sy); __stcg(d + 4, a.sz); __stcg(d + 5, a.mx);
}
__device__ __forceinline__ Acc partial_from_global(const float* s) {
    Acc a;
    a.m = __ldcg(s + 0); a.l = __ldcg(s + 1); a.sx = __ldcg(s + 2); a.sy = __ldcg(s + 3); a.sz = __ldcg(s + 4); a.mx = __ldcg(s + 5);
    a.c = safe_c(a.m); a.lim = a.m + kRebaseSlack;
    return a;
}

// One quad (4 voxels at x0..x0+3, same y, z): weights p_k = 2^(h_k*log2e - c).
__device__ __forceinline__ void acc_quad(Acc& a, const float (&v)[4], float xf, float yf, float zf) {
    const float p0 = ex2(fmaf(v[0], kLog2e, -a.c));
    const float p1 = ex2(fmaf(v[1], kLog2e, -a.c));
    const float p2 = ex2(fmaf(v[2], kLog2e, -a.c));
    const float p3 = ex2(fmaf(v[3], kLog2e, -a.c));
    const float s = (p0 + p1) + (p2 + p3);
    const float w = fmaf(p3, 3.f, fmaf(p2, 2.f, p1));
    a.l += s;
    a.sx = fmaf(s, xf, a.sx + w);
    a.sy = fmaf(s, yf, a.sy);
    a.sz = fmaf(s, zf, a.sz);
}

// ---------------------------------------------------------------------------------------------
// mbarrier + TMA bulk-copy (cp.async.bulk, 1-D: no tensor map needed) wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
#ifdef IHPR_DEBUG_HANG
// debug build: a wait that spins "forever" reports who is stuck on what and traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag = 0) {
    for (unsigned long long spins = 0;; ++spins) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return;
        if (spins > (1ull << 22)) {
            printf("HANG cta %d warp %d lane %d tag %d parity %u\n", blockIdx.x, threadIdx.x >> 5, threadIdx.x & 31, tag, parity);
            __trap();
        }
    }
}
#else
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag = 0) {
    (void)tag;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
#endif
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// global -> shared bulk copy completing on an mbarrier; bytes % 16 == 0, both addresses 16 B aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- thread-block clusters / distributed shared memory
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory object in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t map_to_cta(const void* p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// release at cluster scope: the stores above are visible to whoever acquires this barrier's phase
__device__ __forceinline__ void mbar_arrive_remote(uint32_t addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
#ifdef IHPR_DEBUG_HANG
    for (unsigned long long spins = 0;; ++spins) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return;
        if (spins > (1ull << 22)) {
            printf("HANG (cluster exchange) cta %d lane %d parity %u\n", blockIdx.x, threadIdx.x & 31, parity);
            __trap();
        }
    }
#else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAITC_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONEC_%=;\n\t"
        "bra WAITC_%=;\n\t"
        "DONEC_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
#endif
}

// 16 B loads / stores
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream16(void* p, const uint4& v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds16(const void* p) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(smem_u32(p)));
    return r;
}

// element traits: QPV = quads per 16-byte vector
template <typename T> struct Elem;
template <> struct Elem<float> {
    static constexpr int QPV = 1;
    __device__ static __forceinline__ void unpack(const uint4& u, float (&v)[1][4]) {
        v[0][0] = __uint_as_float(u.x); v[0][1] = __uint_as_float(u.y);
        v[0][2] = __uint_as_float(u.z); v[0][3] = __uint_as_float(u.w);
    }
    __device__ static __forceinline__ uint4 pack(const float (&v)[1][4]) {
        return make_uint4(__float_as_uint(v[0][0]), __float_as_uint(v[0][1]), __float_as_uint(v[0][2]), __float_as_uint(v[0][3]));
    }
    __device__ static __forceinline__ float load1(const float* p) { return __ldg(p); }
    __device__ static __forceinline__ void store1(float* p, float v) { *p = v; }
};
template <> struct Elem<__nv_bfloat16> {
    static constexpr int QPV = 2;
    __device__ static __forceinline__ void unpack(const uint4& u, float (&v)[2][4]) {
        // bf16 -> fp32 is a 16-bit shift
        v[0][0] = __uint_as_float(u.x << 16); v[0][1] = __uint_as_float(u.x & 0xffff0000u);
        v[0][2] = __uint_as_float(u.y << 16); v[0][3] = __uint_as_float(u.y & 0xffff0000u);
        v[1][0] = __uint_as_float(u.z << 16); v[1][1] = __uint_as_float(u.z & 0xffff0000u);
        v[1][2] = __uint_as_float(u.w << 16); v[1][3] = __uint_as_float(u.w & 0xffff0000u);
    }
    __device__ static __forceinline__ uint32_t pk(float lo, float hi) {
        __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    __device__ static __forceinline__ uint4 pack(const float (&v)[2][4]) {
        return make_uint4(pk(v[0][0], v[0][1]), pk(v[0][2], v[0][3]), pk(v[1][0], v[1][1]), pk(v[1][2], v[1][3]));
    }
    __device__ static __forceinline__ float load1(const __nv_bfloat16* p) { return __bfloat162float(*p); }
    __device__ static __forceinline__ void store1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// ---------------------------------------------------------------------------------------------
// kernel parameters (plain data, passed by value)
struct Geometry {
    int R;              // joint-volumes = B*J
    int J;
    int D, H, W;
    uint32_t N;         // voxels per joint-volume
    uint32_t CE;        // voxels per chunk
    uint32_t nch;       // chunks per joint-volume = ceil(N / CE)
    uint64_t Gt;        // total chunks = R * nch
    FastDiv divF;       // quads per x-row (W/4)         [vector paths]
    FastDiv divFv;      // 16-byte vectors per x-row (W/4 fp32, W/8 bf16); d = 0 when W is not a multiple
    FastDiv divW;       // W                             [scalar path]
    FastDiv divH;       // H
};

struct FwdParams {
    Geometry g;
    const void* heat;
    float* coords;          // (R,3)
    float* stats;           // (R,2) or null
    const float* gt;        // (R,3) or null: no loss
    const float* vis;       // (R)
    const float* have_depth;// (B)
    float* loss;            // (1)
    // workspace
    int* row_count;         // (R) tickets, zero between launches
    int* done_rows;         // (1)
    float* row_loss;        // (R)
    float* partials;        // (R, maxslots, 8)
    int maxslots;
};

struct BwdParams {
    Geometry g;
    const void* heat;
    void* grad_heat;
    const float* coords;    // (R,3)
    const float* stats;     // (R,2)
    const float* grad_coords;   // (R,3) or null when the loss is fused
    const float* gt;        // fused loss inputs
    const float* vis;
    const float* have_depth;
    const float* grad_out;  // device scalar, or null: use grad_out_const
    float grad_out_const;
    float loss_scale;       // 1 / (3 * B_total * J): loss.py:50,52 (B_total may exceed this launch's slice)
};

struct FusedParams {
    FwdParams f;            // gt / vis / have_depth / loss are mandatory here
    void* grad_heat;        // d loss / d heat for upstream gradient 1
    uint2* xslots;          // (R, kMaxSplit, 8) {value, tag} pairs: partials traded between the S CTAs of a volume
    int* epoch;             // launch counter in the workspace; tag = epoch + 1
    int S;                  // CTAs per joint-volume
    int xc_chunks, xc_lag;  // K5c: chunks per CTA slice, chunks between the two passes
    float loss_scale;       // 1 / (3 * B * J)
    int debug_no_exchange;  // -DIHPR_TIMING_EXPERIMENTS builds only (IHPR_DEBUG_NOXCHG=1): skip the cross-CTA trade, WRONG results
};

// chunk range of persistent CTA `cta` of G: [cta*Gt/G, (cta+1)*Gt/G)
__device__ __forceinline__ uint64_t range_lo(uint64_t Gt, uint32_t G, uint32_t cta) { return (uint64_t)cta * Gt / G; }
// owner CTA of global chunk g (largest c with range_lo(c) <= g)
__device__ __forceinline__ uint32_t owner_of(uint64_t g, uint64_t Gt, uint32_t G) { return (uint32_t)(((g + 1) * G - 1) / Gt); }

void launch_fwd(const FwdParams& p, int dtype, bool vec_ok, int variant, int num_sms, cudaStream_t s);
int fwd_cluster_size(const Geometry& g, int dtype);      // K1c: CTAs per joint-volume for a small batch, 0 = use the persistent ring kernel
void launch_bwd(const BwdParams& p, int dtype, bool vec_ok, int variant, int num_sms, cudaStream_t s);
Geometry make_geometry(int B, int J, int D, int H, int W, int dtype, bool vec_ok, int variant);
int fused_split(const Geometry& g, int dtype);
cudaError_t launch_fused(const FusedParams& p, int dtype, int num_sms, cudaStream_t s);
struct FusedClusterPlan {
    int cluster;        // CTAs per joint-volume (cluster size); 0 = K5c does not apply
    int chunk_bytes;    // ring stage size
    int chunks;         // chunks per CTA slice
    int lag;            // chunks between pass 1 and pass 2 of the same chunk
};
FusedClusterPlan fused_cluster_plan(const Geometry& g, int dtype);
int fused_cluster_capacity(int dtype, int CS);
cudaError_t launch_fused_cluster(const FusedParams& p, int dtype, const FusedClusterPlan& pl, int nclusters, cudaStream_t s);
void launch_scale(void* grad, size_t n, int dtype, bool aligned, const float* grad_out, int num_sms, cudaStream_t s);
void launch_patches(const unsigned char* images, const int* sizes, int B, int Hs, int Ws, const double* trans, const int* do_flip,
                    const float* color_scale, const float* mean, const float* stdv, int out_h, int out_w, float* out, int channels_last,
                    cudaStream_t s);
void launch_joints(const double* joint_img, const double* joint_vis, const int* sizes, const double* trans, const double* scale, const int* do_flip,
                   const int* perm, int B, int J, int in_h, int in_w, int out_h, int out_w, int depth_dim, double bbox3d_depth, float* gt_coord,
                   float* gt_vis, cudaStream_t s);
void launch_l1_from_coords(const float* coords, const float* gt, const float* vis, const float* hd, int B, int J, float* loss, cudaStream_t s);
void launch_coords_post(const float* coords, const float* flipped, const int* perm, int B, int J, int D, int H, int W, const float* bbox,
                        const float* center, const float* focal, const float* princpt, float bbox3d_depth, int root, float* merged, float* pixel,
                        float* cam, cudaStream_t s);
const char* launch_head_fused(const void* x_nhwc, const void* w, const float* bias, int B, int K, int J, int D, int H, int W, float* coords,
                              float* stats, const float* gt, const float* vis, const float* have_depth, const float* grad_out, void* grad_heat,
                              float* dbias_part, int num_sms, cudaStream_t s);

size_t head_bwd_workspace_bytes(int B, int K, int J, int D, int H, int W);
const char* launch_head_bwd_params(const void* x_nhwc, const void* w, const float* bias, int B, int K, int J, int D, int H, int W, const float* coords,
                                   const float* stats, const float* gt, const float* vis, const float* have_depth, const float* grad_out,
                                   void* dx_nhwc, float* dweight, float* dbias, void* workspace, int num_sms, int pairs, int* launches, cudaStream_t s);

size_t deconv_workspace_bytes(int Cin, int Cout);
void launch_deconv_prepare(const void* weight, const float* gamma, const float* beta, const float* mean, const float* var, float eps, int Cin, int Cout,
                           void* workspace, int* launches, cudaStream_t s);
const char* launch_deconv_bn_relu(const void* x_nhwc, const void* prepared, int B, int Cin, int Cout, int Hin, int Win, void* y_nhwc, int num_sms, int cluster,
                                  int* launches, cudaStream_t s);
// training side of the deconv block (deconv_bn_relu.cu MODE kTrain / kDgrad, bn_train.cu)
void launch_deconv_relayout(const void* weight, int Cin, int Cout, void* wp_fwd, void* wp_dgrad, int* launches, cudaStream_t s);
const char* launch_deconv_train_fwd(const void* x_nhwc, const void* wp_fwd, int B, int Cin, int Cout, int Hin, int Win, void* y_raw_nhwc, float* stat_part,
                                    int* stat_rows, int num_sms, int* launches, cudaStream_t s);
const char* launch_deconv_dgrad(const void* dy_nhwc, const void* wp_dgrad, int B, int Cin, int Cout, int Hin, int Win, void* dx_nhwc, int num_sms, int* launches,
                                cudaStream_t s);
size_t deconv_wgrad_workspace_bytes();
const char* launch_deconv_wgrad(const void* x_nhwc, const void* dy_nhwc, int B, int Hin, int Win, float* dw, void* workspace, int num_sms, int cluster,
                                int* launches, cudaStream_t s);
int bn_bwd_rows(int num_sms);
void launch_bn_stat_finalize(const float* part, int rows, size_t n_per_channel, const float* gamma, const float* beta, float eps, float momentum,
                             float* running_mean, float* running_var, float* mean, float* rstd, float* scale, float* shift, int* launches, cudaStream_t s);
void launch_bn_relu_apply(const void* y, void* out, size_t n_pix, const float* scale, const float* shift, int num_sms, int* launches, cudaStream_t s);
void launch_bn_relu_bwd(const void* dout, const void* y, void* dy, size_t n_pix, const float* scale, const float* shift, const float* mean, const float* rstd,
                        float* dgamma, float* dbeta, float* part, float* cP, float* cQ, int num_sms, int* launches, cudaStream_t s);

}  // namespace ihpr

// K6: test-time post-processing of the (B, J, 3) soft-argmax result, on the device, in one launch.
//
// Follows (arithmetic restated, nothing copied):
//   main/test.py:67-76                       flip-test merge: x' = W - x - 1, swap left/right joints, average
//   common/utils/pose_utils.py:68-75         warp_coord_to_original: heat-map voxel -> image pixel / depth in mm
//   common/utils/pose_utils.py:14-20         pixel2cam: back-projection with focal length and principal point
//   data/Human36M/Human36M.py:226-228        root-joint alignment of the camera-space prediction
//
// The reference does the first step with ~4 + 2*pairs tiny torch launches per batch, then moves the coordinates to the host
// (test.py:93) and does the rest per sample in numpy.  Here one thread handles one (sample, joint); the root joint of the same
// sample is recomputed by every thread instead of exchanged (3 loads + ~12 flops, cheaper than a barrier).
#include "ihpr_common.cuh"

namespace ihpr {

struct PostParams {
    const float* coords;
    const float* flipped;      // nullable: no flip test
    const int* perm;           // nullable: identity; perm[j] = joint whose flipped-pass result lands on j
    const float* bbox;         // (B, 4) x, y, w, h of the crop in the original image
    const float* center;       // (B, 3) camera-space centre of the 3-D box; only [2] is used
    const float* focal;        // (B, 2)
    const float* princpt;      // (B, 2)
    float* merged;             // nullable
    float* pixel;              // nullable
    float* cam;                // nullable
    int B, J, root;
    float w, h, d, half_depth;
};

struct P3 {
    float x, y, z;
};

__device__ __forceinline__ P3 merged_at(const PostParams& p, int b, int j) {
    const float* c = p.coords + ((size_t)b * p.J + j) * 3;
    P3 r{c[0], c[1], c[2]};
    if (p.flipped) {
        const int src = p.perm ? p.perm[j] : j;
        const float* f = p.flipped + ((size_t)b * p.J + src) * 3;
        // (coord + flipped') / 2 with flipped'.x = W - x - 1, evaluated in the reference's order
        r.x = (r.x + (p.w - f[0] - 1.f)) / 2.f;
        r.y = (r.y + f[1]) / 2.f;
        r.z = (r.z + f[2]) / 2.f;
    }
    return r;
}

__device__ __forceinline__ P3 to_pixel(const PostParams& p, int b, P3 m) {
    const float* bb = p.bbox + (size_t)b * 4;
    P3 r;
    // explicit round-to-nearest mul / add: no FMA contraction, so the result does not depend on how the compiler schedules the
    // two inlined copies (the root joint's own aligned coordinate must come out as exactly 0) and follows numpy's un-fused order
    r.x = __fadd_rn(__fmul_rn(__fdiv_rn(m.x, p.w), bb[2]), bb[0]);
    r.y = __fadd_rn(__fmul_rn(__fdiv_rn(m.y, p.h), bb[3]), bb[1]);
    r.z = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fdiv_rn(m.z, p.d), 2.f), -1.f), p.half_depth), p.center[(size_t)b * 3 + 2]);
    return r;
}

__device__ __forceinline__ P3 to_cam(const PostParams& p, int b, P3 q) {
    const float* f = p.focal + (size_t)b * 2;
    const float* c = p.princpt + (size_t)b * 2;
    return P3{__fmul_rn(__fdiv_rn(__fadd_rn(q.x, -c[0]), f[0]), q.z), __fmul_rn(__fdiv_rn(__fadd_rn(q.y, -c[1]), f[1]), q.z), q.z};
}

__global__ void __launch_bounds__(128) coords_post_kernel(PostParams p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.B * p.J) return;
    const int b = i / p.J, j = i - b * p.J;
    const P3 m = merged_at(p, b, j);
    if (p.merged) {
        float* o = p.merged + (size_t)i * 3;
        o[0] = m.x, o[1] = m.y, o[2] = m.z;
    }
    if (!p.pixel && !p.cam) return;
    const P3 q = to_pixel(p, b, m);
    if (p.pixel) {
        float* o = p.pixel + (size_t)i * 3;
        o[0] = q.x, o[1] = q.y, o[2] = q.z;
    }
    if (!p.cam) return;
    P3 c = to_cam(p, b, q);
    if (p.root >= 0) {
        const P3 r = to_cam(p, b, to_pixel(p, b, merged_at(p, b, p.root)));
        c.x -= r.x, c.y -= r.y, c.z -= r.z;
    }
    float* o = p.cam + (size_t)i * 3;
    o[0] = c.x, o[1] = c.y, o[2] = c.z;
}

void launch_coords_post(const float* coords, const float* flipped, const int* perm, int B, int J, int D, int H, int W, const float* bbox,
                        const float* center, const float* focal, const float* princpt, float bbox3d_depth, int root, float* merged, float* pixel,
                        float* cam, cudaStream_t s) {
    PostParams p;
    p.coords = coords, p.flipped = flipped, p.perm = perm, p.bbox = bbox, p.center = center, p.focal = focal, p.princpt = princpt;
    p.merged = merged, p.pixel = pixel, p.cam = cam;
    p.B = B, p.J = J, p.root = root;
    p.w = (float)W, p.h = (float)H, p.d = (float)D;
    p.half_depth = bbox3d_depth * 0.5f;
    const int n = B * J;
    coords_post_kernel<<<(n + 127) / 128, 128, 0, s>>>(p);
}

// JointLocationLoss.forward on coordinates that already exist (common/nets/loss.py:49-52): the fused head (K3) produces the
// coordinates without a heat-map, the masked L1 mean is this one small launch.  One block, fixed summation order:
// thread t adds terms t, t + 256, ... in index order, then a fixed shared-memory tree -- bit-reproducible.
__global__ void __launch_bounds__(256) l1_from_coords_kernel(const float* __restrict__ coords, const float* __restrict__ gt, const float* __restrict__ vis,
                                                            const float* __restrict__ hd, int B, int J, float* __restrict__ loss) {
    __shared__ float part[256];
    const int n = B * J;
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) {
        const float* c = coords + (size_t)i * 3;
        const float* g = gt + (size_t)i * 3;
        const float v = vis[i], d = hd[i / J];
        acc += (fabsf(c[0] - g[0]) * v + fabsf(c[1] - g[1]) * v + fabsf(c[2] - g[2]) * v * d) / 3.f;
    }
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) part[threadIdx.x] += part[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss[0] = part[0] / (float)n;
}

void launch_l1_from_coords(const float* coords, const float* gt, const float* vis, const float* hd, int B, int J, float* loss, cudaStream_t s) {
    l1_from_coords_kernel<<<1, 256, 0, s>>>(coords, gt, vis, hd, B, J, loss);
}

}  // namespace ihpr

// K7 / K8: the per-sample work of the reference's DatasetLoader.__getitem__ (data/dataset.py:84-152) for a whole batch, on the device.
//
//   K7 patch_kernel   dataset.py:201-221 generate_patch_image (flip, cv2.warpAffine INTER_LINEAR, BGR->RGB, float32),
//                     :92-93 colour scale + clip, common/base.py:93-95 ToTensor + Normalize
//   K8 joints_kernel  dataset.py:96-135 joint flip / pair swap, trans_point2d (:259-262), depth normalisation, visibility test,
//                     input-space -> heat-map-space scaling; :148-149 float32 casts
//
// cv2.warpAffine on uint8 is integer arithmetic (OpenCV imgproc: inverse map in double, source coordinates in 10-bit fixed point
// rounded to 1/32 pixel, bilinear weights (32-a)(32-b)*32 summing to 2^15, constant zero border).  K7 restates exactly that, with
// explicit round-to-nearest double operations (no FMA contraction) where OpenCV rounds, so the patch is bit-identical to the
// reference's; the fp32 tail (colour scale, clip, (x - mean) / std) uses the same single operations as numpy / torch.
#include "ihpr_common.cuh"

namespace ihpr {

struct InvMap {
    double a00, a01, a02, a10, a11, a12;
};

// OpenCV's inversion of the forward 2x3 map inside warpAffine
__device__ __forceinline__ InvMap invert_affine(const double* m) {
    double d = __dsub_rn(__dmul_rn(m[0], m[4]), __dmul_rn(m[1], m[3]));
    d = d != 0.0 ? __ddiv_rn(1.0, d) : 0.0;
    InvMap r;
    r.a00 = __dmul_rn(m[4], d);
    r.a11 = __dmul_rn(m[0], d);
    r.a01 = __dmul_rn(m[1], -d);
    r.a10 = __dmul_rn(m[3], -d);
    r.a02 = __dsub_rn(__dmul_rn(-r.a00, m[2]), __dmul_rn(r.a01, m[5]));
    r.a12 = __dsub_rn(__dmul_rn(-r.a10, m[2]), __dmul_rn(r.a11, m[5]));
    return r;
}

struct PatchParams {
    const unsigned char* images;   // (B, Hs, Ws, 3) BGR
    const int* sizes;              // (B, 2) valid rows, cols
    const double* trans;           // (B, 6)
    const int* do_flip;            // (B)
    const float* color_scale;      // (B, 3), RGB order
    float* out;
    int B, Hs, Ws, out_h, out_w, channels_last;
    float mean[3], stdv[3];
};

constexpr int kAbBits = 10, kInterBits = 5;

__global__ void __launch_bounds__(256) patch_kernel(PatchParams p) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y, b = blockIdx.z;
    __shared__ InvMap inv;                                   // one inversion (a double division) per block, not per pixel
    if (threadIdx.x == 0 && threadIdx.y == 0) inv = invert_affine(p.trans + (size_t)b * 6);
    __syncthreads();
    if (x >= p.out_w || y >= p.out_h) return;
    const InvMap im = inv;
    const double ab = (double)(1 << kAbBits);
    const int round_delta = (1 << kAbBits) / (1 << kInterBits) / 2;
    const int adelta = __double2int_rn(__dmul_rn(__dmul_rn(im.a00, (double)x), ab));
    const int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(im.a10, (double)x), ab));
    const int x0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(im.a01, (double)y), im.a02), ab)) + round_delta;
    const int y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(im.a11, (double)y), im.a12), ab)) + round_delta;
    const int X = (x0 + adelta) >> (kAbBits - kInterBits), Y = (y0 + bdelta) >> (kAbBits - kInterBits);
    const int sx = X >> kInterBits, sy = Y >> kInterBits, ax = X & 31, ay = Y & 31;
    const int h = p.sizes[b * 2], w = p.sizes[b * 2 + 1];
    const bool flip = p.do_flip[b] != 0;
    const unsigned char* img = p.images + (size_t)b * p.Hs * p.Ws * 3;

    int acc[3] = {0, 0, 0};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int yy = sy + (t >> 1), xx = sx + (t & 1);
        const int wy = (t >> 1) ? ay : 32 - ay, wx = (t & 1) ? ax : 32 - ax;
        const int wgt = wy * wx * 32;
        if (wgt == 0 || yy < 0 || yy >= h || xx < 0 || xx >= w) continue;       // constant zero border
        const unsigned char* px = img + ((size_t)yy * p.Ws + (flip ? w - 1 - xx : xx)) * 3;
        acc[0] += px[0] * wgt, acc[1] += px[1] * wgt, acc[2] += px[2] * wgt;
    }
    const float* cs = p.color_scale + (size_t)b * 3;
    const size_t plane = (size_t)p.out_h * p.out_w;
#pragma unroll
    for (int c = 0; c < 3; ++c) {                                              // c: RGB channel of the output = BGR channel 2 - c
        const int v = (acc[2 - c] + (1 << 14)) >> 15;
        float f = __fmul_rn((float)v, cs[c]);
        f = fminf(fmaxf(f, 0.f), 255.f);
        f = __fdiv_rn(__fsub_rn(f, p.mean[c]), p.stdv[c]);
        const size_t o = p.channels_last ? (((size_t)b * p.out_h + y) * p.out_w + x) * 3 + c : ((size_t)b * 3 + c) * plane + (size_t)y * p.out_w + x;
        p.out[o] = f;
    }
}

void launch_patches(const unsigned char* images, const int* sizes, int B, int Hs, int Ws, const double* trans, const int* do_flip,
                    const float* color_scale, const float* mean, const float* stdv, int out_h, int out_w, float* out, int channels_last,
                    cudaStream_t s) {
    PatchParams p;
    p.images = images, p.sizes = sizes, p.trans = trans, p.do_flip = do_flip, p.color_scale = color_scale, p.out = out;
    p.B = B, p.Hs = Hs, p.Ws = Ws, p.out_h = out_h, p.out_w = out_w, p.channels_last = channels_last;
    for (int c = 0; c < 3; ++c) p.mean[c] = mean[c], p.stdv[c] = stdv[c];
    dim3 grid((out_w + 31) / 32, (out_h + 7) / 8, B), block(32, 8);
    patch_kernel<<<grid, block, 0, s>>>(p);
}

struct JointParams {
    const double* joint_img;   // (B, J, 3) x, y in source-image pixels, z root-relative depth (mm)
    const double* joint_vis;   // (B, J)
    const int* sizes;          // (B, 2)
    const double* trans;       // (B, 6)
    const double* scale;       // (B)
    const int* do_flip;        // (B)
    const int* perm;           // (J) nullable
    float* gt_coord;           // (B, J, 3)
    float* gt_vis;             // (B, J)
    int B, J;
    double in_h, in_w, out_h, out_w, depth_dim, half_depth;
};

__global__ void __launch_bounds__(128) joints_kernel(JointParams p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.B * p.J) return;
    const int b = i / p.J, j = i - b * p.J;
    const bool flip = p.do_flip[b] != 0;
    const int src = (flip && p.perm) ? p.perm[j] : j;
    const double* q = p.joint_img + ((size_t)b * p.J + src) * 3;
    double x = q[0], y = q[1], z = q[2], vis = p.joint_vis[(size_t)b * p.J + src];
    if (flip) x = (double)p.sizes[b * 2 + 1] - x - 1.0;
    const double* t = p.trans + (size_t)b * 6;
    const double px = t[0] * x + t[1] * y + t[2], py = t[3] * x + t[4] * y + t[5];
    z = z / (p.half_depth * p.scale[b]);
    z = (z + 1.0) / 2.0;
    const bool inside = px >= 0 && px < p.in_w && py >= 0 && py < p.in_h && z >= 0 && z < 1;
    vis *= inside ? 1.0 : 0.0;
    float* o = p.gt_coord + (size_t)i * 3;
    o[0] = (float)(px / p.in_w * p.out_w);
    o[1] = (float)(py / p.in_h * p.out_h);
    o[2] = (float)(z * p.depth_dim);
    p.gt_vis[i] = vis > 0 ? 1.f : 0.f;
}

void launch_joints(const double* joint_img, const double* joint_vis, const int* sizes, const double* trans, const double* scale, const int* do_flip,
                   const int* perm, int B, int J, int in_h, int in_w, int out_h, int out_w, int depth_dim, double bbox3d_depth, float* gt_coord,
                   float* gt_vis, cudaStream_t s) {
    JointParams p;
    p.joint_img = joint_img, p.joint_vis = joint_vis, p.sizes = sizes, p.trans = trans, p.scale = scale, p.do_flip = do_flip, p.perm = perm;
    p.gt_coord = gt_coord, p.gt_vis = gt_vis, p.B = B, p.J = J;
    p.in_h = in_h, p.in_w = in_w, p.out_h = out_h, p.out_w = out_w, p.depth_dim = depth_dim, p.half_depth = bbox3d_depth / 2.;
    const int n = B * J;
    joints_kernel<<<(n + 127) / 128, 128, 0, s>>>(p);
}

}  // namespace ihpr

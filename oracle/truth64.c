/*
 * oracle/truth64.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Independent fp64 CPU restatement of the reference's integral-regression path:
 *   soft_argmax            /root/reference/common/nets/loss.py:13-34
 *   JointLocationLoss      /root/reference/common/nets/loss.py:36-52
 * and of the analytic gradient that torch autograd derives for it.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call this.
 * The shipped CUDA path never links, imports or falls back to it.
 *
 * Pinning: tests/test_oracle.py checks this file and oracle/soft_argmax_ref.py against
 * tests/golden/ (npz files), which oracle/make_golden.py produced by executing the unmodified
 * reference functions (copied to a scratch dir, two-attribute CPU shim) in the build
 * container.
 *
 * Layout: heat is (B, J*D, H, W) contiguous, channel c = j*D + d  (loss.py:16,18),
 * so joint-volume r = b*J + j is the contiguous run heat[r*N .. (r+1)*N), N = D*H*W,
 * voxel i -> x = i % W, y = (i / W) % H, z = i / (W*H).  Output order is (x, y, z)
 * (loss.py:32), each coordinate 0-based after the reference's "1-based arange then -1"
 * (loss.py:24-30).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

/* Forward: coords (R,3), m (R), l (R) with l = sum exp(h - m).   loss.py:16-32 */
void ihpr_oracle_fwd_f64(const float *heat, long R, int D, int H, int W,
                         double *coords, double *mx, double *lsum)
{
    const long N = (long)D * H * W;
    for (long r = 0; r < R; ++r) {
        const float *h = heat + r * N;
        double m = -INFINITY;
        for (long i = 0; i < N; ++i) if ((double)h[i] > m) m = (double)h[i];   /* softmax max-subtraction (loss.py:17) */
        double l = 0.0, sx = 0.0, sy = 0.0, sz = 0.0;
        long i = 0;
        for (int z = 0; z < D; ++z)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x, ++i) {
                    double p = exp((double)h[i] - m);
                    l += p;                  /* softmax denominator */
                    sx += p * x;             /* accu_x * arange, loss.py:20,24,28 (0-based) */
                    sy += p * y;             /* loss.py:21,25,29 */
                    sz += p * z;             /* loss.py:22,26,30 */
                }
        coords[3 * r + 0] = sx / l;
        coords[3 * r + 1] = sy / l;
        coords[3 * r + 2] = sz / l;
        if (mx) mx[r] = m;
        if (lsum) lsum[r] = l;
    }
}

/* Backward of soft_argmax alone: dh_i = p_i * sum_c g_c (c(i) - coord_c).   SURVEY.md 3.4 */
void ihpr_oracle_bwd_f64(const float *heat, long R, int D, int H, int W,
                         const double *coords, const double *mx, const double *lsum,
                         const double *gcoords, double *gheat)
{
    const long N = (long)D * H * W;
    for (long r = 0; r < R; ++r) {
        const float *h = heat + r * N;
        double *g = gheat + r * N;
        const double gx = gcoords[3 * r], gy = gcoords[3 * r + 1], gz = gcoords[3 * r + 2];
        const double a = -(gx * coords[3 * r] + gy * coords[3 * r + 1] + gz * coords[3 * r + 2]);
        const double m = mx[r], il = 1.0 / lsum[r];
        long i = 0;
        for (int z = 0; z < D; ++z)
            for (int y = 0; y < H; ++y)
                for (int x = 0; x < W; ++x, ++i)
                    g[i] = exp((double)h[i] - m) * il * (a + gx * x + gy * y + gz * z);
    }
}

/* JointLocationLoss forward (loss.py:49-52) and d loss / d coords (sign(0) = 0 as in torch.abs).
 * gt (B,J,3), vis (B,J,1), have_depth (B,1).  Returns the scalar loss; gcoords may be NULL. */
double ihpr_oracle_loss_f64(const double *coords, const float *gt, const float *vis,
                            const float *have_depth, long B, long J, double grad_out,
                            double *gcoords)
{
    double acc = 0.0;
    const double inv = 1.0 / (3.0 * (double)(B * J));
    for (long b = 0; b < B; ++b)
        for (long j = 0; j < J; ++j) {
            const long r = b * J + j;
            const double v = vis[r], hd = have_depth[b];
            const double w[3] = {1.0, 1.0, hd};
            double s = 0.0;
            for (int c = 0; c < 3; ++c) {
                const double d = coords[3 * r + c] - (double)gt[3 * r + c];
                s += fabs(d) * v * w[c];
                if (gcoords)
                    gcoords[3 * r + c] = grad_out * ((d > 0) - (d < 0)) * v * w[c] * inv;
            }
            acc += s / 3.0;
        }
    return acc / (double)(B * J);
}

/* fp32 scalar port (same algorithm, float arithmetic) over joint-volumes [r0, r1) of R = B*J;
 * single-threaded: oracle/truth.py fans row ranges out over Python threads (ctypes drops the GIL).
 * Writes coords (R,3), row_loss (R) = per-(b,j) term of loss.py:50 and, if gheat != NULL, d loss/d heat.
 * Used only as the alternative cpu_baseline "port" leg in bench.py. */
void ihpr_oracle_fwd_bwd_f32(const float *heat, long R, long J, long r0, long r1, int D, int H, int W,
                             const float *gt, const float *vis, const float *have_depth,
                             float *coords, float *row_loss, float *gheat)
{
    const long N = (long)D * H * W;
    const float inv = 1.0f / (3.0f * (float)R);
    for (long r = r0; r < r1; ++r) {
        const float *h = heat + r * N;
        float m = -INFINITY;
        for (long i = 0; i < N; ++i) m = h[i] > m ? h[i] : m;
        float l = 0.f, sx = 0.f, sy = 0.f, sz = 0.f;
        long i = 0;
        for (int z = 0; z < D; ++z)
            for (int y = 0; y < H; ++y) {
                float rl = 0.f, rx = 0.f;
                for (int x = 0; x < W; ++x, ++i) {
                    float p = expf(h[i] - m);
                    rl += p;
                    rx += p * (float)x;
                }
                l += rl; sx += rx; sy += rl * (float)y; sz += rl * (float)z;
            }
        const float c[3] = {sx / l, sy / l, sz / l};
        const float v = vis[r], hd = have_depth[r / J];
        const float w[3] = {1.f, 1.f, hd};
        float g[3], s = 0.f;
        for (int k = 0; k < 3; ++k) {
            const float d = c[k] - gt[3 * r + k];
            s += fabsf(d) * v * w[k];
            g[k] = (float)((d > 0) - (d < 0)) * v * w[k] * inv;
            coords[3 * r + k] = c[k];
        }
        row_loss[r] = s / 3.0f;
        if (gheat) {
            const float a = -(g[0] * c[0] + g[1] * c[1] + g[2] * c[2]), il = 1.0f / l;
            float *gh = gheat + r * N;
            i = 0;
            for (int z = 0; z < D; ++z)
                for (int y = 0; y < H; ++y) {
                    const float t0 = a + g[1] * (float)y + g[2] * (float)z;
                    for (int x = 0; x < W; ++x, ++i)
                        gh[i] = expf(h[i] - m) * il * (t0 + g[0] * (float)x);
                }
        }
    }
}

"""oracle/make_golden.py -- produces tests/golden/*.npz by EXECUTING THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference):  python -m oracle.make_golden

For every case the unmodified reference ``soft_argmax`` + ``JointLocationLoss``
(/root/reference/common/nets/loss.py:13-52, loaded by oracle/load_reference.py) is run forward and
backward in fp32 (what the GPU kernels are compared with) and again in fp64 (the reference's own
ops on double input: the "truth" column).  While doing so the two oracle restatements are pinned:
oracle/soft_argmax_ref.py must be bit-identical to the reference in fp32, oracle/truth64.c must
agree with the reference-in-fp64 to 1e-9.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import inputs, truth                     # noqa: E402
from oracle.load_reference import Reference          # noqa: E402
from oracle.soft_argmax_ref import ref_fwd_bwd       # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
GRAD_STRIDE = 61          # big cases keep every 61st gradient element (+ checksums) to stay small

# name, dist, B, J, D, H, W, seed, vis_mode, hd_mode, keep_full
CASES = [
    ("randn1_b2j3_d8h8w8",      "randn1", 2, 3, 8, 8, 8, 1, "ones", "ones", True),
    ("randn3_b1j17_d16h16w16",  "randn3", 1, 17, 16, 16, 16, 2, "ones", "ones", True),
    ("init_b2j18_d8h16w12",     "init",   2, 18, 8, 16, 12, 3, "ones", "ones", True),
    ("odd_b1j2_d5h7w9",         "randn1", 1, 2, 5, 7, 9, 4, "ones", "ones", True),
    ("blobs_b1j4_d16h16w16",    "blobs",  1, 4, 16, 16, 16, 5, "ones", "ones", True),
    ("large_b1j2_d8h8w8",       "large",  1, 2, 8, 8, 8, 6, "ones", "ones", True),
    ("shifted_b1j2_d8h8w8",     "shifted", 1, 2, 8, 8, 8, 7, "ones", "ones", True),
    ("mask_b3j5_d8h8w8",        "randn3", 3, 5, 8, 8, 8, 8, "rand", "alt", True),
    ("nodepth_b2j16_d8h8w8",    "randn1", 2, 16, 8, 8, 8, 9, "rand", "zeros", True),
    ("full_b1j2_d64h64w64",     "randn1", 1, 2, 64, 64, 64, 10, "ones", "ones", False),
    ("fullblobs_b1j2_d64h64w64", "blobs", 1, 2, 64, 64, 64, 11, "ones", "ones", False),
    ("d32_b2j17_d32h64w64",     "randn3", 2, 17, 32, 64, 64, 12, "rand", "alt", False),
]


def main():
    ref = Reference()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(1)        # fixed summation order for the committed fp32 numbers
    for name, dist, B, J, D, H, W, seed, vis_mode, hd_mode, keep_full in CASES:
        heat = inputs.make_heat(dist, B, J, D, H, W, seed)
        gt, vis, hd = inputs.make_targets(B, J, D, H, W, seed, vis_mode, hd_mode)
        th, tgt, tvis, thd = (torch.from_numpy(a) for a in (heat, gt, vis, hd))

        loss32, coords32, grad32 = ref.fwd_bwd(th, tgt, tvis, thd)
        loss64, coords64, grad64 = ref.fwd_bwd(th.double(), tgt.double(), tvis.double(), thd.double())

        # pin the torch restatement: identical ATen sequence -> bit-identical
        l2, c2, g2 = ref_fwd_bwd(th, tgt, tvis, thd)
        assert torch.equal(l2, loss32) and torch.equal(c2, coords32) and torch.equal(g2, grad32), name
        # pin the fp64 C restatement against the reference run in fp64
        lo, co, go = truth.fwd_bwd_f64(heat, gt, vis, hd)
        assert abs(lo - loss64.item()) <= 1e-9 * max(1.0, abs(loss64.item())), name
        assert np.abs(co - coords64.numpy()).max() <= 1e-9 * max(D, H, W), name
        gmax = np.abs(grad64.numpy()).max()
        assert np.abs(go - grad64.numpy()).max() <= 1e-9 * max(gmax, 1e-30), (name, np.abs(go - grad64.numpy()).max(), gmax)

        out = dict(dist=dist, B=B, J=J, D=D, H=H, W=W, seed=seed, vis_mode=vis_mode, hd_mode=hd_mode,
                   gt=gt, vis=vis, have_depth=hd,
                   ref32_loss=np.float32(loss32.item()), ref32_coords=coords32.numpy(),
                   ref64_loss=np.float64(loss64.item()), ref64_coords=coords64.numpy(),
                   heat_sha=np.frombuffer(__import__("hashlib").sha256(heat.tobytes()).digest(), dtype=np.uint8),
                   torch_version=str(torch.__version__))
        g32 = grad32.numpy().reshape(-1); g64 = grad64.numpy().reshape(-1)
        if keep_full:
            out.update(heat=heat, ref32_grad=g32.reshape(heat.shape), ref64_grad=g64.reshape(heat.shape))
        else:
            out.update(grad_stride=GRAD_STRIDE, ref32_grad_sub=g32[::GRAD_STRIDE].copy(),
                       ref64_grad_sub=g64[::GRAD_STRIDE].copy(),
                       ref64_grad_abssum=np.float64(np.abs(g64).sum()), ref64_grad_max=np.float64(gmax))
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        err_c = np.abs(coords32.numpy() - coords64.numpy()).max()
        err_g = np.abs(g32 - g64).max() / max(gmax, 1e-30)
        print(f"{name:28s} loss32={loss32.item():.6f} |coords32-64|={err_c:.2e} grad relerr32-64={err_g:.2e} "
              f"{os.path.getsize(path)/1024:.0f} KiB")


# name, B, J, D, H, W, seed, flip, pairs, root
POST_CASES = [
    ("post_h36m_b4j18_flip", 4, 18, 64, 64, 64, 21, True, ((1, 4), (2, 5), (3, 6), (14, 11), (15, 12), (16, 13)), 0),
    ("post_noflip_b3j17",    3, 17, 64, 64, 64, 22, False, (), 0),
    ("post_odd_b2j5_d16h24w40_chain", 2, 5, 16, 24, 40, 23, True, ((0, 1), (1, 2)), 3),
    ("post_noroot_b1j4",     1, 4, 32, 64, 64, 24, True, ((0, 3),), -1),
]


def make_post_goldens():
    """tests/golden/post_*.npz: the reference's own warp_coord_to_original / pixel2cam (common/utils/pose_utils.py:68-75,14-20)
    chained per sample as data/Human36M/Human36M.py:203-228 does; the flip merge (inline in main/test.py:73-76, not importable)
    is oracle.coords_post_ref.flip_merge.  Pins oracle.coords_post_ref bit for bit while doing so."""
    from oracle import coords_post_ref as cp
    ref = Reference()
    from utils.pose_utils import pixel2cam, warp_coord_to_original            # the reference's functions (scratch copy)
    for name, B, J, D, H, W, seed, flip, pairs, root in POST_CASES:
        coords, flipped, bbox, center, f, c = cp.make_inputs(B, J, D, H, W, seed, flip)
        ref.set_shape(D, H, W)
        depth = float(ref.cfg.bbox_3d_shape[0])
        merged = coords if flipped is None else cp.flip_merge(torch.from_numpy(coords), torch.from_numpy(flipped), W, pairs).numpy()
        pix = np.zeros_like(merged)
        cam = np.zeros(merged.shape, np.float64)
        for n in range(B):
            p2 = merged[n].copy()
            p2[:, 0], p2[:, 1], p2[:, 2] = warp_coord_to_original(p2, bbox[n], center[n])
            p3 = np.zeros((J, 3))
            p3[:, 0], p3[:, 1], p3[:, 2] = pixel2cam(p2, f[n], c[n])
            if root >= 0:
                p3 = p3 - p3[root]
            pix[n], cam[n] = p2, p3
        m2, pix2, cam2 = cp.post_process(coords, flipped, pairs, bbox, center, f, c, root, D, (H, W), depth)
        assert np.array_equal(m2, merged) and np.array_equal(pix2, pix) and np.array_equal(cam2, cam), name
        out = dict(B=B, J=J, D=D, H=H, W=W, seed=seed, flip=flip, pairs=np.array(pairs, np.int32).reshape(-1, 2), root=root,
                   bbox_3d_depth=depth, coords=coords, bbox=bbox, center_cam=center, f=f, c=c, merged=merged, pixel=pix, cam=cam)
        if flipped is not None:
            out["flipped"] = flipped
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name:34s} |cam|max={np.abs(cam).max():.1f} {os.path.getsize(path)/1024:.1f} KiB")


# name, image h, w, J, input_shape, output_shape, depth_dim, seeds (one sample each), flip pairs
AUG_CASES = [
    ("aug_h36m_like", 500, 520, 18, (256, 256), (64, 64), 64, (31, 32, 33, 34, 35, 36), ((1, 4), (2, 5), (3, 6), (14, 11), (15, 12), (16, 13))),
    ("aug_rect_small", 240, 360, 5, (96, 128), (24, 32), 16, (41, 42, 43, 44), ((0, 1), (1, 2))),
]


def make_aug_goldens():
    """tests/golden/aug_*.npz by running the reference's own DatasetLoader.__getitem__ (data/dataset.py:48-152, which calls the
    real cv2.warpAffine / getAffineTransform and torchvision's ToTensor + Normalize as common/base.py:93-95 builds them) on
    seeded synthetic images written as PNG.  utils/vis.py needs matplotlib (absent): a placeholder module with the two names
    dataset.py imports is registered first; no reference line is edited.  Pins oracle/augment_ref.py: patches bit-identical."""
    import hashlib
    import random
    import types
    import cv2
    import torchvision.transforms as transforms
    from oracle import augment_ref as ar
    ref = Reference()
    vis_stub = types.ModuleType("utils.vis")
    vis_stub.vis_keypoints = vis_stub.vis_3d_skeleton = lambda *a, **k: None
    sys.modules.setdefault("utils.vis", vis_stub)
    import dataset as ref_dataset                                   # /root/reference/data/dataset.py (scratch copy)
    cfg = ref.cfg
    for name, h, w, J, in_shape, out_shape, depth_dim, seeds, pairs in AUG_CASES:
        cfg.input_shape, cfg.output_shape, cfg.depth_dim = in_shape, out_shape, depth_dim
        depth = float(cfg.bbox_3d_shape[0])
        tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize(mean=cfg.pixel_mean, std=cfg.pixel_std)])
        out = dict(h=h, w=w, J=J, input_shape=np.array(in_shape), output_shape=np.array(out_shape), depth_dim=depth_dim, seeds=np.array(seeds),
                   pairs=np.array(pairs, np.int32).reshape(-1, 2), bbox_3d_depth=depth, pixel_mean=np.array(cfg.pixel_mean), pixel_std=np.array(cfg.pixel_std),
                   cv2_version=str(cv2.__version__))
        rows = []
        for seed in seeds:
            img = ar.synthetic_image(h, w, seed)
            bbox, joints, vis = ar.synthetic_annotation(h, w, J, seed)
            path = os.path.join(ref.tmp, "img_%d.png" % seed)
            cv2.imwrite(path, img)

            class Db:
                joint_num, skeleton, lr_skeleton, flip_pairs, joints_have_depth = J, (), (), pairs, True

                def load_data(self):
                    return [dict(img_path=path, bbox=bbox.copy(), joint_img=joints.copy(), joint_vis=vis.copy())]

            loader = ref_dataset.DatasetLoader(Db(), True, tf)
            np.random.seed(seed); random.seed(seed)
            aug = ref_dataset.get_aug_config()                      # the draw __getitem__ is about to make
            np.random.seed(seed); random.seed(seed)
            assert ar.get_aug_config() == aug, name
            np.random.seed(seed); random.seed(seed)
            r_img, r_joint, r_vis, r_hd = loader[0]
            scale, rot, do_flip, color_scale = aug
            _, r_trans = ref_dataset.generate_patch_image(img[:1000, :1000], bbox, do_flip, scale, rot)
            o_img, o_joint, o_vis, o_trans = ar.get_item(img, bbox, joints, vis, pairs, aug, in_shape, out_shape, depth_dim, depth,
                                                         cfg.pixel_mean, cfg.pixel_std)
            r_img = r_img.numpy()
            assert np.abs(o_trans - r_trans).max() <= 1e-9, (name, seed, np.abs(o_trans - r_trans).max())
            assert np.array_equal(o_img, r_img), (name, seed, np.abs(o_img - r_img).max(), (o_img != r_img).mean())
            assert np.abs(o_joint - r_joint).max() <= 1e-4 and np.array_equal(o_vis, r_vis), (name, seed)
            rows.append(dict(aug=np.array([scale, rot, float(do_flip)] + list(color_scale)), trans=r_trans, joint=r_joint, vis=r_vis, hd=r_hd,
                             img_sha=np.frombuffer(hashlib.sha256(r_img.tobytes()).digest(), np.uint8), img_sub=r_img.reshape(-1)[::97].copy(),
                             src_sha=np.frombuffer(hashlib.sha256(img.tobytes()).digest(), np.uint8)))
            print(f"{name:16s} seed {seed}: scale {scale:.3f} rot {rot:6.2f} flip {int(do_flip)}  vis {int(r_vis.sum())}/{J}")
        for k in rows[0]:
            out[k] = np.stack([r[k] for r in rows])
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name:34s} {os.path.getsize(path)/1024:.1f} KiB")


DECONV_CASES = (("deconv_block_train_b2_h16w16", 41, 2, 16, 16), ("deconv_block_train_b1_h8w32", 42, 1, 8, 32))


def make_deconv_goldens():
    """Row N1 (training side): the third block of the reference's own HeadNet.deconv_layers (main/model.py:22-38,
    ConvTranspose2d(256, 256, k4 s2 p1) + BatchNorm2d + ReLU) run in TRAINING mode in fp64 on seeded inputs, forward and autograd
    backward; the numpy restatement oracle/deconv_block_ref.py is asserted against it while the fixture is written.  The fixture keeps
    the seed (inputs are regenerated by deconv_block_ref.problem), the per-channel quantities in full and strided samples of the big tensors."""
    import torch
    import torchvision.models.resnet as tvr
    from oracle import deconv_block_ref as R
    if not hasattr(tvr, "model_zoo"):
        tvr.model_zoo, tvr.model_urls = None, {}
    ref = Reference()
    cwd = os.getcwd()
    os.chdir(os.path.join(ref.tmp, "ref", "main"))
    try:
        import model as ref_model
    finally:
        os.chdir(cwd)
    ref.cfg.depth_dim = 64
    head = ref_model.HeadNet(18).double()
    block = head.deconv_layers[6:9]                      # ConvTranspose2d(256, 256), BatchNorm2d(256), ReLU
    conv, bn = block[0], block[1]
    assert isinstance(conv, torch.nn.ConvTranspose2d) and conv.in_channels == 256 and conv.out_channels == 256 and isinstance(bn, torch.nn.BatchNorm2d)
    for name, seed, B, H, W in DECONV_CASES:
        x, w, gamma, beta, rm, rv, dout = R.problem(seed, B, 256, 256, H, W)
        with torch.no_grad():
            conv.weight.copy_(torch.from_numpy(w))
            bn.weight.copy_(torch.from_numpy(gamma))
            bn.bias.copy_(torch.from_numpy(beta))
            bn.running_mean.copy_(torch.from_numpy(rm))
            bn.running_var.copy_(torch.from_numpy(rv))
        block.train()
        for p_ in block.parameters():
            p_.grad = None
        xt = torch.from_numpy(x).requires_grad_(True)
        out = block(xt)
        out.backward(torch.from_numpy(dout))
        got = {"out": out.detach().numpy(), "dx": xt.grad.numpy(), "dw": conv.weight.grad.numpy(), "dgamma": bn.weight.grad.numpy(),
               "dbeta": bn.bias.grad.numpy(), "running_mean": bn.running_mean.numpy().copy(), "running_var": bn.running_var.numpy().copy()}
        block.eval()
        with torch.no_grad():
            got["out_eval"] = block(torch.from_numpy(x)).numpy()       # with the UPDATED running statistics
        # the numpy restatement against the reference's modules
        f = R.forward_train(x, w, gamma, beta, rm, rv)
        b = R.backward_train(x, w, gamma, beta, dout)
        e = R.forward_eval(x, w, gamma, beta, f["running_mean"], f["running_var"])
        for key, val in (("out", f["out"]), ("running_mean", f["running_mean"]), ("running_var", f["running_var"]), ("dx", b["dx"]), ("dw", b["dw"]),
                         ("dgamma", b["dgamma"]), ("dbeta", b["dbeta"]), ("out_eval", e)):
            err = np.abs(val - got[key]).max() / max(np.abs(got[key]).max(), 1e-30)
            assert err <= 1e-10, (name, key, err)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, seed=seed, shape=np.array([B, 256, 256, H, W]), mean=f["mean"], var=f["var"],
                            running_mean=got["running_mean"], running_var=got["running_var"], dgamma=got["dgamma"], dbeta=got["dbeta"],
                            out_sub=got["out"][:, ::8, ::4, ::4], out_eval_sub=got["out_eval"][:, ::8, ::4, ::4], dx_sub=got["dx"][:, ::8, ::2, ::2],
                            dw_sub=got["dw"][::16, ::16], out_sum=got["out"].sum(axis=(0, 2, 3)), dx_sum=got["dx"].sum(axis=(0, 2, 3)),
                            dw_tapsum=got["dw"].sum(axis=(0, 1)))
        print(f"{name:34s} {os.path.getsize(path)/1024:.1f} KiB")


def dump_reference_state_keys():
    """state_dict keys + shapes of the reference's own get_pose_net (main/model.py:105-114) for ResNet-50 / J=18,
    as a fixture for the checkpoint-compatibility test of ihpr_b200.model (the reference's resnet.py imports
    `model_zoo, model_urls` which torchvision >= 0.13 no longer exports: two placeholder attributes let it import)."""
    import torchvision.models.resnet as tvr
    if not hasattr(tvr, "model_zoo"):
        tvr.model_zoo, tvr.model_urls = None, {}
    ref = Reference()
    cwd = os.getcwd()
    os.chdir(os.path.join(ref.tmp, "ref", "main"))
    try:
        import model as ref_model
    finally:
        os.chdir(cwd)
    ref.cfg.resnet_type, ref.cfg.depth_dim = 50, 64
    net = ref_model.get_pose_net(ref.cfg, False, 18)
    path = os.path.join(GOLDEN_DIR, "reference_state_keys_r50_j18.txt")
    with open(path, "w") as f:
        for k, v in net.state_dict().items():
            f.write("module.%s %s\n" % (k, "x".join(str(d) for d in v.shape)))
    print("wrote", path)


if __name__ == "__main__":
    if "--post" in sys.argv:                 # only the post-processing fixtures
        make_post_goldens()
        sys.exit(0)
    if "--aug" in sys.argv:                  # only the augmentation fixtures
        make_aug_goldens()
        sys.exit(0)
    if "--deconv" in sys.argv:               # only the deconv-block fixtures (row N1, training side)
        make_deconv_goldens()
        sys.exit(0)
    if "--keys" not in sys.argv:
        main()
        make_post_goldens()
        make_aug_goldens()
        make_deconv_goldens()
    dump_reference_state_keys()

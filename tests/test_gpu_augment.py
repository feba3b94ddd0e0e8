"""GPU (B200): K7 / K8, the device-side sample preparation (ihpr_augment_patches / ihpr_augment_joints), against
  (1) tests/golden/aug_*.npz -- what the reference's own DatasetLoader.__getitem__ (real cv2) returned for seeded images,
  (2) oracle/augment_ref.py on further seeded cases (mixed image sizes in one padded batch, no augmentation, channels_last).

Bar: the normalised patch is BIT-IDENTICAL (the warp is integer arithmetic, the fp32 tail single IEEE operations);
joint coordinates |a-b| <= 1e-5 * max(|b|, 1) (fp64 on both sides, cast to fp32 at the end); visibility flags equal.
"""
import hashlib

import numpy as np
import pytest
import torch

from conftest import aug_golden_names, load_aug_golden
from oracle import augment_ref as ar

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def joints_close(a, b):
    return bool((np.abs(a - b) <= 1e-5 * np.maximum(np.abs(b), 1.0)).all())


@pytest.mark.parametrize("name", aug_golden_names())
def test_augment_golden(name, dev):
    import ihpr_b200
    g = load_aug_golden(name)
    n = len(g["seeds"])
    imgs = np.stack([ar.synthetic_image(g["h"], g["w"], int(s)) for s in g["seeds"]])
    ann = [ar.synthetic_annotation(g["h"], g["w"], g["J"], int(s)) for s in g["seeds"]]
    out, coord, vis = ihpr_b200.augment_batch(torch.from_numpy(imgs).to(dev), [[g["h"], g["w"]]] * n, [a[0] for a in ann],
                                              np.stack([a[1] for a in ann]), np.stack([a[2] for a in ann]), g["augs"], flip_pairs=g["pairs"],
                                              input_shape=tuple(g["input_shape"]), output_shape=tuple(g["output_shape"]), depth_dim=g["depth_dim"],
                                              bbox_3d_depth=float(g["bbox_3d_depth"]), pixel_mean=g["pixel_mean"], pixel_std=g["pixel_std"])
    torch.cuda.synchronize()
    out, coord, vis = out.cpu().numpy(), coord.cpu().numpy(), vis.cpu().numpy()
    assert out.shape == (n, 3) + tuple(g["input_shape"]) and out.dtype == np.float32
    for i in range(n):
        assert np.array_equal(out[i].reshape(-1)[::97], g["img_sub"][i]), (name, i)
        assert np.array_equal(sha(out[i]), g["img_sha"][i]), (name, i)
        assert joints_close(coord[i], g["joint"][i]), (name, i, np.abs(coord[i] - g["joint"][i]).max())
        assert np.array_equal(vis[i], g["vis"][i]), (name, i)


@pytest.mark.parametrize("channels_last", [False, True])
def test_augment_mixed_sizes_vs_oracle(channels_last, dev):
    """Images of different sizes zero-padded into one batch; rotation / flip / scale from seeded get_aug_config draws plus one
    un-augmented sample (test-time path, dataset.py:88); patch partly outside the image on purpose (constant border)."""
    import random
    import ihpr_b200
    from ihpr_b200 import data
    shapes = [(120, 200), (333, 97), (64, 64), (250, 251), (31, 500)]
    J, in_shape, out_shape, depth_dim = 7, (80, 56), (20, 14), 8
    pairs = ((0, 1), (2, 5))
    Hs, Ws = max(s[0] for s in shapes), max(s[1] for s in shapes)
    batch = np.zeros((len(shapes), Hs, Ws, 3), np.uint8)
    items, augs, ann = [], [], []
    for n, (h, w) in enumerate(shapes):
        img = ar.synthetic_image(h, w, 500 + n)
        batch[n, :h, :w] = img
        bbox, joints, vis = ar.synthetic_annotation(h, w, J, 500 + n)
        if n == 3:
            bbox = np.array([-20., 30., 200., 300.], np.float32)          # sticks out of the image
        np.random.seed(900 + n); random.seed(900 + n)
        aug = data.NO_AUG if n == 2 else data.get_aug_config()
        augs.append(aug); ann.append((bbox, joints, vis))
        items.append(ar.get_item(img, bbox, joints, vis, pairs, aug, in_shape, out_shape, depth_dim, 2000.0, data.PIXEL_MEAN, data.PIXEL_STD))
    out, coord, vis = ihpr_b200.augment_batch(torch.from_numpy(batch).to(dev), shapes, [a[0] for a in ann], np.stack([a[1] for a in ann]),
                                              np.stack([a[2] for a in ann]), augs, flip_pairs=pairs, input_shape=in_shape, output_shape=out_shape,
                                              depth_dim=depth_dim, channels_last=channels_last)
    assert out.shape == (len(shapes), 3) + in_shape
    if channels_last:
        assert out.is_contiguous(memory_format=torch.channels_last)
    out, coord, vis = out.cpu().numpy(), coord.cpu().numpy(), vis.cpu().numpy()
    for n, (o_img, o_joint, o_vis, _) in enumerate(items):
        assert np.array_equal(out[n], o_img), (n, np.abs(out[n] - o_img).max(), (out[n] != o_img).mean())
        assert joints_close(coord[n], o_joint) and np.array_equal(vis[n], o_vis), n


def test_augment_feeds_the_loss(dev):
    """The staged targets go straight into the criterion: augment_batch -> JointLocationLoss runs and masks invisible joints."""
    import ihpr_b200
    g = load_aug_golden("aug_rect_small")
    n = len(g["seeds"])
    imgs = np.stack([ar.synthetic_image(g["h"], g["w"], int(s)) for s in g["seeds"]])
    ann = [ar.synthetic_annotation(g["h"], g["w"], g["J"], int(s)) for s in g["seeds"]]
    _, coord, vis = ihpr_b200.augment_batch(torch.from_numpy(imgs).to(dev), [[g["h"], g["w"]]] * n, [a[0] for a in ann], np.stack([a[1] for a in ann]),
                                            np.stack([a[2] for a in ann]), g["augs"], flip_pairs=g["pairs"], input_shape=tuple(g["input_shape"]),
                                            output_shape=tuple(g["output_shape"]), depth_dim=g["depth_dim"])
    oh, ow = (int(v) for v in g["output_shape"])
    heat = torch.randn(n, g["J"] * g["depth_dim"], oh, ow, device=dev, requires_grad=True)
    loss = ihpr_b200.JointLocationLoss()(heat, coord, vis, torch.ones(n, 1, device=dev))
    loss.backward()
    gv = heat.grad.view(n, g["J"], -1).abs().amax(-1)
    assert torch.isfinite(loss) and bool(((gv > 0) == (vis.view(n, g["J"]) > 0)).all())


def test_augment_argument_errors(dev):
    import ihpr_b200
    from ihpr_b200 import data
    img = torch.zeros(1, 8, 8, 3, dtype=torch.uint8, device=dev)
    args = ([[0, 0, 8, 8]], np.zeros((1, 2, 3)), np.ones((1, 2)), [data.NO_AUG])
    with pytest.raises(ValueError, match="sizes"):
        ihpr_b200.augment_batch(img, [[9, 8]], *args)
    with pytest.raises(ValueError, match="uint8"):
        ihpr_b200.augment_batch(img.float(), [[8, 8]], *args)
    with pytest.raises(ihpr_b200.IhprError, match="pixel_std"):
        ihpr_b200.augment_batch(img, [[8, 8]], *args, pixel_std=(0.2, 0.0, 0.2))
    out, coord, vis = ihpr_b200.augment_batch(img[:0], np.zeros((0, 2)), np.zeros((0, 4)), np.zeros((0, 2, 3)), np.zeros((0, 2)), [])
    assert out.shape == (0, 3, 256, 256) and coord.shape == (0, 2, 3) and vis.shape == (0, 2, 1)

"""GPU (B200): K6, the test-time post-processing of the soft-argmax result (ihpr_coords_to_camera), against the fixtures the
reference's own warp_coord_to_original / pixel2cam produced (tests/golden/post_*.npz) and against oracle/coords_post_ref.py.

Tolerances (fp32 kernel vs the reference's mixed fp32/fp64 numpy):
  merged  bit-identical (same fp32 operations in the same order as main/test.py:73-76)
  pixel   |a-b| <= 2e-6 * max(|b|, 1)
  cam     |a-b| <= 2e-6 * max |depth|   (root alignment subtracts two ~metres-sized numbers: the error scale is the depth)
"""
import numpy as np
import pytest
import torch

from conftest import load_post_golden, post_golden_names
from oracle import coords_post_ref, inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def to_dev(dev, *arrays):
    return [None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in arrays]


def check_against(out, merged, pix, cam):
    assert np.array_equal(out["merged"].cpu().numpy(), merged)
    got_p, got_c = out["pixel"].cpu().numpy().astype(np.float64), out["cam"].cpu().numpy().astype(np.float64)
    assert (np.abs(got_p - pix) <= 2e-6 * np.maximum(np.abs(pix), 1.0)).all(), np.abs(got_p - pix).max()
    scale = np.abs(pix[..., 2]).max()
    assert np.abs(got_c - cam).max() <= 2e-6 * scale, (np.abs(got_c - cam).max(), scale)


@pytest.mark.parametrize("name", post_golden_names())
def test_post_golden(name, dev):
    import ihpr_b200
    g = load_post_golden(name)
    coords, flipped, bbox, center, f, c = to_dev(dev, g["coords"], g["flipped"], g["bbox"], g["center_cam"], g["f"], g["c"])
    out = ihpr_b200.coords_to_camera(coords, bbox, center, f, c, root_idx=g["root"] if g["root"] >= 0 else None, flipped_coord_out=flipped,
                                     flip_pairs=g["pairs"], depth_dim=g["D"], output_shape=(g["H"], g["W"]), bbox_3d_depth=float(g["bbox_3d_depth"]))
    torch.cuda.synchronize()
    assert ihpr_b200.last_launch_count() == 1
    check_against(out, g["merged"], g["pixel"], g["cam"])
    if g["root"] >= 0:
        assert out["cam"][:, g["root"]].abs().max().item() == 0.0
    if flipped is not None:
        m = ihpr_b200.flip_merge(coords, flipped, g["W"], g["pairs"])
        assert np.array_equal(m.cpu().numpy(), g["merged"])


@pytest.mark.parametrize("case", [(1, 1, 64, 64, 64, True), (37, 18, 64, 64, 64, True), (5, 17, 32, 48, 80, False), (1024, 18, 64, 64, 64, True)])
def test_post_seeded_vs_oracle(case, dev):
    import ihpr_b200
    B, J, D, H, W, flip = case
    coords, flipped, bbox, center, f, c = coords_post_ref.make_inputs(B, J, D, H, W, 100 + B, flip)
    pairs = tuple((a, a + 1) for a in range(0, J - 1, 3))
    root = J // 2
    merged, pix, cam = coords_post_ref.post_process(coords, flipped, pairs, bbox, center, f, c, root, D, (H, W), 2000.0)
    t = to_dev(dev, coords, flipped, bbox, center, f, c)
    out = ihpr_b200.coords_to_camera(t[0], t[2], t[3], t[4], t[5], root_idx=root, flipped_coord_out=t[1], flip_pairs=pairs, depth_dim=D,
                                     output_shape=(H, W))
    check_against(out, merged, pix, cam)
    # subset of outputs, shared camera for the whole batch
    only = ihpr_b200.coords_to_camera(t[0], t[2], t[3], f=t[4][0], c=t[5][0], flipped_coord_out=t[1], flip_pairs=pairs, depth_dim=D,
                                      output_shape=(H, W), outputs=("cam",))
    assert set(only) == {"cam"}
    _, _, cam1 = coords_post_ref.post_process(coords, flipped, pairs, bbox, center, np.repeat(f[:1], B, 0), np.repeat(c[:1], B, 0), -1, D, (H, W), 2000.0)
    assert np.abs(only["cam"].cpu().numpy() - cam1).max() <= 2e-6 * np.abs(pix[..., 2]).max()


def test_flip_test_end_to_end(dev):
    """test.py:62-76 with both passes through K1: soft_argmax(heat), soft_argmax(heat of the mirrored image), merge.  A heat-map
    mirrored along W with left/right joint volumes swapped must merge back to the un-flipped coordinates."""
    import ihpr_b200
    B, J, D, H, W = 2, 6, 16, 32, 32
    pairs = ((0, 3), (1, 4))
    heat = torch.from_numpy(inputs.make_heat("blobs", B, J, D, H, W, 77)).to(dev)
    perm = ihpr_b200.flip_perm(J, pairs)
    mirrored = heat.view(B, J, D, H, W)[:, perm].flip(-1).reshape(B, J * D, H, W).contiguous()
    c0 = ihpr_b200.soft_argmax(heat, J)
    c1 = ihpr_b200.soft_argmax(mirrored, J)
    merged = ihpr_b200.flip_merge(c0, c1, W, pairs)
    assert (merged - c0).abs().max().item() <= 1e-4
    want = coords_post_ref.flip_merge(c0.cpu(), c1.cpu(), W, pairs)
    assert torch.equal(merged.cpu(), want)


def test_post_argument_errors(dev):
    import ihpr_b200
    c = torch.rand(2, 4, 3, device=dev)
    with pytest.raises(ValueError, match="required"):
        ihpr_b200.coords_to_camera(c)
    with pytest.raises(ValueError, match="root_idx"):
        ihpr_b200.coords_to_camera(c, torch.rand(2, 4, device=dev), torch.rand(2, 3, device=dev), torch.rand(2, 2, device=dev), torch.rand(2, 2, device=dev),
                                   root_idx=4)
    with pytest.raises(ValueError, match="shape"):
        ihpr_b200.coords_to_camera(c, torch.rand(3, 4, device=dev), torch.rand(2, 3, device=dev), outputs=("pixel",))
    with pytest.raises(ihpr_b200.IhprError):
        ihpr_b200.coords_to_camera(c, flipped_coord_out=torch.rand(2, 4, 3), outputs=("merged",))
    empty = ihpr_b200.coords_to_camera(torch.empty(0, 4, 3, device=dev), outputs=("merged",))
    assert empty["merged"].shape == (0, 4, 3)
    # raw C-ABI: aliasing and a null output set are refused
    from ihpr_b200._lib import lib
    rc = lib().ihpr_coords_to_camera(c.data_ptr(), None, None, 2, 4, 64, 64, 64, None, None, None, None, 2000.0, -1, c.data_ptr(), None, None, None)
    assert rc < 0 and b"alias" in lib().ihpr_last_error()
    rc = lib().ihpr_coords_to_camera(c.data_ptr(), None, None, 2, 4, 64, 64, 64, None, None, None, None, 2000.0, -1, None, None, None, None)
    assert rc < 0

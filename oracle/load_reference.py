"""oracle/load_reference.py -- TEST INFRASTRUCTURE (checker / CPU baseline only; never imported by the product path).

Imports the UNMODIFIED reference functions ``soft_argmax`` / ``JointLocationLoss``
(/root/reference/common/nets/loss.py:13-52) so that goldens can be generated from them and so that
``bench.py --impl reference`` / the ``cpu_baseline`` leg time the reference itself.

Two places the reference can be loaded from:

* ``Reference()``            -- a scratch copy of the whole /root/reference tree (build container only; used by
                                oracle/make_golden.py).  Importing main/config.py creates output directories next to the
                                sources (config.py:63-72) and /root/reference is read-only, hence the copy.
* ``Reference(installed())`` -- ``baseline/_ref/``: the five files the path needs (main/config.py, common/nets/loss.py,
                                common/utils/{__init__,dir_utils,pose_utils}.py) copied there, byte for byte, by
                                ``install()`` (run from ``__graft_entry__.build()`` whenever /root/reference is present).
                                ``baseline/_ref`` is git-ignored (no reference source enters the history) but not
                                gpurun-ignored, so it travels to the GPU box, where /root/reference does not exist.

No reference source line is edited.  What the loader does instead:

* loss.py:24-26 hard-codes ``torch.cuda.FloatTensor`` and ``torch.cuda.comm.broadcast``; for the CPU path the two attributes
  are aliased to their CPU equivalents while the reference runs (``cpu_shim()`` context manager; restored afterwards).
* loss.py:16,18 reads the volume shape from the global ``cfg``; ``set_shape`` sets it.
"""
import contextlib
import os
import shutil
import sys
import tempfile

REFERENCE_ROOT = "/root/reference"
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INSTALL_DIR = os.path.join(_ROOT, "baseline", "_ref")
PATH_FILES = ("main/config.py", "common/nets/loss.py", "common/utils/__init__.py", "common/utils/dir_utils.py",
              "common/utils/pose_utils.py")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "common", "nets"))


def installed():
    """baseline/_ref if install() has put the reference's files for this path there, else None."""
    ok = all(os.path.exists(os.path.join(INSTALL_DIR, f)) for f in PATH_FILES)
    return INSTALL_DIR if ok else None


def install():
    """Copy the reference's own files for this path (unmodified) into baseline/_ref.  Build container only."""
    if not available():
        return installed()
    for f in PATH_FILES:
        dst = os.path.join(INSTALL_DIR, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REFERENCE_ROOT, f), dst)
    return INSTALL_DIR


@contextlib.contextmanager
def cpu_shim():
    """The two attributes loss.py:24-26 needs to run on CPU tensors, set for the duration of the block."""
    import torch
    import torch.cuda.comm
    old = (torch.cuda.FloatTensor, torch.cuda.comm.broadcast)
    torch.cuda.FloatTensor = torch.FloatTensor
    torch.cuda.comm.broadcast = lambda t, devices=None, out=None: (t,)
    try:
        yield
    finally:
        torch.cuda.FloatTensor, torch.cuda.comm.broadcast = old


class Reference:
    def __init__(self, root=None):
        import torch            # noqa: F401
        import torch.cuda.comm  # noqa: F401  (loss.py uses torch.cuda.comm without importing it)
        self.tmp = None
        if root is None:
            if not available():
                raise RuntimeError("reference sources not present (expected only in the build container)")
            self.tmp = tempfile.mkdtemp(prefix="ihpr_ref_")
            root = os.path.join(self.tmp, "ref")
            shutil.copytree(REFERENCE_ROOT, root)
        self.root = root
        # the reference resolves `config`, `nets`, `utils` by bare module name (config.py:63): make sure they come from `root`
        for name in ("config", "nets", "nets.loss", "utils", "utils.pose_utils", "utils.dir_utils"):
            mod = sys.modules.get(name)
            if mod is not None and not str(getattr(mod, "__file__", "")).startswith(root):
                del sys.modules[name]
        sys.path.insert(0, os.path.join(root, "main"))
        cwd = os.getcwd()
        os.chdir(os.path.join(root, "main"))
        try:
            from config import cfg            # noqa: side effects (mkdir output3/*) happen inside `root`
            from nets import loss as ref_loss
        finally:
            os.chdir(cwd)
        assert os.path.abspath(ref_loss.__file__).startswith(os.path.abspath(root)), ref_loss.__file__
        self.cfg = cfg
        self.loss_file = ref_loss.__file__
        self.soft_argmax = ref_loss.soft_argmax
        self.JointLocationLoss = ref_loss.JointLocationLoss

    def set_shape(self, D, H, W):
        self.cfg.depth_dim = D
        self.cfg.output_shape = (H, W)

    def fwd_bwd(self, heat, gt, vis, have_depth, want_coords=True):
        """JointLocationLoss forward + autograd backward (main/train.py:67-71) on CPU tensors."""
        B, C, H, W = heat.shape
        self.set_shape(C // gt.shape[1], H, W)
        with cpu_shim():
            h = heat.detach().clone().requires_grad_(True)
            coords = self.soft_argmax(h, gt.shape[1]).detach() if want_coords else None
            loss = self.JointLocationLoss()(h, gt, vis, have_depth)
            loss.backward()
        return loss.detach(), coords, h.grad

    def step(self, heat, gt, vis, have_depth):
        """What one training iteration executes of this path, nothing else: criterion forward + backward."""
        self.set_shape(heat.shape[1] // gt.shape[1], heat.shape[2], heat.shape[3])
        with cpu_shim():
            h = heat.detach().requires_grad_(True)
            loss = self.JointLocationLoss()(h, gt, vis, have_depth)
            loss.backward()
        return loss

"""oracle/load_reference.py -- TEST INFRASTRUCTURE; build-container only.

Imports the UNMODIFIED reference functions ``soft_argmax`` / ``JointLocationLoss``
(/root/reference/common/nets/loss.py:13-52) so goldens can be generated from them.

* The reference is copied to a scratch directory first: importing main/config.py creates
  output directories next to the sources (config.py:63-72) and /root/reference is read-only.
* loss.py:24-26 hard-codes ``torch.cuda.FloatTensor`` and ``torch.cuda.comm.broadcast``; on a
  CPU-only box the two attributes are aliased to their CPU equivalents.  No reference source
  line is edited.
* loss.py:16,18 reads the volume shape from the global ``cfg``; ``set_shape`` sets it.

/root/reference does not exist on the GPU box: nothing run there may import this module.
"""
import os
import shutil
import sys
import tempfile

REFERENCE_ROOT = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "common", "nets"))


class Reference:
    def __init__(self):
        import torch
        import torch.cuda.comm
        if not available():
            raise RuntimeError("reference sources not present (expected only in the build container)")
        self.tmp = tempfile.mkdtemp(prefix="ihpr_ref_")
        dst = os.path.join(self.tmp, "ref")
        shutil.copytree(REFERENCE_ROOT, dst)
        if not torch.cuda.is_available():
            torch.cuda.FloatTensor = torch.FloatTensor
            torch.cuda.comm.broadcast = lambda t, devices=None, out=None: (t,)
        sys.path.insert(0, os.path.join(dst, "main"))
        cwd = os.getcwd()
        os.chdir(os.path.join(dst, "main"))
        try:
            from config import cfg            # noqa: side effects happen inside the scratch copy
            from nets import loss as ref_loss
        finally:
            os.chdir(cwd)
        self.cfg = cfg
        self.soft_argmax = ref_loss.soft_argmax
        self.JointLocationLoss = ref_loss.JointLocationLoss

    def set_shape(self, D, H, W):
        self.cfg.depth_dim = D
        self.cfg.output_shape = (H, W)

    def fwd_bwd(self, heat, gt, vis, have_depth):
        import torch
        B, C, H, W = heat.shape
        self.set_shape(C // gt.shape[1], H, W)
        h = heat.detach().clone().requires_grad_(True)
        coords = self.soft_argmax(h, gt.shape[1]).detach()
        loss = self.JointLocationLoss()(h, gt, vis, have_depth)
        loss.backward()
        return loss.detach(), coords, h.grad

"""Installs the B200 path behind the reference's own import names.

The reference resolves the hot path by module name: ``from nets.loss import soft_argmax``
(/root/reference/main/train.py:7, main/test.py:11, main/main.py:13) and ``from nets import loss``
style lookups in common/base.py:20,71,139.  ``install_dropin()`` registers this package's
``nets.loss`` under that name BEFORE those imports run, so train.py / test.py work unchanged.
"""
import sys
import types


def install_dropin(force=True):
    from .nets import loss as b200_loss
    try:
        import nets                 # the reference's common/nets when common/ is on sys.path (config.py:63)
    except ImportError:
        nets = types.ModuleType("nets")
        nets.__path__ = []
        sys.modules["nets"] = nets
    if force or "nets.loss" not in sys.modules:
        sys.modules["nets.loss"] = b200_loss
        nets.loss = b200_loss
    return b200_loss

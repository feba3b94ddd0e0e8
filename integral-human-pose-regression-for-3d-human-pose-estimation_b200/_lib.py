"""ctypes binding of lib/libihpr_b200.so (the C-ABI in include/ihpr_b200.h).

ctypes releases the GIL around every call, which is what the reference's threaded criterion
(/root/reference/common/nets/balanced_parallel.py:149-173) needs.  Missing library => loud error;
nothing here ever substitutes a PyTorch / CPU implementation.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "lib", "libihpr_b200.so")
_lock = threading.Lock()
_lib = None

IHPR_F32, IHPR_BF16 = 0, 1

c_void_p, c_int, c_size_t, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_float

# name -> (restype, argtypes): must list every symbol include/ihpr_b200.h declares
SIGNATURES = {
    "ihpr_version": (c_int, []),
    "ihpr_last_error": (ctypes.c_char_p, []),
    "ihpr_workspace_bytes": (c_size_t, [c_int] * 5),
    "ihpr_softargmax3d_fwd": (c_int, [c_void_p, c_int] + [c_int] * 5 + [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ihpr_softargmax3d_bwd": (c_int, [c_void_p, c_int] + [c_int] * 5 + [c_void_p] * 5),
    "ihpr_integral_l1_fwd": (c_int, [c_void_p, c_int] + [c_int] * 5 + [c_void_p] * 6 + [c_void_p, c_size_t, c_void_p]),
    "ihpr_integral_l1_bwd": (c_int, [c_void_p, c_int] + [c_int] * 5 + [c_void_p] * 8),
    "ihpr_integral_l1_fwd_bwd": (c_int, [c_void_p, c_int] + [c_int] * 5 + [c_void_p] * 7 + [c_void_p, c_size_t, c_void_p]),
    "ihpr_scale_grad": (c_int, [c_void_p, c_int, c_size_t, c_void_p, c_void_p]),
    "ihpr_augment_patches": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int] + [c_void_p] * 5 + [c_int, c_int, c_void_p, c_int, c_void_p]),
    "ihpr_augment_joints": (c_int, [c_void_p] * 7 + [c_int] * 7 + [ctypes.c_double] + [c_void_p] * 3),
    "ihpr_integral_l1_from_coords": (c_int, [c_void_p] * 4 + [c_int, c_int, c_void_p, c_void_p]),
    "ihpr_coords_to_camera": (c_int, [c_void_p] * 3 + [c_int] * 5 + [c_void_p] * 4 + [c_float, c_int] + [c_void_p] * 4),
    "ihpr_head_softargmax_fwd": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 6 + [c_void_p, c_void_p, c_void_p]),
    "ihpr_head_integral_l1_bwd": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 6 + [c_void_p] * 9),
    "ihpr_head_bwd_workspace_bytes": (c_size_t, [c_int] * 6),
    "ihpr_head_integral_l1_bwd_params": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 6 + [c_void_p] * 10 + [c_size_t, c_void_p]),
    "ihpr_deconv_bn_relu_workspace_bytes": (c_size_t, [c_int] * 2),
    "ihpr_deconv_bn_relu_prepare": (c_int, [c_void_p] * 5 + [c_float, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "ihpr_deconv_bn_relu": (c_int, [c_void_p, c_void_p] + [c_int] * 5 + [c_void_p, c_void_p]),
    "ihpr_deconv_train_workspace_bytes": (c_size_t, [c_int] * 2),
    "ihpr_deconv_bn_relu_train_fwd": (c_int, [c_void_p] * 6 + [c_float, c_float] + [c_int] * 5 + [c_void_p] * 4 + [c_size_t, c_void_p]),
    "ihpr_deconv_bn_relu_train_bwd": (c_int, [c_void_p] * 4 + [c_int] * 5 + [c_void_p] * 5 + [c_size_t, c_void_p]),
    "ihpr_deconv_wgrad_workspace_bytes": (c_size_t, [c_int] * 2),
    "ihpr_deconv_wgrad": (c_int, [c_void_p, c_void_p] + [c_int] * 5 + [c_void_p, c_void_p, c_size_t, c_void_p]),
    "ihpr_integral_l1_fwd_bwd_host": (c_int, [c_void_p, c_int] + [c_int] * 5 + [c_void_p] * 3 + [c_float] + [c_void_p] * 3 + [c_int, c_int]),
    "ihpr_host_release": (c_int, [c_int]),
    "ihpr_set_variant": (c_int, [c_int]),
    "ihpr_get_variant": (c_int, []),
    "ihpr_last_launch_count": (c_int, []),
    "ihpr_last_path_choice": (c_int, []),
}


class IhprError(RuntimeError):
    pass


def library_path():
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(_LIB_PATH):
                    raise IhprError(
                        "ihpr_b200: %s is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(needs nvcc). There is no CPU / PyTorch fallback for this path." % _LIB_PATH)
                handle = ctypes.CDLL(_LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.restype, fn.argtypes = res, args
                _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().ihpr_last_error()
        raise IhprError("ihpr_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))


def version():
    return lib().ihpr_version()

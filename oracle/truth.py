"""oracle/truth.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes loader for oracle/truth64.c (fp64 restatement of loss.py:13-52 + analytic backward).
Row ranges are fanned out over Python threads (ctypes releases the GIL).
"""
import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libihpr_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "truth64.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-fno-fast-math", "-shared", "-o", _SO, src, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.ihpr_oracle_loss_f64.restype = ctypes.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _ranges(R, nthreads):
    nthreads = max(1, min(nthreads, R))
    edges = [R * k // nthreads for k in range(nthreads + 1)]
    return [(edges[k], edges[k + 1]) for k in range(nthreads) if edges[k + 1] > edges[k]]


def _fan(fn, R, threads):
    rs = _ranges(R, threads or os.cpu_count() or 1)
    if len(rs) == 1:
        fn(*rs[0])
        return
    with ThreadPoolExecutor(len(rs)) as ex:
        list(ex.map(lambda ab: fn(*ab), rs))


def soft_argmax_f64(heat, joint_num, threads=None):
    """heat: float32 ndarray (B, J*D, H, W).  Returns coords (B,J,3), m (B,J), l (B,J) in float64."""
    heat = np.ascontiguousarray(heat, dtype=np.float32)
    B, C, H, W = heat.shape
    D = C // joint_num
    R, N = B * joint_num, D * H * W
    coords = np.empty((R, 3)); m = np.empty(R); l = np.empty(R)
    L = lib()
    flat = heat.reshape(R, N)

    def run(a, b):
        L.ihpr_oracle_fwd_f64(_p(flat[a:b]), ctypes.c_long(b - a), D, H, W, _p(coords[a:b]), _p(m[a:b]), _p(l[a:b]))
    _fan(run, R, threads)
    return coords.reshape(B, joint_num, 3), m.reshape(B, joint_num), l.reshape(B, joint_num)


def soft_argmax_bwd_f64(heat, joint_num, coords, m, l, gcoords, threads=None):
    heat = np.ascontiguousarray(heat, dtype=np.float32)
    B, C, H, W = heat.shape
    D = C // joint_num
    R, N = B * joint_num, D * H * W
    flat = heat.reshape(R, N)
    coords = np.ascontiguousarray(coords, dtype=np.float64).reshape(R, 3)
    gcoords = np.ascontiguousarray(gcoords, dtype=np.float64).reshape(R, 3)
    m = np.ascontiguousarray(m, dtype=np.float64).reshape(R)
    l = np.ascontiguousarray(l, dtype=np.float64).reshape(R)
    g = np.empty((R, N))
    L = lib()

    def run(a, b):
        L.ihpr_oracle_bwd_f64(_p(flat[a:b]), ctypes.c_long(b - a), D, H, W, _p(coords[a:b]), _p(m[a:b]),
                              _p(l[a:b]), _p(gcoords[a:b]), _p(g[a:b]))
    _fan(run, R, threads)
    return g.reshape(heat.shape)


def loss_f64(coords, gt, vis, have_depth, grad_out=1.0, want_grad=True):
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    B, J, _ = coords.shape
    gt = np.ascontiguousarray(gt, dtype=np.float32).reshape(B, J, 3)
    vis = np.ascontiguousarray(vis, dtype=np.float32).reshape(B, J)
    hd = np.ascontiguousarray(have_depth, dtype=np.float32).reshape(B)
    g = np.empty((B, J, 3)) if want_grad else None
    val = lib().ihpr_oracle_loss_f64(_p(coords), _p(gt), _p(vis), _p(hd), ctypes.c_long(B), ctypes.c_long(J),
                                     ctypes.c_double(grad_out), _p(g))
    return val, g


def fwd_bwd_f64(heat, gt, vis, have_depth, grad_out=1.0, threads=None):
    """Full path in fp64: returns (loss, coords (B,J,3), grad_heat like heat) as float64."""
    J = np.asarray(gt).shape[1]
    coords, m, l = soft_argmax_f64(heat, J, threads)
    loss, gc = loss_f64(coords, gt, vis, have_depth, grad_out)
    gh = soft_argmax_bwd_f64(heat, J, coords, m, l, gc, threads)
    return loss, coords, gh


def fwd_bwd_f32_port(heat, gt, vis, have_depth, want_grad=True, threads=None):
    """fp32 scalar C port of the whole path (alternative CPU baseline)."""
    heat = np.ascontiguousarray(heat, dtype=np.float32)
    B, C, H, W = heat.shape
    J = np.asarray(gt).shape[1]
    D = C // J
    R = B * J
    gt = np.ascontiguousarray(gt, dtype=np.float32).reshape(R, 3)
    vis = np.ascontiguousarray(vis, dtype=np.float32).reshape(R)
    hd = np.ascontiguousarray(have_depth, dtype=np.float32).reshape(B)
    coords = np.empty((R, 3), np.float32); row_loss = np.empty(R, np.float32)
    g = np.empty_like(heat) if want_grad else None
    L = lib()

    def run(a, b):
        L.ihpr_oracle_fwd_bwd_f32(_p(heat), ctypes.c_long(R), ctypes.c_long(J), ctypes.c_long(a), ctypes.c_long(b),
                                  D, H, W, _p(gt), _p(vis), _p(hd), _p(coords), _p(row_loss), _p(g))
    _fan(run, R, threads)
    return float(row_loss.astype(np.float64).mean()), coords.reshape(B, J, 3), g

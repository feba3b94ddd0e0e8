#!/usr/bin/env python
"""tools/one_cell.py -- a handful of fwd+bwd launches of one (variant, dtype, shape) cell: the short command ncu wraps."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.sweep import time_cell

ap = argparse.ArgumentParser()
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--dtype", default="f32")
ap.add_argument("--B", type=int, default=32)
ap.add_argument("--J", type=int, default=18)
ap.add_argument("--D", type=int, default=64)
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
print(time_cell(a.B, a.J, a.D, 64, 64, torch.float32 if a.dtype == "f32" else torch.bfloat16, a.variant, a.iters, 1))

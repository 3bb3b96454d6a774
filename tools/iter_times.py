"""Per-launch device time (CUDA events recorded by the library) of the iteration kernels for a few batch sizes."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import vision_transform_codes_b200 as pkg
from vision_transform_codes_b200 import _lib
from vision_transform_codes_b200.analysis_transforms.fully_connected import ista_fista
from oracle import vtc_oracle as oracle

lib = _lib.load()
pkg.config.check_finite = False
S, D, T = 1024, 256, 40
batches = [int(v) for v in sys.argv[1].split(',')] if len(sys.argv) > 1 else [512, 4096, 16384, 65536]
phi = oracle.synthetic_dictionary(S, D).cuda()
for B in batches:
  x = oracle.synthetic_patches(B, D).cuda()
  for prec in ('bf16x3', 'bf16'):
    pkg.config.precision = prec
    for fused in (0, 1):
      lib.vtc_set_fused_iteration(fused)
      ista_fista.run(x, phi, 0.1, T)
      lib.vtc_profile_enable(1)
      ista_fista.run(x, phi, 0.1, T)
      f = [ctypes.c_float() for _ in range(4)]
      n = [ctypes.c_int(), ctypes.c_int()]
      _lib.check(lib.vtc_profile_last(ctypes.byref(f[0]), ctypes.byref(f[1]), ctypes.byref(n[0]), ctypes.byref(n[1]),
                                      ctypes.byref(f[2]), ctypes.byref(f[3])))
      lib.vtc_profile_enable(0)
      print('B=%6d %-6s fused=%d  iter %.4f ms  (launch %.4f, first %.4f)  setup %.3f ms' %
            (B, prec, fused, f[1].value / n[1].value, f[2].value, f[3].value, f[0].value), flush=True)

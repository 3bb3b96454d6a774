"""Per-iteration device time of the persistent iteration kernel with pieces left out (VTC_B200_ABLATE, results wrong).

One process per setting (the switch is read once): tools/ablate.sh loops over the settings.
"""
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
S, D = 1024, 256
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 60
phi = oracle.synthetic_dictionary(S, D).cuda()
x = oracle.synthetic_patches(B, D).cuda()
out = []
for prec in ('bf16x3', 'bf16'):
  pkg.config.precision = prec
  ista_fista.run(x, phi, 0.1, T)
  best = 1e9
  for _ in range(3):
    lib.vtc_profile_enable(1)
    ista_fista.run(x, phi, 0.1, T)
    f = [ctypes.c_float() for _ in range(4)]
    n = [ctypes.c_int(), ctypes.c_int()]
    _lib.check(lib.vtc_profile_last(ctypes.byref(f[0]), ctypes.byref(f[1]), ctypes.byref(n[0]), ctypes.byref(n[1]),
                                    ctypes.byref(f[2]), ctypes.byref(f[3])))
    lib.vtc_profile_enable(0)
    best = min(best, f[1].value / n[1].value)
  out.append('%s %.4f' % (prec, best))
print('ablate=%-4s variant=%-2s  ms/iter: %s' % (os.environ.get('VTC_B200_ABLATE', '0'),
                                                os.environ.get('VTC_B200_ITER_VARIANT', '-'), '   '.join(out)), flush=True)

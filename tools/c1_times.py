"""BASELINE configs[0] (batch 250, 256 atoms, D=256, 300 FISTA iterations): wall time per call, both formulations."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import vision_transform_codes_b200 as pkg
from vision_transform_codes_b200 import _lib
from vision_transform_codes_b200.analysis_transforms.fully_connected import ista_fista
from oracle import vtc_oracle as oracle

lib = _lib.load()
pkg.config.check_finite = False
phi = oracle.synthetic_dictionary(256, 256).cuda()
x = oracle.synthetic_patches(250, 256).cuda()
for form, name in ((1, 'gram'), (2, 'synthesis (persistent)')):
  lib.vtc_set_formulation(form)
  for _ in range(3):
    ista_fista.run(x, phi, 0.1, 300)
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  n = 10
  for _ in range(n):
    a = ista_fista.run(x, phi, 0.1, 300)
  torch.cuda.synchronize()
  print('configs[0] %s: %.3f ms per call' % (name, (time.perf_counter() - t0) / n * 1e3))
want = oracle.ista_fista(x.cpu(), phi.cpu(), 0.1, 300)
print('rel-L2 vs oracle %.2e' % oracle.relative_l2(a.cpu(), want))
t0 = time.perf_counter()
oracle.ista_fista(x.cpu(), phi.cpu(), 0.1, 300)
print('CPU oracle: %.1f ms' % ((time.perf_counter() - t0) * 1e3))

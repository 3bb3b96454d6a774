"""Compare the one-launch (panel-resident) schedule of an iteration with the two-launch schedule on one shape.
usage: python tools/iter_debug.py B S D T [precision] [variant]   -- run one case per process (a trap kills the context)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import vision_transform_codes_b200 as pkg
from vision_transform_codes_b200 import _lib
from vision_transform_codes_b200.analysis_transforms.fully_connected import ista_fista
from oracle import vtc_oracle as oracle

B, S, D, T = (int(v) for v in sys.argv[1:5])
pkg.config.precision = sys.argv[5] if len(sys.argv) > 5 else 'bf16x3'
variant = sys.argv[6] if len(sys.argv) > 6 else 'fista'
lib = _lib.load()
_lib.check(lib.vtc_set_formulation(2))
phi = oracle.synthetic_dictionary(S, D).cuda()
x = oracle.synthetic_patches(B, D).cuda()
out = {}
for fused in (0, 1):
  lib.vtc_set_fused_iteration(fused)
  a = ista_fista.run(x, phi, 0.1, T, variant=variant)
  torch.cuda.synchronize()
  t0 = time.time()
  a = ista_fista.run(x, phi, 0.1, T, variant=variant)
  torch.cuda.synchronize()
  out[fused] = (a, time.time() - t0)
err = oracle.relative_l2(out[1][0].cpu(), out[0][0].cpu())
nz = (out[0][0] != 0).float().mean().item()
print('B=%d S=%d D=%d T=%d %s %s: fused vs two-launch rel-L2 %.3e (nonzeros %.3f)  two-launch %.2f ms  fused %.2f ms' %
      (B, S, D, T, pkg.config.precision, variant, err, nz, out[0][1] * 1e3, out[1][1] * 1e3), flush=True)
if B * S <= 600 * 1024:
  want = oracle.ista_fista(x.cpu(), phi.cpu(), 0.1, T, variant=variant)
  print('   vs oracle: fused %.3e  two-launch %.3e' % (oracle.relative_l2(out[1][0].cpu(), want),
                                                      oracle.relative_l2(out[0][0].cpu(), want)), flush=True)

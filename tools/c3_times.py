"""Per-iteration device time of the subspace path at the configs[3] shape (D = 1024, 4096 atoms, groups of 2)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import vision_transform_codes_b200 as pkg
from oracle import vtc_oracle as oracle
from vision_transform_codes_b200.analysis_transforms.fully_connected import subspace_ista_fista

pkg.config.check_finite = False
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 30
S, D = 4096, 1024
phi = oracle.synthetic_dictionary(S, D).cuda()
x = 0.3 * torch.randn(B, D, device='cuda', generator=torch.Generator(device='cuda').manual_seed(0))
groups = [list(map(int, g)) for g in np.array_split(np.arange(S), S // 2)]
out = []
for prec in ('bf16x3', 'bf16'):
  pkg.config.precision = prec
  subspace_ista_fista.run(x, phi, groups, 0.1, 3)
  best = 1e9
  for _ in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    subspace_ista_fista.run(x, phi, groups, 0.1, T)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / T)
  out.append('%s %.4f' % (prec, best))
print('configs[3] shape, B=%d: ms/iter (setup included): %s' % (B, '   '.join(out)), flush=True)

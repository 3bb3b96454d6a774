"""Short driver for ncu: a few fused FISTA iterations at the BASELINE configs[1] shape. Not a benchmark."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import vision_transform_codes_b200 as pkg  # noqa: E402
from oracle import vtc_oracle as oracle  # noqa: E402  (seeded input generators only)
from vision_transform_codes_b200.analysis_transforms.fully_connected import ista_fista  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--precision', default='bf16x3')
ap.add_argument('--iters', type=int, default=10)
ap.add_argument('--batch', type=int, default=65536)
ap.add_argument('--atoms', type=int, default=1024)
ap.add_argument('--pixels', type=int, default=256)
ap.add_argument('--group', type=int, default=0, help='subspace variant with in-order groups of this size (configs[3]: 2)')
args = ap.parse_args()
pkg.config.precision = args.precision
pkg.config.check_finite = False
dev = torch.device('cuda:0')
phi = oracle.synthetic_dictionary(args.atoms, args.pixels).to(dev)
x = oracle.synthetic_patches(args.batch, args.pixels).to(dev)
if args.group:
  import numpy as np
  from vision_transform_codes_b200.analysis_transforms.fully_connected import subspace_ista_fista
  groups = [list(map(int, g)) for g in np.array_split(np.arange(args.atoms), args.atoms // args.group)]
  codes = subspace_ista_fista.run(x, phi, groups, 0.1, args.iters)
else:
  codes = ista_fista.run(x, phi, 0.1, args.iters)
torch.cuda.synchronize()
print('ok', float(codes.abs().mean()))

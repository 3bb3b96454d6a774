"""Short driver for ncu: a few convolutional FISTA iterations at the BASELINE configs[4] shape (--time: also prints the
device time per iteration of three repeats, for same-box A/B runs with VTC_B200_LIB). Not a benchmark."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import vision_transform_codes_b200 as pkg  # noqa: E402
from oracle import vtc_oracle as oracle  # noqa: E402  (seeded input generators only)
from vision_transform_codes_b200.analysis_transforms.convolutional import ista_fista  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--precision', default='bf16x3')
ap.add_argument('--iters', type=int, default=10)
ap.add_argument('--images', type=int, default=128)
ap.add_argument('--time', action='store_true')
args = ap.parse_args()
pkg.config.precision = args.precision
pkg.config.check_finite = False
dev = torch.device('cuda:0')
nk, k, st, side = 64, (16, 16), (8, 8), 512
few, pad = oracle.synthetic_padded_images(8, 1, side, side, k, st, seed=100)
x = few.repeat((args.images + 7) // 8, 1, 1, 1)[:args.images].contiguous().to(dev)
phi = oracle.synthetic_conv_dictionary(nk, 1, k[0], k[1]).to(dev)
codes = ista_fista.run(x, phi, st, pad, 0.05, args.iters)
torch.cuda.synchronize()
if args.time:
  ms = []
  for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ista_fista.run(x, phi, st, pad, 0.05, args.iters)
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1) / args.iters)
  print('ms/iter: ' + ' '.join('%.4f' % m for m in ms))
print('ok', float(codes.abs().mean()))

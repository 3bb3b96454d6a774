"""Timing of the convolutional path at the BASELINE configs[4] shape. usage: python tools/conv_bench.py [images] [precision] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import vision_transform_codes_b200 as pkg
from oracle import vtc_oracle as oracle
from vision_transform_codes_b200.analysis_transforms.convolutional import ista_fista

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
pkg.config.precision = sys.argv[2] if len(sys.argv) > 2 else 'bf16x3'
T = int(sys.argv[3]) if len(sys.argv) > 3 else 100
pkg.config.check_finite = False
few, pad = oracle.synthetic_padded_images(8, 1, 512, 512, (16, 16), (8, 8))
x = few.repeat((n + 7) // 8, 1, 1, 1)[:n].contiguous().cuda()
phi = oracle.synthetic_conv_dictionary(64, 1, 16, 16).cuda()
for _ in range(2):
  ista_fista.run(x, phi, (8, 8), pad, 0.05, T)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
codes = ista_fista.run(x, phi, (8, 8), pad, 0.05, T)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
rows = n * 66 * 66
P = {'bf16': 1, 'bf16x3': 2, 'bf16x6': 3}[pkg.config.precision]
by = rows * 64 * (12 + 4 + 8 * P)
print('%d images %s: %.2f ms / %d iters = %.4f ms per iteration; %.0f GB/s algorithmic; nonzeros %.3f; finite %s' %
      (n, pkg.config.precision, ms, T, ms / T, by / (ms / T * 1e-3) / 1e9, (codes != 0).float().mean().item(),
       bool(torch.isfinite(codes).all())))
if n <= 8:
  want = oracle.conv_ista_fista(x.cpu(), phi.cpu(), (8, 8), pad, 0.05, T)
  print('   vs oracle rel-L2 %.3e' % oracle.relative_l2(codes.cpu(), want))

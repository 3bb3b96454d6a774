"""BASELINE configs[0] (batch 250, 256 atoms, D=256, 300 FISTA iterations): eager call vs replay of the captured CUDA graph."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import vision_transform_codes_b200 as pkg
from oracle import vtc_oracle as oracle
from vision_transform_codes_b200.analysis_transforms.fully_connected import ista_fista

pkg.config.check_finite = False
phi = oracle.synthetic_dictionary(256, 256).cuda()
x = oracle.synthetic_patches(250, 256).cuda()
side = torch.cuda.Stream()
with torch.cuda.stream(side):
  for _ in range(3):
    eager = ista_fista.run(x, phi, 0.1, 300)
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph, stream=side):
  out = ista_fista.run(x, phi, 0.1, 300)
graph.replay()
torch.cuda.synchronize()
print('replay equals eager:', bool(torch.equal(out, eager)))
for name, fn in (('eager', lambda: ista_fista.run(x, phi, 0.1, 300)), ('graph replay', graph.replay)):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  for _ in range(20):
    fn()
  torch.cuda.synchronize()
  print('configs[0] %s: %.3f ms per call' % (name, (time.perf_counter() - t0) / 20 * 1e3))

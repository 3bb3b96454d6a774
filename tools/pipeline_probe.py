"""Where the host pipeline's time goes: variants of the e2e loop at configs[1] (one GPU)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import vision_transform_codes_b200 as pkg
from oracle import vtc_oracle as oracle
from vision_transform_codes_b200.analysis_transforms.fully_connected import ista_fista
from vision_transform_codes_b200.host_pipeline import HostPipeline

B, S, D, T = 65536, 1024, 256, 300
dev = torch.device('cuda:0')
phi = oracle.synthetic_dictionary(S, D).to(dev)
x_host = oracle.synthetic_patches(B, D).pin_memory()
x = x_host.to(dev)
out = [torch.empty((B, S)).pin_memory() for _ in range(2)]


def device_loop(n):
  for _ in range(n):
    ista_fista.run(x, phi, 0.1, T)
  torch.cuda.synchronize()


def pipe_loop(n, depth, download=True, upload=True):
  pipe = HostPipeline(dev, depth=depth)
  if not download or not upload:
    orig = pipe.submit

  for i in range(n):
    pipe.submit(x_host, phi, 0.1, T, out=out[i % 2])
  pipe.synchronize()


for check in (False, True):
  pkg.config.check_finite = check
  device_loop(2)
  t0 = time.perf_counter(); device_loop(6); t_dev = (time.perf_counter() - t0) / 6 * 1e3
  for depth in (1, 2, 3):
    pipe_loop(2, depth)
    t0 = time.perf_counter(); pipe_loop(6, depth); t = (time.perf_counter() - t0) / 6 * 1e3
    print('check_finite=%s  device-resident %.2f ms/step   pipeline depth %d: %.2f ms/step' % (check, t_dev, depth, t), flush=True)
# copies alone
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(4):
  xd = x_host.to(dev, non_blocking=True)
torch.cuda.synchronize()
print('H2D 67 MB: %.2f ms' % ((time.perf_counter() - t0) / 4 * 1e3))
c = torch.empty((B, S), device=dev)
t0 = time.perf_counter()
for _ in range(4):
  out[0].copy_(c, non_blocking=True)
torch.cuda.synchronize()
print('D2H 268 MB: %.2f ms' % ((time.perf_counter() - t0) / 4 * 1e3))
# a copy running WHILE the kernel runs
s2 = torch.cuda.Stream()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(4):
  ista_fista.run(x, phi, 0.1, T)
  with torch.cuda.stream(s2):
    out[0].copy_(c, non_blocking=True)
torch.cuda.synchronize()
print('4 runs with a concurrent unrelated D2H each: %.2f ms/step' % ((time.perf_counter() - t0) / 4 * 1e3))

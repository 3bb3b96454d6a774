"""Large-shape checks on one B200 (properties that do not need the CPU oracle at full size) + timing:
   C4: subspace FISTA, 32x32 patches (D=1024), 4096 atoms, groups of 2, batch 131072
   C2: fixed point / determinism / shard-independence at batch 65536."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import vision_transform_codes_b200 as pkg  # noqa: E402
from oracle import vtc_oracle as oracle  # noqa: E402
from vision_transform_codes_b200.analysis_transforms.fully_connected import ista_fista, subspace_ista_fista  # noqa: E402

dev = torch.device('cuda:0')
pkg.config.check_finite = False


def timed(fn, n=2):
  fn()
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  for _ in range(n):
    out = fn()
  torch.cuda.synchronize()
  return out, (time.perf_counter() - t0) / n


which = sys.argv[1] if len(sys.argv) > 1 else 'c4'
if which == 'c4':
  B, S, D, T = 131072, 4096, 1024, 300
  phi = oracle.synthetic_dictionary(S, D).to(dev)
  x = oracle.synthetic_patches(B, D).to(dev)
  groups = [list(map(int, g)) for g in np.array_split(np.arange(S), S // 2)]
  for prec in ('bf16x3', 'bf16'):
    pkg.config.precision = prec
    codes, sec = timed(lambda: subspace_ista_fista.run(x, phi, groups, 0.1, T), n=1)
    print('C4 %s: %.3f s per run, %.0f patches/s, nnz frac %.4f' % (prec, sec, B / sec, float((codes != 0).float().mean())))
  # parity on a sub-batch against the oracle (the full batch would take the CPU half an hour)
  pkg.config.precision = 'bf16x3'
  sub = 256
  want = oracle.subspace_ista_fista(x[:sub].cpu(), phi.cpu(), groups, 0.1, T)
  got = subspace_ista_fista.run(x[:sub], phi, groups, 0.1, T).cpu()
  print('C4 sub-batch parity: codes rel-L2 %.3e' % oracle.relative_l2(got, want))
  full = subspace_ista_fista.run(x, phi, groups, 0.1, T)
  print('C4 batch independence: max |full[:sub] - sub| = %.3e' % float((full[:sub].cpu() - got).abs().max()))
else:
  B, S, D, T = 65536, 1024, 256, 300
  phi = oracle.synthetic_dictionary(S, D).to(dev)
  x = oracle.synthetic_patches(B, D, kind='whitened').to(dev)
  a = ista_fista.run(x, phi, 0.1, T)
  b = ista_fista.run(x, phi, 0.1, T)
  print('C2 deterministic:', bool(torch.equal(a, b)))
  half = ista_fista.run(x[B // 2:], phi, 0.1, T)
  print('C2 shard independence: max |full[B/2:] - half| = %.3e' % float((a[B // 2:] - half).abs().max()))
  # a converged code is a fixed point: one more ISTA step from it changes (almost) nothing
  more = ista_fista.run(x, phi, 0.1, 1, variant='ista', initial_codes=a)
  print('C2 fixed point: rel change after one more ISTA step %.3e' % oracle.relative_l2(more.cpu(), a.cpu()))
  # lasso objective must not be worse than the oracle's on a sub-batch
  sub = 1024
  want = oracle.ista_fista(x[:sub].cpu(), phi.cpu(), 0.1, T)
  def obj(c, xx):
    return float(0.5 * ((c @ phi.cpu() - xx) ** 2).sum(1).mean() + 0.1 * c.abs().sum(1).mean())
  print('C2 objective: cuda %.6f oracle %.6f; codes rel-L2 %.3e' % (obj(a[:sub].cpu(), x[:sub].cpu()), obj(want, x[:sub].cpu()), oracle.relative_l2(a[:sub].cpu(), want)))

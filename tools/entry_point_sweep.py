"""
One small call of every CUDA entry point, each checked against the CPU oracle: python tools/entry_point_sweep.py
Sizes are off the tile sizes on purpose (ragged panels, K tails, pitched rows) and small, so that the sweep also suits
an instrumented run (compute-sanitizer is closed on the GPU pool this was developed on; the sweep itself, the parity
tests and the fuzz tools are the out-of-bounds evidence). Prints one line per call.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import vtc_oracle as oracle  # noqa: E402
import vision_transform_codes_b200 as pkg  # noqa: E402
from vision_transform_codes_b200 import _lib  # noqa: E402
from vision_transform_codes_b200.analysis_transforms.convolutional import ista_fista as conv_inf  # noqa: E402
from vision_transform_codes_b200.analysis_transforms.fully_connected import ista_fista, subspace_ista_fista  # noqa: E402
from vision_transform_codes_b200.dict_update_rules.convolutional import sc_cheap_quadratic_descent as conv_cheap  # noqa: E402
from vision_transform_codes_b200.dict_update_rules.fully_connected import (  # noqa: E402
    sc_cheap_quadratic_descent, subspace_sc_cheap_quadratic_descent)
from vision_transform_codes_b200.lean import metrics  # noqa: E402
from vision_transform_codes_b200.utils import dataset_generation  # noqa: E402


def report(name, got, want, tol):
  err = oracle.relative_l2(got.cpu(), want)
  print('%-34s rel-L2 %.2e %s' % (name, err, 'ok' if err <= tol else 'FAIL'), flush=True)
  if not err <= tol:
    sys.exit(1)


def main():
  lib = _lib.load()
  T = 6
  B, S, D = 300, 328, 72
  phi, x = oracle.synthetic_dictionary(S, D), oracle.synthetic_patches(B, D)
  pd, xd = phi.cuda(), x.cuda()
  want = oracle.ista_fista(x, phi, 0.1, T)
  for precision in ('bf16x3', 'bf16'):
    pkg.config.precision = precision
    tol = 1e-4 if precision == 'bf16x3' else 5e-2
    for name, form, fused in (('gram', 1, 1), ('persistent panel-resident', 2, 1), ('two-launch synthesis', 2, 0)):
      _lib.check(lib.vtc_set_formulation(form))
      _lib.check(lib.vtc_set_fused_iteration(fused))
      report('fista %s %s' % (precision, name), ista_fista.run(xd, pd, 0.1, T), want, tol)
  pkg.config.precision = 'bf16x3'
  lib.vtc_set_formulation(0)
  lib.vtc_set_fused_iteration(1)
  got, iters = ista_fista.infer(xd, pd, 0.1, 40, 'ista', None, 1e-2, False, False, 1)
  report('ista early stopping (%d iters)' % iters, got,
         oracle.ista_fista(x, phi, 0.1, 40, variant='ista', early_stopping_epsilon=1e-2), 1e-4)
  groups = [list(range(i, min(i + 3, S))) for i in range(0, S, 3)]
  report('subspace fista (groups of 3)', subspace_ista_fista.run(xd, pd, groups, 0.1, T),
         oracle.subspace_ista_fista(x, phi, groups, 0.1, T), 1e-4)
  h = oracle.hessian_running_mean(torch.zeros(S), want)
  d = pd.clone()
  sc_cheap_quadratic_descent.run(xd, d, want.cuda(), h.cuda(), stepsize=0.1)
  report('dictionary update', d, oracle.sc_dictionary_update(x, phi, want, h, stepsize=0.1), 1e-5)
  d = pd.clone()
  subspace_sc_cheap_quadratic_descent.run(xd, d, want.cuda(), groups, h.cuda(), 0.3, stepsize=0.1)
  report('aligned subspace update', d,
         oracle.sc_dictionary_update(x, phi, want, h, stepsize=0.1, group_assignments=groups, alignment_penalty=0.3), 1e-5)
  m_want = oracle.compute_metrics(x, want, phi, 0.9 * phi, 0.1)
  m_got = metrics.compute_metrics(xd, want.cuda(), pd, 0.9 * pd, 0.1)
  report('validation metrics', torch.tensor([float(np.mean(m_got[k])) for k in sorted(m_want)]),
         torch.tensor([float(np.mean(m_want[k])) for k in sorted(m_want)]), 1e-5)
  # convolutional: two channels, rectangular kernels / strides (one tile per tap) and the 16x16 / stride 8 family (halo)
  for shape in ((2, 2, 21, 30, (8, 12), (4, 6), 10), (2, 1, 40, 48, (16, 16), (8, 8), 24)):
    b, c, hh, ww, k, st, s = shape
    xi, pad = oracle.synthetic_padded_images(b, c, hh, ww, k, st)
    kern = oracle.synthetic_conv_dictionary(s, c, k[0], k[1])
    cw = oracle.conv_ista_fista(xi, kern, st, pad, 0.05, T)
    report('conv fista %dx%d stride %d' % (k[0], k[1], st[0]), conv_inf.run(xi.cuda(), kern.cuda(), st, pad, 0.05, T),
           cw, 1e-4)
    hc = oracle.conv_hessian_running_mean(torch.zeros(s), cw)
    dk = kern.cuda()
    conv_cheap.run(xi.cuda(), dk, cw.cuda(), hc.cuda(), st, pad, stepsize=0.05)
    report('conv dictionary update', dk,
           oracle.conv_sc_dictionary_update(xi, kern, cw, st, pad, hc, stepsize=0.05), 1e-5)
    cm_want = oracle.compute_metrics(xi, cw, kern, kern, 0.05, 'fista', kernel_strides=st, image_padding=pad)
    cm_got = metrics.compute_metrics(xi.cuda(), cw.cuda(), kern.cuda(), kern.cuda(), 0.05, 'fista', kernel_strides=st,
                                     image_padding=pad)
    report('conv validation metrics', torch.tensor([float(np.mean(cm_got[k_])) for k_ in sorted(cm_want)]),
           torch.tensor([float(np.mean(cm_want[k_])) for k_ in sorted(cm_want)]), 1e-5)
  images = torch.randn(3, 40, 52, 1, generator=torch.Generator().manual_seed(3))
  corners = torch.tensor([[0, 0, 0], [2, 24, 36], [1, 5, 7]], dtype=torch.int32)
  report('patch extraction', dataset_generation.extract_patches(images.cuda(), corners.cuda(), (16, 16)),
         oracle.extract_patches(images, corners, (16, 16)), 0.0)
  torch.cuda.synchronize()
  print('all calls ok')


if __name__ == '__main__':
  main()

"""Randomised parity of the convolutional path against the CPU oracle over shapes, strides, channels and thresholds.
usage: python tools/fuzz_conv.py [cases] [seed]"""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import vision_transform_codes_b200 as pkg
from vision_transform_codes_b200.analysis_transforms.convolutional import ista_fista
from vision_transform_codes_b200.dict_update_rules.convolutional import sc_cheap_quadratic_descent, sc_steepest_descent
from oracle import vtc_oracle as oracle

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
pkg.config.precision, pkg.config.update_precision = 'bf16x3', 'bf16x6'
bad = 0
for i in range(cases):
  sy, sx = rng.choice([2, 4, 8]), rng.choice([2, 4, 8])
  ty, tx = rng.choice([1, 2, 2, 3]), rng.choice([1, 2, 2, 3])
  k = (sy * ty, sx * tx)
  c = rng.choice([1, 1, 2, 3])
  s = rng.choice([4, 16, 24, 64, 80])
  b = rng.choice([1, 2, 5])
  h, w = rng.randint(k[0], 70), rng.randint(k[1], 70)
  variant = rng.choice(['ista', 'fista'])
  kw = rng.choice([{}, {'nonnegative_only': True}, {}])
  T = rng.choice([1, 2, 7, 25])
  x, pad = oracle.synthetic_padded_images(b, c, h, w, k, (sy, sx), seed=i)
  if pad[0][1] == 0 or pad[1][1] == 0:
    pad = None   # a trailing padding of 0 zeroes the reference's mask altogether; None is "no mask"
  phi = oracle.synthetic_conv_dictionary(s, c, k[0], k[1], seed=50 + i)
  want = oracle.conv_ista_fista(x, phi, (sy, sx), pad, 0.05, T, variant=variant, **kw)
  got = ista_fista.run(x.cuda(), phi.cuda(), (sy, sx), pad, 0.05, T, variant=variant, **kw).cpu()
  err = oracle.relative_l2(got, want) if float(want.abs().max()) > 0 else float(got.abs().max())
  hd = oracle.conv_hessian_running_mean(torch.zeros(s), want)
  rule = rng.choice(['cheap', 'steepest'])
  if float(want.abs().max()) > 0:
    if rule == 'cheap':
      want_phi = oracle.conv_sc_dictionary_update(x, phi, want, (sy, sx), pad, hd, stepsize=0.05)
      d = phi.cuda()
      sc_cheap_quadratic_descent.run(x.cuda(), d, want.cuda(), hd.cuda(), (sy, sx), pad, stepsize=0.05)
    else:
      want_phi = oracle.conv_sc_dictionary_update(x, phi, want, (sy, sx), pad, None, stepsize=0.05)
      d = phi.cuda()
      sc_steepest_descent.run(x.cuda(), d, want.cuda(), (sy, sx), pad, stepsize=0.05)
    derr = oracle.relative_l2(d.cpu(), want_phi)
  else:
    derr = 0.0
  ok = err <= 1e-4 and derr <= 2e-5 and bool(torch.isfinite(got).all())
  bad += not ok
  print('%s case %3d: b=%d c=%d %3dx%-3d s=%2d kernel %s stride (%d, %d) T=%2d %-5s %-8s codes %.2e  dictionary (%s) %.2e' %
        ('ok  ' if ok else 'FAIL', i, b, c, h, w, s, k, sy, sx, T, variant, 'nonneg' if kw else 'soft', err, rule, derr),
        flush=True)
print('%d of %d cases failed' % (bad, cases))
sys.exit(1 if bad else 0)

"""
Parity of the convolutional CUDA path (drop-in modules -> ctypes -> C ABI) with outputs of the reference
(tests/golden/conv_*.npz, made by make_golden.py from /root/reference) and with the CPU oracle on seeded inputs.
Same tolerances as the fully-connected path: relative L2 <= 1e-4 on codes (bf16x3 arithmetic), 1e-5 on dictionaries.
"""
import os

import pytest
import torch

from conftest import load_golden, record_parity
from oracle import vtc_oracle as oracle

pytestmark = pytest.mark.gpu

CODE_TOL = 1e-4
DICT_TOL = 1e-5


@pytest.fixture(autouse=True)
def default_precision():
  import vision_transform_codes_b200 as pkg
  saved = (pkg.config.precision, pkg.config.update_precision)
  pkg.config.precision, pkg.config.update_precision = 'bf16x3', 'bf16x6'
  yield
  pkg.config.precision, pkg.config.update_precision = saved


def modules():
  from vision_transform_codes_b200.analysis_transforms.convolutional import ista_fista
  from vision_transform_codes_b200.dict_update_rules.convolutional import sc_cheap_quadratic_descent, sc_steepest_descent
  return ista_fista, sc_cheap_quadratic_descent, sc_steepest_descent


def conv_args(g):
  stride = tuple(int(v) for v in g['stride'])
  padding = tuple(tuple(int(v) for v in row) for row in g['padding'])
  return g['images_padded'], g['dictionary'], stride, padding, g['sparsity_weight'], g['num_iters']


def check(got, want, tol=CODE_TOL, band=1e-4, case=''):
  got = got.cpu()
  assert got.shape == want.shape and got.dtype == torch.float32
  assert torch.isfinite(got).all()
  err = oracle.relative_l2(got, want)
  flips, outside = oracle.support_mismatches(got, want, band=0.0 if band is None else band)
  test = os.environ.get('PYTEST_CURRENT_TEST', '').split(' ')[0].split('::')[-1]
  record_parity(test, case, rel_l2=err, flips_total=flips, flips_in_band=flips - outside, flips_outside_band=outside,
                band=band, elements=got.numel(), tol=tol)
  assert err <= tol, err
  if band is not None:
    assert outside == 0, (flips, outside)
  return err


def test_conv_inference_call_matrix_against_reference_outputs():
  """The call matrix of the reference's tests/ista_fista_2.py."""
  ista_fista = modules()[0]
  g = load_golden('conv_small')
  x, phi, st, pad, lam, T = conv_args(g)
  xd, pd = x.cuda(), phi.cuda()
  keep_x, keep_p = xd.clone(), pd.clone()
  check(ista_fista.run(xd, pd, st, pad, lam, T, variant='fista'), g['fista'])
  check(ista_fista.run(xd, pd, st, pad, lam, T, variant='ista'), g['ista'])
  check(ista_fista.run(xd, g['plain_dictionary'].cuda(), st, pad, lam, T, variant='ista'), g['ista_plain'])
  check(ista_fista.run(xd, pd, st, pad, lam, T, nonnegative_only=True), g['fista_nonneg'])
  # hard thresholding at the ordinary tolerance: such calls run in config.hard_threshold_precision (bf16x6)
  got = ista_fista.run(xd, pd, st, pad, lam, T, variant='ista', nonnegative_only=True, hard_threshold=True)
  check(got, g['ista_hard_nonneg'], case='ista_hard_nonneg')
  check(ista_fista.run(xd, pd, st, pad, lam, 500, variant='ista', early_stopping_epsilon=1e-3), g['ista_early'],
        tol=2e-3, band=None)
  warm = g['warm_start'].cuda()
  keep_w = warm.clone()
  out = ista_fista.run(xd, pd, st, pad, lam, T, initial_codes=warm)
  check(out, g['fista_warm'])
  # the reference's own assertions (tests/ista_fista_2.py:56-68): nothing passed in is mutated
  assert torch.equal(xd, keep_x) and torch.equal(pd, keep_p) and torch.equal(warm, keep_w)
  assert not torch.allclose(out, warm)


PROX_VARIANTS = (('soft', {}), ('nonneg', {'nonnegative_only': True}), ('hard', {'hard_threshold': True}),
                 ('hard_nonneg', {'hard_threshold': True, 'nonnegative_only': True}))


@pytest.mark.parametrize('precision', ['bf16x3', 'bf16x6'])
def test_conv_single_step_of_every_threshold_variant_against_the_reference(precision):
  """ONE convolutional iteration from the reference's own iterate a_{T-1} for all four thresholds (analysis_transforms/
  convolutional/ista_fista.py:141-165): pins the prox kernels. Elements whose float64 pre-threshold value ties with the
  cutoff (relative band 1e-5) are excluded and counted."""
  import vision_transform_codes_b200 as pkg
  ista_fista = modules()[0]
  g = load_golden('conv_threshold_steps')
  x, phi, st, pad, lam, _ = conv_args(g)
  xd, pd = x.cuda(), phi.cuda()
  eta64 = 1.0 / float(torch.linalg.eigvalsh(phi.double().flatten(1).t() @ phi.double().flatten(1))[-1])
  mask = oracle.create_mask(x, pad).double()
  for variant in ('ista', 'fista'):
    for name, kw in PROX_VARIANTS:
      warm, want = g['%s_%s_warm' % (variant, name)], g['%s_%s_one' % (variant, name)]
      got, _ = ista_fista.infer(xd, pd, st, pad, lam, 1, variant, warm.cuda(), None, kw.get('nonnegative_only', False),
                                kw.get('hard_threshold', False), precision=pkg.PRECISIONS[precision])
      got = got.cpu()
      y = warm.double()
      resid = mask * (torch.nn.functional.conv_transpose2d(y, phi.double(), stride=st) - x.double())
      u = y - eta64 * torch.nn.functional.conv2d(resid, phi.double(), stride=st)
      mag = u if kw.get('nonnegative_only', False) else u.abs()
      ties = (mag - lam * eta64).abs() <= 1e-5 * torch.clamp(u.abs(), min=1.0)
      keep = ~ties
      err = float(torch.norm((got - want)[keep]) / torch.norm(want[keep]))
      flips = int((((got != 0) != (want != 0)) & keep).sum())
      record_parity('test_conv_single_step_of_every_threshold_variant_against_the_reference',
                    '%s %s %s' % (precision, variant, name), rel_l2=err, flips_outside_ties=flips,
                    ties_excluded=int(ties.sum()), tie_band=1e-5, elements=got.numel())
      assert err <= CODE_TOL, (precision, variant, name, err)
      assert flips == 0, (precision, variant, name, flips)


def test_conv_two_channels_rectangular_kernels():
  ista_fista, cheap, _ = modules()
  g = load_golden('conv_two_channel')
  x, phi, st, pad, lam, T = conv_args(g)
  check(ista_fista.run(x.cuda(), phi.cuda(), st, pad, lam, T), g['fista'])
  d = phi.cuda()
  cheap.run(x.cuda(), d, g['fista'].cuda(), g['hessian_diagonal'].cuda(), st, pad, stepsize=0.05)
  assert oracle.relative_l2(d.cpu(), g['cheap_1']) < DICT_TOL


def test_conv_dictionary_updates_against_reference_outputs():
  _, cheap, steepest = modules()
  g = load_golden('conv_small')
  x, phi, st, pad, _, _ = conv_args(g)
  xd, a, h = x.cuda(), g['fista'].cuda(), g['hessian_diagonal'].cuda()
  keep = (xd.clone(), a.clone(), h.clone())

  def updated(fn, *args, **kw):
    d = phi.cuda()
    assert fn(xd, d, a, *args, **kw) is None  # in place, returns None
    return d.cpu()

  assert oracle.relative_l2(updated(cheap.run, h, st, pad, stepsize=0.05), g['cheap_1']) < DICT_TOL
  assert oracle.relative_l2(updated(cheap.run, h, st, pad, stepsize=0.02, num_iters=2), g['cheap_2']) < 2 * DICT_TOL
  assert oracle.relative_l2(updated(steepest.run, st, pad, stepsize=0.05), g['steepest_1']) < DICT_TOL
  assert oracle.relative_l2(updated(steepest.run, st, pad, stepsize=0.05, normalize_dictionary=False),
                            g['steepest_unnormalized']) < DICT_TOL
  assert torch.equal(xd, keep[0]) and torch.equal(a, keep[1]) and torch.equal(h, keep[2])


def test_conv_hessian_running_mean():
  from vision_transform_codes_b200.dict_update_rules.convolutional import _common
  g = load_golden('conv_small')
  h = torch.zeros(g['fista'].size(1), device='cuda')
  _common.hessian_running_mean(h, g['fista'].cuda())
  assert oracle.relative_l2(h.cpu(), g['hessian_diagonal']) < 1e-6
  h2 = oracle.conv_hessian_running_mean(g['hessian_diagonal'], g['ista'])
  _common.hessian_running_mean(h, g['ista'].cuda())
  assert oracle.relative_l2(h.cpu(), h2) < 1e-6


def test_conv_training_matches_reference_trainer():
  """Three batches of train_dictionary in convolutional mode (reference tests/sparse_coding_4.py) through the drop-ins."""
  from vision_transform_codes_b200.dict_update_rules.convolutional import _common
  ista_fista, cheap, steepest = modules()
  g = load_golden('conv_training_small')
  pad = tuple(tuple(int(v) for v in row) for row in g['padding'])
  for variant, rule, key in (('ista', 'cheap', 'ista_cheap'), ('fista', 'steepest', 'fista_steepest')):
    phi = g['dictionary'].cuda()
    h = torch.zeros(phi.size(0), device='cuda')
    for x in g['batches']:
      x = x.cuda()
      codes = ista_fista.run(x, phi, (8, 8), pad, 0.05, 15, variant=variant)
      if rule == 'cheap':
        _common.hessian_running_mean(h, codes)
        cheap.run(x, phi, codes, h, (8, 8), pad, stepsize=0.05)
      else:
        steepest.run(x, phi, codes, (8, 8), pad, stepsize=0.05)
    assert oracle.relative_l2(phi.cpu(), g[key]) < 5e-5, key


@pytest.mark.parametrize('shape', [
    # (images, channels, height, width, kernels, kernel, stride)
    (2, 1, 64, 64, 64, (16, 16), (8, 8)),      # BASELINE configs[4] family: 64 filters of 16x16 at stride 8
    (3, 3, 30, 44, 20, (12, 8), (4, 4)),       # colour, 3x2 taps
    (1, 1, 40, 40, 16, (8, 8), (8, 8)),        # stride = kernel: no overlap (a single tap)
    (5, 2, 17, 23, 33, (6, 9), (2, 3)),        # 3x3 taps, sizes that need trailing padding
])
def test_conv_against_oracle_on_seeded_images(shape):
  ista_fista, cheap, _ = modules()
  b, c, h, w, s, k, st = shape
  x, pad = oracle.synthetic_padded_images(b, c, h, w, k, st)
  phi = oracle.synthetic_conv_dictionary(s, c, k[0], k[1])
  if pad[0][1] == 0 or pad[1][1] == 0:
    # a trailing padding of 0 makes the reference's create_mask zero everywhere (``mask[..., -0:] = 0``): all codes
    # stay zero, here as there; "no padding" is padding_dims=None
    assert float(oracle.conv_ista_fista(x, phi, st, pad, 0.05, 5).abs().max()) == 0.0
    assert float(ista_fista.run(x.cuda(), phi.cuda(), st, pad, 0.05, 5).abs().max()) == 0.0
    pad = None
  want = oracle.conv_ista_fista(x, phi, st, pad, 0.05, 30)
  got = ista_fista.run(x.cuda(), phi.cuda(), st, pad, 0.05, 30)
  check(got, want)
  hd = oracle.conv_hessian_running_mean(torch.zeros(s), want)
  want_phi = oracle.conv_sc_dictionary_update(x, phi, want, st, pad, hd, stepsize=0.05)
  d = phi.cuda()
  cheap.run(x.cuda(), d, want.cuda(), hd.cuda(), st, pad, stepsize=0.05)
  assert oracle.relative_l2(d.cpu(), want_phi) < DICT_TOL


def test_conv_known_answer_without_overlap():
  """stride = kernel size and orthonormal kernels: the convolutional code is the fully-connected code of every block,
  a = soft(<block, kernel>, lambda) after any number of iterations (step size 1)."""
  ista_fista = modules()[0]
  torch.manual_seed(4)
  q, _ = torch.linalg.qr(torch.randn(16, 16))
  phi = q.reshape(16, 1, 4, 4).contiguous()
  x = 0.5 * torch.randn(3, 1, 12, 20)
  blocks = x.unfold(2, 4, 4).unfold(3, 4, 4).reshape(3, 3, 5, 16)          # (b, i, j, pixels)
  want = oracle.threshold(torch.einsum('bijp,sp->bsij', blocks, q), 0.1)
  for variant in ('ista', 'fista'):
    got = ista_fista.run(x.cuda(), phi.cuda(), (4, 4), None, 0.1, 4, variant=variant).cpu()
    assert (got - want).abs().max() < 3e-5


def test_conv_error_behaviour():
  ista_fista = modules()[0]
  x = torch.zeros(1, 1, 32, 32).cuda()
  phi = oracle.synthetic_conv_dictionary(8, 1, 16, 16).cuda()
  pad = ((8, 8), (8, 8))
  with pytest.raises(AssertionError):
    ista_fista.run(x, phi, (8, 8), pad, 0.1, 3, variant='lista')
  with pytest.raises(UnboundLocalError):
    ista_fista.run(x, phi, (8, 8), pad, 0.1, 0)
  # (a kernel size that is not a multiple of the stride is supported: test_conv_kernel_not_a_multiple_of_the_stride)
  with pytest.raises(RuntimeError):
    ista_fista.run(torch.zeros(1, 1, 35, 32).cuda(), phi, (8, 8), pad, 0.1, 3)
  bad = phi.clone()
  bad[2, 0, 3, 3] = float('inf')
  with pytest.raises(RuntimeError):
    ista_fista.run(x, bad, (8, 8), pad, 0.1, 3)


def test_conv_kernel_not_a_multiple_of_the_stride():
  """The reference takes any kernel size whose strides tile the padded image (analysis_transforms/convolutional/
  ista_fista.py:119-122; ceil in utils/convolutions.py:14-15): 12 x 10 kernels at stride (8, 4), two channels. Codes,
  both dictionary updates and the Hessian mean against the CPU oracle (torch conv2d / conv_transpose2d)."""
  ista_fista, cheap, steepest = modules()
  from vision_transform_codes_b200.dict_update_rules.convolutional import _common
  g = torch.Generator().manual_seed(11)
  kh, kw, st = 12, 10, (8, 4)
  pad = ((4, 4), (6, 2))
  h, w = kh + 4 * st[0], kw + 9 * st[1]          # (h - kh) % sy == 0: 5 x 10 code positions
  img = torch.zeros(3, 2, h, w)
  img[:, :, pad[0][0]:h - pad[0][1], pad[1][0]:w - pad[1][1]] = 0.3 * torch.randn(
      3, 2, h - sum(pad[0]), w - sum(pad[1]), generator=g)
  phi = torch.randn(7, 2, kh, kw, generator=g)
  yy = (torch.arange(kh) - (kh - 1) / 2)[:, None] / (0.18 * kh)
  xx = (torch.arange(kw) - (kw - 1) / 2)[None, :] / (0.18 * kw)
  phi = phi * torch.exp(-0.5 * (yy**2 + xx**2))
  phi = phi / phi.flatten(1).norm(dim=1)[:, None, None, None]
  for variant, kw_ in (('fista', {}), ('ista', {'nonnegative_only': True})):
    want = oracle.conv_ista_fista(img, phi, st, pad, 0.05, 30, variant=variant, **kw_)
    got = ista_fista.run(img.cuda(), phi.cuda(), st, pad, 0.05, 30, variant=variant, **kw_)
    assert tuple(got.shape) == (3, 7, 5, 10)
    check(got, want, case='12x10 kernels, stride (8, 4), ' + variant)
  codes = oracle.conv_ista_fista(img, phi, st, pad, 0.05, 30)
  hd = torch.mean(torch.sum(codes**2, dim=(2, 3)), dim=0) / 100
  d = phi.cuda()
  cheap.run(img.cuda(), d, codes.cuda(), hd.cuda(), st, pad, stepsize=0.05, num_iters=2)
  assert oracle.relative_l2(d.cpu(), oracle.conv_sc_dictionary_update(img, phi, codes, st, pad, hd, stepsize=0.05,
                                                                    num_iters=2)) < DICT_TOL
  d = phi.cuda()
  steepest.run(img.cuda(), d, codes.cuda(), st, pad, stepsize=0.05)
  assert oracle.relative_l2(d.cpu(), oracle.conv_sc_dictionary_update(img, phi, codes, st, pad, None, stepsize=0.05)) < DICT_TOL
  assert tuple(_common.dictionary_gradient(img.cuda(), phi.cuda(), codes.cuda(), st, pad).shape) == (7, 2, kh, kw)


def test_conv_determinism_and_image_shard_independence():
  """What sharding a batch of images over GPUs relies on: every image's code depends on that image only."""
  ista_fista = modules()[0]
  x, pad = oracle.synthetic_padded_images(6, 1, 96, 96, (16, 16), (8, 8))
  phi = oracle.synthetic_conv_dictionary(64, 1, 16, 16).cuda()
  xd = x.cuda()
  a = ista_fista.run(xd, phi, (8, 8), pad, 0.05, 40)
  assert torch.equal(a, ista_fista.run(xd, phi, (8, 8), pad, 0.05, 40))
  assert torch.equal(a[3:], ista_fista.run(xd[3:], phi, (8, 8), pad, 0.05, 40))


def test_conv_size_independent_properties_at_benchmark_shape():
  """BASELINE configs[4] at its full per-GPU size (128 images of 512x512 padded to 528x528, 64 kernels of 16x16 at
  stride 8, codes 64x65x65, 300 iterations), which the oracle cannot cover: determinism, image-shard independence, codes
  confined to what the mask lets through, and oracle parity on one image."""
  ista_fista = modules()[0]
  x, pad = oracle.synthetic_padded_images(128, 1, 512, 512, (16, 16), (8, 8))
  phi = oracle.synthetic_conv_dictionary(64, 1, 16, 16)
  xd, pd = x.cuda(), phi.cuda()
  a = ista_fista.run(xd, pd, (8, 8), pad, 0.05, 300)
  assert tuple(a.shape) == (128, 64, 65, 65) and torch.isfinite(a).all()
  assert torch.equal(a, ista_fista.run(xd, pd, (8, 8), pad, 0.05, 300))
  assert torch.equal(a[96:], ista_fista.run(xd[96:], pd, (8, 8), pad, 0.05, 300))
  assert torch.equal(a[5:6], ista_fista.run(xd[5:6], pd, (8, 8), pad, 0.05, 300))
  check(a[5:6], oracle.conv_ista_fista(x[5:6], phi, (8, 8), pad, 0.05, 300))


def test_conv_lean_trainer_matches_reference_trainer(tmp_path):
  """The package's own train_dictionary in convolutional mode, with the reference's parameter dictionary
  (tests/sparse_coding_4.py), against the unmodified reference trainer's result; checkpoints in its pickle format."""
  import numpy as np
  from vision_transform_codes_b200.lean import sparse_coding as trainer
  g = load_golden('conv_training_small')
  pad = tuple(tuple(int(v) for v in row) for row in g['padding'])
  params = {
      'mode': 'convolutional', 'num_epochs': 1, 'strides': (8, 8), 'padding': pad,
      'code_inference_algorithm': 'ista',
      'inference_param_schedule': {0: {'sparsity_weight': 0.05, 'num_iters': 15}},
      'dictionary_update_algorithm': 'sc_cheap_quadratic_descent',
      'dict_update_param_schedule': {0: {'stepsize': 0.05, 'num_iters': 1}},
      'checkpoint_schedule': {2}, 'logging_folder_fullpath': tmp_path}
  phi = g['dictionary'].cuda()
  trainer.train_dictionary([x.cuda() for x in g['batches']], None, phi, params)
  assert oracle.relative_l2(phi.cpu(), g['ista_cheap']) < 5e-5
  ckpt = trainer.load_newest_dictionary_checkpoint(tmp_path)
  assert isinstance(ckpt, np.ndarray) and ckpt.dtype == np.float32 and ckpt.shape == tuple(phi.shape)
  phi2 = g['dictionary'].cuda()
  trainer.train_dictionary([x.cuda() for x in g['batches']], None, phi2,
                           dict(params, code_inference_algorithm='fista',
                                dictionary_update_algorithm='sc_steepest_descent', checkpoint_schedule=None))
  assert oracle.relative_l2(phi2.cpu(), g['fista_steepest']) < 5e-5


@pytest.mark.parametrize('env', [{'VTC_B200_CONV_HALO': '0'}, {'VTC_B200_CONV_RESIDENT': '0'}])
def test_conv_tile_variants_agree(env):
  """The convolutional launches have three tile variants (one tile per tap on 128-wide tiles; resident dictionary
  operand on 64-wide tiles; + halo staging of the taps, the default). The switches are read once per process, so the
  other two run in a subprocess; all three must agree with the oracle."""
  import os
  import subprocess
  import sys
  from conftest import ROOT
  code = (
      "import sys; sys.path.insert(0, %r)\n"
      "import torch\n"
      "from oracle import vtc_oracle as oracle\n"
      "from vision_transform_codes_b200.analysis_transforms.convolutional import ista_fista\n"
      "x, pad = oracle.synthetic_padded_images(3, 1, 96, 80, (16, 16), (8, 8))\n"
      "phi = oracle.synthetic_conv_dictionary(64, 1, 16, 16)\n"
      "want = oracle.conv_ista_fista(x, phi, (8, 8), pad, 0.05, 30)\n"
      "got = ista_fista.run(x.cuda(), phi.cuda(), (8, 8), pad, 0.05, 30).cpu()\n"
      "err = oracle.relative_l2(got, want)\n"
      "assert err <= 1e-4, err\n"
      "print('VARIANT_OK', err)\n" % ROOT)
  out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, **env))
  assert out.returncode == 0 and 'VARIANT_OK' in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_conv_c_abi_writes_only_the_codes_it_owns():
  """vtc_fista_conv writing its codes into the middle of a larger buffer: the guard zones keep their canaries, the
  padded images, the dictionary and the warm start are not written, and the codes equal the front end's."""
  import ctypes
  import vision_transform_codes_b200 as pkg
  from vision_transform_codes_b200 import _lib
  from vision_transform_codes_b200.analysis_transforms.convolutional import ista_fista
  lib = _lib.load()
  dev = torch.device('cuda:0')
  prec = pkg.config.precision_code()
  canary = 777.0
  for (b, c, h, w, s, k, st, warm) in ((3, 1, 40, 56, 24, (8, 8), (4, 4), False), (2, 2, 32, 32, 16, (16, 16), (8, 8), True),
                                       (5, 1, 64, 48, 64, (16, 16), (8, 8), False)):
    x, pad = oracle.synthetic_padded_images(b, c, h, w, k, st, seed=b)
    x = x.to(dev)
    phi = oracle.synthetic_conv_dictionary(s, c, k[0], k[1]).to(dev)
    geo = ista_fista.geometry(x, phi, st, pad)
    B, C, H, W, S, KH, KW, SY, SX, pt, pb, pl, pr, SH, SW = geo
    init = None
    if warm:
      init = (0.01 * torch.randn(B, S, SH, SW, generator=torch.Generator().manual_seed(4))).to(dev)
    want = ista_fista.run(x, phi, st, pad, 0.05, 10, initial_codes=init)
    n, guard = B * S * SH * SW, 4096
    buf = torch.full((guard + n + guard,), canary, device=dev)
    out = buf[guard:guard + n].view(B, S, SH, SW)
    x0, phi0 = x.clone(), phi.clone()
    init0 = init.clone() if warm else None
    with torch.cuda.device(dev):
      nbytes = lib.vtc_fista_conv_workspace_bytes(B, C, H, W, S, KH, KW, SY, SX, prec)
      assert nbytes > 0
      ws = _lib.workspace(nbytes, dev, 'fista_conv_guard_test')
      iters = ctypes.c_int(0)
      _lib.check(lib.vtc_fista_conv(_lib.ptr(x), _lib.ptr(phi), _lib.ptr(init), _lib.ptr(out), B, C, H, W, S, KH, KW, SY, SX,
                                    pt, pb, pl, pr, 0.05, 10, 1, 0, 0, -1.0, prec, _lib.ptr(ws), ws.numel(),
                                    ctypes.byref(iters), None, _lib.stream_ptr(dev)))
    torch.cuda.synchronize()
    assert torch.equal(out, want), (b, c, h, w, s, k, st)
    assert bool((buf[:guard] == canary).all()) and bool((buf[guard + n:] == canary).all()), 'guard zone written'
    assert torch.equal(x, x0) and torch.equal(phi, phi0)
    if warm:
      assert torch.equal(init, init0)

"""
Device-side validation metrics (training/metrics.py -> vtc_sc_metrics / vtc_sc_conv_metrics /
vtc_dict_change) against the scalars the unmodified reference trainer logged (tests/golden/metrics_*.npz) and against
the CPU oracle's compute_metrics on identical codes. Tolerances: 1e-5 relative when both sides see the same codes (the
metrics are sums of a bf16x6 residual); 2e-3 when the codes come from the device's own bf16x3 inference (support flips
of near-threshold coefficients move the l0 count).
"""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import vtc_oracle as oracle
from test_oracle import metrics_cases, oracle_validation_codes

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def default_precision():
  import vision_transform_codes_b200 as pkg
  saved = (pkg.config.precision, pkg.config.update_precision)
  pkg.config.precision, pkg.config.update_precision = 'bf16x3', 'bf16x6'
  yield
  pkg.config.precision, pkg.config.update_precision = saved


def trainer_modules():
  from vision_transform_codes_b200.lean import metrics, sparse_coding
  return metrics, sparse_coding


def assert_metrics_close(got, want, tol, context):
  for name in want:
    g, w = np.mean(got[name]), np.mean(want[name])
    assert abs(g - w) <= tol * max(1.0, abs(w)), (context, name, g, w)


def test_metrics_on_identical_codes_match_the_oracle():
  metrics, _ = trainer_modules()
  for name, alg, kw in metrics_cases():
    g = load_golden(name)
    kw = dict(kw)
    if 'padding' in g:
      kw['image_padding'] = tuple(tuple(int(v) for v in row) for row in g['padding'])
    phi, prev = g['dictionary_iter_2'], g['dictionary_iter_1']
    for x in g['validation']:
      codes = oracle_validation_codes(g, alg, phi, kw, x)
      want = oracle.compute_metrics(x, codes, phi, prev, g['sparsity_weight'], alg, **kw)
      got = metrics.compute_metrics(x.cuda(), codes.cuda(), phi.cuda(), prev.cuda(), g['sparsity_weight'], alg, **kw)
      assert set(got) == set(want)
      assert_metrics_close(got, want, 1e-5, name)
      assert np.allclose(got[metrics.CHANGE], want[metrics.CHANGE], rtol=1e-5, atol=1e-8)


def test_lean_trainer_logs_the_reference_trainers_validation_metrics(tmp_path):
  """train_dictionary with a 'training_visualization_schedule': the (iteration, metrics) pairs against what the
  unmodified trainer sent to tensorboard at iterations 0 and 2 (tests/golden/make_golden.py:make_metrics)."""
  _, trainer = trainer_modules()
  for name, alg, kw in metrics_cases():
    g = load_golden(name)
    conv = 'padding' in g
    train = g['training']
    if conv:
      padb = tuple(tuple(int(v) for v in row) for row in g['padding'])
      params = {'mode': 'convolutional', 'strides': (8, 8), 'padding': padb, 'code_inference_algorithm': 'ista',
                'dictionary_update_algorithm': 'sc_cheap_quadratic_descent',
                'dict_update_param_schedule': {0: {'stepsize': 0.05, 'num_iters': 1}}}
    else:
      params = {'mode': 'fully-connected', 'code_inference_algorithm': alg,
                'dictionary_update_algorithm': 'sc_cheap_quadratic_descent',
                'dict_update_param_schedule': {0: {'stepsize': 0.1, 'num_iters': 1}}}
      if alg.startswith('subspace'):
        params.update(group_assignments=kw['group_assignments'], subspace_alignment_penalty=0.0,
                      dictionary_update_algorithm='subspace_sc_cheap_quadratic_descent')
    log = []
    params.update(num_epochs=1, training_visualization_schedule={0, 2}, validation_metrics_log=log,
                  logging_folder_fullpath=tmp_path / name,
                  inference_param_schedule={0: {'sparsity_weight': g['sparsity_weight'], 'num_iters': g['num_iters']}})
    phi = g['dictionary'].cuda()
    trainer.train_dictionary(train.cuda(), g['validation'].cuda(), phi, params)
    assert [k for k, _ in log] == [0, 2]
    names = [str(n) for n in g['metric_names']]
    for row, (_, got) in enumerate(log):
      want = {n: float(g['metrics'][row, col]) for col, n in enumerate(names)}
      assert_metrics_close(got, want, 2e-3, (name, row))
    assert oracle.relative_l2(phi.cpu(), g['dictionary_iter_2']) > 0   # three updates in, not two
    assert (tmp_path / name).exists()


def test_metrics_ragged_shapes_pitched_rows_and_exact_reconstructions():
  """Shapes off every tile size, row-pitched inputs, and patches reconstructed exactly (mse == 0: left out of the pSNR
  mean, training/sparse_coding.py:223)."""
  metrics, _ = trainer_modules()
  B, D, S = 1000, 100, 300
  phi = oracle.synthetic_dictionary(S, D)
  x = oracle.synthetic_patches(B, D)
  codes = oracle.ista_fista(x, phi, 0.1, 15)
  codes[5] = 0
  x[5] = 0          # zero codes, zero patch: the residual is exactly zero on both sides
  codes[17] = 0
  codes[17, 3] = 1.0
  x[17] = phi[3]    # one unit coefficient: bf16x6 reproduces the atom exactly
  want = oracle.compute_metrics(x, codes, phi, 0.9 * phi, 0.25)
  wide_x = torch.zeros(B, D + 28).cuda()
  wide_x[:, :D] = x
  wide_a = torch.zeros(B, S + 4).cuda()
  wide_a[:, :S] = codes
  got = metrics.compute_metrics(wide_x[:, :D], wide_a[:, :S], phi.cuda(), (0.9 * phi).cuda(), 0.25)
  assert_metrics_close(got, want, 1e-5, 'ragged')
  totals = metrics.batch_totals(wide_x[:, :D], wide_a[:, :S], phi.cuda()).cpu()
  assert int(totals[4]) == B - 2 and int(totals[7]) == B
  again = metrics.batch_totals(wide_x[:, :D], wide_a[:, :S], phi.cuda()).cpu()
  assert torch.equal(totals, again)   # fixed summation order


def test_conv_metrics_full_size_image_and_no_padding():
  """One image of BASELINE configs[4]'s geometry (512x512 -> 528x528, 64 kernels of 16x16, stride 8), and an un-padded
  call (image_padding None: nothing is cropped, training/sparse_coding.py:188)."""
  metrics, _ = trainer_modules()
  x, pad = oracle.synthetic_padded_images(2, 1, 512, 512, (16, 16), (8, 8))
  phi = oracle.synthetic_conv_dictionary(64, 1, 16, 16)
  codes = oracle.conv_ista_fista(x, phi, (8, 8), pad, 0.05, 3, variant='ista')
  want = oracle.compute_metrics(x, codes, phi, phi, 0.05, 'ista', kernel_strides=(8, 8), image_padding=pad)
  got = metrics.compute_metrics(x.cuda(), codes.cuda(), phi.cuda(), phi.cuda(), 0.05, 'ista', kernel_strides=(8, 8),
                                image_padding=pad)
  assert_metrics_close(got, want, 1e-5, 'configs[4] geometry')
  x2 = oracle.synthetic_patches(3, 48 * 40).view(3, 1, 48, 40)
  phi2 = oracle.synthetic_conv_dictionary(24, 1, 16, 16)
  codes2 = oracle.conv_ista_fista(x2, phi2, (8, 8), None, 0.05, 5, variant='ista')
  want2 = oracle.compute_metrics(x2, codes2, phi2, phi2, 0.05, 'ista', kernel_strides=(8, 8), image_padding=None)
  got2 = metrics.compute_metrics(x2.cuda(), codes2.cuda(), phi2.cuda(), phi2.cuda(), 0.05, 'ista',
                                 kernel_strides=(8, 8), image_padding=None)
  assert_metrics_close(got2, want2, 1e-5, 'no padding')
  with pytest.raises(ValueError):
    metrics.compute_metrics(x.cuda(), codes.cuda(), phi.cuda(), phi.cuda(), 0.05, 'ista', kernel_strides=(8, 8),
                            image_padding=((8, 0), (8, 8)))

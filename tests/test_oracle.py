"""The CPU oracle against outputs of the reference itself (tests/golden, made by make_golden.py) and closed forms."""
import numpy as np
import torch

from conftest import load_golden, ragged_groups
from oracle import vtc_oracle as oracle

TIGHT = 2e-6  # same torch ops in the same order: only eigensolver / summation-order noise is left


def close(a, b, tol=TIGHT):
  assert a.shape == b.shape
  err = oracle.relative_l2(a, b)
  assert err <= tol, err


def test_inference_matches_reference_outputs():
  g = load_golden('inference_small')
  x, phi, lam, T = g['images'], g['dictionary'], g['sparsity_weight'], g['num_iters']
  close(oracle.ista_fista(x, phi, lam, T, variant='fista'), g['fista'])
  close(oracle.ista_fista(x, phi, lam, T, variant='ista'), g['ista'])
  close(oracle.ista_fista(x, phi, lam, T, nonnegative_only=True), g['fista_nonneg'])
  close(oracle.ista_fista(x, phi, lam, T, hard_threshold=True), g['fista_hard'], 1e-5)
  close(oracle.ista_fista(x, phi, lam, T, variant='ista', hard_threshold=True, nonnegative_only=True),
        g['ista_hard_nonneg'], 1e-5)
  close(oracle.ista_fista(x, phi, lam, T, initial_codes=g['warm_start']), g['fista_warm'])
  close(oracle.ista_fista(x, phi, lam, 1000, variant='ista', early_stopping_epsilon=1e-3), g['ista_early'])
  close(oracle.ista_fista(x, phi, lam, 1000, variant='fista', early_stopping_epsilon=1e-3), g['fista_early'])


def test_inference_baseline_configs():
  for name in ('inference_config1', 'inference_overcomplete'):
    g = load_golden(name)
    close(oracle.ista_fista(g['images'], g['dictionary'], g['sparsity_weight'], g['num_iters']), g['fista'], 1e-5)


def test_subspace_matches_reference_outputs():
  g = load_golden('subspace_small')
  x, phi, lam, T = g['images'], g['dictionary'], g['sparsity_weight'], g['num_iters']
  s = phi.size(0)
  pairs = [list(map(int, v)) for v in np.array_split(np.arange(s), s // 2)]
  quads = [list(map(int, v)) for v in np.array_split(np.arange(s), s // 4)]
  close(oracle.subspace_ista_fista(x, phi, pairs, lam, T), g['pairs_fista'])
  close(oracle.subspace_ista_fista(x, phi, pairs, lam, T, variant='ista'), g['pairs_ista'])
  close(oracle.subspace_ista_fista(x, phi, quads, lam, T), g['quads_fista'])
  close(oracle.subspace_ista_fista(x, phi, ragged_groups(g), lam, T), g['ragged_fista'])
  close(oracle.subspace_ista_fista(x, phi, pairs, lam, T, initial_codes=g['warm_start']), g['pairs_warm'])
  close(oracle.subspace_ista_fista(x, phi, pairs, lam, 1000, variant='ista', early_stopping_epsilon=1e-3),
        g['pairs_early'])


def test_dictionary_update_matches_reference_outputs():
  g = load_golden('dict_update_small')
  x, phi, a, h = g['images'], g['dictionary'], g['codes'], g['hessian_diagonal']
  close(oracle.sc_dictionary_update(x, phi, a, h, stepsize=0.1), g['cheap_1'])
  close(oracle.sc_dictionary_update(x, phi, a, h, stepsize=0.05, num_iters=3), g['cheap_3'])
  close(oracle.sc_dictionary_update(x, phi, a, None, stepsize=0.1), g['steepest_1'])
  close(oracle.sc_dictionary_update(x, phi, a, None, stepsize=0.1, num_iters=2, normalize_dictionary=False),
        g['steepest_2_unnormalized'])
  s = phi.size(0)
  pairs = [list(map(int, v)) for v in np.array_split(np.arange(s), s // 2)]
  overlapping = [[0, 2, 5], [1, 7], [2, 3, 4, 5], [6, 7, 8, 9, 10], [11, 12]]
  close(oracle.sc_dictionary_update(x, phi, a, h, stepsize=0.1, group_assignments=pairs, alignment_penalty=0.5),
        g['aligned_pairs'])
  close(oracle.sc_dictionary_update(x, phi, a, h, stepsize=0.05, num_iters=2, group_assignments=overlapping,
                                    alignment_penalty=0.25), g['aligned_overlapping_2'])
  close(oracle.sc_dictionary_update(x, g['unnormalized_in'], a, h, stepsize=0.05, normalize_dictionary=False,
                                    group_assignments=overlapping, alignment_penalty=0.25), g['aligned_unnormalized'])


def test_train_steps_match_reference_trainer():
  g = load_golden('training_small')
  batches, phi0 = g['batches'], g['dictionary']
  s = phi0.size(0)
  phi, _, _ = oracle.train_steps(batches, phi0, 0.1, 30, 0.1)
  close(phi, g['fista_cheap'], 1e-5)
  phi, _, _ = oracle.train_steps(batches, phi0, 0.1, 30, 0.1, variant='ista', update_rule='sc_steepest_descent')
  close(phi, g['ista_steepest'], 1e-5)
  pairs = [list(map(int, v)) for v in np.array_split(np.arange(s), s // 2)]
  phi, _, _ = oracle.train_steps(batches, phi0, 0.1, 30, 0.1, group_assignments=pairs)
  close(phi, g['subspace_cheap'], 1e-5)
  phi, _, _ = oracle.train_steps(batches, phi0, 0.1, 30, 0.1, group_assignments=pairs,
                                 update_rule='subspace_sc_cheap_quadratic_descent', alignment_penalty=0.3)
  close(phi, g['subspace_cheap_aligned'], 1e-5)


def test_known_answer_orthonormal_dictionary():
  # Q orthonormal => L = 1, stepsize = 1, and one step from zero is already the fixed point soft(x Q^T, lambda)
  torch.manual_seed(3)
  q, _ = torch.linalg.qr(torch.randn(32, 32))
  x = 0.5 * torch.randn(20, 32)
  want = oracle.threshold(x @ q.t(), 0.1)
  for variant in ('ista', 'fista'):
    for iters in (1, 7):
      got = oracle.ista_fista(x, q, 0.1, iters, variant=variant)
      assert (got - want).abs().max() < 2e-5


def test_known_answer_groups_of_one_equal_vanilla():
  phi = oracle.synthetic_dictionary(48, 24)
  x = oracle.synthetic_patches(16, 24)
  groups = [[i] for i in range(48)]
  a = oracle.subspace_ista_fista(x, phi, groups, 0.1, 40)
  b = oracle.ista_fista(x, phi, 0.1, 40)
  assert oracle.relative_l2(a, b) < 1e-5


def test_known_answer_zero_codes_leave_dictionary_unchanged():
  phi = oracle.synthetic_dictionary(16, 8)
  x = oracle.synthetic_patches(10, 8)
  out = oracle.sc_dictionary_update(x, phi, torch.zeros(10, 16), torch.zeros(16), stepsize=0.1)
  assert oracle.relative_l2(out, phi) < 1e-6


def test_inputs_not_mutated():
  # the reference's own assertions: tests/ista_fista_1.py:45-54
  phi = oracle.synthetic_dictionary(32, 16)
  x = oracle.synthetic_patches(8, 16)
  warm = oracle.ista_fista(x, phi, 0.1, 3)
  keep = (x.clone(), phi.clone(), warm.clone())
  out = oracle.ista_fista(x, phi, 0.1, 10, initial_codes=warm)
  assert torch.equal(x, keep[0]) and torch.equal(phi, keep[1]) and torch.equal(warm, keep[2])
  assert not torch.allclose(out, warm)


def test_whitened_generator_statistics():
  x = oracle.synthetic_patches(2048, 256, kind='whitened')
  assert x.shape == (2048, 256)
  assert abs(float(x.std()) - 0.3) < 1e-3


def conv_args(g):
  stride = tuple(int(v) for v in g['stride'])
  padding = tuple(tuple(int(v) for v in row) for row in g['padding'])
  return g['images_padded'], g['dictionary'], stride, padding, g['sparsity_weight'], g['num_iters']


def test_conv_inference_matches_reference_outputs():
  """Call matrix of the reference's tests/ista_fista_2.py on outputs of the reference's convolutional path."""
  g = load_golden('conv_small')
  x, phi, st, pad, lam, T = conv_args(g)
  close(oracle.conv_ista_fista(x, phi, st, pad, lam, T, variant='fista'), g['fista'])
  close(oracle.conv_ista_fista(x, phi, st, pad, lam, T, variant='ista'), g['ista'])
  close(oracle.conv_ista_fista(x, g['plain_dictionary'], st, pad, lam, T, variant='ista'), g['ista_plain'])
  close(oracle.conv_ista_fista(x, phi, st, pad, lam, 500, variant='ista', early_stopping_epsilon=1e-3), g['ista_early'])
  close(oracle.conv_ista_fista(x, phi, st, pad, lam, T, nonnegative_only=True), g['fista_nonneg'])
  close(oracle.conv_ista_fista(x, phi, st, pad, lam, T, variant='ista', nonnegative_only=True, hard_threshold=True),
        g['ista_hard_nonneg'], 1e-5)
  close(oracle.conv_ista_fista(x, phi, st, pad, lam, T, initial_codes=g['warm_start']), g['fista_warm'])
  g2 = load_golden('conv_two_channel')
  x, phi, st, pad, lam, T = conv_args(g2)
  close(oracle.conv_ista_fista(x, phi, st, pad, lam, T), g2['fista'])


def test_conv_dictionary_update_matches_reference_outputs():
  g = load_golden('conv_small')
  x, phi, st, pad, _, _ = conv_args(g)
  a, h = g['fista'], g['hessian_diagonal']
  close(oracle.conv_hessian_running_mean(torch.zeros_like(h), a), h)
  close(oracle.conv_sc_dictionary_update(x, phi, a, st, pad, h, stepsize=0.05), g['cheap_1'])
  close(oracle.conv_sc_dictionary_update(x, phi, a, st, pad, h, stepsize=0.02, num_iters=2), g['cheap_2'])
  close(oracle.conv_sc_dictionary_update(x, phi, a, st, pad, None, stepsize=0.05), g['steepest_1'])
  close(oracle.conv_sc_dictionary_update(x, phi, a, st, pad, None, stepsize=0.05, normalize_dictionary=False),
        g['steepest_unnormalized'])
  g2 = load_golden('conv_two_channel')
  x, phi, st, pad, _, _ = conv_args(g2)
  close(oracle.conv_sc_dictionary_update(x, phi, g2['fista'], st, pad, g2['hessian_diagonal'], stepsize=0.05),
        g2['cheap_1'])


def test_conv_training_matches_reference_trainer():
  """Three batches of the unmodified train_dictionary in convolutional mode (reference tests/sparse_coding_4.py)."""
  g = load_golden('conv_training_small')
  pad = tuple(tuple(int(v) for v in row) for row in g['padding'])
  for variant, rule, key in (('ista', 'cheap', 'ista_cheap'), ('fista', 'steepest', 'fista_steepest')):
    phi = g['dictionary'].clone()
    h = torch.zeros(phi.size(0))
    for x in g['batches']:
      codes = oracle.conv_ista_fista(x, phi, (8, 8), pad, 0.05, 15, variant=variant)
      if rule == 'cheap':
        h = oracle.conv_hessian_running_mean(h, codes)
        phi = oracle.conv_sc_dictionary_update(x, phi, codes, (8, 8), pad, h, stepsize=0.05)
      else:
        phi = oracle.conv_sc_dictionary_update(x, phi, codes, (8, 8), pad, None, stepsize=0.05)
    close(phi, g[key], 1e-5)


def test_conv_padding_helpers():
  assert oracle.get_padding_amt(512, 16, 8) == (8, 8)
  assert oracle.get_padding_amt(21, 8, 4) == (4, 7)
  assert oracle.code_dim_from_padded_img_dim(528, 16, 8) == 65
  x, pad = oracle.synthetic_padded_images(2, 1, 40, 48, (16, 16), (8, 8))
  assert tuple(x.shape) == (2, 1, 56, 64) and pad == ((8, 8), (8, 8))
  assert float(x[:, :, :8].abs().max()) == 0.0 and float(x[:, :, :, -8:].abs().max()) == 0.0


def test_patch_extraction_oracle_is_the_literal_crop():
  g = torch.Generator().manual_seed(2)
  images = torch.randn(2, 20, 24, 3, generator=g)
  corners = torch.tensor([[0, 0, 0], [1, 4, 8], [1, 12, 16], [0, 7, 3]], dtype=torch.int32)
  got = oracle.extract_patches(images, corners, (8, 8))
  assert tuple(got.shape) == (4, 8 * 8 * 3)
  assert torch.equal(got[1].view(8, 8, 3), images[1, 4:12, 8:16])
  assert torch.equal(got[3].view(8, 8, 3)[2, 5], images[0, 9, 8])


def metrics_cases():
  """(golden name, oracle inference, compute_metrics keyword arguments) of the three validation-metric goldens."""
  pairs = [list(map(int, g)) for g in np.array_split(np.arange(128), 64)]
  yield 'metrics_fc', 'fista', {}
  yield 'metrics_subspace', 'subspace_fista', {'group_assignments': pairs}
  yield 'metrics_conv', 'ista', {'kernel_strides': (8, 8)}


def oracle_validation_codes(g, alg, phi, kw, x):
  lam, T = g['sparsity_weight'], g['num_iters']
  if 'kernel_strides' in kw:
    return oracle.conv_ista_fista(x, phi, kw['kernel_strides'], kw['image_padding'], lam, T, variant=alg)
  if alg.startswith('subspace'):
    return oracle.subspace_ista_fista(x, phi, kw['group_assignments'], lam, T, variant=alg[9:])
  return oracle.ista_fista(x, phi, lam, T, variant=alg)


def test_validation_metrics_match_reference_trainer():
  """compute_metrics (training/sparse_coding.py:177-229) against the scalars the unmodified trainer sent to its
  tensorboard writer at iterations 0 and 2, averaged over two validation batches (:505-506)."""
  for name, alg, kw in metrics_cases():
    g = load_golden(name)
    kw = dict(kw)
    if 'padding' in g:
      kw['image_padding'] = tuple(tuple(int(v) for v in row) for row in g['padding'])
    names = [str(n) for n in g['metric_names']]
    for row, (phi, prev) in enumerate(((g['dictionary'], g['dictionary']),
                                       (g['dictionary_iter_2'], g['dictionary_iter_1']))):
      per_batch = []
      for x in g['validation']:
        codes = oracle_validation_codes(g, alg, phi, kw, x)
        per_batch.append(oracle.compute_metrics(x, codes, phi, prev, g['sparsity_weight'], alg, **kw))
      for col, metric in enumerate(names):
        got = np.mean([m[metric] for m in per_batch])
        want = float(g['metrics'][row, col])
        assert abs(got - want) <= 2e-5 * max(1.0, abs(want)), (name, row, metric, got, want)


def test_whitening_matches_reference_outputs():
  """utils/image_processing.py:267-308 on the committed reference outputs: transfer function and filtered images, the
  dataset defaults and an un-normalised filter, even and odd DFT sizes, one and three colour channels."""
  g = load_golden('whitening_small')
  cut = {'low': 1e-3, 'high': 0.9}
  for key in ('gray', 'colour'):
    img = g[key].numpy()
    assert np.allclose(oracle.whitening_filter(img.shape, cut), g[key + '_filter'].numpy(), rtol=1e-12, atol=0)
    got = oracle.whiten_center_surround(img, cut)
    assert got.dtype == np.float32 and np.allclose(got, g[key + '_whitened'].numpy(), rtol=0, atol=1e-6)
  raw = {'low': 0.05, 'high': 0.6}
  assert np.allclose(oracle.whitening_filter((40, 52), raw, False), g['raw_filter'].numpy(), rtol=1e-12, atol=0)
  assert np.allclose(oracle.whiten_center_surround(g['gray'].numpy(), raw, False), g['raw_whitened'].numpy(),
                     rtol=0, atol=1e-6)

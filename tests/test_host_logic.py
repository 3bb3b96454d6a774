"""Host-side logic of the drop-in modules that needs no GPU: signatures identical to the reference's (SURVEY.md 8b and
analysis_transforms/convolutional/ista_fista.py:17-21, dict_update_rules/convolutional/*.py), geometry and padding rules of
the convolutional path, refusal of CPU tensors, workspace queries."""
import inspect
import os

import numpy as np

import pytest
import torch

from oracle import vtc_oracle as oracle


def params(fn):
  return [(p.name, p.default if p.default is not inspect._empty else '<required>')
          for p in inspect.signature(fn).parameters.values()]


def test_signatures_match_the_reference():
  from vision_transform_codes_b200.analysis_transforms.convolutional import ista_fista as conv_inf
  from vision_transform_codes_b200.analysis_transforms.fully_connected import ista_fista, subspace_ista_fista
  from vision_transform_codes_b200.dict_update_rules.convolutional import sc_cheap_quadratic_descent as conv_cheap
  from vision_transform_codes_b200.dict_update_rules.convolutional import sc_steepest_descent as conv_steep
  from vision_transform_codes_b200.dict_update_rules.fully_connected import (
      sc_cheap_quadratic_descent, sc_steepest_descent, subspace_sc_cheap_quadratic_descent)
  R = '<required>'
  assert params(ista_fista.run) == [
      ('images', R), ('dictionary', R), ('sparsity_weight', R), ('num_iters', R), ('variant', 'fista'),
      ('initial_codes', None), ('early_stopping_epsilon', None), ('nonnegative_only', False), ('hard_threshold', False)]
  assert params(subspace_ista_fista.run) == [
      ('images', R), ('dictionary', R), ('group_assignments', R), ('sparsity_weight', R), ('num_iters', R),
      ('variant', 'fista'), ('ret_summed_gduplicates', True), ('initial_codes', None),
      ('early_stopping_epsilon', None), ('hard_threshold', False)]
  assert params(sc_cheap_quadratic_descent.run) == [
      ('images', R), ('dictionary', R), ('codes', R), ('hessian_diagonal', R), ('stepsize', 0.001), ('num_iters', 1),
      ('lowest_code_val', 0.001), ('normalize_dictionary', True)]
  assert params(sc_steepest_descent.run) == [
      ('images', R), ('dictionary', R), ('codes', R), ('stepsize', 0.001), ('num_iters', 1),
      ('normalize_dictionary', True)]
  assert params(subspace_sc_cheap_quadratic_descent.run) == [
      ('images', R), ('dictionary', R), ('codes', R), ('group_assignments', R), ('hessian_diagonal', R),
      ('alignment_penalty', R), ('stepsize', 0.001), ('num_iters', 1), ('lowest_code_val', 0.001),
      ('normalize_dictionary', True)]
  assert params(conv_inf.run) == [
      ('images_padded', R), ('dictionary', R), ('kernel_stride', R), ('padding_dims', R), ('sparsity_weight', R),
      ('num_iters', R), ('variant', 'fista'), ('initial_codes', None), ('early_stopping_epsilon', None),
      ('nonnegative_only', False), ('hard_threshold', False)]
  assert params(conv_cheap.run) == [
      ('images_padded', R), ('dictionary', R), ('codes', R), ('hessian_diagonal', R), ('kernel_stride', R),
      ('padding_dims', R), ('stepsize', 0.001), ('num_iters', 1), ('lowest_code_val', 0.001),
      ('normalize_dictionary', True)]
  assert params(conv_steep.run) == [
      ('images_padded', R), ('dictionary', R), ('codes', R), ('kernel_stride', R), ('padding_dims', R),
      ('stepsize', 0.001), ('num_iters', 1), ('normalize_dictionary', True)]


def test_convolution_helpers_match_the_reference_rules():
  from vision_transform_codes_b200.utils import convolutions
  for image_dim, kernel_dim, stride in ((512, 16, 8), (21, 8, 4), (30, 12, 6), (40, 8, 8), (17, 6, 2)):
    assert convolutions.get_padding_amt(image_dim, kernel_dim, stride) == \
        oracle.get_padding_amt(image_dim, kernel_dim, stride)
    lead, trail = convolutions.get_padding_amt(image_dim, kernel_dim, stride)
    padded = image_dim + lead + trail
    assert (padded - kernel_dim) % stride == 0
    assert convolutions.code_dim_from_padded_img_dim(padded, kernel_dim, stride) == \
        oracle.code_dim_from_padded_img_dim(padded, kernel_dim, stride) == (padded - kernel_dim) // stride + 1
  x = torch.randn(2, 1, 12, 14)
  for pad in (((2, 3), (1, 4)), ((2, 0), (1, 1)), None):
    assert torch.equal(convolutions.create_mask(x, pad), oracle.create_mask(x, pad))


def test_conv_geometry_and_padding_rules():
  from vision_transform_codes_b200.analysis_transforms.convolutional.ista_fista import geometry
  x = torch.zeros(3, 1, 528, 528)
  phi = torch.zeros(64, 1, 16, 16)
  geo = geometry(x, phi, (8, 8), ((8, 8), (8, 8)))
  assert geo == (3, 1, 528, 528, 64, 16, 16, 8, 8, 8, 8, 8, 8, 65, 65)
  # no padding information: nothing masked
  assert geometry(x, phi, (8, 8), None)[9:13] == (0, 0, 0, 0)
  # a trailing padding of 0 zeroes the reference's whole mask (create_mask's ``-0:`` slice): an empty un-masked region
  assert geometry(x, phi, (8, 8), ((8, 0), (8, 8)))[9:13] == (528, 0, 528, 0)
  assert float(oracle.create_mask(x, ((8, 0), (8, 8))).abs().max()) == 0.0
  with pytest.raises(RuntimeError):
    geometry(torch.zeros(1, 1, 530, 528), phi, (8, 8), None)      # not kernel + whole strides
  with pytest.raises(ValueError):
    geometry(torch.zeros(1, 2, 528, 528), phi, (8, 8), None)      # channel mismatch


def test_conv_path_refuses_cpu_tensors_and_bad_variants():
  from vision_transform_codes_b200.analysis_transforms.convolutional import ista_fista as conv_inf
  from vision_transform_codes_b200.dict_update_rules.convolutional import sc_cheap_quadratic_descent as conv_cheap
  x, phi = torch.zeros(1, 1, 32, 32), torch.zeros(8, 1, 16, 16)
  pad = ((8, 8), (8, 8))
  with pytest.raises(RuntimeError):
    conv_inf.run(x, phi, (8, 8), pad, 0.1, 3)
  with pytest.raises(RuntimeError):
    conv_cheap.run(x, phi, torch.zeros(1, 8, 3, 3), torch.zeros(8), (8, 8), pad)
  with pytest.raises(AssertionError):
    conv_inf.run(x, phi, (8, 8), pad, 0.1, 3, variant='lista')


def test_conv_workspace_queries_need_no_gpu():
  from vision_transform_codes_b200 import _lib
  if not os.path.exists(_lib.LIB_PATH):
    pytest.skip('library not built yet')
  lib = _lib.load()
  small = lib.vtc_fista_conv_workspace_bytes(2, 1, 80, 80, 64, 16, 16, 8, 8, 3)
  big = lib.vtc_fista_conv_workspace_bytes(128, 1, 528, 528, 64, 16, 16, 8, 8, 3)
  assert 0 < small < big < (16 << 30)
  assert lib.vtc_fista_conv_workspace_bytes(2, 1, 80, 80, 64, 16, 16, 5, 5, 3) == 0   # kernel not a multiple of the stride
  assert lib.vtc_fista_conv_workspace_bytes(2, 1, 81, 80, 64, 16, 16, 8, 8, 3) == 0   # image not kernel + whole strides
  assert lib.vtc_conv_dict_grad_workspace_bytes(128, 1, 528, 528, 64, 16, 16, 8, 8, 6) > 0


def test_trainer_parameter_checks():
  """The lean trainer keeps the reference's parameter dictionary and its refusals (training/sparse_coding.py:289-330,
  :394-436) -- checked before anything touches a device."""
  from vision_transform_codes_b200.lean import sparse_coding as trainer
  phi = oracle.synthetic_conv_dictionary(4, 1, 8, 8)
  base = {'mode': 'convolutional', 'num_epochs': 1, 'strides': (4, 4), 'padding': ((4, 4), (4, 4)),
          'code_inference_algorithm': 'ista', 'inference_param_schedule': {0: {'sparsity_weight': 0.1, 'num_iters': 2}},
          'dictionary_update_algorithm': 'sc_cheap_quadratic_descent',
          'dict_update_param_schedule': {0: {'stepsize': 0.1, 'num_iters': 1}}}
  with pytest.raises(KeyError):
    trainer.train_dictionary([], None, phi, dict(base, code_inference_algorithm='subspace_ista'))
  with pytest.raises(KeyError):
    trainer.train_dictionary([], None, phi, dict(base, dictionary_update_algorithm='subspace_sc_cheap_quadratic_descent',
                                                 group_assignments=[[0, 1], [2, 3]], subspace_alignment_penalty=0.0))
  with pytest.raises(NotImplementedError):
    trainer.train_dictionary([], None, phi, dict(base, dict_element_rp_schedule={0: {}}))
  with pytest.raises(AssertionError):   # a visualisation schedule needs a pathlib.Path to log into (:330-343)
    trainer.train_dictionary([], None, phi, dict(base, training_visualization_schedule={0}))
  with pytest.raises(AssertionError):
    trainer.train_dictionary([], None, 2 * phi, base)   # dictionary not normalised
  assert trainer.train_dictionary([], None, phi, base) is not None   # an empty dataset: nothing to do, returns h


def test_metrics_from_totals_follow_the_reference_definitions():
  """Host side of training/metrics.py: the five scalars from the eight device totals (training/sparse_coding.py:196-225,
  utils/plotting.py:35-39), against the oracle's compute_metrics on a small batch whose totals are formed here."""
  from vision_transform_codes_b200.lean import metrics
  phi = oracle.synthetic_dictionary(24, 16)
  x = oracle.synthetic_patches(10, 16)
  codes = oracle.ista_fista(x, phi, 0.1, 20)
  x[3] = torch.mm(codes, phi)[3]   # an exactly reconstructed patch: pSNR is inf and the reference leaves it out (:223)
  want = oracle.compute_metrics(x, codes, phi, phi, 0.1)
  r2 = ((torch.mm(codes, phi) - x)**2).sum(1).double()
  mse = (r2 / 16).float()
  keep = mse != 0
  assert int(keep.sum()) == 9
  totals = [float(0.5 * r2.sum()), float(codes.abs().sum()), float((codes != 0).sum(1).double().div(24).sum()),
            float(torch.log10(mse[keep].double()).sum()), float(keep.sum()), float(x.min()), float(x.max()), 10.0]
  got = metrics.metrics_from_totals(totals, 0.1)
  for name in got:
    assert abs(got[name] - float(want[name])) <= 1e-5 * max(1.0, abs(float(want[name]))), name
  assert np.isnan(metrics.metrics_from_totals([0, 0, 0, 0, 0, 0.0, 1.0, 4], 0.1)[metrics.PSNR])
  avg = metrics.average_metrics([{'a': 1.0, 'b': np.array([1.0, 3.0])}, {'a': 3.0, 'b': np.array([5.0, 7.0])}])
  assert avg == {'a': 2.0, 'b': 4.0}


def test_tools_and_entry_points_compile():
  """The measurement / debugging scripts under tools/ and the root entry points at least parse (they only run on a GPU)."""
  import glob
  import py_compile
  from conftest import ROOT
  files = sorted(glob.glob(os.path.join(ROOT, 'tools', '*.py'))) + [os.path.join(ROOT, f) for f in
                                                                   ('bench.py', '__graft_entry__.py')]
  assert len(files) > 10
  for f in files:
    py_compile.compile(f, doraise=True)


def test_dataset_pickle_reader_takes_the_references_format(tmp_path):
  """{'training': {'patches': ndarray, ...}, 'validation': {...}} (tests/dset_generation_1.py:13-25)."""
  import pickle
  import numpy as np
  from vision_transform_codes_b200.utils import dataset_generation as dg
  rng = np.random.RandomState(0)
  raw = {'training': {'patches': rng.randn(103, 64).astype('float32'), 'local_contrasts': rng.randn(103, 1)},
         'validation': {'patches': rng.randn(40, 64).astype('float32')}}
  path = tmp_path / 'field_white_8x8.p'
  pickle.dump(raw, open(path, 'wb'))
  got = dg.load_patch_dataset(path, 'cpu')
  assert torch.equal(got['training'], torch.from_numpy(raw['training']['patches']))
  assert torch.equal(got['validation'], torch.from_numpy(raw['validation']['patches']))
  assert set(got['extras']['training']) == {'local_contrasts'}
  batched = dg.load_patch_dataset(path, 'cpu', batch_size=25, shuffle=True)
  seen = torch.cat(list(batched['training']))
  assert len(batched['training']) == 5 and seen.shape == (103, 64)
  assert torch.equal(seen.sort(dim=0).values, got['training'].sort(dim=0).values)   # a permutation of the patches
  assert [b.size(0) for b in dg.DeviceBatches(got['training'], 25, drop_last=True)] == [25] * 4
  assert [b.size(0) for b in batched['validation']] == [40]
  # unflattened (padded) image patches keep their (N, c, h, w) shape (tests/dset_generation_1.py:44-63)
  raw4 = {'training': {'patches': rng.randn(6, 1, 24, 24).astype('float32')}}
  pickle.dump(raw4, open(path, 'wb'))
  assert dg.load_patch_dataset(path, 'cpu')['training'].shape == (6, 1, 24, 24)
  pickle.dump({'nope': 1}, open(path, 'wb'))
  with pytest.raises(ValueError):
    dg.load_patch_dataset(path, 'cpu')

"""The UNMODIFIED reference trainer (training/sparse_coding.py:9-519, staged byte for byte in oracle/_ref) running on
the CUDA drop-ins, exactly as INTEGRATION.md section 1 describes: install() puts this repo's modules first, the
reference root follows, train_dictionary resolves its algorithms by dotted name (:389-439) and calls them with its own
keyword arguments (:126-139, :144-168). Compared with the dictionaries the same trainer produced on the reference's
own CPU implementation (tests/golden/training_small.npz, conv_training_small.npz; tests/golden/make_golden.py)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import reference
from oracle import vtc_oracle as oracle

pytestmark = pytest.mark.gpu

TOL = 1e-4


def fc_params(inference, update, **extra):
  p = {'mode': 'fully-connected', 'num_epochs': 1, 'code_inference_algorithm': inference,
       'inference_param_schedule': {0: {'sparsity_weight': 0.1, 'num_iters': 30}},
       'dictionary_update_algorithm': update,
       'dict_update_param_schedule': {0: {'stepsize': 0.1, 'num_iters': 1}}}
  p.update(extra)
  return p


def assert_resolution(trainer, expected_ours):
  import vision_transform_codes_b200 as pkg
  assert os.path.realpath(trainer.__file__).startswith(os.path.realpath(reference.root())), trainer.__file__
  for name in expected_ours:
    mod = sys.modules[name]   # imported by the reference trainer itself
    assert os.path.realpath(mod.__file__).startswith(os.path.realpath(pkg.PACKAGE_ROOT)), (name, mod.__file__)


def test_reference_trainer_fully_connected_on_the_cuda_drop_ins():
  from vision_transform_codes_b200 import _lib
  assert reference.available(), 'oracle/_ref is not staged (python tools/stage_reference.py)'
  g = load_golden('training_small')
  batches, phi0 = g['batches'].cuda(), g['dictionary']
  s = phi0.size(0)
  pairs = [list(map(int, v)) for v in np.array_split(np.arange(s), s // 2)]
  cases = [
      ('fista_cheap', fc_params('fista', 'sc_cheap_quadratic_descent'),
       ['analysis_transforms.fully_connected.ista_fista', 'dict_update_rules.fully_connected.sc_cheap_quadratic_descent']),
      ('ista_steepest', fc_params('ista', 'sc_steepest_descent'),
       ['analysis_transforms.fully_connected.ista_fista', 'dict_update_rules.fully_connected.sc_steepest_descent']),
      ('subspace_cheap', fc_params('subspace_fista', 'subspace_sc_cheap_quadratic_descent', group_assignments=pairs,
                                   subspace_alignment_penalty=0.0),
       ['analysis_transforms.fully_connected.subspace_ista_fista',
        'dict_update_rules.fully_connected.subspace_sc_cheap_quadratic_descent']),
      ('subspace_cheap_aligned', fc_params('subspace_fista', 'subspace_sc_cheap_quadratic_descent',
                                           group_assignments=pairs, subspace_alignment_penalty=0.3),
       ['analysis_transforms.fully_connected.subspace_ista_fista',
        'dict_update_rules.fully_connected.subspace_sc_cheap_quadratic_descent']),
  ]
  lib = _lib.load()
  for key, params, ours in cases:
    with reference.reference_on_drop_ins():
      trainer = reference.load('training.sparse_coding')
      phi = phi0.cuda()
      launches = lib.vtc_launch_count()
      trainer.train_dictionary(batches, batches[:1], phi, params)
      assert lib.vtc_launch_count() > launches, 'no kernel of libvtc_b200.so was launched'
      assert_resolution(trainer, ours)
    err = oracle.relative_l2(phi.cpu(), g[key])
    assert err < TOL, (key, err)


def test_reference_trainer_selects_the_subspace_steepest_rule_the_reference_lacks():
  """training/sparse_coding.py:421-427 imports dict_update_rules.fully_connected.subspace_sc_steepest_descent, which the
  reference tree does not contain (tests/sparse_coding_5.py:43 selects it): this repo provides it. Oracle = the
  steepest rule with the alignment term (oracle.sc_dictionary_update(h=None, alignment_penalty))."""
  g = load_golden('training_small')
  batches, phi0 = g['batches'], g['dictionary']
  s = phi0.size(0)
  pairs = [list(map(int, v)) for v in np.array_split(np.arange(s), s // 2)]
  params = fc_params('subspace_fista', 'subspace_sc_steepest_descent', group_assignments=pairs,
                     subspace_alignment_penalty=0.3)
  with reference.reference_on_drop_ins():
    trainer = reference.load('training.sparse_coding')
    phi = phi0.cuda()
    trainer.train_dictionary(batches.cuda(), batches[:1].cuda(), phi, params)
    assert_resolution(trainer, ['dict_update_rules.fully_connected.subspace_sc_steepest_descent'])
  want, _, _ = oracle.train_steps(batches, phi0, 0.1, 30, 0.1, variant='fista',
                                  update_rule='subspace_sc_steepest_descent', group_assignments=pairs,
                                  alignment_penalty=0.3)
  err = oracle.relative_l2(phi.cpu(), want)
  assert err < TOL, err


def test_reference_trainer_convolutional_on_the_cuda_drop_ins():
  g = load_golden('conv_training_small')
  xb, phi0 = g['batches'].cuda(), g['dictionary']
  pad = tuple(tuple(int(v) for v in p) for p in g['padding'].tolist())
  params = {'mode': 'convolutional', 'num_epochs': 1, 'strides': (8, 8), 'padding': pad,
            'code_inference_algorithm': 'ista',
            'inference_param_schedule': {0: {'sparsity_weight': 0.05, 'num_iters': 15}},
            'dictionary_update_algorithm': 'sc_cheap_quadratic_descent',
            'dict_update_param_schedule': {0: {'stepsize': 0.05, 'num_iters': 1}}}
  for key, p, ours in [
      ('ista_cheap', params, ['analysis_transforms.convolutional.ista_fista',
                              'dict_update_rules.convolutional.sc_cheap_quadratic_descent']),
      ('fista_steepest', dict(params, code_inference_algorithm='fista',
                              dictionary_update_algorithm='sc_steepest_descent'),
       ['analysis_transforms.convolutional.ista_fista', 'dict_update_rules.convolutional.sc_steepest_descent'])]:
    with reference.reference_on_drop_ins():
      trainer = reference.load('training.sparse_coding')
      d = phi0.cuda()
      trainer.train_dictionary(xb, xb[:1], d, p)
      assert_resolution(trainer, ours)
    err = oracle.relative_l2(d.cpu(), g[key])
    assert err < TOL, (key, err)


def test_reference_trainer_checkpoints_through_its_own_code(tmp_path):
  """checkpoint_schedule is the reference's own code path (training/sparse_coding.py:170-175, yaml dump :367-384)."""
  import pickle
  g = load_golden('training_small')
  with reference.reference_on_drop_ins():
    trainer = reference.load('training.sparse_coding')
    phi = g['dictionary'].cuda()
    trainer.train_dictionary(g['batches'].cuda(), None, phi,
                             fc_params('fista', 'sc_cheap_quadratic_descent', checkpoint_schedule={0, 3},
                                       logging_folder_fullpath=tmp_path / 'log'))
  first = pickle.load(open(tmp_path / 'log' / 'checkpoint_dictionary_iter_0', 'rb'))
  assert np.array_equal(first, g['dictionary'].numpy())
  assert (tmp_path / 'log' / 'training_params.yaml').exists()
  assert oracle.relative_l2(phi.cpu(), g['fista_cheap']) < TOL

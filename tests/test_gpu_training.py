"""Train steps on the GPU against the reference trainer's committed outputs, through both boundaries:
the lean trainer of this package and the reference-style by-name import after install()."""
import importlib
import sys

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import vtc_oracle as oracle

pytestmark = pytest.mark.gpu


def params(inference, update, **extra):
  p = {'mode': 'fully-connected', 'num_epochs': 1, 'code_inference_algorithm': inference,
       'inference_param_schedule': {0: {'sparsity_weight': 0.1, 'num_iters': 30}},
       'dictionary_update_algorithm': update,
       'dict_update_param_schedule': {0: {'stepsize': 0.1, 'num_iters': 1}}}
  p.update(extra)
  return p


def test_trainer_matches_reference_trainer_outputs(capsys):
  from vision_transform_codes_b200.lean import sparse_coding as trainer
  g = load_golden('training_small')
  batches, phi0 = g['batches'].cuda(), g['dictionary']
  s = phi0.size(0)
  phi = phi0.cuda()
  trainer.train_dictionary(batches, batches[:1], phi, params('fista', 'sc_cheap_quadratic_descent'))
  assert oracle.relative_l2(phi.cpu(), g['fista_cheap']) < 1e-4
  phi = phi0.cuda()
  trainer.train_dictionary(batches, batches[:1], phi, params('ista', 'sc_steepest_descent'))
  assert oracle.relative_l2(phi.cpu(), g['ista_steepest']) < 1e-4
  pairs = [list(map(int, v)) for v in np.array_split(np.arange(s), s // 2)]
  phi = phi0.cuda()
  trainer.train_dictionary(batches, batches[:1], phi,
                           params('subspace_fista', 'subspace_sc_cheap_quadratic_descent', group_assignments=pairs,
                                  subspace_alignment_penalty=0.0))
  assert oracle.relative_l2(phi.cpu(), g['subspace_cheap']) < 1e-4
  phi = phi0.cuda()
  trainer.train_dictionary(batches, batches[:1], phi,
                           params('subspace_fista', 'subspace_sc_cheap_quadratic_descent', group_assignments=pairs,
                                  subspace_alignment_penalty=0.3))
  assert oracle.relative_l2(phi.cpu(), g['subspace_cheap_aligned']) < 1e-4


def test_lean_trainer_updates_a_non_contiguous_dictionary_correctly():
  """a strided init_dictionary (a transposed view) must be updated through a contiguous working copy, not
  read and written as if it were dense -- same result, in place in the caller's storage."""
  from vision_transform_codes_b200.lean import sparse_coding as trainer
  g = load_golden('training_small')
  batches, phi0 = g['batches'].cuda(), g['dictionary']
  storage = phi0.t().contiguous().cuda()        # (n, s) storage ...
  phi_view = storage.t()                        # ... seen as a (s, n) dictionary with strides (1, s)
  assert not phi_view.is_contiguous()
  trainer.train_dictionary(batches, batches[:1], phi_view, params('fista', 'sc_cheap_quadratic_descent'))
  assert oracle.relative_l2(phi_view.cpu(), g['fista_cheap']) < 1e-4
  assert oracle.relative_l2(storage.t().cpu(), g['fista_cheap']) < 1e-4   # the caller's own storage was updated
  with pytest.raises(RuntimeError):   # no CPU fallback: a host dictionary is refused before anything runs
    trainer.train_dictionary(batches.cpu(), None, phi0.clone(), params('fista', 'sc_cheap_quadratic_descent'))


def test_checkpoint_format_matches_reference(tmp_path):
  import pickle
  from vision_transform_codes_b200.lean import sparse_coding as trainer
  g = load_golden('training_small')
  phi = g['dictionary'].cuda()
  trainer.train_dictionary(g['batches'].cuda(), None, phi,
                           params('fista', 'sc_cheap_quadratic_descent', checkpoint_schedule={0, 2},
                                  logging_folder_fullpath=tmp_path))
  first = pickle.load(open(tmp_path / 'checkpoint_dictionary_iter_0', 'rb'))
  assert isinstance(first, np.ndarray) and first.dtype == np.float32
  assert np.array_equal(first, g['dictionary'].numpy())
  assert (tmp_path / 'checkpoint_dictionary_iter_2').exists()


def test_by_name_imports_drive_the_cuda_path():
  """What the unmodified reference trainer does at training/sparse_coding.py:389-439: import by dotted name."""
  import vision_transform_codes_b200 as pkg
  saved = list(sys.path)
  try:
    pkg.install()
    inference_alg = importlib.import_module('analysis_transforms.fully_connected.ista_fista')
    dict_update = importlib.import_module('dict_update_rules.fully_connected.sc_cheap_quadratic_descent')
    g = load_golden('training_small')
    x = g['batches'][0].cuda()
    phi = g['dictionary'].cuda()
    h = torch.zeros(phi.size(0), device='cuda')
    # the kwargs the reference trainer passes (:126-139, :144-168)
    codes = inference_alg.run(dictionary=phi, sparsity_weight=0.1, num_iters=30, variant='fista',
                              nonnegative_only=False, hard_threshold=False, images=x)
    h.mul_(0.99).add_(torch.pow(codes, 2).mean(0) / 100)
    dict_update.run(dictionary=phi, codes=codes, stepsize=0.1, num_iters=1, images=x, hessian_diagonal=h)
    want_phi, _, _ = oracle.train_steps(g['batches'][:1], g['dictionary'], 0.1, 30, 0.1)
    assert oracle.relative_l2(phi.cpu(), want_phi) < 1e-4
  finally:
    sys.path[:] = saved
    for name in list(sys.modules):
      if name.split('.')[0] in ('analysis_transforms', 'dict_update_rules'):
        del sys.modules[name]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_two_gpu_train_step_equals_one_gpu(tmp_path):
  import os
  import subprocess
  from conftest import ROOT
  out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                        '--master-addr', '127.0.0.1', '--master-port', '29611',
                        os.path.join(ROOT, 'tools', 'dp_equivalence.py')], capture_output=True, text=True, timeout=600)
  assert out.returncode == 0, out.stdout + out.stderr
  assert 'DP_EQUIVALENCE_OK' in out.stdout


@pytest.mark.gpu
def test_dictionary_gradient_writes_only_its_output():
  """vtc_sc_dict_grad writing into the middle of a larger buffer: the guard zones keep their canaries, the pitched
  images and codes are not written, and the result equals the dense call (ragged shapes, both update precisions)."""
  import ctypes
  import vision_transform_codes_b200 as pkg
  from vision_transform_codes_b200 import _lib
  from vision_transform_codes_b200.dict_update_rules.fully_connected import _common
  lib = _lib.load()
  dev = torch.device('cuda:0')
  canary = 54321.0
  saved = pkg.config.update_precision
  try:
    for precision in ('bf16x3', 'bf16x6'):
      pkg.config.update_precision = precision
      prec = pkg.config.precision_code('update_precision')
      for (b, s, d) in ((1000, 300, 100), (4096, 1024, 256), (77, 64, 20)):
        x = oracle.synthetic_patches(b, d, seed=3).to(dev)
        phi = oracle.synthetic_dictionary(s, d).to(dev)
        codes = torch.relu(torch.randn(b, s, generator=torch.Generator().manual_seed(2)) - 1.0).to(dev)
        want = _common.dictionary_gradient(x, phi, codes)
        ldx, ldc, guard = d + 4, s + 8, 2048
        xb = torch.full((b, ldx), canary, device=dev)
        xb[:, :d] = x
        cb = torch.full((b, ldc), canary, device=dev)
        cb[:, :s] = codes
        x0, c0 = xb.clone(), cb.clone()
        buf = torch.full((guard + s * d + guard,), canary, device=dev)
        out = buf[guard:guard + s * d].view(s, d)
        with torch.cuda.device(dev):
          nbytes = lib.vtc_dict_grad_workspace_bytes(b, s, d, prec)
          ws = _lib.workspace(nbytes, dev, 'dict_grad_guard_test')
          _lib.check(lib.vtc_sc_dict_grad(_lib.ptr(xb), ldx, _lib.ptr(phi), _lib.ptr(cb), ldc, _lib.ptr(out), b, s, d, prec,
                                          _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)))
        torch.cuda.synchronize()
        assert torch.equal(out, want), (precision, b, s, d)
        assert bool((buf[:guard] == canary).all()) and bool((buf[guard + s * d:] == canary).all()), 'guard zone written'
        assert torch.equal(xb, x0) and torch.equal(cb, c0)
  finally:
    pkg.config.update_precision = saved

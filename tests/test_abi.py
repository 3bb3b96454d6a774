"""The C-ABI library loads and exports every symbol include/vtc_b200.h declares (no compute: runs without a GPU)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def declared_symbols():
  text = open(os.path.join(ROOT, 'include', 'vtc_b200.h')).read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(vtc_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_the_expected_entry_points():
  names = declared_symbols()
  for must in ('vtc_fista_fc', 'vtc_sc_dict_grad', 'vtc_sc_dict_apply', 'vtc_hessian_diag_update', 'vtc_lipschitz',
               'vtc_fista_conv', 'vtc_sc_conv_dict_grad', 'vtc_sc_conv_dict_apply', 'vtc_conv_hessian_diag_update'):
    assert must in names


def test_library_exports_every_declared_symbol():
  from vision_transform_codes_b200 import _lib
  if not os.path.exists(_lib.LIB_PATH):
    pytest.skip('library not built yet: run __graft_entry__.build()')
  lib = ctypes.CDLL(_lib.LIB_PATH)
  for name in declared_symbols():
    assert hasattr(lib, name), name
  assert set(declared_symbols()) == set(_lib.SIGNATURES), 'ctypes table and header disagree'
  assert _lib.load().vtc_version() >= 100


def test_workspace_queries_need_no_gpu():
  from vision_transform_codes_b200 import _lib
  if not os.path.exists(_lib.LIB_PATH):
    pytest.skip('library not built yet')
  lib = _lib.load()
  small = lib.vtc_fista_workspace_bytes(250, 256, 256, 3)
  big = lib.vtc_fista_workspace_bytes(65536, 1024, 256, 3)
  assert 0 < small < big < (8 << 30)
  assert lib.vtc_fista_workspace_bytes(250, 256, 256, 2) == 0  # invalid precision
  assert lib.vtc_dict_grad_workspace_bytes(65536, 1024, 256, 6) > 0


def test_no_cpu_fallback():
  """The product path must refuse CPU tensors instead of silently computing somewhere else."""
  from vision_transform_codes_b200.analysis_transforms.fully_connected import ista_fista
  from vision_transform_codes_b200.dict_update_rules.fully_connected import sc_cheap_quadratic_descent
  x, phi = torch.zeros(4, 8), torch.eye(8)
  with pytest.raises(RuntimeError):
    ista_fista.run(x, phi, 0.1, 3)
  with pytest.raises(RuntimeError):
    sc_cheap_quadratic_descent.run(x, phi, torch.zeros(4, 8), torch.zeros(8))
  with pytest.raises(AssertionError):
    ista_fista.run(x, phi, 0.1, 3, variant='lista')


def test_product_code_never_imports_the_oracle():
  pkg = os.path.join(ROOT, 'vision_transform_codes_b200')
  for dirpath, _, files in os.walk(pkg):
    for f in files:
      if f.endswith(('.py', '.cu', '.cuh', '.h')):
        assert 'oracle' not in open(os.path.join(dirpath, f)).read().lower().replace('# oracle-free', ''), f


def test_dropin_names_resolve_through_install():
  import importlib
  import sys
  import vision_transform_codes_b200 as pkg
  saved = list(sys.path)
  try:
    pkg.install()
    for name in ('analysis_transforms.fully_connected.ista_fista',
                 'analysis_transforms.fully_connected.subspace_ista_fista',
                 'dict_update_rules.fully_connected.sc_cheap_quadratic_descent',
                 'dict_update_rules.fully_connected.sc_steepest_descent',
                 'dict_update_rules.fully_connected.subspace_sc_cheap_quadratic_descent',
                 'analysis_transforms.convolutional.ista_fista',
                 'dict_update_rules.convolutional.sc_cheap_quadratic_descent',
                 'dict_update_rules.convolutional.sc_steepest_descent'):
      mod = importlib.import_module(name)
      assert callable(mod.run)
      assert mod.__file__.startswith(pkg.PACKAGE_ROOT)
  finally:
    sys.path[:] = saved
    for name in list(sys.modules):
      if name.split('.')[0] in ('analysis_transforms', 'dict_update_rules'):
        del sys.modules[name]


def test_shipped_library_has_no_trace_points():
  """The timeline hook of tools/iter_trace.py exists only in a -DVTC_TRACE build: the shipped kernels carry no trace
  points, and asking the shipped library for a trace fails loudly instead of silently recording nothing."""
  from vision_transform_codes_b200 import _lib
  if not os.path.exists(_lib.LIB_PATH):
    pytest.skip('library not built yet: run __graft_entry__.build()')
  lib = _lib.load()
  rc = lib.vtc_debug_iter_trace(None)
  assert rc == _lib.VTC_ERR_ARG
  assert b'VTC_TRACE' in lib.vtc_last_error()
  with pytest.raises(ValueError):
    _lib.check(rc)

"""
Subspace (group-LASSO) ISTA / FISTA for fully-connected sparse inference, on one B200.

Drop-in for vision_transform_codes/analysis_transforms/fully_connected/subspace_ista_fista.py:23-192. The
thresholding is applied to the L2 norm of each group of coefficients. As in the reference the codes are regrouped
into a padded (b, num_groups, group_width) layout and the dictionary into the matching "grouped dictionary"
(:94-111); here that is a row gather on the device, after which the same fused GEMM kernel as vanilla FISTA runs
with a group-shrinkage epilogue, and the result is scatter-added back (:184-190).
"""
import os
import sys

import numpy as np
import torch

try:
  from vision_transform_codes_b200 import _lib, config
except ImportError:
  sys.path.append(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
  from vision_transform_codes_b200 import _lib, config
from vision_transform_codes_b200.analysis_transforms.fully_connected import ista_fista as _vanilla


def _slot_table(group_assignments, num_atoms):
  """Atom index of every (group, position) slot, -1 for padding; width = max group size rounded up to 2^k."""
  sizes = [len(g) for g in group_assignments]
  max_size = max(sizes)
  width = 1
  while width < max_size:
    width *= 2
  table = np.full((len(group_assignments), width), -1, dtype=np.int32)
  for g_idx, g in enumerate(group_assignments):
    g = np.asarray(g, dtype=np.int64).reshape(-1)
    if g.size and (g.min() < 0 or g.max() >= num_atoms):
      raise IndexError('group %d refers to a dictionary element outside [0, %d)' % (g_idx, num_atoms))
    table[g_idx, :g.size] = g
  return table, max_size, width


def run(images, dictionary, group_assignments, sparsity_weight,
        num_iters, variant='fista', ret_summed_gduplicates=True,
        initial_codes=None, early_stopping_epsilon=None, hard_threshold=False):
  """
  Runs steps of subspace Iterative Shrinkage/Thresholding.

  Parameters
  ----------
  images : torch.Tensor(float32, size=(b, n))
  dictionary : torch.Tensor(float32, size=(s, n))
  group_assignments : list(array_like)
      group_assignments[g] lists the dictionary elements of group g; groups may differ in size and an element may
      belong to several groups (its duplicated code values are summed in the result).
  sparsity_weight : float
  num_iters : int
  variant : str, optional
      One of {'ista', 'fista'}. Default 'fista'.
  ret_summed_gduplicates : bool, optional
      Only True is implemented (as in the reference).
  initial_codes : torch.Tensor(float32, size=(b, s)), optional
  early_stopping_epsilon : float, optional
  hard_threshold : bool, optional
      Not implemented (as in the reference).

  Returns
  -------
  codes : torch.Tensor(float32, size=(b, s))
  """
  assert variant in ['ista', 'fista']
  if hard_threshold:
    raise NotImplementedError('TODO')
  if not ret_summed_gduplicates:
    raise NotImplementedError('TODO')
  _lib.require_cuda_f32(images, 'images')
  _lib.require_cuda_f32(dictionary, 'dictionary')
  S = dictionary.size(0)
  # groups of up to 16 elements shrink inside the GEMM epilogue; wider ones (padded to 32, 64, ...) take a separate
  # pass per iteration (vtc_fista_fc, wide_group_prox_kernel)
  table, max_size, width = _slot_table(group_assignments, S)
  n_slots = table.size
  identity = (n_slots == S and np.array_equal(table.reshape(-1), np.arange(S, dtype=np.int32)))
  # the reference averages |delta|/stepsize over b * num_groups * max_size slots (:172-177); padding slots stay 0
  eps_scale = float(max_size) / float(width)
  if identity:
    # in-order partition into equal groups (every use in the reference): grouped dictionary == dictionary
    codes, _ = _vanilla.infer(images, dictionary, sparsity_weight, num_iters, variant, initial_codes,
                              early_stopping_epsilon, False, False, group_size=width, eps_scale=eps_scale)
    return codes

  lib = _lib.load()
  device = images.device
  B, D = images.shape
  index = torch.from_numpy(table.reshape(-1)).to(device)
  dictionary_c = dictionary.contiguous()
  grouped_dictionary = torch.empty((n_slots, D), dtype=torch.float32, device=device)
  with torch.cuda.device(device):
    st = _lib.stream_ptr(device)
    _lib.check(lib.vtc_gather_rows(_lib.ptr(dictionary_c), D, _lib.ptr(index), n_slots, D,
                                   _lib.ptr(grouped_dictionary), st))
    grouped_init = None
    if initial_codes is not None:
      _lib.require_cuda_f32(initial_codes, 'initial_codes')
      init_rm, ld_init = _lib.row_major(initial_codes)
      grouped_init = torch.empty((B, n_slots), dtype=torch.float32, device=device)
      _lib.check(lib.vtc_gather_cols(_lib.ptr(init_rm), ld_init, _lib.ptr(index), B, n_slots,
                                     _lib.ptr(grouped_init), n_slots, st))
    grouped_codes, _ = _vanilla.infer(images, grouped_dictionary, sparsity_weight, num_iters, variant, grouped_init,
                                      early_stopping_epsilon, False, False, group_size=width, eps_scale=eps_scale)
    codes = torch.empty((B, S), dtype=torch.float32, device=device)
    _lib.check(lib.vtc_scatter_add_cols(_lib.ptr(grouped_codes), n_slots, _lib.ptr(index), B, n_slots,
                                        _lib.ptr(codes), S, S, st))
  return codes

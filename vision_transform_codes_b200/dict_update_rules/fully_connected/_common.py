"""Shared driver of the fully-connected sparse-coding dictionary updates (gradient contraction + apply step)."""
import os
import sys

import torch

try:
  from vision_transform_codes_b200 import _lib, config
except ImportError:
  sys.path.append(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
  from vision_transform_codes_b200 import _lib, config


def dictionary_gradient(images, dictionary, codes, out=None):
  """grad_sum = codes^T (codes @ dictionary - images), the un-normalised sum over this batch (S, D)."""
  lib = _lib.load()
  device = dictionary.device
  B, D = images.shape
  S = dictionary.size(0)
  images_rm, ld_images = _lib.row_major(images)
  codes_rm, ld_codes = _lib.row_major(codes)
  if out is None:
    out = torch.empty((S, D), dtype=torch.float32, device=device)
  prec = config.precision_code('update_precision')
  with torch.cuda.device(device):
    nbytes = lib.vtc_dict_grad_workspace_bytes(B, S, D, prec)
    ws = _lib.workspace(nbytes, device, 'dict_grad')
    _lib.check(lib.vtc_sc_dict_grad(_lib.ptr(images_rm), ld_images, _lib.ptr(dictionary), _lib.ptr(codes_rm),
                                    ld_codes, _lib.ptr(out), B, S, D, prec, _lib.ptr(ws), ws.numel(),
                                    _lib.stream_ptr(device)))
  return out


def global_batch_size(local_batch, device):
  """Number of samples behind one update: the local batch, or its sum over the data-parallel group."""
  if not config.data_parallel:
    return local_batch
  import torch.distributed as dist
  n = torch.tensor([local_batch], dtype=torch.int64, device=device)
  dist.all_reduce(n, op=dist.ReduceOp.SUM, group=config.process_group)
  return int(n.item())


def group_slot_table(group_assignments, num_atoms, device):
  """(num_groups, width) int32 device table: atom of every (group, position) slot, -1 for padding."""
  import numpy as np
  width = max(len(g) for g in group_assignments)
  table = np.full((len(group_assignments), width), -1, dtype=np.int32)
  for g_idx, g in enumerate(group_assignments):
    g = np.asarray(g, dtype=np.int64).reshape(-1)
    if g.size and (g.min() < 0 or g.max() >= num_atoms):
      raise IndexError('group %d refers to a dictionary element outside [0, %d)' % (g_idx, num_atoms))
    table[g_idx, :g.size] = g
  return torch.from_numpy(table).to(device), width


def descend(images, dictionary, codes, hessian_diagonal, stepsize, num_iters, lowest_code_val, normalize_dictionary,
            batch_global=None, group_assignments=None, alignment_penalty=0.0):
  """
  num_iters steps of  dictionary <- rownorm(dictionary - stepsize * (codes^T (codes dictionary - images) / b) / (h + eps))
  in place (sc_cheap_quadratic_descent.py:42-48; hessian_diagonal=None gives sc_steepest_descent.py:37-41).
  With data parallelism enabled the gradient sum is all-reduced and b is the global batch, so every replica applies
  the identical update.
  """
  for t, name in ((images, 'images'), (dictionary, 'dictionary'), (codes, 'codes')):
    _lib.require_cuda_f32(t, name)
  if hessian_diagonal is not None:
    _lib.require_cuda_f32(hessian_diagonal, 'hessian_diagonal')
    hessian_diagonal = hessian_diagonal.contiguous()
  if images.size(0) != codes.size(0) or codes.size(1) != dictionary.size(0) or images.size(1) != dictionary.size(1):
    raise ValueError('shape mismatch: images %s, dictionary %s, codes %s' %
                     (tuple(images.shape), tuple(dictionary.shape), tuple(codes.shape)))
  lib = _lib.load()
  device = dictionary.device
  S, D = dictionary.shape
  target = dictionary
  work = dictionary if dictionary.is_contiguous() else dictionary.contiguous()
  if batch_global is None:
    batch_global = global_batch_size(codes.size(0), device)
  reg, slots, width = None, None, 0
  if alignment_penalty != 0:
    # within-group alignment regulariser (subspace_sc_cheap_quadratic_descent.py:59-79): depends on the dictionary
    # only, so under data parallelism every replica computes the same term and nothing is exchanged for it
    slots, width = group_slot_table(group_assignments, S, device)
    reg = torch.empty((S, D), dtype=torch.float32, device=device)
  for _ in range(int(num_iters)):
    if reg is not None:
      with torch.cuda.device(device):
        _lib.check(lib.vtc_subspace_alignment_grad(_lib.ptr(work), S, D, _lib.ptr(slots), slots.size(0), width,
                                                   int(bool(normalize_dictionary)), _lib.ptr(reg),
                                                   _lib.stream_ptr(device)))
    grad = dictionary_gradient(images, work, codes)
    if config.data_parallel:
      import torch.distributed as dist
      dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=config.process_group)
    with torch.cuda.device(device):
      _lib.check(lib.vtc_sc_dict_apply(_lib.ptr(work), _lib.ptr(grad), _lib.ptr(hessian_diagonal), _lib.ptr(reg),
                                       float(alignment_penalty), S, D, int(batch_global), float(stepsize),
                                       float(lowest_code_val), int(bool(normalize_dictionary)),
                                       _lib.stream_ptr(device)))
  if work is not target:
    target.copy_(work)

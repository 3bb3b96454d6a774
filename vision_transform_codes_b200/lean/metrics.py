"""
Validation metrics of the sparse-coding trainer, computed on the device.

The reference's ``compute_metrics`` (training/sparse_coding.py:177-229) pulls the images, the codes' reconstructions
and per-sample norms to the host with ``.cpu().numpy()`` and loops over the batch for the pSNR. Here the
reconstruction error comes from one tcgen05 contraction (fully connected: codes * dictionary - images; convolutional:
the masked synthesis contraction of the inference path), the per-item statistics from one pass over residuals, codes
and pixels, and only eight doubles cross to the host. Same metric names and definitions as the reference.

Under data parallelism (``vision_transform_codes_b200.enable_data_parallel()``) every rank passes its shard of the
validation batch and the totals are all-reduced, so every rank reports the metrics of the global batch.
"""
import math

import numpy as np
import torch

from vision_transform_codes_b200 import _lib, config
from vision_transform_codes_b200.analysis_transforms.convolutional.ista_fista import geometry
from vision_transform_codes_b200.analysis_transforms.fully_connected.subspace_ista_fista import _slot_table

L2, LAGRANGE, LOSS, L0, PSNR, CHANGE = (
    'Average LASSO L2 component', 'Average LASSO lagrange component', 'Average LASSO Loss', 'Average Normalized L0',
    'Average pSNR of reconstructions', 'Average change in dictionary kernels')


def _allreduce_totals(totals):
  import torch.distributed as dist
  sums = totals[[0, 1, 2, 3, 4, 7]].contiguous()
  lo, hi = totals[5:6].clone(), totals[6:7].clone()
  dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=config.process_group)
  dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=config.process_group)
  dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=config.process_group)
  totals[[0, 1, 2, 3, 4, 7]] = sums
  totals[5:6], totals[6:7] = lo, hi


def batch_totals(batch_images, batch_codes, dictionary, group_assignments=None, kernel_strides=None,
                 image_padding=None):
  """The eight totals of include/vtc_b200.h:vtc_sc_metrics as a float64 tensor on the device (no host sync)."""
  for t, name in ((batch_images, 'batch_images'), (batch_codes, 'batch_codes'), (dictionary, 'dictionary')):
    _lib.require_cuda_f32(t, name)
  lib = _lib.load()
  device = dictionary.device
  totals = torch.empty(8, dtype=torch.float64, device=device)
  prec = config.precision_code('update_precision')
  with torch.cuda.device(device):
    st = _lib.stream_ptr(device)
    if kernel_strides is None:
      images, ld_images = _lib.row_major(batch_images)
      codes, ld_codes = _lib.row_major(batch_codes)
      B, D = images.shape
      S = dictionary.size(0)
      if tuple(dictionary.shape) != (S, D) or tuple(codes.shape) != (B, S):
        raise ValueError('shapes do not agree: images %s, dictionary %s, codes %s' % (
            tuple(images.shape), tuple(dictionary.shape), tuple(codes.shape)))
      slots, n_groups, width = None, 0, 0
      if group_assignments is not None:
        table, _, width = _slot_table(group_assignments, S)
        n_groups = table.shape[0]
        slots = torch.from_numpy(table).to(device)
      nbytes = max(16, lib.vtc_sc_metrics_workspace_bytes(B, S, D, prec))
      ws = _lib.workspace(nbytes, device, 'metrics')
      _lib.check(lib.vtc_sc_metrics(_lib.ptr(images), ld_images, _lib.ptr(dictionary.contiguous()), _lib.ptr(codes),
                                    ld_codes, B, S, D, _lib.ptr(slots), n_groups, width, prec, _lib.ptr(totals),
                                    _lib.ptr(ws), ws.numel(), st))
    else:
      if group_assignments is not None:
        raise KeyError('Havent implemented subspace ISTA for convolutional yet')
      B, C, H, W, S, KH, KW, SY, SX, pt, pb, pl, pr, SH, SW = geometry(batch_images, dictionary, kernel_strides,
                                                                        image_padding)
      if image_padding is not None and (image_padding[0][1] == 0 or image_padding[1][1] == 0):
        # recons[:, :, top:-0] is empty in the reference (:189-194) and its np.max(...) then raises
        raise ValueError('zero-size array to reduction operation maximum which has no identity '
                         '(a trailing padding of 0 crops everything, training/sparse_coding.py:189-194)')
      if tuple(batch_codes.shape) != (B, S, SH, SW):
        raise ValueError('codes must have shape %s, got %s' % ((B, S, SH, SW), tuple(batch_codes.shape)))
      nbytes = max(16, lib.vtc_sc_conv_metrics_workspace_bytes(B, C, H, W, S, KH, KW, SY, SX, prec))
      ws = _lib.workspace(nbytes, device, 'conv_metrics')
      _lib.check(lib.vtc_sc_conv_metrics(
          _lib.ptr(batch_images.contiguous()), _lib.ptr(dictionary.contiguous()), _lib.ptr(batch_codes.contiguous()),
          B, C, H, W, S, KH, KW, SY, SX, pt, pb, pl, pr, prec, _lib.ptr(totals), _lib.ptr(ws), ws.numel(), st))
  if config.data_parallel:
    _allreduce_totals(totals)
  return totals


def dictionary_change(dictionary, previous_dictionary):
  """Mean |dictionary - previous_dictionary| per dictionary element (training/sparse_coding.py:226-228), on the device."""
  _lib.require_cuda_f32(dictionary, 'dictionary')
  _lib.require_cuda_f32(previous_dictionary, 'previous_dictionary')
  if dictionary.shape != previous_dictionary.shape:
    raise ValueError('dictionary and previous_dictionary differ in shape')
  lib = _lib.load()
  device = dictionary.device
  S = dictionary.size(0)
  out = torch.empty(S, dtype=torch.float32, device=device)
  with torch.cuda.device(device):
    _lib.check(lib.vtc_dict_change(_lib.ptr(dictionary.contiguous()), _lib.ptr(previous_dictionary.contiguous()), S,
                                   dictionary[0].numel(), _lib.ptr(out), _lib.stream_ptr(device)))
  return out


def metrics_from_totals(totals, sparsity_weight):
  """Host side of compute_metrics: the five batch scalars from the eight totals (a sequence of Python floats)."""
  l2_sum, l1_sum, l0_sum, log_mse_sum, n_finite, lo, hi, n = [float(v) for v in totals]
  metrics = {L2: l2_sum / n, LAGRANGE: float(sparsity_weight) * l1_sum / n}
  metrics[LOSS] = metrics[L2] + metrics[LAGRANGE]
  metrics[L0] = l0_sum / n
  sig_mag = hi - lo
  if n_finite > 0 and sig_mag > 0:
    # mean_b 10 log10(sig^2 / mse_b) over the items with mse_b != 0 (utils/plotting.py:35-39, :223-225)
    metrics[PSNR] = 10. * (2. * math.log10(sig_mag) - log_mse_sum / n_finite)
  elif n_finite > 0:
    metrics[PSNR] = -math.inf
  else:
    metrics[PSNR] = math.nan  # np.mean([]) in the reference
  return metrics


def compute_metrics(batch_images, batch_codes, dictionary, previous_dictionary, sparsity_weight, code_inf_alg='fista',
                    group_assignments=None, kernel_strides=None, image_padding=None):
  """
  training/sparse_coding.py:177-229 for one validation batch: a dict with the reference's six metric names. The group
  norms replace the l1 norm for the subspace algorithms (:199-205). One device->host copy of eight doubles plus the
  (s,) vector of dictionary changes.
  """
  groups = group_assignments if code_inf_alg in ('subspace_ista', 'subspace_fista') else None
  totals = batch_totals(batch_images, batch_codes, dictionary, groups, kernel_strides, image_padding)
  change = dictionary_change(dictionary, previous_dictionary)
  metrics = metrics_from_totals(totals.cpu().tolist(), sparsity_weight)
  metrics[CHANGE] = change.cpu().numpy()
  return metrics


def average_metrics(per_batch):
  """training/sparse_coding.py:505-506: the mean over validation batches of every metric (np.mean, arrays included)."""
  return {name: np.mean([m[name] for m in per_batch]) for name in per_batch[0]}

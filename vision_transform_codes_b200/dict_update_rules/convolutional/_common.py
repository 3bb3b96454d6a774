"""Shared driver of the convolutional sparse-coding dictionary updates (gradient contractions + apply step)."""
import os
import sys

import torch

try:
  from vision_transform_codes_b200 import _lib, config
  from vision_transform_codes_b200.analysis_transforms.convolutional.ista_fista import align_to_stride, geometry
  from vision_transform_codes_b200.dict_update_rules.fully_connected._common import global_batch_size
except ImportError:
  sys.path.append(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
  from vision_transform_codes_b200 import _lib, config
  from vision_transform_codes_b200.analysis_transforms.convolutional.ista_fista import align_to_stride, geometry
  from vision_transform_codes_b200.dict_update_rules.fully_connected._common import global_batch_size


def dictionary_gradient(images_padded, dictionary, codes, kernel_stride, padding_dims, out=None):
  """Gradient of the reconstruction error w.r.t. the kernels, summed over this batch (not divided by it), (s, c, kh, kw)."""
  lib = _lib.load()
  device = dictionary.device
  geo = geometry(images_padded, dictionary, kernel_stride, padding_dims)
  kh, kw = dictionary.size(2), dictionary.size(3)
  # a kernel size that is not a multiple of the stride: zero taps up to the next multiple (align_to_stride); the
  # gradient of those taps is not part of the dictionary and is cropped away
  images_padded, dictionary, geo = align_to_stride(images_padded, dictionary, geo)
  B, C, H, W, S, KH, KW, SY, SX, pt, pb, pl, pr, SH, SW = geo
  aligned = (KH, KW) != (kh, kw)
  if tuple(codes.shape) != (B, S, SH, SW):
    raise ValueError('codes must have shape %s, got %s' % ((B, S, SH, SW), tuple(codes.shape)))
  full = out if (out is not None and not aligned) else torch.empty((S, C, KH, KW), dtype=torch.float32, device=device)
  prec = config.precision_code('update_precision')
  with torch.cuda.device(device):
    nbytes = max(16, lib.vtc_conv_dict_grad_workspace_bytes(B, C, H, W, S, KH, KW, SY, SX, prec))
    ws = _lib.workspace(nbytes, device, 'conv_dict_grad')
    _lib.check(lib.vtc_sc_conv_dict_grad(
        _lib.ptr(images_padded.contiguous()), _lib.ptr(dictionary.contiguous()), _lib.ptr(codes.contiguous()),
        _lib.ptr(full), B, C, H, W, S, KH, KW, SY, SX, pt, pb, pl, pr, prec, _lib.ptr(ws), ws.numel(),
        _lib.stream_ptr(device)))
  if not aligned:
    return full
  cropped = full[:, :, :kh, :kw]
  if out is None:
    return cropped.contiguous()
  out.copy_(cropped)
  return out


def descend(images_padded, dictionary, codes, hessian_diagonal, kernel_stride, padding_dims, stepsize, num_iters,
            lowest_code_val, normalize_dictionary, batch_global=None):
  """
  num_iters steps, in place, of convolutional/sc_cheap_quadratic_descent.py:59-79 (hessian_diagonal=None:
  sc_steepest_descent.py:55-72): gradient / b [/ (h + eps)], rescaled to the norm of the dictionary, subtracted, every
  kernel renormalised. With data parallelism enabled the gradient sum is all-reduced and b is the global batch.
  """
  for t, name in ((images_padded, 'images_padded'), (dictionary, 'dictionary'), (codes, 'codes')):
    _lib.require_cuda_f32(t, name)
  if hessian_diagonal is not None:
    _lib.require_cuda_f32(hessian_diagonal, 'hessian_diagonal')
    hessian_diagonal = hessian_diagonal.contiguous()
  lib = _lib.load()
  device = dictionary.device
  S = dictionary.size(0)
  per_kernel = dictionary[0].numel()
  target = dictionary
  work = dictionary if dictionary.is_contiguous() else dictionary.contiguous()
  if batch_global is None:
    batch_global = global_batch_size(images_padded.size(0), device)
  for _ in range(int(num_iters)):
    grad = dictionary_gradient(images_padded, work, codes, kernel_stride, padding_dims)
    if config.data_parallel:
      import torch.distributed as dist
      dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=config.process_group)
    with torch.cuda.device(device):
      _lib.check(lib.vtc_sc_conv_dict_apply(_lib.ptr(work), _lib.ptr(grad), _lib.ptr(hessian_diagonal), S, per_kernel,
                                            int(batch_global), float(stepsize), float(lowest_code_val),
                                            int(bool(normalize_dictionary)), _lib.stream_ptr(device)))
  if work is not target:
    target.copy_(work)


def hessian_running_mean(hessian_diagonal, codes, batch_global=None):
  """In place: h <- 0.99 h + mean_images(sum_positions(codes^2)) / 100 (training/sparse_coding.py:158-161)."""
  _lib.require_cuda_f32(codes, 'codes')
  _lib.require_cuda_f32(hessian_diagonal, 'hessian_diagonal')
  lib = _lib.load()
  device = codes.device
  B, S = codes.size(0), codes.size(1)
  positions = codes[0, 0].numel()
  sq = torch.empty(S, dtype=torch.float32, device=device)
  codes_c = codes.contiguous()
  with torch.cuda.device(device):
    if config.data_parallel:
      import torch.distributed as dist
      _lib.check(lib.vtc_conv_hessian_diag_update(_lib.ptr(codes_c), B, S, positions, B, _lib.ptr(sq), None, 0,
                                                  _lib.stream_ptr(device)))
      dist.all_reduce(sq, op=dist.ReduceOp.SUM, group=config.process_group)
      total = global_batch_size(B, device) if batch_global is None else batch_global
      _lib.check(lib.vtc_hessian_ema(_lib.ptr(hessian_diagonal), _lib.ptr(sq), S, int(total), _lib.stream_ptr(device)))
    else:
      _lib.check(lib.vtc_conv_hessian_diag_update(_lib.ptr(codes_c), B, S, positions, B, _lib.ptr(sq),
                                                  _lib.ptr(hessian_diagonal), 1, _lib.stream_ptr(device)))
  return hessian_diagonal

"""
Iterative Shrinkage/Thresholding (ISTA / FISTA) for fully-connected sparse inference, on one B200.

Drop-in for the reference module of the same dotted name
(vision_transform_codes/analysis_transforms/fully_connected/ista_fista.py:14-148): same signature, same return
value, same exceptions. The arithmetic runs in ``vtc_fista_fc`` (include/vtc_b200.h), which picks the contraction
with fewer flops: for s > 2 n the reference's own synthesis form ``a <- prox(y - eta ((y Phi - x) Phi^T))`` -- all
iterations in ONE persistent launch of the panel-resident tcgen05 kernel when n <= 256 (csrc/fista_iter2_kernel.cuh),
two launches per iteration otherwise -- and for s <= 2 n the Gram form ``a <- prox(y - eta (y G - b))`` with
``G = Phi Phi^T`` and ``b = x Phi^T`` precomputed by tcgen05 GEMMs. In every schedule the gradient step, the threshold
and the FISTA momentum are the epilogue of the contraction (DESIGN.md sections 2 and 4).
"""
import ctypes
import os
import sys

import torch

try:
  from vision_transform_codes_b200 import _lib, config
except ImportError:  # used through sys.path insertion of the package root (install())
  sys.path.append(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
  from vision_transform_codes_b200 import _lib, config

VARIANTS = {'ista': 0, 'fista': 1}


def infer(images, dictionary, sparsity_weight, num_iters, variant, initial_codes, early_stopping_epsilon,
          nonnegative_only, hard_threshold, group_size, eps_scale=1.0, precision=None):
  """Shared by the vanilla and the subspace front ends. Returns (codes, iterations actually run)."""
  _lib.require_cuda_f32(images, 'images')
  _lib.require_cuda_f32(dictionary, 'dictionary')
  if images.dim() != 2 or dictionary.dim() != 2 or images.size(1) != dictionary.size(1):
    raise ValueError('expected images (b, n) and dictionary (s, n), got %s and %s' %
                     (tuple(images.shape), tuple(dictionary.shape)))
  if num_iters < 1:
    # the reference falls out of its while loop and returns a name that was never bound (ista_fista.py:100,148)
    raise UnboundLocalError("cannot access local variable 'codes' where it is not associated with a value")
  device = images.device
  if dictionary.device != device:
    raise RuntimeError('images and dictionary must be on the same device')
  lib = _lib.load()
  B, D = images.shape
  S = dictionary.size(0)
  if B == 0:
    # an empty batch goes through the reference's loop untouched: codes of shape (0, s)
    return torch.empty((0, S), dtype=torch.float32, device=device), int(num_iters)
  images_rm, ld_images = _lib.row_major(images)
  dictionary_c = dictionary.contiguous()
  codes = torch.empty((B, S), dtype=torch.float32, device=device)
  init = None
  if initial_codes is not None:
    _lib.require_cuda_f32(initial_codes, 'initial_codes')
    if tuple(initial_codes.shape) != (B, S):
      raise ValueError('initial_codes must have shape (b, s)')
    init = initial_codes.contiguous()
  prec = config.inference_precision_code(hard_threshold) if precision is None else precision
  with torch.cuda.device(device):
    nbytes = lib.vtc_fista_workspace_bytes(B, S, D, prec)
    ws = _lib.workspace(nbytes, device, 'fista')
    iters_run = ctypes.c_int(0)
    lipschitz = ctypes.c_float(0.0)
    eps = -1.0 if early_stopping_epsilon is None else float(early_stopping_epsilon) * eps_scale
    if early_stopping_epsilon is not None and eps < 0:
      eps = 0.0
    rc = lib.vtc_fista_fc(
        _lib.ptr(images_rm), ld_images, _lib.ptr(dictionary_c), _lib.ptr(init), _lib.ptr(codes), S, B, S, D,
        float(sparsity_weight), int(num_iters), VARIANTS[variant], int(bool(nonnegative_only)),
        int(bool(hard_threshold)), int(group_size), eps, prec, _lib.ptr(ws), ws.numel(), ctypes.byref(iters_run),
        ctypes.byref(lipschitz) if config.check_finite else None, _lib.stream_ptr(device))
  if rc == _lib.VTC_ERR_NONFINITE:
    print('symeig threw an exception. Likely due to one of the dictionary',
          'elements overflowing. The norm of each dictionary element is')
    print(torch.norm(dictionary, dim=1, p=2))
    raise RuntimeError()
  _lib.check(rc)
  return codes, iters_run.value


def run(images, dictionary, sparsity_weight, num_iters, variant='fista',
        initial_codes=None, early_stopping_epsilon=None,
        nonnegative_only=False, hard_threshold=False):
  """
  Runs steps of Iterative Shrinkage/Thresholding with a constant stepsize

  Parameters
  ----------
  images : torch.Tensor(float32, size=(b, n))
      A batch of images (patches) to find the sparse code for.
  dictionary : torch.Tensor(float32, size=(s, n))
      The dictionary of basis functions; atoms are rows.
  sparsity_weight : float or 0-dim tensor
      Weight on the sparsity term (lambda).
  num_iters : int
      Number of steps of ISTA/FISTA to run.
  variant : str, optional
      One of {'ista', 'fista'}. Default 'fista'.
  initial_codes : torch.Tensor(float32, size=(b, s)), optional
      Warm start. Never modified. Default None.
  early_stopping_epsilon : float, optional
      Terminate if the mean absolute change of the codes per component, divided by the stepsize, drops below this
      (checked on the host once per iteration, as in the reference). Default None.
  nonnegative_only : bool, optional
      Shifted-ReLU threshold instead of the two-sided one. Default False.
  hard_threshold : bool, optional
      Identity outside the zeroed region. Default False.

  Returns
  -------
  codes : torch.Tensor(float32, size=(b, s))
  """
  assert variant in ['ista', 'fista']
  codes, _ = infer(images, dictionary, sparsity_weight, num_iters, variant, initial_codes, early_stopping_epsilon,
                   nonnegative_only, hard_threshold, group_size=1)
  return codes

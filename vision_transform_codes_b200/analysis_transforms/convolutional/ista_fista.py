"""
Iterative Shrinkage/Thresholding (ISTA / FISTA) for convolutional sparse inference, on one B200.

Drop-in for the reference module of the same dotted name
(vision_transform_codes/analysis_transforms/convolutional/ista_fista.py:17-197): same signature, same return value,
same exceptions. The arithmetic runs in ``vtc_fista_conv`` (include/vtc_b200.h): the strided ``conv_transpose2d`` /
``conv2d`` pair of every iteration is two tcgen05 GEMMs over stride-sized image blocks (kernel taps = row shifts of the
operand), the reconstruction mask, the gradient step, the threshold and the FISTA momentum are their epilogues.
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


def geometry(images_padded, dictionary, kernel_stride, padding_dims):
  """Validated (B, C, H, W, S, KH, KW, SY, SX, pad_top, pad_bottom, pad_left, pad_right, SH, SW)."""
  if images_padded.dim() != 4 or dictionary.dim() != 4 or images_padded.size(1) != dictionary.size(1):
    raise ValueError('expected images_padded (b, c, h, w) and dictionary (s, c, kh, kw), got %s and %s' %
                     (tuple(images_padded.shape), tuple(dictionary.shape)))
  B, C, H, W = images_padded.shape
  S, _, KH, KW = dictionary.shape
  SY, SX = int(kernel_stride[0]), int(kernel_stride[1])
  if padding_dims is None:
    pads = (0, 0, 0, 0)
  else:
    pads = (int(padding_dims[0][0]), int(padding_dims[0][1]), int(padding_dims[1][0]), int(padding_dims[1][1]))
    if pads[1] == 0 or pads[3] == 0:
      # create_mask (utils/convolutions.py:21,23) writes ``mask[:, :, -padding[0][1]:, :] = 0``: with a trailing padding
      # of 0 the slice ``-0:`` is the WHOLE axis, so the reference's mask is zero everywhere and its codes never leave
      # the starting point's prox. Reproduced as-is (an empty un-masked region); pass padding_dims=None for "no mask".
      pads = (int(images_padded.size(2)), 0, int(images_padded.size(3)), 0)
  if H < KH or W < KW or (H - KH) % SY != 0 or (W - KW) % SX != 0:
    # conv_transpose2d of the codes would not have the shape of images_padded: the reference fails on the subtraction
    raise RuntimeError('padded image size (%d, %d) is not kernel size (%d, %d) plus a whole number of strides (%d, %d)'
                       % (H, W, KH, KW, SY, SX))
  SH, SW = (H - KH) // SY + 1, (W - KW) // SX + 1
  return (B, C, H, W, S, KH, KW, SY, SX) + pads + (SH, SW)


def align_to_stride(images_padded, dictionary, geo):
  """
  The kernels treat a kernel as ceil(k / stride) taps of one stride each, so a kernel size that is not a multiple of
  the stride (the reference takes any, analysis_transforms/convolutional/ista_fista.py:119-122 with the ceil of
  utils/convolutions.py:14-15) is zero-padded at its trailing edge up to the next multiple, and the images get the
  same number of trailing zero rows / columns, counted as padding (masked). Nothing changes numerically: the extra taps
  multiply zeros in the analysis, the extra pixels of the synthesis lie in the masked border, the code grid and the
  Gram matrix of the flattened kernels (step size) are the same. Returns (images, dictionary, geometry) to compute with.
  """
  B, C, H, W, S, KH, KW, SY, SX, pt, pb, pl, pr, SH, SW = geo
  ey, ex = (-KH) % SY, (-KW) % SX
  if ey == 0 and ex == 0:
    return images_padded, dictionary, geo
  images_padded = torch.nn.functional.pad(images_padded, (0, ex, 0, ey))
  dictionary = torch.nn.functional.pad(dictionary, (0, ex, 0, ey))
  return images_padded, dictionary, (B, C, H + ey, W + ex, S, KH + ey, KW + ex, SY, SX, pt, pb + ey, pl, pr + ex, SH, SW)


def infer(images_padded, dictionary, kernel_stride, padding_dims, sparsity_weight, num_iters, variant,
          initial_codes, early_stopping_epsilon, nonnegative_only, hard_threshold, precision=None):
  """Returns (codes, iterations actually run)."""
  _lib.require_cuda_f32(images_padded, 'images_padded')
  _lib.require_cuda_f32(dictionary, 'dictionary')
  geo = geometry(images_padded, dictionary, kernel_stride, padding_dims)
  images_padded, dictionary, geo = align_to_stride(images_padded, dictionary, geo)
  B, C, H, W, S, KH, KW, SY, SX, pt, pb, pl, pr, SH, SW = geo
  if num_iters < 1:
    # the reference falls out of its while loop and returns a name that was never bound (ista_fista.py:141,197)
    raise UnboundLocalError("cannot access local variable 'codes' where it is not associated with a value")
  device = images_padded.device
  if dictionary.device != device:
    raise RuntimeError('images_padded and dictionary must be on the same device')
  lib = _lib.load()
  images_c = images_padded.contiguous()
  dictionary_c = dictionary.contiguous()
  init = None
  if initial_codes is not None:
    _lib.require_cuda_f32(initial_codes, 'initial_codes')
    assert initial_codes.shape[0] == B
    assert initial_codes.shape[1] == S
    assert initial_codes.shape[2] == SH
    assert initial_codes.shape[3] == SW
    init = initial_codes.contiguous()
  codes = torch.empty((B, S, SH, SW), dtype=torch.float32, device=device)
  if B == 0:
    return codes, int(num_iters)   # an empty batch: codes of shape (0, s, sh, sw), as the reference's loop leaves it
  prec = config.inference_precision_code(hard_threshold) if precision is None else precision
  with torch.cuda.device(device):
    nbytes = lib.vtc_fista_conv_workspace_bytes(B, C, H, W, S, KH, KW, SY, SX, prec)
    if nbytes == 0:
      # let the library produce the precise message (unsupported kernel / stride combination, ...)
      nbytes = 16
    ws = _lib.workspace(nbytes, device, 'fista_conv')
    iters_run = ctypes.c_int(0)
    lipschitz = ctypes.c_float(0.0)
    eps = -1.0 if early_stopping_epsilon is None else max(float(early_stopping_epsilon), 0.0)
    rc = lib.vtc_fista_conv(
        _lib.ptr(images_c), _lib.ptr(dictionary_c), _lib.ptr(init), _lib.ptr(codes), B, C, H, W, S, KH, KW, SY, SX,
        pt, pb, pl, pr, float(sparsity_weight), int(num_iters), VARIANTS[variant], int(bool(nonnegative_only)),
        int(bool(hard_threshold)), eps, prec, _lib.ptr(ws), ws.numel(), ctypes.byref(iters_run),
        ctypes.byref(lipschitz) if config.check_finite else None, _lib.stream_ptr(device))
  if rc == _lib.VTC_ERR_NONFINITE:
    print('symeig threw an exception. Likely due to one of the dictionary',
          'elements overflowing. The norm of each dictionary element is')
    print(torch.norm(dictionary, dim=[1, 2, 3], p=2))
    raise RuntimeError()
  _lib.check(rc)
  return codes, iters_run.value


def run(images_padded, dictionary, kernel_stride, padding_dims,
        sparsity_weight, num_iters, variant='fista', initial_codes=None,
        early_stopping_epsilon=None, nonnegative_only=False,
        hard_threshold=False):
  """
  Runs steps of Iterative Shrinkage/Thresholding with a constant stepsize

  Parameters
  ----------
  images_padded : torch.Tensor(float32, size=(b, c, h, w))
      A batch of (already padded) images to find the convolutional sparse code for.
  dictionary : torch.Tensor(float32, size=(s, c, kh, kw))
      The kernels; s is the number of channels of the code.
  kernel_stride : tuple(int, int)
      Vertical and horizontal stride of the kernels.
  padding_dims : tuple(tuple(int, int), tuple(int, int))
      (leading, trailing) padding of the images, vertical then horizontal: the reconstruction error in this border is
      ignored (utils.convolutions.create_mask).
  sparsity_weight : float
      Weight on the sparsity term (lambda).
  num_iters : int
      Number of steps of ISTA/FISTA to run.
  variant : str, optional
      One of {'ista', 'fista'}. Default 'fista'.
  initial_codes : torch.Tensor(float32, size=(b, s, sh, sw)), optional
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
  codes : torch.Tensor(float32, size=(b, s, sh, sw))
  """
  assert variant in ['ista', 'fista']
  codes, _ = infer(images_padded, dictionary, kernel_stride, padding_dims, sparsity_weight, num_iters, variant,
                   initial_codes, early_stopping_epsilon, nonnegative_only, hard_threshold)
  return codes

"""
Host-side geometry helpers of the convolutional transforms.

Same names and results as the reference's vision_transform_codes/utils/convolutions.py:7-24, so that scripts written
against the reference keep working when only this package is on the path (with the reference on the path its own
``utils`` package is the one that resolves). The CUDA path does not use ``create_mask``: the kernels apply the mask
from the padding amounts (csrc/gemm_kernel.cuh, EPI_STORE); it is here for callers that strip or show the border.
"""
import math

import torch


def get_padding_amt(image_dim, kernel_dim, dim_stride):
  """(leading, trailing) padding of one image axis (reference :7-12): one kernel minus one stride on either side,
  and the trailing side additionally completes the last partial stride of the image."""
  overlap = kernel_dim - dim_stride
  remainder = image_dim % dim_stride
  return overlap, overlap + ((dim_stride - remainder) if remainder else 0)


def code_dim_from_padded_img_dim(padded_image_dim, kernel_dim, dim_stride):
  """Number of kernel positions along one axis (reference :14-15)."""
  return int(math.ceil((padded_image_dim - kernel_dim) / dim_stride)) + 1


def _axis_indicator(size, lead, trail, like):
  """1 where an index of this axis belongs to the un-padded image. The reference clears ``[-trail:]`` (:21, :23), so a
  trailing padding of 0 clears the WHOLE axis (``-0:`` is everything) -- reproduced, not fixed."""
  idx = torch.arange(size, device=like.device)
  keep = idx >= lead
  keep &= (idx < size - trail) if trail != 0 else torch.zeros_like(keep)
  return keep.to(like.dtype)


def create_mask(images_with_padding, padding):
  """Ones on the image, zeros on the padded border (reference :17-24); ``padding`` = ((top, bottom), (left, right))."""
  if padding is None:
    return torch.ones_like(images_with_padding)
  h, w = images_with_padding.shape[-2], images_with_padding.shape[-1]
  rows = _axis_indicator(h, padding[0][0], padding[0][1], images_with_padding)
  cols = _axis_indicator(w, padding[1][0], padding[1][1], images_with_padding)
  return (rows[:, None] * cols[None, :]).expand_as(images_with_padding).contiguous()

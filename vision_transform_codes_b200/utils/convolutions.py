"""
Simple host-side utilities for convolutional transforms.

Same names and results as the reference's vision_transform_codes/utils/convolutions.py:7-24 (get_padding_amt,
code_dim_from_padded_img_dim, create_mask), so that example scripts written against the reference keep working when
only this package is on the path. create_mask is not used by the CUDA path (the mask is applied inside the GEMM
epilogue from the padding amounts); it is here for callers that strip or visualise the padded border.
"""
import math

import torch


def get_padding_amt(image_dim, kernel_dim, dim_stride):
  """(leading, trailing) padding so that strided kernels tile the padded image exactly."""
  leading_padding = kernel_dim - dim_stride
  trailing_padding = kernel_dim - dim_stride
  if image_dim % dim_stride != 0:
    trailing_padding += (dim_stride - (image_dim % dim_stride))
  return leading_padding, trailing_padding


def code_dim_from_padded_img_dim(padded_image_dim, kernel_dim, dim_stride):
  return 1 + int(math.ceil((padded_image_dim - kernel_dim) / dim_stride))


def create_mask(images_with_padding, padding):
  mask = torch.ones_like(images_with_padding)
  if padding is not None:
    mask[:, :, 0:padding[0][0], :] = 0.0
    mask[:, :, -padding[0][1]:, :] = 0.0
    mask[:, :, :, 0:padding[1][0]] = 0.0
    mask[:, :, :, -padding[1][1]:] = 0.0
  return mask

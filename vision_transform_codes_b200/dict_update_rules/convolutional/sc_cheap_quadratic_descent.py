"""
Updates dictionary with a modified descent for convolutional sparse coding, on B200.

Drop-in for vision_transform_codes/dict_update_rules/convolutional/sc_cheap_quadratic_descent.py:14-79: the diagonal of
the Hessian rescales the update of every kernel, the whole update is put on the scale of the dictionary.
"""
import os
import sys

try:
  from vision_transform_codes_b200.dict_update_rules.convolutional import _common
except ImportError:
  sys.path.append(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
  from vision_transform_codes_b200.dict_update_rules.convolutional import _common


def run(images_padded, dictionary, codes, hessian_diagonal,
        kernel_stride, padding_dims, stepsize=0.001, num_iters=1,
        lowest_code_val=0.001, normalize_dictionary=True):
  """
  Runs num_iters steps of an approximate quadratic descent, in place on ``dictionary``

  Parameters
  ----------
  images_padded : torch.Tensor(float32, size=(b, c, h, w))
  dictionary : torch.Tensor(float32, size=(s, c, kh, kw))
      Updated in place.
  codes : torch.Tensor(float32, size=(b, s, sh, sw))
  hessian_diagonal : torch.Tensor(float32, size=(s,))
      Estimate of the diagonal of the Hessian, maintained by the caller.
  kernel_stride : tuple(int, int)
  padding_dims : tuple(tuple(int, int), tuple(int, int))
  stepsize : float, optional
      Default 0.001.
  num_iters : int, optional
      Default 1.
  lowest_code_val : float, optional
      Conditions the Hessian diagonal away from zero. Default 0.001
  normalize_dictionary : bool, optional
      Renormalise every kernel to unit L2 norm after each step. Default True.
  """
  _common.descend(images_padded, dictionary, codes, hessian_diagonal, kernel_stride, padding_dims, stepsize,
                  num_iters, lowest_code_val, normalize_dictionary)

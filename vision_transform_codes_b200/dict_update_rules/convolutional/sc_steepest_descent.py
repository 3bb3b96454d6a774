"""
Updates dictionary with steepest descent for convolutional sparse coding, on B200.

Drop-in for vision_transform_codes/dict_update_rules/convolutional/sc_steepest_descent.py:12-72.
"""
import os
import sys

try:
  from vision_transform_codes_b200.dict_update_rules.convolutional import _common
except ImportError:
  sys.path.append(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
  from vision_transform_codes_b200.dict_update_rules.convolutional import _common


def run(images_padded, dictionary, codes, kernel_stride, padding_dims,
        stepsize=0.001, num_iters=1, normalize_dictionary=True):
  """
  Runs num_iters steps of steepest descent, in place on ``dictionary``

  Parameters
  ----------
  images_padded : torch.Tensor(float32, size=(b, c, h, w))
  dictionary : torch.Tensor(float32, size=(s, c, kh, kw))
      Updated in place.
  codes : torch.Tensor(float32, size=(b, s, sh, sw))
  kernel_stride : tuple(int, int)
  padding_dims : tuple(tuple(int, int), tuple(int, int))
  stepsize : float, optional
      Default 0.001.
  num_iters : int, optional
      Default 1.
  normalize_dictionary : bool, optional
      Renormalise every kernel to unit L2 norm after each step. Default True.
  """
  _common.descend(images_padded, dictionary, codes, None, kernel_stride, padding_dims, stepsize, num_iters, 0.0,
                  normalize_dictionary)

"""
Updates dictionary with a modified descent for fully-connected sparse coding, on B200.

Drop-in for vision_transform_codes/dict_update_rules/fully_connected/sc_cheap_quadratic_descent.py:11-48: the
diagonal of the Hessian rescales the steepest-descent update of every dictionary element.
"""
import os
import sys

try:
  from vision_transform_codes_b200.dict_update_rules.fully_connected import _common
except ImportError:
  sys.path.append(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
  from vision_transform_codes_b200.dict_update_rules.fully_connected import _common


def run(images, dictionary, codes, hessian_diagonal, stepsize=0.001,
        num_iters=1, lowest_code_val=0.001, normalize_dictionary=True):
  """
  Runs num_iters steps of an approximate quadratic descent, in place on ``dictionary``

  Parameters
  ----------
  images : torch.Tensor(float32, size=(b, n))
  dictionary : torch.Tensor(float32, size=(s, n))
      Updated in place.
  codes : torch.Tensor(float32, size=(b, s))
  hessian_diagonal : torch.Tensor(float32, size=(s,))
      Estimate of the diagonal of the Hessian, maintained by the caller.
  stepsize : float, optional
      Default 0.001.
  num_iters : int, optional
      Default 1.
  lowest_code_val : float, optional
      Conditions the Hessian diagonal away from zero. Default 0.001
  normalize_dictionary : bool, optional
      Renormalise every dictionary element to unit L2 norm after each step. Default True.
  """
  _common.descend(images, dictionary, codes, hessian_diagonal, stepsize, num_iters, lowest_code_val,
                  normalize_dictionary)

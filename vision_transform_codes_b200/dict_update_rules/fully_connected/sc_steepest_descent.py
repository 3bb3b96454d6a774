"""
Updates dictionary with steepest descent for fully-connected sparse coding, on B200.

Drop-in for vision_transform_codes/dict_update_rules/fully_connected/sc_steepest_descent.py:9-41.
"""
import os
import sys

try:
  from vision_transform_codes_b200.dict_update_rules.fully_connected import _common
except ImportError:
  sys.path.append(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
  from vision_transform_codes_b200.dict_update_rules.fully_connected import _common


def run(images, dictionary, codes, stepsize=0.001, num_iters=1,
        normalize_dictionary=True):
  """
  Runs num_iters steps of SC steepest descent on the dictionary elements, in place on ``dictionary``

  Parameters
  ----------
  images : torch.Tensor(float32, size=(b, n))
  dictionary : torch.Tensor(float32, size=(s, n))
      Updated in place.
  codes : torch.Tensor(float32, size=(b, s))
  stepsize : float, optional
      Default 0.001.
  num_iters : int, optional
      Default 1.
  normalize_dictionary : bool, optional
      Default True.
  """
  _common.descend(images, dictionary, codes, None, stepsize, num_iters, 0.0, normalize_dictionary)

"""
Dictionary update for subspace sparse coding (steepest descent), on B200.

The reference trainer selects this module by name (vision_transform_codes/training/sparse_coding.py:421-427, picked
by tests/sparse_coding_5.py:43) and calls it with ``{dictionary, codes, stepsize, num_iters, images,
group_assignments, alignment_penalty}`` (:144-168) -- but the reference tree ships no such file, so that configuration
ends in an ImportError there. This module is what those call sites expect: sc_steepest_descent
(dict_update_rules/fully_connected/sc_steepest_descent.py:37-41) plus the within-group alignment regulariser of
subspace_sc_cheap_quadratic_descent.py:59-79, :91-127, i.e. that rule without its division by the Hessian diagonal:

  dictionary <- rownorm(dictionary - stepsize * (codes^T (codes dictionary - images) / b + alignment_penalty * R))

The argument order follows subspace_sc_cheap_quadratic_descent.run with ``hessian_diagonal`` removed, the way
sc_steepest_descent.run relates to sc_cheap_quadratic_descent.run.
"""
import os
import sys

try:
  from vision_transform_codes_b200.dict_update_rules.fully_connected import _common
except ImportError:
  sys.path.append(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
  from vision_transform_codes_b200.dict_update_rules.fully_connected import _common


def run(images, dictionary, codes, group_assignments, alignment_penalty,
        stepsize=0.001, num_iters=1, normalize_dictionary=True):
  """
  Runs num_iters steps of steepest descent with the alignment regulariser, in place on ``dictionary``

  Parameters
  ----------
  images : torch.Tensor(float32, size=(b, n))
  dictionary : torch.Tensor(float32, size=(s, n))
      Updated in place.
  codes : torch.Tensor(float32, size=(b, s))
  group_assignments : list(array_like)
      Groups of dictionary elements; an element may belong to several groups.
  alignment_penalty : float
      Weight of the within-group alignment regulariser (0: exactly sc_steepest_descent).
  stepsize : float, optional
      Default 0.001.
  num_iters : int, optional
      Default 1.
  normalize_dictionary : bool, optional
      Default True.
  """
  _common.descend(images, dictionary, codes, None, stepsize, num_iters, 0.0, normalize_dictionary,
                  group_assignments=group_assignments, alignment_penalty=alignment_penalty)

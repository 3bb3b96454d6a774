"""
Dictionary update for subspace sparse coding (cheap quadratic descent), on B200.

Drop-in for vision_transform_codes/dict_update_rules/fully_connected/subspace_sc_cheap_quadratic_descent.py:13-88.
With ``alignment_penalty == 0`` the rule is exactly sc_cheap_quadratic_descent (reference :80-88). Otherwise the
gradient of the within-group alignment penalty (reference :59-79, :91-127) is computed by ``vtc_subspace_alignment_grad``
(one thread block per group) and folded into the apply step.
"""
import os
import sys

try:
  from vision_transform_codes_b200.dict_update_rules.fully_connected import _common
except ImportError:
  sys.path.append(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
  from vision_transform_codes_b200.dict_update_rules.fully_connected import _common


def run(images, dictionary, codes, group_assignments, hessian_diagonal,
        alignment_penalty, stepsize=0.001, num_iters=1,
        lowest_code_val=0.001, normalize_dictionary=True):
  """
  Runs num_iters steps of an approximate quadratic descent, in place on ``dictionary``

  Parameters
  ----------
  images : torch.Tensor(float32, size=(b, n))
  dictionary : torch.Tensor(float32, size=(s, n))
  codes : torch.Tensor(float32, size=(b, s))
  group_assignments : list(array_like)
      Groups of dictionary elements; an element may belong to several groups.
  hessian_diagonal : torch.Tensor(float32, size=(s,))
  alignment_penalty : float
      Weight of the within-group alignment regulariser.
  stepsize, num_iters, lowest_code_val, normalize_dictionary : see sc_cheap_quadratic_descent.run
  """
  _common.descend(images, dictionary, codes, hessian_diagonal, stepsize, num_iters, lowest_code_val,
                  normalize_dictionary, group_assignments=group_assignments, alignment_penalty=alignment_penalty)

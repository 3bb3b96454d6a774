"""
B200-native sparse-coding hot path of spencerkent/vision-transform-codes.

This directory plays the role of the reference's ``vision_transform_codes/`` package root: it holds modules at the
same dotted names the reference trainer imports by string
(``analysis_transforms.fully_connected.{ista_fista,subspace_ista_fista}``,
``dict_update_rules.fully_connected.{sc_cheap_quadratic_descent,sc_steepest_descent,
subspace_sc_cheap_quadratic_descent}``; reference: vision_transform_codes/training/sparse_coding.py:389-439).
``install()`` puts it at the front of ``sys.path`` exactly like the reference's ``examples/_set_the_path.py:6-10``
does for its own root, so an unmodified ``train_dictionary`` picks the CUDA implementations up.

Underneath is one C-ABI shared library (``lib/libvtc_b200.so``, declared in ``include/vtc_b200.h``) of hand-written
sm_100a kernels. There is no CPU or PyTorch fallback: without the library or without a B200 every ``run`` raises.
"""
import os
import sys

PACKAGE_ROOT = os.path.dirname(os.path.abspath(__file__))

PRECISIONS = {'bf16': 1, 'bf16x3': 3, 'bf16x6': 6}


class _Config:
  """Process-wide knobs that the reference API has no argument for."""

  def __init__(self):
    # arithmetic of the tensor-core contractions; 'bf16x3' is the float32-parity path
    self.precision = os.environ.get('VTC_B200_PRECISION', 'bf16x3')
    # hard thresholding is discontinuous: a product error of 2^-17 (bf16x3) is enough to flip a coefficient that sits at
    # the cutoff, and the iteration then follows another trajectory (measured: codes rel-L2 up to 8e-2 after 60
    # iterations, where the reference's own float32 and float64 runs agree to 1e-6). Calls with hard_threshold=True
    # made in the float32-parity mode therefore use this stricter arithmetic; set it to 'bf16x3' to opt out.
    self.hard_threshold_precision = os.environ.get('VTC_B200_HARD_PRECISION', 'bf16x6')
    # precision of the dictionary-gradient contractions (tiny next to inference)
    self.update_precision = os.environ.get('VTC_B200_UPDATE_PRECISION', 'bf16x6')
    # synchronise once per inference call to reproduce the reference's RuntimeError on an overflowed dictionary
    self.check_finite = os.environ.get('VTC_B200_CHECK_FINITE', '1') != '0'
    # torch.distributed process group used to all-reduce the dictionary gradient (None = single GPU)
    self.process_group = None
    self.data_parallel = False

  def inference_precision_code(self, hard_threshold=False):
    """Arithmetic of an inference call: `precision`, except that hard-threshold calls in the parity mode are strict."""
    if hard_threshold and self.precision == 'bf16x3':
      return self.precision_code('hard_threshold_precision')
    return self.precision_code()

  def precision_code(self, which='precision'):
    name = getattr(self, which)
    if name not in PRECISIONS:
      raise ValueError('unknown precision %r, expected one of %s' % (name, sorted(PRECISIONS)))
    return PRECISIONS[name]


config = _Config()


def install():
  """Make ``analysis_transforms.*`` / ``dict_update_rules.*`` resolve to this package's CUDA implementations."""
  if PACKAGE_ROOT in sys.path:
    sys.path.remove(PACKAGE_ROOT)
  sys.path.insert(0, PACKAGE_ROOT)
  for name in list(sys.modules):
    if name.split('.')[0] in ('analysis_transforms', 'dict_update_rules'):
      del sys.modules[name]  # drop anything already resolved from elsewhere


def enable_data_parallel(process_group=None):
  """Shard-the-batch / replicate-the-dictionary mode: every dictionary update all-reduces its gradient (sum)."""
  import torch.distributed as dist
  if not dist.is_initialized():
    raise RuntimeError('torch.distributed is not initialised')
  config.process_group = process_group
  config.data_parallel = True


def disable_data_parallel():
  config.process_group = None
  config.data_parallel = False

"""Test infrastructure: access to the staged, UNMODIFIED reference (oracle/_ref, see tools/stage_reference.py).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use this module; nothing under
vision_transform_codes_b200/ does. Two shims are applied around the reference, neither edits it:
  * torch.symeig (removed from current torch; the reference calls torch.symeig(M)[0][-1],
    analysis_transforms/fully_connected/ista_fista.py:73) -> torch.linalg.eigvalsh with the same conventions
  * utils.plotting is stubbed in sys.modules when its own imports (skimage, matplotlib) are not installed: the trainer
    imports it at module level (training/sparse_coding.py:7) but only uses it under a visualisation schedule
"""
import contextlib
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, '_ref')
TOP_LEVEL = ('analysis_transforms', 'dict_update_rules', 'training', 'utils')


def root():
  """The directory the reference expects on sys.path (its examples/_set_the_path.py:6-10), staged under oracle/_ref."""
  path = os.path.join(STAGED, 'vision_transform_codes')
  if not os.path.isfile(os.path.join(STAGED, 'MANIFEST.json')):
    raise RuntimeError('the reference is not staged: run  python tools/stage_reference.py  where /root/reference exists '
                       '(oracle/_ref is git-ignored and travels to the GPU box with the snapshot)')
  return path


def available():
  return os.path.isfile(os.path.join(STAGED, 'MANIFEST.json'))


def apply_shims():
  import torch
  try:   # current torch keeps the name but raises from it
    torch.symeig(torch.eye(2))
  except Exception:
    torch.symeig = lambda A, eigenvectors=False, upper=True: (
        torch.linalg.eigvalsh(A, UPLO='U' if upper else 'L'), None)
  if 'utils.plotting' not in sys.modules:
    try:
      import skimage  # noqa: F401
      import matplotlib  # noqa: F401
    except ImportError:
      sys.modules['utils.plotting'] = types.ModuleType('utils.plotting')


def forget_modules():
  """Drops every cached module of the reference's top-level names (ours and the reference's share them by design)."""
  for name in list(sys.modules):
    if name.split('.')[0] in TOP_LEVEL:
      del sys.modules[name]


@contextlib.contextmanager
def reference_only():
  """sys.path with the staged reference root in front and none of this repo's drop-in roots: the four top-level names
  resolve to the reference's own CPU implementation."""
  saved = list(sys.path)
  forget_modules()
  apply_shims()
  sys.path[:] = [root()] + [p for p in saved if not os.path.isdir(os.path.join(p, 'analysis_transforms'))]
  try:
    yield
  finally:
    sys.path[:] = saved
    forget_modules()


@contextlib.contextmanager
def reference_on_drop_ins():
  """install() puts this repo's drop-in root first, the reference root comes after it, so
  the reference's trainer (training.sparse_coding) runs on the CUDA implementations it imports by dotted name."""
  import vision_transform_codes_b200 as pkg
  saved = list(sys.path)
  forget_modules()
  apply_shims()
  pkg.install()
  sys.path.append(root())
  try:
    yield
  finally:
    sys.path[:] = saved
    forget_modules()


def load(dotted):
  """importlib.import_module under whichever of the two contexts above is active."""
  return importlib.import_module(dotted)

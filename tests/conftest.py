import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA (sm_100) device; run with -m gpu on the B200 box')


def load_golden(name):
  data = np.load(os.path.join(GOLDEN, name + '.npz'))
  out = {}
  for k in data.files:
    v = data[k]
    out[k] = v.item() if v.ndim == 0 else (torch.from_numpy(v) if v.dtype.kind in 'fiub' else v)
  return out


@pytest.fixture(scope='session')
def golden():
  return load_golden


def ragged_groups(g):
  sizes = g['ragged_sizes'].tolist()
  flat = g['ragged_flat'].tolist()
  groups, pos = [], 0
  for n in sizes:
    groups.append([int(v) for v in flat[pos:pos + n]])
    pos += n
  return groups

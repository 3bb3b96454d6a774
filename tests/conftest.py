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


# ---- parity records: every comparison of the CUDA path with reference outputs appends one record; the session writes
# them to gpurun_out/parity_r02.json (the directory gpurun brings back; copied to profiles/ by hand) -- SURVEY 7.3-1:
# the guard band AND the count of flips inside it belong in the report
PARITY_RECORDS = []


def record_parity(test, case, **numbers):
  PARITY_RECORDS.append(dict(test=test, case=case, **numbers))


def pytest_sessionfinish(session, exitstatus):
  if not PARITY_RECORDS:
    return
  import json
  out_dir = os.path.join(ROOT, 'gpurun_out')
  try:
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, 'parity_r02.json'), 'w') as f:
      json.dump({'records': PARITY_RECORDS, 'device': torch.cuda.get_device_name(0) if torch.cuda.is_available() else None},
                f, indent=1)
  except OSError:
    pass

"""The driver-facing contract of bench.py: stdout is exactly ONE JSON line with the agreed keys, for the reference arm
(CPU, runs anywhere) and for the B200 arm (GPU)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BASE_KEYS = {'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
             'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'gpu_launches'}


def run_bench(args, env=None, timeout=900):
  out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, capture_output=True, text=True,
                       timeout=timeout, env=dict(os.environ, **(env or {})), cwd=ROOT)
  assert out.returncode == 0, out.stderr[-2000:]
  lines = [l for l in out.stdout.split('\n') if l.strip()]
  assert len(lines) == 1, out.stdout[:2000]   # nothing but the JSON line on stdout
  return json.loads(lines[0])


def test_reference_arm_prints_one_json_line():
  line = run_bench(['--impl', 'reference', '--gpus', '1', '--steps', '2', '--warmup', '1'],
                   env={'VTC_BENCH_CPU_SAMPLE': '256'})
  assert BASE_KEYS <= set(line) and line['impl'] == 'reference'
  assert line['metric'] == 'fista_patches_per_sec' and line['unit'] == 'patches/s' and line['higher_is_better'] is True
  assert line['steps'] == 2 and line['warmup'] == 1 and line['n_gpus'] == 1 and line['vs_baseline'] is None
  assert line['value'] > 0 and line['ms_per_step'] > 0 and 'workload' in line['config']
  assert set(line['cpu_baseline']) >= {'value', 'unit', 'cores', 'kind', 'sample'}
  # the unmodified reference when it is staged (oracle/_ref, tools/stage_reference.py), else the oracle port
  from oracle import reference
  assert line['cpu_baseline']['kind'] == ('reference' if reference.available() else 'port')
  assert line['e2e'] == {'value': line['value'], 'unit': 'patches/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
  # both arms describe the SAME workload in `config` (nothing implementation-specific in it)
  sys.path.insert(0, ROOT)
  import bench
  assert line['config'] == bench.shared_config(bench.B_PER_GPU)


def test_reference_arm_other_ranks_print_nothing():
  out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps',
                        '1', '--warmup', '0'], capture_output=True, text=True, timeout=300, cwd=ROOT,
                       env=dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1', VTC_BENCH_CPU_SAMPLE='256'))
  assert out.returncode == 0 and out.stdout.strip() == ''


@pytest.mark.gpu
def test_b200_arm_prints_one_json_line():
  line = run_bench(['--steps', '2', '--warmup', '3', '--batch', '8192', '--no-extras'])
  assert BASE_KEYS <= set(line) and 'impl' not in line or line.get('impl') == 'b200'
  assert line['metric'] == 'fista_patches_per_sec' and line['n_gpus'] == 1 and line['steps'] == 2
  assert line['gpu_launches'] > 0 and line['value'] > 0
  assert set(line['roofline']) >= {'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'}
  assert 0 < line['roofline']['frac'] < 1.2
  assert set(line['e2e']) >= {'value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step'}
  assert line['e2e']['h2d_bytes_per_step'] == 8192 * 256 * 4 and line['e2e']['d2h_bytes_per_step'] == 8192 * 1024 * 4
  assert set(line['clocks']) >= {'sm_mhz', 'sm_max_mhz', 'reasons'}
  assert line['roofline']['bound'] == 'tensor' and set(line['roofline']['hbm']) >= {'achieved', 'peak', 'frac'}
  # the library's own events of the timed steps add up to the driver-timed step (10.7 % was unaccounted)
  br = line['roofline']['step_breakdown']
  assert abs(br['unaccounted_frac']) < 0.05, br
  sys.path.insert(0, ROOT)
  import bench
  assert line['config'] == bench.shared_config(8192)

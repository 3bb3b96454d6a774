"""
bench.py -- the reference's headline metric on B200: FISTA patches/sec (300 iterations) on BASELINE.json configs[1]
(16x16 whitened patches, D=256, 1024 atoms, batch 65,536 per GPU, lambda 0.1), plus train steps/sec on configs[2].

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision bf16x3|bf16|bf16x6]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full inference call (step size, Gram/drive GEMMs, 300 fused iterations) on one batch per GPU.
Inference needs no communication, so N GPUs run N independent shards of an N-times larger batch (weak scaling);
`value` is the whole-job patches/sec with inputs resident in HBM, timed with CUDA events on the launching stream
between barriers, max over ranks. `e2e` is the same call made through the public drop-in API from pinned HOST
buffers, host<->device copies of every step inside the timed region (vision_transform_codes_b200.host_pipeline keeps two
steps in flight so that the copies of one step run under the compute of its neighbours). `--impl reference` times the
UNMODIFIED reference (staged under oracle/_ref by tools/stage_reference.py; kind "reference") on all host cores, each
step a bounded sample of the same workload; without the staged copy it falls back to the oracle port (kind "port").
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

import torch  # noqa: E402

B_PER_GPU, S, D, T, LAM = 65536, 1024, 256, 300, 0.1
TRAIN_GLOBAL_BATCH = 524288
WORKLOAD = ('configs[1]: fully-connected FISTA, 16x16 whitened patches (D=256), 1024 atoms, '
            'batch 65536 per GPU, 300 iters, lambda 0.1')
CPU_SAMPLE = int(os.environ.get('VTC_BENCH_CPU_SAMPLE', '32768'))  # patches per CPU step (the contract test shrinks it)


def shared_config(batch_per_gpu):
  """The `config` of BOTH arms (this one and --impl reference): the workload only, nothing implementation-specific."""
  return {'workload': WORKLOAD, 'batch_per_gpu': batch_per_gpu, 'atoms': S, 'pixels': D, 'iters': T,
          'sparsity_weight': LAM, 'variant': 'fista', 'patches': 'synthetic whitened (oracle.synthetic_patches, seed = rank)',
          'dictionary': 'unit-norm Gaussian rows (oracle.synthetic_dictionary, seed 1)'}
CONV_IMAGES_PER_GPU, CONV_LAM = 128, 0.05


def peaks():
  path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(path):
    p = json.load(open(path))
    return {'bf16_sustained': p.get('bf16_tflops_sustained'), 'bf16_burst': p.get('bf16_tflops'),
            'hbm_gbs': p.get('hbm_gbs'), 'source': 'measured'}
  return {'bf16_sustained': 1400.0, 'bf16_burst': 1590.0, 'hbm_gbs': 6650.0, 'source': 'fallback'}


class ClockSampler:
  """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
  FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
            'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
            'clocks_event_reasons.sw_power_cap')

  def __init__(self, index):
    self.index, self.rows, self.proc, self.thread = index, [], None, None

  def start(self):
    try:
      self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.FIELDS,
                                    '--format=csv,noheader,nounits', '-lms', '100'],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except OSError:
      return
    self.thread = threading.Thread(target=self._read, daemon=True)
    self.thread.start()

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append([c.strip() for c in line.split(',')])

  def stop(self):
    if self.proc is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    self.proc.terminate()
    self.thread.join(timeout=2)
    sm, power, smax, reasons = [], [], None, set()
    for r in self.rows:
      try:
        sm.append(float(r[0]))
        smax = float(r[1])
      except (ValueError, IndexError):
        continue
      try:
        power.append(float(r[2]))
      except (ValueError, IndexError):
        power.append(0.0)
      for name, cell in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
        if cell.lower().startswith('active'):
          reasons.add(name)
    # "under load" = the samples drawing at least 60 % of the highest power seen: an idle GPU of this pool sits at its
    # maximum clock, and a short timed region under torchrun leaves many idle samples around it
    top = max(power) if power else 0.0
    busy = [v for v, w in zip(sm, power) if top > 0 and w >= 0.6 * top] or sm
    return {'sm_mhz': statistics.median(busy) if busy else None, 'sm_max_mhz': smax, 'reasons': sorted(reasons),
            'samples': len(sm)}


def cpu_reference_kind():
  from oracle import reference
  return 'reference' if reference.available() else 'port'


def cpu_reference_patches_per_sec(sample, repeats=1):
  """The reference's own float32 CPU implementation (analysis_transforms/fully_connected/ista_fista.py:14-148, staged
  unmodified under oracle/_ref) on `sample` patches of the same workload, all host cores; the oracle port only when
  the staged copy is absent. Returns (patches/s, cores, seconds, kind)."""
  from oracle import reference
  from oracle import vtc_oracle as oracle
  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  phi = oracle.synthetic_dictionary(S, D)
  x = oracle.synthetic_patches(sample, D, kind='whitened')
  kind = cpu_reference_kind()

  def timed_runs(run):
    run(x[:64], phi, LAM, 3)  # warm the thread pool
    best = float('inf')
    for _ in range(repeats):
      t0 = time.perf_counter()
      run(x, phi, LAM, T)
      best = min(best, time.perf_counter() - t0)
    return best

  if kind == 'reference':
    with reference.reference_only():
      best = timed_runs(reference.load('analysis_transforms.fully_connected.ista_fista').run)
  else:
    best = timed_runs(oracle.ista_fista)
  return sample / best, cores, best, kind


def conv_benchmark(world, rank, dev, timed, pk, precision, images_per_gpu=CONV_IMAGES_PER_GPU):
  """BASELINE.json configs[4]: convolutional FISTA, 64 filters of 16x16 at stride 8 on 512x512 whitened images,
  sharded by image (no data-path collective), plus one conv dictionary update on the same batch."""
  import vision_transform_codes_b200 as pkg
  from oracle import vtc_oracle as oracle  # seeded input generators only
  from vision_transform_codes_b200.analysis_transforms.convolutional import ista_fista as conv_ista_fista
  from vision_transform_codes_b200.dict_update_rules.convolutional import sc_cheap_quadratic_descent as conv_update
  from vision_transform_codes_b200.dict_update_rules.convolutional import _common as conv_common
  nk, k, st, side = 64, (16, 16), (8, 8), 512
  few, pad = oracle.synthetic_padded_images(8, 1, side, side, k, st, seed=100 + rank)
  x = few.repeat((images_per_gpu + 7) // 8, 1, 1, 1)[:images_per_gpu].contiguous().to(dev)
  phi = oracle.synthetic_conv_dictionary(nk, 1, k[0], k[1]).to(dev)
  saved = pkg.config.precision
  pkg.config.precision = precision
  ms, launches = timed(lambda: conv_ista_fista.run(x, phi, st, pad, CONV_LAM, T), 2, 1)
  ms /= 2
  codes = conv_ista_fista.run(x, phi, st, pad, CONV_LAM, T)
  h = torch.zeros(nk, device=dev)

  def update():
    conv_common.hessian_running_mean(h, codes)
    conv_update.run(x, phi.clone(), codes, h, st, pad, stepsize=0.001)

  ums, _ = timed(update, 2, 1)
  ums /= 2
  pkg.config.precision = saved
  sh = (x.shape[2] - k[0]) // st[0] + 1
  rows = images_per_gpu * (x.shape[2] // st[0]) * (x.shape[3] // st[1])
  nparts = {'bf16': 1, 'bf16x3': 2, 'bf16x6': 3}[precision]
  # per grid row and iteration (64 code channels / 64 pixels per block): state 12 B x 64, image 4 B x 64, y and r parts
  # written once and read once each (2 P B x 64 x 2 x 2)
  bytes_iter = rows * 64 * (12 + 4 + 8 * nparts)
  flops_iter = 4.0 * images_per_gpu * nk * 256 * sh * sh
  return {
      'workload': 'configs[4]: convolutional FISTA, 64 filters 16x16, stride 8, %d whitened 512x512 images per GPU '
                  '(padded 528x528, codes 64x65x65), %d iters, lambda %g, sharded by image' % (images_per_gpu, T, CONV_LAM),
      'value': world * images_per_gpu / (ms * 1e-3), 'unit': 'images/s', 'ms_per_step': ms, 'precision': precision,
      'ms_per_iteration': ms / T, 'launches_per_step': int(launches // 2),
      'algorithmic_bytes_per_iteration': bytes_iter,
      'hbm_gbs_achieved': bytes_iter / (ms / T * 1e-3) / 1e9, 'hbm_frac': bytes_iter / (ms / T * 1e-3) / 1e9 / pk['hbm_gbs'],
      'algorithmic_tflops': flops_iter / (ms / T * 1e-3) / 1e12,
      'dictionary_update_ms': ums,
  }


def subspace_benchmark(world, rank, dev, timed, pk, precision, batch=131072):
  """BASELINE.json configs[3]: subspace FISTA with group soft-thresholding (groups of 2), 32x32 patches (D = 1024),
  4096 atoms, batch 131,072 per GPU, 300 iterations. D > 256: the two-launch synthesis schedule, tensor-bound."""
  import numpy as np
  import vision_transform_codes_b200 as pkg
  from oracle import vtc_oracle as oracle  # seeded input generators only
  from vision_transform_codes_b200.analysis_transforms.fully_connected import subspace_ista_fista
  s4, d4 = 4096, 1024
  phi = oracle.synthetic_dictionary(s4, d4).to(dev)
  gx = torch.Generator(device=dev).manual_seed(200 + rank)
  x = 0.3 * torch.randn(batch, d4, generator=gx, device=dev)
  groups = [list(map(int, g)) for g in np.array_split(np.arange(s4), s4 // 2)]
  saved = pkg.config.precision
  pkg.config.precision = precision
  ms, launches = timed(lambda: subspace_ista_fista.run(x, phi, groups, LAM, T), 1, 1)
  pkg.config.precision = saved
  nprod = pkg.PRECISIONS[precision]
  executed = nprod * 4.0 * batch * s4 * d4 * T
  return {
      'workload': 'configs[3]: subspace FISTA, groups of 2, 32x32 patches (D=1024), 4096 atoms, batch %d per GPU, '
                  '%d iters' % (batch, T),
      'value': world * batch / (ms * 1e-3), 'unit': 'patches/s', 'ms_per_step': ms, 'precision': precision,
      'gpu_launches': int(launches),
      'executed_mma_tflops': executed / (ms * 1e-3) / 1e12,
      'executed_frac_of_tensor_peak': executed / (ms * 1e-3) / 1e12 / pk['bf16_sustained'],
      'tflops_gram_equivalent': 2.0 * batch * s4 * s4 * T / (ms * 1e-3) / 1e12,
  }


def config0_benchmark(rank, world, dev, timed, precision):
  """BASELINE.json configs[0], the reference's own CPU-runnable case, timed IN FULL on both sides: batch 250, 256 atoms,
  16x16 patches, 300 FISTA iterations, lambda 0.1. One call = one step (latency-bound on the GPU: 300 launches of the
  Gram form, S <= 2 D); the CPU side is the float32 torch port of the reference on all host cores (rank 0, N = 1 only)."""
  import vision_transform_codes_b200 as pkg
  from oracle import vtc_oracle as oracle
  from vision_transform_codes_b200.analysis_transforms.fully_connected import ista_fista
  phi = oracle.synthetic_dictionary(256, 256)
  x = oracle.synthetic_patches(250, 256, kind='whitened')
  phid, xd = phi.to(dev), x.to(dev)
  saved = pkg.config.precision
  pkg.config.precision = precision
  ms, launches = timed(lambda: ista_fista.run(xd, phid, LAM, T), 10, 3)
  pkg.config.precision = saved
  out = {'workload': 'configs[0]: batch 250, 256 atoms, D=256, 300 FISTA iterations', 'ms_per_call': ms / 10,
         'patches_per_sec': 250 / (ms / 10 * 1e-3), 'gpu_launches_per_call': int(launches // 10)}
  if rank == 0 and world == 1:
    from oracle import reference
    torch.set_num_threads(os.cpu_count() or 1)
    kind = cpu_reference_kind()

    def cpu_call(run):
      run(x, phi, LAM, 10)
      t0 = time.perf_counter()
      run(x, phi, LAM, T)
      return (time.perf_counter() - t0) * 1e3

    if kind == 'reference':
      with reference.reference_only():
        cpu_ms = cpu_call(reference.load('analysis_transforms.fully_connected.ista_fista').run)
    else:
      cpu_ms = cpu_call(oracle.ista_fista)
    out.update({'cpu_reference_ms_per_call': cpu_ms, 'cpu_cores': os.cpu_count() or 1, 'cpu_kind': kind,
                'speedup_over_cpu_reference': cpu_ms / (ms / 10)})
  return out


def metrics_benchmark(rank, world, dev, timed, codes, x, phi):
  """the trainer's validation metrics (training/sparse_coding.py:177-229) of one configs[1] batch, computed
  on the device from resident codes (one residual contraction + one pass over residuals, codes and pixels; 8 doubles
  and 1024 floats cross to the host), next to the reference's host-side definition (the oracle's compute_metrics: numpy
  on copies of images, reconstructions and norms, a Python loop over the batch for the pSNR) on a sample of the batch."""
  from oracle import vtc_oracle as oracle
  from vision_transform_codes_b200.lean import metrics
  prev = phi.clone()
  ms, launches = timed(lambda: metrics.compute_metrics(x, codes, phi, prev, LAM), 5, 2)
  out = {'workload': 'validation metrics of one configs[1] batch (%d patches, %d atoms)' % tuple(codes.shape),
         'ms_per_call': ms / 5, 'gpu_launches_per_call': int(launches // 5),
         'd2h_bytes_per_call': 8 * 8 + 4 * codes.shape[1]}
  if rank == 0 and world == 1:
    n = 8192
    xs, cs, ps = x[:n].cpu(), codes[:n].cpu(), phi.cpu()
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    want = oracle.compute_metrics(xs, cs, ps, ps, LAM)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    got = metrics.compute_metrics(x[:n], codes[:n], phi, prev, LAM)
    err = max(abs(float(got[k]) - float(want[k])) / max(1.0, abs(float(want[k])))
              for k in want if k != metrics.CHANGE)
    out.update({'cpu_reference_ms_per_call_extrapolated': cpu_ms * codes.shape[0] / n,
                'cpu_sample': '%d of %d patches, scaled linearly' % (n, codes.shape[0]),
                'max_rel_difference_on_sample': err})
  return out


def run_reference(args, rank):
  if rank != 0:
    return
  steps = max(1, args.steps)
  # every step is a bounded sample of the workload (~27 s at 32 768 patches on 16 cores); with many steps the sample
  # shrinks so that the whole run still ends within a few minutes (throughput is linear in the batch)
  sample = CPU_SAMPLE if steps <= 6 else max(4096, (CPU_SAMPLE * 6 // steps) // 256 * 256)
  for _ in range(max(0, min(args.warmup, 1))):
    cpu_reference_patches_per_sec(256)
  vals, secs, cores, kind = [], [], 1, 'port'
  for _ in range(steps):
    v, cores, s, kind = cpu_reference_patches_per_sec(sample)
    vals.append(v)
    secs.append(s)
  value = sample * len(vals) / sum(secs)
  line = {
      'impl': 'reference', 'metric': 'fista_patches_per_sec', 'value': value, 'unit': 'patches/s',
      'n_gpus': args.gpus, 'steps': steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * sum(secs) / len(secs),
      'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'e2e': {'value': value, 'unit': 'patches/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
      'gpu_launches': 0,
      'cpu_baseline': {'value': value, 'unit': 'patches/s', 'cores': cores, 'kind': kind,
                       'sample': '%d of %d patches x %d iterations per step; throughput is linear in the batch '
                                 '(every patch is an independent problem)' % (sample, args.batch, T),
                       'what': ('the unmodified reference, analysis_transforms/fully_connected/ista_fista.py run() '
                                'staged under oracle/_ref, float32 torch on all host threads') if kind == 'reference'
                               else 'oracle port of the reference (oracle/_ref not staged), float32 torch on all host threads'},
      'config': shared_config(args.batch),
  }
  emit(line)


RESULT_STREAM = None


def claim_stdout():
  """stdout carries the ONE JSON line and nothing else: keep a private handle on it and point file descriptor 1 at
  stderr, so that whatever a library writes to stdout (NCCL prints its version banner there when NCCL_DEBUG=VERSION,
  and ignores NCCL_DEBUG_FILE at that level) ends up on stderr."""
  global RESULT_STREAM
  if RESULT_STREAM is None:
    sys.stdout.flush()
    RESULT_STREAM = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)


def emit(line):
  RESULT_STREAM.write(json.dumps(line) + '\n')
  RESULT_STREAM.flush()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=5)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
  ap.add_argument('--precision', default=os.environ.get('VTC_B200_PRECISION', 'bf16x3'))
  ap.add_argument('--batch', type=int, default=B_PER_GPU, help='patches per GPU (default: the BASELINE config)')
  ap.add_argument('--no-extras', action='store_true', help='skip the bf16-path, train-step and CPU side measurements')
  args = ap.parse_args()
  claim_stdout()

  rank = int(os.environ.get('RANK', '0'))
  world = int(os.environ.get('WORLD_SIZE', '1'))
  local_rank = int(os.environ.get('LOCAL_RANK', '0'))
  if args.impl == 'reference':
    run_reference(args, rank)
    return

  import torch.distributed as dist
  import vision_transform_codes_b200 as pkg
  from oracle import vtc_oracle as oracle  # inputs only (seeded generators shared with the tests) + cpu_baseline
  from vision_transform_codes_b200 import _lib
  from vision_transform_codes_b200.analysis_transforms.fully_connected import ista_fista

  if not torch.cuda.is_available():
    raise SystemExit('bench.py needs a B200: the CUDA path has no CPU fallback')
  dev = torch.device('cuda', local_rank)
  torch.cuda.set_device(dev)
  if world > 1:
    # NCCL writes its version banner / debug lines to stdout by default; stdout carries the one JSON line only
    os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
    dist.init_process_group('nccl', device_id=dev)
  lib = _lib.load()
  pkg.config.precision = args.precision
  pkg.config.check_finite = False  # no host synchronisation inside the timed region of `value`
  Bn = args.batch

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def max_over_ranks(ms):
    if world == 1:
      return ms
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

  phi = oracle.synthetic_dictionary(S, D).to(dev)
  x_host = oracle.synthetic_patches(Bn, D, seed=rank, kind='whitened').pin_memory()
  x = x_host.to(dev)

  def timed(fn, steps, warmup):
    for _ in range(warmup):
      fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = lib.vtc_launch_count()
    e0.record()
    for _ in range(steps):
      fn()
    e1.record()
    barrier()
    return max_over_ranks(e0.elapsed_time(e1)), lib.vtc_launch_count() - n0

  # ---- headline: inputs resident in HBM. The library records its own CUDA events (setup / iteration launches /
  #      finishing copy of EVERY call, on the launching stream) inside this same timed loop, so the per-kernel figures
  #      below are of the timed region itself, not of a separate call made afterwards.
  import ctypes
  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()
  for _ in range(args.warmup):
    ista_fista.run(x, phi, LAM, T)
  barrier()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  n0 = lib.vtc_launch_count()
  lib.vtc_profile_enable(1)
  e0.record()
  for _ in range(args.steps):
    ista_fista.run(x, phi, LAM, T)
  e1.record()
  barrier()
  total_ms, launches = max_over_ranks(e0.elapsed_time(e1)), lib.vtc_launch_count() - n0
  clocks = sampler.stop() if rank == 0 else None
  ms_per_step = total_ms / args.steps
  value = world * Bn / (ms_per_step * 1e-3)
  per_step = []
  f = [ctypes.c_float() for _ in range(4)]
  for i in range(lib.vtc_profile_history_count()):
    _lib.check(lib.vtc_profile_history(i, *[ctypes.byref(v) for v in f]))
    per_step.append({'setup_ms': f[0].value, 'iter_ms': f[1].value, 'finish_ms': f[2].value, 'gap_to_next_ms': f[3].value})
  setup_ms, iter_ms, n_launch, n_iter, fused_ms, first_ms = (ctypes.c_float(), ctypes.c_float(), ctypes.c_int(),
                                                             ctypes.c_int(), ctypes.c_float(), ctypes.c_float())
  _lib.check(lib.vtc_profile_last(ctypes.byref(setup_ms), ctypes.byref(iter_ms), ctypes.byref(n_launch),
                                  ctypes.byref(n_iter), ctypes.byref(fused_ms), ctypes.byref(first_ms)))
  lib.vtc_profile_enable(0)
  n_launch, n_iter, first_ms = n_launch.value, n_iter.value, first_ms.value
  mean = lambda key: sum(r[key] for r in per_step) / max(1, len(per_step))  # noqa: E731
  setup_ms, iter_ms, finish_ms = mean('setup_ms'), mean('iter_ms'), mean('finish_ms')
  gap_ms = sum(r['gap_to_next_ms'] for r in per_step) / max(1, len(per_step) - 1)
  accounted_ms = setup_ms + iter_ms + finish_ms + gap_ms

  pk = peaks()
  nprod = pkg.PRECISIONS[args.precision]
  nparts = {1: 1, 3: 2, 6: 3}[nprod]
  form = {1: 'gram', 2: 'synthesis'}[lib.vtc_get_formulation(S, D)]
  fused_iter = form == 'synthesis' and bool(lib.vtc_get_fused_iteration(S, D, nprod))
  persistent = fused_iter and n_launch == 1
  gram_flops_iter = 2.0 * Bn * S * S       # north-star (Gram form) algorithmic flops of one iteration
  synth_flops_iter = 4.0 * Bn * S * D      # the reference's own two-contraction form (what is executed, x products)
  launch_iters = n_iter if persistent else 1
  launch_ms = iter_ms if persistent else fused_ms.value   # the dominant kernel's launch: all iterations, or one
  if form == 'gram':
    bytes_iter = Bn * S * (16 + 4 * nparts)
  elif fused_iter:
    # per iteration: reads a_{k-1}, a_{k-2} (8 B) and writes a_k (4 B) per code element; per pixel reads x (4 B), reads
    # r_{k-1} parts (2P B) and writes r_k parts (2P B). y_k stays on chip.
    bytes_iter = Bn * S * 12 + Bn * D * (4 + 4 * nparts)
  else:
    bytes_iter = Bn * S * (12 + 2 * nparts) + Bn * D * 2 * nparts
  executed_flops_iter = nprod * (gram_flops_iter if form == 'gram' else synth_flops_iter)
  # primary roofline, SURVEY 8(d): Gram-form algorithmic flops of the launch / its duration, against the measured
  # sustained cuBLAS bf16 throughput (the kernel is timed inside a long step)
  achieved = launch_iters * gram_flops_iter / (launch_ms * 1e-3) / 1e12
  traffic = None  # DRAM bytes per launch of this kernel from the committed ncu --set full capture of the same shape
  tpath = os.path.join(ROOT, 'profiles', 'r02_traffic.json')
  if os.path.exists(tpath) and form == 'synthesis' and Bn == B_PER_GPU:
    per_iter = json.load(open(tpath)).get(args.precision, {}).get('iter_bytes')
    traffic = per_iter * launch_iters if per_iter else None
  hbm_gbs = launch_iters * bytes_iter / (launch_ms * 1e-3) / 1e9
  roofline = {
      'bound': 'tensor', 'achieved': achieved, 'peak': pk['bf16_sustained'], 'unit': 'TFLOP/s',
      'frac': achieved / pk['bf16_sustained'], 'peak_source': pk['source'] + ' bf16_tflops_sustained (cuBLAS)',
      'basis': 'Gram-form algorithmic flops 2 B S^2 per iteration (SURVEY 8d) x %d iterations per launch' % launch_iters,
      'traffic': traffic,
      'kernel': ('%s<%d> (panel-resident, y_k on chip; %s)' %
                 ('vtc_fista_iter_kernel' if os.environ.get('VTC_B200_ITER_GEN') == '1' else 'vtc_fista_iter2_kernel',
                  nparts, 'ONE launch = all %d iterations' % n_iter if persistent else 'one launch per iteration'))
                if fused_iter else 'vtc_gemm_kernel<EPI_FISTA,%d> (%s form)' % (nparts, form),
      'launch_ms': launch_ms, 'iterations_per_launch': launch_iters, 'ms_per_iteration': launch_ms / launch_iters,
      'frac_of_burst_peak': achieved / pk['bf16_burst'] if pk['bf16_burst'] else None,
      'executed_mma': {'tflops': launch_iters * executed_flops_iter / (launch_ms * 1e-3) / 1e12,
                       'frac': launch_iters * executed_flops_iter / (launch_ms * 1e-3) / 1e12 / pk['bf16_sustained'],
                       'products_per_fp32_product': nprod, 'form': form,
                       'note': 'the %s form executes %.0f%% of the Gram-form flops, times %d bf16 products' %
                               (form, 100 * (1.0 if form == 'gram' else synth_flops_iter / gram_flops_iter), nprod)},
      'hbm': {'achieved': hbm_gbs, 'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': hbm_gbs / pk['hbm_gbs'],
              'algorithmic_bytes_per_iteration': bytes_iter, 'algorithmic_bytes_per_launch': launch_iters * bytes_iter,
              'floor_ms_per_iteration': bytes_iter / (pk['hbm_gbs'] * 1e9) * 1e3},
      'tensor_floor_ms_per_iteration_executed': executed_flops_iter / (pk['bf16_sustained'] * 1e12) * 1e3,
      'first_launch_ms': first_ms if (form == 'synthesis' and not fused_iter) else None,
      # where a timed step goes (means over the %d timed steps, library events on the launching stream)
      'step_breakdown': {'setup_ms': setup_ms, 'iter_ms': iter_ms, 'finish_ms': finish_ms, 'gap_to_next_ms': gap_ms,
                         'sum_ms': accounted_ms, 'ms_per_step': ms_per_step,
                         'unaccounted_frac': (ms_per_step - accounted_ms) / ms_per_step,
                         'per_step': per_step[:8]},
  }

  # ---- end to end through the public API from pinned host memory: every step uploads its images and downloads its
  #      codes; two steps are in flight (host_pipeline), so the copies of step i run under the compute of i - 1 / i + 1
  from vision_transform_codes_b200.host_pipeline import HostPipeline
  pkg.config.check_finite = True  # the default user-facing behaviour
  depth = 2
  codes_host = [torch.empty((Bn, S), dtype=torch.float32).pin_memory() for _ in range(depth)]
  pipe = HostPipeline(dev, depth=depth)
  e2e_steps = max(4, min(args.steps, 8))

  def e2e_run(n):
    for i in range(n):
      pipe.submit(x_host, phi, LAM, T, out=codes_host[i % depth])
    pipe.synchronize()

  e2e_run(2)
  barrier()
  s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  s0.record(pipe.upload)       # the first upload starts after this ...
  e2e_run(e2e_steps)
  s1.record(pipe.download)     # ... and the last download has finished before this
  barrier()
  e2e_ms = max_over_ranks(s0.elapsed_time(s1)) / e2e_steps
  e2e = {'value': world * Bn / (e2e_ms * 1e-3), 'unit': 'patches/s', 'ms_per_step': e2e_ms, 'steps': e2e_steps,
         'h2d_bytes_per_step': x_host.numel() * 4, 'd2h_bytes_per_step': codes_host[0].numel() * 4,
         'frac_of_device_resident_value': (world * Bn / (e2e_ms * 1e-3)) / value,
         'how': 'HostPipeline depth 2: per step one pinned H2D copy of the images, one ista_fista.run, one D2H copy of '
                'the dense fp32 codes; CUDA events from before the first upload to after the last download, max over ranks'}
  # the same without overlap (one step at a time, synchronised): what a naive caller gets
  def serial_step():
    xd = x_host.to(dev, non_blocking=True)
    codes = ista_fista.run(xd, phi, LAM, T)
    codes_host[0].copy_(codes, non_blocking=True)
    torch.cuda.current_stream().synchronize()

  ser_ms, _ = timed(serial_step, 2, 1)
  e2e['serial_ms_per_step'] = ser_ms / 2
  pkg.config.check_finite = False
  del pipe

  line = {
      'metric': 'fista_patches_per_sec', 'value': value, 'unit': 'patches/s', 'n_gpus': world, 'steps': args.steps,
      'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': None, 'dtype': args.precision + ' (bf16 products, fp32 accumulate)', 'data': 'synthetic',
      'train_step': None,   # (filled below; early in the line so that it survives truncated logs)
      'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roofline, 'clocks': clocks,
      'cpu_baseline': None, 'config': shared_config(Bn),
      'implementation': {'precision': args.precision, 'formulation': form,
                         'schedule': ('one persistent launch for all %d iterations' % n_iter if persistent else
                                      'one launch per iteration') if fused_iter else
                                     '%d launch(es) per iteration' % (n_launch // max(1, n_iter)),
                         'parallelism': 'batch sharded over %d GPU(s), dictionary replicated, no data-path '
                                        'collective' % world,
                         'l2': 'inputs larger than L2 (per-iteration state %.0f MB vs 126 MB L2)' % (Bn * S * 4 * 3 / 1e6)},
  }

  if not args.no_extras:
    # the collective-bearing number first (it must survive truncated logs): configs[2], one full train step
    from vision_transform_codes_b200.lean import sparse_coding as trainer
    line['train_step'] = trainer.benchmark_train_step(TRAIN_GLOBAL_BATCH, S, D, T, LAM, world, rank, dev, timed)
    # separately toleranced plain-bf16 path (1 MMA pass per product)
    if args.precision != 'bf16':
      pkg.config.precision = 'bf16'
      ms, _ = timed(lambda: ista_fista.run(x, phi, LAM, T), 3, 2)
      ms /= 3
      line['bf16_path'] = {'value': world * Bn / (ms * 1e-3), 'unit': 'patches/s', 'ms_per_step': ms,
                           'tflops_gram_equivalent_whole_call': (T * gram_flops_iter) / (ms * 1e-3) / 1e12,
                           'frac_of_tensor_peak_gram_equivalent':
                               (T * gram_flops_iter) / (ms * 1e-3) / 1e12 / pk['bf16_sustained'],
                           'tolerance': 'codes rel-L2 <= 1.4e-2, reconstructions <= 4e-3 (tests/test_gpu_parity.py)'}
      if fused_iter:
        # the plain-bf16 iteration kernel is HBM-bound: fp32 state 12 B per code element + x, r parts per pixel; the
        # whole call (setup included) against the measured copy bandwidth
        bytes_bf16 = Bn * S * 12 + Bn * D * (4 + 4 * 1)
        gbs = T * bytes_bf16 / (ms * 1e-3) / 1e9
        line['bf16_path']['roofline'] = {'bound': 'hbm', 'achieved': gbs, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                                         'frac': gbs / pk['hbm_gbs'], 'algorithmic_bytes_per_iteration': bytes_bf16,
                                         'basis': 'whole call / %d iterations' % T}
      pkg.config.precision = args.precision
    line['conv_path'] = conv_benchmark(world, rank, dev, timed, pk, args.precision)
    line['subspace_path'] = subspace_benchmark(world, rank, dev, timed, pk, args.precision)
    line['config0'] = config0_benchmark(rank, world, dev, timed, args.precision)
    line['metrics_path'] = metrics_benchmark(rank, world, dev, timed, ista_fista.run(x, phi, LAM, T), x, phi)
    if rank == 0 and world == 1:
      v, cores, secs, kind = cpu_reference_patches_per_sec(CPU_SAMPLE)
      line['cpu_baseline'] = {'value': v, 'unit': 'patches/s', 'cores': cores, 'kind': kind,
                              'sample': '%d of %d patches x %d iterations (%.1f s), float32 torch on the host'
                              % (CPU_SAMPLE, Bn, T, secs)}
  if rank == 0:
    emit(line)
  if world > 1:
    dist.destroy_process_group()


if __name__ == '__main__':
  main()

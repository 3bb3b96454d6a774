"""
Sparse-coding dictionary training on B200: the per-batch hot path of the reference trainer, optionally data parallel.

The reference's ``training/sparse_coding.py`` (train_dictionary, :9-519) is the CALLER of the hot path and keeps
working unmodified on top of the drop-in modules (``vision_transform_codes_b200.install()``). This module is the
lean equivalent for when the reference tree is not on the machine, and the only trainer that is correct under data
parallelism: it mirrors the reference's parameter dictionary and per-batch body (:460-465 schedules, :513-515
infer -> update, :154 Hessian running mean, :170-175 pickle checkpoints, :498-506 validation metrics on the
'training_visualization_schedule', computed on the device by training/metrics.py) and leaves out what is not on the
hot path: the matplotlib dictionary figures (:237-271) are skipped, dictionary-element reset/prune (:522-764) raises
NotImplementedError.

Data parallel (``vision_transform_codes_b200.enable_data_parallel()``): every rank holds a replica of the
dictionary and its contiguous shard of each batch. Inference needs no communication. Per dictionary-update
iteration there is exactly one collective: a sum all-reduce of [gradient sum | sum of squared codes] (S*D + S
floats), after which every rank applies the identical update, so the replicas stay bit-identical.
"""
import pickle
import time

import torch

from vision_transform_codes_b200 import _lib, config
from vision_transform_codes_b200.analysis_transforms.fully_connected import ista_fista, subspace_ista_fista
from vision_transform_codes_b200.analysis_transforms.convolutional import ista_fista as conv_ista_fista
from vision_transform_codes_b200.dict_update_rules.convolutional import _common as _conv_common
from vision_transform_codes_b200.dict_update_rules.fully_connected import _common
from vision_transform_codes_b200.lean import metrics as _metrics

CHEAP_QUADRATIC = ('sc_cheap_quadratic_descent', 'subspace_sc_cheap_quadratic_descent')
UPDATE_RULES = ('sc_steepest_descent', 'sc_cheap_quadratic_descent', 'subspace_sc_steepest_descent',
                'subspace_sc_cheap_quadratic_descent')


def infer_codes(images, dictionary, code_inf_alg, sparsity_weight, num_iters, nonnegative_only=False,
                hard_threshold=False, group_assignments=None, kernel_strides=None, image_padding=None):
  """training/sparse_coding.py:124-140 (kernel_strides given: the convolutional branch, :132-134)."""
  if kernel_strides is not None:
    return conv_ista_fista.run(images, dictionary, kernel_strides, image_padding, sparsity_weight, num_iters,
                               variant=code_inf_alg, nonnegative_only=nonnegative_only,
                               hard_threshold=hard_threshold)
  if code_inf_alg in ('subspace_ista', 'subspace_fista'):
    return subspace_ista_fista.run(images, dictionary, group_assignments, sparsity_weight, num_iters,
                                   variant=code_inf_alg[9:], hard_threshold=hard_threshold)
  return ista_fista.run(images, dictionary, sparsity_weight, num_iters, variant=code_inf_alg,
                        nonnegative_only=nonnegative_only, hard_threshold=hard_threshold)


class _UpdateState:
  """Buffers reused across steps: packed [gradient sum | sum of squared codes] for the single all-reduce."""

  def __init__(self, dictionary):
    S, D = dictionary.shape
    self.packed = torch.empty(S * D + S, dtype=torch.float32, device=dictionary.device)
    self.grad = self.packed[:S * D].view(S, D)
    self.sq_sum = self.packed[S * D:]
    self.slots, self.width, self.reg = None, 0, None  # subspace alignment penalty


def allreduce_update_buffers(state, with_code_squares):
  """The one collective of a data-parallel update step: sum over ranks of [gradient sum | sum of squared codes]."""
  import torch.distributed as dist
  dist.all_reduce(state.packed if with_code_squares else state.grad, op=dist.ReduceOp.SUM,
                  group=config.process_group)


def update_dictionary(images, dictionary, codes, hessian_diag, stepsize, num_iters, state, lowest_code_val=0.001,
                      normalize_dictionary=True, batch_global=None, group_assignments=None, alignment_penalty=0.0):
  """
  training/sparse_coding.py:142-168 for the fully-connected rules: the Hessian running mean (:154, when
  hessian_diag is given) followed by num_iters descent steps, with ONE all-reduce per step when data parallel.
  """
  lib = _lib.load()
  _lib.require_cuda_f32(dictionary, 'dictionary')
  device = dictionary.device
  S, D = dictionary.shape
  B = codes.size(0)
  # the kernels read and write a dense row-major (S, D) array: a strided view (e.g. a transposed init_dictionary) is
  # updated through a contiguous working copy and written back, like dict_update_rules/fully_connected/_common.descend
  target = dictionary
  if not dictionary.is_contiguous():
    dictionary = dictionary.contiguous()
  if batch_global is None:
    # summed over the ranks EVERY step: a rank cannot tell from its own shard size that another rank's shard changed
    # (ragged last batch), and a stale divisor or a one-sided extra collective would desynchronise the replicas
    batch_global = _common.global_batch_size(B, device)
  codes_rm, ld_codes = _lib.row_major(codes)
  with torch.cuda.device(device):
    st = _lib.stream_ptr(device)
    for it in range(int(num_iters)):
      _common.dictionary_gradient(images, dictionary, codes, out=state.grad)
      first = it == 0 and hessian_diag is not None
      if first:
        _lib.check(lib.vtc_hessian_diag_update(_lib.ptr(codes_rm), ld_codes, B, S, batch_global,
                                               _lib.ptr(state.sq_sum), 0, 0, st))
      if config.data_parallel:
        allreduce_update_buffers(state, first)
      if first:
        _lib.check(lib.vtc_hessian_ema(_lib.ptr(hessian_diag), _lib.ptr(state.sq_sum), S, batch_global, st))
      reg = None
      if alignment_penalty != 0:
        if state.slots is None:
          state.slots, state.width = _common.group_slot_table(group_assignments, S, device)
          state.reg = torch.empty((S, D), dtype=torch.float32, device=device)
        reg = state.reg
        _lib.check(lib.vtc_subspace_alignment_grad(_lib.ptr(dictionary), S, D, _lib.ptr(state.slots),
                                                   state.slots.size(0), state.width, int(bool(normalize_dictionary)),
                                                   _lib.ptr(reg), st))
      _lib.check(lib.vtc_sc_dict_apply(_lib.ptr(dictionary), _lib.ptr(state.grad), _lib.ptr(hessian_diag),
                                       _lib.ptr(reg), float(alignment_penalty), S, D, int(batch_global),
                                       float(stepsize), float(lowest_code_val), int(bool(normalize_dictionary)), st))
  if dictionary is not target:
    target.copy_(dictionary)


def update_dictionary_convolutional(images_padded, dictionary, codes, hessian_diag, kernel_strides, image_padding,
                                    stepsize, num_iters, lowest_code_val=0.001, normalize_dictionary=True):
  """training/sparse_coding.py:142-168, convolutional branch: Hessian running mean (:158-161, when hessian_diag is
  given) then num_iters descent steps; data parallel: the squared-code sums and every gradient are all-reduced."""
  if hessian_diag is not None:
    _conv_common.hessian_running_mean(hessian_diag, codes)
  _conv_common.descend(images_padded, dictionary, codes, hessian_diag, kernel_strides, image_padding, stepsize,
                       num_iters, lowest_code_val, normalize_dictionary)


def load_newest_dictionary_checkpoint(checkpoint_dir):
  """The dictionary (ndarray) of the highest checkpoint iteration in a logging folder: files are raw pickles of the
  float32 ndarray named checkpoint_dictionary_iter_<k>, as the reference writes them (training/sparse_coding.py:
  170-175) and reads them back (utils/misc.py:9-21)."""
  import os
  iter_nums = [int(f[27:]) for f in os.listdir(checkpoint_dir) if f[:27] == 'checkpoint_dictionary_iter_']
  newest = max(iter_nums)
  with open(os.path.join(str(checkpoint_dir), 'checkpoint_dictionary_iter_' + str(newest)), 'rb') as f:
    return pickle.load(f)


def train_dictionary(training_image_dataset, validation_image_dataset, init_dictionary, all_params):
  """
  Train a sparse coding dictionary (fully-connected or convolutional mode), in place on ``init_dictionary``.

  Same arguments as the reference's train_dictionary (training/sparse_coding.py:9-117). Under data parallelism
  ``training_image_dataset`` yields THIS rank's shard of every batch.
  """
  assert 0 in all_params['inference_param_schedule']
  assert 0 in all_params['dict_update_param_schedule']
  coding_mode = all_params['mode']
  num_epochs = all_params['num_epochs']
  code_inf_alg = all_params['code_inference_algorithm']
  inf_param_schedule = all_params['inference_param_schedule']
  dict_update_alg = all_params['dictionary_update_algorithm']
  dict_update_param_schedule = all_params['dict_update_param_schedule']
  assert coding_mode in ['fully-connected', 'convolutional']
  assert code_inf_alg in ['ista', 'fista', 'subspace_ista', 'subspace_fista']
  assert dict_update_alg in UPDATE_RULES
  convolutional = coding_mode == 'convolutional'
  kernel_strides = image_padding = None
  if convolutional:
    # training/sparse_coding.py:295-299, :394-419: plain ISTA/FISTA and the two non-subspace update rules only
    kernel_strides = all_params['strides']
    image_padding = all_params['padding']
    if code_inf_alg not in ('ista', 'fista'):
      raise KeyError('Havent implemented subspace ISTA for convolutional yet')
    if dict_update_alg not in ('sc_steepest_descent', 'sc_cheap_quadratic_descent'):
      raise KeyError('Not implemented for convolutional')
  if 'dict_element_rp_schedule' in all_params:
    raise NotImplementedError('dict_element_rp_schedule is host-side orchestration outside the B200 hot path; run the '
                              'reference trainer on top of vision_transform_codes_b200.install() for it')
  nonneg_only = all_params.get('nonnegative_only', False)
  hard_threshold = all_params.get('hard_threshold', False)
  group_assignments = all_params.get('group_assignments')
  if group_assignments is not None:
    assert all([len(set(x)) == len(x) for x in group_assignments])
    group_assignments = [list(map(int, x)) for x in group_assignments]
  if code_inf_alg.startswith('subspace'):
    assert group_assignments is not None
  alignment_penalty = 0.0
  if dict_update_alg.startswith('subspace'):
    assert group_assignments is not None
    alignment_penalty = all_params['subspace_alignment_penalty']
  if all_params.get('renormalize_dictionary', True):
    norms = init_dictionary.flatten(1).norm(p=2, dim=1)
    assert torch.allclose(norms, torch.ones_like(norms)), 'Please ensure the initial dictionary is already normalized'
  ckpt_sched = all_params.get('checkpoint_schedule')
  logging_path = all_params.get('logging_folder_fullpath')
  if ckpt_sched is not None:
    assert logging_path is not None and not isinstance(logging_path, str), 'should be pathlib.Path'
    logging_path.mkdir(parents=True, exist_ok=True)
  print_interval = all_params.get('stdout_print_interval', 1000)
  # :355-366, :498-506: validation metrics at the scheduled iterations -> tensorboard scalars (when tensorboard is
  # importable) and, as (iteration, metrics) pairs, appended to all_params['validation_metrics_log'] if that is a list
  vis_sched = all_params.get('training_visualization_schedule')
  tb_writer, previous_dictionary = None, None
  metrics_log = all_params.get('validation_metrics_log')
  if vis_sched is not None:
    assert logging_path is not None and not isinstance(logging_path, str), 'should be pathlib.Path'
    previous_dictionary = init_dictionary.clone()

  dictionary = init_dictionary  # no copying, just a new reference
  hessian_diag = None
  if dict_update_alg in CHEAP_QUADRATIC:
    hessian_diag = init_dictionary.new_zeros(init_dictionary.shape[0])
  state = None if convolutional else _UpdateState(dictionary)
  rank0 = (not config.data_parallel) or torch.distributed.get_rank(config.process_group) == 0
  if vis_sched is not None and rank0:
    logging_path.mkdir(parents=True, exist_ok=True)
    try:
      from torch.utils.tensorboard import SummaryWriter
      tb_writer = SummaryWriter(logging_path)
    except ImportError:
      tb_writer = None

  starttime = time.time()
  total_iter_idx = 0
  for epoch_idx in range(num_epochs):
    for t_batch_images in training_image_dataset:
      if total_iter_idx % print_interval == 0 and total_iter_idx != 0 and rank0:
        print(total_iter_idx, 'iterations complete')
        print('Time elapsed:', '{:.1f}'.format(time.time() - starttime), 'seconds')
        print('-----')
      if total_iter_idx in inf_param_schedule:
        sparsity_weight = inf_param_schedule[total_iter_idx]['sparsity_weight']
        inf_num_iters = inf_param_schedule[total_iter_idx]['num_iters']
      if total_iter_idx in dict_update_param_schedule:
        d_upd_stp = dict_update_param_schedule[total_iter_idx]['stepsize']
        d_upd_niters = dict_update_param_schedule[total_iter_idx]['num_iters']
      if ckpt_sched is not None and total_iter_idx in ckpt_sched and rank0:
        pickle.dump(dictionary.cpu().numpy(),
                    open(logging_path / ('checkpoint_dictionary_iter_' + str(total_iter_idx)), 'wb'))
      if vis_sched is not None and total_iter_idx in vis_sched:
        val_metrics = []
        for v_batch_images in validation_image_dataset:
          if dictionary.device != v_batch_images.device:
            v_batch_images = v_batch_images.to(dictionary.device)
          v_codes = infer_codes(v_batch_images, dictionary, code_inf_alg, sparsity_weight, inf_num_iters, nonneg_only,
                                hard_threshold, group_assignments, kernel_strides, image_padding)
          val_metrics.append(_metrics.compute_metrics(
              v_batch_images, v_codes, dictionary, previous_dictionary, sparsity_weight, code_inf_alg,
              group_assignments, kernel_strides, image_padding))
        averaged = _metrics.average_metrics(val_metrics)
        if metrics_log is not None:
          metrics_log.append((total_iter_idx, averaged))
        if tb_writer is not None:
          for name in averaged:
            tb_writer.add_scalar(name, averaged[name], total_iter_idx)
      if dictionary.device != t_batch_images.device:
        t_batch_images = t_batch_images.to(dictionary.device)
      t_codes = infer_codes(t_batch_images, dictionary, code_inf_alg, sparsity_weight, inf_num_iters, nonneg_only,
                            hard_threshold, group_assignments, kernel_strides, image_padding)
      if previous_dictionary is not None:
        previous_dictionary.copy_(dictionary)
      if convolutional:
        update_dictionary_convolutional(t_batch_images, dictionary, t_codes, hessian_diag, kernel_strides,
                                        image_padding, d_upd_stp, d_upd_niters)
      else:
        update_dictionary(t_batch_images, dictionary, t_codes, hessian_diag, d_upd_stp, d_upd_niters, state,
                          group_assignments=group_assignments, alignment_penalty=alignment_penalty)
      total_iter_idx += 1
    if rank0:
      print("Epoch", epoch_idx + 1, "finished")
  return hessian_diag


def sharded_equivalence(S, D, num_iters, sparsity_weight, world, rank, device, reduced_batch=16384, steps=2):
  """
  SURVEY section 4(v) made visible to the driver: `steps` train steps on a reduced global batch, once sharded over the
  `world` ranks (dictionary-gradient all-reduce over NCCL) and once unsharded on rank 0 alone; returns the relative L2
  difference of the two dictionaries (rank 0; 0.0 by construction at world == 1). The shards are seeded per rank so
  that rank 0 can rebuild the whole batch.
  """
  shard = reduced_batch // world

  def shard_of(r):
    g = torch.Generator(device=device).manual_seed(500 + r)
    return 0.3 * torch.randn(shard, D, generator=g, device=device)

  def init():
    g = torch.Generator().manual_seed(1)
    phi = torch.randn(S, D, generator=g)
    return (phi / phi.norm(dim=1, keepdim=True)).to(device), torch.zeros(S, device=device)

  def steps_on(x, phi, h, batch_global):
    state = _UpdateState(phi)
    for _ in range(steps):
      codes = ista_fista.run(x, phi, sparsity_weight, num_iters)
      update_dictionary(x, phi, codes, h, 0.1, 1, state, batch_global=batch_global)

  was_parallel, was_group = config.data_parallel, config.process_group
  try:
    phi_sharded, h = init()
    config.data_parallel, config.process_group = world > 1, None
    steps_on(shard_of(rank), phi_sharded, h, shard * world)
    config.data_parallel = False
    if rank != 0:
      return None
    phi_single, h1 = init()
    steps_on(torch.cat([shard_of(r) for r in range(world)]), phi_single, h1, shard * world)
    return float(torch.norm(phi_sharded - phi_single) / torch.norm(phi_single))
  finally:
    config.data_parallel, config.process_group = was_parallel, was_group


def benchmark_train_step(global_batch, S, D, num_iters, sparsity_weight, world, rank, device, timed, steps=2):
  """
  bench.py helper: BASELINE.json configs[2], one full train step = FISTA inference on this rank's shard of a
  524,288-patch batch + Hessian running mean + cheap-quadratic dictionary update with one all-reduce.
  Strong scaling: the global batch is fixed and sharded over `world` GPUs.
  """
  g = torch.Generator().manual_seed(1)
  phi = torch.randn(S, D, generator=g)
  phi = (phi / phi.norm(dim=1, keepdim=True)).to(device)
  shard = global_batch // world
  gx = torch.Generator(device=device).manual_seed(100 + rank)
  x = 0.3 * torch.randn(shard, D, generator=gx, device=device)
  h = torch.zeros(S, device=device)
  state = _UpdateState(phi)
  was_parallel = config.data_parallel
  if world > 1:
    config.data_parallel, config.process_group = True, None

  def step():
    codes = ista_fista.run(x, phi, sparsity_weight, num_iters)
    update_dictionary(x, phi, codes, h, 0.1, 1, state, batch_global=global_batch)

  try:
    ms, launches = timed(step, steps, 1)
  finally:
    config.data_parallel = was_parallel
  ms /= steps
  replicas_identical = None
  if world > 1:
    import torch.distributed as dist
    ref = phi.clone()
    dist.broadcast(ref, src=0)
    same = torch.tensor([int(torch.equal(ref, phi))], device=device)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    replicas_identical = bool(same.item())
  equivalence = sharded_equivalence(S, D, num_iters, sparsity_weight, world, rank, device)
  return {'steps_per_sec': 1e3 / ms, 'ms_per_step': ms, 'global_batch': global_batch, 'shard_per_gpu': shard,
          'phi_rel_l2_vs_single_gpu': equivalence, 'replicas_bit_identical': replicas_identical,
          'scaling': 'strong', 'update_rule': 'sc_cheap_quadratic_descent', 'allreduce_bytes_per_step': 4 * (S * D + S),
          'patches_per_sec': global_batch * 1e3 / ms, 'gpu_launches': int(launches),
          'workload': 'configs[2]: FISTA-300 inference + SC quadratic-descent update, 16x16 patches, 1024 atoms'}

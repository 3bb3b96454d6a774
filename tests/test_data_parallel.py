"""
Host-side logic of the data-parallel train step on CPU with gloo, world_size 2 (no GPU needed): sharding the batch,
the single packed all-reduce of [gradient sum | sum of squared codes], and that applying the reduced buffers with the
global batch size reproduces the single-process update exactly as the reference defines it.
"""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from oracle import vtc_oracle as oracle


def _free_port():
  s = socket.socket()
  s.bind(('127.0.0.1', 0))
  port = s.getsockname()[1]
  s.close()
  return port


def _worker(rank, world, port, result_path):
  sys.path.insert(0, ROOT)
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  dist.init_process_group('gloo', rank=rank, world_size=world)
  import vision_transform_codes_b200 as pkg
  from vision_transform_codes_b200.lean import sparse_coding as trainer
  pkg.enable_data_parallel()
  torch.manual_seed(0)
  B, S, D = 96, 48, 24
  phi = oracle.synthetic_dictionary(S, D)
  x = oracle.synthetic_patches(B, D)
  codes = oracle.ista_fista(x, phi, 0.1, 20)
  h0 = torch.rand(S, generator=torch.Generator().manual_seed(7)) * 0.01
  shard = slice(rank * B // world, (rank + 1) * B // world)
  xs, cs = x[shard], codes[shard]
  # what the CUDA kernels produce on each rank, restated on the CPU: un-normalised sums over the local shard
  state = trainer._UpdateState(phi)
  state.grad.copy_(cs.t() @ (cs @ phi - xs))
  state.sq_sum.copy_((cs * cs).sum(0))
  trainer.allreduce_update_buffers(state, True)          # the code under test: ONE collective
  h = h0 * 0.99 + (state.sq_sum / B) / 100
  update = 0.1 * (state.grad / B) / (h[:, None] + 0.001)
  new_phi = phi - update
  new_phi = new_phi / new_phi.norm(p=2, dim=1)[:, None]
  # single-process answer
  want_h = oracle.hessian_running_mean(h0, codes)
  want_phi = oracle.sc_dictionary_update(x, phi, codes, want_h, stepsize=0.1)
  ok = oracle.relative_l2(new_phi, want_phi) < 1e-6 and oracle.relative_l2(h, want_h) < 1e-6
  # replicas identical bit for bit
  gathered = [torch.empty_like(new_phi) for _ in range(world)]
  dist.all_gather(gathered, new_phi)
  ok = ok and all(torch.equal(g, gathered[0]) for g in gathered)
  ok = ok and trainer._common.global_batch_size(B // world, torch.device('cpu')) == B
  # convolutional mode shards by image: the gradient sums of the shards add up to the full-batch gradient, and the
  # update applied with the global image count (what dict_update_rules.convolutional._common.descend does around its
  # all-reduce) is the single-process update of the reference
  xi, pad = oracle.synthetic_padded_images(4, 1, 24, 24, (8, 8), (4, 4))
  kern = oracle.synthetic_conv_dictionary(6, 1, 8, 8)
  ci = oracle.conv_ista_fista(xi, kern, (4, 4), pad, 0.05, 10)
  half = slice(rank * 2, rank * 2 + 2)
  g_local = oracle.conv_dictionary_gradient(xi[half], kern, ci[half], (4, 4), pad)
  g_sum = g_local.clone()
  dist.all_reduce(g_sum, op=dist.ReduceOp.SUM)
  hc = oracle.conv_hessian_running_mean(torch.zeros(6), ci)
  got = oracle.conv_sc_dictionary_update(xi[half], kern, ci[half], (4, 4), pad, hc, stepsize=0.05, batch_size=4,
                                         extra_gradient=g_sum - g_local)
  want = oracle.conv_sc_dictionary_update(xi, kern, ci, (4, 4), pad, hc, stepsize=0.05)
  ok = ok and oracle.relative_l2(got, want) < 1e-6
  # validation metrics (training/metrics.py): every rank holds the totals of its shard (restated here on the CPU the way
  # metrics_totals_kernel defines them); after the all-reduce both report the metrics of the whole batch
  from vision_transform_codes_b200.lean import metrics
  r2 = ((torch.mm(cs, phi) - xs)**2).sum(1).double()
  mse = (r2 / D).float()
  totals = torch.tensor([float(0.5 * r2.sum()), float(cs.abs().sum()), float((cs != 0).sum(1).double().div(S).sum()),
                         float(torch.log10(mse.double()).sum()), float(len(mse)), float(xs.min()), float(xs.max()),
                         float(len(mse))], dtype=torch.float64)
  metrics._allreduce_totals(totals)
  got_m = metrics.metrics_from_totals(totals.tolist(), 0.1)
  want_m = oracle.compute_metrics(x, codes, phi, phi, 0.1)
  for name in got_m:
    ok = ok and abs(got_m[name] - float(want_m[name])) <= 1e-5 * max(1.0, abs(float(want_m[name])))
  with open(result_path + str(rank), 'w') as f:
    f.write('ok' if ok else 'fail')
  dist.destroy_process_group()


def test_two_rank_update_equals_single_process(tmp_path):
  world, port = 2, _free_port()
  result = str(tmp_path / 'result')
  mp.spawn(_worker, args=(world, port, result), nprocs=world, join=True)
  for r in range(world):
    assert open(result + str(r)).read() == 'ok'


def test_enable_requires_initialised_process_group():
  import vision_transform_codes_b200 as pkg
  if dist.is_initialized():
    pytest.skip('process group already initialised in this interpreter')
  with pytest.raises(RuntimeError):
    pkg.enable_data_parallel()

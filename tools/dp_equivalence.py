"""torchrun helper: k data-parallel train steps on N GPUs must give the dictionary of the same steps on one GPU."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import vision_transform_codes_b200 as pkg  # noqa: E402
from oracle import vtc_oracle as oracle  # noqa: E402  (seeded inputs + comparison metric)
from vision_transform_codes_b200.lean import sparse_coding as trainer  # noqa: E402

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
dev = torch.device('cuda', int(os.environ['LOCAL_RANK']))
torch.cuda.set_device(dev)
dist.init_process_group('nccl', device_id=dev)
B, S, D, steps = 4096, 512, 256, 3
phi0 = oracle.synthetic_dictionary(S, D)
batches = oracle.synthetic_patches(steps * B, D).view(steps, B, D)
params = {'mode': 'fully-connected', 'num_epochs': 1, 'code_inference_algorithm': 'fista',
          'inference_param_schedule': {0: {'sparsity_weight': 0.1, 'num_iters': 40}},
          'dictionary_update_algorithm': 'sc_cheap_quadratic_descent',
          'dict_update_param_schedule': {0: {'stepsize': 0.1, 'num_iters': 1}}}
# single GPU (every rank computes it redundantly)
single = phi0.to(dev)
trainer.train_dictionary(batches.to(dev), None, single, params)
# data parallel: contiguous shards of every batch
pkg.enable_data_parallel()
shard = B // world
sharded = phi0.to(dev)
trainer.train_dictionary(batches[:, rank * shard:(rank + 1) * shard].to(dev), None, sharded, params)
err = oracle.relative_l2(sharded.cpu(), single.cpu())
ref = sharded.clone()
dist.broadcast(ref, src=0)
same = torch.tensor([int(torch.equal(ref, sharded))], device=dev)
dist.all_reduce(same, op=dist.ReduceOp.MIN)
if rank == 0:
  print('rel-L2 sharded vs single %.3e, replicas identical: %s' % (err, bool(same.item())))
  if err < 1e-5 and same.item() == 1:
    print('DP_EQUIVALENCE_OK')
dist.destroy_process_group()

"""Short driver for ncu: one sparse-coding dictionary update (Hessian mean + cheap quadratic descent) at the shard size
of BASELINE configs[2] on 8 GPUs (65536 patches, 1024 atoms, 16x16 patches). Not a benchmark.
  ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r02_update \
      python tools/profile_update.py"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import vision_transform_codes_b200 as pkg  # noqa: E402
from oracle import vtc_oracle as oracle  # noqa: E402  (seeded input generators only)
from vision_transform_codes_b200.analysis_transforms.fully_connected import ista_fista  # noqa: E402
from vision_transform_codes_b200.lean import sparse_coding as trainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=65536)
ap.add_argument('--atoms', type=int, default=1024)
ap.add_argument('--pixels', type=int, default=256)
args = ap.parse_args()
pkg.config.check_finite = False
dev = torch.device('cuda:0')
phi = oracle.synthetic_dictionary(args.atoms, args.pixels).to(dev)
x = oracle.synthetic_patches(args.batch, args.pixels).to(dev)
codes = ista_fista.run(x, phi, 0.1, 30)
h = torch.zeros(args.atoms, device=dev)
state = trainer._UpdateState(phi)
trainer.update_dictionary(x, phi, codes, h, 0.1, 1, state)   # warm-up (workspaces, tensor maps)
torch.cuda.synchronize()
torch.cuda.profiler.start()
trainer.update_dictionary(x, phi, codes, h, 0.1, 1, state)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('ok', float(phi.abs().mean()))

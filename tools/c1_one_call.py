"""One configs[0] call (batch 250, 256 atoms, D=256, 300 FISTA iterations) after a warm-up, for an ncu launch list."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import vtc_oracle as oracle
from vision_transform_codes_b200.analysis_transforms.fully_connected import ista_fista

phi = oracle.synthetic_dictionary(256, 256).cuda()
x = oracle.synthetic_patches(250, 256).cuda()
for _ in range(2):
  a = ista_fista.run(x, phi, 0.1, 300)
torch.cuda.synchronize()
print('ok', float(a.abs().sum()))

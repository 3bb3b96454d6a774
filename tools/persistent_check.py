"""Checksum of the codes of one configs[1] call; run with VTC_B200_PERSISTENT=0 and =1 and compare the lines."""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import vision_transform_codes_b200 as pkg
from vision_transform_codes_b200.analysis_transforms.fully_connected import ista_fista
from oracle import vtc_oracle as oracle

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = int(sys.argv[2]) if len(sys.argv) > 2 else 300
pkg.config.precision = sys.argv[3] if len(sys.argv) > 3 else 'bf16x3'
phi = oracle.synthetic_dictionary(1024, 256).cuda()
x = oracle.synthetic_patches(B, 256, kind='whitened').cuda()
digests = set()
for _ in range(3):
  a = ista_fista.run(x, phi, 0.1, T)
  digests.add(hashlib.sha256(a.cpu().numpy().tobytes()).hexdigest())
print('B=%d T=%d %s sha256 %s (%d distinct over 3 runs)' % (B, T, pkg.config.precision, sorted(digests)[0][:24], len(digests)))

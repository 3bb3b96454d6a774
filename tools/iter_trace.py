"""Timeline of CTA 0 of one launch of the one-launch iteration kernel (vtc_debug_iter_trace).
Needs a library with the trace points compiled in:
  NVCC_EXTRA=-DVTC_TRACE tools/ab_build.sh HEAD trace
  VTC_B200_LIB=tools/ab/libvtc_b200_trace.so python tools/iter_trace.py [B] [precision]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import vision_transform_codes_b200 as pkg
from vision_transform_codes_b200 import _lib
from vision_transform_codes_b200.analysis_transforms.fully_connected import ista_fista
from oracle import vtc_oracle as oracle

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
pkg.config.precision = sys.argv[2] if len(sys.argv) > 2 else 'bf16x3'
pkg.config.check_finite = False
lib = _lib.load()
S, D = 1024, 256
phi = oracle.synthetic_dictionary(S, D).cuda()
x = oracle.synthetic_patches(B, D).cuda()
ista_fista.run(x, phi, 0.1, 8)
buf = torch.zeros(4 * 2048, dtype=torch.int64, device='cuda')
# trace a middle iteration only: run 4 untraced, then 1 traced via a warm start
warm = ista_fista.run(x, phi, 0.1, 4)
torch.cuda.synchronize()
_lib.check(lib.vtc_debug_iter_trace(_lib.ptr(buf)))   # fails on a library without -DVTC_TRACE
ista_fista.run(x, phi, 0.1, 2, initial_codes=warm)   # the first launch is traced (the last one would skip R)
torch.cuda.synchronize()
host = buf.cpu().tolist()
ev = []
for role in range(4):
  n = host[role * 2048]
  ev += host[role * 2048 + 1: role * 2048 + 1 + min(n, 2047)]
KIND = {1: 'G_BEGIN', 2: 'G_END', 3: 'R_ISSUE', 4: 'E_BEGIN', 5: 'E_SUB', 6: 'E_END', 7: 'G_LOAD', 8: 'START', 9: 'STOP', 10: ' e_in', 11: ' e_ld', 12: ' e_cmp', 13: ' e_yw', 14: ' e_arr'}
rows = sorted(((e & 0xFFFFFFFFFFFF), (e >> 56) & 255, (e >> 48) & 255) for e in ev)
# keep the first launch (between the first START and the first STOP)
t0 = None
has_start = any(kind == 8 for _, kind, _ in rows)
for clk, kind, idx in rows:
  if t0 is None and (kind == 8 or not has_start):
    t0 = clk
  if t0 is None:
    continue
  print('%9.2f us  %-8s %d' % ((clk - t0) / 1.9e3, KIND.get(kind, kind), idx))
  if kind == 9:
    break

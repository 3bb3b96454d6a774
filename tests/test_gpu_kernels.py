"""Building blocks on the GPU: the tcgen05 contraction, the Lipschitz constant, the subspace gathers."""
import ctypes

import pytest
import torch

from oracle import vtc_oracle as oracle

pytestmark = pytest.mark.gpu


def matmul_nt(a, b, sub=None, precision=3):
  from vision_transform_codes_b200 import _lib
  lib = _lib.load()
  M, K = a.shape
  N = b.shape[0]
  out = torch.empty(M, N, device=a.device)
  ws = _lib.workspace(lib.vtc_matmul_nt_workspace_bytes(M, N, K, precision), a.device, 'matmul')
  _lib.check(lib.vtc_matmul_nt(_lib.ptr(a), _lib.ptr(b), _lib.ptr(sub), _lib.ptr(out), M, N, K, precision,
                               _lib.ptr(ws), ws.numel(), _lib.stream_ptr(a.device)))
  torch.cuda.synchronize()
  return out


# relative Frobenius error of the product for each arithmetic mode
TOL = {1: 6e-3, 3: 3e-5, 6: 1e-5}


@pytest.mark.parametrize('shape', [(128, 256, 64), (256, 512, 256), (250, 256, 256), (1000, 200, 72),
                                   (37, 20, 10), (4096, 1024, 1024), (129, 257, 65)])
@pytest.mark.parametrize('precision', [1, 3, 6])
def test_matmul_nt_against_float64(shape, precision):
  M, N, K = shape
  g = torch.Generator().manual_seed(M + N + K)
  a = torch.randn(M, K, generator=g).cuda()
  b = torch.randn(N, K, generator=g).cuda()
  want = a.double() @ b.double().t()
  got = matmul_nt(a, b, precision=precision)
  err = float((got.double() - want).norm() / want.norm())
  assert err < TOL[precision], err


def test_matmul_nt_subtracts_in_the_epilogue():
  g = torch.Generator().manual_seed(0)
  a, b = torch.randn(300, 96, generator=g).cuda(), torch.randn(520, 96, generator=g).cuda()
  sub = torch.randn(300, 520, generator=g).cuda()
  want = a.double() @ b.double().t() - sub.double()
  got = matmul_nt(a, b, sub=sub, precision=6)
  assert float((got.double() - want).norm() / want.norm()) < 1e-5


def test_matmul_nt_is_deterministic():
  g = torch.Generator().manual_seed(1)
  a, b = torch.randn(512, 320, generator=g).cuda(), torch.randn(768, 320, generator=g).cuda()
  assert torch.equal(matmul_nt(a, b), matmul_nt(a, b))


@pytest.mark.parametrize('shape', [(256, 256), (1024, 256), (64, 48), (300, 100), (4096, 1024)])
def test_lipschitz_constant(shape):
  from vision_transform_codes_b200 import _lib
  lib = _lib.load()
  S, D = shape
  phi = oracle.synthetic_dictionary(S, D).cuda()
  out = torch.zeros(1, device='cuda')
  ws = _lib.workspace(lib.vtc_lipschitz_workspace_bytes(S, D), phi.device, 'lip')
  _lib.check(lib.vtc_lipschitz(_lib.ptr(phi), S, D, _lib.ptr(out), _lib.ptr(ws), ws.numel(),
                               _lib.stream_ptr(phi.device)))
  want = torch.linalg.eigvalsh(phi.double().cpu().t() @ phi.double().cpu())[-1]
  assert abs(float(out.item()) - float(want)) / float(want) < 1e-6


def test_lipschitz_clustered_spectrum():
  # two nearly equal top eigenvalues: the squaring iteration must still land inside the cluster
  from vision_transform_codes_b200 import _lib
  lib = _lib.load()
  q, _ = torch.linalg.qr(torch.randn(64, 64, generator=torch.Generator().manual_seed(5)))
  sv = torch.ones(64)
  sv[0], sv[1] = 2.0, 2.0 * (1 - 1e-6)
  phi = (q * sv[None, :]).cuda().contiguous()  # phi^T phi = diag(sv^2)
  out = torch.zeros(1, device='cuda')
  ws = _lib.workspace(lib.vtc_lipschitz_workspace_bytes(64, 64), phi.device, 'lip')
  _lib.check(lib.vtc_lipschitz(_lib.ptr(phi), 64, 64, _lib.ptr(out), _lib.ptr(ws), ws.numel(),
                               _lib.stream_ptr(phi.device)))
  assert abs(float(out.item()) - 4.0) / 4.0 < 5e-6


def test_hessian_diagonal_running_mean():
  from vision_transform_codes_b200 import _lib
  lib = _lib.load()
  g = torch.Generator().manual_seed(2)
  codes = torch.randn(777, 130, generator=g).cuda()
  h = torch.rand(130, generator=g).cuda()
  want = oracle.hessian_running_mean(h.cpu(), codes.cpu())
  sq = torch.empty(130, device='cuda')
  _lib.check(lib.vtc_hessian_diag_update(_lib.ptr(codes), 130, 777, 130, 777, _lib.ptr(sq), _lib.ptr(h), 1,
                                         _lib.stream_ptr(codes.device)))
  torch.cuda.synchronize()
  assert oracle.relative_l2(h.cpu(), want) < 1e-6


def test_patch_extraction_is_the_reference_crop_loop():
  """Device-side data feed: bit-exact copy of the crops the reference's host loop takes."""
  from vision_transform_codes_b200.utils import dataset_generation as dg
  g = torch.Generator().manual_seed(5)
  for shape, patch in (((3, 64, 80), (16, 16)), ((2, 40, 33, 3), (8, 12)), ((1, 16, 16), (16, 16))):
    images = torch.randn(*shape, generator=g)
    full = images if images.dim() == 4 else images.unsqueeze(-1)
    n, h, w = shape[:3]
    edge = 0 if patch == (16, 16) and (h, w) == (16, 16) else 3
    if (h, w) == patch:
      corners = torch.zeros(5, 3, dtype=torch.int32)
    else:
      corners = dg.draw_patch_corners(500, n, (h, w), patch, edge, generator=g)
      assert int(corners[:, 1].min()) >= edge and int(corners[:, 1].max()) < h - patch[0] - edge
      assert int(corners[:, 2].min()) >= edge and int(corners[:, 2].max()) < w - patch[1] - edge
    want = oracle.extract_patches(full, corners, patch)
    got = dg.extract_patches(images.cuda(), corners.cuda(), patch)
    assert torch.equal(got.cpu(), want)
  batch = dg.sample_patches(torch.randn(4, 128, 128).cuda(), 4096, (16, 16), 5)
  assert tuple(batch.shape) == (4096, 256) and torch.isfinite(batch).all()


def test_whitening_against_reference_outputs_and_oracle():
  """Device-side whiten_center_surround (utils/image_processing.py; vtc_whitening_filter + cuFFT + vtc_spectrum_filter)
  against outputs of the reference (tests/golden/whitening_small.npz) and the oracle on a batch of larger images.
  Tolerance: the transfer function is fp64 rounded to float32 (1e-6 relative); the reference transforms in complex128,
  cuFFT in complex64: 1e-5 of the image range."""
  import numpy as np
  from conftest import load_golden
  from vision_transform_codes_b200.utils import image_processing
  g = load_golden('whitening_small')
  cut = {'low': 1e-3, 'high': 0.9}
  for key in ('gray', 'colour'):
    img = g[key].cuda()
    out, filt = image_processing.whiten_center_surround(img, cut, return_filter=True)
    assert tuple(out.shape) == tuple(img.shape) and out.dtype == torch.float32
    assert np.allclose(filt.cpu().numpy(), g[key + '_filter'].numpy(), rtol=1e-6, atol=0)
    assert float((out.cpu() - g[key + '_whitened']).abs().max()) < 1e-5
  raw = {'low': 0.05, 'high': 0.6}
  out, filt = image_processing.whiten_center_surround(g['gray'].cuda(), raw, return_filter=True,
                                                      norm_and_threshold=False)
  # un-normalised: the tail of the low-pass factor falls below the float32 range (absolute tolerance there)
  assert np.allclose(filt.cpu().numpy(), g['raw_filter'].numpy(), rtol=1e-6, atol=1e-30)
  assert float((out.cpu() - g['raw_whitened']).abs().max()) < 1e-5
  # a batch of 512x512 images (the size of BASELINE configs[4]'s inputs), every image filtered independently
  gen = torch.Generator().manual_seed(11)
  batch = torch.rand(3, 512, 512, 1, generator=gen)
  got = image_processing.whiten_center_surround(batch.cuda(), cut).cpu()
  for i in range(3):
    want = torch.from_numpy(oracle.whiten_center_surround(batch[i].numpy(), cut))
    assert float((got[i] - want).abs().max()) < 2e-5
  std = image_processing.standardize_data_range(3.0 * batch.cuda() - 1.0)
  assert float(std.min()) == 0.0 and float(std.max()) == 1.0
  with pytest.raises(RuntimeError):
    image_processing.whiten_center_surround(batch, cut)   # CPU tensor: no fallback


def test_small_batch_kernel_equals_the_tiled_schedule():
  """csrc/fista_small_kernel.cuh (everything on chip, one launch for the whole run; what BASELINE configs[0] takes) and
  the tiled Gram-form schedule (one launch per iteration) are the same arithmetic: bit-identical codes, for ragged
  batch sizes, fewer than 256 atoms, a warm start, ISTA and the non-negative threshold, in both parity precisions."""
  import vision_transform_codes_b200 as pkg
  from vision_transform_codes_b200 import _lib
  from vision_transform_codes_b200.analysis_transforms.fully_connected import ista_fista
  lib = _lib.load()
  saved = pkg.config.precision
  try:
    for precision in ('bf16x3', 'bf16'):
      pkg.config.precision = precision
      for (b, s, d) in ((250, 256, 256), (33, 200, 128), (1, 64, 64), (700, 256, 200)):
        x = oracle.synthetic_patches(b, d, seed=b).cuda()
        phi = oracle.synthetic_dictionary(s, d).cuda()
        warm = ista_fista.run(x, phi, 0.1, 3)
        for kw in ({}, {'variant': 'ista'}, {'nonnegative_only': True}, {'initial_codes': warm}):
          got = {}
          for small in (1, 0):
            _lib.check(lib.vtc_set_small_batch_kernel(small))
            n0 = lib.vtc_launch_count()
            got[small] = ista_fista.run(x, phi, 0.1, 40, **kw)
            got[small, 'launches'] = lib.vtc_launch_count() - n0
          assert torch.equal(got[1], got[0]), (precision, b, s, d, kw)
          assert got[1, 'launches'] < got[0, 'launches'] - 30   # one launch instead of forty
    # the oracle agrees (this shape is tests/golden/inference_config1.npz's)
    pkg.config.precision = 'bf16x3'
    x, phi = oracle.synthetic_patches(250, 256), oracle.synthetic_dictionary(256, 256)
    want = oracle.ista_fista(x, phi, 0.1, 60)
    assert oracle.relative_l2(ista_fista.run(x.cuda(), phi.cuda(), 0.1, 60).cpu(), want) < 1e-4
  finally:
    lib.vtc_set_small_batch_kernel(1)
    pkg.config.precision = saved


@pytest.mark.gpu
def test_workspace_cache_grows_and_gives_back():
  """The scratch cache grows to the largest request of a (device, stream, tag), is reused for smaller ones, and a
  buffer far larger than the request is replaced instead of staying pinned (_lib.workspace)."""
  from vision_transform_codes_b200 import _lib
  dev = torch.device('cuda:0')
  _lib.release_workspaces()
  floor = _lib.WORKSPACE_SHRINK_FLOOR
  a = _lib.workspace(1 << 20, dev, 'test_ws')
  assert a.numel() == 1 << 20
  b = _lib.workspace(2 * floor, dev, 'test_ws')          # grows
  assert b.numel() == 2 * floor
  assert _lib.workspace(floor, dev, 'test_ws') is b      # within the factor: reused
  c = _lib.workspace(1 << 20, dev, 'test_ws')            # far smaller than the cached buffer: given back
  assert c.numel() == 1 << 20 and c is not b
  assert _lib.workspace(1 << 20, dev, 'other_tag') is not c
  _lib.release_workspaces()


@pytest.mark.gpu
def test_c_abi_writes_only_the_codes_it_owns():
  """vtc_fista_fc with a PITCHED output inside a larger buffer: every float outside the (B, S) block -- the pitch
  padding of every row and guard zones before and after -- keeps its canary, the images, the dictionary and the warm
  start are not written (include/vtc_b200.h), and the block equals the dense call. All schedules: the panel-resident
  kernel (S > 2 D), the Gram form, the small-batch kernel, ragged shapes."""
  import vision_transform_codes_b200 as pkg
  from vision_transform_codes_b200 import _lib
  from vision_transform_codes_b200.analysis_transforms.fully_connected import ista_fista
  lib = _lib.load()
  dev = torch.device('cuda:0')
  prec = pkg.config.precision_code()
  canary = 12345.0
  for (b, s, d, warm) in ((600, 1024, 256, False), (257, 1000, 250, True), (250, 256, 256, False), (33, 200, 128, True),
                          (1500, 520, 64, False)):
    x = oracle.synthetic_patches(b, d, seed=b).to(dev)
    phi = oracle.synthetic_dictionary(s, d).to(dev)
    init = (0.01 * torch.randn(b, s, generator=torch.Generator().manual_seed(1))).to(dev) if warm else None
    want = ista_fista.run(x, phi, 0.1, 12, initial_codes=init)
    ld = s + 12                      # multiple of 4 floats: rows stay 16-byte aligned
    guard = 4096
    buf = torch.full((guard + b * ld + guard,), canary, dtype=torch.float32, device=dev)
    out = buf[guard:guard + b * ld].view(b, ld)
    x0, phi0 = x.clone(), phi.clone()
    init_p = None
    if warm:
      init_buf = torch.full((b, ld), canary, dtype=torch.float32, device=dev)
      init_buf[:, :s] = init
      init_p, init0 = init_buf, init_buf.clone()
    with torch.cuda.device(dev):
      nbytes = lib.vtc_fista_workspace_bytes(b, s, d, prec)
      ws = _lib.workspace(nbytes, dev, 'fista_guard_test')
      iters = ctypes.c_int(0)
      _lib.check(lib.vtc_fista_fc(_lib.ptr(x), d, _lib.ptr(phi), _lib.ptr(init_p), _lib.ptr(out), ld, b, s, d, 0.1, 12, 1,
                                  0, 0, 1, -1.0, prec, _lib.ptr(ws), ws.numel(), ctypes.byref(iters), None,
                                  _lib.stream_ptr(dev)))
    torch.cuda.synchronize()
    assert torch.equal(out[:, :s], want), (b, s, d, warm)
    assert bool((out[:, s:] == canary).all()), 'pitch padding written'
    assert bool((buf[:guard] == canary).all()) and bool((buf[guard + b * ld:] == canary).all()), 'guard zone written'
    assert torch.equal(x, x0) and torch.equal(phi, phi0)
    if warm:
      assert torch.equal(init_p, init0)

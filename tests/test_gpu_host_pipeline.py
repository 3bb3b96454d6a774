"""HostPipeline (vision_transform_codes_b200/host_pipeline.py): inference from pinned host buffers with the copies
overlapped; every submitted batch must give exactly what a plain ista_fista.run on the same data gives."""
import pytest
import torch

from oracle import vtc_oracle as oracle

pytestmark = pytest.mark.gpu


def test_pipeline_results_equal_plain_calls_and_arrive_in_their_own_buffers():
  from vision_transform_codes_b200.analysis_transforms.fully_connected import ista_fista
  from vision_transform_codes_b200.host_pipeline import HostPipeline
  dev = torch.device('cuda:0')
  phi = oracle.synthetic_dictionary(1024, 256).to(dev)
  batches = [oracle.synthetic_patches(768, 256, seed=s).pin_memory() for s in range(5)]
  outs = [torch.empty((768, 1024)).pin_memory() for _ in batches]
  pipe = HostPipeline(dev, depth=2)
  for x, o in zip(batches, outs):
    pipe.submit(x, phi, 0.1, 25, out=o)
  pipe.synchronize()
  for x, o in zip(batches, outs):
    want = ista_fista.run(x.to(dev), phi, 0.1, 25).cpu()
    assert torch.equal(o, want)
  assert pipe.h2d_bytes == 5 * 768 * 256 * 4 and pipe.d2h_bytes == 5 * 768 * 1024 * 4
  # keyword arguments of run() pass through
  o = torch.empty((768, 1024)).pin_memory()
  pipe.submit(batches[0], phi, 0.1, 25, out=o, variant='ista', nonnegative_only=True)
  pipe.synchronize()
  assert torch.equal(o, ista_fista.run(batches[0].to(dev), phi, 0.1, 25, variant='ista', nonnegative_only=True).cpu())


def test_pipeline_keeps_the_references_overflow_error():
  """ista_fista.py:75-79: a dictionary that overflowed raises RuntimeError (checked once per dictionary here)."""
  from vision_transform_codes_b200.host_pipeline import HostPipeline
  dev = torch.device('cuda:0')
  phi = oracle.synthetic_dictionary(512, 64).to(dev)
  phi[3, 5] = float('inf')
  x = oracle.synthetic_patches(256, 64).pin_memory()
  pipe = HostPipeline(dev)
  with pytest.raises(RuntimeError):
    pipe.submit(x, phi, 0.1, 5, out=torch.empty((256, 512)).pin_memory())

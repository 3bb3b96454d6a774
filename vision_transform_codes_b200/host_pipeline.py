"""
Inference from pinned HOST buffers with the copies hidden under the compute of neighbouring calls.

The reference API takes tensors that already live on the device (analysis_transforms/fully_connected/ista_fista.py:14-16);
a caller that feeds batches from host memory pays, per batch, one host-to-device copy of the images and one device-to-host
copy of the codes (67 MB and 268 MB at BASELINE configs[1]: about 7 ms of a PCIe Gen5 x16 link against 60-80 ms of
compute, more when eight GPUs share the host's links). ``HostPipeline`` is a small helper (not a new reference API) that
keeps ``depth`` batches in flight on three streams -- upload, compute, download -- so that the upload of batch i + 1 and
the download of batch i - 1 run while batch i computes:

    pipe = HostPipeline(device, depth=2)
    for x_host, codes_host in batches:                      # pinned tensors
      pipe.submit(x_host, dictionary, 0.1, 300, out=codes_host)
    pipe.synchronize()                                      # every codes_host is complete

Each submit is exactly one ``ista_fista.run`` call on the compute stream (same arguments, same results); nothing is
approximated or skipped. The one thing moved is the reference's "dictionary overflowed" check (ista_fista.py:75-79):
``run`` makes it with one host synchronisation per call, which would serialise the pipeline (measured on configs[1]:
86.5 ms per step against 78.3 without, device-resident 76.5; ``tools/pipeline_probe.py``), so the pipeline makes it
ONCE per dictionary (per tensor version) with ``vtc_lipschitz`` and runs the calls themselves unsynchronised.
"""
import torch

import vision_transform_codes_b200 as pkg
from vision_transform_codes_b200 import _lib
from vision_transform_codes_b200.analysis_transforms.fully_connected import ista_fista


class HostPipeline:

  def __init__(self, device, depth=2):
    self.device = torch.device(device)
    self.depth = int(depth)
    if self.depth < 1:
      raise ValueError('depth must be >= 1')
    with torch.cuda.device(self.device):
      self.upload = torch.cuda.Stream()
      self.compute = torch.cuda.Stream()
      self.download = torch.cuda.Stream()
    # per slot: device images, device codes (kept alive until the download has read them), completion event
    self.slots = [{'x': None, 'codes': None, 'done': None} for _ in range(self.depth)]
    self.submitted = 0
    self.h2d_bytes = 0
    self.d2h_bytes = 0
    self._checked = None   # (data_ptr, version, shape) of the dictionary whose step size was found finite

  def _check_dictionary(self, dictionary):
    key = (dictionary.data_ptr(), dictionary._version, tuple(dictionary.shape))
    if key == self._checked:
      return
    lib = _lib.load()
    S, D = dictionary.shape
    with torch.cuda.device(self.device), torch.cuda.stream(self.compute):
      ws = _lib.workspace(lib.vtc_lipschitz_workspace_bytes(S, D), self.device, 'lipschitz')
      out = torch.empty(1, dtype=torch.float32, device=self.device)
      _lib.check(lib.vtc_lipschitz(_lib.ptr(dictionary.contiguous()), S, D, _lib.ptr(out), _lib.ptr(ws), ws.numel(),
                                   _lib.stream_ptr(self.device)))
      value = float(out.item())
    if not (value == value and abs(value) != float('inf')):
      print('symeig threw an exception. Likely due to one of the dictionary',
            'elements overflowing. The norm of each dictionary element is')
      print(torch.norm(dictionary, dim=1, p=2))
      raise RuntimeError()
    self._checked = key

  def submit(self, images_host, dictionary, sparsity_weight, num_iters, out, **run_kwargs):
    """Enqueues upload -> ista_fista.run -> download of one batch; returns the event that marks ``out`` complete."""
    if images_host.device.type != 'cpu' or out.device.type != 'cpu':
      raise ValueError('images_host and out must be host tensors (pinned for the copies to be asynchronous)')
    if pkg.config.check_finite:
      self._check_dictionary(dictionary)
    slot = self.slots[self.submitted % self.depth]
    self.submitted += 1
    with torch.cuda.device(self.device):
      if slot['done'] is not None:
        # the slot's previous batch: its device buffers are reused, and a host thread that submits faster than the
        # device drains must not queue unboundedly
        slot['done'].synchronize()
      if slot['x'] is None or slot['x'].shape != images_host.shape:
        slot['x'] = torch.empty(images_host.shape, dtype=torch.float32, device=self.device)
      uploaded = torch.cuda.Event()
      with torch.cuda.stream(self.upload):
        slot['x'].copy_(images_host, non_blocking=True)
        uploaded.record(self.upload)
      computed = torch.cuda.Event()
      with torch.cuda.stream(self.compute):
        self.compute.wait_event(uploaded)
        saved = pkg.config.check_finite
        pkg.config.check_finite = False   # made once per dictionary above: no host synchronisation inside the call
        try:
          codes = ista_fista.run(slot['x'], dictionary, sparsity_weight, num_iters, **run_kwargs)
        finally:
          pkg.config.check_finite = saved
        computed.record(self.compute)
      done = torch.cuda.Event()
      with torch.cuda.stream(self.download):
        self.download.wait_event(computed)
        out.copy_(codes, non_blocking=True)
        done.record(self.download)
      # (the slot's buffers are only reused after `done` of this batch has been waited for at the top of a later submit,
      # which also keeps `codes` alive until the download stream has read it)
      slot['codes'], slot['done'] = codes, done
      self.h2d_bytes += images_host.numel() * 4
      self.d2h_bytes += out.numel() * 4
    return done

  def synchronize(self):
    for slot in self.slots:
      if slot['done'] is not None:
        slot['done'].synchronize()

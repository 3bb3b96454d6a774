"""
Device-side data feed for dictionary training: random patches cut out of images that already live on the GPU.

The reference builds its training sets on the host (utils/dataset_generation.py:22-289: preprocess whole images, then a
Python loop that copies one random crop per sample, :207-218) and feeds them through a DataLoader. For batches of
hundreds of thousands of patches per step that loop is the bottleneck, so this module keeps the (preprocessed) images on
the device and cuts every batch there: positions are drawn exactly as the reference draws them (image uniformly, top /
left uniformly in [edge_buffer, size - patch - edge_buffer)), the copy is one kernel (``vtc_extract_patches``).
"""
import os
import sys

import torch

try:
  from vision_transform_codes_b200 import _lib
except ImportError:
  sys.path.append(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
  from vision_transform_codes_b200 import _lib


def draw_patch_corners(num_samples, num_images, image_dimensions, patch_dimensions, edge_buffer, generator=None,
                       device=None):
  """(num_samples, 3) int32 [image, top, left], distributed like utils/dataset_generation.py:207-212."""
  h, w = image_dimensions
  max_v, max_h = h - patch_dimensions[0] - edge_buffer, w - patch_dimensions[1] - edge_buffer
  if max_v <= edge_buffer or max_h <= edge_buffer:
    raise ValueError('image too small for this patch size and edge buffer')
  kw = {'generator': generator, 'device': device}
  img = torch.randint(0, num_images, (num_samples,), **kw)
  top = torch.randint(edge_buffer, max_v, (num_samples,), **kw)
  left = torch.randint(edge_buffer, max_h, (num_samples,), **kw)
  return torch.stack([img, top, left], dim=1).to(torch.int32).contiguous()


def extract_patches(images, corners, patch_dimensions, flatten_patches=True):
  """
  Cut patches out of device-resident images.

  Parameters
  ----------
  images : torch.Tensor(float32, size=(n, h, w) or (n, h, w, c))
      Preprocessed images on the GPU (channel last, as the reference holds them).
  corners : torch.Tensor(int32, size=(b, 3))
      [image index, top row, left column] of every patch.
  patch_dimensions : tuple(int, int)
  flatten_patches : bool, optional
      (b, ph*pw*c) if True (the layout the fully-connected trainer takes), else (b, ph, pw, c). Default True.
  """
  _lib.require_cuda_f32(images, 'images')
  if images.dim() == 3:
    images = images.unsqueeze(-1)
  if images.dim() != 4:
    raise ValueError('expected images (n, h, w) or (n, h, w, c)')
  if corners.dtype != torch.int32 or corners.dim() != 2 or corners.size(1) != 3 or not corners.is_cuda:
    raise ValueError('corners must be a CUDA int32 tensor of shape (b, 3)')
  images = images.contiguous()
  corners = corners.contiguous()
  n, h, w, c = images.shape
  ph, pw = int(patch_dimensions[0]), int(patch_dimensions[1])
  b = corners.size(0)
  out = torch.empty((b, ph * pw * c), dtype=torch.float32, device=images.device)
  if b > 0:
    lib = _lib.load()
    with torch.cuda.device(images.device):
      _lib.check(lib.vtc_extract_patches(_lib.ptr(images), n, h, w, c, _lib.ptr(corners), b, ph, pw, _lib.ptr(out),
                                         ph * pw * c, _lib.stream_ptr(images.device)))
  return out if flatten_patches else out.view(b, ph, pw, c)


def sample_patches(images, num_samples, patch_dimensions, edge_buffer, generator=None, flatten_patches=True):
  """One training batch: num_samples random patches of device-resident images (see draw_patch_corners)."""
  n, h, w = images.shape[:3]
  corners = draw_patch_corners(num_samples, n, (h, w), patch_dimensions, edge_buffer, generator, images.device)
  return extract_patches(images, corners, patch_dimensions, flatten_patches)


class DeviceBatches:
  """Iterable of training batches cut from a patch matrix that lives on the device (what the reference's examples
  build from a DataLoader over ``OneOutputDset``, examples/train_sparse_coding.py:83-92): every pass reshuffles the
  patches (one ``randperm`` + one gather on the device) and yields ``(batch, ...)`` tensors; the last, smaller batch is
  kept unless ``drop_last``. ``train_dictionary`` only iterates its dataset argument, so this is a drop-in for it."""

  def __init__(self, patches, batch_size, shuffle=True, drop_last=False, generator=None):
    self.patches, self.batch_size, self.shuffle, self.drop_last = patches, int(batch_size), shuffle, drop_last
    self.generator = generator

  def __len__(self):
    n = self.patches.size(0)
    return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

  def __iter__(self):
    n = self.patches.size(0)
    data = self.patches
    if self.shuffle:
      data = data[torch.randperm(n, device=data.device, generator=self.generator)]
    for i in range(len(self)):
      yield data[i * self.batch_size:(i + 1) * self.batch_size]


def load_patch_dataset(path, device, batch_size=None, validation_batch_size=None, shuffle=True, drop_last=False):
  """
  Reads a dataset pickle in the reference's format and puts it on the device.

  The reference's dataset scripts (tests/dset_generation_1.py:13-25, :28-40, :44-63) pickle
  ``{'training': {'patches': float32 ndarray, ...}, 'validation': {'patches': ...}}`` where ``patches`` is ``(N, D)``
  for flattened patches or ``(N, c, h, w)`` for (padded) image patches; extra keys of the inner dictionaries
  (``local_contrasts``, ``original_component_means``, ... utils/dataset_generation.py:314-333) are kept as numpy arrays.

  Returns ``{'training': ..., 'validation': ...}``: device tensors when ``batch_size`` is None, else ``DeviceBatches``
  iterables ready to be passed to ``train_dictionary`` (validation batches default to ten training batches, as in
  examples/train_sparse_coding.py:88-90), plus ``'extras'`` with the remaining arrays per split.
  """
  import pickle

  import numpy as np
  with open(str(path), 'rb') as f:
    raw = pickle.load(f)
  if not isinstance(raw, dict) or 'training' not in raw:
    raise ValueError("not a dataset pickle of the reference's format: expected a dict with a 'training' entry")
  out, extras = {}, {}
  for split in ('training', 'validation'):
    if split not in raw:
      continue
    entry = raw[split]
    if not isinstance(entry, dict) or 'patches' not in entry:
      raise ValueError("dataset split %r has no 'patches' array" % split)
    patches = torch.from_numpy(np.ascontiguousarray(entry['patches'], dtype=np.float32)).to(device)
    extras[split] = {k: v for k, v in entry.items() if k != 'patches'}
    if batch_size is None:
      out[split] = patches
    elif split == 'training':
      out[split] = DeviceBatches(patches, batch_size, shuffle=shuffle, drop_last=drop_last)
    else:
      out[split] = DeviceBatches(patches, validation_batch_size or 10 * batch_size, shuffle=False)
  out['extras'] = extras
  return out

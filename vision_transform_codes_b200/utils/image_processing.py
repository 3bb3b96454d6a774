"""
Device-side counterpart of the reference's whole-image whitening (utils/image_processing.py:267-308,
whiten_center_surround, as called per image by utils/dataset_generation.py with cutoffs low 1e-3 / high 0.9).

The reference filters one (h, w, c) numpy image at a time on the host. Here a batch (n, h, w, c) of equally sized images
stays on the GPU: the transfer function is built by a CUDA kernel (fp64, like numpy), the DFTs are cuFFT through
torch.fft, the spectrum is multiplied in place by a CUDA kernel. Together with utils/dataset_generation.py (patch
extraction on the device) a 524 288-patch batch is produced without a host round trip.
"""
import torch

from vision_transform_codes_b200 import _lib


def whitening_filter(dft_num_samples, cutoffs, norm_and_threshold=True, device=None, order=8.0):
  """
  The real transfer function rolled_off_ramp * low_pass of whiten_center_surround (utils/image_processing.py:292-302).

  Parameters
  ----------
  dft_num_samples : (int, int)
      Samples of the DFT along the vertical and the horizontal axis (the image size).
  cutoffs : dictionary
      'low' : below this spatial frequency the ramp is flat; 'high' : cutoff of the exponential low-pass filter, as a
      fraction of the Nyquist frequency.
  norm_and_threshold : bool, optional
      Maximum magnitude 1.0, values below 1e-3 set to 1e-3. Default True.

  Returns
  -------
  torch.Tensor(float32, size=dft_num_samples) on `device`
  """
  device = torch.device('cuda' if device is None else device)
  if device.type != 'cuda':
    raise RuntimeError('vision_transform_codes_b200 only runs on a CUDA (sm_100) device and has no CPU fallback')
  h, w = int(dft_num_samples[0]), int(dft_num_samples[1])
  lib = _lib.load()
  out = torch.empty((h, w), dtype=torch.float32, device=device)
  scratch = torch.empty(8, dtype=torch.uint8, device=device)
  with torch.cuda.device(device):
    _lib.check(lib.vtc_whitening_filter(h, w, float(cutoffs['low']), float(cutoffs['high']), float(order),
                                        int(bool(norm_and_threshold)), _lib.ptr(out), _lib.ptr(scratch),
                                        _lib.stream_ptr(device)))
  return out


def filter_fd(images, filter_dft):
  """
  utils/image_processing.py:63-92 for a batch: every colour channel of every image filtered with a real, zero-phase
  transfer function given on the image's own DFT grid.

  Parameters
  ----------
  images : torch.Tensor(float32, size=(n, h, w, c) or (h, w, c)) on the GPU
  filter_dft : torch.Tensor(float32, size=(h, w)) on the GPU
  """
  _lib.require_cuda_f32(images, 'images')
  _lib.require_cuda_f32(filter_dft, 'filter_dft')
  single = images.dim() == 3
  if single:
    images = images[None]
  if images.dim() != 4:
    raise ValueError('images must have shape (n, h, w, c) or (h, w, c), got %s' % (tuple(images.shape),))
  n, h, w, c = images.shape
  if tuple(filter_dft.shape) != (h, w):
    # the reference zero-pads the image up to a larger DFT (:84-85); not needed by the whitening, which sizes the
    # filter to the image
    raise NotImplementedError('filter_dft must have the image size %s, got %s' % ((h, w), tuple(filter_dft.shape)))
  lib = _lib.load()
  device = images.device
  spectrum = torch.fft.fft2(images, dim=(1, 2)).contiguous()   # cuFFT, complex64 (n, h, w, c)
  with torch.cuda.device(device):
    _lib.check(lib.vtc_spectrum_filter(spectrum.data_ptr(), n, h * w, c, _lib.ptr(filter_dft.contiguous()),
                                       _lib.stream_ptr(device)))
  out = torch.fft.ifft2(spectrum, dim=(1, 2)).real.contiguous()
  return out[0] if single else out


def whiten_center_surround(images, cutoffs, return_filter=False, norm_and_threshold=True):
  """
  utils/image_processing.py:267-308 on device-resident images: (n, h, w, c) or a single (h, w, c), float32.
  Same arguments as the reference; returns the filtered images (and the transfer function when return_filter).
  """
  _lib.require_cuda_f32(images, 'images')
  h, w = (images.shape[0], images.shape[1]) if images.dim() == 3 else (images.shape[1], images.shape[2])
  combined_filter = whitening_filter((h, w), cutoffs, norm_and_threshold, images.device)
  filtered = filter_fd(images, combined_filter)
  return (filtered, combined_filter) if return_filter else filtered


def standardize_data_range(images):
  """utils/dataset_generation.py:169-182: the dataset's minimum to 0 and its maximum to 1 (relative luminances kept)."""
  _lib.require_cuda_f32(images, 'images')
  lo, hi = images.amin(), images.amax()
  return (images - lo) / (hi - lo)

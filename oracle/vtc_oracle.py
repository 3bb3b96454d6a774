"""
CPU oracle for the sparse-coding hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A restatement, in plain float32 (or float64) PyTorch on the CPU, of what the reference
(spencerkent/vision-transform-codes) computes on this path. Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; nothing under vision_transform_codes_b200/ does.

Pinning: the reference's own tests hold no numeric golden vectors, so this oracle is pinned
against OUTPUTS OF THE REFERENCE ITSELF, generated in the authoring container by tests/golden/make_golden.py (which
imports /root/reference with a torch.symeig shim) and committed as tests/golden/*.npz; tests/test_oracle.py checks
every function below against those files, plus closed-form known answers (orthonormal dictionary, groups of one).

Reference lines followed (paths relative to vision_transform_codes/):
  step size            analysis_transforms/fully_connected/ista_fista.py:72-80
  gradient step        ista_fista.py:105-106
  thresholds           ista_fista.py:107-121
  FISTA momentum       ista_fista.py:123-133
  early stopping       ista_fista.py:135-144
  subspace grouping    analysis_transforms/fully_connected/subspace_ista_fista.py:94-123, :184-190
  group shrinkage      subspace_ista_fista.py:144-156
  dictionary update    dict_update_rules/fully_connected/sc_cheap_quadratic_descent.py:42-48, sc_steepest_descent.py:37-41
  Hessian running mean training/sparse_coding.py:154
  train step           training/sparse_coding.py:513-515
  convolutional ISTA/FISTA   analysis_transforms/convolutional/ista_fista.py:104-197, utils/convolutions.py:7-24
  convolutional dict update  dict_update_rules/convolutional/sc_cheap_quadratic_descent.py:59-79,
                             sc_steepest_descent.py:55-72; Hessian running mean training/sparse_coding.py:158-161
  validation metrics   training/sparse_coding.py:177-229, utils/plotting.py:17-39 (compute_pSNR)
  image whitening      utils/image_processing.py:63-92 (filter_fd), :173-231, :234-264, :267-308 (whiten_center_surround)
"""
import torch


def lipschitz_constant(dictionary):
  """Largest eigenvalue of dictionary^T dictionary (n x n), as the reference's symeig(...)[0][-1]."""
  return torch.linalg.eigvalsh(torch.mm(dictionary.t(), dictionary))[-1]


def momentum_coefficients(num_iters):
  """beta_k = (t_k - 1) / t_{k+1}, t_1 = 1, t_{k+1} = (1 + sqrt(1 + 4 t_k^2)) / 2, in Python doubles."""
  t_k, out = 1.0, []
  for _ in range(num_iters):
    t_next = (1 + (1 + (4 * t_k**2))**0.5) / 2
    out.append((t_k - 1) / t_next)
    t_k = t_next
  return out


def threshold(pre, cutoff, nonnegative_only=False, hard_threshold=False):
  """The four scalar proximal maps of ista_fista.py:107-121, applied out of place."""
  if hard_threshold:
    out = pre.clone()
    if nonnegative_only:
      out[out < cutoff] = 0
    else:
      out[torch.abs(out) < cutoff] = 0
    return out
  if nonnegative_only:
    return (pre - cutoff).clamp_(min=0.)
  return torch.sign(pre) * (torch.abs(pre) - cutoff).clamp_(min=0.)


def _iterate(images, synthesis, prox, num_iters, variant, start, early_stopping_epsilon, stepsize):
  """
  Common ISTA/FISTA loop: `synthesis` is the (possibly grouped) dictionary, `start` the first gradient-evaluation
  point (flattened codes), `prox` maps the pre-threshold codes to the thresholded ones.
  Returns (codes, iterations run).
  """
  assert variant in ['ista', 'fista']
  eval_pt = start
  previous = start.clone()
  betas = momentum_coefficients(num_iters)
  codes = None
  done = 0
  for k in range(num_iters):
    residual = torch.mm(eval_pt, synthesis) - images
    codes = prox(eval_pt - stepsize * torch.mm(residual, synthesis.t()))
    change = codes - previous
    if variant == 'fista':
      eval_pt = codes + betas[k] * change
    else:
      eval_pt = codes
    previous = codes
    done = k + 1
    if early_stopping_epsilon is not None:
      if bool(torch.mean(torch.abs(change) / stepsize) < early_stopping_epsilon) and k > 0:
        break
  return codes, done


def ista_fista(images, dictionary, sparsity_weight, num_iters, variant='fista', initial_codes=None,
               early_stopping_epsilon=None, nonnegative_only=False, hard_threshold=False, return_iters=False):
  """analysis_transforms/fully_connected/ista_fista.py:14-148."""
  stepsize = 1. / lipschitz_constant(dictionary)
  cutoff = sparsity_weight * stepsize
  if initial_codes is None:
    start = images.new_zeros(images.size(0), dictionary.size(0))
  else:
    start = initial_codes
  codes, done = _iterate(images, dictionary, lambda pre: threshold(pre, cutoff, nonnegative_only, hard_threshold),
                         num_iters, variant, start, early_stopping_epsilon, stepsize)
  return (codes, done) if return_iters else codes


def subspace_ista_fista(images, dictionary, group_assignments, sparsity_weight, num_iters, variant='fista',
                        initial_codes=None, early_stopping_epsilon=None, return_iters=False):
  """analysis_transforms/fully_connected/subspace_ista_fista.py:23-192 (soft group threshold, summed duplicates)."""
  num_groups = len(group_assignments)
  width = max(len(g) for g in group_assignments)
  b = images.size(0)
  # slot (g, j) of the padded layout holds atom group_assignments[g][j]; padding slots have a zero synthesis row
  grouped_dictionary = images.new_zeros(num_groups * width, dictionary.size(1))
  start = images.new_zeros(b, num_groups * width)
  for g, members in enumerate(group_assignments):
    members = list(members)
    grouped_dictionary[g * width:g * width + len(members)] = dictionary[members]
    if initial_codes is not None:
      start[:, g * width:g * width + len(members)] = initial_codes[:, members]
  stepsize = 1. / lipschitz_constant(grouped_dictionary)
  cutoff = sparsity_weight * stepsize

  def group_shrink(pre):
    pre3 = pre.view(b, num_groups, width)
    norms = torch.norm(pre3, p=2, dim=2, keepdim=True)
    norms[norms == 0] = 1.0
    return (pre3 * torch.clamp(1 - (cutoff / norms), min=0.)).reshape(b, num_groups * width)

  grouped, done = _iterate(images, grouped_dictionary, group_shrink, num_iters, variant, start,
                           early_stopping_epsilon, stepsize)
  codes = images.new_zeros(b, dictionary.size(0))
  for g, members in enumerate(group_assignments):
    members = list(members)
    codes[:, members] = codes[:, members] + grouped[:, g * width:g * width + len(members)]
  return (codes, done) if return_iters else codes


def alignment_regularization_gradients(group_rows, dict_is_normalized):
  """
  Gradient of sum_{i != j} |cos(phi_i, phi_j)| over the rows of one group
  (dict_update_rules/fully_connected/subspace_sc_cheap_quadratic_descent.py:91-127).
  """
  m = group_rows.size(0)
  inner = torch.mm(group_rows, group_rows.t())
  if dict_is_normalized:
    cos = inner[:, :, None]
    own = cos * group_rows[:, None, :].expand(m, m, -1)       # depends on phi_i
    other = group_rows[None, :, :].expand(m, m, -1)           # depends on phi_j
  else:
    norms = torch.norm(group_rows, p=2, dim=1, keepdim=True)
    outer = torch.mm(norms, norms.t())
    cos = (inner / outer)[:, :, None]
    own = (cos / (norms**2)[:, None]) * group_rows[:, None, :].expand(m, m, -1)
    other = group_rows[None, :, :].expand(m, m, -1) / outer[:, :, None]
  return torch.sum(torch.sign(cos) * (other - own), dim=1)


def sc_dictionary_update(images, dictionary, codes, hessian_diagonal=None, stepsize=0.001, num_iters=1,
                         lowest_code_val=0.001, normalize_dictionary=True, batch_size=None, extra_gradient=None,
                         group_assignments=None, alignment_penalty=0.0):
  """
  sc_cheap_quadratic_descent.py:42-48 (hessian_diagonal given) / sc_steepest_descent.py:37-41 (None).
  Returns the updated dictionary (the reference updates in place; the oracle is functional).
  batch_size overrides the divisor (global batch of a data-parallel step); extra_gradient is the summed gradient
  contribution of the other shards for the FIRST iteration.
  """
  phi = dictionary.clone()
  divisor = codes.size(0) if batch_size is None else batch_size
  for it in range(num_iters):
    gradient = torch.mm(codes.t(), torch.mm(codes, phi) - images)
    if extra_gradient is not None and it == 0:
      gradient = gradient + extra_gradient
    data_term = gradient / divisor
    if alignment_penalty != 0:
      # subspace_sc_cheap_quadratic_descent.py:59-75: per-group gradients accumulated over groups
      reg = torch.zeros_like(phi)
      for members in group_assignments:
        members = list(members)
        reg[members] = reg[members] + alignment_regularization_gradients(phi[members], normalize_dictionary)
      data_term = data_term + alignment_penalty * reg
    update = stepsize * data_term
    if hessian_diagonal is not None:
      update = update / (hessian_diagonal[:, None] + lowest_code_val)
    phi = phi - update
    if normalize_dictionary:
      phi = phi / phi.norm(p=2, dim=1)[:, None]
  return phi


def hessian_running_mean(hessian_diagonal, codes):
  """training/sparse_coding.py:154 (functional)."""
  return hessian_diagonal * 0.99 + torch.pow(codes, 2).mean(0) / 100


def train_steps(batches, dictionary, sparsity_weight, num_iters, stepsize, variant='fista',
                update_rule='sc_cheap_quadratic_descent', group_assignments=None, alignment_penalty=0.0):
  """
  The per-batch body of train_dictionary (training/sparse_coding.py:513-515 with :139, :154, :168): infer codes,
  update the Hessian running mean, update the dictionary. Returns (dictionary, hessian_diagonal, last codes).
  """
  phi = dictionary.clone()
  h = dictionary.new_zeros(dictionary.size(0))
  codes = None
  for x in batches:
    if group_assignments is None:
      codes = ista_fista(x, phi, sparsity_weight, num_iters, variant=variant)
    else:
      codes = subspace_ista_fista(x, phi, group_assignments, sparsity_weight, num_iters, variant=variant)
    if update_rule.endswith('cheap_quadratic_descent'):
      h = hessian_running_mean(h, codes)
      phi = sc_dictionary_update(x, phi, codes, h, stepsize=stepsize, group_assignments=group_assignments,
                                 alignment_penalty=alignment_penalty)
    else:
      # sc_steepest_descent.py:37-41; 'subspace_sc_steepest_descent' (named at training/sparse_coding.py:421-427 but
      # absent from the reference tree) = the same step with the alignment term of the subspace cheap rule
      phi = sc_dictionary_update(x, phi, codes, None, stepsize=stepsize, group_assignments=group_assignments,
                                 alignment_penalty=alignment_penalty if update_rule.startswith('subspace') else 0.0)
  return phi, h, codes


# ---------------------------------------------------------------------------------------------------------------
# Convolutional sparse coding: images (b, c, h, w) already padded, dictionary (s, c, kh, kw),
# codes (b, s, sh, sw); synthesis = conv_transpose2d, analysis = conv2d, both with stride kernel_stride.
def get_padding_amt(image_dim, kernel_dim, dim_stride):
  """utils/convolutions.py:7-12."""
  leading = kernel_dim - dim_stride
  trailing = kernel_dim - dim_stride
  if image_dim % dim_stride != 0:
    trailing += dim_stride - (image_dim % dim_stride)
  return leading, trailing


def code_dim_from_padded_img_dim(padded_image_dim, kernel_dim, dim_stride):
  """utils/convolutions.py:14-15."""
  import math
  return 1 + int(math.ceil((padded_image_dim - kernel_dim) / dim_stride))


def create_mask(images_with_padding, padding):
  """utils/convolutions.py:17-24: ones, zero on the padded border."""
  mask = torch.ones_like(images_with_padding)
  if padding is not None:
    mask[:, :, 0:padding[0][0], :] = 0.0
    mask[:, :, -padding[0][1]:, :] = 0.0
    mask[:, :, :, 0:padding[1][0]] = 0.0
    mask[:, :, :, -padding[1][1]:] = 0.0
  return mask


def conv_ista_fista(images_padded, dictionary, kernel_stride, padding_dims, sparsity_weight, num_iters,
                    variant='fista', initial_codes=None, early_stopping_epsilon=None, nonnegative_only=False,
                    hard_threshold=False, return_iters=False):
  """analysis_transforms/convolutional/ista_fista.py:104-197."""
  assert variant in ['ista', 'fista']
  F = torch.nn.functional
  flat = torch.flatten(dictionary, start_dim=1)
  stepsize = 1. / torch.linalg.eigvalsh(torch.mm(flat, flat.t()))[-1]
  cutoff = sparsity_weight * stepsize
  sh = code_dim_from_padded_img_dim(images_padded.shape[2], dictionary.shape[2], kernel_stride[0])
  sw = code_dim_from_padded_img_dim(images_padded.shape[3], dictionary.shape[3], kernel_stride[1])
  if initial_codes is None:
    eval_pt = images_padded.new_zeros((images_padded.shape[0], dictionary.shape[0], sh, sw))
  else:
    assert tuple(initial_codes.shape) == (images_padded.shape[0], dictionary.shape[0], sh, sw)
    eval_pt = initial_codes
  previous = eval_pt.clone()
  mask = create_mask(images_padded, padding_dims)
  betas = momentum_coefficients(num_iters)
  codes, done = None, 0
  for k in range(num_iters):
    residual = mask * (F.conv_transpose2d(eval_pt, dictionary, stride=kernel_stride) - images_padded)
    codes = threshold(eval_pt - stepsize * F.conv2d(residual, dictionary, stride=kernel_stride), cutoff,
                      nonnegative_only, hard_threshold)
    change = codes - previous
    eval_pt = codes + betas[k] * change if variant == 'fista' else codes
    previous = codes
    done = k + 1
    if early_stopping_epsilon is not None:
      if bool(torch.mean(torch.abs(change) / stepsize) < early_stopping_epsilon) and k > 0:
        break
  return (codes, done) if return_iters else codes


def conv_dictionary_gradient(images_padded, dictionary, codes, kernel_stride, padding_dims):
  """The data-term gradient, summed over the batch (NOT divided by it):
  dict_update_rules/convolutional/sc_cheap_quadratic_descent.py:65-71 without the division."""
  F = torch.nn.functional
  mask = create_mask(images_padded, padding_dims)
  residual = mask * (F.conv_transpose2d(codes, dictionary, stride=kernel_stride) - images_padded)
  return F.conv2d(residual.transpose(dim0=1, dim1=0), codes.transpose(dim0=1, dim1=0),
                  dilation=kernel_stride).transpose(dim0=1, dim1=0)


def conv_sc_dictionary_update(images_padded, dictionary, codes, kernel_stride, padding_dims, hessian_diagonal=None,
                              stepsize=0.001, num_iters=1, lowest_code_val=0.001, normalize_dictionary=True,
                              batch_size=None, extra_gradient=None):
  """convolutional/sc_cheap_quadratic_descent.py:59-79 (hessian_diagonal given) / sc_steepest_descent.py:55-72 (None).
  Functional: returns the updated dictionary. batch_size / extra_gradient as in sc_dictionary_update."""
  phi = dictionary.clone()
  divisor = images_padded.shape[0] if batch_size is None else batch_size
  for it in range(num_iters):
    gradient = conv_dictionary_gradient(images_padded, phi, codes, kernel_stride, padding_dims)
    if extra_gradient is not None and it == 0:
      gradient = gradient + extra_gradient
    gradient = gradient / divisor
    if hessian_diagonal is not None:
      gradient = gradient / (hessian_diagonal[:, None, None, None] + lowest_code_val)
    gradient = gradient * (phi.norm(p=2) / gradient.norm(p=2))
    phi = phi - stepsize * gradient
    if normalize_dictionary:
      phi = phi / torch.squeeze(phi.norm(p=2, dim=(1, 2, 3)))[:, None, None, None]
  return phi


def conv_hessian_running_mean(hessian_diagonal, codes):
  """training/sparse_coding.py:158-161 (functional)."""
  return hessian_diagonal * 0.99 + torch.mean(torch.sum(codes**2, dim=(2, 3)), dim=0) / 100


def synthetic_conv_dictionary(num_kernels, channels, kh, kw, seed=1, window=0.18):
  """Unit-norm random kernels as every conv example / test initialises (examples/train_convolutional_sparse_coding.py:
  95-99), by default under a Gaussian window of width `window` (in units of the kernel size). The reference's step size
  comes from the Gram matrix of the flattened kernels, which under-estimates the Lipschitz constant of the overlapping
  strided synthesis; with plain random kernels FISTA diverges (window=0 reproduces that), windowed kernels overlap
  little and converge -- the well-posed case for parity and benchmarks."""
  g = torch.Generator().manual_seed(seed)
  phi = torch.randn(num_kernels, channels, kh, kw, generator=g)
  if window:
    yy = (torch.arange(kh) - (kh - 1) / 2)[:, None] / (window * kh)
    xx = (torch.arange(kw) - (kw - 1) / 2)[None, :] / (window * kw)
    phi = phi * torch.exp(-0.5 * (yy**2 + xx**2))
  return phi / torch.squeeze(phi.norm(p=2, dim=(1, 2, 3)))[:, None, None, None]


def synthetic_padded_images(batch, channels, height, width, kernel, stride, seed=0, std=0.3):
  """Whitened 1/f-noise images of (height, width) zero-padded by get_padding_amt on every side; returns
  (images_padded, padding_dims)."""
  g = torch.Generator().manual_seed(seed)
  fy = torch.fft.fftfreq(height)[:, None]
  fx = torch.fft.fftfreq(width)[None, :]
  rad = torch.sqrt(fy**2 + fx**2)
  amp = 1.0 / torch.clamp(rad, min=1.0 / max(height, width))
  whiten = torch.clamp(rad, min=1e-3) * torch.exp(-(rad / (0.5 * 0.9))**8)
  phase = torch.rand(batch, channels, height, width, generator=g) * 2 * torch.pi
  img = torch.fft.ifft2(amp * torch.exp(1j * phase)).real
  lo = img.amin(dim=(2, 3), keepdim=True)
  hi = img.amax(dim=(2, 3), keepdim=True)
  img = (img - lo) / (hi - lo)
  img = torch.fft.ifft2(torch.fft.fft2(img) * whiten).real
  img = (img * (std / img.std())).float()
  pv = get_padding_amt(height, kernel[0], stride[0])
  ph = get_padding_amt(width, kernel[1], stride[1])
  padded = torch.nn.functional.pad(img, (ph[0], ph[1], pv[0], pv[1]))
  return padded.contiguous(), (pv, ph)


def extract_patches(images, corners, patch_dimensions):
  """utils/dataset_generation.py:207-218 followed by the final reshape(N, -1): patch p is
  images[img, top:top+ph, left:left+pw] (channel last), flattened in (y, x, c) order. images (n, h, w, c)."""
  ph, pw = patch_dimensions
  out = images.new_zeros((corners.shape[0], ph * pw * images.shape[3]))
  for p_idx in range(corners.shape[0]):
    img, top, left = (int(v) for v in corners[p_idx])
    out[p_idx] = images[img, top:top + ph, left:left + pw].reshape(-1)
  return out


def whitening_filter(dft_num_samples, cutoffs, norm_and_threshold=True):
  """utils/image_processing.py:292-302: exponential low-pass (order 8, :216-221, un-normalised) times the ramp |f|
  (:251-264, un-normalised) rolled off at cutoffs['low']; numpy float64 (h, w)."""
  import numpy as np
  fv, fh = np.meshgrid(np.fft.fftfreq(dft_num_samples[0]), np.fft.fftfreq(dft_num_samples[1]), indexing='ij')
  mag = np.sqrt(np.square(fv) + np.square(fh))
  lpf = np.exp(-1. * np.power(mag / (0.5 * cutoffs['high']), 8.0))
  combined = np.maximum(mag, cutoffs['low'] * np.ones(mag.shape)) * lpf
  if norm_and_threshold:
    combined /= np.max(np.abs(combined))
    combined[np.abs(combined) < 1e-3] = 1e-3
  return combined


def whiten_center_surround(image, cutoffs, norm_and_threshold=True):
  """utils/image_processing.py:267-308 with filter_fd (:63-92): image ndarray (h, w, c) float32 -> float32, each colour
  channel filtered in the DFT domain."""
  import numpy as np
  filt = whitening_filter(image.shape, cutoffs, norm_and_threshold)
  out = np.zeros(image.shape, dtype='float32')
  for ch in range(image.shape[2]):
    out[:, :, ch] = np.real(np.fft.ifft2(filt * np.fft.fft2(image[:, :, ch], filt.shape),
                                         filt.shape)).astype('float32')[0:image.shape[0], 0:image.shape[1]]
  return out


def compute_psnr(target, reconstruction, manual_sig_mag=None):
  """utils/plotting.py:17-39 (numpy arrays in, numpy scalar out)."""
  import numpy as np
  sig = (np.max(target) - np.min(target)) if manual_sig_mag is None else manual_sig_mag
  mse = np.mean(np.square(target - reconstruction))
  return 10. * np.log10((sig**2) / mse) if mse != 0 else np.inf


def compute_metrics(batch_images, batch_codes, dictionary, previous_dictionary, sparsity_weight, code_inf_alg='fista',
                    group_assignments=None, kernel_strides=None, image_padding=None):
  """training/sparse_coding.py:177-229: the validation metrics of one batch, on the host with numpy like the
  reference (convolutional when kernel_strides is given: both images and reconstructions cropped to the un-padded
  region, :185-195)."""
  import numpy as np
  metrics = {}
  images_np = batch_images.numpy()
  if kernel_strides is None:
    recons = torch.mm(batch_codes, dictionary).numpy()
    axes = 1
  else:
    recons = torch.nn.functional.conv_transpose2d(batch_codes, dictionary, stride=kernel_strides).numpy()
    if image_padding is not None:
      (pt, pb), (pl, pr) = image_padding
      recons = recons[:, :, pt:-pb, pl:-pr]
      images_np = images_np[:, :, pt:-pb, pl:-pr]
    axes = (1, 2, 3)
  metrics['Average LASSO L2 component'] = np.mean(0.5 * np.sum(np.square(recons - images_np), axis=axes))
  if code_inf_alg in ('subspace_ista', 'subspace_fista'):
    group_norms = np.zeros((len(batch_codes),))
    for g in group_assignments:
      group_norms += torch.norm(batch_codes[:, g], p=2, dim=1).numpy()
    metrics['Average LASSO lagrange component'] = np.mean(sparsity_weight * group_norms)
  else:
    metrics['Average LASSO lagrange component'] = np.mean(
        sparsity_weight * torch.norm(batch_codes, p=1, dim=axes).numpy())
  metrics['Average LASSO Loss'] = (metrics['Average LASSO L2 component'] +
                                   metrics['Average LASSO lagrange component'])
  metrics['Average Normalized L0'] = float(torch.mean(
      torch.norm(batch_codes, p=0, dim=axes) / np.prod(batch_codes.shape[1:])).numpy())
  sig = np.max(images_np) - np.min(images_np)
  psnrs = [compute_psnr(images_np[b], recons[b], manual_sig_mag=sig) for b in range(recons.shape[0])]
  metrics['Average pSNR of reconstructions'] = np.mean([v for v in psnrs if v != np.inf])
  metrics['Average change in dictionary kernels'] = torch.mean(
      torch.abs(dictionary - previous_dictionary), dim=axes).numpy()
  return metrics


# ---------------------------------------------------------------------------------------------------------------
# Seeded synthetic inputs; shared by tests, smoke() and bench.py so that the CUDA path and the
# oracle always see identical tensors.
def synthetic_dictionary(num_atoms, num_pixels, seed=1):
  g = torch.Generator().manual_seed(seed)
  phi = torch.randn(num_atoms, num_pixels, generator=g)
  return phi / phi.norm(dim=1, keepdim=True)


def synthetic_patches(batch, num_pixels, seed=0, kind='gaussian', std=0.3):
  """'gaussian': x = std * randn.  'whitened': 1/f noise images passed through the reference's center-surround
  whitening filter max(|f|,1e-3) * exp(-(|f| / 0.45)^8) (utils/image_processing.py:267-308), random crops,
  globally rescaled to the requested per-pixel std."""
  g = torch.Generator().manual_seed(seed)
  if kind == 'gaussian':
    return std * torch.randn(batch, num_pixels, generator=g)
  side = int(round(num_pixels**0.5))
  assert side * side == num_pixels, 'whitened patches must be square'
  size = 256
  fy = torch.fft.fftfreq(size)[:, None]
  fx = torch.fft.fftfreq(size)[None, :]
  rad = torch.sqrt(fy**2 + fx**2)
  amp = 1.0 / torch.clamp(rad, min=1.0 / size)
  whiten = torch.clamp(rad, min=1e-3) * torch.exp(-(rad / (0.5 * 0.9))**8)
  per_image = 1024
  num_images = (batch + per_image - 1) // per_image
  out = torch.empty(num_images * per_image, num_pixels)
  for i in range(num_images):
    phase = torch.rand(size, size, generator=g) * 2 * torch.pi
    spectrum = amp * torch.exp(1j * phase)
    img = torch.fft.ifft2(spectrum).real
    img = (img - img.min()) / (img.max() - img.min())
    img = torch.fft.ifft2(torch.fft.fft2(img) * whiten).real
    ys = torch.randint(5, size - side - 5, (per_image,), generator=g)
    xs = torch.randint(5, size - side - 5, (per_image,), generator=g)
    idx_y = ys[:, None, None] + torch.arange(side)[None, :, None]
    idx_x = xs[:, None, None] + torch.arange(side)[None, None, :]
    out[i * per_image:(i + 1) * per_image] = img[idx_y, idx_x].reshape(per_image, num_pixels)
  out = out[:batch]
  return (out * (std / out.std())).float().contiguous()


def relative_l2(a, b):
  return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300))


def support_mismatches(a, b, band=0.0):
  """
  Entries whose zero / non-zero status differs between a and b, as (total, outside_band): a flip whose non-zero
  side is no larger than `band` is a tie at the threshold and is not counted in
  outside_band.
  """
  flipped = (a != 0) != (b != 0)
  magnitude = torch.maximum(a.abs(), b.abs())
  return int(flipped.sum()), int((flipped & (magnitude > band)).sum())

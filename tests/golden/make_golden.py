"""
Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on the CPU in float32.

Run in the authoring container only (the GPU box has no /root/reference):  python tests/golden/make_golden.py
Two non-invasive shims are applied before import, neither edits the reference:
  * torch.symeig (removed from torch) -> torch.linalg.eigvalsh with the same conventions
  * utils.plotting stubbed in sys.modules (it imports skimage / matplotlib, absent here) for the trainer
"""
import os
import sys
import types

import numpy as np
import torch

REF = '/root/reference/vision_transform_codes'
HERE = os.path.dirname(os.path.abspath(__file__))

torch.symeig = lambda A, eigenvectors=False, upper=True: (
    torch.linalg.eigvalsh(A, UPLO='U' if upper else 'L'), None)
sys.modules['utils.plotting'] = types.ModuleType('utils.plotting')
sys.path.insert(0, REF)

from analysis_transforms.fully_connected import ista_fista  # noqa: E402
from analysis_transforms.fully_connected import subspace_ista_fista  # noqa: E402
from dict_update_rules.fully_connected import sc_cheap_quadratic_descent  # noqa: E402
from dict_update_rules.fully_connected import sc_steepest_descent  # noqa: E402
from dict_update_rules.fully_connected import subspace_sc_cheap_quadratic_descent  # noqa: E402
from training import sparse_coding  # noqa: E402

torch.set_num_threads(4)


def dictionary(s, n, seed=1):
  g = torch.Generator().manual_seed(seed)
  phi = torch.randn(s, n, generator=g)
  return phi / phi.norm(dim=1, keepdim=True)


def patches(b, n, seed=0, std=0.3):
  g = torch.Generator().manual_seed(seed)
  return std * torch.randn(b, n, generator=g)


def save(name, **arrays):
  out = {k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrays.items()}
  np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
  print(name, {k: v.shape for k, v in out.items()})


def make_inference():
  b, n, s, T, lam = 48, 64, 128, 60, 0.1
  x, phi = patches(b, n), dictionary(s, n)
  warm = ista_fista.run(x, phi, lam, 5, variant='fista')
  save('inference_small', images=x, dictionary=phi, sparsity_weight=lam, num_iters=T, warm_start=warm,
       fista=ista_fista.run(x, phi, lam, T, variant='fista'),
       ista=ista_fista.run(x, phi, lam, T, variant='ista'),
       fista_nonneg=ista_fista.run(x, phi, lam, T, variant='fista', nonnegative_only=True),
       fista_hard=ista_fista.run(x, phi, lam, T, variant='fista', hard_threshold=True),
       ista_hard_nonneg=ista_fista.run(x, phi, lam, T, variant='ista', hard_threshold=True, nonnegative_only=True),
       fista_warm=ista_fista.run(x, phi, lam, T, variant='fista', initial_codes=warm),
       ista_early=ista_fista.run(x, phi, lam, 1000, variant='ista', early_stopping_epsilon=1e-3),
       fista_early=ista_fista.run(x, phi, lam, 1000, variant='fista', early_stopping_epsilon=1e-3))


PROX_VARIANTS = (('soft', {}), ('nonneg', {'nonnegative_only': True}), ('hard', {'hard_threshold': True}),
                 ('hard_nonneg', {'hard_threshold': True, 'nonnegative_only': True}))


def make_threshold_steps():
  """What pins the prox kernels themselves: ONE and THREE iterations of the reference, warm-started
  from the reference's own iterate a_{T-1}, for every threshold variant -- a discontinuous prox cannot be judged on a
  whole trajectory. Plus float64 runs of the same reference code for the two hard-threshold trajectories of
  inference_small.npz, so that the float32 reference's own sensitivity is on record."""
  b, n, s, T, lam = 48, 64, 128, 60, 0.1
  x, phi = patches(b, n), dictionary(s, n)
  out = {}
  for variant in ('ista', 'fista'):
    for name, kw in PROX_VARIANTS:
      warm = ista_fista.run(x, phi, lam, T - 1, variant=variant, **kw)
      out['%s_%s_warm' % (variant, name)] = warm
      out['%s_%s_one' % (variant, name)] = ista_fista.run(x, phi, lam, 1, variant=variant, initial_codes=warm, **kw)
      out['%s_%s_three' % (variant, name)] = ista_fista.run(x, phi, lam, 3, variant=variant, initial_codes=warm, **kw)
  xd, pd = x.double(), phi.double()
  out['fista_hard_f64'] = ista_fista.run(xd, pd, lam, T, variant='fista', hard_threshold=True)
  out['ista_hard_nonneg_f64'] = ista_fista.run(xd, pd, lam, T, variant='ista', hard_threshold=True,
                                               nonnegative_only=True)
  out['fista_f64'] = ista_fista.run(xd, pd, lam, T, variant='fista')
  save('threshold_steps', images=x, dictionary=phi, sparsity_weight=lam, num_iters=T, **out)


def make_conv_threshold_steps():
  from analysis_transforms.convolutional import ista_fista as conv_ista_fista
  lam, T = 0.05, 40
  x, phi, pad = conv_inputs(3, 1, (48, 40), (16, 16), (8, 8), 24)
  st = (8, 8)
  out = {}
  for variant in ('ista', 'fista'):
    for name, kw in PROX_VARIANTS:
      warm = conv_ista_fista.run(x, phi, st, pad, lam, T - 1, variant=variant, **kw)
      out['%s_%s_warm' % (variant, name)] = warm
      out['%s_%s_one' % (variant, name)] = conv_ista_fista.run(x, phi, st, pad, lam, 1, variant=variant,
                                                               initial_codes=warm, **kw)
  out['ista_hard_nonneg_f64'] = conv_ista_fista.run(x.double(), phi.double(), st, pad, lam, T, variant='ista',
                                                    nonnegative_only=True, hard_threshold=True)
  save('conv_threshold_steps', images_padded=x, dictionary=phi, stride=np.array(st), padding=np.array(pad),
       sparsity_weight=lam, num_iters=T, **out)


def make_config1():
  # BASELINE.json configs[0]: 16x16 patches, 256 atoms, batch 250, 300 iterations, lambda 0.1
  b, n, s, T, lam = 250, 256, 256, 300, 0.1
  x, phi = patches(b, n), dictionary(s, n)
  save('inference_config1', images=x, dictionary=phi, sparsity_weight=lam, num_iters=T,
       fista=ista_fista.run(x, phi, lam, T, variant='fista'))


def make_overcomplete():
  # configs[1] shape (D=256, 1024 atoms) on a 96-patch sub-batch
  b, n, s, T, lam = 96, 256, 1024, 300, 0.1
  x, phi = patches(b, n), dictionary(s, n)
  save('inference_overcomplete', images=x, dictionary=phi, sparsity_weight=lam, num_iters=T,
       fista=ista_fista.run(x, phi, lam, T, variant='fista'))


def make_subspace():
  b, n, s, T, lam = 40, 48, 64, 50, 0.1
  x, phi = patches(b, n), dictionary(s, n)
  pairs = [list(g) for g in np.array_split(np.arange(s), s // 2)]
  quads = [list(g) for g in np.array_split(np.arange(s), s // 4)]
  ragged = [[0, 2, 5], [1], [2, 3, 4, 5], [6, 7, 8, 9, 10], [11, 12], [13, 14, 15, 16, 17, 18, 19]]
  warm = subspace_ista_fista.run(x, phi, pairs, lam, 5)
  save('subspace_small', images=x, dictionary=phi, sparsity_weight=lam, num_iters=T, warm_start=warm,
       ragged_sizes=np.array([len(g) for g in ragged]), ragged_flat=np.concatenate([np.array(g) for g in ragged]),
       pairs_fista=subspace_ista_fista.run(x, phi, pairs, lam, T, variant='fista'),
       pairs_ista=subspace_ista_fista.run(x, phi, pairs, lam, T, variant='ista'),
       quads_fista=subspace_ista_fista.run(x, phi, quads, lam, T, variant='fista'),
       ragged_fista=subspace_ista_fista.run(x, phi, ragged, lam, T, variant='fista'),
       pairs_warm=subspace_ista_fista.run(x, phi, pairs, lam, T, variant='fista', initial_codes=warm),
       pairs_early=subspace_ista_fista.run(x, phi, pairs, lam, 1000, variant='ista', early_stopping_epsilon=1e-3))


def make_dict_update():
  b, n, s, lam = 96, 64, 128, 0.1
  x, phi = patches(b, n), dictionary(s, n)
  codes = ista_fista.run(x, phi, lam, 40)
  h = torch.pow(codes, 2).mean(0) / 100
  d1 = phi.clone()
  sc_cheap_quadratic_descent.run(x, d1, codes, h, stepsize=0.1, num_iters=1)
  d2 = phi.clone()
  sc_cheap_quadratic_descent.run(x, d2, codes, h, stepsize=0.05, num_iters=3)
  d3 = phi.clone()
  sc_steepest_descent.run(x, d3, codes, stepsize=0.1, num_iters=1)
  d4 = phi.clone()
  sc_steepest_descent.run(x, d4, codes, stepsize=0.1, num_iters=2, normalize_dictionary=False)
  pairs = [list(map(int, g)) for g in np.array_split(np.arange(s), s // 2)]
  overlapping = [[0, 2, 5], [1, 7], [2, 3, 4, 5], [6, 7, 8, 9, 10], [11, 12]]
  d5 = phi.clone()
  subspace_sc_cheap_quadratic_descent.run(x, d5, codes, pairs, h, 0.5, stepsize=0.1, num_iters=1)
  d6 = phi.clone()
  subspace_sc_cheap_quadratic_descent.run(x, d6, codes, overlapping, h, 0.25, stepsize=0.05, num_iters=2)
  d7 = phi.clone() * torch.linspace(0.5, 2.0, s)[:, None]
  d7_in = d7.clone()
  subspace_sc_cheap_quadratic_descent.run(x, d7, codes, overlapping, h, 0.25, stepsize=0.05, num_iters=1,
                                          normalize_dictionary=False)
  save('dict_update_small', images=x, dictionary=phi, codes=codes, hessian_diagonal=h,
       cheap_1=d1, cheap_3=d2, steepest_1=d3, steepest_2_unnormalized=d4,
       aligned_pairs=d5, aligned_overlapping_2=d6, unnormalized_in=d7_in, aligned_unnormalized=d7)


def make_training():
  # 4 steps of the unmodified train_dictionary: fista(30) + cheap quadratic descent
  nb, b, n, s = 4, 64, 64, 128
  x = patches(nb * b, n).view(nb, b, n)
  phi0 = dictionary(s, n)
  params = {
      'mode': 'fully-connected', 'num_epochs': 1,
      'code_inference_algorithm': 'fista',
      'inference_param_schedule': {0: {'sparsity_weight': 0.1, 'num_iters': 30}},
      'dictionary_update_algorithm': 'sc_cheap_quadratic_descent',
      'dict_update_param_schedule': {0: {'stepsize': 0.1, 'num_iters': 1}}}
  phi = phi0.clone()
  sparse_coding.train_dictionary(x, x[:1], phi, params)
  params2 = dict(params, code_inference_algorithm='ista', dictionary_update_algorithm='sc_steepest_descent')
  phi_b = phi0.clone()
  sparse_coding.train_dictionary(x, x[:1], phi_b, params2)
  pairs = [list(map(int, g)) for g in np.array_split(np.arange(s), s // 2)]
  params3 = dict(params, code_inference_algorithm='subspace_fista',
                 dictionary_update_algorithm='subspace_sc_cheap_quadratic_descent',
                 group_assignments=pairs, subspace_alignment_penalty=0.0)
  phi_c = phi0.clone()
  sparse_coding.train_dictionary(x, x[:1], phi_c, params3)
  params4 = dict(params3, subspace_alignment_penalty=0.3)
  phi_d = phi0.clone()
  sparse_coding.train_dictionary(x, x[:1], phi_d, params4)
  save('training_small', batches=x, dictionary=phi0, fista_cheap=phi, ista_steepest=phi_b, subspace_cheap=phi_c,
       subspace_cheap_aligned=phi_d)


def conv_inputs(b, c, hw, k, stride, s, seed=0, std=0.3, window=0.18):
  """Seeded padded images + unit-norm kernels, padded as utils.convolutions.get_padding_amt prescribes.
  window > 0 multiplies the random kernels by a Gaussian of that width (in units of the kernel size): the reference
  takes its step size from the Gram matrix of the FLATTENED kernels, which under-estimates the Lipschitz constant of the
  strided, overlapping synthesis (2.58 against 1.59 for 24 plain random 16x16 kernels at stride 8) -- FISTA then
  diverges and amplifies rounding noise without bound, which makes a useless parity case. Windowed kernels overlap
  little (ratio 1.05), the iteration converges; ISTA is stable either way and is also pinned on plain kernels."""
  from utils.convolutions import get_padding_amt
  g = torch.Generator().manual_seed(seed)
  img = std * torch.randn(b, c, hw[0], hw[1], generator=g)
  pv = get_padding_amt(hw[0], k[0], stride[0])
  ph = get_padding_amt(hw[1], k[1], stride[1])
  x = torch.nn.functional.pad(img, (ph[0], ph[1], pv[0], pv[1])).contiguous()
  g2 = torch.Generator().manual_seed(seed + 1)
  phi = torch.randn(s, c, k[0], k[1], generator=g2)
  if window:
    yy = (torch.arange(k[0]) - (k[0] - 1) / 2)[:, None] / (window * k[0])
    xx = (torch.arange(k[1]) - (k[1] - 1) / 2)[None, :] / (window * k[1])
    phi = phi * torch.exp(-0.5 * (yy**2 + xx**2))
  phi = phi / torch.squeeze(phi.norm(p=2, dim=(1, 2, 3)))[:, None, None, None]
  return x, phi, (pv, ph)


def make_conv():
  """Convolutional path: the call matrix of the reference's tests/ista_fista_2.py, the two conv dictionary updates
  and a short conv training run (tests/sparse_coding_4.py), on small seeded inputs."""
  from analysis_transforms.convolutional import ista_fista as conv_ista_fista
  from dict_update_rules.convolutional import sc_cheap_quadratic_descent as conv_cheap
  from dict_update_rules.convolutional import sc_steepest_descent as conv_steepest
  lam, T = 0.05, 40
  # (a) the shape family of BASELINE configs[4]: 1 channel, 16x16 kernels, stride 8 (here 24 kernels on 48x40 images)
  x, phi, pad = conv_inputs(3, 1, (48, 40), (16, 16), (8, 8), 24)
  st = (8, 8)
  warm = conv_ista_fista.run(x, phi, st, pad, lam, 4, variant='fista')
  codes = conv_ista_fista.run(x, phi, st, pad, lam, T, variant='fista')
  h = torch.mean(torch.sum(codes**2, dim=(2, 3)), dim=0) / 100

  def updated(mod, *args, **kw):
    d = phi.clone()
    mod.run(x, d, codes, *args, **kw)
    return d

  save('conv_small', images_padded=x, dictionary=phi, stride=np.array(st), padding=np.array(pad), sparsity_weight=lam,
       num_iters=T, warm_start=warm, fista=codes,
       ista=conv_ista_fista.run(x, phi, st, pad, lam, T, variant='ista'),
       plain_dictionary=conv_inputs(3, 1, (48, 40), (16, 16), (8, 8), 24, window=0)[1],
       ista_plain=conv_ista_fista.run(x, conv_inputs(3, 1, (48, 40), (16, 16), (8, 8), 24, window=0)[1], st, pad, lam, T,
                                      variant='ista'),
       ista_early=conv_ista_fista.run(x, phi, st, pad, lam, 500, variant='ista', early_stopping_epsilon=1e-3),
       fista_nonneg=conv_ista_fista.run(x, phi, st, pad, lam, T, variant='fista', nonnegative_only=True),
       ista_hard_nonneg=conv_ista_fista.run(x, phi, st, pad, lam, T, variant='ista', nonnegative_only=True,
                                            hard_threshold=True),
       fista_warm=conv_ista_fista.run(x, phi, st, pad, lam, T, variant='fista', initial_codes=warm),
       hessian_diagonal=h,
       cheap_1=updated(conv_cheap, h, st, pad, stepsize=0.05),
       cheap_2=updated(conv_cheap, h, st, pad, stepsize=0.02, num_iters=2),
       steepest_1=updated(conv_steepest, st, pad, stepsize=0.05),
       steepest_unnormalized=updated(conv_steepest, st, pad, stepsize=0.05, normalize_dictionary=False))
  # (b) two channels, rectangular kernels and strides, image size not a multiple of the stride
  x2, phi2, pad2 = conv_inputs(2, 2, (21, 30), (8, 12), (4, 6), 10, seed=5)
  st2 = (4, 6)
  codes2 = conv_ista_fista.run(x2, phi2, st2, pad2, lam, T, variant='fista')
  h2 = torch.mean(torch.sum(codes2**2, dim=(2, 3)), dim=0) / 100
  d2 = phi2.clone()
  conv_cheap.run(x2, d2, codes2, h2, st2, pad2, stepsize=0.05)
  save('conv_two_channel', images_padded=x2, dictionary=phi2, stride=np.array(st2), padding=np.array(pad2),
       sparsity_weight=lam, num_iters=T, fista=codes2, hessian_diagonal=h2, cheap_1=d2)
  # (c) the unmodified trainer in convolutional mode (tests/sparse_coding_4.py): ista + cheap quadratic descent
  nb, b = 3, 2
  xb, phi0, padb = conv_inputs(nb * b, 1, (32, 32), (16, 16), (8, 8), 16, seed=9)
  xb = xb.view(nb, b, *xb.shape[1:])
  params = {
      'mode': 'convolutional', 'num_epochs': 1, 'strides': (8, 8), 'padding': padb,
      'code_inference_algorithm': 'ista',
      'inference_param_schedule': {0: {'sparsity_weight': 0.05, 'num_iters': 15}},
      'dictionary_update_algorithm': 'sc_cheap_quadratic_descent',
      'dict_update_param_schedule': {0: {'stepsize': 0.05, 'num_iters': 1}}}
  d = phi0.clone()
  sparse_coding.train_dictionary(xb, xb[:1], d, params)
  d_b = phi0.clone()
  sparse_coding.train_dictionary(xb, xb[:1], d_b, dict(params, code_inference_algorithm='fista',
                                                        dictionary_update_algorithm='sc_steepest_descent'))
  save('conv_training_small', batches=xb, dictionary=phi0, padding=np.array(padb), ista_cheap=d, fista_steepest=d_b)


def _reference_function(path, name):
  """Executes ONE function definition of a reference source file (here utils/plotting.py, whose module-level imports
  need skimage / matplotlib) and returns it: the reference's own code, unmodified, without importing the module."""
  import ast
  src = open(os.path.join(REF, path)).read()
  node = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == name][0]
  scope = {'np': np}
  exec(compile(ast.Module(body=[node], type_ignores=[]), path, 'exec'), scope)
  return scope[name]


def make_metrics():
  """Validation metrics (training/sparse_coding.py:177-229): the unmodified trainer with a
  'training_visualization_schedule', tensorboard's SummaryWriter replaced by a recorder, the dictionary figures
  (matplotlib, absent here) by an empty list, compute_pSNR executed from the reference's own source."""
  import pathlib
  import pickle
  import tempfile
  recorded = []

  class Recorder:
    def __init__(self, *a, **k):
      pass

    def add_scalar(self, tag, value, step):
      recorded.append((tag, float(value), int(step)))

    def add_image(self, *a, **k):
      pass

  tb = types.ModuleType('torch.utils.tensorboard')
  tb.SummaryWriter = Recorder
  sys.modules['torch.utils.tensorboard'] = tb
  mpl = types.ModuleType('matplotlib')
  mpl.pyplot = types.ModuleType('matplotlib.pyplot')
  sys.modules.setdefault('matplotlib', mpl)
  sys.modules.setdefault('matplotlib.pyplot', mpl.pyplot)
  plotting = sys.modules['utils.plotting']
  plotting.compute_pSNR = _reference_function('utils/plotting.py', 'compute_pSNR')
  plotting.display_dictionary = lambda *a, **k: []

  def run(name, train, val, phi0, params):
    del recorded[:]
    with tempfile.TemporaryDirectory() as tmp:
      log = pathlib.Path(tmp) / 'log'
      phi = phi0.clone()
      sparse_coding.train_dictionary(train, val, phi, dict(
          params, training_visualization_schedule={0, 2}, checkpoint_schedule={1, 2}, logging_folder_fullpath=log))
      d1 = pickle.load(open(log / 'checkpoint_dictionary_iter_1', 'rb'))
      d2 = pickle.load(open(log / 'checkpoint_dictionary_iter_2', 'rb'))
    tags = sorted({t for t, _, _ in recorded})
    table = np.array([[dict((t, v) for t, v, k in recorded if k == step)[tag] for tag in tags] for step in (0, 2)])
    extra = {}
    if 'padding' in params:
      extra = {'padding': np.array(params['padding']), 'stride': np.array(params['strides'])}
    save(name, training=train, validation=val, dictionary=phi0, dictionary_iter_1=d1, dictionary_iter_2=d2, metric_names=np.array(tags),
         metrics=table, sparsity_weight=params['inference_param_schedule'][0]['sparsity_weight'],
         num_iters=params['inference_param_schedule'][0]['num_iters'], **extra)

  nb, b, n, s = 3, 48, 64, 128
  x = patches((nb + 2) * b, n).view(nb + 2, b, n)
  phi0 = dictionary(s, n)
  params = {
      'mode': 'fully-connected', 'num_epochs': 1,
      'code_inference_algorithm': 'fista',
      'inference_param_schedule': {0: {'sparsity_weight': 0.1, 'num_iters': 40}},
      'dictionary_update_algorithm': 'sc_cheap_quadratic_descent',
      'dict_update_param_schedule': {0: {'stepsize': 0.1, 'num_iters': 1}}}
  run('metrics_fc', x[:nb], x[nb:], phi0, params)
  pairs = [list(map(int, g)) for g in np.array_split(np.arange(s), s // 2)]
  run('metrics_subspace', x[:nb], x[nb:], phi0, dict(
      params, code_inference_algorithm='subspace_fista', group_assignments=pairs,
      dictionary_update_algorithm='subspace_sc_cheap_quadratic_descent', subspace_alignment_penalty=0.0))
  xb, phic, padb = conv_inputs((nb + 2) * 2, 1, (32, 32), (16, 16), (8, 8), 16, seed=9)
  xb = xb.view(nb + 2, 2, *xb.shape[1:])
  run('metrics_conv', xb[:nb], xb[nb:], phic, {
      'mode': 'convolutional', 'num_epochs': 1, 'strides': (8, 8), 'padding': padb,
      'code_inference_algorithm': 'ista',
      'inference_param_schedule': {0: {'sparsity_weight': 0.05, 'num_iters': 15}},
      'dictionary_update_algorithm': 'sc_cheap_quadratic_descent',
      'dict_update_param_schedule': {0: {'stepsize': 0.05, 'num_iters': 1}}})


def make_whitening():
  """utils/image_processing.py whiten_center_surround (the module imports matplotlib at the top: stubbed) on two small
  seeded images, with the dataset defaults (cutoffs low 1e-3, high 0.9) and an un-normalised variant."""
  mpl = types.ModuleType('matplotlib')
  mpl.pyplot = types.ModuleType('matplotlib.pyplot')
  sys.modules.setdefault('matplotlib', mpl)
  sys.modules.setdefault('matplotlib.pyplot', mpl.pyplot)
  from utils import image_processing
  rng = np.random.RandomState(4)
  gray = rng.rand(40, 52, 1).astype('float32')
  colour = rng.rand(33, 24, 3).astype('float32')   # odd height: both fftfreq conventions
  cut = {'low': 1e-3, 'high': 0.9}
  g_out, g_filt = image_processing.whiten_center_surround(gray, cut, return_filter=True)
  c_out, c_filt = image_processing.whiten_center_surround(colour, cut, return_filter=True)
  raw_out, raw_filt = image_processing.whiten_center_surround(gray, {'low': 0.05, 'high': 0.6}, return_filter=True,
                                                              norm_and_threshold=False)
  save('whitening_small', gray=gray, gray_whitened=g_out, gray_filter=np.real(g_filt), colour=colour,
       colour_whitened=c_out, colour_filter=np.real(c_filt), raw_whitened=raw_out, raw_filter=np.real(raw_filt))


if __name__ == '__main__':
  which = sys.argv[1:] or ['inference', 'threshold_steps', 'conv_threshold_steps', 'config1', 'overcomplete', 'subspace', 'dict_update', 'training', 'conv', 'metrics', 'whitening']
  for name in which:
    globals()['make_' + name]()

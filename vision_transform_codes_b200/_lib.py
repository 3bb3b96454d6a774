"""ctypes binding of lib/libvtc_b200.so (C ABI in include/vtc_b200.h). Plumbing only: pointers, streams, errors."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# (VTC_B200_LIB: A/B timing of another build of the same library, tools/ab_build.sh)
LIB_PATH = os.environ.get('VTC_B200_LIB') or os.path.join(_HERE, 'lib', 'libvtc_b200.so')

VTC_OK, VTC_ERR_ARG, VTC_ERR_CUDA, VTC_ERR_WORKSPACE, VTC_ERR_UNSUPPORTED, VTC_ERR_NONFINITE = range(6)

_c = ctypes
_i64, _int, _f32, _f64, _ptr, _size = _c.c_int64, _c.c_int, _c.c_float, _c.c_double, _c.c_void_p, _c.c_size_t

# name -> (restype, argtypes); must list every symbol declared in include/vtc_b200.h
SIGNATURES = {
    'vtc_version': (_int, []),
    'vtc_last_error': (_c.c_char_p, []),
    'vtc_launch_count': (_c.c_longlong, []),
    'vtc_profile_enable': (_int, [_int]),
    'vtc_profile_last': (_int, [_c.POINTER(_f32), _c.POINTER(_f32), _c.POINTER(_int), _c.POINTER(_int),
                                _c.POINTER(_f32), _c.POINTER(_f32)]),
    'vtc_profile_history_count': (_int, []),
    'vtc_profile_history': (_int, [_int] + [_c.POINTER(_f32)] * 4),
    'vtc_set_formulation': (_int, [_int]),
    'vtc_get_formulation': (_int, [_i64, _i64]),
    'vtc_set_fused_iteration': (_int, [_int]),
    'vtc_set_small_batch_kernel': (_int, [_int]),
    'vtc_get_fused_iteration': (_int, [_i64, _i64, _int]),
    'vtc_debug_iter_trace': (_int, [_ptr]),
    'vtc_get_chains': (_int, [_i64, _i64, _i64]),
    'vtc_device_info': (_int, [_c.POINTER(_int)] * 3),
    'vtc_fista_workspace_bytes': (_size, [_i64, _i64, _i64, _int]),
    'vtc_fista_fc': (_int, [_ptr, _i64, _ptr, _ptr, _ptr, _i64, _i64, _i64, _i64, _f32, _int, _int, _int, _int, _int,
                            _f32, _int, _ptr, _size, _c.POINTER(_int), _c.POINTER(_f32), _ptr]),
    'vtc_dict_grad_workspace_bytes': (_size, [_i64, _i64, _i64, _int]),
    'vtc_sc_dict_grad': (_int, [_ptr, _i64, _ptr, _ptr, _i64, _ptr, _i64, _i64, _i64, _int, _ptr, _size, _ptr]),
    'vtc_sc_dict_apply': (_int, [_ptr, _ptr, _ptr, _ptr, _f32, _i64, _i64, _i64, _f32, _f32, _int, _ptr]),
    'vtc_subspace_alignment_grad': (_int, [_ptr, _i64, _i64, _ptr, _i64, _i64, _int, _ptr, _ptr]),
    'vtc_hessian_diag_update': (_int, [_ptr, _i64, _i64, _i64, _i64, _ptr, _ptr, _int, _ptr]),
    'vtc_hessian_ema': (_int, [_ptr, _ptr, _i64, _i64, _ptr]),
    'vtc_matmul_nt_workspace_bytes': (_size, [_i64, _i64, _i64, _int]),
    'vtc_matmul_nt': (_int, [_ptr, _ptr, _ptr, _ptr, _i64, _i64, _i64, _int, _ptr, _size, _ptr]),
    'vtc_lipschitz_workspace_bytes': (_size, [_i64, _i64]),
    'vtc_lipschitz': (_int, [_ptr, _i64, _i64, _ptr, _ptr, _size, _ptr]),
    'vtc_gather_rows': (_int, [_ptr, _i64, _ptr, _i64, _i64, _ptr, _ptr]),
    'vtc_gather_cols': (_int, [_ptr, _i64, _ptr, _i64, _i64, _ptr, _i64, _ptr]),
    'vtc_scatter_add_cols': (_int, [_ptr, _i64, _ptr, _i64, _i64, _ptr, _i64, _i64, _ptr]),
    'vtc_fista_conv_workspace_bytes': (_size, [_i64] * 9 + [_int]),
    'vtc_fista_conv': (_int, [_ptr, _ptr, _ptr, _ptr] + [_i64] * 9 + [_int] * 4 + [_f32, _int, _int, _int, _int, _f32,
                              _int, _ptr, _size, _c.POINTER(_int), _c.POINTER(_f32), _ptr]),
    'vtc_conv_dict_grad_workspace_bytes': (_size, [_i64] * 9 + [_int]),
    'vtc_sc_conv_dict_grad': (_int, [_ptr, _ptr, _ptr, _ptr] + [_i64] * 9 + [_int] * 4 + [_int, _ptr, _size, _ptr]),
    'vtc_sc_conv_dict_apply': (_int, [_ptr, _ptr, _ptr, _i64, _i64, _i64, _f32, _f32, _int, _ptr]),
    'vtc_conv_hessian_diag_update': (_int, [_ptr, _i64, _i64, _i64, _i64, _ptr, _ptr, _int, _ptr]),
    'vtc_sc_metrics_workspace_bytes': (_size, [_i64, _i64, _i64, _int]),
    'vtc_sc_metrics': (_int, [_ptr, _i64, _ptr, _ptr, _i64, _i64, _i64, _i64, _ptr, _i64, _i64, _int, _ptr, _ptr, _size,
                              _ptr]),
    'vtc_sc_conv_metrics_workspace_bytes': (_size, [_i64] * 9 + [_int]),
    'vtc_sc_conv_metrics': (_int, [_ptr, _ptr, _ptr] + [_i64] * 9 + [_int] * 4 + [_int, _ptr, _ptr, _size, _ptr]),
    'vtc_dict_change': (_int, [_ptr, _ptr, _i64, _i64, _ptr, _ptr]),
    'vtc_whitening_filter': (_int, [_i64, _i64, _f64, _f64, _f64, _int, _ptr, _ptr, _ptr]),
    'vtc_spectrum_filter': (_int, [_ptr, _i64, _i64, _i64, _ptr, _ptr]),
    'vtc_extract_patches': (_int, [_ptr, _i64, _i64, _i64, _i64, _ptr, _i64, _i64, _i64, _ptr, _i64, _ptr]),
}

_lib = None


def load():
  """Load the shared library (once). Fails loudly: there is no Python/PyTorch implementation to fall back to."""
  global _lib
  if _lib is None:
    if not os.path.exists(LIB_PATH):
      raise ImportError(
          'vision_transform_codes_b200: %s is missing. Build it with `python -c "import __graft_entry__ as g; '
          'g.build()"` (nvcc, sm_100a). There is no CPU fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
      fn = getattr(lib, name)
      fn.restype = restype
      fn.argtypes = argtypes
    _lib = lib
  return _lib


class VtcError(RuntimeError):
  pass


def check(rc):
  """Translate a C return code into the exception the reference would raise for the same condition."""
  if rc == VTC_OK:
    return
  msg = load().vtc_last_error().decode('utf-8', 'replace')
  if rc == VTC_ERR_ARG:
    raise ValueError(msg)
  if rc == VTC_ERR_UNSUPPORTED:
    raise NotImplementedError(msg)
  if rc == VTC_ERR_NONFINITE:
    # ista_fista.py:75-79 prints the atom norms and raises a bare RuntimeError
    raise RuntimeError(msg)
  raise VtcError('vtc_b200 error %d: %s' % (rc, msg))


def require_cuda_f32(t, name):
  if not isinstance(t, torch.Tensor):
    raise TypeError('%s must be a torch.Tensor' % name)
  if not t.is_cuda:
    raise RuntimeError('%s lives on %s: vision_transform_codes_b200 only runs on a CUDA (sm_100) device and has '
                       'no CPU fallback' % (name, t.device))
  if t.dtype != torch.float32:
    raise TypeError('%s must be float32, got %s' % (name, t.dtype))


def row_major(t):
  """Return (tensor, row pitch in elements) for a 2-D tensor whose rows are contiguous; copies only if they are not."""
  if t.dim() != 2:
    raise ValueError('expected a 2-D tensor, got shape %s' % (tuple(t.shape),))
  if t.stride(1) != 1 or t.stride(0) < t.size(1):
    t = t.contiguous()
  return t, t.stride(0) if t.size(0) > 1 else max(t.size(1), t.stride(0))


_workspaces = {}


WORKSPACE_SHRINK_FACTOR = 4          # a cached buffer this many times the request (and over the floor) is given back
WORKSPACE_SHRINK_FLOOR = 256 << 20


def workspace(nbytes, device, tag):
  """Scratch buffer per (device, current stream, tag); the library never allocates device memory itself.
  Keyed by stream so that calls enqueued on different streams never share scratch memory (calls on one stream are
  ordered, so one buffer per stream is enough). The buffer grows to the largest request and is reused; one large call
  does not pin its gigabytes for ever, though: a buffer several times larger than what is asked for goes back to
  torch's allocator (where other tensors can use it) and a fitting one takes its place."""
  key = (device.index, torch.cuda.current_stream(device).cuda_stream, tag)
  buf = _workspaces.get(key)
  if buf is not None and buf.numel() > WORKSPACE_SHRINK_FLOOR and buf.numel() > WORKSPACE_SHRINK_FACTOR * nbytes:
    buf = None
  if buf is None or buf.numel() < nbytes:
    _workspaces.pop(key, None)
    buf = None
    buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
    _workspaces[key] = buf
  return buf


def release_workspaces():
  _workspaces.clear()


def stream_ptr(device):
  return torch.cuda.current_stream(device).cuda_stream


def ptr(t):
  return 0 if t is None else t.data_ptr()

"""ctypes binding of libcpg_b200.so (C ABI declared in include/cpg_b200.h).

There is no CPU fallback: importing works anywhere (so that CPU-only tooling can
introspect the package), but the first call that needs the library raises
`CpgLibraryError` if the shared object is missing, was not built for this
machine, or no CUDA device is present.
"""
import ctypes
import os
import threading
from ctypes import (POINTER, Structure, byref, c_char_p, c_float, c_int, c_int64, c_uint32, c_uint64, c_void_p)

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, 'libcpg_b200.so')
_lock = threading.Lock()
_lib = None
_ctxs = {}

N_PARAM_TENSORS = 19
SC_COUNT = 16
SC = dict(loss=0, recon=1, kl=2, mmd=3, mmdrf=4, logvar_l1=5, logvar_kl=6, z_mu_l1=7, z_logvar=8,
          beta=9, grad_norm=10, ntok=11, nll_sum=12)
ZREGU = {'kl': 0, 'mmd': 1, 'mmdrf': 2}


class CpgLibraryError(RuntimeError):
    pass


class WaeInputs(Structure):
    _fields_ = [('tokens', c_void_p), ('eps', c_void_p), ('c', c_void_p), ('word_drop', c_void_p),
                ('out_keep', c_void_p), ('p_out_dropout', c_float)]


class TrainHparams(Structure):
    _fields_ = [('lr', c_float), ('beta1', c_float), ('beta2', c_float), ('adam_eps', c_float),
                ('clip_norm', c_float), ('beta', c_float), ('lambda_logvar_l1', c_float),
                ('lambda_logvar_kl', c_float), ('z_regu', c_int), ('mmd_sigma', c_float), ('rf_dim', c_int),
                ('compute_full_mmd', c_int), ('adam_step', c_int), ('global_batch', c_int)]


class StepNoiseBuffers(Structure):
    _fields_ = [('eps', c_void_p), ('c', c_void_p), ('word_drop', c_void_p), ('out_keep', c_void_p),
                ('z_prior_full', c_void_p), ('z_prior_rf', c_void_p), ('rf_w', c_void_p), ('rf_b', c_void_p)]


class LossNoise(Structure):
    _fields_ = [('z_prior_full', c_void_p), ('z_prior_rf', c_void_p), ('rf_w', c_void_p), ('rf_b', c_void_p)]


def _require_cuda():
    if not torch.cuda.is_available():
        raise CpgLibraryError('cpg_b200 needs a CUDA device (sm_100a); torch.cuda.is_available() is False '
                              'and there is no CPU fallback')


def _signatures(L):
    P = c_void_p
    I, F, I64 = c_int, c_float, c_int64
    sig = {
        'cpg_abi_version': (I, []),
        'cpg_last_error': (c_char_p, []),
        'cpg_create': (I, [POINTER(c_void_p), I]),
        'cpg_destroy': (I, [P]),
        'cpg_sm_count': (I, [P]),
        'cpg_workspace_bytes': (I64, [P]),
        'cpg_launch_count': (I64, [P]),
        'cpg_stash_generation': (I64, [P]),
        'cpg_check_errors': (I, [P, P]),
        'cpg_vae_param_count': (I64, [I]),
        'cpg_vae_param_layout': (I, [I, POINTER(I64), POINTER(I64)]),
        'cpg_wae_forward': (I, [P, P, P, I, I, I, POINTER(WaeInputs), P, P, P, P, I]),
        'cpg_wae_encode': (I, [P, P, P, I, I, I, P, P, P]),
        'cpg_wae_decode_teacher': (I, [P, P, P, I, I, I, POINTER(WaeInputs), P, P]),
        'cpg_wae_backward': (I, [P, P, P, I, I, I, POINTER(WaeInputs), P, P, P, P, P]),
        'cpg_wae_train_step': (I, [P, P, P, P, P, P, I, I, I, POINTER(WaeInputs), POINTER(LossNoise),
                                   POINTER(TrainHparams), P, P, P, P, P]),
        'cpg_wae_train_step_philox': (I, [P, P, P, P, P, P, I, I, I, P, POINTER(StepNoiseBuffers), POINTER(TrainHparams),
                                          c_uint64, c_uint32, F, F, P]),
        'cpg_coupled_count': (I64, [I]),
        'cpg_wae_step_phase1': (I, [P, P, P, I, I, I, POINTER(WaeInputs), POINTER(LossNoise),
                                    POINTER(TrainHparams), P, P, P, P]),
        'cpg_wae_step_phase2': (I, [P, P, P, P, I, I, I, POINTER(WaeInputs), POINTER(LossNoise),
                                    POINTER(TrainHparams), P, P, P]),
        'cpg_clip_adam_step': (I, [P, P, P, P, P, P, I, POINTER(TrainHparams), P]),
        'cpg_side_stream': (P, [P]),
        'cpg_aux_stream': (P, [P]),
        'cpg_step_dyn_write': (I, [P, P, POINTER(TrainHparams), c_uint32]),
        'cpg_step_dyn_use': (I, [P, I]),
        'cpg_dp_tail_count': (I, []),
        'cpg_dp_pack_tail': (I, [P, P, P]),
        'cpg_dp_apply_tail': (I, [P, P, P, P]),
        'cpg_softmax_xent': (I, [P, P, P, P, I, I, I, P, P]),
        'cpg_latent_stats': (I, [P, P, P, P, I, P]),
        'cpg_mmd_full': (I, [P, P, P, P, I, F, P]),
        'cpg_mmd_full_grad': (I, [P, P, P, P, I, F, P]),
        'cpg_mmd_rf': (I, [P, P, P, P, P, P, I, I, F, P, P]),
        'cpg_fill_step_noise': (I, [P, P, c_uint64, c_uint32, I, I, F, F, P, P, P, P, P, P]),
        'cpg_fill_step_noise_overlapped': (I, [P, P, c_uint64, c_uint32, I, I, F, F, P, P, P, P, P, P]),
        'cpg_fill_normal': (I, [P, P, c_uint64, c_uint32, I64, P]),
        'cpg_fill_uniform': (I, [P, P, c_uint64, c_uint32, F, I64, P]),
        'cpg_beam_decode': (I, [P, P, P, I, I, I, P, P, I, I, P, P, P]),
        'cpg_sample_decode': (I, [P, P, P, I, I, I, P, P, I, F, c_uint64, P, P]),
        'cpg_soft_decode': (I, [P, P, P, I, I, I, P, P, I, F, c_uint64, P, P, P]),
        'cpg_flow_forward': (I, [P, P, P, I, I, P, P, P, P, P, I, P, P]),
        'cpg_cnn_classifier_fwd': (I, [P, P, P, P, P, P, P, P, P, P, P, I, I, I, P, P, P]),
        'cpg_class_score_accept': (I, [P, P, P, P, I64, I, P, P, P, P, P, P, P]),
        'cpg_class_sample': (I, [P, P, P, P, P, I, I, P, P, P, P, c_uint64, I64, I64, P, P, P, P, P, P]),
        'cpg_class_regen': (I, [P, P, P, P, P, I, I, P, P, P, P, c_uint64, P, I64, P, P, P]),
        'cpg_compact_accepted': (I, [P, P, P, I64, I64, I64, P, P]),
        'cpg_gather_rows': (I, [P, P, P, P, I64, I64, I, P]),
        'cpg_dedup_rows': (I, [P, P, P, I64, I, P, P]),
        'cpg_peptide_descriptors': (I, [P, P, P, I64, I, P, I, P, P, ctypes.c_double, F, P, P, P, P]),
        'cpg_feed_batch': (I, [P, P, P, P, I64, I, c_uint64, c_uint64, I, P, P]),
        'cpg_gmm_logpdf': (I, [P, P, P, I64, P, P, P, I, P]),
        'cpg_prior_logpdf': (I, [P, P, P, I64, P]),
        'cpg_gmm_em_step': (I, [P, P, P, I64, I, P, P, P, ctypes.c_double, P, P, P, P, P]),
        'cpg_logreg_newton_stats': (I, [P, P, P, P, I64, P, P]),
        'cpg_logreg_stats_len': (I, []),
        'cpg_set_option': (I, [c_char_p, I]),
        'cpg_profile_enable': (I, [I]),
        'cpg_profile_read': (I, [P, I, P, P, I]),
        'cpg_profile_timeline': (I, [P, I, P, P, I]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    return sig


EXPORTED_SYMBOLS = None   # filled by lib(); tests compare it with include/cpg_b200.h


def lib():
    """The loaded shared library (loads on first use; fails loudly)."""
    global _lib, EXPORTED_SYMBOLS
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.isfile(_LIB_PATH):
                raise CpgLibraryError('%s not found: build it with `python __graft_entry__.py build` '
                                      '(nvcc, sm_100a); there is no CPU fallback' % _LIB_PATH)
            L = ctypes.CDLL(_LIB_PATH)
            EXPORTED_SYMBOLS = sorted(_signatures(L))
            if L.cpg_abi_version() != 1:
                raise CpgLibraryError('ABI version mismatch')
            _lib = L
    return _lib


def last_error():
    return lib().cpg_last_error().decode('utf-8', 'replace')


def check(rc, what=''):
    if rc != 0:
        raise CpgLibraryError('%s failed (code %d): %s' % (what or 'cpg call', rc, last_error()))


def context(device=None):
    """One cpg_ctx per CUDA device (workspace owner)."""
    _require_cuda()
    idx = _device_index(device)
    h = _ctxs.get(idx)
    if h is None:
        out = c_void_p()
        check(lib().cpg_create(byref(out), idx), 'cpg_create')
        h = _ctxs[idx] = out
    return h


def _device_index(device=None):
    if device is None:
        return torch.cuda.current_device()
    device = torch.device(device)
    return device.index if device.index is not None else torch.cuda.current_device()


def tensor_device(device=None):
    _require_cuda()
    return torch.device('cuda', _device_index(device))


def stream_ptr():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _on_device(t):
    return t.is_cuda


def ptr(t, dtype=None, allow_none=True):
    """Raw device pointer of a contiguous tensor (None -> NULL)."""
    if t is None:
        if not allow_none:
            raise ValueError('tensor required')
        return c_void_p(None)
    if not _on_device(t):
        raise CpgLibraryError('expected a CUDA tensor, got device %s (no CPU path)' % t.device)
    if dtype is not None and t.dtype != dtype:
        raise TypeError('expected dtype %s, got %s' % (dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError('tensor must be contiguous')
    return c_void_p(t.data_ptr())


def param_layout(n_vocab):
    offs = (c_int64 * N_PARAM_TENSORS)()
    sizes = (c_int64 * N_PARAM_TENSORS)()
    check(lib().cpg_vae_param_layout(n_vocab, offs, sizes), 'cpg_vae_param_layout')
    return list(offs), list(sizes), int(lib().cpg_vae_param_count(n_vocab))


def launch_count():
    return int(lib().cpg_launch_count(None))


def profile_enable(on=True):
    lib().cpg_profile_enable(1 if on else 0)


def profile_read(cap=128):
    """-> list of (kernel label, total ms, launches) since the last read (synchronises)."""
    stride = 64
    names = ctypes.create_string_buffer(cap * stride)
    ms = (c_float * cap)()
    cnt = (c_int * cap)()
    n = lib().cpg_profile_read(ctypes.cast(names, c_void_p), stride, ctypes.cast(ms, c_void_p),
                               ctypes.cast(cnt, c_void_p), cap)
    out = []
    for i in range(n):
        label = names.raw[i * stride:(i + 1) * stride].split(b'\0', 1)[0].decode()
        out.append((label, float(ms[i]), int(cnt[i])))
    return out


def profile_timeline(cap=4096):
    """-> list of (kernel label, start ms, duration ms) of the launches covered by the last profile_read()."""
    stride = 64
    names = ctypes.create_string_buffer(cap * stride)
    t0 = (c_float * cap)()
    dt = (c_float * cap)()
    n = lib().cpg_profile_timeline(ctypes.cast(names, c_void_p), stride, ctypes.cast(t0, c_void_p),
                                   ctypes.cast(dt, c_void_p), cap)
    return [(names.raw[i * stride:(i + 1) * stride].split(b'\0', 1)[0].decode(), float(t0[i]), float(dt[i])) for i in range(n)]


def set_option(name, value):
    check(lib().cpg_set_option(name.encode(), int(value)), 'cpg_set_option')

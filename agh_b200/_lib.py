"""ctypes binding of libagh_b200.so (the C ABI declared in include/agh_b200.h).

The library is built in-tree (agh_b200/libagh_b200.so) by `build_library()` /
`__graft_entry__.build()` with nvcc for sm_100a.  There is NO fallback: if the
shared object is missing, or a call is made with CPU tensors, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libagh_b200.so')
CSRC = os.path.join(_HERE, 'csrc')
SOURCES = ['decode.cu', 'decode_shared.cu', 'encoder.cu', 'gemm_tc.cu', 'ffn_tc.cu', 'mha_mma.cu', 'mha_tc.cu', 'env.cu', 'train.cu']
NVCC_FLAGS = ['-shared', '-Xcompiler', '-fPIC', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3',
              '-std=c++17']

AGH_GREEDY, AGH_SAMPLING, AGH_REPLAY = 0, 1, 2
AGH_TOUR_INVALID, AGH_CAPACITY_EXCEEDED, AGH_TW_VIOLATED, AGH_DECODE_STUCK = 1, 2, 4, 8
KEY_STRIDE = 128
D = 128


class AghError(RuntimeError):
    pass


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile agh_b200/csrc/*.cu into agh_b200/libagh_b200.so (sm_100a, -lineinfo)."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, 'common.cuh'), os.path.join(CSRC, 'gemm.cuh'), os.path.join(CSRC, 'tc.cuh'), os.path.join(_HERE, '..', 'include', 'agh_b200.h')]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    extra = os.environ.get('AGH_NVCC_EXTRA', '').split()          # e.g. -DAGH_TC_TRANSPOSE_MASK=0x7 for A/B builds
    flags = [f for f in NVCC_FLAGS if f != '-shared'] + extra + (['-Xptxas', '-v'] if verbose else [])
    objdir = os.path.join(CSRC, '_obj')
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):                                          # one translation unit per process, in parallel
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + '.o')
        return obj, subprocess.run([nvcc] + flags + ['-c', src, '-o', obj], capture_output=True, text=True)

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(len(srcs)) as pool:
        results = list(pool.map(compile_one, srcs))
    for obj, res in results:
        if res.returncode != 0:
            raise AghError('nvcc failed:\n' + res.stdout + res.stderr)
        if verbose:
            print(res.stderr)
    res = subprocess.run([nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB_PATH] + [o for o, _ in results],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise AghError('nvcc link failed:\n' + res.stdout + res.stderr)
    return LIB_PATH


class DecodeArgs(C.Structure):
    _fields_ = [
        ('keys', C.c_void_p), ('psc', C.c_void_p), ('qbase', C.c_void_p), ('coord', C.c_void_p),
        ('demand', C.c_void_p), ('tw_left', C.c_void_p), ('tw_right', C.c_void_p), ('duration', C.c_void_p),
        ('distance', C.c_void_p), ('travel', C.c_void_p), ('wblob', C.c_void_p),
        ('key_mod', C.c_int), ('smem_mats', C.c_int), ('mode', C.c_int), ('temp', C.c_float), ('seed', C.c_uint64), ('id_offset', C.c_int64),
        ('actions', C.c_void_p), ('replay_steps', C.c_int), ('inject_logp', C.c_void_p), ('inject_z', C.c_void_p),
        ('t_stride', C.c_int),
        ('pi', C.c_void_p), ('steps', C.c_void_p), ('status', C.c_void_p), ('ll', C.c_void_p), ('cost', C.c_void_p), ('serve', C.c_void_p),
        ('logp_sel', C.c_void_p), ('logp_full', C.c_void_p), ('mask_dump', C.c_void_p), ('used_dump', C.c_void_p),
        ('cft_dump', C.c_void_p), ('q_dump', C.c_void_p), ('attn_dump', C.c_void_p), ('o_dump', C.c_void_p),
        ('lse_dump', C.c_void_p), ('B', C.c_int), ('N', C.c_int),
    ]


class GemmDesc(C.Structure):
    _fields_ = [
        ('A', C.c_void_p), ('lda', C.c_longlong), ('transA', C.c_int),
        ('B', C.c_void_p), ('ldb', C.c_longlong), ('transB', C.c_int),
        ('C', C.c_void_p), ('ldc', C.c_longlong),
        ('M', C.c_int), ('N', C.c_int), ('K', C.c_int), ('batch1', C.c_int), ('batch2', C.c_int),
        ('sA1', C.c_longlong), ('sA2', C.c_longlong), ('sB1', C.c_longlong), ('sB2', C.c_longlong),
        ('sC1', C.c_longlong), ('sC2', C.c_longlong),
        ('alpha', C.c_float), ('accumulate', C.c_int), ('bias', C.c_void_p), ('relu', C.c_int),
        ('mask', C.c_void_p), ('ldm', C.c_longlong), ('splits', C.c_int),
    ]


class StateArgs(C.Structure):
    _fields_ = [
        ('coord', C.c_void_p), ('demand', C.c_void_p), ('tw_left', C.c_void_p), ('tw_right', C.c_void_p),
        ('duration', C.c_void_p), ('distance', C.c_void_p), ('prev_a', C.c_void_p), ('used_capacity', C.c_void_p),
        ('visited', C.c_void_p), ('lengths', C.c_void_p), ('cur_free_time', C.c_void_p), ('serve_time', C.c_void_p),
        ('B', C.c_int), ('N', C.c_int),
    ]


# every symbol include/agh_b200.h declares: name -> (restype, argtypes)
_P, _I = C.c_void_p, C.c_int
SYMBOLS = {
    'agh_last_error': (C.c_char_p, []),
    'agh_version': (_I, []),
    'agh_launch_count': (C.c_ulonglong, []),
    'agh_wblob_floats': (C.c_size_t, []),
    'agh_sizeof_decode_args': (C.c_size_t, []),
    'agh_sizeof_state': (C.c_size_t, []),
    'agh_fleet_task': (_I, [_P, _P, _P, _P, _P, _I, _I, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    'agh_chain_tw_left': (_I, [_P, _I, _P, _I, _I, _P]),
    'agh_encoder_workspace_bytes': (C.c_size_t, [_I, _I]),
    'agh_set_gemm_backend': (_I, [_I]),
    'agh_set_ffn_fused': (_I, [_I]),
    'agh_set_mha_tc': (_I, [_I]),
    'agh_encoder_mha': (_I, [_P, _I, _I, _P, _P]),
    'agh_encoder_ffn_workspace_bytes': (C.c_size_t, [_I]),
    'agh_encoder_ffn': (_I, [_P, _I, _P, _I, _P, _P, _P]),
    'agh_encode': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P]),
    'agh_init_embed': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P, _P]),
    'agh_encode_layers': (_I, [_P, _P, _I, _I, _P, _P, _P]),
    'agh_precompute': (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _P, _P]),
    'agh_decode': (_I, [C.POINTER(DecodeArgs), _P]),
    'agh_get_costs': (_I, [_P, _I, _I, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _P]),
    'agh_state_get_mask': (_I, [C.POINTER(StateArgs), _P, _P]),
    'agh_state_update': (_I, [C.POINTER(StateArgs), _P, _P]),
    'agh_best_of': (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    'agh_generate': (_I, [C.c_uint64, C.c_longlong, _I, _I, _I, C.c_float, _P, _P, _P, _P, _P, _P, _P]),
    'agh_sizeof_gemm_desc': (C.c_size_t, []),
    'agh_gemm': (_I, [C.POINTER(GemmDesc), _P]),
    'agh_gemm_f64': (_I, [_P, _I, _I, _P, _I, _I, _P, _I, _I, _I, _I, C.c_double, _P]),
    'agh_colsum': (_I, [_P, C.c_longlong, _I, _I, _P, _P]),
    'agh_bn_train_fwd': (_I, [_P, _P, _P, _I, C.c_float, C.c_float, _P, _P, _P, _P, _P, _P, _P]),
    'agh_bn_train_bwd': (_I, [_P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P]),
    'agh_mha_fwd': (_I, [_P, _I, _I, _P, _P]),
    'agh_mha_bwd': (_I, [_P, _P, _P, _I, _I, _P, _P]),
    'agh_tf_dlogits': (_I, [_P, _P, _P, _P, _P, C.c_float, _I, _I, _I, _P, _P]),
    'agh_tf_dscore': (_I, [_P, _P, C.c_longlong, _I, _P]),
    'agh_tf_dquery': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P]),
}

_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Load the shared library (once).  Raises AghError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AghError('%s not found: run `python -c "import __graft_entry__ as g; g.build()"` '
                           '(nvcc, sm_100a). There is no CPU or PyTorch fallback.' % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)          # AttributeError if the symbol is missing
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise AghError('libagh_b200 call failed (%d): %s' % (rc, lib().agh_last_error().decode()))


def ptr(t: Optional[torch.Tensor], dtype=None) -> Optional[int]:
    """Device pointer of a contiguous CUDA tensor (None passes NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise AghError('agh_b200 kernels need CUDA tensors (got a %s tensor); there is no CPU path' % t.device)
    if not t.is_contiguous():
        raise AghError('non-contiguous tensor passed to the C ABI')
    if dtype is not None and t.dtype != dtype:
        raise AghError('expected dtype %s, got %s' % (dtype, t.dtype))
    return t.data_ptr()


def stream_ptr(device=None) -> int:
    """Current stream of `device`.  The C ABI launches on the CURRENT device, so the tensors' device must be it."""
    if device is not None:
        idx = torch.device(device).index
        if idx is not None and idx != torch.cuda.current_device():
            raise AghError('tensors live on cuda:%d but the current device is cuda:%d: call torch.cuda.set_device(%d) '
                           '(or wrap the call in torch.cuda.device) first' % (idx, torch.cuda.current_device(), idx))
    return torch.cuda.current_stream(device).cuda_stream

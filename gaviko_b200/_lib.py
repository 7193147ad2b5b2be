"""ctypes binding of libgvk_sm100a.so (include/gvk.h).  There is no CPU fallback: if the library is missing, or a
tensor is not on a CUDA device, the call raises."""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libgvk_sm100a.so')

GVK_F32, GVK_BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_GELU_BWD = 0, 1, 2

_lib = None


class GvkError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GvkError(f'{LIB_PATH} is not built (run `python -m gaviko_b200.build`); gaviko_b200 has no CPU fallback')
        _lib = C.CDLL(LIB_PATH)
        _lib.gvk_last_error.restype = C.c_char_p
        _lib.gvk_launch_count.restype = C.c_uint64
        _lib.gvk_version.restype = C.c_int
    return _lib


def check(status: int, what: str):
    if status != 0:
        raise GvkError(f'{what} failed ({status}): {lib().gvk_last_error().decode()}')


def launch_count() -> int:
    return int(lib().gvk_launch_count())


def dtype_tag(t: torch.dtype) -> int:
    if t == torch.float32:
        return GVK_F32
    if t == torch.bfloat16:
        return GVK_BF16
    raise GvkError(f'unsupported dtype {t}')


def ptr(t):
    """Device pointer of an optional tensor (None -> NULL); refuses CPU tensors (no CPU fallback)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise GvkError('gaviko_b200 kernels need CUDA tensors (there is no CPU fallback)')
    return C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class GemmParams(C.Structure):
    _fields_ = [
        ('a', C.c_void_p), ('b', C.c_void_p), ('ab_dtype', C.c_int), ('M', C.c_int), ('N', C.c_int), ('K', C.c_int),
        ('lda', C.c_int), ('ldb', C.c_int),
        ('bias', C.c_void_p), ('ssf_scale', C.c_void_p), ('ssf_shift', C.c_void_p), ('act', C.c_int),
        ('aux', C.c_void_p), ('aux_dtype', C.c_int), ('ld_aux', C.c_int),
        ('pos', C.c_void_p), ('rows_per_batch', C.c_int), ('out_batch_rows', C.c_int), ('out_row_offset', C.c_int),
        ('res1', C.c_void_p), ('ld_res1', C.c_int), ('res2', C.c_void_p), ('ld_res2', C.c_int),
        ('out', C.c_void_p), ('out_dtype', C.c_int), ('ld_out', C.c_int), ('out2', C.c_void_p), ('ld_out2', C.c_int),
    ]

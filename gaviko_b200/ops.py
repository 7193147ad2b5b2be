"""Thin torch-tensor wrappers over the C ABI (include/gvk.h).  Plumbing only: pointer/shape marshalling and allocation
of outputs; all arithmetic happens in libgvk_sm100a.so on torch's current CUDA stream."""
import ctypes as C

import torch

from . import _lib as L
from ._lib import ACT_GELU, ACT_GELU_BWD, ACT_NONE  # noqa: F401


def _ld(t):
    assert t.dim() == 2 and t.stride(1) == 1, 'expected a row-major 2-D tensor (last stride 1)'
    return t.stride(0)


def gemm(a, b, *, out=None, out_dtype=None, bias=None, ssf_scale=None, ssf_shift=None, act=ACT_NONE, aux=None,
         pos=None, rows_per_batch=0, out_batch_rows=0, out_row_offset=0, res1=None, res2=None, out2=None, out_rows=None):
    """out[row(m), n] = epilogue(sum_k a[m,k] * b[n,k]) — see gvk_gemm in include/gvk.h."""
    M, K = a.shape
    N, Kb = b.shape
    assert K == Kb and a.dtype == b.dtype
    if out is None:
        out = torch.empty((out_rows if out_rows is not None else M, N), device=a.device, dtype=out_dtype or torch.float32)
    p = L.GemmParams()
    p.a, p.b = a.data_ptr(), b.data_ptr()
    if not (a.is_cuda and b.is_cuda and out.is_cuda):
        raise L.GvkError('gaviko_b200 kernels need CUDA tensors (there is no CPU fallback)')
    p.ab_dtype = L.dtype_tag(a.dtype)
    p.M, p.N, p.K = M, N, K
    p.lda, p.ldb = _ld(a), _ld(b)
    for name, t in (('bias', bias), ('ssf_scale', ssf_scale), ('ssf_shift', ssf_shift), ('pos', pos)):
        if t is not None:
            assert t.dtype == torch.float32 and t.is_contiguous() and t.is_cuda
            setattr(p, name, t.data_ptr())
    p.act = act
    if aux is not None:
        p.aux, p.aux_dtype, p.ld_aux = aux.data_ptr(), L.dtype_tag(aux.dtype), _ld(aux)
    p.rows_per_batch, p.out_batch_rows, p.out_row_offset = rows_per_batch, out_batch_rows, out_row_offset
    if res1 is not None:
        assert res1.dtype == torch.float32
        p.res1, p.ld_res1 = res1.data_ptr(), _ld(res1)
    if res2 is not None:
        assert res2.dtype == torch.float32
        p.res2, p.ld_res2 = res2.data_ptr(), _ld(res2)
    p.out, p.out_dtype, p.ld_out = out.data_ptr(), L.dtype_tag(out.dtype), _ld(out)
    if out2 is not None:
        assert out2.dtype == torch.float32
        p.out2, p.ld_out2 = out2.data_ptr(), _ld(out2)
    L.check(L.lib().gvk_gemm(C.byref(p), L.stream()), 'gvk_gemm')
    return out

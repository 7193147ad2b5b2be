"""ctypes binding of libgvk_sm100a.so.  Struct layouts are parsed from include/gvk.h (the single source of truth for the
C ABI) so the Python mirror cannot drift.  There is no CPU fallback: if the library is missing, or a tensor is not on a
CUDA device, the call raises GvkError."""
import ctypes as C
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libgvk_sm100a.so')
HEADER_PATH = os.path.join(os.path.dirname(_HERE), 'include', 'gvk.h')

GVK_F32, GVK_BF16 = 0, 1
ACT_NONE, ACT_GELU, ACT_GELU_BWD, ACT_GELU_SAVE_GRAD, ACT_MUL_AUX = 0, 1, 2, 3, 4
ROWACT_NONE, ROWACT_QUICKGELU, ROWACT_RELU = 0, 1, 2
LOSS_FOCAL, LOSS_CE = 0, 1


class GvkError(RuntimeError):
    pass


_SCALARS = {'int': C.c_int, 'float': C.c_float, 'uint64_t': C.c_uint64, 'size_t': C.c_size_t, 'long long': C.c_longlong,
            'int64_t': C.c_int64, 'uint32_t': C.c_uint32, 'double': C.c_double}


def _parse_header(path):
    """typedef struct { ... } name;  ->  ctypes.Structure subclasses (nested structs by value supported)."""
    src = open(path).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    src = re.sub(r'//[^\n]*', '', src)
    structs = {}
    for body, name in re.findall(r'typedef\s+struct\s*\{(.*?)\}\s*(\w+)\s*;', src, flags=re.S):
        fields = []
        for decl in body.split(';'):
            decl = ' '.join(decl.split())
            if not decl:
                continue
            if '*' in decl:
                base, names = decl.rsplit('*', 1)
                assert ',' not in names, f'one pointer per declaration please: {decl}'
                fields.append((names.strip(), C.c_void_p))
                continue
            m = re.match(r'(?:const\s+)?((?:unsigned\s+)?(?:long long|\w+))\s+(.*)', decl)
            ctype_name, names = m.group(1), m.group(2)
            ctype = _SCALARS.get(ctype_name) or structs.get(ctype_name)
            assert ctype is not None, f'unknown type in gvk.h: {decl}'
            for n in names.split(','):
                fields.append((n.strip(), ctype))
        structs[name] = type(name, (C.Structure,), {'_fields_': fields})
    funcs = re.findall(r'\bint\s+(gvk_\w+)\s*\(', src)
    return structs, sorted(set(funcs))


STRUCTS, FUNCTIONS = _parse_header(HEADER_PATH)
_lib = None


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


def ptr(t, dtype=None):
    """Device address of an optional tensor (None -> NULL).  Refuses CPU tensors: there is no CPU fallback."""
    if t is None:
        return None
    if not t.is_cuda:
        raise GvkError('gaviko_b200 kernels need CUDA tensors (there is no CPU fallback)')
    if dtype is not None and t.dtype != dtype:
        raise GvkError(f'expected {dtype}, got {t.dtype}')
    return t.data_ptr()


def require_cuda(t, what='gaviko_b200 models'):
    """There is no CPU fallback: a CPU tensor is an error, not a slow path."""
    if not t.is_cuda:
        raise GvkError(f'{what} run on CUDA only (no CPU fallback): move the model and the input to a B200')


def device_guard(t):
    """Make the tensor's device current for the enclosed launches: kernels go to the CURRENT device's stream, and the reference's scripts never
    call torch.cuda.set_device (train.py:99)."""
    return torch.cuda.device(t.device)


class untraced:
    """Pause a torch.jit trace around the engine call.  The reference's validation loop jit-TRACES the model (train.py:246-252,405-407:
    torchprofile.profile_macs -> torch.jit._get_trace_graph); under a trace every `tensor.shape[i]` is a traced 0-d tensor, which neither ctypes
    nor the host-side bookkeeping can use, and the kernels are opaque to the tracer anyway.  With the trace paused the forward runs normally and the
    tracer records the logits as a constant of the graph (profile_macs then counts no MACs for the model, as it does for any op without a handler)."""

    def __enter__(self):
        self.state = torch._C._get_tracing_state()
        if self.state is not None:
            torch._C._set_tracing_state(None)

    def __exit__(self, *exc):
        if self.state is not None:
            torch._C._set_tracing_state(self.state)
        return False

    def reattach(self, logits, img):
        """After the paused region: give the traced graph a data dependence of the output on the input (the tracer refuses a graph without one)."""
        if self.state is None:
            return logits
        return logits + (img.reshape(-1)[:1] * 0).to(logits.dtype)


def fptr(t):
    """fp32, contiguous device tensor (or None)."""
    if t is not None and not t.is_contiguous():
        raise GvkError('expected a contiguous tensor')
    return ptr(t, torch.float32)


def stream():
    """torch's current stream of the current device as a raw handle.  (torch.cuda.current_stream() builds a Stream object through several
    Python layers: 14 us per call under the profiler, 20 % of the host time of a training step; the two C calls below are ~0.3 us.)"""
    try:
        return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))
    except AttributeError:      # private torch entry points moved: fall back to the public API
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# Optional per-entry-point timing (tools/profile_step.py): PROFILE = {} enables CUDA-event brackets around every C-ABI call.
PROFILE = None


PROFILE_BY_SITE = False      # key the profile by engine.py call site (name@line) instead of by entry point


def _site(name):
    import sys
    f = sys._getframe(2)
    while f is not None:
        if f.f_code.co_filename.endswith('engine.py'):
            return f'{name}@{f.f_lineno}'
        f = f.f_back
    return name


def call(name, *args):
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(getattr(lib(), name)(*args), name)
        e1.record()
        PROFILE.setdefault(_site(name) if PROFILE_BY_SITE else name, []).append((e0, e1))
        return
    check(getattr(lib(), name)(*args), name)

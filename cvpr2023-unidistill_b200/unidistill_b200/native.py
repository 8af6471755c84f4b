"""ctypes binding of libunidistill_b200.so (the C ABI declared in include/unidistill_b200.h).

There is no CPU fallback anywhere behind this module: if the shared library is missing or a kernel
returns an error, a RuntimeError is raised.
"""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG_ROOT = os.path.dirname(_HERE)
# UD_B200_LIB: an alternative build of the same library (kernel experiments: tools/build_dg_variants.sh)
LIB_PATH = os.environ.get("UD_B200_LIB") or os.path.join(_PKG_ROOT, "libunidistill_b200.so")

_P = ctypes.c_void_p
_I = ctypes.c_int
_F = ctypes.c_float
_Z = ctypes.c_size_t
_ZP = ctypes.POINTER(ctypes.c_size_t)

HEADER_PATH = os.path.join(os.path.dirname(_PKG_ROOT), "include", "unidistill_b200.h")


def parse_header(path=HEADER_PATH):
    """Returns {function name: (restype char, argtype chars)} for every prototype in the C-ABI header.

    'p' pointer, 'i' int, 'f' float, 'z' size_t, 'Z' size_t*, 's' const char* (return only).
    The binding is derived from the header so the two cannot drift apart.
    """
    import re

    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    protos = {}
    for m in re.finditer(r"(const\s+char\s*\*|int)\s+(ud_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, params = m.group(1), m.group(2), m.group(3).strip()
        spec = ""
        if params and params != "void":
            for p in params.split(","):
                p = p.strip()
                if "*" in p:
                    spec += "Z" if re.search(r"\bsize_t\s*\*", p) else "p"
                elif re.search(r"\bsize_t\b", p):
                    spec += "z"
                elif re.search(r"\bfloat\b", p):
                    spec += "f"
                elif re.search(r"\bint\b", p):
                    spec += "i"
                else:
                    raise RuntimeError("unidistill_b200: cannot bind parameter %r of %s" % (p, name))
        protos[name] = ("s" if "char" in ret else "i", spec)
    return protos


_CT = {"p": _P, "i": _I, "f": _F, "z": _Z, "Z": _ZP}

_lib = None
_lock = threading.Lock()


def load():
    """Loads the shared library once; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "unidistill_b200: %s is missing -- run `python cvpr2023-unidistill_b200/build.py` "
                "(or __graft_entry__.build()); there is no fallback path" % LIB_PATH
            )
        cdll = ctypes.CDLL(LIB_PATH)
        lib = _Lib(cdll)
        for name, (ret, spec) in parse_header().items():
            fn = getattr(cdll, name)  # AttributeError here == header/library mismatch
            fn.argtypes = [_CT[c] for c in spec]
            fn.restype = ctypes.c_char_p if ret == "s" else _I
            launches = spec.endswith("p") and not name.endswith(("_bytes", "_sizes", "_supported")) \
                and name not in ("ud_debug_umma_prof",)
            setattr(lib, name, _timed(name, fn) if launches else fn)
        _lib = lib
    return _lib


class _Lib:
    """Attribute bag of the bound entry points (kernel-launching ones pass through the KernelTimer hook)."""

    def __init__(self, cdll):
        self._cdll = cdll


def _timed(name, fn):
    def call(*args):
        if not TIMER.enabled:
            return fn(*args)
        return TIMER.call(name, fn, args)

    call.__name__ = name
    return call


def check(rc, what):
    if rc != 0:
        msg = load().ud_error_string(int(rc)).decode()
        raise RuntimeError("unidistill_b200.%s failed with code %d: %s" % (what, rc, msg))


def ptr(t):
    """Device (or host) pointer of a tensor, None -> NULL."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def host_ints(values):
    arr = (ctypes.c_int * len(values))(*[int(v) for v in values])
    return arr


def host_floats(values):
    arr = (ctypes.c_float * len(values))(*[float(v) for v in values])
    return arr


def host_ptr(arr):
    return ctypes.cast(arr, ctypes.c_void_p)


def query_bytes(fn_name, *args):
    lib = load()
    out = ctypes.c_size_t(0)
    rc = getattr(lib, fn_name)(*args, ctypes.byref(out))
    check(rc, fn_name)
    return int(out.value)


class Workspace:
    """Grow-only scratch buffer per (device, slot, current stream) handed to the C ABI."""

    def __init__(self):
        self._buf = {}

    def get(self, nbytes, device, slot=0):
        dev = torch.device(device)
        stream = torch.cuda.current_stream(dev).cuda_stream if dev.type == "cuda" else 0
        key = (str(device), slot, stream)          # one buffer per stream: work on a side stream never shares scratch
        buf = self._buf.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
            self._buf[key] = buf
        return buf


WORKSPACE = Workspace()


def launch_count(reset=False):
    """Kernels launched by the library since the last reset (every `<<<>>>` of csrc/ counts one)."""
    return int(load().ud_launch_count(1 if reset else 0))


class KernelTimer:
    """CUDA-event timing of every kernel-launching C-ABI call on the stream it is launched on (bench.py's roofline
    numbers are taken live, inside the timed region, through this hook).  Disabled by default."""

    def __init__(self):
        self.enabled = False
        self.records = []
        self._pool = []
        self.meta = None            # optional dict a call site attaches to the NEXT timed call (tensors for lazy counts)

    def prepare(self, n_pairs):
        """Creates (and materialises, by recording once) n_pairs event pairs ahead of the timed region:
        cudaEventCreate inside the region contends with NVML clock sampling for a driver lock."""
        self._pool = []
        for _ in range(int(n_pairs)):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            e1.record()
            self._pool.append((e0, e1))
        torch.cuda.synchronize()

    def call(self, key, fn, args):
        meta, self.meta = self.meta, None
        if self._pool:
            e0, e1 = self._pool.pop()
        else:
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        ints = tuple(a if isinstance(a, (int, float)) else None for a in args)
        self.records.append((key, meta, ints, e0, e1))
        return rc

    def drain(self):
        """-> [(entry point, meta, scalar args (pointers -> None), milliseconds)], after a device synchronize."""
        torch.cuda.synchronize()
        out = [(k, m, a, e0.elapsed_time(e1)) for k, m, a, e0, e1 in self.records]
        self.records = []
        return out


TIMER = KernelTimer()


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "unidistill_b200: this operator only runs on a CUDA device (sm_100a); got a %s tensor. "
                "There is no CPU fallback." % t.device
            )

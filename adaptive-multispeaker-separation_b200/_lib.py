"""ctypes binding of libamss_b200.so, generated from include/amss.h.

There is no fallback: importing this module raises if the shared library is missing, and
every call raises AmssError on a non-zero status.  Device buffers are passed as raw pointers
(torch is only the allocator / stream owner on the host side).
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HEADER = os.path.join(ROOT, "include", "amss.h")
LIB_PATH = os.path.join(HERE, "libamss_b200.so")

AMSS_PREC_FP32, AMSS_PREC_BF16 = 0, 1
AMSS_POOL_MAX, AMSS_POOL_AVG, AMSS_POOL_STRIDE = 0, 1, 2


class AmssError(RuntimeError):
    pass


_CTYPES = {
    "int": ctypes.c_int, "float": ctypes.c_float, "size_t": ctypes.c_size_t,
    "int64_t": ctypes.c_int64, "uint64_t": ctypes.c_uint64, "int32_t": ctypes.c_int32,
    "uint8_t": ctypes.c_uint8,
}


def parse_header(path=HEADER):
    """Returns {name: (restype_str, [(type_str, arg_name), ...])} for every declaration."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
    decls = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(amss_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                mm = re.match(r"(.*?)(\w+)$", a)
                params.append((mm.group(1).strip(), mm.group(2)))
        decls[name] = (ret, params)
    return decls


def _ctype(t):
    t = t.replace("const ", "").strip()
    if t.endswith("*"):
        return ctypes.c_char_p if t == "char*" else ctypes.c_void_p
    return _CTYPES[t]


DECLS = parse_header()

if not os.path.exists(LIB_PATH):
    raise AmssError(
        f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
        "There is no CPU or PyTorch fallback for this package.")

_lib = ctypes.CDLL(LIB_PATH)
_fns = {}
for _name, (_ret, _params) in DECLS.items():
    _f = getattr(_lib, _name)   # AttributeError if the library does not export a declared symbol
    _f.argtypes = [_ctype(t) for t, _ in _params]
    _f.restype = _ctype(_ret)
    _fns[_name] = _f


def last_error():
    return _fns["amss_last_error"]().decode()


def raw(name):
    return _fns[name]


# Optional per-entry-point device timing (bench.py / tools): time_calls([...]) makes call() bracket the named entry
# points with CUDA events on the launching stream; timed_calls() returns {name: [(args, ms), ...]} after a synchronize.
_timed = None


def time_calls(names):
    """Start (names = iterable of entry points, or "all") or stop (names = None) bracketing ABI calls with CUDA events."""
    global _timed
    _timed = None if names is None else {"__all__": names == "all", "names": set(() if names == "all" else names), "ev": []}


def timed_calls():
    import torch
    torch.cuda.synchronize()
    out = {}
    for name, args, e0, e1 in (_timed["ev"] if _timed else ()):
        out.setdefault(name, []).append((args, e0.elapsed_time(e1)))
    return out


def call(name, *args):
    """Call an int-status entry point; raise AmssError with the library's message on failure."""
    ev = None
    if _timed is not None and (_timed["__all__"] or name in _timed["names"]):
        import torch
        if not torch.cuda.is_current_stream_capturing():
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
    rc = _fns[name](*args)
    if ev is not None:
        ev[1].record()
        _timed["ev"].append((name, args, ev[0], ev[1]))
    if rc != 0:
        raise AmssError(f"{name} failed ({rc}): {last_error()}")


def query(name, *args):
    """Call a size/count query (returns the value)."""
    return _fns[name](*args)


def launch_count():
    return int(_fns["amss_launch_count"]())

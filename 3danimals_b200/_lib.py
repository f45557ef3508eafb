"""ctypes binding of libb2a.so, the C-ABI CUDA library behind this package (include/b2a.h).

Prototypes are derived from the header itself, so the binding cannot drift from the declared ABI.  There is NO
fallback: if the shared library is missing or fails to load, importing any compute op raises - the product never
routes through PyTorch eager ops or the CPU oracle for the hot path.
"""
import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), "include", "b2a.h")
LIB_PATH = os.path.join(HERE, "csrc", "_build", "libb2a.so")

_SCALARS = {"int": C.c_int, "int64_t": C.c_int64, "size_t": C.c_size_t, "float": C.c_float, "b2a_stream_t": C.c_void_p}


class B2AError(RuntimeError):
    pass


def parse_header(path=HEADER):
    """-> {name: (restype, [(ctype, param_name), ...])} for every function declared in include/b2a.h."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    protos = {}
    for m in re.finditer(r"(const\s+char\s*\*|int)\s+(b2a_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                pname = re.search(r"(\w+)$", a).group(1)
                typ = a[: -len(pname)].strip()
                if typ.endswith("*"):
                    base = typ[:-1].replace("const", "").strip()
                    ctype = C.POINTER(C.c_size_t) if base == "size_t" else (C.POINTER(C.c_int) if base == "int" else C.c_void_p)
                else:
                    ctype = _SCALARS[typ.replace("const", "").strip()]
                params.append((ctype, pname))
        protos[name] = (C.c_char_p if "char" in ret else C.c_int, params)
    return protos


_lib = None


def lib():
    """The loaded library (built on first use when the toolchain is present; otherwise must already exist)."""
    global _lib
    if _lib is not None:
        return _lib
    path = LIB_PATH
    if not os.path.isfile(path):
        try:
            from . import build as _build
            path = _build.build()
        except Exception as e:  # no nvcc on this host and no prebuilt library: fail loudly
            raise B2AError("libb2a.so not found at %s and could not be built (%s); run `python 3danimals_b200/build.py`"
                           % (LIB_PATH, e)) from e
    handle = C.CDLL(path)
    for name, (restype, params) in parse_header().items():
        fn = getattr(handle, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = restype
        fn.argtypes = [p[0] for p in params]
    if handle.b2a_version() != 100:
        raise B2AError("libb2a.so version mismatch: %d" % handle.b2a_version())
    _lib = handle
    return handle


_fast = False


def fast():
    """The generated METH_FASTCALL binding of the compute entry points (build.py build_fastcall), or None.  Same library,
    same C-ABI: only the Python->C argument conversion differs (ctypes costs ~4 us per 30-argument call)."""
    global _fast
    if _fast is False:
        _fast = None
        lib()       # the extension links against libb2a.so ($ORIGIN rpath): make sure it exists / is built first
        try:
            import importlib.machinery
            import importlib.util
            from . import build as _build
            path = _build.fastcall_path()
            if not os.path.isfile(path):
                path = _build.build_fastcall()
            if path and os.path.isfile(path) and os.environ.get("B2A_CTYPES_ONLY", "0") != "1":
                loader = importlib.machinery.ExtensionFileLoader(_build.FASTCALL_NAME, path)
                spec = importlib.util.spec_from_loader(_build.FASTCALL_NAME, loader)
                mod = importlib.util.module_from_spec(spec)
                loader.exec_module(mod)
                _fast = mod
        except Exception:
            _fast = None
    return _fast


_cpp = False


def cpp_nodes():
    """The C++ autograd layer (csrc/autograd_nodes.cpp, build.py build_autograd), or None: same C-ABI underneath, the nodes' host
    work (allocation, bookkeeping, autograd engine hand-off) in C++ instead of Python.  B2A_CPP_NODES=0 disables it."""
    global _cpp
    if _cpp is False:
        _cpp = None
        if os.environ.get("B2A_CPP_NODES", "1") != "0":
            lib()
            try:
                import importlib.machinery
                import importlib.util
                import torch  # noqa: F401  (libtorch must be loaded first)
                from . import build as _build
                path = _build.autograd_path()
                if not os.path.isfile(path):
                    path = _build.build_autograd()
                if path and os.path.isfile(path):
                    loader = importlib.machinery.ExtensionFileLoader(_build.AUTOGRAD_NAME, path)
                    spec = importlib.util.spec_from_loader(_build.AUTOGRAD_NAME, loader)
                    mod = importlib.util.module_from_spec(spec)
                    loader.exec_module(mod)
                    _cpp = mod
            except Exception:
                _cpp = None
    return _cpp


def check(rc):
    if rc != 0:
        raise B2AError(lib().b2a_last_error_string().decode("utf-8", "replace") or "libb2a error %d" % rc)

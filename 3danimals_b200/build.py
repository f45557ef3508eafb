"""Build libb2a.so (the C-ABI CUDA library of this package) in-tree for sm_100a.

    python 3danimals_b200/build.py [--force] [--verbose]

Output: 3danimals_b200/csrc/_build/libb2a.so (git-ignored; travels to the GPU box with the gpurun snapshot).
Every file: -gencode arch=compute_100a,code=sm_100a  -lineinfo (ncu source pages).
Two arithmetic regimes, chosen per file:
  EXACT  (-fmad=false, IEEE div/sqrt): files whose results are compared BIT-EXACTLY with the oracle - triangle ids /
         barycentrics (raster.cu), extracted vertices and faces (marching_tets.cu), interpolated attributes
         (interpolate.cu), antialiased images (antialias.cu), quantile selection (estimate_bones.cu).  Every fp32 op is
         individually rounded; oracle/raster_ref.c is built with -ffp-contract=off to match.
  FAST   (-use_fast_math: FMA contraction, approximate div / sqrt / exp): files whose contract is the 1e-4 relative
         tolerance - fused g-buffer (gbuffer.cu), skinning (lbs.cu), vertex normals (normals.cu), light shading (shade.cu).  The IEEE division
         subroutine alone was 25 % of the g-buffer backward's instructions (profiles/).
"""
import concurrent.futures
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_DIR = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(SRC_DIR, "_build")
LIB = os.path.join(OUT_DIR, "libb2a.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

COMMON_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden"]
EXACT_FLAGS = ["-fmad=false"]
FAST_FLAGS = ["-use_fast_math"]
FAST_FILES = {"gbuffer.cu", "lbs.cu", "normals.cu", "shade.cu"}


def sources():
    return sorted(glob.glob(os.path.join(SRC_DIR, "*.cu")))


def _deps():
    return sources() + glob.glob(os.path.join(SRC_DIR, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h")) + [os.path.abspath(__file__)]


def _stale():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def _compile(nvcc, src, verbose):
    obj = os.path.join(OUT_DIR, os.path.basename(src)[:-3] + ".o")
    flags = FAST_FLAGS if os.path.basename(src) in FAST_FILES else EXACT_FLAGS
    extra = os.environ.get("B2A_NVCC_DEFINES", "").split()      # profiling builds, e.g. -DB2A_MLP_TRACE (csrc/field_mlp.cu)
    cmd = [nvcc] + COMMON_FLAGS + flags + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return obj, res


def build(force=False, verbose=False):
    if not force and not _stale():
        build_fastcall()
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OUT_DIR, exist_ok=True)
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(lambda s: _compile(nvcc, s, verbose), sources()))
    objs = []
    for obj, res in results:
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed building %s" % obj)
        objs.append(obj)
    res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + ["-lcudart"], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libb2a.so")
    build_fastcall(force=True)
    return LIB


# ----------------------------------------------------------------------------------------------------------------
# Fast binding: a generated CPython extension with one METH_FASTCALL wrapper per compute entry point of include/b2a.h.
# ctypes spends ~4 us converting the ~30 arguments of a typical call (measured); the hot path makes 26 calls per step and
# is host-bound, so the binding layer itself was ~8 % of the step.  The wrappers convert ints / None / floats inline and
# call straight into libb2a.so (same C-ABI, nothing else changes); _lib.py falls back to ctypes when the module is absent.
# ----------------------------------------------------------------------------------------------------------------
FASTCALL_NAME = "_b2a_fastcall"


def fastcall_path():
    import sysconfig
    return os.path.join(OUT_DIR, FASTCALL_NAME + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))


def _fastcall_source(protos):
    conv = {"c_int": ("int", "(int)PyLong_AsLong(%s)"), "c_long": ("int64_t", "(int64_t)PyLong_AsLongLong(%s)"),
            "c_ulong": ("size_t", "(size_t)PyLong_AsUnsignedLongLong(%s)"), "c_float": ("float", "(float)PyFloat_AsDouble(%s)"),
            "c_void_p": ("void*", "(%s == Py_None ? (void*)0 : PyLong_AsVoidPtr(%s))")}
    out = ["#define PY_SSIZE_T_CLEAN", "#include <Python.h>", "#include <stdint.h>", "#include <stddef.h>", ""]
    table = []
    for name, (restype, params) in sorted(protos.items()):
        kinds = [p[0].__name__ for p in params]
        if restype.__name__ != "c_int" or any(k not in conv for k in kinds):
            continue            # size queries / version / error string keep the ctypes path (cold)
        out.append("extern int %s(%s);" % (name, ", ".join(conv[k][0] for k in kinds) or "void"))
        out.append("static PyObject* w_%s(PyObject* self, PyObject* const* a, Py_ssize_t n) {" % name)
        out.append("    if (n != %d) { PyErr_SetString(PyExc_TypeError, \"%s expects %d arguments\"); return NULL; }" % (len(kinds), name, len(kinds)))
        for i, k in enumerate(kinds):
            ctype, expr = conv[k]
            arg = "a[%d]" % i
            out.append("    %s v%d = %s;" % (ctype, i, expr % ((arg, arg) if k == "c_void_p" else (arg,))))
        out.append("    if (PyErr_Occurred()) return NULL;")
        out.append("    return PyLong_FromLong((long)%s(%s));" % (name, ", ".join("v%d" % i for i in range(len(kinds)))))
        out.append("}")
        table.append('    {"%s", (PyCFunction)(void (*)(void))w_%s, METH_FASTCALL, NULL},' % (name, name))
    out.append("static PyMethodDef methods[] = {")
    out += table
    out.append("    {NULL, NULL, 0, NULL}};")
    out.append('static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "%s", NULL, -1, methods};' % FASTCALL_NAME)
    out.append("PyMODINIT_FUNC PyInit_%s(void) { return PyModule_Create(&moddef); }" % FASTCALL_NAME)
    return "\n".join(out) + "\n"


def build_fastcall(force=False):
    """-> path of the extension module, or None when Python.h / gcc are unavailable (the ctypes binding then serves)."""
    import importlib.util
    import sysconfig
    target = fastcall_path()
    if not force and os.path.isfile(target) and os.path.isfile(LIB) and os.path.getmtime(target) >= os.path.getmtime(LIB):
        return target
    inc = sysconfig.get_paths().get("include", "")
    if not os.path.isfile(os.path.join(inc, "Python.h")):
        return None
    spec = importlib.util.spec_from_file_location("_b2a_lib_for_build", os.path.join(HERE, "_lib.py"))
    lib_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(lib_mod)
    src = os.path.join(OUT_DIR, FASTCALL_NAME + ".c")
    with open(src, "w") as f:
        f.write(_fastcall_source(lib_mod.parse_header()))
    cmd = [os.environ.get("CC", "gcc"), "-O2", "-shared", "-fPIC", "-I", inc, src, "-o", target, "-L", OUT_DIR, "-lb2a", "-Wl,-rpath,$ORIGIN"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        return None
    return target


# ----------------------------------------------------------------------------------------------------------------
# C++ autograd layer (csrc/autograd_nodes.cpp): torch::autograd::Function nodes over the C-ABI.  Host code only (g++ against the
# torch headers, linked to libb2a.so); optional - ops.py falls back to its Python `Function`s when the module is absent.
# ----------------------------------------------------------------------------------------------------------------
AUTOGRAD_NAME = "_b2a_autograd"


def autograd_path():
    import sysconfig
    return os.path.join(OUT_DIR, AUTOGRAD_NAME + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))


def build_autograd(force=False, verbose=False):
    import sysconfig
    src = os.path.join(SRC_DIR, "autograd_nodes.cpp")
    target = autograd_path()
    deps = [src, os.path.join(INCLUDE, "b2a.h")]
    if not force and os.path.isfile(target) and all(os.path.getmtime(target) >= os.path.getmtime(d) for d in deps):
        return target
    try:
        import torch
        from torch.utils import cpp_extension as ce
    except Exception:
        return None
    if not os.path.isfile(LIB):
        return None
    inc = [sysconfig.get_paths().get("include", "")] + ce.include_paths() + ["/usr/local/cuda/include"]
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ([os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden", "-DTORCH_EXTENSION_NAME=" + AUTOGRAD_NAME,
            "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
           + ["-I" + i for i in inc if i] + [src, "-o", target, "-L", OUT_DIR, "-lb2a", "-L", torch_lib, "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10",
                                             "-lc10_cuda", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + torch_lib])
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr[-6000:])
    if res.returncode != 0:
        return None
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
    print(build_fastcall(force="--force" in sys.argv))
    print(build_autograd(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

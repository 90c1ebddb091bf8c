"""ctypes binding of libgwbse_b200.so, generated from include/gwbse_b200.h.

The prototypes are parsed from the header so the Python side can never drift from the C ABI.
The library is required: there is no CPU fallback (loading or context creation fails loudly).
"""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gwbse_b200.h")
LIBPATH = os.path.join(ROOT, "votca_b200", "lib", "libgwbse_b200.so")
HOST_HEADER = os.path.join(ROOT, "include", "gwbse_host.h")
HOST_LIBPATH = os.path.join(ROOT, "votca_b200", "lib", "libgwbse_host.so")

_SCALARS = {
    "int": ctypes.c_int,
    "double": ctypes.c_double,
    "float": ctypes.c_float,
    "char": ctypes.c_char,
    "size_t": ctypes.c_size_t,
    "long long": ctypes.c_longlong,
    "void": None,
}


def parse_header(path=HEADER):
    """Returns {name: (restype_str, [(type_str, arg_name), ...])} for every prototype in the header."""
    txt = open(path).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"^\s*#.*$", "", txt, flags=re.M)
    txt = re.sub(r"typedef[^;]*\(\*[^;]*;", "", txt)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(gwbse_\w+)\s*\(([^;{]*?)\)\s*;", txt):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if ret.startswith("typedef"):
            continue
        arglist = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                mm = re.match(r"(.*?)(\w+)$", a)
                arglist.append((mm.group(1).strip(), mm.group(2)))
        protos[name] = (ret, arglist)
    return protos


def _ctype(tstr):
    t = tstr.replace("const", "").strip()
    if t.endswith("*"):
        base = t[:-1].strip()
        if base == "char":
            return ctypes.c_char_p
        return ctypes.c_void_p
    if t == "gwbse_ao3c_fn":
        return ctypes.c_void_p
    if t == "long":
        return ctypes.c_long
    return _SCALARS[t]


class CApi:
    def __init__(self, libpath=LIBPATH, header=HEADER):
        if not os.path.exists(libpath):
            raise RuntimeError(
                f"{libpath} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(gwbse_b200 has no CPU fallback)")
        self.lib = ctypes.CDLL(libpath, mode=ctypes.RTLD_GLOBAL)
        self.protos = parse_header(header)
        for name, (ret, args) in self.protos.items():
            fn = getattr(self.lib, name)
            fn.restype = _ctype(ret) if ret != "void" else None
            fn.argtypes = [_ctype(t) for t, _ in args]

    def __getattr__(self, name):
        return getattr(self.lib, name)


_api = None
_host_api = None


def capi():
    global _api
    if _api is None:
        _api = CApi()
    return _api


def host_api():
    """C++ host layer (GWBSE job facade); loads the kernel library first."""
    global _host_api
    if _host_api is None:
        capi()
        _host_api = CApi(HOST_LIBPATH, HOST_HEADER)
    return _host_api


def ptr(a):
    """Host pointer of a numpy array (must be float64/int32, any memory order handled by caller)."""
    if a is None:
        return None
    return a.ctypes.data_as(ctypes.c_void_p)


def fmat(a):
    """Column-major float64 copy/view."""
    return np.asfortranarray(np.asarray(a, dtype=np.float64))

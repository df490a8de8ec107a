import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}

c_ssize = ctypes.c_ssize_t


def build(want_ref=True):
    """Compile the checkers: liboracle.so always; _ref/ only where /root/reference exists."""
    subprocess.run(["make", "-s", "-C", _HERE, "liboracle.so"], check=True)
    if want_ref and os.path.exists("/root/reference/source/common/primitives.h"):
        subprocess.run(["make", "-s", "-j8", "-C", _HERE, "ref"], check=True)


def orc():
    """ctypes handle of the restatement (liboracle.so)."""
    if "orc" not in _cache:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build(want_ref=False)
        L = ctypes.CDLL(path)
        for name in ("orc_sse_pp", "orc_sse_ss", "orc_ssd_s"):
            getattr(L, name).restype = ctypes.c_uint64
        _cache["orc"] = L
    return _cache["orc"]


def have_ref(depth=8):
    return os.path.exists(os.path.join(_HERE, "_ref", "libx265ref%d.so" % (8 if depth == 8 else 10)))


def ref(depth=8):
    """ctypes handle of the compiled reference (oracle/_ref/libx265ref{8,10}.so) or None."""
    key = "ref%d" % (8 if depth == 8 else 10)
    if key not in _cache:
        path = os.path.join(_HERE, "_ref", "libx265%s.so" % key)
        if not os.path.exists(path):
            _cache[key] = None
        else:
            L = ctypes.CDLL(path)
            L.ref_sse.restype = ctypes.c_uint64
            L.ref_var.restype = ctypes.c_uint64
            L.ref_lambda.restype = ctypes.c_double
            L.ref_lambda2.restype = ctypes.c_double
            L.ref_dct_table.restype = ctypes.c_void_p
            L.ref_luma_filter.restype = ctypes.c_void_p
            L.ref_chroma_filter.restype = ctypes.c_void_p
            L.ref_init()
            _cache[key] = L
    return _cache[key]

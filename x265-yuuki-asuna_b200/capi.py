"""ctypes binding of libx265b200.so (include/x265b200.h) for the test-suite and bench.py.

This is plumbing only: the product is the C-ABI shared library built from csrc/*.cu.  The
loader fails loudly when the library is missing -- there is no Python or CPU fallback.
"""
import ctypes
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libx265b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "x265b200.h")

CMP_SAD, CMP_SATD, CMP_SA8D, CMP_SA8D8, CMP_SSE_PP, CMP_SSE_SS, CMP_SSD_S = range(7)

# LumaPU enum order of the reference (source/common/primitives.h:41-55)
LUMA_PU_SIZES = [
    (4, 4), (8, 8), (16, 16), (32, 32), (64, 64),
    (8, 4), (4, 8), (16, 8), (8, 16), (32, 16), (16, 32), (64, 32), (32, 64),
    (16, 12), (12, 16), (16, 4), (4, 16), (32, 24), (24, 32), (32, 8), (8, 32),
    (64, 48), (48, 64), (64, 16), (16, 64),
]


def partition_from_sizes(w, h):
    """LumaPU enum for a (w, h) block -- primitives.h:435 partitionFromSizes()."""
    return LUMA_PU_SIZES.index((w, h))


_lib = None


def header_symbols():
    """Every function name declared in include/x265b200.h."""
    src = open(HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(x265b200_\w+)\s*\(", src)))


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libx265b200.so is missing (%s); run `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    L.x265b200_last_error.restype = ctypes.c_char_p
    L.x265b200_stream.restype = ctypes.c_void_p
    L.x265b200_launch_count.restype = ctypes.c_uint64
    _lib = L
    return L


class X265B200Error(RuntimeError):
    pass


def _vp(x):
    """void* from a numpy array, a DevBuf, an int address or None."""
    if x is None:
        return ctypes.c_void_p(0)
    if isinstance(x, DevBuf):
        return ctypes.c_void_p(x.ptr)
    if isinstance(x, np.ndarray):
        return ctypes.c_void_p(x.ctypes.data)
    return ctypes.c_void_p(int(x))


class DevBuf:
    """A device allocation owned by a Ctx."""

    def __init__(self, ctx, nbytes):
        self.ctx, self.nbytes = ctx, int(nbytes)
        p = ctypes.c_void_p()
        ctx._chk(ctx.L.x265b200_malloc(ctx.h, ctypes.c_size_t(self.nbytes), ctypes.byref(p)))
        self.ptr = p.value

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        self.ctx._chk(self.ctx.L.x265b200_upload(self.ctx.h, _vp(self.ptr), _vp(arr), ctypes.c_size_t(arr.nbytes)))
        return self

    def download(self, dtype, count=None):
        dt = np.dtype(dtype)
        count = self.nbytes // dt.itemsize if count is None else count
        out = np.empty(count, dtype=dt)
        self.ctx._chk(self.ctx.L.x265b200_download(self.ctx.h, _vp(out), _vp(self.ptr), ctypes.c_size_t(out.nbytes)))
        return out

    def free(self):
        if self.ptr:
            self.ctx.L.x265b200_free(self.ctx.h, _vp(self.ptr))
            self.ptr = 0


class Ctx:
    """One backend context == one GPU + one stream (x265b200_create)."""

    def __init__(self, device=0, stream=None):
        self.L = load()
        h = ctypes.c_void_p()
        rc = self.L.x265b200_create(int(device), _vp(stream), ctypes.byref(h))
        if rc != 0:
            raise X265B200Error(self.L.x265b200_last_error().decode())
        self.h = h

    def _chk(self, rc):
        if rc != 0:
            raise X265B200Error(self.L.x265b200_last_error().decode())

    def close(self):
        if self.h:
            self.L.x265b200_destroy(self.h)
            self.h = None

    def sync(self):
        self._chk(self.L.x265b200_sync(self.h))

    @property
    def stream(self):
        return self.L.x265b200_stream(self.h)

    @property
    def launches(self):
        return int(self.L.x265b200_launch_count(self.h))

    def to_device(self, arr):
        arr = np.ascontiguousarray(arr)
        return DevBuf(self, max(arr.nbytes, 1)).upload(arr)

    def empty(self, nbytes):
        return DevBuf(self, max(int(nbytes), 1))

    # ---- block compare -----------------------------------------------------------------------
    def pixelcmp_host(self, kind, depth, w, h, A, strideA, B, strideB, offA, offB):
        A = np.ascontiguousarray(A)
        B = A if B is None else np.ascontiguousarray(B)
        offA = np.ascontiguousarray(offA, dtype=np.int64)
        offB = np.ascontiguousarray(offB if offB is not None else offA, dtype=np.int64)
        n = len(offA)
        out = np.empty(n, dtype=np.uint64 if kind >= CMP_SSE_PP else np.int32)
        self._chk(self.L.x265b200_pixelcmp_host(
            self.h, kind, depth, w, h, _vp(A), ctypes.c_size_t(A.nbytes), ctypes.c_int64(strideA),
            _vp(B), ctypes.c_size_t(B.nbytes), ctypes.c_int64(strideB), _vp(offA), _vp(offB),
            ctypes.c_int64(n), _vp(out)))
        return out

    def pixelcmp_dev(self, kind, depth, w, h, dA, strideA, dB, strideB, dOffA, dOffB, n, dOut, dMv=None, grid_cols=0):
        self._chk(self.L.x265b200_pixelcmp_dev(
            self.h, kind, depth, w, h, _vp(dA), ctypes.c_int64(strideA), _vp(dB), ctypes.c_int64(strideB),
            _vp(dOffA), _vp(dOffB), _vp(dMv), int(grid_cols), ctypes.c_int64(n), _vp(dOut)))

    def sad_xn_dev(self, depth, K, w, h, dFenc, fencBlockStride, dRef, refStride, dRefOff, n, dRes):
        self._chk(self.L.x265b200_sad_xn_dev(
            self.h, depth, K, w, h, _vp(dFenc), ctypes.c_int64(fencBlockStride), _vp(dRef), ctypes.c_int64(refStride),
            _vp(dRefOff), ctypes.c_int64(n), _vp(dRes)))

"""ctypes binding of libx265b200.so (include/x265b200.h) for the test-suite and bench.py.

This is plumbing only: the product is the C-ABI shared library built from csrc/*.cu.  The
loader fails loudly when the library is missing -- there is no Python or CPU fallback.
"""
import ctypes
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# X265B200_LIB selects another build of the SAME library (e.g. the experimental variant `make -C csrc exp` produces)
LIB_PATH = os.environ.get("X265B200_LIB") or os.path.join(_HERE, "libx265b200.so")
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


def sm_partition(device, sms_first):
    """x265b200_sm_partition: (streamFirst, streamRest, (smsFirst, smsRest)) -- raw cudaStream_t values"""
    L = load()
    a, b = ctypes.c_void_p(), ctypes.c_void_p()
    got = (ctypes.c_int * 2)()
    if L.x265b200_sm_partition(int(device), int(sms_first), ctypes.byref(a), ctypes.byref(b), got) != 0:
        raise X265B200Error(L.x265b200_last_error().decode())
    return a.value, b.value, (got[0], got[1])


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


# ---- transforms / interpolation / intra (appended bindings) ---------------------------------------
IP_HPP, IP_HPS, IP_VPP, IP_VPS, IP_VSP, IP_VSS, IP_HVPP, IP_P2S = range(8)

INTERP_JOB = np.dtype([("srcOff", np.int64), ("dstOff", np.int64), ("idxX", np.int32), ("idxY", np.int32)])
INTRA_JOB = np.dtype([("srcOff", np.int64), ("dstOff", np.int64), ("mode", np.int32), ("bFilter", np.int32)])


def dct_table(N):
    out = np.empty((N, N), dtype=np.int16)
    if load().x265b200_dct_table(int(N), _vp(out)) != 0:
        raise X265B200Error(load().x265b200_last_error().decode())
    return out


def _i64(v):
    return ctypes.c_int64(int(v))


def _ctx_dct_dev(self, sizeIdx, depth, dSrc, srcBlockStride, srcStride, dDst, n):
    self._chk(self.L.x265b200_dct_dev(self.h, sizeIdx, depth, _vp(dSrc), _i64(srcBlockStride), _i64(srcStride), _vp(dDst), _i64(n)))


def _ctx_idct_dev(self, sizeIdx, depth, dSrc, dDst, dstBlockStride, dstStride, n):
    self._chk(self.L.x265b200_idct_dev(self.h, sizeIdx, depth, _vp(dSrc), _vp(dDst), _i64(dstBlockStride), _i64(dstStride), _i64(n)))


def _ctx_quant_dev(self, dCoef, dQuantCoeff, dDeltaU, dQCoef, qBits, add, numCoeff, n, dNumSig, nquant=False):
    if nquant:
        self._chk(self.L.x265b200_nquant_dev(self.h, _vp(dCoef), _vp(dQuantCoeff), _vp(dQCoef), qBits, add, numCoeff, _i64(n), _vp(dNumSig)))
    else:
        self._chk(self.L.x265b200_quant_dev(self.h, _vp(dCoef), _vp(dQuantCoeff), _vp(dDeltaU), _vp(dQCoef), qBits, add, numCoeff, _i64(n), _vp(dNumSig)))


def _ctx_dequant_normal_dev(self, dQ, dCoef, num, n, scale, shift):
    self._chk(self.L.x265b200_dequant_normal_dev(self.h, _vp(dQ), _vp(dCoef), num, _i64(n), scale, shift))


def _ctx_dequant_scaling_dev(self, dQ, dDeq, dCoef, num, n, per, shift):
    self._chk(self.L.x265b200_dequant_scaling_dev(self.h, _vp(dQ), _vp(dDeq), _vp(dCoef), num, _i64(n), per, shift))


def _ctx_count_nonzero_dev(self, dQ, numCoeff, n, dOut):
    self._chk(self.L.x265b200_count_nonzero_dev(self.h, _vp(dQ), numCoeff, _i64(n), _vp(dOut)))


def _ctx_interp_dev(self, kind, taps, depth, w, h, dSrc, srcStride, dDst, dstStride, dJobs, n, isRowExt=0):
    self._chk(self.L.x265b200_interp_dev(self.h, kind, taps, depth, w, h, _vp(dSrc), _i64(srcStride), _vp(dDst), _i64(dstStride),
                                         _vp(dJobs), _i64(n), int(isRowExt)))


class INTERP_SEG(ctypes.Structure):
    """x265b200_interp_seg (include/x265b200.h)"""
    _fields_ = [("kind", ctypes.c_int32), ("w", ctypes.c_int32), ("h", ctypes.c_int32), ("isRowExt", ctypes.c_int32),
                ("src", ctypes.c_void_p), ("srcStride", ctypes.c_int64), ("dst", ctypes.c_void_p), ("dstStride", ctypes.c_int64),
                ("jobs", ctypes.c_void_p), ("n", ctypes.c_int64)]


def _ctx_interp_multi_dev(self, taps, depth, segs):
    """segs: list of (kind, w, h, dSrc, srcStride, dDst, dstStride, dJobs, n[, isRowExt])"""
    arr = (INTERP_SEG * len(segs))()
    for i, sg in enumerate(segs):
        kind, w, h, dSrc, srcStride, dDst, dstStride, dJobs, n = sg[:9]
        arr[i].kind, arr[i].w, arr[i].h, arr[i].isRowExt = int(kind), int(w), int(h), int(sg[9]) if len(sg) > 9 else 0
        arr[i].src, arr[i].srcStride, arr[i].dst, arr[i].dstStride = _vp(dSrc).value, int(srcStride), _vp(dDst).value, int(dstStride)
        arr[i].jobs, arr[i].n = _vp(dJobs).value, int(n)
    self._chk(self.L.x265b200_interp_multi_dev(self.h, int(taps), int(depth), arr, len(segs)))


def _ctx_intra_pred_dev(self, depth, log2N, dNbr, dDst, dstStride, dJobs, n):
    self._chk(self.L.x265b200_intra_pred_dev(self.h, depth, log2N, _vp(dNbr), _vp(dDst), _i64(dstStride), _vp(dJobs), _i64(n)))


def _ctx_intra_filter_dev(self, depth, log2N, dSrc, dDst, n):
    self._chk(self.L.x265b200_intra_filter_dev(self.h, depth, log2N, _vp(dSrc), _vp(dDst), _i64(n)))


def _ctx_intra_allangs_dev(self, depth, log2N, dRef, dFilt, dDest, bLuma, n):
    self._chk(self.L.x265b200_intra_allangs_dev(self.h, depth, log2N, _vp(dRef), _vp(dFilt), _vp(dDest), int(bLuma), _i64(n)))


Ctx.dct_dev = _ctx_dct_dev
Ctx.idct_dev = _ctx_idct_dev
Ctx.quant_dev = _ctx_quant_dev
Ctx.dequant_normal_dev = _ctx_dequant_normal_dev
Ctx.dequant_scaling_dev = _ctx_dequant_scaling_dev
Ctx.count_nonzero_dev = _ctx_count_nonzero_dev
Ctx.interp_multi_dev = _ctx_interp_multi_dev
Ctx.interp_dev = _ctx_interp_dev
Ctx.intra_pred_dev = _ctx_intra_pred_dev
Ctx.intra_filter_dev = _ctx_intra_filter_dev
Ctx.intra_allangs_dev = _ctx_intra_allangs_dev


# ---- glue ----------------------------------------------------------------------------------------
(GL_COPY_PP, GL_COPY_SS, GL_COPY_SP, GL_COPY_PS, GL_FILL_S, GL_CPY2DTO1D_SHL, GL_CPY2DTO1D_SHR, GL_CPY1DTO2D_SHL, GL_CPY1DTO2D_SHR,
 GL_SUB_PS, GL_ADD_PS, GL_ADDAVG, GL_PIXELAVG_PP, GL_TRANSPOSE, GL_WEIGHT_PP, GL_WEIGHT_SP) = range(16)
GLUE_JOB = np.dtype([("dstOff", np.int64), ("src0Off", np.int64), ("src1Off", np.int64)])


def _ctx_glue_dev(self, op, depth, w, h, dDst, dstStride, dSrc0, src0Stride, dSrc1, src1Stride, dJobs, n, p0=0, p1=0, p2=0, p3=0):
    self._chk(self.L.x265b200_glue_dev(self.h, int(op), int(depth), int(w), int(h), _vp(dDst), _i64(dstStride), _vp(dSrc0), _i64(src0Stride),
                                       _vp(dSrc1), _i64(src1Stride), _vp(dJobs), _i64(n), int(p0), int(p1), int(p2), int(p3)))


def _ctx_var_dev(self, depth, size, dSrc, stride, dOff, n, dOut):
    self._chk(self.L.x265b200_var_dev(self.h, int(depth), int(size), _vp(dSrc), _i64(stride), _vp(dOff), _i64(n), _vp(dOut)))


def _ctx_psy_cost_dev(self, depth, size, dSrc, sstride, dRec, rstride, dOffS, dOffR, n, dOut):
    self._chk(self.L.x265b200_psy_cost_dev(self.h, int(depth), int(size), _vp(dSrc), _i64(sstride), _vp(dRec), _i64(rstride),
                                           _vp(dOffS), _vp(dOffR), _i64(n), _vp(dOut)))


def _ctx_copy_cnt_dev(self, size, dCoeff, dResi, stride, dOff, n, dNumSig):
    self._chk(self.L.x265b200_copy_cnt_dev(self.h, int(size), _vp(dCoeff), _vp(dResi), _i64(stride), _vp(dOff), _i64(n), _vp(dNumSig)))


def _ctx_denoise_dct_dev(self, dCoef, dResSum, dOffset, numCoeff, n):
    self._chk(self.L.x265b200_denoise_dct_dev(self.h, _vp(dCoef), _vp(dResSum), _vp(dOffset), int(numCoeff), _i64(n)))


def _ctx_lowpass_dct_dev(self, sizeIdx, depth, dSrc, srcBlockStride, srcStride, dDst, n):
    self._chk(self.L.x265b200_lowpass_dct_dev(self.h, int(sizeIdx), int(depth), _vp(dSrc), _i64(srcBlockStride), _i64(srcStride), _vp(dDst), _i64(n)))


Ctx.glue_dev = _ctx_glue_dev
Ctx.var_dev = _ctx_var_dev
Ctx.psy_cost_dev = _ctx_psy_cost_dev
Ctx.copy_cnt_dev = _ctx_copy_cnt_dev
Ctx.denoise_dct_dev = _ctx_denoise_dct_dev
Ctx.lowpass_dct_dev = _ctx_lowpass_dct_dev


# ---- motion estimation -----------------------------------------------------------------------------
ME_DIA, ME_HEX, ME_UMH, ME_STAR, ME_SEA, ME_FULL = range(6)
ME_REFINE = 6      # MotionEstimate::refineMV (motion.cpp:606-737) through the me_batch entry points
ME_JOB = np.dtype([("puX", np.int32), ("puY", np.int32), ("w", np.int32), ("h", np.int32),
                   ("mvminX", np.int32), ("mvminY", np.int32), ("mvmaxX", np.int32), ("mvmaxY", np.int32),
                   ("mvpX", np.int32), ("mvpY", np.int32), ("numCand", np.int32), ("mvc", np.int32, (8, 2)),
                   ("refIdx", np.int32), ("outMvX", np.int32), ("outMvY", np.int32), ("outCost", np.int32)])


def lambda_for_qp(qp, depth=8):
    f = load().x265b200_lambda
    f.restype = ctypes.c_double
    return float(f(int(qp), int(depth)))


def bitcost_table(lam):
    out = np.empty(4 * 32768 + 1, dtype=np.uint16)
    if load().x265b200_bitcost_table(ctypes.c_double(lam), _vp(out)) != 0:
        raise X265B200Error(load().x265b200_last_error().decode())
    return out


def _ctx_me_batch_dev(self, depth, dFenc, fencStride, dRef, refStride, dJobs, n, maxW, maxH, searchMethod, subpelRefine,
                      merange, lam, maxSlices=1, dRefPlanes=None):
    self._chk(self.L.x265b200_me_batch_dev(self.h, depth, _vp(dFenc), _i64(fencStride), _vp(dRef), _vp(dRefPlanes), _i64(refStride),
                                           _vp(dJobs), _i64(n), int(maxW), int(maxH), int(searchMethod), int(subpelRefine),
                                           int(merange), ctypes.c_double(lam), int(maxSlices)))


Ctx.me_batch_dev = _ctx_me_batch_dev


class ME_CHROMA(ctypes.Structure):
    """x265b200_me_chroma (include/x265b200.h)"""
    _fields_ = [("csp", ctypes.c_int32), ("fencCb", ctypes.c_void_p), ("fencCr", ctypes.c_void_p), ("fencStrideC", ctypes.c_int64),
                ("refCb", ctypes.c_void_p), ("refCr", ctypes.c_void_p), ("refCbPlanes", ctypes.c_void_p), ("refCrPlanes", ctypes.c_void_p),
                ("refStrideC", ctypes.c_int64)]


def _ctx_me_batch_chroma_dev(self, depth, dFenc, fencStride, dRef, refStride, csp, dFencCb, dFencCr, fencStrideC, dRefCb, dRefCr, refStrideC,
                             dJobs, n, maxW, maxH, searchMethod, subpelRefine, merange, lam, maxSlices=1):
    ch = ME_CHROMA(int(csp), _vp(dFencCb), _vp(dFencCr), int(fencStrideC), _vp(dRefCb), _vp(dRefCr), None, None, int(refStrideC))
    self._chk(self.L.x265b200_me_batch_chroma_dev(self.h, depth, _vp(dFenc), _i64(fencStride), _vp(dRef), None, _i64(refStride),
                                                  ctypes.byref(ch), _vp(dJobs), _i64(n), int(maxW), int(maxH), int(searchMethod),
                                                  int(subpelRefine), int(merange), ctypes.c_double(lam), int(maxSlices)))


Ctx.me_batch_chroma_dev = _ctx_me_batch_chroma_dev


# ---- plane forms ---------------------------------------------------------------------------------
def _ctx_dct_plane_dev(self, sizeIdx, depth, dPlane, stride, blocksX, blocksY, dCoef):
    self._chk(self.L.x265b200_dct_plane_dev(self.h, sizeIdx, depth, _vp(dPlane), _i64(stride), int(blocksX), int(blocksY), _vp(dCoef)))


def _ctx_idct_plane_dev(self, sizeIdx, depth, dCoef, dPlane, stride, blocksX, blocksY):
    self._chk(self.L.x265b200_idct_plane_dev(self.h, sizeIdx, depth, _vp(dCoef), _vp(dPlane), _i64(stride), int(blocksX), int(blocksY)))


def _ctx_sub_ps_plane_dev(self, depth, dA, strideA, dB, strideB, dDst, dstStride, w, h):
    self._chk(self.L.x265b200_sub_ps_plane_dev(self.h, depth, _vp(dA), _i64(strideA), _vp(dB), _i64(strideB), _vp(dDst), _i64(dstStride), int(w), int(h)))


def _ctx_add_ps_plane_dev(self, depth, dDst, dstStride, dPred, predStride, dResi, resiStride, w, h):
    self._chk(self.L.x265b200_add_ps_plane_dev(self.h, depth, _vp(dDst), _i64(dstStride), _vp(dPred), _i64(predStride), _vp(dResi), _i64(resiStride), int(w), int(h)))


Ctx.dct_plane_dev = _ctx_dct_plane_dev
Ctx.idct_plane_dev = _ctx_idct_plane_dev
Ctx.sub_ps_plane_dev = _ctx_sub_ps_plane_dev
Ctx.add_ps_plane_dev = _ctx_add_ps_plane_dev


# ---- lookahead -------------------------------------------------------------------------------------
LA_TRIPLE = np.dtype([("b", np.int32), ("p0", np.int32), ("p1", np.int32), ("doSearch", np.int32, (2,)), ("mvSlot", np.int32, (2,)),
                      ("weightIdx0", np.int32), ("weightPlanes0", np.int32)])
LA_WEIGHT_JOB = np.dtype([("fencPlane0", np.uint64), ("intraCost", np.uint64), ("refBuffer", np.uint64, (4,)), ("weighted", np.uint64, (4,)),
                          ("fencSum", np.uint64), ("fencSsd", np.uint64), ("refSum", np.uint64), ("refSsd", np.uint64)])      # x265b200_la_weight_job
LA_WEIGHT = np.dtype([("isWeighted", np.int32), ("inputWeight", np.int32), ("log2WeightDenom", np.int32), ("inputOffset", np.int32),
                      ("origscore", np.uint32), ("score", np.uint32)])                                                        # x265b200_la_weight


def _ctx_lowres_init_dev(self, depth, dSrc, srcStride, planePtrs, dstStride, width, height, marginX, marginY):
    arr = (ctypes.c_void_p * 4)(*[int(p) for p in planePtrs])
    self._chk(self.L.x265b200_lowres_init_dev(self.h, depth, _vp(dSrc), _i64(srcStride), arr, _i64(dstStride), int(width), int(height), int(marginX), int(marginY)))


def _ctx_la_intra_dev(self, depth, dPlane0, stride, wcu, hcu, dInvQ, intraPenalty, dIntraCost, dIntraMode, dLowresCosts, dRowSatds, dSums):
    self._chk(self.L.x265b200_la_intra_dev(self.h, depth, _vp(dPlane0), _i64(stride), int(wcu), int(hcu), _vp(dInvQ), int(intraPenalty),
                                           _vp(dIntraCost), _vp(dIntraMode), _vp(dLowresCosts), _vp(dRowSatds), _vp(dSums)))


def _ctx_la_estimate_dev(self, depth, dPlanes, stride, wcu, hcu, triples, dMvPool, dMvCostPool, dIntraCostPtrs, dInvQPtrs,
                         dLowresCosts, dRowSatds, dSums, lam, lookaheadSlices=0, dWeights=None):
    triples = np.ascontiguousarray(triples, dtype=LA_TRIPLE)
    self._chk(self.L.x265b200_la_estimate_dev(self.h, depth, _vp(dPlanes), _i64(stride), int(wcu), int(hcu), _vp(triples), len(triples),
                                              _vp(dMvPool), _vp(dMvCostPool), _vp(dIntraCostPtrs), _vp(dInvQPtrs), _vp(dLowresCosts),
                                              _vp(dRowSatds), _vp(dSums), ctypes.c_double(lam), int(lookaheadSlices), _vp(dWeights)))


def la_weight_guess(depth, width, lines, fencSum, fencSsd, refSum, refSsd):
    """x265b200_la_weight_guess (host only): dict(measure, curScale, curDenom, curOffset, finScale, finDenom, identity)"""
    L = load()
    out = (ctypes.c_int32 * 7)()
    if L.x265b200_la_weight_guess(int(depth), int(width), int(lines), ctypes.c_uint64(fencSum), ctypes.c_uint64(fencSsd),
                                  ctypes.c_uint64(refSum), ctypes.c_uint64(refSsd), out) != 0:
        raise X265B200Error(L.x265b200_last_error().decode())
    return dict(zip(("measure", "curScale", "curDenom", "curOffset", "finScale", "finDenom", "identity"), [int(v) for v in out]))


def _ctx_la_weights_analyse_dev(self, depth, jobs, stride, paddedLines, padOffset, width, lines, dOut):
    """jobs: numpy array of LA_WEIGHT_JOB (host); dOut: device array of LA_WEIGHT records"""
    j = np.ascontiguousarray(jobs, dtype=LA_WEIGHT_JOB)
    self._chk(self.L.x265b200_la_weights_analyse_dev(self.h, int(depth), _vp(j), len(j), _i64(stride), int(paddedLines), _i64(padOffset),
                                                     int(width), int(lines), _vp(dOut)))


Ctx.la_weights_analyse_dev = _ctx_la_weights_analyse_dev
Ctx.lowres_init_dev = _ctx_lowres_init_dev
Ctx.la_intra_dev = _ctx_la_intra_dev
Ctx.la_estimate_dev = _ctx_la_estimate_dev


def _ctx_sad_pyramid_dev(self, depth, dCur, strideCur, dRefPtrs, numRefs, strideRef, ctuCols, ctuRows, dMvCtu, dOut8, dOut16, dOut32, dOut64):
    self._chk(self.L.x265b200_sad_pyramid_dev(self.h, depth, _vp(dCur), _i64(strideCur), _vp(dRefPtrs), int(numRefs), _i64(strideRef), int(ctuCols), int(ctuRows),
                                              _vp(dMvCtu), _vp(dOut8), _vp(dOut16), _vp(dOut32), _vp(dOut64)))


Ctx.sad_pyramid_dev = _ctx_sad_pyramid_dev


SAD_GROUP = np.dtype([("cur", np.int32), ("ref", np.int32, (8,))])          # x265b200_sad_group


def _ctx_sad_stream_dev(self, depth, poolOrigin, framePitch, stride, marginX, marginY, rowsTotal, numFrames, ctuCols, ctuRows, groups, numRefs,
                        dOut8, dOut16, dOut32, dOut64):
    """groups: numpy array of SAD_GROUP (host)"""
    g = np.ascontiguousarray(groups, dtype=SAD_GROUP)
    self._chk(self.L.x265b200_sad_stream_dev(self.h, int(depth), _vp(poolOrigin), _i64(framePitch), _i64(stride), int(marginX), int(marginY), int(rowsTotal),
                                             int(numFrames), int(ctuCols), int(ctuRows), _vp(g), len(g), int(numRefs), _vp(dOut8), _vp(dOut16), _vp(dOut32), _vp(dOut64)))


Ctx.sad_stream_dev = _ctx_sad_stream_dev


def _ctx_me_frame_dev(self, depth, dCur, curStride, refOrigins, refStride, marginX, marginY, rowsTotal, ctuCols, ctuRows, puMask,
                      dMvpCtu, searchMethod, subpelRefine, merange, lam, dOut):
    arr = (ctypes.c_void_p * len(refOrigins))(*[int(p) for p in refOrigins])
    self._chk(self.L.x265b200_me_frame_dev(self.h, depth, _vp(dCur), _i64(curStride), arr, len(refOrigins), _i64(refStride), int(marginX), int(marginY),
                                           int(rowsTotal), int(ctuCols), int(ctuRows), int(puMask), _vp(dMvpCtu), int(searchMethod), int(subpelRefine),
                                           int(merange), ctypes.c_double(lam), _vp(dOut)))


Ctx.me_frame_dev = _ctx_me_frame_dev


class ME_FRAME_PARAMS(ctypes.Structure):      # x265b200_me_frame_params
    _fields_ = [(k, ctypes.c_int32) for k in ("depth", "ctuSize", "minCuSize", "rect", "amp", "picWidth", "picHeight", "ctuCols", "ctuRows",
                                              "marginX", "marginY", "rowsTotal", "chromaMarginX", "chromaMarginY", "numRefs", "searchMethod", "subpelRefine", "merange", "csp",
                                              "maxCand", "maxSlices", "frameParallel", "firstCtuRow", "sliceTotalRows", "refLagPixels")] + [("lambda_", ctypes.c_double)]


class ME_FRAME_PLANES(ctypes.Structure):      # x265b200_me_frame_planes
    _fields_ = [("curY", ctypes.c_void_p), ("curCb", ctypes.c_void_p), ("curCr", ctypes.c_void_p), ("curStride", ctypes.c_int64), ("curStrideC", ctypes.c_int64),
                ("refY", ctypes.POINTER(ctypes.c_void_p)), ("refCb", ctypes.POINTER(ctypes.c_void_p)), ("refCr", ctypes.POINTER(ctypes.c_void_p)),
                ("refStride", ctypes.c_int64), ("refStrideC", ctypes.c_int64)]


def me_frame_layout(ctuSize, minCuSize, rect, amp):
    """x265b200_me_frame_layout: [nPU][4] = {x, y, w, h} inside the CTU, in the order of the frame search's arrays."""
    L = load()
    n = L.x265b200_me_frame_layout(int(ctuSize), int(minCuSize), int(rect), int(amp), None, 0)
    if n < 0:
        raise X265B200Error(L.x265b200_last_error().decode())
    out = np.zeros((n, 4), dtype=np.int32)
    L.x265b200_me_frame_layout(int(ctuSize), int(minCuSize), int(rect), int(amp), out.ctypes.data_as(ctypes.c_void_p), n)
    return out


def _ctx_me_frame_ex_dev(self, params, curY, curStride, refY, refStride, dOut, dMvpCtu=None, dMvpPu=None, dNumCand=None, dMvc=None,
                         curC=None, curStrideC=0, refCb=None, refCr=None, refStrideC=0):
    """params: dict of x265b200_me_frame_params fields (lambda under 'lambda'); curY / curC = device addresses of plane origins;
    refY / refCb / refCr = lists of device addresses."""
    P, pl, _keep = _me_frame_structs(params, curY, curStride, refY, refStride, curC, curStrideC, refCb, refCr, refStrideC)
    self._chk(self.L.x265b200_me_frame_ex_dev(self.h, ctypes.byref(P), ctypes.byref(pl), _vp(dMvpCtu), _vp(dMvpPu), _vp(dNumCand), _vp(dMvc), _vp(dOut)))


Ctx.me_frame_ex_dev = _ctx_me_frame_ex_dev


def _me_frame_structs(params, curY, curStride, refY, refStride, curC, curStrideC, refCb, refCr, refStrideC):
    P = ME_FRAME_PARAMS()
    for k, v in params.items():
        setattr(P, "lambda_" if k == "lambda" else k, v)
    P.numRefs = len(refY)
    arrs = [(ctypes.c_void_p * len(refY))(*[int(x) for x in lst]) if lst is not None else None for lst in (refY, refCb, refCr)]
    pl = ME_FRAME_PLANES()
    pl.curY = int(curY); pl.curStride = int(curStride); pl.curStrideC = int(curStrideC)
    pl.curCb = int(curC[0]) if curC is not None else None; pl.curCr = int(curC[1]) if curC is not None else None
    pl.refY = ctypes.cast(arrs[0], ctypes.POINTER(ctypes.c_void_p))
    pl.refCb = ctypes.cast(arrs[1], ctypes.POINTER(ctypes.c_void_p)) if arrs[1] is not None else None
    pl.refCr = ctypes.cast(arrs[2], ctypes.POINTER(ctypes.c_void_p)) if arrs[2] is not None else None
    pl.refStride = int(refStride); pl.refStrideC = int(refStrideC)
    return P, pl, arrs


def _ctx_me_frame_ex_host(self, params, curY, curStride, refY, refStride, hostY, devYBase, bytesY, devOut, hostOut, outBytes,
                          dMvpCtu=None, dMvpPu=None, dNumCand=None, dMvc=None, curC=None, curStrideC=0, refCb=None, refCr=None, refStrideC=0,
                          hostC=None, devCBase=None, bytesC=0):
    """x265b200_me_frame_ex_host: hostY / hostC = host addresses of the padded planes (hostC = (Cb, Cr) or None), devYBase / devCBase
    their device destinations; hostOut = host address of the result buffer."""
    P, pl, _keep = _me_frame_structs(params, curY, curStride, refY, refStride, curC, curStrideC, refCb, refCr, refStrideC)
    hc = hostC if hostC is not None else (None, None)
    dc = devCBase if devCBase is not None else (None, None)
    self._chk(self.L.x265b200_me_frame_ex_host(self.h, ctypes.byref(P), ctypes.byref(pl), _vp(hostY), _vp(devYBase), ctypes.c_size_t(bytesY),
                                               _vp(hc[0]), _vp(dc[0]), _vp(hc[1]), _vp(dc[1]), ctypes.c_size_t(bytesC),
                                               _vp(dMvpCtu), _vp(dMvpPu), _vp(dNumCand), _vp(dMvc), _vp(devOut), _vp(hostOut), ctypes.c_size_t(outBytes)))


Ctx.me_frame_ex_host = _ctx_me_frame_ex_host


def _ctx_me_frame_ex_host_begin(self, params, curY, curStride, refY, refStride, hostY, devYBase, bytesY, devOut, hostOut, outBytes,
                                dMvpCtu=None, dMvpPu=None, dNumCand=None, dMvc=None, curC=None, curStrideC=0, refCb=None, refCr=None, refStrideC=0,
                                hostC=None, devCBase=None, bytesC=0):
    """x265b200_me_frame_ex_host_begin (same arguments as me_frame_ex_host); finish with me_frame_host_end()"""
    P, pl, _keep = _me_frame_structs(params, curY, curStride, refY, refStride, curC, curStrideC, refCb, refCr, refStrideC)
    hc = hostC if hostC is not None else (None, None)
    dc = devCBase if devCBase is not None else (None, None)
    self._chk(self.L.x265b200_me_frame_ex_host_begin(self.h, ctypes.byref(P), ctypes.byref(pl), _vp(hostY), _vp(devYBase), ctypes.c_size_t(bytesY),
                                                     _vp(hc[0]), _vp(dc[0]), _vp(hc[1]), _vp(dc[1]), ctypes.c_size_t(bytesC),
                                                     _vp(dMvpCtu), _vp(dMvpPu), _vp(dNumCand), _vp(dMvc), _vp(devOut), _vp(hostOut), ctypes.c_size_t(outBytes)))


def _ctx_me_frame_host_begin(self, depth, hostCurBase, planeBytes, devCurBase, curStride, refOrigins, refStride, marginX, marginY, rowsTotal, ctuCols, ctuRows,
                             puMask, dMvpCtu, searchMethod, subpelRefine, merange, lam, devOut, hostOut, outBytes):
    arr = (ctypes.c_void_p * len(refOrigins))(*[int(p) for p in refOrigins])
    self._chk(self.L.x265b200_me_frame_host_begin(self.h, depth, _vp(hostCurBase), ctypes.c_size_t(planeBytes), _vp(devCurBase), _i64(curStride), arr, len(refOrigins),
                                                  _i64(refStride), int(marginX), int(marginY), int(rowsTotal), int(ctuCols), int(ctuRows), int(puMask), _vp(dMvpCtu),
                                                  int(searchMethod), int(subpelRefine), int(merange), ctypes.c_double(lam), _vp(devOut), _vp(hostOut), ctypes.c_size_t(outBytes)))


def _ctx_me_frame_host_end(self):
    self._chk(self.L.x265b200_me_frame_host_end(self.h))


Ctx.me_frame_ex_host_begin = _ctx_me_frame_ex_host_begin
Ctx.me_frame_host_begin = _ctx_me_frame_host_begin
Ctx.me_frame_host_end = _ctx_me_frame_host_end


def _ctx_me_frame_host(self, depth, hostCurBase, planeBytes, devCurBase, curStride, refOrigins, refStride, marginX, marginY, rowsTotal, ctuCols, ctuRows,
                       puMask, dMvpCtu, searchMethod, subpelRefine, merange, lam, devOut, hostOut, outBytes):
    arr = (ctypes.c_void_p * len(refOrigins))(*[int(p) for p in refOrigins])
    self._chk(self.L.x265b200_me_frame_host(self.h, depth, _vp(hostCurBase), ctypes.c_size_t(planeBytes), _vp(devCurBase), _i64(curStride), arr, len(refOrigins),
                                            _i64(refStride), int(marginX), int(marginY), int(rowsTotal), int(ctuCols), int(ctuRows), int(puMask), _vp(dMvpCtu),
                                            int(searchMethod), int(subpelRefine), int(merange), ctypes.c_double(lam), _vp(devOut), _vp(hostOut), ctypes.c_size_t(outBytes)))


Ctx.me_frame_host = _ctx_me_frame_host


# ---- --me sea --------------------------------------------------------------------------------------
ADS_JOB = np.dtype([("sumsOff", np.int64), ("thresh", np.int32), ("encDC", np.int32, (4,))], align=True)      # 32 bytes, as the C struct
SEA_PLANE_W = [32, 32, 32, 24, 16, 16, 16, 12, 8, 8, 4, 4]       # FrameData::m_meIntegral order (framedata.h:171)
SEA_PLANE_H = [32, 24, 8, 32, 16, 12, 4, 16, 32, 8, 16, 4]


def _ctx_sea_integral_dev(self, depth, dReconOrigin, stride, padX, padY, maxHeight, planeOrigins):
    arr = (ctypes.c_void_p * 12)(*[int(p) for p in planeOrigins])
    self._chk(self.L.x265b200_sea_integral_dev(self.h, depth, _vp(dReconOrigin), _i64(stride), int(padX), int(padY), int(maxHeight), arr))


def _ctx_integral_inith_dev(self, depth, width, dSum, dPix, stride):
    self._chk(self.L.x265b200_integral_inith_dev(self.h, depth, int(width), _vp(dSum), _vp(dPix), _i64(stride)))


def _ctx_integral_initv_dev(self, height, dSum, stride):
    self._chk(self.L.x265b200_integral_initv_dev(self.h, int(height), _vp(dSum), _i64(stride)))


def _ctx_ads_dev(self, kind, lxHalf, dSums, delta, dCostMvX, width, dJobs, n, dMvs, dCounts):
    self._chk(self.L.x265b200_ads_dev(self.h, int(kind), int(lxHalf), _vp(dSums), _i64(delta), _vp(dCostMvX), int(width), _vp(dJobs), _i64(n),
                                      _vp(dMvs), _vp(dCounts)))


def _ctx_me_batch_sea_dev(self, depth, dFenc, fencStride, dRef, refStride, dIntegralPlanes, dJobs, n, maxW, maxH, subpelRefine,
                          merange, lam, maxSlices=1, dRefPlanes=None):
    self._chk(self.L.x265b200_me_batch_sea_dev(self.h, depth, _vp(dFenc), _i64(fencStride), _vp(dRef), _vp(dRefPlanes), _i64(refStride),
                                               _vp(dIntegralPlanes), _vp(dJobs), _i64(n), int(maxW), int(maxH), int(subpelRefine),
                                               int(merange), ctypes.c_double(lam), int(maxSlices)))


Ctx.sea_integral_dev = _ctx_sea_integral_dev
Ctx.integral_inith_dev = _ctx_integral_inith_dev
Ctx.integral_initv_dev = _ctx_integral_initv_dev
Ctx.ads_dev = _ctx_ads_dev
Ctx.me_batch_sea_dev = _ctx_me_batch_sea_dev


def _ctx_tu_pipeline_dev(self, sizeIdx, depth, useDST, dFenc, fencStride, dPred, predStride, dRecon, reconStride, blocksX, blocksY,
                         dQuantCoeff, qBits, add, dDequantCoef, scaleOrPer, dqShift, dCoeff, dNumSig, dSse):
    self._chk(self.L.x265b200_tu_pipeline_dev(self.h, int(sizeIdx), int(depth), int(useDST), _vp(dFenc), _i64(fencStride), _vp(dPred), _i64(predStride),
                                              _vp(dRecon), _i64(reconStride), int(blocksX), int(blocksY), _vp(dQuantCoeff), int(qBits), int(add),
                                              _vp(dDequantCoef), int(scaleOrPer), int(dqShift), _vp(dCoeff), _vp(dNumSig), _vp(dSse)))


Ctx.tu_pipeline_dev = _ctx_tu_pipeline_dev


# ---- motion compensation driver -------------------------------------------------------------------------
MC_JOB = np.dtype([("puX", np.int32), ("puY", np.int32), ("w", np.int32), ("h", np.int32), ("cuX", np.int32), ("cuY", np.int32),
                   ("refIdx", np.int32, (2,)), ("mv", np.int32, (2, 2))])
MC_WEIGHT = np.dtype([("w", np.int32), ("o", np.int32), ("shift", np.int32), ("present", np.int32)])


class MC_DESC(ctypes.Structure):
    """x265b200_mc_desc (include/x265b200.h)"""
    _fields_ = [("csp", ctypes.c_int32), ("isPSlice", ctypes.c_int32), ("weightedPred", ctypes.c_int32), ("weightedBiPred", ctypes.c_int32),
                ("picWidth", ctypes.c_int32), ("picHeight", ctypes.c_int32), ("maxCUSize", ctypes.c_int32), ("maxRefs", ctypes.c_int32),
                ("refs", ctypes.c_void_p), ("refStrideY", ctypes.c_int64), ("refStrideC", ctypes.c_int64),
                ("predY", ctypes.c_void_p), ("predCb", ctypes.c_void_p), ("predCr", ctypes.c_void_p),
                ("predStrideY", ctypes.c_int64), ("predStrideC", ctypes.c_int64), ("weights", ctypes.c_void_p)]


def _ctx_mc_dev(self, depth, desc, dJobs, n, bLuma=1, bChroma=1):
    self._chk(self.L.x265b200_mc_dev(self.h, int(depth), ctypes.byref(desc), _vp(dJobs), _i64(n), int(bLuma), int(bChroma)))


Ctx.mc_dev = _ctx_mc_dev


# ---- --hme lookahead ----------------------------------------------------------------------------------
class LA_HME(ctypes.Structure):
    """x265b200_la_hme (include/x265b200.h)"""
    _fields_ = [("lowerPlanes", ctypes.c_void_p), ("lowerStride", ctypes.c_int64), ("width4", ctypes.c_int32), ("height4", ctypes.c_int32),
                ("lowerMvPool", ctypes.c_void_p), ("lowerMvCostPool", ctypes.c_void_p),
                ("searchMethod", ctypes.c_int32 * 2), ("range", ctypes.c_int32 * 2)]


def _ctx_la_estimate_hme_dev(self, depth, dPlanes, stride, wcu, hcu, hme, triples, dMvPool, dMvCostPool, dIntraCostPtrs, dInvQPtrs,
                             dLowresCosts, dRowSatds, dSums, lam, lookaheadSlices=0, dWeights=None):
    triples = np.ascontiguousarray(triples, dtype=LA_TRIPLE)
    self._chk(self.L.x265b200_la_estimate_hme_dev(self.h, depth, _vp(dPlanes), _i64(stride), int(wcu), int(hcu), ctypes.byref(hme), _vp(triples),
                                                  len(triples), _vp(dMvPool), _vp(dMvCostPool), _vp(dIntraCostPtrs), _vp(dInvQPtrs),
                                                  _vp(dLowresCosts), _vp(dRowSatds), _vp(dSums), ctypes.c_double(lam), int(lookaheadSlices), _vp(dWeights)))


Ctx.la_estimate_hme_dev = _ctx_la_estimate_hme_dev


def _ctx_intra_modes_dev(self, depth, log2N, dNbr, dDest, bLuma, n):
    self._chk(self.L.x265b200_intra_modes_dev(self.h, int(depth), int(log2N), _vp(dNbr), _vp(dDest), int(bLuma), _i64(n)))


Ctx.intra_modes_dev = _ctx_intra_modes_dev


def _ctx_ssim_4x4x2_dev(self, depth, d1, s1, d2, s2, dOff1, dOff2, n, dSums):
    self._chk(self.L.x265b200_ssim_4x4x2_dev(self.h, int(depth), _vp(d1), _i64(s1), _vp(d2), _i64(s2), _vp(dOff1), _vp(dOff2), _i64(n), _vp(dSums)))


def _ctx_ssim_end4_dev(self, depth, dSum0, dSum1, dWidths, n, dOut):
    self._chk(self.L.x265b200_ssim_end4_dev(self.h, int(depth), _vp(dSum0), _vp(dSum1), _vp(dWidths), _i64(n), _vp(dOut)))


def _ctx_plane_clip_max_dev(self, depth, dSrc, stride, width, height, minPix, maxPix, dSum, dMax):
    self._chk(self.L.x265b200_plane_clip_max_dev(self.h, int(depth), _vp(dSrc), _i64(stride), int(width), int(height), int(minPix), int(maxPix), _vp(dSum), _vp(dMax)))


def _ctx_propagate_cost_dev(self, dDst, dIn, dIntra, dInter, dInvQ, fps, n):
    self._chk(self.L.x265b200_propagate_cost_dev(self.h, _vp(dDst), _vp(dIn), _vp(dIntra), _vp(dInter), _vp(dInvQ), ctypes.c_double(fps), _i64(n)))


Ctx.ssim_4x4x2_dev = _ctx_ssim_4x4x2_dev
Ctx.ssim_end4_dev = _ctx_ssim_end4_dev
Ctx.plane_clip_max_dev = _ctx_plane_clip_max_dev
Ctx.propagate_cost_dev = _ctx_propagate_cost_dev


SAO_E0, SAO_E1, SAO_E1_2ROWS, SAO_E2, SAO_E3, SAO_B0 = range(6)
SAO_BO = SAO_B0
SAO_JOB = np.dtype([("recOff", np.int64), ("diffOff", np.int64), ("buf0", np.int64), ("buf1", np.int64), ("offsetOff", np.int64),
                    ("width", np.int32), ("height", np.int32), ("startX", np.int32), ("pad", np.int32)])
DEBLOCK_JOB = np.dtype([("srcOff", np.int64), ("srcStep", np.int64), ("offset", np.int64), ("tcP", np.int32), ("tcQ", np.int32),
                        ("maskQ", np.int32), ("pad", np.int32)])


def _ctx_sao_apply_dev(self, kind, depth, dRec, stride, dJobs, n, dBuf, dOffsets, maxWidth):
    self._chk(self.L.x265b200_sao_apply_dev(self.h, int(kind), int(depth), _vp(dRec), _i64(stride), _vp(dJobs), _i64(n), _vp(dBuf), _vp(dOffsets), int(maxWidth)))


def _ctx_sao_stats_dev(self, kind, depth, dDiff, dRec, stride, dJobs, n, dBuf, dStats, dCount):
    self._chk(self.L.x265b200_sao_stats_dev(self.h, int(kind), int(depth), _vp(dDiff), _vp(dRec), _i64(stride), _vp(dJobs), _i64(n), _vp(dBuf), _vp(dStats), _vp(dCount)))


def _ctx_deblock_dev(self, chroma, depth, dPic, dJobs, n):
    self._chk(self.L.x265b200_deblock_dev(self.h, int(chroma), int(depth), _vp(dPic), _vp(dJobs), _i64(n)))


Ctx.sao_apply_dev = _ctx_sao_apply_dev
Ctx.sao_stats_dev = _ctx_sao_stats_dev
Ctx.deblock_dev = _ctx_deblock_dev


def _ctx_cutree_propagate_dev(self, wcu, hcu, dPropB, dIntra, dLowresCosts, dInvQ, dMvs0, dMvs1, dRef0, dRef1, bipredWeight, fpsFactor):
    self._chk(self.L.x265b200_cutree_propagate_dev(self.h, int(wcu), int(hcu), _vp(dPropB), _vp(dIntra), _vp(dLowresCosts), _vp(dInvQ), _vp(dMvs0), _vp(dMvs1),
                                                   _vp(dRef0), _vp(dRef1), int(bipredWeight), ctypes.c_double(fpsFactor)))


Ctx.cutree_propagate_dev = _ctx_cutree_propagate_dev


def _ctx_aq_energy_dev(self, depth, csp, qgSize, dY, strideY, dCb, dCr, strideC, picW, picH, dEnergy, dWp):
    self._chk(self.L.x265b200_aq_energy_dev(self.h, int(depth), int(csp), int(qgSize), _vp(dY), _i64(strideY), _vp(dCb), _vp(dCr), _i64(strideC),
                                            int(picW), int(picH), _vp(dEnergy), _vp(dWp)))


def _ctx_apply_weight_dev(self, depth, dSrc, dDst, stride, width, height, marginX, marginY, weight, offset, log2Denom):
    self._chk(self.L.x265b200_apply_weight_dev(self.h, int(depth), _vp(dSrc), _vp(dDst), _i64(stride), int(width), int(height), int(marginX), int(marginY),
                                               int(weight), int(offset), int(log2Denom)))


Ctx.aq_energy_dev = _ctx_aq_energy_dev
Ctx.apply_weight_dev = _ctx_apply_weight_dev

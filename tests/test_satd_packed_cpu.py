"""CPU: the packed-word 4x4 SATD used inside the frame-search kernel (csrc/satd_packed.cuh) is plain integer code that
also compiles for the host; here the very same header is built with g++ and compared with the oracle's satd
(pixel.cpp:210-232 restatement) on random, flat, extreme and checkerboard cells."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np

from util import oracle, vp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "x265-yuuki-asuna_b200", "csrc")

HARNESS = r'''
#include "satd_packed.cuh"
#include <string.h>
extern "C" void sp_satd4x4_batch(const uint8_t* f, const uint8_t* o, int n, int* out)
{
    for (int b = 0; b < n; b++)
    {
        uint32_t fw[4], ow[4];
        memcpy(fw, f + 16 * b, 16); memcpy(ow, o + 16 * b, 16);     /* little endian: pixel k in byte k */
        out[b] = satd4x4_packed_u8(fw, ow);
    }
}
extern "C" uint32_t sp_prmt_host(uint32_t a, uint32_t b, uint32_t s) { return sp_prmt(a, b, s); }
'''


def _build(tmp):
    src, so = os.path.join(tmp, "h.cpp"), os.path.join(tmp, "h.so")
    open(src, "w").write(HARNESS)
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-I", CSRC, "-o", so, src], check=True)
    return ctypes.CDLL(so)


def test_packed_satd_equals_oracle():
    with tempfile.TemporaryDirectory() as tmp:
        L = _build(tmp)
        rng = np.random.default_rng(7)
        cells_f = [rng.integers(0, 256, (4000, 16), dtype=np.uint8)]
        cells_o = [rng.integers(0, 256, (4000, 16), dtype=np.uint8)]
        # extremes: every 0/255 source pattern against its complement and against flat cells; checkerboards; near-equal cells
        pat = ((np.arange(65536)[:, None] >> np.arange(16)[None, :]) & 1).astype(np.uint8) * 255
        cells_f += [pat, pat, pat]
        cells_o += [255 - pat, np.zeros_like(pat), np.full_like(pat, 255)]
        near = rng.integers(0, 256, (4000, 16), dtype=np.uint8)
        cells_f.append(near)
        cells_o.append(np.clip(near.astype(np.int32) + rng.integers(-2, 3, near.shape), 0, 255).astype(np.uint8))
        f = np.ascontiguousarray(np.concatenate(cells_f)); o = np.ascontiguousarray(np.concatenate(cells_o))
        n = len(f)
        got = np.empty(n, dtype=np.int32)
        L.sp_satd4x4_batch(vp(f), vp(o), n, vp(got))
        O = oracle.orc()
        O.orc_satd.restype = ctypes.c_int
        want = np.array([O.orc_satd(8, 4, 4, vp(f[i]), ctypes.c_ssize_t(4), vp(o[i]), ctypes.c_ssize_t(4)) for i in range(n)], dtype=np.int32)
        assert np.array_equal(got, want)
        assert int(got.max()) > 2040                     # beyond a single-coefficient cell: the lanes really carry large sums


def test_prmt_emulation_modes():
    """host PRMT model: plain byte selects and the sign-replicate bit (the kernel's selectors use both forms)."""
    with tempfile.TemporaryDirectory() as tmp:
        L = _build(tmp)
        L.sp_prmt_host.restype = ctypes.c_uint32
        a, b = 0x80F17E03, 0x44332211
        assert L.sp_prmt_host(a, b, 0x3210) == a
        assert L.sp_prmt_host(a, b, 0x7654) == b
        assert L.sp_prmt_host(a, 0, 0x4240) == 0x00F10003
        assert L.sp_prmt_host(a, 0, 0x4341) == 0x0080007E
        assert L.sp_prmt_host(a, b, 0xBA98) == 0xFFFF0000 | 0x0000     # signs of bytes 0..3 of a: 03 -> 00, 7E -> 00, F1 -> FF, 80 -> FF

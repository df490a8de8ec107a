"""CPU: the staged packed-word sub-pel arithmetic (csrc/subpel_packed.cuh: horizontal rows of 4, vertical 4x4 cells from
11 source rows, folded rounding / one-instruction clip) compiled for the host and compared with the oracle's
luma_hpp / luma_vpp (ipfilter.cpp:79-118, 164-203) on random and extreme 8-bit data, all three fractions."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np

from util import oracle, vp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "x265-yuuki-asuna_b200", "csrc")
FILTER = np.array([[0, 0, 0, 64, 0, 0, 0, 0], [-1, 4, -10, 58, 17, -5, 1, 0], [-1, 4, -11, 40, 40, -11, 4, -1], [0, 1, -5, 17, 58, -10, 4, -1]], dtype=np.int16)

HARNESS = r'''
#include "subpel_packed.cuh"
#include <string.h>
/* src: n blocks of 11 rows x 12 pixels (row stride 12); the cell is rows 3..6, columns 3..6 */
extern "C" void sp_hpp_cells(const uint8_t* src, int n, const int16_t* taps, uint8_t* out)
{
    const uint32_t clo = sp_taps(taps, 0), chi = sp_taps(taps, 4);
    for (int b = 0; b < n; b++)
        for (int i = 0; i < 4; i++)
        {
            uint32_t w[3];
            memcpy(w, src + (size_t)b * 132 + (3 + i) * 12, 12);
            const uint32_t o = hpp_row4_u8(w, clo, chi);
            memcpy(out + (size_t)b * 16 + 4 * i, &o, 4);
        }
}
extern "C" void sp_vpp_cells(const uint8_t* src, int n, const int16_t* taps, uint8_t* out)
{
    const uint32_t clo = sp_taps(taps, 0), chi = sp_taps(taps, 4);
    for (int b = 0; b < n; b++)
    {
        uint32_t r[11], o[4];
        for (int j = 0; j < 11; j++) memcpy(&r[j], src + (size_t)b * 132 + j * 12 + 3, 4);
        vpp_cell_u8(r, clo, chi, o);
        memcpy(out + (size_t)b * 16, o, 16);
    }
}
'''


def _blocks():
    rng = np.random.default_rng(11)
    a = rng.integers(0, 256, (3000, 11, 12), dtype=np.uint8)
    ext = rng.choice(np.array([0, 255], dtype=np.uint8), (3000, 11, 12))           # saturating patterns: both clip directions
    alt = np.zeros((64, 11, 12), dtype=np.uint8)
    alt[:, ::2, :] = 255                                                              # row stripes (vertical ringing)
    alt2 = np.zeros((64, 11, 12), dtype=np.uint8)
    alt2[:, :, ::2] = 255                                                             # column stripes
    flat = np.stack([np.full((11, 12), v, dtype=np.uint8) for v in (0, 1, 127, 128, 254, 255)])
    return np.ascontiguousarray(np.concatenate([a, ext, alt, alt2, flat]))


def test_packed_subpel_rows_and_cells_equal_oracle():
    with tempfile.TemporaryDirectory() as tmp:
        src, so = os.path.join(tmp, "h.cpp"), os.path.join(tmp, "h.so")
        open(src, "w").write(HARNESS)
        subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-I", CSRC, "-o", so, src], check=True)
        L = ctypes.CDLL(so)
        O = oracle.orc()
        blocks = _blocks()
        n = len(blocks)
        for frac in (1, 2, 3):
            taps = np.ascontiguousarray(FILTER[frac])
            for kind, fn, ix, iy in ((0, L.sp_hpp_cells, frac, 0), (2, L.sp_vpp_cells, frac, 0)):      # single-pass kinds take coeffIdx first
                got = np.empty((n, 4, 4), dtype=np.uint8)
                fn(vp(blocks), n, vp(taps), vp(got))
                want = np.empty((n, 4, 4), dtype=np.uint8)
                for b in range(n):
                    O.orc_interp(8, kind, 8, 4, 4, ctypes.c_void_p(blocks[b].ctypes.data + 3 * 12 + 3), ctypes.c_ssize_t(12),
                                 vp(want[b]), ctypes.c_ssize_t(4), ix, iy, 0)
                assert np.array_equal(got, want), (frac, kind, int(np.count_nonzero(got != want)))

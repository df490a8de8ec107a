// tables.cuh -- HEVC constant tables used by the kernels, generated on the host at library load
// (no numbers copied from the reference; tests pin them against source/common/constants.cpp:250-344
// through oracle/_ref).
#pragma once
#include <stdint.h>

namespace x265b200 {

// |64*sqrt(2)*cos(j*pi/64)| as standardised by HEVC (j = 0..32); entry 0 is the DC basis (64).
static const int kDctMag[33] = { 64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67, 64,
                                 61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4, 0 };

// T_N[k][n] of the N-point HEVC core transform (N = 4, 8, 16, 32)  == g_t4 / g_t8 / g_t16 / g_t32
inline int dct_coef(int N, int k, int n)
{
    int m = ((2 * n + 1) * k * (32 / N)) % 128;   // angle in units of pi/64
    if (m > 64) m = 128 - m;
    return m > 32 ? -kDctMag[64 - m] : kDctMag[m];
}

// 8-tap luma / 4-tap chroma interpolation filters (== g_lumaFilter / g_chromaFilter)
static const int16_t kLumaFilter[4][8] = {
    { 0, 0, 0, 64, 0, 0, 0, 0 }, { -1, 4, -10, 58, 17, -5, 1, 0 },
    { -1, 4, -11, 40, 40, -11, 4, -1 }, { 0, 1, -5, 17, 58, -10, 4, -1 } };
static const int16_t kChromaFilter[8][4] = {
    { 0, 64, 0, 0 }, { -2, 58, 10, -2 }, { -4, 54, 16, -2 }, { -6, 46, 28, -4 },
    { -4, 36, 36, -4 }, { -4, 28, 46, -6 }, { -2, 16, 54, -4 }, { -2, 10, 58, -2 } };

} // namespace x265b200

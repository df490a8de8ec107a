// satd_packed.cuh -- 4x4 SATD (pixel.cpp:210-232, satd_4x4) of 8-bit rows that stay packed four pixels per word.
//
// The reference packs two 16-bit partial sums into one integer (sum2_t, pixel.cpp:199-208) and lets carries ride between
// the halves: a word holds lo + (hi << 16) as ONE two's-complement number, which every add/subtract preserves as long
// as |lo| < 2^15.  The same representation is used here, but on the pixel words themselves:
//   row i:  (d0,d2) = prmt(f) - prmt(o) on the even bytes, (d1,d3) on the odd bytes            (4 PRMT + 2 subtracts)
//           S_i = (d0+d1, d2+d3),  D_i = (d0-d1, d2-d3)                                           (first horizontal stage)
//   columns: a 4-point Hadamard down S_0..S_3 and down D_0..D_3, lane-wise                       (16 adds)
//   the second horizontal stage is never formed: |p + q| + |p - q| = 2 max(|p|, |q|), so a word (p, q) contributes
//   max(|p|, |q|) to satd = (sum of |coefficients|) >> 1 -- exactly (the 16 coefficients share one parity).
// Ranges: |d| <= 255, |S|,|D| lanes <= 510, after the column transform <= 2040.
//
// The header compiles as plain C++ too (tests/test_satd_packed_cpu.py runs it on the host against the oracle); the PRMT
// emulation below follows the PTX prmt default mode including the sign-replicate bit of a selector nibble.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SP_FN __host__ __device__ __forceinline__
#else
#define SP_FN static inline
#endif

SP_FN uint32_t sp_prmt(uint32_t a, uint32_t b, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    const uint64_t src = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++)
    {
        const uint32_t n = (sel >> (4 * i)) & 0xf;
        uint32_t byte = (uint32_t)(src >> (8 * (n & 7))) & 0xff;
        if (n & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= byte << (8 * i);
    }
    return r;
#endif
}

SP_FN void sp_hadamard4(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d)
{
    const uint32_t t0 = a + b, t1 = a - b, t2 = c + d, t3 = c - d;
    a = t0 + t2; c = t0 - t2; b = t1 + t3; d = t1 - t3;
}

// max(|lo|, |hi|) of a word holding lo + (hi << 16)
SP_FN int sp_maxabs2(uint32_t w)
{
    const int lo = (int)(int16_t)(w & 0xffff);
    const int hi = (int)(int32_t)(w + 0x8000u) >> 16;
    const int alo = lo < 0 ? -lo : lo, ahi = hi < 0 ? -hi : hi;
    return alo > ahi ? alo : ahi;
}

// f[i], o[i]: row i of the source / predicted 4x4 cell, pixel k in byte k
SP_FN int satd4x4_packed_u8(const uint32_t f[4], const uint32_t o[4])
{
    uint32_t S[4], D[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 4; i++)
    {
        const uint32_t d02 = sp_prmt(f[i], 0, 0x4240) - sp_prmt(o[i], 0, 0x4240);
        const uint32_t d13 = sp_prmt(f[i], 0, 0x4341) - sp_prmt(o[i], 0, 0x4341);
        S[i] = d02 + d13; D[i] = d02 - d13;
    }
    sp_hadamard4(S[0], S[1], S[2], S[3]);
    sp_hadamard4(D[0], D[1], D[2], D[3]);
    int t = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 4; i++) t += sp_maxabs2(S[i]) + sp_maxabs2(D[i]);
    return t;
}

// subpel_packed.cuh -- 8-bit luma sub-pel arithmetic of the frame search (me_frame_kernels.cu, ME_SUBPEL_PACKED) on
// packed words, written so that the same source also compiles for the host (tests/test_subpel_packed_cpu.py checks it
// against the oracle's luma_hpp / luma_vpp, ipfilter.cpp:79-118 / 164-203).
//
//  * tail: out = clip((sum + 32) >> 6) of an 8-tap sum of 8-bit pixels.  sum lies in [-24*255, 88*255], so the
//    reference's int16 cast is the identity, the +32 rides in the accumulator of the first dot product, and the clip is
//    one max(min(v, 255), 0) (VIMNMX.RELU).
//  * hpp_row4_u8: 4 horizontally adjacent outputs from the 12 pixels under their taps (3 words).
//  * vpp_cell_u8: a whole 4x4 cell from its 11 source rows: the three 4x4 byte blocks are transposed once (24 PRMT), an
//    output row is a byte-shifted window of each column (funnel shifts), every pixel is two u8 x s8 dot products.
#pragma once
#include "satd_packed.cuh"

SP_FN int sp_dp4a_us(uint32_t a, uint32_t b, int c)         // sum_i (u8)a_i * (s8)b_i + c
{
#if defined(__CUDA_ARCH__)
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#else
    for (int i = 0; i < 4; i++) c += (int)((a >> (8 * i)) & 0xff) * (int)(int8_t)((b >> (8 * i)) & 0xff);
    return c;
#endif
}
SP_FN uint32_t sp_funnel_r(uint32_t lo, uint32_t hi, uint32_t sh)   // sh in 0, 8, 16, 24
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}
SP_FN int sp_min_relu(int v, int hi)                        // max(min(v, hi), 0)
{
#if defined(__CUDA_ARCH__)
    return __vimin_s32_relu(v, hi);
#else
    v = v < hi ? v : hi;
    return v < 0 ? 0 : v;
#endif
}
SP_FN uint32_t sp_pack4(int a, int b, int c, int d)         // four values already in 0..255
{
    return sp_prmt(sp_prmt((uint32_t)a, (uint32_t)b, 0x0040), sp_prmt((uint32_t)c, (uint32_t)d, 0x0040), 0x5410);
}
SP_FN uint32_t sp_taps(const int16_t* c, int o)             // taps o..o+3 as signed bytes
{
    return (uint32_t)(c[o] & 0xff) | ((uint32_t)(c[o + 1] & 0xff) << 8) | ((uint32_t)(c[o + 2] & 0xff) << 16) | ((uint32_t)(c[o + 3] & 0xff) << 24);
}

// w[0..2]: the 12 pixels starting 3 left of the first output; clo / chi: taps 0..3 / 4..7
SP_FN uint32_t hpp_row4_u8(const uint32_t w[3], uint32_t clo, uint32_t chi)
{
    int o[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 4; k++)
    {
        const uint32_t lo = sp_funnel_r(w[0], w[1], 8 * k), hi = sp_funnel_r(w[1], w[2], 8 * k);
        o[k] = sp_min_relu(sp_dp4a_us(hi, chi, sp_dp4a_us(lo, clo, 32)) >> 6, 255);
    }
    return sp_pack4(o[0], o[1], o[2], o[3]);
}

// r[j], j = 0..10: the 4 pixels of source row (y0 - 3 + j) of the cell's columns; out[i]: predicted row y0 + i
SP_FN void vpp_cell_u8(const uint32_t r[11], uint32_t cvlo, uint32_t cvhi, uint32_t out[4])
{
    uint32_t c[3][4];                                      // c[b][k]: rows 4b..4b+3 of column k, one byte each
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int b = 0; b < 3; b++)
    {
        const uint32_t r3 = b == 2 ? 0u : r[4 * b + 3];    // row 11 does not exist (and is never under a tap)
        const uint32_t t0 = sp_prmt(r[4 * b], r[4 * b + 1], 0x5140), t1 = sp_prmt(r[4 * b + 2], r3, 0x5140);
        const uint32_t t2 = sp_prmt(r[4 * b], r[4 * b + 1], 0x7362), t3 = sp_prmt(r[4 * b + 2], r3, 0x7362);
        c[b][0] = sp_prmt(t0, t1, 0x5410); c[b][1] = sp_prmt(t0, t1, 0x7632);
        c[b][2] = sp_prmt(t2, t3, 0x5410); c[b][3] = sp_prmt(t2, t3, 0x7632);
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 4; i++)
    {
        int o[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < 4; k++)
        {
            const uint32_t lo = sp_funnel_r(c[0][k], c[1][k], 8 * i), hi = sp_funnel_r(c[1][k], c[2][k], 8 * i);
            o[k] = sp_min_relu(sp_dp4a_us(hi, cvhi, sp_dp4a_us(lo, cvlo, 32)) >> 6, 255);
        }
        out[i] = sp_pack4(o[0], o[1], o[2], o[3]);
    }
}

// ---- two-pass case (luma_hvpp = hps with row extension + vsp, ipfilter.cpp:362-369), 8-bit ------------------------------
// w[j][0..2], j = 0..10: the 12 pixels of source row (y0 - 3 + j) starting 3 left of the cell; clo / chi: horizontal taps;
// cv: vertical taps.  First stage at 8 bits: shift 0, offset -8192 (rides in the accumulator); it lies in [-14312, 14248].
// Second stage: (sum + 2048 + (8192 << 6)) >> 12 lies in [-262, 436], so the int16 cast is the identity here too.
SP_FN void hvpp_cell_u8(const uint32_t w[11][3], uint32_t clo, uint32_t chi, const int16_t cv[8], uint32_t out[4])
{
    int v[11][4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < 11; j++)
    {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < 4; k++)
        {
            const uint32_t lo = sp_funnel_r(w[j][0], w[j][1], 8 * k), hi = sp_funnel_r(w[j][1], w[j][2], 8 * k);
            v[j][k] = sp_dp4a_us(hi, chi, sp_dp4a_us(lo, clo, -8192));
        }
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 4; i++)
    {
        int o[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < 4; k++)
        {
            int sum = 2048 + (8192 << 6);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int t = 0; t < 8; t++) sum += v[i + t][k] * (int)cv[t];
            o[k] = sp_min_relu(sum >> 12, 255);
        }
        out[i] = sp_pack4(o[0], o[1], o[2], o[3]);
    }
}

// ---- the vertical cell in two halves, for loops that carry the transposed blocks from one cell to the next ------------------
// c[k] = rows r0..r3 of column k, one byte each
SP_FN void sp_transpose4(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3, uint32_t c[4])
{
    const uint32_t t0 = sp_prmt(r0, r1, 0x5140), t1 = sp_prmt(r2, r3, 0x5140);
    const uint32_t t2 = sp_prmt(r0, r1, 0x7362), t3 = sp_prmt(r2, r3, 0x7362);
    c[0] = sp_prmt(t0, t1, 0x5410); c[1] = sp_prmt(t0, t1, 0x7632);
    c[2] = sp_prmt(t2, t3, 0x5410); c[3] = sp_prmt(t2, t3, 0x7632);
}
// c0 / c1 / c2: source rows y0-3..y0, y0+1..y0+4, y0+5..y0+8 of the cell's four columns (sp_transpose4); out[i]: predicted row y0+i
SP_FN void vpp_cell_from_cols_u8(const uint32_t c0[4], const uint32_t c1[4], const uint32_t c2[4], uint32_t cvlo, uint32_t cvhi, uint32_t out[4])
{
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 4; i++)
    {
        int o[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < 4; k++)
        {
            const uint32_t lo = sp_funnel_r(c0[k], c1[k], 8 * i), hi = sp_funnel_r(c1[k], c2[k], 8 * i);
            o[k] = sp_min_relu(sp_dp4a_us(hi, cvhi, sp_dp4a_us(lo, cvlo, 32)) >> 6, 255);
        }
        out[i] = sp_pack4(o[0], o[1], o[2], o[3]);
    }
}

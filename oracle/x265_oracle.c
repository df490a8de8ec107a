/* oracle/x265_oracle.c -- TEST INFRASTRUCTURE ONLY (see x265_oracle.h).
 * CPU restatement of the reference's C primitives; each function cites the reference
 * file:line (relative to /root/reference/source) whose arithmetic it follows. */
#include "x265_oracle.h"
#include <stdlib.h>
#include <string.h>

static inline int px(const void* p, int depth, intptr_t i)
{
    return depth == 8 ? ((const uint8_t*)p)[i] : ((const uint16_t*)p)[i];
}
static inline const void* padd(const void* p, int depth, intptr_t i)
{
    return depth == 8 ? (const void*)((const uint8_t*)p + i) : (const void*)((const uint16_t*)p + i);
}

/* common/pixel.cpp:40-55  sad<lx,ly> */
int orc_sad(int depth, int w, int h, const void* a, intptr_t sa, const void* b, intptr_t sb)
{
    int sum = 0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            sum += abs(px(a, depth, y * sa + x) - px(b, depth, y * sb + x));
    return sum;
}

/* common/pixel.cpp:74-119  sad_x3 / sad_x4: fenc stride is FENC_STRIDE (64) */
void orc_sad_xn(int depth, int K, int w, int h, const void* fenc, const void* const* refs, intptr_t refStride, int32_t* res)
{
    for (int k = 0; k < K; k++)
        res[k] = orc_sad(depth, w, h, fenc, 64, refs[k], refStride);
}

/* common/pixel.cpp:190-199 HADAMARD4 */
static void hadamard4(int* d0, int* d1, int* d2, int* d3, int s0, int s1, int s2, int s3)
{
    int t0 = s0 + s1, t1 = s0 - s1, t2 = s2 + s3, t3 = s2 - s3;
    *d0 = t0 + t2; *d2 = t0 - t2; *d1 = t1 + t3; *d3 = t1 - t3;
}

/* common/pixel.cpp:210-236 satd_4x4 (SWAR lanes written out as separate ints), returns sum>>1 */
static int satd_4x4(int depth, const void* a, intptr_t sa, const void* b, intptr_t sb)
{
    int tmp[4][4], sum = 0;
    for (int i = 0; i < 4; i++)
    {
        int d0 = px(a, depth, i * sa + 0) - px(b, depth, i * sb + 0);
        int d1 = px(a, depth, i * sa + 1) - px(b, depth, i * sb + 1);
        int d2 = px(a, depth, i * sa + 2) - px(b, depth, i * sb + 2);
        int d3 = px(a, depth, i * sa + 3) - px(b, depth, i * sb + 3);
        hadamard4(&tmp[i][0], &tmp[i][1], &tmp[i][2], &tmp[i][3], d0, d1, d2, d3);
    }
    for (int i = 0; i < 4; i++)
    {
        int a0, a1, a2, a3;
        hadamard4(&a0, &a1, &a2, &a3, tmp[0][i], tmp[1][i], tmp[2][i], tmp[3][i]);
        sum += abs(a0) + abs(a1) + abs(a2) + abs(a3);
    }
    return sum >> 1;
}

/* common/pixel.cpp:239-261 satd_8x4: two 4x4 Hadamards, ONE shift over their sum */
static int satd_8x4(int depth, const void* a, intptr_t sa, const void* b, intptr_t sb)
{
    int sum = 0;
    for (int half = 0; half < 2; half++)
    {
        int tmp[4][4];
        for (int i = 0; i < 4; i++)
        {
            int d[4];
            for (int k = 0; k < 4; k++)
                d[k] = px(a, depth, i * sa + 4 * half + k) - px(b, depth, i * sb + 4 * half + k);
            hadamard4(&tmp[i][0], &tmp[i][1], &tmp[i][2], &tmp[i][3], d[0], d[1], d[2], d[3]);
        }
        for (int i = 0; i < 4; i++)
        {
            int a0, a1, a2, a3;
            hadamard4(&a0, &a1, &a2, &a3, tmp[0][i], tmp[1][i], tmp[2][i], tmp[3][i]);
            sum += abs(a0) + abs(a1) + abs(a2) + abs(a3);
        }
    }
    return sum >> 1;
}

/* table common/pixel.cpp:1131-1155: satd8<w,h> when w % 8 == 0 (except 8x4 = satd_8x4, same
 * thing), satd4<w,h> otherwise (4x4, 4x8, 12x16, 4x16 ...) */
int orc_satd(int depth, int w, int h, const void* a, intptr_t sa, const void* b, intptr_t sb)
{
    int satd = 0;
    if ((w & 7) == 0)
    {
        for (int row = 0; row < h; row += 4)          /* pixel.cpp:281-297 satd8 */
            for (int col = 0; col < w; col += 8)
                satd += satd_8x4(depth, padd(a, depth, row * sa + col), sa, padd(b, depth, row * sb + col), sb);
    }
    else
    {
        for (int row = 0; row < h; row += 4)          /* pixel.cpp:263-279 satd4 */
            for (int col = 0; col < w; col += 4)
                satd += satd_4x4(depth, padd(a, depth, row * sa + col), sa, padd(b, depth, row * sb + col), sb);
    }
    return satd;
}

/* common/pixel.cpp:299-334 _sa8d_8x8 (unrounded) */
static int sa8d_8x8_raw(int depth, const void* a, intptr_t sa, const void* b, intptr_t sb)
{
    int m[8][8], sum = 0;
    for (int i = 0; i < 8; i++)
    {
        int d[8], e[8];
        for (int k = 0; k < 8; k++) d[k] = px(a, depth, i * sa + k) - px(b, depth, i * sb + k);
        /* b0..b3 hold (sum, diff) pairs of adjacent pixels; then HADAMARD4 across the pairs */
        int s0 = d[0] + d[1], f0 = d[0] - d[1], s1 = d[2] + d[3], f1 = d[2] - d[3];
        int s2 = d[4] + d[5], f2 = d[4] - d[5], s3 = d[6] + d[7], f3 = d[6] - d[7];
        hadamard4(&e[0], &e[1], &e[2], &e[3], s0, s1, s2, s3);
        hadamard4(&e[4], &e[5], &e[6], &e[7], f0, f1, f2, f3);
        for (int k = 0; k < 8; k++) m[i][k] = e[k];
    }
    for (int i = 0; i < 8; i++)
    {
        int a0, a1, a2, a3, a4, a5, a6, a7;
        hadamard4(&a0, &a1, &a2, &a3, m[0][i], m[1][i], m[2][i], m[3][i]);
        hadamard4(&a4, &a5, &a6, &a7, m[4][i], m[5][i], m[6][i], m[7][i]);
        sum += abs(a0 + a4) + abs(a0 - a4) + abs(a1 + a5) + abs(a1 - a5)
             + abs(a2 + a6) + abs(a2 - a6) + abs(a3 + a7) + abs(a3 - a7);
    }
    return sum;
}

/* common/pixel.cpp:336-377: sa8d_8x8 rounds (s+2)>>2 per 8x8; sa8d_16x16 once per 16x16 */
int orc_sa8d(int depth, int w, int h, int per16, const void* a, intptr_t sa, const void* b, intptr_t sb)
{
    int cost = 0;
    if (w == 4 && h == 4) return satd_4x4(depth, a, sa, b, sb);     /* pixel.cpp:1163 */
    if (per16)
    {
        for (int y = 0; y < h; y += 16)
            for (int x = 0; x < w; x += 16)
            {
                int s = 0;
                for (int q = 0; q < 4; q++)
                {
                    int ox = x + (q & 1) * 8, oy = y + (q >> 1) * 8;
                    s += sa8d_8x8_raw(depth, padd(a, depth, oy * sa + ox), sa, padd(b, depth, oy * sb + ox), sb);
                }
                cost += (s + 2) >> 2;
            }
    }
    else
    {
        for (int y = 0; y < h; y += 8)
            for (int x = 0; x < w; x += 8)
                cost += (sa8d_8x8_raw(depth, padd(a, depth, y * sa + x), sa, padd(b, depth, y * sb + x), sb) + 2) >> 2;
    }
    return cost;
}

/* common/pixel.cpp:167-186 sse<>: sse_t is uint32 below 10-bit, uint64 otherwise (common.h:144-148) */
uint64_t orc_sse_pp(int depth, int w, int h, const void* a, intptr_t sa, const void* b, intptr_t sb)
{
    uint64_t sum = 0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
        {
            int tmp = px(a, depth, y * sa + x) - px(b, depth, y * sb + x);
            sum += (uint64_t)(int64_t)(tmp * tmp);
        }
    return depth < 10 ? (sum & 0xffffffffull) : sum;
}

uint64_t orc_sse_ss(int depth, int w, int h, const int16_t* a, intptr_t sa, const int16_t* b, intptr_t sb)
{
    uint64_t sum = 0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
        {
            int tmp = a[y * sa + x] - b[y * sb + x];
            int sq = (int)((uint32_t)tmp * (uint32_t)tmp);       /* int product, wraps like the reference build */
            sum += (uint64_t)(int64_t)sq;
        }
    return depth < 10 ? (sum & 0xffffffffull) : sum;
}

/* common/pixel.cpp:379-391 pixel_ssd_s_c */
uint64_t orc_ssd_s(int depth, int size, const int16_t* a, intptr_t sa)
{
    uint64_t sum = 0;
    for (int y = 0; y < size; y++)
        for (int x = 0; x < size; x++)
            sum += (uint64_t)(int64_t)(a[y * sa + x] * a[y * sa + x]);
    return depth < 10 ? (sum & 0xffffffffull) : sum;
}

/* =====================================================================================================
 * Transforms.  common/dct.cpp:83-240,418-440 (partialButterfly*) compute, for each input row j,
 *   dst[k*line + j] = (sum_n g_tN[k][n] * src[j*N + n] + add) >> shift
 * through an even/odd factorisation; the factorisation is exact, so the dense sum below is the same
 * integer.  Forward results are truncated to int16 (dct.cpp:113), inverse results saturate (dct.cpp:257).
 * The matrices g_t4..g_t32 (common/constants.cpp:270-344) are regenerated from the 33 HEVC cosine
 * magnitudes; tests/test_tables_cpu.py pins them against the reference's tables.
 * ===================================================================================================== */
static const int dct_mag[33] = { 64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67, 64,
                                 61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4, 0 };
static int dct_coef(int N, int k, int n)
{
    int m = ((2 * n + 1) * k * (32 / N)) % 128;
    if (m > 64) m = 128 - m;
    return m > 32 ? -dct_mag[64 - m] : dct_mag[m];
}
static const int dst4_mat[4][4] = { { 29, 55, 74, 84 }, { 74, 74, 0, -74 }, { 84, -29, -74, 55 }, { 55, -84, 74, -29 } };   /* dct.cpp:43-61 expanded */

static int clip16(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }

/* one forward pass: out[k*N + j] = (int16)((sum_n M[k][n] * in[j*N + n] + add) >> shift) */
static void fwd_pass(int N, int isDst, const int16_t* in, int16_t* out, int shift)
{
    int add = 1 << (shift - 1);
    for (int j = 0; j < N; j++)
        for (int k = 0; k < N; k++)
        {
            int sum = 0;
            for (int n = 0; n < N; n++)
                sum += (isDst ? dst4_mat[k][n] : dct_coef(N, k, n)) * in[j * N + n];
            out[k * N + j] = (int16_t)((sum + add) >> shift);
        }
}
/* one inverse pass (dct.cpp:242-416, :63-81): out[j*N + k] = clip16((sum_n M[n][k] * in[n*N + j] + add) >> shift) */
static void inv_pass(int N, int isDst, const int16_t* in, int16_t* out, int shift)
{
    int add = 1 << (shift - 1);
    for (int j = 0; j < N; j++)
        for (int k = 0; k < N; k++)
        {
            int sum = 0;
            for (int n = 0; n < N; n++)
                sum += (isDst ? dst4_mat[n][k] : dct_coef(N, n, k)) * in[n * N + j];
            out[j * N + k] = (int16_t)clip16((sum + add) >> shift);
        }
}

void orc_dct(int depth, int idx, const int16_t* src, int16_t* dst, intptr_t srcStride)
{
    int N = idx == 4 ? 4 : (4 << idx), log2N = idx == 4 ? 2 : idx + 2;
    int16_t block[32 * 32], coef[32 * 32];
    for (int i = 0; i < N; i++) memcpy(block + i * N, src + i * srcStride, N * sizeof(int16_t));
    fwd_pass(N, idx == 4, block, coef, log2N - 1 + (depth - 8));      /* dct.cpp:444,461,478,495,512 shift_1st */
    fwd_pass(N, idx == 4, coef, dst, log2N + 6);                      /* shift_2nd */
}

void orc_idct(int depth, int idx, const int16_t* src, int16_t* dst, intptr_t dstStride)
{
    int N = idx == 4 ? 4 : (4 << idx);
    int16_t coef[32 * 32], block[32 * 32];
    inv_pass(N, idx == 4, src, coef, 7);                              /* dct.cpp:529-530 */
    inv_pass(N, idx == 4, coef, block, 12 - (depth - 8));
    for (int i = 0; i < N; i++) memcpy(dst + i * dstStride, block + i * N, N * sizeof(int16_t));
}

/* dct.cpp:664-686 */
uint32_t orc_quant(const int16_t* coef, const int32_t* quantCoeff, int32_t* deltaU, int16_t* qCoef, int qBits, int add, int numCoeff)
{
    uint32_t numSig = 0;
    for (int i = 0; i < numCoeff; i++)
    {
        int level = coef[i], sign = level < 0 ? -1 : 1;
        int tmplevel = (int)((uint32_t)abs(level) * (uint32_t)quantCoeff[i]);
        level = (tmplevel + add) >> qBits;
        deltaU[i] = (tmplevel - (level << qBits)) >> (qBits - 8);
        if (level) numSig++;
        level *= sign;
        qCoef[i] = (int16_t)clip16(level);
    }
    return numSig;
}
/* dct.cpp:688-713 */
uint32_t orc_nquant(const int16_t* coef, const int32_t* quantCoeff, int16_t* qCoef, int qBits, int add, int numCoeff)
{
    uint32_t numSig = 0;
    for (int i = 0; i < numCoeff; i++)
    {
        int level = coef[i], sign = level < 0 ? -1 : 1;
        int tmplevel = (int)((uint32_t)abs(level) * (uint32_t)quantCoeff[i]);
        level = (tmplevel + add) >> qBits;
        if (level) numSig++;
        level *= sign;
        qCoef[i] = (int16_t)abs(clip16(level));
    }
    return numSig;
}
/* dct.cpp:612-634 */
void orc_dequant_normal(const int16_t* q, int16_t* coef, int num, int scale, int shift)
{
    int add = 1 << (shift - 1);
    for (int n = 0; n < num; n++) coef[n] = (int16_t)clip16((q[n] * scale + add) >> shift);
}
/* dct.cpp:636-662 */
void orc_dequant_scaling(const int16_t* q, const int32_t* deq, int16_t* coef, int num, int per, int shift)
{
    shift += 4;
    if (shift > per)
    {
        int add = 1 << (shift - per - 1);
        for (int n = 0; n < num; n++) coef[n] = (int16_t)clip16((q[n] * deq[n] + add) >> (shift - per));
    }
    else
        for (int n = 0; n < num; n++) coef[n] = (int16_t)clip16((int)((uint32_t)clip16(q[n] * deq[n]) << (per - shift)));
}

/* =====================================================================================================
 * Interpolation (common/ipfilter.cpp:40-370).  Taps: common/constants.cpp:250-268 (HEVC standard values).
 * ===================================================================================================== */
static const int luma_filter[4][8] = { { 0, 0, 0, 64, 0, 0, 0, 0 }, { -1, 4, -10, 58, 17, -5, 1, 0 },
                                       { -1, 4, -11, 40, 40, -11, 4, -1 }, { 0, 1, -5, 17, 58, -10, 4, -1 } };
static const int chroma_filter[8][4] = { { 0, 64, 0, 0 }, { -2, 58, 10, -2 }, { -4, 54, 16, -2 }, { -6, 46, 28, -4 },
                                         { -4, 36, 36, -4 }, { -4, 28, 46, -6 }, { -2, 16, 54, -4 }, { -2, 10, 58, -2 } };
static int tap(int taps, int idx, int t) { return taps == 8 ? luma_filter[idx][t] : chroma_filter[idx][t]; }
static void put_px(void* p, int depth, intptr_t i, int v) { if (depth == 8) ((uint8_t*)p)[i] = (uint8_t)v; else ((uint16_t*)p)[i] = (uint16_t)v; }

/* generic N-tap pass.  srcShort/dstShort select int16 operands; `step` = 1 (horizontal) or srcStride (vertical) */
static void fir_pass(int depth, int taps, int idx, int w, int rows, const void* src, int srcShort, intptr_t srcStride, intptr_t step,
                     void* dst, int dstShort, intptr_t dstStride, int shift, int offset)
{
    int maxVal = (1 << depth) - 1;
    for (int y = 0; y < rows; y++)
        for (int x = 0; x < w; x++)
        {
            int sum = 0;
            for (int t = 0; t < taps; t++)
            {
                intptr_t o = y * srcStride + x + t * step;
                sum += (srcShort ? ((const int16_t*)src)[o] : px(src, depth, o)) * tap(taps, idx, t);
            }
            int val = (int16_t)((sum + offset) >> shift);
            if (dstShort) ((int16_t*)dst)[y * dstStride + x] = (int16_t)val;
            else put_px(dst, depth, y * dstStride + x, val < 0 ? 0 : (val > maxVal ? maxVal : val));
        }
}

void orc_interp(int depth, int kind, int taps, int w, int h, const void* src, intptr_t srcStride, void* dst, intptr_t dstStride,
                int idxX, int idxY, int isRowExt)
{
    int headRoom = 14 - depth, half = taps / 2 - 1;
    const void* sh = padd(src, depth, -half);                               /* src - (N/2-1)            */
    const void* sv = padd(src, depth, -half * srcStride);                   /* src - (N/2-1)*srcStride  */
    switch (kind)
    {
    case 0: fir_pass(depth, taps, idxX, w, h, sh, 0, srcStride, 1, dst, 0, dstStride, 6, 32); break;                               /* :79-118  */
    case 1:                                                                                                                          /* :120-162 */
    {
        int shift = 6 - headRoom, offset = (int)((unsigned)-8192 << shift), rows = h;
        if (isRowExt) { sh = padd(sh, depth, -half * srcStride); rows += taps - 1; }
        fir_pass(depth, taps, idxX, w, rows, sh, 0, srcStride, 1, dst, 1, dstStride, shift, offset);
        break;
    }
    case 2: fir_pass(depth, taps, idxX, w, h, sv, 0, srcStride, srcStride, dst, 0, dstStride, 6, 32); break;                       /* :164-203 */
    case 3: { int shift = 6 - headRoom; fir_pass(depth, taps, idxX, w, h, sv, 0, srcStride, srcStride, dst, 1, dstStride, shift, (int)((unsigned)-8192 << shift)); break; } /* :205-239 */
    case 4: { int shift = 6 + headRoom; fir_pass(depth, taps, idxX, w, h, (const int16_t*)src - half * srcStride, 1, srcStride, srcStride, dst, 0, dstStride, shift, (1 << (shift - 1)) + (8192 << 6)); break; } /* :241-282 */
    case 5: fir_pass(depth, taps, idxX, w, h, (const int16_t*)src - half * srcStride, 1, srcStride, srcStride, dst, 1, dstStride, 6, 0); break;  /* :284-317 */
    case 6:                                                                                                                          /* :362-369 */
    {
        int16_t immed[64 * (64 + 7)];
        int shift = 6 - headRoom, shift2 = 6 + headRoom;
        fir_pass(depth, taps, idxX, w, h + taps - 1, padd(sh, depth, -half * srcStride), 0, srcStride, 1, immed, 1, w, shift, (int)((unsigned)-8192 << shift));
        fir_pass(depth, taps, idxY, w, h, immed, 1, w, w, dst, 0, dstStride, shift2, (1 << (shift2 - 1)) + (8192 << 6));
        break;
    }
    case 7:                                                                                                                          /* :40-57 */
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++)
            {
                int16_t val = (int16_t)(px(src, depth, y * srcStride + x) << headRoom);
                ((int16_t*)dst)[y * dstStride + x] = (int16_t)(val - (int16_t)8192);
            }
        break;
    }
}

/* =====================================================================================================
 * Intra prediction (common/intrapred.cpp:31-204)
 * ===================================================================================================== */
void orc_intra_filter(int depth, int log2N, const void* s, void* f)         /* :31-51 */
{
    int N = 1 << log2N, N2 = 2 * N;
    put_px(f, depth, 0, ((px(s, depth, 0) << 1) + px(s, depth, 1) + px(s, depth, N2 + 1) + 2) >> 2);
    for (int i = 1; i < N2; i++) put_px(f, depth, i, ((px(s, depth, i) << 1) + px(s, depth, i - 1) + px(s, depth, i + 1) + 2) >> 2);
    put_px(f, depth, N2, px(s, depth, N2));
    put_px(f, depth, N2 + 1, ((px(s, depth, N2 + 1) << 1) + px(s, depth, 0) + px(s, depth, N2 + 2) + 2) >> 2);
    for (int i = N2 + 2; i < 2 * N2; i++) put_px(f, depth, i, ((px(s, depth, i) << 1) + px(s, depth, i - 1) + px(s, depth, i + 1) + 2) >> 2);
    put_px(f, depth, 2 * N2, px(s, depth, 2 * N2));
}

void orc_intra_pred(int depth, int log2N, int mode, int bFilter, const void* srcPix, void* dst, intptr_t ds)
{
    static const int angleTable[17] = { -32, -26, -21, -17, -13, -9, -5, -2, 0, 2, 5, 9, 13, 17, 21, 26, 32 };
    static const int invAngleTable[8] = { 4096, 1638, 910, 630, 482, 390, 315, 256 };
    int N = 1 << log2N, N2 = 2 * N, maxVal = (1 << depth) - 1;
    int s[129];
    for (int i = 0; i < 4 * N + 1; i++) s[i] = px(srcPix, depth, i);
    if (mode == 0)                                                            /* planar :87-100 */
    {
        for (int y = 0; y < N; y++)
            for (int x = 0; x < N; x++)
                put_px(dst, depth, y * ds + x, ((N - 1 - x) * s[N2 + 1 + y] + (N - 1 - y) * s[1 + x] + (x + 1) * s[1 + N] + (y + 1) * s[N2 + 1 + N] + N) >> (log2N + 1));
        return;
    }
    if (mode == 1)                                                            /* DC :53-85 */
    {
        int dc = N;
        for (int i = 0; i < N; i++) dc += s[1 + i] + s[N2 + 1 + i];
        dc /= 2 * N;
        for (int y = 0; y < N; y++) for (int x = 0; x < N; x++) put_px(dst, depth, y * ds + x, dc);
        if (bFilter)
        {
            put_px(dst, depth, 0, (s[1] + s[N2 + 1] + 2 * dc + 2) >> 2);
            for (int x = 1; x < N; x++) put_px(dst, depth, x, (s[1 + x] + 3 * dc + 2) >> 2);
            for (int y = 1; y < N; y++) put_px(dst, depth, y * ds, (s[N2 + 1 + y] + 3 * dc + 2) >> 2);
        }
        return;
    }
    /* angular :102-204 */
    int hor = mode < 18, nbuf[129];
    if (hor)
    {
        nbuf[0] = s[0];
        for (int i = 0; i < N2; i++) { nbuf[1 + i] = s[N2 + 1 + i]; nbuf[N2 + 1 + i] = s[1 + i]; }
    }
    else memcpy(nbuf, s, sizeof(int) * (4 * N + 1));
    int angleOffset = hor ? 10 - mode : mode - 26, angle = angleTable[8 + angleOffset];
    int pred[32][32];
    if (!angle)
    {
        for (int y = 0; y < N; y++) for (int x = 0; x < N; x++) pred[y][x] = nbuf[1 + x];
        if (bFilter)
            for (int y = 0; y < N; y++)
            {
                int v = (int16_t)(nbuf[1] + ((nbuf[N2 + 1 + y] - nbuf[0]) >> 1));
                pred[y][0] = v < 0 ? 0 : (v > maxVal ? maxVal : v);
            }
    }
    else
    {
        int refBuf[64 + 1], *ref;
        if (angle < 0)
        {
            int nbProjected = -((N * angle) >> 5) - 1;
            int* ref_pix = refBuf + nbProjected + 1;
            int invAngle = invAngleTable[-angleOffset - 1], invAngleSum = 128;
            for (int i = 0; i < nbProjected; i++) { invAngleSum += invAngle; ref_pix[-2 - i] = nbuf[N2 + (invAngleSum >> 8)]; }
            for (int i = 0; i < N + 1; i++) ref_pix[-1 + i] = nbuf[i];
            ref = ref_pix;
        }
        else ref = nbuf + 1;
        int angleSum = 0;
        for (int y = 0; y < N; y++)
        {
            angleSum += angle;
            int offset = angleSum >> 5, fraction = angleSum & 31;
            for (int x = 0; x < N; x++)
                pred[y][x] = fraction ? ((32 - fraction) * ref[offset + x] + fraction * ref[offset + x + 1] + 16) >> 5 : ref[offset + x];
        }
    }
    for (int y = 0; y < N; y++)
        for (int x = 0; x < N; x++)
            put_px(dst, depth, y * ds + x, hor ? pred[x][y] : pred[y][x]);
}

/* ======================================================================================================================
 * Families added late in round 1: SEA support, in-loop filters, cuTree helpers, the residual pipeline chain.
 * ====================================================================================================================== */

/* common/pixel.cpp:121-165  ads_x1 / ads_x2 / ads_x4<lx,ly>: kind = 1, 2, 4; lxHalf = lx >> 1 */
int orc_ads(int kind, int lxHalf, const int* encDC, const uint32_t* sums, intptr_t delta, const uint16_t* costMvX, int16_t* mvs, int width, int thresh)
{
    int nmv = 0;
    for (int i = 0; i < width; i++, sums++)
    {
        long long ads = llabs((long long)encDC[0] - (long long)sums[0]);
        if (kind == 2) ads += llabs((long long)encDC[1] - (long long)sums[delta]);
        if (kind == 4)
            ads += llabs((long long)encDC[1] - (long long)sums[lxHalf]) + llabs((long long)encDC[2] - (long long)sums[delta]) +
                   llabs((long long)encDC[3] - (long long)sums[delta + lxHalf]);
        ads += costMvX[i];
        if ((int)ads < thresh) mvs[nmv++] = (int16_t)i;
    }
    return nmv;
}

/* encoder/framefilter.cpp:39-140 (integral_init{4..32}h / v) driven as FrameFilter::computeMEIntegral drives them (:722-825):
 * running column sums of horizontal w-sums, turned into w x h box sums h rows later.  planes[k]: origin (pixel 0,0). */
void orc_sea_integral(int depth, const void* reconOrigin, intptr_t stride, int padX, int padY, int maxHeight, uint32_t* const planes[12])
{
    static const int W[12] = { 32, 32, 32, 24, 16, 16, 16, 12, 8, 8, 4, 4 }, H[12] = { 32, 24, 8, 32, 16, 12, 4, 16, 32, 8, 16, 4 };
    for (int k = 0; k < 12; k++)
    {
        uint32_t* base = planes[k] - padY * stride - padX;
        memset(base, 0, stride * sizeof(uint32_t));                                   /* :757 */
        const int height = maxHeight + padY - 1;                                      /* lastRow: height += padY - 1 */
        for (int y = -padY; y < height; y++)
        {
            const void* pix = padd(reconOrigin, depth, y * stride - padX);
            uint32_t* sum = planes[k] + (y + 1) * stride - padX;
            int32_t v = 0;
            for (int i = 0; i < W[k]; i++) v += px(pix, depth, i);
            for (int x = 0; x < stride - W[k]; x++)                                    /* integral_initNh_c */
            {
                sum[x] = (uint32_t)v + sum[x - stride];
                v += px(pix, depth, x + W[k]) - px(pix, depth, x);
            }
            if (y >= H[k] - padY)                                                      /* integral_initNv_c on the row H rows up */
            {
                uint32_t* top = sum - H[k] * stride;
                for (int x = 0; x < stride; x++) top[x] = top[x + H[k] * stride] - top[x];
            }
        }
    }
}

static int sgn_of(int x) { return (x > 0) - (x < 0); }                                /* common/loopfilter.cpp:33-36 signOf */
static int clip_px(int v, int depth) { int m = (1 << depth) - 1; return v < 0 ? 0 : (v > m ? m : v); }
static void put_pix(void* p, int depth, intptr_t i, int v)
{
    if (depth == 8) ((uint8_t*)p)[i] = (uint8_t)v; else ((uint16_t*)p)[i] = (uint16_t)v;
}

/* common/loopfilter.cpp:45-139: kind 0 processSaoCUE0, 1 E1, 2 E1_2Rows, 3 E2, 4 E3, 5 B0 (sequential, in place, as written there) */
void orc_sao_apply(int kind, int depth, void* rec, intptr_t stride, int8_t* buf0, int8_t* buf1, const int8_t* offset, int width, int height, int startX)
{
    if (kind == 0)
        for (int y = 0; y < 2; y++)
        {
            int signLeft0 = buf0[y];
            for (int x = 0; x < width; x++)
            {
                int d = px(rec, depth, y * stride + x) - px(rec, depth, y * stride + x + 1);
                int signRight = sgn_of(d), edgeType = signRight + signLeft0 + 2;
                signLeft0 = -signRight;
                put_pix(rec, depth, y * stride + x, clip_px(px(rec, depth, y * stride + x) + offset[edgeType], depth));
            }
        }
    else if (kind == 1 || kind == 2)
        for (int y = 0; y < (kind == 1 ? 1 : 2); y++)
            for (int x = 0; x < width; x++)
            {
                int signDown = sgn_of(px(rec, depth, y * stride + x) - px(rec, depth, (y + 1) * stride + x));
                int edgeType = signDown + buf0[x] + 2;
                buf0[x] = (int8_t)-signDown;
                put_pix(rec, depth, y * stride + x, clip_px(px(rec, depth, y * stride + x) + offset[edgeType], depth));
            }
    else if (kind == 3)
        for (int x = 0; x < width; x++)
        {
            int signDown = sgn_of(px(rec, depth, x) - px(rec, depth, x + stride + 1));
            int edgeType = signDown + buf1[x] + 2;
            buf0[x + 1] = (int8_t)-signDown;                                          /* bufft */
            put_pix(rec, depth, x, clip_px(px(rec, depth, x) + offset[edgeType], depth));
        }
    else if (kind == 4)
        for (int x = startX + 1; x < width; x++)                                       /* width carries endX */
        {
            int signDown = sgn_of(px(rec, depth, x) - px(rec, depth, x + stride));
            int edgeType = signDown + buf0[x] + 2;
            buf0[x - 1] = (int8_t)-signDown;
            put_pix(rec, depth, x, clip_px(px(rec, depth, x) + offset[edgeType], depth));
        }
    else
        for (int y = 0; y < height; y++)
            for (int x = 0; x < width; x++)
            {
                int v = px(rec, depth, y * stride + x);
                put_pix(rec, depth, y * stride + x, clip_px(v + offset[v >> (depth - 5)], depth));
            }
}

/* encoder/sao.cpp:1762-1926: kind 5 saoCuStatsBO, 0 E0, 1 E1, 3 E2, 4 E3; diff pitch is MAX_CU_SIZE = 64 */
void orc_sao_stats(int kind, int depth, const int16_t* diff, const void* rec, intptr_t stride, int8_t* upBuff1, int8_t* upBufft,
                   int endX, int endY, int32_t* stats, int32_t* count)
{
    static const int eoTable[5] = { 1, 2, 0, 3, 4 };                                   /* SAO::s_eoTable, sao.cpp:65-72 */
    int32_t ts[5] = { 0 }, tc[5] = { 0 };
    for (int y = 0; y < endY; y++)
    {
        const intptr_t r = y * stride;
        if (kind == 5)
        {
            for (int x = 0; x < endX; x++) { int c = px(rec, depth, r + x) >> (depth - 5); stats[c] += diff[y * 64 + x]; count[c]++; }
            continue;
        }
        int signLeft = kind == 0 ? sgn_of(px(rec, depth, r) - px(rec, depth, r - 1)) : 0;
        if (kind == 3) upBufft[0] = (int8_t)sgn_of(px(rec, depth, r + stride) - px(rec, depth, r - 1));
        for (int x = 0; x < endX; x++)
        {
            int c = px(rec, depth, r + x), edgeType;
            if (kind == 0) { int sr = sgn_of(c - px(rec, depth, r + x + 1)); edgeType = sr + signLeft + 2; signLeft = -sr; }
            else if (kind == 1) { int sd = sgn_of(c - px(rec, depth, r + x + stride)); edgeType = sd + upBuff1[x] + 2; upBuff1[x] = (int8_t)-sd; }
            else if (kind == 3) { int sd = sgn_of(c - px(rec, depth, r + x + stride + 1)); edgeType = sd + upBuff1[x] + 2; upBufft[x + 1] = (int8_t)-sd; }
            else { int sd = sgn_of(c - px(rec, depth, r + x + stride - 1)); edgeType = sd + upBuff1[x] + 2; upBuff1[x - 1] = (int8_t)-sd; }
            ts[edgeType] += diff[y * 64 + x]; tc[edgeType]++;
        }
        if (kind == 3) { int8_t* t = upBuff1; upBuff1 = upBufft; upBufft = t; }
        if (kind == 4) upBuff1[endX - 1] = (int8_t)sgn_of(px(rec, depth, r + endX - 1 + stride) - px(rec, depth, r + endX));
    }
    if (kind != 5)
        for (int x = 0; x < 5; x++) { stats[eoTable[x]] += ts[x]; count[eoTable[x]] += tc[x]; }
}

/* common/loopfilter.cpp:141-180: pelFilterLumaStrong_c (chroma = 0: a = tcP, b = tcQ) / pelFilterChroma_c (a = tc, b = maskP, c = maskQ) */
void orc_deblock(int chroma, int depth, void* src, intptr_t srcStep, intptr_t o, int a, int b, int c)
{
#define CL3(lo, hi, v) ((v) < (lo) ? (lo) : ((v) > (hi) ? (hi) : (v)))
    for (int i = 0; i < 4; i++)
    {
        const intptr_t p = i * srcStep;
        int m4 = (int16_t)px(src, depth, p), m3 = (int16_t)px(src, depth, p - o), m5 = (int16_t)px(src, depth, p + o), m2 = (int16_t)px(src, depth, p - 2 * o);
        if (chroma)
        {
            int delta = CL3(-a, a, ((((m4 - m3) * 4) + m2 - m5 + 4) >> 3));
            put_pix(src, depth, p - o, clip_px(m3 + (delta & b), depth));
            put_pix(src, depth, p, clip_px(m4 - (delta & c), depth));
            continue;
        }
        int m6 = (int16_t)px(src, depth, p + 2 * o), m1 = (int16_t)px(src, depth, p - 3 * o), m7 = (int16_t)px(src, depth, p + 3 * o), m0 = (int16_t)px(src, depth, p - 4 * o);
        put_pix(src, depth, p - 3 * o, CL3(-a, a, ((2 * m0 + 3 * m1 + m2 + m3 + m4 + 4) >> 3) - m1) + m1);
        put_pix(src, depth, p - 2 * o, CL3(-a, a, ((m1 + m2 + m3 + m4 + 2) >> 2) - m2) + m2);
        put_pix(src, depth, p - o,     CL3(-a, a, ((m1 + 2 * m2 + 2 * m3 + 2 * m4 + m5 + 4) >> 3) - m3) + m3);
        put_pix(src, depth, p,         CL3(-b, b, ((m2 + 2 * m3 + 2 * m4 + 2 * m5 + m6 + 4) >> 3) - m4) + m4);
        put_pix(src, depth, p + o,     CL3(-b, b, ((m3 + m4 + m5 + m6 + 2) >> 2) - m5) + m5);
        put_pix(src, depth, p + 2 * o, CL3(-b, b, ((m3 + m4 + m5 + 3 * m6 + 2 * m7 + 4) >> 3) - m6) + m6);
    }
#undef CL3
}

/* common/pixel.cpp:914-940 estimateCUPropagateCost (double arithmetic, one rounding per operation) */
void orc_propagate_cost(int* dst, const uint16_t* propagateIn, const int32_t* intraCosts, const uint16_t* interCosts, const int32_t* invQscales,
                        double fpsFactor, int len)
{
    volatile double fps = fpsFactor / 256;
    for (int i = 0; i < len; i++)
    {
        int intraCost = intraCosts[i];
        int inter = interCosts[i] & ((1 << 14) - 1);
        int interCost = intraCost < inter ? intraCost : inter;
        volatile double propagateIntra = (double)(int)((uint32_t)intraCost * (uint32_t)invQscales[i]);
        volatile double t = propagateIntra * fps;
        volatile double propagateAmount = (double)propagateIn[i] + t;
        volatile double propagateNum = (double)(intraCost - interCost);
        volatile double q = propagateAmount * propagateNum;
        volatile double r = q / (double)intraCost;
        double v = r + 0.5;
        dst[i] = (v > -2147483649.0 && v < 2147483648.0) ? (int)v : (int)0x80000000;   /* cvttsd2si semantics */
    }
}

/* the TU chain of Quant::transformNxN + invtransformNxN (common/quant.cpp:397-480, :543-605; rdoq 0, no sign hiding) around
 * cu[].sub_ps / add_ps / sse_pp for one N x N TU; returns numSig */
uint32_t orc_tu_chain(int depth, int sizeIdx, int useDST, const void* fenc, intptr_t fencStride, const void* pred, intptr_t predStride,
                      void* recon, intptr_t reconStride, const int32_t* quantCoeff, int qBits, int add,
                      const int32_t* dequantCoef, int scaleOrPer, int dqShift, int16_t* coeff, uint64_t* sse)
{
    const int N = 4 << sizeIdx, n2 = N * N;
    int16_t resi[32 * 32], dctc[32 * 32];
    int32_t deltaU[32 * 32];
    for (int y = 0; y < N; y++)
        for (int x = 0; x < N; x++) resi[y * N + x] = (int16_t)(px(fenc, depth, y * fencStride + x) - px(pred, depth, y * predStride + x));
    orc_dct(depth, useDST ? 4 : sizeIdx, resi, dctc, N);
    uint32_t numSig = orc_quant(dctc, quantCoeff, deltaU, coeff, qBits, add, n2);
    if (numSig)
    {
        if (dequantCoef) orc_dequant_scaling(coeff, dequantCoef, dctc, n2, scaleOrPer, dqShift);
        else orc_dequant_normal(coeff, dctc, n2, scaleOrPer, dqShift);
        if (numSig == 1 && coeff[0] != 0 && !useDST)
        {
            const int shift_2nd = 12 - (depth - 8) - 3, add_2nd = 1 << (shift_2nd - 1);
            int dc_val = (((dctc[0] * (64 >> 6) + 1) >> 1) * (64 >> 3) + add_2nd) >> shift_2nd;
            for (int i = 0; i < n2; i++) resi[i] = (int16_t)dc_val;
        }
        else orc_idct(depth, useDST ? 4 : sizeIdx, dctc, resi, N);
    }
    else memset(resi, 0, sizeof(int16_t) * n2);
    uint64_t s = 0;
    for (int y = 0; y < N; y++)
        for (int x = 0; x < N; x++)
        {
            int r = clip_px(px(pred, depth, y * predStride + x) + resi[y * N + x], depth);
            put_pix(recon, depth, y * reconStride + x, r);
            int d = px(fenc, depth, y * fencStride + x) - r;
            s += (uint64_t)(d * d);
        }
    *sse = s;
    return numSig;
}

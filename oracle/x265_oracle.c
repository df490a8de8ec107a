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

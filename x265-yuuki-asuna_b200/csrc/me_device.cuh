// me_device.cuh -- warp-cooperative device implementation of x265's MotionEstimate::motionEstimate
// (source/encoder/motion.cpp:739-1569) and subpelCompare (:1571-1664).  One warp runs one PU
// search; every lane executes the same control flow on warp-uniform values (costs are reduced
// with warp shuffles and broadcast), so the data-dependent search path of the reference is
// reproduced step for step, including its quirks (SURVEY.md 8a "quirks").
//
// Shared by me_kernels.cu (full-resolution batched ME) and lookahead_kernels.cu (lowres path).
#pragma once
#include "common.cuh"
#include "tables.cuh"

namespace x265b200 {

enum { ME_DIA = 0, ME_HEX = 1, ME_UMH = 2, ME_STAR = 3, ME_SEA = 4, ME_FULL = 5 };   // x265.h:492-497
#define ME_COST_MAX (1 << 28)                                                       // motion.h:65

struct MV2 { int x, y; };
__device__ __forceinline__ MV2 mv2(int x, int y) { MV2 m; m.x = x; m.y = y; return m; }

// subpel workloads, motion.cpp:48-58
struct SubpelWL { int hpel_iters, hpel_dirs, qpel_iters, qpel_dirs, hpel_satd; };
__constant__ SubpelWL c_workload[8] = {
    { 1, 4, 0, 4, 0 }, { 1, 4, 1, 4, 0 }, { 1, 4, 1, 4, 1 }, { 2, 4, 1, 4, 1 },
    { 2, 4, 2, 4, 1 }, { 1, 8, 1, 8, 1 }, { 2, 8, 1, 8, 1 }, { 2, 8, 2, 8, 1 } };
// pattern tables, motion.cpp:64-84
__constant__ int8_t c_hex2[8][2]    = { {-1,-2}, {-2,0}, {-1,2}, {1,2}, {2,0}, {1,-2}, {-1,-2}, {-2,0} };
__constant__ uint8_t c_mod6m1[8]    = { 5, 0, 1, 2, 3, 4, 5, 0 };
__constant__ int8_t c_square1[9][2] = { {0,0}, {0,-1}, {0,1}, {-1,0}, {1,0}, {-1,-1}, {-1,1}, {1,-1}, {1,1} };
__constant__ int8_t c_hex4[16][2]   = { {0,-4}, {0,4}, {-2,-3}, {2,-3}, {-4,-2}, {4,-2}, {-4,-1}, {4,-1},
                                        {-4,0}, {4,0}, {-4,1}, {4,1}, {-4,2}, {4,2}, {-2,3}, {2,3} };
__constant__ int8_t c_offsets[16][2] = { {-1,0}, {0,-1}, {-1,-1}, {1,-1}, {-1,0}, {1,0}, {-1,1}, {-1,-1},
                                         {1,-1}, {1,1}, {-1,0}, {0,1}, {-1,1}, {1,1}, {1,0}, {0,1} };
__constant__ uint8_t c_rangeMul[4][4] = { {3,3,4,4}, {3,4,4,4}, {4,4,4,5}, {4,4,5,6} };
__constant__ int16_t c_meLumaFilter[4][8] = {
    { 0, 0, 0, 64, 0, 0, 0, 0 }, { -1, 4, -10, 58, 17, -5, 1, 0 },
    { -1, 4, -11, 40, 40, -11, 4, -1 }, { 0, 1, -5, 17, 58, -10, 4, -1 } };

template<typename pixel>
struct MEState
{
    // per-warp shared memory
    pixel*   fenc;       // cached PU, row stride 64 (FENC_STRIDE), motion.cpp:189
    pixel*   pred;       // subpel prediction, row stride = w (motion.cpp:1581 subpelbuf)
    int16_t* immed;      // hvpp intermediate, w * (h + 7)
    // reference
    const pixel* fref;   // fpelPlane[0] + blockOffset
    int64_t  stride;
    const pixel* lowres[4];   // lowres hpel planes + blockOffset (isLowres path), else unused
    bool     isLowres;
    int      w, h, lane, depth, partSizeScale;
    const uint16_t* cost;     // centred lambda-scaled MV cost table (bitcost.cpp:31-60)
    int      mvpx, mvpy;      // setMVP(qmvp), bitcost.h:41
};

constexpr int kMvTableHalf = 2 * 32768;

template<typename pixel>
__device__ __forceinline__ int mvcost(const MEState<pixel>& s, int qx, int qy)
{
    int ix = clip3i(-kMvTableHalf, kMvTableHalf, qx - s.mvpx);
    int iy = clip3i(-kMvTableHalf, kMvTableHalf, qy - s.mvpy);
    return ((int)__ldg(s.cost + ix) + (int)__ldg(s.cost + iy)) & 0xffff;      // bitcost.h:45 returns uint16_t
}

// ---- block compares (warp cooperative) ---------------------------------------------------------
template<typename pixel> __device__ __forceinline__ int sad4_sg(const pixel* fs, const pixel* rg);
template<> __device__ __forceinline__ int sad4_sg<uint8_t>(const uint8_t* fs, const uint8_t* rg)
{
    return (int)__vsadu4(*(const uint32_t*)fs, ld_px4(rg));
}
template<> __device__ __forceinline__ int sad4_sg<uint16_t>(const uint16_t* fs, const uint16_t* rg)
{
    const uint32_t* f = (const uint32_t*)fs;
    return (int)(__vsadu2(f[0], ld_px2(rg)) + __vsadu2(f[1], ld_px2(rg + 2)));
}
template<typename pixel> __device__ __forceinline__ int sad4_ss(const pixel* fs, const pixel* ps);
template<> __device__ __forceinline__ int sad4_ss<uint8_t>(const uint8_t* fs, const uint8_t* ps)
{
    return (int)__vsadu4(*(const uint32_t*)fs, *(const uint32_t*)ps);
}
template<> __device__ __forceinline__ int sad4_ss<uint16_t>(const uint16_t* fs, const uint16_t* ps)
{
    const uint32_t* f = (const uint32_t*)fs; const uint32_t* p = (const uint32_t*)ps;
    return (int)(__vsadu2(f[0], p[0]) + __vsadu2(f[1], p[1]));
}

// SAD of the cached PU against a GLOBAL block (any alignment)
template<typename pixel>
__device__ __forceinline__ int warp_sad_g(const MEState<pixel>& s, const pixel* r, int64_t rs)
{
    const int gw = s.w >> 2, ng = gw * s.h;
    int acc = 0;
    for (int u = s.lane; u < ng; u += 32)
    {
        int y = u / gw, x = (u - y * gw) << 2;
        acc += sad4_sg<pixel>(s.fenc + y * 64 + x, r + (int64_t)y * rs + x);
    }
    return warp_sum(acc);
}
// SAD of the cached PU against a SHARED block with row stride ps (4-pixel aligned)
template<typename pixel>
__device__ __forceinline__ int warp_sad_s(const MEState<pixel>& s, const pixel* p, int ps)
{
    const int gw = s.w >> 2, ng = gw * s.h;
    int acc = 0;
    for (int u = s.lane; u < ng; u += 32)
    {
        int y = u / gw, x = (u - y * gw) << 2;
        acc += sad4_ss<pixel>(s.fenc + y * 64 + x, p + y * ps + x);
    }
    return warp_sum(acc);
}

__device__ __forceinline__ void me_hadamard4(int& a, int& b, int& c, int& d)
{
    int t0 = a + b, t1 = a - b, t2 = c + d, t3 = c - d;
    a = t0 + t2; c = t0 - t2; b = t1 + t3; d = t1 - t3;
}

template<typename pixel> __device__ __forceinline__ void px4_s(const pixel* p, int v[4]);
template<> __device__ __forceinline__ void px4_s<uint8_t>(const uint8_t* p, int v[4])
{
    uint32_t x = *(const uint32_t*)p;
    v[0] = x & 0xff; v[1] = (x >> 8) & 0xff; v[2] = (x >> 16) & 0xff; v[3] = x >> 24;
}
template<> __device__ __forceinline__ void px4_s<uint16_t>(const uint16_t* p, int v[4])
{
    const uint32_t* q = (const uint32_t*)p;
    v[0] = q[0] & 0xffff; v[1] = q[0] >> 16; v[2] = q[1] & 0xffff; v[3] = q[1] >> 16;
}

// SATD (pixel.cpp:210-297): sum over 4x4 cells of (sum|H d H^T| >> 1); REFG: ref in global memory
template<typename pixel, bool REFG>
__device__ __forceinline__ int warp_satd(const MEState<pixel>& s, const pixel* r, int64_t rs)
{
    const int cw = s.w >> 2, nc = cw * (s.h >> 2);
    int acc = 0;
    for (int c = s.lane; c < nc; c += 32)
    {
        int cy = c / cw, cx = c - cy * cw;
        const pixel* f = s.fenc + cy * 4 * 64 + cx * 4;
        const pixel* q = r + (int64_t)cy * 4 * rs + cx * 4;
        int d[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
            int a[4], b[4];
            px4_s<pixel>(f + i * 64, a);
            if (REFG) ld4i<pixel>(q + i * rs, b); else px4_s<pixel>(q + i * rs, b);
#pragma unroll
            for (int k = 0; k < 4; k++) d[i][k] = a[k] - b[k];
            me_hadamard4(d[i][0], d[i][1], d[i][2], d[i][3]);
        }
        int t = 0;
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            me_hadamard4(d[0][k], d[1][k], d[2][k], d[3][k]);
            t += abs(d[0][k]) + abs(d[1][k]) + abs(d[2][k]) + abs(d[3][k]);
        }
        acc += t >> 1;
    }
    return warp_sum(acc);
}

// ---- subpel (motion.cpp:1571-1599, luma only) ---------------------------------------------------
// interpolates the PU at fractional (xFrac,yFrac) of the block starting at `src` into s.pred
template<typename pixel>
__device__ __forceinline__ void warp_interp_luma(const MEState<pixel>& s, const pixel* src, int xFrac, int yFrac)
{
    const int w = s.w, h = s.h, maxVal = (1 << s.depth) - 1, headRoom = 14 - s.depth;
    if (!yFrac || !xFrac)
    {
        // luma_hpp / luma_vpp : ipfilter.cpp:79-118, :164-203
        const int idx = yFrac ? yFrac : xFrac;
        const int64_t step = yFrac ? s.stride : 1;
        int c[8];
#pragma unroll
        for (int t = 0; t < 8; t++) c[t] = c_meLumaFilter[idx][t];
        const pixel* base = src - 3 * step;
        for (int e = s.lane; e < w * h; e += 32)
        {
            int y = e / w, x = e - y * w;
            const pixel* q = base + (int64_t)y * s.stride + x;
            int sum = 0;
#pragma unroll
            for (int t = 0; t < 8; t++) sum += (int)__ldg(q + t * step) * c[t];
            int val = (int16_t)((sum + 32) >> 6);
            s.pred[e] = (pixel)(val < 0 ? 0 : (val > maxVal ? maxVal : val));
        }
    }
    else
    {
        // luma_hvpp = hps(isRowExt) + vsp : ipfilter.cpp:362-369
        int c[8];
#pragma unroll
        for (int t = 0; t < 8; t++) c[t] = c_meLumaFilter[xFrac][t];
        const int shift = 6 - headRoom, offset = (int)((unsigned)-8192 << shift);
        const pixel* base = src - 3 - 3 * s.stride;
        for (int e = s.lane; e < w * (h + 7); e += 32)
        {
            int y = e / w, x = e - y * w;
            const pixel* q = base + (int64_t)y * s.stride + x;
            int sum = 0;
#pragma unroll
            for (int t = 0; t < 8; t++) sum += (int)__ldg(q + t) * c[t];
            s.immed[e] = (int16_t)((sum + offset) >> shift);
        }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 8; t++) c[t] = c_meLumaFilter[yFrac][t];
        const int shift2 = 6 + headRoom, offset2 = (1 << (shift2 - 1)) + (8192 << 6);
        for (int e = s.lane; e < w * h; e += 32)
        {
            int y = e / w, x = e - y * w;
            int sum = 0;
#pragma unroll
            for (int t = 0; t < 8; t++) sum += (int)s.immed[(y + t) * w + x] * c[t];
            int val = (int16_t)((sum + offset2) >> shift2);
            s.pred[e] = (pixel)(val < 0 ? 0 : (val > maxVal ? maxVal : val));
        }
    }
    __syncwarp();
}

// MotionEstimate::subpelCompare (motion.cpp:1571-1599); useSatd selects cmp
template<typename pixel>
__device__ __forceinline__ int subpel_compare(const MEState<pixel>& s, int qx, int qy, bool useSatd)
{
    const pixel* fref = s.fref + (qx >> 2) + (int64_t)(qy >> 2) * s.stride;
    const int xFrac = qx & 3, yFrac = qy & 3;
    if (!(xFrac | yFrac))
        return useSatd ? warp_satd<pixel, true>(s, fref, s.stride) : warp_sad_g<pixel>(s, fref, s.stride);
    __syncwarp();
    warp_interp_luma<pixel>(s, fref, xFrac, yFrac);
    int c = useSatd ? warp_satd<pixel, false>(s, s.pred, s.w) : warp_sad_s<pixel>(s, s.pred, s.w);
    __syncwarp();
    return c;
}

// ReferencePlanes::lowresQPelCost (common/lowres.h:94-120), 8x8 lowres blocks, hme = false
template<typename pixel>
__device__ __forceinline__ int lowres_qpel_cost(const MEState<pixel>& s, int qx, int qy, bool useSatd)
{
    if ((qx | qy) & 1)
    {
        int hpelA = (qy & 2) | ((qx & 2) >> 1);
        const pixel* frefA = s.lowres[hpelA] + (qx >> 2) + (int64_t)(qy >> 2) * s.stride;
        int qmvx = qx + (qx & 1), qmvy = qy + (qy & 1);
        int hpelB = (qmvy & 2) | ((qmvx & 2) >> 1);
        const pixel* frefB = s.lowres[hpelB] + (qmvx >> 2) + (int64_t)(qmvy >> 2) * s.stride;
        __syncwarp();
        for (int e = s.lane; e < s.w * s.h; e += 32)          // pixelavg_pp, pixel.cpp:545-557
        {
            int y = e / s.w, x = e - y * s.w;
            s.pred[e] = (pixel)(((int)__ldg(frefA + (int64_t)y * s.stride + x) + (int)__ldg(frefB + (int64_t)y * s.stride + x) + 1) >> 1);
        }
        __syncwarp();
        int c = useSatd ? warp_satd<pixel, false>(s, s.pred, s.w) : warp_sad_s<pixel>(s, s.pred, s.w);
        __syncwarp();
        return c;
    }
    int hpel = (qy & 2) | ((qx & 2) >> 1);
    const pixel* fref = s.lowres[hpel] + (qx >> 2) + (int64_t)(qy >> 2) * s.stride;
    return useSatd ? warp_satd<pixel, true>(s, fref, s.stride) : warp_sad_g<pixel>(s, fref, s.stride);
}

// ---- the search ------------------------------------------------------------------------------------
template<typename pixel>
struct MESearch
{
    const MEState<pixel>& s;
    MV2 mvmin, mvmax;
    MV2 bmv; int bcost;

    __device__ __forceinline__ MESearch(const MEState<pixel>& st) : s(st) {}

    __device__ __forceinline__ int sadAt(int mx, int my) const { return warp_sad_g<pixel>(s, s.fref + mx + (int64_t)my * s.stride, s.stride); }
    __device__ __forceinline__ int fcost(int mx, int my) const { return mvcost(s, mx << 2, my << 2); }
    __device__ __forceinline__ bool inRange(int x, int y) const { return x >= mvmin.x && x <= mvmax.x && y >= mvmin.y && y <= mvmax.y; }
    __device__ __forceinline__ bool yOk(int y) const { return (y >= mvmin.y) & (y <= mvmax.y); }

    // COST_MV (motion.cpp:238-244)
    __device__ __forceinline__ void costMv(int mx, int my)
    {
        int cost = sadAt(mx, my) + fcost(mx, my);
        if (cost < bcost) { bcost = cost; bmv = mv2(mx, my); }
    }
    // COST_MV_X4 (motion.cpp:277-298): only the y range is checked (quirk)
    __device__ __forceinline__ void costMvX4(MV2 omv, int x0, int y0, int x1, int y1, int x2, int y2, int x3, int y3)
    {
        const int dx[4] = { x0, x1, x2, x3 }, dy[4] = { y0, y1, y2, y3 };
        int costs[4];
#pragma unroll
        for (int k = 0; k < 4; k++) costs[k] = sadAt(omv.x + dx[k], omv.y + dy[k]) + fcost(omv.x + dx[k], omv.y + dy[k]);
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (yOk(omv.y + dy[k]) && costs[k] < bcost) { bcost = costs[k]; bmv = mv2(omv.x + dx[k], omv.y + dy[k]); }
    }
    // COST_MV_X4_DIR / COST_MV_X3_DIR (motion.cpp:246-257, :315-328): costs relative to bmv, no update
    __device__ __forceinline__ void dirCosts(int n, const int dx[], const int dy[], int costs[]) const
    {
        for (int k = 0; k < n; k++) costs[k] = sadAt(bmv.x + dx[k], bmv.y + dy[k]) + fcost(bmv.x + dx[k], bmv.y + dy[k]);
    }
    // CROSS (motion.cpp:336-360)
    __device__ __forceinline__ void cross(MV2 omv, int start, int x_max, int y_max)
    {
        int i = start;
        if (x_max <= min(mvmax.x - omv.x, omv.x - mvmin.x))
            for (; i < x_max - 2; i += 4) costMvX4(omv, i, 0, -i, 0, i + 2, 0, -i - 2, 0);
        for (; i < x_max; i += 2)
        {
            if (omv.x + i <= mvmax.x) costMv(omv.x + i, omv.y);
            if (omv.x - i >= mvmin.x) costMv(omv.x - i, omv.y);
        }
        i = start;
        if (y_max <= min(mvmax.y - omv.y, omv.y - mvmin.y))
            for (; i < y_max - 2; i += 4) costMvX4(omv, 0, i, 0, -i, 0, i + 2, 0, -i - 2);
        for (; i < y_max; i += 2)
        {
            if (omv.y + i <= mvmax.y) costMv(omv.x, omv.y + i);
            if (omv.y - i >= mvmin.y) costMv(omv.x, omv.y - i);
        }
    }

    // COST_MV_PT_DIST (motion.cpp:224-236)
    __device__ __forceinline__ void ptDist(int mx, int my, int point, int dist, int& bPointNr, int& bDistance)
    {
        int cost = sadAt(mx, my) + fcost(mx, my);
        if (cost < bcost) { bcost = cost; bmv = mv2(mx, my); bPointNr = point; bDistance = dist; }
    }

    // StarPatternSearch (motion.cpp:362-604)
    __device__ void starPattern(int& bPointNr, int& bDistance, int earlyExitIters, int merange)
    {
        const MV2 omv = bmv;
        int saved = bcost, rounds = 0;
        {
            const int dist = 1;
            const int top = omv.y - dist, bottom = omv.y + dist, left = omv.x - dist, right = omv.x + dist;
            // the x4 form and the guarded form evaluate the same points in the same order
            if (top >= mvmin.y) ptDist(omv.x, top, 2, dist, bPointNr, bDistance);
            if (left >= mvmin.x) ptDist(left, omv.y, 4, dist, bPointNr, bDistance);
            if (right <= mvmax.x) ptDist(right, omv.y, 5, dist, bPointNr, bDistance);
            if (bottom <= mvmax.y) ptDist(omv.x, bottom, 7, dist, bPointNr, bDistance);
            if (bcost < saved) rounds = 0;
            else if (++rounds >= earlyExitIters) return;
        }
        for (int dist = 2; dist <= 8; dist <<= 1)
        {
            const int top = omv.y - dist, bottom = omv.y + dist, left = omv.x - dist, right = omv.x + dist;
            const int top2 = omv.y - (dist >> 1), bottom2 = omv.y + (dist >> 1), left2 = omv.x - (dist >> 1), right2 = omv.x + (dist >> 1);
            saved = bcost;
            if (top >= mvmin.y && left >= mvmin.x && right <= mvmax.x && bottom <= mvmax.y)
            {
                // x4 order (motion.cpp:451-458): 2,1,3,4 then 5,6,8,7
                ptDist(omv.x, top, 2, dist, bPointNr, bDistance);
                ptDist(left2, top2, 1, dist >> 1, bPointNr, bDistance);
                ptDist(right2, top2, 3, dist >> 1, bPointNr, bDistance);
                ptDist(left, omv.y, 4, dist, bPointNr, bDistance);
                ptDist(right, omv.y, 5, dist, bPointNr, bDistance);
                ptDist(left2, bottom2, 6, dist >> 1, bPointNr, bDistance);
                ptDist(right2, bottom2, 8, dist >> 1, bPointNr, bDistance);
                ptDist(omv.x, bottom, 7, dist, bPointNr, bDistance);
            }
            else
            {
                if (top >= mvmin.y) ptDist(omv.x, top, 2, dist, bPointNr, bDistance);
                if (top2 >= mvmin.y)
                {
                    if (left2 >= mvmin.x) ptDist(left2, top2, 1, dist >> 1, bPointNr, bDistance);
                    if (right2 <= mvmax.x) ptDist(right2, top2, 3, dist >> 1, bPointNr, bDistance);
                }
                if (left >= mvmin.x) ptDist(left, omv.y, 4, dist, bPointNr, bDistance);
                if (right <= mvmax.x) ptDist(right, omv.y, 5, dist, bPointNr, bDistance);
                if (bottom2 <= mvmax.y)
                {
                    if (left2 >= mvmin.x) ptDist(left2, bottom2, 6, dist >> 1, bPointNr, bDistance);
                    if (right2 <= mvmax.x) ptDist(right2, bottom2, 8, dist >> 1, bPointNr, bDistance);
                }
                if (bottom <= mvmax.y) ptDist(omv.x, bottom, 7, dist, bPointNr, bDistance);
            }
            if (bcost < saved) rounds = 0;
            else if (++rounds >= earlyExitIters) return;
        }
        for (int dist = 16; dist <= (int)(int16_t)merange; dist <<= 1)
        {
            const int top = omv.y - dist, bottom = omv.y + dist, left = omv.x - dist, right = omv.x + dist;
            saved = bcost;
            if (top >= mvmin.y && left >= mvmin.x && right <= mvmax.x && bottom <= mvmax.y)
            {
                ptDist(omv.x, top, 0, dist, bPointNr, bDistance);
                ptDist(left, omv.y, 0, dist, bPointNr, bDistance);
                ptDist(right, omv.y, 0, dist, bPointNr, bDistance);
                ptDist(omv.x, bottom, 0, dist, bPointNr, bDistance);
                for (int index = 1; index < 4; index++)
                {
                    int posYT = top + ((dist >> 2) * index), posYB = bottom - ((dist >> 2) * index);
                    int posXL = omv.x - ((dist >> 2) * index), posXR = omv.x + ((dist >> 2) * index);
                    ptDist(posXL, posYT, 0, dist, bPointNr, bDistance);
                    ptDist(posXR, posYT, 0, dist, bPointNr, bDistance);
                    ptDist(posXL, posYB, 0, dist, bPointNr, bDistance);
                    ptDist(posXR, posYB, 0, dist, bPointNr, bDistance);
                }
            }
            else
            {
                if (top >= mvmin.y) ptDist(omv.x, top, 0, dist, bPointNr, bDistance);
                if (left >= mvmin.x) ptDist(left, omv.y, 0, dist, bPointNr, bDistance);
                if (right <= mvmax.x) ptDist(right, omv.y, 0, dist, bPointNr, bDistance);
                if (bottom <= mvmax.y) ptDist(omv.x, bottom, 0, dist, bPointNr, bDistance);
                for (int index = 1; index < 4; index++)
                {
                    int posYT = top + ((dist >> 2) * index), posYB = bottom - ((dist >> 2) * index);
                    int posXL = omv.x - ((dist >> 2) * index), posXR = omv.x + ((dist >> 2) * index);
                    if (posYT >= mvmin.y)
                    {
                        if (posXL >= mvmin.x) ptDist(posXL, posYT, 0, dist, bPointNr, bDistance);
                        if (posXR <= mvmax.x) ptDist(posXR, posYT, 0, dist, bPointNr, bDistance);
                    }
                    if (posYB <= mvmax.y)
                    {
                        if (posXL >= mvmin.x) ptDist(posXL, posYB, 0, dist, bPointNr, bDistance);
                        if (posXR <= mvmax.x) ptDist(posXR, posYB, 0, dist, bPointNr, bDistance);
                    }
                }
            }
            if (bcost < saved) rounds = 0;
            else if (++rounds >= earlyExitIters) return;
        }
    }

    // square refine shared by HEX (motion.cpp:924-942)
    __device__ __forceinline__ void squareRefine()
    {
        int costs[4];
        int dir = 0;
        { const int dx[4] = { 0, 0, -1, 1 }, dy[4] = { -1, 1, 0, 0 }; dirCosts(4, dx, dy, costs); }
        if (yOk(bmv.y - 1) && costs[0] < bcost) { bcost = costs[0]; dir = 1; }
        if (yOk(bmv.y + 1) && costs[1] < bcost) { bcost = costs[1]; dir = 2; }
        if (costs[2] < bcost) { bcost = costs[2]; dir = 3; }
        if (costs[3] < bcost) { bcost = costs[3]; dir = 4; }
        { const int dx[4] = { -1, -1, 1, 1 }, dy[4] = { -1, 1, -1, 1 }; dirCosts(4, dx, dy, costs); }
        if (yOk(bmv.y - 1) && costs[0] < bcost) { bcost = costs[0]; dir = 5; }
        if (yOk(bmv.y + 1) && costs[1] < bcost) { bcost = costs[1]; dir = 6; }
        if (yOk(bmv.y - 1) && costs[2] < bcost) { bcost = costs[2]; dir = 7; }
        if (yOk(bmv.y + 1) && costs[3] < bcost) { bcost = costs[3]; dir = 8; }
        bmv.x += c_square1[dir][0]; bmv.y += c_square1[dir][1];
    }

    // me_hex2 (motion.cpp:845-944)
    __device__ void hexSearch(int merange)
    {
        int costs[4];
        { const int dx[3] = { -2, -1, 1 }, dy[3] = { 0, 2, 2 }; dirCosts(3, dx, dy, costs); }
        bcost <<= 3;
        if (yOk(bmv.y) && (costs[0] << 3) + 2 < bcost) bcost = (costs[0] << 3) + 2;
        if (yOk(bmv.y + 2))
        {
            if ((costs[1] << 3) + 3 < bcost) bcost = (costs[1] << 3) + 3;
            if ((costs[2] << 3) + 4 < bcost) bcost = (costs[2] << 3) + 4;
        }
        { const int dx[3] = { 2, 1, -1 }, dy[3] = { 0, -2, -2 }; dirCosts(3, dx, dy, costs); }
        if (yOk(bmv.y) && (costs[0] << 3) + 5 < bcost) bcost = (costs[0] << 3) + 5;
        if (yOk(bmv.y - 2))
        {
            if ((costs[1] << 3) + 6 < bcost) bcost = (costs[1] << 3) + 6;
            if ((costs[2] << 3) + 7 < bcost) bcost = (costs[2] << 3) + 7;
        }
        if (bcost & 7)
        {
            int dir = (bcost & 7) - 2;
            if (yOk(bmv.y + c_hex2[dir + 1][1]))
            {
                bmv.x += c_hex2[dir + 1][0]; bmv.y += c_hex2[dir + 1][1];
                for (int i = (merange >> 1) - 1; i > 0 && inRange(bmv.x, bmv.y); i--)
                {
                    const int dx[3] = { c_hex2[dir][0], c_hex2[dir + 1][0], c_hex2[dir + 2][0] };
                    const int dy[3] = { c_hex2[dir][1], c_hex2[dir + 1][1], c_hex2[dir + 2][1] };
                    dirCosts(3, dx, dy, costs);
                    bcost &= ~7;
                    if (yOk(bmv.y + dy[0]) && (costs[0] << 3) + 1 < bcost) bcost = (costs[0] << 3) + 1;
                    if (yOk(bmv.y + dy[1]) && (costs[1] << 3) + 2 < bcost) bcost = (costs[1] << 3) + 2;
                    if (yOk(bmv.y + dy[2]) && (costs[2] << 3) + 3 < bcost) bcost = (costs[2] << 3) + 3;
                    if (!(bcost & 7)) break;
                    dir += (bcost & 7) - 2;
                    dir = c_mod6m1[dir + 1];
                    bmv.x += c_hex2[dir + 1][0]; bmv.y += c_hex2[dir + 1][1];
                }
            }
        }
        bcost >>= 3;
        squareRefine();
    }

    // two-point refinement after a distance-1 star result (motion.cpp:1151-1166, :1225-1235)
    __device__ __forceinline__ void twoPoint(int bPointNr)
    {
        const MV2 b = bmv;
        const int x1 = b.x + c_offsets[(bPointNr - 1) * 2][0], y1 = b.y + c_offsets[(bPointNr - 1) * 2][1];
        const int x2 = b.x + c_offsets[(bPointNr - 1) * 2 + 1][0], y2 = b.y + c_offsets[(bPointNr - 1) * 2 + 1][1];
        if (inRange(x1, y1)) costMv(x1, y1);
        if (inRange(x2, y2)) costMv(x2, y2);
    }
};

// Full MotionEstimate::motionEstimate.  Returns the cost; (outx,outy) = outQMv.
template<typename pixel>
__device__ int motion_estimate(const MEState<pixel>& s, MV2 mvmin, MV2 mvmax, MV2 qmvp, int numCand, const int* mvc /* [numCand][2] */,
                               int merange, int searchMethod, int subpelRefine, int maxSlices, int partEnumIs64, int& outx, int& outy)
{
    MESearch<pixel> S(s);
    S.mvmin = mvmin; S.mvmax = mvmax;
    const MV2 qmvmin = mv2(mvmin.x << 2, mvmin.y << 2), qmvmax = mv2(mvmax.x << 2, mvmax.y << 2);

    // measure SAD cost at clipped QPEL MVP (motion.cpp:770-778)
    MV2 pmv = mv2(max(min(qmvp.x, qmvmax.x), qmvmin.x), max(min(qmvp.y, qmvmax.y), qmvmin.y));
    MV2 bestpre = pmv;
    int bprecost = s.isLowres ? lowres_qpel_cost<pixel>(s, pmv.x, pmv.y, false) : subpel_compare<pixel>(s, pmv.x, pmv.y, false);

    // re-measure full pel rounded MVP with SAD as search start point (:780-784)
    S.bmv = mv2((pmv.x + 2) >> 2, (pmv.y + 2) >> 2);
    S.bcost = bprecost;
    if ((pmv.x | pmv.y) & 3)
        S.bcost = S.sadAt(S.bmv.x, S.bmv.y) + S.fcost(S.bmv.x, S.bmv.y);

    // measure SAD cost at MV(0) if MVP is not zero (:786-796)
    if (pmv.x | pmv.y)
    {
        int cost = S.sadAt(0, 0) + mvcost(s, 0, 0);
        if (cost < S.bcost)
        {
            S.bcost = cost;
            S.bmv = mv2(0, max(min(0, mvmax.y), mvmin.y));      // quirk: only y is clamped (:794)
        }
    }

    // each QPEL candidate (:799-812)
    for (int i = 0; i < numCand; i++)
    {
        MV2 m = mv2(max(min(mvc[2 * i], qmvmax.x), qmvmin.x), max(min(mvc[2 * i + 1], qmvmax.y), qmvmin.y));
        bool notZero = (m.x | m.y) != 0, nePmv = (m.x != pmv.x) | (m.y != pmv.y), nePre = (m.x != bestpre.x) | (m.y != bestpre.y);
        if (notZero & nePmv & nePre)
        {
            int cost = subpel_compare<pixel>(s, m.x, m.y, false) + mvcost(s, m.x, m.y);
            if (cost < bprecost) { bprecost = cost; bestpre = m; }
        }
    }

    pmv = mv2((pmv.x + 2) >> 2, (pmv.y + 2) >> 2);
    MV2 omv = S.bmv;

    switch (searchMethod)
    {
    case ME_DIA:      // motion.cpp:820-843
    {
        S.bcost <<= 4;
        int i = merange;
        do
        {
            int costs[4];
            const int dx[4] = { 0, 0, -1, 1 }, dy[4] = { -1, 1, 0, 0 };
            S.dirCosts(4, dx, dy, costs);
            if (S.yOk(S.bmv.y - 1) && (costs[0] << 4) + 1 < S.bcost) S.bcost = (costs[0] << 4) + 1;
            if (S.yOk(S.bmv.y + 1) && (costs[1] << 4) + 3 < S.bcost) S.bcost = (costs[1] << 4) + 3;
            if ((costs[2] << 4) + 4 < S.bcost) S.bcost = (costs[2] << 4) + 4;
            if ((costs[3] << 4) + 12 < S.bcost) S.bcost = (costs[3] << 4) + 12;
            if (!(S.bcost & 15)) break;
            S.bmv.x -= ((int)((unsigned)S.bcost << 28)) >> 30;
            S.bmv.y -= ((int)((unsigned)S.bcost << 30)) >> 30;
            S.bcost &= ~15;
        }
        while (--i && S.inRange(S.bmv.x, S.bmv.y));
        S.bcost >>= 4;
        break;
    }
    case ME_HEX:
        S.hexSearch(merange);
        break;
    case ME_UMH:      // motion.cpp:946-1130
    {
        int ucost1, ucost2;
        int cross_start = 1;
        bool done = false;
        omv = S.bmv;
        ucost1 = S.bcost;
        S.costMvX4(mv2(pmv.x, pmv.y), 0, -1, 0, 1, -1, 0, 1, 0);               // DIA1_ITER(pmv)
        if (pmv.x | pmv.y) S.costMvX4(mv2(0, 0), 0, -1, 0, 1, -1, 0, 1, 0);
        ucost2 = S.bcost;
        if ((S.bmv.x | S.bmv.y) && ((S.bmv.x != pmv.x) | (S.bmv.y != pmv.y)))
            S.costMvX4(S.bmv, 0, -1, 0, 1, -1, 0, 1, 0);
        if (S.bcost == ucost2) cross_start = 3;

        omv = S.bmv;
#define ME_SAD_THRESH(v) (S.bcost < (((v) >> 4) * s.partSizeScale))
        if (S.bcost == ucost2 && ME_SAD_THRESH(2000))
        {
            S.costMvX4(omv, 0, -2, -1, -1, 1, -1, -2, 0);
            S.costMvX4(omv, 2, 0, -1, 1, 1, 1, 0, 2);
            if (S.bcost == ucost1 && ME_SAD_THRESH(500)) done = true;
            else if (S.bcost == ucost2)
            {
                int range = (int)(int16_t)((merange >> 1) | 1);
                S.cross(omv, 3, range, range);
                S.costMvX4(omv, -1, -2, 1, -2, -2, -1, 2, -1);
                S.costMvX4(omv, -2, 1, 2, 1, -1, 2, 1, 2);
                if (S.bcost == ucost2) done = true;
                else cross_start = range + 2;
            }
        }
        if (done) break;

        if (numCand)      // adaptive range (:989-1040)
        {
            int mvd, denom = 1;
            if (numCand == 1)
            {
                if (partEnumIs64) mvd = 25;
                else mvd = abs(qmvp.x - mvc[0]) + abs(qmvp.y - mvc[1]);
            }
            else
            {
                denom = numCand - 1;
                mvd = 0;
                if (!partEnumIs64) { mvd = abs(qmvp.x - mvc[0]) + abs(qmvp.y - mvc[1]); denom++; }
                for (int i = 0; i < numCand - 1; i++)
                    mvd += abs(mvc[2 * i] - mvc[2 * i + 2]) + abs(mvc[2 * i + 1] - mvc[2 * i + 3]);
            }
            int sad_ctx = ME_SAD_THRESH(1000) ? 0 : ME_SAD_THRESH(2000) ? 1 : ME_SAD_THRESH(4000) ? 2 : 3;
            int mvd_ctx = mvd < 10 * denom ? 0 : mvd < 20 * denom ? 1 : mvd < 40 * denom ? 2 : 3;
            merange = (merange * c_rangeMul[mvd_ctx][sad_ctx]) >> 2;
        }
#undef ME_SAD_THRESH
        S.cross(omv, cross_start, merange, merange >> 1);
        S.costMvX4(omv, -2, -2, -2, 2, 2, -2, 2, 2);

        // hexagon grid (:1047-1126)
        omv = S.bmv;
        int i = 1;
        do
        {
            if (4 * i > min(min(mvmax.x - omv.x, omv.x - mvmin.x), min(mvmax.y - omv.y, omv.y - mvmin.y)))
            {
                for (int j = 0; j < 16; j++)
                {
                    int mx = omv.x + c_hex4[j][0] * i, my = omv.y + c_hex4[j][1] * i;
                    if (S.inRange(mx, my)) S.costMv(mx, my);
                }
            }
            else
            {
                int dir = 0;
                for (int k = 0; k < 16; k++)
                {
                    int hx = c_hex4[k][0], hy = c_hex4[k][1];
                    int mx = omv.x + hx * i, my = omv.y + hy * i;
                    int cost = S.sadAt(mx, my) + S.fcost(mx, my);
                    // MIN_MV checks the UNSCALED dy (quirk, :1078)
                    if (S.yOk(omv.y + hy) && cost < S.bcost) { S.bcost = cost; dir = hx * 16 + (hy & 15); }
                }
                if (dir)
                {
                    S.bmv.x = omv.x + i * (dir >> 4);
                    S.bmv.y = omv.y + i * (((int)((unsigned)dir << 28)) >> 28);
                }
            }
        }
        while (++i <= merange >> 2);
        if (S.inRange(S.bmv.x, S.bmv.y)) S.hexSearch(merange);
        break;
    }
    case ME_STAR:     // motion.cpp:1132-1240
    {
        int bPointNr = 0, bDistance = 0;
        S.starPattern(bPointNr, bDistance, 3, merange);
        bool stop = false;
        if (bDistance == 1)
        {
            if (bPointNr)
            {
                int saved = S.bcost;
                S.twoPoint(bPointNr);
                if (S.bcost == saved) stop = true;
            }
            else stop = true;
        }
        if (stop) break;
        const int RasterDistance = 5;
        if (bDistance > RasterDistance)
        {
            for (int ty = mvmin.y; ty <= mvmax.y; ty += RasterDistance)
                for (int tx = mvmin.x; tx <= mvmax.x; tx += RasterDistance)
                {
                    if (tx + RasterDistance * 3 <= mvmax.x)
                    {
                        int c0 = S.sadAt(tx, ty), c1 = S.sadAt(tx + 5, ty), c2 = S.sadAt(tx + 10, ty), c3 = S.sadAt(tx + 15, ty);
                        c0 += S.fcost(tx, ty);
                        if (c0 < S.bcost) { S.bcost = c0; S.bmv = mv2(tx, ty); }
                        tx += RasterDistance;
                        c1 += S.fcost(tx, ty);
                        if (c1 < S.bcost) { S.bcost = c1; S.bmv = mv2(tx, ty); }
                        tx += RasterDistance;
                        c2 += S.fcost(tx, ty);
                        if (c2 < S.bcost) { S.bcost = c2; S.bmv = mv2(tx, ty); }
                        tx += RasterDistance;
                        c3 += mvcost(s, tx << 3, ty << 3);                   // quirk: << 3 (motion.cpp:1196)
                        if (c3 < S.bcost) { S.bcost = c3; S.bmv = mv2(tx, ty); }
                    }
                    else
                        S.costMv(tx, ty);
                }
        }
        while (bDistance > 0)
        {
            bDistance = 0; bPointNr = 0;
            S.starPattern(bPointNr, bDistance, 32, merange);
            if (bDistance == 1)
            {
                if (!bPointNr) break;
                S.twoPoint(bPointNr);
                break;
            }
        }
        break;
    }
    case ME_FULL:     // motion.cpp:1397-1441 (non-HME)
    {
        for (int ty = mvmin.y; ty <= mvmax.y; ty++)
            for (int tx = mvmin.x; tx <= mvmax.x; tx++)
                S.costMv(tx, ty);       // the x4 grouping of the reference evaluates the same points in the same order
        break;
    }
    default:
        break;
    }

    // choose between the search result and the best predictor (:1448-1454)
    MV2 bmv; int bcost;
    if (bprecost < S.bcost) { bmv = bestpre; bcost = bprecost; }
    else { bmv = mv2(S.bmv.x << 2, S.bmv.y << 2); bcost = S.bcost; }

    const SubpelWL wl = c_workload[subpelRefine];

    // slice bound clamp (:1458-1463)
    if ((maxSlices > 1) & ((bmv.y < qmvmin.y) | (bmv.y > qmvmax.y)))
    {
        bmv.y = min(max(bmv.y, qmvmin.y), qmvmax.y);
        bcost = subpel_compare<pixel>(s, bmv.x, bmv.y, true) + mvcost(s, bmv.x, bmv.y);
    }

    if (!bcost)
        bcost = mvcost(s, bmv.x, bmv.y);                                  // :1465-1470
    else if (s.isLowres)                                                  // :1471-1503
    {
        int bdir = 0;
        for (int i = 1; i <= wl.hpel_dirs; i++)
        {
            int qx = bmv.x + c_square1[i][0] * 2, qy = bmv.y + c_square1[i][1] * 2;
            if ((qy < qmvmin.y) | (qy > qmvmax.y)) continue;
            int cost = lowres_qpel_cost<pixel>(s, qx, qy, false) + mvcost(s, qx, qy);
            if (cost < bcost) { bcost = cost; bdir = i; }
        }
        bmv.x += c_square1[bdir][0] * 2; bmv.y += c_square1[bdir][1] * 2;
        bcost = lowres_qpel_cost<pixel>(s, bmv.x, bmv.y, true) + mvcost(s, bmv.x, bmv.y);
        bdir = 0;
        for (int i = 1; i <= wl.qpel_dirs; i++)
        {
            int qx = bmv.x + c_square1[i][0], qy = bmv.y + c_square1[i][1];
            if ((qy < qmvmin.y) | (qy > qmvmax.y)) continue;
            int cost = lowres_qpel_cost<pixel>(s, qx, qy, true) + mvcost(s, qx, qy);
            if (cost < bcost) { bcost = cost; bdir = i; }
        }
        bmv.x += c_square1[bdir][0]; bmv.y += c_square1[bdir][1];
    }
    else                                                                  // :1504-1561
    {
        bool hpelSatd = wl.hpel_satd != 0;
        if (hpelSatd)
            bcost = subpel_compare<pixel>(s, bmv.x, bmv.y, true) + mvcost(s, bmv.x, bmv.y);
        for (int iter = 0; iter < wl.hpel_iters; iter++)
        {
            int bdir = 0;
            for (int i = 1; i <= wl.hpel_dirs; i++)
            {
                int qx = bmv.x + c_square1[i][0] * 2, qy = bmv.y + c_square1[i][1] * 2;
                if ((qy < qmvmin.y) | (qy > qmvmax.y)) continue;
                int cost = subpel_compare<pixel>(s, qx, qy, hpelSatd) + mvcost(s, qx, qy);
                if (cost < bcost) { bcost = cost; bdir = i; }
            }
            if (bdir) { bmv.x += c_square1[bdir][0] * 2; bmv.y += c_square1[bdir][1] * 2; }
            else break;
        }
        if (!hpelSatd)
            bcost = subpel_compare<pixel>(s, bmv.x, bmv.y, true) + mvcost(s, bmv.x, bmv.y);
        for (int iter = 0; iter < wl.qpel_iters; iter++)
        {
            int bdir = 0;
            for (int i = 1; i <= wl.qpel_dirs; i++)
            {
                int qx = bmv.x + c_square1[i][0], qy = bmv.y + c_square1[i][1];
                if ((qy < qmvmin.y) | (qy > qmvmax.y)) continue;
                int cost = subpel_compare<pixel>(s, qx, qy, true) + mvcost(s, qx, qy);
                if (cost < bcost) { bcost = cost; bdir = i; }
            }
            if (bdir) { bmv.x += c_square1[bdir][0]; bmv.y += c_square1[bdir][1]; }
            else break;
        }
    }
    outx = bmv.x; outy = bmv.y;
    return bcost;
}

} // namespace x265b200

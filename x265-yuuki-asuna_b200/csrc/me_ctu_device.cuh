// me_ctu_device.cuh -- per-lane body of the general frame search (me_ctu_kernels.cu), kept in a header so that the host
// emulation (tests/host_emu/me_ctu_emu.cpp) compiles and runs exactly this code.
//
// One CTA = one (CTU, reference) pair.  The reference search window, the source CTU and -- when the chroma SATD term of
// subpelCompare is on (subme > 2, motion.cpp:212) -- the Cb / Cr windows and source blocks sit in shared memory; the PUs of
// the CTU (any partition set: 2Nx2N, rect, AMP, CTU 64 / 32 / 16) are searched by work items of 32 lanes taken from a
// per-CTA queue.  A lane owns a sub-block of its PU and runs the whole bit-exact search of me_device.cuh on it; the lanes of
// a PU add their partial costs, so they hold identical costs and follow identical control flow.
// Per-PU semantics = Search::predInterSearch's call (search.cpp:2181-2436): setSearchRange(cu, mvp, merange) (:2724-2769,
// incl. CUData::clipMv, the slice rows and the reference-lag clamp), then motionEstimate(ref, mvmin, mvmax, mvp, numCand, mvc,
// merange, outMV, maxSlices) with the encode-style setSourcePU (bChromaSATD).
#pragma once
#define ME_CTU_KERNEL 1
#define ME_FORCE_THREAD 1
#define ME_FULLRES_ONLY 1
#define ME_REF_IN_SMEM 1
#define ME_WINDOW_FAST 1
#define ME_SUBPEL_PACKED 1
#define ME_BATCH_GROUPSUM 1
#define ME_VCELL_REUSE 1
#define ME_LANE_W_VAR 1           /* lanes own 4- or 8-wide sub-blocks                                     */
#define ME_WINDOW_CHECK 1         /* blocks outside the staged window are read from the global plane       */
#define ME_COST_SMEM 1            /* the MV-cost entries a search can reach sit in shared memory          */
#define ME_THREAD_CHROMA 1        /* per-thread chroma SATD term                                           */
#include "me_device.cuh"
#include "me_ctu_layout.h"

namespace x265b200 {

constexpr int MC_MAX_REFS = 8;

struct MECtuArgs
{
    const void* const* refY;          // device arrays [numRefs] of plane ORIGINS (pixel 0,0)
    const void* const* refCb;
    const void* const* refCr;
    int64_t refStride, refStrideC;
    int ctuCols, ctuRows, numRefs, ctuSize;
    int marginX, marginY, cmarginX, cmarginY;
    int picW, picH;                   // luma samples (CUData::clipMv; CUs that leave the picture are not searched)
    int firstCtuRow;                  // CTU row of the plane origin inside the whole frame (band calls); picH is the frame's
    const MECtuPU* pus; int numPu;
    const uint32_t* items; int numItems;
    const int32_t* mvpCtu;            // [ref][ctu][2] quarter-pel: window centre and default predictor, or null (0)
    const int32_t* mvpPu;             // [ref][ctu][pu][2] or null
    const uint8_t* numCandPu;         // [ref][ctu][pu] or null
    const int32_t* mvcPu;             // [ref][ctu][pu][maxCand][2]
    int maxCand;
    int32_t* out;                     // [ref][ctu][pu][3]
    const uint16_t* cost;             // full lambda-scaled table (global), centred at + 2*32768
    int costK;
    int searchMethod, subpelRefine, merange, depth, R;
    int winPitch, winRows;            // luma window: pixels per row (16-byte multiple), rows staged
    int csp, hshift, vshift, chromaSatd;
    int cwinPitch, cwinRows;
    const int32_t* sliceBounds;       // [ctuRows][2] quarter-pel m_sliceMinY / m_sliceMaxY (frameencoder.cpp:1448-1453), or null
    int maxSlices;
    int refLagPixels;                 // Search::m_refLagPixels (search.cpp:92), full-pel
    int rawRange;                     // 1: mvmin / mvmax = (mvp >> 2) -+ merange, no picture clipping (x265b200_me_frame_dev semantics)
};

// where the CTA's staged data lives (shared memory) and which picture area it shows
template<typename pixel>
struct MECtuStage
{
    const pixel* window;   int winLeft, winTop;          // picture coordinates of window[0]
    const pixel* cwindow[2]; int cwinLeft, cwinTop;      // chroma sample coordinates
    const pixel* fenc;                                   // source CTU, row pitch 64
    const pixel* fencC[2];                               // source Cb / Cr of the CTU, row pitch 64
    const uint16_t* costS;                               // cost[-costK .. costK], pointer to the centre
};

__device__ __forceinline__ int me_ctu_clip(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

// One lane of one work item.  `laneInWarp` positions the lane's group inside the warp (shuffle mask).
template<typename pixel>
__device__ void me_ctu_lane(const MECtuArgs& p, const MECtuStage<pixel>& st, uint32_t word, int ctuX, int ctuY, int ref, int laneInWarp)
{
    const int puIdx = (int)(word & 1023u);
    const int sx = (int)((word >> 10) & 15u) << 2, sy = (int)((word >> 14) & 15u) << 2;
    const int sw = ((word >> 18) & 1u) ? 8 : 4, sh = (int)((word >> 19) & 31u) << 2;
    const int gLog2 = (int)((word >> 24) & 7u), G = 1 << gLog2;
    const MECtuPU pu = p.pus[puIdx];
    const int C = p.ctuSize;
    const int ctu = ctuY * p.ctuCols + ctuX;
    const int64_t puSlot = ((int64_t)ref * p.ctuCols * p.ctuRows + ctu) * p.numPu + puIdx;
    const int q = laneInWarp & (G - 1);
    int32_t* o = p.out + puSlot * 3;

    // CUs that leave the picture are never analysed (the CU quadtree splits until it fits, analysis.cpp compressCTU)
    const int cuPelX = ctuX * C + pu.cuX, cuPelY = (ctuY + p.firstCtuRow) * C + pu.cuY;       // frame coordinates
    if (cuPelX + pu.cuSize > p.picW || cuPelY + pu.cuSize > p.picH)
    {
        if (q == 0) { o[0] = 0; o[1] = 0; o[2] = -1; }
        return;
    }
    const int PX = ctuX * C + pu.x, PY = ctuY * C + pu.y;

    int mvpx = 0, mvpy = 0;
    if (p.mvpPu) { mvpx = p.mvpPu[puSlot * 2]; mvpy = p.mvpPu[puSlot * 2 + 1]; }
    else if (p.mvpCtu) { const int32_t* m = p.mvpCtu + ((int64_t)ref * p.ctuCols * p.ctuRows + ctu) * 2; mvpx = m[0]; mvpy = m[1]; }

    // Search::setSearchRange (search.cpp:2724-2769)
    int minx = mvpx - (p.merange << 2), miny = mvpy - (p.merange << 2), maxx = mvpx + (p.merange << 2), maxy = mvpy + (p.merange << 2);
    if (!p.rawRange)
    {
        {
            // CUData::clipMv (cudata.cpp:1915-1928)
            const int xmax = (p.picW + 8 - cuPelX - 1) << 2, xmin = -((C + 8 + cuPelX - 1) << 2);
            const int ymax = (p.picH + 8 - cuPelY - 1) << 2, ymin = -((C + 8 + cuPelY - 1) << 2);
            minx = min(xmax, max(xmin, minx)); miny = min(ymax, max(ymin, miny));
            maxx = min(xmax, max(xmin, maxx)); maxy = min(ymax, max(ymin, maxy));
        }
        if (p.sliceBounds)          // (m_param->maxSlices > 1) & m_bFrameParallel
        {
            miny = max(miny, p.sliceBounds[2 * ctuY]);
            maxy = min(maxy, p.sliceBounds[2 * ctuY + 1]);
        }
        const int maxMvLen = (1 << 15) - 1;
        minx = max(minx, -maxMvLen); miny = max(miny, -maxMvLen); maxx = min(maxx, maxMvLen); maxy = min(maxy, maxMvLen);
    }
    minx >>= 2; miny >>= 2; maxx >>= 2; maxy >>= 2;
    miny = min(miny, p.refLagPixels); maxy = min(maxy, p.refLagPixels);
    maxy = max(maxy, miny);

    MEState<pixel> s;
    s.isLowres = false; s.perThread = true; s.lane = 0; s.depth = p.depth;
    s.pred = nullptr; s.immed = nullptr;
    s.w = sw; s.h = sh;
    s.groupSize = G;
    s.groupMask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (laneInWarp & ~(G - 1)));
    s.partSizeScale = ((int)pu.h * (int)pu.h) >> 4;                       // motion.cpp:125-126 sizeScale = (H * H) >> 4
    s.cost = p.cost + 2 * 32768; s.costS = st.costS; s.costK = p.costK;
    s.mvpx = mvpx; s.mvpy = mvpy;
    s.stride = p.winPitch; s.gstride = p.refStride;
    s.fenc = const_cast<pixel*>(st.fenc) + (pu.y + sy) * 64 + pu.x + sx;
    s.fref = st.window + (int64_t)(PY + sy - st.winTop) * p.winPitch + (PX + sx - st.winLeft);
    s.gfref = (const pixel*)p.refY[ref] + (PX + sx) + (int64_t)(PY + sy) * p.refStride;
    s.winX0 = st.winLeft - PX; s.winX1 = st.winLeft + p.winPitch - (int)pu.w - PX;
    s.winY0 = st.winTop - PY;  s.winY1 = st.winTop + p.winRows - (int)pu.h - PY;
    s.chromaSatd = p.chromaSatd && pu.chromaOk;
    s.hshift = p.hshift; s.vshift = p.vshift; s.csize = 64;
    s.strideC = p.cwinPitch; s.gstrideC = p.refStrideC;
    if (s.chromaSatd)
    {
        const int cx = (PX + sx) >> p.hshift, cy = (PY + sy) >> p.vshift;
        const int fx = (pu.x + sx) >> p.hshift, fy = (pu.y + sy) >> p.vshift;
        s.fencC[0] = st.fencC[0] + fy * 64 + fx; s.fencC[1] = st.fencC[1] + fy * 64 + fx;
        s.frefC[0] = st.cwindow[0] + (int64_t)(cy - st.cwinTop) * p.cwinPitch + (cx - st.cwinLeft);
        s.frefC[1] = st.cwindow[1] + (int64_t)(cy - st.cwinTop) * p.cwinPitch + (cx - st.cwinLeft);
        s.gfrefC[0] = (const pixel*)p.refCb[ref] + cx + (int64_t)cy * p.refStrideC;
        s.gfrefC[1] = (const pixel*)p.refCr[ref] + cx + (int64_t)cy * p.refStrideC;
        // whole chroma samples (mv >> 3) for which the PU's chroma block plus the 4-tap footprint (1 left / above, 2 right / below)
        // is inside the chroma window
        const int pcx = PX >> p.hshift, pcy = PY >> p.vshift, wC = pu.w >> p.hshift, hC = pu.h >> p.vshift;
        s.cwinX0 = st.cwinLeft - pcx + 1; s.cwinX1 = st.cwinLeft + p.cwinPitch - wC - 2 - pcx;
        s.cwinY0 = st.cwinTop - pcy + 1;  s.cwinY1 = st.cwinTop + p.cwinRows - hC - 2 - pcy;
    }
    else
    {
        s.fencC[0] = s.fencC[1] = nullptr; s.frefC[0] = s.frefC[1] = nullptr; s.gfrefC[0] = s.gfrefC[1] = nullptr;
        s.cwinX0 = s.cwinY0 = 0; s.cwinX1 = s.cwinY1 = -1;
    }
    s.integral = nullptr; s.integralOff = 0;
    s.lowres[0] = s.lowres[1] = s.lowres[2] = s.lowres[3] = nullptr;

    int numCand = 0;
    const int* mvc = nullptr;
    if (p.numCandPu && p.maxCand > 0)
    {
        numCand = min((int)p.numCandPu[puSlot], p.maxCand);
        mvc = p.mvcPu + puSlot * p.maxCand * 2;
    }
    int ox, oy;
    const int cost = motion_estimate<pixel>(s, mv2(minx, miny), mv2(maxx, maxy), mv2(mvpx, mvpy), numCand, mvc, p.merange, p.searchMethod,
                                            p.subpelRefine, p.maxSlices, pu.w == 64 && pu.h == 64, ox, oy);
    if (q == 0) { o[0] = ox; o[1] = oy; o[2] = cost; }
}

} // namespace x265b200

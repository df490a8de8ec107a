// me_ctu_emu.cpp -- TEST INFRASTRUCTURE: one (CTU, reference) pair of me_ctu_kernel (csrc/me_ctu_kernels.cu) on the host.
// The per-lane body of the kernel (csrc/me_ctu_device.cuh -> me_device.cuh) and the layout / geometry code
// (csrc/me_ctu_layout.h) are compiled unchanged through tests/host_emu/me_host_emu.h; this file restates only the kernel's
// thin CTA code: stage the windows, the source CTU and the cost entries in (emulated) shared memory the way the TMA loads
// do, then run every work item with the lanes of a PU as host threads.
#define ME_HOST_EMU 1
#include "me_host_emu.h"
#include "me_ctu_device.cuh"
#include <vector>
#include <climits>

namespace x265b200 {
namespace emu {
unsigned char* smem_base = nullptr;
thread_local Group* t_group = nullptr;
thread_local int t_q = 0;
}

struct EmuCtuParams            // mirrors x265b200_me_frame_params (plus the per-call pointers the kernel gets)
{
    int32_t depth, ctuSize, minCuSize, rect, amp, picWidth, picHeight, ctuCols, ctuRows, marginX, marginY, rowsTotal, numRefs;
    int32_t searchMethod, subpelRefine, merange, csp, maxCand, maxSlices, refLagPixels;
};

template<typename pixel>
struct LaneJob
{
    const MECtuArgs* p; const MECtuStage<pixel>* st; uint32_t word; int ctuX, ctuY, ref, lane;
    emu::Group* group; int q;
};
template<typename pixel>
static void* lane_main(void* v)
{
    LaneJob<pixel>* j = (LaneJob<pixel>*)v;
    emu::t_group = j->group; emu::t_q = j->q;
    me_ctu_lane<pixel>(*j->p, *j->st, j->word, j->ctuX, j->ctuY, j->ref, j->lane);
    return nullptr;
}

// copy a box the way a TMA tile load does (rows of `rowBytes` bytes; the planes of the tests are padded generously)
static void stage_box(unsigned char* dst, const unsigned char* planeBase, int64_t strideBytes, int64_t xBytes, int y, int rowBytes, int rows)
{
    for (int r = 0; r < rows; r++) memcpy(dst + (size_t)r * rowBytes, planeBase + (int64_t)(y + r) * strideBytes + xBytes, rowBytes);
}

template<typename pixel>
static int run(const EmuCtuParams& P, const pixel* curY, const pixel* curCb, const pixel* curCr, int64_t curStride, int64_t curStrideC,
               const pixel* refY, const pixel* refCb, const pixel* refCr, int64_t refStride, int64_t refStrideC,
               int ctuX, int ctuY, const int32_t* mvpCtu, const int32_t* mvpPu, const uint8_t* numCand, const int32_t* mvc,
               const int32_t* sliceBounds, const uint16_t* costTable, int32_t* out, int32_t* numPuOut)
{
    const int px = (int)sizeof(pixel), APX = 16 / px, C = P.ctuSize;
    const int hs = (P.csp == 1 || P.csp == 2) ? 1 : 0, vs = P.csp == 1 ? 1 : 0;
    const bool chromaSatd = P.csp != 0 && P.subpelRefine > 2;
    MECtuGeom g;
    if (me_ctu_geometry(P.depth, C, P.merange, P.csp, chromaSatd, g)) return -3;
    MECtuLayout L;
    me_ctu_build_layout(C, P.minCuSize, P.rect != 0, P.amp != 0, P.csp, chromaSatd, L);
    *numPuOut = (int)L.pus.size();

    unsigned char* smem = (unsigned char*)aligned_alloc(128, (g.smemBytes + 255) & ~(size_t)127);
    if (!smem) return -1;
    memset(smem, 0xA5, g.smemBytes);
    emu::smem_base = smem;

    const int cmarginX = P.marginX >> hs, cmarginY = P.marginY >> vs;
    // the kernel sees ONE (ctu, ref): present it as a 1-reference, whole-frame call restricted to that CTU
    const void* refYp[1] = { refY }; const void* refCbp[1] = { refCb }; const void* refCrp[1] = { refCr };
    MECtuArgs a; memset(&a, 0, sizeof(a));
    a.refY = refYp; a.refCb = refCbp; a.refCr = refCrp; a.refStride = refStride; a.refStrideC = refStrideC;
    a.ctuCols = P.ctuCols; a.ctuRows = P.ctuRows; a.numRefs = 1; a.ctuSize = C;
    a.marginX = P.marginX; a.marginY = P.marginY; a.cmarginX = cmarginX; a.cmarginY = cmarginY;
    a.picW = P.picWidth > 0 ? P.picWidth : P.ctuCols * C; a.picH = P.picHeight > 0 ? P.picHeight : P.ctuRows * C; a.firstCtuRow = 0;
    a.pus = L.pus.data(); a.numPu = (int)L.pus.size(); a.items = L.items.data(); a.numItems = L.numItems;
    // per-(ctu) arrays are addressed by the kernel as [ref][ctu][...]: shift the bases so that index (0, ctu) hits the caller's arrays
    const int ctu = ctuY * P.ctuCols + ctuX;
    const int64_t slot0 = (int64_t)ctu * a.numPu;
    a.mvpCtu = mvpCtu ? mvpCtu - (int64_t)ctu * 2 : nullptr;
    a.mvpPu = mvpPu ? mvpPu - slot0 * 2 : nullptr;
    a.numCandPu = (numCand && P.maxCand) ? numCand - slot0 : nullptr;
    a.mvcPu = mvc ? mvc - slot0 * P.maxCand * 2 : nullptr; a.maxCand = P.maxCand;
    a.out = out - slot0 * 3;
    a.cost = costTable; a.costK = g.costK;
    a.searchMethod = P.searchMethod; a.subpelRefine = P.subpelRefine; a.merange = P.merange; a.depth = P.depth; a.R = g.R;
    a.winPitch = g.winPitch / px; a.winRows = g.winRows;
    a.csp = P.csp; a.hshift = hs; a.vshift = vs; a.chromaSatd = chromaSatd;
    a.cwinPitch = g.cwinPitch / px; a.cwinRows = g.cwinRows;
    a.sliceBounds = sliceBounds; a.maxSlices = P.maxSlices > 1 ? P.maxSlices : 1;
    a.refLagPixels = P.refLagPixels > 0 ? P.refLagPixels : INT_MAX / 2;

    int mvpx = 0, mvpy = 0;
    if (mvpCtu) { mvpx = mvpCtu[0]; mvpy = mvpCtu[1]; }
    const int cx = mvpx >> 2, cy = mvpy >> 2;
    MECtuStage<pixel> st;
    const int tx = ctuX * C + cx - g.R + P.marginX, ax = tx & ~(APX - 1);
    st.window = (const pixel*)smem; st.winLeft = ax - P.marginX; st.winTop = ctuY * C + cy - g.R;
    st.fenc = (const pixel*)(smem + g.offFenc);
    {
        // cost[-K .. K]
        uint16_t* cs = (uint16_t*)(smem + g.offCost);
        for (int i = 0; i < 2 * g.costK + 1; i++) cs[i] = costTable[2 * 32768 - g.costK + i];
        st.costS = cs + g.costK;
    }
    const unsigned char* baseY = (const unsigned char*)refY - ((int64_t)P.marginY * refStride + P.marginX) * px;
    stage_box(smem, baseY, refStride * px, (int64_t)ax * px, st.winTop + P.marginY, g.winPitch, g.winRows);
    const unsigned char* baseCur = (const unsigned char*)curY - ((int64_t)P.marginY * curStride + P.marginX) * px;
    stage_box(smem + g.offFenc, baseCur, curStride * px, (int64_t)(ctuX * C + P.marginX) * px, ctuY * C + P.marginY, 64 * px, g.fencRows);
    if (chromaSatd)
    {
        const int ctx0 = ((ctuX * C + cx - g.R) >> hs) - 2 + cmarginX, cax = ctx0 & ~(APX - 1);
        st.cwinLeft = cax - cmarginX; st.cwinTop = ((ctuY * C + cy - g.R) >> vs) - 2;
        const pixel* rc[2] = { refCb, refCr }; const pixel* cc[2] = { curCb, curCr };
        for (int c = 0; c < 2; c++)
        {
            st.cwindow[c] = (const pixel*)(smem + g.offCwin[c]); st.fencC[c] = (const pixel*)(smem + g.offFencC[c]);
            const unsigned char* b = (const unsigned char*)rc[c] - ((int64_t)cmarginY * refStrideC + cmarginX) * px;
            stage_box(smem + g.offCwin[c], b, refStrideC * px, (int64_t)cax * px, st.cwinTop + cmarginY, g.cwinPitch, g.cwinRows);
            const unsigned char* bc = (const unsigned char*)cc[c] - ((int64_t)cmarginY * curStrideC + cmarginX) * px;
            stage_box(smem + g.offFencC[c], bc, curStrideC * px, (int64_t)(((ctuX * C) >> hs) + cmarginX) * px, ((ctuY * C) >> vs) + cmarginY, 64 * px, g.fencCRows);
        }
    }
    else
    {
        st.cwinLeft = st.cwinTop = 0; st.cwindow[0] = st.cwindow[1] = nullptr; st.fencC[0] = st.fencC[1] = nullptr;
    }

    for (int it = 0; it < L.numItems; it++)
    {
        const uint32_t* words = &L.items[(size_t)it * 32];
        int lane = 0;
        while (lane < 32)
        {
            const uint32_t w = words[lane];
            if (!(w & 0x80000000u)) { lane++; continue; }
            const int G = 1 << ((w >> 24) & 7);
            emu::Group grp; grp.n = G;
            pthread_barrier_init(&grp.bar, nullptr, G);
            std::vector<LaneJob<pixel>> jobs(G);
            std::vector<pthread_t> th(G);
            for (int q = 0; q < G; q++)
            {
                jobs[q] = LaneJob<pixel>{ &a, &st, words[lane + q], ctuX, ctuY, 0, lane + q, &grp, q };
                if (G == 1) lane_main<pixel>(&jobs[q]);
                else pthread_create(&th[q], nullptr, lane_main<pixel>, &jobs[q]);
            }
            if (G > 1) for (int q = 0; q < G; q++) pthread_join(th[q], nullptr);
            pthread_barrier_destroy(&grp.bar);
            lane += G;
        }
    }
    emu::smem_base = nullptr;
    free(smem);
    return 0;
}
} // namespace x265b200

extern "C" int emu_me_ctu(const x265b200::EmuCtuParams* P, const void* curY, const void* curCb, const void* curCr, int64_t curStride, int64_t curStrideC,
                          const void* refY, const void* refCb, const void* refCr, int64_t refStride, int64_t refStrideC,
                          int ctuX, int ctuY, const int32_t* mvpCtu, const int32_t* mvpPu, const uint8_t* numCand, const int32_t* mvc,
                          const int32_t* sliceBounds, const uint16_t* costTable, int32_t* out, int32_t* numPuOut)
{
    using namespace x265b200;
    if (P->depth > 8)
        return run<uint16_t>(*P, (const uint16_t*)curY, (const uint16_t*)curCb, (const uint16_t*)curCr, curStride, curStrideC,
                             (const uint16_t*)refY, (const uint16_t*)refCb, (const uint16_t*)refCr, refStride, refStrideC,
                             ctuX, ctuY, mvpCtu, mvpPu, numCand, mvc, sliceBounds, costTable, out, numPuOut);
    return run<uint8_t>(*P, (const uint8_t*)curY, (const uint8_t*)curCb, (const uint8_t*)curCr, curStride, curStrideC,
                        (const uint8_t*)refY, (const uint8_t*)refCb, (const uint8_t*)refCr, refStride, refStrideC,
                        ctuX, ctuY, mvpCtu, mvpPu, numCand, mvc, sliceBounds, costTable, out, numPuOut);
}

extern "C" int emu_me_ctu_layout(int ctuSize, int minCu, int rect, int amp, int32_t* outXYWH, int cap)
{
    std::vector<x265b200::MECtuPU> v;
    x265b200::me_ctu_build_pus(ctuSize, minCu, rect != 0, amp != 0, 0, v);
    for (int i = 0; i < (int)v.size() && i < cap; i++) { outXYWH[4 * i] = v[i].x; outXYWH[4 * i + 1] = v[i].y; outXYWH[4 * i + 2] = v[i].w; outXYWH[4 * i + 3] = v[i].h; }
    return (int)v.size();
}

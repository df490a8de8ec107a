// la_search_emu.cpp -- TEST INFRASTRUCTURE: the lookahead list search (csrc/la_search_thread.cu = estimateCUCost phase 1,
// slicetype.cpp:3216-3325) on the host.  The device source of the search (csrc/me_device.cuh, lowres thread-only build) is
// compiled unchanged through me_host_emu.h; this file restates only the kernel's thin per-CU role code and walks the CUs in
// reverse raster order (the order the wavefront of the kernel is equivalent to).
#define ME_HOST_EMU 1
#define ME_FORCE_THREAD 1
#define ME_LOWRES_ONLY 1
#define ME_BATCH_GROUPSUM 1          /* the candidate SADs of a search step through the register-only call (me_device.cuh thread_cand_sads) */
#ifndef EMU_LA_PACKED_SATD_OFF
#define ME_PACKED_SATD 1
#endif
#include "me_device.cuh"
#include <vector>

namespace x265b200 {
namespace emu {
unsigned char* smem_base = nullptr;
thread_local Group* t_group = nullptr;
thread_local int t_q = 0;
}

#ifndef EMU_LA_LANES
#define EMU_LA_LANES 2            /* lanes per CU, as LA_LANES of csrc/la_search_thread.cu: each owns 8 / lanes rows of the CU */
#endif

template<typename pixel>
struct CuLane
{
    MEState<pixel> s; emu::Group* group; int q;
    const int (*mvc)[2]; int numc, bBidir, merange, cuX, cuY, W, Hc;
    int ox, oy, cost;
};

template<typename pixel>
static void* cu_lane_main(void* v)
{
    CuLane<pixel>& a = *(CuLane<pixel>*)v;
    emu::t_group = a.group; emu::t_q = a.q;
    MEState<pixel>& s = a.s;
    int mvpx = 0, mvpy = 0, skipCost = 0x7fffffff;
    if (a.numc)
    {
        int mvpcost = ME_COST_MAX;
        for (int idx = 0; idx < a.numc; idx++)
        {
            int cost = lowres_qpel_cost<pixel>(s, a.mvc[idx][0], a.mvc[idx][1], true);
            if (cost < mvpcost) { mvpcost = cost; mvpx = a.mvc[idx][0]; mvpy = a.mvc[idx][1]; }
            if (!(mvpx | mvpy) && a.bBidir) skipCost = cost;
        }
    }
    s.mvpx = mvpx; s.mvpy = mvpy;
    const MV2 mvmin = mv2(-a.cuX * 8 - 8, -a.cuY * 8 - 8), mvmax = mv2((a.W - a.cuX - 1) * 8 + 8, (a.Hc - a.cuY - 1) * 8 + 8);
    int ox, oy;
    int fencCost = motion_estimate<pixel>(s, mvmin, mvmax, mv2(mvpx, mvpy), 0, nullptr, a.merange, (int)ME_HEX, 1, 1, 0, ox, oy);
    if (skipCost < 64 && skipCost < fencCost && a.bBidir) { fencCost = skipCost; ox = 0; oy = 0; }
    a.ox = ox; a.oy = oy; a.cost = fencCost;
    return nullptr;
}

template<typename pixel>
static int run_field(int depth, const pixel* const fencPlanes[4], const pixel* const refPlanes[4], int64_t stride, int W, int Hc, int bBidir,
                     int merange, const uint16_t* costTable, int32_t* mvs, int32_t* mvcosts)
{
    constexpr int L = EMU_LA_LANES, HR = 8 / L;
    std::vector<pixel> fencBuf(8 * 64);
    for (int cuY = Hc - 1; cuY >= 0; cuY--)
        for (int cuX = W - 1; cuX >= 0; cuX--)
        {
            const int cuXY = cuX + cuY * W;
            const bool lastRow = cuY == Hc - 1;
            const int64_t pelOffset = 8 * cuX + (int64_t)8 * cuY * stride;
            for (int y = 0; y < 8; y++) memcpy(fencBuf.data() + y * 64, fencPlanes[0] + pelOffset + (int64_t)y * stride, 8 * sizeof(pixel));
            int mvc[5][2], numc = 0;
            if (cuX < W - 1) { mvc[numc][0] = mvs[(cuXY + 1) * 2]; mvc[numc][1] = mvs[(cuXY + 1) * 2 + 1]; numc++; }
            if (!lastRow)
            {
                const int32_t* row = mvs + (int64_t)(cuXY + W) * 2;
                mvc[numc][0] = row[0]; mvc[numc][1] = row[1]; numc++;
                if (cuX > 0) { mvc[numc][0] = row[-2]; mvc[numc][1] = row[-1]; numc++; }
                if (cuX < W - 1) { mvc[numc][0] = row[2]; mvc[numc][1] = row[3]; numc++; }
            }
            emu::Group g; g.n = L;
            pthread_barrier_init(&g.bar, nullptr, L);
            std::vector<CuLane<pixel>> la(L);
            std::vector<pthread_t> th(L);
            for (int q = 0; q < L; q++)
            {
                CuLane<pixel>& a = la[q];
                MEState<pixel>& s = a.s;
                memset(&s, 0, sizeof(s));
                s.fenc = fencBuf.data() + q * HR * 64;
                s.stride = stride; s.isLowres = true; s.perThread = true; s.chromaSatd = false; s.groupSize = L; s.groupMask = 0xffffffffu;
                s.w = 8; s.h = HR; s.lane = 0; s.depth = depth; s.partSizeScale = 4; s.cost = costTable + 2 * 32768;
                for (int k = 0; k < 4; k++) s.lowres[k] = refPlanes[k] + pelOffset + (int64_t)q * HR * stride;
                s.fref = s.lowres[0]; s.gfref = s.lowres[0]; s.gstride = stride;
                a.group = &g; a.q = q; a.mvc = mvc; a.numc = numc; a.bBidir = bBidir; a.merange = merange; a.cuX = cuX; a.cuY = cuY; a.W = W; a.Hc = Hc;
                if (L == 1) cu_lane_main<pixel>(&a);
                else pthread_create(&th[q], nullptr, cu_lane_main<pixel>, &a);
            }
            if (L > 1) for (int q = 0; q < L; q++) pthread_join(th[q], nullptr);
            pthread_barrier_destroy(&g.bar);
            for (int q = 1; q < L; q++)
                if (la[q].ox != la[0].ox || la[q].oy != la[0].oy || la[q].cost != la[0].cost) return -2;        // the lanes of a CU must agree
            mvs[cuXY * 2] = la[0].ox; mvs[cuXY * 2 + 1] = la[0].oy; mvcosts[cuXY] = la[0].cost;
        }
    return 0;
}
} // namespace x265b200

extern "C" int emu_la_search_field(int depth, const void* const fencPlanes[4], const void* const refPlanes[4], int64_t stride, int widthInCU, int heightInCU,
                                   int bBidir, int merange, const uint16_t* costTable, int32_t* mvs, int32_t* mvcosts)
{
    using namespace x265b200;
    if (depth > 8) return run_field<uint16_t>(depth, (const uint16_t* const*)fencPlanes, (const uint16_t* const*)refPlanes, stride, widthInCU, heightInCU, bBidir, merange, costTable, mvs, mvcosts);
    return run_field<uint8_t>(depth, (const uint8_t* const*)fencPlanes, (const uint8_t* const*)refPlanes, stride, widthInCU, heightInCU, bBidir, merange, costTable, mvs, mvcosts);
}

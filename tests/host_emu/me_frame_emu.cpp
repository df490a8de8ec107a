// me_frame_emu.cpp -- TEST INFRASTRUCTURE: one (CTU, reference) pair of me_frame_kernel (csrc/me_frame_kernels.cu) on the host.
// The device source of the search (csrc/me_device.cuh) is compiled unchanged through tests/host_emu/me_host_emu.h; this file
// restates only the kernel's thin role code: stage the window and the source CTU in (emulated) shared memory, then run the
// 85 2Nx2N PU searches with the kernel's lane layout -- a lane owns an 8-wide sub-block, the lanes of a PU are host threads.
#define ME_HOST_EMU 1
#define ME_FORCE_THREAD 1
#define ME_FULLRES_ONLY 1
#define ME_REF_IN_SMEM 1
#ifndef EMU_WINDOW_SLOW
#define ME_WINDOW_FAST 1
#endif
#ifndef EMU_SUBPEL_PACKED_OFF
#define ME_SUBPEL_PACKED 1
#endif
#ifndef EMU_BATCH_GROUPSUM_OFF
#define ME_BATCH_GROUPSUM 1
#endif
#ifndef EMU_VCELL_REUSE_OFF
#define ME_VCELL_REUSE 1
#endif
#include "me_device.cuh"
#include <vector>

namespace x265b200 {
namespace emu {
unsigned char* smem_base = nullptr;
thread_local Group* t_group = nullptr;
thread_local int t_q = 0;
}

template<typename pixel>
struct LaneArgs
{
    MEState<pixel> s;
    emu::Group* group; int q;
    int cx, cy, mvpx, mvpy, merange, method, subme, is64;
    int ox, oy, cost;
};

template<typename pixel>
static void* lane_main(void* p)
{
    LaneArgs<pixel>* a = (LaneArgs<pixel>*)p;
    emu::t_group = a->group; emu::t_q = a->q;
    a->cost = motion_estimate<pixel>(a->s, mv2(a->cx - a->merange, a->cy - a->merange), mv2(a->cx + a->merange, a->cy + a->merange),
                                     mv2(a->mvpx, a->mvpy), 0, nullptr, a->merange, a->method, a->subme, 1, a->is64, a->ox, a->oy);
    return nullptr;
}

// out: [85][3] = {mvx, mvy, cost} in level order (64x64; 32x32 raster; 16x16 raster; 8x8 raster)
template<typename pixel>
static int run_ctu(int depth, const pixel* curOrigin, int64_t curStride, const pixel* refOrigin, int64_t refStride, int marginX,
                   int ctuX, int ctuY, int mvpx, int mvpy, int method, int subme, int merange, const uint16_t* costTable, int32_t* out)
{
    const int px = (int)sizeof(pixel), R = merange + 8;
    int winW = 64 + 2 * R + (16 / px - 1);
    int pitchBytes = ((winW * px + 15) / 16) * 16;
    winW = pitchBytes / px;
    const int winH = 64 + 2 * R;
    const size_t winBytes = ((size_t)winW * winH * px + 127) & ~(size_t)127;
    unsigned char* smem = (unsigned char*)aligned_alloc(128, winBytes + 64 * 64 * px + 128);
    if (!smem) return -1;
    memset(smem, 0xA5, winBytes + 64 * 64 * px + 128);
    emu::smem_base = smem;
    pixel* window = (pixel*)smem;
    pixel* fencCtu = (pixel*)(smem + winBytes);
    const int cx = mvpx >> 2, cy = mvpy >> 2;
    const int wx0 = ctuX * 64 + cx - R, wy0 = ctuY * 64 + cy - R;
    const int tx = wx0 + marginX, ax = tx & ~(16 / px - 1), ex = tx - ax;          // the kernel's 16-byte aligned TMA box start
    for (int y = 0; y < winH; y++)
        memcpy(window + (size_t)y * winW, refOrigin + (int64_t)(wy0 + y) * refStride + (ax - marginX), (size_t)winW * px);
    for (int y = 0; y < 64; y++)
        memcpy(fencCtu + y * 64, curOrigin + (int64_t)(ctuY * 64 + y) * curStride + ctuX * 64, 64 * px);

    int o = 0;
    for (int level = 0; level < 4; level++)
    {
        const int lanesLog2 = level == 0 ? 5 : 6 - 2 * level, gs = 1 << lanesLog2;
        const int subH = level == 0 ? 16 : 8, subCols = level == 0 ? 8 : (8 >> level);
        const int sz = 64 >> level, per = 1 << level;
        for (int idx = 0; idx < per * per; idx++, o++)
        {
            const int puy = (idx / per) * sz, pux = (idx % per) * sz;
            emu::Group g; g.n = gs;
            pthread_barrier_init(&g.bar, nullptr, gs);
            std::vector<LaneArgs<pixel>> la(gs);
            std::vector<pthread_t> th(gs);
            for (int q = 0; q < gs; q++)
            {
                LaneArgs<pixel>& a = la[q];
                MEState<pixel>& s = a.s;
                memset(&s, 0, sizeof(s));
                const int sx = (q % subCols) * 8, sy = (q / subCols) * subH;
                s.stride = winW; s.isLowres = false; s.perThread = true; s.chromaSatd = false; s.lane = 0; s.depth = depth;
                s.cost = costTable + 2 * 32768; s.mvpx = mvpx; s.mvpy = mvpy; s.gstride = refStride;
                s.groupSize = gs; s.groupMask = 0xffffffffu; s.pred = nullptr; s.immed = nullptr; s.w = 8; s.h = subH;
                s.partSizeScale = (sz * sz) >> 4;
                s.fenc = fencCtu + (puy + sy) * 64 + pux + sx;
                s.fref = window + (int64_t)(puy + sy - cy + R) * winW + (pux + sx - cx + R + ex);
                s.gfref = refOrigin + (ctuX * 64 + pux + sx) + (int64_t)(ctuY * 64 + puy + sy) * refStride;
                a.group = &g; a.q = q; a.cx = cx; a.cy = cy; a.mvpx = mvpx; a.mvpy = mvpy; a.merange = merange; a.method = method;
                a.subme = subme; a.is64 = sz == 64;
                if (gs == 1) lane_main<pixel>(&a);
                else pthread_create(&th[q], nullptr, lane_main<pixel>, &a);
            }
            if (gs > 1) for (int q = 0; q < gs; q++) pthread_join(th[q], nullptr);
            pthread_barrier_destroy(&g.bar);
            for (int q = 1; q < gs; q++)
                if (la[q].ox != la[0].ox || la[q].oy != la[0].oy || la[q].cost != la[0].cost) { free(smem); return -2; }    // lanes must agree
            out[3 * o] = la[0].ox; out[3 * o + 1] = la[0].oy; out[3 * o + 2] = la[0].cost;
        }
    }
    emu::smem_base = nullptr;
    free(smem);
    return 0;
}
} // namespace x265b200

extern "C" int emu_me_frame_ctu(int depth, const void* curOrigin, int64_t curStride, const void* refOrigin, int64_t refStride, int marginX,
                                int ctuX, int ctuY, int mvpx, int mvpy, int method, int subme, int merange, const uint16_t* costTable, int32_t* out)
{
    using namespace x265b200;
    if (depth > 8)
        return run_ctu<uint16_t>(depth, (const uint16_t*)curOrigin, curStride, (const uint16_t*)refOrigin, refStride, marginX, ctuX, ctuY, mvpx, mvpy,
                                 method, subme, merange, costTable, out);
    return run_ctu<uint8_t>(depth, (const uint8_t*)curOrigin, curStride, (const uint8_t*)refOrigin, refStride, marginX, ctuX, ctuY, mvpx, mvpy,
                            method, subme, merange, costTable, out);
}

// me_ctu_layout.h -- host-side description of the PUs one CTU searches and of how they are spread over warp lanes
// (plain C++, no CUDA: shared by me_ctu_kernels.cu and by the host emulation in tests/host_emu/me_ctu_emu.cpp).
//
// Which PUs exist follows the reference's partition modes (common/cudata.h PartSize, analysis.cpp: 2Nx2N at every CU
// size from the CTU down to minCUSize; 2NxN / Nx2N with --rect; 2NxnU / 2NxnD / nLx2N / nRx2N with --amp for CUs of 16 and
// larger, param.cpp:396-555 per preset).  The ORDER is ours and is what the C ABI documents (x265b200_me_frame_layout):
// CU sizes from the CTU size down, CUs of one size in raster order, per CU the part modes in PartSize order, per mode its PUs
// in partition order.
#pragma once
#include <stdint.h>
#include <vector>
#include <algorithm>

namespace x265b200 {

struct MECtuPU            // one PU of the CTU layout
{
    uint8_t x, y, w, h;   // position inside the CTU and size, luma samples
    uint8_t cuX, cuY;     // the CU that owns it (CUData::clipMv uses the CU position, cudata.cpp:1915-1928)
    uint8_t cuSize;
    uint8_t chromaOk;     // chroma[csp].pu[].satd exists for this shape (pixel.cpp:1200-1226, :1279-1305): both chroma
                          // dimensions are multiples of 4
};

// one lane of a work item (32 lanes = one warp pass): which PU, which sub-block of it
//   bits 0..9 PU index | 10..13 sx/4 | 14..17 sy/4 | 18 sw == 8 (else 4) | 19..23 sh/4 (0 = idle lane of the group)
//   | 24..26 log2(lanes of the PU's group) | 31 valid
inline uint32_t me_ctu_lane_word(int pu, int sx, int sy, int sw, int sh, int gLog2)
{
    return (uint32_t)pu | ((uint32_t)(sx >> 2) << 10) | ((uint32_t)(sy >> 2) << 14) | ((uint32_t)(sw == 8) << 18) |
           ((uint32_t)(sh >> 2) << 19) | ((uint32_t)gLog2 << 24) | 0x80000000u;
}

struct MECtuLayout
{
    std::vector<MECtuPU> pus;
    std::vector<uint32_t> items;      // numItems x 32 lane words, heaviest items first
    int numItems = 0;
};

inline void me_ctu_add_pu(std::vector<MECtuPU>& v, int cuX, int cuY, int S, int x, int y, int w, int h, int hshift, int vshift, bool chroma)
{
    MECtuPU p;
    p.x = (uint8_t)(cuX + x); p.y = (uint8_t)(cuY + y); p.w = (uint8_t)w; p.h = (uint8_t)h;
    p.cuX = (uint8_t)cuX; p.cuY = (uint8_t)cuY; p.cuSize = (uint8_t)S;
    p.chromaOk = chroma && !((w >> hshift) & 3) && !((h >> vshift) & 3);
    v.push_back(p);
}

// csp: 0 = luma only, 1 = 4:2:0, 2 = 4:2:2, 3 = 4:4:4 (x265.h:588-592)
// sizeMask: bit k set = CUs of size 64 >> k are searched (0 = all sizes from ctuSize down to minCu)
inline void me_ctu_build_pus(int ctuSize, int minCu, bool rect, bool amp, int csp, std::vector<MECtuPU>& v, int sizeMask = 0)
{
    const int hs = (csp == 1 || csp == 2) ? 1 : 0, vs = csp == 1 ? 1 : 0;
    const bool ch = csp != 0;
    v.clear();
    for (int S = ctuSize; S >= minCu; S >>= 1)
    {
        if (sizeMask && !(sizeMask & (64 / S))) continue;
        for (int cy = 0; cy < ctuSize; cy += S)
            for (int cx = 0; cx < ctuSize; cx += S)
            {
                me_ctu_add_pu(v, cx, cy, S, 0, 0, S, S, hs, vs, ch);                                  // SIZE_2Nx2N
                if (rect)
                {
                    me_ctu_add_pu(v, cx, cy, S, 0, 0, S, S / 2, hs, vs, ch);                          // SIZE_2NxN
                    me_ctu_add_pu(v, cx, cy, S, 0, S / 2, S, S / 2, hs, vs, ch);
                    me_ctu_add_pu(v, cx, cy, S, 0, 0, S / 2, S, hs, vs, ch);                          // SIZE_Nx2N
                    me_ctu_add_pu(v, cx, cy, S, S / 2, 0, S / 2, S, hs, vs, ch);
                }
                if (amp && S >= 16)
                {
                    me_ctu_add_pu(v, cx, cy, S, 0, 0, S, S / 4, hs, vs, ch);                          // SIZE_2NxnU
                    me_ctu_add_pu(v, cx, cy, S, 0, S / 4, S, 3 * S / 4, hs, vs, ch);
                    me_ctu_add_pu(v, cx, cy, S, 0, 0, S, 3 * S / 4, hs, vs, ch);                      // SIZE_2NxnD
                    me_ctu_add_pu(v, cx, cy, S, 0, 3 * S / 4, S, S / 4, hs, vs, ch);
                    me_ctu_add_pu(v, cx, cy, S, 0, 0, S / 4, S, hs, vs, ch);                          // SIZE_nLx2N
                    me_ctu_add_pu(v, cx, cy, S, S / 4, 0, 3 * S / 4, S, hs, vs, ch);
                    me_ctu_add_pu(v, cx, cy, S, 0, 0, 3 * S / 4, S, hs, vs, ch);                      // SIZE_nRx2N
                    me_ctu_add_pu(v, cx, cy, S, 3 * S / 4, 0, S / 4, S, hs, vs, ch);
                }
            }
    }
}

// Lanes of one PU.  A lane owns a sub-block 8 (or, for the 4-wide remainder of 4- and 12-wide PUs, 4) pixels wide and a
// multiple of 4 rows tall -- of 4 << vshift rows when the PU carries the chroma SATD term, so that the lane's chroma
// sub-block is whole 4x4 cells too.  SAD and SATD are additive over 4x4 cells, so the lanes of a PU add their partial costs
// (xor butterfly over a power-of-two group; lanes the shape cannot feed stay idle with a 0-row sub-block).
struct MECtuSub { int sx, sy, sw, sh; };
inline int me_ctu_split_pu(const MECtuPU& pu, bool chromaSatd, int vshift, MECtuSub out[32])
{
    const int w = pu.w, h = pu.h;
    // columns are cut at multiples of 8 of the CTU's x axis, so that an 8-wide sub-block is 8-pixel aligned in the cached
    // source CTU (its rows are read with one vector load): the 12-wide AMP piece at x = 4 is 4 + 8, the one at x = 0 is 8 + 4
    int colX[9], colW[9], cols = 0;
    for (int x = 0; x < w;)
    {
        const int ax = pu.x + x;
        const int cw = ((ax & 7) || w - x < 8) ? 4 : 8;
        colX[cols] = x; colW[cols] = cw; cols++;
        x += cw;
    }
    int unit = (chromaSatd && pu.chromaOk) ? (4 << vshift) : 4;
    if (h % unit) unit = 4;
    const int units = h / unit;
    int G = 1;
    while (G < 32 && G * 2 <= (w * h) / 64) G *= 2;
    while (G < cols) G *= 2;
    int rgroups = G / cols;
    if (rgroups > units) rgroups = units;
    int n = 0, y = 0;
    for (int g = 0; g < rgroups; g++)
    {
        const int u = units / rgroups + (g < units % rgroups ? 1 : 0);
        for (int c = 0; c < cols; c++)
        {
            out[n].sx = colX[c]; out[n].sy = y; out[n].sw = colW[c]; out[n].sh = u * unit;
            n++;
        }
        y += u * unit;
    }
    for (; n < G; n++) { out[n].sx = 0; out[n].sy = 0; out[n].sw = 8; out[n].sh = 0; }
    return G;
}

inline void me_ctu_build_layout(int ctuSize, int minCu, bool rect, bool amp, int csp, bool chromaSatd, MECtuLayout& L, int sizeMask = 0)
{
    me_ctu_build_pus(ctuSize, minCu, rect, amp, csp, L.pus, sizeMask);
    const int vs = csp == 1 ? 1 : 0;
    struct Grp { int pu, G, work; MECtuSub sub[32]; };
    std::vector<Grp> groups(L.pus.size());
    for (size_t i = 0; i < L.pus.size(); i++)
    {
        Grp& g = groups[i];
        g.pu = (int)i;
        g.G = me_ctu_split_pu(L.pus[i], chromaSatd, vs, g.sub);
        g.work = 0;
        for (int k = 0; k < g.G; k++) g.work = std::max(g.work, g.sub[k].sw * g.sub[k].sh);
    }
    // groups of equal lane count share items; within a lane count, heavier sub-blocks first so items are homogeneous
    std::stable_sort(groups.begin(), groups.end(), [](const Grp& a, const Grp& b) { return a.G != b.G ? a.G > b.G : a.work > b.work; });
    struct Item { uint32_t w[32]; int work; };
    std::vector<Item> items;
    size_t i = 0;
    while (i < groups.size())
    {
        Item it; it.work = 0;
        for (int k = 0; k < 32; k++) it.w[k] = 0;
        const int G = groups[i].G;
        int lane = 0;
        while (i < groups.size() && groups[i].G == G && lane + G <= 32)
        {
            int gl = 0; while ((1 << gl) < G) gl++;
            for (int k = 0; k < G; k++)
                it.w[lane + k] = me_ctu_lane_word(groups[i].pu, groups[i].sub[k].sx, groups[i].sub[k].sy, groups[i].sub[k].sw, groups[i].sub[k].sh, gl);
            it.work = std::max(it.work, groups[i].work);
            lane += G; i++;
        }
        items.push_back(it);
    }
    std::stable_sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.work > b.work; });
    L.numItems = (int)items.size();
    L.items.resize(items.size() * 32);
    for (size_t k = 0; k < items.size(); k++)
        for (int l = 0; l < 32; l++) L.items[k * 32 + l] = items[k].w[l];
}

// Shared-memory geometry of one CTA of the search: the luma window (rows of winPitch bytes, staged in winBoxes TMA boxes of
// winBoxRows rows), the Cb / Cr windows, the source CTU (and its Cb / Cr) at a row pitch of 64 pixels, the reachable MV-cost
// entries, the mbarrier and the work-queue counter.
struct MECtuGeom
{
    int R, costK;
    int winPitch, winRows, winBoxes, winBoxRows;          // winPitch / cwinPitch in BYTES
    int cwinPitch, cwinRows, cwinBoxes, cwinBoxRows;
    int fencRows, fencCRows;
    uint32_t offCwin[2], offFenc, offFencC[2], offCost, offBar;
    uint32_t txBytes;
    size_t smemBytes;
};
// returns nullptr, or why the configuration does not fit
inline const char* me_ctu_geometry(int depth, int C, int merange, int csp, bool chromaSatd, MECtuGeom& g)
{
    const int px = depth > 8 ? 2 : 1, APX = 16 / px;
    const int hs = (csp == 1 || csp == 2) ? 1 : 0, vs = csp == 1 ? 1 : 0;
    // the integer candidates reach merange (+3: hex / square overshoot in x, which the reference range-checks in y only), a
    // quarter-pel step and the 8-tap footprint add 1 + 4
    g.R = merange + 8;
    g.costK = 4 * (merange + 8);
    int pitch = ((C + 2 * g.R + (APX - 1)) * px + 15) / 16 * 16;
    while (((pitch / 4) % 8) != 4) pitch += 16;            // 8 consecutive rows in distinct bank groups
    if (pitch > 1024) return "window rows above the 1024-byte TMA box";
    g.winPitch = pitch;
    const int winH = C + 2 * g.R + 1;                      // + 1: the shared vertical cells read one row past the 8-tap footprint
    // a TMA tile lands at a 128-byte aligned shared address: with rows that are multiples of 16 bytes, boxes of a multiple of 8 rows
    g.winBoxes = (winH + 255) / 256; g.winBoxRows = (winH + g.winBoxes - 1) / g.winBoxes;
    if (g.winBoxes > 1) g.winBoxRows = (g.winBoxRows + 7) & ~7;
    g.winRows = g.winBoxes * g.winBoxRows;
    size_t off = ((size_t)g.winPitch * g.winRows + 127) & ~(size_t)127;
    g.cwinPitch = g.cwinRows = g.cwinBoxes = g.cwinBoxRows = 0;
    g.offCwin[0] = g.offCwin[1] = g.offFencC[0] = g.offFencC[1] = 0;
    if (chromaSatd)
    {
        g.cwinPitch = ((((C + 2 * g.R) >> hs) + 6 + (APX - 1)) * px + 15) / 16 * 16;
        if (g.cwinPitch > 1024) return "chroma window rows above the 1024-byte TMA box";
        const int cwinH = ((C + 2 * g.R) >> vs) + 6;
        g.cwinBoxes = (cwinH + 255) / 256; g.cwinBoxRows = (cwinH + g.cwinBoxes - 1) / g.cwinBoxes;
        if (g.cwinBoxes > 1) g.cwinBoxRows = (g.cwinBoxRows + 7) & ~7;
        g.cwinRows = g.cwinBoxes * g.cwinBoxRows;
        for (int c = 0; c < 2; c++) { g.offCwin[c] = (uint32_t)off; off += ((size_t)g.cwinPitch * g.cwinRows + 127) & ~(size_t)127; }
    }
    g.fencRows = C; g.fencCRows = C >> vs;
    g.offFenc = (uint32_t)off; off += (size_t)64 * g.fencRows * px;
    if (chromaSatd)
        for (int c = 0; c < 2; c++) { g.offFencC[c] = (uint32_t)off; off += (size_t)64 * g.fencCRows * px; }
    g.offCost = (uint32_t)off; off += (((size_t)(2 * g.costK + 1) * 2) + 15) & ~(size_t)15;
    g.offBar = (uint32_t)off; off += 16;
    g.smemBytes = off + 16;                                // + slack: row loaders read one word past an aligned run
    g.txBytes = (uint32_t)((size_t)g.winPitch * g.winRows + (size_t)64 * g.fencRows * px +
                           (chromaSatd ? 2 * ((size_t)g.cwinPitch * g.cwinRows + (size_t)64 * g.fencCRows * px) : 0));
    if (g.smemBytes > 227 * 1024) return "shared memory per CTA above 232448 bytes";
    return nullptr;
}

} // namespace x265b200

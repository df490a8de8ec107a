// me_ctu_kernels.cu -- the general frame-shaped motion search: any CTU size (64 / 32 / 16), any partition set (2Nx2N, rect,
// AMP), per-PU predictors and MV candidates, the chroma SATD term, 8- and 10-bit, search ranges up to the shared-memory
// capacity (merange 128 at 8 bits).  One CTA per (CTU, reference):
//   * TMA (cp.async.bulk.tensor.2d + mbarrier, SASS UTMALDG) stages the reference search window -- in boxes of at most 256
//     rows; rows are addressed as 32-bit elements so a row of up to 1024 bytes is one box --, the source CTU and, with the
//     chroma term on, the Cb / Cr windows and source blocks;
//   * the MV-cost entries a search can reach are copied next to them;
//   * warps take work items (32 lanes' worth of PUs, me_ctu_layout.h) from a shared-memory queue, heaviest first, and run
//     the per-lane search of me_ctu_device.cuh.
// Blocks outside the staged window (a PU predictor far from the CTU's window centre, the zero-MV candidate) are read from
// the global plane with the same arithmetic, so results never depend on the window placement.
#include "me_ctu_device.cuh"
#include "x265b200.h"
#include <cuda.h>
#include <climits>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

namespace x265b200 {

int scratch_dev(Ctx* ctx, int slot, size_t bytes, void** out);
int ensure_mvcost(Ctx* ctx, double lambda);

struct alignas(64) MECtuMaps
{
    CUtensorMap cur[3];                   // source Y / Cb / Cr: box 64 x ctuRowsOfPlane
    CUtensorMap ref[3][MC_MAX_REFS];      // reference Y / Cb / Cr windows
};

__device__ __forceinline__ uint32_t mc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mc_tma_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(mc_smem_u32(dst)), "l"(map), "r"(mc_smem_u32(bar)), "r"(x), "r"(y) : "memory");
}

template<typename pixel>
__global__ void __launch_bounds__(512, 1)
me_ctu_kernel(const __grid_constant__ MECtuMaps maps, const MECtuArgs p, const MECtuGeom L)
{
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int px = (int)sizeof(pixel);
    constexpr int APX = 16 / px;                     // pixels per 16 bytes: TMA box starts are aligned down to this
    uint64_t* bar = (uint64_t*)(smem + L.offBar);
    int* nextItem = (int*)(smem + L.offBar + 8);
    const int lane = threadIdx.x & 31;
    const int ctu = blockIdx.x, ref = blockIdx.y;
    const int ctuX = ctu % p.ctuCols, ctuY = ctu / p.ctuCols;
    const int C = p.ctuSize;

    int mvpx = 0, mvpy = 0;
    if (p.mvpCtu) { const int32_t* m = p.mvpCtu + ((int64_t)ref * p.ctuCols * p.ctuRows + ctu) * 2; mvpx = m[0]; mvpy = m[1]; }
    const int cx = mvpx >> 2, cy = mvpy >> 2;

    MECtuStage<pixel> st;
    // luma window: picture columns [winLeft, winLeft + winPitch), rows [winTop, winTop + winRows)
    const int tx = ctuX * C + cx - p.R + p.marginX, ax = tx & ~(APX - 1);
    st.window = (const pixel*)smem; st.winLeft = ax - p.marginX; st.winTop = ctuY * C + cy - p.R;
    st.fenc = (const pixel*)(smem + L.offFenc);
    st.costS = (const uint16_t*)(smem + L.offCost) + p.costK;
    int cax = 0, fcax = 0;
    if (p.chromaSatd)
    {
        const int ctx0 = ((ctuX * C + cx - p.R) >> p.hshift) - 2 + p.cmarginX;
        cax = ctx0 & ~(APX - 1);
        st.cwinLeft = cax - p.cmarginX; st.cwinTop = ((ctuY * C + cy - p.R) >> p.vshift) - 2;
        st.cwindow[0] = (const pixel*)(smem + L.offCwin[0]); st.cwindow[1] = (const pixel*)(smem + L.offCwin[1]);
        // the source boxes start 16-byte aligned too (a 16-pixel CTU's chroma is 8 bytes wide at 8 bits): the 64-pixel box has room
        const int fcx = ((ctuX * C) >> p.hshift) + p.cmarginX;
        fcax = fcx & ~(APX - 1);
        st.fencC[0] = (const pixel*)(smem + L.offFencC[0]) + (fcx - fcax); st.fencC[1] = (const pixel*)(smem + L.offFencC[1]) + (fcx - fcax);
    }
    else
    {
        st.cwinLeft = st.cwinTop = 0; st.cwindow[0] = st.cwindow[1] = nullptr; st.fencC[0] = st.fencC[1] = nullptr;
    }

    if (threadIdx.x == 0)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mc_smem_u32(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        *nextItem = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mc_smem_u32(bar)), "r"(L.txBytes) : "memory");
        // reference rows travel as 32-bit elements: x coordinates in units of 4 bytes
        for (int b = 0; b < L.winBoxes; b++)
            mc_tma_2d(smem + (size_t)b * L.winBoxRows * L.winPitch, &maps.ref[0][ref], bar, ax * px / 4, st.winTop + p.marginY + b * L.winBoxRows);
        mc_tma_2d(smem + L.offFenc, &maps.cur[0], bar, ctuX * C + p.marginX, ctuY * C + p.marginY);
        if (p.chromaSatd)
        {
            for (int c = 0; c < 2; c++)
            {
                for (int b = 0; b < L.cwinBoxes; b++)
                    mc_tma_2d(smem + L.offCwin[c] + (size_t)b * L.cwinBoxRows * L.cwinPitch, &maps.ref[1 + c][ref], bar, cax * px / 4,
                              st.cwinTop + p.cmarginY + b * L.cwinBoxRows);
                mc_tma_2d(smem + L.offFencC[c], &maps.cur[1 + c], bar, fcax, ((ctuY * C) >> p.vshift) + p.cmarginY);
            }
        }
    }
    // the cost entries within reach of a search, while the tiles are in flight
    {
        uint16_t* cs = (uint16_t*)(smem + L.offCost);
        const uint16_t* g = p.cost + 2 * 32768 - p.costK;
        for (int i = threadIdx.x; i < 2 * p.costK + 1; i += blockDim.x) cs[i] = g[i];
    }
    __syncthreads();
    {
        // wait for the tiles (phase 0 of the barrier)
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "WAIT_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra DONE_%=;\n"
            "bra WAIT_%=;\n"
            "DONE_%=:\n"
            "}\n" :: "r"(mc_smem_u32(bar)), "r"(0) : "memory");
    }

    for (;;)
    {
        int it = 0;
        if (lane == 0) it = atomicAdd(nextItem, 1);
        it = __shfl_sync(0xffffffffu, it, 0);
        if (it >= p.numItems) break;
        const uint32_t word = p.items[it * 32 + lane];
        if (word & 0x80000000u) me_ctu_lane<pixel>(p, st, word, ctuX, ctuY, ref, lane);
        __syncwarp();
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------------
typedef CUresult (*MCEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static MCEncodeTiledFn mc_get_encode()
{
    static MCEncodeTiledFn fn = []() -> MCEncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            return (MCEncodeTiledFn)p;
        return nullptr;
    }();
    return fn;
}

// A plane as a 2-D tensor.  asWords: rows are described as 32-bit elements (box widths up to 1024 bytes); else as pixels.
static int mc_encode_plane(CUtensorMap* m, int depth, const void* origin, int64_t stride, int marginX, int marginY, int rowsTotal,
                           int boxWpx, int boxH, bool asWords, const char* what)
{
    MCEncodeTiledFn enc = mc_get_encode();
    if (!enc) { set_error("me_frame_ex: cuTensorMapEncodeTiled not available from the driver"); return -1; }
    const int px = depth > 8 ? 2 : 1;
    const char* base = (const char*)origin - ((int64_t)marginY * stride + marginX) * px;
    if (((uintptr_t)base & 15) || ((stride * px) & 15)) { set_error("me_frame_ex: %s plane base / stride must be 16-byte aligned for TMA", what); return -1; }
    cuuint64_t gdim[2] = { asWords ? (cuuint64_t)(stride * px / 4) : (cuuint64_t)stride, (cuuint64_t)rowsTotal };
    cuuint64_t gstr[1] = { (cuuint64_t)stride * px };
    cuuint32_t box[2] = { asWords ? (cuuint32_t)(boxWpx * px / 4) : (cuuint32_t)boxWpx, (cuuint32_t)boxH };
    cuuint32_t estr[2] = { 1, 1 };
    if (box[0] > 256 || box[1] > 256) { set_error("me_frame_ex: %s box %ux%u above the TMA limit", what, box[0], box[1]); return -1; }
    const CUtensorMapDataType dt = asWords ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : (depth > 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8);
    CUresult r = enc(m, dt, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("me_frame_ex: cuTensorMapEncodeTiled(%s) failed (%d)", what, (int)r); return -1; }
    return 0;
}

// the PU / item tables of a layout live on the device for the lifetime of the context
struct MECtuTables { void* dPus; void* dItems; int numPu, numItems; };
static std::map<std::tuple<Ctx*, int, int, int, int, int, int, int>, MECtuTables> g_mcTables;
static std::mutex g_mcLock;

static int mc_tables(Ctx* ctx, int ctuSize, int minCu, int rect, int amp, int csp, int chromaSatd, int sizeMask, MECtuTables& out)
{
    std::lock_guard<std::mutex> lk(g_mcLock);
    auto key = std::make_tuple(ctx, ctuSize, minCu, rect, amp, csp, chromaSatd, sizeMask);
    auto it = g_mcTables.find(key);
    if (it != g_mcTables.end()) { out = it->second; return 0; }
    MECtuLayout L;
    me_ctu_build_layout(ctuSize, minCu, rect != 0, amp != 0, csp, chromaSatd != 0, L, sizeMask);
    MECtuTables t; t.numPu = (int)L.pus.size(); t.numItems = L.numItems;
    X265B200_CHECK(cudaMalloc(&t.dPus, L.pus.size() * sizeof(MECtuPU)));
    X265B200_CHECK(cudaMalloc(&t.dItems, L.items.size() * sizeof(uint32_t)));
    X265B200_CHECK(cudaMemcpyAsync(t.dPus, L.pus.data(), L.pus.size() * sizeof(MECtuPU), cudaMemcpyHostToDevice, ctx->stream));
    X265B200_CHECK(cudaMemcpyAsync(t.dItems, L.items.data(), L.items.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    X265B200_CHECK(cudaStreamSynchronize(ctx->stream));         // the host vectors die with this scope
    g_mcTables[key] = t;
    out = t;
    return 0;
}

void me_ctu_release(Ctx* ctx)
{
    std::lock_guard<std::mutex> lk(g_mcLock);
    for (auto it = g_mcTables.begin(); it != g_mcTables.end();)
    {
        if (std::get<0>(it->first) == ctx) { cudaFree(it->second.dPus); cudaFree(it->second.dItems); it = g_mcTables.erase(it); }
        else ++it;
    }
}

static bool mc_valid_ctu(int ctuSize, int minCu) { return (ctuSize == 64 || ctuSize == 32 || ctuSize == 16) && (minCu == 8 || minCu == 16 || minCu == 32 || minCu == 64) && minCu <= ctuSize; }

int me_frame_layout(int ctuSize, int minCuSize, int rect, int amp, int32_t* outXYWH, int cap)
{
    if (!mc_valid_ctu(ctuSize, minCuSize)) { set_error("me_frame_layout: ctuSize %d / minCuSize %d", ctuSize, minCuSize); return -1; }
    std::vector<MECtuPU> v;
    me_ctu_build_pus(ctuSize, minCuSize, rect != 0, amp != 0, 0, v);
    for (int i = 0; i < (int)v.size() && i < cap; i++)
    {
        outXYWH[4 * i] = v[i].x; outXYWH[4 * i + 1] = v[i].y; outXYWH[4 * i + 2] = v[i].w; outXYWH[4 * i + 3] = v[i].h;
    }
    return (int)v.size();
}

static int me_ctu_launch(Ctx* ctx, const x265b200_me_frame_params* P, const x265b200_me_frame_planes* pl, const int32_t* mvpCtu, const int32_t* mvpPu,
                         const uint8_t* numCandPu, const int32_t* mvcPu, int32_t* out, int rawRange, int sizeMask)
{
    if (!P || !pl) { set_error("me_frame_ex: null params"); return -1; }
    if (P->numRefs <= 0 || P->ctuCols <= 0 || P->ctuRows <= 0) return 0;
    if (P->numRefs > MC_MAX_REFS) { set_error("me_frame_ex: at most %d references per call", MC_MAX_REFS); return -1; }
    if (!mc_valid_ctu(P->ctuSize, P->minCuSize)) { set_error("me_frame_ex: ctuSize %d / minCuSize %d", P->ctuSize, P->minCuSize); return -1; }
    if (P->searchMethod == ME_SEA || P->searchMethod < 0 || P->searchMethod > ME_FULL) { set_error("me_frame_ex: searchMethod %d unsupported (use x265b200_me_batch_sea_dev for --me sea)", P->searchMethod); return -1; }
    if (P->subpelRefine < 0 || P->subpelRefine > 7) { set_error("me_frame_ex: subpelRefine %d", P->subpelRefine); return -1; }
    if (P->merange < 1 || P->merange > 512) { set_error("me_frame_ex: merange %d", P->merange); return -1; }
    if (P->csp < 0 || P->csp > 3) { set_error("me_frame_ex: csp %d (0 = luma only, 1 = 4:2:0, 2 = 4:2:2, 3 = 4:4:4)", P->csp); return -1; }
    if (P->maxCand < 0 || P->maxCand > 16) { set_error("me_frame_ex: maxCand %d", P->maxCand); return -1; }
    if (P->maxCand && (!numCandPu || !mvcPu)) { set_error("me_frame_ex: maxCand > 0 needs numCandPu and mvcPu"); return -1; }
    const int depth = P->depth, px = depth > 8 ? 2 : 1;
    const int C = P->ctuSize;
    const int hs = (P->csp == 1 || P->csp == 2) ? 1 : 0, vs = P->csp == 1 ? 1 : 0;
    const int chromaSatd = P->csp != 0 && P->subpelRefine > 2;            // motion.cpp:212 (per PU: && chroma satd exists)
    if (chromaSatd && (!pl->curCb || !pl->curCr || !pl->refCb || !pl->refCr)) { set_error("me_frame_ex: csp %d with subme > 2 needs the Cb / Cr planes", P->csp); return -1; }

    MECtuGeom L;
    if (const char* why = me_ctu_geometry(depth, C, P->merange, P->csp, chromaSatd != 0, L))
    {
        set_error("me_frame_ex: merange %d at %d bits: %s (use x265b200_me_batch_dev)", P->merange, depth, why);
        return -1;
    }
    const int R = L.R, costK = L.costK, winPitch = L.winPitch, winRows = L.winRows, cwinPitch = L.cwinPitch, cwinRows = L.cwinRows;
    const size_t smem = L.smemBytes;

    const int cmarginX = P->chromaMarginX > 0 ? P->chromaMarginX : P->marginX >> hs, cmarginY = P->chromaMarginY > 0 ? P->chromaMarginY : P->marginY >> vs;
    if ((P->marginX * px) & 15) { set_error("me_frame_ex: marginX*sizeof(pixel) must be a multiple of 16 bytes (TMA box alignment)"); return -1; }
    if (chromaSatd && ((cmarginX * px) & 15)) { set_error("me_frame_ex: chroma marginX*sizeof(pixel) must be a multiple of 16 bytes"); return -1; }
    if (P->marginX < 64 || P->marginY < 16) { set_error("me_frame_ex: plane margins (%d,%d) smaller than the reference's PicYuv padding", P->marginX, P->marginY); return -1; }
    if (ensure_mvcost(ctx, P->lambda)) return -1;

    MECtuTables T;
    if (mc_tables(ctx, C, P->minCuSize, P->rect, P->amp, P->csp, chromaSatd, sizeMask, T)) return -1;

    MECtuMaps maps; memset(&maps, 0, sizeof(maps));
    const int rowsC = ((P->rowsTotal - 2 * P->marginY) >> vs) + 2 * cmarginY;
    if (mc_encode_plane(&maps.cur[0], depth, pl->curY, pl->curStride, P->marginX, P->marginY, P->rowsTotal, 64, L.fencRows, false, "source luma")) return -1;
    for (int r = 0; r < P->numRefs; r++)
        if (mc_encode_plane(&maps.ref[0][r], depth, pl->refY[r], pl->refStride, P->marginX, P->marginY, P->rowsTotal, winPitch / px, L.winBoxRows, true, "reference luma")) return -1;
    if (chromaSatd)
    {
        if (mc_encode_plane(&maps.cur[1], depth, pl->curCb, pl->curStrideC, cmarginX, cmarginY, rowsC, 64, L.fencCRows, false, "source Cb")) return -1;
        if (mc_encode_plane(&maps.cur[2], depth, pl->curCr, pl->curStrideC, cmarginX, cmarginY, rowsC, 64, L.fencCRows, false, "source Cr")) return -1;
        for (int r = 0; r < P->numRefs; r++)
        {
            if (mc_encode_plane(&maps.ref[1][r], depth, pl->refCb[r], pl->refStrideC, cmarginX, cmarginY, rowsC, cwinPitch / px, L.cwinBoxRows, true, "reference Cb")) return -1;
            if (mc_encode_plane(&maps.ref[2][r], depth, pl->refCr[r], pl->refStrideC, cmarginX, cmarginY, rowsC, cwinPitch / px, L.cwinBoxRows, true, "reference Cr")) return -1;
        }
    }

    // device-side arrays: plane origins, slice bounds
    const void* ptrs[3 * MC_MAX_REFS];
    for (int r = 0; r < P->numRefs; r++)
    {
        ptrs[r] = pl->refY[r];
        ptrs[MC_MAX_REFS + r] = chromaSatd ? pl->refCb[r] : nullptr;
        ptrs[2 * MC_MAX_REFS + r] = chromaSatd ? pl->refCr[r] : nullptr;
    }
    std::vector<int32_t> slice;
    const bool sliced = P->maxSlices > 1 && P->frameParallel;
    if (sliced)
    {
        // FrameEncoder: rows of slice k = [m_sliceBaseRow[k], m_sliceBaseRow[k+1]) (frameencoder.cpp:124-139); per CTU row
        // m_sliceMinY / m_sliceMaxY (frameencoder.cpp:1448-1453)
        const int numRows = P->sliceTotalRows > 0 ? P->sliceTotalRows : P->firstCtuRow + P->ctuRows;
        std::vector<int> base(P->maxSlices + 1, 0);
        {
            const uint32_t accu = ((uint32_t)numRows << 8) / (uint32_t)P->maxSlices;
            uint32_t rowSum = accu, sidx = 0;
            for (uint32_t i = 0; i < (uint32_t)numRows; i++)
            {
                const uint32_t rowRange = rowSum >> 8;
                if ((i >= rowRange) & (sidx != (uint32_t)P->maxSlices - 1)) { rowSum += accu; base[++sidx] = (int)i; }
            }
            for (int k = (int)sidx + 1; k < P->maxSlices; k++) base[k] = numRows;      // slices that received no row
            base[0] = 0; base[P->maxSlices] = numRows;
        }
        slice.resize((size_t)P->ctuRows * 2);
        for (int y = 0; y < P->ctuRows; y++)
        {
            const int row = y + P->firstCtuRow;
            int k = 0;
            while (k + 1 < P->maxSlices && row >= base[k + 1]) k++;
            const int rowInSlice = row - base[k], endRowPlus1 = base[k + 1];
            int mn = -(rowInSlice * C * 4) + 3 * 4, mx = (endRowPlus1 - 1 - row) * (C * 4) - 4 * 4;
            if (mx < mn) mx = mn = 0;
            slice[2 * y] = mn; slice[2 * y + 1] = mx;
        }
    }
    const size_t ptrBytes = sizeof(ptrs), sliceBytes = slice.size() * sizeof(int32_t);
    void* dScr = nullptr;
    if (scratch_dev(ctx, 5, ptrBytes + sliceBytes + 64, &dScr)) return -1;
    if (stage_small(ctx, dScr, ptrs, ptrBytes)) return -1;                     // host temporaries: staged through pinned memory, no sync
    if (sliceBytes && stage_small(ctx, (char*)dScr + ptrBytes, slice.data(), sliceBytes)) return -1;

    MECtuArgs a; memset(&a, 0, sizeof(a));
    a.refY = (const void* const*)dScr; a.refCb = a.refY + MC_MAX_REFS; a.refCr = a.refY + 2 * MC_MAX_REFS;
    a.refStride = pl->refStride; a.refStrideC = pl->refStrideC;
    a.ctuCols = P->ctuCols; a.ctuRows = P->ctuRows; a.numRefs = P->numRefs; a.ctuSize = C;
    a.marginX = P->marginX; a.marginY = P->marginY; a.cmarginX = cmarginX; a.cmarginY = cmarginY;
    a.picW = P->picWidth > 0 ? P->picWidth : P->ctuCols * C; a.picH = P->picHeight > 0 ? P->picHeight : (P->firstCtuRow + P->ctuRows) * C;
    a.firstCtuRow = P->firstCtuRow;
    a.pus = (const MECtuPU*)T.dPus; a.numPu = T.numPu; a.items = (const uint32_t*)T.dItems; a.numItems = T.numItems;
    a.mvpCtu = mvpCtu; a.mvpPu = mvpPu; a.numCandPu = P->maxCand ? numCandPu : nullptr; a.mvcPu = mvcPu; a.maxCand = P->maxCand;
    a.out = out; a.cost = ctx->dMvCost; a.costK = costK;
    a.searchMethod = P->searchMethod; a.subpelRefine = P->subpelRefine; a.merange = P->merange; a.depth = depth; a.R = R;
    a.winPitch = winPitch / px; a.winRows = winRows;
    a.csp = P->csp; a.hshift = hs; a.vshift = vs; a.chromaSatd = chromaSatd;
    a.cwinPitch = cwinPitch / px; a.cwinRows = cwinRows;
    a.sliceBounds = sliced ? (const int32_t*)((char*)dScr + ptrBytes) : nullptr;
    a.maxSlices = P->maxSlices > 1 ? P->maxSlices : 1;
    a.refLagPixels = P->refLagPixels > 0 ? P->refLagPixels : INT_MAX / 2;
    a.rawRange = rawRange;

    // warps per CTA: about 16 resident warps per SM at 128 registers per thread
    const int ctasPerSm = (int)std::max<size_t>(1, std::min<size_t>(4, (size_t)(227 * 1024) / (smem + 1024)));
    int warps = 16 / ctasPerSm;
    if (warps > T.numItems) warps = T.numItems;
    if (warps < 1) warps = 1;
    dim3 grid(P->ctuCols * P->ctuRows, P->numRefs);
    if (depth > 8)
    {
        X265B200_CHECK(cudaFuncSetAttribute(me_ctu_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        me_ctu_kernel<uint16_t><<<grid, warps * 32, smem, ctx->stream>>>(maps, a, L);
    }
    else
    {
        X265B200_CHECK(cudaFuncSetAttribute(me_ctu_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        me_ctu_kernel<uint8_t><<<grid, warps * 32, smem, ctx->stream>>>(maps, a, L);
    }
    ctx->launches++;
    return check(cudaGetLastError(), "me_frame_ex kernel launch");
}

int me_frame_ex_dev(Ctx* ctx, const x265b200_me_frame_params* P, const x265b200_me_frame_planes* pl, const int32_t* mvpCtu, const int32_t* mvpPu,
                    const uint8_t* numCandPu, const int32_t* mvcPu, int32_t* out)
{
    return me_ctu_launch(ctx, P, pl, mvpCtu, mvpPu, numCandPu, mvcPu, out, 0, 0);
}

// ---- x265b200_me_frame_dev with per-CTU predictors ---------------------------------------------------------------------------
// The 2Nx2N-only kernel of me_frame_kernels.cu reads every block from its staged window, which is only safe while the search
// cannot leave it -- true for mvp = 0, not for arbitrary predictors (a zero-MV winner far from the window centre).  With
// predictors the entry therefore runs on the general kernel (window test per block, plane fallback) with the entry's
// unclipped range semantics, and its [ref][ctu][pu] results are re-ordered into the entry's level grids.
__global__ void me_ctu_to_levels_kernel(const int32_t* in, int32_t* out, const MECtuPU* pus, int numPu, int ctuCols, int ctuRows, int numRefs,
                                        int64_t off64, int64_t off32, int64_t off16, int64_t off8, int64_t perRef)
{
    const int64_t total = (int64_t)numRefs * ctuCols * ctuRows * numPu;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    {
        const int pu = (int)(i % numPu);
        const int64_t t = i / numPu;
        const int ctu = (int)(t % (ctuCols * ctuRows)), ref = (int)(t / (ctuCols * ctuRows));
        const MECtuPU u = pus[pu];
        const int S = u.w, per = 64 / S;
        const int64_t lo = S == 64 ? off64 : (S == 32 ? off32 : (S == 16 ? off16 : off8));
        const int gx = (ctu % ctuCols) * per + u.x / S, gy = (ctu / ctuCols) * per + u.y / S;
        int32_t* o = out + ((int64_t)ref * perRef + lo + (int64_t)gy * (ctuCols * per) + gx) * 3;
        o[0] = in[3 * i]; o[1] = in[3 * i + 1]; o[2] = in[3 * i + 2];
    }
}

int me_frame_general_2Nx2N(Ctx* ctx, int depth, const void* curOrigin, int64_t curStride, const void* const* refOriginsHost, int numRefs, int64_t refStride,
                           int marginX, int marginY, int rowsTotal, int ctuCols, int ctuRows, int puMask, const int32_t* mvpCtu,
                           int searchMethod, int subpelRefine, int merange, double lambda, int32_t* out)
{
    puMask &= 15;
    if (!puMask) return 0;
    x265b200_me_frame_params P; memset(&P, 0, sizeof(P));
    P.depth = depth; P.ctuSize = 64; P.minCuSize = 8; P.ctuCols = ctuCols; P.ctuRows = ctuRows;
    P.marginX = marginX; P.marginY = marginY; P.rowsTotal = rowsTotal; P.numRefs = numRefs;
    P.searchMethod = searchMethod; P.subpelRefine = subpelRefine; P.merange = merange; P.maxSlices = 1; P.lambda = lambda;
    x265b200_me_frame_planes pl; memset(&pl, 0, sizeof(pl));
    pl.curY = curOrigin; pl.curStride = curStride; pl.refY = refOriginsHost; pl.refStride = refStride;
    MECtuTables T;
    if (mc_tables(ctx, 64, 8, 0, 0, 0, 0, puMask, T)) return -1;
    void* tmp = nullptr;
    if (scratch_dev(ctx, 6, (size_t)numRefs * ctuCols * ctuRows * T.numPu * 12, &tmp)) return -1;
    if (me_ctu_launch(ctx, &P, &pl, mvpCtu, nullptr, nullptr, nullptr, (int32_t*)tmp, 1, puMask)) return -1;
    int64_t off[4], o = 0;
    for (int l = 0; l < 4; l++) { off[l] = o; if (puMask & (1 << l)) o += (int64_t)ctuCols * ctuRows * (1 << l) * (1 << l); }
    const int64_t total = (int64_t)numRefs * ctuCols * ctuRows * T.numPu;
    const unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)ctx->smCount * 8);
    me_ctu_to_levels_kernel<<<blocks, 256, 0, ctx->stream>>>((const int32_t*)tmp, out, (const MECtuPU*)T.dPus, T.numPu, ctuCols, ctuRows, numRefs,
                                                              off[0], off[1], off[2], off[3], o);
    ctx->launches++;
    return check(cudaGetLastError(), "me_frame level re-order launch");
}

} // namespace x265b200

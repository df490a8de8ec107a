// la_search_thread.cu -- estimateCUCost phase 1 (slicetype.cpp:3216-3325: MV candidates, mvp choice, motionEstimate
// HEX / subme 1 / merange 16 on one 8x8 lowres CU) with ONE LANE PER CU.
//
// estimateCUCost walks the CUs of a (frame, list) field in reverse raster order and predicts each CU from its right
// neighbour and the three neighbours in the row below (slicetype.cpp:3269-3280), so a CU can start once the row below is
// two CUs ahead.  Here a warp owns a BAND of 32 consecutive CU rows: lane r searches row (bandBottom - r) and runs two
// columns behind lane r-1, one __syncwarp per step keeps the stagger (and publishes the MVs lane r-1 just wrote).  Only
// lane 0 waits on another warp (the top row of the band below, through a progress word in global memory); bands are
// claimed bottom-up from an atomic counter, so a band's producer is always resident.  Every lane runs the whole
// bit-exact search in per-thread mode (me_device.cuh, thread-only build: 8x8 block, SAD rows batched, the lowres
// quarter-pel average built and costed per 4x4 cell in registers) -- a 64-pixel CU cannot feed 32 lanes (the
// one-warp-per-CU kernel in lookahead_kernels.cu tops out at 0.85 ms per 32 640-CU field; profiles/r01_lookahead.txt).
#define ME_FORCE_THREAD 1
#define ME_LOWRES_ONLY 1
#define ME_BATCH_GROUPSUM 1          /* the candidate SADs of a search step through the register-only call (me_device.cuh thread_cand_sads) */
#ifndef LA_PACKED_SATD_OFF            /* packed-word 4x4 SATD (satd_packed.cuh) in the lowres search: -2 % (profiles/r02_staged_ab.txt) */
#define ME_PACKED_SATD 1
#endif
#include "me_device.cuh"
#include "lookahead_args.cuh"

namespace x265b200 {

constexpr int LAT_WARPS = 2;
// Lanes per CU.  The search of a field is a wavefront whose length is fixed by the reference's candidate order (a slice of R rows
// and W columns needs W + 2R dependent steps), so what a launch costs is the latency of ONE CU's search; with 2 lanes per CU each
// lane owns four of the CU's eight rows (SAD and SATD are sums over 4x4 cells: the two halves meet in one shuffle per cost, and
// both lanes take the same decisions), a warp covers 16 rows and twice as many warps are resident.
#ifndef LA_LANES
#define LA_LANES 1
#endif
constexpr int LAT_ROWS = 32 / LA_LANES;        // CU rows per warp
constexpr int LAT_HR = 8 / LA_LANES;           // rows of the CU a lane owns

// HME = false: the plain lookahead (search method and range are compile-time constants of the hot kernel);
// HME = true: either level of --hme (runtime method / range, optional quarter-resolution candidate)
template<typename pixel, bool HME>
__global__ void __launch_bounds__(LAT_WARPS * 32, 8)
la_search_thread_kernel(LASearchArgs p)
{
    // 8 lanes share one 8-row x 64-pixel tile: lane L's cached source CU sits at column (L & 7) * 8, row stride 64
    __shared__ __align__(16) pixel sFenc[LAT_WARPS][4][8 * 64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = p.widthInCU, Hc = p.heightInCU, ncu = W * Hc;
    const int row = lane / LA_LANES, half = lane % LA_LANES;          // CU row of the band this lane works on; which rows of the CU
    // work items: (chain, LAT_ROWS-row band of a slice), bands counted bottom-up inside their slice; item ids are ordered band-major
    // inside a chain, so the band an item waits for (the one below it in the same slice) always has a smaller id
    const int nSl = p.numSlices, R = p.rowsPerSlice;
    const int maxRows = Hc - (nSl - 1) * R;                          // the last slice is the tallest (>= R rows)
    const int bands = (max(maxRows, R) + LAT_ROWS - 1) / LAT_ROWS, perChain = bands * nSl;

    for (;;)
    {
        int id = 0;
        if (lane == 0) id = atomicAdd(p.workCounter, 1);
        id = __shfl_sync(0xffffffffu, id, 0);
        if (id >= p.numChains * perChain) return;
        const int chain = id / perChain, rem = id - chain * perChain;
        const int band = rem / nSl, slice = rem - band * nSl;
        const LAChain ch = p.chains[chain];
        const int firstY = slice * R, lastY = slice == nSl - 1 ? Hc - 1 : (slice + 1) * R - 1;
        const int cuY = lastY - band * LAT_ROWS - row;              // this lane's row (bottom-up inside the slice)
        const bool rowValid = cuY >= firstY;
        const bool lastRow = cuY == lastY;                          // estimateCUCost's lastRow: bottom row of the slice
        volatile int* below = p.progress + chain * Hc + cuY + 1;
        int32_t* mvs = p.mvPool + (int64_t)ch.mvSlot * ncu * 2;
        int32_t* mvcosts = p.mvCostPool + (int64_t)ch.mvSlot * ncu;
        const pixel* const* fencPlanes = (const pixel* const*)p.planes + ch.b * 4;
        const bool useW = p.weights && ch.wIdx >= 0 && p.weights[ch.wIdx].isWeighted;                   // wfref0, slicetype.cpp:3222
        const pixel* const* refPlanes = (const pixel* const*)p.planes + (useW ? ch.wref : ch.ref) * 4;
        // the row above this lane belongs to another warp when this is the band's top lane and the slice goes on above it
        const bool publishes = rowValid && cuY > firstY && (lane == 32 - LA_LANES);

        MEState<pixel> s;
        s.fenc = &sFenc[warp][row >> 3][(row & 7) * 8] + half * LAT_HR * 64;
        s.pred = nullptr; s.immed = nullptr;
        s.stride = p.stride; s.isLowres = true; s.perThread = true; s.chromaSatd = false; s.groupSize = LA_LANES; s.groupMask = ((1u << LA_LANES) - 1u) << (lane - half);
        s.w = 8; s.h = LAT_HR; s.lane = 0; s.depth = p.depth; s.partSizeScale = 4; s.cost = p.cost + 2 * 32768;

        int rightX = 0, rightY = 0;                                  // fencMV of the CU to the right (previous step of this lane)
        const int steps = W + 2 * (LAT_ROWS - 1);
#pragma unroll 1
        for (int t = 0; t < steps; t++)
        {
            const int cuX = W - 1 - (t - 2 * row);
            if (rowValid && cuX >= 0 && cuX < W)
            {
                const int cuXY = cuX + cuY * W;
                const int64_t pelOffset = 8 * cuX + (int64_t)8 * cuY * p.stride;
                if (row == 0 && !lastRow)
                {
                    const int need = min(W, W - cuX + 1);
                    while (*below < need) __nanosleep(64);
                    __threadfence();
                }
                // setSourcePU: cache the 8x8 source block (motion.cpp:188-189)
                {
                    constexpr int NW = 8 * (int)sizeof(pixel) / 4;
                    const pixel* fp = fencPlanes[0] + pelOffset + (int64_t)half * LAT_HR * p.stride;     // this lane's rows of the CU
#pragma unroll
                    for (int y = 0; y < LAT_HR; y++)
                    {
                        uint32_t wv[NW];
                        ld_words<pixel, NW>(fp + (int64_t)y * p.stride, wv);
                        uint32_t* d = (uint32_t*)(const_cast<pixel*>(s.fenc) + y * 64);
#pragma unroll
                        for (int i = 0; i < NW; i++) d[i] = wv[i];
                    }
                }
                for (int k = 0; k < 4; k++) s.lowres[k] = refPlanes[k] + pelOffset + (int64_t)half * LAT_HR * p.stride;
                s.fref = s.lowres[0]; s.gfref = s.lowres[0]; s.gstride = p.stride;

                // reverse-order MV prediction candidates (slicetype.cpp:3269-3280)
                int mvc[5][2], numc = 0;
                if (cuX < W - 1) { mvc[numc][0] = rightX; mvc[numc][1] = rightY; numc++; }
                if (!lastRow)
                {
                    const volatile int32_t* row = mvs + (int64_t)(cuXY + W) * 2;
                    mvc[numc][0] = row[0]; mvc[numc][1] = row[1]; numc++;
                    if (cuX > 0) { mvc[numc][0] = row[-2]; mvc[numc][1] = row[-1]; numc++; }
                    if (cuX < W - 1) { mvc[numc][0] = row[2]; mvc[numc][1] = row[3]; numc++; }
                }
                if (HME && p.hmeMvPool)
                {
                    // the quarter-resolution MV of the co-located CU, doubled (slicetype.cpp:3281-3284).  Index as written at
                    // :3228: (cuX / 2) + (cuY / 2) * widthInCU / 2 -- the product is halved, not the width, and the pitch is
                    // this level's width rather than m_4x4Width
                    const int64_t i4 = (int64_t)ch.mvSlot * p.hmeNcu + (cuX >> 1) + (((cuY >> 1) * W) >> 1);
                    if (p.hmeMvCostPool[i4] > 0) { mvc[numc][0] = p.hmeMvPool[i4 * 2] * 2; mvc[numc][1] = p.hmeMvPool[i4 * 2 + 1] * 2; numc++; }
                }
                int mvpx = 0, mvpy = 0, skipCost = 0x7fffffff;
                if (numc)
                {
                    int mvpcost = ME_COST_MAX;
                    for (int idx = 0; idx < numc; idx++)
                    {
                        int cost = lowres_qpel_cost<pixel>(s, mvc[idx][0], mvc[idx][1], true);     // lowresMC + bufSATD (:3292-3303)
                        if (cost < mvpcost) { mvpcost = cost; mvpx = mvc[idx][0]; mvpy = mvc[idx][1]; }
                        if (!(mvpx | mvpy) && ch.bBidir) skipCost = cost;                          // :3304-3305 (as written)
                    }
                }
                s.mvpx = mvpx; s.mvpy = mvpy;
                const MV2 mvmin = mv2(-cuX * 8 - 8, -cuY * 8 - 8), mvmax = mv2((W - cuX - 1) * 8 + 8, (Hc - cuY - 1) * 8 + 8);
                int ox, oy;
                int fencCost = motion_estimate<pixel>(s, mvmin, mvmax, mv2(mvpx, mvpy), 0, nullptr, p.merange, HME ? p.searchMethod : (int)ME_HEX, 1, p.maxSlices, 0, ox, oy);
                if (skipCost < 64 && skipCost < fencCost && ch.bBidir) { fencCost = skipCost; ox = 0; oy = 0; }
                rightX = ox; rightY = oy;
                if (half == 0) { mvs[cuXY * 2] = ox; mvs[cuXY * 2 + 1] = oy; mvcosts[cuXY] = fencCost; }
                if (publishes)
                {
                    __threadfence();
                    *(volatile int*)(p.progress + chain * Hc + cuY) = W - cuX;
                }
            }
            __syncwarp();          // keeps the two-column stagger and orders lane r-1's MV stores before lane r's loads
        }
    }
}

int la_search_thread_launch(Ctx* ctx, int depth, const LASearchArgs& a)
{
    const int maxRows = a.heightInCU - (a.numSlices - 1) * a.rowsPerSlice;
    const int bands = ((maxRows > a.rowsPerSlice ? maxRows : a.rowsPerSlice) + LAT_ROWS - 1) / LAT_ROWS;
    const int64_t items = (int64_t)a.numChains * bands * a.numSlices;
    int64_t blocksWanted = (items + LAT_WARPS - 1) / LAT_WARPS;
    int64_t cap = (int64_t)ctx->smCount * 8;
    unsigned blocks = (unsigned)(blocksWanted < cap ? blocksWanted : cap);
    if (a.hme)
    {
        if (depth > 8) la_search_thread_kernel<uint16_t, true><<<blocks, LAT_WARPS * 32, 0, ctx->stream>>>(a);
        else           la_search_thread_kernel<uint8_t, true><<<blocks, LAT_WARPS * 32, 0, ctx->stream>>>(a);
    }
    else if (depth > 8) la_search_thread_kernel<uint16_t, false><<<blocks, LAT_WARPS * 32, 0, ctx->stream>>>(a);
    else                la_search_thread_kernel<uint8_t, false><<<blocks, LAT_WARPS * 32, 0, ctx->stream>>>(a);
    ctx->launches++;
    return check(cudaGetLastError(), "la_search (per-thread) launch");
}

} // namespace x265b200

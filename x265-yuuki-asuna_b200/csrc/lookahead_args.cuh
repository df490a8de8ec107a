// lookahead_args.cuh -- argument blocks shared by the two implementations of estimateCUCost phase 1
// (lookahead_kernels.cu: one warp per CU; la_search_thread.cu: one lane per CU)
#pragma once
#include "common.cuh"
#include "x265b200.h"

namespace x265b200 {

struct LAChain { int32_t b, ref, bBidir, mvSlot; int32_t wref, wIdx; };   // frame indices; MV/cost pool slot; weighted planes of list 0 (wIdx < 0: none)

struct LASearchArgs
{
    const void* const* planes;     // [numFrames][4] plane origins
    int64_t stride;
    const LAChain* chains; int numChains;
    int widthInCU, heightInCU, depth, merange, maxSlices;
    int32_t* mvPool;               // [slot][ncu][2]
    int32_t* mvCostPool;           // [slot][ncu]
    int* progress;                 // [numChains][heightInCU], zeroed
    int* workCounter;              // zeroed
    const uint16_t* cost;
    // --hme (slicetype.cpp:3216-3325 with bEnableHME): the search method / range of this level (hmeSearchMethod[],
    // hmeRange[]) and, for the 8x8 level, the quarter-resolution result that becomes one more MV candidate (:3281-3284)
    int hme;                       // 0: plain lookahead (HEX, s_merange)
    int searchMethod;
    const int32_t* hmeMvPool;      // [slot][hmeNcu][2] lowerResMvs, or nullptr (the quarter-resolution level itself)
    const int32_t* hmeMvCostPool;  // [slot][hmeNcu]    lowerResMvCosts
    int hmeNcu;
    // cooperative slices (--lookahead-slices, slicetype.cpp:3079-3107): rows [k*rowsPerSlice, (k+1)*rowsPerSlice) form slice k
    // (the last slice runs to the bottom row); the bottom row of every slice is searched with lastRow = true, i.e. without the
    // MV candidates of the row below, so the slices are independent wavefronts.  numSlices = 1: the whole field is one slice.
    int rowsPerSlice, numSlices;
    // weightp (slicetype.cpp:3222): a list-0 chain searches planes[wref] when weights[wIdx].isWeighted (written by la_weights_analyse)
    const x265b200_la_weight* weights;
};

int la_search_thread_launch(Ctx* ctx, int depth, const LASearchArgs& a);

} // namespace x265b200

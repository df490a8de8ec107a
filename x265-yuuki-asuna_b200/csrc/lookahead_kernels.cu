// lookahead_kernels.cu -- the lookahead's lowres pre-analysis on the device:
//   * frame_init_lowres_core   source/common/pixel.cpp:604-628   (+ extendPicBorder :1027-1041)
//   * lowresIntraEstimate      source/encoder/slicetype.cpp:696-805
//   * estimateFrameCost / estimateCUCost   source/encoder/slicetype.cpp:3115-3388
//     (non-cooperative path: --lookahead-slices 0 / batch mode; weightp off)
//
// estimateCUCost predicts each 8x8 CU's MV from its right / below neighbours in reverse raster
// order (slicetype.cpp:3269-3280), so CUs of one (frame-triple, list) form a wavefront.  Phase 1 runs
// one warp per CU ROW; rows are claimed bottom-up from an atomic work counter and a row waits (spin
// on a per-row progress word in global memory) until the row below is two CUs ahead.  Many
// (triple, list) chains are in flight at once -- that is where the parallelism comes from.
// Phase 2 (bidir / intra compare / accumulation) is embarrassingly parallel, one thread per CU.
#include "me_device.cuh"
#include "lookahead_args.cuh"
#include "x265b200.h"
#include <vector>
#include <cstdlib>

namespace x265b200 {

int scratch_dev(Ctx* ctx, int slot, size_t bytes, void** out);
int ensure_mvcost(Ctx* ctx, double lambda);

// ---------------------------------------------------------------------------------------------
// lowres init: 2x downscale into 4 hpel planes (pixel.cpp:604-628), one thread per output pixel
// ---------------------------------------------------------------------------------------------
template<typename pixel>
__global__ void __launch_bounds__(256)
lowres_init_kernel(const pixel* __restrict__ src, int64_t srcStride, pixel* __restrict__ d0, pixel* __restrict__ dh,
                   pixel* __restrict__ dv, pixel* __restrict__ dc, int64_t dstStride, int width, int height)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= width || y >= height) return;
    const pixel* s0 = src + (int64_t)2 * y * srcStride + 2 * x;
    const pixel* s1 = s0 + srcStride;
    const pixel* s2 = s1 + srcStride;
#define LR_FILTER(a, b, c, d) ((((a + b + 1) >> 1) + ((c + d + 1) >> 1) + 1) >> 1)
    int a0 = s0[0], a1 = s0[1], a2 = s0[2], b0 = s1[0], b1 = s1[1], b2 = s1[2], c0 = s2[0], c1 = s2[1], c2 = s2[2];
    int64_t o = (int64_t)y * dstStride + x;
    d0[o] = (pixel)LR_FILTER(a0, b0, a1, b1);
    dh[o] = (pixel)LR_FILTER(a1, b1, a2, b2);
    dv[o] = (pixel)LR_FILTER(b0, c0, b1, c1);
    dc[o] = (pixel)LR_FILTER(b1, c1, b2, c2);
#undef LR_FILTER
}

// extendPicBorder (pixel.cpp:1027-1041): replicate edges into the margins; one thread per margin pixel row/col
template<typename pixel>
__global__ void __launch_bounds__(256)
extend_border_kernel(pixel* __restrict__ pic, int64_t stride, int width, int height, int marginX, int marginY)
{
    const int fullW = width + 2 * marginX, fullH = height + 2 * marginY;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)fullW * fullH) return;
    int yy = (int)(i / fullW) - marginY, xx = (int)(i % fullW) - marginX;
    if (xx >= 0 && xx < width && yy >= 0 && yy < height) return;
    int sx = min(max(xx, 0), width - 1), sy = min(max(yy, 0), height - 1);
    pic[(int64_t)yy * stride + xx] = pic[(int64_t)sy * stride + sx];
}

// the four hpel planes of a Lowres in one launch (blockIdx.y = plane)
template<typename pixel>
__global__ void __launch_bounds__(256)
extend_border4_kernel(pixel* __restrict__ p0, pixel* __restrict__ p1, pixel* __restrict__ p2, pixel* __restrict__ p3, int64_t stride, int width, int height, int marginX, int marginY)
{
    pixel* pic = blockIdx.y == 0 ? p0 : (blockIdx.y == 1 ? p1 : (blockIdx.y == 2 ? p2 : p3));
    const int fullW = width + 2 * marginX, fullH = height + 2 * marginY;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)fullW * fullH) return;
    int yy = (int)(i / fullW) - marginY, xx = (int)(i % fullW) - marginX;
    if (xx >= 0 && xx < width && yy >= 0 && yy < height) return;
    int sx = min(max(xx, 0), width - 1), sy = min(max(yy, 0), height - 1);
    pic[(int64_t)yy * stride + xx] = pic[(int64_t)sy * stride + sx];
}

int lowres_init_dev(Ctx* ctx, int depth, const void* src, int64_t srcStride, void* const planes[4], int64_t dstStride,
                    int width, int height, int marginX, int marginY)
{
    if (width <= 0 || height <= 0) return 0;
    if (height > 65535) { set_error("lowres_init: height %d", height); return -1; }
    dim3 grid((width + 255) / 256, height), block(256);
    if (depth > 8)
        lowres_init_kernel<uint16_t><<<grid, block, 0, ctx->stream>>>((const uint16_t*)src, srcStride, (uint16_t*)planes[0], (uint16_t*)planes[1], (uint16_t*)planes[2], (uint16_t*)planes[3], dstStride, width, height);
    else
        lowres_init_kernel<uint8_t><<<grid, block, 0, ctx->stream>>>((const uint8_t*)src, srcStride, (uint8_t*)planes[0], (uint8_t*)planes[1], (uint8_t*)planes[2], (uint8_t*)planes[3], dstStride, width, height);
    ctx->launches++;
    if (check(cudaGetLastError(), "lowres_init launch")) return -1;
    if (marginX > 0 || marginY > 0)
    {
        int64_t total = (int64_t)(width + 2 * marginX) * (height + 2 * marginY);
        unsigned blocks = (unsigned)((total + 255) / 256);
        dim3 g4(blocks, 4);
        if (depth > 8) extend_border4_kernel<uint16_t><<<g4, 256, 0, ctx->stream>>>((uint16_t*)planes[0], (uint16_t*)planes[1], (uint16_t*)planes[2], (uint16_t*)planes[3], dstStride, width, height, marginX, marginY);
        else           extend_border4_kernel<uint8_t><<<g4, 256, 0, ctx->stream>>>((uint8_t*)planes[0], (uint8_t*)planes[1], (uint8_t*)planes[2], (uint8_t*)planes[3], dstStride, width, height, marginX, marginY);
        ctx->launches++;
        if (check(cudaGetLastError(), "extend_border launch")) return -1;
    }
    return 0;
}

// extendPicBorder (pixel.cpp:1027-1041) / extendCURowColBorder (ipfilter.cpp:59-77, marginY = 0) on their own
int extend_border_dev(Ctx* ctx, int depth, void* origin, int64_t stride, int width, int height, int marginX, int marginY)
{
    if (width <= 0 || height <= 0 || (marginX <= 0 && marginY <= 0)) return 0;
    const int64_t total = (int64_t)(width + 2 * marginX) * (height + 2 * marginY);
    const unsigned blocks = (unsigned)((total + 255) / 256);
    if (depth > 8) extend_border_kernel<uint16_t><<<blocks, 256, 0, ctx->stream>>>((uint16_t*)origin, stride, width, height, marginX, marginY);
    else           extend_border_kernel<uint8_t><<<blocks, 256, 0, ctx->stream>>>((uint8_t*)origin, stride, width, height, marginX, marginY);
    ctx->launches++;
    return check(cudaGetLastError(), "extend_border launch");
}

// ---------------------------------------------------------------------------------------------
// small per-thread 8x8 helpers (phase 2 and intra estimate)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int satd8x8_arr(const int* a /* [64] fenc */, const int* b /* [64] pred */)
{
    int total = 0;
#pragma unroll
    for (int cy = 0; cy < 2; cy++)
#pragma unroll
        for (int cx = 0; cx < 2; cx++)
        {
            int d[4][4];
#pragma unroll
            for (int i = 0; i < 4; i++)
            {
#pragma unroll
                for (int k = 0; k < 4; k++) d[i][k] = a[(cy * 4 + i) * 8 + cx * 4 + k] - b[(cy * 4 + i) * 8 + cx * 4 + k];
                me_hadamard4(d[i][0], d[i][1], d[i][2], d[i][3]);
            }
            int t = 0;
#pragma unroll
            for (int k = 0; k < 4; k++)
            {
                me_hadamard4(d[0][k], d[1][k], d[2][k], d[3][k]);
                t += abs(d[0][k]) + abs(d[1][k]) + abs(d[2][k]) + abs(d[3][k]);
            }
            total += t >> 1;
        }
    return total;
}

// ReferencePlanes::lowresMC (lowres.h:67-92) into an int[64] array (stride 8)
template<typename pixel>
__device__ __forceinline__ void lowres_mc_arr(const pixel* const planes[4], int64_t pelOffset, int64_t stride, int qx, int qy, int* out)
{
    if ((qx | qy) & 1)
    {
        int hpelA = (qy & 2) | ((qx & 2) >> 1);
        const pixel* fa = planes[hpelA] + pelOffset + (qx >> 2) + (int64_t)(qy >> 2) * stride;
        int qmvx = qx + (qx & 1), qmvy = qy + (qy & 1);
        int hpelB = (qmvy & 2) | ((qmvx & 2) >> 1);
        const pixel* fb = planes[hpelB] + pelOffset + (qmvx >> 2) + (int64_t)(qmvy >> 2) * stride;
        for (int y = 0; y < 8; y++)
            for (int x = 0; x < 8; x++)
                out[y * 8 + x] = ((int)fa[y * stride + x] + (int)fb[y * stride + x] + 1) >> 1;
    }
    else
    {
        int hpel = (qy & 2) | ((qx & 2) >> 1);
        const pixel* f = planes[hpel] + pelOffset + (qx >> 2) + (int64_t)(qy >> 2) * stride;
        for (int y = 0; y < 8; y++)
            for (int x = 0; x < 8; x++)
                out[y * 8 + x] = f[y * stride + x];
    }
}

// ---------------------------------------------------------------------------------------------
// lowresIntraEstimate (slicetype.cpp:696-805), four lanes per 8x8 CU
// ---------------------------------------------------------------------------------------------
// [host-testable: tests/test_la_intra_cell_cpu.py compiles the functions between these markers for the CPU]
__device__ __forceinline__ int la_intra_pixel(const int* s /* 33 neighbours */, int mode, int bFilter, int y, int x, int dcVal, int depth)
{
    // N = 8 specialisation of intrapred.cpp:69-204 (see intra_kernels.cu for the general form)
    const int N = 8, N2 = 16;
    if (mode == 0)
        return ((N - 1 - x) * s[N2 + 1 + y] + (N - 1 - y) * s[1 + x] + (x + 1) * s[1 + N] + (y + 1) * s[N2 + 1 + N] + N) >> 4;
    if (mode == 1)
    {
        if (!bFilter) return dcVal;
        if (x == 0 && y == 0) return (s[1] + s[N2 + 1] + 2 * dcVal + 2) >> 2;
        if (y == 0) return (s[1 + x] + 3 * dcVal + 2) >> 2;
        if (x == 0) return (s[N2 + 1 + y] + 3 * dcVal + 2) >> 2;
        return dcVal;
    }
    const int8_t angleTable[17] = { -32, -26, -21, -17, -13, -9, -5, -2, 0, 2, 5, 9, 13, 17, 21, 26, 32 };
    const int16_t invAngleTable[8] = { 4096, 1638, 910, 630, 482, 390, 315, 256 };
    const bool hor = mode < 18;
    auto nb = [&](int i) -> int { if (!hor || i == 0) return s[i]; return i <= N2 ? s[N2 + i] : s[i - N2]; };
    const int yy = hor ? x : y, xx = hor ? y : x;
    const int angleOffset = hor ? 10 - mode : mode - 26;
    const int angle = angleTable[8 + angleOffset];
    if (!angle)
    {
        if (bFilter && xx == 0)
        {
            int v = (int16_t)(nb(1) + ((nb(N2 + 1 + yy) - nb(0)) >> 1));
            return clip3i(0, (1 << depth) - 1, v);
        }
        return nb(1 + xx);
    }
    const int angleSum = (yy + 1) * angle, offset = angleSum >> 5, fraction = angleSum & 31;
    auto ref = [&](int idx) -> int {
        if (angle > 0 || idx >= -1) return nb(idx + 1);
        int i = -2 - idx;
        int invAngleSum = 128 + (i + 1) * invAngleTable[-angleOffset - 1];
        return nb(N2 + (invAngleSum >> 8));
    };
    if (fraction) return ((32 - fraction) * ref(offset + xx) + fraction * ref(offset + xx + 1) + 16) >> 5;
    return ref(offset + xx);
}

struct LAIntraArgs
{
    const void* plane0; int64_t stride;         // lowresPlane[0] origin
    const int32_t* invQscale;                   // may be null
    int widthInCU, heightInCU, depth, intraPenalty;
    int32_t* intraCost; uint8_t* intraMode; uint16_t* lowresCosts; int32_t* rowSatds; int32_t* sums /* [2]: costEst, costEstAq */;
};

// SATD of one predicted 4x4 cell (rows y0..y0+3, columns x0..x0+3 of the 8x8 CU) of `mode` against the lane's source cell.
// Out of line: the three callers (DC, planar, the angular candidates) share one copy of the 16 inlined pixel formulas.
__device__ __noinline__ int la_intra_cell_cost(const int* nb /* 33 neighbours, shared memory */, int mode, int bFilter, int dcVal, int depth,
                                               int x0, int y0, const int* fenc /* [16] */)
{
    int d[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
#pragma unroll
        for (int k = 0; k < 4; k++) d[i][k] = fenc[i * 4 + k] - la_intra_pixel(nb, mode, bFilter, y0 + i, x0 + k, dcVal, depth);
        me_hadamard4(d[i][0], d[i][1], d[i][2], d[i][3]);
    }
    int t = 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
        me_hadamard4(d[0][k], d[1][k], d[2][k], d[3][k]);
        t += abs(d[0][k]) + abs(d[1][k]) + abs(d[2][k]) + abs(d[3][k]);
    }
    return t >> 1;
}

// One angular mode on one 4x4 cell, with the mode as DATA: the lanes of a warp belong to eight CUs whose refinement candidates
// (best +-2, +-1) differ, and the per-pixel form above then runs its direction / angle-sign / fraction branches once per distinct
// mode.  Here the four lanes of a CU first write the mode's projected reference line (intrapred.cpp:146-172: ref[-8 .. 16], the
// side neighbours projected through the inverse angle for negative angles) to shared memory, then every pixel is one two-tap
// interpolation along that line -- the same instructions whatever the mode; only the pure horizontal / vertical modes (edge
// filter on the first column, :176-189) take a short branch.
__device__ __noinline__ int la_intra_ang_cell_cost(const int* smp, const int* flt, int* line /* [25], shared by the CU's lanes */, const int* tabs /* angle[17] | invAngle[8] */,
                                                   int mode, int depth, int cell, int x0, int y0, const int* fenc /* [16] */)
{
    const bool hor = mode < 18;
    const int angleOffset = hor ? 10 - mode : mode - 26;
    const int angle = tabs[8 + angleOffset];
    const int inv = angle < 0 ? tabs[17 - angleOffset - 1] : 0;
    const int* src = min(abs(mode - 26), abs(mode - 10)) > 7 ? flt : smp;            // g_intraFilterFlags[mode] & 8
    for (int L = cell; L < 25; L += 4)
    {
        const int idx = L - 8;
        int i = (angle >= 0 || idx >= -1) ? idx + 1 : 16 + ((128 + (-1 - idx) * inv) >> 8);
        i = max(0, min(i, 32));                                                      // entries a mode never reads
        const int j = (!hor || i == 0) ? i : (i <= 16 ? 16 + i : i - 16);            // the flipped neighbour view of the horizontal modes
        line[L] = src[j];
    }
    __syncwarp();
    const int yyB = hor ? x0 : y0, xxB = hor ? y0 : x0;
    int pv[4][4];
#pragma unroll
    for (int t = 0; t < 4; t++)
    {
        const int angleSum = (yyB + t + 1) * angle, off = angleSum >> 5, f = angleSum & 31;
        const int* a = line + off + xxB + 8;
#pragma unroll
        for (int u = 0; u < 4; u++) pv[t][u] = ((32 - f) * a[u] + f * a[u + 1] + 16) >> 5;
    }
    if (!angle && xxB == 0)
    {
        // first column (row for mode 10) of the pure vertical / horizontal modes: nb(1) + ((nb(N2 + 1 + yy) - nb(0)) >> 1), clipped
        const int n1 = hor ? smp[17] : smp[1], n0 = smp[0];
#pragma unroll
        for (int t = 0; t < 4; t++)
        {
            const int side = hor ? smp[1 + yyB + t] : smp[17 + yyB + t];
            pv[t][0] = clip3i(0, (1 << depth) - 1, (int)(int16_t)(n1 + ((side - n0) >> 1)));
        }
    }
    __syncwarp();                                                                    // the line is rewritten by the next mode
    int d[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
#pragma unroll
        for (int k = 0; k < 4; k++) d[i][k] = fenc[i * 4 + k] - (hor ? pv[k][i] : pv[i][k]);
        me_hadamard4(d[i][0], d[i][1], d[i][2], d[i][3]);
    }
    int t = 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
        me_hadamard4(d[0][k], d[1][k], d[2][k], d[3][k]);
        t += abs(d[0][k]) + abs(d[1][k]) + abs(d[2][k]) + abs(d[3][k]);
    }
    return t >> 1;
}

// [host-testable end]
// FOUR LANES PER CU: a lane owns one 4x4 cell of the 8x8 CU (pu[LUMA_8x8].satd is the sum of its four 4x4 Hadamard costs,
// pixel.cpp:210-261), predicts only that cell for every mode tried and the four partial costs meet in two shuffles, so the
// lanes of a CU hold identical costs and walk the mode decision together.  The raw and the 1:2:1-filtered neighbour arrays of
// a CU are built once in shared memory (the per-thread form kept 64-entry fenc / pred arrays and two 33-entry neighbour
// arrays in local memory: 0.31 ms per 2160p frame, 15 % of the serialised step in profiles/r02_launches_summary.txt).
constexpr int LAI_CUS = 16;                                     // CUs per CTA of 64 threads
template<typename pixel>
__global__ void __launch_bounds__(64)
la_intra_kernel(LAIntraArgs p)
{
    __shared__ int sSmp[LAI_CUS][34], sFlt[LAI_CUS][34], sLine[LAI_CUS][26], sTabs[25];
    const int slot = threadIdx.x >> 2, cell = threadIdx.x & 3;
    if (threadIdx.x < 25)
    {
        const int8_t angleTable[17] = { -32, -26, -21, -17, -13, -9, -5, -2, 0, 2, 5, 9, 13, 17, 21, 26, 32 };
        const int16_t invAngleTable[8] = { 4096, 1638, 910, 630, 482, 390, 315, 256 };
        sTabs[threadIdx.x] = threadIdx.x < 17 ? (int)angleTable[threadIdx.x] : (int)invAngleTable[threadIdx.x - 17];
    }
    __syncthreads();
    const int ncu = p.widthInCU * p.heightInCU;
    const int cuXY = min(blockIdx.x * LAI_CUS + slot, ncu - 1);     // surplus lanes shadow the last CU (they take part in the shuffles)
    const bool live = blockIdx.x * LAI_CUS + slot < ncu;
    const int cuX = cuXY % p.widthInCU, cuY = cuXY / p.widthInCU;
    const pixel* pixCur = (const pixel*)p.plane0 + 8 * cuX + (int64_t)8 * cuY * p.stride;
    const int x0 = (cell & 1) * 4, y0 = (cell >> 1) * 4;
    int fenc[16];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int k = 0; k < 4; k++) fenc[i * 4 + k] = pixCur[(int64_t)(y0 + i) * p.stride + x0 + k];
    // neighbours (slicetype.cpp:731-735) and their 1:2:1 filtered version (intrapred.cpp:31-51): the 4 lanes of a CU share the work
    int* smp = sSmp[slot]; int* flt = sFlt[slot];
    const pixel* tl = pixCur - p.stride - 1;
    for (int i = cell; i <= 16; i += 4) smp[i] = tl[i];
    for (int i = 1 + cell; i <= 16; i += 4) smp[16 + i] = tl[(int64_t)i * p.stride];
    __syncwarp();
    for (int i = cell; i < 33; i += 4)
    {
        int v;
        if (i == 0) v = ((smp[0] << 1) + smp[1] + smp[17] + 2) >> 2;
        else if (i == 16 || i == 32) v = smp[i];
        else if (i == 17) v = ((smp[17] << 1) + smp[0] + smp[18] + 2) >> 2;
        else v = ((smp[i] << 1) + smp[i - 1] + smp[i + 1] + 2) >> 2;
        flt[i] = v;
    }
    __syncwarp();
    auto cuSum = [](int v) { v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2); return v; };

    int icost = ME_COST_MAX, ilowmode = 0;
    // DC (unfiltered neighbours, edge filter on since cuSize <= 16)
    {
        int dc = 8;
        for (int i = 0; i < 8; i++) dc += smp[1 + i] + smp[17 + i];
        dc = dc / 16;
        const int cost = cuSum(la_intra_cell_cost(smp, 1, 1, dc, p.depth, x0, y0, fenc));
        if (cost < icost) { icost = cost; ilowmode = 1; }
    }
    // planar uses the FILTERED neighbours (planar = !!(cuSize >= 8), slicetype.cpp:712,751)
    {
        const int cost = cuSum(la_intra_cell_cost(flt, 0, 0, 0, p.depth, x0, y0, fenc));
        if (cost < icost) { icost = cost; ilowmode = 0; }
    }
    auto angCost = [&](int mode) -> int {
        return cuSum(la_intra_ang_cell_cost(smp, flt, sLine[slot], sTabs, mode, p.depth, cell, x0, y0, fenc));
    };
    int acost = ME_COST_MAX, alowmode = 4;
#pragma unroll 1
    for (int mode = 5; mode < 35; mode += 5)
    {
        int cost = angCost(mode);
        if (cost < acost) { acost = cost; alowmode = mode; }
    }
#pragma unroll 1
    for (int dist = 2; dist >= 1; dist--)
    {
        int minusmode = alowmode - dist, plusmode = alowmode + dist;
        int cost = angCost(minusmode);
        if (cost < acost) { acost = cost; alowmode = minusmode; }
        cost = angCost(plusmode);
        if (cost < acost) { acost = cost; alowmode = plusmode; }
    }
    if (acost < icost) { icost = acost; ilowmode = alowmode; }
    icost += p.intraPenalty + 4;

    if (!live || cell) return;
    p.lowresCosts[cuXY] = (uint16_t)min(icost, (1 << 14) - 1);
    p.intraCost[cuXY] = icost;
    p.intraMode[cuXY] = (uint8_t)ilowmode;
    const bool bFrameScoreCU = (cuX > 0 && cuX < p.widthInCU - 1 && cuY > 0 && cuY < p.heightInCU - 1) || p.widthInCU <= 2 || p.heightInCU <= 2;
    int icostAq = (bFrameScoreCU && p.invQscale) ? ((icost * p.invQscale[cuXY] + 128) >> 8) : icost;
    if (bFrameScoreCU) { atomicAdd(&p.sums[0], icost); atomicAdd(&p.sums[1], icostAq); }
    atomicAdd(&p.rowSatds[cuY], icostAq);
}

int la_intra_dev(Ctx* ctx, int depth, const void* plane0, int64_t stride, int widthInCU, int heightInCU, const int32_t* invQscale,
                 int intraPenalty, int32_t* intraCost, uint8_t* intraMode, uint16_t* lowresCosts, int32_t* rowSatds, int32_t* sums)
{
    if (widthInCU <= 0 || heightInCU <= 0) return 0;
    LAIntraArgs a; a.plane0 = plane0; a.stride = stride; a.invQscale = invQscale; a.widthInCU = widthInCU; a.heightInCU = heightInCU;
    a.depth = depth; a.intraPenalty = intraPenalty; a.intraCost = intraCost; a.intraMode = intraMode; a.lowresCosts = lowresCosts;
    a.rowSatds = rowSatds; a.sums = sums;
    X265B200_CHECK(cudaMemsetAsync(rowSatds, 0, sizeof(int32_t) * heightInCU, ctx->stream));
    X265B200_CHECK(cudaMemsetAsync(sums, 0, sizeof(int32_t) * 2, ctx->stream));
    int n = widthInCU * heightInCU;
    if (depth > 8) la_intra_kernel<uint16_t><<<(n + LAI_CUS - 1) / LAI_CUS, 64, 0, ctx->stream>>>(a);
    else           la_intra_kernel<uint8_t><<<(n + LAI_CUS - 1) / LAI_CUS, 64, 0, ctx->stream>>>(a);
    ctx->launches++;
    return check(cudaGetLastError(), "la_intra launch");
}

// ---------------------------------------------------------------------------------------------
// estimateCUCost phase 1 (the MV search chains of every (triple, list) field) lives in la_search_thread.cu: one lane per CU
// ---------------------------------------------------------------------------------------------

// ---------------------------------------------------------------------------------------------
// estimateCUCost phase 2 (slicetype.cpp:3253-3262 list costs, :3326-3387): one thread per CU
// ---------------------------------------------------------------------------------------------
struct LAFinishArgs
{
    const void* const* planes; int64_t stride;
    const x265b200_la_triple* triples; int numTriples;
    int widthInCU, heightInCU, depth;
    const int32_t* mvPool; const int32_t* mvCostPool;
    const int32_t* const* intraCost;      // [numFrames] per-frame intra cost arrays
    const int32_t* const* invQscale;      // [numFrames] or null entries
    uint16_t* lowresCosts;                // [numTriples][ncu]
    int32_t* rowSatds;                    // [numTriples][heightInCU]
    int32_t* sums;                        // [numTriples][4]: costEst, costEstAq, intraMbs
};

template<typename pixel>
__global__ void __launch_bounds__(64)
la_finish_kernel(LAFinishArgs p)
{
    const int ncu = p.widthInCU * p.heightInCU;
    const int t = blockIdx.y;
    const int cuXY = blockIdx.x * blockDim.x + threadIdx.x;
    if (cuXY >= ncu) return;
    const x265b200_la_triple tr = p.triples[t];
    const int cuX = cuXY % p.widthInCU, cuY = cuXY / p.widthInCU;
    const int bBidir = tr.b < tr.p1;
    const int64_t pelOffset = 8 * cuX + (int64_t)8 * cuY * p.stride;
    int bcost = ME_COST_MAX, listused = 0;
    for (int i = 0; i < 1 + bBidir; i++)
    {
        int fencCost = p.mvCostPool[(int64_t)tr.mvSlot[i] * ncu + cuXY];
        if (fencCost < bcost) { bcost = fencCost; listused = i + 1; }
    }
    if (bBidir)
    {
        const pixel* const* f0 = (const pixel* const*)p.planes + tr.p0 * 4;
        const pixel* const* f1 = (const pixel* const*)p.planes + tr.p1 * 4;
        const pixel* fb = ((const pixel* const*)p.planes)[tr.b * 4] + pelOffset;
        int fenc[64], a[64], b2[64];
        for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) fenc[y * 8 + x] = fb[y * p.stride + x];
        const int32_t* mv0 = p.mvPool + ((int64_t)tr.mvSlot[0] * ncu + cuXY) * 2;
        const int32_t* mv1 = p.mvPool + ((int64_t)tr.mvSlot[1] * ncu + cuXY) * 2;
        lowres_mc_arr<pixel>(f0, pelOffset, p.stride, mv0[0], mv0[1], a);
        lowres_mc_arr<pixel>(f1, pelOffset, p.stride, mv1[0], mv1[1], b2);
        for (int e = 0; e < 64; e++) a[e] = (a[e] + b2[e] + 1) >> 1;
        int bicost = satd8x8_arr(fenc, a);
        if (bicost < bcost) { bcost = bicost; listused = 3; }
        const pixel* s0 = f0[0] + pelOffset; const pixel* s1 = f1[0] + pelOffset;       // coloc candidate
        for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) a[y * 8 + x] = ((int)s0[y * p.stride + x] + (int)s1[y * p.stride + x] + 1) >> 1;
        bicost = satd8x8_arr(fenc, a);
        if (bicost < bcost) { bcost = bicost; listused = 3; }
        bcost += 4;
    }
    else
    {
        bcost += 4;
        int ic = p.intraCost[tr.b][cuXY];
        if (ic < bcost) { bcost = ic; listused = 0; }
    }
    const bool bFrameScoreCU = (cuX > 0 && cuX < p.widthInCU - 1 && cuY > 0 && cuY < p.heightInCU - 1) || p.widthInCU <= 2 || p.heightInCU <= 2;
    const int32_t* iq = p.invQscale ? p.invQscale[tr.b] : nullptr;
    int bcostAq = (bFrameScoreCU && iq) ? ((bcost * iq[cuXY] + 128) >> 8) : bcost;
    if (bFrameScoreCU)
    {
        atomicAdd(&p.sums[t * 4 + 0], bcost);
        atomicAdd(&p.sums[t * 4 + 1], bcostAq);
        if (!listused && !bBidir) atomicAdd(&p.sums[t * 4 + 2], 1);
    }
    atomicAdd(&p.rowSatds[t * p.heightInCU + cuY], bcostAq);
    p.lowresCosts[(int64_t)t * ncu + cuXY] = (uint16_t)(min(bcost, (1 << 14) - 1) | (listused << 14));
}

int la_estimate_dev(Ctx* ctx, int depth, const void* const* planes, int64_t stride, int widthInCU, int heightInCU,
                    const x265b200_la_triple* triplesHost, int numTriples,
                    int32_t* mvPool, int32_t* mvCostPool, const int32_t* const* intraCost, const int32_t* const* invQscale,
                    uint16_t* lowresCosts, int32_t* rowSatds, int32_t* sums, double lambda, int maxSlices, int lookaheadSlices, const x265b200_la_hme* hme,
                    const x265b200_la_weight* weights)
{
    if (numTriples <= 0) return 0;
    if (hme)
    {
        if (!hme->lowerPlanes || !hme->lowerMvPool || !hme->lowerMvCostPool || hme->width4 <= 0 || hme->height4 <= 0) { set_error("la_estimate: incomplete HME descriptor"); return -1; }
        for (int l = 0; l < 2; l++)
            if (hme->searchMethod[l] < 0 || hme->searchMethod[l] > ME_FULL || hme->searchMethod[l] == ME_SEA || hme->range[l] < 1)
            { set_error("la_estimate: HME level %d search method %d / range %d", l, hme->searchMethod[l], hme->range[l]); return -1; }
    }
    if (maxSlices > 1) { set_error("la_estimate: maxSlices > 1 (--slices) is not supported on the lowres path"); return -1; }
    // cooperative slices (Lookahead::create, slicetype.cpp:1029-1041): at least 10 rows per slice, no more than the field
    int rowsPerSlice = heightInCU, numSlices = 1;
    if (lookaheadSlices > 1)
    {
        rowsPerSlice = heightInCU / lookaheadSlices;
        if (rowsPerSlice < 10) rowsPerSlice = 10;
        if (rowsPerSlice > heightInCU) rowsPerSlice = heightInCU;
        numSlices = heightInCU / rowsPerSlice;
    }
    if (ensure_mvcost(ctx, lambda)) return -1;
    void* dTriplesV = nullptr;
    if (scratch_dev(ctx, 6, sizeof(x265b200_la_triple) * numTriples, &dTriplesV)) return -1;
    if (stage_small(ctx, dTriplesV, triplesHost, sizeof(x265b200_la_triple) * numTriples)) return -1;
    const x265b200_la_triple* triples = (const x265b200_la_triple*)dTriplesV;
    // build the chain list (host) and upload it with the sync words
    std::vector<LAChain> chains;
    for (int t = 0; t < numTriples; t++)
    {
        const x265b200_la_triple& tr = triplesHost[t];
        int bBidir = tr.b < tr.p1;
        if (tr.doSearch[0])
        {
            LAChain c; c.b = tr.b; c.ref = tr.p0; c.bBidir = bBidir; c.mvSlot = tr.mvSlot[0];
            c.wIdx = weights && tr.weightIdx0 > 0 ? tr.weightIdx0 - 1 : -1; c.wref = tr.weightPlanes0;
            chains.push_back(c);
        }
        if (bBidir && tr.doSearch[1]) { LAChain c; c.b = tr.b; c.ref = tr.p1; c.bBidir = bBidir; c.mvSlot = tr.mvSlot[1]; c.wIdx = -1; c.wref = 0; chains.push_back(c); }
    }
    const int numChains = (int)chains.size();
    void* scratch = nullptr;
    size_t chainBytes = ((sizeof(LAChain) * (numChains ? numChains : 1) + 255) / 256) * 256;
    size_t progBytes = sizeof(int) * ((size_t)numChains * heightInCU + 64) * (hme ? 2 : 1);
    if (scratch_dev(ctx, 7, chainBytes + progBytes, &scratch)) return -1;
    LAChain* dChains = (LAChain*)scratch;
    int* dProg = (int*)((char*)scratch + chainBytes);
    if (numChains)
    {
        if (stage_small(ctx, dChains, chains.data(), sizeof(LAChain) * numChains)) return -1;       // no stream-wide sync: the lookahead runs beside other streams
        X265B200_CHECK(cudaMemsetAsync(dProg, 0, progBytes, ctx->stream));
        LASearchArgs a; a.planes = planes; a.stride = stride; a.chains = dChains; a.numChains = numChains;
        a.widthInCU = widthInCU; a.heightInCU = heightInCU; a.depth = depth; a.merange = 16; a.maxSlices = maxSlices;   // s_merange, slicetype.h:259
        a.mvPool = mvPool; a.mvCostPool = mvCostPool; a.progress = dProg; a.workCounter = dProg + (size_t)numChains * heightInCU; a.cost = ctx->dMvCost;
        a.hme = 0; a.searchMethod = ME_HEX; a.hmeMvPool = nullptr; a.hmeMvCostPool = nullptr; a.hmeNcu = 0;
        a.rowsPerSlice = rowsPerSlice; a.numSlices = numSlices; a.weights = weights;
        if (hme)
        {
            // level 0 (slicetype.cpp:3177-3188): the same chains over the quarter-resolution planes, hmeSearchMethod[0] / hmeRange[0];
            // its MVs / costs (lowerResMvs, lowerResMvCosts) feed the 8x8 level, which runs hmeSearchMethod[1] / hmeRange[1]
            // (motion.cpp:817: isHMELowres selects searchMethodL0 / L1)
            LASearchArgs h = a;
            h.planes = hme->lowerPlanes; h.stride = hme->lowerStride; h.widthInCU = hme->width4; h.heightInCU = hme->height4;
            h.merange = hme->range[0]; h.searchMethod = hme->searchMethod[0]; h.hme = 1;
            h.mvPool = hme->lowerMvPool; h.mvCostPool = hme->lowerMvCostPool; h.weights = nullptr;        // wfref0 only without hme (:3222)
            h.progress = dProg + (size_t)numChains * heightInCU + 64; h.workCounter = h.progress + (size_t)numChains * hme->height4;
            if (numSlices > 1)
            {
                // the quarter-resolution level is cut into the same NUMBER of slices (processTasks, slicetype.cpp:3086-3092)
                int r4 = hme->height4 / lookaheadSlices;
                r4 = r4 < 5 ? 5 : r4; r4 = r4 > hme->height4 ? hme->height4 : r4;
                if ((int64_t)r4 * (numSlices - 1) >= hme->height4) { set_error("la_estimate: %d slices do not fit the %d rows of the HME level", numSlices, hme->height4); return -1; }
                h.rowsPerSlice = r4; h.numSlices = numSlices;
            }
            else { h.rowsPerSlice = hme->height4; h.numSlices = 1; }
            if (hme->height4 > heightInCU) { set_error("la_estimate: HME level taller than the 8x8 level"); return -1; }
            if (la_search_thread_launch(ctx, depth, h)) return -1;
            a.hme = 1; a.searchMethod = hme->searchMethod[1]; a.merange = hme->range[1];
            a.hmeMvPool = hme->lowerMvPool; a.hmeMvCostPool = hme->lowerMvCostPool; a.hmeNcu = hme->width4 * hme->height4;
        }
        if (la_search_thread_launch(ctx, depth, a)) return -1;
    }
    X265B200_CHECK(cudaMemsetAsync(rowSatds, 0, sizeof(int32_t) * (size_t)numTriples * heightInCU, ctx->stream));
    X265B200_CHECK(cudaMemsetAsync(sums, 0, sizeof(int32_t) * (size_t)numTriples * 4, ctx->stream));
    LAFinishArgs f; f.planes = planes; f.stride = stride; f.triples = triples; f.numTriples = numTriples; f.widthInCU = widthInCU; f.heightInCU = heightInCU;
    f.depth = depth; f.mvPool = mvPool; f.mvCostPool = mvCostPool; f.intraCost = intraCost; f.invQscale = invQscale;
    f.lowresCosts = lowresCosts; f.rowSatds = rowSatds; f.sums = sums;
    dim3 grid((widthInCU * heightInCU + 63) / 64, numTriples);
    if (depth > 8) la_finish_kernel<uint16_t><<<grid, 64, 0, ctx->stream>>>(f);
    else           la_finish_kernel<uint8_t><<<grid, 64, 0, ctx->stream>>>(f);
    ctx->launches++;
    return check(cudaGetLastError(), "la_finish launch");
}

// ---------------------------------------------------------------------------------------------
// weightp in the lookahead: LookaheadTLD::weightsAnalyse (slicetype.cpp:860-961) for a batch of (fenc, ref) pairs.
// Host: the float scale / offset guess (:886-923), evaluated with the reference's own operation order.  Device: the two
// weightCostLuma passes (:805-841) in one kernel, then the accept test (:935) folded into the kernel that weights the 4 planes.
// ---------------------------------------------------------------------------------------------
struct LAWeightDev
{
    const void* fenc; const int32_t* intraCost; const void* refBuf[4]; void* wbuf[4];
    int32_t measure;                          // 0: early termination (:895-897), nothing is measured or weighted
    int32_t identity;                         // the candidate reduces to weight 1, offset 0 (:935): measured, never applied
    int32_t curScale, curDenom, curOffset;    // the candidate the second weightCostLuma pass measures (:923)
    int32_t finScale, finDenom;               // the same weight over the smaller denominator (:926-933)
};

template<typename pixel>
__global__ void __launch_bounds__(128)
la_weight_cost_kernel(const LAWeightDev* jobs, x265b200_la_weight* out, int64_t stride, int64_t padOffset, int widthInCU, int heightInCU, int depth)
{
    const LAWeightDev j = jobs[blockIdx.y];
    if (!j.measure) return;
    const int mb = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned c0 = 0, c1 = 0;
    if (mb < widthInCU * heightInCU)
    {
        const int cuX = mb % widthInCU, cuY = mb / widthInCU;
        const int64_t pixoff = (int64_t)8 * cuY * stride + 8 * cuX;
        const pixel* f = (const pixel*)j.fenc + pixoff;
        const pixel* r = (const pixel*)j.refBuf[0] + padOffset + pixoff;
        const int corr = 14 - depth, maxVal = (1 << depth) - 1;
        const int offset = j.curOffset * (1 << (depth - 8)), shift = j.curDenom + corr;
        const int round = (j.curDenom ? 1 << (j.curDenom - 1) : 0) << corr;
        int fe[64], a[64];
        constexpr int NW = 8 * (int)sizeof(pixel) / 4;
        for (int y = 0; y < 8; y++)
        {
            uint32_t wf[NW], wr[NW];
            ld_words<pixel, NW>(f + y * stride, wf);
            ld_words<pixel, NW>(r + y * stride, wr);
#pragma unroll
            for (int x = 0; x < 8; x++)
            {
                constexpr int PER = 4 / (int)sizeof(pixel), BITS = 8 * (int)sizeof(pixel);
                fe[y * 8 + x] = (wf[x / PER] >> ((x % PER) * BITS)) & ((1u << BITS) - 1);
                a[y * 8 + x] = (wr[x / PER] >> ((x % PER) * BITS)) & ((1u << BITS) - 1);
            }
        }
        const int ic = j.intraCost[mb];
        c0 = (unsigned)min(satd8x8_arr(fe, a), ic);                                      // wtPresent = 0: the unweighted reference
        for (int e = 0; e < 64; e++)
        {
            const int v = ((j.curScale * (int)(int16_t)(a[e] << corr) + round) >> shift) + offset;     // weight_pp_c, pixel.cpp:518-543
            a[e] = v < 0 ? 0 : (v > maxVal ? maxVal : v);
        }
        c1 = (unsigned)min(satd8x8_arr(fe, a), ic);
    }
    for (int o = 16; o; o >>= 1) { c0 += __shfl_xor_sync(0xffffffffu, c0, o); c1 += __shfl_xor_sync(0xffffffffu, c1, o); }
    if ((threadIdx.x & 31) == 0)
    {
        if (c0) atomicAdd(&out[blockIdx.y].origscore, c0);
        if (c1) atomicAdd(&out[blockIdx.y].score, c1);
    }
}

// the accept test of weightsAnalyse (:935) on the two sums; every thread evaluates it, one records it
__device__ __forceinline__ bool la_weight_accept(const LAWeightDev& j, unsigned origscore, unsigned score)
{
    if (!j.measure || j.identity || !origscore || !(score < origscore)) return false;             // !minscore (:907), !found, identity
    return !(__fdiv_rn(__uint2float_rn(score), __uint2float_rn(origscore)) > 0.998f);
}

template<typename pixel, int V>
__global__ void __launch_bounds__(256)
la_weight_planes_kernel(const LAWeightDev* jobs, x265b200_la_weight* out, int64_t count /* stride * paddedLines */, int depth)
{
    const int job = blockIdx.y >> 2, plane = blockIdx.y & 3;
    const LAWeightDev j = jobs[job];
    const bool on = la_weight_accept(j, out[job].origscore, out[job].score);
    if (plane == 0 && blockIdx.x == 0 && threadIdx.x == 0)
    {
        out[job].isWeighted = on;
        out[job].inputWeight = on ? j.finScale : 0; out[job].log2WeightDenom = on ? j.finDenom : 0; out[job].inputOffset = on ? j.curOffset : 0;
    }
    if (!on) return;
    const int corr = 14 - depth, maxVal = (1 << depth) - 1;
    const int offset = j.curOffset * (1 << (depth - 8)), shift = j.finDenom + corr, scale = j.finScale;
    const int round = (j.finDenom ? 1 << (j.finDenom - 1) : 0) << corr;
    const pixel* src = (const pixel*)j.refBuf[plane];
    pixel* dst = (pixel*)j.wbuf[plane];
    const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
    if (i0 >= count) return;
    pixel v[V];
    if (V > 1 && i0 + V <= count) *(uint4*)v = *(const uint4*)(src + i0);
    else for (int e = 0; e < V && i0 + e < count; e++) v[e] = src[i0 + e];
#pragma unroll
    for (int e = 0; e < V; e++)
    {
        const int w = ((scale * (int)(int16_t)((int)v[e] << corr) + round) >> shift) + offset;
        v[e] = (pixel)(w < 0 ? 0 : (w > maxVal ? maxVal : w));
    }
    if (V > 1 && i0 + V <= count) *(uint4*)(dst + i0) = *(const uint4*)v;
    else for (int e = 0; e < V && i0 + e < count; e++) dst[i0 + e] = v[e];
}

// The host half of weightsAnalyse (slicetype.cpp:886-933): the float scale / offset guess in the reference's operation order (no
// contraction: -ffp-contract=off in the Makefile).  out = { measure (0: early termination :895-897), curScale, curDenom, curOffset (the
// candidate the second weightCostLuma pass measures, :923), finScale, finDenom (the same weight over the smaller denominator, :926-933;
// the candidate only survives when it beat the unweighted cost, so the reduction applies to it), identity (weight 1, offset 0: :935) }.
void la_weight_guess(int depth, int width, int lines, uint64_t fencSum, uint64_t fencSsd, uint64_t refSum, uint64_t refSsd, int out[7])
{
    for (int i = 0; i < 7; i++) out[i] = 0;
    const float epsilon = 1.f / 128.f;
    float guessScale, fencMean, refMean;
    if (fencSsd && refSsd) guessScale = sqrtf((float)fencSsd / refSsd);
    else guessScale = 1.0f;
    fencMean = (float)fencSum / (lines * width) / (1 << (depth - 8));
    refMean = (float)refSum / (lines * width) / (1 << (depth - 8));
    if (fabsf(refMean - fencMean) < 0.5f && fabsf(1.f - guessScale) < epsilon) return;
    out[0] = 1;
    // WeightParam::setFromWeightAndOffset(w, 0, 7, true) (slice.h:304-316)
    int minscale = (int)(guessScale * 128 + 0.5f), mindenom = 7;
    while (mindenom > 0 && minscale > 127) { mindenom--; minscale >>= 1; }
    minscale = minscale < 127 ? minscale : 127;
    int curScale = minscale;
    int curOffset = (int)(fencMean - refMean * curScale / (1 << mindenom) + 0.5f);
    if (curOffset < -128 || curOffset > 127)
    {
        curOffset = curOffset < -128 ? -128 : 127;
        curScale = (int)((1 << mindenom) * (fencMean - curOffset) / refMean + 0.5f);
        curScale = curScale < 0 ? 0 : (curScale > 127 ? 127 : curScale);
    }
    out[1] = curScale; out[2] = mindenom; out[3] = curOffset;
    int fs = curScale, fd = mindenom;
    if (fd > 0 && !(fs & 1))
    {
        int idx = 0;
        if (fs) while (!((fs >> idx) & 1)) idx++; else idx = 32;
        int sh = idx < fd ? idx : fd;
        fd -= sh; fs >>= sh;
    }
    out[4] = fs; out[5] = fd;
    out[6] = (fs == 1 << fd && curOffset == 0);
}

int la_weights_analyse_dev(Ctx* ctx, int depth, const x265b200_la_weight_job* jobsHost, int numJobs, int64_t stride, int paddedLines, int64_t padOffset,
                           int width, int lines, x265b200_la_weight* out)
{
    if (numJobs <= 0) return 0;
    if (!jobsHost || !out || width <= 0 || lines <= 0 || stride <= 0 || paddedLines <= 0) { set_error("la_weights_analyse: bad arguments"); return -1; }
    std::vector<LAWeightDev> jobs(numJobs);
    bool vec = true, any = false;
    for (int n = 0; n < numJobs; n++)
    {
        const x265b200_la_weight_job& J = jobsHost[n];
        LAWeightDev& d = jobs[n];
        memset(&d, 0, sizeof(d));
        d.fenc = J.fencPlane0; d.intraCost = J.intraCost;
        for (int k = 0; k < 4; k++)
        {
            d.refBuf[k] = J.refBuffer[k]; d.wbuf[k] = J.weighted[k];
            if (!J.refBuffer[k] || !J.weighted[k]) { set_error("la_weights_analyse: job %d plane %d is NULL", n, k); return -1; }
            vec = vec && !(((uintptr_t)J.refBuffer[k] | (uintptr_t)J.weighted[k]) & 15);
        }
        if (!J.fencPlane0 || !J.intraCost) { set_error("la_weights_analyse: job %d source is NULL", n); return -1; }
        int g[7];
        la_weight_guess(depth, width, lines, J.fencSum, J.fencSsd, J.refSum, J.refSsd, g);
        if (!g[0]) continue;                                                                           // early termination: measure = 0
        d.measure = 1; any = true;
        d.curScale = g[1]; d.curDenom = g[2]; d.curOffset = g[3]; d.finScale = g[4]; d.finDenom = g[5]; d.identity = g[6];
    }
    void* dJobsV = nullptr;
    if (scratch_dev(ctx, 5, sizeof(LAWeightDev) * numJobs, &dJobsV)) return -1;
    if (stage_small(ctx, dJobsV, jobs.data(), sizeof(LAWeightDev) * numJobs)) return -1;
    const LAWeightDev* dJobs = (const LAWeightDev*)dJobsV;
    X265B200_CHECK(cudaMemsetAsync(out, 0, sizeof(x265b200_la_weight) * numJobs, ctx->stream));
    if (!any) return 0;                                                                        // every pair terminated early
    const int widthInCU = (width + 7) >> 3, heightInCU = (lines + 7) >> 3;
    dim3 gc((widthInCU * heightInCU + 127) / 128, numJobs);
    if (depth > 8) la_weight_cost_kernel<uint16_t><<<gc, 128, 0, ctx->stream>>>(dJobs, out, stride, padOffset, widthInCU, heightInCU, depth);
    else           la_weight_cost_kernel<uint8_t><<<gc, 128, 0, ctx->stream>>>(dJobs, out, stride, padOffset, widthInCU, heightInCU, depth);
    const int64_t count = stride * paddedLines;
    const int V = vec ? 16 / (depth > 8 ? 2 : 1) : 1;
    dim3 gw((unsigned)((count + 256 * V - 1) / (256 * V)), numJobs * 4);
    if (depth > 8) { if (vec) la_weight_planes_kernel<uint16_t, 8><<<gw, 256, 0, ctx->stream>>>(dJobs, out, count, depth); else la_weight_planes_kernel<uint16_t, 1><<<gw, 256, 0, ctx->stream>>>(dJobs, out, count, depth); }
    else           { if (vec) la_weight_planes_kernel<uint8_t, 16><<<gw, 256, 0, ctx->stream>>>(dJobs, out, count, depth); else la_weight_planes_kernel<uint8_t, 1><<<gw, 256, 0, ctx->stream>>>(dJobs, out, count, depth); }
    ctx->launches += 2;
    return check(cudaGetLastError(), "la_weights_analyse launch");
}

} // namespace x265b200

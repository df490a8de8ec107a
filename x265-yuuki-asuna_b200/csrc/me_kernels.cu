// me_kernels.cu -- batched full-resolution motion estimation: one warp per PU search, bit-exact
// with MotionEstimate::motionEstimate (source/encoder/motion.cpp:739-1569) for
// DIA / HEX / UMH / STAR / SEA / FULL integer search + hpel/qpel refinement (luma, plus the chroma
// SATD term of subme > 2 in the encode-style form).  The device algorithm lives in me_device.cuh.
//
// Also here: the BitCost lambda-scaled MV cost table (source/encoder/bitcost.cpp:31-110), built on
// the HOST with the reference's own float/double expression order and uploaded once per lambda.
//
// Roofline: bytes = 2*W*H*sizeof(pixel) + nPU*8 per frame pair; the pattern search
// itself is ALU/latency-bound (DESIGN.md "ME arithmetic"), the planes stay L2-resident.
#include "me_device.cuh"
#include "x265b200.h"
#include <cmath>
#include <vector>

namespace x265b200 {

// bitcost.cpp:95-110 CalculateLogs + :31-60 setQP.  `log` is the double overload applied to a float
// argument, multiplied by the float 2/log(2) (double arithmetic), rounded to float on store.
void host_bitcost_table(double lambda, uint16_t* out /* 4*32768+1, centred at out[2*32768] */)
{
    const int M = 32768;
    // built once, thread-safely (a C++11 magic static; the reference guards the same initialisation with s_costCalcLock,
    // bitcost.cpp:34-58): contexts are per thread, so two encoder worker threads can arrive here together
    static const std::vector<float> bits = []() {
        std::vector<float> b(2 * 32768 + 1);
        b[0] = 0.718f;
        // NB: inside a .cu file `log(float)` would bind to CUDA's float overload; the reference (plain
        // g++) calls the double `log`, so spell the promotions out.
        const float log2_2 = (float)((double)2.0f / log((double)2.0f));
        for (int i = 1; i <= 2 * 32768; i++)
            b[i] = (float)(log((double)(float)(i + 1)) * (double)log2_2 + (double)1.718f);
        return b;
    }();
    uint16_t* c = out + 2 * M;
    for (int i = 0; i <= 2 * M; i++)
    {
        double v = (double)bits[i] * lambda + (double)0.5f;
        double lim = (1 << 15) - 1;
        c[i] = c[-i] = (uint16_t)(v < lim ? v : lim);
    }
}

int ensure_mvcost(Ctx* ctx, double lambda)
{
    if (ctx->mvCostValid && ctx->mvCostLambda == lambda) return 0;
    const size_t n = 4 * 32768 + 1;
    if (!ctx->dMvCost) X265B200_CHECK(cudaMalloc((void**)&ctx->dMvCost, (n + 3) * sizeof(uint16_t)));
    std::vector<uint16_t> h(n);
    host_bitcost_table(lambda, h.data());
    X265B200_CHECK(cudaStreamSynchronize(ctx->stream));
    X265B200_CHECK(cudaMemcpy(ctx->dMvCost, h.data(), n * sizeof(uint16_t), cudaMemcpyHostToDevice));
    ctx->mvCostLambda = lambda; ctx->mvCostValid = true;
    return 0;
}

struct MEArgs
{
    const void* fencPlane; int64_t fencStride;
    const void* const* refPlanes;     // device array of plane pointers (indexed by job.refIdx), or nullptr
    const void* refPlane0; int64_t refStride;
    x265b200_me_job* jobs; int64_t n;
    const uint16_t* cost;
    int searchMethod, subpelRefine, merange, maxSlices, depth;
    int maxW, maxH;
    // encode-style setSourcePU (motion.cpp:193-222): Cb/Cr planes for the bChromaSATD term; csp 0 = luma only
    const void* fencC[2]; int64_t fencStrideC;
    const void* refC0[2]; const void* const* refCPlanes[2]; int64_t refStrideC;
    int csp, hshift, vshift;
    // --me sea: device array [numRefs][12] of integral-plane pointers addressed like the reference planes (else nullptr)
    const uint32_t* const* seaPlanes;
};

constexpr int ME_WARPS = 4;

template<typename pixel>
__global__ void __launch_bounds__(ME_WARPS * 32)
me_batch_kernel(MEArgs p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t fencBytes = (size_t)64 * p.maxH * sizeof(pixel);
    const size_t predBytes = ((size_t)p.maxW * p.maxH * sizeof(pixel) + 15) & ~(size_t)15;
    const size_t immedBytes = ((size_t)p.maxW * (p.maxH + 7) * sizeof(int16_t) + 15) & ~(size_t)15;
    const int csize = 64 >> p.hshift, maxHC = p.maxH >> p.vshift;
    const size_t fencCBytes = p.csp ? (((size_t)csize * maxHC * sizeof(pixel) + 15) & ~(size_t)15) : 0;
    unsigned char* base = smem + warp * (fencBytes + predBytes + immedBytes + 2 * fencCBytes);

    const int64_t j = (int64_t)blockIdx.x * ME_WARPS + warp;
    if (j >= p.n) return;
    x265b200_me_job job = p.jobs[j];
    // a job the launch was not sized for (or not a multiple of 4x4) would overrun this warp's shared-memory areas: refuse it
    // visibly instead (cost -1, MV untouched); x265b200_me_batch_dev documents maxW / maxH
    if (job.w > p.maxW || job.h > p.maxH || job.w < 4 || job.h < 4 || ((job.w | job.h) & 3))
    {
        if (lane == 0) p.jobs[j].outCost = -1;
        return;
    }

    MEState<pixel> s;
    s.fenc = (pixel*)base;
    s.pred = (pixel*)(base + fencBytes);
    s.immed = (int16_t*)(base + fencBytes + predBytes);
    s.stride = p.refStride;
    const pixel* refPlane = (const pixel*)(p.refPlanes ? p.refPlanes[job.refIdx] : p.refPlane0);
    s.fref = refPlane + job.puX + (int64_t)job.puY * p.refStride;
    s.gfref = s.fref; s.gstride = p.refStride;
    s.isLowres = false; s.perThread = false; s.groupSize = 1; s.groupMask = 0xffffffffu;
    // bChromaSATD = subpelRefine > 2 && chromaSatd != NULL && csp != I400 (motion.cpp:212); chroma[csp].pu[].satd exists
    // exactly when the chroma block is a multiple of 4x4 (pixel.cpp:1200-1226, :1262-1290; 4:4:4 aliases the luma satd)
    const int wC = job.w >> p.hshift, hC = job.h >> p.vshift;
    s.chromaSatd = p.csp != 0 && p.subpelRefine > 2 && !(wC & 3) && !(hC & 3);
    s.csize = csize; s.hshift = p.hshift; s.vshift = p.vshift; s.strideC = p.refStrideC;
    s.fencC[0] = s.fencC[1] = nullptr; s.frefC[0] = s.frefC[1] = nullptr;
    s.w = job.w; s.h = job.h; s.lane = lane; s.depth = p.depth;
    s.partSizeScale = (job.h * job.h) >> 4;                      // motion.cpp:125-126 sizeScale
    s.cost = p.cost + 2 * 32768;
    s.mvpx = job.mvpX; s.mvpy = job.mvpY;
    s.integral = p.seaPlanes ? p.seaPlanes + 12 * (p.refPlanes ? job.refIdx : 0) : nullptr;
    s.integralOff = job.puX + (int64_t)job.puY * p.refStride;

    // setSourcePU: copy the PU into the 64-stride cache (motion.cpp:188-189)
    const pixel* fp = (const pixel*)p.fencPlane + job.puX + (int64_t)job.puY * p.fencStride;
    const int gw = job.w >> 2;
    for (int u = lane; u < gw * job.h; u += 32)
    {
        int y = u / gw, x = (u - y * gw) << 2;
        if (sizeof(pixel) == 1) *(uint32_t*)((uint8_t*)s.fenc + y * 64 + x) = ld_px4((const uint8_t*)fp + (int64_t)y * p.fencStride + x);
        else
        {
            uint32_t* d = (uint32_t*)((uint16_t*)s.fenc + y * 64 + x);
            const uint16_t* q = (const uint16_t*)fp + (int64_t)y * p.fencStride + x;
            d[0] = ld_px2(q); d[1] = ld_px2(q + 2);
        }
    }
    if (s.chromaSatd)
    {
        // Yuv::copyPUFromYuv with bChroma (yuv.cpp:126-140): Cb, Cr PU into the m_csize-stride cache
        const int64_t coff = (job.puX >> p.hshift) + (int64_t)(job.puY >> p.vshift);
        for (int c = 0; c < 2; c++)
        {
            pixel* dst = (pixel*)(base + fencBytes + predBytes + immedBytes + c * fencCBytes);
            const pixel* fc = (const pixel*)p.fencC[c] + (job.puX >> p.hshift) + (int64_t)(job.puY >> p.vshift) * p.fencStrideC;
            for (int e = lane; e < wC * hC; e += 32)
            {
                const int y = e / wC, x = e - y * wC;
                dst[y * csize + x] = fc[(int64_t)y * p.fencStrideC + x];
            }
            s.fencC[c] = dst;
            const pixel* rc = (const pixel*)(p.refCPlanes[c] ? p.refCPlanes[c][job.refIdx] : p.refC0[c]);
            s.frefC[c] = rc + (job.puX >> p.hshift) + (int64_t)(job.puY >> p.vshift) * p.refStrideC;
        }
        (void)coff;
    }
    __syncwarp();

    int ox, oy;
    int cost = motion_estimate<pixel>(s, mv2(job.mvminX, job.mvminY), mv2(job.mvmaxX, job.mvmaxY), mv2(job.mvpX, job.mvpY),
                                      job.numCand, &job.mvc[0][0], p.merange, p.searchMethod, p.subpelRefine, p.maxSlices,
                                      (job.w == 64 && job.h == 64), ox, oy);
    if (lane == 0)
    {
        p.jobs[j].outMvX = ox; p.jobs[j].outMvY = oy; p.jobs[j].outCost = cost;
    }
}

int me_batch_dev(Ctx* ctx, int depth, const void* fencPlane, int64_t fencStride, const void* refPlane, const void* const* refPlanes,
                 int64_t refStride, const x265b200_me_chroma* chroma, const uint32_t* const* seaPlanes, x265b200_me_job* jobs, int64_t n,
                 int maxW, int maxH, int searchMethod, int subpelRefine, int merange, double lambda, int maxSlices)
{
    if (n <= 0) return 0;
    if (searchMethod == ME_SEA && !seaPlanes) { set_error("me_batch: --me sea needs the 12 integral planes of every reference (x265b200_me_batch_sea_dev)"); return -1; }
    if (searchMethod == ME_SEA && (merange < 0 || merange > 16000)) { set_error("me_batch: merange %d", merange); return -1; }
    if (searchMethod < 0 || searchMethod > ME_REFINE) { set_error("me_batch: searchMethod %d", searchMethod); return -1; }
    if (subpelRefine < 0 || subpelRefine > 7) { set_error("me_batch: subpelRefine %d", subpelRefine); return -1; }
    if (maxW < 8 && maxH < 8) { set_error("me_batch: inter PUs are at least 8x4 / 4x8"); return -1; }
    if (maxW > 64 || maxH > 64 || (maxW & 3) || (maxH & 3)) { set_error("me_batch: maxW/maxH %dx%d", maxW, maxH); return -1; }
    if (ensure_mvcost(ctx, lambda)) return -1;
    MEArgs a;
    a.fencPlane = fencPlane; a.fencStride = fencStride; a.refPlanes = refPlanes; a.refPlane0 = refPlane; a.refStride = refStride;
    a.jobs = jobs; a.n = n; a.cost = ctx->dMvCost; a.searchMethod = searchMethod; a.subpelRefine = subpelRefine;
    a.merange = merange; a.maxSlices = maxSlices; a.depth = depth; a.maxW = maxW; a.maxH = maxH;
    a.seaPlanes = searchMethod == ME_SEA ? seaPlanes : nullptr;
    a.csp = 0; a.hshift = a.vshift = 0; a.fencStrideC = a.refStrideC = 0;
    a.fencC[0] = a.fencC[1] = a.refC0[0] = a.refC0[1] = nullptr; a.refCPlanes[0] = a.refCPlanes[1] = nullptr;
    if (chroma && chroma->csp)
    {
        if (chroma->csp < 1 || chroma->csp > 3) { set_error("me_batch: csp %d (1 = 4:2:0, 2 = 4:2:2, 3 = 4:4:4)", chroma->csp); return -1; }
        a.csp = chroma->csp; a.hshift = chroma->csp != 3; a.vshift = chroma->csp == 1;      // x265.h:588-592, Yuv::create
        a.fencC[0] = chroma->fencCb; a.fencC[1] = chroma->fencCr; a.fencStrideC = chroma->fencStrideC;
        a.refC0[0] = chroma->refCb; a.refC0[1] = chroma->refCr; a.refStrideC = chroma->refStrideC;
        a.refCPlanes[0] = chroma->refCbPlanes; a.refCPlanes[1] = chroma->refCrPlanes;
        if (!a.fencC[0] || !a.fencC[1] || (!a.refC0[0] && !a.refCPlanes[0]) || (!a.refC0[1] && !a.refCPlanes[1])) { set_error("me_batch: chroma planes missing"); return -1; }
    }
    const size_t px = depth > 8 ? 2 : 1;
    size_t perWarp = (size_t)64 * maxH * px + (((size_t)maxW * maxH * px + 15) & ~(size_t)15) + (((size_t)maxW * (maxH + 7) * 2 + 15) & ~(size_t)15);
    if (a.csp) perWarp += 2 * ((((size_t)(64 >> a.hshift) * (maxH >> a.vshift) * px) + 15) & ~(size_t)15);
    size_t smem = perWarp * ME_WARPS;
    unsigned blocks = (unsigned)((n + ME_WARPS - 1) / ME_WARPS);
    if (depth > 8)
    {
        X265B200_CHECK(cudaFuncSetAttribute(me_batch_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        me_batch_kernel<uint16_t><<<blocks, ME_WARPS * 32, smem, ctx->stream>>>(a);
    }
    else
    {
        X265B200_CHECK(cudaFuncSetAttribute(me_batch_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        me_batch_kernel<uint8_t><<<blocks, ME_WARPS * 32, smem, ctx->stream>>>(a);
    }
    ctx->launches++;
    return check(cudaGetLastError(), "me_batch kernel launch");
}

} // namespace x265b200

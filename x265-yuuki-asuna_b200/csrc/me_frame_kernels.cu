// me_frame_kernels.cu -- frame-shaped motion estimation with TMA-staged search windows.
//
// One CTA = one (CTU, reference) pair.  The 64x64 source CTU and the reference search window
// (64 + 2*(merange+8) pixels square, enough for every integer candidate, the hex/square overshoot and the
// 8-tap sub-pel footprint) are brought into shared memory by two TMA tile loads
// (cp.async.bulk.tensor.2d, SASS UTMALDG) signalled on an mbarrier; all 85 2Nx2N PU searches of the CTU
// (64x64, 4x 32x32, 16x 16x16, 64x 8x8) then run out of shared memory with the per-thread form of the
// bit-exact device algorithm of me_device.cuh (= MotionEstimate::motionEstimate, motion.cpp:739-1569): a lane
// owns an 8-wide sub-block of a PU, the lanes of a PU add their partial costs (roles below).  The window pitch
// is 16*k bytes chosen so that 8 consecutive rows fall in distinct banks; the box start is aligned down to 16
// bytes as TMA requires.
//
// Per-PU semantics: motionEstimate(ref, mvmin = (mvp>>2) - merange, mvmax = (mvp>>2) + merange, qmvp = mvp,
// numCandidates = 0, merange, ...) with the CTU's predictor mvp shared by all its PUs (Search::setSearchRange
// shape, search.cpp:2724-2769, before picture-boundary clipping).
#define ME_FORCE_THREAD 1
#define ME_FULLRES_ONLY 1
#define ME_REF_IN_SMEM 1          /* every reference block the search reads is inside the TMA-staged window */
// Measured variants of the shared device code, all on by default here (A/B history: profiles/r01_me_frame_v10_ab.txt).  Each
// can be switched off for an A/B build (`make -C csrc exp EXPFLAGS=-DME_..._OFF=1`, scripts/ab_me_frame.py):
#ifndef ME_WINDOW_SLOW
#define ME_WINDOW_FAST 1          /* 32-bit shared-memory row addressing + packed-word 4x4 SATD (me_device.cuh, satd_packed.cuh) */
#endif
#ifndef ME_SUBPEL_PACKED_OFF
#define ME_SUBPEL_PACKED 1        /* 8-bit one-pass sub-pel paths on packed words; vertical cells from 11 rows (subpel_packed.cuh) */
#endif
#ifndef ME_BATCH_GROUPSUM_OFF
#define ME_BATCH_GROUPSUM 1       /* the K partial SADs of a sad_x3/x4 step reduced over the PU's lanes together                   */
#endif
#ifndef ME_VCELL_REUSE_OFF
#define ME_VCELL_REUSE 1          /* vertical cells share their transposed row blocks (+2.4 %, profiles/r02_staged_ab.txt)         */
#endif
#include "me_device.cuh"
#include "x265b200.h"
#include <cuda.h>
#include <vector>
#include <cstring>

namespace x265b200 {

int scratch_dev(Ctx* ctx, int slot, size_t bytes, void** out);
int ensure_mvcost(Ctx* ctx, double lambda);
int me_frame_general_2Nx2N(Ctx* ctx, int depth, const void* curOrigin, int64_t curStride, const void* const* refOriginsHost, int numRefs, int64_t refStride,
                           int marginX, int marginY, int rowsTotal, int ctuCols, int ctuRows, int puMask, const int32_t* mvpCtu,
                           int searchMethod, int subpelRefine, int merange, double lambda, int32_t* out);

constexpr int MF_MAX_REFS = 6;
// TMA descriptors travel as a __grid_constant__ kernel parameter (the canonical, fence-free way)
struct alignas(64) MEFrameMaps
{
    CUtensorMap cur;                  // current plane, box 64x64
    CUtensorMap ref[MF_MAX_REFS];     // reference planes, box winW x winH
};

struct MEFrameArgs
{
    const void* const* refOrigins;    // device array of plane origins (zero-MV candidate outside the window)
    int64_t refStride;
    int ctuCols, ctuRows, numRefs, marginX, marginY;
    const int32_t* mvpCtu;            // [ref][ctu][2] quarter-pel, or null (= 0)
    int32_t* out;                     // [ref][level grid][3]
    int64_t levelOff[4];              // element offset (in PUs) of level 64/32/16/8 inside one reference's block
    int64_t perRef;                   // PUs per reference
    int puMask;
    const uint16_t* cost;
    int searchMethod, subpelRefine, merange, depth, R, winW, winH;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

constexpr int MF_WARPS = 4;

// CTA roles (4 warps).  EVERY search runs in per-thread mode: a lane owns an 8-pixel-wide sub-block of the PU, runs the
// whole bit-exact search on it and the lanes of one PU sum their SAD/SATD partials with shuffles (they then hold
// identical costs and follow identical control flow).  One code path for all PU sizes keeps the hot code inside the
// instruction cache: with warp-cooperative 64x64/32x32 roles next to per-thread 16x16/8x8 roles the profile showed
// 42% of warp time in stall_no_instruction (profiles/r01_me_frame_roles.txt), each role alone < 3%.
//   warp 0 : the 64x64 PU, 32 lanes x (8 wide x 16 tall)
//   warp 1 : the four 32x32 PUs, two at a time, 16 lanes x 8x8 each
//   warp 2 : the sixteen 16x16 PUs, eight at a time, 4 lanes x 8x8 each
//   warp 3 : the sixty-four 8x8 PUs, thirty-two at a time, one lane each
// No scratch at all: sub-pel predictions are produced and costed cell by cell in registers (thread_subpel_cost).
// Residency: 4 CTAs per SM at 128 registers (shared-memory limited: 43 KB window + 4 KB fenc); 3 CTAs at 168 registers measured
// the same after the sub-pel rewrite, 2 CTAs at 212 registers slower.
template<typename pixel>
__global__ void __launch_bounds__(MF_WARPS * 32, 4)
me_frame_kernel(const __grid_constant__ MEFrameMaps maps, MEFrameArgs p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int px = (int)sizeof(pixel);
    const size_t winBytes = ((size_t)p.winW * p.winH * px + 127) & ~(size_t)127;
    pixel* window = (pixel*)smem;
    pixel* fencCtu = (pixel*)(smem + winBytes);
    uint64_t* bar = (uint64_t*)(smem + winBytes + 64 * 64 * px);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int ctu = blockIdx.x, ref = blockIdx.y;
    const int ctuX = ctu % p.ctuCols, ctuY = ctu / p.ctuCols;
    int mvpx = 0, mvpy = 0;
    if (p.mvpCtu) { const int32_t* m = p.mvpCtu + ((int64_t)ref * p.ctuCols * p.ctuRows + ctu) * 2; mvpx = m[0]; mvpy = m[1]; }
    const int cx = mvpx >> 2, cy = mvpy >> 2;
    const int wx0 = ctuX * 64 + cx - p.R, wy0 = ctuY * 64 + cy - p.R;        // picture coordinates of the search window
    // TMA needs the box's inner start coordinate 16-byte aligned (measured on this B200: an unaligned x raises
    // "illegal instruction"), so the box starts at the aligned-down column and `ex` pixels are skipped.
    const int tx = wx0 + p.marginX, ax = tx & ~(16 / px - 1), ex = tx - ax;

    if (threadIdx.x == 0)
    {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        mbar_expect_tx(bar, (uint32_t)((size_t)p.winW * p.winH * px + 64 * 64 * px));
        tma_load_2d(window, &maps.ref[ref], bar, ax, wy0 + p.marginY);
        tma_load_2d(fencCtu, &maps.cur, bar, ctuX * 64 + p.marginX, ctuY * 64 + p.marginY);
    }
    mbar_wait(bar, 0);

    // ---- roles ---------------------------------------------------------------------------------------------------
    MEState<pixel> s;
    s.stride = p.winW; s.isLowres = false; s.perThread = true; s.chromaSatd = false; s.lane = 0; s.depth = p.depth;
    s.cost = p.cost + 2 * 32768; s.mvpx = mvpx; s.mvpy = mvpy; s.gstride = p.refStride;

    // level = log2(64 / PU size); per round a warp searches 32 / lanesPerPu PUs; lane q of a PU owns sub-block (sx, sy)
    const int level = warp, rounds = warp == 0 ? 1 : 2;
    const int lanesLog2 = warp == 0 ? 5 : 6 - 2 * warp;                   // 32, 16, 4, 1 lanes per PU
    const int q = lane & ((1 << lanesLog2) - 1), pusPerRound = 32 >> lanesLog2;
    const int subH = warp == 0 ? 16 : 8;
    const int subCols = warp == 0 ? 8 : (8 >> warp);                       // sub-blocks per PU row: 8, 4, 2, 1
    const int sx = (q % subCols) * 8, sy = (q / subCols) * subH;
    s.groupSize = 1 << lanesLog2;
    s.groupMask = lanesLog2 == 5 ? 0xffffffffu : (((1u << (1 << lanesLog2)) - 1u) << (lane & ~((1 << lanesLog2) - 1)));
    s.pred = nullptr; s.immed = nullptr;       // per-thread searches never store a prediction (thread_subpel_cost)
    s.w = 8; s.h = subH;

    if ((p.puMask >> level) & 1)
    {
        const int sz = 64 >> level, per = 1 << level;
        s.partSizeScale = (sz * sz) >> 4;
#pragma unroll 1
        for (int i = 0; i < rounds; i++)
        {
            const int idx = i * pusPerRound + (lane >> lanesLog2);
            const int puy = (idx / per) * sz, pux = (idx % per) * sz;
            s.fenc = fencCtu + (puy + sy) * 64 + pux + sx;
            s.fref = window + (int64_t)(puy + sy - cy + p.R) * p.winW + (pux + sx - cx + p.R + ex);
            s.gfref = (const pixel*)p.refOrigins[ref] + (ctuX * 64 + pux + sx) + (int64_t)(ctuY * 64 + puy + sy) * p.refStride;
            int ox, oy;
            int cost = motion_estimate<pixel>(s, mv2(cx - p.merange, cy - p.merange), mv2(cx + p.merange, cy + p.merange), mv2(mvpx, mvpy),
                                              0, nullptr, p.merange, p.searchMethod, p.subpelRefine, 1, sz == 64, ox, oy);
            if (q == 0)
            {
                const int gx = ctuX * per + pux / sz, gy = ctuY * per + puy / sz;
                int32_t* o = p.out + ((int64_t)ref * p.perRef + p.levelOff[level] + (int64_t)gy * (p.ctuCols * per) + gx) * 3;
                o[0] = ox; o[1] = oy; o[2] = cost;
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn)
    {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

static int encode_plane(CUtensorMap* m, int depth, const void* origin, int64_t stride, int marginX, int marginY, int rowsTotal, int boxW, int boxH)
{
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("me_frame: cuTensorMapEncodeTiled not available from the driver"); return -1; }
    const int px = depth > 8 ? 2 : 1;
    const char* base = (const char*)origin - ((int64_t)marginY * stride + marginX) * px;
    if (((uintptr_t)base & 15) || ((stride * px) & 15)) { set_error("me_frame: plane base / stride must be 16-byte aligned for TMA"); return -1; }
    cuuint64_t gdim[2] = { (cuuint64_t)stride, (cuuint64_t)rowsTotal };
    cuuint64_t gstr[1] = { (cuuint64_t)stride * px };
    cuuint32_t box[2] = { (cuuint32_t)boxW, (cuuint32_t)boxH };
    cuuint32_t estr[2] = { 1, 1 };
    CUresult r = enc(m, depth > 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)base, gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("me_frame: cuTensorMapEncodeTiled failed (%d)", (int)r); return -1; }
    return 0;
}

int me_frame_dev(Ctx* ctx, int depth, const void* curOrigin, int64_t curStride, const void* const* refOriginsHost, int numRefs, int64_t refStride,
                 int marginX, int marginY, int rowsTotal, int ctuCols, int ctuRows, int puMask, const int32_t* mvpCtu,
                 int searchMethod, int subpelRefine, int merange, double lambda, int32_t* out)
{
    if (numRefs <= 0 || ctuCols <= 0 || ctuRows <= 0) return 0;
    if (searchMethod == ME_SEA || searchMethod < 0 || searchMethod > ME_FULL) { set_error("me_frame: searchMethod %d unsupported", searchMethod); return -1; }
    if (subpelRefine < 0 || subpelRefine > 7) { set_error("me_frame: subpelRefine %d", subpelRefine); return -1; }
    // This kernel reads every reference block from its staged window.  That is safe exactly when no search can leave the
    // window: predictor 0 (then the zero-MV candidate IS the predictor, and the patterns stay within merange + 3).  With per-CTU
    // predictors a zero-MV winner can sit anywhere, so those calls run on the general kernel, which tests every block against
    // its window and falls back to the plane (me_ctu_kernels.cu); same results, same output order.
    if (mvpCtu)
        return me_frame_general_2Nx2N(ctx, depth, curOrigin, curStride, refOriginsHost, numRefs, refStride, marginX, marginY, rowsTotal, ctuCols, ctuRows,
                                      puMask, mvpCtu, searchMethod, subpelRefine, merange, lambda, out);
    const int px = depth > 8 ? 2 : 1;
    const int R = merange + 8;
    int winW = 64 + 2 * R + (16 / px - 1);            // + slack for the 16-byte alignment of the TMA box start
    // pitch: a multiple of 16 bytes whose word count is 4 mod 8 -> 8 consecutive rows hit distinct bank groups
    int pitchBytes = ((winW * px + 15) / 16) * 16;
    // prefer a pitch whose word count is 4 mod 8 (8 consecutive rows in distinct bank groups) unless that would push
    // the CTA past a quarter of the SM's shared memory (4 resident CTAs matter more than 2-way conflicts)
    {
        int alt = pitchBytes;
        while (((alt / 4) % 8) != 4) alt += 16;
        size_t fixedBytes = (size_t)64 * 64 * px + 16 + 1024;
        if ((((size_t)alt * (64 + 2 * R) + 127) & ~(size_t)127) + fixedBytes <= 233472 / 4) pitchBytes = alt;
        else if (((pitchBytes / 4) % 32) == 0) pitchBytes += 16;
    }
    winW = pitchBytes / px;
    const int winH = 64 + 2 * R;
    if (winW > 256 || winH > 256) { set_error("me_frame: merange %d needs a %dx%d window, above the 256-element TMA box limit; use x265b200_me_batch_dev", merange, winW, winH); return -1; }
    if ((marginX * px) & 15) { set_error("me_frame: marginX*sizeof(pixel) must be a multiple of 16 bytes (TMA box alignment)"); return -1; }
    if (marginX < R + 8 || marginY < R) { set_error("me_frame: plane margins (%d,%d) smaller than the search window reach %d", marginX, marginY, R); return -1; }
    if (ensure_mvcost(ctx, lambda)) return -1;

    if (numRefs > MF_MAX_REFS) { set_error("me_frame: at most %d references per call", MF_MAX_REFS); return -1; }
    MEFrameMaps maps;
    memset(&maps, 0, sizeof(maps));
    if (encode_plane(&maps.cur, depth, curOrigin, curStride, marginX, marginY, rowsTotal, 64, 64)) return -1;
    for (int r = 0; r < numRefs; r++)
        if (encode_plane(&maps.ref[r], depth, refOriginsHost[r], refStride, marginX, marginY, rowsTotal, winW, winH)) return -1;
    void* dScr = nullptr;
    size_t ptrBytes = sizeof(void*) * numRefs;
    if (scratch_dev(ctx, 5, ptrBytes + 64, &dScr)) return -1;
    if (stage_small(ctx, dScr, refOriginsHost, ptrBytes)) return -1;

    MEFrameArgs a;
    a.refOrigins = (const void* const*)dScr; a.refStride = refStride;
    a.ctuCols = ctuCols; a.ctuRows = ctuRows; a.numRefs = numRefs; a.marginX = marginX; a.marginY = marginY;
    a.mvpCtu = mvpCtu; a.out = out; a.puMask = puMask; a.cost = ctx->dMvCost;
    a.searchMethod = searchMethod; a.subpelRefine = subpelRefine; a.merange = merange; a.depth = depth; a.R = R; a.winW = winW; a.winH = winH;
    int64_t off = 0;
    for (int l = 0; l < 4; l++)
    {
        a.levelOff[l] = off;
        if (puMask & (1 << l)) off += (int64_t)ctuCols * ctuRows * (1 << l) * (1 << l);
    }
    a.perRef = off;
    size_t smem = (((size_t)winW * winH * px + 127) & ~(size_t)127) + (size_t)64 * 64 * px + 16;
    dim3 grid(ctuCols * ctuRows, numRefs);
    if (depth > 8)
    {
        X265B200_CHECK(cudaFuncSetAttribute(me_frame_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        me_frame_kernel<uint16_t><<<grid, MF_WARPS * 32, smem, ctx->stream>>>(maps, a);
    }
    else
    {
        X265B200_CHECK(cudaFuncSetAttribute(me_frame_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        me_frame_kernel<uint8_t><<<grid, MF_WARPS * 32, smem, ctx->stream>>>(maps, a);
    }
    ctx->launches++;
    return check(cudaGetLastError(), "me_frame kernel launch");
}

} // namespace x265b200

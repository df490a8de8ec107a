// sad_stream_kernels.cu -- the streaming ME-SAD kernel: SAD at the (zero) predictor of every 2Nx2N PU 8x8 .. 64x64 of every
// CTU of a frame against each of its references (step 1 of MotionEstimate::motionEstimate, motion.cpp:771-784, for all PU
// levels at once; pu[].sad of pixel.cpp:40-55 is additive over sub-blocks), for MANY (frame, references) groups in ONE launch.
//
// This is the kernel BASELINE's "ME SAD achieved HBM GB/s" is measured on, so it is built as a pure stream:
//   * frames live in a pool of equal planes (the DPB shape: base + frame pitch), described by ONE 3-D tensor map;
//   * persistent CTAs (2 per SM) walk a static list of tile jobs -- (group, CTU row, 256-byte column chunk) -- and keep a ring
//     of 6 x 16 KB shared-memory stages full with TMA tile loads (cp.async.bulk.tensor.3d, SASS UTMALDG) signalled on
//     mbarriers: per job the source tile arrives once and is held while the tiles of the references stream past it, so every
//     plane byte crosses HBM -> SM exactly once per use and 80 KB per CTA are in flight;
//   * a warp owns a 64-byte column strip of the tile (one CTU at 8 bits), a lane a 16-byte x 8-row piece: 8 x 2 LDS.128,
//     VABSDIFF4 (8-bit) / packed 16-bit absolute differences, then shuffles up the 8 -> 16 -> 32 -> 64 pyramid.
// Algorithmic bytes per group: (1 + numRefs) * W * H * sizeof(pixel) + 4 * numRefs * (n8 + n16 + n32 + n64).
#include "common.cuh"
#include "x265b200.h"
#include <cuda.h>
#include <cstring>
#include <vector>

namespace x265b200 {

constexpr int SS_STAGES = 6;
constexpr int SS_TILE_BYTES = 256 * 64;          // 64 rows of 256 bytes
constexpr int SS_MAX_REFS = 8;

struct SSGroup { int32_t cur; int32_t ref[SS_MAX_REFS]; };

constexpr int SS_MAX_GROUPS = 16;                // groups per launch: they travel in the kernel parameters (no upload, no sync)
struct SSArgs
{
    SSGroup groups[SS_MAX_GROUPS]; int numGroups, numRefs;
    int ctuCols, ctuRows, marginX, marginY, chunks;      // chunks = 256-byte column chunks per CTU row
    int32_t* out8; int32_t* out16; int32_t* out32; int32_t* out64;
};

__device__ __forceinline__ uint32_t ss_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ss_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(ss_smem(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void ss_tma_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(ss_smem(bar)), "r"(SS_TILE_BYTES) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(ss_smem(dst)), "l"(map), "r"(ss_smem(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}

template<typename pixel> __device__ __forceinline__ uint32_t ss_sad_word(uint32_t a, uint32_t b);
template<> __device__ __forceinline__ uint32_t ss_sad_word<uint8_t>(uint32_t a, uint32_t b) { return __vsadu4(a, b); }
template<> __device__ __forceinline__ uint32_t ss_sad_word<uint16_t>(uint32_t a, uint32_t b) { return sad_u16x2(a, b); }

template<typename pixel>
__global__ void __launch_bounds__(128, 2)
sad_stream_kernel(const __grid_constant__ CUtensorMap pool, const __grid_constant__ SSArgs p)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* full = (uint64_t*)(smem + SS_STAGES * SS_TILE_BYTES);          // [SS_STAGES] tile landed
    uint64_t* empty = full + SS_STAGES;                                      // [SS_STAGES] all 4 warps done with the stage
    int* half64 = (int*)(empty + SS_STAGES);                                 // 16-bit pixels: the two half-CTU sums of a 64x64 PU
    constexpr int px = (int)sizeof(pixel);
    constexpr int PXT = 256 / px;                    // pixels per tile row: 4 CTUs at 8 bits, 2 at 16
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int planes = 1 + p.numRefs;
    const int64_t jobsTotal = (int64_t)p.numGroups * p.ctuRows * p.chunks;
    // this CTA's jobs: blockIdx.x, blockIdx.x + gridDim.x, ...
    const int64_t myJobs = jobsTotal > blockIdx.x ? (jobsTotal - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t myTiles = myJobs * planes;

    if (threadIdx.x == 0)
    {
        for (int s = 0; s < SS_STAGES; s++)
        {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(ss_smem(&full[s])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(ss_smem(&empty[s])), "r"(4));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // tile t of this CTA -> (job, plane) -> tensor coordinates
    auto issue = [&](int64_t t) {
        const int64_t j = blockIdx.x + (t / planes) * (int64_t)gridDim.x;
        const int which = (int)(t % planes);
        const int chunk = (int)(j % p.chunks);
        const int64_t r = j / p.chunks;
        const int ctuY = (int)(r % p.ctuRows), g = (int)(r / p.ctuRows);
        const SSGroup& grp = p.groups[g];      // (kernel parameter space: indexed loads from the constant bank)
        const int frame = which == 0 ? grp.cur : grp.ref[which - 1];
        const int s = (int)(t % SS_STAGES);
        ss_tma_3d(smem + s * SS_TILE_BYTES, &pool, &full[s], p.marginX + chunk * PXT, p.marginY + ctuY * 64, frame);
    };
    if (threadIdx.x == 0)
        for (int64_t t = 0; t < SS_STAGES && t < myTiles; t++) issue(t);

    const int cg = lane & 3, rg = lane >> 2;                 // 16-byte column of the strip, 8-row group
    const uint32_t laneOff = (uint32_t)(rg * 8) * 256u + (uint32_t)warp * 64u + (uint32_t)cg * 16u;
    const int64_t nctu = (int64_t)p.ctuCols * p.ctuRows;
    const int n8x = p.ctuCols * 8, n16x = p.ctuCols * 4, n32x = p.ctuCols * 2;

    int64_t t = 0;
    int meet = 0;                                             // 16-bit pixels: meetings of a warp pair so far
    for (int64_t k = 0; k < myJobs; k++)
    {
        const int64_t j = blockIdx.x + k * (int64_t)gridDim.x;
        const int chunk = (int)(j % p.chunks);
        const int64_t rr = j / p.chunks;
        const int ctuY = (int)(rr % p.ctuRows), g = (int)(rr / p.ctuRows);
        // pixel position of this lane's piece inside the picture
        const int x0 = chunk * PXT + (warp * 64 + cg * 16) / px, y0 = ctuY * 64 + rg * 8;
        const bool valid = x0 < p.ctuCols * 64;

        const int64_t tc = t;                                 // the source tile of the job
        const int sc = (int)(tc % SS_STAGES);
        ss_wait(&full[sc], (uint32_t)((tc / SS_STAGES) & 1));
        uint4 c[8];
#pragma unroll
        for (int y = 0; y < 8; y++) c[y] = *(const uint4*)(smem + sc * SS_TILE_BYTES + laneOff + y * 256);
        // the source rows now live in registers: the stage can be refilled
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(ss_smem(&empty[sc])) : "memory");
        if (threadIdx.x == 0 && tc + SS_STAGES < myTiles)
        {
            ss_wait(&empty[sc], (uint32_t)((tc / SS_STAGES) & 1));
            issue(tc + SS_STAGES);
        }
        t++;

        for (int ref = 0; ref < p.numRefs; ref++, t++)
        {
            const int sr = (int)(t % SS_STAGES);
            ss_wait(&full[sr], (uint32_t)((t / SS_STAGES) & 1));
            uint32_t sa = 0, sb = 0;                          // the two 8-byte halves of the 16-byte piece
#pragma unroll
            for (int y = 0; y < 8; y++)
            {
                const uint4 b = *(const uint4*)(smem + sr * SS_TILE_BYTES + laneOff + y * 256);
                sa += ss_sad_word<pixel>(c[y].x, b.x) + ss_sad_word<pixel>(c[y].y, b.y);
                sb += ss_sad_word<pixel>(c[y].z, b.z) + ss_sad_word<pixel>(c[y].w, b.w);
            }
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(ss_smem(&empty[sr])) : "memory");
            if (threadIdx.x == 0 && t + SS_STAGES < myTiles)
            {
                ss_wait(&empty[sr], (uint32_t)((t / SS_STAGES) & 1));
                issue(t + SS_STAGES);
            }

            // ---- the pyramid ------------------------------------------------------------------------------------
            const int64_t oBase = (int64_t)g * p.numRefs + ref;
            int32_t* o8 = p.out8 + oBase * nctu * 64; int32_t* o16 = p.out16 + oBase * nctu * 16;
            int32_t* o32 = p.out32 + oBase * nctu * 4; int32_t* o64 = p.out64 + oBase * nctu;
            if (px == 1)
            {
                // lane = 16 px x 8 rows: two 8x8 PUs; 16x16 = lanes l, l^4; 32x32 = + l^1, l^8; 64x64 = the warp
                if (valid) *(int2*)(o8 + (int64_t)(y0 >> 3) * n8x + (x0 >> 3)) = make_int2((int)sa, (int)sb);
                int s16 = (int)(sa + sb);
                s16 += __shfl_xor_sync(0xffffffffu, s16, 4);
                if (valid && !(rg & 1)) o16[(int64_t)(y0 >> 4) * n16x + (x0 >> 4)] = s16;
                int s32 = s16 + __shfl_xor_sync(0xffffffffu, s16, 1);
                s32 += __shfl_xor_sync(0xffffffffu, s32, 8);
                if (valid && !(rg & 3) && !(cg & 1)) o32[(int64_t)(y0 >> 5) * n32x + (x0 >> 5)] = s32;
                int s64 = s32 + __shfl_xor_sync(0xffffffffu, s32, 2);
                s64 += __shfl_xor_sync(0xffffffffu, s64, 16);
                if (valid && lane == 0) o64[(int64_t)ctuY * p.ctuCols + (x0 >> 6)] = s64;
            }
            else
            {
                // lane = 8 px x 8 rows: one 8x8 PU; 16x16 = l, l^1, l^4, l^5; 32x32 = the strip's 4 x 4 lanes; 64x64 = two warps
                const int s8 = (int)(sa + sb);
                if (valid) o8[(int64_t)(y0 >> 3) * n8x + (x0 >> 3)] = s8;
                int s16 = s8 + __shfl_xor_sync(0xffffffffu, s8, 1);
                s16 += __shfl_xor_sync(0xffffffffu, s16, 4);
                if (valid && !(rg & 1) && !(cg & 1)) o16[(int64_t)(y0 >> 4) * n16x + (x0 >> 4)] = s16;
                int s32 = s16 + __shfl_xor_sync(0xffffffffu, s16, 2);
                s32 += __shfl_xor_sync(0xffffffffu, s32, 8);
                if (valid && !(rg & 3) && cg == 0) o32[(int64_t)(y0 >> 5) * n32x + (x0 >> 5)] = s32;
                const int sHalf = s32 + __shfl_xor_sync(0xffffffffu, s32, 16);          // 32 px x 64 rows
                // the two warps of a CTU meet at a named barrier of their own (the other pair keeps streaming); the exchange slot
                // alternates, so the pair's next meeting orders this read before the slot's next write
                int* slot = half64 + (meet++ & 1) * 4;
                if (lane == 0) slot[warp] = sHalf;
                asm volatile("bar.sync %0, 64;" :: "r"(1 + (warp >> 1)) : "memory");
                if (valid && lane == 0 && !(warp & 1)) o64[(int64_t)ctuY * p.ctuCols + (x0 >> 6)] = slot[warp] + slot[warp + 1];
            }
        }
    }
}

typedef CUresult (*SSEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int sad_stream_dev(Ctx* ctx, int depth, const void* poolOrigin, int64_t framePitch, int64_t stride, int marginX, int marginY, int rowsTotal, int numFrames,
                   int ctuCols, int ctuRows, const x265b200_sad_group* groupsHost, int numGroups, int numRefs,
                   int32_t* out8, int32_t* out16, int32_t* out32, int32_t* out64)
{
    if (numGroups <= 0 || numRefs <= 0 || ctuCols <= 0 || ctuRows <= 0) return 0;
    if (numRefs > SS_MAX_REFS) { set_error("sad_stream: at most %d references per group", SS_MAX_REFS); return -1; }
    if (depth != 8 && depth != 10 && depth != 12) { set_error("sad_stream: depth %d", depth); return -1; }
    const int px = depth > 8 ? 2 : 1;
    static SSEncodeTiledFn enc = []() -> SSEncodeTiledFn {
        void* f = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) return (SSEncodeTiledFn)f;
        return nullptr;
    }();
    if (!enc) { set_error("sad_stream: cuTensorMapEncodeTiled not available from the driver"); return -1; }
    const char* base = (const char*)poolOrigin - ((int64_t)marginY * stride + marginX) * px;
    if (((uintptr_t)base & 15) || ((stride * px) & 15) || ((framePitch * px) & 15) || ((marginX * px) & 15))
    { set_error("sad_stream: pool base, row stride, frame pitch and marginX must be multiples of 16 bytes (TMA)"); return -1; }
    if (((uintptr_t)out8 & 7)) { set_error("sad_stream: out8 must be 8-byte aligned"); return -1; }
    for (int g = 0; g < numGroups; g++)
    {
        if (groupsHost[g].cur < 0 || groupsHost[g].cur >= numFrames) { set_error("sad_stream: group %d: source frame %d outside the pool", g, groupsHost[g].cur); return -1; }
        for (int r = 0; r < numRefs; r++)
            if (groupsHost[g].ref[r] < 0 || groupsHost[g].ref[r] >= numFrames) { set_error("sad_stream: group %d: reference %d outside the pool", g, groupsHost[g].ref[r]); return -1; }
    }
    CUtensorMap map;
    cuuint64_t gdim[3] = { (cuuint64_t)stride, (cuuint64_t)rowsTotal, (cuuint64_t)numFrames };
    cuuint64_t gstr[2] = { (cuuint64_t)stride * px, (cuuint64_t)framePitch * px };
    cuuint32_t box[3] = { (cuuint32_t)(256 / px), 64, 1 };
    cuuint32_t estr[3] = { 1, 1, 1 };
    CUresult r = enc(&map, depth > 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)base, gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("sad_stream: cuTensorMapEncodeTiled failed (%d)", (int)r); return -1; }

    static_assert(sizeof(SSGroup) == sizeof(x265b200_sad_group), "group record layout");
    const size_t smem = (size_t)SS_STAGES * SS_TILE_BYTES + 2 * SS_STAGES * 8 + 64;
    if (depth > 8) X265B200_CHECK(cudaFuncSetAttribute(sad_stream_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else           X265B200_CHECK(cudaFuncSetAttribute(sad_stream_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t nctu = (int64_t)ctuCols * ctuRows;
    for (int g0 = 0; g0 < numGroups; g0 += SS_MAX_GROUPS)
    {
        SSArgs a;
        a.numGroups = std::min(SS_MAX_GROUPS, numGroups - g0); a.numRefs = numRefs;
        memcpy(a.groups, groupsHost + g0, sizeof(SSGroup) * (size_t)a.numGroups);
        a.ctuCols = ctuCols; a.ctuRows = ctuRows; a.marginX = marginX; a.marginY = marginY;
        a.chunks = (ctuCols * 64 * px + 255) / 256;
        const int64_t ob = (int64_t)g0 * numRefs * nctu;
        a.out8 = out8 + ob * 64; a.out16 = out16 + ob * 16; a.out32 = out32 + ob * 4; a.out64 = out64 + ob;
        const int64_t jobs = (int64_t)a.numGroups * ctuRows * a.chunks;
        const int grid = (int)std::min<int64_t>(jobs, (int64_t)ctx->smCount * 2);
        if (depth > 8) sad_stream_kernel<uint16_t><<<grid, 128, smem, ctx->stream>>>(map, a);
        else           sad_stream_kernel<uint8_t><<<grid, 128, smem, ctx->stream>>>(map, a);
        ctx->launches++;
    }
    return check(cudaGetLastError(), "sad_stream launch");
}

} // namespace x265b200

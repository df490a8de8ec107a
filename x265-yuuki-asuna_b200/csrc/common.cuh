// common.cuh -- shared device helpers for the x265 B200 primitive backend (sm_100a only).
//
// Conventions used by every kernel in csrc/:
//   * `pixel` is uint8_t (8-bit build, X265_DEPTH=8) or uint16_t (HIGH_BIT_DEPTH, depth 10/12),
//     mirroring source/common/common.h:126-142 of the reference.
//   * strides and offsets are in ELEMENTS (pixels / int16 coefficients), never bytes.
//   * reference-plane reads may start at any pixel (motion vectors), so all block loads go
//     through the ld_px4 helpers that build an unaligned 4-pixel group out of aligned words.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define X265B200_WARP 32

namespace x265b200 {

struct Ctx
{
    int          device;
    cudaStream_t stream;
    bool         ownsStream;
    int          smCount;
    // scratch for host-pointer entry points (grown on demand)
    void*        dScratch[8];
    size_t       dScratchBytes[8];
    void*        hPinned[8];
    size_t       hPinnedBytes[8];
    // kernel launch counter (bench.py reports it as gpu_launches)
    unsigned long long launches;
    // bit-cost tables cached per lambda (me_kernels.cu)
    uint16_t*    dMvCost;
    double       mvCostLambda;
    bool         mvCostValid;
    // small host -> device parameter uploads (job lists, pointer tables) without a stream-wide sync: a ring of pinned staging
    // buffers, each guarded by the event of its last copy (stage_small, capi.cu)
    void*        hStage[4];
    size_t       hStageBytes[4];
    cudaEvent_t  hStageEv[4];
    int          hStageNext;
    // host-buffer frame searches (x265b200_me_frame*_host_begin / _host_end): a copy stream next to the compute stream, and the
    // events that order H2D -> search -> D2H across them
    cudaStream_t copyStream;
    cudaEvent_t  evH2D, evSearch, evD2H[2];
    int          hostHead, hostCount;      // ring of pending host calls (at most two: the caller may queue the next frame before it reads this one)
};

void set_error(const char* fmt, ...);
int  stage_small(Ctx* ctx, void* dDst, const void* hostSrc, size_t bytes);
int  check(cudaError_t e, const char* what);

#define X265B200_CHECK(expr) do { if (x265b200::check((expr), #expr)) return -1; } while (0)

// ---- device helpers ------------------------------------------------------------------------

__device__ __forceinline__ int warp_sum(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 4 consecutive 8-bit pixels starting at any byte address, packed little-endian in a u32.
__device__ __forceinline__ uint32_t ld_px4(const uint8_t* p)
{
    uintptr_t a = (uintptr_t)p;
    const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
    uint32_t sh = (uint32_t)(a & 3);
    uint32_t lo = __ldg(w);
    if (sh == 0) return lo;
    uint32_t hi = __ldg(w + 1);
    return __funnelshift_r(lo, hi, sh * 8);
}

// 2 consecutive 16-bit pixels starting at any 2-byte aligned address, packed in a u32.
__device__ __forceinline__ uint32_t ld_px2(const uint16_t* p)
{
    uintptr_t a = (uintptr_t)p;
    const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
    uint32_t lo = __ldg(w);
    if ((a & 2) == 0) return lo;
    uint32_t hi = __ldg(w + 1);
    return __funnelshift_r(lo, hi, 16);
}

// Load 4 pixels as ints (generic over pixel type).
template<typename pixel> __device__ __forceinline__ void ld4i(const pixel* p, int v[4]);
template<> __device__ __forceinline__ void ld4i<uint8_t>(const uint8_t* p, int v[4])
{
    uint32_t x = ld_px4(p);
    v[0] = x & 0xff; v[1] = (x >> 8) & 0xff; v[2] = (x >> 16) & 0xff; v[3] = x >> 24;
}
template<> __device__ __forceinline__ void ld4i<uint16_t>(const uint16_t* p, int v[4])
{
    uint32_t x = ld_px2(p), y = ld_px2(p + 2);
    v[0] = x & 0xffff; v[1] = x >> 16; v[2] = y & 0xffff; v[3] = y >> 16;
}
// int16 coefficients / residuals (2-byte aligned)
__device__ __forceinline__ void ld4s(const int16_t* p, int v[4])
{
    uint32_t x = ld_px2((const uint16_t*)p), y = ld_px2((const uint16_t*)p + 2);
    v[0] = (int16_t)(x & 0xffff); v[1] = (int16_t)(x >> 16); v[2] = (int16_t)(y & 0xffff); v[3] = (int16_t)(y >> 16);
}

// SAD of the two 16-bit pixels of a word pair.  __vsadu2 is emulated on sm_100a (~10 instructions: PRMT / IMAD / IABS per lane,
// scripts' SASS count of the first 16-bit streaming kernel); per-halfword max - min cannot borrow across the halves, so
// VIMNMX.U16x2 x2 + IADD + IDP.2A (both halves summed by a dot product with 1, 1) gives the same number in 4.
__device__ __forceinline__ uint32_t sad_u16x2(uint32_t a, uint32_t b)
{
    return __dp2a_lo(__vmaxu2(a, b) - __vminu2(a, b), 0x0101u, 0u);
}

// sum of absolute differences of 4 pixels
template<typename pixel> __device__ __forceinline__ int sad4(const pixel* a, const pixel* b);
template<> __device__ __forceinline__ int sad4<uint8_t>(const uint8_t* a, const uint8_t* b)
{
    return (int)__vsadu4(ld_px4(a), ld_px4(b));
}
template<> __device__ __forceinline__ int sad4<uint16_t>(const uint16_t* a, const uint16_t* b)
{
    return (int)(sad_u16x2(ld_px2(a), ld_px2(b)) + sad_u16x2(ld_px2(a + 2), ld_px2(b + 2)));
}

__device__ __forceinline__ int clip3i(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

} // namespace x265b200

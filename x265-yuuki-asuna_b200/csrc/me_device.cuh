// me_device.cuh -- warp-cooperative device implementation of x265's MotionEstimate::motionEstimate
// (source/encoder/motion.cpp:739-1569) and subpelCompare (:1571-1664).  One warp runs one PU
// search; every lane executes the same control flow on warp-uniform values (costs are reduced
// with warp shuffles and broadcast), so the data-dependent search path of the reference is
// reproduced step for step, including its quirks (SURVEY.md 8a "quirks").
//
// Shared by me_kernels.cu (full-resolution batched ME) and lookahead_kernels.cu (lowres path).
#pragma once
#ifdef ME_HOST_EMU                      /* tests/host_emu: the same source compiled for the host (test infrastructure) */
#include "me_host_emu.h"
#else
#include "common.cuh"
#endif
#include "tables.cuh"
#include "satd_packed.cuh"
#include "subpel_packed.cuh"
#include <type_traits>

// A translation unit whose searches ALL run in per-thread mode defines ME_FORCE_THREAD before including this header:
// the warp-cooperative bodies are then compiled out (smaller kernel, see me_frame_kernels.cu).  The variant lives in
// its own inline namespace so the differently-compiled templates never share a symbol.
// A thread-only translation unit may also fix the plane kind at compile time: ME_LOWRES_ONLY (the lookahead: every search
// runs on the 4 hpel planes) or ME_FULLRES_ONLY (the frame search) drops the other sub-pel machinery from the kernel, which
// keeps its register allocation independent of code it never executes.
#ifdef ME_FORCE_THREAD
#define ME_IS_THREAD(s) true
#ifdef ME_LANE_W_VAR
#define ME_PU_W(s) ((s).w)                /* a lane searches a 4- or 8-pixel-wide sub-block (rect / AMP PUs, me_ctu_kernels.cu) */
#else
#define ME_PU_W(s) 8                      /* every lane searches an 8-pixel-wide sub-block */
#endif
#if defined(ME_CTU_KERNEL)
#define ME_VARIANT me_thread_ctu
#define ME_IS_LOWRES(s) false
#elif defined(ME_LOWRES_ONLY)
#define ME_VARIANT me_thread_lowres
#define ME_IS_LOWRES(s) true
#elif defined(ME_FULLRES_ONLY)
#define ME_VARIANT me_thread_fullres
#define ME_IS_LOWRES(s) false
#else
#define ME_VARIANT me_thread_only
#endif
#else
#define ME_IS_THREAD(s) ((s).perThread)
#define ME_PU_W(s) ((s).w)
#define ME_VARIANT me_generic
#endif
#ifndef ME_IS_LOWRES
#define ME_IS_LOWRES(s) ((s).isLowres)
#endif

namespace x265b200 {
inline namespace ME_VARIANT {

enum { ME_DIA = 0, ME_HEX = 1, ME_UMH = 2, ME_STAR = 3, ME_SEA = 4, ME_FULL = 5,   // x265.h:492-497
       ME_REFINE = 6 };   // not a search method of the reference: selects MotionEstimate::refineMV (motion.cpp:606-737)
#define ME_COST_MAX (1 << 28)                                                       // motion.h:65

struct MV2 { int x, y; };
__device__ __forceinline__ MV2 mv2(int x, int y) { MV2 m; m.x = x; m.y = y; return m; }

// subpel workloads, motion.cpp:48-58
struct SubpelWL { int hpel_iters, hpel_dirs, qpel_iters, qpel_dirs, hpel_satd; };
__constant__ SubpelWL c_workload[8] = {
    { 1, 4, 0, 4, 0 }, { 1, 4, 1, 4, 0 }, { 1, 4, 1, 4, 1 }, { 2, 4, 1, 4, 1 },
    { 2, 4, 2, 4, 1 }, { 1, 8, 1, 8, 1 }, { 2, 8, 1, 8, 1 }, { 2, 8, 2, 8, 1 } };
// pattern tables, motion.cpp:64-84
__constant__ int8_t c_hex2[8][2]    = { {-1,-2}, {-2,0}, {-1,2}, {1,2}, {2,0}, {1,-2}, {-1,-2}, {-2,0} };
__constant__ uint8_t c_mod6m1[8]    = { 5, 0, 1, 2, 3, 4, 5, 0 };
__constant__ int8_t c_square1[9][2] = { {0,0}, {0,-1}, {0,1}, {-1,0}, {1,0}, {-1,-1}, {-1,1}, {1,-1}, {1,1} };
__constant__ int8_t c_hex4[16][2]   = { {0,-4}, {0,4}, {-2,-3}, {2,-3}, {-4,-2}, {4,-2}, {-4,-1}, {4,-1},
                                        {-4,0}, {4,0}, {-4,1}, {4,1}, {-4,2}, {4,2}, {-2,3}, {2,3} };
__constant__ int8_t c_offsets[16][2] = { {-1,0}, {0,-1}, {-1,-1}, {1,-1}, {-1,0}, {1,0}, {-1,1}, {-1,-1},
                                         {1,-1}, {1,1}, {-1,0}, {0,1}, {-1,1}, {1,1}, {1,0}, {0,1} };
__constant__ uint8_t c_rangeMul[4][4] = { {3,3,4,4}, {3,4,4,4}, {4,4,4,5}, {4,4,5,6} };
__constant__ int16_t c_meLumaFilter[4][8] = {
    { 0, 0, 0, 64, 0, 0, 0, 0 }, { -1, 4, -10, 58, 17, -5, 1, 0 },
    { -1, 4, -11, 40, 40, -11, 4, -1 }, { 0, 1, -5, 17, 58, -10, 4, -1 } };

template<typename pixel>
struct MEState
{
    // per-warp shared memory
    pixel*   fenc;       // cached PU, row stride 64 (FENC_STRIDE), motion.cpp:189
    pixel*   pred;       // subpel prediction, row stride = w (motion.cpp:1581 subpelbuf)
    int16_t* immed;      // hvpp intermediate, w * (h + 7)
    // reference
    const pixel* fref;   // fpelPlane[0] + blockOffset (global plane, or the TMA-staged shared-memory window)
    int64_t  stride;
    const pixel* gfref;  // the same block in the GLOBAL plane (for the zero-MV candidate, which may lie outside a window)
    int64_t  gstride;
    const pixel* lowres[4];   // lowres hpel planes + blockOffset (isLowres path), else unused
    bool     isLowres;
    bool     perThread;   // true: ONE THREAD runs the whole search on its (sub-)block; false: one warp (lanes cooperate)
    int      groupSize;   // perThread only: 1, or 2/4 lanes that each own a sub-block of the PU and sum their costs
    unsigned groupMask;   // lanes of this group (shuffle mask)
    int      w, h, lane, depth, partSizeScale;
    // chroma residual cost of subpelCompare (bChromaSATD, motion.cpp:212,1601-1661); warp-cooperative searches only
    bool     chromaSatd;
    const pixel* fencC[2];    // cached Cb / Cr PU, row stride csize (= 64 >> hshift, Yuv::m_csize)
    const pixel* frefC[2];    // Cb / Cr reference planes at the PU position (ReferencePlanes::getCbAddr/getCrAddr)
    int64_t  strideC;
    int      csize, hshift, vshift;
    const uint16_t* cost;     // centred lambda-scaled MV cost table (bitcost.cpp:31-60)
    int      mvpx, mvpy;      // setMVP(qmvp), bitcost.h:41
    // --me sea (motion.cpp:1242-1395): MotionEstimate::integral[12] of this reference (plane pointers addressed like the
    // reference plane) and the PU's element offset in them (search.cpp:2264); warp-cooperative searches only
    const uint32_t* const* integral;
    int64_t  integralOff;
#ifdef ME_WINDOW_CHECK
    // Full-pel MVs (relative to the PU) for which the whole PU block lies inside the staged shared-memory window, inclusive;
    // anything else is read from the global plane through gfref (predictors far from the CTU's window centre, the zero-MV
    // candidate, overshoot of the pattern searches).  Sub-pel reads shrink the rectangle by the filter footprint.
    int      winX0, winX1, winY0, winY1;
#endif
#ifdef ME_COST_SMEM
    const uint16_t* costS;    // cost[-costK .. costK] staged in shared memory, centred like `cost`
    int      costK;
#endif
#ifdef ME_THREAD_CHROMA
    // per-thread chroma term: the lane's Cb / Cr sub-block in the staged chroma windows (frefC, strideC) and in the global
    // planes (gfrefC, gstrideC); the window rectangle in whole chroma samples, inclusive, footprint of the 4-tap filter included
    const pixel* gfrefC[2];
    int64_t  gstrideC;
    int      cwinX0, cwinX1, cwinY0, cwinY1;
#endif
};

// The cached source PU (fenc) lives in shared memory in every kernel that uses this header: say so at the load sites.
template<typename T> __device__ __forceinline__ const T* smem_hint(const T* p)
{
#ifdef ME_HOST_EMU
    if ((uintptr_t)p % sizeof(T)) abort();      // the host emulation enforces the natural alignment the GPU faults on
#endif
    __builtin_assume(__isShared(p));
    return p;
}

constexpr int kMvTableHalf = 2 * 32768;

template<typename pixel>
__device__ __forceinline__ int mvcost(const MEState<pixel>& s, int qx, int qy)
{
    int ix = clip3i(-kMvTableHalf, kMvTableHalf, qx - s.mvpx);
    int iy = clip3i(-kMvTableHalf, kMvTableHalf, qy - s.mvpy);
#ifdef ME_COST_SMEM
    // the differences a search produces stay within a few multiples of merange: those entries sit in shared memory
    // (one LDS instead of an L2 round trip on the critical path of every candidate); the rest comes from the full table
    const int K = s.costK;
    const int cx = (ix >= -K && ix <= K) ? (int)smem_hint(s.costS)[ix] : (int)s.cost[ix];
    const int cy = (iy >= -K && iy <= K) ? (int)smem_hint(s.costS)[iy] : (int)s.cost[iy];
    return (cx + cy) & 0xffff;
#else
    return ((int)s.cost[ix] + (int)s.cost[iy]) & 0xffff;      // bitcost.h:45 returns uint16_t
#endif
}


// ---- row loaders ---------------------------------------------------------------------------------
// NW consecutive 32-bit words of pixels starting at ANY pixel address (generic pointer: global plane or
// shared-memory window).  Loads NW+1 aligned words and funnel-shifts.
// A translation unit whose reference blocks always sit in the shared-memory window defines ME_REF_IN_SMEM: the loader
// then tells the compiler so (LDS with 32-bit addresses instead of generic LD + address-space resolution).
template<typename pixel, int NW, bool ANYSPACE = false>
__device__ __forceinline__ void ld_words(const pixel* p, uint32_t out[NW])
{
    uintptr_t a = (uintptr_t)p;
    const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
#ifdef ME_REF_IN_SMEM
    if (!ANYSPACE) __builtin_assume(__isShared(w));
#endif
    const uint32_t sh = (uint32_t)(a & 3) * 8;
    uint32_t t[NW + 1];
#pragma unroll
    for (int i = 0; i < NW; i++) t[i] = w[i];
    t[NW] = sh ? w[NW] : 0u;             // predicated: never touches the word after an aligned run
#pragma unroll
    for (int i = 0; i < NW; i++) out[i] = __funnelshift_r(t[i], t[i + 1], sh);
}

#if defined(ME_REF_IN_SMEM) && defined(ME_WINDOW_FAST)
#define ME_SMEM_FAST 1
// Rows of the shared-memory window through 32-bit shared addresses.  The window pitch is a multiple of 16 bytes
// (me_frame_dev), so the byte misalignment of a block is the same on every row: a caller resolves the pointer ONCE into
// (4-byte-aligned shared address, bit shift) and then walks rows / 4-pixel columns with 32-bit adds -- the generic form
// above spends as many instructions on 64-bit row addresses, alignment and the predicate of the last word as on the
// loads themselves (profiles/r01_me_frame_v8.txt: ld_words = 19 % of the executed instructions).  NW + 1 words are always
// read: the word after an aligned run is ignored by the funnel shift and lies inside the CTA's allocation (the window
// is followed by the fenc tile).
template<int OFF> __device__ __forceinline__ uint32_t lds_u32(uint32_t a)
{
#ifdef ME_HOST_EMU
    return *(const uint32_t*)(emu::smem_base + a + OFF);
#else
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF));
    return v;
#endif
}
template<int N, int I = 0> __device__ __forceinline__ void lds_run(uint32_t a, uint32_t* t)
{
    if constexpr (I < N) { t[I] = lds_u32<4 * I>(a); lds_run<N, I + 1>(a, t); }
}
struct SRow { uint32_t a, sh; };
template<typename pixel> __device__ __forceinline__ SRow srow(const pixel* p)
{
    const uint32_t g = (uint32_t)__cvta_generic_to_shared(p);
    SRow r; r.a = g & ~3u; r.sh = (g & 3u) * 8u;
    return r;
}
template<int NW> __device__ __forceinline__ void lds_words(uint32_t a, uint32_t sh, uint32_t out[NW])
{
    uint32_t t[NW + 1];
    lds_run<NW + 1>(a, t);
#pragma unroll
    for (int i = 0; i < NW; i++) out[i] = __funnelshift_r(t[i], t[i + 1], sh);
}
#endif

template<typename pixel> __device__ __forceinline__ uint32_t sad_word(uint32_t a, uint32_t b);
template<> __device__ __forceinline__ uint32_t sad_word<uint8_t>(uint32_t a, uint32_t b) { return __vsadu4(a, b); }
template<> __device__ __forceinline__ uint32_t sad_word<uint16_t>(uint32_t a, uint32_t b) { return sad_u16x2(a, b); }      // common.cuh: 4 native instructions

// SAD of one SEG-pixel row segment: fenc in smem (aligned), ref anywhere
template<typename pixel, int SEG>
__device__ __forceinline__ int sad_seg(const pixel* f, const pixel* r)
{
    constexpr int NW = SEG * (int)sizeof(pixel) / 4;
    uint32_t rw[NW];
    ld_words<pixel, NW>(r, rw);
    const uint32_t* fw = smem_hint((const uint32_t*)f);
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < NW; i++) acc += sad_word<pixel>(fw[i], rw[i]);
    return (int)acc;
}

// K (1..4) candidate SADs of the cached PU at full-pel offsets (ox[k], oy[k]) from s.fref, all at once:
// lanes are spread over candidates x rows x SEG-pixel segments, per-lane partial sums are combined with a
// 10-shuffle transpose-reduce and every lane receives all K totals.
template<typename pixel, int SEG>
__device__ __forceinline__ void sad_k_impl(const MEState<pixel>& s, int K, const int ox[4], const int oy[4], int costs[4])
{
    const int segs = ME_PU_W(s) / SEG, upc = s.h * segs, total = K * upc;
    int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    int cand = 0, rem = s.lane;
    while (rem >= upc) { rem -= upc; cand++; }
    for (int u = s.lane; u < total; u += 32)
    {
        const int y = rem / segs, x = (rem - y * segs) * SEG;
        const pixel* r = s.fref + (ox[cand & 3] + x) + (int64_t)(oy[cand & 3] + y) * s.stride;
        const int v = sad_seg<pixel, SEG>(s.fenc + y * 64 + x, r);
        a0 += cand == 0 ? v : 0; a1 += cand == 1 ? v : 0; a2 += cand == 2 ? v : 0; a3 += cand == 3 ? v : 0;
        rem += 32;
        while (rem >= upc) { rem -= upc; cand++; }
    }
    if (K == 1) { costs[0] = warp_sum(a0); return; }
    const bool b4 = s.lane & 16, b3 = s.lane & 8;
    int b0 = b4 ? a2 : a0, s0 = b4 ? a0 : a2; b0 += __shfl_xor_sync(0xffffffffu, s0, 16);
    int b1 = b4 ? a3 : a1, s1 = b4 ? a1 : a3; b1 += __shfl_xor_sync(0xffffffffu, s1, 16);
    int c = b3 ? b1 : b0, t = b3 ? b0 : b1; c += __shfl_xor_sync(0xffffffffu, t, 8);
    c += __shfl_xor_sync(0xffffffffu, c, 4);
    c += __shfl_xor_sync(0xffffffffu, c, 2);
    c += __shfl_xor_sync(0xffffffffu, c, 1);
    // lanes 0-7 hold cand 0, 8-15 cand 1, 16-23 cand 2, 24-31 cand 3
    costs[0] = __shfl_sync(0xffffffffu, c, 0);
    costs[1] = __shfl_sync(0xffffffffu, c, 8);
    costs[2] = __shfl_sync(0xffffffffu, c, 16);
    costs[3] = __shfl_sync(0xffffffffu, c, 24);
}


__device__ __forceinline__ void me_hadamard4(int& a, int& b, int& c, int& d)
{
    int t0 = a + b, t1 = a - b, t2 = c + d, t3 = c - d;
    a = t0 + t2; c = t0 - t2; b = t1 + t3; d = t1 - t3;
}

template<typename pixel> __device__ __forceinline__ void unpack4(const uint32_t* w, int v[4]);
template<> __device__ __forceinline__ void unpack4<uint8_t>(const uint32_t* w, int v[4])
{
    v[0] = w[0] & 0xff; v[1] = (w[0] >> 8) & 0xff; v[2] = (w[0] >> 16) & 0xff; v[3] = w[0] >> 24;
}
template<> __device__ __forceinline__ void unpack4<uint16_t>(const uint32_t* w, int v[4])
{
    v[0] = w[0] & 0xffff; v[1] = w[0] >> 16; v[2] = w[1] & 0xffff; v[3] = w[1] >> 16;
}


// ---- per-thread primitives (perThread mode: no full-warp shuffles, no __syncwarp; used for 8x8 / 16x16 PUs) ----
// A PU may be split over groupSize = 2 or 4 adjacent lanes, each owning a sub-block (SAD and SATD are additive over
// sub-blocks of 4x4 cells): the lanes of a group compute identical totals, so they follow the same control flow.
template<typename pixel>
__device__ __forceinline__ int group_sum(const MEState<pixel>& s, int v)
{
    for (int o = 1; o < s.groupSize; o <<= 1) v += __shfl_xor_sync(s.groupMask, v, o);
    return v;
}

template<typename pixel, int SEG>
__device__ __forceinline__ int thread_sad_one(const MEState<pixel>& s, const pixel* r, int64_t rs)
{
    int acc = 0;
    for (int y = 0; y < s.h; y++)
        for (int x = 0; x < ME_PU_W(s); x += SEG)
            acc += sad_seg<pixel, SEG>(s.fenc + y * 64 + x, r + (int64_t)y * rs + x);
    return acc;
}
// 8-pixel-wide sub-block (the me_frame layout): 8 rows at a time with every load issued before the first use, so one
// candidate costs one load latency per 8 rows instead of one per row; fenc rows are 8-pixel aligned (one vector load).
// Heights that are not a multiple of 8 (12- and 24-row AMP pieces, 4-row tails) finish with a batch of 4 rows.
template<typename pixel>
__device__ __forceinline__ int thread_sad_w8(const MEState<pixel>& s, const pixel* r, int64_t rs)
{
    constexpr int NW = 8 * (int)sizeof(pixel) / 4;
    typedef typename std::conditional<sizeof(pixel) == 1, uint2, uint4>::type fvec;
    uint32_t acc = 0;
#ifdef ME_SMEM_FAST
    const SRow base = srow(r);
    const uint32_t rsB = (uint32_t)rs * (uint32_t)sizeof(pixel);
#endif
    int y0 = 0;
#pragma unroll 1
    for (; y0 + 8 <= s.h; y0 += 8)
    {
        uint32_t rw[8][NW];
        fvec f[8];
#pragma unroll
        for (int y = 0; y < 8; y++)
        {
#ifdef ME_SMEM_FAST
            lds_words<NW>(base.a + (uint32_t)(y0 + y) * rsB, base.sh, rw[y]);
#else
            ld_words<pixel, NW>(r + (int64_t)(y0 + y) * rs, rw[y]);
#endif
            f[y] = *smem_hint((const fvec*)(s.fenc + (y0 + y) * 64));
        }
#pragma unroll
        for (int y = 0; y < 8; y++)
        {
            const uint32_t* fw = (const uint32_t*)&f[y];
#pragma unroll
            for (int i = 0; i < NW; i++) acc += sad_word<pixel>(fw[i], rw[y][i]);
        }
    }
#ifdef ME_LANE_W_VAR
    if (y0 < s.h)
    {
        uint32_t rw[4][NW];
        fvec f[4];
#pragma unroll
        for (int y = 0; y < 4; y++)
        {
#ifdef ME_SMEM_FAST
            lds_words<NW>(base.a + (uint32_t)(y0 + y) * rsB, base.sh, rw[y]);
#else
            ld_words<pixel, NW>(r + (int64_t)(y0 + y) * rs, rw[y]);
#endif
            f[y] = *smem_hint((const fvec*)(s.fenc + (y0 + y) * 64));
        }
#pragma unroll
        for (int y = 0; y < 4; y++)
        {
            const uint32_t* fw = (const uint32_t*)&f[y];
#pragma unroll
            for (int i = 0; i < NW; i++) acc += sad_word<pixel>(fw[i], rw[y][i]);
        }
    }
#endif
    return (int)acc;
}
// per-thread SAD of the sub-block against a block in ANY address space (row by row; rare path: the zero-MV candidate of
// the frame search, and every block outside the staged window when ME_WINDOW_CHECK is on)
template<typename pixel>
__device__ __noinline__ int thread_sad_anyspace(const MEState<pixel>& s, const pixel* r, int64_t rs)
{
    constexpr int NW = 4 * (int)sizeof(pixel) / 4;
    uint32_t acc = 0;
#pragma unroll 1
    for (int y = 0; y < s.h; y++)
#pragma unroll 1
        for (int x = 0; x < ME_PU_W(s); x += 4)
        {
            uint32_t rw[NW];
            ld_words<pixel, NW, true>(r + (int64_t)y * rs + x, rw);
            const uint32_t* fw = smem_hint((const uint32_t*)(s.fenc + y * 64 + x));
#pragma unroll
            for (int i = 0; i < NW; i++) acc += sad_word<pixel>(fw[i], rw[i]);
        }
    return (int)acc;
}
template<typename pixel>
__device__ __forceinline__ int thread_sad_any(const MEState<pixel>& s, const pixel* r, int64_t rs)
{
#ifdef ME_LANE_W_VAR
    if (ME_PU_W(s) == 8) return thread_sad_w8<pixel>(s, r, rs);
    return thread_sad_one<pixel, 4>(s, r, rs);
#else
    if (ME_PU_W(s) == 8 && !(s.h & 7)) return thread_sad_w8<pixel>(s, r, rs);
    if (!(ME_PU_W(s) & 15)) return thread_sad_one<pixel, 16>(s, r, rs);
    if (!(ME_PU_W(s) & 7))  return thread_sad_one<pixel, 8>(s, r, rs);
    return thread_sad_one<pixel, 4>(s, r, rs);
#endif
}
#ifdef ME_WINDOW_CHECK
// is the PU block at full-pel MV (mx, my), grown by `lt` pixels left / above and `rb` right / below, inside the staged window?
template<typename pixel>
__device__ __forceinline__ bool in_window(const MEState<pixel>& s, int mx, int my, int lt, int rb)
{
    return (mx >= s.winX0 + lt) & (mx <= s.winX1 - rb) & (my >= s.winY0 + lt) & (my <= s.winY1 - rb);
}
#endif
// 4x4 Hadamard cost of d[i][k] = a - b rows already differenced: sum |H d H^T| >> 1
__device__ __forceinline__ int satd_cell(int d[4][4])
{
#pragma unroll
    for (int i = 0; i < 4; i++) me_hadamard4(d[i][0], d[i][1], d[i][2], d[i][3]);
    int t = 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
        me_hadamard4(d[0][k], d[1][k], d[2][k], d[3][k]);
        t += abs(d[0][k]) + abs(d[1][k]) + abs(d[2][k]) + abs(d[3][k]);
    }
    return t >> 1;
}

template<typename pixel>
__device__ __forceinline__ int thread_satd(const MEState<pixel>& s, const pixel* r, int64_t rs)
{
    int acc = 0;
    constexpr int NW = 4 * (int)sizeof(pixel) / 4;
#ifdef ME_SMEM_FAST
    const SRow base = srow(r);
    const uint32_t rsB = (uint32_t)rs * (uint32_t)sizeof(pixel);
#pragma unroll 1
    for (int cy = 0; cy < s.h; cy += 4)
#pragma unroll 1
        for (int cx = 0; cx < ME_PU_W(s); cx += 4)
        {
            const uint32_t a0 = base.a + (uint32_t)cy * rsB + (uint32_t)cx * (uint32_t)sizeof(pixel);
            if constexpr (sizeof(pixel) == 1)
            {
                uint32_t fw[4], ow[4];
#pragma unroll
                for (int i = 0; i < 4; i++)
                {
                    fw[i] = *smem_hint((const uint32_t*)(s.fenc + (cy + i) * 64 + cx));
                    lds_words<1>(a0 + (uint32_t)i * rsB, base.sh, &ow[i]);
                }
                acc += satd4x4_packed_u8(fw, ow);
            }
            else
            {
                int d[4][4];
#pragma unroll
                for (int i = 0; i < 4; i++)
                {
                    int a[4], b[4];
                    uint32_t rw[NW];
                    unpack4<pixel>(smem_hint((const uint32_t*)(s.fenc + (cy + i) * 64 + cx)), a);
                    lds_words<NW>(a0 + (uint32_t)i * rsB, base.sh, rw);
                    unpack4<pixel>(rw, b);
#pragma unroll
                    for (int k = 0; k < 4; k++) d[i][k] = a[k] - b[k];
                }
                acc += satd_cell(d);
            }
        }
    return acc;
#else
#if defined(ME_PACKED_SATD) && defined(ME_LOWRES_ONLY) && !defined(ME_LA_SATD8_OFF)
    // the lookahead's 8x8 CU at 8 bits: one pass over the eight rows (each row loaded once for its two 4x4 cells), fully unrolled
    if constexpr (sizeof(pixel) == 1)
    {
        if (ME_PU_W(s) == 8 && s.h == 8)
        {
            uint32_t rw[8][2];
            uint2 fw[8];
#pragma unroll
            for (int i = 0; i < 8; i++)
            {
                ld_words<pixel, 2>(r + (int64_t)i * rs, rw[i]);
                fw[i] = *smem_hint((const uint2*)(s.fenc + i * 64));
            }
#pragma unroll
            for (int q = 0; q < 2; q++)
            {
                uint32_t fl[4], fr[4], ol[4], orr[4];
#pragma unroll
                for (int i = 0; i < 4; i++) { fl[i] = fw[4 * q + i].x; fr[i] = fw[4 * q + i].y; ol[i] = rw[4 * q + i][0]; orr[i] = rw[4 * q + i][1]; }
                acc += satd4x4_packed_u8(fl, ol) + satd4x4_packed_u8(fr, orr);
            }
            return acc;
        }
    }
#endif
#ifdef ME_PACKED_SATD
    // the lookahead kernel: the packed-word SATD on rows from any address space (-2 % per launch, profiles/r02_staged_ab.txt)
    if constexpr (sizeof(pixel) == 1)
    {
#pragma unroll 1
        for (int cy = 0; cy < s.h; cy += 4)
#pragma unroll 1
            for (int cx = 0; cx < ME_PU_W(s); cx += 4)
            {
                uint32_t fw[4], ow[4];
#pragma unroll
                for (int i = 0; i < 4; i++)
                {
                    fw[i] = *smem_hint((const uint32_t*)(s.fenc + (cy + i) * 64 + cx));
                    ld_words<pixel, 1>(r + (int64_t)(cy + i) * rs + cx, &ow[i]);
                }
                acc += satd4x4_packed_u8(fw, ow);
            }
    }
    else
#endif
    {
#pragma unroll 1
    for (int cy = 0; cy < s.h; cy += 4)
#pragma unroll 1
        for (int cx = 0; cx < ME_PU_W(s); cx += 4)
        {
            int d[4][4];
#pragma unroll
            for (int i = 0; i < 4; i++)
            {
                int a[4], b[4];
                uint32_t rw[NW];
                unpack4<pixel>(smem_hint((const uint32_t*)(s.fenc + (cy + i) * 64 + cx)), a);
                ld_words<pixel, NW>(r + (int64_t)(cy + i) * rs + cx, rw);
                unpack4<pixel>(rw, b);
#pragma unroll
                for (int k = 0; k < 4; k++) d[i][k] = a[k] - b[k];
            }
            acc += satd_cell(d);
        }
    }
    return acc;
#endif
}

__device__ __forceinline__ uint32_t pack_cand(int x, int y) { return ((uint32_t)x & 0xffffu) | ((uint32_t)y << 16); }
#ifdef ME_WINDOW_CHECK
// Frame-search form (me_ctu_kernels.cu): ONE out-of-line function costs the K (1..4) full-pel candidates of a search step.
// The candidate offsets arrive packed in registers ((x & 0xffff) | (y << 16); as ox[] / oy[] arrays of a __noinline__ function
// they lived in local memory and every candidate began with two dependent LDL -- 7 % of the stall samples of
// profiles/r02_me_frame_v11.txt), each candidate is read from the staged window or, outside it, from the plane, the K partial
// SADs are reduced over the PU's lanes together, and the MV cost of each candidate is added here -- every caller wants
// SAD + mvcost, and the callers (hexSearch, squareRefine, starPattern ...) stay small enough for the instruction caches
// (inlined, this code made them 5-13 KB each and stall_no_instruction rose from 0.23 to 1.75 per issue,
// profiles/r02_me_ctu_v1.txt).
template<typename pixel>
__device__ __noinline__ int4 thread_cand_costs(const MEState<pixel>& s, int K, uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3, bool addMvCost)
{
    int part[4] = { 0, 0, 0, 0 };
#pragma unroll 1
    for (int k = 0; k < K; k++)
    {
        const uint32_t pk = k == 0 ? p0 : (k == 1 ? p1 : (k == 2 ? p2 : p3));
        const int mx = (int)(int16_t)(pk & 0xffffu), my = (int)pk >> 16;
        int v;
        if (in_window(s, mx, my, 0, 0)) v = thread_sad_any<pixel>(s, s.fref + mx + (int64_t)my * s.stride, s.stride);
        else v = thread_sad_anyspace<pixel>(s, s.gfref + mx + (int64_t)my * s.gstride, s.gstride);
#pragma unroll
        for (int j = 0; j < 4; j++) part[j] = j == k ? v : part[j];      // stays in registers (no dynamic indexing)
    }
    for (int o = 1; o < s.groupSize; o <<= 1)
    {
#pragma unroll
        for (int k = 0; k < 4; k++) part[k] += __shfl_xor_sync(s.groupMask, part[k], o);
    }
    if (addMvCost)
    {
#pragma unroll 1
        for (int k = 0; k < K; k++)
        {
            const uint32_t pk = k == 0 ? p0 : (k == 1 ? p1 : (k == 2 ? p2 : p3));
            const int c = mvcost(s, (int)(int16_t)(pk & 0xffffu) << 2, ((int)pk >> 16) << 2);
#pragma unroll
            for (int j = 0; j < 4; j++) part[j] += j == k ? c : 0;
        }
    }
    return make_int4(part[0], part[1], part[2], part[3]);
}
// the array form the search code uses; addMvCost = false leaves the plain SADs
template<typename pixel>
__device__ __forceinline__ void warp_sad_k(const MEState<pixel>& s, int K, const int ox[4], const int oy[4], int costs[4], bool addMvCost = false)
{
    const int4 c = thread_cand_costs<pixel>(s, K, pack_cand(ox[0], oy[0]), K > 1 ? pack_cand(ox[1], oy[1]) : 0u, K > 2 ? pack_cand(ox[2], oy[2]) : 0u,
                                            K > 3 ? pack_cand(ox[3], oy[3]) : 0u, addMvCost);
    costs[0] = c.x;
    if (K > 1) costs[1] = c.y;
    if (K > 2) costs[2] = c.z;
    if (K > 3) costs[3] = c.w;
}
#elif defined(ME_FORCE_THREAD) && defined(ME_BATCH_GROUPSUM) && !defined(ME_SADK_ARRAYS)
// Thread-only builds without the window test (the 2Nx2N frame search, the lookahead search): the same register-only calling form --
// the K candidate offsets arrive packed in registers and the K sums return in an int4 (as arrays of an out-of-line function they
// lived in local memory: every candidate began with two dependent LDL, 7 % of the stall samples of profiles/r02_me_frame_v11.txt).
template<typename pixel>
__device__ __noinline__ int4 thread_cand_sads(const MEState<pixel>& s, int K, uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3)
{
    int part[4] = { 0, 0, 0, 0 };
#pragma unroll 1
    for (int k = 0; k < K; k++)
    {
        const uint32_t pk = k == 0 ? p0 : (k == 1 ? p1 : (k == 2 ? p2 : p3));
        const int mx = (int)(int16_t)(pk & 0xffffu), my = (int)pk >> 16;
        const int v = thread_sad_any<pixel>(s, s.fref + mx + (int64_t)my * s.stride, s.stride);
#pragma unroll
        for (int j = 0; j < 4; j++) part[j] = j == k ? v : part[j];
    }
    for (int o = 1; o < s.groupSize; o <<= 1)
    {
#pragma unroll
        for (int k = 0; k < 4; k++) part[k] += __shfl_xor_sync(s.groupMask, part[k], o);
    }
    return make_int4(part[0], part[1], part[2], part[3]);
}
template<typename pixel>
__device__ __forceinline__ void warp_sad_k(const MEState<pixel>& s, int K, const int ox[4], const int oy[4], int costs[4])
{
    const int4 c = thread_cand_sads<pixel>(s, K, pack_cand(ox[0], oy[0]), K > 1 ? pack_cand(ox[1], oy[1]) : 0u, K > 2 ? pack_cand(ox[2], oy[2]) : 0u,
                                           K > 3 ? pack_cand(ox[3], oy[3]) : 0u);
    costs[0] = c.x;
    if (K > 1) costs[1] = c.y;
    if (K > 2) costs[2] = c.z;
    if (K > 3) costs[3] = c.w;
}
#else
template<typename pixel>
__device__ __noinline__ void warp_sad_k(const MEState<pixel>& s, int K, const int ox[4], const int oy[4], int costs[4])
{
    if (ME_IS_THREAD(s))
    {
#ifdef ME_BATCH_GROUPSUM
        // the K partial SADs first, then ONE butterfly over the lanes of the PU for all of them -- K independent shuffles per
        // round instead of K dependent 5-round chains (the reduction was 2.6 % of the instructions but 8.8 % of the stall
        // samples in profiles/r01_me_frame_v8_lines.txt; -3.8 % kernel time, profiles/r01_me_frame_v10_ab.txt).  Same sums.
        int part[4] = { 0, 0, 0, 0 };
#pragma unroll 1
        for (int k = 0; k < K; k++)
        {
            const int v = thread_sad_any<pixel>(s, s.fref + ox[k] + (int64_t)oy[k] * s.stride, s.stride);
#pragma unroll
            for (int j = 0; j < 4; j++) part[j] = j == k ? v : part[j];      // stays in registers (no dynamic indexing)
        }
        for (int o = 1; o < s.groupSize; o <<= 1)
        {
#pragma unroll
            for (int k = 0; k < 4; k++) part[k] += __shfl_xor_sync(s.groupMask, part[k], o);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) if (k < K) costs[k] = part[k];
#else
        for (int k = 0; k < K; k++)
            costs[k] = group_sum<pixel>(s, thread_sad_any<pixel>(s, s.fref + ox[k] + (int64_t)oy[k] * s.stride, s.stride));
#endif
        return;
    }
#ifndef ME_FORCE_THREAD
    if (!(ME_PU_W(s) & 15))     sad_k_impl<pixel, 16>(s, K, ox, oy, costs);
    else if (!(ME_PU_W(s) & 7)) sad_k_impl<pixel, 8>(s, K, ox, oy, costs);
    else                 sad_k_impl<pixel, 4>(s, K, ox, oy, costs);
#endif
}
#endif

// SAD of the cached PU against an arbitrary block (global or shared) with row stride rs
template<typename pixel>
__device__ __noinline__ int warp_sad_block(const MEState<pixel>& s, const pixel* r, int64_t rs)
{
    if (ME_IS_THREAD(s)) return group_sum<pixel>(s, thread_sad_any<pixel>(s, r, rs));
    const int gw = ME_PU_W(s) >> 2, ng = gw * s.h;
    int acc = 0;
    for (int u = s.lane; u < ng; u += 32)
    {
        int y = u / gw, x = (u - y * gw) << 2;
        acc += sad_seg<pixel, 4>(s.fenc + y * 64 + x, r + (int64_t)y * rs + x);
    }
    return warp_sum(acc);
}

// SATD (pixel.cpp:210-297): sum over 4x4 cells of (sum|H d H^T| >> 1); ref anywhere (generic pointer)
template<typename pixel>
__device__ __noinline__ int warp_satd(const MEState<pixel>& s, const pixel* r, int64_t rs)
{
    if (ME_IS_THREAD(s)) return group_sum<pixel>(s, thread_satd<pixel>(s, r, rs));
    constexpr int NW = 4 * (int)sizeof(pixel) / 4;
    const int cw = ME_PU_W(s) >> 2, nc = cw * (s.h >> 2);
    int acc = 0;
#pragma unroll 1
    for (int c = s.lane; c < nc; c += 32)
    {
        int cy = c / cw, cx = c - cy * cw;
        const pixel* f = s.fenc + cy * 4 * 64 + cx * 4;
        const pixel* q = r + (int64_t)cy * 4 * rs + cx * 4;
        int d[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
            int a[4], b[4];
            uint32_t rw[NW];
            unpack4<pixel>(smem_hint((const uint32_t*)(f + i * 64)), a);
            ld_words<pixel, NW>(q + i * rs, rw);
            unpack4<pixel>(rw, b);
#pragma unroll
            for (int k = 0; k < 4; k++) d[i][k] = a[k] - b[k];
            me_hadamard4(d[i][0], d[i][1], d[i][2], d[i][3]);
        }
        int t = 0;
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            me_hadamard4(d[0][k], d[1][k], d[2][k], d[3][k]);
            t += abs(d[0][k]) + abs(d[1][k]) + abs(d[2][k]) + abs(d[3][k]);
        }
        acc += t >> 1;
    }
    return warp_sum(acc);
}

// ---- subpel (motion.cpp:1571-1599, luma only) ---------------------------------------------------
template<typename pixel> __device__ __forceinline__ void unpack_row12(const uint32_t* w, int v[12]);
template<> __device__ __forceinline__ void unpack_row12<uint8_t>(const uint32_t* w, int v[12])
{
#pragma unroll
    for (int i = 0; i < 3; i++) { v[4 * i] = w[i] & 0xff; v[4 * i + 1] = (w[i] >> 8) & 0xff; v[4 * i + 2] = (w[i] >> 16) & 0xff; v[4 * i + 3] = w[i] >> 24; }
}
template<> __device__ __forceinline__ void unpack_row12<uint16_t>(const uint32_t* w, int v[12])
{
#pragma unroll
    for (int i = 0; i < 6; i++) { v[2 * i] = w[i] & 0xffff; v[2 * i + 1] = w[i] >> 16; }
}
template<typename pixel> __device__ __forceinline__ void store_px4(pixel* p, const int v[4]);
template<> __device__ __forceinline__ void store_px4<uint8_t>(uint8_t* p, const int v[4])
{
    *(uint32_t*)p = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24);
}
template<> __device__ __forceinline__ void store_px4<uint16_t>(uint16_t* p, const int v[4])
{
    uint32_t* q = (uint32_t*)p;
    q[0] = (uint32_t)v[0] | ((uint32_t)v[1] << 16); q[1] = (uint32_t)v[2] | ((uint32_t)v[3] << 16);
}

// 8-tap horizontal FIR of 4 adjacent 8-bit outputs from 3 packed words (12 pixels) with two u8 x s8 dot products
// per output (DP4A): out[k] = sum_t px[k+t]*c[t], taps packed 4 per word (|c| <= 58 fits s8).
__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c)
{
#ifdef ME_HOST_EMU
    return sp_dp4a_us(a, b, c);
#else
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#endif
}
__device__ __forceinline__ void hfir4_u8(const uint32_t w[3], uint32_t clo, uint32_t chi, int out[4])
{
    out[0] = dp4a_us(w[1], chi, dp4a_us(w[0], clo, 0));
#pragma unroll
    for (int k = 1; k < 4; k++)
        out[k] = dp4a_us(__funnelshift_r(w[1], w[2], 8 * k), chi, dp4a_us(__funnelshift_r(w[0], w[1], 8 * k), clo, 0));
}
__device__ __forceinline__ uint32_t pack_taps(const int c[8], int o)
{
    return (uint32_t)(c[o] & 0xff) | ((uint32_t)(c[o + 1] & 0xff) << 8) | ((uint32_t)(c[o + 2] & 0xff) << 16) | ((uint32_t)(c[o + 3] & 0xff) << 24);
}

// interpolates the PU at fractional (xFrac,yFrac) of the block starting at `src` into s.pred (stride w).
// Each lane produces groups of 4 horizontally adjacent pixels from word loads (3 words per row for the
// horizontal taps, one word per tap row for the vertical ones).
template<typename pixel>
__device__ __noinline__ void warp_interp_luma(const MEState<pixel>& s, const pixel* src, int xFrac, int yFrac)
{
    constexpr int NW12 = 12 * (int)sizeof(pixel) / 4, NW4 = 4 * (int)sizeof(pixel) / 4;
    const int w = ME_PU_W(s), h = s.h, maxVal = (1 << s.depth) - 1, headRoom = 14 - s.depth;
    const int gw = w >> 2;
    const int u0 = ME_IS_THREAD(s) ? 0 : s.lane, du = ME_IS_THREAD(s) ? 1 : 32;
    int c[8];
    if (!yFrac)
    {
        // luma_hpp : ipfilter.cpp:79-118
#pragma unroll
        for (int t = 0; t < 8; t++) c[t] = c_meLumaFilter[xFrac][t];
#pragma unroll 1
        for (int u = u0; u < gw * h; u += du)
        {
            int y = u / gw, x = (u - y * gw) << 2;
            uint32_t rw[NW12]; int v[12], o[4], sums[4];
            ld_words<pixel, NW12>(src + (int64_t)y * s.stride + x - 3, rw);
            if (sizeof(pixel) == 1) hfir4_u8(rw, pack_taps(c, 0), pack_taps(c, 4), sums);
            else
            {
                unpack_row12<pixel>(rw, v);
#pragma unroll
                for (int k = 0; k < 4; k++)
                {
                    sums[k] = 0;
#pragma unroll
                    for (int t = 0; t < 8; t++) sums[k] += v[k + t] * c[t];
                }
            }
#pragma unroll
            for (int k = 0; k < 4; k++)
            {
                int val = (int16_t)((sums[k] + 32) >> 6);
                o[k] = val < 0 ? 0 : (val > maxVal ? maxVal : val);
            }
            store_px4<pixel>(s.pred + y * w + x, o);
        }
    }
    else if (!xFrac)
    {
        // luma_vpp : ipfilter.cpp:164-203.  Unit = 4 columns x 4 rows: the 11 source rows are loaded and unpacked once
        // and reused by the 4 output rows (sliding window) instead of 8 loads per output row.
#pragma unroll
        for (int t = 0; t < 8; t++) c[t] = c_meLumaFilter[yFrac][t];
#pragma unroll 1
        for (int u = u0; u < gw * (h >> 2); u += du)
        {
            const int rb = u / gw, x = (u - rb * gw) << 2, y0 = rb << 2;
            int win[11][4];
#pragma unroll
            for (int r = 0; r < 11; r++)
            {
                uint32_t rw[NW4];
                ld_words<pixel, NW4>(src + (int64_t)(y0 + r - 3) * s.stride + x, rw);
                unpack4<pixel>(rw, win[r]);
            }
#pragma unroll
            for (int r = 0; r < 4; r++)
            {
                int o[4];
#pragma unroll
                for (int k = 0; k < 4; k++)
                {
                    int sum = 0;
#pragma unroll
                    for (int t = 0; t < 8; t++) sum += win[r + t][k] * c[t];
                    int val = (int16_t)((sum + 32) >> 6);
                    o[k] = val < 0 ? 0 : (val > maxVal ? maxVal : val);
                }
                store_px4<pixel>(s.pred + (y0 + r) * w + x, o);
            }
        }
    }
    else
    {
        // luma_hvpp = hps(isRowExt) + vsp : ipfilter.cpp:362-369
#pragma unroll
        for (int t = 0; t < 8; t++) c[t] = c_meLumaFilter[xFrac][t];
        const int shift = 6 - headRoom, offset = (int)((unsigned)-8192 << shift);
#pragma unroll 1
        for (int u = u0; u < gw * (h + 7); u += du)
        {
            int y = u / gw, x = (u - y * gw) << 2;
            uint32_t rw[NW12]; int v[12], sums[4];
            ld_words<pixel, NW12>(src + (int64_t)(y - 3) * s.stride + x - 3, rw);
            if (sizeof(pixel) == 1) hfir4_u8(rw, pack_taps(c, 0), pack_taps(c, 4), sums);
            else
            {
                unpack_row12<pixel>(rw, v);
#pragma unroll
                for (int k = 0; k < 4; k++)
                {
                    sums[k] = 0;
#pragma unroll
                    for (int t = 0; t < 8; t++) sums[k] += v[k + t] * c[t];
                }
            }
            int o[4];
#pragma unroll
            for (int k = 0; k < 4; k++) o[k] = (int16_t)((sums[k] + offset) >> shift);
            uint32_t* d = (uint32_t*)(s.immed + y * w + x);
            d[0] = (uint32_t)(o[0] & 0xffff) | ((uint32_t)o[1] << 16);
            d[1] = (uint32_t)(o[2] & 0xffff) | ((uint32_t)o[3] << 16);
        }
        if (!ME_IS_THREAD(s)) __syncwarp();
#pragma unroll
        for (int t = 0; t < 8; t++) c[t] = c_meLumaFilter[yFrac][t];
        const int shift2 = 6 + headRoom, offset2 = (1 << (shift2 - 1)) + (8192 << 6);
#pragma unroll 1
        for (int u = u0; u < gw * (h >> 2); u += du)
        {
            const int rb = u / gw, x = (u - rb * gw) << 2, y0 = rb << 2;
            int win[11][4];
#pragma unroll
            for (int r = 0; r < 11; r++)
            {
                const uint32_t* q = (const uint32_t*)(s.immed + (y0 + r) * w + x);
                const uint32_t w0 = q[0], w1 = q[1];
                win[r][0] = (int)(int16_t)(w0 & 0xffff); win[r][1] = (int)w0 >> 16;
                win[r][2] = (int)(int16_t)(w1 & 0xffff); win[r][3] = (int)w1 >> 16;
            }
#pragma unroll
            for (int r = 0; r < 4; r++)
            {
                int o[4];
#pragma unroll
                for (int k = 0; k < 4; k++)
                {
                    int sum = 0;
#pragma unroll
                    for (int t = 0; t < 8; t++) sum += win[r + t][k] * c[t];
                    int val = (int16_t)((sum + offset2) >> shift2);
                    o[k] = val < 0 ? 0 : (val > maxVal ? maxVal : val);
                }
                store_px4<pixel>(s.pred + (y0 + r) * w + x, o);
            }
        }
    }
    if (!ME_IS_THREAD(s)) __syncwarp();
}

// ---- fused sub-pel cost, per-thread mode ------------------------------------------------------------------------
// The prediction of a sub-block is never stored: it is produced one 4x4 cell at a time in registers (the same
// luma_hpp / luma_vpp / luma_hvpp arithmetic as warp_interp_luma above, ipfilter.cpp:79-118,164-203,362-369) and
// the cell's SAD or SATD against fenc is accumulated on the spot.  Vertical taps slide down a 4-column strip: the
// 11-row window keeps its last 7 rows between cells, so every source (or first-stage) row is produced once.
template<typename pixel>
__device__ __forceinline__ void pack4(const int v[4], uint32_t* out);
template<> __device__ __forceinline__ void pack4<uint8_t>(const int v[4], uint32_t* out)
{
    out[0] = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24);
}
template<> __device__ __forceinline__ void pack4<uint16_t>(const int v[4], uint32_t* out)
{
    out[0] = (uint32_t)v[0] | ((uint32_t)v[1] << 16); out[1] = (uint32_t)v[2] | ((uint32_t)v[3] << 16);
}

// cost of one predicted 4x4 cell against fenc.  NOT inlined: the three producers of thread_subpel_cost (and the lowres
// average) share one copy -- the sub-pel code is instruction-fetch bound (profiles/r01_me_frame_v5.txt), so the prediction
// rows travel packed in registers (4 x NW words) and the Hadamard / SAD code exists once.
template<typename pixel> struct CellRows { uint32_t w[4 * (4 * (int)sizeof(pixel) / 4)]; };
template<typename pixel>
__device__ __noinline__ int cell_cost_packed(const pixel* f, CellRows<pixel> rows, bool useSatd)
{
    constexpr int NW = 4 * (int)sizeof(pixel) / 4;
    if (useSatd)
    {
#if defined(ME_WINDOW_FAST) || defined(ME_PACKED_SATD)
        if constexpr (sizeof(pixel) == 1)
        {
            uint32_t fw[4];
#pragma unroll
            for (int i = 0; i < 4; i++) fw[i] = *smem_hint((const uint32_t*)(f + i * 64));
            return satd4x4_packed_u8(fw, rows.w);
        }
        else
#endif
        {
        int d[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
            int a[4], o[4];
            unpack4<pixel>(smem_hint((const uint32_t*)(f + i * 64)), a);
            unpack4<pixel>(&rows.w[i * NW], o);
#pragma unroll
            for (int k = 0; k < 4; k++) d[i][k] = a[k] - o[k];
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
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        const uint32_t* fw = smem_hint((const uint32_t*)(f + i * 64));
#pragma unroll
        for (int j = 0; j < NW; j++) acc += sad_word<pixel>(fw[j], rows.w[i * NW + j]);
    }
    return (int)acc;
}
template<typename pixel>
__device__ __forceinline__ int cell_cost(const pixel* f, const int o[4][4], bool useSatd)
{
    constexpr int NW = 4 * (int)sizeof(pixel) / 4;
    CellRows<pixel> rows;
#pragma unroll
    for (int i = 0; i < 4; i++) pack4<pixel>(o[i], &rows.w[i * NW]);
    return cell_cost_packed<pixel>(f, rows, useSatd);
}

template<typename pixel>
__device__ __noinline__ int thread_subpel_cost(const MEState<pixel>& s, const pixel* src, int xFrac, int yFrac, bool useSatd)
{
    constexpr int NW12 = 12 * (int)sizeof(pixel) / 4, NW4 = 4 * (int)sizeof(pixel) / 4;
    const int maxVal = (1 << s.depth) - 1, headRoom = 14 - s.depth;
    const int W = ME_PU_W(s), H = s.h;
    int ch[8];
#pragma unroll
    for (int t = 0; t < 8; t++) ch[t] = c_meLumaFilter[xFrac][t];
    const uint32_t clo = pack_taps(ch, 0), chi = pack_taps(ch, 4);
    // horizontal 8-tap sums of 4 adjacent outputs whose first tap sits at p
#ifdef ME_SMEM_FAST
    // rows by 32-bit shared address: b0 resolves src, b3 the first tap (src - 3); columns advance by whole words
    const SRow b0 = srow(src), b3 = srow(src - 3);
    const uint32_t rsB = (uint32_t)s.stride * (uint32_t)sizeof(pixel);
    constexpr uint32_t PXB = (uint32_t)sizeof(pixel);
    auto hsums = [&](uint32_t a, int sums[4]) {
        uint32_t rw[NW12];
        lds_words<NW12>(a, b3.sh, rw);
#else
    auto hsums = [&](const pixel* p, int sums[4]) {
        uint32_t rw[NW12];
        ld_words<pixel, NW12>(p, rw);
#endif
        if (sizeof(pixel) == 1) hfir4_u8(rw, clo, chi, sums);
        else
        {
            int v[12];
            unpack_row12<pixel>(rw, v);
#pragma unroll
            for (int k = 0; k < 4; k++)
            {
                sums[k] = 0;
#pragma unroll
                for (int t = 0; t < 8; t++) sums[k] += v[k + t] * ch[t];
            }
        }
    };
    int acc = 0;
#if defined(ME_SMEM_FAST) && defined(ME_SUBPEL_PACKED)
    // the 8-bit one-pass cases on packed words (subpel_packed.cuh, checked on the host against the oracle): horizontal rows
    // with the folded rounding / one-instruction clip, vertical cells from their 11 source rows with ONE transpose per cell
    // instead of one per row and no sliding window (-6 % kernel time, profiles/r01_me_frame_v10_ab.txt)
    if constexpr (sizeof(pixel) == 1)
    {
        if (!yFrac)
        {
#pragma unroll 1
            for (int y0 = 0; y0 < H; y0 += 4)
#pragma unroll 1
                for (int x = 0; x < W; x += 4)
                {
                    CellRows<pixel> rows;
#pragma unroll
                    for (int j = 0; j < 4; j++) rows.w[j] = 0;
#pragma unroll 1
                    for (int r = 0; r < 4; r++)
                    {
                        uint32_t rw[3];
                        lds_words<3>(b3.a + (uint32_t)(y0 + r) * rsB + (uint32_t)x, b3.sh, rw);
                        rows.w[0] = rows.w[1]; rows.w[1] = rows.w[2]; rows.w[2] = rows.w[3];
                        rows.w[3] = hpp_row4_u8(rw, clo, chi);
                    }
                    acc += cell_cost_packed<pixel>(s.fenc + y0 * 64 + x, rows, useSatd);
                }
            return acc;
        }
        if (!xFrac)
        {
            int16_t cvt[8];
#pragma unroll
            for (int t = 0; t < 8; t++) cvt[t] = c_meLumaFilter[yFrac][t];
            const uint32_t cvlo = sp_taps(cvt, 0), cvhi = sp_taps(cvt, 4);
#ifdef ME_VCELL_REUSE
            // vertically adjacent cells share two of their three transposed 4-row blocks (+2.4 % on the frame search, profiles/r02_staged_ab.txt);
            // per cell only rows y0+5..y0+8 are loaded and transposed (the window keeps spare rows below the 8-tap footprint)
#pragma unroll 1
            for (int x = 0; x < W; x += 4)
            {
                uint32_t c0[4], c1[4], c2[4], r[4];
#pragma unroll
                for (int j = 0; j < 4; j++) lds_words<1>(b0.a + (uint32_t)(j - 3) * rsB + (uint32_t)x, b0.sh, &r[j]);
                sp_transpose4(r[0], r[1], r[2], r[3], c0);
#pragma unroll
                for (int j = 0; j < 4; j++) lds_words<1>(b0.a + (uint32_t)(j + 1) * rsB + (uint32_t)x, b0.sh, &r[j]);
                sp_transpose4(r[0], r[1], r[2], r[3], c1);
#pragma unroll 1
                for (int y0 = 0; y0 < H; y0 += 4)
                {
#pragma unroll
                    for (int j = 0; j < 4; j++) lds_words<1>(b0.a + (uint32_t)(y0 + 5 + j) * rsB + (uint32_t)x, b0.sh, &r[j]);
                    sp_transpose4(r[0], r[1], r[2], r[3], c2);
                    CellRows<pixel> rows;
                    vpp_cell_from_cols_u8(c0, c1, c2, cvlo, cvhi, rows.w);
                    acc += cell_cost_packed<pixel>(s.fenc + y0 * 64 + x, rows, useSatd);
#pragma unroll
                    for (int k = 0; k < 4; k++) { c0[k] = c1[k]; c1[k] = c2[k]; }
                }
            }
            return acc;
#else
#pragma unroll 1
            for (int x = 0; x < W; x += 4)
#pragma unroll 1
                for (int y0 = 0; y0 < H; y0 += 4)
                {
                    uint32_t r[11];
#pragma unroll
                    for (int j = 0; j < 11; j++) lds_words<1>(b0.a + (uint32_t)(y0 - 3 + j) * rsB + (uint32_t)x, b0.sh, &r[j]);
                    CellRows<pixel> rows;
                    vpp_cell_u8(r, cvlo, cvhi, rows.w);
                    acc += cell_cost_packed<pixel>(s.fenc + y0 * 64 + x, rows, useSatd);
                }
            return acc;
#endif
        }
    }
#endif
    if (!yFrac)
    {
        // luma_hpp.  One row of 4 pixels per loop iteration (~55 instructions) so that the loop plus the shared cell cost stay
        // inside the per-scheduler L0 instruction cache: the kernel is bound by instruction supply (ncu: stall_no_instruction is
        // ~60 % of the samples everywhere in this function; L1.5 -> L0 delivers ~2 instructions/clk/SM), not by issue slots.
#pragma unroll 1
        for (int y0 = 0; y0 < H; y0 += 4)
#pragma unroll 1
            for (int x = 0; x < W; x += 4)
            {
                CellRows<pixel> rows;
#pragma unroll
                for (int j = 0; j < 4 * NW4; j++) rows.w[j] = 0;
#pragma unroll 1
                for (int r = 0; r < 4; r++)
                {
                    int sums[4], o[4];
#ifdef ME_SMEM_FAST
                    hsums(b3.a + (uint32_t)(y0 + r) * rsB + (uint32_t)x * PXB, sums);
#else
                    hsums(src + (int64_t)(y0 + r) * s.stride + x - 3, sums);
#endif
#pragma unroll
                    for (int k = 0; k < 4; k++)
                    {
                        int val = (int16_t)((sums[k] + 32) >> 6);
                        o[k] = val < 0 ? 0 : (val > maxVal ? maxVal : val);
                    }
#pragma unroll
                    for (int j = 0; j < 3 * NW4; j++) rows.w[j] = rows.w[j + NW4];
                    pack4<pixel>(o, &rows.w[3 * NW4]);
                }
                acc += cell_cost_packed<pixel>(s.fenc + y0 * 64 + x, rows, useSatd);
            }
        return acc;
    }
    // luma_vpp (xFrac == 0) or luma_hvpp = hps(isRowExt) + vsp, again one (first-stage) row per loop iteration: the 8 rows
    // under the vertical taps live in a small register window that moves down by one row per iteration.  Output row i needs
    // source rows i-3 .. i+4, so iteration i brings in row i+4; the first 7 iterations only fill the window.
    int cv[8];
#pragma unroll
    for (int t = 0; t < 8; t++) cv[t] = c_meLumaFilter[yFrac][t];
    if constexpr (sizeof(pixel) == 1)
    {
        if (!xFrac)
        {
            // 8-bit luma_vpp: the window holds pixel words; per output row the two 4x4 byte blocks are transposed (8 PRMT each)
            // and every pixel is two u8 x s8 dot products (DP4A) down its column
            const uint32_t cvlo = pack_taps(cv, 0), cvhi = pack_taps(cv, 4);
#pragma unroll 1
            for (int x = 0; x < W; x += 4)
            {
                uint32_t wq[8];
#pragma unroll
                for (int t = 0; t < 8; t++) wq[t] = 0;
                CellRows<pixel> rows;
#pragma unroll
                for (int j = 0; j < 4; j++) rows.w[j] = 0;
#pragma unroll 1
                for (int i = -7; i < H; i++)
                {
#pragma unroll
                    for (int t = 0; t < 7; t++) wq[t] = wq[t + 1];
#ifdef ME_SMEM_FAST
                    lds_words<1>(b0.a + (uint32_t)(i + 4) * rsB + (uint32_t)x, b0.sh, &wq[7]);
#else
                    ld_words<pixel, 1>(src + (int64_t)(i + 4) * s.stride + x, &wq[7]);
#endif
                    if (i < 0) continue;
                    uint32_t ca[4], cb[4];
                    {
                        const uint32_t t0 = __byte_perm(wq[0], wq[1], 0x5140), t1 = __byte_perm(wq[2], wq[3], 0x5140);
                        const uint32_t t2 = __byte_perm(wq[0], wq[1], 0x7362), t3 = __byte_perm(wq[2], wq[3], 0x7362);
                        ca[0] = __byte_perm(t0, t1, 0x5410); ca[1] = __byte_perm(t0, t1, 0x7632);
                        ca[2] = __byte_perm(t2, t3, 0x5410); ca[3] = __byte_perm(t2, t3, 0x7632);
                        const uint32_t u0 = __byte_perm(wq[4], wq[5], 0x5140), u1 = __byte_perm(wq[6], wq[7], 0x5140);
                        const uint32_t u2 = __byte_perm(wq[4], wq[5], 0x7362), u3 = __byte_perm(wq[6], wq[7], 0x7362);
                        cb[0] = __byte_perm(u0, u1, 0x5410); cb[1] = __byte_perm(u0, u1, 0x7632);
                        cb[2] = __byte_perm(u2, u3, 0x5410); cb[3] = __byte_perm(u2, u3, 0x7632);
                    }
                    int o[4];
#pragma unroll
                    for (int k = 0; k < 4; k++)
                    {
                        const int val = (int16_t)((dp4a_us(cb[k], cvhi, dp4a_us(ca[k], cvlo, 0)) + 32) >> 6);
                        o[k] = val < 0 ? 0 : (val > maxVal ? maxVal : val);
                    }
                    rows.w[0] = rows.w[1]; rows.w[1] = rows.w[2]; rows.w[2] = rows.w[3];
                    pack4<pixel>(o, &rows.w[3]);
                    if ((i & 3) == 3) acc += cell_cost_packed<pixel>(s.fenc + (i - 3) * 64 + x, rows, useSatd);
                }
            }
            return acc;
        }
    }
    // first-stage rows as int16 pairs (two words per row of 4); the vertical taps run as DP2A on row pairs
    const int sh1 = 6 - headRoom, off1 = (int)((unsigned)-8192 << sh1);
    const int sh2 = xFrac ? 6 + headRoom : 6, off2 = xFrac ? (1 << (sh2 - 1)) + (8192 << 6) : 32;
    uint32_t tp[4];
#pragma unroll
    for (int q = 0; q < 4; q++) tp[q] = (uint32_t)(cv[2 * q] & 0xff) | ((uint32_t)(cv[2 * q + 1] & 0xff) << 8);
#pragma unroll 1
    for (int x = 0; x < W; x += 4)
    {
        uint32_t wp[8][2];
#pragma unroll
        for (int t = 0; t < 8; t++) { wp[t][0] = 0; wp[t][1] = 0; }
        CellRows<pixel> rows;
#pragma unroll
        for (int j = 0; j < 4 * NW4; j++) rows.w[j] = 0;
#pragma unroll 1
        for (int i = -7; i < H; i++)
        {
#pragma unroll
            for (int t = 0; t < 7; t++) { wp[t][0] = wp[t + 1][0]; wp[t][1] = wp[t + 1][1]; }
            {
                int v[4];
#ifdef ME_SMEM_FAST
                const uint32_t rowOff = (uint32_t)(i + 4) * rsB + (uint32_t)x * PXB;
#else
                const pixel* p = src + (int64_t)(i + 4) * s.stride + x;
#endif
                if (xFrac)
                {
                    int sums[4];
#ifdef ME_SMEM_FAST
                    hsums(b3.a + rowOff, sums);
#else
                    hsums(p - 3, sums);
#endif
#pragma unroll
                    for (int k = 0; k < 4; k++) v[k] = (sums[k] + off1) >> sh1;
                }
                else
                {
                    uint32_t rw[NW4];
#ifdef ME_SMEM_FAST
                    lds_words<NW4>(b0.a + rowOff, b0.sh, rw);
#else
                    ld_words<pixel, NW4>(p, rw);
#endif
                    unpack4<pixel>(rw, v);
                }
                wp[7][0] = (uint32_t)(v[0] & 0xffff) | ((uint32_t)v[1] << 16);
                wp[7][1] = (uint32_t)(v[2] & 0xffff) | ((uint32_t)v[3] << 16);
            }
            if (i < 0) continue;
            int o[4];
#pragma unroll
            for (int hf = 0; hf < 2; hf++)
            {
                int s0 = 0, s1 = 0;
#pragma unroll
                for (int q = 0; q < 4; q++)
                {
                    const uint32_t ra = wp[2 * q][hf], rb = wp[2 * q + 1][hf];
                    s0 = __dp2a_lo((int)__byte_perm(ra, rb, 0x5410), (int)tp[q], s0);      // pixel 2*hf    : rows 2q, 2q+1
                    s1 = __dp2a_lo((int)__byte_perm(ra, rb, 0x7632), (int)tp[q], s1);      // pixel 2*hf + 1
                }
                const int v0 = (int16_t)((s0 + off2) >> sh2), v1 = (int16_t)((s1 + off2) >> sh2);
                o[2 * hf] = v0 < 0 ? 0 : (v0 > maxVal ? maxVal : v0);
                o[2 * hf + 1] = v1 < 0 ? 0 : (v1 > maxVal ? maxVal : v1);
            }
#pragma unroll
            for (int j = 0; j < 3 * NW4; j++) rows.w[j] = rows.w[j + NW4];
            pack4<pixel>(o, &rows.w[3 * NW4]);
            if ((i & 3) == 3) acc += cell_cost_packed<pixel>(s.fenc + (i - 3) * 64 + x, rows, useSatd);
        }
    }
    return acc;
}

#ifndef ME_FORCE_THREAD
// ---- chroma term of subpelCompare (motion.cpp:1601-1661), warp-cooperative ---------------------------------------
__constant__ int16_t c_meChromaFilter[8][4] = {
    { 0, 64, 0, 0 }, { -2, 58, 10, -2 }, { -4, 54, 16, -2 }, { -6, 46, 28, -4 },
    { -4, 36, 36, -4 }, { -4, 28, 46, -6 }, { -2, 16, 54, -4 }, { -2, 10, 58, -2 } };   // == g_chromaFilter, constants.cpp:258-268

// SATD of a w x h block (multiples of 4) f (stride fs, 4-pixel aligned) vs r (any alignment)
template<typename pixel>
__device__ __noinline__ int warp_satd_blk(const pixel* f, int fs, const pixel* r, int64_t rs, int w, int h, int lane)
{
    constexpr int NW = 4 * (int)sizeof(pixel) / 4;
    const int cw = w >> 2, nc = cw * (h >> 2);
    int acc = 0;
#pragma unroll 1
    for (int c = lane; c < nc; c += 32)
    {
        int cy = c / cw, cx = c - cy * cw;
        const pixel* ff = f + cy * 4 * fs + cx * 4;
        const pixel* q = r + (int64_t)cy * 4 * rs + cx * 4;
        int d[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
            int a[4], b[4];
            uint32_t rw[NW];
            unpack4<pixel>((const uint32_t*)(ff + i * fs), a);
            ld_words<pixel, NW>(q + i * rs, rw);
            unpack4<pixel>(rw, b);
#pragma unroll
            for (int k = 0; k < 4; k++) d[i][k] = a[k] - b[k];
            me_hadamard4(d[i][0], d[i][1], d[i][2], d[i][3]);
        }
        int t = 0;
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            me_hadamard4(d[0][k], d[1][k], d[2][k], d[3][k]);
            t += abs(d[0][k]) + abs(d[1][k]) + abs(d[2][k]) + abs(d[3][k]);
        }
        acc += t >> 1;
    }
    return warp_sum(acc);
}

// chroma[csp].pu[].filter_hpp / filter_vpp / filter_hps(isRowExt) + filter_vsp (ipfilter.cpp:79-118,120-162,164-203,
// 241-282 with N = 4) of a wC x hC block at eighth-pel fractions into s.pred (stride wC)
template<typename pixel>
__device__ __noinline__ void warp_interp_chroma(const MEState<pixel>& s, const pixel* src, int xFrac, int yFrac, int wC, int hC)
{
    const int maxVal = (1 << s.depth) - 1, headRoom = 14 - s.depth;
    const int64_t st = s.strideC;
    if (!yFrac || !xFrac)
    {
        const int frac = yFrac ? yFrac : xFrac;
        const int64_t step = yFrac ? st : 1;
        const int c0 = c_meChromaFilter[frac][0], c1 = c_meChromaFilter[frac][1], c2 = c_meChromaFilter[frac][2], c3 = c_meChromaFilter[frac][3];
        for (int e = s.lane; e < wC * hC; e += 32)
        {
            const int y = e / wC, x = e - y * wC;
            const pixel* q = src + (int64_t)y * st + x - step;
            const int sum = (int)q[0] * c0 + (int)q[step] * c1 + (int)q[2 * step] * c2 + (int)q[3 * step] * c3;
            const int val = (int16_t)((sum + 32) >> 6);
            s.pred[e] = (pixel)(val < 0 ? 0 : (val > maxVal ? maxVal : val));
        }
    }
    else
    {
        {
            const int c0 = c_meChromaFilter[xFrac][0], c1 = c_meChromaFilter[xFrac][1], c2 = c_meChromaFilter[xFrac][2], c3 = c_meChromaFilter[xFrac][3];
            const int shift = 6 - headRoom, offset = (int)((unsigned)-8192 << shift);
            for (int e = s.lane; e < wC * (hC + 3); e += 32)       // rows -1 .. hC+1
            {
                const int y = e / wC, x = e - y * wC;
                const pixel* q = src + (int64_t)(y - 1) * st + x - 1;
                const int sum = (int)q[0] * c0 + (int)q[1] * c1 + (int)q[2] * c2 + (int)q[3] * c3;
                s.immed[e] = (int16_t)((sum + offset) >> shift);
            }
        }
        __syncwarp();
        const int c0 = c_meChromaFilter[yFrac][0], c1 = c_meChromaFilter[yFrac][1], c2 = c_meChromaFilter[yFrac][2], c3 = c_meChromaFilter[yFrac][3];
        const int shift = 6 + headRoom, offset = (1 << (shift - 1)) + (8192 << 6);
        for (int e = s.lane; e < wC * hC; e += 32)
        {
            const int16_t* q = s.immed + e;
            const int sum = (int)q[0] * c0 + (int)q[wC] * c1 + (int)q[2 * wC] * c2 + (int)q[3 * wC] * c3;
            const int val = (int16_t)((sum + offset) >> shift);
            s.pred[e] = (pixel)(val < 0 ? 0 : (val > maxVal ? maxVal : val));
        }
    }
    __syncwarp();
}

template<typename pixel>
__device__ __noinline__ int warp_chroma_cost(const MEState<pixel>& s, int qx, int qy)
{
    const int mvx = qx << (1 - s.hshift), mvy = qy << (1 - s.vshift);
    const int64_t off = (mvx >> 3) + (int64_t)(mvy >> 3) * s.strideC;
    const int xFrac = mvx & 7, yFrac = mvy & 7;
    const int wC = s.w >> s.hshift, hC = s.h >> s.vshift;
    int cost = 0;
    for (int c = 0; c < 2; c++)
    {
        const pixel* r = s.frefC[c] + off;
        if (!(xFrac | yFrac))
            cost += warp_satd_blk<pixel>(s.fencC[c], s.csize, r, s.strideC, wC, hC, s.lane);
        else
        {
            __syncwarp();
            warp_interp_chroma<pixel>(s, r, xFrac, yFrac, wC, hC);
            cost += warp_satd_blk<pixel>(s.fencC[c], s.csize, s.pred, wC, wC, hC, s.lane);
            __syncwarp();
        }
    }
    return cost;
}
#endif

#ifdef ME_WINDOW_CHECK
// ---- blocks outside the staged window: the same arithmetic from the global plane (generic pointers, compact loops) -------
// N consecutive pixels from any address in any space as ints
template<typename pixel, int N>
__device__ __forceinline__ void ld_px_run(const pixel* p, int v[N])
{
    constexpr int PW = 4 / (int)sizeof(pixel), NW = (N + PW - 1) / PW;
    uint32_t w[NW];
    ld_words<pixel, NW, true>(p, w);
#pragma unroll
    for (int i = 0; i < N; i++)
        v[i] = sizeof(pixel) == 1 ? (int)((w[i / PW] >> (8 * (i % PW))) & 0xff) : (int)((w[i / PW] >> (16 * (i % PW))) & 0xffff);
}
template<typename pixel>
__device__ __noinline__ int thread_satd_gen(const MEState<pixel>& s, const pixel* r, int64_t rs)
{
    int acc = 0;
#pragma unroll 1
    for (int cy = 0; cy < s.h; cy += 4)
#pragma unroll 1
        for (int cx = 0; cx < ME_PU_W(s); cx += 4)
        {
            int o[4][4];
#pragma unroll
            for (int i = 0; i < 4; i++) ld_px_run<pixel, 4>(r + (int64_t)(cy + i) * rs + cx, o[i]);
            acc += cell_cost<pixel>(s.fenc + cy * 64 + cx, o, true);
        }
    return acc;
}
// luma_hpp / luma_vpp / luma_hvpp (ipfilter.cpp:79-118,164-203,362-369) of the lane's sub-block, 4x4 cell by cell
template<typename pixel>
__device__ __noinline__ int thread_subpel_cost_gen(const MEState<pixel>& s, const pixel* src, int64_t rs, int xFrac, int yFrac, bool useSatd)
{
    const int maxVal = (1 << s.depth) - 1, headRoom = 14 - s.depth;
    int ch[8], cv[8];
#pragma unroll
    for (int t = 0; t < 8; t++) { ch[t] = c_meLumaFilter[xFrac][t]; cv[t] = c_meLumaFilter[yFrac][t]; }
    const int sh1 = 6 - headRoom, off1 = (int)((unsigned)-8192 << sh1);
    const int sh2 = xFrac ? 6 + headRoom : 6, off2 = xFrac ? (1 << (sh2 - 1)) + (8192 << 6) : 32;
    int acc = 0;
#pragma unroll 1
    for (int cy = 0; cy < s.h; cy += 4)
#pragma unroll 1
        for (int cx = 0; cx < ME_PU_W(s); cx += 4)
        {
            int o[4][4];
            if (!yFrac)
            {
#pragma unroll
                for (int i = 0; i < 4; i++)
                {
                    int v[11];
                    ld_px_run<pixel, 11>(src + (int64_t)(cy + i) * rs + cx - 3, v);
#pragma unroll
                    for (int k = 0; k < 4; k++)
                    {
                        int sum = 0;
#pragma unroll
                        for (int t = 0; t < 8; t++) sum += v[k + t] * ch[t];
                        const int val = (int16_t)((sum + 32) >> 6);
                        o[i][k] = val < 0 ? 0 : (val > maxVal ? maxVal : val);
                    }
                }
            }
            else
            {
                int win[11][4];
#pragma unroll
                for (int r = 0; r < 11; r++)
                {
                    int q[4];
                    if (xFrac)
                    {
                        int v[11];
                        ld_px_run<pixel, 11>(src + (int64_t)(cy + r - 3) * rs + cx - 3, v);
#pragma unroll
                        for (int k = 0; k < 4; k++)
                        {
                            int sum = 0;
#pragma unroll
                            for (int t = 0; t < 8; t++) sum += v[k + t] * ch[t];
                            q[k] = (int16_t)((sum + off1) >> sh1);
                        }
                    }
                    else
                        ld_px_run<pixel, 4>(src + (int64_t)(cy + r - 3) * rs + cx, q);
#pragma unroll
                    for (int k = 0; k < 4; k++) win[r][k] = q[k];
                }
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int k = 0; k < 4; k++)
                    {
                        int sum = 0;
#pragma unroll
                        for (int t = 0; t < 8; t++) sum += win[i + t][k] * cv[t];
                        const int val = (int16_t)((sum + off2) >> sh2);
                        o[i][k] = val < 0 ? 0 : (val > maxVal ? maxVal : val);
                    }
            }
            acc += cell_cost<pixel>(s.fenc + cy * 64 + cx, o, useSatd);
        }
    return acc;
}
#endif

#ifdef ME_THREAD_CHROMA
// ---- chroma term of subpelCompare (motion.cpp:1601-1661), per-thread form ---------------------------------------------------
// The lane's Cb / Cr sub-blocks (its luma sub-block >> the chroma shifts; the host splits PUs so that they are multiples of
// 4x4) are predicted 4x4 cell by cell with the 4-tap filters (chroma[csp].pu[].filter_hpp / filter_vpp / filter_hps(isRowExt)
// + filter_vsp, ipfilter.cpp:79-162,164-203,241-282 with N = 4) and costed with the SATD on the spot.  The source CTU's Cb /
// Cr live in shared memory at a row pitch of 64 like the luma, the reference samples come from the staged chroma windows
// or, outside them, from the global planes.
__constant__ int16_t c_meChromaFilterT[8][4] = {
    { 0, 64, 0, 0 }, { -2, 58, 10, -2 }, { -4, 54, 16, -2 }, { -6, 46, 28, -4 },
    { -4, 36, 36, -4 }, { -4, 28, 46, -6 }, { -2, 16, 54, -4 }, { -2, 10, 58, -2 } };   // == g_chromaFilter, constants.cpp:258-268
template<typename pixel>
__device__ __noinline__ int thread_chroma_cost(const MEState<pixel>& s, int qx, int qy)
{
    const int mvx = qx << (1 - s.hshift), mvy = qy << (1 - s.vshift);
    const int ix = mvx >> 3, iy = mvy >> 3, xFrac = mvx & 7, yFrac = mvy & 7;
    const int wC = ME_PU_W(s) >> s.hshift, hC = s.h >> s.vshift;
    const int maxVal = (1 << s.depth) - 1, headRoom = 14 - s.depth;
    const bool inw = (ix >= s.cwinX0) & (ix <= s.cwinX1) & (iy >= s.cwinY0) & (iy <= s.cwinY1);
    const int64_t rs = inw ? s.strideC : s.gstrideC;
    int ch[4], cv[4];
#pragma unroll
    for (int t = 0; t < 4; t++) { ch[t] = c_meChromaFilterT[xFrac][t]; cv[t] = c_meChromaFilterT[yFrac][t]; }
    const int sh1 = 6 - headRoom, off1 = (int)((unsigned)-8192 << sh1);
    const int sh2 = xFrac ? 6 + headRoom : 6, off2 = xFrac ? (1 << (sh2 - 1)) + (8192 << 6) : 32;
    int acc = 0;
#pragma unroll 1
    for (int c = 0; c < 2; c++)
    {
        const pixel* src = (inw ? s.frefC[c] : s.gfrefC[c]) + ix + (int64_t)iy * rs;
#pragma unroll 1
        for (int cy = 0; cy < hC; cy += 4)
#pragma unroll 1
            for (int cx = 0; cx < wC; cx += 4)
            {
                int o[4][4];
                if (!(xFrac | yFrac))
                {
#pragma unroll
                    for (int i = 0; i < 4; i++) ld_px_run<pixel, 4>(src + (int64_t)(cy + i) * rs + cx, o[i]);
                }
                else if (!yFrac)
                {
#pragma unroll
                    for (int i = 0; i < 4; i++)
                    {
                        int v[7];
                        ld_px_run<pixel, 7>(src + (int64_t)(cy + i) * rs + cx - 1, v);
#pragma unroll
                        for (int k = 0; k < 4; k++)
                        {
                            const int val = (int16_t)((v[k] * ch[0] + v[k + 1] * ch[1] + v[k + 2] * ch[2] + v[k + 3] * ch[3] + 32) >> 6);
                            o[i][k] = val < 0 ? 0 : (val > maxVal ? maxVal : val);
                        }
                    }
                }
                else
                {
                    int win[7][4];
#pragma unroll
                    for (int r = 0; r < 7; r++)
                    {
                        if (xFrac)
                        {
                            int v[7];
                            ld_px_run<pixel, 7>(src + (int64_t)(cy + r - 1) * rs + cx - 1, v);
#pragma unroll
                            for (int k = 0; k < 4; k++)
                                win[r][k] = (int16_t)((v[k] * ch[0] + v[k + 1] * ch[1] + v[k + 2] * ch[2] + v[k + 3] * ch[3] + off1) >> sh1);
                        }
                        else
                            ld_px_run<pixel, 4>(src + (int64_t)(cy + r - 1) * rs + cx, win[r]);
                    }
#pragma unroll
                    for (int i = 0; i < 4; i++)
#pragma unroll
                        for (int k = 0; k < 4; k++)
                        {
                            const int val = (int16_t)((win[i][k] * cv[0] + win[i + 1][k] * cv[1] + win[i + 2][k] * cv[2] + win[i + 3][k] * cv[3] + off2) >> sh2);
                            o[i][k] = val < 0 ? 0 : (val > maxVal ? maxVal : val);
                        }
                }
                acc += cell_cost<pixel>(s.fencC[c] + cy * 64 + cx, o, true);
            }
    }
    return acc;
}
#endif

// MotionEstimate::subpelCompare (motion.cpp:1571-1599); useSatd selects cmp
// In a thread-only build without the window test this is a trampoline (pick the full-pel cost or the sub-pel one, then add the lanes'
// partials): inlined into its callers it saves a call level whose only memory traffic was re-loading the state.
#if defined(ME_FORCE_THREAD) && !defined(ME_WINDOW_CHECK) && !defined(ME_SUBPEL_CALL)
#define ME_SUBPEL_INLINE __forceinline__
#else
#define ME_SUBPEL_INLINE __noinline__
#endif
template<typename pixel>
__device__ ME_SUBPEL_INLINE int subpel_compare(const MEState<pixel>& s, int qx, int qy, bool useSatd)
{
    const pixel* fref = s.fref + (qx >> 2) + (int64_t)(qy >> 2) * s.stride;
    const int xFrac = qx & 3, yFrac = qy & 3;
    if (ME_IS_THREAD(s))
    {
#ifdef ME_WINDOW_CHECK
        // lane partials first (luma from the window or, outside it, from the plane; then the chroma term), ONE reduction
        int c;
        const int mx = qx >> 2, my = qy >> 2;
        if (in_window(s, mx, my, 4, 5))
        {
            if (!(xFrac | yFrac)) c = useSatd ? thread_satd<pixel>(s, fref, s.stride) : thread_sad_any<pixel>(s, fref, s.stride);
            else c = thread_subpel_cost<pixel>(s, fref, xFrac, yFrac, useSatd);
        }
        else
        {
            const pixel* g = s.gfref + mx + (int64_t)my * s.gstride;
            if (!(xFrac | yFrac)) c = useSatd ? thread_satd_gen<pixel>(s, g, s.gstride) : thread_sad_anyspace<pixel>(s, g, s.gstride);
            else c = thread_subpel_cost_gen<pixel>(s, g, s.gstride, xFrac, yFrac, useSatd);
        }
#ifdef ME_THREAD_CHROMA
        if (s.chromaSatd) c += thread_chroma_cost<pixel>(s, qx, qy);      // motion.cpp:1601
#endif
        return group_sum<pixel>(s, c);
#else
        if (!(xFrac | yFrac))
            return useSatd ? warp_satd<pixel>(s, fref, s.stride) : warp_sad_block<pixel>(s, fref, s.stride);
        const int gs = s.groupSize; const unsigned gm = s.groupMask;          // read before the call: not re-loaded on its critical path
        int v = thread_subpel_cost<pixel>(s, fref, xFrac, yFrac, useSatd);
        for (int o = 1; o < gs; o <<= 1) v += __shfl_xor_sync(gm, v, o);
        return v;
#endif
    }
#ifndef ME_FORCE_THREAD
    int c;
    if (!(xFrac | yFrac))
        c = useSatd ? warp_satd<pixel>(s, fref, s.stride) : warp_sad_block<pixel>(s, fref, s.stride);
    else
    {
        __syncwarp();
        warp_interp_luma<pixel>(s, fref, xFrac, yFrac);
        c = useSatd ? warp_satd<pixel>(s, s.pred, ME_PU_W(s)) : warp_sad_block<pixel>(s, s.pred, ME_PU_W(s));
        __syncwarp();
    }
    if (s.chromaSatd) c += warp_chroma_cost<pixel>(s, qx, qy);      // motion.cpp:1601
    return c;
#else
    return 0;
#endif
}

// ReferencePlanes::lowresQPelCost (common/lowres.h:94-120), 8x8 lowres blocks, hme = false
// per-thread form: the pixelavg_pp of the two hpel planes (pixel.cpp:545-557) is built 4x4 cell by cell in registers
// (the byte-wise rounded average is exactly (a + b + 1) >> 1) and costed on the spot
template<typename pixel> __device__ __forceinline__ uint32_t avg_word(uint32_t a, uint32_t b);
template<> __device__ __forceinline__ uint32_t avg_word<uint8_t>(uint32_t a, uint32_t b) { return __vavgu4(a, b); }
template<> __device__ __forceinline__ uint32_t avg_word<uint16_t>(uint32_t a, uint32_t b) { return __vavgu2(a, b); }

template<typename pixel>
__device__ __forceinline__ int thread_lowres_avg_cost(const MEState<pixel>& s, const pixel* A, const pixel* B, bool useSatd)
{
    // lowres CUs are 8 pixels wide: both 4x4 cells of a row group come from one batch of (A, B) row loads
    constexpr int NW8 = 8 * (int)sizeof(pixel) / 4, NW4 = NW8 / 2;
    int acc = 0;
#if defined(ME_PACKED_SATD) && !defined(ME_LA_SATD8_OFF)
    // 8-bit SATD (the quarter-pel candidates and the neighbour-MV candidates of the lookahead): the averaged rows ARE the packed words
    // the packed-word SATD takes -- no unpacking into cells, no out-of-line cell cost; four rows (two cells) per batch
    if constexpr (sizeof(pixel) == 1)
    {
        if (useSatd)
        {
#pragma unroll 1
            for (int y0 = 0; y0 < s.h; y0 += 4)
            {
                uint32_t wl[4], wr[4], fl[4], fr[4];
#pragma unroll
                for (int r = 0; r < 4; r++)
                {
                    uint32_t wa[2], wb[2];
                    ld_words<pixel, 2>(A + (int64_t)(y0 + r) * s.stride, wa);
                    ld_words<pixel, 2>(B + (int64_t)(y0 + r) * s.stride, wb);
                    wl[r] = __vavgu4(wa[0], wb[0]); wr[r] = __vavgu4(wa[1], wb[1]);
                    const uint2 f = *smem_hint((const uint2*)(s.fenc + (y0 + r) * 64));
                    fl[r] = f.x; fr[r] = f.y;
                }
                acc += satd4x4_packed_u8(fl, wl) + satd4x4_packed_u8(fr, wr);
            }
            return acc;
        }
    }
#endif
#pragma unroll 1
    for (int y0 = 0; y0 < s.h; y0 += 4)
    {
        uint32_t wv[4][NW8];
#pragma unroll
        for (int r = 0; r < 4; r++)
        {
            uint32_t wa[NW8], wb[NW8];
            ld_words<pixel, NW8>(A + (int64_t)(y0 + r) * s.stride, wa);
            ld_words<pixel, NW8>(B + (int64_t)(y0 + r) * s.stride, wb);
#pragma unroll
            for (int i = 0; i < NW8; i++) wv[r][i] = avg_word<pixel>(wa[i], wb[i]);
        }
#pragma unroll
        for (int c = 0; c < 2; c++)
        {
            int o[4][4];
#pragma unroll
            for (int r = 0; r < 4; r++) unpack4<pixel>(wv[r] + c * NW4, o[r]);
            acc += cell_cost<pixel>(s.fenc + y0 * 64 + c * 4, o, useSatd);
        }
    }
    return acc;
}

template<typename pixel>
__device__ __noinline__ int lowres_qpel_cost(const MEState<pixel>& s, int qx, int qy, bool useSatd)
{
    if (ME_IS_THREAD(s))
    {
        if ((qx | qy) & 1)
        {
            const int hpelA = (qy & 2) | ((qx & 2) >> 1);
            const pixel* frefA = s.lowres[hpelA] + (qx >> 2) + (int64_t)(qy >> 2) * s.stride;
            const int qmvx = qx + (qx & 1), qmvy = qy + (qy & 1);
            const int hpelB = (qmvy & 2) | ((qmvx & 2) >> 1);
            const pixel* frefB = s.lowres[hpelB] + (qmvx >> 2) + (int64_t)(qmvy >> 2) * s.stride;
            return group_sum<pixel>(s, thread_lowres_avg_cost<pixel>(s, frefA, frefB, useSatd));
        }
        const int hpel = (qy & 2) | ((qx & 2) >> 1);
        const pixel* fref = s.lowres[hpel] + (qx >> 2) + (int64_t)(qy >> 2) * s.stride;
        return useSatd ? warp_satd<pixel>(s, fref, s.stride) : warp_sad_block<pixel>(s, fref, s.stride);
    }
#ifndef ME_FORCE_THREAD
    if ((qx | qy) & 1)
    {
        int hpelA = (qy & 2) | ((qx & 2) >> 1);
        const pixel* frefA = s.lowres[hpelA] + (qx >> 2) + (int64_t)(qy >> 2) * s.stride;
        int qmvx = qx + (qx & 1), qmvy = qy + (qy & 1);
        int hpelB = (qmvy & 2) | ((qmvx & 2) >> 1);
        const pixel* frefB = s.lowres[hpelB] + (qmvx >> 2) + (int64_t)(qmvy >> 2) * s.stride;
        __syncwarp();
        for (int e = s.lane; e < ME_PU_W(s) * s.h; e += 32)          // pixelavg_pp, pixel.cpp:545-557
        {
            int y = e / ME_PU_W(s), x = e - y * ME_PU_W(s);
            s.pred[e] = (pixel)(((int)frefA[(int64_t)y * s.stride + x] + (int)frefB[(int64_t)y * s.stride + x] + 1) >> 1);
        }
        __syncwarp();
        int c = useSatd ? warp_satd<pixel>(s, s.pred, ME_PU_W(s)) : warp_sad_block<pixel>(s, s.pred, ME_PU_W(s));
        __syncwarp();
        return c;
    }
    int hpel = (qy & 2) | ((qx & 2) >> 1);
    const pixel* fref = s.lowres[hpel] + (qx >> 2) + (int64_t)(qy >> 2) * s.stride;
    return useSatd ? warp_satd<pixel>(s, fref, s.stride) : warp_sad_block<pixel>(s, fref, s.stride);
#else
    return 0;
#endif
}

// ---- the search ------------------------------------------------------------------------------------
template<typename pixel>
struct MESearch
{
    const MEState<pixel>& s;
    MV2 mvmin, mvmax;
    MV2 bmv; int bcost;

    __device__ __forceinline__ MESearch(const MEState<pixel>& st) : s(st) {}

    __device__ __forceinline__ int sadAt(int mx, int my) const
    {
        int ox[4] = { mx, 0, 0, 0 }, oy[4] = { my, 0, 0, 0 }, c[4];
        warp_sad_k<pixel>(s, 1, ox, oy, c);
        return c[0];
    }
    __device__ __forceinline__ int fcost(int mx, int my) const { return mvcost(s, mx << 2, my << 2); }
    // SAD + mvcost of one / K full-pel candidates (COST_MV's and COST_MV_X4's operands, motion.cpp:238-298)
    __device__ __forceinline__ int costAt(int mx, int my) const
    {
#ifdef ME_WINDOW_CHECK
        return thread_cand_costs<pixel>(s, 1, pack_cand(mx, my), 0u, 0u, 0u, true).x;
#else
        return sadAt(mx, my) + fcost(mx, my);
#endif
    }
    __device__ __forceinline__ void candCosts(int K, const int ox[4], const int oy[4], int costs[4]) const
    {
#ifdef ME_WINDOW_CHECK
        warp_sad_k<pixel>(s, K, ox, oy, costs, true);
#else
        warp_sad_k<pixel>(s, K, ox, oy, costs);
        for (int k = 0; k < K; k++) costs[k] += fcost(ox[k], oy[k]);
#endif
    }
    __device__ __forceinline__ bool inRange(int x, int y) const { return x >= mvmin.x && x <= mvmax.x && y >= mvmin.y && y <= mvmax.y; }
    __device__ __forceinline__ bool yOk(int y) const { return (y >= mvmin.y) & (y <= mvmax.y); }

    // COST_MV (motion.cpp:238-244)
    __device__ __noinline__ void costMv(int mx, int my)
    {
        int cost = costAt(mx, my);
        if (cost < bcost) { bcost = cost; bmv = mv2(mx, my); }
    }
    // COST_MV_X4 (motion.cpp:277-298): only the y range is checked (quirk)
    __device__ __noinline__ void costMvX4(MV2 omv, int x0, int y0, int x1, int y1, int x2, int y2, int x3, int y3)
    {
        const int dx[4] = { x0, x1, x2, x3 }, dy[4] = { y0, y1, y2, y3 };
        int costs[4], ox[4], oy[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { ox[k] = omv.x + dx[k]; oy[k] = omv.y + dy[k]; }
        candCosts(4, ox, oy, costs);
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (yOk(omv.y + dy[k]) && costs[k] < bcost) { bcost = costs[k]; bmv = mv2(omv.x + dx[k], omv.y + dy[k]); }
    }
    // COST_MV_X4_DIR / COST_MV_X3_DIR (motion.cpp:246-257, :315-328): costs relative to bmv, no update
    __device__ __forceinline__ void dirCosts(int n, const int dx[], const int dy[], int costs[]) const
    {
        int ox[4] = { 0, 0, 0, 0 }, oy[4] = { 0, 0, 0, 0 };
        for (int k = 0; k < n; k++) { ox[k] = bmv.x + dx[k]; oy[k] = bmv.y + dy[k]; }
        candCosts(n, ox, oy, costs);
    }
    // CROSS (motion.cpp:336-360)
    __device__ __noinline__ void cross(MV2 omv, int start, int x_max, int y_max)
    {
        int i = start;
        if (x_max <= min(mvmax.x - omv.x, omv.x - mvmin.x))
            for (; i < x_max - 2; i += 4) costMvX4(omv, i, 0, -i, 0, i + 2, 0, -i - 2, 0);
        for (; i < x_max; i += 2)
        {
            if (omv.x + i <= mvmax.x) costMv(omv.x + i, omv.y);
            if (omv.x - i >= mvmin.x) costMv(omv.x - i, omv.y);
        }
        i = start;
        if (y_max <= min(mvmax.y - omv.y, omv.y - mvmin.y))
            for (; i < y_max - 2; i += 4) costMvX4(omv, 0, i, 0, -i, 0, i + 2, 0, -i - 2);
        for (; i < y_max; i += 2)
        {
            if (omv.y + i <= mvmax.y) costMv(omv.x, omv.y + i);
            if (omv.y - i >= mvmin.y) costMv(omv.x, omv.y - i);
        }
    }

    // COST_MV_PT_DIST (motion.cpp:224-236)
    __device__ __noinline__ void ptDist(int mx, int my, int point, int dist, int& bPointNr, int& bDistance)
    {
        int cost = costAt(mx, my);
        if (cost < bcost) { bcost = cost; bmv = mv2(mx, my); bPointNr = point; bDistance = dist; }
    }

    // StarPatternSearch (motion.cpp:362-604)
    __device__ __noinline__ void starPattern(int& bPointNr, int& bDistance, int earlyExitIters, int merange)
    {
        const MV2 omv = bmv;
        int saved = bcost, rounds = 0;
        {
            const int dist = 1;
            const int top = omv.y - dist, bottom = omv.y + dist, left = omv.x - dist, right = omv.x + dist;
            // the x4 form and the guarded form evaluate the same points in the same order
            if (top >= mvmin.y) ptDist(omv.x, top, 2, dist, bPointNr, bDistance);
            if (left >= mvmin.x) ptDist(left, omv.y, 4, dist, bPointNr, bDistance);
            if (right <= mvmax.x) ptDist(right, omv.y, 5, dist, bPointNr, bDistance);
            if (bottom <= mvmax.y) ptDist(omv.x, bottom, 7, dist, bPointNr, bDistance);
            if (bcost < saved) rounds = 0;
            else if (++rounds >= earlyExitIters) return;
        }
        for (int dist = 2; dist <= 8; dist <<= 1)
        {
            const int top = omv.y - dist, bottom = omv.y + dist, left = omv.x - dist, right = omv.x + dist;
            const int top2 = omv.y - (dist >> 1), bottom2 = omv.y + (dist >> 1), left2 = omv.x - (dist >> 1), right2 = omv.x + (dist >> 1);
            saved = bcost;
            if (top >= mvmin.y && left >= mvmin.x && right <= mvmax.x && bottom <= mvmax.y)
            {
                // x4 order (motion.cpp:451-458): 2,1,3,4 then 5,6,8,7
                ptDist(omv.x, top, 2, dist, bPointNr, bDistance);
                ptDist(left2, top2, 1, dist >> 1, bPointNr, bDistance);
                ptDist(right2, top2, 3, dist >> 1, bPointNr, bDistance);
                ptDist(left, omv.y, 4, dist, bPointNr, bDistance);
                ptDist(right, omv.y, 5, dist, bPointNr, bDistance);
                ptDist(left2, bottom2, 6, dist >> 1, bPointNr, bDistance);
                ptDist(right2, bottom2, 8, dist >> 1, bPointNr, bDistance);
                ptDist(omv.x, bottom, 7, dist, bPointNr, bDistance);
            }
            else
            {
                if (top >= mvmin.y) ptDist(omv.x, top, 2, dist, bPointNr, bDistance);
                if (top2 >= mvmin.y)
                {
                    if (left2 >= mvmin.x) ptDist(left2, top2, 1, dist >> 1, bPointNr, bDistance);
                    if (right2 <= mvmax.x) ptDist(right2, top2, 3, dist >> 1, bPointNr, bDistance);
                }
                if (left >= mvmin.x) ptDist(left, omv.y, 4, dist, bPointNr, bDistance);
                if (right <= mvmax.x) ptDist(right, omv.y, 5, dist, bPointNr, bDistance);
                if (bottom2 <= mvmax.y)
                {
                    if (left2 >= mvmin.x) ptDist(left2, bottom2, 6, dist >> 1, bPointNr, bDistance);
                    if (right2 <= mvmax.x) ptDist(right2, bottom2, 8, dist >> 1, bPointNr, bDistance);
                }
                if (bottom <= mvmax.y) ptDist(omv.x, bottom, 7, dist, bPointNr, bDistance);
            }
            if (bcost < saved) rounds = 0;
            else if (++rounds >= earlyExitIters) return;
        }
        for (int dist = 16; dist <= (int)(int16_t)merange; dist <<= 1)
        {
            const int top = omv.y - dist, bottom = omv.y + dist, left = omv.x - dist, right = omv.x + dist;
            saved = bcost;
            if (top >= mvmin.y && left >= mvmin.x && right <= mvmax.x && bottom <= mvmax.y)
            {
                ptDist(omv.x, top, 0, dist, bPointNr, bDistance);
                ptDist(left, omv.y, 0, dist, bPointNr, bDistance);
                ptDist(right, omv.y, 0, dist, bPointNr, bDistance);
                ptDist(omv.x, bottom, 0, dist, bPointNr, bDistance);
                for (int index = 1; index < 4; index++)
                {
                    int posYT = top + ((dist >> 2) * index), posYB = bottom - ((dist >> 2) * index);
                    int posXL = omv.x - ((dist >> 2) * index), posXR = omv.x + ((dist >> 2) * index);
                    ptDist(posXL, posYT, 0, dist, bPointNr, bDistance);
                    ptDist(posXR, posYT, 0, dist, bPointNr, bDistance);
                    ptDist(posXL, posYB, 0, dist, bPointNr, bDistance);
                    ptDist(posXR, posYB, 0, dist, bPointNr, bDistance);
                }
            }
            else
            {
                if (top >= mvmin.y) ptDist(omv.x, top, 0, dist, bPointNr, bDistance);
                if (left >= mvmin.x) ptDist(left, omv.y, 0, dist, bPointNr, bDistance);
                if (right <= mvmax.x) ptDist(right, omv.y, 0, dist, bPointNr, bDistance);
                if (bottom <= mvmax.y) ptDist(omv.x, bottom, 0, dist, bPointNr, bDistance);
                for (int index = 1; index < 4; index++)
                {
                    int posYT = top + ((dist >> 2) * index), posYB = bottom - ((dist >> 2) * index);
                    int posXL = omv.x - ((dist >> 2) * index), posXR = omv.x + ((dist >> 2) * index);
                    if (posYT >= mvmin.y)
                    {
                        if (posXL >= mvmin.x) ptDist(posXL, posYT, 0, dist, bPointNr, bDistance);
                        if (posXR <= mvmax.x) ptDist(posXR, posYT, 0, dist, bPointNr, bDistance);
                    }
                    if (posYB <= mvmax.y)
                    {
                        if (posXL >= mvmin.x) ptDist(posXL, posYB, 0, dist, bPointNr, bDistance);
                        if (posXR <= mvmax.x) ptDist(posXR, posYB, 0, dist, bPointNr, bDistance);
                    }
                }
            }
            if (bcost < saved) rounds = 0;
            else if (++rounds >= earlyExitIters) return;
        }
    }

    // square refine shared by HEX (motion.cpp:924-942)
    __device__ __noinline__ void squareRefine()
    {
        int costs[4];
        int dir = 0;
        { const int dx[4] = { 0, 0, -1, 1 }, dy[4] = { -1, 1, 0, 0 }; dirCosts(4, dx, dy, costs); }
        if (yOk(bmv.y - 1) && costs[0] < bcost) { bcost = costs[0]; dir = 1; }
        if (yOk(bmv.y + 1) && costs[1] < bcost) { bcost = costs[1]; dir = 2; }
        if (costs[2] < bcost) { bcost = costs[2]; dir = 3; }
        if (costs[3] < bcost) { bcost = costs[3]; dir = 4; }
        { const int dx[4] = { -1, -1, 1, 1 }, dy[4] = { -1, 1, -1, 1 }; dirCosts(4, dx, dy, costs); }
        if (yOk(bmv.y - 1) && costs[0] < bcost) { bcost = costs[0]; dir = 5; }
        if (yOk(bmv.y + 1) && costs[1] < bcost) { bcost = costs[1]; dir = 6; }
        if (yOk(bmv.y - 1) && costs[2] < bcost) { bcost = costs[2]; dir = 7; }
        if (yOk(bmv.y + 1) && costs[3] < bcost) { bcost = costs[3]; dir = 8; }
        bmv.x += c_square1[dir][0]; bmv.y += c_square1[dir][1];
    }

    // me_hex2 (motion.cpp:845-944)
#if defined(ME_WINDOW_CHECK) || defined(ME_HEX_MEMBERS) || defined(ME_LOWRES_ONLY)      /* the lookahead build measured 1 % slower with the register form */
    __device__ __noinline__ void hexSearch(int merange)
    {
        int costs[4];
        { const int dx[3] = { -2, -1, 1 }, dy[3] = { 0, 2, 2 }; dirCosts(3, dx, dy, costs); }
        bcost <<= 3;
        if (yOk(bmv.y) && (costs[0] << 3) + 2 < bcost) bcost = (costs[0] << 3) + 2;
        if (yOk(bmv.y + 2))
        {
            if ((costs[1] << 3) + 3 < bcost) bcost = (costs[1] << 3) + 3;
            if ((costs[2] << 3) + 4 < bcost) bcost = (costs[2] << 3) + 4;
        }
        { const int dx[3] = { 2, 1, -1 }, dy[3] = { 0, -2, -2 }; dirCosts(3, dx, dy, costs); }
        if (yOk(bmv.y) && (costs[0] << 3) + 5 < bcost) bcost = (costs[0] << 3) + 5;
        if (yOk(bmv.y - 2))
        {
            if ((costs[1] << 3) + 6 < bcost) bcost = (costs[1] << 3) + 6;
            if ((costs[2] << 3) + 7 < bcost) bcost = (costs[2] << 3) + 7;
        }
        if (bcost & 7)
        {
            int dir = (bcost & 7) - 2;
            if (yOk(bmv.y + c_hex2[dir + 1][1]))
            {
                bmv.x += c_hex2[dir + 1][0]; bmv.y += c_hex2[dir + 1][1];
                for (int i = (merange >> 1) - 1; i > 0 && inRange(bmv.x, bmv.y); i--)
                {
                    const int dx[3] = { c_hex2[dir][0], c_hex2[dir + 1][0], c_hex2[dir + 2][0] };
                    const int dy[3] = { c_hex2[dir][1], c_hex2[dir + 1][1], c_hex2[dir + 2][1] };
                    dirCosts(3, dx, dy, costs);
                    bcost &= ~7;
                    if (yOk(bmv.y + dy[0]) && (costs[0] << 3) + 1 < bcost) bcost = (costs[0] << 3) + 1;
                    if (yOk(bmv.y + dy[1]) && (costs[1] << 3) + 2 < bcost) bcost = (costs[1] << 3) + 2;
                    if (yOk(bmv.y + dy[2]) && (costs[2] << 3) + 3 < bcost) bcost = (costs[2] << 3) + 3;
                    if (!(bcost & 7)) break;
                    dir += (bcost & 7) - 2;
                    dir = c_mod6m1[dir + 1];
                    bmv.x += c_hex2[dir + 1][0]; bmv.y += c_hex2[dir + 1][1];
                }
            }
        }
        bcost >>= 3;
        squareRefine();
    }
#else
    // The same walk with the search state in registers: as members of this object (which lives in local memory, like the MEState it
    // points to) the best cost / MV, the range and the MV-cost operands were re-loaded after every out-of-line candidate call --
    // three dependent loads (this -> s -> cost table) on the critical path of each of the ~8 steps (hexSearch: 27 M instructions but
    // 9 % of the stall samples in profiles/r02_me_frame_v12.txt).
    __device__ __noinline__ void hexSearch(int merange)
    {
        const MEState<pixel>& st = s;
        const uint16_t* const costT = st.cost;
        const int px = st.mvpx, py = st.mvpy;
        const MV2 mn = mvmin, mx = mvmax;
        MV2 b = bmv; int bc = bcost;
        auto yok = [&](int y) { return (y >= mn.y) & (y <= mx.y); };
        auto inr = [&](int x, int y) { return x >= mn.x && x <= mx.x && y >= mn.y && y <= mx.y; };
        auto dir3 = [&](int dx0, int dy0, int dx1, int dy1, int dx2, int dy2, int costs[4]) {
            int ox[4] = { b.x + dx0, b.x + dx1, b.x + dx2, 0 }, oy[4] = { b.y + dy0, b.y + dy1, b.y + dy2, 0 };
            warp_sad_k<pixel>(st, 3, ox, oy, costs);
#pragma unroll
            for (int k = 0; k < 3; k++)
            {
                const int ix = clip3i(-kMvTableHalf, kMvTableHalf, (ox[k] << 2) - px), iy = clip3i(-kMvTableHalf, kMvTableHalf, (oy[k] << 2) - py);
                costs[k] += ((int)costT[ix] + (int)costT[iy]) & 0xffff;                    // mvcost(), bitcost.h:45
            }
        };
        int costs[4];
        dir3(-2, 0, -1, 2, 1, 2, costs);
        bc <<= 3;
        if (yok(b.y) && (costs[0] << 3) + 2 < bc) bc = (costs[0] << 3) + 2;
        if (yok(b.y + 2))
        {
            if ((costs[1] << 3) + 3 < bc) bc = (costs[1] << 3) + 3;
            if ((costs[2] << 3) + 4 < bc) bc = (costs[2] << 3) + 4;
        }
        dir3(2, 0, 1, -2, -1, -2, costs);
        if (yok(b.y) && (costs[0] << 3) + 5 < bc) bc = (costs[0] << 3) + 5;
        if (yok(b.y - 2))
        {
            if ((costs[1] << 3) + 6 < bc) bc = (costs[1] << 3) + 6;
            if ((costs[2] << 3) + 7 < bc) bc = (costs[2] << 3) + 7;
        }
        if (bc & 7)
        {
            int dir = (bc & 7) - 2;
            if (yok(b.y + c_hex2[dir + 1][1]))
            {
                b.x += c_hex2[dir + 1][0]; b.y += c_hex2[dir + 1][1];
                for (int i = (merange >> 1) - 1; i > 0 && inr(b.x, b.y); i--)
                {
                    const int dy0 = c_hex2[dir][1], dy1 = c_hex2[dir + 1][1], dy2 = c_hex2[dir + 2][1];
                    dir3(c_hex2[dir][0], dy0, c_hex2[dir + 1][0], dy1, c_hex2[dir + 2][0], dy2, costs);
                    bc &= ~7;
                    if (yok(b.y + dy0) && (costs[0] << 3) + 1 < bc) bc = (costs[0] << 3) + 1;
                    if (yok(b.y + dy1) && (costs[1] << 3) + 2 < bc) bc = (costs[1] << 3) + 2;
                    if (yok(b.y + dy2) && (costs[2] << 3) + 3 < bc) bc = (costs[2] << 3) + 3;
                    if (!(bc & 7)) break;
                    dir += (bc & 7) - 2;
                    dir = c_mod6m1[dir + 1];
                    b.x += c_hex2[dir + 1][0]; b.y += c_hex2[dir + 1][1];
                }
            }
        }
        bc >>= 3;
        // square refine (motion.cpp:933-943), same registers
        auto dir4 = [&](int dx0, int dy0, int dx1, int dy1, int dx2, int dy2, int dx3, int dy3, int costs[4]) {
            int ox[4] = { b.x + dx0, b.x + dx1, b.x + dx2, b.x + dx3 }, oy[4] = { b.y + dy0, b.y + dy1, b.y + dy2, b.y + dy3 };
            warp_sad_k<pixel>(st, 4, ox, oy, costs);
#pragma unroll
            for (int k = 0; k < 4; k++)
            {
                const int ix = clip3i(-kMvTableHalf, kMvTableHalf, (ox[k] << 2) - px), iy = clip3i(-kMvTableHalf, kMvTableHalf, (oy[k] << 2) - py);
                costs[k] += ((int)costT[ix] + (int)costT[iy]) & 0xffff;
            }
        };
        int sq = 0;
        dir4(0, -1, 0, 1, -1, 0, 1, 0, costs);
        if (yok(b.y - 1) && costs[0] < bc) { bc = costs[0]; sq = 1; }
        if (yok(b.y + 1) && costs[1] < bc) { bc = costs[1]; sq = 2; }
        if (costs[2] < bc) { bc = costs[2]; sq = 3; }
        if (costs[3] < bc) { bc = costs[3]; sq = 4; }
        dir4(-1, -1, -1, 1, 1, -1, 1, 1, costs);
        if (yok(b.y - 1) && costs[0] < bc) { bc = costs[0]; sq = 5; }
        if (yok(b.y + 1) && costs[1] < bc) { bc = costs[1]; sq = 6; }
        if (yok(b.y - 1) && costs[2] < bc) { bc = costs[2]; sq = 7; }
        if (yok(b.y + 1) && costs[3] < bc) { bc = costs[3]; sq = 8; }
        b.x += c_square1[sq][0]; b.y += c_square1[sq][1];
        bmv = b; bcost = bc;
    }
#endif

    // two-point refinement after a distance-1 star result (motion.cpp:1151-1166, :1225-1235)
    __device__ __forceinline__ void twoPoint(int bPointNr)
    {
        const MV2 b = bmv;
        const int x1 = b.x + c_offsets[(bPointNr - 1) * 2][0], y1 = b.y + c_offsets[(bPointNr - 1) * 2][1];
        const int x2 = b.x + c_offsets[(bPointNr - 1) * 2 + 1][0], y2 = b.y + c_offsets[(bPointNr - 1) * 2 + 1][1];
        if (inRange(x1, y1)) costMv(x1, y1);
        if (inRange(x2, y2)) costMv(x2, y2);
    }
};

#ifndef ME_FORCE_THREAD
// ---- --me sea: successive elimination (motion.cpp:1242-1395) -------------------------------------------------------
// One table lookup of the lambda-scaled cost, index clipped to the table like mvcost() above.
template<typename pixel>
__device__ __forceinline__ int cost_at(const MEState<pixel>& s, int idx)
{
    return (int)s.cost[clip3i(-kMvTableHalf, kMvTableHalf, idx)];
}

template<typename pixel>
__device__ __noinline__ void sea_search(MESearch<pixel>& S, const MEState<pixel>& s, MV2 omv, int merange)
{
    const int w = s.w, h = s.h, lane = s.lane;
    const int minX = max(omv.x - merange, S.mvmin.x), minY = max(omv.y - merange, S.mvmin.y);
    const int maxX = min(omv.x + merange, S.mvmax.x), maxY = min(omv.y + merange, S.mvmax.y);
    const int width = (maxX - minX + 3) & ~3;                                   // "SEA is fastest in multiples of 4" (:1259-1260)
    int deltaX = (w <= 8) ? w : (w >> 1);
    int deltaY = (h <= 8) ? h : (h >> 1);

    // partition classes of :1268-1283
    const int wh = (w << 8) | h;
#define ME_IS(a, b) (wh == (((a) << 8) | (b)))
    const bool smallRect = ME_IS(4, 4) || ME_IS(16, 12) || ME_IS(12, 16) || ME_IS(16, 4) || ME_IS(4, 16);
    const bool verticalRect = ME_IS(32, 64) || ME_IS(16, 32) || ME_IS(8, 16) || ME_IS(4, 8);
    const bool horizontalRect = ME_IS(64, 32) || ME_IS(32, 16) || ME_IS(16, 8) || ME_IS(8, 4);
    const bool asymV = ME_IS(12, 16) || ME_IS(4, 16) || ME_IS(24, 32) || ME_IS(8, 32) || ME_IS(48, 64) || ME_IS(16, 64);
    const bool asymH = ME_IS(16, 12) || ME_IS(16, 4) || ME_IS(32, 24) || ME_IS(32, 8) || ME_IS(64, 48) || ME_IS(64, 16);
    // pu[partEnum].ads (pixel.cpp:1105-1129): ads_x1 / ads_x2 / ads_x4 <lx, ly>
    const int adsKind = (ME_IS(4, 4) || ME_IS(8, 8) || ME_IS(16, 12) || ME_IS(12, 16) || ME_IS(16, 4) || ME_IS(4, 16)) ? 1
                        : (verticalRect || horizontalRect) ? 2 : 4;
    // deltaY counts rows of the integral plane for these shapes only (:1349-1355)
    const bool rowDelta = ME_IS(64, 64) || ME_IS(32, 32) || ME_IS(16, 16) || verticalRect || asymV;
#undef ME_IS
    const int lxHalf = w >> 1;

    // tempPartEnum (:1285-1301): the sub-block whose pixel sums form encDC
    int tw, th;
    if (verticalRect) { tw = w; th = h >> 1; }
    else if (horizontalRect) { tw = w >> 1; th = h; }
    else if (asymV || asymH) { tw = smallRect ? w : (w >> 1); th = smallRect ? h : (h >> 1); }
    else { tw = (w <= 8) ? w : (w >> 1); th = (w <= 8) ? h : (h >> 1); }

    // encDC = sad_x4(zero, fenc, fenc + deltaX, fenc + deltaY*64, fenc + deltaX + deltaY*64) (:1305-1311).  Where a
    // sub-block leaves the PU the reference reads whatever an earlier PU left in fencPUYuv; this backend defines those
    // pixels as 0 (the cache is conceptually cleared by setSourcePU).
    int encDC[4];
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
        const int bx = (k & 1) ? deltaX : 0, by = (k & 2) ? deltaY : 0;
        int acc = 0;
        for (int e = lane; e < tw * th; e += 32)
        {
            const int y = by + e / tw, x = bx + e % tw;
            if (x < w && y < h) acc += (int)s.fenc[y * 64 + x];
        }
        encDC[k] = warp_sum(acc);
    }

    // integral plane (:1313-1347)
    int plane;
    switch (deltaX)
    {
    case 32: plane = (deltaY % 24 == 0) ? 1 : (deltaY == 8) ? 2 : 0; break;
    case 24: plane = 3; break;
    case 16: plane = (deltaY % 12 == 0) ? 5 : (deltaY == 4) ? 6 : 4; break;
    case 12: plane = 7; break;
    case 8:  plane = (deltaY == 32) ? 8 : 9; break;
    case 4:  plane = (deltaY == 16) ? 10 : 11; break;
    default: plane = 11; break;
    }
    const uint32_t* sumsBase = s.integral[plane] + s.integralOff;
    int64_t delta = rowDelta ? (int64_t)deltaY * s.stride : (int64_t)deltaY;
    if (verticalRect) encDC[1] = encDC[2];
    if (horizontalRect) delta = deltaX;

    for (int ty = minY; ty <= maxY; ty++)
    {
        // p_cost_mvy = m_cost_mvy - qmvp.y is indexed with the FULL-PEL y and scaled by 4 (:1250, :1363)
        const int ycost = cost_at(s, ty - 2 * s.mvpy) << 2;
        if (S.bcost <= ycost) continue;
        int bc = S.bcost - ycost;
        const int thresh = bc;
        const uint32_t* sums = sumsBase + minX + (int64_t)ty * s.stride;

        // ads() of the row (pixel.cpp:121-165): which x offsets survive, and how many (xn)
        int xn = 0;
        for (int base = 0; base < width; base += 32)
        {
            const int i = base + lane;
            bool pass = false;
            if (i < width)
            {
                int64_t ads = llabs((int64_t)encDC[0] - (int64_t)sums[i]);
                if (adsKind == 2) ads += llabs((int64_t)encDC[1] - (int64_t)sums[i + delta]);
                if (adsKind == 4)
                    ads += llabs((int64_t)encDC[1] - (int64_t)sums[i + lxHalf]) + llabs((int64_t)encDC[2] - (int64_t)sums[i + delta]) +
                           llabs((int64_t)encDC[3] - (int64_t)sums[i + delta + lxHalf]);
                ads += cost_at(s, ((minX + i) << 2) - s.mvpx);                   // fpelCostMvX[minX + i] (bitcost.cpp:56-75)
                pass = (int)ads < thresh;
            }
            xn += __popc(__ballot_sync(0xffffffffu, pass));
        }
        const int nx3 = (xn / 3) * 3;        // survivors costed by COST_MV_X3_ABS; the last xn % 3 go through COST_MV (:1383-1390)

        int k = 0;
        bool restored = false;
        for (int base = 0; base < width && k < xn; base += 32)
        {
            const int i = base + lane;
            bool pass = false;
            if (i < width)
            {
                int64_t ads = llabs((int64_t)encDC[0] - (int64_t)sums[i]);
                if (adsKind == 2) ads += llabs((int64_t)encDC[1] - (int64_t)sums[i + delta]);
                if (adsKind == 4)
                    ads += llabs((int64_t)encDC[1] - (int64_t)sums[i + lxHalf]) + llabs((int64_t)encDC[2] - (int64_t)sums[i + delta]) +
                           llabs((int64_t)encDC[3] - (int64_t)sums[i + delta + lxHalf]);
                ads += cost_at(s, ((minX + i) << 2) - s.mvpx);
                pass = (int)ads < thresh;
            }
            unsigned m = __ballot_sync(0xffffffffu, pass);
            while (m)
            {
                int ox[4] = { 0, 0, 0, 0 }, oy[4] = { ty, ty, ty, ty }, c[4];
                int cnt = 0;
                while (m && cnt < 4) { const int b = __ffs(m) - 1; m &= m - 1; ox[cnt++] = minX + base + b; }
                warp_sad_k<pixel>(s, cnt, ox, oy, c);
                for (int q = 0; q < cnt; q++, k++)
                {
                    if (k < nx3)
                    {
                        // COST_MV_X3_ABS (:300-313): x cost only, through the doubly-offset p_cost_mvx; compared with bcost - ycost
                        const int cost = c[q] + cost_at(s, (ox[q] << 2) - 2 * s.mvpx);
                        if (cost < bc) { bc = cost; S.bmv = mv2(ox[q], ty); }
                    }
                    else
                    {
                        if (!restored) { bc += ycost; restored = true; }
                        const int cost = c[q] + S.fcost(ox[q], ty);                // COST_MV (:238-244)
                        if (cost < bc) { bc = cost; S.bmv = mv2(ox[q], ty); }
                    }
                }
            }
        }
        if (!restored) bc += ycost;
        S.bcost = bc;
    }
}
#endif

// Full MotionEstimate::motionEstimate.  Returns the cost; (outx,outy) = outQMv.
template<typename pixel>
__device__ int motion_estimate(const MEState<pixel>& s, MV2 mvmin, MV2 mvmax, MV2 qmvp, int numCand, const int* mvc /* [numCand][2] */,
                               int merange, int searchMethod, int subpelRefine, int maxSlices, int partEnumIs64, int& outx, int& outy)
{
#ifdef ME_LOWRES_ONLY
    numCand = 0; maxSlices = 1;      // the lookahead passes neither MV candidates nor slice bounds (slicetype.cpp:3313-3315)
#endif
    MESearch<pixel> S(s);
    S.mvmin = mvmin; S.mvmax = mvmax;
    const MV2 qmvmin = mv2(mvmin.x << 2, mvmin.y << 2), qmvmax = mv2(mvmax.x << 2, mvmax.y << 2);

    // measure SAD cost at clipped QPEL MVP (motion.cpp:770-778)
    MV2 pmv = mv2(max(min(qmvp.x, qmvmax.x), qmvmin.x), max(min(qmvp.y, qmvmax.y), qmvmin.y));
    MV2 bestpre = pmv;
    int bprecost = ME_IS_LOWRES(s) ? lowres_qpel_cost<pixel>(s, pmv.x, pmv.y, false) : subpel_compare<pixel>(s, pmv.x, pmv.y, false);

    // re-measure full pel rounded MVP with SAD as search start point (:780-784)
    S.bmv = mv2((pmv.x + 2) >> 2, (pmv.y + 2) >> 2);
    S.bcost = bprecost;
    if ((pmv.x | pmv.y) & 3)
        S.bcost = S.costAt(S.bmv.x, S.bmv.y);

    // refineMV (:606-737) measures neither the zero MV nor candidates: predictor, square refine, workload[5] subpel
    const bool refineOnly = searchMethod == ME_REFINE;
    if (refineOnly) numCand = 0;

    // measure SAD cost at MV(0) if MVP is not zero (:786-796)
    if ((pmv.x | pmv.y) && !refineOnly)
    {
#if defined(ME_WINDOW_CHECK)
        int cost = S.costAt(0, 0);                             // the window test of every candidate covers it
#elif defined(ME_REF_IN_SMEM)
        // the zero-MV block may lie outside the staged window: read it from the global plane
        int cost = group_sum<pixel>(s, thread_sad_anyspace<pixel>(s, s.gfref, s.gstride)) + mvcost(s, 0, 0);
#else
        int cost = warp_sad_block<pixel>(s, s.gfref, s.gstride) + mvcost(s, 0, 0);
#endif
        if (cost < S.bcost)
        {
            S.bcost = cost;
            S.bmv = mv2(0, max(min(0, mvmax.y), mvmin.y));      // quirk: only y is clamped (:794)
        }
    }

    // each QPEL candidate (:799-812)
    for (int i = 0; i < numCand; i++)
    {
        MV2 m = mv2(max(min(mvc[2 * i], qmvmax.x), qmvmin.x), max(min(mvc[2 * i + 1], qmvmax.y), qmvmin.y));
        bool notZero = (m.x | m.y) != 0, nePmv = (m.x != pmv.x) | (m.y != pmv.y), nePre = (m.x != bestpre.x) | (m.y != bestpre.y);
        if (notZero & nePmv & nePre)
        {
            int cost = subpel_compare<pixel>(s, m.x, m.y, false) + mvcost(s, m.x, m.y);
            if (cost < bprecost) { bprecost = cost; bestpre = m; }
        }
    }

    pmv = mv2((pmv.x + 2) >> 2, (pmv.y + 2) >> 2);
    MV2 omv = S.bmv;

    switch (searchMethod)
    {
    case ME_DIA:      // motion.cpp:820-843
    {
        S.bcost <<= 4;
        int i = merange;
        do
        {
            int costs[4];
            const int dx[4] = { 0, 0, -1, 1 }, dy[4] = { -1, 1, 0, 0 };
            S.dirCosts(4, dx, dy, costs);
            if (S.yOk(S.bmv.y - 1) && (costs[0] << 4) + 1 < S.bcost) S.bcost = (costs[0] << 4) + 1;
            if (S.yOk(S.bmv.y + 1) && (costs[1] << 4) + 3 < S.bcost) S.bcost = (costs[1] << 4) + 3;
            if ((costs[2] << 4) + 4 < S.bcost) S.bcost = (costs[2] << 4) + 4;
            if ((costs[3] << 4) + 12 < S.bcost) S.bcost = (costs[3] << 4) + 12;
            if (!(S.bcost & 15)) break;
            S.bmv.x -= ((int)((unsigned)S.bcost << 28)) >> 30;
            S.bmv.y -= ((int)((unsigned)S.bcost << 30)) >> 30;
            S.bcost &= ~15;
        }
        while (--i && S.inRange(S.bmv.x, S.bmv.y));
        S.bcost >>= 4;
        break;
    }
    case ME_HEX:
        S.hexSearch(merange);
        break;
    case ME_REFINE:       // motion.cpp:644-663
        S.squareRefine();
        break;
    case ME_UMH:      // motion.cpp:946-1130
    {
        int ucost1, ucost2;
        int cross_start = 1;
        bool done = false;
        omv = S.bmv;
        ucost1 = S.bcost;
        S.costMvX4(mv2(pmv.x, pmv.y), 0, -1, 0, 1, -1, 0, 1, 0);               // DIA1_ITER(pmv)
        if (pmv.x | pmv.y) S.costMvX4(mv2(0, 0), 0, -1, 0, 1, -1, 0, 1, 0);
        ucost2 = S.bcost;
        if ((S.bmv.x | S.bmv.y) && ((S.bmv.x != pmv.x) | (S.bmv.y != pmv.y)))
            S.costMvX4(S.bmv, 0, -1, 0, 1, -1, 0, 1, 0);
        if (S.bcost == ucost2) cross_start = 3;

        omv = S.bmv;
#define ME_SAD_THRESH(v) (S.bcost < (((v) >> 4) * s.partSizeScale))
        if (S.bcost == ucost2 && ME_SAD_THRESH(2000))
        {
            S.costMvX4(omv, 0, -2, -1, -1, 1, -1, -2, 0);
            S.costMvX4(omv, 2, 0, -1, 1, 1, 1, 0, 2);
            if (S.bcost == ucost1 && ME_SAD_THRESH(500)) done = true;
            else if (S.bcost == ucost2)
            {
                int range = (int)(int16_t)((merange >> 1) | 1);
                S.cross(omv, 3, range, range);
                S.costMvX4(omv, -1, -2, 1, -2, -2, -1, 2, -1);
                S.costMvX4(omv, -2, 1, 2, 1, -1, 2, 1, 2);
                if (S.bcost == ucost2) done = true;
                else cross_start = range + 2;
            }
        }
        if (done) break;

        if (numCand)      // adaptive range (:989-1040)
        {
            int mvd, denom = 1;
            if (numCand == 1)
            {
                if (partEnumIs64) mvd = 25;
                else mvd = abs(qmvp.x - mvc[0]) + abs(qmvp.y - mvc[1]);
            }
            else
            {
                denom = numCand - 1;
                mvd = 0;
                if (!partEnumIs64) { mvd = abs(qmvp.x - mvc[0]) + abs(qmvp.y - mvc[1]); denom++; }
                for (int i = 0; i < numCand - 1; i++)
                    mvd += abs(mvc[2 * i] - mvc[2 * i + 2]) + abs(mvc[2 * i + 1] - mvc[2 * i + 3]);
            }
            int sad_ctx = ME_SAD_THRESH(1000) ? 0 : ME_SAD_THRESH(2000) ? 1 : ME_SAD_THRESH(4000) ? 2 : 3;
            int mvd_ctx = mvd < 10 * denom ? 0 : mvd < 20 * denom ? 1 : mvd < 40 * denom ? 2 : 3;
            merange = (merange * c_rangeMul[mvd_ctx][sad_ctx]) >> 2;
        }
#undef ME_SAD_THRESH
        S.cross(omv, cross_start, merange, merange >> 1);
        S.costMvX4(omv, -2, -2, -2, 2, 2, -2, 2, 2);

        // hexagon grid (:1047-1126)
        omv = S.bmv;
        int i = 1;
        do
        {
            if (4 * i > min(min(mvmax.x - omv.x, omv.x - mvmin.x), min(mvmax.y - omv.y, omv.y - mvmin.y)))
            {
                for (int j = 0; j < 16; j++)
                {
                    int mx = omv.x + c_hex4[j][0] * i, my = omv.y + c_hex4[j][1] * i;
                    if (S.inRange(mx, my)) S.costMv(mx, my);
                }
            }
            else
            {
                int dir = 0;
                for (int k = 0; k < 16; k++)
                {
                    int hx = c_hex4[k][0], hy = c_hex4[k][1];
                    int mx = omv.x + hx * i, my = omv.y + hy * i;
                    int cost = S.costAt(mx, my);
                    // MIN_MV checks the UNSCALED dy (quirk, :1078)
                    if (S.yOk(omv.y + hy) && cost < S.bcost) { S.bcost = cost; dir = hx * 16 + (hy & 15); }
                }
                if (dir)
                {
                    S.bmv.x = omv.x + i * (dir >> 4);
                    S.bmv.y = omv.y + i * (((int)((unsigned)dir << 28)) >> 28);
                }
            }
        }
        while (++i <= merange >> 2);
        if (S.inRange(S.bmv.x, S.bmv.y)) S.hexSearch(merange);
        break;
    }
    case ME_STAR:     // motion.cpp:1132-1240
    {
        int bPointNr = 0, bDistance = 0;
        S.starPattern(bPointNr, bDistance, 3, merange);
        bool stop = false;
        if (bDistance == 1)
        {
            if (bPointNr)
            {
                int saved = S.bcost;
                S.twoPoint(bPointNr);
                if (S.bcost == saved) stop = true;
            }
            else stop = true;
        }
        if (stop) break;
        const int RasterDistance = 5;
        if (bDistance > RasterDistance)
        {
            for (int ty = mvmin.y; ty <= mvmax.y; ty += RasterDistance)
                for (int tx = mvmin.x; tx <= mvmax.x; tx += RasterDistance)
                {
                    if (tx + RasterDistance * 3 <= mvmax.x)
                    {
                        int rx[4] = { tx, tx + 5, tx + 10, tx + 15 }, ry[4] = { ty, ty, ty, ty }, rc[4];
                        warp_sad_k<pixel>(s, 4, rx, ry, rc);
                        int c0 = rc[0], c1 = rc[1], c2 = rc[2], c3 = rc[3];
                        c0 += S.fcost(tx, ty);
                        if (c0 < S.bcost) { S.bcost = c0; S.bmv = mv2(tx, ty); }
                        tx += RasterDistance;
                        c1 += S.fcost(tx, ty);
                        if (c1 < S.bcost) { S.bcost = c1; S.bmv = mv2(tx, ty); }
                        tx += RasterDistance;
                        c2 += S.fcost(tx, ty);
                        if (c2 < S.bcost) { S.bcost = c2; S.bmv = mv2(tx, ty); }
                        tx += RasterDistance;
                        c3 += mvcost(s, tx << 3, ty << 3);                   // quirk: << 3 (motion.cpp:1196)
                        if (c3 < S.bcost) { S.bcost = c3; S.bmv = mv2(tx, ty); }
                    }
                    else
                        S.costMv(tx, ty);
                }
        }
        while (bDistance > 0)
        {
            bDistance = 0; bPointNr = 0;
            S.starPattern(bPointNr, bDistance, 32, merange);
            if (bDistance == 1)
            {
                if (!bPointNr) break;
                S.twoPoint(bPointNr);
                break;
            }
        }
        break;
    }
#ifndef ME_FORCE_THREAD
    case ME_SEA:      // motion.cpp:1242-1395
        sea_search<pixel>(S, s, omv, merange);
        break;
#endif
    case ME_FULL:     // motion.cpp:1397-1441 (non-HME)
    {
        for (int ty = mvmin.y; ty <= mvmax.y; ty++)
            for (int tx = mvmin.x; tx <= mvmax.x; tx++)
                S.costMv(tx, ty);       // the x4 grouping of the reference evaluates the same points in the same order
        break;
    }
    default:
        break;
    }

    // choose between the search result and the best predictor (:1448-1454)
    MV2 bmv; int bcost;
    if (bprecost < S.bcost) { bmv = bestpre; bcost = bprecost; }
    else { bmv = mv2(S.bmv.x << 2, S.bmv.y << 2); bcost = S.bcost; }

    const SubpelWL wl = c_workload[refineOnly ? 5 : subpelRefine];          // refineMV: fixed workload[5] (:673)

    // slice bound clamp (:1458-1463)
    if ((maxSlices > 1) & !refineOnly & ((bmv.y < qmvmin.y) | (bmv.y > qmvmax.y)))
    {
        bmv.y = min(max(bmv.y, qmvmin.y), qmvmax.y);
        bcost = subpel_compare<pixel>(s, bmv.x, bmv.y, true) + mvcost(s, bmv.x, bmv.y);
    }

    if (!bcost && !refineOnly)
        bcost = mvcost(s, bmv.x, bmv.y);                                  // :1465-1470 (refineMV has no such exit)
    else if (ME_IS_LOWRES(s))                                             // :1471-1503
    {
        int bdir = 0;
        for (int i = 1; i <= wl.hpel_dirs; i++)
        {
            int qx = bmv.x + c_square1[i][0] * 2, qy = bmv.y + c_square1[i][1] * 2;
            if ((qy < qmvmin.y) | (qy > qmvmax.y)) continue;
            int cost = lowres_qpel_cost<pixel>(s, qx, qy, false) + mvcost(s, qx, qy);
            if (cost < bcost) { bcost = cost; bdir = i; }
        }
        bmv.x += c_square1[bdir][0] * 2; bmv.y += c_square1[bdir][1] * 2;
        bcost = lowres_qpel_cost<pixel>(s, bmv.x, bmv.y, true) + mvcost(s, bmv.x, bmv.y);
        bdir = 0;
        for (int i = 1; i <= wl.qpel_dirs; i++)
        {
            int qx = bmv.x + c_square1[i][0], qy = bmv.y + c_square1[i][1];
            if ((qy < qmvmin.y) | (qy > qmvmax.y)) continue;
            int cost = lowres_qpel_cost<pixel>(s, qx, qy, true) + mvcost(s, qx, qy);
            if (cost < bcost) { bcost = cost; bdir = i; }
        }
        bmv.x += c_square1[bdir][0]; bmv.y += c_square1[bdir][1];
    }
    else                                                                  // :1504-1561
    {
        bool hpelSatd = wl.hpel_satd != 0;
        if (hpelSatd)
            { const int mc = mvcost(s, bmv.x, bmv.y); bcost = subpel_compare<pixel>(s, bmv.x, bmv.y, true) + mc; }
        for (int iter = 0; iter < wl.hpel_iters; iter++)
        {
            int bdir = 0;
            for (int i = 1; i <= wl.hpel_dirs; i++)
            {
                int qx = bmv.x + c_square1[i][0] * 2, qy = bmv.y + c_square1[i][1] * 2;
                if ((qy < qmvmin.y) | (qy > qmvmax.y)) continue;
                int cost = mvcost(s, qx, qy);                           // first: its table loads are in flight under the interpolation
                cost += subpel_compare<pixel>(s, qx, qy, hpelSatd);
                if (cost < bcost) { bcost = cost; bdir = i; }
            }
            if (bdir) { bmv.x += c_square1[bdir][0] * 2; bmv.y += c_square1[bdir][1] * 2; }
            else break;
        }
        if (!hpelSatd)
            { const int mc = mvcost(s, bmv.x, bmv.y); bcost = subpel_compare<pixel>(s, bmv.x, bmv.y, true) + mc; }
        for (int iter = 0; iter < wl.qpel_iters; iter++)
        {
            int bdir = 0;
            for (int i = 1; i <= wl.qpel_dirs; i++)
            {
                int qx = bmv.x + c_square1[i][0], qy = bmv.y + c_square1[i][1];
                if ((qy < qmvmin.y) | (qy > qmvmax.y)) continue;
                int cost = mvcost(s, qx, qy);                           // first: its table loads are in flight under the interpolation
                cost += subpel_compare<pixel>(s, qx, qy, true);
                if (cost < bcost) { bcost = cost; bdir = i; }
            }
            if (bdir) { bmv.x += c_square1[bdir][0]; bmv.y += c_square1[bdir][1]; }
            else break;
        }
    }
    outx = bmv.x; outy = bmv.y;
    return bcost;
}

} // inline namespace ME_VARIANT
} // namespace x265b200

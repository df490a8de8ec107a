// interp_cell.cuh -- 8-bit luma pp interpolation of the batched entry (x265b200_interp_dev; measured 2.78x, profiles/r02_staged_ab.txt):
// (luma_hpp / luma_vpp / luma_hvpp) with ONE THREAD PER 4x4 OUTPUT CELL on packed words (subpel_packed.cuh) instead of one
// thread per pixel with byte loads: ~10 / 12 / 26 instructions per pixel against ~45 / 45 / 115.
// The whole per-thread function lives here so that the same source runs on the host (tests/test_interp_cell_cpu.py executes
// it thread by thread against the oracle, with the alignment of every 32-bit access asserted); the kernel in
// interp_kernels.cu only turns (blockIdx, threadIdx) into the thread number.
#pragma once
#include "subpel_packed.cuh"
#include "x265b200.h"

namespace x265b200 {

struct InterpArgs
{
    const void* src; int64_t srcStride;
    void* dst;       int64_t dstStride;
    const x265b200_interp_job* jobs; int64_t n;
    int kind, taps, depth, w, h, isRowExt;
};

#if defined(INTERP_CELL_HOST_TEST)
void ic_check_access(const void* p, int store);            // host harness: aborts on a misaligned or out-of-buffer 32-bit access
SP_FN uint32_t ic_ld32(const uint32_t* p) { ic_check_access(p, 0); return *p; }
SP_FN void ic_st32(uint32_t* p, uint32_t v) { ic_check_access(p, 1); *p = v; }
#elif defined(__CUDA_ARCH__)
SP_FN uint32_t ic_ld32(const uint32_t* p) { return __ldg(p); }
SP_FN void ic_st32(uint32_t* p, uint32_t v) { *p = v; }
#else
SP_FN uint32_t ic_ld32(const uint32_t* p) { return *p; }
SP_FN void ic_st32(uint32_t* p, uint32_t v) { *p = v; }
#endif

// NW words of pixels starting at any byte address (aligned word loads + funnel shift; the word after an aligned run is not read)
template<int NW> SP_FN void ic_words(const uint8_t* p, uint32_t out[NW])
{
    const uintptr_t a = (uintptr_t)p;
    const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3) * 8;
    uint32_t t[NW + 1];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < NW; i++) t[i] = ic_ld32(w + i);
    t[NW] = sh ? ic_ld32(w + NW) : 0u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < NW; i++) out[i] = sp_funnel_r(t[i], t[i + 1], sh);
}

// thread number g -> (job, 4x4 cell); filt = the luma filter table (constants.cpp:250-268 values, tables.cuh)
SP_FN void interp_pp8_cell_thread(const InterpArgs& p, int64_t g, const int16_t (*filt)[8])
{
    const int cellsX = p.w >> 2, cells = cellsX * (p.h >> 2);
    const int64_t jidx = g / cells;
    if (jidx >= p.n) return;
    const int cell = (int)(g - jidx * cells), cy = cell / cellsX, cx = cell - cy * cellsX;
    const x265b200_interp_job job = p.jobs[jidx];
    const uint8_t* src = (const uint8_t*)p.src + job.srcOff + (int64_t)(cy * 4) * p.srcStride + cx * 4;
    uint8_t* dst = (uint8_t*)p.dst + job.dstOff + (int64_t)(cy * 4) * p.dstStride + cx * 4;
    uint32_t out[4];
    if (p.kind == X265B200_IP_HPP)
    {
        const uint32_t clo = sp_taps(filt[job.idxX], 0), chi = sp_taps(filt[job.idxX], 4);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int r = 0; r < 4; r++)
        {
            uint32_t w[3];
            ic_words<3>(src + (int64_t)r * p.srcStride - 3, w);
            out[r] = hpp_row4_u8(w, clo, chi);
        }
    }
    else if (p.kind == X265B200_IP_VPP)
    {
        const uint32_t clo = sp_taps(filt[job.idxX], 0), chi = sp_taps(filt[job.idxX], 4);      // single-pass kinds take coeffIdx in idxX
        uint32_t r[11];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 11; j++) ic_words<1>(src + (int64_t)(j - 3) * p.srcStride, &r[j]);
        vpp_cell_u8(r, clo, chi, out);
    }
    else    // X265B200_IP_HVPP
    {
        const uint32_t clo = sp_taps(filt[job.idxX], 0), chi = sp_taps(filt[job.idxX], 4);
        uint32_t w[11][3];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 11; j++) ic_words<3>(src + (int64_t)(j - 3) * p.srcStride - 3, w[j]);
        hvpp_cell_u8(w, clo, chi, filt[job.idxY], out);
    }
    if ((((uintptr_t)dst | (uintptr_t)p.dstStride) & 3) == 0)
    {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int r = 0; r < 4; r++) ic_st32((uint32_t*)(dst + (int64_t)r * p.dstStride), out[r]);
    }
    else
    {
        for (int r = 0; r < 4; r++)
            for (int k = 0; k < 4; k++) dst[(int64_t)r * p.dstStride + k] = (uint8_t)(out[r] >> (8 * k));
    }
}

} // namespace x265b200

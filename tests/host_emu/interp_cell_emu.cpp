// interp_cell_emu.cpp -- TEST INFRASTRUCTURE: runs the staged cell form of the 8-bit luma interpolation (csrc/interp_cell.cuh,
// the function the kernel calls with its thread number) for every thread of the grid on the host; every 32-bit access is
// checked for alignment and for staying inside the source / destination buffers.
#define INTERP_CELL_HOST_TEST 1
#include "interp_cell.cuh"
#include "tables.cuh"
#include <stdlib.h>
#include <stdio.h>

namespace x265b200 {
static const unsigned char *g_srcLo, *g_srcHi, *g_dstLo, *g_dstHi;
static int g_fault = 0;
void ic_check_access(const void* p, int store)
{
    const unsigned char* q = (const unsigned char*)p;
    if ((uintptr_t)p & 3) g_fault |= 1;
    if (store) { if (q < g_dstLo || q + 4 > g_dstHi) g_fault |= 2; }
    else       { if (q < g_srcLo || q + 4 > g_srcHi) g_fault |= 4; }
    if (g_fault) { fprintf(stderr, "interp_cell_emu: bad access (fault %d, store %d)\n", g_fault, store); abort(); }
}
}

extern "C" int ic_run(const void* src, size_t srcBytes, int64_t srcStride, void* dst, size_t dstBytes, int64_t dstStride,
                      const x265b200_interp_job* jobs, int64_t n, int kind, int w, int h)
{
    using namespace x265b200;
    g_srcLo = (const unsigned char*)src; g_srcHi = g_srcLo + srcBytes;
    g_dstLo = (const unsigned char*)dst; g_dstHi = g_dstLo + dstBytes;
    InterpArgs a; a.src = src; a.srcStride = srcStride; a.dst = dst; a.dstStride = dstStride; a.jobs = jobs; a.n = n;
    a.kind = kind; a.taps = 8; a.depth = 8; a.w = w; a.h = h; a.isRowExt = 0;
    const int64_t threads = n * (w >> 2) * (h >> 2), blocks = (threads + 127) / 128;
    for (int64_t g = 0; g < blocks * 128; g++)                    // the kernel: one call per thread of the grid, tail threads included
        interp_pp8_cell_thread(a, g, kLumaFilter);
    return g_fault;
}

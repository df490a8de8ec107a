// intra_cell_emu.cpp -- TEST INFRASTRUCTURE: runs the staged cell form of the 8-bit all-modes intra prediction
// (csrc/intra_cell.cuh, the function the kernel calls with its thread number) for every thread of the grid on the host; every
// access is checked for alignment and for staying inside the neighbour / destination buffers.
#define INTRA_CELL_HOST_TEST 1
#include "intra_cell.cuh"
#include <stdlib.h>
#include <stdio.h>

namespace x265b200 {
static const unsigned char *g_lo[3], *g_hi[3];
void xc_check(const void* p, int bytes, int store)
{
    const unsigned char* q = (const unsigned char*)p;
    int fault = 0;
    if (bytes == 4 && ((uintptr_t)p & 3)) fault |= 1;
    if (store) { if (q < g_lo[2] || q + bytes > g_hi[2]) fault |= 2; }
    // loads: an ALIGNED word may hang over either end of the array by up to 3 bytes (it still holds at least one byte of it --
    // the same aligned-word-plus-funnel-shift loads every kernel of the library uses); anything further out is a fault
    else if (!((q + bytes > g_lo[0] && q < g_hi[0]) || (q + bytes > g_lo[1] && q < g_hi[1]))) fault |= 4;
    if (fault) { fprintf(stderr, "intra_cell_emu: bad access (fault %d, %d bytes, store %d)\n", fault, bytes, store); abort(); }
}
}

extern "C" int xc_run(const uint8_t* raw, const uint8_t* filt, size_t nbrBytes, uint8_t* dest, size_t destBytes, int log2N, int bLuma, int all35, int64_t n)
{
    using namespace x265b200;
    g_lo[0] = raw; g_hi[0] = raw + nbrBytes; g_lo[1] = filt; g_hi[1] = filt + nbrBytes; g_lo[2] = dest; g_hi[2] = dest + destBytes;
    IntraCellArgs a; a.raw = raw; a.filt = filt; a.dest = dest; a.log2N = log2N; a.bLuma = bLuma; a.all35 = all35; a.n = n;
    const int N = 1 << log2N;
    const int64_t threads = n * (all35 ? 35 : 33) * ((N * N) >> 4), blocks = (threads + 127) / 128;
    for (int64_t g = 0; g < blocks * 128; g++) intra_modes8_cell_thread(a, g);
    return 0;
}

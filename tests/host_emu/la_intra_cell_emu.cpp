// la_intra_cell_emu.cpp -- TEST INFRASTRUCTURE: the per-cell prediction + cost functions of la_intra_kernel
// (csrc/lookahead_kernels.cu, between the [host-testable] markers; pasted into la_intra_cell_src.inc by the test) on the host.
// The four lanes of a CU are run in turn: first every lane's share of the projected reference line (the code before the first
// __syncwarp of la_intra_ang_cell_cost), then every lane's cost.
#include <cstdint>
#include <cstdlib>
#include <algorithm>
using std::min; using std::max;
#define __device__
#define __forceinline__ inline
#define __noinline__
static int g_phase = 0;
#define __syncwarp() do { if (g_phase == 0) return -12345; } while (0)
inline int clip3i(int a, int b, int v) { return v < a ? a : (v > b ? b : v); }
inline void me_hadamard4(int& a, int& b, int& c, int& d) { int s01 = a + b, d01 = a - b, s23 = c + d, d23 = c - d; a = s01 + s23; b = d01 + d23; c = s01 - s23; d = d01 - d23; }
struct LAIntraArgs;
#include "la_intra_cell_src.inc"

static const int kTabs[25] = { -32, -26, -21, -17, -13, -9, -5, -2, 0, 2, 5, 9, 13, 17, 21, 26, 32, 4096, 1638, 910, 630, 482, 390, 315, 256 };

// cost of one mode on an 8x8 CU as the kernel computes it: sum of the four lanes' 4x4 cell costs.
// smp / flt: the 33 raw / 1:2:1-filtered neighbours (ints), fenc: 64 source pixels (row-major), dc: the DC value (mode 1)
extern "C" int emu_la_intra_mode_cost(const int* smp, const int* flt, const int* fenc, int mode, int dc, int depth)
{
    int cell[4][16];
    for (int c = 0; c < 4; c++)
        for (int i = 0; i < 4; i++)
            for (int k = 0; k < 4; k++) cell[c][i * 4 + k] = fenc[((c >> 1) * 4 + i) * 8 + (c & 1) * 4 + k];
    int total = 0;
    if (mode == 1) { for (int c = 0; c < 4; c++) total += la_intra_cell_cost(smp, 1, 1, dc, depth, (c & 1) * 4, (c >> 1) * 4, cell[c]); return total; }
    if (mode == 0) { for (int c = 0; c < 4; c++) total += la_intra_cell_cost(flt, 0, 0, 0, depth, (c & 1) * 4, (c >> 1) * 4, cell[c]); return total; }
    int line[26];
    g_phase = 0;
    for (int c = 0; c < 4; c++) la_intra_ang_cell_cost(smp, flt, line, kTabs, mode, depth, c, (c & 1) * 4, (c >> 1) * 4, cell[c]);
    g_phase = 1;
    for (int c = 0; c < 4; c++) total += la_intra_ang_cell_cost(smp, flt, line, kTabs, mode, depth, c, (c & 1) * 4, (c >> 1) * 4, cell[c]);
    return total;
}

// intra_cell.cuh -- STAGED (off unless X265B200_INTRA_FAST=1; not yet run on a GPU): the 8-bit all-modes intra prediction
// (x265b200_intra_modes_dev / x265b200_intra_allangs_dev, N = 8 / 16 / 32) with ONE THREAD PER 16 OUTPUT BYTES and no shared
// memory: an angular row segment is a 2-tap interpolation of CONSECUTIVE entries of the projected reference line
// (intra_pred_ang_c, intrapred.cpp:102-204); for non-negative offsets that line is a contiguous run of the neighbour array
// itself (top row for vertical modes, left column for horizontal ones -- all_angs_pred_c leaves horizontal modes
// un-transposed, :206-234), so the entries are unaligned word loads; only segments that reach the inverse-angle projection
// (negative angles) gather bytes.  No per-block line build, no barriers, no (mode, entry) -> index table.
// The per-thread function lives here so that the same source runs on the host (tests/test_intra_cell_cpu.py executes it for
// every thread of the grid against the oracle with every access bounds-checked); the kernel in intra_kernels.cu only turns
// (blockIdx, threadIdx) into the thread number.
#pragma once
#include "subpel_packed.cuh"

namespace x265b200 {

struct IntraCellArgs
{
    const uint8_t* raw;     // [n][4N+1] neighbours: topLeft, top 2N, left 2N (intrapred.cpp:36-50 layout)
    const uint8_t* filt;    // [n][4N+1] the same after intraFilter<N> (:31-51)
    uint8_t* dest;          // [n][NM][N*N], NM = 35 (planar, DC, 33 angular) or 33
    int log2N, bLuma, all35;
    int64_t n;
};

#if defined(INTRA_CELL_HOST_TEST)
void xc_check(const void* p, int bytes, int store);        // host harness: aborts on a misaligned or out-of-buffer access
SP_FN uint32_t xc_ld32(const uint32_t* p) { xc_check(p, 4, 0); return *p; }
SP_FN uint32_t xc_ld8(const uint8_t* p) { xc_check(p, 1, 0); return *p; }
SP_FN void xc_st32(uint32_t* p, uint32_t v) { xc_check(p, 4, 1); *p = v; }
SP_FN void xc_st8(uint8_t* p, uint32_t v) { xc_check(p, 1, 1); *p = (uint8_t)v; }
#elif defined(__CUDA_ARCH__)
SP_FN uint32_t xc_ld32(const uint32_t* p) { return __ldg(p); }
SP_FN uint32_t xc_ld8(const uint8_t* p) { return __ldg(p); }
SP_FN void xc_st32(uint32_t* p, uint32_t v) { *p = v; }
SP_FN void xc_st8(uint8_t* p, uint32_t v) { *p = (uint8_t)v; }
#else
SP_FN uint32_t xc_ld32(const uint32_t* p) { return *p; }
SP_FN uint32_t xc_ld8(const uint8_t* p) { return *p; }
SP_FN void xc_st32(uint32_t* p, uint32_t v) { *p = v; }
SP_FN void xc_st8(uint8_t* p, uint32_t v) { *p = (uint8_t)v; }
#endif

SP_FN uint32_t sp_sum4(uint32_t w)                          // sum of the four bytes
{
#if defined(__CUDA_ARCH__)
    return __vsadu4(w, 0u);
#else
    return (w & 0xff) + ((w >> 8) & 0xff) + ((w >> 16) & 0xff) + (w >> 24);
#endif
}

// out[i] = bytes 4i .. 4i+3 of the run that starts at p; only the aligned words that hold bytes [p, p + nbytes) are read
template<int NW> SP_FN void xc_bytes(const uint8_t* p, int nbytes, uint32_t out[NW])
{
    const uintptr_t a = (uintptr_t)p;
    const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3) * 8;
    const int last = (int)(((a + (uintptr_t)nbytes - 1) >> 2) - (a >> 2));
    uint32_t t[NW + 1];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i <= NW; i++) t[i] = i <= last ? xc_ld32(w + i) : 0u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < NW; i++) out[i] = sp_funnel_r(t[i], t[i + 1], sh);
}

SP_FN int intra_cell_angle(int angleOffset)                 // intrapred.cpp:121-124 angle table, angleOffset in -8..8
{
    const int a = angleOffset < 0 ? -angleOffset : angleOffset;
    const int mag = a == 0 ? 0 : (a == 1 ? 2 : (a == 2 ? 5 : (a == 3 ? 9 : (a == 4 ? 13 : (a == 5 ? 17 : (a == 6 ? 21 : (a == 7 ? 26 : 32)))))));
    return angleOffset < 0 ? -mag : mag;
}
SP_FN int intra_cell_inv_angle(int k)                       // :124 invAngle table, k = -angleOffset - 1 in 0..7
{
    return k == 0 ? 4096 : (k == 1 ? 1638 : (k == 2 ? 910 : (k == 3 ? 630 : (k == 4 ? 482 : (k == 5 ? 390 : (k == 6 ? 315 : 256))))));
}

// out[i] = bytes 4i .. 4i+3 of the run that starts at p, of which only bytes [from, nbytes) are wanted: aligned words that hold
// none of them are not read, and the unwanted low bytes come back as zero
template<int NW> SP_FN void xc_bytes_from(const uint8_t* p, int from, int nbytes, uint32_t out[NW])
{
    const intptr_t a = (intptr_t)p;
    const intptr_t w0 = a & ~(intptr_t)3;
    const uint32_t sh = (uint32_t)(a & 3) * 8;
    const int first = (int)(((a + from) >> 2) - (a >> 2)), last = (int)(((a + nbytes - 1) >> 2) - (a >> 2));
    uint32_t t[NW + 1];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i <= NW; i++) t[i] = (i >= first && i <= last) ? xc_ld32((const uint32_t*)(w0 + 4 * i)) : 0u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < NW; i++)
    {
        const int lo = from - 4 * i;                         // bytes below `lo` of this word are not wanted
        const uint32_t keep = lo <= 0 ? 0xFFFFFFFFu : (lo >= 4 ? 0u : 0xFFFFFFFFu << (8 * lo));
        out[i] = sp_funnel_r(t[i], t[i + 1], sh) & keep;
    }
}

// CH (8 or 16) bytes of row y starting at column x of angular mode `mode` predicted from neighbour array arr
template<int CH> SP_FN void intra_cell_ang_seg(const uint8_t* arr, int N, int mode, int bEdge, int y, int x, uint32_t* ow)
{
    constexpr int NW = CH / 4, NE = NW + 1;
    const int N2 = N << 1;
    const bool hor = mode < 18;
    const int angleOffset = hor ? 10 - mode : mode - 26;
    const int angle = intra_cell_angle(angleOffset);
    const uint8_t* mainp = arr + (hor ? N2 : 0);            // ref[k] = mainp[1 + k] for k >= 0 (flipped view for horizontal modes, :111-120)
    const uint8_t* side = arr + (hor ? 0 : N2);             // nb(2N + j) = side[j]
    const int angleSum = (y + 1) * angle, offset = angleSum >> 5;
    const uint32_t f = (uint32_t)(angleSum & 31), g = 32u - f;
    const int first = offset + x, need = CH + (f ? 1 : 0);
    uint32_t e[NE];
    if (first >= 0) xc_bytes<NE>(mainp + 1 + first, need, e);
    else
    {
        // the first ng entries lie left of ref[0]: ref[-1] is the corner, ref[idx < -1] the inverse-angle projection of the side
        // (:146-172); they are gathered byte by byte, the rest of the segment is the contiguous run that starts at ref[0]
        const int ng = -first < need ? -first : need;
        if (ng < need) xc_bytes_from<NE>(mainp + 1 + first, ng, need, e);
        else
        {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int k = 0; k < NE; k++) e[k] = 0;
        }
        const int inv = intra_cell_inv_angle(-angleOffset - 1);
        for (int k = 0; k < ng; k++)
        {
            const int idx = first + k;
            const uint32_t v = xc_ld8(idx == -1 ? arr : side + ((128 + (-1 - idx) * inv) >> 8)) << ((uint32_t)(k & 3) * 8);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int j = 0; j < NE; j++) e[j] |= j == (k >> 2) ? v : 0u;
        }
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < NW; k++)
    {
        const uint32_t A = e[k];
        if (f)
        {
            // two pixels per 32-bit multiply-add in 16-bit lanes: 255 * 32 + 16 < 2^16
            const uint32_t B = sp_funnel_r(e[k], e[k + 1], 8);
            const uint32_t r02 = (((A & 0x00FF00FFu) * g + (B & 0x00FF00FFu) * f + 0x00100010u) >> 5) & 0x00FF00FFu;
            const uint32_t r13 = ((((A >> 8) & 0x00FF00FFu) * g + ((B >> 8) & 0x00FF00FFu) * f + 0x00100010u) >> 5) & 0x00FF00FFu;
            ow[k] = r02 | (r13 << 8);
        }
        else ow[k] = A;
    }
    if (!angle && bEdge && x == 0)
    {
        // pure horizontal / vertical: the first column is edge-filtered for luma (:176-189)
        const int v = (int)xc_ld8(mainp + 1) + (((int)xc_ld8(side + 1 + y) - (int)xc_ld8(arr)) >> 1);
        ow[0] = (ow[0] & 0xFFFFFF00u) | (uint32_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
}

// thread number g -> (block, mode, 16-byte chunk)
SP_FN void intra_modes8_cell_thread(const IntraCellArgs& p, int64_t g)
{
    const int log2N = p.log2N, N = 1 << log2N, N2 = N << 1, LEN = 4 * N + 1;
    const int NM = p.all35 ? 35 : 33, CPM = (N * N) >> 4, CPB = NM * CPM;
    const int64_t b = g / CPB;
    if (b >= p.n) return;
    const int q = (int)(g - b * CPB), mi = q / CPM, r = (q - mi * CPM) << 4;
    const int mode = p.all35 ? mi : mi + 2;
    const uint8_t* raw = p.raw + b * LEN;
    const uint8_t* filt = p.filt + b * LEN;
    const int thr = N == 8 ? 7 : (N == 16 ? 1 : 0);         // constants.cpp:561 g_intraFilterFlags as distance thresholds
    const int CH = N < 16 ? N : 16, nw = CH >> 2, segs = 16 / CH;
    uint32_t ow[4];
    int dc = 0;
    if (mode == 1)
    {
        // intra_pred_dc_c :69-85: dcVal = (N + sum of the N top and N left neighbours) / 2N
        uint32_t sum = 0;
        for (int i = 0; i < N; i += 16)
        {
            uint32_t wa[4], wl[4];
            xc_bytes<4>(raw + 1 + i, N - i < 16 ? N - i : 16, wa);
            xc_bytes<4>(raw + N2 + 1 + i, N - i < 16 ? N - i : 16, wl);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int k = 0; k < 4; k++) if (4 * k < N - i) sum += sp_sum4(wa[k]) + sp_sum4(wl[k]);
        }
        dc = (int)((N + sum) >> (log2N + 1));
    }
    for (int sg = 0; sg < segs; sg++)
    {
        const int rr = r + sg * CH, y = rr >> log2N, x = rr & (N - 1);
        uint32_t* o = ow + sg * nw;
        if (mode == 0)
        {
            // planar_pred_c :87-100, from the smoothed neighbours for N >= 8 (Search::estIntraPredQT, search.cpp:1358-1375)
            const uint8_t* above = filt + 1;
            const uint8_t* left = filt + N2 + 1;
            const int topRight = (int)xc_ld8(above + N), bottomLeft = (int)xc_ld8(left + N), l = (int)xc_ld8(left + y);
            for (int k = 0; k < nw; k++)
            {
                uint32_t word = 0;
                for (int j = 0; j < 4; j++)
                {
                    const int xx = x + 4 * k + j;
                    const int v = ((N - 1 - xx) * l + (N - 1 - y) * (int)xc_ld8(above + xx) + (xx + 1) * topRight + (y + 1) * bottomLeft + N) >> (log2N + 1);
                    word |= (uint32_t)v << (8 * j);
                }
                o[k] = word;
            }
        }
        else if (mode == 1)
        {
            // DC from the raw neighbours, edges filtered when bLuma (dcPredFilter :53-67)
            const uint32_t d4 = (uint32_t)dc * 0x01010101u;
            for (int k = 0; k < nw; k++) o[k] = d4;
            if (p.bLuma)
            {
                const uint8_t* above = raw + 1;
                const uint8_t* left = raw + N2 + 1;
                if (y == 0)
                    for (int k = 0; k < nw; k++)
                    {
                        uint32_t word = 0;
                        for (int j = 0; j < 4; j++) word |= (uint32_t)(((int)xc_ld8(above + x + 4 * k + j) + 3 * dc + 2) >> 2) << (8 * j);
                        o[k] = word;
                    }
                if (x == 0)
                {
                    const int v = y == 0 ? ((int)xc_ld8(above) + (int)xc_ld8(left) + 2 * dc + 2) >> 2 : ((int)xc_ld8(left + y) + 3 * dc + 2) >> 2;
                    o[0] = (o[0] & 0xFFFFFF00u) | (uint32_t)v;
                }
            }
        }
        else
        {
            const int d26 = mode > 26 ? mode - 26 : 26 - mode, d10 = mode > 10 ? mode - 10 : 10 - mode;
            const uint8_t* arr = (d26 < d10 ? d26 : d10) > thr ? filt : raw;
            if (CH == 8) intra_cell_ang_seg<8>(arr, N, mode, p.bLuma, y, x, o);
            else         intra_cell_ang_seg<16>(arr, N, mode, p.bLuma, y, x, o);
        }
    }
    uint8_t* dst = p.dest + ((int64_t)b * NM + mi) * (N * N) + r;
    if (((uintptr_t)dst & 3) == 0)
    {
#if defined(__CUDA_ARCH__)
        if (((uintptr_t)dst & 15) == 0) { *(uint4*)dst = make_uint4(ow[0], ow[1], ow[2], ow[3]); return; }
#endif
        for (int k = 0; k < 4; k++) xc_st32((uint32_t*)dst + k, ow[k]);
    }
    else
        for (int k = 0; k < 16; k++) xc_st8(dst + k, (ow[k >> 2] >> (8 * (k & 3))) & 0xff);
}

} // namespace x265b200

// intra_kernels.cu -- HEVC intra prediction (DC / planar / 33 angular modes, reference smoothing,
// all-angles) as batched sm_100a kernels.  Every output pixel is a closed-form function of the
// 4N+1 neighbour array, so one thread produces one pixel; the neighbour array is staged in smem.
//
// Reference semantics (bit-exact), source/common/intrapred.cpp:
//   intraFilter<N>      :31-51    1:2:1 smoothing of [topLeft, top 2N, left 2N]
//   intra_pred_dc_c     :69-85    dcVal = (N + sum)/(2N), optional edge filter :53-67
//   planar_pred_c       :87-100
//   intra_pred_ang_c    :102-204  horizontal modes = neighbour flip + transpose; angle==0 edge filter
//   all_angs_pred_c     :206-234  33 NxN blocks, horizontal modes left un-transposed,
//                                 filtered/unfiltered neighbours chosen by g_intraFilterFlags
//                                 (constants.cpp:561) = |mode-26|,|mode-10| distance thresholds.
// Roofline: HBM-bound, (4N+1) + N*N pixels per (block, mode).
#include "common.cuh"
#include "x265b200.h"
#include <cstdlib>

namespace x265b200 {

__constant__ int8_t  c_angle[17]   = { -32, -26, -21, -17, -13, -9, -5, -2, 0, 2, 5, 9, 13, 17, 21, 26, 32 };
__constant__ int16_t c_invAngle[8] = { 4096, 1638, 910, 630, 482, 390, 315, 256 };

// neighbours `s` are in the reference layout [0]=topLeft, [1..2N]=top, [2N+1..4N]=left.
// Returns the UNFLIPPED angular prediction at (yy, xx) computed from flipped-or-not neighbours,
// i.e. intra_pred_ang_c :126-189 with srcPix access redirected through nb().
template<typename pixel>
__device__ __forceinline__ int ang_pixel(const pixel* s, int N, int mode, int bFilter, int y, int x, int depth)
{
    const int N2 = N << 1;
    const bool hor = mode < 18;
    // flipped neighbour view (:111-120): nb(0)=s[0]; nb(1+i)=s[N2+1+i]; nb(N2+1+i)=s[1+i]
    auto nb = [&](int i) -> int {
        if (!hor || i == 0) return (int)s[i];
        return i <= N2 ? (int)s[N2 + i] : (int)s[i - N2];
    };
    const int yy = hor ? x : y, xx = hor ? y : x;     // transpose back at the end (:192-203)
    const int angleOffset = hor ? 10 - mode : mode - 26;
    const int angle = c_angle[8 + angleOffset];
    if (!angle)
    {
        if (bFilter && xx == 0)
        {
            int v = (int16_t)(nb(1) + ((nb(N2 + 1 + yy) - nb(0)) >> 1));
            return clip3i(0, (1 << depth) - 1, v);
        }
        return nb(1 + xx);
    }
    const int angleSum = (yy + 1) * angle;
    const int offset = angleSum >> 5, fraction = angleSum & 31;
    // ref[idx] (:146-172)
    auto ref = [&](int idx) -> int {
        if (angle > 0 || idx >= -1) return nb(idx + 1);
        int i = -2 - idx;
        int invAngleSum = 128 + (i + 1) * c_invAngle[-angleOffset - 1];
        return nb(N2 + (invAngleSum >> 8));
    };
    if (fraction)
        return ((32 - fraction) * ref(offset + xx) + fraction * ref(offset + xx + 1) + 16) >> 5;
    return ref(offset + xx);
}

template<typename pixel>
__device__ __forceinline__ int intra_pixel(const pixel* s, int N, int log2N, int mode, int bFilter, int y, int x, int depth, int dcVal)
{
    const int N2 = N << 1;
    if (mode == 0)        // planar :87-100
    {
        const pixel* above = s + 1; const pixel* left = s + N2 + 1;
        return ((N - 1 - x) * left[y] + (N - 1 - y) * above[x] + (x + 1) * above[N] + (y + 1) * left[N] + N) >> (log2N + 1);
    }
    if (mode == 1)        // DC :69-85 (+ dcPredFilter :53-67)
    {
        if (!bFilter) return dcVal;
        const pixel* above = s + 1; const pixel* left = s + N2 + 1;
        if (x == 0 && y == 0) return (above[0] + left[0] + 2 * dcVal + 2) >> 2;
        if (y == 0) return (above[x] + 3 * dcVal + 2) >> 2;
        if (x == 0) return (left[y] + 3 * dcVal + 2) >> 2;
        return dcVal;
    }
    return ang_pixel<pixel>(s, N, mode, bFilter, y, x, depth);
}

struct IntraArgs
{
    const void* nbr;          // neighbour arrays
    void* dst; int64_t dstStride;
    const x265b200_intra_job* jobs; int64_t n;
    int log2N, depth;
};

template<typename pixel>
__global__ void __launch_bounds__(256)
intra_pred_kernel(IntraArgs p)
{
    __shared__ pixel s[132];
    __shared__ int sDc;
    const x265b200_intra_job job = p.jobs[blockIdx.x];
    const int N = 1 << p.log2N;
    const pixel* src = (const pixel*)p.nbr + job.srcOff;
    for (int i = threadIdx.x; i < 4 * N + 1; i += blockDim.x) s[i] = src[i];
    __syncthreads();
    if (job.mode == 1)
    {
        if (threadIdx.x < 32)
        {
            int v = 0;
            for (int i = threadIdx.x; i < N; i += 32) v += s[1 + i] + s[2 * N + 1 + i];
            v = warp_sum(v);
            if (threadIdx.x == 0) sDc = (N + v) / (N + N);
        }
        __syncthreads();
    }
    const int dc = job.mode == 1 ? sDc : 0;
    pixel* d = (pixel*)p.dst + job.dstOff;
    for (int e = threadIdx.x; e < N * N; e += blockDim.x)
    {
        int y = e >> p.log2N, x = e & (N - 1);
        d[(int64_t)y * p.dstStride + x] = (pixel)intra_pixel<pixel>(s, N, p.log2N, job.mode, job.bFilter, y, x, p.depth, dc);
    }
}

// intraFilter<N> :31-51, n neighbour arrays of 4N+1 pixels at srcOff/dstOff = i*(4N+1) unless jobs given
template<typename pixel>
__global__ void __launch_bounds__(128)
intra_filter_kernel(const pixel* __restrict__ src, pixel* __restrict__ dst, int N, int64_t n)
{
    const int64_t b = blockIdx.x;
    const int len = 4 * N + 1, N2 = 2 * N;
    const pixel* s = src + b * len;
    pixel* f = dst + b * len;
    for (int i = threadIdx.x; i < len; i += blockDim.x)
    {
        int v;
        if (i == 0) v = ((s[0] << 1) + s[1] + s[N2 + 1] + 2) >> 2;
        else if (i == N2 || i == 2 * N2) v = s[i];
        else if (i == N2 + 1) v = ((s[N2 + 1] << 1) + s[0] + s[N2 + 2] + 2) >> 2;
        else v = ((s[i] << 1) + s[i - 1] + s[i + 1] + 2) >> 2;
        f[i] = (pixel)v;
    }
}

// all_angs_pred_c :206-234: block b -> dest[b][33][N*N]; refPix/filtPix are [n][4N+1].
// One CTA per block.  Phase 1 builds, for each of the 33 modes, the projected reference line ref[-N .. 2N] of
// intra_pred_ang_c (:146-172: flipped neighbours for horizontal modes, inverse-angle projection of the side samples for
// negative angles) in shared memory; phase 2 is then a 2-tap interpolation of CONSECUTIVE line entries, so a thread
// produces 4 horizontally adjacent pixels from 5 line bytes and issues one 32-bit (64-bit for 16-bit pixels) store.
// (The first version computed one pixel per thread through ang_pixel: 620 us per 274 MB size at 2160p, 15x off the
// write roofline; profiles/r01_launches_v5.csv.)
constexpr int ALLANGS_PITCH = 120;      // bytes/elements per line row: 3N + 2 = 98 entries + the 16-byte read-ahead of the word path
template<typename pixel>
__global__ void __launch_bounds__(256)
intra_allangs_kernel(const pixel* __restrict__ refPix, const pixel* __restrict__ filtPix, pixel* __restrict__ dest,
                     int log2N, int bLuma, int depth, int64_t n, int vec16, int all35)
{
    __shared__ pixel sr[132], sf[132];
    __shared__ __align__(16) pixel line[33][ALLANGS_PITCH];   // entry L <-> ref[L - N], L in [0, 3N + 1]
    const int N = 1 << log2N, N2 = N << 1, len = 4 * N + 1, LW = 3 * N + 2;
    const int64_t b = blockIdx.x;
    for (int i = threadIdx.x; i < len; i += blockDim.x) { sr[i] = refPix[b * len + i]; if (!all35) sf[i] = filtPix[b * len + i]; }
    __syncthreads();
    if (all35)
    {
        // all-35-modes form: the smoothed neighbours are produced here (intraFilter<N>, :31-51)
        for (int i = threadIdx.x; i < len; i += blockDim.x)
        {
            int v;
            if (i == 0) v = ((sr[0] << 1) + sr[1] + sr[N2 + 1] + 2) >> 2;
            else if (i == N2 || i == 2 * N2) v = sr[i];
            else if (i == N2 + 1) v = ((sr[N2 + 1] << 1) + sr[0] + sr[N2 + 2] + 2) >> 2;
            else v = ((sr[i] << 1) + sr[i - 1] + sr[i + 1] + 2) >> 2;
            sf[i] = (pixel)v;
        }
        __syncthreads();
    }
    // smoothing thresholds (constants.cpp:561): filtered when min(|m-26|,|m-10|) > {7,1,0} for N = 8,16,32; never for 4
    const int thr = N == 8 ? 7 : (N == 16 ? 1 : (N == 32 ? 0 : 99));
    for (int e = threadIdx.x; e < 33 * LW; e += blockDim.x)
    {
        const int m = e / LW, L = e - m * LW, idx = L - N, mode = m + 2;
        const bool hor = mode < 18;
        const int angleOffset = hor ? 10 - mode : mode - 26;
        const int angle = c_angle[8 + angleOffset];
        const pixel* s = min(abs(mode - 26), abs(mode - 10)) > thr ? sf : sr;
        int i;                                             // index into the (flipped) neighbour view
        if (angle >= 0 || idx >= -1) i = idx + 1;
        else i = N2 + ((128 + (-1 - idx) * c_invAngle[-angleOffset - 1]) >> 8);
        i = max(0, min(i, 4 * N));                         // entries outside a mode's reach are never read back
        const int j = (!hor || i == 0) ? i : (i <= N2 ? N2 + i : i - N2);
        line[m][L] = s[j];
    }
    __syncthreads();
    pixel* out = dest + (all35 ? b * 35 * N * N + 2 * N * N : b * 33 * N * N);
    const int maxVal = (1 << depth) - 1;
    if (all35)
    {
        // planar from the smoothed neighbours for N = 8/16/32, DC from the raw ones with the edge filter for N <= 16
        // (Search::estIntraPredQT, search.cpp:1358-1375)
        __shared__ int sDcAll;
        if (threadIdx.x < 32)
        {
            int v = 0;
            for (int i = threadIdx.x; i < N; i += 32) v += sr[1 + i] + sr[N2 + 1 + i];
            v = warp_sum(v);
            if (threadIdx.x == 0) sDcAll = (N + v) / (N + N);
        }
        __syncthreads();
        pixel* pd = dest + b * 35 * N * N;
        const pixel* pl = N >= 8 ? sf : sr;
        for (int e = threadIdx.x; e < 2 * N * N; e += blockDim.x)
        {
            const int mode = e >> (2 * log2N), r = e & (N * N - 1), y = r >> log2N, x = r & (N - 1);
            pd[e] = (pixel)intra_pixel<pixel>(mode ? sr : pl, N, log2N, mode, mode ? bLuma : 0, y, x, depth, sDcAll);
        }
    }
    if (sizeof(pixel) == 1 && vec16)
    {
        // 8-bit fast path: a thread produces 16 consecutive output bytes (one row segment for N >= 16, 2 rows of 8, 4 rows
        // of 4) from aligned word loads of the line, two pixels per 32-bit multiply-add (16-bit lanes: 255*32 + 16 < 2^16),
        // and issues one 128-bit store
        const uint32_t* lw = (const uint32_t*)&line[0][0];
        const int CH = N < 16 ? N : 16, nw = CH >> 2;
        for (int q = threadIdx.x; q < 33 * N * N / 16; q += blockDim.x)
        {
            const int e = q * 16;
            const int m = e >> (2 * log2N), r = e & (N * N - 1), mode = m + 2;
            const bool hor = mode < 18;
            const int angle = c_angle[8 + (hor ? 10 - mode : mode - 26)];
            uint32_t ow[4];
            for (int c = 0; c < 16 / CH; c++)
            {
                const int rr = r + c * CH, y = rr >> log2N, x = rr & (N - 1);
                const int angleSum = (y + 1) * angle, offset = angleSum >> 5;
                const uint32_t f = (uint32_t)(angleSum & 31), g = 32u - f;
                const int base = offset + x + N, sh = (base & 3) * 8;
                const uint32_t* src = lw + m * (ALLANGS_PITCH / 4) + (base >> 2);
                uint32_t w[5];
#pragma unroll
                for (int k = 0; k < 5; k++) w[k] = k <= nw ? src[k] : 0u;
#pragma unroll
                for (int k = 0; k < 4; k++)
                {
                    if (k >= nw) break;
                    const uint32_t A = __funnelshift_r(w[k], w[k + 1], sh), B = __funnelshift_rc(w[k], w[k + 1], sh + 8);
                    const uint32_t r02 = (((A & 0x00FF00FFu) * g + (B & 0x00FF00FFu) * f + 0x00100010u) >> 5) & 0x00FF00FFu;
                    const uint32_t r13 = ((((A >> 8) & 0x00FF00FFu) * g + ((B >> 8) & 0x00FF00FFu) * f + 0x00100010u) >> 5) & 0x00FF00FFu;
                    ow[c * nw + k] = r02 | (r13 << 8);
                }
                if (!angle && bLuma && x == 0)
                {
                    // pure horizontal / vertical, first column edge-filtered for luma (:176-189)
                    const pixel* s = min(abs(mode - 26), abs(mode - 10)) > thr ? sf : sr;
                    auto nb = [&](int i) -> int { return (!hor || i == 0) ? (int)s[i] : (i <= N2 ? (int)s[N2 + i] : (int)s[i - N2]); };
                    const int v = (int16_t)(nb(1) + ((nb(N2 + 1 + y) - nb(0)) >> 1));
                    ow[c * nw] = (ow[c * nw] & 0xFFFFFF00u) | (uint32_t)(v < 0 ? 0 : (v > maxVal ? maxVal : v));
                }
            }
            *(uint4*)(out + e) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        }
        return;
    }
    for (int q = threadIdx.x; q < 33 * N * N / 4; q += blockDim.x)
    {
        const int e = q * 4;
        const int m = e >> (2 * log2N), r = e & (N * N - 1), mode = m + 2;
        const int y = r >> log2N, x = r & (N - 1);
        const bool hor = mode < 18;
        const int angle = c_angle[8 + (hor ? 10 - mode : mode - 26)];
        int o[4];
        if (!angle)
        {
            // pure horizontal / vertical: copy of ref[1 + x], first column edge-filtered for luma (:176-189)
            const pixel* a = &line[m][N + x];
#pragma unroll
            for (int k = 0; k < 4; k++) o[k] = a[k];
            if (bLuma && x == 0)
            {
                const pixel* s = min(abs(mode - 26), abs(mode - 10)) > thr ? sf : sr;
                auto nb = [&](int i) -> int { return (!hor || i == 0) ? (int)s[i] : (i <= N2 ? (int)s[N2 + i] : (int)s[i - N2]); };
                const int v = (int16_t)(nb(1) + ((nb(N2 + 1 + y) - nb(0)) >> 1));
                o[0] = v < 0 ? 0 : (v > maxVal ? maxVal : v);
            }
        }
        else
        {
            const int angleSum = (y + 1) * angle, offset = angleSum >> 5, fraction = angleSum & 31;
            const pixel* a = &line[m][offset + x + N];
            int pv[5];
#pragma unroll
            for (int k = 0; k < 5; k++) pv[k] = a[k];
#pragma unroll
            for (int k = 0; k < 4; k++) o[k] = fraction ? ((32 - fraction) * pv[k] + fraction * pv[k + 1] + 16) >> 5 : pv[k];
        }
        if (sizeof(pixel) == 1)
            *(uint32_t*)(out + e) = (uint32_t)o[0] | ((uint32_t)o[1] << 8) | ((uint32_t)o[2] << 16) | ((uint32_t)o[3] << 24);
        else
            *(uint2*)(out + e) = make_uint2((uint32_t)o[0] | ((uint32_t)o[1] << 16), (uint32_t)o[2] | ((uint32_t)o[3] << 16));
    }
}

// 8-bit all-angles, persistent form (the default for 16-byte aligned 8-bit output).  The first two versions launched one
// CTA per block: at 2160p that is 130 560 CTAs of 8x8 blocks whose fixed cost (neighbour fetch -> barrier -> line build with
// per-entry divisions and divergent __constant__ lookups -> barrier) dominated: 450 / 251 / 185 us for the 8 / 16 / 32 sizes
// against a 45 us write floor (profiles/r01_launches_v4.csv).  Here a CTA loops over groups of G blocks: the (mode, L) ->
// neighbour-index table is built ONCE per CTA, the next group's neighbour bytes are prefetched into registers while the
// current group is predicted, a thread emits 16 output bytes per 128-bit store from aligned word loads of the line with
// two pixels per 32-bit multiply-add (16-bit lanes: 255*32 + 16 < 2^16).
// ALL35: the fused mode-search form (x265b200_intra_modes_dev): only the raw neighbours are read, the smoothed ones are
// produced in shared memory, and planar + DC are emitted in front of the 33 angular blocks (35 x N x N per block).
template<int LOG2N, bool ALL35>
__global__ void __launch_bounds__(256)
intra_allangs8_kernel(const uint8_t* __restrict__ refPix, const uint8_t* __restrict__ filtPix, uint8_t* __restrict__ dest, int bLuma, int64_t n)
{
    constexpr int N = 1 << LOG2N, N2 = N << 1, LEN = 4 * N + 1, LW = 3 * N + 2;
    constexpr int G = N == 32 ? 1 : (N == 16 ? 2 : (N == 8 ? 8 : 16));          // blocks per group (work between two barriers)
    constexpr int PITCH = (LW + 16 + 7) & ~7;                  // line row: LW entries + the read-ahead of the word path
    constexpr int SP = 136;                                     // neighbour array pitch (>= LEN)
    constexpr int NM = ALL35 ? 35 : 33;                         // prediction blocks per input block
    constexpr int CPB = NM * N * N / 16;                        // 16-byte chunks per block
    constexpr int CH = N < 16 ? N : 16, NWD = CH / 4;
    constexpr int TOT = G * (ALL35 ? 1 : 2) * LEN;              // neighbour bytes staged per group
    constexpr int NU = (TOT + 255) / 256;                       // ... = NU bytes per thread
    __shared__ int sDc[G];
    __shared__ __align__(8) uint16_t jtab[33 * PITCH];          // (mode, L) -> neighbour index, rows padded to the line pitch
    __shared__ uint8_t s2[G][2 * SP];                           // per block: [unfiltered | filtered] neighbours
    __shared__ __align__(16) uint8_t line[G][33][PITCH];        // entry L <-> ref[L - N]
    const int tid = threadIdx.x;
    const int thr = N == 8 ? 7 : (N == 16 ? 1 : (N == 32 ? 0 : 99));      // constants.cpp:561 g_intraFilterFlags

    for (int e = tid; e < 33 * PITCH; e += 256)
    {
        const int m = e / PITCH, L = e - m * PITCH, idx = L - N, mode = m + 2;
        const bool hor = mode < 18;
        const int angleOffset = hor ? 10 - mode : mode - 26;
        const int angle = c_angle[8 + angleOffset];
        int i;                                                  // index into the (flipped) neighbour view, intrapred.cpp:146-172
        if (angle >= 0 || idx >= -1) i = idx + 1;
        else i = N2 + ((128 + (-1 - idx) * c_invAngle[-angleOffset - 1]) >> 8);
        i = max(0, min(i, 4 * N));                              // entries outside a mode's reach (and the row padding) are never used
        const int j = (!hor || i == 0) ? i : (i <= N2 ? N2 + i : i - N2);
        jtab[e] = (uint16_t)(j + (min(abs(mode - 26), abs(mode - 10)) > thr ? SP : 0));
    }

    // staging map of this thread's (up to two) neighbour bytes
    int sOff[NU], gOf[NU]; int64_t srcOff[NU]; bool sFilt[NU], sOk[NU];
#pragma unroll
    for (int u = 0; u < NU; u++)
    {
        const int t = tid + u * 256;
        sOk[u] = t < TOT;
        const int g = t / ((ALL35 ? 1 : 2) * LEN), rem = t - g * ((ALL35 ? 1 : 2) * LEN), which = rem >= LEN, i = rem - which * LEN;
        gOf[u] = g; sFilt[u] = which; sOff[u] = g * (2 * SP) + which * SP + i; srcOff[u] = (int64_t)g * LEN + i;
    }
    const int64_t stride = (int64_t)gridDim.x * G;
    int64_t b0 = (int64_t)blockIdx.x * G;
    uint8_t v[NU];
#pragma unroll
    for (int u = 0; u < NU; u++)
    {
        v[u] = 0;
        if (sOk[u] && b0 + gOf[u] < n) v[u] = (sFilt[u] ? filtPix : refPix)[b0 * LEN + srcOff[u]];
    }

    const uint32_t* lw = (const uint32_t*)&line[0][0][0];
    for (; b0 < n; b0 += stride)
    {
#pragma unroll
        for (int u = 0; u < NU; u++) if (sOk[u]) (&s2[0][0])[sOff[u]] = v[u];
        __syncthreads();
        if (ALL35)
        {
            // intraFilter<N> (intrapred.cpp:31-51) into the second half of s2, and the DC value of each block
            for (int e = tid; e < G * LEN; e += 256)
            {
                const int g = e / LEN, i = e - g * LEN;
                const uint8_t* sr = s2[g];
                int fv;
                if (i == 0) fv = ((sr[0] << 1) + sr[1] + sr[N2 + 1] + 2) >> 2;
                else if (i == N2 || i == 2 * N2) fv = sr[i];
                else if (i == N2 + 1) fv = ((sr[N2 + 1] << 1) + sr[0] + sr[N2 + 2] + 2) >> 2;
                else fv = ((sr[i] << 1) + sr[i - 1] + sr[i + 1] + 2) >> 2;
                s2[g][SP + i] = (uint8_t)fv;
            }
            // DC values (intrapred.cpp:72-78): warp w sums the 2N neighbours of blocks w, w + 8, ...
            for (int g = tid >> 5; g < G; g += 8)
            {
                const int lane = tid & 31;
                int sum = lane < N ? (int)s2[g][1 + lane] + (int)s2[g][N2 + 1 + lane] : 0;
                sum = warp_sum(sum);
                if (lane == 0) sDc[g] = (N + sum) / (N + N);
            }
            __syncthreads();
        }
        const int64_t nb0 = b0 + stride;
#pragma unroll
        for (int u = 0; u < NU; u++)
            if (sOk[u] && nb0 + gOf[u] < n) v[u] = (sFilt[u] ? filtPix : refPix)[nb0 * LEN + srcOff[u]];

        // the 33 projected reference lines of every block, four entries (one word) per thread
        for (int e = tid; e < G * 33 * (PITCH / 4); e += 256)
        {
            const int g = e / (33 * (PITCH / 4)), r = e - g * (33 * (PITCH / 4));
            const uint2 jt = *(const uint2*)&jtab[r * 4];
            const uint8_t* sn = s2[g];
            const uint32_t word = (uint32_t)sn[jt.x & 0xffff] | ((uint32_t)sn[jt.x >> 16] << 8) | ((uint32_t)sn[jt.y & 0xffff] << 16) | ((uint32_t)sn[jt.y >> 16] << 24);
            ((uint32_t*)&line[g][0][0])[r] = word;
        }
        __syncthreads();

        for (int q = tid; q < G * CPB; q += 256)
        {
            const int g = q / CPB, e = (q - g * CPB) * 16;
            if (b0 + g >= n) continue;
            if (ALL35 && e < 2 * N * N)
            {
                // planar (smoothed neighbours for N >= 8) / DC (raw neighbours, edge filter when bLuma): search.cpp:1358-1375
                const int pm = e >> (2 * LOG2N), r0 = e & (N * N - 1);
                const uint8_t* sn = &s2[g][(pm == 0 && N >= 8) ? SP : 0];
                uint32_t ow[4];
#pragma unroll
                for (int wq = 0; wq < 4; wq++)
                {
                    uint32_t word = 0;
#pragma unroll
                    for (int c = 0; c < 4; c++)
                    {
                        const int rr = r0 + wq * 4 + c, y = rr >> LOG2N, x = rr & (N - 1);
                        word |= (uint32_t)intra_pixel<uint8_t>(sn, N, LOG2N, pm, pm ? bLuma : 0, y, x, 8, sDc[g]) << (8 * c);
                    }
                    ow[wq] = word;
                }
                *(uint4*)(dest + (b0 + g) * (int64_t)(NM * N * N) + e) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                continue;
            }
            const int eAng = ALL35 ? e - 2 * N * N : e;
            const int m = eAng >> (2 * LOG2N), r = eAng & (N * N - 1), mode = m + 2;
            const bool hor = mode < 18;
            const int angle = c_angle[8 + (hor ? 10 - mode : mode - 26)];
            uint32_t ow[4];
#pragma unroll
            for (int c = 0; c < 16 / CH; c++)
            {
                const int rr = r + c * CH, y = rr >> LOG2N, x = rr & (N - 1);
                const int angleSum = (y + 1) * angle, offset = angleSum >> 5;
                const uint32_t f = (uint32_t)(angleSum & 31), gg = 32u - f;
                const int base = offset + x + N, sh = (base & 3) * 8;
                const uint32_t* src = lw + (g * 33 + m) * (PITCH / 4) + (base >> 2);
                uint32_t w[NWD + 1];
#pragma unroll
                for (int k = 0; k <= NWD; k++) w[k] = src[k];
                // source bytes s0..s4k+4 of the line; E = (s0, s2), O = (s1, s3), X = (s2, s4) as 16-bit lanes per word:
                // even outputs g*E + f*O, odd outputs g*O + f*X (255*32 + 16 < 2^16: no carry between lanes), PRMT repacks
                uint32_t E[NWD + 1], O[NWD];
#pragma unroll
                for (int k = 0; k < NWD; k++)
                {
                    const uint32_t A = __funnelshift_r(w[k], w[k + 1], sh);
                    E[k] = __byte_perm(A, 0u, 0x4240); O[k] = __byte_perm(A, 0u, 0x4341);
                }
                E[NWD] = __byte_perm(w[NWD] >> sh, 0u, 0x4240);
#pragma unroll
                for (int k = 0; k < NWD; k++)
                {
                    const uint32_t X = __byte_perm(E[k], E[k + 1], 0x5432);
                    const uint32_t ev = (E[k] * gg + (O[k] * f + 0x00100010u)) >> 5;
                    const uint32_t od = (O[k] * gg + (X * f + 0x00100010u)) >> 5;
                    ow[c * NWD + k] = __byte_perm(ev, od, 0x6240);
                }
                if (!angle && bLuma && x == 0)
                {
                    // pure horizontal / vertical: first column edge-filtered for luma (intrapred.cpp:176-189)
                    const uint8_t* s = &s2[g][min(abs(mode - 26), abs(mode - 10)) > thr ? SP : 0];
                    auto nb = [&](int i) -> int { return (!hor || i == 0) ? (int)s[i] : (i <= N2 ? (int)s[N2 + i] : (int)s[i - N2]); };
                    const int vv = (int16_t)(nb(1) + ((nb(N2 + 1 + y) - nb(0)) >> 1));
                    ow[c * NWD] = (ow[c * NWD] & 0xFFFFFF00u) | (uint32_t)(vv < 0 ? 0 : (vv > 255 ? 255 : vv));
                }
            }
            *(uint4*)(dest + (b0 + g) * (int64_t)(NM * N * N) + e) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        }
        __syncthreads();
    }
}

int intra_pred_dev(Ctx* ctx, int depth, int log2N, const void* nbr, void* dst, int64_t dstStride,
                   const x265b200_intra_job* jobs, int64_t n)
{
    if (n <= 0) return 0;
    if (log2N < 2 || log2N > 5) { set_error("intra: log2N %d", log2N); return -1; }
    IntraArgs a; a.nbr = nbr; a.dst = dst; a.dstStride = dstStride; a.jobs = jobs; a.n = n; a.log2N = log2N; a.depth = depth;
    int threads = (1 << (2 * log2N)) < 256 ? max(32, 1 << (2 * log2N)) : 256;
    if (depth > 8) intra_pred_kernel<uint16_t><<<(unsigned)n, threads, 0, ctx->stream>>>(a);
    else           intra_pred_kernel<uint8_t><<<(unsigned)n, threads, 0, ctx->stream>>>(a);
    ctx->launches++;
    return check(cudaGetLastError(), "intra_pred kernel launch");
}

int intra_filter_dev(Ctx* ctx, int depth, int log2N, const void* src, void* dst, int64_t n)
{
    if (n <= 0) return 0;
    if (log2N < 2 || log2N > 5) { set_error("intra_filter: log2N %d", log2N); return -1; }
    if (depth > 8) intra_filter_kernel<uint16_t><<<(unsigned)n, 128, 0, ctx->stream>>>((const uint16_t*)src, (uint16_t*)dst, 1 << log2N, n);
    else           intra_filter_kernel<uint8_t><<<(unsigned)n, 128, 0, ctx->stream>>>((const uint8_t*)src, (uint8_t*)dst, 1 << log2N, n);
    ctx->launches++;
    return check(cudaGetLastError(), "intra_filter kernel launch");
}

int scratch_dev(Ctx* ctx, int slot, size_t bytes, void** out);

static int intra_allangs_launch(Ctx* ctx, int depth, int log2N, const void* refPix, const void* filtPix, void* dest, int bLuma, int64_t n, int all35)
{
    if (n <= 0) return 0;
    if (log2N < 2 || log2N > 5) { set_error("intra_allangs: log2N %d", log2N); return -1; }
    if ((uintptr_t)dest & (depth > 8 ? 7 : 3)) { set_error("intra_allangs: dest must be %d-byte aligned (4 pixels per store)", depth > 8 ? 8 : 4); return -1; }
    if (depth > 8) intra_allangs_kernel<uint16_t><<<(unsigned)n, 256, 0, ctx->stream>>>((const uint16_t*)refPix, (const uint16_t*)filtPix, (uint16_t*)dest, log2N, bLuma, depth, n, 0, all35);
    else if (((uintptr_t)dest & 15) == 0)
    {
        const int G = log2N == 5 ? 1 : (log2N == 4 ? 2 : (log2N == 3 ? 8 : 16));
        const int64_t groups = (n + G - 1) / G, cap = (int64_t)ctx->smCount * 8;
        const unsigned grid = (unsigned)(groups < cap ? groups : cap);
        const uint8_t* r = (const uint8_t*)refPix; const uint8_t* f = (const uint8_t*)filtPix; uint8_t* d = (uint8_t*)dest;
#define AA_LAUNCH(L) do { if (all35) intra_allangs8_kernel<L, true><<<grid, 256, 0, ctx->stream>>>(r, f, d, bLuma, n); \
                          else intra_allangs8_kernel<L, false><<<grid, 256, 0, ctx->stream>>>(r, f, d, bLuma, n); } while (0)
        switch (log2N)
        {
        case 2: AA_LAUNCH(2); break;
        case 3: AA_LAUNCH(3); break;
        case 4: AA_LAUNCH(4); break;
        default: AA_LAUNCH(5); break;
        }
#undef AA_LAUNCH
    }
    else           intra_allangs_kernel<uint8_t><<<(unsigned)n, 256, 0, ctx->stream>>>((const uint8_t*)refPix, (const uint8_t*)filtPix, (uint8_t*)dest, log2N, bLuma, depth, n, 0, all35);
    ctx->launches++;
    return check(cudaGetLastError(), "intra_allangs kernel launch");
}

int intra_allangs_dev(Ctx* ctx, int depth, int log2N, const void* refPix, const void* filtPix, void* dest, int bLuma, int64_t n)
{
    return intra_allangs_launch(ctx, depth, log2N, refPix, filtPix, dest, bLuma, n, 0);
}
// the prediction half of the intra mode search (Search::estIntraPredQT, search.cpp:1358-1400, as the asm table drives it):
// intra_filter + DC + planar + intra_pred_allangs of every block in one launch, 35 x N x N predictions per block
int intra_modes_dev(Ctx* ctx, int depth, int log2N, const void* neighbours, void* dest, int bLuma, int64_t n)
{
    return intra_allangs_launch(ctx, depth, log2N, neighbours, nullptr, dest, bLuma, n, 1);
}

} // namespace x265b200

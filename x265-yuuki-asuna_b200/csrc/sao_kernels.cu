// sao_kernels.cu -- the in-loop filter entries of the table: SAO offset application
// (saoCuOrgE0/E1/E1_2Rows/E2/E3/B0, source/common/loopfilter.cpp:45-139), SAO statistics (saoCuStatsBO/E0..E3,
// source/encoder/sao.cpp:1762-1926), primitives.sign (loopfilter.cpp:39-43) and the deblocking line filters
// pelFilterLumaStrong / pelFilterChroma (loopfilter.cpp:141-180).
//
// The C code walks pixels in raster order carrying "sign of the previous comparison" state (signLeft, upBuff1, upBufft).
// Every one of those states is the sign of a difference of two ORIGINAL pixels -- the in-place updates never feed a later
// comparison (each row is compared with the row below / the pixel to the right before either is rewritten) -- so the edge
// class of a pixel is a closed form of the input picture plus, for the first row / column only, the caller's buffers.
// That makes every entry a one-pass, fully parallel kernel: one CTA per job (CTU-sized block), one thread per pixel.
#include "common.cuh"
#include "x265b200.h"

namespace x265b200 {

namespace {

__device__ __forceinline__ int sgn(int x) { return (x > 0) - (x < 0); }                   // signOf, loopfilter.cpp:33-36
__constant__ int c_eoTable[5] = { 1, 2, 0, 3, 4 };                                        // SAO::s_eoTable, sao.cpp:65-72

struct SaoApplyArgs
{
    void* rec; int64_t stride;
    const x265b200_sao_job* jobs; int64_t n;
    int8_t* buf;                 // sign buffers (upBuff1 / bufft / signLeft), addressed by job.buf0 / job.buf1
    const int8_t* offsets;       // offsetEo[5] / offsetBo[32] tables, addressed by job.offsetOff
    int kind, depth;
};

// one CTA per job; E0..E3: blockDim.x >= width (one thread per column, every buffer read happens before the barrier that
// precedes the buffer writes); B0: threads stride over the block
template<typename pixel>
__global__ void __launch_bounds__(256)
sao_apply_kernel(SaoApplyArgs a)
{
    const x265b200_sao_job j = a.jobs[blockIdx.x];
    pixel* rec = (pixel*)a.rec + j.recOff;
    const int8_t* off = a.offsets + j.offsetOff;
    const int maxVal = (1 << a.depth) - 1;
    const int64_t s = a.stride;
    const int x = threadIdx.x;
    auto clipAdd = [&](int v, int o) -> pixel { v += o; return (pixel)(v < 0 ? 0 : (v > maxVal ? maxVal : v)); };
    switch (a.kind)
    {
    case X265B200_SAO_E0:            // processSaoCUE0: 2 rows, left/right neighbours, signLeft[2] = buf0
    {
        int e[2] = { 0, 0 };
        if (x < j.width)
            for (int y = 0; y < 2; y++)
            {
                const pixel* r = rec + y * s;
                const int signRight = sgn((int)r[x] - (int)r[x + 1]);
                const int signLeft = x ? sgn((int)r[x] - (int)r[x - 1]) : (int)a.buf[j.buf0 + y];
                e[y] = signRight + signLeft + 2;
            }
        __syncthreads();
        if (x < j.width)
            for (int y = 0; y < 2; y++) rec[y * s + x] = clipAdd(rec[y * s + x], off[e[y]]);
        break;
    }
    case X265B200_SAO_E1:            // processSaoCUE1 (rows = 1) / processSaoCUE1_2Rows (rows = 2), upBuff1 = buf0
    case X265B200_SAO_E1_2ROWS:
    {
        const int rows = a.kind == X265B200_SAO_E1 ? 1 : 2;
        int e[2] = { 0, 0 }, last = 0;
        if (x < j.width)
            for (int y = 0; y < rows; y++)
            {
                const int signDown = sgn((int)rec[y * s + x] - (int)rec[(y + 1) * s + x]);
                const int up = y ? sgn((int)rec[y * s + x] - (int)rec[(y - 1) * s + x]) : (int)a.buf[j.buf0 + x];
                e[y] = signDown + up + 2; last = -signDown;
            }
        __syncthreads();
        if (x < j.width)
        {
            for (int y = 0; y < rows; y++) rec[y * s + x] = clipAdd(rec[y * s + x], off[e[y]]);
            a.buf[j.buf0 + x] = (int8_t)last;
        }
        break;
    }
    case X265B200_SAO_E2:            // processSaoCUE2: bufft = buf0 (written at x + 1), buff1 = buf1
    {
        int e = 0, sd = 0;
        if (x < j.width) { sd = sgn((int)rec[x] - (int)rec[x + s + 1]); e = sd + (int)a.buf[j.buf1 + x] + 2; }
        __syncthreads();
        if (x < j.width) { a.buf[j.buf0 + x + 1] = (int8_t)(-sd); rec[x] = clipAdd(rec[x], off[e]); }
        break;
    }
    case X265B200_SAO_E3:            // processSaoCUE3: columns (startX, endX), upBuff1 = buf0 read at x, written at x - 1
    {
        int e = 0, sd = 0;
        const bool on = x > j.startX && x < j.width;            // width carries endX
        if (on) { sd = sgn((int)rec[x] - (int)rec[x + s]); e = sd + (int)a.buf[j.buf0 + x] + 2; }
        __syncthreads();
        if (on) { a.buf[j.buf0 + x - 1] = (int8_t)(-sd); rec[x] = clipAdd(rec[x], off[e]); }
        break;
    }
    default:                         // processSaoCUB0: band offset over width x height
    {
        const int boShift = a.depth - 5;
        for (int e = threadIdx.x; e < j.width * j.height; e += blockDim.x)
        {
            const int y = e / j.width, xx = e - y * j.width;
            const int v = rec[y * s + xx];
            rec[y * s + xx] = clipAdd(v, off[v >> boShift]);
        }
        break;
    }
    }
}

struct SaoStatsArgs
{
    const int16_t* diff;         // fenc - rec, row pitch 64 (MAX_CU_SIZE)
    const void* rec; int64_t stride;
    const x265b200_sao_job* jobs; int64_t n;
    int8_t* buf; int32_t* stats; int32_t* count;
    int kind, depth;
};

// saoCuStatsBO/E0/E1/E2/E3_c: class histogram of diff over endX x endY pixels, added to stats/count at job.offsetOff
// (E kinds: 5 classes mapped through s_eoTable; BO: 32 bands).  job.width/height = endX/endY.
template<typename pixel>
__global__ void __launch_bounds__(256)
sao_stats_kernel(SaoStatsArgs a)
{
    __shared__ int sStat[32], sCnt[32];
    const x265b200_sao_job j = a.jobs[blockIdx.x];
    const pixel* rec = (const pixel*)a.rec + j.recOff;
    const int16_t* diff = a.diff + j.diffOff;
    const int64_t s = a.stride;
    const int endX = j.width, endY = j.height;
    if (threadIdx.x < 32) { sStat[threadIdx.x] = 0; sCnt[threadIdx.x] = 0; }
    __syncthreads();
    const int boShift = a.depth - 5;
    for (int e = threadIdx.x; e < endX * endY; e += blockDim.x)
    {
        const int y = e / endX, x = e - y * endX;
        const pixel* r = rec + y * s;
        const int c = r[x];
        int cls;
        switch (a.kind)
        {
        case X265B200_SAO_BO: cls = c >> boShift; break;
        case X265B200_SAO_E0: cls = sgn(c - (int)r[x + 1]) + sgn(c - (int)r[x - 1]) + 2; break;                       // signLeft(x) = sign(rec[x] - rec[x-1])
        case X265B200_SAO_E1: cls = sgn(c - (int)r[x + s]) + (y ? sgn(c - (int)r[x - s]) : (int)a.buf[j.buf0 + x]) + 2; break;
        case X265B200_SAO_E2: cls = sgn(c - (int)r[x + s + 1]) + (y ? sgn(c - (int)r[x - s - 1]) : (int)a.buf[j.buf0 + x]) + 2; break;
        default:              cls = sgn(c - (int)r[x + s - 1]) + (y ? sgn(c - (int)r[x - s + 1]) : (int)a.buf[j.buf0 + x]) + 2; break;   // E3
        }
        atomicAdd(&sStat[cls], (int)diff[y * 64 + x]);
        atomicAdd(&sCnt[cls], 1);
    }
    __syncthreads();
    const int ncls = a.kind == X265B200_SAO_BO ? 32 : 5;
    if (threadIdx.x < ncls)
    {
        const int dst = a.kind == X265B200_SAO_BO ? threadIdx.x : c_eoTable[threadIdx.x];
        atomicAdd(&a.stats[j.offsetOff + dst], sStat[threadIdx.x]);
        atomicAdd(&a.count[j.offsetOff + dst], sCnt[threadIdx.x]);
    }
    // final state of the caller's sign buffers (what the row-by-row C loops leave behind)
    if (endY <= 0 || endX <= 0) return;
    if (a.kind == X265B200_SAO_E1)
    {
        for (int x = threadIdx.x; x < endX; x += blockDim.x)
            a.buf[j.buf0 + x] = (int8_t)sgn((int)rec[endY * s + x] - (int)rec[(endY - 1) * s + x]);
    }
    else if (a.kind == X265B200_SAO_E2)
    {
        // row y writes W_y[i] = sign(rec[y+1][i] - rec[y][i-1]), i = 0..endX, into upBufft and the two pointers are swapped:
        // even rows land in the caller's upBufft (buf1), odd rows in the caller's upBuff1 (buf0)
        const int lastEven = (endY - 1) & ~1, lastOdd = ((endY - 2) & ~1) + 1;
        for (int i = threadIdx.x; i <= endX; i += blockDim.x)
        {
            a.buf[j.buf1 + i] = (int8_t)sgn((int)rec[(lastEven + 1) * s + i] - (int)rec[lastEven * s + i - 1]);
            if (endY >= 2) a.buf[j.buf0 + i] = (int8_t)sgn((int)rec[(lastOdd + 1) * s + i] - (int)rec[lastOdd * s + i - 1]);
        }
    }
    else if (a.kind == X265B200_SAO_E3)
    {
        const int y = endY - 1;
        for (int i = (int)threadIdx.x - 1; i <= endX - 1; i += blockDim.x)
            a.buf[j.buf0 + i] = (int8_t)sgn((int)rec[(y + 1) * s + i] - (int)rec[y * s + i + 1]);
    }
}

template<typename pixel>
__global__ void sign_kernel(int8_t* dst, const pixel* s1, const pixel* s2, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (int8_t)sgn((int)s1[i] - (int)s2[i]);
}

// pelFilterLumaStrong_c / pelFilterChroma_c: 4 lines (UNIT_SIZE) per job, one thread per line
template<typename pixel>
__global__ void deblock_kernel(pixel* pic, const x265b200_deblock_job* jobs, int64_t n, int chroma, int depth)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * 4) return;
    const x265b200_deblock_job j = jobs[t >> 2];
    pixel* src = pic + j.srcOff + (t & 3) * j.srcStep;
    const int64_t o = j.offset;
    const int maxVal = (1 << depth) - 1;
    const int m4 = (int16_t)src[0], m3 = (int16_t)src[-o], m5 = (int16_t)src[o], m2 = (int16_t)src[-o * 2];
    if (chroma)
    {
        const int tc = j.tcP, maskP = j.tcQ, maskQ = j.maskQ;
        const int delta = min(max(-tc, (((m4 - m3) * 4) + m2 - m5 + 4) >> 3), tc);
        const int a = m3 + (delta & maskP), b = m4 - (delta & maskQ);
        src[-o] = (pixel)(a < 0 ? 0 : (a > maxVal ? maxVal : a));
        src[0] = (pixel)(b < 0 ? 0 : (b > maxVal ? maxVal : b));
        return;
    }
    const int m6 = (int16_t)src[o * 2], m1 = (int16_t)src[-o * 3], m7 = (int16_t)src[o * 3], m0 = (int16_t)src[-o * 4];
    const int tcP = j.tcP, tcQ = j.tcQ;
    auto c3 = [](int lo, int hi, int v) { return min(max(lo, v), hi); };                 // x265_clip3
    src[-o * 3] = (pixel)(c3(-tcP, tcP, ((2 * m0 + 3 * m1 + m2 + m3 + m4 + 4) >> 3) - m1) + m1);
    src[-o * 2] = (pixel)(c3(-tcP, tcP, ((m1 + m2 + m3 + m4 + 2) >> 2) - m2) + m2);
    src[-o]     = (pixel)(c3(-tcP, tcP, ((m1 + 2 * m2 + 2 * m3 + 2 * m4 + m5 + 4) >> 3) - m3) + m3);
    src[0]      = (pixel)(c3(-tcQ, tcQ, ((m2 + 2 * m3 + 2 * m4 + 2 * m5 + m6 + 4) >> 3) - m4) + m4);
    src[o]      = (pixel)(c3(-tcQ, tcQ, ((m3 + m4 + m5 + m6 + 2) >> 2) - m5) + m5);
    src[o * 2]  = (pixel)(c3(-tcQ, tcQ, ((m3 + m4 + m5 + 3 * m6 + 2 * m7 + 4) >> 3) - m6) + m6);
}

} // namespace

int sao_apply_dev(Ctx* ctx, int kind, int depth, void* rec, int64_t stride, const x265b200_sao_job* jobs, int64_t n,
                  int8_t* signBuf, const int8_t* offsets, int maxWidth)
{
    if (n <= 0) return 0;
    if (kind < X265B200_SAO_E0 || kind > X265B200_SAO_B0) { set_error("sao_apply: kind %d", kind); return -1; }
    if (kind != X265B200_SAO_B0 && (maxWidth < 1 || maxWidth > 256)) { set_error("sao_apply: width %d (1..256)", maxWidth); return -1; }
    SaoApplyArgs a; a.rec = rec; a.stride = stride; a.jobs = jobs; a.n = n; a.buf = signBuf; a.offsets = offsets; a.kind = kind; a.depth = depth;
    const int threads = kind == X265B200_SAO_B0 ? 256 : ((maxWidth + 31) & ~31);
    if (depth > 8) sao_apply_kernel<uint16_t><<<(unsigned)n, threads, 0, ctx->stream>>>(a);
    else           sao_apply_kernel<uint8_t><<<(unsigned)n, threads, 0, ctx->stream>>>(a);
    ctx->launches++;
    return check(cudaGetLastError(), "sao_apply kernel launch");
}

int sao_stats_dev(Ctx* ctx, int kind, int depth, const int16_t* diff, const void* rec, int64_t stride, const x265b200_sao_job* jobs, int64_t n,
                  int8_t* signBuf, int32_t* stats, int32_t* count)
{
    if (n <= 0) return 0;
    if (kind != X265B200_SAO_BO && (kind < X265B200_SAO_E0 || kind > X265B200_SAO_E3 || kind == X265B200_SAO_E1_2ROWS)) { set_error("sao_stats: kind %d", kind); return -1; }
    SaoStatsArgs a; a.diff = diff; a.rec = rec; a.stride = stride; a.jobs = jobs; a.n = n; a.buf = signBuf; a.stats = stats; a.count = count;
    a.kind = kind; a.depth = depth;
    if (depth > 8) sao_stats_kernel<uint16_t><<<(unsigned)n, 256, 0, ctx->stream>>>(a);
    else           sao_stats_kernel<uint8_t><<<(unsigned)n, 256, 0, ctx->stream>>>(a);
    ctx->launches++;
    return check(cudaGetLastError(), "sao_stats kernel launch");
}

int sign_dev(Ctx* ctx, int depth, int8_t* dst, const void* src1, const void* src2, int64_t n)
{
    if (n <= 0) return 0;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (depth > 8) sign_kernel<uint16_t><<<blocks, 256, 0, ctx->stream>>>(dst, (const uint16_t*)src1, (const uint16_t*)src2, n);
    else           sign_kernel<uint8_t><<<blocks, 256, 0, ctx->stream>>>(dst, (const uint8_t*)src1, (const uint8_t*)src2, n);
    ctx->launches++;
    return check(cudaGetLastError(), "sign kernel launch");
}

int deblock_dev(Ctx* ctx, int chroma, int depth, void* pic, const x265b200_deblock_job* jobs, int64_t n)
{
    if (n <= 0) return 0;
    const unsigned blocks = (unsigned)((n * 4 + 127) / 128);
    if (depth > 8) deblock_kernel<uint16_t><<<blocks, 128, 0, ctx->stream>>>((uint16_t*)pic, jobs, n, chroma, depth);
    else           deblock_kernel<uint8_t><<<blocks, 128, 0, ctx->stream>>>((uint8_t*)pic, jobs, n, chroma, depth);
    ctx->launches++;
    return check(cudaGetLastError(), "deblock kernel launch");
}

} // namespace x265b200

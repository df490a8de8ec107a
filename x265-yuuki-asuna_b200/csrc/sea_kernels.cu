// sea_kernels.cu -- the SEA (successive elimination) support primitives of the table:
//   * integral_inith[6] / integral_initv[6] (source/encoder/framefilter.cpp:39-140) and the whole-frame form of
//     FrameFilter::computeMEIntegral (framefilter.cpp:722-825): the 12 box-sum planes 32x32, 32x24, 32x8, 24x32,
//     16x16, 16x12, 16x4, 12x16, 8x32, 8x8, 4x16, 4x4 that MotionEstimate reads through `integral[]`
//   * pu[].ads = ads_x1 / ads_x2 / ads_x4 (source/common/pixel.cpp:121-165)
// The SEA search itself (motion.cpp:1242-1395) lives in me_device.cuh.
//
// Whole-frame kernel: the reference builds each plane as a running column sum of horizontal w-sums and turns row Y
// into a box sum h rows later (sum[Y] = sum[Y+h] - sum[Y]); the result at (x, Y) is the sum of the w x h pixels whose
// top-left corner is (x, Y).  Here one CTA owns a 128 x 32 tile of box origins: the (32+31) x (128+32) pixel tile is
// read once, turned into per-row prefix sums in shared memory, and each thread marches down its column with a sliding
// vertical window per plane -- every pixel is read from HBM ~1.5 times, the 12 uint32 planes are written once
// (the write stream, 48 B per pixel, is what bounds the kernel).
#include "common.cuh"
#include "x265b200.h"

namespace x265b200 {

namespace {

__constant__ int c_intW[12] = { 32, 32, 32, 24, 16, 16, 16, 12, 8, 8, 4, 4 };     // framefilter.cpp:767-778
__constant__ int c_intH[12] = { 32, 24, 8, 32, 16, 12, 4, 16, 32, 8, 16, 4 };

struct IntegralArgs
{
    const void* pix;          // first allocated pixel (origin - padY*stride - padX)
    int64_t stride;
    int rowsTotal;            // allocated pixel rows = maxHeight + 2*padY
    uint32_t* planes[12];     // first allocated element of every plane (same geometry as the pixel plane)
};

constexpr int IT_X = 128, IT_R = 32, IT_HALO = 32;
constexpr int IT_ROWS = IT_R + IT_HALO - 1;       // 63 pixel rows per tile
constexpr int IT_COLS = IT_X + IT_HALO;           // 160 pixel columns per tile
constexpr int IT_PITCH = IT_COLS + 1;             // prefix row: P[0] = 0 .. P[160]

template<typename pixel>
__global__ void __launch_bounds__(IT_X)
sea_integral_kernel(IntegralArgs a)
{
    __shared__ uint32_t P[IT_ROWS][IT_PITCH];
    const int c0 = blockIdx.x * IT_X, r0 = blockIdx.y * IT_R;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const pixel* pix = (const pixel*)a.pix;

    // per-row exclusive prefix sums of the tile; pixels outside the rows the reference reads (the last allocated
    // row is never summed, framefilter.cpp:761-763) or outside the stride count as 0
    for (int rr = warp; rr < IT_ROWS; rr += IT_X / 32)
    {
        const int r = r0 + rr;
        const bool rowOk = r < a.rowsTotal - 1;
        uint32_t carry = 0;
        if (lane == 0) P[rr][0] = 0;
        for (int cc = 0; cc < IT_COLS; cc += 32)
        {
            const int c = c0 + cc + lane;
            uint32_t v = (rowOk && c < a.stride) ? (uint32_t)pix[(int64_t)r * a.stride + c] : 0u;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += t;
            }
            P[rr][cc + lane + 1] = carry + v;
            carry += __shfl_sync(0xffffffffu, v, 31);
        }
    }
    __syncthreads();

    const int t = threadIdx.x, c = c0 + t;
    if (c >= a.stride) return;
    for (int k = 0; k < 12; k++)
    {
        const int w = c_intW[k], h = c_intH[k];
        uint32_t* out = a.planes[k];
        // box origins that the reference finalises: columns [0, stride - w), rows [1, rowsTotal - 1 - h]; row 0 is the
        // memset row (framefilter.cpp:757) and stays 0; everything else is written as 0 (the reference leaves running
        // sums there, which no in-bounds search reads)
        const bool colOk = c < a.stride - w;
        uint32_t acc = 0;
        for (int j = 0; j < h; j++) acc += P[j][t + w] - P[j][t];
        for (int rr = 0; rr < IT_R; rr++)
        {
            const int r = r0 + rr;
            if (r >= a.rowsTotal) break;
            const bool ok = colOk && r >= 1 && r <= a.rowsTotal - 1 - h;
            out[(int64_t)r * a.stride + c] = ok ? acc : 0u;
            if (rr + h < IT_ROWS) acc += (P[rr + h][t + w] - P[rr + h][t]) - (P[rr][t + w] - P[rr][t]);
        }
    }
}

// integral_init{4,8,12,16,24,32}h_c (framefilter.cpp:39-104): sum[x] = (pix[x] + ... + pix[x+w-1]) + sum[x - stride]
template<typename pixel>
__global__ void integral_inith_kernel(uint32_t* sum, const pixel* pix, int64_t stride, int w)
{
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= stride - w) return;
    uint32_t v = 0;
    for (int i = 0; i < w; i++) v += pix[x + i];
    sum[x] = v + sum[x - stride];
}

// integral_init{4,...,32}v_c (framefilter.cpp:106-140): sum[x] = sum[x + h*stride] - sum[x]
__global__ void integral_initv_kernel(uint32_t* sum, int64_t stride, int h)
{
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= stride) return;
    sum[x] = sum[x + (int64_t)h * stride] - sum[x];
}

// ads_x1 / ads_x2 / ads_x4 (pixel.cpp:121-165): one warp per job, candidates compacted in order with ballots
__global__ void ads_kernel(int kind, int lxHalf, const uint32_t* sums, int64_t delta, const uint16_t* costMvX, int width,
                           const x265b200_ads_job* jobs, int64_t n, int16_t* mvs, int32_t* counts)
{
    const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (j >= n) return;
    const x265b200_ads_job job = jobs[j];
    const uint32_t* s = sums + job.sumsOff;
    int16_t* out = mvs + j * width;
    int nmv = 0;
    for (int base = 0; base < width; base += 32)
    {
        const int i = base + lane;
        bool pass = false;
        if (i < width)
        {
            // `long` arithmetic truncated to int on assignment (pixel.cpp:127-131)
            int64_t ads = llabs((int64_t)job.encDC[0] - (int64_t)s[i]);
            if (kind == 2) ads += llabs((int64_t)job.encDC[1] - (int64_t)s[i + delta]);
            if (kind == 4)
                ads += llabs((int64_t)job.encDC[1] - (int64_t)s[i + lxHalf]) + llabs((int64_t)job.encDC[2] - (int64_t)s[i + delta]) +
                       llabs((int64_t)job.encDC[3] - (int64_t)s[i + delta + lxHalf]);
            ads += costMvX[i];
            pass = (int)ads < job.thresh;
        }
        const unsigned m = __ballot_sync(0xffffffffu, pass);
        if (pass) out[nmv + __popc(m & ((1u << lane) - 1))] = (int16_t)i;
        nmv += __popc(m);
    }
    if (lane == 0) counts[j] = nmv;
}

} // namespace

int sea_integral_dev(Ctx* ctx, int depth, const void* reconOrigin, int64_t stride, int padX, int padY, int maxHeight, uint32_t* const planes[12])
{
    if (stride <= 32 || maxHeight <= 0 || padX < 0 || padY < 0) { set_error("sea_integral: bad geometry"); return -1; }
    IntegralArgs a;
    const size_t px = depth > 8 ? 2 : 1;
    a.pix = (const char*)reconOrigin - ((int64_t)padY * stride + padX) * (int64_t)px;
    a.stride = stride; a.rowsTotal = maxHeight + 2 * padY;
    for (int k = 0; k < 12; k++)
    {
        if (!planes[k]) { set_error("sea_integral: plane %d is NULL", k); return -1; }
        a.planes[k] = planes[k] - ((int64_t)padY * stride + padX);
    }
    dim3 grid((unsigned)((stride + IT_X - 1) / IT_X), (unsigned)((a.rowsTotal + IT_R - 1) / IT_R));
    if (depth > 8) sea_integral_kernel<uint16_t><<<grid, IT_X, 0, ctx->stream>>>(a);
    else sea_integral_kernel<uint8_t><<<grid, IT_X, 0, ctx->stream>>>(a);
    ctx->launches++;
    return check(cudaGetLastError(), "sea_integral kernel launch");
}

int integral_inith_dev(Ctx* ctx, int depth, int w, uint32_t* sum, const void* pix, int64_t stride)
{
    if (w != 4 && w != 8 && w != 12 && w != 16 && w != 24 && w != 32) { set_error("integral_inith: width %d", w); return -1; }
    if (stride <= w) return 0;
    const unsigned blocks = (unsigned)((stride - w + 255) / 256);
    if (depth > 8) integral_inith_kernel<uint16_t><<<blocks, 256, 0, ctx->stream>>>(sum, (const uint16_t*)pix, stride, w);
    else integral_inith_kernel<uint8_t><<<blocks, 256, 0, ctx->stream>>>(sum, (const uint8_t*)pix, stride, w);
    ctx->launches++;
    return check(cudaGetLastError(), "integral_inith kernel launch");
}

int integral_initv_dev(Ctx* ctx, int h, uint32_t* sum, int64_t stride)
{
    if (h != 4 && h != 8 && h != 12 && h != 16 && h != 24 && h != 32) { set_error("integral_initv: height %d", h); return -1; }
    if (stride <= 0) return 0;
    integral_initv_kernel<<<(unsigned)((stride + 255) / 256), 256, 0, ctx->stream>>>(sum, stride, h);
    ctx->launches++;
    return check(cudaGetLastError(), "integral_initv kernel launch");
}

int ads_dev(Ctx* ctx, int kind, int lxHalf, const uint32_t* sums, int64_t delta, const uint16_t* costMvX, int width,
            const x265b200_ads_job* jobs, int64_t n, int16_t* mvs, int32_t* counts)
{
    if (kind != 1 && kind != 2 && kind != 4) { set_error("ads: kind %d (1, 2 or 4)", kind); return -1; }
    if (n <= 0) return 0;
    if (width < 0 || width > 32767) { set_error("ads: width %d", width); return -1; }
    ads_kernel<<<(unsigned)((n + 3) / 4), 128, 0, ctx->stream>>>(kind, lxHalf, sums, delta, costMvX, width, jobs, n, mvs, counts);
    ctx->launches++;
    return check(cudaGetLastError(), "ads kernel launch");
}

} // namespace x265b200

// pixel_kernels.cu -- block-compare primitives of the x265 EncoderPrimitives table as batched
// sm_100a kernels: SAD / SATD / SA8D / SSE / ssd_s over N independent block pairs, and the
// sad_x3 / sad_x4 multi-candidate forms.
//
// Reference semantics (bit-exact, integer):
//   sad<lx,ly>            source/common/pixel.cpp:40-55
//   sad_x3 / sad_x4       source/common/pixel.cpp:74-119   (fenc stride hard-wired to FENC_STRIDE=64)
//   satd_4x4 / satd_8x4   source/common/pixel.cpp:210-297  (sum|H4 d H4^T| >> 1 per 4x4 / 8x4)
//   _sa8d_8x8 / sa8d_*    source/common/pixel.cpp:299-377  ((s+2)>>2 per 8x8, or once per 16x16)
//   sse<> / ssd_s         source/common/pixel.cpp:167-186, :379-391
//
// The SWAR tricks of the C reference (two 16-bit lanes per 32-bit word) cannot overflow for
// legal pixel ranges (see DESIGN.md "exactness notes"), so the kernels compute the plain
// mathematical Hadamard sums; the per-sub-block shifts and rounding granularity are kept.
//
// Roofline: every kernel here is HBM-bound: algorithmic bytes = N * (2*w*h*sizeof(pixel) + 4).
// Work mapping: one lane per 4x4 cell (8x8 cell for SA8D); G = min(32, pow2ceil(cells)) lanes
// cooperate on one block and finish with a segmented warp-shuffle reduction.
#include "common.cuh"
#include "x265b200.h"

namespace x265b200 {

// ---------------------------------------------------------------------------------------------
// cells
// ---------------------------------------------------------------------------------------------
template<typename pixel>
__device__ __forceinline__ int cell_sad(const pixel* a, int64_t sa, const pixel* b, int64_t sb)
{
    int s = 0;
#pragma unroll
    for (int r = 0; r < 4; r++) s += sad4<pixel>(a + r * sa, b + r * sb);
    return s;
}

__device__ __forceinline__ void hadamard4(int& a, int& b, int& c, int& d)
{
    int t0 = a + b, t1 = a - b, t2 = c + d, t3 = c - d;
    a = t0 + t2; c = t0 - t2; b = t1 + t3; d = t1 - t3;
}

// sum |H4 d H4^T| / 2 of one 4x4 (pixel.cpp:210-236; the /2 is exact: all 16 coefficients
// share the parity of the block sum)
template<typename pixel>
__device__ __forceinline__ int cell_satd(const pixel* a, int64_t sa, const pixel* b, int64_t sb)
{
    int d[4][4];
#pragma unroll
    for (int r = 0; r < 4; r++)
    {
        int x[4], y[4];
        ld4i<pixel>(a + r * sa, x);
        ld4i<pixel>(b + r * sb, y);
#pragma unroll
        for (int c = 0; c < 4; c++) d[r][c] = x[c] - y[c];
        hadamard4(d[r][0], d[r][1], d[r][2], d[r][3]);
    }
    int s = 0;
#pragma unroll
    for (int c = 0; c < 4; c++)
    {
        hadamard4(d[0][c], d[1][c], d[2][c], d[3][c]);
        s += abs(d[0][c]) + abs(d[1][c]) + abs(d[2][c]) + abs(d[3][c]);
    }
    return s >> 1;
}

__device__ __forceinline__ void hadamard8(int v[8])
{
    hadamard4(v[0], v[1], v[2], v[3]);
    hadamard4(v[4], v[5], v[6], v[7]);
#pragma unroll
    for (int i = 0; i < 4; i++) { int p = v[i] + v[i + 4], m = v[i] - v[i + 4]; v[i] = p; v[i + 4] = m; }
}

// raw sum |H8 d H8^T| of one 8x8 (pixel.cpp:299-334 `_sa8d_8x8`, before rounding)
template<typename pixel>
__device__ __forceinline__ int cell_sa8d_raw(const pixel* a, int64_t sa, const pixel* b, int64_t sb)
{
    int d[8][8];
#pragma unroll
    for (int r = 0; r < 8; r++)
    {
        int x[4], y[4];
        ld4i<pixel>(a + r * sa, x); ld4i<pixel>(b + r * sb, y);
#pragma unroll
        for (int c = 0; c < 4; c++) d[r][c] = x[c] - y[c];
        ld4i<pixel>(a + r * sa + 4, x); ld4i<pixel>(b + r * sb + 4, y);
#pragma unroll
        for (int c = 0; c < 4; c++) d[r][4 + c] = x[c] - y[c];
        hadamard8(d[r]);
    }
    int s = 0;
#pragma unroll
    for (int c = 0; c < 8; c++)
    {
        int v[8];
#pragma unroll
        for (int r = 0; r < 8; r++) v[r] = d[r][c];
        hadamard8(v);
#pragma unroll
        for (int r = 0; r < 8; r++) s += abs(v[r]);
    }
    return s;
}

template<typename pixel>
__device__ __forceinline__ uint64_t cell_sse_pp(const pixel* a, int64_t sa, const pixel* b, int64_t sb)
{
    uint32_t s = 0;   // 16 * 4095^2 < 2^32
#pragma unroll
    for (int r = 0; r < 4; r++)
    {
        int x[4], y[4];
        ld4i<pixel>(a + r * sa, x); ld4i<pixel>(b + r * sb, y);
#pragma unroll
        for (int c = 0; c < 4; c++) { int t = x[c] - y[c]; s += (uint32_t)(t * t); }
    }
    return s;
}

// sse<int16,int16>: `tmp*tmp` is a wrapping int product, sign-extended when added to sse_t
__device__ __forceinline__ uint64_t cell_sse_ss(const int16_t* a, int64_t sa, const int16_t* b, int64_t sb)
{
    uint64_t s = 0;
#pragma unroll
    for (int r = 0; r < 4; r++)
    {
        int x[4], y[4];
        ld4s(a + r * sa, x); ld4s(b + r * sb, y);
#pragma unroll
        for (int c = 0; c < 4; c++)
        {
            int t = x[c] - y[c];
            int p = (int)((uint32_t)t * (uint32_t)t);
            s += (uint64_t)(int64_t)p;
        }
    }
    return s;
}

__device__ __forceinline__ uint64_t cell_ssd_s(const int16_t* a, int64_t sa)
{
    uint64_t s = 0;
#pragma unroll
    for (int r = 0; r < 4; r++)
    {
        int x[4];
        ld4s(a + r * sa, x);
#pragma unroll
        for (int c = 0; c < 4; c++) s += (uint64_t)(int64_t)(x[c] * x[c]);
    }
    return s;
}

// ---------------------------------------------------------------------------------------------
// batched compare kernel
// ---------------------------------------------------------------------------------------------
struct CmpArgs
{
    const void* A; int64_t strideA;
    const void* B; int64_t strideB;
    const int64_t* offA;     // element offsets, or nullptr => grid mode
    const int64_t* offB;
    const int16_t* mv;       // grid mode: optional full-pel {x,y} per block applied to B
    int   gridCols;          // grid mode: blocks per row
    int64_t n;
    int   w, h;
    int   depth;
    void* out;               // int32[n] (SAD/SATD/SA8D) or uint64[n] (SSE kinds)
};

template<typename pixel, int KIND>
__global__ void __launch_bounds__(256)
cmp_batch_kernel(CmpArgs p)
{
    constexpr bool is8 = (KIND == X265B200_CMP_SA8D || KIND == X265B200_CMP_SA8D8);
    constexpr int CW = is8 ? 8 : 4;
    const int cellsX = p.w / CW, cellsY = p.h / CW;
    const int ncell = cellsX * cellsY;
    int G = 1; while (G < ncell && G < 32) G <<= 1;
    const int perWarp = 32 / G;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int g = lane / G, l = lane % G;
    const int64_t blk = warp * perWarp + g;
    const bool active = blk < p.n;

    int64_t oa = 0, ob = 0;
    if (active)
    {
        if (p.offA) { oa = p.offA[blk]; ob = p.offB ? p.offB[blk] : oa; }
        else
        {
            int bx = (int)(blk % p.gridCols), by = (int)(blk / p.gridCols);
            oa = (int64_t)by * p.h * p.strideA + (int64_t)bx * p.w;
            ob = (int64_t)by * p.h * p.strideB + (int64_t)bx * p.w;
            if (p.mv) ob += p.mv[2 * blk] + (int64_t)p.mv[2 * blk + 1] * p.strideB;
        }
    }

    if (KIND == X265B200_CMP_SSE_PP || KIND == X265B200_CMP_SSE_SS || KIND == X265B200_CMP_SSD_S)
    {
        uint64_t acc = 0;
        if (active)
            for (int c = l; c < ncell; c += G)
            {
                int cx = c % cellsX, cy = c / cellsX;
                if (KIND == X265B200_CMP_SSE_PP)
                    acc += cell_sse_pp<pixel>((const pixel*)p.A + oa + (int64_t)cy * 4 * p.strideA + cx * 4, p.strideA,
                                              (const pixel*)p.B + ob + (int64_t)cy * 4 * p.strideB + cx * 4, p.strideB);
                else if (KIND == X265B200_CMP_SSE_SS)
                    acc += cell_sse_ss((const int16_t*)p.A + oa + (int64_t)cy * 4 * p.strideA + cx * 4, p.strideA,
                                       (const int16_t*)p.B + ob + (int64_t)cy * 4 * p.strideB + cx * 4, p.strideB);
                else
                    acc += cell_ssd_s((const int16_t*)p.A + oa + (int64_t)cy * 4 * p.strideA + cx * 4, p.strideA);
            }
        for (int o = G >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (active && l == 0)
            ((uint64_t*)p.out)[blk] = p.depth < 10 ? (acc & 0xffffffffull) : acc;   // sse_t: common.h:144-148
        return;
    }

    int acc = 0;
    if (is8)
    {
        // cell order: 4 consecutive cells = the four 8x8 of one 16x16 (pixel.cpp:341-351 rounds once per 16x16)
        const bool r16 = (KIND == X265B200_CMP_SA8D) && p.w >= 16 && p.h >= 16;
        const int iters = (ncell + G - 1) / G;
        for (int it = 0; it < iters; it++)
        {
            int c = it * G + l;
            int v = 0;
            if (active && c < ncell)
            {
                int cx, cy;
                if (r16) { int q = c >> 2, s = c & 3; int qx = q % (cellsX >> 1), qy = q / (cellsX >> 1); cx = qx * 2 + (s & 1); cy = qy * 2 + (s >> 1); }
                else { cx = c % cellsX; cy = c / cellsX; }
                v = cell_sa8d_raw<pixel>((const pixel*)p.A + oa + (int64_t)cy * 8 * p.strideA + cx * 8, p.strideA,
                                         (const pixel*)p.B + ob + (int64_t)cy * 8 * p.strideB + cx * 8, p.strideB);
            }
            if (r16)
            {
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                if ((l & 3) == 0) acc += (v + 2) >> 2;
            }
            else
                acc += (v + 2) >> 2;   // sa8d_8x8, pixel.cpp:336-339
        }
    }
    else if (active)
    {
        for (int c = l; c < ncell; c += G)
        {
            int cx = c % cellsX, cy = c / cellsX;
            const pixel* a = (const pixel*)p.A + oa + (int64_t)cy * 4 * p.strideA + cx * 4;
            const pixel* b = (const pixel*)p.B + ob + (int64_t)cy * 4 * p.strideB + cx * 4;
            acc += (KIND == X265B200_CMP_SAD) ? cell_sad<pixel>(a, p.strideA, b, p.strideB)
                                              : cell_satd<pixel>(a, p.strideA, b, p.strideB);
        }
    }
    for (int o = G >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (active && l == 0) ((int32_t*)p.out)[blk] = acc;
}

template<typename pixel>
static int launch_cmp(Ctx* ctx, int kind, const CmpArgs& a)
{
    const bool is8 = (kind == X265B200_CMP_SA8D || kind == X265B200_CMP_SA8D8);
    int cw = is8 ? 8 : 4;
    int ncell = (a.w / cw) * (a.h / cw);
    int G = 1; while (G < ncell && G < 32) G <<= 1;
    int perWarp = 32 / G;
    const int threads = 256;
    int64_t warps = (a.n + perWarp - 1) / perWarp;
    int64_t blocks = (warps * 32 + threads - 1) / threads;
    if (blocks <= 0) return 0;
    dim3 grid((unsigned)blocks), block(threads);
    switch (kind)
    {
    case X265B200_CMP_SAD:    cmp_batch_kernel<pixel, X265B200_CMP_SAD><<<grid, block, 0, ctx->stream>>>(a); break;
    case X265B200_CMP_SATD:   cmp_batch_kernel<pixel, X265B200_CMP_SATD><<<grid, block, 0, ctx->stream>>>(a); break;
    case X265B200_CMP_SA8D:   cmp_batch_kernel<pixel, X265B200_CMP_SA8D><<<grid, block, 0, ctx->stream>>>(a); break;
    case X265B200_CMP_SA8D8:  cmp_batch_kernel<pixel, X265B200_CMP_SA8D8><<<grid, block, 0, ctx->stream>>>(a); break;
    case X265B200_CMP_SSE_PP: cmp_batch_kernel<pixel, X265B200_CMP_SSE_PP><<<grid, block, 0, ctx->stream>>>(a); break;
    case X265B200_CMP_SSE_SS: cmp_batch_kernel<pixel, X265B200_CMP_SSE_SS><<<grid, block, 0, ctx->stream>>>(a); break;
    case X265B200_CMP_SSD_S:  cmp_batch_kernel<pixel, X265B200_CMP_SSD_S><<<grid, block, 0, ctx->stream>>>(a); break;
    default: set_error("pixelcmp: unknown kind %d", kind); return -1;
    }
    ctx->launches++;
    return check(cudaGetLastError(), "cmp_batch_kernel launch");
}

int pixelcmp_dev(Ctx* ctx, int kind, int depth, int w, int h,
                 const void* A, int64_t strideA, const void* B, int64_t strideB,
                 const int64_t* offA, const int64_t* offB, const int16_t* mv, int gridCols,
                 int64_t n, void* out)
{
    const bool is8 = (kind == X265B200_CMP_SA8D || kind == X265B200_CMP_SA8D8);
    if (is8 && w == 4 && h == 4) kind = X265B200_CMP_SATD;          // cu[BLOCK_4x4].sa8d = satd_4x4 (pixel.cpp:1163)
    else if (is8 && ((w | h) & 7)) { set_error("sa8d needs w,h multiples of 8 (got %dx%d)", w, h); return -1; }
    if (w <= 0 || h <= 0 || ((w | h) & 3) || w > 64 || h > 64) { set_error("pixelcmp: unsupported block %dx%d", w, h); return -1; }
    CmpArgs a; a.A = A; a.strideA = strideA; a.B = B; a.strideB = strideB; a.offA = offA; a.offB = offB; a.mv = mv;
    a.gridCols = gridCols; a.n = n; a.w = w; a.h = h; a.depth = depth; a.out = out;
    return depth > 8 ? launch_cmp<uint16_t>(ctx, kind, a) : launch_cmp<uint8_t>(ctx, kind, a);
}

// ---------------------------------------------------------------------------------------------
// sad_x3 / sad_x4: one cached 64-stride fenc block against K reference blocks sharing a stride
// (pixel.cpp:74-119).  One warp per item: lanes stride over the K*w*h/4 four-pixel groups.
// ---------------------------------------------------------------------------------------------
template<typename pixel>
__global__ void __launch_bounds__(256)
sad_xn_kernel(const pixel* __restrict__ fenc, int64_t fencBlockStride,   // item i's cached PU at fenc + i*fencBlockStride (row stride 64)
              const pixel* __restrict__ ref, int64_t refStride,
              const int64_t* __restrict__ refOff,                        // [n][K] element offsets into ref
              int K, int w, int h, int64_t n, int32_t* __restrict__ res) // [n][K]
{
    const int lane = threadIdx.x & 31;
    const int64_t item = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (item >= n) return;
    const pixel* f = fenc + item * fencBlockStride;
    const int gw = w >> 2, ngroups = gw * h;
    for (int k = 0; k < K; k++)
    {
        const pixel* r = ref + refOff[item * K + k];
        int s = 0;
        for (int u = lane; u < ngroups; u += 32)
        {
            int y = u / gw, x = (u - y * gw) << 2;
            s += sad4<pixel>(f + y * 64 + x, r + (int64_t)y * refStride + x);
        }
        s = warp_sum(s);
        if (lane == 0) res[item * K + k] = s;
    }
}

int sad_xn_dev(Ctx* ctx, int depth, int K, int w, int h, const void* fenc, int64_t fencBlockStride,
               const void* ref, int64_t refStride, const int64_t* refOff, int64_t n, int32_t* res)
{
    if (K < 1 || K > 4) { set_error("sad_xN: K=%d", K); return -1; }
    if (w <= 0 || h <= 0 || (w & 3) || w > 64 || h > 64) { set_error("sad_xN: unsupported block %dx%d", w, h); return -1; }
    const int threads = 256;
    int64_t blocks = (n * 32 + threads - 1) / threads;
    if (blocks <= 0) return 0;
    if (depth > 8)
        sad_xn_kernel<uint16_t><<<(unsigned)blocks, threads, 0, ctx->stream>>>((const uint16_t*)fenc, fencBlockStride, (const uint16_t*)ref, refStride, refOff, K, w, h, n, res);
    else
        sad_xn_kernel<uint8_t><<<(unsigned)blocks, threads, 0, ctx->stream>>>((const uint8_t*)fenc, fencBlockStride, (const uint8_t*)ref, refStride, refOff, K, w, h, n, res);
    ctx->launches++;
    return check(cudaGetLastError(), "sad_xn_kernel launch");
}

// ---------------------------------------------------------------------------------------------
// plane-wide sub_ps / add_ps (pixel.cpp:814-840): 4 pixels per thread
// ---------------------------------------------------------------------------------------------
template<typename pixel>
__global__ void __launch_bounds__(256)
sub_ps_plane_kernel(const pixel* __restrict__ a, int64_t sa, const pixel* __restrict__ b, int64_t sb, int16_t* __restrict__ d, int64_t sd, int w, int h)
{
    const int gw = (w + 3) >> 2;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)gw * h) return;
    int y = (int)(i / gw), x = (int)(i - (int64_t)y * gw) << 2;
    for (int k = 0; k < 4 && x + k < w; k++)
        d[y * sd + x + k] = (int16_t)((int)a[y * sa + x + k] - (int)b[y * sb + x + k]);
}
template<typename pixel>
__global__ void __launch_bounds__(256)
add_ps_plane_kernel(pixel* __restrict__ d, int64_t sd, const pixel* __restrict__ p, int64_t sp, const int16_t* __restrict__ r, int64_t sr, int w, int h, int maxVal)
{
    const int gw = (w + 3) >> 2;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)gw * h) return;
    int y = (int)(i / gw), x = (int)(i - (int64_t)y * gw) << 2;
    for (int k = 0; k < 4 && x + k < w; k++)
        d[y * sd + x + k] = (pixel)clip3i(0, maxVal, (int)p[y * sp + x + k] + (int)r[y * sr + x + k]);
}

int sub_ps_plane_dev(Ctx* ctx, int depth, const void* a, int64_t strideA, const void* b, int64_t strideB, int16_t* dst, int64_t dstStride, int w, int h)
{
    int64_t n = (int64_t)((w + 3) >> 2) * h;
    if (n <= 0) return 0;
    unsigned blocks = (unsigned)((n + 255) / 256);
    if (depth > 8) sub_ps_plane_kernel<uint16_t><<<blocks, 256, 0, ctx->stream>>>((const uint16_t*)a, strideA, (const uint16_t*)b, strideB, dst, dstStride, w, h);
    else           sub_ps_plane_kernel<uint8_t><<<blocks, 256, 0, ctx->stream>>>((const uint8_t*)a, strideA, (const uint8_t*)b, strideB, dst, dstStride, w, h);
    ctx->launches++;
    return check(cudaGetLastError(), "sub_ps_plane launch");
}

int add_ps_plane_dev(Ctx* ctx, int depth, void* dst, int64_t dstStride, const void* pred, int64_t predStride, const int16_t* resi, int64_t resiStride, int w, int h)
{
    int64_t n = (int64_t)((w + 3) >> 2) * h;
    if (n <= 0) return 0;
    unsigned blocks = (unsigned)((n + 255) / 256);
    if (depth > 8) add_ps_plane_kernel<uint16_t><<<blocks, 256, 0, ctx->stream>>>((uint16_t*)dst, dstStride, (const uint16_t*)pred, predStride, resi, resiStride, w, h, (1 << depth) - 1);
    else           add_ps_plane_kernel<uint8_t><<<blocks, 256, 0, ctx->stream>>>((uint8_t*)dst, dstStride, (const uint8_t*)pred, predStride, resi, resiStride, w, h, 255);
    ctx->launches++;
    return check(cudaGetLastError(), "add_ps_plane launch");
}


// ---------------------------------------------------------------------------------------------
// SAD pyramid: ONE streaming pass over a frame pair produces the SADs of every 8x8, 16x16, 32x32 and
// 64x64 2Nx2N PU of every CTU (the "cost at the predictor" step of motionEstimate, motion.cpp:771-784,
// for all PU levels at once; sad<lx,ly> of pixel.cpp:40-55 is additive over sub-blocks).  Each plane
// byte is fetched exactly once with 128-bit loads; per-CTU full-pel displacement optional; all references of a
// frame are processed by one launch (grid.y = reference).
// Mapping: see the kernel comment (CTA = 64 rows x 512 px, lane = 16x16 block).
// Algorithmic bytes per launch: 2*W*H + 4*(n8 + n16 + n32 + n64).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ld16_any(const uint8_t* p)
{
    uintptr_t a = (uintptr_t)p;
    if ((a & 15) == 0) return __ldg((const uint4*)p);
    const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3) * 8;
    uint32_t t0 = __ldg(w), t1 = __ldg(w + 1), t2 = __ldg(w + 2), t3 = __ldg(w + 3);
    if (!sh) return make_uint4(t0, t1, t2, t3);
    uint32_t t4 = __ldg(w + 4);
    return make_uint4(__funnelshift_r(t0, t1, sh), __funnelshift_r(t1, t2, sh), __funnelshift_r(t2, t3, sh), __funnelshift_r(t3, t4, sh));
}

// CTA = 4 warps = one CTU row (64 rows) x 512 px (8 CTUs): warp w owns the 16-row band w, lane l the 16-px column l, so
// every row is read as one contiguous 512-byte run per warp (DRAM-friendly); 32x32 / 64x64 sums combine lanes by
// shuffle and the 4 row bands through shared memory.
template<typename pixel>
__global__ void __launch_bounds__(128, 8)
sad_pyramid_kernel(const pixel* __restrict__ cur, int64_t strideC, const pixel* const* __restrict__ refs, int64_t strideR,
                   int ctuCols, int ctuRows, const int16_t* __restrict__ mvCtu,
                   int32_t* __restrict__ out8, int32_t* __restrict__ out16, int32_t* __restrict__ out32, int32_t* __restrict__ out64)
{
    __shared__ int sBand[4][32];
    // blockIdx.y = reference index: all references of a frame in ONE launch (the source CTU rows are shared in L2)
    const pixel* __restrict__ ref = refs[blockIdx.y];
    {
        const int64_t nctu = (int64_t)ctuCols * ctuRows;
        out8 += blockIdx.y * nctu * 64; out16 += blockIdx.y * nctu * 16; out32 += blockIdx.y * nctu * 4; out64 += blockIdx.y * nctu;
        if (mvCtu) mvCtu += blockIdx.y * nctu * 2;
    }
    const int lane = threadIdx.x & 31, rsub = threadIdx.x >> 5;
    const int chunksPerRow = (ctuCols + 7) >> 3;
    const int ctuY = blockIdx.x / chunksPerRow, chunk = blockIdx.x % chunksPerRow;
    const int ctuX = chunk * 8 + (lane >> 2);
    const bool valid = ctuX < ctuCols;
    const int x0 = chunk * 512 + lane * 16, y0 = ctuY * 64 + rsub * 16;
    int mvx = 0, mvy = 0;
    if (mvCtu && valid) { mvx = mvCtu[2 * (ctuY * ctuCols + ctuX)]; mvy = mvCtu[2 * (ctuY * ctuCols + ctuX) + 1]; }
    const pixel* c = cur + (int64_t)y0 * strideC + x0;
    const pixel* r = ref + (int64_t)(y0 + mvy) * strideR + x0 + mvx;
    uint32_t s8[4] = { 0, 0, 0, 0 };        // [row group 0/1][left/right 8x8]
    if (valid)
    {
        // 4 batches of 4 rows: 8 (16-bit: 16) independent 128-bit loads in flight per lane
#pragma unroll
        for (int g = 0; g < 4; g++)
        {
            if constexpr (sizeof(pixel) == 1)
            {
                uint4 a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; i++) { a[i] = __ldg((const uint4*)(c + (int64_t)(g * 4 + i) * strideC)); b[i] = ld16_any(r + (int64_t)(g * 4 + i) * strideR); }
#pragma unroll
                for (int i = 0; i < 4; i++)
                {
                    s8[(g >> 1) * 2 + 0] += __vsadu4(a[i].x, b[i].x) + __vsadu4(a[i].y, b[i].y);
                    s8[(g >> 1) * 2 + 1] += __vsadu4(a[i].z, b[i].z) + __vsadu4(a[i].w, b[i].w);
                }
            }
            else
            {
                // 16-bit samples: a row of the lane's 16 pixels is two vectors, one per 8x8 half
                uint4 a[4][2], b[4][2];
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int h = 0; h < 2; h++)
                    {
                        a[i][h] = __ldg((const uint4*)(c + (int64_t)(g * 4 + i) * strideC + 8 * h));
                        b[i][h] = ld16_any((const uint8_t*)(r + (int64_t)(g * 4 + i) * strideR + 8 * h));
                    }
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int h = 0; h < 2; h++)
                        s8[(g >> 1) * 2 + h] += sad_u16x2(a[i][h].x, b[i][h].x) + sad_u16x2(a[i][h].y, b[i][h].y) + sad_u16x2(a[i][h].z, b[i][h].z) + sad_u16x2(a[i][h].w, b[i][h].w);
            }
        }
    }
    const int n8x = ctuCols * 8, n16x = ctuCols * 4, n32x = ctuCols * 2;
    if (valid)
    {
        const int bx = x0 >> 3, by = y0 >> 3;
        *(int2*)(out8 + (int64_t)by * n8x + bx) = make_int2((int)s8[0], (int)s8[1]);
        *(int2*)(out8 + (int64_t)(by + 1) * n8x + bx) = make_int2((int)s8[2], (int)s8[3]);
    }
    const int s16 = (int)(s8[0] + s8[1] + s8[2] + s8[3]);
    if (valid) out16[(int64_t)(y0 >> 4) * n16x + (x0 >> 4)] = s16;
    const int h32 = s16 + __shfl_xor_sync(0xffffffffu, s16, 1);            // 32 px wide, 16 rows
    sBand[rsub][lane] = h32;
    __syncthreads();
    if (valid && !(lane & 1) && !(rsub & 1))
        out32[(int64_t)(y0 >> 5) * n32x + (x0 >> 5)] = sBand[rsub][lane] + sBand[rsub + 1][lane];
    if (valid && !(lane & 3) && rsub == 0)
    {
        int s64 = 0;
#pragma unroll
        for (int w = 0; w < 4; w++) s64 += sBand[w][lane] + sBand[w][lane + 2];
        out64[(int64_t)ctuY * ctuCols + ctuX] = s64;
    }
}

int sad_pyramid_dev(Ctx* ctx, int depth, const void* cur, int64_t strideC, const void* const* refs, int numRefs, int64_t strideR, int ctuCols, int ctuRows,
                    const int16_t* mvCtu, int32_t* out8, int32_t* out16, int32_t* out32, int32_t* out64)
{
    const int px = depth > 8 ? 2 : 1;
    if (((uintptr_t)cur & 15) || ((strideC * px) & 15)) { set_error("sad_pyramid: cur plane and row pitch must be 16-byte aligned"); return -1; }
    if (((uintptr_t)out8 & 7)) { set_error("sad_pyramid: out8 must be 8-byte aligned"); return -1; }
    int64_t ctas = (int64_t)((ctuCols + 7) / 8) * ctuRows;
    if (ctas <= 0 || numRefs <= 0) return 0;
    dim3 grid((unsigned)ctas, (unsigned)numRefs);
    if (depth > 8) sad_pyramid_kernel<uint16_t><<<grid, 128, 0, ctx->stream>>>((const uint16_t*)cur, strideC, (const uint16_t* const*)refs, strideR, ctuCols, ctuRows, mvCtu, out8, out16, out32, out64);
    else           sad_pyramid_kernel<uint8_t><<<grid, 128, 0, ctx->stream>>>((const uint8_t*)cur, strideC, (const uint8_t* const*)refs, strideR, ctuCols, ctuRows, mvCtu, out8, out16, out32, out64);
    ctx->launches++;
    return check(cudaGetLastError(), "sad_pyramid launch");
}

} // namespace x265b200

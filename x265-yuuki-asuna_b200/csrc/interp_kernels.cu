// interp_kernels.cu -- HEVC sub-pel interpolation filters (8-tap luma / 4-tap chroma) as a
// batched sm_100a kernel.  One CTA per job; hvpp stages the horizontal int16 intermediate in
// shared memory so every source pixel is fetched once.
//
// Reference semantics (bit-exact), all in source/common/ipfilter.cpp:
//   HPP  interp_horiz_pp_c :79-118    (sum+32)>>6, cast to int16, then clip to [0,maxVal]
//   HPS  interp_horiz_ps_c :120-162   offset = -8192<<shift, shift = depth-8; isRowExt adds N-1 rows
//   VPP  interp_vert_pp_c  :164-203
//   VPS  interp_vert_ps_c  :205-239
//   VSP  interp_vert_sp_c  :241-282   shift = 6+(14-depth), offset = (1<<(shift-1)) + (8192<<6)
//   VSS  interp_vert_ss_c  :284-317   sum>>6
//   HVPP interp_hv_pp_c    :362-369   = HPS(isRowExt=1) into a width-stride buffer, then VSP
//   P2S  filterPixelToShort_c :40-57  (src<<(14-depth)) - 8192
// Filter taps: constants.cpp:250-268 (regenerated in tables.cuh).
//
// Roofline: HBM-bound, algorithmic bytes per job = (w+N-1)(h+N-1)*sizeof(pixel) + w*h*sizeof(out).
#include "common.cuh"
#include "tables.cuh"
#include "x265b200.h"

namespace x265b200 {

__constant__ int16_t c_lumaFilter[4][8];
__constant__ int16_t c_chromaFilter[8][4];
static bool g_filtUploaded[16] = { false };

int upload_filters(Ctx* ctx)
{
    if (ctx->device < 16 && g_filtUploaded[ctx->device]) return 0;
    X265B200_CHECK(cudaMemcpyToSymbolAsync(c_lumaFilter, kLumaFilter, sizeof(kLumaFilter), 0, cudaMemcpyHostToDevice, ctx->stream));
    X265B200_CHECK(cudaMemcpyToSymbolAsync(c_chromaFilter, kChromaFilter, sizeof(kChromaFilter), 0, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->device < 16) g_filtUploaded[ctx->device] = true;
    return 0;
}

struct InterpArgs
{
    const void* src; int64_t srcStride;
    void* dst;       int64_t dstStride;
    const x265b200_interp_job* jobs; int64_t n;
    int kind, taps, depth, w, h, isRowExt;
};

template<int N> __device__ __forceinline__ void load_taps(int idx, int c[8])
{
#pragma unroll
    for (int i = 0; i < N; i++) c[i] = (N == 8) ? c_lumaFilter[idx][i] : c_chromaFilter[idx][i];
}

template<typename pixel, int N>
__global__ void __launch_bounds__(128)
interp_kernel(InterpArgs p)
{
    extern __shared__ int16_t immed[];        // HVPP only: w * (h + N - 1)
    const x265b200_interp_job job = p.jobs[blockIdx.x];
    const int w = p.w, h = p.h, depth = p.depth;
    const int maxVal = (1 << depth) - 1;
    const int headRoom = 14 - depth;
    int c[8];

    if (p.kind == X265B200_IP_HVPP)
    {
        // horizontal ps pass with row extension: rows -(N/2-1) .. h+N/2-1
        load_taps<N>(job.idxX, c);
        const pixel* s = (const pixel*)p.src + job.srcOff - (N / 2 - 1) - (int64_t)(N / 2 - 1) * p.srcStride;
        const int shift = 6 - headRoom, offset = (int)((unsigned)-8192 << shift);
        const int rows = h + N - 1;
        for (int e = threadIdx.x; e < rows * w; e += blockDim.x)
        {
            int y = e / w, x = e - y * w;
            const pixel* q = s + (int64_t)y * p.srcStride + x;
            int sum = 0;
#pragma unroll
            for (int t = 0; t < N; t++) sum += (int)q[t] * c[t];
            immed[e] = (int16_t)((sum + offset) >> shift);
        }
        __syncthreads();
        load_taps<N>(job.idxY, c);
        const int shift2 = 6 + headRoom, offset2 = (1 << (shift2 - 1)) + (8192 << 6);
        pixel* d = (pixel*)p.dst + job.dstOff;
        for (int e = threadIdx.x; e < h * w; e += blockDim.x)
        {
            int y = e / w, x = e - y * w;
            int sum = 0;
#pragma unroll
            for (int t = 0; t < N; t++) sum += (int)immed[(y + t) * w + x] * c[t];
            int val = (int16_t)((sum + offset2) >> shift2);
            val = val < 0 ? 0 : (val > maxVal ? maxVal : val);
            d[(int64_t)y * p.dstStride + x] = (pixel)val;
        }
        return;
    }

    if (p.kind == X265B200_IP_P2S)
    {
        const pixel* s = (const pixel*)p.src + job.srcOff;
        int16_t* d = (int16_t*)p.dst + job.dstOff;
        for (int e = threadIdx.x; e < h * w; e += blockDim.x)
        {
            int y = e / w, x = e - y * w;
            int16_t val = (int16_t)((int)s[(int64_t)y * p.srcStride + x] << headRoom);
            d[(int64_t)y * p.dstStride + x] = (int16_t)(val - (int16_t)8192);
        }
        return;
    }

    load_taps<N>(job.idxX, c);
    const bool horiz = (p.kind == X265B200_IP_HPP || p.kind == X265B200_IP_HPS);
    const bool srcShort = (p.kind == X265B200_IP_VSP || p.kind == X265B200_IP_VSS);
    const bool dstShort = (p.kind == X265B200_IP_HPS || p.kind == X265B200_IP_VPS || p.kind == X265B200_IP_VSS);
    const int64_t tapStep = horiz ? 1 : p.srcStride;
    int rows = h;
    int64_t base = job.srcOff - (N / 2 - 1) * tapStep;
    if (p.kind == X265B200_IP_HPS && p.isRowExt) { base -= (int64_t)(N / 2 - 1) * p.srcStride; rows += N - 1; }

    int shift, offset;
    switch (p.kind)
    {
    case X265B200_IP_HPP: case X265B200_IP_VPP: shift = 6; offset = 32; break;
    case X265B200_IP_HPS: case X265B200_IP_VPS: shift = 6 - headRoom; offset = (int)((unsigned)-8192 << shift); break;
    case X265B200_IP_VSP: shift = 6 + headRoom; offset = (1 << (shift - 1)) + (8192 << 6); break;
    default /* VSS */:    shift = 6; offset = 0; break;
    }

    for (int e = threadIdx.x; e < rows * w; e += blockDim.x)
    {
        int y = e / w, x = e - y * w;
        int64_t o = base + (int64_t)y * p.srcStride + x;
        int sum = 0;
        if (srcShort)
        {
            const int16_t* q = (const int16_t*)p.src + o;
#pragma unroll
            for (int t = 0; t < N; t++) sum += (int)q[t * tapStep] * c[t];
        }
        else
        {
            const pixel* q = (const pixel*)p.src + o;
#pragma unroll
            for (int t = 0; t < N; t++) sum += (int)q[t * tapStep] * c[t];
        }
        int val = (int16_t)((sum + offset) >> shift);
        int64_t di = job.dstOff + (int64_t)y * p.dstStride + x;
        if (dstShort) ((int16_t*)p.dst)[di] = (int16_t)val;
        else
        {
            val = val < 0 ? 0 : (val > maxVal ? maxVal : val);
            ((pixel*)p.dst)[di] = (pixel)val;
        }
    }
}

int interp_dev(Ctx* ctx, int kind, int taps, int depth, int w, int h, const void* src, int64_t srcStride,
               void* dst, int64_t dstStride, const x265b200_interp_job* jobs, int64_t n, int isRowExt)
{
    if (n <= 0) return 0;
    if (taps != 8 && taps != 4) { set_error("interp: taps must be 8 (luma) or 4 (chroma), got %d", taps); return -1; }
    if (kind < 0 || kind > X265B200_IP_P2S) { set_error("interp: kind %d", kind); return -1; }
    if (w < 2 || h < 2 || w > 64 || h > 64) { set_error("interp: block %dx%d", w, h); return -1; }
    if (upload_filters(ctx)) return -1;
    InterpArgs a; a.src = src; a.srcStride = srcStride; a.dst = dst; a.dstStride = dstStride; a.jobs = jobs; a.n = n;
    a.kind = kind; a.taps = taps; a.depth = depth; a.w = w; a.h = h; a.isRowExt = isRowExt;
    size_t smem = kind == X265B200_IP_HVPP ? (size_t)w * (h + taps - 1) * sizeof(int16_t) : 0;
    dim3 grid((unsigned)n), block(128);
    if (depth > 8)
    {
        if (taps == 8) interp_kernel<uint16_t, 8><<<grid, block, smem, ctx->stream>>>(a);
        else           interp_kernel<uint16_t, 4><<<grid, block, smem, ctx->stream>>>(a);
    }
    else
    {
        if (taps == 8) interp_kernel<uint8_t, 8><<<grid, block, smem, ctx->stream>>>(a);
        else           interp_kernel<uint8_t, 4><<<grid, block, smem, ctx->stream>>>(a);
    }
    ctx->launches++;
    return check(cudaGetLastError(), "interp kernel launch");
}

} // namespace x265b200

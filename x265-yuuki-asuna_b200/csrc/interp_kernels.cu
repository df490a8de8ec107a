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
#include "interp_cell.cuh"
#include <cstdlib>

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

template<int N> __device__ __forceinline__ void load_taps(int idx, int c[8])
{
#pragma unroll
    for (int i = 0; i < N; i++) c[i] = (N == 8) ? c_lumaFilter[idx][i] : c_chromaFilter[idx][i];
}

// One CTA serves JPC jobs (blocks of <= 16x16 share a CTA: 128 threads on a 64-pixel block would idle half of them and
// cost one CTA launch per 8x8 block); the TPJ = 128 / JPC threads of a job stride over its pixels.
template<typename pixel, int N>
__global__ void __launch_bounds__(128)
interp_kernel(InterpArgs p, int jpc)
{
    extern __shared__ int16_t immedAll[];     // HVPP only: jpc * w * (h + N - 1)
    const int tpj = 128 / jpc, sub = threadIdx.x / tpj, tid = threadIdx.x - sub * tpj;
    const int64_t jidx = (int64_t)blockIdx.x * jpc + sub;
    const bool live = jidx < p.n;
    x265b200_interp_job job; job.srcOff = 0; job.dstOff = 0; job.idxX = 0; job.idxY = 0;
    if (live) job = p.jobs[jidx];
    const int w = p.w, h = p.h, depth = p.depth;
    const int maxVal = (1 << depth) - 1;
    const int headRoom = 14 - depth;
    int c[8];

    if (p.kind == X265B200_IP_HVPP)
    {
        int16_t* immed = immedAll + (size_t)sub * w * (h + N - 1);
        // horizontal ps pass with row extension: rows -(N/2-1) .. h+N/2-1
        load_taps<N>(job.idxX, c);
        const pixel* s = (const pixel*)p.src + job.srcOff - (N / 2 - 1) - (int64_t)(N / 2 - 1) * p.srcStride;
        const int shift = 6 - headRoom, offset = (int)((unsigned)-8192 << shift);
        const int rows = h + N - 1;
        if (live)
            for (int e = tid; e < rows * w; e += tpj)
            {
                int y = e / w, x = e - y * w;
                const pixel* q = s + (int64_t)y * p.srcStride + x;
                int sum = 0;
#pragma unroll
                for (int t = 0; t < N; t++) sum += (int)q[t] * c[t];
                immed[e] = (int16_t)((sum + offset) >> shift);
            }
        __syncthreads();
        if (!live) return;
        load_taps<N>(job.idxY, c);
        const int shift2 = 6 + headRoom, offset2 = (1 << (shift2 - 1)) + (8192 << 6);
        pixel* d = (pixel*)p.dst + job.dstOff;
        for (int e = tid; e < h * w; e += tpj)
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
    if (!live) return;

    if (p.kind == X265B200_IP_P2S)
    {
        const pixel* s = (const pixel*)p.src + job.srcOff;
        int16_t* d = (int16_t*)p.dst + job.dstOff;
        for (int e = tid; e < h * w; e += tpj)
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

    for (int e = tid; e < rows * w; e += tpj)
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

// Cell form of the 8-bit luma pp kinds (interp_cell.cuh): one thread per 4x4 output cell on packed words.  Measured 2.78x the
// pixel-per-thread kernel on the bench's 173 400 blocks with identical planes (profiles/r02_staged_ab.txt); the same source
// runs on the host in tests/test_interp_cell_cpu.py.
__global__ void __launch_bounds__(128)
interp_pp8_cell_kernel(InterpArgs p)
{
    interp_pp8_cell_thread(p, (int64_t)blockIdx.x * 128 + threadIdx.x, c_lumaFilter);
}

// Several (kind, block size) segments in ONE launch: the bench's MC stage is 12 short launches of this kernel (one per PU level and
// filter kind, 8 us each, mostly ramp and tail); here a CTA finds its segment from the prefix sums of the segments' CTA counts.
constexpr int IP_MAX_SEGS = 16;
struct InterpMultiArgs { InterpArgs seg[IP_MAX_SEGS]; uint32_t firstCta[IP_MAX_SEGS + 1]; int numSegs; };
__global__ void __launch_bounds__(128)
interp_pp8_cell_multi_kernel(const __grid_constant__ InterpMultiArgs m)
{
    int k = 0;
#pragma unroll 1
    while (k + 1 < m.numSegs && blockIdx.x >= m.firstCta[k + 1]) k++;
    interp_pp8_cell_thread(m.seg[k], (int64_t)(blockIdx.x - m.firstCta[k]) * 128 + threadIdx.x, c_lumaFilter);
}

static bool interp_cell_ok(int kind, int taps, int depth, int w, int h, int64_t srcStride)
{
    return depth == 8 && taps == 8 && !(w & 3) && !(h & 3) && !(srcStride & 3) &&
           (kind == X265B200_IP_HPP || kind == X265B200_IP_VPP || kind == X265B200_IP_HVPP);
}

int interp_dev(Ctx* ctx, int kind, int taps, int depth, int w, int h, const void* src, int64_t srcStride,
               void* dst, int64_t dstStride, const x265b200_interp_job* jobs, int64_t n, int isRowExt);

int interp_multi_dev(Ctx* ctx, int taps, int depth, const x265b200_interp_seg* segs, int numSegs)
{
    if (numSegs <= 0) return 0;
    if (!segs) { set_error("interp_multi: segs == NULL"); return -1; }
    int done = 0;
    while (done < numSegs)
    {
        // greedily pack consecutive segments the cell kernel takes into one launch; anything else goes through interp_dev
        InterpMultiArgs m; m.numSegs = 0; m.firstCta[0] = 0;
        int k = done;
        for (; k < numSegs && m.numSegs < IP_MAX_SEGS; k++)
        {
            const x265b200_interp_seg& g = segs[k];
            if (g.n <= 0) continue;
            if (!interp_cell_ok(g.kind, taps, depth, g.w, g.h, g.srcStride)) break;
            const int64_t ctas = (g.n * (g.w >> 2) * (g.h >> 2) + 127) / 128;
            if ((int64_t)m.firstCta[m.numSegs] + ctas >= 0x7fffffffLL) break;
            InterpArgs& a = m.seg[m.numSegs];
            a.src = g.src; a.srcStride = g.srcStride; a.dst = g.dst; a.dstStride = g.dstStride; a.jobs = g.jobs; a.n = g.n;
            a.kind = g.kind; a.taps = taps; a.depth = depth; a.w = g.w; a.h = g.h; a.isRowExt = g.isRowExt;
            m.firstCta[m.numSegs + 1] = m.firstCta[m.numSegs] + (uint32_t)ctas;
            m.numSegs++;
        }
        if (m.numSegs)
        {
            if (upload_filters(ctx)) return -1;
            interp_pp8_cell_multi_kernel<<<m.firstCta[m.numSegs], 128, 0, ctx->stream>>>(m);
            ctx->launches++;
            if (check(cudaGetLastError(), "interp multi kernel launch")) return -1;
        }
        if (k < numSegs && k == done + 0 && !m.numSegs)
        {
            const x265b200_interp_seg& g = segs[k];
            if (g.n > 0 && interp_dev(ctx, g.kind, taps, depth, g.w, g.h, g.src, g.srcStride, g.dst, g.dstStride, g.jobs, g.n, g.isRowExt)) return -1;
            k++;
        }
        done = k;
    }
    return 0;
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
    if (depth == 8 && taps == 8 && !(w & 3) && !(h & 3) && !(srcStride & 3) &&
        (kind == X265B200_IP_HPP || kind == X265B200_IP_VPP || kind == X265B200_IP_HVPP))
    {
        const int64_t threads = n * (w >> 2) * (h >> 2);
        if ((threads + 127) / 128 < 0x7fffffffLL)
        {
            interp_pp8_cell_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, ctx->stream>>>(a);
            ctx->launches++;
            return check(cudaGetLastError(), "interp cell kernel launch");
        }
    }
    // jobs per CTA: ~64 output pixels per thread-group keeps every thread busy for small blocks
    const int jpc = w * h <= 32 ? 8 : (w * h <= 64 ? 4 : (w * h <= 128 ? 2 : 1));
    size_t smem = kind == X265B200_IP_HVPP ? (size_t)jpc * w * (h + taps - 1) * sizeof(int16_t) : 0;
    dim3 grid((unsigned)((n + jpc - 1) / jpc)), block(128);
    if (depth > 8)
    {
        if (taps == 8) interp_kernel<uint16_t, 8><<<grid, block, smem, ctx->stream>>>(a, jpc);
        else           interp_kernel<uint16_t, 4><<<grid, block, smem, ctx->stream>>>(a, jpc);
    }
    else
    {
        if (taps == 8) interp_kernel<uint8_t, 8><<<grid, block, smem, ctx->stream>>>(a, jpc);
        else           interp_kernel<uint8_t, 4><<<grid, block, smem, ctx->stream>>>(a, jpc);
    }
    ctx->launches++;
    return check(cudaGetLastError(), "interp kernel launch");
}

// ---------------------------------------------------------------------------------------------
// Motion compensation driver: Predict::motionCompensation (common/predict.cpp:77-257) for a batch of
// PUs -- CUData::clipMv (cudata.cpp:1915-1928), predInterLumaPixel/Short + predInterChromaPixel/Short (predict.cpp:259-406:
// copy / hpp / vpp / hps+vsp for the pixel path, p2s / hps / vps / hps+vss for the 14-bit path), then bi-prediction
// addAvg (pixel.cpp:842), addWeightUni = weight_sp (predict.cpp:525-575, pixel.cpp:493) or addWeightBi (predict.cpp:411-522).
// One CTA per (PU, plane); the 14-bit intermediates of the two lists live in shared memory and never touch HBM.
// ---------------------------------------------------------------------------------------------
struct McArgs
{
    const void* const* refs;                 // device array [2][maxRefs][3] of plane origins
    int64_t refStride[3];
    void* pred[3]; int64_t predStride[3];
    const x265b200_mc_job* jobs; int64_t n;
    const x265b200_mc_weight* weights;       // device array [2][maxRefs][3], or nullptr
    int maxRefs, hshift, vshift, depth, isP, wpP, wpB, picW, picH, maxCU;
    int planes[3], nPlanes;
};

// One block through the interpolation filters.  toShort: 14-bit intermediate into dsh (row pitch w); else pixels into dpix.
template<typename pixel, int N>
__device__ void mc_fir(const pixel* src, int64_t ss, int w, int h, int xf, int yf, int depth, bool toShort,
                       pixel* dpix, int64_t dps, int16_t* dsh, int16_t* immed)
{
    const int headRoom = 14 - depth, maxVal = (1 << depth) - 1, half = N / 2 - 1;
    const int shPS = 6 - headRoom, offPS = (int)((unsigned)-8192 << shPS);
    int cx[N], cy[N];
#pragma unroll
    for (int t = 0; t < N; t++) { cx[t] = N == 8 ? c_lumaFilter[xf & 3][t] : c_chromaFilter[xf][t]; cy[t] = N == 8 ? c_lumaFilter[yf & 3][t] : c_chromaFilter[yf][t]; }
    if (!xf && !yf)
    {
        for (int e = threadIdx.x; e < w * h; e += blockDim.x)
        {
            const int y = e / w, x = e - y * w;
            const int v = src[(int64_t)y * ss + x];
            if (toShort) dsh[e] = (int16_t)((int16_t)(v << headRoom) - (int16_t)8192);          // filterPixelToShort_c, ipfilter.cpp:40-57
            else dpix[(int64_t)y * dps + x] = (pixel)v;                                         // copy_pp
        }
        return;
    }
    if (!xf || !yf)
    {
        const int64_t step = yf ? ss : 1;
        const int* c = yf ? cy : cx;
        for (int e = threadIdx.x; e < w * h; e += blockDim.x)
        {
            const int y = e / w, x = e - y * w;
            const pixel* q = src + (int64_t)y * ss + x - half * step;
            int sum = 0;
#pragma unroll
            for (int t = 0; t < N; t++) sum += (int)q[t * step] * c[t];
            if (toShort) dsh[e] = (int16_t)((sum + offPS) >> shPS);                              // interp_horiz_ps_c / interp_vert_ps_c
            else
            {
                const int val = (int16_t)((sum + 32) >> 6);                                      // interp_horiz_pp_c / interp_vert_pp_c
                dpix[(int64_t)y * dps + x] = (pixel)(val < 0 ? 0 : (val > maxVal ? maxVal : val));
            }
        }
        return;
    }
    // hps with row extension into immed (pitch w), then vsp (pixel path, predict.cpp:266 luma_hvpp / :322-326) or vss (:297-303)
    const int rows = h + N - 1;
    for (int e = threadIdx.x; e < rows * w; e += blockDim.x)
    {
        const int y = e / w, x = e - y * w;
        const pixel* q = src + (int64_t)(y - half) * ss + x - half;
        int sum = 0;
#pragma unroll
        for (int t = 0; t < N; t++) sum += (int)q[t] * cx[t];
        immed[e] = (int16_t)((sum + offPS) >> shPS);
    }
    __syncthreads();
    const int shSP = 6 + headRoom, offSP = (1 << (shSP - 1)) + (8192 << 6);
    for (int e = threadIdx.x; e < w * h; e += blockDim.x)
    {
        const int y = e / w, x = e - y * w;
        int sum = 0;
#pragma unroll
        for (int t = 0; t < N; t++) sum += (int)immed[(y + t) * w + x] * cy[t];
        if (toShort) dsh[e] = (int16_t)(sum >> 6);                                                // interp_vert_ss_c
        else
        {
            const int val = (int16_t)((sum + offSP) >> shSP);                                     // interp_vert_sp_c
            dpix[(int64_t)y * dps + x] = (pixel)(val < 0 ? 0 : (val > maxVal ? maxVal : val));
        }
    }
    __syncthreads();
}

template<typename pixel>
__global__ void __launch_bounds__(128)
mc_kernel(McArgs a)
{
    __shared__ int16_t sh[2][64 * 64];
    __shared__ int16_t immed[64 * (64 + 7)];
    const x265b200_mc_job job = a.jobs[blockIdx.x];
    const int p = a.planes[blockIdx.y];
    const int hs = p ? a.hshift : 0, vs = p ? a.vshift : 0;
    const int w = job.w >> hs, h = job.h >> vs, x0 = job.puX >> hs, y0 = job.puY >> vs;
    if (w <= 0 || h <= 0) return;
    const int maxVal = (1 << a.depth) - 1, shiftNum = 14 - a.depth;

    // CUData::clipMv
    const int xmax = (a.picW + 8 - job.cuX - 1) << 2, xmin = -((a.maxCU + 8 + job.cuX - 1) << 2);
    const int ymax = (a.picH + 8 - job.cuY - 1) << 2, ymin = -((a.maxCU + 8 + job.cuY - 1) << 2);
    int mvx[2], mvy[2];
#pragma unroll
    for (int l = 0; l < 2; l++) { mvx[l] = min(xmax, max(xmin, job.mv[l][0])); mvy[l] = min(ymax, max(ymin, job.mv[l][1])); }

    const int r0 = job.refIdx[0], r1 = a.isP ? -1 : job.refIdx[1];
    auto W = [&](int l, int r) -> const x265b200_mc_weight* { return a.weights + ((int64_t)l * a.maxRefs + r) * 3; };
    // mode: 0 uni pixel, 1 uni weighted, 2 bi average, 3 bi weighted
    int mode, l0 = 0;
    if (a.isP) mode = (a.wpP && a.weights && W(0, r0)[0].present) ? 1 : 0;
    else
    {
        const bool pw0 = a.wpB && a.weights && r0 >= 0, pw1 = a.wpB && a.weights && r1 >= 0;
        if (r0 >= 0 && r1 >= 0) mode = (pw0 && pw1 && (W(0, r0)[0].present || W(1, r1)[0].present)) ? 3 : 2;
        else if (r0 >= 0) mode = (pw0 && W(0, r0)[0].present) ? 1 : 0;
        else { l0 = 1; mode = (pw1 && W(1, r1)[0].present) ? 1 : 0; }
    }
    pixel* dst = (pixel*)a.pred[p] + (int64_t)y0 * a.predStride[p] + x0;
    const int64_t ds = a.predStride[p];

    const int nl = mode >= 2 ? 2 : 1;
    for (int k = 0; k < nl; k++)
    {
        const int l = mode >= 2 ? k : l0, r = l ? r1 : r0;
        const pixel* plane = (const pixel*)a.refs[((int64_t)l * a.maxRefs + r) * 3 + p];
        int ix, iy, xf, yf;
        if (!p) { ix = mvx[l] >> 2; iy = mvy[l] >> 2; xf = mvx[l] & 3; yf = mvy[l] & 3; }
        else
        {
            const int cx = mvx[l] << (1 - hs), cy = mvy[l] << (1 - vs);                         // predict.cpp:312-313
            ix = cx >> 3; iy = cy >> 3; xf = cx & 7; yf = cy & 7;
        }
        const pixel* src = plane + (int64_t)(y0 + iy) * a.refStride[p] + x0 + ix;
        if (!p) mc_fir<pixel, 8>(src, a.refStride[p], w, h, xf, yf, a.depth, mode != 0, dst, ds, sh[k], immed);
        else    mc_fir<pixel, 4>(src, a.refStride[p], w, h, xf, yf, a.depth, mode != 0, dst, ds, sh[k], immed);
    }
    if (mode == 0) return;
    __syncthreads();
    if (mode == 2)
    {
        const int sn = shiftNum + 1, offset = (1 << (sn - 1)) + 2 * 8192;                        // addAvg, pixel.cpp:842-862
        for (int e = threadIdx.x; e < w * h; e += blockDim.x)
        {
            const int y = e / w, x = e - y * w, v = ((int)sh[0][e] + (int)sh[1][e] + offset) >> sn;
            dst[(int64_t)y * ds + x] = (pixel)(v < 0 ? 0 : (v > maxVal ? maxVal : v));
        }
    }
    else if (mode == 1)
    {
        const x265b200_mc_weight wp = W(l0, l0 ? r1 : r0)[p];                                    // addWeightUni -> weight_sp_c
        const int shift = wp.shift + shiftNum, round = shift ? (1 << (shift - 1)) : 0, offset = wp.o * (1 << (a.depth - 8));
        for (int e = threadIdx.x; e < w * h; e += blockDim.x)
        {
            const int y = e / w, x = e - y * w, v = ((wp.w * ((int)sh[0][e] + 8192) + round) >> shift) + offset;
            dst[(int64_t)y * ds + x] = (pixel)(v < 0 ? 0 : (v > maxVal ? maxVal : v));
        }
    }
    else
    {
        const x265b200_mc_weight w0 = W(0, r0)[p], w1 = W(1, r1)[p];                             // addWeightBi / weightBidir
        const int offset = (w0.o + w1.o) * (1 << (a.depth - 8)), shift = w0.shift + shiftNum + 1, round = shift ? (1 << (shift - 1)) : 0;
        for (int e = threadIdx.x; e < w * h; e += blockDim.x)
        {
            const int y = e / w, x = e - y * w;
            const int v = (w0.w * ((int)sh[0][e] + 8192) + w1.w * ((int)sh[1][e] + 8192) + round + (offset * (1 << (shift - 1)))) >> shift;
            dst[(int64_t)y * ds + x] = (pixel)(v < 0 ? 0 : (v > maxVal ? maxVal : v));
        }
    }
}

int mc_dev(Ctx* ctx, int depth, const x265b200_mc_desc* d, const x265b200_mc_job* jobs, int64_t n, int bLuma, int bChroma)
{
    if (n <= 0) return 0;
    if (!d || !d->refs || !jobs) { set_error("mc: NULL descriptor / reference table / jobs"); return -1; }
    if (d->csp < 0 || d->csp > 3) { set_error("mc: csp %d", d->csp); return -1; }
    if (d->maxRefs <= 0) { set_error("mc: maxRefs %d", d->maxRefs); return -1; }
    if (upload_filters(ctx)) return -1;
    McArgs a;
    a.refs = d->refs; a.jobs = jobs; a.n = n; a.weights = d->weights; a.maxRefs = d->maxRefs;
    a.refStride[0] = d->refStrideY; a.refStride[1] = a.refStride[2] = d->refStrideC;
    a.pred[0] = d->predY; a.pred[1] = d->predCb; a.pred[2] = d->predCr;
    a.predStride[0] = d->predStrideY; a.predStride[1] = a.predStride[2] = d->predStrideC;
    a.hshift = d->csp == 1 || d->csp == 2; a.vshift = d->csp == 1;
    a.depth = depth; a.isP = d->isPSlice; a.wpP = d->weightedPred; a.wpB = d->weightedBiPred;
    a.picW = d->picWidth; a.picH = d->picHeight; a.maxCU = d->maxCUSize;
    a.nPlanes = 0;
    if (bLuma) a.planes[a.nPlanes++] = 0;
    if (bChroma && d->csp != 0) { a.planes[a.nPlanes++] = 1; a.planes[a.nPlanes++] = 2; }
    if (!a.nPlanes) return 0;
    for (int i = 0; i < a.nPlanes; i++) if (!a.pred[a.planes[i]]) { set_error("mc: prediction plane %d is NULL", a.planes[i]); return -1; }
    dim3 grid((unsigned)n, (unsigned)a.nPlanes);
    if (depth > 8) mc_kernel<uint16_t><<<grid, 128, 0, ctx->stream>>>(a);
    else           mc_kernel<uint8_t><<<grid, 128, 0, ctx->stream>>>(a);
    ctx->launches++;
    return check(cudaGetLastError(), "mc kernel launch");
}

} // namespace x265b200

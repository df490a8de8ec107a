// misc_kernels.cu -- the remaining pointer-level entries next to the hot path (SURVEY.md 8f-4 and the "full table" of config 3):
//   propagateCost = estimateCUPropagateCost, fix8Pack / fix8Unpack   (source/common/pixel.cpp:912-956; cuTree, FP double)
//   planecopy_cp / planecopy_sp / planecopy_sp_shl / planecopy_pp_shr  (pixel.cpp:864-910; picture ingest)
//   cu[].ssimDist / cu[].normFact                                      (pixel.cpp:958-994; --ssim-rd)
// The double-precision entries use explicit round-to-nearest multiplies / adds / divides (no FMA contraction), which is what
// the reference's scalar C compiles to, and the x86 "integer indefinite" result (INT_MIN) for out-of-range double -> int
// conversions, so the results stay bit-identical even on degenerate inputs (intraCost == 0).
#include "common.cuh"
#include "x265b200.h"

namespace x265b200 {

namespace {

// cvttsd2si: truncation, 0x80000000 when the value does not fit (or is NaN)
__device__ __forceinline__ int x86_double_to_int(double r)
{
    return (r > -2147483649.0 && r < 2147483648.0) ? (int)r : (int)0x80000000;
}

__global__ void propagate_cost_kernel(int* dst, const uint16_t* propagateIn, const int32_t* intraCosts, const uint16_t* interCosts,
                                      const int32_t* invQscales, double fpsFactor, int64_t len)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    const double fps = __ddiv_rn(fpsFactor, 256.0);
    const int intraCost = intraCosts[i];
    const int interCost = min(intraCost, (int)(interCosts[i] & ((1 << 14) - 1)));                 // LOWRES_COST_MASK
    const double propagateIntra = (double)(int)((uint32_t)intraCost * (uint32_t)invQscales[i]);  // int * int, then converted
    const double propagateAmount = __dadd_rn((double)propagateIn[i], __dmul_rn(propagateIntra, fps));
    const double propagateNum = (double)(intraCost - interCost);
    const double propagateDenom = (double)intraCost;
    dst[i] = x86_double_to_int(__dadd_rn(__ddiv_rn(__dmul_rn(propagateAmount, propagateNum), propagateDenom), 0.5));
}

// ---- Lookahead::estimateCUPropagate (slicetype.cpp:2641-2747): the cuTree propagation step of one (p0, p1, b) ------------------
// Per 8x8 CU of frame b: the amount to propagate (estimateCUPropagateCost, the arithmetic of propagate_cost_kernel above), split by
// the lists the CU used, follows its lowres MVs into the reference frames and is spread bilinearly over the four CUs it lands on.
// The reference adds with CLIP_ADD (saturate at 65535) in raster order; every addend is >= 0, so the final value of an entry is
// min(old + sum of its addends, 65535) whatever the order: the addends are summed with 64-bit atomics, then one pass clamps.
struct CuTreeArgs
{
    const uint16_t* propagateIn;          // frames[b]->propagateCost, or nullptr (non-referenced frame: zeros)
    const int32_t* intraCost; const uint16_t* lowresCosts; const int32_t* invQscale;
    const int32_t* mvs[2];                // lowresMvs[list][listDist[list]] ({x, y} int32 pairs), may be nullptr when the list is unused
    unsigned long long* acc[2];           // zeroed accumulators for frames[p0] / frames[p1]->propagateCost
    int widthInCU, heightInCU, bipredWeight;
    double fpsFactor;
};
__global__ void cutree_propagate_kernel(CuTreeArgs p)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int W = p.widthInCU, Hc = p.heightInCU;
    if (i >= W * Hc) return;
    const int blockx = i % W, blocky = i / W;
    const double fps = __ddiv_rn(p.fpsFactor, 256.0);
    const int intraCost = p.intraCost[i];
    const int interCost = min(intraCost, (int)(p.lowresCosts[i] & ((1 << 14) - 1)));
    const double propagateIntra = (double)(int)((uint32_t)intraCost * (uint32_t)p.invQscale[i]);
    const double propagateAmount = __dadd_rn((double)(p.propagateIn ? p.propagateIn[i] : (uint16_t)0), __dmul_rn(propagateIntra, fps));
    const int amount = x86_double_to_int(__dadd_rn(__ddiv_rn(__dmul_rn(propagateAmount, (double)(intraCost - interCost)), (double)intraCost), 0.5));
    if (amount <= 0) return;                                       // don't propagate for an intra block
    const int listsUsed = p.lowresCosts[i] >> 14;
    for (int list = 0; list < 2; list++)
    {
        if (!((listsUsed >> list) & 1)) continue;
        int listamount = amount;
        if (listsUsed == 3) listamount = (listamount * (list ? 64 - p.bipredWeight : p.bipredWeight) + 32) >> 6;
        unsigned long long* acc = p.acc[list];
        int x = p.mvs[list][2 * i], y = p.mvs[list][2 * i + 1];
        if (!(x | y)) { atomicAdd(&acc[i], (unsigned long long)listamount); continue; }
        const int cux = (x >> 5) + blockx, cuy = (y >> 5) + blocky;
        const int idx0 = cux + cuy * W;
        x &= 31; y &= 31;
        const int w0 = (32 - y) * (32 - x), w1 = (32 - y) * x, w2 = y * (32 - x), w3 = y * x;
        // (the reference's fast path for fully-inside targets and its per-target checks select the same targets)
        if (cux >= 0 && cux < W && cuy >= 0 && cuy < Hc)             atomicAdd(&acc[idx0], (unsigned long long)((listamount * w0 + 512) >> 10));
        if (cux + 1 >= 0 && cux + 1 < W && cuy >= 0 && cuy < Hc)     atomicAdd(&acc[idx0 + 1], (unsigned long long)((listamount * w1 + 512) >> 10));
        if (cux >= 0 && cux < W && cuy + 1 >= 0 && cuy + 1 < Hc)     atomicAdd(&acc[idx0 + W], (unsigned long long)((listamount * w2 + 512) >> 10));
        if (cux + 1 >= 0 && cux + 1 < W && cuy + 1 >= 0 && cuy + 1 < Hc) atomicAdd(&acc[idx0 + W + 1], (unsigned long long)((listamount * w3 + 512) >> 10));
    }
}
__global__ void cutree_clamp_kernel(uint16_t* cost, const unsigned long long* acc, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long v = (unsigned long long)cost[i] + acc[i];
    cost[i] = (uint16_t)(v < 65535ull ? v : 65535ull);
}

__global__ void fix8_pack_kernel(uint16_t* dst, const double* src, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (uint16_t)(int16_t)x86_double_to_int(__dmul_rn(src[i], 256.0));
}
__global__ void fix8_unpack_kernel(double* dst, const uint16_t* src, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __ddiv_rn((double)(int16_t)src[i], 256.0);
}

template<typename pixel>
__global__ void planecopy_kernel(int mode, const void* src, int64_t srcStride, pixel* dst, int64_t dstStride, int width, int height, int shift, int mask)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)width * height) return;
    const int r = (int)(i / width), c = (int)(i - (int64_t)r * width);
    const int64_t so = (int64_t)r * srcStride + c, d = (int64_t)r * dstStride + c;
    switch (mode)
    {
    case X265B200_PC_CP:     dst[d] = (pixel)(((pixel)((const uint8_t*)src)[so]) << shift); break;
    case X265B200_PC_SP:     dst[d] = (pixel)((((const uint16_t*)src)[so] >> shift) & mask); break;
    case X265B200_PC_SP_SHL: dst[d] = (pixel)((((const uint16_t*)src)[so] << shift) & mask); break;
    default:                 dst[d] = (pixel)(((const pixel*)src)[so] >> shift); break;
    }
}

// ssimDist_c<log2TrSize>: ssBlock = sum (fenc - recon)^2 ; ac_k = sum (fenc >> shift)^2 (pixel.cpp:958-981); one warp per block
template<typename pixel>
__global__ void ssim_dist_kernel(const pixel* fenc, int64_t fStride, const pixel* recon, int64_t rStride, const int64_t* offF, const int64_t* offR,
                                 int64_t n, int trSize, int shift, uint64_t* ssBlock, uint64_t* ack)
{
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= n) return;
    const pixel* f = fenc + offF[b]; const pixel* r = recon + offR[b];
    unsigned long long ss = 0, ac = 0;
    for (int e = lane; e < trSize * trSize; e += 32)
    {
        const int y = e / trSize, x = e - y * trSize;
        const int fv = f[y * fStride + x], t = fv - (int)r[y * rStride + x];
        ss += (unsigned long long)(long long)(t * t);
        const uint32_t u = (uint32_t)fv >> shift;
        ac += (unsigned long long)(u * u);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        ss += ((unsigned long long)__shfl_xor_sync(0xffffffffu, (unsigned)(ss >> 32), o) << 32) | __shfl_xor_sync(0xffffffffu, (unsigned)ss, o);
        ac += ((unsigned long long)__shfl_xor_sync(0xffffffffu, (unsigned)(ac >> 32), o) << 32) | __shfl_xor_sync(0xffffffffu, (unsigned)ac, o);
    }
    if (lane == 0) { ssBlock[b] = ss; ack[b] = ac; }
}

// normFact_c: z_k = sum (src >> shift)^2 over a contiguous blockSize x blockSize block (pixel.cpp:983-994)
template<typename pixel>
__global__ void norm_fact_kernel(const pixel* src, const int64_t* off, int64_t n, int blockSize, int shift, uint64_t* zk)
{
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= n) return;
    const pixel* s = src + off[b];
    unsigned long long z = 0;
    for (int e = lane; e < blockSize * blockSize; e += 32) { const uint32_t u = (uint32_t)s[e] >> shift; z += (unsigned long long)(u * u); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        z += ((unsigned long long)__shfl_xor_sync(0xffffffffu, (unsigned)(z >> 32), o) << 32) | __shfl_xor_sync(0xffffffffu, (unsigned)z, o);
    if (lane == 0) zk[b] = z;
}

// ssim_4x4x2_core (pixel.cpp:631-658): the four sums of two horizontally adjacent 4x4 blocks; one thread per job
template<typename pixel>
__global__ void ssim_core_kernel(const pixel* p1, int64_t s1, const pixel* p2, int64_t s2, const int64_t* off1, const int64_t* off2, int64_t n, int32_t* sums)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * 2) return;
    const int64_t j = t >> 1; const int z = (int)(t & 1);
    const pixel* a = p1 + off1[j] + 4 * z; const pixel* b = p2 + off2[j] + 4 * z;
    uint32_t v1 = 0, v2 = 0, ss = 0, s12 = 0;
    for (int y = 0; y < 4; y++)
        for (int x = 0; x < 4; x++)
        {
            const uint32_t pa = a[x + y * s1], pb = b[x + y * s2];
            v1 += pa; v2 += pb; ss += pa * pa; ss += pb * pb; s12 += pa * pb;
        }
    int32_t* o = sums + j * 8 + z * 4;
    o[0] = (int32_t)v1; o[1] = (int32_t)v2; o[2] = (int32_t)ss; o[3] = (int32_t)s12;
}

// ssim_end_1 / ssim_end_4 (pixel.cpp:660-702): integer terms for 8-bit, float terms for HIGH_BIT_DEPTH, every float
// operation individually rounded (no contraction) in the reference's order
__device__ float ssim_end_1_dev(int s1, int s2, int ss, int s12, int depth)
{
    const double pmax = (double)((1 << depth) - 1);
    if (depth == 8)
    {
        const int c1 = (int)(.01 * .01 * pmax * pmax * 64 + .5), c2 = (int)(.03 * .03 * pmax * pmax * 64 * 63 + .5);
        const int vars = (int)((uint32_t)ss * 64u - (uint32_t)s1 * (uint32_t)s1 - (uint32_t)s2 * (uint32_t)s2);
        const int covar = (int)((uint32_t)s12 * 64u - (uint32_t)s1 * (uint32_t)s2);
        const float num = __fmul_rn((float)(int)(2u * (uint32_t)s1 * (uint32_t)s2 + (uint32_t)c1), (float)(int)(2u * (uint32_t)covar + (uint32_t)c2));
        const float den = __fmul_rn((float)(int)((uint32_t)s1 * (uint32_t)s1 + (uint32_t)s2 * (uint32_t)s2 + (uint32_t)c1), (float)(int)((uint32_t)vars + (uint32_t)c2));
        return __fdiv_rn(num, den);
    }
    const float c1 = (float)(.01 * .01 * pmax * pmax * 64), c2 = (float)(.03 * .03 * pmax * pmax * 64 * 63);
    const float fs1 = (float)s1, fs2 = (float)s2, fss = (float)ss, fs12 = (float)s12;
    const float vars = __fsub_rn(__fsub_rn(__fmul_rn(fss, 64.f), __fmul_rn(fs1, fs1)), __fmul_rn(fs2, fs2));
    const float covar = __fsub_rn(__fmul_rn(fs12, 64.f), __fmul_rn(fs1, fs2));
    const float num = __fmul_rn(__fadd_rn(__fmul_rn(__fmul_rn(2.f, fs1), fs2), c1), __fadd_rn(__fmul_rn(2.f, covar), c2));
    const float den = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(fs1, fs1), __fmul_rn(fs2, fs2)), c1), __fadd_rn(vars, c2));
    return __fdiv_rn(num, den);
}
__global__ void ssim_end4_kernel(const int32_t* sum0, const int32_t* sum1, const int32_t* widths, int64_t n, int depth, float* out)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int32_t* a = sum0 + j * 20; const int32_t* b = sum1 + j * 20;
    float ssim = 0.f;
    for (int i = 0; i < widths[j]; i++)
        ssim = __fadd_rn(ssim, ssim_end_1_dev(a[i * 4 + 0] + a[i * 4 + 4] + b[i * 4 + 0] + b[i * 4 + 4], a[i * 4 + 1] + a[i * 4 + 5] + b[i * 4 + 1] + b[i * 4 + 5],
                                              a[i * 4 + 2] + a[i * 4 + 6] + b[i * 4 + 2] + b[i * 4 + 6], a[i * 4 + 3] + a[i * 4 + 7] + b[i * 4 + 3] + b[i * 4 + 7], depth));
    out[j] = ssim;
}

// planeClipAndMax_c (pixel.cpp:996-1016, HIGH_BIT_DEPTH builds): clip the plane in place, return max and sum
template<typename pixel>
__global__ void plane_clip_max_kernel(pixel* src, int64_t stride, int width, int height, int minPix, int maxPix, unsigned long long* sum, unsigned* maxOut)
{
    unsigned long long s = 0; unsigned m = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)width * height; i += (int64_t)gridDim.x * blockDim.x)
    {
        const int r = (int)(i / width), c = (int)(i - (int64_t)r * width);
        int v = src[(int64_t)r * stride + c];
        v = v < minPix ? minPix : (v > maxPix ? maxPix : v);                     // x265_clip3(min, max, v)
        src[(int64_t)r * stride + c] = (pixel)v;
        s += (unsigned)v; m = max(m, (unsigned)v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        s += ((unsigned long long)__shfl_xor_sync(0xffffffffu, (unsigned)(s >> 32), o) << 32) | __shfl_xor_sync(0xffffffffu, (unsigned)s, o);
        m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(sum, s); atomicMax(maxOut, m); }
}

} // namespace

int scratch_dev(Ctx* ctx, int slot, size_t bytes, void** out);
int extend_border_dev(Ctx* ctx, int depth, void* origin, int64_t stride, int width, int height, int marginX, int marginY);

// ---- LookaheadTLD::acEnergyCu for every quantisation-group block of a frame (slicetype.cpp:49-86, :252-275) ----------------------
// One warp per block: pixel_var (sum | sqr << 32, pixel.cpp:703-720) of the luma block and of the two chroma blocks, each folded as
// acEnergyVar folds it (ssd - (sum * sum >> shift), 32-bit) and added; the per-plane sums / ssds also go to the frame's wp_sum /
// wp_ssd (the inputs of weightsAnalyse).  The float qp-offset arithmetic of calcAdaptiveQuantFrame stays on the host.
template<typename pixel>
__global__ void aq_energy_kernel(const pixel* y, int64_t strideY, const pixel* cb, const pixel* cr, int64_t strideC, int hShift, int vShift,
                                 int blocksX, int blocksY, int qg, int csp, uint32_t* energy, unsigned long long* wp /* [6]: sum Y,Cb,Cr, ssd Y,Cb,Cr */)
{
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= blocksX * blocksY) return;
    const int bx = (b % blocksX) * qg, by = (b / blocksX) * qg;
    uint32_t total = 0;
    for (int plane = 0; plane < (csp ? 3 : 1); plane++)
    {
        const bool sub = plane && csp != 3;                       // 4:2:0 / 4:2:2 chroma blocks are half size (copied to 4x4 / 8x8, :62-77)
        const int size = sub ? qg >> 1 : qg;
        const pixel* p = plane == 0 ? y + bx + (int64_t)by * strideY
                                    : (plane == 1 ? cb : cr) + (bx >> hShift) + (int64_t)(by >> vShift) * strideC;
        const int64_t st = plane ? strideC : strideY;
        uint32_t sum = 0, sqr = 0;
        for (int e = lane; e < size * size; e += 32)
        {
            const int yy = e / size, xx = e - yy * size;
            const uint32_t v = p[(int64_t)yy * st + xx];
            sum += v; sqr += v * v;
        }
        sum = (uint32_t)warp_sum((int)sum); sqr = (uint32_t)warp_sum((int)sqr);
        const int shift = size == 16 ? 8 : (size == 8 ? 6 : 4);
        total += sqr - (uint32_t)(((unsigned long long)sum * sum) >> shift);
        if (lane == 0) { atomicAdd(&wp[plane], (unsigned long long)sum); atomicAdd(&wp[3 + plane], (unsigned long long)sqr); }
    }
    if (lane == 0) energy[b] = total;
}

int aq_energy_dev(Ctx* ctx, int depth, int csp, int qgSize, const void* y, int64_t strideY, const void* cb, const void* cr, int64_t strideC,
                  int picWidth, int picHeight, uint32_t* energy, uint64_t* wpSumSsd)
{
    if (qgSize != 8 && qgSize != 16) { set_error("aq_energy: qgSize %d (8 or 16)", qgSize); return -1; }
    if (csp < 0 || csp > 3 || (csp && (!cb || !cr))) { set_error("aq_energy: csp %d / chroma planes", csp); return -1; }
    const int bxn = (picWidth + qgSize - 1) / qgSize, byn = (picHeight + qgSize - 1) / qgSize;
    if (bxn <= 0 || byn <= 0) return 0;
    X265B200_CHECK(cudaMemsetAsync(wpSumSsd, 0, 6 * sizeof(uint64_t), ctx->stream));
    const int hs = (csp == 1 || csp == 2) ? 1 : 0, vs = csp == 1 ? 1 : 0;
    const int n = bxn * byn;
    if (depth > 8) aq_energy_kernel<uint16_t><<<(n + 7) / 8, 256, 0, ctx->stream>>>((const uint16_t*)y, strideY, (const uint16_t*)cb, (const uint16_t*)cr, strideC, hs, vs, bxn, byn, qgSize, csp, energy, (unsigned long long*)wpSumSsd);
    else           aq_energy_kernel<uint8_t><<<(n + 7) / 8, 256, 0, ctx->stream>>>((const uint8_t*)y, strideY, (const uint8_t*)cb, (const uint8_t*)cr, strideC, hs, vs, bxn, byn, qgSize, csp, energy, (unsigned long long*)wpSumSsd);
    ctx->launches++;
    return check(cudaGetLastError(), "aq_energy launch");
}

// ---- MotionReference::applyWeight for a whole plane (reference.cpp:119-185) ------------------------------------------------------
template<typename pixel>
__global__ void apply_weight_kernel(const pixel* src, pixel* dst, int64_t stride, int width, int height, int w0, int round, int shift, int offset, int corr, int maxVal)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)width * height) return;
    const int r = (int)(i / width), c = (int)(i - (int64_t)r * width);
    const int v = ((w0 * ((int)src[(int64_t)r * stride + c] << corr) + round) >> shift) + offset;          // weight_pp_c, pixel.cpp:516-543
    dst[(int64_t)r * stride + c] = (pixel)(v < 0 ? 0 : (v > maxVal ? maxVal : v));
}

int apply_weight_dev(Ctx* ctx, int depth, const void* srcOrigin, void* dstOrigin, int64_t stride, int width, int height, int marginX, int marginY,
                     int inputWeight, int inputOffset, int log2WeightDenom)
{
    if (width <= 0 || height <= 0) return 0;
    // MotionReference::init (:100-103) and applyWeight (:160-161): offset scaled to the bit depth, rounding and shift in the
    // 14-bit intermediate domain of weight_pp
    const int corr = 14 - depth;
    const int offset = inputOffset * (1 << (depth - 8));
    const int round = (log2WeightDenom ? 1 << (log2WeightDenom - 1) : 0) << corr, shift = log2WeightDenom + corr;
    const int64_t n = (int64_t)width * height;
    if (depth > 8) apply_weight_kernel<uint16_t><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const uint16_t*)srcOrigin, (uint16_t*)dstOrigin, stride, width, height, inputWeight, round, shift, offset, corr, (1 << depth) - 1);
    else           apply_weight_kernel<uint8_t><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const uint8_t*)srcOrigin, (uint8_t*)dstOrigin, stride, width, height, inputWeight, round, shift, offset, corr, 255);
    ctx->launches++;
    if (check(cudaGetLastError(), "apply_weight launch")) return -1;
    return extend_border_dev(ctx, depth, dstOrigin, stride, width, height, marginX, marginY);     // extendRowBorder + the rows above / below (:163-182)
}

int cutree_propagate_dev(Ctx* ctx, int widthInCU, int heightInCU, const uint16_t* propagateCostB, const int32_t* intraCost, const uint16_t* lowresCosts,
                         const int32_t* invQscale, const int32_t* mvs0, const int32_t* mvs1, uint16_t* refCost0, uint16_t* refCost1,
                         int bipredWeight, double fpsFactor)
{
    const int n = widthInCU * heightInCU;
    if (n <= 0) return 0;
    if (!intraCost || !lowresCosts || !invQscale || !refCost0) { set_error("cutree_propagate: null array"); return -1; }
    void* scr = nullptr;
    if (scratch_dev(ctx, 6, (size_t)n * 16, &scr)) return -1;
    X265B200_CHECK(cudaMemsetAsync(scr, 0, (size_t)n * 16, ctx->stream));
    CuTreeArgs a;
    a.propagateIn = propagateCostB; a.intraCost = intraCost; a.lowresCosts = lowresCosts; a.invQscale = invQscale;
    a.mvs[0] = mvs0; a.mvs[1] = mvs1; a.acc[0] = (unsigned long long*)scr; a.acc[1] = (unsigned long long*)scr + n;
    a.widthInCU = widthInCU; a.heightInCU = heightInCU; a.bipredWeight = bipredWeight; a.fpsFactor = fpsFactor;
    if (!mvs0) { set_error("cutree_propagate: list 0 MVs missing"); return -1; }
    if (!mvs1) a.mvs[1] = mvs0;                          // never dereferenced: list 1 is only used by CUs of B frames
    cutree_propagate_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(a);
    cutree_clamp_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(refCost0, a.acc[0], n);
    if (refCost1 && refCost1 != refCost0) cutree_clamp_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(refCost1, a.acc[1], n);
    ctx->launches += 2 + (refCost1 && refCost1 != refCost0);
    return check(cudaGetLastError(), "cutree_propagate launch");
}

int propagate_cost_dev(Ctx* ctx, int* dst, const uint16_t* propagateIn, const int32_t* intraCosts, const uint16_t* interCosts,
                       const int32_t* invQscales, double fpsFactor, int64_t len)
{
    if (len <= 0) return 0;
    propagate_cost_kernel<<<(unsigned)((len + 255) / 256), 256, 0, ctx->stream>>>(dst, propagateIn, intraCosts, interCosts, invQscales, fpsFactor, len);
    ctx->launches++;
    return check(cudaGetLastError(), "propagate_cost kernel launch");
}
int fix8_dev(Ctx* ctx, int unpack, void* dst, const void* src, int64_t n)
{
    if (n <= 0) return 0;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (unpack) fix8_unpack_kernel<<<blocks, 256, 0, ctx->stream>>>((double*)dst, (const uint16_t*)src, n);
    else        fix8_pack_kernel<<<blocks, 256, 0, ctx->stream>>>((uint16_t*)dst, (const double*)src, n);
    ctx->launches++;
    return check(cudaGetLastError(), "fix8 kernel launch");
}
int planecopy_dev(Ctx* ctx, int mode, int depth, const void* src, int64_t srcStride, void* dst, int64_t dstStride, int width, int height, int shift, int mask)
{
    if (width <= 0 || height <= 0) return 0;
    if (mode < X265B200_PC_CP || mode > X265B200_PC_PP_SHR) { set_error("planecopy: mode %d", mode); return -1; }
    const unsigned blocks = (unsigned)(((int64_t)width * height + 255) / 256);
    if (depth > 8) planecopy_kernel<uint16_t><<<blocks, 256, 0, ctx->stream>>>(mode, src, srcStride, (uint16_t*)dst, dstStride, width, height, shift, mask);
    else           planecopy_kernel<uint8_t><<<blocks, 256, 0, ctx->stream>>>(mode, src, srcStride, (uint8_t*)dst, dstStride, width, height, shift, mask);
    ctx->launches++;
    return check(cudaGetLastError(), "planecopy kernel launch");
}
int ssim_dist_dev(Ctx* ctx, int depth, int log2TrSize, const void* fenc, int64_t fStride, const void* recon, int64_t rStride,
                  const int64_t* offF, const int64_t* offR, int64_t n, int shift, uint64_t* ssBlock, uint64_t* ack)
{
    if (n <= 0) return 0;
    if (log2TrSize < 2 || log2TrSize > 6) { set_error("ssim_dist: log2TrSize %d", log2TrSize); return -1; }
    const unsigned blocks = (unsigned)((n + 3) / 4);
    if (depth > 8) ssim_dist_kernel<uint16_t><<<blocks, 128, 0, ctx->stream>>>((const uint16_t*)fenc, fStride, (const uint16_t*)recon, rStride, offF, offR, n, 1 << log2TrSize, shift, ssBlock, ack);
    else           ssim_dist_kernel<uint8_t><<<blocks, 128, 0, ctx->stream>>>((const uint8_t*)fenc, fStride, (const uint8_t*)recon, rStride, offF, offR, n, 1 << log2TrSize, shift, ssBlock, ack);
    ctx->launches++;
    return check(cudaGetLastError(), "ssim_dist kernel launch");
}
int norm_fact_dev(Ctx* ctx, int depth, const void* src, const int64_t* off, int64_t n, int blockSize, int shift, uint64_t* zk)
{
    if (n <= 0) return 0;
    const unsigned blocks = (unsigned)((n + 3) / 4);
    if (depth > 8) norm_fact_kernel<uint16_t><<<blocks, 128, 0, ctx->stream>>>((const uint16_t*)src, off, n, blockSize, shift, zk);
    else           norm_fact_kernel<uint8_t><<<blocks, 128, 0, ctx->stream>>>((const uint8_t*)src, off, n, blockSize, shift, zk);
    ctx->launches++;
    return check(cudaGetLastError(), "norm_fact kernel launch");
}

int ssim_core_dev(Ctx* ctx, int depth, const void* p1, int64_t s1, const void* p2, int64_t s2, const int64_t* off1, const int64_t* off2, int64_t n, int32_t* sums)
{
    if (n <= 0) return 0;
    const unsigned blocks = (unsigned)((n * 2 + 127) / 128);
    if (depth > 8) ssim_core_kernel<uint16_t><<<blocks, 128, 0, ctx->stream>>>((const uint16_t*)p1, s1, (const uint16_t*)p2, s2, off1, off2, n, sums);
    else           ssim_core_kernel<uint8_t><<<blocks, 128, 0, ctx->stream>>>((const uint8_t*)p1, s1, (const uint8_t*)p2, s2, off1, off2, n, sums);
    ctx->launches++;
    return check(cudaGetLastError(), "ssim_core kernel launch");
}
int ssim_end4_dev(Ctx* ctx, int depth, const int32_t* sum0, const int32_t* sum1, const int32_t* widths, int64_t n, float* out)
{
    if (n <= 0) return 0;
    ssim_end4_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(sum0, sum1, widths, n, depth, out);
    ctx->launches++;
    return check(cudaGetLastError(), "ssim_end4 kernel launch");
}
int plane_clip_max_dev(Ctx* ctx, int depth, void* src, int64_t stride, int width, int height, int minPix, int maxPix, uint64_t* outsum, uint32_t* outmax)
{
    X265B200_CHECK(cudaMemsetAsync(outsum, 0, 8, ctx->stream));
    X265B200_CHECK(cudaMemsetAsync(outmax, 0, 4, ctx->stream));
    if (width <= 0 || height <= 0) return 0;
    int64_t total = (int64_t)width * height;
    unsigned blocks = (unsigned)((total + 255) / 256);
    if (blocks > (unsigned)ctx->smCount * 8) blocks = (unsigned)ctx->smCount * 8;
    if (depth > 8) plane_clip_max_kernel<uint16_t><<<blocks, 256, 0, ctx->stream>>>((uint16_t*)src, stride, width, height, minPix, maxPix, (unsigned long long*)outsum, outmax);
    else           plane_clip_max_kernel<uint8_t><<<blocks, 256, 0, ctx->stream>>>((uint8_t*)src, stride, width, height, minPix, maxPix, (unsigned long long*)outsum, outmax);
    ctx->launches++;
    return check(cudaGetLastError(), "plane_clip_max kernel launch");
}

} // namespace x265b200

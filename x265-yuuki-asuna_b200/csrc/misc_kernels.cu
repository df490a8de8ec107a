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

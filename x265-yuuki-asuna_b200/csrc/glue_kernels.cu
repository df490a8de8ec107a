// glue_kernels.cu -- the element-wise "glue" entries of EncoderPrimitives (SURVEY.md 8a rows a-7, a-10, a-13, a-17;
// class G of the slot checklist): block copies and fills, cpy2Dto1D/1Dto2D shifts, residual / reconstruction,
// bi-prediction averages, explicit weighted prediction, transpose, plus the small reductions that sit next to them
// (variance, psy-cost energy, copy_cnt, denoiseDct) and the low-pass DCT front/back ends.
//
// One generic kernel does every element-wise op: a thread produces one output element of one job
// (job = one w x h block); all of them are HBM-bound copies with <= 4 integer ops per element.
// Reference semantics (all integer, bit-exact): source/common/pixel.cpp:393-557, :703-862; dct.cpp:728-755;
// lowpassdct.cpp:33-111.
#include "common.cuh"
#include "x265b200.h"

namespace x265b200 {

struct GlueArgs
{
    void* dst; int64_t dstStride;
    const void* src0; int64_t src0Stride;
    const void* src1; int64_t src1Stride;
    const x265b200_glue_job* jobs; int64_t n;
    int w, h, depth, p0, p1, p2, p3;
};

__device__ __forceinline__ int clip_px(int v, int maxVal) { return v < 0 ? 0 : (v > maxVal ? maxVal : v); }

template<typename pixel, int OP>
__global__ void __launch_bounds__(256) glue_kernel(GlueArgs a)
{
    const int64_t per = (int64_t)a.w * a.h;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= per * a.n) return;
    const int64_t j = idx / per;
    const int e = (int)(idx - j * per);
    const int y = e / a.w, x = e - y * a.w;
    const x265b200_glue_job job = a.jobs[j];
    const int maxVal = (1 << a.depth) - 1;
    const int64_t d = job.dstOff + (int64_t)y * a.dstStride + x;
    const int64_t s0 = job.src0Off + (int64_t)y * a.src0Stride + x;
    const int64_t s1 = job.src1Off + (int64_t)y * a.src1Stride + x;
    pixel* dp = (pixel*)a.dst; int16_t* ds = (int16_t*)a.dst;
    const pixel* p0 = (const pixel*)a.src0; const int16_t* q0 = (const int16_t*)a.src0;
    const pixel* p1 = (const pixel*)a.src1; const int16_t* q1 = (const int16_t*)a.src1;
    switch (OP)
    {
    case X265B200_GL_COPY_PP: dp[d] = p0[s0]; break;                                        // blockcopy_pp_c  pixel.cpp:759
    case X265B200_GL_COPY_SS: ds[d] = q0[s0]; break;                                        // blockcopy_ss_c  :772
    case X265B200_GL_COPY_SP: dp[d] = (pixel)q0[s0]; break;                                 // blockcopy_sp_c  :785
    case X265B200_GL_COPY_PS: ds[d] = (int16_t)p0[s0]; break;                               // blockcopy_ps_c  :801
    case X265B200_GL_FILL_S:  ds[d] = (int16_t)a.p0; break;                                 // blockfill_s_c   :393
    case X265B200_GL_CPY2DTO1D_SHL:                                                         // :401 (dst contiguous: dstStride = w)
    case X265B200_GL_CPY1DTO2D_SHL: ds[d] = (int16_t)((int)q0[s0] << a.p0); break;          // :436 (src contiguous)
    case X265B200_GL_CPY2DTO1D_SHR:                                                         // :418
    case X265B200_GL_CPY1DTO2D_SHR: ds[d] = (int16_t)(((int)q0[s0] + (int)(int16_t)(1 << (a.p0 - 1))) >> a.p0); break;   // :453
    case X265B200_GL_SUB_PS:  ds[d] = (int16_t)((int)p0[s0] - (int)p1[s1]); break;          // pixel_sub_ps_c :814, getResidual :471
    case X265B200_GL_ADD_PS:  dp[d] = (pixel)clip_px((int)p0[s0] + (int)q1[s1], maxVal); break;   // pixel_add_ps_c :828
    case X265B200_GL_ADDAVG:                                                                // addAvg :842
    {
        const int shiftNum = 14 + 1 - a.depth, offset = (1 << (shiftNum - 1)) + 2 * 8192;
        dp[d] = (pixel)clip_px(((int)q0[s0] + (int)q1[s1] + offset) >> shiftNum, maxVal);
        break;
    }
    case X265B200_GL_PIXELAVG_PP: dp[d] = (pixel)(((int)p0[s0] + (int)p1[s1] + 1) >> 1); break;   // pixelavg_pp :545
    case X265B200_GL_TRANSPOSE:                                                             // transpose :485: dst[k*N+l] = src[l*stride+k]
        dp[job.dstOff + (int64_t)y * a.w + x] = p0[job.src0Off + (int64_t)x * a.src0Stride + y];
        break;
    case X265B200_GL_WEIGHT_PP:                                                             // weight_pp_c :516
    {
        const int val = (int16_t)((int)p0[s0] << (14 - a.depth));
        dp[d] = (pixel)clip_px(((a.p0 * val + a.p1) >> a.p2) + a.p3, maxVal);
        break;
    }
    case X265B200_GL_WEIGHT_SP:                                                             // weight_sp_c :493
        dp[d] = (pixel)clip_px(((a.p0 * ((int)q0[s0] + 8192) + a.p1) >> a.p2) + a.p3, maxVal);
        break;
    case X265B200_GL_SCALE1D_128TO64:                                                       // scale1D_128to64 :559 (w = 128, h = 1)
    {
        const pixel* s = p0 + job.src0Off + (x >> 6) * 128 + ((x & 63) << 1);
        dp[job.dstOff + x] = (pixel)(((int)s[0] + (int)s[1] + 1) >> 1);
        break;
    }
    case X265B200_GL_SCALE2D_64TO32:                                                        // scale2D_64to32 :585 (w = h = 32)
    {
        const pixel* s = p0 + job.src0Off + (int64_t)(2 * y) * a.src0Stride + 2 * x;
        dp[job.dstOff + y * 32 + x] = (pixel)(((int)s[0] + (int)s[1] + (int)s[a.src0Stride] + (int)s[a.src0Stride + 1] + 2) >> 2);
        break;
    }
    }
}

template<typename pixel>
static int launch_glue(Ctx* ctx, int op, const GlueArgs& a)
{
    const int64_t total = (int64_t)a.w * a.h * a.n;
    const unsigned blocks = (unsigned)((total + 255) / 256);
#define GL_CASE(OPC) case OPC: glue_kernel<pixel, OPC><<<blocks, 256, 0, ctx->stream>>>(a); break;
    switch (op)
    {
    GL_CASE(X265B200_GL_COPY_PP) GL_CASE(X265B200_GL_COPY_SS) GL_CASE(X265B200_GL_COPY_SP) GL_CASE(X265B200_GL_COPY_PS)
    GL_CASE(X265B200_GL_FILL_S) GL_CASE(X265B200_GL_CPY2DTO1D_SHL) GL_CASE(X265B200_GL_CPY2DTO1D_SHR)
    GL_CASE(X265B200_GL_CPY1DTO2D_SHL) GL_CASE(X265B200_GL_CPY1DTO2D_SHR) GL_CASE(X265B200_GL_SUB_PS) GL_CASE(X265B200_GL_ADD_PS)
    GL_CASE(X265B200_GL_ADDAVG) GL_CASE(X265B200_GL_PIXELAVG_PP) GL_CASE(X265B200_GL_TRANSPOSE)
    GL_CASE(X265B200_GL_WEIGHT_PP) GL_CASE(X265B200_GL_WEIGHT_SP) GL_CASE(X265B200_GL_SCALE1D_128TO64) GL_CASE(X265B200_GL_SCALE2D_64TO32)
    default: set_error("glue: unknown op %d", op); return -1;
    }
#undef GL_CASE
    ctx->launches++;
    return check(cudaGetLastError(), "glue kernel launch");
}

int glue_dev(Ctx* ctx, int op, int depth, int w, int h, void* dst, int64_t dstStride, const void* src0, int64_t src0Stride,
             const void* src1, int64_t src1Stride, const x265b200_glue_job* jobs, int64_t n, int p0, int p1, int p2, int p3)
{
    if (n <= 0 || w <= 0 || h <= 0) return 0;
    if ((op == X265B200_GL_CPY2DTO1D_SHR || op == X265B200_GL_CPY1DTO2D_SHR) && p0 < 1) { set_error("glue: shr needs shift > 0"); return -1; }
    if (op == X265B200_GL_SCALE1D_128TO64) { w = 128; h = 1; }
    if (op == X265B200_GL_SCALE2D_64TO32) { w = 32; h = 32; }
    GlueArgs a;
    a.dst = dst; a.dstStride = dstStride; a.src0 = src0; a.src0Stride = src0Stride; a.src1 = src1; a.src1Stride = src1Stride;
    a.jobs = jobs; a.n = n; a.w = w; a.h = h; a.depth = depth; a.p0 = p0; a.p1 = p1; a.p2 = p2; a.p3 = p3;
    return depth > 8 ? launch_glue<uint16_t>(ctx, op, a) : launch_glue<uint8_t>(ctx, op, a);
}

// ---- reductions -------------------------------------------------------------------------------------------------
// pixel_var<size> (pixel.cpp:703-720): sum | (uint64)sqr << 32 with 32-bit wrapping accumulators; one warp per block
template<typename pixel>
__global__ void var_kernel(const pixel* src, int64_t stride, const int64_t* off, int64_t n, int size, uint64_t* out)
{
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= n) return;
    const int lane = threadIdx.x & 31;
    const pixel* p = src + off[b];
    uint32_t sum = 0, sqr = 0;
    for (int e = lane; e < size * size; e += 32)
    {
        const int y = e / size, x = e - y * size;
        const uint32_t v = p[(int64_t)y * stride + x];
        sum += v; sqr += v * v;
    }
    sum = (uint32_t)warp_sum((int)sum); sqr = (uint32_t)warp_sum((int)sqr);
    if (lane == 0) out[b] = (uint64_t)sum + ((uint64_t)sqr << 32);
}

// sa8d_8x8(block, zero) = (sum |H8 X H8^T| + 2) >> 2 and sad<8,8>(block, zero) = sum X (pixel.cpp:299-341, :40-55)
template<typename pixel>
__device__ int ac_energy_8x8(const pixel* p, int64_t stride)
{
    int m[8][8];
    int sad = 0;
#pragma unroll
    for (int y = 0; y < 8; y++)
#pragma unroll
        for (int x = 0; x < 8; x++) { m[y][x] = p[(int64_t)y * stride + x]; sad += m[y][x]; }
#pragma unroll
    for (int pass = 0; pass < 2; pass++)
#pragma unroll
        for (int i = 0; i < 8; i++)
        {
            int v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = pass ? m[k][i] : m[i][k];
#pragma unroll
            for (int s = 1; s < 8; s <<= 1)
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if (!(k & s)) { int a = v[k], b = v[k | s]; v[k] = a + b; v[k | s] = a - b; }
#pragma unroll
            for (int k = 0; k < 8; k++) { if (pass) m[k][i] = v[k]; else m[i][k] = v[k]; }
        }
    int sum = 0;
#pragma unroll
    for (int y = 0; y < 8; y++)
#pragma unroll
        for (int x = 0; x < 8; x++) sum += abs(m[y][x]);
    return ((sum + 2) >> 2) - (sad >> 2);
}
// satd_4x4(block, zero) - (sad<4,4> >> 2)   (pixel.cpp:190-208)
template<typename pixel>
__device__ int ac_energy_4x4(const pixel* p, int64_t stride)
{
    int m[4][4], sad = 0;
#pragma unroll
    for (int y = 0; y < 4; y++)
#pragma unroll
        for (int x = 0; x < 4; x++) { m[y][x] = p[(int64_t)y * stride + x]; sad += m[y][x]; }
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        int t0 = m[i][0] + m[i][1], t1 = m[i][0] - m[i][1], t2 = m[i][2] + m[i][3], t3 = m[i][2] - m[i][3];
        m[i][0] = t0 + t2; m[i][2] = t0 - t2; m[i][1] = t1 + t3; m[i][3] = t1 - t3;
    }
    int sum = 0;
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        int t0 = m[0][i] + m[1][i], t1 = m[0][i] - m[1][i], t2 = m[2][i] + m[3][i], t3 = m[2][i] - m[3][i];
        sum += abs(t0 + t2) + abs(t0 - t2) + abs(t1 + t3) + abs(t1 - t3);
    }
    return (sum >> 1) - (sad >> 2);
}

// psyCost_pp<size> (pixel.cpp:726-757): sum over 8x8 sub-blocks of |AC energy(source) - AC energy(recon)|
template<typename pixel>
__global__ void psycost_kernel(const pixel* src, int64_t sstride, const pixel* rec, int64_t rstride, const int64_t* offS, const int64_t* offR,
                               int64_t n, int dim, int32_t* out)
{
    const int sub = dim >= 8 ? (dim >> 3) * (dim >> 3) : 1;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * sub) return;
    const int64_t b = idx / sub;
    const int k = (int)(idx - b * sub);
    int e;
    if (dim >= 8)
    {
        const int per = dim >> 3, i = (k / per) * 8, j = (k % per) * 8;
        e = abs(ac_energy_8x8<pixel>(src + offS[b] + (int64_t)i * sstride + j, sstride) - ac_energy_8x8<pixel>(rec + offR[b] + (int64_t)i * rstride + j, rstride));
    }
    else
        e = abs(ac_energy_4x4<pixel>(src + offS[b], sstride) - ac_energy_4x4<pixel>(rec + offR[b], rstride));
    atomicAdd(&out[b], e);
}

// copy_count<trSize> (dct.cpp:728-743): coeff (contiguous) = residual block, returns the number of non-zeros
__global__ void copy_cnt_kernel(int16_t* coeff, const int16_t* resi, int64_t stride, const int64_t* off, int64_t n, int size, uint32_t* numSig)
{
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= n) return;
    const int lane = threadIdx.x & 31;
    int cnt = 0;
    for (int e = lane; e < size * size; e += 32)
    {
        const int y = e / size, x = e - y * size;
        const int16_t v = resi[off[b] + (int64_t)y * stride + x];
        coeff[b * size * size + e] = v;
        cnt += v != 0;
    }
    cnt = warp_sum(cnt);
    if (lane == 0) numSig[b] = (uint32_t)cnt;
}

// denoiseDct_c (dct.cpp:745-755) over n TUs that share resSum[] / offset[]: the per-position sums commute, so the
// batch accumulates them with atomics (uint32 wrap-around like the reference's +=)
__global__ void denoise_kernel(int16_t* coef, uint32_t* resSum, const uint16_t* offset, int numCoeff, int64_t n)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * numCoeff) return;
    const int i = (int)(idx % numCoeff);
    int level = coef[idx];
    const int sign = level >> 31;
    level = (level + sign) ^ sign;
    atomicAdd(&resSum[i], (uint32_t)level);
    level -= offset[i];
    coef[idx] = (int16_t)(level < 0 ? 0 : (level ^ sign) - sign);
}

// lowPassDct8/16/32_c (lowpassdct.cpp:33-111) front end: 2x2 sums (int16 wrap like the reference's `int16_t sum`),
// averages to a contiguous (N/2)^2 block, block total (int16 accumulator for N = 8, int32 otherwise)
__global__ void lowpass_avg_kernel(const int16_t* src, int64_t srcBlockStride, int64_t srcStride, int64_t n, int N, int16_t* avg, int32_t* total)
{
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= n) return;
    const int lane = threadIdx.x & 31, half = N >> 1;
    const int16_t* p = src + b * srcBlockStride;
    int tot = 0;
    for (int e = lane; e < half * half; e += 32)
    {
        const int i = e / half, j = e - i * half;
        const int16_t s = (int16_t)((int)p[(int64_t)(2 * i) * srcStride + 2 * j] + (int)p[(int64_t)(2 * i) * srcStride + 2 * j + 1] +
                                    (int)p[(int64_t)(2 * i + 1) * srcStride + 2 * j] + (int)p[(int64_t)(2 * i + 1) * srcStride + 2 * j + 1]);
        avg[b * half * half + e] = (int16_t)(s >> 2);
        tot += s;
    }
    tot = warp_sum(tot);
    if (lane == 0) total[b] = N == 8 ? (int)(int16_t)tot : tot;      // `int16_t totalSum` in lowPassDct8_c (:37)
}
// back end: zero-padded N x N with the half-size coefficients in the top-left quadrant, DC replaced (:58, :86, :110)
__global__ void lowpass_place_kernel(const int16_t* coefHalf, const int32_t* total, int64_t n, int N, int16_t* dst)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * N * N) return;
    const int64_t b = idx / (N * N);
    const int e = (int)(idx - b * N * N), y = e / N, x = e - y * N, half = N >> 1;
    int16_t v = 0;
    if (y < half && x < half) v = coefHalf[b * half * half + y * half + x];
    if (e == 0) v = N == 8 ? (int16_t)(total[b] << 1) : (N == 16 ? (int16_t)(total[b] >> 1) : (int16_t)(total[b] >> 3));
    dst[idx] = v;
}

int var_dev(Ctx* ctx, int depth, int size, const void* src, int64_t stride, const int64_t* off, int64_t n, uint64_t* out)
{
    if (n <= 0) return 0;
    const unsigned blocks = (unsigned)((n + 7) / 8);
    if (depth > 8) var_kernel<uint16_t><<<blocks, 256, 0, ctx->stream>>>((const uint16_t*)src, stride, off, n, size, out);
    else           var_kernel<uint8_t><<<blocks, 256, 0, ctx->stream>>>((const uint8_t*)src, stride, off, n, size, out);
    ctx->launches++;
    return check(cudaGetLastError(), "var launch");
}

int psycost_dev(Ctx* ctx, int depth, int dim, const void* src, int64_t sstride, const void* rec, int64_t rstride,
                const int64_t* offS, const int64_t* offR, int64_t n, int32_t* out)
{
    if (n <= 0) return 0;
    if (dim != 4 && dim != 8 && dim != 16 && dim != 32 && dim != 64) { set_error("psy_cost: size %d", dim); return -1; }
    X265B200_CHECK(cudaMemsetAsync(out, 0, sizeof(int32_t) * n, ctx->stream));
    const int sub = dim >= 8 ? (dim >> 3) * (dim >> 3) : 1;
    const unsigned blocks = (unsigned)((n * sub + 63) / 64);
    if (depth > 8) psycost_kernel<uint16_t><<<blocks, 64, 0, ctx->stream>>>((const uint16_t*)src, sstride, (const uint16_t*)rec, rstride, offS, offR, n, dim, out);
    else           psycost_kernel<uint8_t><<<blocks, 64, 0, ctx->stream>>>((const uint8_t*)src, sstride, (const uint8_t*)rec, rstride, offS, offR, n, dim, out);
    ctx->launches++;
    return check(cudaGetLastError(), "psy_cost launch");
}

int copy_cnt_dev(Ctx* ctx, int size, int16_t* coeff, const int16_t* resi, int64_t stride, const int64_t* off, int64_t n, uint32_t* numSig)
{
    if (n <= 0) return 0;
    copy_cnt_kernel<<<(unsigned)((n + 7) / 8), 256, 0, ctx->stream>>>(coeff, resi, stride, off, n, size, numSig);
    ctx->launches++;
    return check(cudaGetLastError(), "copy_cnt launch");
}

int denoise_dct_dev(Ctx* ctx, int16_t* coef, uint32_t* resSum, const uint16_t* offset, int numCoeff, int64_t n)
{
    if (n <= 0) return 0;
    denoise_kernel<<<(unsigned)((n * numCoeff + 255) / 256), 256, 0, ctx->stream>>>(coef, resSum, offset, numCoeff, n);
    ctx->launches++;
    return check(cudaGetLastError(), "denoiseDct launch");
}

int lowpass_front_dev(Ctx* ctx, const int16_t* src, int64_t srcBlockStride, int64_t srcStride, int64_t n, int N, int16_t* avg, int32_t* total)
{
    lowpass_avg_kernel<<<(unsigned)((n + 7) / 8), 256, 0, ctx->stream>>>(src, srcBlockStride, srcStride, n, N, avg, total);
    ctx->launches++;
    return check(cudaGetLastError(), "lowpass_dct front launch");
}
int lowpass_back_dev(Ctx* ctx, const int16_t* coefHalf, const int32_t* total, int64_t n, int N, int16_t* dst)
{
    lowpass_place_kernel<<<(unsigned)((n * N * N + 255) / 256), 256, 0, ctx->stream>>>(coefHalf, total, n, N, dst);
    ctx->launches++;
    return check(cudaGetLastError(), "lowpass_dct back launch");
}

} // namespace x265b200

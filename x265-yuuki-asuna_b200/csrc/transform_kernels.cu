// transform_kernels.cu -- HEVC core transforms and quantisation as batched sm_100a kernels.
//
// Reference semantics (bit-exact):
//   dct4/8/16/32_c, dst4_c     source/common/dct.cpp:442-525 (butterflies :83-240, :418-440)
//   idct4/8/16/32_c, idst4_c   source/common/dct.cpp:527-610 (inverse butterflies :242-416, :63-81)
//   quant_c / nquant_c         source/common/dct.cpp:664-713
//   dequant_normal/scaling_c   source/common/dct.cpp:612-662
//   count_nonzero / copy_count / denoiseDct   source/common/dct.cpp:714-755
//
// The partial butterflies of the reference are an exact factorisation of the dense product
//       pass(In)[k][j] = ( sum_n T[k][n] * In[j][n] + add ) >> shift
// (all arithmetic is in the ring Z/2^32, so re-association cannot change a bit), forward DCT =
// two such passes with T = g_tN, the result WRAPPED to int16 (dct.cpp:113); the inverse = two
// passes with T^T, each SATURATED to int16 (dct.cpp:257).  For N = 8/16/32 the dense product runs
// on the integer tensor cores (IMMA, mma.sync m16n8k32 / m16n8k16 s8): coefficients (|c| <= 90)
// are s8, int16 samples are split x = 256*hi + lo with hi s8 and lo u8, so
//       T*x = 256*(T*hi) + (T*lo)      exactly in s32.
// 4x4 (DCT and DST) is scalar, one thread per block.
//
// Roofline: HBM-bound when streamed (4*N^2 bytes per block in+out vs 4*N^3 MACs);
// DESIGN.md lists the per-size ridge points.
#include "common.cuh"
#include "tables.cuh"
#include "x265b200.h"

namespace x265b200 {

__constant__ int c_dctMag[33];
// per-lane mma A fragments: [forward/inverse][N = 32, 16, 8][lane][8 regs]
__device__ uint32_t g_xfFrag[2][3][32][8];
static bool g_tablesUploaded[16] = { false };

static int upload_tables(Ctx* ctx)
{
    if (ctx->device < 16 && g_tablesUploaded[ctx->device]) return 0;
    X265B200_CHECK(cudaMemcpyToSymbolAsync(c_dctMag, kDctMag, sizeof(kDctMag), 0, cudaMemcpyHostToDevice, ctx->stream));
    static uint32_t frag[2][3][32][8];
    auto pk = [](int a, int b, int c, int d) { return (uint32_t)(a & 0xff) | ((uint32_t)(b & 0xff) << 8) | ((uint32_t)(c & 0xff) << 16) | ((uint32_t)(d & 0xff) << 24); };
    for (int inv = 0; inv < 2; inv++)
        for (int lane = 0; lane < 32; lane++)
        {
            const int gid = lane >> 2, tig = lane & 3;
            auto M = [&](int N, int k, int n) { return inv ? dct_coef(N, n, k) : dct_coef(N, k, n); };
            for (int mi = 0; mi < 2; mi++)
                for (int r = 0; r < 4; r++)
                {
                    int row = 16 * mi + gid + ((r & 1) ? 8 : 0), k0 = 4 * tig + ((r & 2) ? 16 : 0);
                    frag[inv][0][lane][mi * 4 + r] = pk(M(32, row, k0), M(32, row, k0 + 1), M(32, row, k0 + 2), M(32, row, k0 + 3));
                }
            for (int r = 0; r < 2; r++)
            {
                int row = gid + 8 * r, k0 = 4 * tig;
                frag[inv][1][lane][r] = pk(M(16, row, k0), M(16, row, k0 + 1), M(16, row, k0 + 2), M(16, row, k0 + 3));
            }
            int k0 = 4 * (tig & 1);
            uint32_t v = pk(M(8, gid, k0), M(8, gid, k0 + 1), M(8, gid, k0 + 2), M(8, gid, k0 + 3));
            frag[inv][2][lane][0] = tig < 2 ? v : 0u;        // rows 0-7 see k 0-7
            frag[inv][2][lane][1] = tig >= 2 ? v : 0u;       // rows 8-15 see k 8-15
            for (int r = 2; r < 8; r++) { frag[inv][1][lane][r] = 0; frag[inv][2][lane][r] = 0; }
        }
    X265B200_CHECK(cudaMemcpyToSymbolAsync(g_xfFrag, frag, sizeof(frag), 0, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->device < 16) g_tablesUploaded[ctx->device] = true;
    return 0;
}

__device__ __forceinline__ int dct_coef_dev(int N, int k, int n)
{
    int m = ((2 * n + 1) * k * (32 / N)) & 127;
    if (m > 64) m = 128 - m;
    return m > 32 ? -c_dctMag[64 - m] : c_dctMag[m];
}

// M[k][n]: forward uses T[k][n], inverse uses T[n][k]
__device__ __forceinline__ int mcoef(int N, bool inverse, int k, int n)
{
    return inverse ? dct_coef_dev(N, n, k) : dct_coef_dev(N, k, n);
}

__device__ __forceinline__ uint32_t pack4(int b0, int b1, int b2, int b3)
{
    return (uint32_t)(b0 & 0xff) | ((uint32_t)(b1 & 0xff) << 8) | ((uint32_t)(b2 & 0xff) << 16) | ((uint32_t)(b3 & 0xff) << 24);
}

__device__ __forceinline__ void mma_k32_s8u8(int d[4], const uint32_t a[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_k32_s8s8(int d[4], const uint32_t a[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_k16_s8u8(int d[4], const uint32_t a[2], uint32_t b0)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(b0));
}
__device__ __forceinline__ void mma_k16_s8s8(int d[4], const uint32_t a[2], uint32_t b0)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(b0));
}

template<bool SAT> __device__ __forceinline__ int finish(int hi, int lo, int add, int shift)
{
    int v = (int)((uint32_t)hi * 256u + (uint32_t)lo);
    v = (v + add) >> shift;
    if (SAT) v = clip3i(-32768, 32767, v);
    return v & 0xffff;     // int16 wrap (forward) / already in range (inverse)
}

// LD (row pitch in int16) of the per-warp smem tiles
template<int N> struct Tile { static constexpr int LD = N + 8; static constexpr int ELEMS = N * (N + 8); };

// One dense pass on a 32x32 tile held in smem: Out[k][j] = f(sum_n M[k][n] In[j][n])
template<bool SAT>
__device__ __forceinline__ void pass32(const uint32_t a[2][4], const int16_t* In, int16_t* Out, int add, int shift, int gid, int tig)
{
    constexpr int LD = Tile<32>::LD;
#pragma unroll
    for (int ni = 0; ni < 4; ni++)
    {
        const int16_t* p = In + (8 * ni + gid) * LD + 4 * tig;
        uint2 w0 = *(const uint2*)p, w1 = *(const uint2*)(p + 16);
        uint32_t b0lo = __byte_perm(w0.x, w0.y, 0x6420), b0hi = __byte_perm(w0.x, w0.y, 0x7531);
        uint32_t b1lo = __byte_perm(w1.x, w1.y, 0x6420), b1hi = __byte_perm(w1.x, w1.y, 0x7531);
#pragma unroll
        for (int mi = 0; mi < 2; mi++)
        {
            int lo[4] = { 0, 0, 0, 0 }, hi[4] = { 0, 0, 0, 0 };
            mma_k32_s8u8(lo, a[mi], b0lo, b1lo);
            mma_k32_s8s8(hi, a[mi], b0hi, b1hi);
            int r0 = finish<SAT>(hi[0], lo[0], add, shift), r1 = finish<SAT>(hi[1], lo[1], add, shift);
            int r2 = finish<SAT>(hi[2], lo[2], add, shift), r3 = finish<SAT>(hi[3], lo[3], add, shift);
            *(uint32_t*)(Out + (16 * mi + gid) * LD + 8 * ni + 2 * tig) = (uint32_t)r0 | ((uint32_t)r1 << 16);
            *(uint32_t*)(Out + (16 * mi + gid + 8) * LD + 8 * ni + 2 * tig) = (uint32_t)r2 | ((uint32_t)r3 << 16);
        }
    }
}

template<bool SAT>
__device__ __forceinline__ void pass16(const uint32_t a[2], const int16_t* In, int16_t* Out, int add, int shift, int gid, int tig)
{
    constexpr int LD = Tile<16>::LD;
#pragma unroll
    for (int ni = 0; ni < 2; ni++)
    {
        const int16_t* p = In + (8 * ni + gid) * LD + 4 * tig;
        uint2 w0 = *(const uint2*)p;
        uint32_t blo = __byte_perm(w0.x, w0.y, 0x6420), bhi = __byte_perm(w0.x, w0.y, 0x7531);
        int lo[4] = { 0, 0, 0, 0 }, hi[4] = { 0, 0, 0, 0 };
        mma_k16_s8u8(lo, a, blo);
        mma_k16_s8s8(hi, a, bhi);
        int r0 = finish<SAT>(hi[0], lo[0], add, shift), r1 = finish<SAT>(hi[1], lo[1], add, shift);
        int r2 = finish<SAT>(hi[2], lo[2], add, shift), r3 = finish<SAT>(hi[3], lo[3], add, shift);
        *(uint32_t*)(Out + gid * LD + 8 * ni + 2 * tig) = (uint32_t)r0 | ((uint32_t)r1 << 16);
        *(uint32_t*)(Out + (gid + 8) * LD + 8 * ni + 2 * tig) = (uint32_t)r2 | ((uint32_t)r3 << 16);
    }
}

// Two 8x8 blocks per mma through a block-diagonal A: rows 0-7 x k 0-7 = M8 (block 0),
// rows 8-15 x k 8-15 = M8 (block 1).  In/Out hold the two blocks back to back ([2][8][LD]).
template<bool SAT>
__device__ __forceinline__ void pass8(const uint32_t a[2], const int16_t* In, int16_t* Out, int add, int shift, int gid, int tig)
{
    constexpr int LD = Tile<8>::LD;
    const int blk = tig >> 1;                                   // k 0..7 -> block 0, k 8..15 -> block 1
    const int16_t* p = In + blk * 8 * LD + gid * LD + 4 * (tig & 1);
    uint2 w0 = *(const uint2*)p;
    uint32_t blo = __byte_perm(w0.x, w0.y, 0x6420), bhi = __byte_perm(w0.x, w0.y, 0x7531);
    int lo[4] = { 0, 0, 0, 0 }, hi[4] = { 0, 0, 0, 0 };
    mma_k16_s8u8(lo, a, blo);
    mma_k16_s8s8(hi, a, bhi);
    int r0 = finish<SAT>(hi[0], lo[0], add, shift), r1 = finish<SAT>(hi[1], lo[1], add, shift);
    int r2 = finish<SAT>(hi[2], lo[2], add, shift), r3 = finish<SAT>(hi[3], lo[3], add, shift);
    *(uint32_t*)(Out + gid * LD + 2 * tig) = (uint32_t)r0 | ((uint32_t)r1 << 16);                 // block 0: R[k=gid][j]
    *(uint32_t*)(Out + 8 * LD + gid * LD + 2 * tig) = (uint32_t)r2 | ((uint32_t)r3 << 16);        // block 1
}

struct XformArgs
{
    const int16_t* src; int64_t srcBlockStride; int64_t srcStride;   // forward: strided src, contiguous dst
    int16_t* dst;       int64_t dstBlockStride; int64_t dstStride;   // inverse: contiguous src, strided dst
    int64_t n;
    int64_t bpr;          // blocks per row of a 2-D block grid (>= n: plain linear list)
    int64_t srcRowStride, dstRowStride;   // element offset between block rows of the grid
    int shift1, shift2;
};

__device__ __forceinline__ int64_t xf_src_off(const XformArgs& p, int64_t b) { return (b / p.bpr) * p.srcRowStride + (b % p.bpr) * p.srcBlockStride; }
__device__ __forceinline__ int64_t xf_dst_off(const XformArgs& p, int64_t b) { return (b / p.bpr) * p.dstRowStride + (b % p.bpr) * p.dstBlockStride; }

constexpr int XF_WARPS = 8;

// N = 32 or 16: one block per warp-iteration.  N = 8: two blocks per warp-iteration.
template<int N, bool INVERSE>
__global__ void __launch_bounds__(XF_WARPS * 32)
xform_mma_kernel(XformArgs p)
{
    constexpr int LD = Tile<N>::LD;
    constexpr int PER = (N == 8) ? 2 : 1;
    constexpr int TILE = PER * N * LD;
    __shared__ __align__(16) int16_t smem[XF_WARPS][2][TILE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
    int16_t* bufA = smem[warp][0];
    int16_t* bufB = smem[warp][1];

    // A fragments (coefficient matrix) live in registers for the whole kernel; they come from a per-lane table
    // built once on the host (g_xfFrag) -- computing them in-kernel from __constant__ memory with divergent
    // indices saturated the constant-cache address unit (ncu: ADU 87%).
    uint32_t a32[2][4]; uint32_t a16[2];
    {
        const uint32_t* f = g_xfFrag[INVERSE ? 1 : 0][N == 32 ? 0 : (N == 16 ? 1 : 2)][lane];
        if (N == 32)
        {
#pragma unroll
            for (int mi = 0; mi < 2; mi++)
#pragma unroll
                for (int r = 0; r < 4; r++) a32[mi][r] = f[mi * 4 + r];
        }
        else { a16[0] = f[0]; a16[1] = f[1]; }
    }

    const int add1 = 1 << (p.shift1 - 1), add2 = 1 << (p.shift2 - 1);
    const int64_t units = (p.n + PER - 1) / PER;
    constexpr int CPB = N * N / 8;                 // 16-byte chunks per block
    constexpr int CH = PER * CPB;                  // chunks per unit
    constexpr int CNT = (CH + 31) / 32;            // chunks per lane
    constexpr int LPR = N / 8;                     // chunks per row

    // 128-bit path: whole unit addressable with aligned uint4 loads?
    auto vec_ok = [&](int64_t u) -> bool {
        bool ok = true;
#pragma unroll
        for (int s = 0; s < PER; s++)
        {
            int64_t b = u * PER + s;
            if (b >= p.n) continue;
            uintptr_t a = (uintptr_t)(p.src + xf_src_off(p, b));
            if (!INVERSE) a |= (uintptr_t)(p.srcStride * 2);
            ok = ok && !(a & 15);
        }
        return ok;
    };
    auto fetch = [&](int64_t u, uint4 regs[CNT]) {
#pragma unroll
        for (int i = 0; i < CNT; i++)
        {
            const int c = lane + 32 * i;
            regs[i] = make_uint4(0, 0, 0, 0);
            if (c < CH)
            {
                const int s = c / CPB, cc = c % CPB, row = cc / LPR, c8 = (cc % LPR) * 8;
                const int64_t b = u * PER + s;
                if (b < p.n)
                    regs[i] = __ldg((const uint4*)(p.src + xf_src_off(p, b) + (INVERSE ? (int64_t)row * N : (int64_t)row * p.srcStride) + c8));
            }
        }
    };
    auto commit = [&](const uint4 regs[CNT]) {
#pragma unroll
        for (int i = 0; i < CNT; i++)
        {
            const int c = lane + 32 * i;
            if (c < CH)
            {
                const int s = c / CPB, cc = c % CPB, row = cc / LPR, c8 = (cc % LPR) * 8;
                int16_t* tile = bufA + s * N * LD;
                if (!INVERSE) *(uint4*)(tile + row * LD + c8) = regs[i];
                else
                {
                    const uint32_t w[4] = { regs[i].x, regs[i].y, regs[i].z, regs[i].w };
#pragma unroll
                    for (int k = 0; k < 4; k++)
                    {
                        tile[(c8 + 2 * k) * LD + row] = (int16_t)(w[k] & 0xffff);
                        tile[(c8 + 2 * k + 1) * LD + row] = (int16_t)(w[k] >> 16);
                    }
                }
            }
        }
    };
    auto scalar_load = [&](int64_t u) {
#pragma unroll
        for (int s = 0; s < PER; s++)
        {
            int64_t b = u * PER + s;
            int16_t* tile = bufA + s * N * LD;
            if (b < p.n)
            {
                const int16_t* sp = p.src + xf_src_off(p, b);
                for (int e = lane; e < N * N / 2; e += 32)
                {
                    int row = (2 * e) / N, col = (2 * e) % N;
                    uint32_t w = ld_px2((const uint16_t*)(sp + (INVERSE ? (int64_t)row * N : (int64_t)row * p.srcStride) + col));
                    if (!INVERSE) *(uint32_t*)(tile + row * LD + col) = w;
                    else { tile[col * LD + row] = (int16_t)(w & 0xffff); tile[(col + 1) * LD + row] = (int16_t)(w >> 16); }
                }
            }
            else
                for (int e = lane; e < N * LD / 2; e += 32) ((uint32_t*)tile)[e] = 0;
        }
    };

    const int64_t ustride = (int64_t)gridDim.x * XF_WARPS;
    int64_t u = (int64_t)blockIdx.x * XF_WARPS + warp;
    uint4 pre[CNT];
    bool pv = false;
    if (u < units) { pv = vec_ok(u); if (pv) fetch(u, pre); }
    for (; u < units; u += ustride)
    {
        // ---- load: In[j][n] = src[j][n] (forward) or src[n][j] (inverse, contiguous NxN) ----
        if (pv) commit(pre); else scalar_load(u);
        // software pipeline: the next unit's global loads are in flight during this unit's tensor-core passes
        const int64_t nu = u + ustride;
        bool nv = false;
        if (nu < units) { nv = vec_ok(nu); if (nv) fetch(nu, pre); }
        __syncwarp();
        if (N == 32)      { pass32<INVERSE>(a32, bufA, bufB, add1, p.shift1, gid, tig); __syncwarp(); pass32<INVERSE>(a32, bufB, bufA, add2, p.shift2, gid, tig); }
        else if (N == 16) { pass16<INVERSE>(a16, bufA, bufB, add1, p.shift1, gid, tig); __syncwarp(); pass16<INVERSE>(a16, bufB, bufA, add2, p.shift2, gid, tig); }
        else              { pass8<INVERSE>(a16, bufA, bufB, add1, p.shift1, gid, tig);  __syncwarp(); pass8<INVERSE>(a16, bufB, bufA, add2, p.shift2, gid, tig); }
        __syncwarp();
        // ---- store: forward dst[k*N + j] = R[k][j]; inverse dst[j*dstStride + k] = R[k][j] ----
#pragma unroll
        for (int s = 0; s < PER; s++)
        {
            int64_t b = u * PER + s;
            if (b >= p.n) continue;
            const int16_t* tile = bufA + s * N * LD;
            int16_t* dp = p.dst + xf_dst_off(p, b);
            if (!INVERSE && !((uintptr_t)dp & 15))
            {
                for (int e = lane; e < CPB; e += 32)
                {
                    int row = e / LPR, c8 = (e % LPR) * 8;
                    *(uint4*)(dp + row * N + c8) = *(const uint4*)(tile + row * LD + c8);
                }
            }
            else if (!INVERSE)
                for (int e = lane; e < N * N / 2; e += 32)
                {
                    int row = (2 * e) / N, col = (2 * e) % N;
                    if ((uintptr_t)dp & 3) { dp[row * N + col] = tile[row * LD + col]; dp[row * N + col + 1] = tile[row * LD + col + 1]; }
                    else *(uint32_t*)(dp + row * N + col) = *(const uint32_t*)(tile + row * LD + col);
                }
            else if (!(((uintptr_t)dp | (uintptr_t)(p.dstStride * 2)) & 15))
                for (int e = lane; e < CPB; e += 32)
                {
                    int j = e / LPR, c8 = (e % LPR) * 8;           // output row j, columns c8..c8+7 = R[c8+k][j]
                    uint32_t w[4];
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        w[k] = (uint32_t)(uint16_t)tile[(c8 + 2 * k) * LD + j] | ((uint32_t)(uint16_t)tile[(c8 + 2 * k + 1) * LD + j] << 16);
                    *(uint4*)(dp + (int64_t)j * p.dstStride + c8) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            else
                for (int e = lane; e < N * N; e += 32)
                {
                    int j = e / N, k = e % N;
                    dp[(int64_t)j * p.dstStride + k] = tile[k * LD + j];
                }
        }
        __syncwarp();
        pv = nv;
    }
}

// ---- 4x4: DCT (kind 0) / DST (kind 1), one thread per block ----------------------------------
__device__ __forceinline__ int16_t wrap16(int v) { return (int16_t)v; }
__device__ __forceinline__ int16_t sat16(int v) { return (int16_t)clip3i(-32768, 32767, v); }

// dct.cpp:418-440 partialButterfly4 / :43-61 fastForwardDst : out[k*4 + j] from in[j*4 + n]
__device__ __forceinline__ void fwd4_pass(const int in[16], int out[16], int shift, bool dst)
{
    const int add = 1 << (shift - 1);
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
        const int s0 = in[4 * j], s1 = in[4 * j + 1], s2 = in[4 * j + 2], s3 = in[4 * j + 3];
        if (!dst)
        {
            int E0 = s0 + s3, O0 = s0 - s3, E1 = s1 + s2, O1 = s1 - s2;
            out[j]      = wrap16((64 * E0 + 64 * E1 + add) >> shift);
            out[8 + j]  = wrap16((64 * E0 - 64 * E1 + add) >> shift);
            out[4 + j]  = wrap16((83 * O0 + 36 * O1 + add) >> shift);
            out[12 + j] = wrap16((36 * O0 - 83 * O1 + add) >> shift);
        }
        else
        {
            int c0 = s0 + s3, c1 = s1 + s3, c2 = s0 - s1, c3 = 74 * s2;
            out[j]      = wrap16((29 * c0 + 55 * c1 + c3 + add) >> shift);
            out[4 + j]  = wrap16((74 * (s0 + s1 - s3) + add) >> shift);
            out[8 + j]  = wrap16((29 * c2 + 55 * c0 - c3 + add) >> shift);
            out[12 + j] = wrap16((55 * c2 - 29 * c1 + c3 + add) >> shift);
        }
    }
}

// dct.cpp:242-265 partialButterflyInverse4 / :63-81 inversedst : out[j*4 + k] from in[n*4 + j]
__device__ __forceinline__ void inv4_pass(const int in[16], int out[16], int shift, bool dst)
{
    const int add = 1 << (shift - 1);
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
        const int t0 = in[j], t1 = in[4 + j], t2 = in[8 + j], t3 = in[12 + j];
        if (!dst)
        {
            int O0 = 83 * t1 + 36 * t3, O1 = 36 * t1 - 83 * t3, E0 = 64 * t0 + 64 * t2, E1 = 64 * t0 - 64 * t2;
            out[4 * j]     = sat16((E0 + O0 + add) >> shift);
            out[4 * j + 1] = sat16((E1 + O1 + add) >> shift);
            out[4 * j + 2] = sat16((E1 - O1 + add) >> shift);
            out[4 * j + 3] = sat16((E0 - O0 + add) >> shift);
        }
        else
        {
            int c0 = t0 + t2, c1 = t2 + t3, c2 = t0 - t3, c3 = 74 * t1;
            out[4 * j]     = sat16((29 * c0 + 55 * c1 + c3 + add) >> shift);
            out[4 * j + 1] = sat16((55 * c2 - 29 * c1 + c3 + add) >> shift);
            out[4 * j + 2] = sat16((74 * (t0 - t2 + t3) + add) >> shift);
            out[4 * j + 3] = sat16((55 * c0 + 29 * c2 - c3 + add) >> shift);
        }
    }
}

template<bool INVERSE>
__global__ void __launch_bounds__(128)
xform4_kernel(XformArgs p, int isDst)
{
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.n) return;
    int in[16], tmp[16], out[16];
    const int16_t* sp = p.src + xf_src_off(p, b);
#pragma unroll
    for (int r = 0; r < 4; r++)
    {
        int v[4];
        ld4s(sp + (INVERSE ? r * 4 : (int64_t)r * p.srcStride), v);
#pragma unroll
        for (int c = 0; c < 4; c++) in[4 * r + c] = v[c];
    }
    if (!INVERSE) { fwd4_pass(in, tmp, p.shift1, isDst); fwd4_pass(tmp, out, p.shift2, isDst); }
    else          { inv4_pass(in, tmp, p.shift1, isDst); inv4_pass(tmp, out, p.shift2, isDst); }
    int16_t* dp = p.dst + xf_dst_off(p, b);
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++)
            dp[(INVERSE ? (int64_t)r * p.dstStride : r * 4) + c] = (int16_t)out[4 * r + c];
}

// sizeIdx: 0..3 = 4/8/16/32 DCT, 4 = 4x4 DST (same numbering as the reference TestBench harness)
int transform_dev(Ctx* ctx, int inverse, int sizeIdx, int depth, const int16_t* src, int64_t srcBlockStride, int64_t srcStride,
                  int16_t* dst, int64_t dstBlockStride, int64_t dstStride, int64_t n, int64_t blocksPerRow, int64_t rowStride)
{
    if (sizeIdx < 0 || sizeIdx > 4) { set_error("transform: sizeIdx %d", sizeIdx); return -1; }
    if (n <= 0) return 0;
    if (upload_tables(ctx)) return -1;
    const int N = sizeIdx == 4 ? 4 : (4 << sizeIdx);
    const int log2N = sizeIdx == 4 ? 2 : sizeIdx + 2;
    XformArgs a;
    a.src = src; a.srcBlockStride = srcBlockStride; a.srcStride = srcStride;
    a.dst = dst; a.dstBlockStride = dstBlockStride; a.dstStride = dstStride; a.n = n;
    a.bpr = blocksPerRow > 0 ? blocksPerRow : (n > 0 ? n : 1);
    // forward: the strided side is src, dst is a linear list; inverse: the other way round
    a.srcRowStride = inverse ? a.bpr * a.srcBlockStride : rowStride;
    a.dstRowStride = inverse ? rowStride : a.bpr * a.dstBlockStride;
    if (blocksPerRow <= 0) { a.srcRowStride = a.bpr * a.srcBlockStride; a.dstRowStride = a.bpr * a.dstBlockStride; }
    if (!inverse) { a.shift1 = log2N - 1 + (depth - 8); a.shift2 = log2N + 6; }      // dct.cpp:444-445, 478-479, ...
    else          { a.shift1 = 7; a.shift2 = 12 - (depth - 8); }                     // dct.cpp:529-530
    if (N == 4)
    {
        unsigned blocks = (unsigned)((n + 127) / 128);
        if (!inverse) xform4_kernel<false><<<blocks, 128, 0, ctx->stream>>>(a, sizeIdx == 4);
        else          xform4_kernel<true><<<blocks, 128, 0, ctx->stream>>>(a, sizeIdx == 4);
    }
    else
    {
        int64_t units = N == 8 ? (n + 1) / 2 : n;
        int64_t want = (units + XF_WARPS - 1) / XF_WARPS;
        int64_t cap = (int64_t)ctx->smCount * 3;      // ~2-3 units per warp so the load prefetch has something to overlap
        unsigned blocks = (unsigned)(want < cap ? want : cap);
        dim3 block(XF_WARPS * 32);
        if (N == 32) { if (!inverse) xform_mma_kernel<32, false><<<blocks, block, 0, ctx->stream>>>(a); else xform_mma_kernel<32, true><<<blocks, block, 0, ctx->stream>>>(a); }
        if (N == 16) { if (!inverse) xform_mma_kernel<16, false><<<blocks, block, 0, ctx->stream>>>(a); else xform_mma_kernel<16, true><<<blocks, block, 0, ctx->stream>>>(a); }
        if (N == 8)  { if (!inverse) xform_mma_kernel<8, false><<<blocks, block, 0, ctx->stream>>>(a);  else xform_mma_kernel<8, true><<<blocks, block, 0, ctx->stream>>>(a); }
    }
    ctx->launches++;
    return check(cudaGetLastError(), "transform kernel launch");
}

// ---------------------------------------------------------------------------------------------
// quant / nquant / dequant / count_nonzero / denoise : one warp per TU block
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
quant_kernel(const int16_t* __restrict__ coef, const int32_t* __restrict__ quantCoeff, int32_t* __restrict__ deltaU,
             int16_t* __restrict__ qCoef, int qBits, int add, int numCoeff, int64_t n, uint32_t* __restrict__ numSig, int isN)
{
    const int lane = threadIdx.x & 31;
    const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= n) return;
    const int16_t* c = coef + b * numCoeff;
    int cnt = 0;
    for (int i = lane; i < numCoeff; i += 32)
    {
        int level = c[i];
        int sign = level < 0 ? -1 : 1;
        int tmplevel = (int)((uint32_t)abs(level) * (uint32_t)quantCoeff[i]);
        level = (tmplevel + add) >> qBits;
        if (!isN && deltaU) deltaU[b * numCoeff + i] = (tmplevel - (level << qBits)) >> (qBits - 8);
        if (level) cnt++;
        level *= sign;
        int q = clip3i(-32768, 32767, level);
        qCoef[b * numCoeff + i] = (int16_t)(isN ? abs(q) : q);      // dct.cpp:682 / :709
    }
    cnt = warp_sum(cnt);
    if (lane == 0 && numSig) numSig[b] = (uint32_t)cnt;
}

int quant_dev(Ctx* ctx, int isN, const int16_t* coef, const int32_t* quantCoeff, int32_t* deltaU, int16_t* qCoef,
              int qBits, int add, int numCoeff, int64_t n, uint32_t* numSig)
{
    if (n <= 0) return 0;
    unsigned blocks = (unsigned)((n * 32 + 255) / 256);
    quant_kernel<<<blocks, 256, 0, ctx->stream>>>(coef, quantCoeff, deltaU, qCoef, qBits, add, numCoeff, n, numSig, isN);
    ctx->launches++;
    return check(cudaGetLastError(), "quant kernel launch");
}

__global__ void __launch_bounds__(256)
dequant_kernel(const int16_t* __restrict__ q, const int32_t* __restrict__ deqCoef, int16_t* __restrict__ coef,
               int num, int64_t total, int scaleOrPer, int shift, int scaling)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int v = q[i];
    int r;
    if (!scaling)
    {
        int add = 1 << (shift - 1);
        r = clip3i(-32768, 32767, (v * scaleOrPer + add) >> shift);                  // dct.cpp:631-632
    }
    else
    {
        int per = scaleOrPer, sh = shift + 4, d = deqCoef[i % num];
        if (sh > per) { int add = 1 << (sh - per - 1); r = clip3i(-32768, 32767, (v * d + add) >> (sh - per)); }       // :646-652
        else          { int c = clip3i(-32768, 32767, v * d); r = clip3i(-32768, 32767, (int)((uint32_t)c << (per - sh))); } // :656-660
    }
    coef[i] = (int16_t)r;
}

int dequant_dev(Ctx* ctx, int scaling, const int16_t* q, const int32_t* deqCoef, int16_t* coef, int num, int64_t n,
                int scaleOrPer, int shift)
{
    int64_t total = n * num;
    if (total <= 0) return 0;
    unsigned blocks = (unsigned)((total + 255) / 256);
    dequant_kernel<<<blocks, 256, 0, ctx->stream>>>(q, deqCoef, coef, num, total, scaleOrPer, shift, scaling);
    ctx->launches++;
    return check(cudaGetLastError(), "dequant kernel launch");
}

// count_nonzero (dct.cpp:714-726) over n contiguous blocks of numCoeff
__global__ void __launch_bounds__(256)
count_nonzero_kernel(const int16_t* __restrict__ q, int numCoeff, int64_t n, int32_t* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= n) return;
    int cnt = 0;
    for (int i = lane; i < numCoeff; i += 32) cnt += q[b * numCoeff + i] != 0;
    cnt = warp_sum(cnt);
    if (lane == 0) out[b] = cnt;
}

int count_nonzero_dev(Ctx* ctx, const int16_t* q, int numCoeff, int64_t n, int32_t* out)
{
    if (n <= 0) return 0;
    unsigned blocks = (unsigned)((n * 32 + 255) / 256);
    count_nonzero_kernel<<<blocks, 256, 0, ctx->stream>>>(q, numCoeff, n, out);
    ctx->launches++;
    return check(cudaGetLastError(), "count_nonzero kernel launch");
}

// ---------------------------------------------------------------------------------------------
// Fused residual pipeline: per TU, in one kernel and with nothing but pixels in / pixels + levels out,
//   residual = fenc - pred                      cu[].sub_ps           (pixel.cpp:814)
//   coef     = DCT(residual)  (DST for 4x4 luma intra)               Quant::transformNxN (quant.cpp:397-480, rdoq 0,
//   level    = quant(coef), numSig              primitives.quant      no sign hiding / noise reduction / transform skip)
//   coef'    = dequant_normal | dequant_scaling(level)                Quant::invtransformNxN (quant.cpp:543-605) incl. its
//   resi'    = 0 | DC fill | IDCT(coef')                              numSig == 0 (search.cpp blockfill_s 0) and DC-only shortcuts
//   recon    = clip(pred + resi')               cu[].add_ps           (pixel.cpp:828)
//   sse      = sum (fenc - recon)^2             cu[].sse_pp           (pixel.cpp:167)
// N = 8/16/32: one warp per TU (two 8x8 TUs per warp), both transforms on the integer tensor cores exactly as
// xform_mma_kernel, quant + dequant between them on the smem tile (written back transposed for the inverse).
// Algorithmic bytes per TU: N*N*(3*sizeof(pixel) + 2) + 12.
// ---------------------------------------------------------------------------------------------
struct TuArgs
{
    const void* fenc; int64_t fencStride; const void* pred; int64_t predStride; void* recon; int64_t reconStride;
    int blocksX; int64_t n;
    const int32_t* quantCoeff; int qBits, add;
    const int32_t* dequantCoef; int scaleOrPer, dqShift;
    int16_t* coeff; uint32_t* numSig; uint64_t* sse;
    int depth, useDST, wordStores;
};

__device__ __forceinline__ int tu_quant(int c, int q, int add, int qBits, int& cnt)
{
    const int sign = c < 0 ? -1 : 1;
    const int tmplevel = (int)((uint32_t)abs(c) * (uint32_t)q);
    int level = (tmplevel + add) >> qBits;
    if (level) cnt++;
    level *= sign;
    return clip3i(-32768, 32767, level);
}
__device__ __forceinline__ int tu_dequant(const TuArgs& p, int v, int idx)
{
    if (!p.dequantCoef)
        return clip3i(-32768, 32767, (v * p.scaleOrPer + (1 << (p.dqShift - 1))) >> p.dqShift);              // dct.cpp:631-632
    const int per = p.scaleOrPer, sh = p.dqShift + 4, d = __ldg(p.dequantCoef + idx);
    if (sh > per) return clip3i(-32768, 32767, (v * d + (1 << (sh - per - 1))) >> (sh - per));                 // :646-652
    const int c = clip3i(-32768, 32767, v * d);
    return clip3i(-32768, 32767, (int)((uint32_t)c << (per - sh)));                                           // :656-660
}
// DC-only inverse (quant.cpp:585-595)
__device__ __forceinline__ int tu_dc_val(int deq0, int depth)
{
    const int shift_2nd = 12 - (depth - 8) - 3, add_2nd = 1 << (shift_2nd - 1);
    return (int)(int16_t)(((((deq0 * (64 >> 6) + 1) >> 1) * (64 >> 3)) + add_2nd) >> shift_2nd);
}
template<typename pixel> __device__ __forceinline__ uint32_t tu_ldw(const pixel* p);
template<> __device__ __forceinline__ uint32_t tu_ldw<uint8_t>(const uint8_t* p) { return ld_px4(p); }
template<> __device__ __forceinline__ uint32_t tu_ldw<uint16_t>(const uint16_t* p) { return ld_px2(p); }

template<typename pixel, int N>
__global__ void __launch_bounds__(XF_WARPS * 32)
tu_pipeline_kernel(TuArgs p)
{
    constexpr int LD = Tile<N>::LD;
    constexpr int PER = (N == 8) ? 2 : 1;
    constexpr int TILE = PER * N * LD;
    constexpr int PW = 4 / (int)sizeof(pixel);                  // pixels per 32-bit word
    constexpr int WPT = N * N / PW;                             // words per TU
    constexpr int WPU = PER * WPT;
    constexpr int CNT = (WPU + 31) / 32;
    constexpr int PAIRS = PER * N * N / 2;
    __shared__ __align__(16) int16_t smem[XF_WARPS][2][TILE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, tig = lane & 3;
    int16_t* bufA = smem[warp][0];
    int16_t* bufB = smem[warp][1];
    const int maxVal = (1 << p.depth) - 1;
    const int log2N = N == 32 ? 5 : (N == 16 ? 4 : 3);

    uint32_t af32[2][4], ai32[2][4]; uint32_t af16[2], ai16[2];
    {
        const int sz = N == 32 ? 0 : (N == 16 ? 1 : 2);
        const uint32_t* ff = g_xfFrag[0][sz][lane];
        const uint32_t* fi = g_xfFrag[1][sz][lane];
        if (N == 32)
        {
#pragma unroll
            for (int mi = 0; mi < 2; mi++)
#pragma unroll
                for (int r = 0; r < 4; r++) { af32[mi][r] = ff[mi * 4 + r]; ai32[mi][r] = fi[mi * 4 + r]; }
        }
        else { af16[0] = ff[0]; af16[1] = ff[1]; ai16[0] = fi[0]; ai16[1] = fi[1]; }
    }
    const int fs1 = log2N - 1 + (p.depth - 8), fs2 = log2N + 6;              // dct.cpp:478-479 etc.
    const int is1 = 7, is2 = 12 - (p.depth - 8);                             // dct.cpp:529-530
    const pixel* fenc = (const pixel*)p.fenc; const pixel* pred = (const pixel*)p.pred; pixel* recon = (pixel*)p.recon;

    const int64_t units = (p.n + PER - 1) / PER;
    for (int64_t u = (int64_t)blockIdx.x * XF_WARPS + warp; u < units; u += (int64_t)gridDim.x * XF_WARPS)
    {
        // ---- residual = fenc - pred into the forward tile In[j][n] ----
        uint32_t fw[CNT], pw[CNT];
#pragma unroll
        for (int i = 0; i < CNT; i++)
        {
            const int e = lane + 32 * i;
            fw[i] = pw[i] = 0;
            if (e < WPU)
            {
                const int s = e / WPT, ee = e - s * WPT, row = (ee * PW) >> log2N, col = (ee * PW) & (N - 1);
                const int64_t b = u * PER + s;
                int16_t* t = bufA + s * N * LD + row * LD + col;
                if (b < p.n)
                {
                    const int64_t by = b / p.blocksX, bx = b - by * p.blocksX;
                    fw[i] = tu_ldw<pixel>(fenc + (by * N + row) * p.fencStride + bx * N + col);
                    pw[i] = tu_ldw<pixel>(pred + (by * N + row) * p.predStride + bx * N + col);
                }
                if (sizeof(pixel) == 1)
                {
                    const int d0 = (int)(fw[i] & 0xff) - (int)(pw[i] & 0xff), d1 = (int)((fw[i] >> 8) & 0xff) - (int)((pw[i] >> 8) & 0xff);
                    const int d2 = (int)((fw[i] >> 16) & 0xff) - (int)((pw[i] >> 16) & 0xff), d3 = (int)(fw[i] >> 24) - (int)(pw[i] >> 24);
                    *(uint2*)t = make_uint2((uint32_t)(d0 & 0xffff) | ((uint32_t)d1 << 16), (uint32_t)(d2 & 0xffff) | ((uint32_t)d3 << 16));
                }
                else
                {
                    const int d0 = (int)(fw[i] & 0xffff) - (int)(pw[i] & 0xffff), d1 = (int)(fw[i] >> 16) - (int)(pw[i] >> 16);
                    *(uint32_t*)t = (uint32_t)(d0 & 0xffff) | ((uint32_t)d1 << 16);
                }
            }
        }
        __syncwarp();
        if (N == 32)      { pass32<false>(af32, bufA, bufB, 1 << (fs1 - 1), fs1, gid, tig); __syncwarp(); pass32<false>(af32, bufB, bufA, 1 << (fs2 - 1), fs2, gid, tig); }
        else if (N == 16) { pass16<false>(af16, bufA, bufB, 1 << (fs1 - 1), fs1, gid, tig); __syncwarp(); pass16<false>(af16, bufB, bufA, 1 << (fs2 - 1), fs2, gid, tig); }
        else              { pass8<false>(af16, bufA, bufB, 1 << (fs1 - 1), fs1, gid, tig);  __syncwarp(); pass8<false>(af16, bufB, bufA, 1 << (fs2 - 1), fs2, gid, tig); }
        __syncwarp();

        // ---- quant (levels out) + dequant into the inverse tile In'[j][n] = coef'[n][j] ----
        int cnt[PER], dcq[PER], dcd[PER];
#pragma unroll
        for (int s = 0; s < PER; s++) { cnt[s] = 0; dcq[s] = 0; dcd[s] = 0; }
#pragma unroll 4
        for (int e2 = lane; e2 < PAIRS; e2 += 32)
        {
            const int s = e2 / (N * N / 2), idx = (e2 - s * (N * N / 2)) * 2, k = idx >> log2N, j = idx & (N - 1);
            const int64_t b = u * PER + s;
            if (b >= p.n) continue;
            const uint32_t cw = *(const uint32_t*)(bufA + s * N * LD + k * LD + j);
            int c0 = 0, c1 = 0;
            const int l0 = tu_quant((int)(int16_t)(cw & 0xffff), __ldg(p.quantCoeff + idx), p.add, p.qBits, c0);
            const int l1 = tu_quant((int)(int16_t)(cw >> 16), __ldg(p.quantCoeff + idx + 1), p.add, p.qBits, c1);
            *(uint32_t*)(p.coeff + b * (N * N) + idx) = (uint32_t)(l0 & 0xffff) | ((uint32_t)l1 << 16);
            const int d0 = tu_dequant(p, l0, idx), d1 = tu_dequant(p, l1, idx + 1);
            bufB[s * N * LD + j * LD + k] = (int16_t)d0;
            bufB[s * N * LD + (j + 1) * LD + k] = (int16_t)d1;
#pragma unroll
            for (int ss = 0; ss < PER; ss++)
                if (ss == s) { cnt[ss] += c0 + c1; if (idx == 0) { dcq[ss] = l0; dcd[ss] = d0; } }
        }
        int mode[PER], dcv[PER];           // 0: zero residual, 1: DC fill, 2: full inverse
        bool anyInv = false;
#pragma unroll
        for (int s = 0; s < PER; s++)
        {
            cnt[s] = warp_sum(cnt[s]);
            // lane 0 handled idx 0 of TU 0 (and of TU 1 when N == 8: pair N*N/2 = 32 -> lane 0 again)
            dcq[s] = __shfl_sync(0xffffffffu, dcq[s], 0); dcd[s] = __shfl_sync(0xffffffffu, dcd[s], 0);
            mode[s] = cnt[s] == 0 ? 0 : ((cnt[s] == 1 && dcq[s] != 0) ? 1 : 2);
            dcv[s] = tu_dc_val(dcd[s], p.depth);
            anyInv = anyInv || mode[s] == 2;
        }
        __syncwarp();
        if (anyInv)
        {
            if (N == 32)      { pass32<true>(ai32, bufB, bufA, 1 << (is1 - 1), is1, gid, tig); __syncwarp(); pass32<true>(ai32, bufA, bufB, 1 << (is2 - 1), is2, gid, tig); }
            else if (N == 16) { pass16<true>(ai16, bufB, bufA, 1 << (is1 - 1), is1, gid, tig); __syncwarp(); pass16<true>(ai16, bufA, bufB, 1 << (is2 - 1), is2, gid, tig); }
            else              { pass8<true>(ai16, bufB, bufA, 1 << (is1 - 1), is1, gid, tig);  __syncwarp(); pass8<true>(ai16, bufA, bufB, 1 << (is2 - 1), is2, gid, tig); }
            __syncwarp();
        }

        // ---- recon = clip(pred + resi'), sse vs fenc; bufB[k][j] = resi'[row j][col k] ----
        unsigned long long sse[PER];
#pragma unroll
        for (int s = 0; s < PER; s++) sse[s] = 0;
#pragma unroll
        for (int i = 0; i < CNT; i++)
        {
            const int e = lane + 32 * i;
            if (e >= WPU) continue;
            const int s = e / WPT, ee = e - s * WPT, row = (ee * PW) >> log2N, col = (ee * PW) & (N - 1);
            const int64_t b = u * PER + s;
            if (b >= p.n) continue;
            int md = mode[0], dv = dcv[0];
#pragma unroll
            for (int ss = 1; ss < PER; ss++) if (ss == s) { md = mode[ss]; dv = dcv[ss]; }
            uint32_t out = 0; unsigned acc = 0;
#pragma unroll
            for (int c = 0; c < PW; c++)
            {
                const int sh = c * (32 / PW), msk = sizeof(pixel) == 1 ? 0xff : 0xffff;
                const int pv = (int)((pw[i] >> sh) & msk), fv = (int)((fw[i] >> sh) & msk);
                const int r = md == 0 ? 0 : (md == 1 ? dv : (int)bufB[s * N * LD + (col + c) * LD + row]);
                const int rec = clip3i(0, maxVal, pv + r);
                out |= (uint32_t)rec << sh;
                acc += (unsigned)((fv - rec) * (fv - rec));
            }
            const int64_t by = b / p.blocksX, bx = b - by * p.blocksX;
            pixel* rp = recon + (by * N + row) * p.reconStride + bx * N + col;
            if (p.wordStores) *(uint32_t*)rp = out;
            else
            {
#pragma unroll
                for (int c = 0; c < PW; c++) rp[c] = (pixel)((out >> (c * (32 / PW))) & (sizeof(pixel) == 1 ? 0xff : 0xffff));
            }
#pragma unroll
            for (int ss = 0; ss < PER; ss++) if (ss == s) sse[ss] += acc;
        }
#pragma unroll
        for (int s = 0; s < PER; s++)
        {
            unsigned lo = (unsigned)sse[s], hi = (unsigned)(sse[s] >> 32);
            // per-lane sums stay below 2^32 (<= 32 px * 4095^2); add across lanes in 64 bits
            unsigned long long tot = sse[s];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
            {
                const unsigned l2 = __shfl_xor_sync(0xffffffffu, (unsigned)tot, o), h2 = __shfl_xor_sync(0xffffffffu, (unsigned)(tot >> 32), o);
                tot += ((unsigned long long)h2 << 32) | l2;
            }
            (void)lo; (void)hi;
            const int64_t b = u * PER + s;
            if (lane == 0 && b < p.n) { p.numSig[b] = (uint32_t)cnt[s]; p.sse[b] = tot; }
        }
        __syncwarp();
    }
}

// 4x4 TUs: one thread per TU, scalar butterflies (DCT or DST)
template<typename pixel>
__global__ void __launch_bounds__(128)
tu4_pipeline_kernel(TuArgs p)
{
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.n) return;
    const int64_t by = b / p.blocksX, bx = b - by * p.blocksX;
    const pixel* fenc = (const pixel*)p.fenc + by * 4 * p.fencStride + bx * 4;
    const pixel* pred = (const pixel*)p.pred + by * 4 * p.predStride + bx * 4;
    pixel* recon = (pixel*)p.recon + by * 4 * p.reconStride + bx * 4;
    const int maxVal = (1 << p.depth) - 1;
    int f[16], q[16], in[16], tmp[16], out[16];
#pragma unroll
    for (int r = 0; r < 4; r++)
    {
        int a[4], c[4];
        ld4i<pixel>(fenc + r * p.fencStride, a); ld4i<pixel>(pred + r * p.predStride, c);
#pragma unroll
        for (int x = 0; x < 4; x++) { f[4 * r + x] = a[x]; q[4 * r + x] = c[x]; in[4 * r + x] = a[x] - c[x]; }
    }
    const bool dst = p.useDST != 0;
    fwd4_pass(in, tmp, 1 + (p.depth - 8), dst); fwd4_pass(tmp, out, 8, dst);          // dct.cpp:444-445 / :461-462
    int cnt = 0, lv0 = 0;
#pragma unroll
    for (int i = 0; i < 16; i++)
    {
        const int l = tu_quant((int)(int16_t)out[i], __ldg(p.quantCoeff + i), p.add, p.qBits, cnt);
        p.coeff[b * 16 + i] = (int16_t)l;
        if (i == 0) lv0 = l;
        in[i] = tu_dequant(p, l, i);
    }
    const int mode = cnt == 0 ? 0 : ((cnt == 1 && lv0 != 0 && !dst) ? 1 : 2);
    if (mode == 2) { inv4_pass(in, tmp, 7, dst); inv4_pass(tmp, out, 12 - (p.depth - 8), dst); }
    const int dcv = tu_dc_val(in[0], p.depth);
    unsigned long long sse = 0;
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int x = 0; x < 4; x++)
        {
            const int rs = mode == 0 ? 0 : (mode == 1 ? dcv : out[4 * r + x]);
            const int rec = clip3i(0, maxVal, q[4 * r + x] + rs);
            recon[r * p.reconStride + x] = (pixel)rec;
            sse += (unsigned)((f[4 * r + x] - rec) * (f[4 * r + x] - rec));
        }
    p.numSig[b] = (uint32_t)cnt; p.sse[b] = sse;
}

int tu_pipeline_dev(Ctx* ctx, int sizeIdx, int depth, int useDST, const void* fenc, int64_t fencStride, const void* pred, int64_t predStride,
                    void* recon, int64_t reconStride, int blocksX, int blocksY, const int32_t* quantCoeff, int qBits, int add,
                    const int32_t* dequantCoef, int scaleOrPer, int dqShift, int16_t* coeff, uint32_t* numSig, uint64_t* sse)
{
    if (sizeIdx < 0 || sizeIdx > 3) { set_error("tu_pipeline: sizeIdx %d (0..3 = 4/8/16/32)", sizeIdx); return -1; }
    if (useDST && sizeIdx != 0) { set_error("tu_pipeline: DST is the 4x4 luma intra transform only"); return -1; }
    if (blocksX <= 0 || blocksY <= 0) return 0;
    if (!fenc || !pred || !recon || !quantCoeff || !coeff || !numSig || !sse) { set_error("tu_pipeline: NULL operand"); return -1; }
    if (qBits < 8 || qBits > 30 || (!dequantCoef && (dqShift < 1 || dqShift > 10))) { set_error("tu_pipeline: qBits %d / shift %d", qBits, dqShift); return -1; }
    if ((uintptr_t)coeff & 3) { set_error("tu_pipeline: coeff must be 4-byte aligned"); return -1; }
    if (upload_tables(ctx)) return -1;
    TuArgs a;
    a.fenc = fenc; a.fencStride = fencStride; a.pred = pred; a.predStride = predStride; a.recon = recon; a.reconStride = reconStride;
    a.blocksX = blocksX; a.n = (int64_t)blocksX * blocksY; a.quantCoeff = quantCoeff; a.qBits = qBits; a.add = add;
    a.dequantCoef = dequantCoef; a.scaleOrPer = scaleOrPer; a.dqShift = dqShift; a.coeff = coeff; a.numSig = numSig; a.sse = sse;
    a.depth = depth; a.useDST = useDST;
    const size_t px = depth > 8 ? 2 : 1;
    a.wordStores = !(((uintptr_t)recon | (uintptr_t)(reconStride * (int64_t)px)) & 3);
    const int N = 4 << sizeIdx;
    if (N == 4)
    {
        const unsigned blocks = (unsigned)((a.n + 127) / 128);
        if (depth > 8) tu4_pipeline_kernel<uint16_t><<<blocks, 128, 0, ctx->stream>>>(a);
        else           tu4_pipeline_kernel<uint8_t><<<blocks, 128, 0, ctx->stream>>>(a);
    }
    else
    {
        const int64_t units = N == 8 ? (a.n + 1) / 2 : a.n;
        const int64_t want = (units + XF_WARPS - 1) / XF_WARPS, cap = (int64_t)ctx->smCount * 4;
        const unsigned blocks = (unsigned)(want < cap ? want : cap);
        dim3 block(XF_WARPS * 32);
#define TU_LAUNCH(PX, NN) tu_pipeline_kernel<PX, NN><<<blocks, block, 0, ctx->stream>>>(a)
        if (depth > 8) { if (N == 32) TU_LAUNCH(uint16_t, 32); else if (N == 16) TU_LAUNCH(uint16_t, 16); else TU_LAUNCH(uint16_t, 8); }
        else           { if (N == 32) TU_LAUNCH(uint8_t, 32);  else if (N == 16) TU_LAUNCH(uint8_t, 16);  else TU_LAUNCH(uint8_t, 8); }
#undef TU_LAUNCH
    }
    ctx->launches++;
    return check(cudaGetLastError(), "tu_pipeline kernel launch");
}

void host_dct_table(int N, int16_t* out)
{
    for (int k = 0; k < N; k++)
        for (int n = 0; n < N; n++) out[k * N + n] = (int16_t)dct_coef(N, k, n);
}

} // namespace x265b200

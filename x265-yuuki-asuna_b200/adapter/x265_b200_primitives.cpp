// x265_b200_primitives.cpp -- the drop-in adapter: provides
//
//     namespace X265_NS { void setupAssemblyPrimitives(EncoderPrimitives& p, int cpuMask); }
//
// exactly as declared by the reference (source/common/primitives.h:470) and called from
// x265_setup_primitives() (source/common/primitives.cpp:264) and TestBench (source/test/testbench.cpp:211).
// It is compiled AGAINST THE REFERENCE'S OWN HEADERS (-I<x265>/source/common) once per bit depth
// (X265_DEPTH / HIGH_BIT_DEPTH / X265_NS as in the x265 build) and linked with libx265b200.so, which
// is all a maintainer has to add to an x265 build configured with ENABLE_ASSEMBLY.
//
// Phase A of SURVEY.md 7 / 8b: every installed slot is a pointer-compatible, synchronous HOST-pointer
// thunk -- stage the operands (H2D), run the same sm_100a kernel the batched API uses, copy the result
// back -- so the unmodified encoder and the unmodified TestBench link it.  These thunks exist for parity
// (a ~30 us round trip per call can never be fast); throughput comes from the batched C ABI.
// There is no CPU fallback: if the CUDA backend cannot start, setupAssemblyPrimitives aborts loudly.
#include "common.h"
#include "primitives.h"
#include "x265b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

using namespace X265_NS;

namespace {

struct ThreadCtx
{
    x265b200_ctx* ctx;
    void* buf[8]; size_t cap[8];
    ~ThreadCtx() { if (ctx) x265b200_destroy(ctx); }
};
thread_local ThreadCtx tl = { nullptr, { 0 }, { 0 } };

[[noreturn]] void die(const char* what)
{
    fprintf(stderr, "x265 [error]: B200 primitive backend: %s: %s\n", what, x265b200_last_error());
    abort();            // encoder-level convention: log + abort (encoder.cpp:332-454); never a silent CPU fallback
}
#define CK(e) do { if ((e) != 0) die(#e); } while (0)

x265b200_ctx* C()
{
    if (!tl.ctx)
    {
        const char* d = getenv("X265B200_DEVICE");
        CK(x265b200_create(d ? atoi(d) : 0, nullptr, &tl.ctx));
    }
    return tl.ctx;
}

void* dev(int slot, size_t bytes)
{
    if (tl.cap[slot] < bytes)
    {
        if (tl.buf[slot]) CK(x265b200_free(C(), tl.buf[slot]));
        size_t cap = bytes * 2 + 256;
        CK(x265b200_malloc(C(), cap, &tl.buf[slot]));
        tl.cap[slot] = cap;
    }
    return tl.buf[slot];
}

// stage a w x h region (elements of `es` bytes, row stride `stride` elements) as ONE contiguous span of
// (h-1)*stride + w elements; the device copy keeps the caller's stride (TestBench uses strides smaller than
// the block width, e.g. FENC_STRIDE-5 for sad_x3/x4, so a pitched 2-D copy is not generally possible).
void* up_span(int slot, const void* src, intptr_t stride, int w, int h, int es)
{
    size_t bytes = ((size_t)(h - 1) * stride + w) * es;
    void* d = dev(slot, bytes);
    CK(x265b200_upload(C(), d, src, bytes));
    return d;
}
void* up1d(int slot, const void* src, size_t bytes)
{
    void* d = dev(slot, bytes);
    CK(x265b200_upload(C(), d, src, bytes));
    return d;
}
void down2d(void* dst, intptr_t stride, const void* d, int w, int h, int es)
{
    if (stride >= w) { CK(x265b200_download2d(C(), dst, (size_t)stride * es, d, (size_t)w * es, (size_t)w * es, h)); return; }
    for (int y = 0; y < h; y++)      // overlapping destination rows: copy in the order the C code writes them
        CK(x265b200_download(C(), (char*)dst + (size_t)y * stride * es, (const char*)d + (size_t)y * w * es, (size_t)w * es));
}

const int PX = (int)sizeof(pixel);

// ---- block compare ------------------------------------------------------------------------------
template<int W, int H, int KIND>
int cmp_thunk(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb)
{
    void* dA = up_span(0, a, sa, W, H, PX);
    void* dB = up_span(1, b, sb, W, H, PX);
    void* dO = dev(2, 8);
    CK(x265b200_pixelcmp_dev(C(), KIND, X265_DEPTH, W, H, dA, sa, dB, sb, nullptr, nullptr, nullptr, 1, 1, dO));
    int32_t r; CK(x265b200_download(C(), &r, dO, 4));
    return r;
}
template<int W, int H>
sse_t sse_pp_thunk(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb)
{
    void* dA = up_span(0, a, sa, W, H, PX);
    void* dB = up_span(1, b, sb, W, H, PX);
    void* dO = dev(2, 8);
    CK(x265b200_pixelcmp_dev(C(), X265B200_CMP_SSE_PP, X265_DEPTH, W, H, dA, sa, dB, sb, nullptr, nullptr, nullptr, 1, 1, dO));
    uint64_t r; CK(x265b200_download(C(), &r, dO, 8));
    return (sse_t)r;
}
template<int W, int H>
sse_t sse_ss_thunk(const int16_t* a, intptr_t sa, const int16_t* b, intptr_t sb)
{
    void* dA = up_span(0, a, sa, W, H, 2);
    void* dB = up_span(1, b, sb, W, H, 2);
    void* dO = dev(2, 8);
    CK(x265b200_pixelcmp_dev(C(), X265B200_CMP_SSE_SS, X265_DEPTH, W, H, dA, sa, dB, sb, nullptr, nullptr, nullptr, 1, 1, dO));
    uint64_t r; CK(x265b200_download(C(), &r, dO, 8));
    return (sse_t)r;
}
template<int N>
sse_t ssd_s_thunk(const int16_t* a, intptr_t sa)
{
    void* dA = up_span(0, a, sa, N, N, 2);
    void* dO = dev(2, 8);
    CK(x265b200_pixelcmp_dev(C(), X265B200_CMP_SSD_S, X265_DEPTH, N, N, dA, sa, dA, sa, nullptr, nullptr, nullptr, 1, 1, dO));
    uint64_t r; CK(x265b200_download(C(), &r, dO, 8));
    return (sse_t)r;
}
template<int W, int H, int K>
void sad_xn(const pixel* fenc, const pixel* const refs[K], intptr_t stride, int32_t* res)
{
    void* dF = up_span(0, fenc, FENC_STRIDE, W, H, PX);                         // 64-pixel pitch kept (pixel.cpp:89)
    const size_t span = (size_t)(H - 1) * stride + W;                           // elements per reference block
    char* dR = (char*)dev(1, (size_t)K * span * PX);
    int64_t off[4];
    for (int k = 0; k < K; k++)
    {
        CK(x265b200_upload(C(), dR + (size_t)k * span * PX, refs[k], span * PX));
        off[k] = (int64_t)k * span;
    }
    void* dOff = up1d(2, off, sizeof(int64_t) * K);
    void* dO = dev(3, 16);
    CK(x265b200_sad_xn_dev(C(), X265_DEPTH, K, W, H, dF, 0, dR, stride, (const int64_t*)dOff, 1, (int32_t*)dO));
    CK(x265b200_download(C(), res, dO, 4 * K));
}
template<int W, int H>
void sad_x3_thunk(const pixel* fenc, const pixel* r0, const pixel* r1, const pixel* r2, intptr_t stride, int32_t* res)
{
    const pixel* refs[3] = { r0, r1, r2 };
    sad_xn<W, H, 3>(fenc, refs, stride, res);
}
template<int W, int H>
void sad_x4_thunk(const pixel* fenc, const pixel* r0, const pixel* r1, const pixel* r2, const pixel* r3, intptr_t stride, int32_t* res)
{
    const pixel* refs[4] = { r0, r1, r2, r3 };
    sad_xn<W, H, 4>(fenc, refs, stride, res);
}

// ---- transforms ---------------------------------------------------------------------------------
template<int IDX, int N>
void dct_thunk(const int16_t* src, int16_t* dst, intptr_t srcStride)
{
    void* dS = up_span(0, src, srcStride, N, N, 2);
    void* dD = dev(1, N * N * 2);
    CK(x265b200_dct_dev(C(), IDX, X265_DEPTH, (const int16_t*)dS, 0, srcStride, (int16_t*)dD, 1));
    CK(x265b200_download(C(), dst, dD, N * N * 2));
}
template<int IDX, int N>
void idct_thunk(const int16_t* src, int16_t* dst, intptr_t dstStride)
{
    void* dS = up1d(0, src, N * N * 2);
    void* dD = dev(1, N * N * 2);
    CK(x265b200_idct_dev(C(), IDX, X265_DEPTH, (const int16_t*)dS, (int16_t*)dD, N * N, N, 1));
    down2d(dst, dstStride, dD, N, N, 2);
}
uint32_t quant_thunk(const int16_t* coef, const int32_t* quantCoeff, int32_t* deltaU, int16_t* qCoef, int qBits, int add, int numCoeff)
{
    void* dC = up1d(0, coef, numCoeff * 2); void* dQ = up1d(1, quantCoeff, numCoeff * 4);
    void* dU = dev(2, numCoeff * 4); void* dO = dev(3, numCoeff * 2); void* dN = dev(4, 4);
    CK(x265b200_quant_dev(C(), (const int16_t*)dC, (const int32_t*)dQ, (int32_t*)dU, (int16_t*)dO, qBits, add, numCoeff, 1, (uint32_t*)dN));
    CK(x265b200_download(C(), deltaU, dU, numCoeff * 4)); CK(x265b200_download(C(), qCoef, dO, numCoeff * 2));
    uint32_t n; CK(x265b200_download(C(), &n, dN, 4));
    return n;
}
uint32_t nquant_thunk(const int16_t* coef, const int32_t* quantCoeff, int16_t* qCoef, int qBits, int add, int numCoeff)
{
    void* dC = up1d(0, coef, numCoeff * 2); void* dQ = up1d(1, quantCoeff, numCoeff * 4);
    void* dO = dev(3, numCoeff * 2); void* dN = dev(4, 4);
    CK(x265b200_nquant_dev(C(), (const int16_t*)dC, (const int32_t*)dQ, (int16_t*)dO, qBits, add, numCoeff, 1, (uint32_t*)dN));
    CK(x265b200_download(C(), qCoef, dO, numCoeff * 2));
    uint32_t n; CK(x265b200_download(C(), &n, dN, 4));
    return n;
}
void dequant_normal_thunk(const int16_t* q, int16_t* coef, int num, int scale, int shift)
{
    void* dQ = up1d(0, q, num * 2); void* dO = dev(1, num * 2);
    CK(x265b200_dequant_normal_dev(C(), (const int16_t*)dQ, (int16_t*)dO, num, 1, scale, shift));
    CK(x265b200_download(C(), coef, dO, num * 2));
}
void dequant_scaling_thunk(const int16_t* q, const int32_t* deq, int16_t* coef, int num, int per, int shift)
{
    void* dQ = up1d(0, q, num * 2); void* dD = up1d(2, deq, num * 4); void* dO = dev(1, num * 2);
    CK(x265b200_dequant_scaling_dev(C(), (const int16_t*)dQ, (const int32_t*)dD, (int16_t*)dO, num, 1, per, shift));
    CK(x265b200_download(C(), coef, dO, num * 2));
}
template<int N>
int count_nonzero_thunk(const int16_t* q)
{
    void* dQ = up1d(0, q, N * N * 2); void* dO = dev(1, 4);
    CK(x265b200_count_nonzero_dev(C(), (const int16_t*)dQ, N * N, 1, (int32_t*)dO));
    int32_t n; CK(x265b200_download(C(), &n, dO, 4));
    return n;
}

// ---- interpolation -----------------------------------------------------------------------------
// stages exactly the footprint the reference function reads (SURVEY.md 8a "interpolation read footprint")
template<int TAPS, int W, int H, int KIND>
void interp_run(const void* src, intptr_t srcStride, void* dst, intptr_t dstStride, int idxX, int idxY, int isRowExt)
{
    const bool srcShort = (KIND == X265B200_IP_VSP || KIND == X265B200_IP_VSS);
    const bool dstShort = (KIND == X265B200_IP_HPS || KIND == X265B200_IP_VPS || KIND == X265B200_IP_VSS || KIND == X265B200_IP_P2S);
    const bool horiz = (KIND == X265B200_IP_HPP || KIND == X265B200_IP_HPS || KIND == X265B200_IP_HVPP);
    const bool vert = (KIND == X265B200_IP_VPP || KIND == X265B200_IP_VPS || KIND == X265B200_IP_VSP || KIND == X265B200_IP_VSS || KIND == X265B200_IP_HVPP ||
                       (KIND == X265B200_IP_HPS && isRowExt));
    const int half = TAPS / 2 - 1;
    const int left = horiz ? half : 0, top = vert ? half : 0;
    const int rw = W + (horiz ? TAPS - 1 : 0), rh = H + (vert ? TAPS - 1 : 0);
    const int ses = srcShort ? 2 : PX, des = dstShort ? 2 : PX;
    const char* origin = (const char*)src - ((intptr_t)top * srcStride + left) * ses;
    void* dS = up_span(0, origin, srcStride, rw, rh, ses);
    const int outRows = H + ((KIND == X265B200_IP_HPS && isRowExt) ? TAPS - 1 : 0);
    void* dD = dev(1, (size_t)W * outRows * des);
    x265b200_interp_job job; job.srcOff = (int64_t)top * srcStride + left; job.dstOff = 0; job.idxX = idxX; job.idxY = idxY;
    void* dJ = up1d(2, &job, sizeof(job));
    CK(x265b200_interp_dev(C(), KIND, TAPS, X265_DEPTH, W, H, dS, srcStride, dD, W, (const x265b200_interp_job*)dJ, 1, isRowExt));
    down2d(dst, dstStride, dD, W, outRows, des);
}
template<int T, int W, int H, int K> void ip_pp(const pixel* s, intptr_t ss, pixel* d, intptr_t ds, int c) { interp_run<T, W, H, K>(s, ss, d, ds, c, 0, 0); }
template<int T, int W, int H> void ip_hps(const pixel* s, intptr_t ss, int16_t* d, intptr_t ds, int c, int ext) { interp_run<T, W, H, X265B200_IP_HPS>(s, ss, d, ds, c, 0, ext); }
template<int T, int W, int H> void ip_vps(const pixel* s, intptr_t ss, int16_t* d, intptr_t ds, int c) { interp_run<T, W, H, X265B200_IP_VPS>(s, ss, d, ds, c, 0, 0); }
template<int T, int W, int H> void ip_vsp(const int16_t* s, intptr_t ss, pixel* d, intptr_t ds, int c) { interp_run<T, W, H, X265B200_IP_VSP>(s, ss, d, ds, c, 0, 0); }
template<int T, int W, int H> void ip_vss(const int16_t* s, intptr_t ss, int16_t* d, intptr_t ds, int c) { interp_run<T, W, H, X265B200_IP_VSS>(s, ss, d, ds, c, 0, 0); }
template<int W, int H> void ip_hvpp(const pixel* s, intptr_t ss, pixel* d, intptr_t ds, int cx, int cy) { interp_run<8, W, H, X265B200_IP_HVPP>(s, ss, d, ds, cx, cy, 0); }
template<int W, int H> void ip_p2s(const pixel* s, intptr_t ss, int16_t* d, intptr_t ds) { interp_run<8, W, H, X265B200_IP_P2S>(s, ss, d, ds, 0, 0, 0); }

// ---- intra ---------------------------------------------------------------------------------------
// SLOT: 0 = intra_pred[PLANAR_IDX], 1 = intra_pred[DC_IDX] (both ignore dirMode, intrapred.cpp:69-100 -- TestBench
// calls them with dirMode = 0), 2 = the angular slots (mode = dirMode)
template<int LOG2, int SLOT>
void intra_pred_thunk(pixel* dst, intptr_t dstStride, const pixel* srcPix, int dirMode, int bFilter)
{
    const int N = 1 << LOG2;
    if (SLOT < 2) dirMode = SLOT;
    void* dS = up1d(0, srcPix, (4 * N + 1) * PX);
    void* dD = dev(1, N * N * PX);
    x265b200_intra_job job; job.srcOff = 0; job.dstOff = 0; job.mode = dirMode; job.bFilter = bFilter;
    void* dJ = up1d(2, &job, sizeof(job));
    CK(x265b200_intra_pred_dev(C(), X265_DEPTH, LOG2, dS, dD, N, (const x265b200_intra_job*)dJ, 1));
    down2d(dst, dstStride, dD, N, N, PX);
}
template<int LOG2>
void intra_filter_thunk(const pixel* ref, pixel* filtered)
{
    const int N = 1 << LOG2;
    void* dS = up1d(0, ref, (4 * N + 1) * PX); void* dD = dev(1, (4 * N + 1) * PX);
    CK(x265b200_intra_filter_dev(C(), X265_DEPTH, LOG2, dS, dD, 1));
    CK(x265b200_download(C(), filtered, dD, (4 * N + 1) * PX));
}
template<int LOG2>
void intra_allangs_thunk(pixel* dest, pixel* refPix, pixel* filtPix, int bLuma)
{
    const int N = 1 << LOG2;
    void* dR = up1d(0, refPix, (4 * N + 1) * PX); void* dF = up1d(2, filtPix, (4 * N + 1) * PX);
    void* dD = dev(1, 33 * N * N * PX);
    CK(x265b200_intra_allangs_dev(C(), X265_DEPTH, LOG2, dR, dF, dD, bLuma, 1));
    CK(x265b200_download(C(), dest, dD, 33 * N * N * PX));
}

// ---- glue (class G of SURVEY.md 8a-bis): one-job launches of the generic element-wise kernel ---------------------
// e0/e1/ed: element sizes of src0 / src1 / dst
void glue_run(int op, int w, int h, void* dst, intptr_t dstStride, int ed, const void* s0, intptr_t s0Stride, int e0,
              const void* s1, intptr_t s1Stride, int e1, int p0 = 0, int p1 = 0, int p2 = 0, int p3 = 0, int dw = -1, int dh = -1)
{
    if (dw < 0) { dw = w; dh = h; }
    void* d0 = s0 ? up_span(0, s0, s0Stride, op == X265B200_GL_TRANSPOSE ? h : w, op == X265B200_GL_TRANSPOSE ? w : h, e0) : nullptr;
    void* d1 = s1 ? up_span(1, s1, s1Stride, w, h, e1) : nullptr;
    void* dD = dev(2, (size_t)dw * dh * ed);
    x265b200_glue_job job = { 0, 0, 0 };
    void* dJ = up1d(3, &job, sizeof(job));
    CK(x265b200_glue_dev(C(), op, X265_DEPTH, w, h, dD, dw, d0, s0Stride, d1, s1Stride, (const x265b200_glue_job*)dJ, 1, p0, p1, p2, p3));
    down2d(dst, dstStride, dD, dw, dh, ed);
}
template<int W, int H> void copy_pp_thunk(pixel* d, intptr_t ds, const pixel* s, intptr_t ss) { glue_run(X265B200_GL_COPY_PP, W, H, d, ds, PX, s, ss, PX, nullptr, 0, 0); }
template<int W, int H> void copy_sp_thunk(pixel* d, intptr_t ds, const int16_t* s, intptr_t ss) { glue_run(X265B200_GL_COPY_SP, W, H, d, ds, PX, s, ss, 2, nullptr, 0, 0); }
template<int W, int H> void copy_ps_thunk(int16_t* d, intptr_t ds, const pixel* s, intptr_t ss) { glue_run(X265B200_GL_COPY_PS, W, H, d, ds, 2, s, ss, PX, nullptr, 0, 0); }
template<int W, int H> void copy_ss_thunk(int16_t* d, intptr_t ds, const int16_t* s, intptr_t ss) { glue_run(X265B200_GL_COPY_SS, W, H, d, ds, 2, s, ss, 2, nullptr, 0, 0); }
template<int N> void blockfill_s_thunk(int16_t* d, intptr_t ds, int16_t val) { glue_run(X265B200_GL_FILL_S, N, N, d, ds, 2, nullptr, 0, 0, nullptr, 0, 0, val); }
template<int N> void cpy2Dto1D_shl_thunk(int16_t* d, const int16_t* s, intptr_t ss, int shift) { glue_run(X265B200_GL_CPY2DTO1D_SHL, N, N, d, N, 2, s, ss, 2, nullptr, 0, 0, shift); }
template<int N> void cpy2Dto1D_shr_thunk(int16_t* d, const int16_t* s, intptr_t ss, int shift) { glue_run(X265B200_GL_CPY2DTO1D_SHR, N, N, d, N, 2, s, ss, 2, nullptr, 0, 0, shift); }
template<int N> void cpy1Dto2D_shl_thunk(int16_t* d, const int16_t* s, intptr_t ds, int shift) { glue_run(X265B200_GL_CPY1DTO2D_SHL, N, N, d, ds, 2, s, N, 2, nullptr, 0, 0, shift); }
template<int N> void cpy1Dto2D_shr_thunk(int16_t* d, const int16_t* s, intptr_t ds, int shift) { glue_run(X265B200_GL_CPY1DTO2D_SHR, N, N, d, ds, 2, s, N, 2, nullptr, 0, 0, shift); }
template<int N> void calcresidual_thunk(const pixel* fenc, const pixel* pred, int16_t* resi, intptr_t stride) { glue_run(X265B200_GL_SUB_PS, N, N, resi, stride, 2, fenc, stride, PX, pred, stride, PX); }
template<int W, int H> void sub_ps_thunk(int16_t* d, intptr_t ds, const pixel* a, const pixel* b, intptr_t sa, intptr_t sb) { glue_run(X265B200_GL_SUB_PS, W, H, d, ds, 2, a, sa, PX, b, sb, PX); }
template<int W, int H> void add_ps_thunk(pixel* d, intptr_t ds, const pixel* a, const int16_t* b, intptr_t sa, intptr_t sb) { glue_run(X265B200_GL_ADD_PS, W, H, d, ds, PX, a, sa, PX, b, sb, 2); }
template<int W, int H> void pixelavg_pp_thunk(pixel* d, intptr_t ds, const pixel* a, intptr_t sa, const pixel* b, intptr_t sb, int) { glue_run(X265B200_GL_PIXELAVG_PP, W, H, d, ds, PX, a, sa, PX, b, sb, PX); }
template<int W, int H> void addAvg_thunk(const int16_t* a, const int16_t* b, pixel* d, intptr_t sa, intptr_t sb, intptr_t ds) { glue_run(X265B200_GL_ADDAVG, W, H, d, ds, PX, a, sa, 2, b, sb, 2); }
template<int N> void transpose_thunk(pixel* d, const pixel* s, intptr_t stride) { glue_run(X265B200_GL_TRANSPOSE, N, N, d, N, PX, s, stride, PX, nullptr, 0, 0); }
void weight_pp_thunk(const pixel* s, pixel* d, intptr_t stride, int w, int h, int w0, int round, int shift, int offset)
{
    glue_run(X265B200_GL_WEIGHT_PP, w, h, d, stride, PX, s, stride, PX, nullptr, 0, 0, w0, round, shift, offset);
}
void weight_sp_thunk(const int16_t* s, pixel* d, intptr_t ss, intptr_t ds, int w, int h, int w0, int round, int shift, int offset)
{
    glue_run(X265B200_GL_WEIGHT_SP, w, h, d, ds, PX, s, ss, 2, nullptr, 0, 0, w0, round, shift, offset);
}
template<int N> uint64_t var_thunk(const pixel* pix, intptr_t stride)
{
    void* dS = up_span(0, pix, stride, N, N, PX);
    int64_t zero = 0; void* dOff = up1d(1, &zero, 8); void* dO = dev(2, 8);
    CK(x265b200_var_dev(C(), X265_DEPTH, N, dS, stride, (const int64_t*)dOff, 1, (uint64_t*)dO));
    uint64_t r; CK(x265b200_download(C(), &r, dO, 8));
    return r;
}
template<int N> int psy_cost_pp_thunk(const pixel* src, intptr_t ss, const pixel* rec, intptr_t rs)
{
    void* dS = up_span(0, src, ss, N, N, PX); void* dR = up_span(1, rec, rs, N, N, PX);
    int64_t zero = 0; void* dOff = up1d(3, &zero, 8); void* dO = dev(2, 8);
    CK(x265b200_psy_cost_dev(C(), X265_DEPTH, N, dS, ss, dR, rs, (const int64_t*)dOff, (const int64_t*)dOff, 1, (int32_t*)dO));
    int32_t r; CK(x265b200_download(C(), &r, dO, 4));
    return r;
}
template<int N> uint32_t copy_cnt_thunk(int16_t* coeff, const int16_t* resi, intptr_t stride)
{
    void* dS = up_span(0, resi, stride, N, N, 2);
    int64_t zero = 0; void* dOff = up1d(3, &zero, 8); void* dC = dev(1, N * N * 2); void* dO = dev(2, 8);
    CK(x265b200_copy_cnt_dev(C(), N, (int16_t*)dC, (const int16_t*)dS, stride, (const int64_t*)dOff, 1, (uint32_t*)dO));
    CK(x265b200_download(C(), coeff, dC, N * N * 2));
    uint32_t r; CK(x265b200_download(C(), &r, dO, 4));
    return r;
}
void denoise_dct_thunk(int16_t* coef, uint32_t* resSum, const uint16_t* offset, int numCoeff)
{
    void* dC = up1d(0, coef, numCoeff * 2); void* dR = up1d(1, resSum, numCoeff * 4); void* dF = up1d(2, offset, numCoeff * 2);
    CK(x265b200_denoise_dct_dev(C(), (int16_t*)dC, (uint32_t*)dR, (const uint16_t*)dF, numCoeff, 1));
    CK(x265b200_download(C(), coef, dC, numCoeff * 2)); CK(x265b200_download(C(), resSum, dR, numCoeff * 4));
}
template<int IDX, int N>
void lowpass_dct_thunk(const int16_t* src, int16_t* dst, intptr_t srcStride)
{
    void* dS = up_span(0, src, srcStride, N, N, 2);
    void* dD = dev(1, N * N * 2);
    CK(x265b200_lowpass_dct_dev(C(), IDX, X265_DEPTH, (const int16_t*)dS, 0, srcStride, (int16_t*)dD, 1));
    CK(x265b200_download(C(), dst, dD, N * N * 2));
}

// ---- --me sea support (pixel.cpp:121-165, framefilter.cpp:39-140) ---------------------------------------------------
template<int KIND, int LXH>
int ads_thunk(int encDC[], uint32_t* sums, int delta, uint16_t* costMvX, int16_t* mvs, int width, int thresh)
{
    if (width <= 0) return 0;
    const size_t seg = (size_t)width + LXH;                 // elements read from each of the (up to) two plane rows
    uint32_t* dS = (uint32_t*)dev(0, 2 * seg * 4);
    CK(x265b200_upload(C(), dS, sums, seg * 4));
    if (KIND != 1) CK(x265b200_upload(C(), dS + seg, sums + delta, seg * 4));
    void* dC = up1d(1, costMvX, (size_t)width * 2);
    x265b200_ads_job job; job.sumsOff = 0; job.thresh = thresh;
    for (int k = 0; k < 4; k++) job.encDC[k] = k < KIND ? encDC[k] : 0;
    void* dJ = up1d(2, &job, sizeof(job));
    void* dM = dev(3, (size_t)width * 2); void* dN = dev(4, 4);
    CK(x265b200_ads_dev(C(), KIND, LXH, dS, (int64_t)seg, (const uint16_t*)dC, width, (const x265b200_ads_job*)dJ, 1, (int16_t*)dM, (int32_t*)dN));
    int32_t n; CK(x265b200_download(C(), &n, dN, 4));
    if (n > 0) CK(x265b200_download(C(), mvs, dM, (size_t)n * 2));
    return n;
}
template<int W>
void integral_inith_thunk(uint32_t* sum, pixel* pix, intptr_t stride)
{
    if (stride <= W) return;
    uint32_t* dS = (uint32_t*)dev(0, (size_t)2 * stride * 4);
    CK(x265b200_upload(C(), dS, sum - stride, (size_t)stride * 4));
    void* dP = up1d(1, pix, (size_t)stride * PX);
    CK(x265b200_integral_inith_dev(C(), X265_DEPTH, W, dS + stride, dP, stride));
    CK(x265b200_download(C(), sum, dS + stride, (size_t)(stride - W) * 4));
}
template<int H>
void integral_initv_thunk(uint32_t* sum, intptr_t stride)
{
    if (stride <= 0) return;
    uint32_t* dS = (uint32_t*)dev(0, (size_t)(H + 1) * stride * 4);
    CK(x265b200_upload(C(), dS, sum, (size_t)stride * 4));
    CK(x265b200_upload(C(), dS + (size_t)H * stride, sum + (size_t)H * stride, (size_t)stride * 4));
    CK(x265b200_integral_initv_dev(C(), H, dS, stride));
    CK(x265b200_download(C(), sum, dS, (size_t)stride * 4));
}

// ---- lowres / borders / pre-scale (pixel.cpp:559-628, ipfilter.cpp:59-77) ------------------------------------------------
void frame_init_lowres_thunk(const pixel* src0, pixel* dst0, pixel* dsth, pixel* dstv, pixel* dstc, intptr_t srcStride, intptr_t dstStride,
                             int width, int height)
{
    if (width <= 0 || height <= 0) return;
    void* dS = up_span(0, src0, srcStride, 2 * width + 1, 2 * height + 1, PX);       // the C loop reads src0[2x+2] and the row 2y+2
    void* planes[4];
    for (int k = 0; k < 4; k++) planes[k] = dev(1 + k, (size_t)width * height * PX);
    CK(x265b200_lowres_init_dev(C(), X265_DEPTH, dS, srcStride, planes, width, width, height, 0, 0));
    pixel* dst[4] = { dst0, dsth, dstv, dstc };
    for (int k = 0; k < 4; k++) down2d(dst[k], dstStride, planes[k], width, height, PX);
}
void extend_row_border_thunk(pixel* txt, intptr_t stride, int width, int height, int marginX)
{
    if (height <= 0 || marginX <= 0) return;
    const size_t span = (size_t)(height - 1) * stride + width + 2 * (size_t)marginX;
    char* d = (char*)dev(0, span * PX);
    CK(x265b200_upload(C(), d, txt - marginX, span * PX));
    CK(x265b200_extend_border_dev(C(), X265_DEPTH, d + (size_t)marginX * PX, stride, width, height, marginX, 0));
    // only what extendCURowColBorder writes (ipfilter.cpp:59-77) travels back: the left and right margins of each row.  The
    // picture columns in between belong to other rows' worker threads (WPP / frame filter) and must not be rewritten.
    CK(x265b200_download2d(C(), txt - marginX, (size_t)stride * PX, d, (size_t)stride * PX, (size_t)marginX * PX, height));
    CK(x265b200_download2d(C(), txt + width, (size_t)stride * PX, d + (size_t)(marginX + width) * PX, (size_t)stride * PX, (size_t)marginX * PX, height));
}
void scale1D_thunk(pixel* dst, const pixel* src)
{
    void* dS = up1d(0, src, 256 * PX); void* dD = dev(1, 128 * PX);
    x265b200_glue_job job = { 0, 0, 0 };
    void* dJ = up1d(3, &job, sizeof(job));
    CK(x265b200_glue_dev(C(), X265B200_GL_SCALE1D_128TO64, X265_DEPTH, 128, 1, dD, 128, dS, 256, nullptr, 0, (const x265b200_glue_job*)dJ, 1, 0, 0, 0, 0));
    CK(x265b200_download(C(), dst, dD, 128 * PX));
}
void scale2D_thunk(pixel* dst, const pixel* src, intptr_t stride)
{
    void* dS = up_span(0, src, stride, 64, 64, PX); void* dD = dev(1, 32 * 32 * PX);
    x265b200_glue_job job = { 0, 0, 0 };
    void* dJ = up1d(3, &job, sizeof(job));
    CK(x265b200_glue_dev(C(), X265B200_GL_SCALE2D_64TO32, X265_DEPTH, 32, 32, dD, 32, dS, stride, nullptr, 0, (const x265b200_glue_job*)dJ, 1, 0, 0, 0, 0));
    CK(x265b200_download(C(), dst, dD, 32 * 32 * PX));
}

// ---- SAO / deblock / sign (loopfilter.cpp:39-180, sao.cpp:1762-1926) -------------------------------------------------
x265b200_sao_job sao_job(int64_t recOff, int64_t buf0, int64_t buf1, int width, int height, int startX)
{
    x265b200_sao_job j; memset(&j, 0, sizeof(j));
    j.recOff = recOff; j.buf0 = buf0; j.buf1 = buf1; j.width = width; j.height = height; j.startX = startX;
    return j;
}
// runs one apply job on a staged copy of rec[0, ext) and of the sign buffers.  Only what the C loop writes travels back
// (loopfilter.cpp:45-139): columns [x0, x1) of `height` rows of rec, and bytes [w0, w1) of hostBuf0 -- the neighbour rows and
// columns that were staged for reading belong to other CTUs' worker threads and are never rewritten.
void sao_apply_run(int kind, pixel* rec, size_t ext, intptr_t stride, const int8_t* offsets, int nOffsets, int width, int height, int startX,
                   int x0, int x1, int8_t* hostBuf0, size_t bytes0, size_t w0, size_t w1, const int8_t* hostBuf1, size_t bytes1)
{
    char* dR = (char*)up1d(0, rec, ext * PX);
    const size_t off1 = (bytes0 + 15) & ~(size_t)15;
    int8_t* dB = (int8_t*)dev(1, off1 + bytes1 + 16);
    if (bytes0) CK(x265b200_upload(C(), dB, hostBuf0, bytes0));
    if (bytes1) CK(x265b200_upload(C(), dB + off1, hostBuf1, bytes1));
    void* dO = up1d(2, offsets, nOffsets);
    x265b200_sao_job job = sao_job(0, 0, (int64_t)off1, width, height, startX);
    void* dJ = up1d(3, &job, sizeof(job));
    CK(x265b200_sao_apply_dev(C(), kind, X265_DEPTH, dR, stride, (const x265b200_sao_job*)dJ, 1, dB, (const int8_t*)dO, width));
    if (x1 > x0)
    {
        // (a single row may be longer than the stride: TestBench calls saoCuOrgE3 with endX = 64 at stride 32)
        if (height > 1 && stride >= x1 - x0)
            CK(x265b200_download2d(C(), rec + x0, (size_t)stride * PX, dR + (size_t)x0 * PX, (size_t)stride * PX, (size_t)(x1 - x0) * PX, height));
        else
            for (int y = 0; y < height; y++)
                CK(x265b200_download(C(), rec + (size_t)y * stride + x0, dR + ((size_t)y * stride + x0) * PX, (size_t)(x1 - x0) * PX));
    }
    if (w1 > w0) CK(x265b200_download(C(), hostBuf0 + w0, dB + w0, w1 - w0));
}
void saoE0_thunk(pixel* rec, int8_t* offsetEo, int width, int8_t* signLeft, intptr_t stride)
{
    sao_apply_run(X265B200_SAO_E0, rec, (size_t)stride + width + 1, stride, offsetEo, 5, width, 2, 0, 0, width, signLeft, 2, 0, 0, nullptr, 0);
}
void saoE1_thunk(pixel* rec, int8_t* upBuff1, int8_t* offsetEo, intptr_t stride, int width)
{
    sao_apply_run(X265B200_SAO_E1, rec, (size_t)stride + width, stride, offsetEo, 5, width, 1, 0, 0, width, upBuff1, width, 0, width, nullptr, 0);
}
void saoE1_2rows_thunk(pixel* rec, int8_t* upBuff1, int8_t* offsetEo, intptr_t stride, int width)
{
    sao_apply_run(X265B200_SAO_E1_2ROWS, rec, (size_t)2 * stride + width, stride, offsetEo, 5, width, 2, 0, 0, width, upBuff1, width, 0, width, nullptr, 0);
}
void saoE2_thunk(pixel* rec, int8_t* bufft, int8_t* buff1, int8_t* offsetEo, int width, intptr_t stride)
{
    // bufft[x + 1] is written for x in [0, width)
    sao_apply_run(X265B200_SAO_E2, rec, (size_t)stride + width + 1, stride, offsetEo, 5, width, 1, 0, 0, width, bufft, (size_t)width + 1, 1, (size_t)width + 1, buff1, width);
}
void saoE3_thunk(pixel* rec, int8_t* upBuff1, int8_t* offsetEo, intptr_t stride, int startX, int endX)
{
    if (endX <= startX + 1) return;
    // x runs over (startX, endX): rec[x] and upBuff1[x - 1] are written
    sao_apply_run(X265B200_SAO_E3, rec, (size_t)stride + endX, stride, offsetEo, 5, endX, 1, startX, startX + 1, endX, upBuff1, endX, startX, (size_t)endX - 1, nullptr, 0);
}
void saoB0_thunk(pixel* rec, const int8_t* offsetBo, int ctuWidth, int ctuHeight, intptr_t stride)
{
    if (ctuWidth <= 0 || ctuHeight <= 0) return;
    sao_apply_run(X265B200_SAO_B0, rec, (size_t)(ctuHeight - 1) * stride + ctuWidth, stride, offsetBo, 32, ctuWidth, ctuHeight, 0, 0, ctuWidth, nullptr, 0, 0, 0, nullptr, 0);
}

void sao_stats_run(int kind, const int16_t* diff, const pixel* rec, intptr_t stride, int8_t* up1, int8_t* upt, int endX, int endY,
                   int32_t* stats, int32_t* count)
{
    if (endX <= 0 || endY <= 0) return;
    const int ncls = kind == X265B200_SAO_BO ? 32 : 5;
    void* dD = up1d(0, diff, ((size_t)(endY - 1) * 64 + endX) * 2);
    void* dR = up1d(1, rec - 1, ((size_t)endY * stride + endX + 2) * PX);
    const size_t seg = ((size_t)endX + 2 + 15) & ~(size_t)15;               // bytes [-1, endX] of each sign buffer
    int8_t* dB = (int8_t*)dev(2, 2 * seg + 16);
    if (up1) CK(x265b200_upload(C(), dB, up1 - 1, (size_t)endX + 2));
    if (upt) CK(x265b200_upload(C(), dB + seg, upt - 1, (size_t)endX + 2));
    int32_t* dS = (int32_t*)dev(3, 64 * 4);
    CK(x265b200_upload(C(), dS, stats, ncls * 4)); CK(x265b200_upload(C(), dS + 32, count, ncls * 4));
    x265b200_sao_job job = sao_job(1, 1, (int64_t)seg + 1, endX, endY, 0);
    void* dJ = up1d(4, &job, sizeof(job));
    CK(x265b200_sao_stats_dev(C(), kind, X265_DEPTH, (const int16_t*)dD, dR, stride, (const x265b200_sao_job*)dJ, 1, dB, dS, dS + 32));
    CK(x265b200_download(C(), stats, dS, ncls * 4)); CK(x265b200_download(C(), count, dS + 32, ncls * 4));
    if (up1 && kind != X265B200_SAO_E0 && kind != X265B200_SAO_BO) CK(x265b200_download(C(), up1 - 1, dB, (size_t)endX + 2));
    if (upt) CK(x265b200_download(C(), upt - 1, dB + seg, (size_t)endX + 2));
}
void saoStatsBO_thunk(const int16_t* diff, const pixel* rec, intptr_t stride, int endX, int endY, int32_t* stats, int32_t* count)
{ sao_stats_run(X265B200_SAO_BO, diff, rec, stride, nullptr, nullptr, endX, endY, stats, count); }
void saoStatsE0_thunk(const int16_t* diff, const pixel* rec, intptr_t stride, int endX, int endY, int32_t* stats, int32_t* count)
{ sao_stats_run(X265B200_SAO_E0, diff, rec, stride, nullptr, nullptr, endX, endY, stats, count); }
void saoStatsE1_thunk(const int16_t* diff, const pixel* rec, intptr_t stride, int8_t* upBuff1, int endX, int endY, int32_t* stats, int32_t* count)
{ sao_stats_run(X265B200_SAO_E1, diff, rec, stride, upBuff1, nullptr, endX, endY, stats, count); }
void saoStatsE2_thunk(const int16_t* diff, const pixel* rec, intptr_t stride, int8_t* upBuff1, int8_t* upBufft, int endX, int endY, int32_t* stats, int32_t* count)
{ sao_stats_run(X265B200_SAO_E2, diff, rec, stride, upBuff1, upBufft, endX, endY, stats, count); }
void saoStatsE3_thunk(const int16_t* diff, const pixel* rec, intptr_t stride, int8_t* upBuff1, int endX, int endY, int32_t* stats, int32_t* count)
{ sao_stats_run(X265B200_SAO_E3, diff, rec, stride, upBuff1, nullptr, endX, endY, stats, count); }
void sign_thunk(int8_t* dst, const pixel* src1, const pixel* src2, const int endX)
{
    if (endX <= 0) return;
    void* d1 = up1d(0, src1, (size_t)endX * PX); void* d2 = up1d(1, src2, (size_t)endX * PX); void* dD = dev(2, endX);
    CK(x265b200_sign_dev(C(), X265_DEPTH, (int8_t*)dD, d1, d2, endX));
    CK(x265b200_download(C(), dst, dD, endX));
}
void deblock_run(int chroma, pixel* src, intptr_t srcStep, intptr_t offset, int lowTap, int highTap, int32_t a, int32_t b, int32_t c)
{
    intptr_t lo = 0, hi = 0;
    for (int k = 0; k < 4; k++)
        for (int m = lowTap; m <= highTap; m++)
        {
            const intptr_t o = k * srcStep + m * offset;
            lo = o < lo ? o : lo; hi = o > hi ? o : hi;
        }
    const size_t ext = (size_t)(hi - lo + 1);
    void* dP = up1d(0, src + lo, ext * PX);
    x265b200_deblock_job job; memset(&job, 0, sizeof(job));
    job.srcOff = -lo; job.srcStep = srcStep; job.offset = offset; job.tcP = a; job.tcQ = b; job.maskQ = c;
    void* dJ = up1d(1, &job, sizeof(job));
    CK(x265b200_deblock_dev(C(), chroma, X265_DEPTH, dP, (const x265b200_deblock_job*)dJ, 1));
    CK(x265b200_download(C(), src + lo, dP, ext * PX));
}
void pel_filter_luma_thunk(pixel* src, intptr_t srcStep, intptr_t offset, int32_t tcP, int32_t tcQ) { deblock_run(0, src, srcStep, offset, -4, 3, tcP, tcQ, 0); }
void pel_filter_chroma_thunk(pixel* src, intptr_t srcStep, intptr_t offset, int32_t tc, int32_t maskP, int32_t maskQ) { deblock_run(1, src, srcStep, offset, -2, 1, tc, maskP, maskQ); }

// ---- cuTree / ingest / ssim-rd (pixel.cpp:864-994) -----------------------------------------------------------------------
void propagate_cost_thunk(int* dst, const uint16_t* propagateIn, const int32_t* intraCosts, const uint16_t* interCosts,
                          const int32_t* invQscales, const double* fpsFactor, int len)
{
    if (len <= 0) return;
    void* dP = up1d(0, propagateIn, (size_t)len * 2); void* dI = up1d(1, intraCosts, (size_t)len * 4);
    void* dE = up1d(2, interCosts, (size_t)len * 2);  void* dQ = up1d(3, invQscales, (size_t)len * 4);
    void* dD = dev(4, (size_t)len * 4);
    CK(x265b200_propagate_cost_dev(C(), (int*)dD, (const uint16_t*)dP, (const int32_t*)dI, (const uint16_t*)dE, (const int32_t*)dQ, *fpsFactor, len));
    CK(x265b200_download(C(), dst, dD, (size_t)len * 4));
}
void fix8_pack_thunk(uint16_t* dst, double* src, int count)
{
    if (count <= 0) return;
    void* dS = up1d(0, src, (size_t)count * 8); void* dD = dev(1, (size_t)count * 2);
    CK(x265b200_fix8_pack_dev(C(), (uint16_t*)dD, (const double*)dS, count));
    CK(x265b200_download(C(), dst, dD, (size_t)count * 2));
}
void fix8_unpack_thunk(double* dst, uint16_t* src, int count)
{
    if (count <= 0) return;
    void* dS = up1d(0, src, (size_t)count * 2); void* dD = dev(1, (size_t)count * 8);
    CK(x265b200_fix8_unpack_dev(C(), (double*)dD, (const uint16_t*)dS, count));
    CK(x265b200_download(C(), dst, dD, (size_t)count * 8));
}
void planecopy_run(int mode, const void* src, intptr_t srcStride, int ses, pixel* dst, intptr_t dstStride, int width, int height, int shift, int mask)
{
    if (width <= 0 || height <= 0) return;
    void* dS = up_span(0, src, srcStride, width, height, ses);
    void* dD = dev(1, (size_t)width * height * PX);
    CK(x265b200_planecopy_dev(C(), mode, X265_DEPTH, dS, srcStride, dD, width, width, height, shift, mask));
    down2d(dst, dstStride, dD, width, height, PX);
}
void planecopy_cp_thunk(const uint8_t* src, intptr_t srcStride, pixel* dst, intptr_t dstStride, int width, int height, int shift)
{ planecopy_run(X265B200_PC_CP, src, srcStride, 1, dst, dstStride, width, height, shift, 0); }
void planecopy_sp_thunk(const uint16_t* src, intptr_t srcStride, pixel* dst, intptr_t dstStride, int width, int height, int shift, uint16_t mask)
{ planecopy_run(X265B200_PC_SP, src, srcStride, 2, dst, dstStride, width, height, shift, mask); }
void planecopy_sp_shl_thunk(const uint16_t* src, intptr_t srcStride, pixel* dst, intptr_t dstStride, int width, int height, int shift, uint16_t mask)
{ planecopy_run(X265B200_PC_SP_SHL, src, srcStride, 2, dst, dstStride, width, height, shift, mask); }
void planecopy_pp_shr_thunk(const pixel* src, intptr_t srcStride, pixel* dst, intptr_t dstStride, int width, int height, int shift)
{ planecopy_run(X265B200_PC_PP_SHR, src, srcStride, PX, dst, dstStride, width, height, shift, 0); }
template<int LOG2>
void ssim_dist_thunk(const pixel* fenc, uint32_t fStride, const pixel* recon, intptr_t rstride, uint64_t* ssBlock, int shift, uint64_t* ac_k)
{
    const int N = 1 << LOG2;
    void* dF = up_span(0, fenc, fStride, N, N, PX); void* dR = up_span(1, recon, rstride, N, N, PX);
    int64_t zero = 0; void* dOff = up1d(2, &zero, 8); uint64_t* dO = (uint64_t*)dev(3, 16);
    CK(x265b200_ssim_dist_dev(C(), X265_DEPTH, LOG2, dF, fStride, dR, rstride, (const int64_t*)dOff, (const int64_t*)dOff, 1, shift, dO, dO + 1));
    uint64_t r[2]; CK(x265b200_download(C(), r, dO, 16));
    *ssBlock = r[0]; *ac_k = r[1];
}
void norm_fact_thunk(const pixel* src, uint32_t blockSize, int shift, uint64_t* z_k)
{
    void* dS = up1d(0, src, (size_t)blockSize * blockSize * PX);
    int64_t zero = 0; void* dOff = up1d(1, &zero, 8); void* dO = dev(2, 8);
    CK(x265b200_norm_fact_dev(C(), X265_DEPTH, dS, (const int64_t*)dOff, 1, (int)blockSize, shift, (uint64_t*)dO));
    CK(x265b200_download(C(), z_k, dO, 8));
}

// ---- SSIM metric helpers and the HDR luma clip (pixel.cpp:631-702, :996-1016) ------------------------------------------------
void ssim_core_thunk(const pixel* pix1, intptr_t stride1, const pixel* pix2, intptr_t stride2, int sums[2][4])
{
    void* d1 = up_span(0, pix1, stride1, 8, 4, PX); void* d2 = up_span(1, pix2, stride2, 8, 4, PX);
    int64_t zero = 0; void* dOff = up1d(2, &zero, 8); void* dO = dev(3, 32);
    CK(x265b200_ssim_4x4x2_dev(C(), X265_DEPTH, d1, stride1, d2, stride2, (const int64_t*)dOff, (const int64_t*)dOff, 1, (int32_t*)dO));
    CK(x265b200_download(C(), sums, dO, 32));
}
float ssim_end4_thunk(int sum0[5][4], int sum1[5][4], int width)
{
    void* d0 = up1d(0, sum0, 80); void* d1 = up1d(1, sum1, 80);
    int32_t w = width; void* dW = up1d(2, &w, 4); void* dO = dev(3, 4);
    CK(x265b200_ssim_end4_dev(C(), X265_DEPTH, (const int32_t*)d0, (const int32_t*)d1, (const int32_t*)dW, 1, (float*)dO));
    float r; CK(x265b200_download(C(), &r, dO, 4));
    return r;
}
#if HIGH_BIT_DEPTH
pixel plane_clip_max_thunk(pixel* src, intptr_t stride, int width, int height, uint64_t* outsum, const pixel minPix, const pixel maxPix)
{
    *outsum = 0;
    if (width <= 0 || height <= 0) return 0;
    void* dS = up_span(0, src, stride, width, height, PX);
    char* dO = (char*)dev(1, 16);
    CK(x265b200_plane_clip_max_dev(C(), X265_DEPTH, dS, stride, width, height, minPix, maxPix, (uint64_t*)dO, (uint32_t*)(dO + 8)));
    CK(x265b200_download(C(), src, dS, ((size_t)(height - 1) * stride + width) * PX));
    uint64_t r[2]; CK(x265b200_download(C(), r, dO, 16));
    *outsum = r[0];
    return (pixel)(uint32_t)r[1];
}
#endif

} // namespace

namespace X265_NS {

void setupAssemblyPrimitives(EncoderPrimitives& p, int /*cpuMask: SIMD flags are meaningless for this backend*/)
{
    if (x265b200_device_count() <= 0)
    {
        fprintf(stderr, "x265 [error]: B200 primitive backend linked but no CUDA device is visible (no CPU fallback)\n");
        abort();
    }
    C();

#define PU(W, H) \
    p.pu[LUMA_ ## W ## x ## H].sad      = cmp_thunk<W, H, X265B200_CMP_SAD>; \
    p.pu[LUMA_ ## W ## x ## H].satd     = cmp_thunk<W, H, X265B200_CMP_SATD>; \
    p.pu[LUMA_ ## W ## x ## H].sad_x3   = sad_x3_thunk<W, H>; \
    p.pu[LUMA_ ## W ## x ## H].sad_x4   = sad_x4_thunk<W, H>; \
    p.pu[LUMA_ ## W ## x ## H].luma_hpp = ip_pp<8, W, H, X265B200_IP_HPP>; \
    p.pu[LUMA_ ## W ## x ## H].luma_vpp = ip_pp<8, W, H, X265B200_IP_VPP>; \
    p.pu[LUMA_ ## W ## x ## H].luma_hps = ip_hps<8, W, H>; \
    p.pu[LUMA_ ## W ## x ## H].luma_vps = ip_vps<8, W, H>; \
    p.pu[LUMA_ ## W ## x ## H].luma_vsp = ip_vsp<8, W, H>; \
    p.pu[LUMA_ ## W ## x ## H].luma_vss = ip_vss<8, W, H>; \
    p.pu[LUMA_ ## W ## x ## H].luma_hvpp = ip_hvpp<W, H>; \
    p.pu[LUMA_ ## W ## x ## H].convert_p2s[NONALIGNED] = ip_p2s<W, H>; \
    p.pu[LUMA_ ## W ## x ## H].convert_p2s[ALIGNED] = ip_p2s<W, H>; \
    p.pu[LUMA_ ## W ## x ## H].copy_pp = copy_pp_thunk<W, H>; \
    p.pu[LUMA_ ## W ## x ## H].pixelavg_pp[NONALIGNED] = pixelavg_pp_thunk<W, H>; \
    p.pu[LUMA_ ## W ## x ## H].pixelavg_pp[ALIGNED] = pixelavg_pp_thunk<W, H>; \
    p.pu[LUMA_ ## W ## x ## H].addAvg[NONALIGNED] = addAvg_thunk<W, H>; \
    p.pu[LUMA_ ## W ## x ## H].addAvg[ALIGNED] = addAvg_thunk<W, H>;
    PU(4, 4) PU(8, 8) PU(16, 16) PU(32, 32) PU(64, 64)
    PU(8, 4) PU(4, 8) PU(16, 8) PU(8, 16) PU(32, 16) PU(16, 32) PU(64, 32) PU(32, 64)
    PU(16, 12) PU(12, 16) PU(16, 4) PU(4, 16) PU(32, 24) PU(24, 32) PU(32, 8) PU(8, 32)
    PU(64, 48) PU(48, 64) PU(64, 16) PU(16, 64)
#undef PU

    // chroma tables (ipfilter.cpp:375-403, pixel.cpp:1168-1320): 4-tap filters + p2s for every chroma PU shape of 4:2:0 / 4:2:2 /
    // 4:4:4, addAvg / copy_pp for 4:2:0 / 4:2:2 (the 4:4:4 ones, and all chroma satd, are aliases of the luma pointers installed by
    // setupAliasPrimitives, primitives.cpp:139-208)
#define CH_FILT(CSP, PART, W, H) \
    p.chroma[CSP].pu[PART].filter_hpp = ip_pp<4, W, H, X265B200_IP_HPP>; \
    p.chroma[CSP].pu[PART].filter_vpp = ip_pp<4, W, H, X265B200_IP_VPP>; \
    p.chroma[CSP].pu[PART].filter_hps = ip_hps<4, W, H>; \
    p.chroma[CSP].pu[PART].filter_vps = ip_vps<4, W, H>; \
    p.chroma[CSP].pu[PART].filter_vsp = ip_vsp<4, W, H>; \
    p.chroma[CSP].pu[PART].filter_vss = ip_vss<4, W, H>; \
    p.chroma[CSP].pu[PART].p2s[NONALIGNED] = ip_p2s<W, H>; \
    p.chroma[CSP].pu[PART].p2s[ALIGNED] = ip_p2s<W, H>;
#define CH_GLUE(CSP, PART, W, H) \
    p.chroma[CSP].pu[PART].addAvg[NONALIGNED] = addAvg_thunk<W, H>; \
    p.chroma[CSP].pu[PART].addAvg[ALIGNED] = addAvg_thunk<W, H>; \
    p.chroma[CSP].pu[PART].copy_pp = copy_pp_thunk<W, H>;
#define CH420(W, H) CH_FILT(X265_CSP_I420, CHROMA_420_ ## W ## x ## H, W, H) CH_GLUE(X265_CSP_I420, CHROMA_420_ ## W ## x ## H, W, H)
#define CH422(W, H) CH_FILT(X265_CSP_I422, CHROMA_422_ ## W ## x ## H, W, H) CH_GLUE(X265_CSP_I422, CHROMA_422_ ## W ## x ## H, W, H)
#define CH444(W, H) CH_FILT(X265_CSP_I444, LUMA_ ## W ## x ## H, W, H)
    CH_GLUE(X265_CSP_I420, CHROMA_420_2x2, 2, 2)                 // no 2x2 filters in the table (ipfilter.cpp:419-421)
    CH420(4, 4) CH420(2, 4) CH420(4, 2) CH420(8, 8) CH420(8, 4) CH420(4, 8) CH420(8, 6) CH420(6, 8) CH420(8, 2) CH420(2, 8)
    CH420(16, 16) CH420(16, 8) CH420(8, 16) CH420(16, 12) CH420(12, 16) CH420(16, 4) CH420(4, 16)
    CH420(32, 32) CH420(32, 16) CH420(16, 32) CH420(32, 24) CH420(24, 32) CH420(32, 8) CH420(8, 32)
    CH422(4, 8) CH422(4, 4) CH422(2, 4) CH422(2, 8) CH422(8, 16) CH422(8, 8) CH422(4, 16) CH422(8, 12) CH422(6, 16) CH422(8, 4) CH422(2, 16)
    CH422(16, 32) CH422(16, 16) CH422(8, 32) CH422(16, 24) CH422(12, 32) CH422(16, 8) CH422(4, 32)
    CH422(32, 64) CH422(32, 32) CH422(16, 64) CH422(32, 48) CH422(24, 64) CH422(32, 16) CH422(8, 64)
    CH444(4, 4) CH444(8, 8) CH444(4, 8) CH444(8, 4) CH444(16, 16) CH444(16, 8) CH444(8, 16) CH444(16, 12) CH444(12, 16) CH444(16, 4) CH444(4, 16)
    CH444(32, 32) CH444(32, 16) CH444(16, 32) CH444(32, 24) CH444(24, 32) CH444(32, 8) CH444(8, 32)
    CH444(64, 64) CH444(64, 32) CH444(32, 64) CH444(64, 48) CH444(48, 64) CH444(64, 16) CH444(16, 64)
#undef CH420
#undef CH422
#undef CH444
#undef CH_FILT
#undef CH_GLUE

    // chroma CU blocks (pixel.cpp:1228-1246, :1307-1325): copies / residual / recon, sse_pp, and the sa8d compositions
#define CH_CU(CSP, IDX, W, H) \
    p.chroma[CSP].cu[IDX].copy_sp = copy_sp_thunk<W, H>; p.chroma[CSP].cu[IDX].copy_ps = copy_ps_thunk<W, H>; \
    p.chroma[CSP].cu[IDX].copy_ss = copy_ss_thunk<W, H>; p.chroma[CSP].cu[IDX].sub_ps = sub_ps_thunk<W, H>; \
    p.chroma[CSP].cu[IDX].add_ps[NONALIGNED] = add_ps_thunk<W, H>; p.chroma[CSP].cu[IDX].add_ps[ALIGNED] = add_ps_thunk<W, H>;
    CH_CU(X265_CSP_I420, BLOCK_420_2x2, 2, 2) CH_CU(X265_CSP_I420, BLOCK_420_4x4, 4, 4) CH_CU(X265_CSP_I420, BLOCK_420_8x8, 8, 8)
    CH_CU(X265_CSP_I420, BLOCK_420_16x16, 16, 16) CH_CU(X265_CSP_I420, BLOCK_420_32x32, 32, 32)
    CH_CU(X265_CSP_I422, BLOCK_422_2x4, 2, 4) CH_CU(X265_CSP_I422, BLOCK_422_4x8, 4, 8) CH_CU(X265_CSP_I422, BLOCK_422_8x16, 8, 16)
    CH_CU(X265_CSP_I422, BLOCK_422_16x32, 16, 32) CH_CU(X265_CSP_I422, BLOCK_422_32x64, 32, 64)
#undef CH_CU
    p.chroma[X265_CSP_I420].cu[BLOCK_420_4x4].sse_pp = sse_pp_thunk<4, 4>;     p.chroma[X265_CSP_I420].cu[BLOCK_420_8x8].sse_pp = sse_pp_thunk<8, 8>;
    p.chroma[X265_CSP_I420].cu[BLOCK_420_16x16].sse_pp = sse_pp_thunk<16, 16>; p.chroma[X265_CSP_I420].cu[BLOCK_420_32x32].sse_pp = sse_pp_thunk<32, 32>;
    p.chroma[X265_CSP_I422].cu[BLOCK_422_4x8].sse_pp = sse_pp_thunk<4, 8>;     p.chroma[X265_CSP_I422].cu[BLOCK_422_8x16].sse_pp = sse_pp_thunk<8, 16>;
    p.chroma[X265_CSP_I422].cu[BLOCK_422_16x32].sse_pp = sse_pp_thunk<16, 32>; p.chroma[X265_CSP_I422].cu[BLOCK_422_32x64].sse_pp = sse_pp_thunk<32, 64>;
    p.chroma[X265_CSP_I420].cu[BLOCK_8x8].sa8d   = cmp_thunk<4, 4, X265B200_CMP_SATD>;       // = chroma pu 4x4 satd (pixel.cpp:1243)
    p.chroma[X265_CSP_I420].cu[BLOCK_16x16].sa8d = cmp_thunk<8, 8, X265B200_CMP_SA8D8>;      // sa8d8<8, 8>
    p.chroma[X265_CSP_I420].cu[BLOCK_32x32].sa8d = cmp_thunk<16, 16, X265B200_CMP_SA8D>;     // sa8d16<16, 16>
    p.chroma[X265_CSP_I420].cu[BLOCK_64x64].sa8d = cmp_thunk<32, 32, X265B200_CMP_SA8D>;     // sa8d16<32, 32>
    p.chroma[X265_CSP_I422].cu[BLOCK_8x8].sa8d   = cmp_thunk<4, 8, X265B200_CMP_SATD>;       // satd4<4, 8> (pixel.cpp:1322)
    p.chroma[X265_CSP_I422].cu[BLOCK_16x16].sa8d = cmp_thunk<8, 16, X265B200_CMP_SA8D8>;     // sa8d8<8, 16>
    p.chroma[X265_CSP_I422].cu[BLOCK_32x32].sa8d = cmp_thunk<16, 32, X265B200_CMP_SA8D>;     // sa8d16<16, 32>
    p.chroma[X265_CSP_I422].cu[BLOCK_64x64].sa8d = cmp_thunk<32, 64, X265B200_CMP_SA8D>;     // sa8d16<32, 64>

#define CU(IDX, N, LOG2) \
    p.cu[IDX].sa8d   = cmp_thunk<N, N, X265B200_CMP_SA8D>; \
    p.cu[IDX].sse_pp = sse_pp_thunk<N, N>; \
    p.cu[IDX].sse_ss = sse_ss_thunk<N, N>; \
    p.cu[IDX].ssd_s[NONALIGNED] = ssd_s_thunk<N>; \
    p.cu[IDX].ssd_s[ALIGNED] = ssd_s_thunk<N>; \
    p.cu[IDX].copy_pp = copy_pp_thunk<N, N>; p.cu[IDX].copy_sp = copy_sp_thunk<N, N>; \
    p.cu[IDX].copy_ps = copy_ps_thunk<N, N>; p.cu[IDX].copy_ss = copy_ss_thunk<N, N>; \
    p.cu[IDX].sub_ps = sub_ps_thunk<N, N>; \
    p.cu[IDX].add_ps[NONALIGNED] = add_ps_thunk<N, N>; p.cu[IDX].add_ps[ALIGNED] = add_ps_thunk<N, N>; \
    p.cu[IDX].blockfill_s[NONALIGNED] = blockfill_s_thunk<N>; p.cu[IDX].blockfill_s[ALIGNED] = blockfill_s_thunk<N>; \
    p.cu[IDX].var = var_thunk<N>; \
    p.cu[IDX].psy_cost_pp = psy_cost_pp_thunk<N>;
    CU(BLOCK_4x4, 4, 2) CU(BLOCK_8x8, 8, 3) CU(BLOCK_16x16, 16, 4) CU(BLOCK_32x32, 32, 5) CU(BLOCK_64x64, 64, 6)
#undef CU

#define TU(IDX, N, LOG2) \
    p.cu[IDX].dct = dct_thunk<IDX, N>; p.cu[IDX].standard_dct = dct_thunk<IDX, N>; \
    p.cu[IDX].idct = idct_thunk<IDX, N>; \
    p.cu[IDX].count_nonzero = count_nonzero_thunk<N>; \
    p.cu[IDX].copy_cnt = copy_cnt_thunk<N>; \
    p.cu[IDX].calcresidual[NONALIGNED] = calcresidual_thunk<N>; p.cu[IDX].calcresidual[ALIGNED] = calcresidual_thunk<N>; \
    p.cu[IDX].transpose = transpose_thunk<N>; \
    p.cu[IDX].cpy2Dto1D_shl = cpy2Dto1D_shl_thunk<N>; p.cu[IDX].cpy2Dto1D_shr = cpy2Dto1D_shr_thunk<N>; \
    p.cu[IDX].cpy1Dto2D_shl[NONALIGNED] = cpy1Dto2D_shl_thunk<N>; p.cu[IDX].cpy1Dto2D_shl[ALIGNED] = cpy1Dto2D_shl_thunk<N>; \
    p.cu[IDX].cpy1Dto2D_shr = cpy1Dto2D_shr_thunk<N>; \
    p.cu[IDX].intra_filter = intra_filter_thunk<LOG2>; \
    p.cu[IDX].intra_pred_allangs = intra_allangs_thunk<LOG2>; \
    p.cu[IDX].intra_pred[PLANAR_IDX] = intra_pred_thunk<LOG2, 0>; p.cu[IDX].intra_pred[DC_IDX] = intra_pred_thunk<LOG2, 1>; \
    for (int m = 2; m < NUM_INTRA_MODE; m++) p.cu[IDX].intra_pred[m] = intra_pred_thunk<LOG2, 2>;
    TU(BLOCK_4x4, 4, 2) TU(BLOCK_8x8, 8, 3) TU(BLOCK_16x16, 16, 4) TU(BLOCK_32x32, 32, 5)
#undef TU
    p.dst4x4 = dct_thunk<4, 4>;
    p.idst4x4 = idct_thunk<4, 4>;
    p.quant = quant_thunk;
    p.nquant = nquant_thunk;
    p.dequant_normal = dequant_normal_thunk;
    p.dequant_scaling = dequant_scaling_thunk;
    p.cu[BLOCK_64x64].transpose = transpose_thunk<64>;
    p.weight_pp = weight_pp_thunk;
    p.weight_sp = weight_sp_thunk;
    p.denoiseDct = denoise_dct_thunk;
    // installed unconditionally; x265_setup_primitives only routes cu[].dct through it with --lowpass-dct (primitives.cpp:75-86)
    p.cu[BLOCK_8x8].lowpass_dct = lowpass_dct_thunk<1, 8>;
    p.cu[BLOCK_16x16].lowpass_dct = lowpass_dct_thunk<2, 16>;
    p.cu[BLOCK_32x32].lowpass_dct = lowpass_dct_thunk<3, 32>;

    // --me sea: pu[].ads per pixel.cpp:1105-1129 and the integral row primitives (framefilter.cpp:142-156)
#define ADS(W, H, K) p.pu[LUMA_ ## W ## x ## H].ads = ads_thunk<K, (W >> 1)>;
    ADS(4, 4, 1) ADS(8, 8, 1) ADS(8, 4, 2) ADS(4, 8, 2) ADS(16, 16, 4) ADS(16, 8, 2) ADS(8, 16, 2) ADS(16, 12, 1) ADS(12, 16, 1)
    ADS(16, 4, 1) ADS(4, 16, 1) ADS(32, 32, 4) ADS(32, 16, 2) ADS(16, 32, 2) ADS(32, 24, 4) ADS(24, 32, 4) ADS(32, 8, 4) ADS(8, 32, 4)
    ADS(64, 64, 4) ADS(64, 32, 2) ADS(32, 64, 2) ADS(64, 48, 4) ADS(48, 64, 4) ADS(64, 16, 4) ADS(16, 64, 4)
#undef ADS
    p.integral_inith[INTEGRAL_4] = integral_inith_thunk<4>;   p.integral_initv[INTEGRAL_4] = integral_initv_thunk<4>;
    p.integral_inith[INTEGRAL_8] = integral_inith_thunk<8>;   p.integral_initv[INTEGRAL_8] = integral_initv_thunk<8>;
    p.integral_inith[INTEGRAL_12] = integral_inith_thunk<12>; p.integral_initv[INTEGRAL_12] = integral_initv_thunk<12>;
    p.integral_inith[INTEGRAL_16] = integral_inith_thunk<16>; p.integral_initv[INTEGRAL_16] = integral_initv_thunk<16>;
    p.integral_inith[INTEGRAL_24] = integral_inith_thunk<24>; p.integral_initv[INTEGRAL_24] = integral_initv_thunk<24>;
    p.integral_inith[INTEGRAL_32] = integral_inith_thunk<32>; p.integral_initv[INTEGRAL_32] = integral_initv_thunk<32>;

    p.frameInitLowres = frame_init_lowres_thunk;
    p.frameInitLowerRes = frame_init_lowres_thunk;
    p.extendRowBorder = extend_row_border_thunk;
    p.scale1D_128to64[NONALIGNED] = scale1D_thunk; p.scale1D_128to64[ALIGNED] = scale1D_thunk;
    p.scale2D_64to32 = scale2D_thunk;

    // in-loop filters (setupLoopFilterPrimitives_c loopfilter.cpp:184-201, setupSaoPrimitives_c sao.cpp:1927-1935)
    p.saoCuOrgE0 = saoE0_thunk; p.saoCuOrgE1 = saoE1_thunk; p.saoCuOrgE1_2Rows = saoE1_2rows_thunk;
    p.saoCuOrgE2[0] = saoE2_thunk; p.saoCuOrgE2[1] = saoE2_thunk;
    p.saoCuOrgE3[0] = saoE3_thunk; p.saoCuOrgE3[1] = saoE3_thunk;
    p.saoCuOrgB0 = saoB0_thunk;
    p.saoCuStatsBO = saoStatsBO_thunk; p.saoCuStatsE0 = saoStatsE0_thunk; p.saoCuStatsE1 = saoStatsE1_thunk;
    p.saoCuStatsE2 = saoStatsE2_thunk; p.saoCuStatsE3 = saoStatsE3_thunk;
    p.sign = sign_thunk;
    p.pelFilterLumaStrong[0] = pel_filter_luma_thunk; p.pelFilterLumaStrong[1] = pel_filter_luma_thunk;
    p.pelFilterChroma[0] = pel_filter_chroma_thunk;   p.pelFilterChroma[1] = pel_filter_chroma_thunk;

    // cuTree propagation / fixed-point packing, picture ingest copies, --ssim-rd sums (pixel.cpp:1337-1357)
    p.propagateCost = propagate_cost_thunk; p.fix8Pack = fix8_pack_thunk; p.fix8Unpack = fix8_unpack_thunk;
    p.planecopy_cp = planecopy_cp_thunk; p.planecopy_sp = planecopy_sp_thunk; p.planecopy_sp_shl = planecopy_sp_shl_thunk;
    p.planecopy_pp_shr = planecopy_pp_shr_thunk;
    p.cu[BLOCK_4x4].ssimDist = ssim_dist_thunk<2>; p.cu[BLOCK_8x8].ssimDist = ssim_dist_thunk<3>; p.cu[BLOCK_16x16].ssimDist = ssim_dist_thunk<4>;
    p.cu[BLOCK_32x32].ssimDist = ssim_dist_thunk<5>; p.cu[BLOCK_64x64].ssimDist = ssim_dist_thunk<6>;
    p.cu[BLOCK_8x8].normFact = norm_fact_thunk; p.cu[BLOCK_16x16].normFact = norm_fact_thunk;
    p.cu[BLOCK_32x32].normFact = norm_fact_thunk; p.cu[BLOCK_64x64].normFact = norm_fact_thunk;
    p.ssim_4x4x2_core = ssim_core_thunk; p.ssim_end_4 = ssim_end4_thunk;
#if HIGH_BIT_DEPTH
    p.planeClipAndMax = plane_clip_max_thunk;
#endif
}

} // namespace X265_NS

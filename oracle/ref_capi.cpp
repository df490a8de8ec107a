/* oracle/ref_capi.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * A thin extern "C" accessor over the UNMODIFIED reference (x265 3.5+1) compiled from
 * /root/reference by oracle/Makefile.  It lets tests/ and bench.py's cpu_baseline /
 * `--impl reference` arm call the reference's own C primitive table
 * (source/common/primitives.h:237 `struct EncoderPrimitives`, filled by
 * x265_setup_primitives(), source/common/primitives.cpp:248) and the reference's own
 * MotionEstimate class (source/encoder/motion.h:37, motion.cpp:739) through ctypes.
 *
 * This file is OUR code; it includes the reference headers where they lie and contains no
 * reference source.  Built twice: X265_NS=x265 (8-bit) and X265_NS=x265_10bit.
 */
#include "common.h"
#include "primitives.h"
#include "constants.h"
#include "motion.h"
#include "bitcost.h"
#include "lowres.h"
#include "x265.h"

#include <thread>
#include <atomic>
#include <vector>
#include <cstring>

using namespace X265_NS;

namespace {
bool g_init = false;
void ensure_init()
{
    if (g_init) return;
    static x265_param param;
    x265_param_default(&param);
    param.logLevel = X265_LOG_NONE;
    x265_setup_primitives(&param);   /* primitives.cpp:248: C table + aliases (ENABLE_ASSEMBLY=0) */
    MotionEstimate::initScales();    /* motion.cpp:123 */
    g_init = true;
}

/* run fn(i, t) for i in [0,n) on `threads` std::threads.  Work is handed out dynamically in small chunks from an atomic
 * counter (jobs differ in cost by two orders of magnitude -- a 64x64 search against an 8x8 one -- so static contiguous ranges
 * would leave most threads idle).  t is the worker index, stable for the lifetime of the call. */
template<class F> void parallel_for(int64_t n, int threads, F fn)
{
    if (threads <= 1 || n < 2) { for (int64_t i = 0; i < n; i++) fn(i, 0); return; }
    if ((int64_t)threads > n) threads = (int)n;
    std::atomic<int64_t> next(0);
    int64_t chunk = n / ((int64_t)threads * 64);
    if (chunk < 1) chunk = 1;
    if (chunk > 64) chunk = 64;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&, t]() {
            for (;;)
            {
                int64_t lo = next.fetch_add(chunk);
                if (lo >= n) break;
                int64_t hi = lo + chunk < n ? lo + chunk : n;
                for (int64_t i = lo; i < hi; i++) fn(i, t);
            }
        });
    for (auto& th : pool) th.join();
}
}

extern "C" {

/* Installs the reference's SSE-intrinsic entries (vec/vec-primitives.cpp:62 setupInstrinsicPrimitives: cu[8/16/32].idct,
 * cu[16/32].dct, dequant_scaling) over the C table -- the only optimised table that builds without nasm.  Opt-in: the parity
 * tests keep comparing against the plain C table; bench.py's CPU legs call this so the baseline is the reference's best
 * buildable path.  Returns the number of slots that changed. */
int ref_enable_intrinsics(void)
{
    ensure_init();
    EncoderPrimitives before = primitives;
    setupInstrinsicPrimitives(primitives, X265_CPU_SSE2 | X265_CPU_SSE3 | X265_CPU_SSSE3 | X265_CPU_SSE4);
    int changed = 0;
    for (int i = 0; i < NUM_CU_SIZES; i++)
        changed += (before.cu[i].dct != primitives.cu[i].dct) + (before.cu[i].idct != primitives.cu[i].idct);
    changed += before.dequant_scaling != primitives.dequant_scaling;
    for (int i = 0; i < NUM_CU_SIZES; i++) primitives.cu[i].standard_dct = primitives.cu[i].dct;
    return changed;
}

int ref_depth(void) { return X265_DEPTH; }
int ref_pixel_bytes(void) { return (int)sizeof(pixel); }
int ref_sse_bytes(void) { return (int)sizeof(sse_t); }
void ref_init(void) { ensure_init(); }

/* partitionFromSizes() LUT, primitives.h:435 */
int ref_partition_from_sizes(int w, int h) { ensure_init(); return partitionFromSizes(w, h); }

/* ---- pixel compare family (primitives.h:133 pixelcmp_t) --------------------------------
 * fam: 0 pu[idx].sad  1 pu[idx].satd  2 cu[idx].sa8d  3 cu[idx].psy_cost_pp
 *      4 chroma[csp=1].pu[idx].satd   5 chroma[1].cu[idx].sa8d   6 chroma[2].cu[idx].sa8d */
static pixelcmp_t get_cmp(int fam, int idx)
{
    switch (fam)
    {
    case 0: return primitives.pu[idx].sad;
    case 1: return primitives.pu[idx].satd;
    case 2: return primitives.cu[idx].sa8d;
    case 3: return primitives.cu[idx].psy_cost_pp;
    case 4: return primitives.chroma[X265_CSP_I420].pu[idx].satd;
    case 5: return primitives.chroma[X265_CSP_I420].cu[idx].sa8d;
    case 6: return primitives.chroma[X265_CSP_I422].cu[idx].sa8d;
    }
    return NULL;
}

int ref_pixelcmp(int fam, int idx, const void* a, intptr_t sa, const void* b, intptr_t sb)
{
    ensure_init();
    pixelcmp_t f = get_cmp(fam, idx);
    return f ? f((const pixel*)a, sa, (const pixel*)b, sb) : -1;
}

/* n independent block pairs: a = planeA + offA[i], b = planeB + offB[i] (element offsets) */
int ref_pixelcmp_batch(int fam, int idx, const void* planeA, intptr_t sa, const void* planeB, intptr_t sb,
                       const int64_t* offA, const int64_t* offB, int64_t n, int32_t* out, int threads)
{
    ensure_init();
    pixelcmp_t f = get_cmp(fam, idx);
    if (!f) return -1;
    const pixel* A = (const pixel*)planeA; const pixel* B = (const pixel*)planeB;
    parallel_for(n, threads, [=](int64_t i, int) { out[i] = f(A + offA[i], sa, B + offB[i], sb); });
    return 0;
}

/* fam: 0 cu[idx].sse_pp (pixel,pixel)  1 cu[idx].sse_ss (int16,int16)  2 cu[idx].ssd_s (int16; b ignored) */
uint64_t ref_sse(int fam, int idx, const void* a, intptr_t sa, const void* b, intptr_t sb)
{
    ensure_init();
    switch (fam)
    {
    case 0: return primitives.cu[idx].sse_pp((const pixel*)a, sa, (const pixel*)b, sb);
    case 1: return primitives.cu[idx].sse_ss((const int16_t*)a, sa, (const int16_t*)b, sb);
    case 2: return primitives.cu[idx].ssd_s[NONALIGNED]((const int16_t*)a, sa);
    }
    return 0;
}

int ref_sse_batch(int fam, int idx, const void* planeA, intptr_t sa, const void* planeB, intptr_t sb,
                  const int64_t* offA, const int64_t* offB, int64_t n, uint64_t* out, int threads)
{
    ensure_init();
    parallel_for(n, threads, [=](int64_t i, int) {
        if (fam == 0) out[i] = primitives.cu[idx].sse_pp((const pixel*)planeA + offA[i], sa, (const pixel*)planeB + offB[i], sb);
        else if (fam == 1) out[i] = primitives.cu[idx].sse_ss((const int16_t*)planeA + offA[i], sa, (const int16_t*)planeB + offB[i], sb);
        else out[i] = primitives.cu[idx].ssd_s[NONALIGNED]((const int16_t*)planeA + offA[i], sa);
    });
    return 0;
}

/* sad_x3 / sad_x4 (primitives.h:139-140): fenc has the hard-wired FENC_STRIDE (pixel.cpp:89,113) */
void ref_sad_x3(int part, const void* fenc, const void* r0, const void* r1, const void* r2, intptr_t stride, int32_t* res)
{
    ensure_init();
    primitives.pu[part].sad_x3((const pixel*)fenc, (const pixel*)r0, (const pixel*)r1, (const pixel*)r2, stride, res);
}
void ref_sad_x4(int part, const void* fenc, const void* r0, const void* r1, const void* r2, const void* r3, intptr_t stride, int32_t* res)
{
    ensure_init();
    primitives.pu[part].sad_x4((const pixel*)fenc, (const pixel*)r0, (const pixel*)r1, (const pixel*)r2, (const pixel*)r3, stride, res);
}

/* ads (primitives.h:138) */
int ref_ads(int part, int* encDC, uint32_t* sums, int delta, uint16_t* costMvX, int16_t* mvs, int width, int thresh)
{
    ensure_init();
    return primitives.pu[part].ads(encDC, sums, delta, costMvX, mvs, width, thresh);
}

/* ---- transforms (primitives.h:153-162) -------------------------------------------------
 * idx 0..3 = cu[BLOCK_4x4..32x32].dct / idct, idx 4 = dst4x4 / idst4x4 */
void ref_dct(int idx, const int16_t* src, int16_t* dst, intptr_t srcStride)
{
    ensure_init();
    (idx == 4 ? primitives.dst4x4 : primitives.cu[idx].dct)(src, dst, srcStride);
}
void ref_idct(int idx, const int16_t* src, int16_t* dst, intptr_t dstStride)
{
    ensure_init();
    (idx == 4 ? primitives.idst4x4 : primitives.cu[idx].idct)(src, dst, dstStride);
}
/* n blocks, src block i at src + i*srcBlockStride (elements), contiguous dst blocks */
void ref_dct_batch(int idx, const int16_t* src, int64_t srcBlockStride, intptr_t srcStride, int16_t* dst, int64_t n, int threads)
{
    ensure_init();
    int sz = idx == 4 ? 4 : (4 << idx);
    dct_t f = idx == 4 ? primitives.dst4x4 : primitives.cu[idx].dct;
    parallel_for(n, threads, [=](int64_t i, int) { f(src + i * srcBlockStride, dst + i * sz * sz, srcStride); });
}
void ref_idct_batch(int idx, const int16_t* src, int16_t* dst, int64_t dstBlockStride, intptr_t dstStride, int64_t n, int threads)
{
    ensure_init();
    int sz = idx == 4 ? 4 : (4 << idx);
    idct_t f = idx == 4 ? primitives.idst4x4 : primitives.cu[idx].idct;
    parallel_for(n, threads, [=](int64_t i, int) { f(src + i * sz * sz, dst + i * dstBlockStride, dstStride); });
}
uint32_t ref_quant(const int16_t* coef, const int32_t* quantCoeff, int32_t* deltaU, int16_t* qCoef, int qBits, int add, int numCoeff)
{
    ensure_init();
    return primitives.quant(coef, quantCoeff, deltaU, qCoef, qBits, add, numCoeff);
}
uint32_t ref_nquant(const int16_t* coef, const int32_t* quantCoeff, int16_t* qCoef, int qBits, int add, int numCoeff)
{
    ensure_init();
    return primitives.nquant(coef, quantCoeff, qCoef, qBits, add, numCoeff);
}
void ref_dequant_normal(const int16_t* quantCoef, int16_t* coef, int num, int scale, int shift)
{
    ensure_init();
    primitives.dequant_normal(quantCoef, coef, num, scale, shift);
}
void ref_dequant_scaling(const int16_t* src, const int32_t* dequantCoef, int16_t* dst, int num, int mcqp_miper, int shift)
{
    ensure_init();
    primitives.dequant_scaling(src, dequantCoef, dst, num, mcqp_miper, shift);
}
int ref_count_nonzero(int idx, const int16_t* q) { ensure_init(); return primitives.cu[idx].count_nonzero(q); }
uint32_t ref_copy_cnt(int idx, int16_t* coeff, const int16_t* residual, intptr_t stride) { ensure_init(); return primitives.cu[idx].copy_cnt(coeff, residual, stride); }
void ref_denoise_dct(int16_t* dctCoef, uint32_t* resSum, const uint16_t* offset, int numCoeff) { ensure_init(); primitives.denoiseDct(dctCoef, resSum, offset, numCoeff); }

/* constant tables (constants.cpp:250-344) so tests can pin the product's generated tables */
const int16_t* ref_dct_table(int idx)
{
    switch (idx) { case 0: return &g_t4[0][0]; case 1: return &g_t8[0][0]; case 2: return &g_t16[0][0]; case 3: return &g_t32[0][0]; }
    return NULL;
}
const int16_t* ref_luma_filter(void) { return &g_lumaFilter[0][0]; }
const int16_t* ref_chroma_filter(void) { return &g_chromaFilter[0][0]; }
double ref_lambda(int qp) { return x265_lambda_tab[qp]; }
double ref_lambda2(int qp) { return x265_lambda2_tab[qp]; }

/* ---- interpolation (primitives.h:176-182) ----------------------------------------------
 * csp < 0: luma pu[part].luma_*; csp >= 0: chroma[csp].pu[part].filter_*
 * kind: 0 hpp 1 hps 2 vpp 3 vps 4 vsp 5 vss 6 hvpp(luma only) 7 p2s */
void ref_interp(int kind, int csp, int part, const void* src, intptr_t srcStride, void* dst, intptr_t dstStride,
                int coeffIdx, int arg2 /* isRowExt for hps, idxY for hvpp */)
{
    ensure_init();
    if (csp < 0)
    {
        EncoderPrimitives::PU& p = primitives.pu[part];
        switch (kind)
        {
        case 0: p.luma_hpp((const pixel*)src, srcStride, (pixel*)dst, dstStride, coeffIdx); break;
        case 1: p.luma_hps((const pixel*)src, srcStride, (int16_t*)dst, dstStride, coeffIdx, arg2); break;
        case 2: p.luma_vpp((const pixel*)src, srcStride, (pixel*)dst, dstStride, coeffIdx); break;
        case 3: p.luma_vps((const pixel*)src, srcStride, (int16_t*)dst, dstStride, coeffIdx); break;
        case 4: p.luma_vsp((const int16_t*)src, srcStride, (pixel*)dst, dstStride, coeffIdx); break;
        case 5: p.luma_vss((const int16_t*)src, srcStride, (int16_t*)dst, dstStride, coeffIdx); break;
        case 6: p.luma_hvpp((const pixel*)src, srcStride, (pixel*)dst, dstStride, coeffIdx, arg2); break;
        case 7: p.convert_p2s[NONALIGNED]((const pixel*)src, srcStride, (int16_t*)dst, dstStride); break;
        }
    }
    else
    {
        EncoderPrimitives::Chroma::PUChroma& p = primitives.chroma[csp].pu[part];
        switch (kind)
        {
        case 0: p.filter_hpp((const pixel*)src, srcStride, (pixel*)dst, dstStride, coeffIdx); break;
        case 1: p.filter_hps((const pixel*)src, srcStride, (int16_t*)dst, dstStride, coeffIdx, arg2); break;
        case 2: p.filter_vpp((const pixel*)src, srcStride, (pixel*)dst, dstStride, coeffIdx); break;
        case 3: p.filter_vps((const pixel*)src, srcStride, (int16_t*)dst, dstStride, coeffIdx); break;
        case 4: p.filter_vsp((const int16_t*)src, srcStride, (pixel*)dst, dstStride, coeffIdx); break;
        case 5: p.filter_vss((const int16_t*)src, srcStride, (int16_t*)dst, dstStride, coeffIdx); break;
        case 7: p.p2s[NONALIGNED]((const pixel*)src, srcStride, (int16_t*)dst, dstStride); break;
        }
    }
}
int ref_interp_available(int kind, int csp, int part)
{
    ensure_init();
    if (csp < 0) return 1;
    EncoderPrimitives::Chroma::PUChroma& p = primitives.chroma[csp].pu[part];
    switch (kind)
    {
    case 0: return p.filter_hpp != NULL; case 1: return p.filter_hps != NULL; case 2: return p.filter_vpp != NULL;
    case 3: return p.filter_vps != NULL; case 4: return p.filter_vsp != NULL; case 5: return p.filter_vss != NULL;
    case 7: return p.p2s[NONALIGNED] != NULL;
    }
    return 0;
}

/* ---- intra (primitives.h:143-145) ------------------------------------------------------ */
void ref_intra_pred(int sizeIdx, int mode, void* dst, intptr_t dstStride, const void* srcPix, int bFilter)
{
    ensure_init();
    primitives.cu[sizeIdx].intra_pred[mode]((pixel*)dst, dstStride, (const pixel*)srcPix, mode, bFilter);
}
void ref_intra_filter(int sizeIdx, const void* ref, void* filtered)
{
    ensure_init();
    primitives.cu[sizeIdx].intra_filter((const pixel*)ref, (pixel*)filtered);
}

/* ---- glue ------------------------------------------------------------------------------- */
void ref_pixelavg_pp(int part, void* dst, intptr_t ds, const void* s0, intptr_t ss0, const void* s1, intptr_t ss1)
{
    ensure_init();
    primitives.pu[part].pixelavg_pp[NONALIGNED]((pixel*)dst, ds, (const pixel*)s0, ss0, (const pixel*)s1, ss1, 32);
}
void ref_sub_ps(int idx, int16_t* dst, intptr_t ds, const void* s0, const void* s1, intptr_t ss0, intptr_t ss1)
{
    ensure_init();
    primitives.cu[idx].sub_ps(dst, ds, (const pixel*)s0, (const pixel*)s1, ss0, ss1);
}
void ref_add_ps(int idx, void* dst, intptr_t ds, const void* s0, const int16_t* s1, intptr_t ss0, intptr_t ss1)
{
    ensure_init();
    primitives.cu[idx].add_ps[NONALIGNED]((pixel*)dst, ds, (const pixel*)s0, s1, ss0, ss1);
}
void ref_frame_init_lowres(const void* src0, void* dst0, void* dsth, void* dstv, void* dstc,
                           intptr_t srcStride, intptr_t dstStride, int width, int height)
{
    ensure_init();
    primitives.frameInitLowres((const pixel*)src0, (pixel*)dst0, (pixel*)dsth, (pixel*)dstv, (pixel*)dstc, srcStride, dstStride, width, height);
}
/* glue slots used by tests/test_glue_gpu.py (the rest is pinned by the reference's own TestBench) */
void ref_lowpass_dct(int idx, const int16_t* src, int16_t* dst, intptr_t srcStride)
{
    ensure_init();
    /* lowPassDct*_c call through cu[].standard_dct, which only enableLowpassDCTPrimitives (primitives.cpp:75-86,
     * --lowpass-dct) fills in; do that half of it here without rerouting cu[].dct */
    for (int i = BLOCK_4x4; i <= BLOCK_32x32; i++)
        if (!primitives.cu[i].standard_dct) primitives.cu[i].standard_dct = primitives.cu[i].dct;
    primitives.cu[idx].lowpass_dct(src, dst, srcStride);
}
void ref_weight_pp(const void* src, void* dst, intptr_t stride, int w, int h, int w0, int round, int shift, int offset)
{ ensure_init(); primitives.weight_pp((const pixel*)src, (pixel*)dst, stride, w, h, w0, round, shift, offset); }
void ref_weight_sp(const int16_t* src, void* dst, intptr_t ss, intptr_t ds, int w, int h, int w0, int round, int shift, int offset)
{ ensure_init(); primitives.weight_sp(src, (pixel*)dst, ss, ds, w, h, w0, round, shift, offset); }
int ref_psy_cost(int idx, const void* s, intptr_t ss, const void* r, intptr_t rs) { ensure_init(); return primitives.cu[idx].psy_cost_pp((const pixel*)s, ss, (const pixel*)r, rs); }
void ref_addavg(int part, const int16_t* a, const int16_t* b, void* dst, intptr_t sa, intptr_t sb, intptr_t ds)
{ ensure_init(); primitives.pu[part].addAvg[NONALIGNED](a, b, (pixel*)dst, sa, sb, ds); }
uint64_t ref_var(int idx, const void* pix, intptr_t stride) { ensure_init(); return primitives.cu[idx].var((const pixel*)pix, stride); }

/* ---- BitCost table (bitcost.cpp:31-110): cost[i] for i in [-2*BC_MAX_MV, 2*BC_MAX_MV] --- */
struct BitCostPeek : public BitCost { const uint16_t* table() const { return m_cost; } };
void ref_bitcost_table(int qp, uint16_t* out /* 4*32768+1 entries, out[2*32768+i] = cost[i] */)
{
    ensure_init();
    BitCostPeek bc; bc.setQP(qp);
    memcpy(out, bc.table() - 2 * 32768, sizeof(uint16_t) * (4 * 32768 + 1));
}

/* ---- MotionEstimate::motionEstimate (motion.cpp:739) ------------------------------------
 * One descriptor per PU search, luma only (the lookahead-style setSourcePU, motion.cpp:167,
 * which leaves ctuAddr = -1 so no PicYuv is needed).  All MVs as {x,y} int32 pairs. */
struct RefMEJob
{
    int32_t  puX, puY;        /* top-left of the PU in the fenc / ref planes (pixels)        */
    int32_t  w, h;            /* PU size                                                      */
    int32_t  mvminX, mvminY, mvmaxX, mvmaxY;   /* full-pel, inclusive (motion.cpp:740-741)   */
    int32_t  mvpX, mvpY;      /* qpel predictor                                               */
    int32_t  numCand;         /* <= 8                                                         */
    int32_t  mvc[8][2];       /* qpel candidates                                              */
    int32_t  outMvX, outMvY;  /* OUT: qpel                                                    */
    int32_t  outCost;         /* OUT                                                          */
};

int ref_me_batch(const void* fencPlane, intptr_t fencStride, const void* refPlane, intptr_t refStride,
                 RefMEJob* jobs, int64_t n, int searchMethod, int subpelRefine, int merange, int qp,
                 int maxSlices, int threads)
{
    ensure_init();
    { BitCost warm; warm.setQP(qp); }            /* build the shared cost table before threading */
    int nt = threads < 1 ? 1 : threads;
    std::vector<MotionEstimate*> mes(nt);
    for (int t = 0; t < nt; t++)
    {
        mes[t] = new MotionEstimate;
        mes[t]->init(X265_CSP_I400);             /* motion.cpp:118, luma-only fencPUYuv */
        mes[t]->setQP(qp);
    }
    parallel_for(n, nt, [&](int64_t i, int t) {
        RefMEJob& j = jobs[i];
        MotionEstimate& me = *mes[t];
        ReferencePlanes ref;
        ref.fpelPlane[0] = (pixel*)refPlane;
        ref.lumaStride = refStride;
        ref.isLowres = false;
        intptr_t off = j.puX + (intptr_t)j.puY * refStride;
        /* fenc and ref planes must share one stride for this entry point (blockOffset is reused) */
        const int sm = searchMethod == 6 ? X265_HEX_SEARCH : searchMethod;
        me.setSourcePU((pixel*)fencPlane, fencStride, off, j.w, j.h, sm, sm, sm, subpelRefine);
        MV mvmin(j.mvminX, j.mvminY), mvmax(j.mvmaxX, j.mvmaxY), mvp(j.mvpX, j.mvpY), out(0, 0);
        MV mvc[8];
        for (int k = 0; k < j.numCand; k++) mvc[k] = MV(j.mvc[k][0], j.mvc[k][1]);
        if (searchMethod == 6)      /* MotionEstimate::refineMV (motion.cpp:606-737): returns only the MV */
        {
            me.refineMV(&ref, mvmin, mvmax, mvp, out);
            j.outCost = 0;
        }
        else
            j.outCost = me.motionEstimate(&ref, mvmin, mvmax, mvp, j.numCand, mvc, merange, out, (uint32_t)maxSlices);
        j.outMvX = out.x; j.outMvY = out.y;
    });
    for (int t = 0; t < nt; t++) delete mes[t];
    return 0;
}

} /* extern "C" */

/* ---- lookahead: Lowres::init, LookaheadTLD::lowresIntraEstimate, CostEstimateGroup::singleCost ----
 * Built on the reference's own Lookahead / Lowres / PicYuv objects (no pool => non-cooperative path,
 * slicetype.cpp:3175-3199).  weightp, HME and cutree are off; AQ factors can be injected. */
#define protected public          /* test infrastructure: reach Lookahead::estimateCUPropagate (the reference source is not modified) */
#include "slicetype.h"
#undef protected
#include "picyuv.h"
#include "frame.h"

namespace {
struct RefLA
{
    x265_param* param;
    Lookahead*  la;
    std::vector<PicYuv*> pics;
    std::vector<Lowres*> lowres;
    std::vector<Lowres*> framePtrs;
};
}

extern "C" {

static void* la_create(int width, int height, int bframes, int aq, int hme);
void* ref_la_create(int width, int height, int bframes, int aq) { return la_create(width, height, bframes, aq, 0); }
/* --hme with the defaults of x265_param_default (hmeSearchMethod = hex, umh, umh; hmeRange = 16, 32, 48) */
void* ref_la_create_hme(int width, int height, int bframes, int aq) { return la_create(width, height, bframes, aq, 1); }
static void* la_create(int width, int height, int bframes, int aq, int hme)
{
    ensure_init();
    RefLA* h = new RefLA;
    h->param = x265_param_alloc();
    x265_param_default_preset(h->param, "medium", NULL);
    h->param->sourceWidth = width; h->param->sourceHeight = height;
    h->param->fpsNum = 25; h->param->fpsDenom = 1;             /* the CLI always sets them; estimateCUPropagate divides by fpsNum */
    h->param->internalCsp = X265_CSP_I400;
    h->param->internalBitDepth = X265_DEPTH;
    h->param->bframes = bframes;
    h->param->bEnableWeightedPred = 0; h->param->bEnableWeightedBiPred = 0;
    h->param->lookaheadSlices = 0;
    h->param->bEnableHME = hme;
    h->param->rc.aqMode = aq ? X265_AQ_VARIANCE : X265_AQ_NONE;
    h->param->rc.cuTree = 0;
    h->param->logLevel = X265_LOG_NONE;
    h->param->maxSlices = 1;
    h->param->bCopyPicToFrame = 1;
    h->la = new Lookahead(h->param, NULL);
    h->la->create();
    return h;
}

int ref_la_add_frame(void* hv, const void* luma, intptr_t strideElems)
{
    RefLA* h = (RefLA*)hv;
    PicYuv* pic = new PicYuv;
    pic->create(h->param, true);
    x265_picture in;
    x265_picture_init(h->param, &in);
    in.planes[0] = (void*)luma; in.stride[0] = (int)(strideElems * sizeof(pixel));
    in.bitDepth = X265_DEPTH; in.colorSpace = X265_CSP_I400;
    pic->copyFromPicture(in, *h->param, 0, 0);
    Lowres* lr = new Lowres;
    memset((void*)lr, 0, sizeof(Lowres));
    lr->create(h->param, pic, h->param->rc.qgSize);
    lr->init(pic, (int)h->lowres.size());
    h->pics.push_back(pic); h->lowres.push_back(lr); h->framePtrs.push_back(lr);
    return (int)h->lowres.size() - 1;
}

/* geometry: out[0]=lowres width, [1]=lines, [2]=lumaStride, [3]=marginX, [4]=marginY, [5]=widthInCU, [6]=heightInCU,
 * [7]=fullres stride, [8]=fullres marginX, [9]=fullres marginY, [10]=fullres buffer rows */
void ref_la_geometry(void* hv, int64_t* out)
{
    RefLA* h = (RefLA*)hv;
    Lowres* lr = h->lowres[0]; PicYuv* p = h->pics[0];
    out[0] = lr->width; out[1] = lr->lines; out[2] = lr->lumaStride; out[3] = p->m_lumaMarginX; out[4] = p->m_lumaMarginY;
    out[5] = lr->maxBlocksInRow; out[6] = lr->maxBlocksInCol; out[7] = p->m_stride; out[8] = p->m_lumaMarginX; out[9] = p->m_lumaMarginY;
    uint32_t numCuInHeight = (p->m_picHeight + h->param->maxCUSize - 1) / h->param->maxCUSize;
    out[10] = numCuInHeight * h->param->maxCUSize + 2 * p->m_lumaMarginY;
}
/* full padded buffers (start of allocation) */
const void* ref_la_fullres_buffer(void* hv, int idx) { RefLA* h = (RefLA*)hv; return h->pics[idx]->m_picBuf[0]; }
const void* ref_la_lowres_buffer(void* hv, int idx, int k) { RefLA* h = (RefLA*)hv; return h->lowres[idx]->buffer[k]; }
int32_t* ref_la_inv_qscale(void* hv, int idx) { RefLA* h = (RefLA*)hv; return h->lowres[idx]->invQscaleFactor; }

void ref_la_intra(void* hv, int idx)
{
    RefLA* h = (RefLA*)hv;
    h->la->m_tld[0].lowresIntraEstimate(*h->lowres[idx], h->param->rc.qgSize);
}
const int32_t* ref_la_intra_cost(void* hv, int idx) { return ((RefLA*)hv)->lowres[idx]->intraCost; }
const uint8_t* ref_la_intra_mode(void* hv, int idx) { return ((RefLA*)hv)->lowres[idx]->intraMode; }

int64_t ref_la_frame_cost(void* hv, int p0, int p1, int b, int intraPenalty)
{
    RefLA* h = (RefLA*)hv;
    CostEstimateGroup est(*h->la, h->framePtrs.data());
    return est.singleCost(p0, p1, b, !!intraPenalty);
}
/* The cooperative-slice path of estimateFrameCost (slicetype.cpp:3143-3175), which the reference only takes with a thread pool:
 * the slice geometry is set as Lookahead::create sets it (slicetype.cpp:1029-1041), the Coop job is posted as :3150-3160 post it,
 * and the reference's own processTasks (:3049-3113, non-batch branch) runs every slice on this thread -- the same
 * estimateCUCost calls with the same lastRow / slice arguments the pool workers would make, in slice order.  The per-slice sums
 * are then folded as :3166-3173 and the frame score scaled as :3201-3206.  lookaheadSlices is the value the CLI would pass. */
int64_t ref_la_frame_cost_slices(void* hv, int p0, int p1, int b, int lookaheadSlices)
{
    RefLA* h = (RefLA*)hv;
    Lookahead& la = *h->la;
    if (lookaheadSlices > 1)
    {
        la.m_numRowsPerSlice = la.m_8x8Height / lookaheadSlices;
        la.m_numRowsPerSlice = X265_MAX(la.m_numRowsPerSlice, 10);
        la.m_numRowsPerSlice = X265_MIN(la.m_numRowsPerSlice, la.m_8x8Height);
        la.m_numCoopSlices = la.m_8x8Height / la.m_numRowsPerSlice;
    }
    else { la.m_numRowsPerSlice = la.m_8x8Height; la.m_numCoopSlices = 1; }
    h->param->lookaheadSlices = la.m_numCoopSlices;            /* the HME branch of processTasks divides by it */
    CostEstimateGroup est(la, h->framePtrs.data());
    if (la.m_numCoopSlices <= 1) return est.singleCost(p0, p1, b, false);
    Lowres* fenc = h->lowres[b];
    bool doSearch[2];
    doSearch[0] = fenc->lowresMvs[0][b - p0][0].x == 0x7FFF;
    doSearch[1] = p1 > b && fenc->lowresMvs[1][p1 - b][0].x == 0x7FFF;
    fenc->weightedRef[b - p0].isWeighted = false;
    if (h->param->bEnableWeightedPred && doSearch[0])                 /* :3136-3138 */
        la.m_tld[0].weightsAnalyse(*fenc, *h->lowres[p0]);
    fenc->costEst[b - p0][p1 - b] = 0;
    fenc->costEstAq[b - p0][p1 - b] = 0;
    memset(&est.m_slice, 0, sizeof(est.m_slice[0]) * la.m_numCoopSlices);
    est.m_coop.p0 = p0; est.m_coop.p1 = p1; est.m_coop.b = b;
    est.m_coop.bDoSearch[0] = doSearch[0]; est.m_coop.bDoSearch[1] = doSearch[1];
    est.m_jobTotal = la.m_numCoopSlices; est.m_jobAcquired = 0;
    static_cast<BondedTaskGroup&>(est).processTasks(-1);
    for (int i = 0; i < la.m_numCoopSlices; i++)
    {
        fenc->costEst[b - p0][p1 - b] += est.m_slice[i].costEst;
        fenc->costEstAq[b - p0][p1 - b] += est.m_slice[i].costEstAq;
        if (p1 == b) fenc->intraMbs[b - p0] += est.m_slice[i].intraMbs;
    }
    int64_t score = fenc->costEst[b - p0][p1 - b];
    if (b != p1) score = score * 100 / (130 + h->param->bFrameBias);
    fenc->costEst[b - p0][p1 - b] = score;
    est.m_jobTotal = est.m_jobAcquired = 0;
    return score;
}
/* ---- weightp in the lookahead (slicetype.cpp:860-961, hook at :3136-3138) ----
 * ref_la_set_weightp: param->bEnableWeightedPred of the handle (estimateFrameCost then calls weightsAnalyse before a list-0 search).
 * ref_la_set_wp_stats: Lowres::wp_sum[0] / wp_ssd[0] as calcAdaptiveQuantFrame leaves them (:672-674).
 * ref_la_weights_analyse: LookaheadTLD::weightsAnalyse(frames[b], frames[p0]) alone; out = {isWeighted, paddedLines}.
 * ref_la_weighted_buffer: LookaheadTLD::wbuffer[k] (start of the padded plane) of the TLD singleCost uses without a pool. */
void ref_la_set_weightp(void* hv, int on) { ((RefLA*)hv)->param->bEnableWeightedPred = on; }
void ref_la_set_wp_stats(void* hv, int idx, uint64_t sum0, uint64_t ssd0)
{
    Lowres* l = ((RefLA*)hv)->lowres[idx];
    l->wp_sum[0] = sum0; l->wp_ssd[0] = ssd0;
}
void ref_la_weights_analyse(void* hv, int b, int p0, int32_t* out)
{
    RefLA* h = (RefLA*)hv;
    Lowres* fenc = h->lowres[b];
    fenc->weightedRef[b - p0].isWeighted = false;
    h->la->m_tld[0].weightsAnalyse(*fenc, *h->lowres[p0]);
    out[0] = fenc->weightedRef[b - p0].isWeighted; out[1] = h->la->m_tld[0].paddedLines;
}
int ref_la_is_weighted(void* hv, int b, int d0) { return ((RefLA*)hv)->lowres[b]->weightedRef[d0].isWeighted; }
const void* ref_la_weighted_buffer(void* hv, int k) { return ((RefLA*)hv)->la->m_tld[0].wbuffer[k]; }

/* Lookahead::estimateCUPropagate (slicetype.cpp:2641-2747) on the frames of the handle; fps = fpsNum / fpsDenom of the param */
void ref_la_cutree_propagate(void* hv, int p0, int p1, int b, int referenced, double averageDuration)
{
    RefLA* h = (RefLA*)hv;
    h->la->estimateCUPropagate(h->framePtrs.data(), averageDuration, p0, p1, b, referenced);
}
uint16_t* ref_la_propagate_cost(void* hv, int idx) { return ((RefLA*)hv)->lowres[idx]->propagateCost; }
double ref_la_fps_factor(void* hv, double averageDuration)
{
    RefLA* h = (RefLA*)hv;
    double a = (double)h->param->fpsDenom / h->param->fpsNum;
    a = a < 0.01 ? 0.01 : (a > 1.00 ? 1.00 : a);                 /* CLIP_DURATION, ratecontrol.h:45-47 */
    double d = averageDuration < 0.01 ? 0.01 : (averageDuration > 1.00 ? 1.00 : averageDuration);
    return a / d;
}
int ref_la_num_coop_slices(void* hv) { return ((RefLA*)hv)->la->m_numCoopSlices; }
/* outputs cached on frame b (common/lowres.h) */
const int32_t* ref_la_mvs(void* hv, int b, int list, int dist) { return (const int32_t*)((RefLA*)hv)->lowres[b]->lowresMvs[list][dist]; }
const int32_t* ref_la_mvcosts(void* hv, int b, int list, int dist) { return ((RefLA*)hv)->lowres[b]->lowresMvCosts[list][dist]; }
const uint16_t* ref_la_lowres_costs(void* hv, int b, int d0, int d1) { return ((RefLA*)hv)->lowres[b]->lowresCosts[d0][d1]; }
const int32_t* ref_la_row_satds(void* hv, int b, int d0, int d1) { return ((RefLA*)hv)->lowres[b]->rowSatds[d0][d1]; }
int64_t ref_la_cost_est(void* hv, int b, int d0, int d1, int aq) { Lowres* l = ((RefLA*)hv)->lowres[b]; return aq ? l->costEstAq[d0][d1] : l->costEst[d0][d1]; }
int ref_la_intra_mbs(void* hv, int b, int d0) { return ((RefLA*)hv)->lowres[b]->intraMbs[d0]; }

/* --hme: geometry out[0] = m_4x4Width, [1] = m_4x4Height, [2] = hmeSearchMethod[0], [3] = [1], [4] = hmeRange[0], [5] = [1],
 * [6] = element offset of lowerResPlane[0] inside lowerResBuffer[0], [7] = elements between consecutive lower-res planes */
void ref_la_hme_geometry(void* hv, int64_t* out)
{
    RefLA* h = (RefLA*)hv; Lowres* lr = h->lowres[0];
    out[0] = h->la->m_4x4Width; out[1] = h->la->m_4x4Height;
    out[2] = h->param->hmeSearchMethod[0]; out[3] = h->param->hmeSearchMethod[1]; out[4] = h->param->hmeRange[0]; out[5] = h->param->hmeRange[1];
    out[6] = lr->lowerResPlane[0] - lr->lowerResBuffer[0]; out[7] = lr->lowerResBuffer[1] - lr->lowerResBuffer[0];
}
const void* ref_la_lower_buffer(void* hv, int idx) { return ((RefLA*)hv)->lowres[idx]->lowerResBuffer[0]; }
const int32_t* ref_la_lower_mvs(void* hv, int b, int list, int dist) { return (const int32_t*)((RefLA*)hv)->lowres[b]->lowerResMvs[list][dist]; }
const int32_t* ref_la_lower_mvcosts(void* hv, int b, int list, int dist) { return ((RefLA*)hv)->lowres[b]->lowerResMvCosts[list][dist]; }

void ref_la_destroy(void* hv)
{
    RefLA* h = (RefLA*)hv;
    for (auto l : h->lowres) { l->destroy(); delete l; }
    for (auto p : h->pics) { p->destroy(); delete p; }
    h->la->destroy(); delete h->la;
    x265_param_free(h->param);
    delete h;
}

} /* extern "C" */

/* ---- MotionEstimate with the chroma residual term --------------------------------------------------
 * The encode-style setSourcePU (motion.cpp:193-222, the one Search::predInterSearch uses): bChromaSATD is
 * decided by the reference itself (subpelRefine > 2 && chroma satd exists).  The CU Yuv holds the PU at
 * partition 0; the reference picture is a PicYuv shell whose offset tables are {0}, so
 * getCbAddr(0, 0) = fpelPlane[1] = the chroma block under the PU. */
#include "yuv.h"
extern "C" int ref_me_batch_chroma(const void* fencY, const void* fencCb, const void* fencCr, intptr_t fencStrideY, intptr_t fencStrideC,
                                   const void* refY, const void* refCb, const void* refCr, intptr_t refStrideY, intptr_t refStrideC,
                                   int csp, RefMEJob* jobs, int64_t n, int searchMethod, int subpelRefine, int merange, int qp,
                                   int maxSlices, int threads)
{
    ensure_init();
    { BitCost warm; warm.setQP(qp); }
    const int hs = csp != X265_CSP_I444, vs = csp == X265_CSP_I420;
    int nt = threads < 1 ? 1 : threads;
    std::vector<MotionEstimate*> mes(nt);
    std::vector<Yuv*> cus(nt);
    for (int t = 0; t < nt; t++)
    {
        mes[t] = new MotionEstimate;
        mes[t]->init(csp);
        mes[t]->setQP(qp);
        cus[t] = new Yuv;
        if (!cus[t]->create(64, csp)) return -1;
    }
    parallel_for(n, nt, [&](int64_t i, int t) {
        RefMEJob& j = jobs[i];
        MotionEstimate& me = *mes[t];
        Yuv& cu = *cus[t];
        const pixel* fy = (const pixel*)fencY + j.puX + (intptr_t)j.puY * fencStrideY;
        for (int y = 0; y < j.h; y++) memcpy(cu.m_buf[0] + y * cu.m_size, fy + y * fencStrideY, j.w * sizeof(pixel));
        const pixel* fc[2] = { (const pixel*)fencCb, (const pixel*)fencCr };
        for (int c = 0; c < 2; c++)
        {
            const pixel* p = fc[c] + (j.puX >> hs) + (intptr_t)(j.puY >> vs) * fencStrideC;
            for (int y = 0; y < (j.h >> vs); y++) memcpy(cu.m_buf[1 + c] + y * cu.m_csize, p + y * fencStrideC, (j.w >> hs) * sizeof(pixel));
        }
        me.setSourcePU(cu, 0, 0, 0, j.w, j.h, searchMethod == 6 ? X265_HEX_SEARCH : searchMethod, subpelRefine, true);

        intptr_t zero = 0;
        PicYuv pic;
        pic.m_cuOffsetY = pic.m_cuOffsetC = pic.m_buOffsetY = pic.m_buOffsetC = &zero;
        pic.m_stride = refStrideY; pic.m_strideC = refStrideC;
        ReferencePlanes ref;
        ref.reconPic = &pic;
        ref.fpelPlane[0] = (pixel*)refY + j.puX + (intptr_t)j.puY * refStrideY;
        ref.fpelPlane[1] = (pixel*)refCb + (j.puX >> hs) + (intptr_t)(j.puY >> vs) * refStrideC;
        ref.fpelPlane[2] = (pixel*)refCr + (j.puX >> hs) + (intptr_t)(j.puY >> vs) * refStrideC;
        ref.lumaStride = refStrideY; ref.chromaStride = refStrideC;
        ref.isLowres = false;
        MV mvmin(j.mvminX, j.mvminY), mvmax(j.mvmaxX, j.mvmaxY), mvp(j.mvpX, j.mvpY), out(0, 0);
        MV mvc[8];
        for (int k = 0; k < j.numCand; k++) mvc[k] = MV(j.mvc[k][0], j.mvc[k][1]);
        if (searchMethod == 6)
        {
            me.refineMV(&ref, mvmin, mvmax, mvp, out);
            j.outCost = 0;
        }
        else
            j.outCost = me.motionEstimate(&ref, mvmin, mvmax, mvp, j.numCand, mvc, merange, out, (uint32_t)maxSlices);
        j.outMvX = out.x; j.outMvY = out.y;
        pic.m_cuOffsetY = pic.m_cuOffsetC = pic.m_buOffsetY = pic.m_buOffsetC = NULL;
    });
    for (int t = 0; t < nt; t++) { delete mes[t]; cus[t]->destroy(); delete cus[t]; }
    return 0;
}

/* ---- batched table calls for bench.py's CPU arm (the reference's own C primitives over job lists, threaded) ---- */
extern "C" {
struct RefInterpJob { int64_t srcOff, dstOff; int32_t idxX, idxY; };      /* same layout as x265b200_interp_job */
/* luma 8-tap: kind 0 = luma_hpp, 2 = luma_vpp, 6 = luma_hvpp */
void ref_interp_batch(int kind, int part, const void* src, intptr_t srcStride, void* dst, intptr_t dstStride,
                      const RefInterpJob* jobs, int64_t n, int threads)
{
    ensure_init();
    EncoderPrimitives::PU& p = primitives.pu[part];
    parallel_for(n, threads, [&](int64_t i, int) {
        const pixel* s = (const pixel*)src + jobs[i].srcOff; pixel* d = (pixel*)dst + jobs[i].dstOff;
        if (kind == 0) p.luma_hpp(s, srcStride, d, dstStride, jobs[i].idxX);
        else if (kind == 2) p.luma_vpp(s, srcStride, d, dstStride, jobs[i].idxX);
        else p.luma_hvpp(s, srcStride, d, dstStride, jobs[i].idxX, jobs[i].idxY);
    });
}
/* per block: intra_filter, the 33 angular modes (the C table has no intra_pred_allangs, primitives.cpp:257, so the
 * encoder loops over intra_pred[mode], search.cpp:1385-1400), planar and DC.  neigh/filt: n arrays of 4N+1 pixels;
 * dest: n x 35 x N*N pixels (modes 0..34) */
void ref_intra_batch(int sizeIdx, const void* neigh, void* filt, void* dest, int64_t n, int threads)
{
    ensure_init();
    const int N = 4 << sizeIdx, A = 4 * N + 1;
    parallel_for(n, threads, [&](int64_t i, int) {
        const pixel* ref = (const pixel*)neigh + i * A; pixel* f = (pixel*)filt + i * A;
        pixel* out = (pixel*)dest + i * 35 * N * N;
        primitives.cu[sizeIdx].intra_filter(ref, f);
        for (int mode = 0; mode < 35; mode++)
        {
            const pixel* srcPix = (g_intraFilterFlags[mode] & N) ? f : ref;
            primitives.cu[sizeIdx].intra_pred[mode](out + mode * N * N, N, srcPix, mode, N <= 16);
        }
    });
}
void ref_quant_dequant_batch(const int16_t* coef, const int32_t* quantCoeff, int16_t* qCoef, int16_t* deq, int32_t* deltaU,
                             int qBits, int add, int numCoeff, int64_t n, int scale, int shift, int threads)
{
    ensure_init();
    parallel_for(n, threads, [&](int64_t i, int t) {
        primitives.quant(coef + i * numCoeff, quantCoeff, deltaU + (int64_t)t * numCoeff, qCoef + i * numCoeff, qBits, add, numCoeff);
        primitives.dequant_normal(qCoef + i * numCoeff, deq + i * numCoeff, numCoeff, scale, shift);
    });
}
}

/* ---- --me sea: the integral planes exactly as FrameFilter::computeMEIntegral drives the table
 * (framefilter.cpp:722-825: integral_inith per pixel row, integral_initv h rows later), and the reference's own
 * MotionEstimate running X265_SEA against them.  planes[k]: ORIGIN (element of pixel 0,0) of plane k. */
extern "C" {
void ref_sea_integrals(const void* reconOrigin, intptr_t stride, int padX, int padY, int maxHeight, uint32_t* const planes[12])
{
    ensure_init();
    static const int W[12] = { 32, 32, 32, 24, 16, 16, 16, 12, 8, 8, 4, 4 };
    static const int H[12] = { 32, 24, 8, 32, 16, 12, 4, 16, 32, 8, 16, 4 };
    static const int idx[33] = { 0, 0, 0, 0, INTEGRAL_4, 0, 0, 0, INTEGRAL_8, 0, 0, 0, INTEGRAL_12, 0, 0, 0, INTEGRAL_16,
                                 0, 0, 0, 0, 0, 0, 0, INTEGRAL_24, 0, 0, 0, 0, 0, 0, 0, INTEGRAL_32 };
    for (int k = 0; k < 12; k++)
        memset(planes[k] - padY * stride - padX, 0, stride * sizeof(uint32_t));
    const int height = maxHeight + padY - 1;
    for (int y = -padY; y < height; y++)
    {
        pixel* pix = (pixel*)reconOrigin + y * stride - padX;
        for (int k = 0; k < 12; k++)
        {
            uint32_t* sum = planes[k] + (y + 1) * stride - padX;
            primitives.integral_inith[idx[W[k]]](sum, pix, stride);
            if (y >= H[k] - padY)
                primitives.integral_initv[idx[H[k]]](sum - H[k] * stride, stride);
        }
    }
}

/* integralPlanes[12]: pointers addressed like refPlane (element 0 <-> refPlane[0]) */
int ref_me_batch_sea(const void* fencPlane, intptr_t fencStride, const void* refPlane, intptr_t refStride,
                     uint32_t* const integralPlanes[12], RefMEJob* jobs, int64_t n, int subpelRefine, int merange, int qp,
                     int maxSlices, int threads)
{
    ensure_init();
    { BitCost warm; warm.setQP(qp); }
    int nt = threads < 1 ? 1 : threads;
    std::vector<MotionEstimate*> mes(nt);
    for (int t = 0; t < nt; t++)
    {
        mes[t] = new MotionEstimate;
        mes[t]->init(X265_CSP_I400);
        mes[t]->setQP(qp);
    }
    parallel_for(n, nt, [&](int64_t i, int t) {
        RefMEJob& j = jobs[i];
        MotionEstimate& me = *mes[t];
        ReferencePlanes ref;
        ref.fpelPlane[0] = (pixel*)refPlane;
        ref.lumaStride = refStride;
        ref.isLowres = false;
        intptr_t off = j.puX + (intptr_t)j.puY * refStride;
        /* pixels of fencPUYuv outside the PU are stale in the encoder; the backend defines them as 0 */
        memset(me.fencPUYuv.m_buf[0], 0, sizeof(pixel) * FENC_STRIDE * FENC_STRIDE);
        me.setSourcePU((pixel*)fencPlane, fencStride, off, j.w, j.h, X265_SEA, X265_SEA, X265_SEA, subpelRefine);
        for (int k = 0; k < INTEGRAL_PLANE_NUM; k++) me.integral[k] = integralPlanes[k] + off;      /* search.cpp:2264 */
        MV mvmin(j.mvminX, j.mvminY), mvmax(j.mvmaxX, j.mvmaxY), mvp(j.mvpX, j.mvpY), out(0, 0);
        MV mvc[8];
        for (int k = 0; k < j.numCand; k++) mvc[k] = MV(j.mvc[k][0], j.mvc[k][1]);
        j.outCost = me.motionEstimate(&ref, mvmin, mvmax, mvp, j.numCand, mvc, merange, out, (uint32_t)maxSlices);
        j.outMvX = out.x; j.outMvY = out.y;
    });
    for (int t = 0; t < nt; t++) delete mes[t];
    return 0;
}

void ref_integral_inith(int w, uint32_t* sum, const void* pix, intptr_t stride)
{
    ensure_init();
    const int k = w == 4 ? INTEGRAL_4 : w == 8 ? INTEGRAL_8 : w == 12 ? INTEGRAL_12 : w == 16 ? INTEGRAL_16 : w == 24 ? INTEGRAL_24 : INTEGRAL_32;
    primitives.integral_inith[k](sum, (pixel*)pix, stride);
}
void ref_integral_initv(int h, uint32_t* sum, intptr_t stride)
{
    ensure_init();
    const int k = h == 4 ? INTEGRAL_4 : h == 8 ? INTEGRAL_8 : h == 12 ? INTEGRAL_12 : h == 16 ? INTEGRAL_16 : h == 24 ? INTEGRAL_24 : INTEGRAL_32;
    primitives.integral_initv[k](sum, stride);
}
} /* extern "C" */

/* ---- the residual pipeline of one TU exactly as the encoder chains the table entries: Search (sub_ps) ->
 * Quant::transformNxN (quant.cpp:397-480, rdoq 0 / no sign hiding) -> Quant::invtransformNxN (quant.cpp:543-605) ->
 * add_ps -> sse_pp, over the blocksX x blocksY TU grid of a plane.  dequantCoef == NULL selects dequant_normal. */
extern "C" void ref_tu_pipeline(int sizeIdx, int useDST, const void* fencV, intptr_t fencStride, const void* predV, intptr_t predStride,
                                void* reconV, intptr_t reconStride, int blocksX, int blocksY, const int32_t* quantCoeff, int qBits, int add,
                                const int32_t* dequantCoef, int scaleOrPer, int dqShift, int16_t* coeff, uint32_t* numSigOut, uint64_t* sseOut,
                                int threads)
{
    ensure_init();
    const int N = 4 << sizeIdx, numCoeff = N * N;
    parallel_for((int64_t)blocksX * blocksY, threads, [&](int64_t b, int) {
        ALIGN_VAR_32(int16_t, resi[32 * 32]);
        ALIGN_VAR_32(int16_t, dctc[32 * 32]);
        ALIGN_VAR_32(int32_t, deltaU[32 * 32]);
        const int by = (int)(b / blocksX), bx = (int)(b % blocksX);
        const pixel* fenc = (const pixel*)fencV + (intptr_t)by * N * fencStride + bx * N;
        const pixel* pred = (const pixel*)predV + (intptr_t)by * N * predStride + bx * N;
        pixel* recon = (pixel*)reconV + (intptr_t)by * N * reconStride + bx * N;
        int16_t* qc = coeff + b * numCoeff;
        primitives.cu[sizeIdx].sub_ps(resi, N, fenc, pred, fencStride, predStride);
        if (useDST) primitives.dst4x4(resi, dctc, N); else primitives.cu[sizeIdx].dct(resi, dctc, N);
        uint32_t numSig = primitives.quant(dctc, quantCoeff, deltaU, qc, qBits, add, numCoeff);
        if (numSig)
        {
            if (dequantCoef) primitives.dequant_scaling(qc, dequantCoef, dctc, numCoeff, scaleOrPer, dqShift);
            else primitives.dequant_normal(qc, dctc, numCoeff, scaleOrPer, dqShift);
            if (numSig == 1 && qc[0] != 0 && !useDST)
            {
                const int shift_1st = 7 - 6, add_1st = 1 << (shift_1st - 1);
                const int shift_2nd = 12 - (X265_DEPTH - 8) - 3, add_2nd = 1 << (shift_2nd - 1);
                int dc_val = (((dctc[0] * (64 >> 6) + add_1st) >> shift_1st) * (64 >> 3) + add_2nd) >> shift_2nd;
                primitives.cu[sizeIdx].blockfill_s[0](resi, N, (int16_t)dc_val);
            }
            else if (useDST) primitives.idst4x4(dctc, resi, N);
            else primitives.cu[sizeIdx].idct(dctc, resi, N);
        }
        else primitives.cu[sizeIdx].blockfill_s[0](resi, N, 0);
        primitives.cu[sizeIdx].add_ps[0](recon, reconStride, pred, resi, predStride, N);
        numSigOut[b] = numSig;
        sseOut[b] = (uint64_t)primitives.cu[sizeIdx].sse_pp(fenc, fencStride, recon, reconStride);
    });
}

/* ---- motion compensation: the reference's OWN Predict::motionCompensation (common/predict.cpp:77-257) driven through
 * shell CUData / Slice / PicYuv objects (offset tables = {0}, so getLumaAddr(0, 0) is the plane pointer we set per job). */
#include "predict.h"
#include "cudata.h"
#include "slice.h"
#include "framedata.h"
#include "shortyuv.h"

extern "C" {
struct RefMCJob { int32_t puX, puY, w, h, cuX, cuY; int32_t refIdx[2]; int32_t mv[2][2]; };
struct RefMCWeight { int32_t w, o, shift, present; };

/* refs: [2][maxRefs][3] plane ORIGINS; weights: [2][maxRefs][3] or NULL; pred*: output planes (PU written at its position) */
int ref_mc_batch(int csp, int isP, int wpP, int wpB, int picW, int picH, int maxCU, int maxRefs,
                 const void* const* refs, intptr_t refStrideY, intptr_t refStrideC,
                 void* predY, void* predCb, void* predCr, intptr_t predStrideY, intptr_t predStrideC,
                 const RefMCWeight* weights, const RefMCJob* jobs, int64_t n, int bLuma, int bChroma)
{
    ensure_init();
    if (maxRefs > MAX_NUM_REF) return -1;
    const int hs = CHROMA_H_SHIFT(csp), vs = CHROMA_V_SHIFT(csp);
    x265_param param;
    x265_param_default(&param);
    param.maxCUSize = maxCU; param.internalCsp = csp;
    SPS sps; memset(&sps, 0, sizeof(sps));
    sps.picWidthInLumaSamples = picW; sps.picHeightInLumaSamples = picH;
    PPS pps; memset(&pps, 0, sizeof(pps));
    pps.bUseWeightPred = !!wpP; pps.bUseWeightedBiPred = !!wpB;
    Slice slice;
    slice.m_sps = &sps; slice.m_pps = &pps; slice.m_sliceType = isP ? P_SLICE : B_SLICE;
    slice.m_numRefIdx[0] = slice.m_numRefIdx[1] = maxRefs;
    FrameData fd;
    fd.m_param = &param; fd.m_slice = &slice;
    intptr_t zero = 0;
    std::vector<PicYuv> pics(2 * maxRefs);
    for (int l = 0; l < 2; l++)
        for (int r = 0; r < maxRefs; r++)
        {
            PicYuv& pic = pics[l * maxRefs + r];
            pic.m_cuOffsetY = pic.m_cuOffsetC = pic.m_buOffsetY = pic.m_buOffsetC = &zero;
            pic.m_stride = refStrideY; pic.m_strideC = refStrideC;
            slice.m_refReconPicList[l][r] = &pic;
            for (int p = 0; p < 3; p++)
            {
                WeightParam& wp = slice.m_weightPredTable[l][r][p];
                if (weights)
                {
                    const RefMCWeight& w = weights[(l * maxRefs + r) * 3 + p];
                    wp.inputWeight = w.w; wp.inputOffset = w.o; wp.log2WeightDenom = (uint32_t)w.shift; wp.wtPresent = w.present;
                }
                else { wp.inputWeight = 1; wp.inputOffset = 0; wp.log2WeightDenom = 0; wp.wtPresent = 0; }
            }
        }
    Predict pred;
    if (!pred.allocBuffers(csp)) return -1;
    Yuv predYuv;
    if (!predYuv.create(64, csp)) return -1;
    CUData cu;
    int8_t refIdx[2][4]; MV mv[2][4];
    cu.m_slice = &slice; cu.m_encData = &fd;
    cu.m_refIdx[0] = refIdx[0]; cu.m_refIdx[1] = refIdx[1]; cu.m_mv[0] = mv[0]; cu.m_mv[1] = mv[1];
    alignas(16) unsigned char puStore[sizeof(PredictionUnit)];
    PredictionUnit& pu = *(PredictionUnit*)puStore;
    for (int64_t i = 0; i < n; i++)
    {
        const RefMCJob& j = jobs[i];
        cu.m_cuPelX = j.cuX; cu.m_cuPelY = j.cuY;
        for (int l = 0; l < 2; l++) { refIdx[l][0] = (int8_t)j.refIdx[l]; mv[l][0] = MV(j.mv[l][0], j.mv[l][1]); }
        const intptr_t offY = j.puX + (intptr_t)j.puY * refStrideY, offC = (j.puX >> hs) + (intptr_t)(j.puY >> vs) * refStrideC;
        for (int l = 0; l < 2; l++)
            for (int r = 0; r < maxRefs; r++)
            {
                PicYuv& pic = pics[l * maxRefs + r];
                const void* const* pl = refs + (l * maxRefs + r) * 3;
                pic.m_picOrg[0] = (pixel*)pl[0] + offY;
                pic.m_picOrg[1] = pl[1] ? (pixel*)pl[1] + offC : NULL;
                pic.m_picOrg[2] = pl[2] ? (pixel*)pl[2] + offC : NULL;
            }
        pu.ctuAddr = 0; pu.cuAbsPartIdx = 0; pu.puAbsPartIdx = 0; pu.width = j.w; pu.height = j.h;
        pred.motionCompensation(cu, pu, predYuv, !!bLuma, !!bChroma && csp != X265_CSP_I400);
        if (bLuma)
            for (int y = 0; y < j.h; y++)
                memcpy((pixel*)predY + j.puX + (intptr_t)(j.puY + y) * predStrideY, predYuv.m_buf[0] + y * predYuv.m_size, j.w * sizeof(pixel));
        if (bChroma && csp != X265_CSP_I400)
            for (int c = 0; c < 2; c++)
            {
                pixel* d = (pixel*)(c ? predCr : predCb) + (j.puX >> hs) + (intptr_t)(j.puY >> vs) * predStrideC;
                for (int y = 0; y < (j.h >> vs); y++)
                    memcpy(d + (intptr_t)y * predStrideC, predYuv.m_buf[1 + c] + y * predYuv.m_csize, (j.w >> hs) * sizeof(pixel));
            }
    }
    for (auto& pic : pics) { pic.m_cuOffsetY = pic.m_cuOffsetC = pic.m_buOffsetY = pic.m_buOffsetC = NULL; pic.m_picOrg[0] = pic.m_picOrg[1] = pic.m_picOrg[2] = NULL; }
    cu.m_refIdx[0] = cu.m_refIdx[1] = NULL; cu.m_mv[0] = cu.m_mv[1] = NULL;
    predYuv.destroy();
    return 0;
}
} /* extern "C" */

extern "C" {
float ref_ssim_end4(int (*sum0)[4], int (*sum1)[4], int width) { ensure_init(); return primitives.ssim_end_4(sum0, sum1, width); }
void ref_ssim_core(const void* p1, intptr_t s1, const void* p2, intptr_t s2, int (*sums)[4]) { ensure_init(); primitives.ssim_4x4x2_core((const pixel*)p1, s1, (const pixel*)p2, s2, sums); }
int ref_plane_clip_max(void* src, intptr_t stride, int width, int height, uint64_t* outsum, int minPix, int maxPix)
{
    ensure_init();
    if (!primitives.planeClipAndMax) return -1;
    return (int)primitives.planeClipAndMax((pixel*)src, stride, width, height, outsum, (pixel)minPix, (pixel)maxPix);
}
void ref_propagate_cost(int* dst, const uint16_t* propagateIn, const int32_t* intraCosts, const uint16_t* interCosts, const int32_t* invQscales, double fps, int len)
{ ensure_init(); primitives.propagateCost(dst, propagateIn, intraCosts, interCosts, invQscales, &fps, len); }
}

/* ---- SAO / deblock table entries, one call each (kinds as X265B200_SAO_*: 0 E0, 1 E1, 2 E1_2Rows, 3 E2, 4 E3, 5 B0 / BO) ---- */
extern "C" {
void ref_sao_apply(int kind, void* recV, intptr_t stride, int8_t* buf0, int8_t* buf1, int8_t* offsets, int width, int height, int startX)
{
    ensure_init();
    pixel* rec = (pixel*)recV;
    switch (kind)
    {
    case 0: primitives.saoCuOrgE0(rec, offsets, width, buf0, stride); break;
    case 1: primitives.saoCuOrgE1(rec, buf0, offsets, stride, width); break;
    case 2: primitives.saoCuOrgE1_2Rows(rec, buf0, offsets, stride, width); break;
    case 3: primitives.saoCuOrgE2[0](rec, buf0, buf1, offsets, width, stride); break;
    case 4: primitives.saoCuOrgE3[0](rec, buf0, offsets, stride, startX, width); break;
    default: primitives.saoCuOrgB0(rec, offsets, width, height, stride); break;
    }
}
void ref_sao_stats(int kind, const int16_t* diff, const void* recV, intptr_t stride, int8_t* up1, int8_t* upt, int endX, int endY, int32_t* stats, int32_t* count)
{
    ensure_init();
    const pixel* rec = (const pixel*)recV;
    switch (kind)
    {
    case 0: primitives.saoCuStatsE0(diff, rec, stride, endX, endY, stats, count); break;
    case 1: primitives.saoCuStatsE1(diff, rec, stride, up1, endX, endY, stats, count); break;
    case 3: primitives.saoCuStatsE2(diff, rec, stride, up1, upt, endX, endY, stats, count); break;
    case 4: primitives.saoCuStatsE3(diff, rec, stride, up1, endX, endY, stats, count); break;
    default: primitives.saoCuStatsBO(diff, rec, stride, endX, endY, stats, count); break;
    }
}
void ref_deblock(int chroma, void* src, intptr_t srcStep, intptr_t offset, int a, int b, int c)
{
    ensure_init();
    if (chroma) primitives.pelFilterChroma[0]((pixel*)src, srcStep, offset, a, b, c);
    else primitives.pelFilterLumaStrong[0]((pixel*)src, srcStep, offset, a, b);
}
}


/* ---- AQ energies and weighted reference planes: the reference's own LookaheadTLD::acEnergyCu and MotionReference::applyWeight ---- */
#include "reference.h"
extern "C" {
struct RefPic { x265_param* param; PicYuv* pic; Frame* frame; };
/* a PicYuv (with the encoder's padding) from planar input; csp as x265.h:588-592 */
void* ref_pic_create(int width, int height, int csp, const void* y, const void* cb, const void* cr, intptr_t strideY, intptr_t strideC)
{
    ensure_init();
    RefPic* h = new RefPic;
    h->param = x265_param_alloc();
    x265_param_default_preset(h->param, "medium", NULL);
    h->param->sourceWidth = width; h->param->sourceHeight = height;
    h->param->internalCsp = csp; h->param->internalBitDepth = X265_DEPTH; h->param->logLevel = X265_LOG_NONE;
    h->param->bCopyPicToFrame = 1;
    h->pic = new PicYuv;
    h->pic->create(h->param, true);
    x265_picture in;
    x265_picture_init(h->param, &in);
    in.planes[0] = (void*)y; in.planes[1] = (void*)cb; in.planes[2] = (void*)cr;
    in.stride[0] = (int)(strideY * sizeof(pixel)); in.stride[1] = in.stride[2] = (int)(strideC * sizeof(pixel));
    in.bitDepth = X265_DEPTH; in.colorSpace = csp;
    h->pic->copyFromPicture(in, *h->param, 0, 0);
    h->frame = new Frame;
    h->frame->m_fencPic = h->pic;
    h->frame->m_param = h->param;
    return h;
}
/* out[0] = stride, [1] = strideC, [2] = lumaMarginX, [3] = lumaMarginY, [4] = chromaMarginX, [5] = chromaMarginY, [6] = luma rows allocated, [7] = chroma rows */
void ref_pic_geometry(void* hv, int64_t* out)
{
    RefPic* h = (RefPic*)hv; PicYuv* p = h->pic;
    uint32_t nh = (p->m_picHeight + h->param->maxCUSize - 1) / h->param->maxCUSize;
    out[0] = p->m_stride; out[1] = p->m_strideC; out[2] = p->m_lumaMarginX; out[3] = p->m_lumaMarginY; out[4] = p->m_chromaMarginX; out[5] = p->m_chromaMarginY;
    out[6] = nh * h->param->maxCUSize + 2 * p->m_lumaMarginY; out[7] = ((nh * h->param->maxCUSize) >> p->m_vChromaShift) + 2 * p->m_chromaMarginY;
}
const void* ref_pic_buffer(void* hv, int plane) { return ((RefPic*)hv)->pic->m_picBuf[plane]; }
/* acEnergyCu over the block loop of calcAdaptiveQuantFrame (slicetype.cpp:519-523); wp[0..2] = wp_sum, wp[3..5] = wp_ssd */
void ref_aq_energy(void* hv, int qgSize, uint32_t* energy, uint64_t* wp)
{
    RefPic* h = (RefPic*)hv;
    LookaheadTLD tld;
    for (int i = 0; i < 3; i++) h->frame->m_lowres.wp_sum[i] = h->frame->m_lowres.wp_ssd[i] = 0;
    int k = 0;
    for (int by = 0; by < (int)h->pic->m_picHeight; by += qgSize)
        for (int bx = 0; bx < (int)h->pic->m_picWidth; bx += qgSize)
            energy[k++] = tld.acEnergyCu(h->frame, bx, by, h->param->internalCsp, qgSize);
    for (int i = 0; i < 3; i++) { wp[i] = h->frame->m_lowres.wp_sum[i]; wp[3 + i] = h->frame->m_lowres.wp_ssd[i]; }
}
/* MotionReference::init + applyWeight over all rows (reference.cpp:51-185); outFull receives the whole padded weighted luma buffer */
int ref_apply_weight(void* hv, int inputWeight, int inputOffset, int log2Denom, void* outFull)
{
    RefPic* h = (RefPic*)hv;
    x265_param p = *h->param;
    p.subpelRefine = 2; p.maxSlices = 1;
    WeightParam wp[3];
    memset(wp, 0, sizeof(wp));
    wp[0].wtPresent = 1; wp[0].inputWeight = inputWeight; wp[0].inputOffset = inputOffset; wp[0].log2WeightDenom = log2Denom;
    MotionReference ref;
    if (ref.init(h->pic, wp, p)) return -1;
    const uint32_t rows = (h->pic->m_picHeight + p.maxCUSize - 1) / p.maxCUSize;
    ref.applyWeight(rows - 1, rows, rows, 0);
    const size_t padheight = rows * p.maxCUSize + 2 * h->pic->m_lumaMarginY;
    memcpy(outFull, ref.weightBuffer[0], (size_t)h->pic->m_stride * padheight * sizeof(pixel));
    return 0;
}
void ref_pic_destroy(void* hv)
{
    RefPic* h = (RefPic*)hv;
    h->frame->m_fencPic = NULL;
    delete h->frame;
    h->pic->destroy(); delete h->pic;
    x265_param_free(h->param);
    delete h;
}
} /* extern "C" */

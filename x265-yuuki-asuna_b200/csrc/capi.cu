// capi.cu -- the extern "C" shim of libx265b200.so (declared in include/x265b200.h).
// Context/memory management plus the *_dev / *_host entry points that forward to the
// kernel launchers in the other translation units.  No CPU fallback: every compute entry
// point needs a live CUDA context and fails loudly otherwise.
#include "common.cuh"
#include "x265b200.h"
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <cmath>
#include <vector>
#include <cuda.h>          // types of the green-context driver API only: entry points are fetched at run time (no link against libcuda)

namespace x265b200 {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...)
{
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check(cudaError_t e, const char* what)
{
    if (e == cudaSuccess) return 0;
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return -1;
}

// scratch management for host-pointer entry points ------------------------------------------
int scratch_dev(Ctx* ctx, int slot, size_t bytes, void** out)
{
    if (ctx->dScratchBytes[slot] < bytes)
    {
        if (ctx->dScratch[slot]) { X265B200_CHECK(cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->dScratch[slot]); ctx->dScratch[slot] = nullptr; ctx->dScratchBytes[slot] = 0; }
        size_t cap = bytes + bytes / 4 + 256;
        X265B200_CHECK(cudaMalloc(&ctx->dScratch[slot], cap));
        ctx->dScratchBytes[slot] = cap;
    }
    *out = ctx->dScratch[slot];
    return 0;
}

int scratch_pinned(Ctx* ctx, int slot, size_t bytes, void** out)
{
    if (ctx->hPinnedBytes[slot] < bytes)
    {
        if (ctx->hPinned[slot]) { X265B200_CHECK(cudaStreamSynchronize(ctx->stream)); cudaFreeHost(ctx->hPinned[slot]); ctx->hPinned[slot] = nullptr; ctx->hPinnedBytes[slot] = 0; }
        size_t cap = bytes + bytes / 4 + 256;
        X265B200_CHECK(cudaMallocHost(&ctx->hPinned[slot], cap));
        ctx->hPinnedBytes[slot] = cap;
    }
    *out = ctx->hPinned[slot];
    return 0;
}

// Asynchronous upload of a small host temporary (the caller may free or reuse hostSrc as soon as this returns): the bytes are
// copied into one of four pinned staging buffers and travel from there.  A buffer is reused only after the copy that last read
// it has completed (its event; in practice long ago), so no call waits for the stream's other work.
int stage_small(Ctx* ctx, void* dDst, const void* hostSrc, size_t bytes)
{
    if (!bytes) return 0;
    const int k = ctx->hStageNext++ & 3;
    if (ctx->hStageEv[k]) X265B200_CHECK(cudaEventSynchronize(ctx->hStageEv[k]));
    else X265B200_CHECK(cudaEventCreateWithFlags(&ctx->hStageEv[k], cudaEventDisableTiming));
    if (ctx->hStageBytes[k] < bytes)
    {
        if (ctx->hStage[k]) cudaFreeHost(ctx->hStage[k]);
        ctx->hStage[k] = nullptr; ctx->hStageBytes[k] = 0;
        const size_t cap = bytes + bytes / 2 + 4096;
        X265B200_CHECK(cudaMallocHost(&ctx->hStage[k], cap));
        ctx->hStageBytes[k] = cap;
    }
    memcpy(ctx->hStage[k], hostSrc, bytes);
    X265B200_CHECK(cudaMemcpyAsync(dDst, ctx->hStage[k], bytes, cudaMemcpyHostToDevice, ctx->stream));
    X265B200_CHECK(cudaEventRecord(ctx->hStageEv[k], ctx->stream));
    return 0;
}

// H2D of an arbitrary (possibly pageable) host buffer into scratch slot `slot`
int stage_in(Ctx* ctx, int slot, const void* host, size_t bytes, void** dev)
{
    if (scratch_dev(ctx, slot, bytes ? bytes : 1, dev)) return -1;
    if (bytes) X265B200_CHECK(cudaMemcpyAsync(*dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

// forward declarations of launchers -----------------------------------------------------------
int pixelcmp_dev(Ctx*, int kind, int depth, int w, int h, const void* A, int64_t strideA, const void* B, int64_t strideB,
                 const int64_t* offA, const int64_t* offB, const int16_t* mv, int gridCols, int64_t n, void* out);
int sad_pyramid_dev(Ctx*, int depth, const void* cur, int64_t strideC, const void* const* refs, int numRefs, int64_t strideR, int ctuCols, int ctuRows,
                    const int16_t* mvCtu, int32_t* out8, int32_t* out16, int32_t* out32, int32_t* out64);
int sad_xn_dev(Ctx*, int depth, int K, int w, int h, const void* fenc, int64_t fencBlockStride,
               const void* ref, int64_t refStride, const int64_t* refOff, int64_t n, int32_t* res);

int transform_dev(Ctx*, int inverse, int sizeIdx, int depth, const int16_t* src, int64_t srcBlockStride, int64_t srcStride,
                  int16_t* dst, int64_t dstBlockStride, int64_t dstStride, int64_t n, int64_t blocksPerRow, int64_t rowStride);
int quant_dev(Ctx*, int isN, const int16_t* coef, const int32_t* quantCoeff, int32_t* deltaU, int16_t* qCoef,
              int qBits, int add, int numCoeff, int64_t n, uint32_t* numSig);
int dequant_dev(Ctx*, int scaling, const int16_t* q, const int32_t* deqCoef, int16_t* coef, int num, int64_t n, int scaleOrPer, int shift);
int count_nonzero_dev(Ctx*, const int16_t* q, int numCoeff, int64_t n, int32_t* out);
int tu_pipeline_dev(Ctx*, int sizeIdx, int depth, int useDST, const void* fenc, int64_t fencStride, const void* pred, int64_t predStride,
                    void* recon, int64_t reconStride, int blocksX, int blocksY, const int32_t* quantCoeff, int qBits, int add,
                    const int32_t* dequantCoef, int scaleOrPer, int dqShift, int16_t* coeff, uint32_t* numSig, uint64_t* sse);
void host_dct_table(int N, int16_t* out);
int interp_dev(Ctx*, int kind, int taps, int depth, int w, int h, const void* src, int64_t srcStride,
               void* dst, int64_t dstStride, const x265b200_interp_job* jobs, int64_t n, int isRowExt);
int interp_multi_dev(Ctx*, int taps, int depth, const x265b200_interp_seg* segs, int numSegs);
int intra_modes_dev(Ctx*, int depth, int log2N, const void* neighbours, void* dest, int bLuma, int64_t n);
int mc_dev(Ctx*, int depth, const x265b200_mc_desc* d, const x265b200_mc_job* jobs, int64_t n, int bLuma, int bChroma);
int sao_apply_dev(Ctx*, int kind, int depth, void* rec, int64_t stride, const x265b200_sao_job* jobs, int64_t n, int8_t* signBuf, const int8_t* offsets, int maxWidth);
int sao_stats_dev(Ctx*, int kind, int depth, const int16_t* diff, const void* rec, int64_t stride, const x265b200_sao_job* jobs, int64_t n,
                  int8_t* signBuf, int32_t* stats, int32_t* count);
int sign_dev(Ctx*, int depth, int8_t* dst, const void* src1, const void* src2, int64_t n);
int deblock_dev(Ctx*, int chroma, int depth, void* pic, const x265b200_deblock_job* jobs, int64_t n);
int propagate_cost_dev(Ctx*, int* dst, const uint16_t* propagateIn, const int32_t* intraCosts, const uint16_t* interCosts, const int32_t* invQscales, double fpsFactor, int64_t len);
int fix8_dev(Ctx*, int unpack, void* dst, const void* src, int64_t n);
int planecopy_dev(Ctx*, int mode, int depth, const void* src, int64_t srcStride, void* dst, int64_t dstStride, int width, int height, int shift, int mask);
int ssim_dist_dev(Ctx*, int depth, int log2TrSize, const void* fenc, int64_t fStride, const void* recon, int64_t rStride, const int64_t* offF, const int64_t* offR,
                  int64_t n, int shift, uint64_t* ssBlock, uint64_t* ack);
int norm_fact_dev(Ctx*, int depth, const void* src, const int64_t* off, int64_t n, int blockSize, int shift, uint64_t* zk);
int ssim_core_dev(Ctx*, int depth, const void* p1, int64_t s1, const void* p2, int64_t s2, const int64_t* off1, const int64_t* off2, int64_t n, int32_t* sums);
int ssim_end4_dev(Ctx*, int depth, const int32_t* sum0, const int32_t* sum1, const int32_t* widths, int64_t n, float* out);
int plane_clip_max_dev(Ctx*, int depth, void* src, int64_t stride, int width, int height, int minPix, int maxPix, uint64_t* outsum, uint32_t* outmax);
int intra_pred_dev(Ctx*, int depth, int log2N, const void* nbr, void* dst, int64_t dstStride, const x265b200_intra_job* jobs, int64_t n);
int intra_filter_dev(Ctx*, int depth, int log2N, const void* src, void* dst, int64_t n);
int intra_allangs_dev(Ctx*, int depth, int log2N, const void* refPix, const void* filtPix, void* dest, int bLuma, int64_t n);

int me_batch_dev(Ctx*, int depth, const void* fencPlane, int64_t fencStride, const void* refPlane, const void* const* refPlanes,
                 int64_t refStride, const x265b200_me_chroma* chroma, const uint32_t* const* seaPlanes, x265b200_me_job* jobs, int64_t n,
                 int maxW, int maxH, int searchMethod, int subpelRefine, int merange, double lambda, int maxSlices);
int sea_integral_dev(Ctx*, int depth, const void* reconOrigin, int64_t stride, int padX, int padY, int maxHeight, uint32_t* const planes[12]);
int integral_inith_dev(Ctx*, int depth, int w, uint32_t* sum, const void* pix, int64_t stride);
int integral_initv_dev(Ctx*, int h, uint32_t* sum, int64_t stride);
int ads_dev(Ctx*, int kind, int lxHalf, const uint32_t* sums, int64_t delta, const uint16_t* costMvX, int width,
            const x265b200_ads_job* jobs, int64_t n, int16_t* mvs, int32_t* counts);
void host_bitcost_table(double lambda, uint16_t* out);
int glue_dev(Ctx*, int op, int depth, int w, int h, void* dst, int64_t dstStride, const void* src0, int64_t src0Stride,
             const void* src1, int64_t src1Stride, const x265b200_glue_job* jobs, int64_t n, int p0, int p1, int p2, int p3);
int var_dev(Ctx*, int depth, int size, const void* src, int64_t stride, const int64_t* off, int64_t n, uint64_t* out);
int psycost_dev(Ctx*, int depth, int dim, const void* src, int64_t sstride, const void* rec, int64_t rstride,
                const int64_t* offS, const int64_t* offR, int64_t n, int32_t* out);
int copy_cnt_dev(Ctx*, int size, int16_t* coeff, const int16_t* resi, int64_t stride, const int64_t* off, int64_t n, uint32_t* numSig);
int denoise_dct_dev(Ctx*, int16_t* coef, uint32_t* resSum, const uint16_t* offset, int numCoeff, int64_t n);
int lowpass_front_dev(Ctx*, const int16_t* src, int64_t srcBlockStride, int64_t srcStride, int64_t n, int N, int16_t* avg, int32_t* total);
int lowpass_back_dev(Ctx*, const int16_t* coefHalf, const int32_t* total, int64_t n, int N, int16_t* dst);
int me_frame_ex_dev(Ctx*, const x265b200_me_frame_params* P, const x265b200_me_frame_planes* pl, const int32_t* mvpCtu, const int32_t* mvpPu,
                    const uint8_t* numCandPu, const int32_t* mvcPu, int32_t* out);
int sad_stream_dev(Ctx*, int depth, const void* poolOrigin, int64_t framePitch, int64_t stride, int marginX, int marginY, int rowsTotal, int numFrames,
                   int ctuCols, int ctuRows, const x265b200_sad_group* groupsHost, int numGroups, int numRefs,
                   int32_t* out8, int32_t* out16, int32_t* out32, int32_t* out64);
int cutree_propagate_dev(Ctx*, int widthInCU, int heightInCU, const uint16_t* propagateCostB, const int32_t* intraCost, const uint16_t* lowresCosts,
                         const int32_t* invQscale, const int32_t* mvs0, const int32_t* mvs1, uint16_t* refCost0, uint16_t* refCost1,
                         int bipredWeight, double fpsFactor);
int aq_energy_dev(Ctx*, int depth, int csp, int qgSize, const void* y, int64_t strideY, const void* cb, const void* cr, int64_t strideC,
                  int picWidth, int picHeight, uint32_t* energy, uint64_t* wpSumSsd);
int apply_weight_dev(Ctx*, int depth, const void* srcOrigin, void* dstOrigin, int64_t stride, int width, int height, int marginX, int marginY,
                     int inputWeight, int inputOffset, int log2WeightDenom);
int me_frame_layout(int ctuSize, int minCuSize, int rect, int amp, int32_t* outXYWH, int cap);
void me_ctu_release(Ctx* ctx);
int me_frame_dev(Ctx*, int depth, const void* curOrigin, int64_t curStride, const void* const* refOriginsHost, int numRefs, int64_t refStride,
                 int marginX, int marginY, int rowsTotal, int ctuCols, int ctuRows, int puMask, const int32_t* mvpCtu,
                 int searchMethod, int subpelRefine, int merange, double lambda, int32_t* out);
int lowres_init_dev(Ctx*, int depth, const void* src, int64_t srcStride, void* const planes[4], int64_t dstStride, int width, int height, int marginX, int marginY);
int extend_border_dev(Ctx*, int depth, void* origin, int64_t stride, int width, int height, int marginX, int marginY);
int la_intra_dev(Ctx*, int depth, const void* plane0, int64_t stride, int widthInCU, int heightInCU, const int32_t* invQscale,
                 int intraPenalty, int32_t* intraCost, uint8_t* intraMode, uint16_t* lowresCosts, int32_t* rowSatds, int32_t* sums);
int la_estimate_dev(Ctx*, int depth, const void* const* planes, int64_t stride, int widthInCU, int heightInCU,
                    const x265b200_la_triple* triplesHost, int numTriples, int32_t* mvPool, int32_t* mvCostPool,
                    const int32_t* const* intraCost, const int32_t* const* invQscale, uint16_t* lowresCosts, int32_t* rowSatds, int32_t* sums,
                    double lambda, int maxSlices, int lookaheadSlices, const x265b200_la_hme* hme, const x265b200_la_weight* weights);
void la_weight_guess(int depth, int width, int lines, uint64_t fencSum, uint64_t fencSsd, uint64_t refSum, uint64_t refSsd, int out[7]);
int la_weights_analyse_dev(Ctx*, int depth, const x265b200_la_weight_job* jobsHost, int numJobs, int64_t stride, int paddedLines, int64_t padOffset,
                           int width, int lines, x265b200_la_weight* out);
int sub_ps_plane_dev(Ctx*, int depth, const void* a, int64_t strideA, const void* b, int64_t strideB, int16_t* dst, int64_t dstStride, int w, int h);
int add_ps_plane_dev(Ctx*, int depth, void* dst, int64_t dstStride, const void* pred, int64_t predStride, const int16_t* resi, int64_t resiStride, int w, int h);

} // namespace x265b200

using namespace x265b200;

struct x265b200_ctx { Ctx c; };

#define CTX(ctx) (&(ctx)->c)
#define REQUIRE_CTX(ctx) do { if (!(ctx)) { set_error("null x265b200 context (CUDA backend not initialised; there is no CPU fallback)"); return -1; } \
                              if (check(cudaSetDevice((ctx)->c.device), "cudaSetDevice")) return -1; } while (0)

// ---- SM partitioning through green contexts (driver API >= 12.4, fetched with cudaGetDriverEntryPoint) ----------------------------
namespace {
template<typename F> bool drv_entry(const char* name, F& fn)
{
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) { cudaGetLastError(); return false; }
    fn = (F)p;
    return true;
}
struct GreenPartition { CUgreenCtx ctx[2]; CUstream stream[2]; };
std::mutex g_partMutex;
std::vector<GreenPartition> g_partitions;

// number of SMs the stream's kernels may run on: its green context's share, else the whole device
int stream_sm_count(cudaStream_t stream, int deviceSms)
{
    CUresult (*getGreen)(CUstream, CUgreenCtx*) = nullptr;
    CUresult (*getRes)(CUgreenCtx, CUdevResource*, CUdevResourceType) = nullptr;
    if (!stream || !drv_entry("cuStreamGetGreenCtx", getGreen) || !drv_entry("cuGreenCtxGetDevResource", getRes)) return deviceSms;
    CUgreenCtx g = nullptr;
    if (getGreen((CUstream)stream, &g) != CUDA_SUCCESS || !g) return deviceSms;
    CUdevResource r;
    if (getRes(g, &r, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS || !r.sm.smCount) return deviceSms;
    return (int)r.sm.smCount;
}
}


extern "C" {

int x265b200_version(void) { return 100; }
const char* x265b200_last_error(void) { return g_err; }

int x265b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int x265b200_sm_partition(int device, int smsFirst, void** streamFirst, void** streamRest, int smsGot[2])
{
    if (!streamFirst || !streamRest || smsFirst <= 0) { set_error("x265b200_sm_partition: bad arguments"); return -1; }
    int n = x265b200_device_count();
    if (device < 0 || device >= n) { set_error("x265b200_sm_partition: device %d out of range [0,%d)", device, n); return -1; }
    X265B200_CHECK(cudaSetDevice(device));
    X265B200_CHECK(cudaFree(0));                                            // the primary context the green contexts derive from
    CUresult (*devGet)(CUdevice*, int) = nullptr;
    CUresult (*getRes)(CUdevice, CUdevResource*, CUdevResourceType) = nullptr;
    CUresult (*split)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int) = nullptr;
    CUresult (*genDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int) = nullptr;
    CUresult (*gcreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
    CUresult (*gdestroy)(CUgreenCtx) = nullptr;
    CUresult (*gstream)(CUstream*, CUgreenCtx, unsigned int, int) = nullptr;
    if (!drv_entry("cuDeviceGet", devGet) || !drv_entry("cuDeviceGetDevResource", getRes) || !drv_entry("cuDevSmResourceSplitByCount", split) ||
        !drv_entry("cuDevResourceGenerateDesc", genDesc) || !drv_entry("cuGreenCtxCreate", gcreate) || !drv_entry("cuGreenCtxDestroy", gdestroy) ||
        !drv_entry("cuGreenCtxStreamCreate", gstream))
    { set_error("x265b200_sm_partition: this driver has no green-context API"); return -1; }
    CUdevice dev;
    CUdevResource all, first, rest;
    unsigned int groups = 1;
    CUresult r;
#define DRV(call, what) if ((r = (call)) != CUDA_SUCCESS) { set_error("x265b200_sm_partition: %s failed (CUresult %d)", what, (int)r); return -1; }
    DRV(devGet(&dev, device), "cuDeviceGet");
    DRV(getRes(dev, &all, CU_DEV_RESOURCE_TYPE_SM), "cuDeviceGetDevResource");
    if ((unsigned)smsFirst >= all.sm.smCount) { set_error("x265b200_sm_partition: %d SMs asked of %u", smsFirst, all.sm.smCount); return -1; }
    DRV(split(&first, &groups, &all, &rest, 0, (unsigned)smsFirst), "cuDevSmResourceSplitByCount");
    if (groups != 1 || !rest.sm.smCount) { set_error("x265b200_sm_partition: the split left no second group"); return -1; }
    GreenPartition P; memset(&P, 0, sizeof(P));
    CUdevResource* parts[2] = { &first, &rest };
    for (int k = 0; k < 2; k++)
    {
        CUdevResourceDesc desc;
        DRV(genDesc(&desc, parts[k], 1), "cuDevResourceGenerateDesc");
        DRV(gcreate(&P.ctx[k], desc, dev, CU_GREEN_CTX_DEFAULT_STREAM), "cuGreenCtxCreate");
        DRV(gstream(&P.stream[k], P.ctx[k], CU_STREAM_NON_BLOCKING, 0), "cuGreenCtxStreamCreate");
    }
#undef DRV
    { std::lock_guard<std::mutex> lk(g_partMutex); g_partitions.push_back(P); }
    *streamFirst = (void*)P.stream[0]; *streamRest = (void*)P.stream[1];
    if (smsGot) { smsGot[0] = (int)first.sm.smCount; smsGot[1] = (int)rest.sm.smCount; }
    return 0;
}

void x265b200_sm_partition_release(void)
{
    CUresult (*gdestroy)(CUgreenCtx) = nullptr;
    std::lock_guard<std::mutex> lk(g_partMutex);
    if (!drv_entry("cuGreenCtxDestroy", gdestroy)) { g_partitions.clear(); return; }
    for (auto& P : g_partitions)
        for (int k = 0; k < 2; k++)
        {
            if (P.stream[k]) { cudaStreamSynchronize((cudaStream_t)P.stream[k]); cudaStreamDestroy((cudaStream_t)P.stream[k]); }
            if (P.ctx[k]) gdestroy(P.ctx[k]);
        }
    g_partitions.clear();
}

int x265b200_create(int device, void* stream, x265b200_ctx** out)
{
    if (!out) { set_error("x265b200_create: out == NULL"); return -1; }
    *out = nullptr;
    int n = x265b200_device_count();
    if (n <= 0) { set_error("x265b200_create: no CUDA device visible; this backend has no CPU fallback"); return -1; }
    if (device < 0 || device >= n) { set_error("x265b200_create: device %d out of range [0,%d)", device, n); return -1; }
    X265B200_CHECK(cudaSetDevice(device));
    x265b200_ctx* ctx = new x265b200_ctx;
    memset(&ctx->c, 0, sizeof(Ctx));
    ctx->c.device = device;
    if (stream) { ctx->c.stream = (cudaStream_t)stream; ctx->c.ownsStream = false; }
    else
    {
        if (check(cudaStreamCreateWithFlags(&ctx->c.stream, cudaStreamNonBlocking), "cudaStreamCreate")) { delete ctx; return -1; }
        ctx->c.ownsStream = true;
    }
    cudaDeviceProp prop;
    if (check(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties")) { delete ctx; return -1; }
    ctx->c.smCount = stream_sm_count(ctx->c.stream, prop.multiProcessorCount);      // a stream of an SM partition: its share
    *out = ctx;
    return 0;
}

void x265b200_destroy(x265b200_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->c.device);
    cudaStreamSynchronize(ctx->c.stream);
    for (int i = 0; i < 8; i++)
    {
        if (ctx->c.dScratch[i]) cudaFree(ctx->c.dScratch[i]);
        if (ctx->c.hPinned[i]) cudaFreeHost(ctx->c.hPinned[i]);
    }
    if (ctx->c.dMvCost) cudaFree(ctx->c.dMvCost);
    for (int i = 0; i < 4; i++)
    {
        if (ctx->c.hStage[i]) cudaFreeHost(ctx->c.hStage[i]);
        if (ctx->c.hStageEv[i]) cudaEventDestroy(ctx->c.hStageEv[i]);
    }
    if (ctx->c.copyStream)
    {
        cudaStreamSynchronize(ctx->c.copyStream); cudaStreamDestroy(ctx->c.copyStream);
        cudaEventDestroy(ctx->c.evH2D); cudaEventDestroy(ctx->c.evSearch); cudaEventDestroy(ctx->c.evD2H[0]); cudaEventDestroy(ctx->c.evD2H[1]);
    }
    me_ctu_release(&ctx->c);
    if (ctx->c.ownsStream) cudaStreamDestroy(ctx->c.stream);
    delete ctx;
}

int x265b200_sync(x265b200_ctx* ctx) { REQUIRE_CTX(ctx); X265B200_CHECK(cudaStreamSynchronize(ctx->c.stream)); return 0; }
void* x265b200_stream(x265b200_ctx* ctx) { return ctx ? (void*)ctx->c.stream : nullptr; }
uint64_t x265b200_launch_count(x265b200_ctx* ctx) { return ctx ? ctx->c.launches : 0; }

int x265b200_malloc(x265b200_ctx* ctx, size_t bytes, void** p) { REQUIRE_CTX(ctx); X265B200_CHECK(cudaMalloc(p, bytes ? bytes : 1)); return 0; }
int x265b200_free(x265b200_ctx* ctx, void* p) { REQUIRE_CTX(ctx); X265B200_CHECK(cudaStreamSynchronize(ctx->c.stream)); X265B200_CHECK(cudaFree(p)); return 0; }
int x265b200_upload(x265b200_ctx* ctx, void* d, const void* h, size_t bytes)
{
    REQUIRE_CTX(ctx);
    X265B200_CHECK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->c.stream));
    return 0;
}
int x265b200_download(x265b200_ctx* ctx, void* h, const void* d, size_t bytes)
{
    REQUIRE_CTX(ctx);
    X265B200_CHECK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, ctx->c.stream));
    X265B200_CHECK(cudaStreamSynchronize(ctx->c.stream));
    return 0;
}
int x265b200_upload2d(x265b200_ctx* ctx, void* d, size_t dpitch, const void* h, size_t spitch, size_t wbytes, size_t height)
{
    REQUIRE_CTX(ctx);
    if (!wbytes || !height) return 0;
    X265B200_CHECK(cudaMemcpy2DAsync(d, dpitch, h, spitch, wbytes, height, cudaMemcpyHostToDevice, ctx->c.stream));
    return 0;
}
int x265b200_download2d(x265b200_ctx* ctx, void* h, size_t dpitch, const void* d, size_t spitch, size_t wbytes, size_t height)
{
    REQUIRE_CTX(ctx);
    if (wbytes && height) X265B200_CHECK(cudaMemcpy2DAsync(h, dpitch, d, spitch, wbytes, height, cudaMemcpyDeviceToHost, ctx->c.stream));
    X265B200_CHECK(cudaStreamSynchronize(ctx->c.stream));
    return 0;
}
int x265b200_malloc_host(size_t bytes, void** p) { X265B200_CHECK(cudaMallocHost(p, bytes ? bytes : 1)); return 0; }
int x265b200_free_host(void* p) { X265B200_CHECK(cudaFreeHost(p)); return 0; }

// ---- block compare ----------------------------------------------------------------------------
int x265b200_pixelcmp_dev(x265b200_ctx* ctx, int kind, int depth, int w, int h,
                          const void* A, int64_t strideA, const void* B, int64_t strideB,
                          const int64_t* offA, const int64_t* offB, const int16_t* mv, int gridCols,
                          int64_t n, void* out)
{
    REQUIRE_CTX(ctx);
    return pixelcmp_dev(CTX(ctx), kind, depth, w, h, A, strideA, B, strideB, offA, offB, mv, gridCols, n, out);
}

int x265b200_pixelcmp_host(x265b200_ctx* ctx, int kind, int depth, int w, int h,
                           const void* A, size_t bytesA, int64_t strideA,
                           const void* B, size_t bytesB, int64_t strideB,
                           const int64_t* offA, const int64_t* offB, int64_t n, void* out)
{
    REQUIRE_CTX(ctx);
    if (!offA) { set_error("pixelcmp_host: offA required"); return -1; }
    Ctx* c = CTX(ctx);
    void *dA, *dB = nullptr, *dOA, *dOB = nullptr, *dOut;
    const bool wide = kind >= X265B200_CMP_SSE_PP;
    size_t outBytes = (size_t)n * (wide ? 8 : 4);
    if (stage_in(c, 0, A, bytesA, &dA)) return -1;
    if (B && kind != X265B200_CMP_SSD_S) { if (stage_in(c, 1, B, bytesB, &dB)) return -1; } else dB = dA;
    if (stage_in(c, 2, offA, (size_t)n * 8, &dOA)) return -1;
    if (offB) { if (stage_in(c, 3, offB, (size_t)n * 8, &dOB)) return -1; }
    if (scratch_dev(c, 4, outBytes ? outBytes : 1, &dOut)) return -1;
    if (pixelcmp_dev(c, kind, depth, w, h, dA, strideA, dB, strideB, (const int64_t*)dOA, (const int64_t*)dOB, nullptr, 0, n, dOut)) return -1;
    X265B200_CHECK(cudaMemcpyAsync(out, dOut, outBytes, cudaMemcpyDeviceToHost, c->stream));
    X265B200_CHECK(cudaStreamSynchronize(c->stream));
    return 0;
}

int x265b200_sad_pyramid_dev(x265b200_ctx* ctx, int depth, const void* cur, int64_t strideCur, const void* const* refsDev, int numRefs, int64_t strideRef,
                             int ctuCols, int ctuRows, const int16_t* mvCtu, int32_t* out8, int32_t* out16, int32_t* out32, int32_t* out64)
{
    REQUIRE_CTX(ctx);
    return sad_pyramid_dev(CTX(ctx), depth, cur, strideCur, refsDev, numRefs, strideRef, ctuCols, ctuRows, mvCtu, out8, out16, out32, out64);
}

int x265b200_sad_xn_dev(x265b200_ctx* ctx, int depth, int K, int w, int h, const void* fenc, int64_t fencBlockStride,
                        const void* ref, int64_t refStride, const int64_t* refOff, int64_t n, int32_t* res)
{
    REQUIRE_CTX(ctx);
    return sad_xn_dev(CTX(ctx), depth, K, w, h, fenc, fencBlockStride, ref, refStride, refOff, n, res);
}

// ---- transforms ---------------------------------------------------------------------------------
int x265b200_dct_dev(x265b200_ctx* ctx, int sizeIdx, int depth, const int16_t* src, int64_t srcBlockStride, int64_t srcStride, int16_t* dst, int64_t n)
{
    REQUIRE_CTX(ctx);
    int N = sizeIdx == 4 ? 4 : (4 << sizeIdx);
    return transform_dev(CTX(ctx), 0, sizeIdx, depth, src, srcBlockStride, srcStride, dst, (int64_t)N * N, N, n, 0, 0);
}
int x265b200_idct_dev(x265b200_ctx* ctx, int sizeIdx, int depth, const int16_t* src, int16_t* dst, int64_t dstBlockStride, int64_t dstStride, int64_t n)
{
    REQUIRE_CTX(ctx);
    int N = sizeIdx == 4 ? 4 : (4 << sizeIdx);
    return transform_dev(CTX(ctx), 1, sizeIdx, depth, src, (int64_t)N * N, N, dst, dstBlockStride, dstStride, n, 0, 0);
}
int x265b200_dct_plane_dev(x265b200_ctx* ctx, int sizeIdx, int depth, const int16_t* plane, int64_t stride, int blocksX, int blocksY, int16_t* coef)
{
    REQUIRE_CTX(ctx);
    int N = sizeIdx == 4 ? 4 : (4 << sizeIdx);
    return transform_dev(CTX(ctx), 0, sizeIdx, depth, plane, N, stride, coef, (int64_t)N * N, N, (int64_t)blocksX * blocksY, blocksX, (int64_t)N * stride);
}
int x265b200_idct_plane_dev(x265b200_ctx* ctx, int sizeIdx, int depth, const int16_t* coef, int16_t* plane, int64_t stride, int blocksX, int blocksY)
{
    REQUIRE_CTX(ctx);
    int N = sizeIdx == 4 ? 4 : (4 << sizeIdx);
    return transform_dev(CTX(ctx), 1, sizeIdx, depth, coef, (int64_t)N * N, N, plane, N, stride, (int64_t)blocksX * blocksY, blocksX, (int64_t)N * stride);
}
int x265b200_sub_ps_plane_dev(x265b200_ctx* ctx, int depth, const void* a, int64_t strideA, const void* b, int64_t strideB, int16_t* dst, int64_t dstStride, int w, int h)
{
    REQUIRE_CTX(ctx);
    return sub_ps_plane_dev(CTX(ctx), depth, a, strideA, b, strideB, dst, dstStride, w, h);
}
int x265b200_add_ps_plane_dev(x265b200_ctx* ctx, int depth, void* dst, int64_t dstStride, const void* pred, int64_t predStride, const int16_t* resi, int64_t resiStride, int w, int h)
{
    REQUIRE_CTX(ctx);
    return add_ps_plane_dev(CTX(ctx), depth, dst, dstStride, pred, predStride, resi, resiStride, w, h);
}
int x265b200_quant_dev(x265b200_ctx* ctx, const int16_t* coef, const int32_t* quantCoeff, int32_t* deltaU, int16_t* qCoef, int qBits, int add, int numCoeff, int64_t n, uint32_t* numSig)
{
    REQUIRE_CTX(ctx);
    return quant_dev(CTX(ctx), 0, coef, quantCoeff, deltaU, qCoef, qBits, add, numCoeff, n, numSig);
}
int x265b200_nquant_dev(x265b200_ctx* ctx, const int16_t* coef, const int32_t* quantCoeff, int16_t* qCoef, int qBits, int add, int numCoeff, int64_t n, uint32_t* numSig)
{
    REQUIRE_CTX(ctx);
    return quant_dev(CTX(ctx), 1, coef, quantCoeff, nullptr, qCoef, qBits, add, numCoeff, n, numSig);
}
int x265b200_dequant_normal_dev(x265b200_ctx* ctx, const int16_t* q, int16_t* coef, int num, int64_t n, int scale, int shift)
{
    REQUIRE_CTX(ctx);
    return dequant_dev(CTX(ctx), 0, q, nullptr, coef, num, n, scale, shift);
}
int x265b200_dequant_scaling_dev(x265b200_ctx* ctx, const int16_t* q, const int32_t* deq, int16_t* coef, int num, int64_t n, int per, int shift)
{
    REQUIRE_CTX(ctx);
    return dequant_dev(CTX(ctx), 1, q, deq, coef, num, n, per, shift);
}
int x265b200_count_nonzero_dev(x265b200_ctx* ctx, const int16_t* q, int numCoeff, int64_t n, int32_t* out)
{
    REQUIRE_CTX(ctx);
    return count_nonzero_dev(CTX(ctx), q, numCoeff, n, out);
}
int x265b200_tu_pipeline_dev(x265b200_ctx* ctx, int sizeIdx, int depth, int useDST, const void* fenc, int64_t fencStride,
                             const void* pred, int64_t predStride, void* recon, int64_t reconStride, int blocksX, int blocksY,
                             const int32_t* quantCoeff, int qBits, int add, const int32_t* dequantCoef, int scaleOrPer, int dqShift,
                             int16_t* coeff, uint32_t* numSig, uint64_t* sse)
{
    REQUIRE_CTX(ctx);
    return tu_pipeline_dev(CTX(ctx), sizeIdx, depth, useDST, fenc, fencStride, pred, predStride, recon, reconStride, blocksX, blocksY,
                           quantCoeff, qBits, add, dequantCoef, scaleOrPer, dqShift, coeff, numSig, sse);
}
int x265b200_dct_table(int N, int16_t* out)
{
    if (N != 4 && N != 8 && N != 16 && N != 32) { set_error("dct_table: N=%d", N); return -1; }
    host_dct_table(N, out);
    return 0;
}

// ---- interpolation / intra ----------------------------------------------------------------------
int x265b200_interp_multi_dev(x265b200_ctx* ctx, int taps, int depth, const x265b200_interp_seg* segsHost, int numSegs)
{
    REQUIRE_CTX(ctx);
    return interp_multi_dev(CTX(ctx), taps, depth, segsHost, numSegs);
}
int x265b200_interp_dev(x265b200_ctx* ctx, int kind, int taps, int depth, int w, int h, const void* src, int64_t srcStride,
                        void* dst, int64_t dstStride, const x265b200_interp_job* jobs, int64_t n, int isRowExt)
{
    REQUIRE_CTX(ctx);
    return interp_dev(CTX(ctx), kind, taps, depth, w, h, src, srcStride, dst, dstStride, jobs, n, isRowExt);
}
int x265b200_mc_dev(x265b200_ctx* ctx, int depth, const x265b200_mc_desc* desc, const x265b200_mc_job* jobs, int64_t n, int bLuma, int bChroma)
{
    REQUIRE_CTX(ctx);
    return mc_dev(CTX(ctx), depth, desc, jobs, n, bLuma, bChroma);
}
int x265b200_sao_apply_dev(x265b200_ctx* ctx, int kind, int depth, void* rec, int64_t stride, const x265b200_sao_job* jobs, int64_t n,
                           int8_t* signBuf, const int8_t* offsets, int maxWidth)
{
    REQUIRE_CTX(ctx);
    return sao_apply_dev(CTX(ctx), kind, depth, rec, stride, jobs, n, signBuf, offsets, maxWidth);
}
int x265b200_sao_stats_dev(x265b200_ctx* ctx, int kind, int depth, const int16_t* diff, const void* rec, int64_t stride,
                           const x265b200_sao_job* jobs, int64_t n, int8_t* signBuf, int32_t* stats, int32_t* count)
{
    REQUIRE_CTX(ctx);
    return sao_stats_dev(CTX(ctx), kind, depth, diff, rec, stride, jobs, n, signBuf, stats, count);
}
int x265b200_sign_dev(x265b200_ctx* ctx, int depth, int8_t* dst, const void* src1, const void* src2, int64_t n)
{
    REQUIRE_CTX(ctx);
    return sign_dev(CTX(ctx), depth, dst, src1, src2, n);
}
int x265b200_deblock_dev(x265b200_ctx* ctx, int chroma, int depth, void* pic, const x265b200_deblock_job* jobs, int64_t n)
{
    REQUIRE_CTX(ctx);
    return deblock_dev(CTX(ctx), chroma, depth, pic, jobs, n);
}
int x265b200_propagate_cost_dev(x265b200_ctx* ctx, int* dst, const uint16_t* propagateIn, const int32_t* intraCosts, const uint16_t* interCosts,
                                const int32_t* invQscales, double fpsFactor, int64_t len)
{
    REQUIRE_CTX(ctx);
    return propagate_cost_dev(CTX(ctx), dst, propagateIn, intraCosts, interCosts, invQscales, fpsFactor, len);
}
int x265b200_fix8_pack_dev(x265b200_ctx* ctx, uint16_t* dst, const double* src, int64_t count) { REQUIRE_CTX(ctx); return fix8_dev(CTX(ctx), 0, dst, src, count); }
int x265b200_fix8_unpack_dev(x265b200_ctx* ctx, double* dst, const uint16_t* src, int64_t count) { REQUIRE_CTX(ctx); return fix8_dev(CTX(ctx), 1, dst, src, count); }
int x265b200_planecopy_dev(x265b200_ctx* ctx, int mode, int depth, const void* src, int64_t srcStride, void* dst, int64_t dstStride,
                           int width, int height, int shift, int mask)
{
    REQUIRE_CTX(ctx);
    return planecopy_dev(CTX(ctx), mode, depth, src, srcStride, dst, dstStride, width, height, shift, mask);
}
int x265b200_ssim_dist_dev(x265b200_ctx* ctx, int depth, int log2TrSize, const void* fenc, int64_t fStride, const void* recon, int64_t rStride,
                           const int64_t* offF, const int64_t* offR, int64_t n, int shift, uint64_t* ssBlock, uint64_t* ac_k)
{
    REQUIRE_CTX(ctx);
    return ssim_dist_dev(CTX(ctx), depth, log2TrSize, fenc, fStride, recon, rStride, offF, offR, n, shift, ssBlock, ac_k);
}
int x265b200_norm_fact_dev(x265b200_ctx* ctx, int depth, const void* src, const int64_t* off, int64_t n, int blockSize, int shift, uint64_t* z_k)
{
    REQUIRE_CTX(ctx);
    return norm_fact_dev(CTX(ctx), depth, src, off, n, blockSize, shift, z_k);
}
int x265b200_ssim_4x4x2_dev(x265b200_ctx* ctx, int depth, const void* pix1, int64_t stride1, const void* pix2, int64_t stride2,
                            const int64_t* off1, const int64_t* off2, int64_t n, int32_t* sums)
{
    REQUIRE_CTX(ctx);
    return ssim_core_dev(CTX(ctx), depth, pix1, stride1, pix2, stride2, off1, off2, n, sums);
}
int x265b200_ssim_end4_dev(x265b200_ctx* ctx, int depth, const int32_t* sum0, const int32_t* sum1, const int32_t* widths, int64_t n, float* out)
{
    REQUIRE_CTX(ctx);
    return ssim_end4_dev(CTX(ctx), depth, sum0, sum1, widths, n, out);
}
int x265b200_plane_clip_max_dev(x265b200_ctx* ctx, int depth, void* src, int64_t stride, int width, int height, int minPix, int maxPix,
                                uint64_t* outsum, uint32_t* outmax)
{
    REQUIRE_CTX(ctx);
    return plane_clip_max_dev(CTX(ctx), depth, src, stride, width, height, minPix, maxPix, outsum, outmax);
}
int x265b200_intra_pred_dev(x265b200_ctx* ctx, int depth, int log2N, const void* nbr, void* dst, int64_t dstStride, const x265b200_intra_job* jobs, int64_t n)
{
    REQUIRE_CTX(ctx);
    return intra_pred_dev(CTX(ctx), depth, log2N, nbr, dst, dstStride, jobs, n);
}
int x265b200_intra_filter_dev(x265b200_ctx* ctx, int depth, int log2N, const void* src, void* dst, int64_t n)
{
    REQUIRE_CTX(ctx);
    return intra_filter_dev(CTX(ctx), depth, log2N, src, dst, n);
}
int x265b200_intra_allangs_dev(x265b200_ctx* ctx, int depth, int log2N, const void* refPix, const void* filtPix, void* dest, int bLuma, int64_t n)
{
    REQUIRE_CTX(ctx);
    return intra_allangs_dev(CTX(ctx), depth, log2N, refPix, filtPix, dest, bLuma, n);
}

int x265b200_intra_modes_dev(x265b200_ctx* ctx, int depth, int log2N, const void* neighbours, void* dest, int bLuma, int64_t n)
{
    REQUIRE_CTX(ctx);
    return intra_modes_dev(CTX(ctx), depth, log2N, neighbours, dest, bLuma, n);
}

// ---- glue ---------------------------------------------------------------------------------------
int x265b200_glue_dev(x265b200_ctx* ctx, int op, int depth, int w, int h, void* dst, int64_t dstStride,
                      const void* src0, int64_t src0Stride, const void* src1, int64_t src1Stride,
                      const x265b200_glue_job* jobs, int64_t n, int p0, int p1, int p2, int p3)
{
    REQUIRE_CTX(ctx);
    return glue_dev(CTX(ctx), op, depth, w, h, dst, dstStride, src0, src0Stride, src1, src1Stride, jobs, n, p0, p1, p2, p3);
}
int x265b200_var_dev(x265b200_ctx* ctx, int depth, int size, const void* src, int64_t stride, const int64_t* off, int64_t n, uint64_t* out)
{
    REQUIRE_CTX(ctx);
    return var_dev(CTX(ctx), depth, size, src, stride, off, n, out);
}
int x265b200_psy_cost_dev(x265b200_ctx* ctx, int depth, int size, const void* source, int64_t sstride, const void* recon, int64_t rstride,
                          const int64_t* offS, const int64_t* offR, int64_t n, int32_t* out)
{
    REQUIRE_CTX(ctx);
    return psycost_dev(CTX(ctx), depth, size, source, sstride, recon, rstride, offS, offR, n, out);
}
int x265b200_copy_cnt_dev(x265b200_ctx* ctx, int size, int16_t* coeff, const int16_t* residual, int64_t resiStride,
                          const int64_t* off, int64_t n, uint32_t* numSig)
{
    REQUIRE_CTX(ctx);
    return copy_cnt_dev(CTX(ctx), size, coeff, residual, resiStride, off, n, numSig);
}
int x265b200_denoise_dct_dev(x265b200_ctx* ctx, int16_t* dctCoef, uint32_t* resSum, const uint16_t* offset, int numCoeff, int64_t n)
{
    REQUIRE_CTX(ctx);
    return denoise_dct_dev(CTX(ctx), dctCoef, resSum, offset, numCoeff, n);
}
int x265b200_lowpass_dct_dev(x265b200_ctx* ctx, int sizeIdx, int depth, const int16_t* src, int64_t srcBlockStride,
                             int64_t srcStride, int16_t* dst, int64_t n)
{
    REQUIRE_CTX(ctx);
    if (sizeIdx < 1 || sizeIdx > 3) { set_error("lowpass_dct: sizeIdx %d (1..3 = 8/16/32)", sizeIdx); return -1; }
    if (n <= 0) return 0;
    const int N = 4 << sizeIdx, half = N >> 1;
    void *dAvg = nullptr, *dCoef = nullptr, *dTot = nullptr;
    if (scratch_dev(CTX(ctx), 0, (size_t)n * half * half * 2, &dAvg) || scratch_dev(CTX(ctx), 1, (size_t)n * half * half * 2, &dCoef) ||
        scratch_dev(CTX(ctx), 2, (size_t)n * 4, &dTot)) return -1;
    if (lowpass_front_dev(CTX(ctx), src, srcBlockStride, srcStride, n, N, (int16_t*)dAvg, (int32_t*)dTot)) return -1;
    // the half-size transform is the table's own standard_dct (lowpassdct.cpp:117-119)
    if (x265b200_dct_dev(ctx, sizeIdx - 1, depth, (const int16_t*)dAvg, (int64_t)half * half, half, (int16_t*)dCoef, n)) return -1;
    return lowpass_back_dev(CTX(ctx), (const int16_t*)dCoef, (const int32_t*)dTot, n, N, dst);
}

// ---- motion estimation ----------------------------------------------------------------------------
int x265b200_me_batch_dev(x265b200_ctx* ctx, int depth, const void* fencPlane, int64_t fencStride,
                          const void* refPlane, const void* const* refPlanes, int64_t refStride,
                          x265b200_me_job* jobs, int64_t n, int maxW, int maxH,
                          int searchMethod, int subpelRefine, int merange, double lambda, int maxSlices)
{
    REQUIRE_CTX(ctx);
    return me_batch_dev(CTX(ctx), depth, fencPlane, fencStride, refPlane, refPlanes, refStride, nullptr, nullptr, jobs, n, maxW, maxH,
                        searchMethod, subpelRefine, merange, lambda, maxSlices);
}
int x265b200_me_batch_sea_dev(x265b200_ctx* ctx, int depth, const void* fencPlane, int64_t fencStride,
                              const void* refPlane, const void* const* refPlanes, int64_t refStride,
                              const uint32_t* const* integralPlanes, x265b200_me_job* jobs, int64_t n, int maxW, int maxH,
                              int subpelRefine, int merange, double lambda, int maxSlices)
{
    REQUIRE_CTX(ctx);
    if (!integralPlanes) { set_error("me_batch_sea: integralPlanes is NULL"); return -1; }
    return me_batch_dev(CTX(ctx), depth, fencPlane, fencStride, refPlane, refPlanes, refStride, nullptr, integralPlanes, jobs, n, maxW, maxH,
                        4 /* X265_SEA */, subpelRefine, merange, lambda, maxSlices);
}
int x265b200_sea_integral_dev(x265b200_ctx* ctx, int depth, const void* reconOrigin, int64_t stride, int padX, int padY, int maxHeight,
                              uint32_t* const planes[12])
{
    REQUIRE_CTX(ctx);
    if (!planes) { set_error("sea_integral: planes is NULL"); return -1; }
    return sea_integral_dev(CTX(ctx), depth, reconOrigin, stride, padX, padY, maxHeight, planes);
}
int x265b200_integral_inith_dev(x265b200_ctx* ctx, int depth, int width, uint32_t* sum, const void* pix, int64_t stride)
{
    REQUIRE_CTX(ctx);
    return integral_inith_dev(CTX(ctx), depth, width, sum, pix, stride);
}
int x265b200_integral_initv_dev(x265b200_ctx* ctx, int height, uint32_t* sum, int64_t stride)
{
    REQUIRE_CTX(ctx);
    return integral_initv_dev(CTX(ctx), height, sum, stride);
}
int x265b200_ads_dev(x265b200_ctx* ctx, int kind, int lxHalf, const uint32_t* sums, int64_t delta, const uint16_t* costMvX, int width,
                     const x265b200_ads_job* jobs, int64_t n, int16_t* mvs, int32_t* counts)
{
    REQUIRE_CTX(ctx);
    return ads_dev(CTX(ctx), kind, lxHalf, sums, delta, costMvX, width, jobs, n, mvs, counts);
}
int x265b200_me_batch_chroma_dev(x265b200_ctx* ctx, int depth, const void* fencPlane, int64_t fencStride,
                                 const void* refPlane, const void* const* refPlanes, int64_t refStride,
                                 const x265b200_me_chroma* chroma, x265b200_me_job* jobs, int64_t n, int maxW, int maxH,
                                 int searchMethod, int subpelRefine, int merange, double lambda, int maxSlices)
{
    REQUIRE_CTX(ctx);
    if (!chroma) { set_error("me_batch_chroma: chroma descriptor is NULL"); return -1; }
    return me_batch_dev(CTX(ctx), depth, fencPlane, fencStride, refPlane, refPlanes, refStride, chroma, nullptr, jobs, n, maxW, maxH,
                        searchMethod, subpelRefine, merange, lambda, maxSlices);
}
int x265b200_me_frame_dev(x265b200_ctx* ctx, int depth, const void* curOrigin, int64_t curStride, const void* const* refOriginsHost, int numRefs,
                          int64_t refStride, int marginX, int marginY, int rowsTotal, int ctuCols, int ctuRows, int puMask,
                          const int32_t* mvpCtu, int searchMethod, int subpelRefine, int merange, double lambda, int32_t* out)
{
    REQUIRE_CTX(ctx);
    return me_frame_dev(CTX(ctx), depth, curOrigin, curStride, refOriginsHost, numRefs, refStride, marginX, marginY, rowsTotal,
                        ctuCols, ctuRows, puMask, mvpCtu, searchMethod, subpelRefine, merange, lambda, out);
}
int x265b200_me_frame_ex_dev(x265b200_ctx* ctx, const x265b200_me_frame_params* params, const x265b200_me_frame_planes* planes,
                             const int32_t* mvpCtu, const int32_t* mvpPu, const uint8_t* numCandPu, const int32_t* mvcPu, int32_t* out)
{
    REQUIRE_CTX(ctx);
    return me_frame_ex_dev(CTX(ctx), params, planes, mvpCtu, mvpPu, numCandPu, mvcPu, out);
}
int x265b200_sad_stream_dev(x265b200_ctx* ctx, int depth, const void* poolOrigin, int64_t framePitch, int64_t stride,
                            int marginX, int marginY, int rowsTotal, int numFrames, int ctuCols, int ctuRows,
                            const x265b200_sad_group* groupsHost, int numGroups, int numRefs,
                            int32_t* out8, int32_t* out16, int32_t* out32, int32_t* out64)
{
    REQUIRE_CTX(ctx);
    return sad_stream_dev(CTX(ctx), depth, poolOrigin, framePitch, stride, marginX, marginY, rowsTotal, numFrames, ctuCols, ctuRows,
                          groupsHost, numGroups, numRefs, out8, out16, out32, out64);
}
// host-buffer forms: H2D and D2H run on the context's copy stream, ordered against the search on the compute stream by events, so the
// copies of one call overlap whatever else the caller has queued on the compute stream (the stages that follow the previous search)
static int host_copy_setup(x265b200_ctx* ctx)
{
    Ctx& c = ctx->c;
    if (c.copyStream) return 0;
    X265B200_CHECK(cudaStreamCreateWithFlags(&c.copyStream, cudaStreamNonBlocking));
    X265B200_CHECK(cudaEventCreateWithFlags(&c.evH2D, cudaEventDisableTiming));
    X265B200_CHECK(cudaEventCreateWithFlags(&c.evSearch, cudaEventDisableTiming));
    X265B200_CHECK(cudaEventCreateWithFlags(&c.evD2H[0], cudaEventDisableTiming));
    X265B200_CHECK(cudaEventCreateWithFlags(&c.evD2H[1], cudaEventDisableTiming));
    return 0;
}
static int host_results_back(x265b200_ctx* ctx, int32_t* devOut, int32_t* hostOut, size_t outBytes)
{
    Ctx& c = ctx->c;
    X265B200_CHECK(cudaEventRecord(c.evSearch, c.stream));
    X265B200_CHECK(cudaStreamWaitEvent(c.copyStream, c.evSearch, 0));
    X265B200_CHECK(cudaMemcpyAsync(hostOut, devOut, outBytes, cudaMemcpyDeviceToHost, c.copyStream));
    X265B200_CHECK(cudaEventRecord(c.evD2H[(c.hostHead + c.hostCount) & 1], c.copyStream));
    c.hostCount++;
    return 0;
}
int x265b200_me_frame_host_begin(x265b200_ctx* ctx, int depth, const void* hostCurBase, size_t planeBytes, void* devCurBase, int64_t curStride,
                                 const void* const* refOriginsHost, int numRefs, int64_t refStride,
                                 int marginX, int marginY, int rowsTotal, int ctuCols, int ctuRows, int puMask,
                                 const int32_t* mvpCtu, int searchMethod, int subpelRefine, int merange, double lambda,
                                 int32_t* devOut, int32_t* hostOut, size_t outBytes)
{
    REQUIRE_CTX(ctx);
    if (!hostCurBase || !devCurBase || !devOut || !hostOut) { set_error("me_frame_host: null buffer"); return -1; }
    if (ctx->c.hostCount >= 2) { set_error("me_frame_host_begin: two calls are pending already (x265b200_me_frame_host_end)"); return -1; }
    if (host_copy_setup(ctx)) return -1;
    const int px = depth > 8 ? 2 : 1;
    X265B200_CHECK(cudaMemcpyAsync(devCurBase, hostCurBase, planeBytes, cudaMemcpyHostToDevice, ctx->c.copyStream));
    X265B200_CHECK(cudaEventRecord(ctx->c.evH2D, ctx->c.copyStream));
    X265B200_CHECK(cudaStreamWaitEvent(ctx->c.stream, ctx->c.evH2D, 0));
    const char* origin = (const char*)devCurBase + ((int64_t)marginY * curStride + marginX) * px;
    if (me_frame_dev(CTX(ctx), depth, origin, curStride, refOriginsHost, numRefs, refStride, marginX, marginY, rowsTotal, ctuCols, ctuRows, puMask,
                     mvpCtu, searchMethod, subpelRefine, merange, lambda, devOut)) return -1;
    return host_results_back(ctx, devOut, hostOut, outBytes);
}
int x265b200_me_frame_ex_host_begin(x265b200_ctx* ctx, const x265b200_me_frame_params* params, const x265b200_me_frame_planes* planes,
                                    const void* hostCurYBase, void* devCurYBase, size_t bytesY,
                                    const void* hostCurCbBase, void* devCurCbBase, const void* hostCurCrBase, void* devCurCrBase, size_t bytesC,
                                    const int32_t* mvpCtu, const int32_t* mvpPu, const uint8_t* numCandPu, const int32_t* mvcPu,
                                    int32_t* devOut, int32_t* hostOut, size_t outBytes)
{
    REQUIRE_CTX(ctx);
    if (!hostCurYBase || !devCurYBase || !devOut || !hostOut) { set_error("me_frame_ex_host: null buffer"); return -1; }
    if (ctx->c.hostCount >= 2) { set_error("me_frame_ex_host_begin: two calls are pending already (x265b200_me_frame_host_end)"); return -1; }
    if (host_copy_setup(ctx)) return -1;
    cudaStream_t cs = ctx->c.copyStream;
    X265B200_CHECK(cudaMemcpyAsync(devCurYBase, hostCurYBase, bytesY, cudaMemcpyHostToDevice, cs));
    if (hostCurCbBase && devCurCbBase) X265B200_CHECK(cudaMemcpyAsync(devCurCbBase, hostCurCbBase, bytesC, cudaMemcpyHostToDevice, cs));
    if (hostCurCrBase && devCurCrBase) X265B200_CHECK(cudaMemcpyAsync(devCurCrBase, hostCurCrBase, bytesC, cudaMemcpyHostToDevice, cs));
    X265B200_CHECK(cudaEventRecord(ctx->c.evH2D, cs));
    X265B200_CHECK(cudaStreamWaitEvent(ctx->c.stream, ctx->c.evH2D, 0));
    if (me_frame_ex_dev(CTX(ctx), params, planes, mvpCtu, mvpPu, numCandPu, mvcPu, devOut)) return -1;
    return host_results_back(ctx, devOut, hostOut, outBytes);
}
int x265b200_me_frame_host_end(x265b200_ctx* ctx)
{
    REQUIRE_CTX(ctx);
    if (!ctx->c.hostCount) return 0;
    const int slot = ctx->c.hostHead;
    ctx->c.hostHead ^= 1; ctx->c.hostCount--;
    X265B200_CHECK(cudaEventSynchronize(ctx->c.evD2H[slot]));                 // the OLDEST pending call
    return 0;
}
int x265b200_me_frame_host(x265b200_ctx* ctx, int depth, const void* hostCurBase, size_t planeBytes, void* devCurBase, int64_t curStride,
                           const void* const* refOriginsHost, int numRefs, int64_t refStride,
                           int marginX, int marginY, int rowsTotal, int ctuCols, int ctuRows, int puMask,
                           const int32_t* mvpCtu, int searchMethod, int subpelRefine, int merange, double lambda,
                           int32_t* devOut, int32_t* hostOut, size_t outBytes)
{
    if (x265b200_me_frame_host_begin(ctx, depth, hostCurBase, planeBytes, devCurBase, curStride, refOriginsHost, numRefs, refStride, marginX, marginY, rowsTotal,
                                     ctuCols, ctuRows, puMask, mvpCtu, searchMethod, subpelRefine, merange, lambda, devOut, hostOut, outBytes)) return -1;
    return x265b200_me_frame_host_end(ctx);
}
int x265b200_me_frame_ex_host(x265b200_ctx* ctx, const x265b200_me_frame_params* params, const x265b200_me_frame_planes* planes,
                              const void* hostCurYBase, void* devCurYBase, size_t bytesY,
                              const void* hostCurCbBase, void* devCurCbBase, const void* hostCurCrBase, void* devCurCrBase, size_t bytesC,
                              const int32_t* mvpCtu, const int32_t* mvpPu, const uint8_t* numCandPu, const int32_t* mvcPu,
                              int32_t* devOut, int32_t* hostOut, size_t outBytes)
{
    if (x265b200_me_frame_ex_host_begin(ctx, params, planes, hostCurYBase, devCurYBase, bytesY, hostCurCbBase, devCurCbBase, hostCurCrBase, devCurCrBase, bytesC,
                                        mvpCtu, mvpPu, numCandPu, mvcPu, devOut, hostOut, outBytes)) return -1;
    return x265b200_me_frame_host_end(ctx);
}

int x265b200_cutree_propagate_dev(x265b200_ctx* ctx, int widthInCU, int heightInCU, const uint16_t* propagateCostB, const int32_t* intraCost,
                                  const uint16_t* lowresCosts, const int32_t* invQscale, const int32_t* mvs0, const int32_t* mvs1,
                                  uint16_t* refCost0, uint16_t* refCost1, int bipredWeight, double fpsFactor)
{
    REQUIRE_CTX(ctx);
    return cutree_propagate_dev(CTX(ctx), widthInCU, heightInCU, propagateCostB, intraCost, lowresCosts, invQscale, mvs0, mvs1, refCost0, refCost1,
                                bipredWeight, fpsFactor);
}
int x265b200_aq_energy_dev(x265b200_ctx* ctx, int depth, int csp, int qgSize, const void* y, int64_t strideY, const void* cb, const void* cr,
                           int64_t strideC, int picWidth, int picHeight, uint32_t* energy, uint64_t* wpSumSsd)
{
    REQUIRE_CTX(ctx);
    return aq_energy_dev(CTX(ctx), depth, csp, qgSize, y, strideY, cb, cr, strideC, picWidth, picHeight, energy, wpSumSsd);
}
int x265b200_apply_weight_dev(x265b200_ctx* ctx, int depth, const void* srcOrigin, void* dstOrigin, int64_t stride, int width, int height,
                              int marginX, int marginY, int inputWeight, int inputOffset, int log2WeightDenom)
{
    REQUIRE_CTX(ctx);
    return apply_weight_dev(CTX(ctx), depth, srcOrigin, dstOrigin, stride, width, height, marginX, marginY, inputWeight, inputOffset, log2WeightDenom);
}
int x265b200_me_frame_layout(int ctuSize, int minCuSize, int rect, int amp, int32_t* outXYWH, int cap)
{
    return me_frame_layout(ctuSize, minCuSize, rect, amp, outXYWH, cap);
}
int x265b200_bitcost_table(double lambda, uint16_t* out)
{
    if (!out) { set_error("bitcost_table: out == NULL"); return -1; }
    host_bitcost_table(lambda, out);
    return 0;
}
double x265b200_lambda(int qp, int depth)
{
    double v = pow(2.0, (qp - 12) / 6.0 + (depth - 8));
    return floor(v * 10000.0 + 0.5) / 10000.0;
}

// ---- lookahead -------------------------------------------------------------------------------------
int x265b200_lowres_init_dev(x265b200_ctx* ctx, int depth, const void* src, int64_t srcStride, void* const planes[4], int64_t dstStride,
                             int width, int height, int marginX, int marginY)
{
    REQUIRE_CTX(ctx);
    return lowres_init_dev(CTX(ctx), depth, src, srcStride, planes, dstStride, width, height, marginX, marginY);
}
int x265b200_extend_border_dev(x265b200_ctx* ctx, int depth, void* origin, int64_t stride, int width, int height, int marginX, int marginY)
{
    REQUIRE_CTX(ctx);
    return extend_border_dev(CTX(ctx), depth, origin, stride, width, height, marginX, marginY);
}
int x265b200_la_intra_dev(x265b200_ctx* ctx, int depth, const void* plane0, int64_t stride, int widthInCU, int heightInCU,
                          const int32_t* invQscale, int intraPenalty, int32_t* intraCost, uint8_t* intraMode,
                          uint16_t* lowresCosts, int32_t* rowSatds, int32_t* sums)
{
    REQUIRE_CTX(ctx);
    return la_intra_dev(CTX(ctx), depth, plane0, stride, widthInCU, heightInCU, invQscale, intraPenalty, intraCost, intraMode, lowresCosts, rowSatds, sums);
}
int x265b200_la_estimate_dev(x265b200_ctx* ctx, int depth, const void* const* planes, int64_t stride, int widthInCU, int heightInCU,
                             const x265b200_la_triple* triplesHost, int numTriples, int32_t* mvPool, int32_t* mvCostPool,
                             const int32_t* const* intraCost, const int32_t* const* invQscale, uint16_t* lowresCosts,
                             int32_t* rowSatds, int32_t* sums, double lambda, int lookaheadSlices, const x265b200_la_weight* weights)
{
    REQUIRE_CTX(ctx);
    return la_estimate_dev(CTX(ctx), depth, planes, stride, widthInCU, heightInCU, triplesHost, numTriples, mvPool, mvCostPool,
                           intraCost, invQscale, lowresCosts, rowSatds, sums, lambda, 1, lookaheadSlices, nullptr, weights);
}
int x265b200_la_weight_guess(int depth, int width, int lines, uint64_t fencSum, uint64_t fencSsd, uint64_t refSum, uint64_t refSsd, int32_t out[7])
{
    if (!out || width <= 0 || lines <= 0 || depth < 8) { set_error("la_weight_guess: bad arguments"); return -1; }
    la_weight_guess(depth, width, lines, fencSum, fencSsd, refSum, refSsd, out);
    return 0;
}
int x265b200_la_weights_analyse_dev(x265b200_ctx* ctx, int depth, const x265b200_la_weight_job* jobsHost, int numJobs,
                                    int64_t stride, int paddedLines, int64_t padOffset, int width, int lines, x265b200_la_weight* out)
{
    REQUIRE_CTX(ctx);
    return la_weights_analyse_dev(CTX(ctx), depth, jobsHost, numJobs, stride, paddedLines, padOffset, width, lines, out);
}
int x265b200_la_estimate_hme_dev(x265b200_ctx* ctx, int depth, const void* const* planes, int64_t stride, int widthInCU, int heightInCU,
                                 const x265b200_la_hme* hme, const x265b200_la_triple* triplesHost, int numTriples, int32_t* mvPool, int32_t* mvCostPool,
                                 const int32_t* const* intraCost, const int32_t* const* invQscale, uint16_t* lowresCosts,
                                 int32_t* rowSatds, int32_t* sums, double lambda, int lookaheadSlices, const x265b200_la_weight* weights)
{
    REQUIRE_CTX(ctx);
    if (!hme) { set_error("la_estimate_hme: hme descriptor is NULL"); return -1; }
    return la_estimate_dev(CTX(ctx), depth, planes, stride, widthInCU, heightInCU, triplesHost, numTriples, mvPool, mvCostPool,
                           intraCost, invQscale, lowresCosts, rowSatds, sums, lambda, 1, lookaheadSlices, hme, weights);
}

} // extern "C"

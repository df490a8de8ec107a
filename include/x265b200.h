/* x265b200.h -- C ABI of the B200-native primitive backend for x265 (libx265b200.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.
 * Each entry point names the reference interface it replaces (file:line under
 * /root/reference/source).  Two flavours of entry point:
 *   *_dev  : all pointers are DEVICE pointers; asynchronous on the context's stream.  Every family has it.
 *   *_host : HOST buffers in and out; the call stages H2D, runs the same kernel and copies the results back before
 *            returning.  Provided where a caller hands over host memory per call: the block compares
 *            (x265b200_pixelcmp_host) and the frame searches (x265b200_me_frame_host / x265b200_me_frame_ex_host: the new
 *            source picture comes from the host, the reference pictures stay resident on the device like a decoded-picture
 *            buffer, the {mv, cost} results return to the host) -- the latter are what bench.py's end-to-end number goes
 *            through.  The pointer-compatible EncoderPrimitives slots (adapter/) stage their operands with
 *            x265b200_upload / x265b200_download around the *_dev entries.
 * Return value: 0 on success, negative on error (message via x265b200_last_error()).
 * There is NO CPU fallback anywhere in this library: without a CUDA device every compute
 * entry point fails loudly.
 *
 * Conventions: `depth` is the x265 build bit depth (8 => pixel = uint8_t, 10/12 => pixel =
 * uint16_t, common/common.h:126-142); strides/offsets are in elements; MVs are int32 {x,y}.
 */
#ifndef X265B200_H
#define X265B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct x265b200_ctx x265b200_ctx;

/* ---- context / memory -------------------------------------------------------------------- */
int         x265b200_version(void);
const char* x265b200_last_error(void);
int         x265b200_device_count(void);
/* stream == NULL: the context creates its own non-blocking stream. */
int         x265b200_create(int device, void* cudaStream, x265b200_ctx** out);
void        x265b200_destroy(x265b200_ctx* ctx);
/* SM partitioning (CUDA green contexts): splits the SMs of `device` into two disjoint groups and returns one stream per group --
 * streamFirst runs on at least smsFirst SMs (rounded up to the hardware's granularity, 8 on sm_100), streamRest on all the others;
 * smsGot[2] = the two group sizes.  Contexts created on these streams (x265b200_create) size their persistent grids for their
 * group.  Used to run the latency-bound lookahead (a few warps per SM, x265b200_la_*) BESIDE the frame search, which otherwise
 * fills every SM's register file and serialises with it.  The streams live until x265b200_sm_partition_release(). */
int         x265b200_sm_partition(int device, int smsFirst, void** streamFirst, void** streamRest, int smsGot[2]);
void        x265b200_sm_partition_release(void);
int         x265b200_sync(x265b200_ctx* ctx);
void*       x265b200_stream(x265b200_ctx* ctx);                 /* cudaStream_t the kernels run on */
uint64_t    x265b200_launch_count(x265b200_ctx* ctx);           /* kernels launched so far        */
int         x265b200_malloc(x265b200_ctx* ctx, size_t bytes, void** devPtr);
int         x265b200_free(x265b200_ctx* ctx, void* devPtr);
int         x265b200_upload(x265b200_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int         x265b200_download(x265b200_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
/* 2-D strided copies (pitches and width in BYTES), e.g. a w x h block out of / into a strided host plane */
int         x265b200_upload2d(x265b200_ctx* ctx, void* dst_dev, size_t dpitch, const void* src_host, size_t spitch, size_t widthBytes, size_t height);
int         x265b200_download2d(x265b200_ctx* ctx, void* dst_host, size_t dpitch, const void* src_dev, size_t spitch, size_t widthBytes, size_t height);
int         x265b200_malloc_host(size_t bytes, void** hostPtr);  /* pinned */
int         x265b200_free_host(void* hostPtr);

/* ---- block compare: replaces EncoderPrimitives pu[].sad / pu[].satd / cu[].sa8d /
 *      cu[].sse_pp / cu[].sse_ss / cu[].ssd_s  (common/primitives.h:133-137,247-302;
 *      C reference common/pixel.cpp:40-55,167-186,210-391) over n independent block pairs. */
enum {
    X265B200_CMP_SAD    = 0,   /* pixelcmp_t  sad<w,h>                       -> int32  */
    X265B200_CMP_SATD   = 1,   /* pixelcmp_t  satd (4x4 Hadamard)            -> int32  */
    X265B200_CMP_SA8D   = 2,   /* pixelcmp_t  cu[].sa8d: 8x8 Hadamard, rounded per 16x16 when w,h>=16 -> int32 */
    X265B200_CMP_SA8D8  = 3,   /* sa8d8<w,h>: rounded per 8x8 (chroma 4:2:0/4:2:2 CU aliases, pixel.cpp:1244,1323) */
    X265B200_CMP_SSE_PP = 4,   /* pixel_sse_t    sse<pixel,pixel>            -> uint64 (sse_t) */
    X265B200_CMP_SSE_SS = 5,   /* pixel_sse_ss_t sse<int16,int16>            -> uint64 */
    X265B200_CMP_SSD_S  = 6    /* pixel_ssd_s_t  sum a^2 over int16 (B unused) -> uint64 */
};
/* List mode: block i = A + offA[i] vs B + offB[i].  Grid mode (offA == NULL): block i is the
 * i-th w x h tile of a gridCols-wide tiling; `mv` (optional, int16 {x,y} per block, full-pel)
 * displaces the B block: the "cost at predictor" pass of motionEstimate (motion.cpp:771-796)
 * for every PU of a frame.  out: int32[n] for kinds 0-3, uint64[n] for kinds 4-6. */
int x265b200_pixelcmp_dev(x265b200_ctx* ctx, int kind, int depth, int w, int h,
                          const void* A, int64_t strideA, const void* B, int64_t strideB,
                          const int64_t* offA, const int64_t* offB,
                          const int16_t* mv, int gridCols, int64_t n, void* out);
/* Host form of list mode; planes are copied in [0, bytesA) / [0, bytesB). */
int x265b200_pixelcmp_host(x265b200_ctx* ctx, int kind, int depth, int w, int h,
                           const void* A, size_t bytesA, int64_t strideA,
                           const void* B, size_t bytesB, int64_t strideB,
                           const int64_t* offA, const int64_t* offB, int64_t n, void* out);

/* SAD pyramid: one streaming pass over a frame pair -> sad<8,8>, sad<16,16>, sad<32,32>, sad<64,64>
 * (pixel.cpp:40-55) of every 2Nx2N PU of every 64x64 CTU, i.e. the cost-at-predictor step of
 * motionEstimate (motion.cpp:771-784) for all PU levels at once, against numRefs references in ONE launch.
 * refsDev: DEVICE array of numRefs plane pointers.  mvCtu: optional full-pel {x,y} per [ref][CTU] applied to the
 * reference.  outN holds numRefs consecutive raster grids of NxN blocks (ctuCols*64/N per row).  8- and 16-bit samples. */
int x265b200_sad_pyramid_dev(x265b200_ctx* ctx, int depth, const void* cur, int64_t strideCur, const void* const* refsDev, int numRefs, int64_t strideRef,
                             int ctuCols, int ctuRows, const int16_t* mvCtu,
                             int32_t* out8, int32_t* out16, int32_t* out32, int32_t* out64);

/* Streaming form over a frame POOL, many (frame, references) groups per launch -- the kernel BASELINE's "ME SAD achieved HBM
 * GB/s" is measured on.  The pool is numFrames equal planes (PicYuv layout: `stride` pixels per row, rowsTotal rows, marginX /
 * marginY of padding), frame f at poolOrigin + f * framePitch pixels (poolOrigin = pixel (0,0) of frame 0) -- the shape of a
 * decoded-picture buffer.  groupsHost[g] = { source frame, reference frames[numRefs] } (indices into the pool).  Per group and
 * reference the SADs at zero displacement of every 2Nx2N PU (pu[LUMA_8x8 .. LUMA_64x64].sad, pixel.cpp:40-55; the
 * cost-at-predictor step of motionEstimate, motion.cpp:771-784, with mvp = 0) are written to outN[(g * numRefs + r)] as raster
 * grids of NxN blocks (ctuCols*64/N per row).  8-, 10- and 12-bit.  Persistent CTAs stream 16 KB tiles through a 6-stage
 * shared-memory ring with TMA; every plane byte is fetched once per use.  Base, stride, frame pitch and marginX must be
 * multiples of 16 bytes. */
typedef struct { int32_t cur; int32_t ref[8]; } x265b200_sad_group;
int x265b200_sad_stream_dev(x265b200_ctx* ctx, int depth, const void* poolOrigin, int64_t framePitch, int64_t stride,
                            int marginX, int marginY, int rowsTotal, int numFrames, int ctuCols, int ctuRows,
                            const x265b200_sad_group* groupsHost, int numGroups, int numRefs,
                            int32_t* out8, int32_t* out16, int32_t* out32, int32_t* out64);

/* sad_x3 / sad_x4: replaces pu[].sad_x3 / pu[].sad_x4 (primitives.h:139-140,248-249;
 * pixel.cpp:74-119).  Item i compares the cached 64-stride PU at fenc + i*fencBlockStride with
 * K (=3 or 4; 1..4 accepted) blocks ref + refOff[i*K+k], common refStride.  res: int32[n][K]. */
int x265b200_sad_xn_dev(x265b200_ctx* ctx, int depth, int K, int w, int h,
                        const void* fenc, int64_t fencBlockStride,
                        const void* ref, int64_t refStride, const int64_t* refOff,
                        int64_t n, int32_t* res);

/* ---- transforms: replaces cu[].dct / cu[].idct / dst4x4 / idst4x4 (primitives.h:153-154,
 *      275-276,318-319; dct.cpp:442-610).  sizeIdx 0..3 = 4/8/16/32-point DCT, 4 = 4x4 DST.
 * Forward: block i reads src + i*srcBlockStride with row pitch srcStride, writes N*N contiguous
 * coefficients at dst + i*N*N.  Inverse: reads N*N contiguous at src + i*N*N, writes
 * dst + i*dstBlockStride with row pitch dstStride.  8/16/32 run on the integer tensor cores. */
int x265b200_dct_dev(x265b200_ctx* ctx, int sizeIdx, int depth, const int16_t* src, int64_t srcBlockStride,
                     int64_t srcStride, int16_t* dst, int64_t n);
int x265b200_idct_dev(x265b200_ctx* ctx, int sizeIdx, int depth, const int16_t* src, int16_t* dst,
                      int64_t dstBlockStride, int64_t dstStride, int64_t n);
/* Plane forms: the blocksX x blocksY grid of N x N TUs of an int16 plane (row pitch `stride`), TU
 * (bx,by) <-> coefficient block by*blocksX+bx.  One launch per plane (the residual pipeline shape). */
int x265b200_dct_plane_dev(x265b200_ctx* ctx, int sizeIdx, int depth, const int16_t* plane, int64_t stride,
                           int blocksX, int blocksY, int16_t* coef);
int x265b200_idct_plane_dev(x265b200_ctx* ctx, int sizeIdx, int depth, const int16_t* coef, int16_t* plane,
                            int64_t stride, int blocksX, int blocksY);
/* cu[].sub_ps / cu[].add_ps (primitives.h:189-190,282-283; pixel.cpp:814-840) over a whole w x h plane:
 * dst = a - b (int16);  dst = clip(pred + resi). */
int x265b200_sub_ps_plane_dev(x265b200_ctx* ctx, int depth, const void* a, int64_t strideA, const void* b, int64_t strideB,
                              int16_t* dst, int64_t dstStride, int w, int h);
int x265b200_add_ps_plane_dev(x265b200_ctx* ctx, int depth, void* dst, int64_t dstStride, const void* pred, int64_t predStride,
                              const int16_t* resi, int64_t resiStride, int w, int h);
/* quant / nquant (primitives.h:159-160,321-322; dct.cpp:664-713) over n TUs of numCoeff
 * coefficients sharing one quantCoeff[numCoeff] table.  deltaU may be NULL; numSig: uint32[n]. */
int x265b200_quant_dev(x265b200_ctx* ctx, const int16_t* coef, const int32_t* quantCoeff, int32_t* deltaU,
                       int16_t* qCoef, int qBits, int add, int numCoeff, int64_t n, uint32_t* numSig);
int x265b200_nquant_dev(x265b200_ctx* ctx, const int16_t* coef, const int32_t* quantCoeff, int16_t* qCoef,
                        int qBits, int add, int numCoeff, int64_t n, uint32_t* numSig);
/* dequant_normal / dequant_scaling (primitives.h:161-162,323-324; dct.cpp:612-662), n TUs of num coeffs */
int x265b200_dequant_normal_dev(x265b200_ctx* ctx, const int16_t* quantCoef, int16_t* coef, int num, int64_t n,
                                int scale, int shift);
int x265b200_dequant_scaling_dev(x265b200_ctx* ctx, const int16_t* quantCoef, const int32_t* deQuantCoef,
                                 int16_t* coef, int num, int64_t n, int mcqp_miper, int shift);
/* cu[].count_nonzero (primitives.h:163; dct.cpp:714-726): out int32[n] */
int x265b200_count_nonzero_dev(x265b200_ctx* ctx, const int16_t* quantCoeff, int numCoeff, int64_t n, int32_t* out);
/* Fused residual pipeline: what Search::residualTransformQuantInter / codeIntraLumaQT do per TU through
 * Quant::transformNxN (common/quant.cpp:397-480) and Quant::invtransformNxN (quant.cpp:543-605) with rdoqLevel 0, no sign
 * hiding, no noise reduction, no transform skip / bypass -- for the blocksX x blocksY grid of N x N TUs of a plane, one launch:
 *   resi = fenc - pred (cu[].sub_ps) -> cu[].dct (dst4x4 when useDST: 4x4 luma intra) -> primitives.quant(qBits, add,
 *   quantCoeff[N*N]) -> levels coeff[tu][N*N] + numSig[tu] -> dequant_normal(scale = scaleOrPer, dqShift) or, when
 *   dequantCoef != NULL, dequant_scaling(dequantCoef[N*N], per = scaleOrPer, dqShift) -> residual' = 0 when numSig == 0
 *   (search.cpp: blockfill_s 0), the DC-only fill of quant.cpp:585-595 when numSig == 1 && level[0] != 0 && !useDST, else
 *   cu[].idct / idst4x4 -> recon = clip(pred + residual') (cu[].add_ps) -> sse[tu] = cu[].sse_pp(fenc, recon).
 * TU (bx, by) covers pixels (bx*N, by*N) of all three planes and is entry by*blocksX + bx of coeff / numSig / sse.
 * sizeIdx 0..3 = 4/8/16/32.  8/16/32 run both transforms on the integer tensor cores inside the same kernel. */
int x265b200_tu_pipeline_dev(x265b200_ctx* ctx, int sizeIdx, int depth, int useDST, const void* fenc, int64_t fencStride,
                             const void* pred, int64_t predStride, void* recon, int64_t reconStride, int blocksX, int blocksY,
                             const int32_t* quantCoeff, int qBits, int add, const int32_t* dequantCoef, int scaleOrPer, int dqShift,
                             int16_t* coeff, uint32_t* numSig, uint64_t* sse);
/* The N x N HEVC core-transform matrix the kernels use (host side, no GPU needed); N = 4,8,16,32. */
int x265b200_dct_table(int N, int16_t* out);

/* ---- glue: the element-wise entries of the table (pixel.cpp:393-557, :759-862).
 * Job i works on the w x h block at dst + dstOff / src0 + src0Off / src1 + src1Off (offsets and strides in elements of
 * the operand's own type: pixel or int16).  Operand types per op (P = pixel, S = int16):
 *   COPY_PP P<-P  COPY_SS S<-S  COPY_SP P<-S  COPY_PS S<-P   cu[]/pu[].copy_* (blockcopy_*_c)
 *   FILL_S  S<-p0                                             cu[].blockfill_s
 *   CPY2DTO1D_SHL/SHR  S<-S, dst contiguous (dstStride = w), p0 = shift      cu[].cpy2Dto1D_shl/shr
 *   CPY1DTO2D_SHL/SHR  S<-S, src contiguous (src0Stride = w), p0 = shift     cu[].cpy1Dto2D_shl/shr
 *   SUB_PS  S<-P-P (also calcresidual)   ADD_PS P<-clip(P+S)                 cu[].sub_ps / add_ps / calcresidual
 *   ADDAVG  P<-clip((S+S+offset)>>shift)  PIXELAVG_PP P<-(P+P+1)>>1          pu[].addAvg / pixelavg_pp
 *   TRANSPOSE P<-P, dst contiguous w x w                                     cu[].transpose
 *   WEIGHT_PP P<-P, WEIGHT_SP P<-S: p0 = w0, p1 = round, p2 = shift, p3 = offset   weight_pp / weight_sp
 *   SCALE1D_128TO64 P<-P: two 128-pixel lines at src0 (+0, +128) -> two 64-pixel lines at dst (+0, +64)  scale1D_128to64 (pixel.cpp:559)
 *   SCALE2D_64TO32  P<-P: 64x64 block (src0Stride) -> contiguous 32x32                                   scale2D_64to32 (pixel.cpp:585)
 *   (w, h are ignored for the two scale ops) */
enum { X265B200_GL_COPY_PP = 0, X265B200_GL_COPY_SS, X265B200_GL_COPY_SP, X265B200_GL_COPY_PS, X265B200_GL_FILL_S,
       X265B200_GL_CPY2DTO1D_SHL, X265B200_GL_CPY2DTO1D_SHR, X265B200_GL_CPY1DTO2D_SHL, X265B200_GL_CPY1DTO2D_SHR,
       X265B200_GL_SUB_PS, X265B200_GL_ADD_PS, X265B200_GL_ADDAVG, X265B200_GL_PIXELAVG_PP, X265B200_GL_TRANSPOSE,
       X265B200_GL_WEIGHT_PP, X265B200_GL_WEIGHT_SP, X265B200_GL_SCALE1D_128TO64, X265B200_GL_SCALE2D_64TO32 };
typedef struct { int64_t dstOff, src0Off, src1Off; } x265b200_glue_job;
int x265b200_glue_dev(x265b200_ctx* ctx, int op, int depth, int w, int h, void* dst, int64_t dstStride,
                      const void* src0, int64_t src0Stride, const void* src1, int64_t src1Stride,
                      const x265b200_glue_job* jobs, int64_t n, int p0, int p1, int p2, int p3);
/* cu[].var (pixel_var, pixel.cpp:703-720): out[i] = sum | (uint64)sqr << 32 of the size x size block at src + off[i] */
int x265b200_var_dev(x265b200_ctx* ctx, int depth, int size, const void* src, int64_t stride, const int64_t* off, int64_t n, uint64_t* out);
/* cu[].psy_cost_pp (psyCost_pp, pixel.cpp:726-757), size = 4, 8, 16, 32 or 64 */
int x265b200_psy_cost_dev(x265b200_ctx* ctx, int depth, int size, const void* source, int64_t sstride, const void* recon, int64_t rstride,
                          const int64_t* offS, const int64_t* offR, int64_t n, int32_t* out);
/* cu[].copy_cnt (copy_count, dct.cpp:728-743): coeff + i*size*size (contiguous) <- residual + off[i]; numSig[i] = non-zeros */
int x265b200_copy_cnt_dev(x265b200_ctx* ctx, int size, int16_t* coeff, const int16_t* residual, int64_t resiStride,
                          const int64_t* off, int64_t n, uint32_t* numSig);
/* denoiseDct (dct.cpp:745-755) over n TUs of numCoeff coefficients (contiguous) that share resSum[numCoeff] (accumulated
 * into, as NoiseReduction does per TU size/type) and offset[numCoeff] */
int x265b200_denoise_dct_dev(x265b200_ctx* ctx, int16_t* dctCoef, uint32_t* resSum, const uint16_t* offset, int numCoeff, int64_t n);
/* cu[].lowpass_dct (lowPassDct8/16/32_c, lowpassdct.cpp:33-111), sizeIdx 1..3 = 8/16/32: 2x2 average -> half-size
 * standard DCT -> zero-padded N x N with the DC replaced by the scaled block sum.  Same addressing as x265b200_dct_dev. */
int x265b200_lowpass_dct_dev(x265b200_ctx* ctx, int sizeIdx, int depth, const int16_t* src, int64_t srcBlockStride,
                             int64_t srcStride, int16_t* dst, int64_t n);

/* ---- interpolation: replaces pu[].luma_hpp/hps/vpp/vps/vsp/vss/hvpp/convert_p2s and
 *      chroma[].pu[].filter_* / p2s (primitives.h:176-182,253-263,398-407; ipfilter.cpp:40-370).
 * taps = 8 (luma) or 4 (chroma).  Job i filters the w x h block at src + srcOff into dst + dstOff
 * with coeffIdx = idxX (idxY: vertical fraction of HVPP).  srcOff points at the block itself; the
 * kernel applies the reference's -(N/2-1) tap offset (and the HPS isRowExt row offset) itself. */
enum { X265B200_IP_HPP = 0, X265B200_IP_HPS = 1, X265B200_IP_VPP = 2, X265B200_IP_VPS = 3,
       X265B200_IP_VSP = 4, X265B200_IP_VSS = 5, X265B200_IP_HVPP = 6, X265B200_IP_P2S = 7 };
typedef struct { int64_t srcOff, dstOff; int32_t idxX, idxY; } x265b200_interp_job;
int x265b200_interp_dev(x265b200_ctx* ctx, int kind, int taps, int depth, int w, int h,
                        const void* src, int64_t srcStride, void* dst, int64_t dstStride,
                        const x265b200_interp_job* jobs, int64_t n, int isRowExt);
/* The same for several (kind, block size) segments at once -- e.g. the luma predictions of every PU level of a frame: segments the
 * packed-word kernel takes (8-bit, 8 taps, HPP / VPP / HVPP, sizes and strides multiples of 4) share launches (up to 16 per launch),
 * anything else is forwarded to x265b200_interp_dev segment by segment.  segsHost is a HOST array. */
typedef struct { int32_t kind, w, h, isRowExt; const void* src; int64_t srcStride; void* dst; int64_t dstStride;
                 const x265b200_interp_job* jobs; int64_t n; } x265b200_interp_seg;
int x265b200_interp_multi_dev(x265b200_ctx* ctx, int taps, int depth, const x265b200_interp_seg* segsHost, int numSegs);

/* ---- motion compensation driver: Predict::motionCompensation (common/predict.cpp:77-257) for n PUs.
 * Per job: CUData::clipMv on both MVs (cudata.cpp:1915-1928, from cuX/cuY and the descriptor's picture size / maxCUSize),
 * then exactly the reference's choice of path -- P slice: list 0 only, weighted (predInter*Short + addWeightUni) when
 * weightedPred && weight[0][ref][0].present, else predInter*Pixel; B slice: both lists -> addWeightBi when weightedBiPred and
 * either luma weight is present, else addAvg; one list -> weighted uni or pixel path.  Luma uses the 8-tap filters at
 * quarter-pel fractions, chroma the 4-tap filters at eighth-pel fractions of mv << (1 - shift) (predict.cpp:312-318).
 * The prediction of PU (puX, puY, w, h) is written at that position of predY / predCb / predCr (chroma at the shifted
 * position and size).  refs: DEVICE array [2][maxRefs][3] of plane ORIGINS ({Y, Cb, Cr} of reference r of list l);
 * weights: DEVICE array [2][maxRefs][3] {inputWeight, inputOffset, log2WeightDenom, wtPresent} or NULL. */
typedef struct { int32_t puX, puY, w, h, cuX, cuY; int32_t refIdx[2]; int32_t mv[2][2]; } x265b200_mc_job;
typedef struct { int32_t w, o, shift, present; } x265b200_mc_weight;
typedef struct {
    int32_t csp;                                  /* x265.h:588-592: 0 = 4:0:0 ... 3 = 4:4:4 */
    int32_t isPSlice, weightedPred, weightedBiPred;
    int32_t picWidth, picHeight, maxCUSize, maxRefs;
    const void* const* refs; int64_t refStrideY, refStrideC;
    void* predY; void* predCb; void* predCr; int64_t predStrideY, predStrideC;
    const x265b200_mc_weight* weights;
} x265b200_mc_desc;
int x265b200_mc_dev(x265b200_ctx* ctx, int depth, const x265b200_mc_desc* desc, const x265b200_mc_job* jobs, int64_t n, int bLuma, int bChroma);

/* ---- in-loop filter entries: SAO + deblocking line filters ----------------------------------------------
 * One job = one call of the reference primitive on a CTU-sized block.  Offsets are in elements of the operand they index:
 * recOff into `rec` (pixels), diffOff into `diff` (int16, row pitch 64 = MAX_CU_SIZE), buf0/buf1 into signBuf (int8),
 * offsetOff into the offset tables (apply) or into stats/count (statistics, which are ADDED to, as the C code does).
 *   apply (x265b200_sao_apply_dev), replaces primitives.saoCuOrg* (primitives.h:343-349; loopfilter.cpp:45-139):
 *     E0       processSaoCUE0(rec, offsetEo, width, signLeft = buf0[2], stride)              2 rows
 *     E1       processSaoCUE1(rec, upBuff1 = buf0, offsetEo, stride, width)                  1 row
 *     E1_2ROWS processSaoCUE1_2Rows(...)                                                      2 rows
 *     E2       processSaoCUE2(rec, bufft = buf0, buff1 = buf1, offsetEo, width, stride)
 *     E3       processSaoCUE3(rec, upBuff1 = buf0, offsetEo, stride, startX, endX = width)
 *     B0       processSaoCUB0(rec, offsetBo[32], ctuWidth = width, ctuHeight = height, stride)
 *     maxWidth = largest job.width (<= 256) of the launch; it sizes the CTAs (one thread per column), so the edge-offset kinds do NOT
 *     process columns at or beyond it -- a job wider than maxWidth is the caller's error (the jobs live in device memory, the host cannot
 *     check them; CTUs are at most 64 wide, the adapter passes the slot's own width).
 *   statistics (x265b200_sao_stats_dev), replaces primitives.saoCuStats* (primitives.h:351-355; sao.cpp:1762-1926):
 *     BO / E0 / E1 (upBuff1 = buf0) / E2 (upBuff1 = buf0, upBufft = buf1) / E3 (upBuff1 = buf0), endX = width, endY = height;
 *     the sign buffers end in the state the row-by-row C loops leave them in. */
enum { X265B200_SAO_E0 = 0, X265B200_SAO_E1, X265B200_SAO_E1_2ROWS, X265B200_SAO_E2, X265B200_SAO_E3, X265B200_SAO_B0, X265B200_SAO_BO = X265B200_SAO_B0 };
typedef struct { int64_t recOff, diffOff, buf0, buf1, offsetOff; int32_t width, height, startX, pad; } x265b200_sao_job;
int x265b200_sao_apply_dev(x265b200_ctx* ctx, int kind, int depth, void* rec, int64_t stride, const x265b200_sao_job* jobs, int64_t n,
                           int8_t* signBuf, const int8_t* offsets, int maxWidth);
int x265b200_sao_stats_dev(x265b200_ctx* ctx, int kind, int depth, const int16_t* diff, const void* rec, int64_t stride,
                           const x265b200_sao_job* jobs, int64_t n, int8_t* signBuf, int32_t* stats, int32_t* count);
/* primitives.sign = calSign (primitives.h:357; loopfilter.cpp:39-43): dst[x] = sign(src1[x] - src2[x]) */
int x265b200_sign_dev(x265b200_ctx* ctx, int depth, int8_t* dst, const void* src1, const void* src2, int64_t n);
/* pelFilterLumaStrong[2] / pelFilterChroma[2] (primitives.h:372-373; loopfilter.cpp:141-180): job i filters the 4 lines
 * (UNIT_SIZE) src + k*srcStep, k = 0..3, across the edge at `offset` spacing.  luma: tcP, tcQ; chroma (chroma != 0): tcP = tc,
 * tcQ = maskP, maskQ = maskQ. */
typedef struct { int64_t srcOff, srcStep, offset; int32_t tcP, tcQ, maskQ, pad; } x265b200_deblock_job;
int x265b200_deblock_dev(x265b200_ctx* ctx, int chroma, int depth, void* pic, const x265b200_deblock_job* jobs, int64_t n);

/* ---- cuTree / ingest / ssim-rd entries (SURVEY.md 8f-4 and the rest of the pointer table) -------------------------------
 * propagateCost = estimateCUPropagateCost (primitives.h:360; pixel.cpp:914-940), fix8Pack / fix8Unpack (pixel.cpp:943-956):
 * IEEE double arithmetic in the reference's operation order (no FMA contraction), x86 double -> int conversion semantics. */
int x265b200_propagate_cost_dev(x265b200_ctx* ctx, int* dst, const uint16_t* propagateIn, const int32_t* intraCosts, const uint16_t* interCosts,
                                const int32_t* invQscales, double fpsFactor, int64_t len);
int x265b200_fix8_pack_dev(x265b200_ctx* ctx, uint16_t* dst, const double* src, int64_t count);
int x265b200_fix8_unpack_dev(x265b200_ctx* ctx, double* dst, const uint16_t* src, int64_t count);
/* planecopy_cp (uint8 source, << shift), planecopy_sp ((uint16 >> shift) & mask), planecopy_sp_shl ((uint16 << shift) & mask),
 * planecopy_pp_shr (pixel >> shift) (primitives.h:332-335; pixel.cpp:864-910); strides in elements of each side's own type */
enum { X265B200_PC_CP = 0, X265B200_PC_SP, X265B200_PC_SP_SHL, X265B200_PC_PP_SHR };
int x265b200_planecopy_dev(x265b200_ctx* ctx, int mode, int depth, const void* src, int64_t srcStride, void* dst, int64_t dstStride,
                           int width, int height, int shift, int mask);
/* cu[].ssimDist (ssimDist_c<log2TrSize>, pixel.cpp:958-981) and cu[].normFact (normFact_c, :983-994) over n blocks */
int x265b200_ssim_dist_dev(x265b200_ctx* ctx, int depth, int log2TrSize, const void* fenc, int64_t fStride, const void* recon, int64_t rStride,
                           const int64_t* offF, const int64_t* offR, int64_t n, int shift, uint64_t* ssBlock, uint64_t* ac_k);
int x265b200_norm_fact_dev(x265b200_ctx* ctx, int depth, const void* src, const int64_t* off, int64_t n, int blockSize, int shift, uint64_t* z_k);

/* ssim_4x4x2_core (primitives.h:345; pixel.cpp:631-658): sums[i][2][4] of the two adjacent 4x4 blocks at pix1 + off1[i] / pix2 + off2[i];
 * ssim_end_4 (pixel.cpp:660-702): out[i] = sum over widths[i] (1..4) window positions of ssim_end_1 on sum0[i][5][4] / sum1[i][5][4];
 * planeClipAndMax (pixel.cpp:996-1016): clip the plane to [minPix, maxPix] in place, *outsum = sum, *outmax = maximum. */
int x265b200_ssim_4x4x2_dev(x265b200_ctx* ctx, int depth, const void* pix1, int64_t stride1, const void* pix2, int64_t stride2,
                            const int64_t* off1, const int64_t* off2, int64_t n, int32_t* sums);
int x265b200_ssim_end4_dev(x265b200_ctx* ctx, int depth, const int32_t* sum0, const int32_t* sum1, const int32_t* widths, int64_t n, float* out);
int x265b200_plane_clip_max_dev(x265b200_ctx* ctx, int depth, void* src, int64_t stride, int width, int height, int minPix, int maxPix,
                                uint64_t* outsum, uint32_t* outmax);

/* ---- intra prediction: replaces cu[].intra_pred[35] / intra_filter / intra_pred_allangs
 *      (primitives.h:143-145,304-306; intrapred.cpp:31-234).  Neighbour arrays use the reference
 *      layout [topLeft, top 2N, left 2N] (4N+1 pixels).  log2N = 2..5. */
typedef struct { int64_t srcOff, dstOff; int32_t mode, bFilter; } x265b200_intra_job;
int x265b200_intra_pred_dev(x265b200_ctx* ctx, int depth, int log2N, const void* neighbours,
                            void* dst, int64_t dstStride, const x265b200_intra_job* jobs, int64_t n);
/* n arrays of 4N+1 pixels, array i at i*(4N+1) in both src and dst */
int x265b200_intra_filter_dev(x265b200_ctx* ctx, int depth, int log2N, const void* src, void* dst, int64_t n);
/* block i: refPix/filtPix arrays at i*(4N+1), 33 predictions (modes 2..34) of N*N at dest + i*33*N*N */
int x265b200_intra_allangs_dev(x265b200_ctx* ctx, int depth, int log2N, const void* refPix, const void* filtPix,
                               void* dest, int bLuma, int64_t n);

/* The prediction half of the intra mode search as the asm table drives it (Search::estIntraPredQT, encoder/search.cpp:1358-1400):
 * per block, cu[].intra_filter on the neighbour array, cu[].intra_pred[DC_IDX] (raw neighbours, edge filter = bLuma),
 * cu[].intra_pred[PLANAR_IDX] (smoothed neighbours for N = 8/16/32, raw for N = 4) and cu[].intra_pred_allangs(raw, smoothed,
 * bLuma) -- one launch, only the raw neighbour arrays are read.  dest: block i holds 35 predictions of N*N at
 * dest + i*35*N*N: [0] planar, [1] DC, [2..34] the all-angles layout (modes < 18 transposed, intrapred.cpp:206-234).
 * bLuma = (N <= 16) in the reference's calls. */
int x265b200_intra_modes_dev(x265b200_ctx* ctx, int depth, int log2N, const void* neighbours, void* dest, int bLuma, int64_t n);

/* ---- motion estimation: replaces MotionEstimate::setSourcePU + MotionEstimate::motionEstimate
 *      (encoder/motion.h:81-98; encoder/motion.cpp:167-191, :739-1569) for n independent PU
 *      searches.  One job = one call of motionEstimate(); per-job semantics are identical:
 *      mvmin/mvmax are full-pel inclusive bounds, mvp and mvc[] are quarter-pel, the result
 *      (outMv, outCost) is `outQMv` and the return value.  searchMethod uses the X265_*_SEARCH
 *      numbering of x265.h:492-497 (0 DIA, 1 HEX, 2 UMH, 3 STAR, 5 FULL; 4 SEA goes through x265b200_me_batch_sea_dev);
 *      searchMethod 6 (X265B200_ME_REFINE) runs MotionEstimate::refineMV instead (motion.cpp:606-737:
 *      predictor, one square refine, sub-pel with the fixed workload[5]; mvc/merange/subpelRefine/maxSlices unused;
 *      the reference returns only the MV, outCost is the final SATD + mvcost).
 *      lambda = x265_lambda_tab[qp] of BitCost::setQP (bitcost.cpp:31-60).  x265b200_me_batch_dev is the luma-only
 *      form (bChromaSATD = false: subme <= 2, or the lookahead-style setSourcePU of motion.cpp:167);
 *      x265b200_me_batch_chroma_dev below is the encode-style form with the chroma residual term. */
#define X265B200_ME_REFINE 6
typedef struct {
    int32_t puX, puY;                          /* PU position in the fenc/ref planes (pixels)   */
    int32_t w, h;                              /* PU size (any of the 24 inter LumaPU shapes)   */
    int32_t mvminX, mvminY, mvmaxX, mvmaxY;
    int32_t mvpX, mvpY;
    int32_t numCand;                           /* 0..8                                          */
    int32_t mvc[8][2];
    int32_t refIdx;                            /* index into refPlanes[] (0 when single plane)  */
    int32_t outMvX, outMvY, outCost;           /* OUT                                           */
} x265b200_me_job;
/* refPlanes: optional device array of plane base pointers selected by job.refIdx; when NULL every
 * job searches refPlane.  All planes share refStride; (puX,puY) addresses the same pixel in the
 * fenc plane and in every reference plane.  maxW/maxH: largest PU in the batch (sizes smem). */
int x265b200_me_batch_dev(x265b200_ctx* ctx, int depth, const void* fencPlane, int64_t fencStride,
                          const void* refPlane, const void* const* refPlanes, int64_t refStride,
                          x265b200_me_job* jobs, int64_t n, int maxW, int maxH,
                          int searchMethod, int subpelRefine, int merange, double lambda, int maxSlices);
/* Encode-style form: the Yuv-based MotionEstimate::setSourcePU (motion.cpp:193-222) used by Search::predInterSearch.
 * With subpelRefine > 2, csp != 4:0:0 and a chroma block that is a multiple of 4x4 (i.e. chroma[csp].pu[part].satd
 * exists, pixel.cpp:1200-1290) bChromaSATD is set and EVERY subpelCompare adds the SATD of the interpolated Cb and Cr
 * blocks (4-tap filters at eighth-pel fractions, motion.cpp:1601-1661).  csp uses x265.h:588-592 (1 = 4:2:0,
 * 2 = 4:2:2, 3 = 4:4:4).  Chroma planes are addressed like the luma ones: pixel (puX >> hshift, puY >> vshift);
 * ref{Cb,Cr}Planes are optional device arrays indexed by job.refIdx (else refCb / refCr). */
typedef struct {
    int32_t csp;
    const void* fencCb; const void* fencCr; int64_t fencStrideC;
    const void* refCb; const void* refCr;
    const void* const* refCbPlanes; const void* const* refCrPlanes;
    int64_t refStrideC;
} x265b200_me_chroma;
int x265b200_me_batch_chroma_dev(x265b200_ctx* ctx, int depth, const void* fencPlane, int64_t fencStride,
                                 const void* refPlane, const void* const* refPlanes, int64_t refStride,
                                 const x265b200_me_chroma* chroma, x265b200_me_job* jobs, int64_t n, int maxW, int maxH,
                                 int searchMethod, int subpelRefine, int merange, double lambda, int maxSlices);
/* --me sea (X265_SEA, motion.cpp:1242-1395): the same entry with MotionEstimate::integral[] supplied.
 * integralPlanes: DEVICE array [numRefs][12] (numRefs = 1 when refPlanes is NULL) of uint32 plane pointers in the order of
 * FrameData::m_meIntegral (framedata.h:171: 32x32, 32x24, 32x8, 24x32, 16x16, 16x12, 16x4, 12x16, 8x32, 8x8, 4x16, 4x4);
 * plane element (puX + puY*refStride) holds the box sum whose top-left pixel is (puX, puY) of the matching reference
 * plane (search.cpp:2264).  Where the reference's ads/sad_x4 read fencPUYuv pixels outside the PU (8x4, 4x8, 32x8, 8x32:
 * stale data of an earlier PU), this backend reads 0. */
int x265b200_me_batch_sea_dev(x265b200_ctx* ctx, int depth, const void* fencPlane, int64_t fencStride,
                              const void* refPlane, const void* const* refPlanes, int64_t refStride,
                              const uint32_t* const* integralPlanes, x265b200_me_job* jobs, int64_t n, int maxW, int maxH,
                              int subpelRefine, int merange, double lambda, int maxSlices);
/* FrameFilter::computeMEIntegral (framefilter.cpp:722-825) for a whole reconstructed frame: the 12 integral planes of a
 * PicYuv luma plane (origin pointer, `stride`, margins padX = maxCU+32 / padY = maxCU+16, maxHeight = numCuInHeight*maxCU).
 * planes[k] = ORIGIN (element matching pixel 0,0) of plane k, same geometry as the pixel plane.  Elements the reference
 * finalises -- box origins in columns [-padX, stride-padX-w) and rows [1-padY, maxHeight+padY-1-h] -- hold the w x h box
 * sum; the first row and the trailing rows/columns (running sums / never written in the reference) are written as 0. */
int x265b200_sea_integral_dev(x265b200_ctx* ctx, int depth, const void* reconOrigin, int64_t stride, int padX, int padY, int maxHeight,
                              uint32_t* const planes[12]);
/* integral_inith[INTEGRAL_w] / integral_initv[INTEGRAL_h] (primitives.h:368-369; framefilter.cpp:39-140), one row:
 * inith: sum[x] = pix[x] + .. + pix[x+width-1] + sum[x - stride] for x < stride - width;
 * initv: sum[x] = sum[x + height*stride] - sum[x] for x < stride.  width/height in {4, 8, 12, 16, 24, 32}. */
int x265b200_integral_inith_dev(x265b200_ctx* ctx, int depth, int width, uint32_t* sum, const void* pix, int64_t stride);
int x265b200_integral_initv_dev(x265b200_ctx* ctx, int height, uint32_t* sum, int64_t stride);
/* pu[].ads = ads_x1 / ads_x2 / ads_x4<lx,ly> (primitives.h:138,250; pixel.cpp:121-165) for n rows: job i tests `width`
 * consecutive x offsets of sums + sumsOff against encDC / thresh (kind = 1, 2 or 4; lxHalf = lx >> 1; costMvX[width] is
 * shared) and writes the surviving offsets, in order, to mvs + i*width and their number to counts[i]. */
typedef struct { int64_t sumsOff; int32_t thresh; int32_t encDC[4]; } x265b200_ads_job;
int x265b200_ads_dev(x265b200_ctx* ctx, int kind, int lxHalf, const uint32_t* sums, int64_t delta, const uint16_t* costMvX, int width,
                     const x265b200_ads_job* jobs, int64_t n, int16_t* mvs, int32_t* counts);
/* Frame form of the same search: every 2Nx2N PU (levels selected by puMask: bit0 64x64, bit1 32x32, bit2 16x16,
 * bit3 8x8) of every 64x64 CTU against numRefs reference planes.  One CTA per (CTU, reference) stages the
 * source CTU and the search window in shared memory with TMA, then runs the PU searches from there.
 * Per-PU semantics = motionEstimate(mvmin = (mvp>>2) - merange, mvmax = (mvp>>2) + merange, qmvp = mvp, no mvc)
 * with the CTU's mvp (mvpCtu[ref][ctu][2], quarter-pel, NULL = 0).  Planes are given by their ORIGIN (pixel 0,0);
 * marginX/marginY pixels around the picture and rowsTotal rows of `stride` pixels must be allocated (the
 * reference's PicYuv layout).  refOriginsHost is a HOST array of device pointers.
 * out: int32 {mvx, mvy, cost} per PU, ordered [ref][level 64,32,16,8 (only those in puMask)][raster grid of PUs]. */
int x265b200_me_frame_dev(x265b200_ctx* ctx, int depth, const void* curOrigin, int64_t curStride,
                          const void* const* refOriginsHost, int numRefs, int64_t refStride,
                          int marginX, int marginY, int rowsTotal, int ctuCols, int ctuRows, int puMask,
                          const int32_t* mvpCtu, int searchMethod, int subpelRefine, int merange, double lambda, int32_t* out);
/* General frame form: what Search::predInterSearch asks of MotionEstimate for every PU of every CTU of a frame
 * (search.cpp:2181-2436 -> setSearchRange :2724-2769 -> motionEstimate motion.cpp:739-1569 with the encode-style setSourcePU
 * :193-222), for the partition sets, CTU sizes, bit depths and ranges of all presets (param.cpp:396-555):
 *   ctuSize 64 / 32 / 16 with CUs down to minCuSize; 2Nx2N always, 2NxN / Nx2N with `rect`, the four AMP modes with `amp`
 *   (CUs >= 16); per-PU predictor mvp and up to maxCand MV candidates mvc[] (AMVP / neighbour / lowres MVs: produced by the
 *   caller's analysis, they are inputs here); csp != 0 with subpelRefine > 2 adds the Cb / Cr SATD to every subpelCompare
 *   (bChromaSATD, motion.cpp:212,1601-1661) for the shapes whose chroma block is a multiple of 4x4; DIA/HEX/UMH/STAR/FULL.
 * The search range of a PU is computed as the reference does: mvp +- merange, CUData::clipMv against the picture
 * (picWidth x picHeight; 0 = ctuCols*ctuSize x (firstCtuRow+ctuRows)*ctuSize) with the PU's CU position, slice rows when
 * maxSlices > 1 and frameParallel (the band semantics of --slices), refLagPixels (0 = no clamp), maxMvLen.
 * A call may cover a BAND of CTU rows of a larger frame (CTU-row sharding): pass the plane origins advanced to the
 * band's first CTU row (marginY grown by the same rows), firstCtuRow = that row inside the whole frame, picHeight = the whole
 * frame's and sliceTotalRows = the whole frame's CTU rows; CU positions for clipMv and the slice rows are then the frame's.  PUs of CUs that leave the picture
 * are not searched (cost -1).
 * One CTA per (CTU, reference) stages the search window around the CTU's window centre mvpCtu[ref][ctu] (quarter-pel, NULL =
 * 0; also the predictor when mvpPu is NULL) with TMA; blocks outside the window are read from the plane, so the results do
 * not depend on the centre.  Planes are given by their ORIGIN (pixel 0,0) with marginX / marginY pixels of padding around
 * the picture and rowsTotal rows allocated (PicYuv layout; chroma rows are the luma ones >> the vertical chroma shift, chroma margins see
 * chromaMarginX / chromaMarginY).
 * PU order inside a CTU: x265b200_me_frame_layout.  Arrays: mvpPu int32 [ref][ctu][pu][2], numCandPu uint8 [ref][ctu][pu],
 * mvcPu int32 [ref][ctu][pu][maxCand][2], out int32 [ref][ctu][pu][3] = {mvx, mvy, cost} (all device memory). */
typedef struct
{
    int32_t depth;                            /* 8, 10 or 12 */
    int32_t ctuSize, minCuSize, rect, amp;
    int32_t picWidth, picHeight;
    int32_t ctuCols, ctuRows;
    int32_t marginX, marginY, rowsTotal;
    int32_t chromaMarginX, chromaMarginY;     /* padding of the Cb / Cr planes; 0 = marginX >> hshift / marginY >> vshift.  PicYuv itself keeps
                                                 chromaMarginX = lumaMarginX (picyuv.cpp:104-106): pass its values when binding real PicYuv planes */
    int32_t numRefs;
    int32_t searchMethod, subpelRefine, merange;
    int32_t csp;                              /* 0 = luma only, 1 = 4:2:0, 2 = 4:2:2, 3 = 4:4:4 (x265.h:588-592) */
    int32_t maxCand;
    int32_t maxSlices, frameParallel, firstCtuRow, sliceTotalRows;
    int32_t refLagPixels;
    double  lambda;
} x265b200_me_frame_params;
typedef struct
{
    const void* curY; const void* curCb; const void* curCr;      /* device plane origins of the source picture */
    int64_t curStride, curStrideC;
    const void* const* refY; const void* const* refCb; const void* const* refCr;   /* HOST arrays [numRefs] of device plane origins */
    int64_t refStride, refStrideC;
} x265b200_me_frame_planes;
int x265b200_me_frame_ex_dev(x265b200_ctx* ctx, const x265b200_me_frame_params* params, const x265b200_me_frame_planes* planes,
                             const int32_t* mvpCtu, const int32_t* mvpPu, const uint8_t* numCandPu, const int32_t* mvcPu, int32_t* out);
/* Host-buffer forms of the two frame searches (the end-to-end path): the source picture arrives from HOST memory -- hostCurBase
 * points at the first byte of the padded plane (origin - marginY*stride - marginX), planeBytes bytes are copied to devCurBase
 * (caller-owned device plane of the same layout: it becomes a reference picture later) --, the search runs against the
 * device-resident references, and the {mvx, mvy, cost} records are copied to hostOut (outBytes; pinned memory makes both copies
 * DMA transfers).  Synchronous: returns when hostOut is complete.  devOut is the caller's device result buffer (same layout). */
int x265b200_me_frame_host(x265b200_ctx* ctx, int depth, const void* hostCurBase, size_t planeBytes, void* devCurBase, int64_t curStride,
                           const void* const* refOriginsHost, int numRefs, int64_t refStride,
                           int marginX, int marginY, int rowsTotal, int ctuCols, int ctuRows, int puMask,
                           const int32_t* mvpCtu, int searchMethod, int subpelRefine, int merange, double lambda,
                           int32_t* devOut, int32_t* hostOut, size_t outBytes);
/* The same in two halves, for callers that have more device work to queue behind the search (the stages that consume the MVs):
 * _begin returns once the transfers and the search are queued -- the source picture travels on the context's own copy stream, so it
 * overlaps whatever is still running on the compute stream, the search waits for it, and the results travel back on the copy stream
 * while the compute stream goes on; x265b200_me_frame_host_end blocks until the hostOut of the OLDEST pending call is complete.  Two calls
 * may be pending per context (queue frame t+1, then read frame t: the host never idles the device);
 * the destination plane must not be read by work queued before _begin (a picture buffer that left the reference list). */
int x265b200_me_frame_host_begin(x265b200_ctx* ctx, int depth, const void* hostCurBase, size_t planeBytes, void* devCurBase, int64_t curStride,
                                 const void* const* refOriginsHost, int numRefs, int64_t refStride,
                                 int marginX, int marginY, int rowsTotal, int ctuCols, int ctuRows, int puMask,
                                 const int32_t* mvpCtu, int searchMethod, int subpelRefine, int merange, double lambda,
                                 int32_t* devOut, int32_t* hostOut, size_t outBytes);
int x265b200_me_frame_host_end(x265b200_ctx* ctx);
/* planes->curY / curCb / curCr are the device ORIGINS as in x265b200_me_frame_ex_dev; hostCur*Base / devCur*Base the padded
 * planes' first bytes (Cb / Cr may be NULL when params->csp == 0 or subpelRefine <= 2). */
int x265b200_me_frame_ex_host(x265b200_ctx* ctx, const x265b200_me_frame_params* params, const x265b200_me_frame_planes* planes,
                              const void* hostCurYBase, void* devCurYBase, size_t bytesY,
                              const void* hostCurCbBase, void* devCurCbBase, const void* hostCurCrBase, void* devCurCrBase, size_t bytesC,
                              const int32_t* mvpCtu, const int32_t* mvpPu, const uint8_t* numCandPu, const int32_t* mvcPu,
                              int32_t* devOut, int32_t* hostOut, size_t outBytes);
int x265b200_me_frame_ex_host_begin(x265b200_ctx* ctx, const x265b200_me_frame_params* params, const x265b200_me_frame_planes* planes,
                              const void* hostCurYBase, void* devCurYBase, size_t bytesY,
                              const void* hostCurCbBase, void* devCurCbBase, const void* hostCurCrBase, void* devCurCrBase, size_t bytesC,
                              const int32_t* mvpCtu, const int32_t* mvpPu, const uint8_t* numCandPu, const int32_t* mvcPu,
                              int32_t* devOut, int32_t* hostOut, size_t outBytes);      /* ended by x265b200_me_frame_host_end */
/* The PUs of one CTU in the order x265b200_me_frame_ex_dev uses: {x, y, w, h} inside the CTU, CU sizes from ctuSize down to
 * minCuSize, CUs in raster order, per CU the part modes in PartSize order (2Nx2N, 2NxN, Nx2N, 2NxnU, 2NxnD, nLx2N, nRx2N;
 * common/cudata.h).  Writes at most cap entries; returns the number of PUs (host side, no GPU needed), -1 on bad arguments. */
int x265b200_me_frame_layout(int ctuSize, int minCuSize, int rect, int amp, int32_t* outXYWH, int cap);
/* The lambda-scaled MV cost table the kernels use (host side, no GPU needed):
 * out[2*32768 + i] = cost of a quarter-pel MV difference i, i in [-65536, 65536]. */
int x265b200_bitcost_table(double lambda, uint16_t* out);
/* x265_lambda_tab[qp] for a bit depth (constants.cpp:34-150), regenerated as round(2^((qp-12)/6 + (depth-8)), 4 dp) */
double x265b200_lambda(int qp, int depth);

/* ---- lookahead (lowres pre-analysis) ----------------------------------------------------------
 * x265b200_lowres_init_dev: replaces primitives.frameInitLowres + extendPicBorder x4 as used by
 *   Lowres::init (common/lowres.cpp:294-302; pixel.cpp:604-628, :1027-1041).  planes[4] are the
 *   ORIGINS (pixel 0,0) of the four hpel planes; margins are filled by edge replication.
 * x265b200_la_intra_dev: replaces LookaheadTLD::lowresIntraEstimate (slicetype.cpp:696-805) for one
 *   frame.  intraPenalty = 5 * (int)x265_lambda_tab[X265_LOOKAHEAD_QP].  sums[2] = costEst, costEstAq.
 * x265b200_la_estimate_dev: replaces CostEstimateGroup::estimateFrameCost / estimateCUCost
 *   (slicetype.cpp:3115-3388; HME off) for a batch of frame triples.  lookaheadSlices = the effective
 *   --lookahead-slices (param.cpp:173 default 8 at medium; 0 or 1 = the non-cooperative path): the field is cut into
 *   cooperative slices as Lookahead::create does (slicetype.cpp:1029-1041: heightInCU / slices rows each, at least 10; the
 *   last slice runs to the bottom) and the bottom row of every slice is searched with lastRow = true (processTasks,
 *   slicetype.cpp:3079-3107) -- bit-exact with the reference's coop-slice path, and the slices are independent wavefronts.  `planes` = device array [numFrames][4] of plane origins; MVs and MV costs of list i live
 *   in slot mvSlot[i] of mvPool ([slot][ncu][2] int32) / mvCostPool ([slot][ncu] int32) -- the
 *   reference's lowresMvs[i][dist] / lowresMvCosts[i][dist] cache; doSearch[i] = 0 reuses the slot.
 *   Outputs per triple t: lowresCosts[t][ncu] (uint16), rowSatds[t][heightInCU], sums[t][4] =
 *   {costEst (raw sum, before the b-frame 100/130 scaling of :3203-3204), costEstAq, intraMbs, 0}. */
typedef struct {
    int32_t b, p0, p1; int32_t doSearch[2]; int32_t mvSlot[2];
    int32_t weightIdx0;            /* weightp: 0 = list 0 searches p0's planes; k > 0 = when weights[k-1].isWeighted (x265b200_la_weights_analyse_dev) ... */
    int32_t weightPlanes0;         /* ... the list-0 search reads planes[weightPlanes0][4] (the wbuffer planes) instead (slicetype.cpp:3222) */
} x265b200_la_triple;
/* weightp in the lookahead: LookaheadTLD::weightsAnalyse (slicetype.cpp:860-961; weightCostLuma :805-841), called by
 * estimateFrameCost before every list-0 search when param->bEnableWeightedPred (:3136-3138).  One job = one (fenc, ref) pair:
 * fencPlane0 = fenc.lowresPlane[0] (origin), intraCost = fenc.intraCost, refBuffer[k] = ref.buffer[k] and weighted[k] = wbuffer[k]
 * (the FIRST element of each padded plane: weight_pp runs over stride x paddedLines elements, :947-956), fencSum / fencSsd /
 * refSum / refSsd = wp_sum[0] / wp_ssd[0] of the two frames as calcAdaptiveQuantFrame leaves them (:672-674;
 * x265b200_aq_energy_dev yields the raw sums).  The float scale / offset guess is evaluated on the host exactly as the reference
 * evaluates it; the two weightCostLuma passes (8x8 SATD against the unweighted and the weighted reference, each block capped by
 * its intra cost), the accept test and the weighting of the four planes run on the device without a host round trip.
 * out[j] (device): isWeighted and the chosen weight, origscore / score = the two weightCostLuma sums (0 when the early
 * termination of :895-897 applied).  x265b200_la_estimate_dev reads isWeighted through x265b200_la_triple.weightIdx0. */
typedef struct {
    const void* fencPlane0; const int32_t* intraCost;
    const void* refBuffer[4]; void* weighted[4];
    uint64_t fencSum, fencSsd, refSum, refSsd;
} x265b200_la_weight_job;
typedef struct { int32_t isWeighted, inputWeight, log2WeightDenom, inputOffset; uint32_t origscore, score; } x265b200_la_weight;
/* The host half of the above on its own (no GPU needed, like x265b200_bitcost_table): out = { measure, curScale, curDenom, curOffset, finScale,
 * finDenom, identity } -- measure 0 = the early termination of slicetype.cpp:895-897; (curScale, curDenom, curOffset) = the candidate the
 * second weightCostLuma pass measures; (finScale, finDenom, curOffset) = the weight applied when the candidate is accepted. */
int x265b200_la_weight_guess(int depth, int width, int lines, uint64_t fencSum, uint64_t fencSsd, uint64_t refSum, uint64_t refSsd, int32_t out[7]);
int x265b200_la_weights_analyse_dev(x265b200_ctx* ctx, int depth, const x265b200_la_weight_job* jobsHost, int numJobs,
                                    int64_t stride, int paddedLines, int64_t padOffset, int width, int lines, x265b200_la_weight* out);
int x265b200_lowres_init_dev(x265b200_ctx* ctx, int depth, const void* src, int64_t srcStride,
                             void* const planes[4], int64_t dstStride, int width, int height, int marginX, int marginY);
/* extendPicBorder (pixel.cpp:1027-1041: PicYuv / Lowres margins) and, with marginY = 0, primitives.extendRowBorder =
 * extendCURowColBorder (primitives.h:340; ipfilter.cpp:59-77): replicate the edge pixels of the width x height picture
 * at `origin` into marginX columns left/right and marginY rows above/below. */
int x265b200_extend_border_dev(x265b200_ctx* ctx, int depth, void* origin, int64_t stride, int width, int height, int marginX, int marginY);
int x265b200_la_intra_dev(x265b200_ctx* ctx, int depth, const void* plane0, int64_t stride, int widthInCU, int heightInCU,
                          const int32_t* invQscale, int intraPenalty, int32_t* intraCost, uint8_t* intraMode,
                          uint16_t* lowresCosts, int32_t* rowSatds, int32_t* sums);
int x265b200_la_estimate_dev(x265b200_ctx* ctx, int depth, const void* const* planes, int64_t stride,
                             int widthInCU, int heightInCU, const x265b200_la_triple* triplesHost, int numTriples,
                             int32_t* mvPool, int32_t* mvCostPool, const int32_t* const* intraCost,
                             const int32_t* const* invQscale, uint16_t* lowresCosts, int32_t* rowSatds, int32_t* sums,
                             double lambda, int lookaheadSlices, const x265b200_la_weight* weights /* device, or NULL: weightp off */);

/* --hme (hierarchical ME, param bEnableHME): estimateFrameCost first runs estimateCUCost(..., hme = true) over the
 * quarter-resolution planes (Lowres::lowerResPlane[4], built by primitives.frameInitLowerRes + extendPicBorder,
 * lowres.cpp:304-313 -- use x265b200_lowres_init_dev on lowresPlane[0]) with hmeSearchMethod[0] / hmeRange[0], then the
 * 8x8 level with hmeSearchMethod[1] / hmeRange[1] and the doubled quarter-resolution MV as one more candidate
 * (slicetype.cpp:3177-3188, :3216-3325; motion.cpp:751-754,817).  lowerPlanes = device array [numFrames][4] of plane
 * origins (row pitch lowerStride = lumaStride / 2); width4 x height4 = Lookahead::m_4x4Width x m_4x4Height
 * (slicetype.cpp:979-980); lowerMvPool / lowerMvCostPool = lowerResMvs / lowerResMvCosts, same slot numbering as mvPool
 * with width4*height4 entries per slot.  Everything else as x265b200_la_estimate_dev. */
typedef struct {
    const void* const* lowerPlanes; int64_t lowerStride;
    int32_t width4, height4;
    int32_t* lowerMvPool; int32_t* lowerMvCostPool;
    int32_t searchMethod[2];       /* hmeSearchMethod[0], [1] (x265.h:492-497 numbering; defaults HEX, UMH) */
    int32_t range[2];              /* hmeRange[0], [1] (defaults 16, 32) */
} x265b200_la_hme;
int x265b200_la_estimate_hme_dev(x265b200_ctx* ctx, int depth, const void* const* planes, int64_t stride, int widthInCU, int heightInCU,
                                 const x265b200_la_hme* hme, const x265b200_la_triple* triplesHost, int numTriples, int32_t* mvPool, int32_t* mvCostPool,
                                 const int32_t* const* intraCost, const int32_t* const* invQscale, uint16_t* lowresCosts,
                                 int32_t* rowSatds, int32_t* sums, double lambda, int lookaheadSlices, const x265b200_la_weight* weights);

/* Lookahead::estimateCUPropagate (slicetype.cpp:2641-2747), the cuTree propagation step of one (p0, p1, b): per 8x8 CU of frame b
 * the amount estimateCUPropagateCost yields (from propagateCostB = frames[b]->propagateCost, NULL for a non-referenced frame whose
 * source costs are zero; intraCost, lowresCosts[b-p0][p1-b], invQscaleFactor(8x8) of frame b; fpsFactor as computed at :2654) is
 * split over the lists the CU used (bipredWeight = 32, or 64 - (distScaleFactor >> 2) with weighted bi-prediction, :2644-2646),
 * follows mvs0 / mvs1 = lowresMvs[list][listDist[list]] and is added, saturating at 65535, to refCost0 / refCost1 =
 * frames[p0] / frames[p1]->propagateCost (refCost1 may be NULL or equal refCost0's frame for P frames).  The cuTreeFinish call at
 * the end of the reference function (VBV lookahead only) stays with the caller.  All arrays are device memory, widthInCU*heightInCU. */
int x265b200_cutree_propagate_dev(x265b200_ctx* ctx, int widthInCU, int heightInCU, const uint16_t* propagateCostB, const int32_t* intraCost,
                                  const uint16_t* lowresCosts, const int32_t* invQscale, const int32_t* mvs0, const int32_t* mvs1,
                                  uint16_t* refCost0, uint16_t* refCost1, int bipredWeight, double fpsFactor);

/* The block loop of LookaheadTLD::calcAdaptiveQuantFrame (slicetype.cpp:444-694): acEnergyCu (:252-275 = acEnergyPlane + acEnergyVar,
 * :49-86) for every qgSize x qgSize block of a frame, blocks in raster order over ceil(picWidth / qgSize) x ceil(picHeight / qgSize)
 * (edge blocks read the picture's padding, as the reference does).  energy[block] = the 32-bit sum over Y, Cb, Cr of
 * ssd - (sum^2 >> shift); wpSumSsd[0..2] / [3..5] = Lowres::wp_sum / wp_ssd of the three planes (what weightsAnalyse consumes).
 * csp: 0 = 4:0:0, 1 = 4:2:0, 2 = 4:2:2, 3 = 4:4:4; planes by ORIGIN.  The qp-offset formulas of the AQ modes are float / double
 * (X265_LOG2, pow) and stay on the host, fed with these integers. */
int x265b200_aq_energy_dev(x265b200_ctx* ctx, int depth, int csp, int qgSize, const void* y, int64_t strideY, const void* cb, const void* cr,
                           int64_t strideC, int picWidth, int picHeight, uint32_t* energy, uint64_t* wpSumSsd);
/* MotionReference::applyWeight for a whole plane (encoder/reference.cpp:119-185, weights set up as MotionReference::init :96-103):
 * dst = weight_pp(src) with (inputWeight, inputOffset, log2WeightDenom) of the slice's WeightParam, then the left / right / top /
 * bottom margins replicated -- the weighted reference plane the main motion search reads when weightp is on. */
int x265b200_apply_weight_dev(x265b200_ctx* ctx, int depth, const void* srcOrigin, void* dstOrigin, int64_t stride, int width, int height,
                              int marginX, int marginY, int inputWeight, int inputOffset, int log2WeightDenom);

#ifdef __cplusplus
}
#endif
#endif /* X265B200_H */

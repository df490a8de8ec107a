/* oracle/x265_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference algorithms on the hot path (x265 3.5+1,
 * msg7086/x265-Yuuki-Asuna).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this; the product library never does.
 * Parity status: PINNED -- every function here is checked against the reference itself
 * (oracle/_ref/libx265ref{8,10}.so compiled from /root/reference by oracle/Makefile) in
 * tests/test_oracle_vs_ref.py and against the committed fixtures in tests/golden/.
 *
 * All functions take `depth` (8, 10 or 12).  depth == 8 -> pixel is uint8_t, else uint16_t
 * (common/common.h:126-142).  Strides are in elements.
 */
#ifndef X265_ORACLE_H
#define X265_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

int      orc_sad(int depth, int w, int h, const void* a, intptr_t sa, const void* b, intptr_t sb);
void     orc_sad_xn(int depth, int K, int w, int h, const void* fenc, const void* const* refs, intptr_t refStride, int32_t* res);
int      orc_satd(int depth, int w, int h, const void* a, intptr_t sa, const void* b, intptr_t sb);
/* per16 != 0: round once per 16x16 (sa8d_16x16 / sa8d16<>), else per 8x8 (sa8d_8x8 / sa8d8<>) */
int      orc_sa8d(int depth, int w, int h, int per16, const void* a, intptr_t sa, const void* b, intptr_t sb);
uint64_t orc_sse_pp(int depth, int w, int h, const void* a, intptr_t sa, const void* b, intptr_t sb);
uint64_t orc_sse_ss(int depth, int w, int h, const int16_t* a, intptr_t sa, const int16_t* b, intptr_t sb);
uint64_t orc_ssd_s(int depth, int size, const int16_t* a, intptr_t sa);

#ifdef __cplusplus
}
#endif
#endif

/* ---- transforms (restatement of common/dct.cpp) ---- */
#ifdef __cplusplus
extern "C" {
#endif
/* idx 0..3 = 4/8/16/32-point DCT, 4 = 4x4 DST.  Forward: strided src -> N*N contiguous; inverse: N*N contiguous -> strided dst */
void     orc_dct(int depth, int idx, const int16_t* src, int16_t* dst, intptr_t srcStride);
void     orc_idct(int depth, int idx, const int16_t* src, int16_t* dst, intptr_t dstStride);
uint32_t orc_quant(const int16_t* coef, const int32_t* quantCoeff, int32_t* deltaU, int16_t* qCoef, int qBits, int add, int numCoeff);
uint32_t orc_nquant(const int16_t* coef, const int32_t* quantCoeff, int16_t* qCoef, int qBits, int add, int numCoeff);
void     orc_dequant_normal(const int16_t* quantCoef, int16_t* coef, int num, int scale, int shift);
void     orc_dequant_scaling(const int16_t* quantCoef, const int32_t* deQuantCoef, int16_t* coef, int num, int per, int shift);
/* interpolation: kind 0 hpp 1 hps 2 vpp 3 vps 4 vsp 5 vss 6 hvpp 7 p2s; taps 8 (luma) / 4 (chroma) */
void     orc_interp(int depth, int kind, int taps, int w, int h, const void* src, intptr_t srcStride, void* dst, intptr_t dstStride,
                    int idxX, int idxY, int isRowExt);
/* intra: neighbours [topLeft, top 2N, left 2N]; mode 0 planar, 1 DC, 2..34 angular */
void     orc_intra_pred(int depth, int log2N, int mode, int bFilter, const void* srcPix, void* dst, intptr_t dstStride);
void     orc_intra_filter(int depth, int log2N, const void* src, void* dst);
#ifdef __cplusplus
}
#endif

/* ---- SEA support, in-loop filters, cuTree, the residual chain (restatements of pixel.cpp:121-165, framefilter.cpp:39-140 + :722-825,
 *      loopfilter.cpp:39-180, sao.cpp:1762-1926, pixel.cpp:914-940, quant.cpp:397-480 + :543-605) ---- */
#ifdef __cplusplus
extern "C" {
#endif
int      orc_ads(int kind, int lxHalf, const int* encDC, const uint32_t* sums, intptr_t delta, const uint16_t* costMvX, int16_t* mvs, int width, int thresh);
void     orc_sea_integral(int depth, const void* reconOrigin, intptr_t stride, int padX, int padY, int maxHeight, uint32_t* const planes[12]);
void     orc_sao_apply(int kind, int depth, void* rec, intptr_t stride, int8_t* buf0, int8_t* buf1, const int8_t* offset, int width, int height, int startX);
void     orc_sao_stats(int kind, int depth, const int16_t* diff, const void* rec, intptr_t stride, int8_t* upBuff1, int8_t* upBufft,
                       int endX, int endY, int32_t* stats, int32_t* count);
void     orc_deblock(int chroma, int depth, void* src, intptr_t srcStep, intptr_t offset, int a, int b, int c);
void     orc_propagate_cost(int* dst, const uint16_t* propagateIn, const int32_t* intraCosts, const uint16_t* interCosts, const int32_t* invQscales,
                            double fpsFactor, int len);
uint32_t orc_tu_chain(int depth, int sizeIdx, int useDST, const void* fenc, intptr_t fencStride, const void* pred, intptr_t predStride,
                      void* recon, intptr_t reconStride, const int32_t* quantCoeff, int qBits, int add,
                      const int32_t* dequantCoef, int scaleOrPer, int dqShift, int16_t* coeff, uint64_t* sse);
#ifdef __cplusplus
}
#endif

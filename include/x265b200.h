/* x265b200.h -- C ABI of the B200-native primitive backend for x265 (libx265b200.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.
 * Each entry point names the reference interface it replaces (file:line under
 * /root/reference/source).  Two flavours exist for every family:
 *   *_dev  : all pointers are DEVICE pointers; asynchronous on the context's stream.
 *   *_host : all pointers are HOST pointers; the call stages H2D, runs the same kernel and
 *            copies results back before returning (this is what the pointer-compatible
 *            EncoderPrimitives slots and the end-to-end bench use).
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
int         x265b200_sync(x265b200_ctx* ctx);
void*       x265b200_stream(x265b200_ctx* ctx);                 /* cudaStream_t the kernels run on */
uint64_t    x265b200_launch_count(x265b200_ctx* ctx);           /* kernels launched so far        */
int         x265b200_malloc(x265b200_ctx* ctx, size_t bytes, void** devPtr);
int         x265b200_free(x265b200_ctx* ctx, void* devPtr);
int         x265b200_upload(x265b200_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int         x265b200_download(x265b200_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
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

/* sad_x3 / sad_x4: replaces pu[].sad_x3 / pu[].sad_x4 (primitives.h:139-140,248-249;
 * pixel.cpp:74-119).  Item i compares the cached 64-stride PU at fenc + i*fencBlockStride with
 * K (=3 or 4; 1..4 accepted) blocks ref + refOff[i*K+k], common refStride.  res: int32[n][K]. */
int x265b200_sad_xn_dev(x265b200_ctx* ctx, int depth, int K, int w, int h,
                        const void* fenc, int64_t fencBlockStride,
                        const void* ref, int64_t refStride, const int64_t* refOff,
                        int64_t n, int32_t* res);

#ifdef __cplusplus
}
#endif
#endif /* X265B200_H */

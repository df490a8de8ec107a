// scripts/tma_probe.cu -- standalone probe: which TMA tile-load configurations execute on this B200?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o scripts/tma_probe scripts/tma_probe.cu
// usage: tma_probe <boxW> <boxH> <variant>   variant 0: plain, 1: no ".tile", 2: shared::cta dst
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template<int VAR>
__global__ void probe(const __grid_constant__ CUtensorMap map, int boxW, int boxH, int x, int y, unsigned* out)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar)), "r"(boxW * boxH) : "memory");
        if (VAR == 0)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         :: "r"(smem_u32(smem)), "l"(&map), "r"(smem_u32(&bar)), "r"(x), "r"(y) : "memory");
        else if (VAR == 1)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         :: "r"(smem_u32(smem)), "l"(&map), "r"(smem_u32(&bar)), "r"(x), "r"(y) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         :: "r"(smem_u32(smem)), "l"(&map), "r"(smem_u32(&bar)), "r"(x), "r"(y) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" :: "r"(smem_u32(&bar)), "r"(0) : "memory");
    unsigned s = 0;
    for (int i = threadIdx.x; i < boxW * boxH; i += blockDim.x) s += smem[i];
    atomicAdd(out, s);
}

int main(int argc, char** argv)
{
    int boxW = argc > 1 ? atoi(argv[1]) : 64, boxH = argc > 2 ? atoi(argv[2]) : 64, var = argc > 3 ? atoi(argv[3]) : 0;
    const int W = 1024, H = 512;
    std::vector<uint8_t> h(W * H);
    for (int i = 0; i < W * H; i++) h[i] = (uint8_t)(i * 7 + (i >> 10));
    uint8_t* d; cudaMalloc(&d, W * H); cudaMemcpy(d, h.data(), W * H, cudaMemcpyHostToDevice);
    unsigned* dout; cudaMalloc(&dout, 4); cudaMemset(dout, 0, 4);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                            CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    CUtensorMap map;
    cuuint64_t gdim[2] = { W, H }; cuuint64_t gstr[1] = { W }; cuuint32_t box[2] = { (cuuint32_t)boxW, (cuuint32_t)boxH }; cuuint32_t es[2] = { 1, 1 };
    CUresult r = ((Enc)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box %dx%d var %d: encode=%d ", boxW, boxH, var, (int)r);
    size_t smem = (size_t)boxW * boxH + 256;
    int x = argc > 4 ? atoi(argv[4]) : 37, y = argc > 5 ? atoi(argv[5]) : 11;
    if (var == 0) { cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<0><<<1, 128, smem>>>(map, boxW, boxH, x, y, dout); }
    if (var == 1) { cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<1><<<1, 128, smem>>>(map, boxW, boxH, x, y, dout); }
    if (var == 2) { cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); probe<2><<<1, 128, smem>>>(map, boxW, boxH, x, y, dout); }
    cudaError_t e = cudaDeviceSynchronize();
    unsigned got = 0; cudaMemcpy(&got, dout, 4, cudaMemcpyDeviceToHost);
    unsigned exp = 0;
    for (int yy = 0; yy < boxH; yy++) for (int xx = 0; xx < boxW; xx++) exp += h[(y + yy) * W + x + xx];
    printf("sync=%s got=%u exp=%u %s\n", cudaGetErrorString(e), got, exp, (e == cudaSuccess && got == exp) ? "OK" : "FAIL");
    return 0;
}

// bisect: V=0 mbarrier only; V=1 1-D bulk copy; V=2 2-D tensor box (u16); V=3 3-D tensor box
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>
#ifndef CX
#define CX 8
#endif
#ifndef BW
#define BW 32
#endif
#ifndef BH
#define BH 16
#endif
#ifndef V
#define V 0
#endif
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap tmap, const uint16_t *src, uint16_t *out)
{
    __shared__ alignas(128) uint8_t box[2048];
    __shared__ alignas(8) uint64_t bar;
    const int lane = threadIdx.x;
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (lane == 0) {
#if V == 0
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(&bar)) : "memory");
#elif V == 1
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar)), "r"(1024) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(smem_u32(box)), "l"(src), "r"(1024), "r"(smem_u32(&bar)) : "memory");
#elif V == 2
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar)), "r"(32 * 16 * 2) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     :: "r"(smem_u32(box)), "l"(&tmap), "r"(8), "r"(4), "r"(smem_u32(&bar)) : "memory");
#else
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar)), "r"(BW * BH * 2) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     :: "r"(smem_u32(box)), "l"(&tmap), "r"(CX), "r"(4), "r"(1), "r"(smem_u32(&bar)) : "memory");
#endif
    }
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n"
        :: "r"(smem_u32(&bar)), "r"(0) : "memory");
    for (int i = lane; i < 512; i += 32)
        out[i] = reinterpret_cast<uint16_t *>(box)[i];
}
int main()
{
    const int W = 512, H = 240, B = 3;
    std::vector<uint16_t> h((size_t)W * H * B);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint16_t)i;
    uint16_t *d, *out;
    cudaMalloc(&d, h.size() * 2); cudaMalloc(&out, 4096);
    cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap map; memset(&map, 0, sizeof(map));
    const cuuint64_t dims[3] = { W, H, B }, strides[2] = { W * 2, (cuuint64_t)W * H * 2 };
    const cuuint32_t box[3] = { BW, BH, 1 }, es[3] = { 1, 1, 1 };
    CUresult r = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, V == 2 ? 2 : 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("V=%d encode: %d\n", V, (int)r);
    probe<<<1, 32>>>(map, d, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("V=%d: %s\n", V, cudaGetErrorString(e));
    if (e == cudaSuccess) {
        uint16_t o[8]; cudaMemcpy(o, out, 16, cudaMemcpyDeviceToHost);
        printf("  first: %u %u %u %u\n", o[0], o[1], o[2], o[3]);
    }
    return 0;
}

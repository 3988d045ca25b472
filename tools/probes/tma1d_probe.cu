// Probe: which 1-D tensor-map loads does sm_100a accept?  (aligned / unaligned start coordinate, partially out of bounds)
// nvcc -gencode arch=compute_100a,code=sm_100a -o tma1d_probe tma1d_probe.cu -lcuda ; ./tma1d_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

__global__ void probe(const __grid_constant__ CUtensorMap map, int x, float* out) {
    __shared__ __align__(128) float buf[128];
    __shared__ __align__(8) unsigned long long bar;
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(buf);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(512u) : "memory");
        asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3}], [%2];"
                     ::"r"(d), "l"(&map), "r"(b), "r"(x) : "memory");
    }
    __syncthreads();
    asm volatile("{ .reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0; @p bra D; bra W; D: }" ::"r"(b) : "memory");
    out[threadIdx.x] = buf[threadIdx.x];
}

int main() {
    const int n = 1000;
    float* src; float* out;
    cudaMalloc(&src, 4096 * 4); cudaMalloc(&out, 128 * 4);
    float h[4096]; for (int i = 0; i < 4096; ++i) h[i] = (float)i;
    cudaMemcpy(src, h, sizeof(h), cudaMemcpyHostToDevice);
    typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* sym; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    Enc enc = (Enc)sym;
    for (int promo = 0; promo < 2; ++promo) {
        CUtensorMap map;
        cuuint64_t dims[1] = {(cuuint64_t)n}, strides[1] = {0};
        cuuint32_t box[1] = {128}, es[1] = {1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 1, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode rank1 promo=%d -> %d\n", promo, (int)r);
        const int xs[5] = {0, 128, 1, 37, 950};
        for (int k = 0; k < 5; ++k) {
            probe<<<1, 128>>>(map, xs[k], out);
            cudaError_t e = cudaDeviceSynchronize();
            float o[128]; cudaMemcpy(o, out, sizeof(o), cudaMemcpyDeviceToHost);
            printf("  x=%d: %s  first=%g last=%g\n", xs[k], cudaGetErrorString(e), o[0], o[127]);
            if (e != cudaSuccess) return 1;
        }
    }
    return 0;
}

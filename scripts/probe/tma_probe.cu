// tma_probe.cu -- developer probe: does cuTensorMapEncodeTiled accept FP64 tensor maps whose
// dimension strides are NOT monotonically increasing (a digit-reversing store box), and do
// cp.async.bulk.tensor load/store round-trip such a box?  Prints PASS/FAIL lines.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// load a (C, R) box (natural rows), store it through a (C, r2, r1) box with strides (.., r1*rowbytes, rowbytes)
__global__ void probe_kernel(const __grid_constant__ CUtensorMap lmap, const __grid_constant__ CUtensorMap smap, int C, int R, int col0)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long mbar;
    double *tile = (double *)smem;
    const uint32_t bar = smem_u32(&mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)(C * R * 8);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(tile)), "l"(&lmap), "r"(col0), "r"(0), "r"(bar) : "memory");
    }
    // everybody waits for phase 0
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(bar), "r"(0) : "memory");
        }
    }
    for (int i = threadIdx.x; i < C * R; i += blockDim.x) tile[i] = tile[i] * 2.0 + 1.0;   // touch in the generic proxy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                     ::"l"(&smap), "r"(col0), "r"(0), "r"(0), "r"(smem_u32(tile)) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

int main()
{
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &q) != cudaSuccess || !encode) {
        printf("FAIL no cuTensorMapEncodeTiled\n");
        return 1;
    }
    const int r1 = 4, r2 = 8, R = r1 * r2, inner_cplx = 24, C = 8 /* doubles per box row = T*2, T=4 */;
    const int rowd = 2 * inner_cplx;   // doubles per global row
    std::vector<double> hin((size_t)R * rowd), hout((size_t)R * rowd, -1.0);
    for (size_t i = 0; i < hin.size(); ++i) hin[i] = (double)i;
    double *din, *dout;
    cudaMalloc(&din, hin.size() * 8); cudaMalloc(&dout, hout.size() * 8);
    cudaMemcpy(din, hin.data(), hin.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dout, hout.data(), hout.size() * 8, cudaMemcpyHostToDevice);
    CUtensorMap lmap, smap;
    {
        cuuint64_t dims[2] = {(cuuint64_t)rowd, (cuuint64_t)R};
        cuuint64_t strides[1] = {(cuuint64_t)rowd * 8};
        cuuint32_t box[2] = {(cuuint32_t)C, (cuuint32_t)R};
        cuuint32_t es[2] = {1, 1};
        CUresult rc = encode(&lmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, din, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("%s encode load map rc=%d\n", rc == CUDA_SUCCESS ? "PASS" : "FAIL", (int)rc);
        if (rc != CUDA_SUCCESS) return 1;
    }
    {
        // smem row p = k1*r2 + k2 (k2 fastest)  ->  global row K = k1 + r1*k2
        cuuint64_t dims[3] = {(cuuint64_t)rowd, (cuuint64_t)r2, (cuuint64_t)r1};
        cuuint64_t strides[2] = {(cuuint64_t)r1 * rowd * 8, (cuuint64_t)rowd * 8};     // DEcreasing
        cuuint32_t box[3] = {(cuuint32_t)C, (cuuint32_t)r2, (cuuint32_t)r1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult rc = encode(&smap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, dout, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("%s encode digit-reversing store map (decreasing strides) rc=%d\n", rc == CUDA_SUCCESS ? "PASS" : "FAIL", (int)rc);
        if (rc != CUDA_SUCCESS) return 2;
    }
    const int col0 = 16;   // doubles
    probe_kernel<<<1, 128, C * R * 8>>>(lmap, smap, C, R, col0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s kernel: %s\n", e == cudaSuccess ? "PASS" : "FAIL", cudaGetErrorString(e));
    if (e != cudaSuccess) return 3;
    cudaMemcpy(hout.data(), dout, hout.size() * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int p = 0; p < R; ++p) {
        const int k1 = p / r2, k2 = p % r2, K = k1 + r1 * k2;
        for (int c = 0; c < C; ++c) {
            const double want = hin[(size_t)p * rowd + col0 + c] * 2.0 + 1.0;
            if (hout[(size_t)K * rowd + col0 + c] != want) ++bad;
        }
    }
    int untouched = 0;
    for (int K = 0; K < R; ++K) for (int c = 0; c < rowd; ++c) if ((c < col0 || c >= col0 + C) && hout[(size_t)K * rowd + c] != -1.0) ++untouched;
    printf("%s permuted store: %d wrong, %d stray writes\n", (bad == 0 && untouched == 0) ? "PASS" : "FAIL", bad, untouched);
    return bad || untouched;
}

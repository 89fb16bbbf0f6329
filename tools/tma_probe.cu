// Micro-probe: how fast can ONE SM ingest 128B-swizzled bf16 tiles through TMA (no MMA)?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_probe.bin tools/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include "../point-sam_b200/csrc/psam_common.cuh"
using namespace psam;

struct P { int iters, stages, boxes, box_rows, shared_rows, rank5, nrows; };

__global__ void __launch_bounds__(64, 1) probe(const __grid_constant__ CUtensorMap tm, P p) {
    extern __shared__ unsigned char smem[];
    __shared__ __align__(8) uint64_t full[16], empty[16];
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
        fence_mbar_init();
    }
    __syncthreads();
    const uint32_t box_bytes = p.box_rows * 128;
    const uint32_t stage_bytes = p.boxes * box_bytes;
    if (warp == 0 && lane == 0) {
        for (int i = 0; i < p.iters; ++i) {
            const int s = i % p.stages; const uint32_t ph = (i / p.stages) & 1;
            mbar_wait(smem_u32(&empty[s]), ph ^ 1);
            mbar_arrive_expect_tx(smem_u32(&full[s]), stage_bytes);
            for (int b = 0; b < p.boxes; ++b) {
                const int k0 = (i % 16) * 64;
                int row = p.shared_rows ? (b * p.box_rows) : ((blockIdx.x * p.boxes + b) * p.box_rows) % p.nrows;
                if (p.rank5) tma_load_5d(base + s * stage_bytes + b * box_bytes, &tm, smem_u32(&full[s]), k0, row, 0, 0, 0);
                else asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                                  ::"r"(base + s * stage_bytes + b * box_bytes), "l"(&tm), "r"(smem_u32(&full[s])), "r"(k0), "r"(row) : "memory");
            }
        }
    } else if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.iters; ++i) {
            const int s = i % p.stages; const uint32_t ph = (i / p.stages) & 1;
            mbar_wait(smem_u32(&full[s]), ph);
            mbar_arrive(smem_u32(&empty[s]));
        }
    }
}

typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    setvbuf(stdout, NULL, _IONBF, 0);
    const int K = 1024, NROWS = 65536;
    __nv_bfloat16* d; cudaMalloc(&d, (size_t)2 * NROWS * K * 2); cudaMemset(d, 0, (size_t)2 * NROWS * K * 2);
    void* fp; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    PFN enc = (PFN)fp;
    printf("ctas stages boxes box_rows shared rank5 | us  GB/s total  GB/s per SM\n");
    int cfgs[][6] = {
        {1, 3, 4, 128, 0, 1}, {1, 6, 2, 128, 0, 1}, {1, 3, 4, 128, 0, 0}, {1, 6, 2, 128, 0, 0}, {1, 12, 1, 128, 0, 0}, {1, 3, 2, 256, 0, 0},
        {148, 3, 4, 128, 0, 1}, {148, 3, 4, 128, 0, 0}, {148, 6, 2, 128, 0, 0}, {148, 3, 4, 128, 1, 0}, {148, 3, 4, 128, 1, 1},
        {96, 3, 4, 128, 0, 1}, {96, 3, 4, 128, 0, 0}, {32, 3, 4, 128, 0, 0}, {148, 3, 2, 256, 0, 0}, {148, 7, 1, 256, 0, 0}, {148, 4, 4, 96, 0, 0}};
    for (auto& c : cfgs) {
        P p; p.iters = 256; p.stages = c[1]; p.boxes = c[2]; p.box_rows = c[3]; p.shared_rows = c[4]; p.rank5 = c[5]; p.nrows = NROWS - 256;
        CUtensorMap tm;
        if (p.rank5) {
            cuuint64_t dims[5] = {K, NROWS, 2, 1, 1}; cuuint64_t str[4] = {K * 2, (cuuint64_t)NROWS * K * 2, 16, 16};
            cuuint32_t box[5] = {64, (cuuint32_t)p.box_rows, 1, 1, 1}, es[5] = {1, 1, 1, 1, 1};
            enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else {
            cuuint64_t dims[2] = {K, NROWS}; cuuint64_t str[1] = {K * 2};
            cuuint32_t box[2] = {64, (cuuint32_t)p.box_rows}, es[2] = {1, 1};
            enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        size_t smem = (size_t)p.stages * p.boxes * p.box_rows * 128 + 1024;
        cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        probe<<<c[0], 64, smem>>>(tm, p); cudaDeviceSynchronize();
        cudaEventRecord(e0); probe<<<c[0], 64, smem>>>(tm, p); cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double bytes = (double)c[0] * p.iters * p.boxes * p.box_rows * 128;
        printf("%4d %2d %2d %4d %d %d | %8.1f %9.1f %8.1f  (%s)\n", c[0], c[1], c[2], c[3], c[4], c[5], ms * 1e3, bytes / ms / 1e6, bytes / ms / 1e6 / c[0], cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}

// Micro-probe: tcgen05.mma dispatch rate on one SM.  A CTA issues `iters` UMMA 128 x N x 16 (kind::f16) back to back on
// resident operands (no TMA in the loop) and reports clocks per MMA, for
//   mode 0: A and B from shared memory (SS), one issuing thread
//   mode 1: A from tensor memory (TS), B from shared memory, one issuing thread
//   mode 2: SS, TWO issuing threads (different warps) on disjoint accumulators
//   mode 3: SS + a second warp streaming tcgen05.ld of 128 other TMEM columns meanwhile (TMEM read-port contention)
//   mode 4: TS + the tcgen05.ld stream
//   mode 5/6/7: ONE issuing thread cycling over 2 / 4 / 3 disjoint accumulators (SS)   - is the floor a per-accumulator
//   mode 8: ONE issuing thread cycling over 2 disjoint accumulators, A from TMEM (TS)    dependency or a per-thread issue cost?
//   mode 9: SS, TWO issuing threads accumulating into the SAME accumulator (k-steps split between them)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_issue_probe.bin tools/mma_issue_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdint.h>
#include "../point-sam_b200/csrc/psam_common.cuh"
using namespace psam;

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

struct R { long long clocks; };

__global__ void __launch_bounds__(192, 1) probe(int mode, int N, int iters, R* out) {
    extern __shared__ unsigned char smem[];
    __shared__ __align__(8) uint64_t done[2];
    __shared__ uint32_t tmem_slot;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;  // A tile 128 x 64 (16 KB) then B tile 256 x 64 (32 KB), zeros
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem + (base - smem_u32(smem)))[i] = 0u;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&done[0]), 1); mbar_init(smem_u32(&done[1]), 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_slot;
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint64_t adesc = umma_desc_k_sw128(base), bdesc = umma_desc_k_sw128(base + 16384);
    const bool two = mode == 2 || mode == 9;
    long long t0 = clock64();
    if ((warp == 0 || (two && warp == 1)) && lane == 0) {
        const uint32_t d = tm + ((warp == 1 && mode != 9) ? 256u : 0u);
        const int n = two ? iters / 2 : iters;
        const int chains = mode == 5 || mode == 8 ? 2 : (mode == 6 ? 4 : (mode == 7 ? 3 : 1));
        if (chains > 1) {
            for (int i = 0; i < n; ++i) {
                const uint32_t dd = tm + (uint32_t)((i % chains) * (N <= 64 ? 64 : 128));
                if (mode == 8) umma_ts(dd, tm + 384u + (uint32_t)((i & 3) * 8), bdesc + (uint64_t)((i & 3) * 2), idesc, 1u);
                else umma_bf16(dd, adesc + (uint64_t)((i & 3) * 2), bdesc + (uint64_t)((i & 3) * 2), idesc, 1u);
            }
        } else
        for (int i = 0; i < n; ++i) {
            if (mode == 1 || mode == 4) umma_ts(d, tm + 384u + (uint32_t)((i & 3) * 8), bdesc + (uint64_t)((i & 3) * 2), idesc, 1u);
            else umma_bf16(d, adesc + (uint64_t)((i & 3) * 2), bdesc + (uint64_t)((i & 3) * 2), idesc, 1u);
        }
        umma_commit(smem_u32(&done[warp]));
    }
    if ((mode == 3 || mode == 4) && warp >= 2) {  // 4 warps read their lane quarter of 128 columns over and over
        const uint32_t q = (uint32_t)((warp & 3) * 32) << 16;
        uint32_t acc = 0;
        for (int r = 0; r < iters / 8; ++r) {
            uint32_t v[32];
            tmem_ld_32x32(tm + q + 256u + (uint32_t)((r & 3) * 32), v);
            tmem_ld_wait();
            acc += v[0] + v[31];
        }
        if (acc == 0x12345u) out[1].clocks = acc;
    }
    if (warp == 0) {
        mbar_wait(smem_u32(&done[0]), 0);
        if (two) mbar_wait(smem_u32(&done[1]), 0);
        if (lane == 0) out[0].clocks = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
    R* d; cudaMalloc(&d, 2 * sizeof(R));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768 + 1024);
    const int iters = 4096;
    const char* names[] = {"SS one issuer", "TS (A in TMEM) one issuer", "SS two issuers", "SS + tcgen05.ld stream", "TS + tcgen05.ld stream",
                           "SS one issuer, 2 accum", "SS one issuer, 4 accum", "SS one issuer, 3 accum", "TS one issuer, 2 accum",
                           "SS two issuers, SAME accum"};
    printf("mode | N | clocks per MMA | ideal (128 N / 256)\n");
    for (int mode = 0; mode < 10; ++mode)
        for (int N = 64; N <= 256; N *= 2) {
            if (mode >= 5 && mode <= 8 && N == 256) continue;  // 4 x 128 columns at most (and column 384+ holds the TS operand)
            if (mode == 6 && N == 128) continue;
            probe<<<1, 192, 16384 + 32768 + 1024>>>(mode, N, iters, d);
            probe<<<1, 192, 16384 + 32768 + 1024>>>(mode, N, iters, d);
            R h[2]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            cudaError_t e = cudaGetLastError();
            printf("%-28s | %3d | %7.1f | %5.1f %s\n", names[mode], N, (double)h[0].clocks / iters, 128.0 * N / 256.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}

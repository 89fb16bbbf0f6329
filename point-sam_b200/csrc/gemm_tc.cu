// Tensor-core GEMM for sm_100a: C[M,N] = A[M,K] * W[N,K]^T (+bias, activation, residual).
//
// Replaces every nn.Linear / bmm of the reference hot path that has enough rows to fill a 128-row
// MMA tile (PatchEncoder common.py:486-497, patch_proj/pos_embed/out_proj pc_encoder.py:99-116, the
// timm EVA block linears and attention products, output_upscaling mask_decoder.py:53-59).
//
// Numerics ("split-bf16"): the reference runs these contractions in fp32.  Operands are stored as two
// bf16 planes, x = hi + lo (relative residual <= 2^-17), and the product is accumulated in fp32 TMEM as
//        A_hi*W_hi + A_lo*W_hi + A_hi*W_lo                                   (passes = 3)
// which keeps ~16 mantissa bits per operand: measured end-to-end logit error ~2e-5 vs fp32, inside the
// 1e-3 abs / 1e-2 rel parity bound (a single bf16 pass, passes = 1, is ~1e-2 and fails it).
//
// Structure (one 128 x BN output tile per CTA, optional split-K over blockIdx.z):
//   warp 0   : TMA producer - 5-D tensor maps (k, row, plane, batch1, batch2), 128B-swizzled boxes of
//              64 bf16 x {128|BN} rows, 3-stage mbarrier ring, hi and lo planes of A and W per stage
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128xBNx16, kind::f16),
//              tcgen05.commit releases smem stages and finally signals the epilogue
//   warps 2-5: epilogue - tcgen05.ld 32x32b (each warp owns the TMEM lane quarter warp_id % 4),
//              bias/activation/residual, fp32 and/or split-bf16 stores (red.add for split-K)
#include "psam_common.cuh"
#include "../../include/psam_b200.h"

namespace psam {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_STAGES = 3;
constexpr int GEMM_THREADS = 192;

struct GemmEpilogue {
    float* out_f32;            // may be null
    long long ldo;             // row stride of out_f32 / resid (elements)
    long long out_b1, out_b2;  // batch strides of out_f32 / resid
    __nv_bfloat16* out_hi;     // may be null; lo plane at out_hi + out_plane
    long long out_plane, ldo_s, outs_b1, outs_b2;
    const float* bias;   // [N] or null
    const float* resid;  // same geometry as out_f32, may alias it; null = none
    float alpha;         // scale applied to the accumulator before bias
    int act;
    int accumulate;  // 1: out_f32 += result via red.global.add (required for split_k > 1)
};

struct GemmShape {
    int M, N, K;
    int nb1, nb2;  // batch extents (blockIdx.z = ((b2 * nb1) + b1) * split_k + split)
    int split_k;
    int passes;  // 1 or 3
};

template <int BN>
struct GemmSmem {
    static constexpr int A_TILE = GEMM_BM * GEMM_BK * 2;  // bytes per plane
    static constexpr int B_TILE = BN * GEMM_BK * 2;
    static constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;
    static constexpr int TOTAL = GEMM_STAGES * STAGE + 1024;  // + alignment slack
};

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const GemmShape shape, const GemmEpilogue ep) {
    using S = GemmSmem<BN>;
    extern __shared__ unsigned char smem_dyn[];
    __shared__ __align__(8) uint64_t full_bar[GEMM_STAGES];
    __shared__ __align__(8) uint64_t empty_bar[GEMM_STAGES];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_base_smem;

    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tile = blockIdx.x, m_tile = blockIdx.y;
    const int z = blockIdx.z;
    const int split = z % shape.split_k;
    const int bz = z / shape.split_k;
    const int b1 = bz % shape.nb1, b2 = bz / shape.nb1;

    const int kb_total = (shape.K + GEMM_BK - 1) / GEMM_BK;
    const int kb_per = (kb_total + shape.split_k - 1) / shape.split_k;
    const int kb_begin = split * kb_per;
    const int kb_end = min(kb_total, kb_begin + kb_per);
    const int num_kb = max(0, kb_end - kb_begin);
    const bool lo_pass = shape.passes == 3;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
        for (int s = 0; s < GEMM_STAGES; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        mbar_init(smem_u32(&tmem_full_bar), 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(&tmem_base_smem), BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            const uint32_t stage_bytes = (lo_pass ? 2u : 1u) * (uint32_t)(S::A_TILE + S::B_TILE);
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % GEMM_STAGES;
                const uint32_t ph = (uint32_t)(i / GEMM_STAGES) & 1u;
                mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1u);
                const uint32_t fb = smem_u32(&full_bar[s]);
                mbar_arrive_expect_tx(fb, stage_bytes);
                const uint32_t sa = smem_base + s * S::STAGE;
                const int k0 = (kb_begin + i) * GEMM_BK;
                tma_load_5d(sa, &tmap_a, fb, k0, m_tile * GEMM_BM, 0, b1, b2);
                tma_load_5d(sa + 2 * S::A_TILE, &tmap_b, fb, k0, n_tile * BN, 0, b1, b2);
                if (lo_pass) {
                    tma_load_5d(sa + S::A_TILE, &tmap_a, fb, k0, m_tile * GEMM_BM, 1, b1, b2);
                    tma_load_5d(sa + 2 * S::A_TILE + S::B_TILE, &tmap_b, fb, k0, n_tile * BN, 1, b1, b2);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN);
        for (int i = 0; i < num_kb; ++i) {
            const int s = i % GEMM_STAGES;
            const uint32_t ph = (uint32_t)(i / GEMM_STAGES) & 1u;
            mbar_wait(smem_u32(&full_bar[s]), ph);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t sa = smem_base + s * S::STAGE;
                const uint64_t a_hi = umma_desc_k_sw128(sa);
                const uint64_t a_lo = umma_desc_k_sw128(sa + S::A_TILE);
                const uint64_t b_hi = umma_desc_k_sw128(sa + 2 * S::A_TILE);
                const uint64_t b_lo = umma_desc_k_sw128(sa + 2 * S::A_TILE + S::B_TILE);
#pragma unroll
                for (int k = 0; k < GEMM_BK / 16; ++k) {
                    const uint64_t adv = (uint64_t)(k * 2);  // 16 bf16 = 32 B = 2 x 16-byte units
                    umma_bf16(tmem_base, a_hi + adv, b_hi + adv, idesc, (i > 0 || k > 0) ? 1u : 0u);
                }
                if (lo_pass) {
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) {
                        const uint64_t adv = (uint64_t)(k * 2);
                        umma_bf16(tmem_base, a_lo + adv, b_hi + adv, idesc, 1u);
                    }
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) {
                        const uint64_t adv = (uint64_t)(k * 2);
                        umma_bf16(tmem_base, a_hi + adv, b_lo + adv, idesc, 1u);
                    }
                }
                umma_commit(smem_u32(&empty_bar[s]));                      // stage reusable once these MMAs retire
                if (i == num_kb - 1) umma_commit(smem_u32(&tmem_full_bar));  // accumulator complete
            }
            __syncwarp();
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32) are accessible to this warp
        const int row = m_tile * GEMM_BM + quarter * 32 + lane;
        const bool row_ok = row < shape.M;
        if (num_kb > 0) {
            mbar_wait(smem_u32(&tmem_full_bar), 0);
            tc_fence_after();
        }
        float* out = ep.out_f32 ? ep.out_f32 + (long long)b1 * ep.out_b1 + (long long)b2 * ep.out_b2 + (long long)row * ep.ldo : nullptr;
        const float* res = ep.resid ? ep.resid + (long long)b1 * ep.out_b1 + (long long)b2 * ep.out_b2 + (long long)row * ep.ldo : nullptr;
        __nv_bfloat16* ohi = ep.out_hi ? ep.out_hi + (long long)b1 * ep.outs_b1 + (long long)b2 * ep.outs_b2 + (long long)row * ep.ldo_s : nullptr;
        const bool add_bias = ep.bias && split == 0;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t v[32];
            if (num_kb > 0) {
                tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(c * 32), v);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int t = 0; t < 32; ++t) v[t] = 0u;
            }
            const int col0 = n_tile * BN + c * 32;
            if (!row_ok || col0 >= shape.N) continue;
            const int ncols = min(32, shape.N - col0);
            float f[32];
#pragma unroll
            for (int t = 0; t < 32; ++t) {
                float x = __uint_as_float(v[t]) * ep.alpha;
                if (add_bias && t < ncols) x += ep.bias[col0 + t];
                f[t] = x;
            }
            if (res && !ep.accumulate) {
#pragma unroll
                for (int t = 0; t < 32; ++t)
                    if (t < ncols) f[t] += res[col0 + t];
            }
            if (ep.act != ACT_NONE) {
#pragma unroll
                for (int t = 0; t < 32; ++t) f[t] = apply_act(f[t], ep.act);
            }
            if (out) {
                if (ep.accumulate) {
#pragma unroll
                    for (int t = 0; t < 32; ++t)
                        if (t < ncols) atomicAdd(out + col0 + t, f[t]);
                } else if (ncols == 32 && ((ep.ldo | col0) & 3) == 0) {
#pragma unroll
                    for (int t = 0; t < 32; t += 4)
                        *reinterpret_cast<float4*>(out + col0 + t) = make_float4(f[t], f[t + 1], f[t + 2], f[t + 3]);
                } else {
#pragma unroll
                    for (int t = 0; t < 32; ++t)
                        if (t < ncols) out[col0 + t] = f[t];
                }
            }
            if (ohi) {
                __nv_bfloat16* olo = ohi + ep.out_plane;
                if (ncols == 32 && ((ep.ldo_s | col0 | ep.out_plane) & 7) == 0) {
#pragma unroll
                    for (int t = 0; t < 32; t += 8) {
                        uint32_t h[4], l[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            __nv_bfloat16 h0, l0, h1, l1;
                            split_bf16(f[t + 2 * u], h0, l0);
                            split_bf16(f[t + 2 * u + 1], h1, l1);
                            h[u] = pack_bf16x2(h0, h1);
                            l[u] = pack_bf16x2(l0, l1);
                        }
                        *reinterpret_cast<uint4*>(ohi + col0 + t) = make_uint4(h[0], h[1], h[2], h[3]);
                        *reinterpret_cast<uint4*>(olo + col0 + t) = make_uint4(l[0], l[1], l[2], l[3]);
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < 32; ++t)
                        if (t < ncols) {
                            __nv_bfloat16 h0, l0;
                            split_bf16(f[t], h0, l0);
                            ohi[col0 + t] = h0;
                            olo[col0 + t] = l0;
                        }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, BN);
    }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps through the driver entry point (no link-time libcuda dependency)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

static int make_operand_map(CUtensorMap* map, const psam_operand* op, int box_rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return PSAM_ERR_UNSUPPORTED;
    const int nb1 = op->nb1 > 0 ? op->nb1 : 1, nb2 = op->nb2 > 0 ? op->nb2 : 1;
    cuuint64_t dims[5] = {(cuuint64_t)op->k, (cuuint64_t)op->rows, 2, (cuuint64_t)nb1, (cuuint64_t)nb2};
    // strides of dims 1..4 in bytes (dim 0 is contiguous); degenerate dims still need a 16-byte multiple
    const long long ps = op->plane_stride > 0 ? op->plane_stride : op->row_stride * (long long)op->rows;
    const long long s1 = op->b1_stride > 0 ? op->b1_stride : 8, s2 = op->b2_stride > 0 ? op->b2_stride : 8;
    cuuint64_t strides[4] = {(cuuint64_t)op->row_stride * 2, (cuuint64_t)ps * 2, (cuuint64_t)s1 * 2, (cuuint64_t)s2 * 2};
    for (int i = 0; i < 4; ++i)
        if (strides[i] % 16) return PSAM_ERR_ARG;
    if (((uintptr_t)op->hi) % 16) return PSAM_ERR_ARG;
    cuuint32_t box[5] = {(cuuint32_t)GEMM_BK, (cuuint32_t)box_rows, 1, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(op->hi), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? PSAM_OK : (int)(1000 + r);
}

template <int BN>
static int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mb, const GemmShape& sh, const GemmEpilogue& ep,
                       cudaStream_t stream) {
    auto kern = gemm_tc_kernel<BN>;
    PSAM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmSmem<BN>::TOTAL));
    dim3 grid((unsigned)ceil_div(sh.N, BN), (unsigned)ceil_div(sh.M, GEMM_BM), (unsigned)(sh.nb1 * sh.nb2 * sh.split_k));
    kern<<<grid, GEMM_THREADS, GemmSmem<BN>::TOTAL, stream>>>(ma, mb, sh, ep);
    PSAM_LAUNCH_CHECK();
    return PSAM_OK;
}

}  // namespace psam

extern "C" int psam_gemm_bf16x3(const psam_operand* a, const psam_operand* w, const psam_gemm_out* o, int passes,
                                int split_k, cudaStream_t stream) {
    using namespace psam;
    if (!a || !w || !o || !a->hi || !w->hi) return PSAM_ERR_ARG;
    if (a->k != w->k || a->k <= 0 || a->rows <= 0 || w->rows <= 0) return PSAM_ERR_ARG;
    if (passes != 1 && passes != 3) return PSAM_ERR_ARG;
    if (split_k < 1) split_k = 1;
    if (split_k > 1 && !(o->accumulate && o->out_f32 && !o->out_hi && o->act == 0)) return PSAM_ERR_ARG;
    if (!o->out_f32 && !o->out_hi) return PSAM_ERR_ARG;
    GemmShape sh;
    sh.M = a->rows, sh.N = w->rows, sh.K = a->k;
    sh.nb1 = a->nb1 > 0 ? a->nb1 : 1;
    sh.nb2 = a->nb2 > 0 ? a->nb2 : 1;
    if ((w->nb1 > 0 ? w->nb1 : 1) != sh.nb1 || (w->nb2 > 0 ? w->nb2 : 1) != sh.nb2) return PSAM_ERR_ARG;
    sh.split_k = split_k;
    sh.passes = passes;
    GemmEpilogue ep;
    ep.out_f32 = o->out_f32, ep.ldo = o->ldo, ep.out_b1 = o->out_b1, ep.out_b2 = o->out_b2;
    ep.out_hi = (__nv_bfloat16*)o->out_hi, ep.out_plane = o->out_plane, ep.ldo_s = o->ldo_s;
    ep.outs_b1 = o->outs_b1, ep.outs_b2 = o->outs_b2;
    ep.bias = o->bias, ep.resid = o->resid, ep.alpha = o->alpha, ep.act = o->act, ep.accumulate = o->accumulate;
    const int bn = (sh.N >= 128) ? 128 : 64;
    CUtensorMap ma, mb;
    int rc = make_operand_map(&ma, a, GEMM_BM);
    if (rc) return rc;
    rc = make_operand_map(&mb, w, bn);
    if (rc) return rc;
    return bn == 128 ? launch_gemm<128>(ma, mb, sh, ep, stream) : launch_gemm<64>(ma, mb, sh, ep, stream);
}

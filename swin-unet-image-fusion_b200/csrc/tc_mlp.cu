// Fused MLP on tcgen05 (one CTA per 128-token tile): LN -> GEMM1 -> +b1 -> ELU -> bf16 back to smem
// as the A operand of GEMM2, accumulated over hidden chunks in TMEM -> +b2 + residual.  The 4x
// hidden activation never leaves the SM.  Used for the small-channel stages where that activation
// would otherwise dominate HBM traffic.
#include <initializer_list>
#include "bf16_kernels.cuh"
#include "tc_common.cuh"

namespace sf {
using namespace tc;

static constexpr uint32_t LBO_A = lbo_padded(128);
static constexpr uint32_t SBO = 128;
static constexpr int TC_THREADS = 128;
static constexpr size_t SMEM_LIMIT = 227 * 1024;
__host__ __device__ static inline uint32_t align128(uint32_t v) { return (v + 127) & ~127u; }

// fp32 rows (optionally LayerNorm-ed) -> bf16.  Lanes of a warp split into groups of LPR lanes,
// one group per row, float4 per lane: global reads are coalesced and the row stays in registers
// between the statistics and the normalisation.
template <bool LN>
__device__ __forceinline__ void produce_a_f32(uint8_t* sA, const float* __restrict__ A, long long lda, long long M, long long m0,
                                              int K, int Kpad, const float* __restrict__ g, const float* __restrict__ b, float eps) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nf4 = K >> 2, nslots = Kpad >> 2;
    int LPR = 1;
    while (LPR < 32 && LPR < nslots) LPR <<= 1;
    const int RPW = 32 / LPR;
    const int gl = lane & (LPR - 1), gr = lane / LPR;
    for (int it = 0; it < 32 / RPW; it++) {
        const int r = warp * 32 + it * RPW + gr;
        const long long m = m0 + r;
        const bool rowok = m < M;
        float4 v[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            int q = gl + i * LPR;
            v[i] = (rowok && q < nf4) ? *reinterpret_cast<const float4*>(A + m * lda + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (LN) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 3; i++) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
            for (int o = LPR >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const float mean = s / (float)K;
            float ss = 0.f;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                if (gl + i * LPR < nf4) {
                    float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
                    ss += (dx * dx + dy * dy) + (dz * dz + dw * dw);
                }
            }
            for (int o = LPR >> 1; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            const float rstd = rsqrtf(ss / (float)K + eps);
#pragma unroll
            for (int i = 0; i < 3; i++) {
                int q = gl + i * LPR;
                if (rowok && q < nf4) {
                    float4 gg = __ldg(reinterpret_cast<const float4*>(g) + q), bb = __ldg(reinterpret_cast<const float4*>(b) + q);
                    v[i].x = (v[i].x - mean) * rstd * gg.x + bb.x;
                    v[i].y = (v[i].y - mean) * rstd * gg.y + bb.y;
                    v[i].z = (v[i].z - mean) * rstd * gg.z + bb.z;
                    v[i].w = (v[i].w - mean) * rstd * gg.w + bb.w;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 3; i++) {
            int q = gl + i * LPR;
            if (q < nslots) {
                uint2 pk = make_uint2(pack_bf16x2(v[i].x, v[i].y), pack_bf16x2(v[i].z, v[i].w));
                *reinterpret_cast<uint2*>(sA + (uint32_t)(q >> 1) * LBO_A + (uint32_t)r * 16 + (q & 1) * 8) = pk;
            }
        }
    }
}

// =============================================================================================
// k_tc_mlp: out = residual + W2 ELU(W1 LN(x) + b1) + b2, hidden activation kept on chip
// =============================================================================================

template <bool LN>
__global__ void __launch_bounds__(TC_THREADS) k_tc_mlp(TcMlp p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t Cpad = (uint32_t)p.Cpad, HC = (uint32_t)p.HC;
    uint8_t* sA1 = smem;
    uint8_t* sA2 = sA1 + align128((Cpad >> 3) * LBO_A);
    uint8_t* sW1 = sA2 + align128((HC >> 3) * LBO_A);
    const uint32_t w_bytes = HC * Cpad * 2u;  // both weight chunks have HC*Cpad elements
    uint8_t* sW2 = sW1 + align128(w_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW2 + align128(w_bytes));  // w1, w2, mma1, mma2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    const long long m0 = (long long)blockIdx.x * 128;
    const uint32_t d2off = (HC + 31u) & ~31u;
    const uint32_t ncols = tmem_cols_pow2(d2off + Cpad);

    if (tid == 0) {
        for (int i = 0; i < 4; i++) mbar_init(&bars[i], 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(&bars[0], w_bytes);
        bulk_g2s(sW1, p.W1p, w_bytes, &bars[0]);
        mbar_arrive_expect_tx(&bars[1], w_bytes);
        bulk_g2s(sW2, p.W2p, w_bytes, &bars[1]);
    }
    if (warp == 1) tmem_alloc(tmem_slot, ncols);
    produce_a_f32<LN>(sA1, p.x, p.C, p.M, m0, p.C, p.Cpad, p.ln_g, p.ln_b, p.eps);
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t a1 = smem_u32(sA1), a2 = smem_u32(sA2), w1 = smem_u32(sW1), w2 = smem_u32(sW2);
    const int row = warp * 32 + lane;
    const long long m = m0 + row;
    const uint32_t tlane = tmem_base + ((uint32_t)(warp * 32) << 16);

    for (int hc = 0; hc < p.n_hc; hc++) {
        const uint32_t par = (uint32_t)hc & 1u;
        if (tid == 0) {  // GEMM1: D1[128 x HC] = A1 * W1chunk^T
            mbar_wait(&bars[0], par);
            tc_fence_after_sync();
            const uint32_t idesc = make_idesc_bf16(128, HC);
            const uint32_t lbo_w = lbo_dense(HC);
            for (uint32_t ks = 0; ks < (Cpad >> 4); ks++)
                umma_bf16(tmem_base, make_smem_desc(a1 + ks * 2u * LBO_A, LBO_A, SBO), make_smem_desc(w1 + ks * 2u * lbo_w, lbo_w, SBO), idesc, ks > 0);
            umma_commit(&bars[2]);
        }
        mbar_wait(&bars[2], par);
        tc_fence_after_sync();
        if (hc > 0) {  // GEMM2 of the previous chunk must be done before sA2 / sW2 are overwritten
            mbar_wait(&bars[3], par ^ 1u);
            tc_fence_after_sync();
        }
        if (tid == 0) {
            if (hc + 1 < p.n_hc) {  // sW1 is free (GEMM1 done): prefetch the next W1 chunk
                mbar_arrive_expect_tx(&bars[0], w_bytes);
                bulk_g2s(sW1, p.W1p + (size_t)(hc + 1) * HC * Cpad, w_bytes, &bars[0]);
            }
            if (hc > 0) {  // sW2 is free: fetch this chunk's W2 (it lands while the ELU epilogue runs)
                mbar_arrive_expect_tx(&bars[1], w_bytes);
                bulk_g2s(sW2, p.W2p + (size_t)hc * HC * Cpad, w_bytes, &bars[1]);
            }
        }
        __syncwarp();
        // epilogue 1: D1 -> +b1 -> ELU -> bf16 -> sA2 (A operand of GEMM2)
        for (uint32_t c16 = 0; c16 < HC; c16 += 16) {
            float v[16];
            tmem_ld16(tlane + c16, v);
            const float* bb = p.b1 + (size_t)hc * HC + c16;
#pragma unroll
            for (int i = 0; i < 16; i++) v[i] = elu1(v[i] + __ldg(bb + i));
            uint8_t* dst = sA2 + (c16 >> 3) * LBO_A + (uint32_t)row * 16;
            *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
            *reinterpret_cast<uint4*>(dst + LBO_A) = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
        }
        fence_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        tc_fence_after_sync();
        if (tid == 0) {  // GEMM2: D2[128 x Cpad] += A2 * W2chunk^T   (N split in <=256-wide pieces)
            mbar_wait(&bars[1], par);
            tc_fence_after_sync();
            const uint32_t lbo_w = lbo_dense(Cpad);
            for (uint32_t n0 = 0; n0 < Cpad; n0 += 256) {
                const uint32_t nsz = (Cpad - n0) < 256u ? (Cpad - n0) : 256u;
                const uint32_t idesc = make_idesc_bf16(128, nsz);
                for (uint32_t ks = 0; ks < (HC >> 4); ks++)
                    umma_bf16(tmem_base + d2off + n0, make_smem_desc(a2 + ks * 2u * LBO_A, LBO_A, SBO),
                              make_smem_desc(w2 + ks * 2u * lbo_w + n0 * 16u, lbo_w, SBO), idesc, hc > 0 || ks > 0);
            }
            umma_commit(&bars[3]);
        }
    }
    mbar_wait(&bars[3], (uint32_t)(p.n_hc - 1) & 1u);
    __syncwarp();
    tc_fence_after_sync();
    // epilogue 2: D2 + b2 + residual -> out (fp32)
    for (uint32_t c16 = 0; c16 < Cpad; c16 += 16) {
        if ((int)c16 >= p.C) break;
        float v[16];
        tmem_ld16(tlane + d2off + c16, v);
        if (m < p.M) {
            float* o = p.out + m * p.C + c16;
            const float* rs = p.residual ? p.residual + m * p.C + c16 : nullptr;
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                if ((int)c16 + i + 4 <= p.C) {
                    float4 b = __ldg(reinterpret_cast<const float4*>(p.b2 + c16 + i));
                    float4 t = make_float4(v[i] + b.x, v[i + 1] + b.y, v[i + 2] + b.z, v[i + 3] + b.w);
                    if (rs) {
                        float4 rr = *reinterpret_cast<const float4*>(rs + i);
                        t.x += rr.x; t.y += rr.y; t.z += rr.z; t.w += rr.w;
                    }
                    *reinterpret_cast<float4*>(o + i) = t;
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, ncols);
}

static size_t tc_mlp_smem(int Cpad, int HC) {
    return align128((uint32_t)(Cpad / 8) * LBO_A) + align128((uint32_t)(HC / 8) * LBO_A) + 2 * (size_t)align128((uint32_t)HC * Cpad * 2) + 64;
}

int tc_mlp_pick_hc(int Cpad, int hidden) {
    int hpad = (int)pad16((uint32_t)hidden);
    for (int hc : {128, 64, 32, 16}) {
        if (hc > hpad && hc != 16) continue;
        if (tc_mlp_smem(Cpad, hc) <= SMEM_LIMIT && ((hc + 31) / 32 * 32 + Cpad) <= 512) return hc;
    }
    return 0;
}


int launch_tc_mlp(const TcMlp& t, cudaStream_t st) {
    size_t smem = tc_mlp_smem(t.Cpad, t.HC);
    static thread_local bool configured = false;
    if (!configured) {
        cudaError_t e1 = cudaFuncSetAttribute(k_tc_mlp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
        cudaError_t e2 = cudaFuncSetAttribute(k_tc_mlp<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
        if (e1 != cudaSuccess || e2 != cudaSuccess) { set_error("tc_mlp: cudaFuncSetAttribute failed"); return SF_ERR_CUDA; }
        configured = true;
    }
    long long tiles = (t.M + 127) / 128;
    SF_CHECK_ARG(tiles <= 2147483647LL, "tc_mlp: M too large");
    ProfScope ps("tc_mlp_fused", 4.0 * (double)t.M * t.C * t.hidden,
                 4.0 * (double)t.M * t.C * (t.residual && t.residual != t.x ? 3.0 : 2.0) + 4.0 * t.C * t.hidden, st);
    if (t.ln_g) k_tc_mlp<true><<<(unsigned)tiles, TC_THREADS, smem, st>>>(t);
    else k_tc_mlp<false><<<(unsigned)tiles, TC_THREADS, smem, st>>>(t);
    SF_CHECK_LAUNCH("tc_mlp");
    return SF_OK;
}

}  // namespace sf

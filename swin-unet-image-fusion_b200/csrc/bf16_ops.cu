// bf16 tensor-core path (SF_PREC_BF16): tcgen05.mma with TMEM accumulators for every linear
// layer of the hot path (QKV / output projection, the MLP, the patch 1x1 convs); LayerNorm,
// softmax, bias/mask, residual stream and all accumulators stay fp32.
//
//   k_pack_w        fp32 nn.Linear / 1x1-conv weights -> bf16 UMMA operand images (k-chunk major)
//   k_tc_gemm       [LN ->] token GEMM: A tile produced by the CTA's threads straight into the UMMA
//                   smem layout (fp32 -> LN -> bf16), weights arrive with one cp.async.bulk (TMA
//                   bulk engine) per CTA, one thread issues tcgen05.mma, 4 warps drain TMEM with
//                   tcgen05.ld and apply bias / residual
//   k_tc_mlp        fused MLP: LN -> GEMM1 -> +b1 -> ELU -> (bf16, back to smem as the next A
//                   operand) -> GEMM2 accumulated over hidden chunks in TMEM -> +b2 + residual.
//                   The 4x hidden activation never touches HBM.
//   k_attn_core_bf16  per-window attention core on CUDA cores (head_dim 3..48: the work is
//                   dominated by the 49x49 softmax, not by the two tiny GEMMs), bf16 I/O, fp32 math.
#include <initializer_list>
#include "bf16_kernels.cuh"
#include "fp32_kernels.cuh"
#include "tc_common.cuh"

namespace sf {
using namespace tc;
using bf16 = __nv_bfloat16;

static constexpr uint32_t LBO_A = lbo_padded(128);  // 2064 B between k-chunks of a 128-row A tile
static constexpr uint32_t SBO = 128;                // 8 rows x 16 B
static constexpr int TC_THREADS = 128;
static constexpr int MAX_KPAD = 384;
static constexpr size_t SMEM_LIMIT = 227 * 1024;

__host__ __device__ static inline uint32_t align128(uint32_t v) { return (v + 127) & ~127u; }

// =============================================================================================
// weight packing
// =============================================================================================
// Stacks up to 3 [Neach x K] fp32 matrices along N, splits rows into n_chunks of NR and columns
// into k_chunks of KR, and writes for every (jn, jk) a bf16 image img[kc][r][8]
// (= W[jn*NR + r][jk*KR + kc*8 + e], zero outside), images ordered jn-major.
struct PackSrc { const float* w[3]; const float* b[3]; };

__global__ void k_pack_w(PackSrc src, int nsrc, int Neach, int K, bf16* __restrict__ out, float* __restrict__ bias_out,
                         int NR, int KR, int n_chunks, int k_chunks) {
    const int Ntot = nsrc * Neach;
    const long long cpi = (long long)NR * KR / 8;  // 16-byte chunks per image
    const long long total = cpi * n_chunks * k_chunks;
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < total; c += (long long)gridDim.x * blockDim.x) {
        long long img = c / cpi;
        int ci = (int)(c - img * cpi);
        int jk = (int)(img % k_chunks), jn = (int)(img / k_chunks);
        int kc = ci / NR, r = ci - kc * NR;
        int n = jn * NR + r;
        uint32_t pk[4] = {0, 0, 0, 0};
        if (n < Ntot) {
            int s = n / Neach;
            const float* row = src.w[s] + (long long)(n - s * Neach) * K;
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {
                int k = jk * KR + kc * 8 + e;
                v[e] = k < K ? row[k] : 0.f;
            }
#pragma unroll
            for (int e = 0; e < 4; e++) pk[e] = pack_bf16x2(v[2 * e], v[2 * e + 1]);
        }
        *reinterpret_cast<uint4*>(out + c * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
    if (bias_out) {
        for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < n_chunks * NR; n += gridDim.x * blockDim.x) {
            float b = 0.f;
            if (n < Ntot) {
                int s = n / Neach;
                if (src.b[s]) b = src.b[s][n - s * Neach];
            }
            bias_out[n] = b;
        }
    }
}

static int launch_pack(const PackSrc& src, int nsrc, int Neach, int K, bf16* out, float* bias_out, int NR, int KR,
                       int n_chunks, int k_chunks, cudaStream_t st) {
    long long total = (long long)NR * KR / 8 * n_chunks * k_chunks;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    ProfScope ps("pack_weights_bf16", 0.0, 6.0 * (double)nsrc * Neach * K, st);
    k_pack_w<<<blocks, 256, 0, st>>>(src, nsrc, Neach, K, out, bias_out, NR, KR, n_chunks, k_chunks);
    SF_CHECK_LAUNCH("pack_weights_bf16");
    return SF_OK;
}

// =============================================================================================
// A-tile producers: write a [128 x Kpad] bf16 operand into smem in the UMMA layout
// =============================================================================================
enum { AM_F32 = 0, AM_F32_LN = 1, AM_BF16 = 2, AM_MERGE = 3 };

struct TcGemm {
    const void* A;            // fp32 (AM_F32*, AM_MERGE) or bf16 (AM_BF16)
    long long M;
    int K, Kpad;
    long long lda;            // elements
    const float* ln_g; const float* ln_b; float eps;
    const bf16* Wp; int NCH;  // packed weights, one image [Kpad/8][NCH][8] per blockIdx.y
    const float* bias;        // [n_chunks*NCH] (packed, zero padded) or null
    const float* residual; long long ldr;
    void* out; long long ldo; int out_col0; int N;
    int Hf, Wf, Cin, mh, mw;  // AM_MERGE: fine map (B,Hf,Wf,Cin), merging factors
};

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

// bf16 rows: 16-byte cp.async per k-chunk (global chunk == smem chunk)
__device__ __forceinline__ void produce_a_bf16(uint8_t* sA, const bf16* __restrict__ A, long long lda, long long M, long long m0, int K,
                                               int Kpad) {
    const int nkc = Kpad >> 3;
    for (int idx = threadIdx.x; idx < 128 * nkc; idx += blockDim.x) {
        int r = idx / nkc, kc = idx - r * nkc;
        uint8_t* dst = sA + (uint32_t)kc * LBO_A + (uint32_t)r * 16;
        long long m = m0 + r;
        if (m < M && kc * 8 + 8 <= K) cp_async16(dst, A + m * lda + kc * 8);
        else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
    cp_async_wait_all();
}

// patch-merge gather (a011:87-93): row (b,Y,X), k = (ph*mw+pw)*Cin + c <- in[b][Y*mh+ph][X*mw+pw][c]
__device__ __forceinline__ void produce_a_merge(uint8_t* sA, const TcGemm& p, long long m0) {
    const float* __restrict__ in = reinterpret_cast<const float*>(p.A);
    const int nkc = p.Kpad >> 3;
    const int Hc = p.Hf / p.mh, Wc = p.Wf / p.mw;
    for (int idx = threadIdx.x; idx < 128 * nkc; idx += blockDim.x) {
        int kc = idx >> 7, r = idx & 127;  // consecutive threads -> consecutive rows (conflict-free st.shared)
        long long m = m0 + r;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; e++) v[e] = 0.f;
        if (m < p.M) {
            int X = (int)(m % Wc);
            long long t = m / Wc;
            int Y = (int)(t % Hc);
            long long b = t / Hc;
#pragma unroll
            for (int e = 0; e < 8; e++) {
                int k = kc * 8 + e;
                if (k < p.K) {
                    int q = k / p.Cin, c = k - q * p.Cin;
                    int ph = q / p.mw, pw = q - ph * p.mw;
                    v[e] = in[((b * p.Hf + (Y * p.mh + ph)) * p.Wf + (X * p.mw + pw)) * p.Cin + c];
                }
            }
        }
        *reinterpret_cast<uint4*>(sA + (uint32_t)kc * LBO_A + (uint32_t)r * 16) =
            make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
}

// =============================================================================================
// k_tc_gemm: out[m0:m0+128, chunk] = A_tile * Wchunk^T (+bias)(+residual)
// =============================================================================================
template <int AMODE, bool OUT_BF16>
__global__ void __launch_bounds__(TC_THREADS) k_tc_gemm(TcGemm p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t nkc = (uint32_t)p.Kpad >> 3;
    uint8_t* sA = smem;
    uint8_t* sW = smem + align128(nkc * LBO_A);
    const uint32_t w_bytes = (uint32_t)p.NCH * (uint32_t)p.Kpad * 2u;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW + align128(w_bytes));  // [0] weights landed, [1] MMAs done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    const long long m0 = (long long)blockIdx.x * 128;
    const int chunk = blockIdx.y;
    const uint32_t ncols = tmem_cols_pow2((uint32_t)p.NCH);

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(&bars[0], w_bytes);
        bulk_g2s(sW, p.Wp + (size_t)chunk * p.NCH * p.Kpad, w_bytes, &bars[0]);
    }
    if (warp == 1) tmem_alloc(tmem_slot, ncols);

    if (AMODE == AM_F32) produce_a_f32<false>(sA, reinterpret_cast<const float*>(p.A), p.lda, p.M, m0, p.K, p.Kpad, nullptr, nullptr, 0.f);
    else if (AMODE == AM_F32_LN) produce_a_f32<true>(sA, reinterpret_cast<const float*>(p.A), p.lda, p.M, m0, p.K, p.Kpad, p.ln_g, p.ln_b, p.eps);
    else if (AMODE == AM_BF16) produce_a_bf16(sA, reinterpret_cast<const bf16*>(p.A), p.lda, p.M, m0, p.K, p.Kpad);
    else produce_a_merge(sA, p, m0);

    fence_async_smem();       // generic-proxy smem writes -> visible to the tensor-core (async) proxy
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (tid == 0) {
        mbar_wait(&bars[0], 0);
        tc_fence_after_sync();
        const uint32_t idesc = make_idesc_bf16(128, (uint32_t)p.NCH);
        const uint32_t a0 = smem_u32(sA), w0 = smem_u32(sW);
        const uint32_t lbo_w = lbo_dense((uint32_t)p.NCH);
        const int ksteps = p.Kpad >> 4;
        for (int ks = 0; ks < ksteps; ks++) {
            uint64_t da = make_smem_desc(a0 + (uint32_t)ks * 2u * LBO_A, LBO_A, SBO);
            uint64_t db = make_smem_desc(w0 + (uint32_t)ks * 2u * lbo_w, lbo_w, SBO);
            umma_bf16(tmem_base, da, db, idesc, ks > 0);
        }
        umma_commit(&bars[1]);
    }
    mbar_wait(&bars[1], 0);
    __syncwarp();             // tcgen05.ld is .sync.aligned: the warp must be converged
    tc_fence_after_sync();

    // ---- epilogue: thread <-> row (TMEM lane), 16 columns per tcgen05.ld ----------------------------
    const int row = warp * 32 + lane;
    const long long m = m0 + row;
    const uint32_t tlane = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int ncol0 = chunk * p.NCH;
    for (int c16 = 0; c16 < p.NCH; c16 += 16) {
        if (ncol0 + c16 >= p.N) break;  // uniform across the CTA
        float v[16];
        tmem_ld16(tlane + (uint32_t)c16, v);
        if (m < p.M) {
            const int n0 = ncol0 + c16;
            if (p.bias) {
#pragma unroll
                for (int i = 0; i < 16; i++) v[i] += __ldg(p.bias + n0 + i);
            }
            if (OUT_BF16) {
                bf16* o = reinterpret_cast<bf16*>(p.out) + m * p.ldo + p.out_col0 + n0;
                if (n0 + 16 <= p.N && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
                    uint4 a = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
                    uint4 c = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
                    reinterpret_cast<uint4*>(o)[0] = a;
                    reinterpret_cast<uint4*>(o)[1] = c;
                } else {
#pragma unroll
                    for (int i = 0; i < 16; i++) if (n0 + i < p.N) o[i] = __float2bfloat16_rn(v[i]);
                }
            } else {
                float* o = reinterpret_cast<float*>(p.out) + m * p.ldo + p.out_col0 + n0;
                const float* rs = p.residual ? p.residual + m * p.ldr + n0 : nullptr;
                const bool vec = ((reinterpret_cast<uintptr_t>(o) & 15) == 0) && (!rs || (reinterpret_cast<uintptr_t>(rs) & 15) == 0);
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    if (vec && n0 + i + 4 <= p.N) {
                        float4 t = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        if (rs) {
                            float4 rr = *reinterpret_cast<const float4*>(rs + i);
                            t.x += rr.x; t.y += rr.y; t.z += rr.z; t.w += rr.w;
                        }
                        *reinterpret_cast<float4*>(o + i) = t;
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; e++)
                            if (n0 + i + e < p.N) o[i + e] = v[i + e] + (rs ? rs[i + e] : 0.f);
                    }
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, ncols);
}

static size_t tc_gemm_smem(int Kpad, int NCH) {
    return align128((uint32_t)(Kpad / 8) * LBO_A) + align128((uint32_t)NCH * Kpad * 2) + 64;
}

// chunk width along N: as wide as TMEM / smem allow, balanced over the chunks
static void pick_nchunk(int Ntot, int Kpad, int* NCH, int* n_chunks) {
    int cap = Kpad > 192 ? 128 : 256;
    int npad = (int)pad16((uint32_t)Ntot);
    int nc = (npad + cap - 1) / cap;
    *NCH = (int)pad16((uint32_t)((npad + nc - 1) / nc));
    *n_chunks = nc;
}

template <int AMODE, bool OUT_BF16>
static int launch_tc_gemm_t(const TcGemm& p, int n_chunks, cudaStream_t st) {
    size_t smem = tc_gemm_smem(p.Kpad, p.NCH);
    SF_CHECK_ARG(smem <= SMEM_LIMIT, "tc_gemm: %zu B of shared memory needed (K=%d, chunk=%d)", smem, p.K, p.NCH);
    static thread_local size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(k_tc_gemm<AMODE, OUT_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
        if (e != cudaSuccess) { set_error("tc_gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SF_ERR_CUDA; }
        configured = SMEM_LIMIT;
    }
    long long tiles = (p.M + 127) / 128;
    SF_CHECK_ARG(tiles <= 2147483647LL, "tc_gemm: M too large");
    dim3 grid((unsigned)tiles, (unsigned)n_chunks);
    const double abytes = (AMODE == AM_BF16 ? 2.0 : 4.0) * (double)p.M * p.K;
    const double obytes = (OUT_BF16 ? 2.0 : 4.0) * (double)p.M * p.N * (p.residual ? 2.0 : 1.0);
    ProfScope ps(AMODE == AM_F32_LN ? "tc_gemm_ln" : (AMODE == AM_BF16 ? "tc_gemm_bf16in" : (AMODE == AM_MERGE ? "tc_gemm_merge" : "tc_gemm_f32in")),
                 2.0 * (double)p.M * p.N * p.K, abytes + obytes + 2.0 * p.N * p.K, st);
    k_tc_gemm<AMODE, OUT_BF16><<<grid, TC_THREADS, smem, st>>>(p);
    SF_CHECK_LAUNCH("tc_gemm");
    return SF_OK;
}

static int launch_tc_gemm(int amode, bool out_bf16, const TcGemm& p, int n_chunks, cudaStream_t st) {
    SF_CHECK_ARG(p.Kpad <= MAX_KPAD && p.Kpad % 16 == 0 && p.NCH % 16 == 0 && p.NCH <= 256, "tc_gemm: unsupported tile (Kpad=%d, NCH=%d)", p.Kpad, p.NCH);
    if (amode == AM_F32_LN && !out_bf16) return launch_tc_gemm_t<AM_F32_LN, false>(p, n_chunks, st);
    if (amode == AM_F32_LN && out_bf16) return launch_tc_gemm_t<AM_F32_LN, true>(p, n_chunks, st);
    if (amode == AM_F32 && !out_bf16) return launch_tc_gemm_t<AM_F32, false>(p, n_chunks, st);
    if (amode == AM_F32 && out_bf16) return launch_tc_gemm_t<AM_F32, true>(p, n_chunks, st);
    if (amode == AM_BF16 && !out_bf16) return launch_tc_gemm_t<AM_BF16, false>(p, n_chunks, st);
    if (amode == AM_MERGE && !out_bf16) return launch_tc_gemm_t<AM_MERGE, false>(p, n_chunks, st);
    set_error("tc_gemm: unsupported mode combination");
    return SF_ERR_INVALID;
}

// =============================================================================================
// k_tc_mlp: out = residual + W2 ELU(W1 LN(x) + b1) + b2, hidden activation kept on chip
// =============================================================================================
struct TcMlp {
    const float* x; const float* residual; float* out;
    long long M;
    int C, Cpad, hidden, HC, n_hc;
    const float* ln_g; const float* ln_b; float eps;
    const bf16* W1p;   // n_hc images [Cpad/8][HC][8]      (rows = hidden chunk)
    const bf16* W2p;   // n_hc images [HC/8][Cpad][8]      (rows = output channel, k = hidden chunk)
    const float* b1;   // [n_hc*HC] zero padded
    const float* b2;   // [C]
};

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

static int pick_hc(int Cpad, int hidden) {
    int hpad = (int)pad16((uint32_t)hidden);
    for (int hc : {128, 64, 32, 16}) {
        if (hc > hpad && hc != 16) continue;
        if (tc_mlp_smem(Cpad, hc) <= SMEM_LIMIT && ((hc + 31) / 32 * 32 + Cpad) <= 512) return hc;
    }
    return 0;
}

// =============================================================================================
// attention core, bf16 I/O (a001:317-354), one CTA per window, all heads
// =============================================================================================
// REGS: T <= 64 and small head_dim: the T scores of a row stay in registers between the max
//       and the exp/PV passes and the (T x T) bias matrix is staged in smem.
template <int DMAX, bool REGS>
__global__ void __launch_bounds__(256) k_attn_core_bf16(const bf16* __restrict__ qkv, long long ld, int koff, int voff,
                                                        bf16* __restrict__ O, long long ldo, const float* __restrict__ table,
                                                        WinGeom g, int nh, int d, float scale) {
    extern __shared__ __align__(16) float smf[];
    const int T = g.T, inner = nh * d;
    float* Ks = smf;                              // [T][inner]
    float* Vs = Ks + (size_t)T * inner;           // [T][inner]
    float* bias = Vs + (size_t)T * inner;         // REGS: [T][T] ; else the raw table
    const int tw = 2 * g.wsw - 1, tabn = (2 * g.wsh - 1) * tw;
    long long* rows = reinterpret_cast<long long*>(bias + (REGS ? ((T * T + 1) & ~1) : ((tabn + 1) & ~1)));  // 8-byte aligned
    int* regs = reinterpret_cast<int*>(rows + T);
    __shared__ int s_has_mask;
    const int win = blockIdx.x;

    if (threadIdx.x == 0) s_has_mask = 0;
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        int rg;
        rows[t] = win_token_src(g, win, t, &rg);
        regs[t] = rg;
        if (rg != 0) s_has_mask = 1;  // any region other than 0 -> the window straddles a shift boundary
    }
    if (REGS) {
        for (int i = threadIdx.x; i < T * T; i += blockDim.x) {
            int qi = i / T, kj = i - qi * T;
            bias[i] = table[(kj / g.wsw - qi / g.wsw + g.wsh - 1) * tw + (kj % g.wsw - qi % g.wsw + g.wsw - 1)];
        }
    } else {
        for (int i = threadIdx.x; i < tabn; i += blockDim.x) bias[i] = table[i];
    }
    __syncthreads();
    // stage K and V of the whole window (all heads) as fp32
    if ((inner & 7) == 0 && (ld & 7) == 0 && (koff & 7) == 0 && (voff & 7) == 0) {
        const int nch = inner >> 3;
        for (int i = threadIdx.x; i < T * nch; i += blockDim.x) {
            int t = i / nch, c = i - t * nch;
            const bf16* src = qkv + rows[t] * ld + c * 8;
            uint4 kk = *reinterpret_cast<const uint4*>(src + koff), vv = *reinterpret_cast<const uint4*>(src + voff);
            const __nv_bfloat162* k2 = reinterpret_cast<const __nv_bfloat162*>(&kk);
            const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&vv);
#pragma unroll
            for (int e = 0; e < 4; e++) {
                float2 a = __bfloat1622float2(k2[e]), b = __bfloat1622float2(v2[e]);
                Ks[t * inner + c * 8 + 2 * e] = a.x; Ks[t * inner + c * 8 + 2 * e + 1] = a.y;
                Vs[t * inner + c * 8 + 2 * e] = b.x; Vs[t * inner + c * 8 + 2 * e + 1] = b.y;
            }
        }
    } else {
        for (int i = threadIdx.x; i < T * inner; i += blockDim.x) {
            int t = i / inner, c = i - t * inner;
            Ks[i] = __bfloat162float(qkv[rows[t] * ld + koff + c]);
            Vs[i] = __bfloat162float(qkv[rows[t] * ld + voff + c]);
        }
    }
    __syncthreads();
    const bool has_mask = s_has_mask != 0;

    for (int item = threadIdx.x; item < nh * T; item += blockDim.x) {
        const int head = item / T, qi = item - head * T;
        const int hoff = head * d;
        float q[DMAX];
        const bf16* qp = qkv + rows[qi] * ld + hoff;
#pragma unroll
        for (int dd = 0; dd < DMAX; dd++) q[dd] = dd < d ? __bfloat162float(qp[dd]) * scale : 0.f;
        const int qreg = regs[qi];
        float acc[DMAX];
#pragma unroll
        for (int dd = 0; dd < DMAX; dd++) acc[dd] = 0.f;
        float mx = -INFINITY, sum = 0.f;
        if (REGS) {
            float s[64];
#pragma unroll
            for (int j = 0; j < 64; j++) {
                if (j < T) {
                    float a = 0.f;
#pragma unroll
                    for (int dd = 0; dd < DMAX; dd++) a = fmaf(q[dd], Ks[j * inner + hoff + (dd < d ? dd : 0)], a);
                    a += bias[qi * T + j];
                    if (has_mask && regs[j] != qreg) a = -1e10f;
                    s[j] = a;
                    mx = fmaxf(mx, a);
                }
            }
#pragma unroll
            for (int j = 0; j < 64; j++) {
                if (j < T) {
                    float pj = __expf(s[j] - mx);
                    sum += pj;
#pragma unroll
                    for (int dd = 0; dd < DMAX; dd++) acc[dd] = fmaf(pj, Vs[j * inner + hoff + (dd < d ? dd : 0)], acc[dd]);
                }
            }
        } else {
            const int qr = qi / g.wsw, qc = qi - qr * g.wsw;
            for (int j = 0; j < T; j++) {
                float a = 0.f;
#pragma unroll
                for (int dd = 0; dd < DMAX; dd++) a = fmaf(q[dd], Ks[j * inner + hoff + (dd < d ? dd : 0)], a);
                int jr = j / g.wsw, jc = j - jr * g.wsw;
                a += bias[(jr - qr + g.wsh - 1) * tw + (jc - qc + g.wsw - 1)];
                if (has_mask && regs[j] != qreg) a = -1e10f;
                mx = fmaxf(mx, a);
            }
            for (int j = 0; j < T; j++) {
                float a = 0.f;
#pragma unroll
                for (int dd = 0; dd < DMAX; dd++) a = fmaf(q[dd], Ks[j * inner + hoff + (dd < d ? dd : 0)], a);
                int jr = j / g.wsw, jc = j - jr * g.wsw;
                a += bias[(jr - qr + g.wsh - 1) * tw + (jc - qc + g.wsw - 1)];
                if (has_mask && regs[j] != qreg) a = -1e10f;
                float pj = __expf(a - mx);
                sum += pj;
#pragma unroll
                for (int dd = 0; dd < DMAX; dd++) acc[dd] = fmaf(pj, Vs[j * inner + hoff + (dd < d ? dd : 0)], acc[dd]);
            }
        }
        const float inv = 1.f / sum;
        bf16* op = O + rows[qi] * ldo + hoff;
#pragma unroll
        for (int dd = 0; dd < DMAX; dd++)
            if (dd < d) op[dd] = __float2bfloat16_rn(acc[dd] * inv);
    }
}

static size_t attn_bf16_smem(const WinGeom& g, int inner, bool regs_variant) {
    int tabn = (2 * g.wsh - 1) * (2 * g.wsw - 1);
    size_t f = (size_t)2 * g.T * inner + (regs_variant ? (size_t)((g.T * g.T + 1) & ~1) : (size_t)((tabn + 1) & ~1));
    return f * sizeof(float) + (size_t)g.T * (sizeof(long long) + sizeof(int)) + 16;
}

template <int DMAX, bool REGS>
static int launch_attn_bf16_t(const bf16* qkv, long long ld, int koff, int voff, bf16* O, long long ldo, const float* table,
                              const WinGeom& g, int nh, int d, cudaStream_t st) {
    const int inner = nh * d;
    size_t smem = attn_bf16_smem(g, inner, REGS);
    SF_CHECK_ARG(smem <= SMEM_LIMIT, "attention core: window of %d tokens x %d channels needs %zu B of shared memory", g.T, inner, smem);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_attn_core_bf16<DMAX, REGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("attention core: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SF_ERR_CUDA; }
    }
    long long nwin = (long long)g.B * g.nWh * g.nWw;
    SF_CHECK_ARG(nwin <= 2147483647LL, "attention core: too many windows");
    int items = nh * g.T;
    int threads = items >= 256 ? 256 : ((items + 31) / 32 * 32);
    const double mtok = (double)nwin * g.T;
    ProfScope ps("attn_core_bf16", 4.0 * g.T * mtok * inner, 8.0 * mtok * inner, st);
    k_attn_core_bf16<DMAX, REGS><<<(unsigned)nwin, threads, smem, st>>>(qkv, ld, koff, voff, O, ldo, table, g, nh, d, 1.0f / sqrtf((float)d));
    SF_CHECK_LAUNCH("attn_core_bf16");
    return SF_OK;
}

static int launch_attn_bf16(const bf16* qkv, long long ld, int koff, int voff, bf16* O, long long ldo, const float* table,
                            const WinGeom& g, int nh, int d, cudaStream_t st) {
    const bool small = g.T <= 64;
    if (small && d <= 4) return launch_attn_bf16_t<4, true>(qkv, ld, koff, voff, O, ldo, table, g, nh, d, st);
    if (small && d <= 8) return launch_attn_bf16_t<8, true>(qkv, ld, koff, voff, O, ldo, table, g, nh, d, st);
    if (small && d <= 16) return launch_attn_bf16_t<16, true>(qkv, ld, koff, voff, O, ldo, table, g, nh, d, st);
    if (d <= 16) return launch_attn_bf16_t<16, false>(qkv, ld, koff, voff, O, ldo, table, g, nh, d, st);
    if (d <= 32) return launch_attn_bf16_t<32, false>(qkv, ld, koff, voff, O, ldo, table, g, nh, d, st);
    if (d <= 64) return launch_attn_bf16_t<64, false>(qkv, ld, koff, voff, O, ldo, table, g, nh, d, st);
    set_error("attention core: head_dim %d > 64 is not supported", d);
    return SF_ERR_UNSUPPORTED;
}

// =============================================================================================
// operator-level host code
// =============================================================================================
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int wa_check_bf16(const sf_window_attn_params* p) {
    const int inner = p->num_heads * p->head_dim;
    if (p->C % 4 != 0 || (int)pad16(p->C) > MAX_KPAD || inner % 8 != 0 || (int)pad16(inner) > MAX_KPAD ||
        !aligned16(p->q_src) || !aligned16(p->kv_src)) {
        set_error("bf16 window attention supports C %% 4 == 0, heads*head_dim %% 8 == 0, both <= %d and 16-byte aligned maps "
                  "(got C=%d, inner=%d)", MAX_KPAD, p->C, inner);
        return SF_ERR_UNSUPPORTED;
    }
    return SF_OK;
}

struct WaPlan {
    int inner, Kpad, KpadO;
    int nch_q, nc_q;     // self: stacked q,k,v ; cross: q only
    int nch_kv, nc_kv;   // cross: stacked k,v
    int nch_o, nc_o;
    size_t off_qkv, off_o, off_wq, off_wkv, off_wo, off_bq, off_bkv, off_bo, total;
};

static WaPlan wa_plan(const sf_window_attn_params* p) {
    WaPlan w{};
    const size_t M = (size_t)p->B * p->Hp * p->Wp;
    const bool self_attn = p->kv_src == p->q_src && p->ln_q_gamma == p->ln_kv_gamma && p->ln_q_beta == p->ln_kv_beta;
    w.inner = p->num_heads * p->head_dim;
    w.Kpad = (int)pad16(p->C);
    w.KpadO = (int)pad16(w.inner);
    pick_nchunk(self_attn ? 3 * w.inner : w.inner, w.Kpad, &w.nch_q, &w.nc_q);
    if (!self_attn) pick_nchunk(2 * w.inner, w.Kpad, &w.nch_kv, &w.nc_kv);
    pick_nchunk(p->C, w.KpadO, &w.nch_o, &w.nc_o);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
    w.off_qkv = take(M * 3 * w.inner * sizeof(bf16));
    w.off_o = take(M * w.inner * sizeof(bf16));
    w.off_wq = take((size_t)w.nc_q * w.nch_q * w.Kpad * sizeof(bf16));
    w.off_wkv = take((size_t)w.nc_kv * w.nch_kv * w.Kpad * sizeof(bf16));
    w.off_wo = take((size_t)w.nc_o * w.nch_o * w.KpadO * sizeof(bf16));
    w.off_bq = take((size_t)w.nc_q * w.nch_q * sizeof(float));
    w.off_bkv = take((size_t)w.nc_kv * w.nch_kv * sizeof(float) + 16);
    w.off_bo = take((size_t)w.nc_o * w.nch_o * sizeof(float));
    w.total = off;
    return w;
}

size_t window_attn_ws_bf16(const sf_window_attn_params* p) { return wa_plan(p).total; }

int window_attn_fwd_bf16(const sf_window_attn_params* p, void* ws_ptr, size_t ws_bytes, cudaStream_t st) {
    SF_TRY(wa_check_bf16(p));
    const WaPlan w = wa_plan(p);
    if (ws_bytes < w.total || !ws_ptr) { set_error("sf_window_attn_fwd: workspace too small (%zu B given, %zu needed)", ws_bytes, w.total); return SF_ERR_WORKSPACE; }
    char* base = reinterpret_cast<char*>(ws_ptr);
    const long long M = (long long)p->B * p->Hp * p->Wp;
    const int inner = w.inner, C = p->C;
    const bool self_attn = w.nc_kv == 0;
    bf16* qkv = reinterpret_cast<bf16*>(base + w.off_qkv);
    bf16* O = reinterpret_cast<bf16*>(base + w.off_o);
    bf16* wq = reinterpret_cast<bf16*>(base + w.off_wq);
    bf16* wkv = reinterpret_cast<bf16*>(base + w.off_wkv);
    bf16* wo = reinterpret_cast<bf16*>(base + w.off_wo);
    float* bq = reinterpret_cast<float*>(base + w.off_bq);
    float* bkv = reinterpret_cast<float*>(base + w.off_bkv);
    float* bo = reinterpret_cast<float*>(base + w.off_bo);

    // 1. pack weights (bf16 UMMA images) + stacked biases
    if (self_attn) {
        PackSrc s{{p->wq, p->wk, p->wv}, {p->bq, p->bk, p->bv}};
        SF_TRY(launch_pack(s, 3, inner, C, wq, bq, w.nch_q, w.Kpad, w.nc_q, 1, st));
    } else {
        PackSrc s1{{p->wq, nullptr, nullptr}, {p->bq, nullptr, nullptr}};
        SF_TRY(launch_pack(s1, 1, inner, C, wq, bq, w.nch_q, w.Kpad, w.nc_q, 1, st));
        PackSrc s2{{p->wk, p->wv, nullptr}, {p->bk, p->bv, nullptr}};
        SF_TRY(launch_pack(s2, 2, inner, C, wkv, bkv, w.nch_kv, w.Kpad, w.nc_kv, 1, st));
    }
    PackSrc so{{p->wo, nullptr, nullptr}, {p->bo, nullptr, nullptr}};
    SF_TRY(launch_pack(so, 1, C, inner, wo, bo, w.nch_o, w.KpadO, w.nc_o, 1, st));

    // 2. projections -> qkv [M][3*inner] bf16
    TcGemm g{};
    g.M = M; g.K = C; g.Kpad = w.Kpad; g.lda = C; g.eps = p->ln_eps;
    g.out = qkv; g.ldo = 3 * inner;
    g.A = p->q_src; g.ln_g = p->ln_q_gamma; g.ln_b = p->ln_q_beta;
    g.Wp = wq; g.NCH = w.nch_q; g.bias = bq; g.out_col0 = 0; g.N = self_attn ? 3 * inner : inner;
    SF_TRY(launch_tc_gemm(p->ln_q_gamma ? AM_F32_LN : AM_F32, true, g, w.nc_q, st));
    if (!self_attn) {
        g.A = p->kv_src; g.ln_g = p->ln_kv_gamma; g.ln_b = p->ln_kv_beta;
        g.Wp = wkv; g.NCH = w.nch_kv; g.bias = bkv; g.out_col0 = inner; g.N = 2 * inner;
        SF_TRY(launch_tc_gemm(p->ln_kv_gamma ? AM_F32_LN : AM_F32, true, g, w.nc_kv, st));
    }
    // 3. attention core -> O [M][inner] bf16
    WinGeom geom = make_geom(p->B, p->Hp, p->Wp, p->wsh, p->wsw, p->shift);
    SF_TRY(launch_attn_bf16(qkv, 3 * inner, inner, 2 * inner, O, inner, p->bias_table, geom, p->num_heads, p->head_dim, st));
    // 4. output projection (+ residual) -> out fp32
    TcGemm o{};
    o.M = M; o.K = inner; o.Kpad = w.KpadO; o.lda = inner; o.A = O;
    o.Wp = wo; o.NCH = w.nch_o; o.bias = bo; o.residual = p->residual; o.ldr = C;
    o.out = p->out; o.ldo = C; o.out_col0 = 0; o.N = C;
    SF_TRY(launch_tc_gemm(AM_BF16, false, o, w.nc_o, st));
    return SF_OK;
}

// ---- MLP -------------------------------------------------------------------------------------------
struct MlpPlan { int Cpad, HC, n_hc; size_t off_w1, off_w2, off_b1, total; };

static MlpPlan mlp_plan(const sf_mlp_params* p) {
    MlpPlan m{};
    m.Cpad = (int)pad16(p->C);
    m.HC = m.Cpad <= MAX_KPAD ? pick_hc(m.Cpad, p->hidden) : 0;
    if (m.HC == 0) return m;
    m.n_hc = ((int)pad16(p->hidden) + m.HC - 1) / m.HC;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
    m.off_w1 = take((size_t)m.n_hc * m.HC * m.Cpad * sizeof(bf16));
    m.off_w2 = take((size_t)m.n_hc * m.HC * m.Cpad * sizeof(bf16));
    m.off_b1 = take((size_t)m.n_hc * m.HC * sizeof(float));
    m.total = off;
    return m;
}

size_t mlp_ws_bf16(const sf_mlp_params* p) { return mlp_plan(p).total; }

int mlp_fwd_bf16(const sf_mlp_params* p, void* ws_ptr, size_t ws_bytes, cudaStream_t st) {
    const MlpPlan m = mlp_plan(p);
    if (m.HC == 0 || p->C % 4 != 0 || !aligned16(p->in) || !aligned16(p->out) || (p->residual && !aligned16(p->residual))) {
        set_error("bf16 MLP supports C %% 4 == 0, C <= %d and 16-byte aligned maps (got C=%d, hidden=%d)", MAX_KPAD, p->C, p->hidden);
        return SF_ERR_UNSUPPORTED;
    }
    if (ws_bytes < m.total || !ws_ptr) { set_error("sf_mlp_fwd: workspace too small (%zu B given, %zu needed)", ws_bytes, m.total); return SF_ERR_WORKSPACE; }
    char* base = reinterpret_cast<char*>(ws_ptr);
    bf16* w1 = reinterpret_cast<bf16*>(base + m.off_w1);
    bf16* w2 = reinterpret_cast<bf16*>(base + m.off_w2);
    float* b1 = reinterpret_cast<float*>(base + m.off_b1);
    PackSrc s1{{p->w1, nullptr, nullptr}, {p->b1, nullptr, nullptr}};
    SF_TRY(launch_pack(s1, 1, p->hidden, p->C, w1, b1, m.HC, m.Cpad, m.n_hc, 1, st));       // rows = hidden chunks
    PackSrc s2{{p->w2, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    SF_TRY(launch_pack(s2, 1, p->C, p->hidden, w2, nullptr, m.Cpad, m.HC, 1, m.n_hc, st));   // k = hidden chunks
    TcMlp t{};
    t.x = p->in; t.residual = p->residual; t.out = p->out; t.M = p->M;
    t.C = p->C; t.Cpad = m.Cpad; t.hidden = p->hidden; t.HC = m.HC; t.n_hc = m.n_hc;
    t.ln_g = p->ln_gamma; t.ln_b = p->ln_beta; t.eps = p->ln_eps;
    t.W1p = w1; t.W2p = w2; t.b1 = b1; t.b2 = p->b2;
    size_t smem = tc_mlp_smem(m.Cpad, m.HC);
    static thread_local bool configured = false;
    if (!configured) {
        cudaError_t e1 = cudaFuncSetAttribute(k_tc_mlp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
        cudaError_t e2 = cudaFuncSetAttribute(k_tc_mlp<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
        if (e1 != cudaSuccess || e2 != cudaSuccess) { set_error("tc_mlp: cudaFuncSetAttribute failed"); return SF_ERR_CUDA; }
        configured = true;
    }
    long long tiles = (p->M + 127) / 128;
    SF_CHECK_ARG(tiles <= 2147483647LL, "tc_mlp: M too large");
    ProfScope ps("tc_mlp_fused", 4.0 * (double)p->M * p->C * p->hidden,
                 4.0 * (double)p->M * p->C * (p->residual && p->residual != p->in ? 3.0 : 2.0) + 4.0 * p->C * p->hidden, st);
    if (p->ln_gamma) k_tc_mlp<true><<<(unsigned)tiles, TC_THREADS, smem, st>>>(t);
    else k_tc_mlp<false><<<(unsigned)tiles, TC_THREADS, smem, st>>>(t);
    SF_CHECK_LAUNCH("tc_mlp");
    return SF_OK;
}

// ---- patch layers -------------------------------------------------------------------------------------
struct PatchPlan { bool tc; int K, N, Kpad, nch, nc; long long Mrows; size_t off_lin, off_w, off_b, off_f32, total; };

static PatchPlan patch_plan(const sf_patch_params* p) {
    PatchPlan q{};
    const int mm = p->mh * p->mw;
    if (p->encoder) { q.K = mm * p->Cin; q.N = p->Cout; q.Mrows = (long long)p->B * (p->H / p->mh) * (p->W / p->mw); }
    else { q.K = p->Cin; q.N = mm * p->Cout; q.Mrows = (long long)p->B * p->H * p->W; }
    q.Kpad = (int)pad16(q.K);
    // layers outside the tensor-core tile limits run the fp32 kernels (higher precision, same ABI)
    q.tc = q.Kpad <= MAX_KPAD && (p->encoder || p->Cin % 4 == 0) && aligned16(p->in);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
    if (q.tc) {
        pick_nchunk(q.N, q.Kpad, &q.nch, &q.nc);
        q.off_lin = take((size_t)q.Mrows * q.N * sizeof(float));
        q.off_w = take((size_t)q.nc * q.nch * q.Kpad * sizeof(bf16));
        q.off_b = take((size_t)q.nc * q.nch * sizeof(float));
    } else {
        q.off_f32 = take(patch_ws_f32(p));
    }
    q.total = off;
    return q;
}

size_t patch_ws_bf16(const sf_patch_params* p) { return patch_plan(p).total; }

int patch_fwd_bf16(const sf_patch_params* p, void* ws_ptr, size_t ws_bytes, cudaStream_t st) {
    const PatchPlan q = patch_plan(p);
    if (ws_bytes < q.total || !ws_ptr) { set_error("sf_patch_fwd: workspace too small (%zu B given, %zu needed)", ws_bytes, q.total); return SF_ERR_WORKSPACE; }
    char* base = reinterpret_cast<char*>(ws_ptr);
    if (!q.tc) return patch_fwd_f32(p, base + q.off_f32, ws_bytes - q.off_f32, st);
    float* lin = reinterpret_cast<float*>(base + q.off_lin);
    bf16* wp = reinterpret_cast<bf16*>(base + q.off_w);
    float* bp = reinterpret_cast<float*>(base + q.off_b);
    PackSrc s{{p->w, nullptr, nullptr}, {p->b, nullptr, nullptr}};
    SF_TRY(launch_pack(s, 1, q.N, q.K, wp, bp, q.nch, q.Kpad, q.nc, 1, st));
    TcGemm g{};
    g.A = p->in; g.M = q.Mrows; g.K = q.K; g.Kpad = q.Kpad; g.lda = q.K;
    g.Wp = wp; g.NCH = q.nch; g.bias = bp; g.out = lin; g.ldo = q.N; g.N = q.N;
    if (p->encoder) {
        g.Hf = p->H; g.Wf = p->W; g.Cin = p->Cin; g.mh = p->mh; g.mw = p->mw;
        SF_TRY(launch_tc_gemm(AM_MERGE, false, g, q.nc, st));
        SF_TRY(launch_layernorm(lin, p->ln_gamma, p->ln_beta, p->out, q.Mrows, q.N, p->ln_eps, 1, nullptr, st));
    } else {
        SF_TRY(launch_tc_gemm(AM_F32, false, g, q.nc, st));
        UnmergeGeom ug{p->H, p->W, p->mh, p->mw, p->Cout};
        SF_TRY(launch_layernorm(lin, p->ln_gamma, p->ln_beta, p->out, q.Mrows, q.N, p->ln_eps, 1, &ug, st));
    }
    return SF_OK;
}

}  // namespace sf

// Persistent, warp-specialised tcgen05 token GEMM (sm_100a) and the weight packer.
//
//   out[M, N] = epilogue( prologue(A)[M, K] * W[N, K]^T )
//
// prologue: fp32 rows [-> LayerNorm] -> bf16 (written by 4 producer warps straight into the UMMA
//           smem layout), patch-merge gather, or bf16 activations already stored in the UMMA-tiled
//           layout in HBM (brought in by the bulk-copy engine, no thread touches them);
// epilogue: + bias, ELU, + residual; fp32 rows, bf16 rows, or bf16 UMMA-tiled (the next GEMM's A).
//
// One CTA per SM (two when the tile is small) loops over work items (m-tile of 128 tokens x group
// of n-chunks).  Warp roles:
//   warps 0-3, 4-7  two A-producer groups, each filling its own A buffer (alternate work items, so two
//                   tiles' global loads are in flight); idle in the STREAM flavour
//   warp  8    MMA issuer         one lane issues tcgen05.mma, commits to mbarriers
//   warp  9    loader             one lane drives cp.async.bulk for weight (and A) k-slabs
//   warps 10-17 epilogue          tcgen05.ld -> registers -> global; two warps per TMEM lane quarter, each
//                                 taking every other 16-column group (the epilogue is the longest stage)
// Pipelines: A buffers (full/empty), k-slab ring (full/empty), two TMEM accumulators (full/empty),
// so the producers work on item i+1 and the epilogue on chunk j-1 while the tensor core runs chunk j.
#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include <initializer_list>
#include "bf16_kernels.cuh"
#include "tc_common.cuh"

namespace sf {
using namespace tc;

static constexpr uint32_t SBO = 128;
static constexpr int G_THREADS = 576;          // RESIDENT flavour: 8 producer + MMA + loader + 8 epilogue warps
static constexpr int G_THREADS_STREAM = 320;   // STREAM flavour: no producer warps
static constexpr size_t SMEM_LIMIT = 227 * 1024;
static constexpr uint32_t EPI_TR_PITCH = 20;                           // floats per staged row: 16 columns + 4 (bank spread)
static constexpr uint32_t EPI_TR_BYTES = 32 * EPI_TR_PITCH * 4;        // one epilogue warp: 32 rows x 16 columns

__host__ __device__ static inline uint32_t align128(uint32_t v) { return (v + 127) & ~127u; }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// =============================================================================================
// weight packing
// =============================================================================================
// tr != 0: the (single) source is stored [K][Neach] row-major and the image is that of its transpose
// (out[n][k] = src[k][n]): the packed W^T of the backward data-gradient GEMMs dX = dY W.
__global__ void k_pack_w(PackSrc src, int nsrc, int Neach, int K, bf16* __restrict__ out, float* __restrict__ bias_out,
                         int NR, int KR, int n_chunks, int k_chunks, int tr) {
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
                v[e] = k < K ? (tr ? src.w[0][(long long)k * Neach + n] : row[k]) : 0.f;
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

int launch_pack(const PackSrc& src, int nsrc, int Neach, int K, bf16* out, float* bias_out, int NR, int KR, int n_chunks,
                int k_chunks, cudaStream_t st, int transposed) {
    SF_CHECK_ARG(!transposed || nsrc == 1, "pack: the transposed form takes one source");
    long long total = (long long)NR * KR / 8 * n_chunks * k_chunks;
    int blocks = (int)((total + 255) / 256);
    if (blocks > sm_count() * 8) blocks = sm_count() * 8;
    if (blocks < 1) blocks = 1;
    ProfScope ps("pack_weights_bf16", 0.0, 6.0 * (double)nsrc * Neach * K, st);
    k_pack_w<<<blocks, 256, 0, st>>>(src, nsrc, Neach, K, out, bias_out, NR, KR, n_chunks, k_chunks, transposed);
    SF_CHECK_LAUNCH("pack_weights_bf16");
    return SF_OK;
}

__global__ void k_pack_w_mapped(PackSrc src, PackMap map, int K, bf16* __restrict__ out, float* __restrict__ bias_out,
                                int NR, int KR, int n_chunks, int k_chunks) {
    const long long cpi = (long long)NR * KR / 8;
    const long long total = cpi * n_chunks * k_chunks;
    // packed row n -> (source, source row) or nothing
    auto locate = [&](int n, int* s_out) -> int {
        int start = 0;
        for (int s = 0; s < map.nsrc; s++) {
            if (n < start + map.rows[s]) {
                int np = n - start;
                *s_out = s;
                if (!map.head_padded[s]) return np;
                int h = np / map.dp, dd = np - h * map.dp;
                if (dd == map.d && map.ones_pad[s]) return -2;   // zero weights, bias 1: a column of ones
                return dd < map.d ? h * map.d + dd : -1;
            }
            start += map.rows[s];
        }
        *s_out = 0;
        return -1;
    };
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < total; c += (long long)gridDim.x * blockDim.x) {
        long long img = c / cpi;
        int ci = (int)(c - img * cpi);
        int jk = (int)(img % k_chunks), jn = (int)(img / k_chunks);
        int kc = ci / NR, r = ci - kc * NR;
        int sidx;
        const int srow = locate(jn * NR + r, &sidx);
        uint32_t pk[4] = {0, 0, 0, 0};
        if (srow >= 0) {
            const float* row = src.w[sidx] + (long long)srow * K;
            const float sc = map.scale[sidx];
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {
                int k = jk * KR + kc * 8 + e;
                if (map.kdp > 0) {   // head-padded K axis
                    int h = k / map.kdp, dd = k - h * map.kdp;
                    k = dd < map.kd ? h * map.kd + dd : K;
                }
                v[e] = k < K ? row[k] * sc : 0.f;
            }
#pragma unroll
            for (int e = 0; e < 4; e++) pk[e] = pack_bf16x2(v[2 * e], v[2 * e + 1]);
        }
        *reinterpret_cast<uint4*>(out + c * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
    if (bias_out) {
        for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < n_chunks * NR; n += gridDim.x * blockDim.x) {
            int sidx;
            const int srow = locate(n, &sidx);
            bias_out[n] = srow == -2 ? 1.f : ((srow >= 0 && src.b[sidx]) ? src.b[sidx][srow] * map.scale[sidx] : 0.f);
        }
    }
}

// several plain / transposed images in ONE launch (blockIdx.y = job): the backward pass packs up to seven images per operator
__global__ void k_pack_jobs(PackJobs jobs) {
    const PackJob& j = jobs.job[blockIdx.y];
    const long long cpi = (long long)j.NR * j.KR / 8;
    const long long total = cpi * j.n_chunks * j.k_chunks;
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < total; c += (long long)gridDim.x * blockDim.x) {
        const long long img = c / cpi;
        const int ci = (int)(c - img * cpi);
        const int jk = (int)(img % j.k_chunks), jn = (int)(img / j.k_chunks);
        const int kc = ci / j.NR, r = ci - kc * j.NR;
        const int n = jn * j.NR + r;
        uint32_t pk[4] = {0, 0, 0, 0};
        if (n < j.N) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const int k = jk * j.KR + kc * 8 + e;
                float x = 0.f;
                if (k < j.K) {
                    if (j.transposed) { const int s = k / j.each; x = j.w[s][(long long)(k - s * j.each) * j.N + n]; }   // sources stacked along K
                    else { const int s = n / j.each; x = j.w[s][(long long)(n - s * j.each) * j.K + k]; }               // stacked along N
                }
                v[e] = x;
            }
#pragma unroll
            for (int e = 0; e < 4; e++) pk[e] = pack_bf16x2(v[2 * e], v[2 * e + 1]);
        }
        *reinterpret_cast<uint4*>(j.out + c * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < j.n_chunks * j.NR; n += gridDim.x * blockDim.x) {
        float b = 0.f;
        if (n < j.N && !j.transposed) { const int s = n / j.each; if (j.b[s]) b = j.b[s][n - s * j.each]; }
        j.bias_out[n] = b;
    }
}

int launch_pack_jobs(const PackJobs& jobs, cudaStream_t st) {
    SF_CHECK_ARG(jobs.n >= 1 && jobs.n <= PackJobs::MAX, "pack: bad job count %d", jobs.n);
    long long most = 0;
    double bytes = 0;
    for (int i = 0; i < jobs.n; i++) {
        const PackJob& j = jobs.job[i];
        most = std::max(most, (long long)j.NR * j.KR / 8 * j.n_chunks * j.k_chunks);
        bytes += 6.0 * j.N * j.K;
    }
    int bx = (int)std::min<long long>((most + 255) / 256, 64);
    ProfScope ps("pack_weights_bf16", 0.0, bytes, st);
    k_pack_jobs<<<dim3((unsigned)std::max(bx, 1), (unsigned)jobs.n), 256, 0, st>>>(jobs);
    SF_CHECK_LAUNCH("pack_weights_bf16");
    return SF_OK;
}

int launch_pack_mapped(const PackSrc& src, const PackMap& map, int K, bf16* out, float* bias_out, int NR, int KR, int n_chunks,
                       int k_chunks, cudaStream_t st) {
    long long total = (long long)NR * KR / 8 * n_chunks * k_chunks;
    int blocks = (int)((total + 255) / 256);
    if (blocks > sm_count() * 8) blocks = sm_count() * 8;
    if (blocks < 1) blocks = 1;
    double rows = 0;
    for (int s = 0; s < map.nsrc; s++) rows += map.rows[s];
    ProfScope ps("pack_weights_bf16", 0.0, 6.0 * rows * K, st);
    k_pack_w_mapped<<<blocks, 256, 0, st>>>(src, map, K, out, bias_out, NR, KR, n_chunks, k_chunks);
    SF_CHECK_LAUNCH("pack_weights_bf16");
    return SF_OK;
}

// =============================================================================================
// A producers (RESIDENT flavour): a [128 x Kpad] bf16 operand in the UMMA layout, LBO = 2064
// =============================================================================================
static constexpr uint32_t LBO_P = lbo_padded(128);  // producer-written A: 2064 B between k-chunks
static constexpr uint32_t LBO_T = lbo_dense(128);   // bulk-copied (tiled) A: 2048 B

// fp32 rows (optionally LayerNorm-ed) -> bf16.  Lanes of a warp split into groups of LPR lanes,
// one group per row, NPER float4 per lane; RB row-iterations are loaded before any is processed
// so that each lane keeps RB*NPER 16-byte loads in flight.
template <bool LN, int NPER, int RB>
__device__ __forceinline__ void produce_rows(uint8_t* sA, const float* __restrict__ A, long long lda, long long M, long long m0,
                                             int K, int Kpad, const float* __restrict__ g, const float* __restrict__ b, float eps,
                                             int ptid, const WinOrder* wo) {
    const int lane = ptid & 31, warp = ptid >> 5;
    const int nf4 = K >> 2, nslots = Kpad >> 2;
    int LPR = 1;
    while (LPR < 32 && LPR < nslots) LPR <<= 1;
    const int RPW = 32 / LPR;
    const int gl = lane & (LPR - 1), gr = lane / LPR;
    const int iters = 32 / RPW;
    for (int it0 = 0; it0 < iters; it0 += RB) {
        float4 v[RB][NPER];
#pragma unroll
        for (int u = 0; u < RB; u++) {
            const long long m = m0 + warp * 32 + (it0 + u) * RPW + gr;
            const bool rowok = (it0 + u) < iters && m < M;
            const long long ms = (wo && rowok) ? win_order_token(*wo, (uint32_t)m) : m;   // window order: gather the source row
#pragma unroll
            for (int i = 0; i < NPER; i++) {
                int q = gl + i * LPR;
                v[u][i] = (rowok && q < nf4) ? *reinterpret_cast<const float4*>(A + ms * lda + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int u = 0; u < RB; u++) {
            if (it0 + u >= iters) break;  // uniform
            const int r = warp * 32 + (it0 + u) * RPW + gr;
            const bool rowok = (m0 + r) < M;
            if (LN) {
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < NPER; i++) s += (v[u][i].x + v[u][i].y) + (v[u][i].z + v[u][i].w);
                for (int o = LPR >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                const float mean = s / (float)K;
                float ss = 0.f;
#pragma unroll
                for (int i = 0; i < NPER; i++) {
                    if (gl + i * LPR < nf4) {
                        float dx = v[u][i].x - mean, dy = v[u][i].y - mean, dz = v[u][i].z - mean, dw = v[u][i].w - mean;
                        ss += (dx * dx + dy * dy) + (dz * dz + dw * dw);
                    }
                }
                for (int o = LPR >> 1; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
                const float rstd = rsqrtf(ss / (float)K + eps);
#pragma unroll
                for (int i = 0; i < NPER; i++) {
                    int q = gl + i * LPR;
                    if (rowok && q < nf4) {
                        float4 gg = __ldg(reinterpret_cast<const float4*>(g) + q), bb = __ldg(reinterpret_cast<const float4*>(b) + q);
                        v[u][i].x = (v[u][i].x - mean) * rstd * gg.x + bb.x;
                        v[u][i].y = (v[u][i].y - mean) * rstd * gg.y + bb.y;
                        v[u][i].z = (v[u][i].z - mean) * rstd * gg.z + bb.z;
                        v[u][i].w = (v[u][i].w - mean) * rstd * gg.w + bb.w;
                    } else {
                        v[u][i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < NPER; i++) {
                int q = gl + i * LPR;
                if (q < nslots) {
                    uint2 pk = make_uint2(pack_bf16x2(v[u][i].x, v[u][i].y), pack_bf16x2(v[u][i].z, v[u][i].w));
                    *reinterpret_cast<uint2*>(sA + (uint32_t)(q >> 1) * LBO_P + (uint32_t)r * 16 + (q & 1) * 8) = pk;
                }
            }
        }
    }
}

// Short rows (K <= 64): one thread per row, the whole row in registers -- no shuffles, the 32 rows
// of a warp are independent instruction streams, st.shared of 16-byte chunks is conflict-free
// (consecutive lanes = consecutive rows = consecutive 16 bytes).  A warp reads 32 consecutive rows,
// i.e. one contiguous 32*K*4-byte span; the sectors a single LDG.128 half-uses are completed by the
// next one out of L1.  Split into a load and a finish phase so that the producer loop can keep the
// NEXT tile's rows in flight while it normalises the current one (these GEMMs are HBM-bound and a
// producer group that waits out every load latency cannot cover it).
template <int NF4MAX>
__device__ __forceinline__ void load_row_thread(float4 (&v)[NF4MAX], const float* __restrict__ A, long long lda, long long M,
                                                long long m0, int K, int ptid, const WinOrder* wo) {
    const long long m = m0 + ptid;
    const int nf4 = K >> 2;
    const bool rowok = m < M;
    const long long ms = (wo && rowok) ? win_order_token(*wo, (uint32_t)m) : m;   // window order: gather the source row
    const float4* src = reinterpret_cast<const float4*>(A + ms * lda);
#pragma unroll
    for (int i = 0; i < NF4MAX; i++) v[i] = (rowok && i < nf4) ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
}

template <bool LN, int NF4MAX>
__device__ __forceinline__ void finish_row_thread(uint8_t* sA, float4 (&v)[NF4MAX], long long M, long long m0, int K, int Kpad,
                                                  const float* __restrict__ g, const float* __restrict__ b, float eps, int ptid) {
    const int r = ptid;
    const int nf4 = K >> 2, nkc = Kpad >> 3;
    const bool rowok = m0 + r < M;
    if (LN) {
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < NF4MAX; i++) { s0 += v[i].x + v[i].y; s1 += v[i].z + v[i].w; }
        const float invk = 1.f / (float)K;
        const float mean = (s0 + s1) * invk;
        float q0 = 0.f, q1 = 0.f;
#pragma unroll
        for (int i = 0; i < NF4MAX; i++) {
            if (i < nf4) {
                float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
                q0 += dx * dx + dy * dy; q1 += dz * dz + dw * dw;
            }
        }
        const float rstd = rsqrtf((q0 + q1) * invk + eps);
#pragma unroll
        for (int i = 0; i < NF4MAX; i++) {
            if (i < nf4) {
                float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i), bb = __ldg(reinterpret_cast<const float4*>(b) + i);
                v[i].x = (v[i].x - mean) * rstd * gg.x + bb.x;
                v[i].y = (v[i].y - mean) * rstd * gg.y + bb.y;
                v[i].z = (v[i].z - mean) * rstd * gg.z + bb.z;
                v[i].w = (v[i].w - mean) * rstd * gg.w + bb.w;
            }
        }
        if (!rowok) {
#pragma unroll
            for (int i = 0; i < NF4MAX; i++) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
#pragma unroll
    for (int c = 0; c < NF4MAX / 2; c++) {
        if (c < nkc) {
            uint4 pk = make_uint4(pack_bf16x2(v[2 * c].x, v[2 * c].y), pack_bf16x2(v[2 * c].z, v[2 * c].w),
                                  pack_bf16x2(v[2 * c + 1].x, v[2 * c + 1].y), pack_bf16x2(v[2 * c + 1].z, v[2 * c + 1].w));
            *reinterpret_cast<uint4*>(sA + (uint32_t)c * LBO_P + (uint32_t)r * 16) = pk;
        }
    }
}

template <bool LN, int NF4MAX>
__device__ __forceinline__ void produce_rows_thread(uint8_t* sA, const float* __restrict__ A, long long lda, long long M,
                                                    long long m0, int K, int Kpad, const float* __restrict__ g,
                                                    const float* __restrict__ b, float eps, int ptid, const WinOrder* wo) {
    float4 v[NF4MAX];
    load_row_thread<NF4MAX>(v, A, lda, M, m0, K, ptid, wo);
    finish_row_thread<LN, NF4MAX>(sA, v, M, m0, K, Kpad, g, b, eps, ptid);
}

template <bool LN>
__device__ __forceinline__ void produce_a_f32(uint8_t* sA, const TcGemm& p, long long m0, int ptid) {
    const float* A = reinterpret_cast<const float*>(p.A);
    const int nslots = p.Kpad >> 2;
    const WinOrder* wo = p.win_order ? &p.wo : nullptr;
    if (nslots <= 8) produce_rows_thread<LN, 8>(sA, A, p.lda, p.M, m0, p.K, p.Kpad, p.ln_g, p.ln_b, p.eps, ptid, wo);
    else if (nslots <= 16) produce_rows_thread<LN, 16>(sA, A, p.lda, p.M, m0, p.K, p.Kpad, p.ln_g, p.ln_b, p.eps, ptid, wo);
    else if (nslots <= 32) produce_rows<LN, 1, 8>(sA, A, p.lda, p.M, m0, p.K, p.Kpad, p.ln_g, p.ln_b, p.eps, ptid, wo);
    else if (nslots <= 64) produce_rows<LN, 2, 4>(sA, A, p.lda, p.M, m0, p.K, p.Kpad, p.ln_g, p.ln_b, p.eps, ptid, wo);
    else produce_rows<LN, 3, 4>(sA, A, p.lda, p.M, m0, p.K, p.Kpad, p.ln_g, p.ln_b, p.eps, ptid, wo);
}

// patch-merge gather (a011:87-93): row (b,Y,X), k = (ph*mw+pw)*Cin + c <- in[b][Y*mh+ph][X*mw+pw][c]
// One thread per row: the row's pixel coordinates are decoded once; with Cin % 8 == 0 every 8-element
// k-chunk is 8 consecutive channels of one source pixel (two 16-byte loads).
__device__ __forceinline__ void produce_a_merge(uint8_t* sA, const TcGemm& p, long long m0, int ptid) {
    const float* __restrict__ in = reinterpret_cast<const float*>(p.A);
    const int nkc = p.Kpad >> 3;
    const int Hc = p.Hf / p.mh, Wc = p.Wf / p.mw;
    const int r = ptid;
    const long long m = m0 + r;
    const bool rowok = m < p.M;
    int X = 0, Y = 0;
    long long b = 0;
    if (rowok) {
        X = (int)(m % Wc);
        long long t = m / Wc;
        Y = (int)(t % Hc);
        b = t / Hc;
    }
    const float* base = in + ((b * p.Hf + (long long)Y * p.mh) * p.Wf + (long long)X * p.mw) * p.Cin;   // pixel (ph, pw) = (0, 0)
    const long long rowpitch = (long long)p.Wf * p.Cin;
    if ((p.Cin & 7) == 0) {
        const int cpp = p.Cin >> 3;   // chunks per source pixel
        int q = 0, cc = 0, ph = 0, pw = 0;
        for (int kc = 0; kc < nkc; kc++) {
            uint4 pk = make_uint4(0u, 0u, 0u, 0u);
            if (rowok && kc * 8 < p.K) {
                const float4* src = reinterpret_cast<const float4*>(base + ph * rowpitch + (long long)pw * p.Cin + cc * 8);
                const float4 lo = src[0], hi = src[1];
                pk = make_uint4(pack_bf16x2(lo.x, lo.y), pack_bf16x2(lo.z, lo.w), pack_bf16x2(hi.x, hi.y), pack_bf16x2(hi.z, hi.w));
            }
            *reinterpret_cast<uint4*>(sA + (uint32_t)kc * LBO_P + (uint32_t)r * 16) = pk;
            if (++cc == cpp) { cc = 0; q++; if (++pw == p.mw) { pw = 0; ph++; } }
        }
    } else {
        for (int kc = 0; kc < nkc; kc++) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const int k = kc * 8 + e;
                v[e] = 0.f;
                if (rowok && k < p.K) {
                    const int q = k / p.Cin, c = k - q * p.Cin;
                    const int ph = q / p.mw, pw = q - ph * p.mw;
                    v[e] = base[ph * rowpitch + (long long)pw * p.Cin + c];
                }
            }
            *reinterpret_cast<uint4*>(sA + (uint32_t)kc * LBO_P + (uint32_t)r * 16) =
                make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        }
    }
}


// =============================================================================================
// LayerNorm pre-pass for wide rows: fp32 rows -> LN -> bf16 in the UMMA-tiled layout, so that the
// GEMMs consuming it run in the STREAM flavour (A arrives by bulk copy, no producer warps in the
// critical path and no per-n-group recomputation of the LayerNorm).  One warp per row.
// =============================================================================================
template <int LPR>
__global__ void __launch_bounds__(256) k_ln_to_tiled(const float* __restrict__ in, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     bf16* __restrict__ out, long long M, int C, int Kpad, float eps, int gather, WinOrder wo,
                                                     long long ld_in, int out_nkc, int out_kc0) {
    // LPR lanes per row, three float4 per lane (C <= 12 * LPR), 32 / LPR rows per warp.
    // gamma == nullptr: no LayerNorm, a plain fp32 -> bf16 cast into the tiled layout (backward pass operands).
    // Rows M .. 128*ceil(M/128)-1 of the last tile are written as zeros: the weight-gradient GEMM sums over token rows.
    const int lane = threadIdx.x & 31, l = lane & (LPR - 1), sub = lane / LPR;
    constexpr int RPW = 32 / LPR;
    const int nf4 = C >> 2, nslots = Kpad >> 2;
    const long long Mpad = (M + 127) / 128 * 128;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r0 = warp0 * RPW; r0 < Mpad; r0 += nwarps * RPW) {
        const long long row = r0 + sub;
        float4 v[3];
        const bool real = row < M;
        const long long srow = (gather && real) ? win_order_token(wo, (uint32_t)row) : row;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            int q = l + LPR * i;
            v[i] = (real && q < nf4) ? *reinterpret_cast<const float4*>(in + srow * ld_in + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float mean = 0.f, rstd = 1.f;
        if (gamma) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 3; i++) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
            for (int o = LPR >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            mean = s / (float)C;
            float ss = 0.f;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                if (l + LPR * i < nf4) {
                    float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
                    ss += (dx * dx + dy * dy) + (dz * dz + dw * dw);
                }
            }
#pragma unroll
            for (int o = LPR >> 1; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            rstd = rsqrtf(ss / (float)C + eps);
        }
        if (row >= Mpad) continue;
        const long long tile = row >> 7;
        const int r = (int)(row & 127);
#pragma unroll
        for (int i = 0; i < 3; i++) {
            int q = l + LPR * i;
            if (q < nslots) {
                uint2 pk = make_uint2(0u, 0u);
                if (real && q < nf4) {
                    if (gamma) {
                        float4 gg = __ldg(reinterpret_cast<const float4*>(gamma) + q), bb = __ldg(reinterpret_cast<const float4*>(beta) + q);
                        pk = make_uint2(pack_bf16x2((v[i].x - mean) * rstd * gg.x + bb.x, (v[i].y - mean) * rstd * gg.y + bb.y),
                                        pack_bf16x2((v[i].z - mean) * rstd * gg.z + bb.z, (v[i].w - mean) * rstd * gg.w + bb.w));
                    } else {
                        pk = make_uint2(pack_bf16x2(v[i].x, v[i].y), pack_bf16x2(v[i].z, v[i].w));
                    }
                }
                *reinterpret_cast<uint2*>(out + ((tile * out_nkc + out_kc0 + (q >> 1)) * 128 + r) * 8 + (q & 1) * 4) = pk;
            }
        }
    }
}

// Narrow rows (Kpad <= 64): one thread per row -- a warp-per-row mapping would leave most lanes idle at C = 24 / 48.
// A warp reads 32 consecutive rows (one contiguous span, completed out of L1 over the NF4 loads) and stores
// 32 consecutive 16-byte chunks per k-chunk.
template <int NF4>
__global__ void __launch_bounds__(256) k_ln_to_tiled_rows(const float* __restrict__ in, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, bf16* __restrict__ out, long long M, int C, int Kpad,
                                                          float eps, int gather, WinOrder wo, long long ld_in, int out_nkc, int out_kc0) {
    const int nf4 = C >> 2, nkc = Kpad >> 3;
    const long long Mpad = (M + 127) / 128 * 128;
    for (long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x; row < Mpad; row += (long long)gridDim.x * blockDim.x) {
        const bool real = row < M;
        const long long srow = (gather && real) ? win_order_token(wo, (uint32_t)row) : row;
        const float4* src = reinterpret_cast<const float4*>(in + srow * ld_in);
        float4 v[NF4];
#pragma unroll
        for (int i = 0; i < NF4; i++) v[i] = (real && i < nf4) ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        if (gamma && real) {
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int i = 0; i < NF4; i++) { s0 += v[i].x + v[i].y; s1 += v[i].z + v[i].w; }
            const float invc = 1.f / (float)C;
            const float mean = (s0 + s1) * invc;
            float q0 = 0.f, q1 = 0.f;
#pragma unroll
            for (int i = 0; i < NF4; i++) {
                if (i < nf4) {
                    const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
                    q0 += dx * dx + dy * dy; q1 += dz * dz + dw * dw;
                }
            }
            const float rstd = rsqrtf((q0 + q1) * invc + eps);
#pragma unroll
            for (int i = 0; i < NF4; i++) {
                if (i < nf4) {
                    const float4 gg = __ldg(reinterpret_cast<const float4*>(gamma) + i), bb = __ldg(reinterpret_cast<const float4*>(beta) + i);
                    v[i].x = (v[i].x - mean) * rstd * gg.x + bb.x;
                    v[i].y = (v[i].y - mean) * rstd * gg.y + bb.y;
                    v[i].z = (v[i].z - mean) * rstd * gg.z + bb.z;
                    v[i].w = (v[i].w - mean) * rstd * gg.w + bb.w;
                }
            }
        }
        const long long tile = row >> 7;
        const int r = (int)(row & 127);
#pragma unroll
        for (int c = 0; c < NF4 / 2; c++) {
            if (c < nkc)
                *reinterpret_cast<uint4*>(out + ((tile * out_nkc + out_kc0 + c) * 128 + r) * 8) =
                    make_uint4(pack_bf16x2(v[2 * c].x, v[2 * c].y), pack_bf16x2(v[2 * c].z, v[2 * c].w),
                               pack_bf16x2(v[2 * c + 1].x, v[2 * c + 1].y), pack_bf16x2(v[2 * c + 1].z, v[2 * c + 1].w));
        }
    }
}

int launch_ln_to_tiled(const float* in, const float* gamma, const float* beta, bf16* out, long long M, int C, float eps, cudaStream_t st,
                       const WinOrder* wo, long long ld_in, int out_nkc, int out_kc0) {
    const int Kpad = (int)pad16((uint32_t)C);
    if (ld_in <= 0) ld_in = C;
    if (out_nkc <= 0) { out_nkc = Kpad >> 3; out_kc0 = 0; }
    SF_CHECK_ARG(ld_in % 4 == 0 && ((reinterpret_cast<uintptr_t>(in) & 15) == 0), "ln_to_tiled: rows must be 16-byte aligned");
    SF_CHECK_ARG(C % 4 == 0 && Kpad <= TC_MAX_KPAD, "ln_to_tiled: unsupported row width %d", C);
    long long blocks = (M * 32 + 255) / 256;
    if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
    if (blocks < 1) blocks = 1;
    ProfScope ps(prof_name(gamma ? "ln_to_tiled_c%d" : "cast_to_tiled_c%d", C), gamma ? 8.0 * (double)M * C : 0.0, 6.0 * (double)M * C, st);
    SF_CHECK_ARG(!wo || M < 2147483647LL, "ln_to_tiled: %lld rows exceed the window-order index range", M);
    if (Kpad <= 64) {
        long long rb = (M + 255) / 256;
        if (rb > (long long)sm_count() * 8) rb = (long long)sm_count() * 8;
        if (rb < 1) rb = 1;
        if (Kpad <= 32) k_ln_to_tiled_rows<8><<<(unsigned)rb, 256, 0, st>>>(in, gamma, beta, out, M, C, Kpad, eps, wo ? 1 : 0, wo ? *wo : WinOrder{}, ld_in, out_nkc, out_kc0);
        else k_ln_to_tiled_rows<16><<<(unsigned)rb, 256, 0, st>>>(in, gamma, beta, out, M, C, Kpad, eps, wo ? 1 : 0, wo ? *wo : WinOrder{}, ld_in, out_nkc, out_kc0);
        SF_CHECK_LAUNCH("ln_to_tiled");
        return SF_OK;
    }
    const WinOrder woa = wo ? *wo : WinOrder{};
    // (Round 2 tried a variant that builds the tile image in shared memory and writes it out with cp.async.bulk, whole lines instead
    // of 8-byte fragments 2 KB apart: bit-identical output, same device time within 0.1 % in a same-box A/B -- a 12 us kernel moving
    // 29 MB is launch ramp and DRAM page opening, not store efficiency.  Removed.  Likewise requesting the first k-slabs before the
    // GEMM's prologue barrier: +1 % step time, removed.)
    if (Kpad <= 96) {
        blocks = std::min<long long>((M / 4 * 32 + 255) / 256 + 1, (long long)sm_count() * 16);
        k_ln_to_tiled<8><<<(unsigned)blocks, 256, 0, st>>>(in, gamma, beta, out, M, C, Kpad, eps, wo ? 1 : 0, woa, ld_in, out_nkc, out_kc0);
    } else if (Kpad <= 192) {
        blocks = std::min<long long>((M / 2 * 32 + 255) / 256 + 1, (long long)sm_count() * 16);
        k_ln_to_tiled<16><<<(unsigned)blocks, 256, 0, st>>>(in, gamma, beta, out, M, C, Kpad, eps, wo ? 1 : 0, woa, ld_in, out_nkc, out_kc0);
    } else {
        k_ln_to_tiled<32><<<(unsigned)blocks, 256, 0, st>>>(in, gamma, beta, out, M, C, Kpad, eps, wo ? 1 : 0, woa, ld_in, out_nkc, out_kc0);
    }
    SF_CHECK_LAUNCH("ln_to_tiled");
    return SF_OK;
}

// =============================================================================================
// the kernel
// =============================================================================================
struct GemmSmem {
    uint32_t a_bytes, a_off[2], w_off, stage_bytes, slabA_bytes, slabW_bytes, bias_off, tr_off, bar_off, total;
};

__host__ __device__ static inline GemmSmem gemm_smem_layout(const TcGemm& p) {
    GemmSmem s{};
    const bool stream = p.a_mode == AM_TILED;
    s.a_bytes = stream ? 0u : align128((uint32_t)(p.Kpad >> 3) * LBO_P);
    s.a_off[0] = 0;
    s.a_off[1] = s.a_bytes;
    s.w_off = (uint32_t)p.NA * s.a_bytes;
    s.slabW_bytes = (uint32_t)p.NCH * (uint32_t)p.KS * 2u;
    s.slabA_bytes = stream ? (uint32_t)(p.KS >> 3) * LBO_T : 0u;
    s.stage_bytes = align128(s.slabW_bytes) + align128(s.slabA_bytes);
    s.bias_off = s.w_off + (uint32_t)p.NS * s.stage_bytes;
    s.tr_off = s.bias_off + align128(p.bias ? (uint32_t)p.n_chunks * (uint32_t)p.NCH * 4u : 0u);
    s.bar_off = s.tr_off + ((p.out_mode == OUT_F32 && p.coal) ? 8u * EPI_TR_BYTES : 0u);   // fp32 rows leave through a per-warp transposition buffer
    s.total = s.bar_off + 256;
    return s;
}

template <int AMODE, int OUTMODE>
__global__ void __launch_bounds__(AMODE == AM_TILED ? G_THREADS_STREAM : G_THREADS, 1) k_tc_gemm2(TcGemm p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr bool STREAM = (AMODE == AM_TILED);
    constexpr int PW = STREAM ? 0 : 8;          // producer warps; roles after them: MMA, loader, 8 epilogue warps
    const GemmSmem L = gemm_smem_layout(p);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint64_t* a_full = bars;            // [2]
    uint64_t* a_empty = bars + 2;       // [2]
    uint64_t* w_full = bars + 4;        // [NS <= 8]
    uint64_t* w_empty = bars + 12;      // [NS <= 8]
    uint64_t* d_full = bars + 20;       // [2]
    uint64_t* d_empty = bars + 22;      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
    const uint32_t acc_stride = ((uint32_t)p.NCH + 31u) & ~31u;
    const uint32_t ncols = tmem_cols_pow2(2u * acc_stride);
    const int NS = p.NS, NA = p.NA, n_slabs = p.n_slabs, ksteps = p.KS >> 4;
    const long long m_tiles = (p.M + 127) / 128;
    const long long items = m_tiles * p.n_groups;

    if (tid == PW * 32) {  // MMA warp, lane 0
        for (int i = 0; i < 2; i++) { mbar_init(&a_full[i], 128); mbar_init(&a_empty[i], 1); mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], 8); }
        for (int i = 0; i < NS; i++) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        fence_mbar_init();
    }
    if (warp == PW + 1) tmem_alloc(tmem_slot, ncols);
    // bias vector of all n-chunks -> shared memory (the epilogue reads it as broadcast float4s)
    const float* sbias = reinterpret_cast<const float*>(smem + L.bias_off);
    if (p.bias) {
        float* sb = reinterpret_cast<float*>(smem + L.bias_off);
        for (int i = tid; i < p.n_chunks * p.NCH; i += blockDim.x) sb[i] = __ldg(p.bias + i);
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < PW) {
        // ------------------------------ A producers -------------------------------------------------
        const uint32_t grp = (uint32_t)warp >> 2;
        const int ptid = tid & 127;
        const int nslots_a = p.Kpad >> 2;
        if ((AMODE == AM_F32 || AMODE == AM_F32_LN) && nslots_a <= 8 && (NA == 2 || grp == 0)) {
            // thread-per-row producers, software pipelined: the rows of this group's NEXT item are in flight
            // (registers) while the current ones are normalised and written to shared memory
            const float* Af = reinterpret_cast<const float*>(p.A);
            const WinOrder* wo = p.win_order ? &p.wo : nullptr;
            const long long step = (long long)gridDim.x * NA;
            auto run = [&](auto nf4tag) {
                constexpr int NF = decltype(nf4tag)::value;
                float4 vn[NF];
                long long item = blockIdx.x + (long long)(NA == 2 ? grp : 0) * gridDim.x;
                if (item < items) load_row_thread<NF>(vn, Af, p.lda, p.M, (item / p.n_groups) * 128, p.K, ptid, wo);
                for (uint32_t it = 0; item < items; item += step, it++) {
                    float4 v[NF];
#pragma unroll
                    for (int i = 0; i < NF; i++) v[i] = vn[i];
                    const long long m0 = (item / p.n_groups) * 128;
                    if (item + step < items) load_row_thread<NF>(vn, Af, p.lda, p.M, ((item + step) / p.n_groups) * 128, p.K, ptid, wo);
                    const uint32_t ab = NA == 2 ? grp : 0u;
                    mbar_wait_relaxed(&a_empty[ab], (it & 1u) ^ 1u);
                    finish_row_thread<AMODE == AM_F32_LN, NF>(smem + L.a_off[ab], v, p.M, m0, p.K, p.Kpad, p.ln_g, p.ln_b, p.eps, ptid);
                    fence_async_smem();
                    mbar_arrive(&a_full[ab]);
                }
            };
            run(std::integral_constant<int, 8>{});   // wider rows would spill at this CTA size: they take the plain loop below
        } else if (!STREAM && (NA == 2 || grp == 0)) {
            uint32_t acount = 0;
            for (long long item = blockIdx.x; item < items; item += gridDim.x, acount++) {
                const long long tile = item / p.n_groups;
                const uint32_t ab = acount % (uint32_t)NA;
                if (NA == 2 && ab != grp) continue;
                mbar_wait_relaxed(&a_empty[ab], ((acount / (uint32_t)NA) & 1u) ^ 1u);
                uint8_t* sA = smem + L.a_off[ab];
                if (AMODE == AM_F32_LN) produce_a_f32<true>(sA, p, tile * 128, ptid);
                else if (AMODE == AM_F32) produce_a_f32<false>(sA, p, tile * 128, ptid);
                else produce_a_merge(sA, p, tile * 128, ptid);
                fence_async_smem();
                mbar_arrive(&a_full[ab]);
            }
        }
    } else if (warp == PW) {
        // ------------------------------ MMA issuer ----------------------------------------------------
        if (lane == 0) {
            uint32_t acount = 0, wit = 0, tcount = 0;
            const uint32_t idesc = make_idesc_bf16(128, (uint32_t)p.NCH);
            const uint32_t lbo_w = lbo_dense((uint32_t)p.NCH);
            const uint32_t lbo_a = STREAM ? LBO_T : LBO_P;
            for (long long item = blockIdx.x; item < items; item += gridDim.x, acount++) {
                const int group = (int)(item % p.n_groups);
                const uint32_t ab = acount % (uint32_t)NA;
                if (!STREAM) {
                    mbar_wait(&a_full[ab], (acount / (uint32_t)NA) & 1u);
                    tc_fence_after_sync();
                }
                const int c0 = group * p.chunks_per_group;
                const int c1 = min(c0 + p.chunks_per_group, p.n_chunks);
                for (int c = c0; c < c1; c++, tcount++) {
                    const uint32_t acc = tcount & 1u;
                    mbar_wait(&d_empty[acc], ((tcount >> 1) & 1u) ^ 1u);
                    tc_fence_after_sync();
                    const uint32_t tacc = tmem_base + acc * acc_stride;
                    for (int s = 0; s < n_slabs; s++, wit++) {
                        const uint32_t stg = wit % (uint32_t)NS;
                        mbar_wait(&w_full[stg], (wit / (uint32_t)NS) & 1u);
                        tc_fence_after_sync();
                        const uint32_t wbase = smem_u32(smem + L.w_off + stg * L.stage_bytes);
                        const uint32_t abase = STREAM ? wbase + align128(L.slabW_bytes)
                                                      : smem_u32(smem + L.a_off[ab]) + (uint32_t)s * (uint32_t)(p.KS >> 3) * lbo_a;
                        for (int ks = 0; ks < ksteps; ks++) {
                            uint64_t da = make_smem_desc(abase + (uint32_t)ks * 2u * lbo_a, lbo_a, SBO);
                            uint64_t db = make_smem_desc(wbase + (uint32_t)ks * 2u * lbo_w, lbo_w, SBO);
                            umma_bf16(tacc, da, db, idesc, (s | ks) != 0);
                        }
                        umma_commit(&w_empty[stg]);
                    }
                    umma_commit(&d_full[acc]);
                }
                if (!STREAM) umma_commit(&a_empty[ab]);
            }
        }
    } else if (warp == PW + 1) {
        // ------------------------------ loader (bulk-copy engine) --------------------------------------
        if (lane == 0) {
            uint32_t wit = 0;
            for (long long item = blockIdx.x; item < items; item += gridDim.x) {
                const long long tile = item / p.n_groups;
                const int group = (int)(item % p.n_groups);
                const int c0 = group * p.chunks_per_group;
                const int c1 = min(c0 + p.chunks_per_group, p.n_chunks);
                for (int c = c0; c < c1; c++) {
                    for (int s = 0; s < n_slabs; s++, wit++) {
                        const uint32_t stg = wit % (uint32_t)NS;
                        mbar_wait(&w_empty[stg], ((wit / (uint32_t)NS) & 1u) ^ 1u);
                        uint8_t* dstW = smem + L.w_off + stg * L.stage_bytes;
                        mbar_arrive_expect_tx(&w_full[stg], L.slabW_bytes + L.slabA_bytes);
                        bulk_g2s(dstW, p.Wp + ((size_t)c * n_slabs + s) * (size_t)p.NCH * p.KS, L.slabW_bytes, &w_full[stg]);
                        if (STREAM) {
                            const bf16* srcA = reinterpret_cast<const bf16*>(p.A) + ((size_t)tile * p.a_tile_nkc + p.a_kc0 + (size_t)s * (p.KS >> 3)) * 128 * 8;
                            bulk_g2s(dstW + align128(L.slabW_bytes), srcA, L.slabA_bytes, &w_full[stg]);
                        }
                    }
                }
            }
        }
    } else {
        // ------------------------------ epilogue ---------------------------------------------------------
        const int rb = warp & 3;                 // TMEM lane quarter this warp may access (hardware: warp id % 4)
        const int eg = (warp - (PW + 2)) >> 2;    // which half of the 16-column groups
        const int row = rb * 32 + lane;
        uint32_t tcount = 0;
        for (long long item = blockIdx.x; item < items; item += gridDim.x) {
            const long long tile = item / p.n_groups;
            const int group = (int)(item % p.n_groups);
            const long long m = tile * 128 + row;
            // window-order rows: destination token of this row (projection scatter: window reverse + un-shift)
            long long mo = m;
            if (OUTMODE == OUT_F32 && p.win_order && m < p.M) mo = win_order_token(p.wo, (uint32_t)m);
            const int c0 = group * p.chunks_per_group;
            const int c1 = min(c0 + p.chunks_per_group, p.n_chunks);
            // fp32 rows, coalesced (p.coal): a thread owns a ROW of the accumulator (TMEM lane), so storing its 64-byte segment directly
            // makes every warp instruction touch 32 lines.  The 32 x 16 block goes through shared memory instead and leaves as 8 rows x
            // 64 contiguous bytes per instruction (lane l: rows 8k + l/4, float4 l%4); the residual arrives in the same mapping.
            const bool coal = OUTMODE == OUT_F32 && STREAM && p.coal;   // the producer flavours (576 threads, 96 registers) keep direct stores
            long long mo4[4];
            bool ok4[4];
            if (OUTMODE == OUT_F32 && STREAM && coal) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int src = 8 * k + (lane >> 2);
                    mo4[k] = __shfl_sync(0xffffffffu, mo, src);
                    ok4[k] = tile * 128 + rb * 32 + src < p.M;
                }
            }
            float* trbuf = reinterpret_cast<float*>(smem + L.tr_off + (uint32_t)(warp - (PW + 2)) * EPI_TR_BYTES);
            for (int c = c0; c < c1; c++, tcount++) {
                const uint32_t acc = tcount & 1u;
                mbar_wait_relaxed(&d_full[acc], (tcount >> 1) & 1u);
                __syncwarp();
                tc_fence_after_sync();
                const uint32_t tlane = tmem_base + acc * acc_stride + ((uint32_t)(rb * 32) << 16);
                const int ncol0 = c * p.NCH;
                // 16-column groups of this warp: eg*16, eg*16 + 32, ...  Software pipeline: the TMEM load of group i+1 and the
                // residual row segment of group i+1 (an L2 / HBM access, scattered in window order) are in flight while group i is
                // converted and stored -- the serial form paid one memory latency per group (most of a wide projection's time).
                int ngr = 0;
                for (int c16 = eg * 16; c16 < p.NCH && ncol0 + c16 < p.N; c16 += 32) ngr++;
                const bool res_vec = OUTMODE == OUT_F32 && p.residual && m < p.M && ((p.ldr & 3) == 0) && ((p.N & 3) == 0) &&
                                     ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0);
                uint32_t rnext[16];
                float4 rsn[4];
                if (ngr > 0) {
                    tmem_ld16_issue(tlane + (uint32_t)(eg * 16), rnext);
                    if (OUTMODE == OUT_F32 && STREAM && coal) {
                        if (p.residual) {
#pragma unroll
                            for (int k = 0; k < 4; k++)
                                rsn[k] = ok4[k] ? *reinterpret_cast<const float4*>(p.residual + mo4[k] * p.ldr + ncol0 + eg * 16 + (lane & 3) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    } else if (res_vec) {
                        const float* rp = p.residual + mo * p.ldr + ncol0 + eg * 16;
#pragma unroll
                        for (int i = 0; i < 4; i++) rsn[i] = (ncol0 + eg * 16 + i * 4 + 4 <= p.N) ? *reinterpret_cast<const float4*>(rp + i * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                for (int gi = 0; gi < ngr; gi++) {
                    const int c16 = eg * 16 + 32 * gi;
                    const int n0 = ncol0 + c16;
                    float v[16];
                    float4 rsc[4];
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(rnext[i]);
#pragma unroll
                    for (int i = 0; i < 4; i++) rsc[i] = rsn[i];
                    if (gi + 1 < ngr) {
                        tmem_ld16_issue(tlane + (uint32_t)(c16 + 32), rnext);
                        if (OUTMODE == OUT_F32 && STREAM && coal) {
                            if (p.residual) {
#pragma unroll
                                for (int k = 0; k < 4; k++)
                                    rsn[k] = ok4[k] ? *reinterpret_cast<const float4*>(p.residual + mo4[k] * p.ldr + n0 + 32 + (lane & 3) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                            }
                        } else if (res_vec) {
                            const float* rp = p.residual + mo * p.ldr + n0 + 32;
#pragma unroll
                            for (int i = 0; i < 4; i++) rsn[i] = (n0 + 32 + i * 4 + 4 <= p.N) ? *reinterpret_cast<const float4*>(rp + i * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                    if (OUTMODE == OUT_F32 && STREAM && coal) {
                        if (p.bias) {
#pragma unroll
                            for (int i = 0; i < 16; i += 4) {
                                const float4 bb = *reinterpret_cast<const float4*>(sbias + n0 + i);
                                v[i] += bb.x; v[i + 1] += bb.y; v[i + 2] += bb.z; v[i + 3] += bb.w;
                            }
                        }
                        if (p.elu) {
#pragma unroll
                            for (int i = 0; i < 16; i++) v[i] = elu_fast(v[i]);
                        }
                        float4* mine = reinterpret_cast<float4*>(trbuf + (uint32_t)lane * EPI_TR_PITCH);
#pragma unroll
                        for (int i = 0; i < 4; i++) mine[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                        __syncwarp();
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            float4 t = *reinterpret_cast<const float4*>(trbuf + (uint32_t)(8 * k + (lane >> 2)) * EPI_TR_PITCH + (uint32_t)(lane & 3) * 4u);
                            if (ok4[k]) {
                                if (p.residual) { t.x += rsc[k].x; t.y += rsc[k].y; t.z += rsc[k].z; t.w += rsc[k].w; }
                                *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + mo4[k] * p.ldo + p.out_col0 + n0 + (lane & 3) * 4) = t;
                            }
                        }
                        __syncwarp();   // the block is read before the next group overwrites it
                        continue;
                    }
                    if (OUTMODE == OUT_TILED && m >= p.M && p.zero_tail) {
                        // rows past M of the last tile: zeros (the weight-gradient GEMM sums over the rows of this tensor)
                        bf16* o = reinterpret_cast<bf16*>(p.out);
                        const int kc = (p.out_col0 + n0) >> 3;
                        *reinterpret_cast<uint4*>(o + (((size_t)tile * p.out_nkc + kc) * 128 + row) * 8) = make_uint4(0u, 0u, 0u, 0u);
                        if (kc + 1 < p.out_nkc) *reinterpret_cast<uint4*>(o + (((size_t)tile * p.out_nkc + kc + 1) * 128 + row) * 8) = make_uint4(0u, 0u, 0u, 0u);
                    }
                    if (m < p.M) {
                        if (p.bias) {
#pragma unroll
                            for (int i = 0; i < 16; i += 4) {
                                const float4 bb = *reinterpret_cast<const float4*>(sbias + n0 + i);
                                v[i] += bb.x; v[i + 1] += bb.y; v[i + 2] += bb.z; v[i + 3] += bb.w;
                            }
                        }
                        if (p.elu) {
#pragma unroll
                            for (int i = 0; i < 16; i++) v[i] = elu_fast(v[i]);
                        }
                        if (OUTMODE == OUT_TILED && p.elu_aux) {
                            // backward of ELU: multiply by ELU'(pre) read off the saved activation a = ELU(pre):
                            // a > 0 -> 1, else e^pre = a + 1   (a003:23, alpha = 1)
                            const bf16* ax = p.elu_aux + (((size_t)tile * p.out_nkc + ((p.out_col0 + n0) >> 3)) * 128 + row) * 8;
                            const uint4 a0 = *reinterpret_cast<const uint4*>(ax);
                            const uint4 a1 = ((p.out_col0 + n0) >> 3) + 1 < p.out_nkc ? *reinterpret_cast<const uint4*>(ax + 1024) : make_uint4(0u, 0u, 0u, 0u);
                            const uint32_t aw[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                            for (int i = 0; i < 8; i++) {
                                const float lo = __uint_as_float(aw[i] << 16), hi = __uint_as_float(aw[i] & 0xffff0000u);
                                v[2 * i] *= lo > 0.f ? 1.f : lo + 1.f;
                                v[2 * i + 1] *= hi > 0.f ? 1.f : hi + 1.f;
                            }
                        }
                        if (OUTMODE == OUT_TILED) {
                            // chunk (tile, kc, r) at ((tile*out_nkc + kc)*128 + r)*8 elements; columns >= N are zero
                            // (zero weights, zero bias, ELU(0) = 0) which is exactly the K padding the next GEMM needs
                            bf16* o = reinterpret_cast<bf16*>(p.out);
                            const int kc = (p.out_col0 + n0) >> 3;
                            uint32_t pk[8];
#pragma unroll
                            for (int i = 0; i < 8; i++) pk[i] = p.out_fp16 ? pack_f16x2(v[2 * i], v[2 * i + 1]) : pack_bf16x2(v[2 * i], v[2 * i + 1]);
                            uint4 lo = make_uint4(pk[0], pk[1], pk[2], pk[3]), hi = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                            *reinterpret_cast<uint4*>(o + (((size_t)tile * p.out_nkc + kc) * 128 + row) * 8) = lo;
                            if (kc + 1 < p.out_nkc) *reinterpret_cast<uint4*>(o + (((size_t)tile * p.out_nkc + kc + 1) * 128 + row) * 8) = hi;
                        } else if (OUTMODE == OUT_BF16) {
                            // 16-bit rows: bf16, or fp16 (out_fp16) for q/k/v whose consumer is the fp32 softmax
                            // kernel, not the tensor core -- 3 more mantissa bits on the attention scores
                            bf16* o = reinterpret_cast<bf16*>(p.out) + m * p.ldo + p.out_col0 + n0;
                            uint32_t pk[8];
#pragma unroll
                            for (int i = 0; i < 8; i++) pk[i] = p.out_fp16 ? pack_f16x2(v[2 * i], v[2 * i + 1]) : pack_bf16x2(v[2 * i], v[2 * i + 1]);
                            if (n0 + 16 <= p.N && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
                                reinterpret_cast<uint4*>(o)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                                reinterpret_cast<uint4*>(o)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                            } else {
                                const uint16_t* h = reinterpret_cast<const uint16_t*>(pk);
                                uint16_t* o16 = reinterpret_cast<uint16_t*>(o);
#pragma unroll
                                for (int i = 0; i < 16; i++) if (n0 + i < p.N) o16[i] = h[i];
                            }
                        } else {
                            float* o = reinterpret_cast<float*>(p.out) + mo * p.ldo + p.out_col0 + n0;
                            const float* rs = p.residual ? p.residual + mo * p.ldr + n0 : nullptr;
                            const bool vec = ((reinterpret_cast<uintptr_t>(o) & 15) == 0) && (!rs || (reinterpret_cast<uintptr_t>(rs) & 15) == 0);
#pragma unroll
                            for (int i = 0; i < 16; i += 4) {
                                if (vec && n0 + i + 4 <= p.N) {
                                    float4 t = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                                    if (rs) {
                                        const float4 rr = res_vec ? rsc[i >> 2] : *reinterpret_cast<const float4*>(rs + i);
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
                __syncwarp();
                if (lane == 0) mbar_arrive(&d_empty[acc]);
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == PW + 1) tmem_dealloc(tmem_base, ncols);
}

// =============================================================================================
// host side
// =============================================================================================
// k-slab width: the largest of {64, 48, 32, 16} dividing Kpad
int tc_gemm_pick_ks(int Kpad) {
    for (int ks : {64, 48, 32, 16})
        if (Kpad % ks == 0) return ks;
    return 16;
}

// chunk width along N (<= 256 so that two accumulators fit the 512 TMEM columns), balanced
void tc_gemm_pick_nchunk(int Ntot, int* NCH, int* n_chunks) {
    int npad = (int)pad16((uint32_t)Ntot);
    int nc = (npad + 255) / 256;
    *NCH = (int)pad16((uint32_t)((npad + nc - 1) / nc));
    *n_chunks = nc;
}

int tc_gemm_plan(TcGemm* p) {
    p->Kpad = (int)pad16((uint32_t)p->K);
    p->KS = tc_gemm_pick_ks(p->Kpad);
    p->n_slabs = p->Kpad / p->KS;
    const bool stream = p->a_mode == AM_TILED;
    SF_CHECK_ARG(p->NCH % 16 == 0 && p->NCH >= 16 && p->NCH <= 256, "tc_gemm: bad n-chunk %d", p->NCH);
    if (!stream) {
        SF_CHECK_ARG(p->Kpad <= TC_MAX_KPAD, "tc_gemm: K=%d exceeds the resident-A limit %d", p->K, TC_MAX_KPAD);
        if (p->a_mode != AM_MERGE) SF_CHECK_ARG(p->K % 4 == 0 && p->lda % 4 == 0, "tc_gemm: fp32 A needs K %% 4 == 0 (K=%d)", p->K);
    } else {
        p->a_nkc = p->Kpad >> 3;
        if (p->a_tile_nkc <= 0) { p->a_tile_nkc = p->a_nkc; p->a_kc0 = 0; }   // else: A is a column range of a wider tiled tensor
    }
    SF_CHECK_ARG(!p->win_order || p->M < 2147483647LL, "tc_gemm: %lld rows exceed the window-order index range", p->M);
    {   // fp32 rows through the shared-memory transposition when every access can be a float4 (SWINFUSE_GEMM_COAL=0: direct stores)
        static const bool coal_on = [] { const char* e = getenv("SWINFUSE_GEMM_COAL"); return !(e && e[0] == '0'); }();
        p->coal = (coal_on && p->out_mode == OUT_F32 && p->a_mode == AM_TILED && (p->N & 15) == 0 && (p->ldo & 3) == 0 && (p->out_col0 & 3) == 0 && aligned16(p->out) &&
                   (!p->residual || ((p->ldr & 3) == 0 && aligned16(p->residual)))) ? 1 : 0;
    }
    const long long m_tiles = (p->M + 127) / 128;
    // spread the n-chunks of one m-tile over several CTAs only when there are too few m-tiles
    int groups = 1;
    if (m_tiles < 296 && p->n_chunks > 1) {
        groups = (int)((296 + m_tiles - 1) / m_tiles);
        if (groups > p->n_chunks) groups = p->n_chunks;
    }
    p->chunks_per_group = (p->n_chunks + groups - 1) / groups;
    p->n_groups = (p->n_chunks + p->chunks_per_group - 1) / p->chunks_per_group;
    // shared memory: A buffers + k-slab ring
    p->NA = stream ? 0 : 2;
    p->NS = 4;
    for (;;) {
        GemmSmem L = gemm_smem_layout(*p);
        if (L.total <= SMEM_LIMIT) break;
        if (p->NS > 2) p->NS--;
        else if (p->NA > 1) { p->NA = 1; p->NS = 4; }
        else { set_error("tc_gemm: tile does not fit shared memory (K=%d, chunk=%d)", p->K, p->NCH); return SF_ERR_UNSUPPORTED; }
    }
    return SF_OK;
}

template <int AMODE, int OUTMODE>
static int launch_t(const TcGemm& p, const char* name, cudaStream_t st) {
    GemmSmem L = gemm_smem_layout(p);
    static DeviceOnce configured;
    if (configured.need()) {
        cudaError_t e = cudaFuncSetAttribute(k_tc_gemm2<AMODE, OUTMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
        if (e != cudaSuccess) { set_error("tc_gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SF_ERR_CUDA; }
        configured.done();
    }
    const long long m_tiles = (p.M + 127) / 128;
    const long long items = m_tiles * p.n_groups;
    const uint32_t ncols = tmem_cols_pow2(2u * (((uint32_t)p.NCH + 31u) & ~31u));
    const int threads = AMODE == AM_TILED ? G_THREADS_STREAM : G_THREADS;
    int per_sm = (int)(SMEM_LIMIT / (L.total + 1024));
    per_sm = std::min(per_sm, (int)(512u / ncols));
    per_sm = std::min(per_sm, 65536 / (threads * (AMODE == AM_TILED ? 66 : 100)));
    per_sm = std::max(1, std::min(per_sm, 4));
    long long grid = (long long)sm_count() * per_sm;
    if (grid > items) grid = items;
    const double abytes = (AMODE == AM_TILED ? 2.0 : 4.0) * (double)p.M * p.K;
    const double obytes = (OUTMODE == OUT_F32 ? 4.0 : 2.0) * (double)p.M * p.N + (p.residual ? 4.0 * (double)p.M * p.N : 0.0);
    ProfScope ps(name, 2.0 * (double)p.M * p.N * p.K, abytes + obytes + 2.0 * (double)p.N * p.K, st);
    k_tc_gemm2<AMODE, OUTMODE><<<(unsigned)grid, threads, L.total, st>>>(p);
    SF_CHECK_LAUNCH("tc_gemm");
    return SF_OK;
}

int launch_tc_gemm(const TcGemm& p, const char* name, cudaStream_t st) {
    if (p.a_mode == AM_F32_LN && p.out_mode == OUT_BF16) return launch_t<AM_F32_LN, OUT_BF16>(p, name, st);
    if (p.a_mode == AM_F32 && p.out_mode == OUT_BF16) return launch_t<AM_F32, OUT_BF16>(p, name, st);
    if (p.a_mode == AM_F32_LN && p.out_mode == OUT_TILED) return launch_t<AM_F32_LN, OUT_TILED>(p, name, st);
    if (p.a_mode == AM_F32 && p.out_mode == OUT_TILED) return launch_t<AM_F32, OUT_TILED>(p, name, st);
    if (p.a_mode == AM_TILED && p.out_mode == OUT_F32) return launch_t<AM_TILED, OUT_F32>(p, name, st);
    if (p.a_mode == AM_TILED && p.out_mode == OUT_BF16) return launch_t<AM_TILED, OUT_BF16>(p, name, st);
    if (p.a_mode == AM_TILED && p.out_mode == OUT_TILED) return launch_t<AM_TILED, OUT_TILED>(p, name, st);
    if (p.a_mode == AM_F32 && p.out_mode == OUT_F32) return launch_t<AM_F32, OUT_F32>(p, name, st);
    if (p.a_mode == AM_MERGE && p.out_mode == OUT_F32) return launch_t<AM_MERGE, OUT_F32>(p, name, st);
    set_error("tc_gemm: unsupported mode combination (%d -> %d)", p.a_mode, p.out_mode);
    return SF_ERR_INVALID;
}

}  // namespace sf

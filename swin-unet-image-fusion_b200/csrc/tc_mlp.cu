// Fused MLP on tcgen05 for the narrow stages (a003:21-50):  out = x + W2 ELU(W1 LN(x) + b1) + b2.
//
// For C <= 64 the two GEMMs are HBM-bound and the 4x hidden activation dominates their traffic
// (per token: 4C B in, 8C B hidden out, 8C B hidden in, 4C B residual, 4C B out).  Here the hidden
// activation never leaves the SM: a persistent CTA per SM keeps both weight matrices resident in
// shared memory and pipelines, tile (128 tokens) by tile, over double-buffered stages
//
//   warp  17     loader       one cp.async.bulk per tile: the 128 fp32 rows (contiguous) -> a ring of raw tiles in
//                             shared memory, several tiles ahead (HBM latency is covered by the ring depth, not
//                             by registers); with a deep ring the tile is kept until the output stage has
//                             taken the residual from it, so x is read from HBM exactly once
//   warps 0-3    producers    raw row (one thread per row) -> LayerNorm -> bf16 A1 in the UMMA layout
//   warp  4      MMA issuer   D1[128 x Hpad] = A1 W1^T ;  D2[128 x Cpad] = A2 W2^T   (tcgen05.mma, TMEM)
//   warps 5-12   hidden       tcgen05.ld D1 -> +b1 -> ELU -> bf16 -> A2 (shared memory, UMMA layout)
//   warps 13-16  output       tcgen05.ld D2 -> +b2 + residual -> fp32 rows
//
// mbarrier pipelines: A1 full/empty, D1 full/empty, A2 full/empty, D2 full/empty (all two deep), so
// the producers run one tile ahead, GEMM1 of tile t+1 overlaps the ELU stage of tile t, and GEMM2 of
// tile t overlaps the output stage of tile t-1.
#include <type_traits>
#include <cstdlib>
#include "bf16_kernels.cuh"
#include "tc_common.cuh"

namespace sf {
using namespace tc;

static constexpr uint32_t LBO_A = lbo_padded(128);   // thread-written A operands: 2064 B between k-chunks
static constexpr uint32_t SBO_M = 128;
static constexpr int M_THREADS = 18 * 32;
static constexpr int M_LOADER = 17;                  // loader warp
static constexpr int M_MAX_STAGES = 8;
static constexpr int M_RES_MIN_STAGES = 5;           // ring deep enough to hold a tile until its output stage: residual from the ring
static constexpr int M_PW = 4;                       // producer warps; then 1 MMA warp, 8 hidden warps, 4 output warps
static constexpr size_t M_SMEM_LIMIT = 227 * 1024;
__host__ __device__ static inline uint32_t al128(uint32_t v) { return (v + 127) & ~127u; }
__device__ __forceinline__ void mbar_arrive1(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct MlpSmem { uint32_t w1, w2, a1[2], a2[2], b1, b2, bars, ring, tile_bytes, nstage, total; };
__host__ __device__ static inline MlpSmem mlp_smem_layout(int Kpad, int Hpad, int N2, int C, int max_stages = M_MAX_STAGES) {
    MlpSmem s{};
    uint32_t o = 0;
    s.w1 = o; o += al128((uint32_t)Hpad * Kpad * 2);
    s.w2 = o; o += al128((uint32_t)N2 * Hpad * 2);
    for (int i = 0; i < 2; i++) { s.a1[i] = o; o += al128((uint32_t)(Kpad >> 3) * LBO_A); }
    for (int i = 0; i < 2; i++) { s.a2[i] = o; o += al128((uint32_t)(Hpad >> 3) * LBO_A); }
    s.b1 = o; o += al128((uint32_t)Hpad * 4);
    s.b2 = o; o += al128((uint32_t)N2 * 4);
    s.bars = o; o += 384;
    s.ring = o;
    s.tile_bytes = al128(128u * (uint32_t)C * 4u);
    const uint32_t room = o < (uint32_t)M_SMEM_LIMIT ? (uint32_t)M_SMEM_LIMIT - o : 0u;
    s.nstage = room / s.tile_bytes;
    if (s.nstage > (uint32_t)max_stages) s.nstage = max_stages;
    s.total = o + s.nstage * s.tile_bytes;
    return s;
}

// fp32 row (one thread per row) -> registers / registers -> [LayerNorm] -> bf16 UMMA chunks
template <int NF>
__device__ __forceinline__ void mlp_load_row(float4 (&v)[NF], const uint8_t* raw_tile, int r, bool rowok, int C) {
    const int nf4 = C >> 2;
    const float4* src = reinterpret_cast<const float4*>(raw_tile + (size_t)r * C * 4);
#pragma unroll
    for (int i = 0; i < NF; i++) v[i] = (rowok && i < nf4) ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
}
template <int NF>
__device__ __forceinline__ void mlp_finish_row(uint8_t* sA, float4 (&v)[NF], int r, int C, int Kpad, const float* __restrict__ g,
                                               const float* __restrict__ b, float eps) {
    const int nf4 = C >> 2, nkc = Kpad >> 3;
    if (g) {
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < NF; i++) { s0 += v[i].x + v[i].y; s1 += v[i].z + v[i].w; }
        const float invk = 1.f / (float)C;
        const float mean = (s0 + s1) * invk;
        float q0 = 0.f, q1 = 0.f;
#pragma unroll
        for (int i = 0; i < NF; i++) {
            if (i < nf4) {
                float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
                q0 += dx * dx + dy * dy; q1 += dz * dz + dw * dw;
            }
        }
        const float rstd = rsqrtf((q0 + q1) * invk + eps);
#pragma unroll
        for (int i = 0; i < NF; i++) {
            if (i < nf4) {
                float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i), bb = __ldg(reinterpret_cast<const float4*>(b) + i);
                v[i].x = (v[i].x - mean) * rstd * gg.x + bb.x;
                v[i].y = (v[i].y - mean) * rstd * gg.y + bb.y;
                v[i].z = (v[i].z - mean) * rstd * gg.z + bb.z;
                v[i].w = (v[i].w - mean) * rstd * gg.w + bb.w;
            }
        }
    }
#pragma unroll
    for (int c = 0; c < NF / 2; c++) {
        if (c < nkc) {
            uint4 pk = make_uint4(pack_bf16x2(v[2 * c].x, v[2 * c].y), pack_bf16x2(v[2 * c].z, v[2 * c].w),
                                  pack_bf16x2(v[2 * c + 1].x, v[2 * c + 1].y), pack_bf16x2(v[2 * c + 1].z, v[2 * c + 1].w));
            *reinterpret_cast<uint4*>(sA + (uint32_t)c * LBO_A + (uint32_t)r * 16) = pk;
        }
    }
}

__global__ void __launch_bounds__(M_THREADS, 1) k_tc_mlp(TcMlp p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Kpad = p.Cpad, Hpad = p.Hpad, N2 = p.Cpad;
    const MlpSmem L = mlp_smem_layout(Kpad, Hpad, N2, p.C, p.max_stages);
    const uint32_t NS = L.nstage;
    // the residual is x itself (a003 through a004:29-38) and the ring is deep enough: take it from the ring
    const bool ring_res = p.residual == p.x && NS >= (uint32_t)M_RES_MIN_STAGES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
    uint64_t* a1_full = bars;        // [2]
    uint64_t* a1_empty = bars + 2;   // [2]
    uint64_t* d1_full = bars + 4;
    uint64_t* d1_empty = bars + 6;
    uint64_t* a2_full = bars + 8;
    uint64_t* a2_empty = bars + 10;
    uint64_t* d2_full = bars + 12;
    uint64_t* d2_empty = bars + 14;
    uint64_t* w_full = bars + 16;
    uint64_t* x_full = bars + 17;    // [M_MAX_STAGES]
    uint64_t* x_empty = bars + 25;   // [M_MAX_STAGES]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 34);
    const uint32_t d1_stride = ((uint32_t)Hpad + 31u) & ~31u, d2_stride = ((uint32_t)N2 + 31u) & ~31u;
    const uint32_t ncols = tmem_cols_pow2(2u * d1_stride + 2u * d2_stride);
    const long long tiles = (p.M + 127) / 128;
    const long long my_tiles = blockIdx.x < tiles ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const uint32_t w1_bytes = (uint32_t)Hpad * Kpad * 2, w2_bytes = (uint32_t)N2 * Hpad * 2;

    if (tid == M_PW * 32) {
        for (int i = 0; i < 2; i++) {
            mbar_init(&a1_full[i], 128); mbar_init(&a1_empty[i], 1);
            mbar_init(&d1_full[i], 1); mbar_init(&d1_empty[i], 8);
            mbar_init(&a2_full[i], 8); mbar_init(&a2_empty[i], 1);
            mbar_init(&d2_full[i], 1); mbar_init(&d2_empty[i], 4);
        }
        mbar_init(w_full, 1);
        for (uint32_t i = 0; i < NS; i++) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], ring_res ? 128 + 4 : 128); }
        fence_mbar_init();
        // both weight matrices stay in shared memory for the life of the CTA
        mbar_arrive_expect_tx(w_full, w1_bytes + w2_bytes);
        bulk_g2s(smem + L.w1, p.W1p, w1_bytes, w_full);
        bulk_g2s(smem + L.w2, p.W2p, w2_bytes, w_full);
    }
    if (warp == M_PW + 1) tmem_alloc(tmem_slot, ncols);
    {
        float* sb1 = reinterpret_cast<float*>(smem + L.b1);
        float* sb2 = reinterpret_cast<float*>(smem + L.b2);
        for (int i = tid; i < Hpad; i += M_THREADS) sb1[i] = __ldg(p.b1 + i);
        for (int i = tid; i < N2; i += M_THREADS) sb2[i] = __ldg(p.b2 + i);
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const float* sb1 = reinterpret_cast<const float*>(smem + L.b1);
    const float* sb2 = reinterpret_cast<const float*>(smem + L.b2);

    if (warp < M_PW) {
        // ------------------------------ producers ---------------------------------------------------------
        const int r = tid;
        auto run = [&](auto tag) {
            constexpr int NF = decltype(tag)::value;
            for (long long t = 0; t < my_tiles; t++) {
                const long long tile = blockIdx.x + t * gridDim.x;
                const uint32_t st = (uint32_t)(t % NS), sp = (uint32_t)((t / NS) & 1);
                float4 v[NF];
                mbar_wait_relaxed(&x_full[st], sp);
                mlp_load_row<NF>(v, smem + L.ring + st * L.tile_bytes, r, tile * 128 + r < p.M, p.C);
                const uint32_t b = (uint32_t)t & 1u;
                mbar_wait_relaxed(&a1_empty[b], (((uint32_t)t >> 1) & 1u) ^ 1u);
                mlp_finish_row<NF>(smem + L.a1[b], v, r, p.C, Kpad, p.ln_g, p.ln_b, p.eps);
                fence_async_smem();
                mbar_arrive1(&a1_full[b]);
                // release the ring slot only now: the row has been consumed, so its ld.shared have returned
                // (an arrive issued right behind the loads can let the loader's bulk copy overwrite the slot
                // while they are still in flight -- observed as a data race at C = 48)
                mbar_arrive1(&x_empty[st]);
            }
        };
        if (Kpad <= 32) run(std::integral_constant<int, 8>{});
        else run(std::integral_constant<int, 16>{});
    } else if (warp == M_LOADER) {
        // ------------------------------ loader: raw fp32 tiles -> ring ------------------------------------------
        if (lane == 0) {
            for (long long t = 0; t < my_tiles; t++) {
                const long long tile = blockIdx.x + t * gridDim.x;
                const uint32_t st = (uint32_t)(t % NS), sp = (uint32_t)((t / NS) & 1);
                mbar_wait(&x_empty[st], sp ^ 1u);
                const long long rows = (p.M - tile * 128) < 128 ? (p.M - tile * 128) : 128;
                const uint32_t bytes = (uint32_t)rows * (uint32_t)p.C * 4u;
                mbar_arrive_expect_tx(&x_full[st], bytes);
                bulk_g2s(smem + L.ring + st * L.tile_bytes, p.x + tile * 128 * p.C, bytes, &x_full[st]);
            }
        }
    } else if (warp == M_PW) {
        // ------------------------------ MMA issuer --------------------------------------------------------
        if (lane == 0 && my_tiles > 0) {
            mbar_wait(w_full, 0);
            tc_fence_after_sync();
            const uint32_t idesc1 = make_idesc_bf16(128, (uint32_t)Hpad), idesc2 = make_idesc_bf16(128, (uint32_t)N2);
            const uint32_t lbo_w1 = lbo_dense((uint32_t)Hpad), lbo_w2 = lbo_dense((uint32_t)N2);
            const uint32_t w1 = smem_u32(smem + L.w1), w2 = smem_u32(smem + L.w2);
            auto gemm1 = [&](long long t) {
                const uint32_t b = (uint32_t)t & 1u, par = ((uint32_t)t >> 1) & 1u;
                mbar_wait(&a1_full[b], par);
                mbar_wait(&d1_empty[b], par ^ 1u);
                tc_fence_after_sync();
                const uint32_t a = smem_u32(smem + L.a1[b]);
                for (int ks = 0; ks < (Kpad >> 4); ks++)
                    umma_bf16(tmem_base + b * d1_stride, make_smem_desc(a + (uint32_t)ks * 2u * LBO_A, LBO_A, SBO_M),
                              make_smem_desc(w1 + (uint32_t)ks * 2u * lbo_w1, lbo_w1, SBO_M), idesc1, ks > 0);
                umma_commit(&a1_empty[b]);
                umma_commit(&d1_full[b]);
            };
            auto gemm2 = [&](long long t) {
                const uint32_t b = (uint32_t)t & 1u, par = ((uint32_t)t >> 1) & 1u;
                mbar_wait(&a2_full[b], par);
                mbar_wait(&d2_empty[b], par ^ 1u);
                tc_fence_after_sync();
                const uint32_t a = smem_u32(smem + L.a2[b]);
                for (int ks = 0; ks < (Hpad >> 4); ks++)
                    umma_bf16(tmem_base + 2u * d1_stride + b * d2_stride, make_smem_desc(a + (uint32_t)ks * 2u * LBO_A, LBO_A, SBO_M),
                              make_smem_desc(w2 + (uint32_t)ks * 2u * lbo_w2, lbo_w2, SBO_M), idesc2, ks > 0);
                umma_commit(&a2_empty[b]);
                umma_commit(&d2_full[b]);
            };
            gemm1(0);
            for (long long t = 0; t < my_tiles; t++) {
                if (t + 1 < my_tiles) gemm1(t + 1);
                gemm2(t);
            }
        }
    } else if (warp >= M_PW + 1 && warp < M_PW + 1 + 8) {
        // ------------------------------ hidden stage: D1 -> +b1 -> ELU -> bf16 A2 -----------------------------
        const int rb = warp & 3, eg = (warp - (M_PW + 1)) >> 2;
        const int row = rb * 32 + lane;
        for (long long t = 0; t < my_tiles; t++) {
            const uint32_t b = (uint32_t)t & 1u, par = ((uint32_t)t >> 1) & 1u;
            mbar_wait_relaxed(&d1_full[b], par);
            mbar_wait_relaxed(&a2_empty[b], par ^ 1u);
            __syncwarp();
            tc_fence_after_sync();
            const uint32_t tlane = tmem_base + b * d1_stride + ((uint32_t)(rb * 32) << 16);
            uint8_t* sA2 = smem + L.a2[b];
            for (int c16 = eg * 16; c16 < Hpad; c16 += 32) {
                float v[16];
                tmem_ld16(tlane + (uint32_t)c16, v);
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const float4 bb = *reinterpret_cast<const float4*>(sb1 + c16 + i);
                    v[i] += bb.x; v[i + 1] += bb.y; v[i + 2] += bb.z; v[i + 3] += bb.w;
                }
#pragma unroll
                for (int i = 0; i < 16; i++) v[i] = elu_fast(v[i]);
                uint8_t* dst = sA2 + (uint32_t)(c16 >> 3) * LBO_A + (uint32_t)row * 16;
                *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
                *reinterpret_cast<uint4*>(dst + LBO_A) = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
            }
            fence_async_smem();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) { mbar_arrive1(&d1_empty[b]); mbar_arrive1(&a2_full[b]); }
        }
    } else if (warp >= M_PW + 9 && warp < M_PW + 13) {
        // ------------------------------ output stage: D2 + b2 + residual -> fp32 rows --------------------------
        const int rb = warp & 3;
        const int row = rb * 32 + lane;
        for (long long t = 0; t < my_tiles; t++) {
            const long long tile = blockIdx.x + t * gridDim.x;
            const long long m = tile * 128 + row;
            const uint32_t b = (uint32_t)t & 1u, par = ((uint32_t)t >> 1) & 1u;
            // the residual row: from the ring (x itself, read once from HBM) or in flight from global while the
            // accumulator is awaited
            float4 rr[16];
            const bool rowok = m < p.M;
            const int nf4 = p.C >> 2;
            if (ring_res) {
                const uint32_t st = (uint32_t)(t % NS), sp = (uint32_t)((t / NS) & 1);
                mbar_wait_relaxed(&x_full[st], sp);   // completed long ago; makes the bulk-copied bytes visible to this thread
                const float4* rs = reinterpret_cast<const float4*>(smem + L.ring + st * L.tile_bytes + (size_t)row * p.C * 4);
#pragma unroll
                for (int i = 0; i < 16; i++) rr[i] = (rowok && i < nf4) ? rs[i] : make_float4(0.f, 0.f, 0.f, 0.f);
            } else if (p.residual) {
                const float4* rs = reinterpret_cast<const float4*>(p.residual + m * p.C);
#pragma unroll
                for (int i = 0; i < 16; i++) rr[i] = (rowok && i < nf4) ? rs[i] : make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
#pragma unroll
                for (int i = 0; i < 16; i++) rr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            mbar_wait_relaxed(&d2_full[b], par);
            __syncwarp();
            tc_fence_after_sync();
            const uint32_t tlane = tmem_base + 2u * d1_stride + b * d2_stride + ((uint32_t)(rb * 32) << 16);
            if (ring_res && p.bulk_out) {
                // The tile's rows are contiguous in HBM: each row is completed IN PLACE in its ring slot (x + MLP(x)) and the warp's 32
                // rows leave as one bulk copy -- no thread-per-row stores (32 lines per instruction at a 4C-byte pitch).
                float4* slot = reinterpret_cast<float4*>(smem + L.ring + (uint32_t)(t % NS) * L.tile_bytes + (size_t)row * p.C * 4);
#pragma unroll
                for (int g16 = 0; g16 < 4; g16++) {
                    const int c16 = g16 * 16;
                    if (c16 < p.C) {   // uniform
                        float v[16];
                        tmem_ld16(tlane + (uint32_t)c16, v);
#pragma unroll
                        for (int i = 0; i < 16; i += 4) {
                            if (c16 + i + 4 <= p.C) {
                                const float4 bb = *reinterpret_cast<const float4*>(sb2 + c16 + i);
                                const float4 q4 = rr[g16 * 4 + (i >> 2)];
                                slot[(c16 + i) >> 2] = make_float4(v[i] + bb.x + q4.x, v[i + 1] + bb.y + q4.y, v[i + 2] + bb.z + q4.z, v[i + 3] + bb.w + q4.w);
                            }
                        }
                    }
                }
                tc_fence_before_sync();
                fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive1(&d2_empty[b]);
                    const long long r0 = tile * 128 + rb * 32;
                    const long long nrow = p.M - r0 < 32 ? p.M - r0 : 32;
                    if (nrow > 0) {
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p.out + r0 * p.C),
                                     "r"(smem_u32(smem + L.ring + (uint32_t)(t % NS) * L.tile_bytes + (uint32_t)(rb * 32) * (uint32_t)p.C * 4u)),
                                     "r"((uint32_t)nrow * (uint32_t)p.C * 4u)
                                     : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the slot is refilled once its rows have been read
                    }
                    mbar_arrive1(&x_empty[(uint32_t)(t % NS)]);
                }
                continue;
            }
#pragma unroll
            for (int g16 = 0; g16 < 4; g16++) {
                const int c16 = g16 * 16;
                if (c16 < p.C) {   // uniform
                    float v[16];
                    tmem_ld16(tlane + (uint32_t)c16, v);
                    if (rowok) {
                        float* o = p.out + m * p.C + c16;
#pragma unroll
                        for (int i = 0; i < 16; i += 4) {
                            if (c16 + i + 4 <= p.C) {
                                const float4 bb = *reinterpret_cast<const float4*>(sb2 + c16 + i);
                                const float4 q4 = rr[g16 * 4 + (i >> 2)];
                                *reinterpret_cast<float4*>(o + i) =
                                    make_float4(v[i] + bb.x + q4.x, v[i + 1] + bb.y + q4.y, v[i + 2] + bb.z + q4.z, v[i + 3] + bb.w + q4.w);
                            }
                        }
                    }
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive1(&d2_empty[b]);
                if (ring_res) mbar_arrive1(&x_empty[(uint32_t)(t % NS)]);   // residual consumed (stores issued): slot free
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == M_PW + 1) tmem_dealloc(tmem_base, ncols);
}

bool tc_mlp_supported(int C, int hidden) {
    if (C % 4 != 0 || C > 64 || hidden < 1) return false;
    const int Kpad = (int)pad16((uint32_t)C), Hpad = (int)pad16((uint32_t)hidden);
    if (Hpad > 256) return false;
    const uint32_t cols = 2u * (((uint32_t)Hpad + 31u) & ~31u) + 2u * (((uint32_t)Kpad + 31u) & ~31u);
    const MlpSmem L = mlp_smem_layout(Kpad, Hpad, Kpad, C);
    return cols <= 512 && L.nstage >= 2 && L.total <= M_SMEM_LIMIT;
}

int launch_tc_mlp(const TcMlp& t_in, cudaStream_t st) {
    TcMlp t = t_in;
    {
        static const bool bulk = [] { const char* e = getenv("SWINFUSE_MLP_BULK_OUT"); return !(e && e[0] == '0'); }();
        t.bulk_out = (bulk && ((reinterpret_cast<uintptr_t>(t.out) & 15) == 0)) ? 1 : 0;
    }
    SF_CHECK_ARG(tc_mlp_supported(t.C, t.hidden), "tc_mlp: unsupported shape C=%d hidden=%d", t.C, t.hidden);
    const MlpSmem L = mlp_smem_layout(t.Cpad, t.Hpad, t.Cpad, t.C, t.max_stages);
    static DeviceOnce configured;
    if (configured.need()) {
        cudaError_t e = cudaFuncSetAttribute(k_tc_mlp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)M_SMEM_LIMIT);
        if (e != cudaSuccess) { set_error("tc_mlp: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SF_ERR_CUDA; }
        configured.done();
    }
    const long long tiles = (t.M + 127) / 128;
    long long grid = sm_count();
    if (grid > tiles) grid = tiles;
    // algorithmic work: two GEMMs; bytes: x in, residual in (when it is another tensor), out, weights once
    ProfScope ps(prof_name("tc_mlp_fused_c%d", t.C), 4.0 * (double)t.M * t.C * t.hidden,
                 4.0 * (double)t.M * t.C * (t.residual && t.residual != t.x ? 3.0 : 2.0) + 4.0 * t.C * t.hidden, st);
    k_tc_mlp<<<(unsigned)grid, M_THREADS, L.total, st>>>(t);
    SF_CHECK_LAUNCH("tc_mlp");
    return SF_OK;
}

}  // namespace sf

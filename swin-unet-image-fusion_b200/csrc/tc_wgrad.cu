// Weight-gradient GEMM of the backward pass on tcgen05 (autograd of the 1x1 convolutions / linears of a001, a003, a011):
//
//     Wg[N][K] += sum over token rows t of  G[t][n] * A[t][k]            (dW = dY^T X;  a016:164 loss.backward())
//     bg[N]    += sum over t of G[t][n]                                  (bias gradient, same pass)
//
// Both operands arrive as bf16 tensors in the UMMA-tiled layout [tile of 128 tokens][8-feature chunk][token][8]
// (the layout every forward GEMM of this library already produces / consumes).  Seen from THIS product the
// reduction runs over tokens, i.e. both operands are "MN-major": a chunk-column of a tile is exactly a stack of
// 8-token x 8-feature core matrices whose 8 features are contiguous, so the same bytes are handed to tcgen05.mma
// with the major bits of the instruction descriptor set and LBO / SBO swapped -- no transposing pass, no thread
// touches the operands: a loader lane bulk-copies (m-block of G, n-block of A) slabs of a token tile into a
// two-deep shared-memory ring, an issuer lane runs 8 K-steps (16 tokens each) per tile into ONE TMEM accumulator
// that lives across all token tiles of the work item, and four epilogue warps add it to Wg once per item
// (fp32 red.global.add: the token range is split over CTAs).  Bias gradient: one more N=16 MMA per K-step against a
// constant all-ones operand.
#include <algorithm>
#include "bf16_kernels.cuh"
#include "tc_common.cuh"

namespace sf {
using namespace tc;

static constexpr int WG_THREADS = 192;         // warp 0: loader lane, warp 1: MMA issuer lane, warps 2-5: epilogue
static constexpr uint32_t WG_CHUNK = 2048;     // bytes of one 8-feature chunk of a 128-token tile
static constexpr uint32_t WG_GBYTES = 16 * WG_CHUNK;   // m-block of G: 128 features
static constexpr size_t WG_SMEM_LIMIT = 227 * 1024;

struct TcWgrad {
    const bf16* G; int g_nkc;     // gradient rows, tiled; g_nkc chunks of this operand per tile (= pad16(N) / 8)
    const bf16* A; int a_nkc;     // activation rows, tiled (= pad16(K) / 8 chunks)
    int g_tile_nkc, g_kc0;        // G is the chunk range [g_kc0, g_kc0 + g_nkc) of a tiled tensor with g_tile_nkc chunks per tile
    int a_tile_nkc, a_kc0;        // same for A
    long long M;                  // token rows
    int N, K;
    int each;                     // rows [s*each, (s+1)*each) of the product go to Wg[s] / bias_grad[s] (stacked weight matrices)
    float* Wg[3];                 // each [each][K] fp32, accumulated
    float* bias_grad[3];          // each [each] fp32, accumulated; or null
    int m_blocks, n_blocks, nbc;  // nbc: chunks of A per n-block (<= 32)
    int splits;                   // token-range splits
    long long tiles, tiles_per_split;
    uint32_t bias_col, ncols;     // TMEM: column of the bias-gradient accumulator, columns allocated
    int vec4;                     // 16-byte vector reductions into Wg
};

// MN-major, SWIZZLE_NONE operand: `lbo` = bytes between 8-token groups (K direction), `sbo` = bytes between 8-feature
// chunks (M / N direction) -- the roles the two fields have for a K-major operand, exchanged
__device__ __forceinline__ uint64_t wg_desc(uint32_t addr) { return make_smem_desc(addr, 128u, WG_CHUNK); }
__host__ __device__ constexpr uint32_t wg_idesc(uint32_t M, uint32_t N) {
    return make_idesc_bf16(M, N) | (1u << 15) | (1u << 16);   // a_major = b_major = MN
}
// Wg[0..3] += v  (one 16-byte reduction at L2 instead of four)
__device__ __forceinline__ void wg_red4(float* dst, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void wg_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(WG_THREADS, 1) k_tc_wgrad(TcWgrad p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t abytes = (uint32_t)p.nbc * WG_CHUNK;
    const uint32_t stage_bytes = WG_GBYTES + abytes;
    uint8_t* ones = smem + 2 * stage_bytes;                       // [2 token groups][16 columns][8] bf16 = 512 B of 1.0
    uint64_t* bars = reinterpret_cast<uint64_t*>(ones + 512);
    uint64_t* full = bars;        // [2]
    uint64_t* empty = bars + 2;   // [2]
    uint64_t* d_full = bars + 4;
    uint64_t* d_empty = bars + 5;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
    const bool has_bias = p.bias_grad[0] != nullptr;
    const uint32_t ncols = p.ncols;
    const long long items = (long long)p.m_blocks * p.n_blocks * p.splits;

    if (tid == 0) {
        for (int i = 0; i < 2; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(d_full, 1); mbar_init(d_empty, 4);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, ncols);
    {
        // chunks of an m-block beyond the tensor's last chunk are never copied: they must read as zeros
        uint4* z = reinterpret_cast<uint4*>(smem);
        for (uint32_t i = tid; i < (2 * stage_bytes) >> 4; i += WG_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
        uint32_t* o = reinterpret_cast<uint32_t*>(ones);
        for (int i = tid; i < 128; i += WG_THREADS) o[i] = 0x3f803f80u;   // bf16 1.0 pairs
    }
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------ loader ------------------------------------------------------------------------
        if (lane == 0) {
            uint32_t cnt = 0;
            for (long long item = blockIdx.x; item < items; item += gridDim.x) {
                const int sp = (int)(item % p.splits);
                const int nb = (int)((item / p.splits) % p.n_blocks), mb = (int)(item / ((long long)p.splits * p.n_blocks));
                const long long t0 = sp * p.tiles_per_split, t1 = std::min(p.tiles, t0 + p.tiles_per_split);
                const uint32_t gch = (uint32_t)std::min(16, p.g_nkc - mb * 16), ach = (uint32_t)std::min(p.nbc, p.a_nkc - nb * p.nbc);
                for (long long t = t0; t < t1; t++, cnt++) {
                    const uint32_t st = cnt & 1u;
                    mbar_wait(&empty[st], ((cnt >> 1) & 1u) ^ 1u);
                    uint8_t* dst = smem + st * stage_bytes;
                    mbar_arrive_expect_tx(&full[st], (gch + ach) * WG_CHUNK);
                    bulk_g2s(dst, p.G + ((size_t)t * p.g_tile_nkc + p.g_kc0 + (size_t)mb * 16) * 1024, gch * WG_CHUNK, &full[st]);
                    bulk_g2s(dst + WG_GBYTES, p.A + ((size_t)t * p.a_tile_nkc + p.a_kc0 + (size_t)nb * p.nbc) * 1024, ach * WG_CHUNK, &full[st]);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer ----------------------------------------------------------------------
        if (lane == 0) {
            uint32_t cnt = 0, icount = 0;
            const uint32_t idesc = wg_idesc(128, (uint32_t)p.nbc * 8u), idesc_b = wg_idesc(128, 16);
            const uint32_t ones_a = smem_u32(ones);
            for (long long item = blockIdx.x; item < items; item += gridDim.x, icount++) {
                const int sp = (int)(item % p.splits);
                const int nb = (int)((item / p.splits) % p.n_blocks);
                const long long t0 = sp * p.tiles_per_split, t1 = std::min(p.tiles, t0 + p.tiles_per_split);
                mbar_wait(d_empty, (icount & 1u) ^ 1u);
                tc_fence_after_sync();
                bool first = true;
                for (long long t = t0; t < t1; t++, cnt++) {
                    const uint32_t st = cnt & 1u;
                    mbar_wait(&full[st], (cnt >> 1) & 1u);
                    tc_fence_after_sync();
                    const uint32_t gbase = smem_u32(smem + st * stage_bytes), abase = gbase + WG_GBYTES;
                    for (uint32_t ks = 0; ks < 8; ks++) {   // 16 tokens per step: two 8-token groups of 128 B
                        const uint64_t dg = wg_desc(gbase + ks * 256u);
                        umma_bf16(tmem_base, dg, wg_desc(abase + ks * 256u), idesc, !first);
                        if (has_bias && nb == 0) umma_bf16(tmem_base + p.bias_col, dg, make_smem_desc(ones_a, 128u, 256u), idesc_b, !first);
                        first = false;
                    }
                    umma_commit(&empty[st]);
                }
                umma_commit(d_full);
            }
        }
    } else {
        // ------------------------------ epilogue: accumulator -> Wg (+=) ------------------------------------------------------
        const int rb = warp & 3;
        const int rowl = rb * 32 + lane;
        uint32_t icount = 0;
        for (long long item = blockIdx.x; item < items; item += gridDim.x, icount++) {
            const int nb = (int)((item / p.splits) % p.n_blocks), mb = (int)(item / ((long long)p.splits * p.n_blocks));
            const int n = mb * 128 + rowl;
            const int k0 = nb * p.nbc * 8;
            const int kcols = std::min(p.nbc * 8, p.K - k0);
            mbar_wait_relaxed(d_full, icount & 1u);
            __syncwarp();
            tc_fence_after_sync();
            const uint32_t tl = tmem_base + ((uint32_t)(rb * 32) << 16);
            for (int c16 = 0; c16 < kcols; c16 += 16) {
                float v[16];
                tmem_ld16(tl + (uint32_t)c16, v);
                if (n < p.N) {
                    const int src = n / p.each;
                    float* dst = p.Wg[src] + (size_t)(n - src * p.each) * p.K + k0 + c16;
                    if (p.vec4) {   // K % 4 == 0 and Wg 16-byte aligned: every group of four columns is one aligned vector
#pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            if (c16 + i < kcols) wg_red4(dst + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; i++)
                            if (c16 + i < kcols) atomicAdd(dst + i, v[i]);
                    }
                }
            }
            if (has_bias && nb == 0) {
                float v[16];
                tmem_ld16(tl + p.bias_col, v);
                if (n < p.N) { const int src = n / p.each; atomicAdd(p.bias_grad[src] + (n - src * p.each), v[0]); }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) wg_arrive(d_empty);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, ncols);
}

int launch_tc_wgrad(const bf16* G, const bf16* A, float* Wg, float* bias_grad, long long M, int N, int K, const char* name, cudaStream_t st) {
    TcWgradArgs a{};
    a.G = G; a.A = A; a.M = M; a.N = N; a.K = K; a.nout = 1; a.Wg[0] = Wg; a.bias_grad[0] = bias_grad;
    return launch_tc_wgrad_ex(a, name, st);
}

int launch_tc_wgrad_ex(const TcWgradArgs& a, const char* name, cudaStream_t st) {
    const bf16* G = a.G; const bf16* A = a.A; const long long M = a.M; const int N = a.N, K = a.K;
    float* Wg = a.Wg[0]; float* bias_grad = a.bias_grad[0];
    SF_CHECK_ARG(G && A && Wg && M > 0 && N > 0 && K > 0 && a.nout >= 1 && a.nout <= 3 && N % a.nout == 0, "tc_wgrad: bad arguments");
    TcWgrad p{};
    p.G = G; p.A = A; p.M = M; p.N = N; p.K = K;
    p.each = N / a.nout;
    for (int i = 0; i < 3; i++) { p.Wg[i] = i < a.nout ? a.Wg[i] : nullptr; p.bias_grad[i] = i < a.nout ? a.bias_grad[i] : nullptr; }
    SF_CHECK_ARG(!bias_grad || a.nout == 1 || (a.bias_grad[1] && (a.nout < 3 || a.bias_grad[2])), "tc_wgrad: bias gradients are all or none");
    p.g_nkc = (int)pad16((uint32_t)N) / 8;
    p.a_nkc = (int)pad16((uint32_t)K) / 8;
    p.g_tile_nkc = a.g_tile_nkc > 0 ? a.g_tile_nkc : p.g_nkc; p.g_kc0 = a.g_tile_nkc > 0 ? a.g_kc0 : 0;
    p.a_tile_nkc = a.a_tile_nkc > 0 ? a.a_tile_nkc : p.a_nkc; p.a_kc0 = a.a_tile_nkc > 0 ? a.a_kc0 : 0;
    p.m_blocks = (p.g_nkc + 15) / 16;
    p.n_blocks = (p.a_nkc + 31) / 32;
    p.nbc = (p.a_nkc + p.n_blocks - 1) / p.n_blocks;
    p.nbc = (p.nbc + 1) & ~1;                       // N of the MMA is a multiple of 16
    p.n_blocks = (p.a_nkc + p.nbc - 1) / p.nbc;
    p.tiles = (M + 127) / 128;
    const long long blocks = (long long)p.m_blocks * p.n_blocks;
    // token-range splits: about one work item per SM, and at least eight token tiles per item so that the item's epilogue
    // (128 x nbc*8 reductions into Wg) is paid for by its streaming phase
    long long splits = ((long long)sm_count() + blocks - 1) / blocks;
    splits = std::max(1LL, std::min(splits, (p.tiles + 7) / 8));
    p.vec4 = (K % 4 == 0);
    for (int i = 0; i < a.nout; i++) p.vec4 = p.vec4 && ((reinterpret_cast<uintptr_t>(a.Wg[i]) & 15) == 0);
    p.tiles_per_split = (p.tiles + splits - 1) / splits;
    p.splits = (int)((p.tiles + p.tiles_per_split - 1) / p.tiles_per_split);
    const uint32_t stage_bytes = WG_GBYTES + (uint32_t)p.nbc * WG_CHUNK;
    const size_t smem = 2 * (size_t)stage_bytes + 512 + 64;
    SF_CHECK_ARG(smem <= WG_SMEM_LIMIT, "tc_wgrad: tile does not fit shared memory");
    static DeviceOnce configured;
    if (configured.need()) {
        cudaError_t e = cudaFuncSetAttribute(k_tc_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WG_SMEM_LIMIT);
        if (e != cudaSuccess) { set_error("tc_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SF_ERR_CUDA; }
        configured.done();
    }
    const long long items = blocks * p.splits;
    p.bias_col = ((uint32_t)p.nbc * 8u + 31u) & ~31u;
    p.ncols = tmem_cols_pow2(p.bias_col + (bias_grad ? 32u : 0u));
    long long per_sm = std::min<long long>((long long)(WG_SMEM_LIMIT / (smem + 1024)), (long long)(512u / p.ncols));
    per_sm = std::max(1LL, std::min(per_sm, 2LL));
    long long grid = std::min(items, (long long)sm_count() * per_sm);
    // algorithmic work: 2*M*N*K FLOP; both operands read once per (m-block, n-block) pair that needs them
    ProfScope ps(name, 2.0 * (double)M * N * K, 2.0 * (double)M * ((double)N * p.n_blocks + (double)K * p.m_blocks) + 4.0 * (double)N * K, st);
    k_tc_wgrad<<<(unsigned)grid, WG_THREADS, smem, st>>>(p);
    SF_CHECK_LAUNCH("tc_wgrad");
    return SF_OK;
}

}  // namespace sf

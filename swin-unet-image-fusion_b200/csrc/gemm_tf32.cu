// TF32 tensor-core GEMMs of the backward pass of SF_PREC_BF16 operators (mma.sync.m16n8k8.tf32, fp32 accumulation):
//   gemm_tf32_rows : C[M,N] (+)= A[M,K] op(B) (+ bias) (* ELU'(aux))     op(B) = B[K][N]  (dX = dY W, a018 adjoints)
//                                                                        or   = W[N][K]^T (y = x W^T + b: the forward
//                                                                               recompute of q/k/v, hidden, patch linear)
//   gemm_tf32_wgrad: Wg[N,K] += G[M,N]^T f(A[M,K])   (f = identity or ELU), reduction over M split across CTAs
// These are streaming kernels (M = 10^5..10^6 token rows, N and K = 24..1536): what matters is bytes in flight.
// Both use 128-bit global loads, register prefetch of the next slice while the current one is multiplied, and
// double-buffered shared memory (one barrier per slice).  Operands are rounded to tf32 (cvt.rna) on their way into
// shared memory; tile 64 x 64, 8 warps (16 rows x 32 columns each), same fragment routine for every variant.
#include "bwd_kernels.cuh"
#include "fp32_kernels.cuh"

namespace sf {
namespace {

constexpr int LD = 64 + 8;   // row pitch: 72 % 32 == 8 -> the (tq, gq) fragment reads hit 32 distinct banks

__device__ __forceinline__ float tf32r(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// acc[nt][*] += P^T Q over one 8-deep slice kk: P = Ps[.][LD] (reduction index major, this warp's 16 columns at mrow), Q likewise
__device__ __forceinline__ void slice_mma(const float (*Ps)[LD], const float (*Qs)[LD], int kk, int mrow, int ncol, int gq, int tq, float (&acc)[4][4]) {
    const uint32_t a0 = __float_as_uint(Ps[kk + tq][mrow + gq]), a1 = __float_as_uint(Ps[kk + tq][mrow + gq + 8]);
    const uint32_t a2 = __float_as_uint(Ps[kk + tq + 4][mrow + gq]), a3 = __float_as_uint(Ps[kk + tq + 4][mrow + gq + 8]);
#pragma unroll
    for (int nt = 0; nt < 4; nt++) {
        const uint32_t b0 = __float_as_uint(Qs[kk + tq][ncol + nt * 8 + gq]), b1 = __float_as_uint(Qs[kk + tq + 4][ncol + nt * 8 + gq]);
        mma_tf32(acc[nt], a0, a1, a2, a3, b0, b1);
    }
}
__device__ __forceinline__ void ldsm4(const float* p, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
// four consecutive floats of a row of `ld` elements starting at column c (zero beyond `cols` / when !row_ok)
__device__ __forceinline__ float4 ld4(const float* __restrict__ base, long long row, int ld, int c, int cols, bool row_ok, bool vec) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!row_ok || c >= cols) return v;
    const float* p = base + row * ld + c;
    if (vec && c + 3 < cols) return __ldg(reinterpret_cast<const float4*>(p));
    v.x = __ldg(p);
    if (c + 1 < cols) v.y = __ldg(p + 1);
    if (c + 2 < cols) v.z = __ldg(p + 2);
    if (c + 3 < cols) v.w = __ldg(p + 3);
    return v;
}

struct RowsBatch { const float* A[3]; const float* B[3]; const float* bias[3]; float* C[3]; };

// grid (ceil(M/64), ceil(N/64), nbatch); reduction over K in slices of 32
template <bool ACCUM, bool ELUAUX, bool BT>
__global__ void __launch_bounds__(256) k_gemm_tf32_rows(const RowsBatch rb, const float* __restrict__ aux, long long M, int N, int K, int vecA, int vecB) {
    constexpr int KS = 32;      // reduction slice: 64 x 32 of A (8 KB) + the B slice per CTA in flight while the previous one is multiplied
    constexpr int LDR = KS + 4; // row pitch 36 words: the eight 16-byte rows of an ldmatrix tile fall in distinct bank groups
    // both operands are kept [row][k] (A: [m][k], B: [n][k]), so that whole mma fragments come from ldmatrix:
    // one x4 for the 16 x 8 A fragment, one x4 for the B fragments of two 8-wide column tiles
    __shared__ __align__(16) float As[2][64][LDR];
    __shared__ __align__(16) float Bs[2][64][LDR];
    const int z = blockIdx.z;   // selects instead of a dynamic index: the parameter struct stays in the constant bank
    const float* __restrict__ A = z == 0 ? rb.A[0] : (z == 1 ? rb.A[1] : rb.A[2]);
    const float* __restrict__ B = z == 0 ? rb.B[0] : (z == 1 ? rb.B[1] : rb.B[2]);
    const float* __restrict__ bias = z == 0 ? rb.bias[0] : (z == 1 ? rb.bias[1] : rb.bias[2]);
    float* __restrict__ C = z == 0 ? rb.C[0] : (z == 1 ? rb.C[1] : rb.C[2]);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
    const int n0 = blockIdx.y * 64;
    const int mrow = (warp & 3) * 16, ncol = (warp >> 2) * 32;
    // A slice 64 rows x 32 k: thread -> (row tid/4, k (tid%4)*4 + 16 j), stored as it is read (128-bit).
    // B as W[N][K]: same mapping and the same [n][k] tile.  B as [K][N] (weights of the dX products) stays k-major in
    // shared memory ([k][72], the tile exactly as it is read, 128-bit both ways) and feeds scalar fragment loads.
    const int arow = tid >> 2, ak = (tid & 3) * 4;
    const int bk = tid >> 4, bn = (tid & 15) * 4;
    const int nsteps = (K + KS - 1) / KS;
    const long long ntiles = (M + 63) / 64;
    float4 ra[2], rbv[2];
    auto fetch = [&](long long tile, int s) {
        const int k0 = s * KS;
        const long long m0 = tile * 64;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            ra[j] = ld4(A, m0 + arow, K, k0 + ak + 16 * j, K, m0 + arow < M, vecA);
            if (BT) rbv[j] = ld4(B, n0 + arow, K, k0 + ak + 16 * j, K, n0 + arow < N, vecB);
            else rbv[j] = ld4(B, k0 + bk + 16 * j, N, n0 + bn, N, k0 + bk + 16 * j < K, vecB);
        }
    };
    auto r4 = [](float4 v) { return make_float4(tf32r(v.x), tf32r(v.y), tf32r(v.z), tf32r(v.w)); };
    auto stash = [&](int buf) {
#pragma unroll
        for (int j = 0; j < 2; j++) {
            *reinterpret_cast<float4*>(&As[buf][arow][ak + 16 * j]) = r4(ra[j]);
            if (BT) {
                *reinterpret_cast<float4*>(&Bs[buf][arow][ak + 16 * j]) = r4(rbv[j]);
            } else {
                float (*Bk)[LD] = reinterpret_cast<float (*)[LD]>(&Bs[buf][0][0]);     // [32][72] == [64][36] floats
                *reinterpret_cast<float4*>(&Bk[bk + 16 * j][bn]) = r4(rbv[j]);
            }
        }
    };
    // ldmatrix lane addresses (b16 8x8 tile = 8 rows x 4 tf32): A x4 = {rows 0-7 | 8-15} x {k 0-3 | 4-7};
    // B x4 = {n 0-7: k 0-3, k 4-7 ; n 8-15: k 0-3, k 4-7}
    const int a_r = mrow + (lane & 7) + ((lane >> 3) & 1) * 8, a_k = (lane >> 4) * 4;
    const int b_r = ncol + (lane & 7) + (lane >> 4) * 8, b_k = ((lane >> 3) & 1) * 4;
    auto mma_slice = [&](int buf, int kk, float (&acc)[4][4]) {
        uint32_t a0, a1, a2, a3;
        ldsm4(&As[buf][a_r][kk + a_k], a0, a1, a2, a3);
        if constexpr (!BT) {
            const float (*Bk)[LD] = reinterpret_cast<const float (*)[LD]>(&Bs[buf][0][0]);
#pragma unroll
            for (int nt = 0; nt < 4; nt++) {
                const uint32_t b0 = __float_as_uint(Bk[kk + tq][ncol + nt * 8 + gq]), b1 = __float_as_uint(Bk[kk + tq + 4][ncol + nt * 8 + gq]);
                mma_tf32(acc[nt], a0, a1, a2, a3, b0, b1);
            }
        } else {
#pragma unroll
            for (int np = 0; np < 2; np++) {
                uint32_t b0, b1, b2, b3;
                ldsm4(&Bs[buf][b_r + 16 * np][kk + b_k], b0, b1, b2, b3);
                mma_tf32(acc[2 * np], a0, a1, a2, a3, b0, b1);
                mma_tf32(acc[2 * np + 1], a0, a1, a2, a3, b2, b3);
            }
        }
    };
    // Persistent over row tiles: (tile, slice) pairs form one stream, the loads of the next pair are in flight while
    // the current slice is multiplied -- also across tile boundaries (most of these GEMMs have a single slice per tile).
    float acc[4][4] = {};
    long long tile = blockIdx.x;
    int s = 0, buf = 0;
    if (tile < ntiles) {
        fetch(tile, 0);
        stash(0);
    }
    __syncthreads();
    while (tile < ntiles) {
        int ns = s + 1;
        long long ntile = tile;
        if (ns == nsteps) { ns = 0; ntile += gridDim.x; }
        const bool more = ntile < ntiles;
        if (more) fetch(ntile, ns);
        const int kleft = K - s * KS;      // skip the all-zero tail of the last slice
#pragma unroll
        for (int kk = 0; kk < KS; kk += 8)
            if (kk < kleft) mma_slice(buf, kk, acc);
        if (ns == 0) {      // tile finished: epilogue from registers
            const long long m0 = tile * 64;
#pragma unroll
            for (int nt = 0; nt < 4; nt++) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const long long m = m0 + mrow + gq + h * 8;
                    const int n = n0 + ncol + nt * 8 + 2 * tq;
                    float v0 = acc[nt][2 * h], v1 = acc[nt][2 * h + 1];
                    acc[nt][2 * h] = 0.f; acc[nt][2 * h + 1] = 0.f;
                    if (m >= M || n >= N) continue;
                    const bool two = n + 1 < N;
                    if (bias) { v0 += __ldg(bias + n); if (two) v1 += __ldg(bias + n + 1); }
                    if (ELUAUX) {
                        const float x0 = aux[m * N + n];
                        v0 *= x0 > 0.f ? 1.f : __expf(x0);
                        if (two) { const float x1 = aux[m * N + n + 1]; v1 *= x1 > 0.f ? 1.f : __expf(x1); }
                    }
                    float* dst = C + m * N + n;
                    if (ACCUM) { v0 += dst[0]; if (two) v1 += dst[1]; }
                    if (two && (N & 1) == 0) *reinterpret_cast<float2*>(dst) = make_float2(v0, v1);
                    else { dst[0] = v0; if (two) dst[1] = v1; }
                }
            }
        }
        if (more) stash(buf ^ 1);
        __syncthreads();
        buf ^= 1; tile = ntile; s = ns;
    }
}

// grid (ceil(N/64), ceil(K/64), splits); reduction over rows in slices of 32
template <bool ELU_A>
__global__ void __launch_bounds__(256) k_gemm_tf32_wgrad(const float* __restrict__ G, const float* __restrict__ A, float* __restrict__ Wg,
                                                         float* __restrict__ bg, long long M, int N, int K, long long rows_per_split,
                                                         int vecG, int vecA) {
    __shared__ __align__(16) float Gs[2][32][LD];
    __shared__ __align__(16) float As[2][32][LD];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
    const int n0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
    const long long r0 = (long long)blockIdx.z * rows_per_split;
    const long long r1 = min(M, r0 + rows_per_split);
    const int mrow = (warp & 3) * 16, ncol = (warp >> 2) * 32;
    const int lr = tid >> 4, lc = (tid & 15) * 4;
    float4 g0, g1, a0, a1;
    auto fetch = [&](long long r) {
        g0 = ld4(G, r + lr, N, n0 + lc, N, r + lr < r1, vecG);
        g1 = ld4(G, r + lr + 16, N, n0 + lc, N, r + lr + 16 < r1, vecG);
        a0 = ld4(A, r + lr, K, k0 + lc, K, r + lr < r1, vecA);
        a1 = ld4(A, r + lr + 16, K, k0 + lc, K, r + lr + 16 < r1, vecA);
    };
    // ELU through the fast exponential: its absolute error (1e-7) is far below the tf32 rounding applied next
    auto f = [](float v) { return tf32r(ELU_A ? (v > 0.f ? v : __expf(v) - 1.f) : v); };
    // bias gradient bg[n] += sum_rows G[row][n]: a by-product of the G slices of the CTAs of the first k-tile column
    // (exact fp32 values, before the tf32 rounding); every slice is stashed exactly once
    const bool do_bg = bg != nullptr && blockIdx.y == 0;
    float cs0 = 0.f, cs1 = 0.f, cs2 = 0.f, cs3 = 0.f;
    auto stash = [&](int buf) {
        if (do_bg) { cs0 += g0.x + g1.x; cs1 += g0.y + g1.y; cs2 += g0.z + g1.z; cs3 += g0.w + g1.w; }
        *reinterpret_cast<float4*>(&Gs[buf][lr][lc]) = make_float4(tf32r(g0.x), tf32r(g0.y), tf32r(g0.z), tf32r(g0.w));
        *reinterpret_cast<float4*>(&Gs[buf][lr + 16][lc]) = make_float4(tf32r(g1.x), tf32r(g1.y), tf32r(g1.z), tf32r(g1.w));
        *reinterpret_cast<float4*>(&As[buf][lr][lc]) = make_float4(f(a0.x), f(a0.y), f(a0.z), f(a0.w));
        *reinterpret_cast<float4*>(&As[buf][lr + 16][lc]) = make_float4(f(a1.x), f(a1.y), f(a1.z), f(a1.w));
    };
    float acc[4][4] = {};
    fetch(r0);
    stash(0);
    __syncthreads();
    int buf = 0;
    for (long long r = r0; r < r1; r += 32, buf ^= 1) {
        const bool more = r + 32 < r1;
        if (more) fetch(r + 32);
#pragma unroll
        for (int kk = 0; kk < 32; kk += 8) slice_mma(Gs[buf], As[buf], kk, mrow, ncol, gq, tq, acc);
        if (more) stash(buf ^ 1);
        __syncthreads();
    }
#pragma unroll
    for (int nt = 0; nt < 4; nt++) {
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const int n = n0 + mrow + gq + (e >> 1) * 8;
            const int k = k0 + ncol + nt * 8 + 2 * tq + (e & 1);
            if (n < N && k < K) atomicAdd(&Wg[(long long)n * K + k], acc[nt][e]);
        }
    }
    if (do_bg) {     // block-uniform; the main loop ended with a barrier, so slice buffer 0 is free: [16 row lanes][64 columns]
        float* red = &Gs[0][0][0];
        *reinterpret_cast<float4*>(&red[lr * 64 + lc]) = make_float4(cs0, cs1, cs2, cs3);
        __syncthreads();
        if (tid < 64 && n0 + tid < N) {
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < 16; i++) sum += red[i * 64 + tid];
            atomicAdd(&bg[n0 + tid], sum);
        }
    }
}

// row-tile CTAs (the kernel strides over tiles if the grid is smaller than their number)
inline unsigned rows_grid_x(long long M, int others) {
    (void)others;
    const long long tiles = (M + 63) / 64;       // one row tile per CTA measured 10 % faster than 4 resident CTAs per SM looping
    long long gx = tiles < 2147483647LL ? tiles : 2147483647LL;
    if (gx < 1) gx = 1;
    return (unsigned)gx;
}
inline int vec_ok(const void* p, int ld) { return ((reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld & 3) == 0) ? 1 : 0; }

}  // namespace

// C (+)= A B (* ELU'(aux)),  B row-major [K][N]
int gemm_tf32_nn(const float* A, const float* B, const float* aux, float* C, long long M, int N, int K, bool accum, cudaStream_t st) {
    dim3 grid(rows_grid_x(M, (N + 63) / 64), (unsigned)((N + 63) / 64), 1);
    ProfScope ps("bwd_gemm_nn_tf32", 2.0 * (double)M * N * K, 4.0 * ((double)M * K + (double)M * N * (accum ? 2 : 1) + (double)K * N), st);
    RowsBatch rb{};
    rb.A[0] = A; rb.B[0] = B; rb.bias[0] = nullptr; rb.C[0] = C;
    const int va = vec_ok(A, K), vb = vec_ok(B, N);
    if (aux) {
        if (accum) k_gemm_tf32_rows<true, true, false><<<grid, 256, 0, st>>>(rb, aux, M, N, K, va, vb);
        else k_gemm_tf32_rows<false, true, false><<<grid, 256, 0, st>>>(rb, aux, M, N, K, va, vb);
    } else {
        if (accum) k_gemm_tf32_rows<true, false, false><<<grid, 256, 0, st>>>(rb, aux, M, N, K, va, vb);
        else k_gemm_tf32_rows<false, false, false><<<grid, 256, 0, st>>>(rb, aux, M, N, K, va, vb);
    }
    SF_CHECK_LAUNCH("bwd_gemm_nn_tf32");
    return SF_OK;
}

// up to three problems C_z = A_z W_z^T + bias_z in one launch (W as nn.Linear stores it: [N][K])
int gemm_tf32_nt(const GemmBatch& batch, int nbatch, long long M, int N, int K, cudaStream_t st) {
    SF_CHECK_ARG(nbatch >= 1 && nbatch <= 3 && M > 0 && N > 0 && K > 0, "gemm_tf32_nt: bad shape M=%lld N=%d K=%d", M, N, K);
    dim3 grid(rows_grid_x(M, ((N + 63) / 64) * nbatch), (unsigned)((N + 63) / 64), (unsigned)nbatch);
    ProfScope ps("bwd_recompute_gemm_tf32", 2.0 * (double)M * N * K * nbatch, 4.0 * nbatch * ((double)M * K + (double)N * K + (double)M * N), st);
    RowsBatch rb{};
    int va = 1, vb = 1;
    for (int z = 0; z < nbatch; z++) {
        SF_CHECK_ARG(batch.p[z].residual == nullptr, "gemm_tf32_nt: residual is not supported");
        rb.A[z] = batch.p[z].A; rb.B[z] = batch.p[z].W; rb.bias[z] = batch.p[z].bias; rb.C[z] = batch.p[z].C;
        va &= vec_ok(batch.p[z].A, K);
        vb &= vec_ok(batch.p[z].W, K);
    }
    k_gemm_tf32_rows<false, false, true><<<grid, 256, 0, st>>>(rb, nullptr, M, N, K, va, vb);
    SF_CHECK_LAUNCH("bwd_recompute_gemm_tf32");
    return SF_OK;
}

// Wg[N,K] += G^T f(A) ; bias_grad[N] += column sums of G (optional)
int gemm_tf32_wgrad(const float* G, const float* A, float* Wg, float* bias_grad, long long M, int N, int K, bool elu_a, cudaStream_t st) {
    SF_CHECK_ARG(Wg, "gemm_tf32_wgrad: null weight gradient");
    const int tiles = ((N + 63) / 64) * ((K + 63) / 64);
    long long splits = ((long long)sm_count() * 4 + tiles - 1) / tiles;
    long long max_splits = (M + 255) / 256;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    long long rps = ((M + splits - 1) / splits + 31) / 32 * 32;
    splits = (M + rps - 1) / rps;
    dim3 grid((unsigned)((N + 63) / 64), (unsigned)((K + 63) / 64), (unsigned)splits);
    ProfScope ps("bwd_gemm_wgrad_tf32", 2.0 * (double)M * N * K, 4.0 * ((double)M * K + (double)M * N), st);
    const int vg = vec_ok(G, N), va = vec_ok(A, K);
    if (elu_a) k_gemm_tf32_wgrad<true><<<grid, 256, 0, st>>>(G, A, Wg, bias_grad, M, N, K, rps, vg, va);
    else k_gemm_tf32_wgrad<false><<<grid, 256, 0, st>>>(G, A, Wg, bias_grad, M, N, K, rps, vg, va);
    SF_CHECK_LAUNCH("bwd_gemm_wgrad_tf32");
    return SF_OK;
}

}  // namespace sf

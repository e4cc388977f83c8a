// Backward (gradient) kernels of the fused operators -- row a18 of SURVEY.md section 8.  The
// reference defines no custom backward (pure autograd, a016:164); these kernels restate the
// adjoints of a001/a003/a004/a011/a013 and are checked against autograd over the oracle.
//
// Every backward recomputes the forward intermediates from the operator inputs (nothing but the
// inputs is saved by the forward pass): exact fp32 FFMA kernels for SF_PREC_FP32 operators, TF32 /
// fp16 tensor-core kernels with fp32 accumulation for SF_PREC_BF16 operators (gemm_tf32.cu,
// attn_bwd_mma.cu).  Weight / bias / table gradients are ACCUMULATED into caller-zeroed buffers
// with fp32 atomics; activation gradients are written.
#include <cstdlib>
#include "bf16_kernels.cuh"
#include "bwd_kernels.cuh"
#include "fp32_kernels.cuh"

namespace sf {

__device__ __forceinline__ float elu_grad(float pre) { return pre > 0.f ? 1.f : expf(pre); }

// (the TF32 tensor-core variants used by SF_PREC_BF16 operators live in gemm_tf32.cu)

// =============================================================================================
// C[M,N] (+)= (A[M,K] * B[K,N]) (* ELU'(aux[M,N]))          B row-major [K][N]
// =============================================================================================
template <bool ACCUM, bool ELUAUX>
__global__ void __launch_bounds__(256) k_gemm_nn(const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ aux,
                                                 float* __restrict__ C, long long M, int N, int K) {
    __shared__ __align__(16) float As[16][64 + 4];
    __shared__ __align__(16) float Bs[16][64 + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long m0 = (long long)blockIdx.x * 64;
    const int n0 = blockIdx.y * 64;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        {   // A tile: 64 rows x 16 k, thread -> (row = tid/4, 4 consecutive k)
            const int lrow = tid >> 2, lk = (tid & 3) * 4;
            const long long m = m0 + lrow;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                int k = k0 + lk + e;
                As[lk + e][lrow] = (m < M && k < K) ? A[m * K + k] : 0.f;
            }
        }
        {   // B tile: 16 k x 64 n, thread -> (k = tid/16, 4 consecutive n)
            const int lk = tid >> 4, ln = (tid & 15) * 4;
            const int k = k0 + lk;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                int n = n0 + ln + e;
                Bs[lk][ln + e] = (k < K && n < N) ? B[(long long)k * N + n] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; k++) {
            float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const long long m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (ELUAUX) v *= elu_grad(aux[m * N + n]);
            if (ACCUM) v += C[m * N + n];
            C[m * N + n] = v;
        }
    }
}

static int launch_gemm_nn(const float* A, const float* B, const float* aux, float* C, long long M, int N, int K, bool accum,
                          cudaStream_t st, bool tf32 = false) {
    if (tf32) return gemm_tf32_nn(A, B, aux, C, M, N, K, accum, st);
    dim3 grid((unsigned)((M + 63) / 64), (unsigned)((N + 63) / 64));
    ProfScope ps("bwd_gemm_nn_f32", 2.0 * (double)M * N * K,
                 4.0 * ((double)M * K + (double)M * N * (accum ? 2 : 1) + (double)K * N), st);
    if (aux) {
        if (accum) k_gemm_nn<true, true><<<grid, 256, 0, st>>>(A, B, aux, C, M, N, K);
        else k_gemm_nn<false, true><<<grid, 256, 0, st>>>(A, B, aux, C, M, N, K);
    } else {
        if (accum) k_gemm_nn<true, false><<<grid, 256, 0, st>>>(A, B, aux, C, M, N, K);
        else k_gemm_nn<false, false><<<grid, 256, 0, st>>>(A, B, aux, C, M, N, K);
    }
    SF_CHECK_LAUNCH("bwd_gemm_nn");
    return SF_OK;
}

// =============================================================================================
// Wg[N,K] += G[M,N]^T * f(A[M,K])      (f = identity or ELU); reduction over M split across CTAs
// =============================================================================================
template <bool ELU_A>
__global__ void __launch_bounds__(256) k_gemm_tn_reduce(const float* __restrict__ G, const float* __restrict__ A, float* __restrict__ Wg,
                                                        long long M, int N, int K, long long rows_per_split) {
    __shared__ __align__(16) float Gs[16][64 + 4];
    __shared__ __align__(16) float As[16][64 + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int n0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
    const long long r0 = (long long)blockIdx.z * rows_per_split;
    const long long r1 = min(M, r0 + rows_per_split);
    float acc[4][4] = {};
    const int lr = tid >> 4, lc = (tid & 15) * 4;
    for (long long r = r0; r < r1; r += 16) {
        const long long m = r + lr;
#pragma unroll
        for (int e = 0; e < 4; e++) {
            int n = n0 + lc + e, k = k0 + lc + e;
            Gs[lr][lc + e] = (m < r1 && n < N) ? G[m * N + n] : 0.f;
            float a = (m < r1 && k < K) ? A[m * K + k] : 0.f;
            As[lr][lc + e] = ELU_A ? elu1(a) : a;
        }
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < 16; rr++) {
            float4 g = *reinterpret_cast<const float4*>(&Gs[rr][ty * 4]);
            float4 a = *reinterpret_cast<const float4*>(&As[rr][tx * 4]);
            float gv[4] = {g.x, g.y, g.z, g.w}, av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(gv[i], av[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int n = n0 + ty * 4 + i;
        if (n >= N) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int k = k0 + tx * 4 + j;
            if (k < K) atomicAdd(&Wg[(long long)n * K + k], acc[i][j]);
        }
    }
}

static int launch_colsum(const float* G, float* out, long long M, int N, cudaStream_t st);
// bias_grad (optional): += column sums of G -- fused into the tensor-core kernel, a separate pass for the exact fp32 one
static int launch_gemm_tn_reduce(const float* G, const float* A, float* Wg, long long M, int N, int K, bool elu_a, cudaStream_t st,
                                 bool tf32 = false, float* bias_grad = nullptr) {
    if (tf32 && Wg) return gemm_tf32_wgrad(G, A, Wg, bias_grad, M, N, K, elu_a, st);
    if (bias_grad) SF_TRY(launch_colsum(G, bias_grad, M, N, st));
    if (!Wg) return SF_OK;
    const int tiles = ((N + 63) / 64) * ((K + 63) / 64);
    long long splits = ((long long)sm_count() * 4 + tiles - 1) / tiles;
    long long max_splits = (M + 255) / 256;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    long long rps = ((M + splits - 1) / splits + 15) / 16 * 16;
    splits = (M + rps - 1) / rps;
    dim3 grid((unsigned)((N + 63) / 64), (unsigned)((K + 63) / 64), (unsigned)splits);
    ProfScope ps("bwd_gemm_wgrad_f32", 2.0 * (double)M * N * K, 4.0 * ((double)M * K + (double)M * N), st);
    if (elu_a) k_gemm_tn_reduce<true><<<grid, 256, 0, st>>>(G, A, Wg, M, N, K, rps);
    else k_gemm_tn_reduce<false><<<grid, 256, 0, st>>>(G, A, Wg, M, N, K, rps);
    SF_CHECK_LAUNCH("bwd_gemm_wgrad");
    return SF_OK;
}

// out[n] += sum_m G[m][n]
// Narrow rows (N <= 256): the block is a whole number of rows wide (threads = (256 / N) * N), so a thread keeps one
// column while the block walks the flat array with fully coalesced loads; row groups are then summed through shared memory.
__global__ void __launch_bounds__(256) k_colsum_narrow(const float* __restrict__ G, float* __restrict__ out, long long total, int N,
                                                       long long elems_per_block) {
    __shared__ float sacc[256];
    const long long e0 = (long long)blockIdx.x * elems_per_block, e1 = min(total, e0 + elems_per_block);
    const int nt = blockDim.x;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    long long i = e0 + threadIdx.x;
    for (; i + 3LL * nt < e1; i += 4LL * nt) {
        s0 += __ldg(G + i); s1 += __ldg(G + i + nt); s2 += __ldg(G + i + 2LL * nt); s3 += __ldg(G + i + 3LL * nt);
    }
    for (; i < e1; i += nt) s0 += __ldg(G + i);
    sacc[threadIdx.x] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (threadIdx.x < N) {
        float s = 0.f;
        for (int t = threadIdx.x; t < nt; t += N) s += sacc[t];
        atomicAdd(&out[threadIdx.x], s);
    }
}
__global__ void k_colsum(const float* __restrict__ G, float* __restrict__ out, long long M, int N, long long rows_per_block) {
    const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(M, r0 + rows_per_block);
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        float s = 0.f;
        for (long long m = r0; m < r1; m++) s += G[m * N + n];
        atomicAdd(&out[n], s);
    }
}

static int launch_colsum(const float* G, float* out, long long M, int N, cudaStream_t st) {
    if (!out) return SF_OK;
    ProfScope ps("bwd_colsum", (double)M * N, 4.0 * (double)M * N, st);
    if (N <= 256) {
        const int threads = (256 / N) * N;              // a multiple of the row length: e0 and the stride keep columns fixed
        const long long total = M * N;
        long long blocks = (long long)sm_count() * 8;
        long long epb = (total + blocks - 1) / blocks;
        epb = (epb + threads - 1) / threads * threads;  // whole block-strides, hence whole rows
        if (epb < 4LL * threads) epb = 4LL * threads;
        blocks = (total + epb - 1) / epb;
        k_colsum_narrow<<<(unsigned)blocks, threads, 0, st>>>(G, out, total, N, epb);
        SF_CHECK_LAUNCH("bwd_colsum");
        return SF_OK;
    }
    long long blocks = (long long)sm_count() * 8;
    if (blocks > (M + 63) / 64) blocks = (M + 63) / 64;
    if (blocks < 1) blocks = 1;
    long long rpb = (M + blocks - 1) / blocks;
    blocks = (M + rpb - 1) / rpb;
    k_colsum<<<(unsigned)blocks, 256, 0, st>>>(G, out, M, N, rpb);
    SF_CHECK_LAUNCH("bwd_colsum");
    return SF_OK;
}

// =============================================================================================
// LayerNorm backward (adjoint of my_layer_norm, a004:54-72), one warp per row.
//   y = (x - mean) * rstd * gamma + beta ;  gy_eff = ELUOUT ? gy * ELU'(y) : gy
//   gx (=|+=) rstd * (gyg - mean(gyg) - xhat * mean(gyg * xhat)),  gyg = gy_eff * gamma
//   ggamma += sum_rows gy_eff * xhat ; gbeta += sum_rows gy_eff
// =============================================================================================
template <bool ELUOUT, bool ACCUM>
__global__ void k_ln_bwd(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                         const float* __restrict__ gy, float* gx, const float* gadd, float* __restrict__ ggamma, float* __restrict__ gbeta,
                         long long M, int C, float eps) {
    // per-warp private [2][C] accumulators: lane l owns columns l, l+32, ... of its warp's copy, so the updates need
    // neither atomics nor synchronisation; the copies are summed once at the end
    extern __shared__ float sacc_all[];  // [warps][2][C]
    for (int c = threadIdx.x; c < (int)(blockDim.x >> 5) * 2 * C; c += blockDim.x) sacc_all[c] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    float* sacc = sacc_all + (threadIdx.x >> 5) * 2 * C;
    long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long stride = ((long long)gridDim.x * blockDim.x) >> 5;
    for (; row < M; row += stride) {
        const float* xr = x + row * C;
        const float* gr = gy + row * C;
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s += xr[c];
        s = warp_sum(s);
        const float mean = s / (float)C;
        float v = 0.f;
        for (int c = lane; c < C; c += 32) { float d = xr[c] - mean; v += d * d; }
        v = warp_sum(v);
        const float rstd = rsqrtf(v / (float)C + eps);
        float s1 = 0.f, s2 = 0.f;
        for (int c = lane; c < C; c += 32) {
            float xh = (xr[c] - mean) * rstd;
            float g = gr[c];
            if (ELUOUT) g *= elu_grad(xh * gamma[c] + beta[c]);
            float gg = g * gamma[c];
            s1 += gg;
            s2 += gg * xh;
            sacc[c] += g * xh;
            sacc[C + c] += g;
        }
        s1 = warp_sum(s1) / (float)C;
        s2 = warp_sum(s2) / (float)C;
        for (int c = lane; c < C; c += 32) {
            float xh = (xr[c] - mean) * rstd;
            float g = gr[c];
            if (ELUOUT) g *= elu_grad(xh * gamma[c] + beta[c]);
            float val = rstd * (g * gamma[c] - s1 - xh * s2);
            if (ACCUM) val += gadd[row * C + c];
            gx[row * C + c] = val;
        }
    }
    __syncthreads();
    const int nw = blockDim.x >> 5;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = 0.f, b = 0.f;
        for (int w = 0; w < nw; w++) { a += sacc_all[w * 2 * C + c]; b += sacc_all[w * 2 * C + C + c]; }
        if (ggamma) atomicAdd(&ggamma[c], a);
        if (gbeta) atomicAdd(&gbeta[c], b);
    }
}

// Short rows (C <= 128, C % 4 == 0): a group of G lanes (8, 16 or 32) owns one row, four consecutive channels per lane
// (128-bit accesses), so a warp works on 32 / G rows at once, the row reductions are log2(G) shuffle steps, and the
// lane's channels never change: ggamma / gbeta partial sums stay in eight registers until the end of the kernel.
template <int G, bool ELUOUT, bool ACCUM>
__global__ void __launch_bounds__(256) k_ln_bwd_grp(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                    const float* __restrict__ gy, float* gx, const float* gadd, float* __restrict__ ggamma,
                                                    float* __restrict__ gbeta, long long M, int C, float eps) {
    __shared__ float sacc[8][2][128];
    constexpr int RPW = 32 / G;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = (lane % G) * 4, sub = lane / G;
    const bool act = c0 < C;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 gm = act ? __ldg(reinterpret_cast<const float4*>(gamma + c0)) : z4;
    const float4 bt = act ? __ldg(reinterpret_cast<const float4*>(beta + c0)) : z4;
    const float invC = 1.f / (float)C;
    float ag[4] = {0.f, 0.f, 0.f, 0.f}, ab[4] = {0.f, 0.f, 0.f, 0.f};
    const long long wstride = (long long)gridDim.x * 8 * RPW;
    for (long long base = ((long long)blockIdx.x * 8 + warp) * RPW; base < M; base += wstride) {
        const long long row = base + sub;
        const bool ok = act && row < M;
        const float4 xv = ok ? __ldg(reinterpret_cast<const float4*>(x + row * C + c0)) : z4;
        float4 gv = ok ? __ldg(reinterpret_cast<const float4*>(gy + row * C + c0)) : z4;
        float sx = (xv.x + xv.y) + (xv.z + xv.w);
#pragma unroll
        for (int o = G / 2; o; o >>= 1) sx += __shfl_xor_sync(0xffffffffu, sx, o);
        const float mean = sx * invC;
        float d0 = xv.x - mean, d1 = xv.y - mean, d2 = xv.z - mean, d3 = xv.w - mean;
        if (!act) { d0 = 0.f; d1 = 0.f; d2 = 0.f; d3 = 0.f; }
        float var = (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
#pragma unroll
        for (int o = G / 2; o; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
        const float rstd = rsqrtf(var * invC + eps);
        const float h0 = d0 * rstd, h1 = d1 * rstd, h2 = d2 * rstd, h3 = d3 * rstd;
        if (ELUOUT) {
            gv.x *= elu_grad(h0 * gm.x + bt.x); gv.y *= elu_grad(h1 * gm.y + bt.y);
            gv.z *= elu_grad(h2 * gm.z + bt.z); gv.w *= elu_grad(h3 * gm.w + bt.w);
        }
        const float g0 = gv.x * gm.x, g1 = gv.y * gm.y, g2 = gv.z * gm.z, g3 = gv.w * gm.w;
        float s1 = (g0 + g1) + (g2 + g3), s2 = (g0 * h0 + g1 * h1) + (g2 * h2 + g3 * h3);
#pragma unroll
        for (int o = G / 2; o; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        s1 *= invC; s2 *= invC;
        if (ok) {
            float4 o4 = make_float4(rstd * (g0 - s1 - h0 * s2), rstd * (g1 - s1 - h1 * s2), rstd * (g2 - s1 - h2 * s2), rstd * (g3 - s1 - h3 * s2));
            float4* dst = reinterpret_cast<float4*>(gx + row * C + c0);
            if (ACCUM) { const float4 old = *reinterpret_cast<const float4*>(gadd + row * C + c0); o4.x += old.x; o4.y += old.y; o4.z += old.z; o4.w += old.w; }
            *dst = o4;
            ag[0] = fmaf(gv.x, h0, ag[0]); ag[1] = fmaf(gv.y, h1, ag[1]); ag[2] = fmaf(gv.z, h2, ag[2]); ag[3] = fmaf(gv.w, h3, ag[3]);
            ab[0] += gv.x; ab[1] += gv.y; ab[2] += gv.z; ab[3] += gv.w;
        }
    }
    // lanes of the other row groups of the warp hold the same channels
#pragma unroll
    for (int o = G; o < 32; o <<= 1)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            ag[j] += __shfl_xor_sync(0xffffffffu, ag[j], o);
            ab[j] += __shfl_xor_sync(0xffffffffu, ab[j], o);
        }
    if (sub == 0 && act) {
#pragma unroll
        for (int j = 0; j < 4; j++) { sacc[warp][0][c0 + j] = ag[j]; sacc[warp][1][c0 + j] = ab[j]; }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) { a += sacc[w][0][c]; b += sacc[w][1][c]; }
        if (ggamma) atomicAdd(&ggamma[c], a);
        if (gbeta) atomicAdd(&gbeta[c], b);
    }
}

template <int G>
static void launch_ln_bwd_grp(const float* x, const float* gamma, const float* beta, const float* gy, float* gx, const float* gadd, float* ggamma,
                              float* gbeta, long long M, int C, float eps, bool eluout, cudaStream_t st) {
    const bool accum = gadd != nullptr;
    const long long rows_per_block = 8 * (32 / G);
    long long blocks = (M + rows_per_block - 1) / rows_per_block;
    if (blocks > (long long)sm_count() * 8) blocks = (long long)sm_count() * 8;
    if (eluout) {
        if (accum) k_ln_bwd_grp<G, true, true><<<(unsigned)blocks, 256, 0, st>>>(x, gamma, beta, gy, gx, gadd, ggamma, gbeta, M, C, eps);
        else k_ln_bwd_grp<G, true, false><<<(unsigned)blocks, 256, 0, st>>>(x, gamma, beta, gy, gx, gadd, ggamma, gbeta, M, C, eps);
    } else {
        if (accum) k_ln_bwd_grp<G, false, true><<<(unsigned)blocks, 256, 0, st>>>(x, gamma, beta, gy, gx, gadd, ggamma, gbeta, M, C, eps);
        else k_ln_bwd_grp<G, false, false><<<(unsigned)blocks, 256, 0, st>>>(x, gamma, beta, gy, gx, gadd, ggamma, gbeta, M, C, eps);
    }
}

// gadd (optional): a (M,C) tensor added to the result -- gx itself to accumulate in place, or another gradient branch
static int launch_ln_bwd(const float* x, const float* gamma, const float* beta, const float* gy, float* gx, float* ggamma,
                         float* gbeta, long long M, int C, float eps, bool eluout, const float* gadd, cudaStream_t st) {
    const bool accum = gadd != nullptr;
    const bool al16 = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gy) | reinterpret_cast<uintptr_t>(gx) |
                        reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) | reinterpret_cast<uintptr_t>(gadd)) & 15) == 0;
    if (C <= 128 && (C & 3) == 0 && al16) {
        ProfScope ps("bwd_layernorm", 20.0 * (double)M * C, 12.0 * (double)M * C, st);
        if (C <= 32) launch_ln_bwd_grp<8>(x, gamma, beta, gy, gx, gadd, ggamma, gbeta, M, C, eps, eluout, st);
        else if (C <= 64) launch_ln_bwd_grp<16>(x, gamma, beta, gy, gx, gadd, ggamma, gbeta, M, C, eps, eluout, st);
        else launch_ln_bwd_grp<32>(x, gamma, beta, gy, gx, gadd, ggamma, gbeta, M, C, eps, eluout, st);
        SF_CHECK_LAUNCH("bwd_layernorm");
        return SF_OK;
    }
    const int threads = 256;
    long long blocks = (M * 32 + threads - 1) / threads;
    if (blocks > (long long)sm_count() * 4) blocks = (long long)sm_count() * 4;
    if (blocks < 1) blocks = 1;
    size_t smem = (size_t)(threads / 32) * 2 * (size_t)C * sizeof(float);
    SF_CHECK_ARG(smem <= 200 * 1024, "LayerNorm backward: row length %d needs %zu B of shared memory", C, smem);
    if (smem > 48 * 1024) {
        static DeviceOnce configured;
        if (configured.need()) {
            cudaFuncSetAttribute(k_ln_bwd<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            cudaFuncSetAttribute(k_ln_bwd<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            cudaFuncSetAttribute(k_ln_bwd<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            cudaFuncSetAttribute(k_ln_bwd<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            configured.done();
        }
    }
    ProfScope ps("bwd_layernorm", 20.0 * (double)M * C, 12.0 * (double)M * C, st);
    if (eluout) {
        if (accum) k_ln_bwd<true, true><<<(unsigned)blocks, threads, smem, st>>>(x, gamma, beta, gy, gx, gadd, ggamma, gbeta, M, C, eps);
        else k_ln_bwd<true, false><<<(unsigned)blocks, threads, smem, st>>>(x, gamma, beta, gy, gx, gadd, ggamma, gbeta, M, C, eps);
    } else {
        if (accum) k_ln_bwd<false, true><<<(unsigned)blocks, threads, smem, st>>>(x, gamma, beta, gy, gx, gadd, ggamma, gbeta, M, C, eps);
        else k_ln_bwd<false, false><<<(unsigned)blocks, threads, smem, st>>>(x, gamma, beta, gy, gx, gadd, ggamma, gbeta, M, C, eps);
    }
    SF_CHECK_LAUNCH("bwd_layernorm");
    return SF_OK;
}

// =============================================================================================
// attention core backward (adjoint of a001:317-354), one (window, head) at a time per CTA
// =============================================================================================
// dV = P^T dO ; dP = dO V^T ; dS = P o (dP - rowsum(dP o P)) ; dtable[idx(i,j)] += dS ;
// dQ = scale * dS K ; dK = scale * dS^T Q.   Masked entries have P = 0 hence dS = 0.
__global__ void k_attn_core_bwd(const float* __restrict__ Q, const float* __restrict__ Kt, const float* __restrict__ V,
                                const float* __restrict__ gO, float* __restrict__ dQ, float* __restrict__ dK, float* __restrict__ dV,
                                const float* __restrict__ table, float* __restrict__ gtable, WinGeom g, int inner, int nh, int d,
                                float scale, long long nitems) {
    extern __shared__ float smb[];
    const int T = g.T, TS = T + 1;
    const int tw = 2 * g.wsw - 1, tabn = (2 * g.wsh - 1) * tw;
    float* Qs = smb;                    // [T][d]
    float* Ks = Qs + T * d;
    float* Vs = Ks + T * d;
    float* Gs = Vs + T * d;             // dO
    float* Ps = Gs + T * d;             // [T][TS]
    float* Ss = Ps + T * TS;            // dS [T][TS]
    float* tab = Ss + T * TS;           // [tabn]
    float* tacc = tab + tabn;           // [tabn]
    long long* rows = reinterpret_cast<long long*>(tacc + tabn + ((2 * tabn + 4 * T * d + 2 * T * TS) & 1));
    int* regs = reinterpret_cast<int*>(rows + T);
    for (int i = threadIdx.x; i < tabn; i += blockDim.x) { tab[i] = table[i]; tacc[i] = 0.f; }

    for (long long item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int win = (int)(item / nh), head = (int)(item % nh);
        __syncthreads();
        for (int t = threadIdx.x; t < T; t += blockDim.x) {
            int rg;
            rows[t] = win_token_src(g, win, t, &rg);
            regs[t] = rg;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < T * d; i += blockDim.x) {
            int t = i / d, dd = i - t * d;
            long long off = rows[t] * inner + head * d + dd;
            Qs[i] = Q[off]; Ks[i] = Kt[off]; Vs[i] = V[off]; Gs[i] = gO[off];
        }
        __syncthreads();
        for (int qi = threadIdx.x; qi < T; qi += blockDim.x) {
            const int qr = qi / g.wsw, qc = qi - qr * g.wsw, qreg = regs[qi];
            float mx = -INFINITY;
            for (int j = 0; j < T; j++) {
                float s = 0.f;
                for (int dd = 0; dd < d; dd++) s = fmaf(Qs[qi * d + dd], Ks[j * d + dd], s);
                int jr = j / g.wsw, jc = j - jr * g.wsw;
                s = s * scale + tab[(jr - qr + g.wsh - 1) * tw + (jc - qc + g.wsw - 1)];
                if (regs[j] != qreg) s = -1e10f;
                Ps[qi * TS + j] = s;
                mx = fmaxf(mx, s);
            }
            float sum = 0.f;
            for (int j = 0; j < T; j++) { float e = expf(Ps[qi * TS + j] - mx); Ps[qi * TS + j] = e; sum += e; }
            const float inv = 1.f / sum;
            float delta = 0.f;
            for (int j = 0; j < T; j++) {
                float p = Ps[qi * TS + j] * inv;
                float dp = 0.f;
                for (int dd = 0; dd < d; dd++) dp = fmaf(Gs[qi * d + dd], Vs[j * d + dd], dp);
                Ps[qi * TS + j] = p;
                Ss[qi * TS + j] = dp;
                delta = fmaf(p, dp, delta);
            }
            for (int j = 0; j < T; j++) {
                float ds = Ps[qi * TS + j] * (Ss[qi * TS + j] - delta);
                Ss[qi * TS + j] = ds;
                int jr = j / g.wsw, jc = j - jr * g.wsw;
                if (gtable && ds != 0.f) atomicAdd(&tacc[(jr - qr + g.wsh - 1) * tw + (jc - qc + g.wsw - 1)], ds);
            }
            const long long off = rows[qi] * inner + head * d;
            for (int dd = 0; dd < d; dd++) {
                float a = 0.f;
                for (int j = 0; j < T; j++) a = fmaf(Ss[qi * TS + j], Ks[j * d + dd], a);
                dQ[off + dd] = a * scale;
            }
        }
        __syncthreads();
        for (int kj = threadIdx.x; kj < T; kj += blockDim.x) {
            const long long off = rows[kj] * inner + head * d;
            for (int dd = 0; dd < d; dd++) {
                float a = 0.f, b = 0.f;
                for (int i = 0; i < T; i++) {
                    a = fmaf(Ss[i * TS + kj], Qs[i * d + dd], a);
                    b = fmaf(Ps[i * TS + kj], Gs[i * d + dd], b);
                }
                dK[off + dd] = a * scale;
                dV[off + dd] = b;
            }
        }
    }
    __syncthreads();
    if (gtable)
        for (int i = threadIdx.x; i < tabn; i += blockDim.x) atomicAdd(&gtable[i], tacc[i]);
}

// 7x7 windows (the model's only window size): the same algorithm with compile-time window constants --
// no integer divisions in the inner loops, unrolled key loops, and the gradient of the 13x13 bias table
// accumulated in registers (each query row owns 49 distinct table entries) instead of 2401 shared-memory
// atomics per (window, head); one global atomicAdd per entry and thread at the end of the kernel.
__global__ void __launch_bounds__(64) k_attn_core_bwd_w7(const float* __restrict__ Q, const float* __restrict__ Kt, const float* __restrict__ V,
                                                       const float* __restrict__ gO, float* __restrict__ dQ, float* __restrict__ dK,
                                                       float* __restrict__ dV, const float* __restrict__ table, float* __restrict__ gtable,
                                                       WinGeom g, int inner, int nh, int d, float scale, long long nitems) {
    constexpr int T = 49, TS = 51, TW = 13;   // TS odd: row-strided accesses of a warp are conflict-free
    extern __shared__ float smb[];
    float* Qs = smb;                    // [T][d]
    float* Ks = Qs + T * d;
    float* Vs = Ks + T * d;
    float* Gs = Vs + T * d;             // dO
    float* Ps = Gs + T * d;             // [T][TS]
    float* Ss = Ps + T * TS;            // dS [T][TS]
    float* tab = Ss + T * TS;           // [169]
    long long* rows = reinterpret_cast<long long*>(tab + 169 + ((169 + 4 * T * d + 2 * T * TS) & 1));
    int* regs = reinterpret_cast<int*>(rows + T);
    for (int i = threadIdx.x; i < 169; i += blockDim.x) tab[i] = table[i];
    const int qi = threadIdx.x;
    const bool active = qi < T;
    const int qr = qi / 7, qc = qi - qr * 7;
    const int qoff = (6 - qr) * TW + (6 - qc);   // table index of (query qi, key j) = qoff + (j/7)*13 + j%7
    float tl[T];                                  // this query row's 49 table-gradient entries
#pragma unroll
    for (int j = 0; j < T; j++) tl[j] = 0.f;

    for (long long item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int win = (int)(item / nh), head = (int)(item % nh);
        __syncthreads();
        if (active) {
            int rg;
            rows[qi] = win_token_src(g, win, qi, &rg);
            regs[qi] = rg;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < T * d; i += blockDim.x) {
            int t = i / d, dd = i - t * d;
            long long off = rows[t] * inner + head * d + dd;
            Qs[i] = Q[off]; Ks[i] = Kt[off]; Vs[i] = V[off]; Gs[i] = gO[off];
        }
        __syncthreads();
        if (active) {
            const int qreg = regs[qi];
            float* Pr = Ps + qi * TS;
            float* Sr = Ss + qi * TS;
            float mx = -INFINITY;
#pragma unroll 7
            for (int j = 0; j < T; j++) {
                float s = 0.f, dp = 0.f;
                for (int dd = 0; dd < d; dd++) {
                    s = fmaf(Qs[qi * d + dd], Ks[j * d + dd], s);
                    dp = fmaf(Gs[qi * d + dd], Vs[j * d + dd], dp);
                }
                s = s * scale + tab[qoff + (j / 7) * TW + (j % 7)];
                if (regs[j] != qreg) s = -1e10f;
                Pr[j] = s;
                Sr[j] = dp;
                mx = fmaxf(mx, s);
            }
            float sum = 0.f;
#pragma unroll 7
            for (int j = 0; j < T; j++) { float e = expf(Pr[j] - mx); Pr[j] = e; sum += e; }
            const float inv = 1.f / sum;
            float delta = 0.f;
#pragma unroll 7
            for (int j = 0; j < T; j++) {
                float p = Pr[j] * inv;
                Pr[j] = p;
                delta = fmaf(p, Sr[j], delta);
            }
#pragma unroll
            for (int j = 0; j < T; j++) {
                const float ds = Pr[j] * (Sr[j] - delta);
                Sr[j] = ds;
                tl[j] += ds;
            }
            const long long off = rows[qi] * inner + head * d;
            for (int dd = 0; dd < d; dd++) {
                float a = 0.f;
#pragma unroll 7
                for (int j = 0; j < T; j++) a = fmaf(Sr[j], Ks[j * d + dd], a);
                dQ[off + dd] = a * scale;
            }
        }
        __syncthreads();
        if (active) {
            const int kj = qi;
            const long long off = rows[kj] * inner + head * d;
            for (int dd = 0; dd < d; dd++) {
                float a = 0.f, b = 0.f;
#pragma unroll 7
                for (int i = 0; i < T; i++) {
                    a = fmaf(Ss[i * TS + kj], Qs[i * d + dd], a);
                    b = fmaf(Ps[i * TS + kj], Gs[i * d + dd], b);
                }
                dK[off + dd] = a * scale;
                dV[off + dd] = b;
            }
        }
    }
    if (gtable && active) {
#pragma unroll
        for (int j = 0; j < T; j++)
            if (tl[j] != 0.f) atomicAdd(&gtable[qoff + (j / 7) * TW + (j % 7)], tl[j]);
    }
}

static int launch_attn_core_bwd(const float* Q, const float* K, const float* V, const float* gO, float* dQ, float* dK, float* dV,
                                const float* table, float* gtable, const WinGeom& g, int inner, int nh, int d, cudaStream_t st,
                                bool tensor_cores, float* O_out = nullptr) {
    if (tensor_cores && attn_core_bwd_mma_supported(g, d, nh))
        return launch_attn_core_bwd_mma(Q, K, V, gO, dQ, dK, dV, O_out, table, gtable, g, inner, nh, d, st);
    const int tabn = (2 * g.wsh - 1) * (2 * g.wsw - 1);
    size_t smem = ((size_t)4 * g.T * d + 2 * (size_t)g.T * (g.T + 1) + 2 * tabn + 2) * sizeof(float) + (size_t)g.T * 12 + 16;
    SF_CHECK_ARG(smem <= 200 * 1024, "attention backward: window of %d tokens x head_dim %d needs %zu B of shared memory", g.T, d, smem);
    static DeviceOnce configured;
    if (configured.need()) {
        cudaError_t e = cudaFuncSetAttribute(k_attn_core_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) { set_error("attention backward: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SF_ERR_CUDA; }
        configured.done();
    }
    const long long nitems = (long long)g.B * g.nWh * g.nWw * nh;
    int threads = g.T <= 64 ? 64 : (g.T <= 128 ? 128 : 256);
    int per_sm = (int)((200 * 1024) / (smem + 1024));
    if (per_sm > 16) per_sm = 16;
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)sm_count() * per_sm;
    if (grid > nitems) grid = nitems;
    const double mtok = (double)g.B * g.Hp * g.Wp;
    ProfScope ps("bwd_attn_core_f32", 12.0 * g.T * mtok * inner, 28.0 * mtok * inner, st);
    if (g.wsh == 7 && g.wsw == 7) {
        const size_t smem7 = ((size_t)4 * 49 * d + 2 * 49 * 51 + 169 + 2) * sizeof(float) + 49 * 12 + 16;
        static DeviceOnce configured7;
        if (configured7.need()) {
            cudaError_t e = cudaFuncSetAttribute(k_attn_core_bwd_w7, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e != cudaSuccess) { set_error("attention backward: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SF_ERR_CUDA; }
            configured7.done();
        }
        int psm = (int)((200 * 1024) / (smem7 + 1024));
        if (psm > 16) psm = 16;
        if (psm < 1) psm = 1;
        long long grid7 = (long long)sm_count() * psm;
        if (grid7 > nitems) grid7 = nitems;
        k_attn_core_bwd_w7<<<(unsigned)grid7, 64, smem7, st>>>(Q, K, V, gO, dQ, dK, dV, table, gtable, g, inner, nh, d, 1.0f / sqrtf((float)d), nitems);
        SF_CHECK_LAUNCH("bwd_attn_core");
        return SF_OK;
    }
    k_attn_core_bwd<<<(unsigned)grid, threads, smem, st>>>(Q, K, V, gO, dQ, dK, dV, table, gtable, g, inner, nh, d, 1.0f / sqrtf((float)d), nitems);
    SF_CHECK_LAUNCH("bwd_attn_core");
    return SF_OK;
}


// =============================================================================================
// SF_PREC_BF16 operators: every backward GEMM on tcgen05
// =============================================================================================
// Activations and gradients that feed a GEMM are bf16 tensors in the UMMA-tiled layout (tail rows zero): produced by the
// LayerNorm / cast pre-pass (k_ln_to_tiled) or by a GEMM epilogue, consumed by bulk copies.  Forward recompute and the
// data gradients dX = dY W run through k_tc_gemm2 (dX with a TRANSPOSED weight image), the weight gradients
// dW = dY^T X through k_tc_wgrad (tc_wgrad.cu), accumulation in fp32 (TMEM) throughout.
static bool bwd_tc_enabled() {
    static const bool on = [] { const char* e = getenv("SWINFUSE_BWD_TC"); return !(e && e[0] == '0'); }();
    return on;
}
static inline size_t tiled_bytes(long long M, int cols) { return tiled_elems(M, cols) * sizeof(bf16); }

// image of W (rows = output features, as stored) with its bias: the forward GEMM  y = x W^T + b
static void pack_fwd(PackJobs& jobs, const float* W, const float* b, int N, int K, const PackedGemm& g, char* base) {
    jobs.job[jobs.n++] = PackJob{{W, nullptr, nullptr}, {b, nullptr, nullptr}, N, N, K, 0, (bf16*)(base + g.off_w), (float*)(base + g.off_b), g.nch, g.ks, g.nc, g.nslabs};
}
// stacked forward image [W0; W1; (W2)] (each [Neach][K]) with the stacked biases
static void pack_fwd_stack(PackJobs& jobs, int nsrc, const float* const* W, const float* const* b, int Neach, int K, const PackedGemm& g, char* base) {
    PackJob j{{W[0], nsrc > 1 ? W[1] : nullptr, nsrc > 2 ? W[2] : nullptr}, {b[0], nsrc > 1 ? b[1] : nullptr, nsrc > 2 ? b[2] : nullptr},
              Neach, nsrc * Neach, K, 0, (bf16*)(base + g.off_w), (float*)(base + g.off_b), g.nch, g.ks, g.nc, g.nslabs};
    jobs.job[jobs.n++] = j;
}
// image of [W0; W1; (W2)]^T (each W stored [R][Cc]): dX[M x Cc] = [dY0 | dY1 | dY2][M x nsrc*R] [W0; W1; W2]
static void pack_tr_stack(PackJobs& jobs, int nsrc, const float* const* W, int R, int Cc, const PackedGemm& g, char* base) {
    PackJob j{{W[0], nsrc > 1 ? W[1] : nullptr, nsrc > 2 ? W[2] : nullptr}, {nullptr, nullptr, nullptr},
              R, Cc, nsrc * R, 1, (bf16*)(base + g.off_w), (float*)(base + g.off_b), g.nch, g.ks, g.nc, g.nslabs};
    jobs.job[jobs.n++] = j;
}
// image of W^T for W stored [R][Cc]: the data-gradient GEMM  dX[M x Cc] = dY[M x R] W
static void pack_tr(PackJobs& jobs, const float* W, int R, int Cc, const PackedGemm& g, char* base) {
    jobs.job[jobs.n++] = PackJob{{W, nullptr, nullptr}, {nullptr, nullptr, nullptr}, R, Cc, R, 1, (bf16*)(base + g.off_w), (float*)(base + g.off_b), g.nch, g.ks, g.nc, g.nslabs};
}
// out = A_tiled[M x K] * image^T:  fp32 rows (optionally accumulated into `out_f32`) or bf16 tiled (optionally x ELU'(aux))
static int tc_gemm_tiled(const bf16* A, long long M, int K, int N, const PackedGemm& g, const char* pk, bool with_bias, bool elu,
                         float* out_f32, bool accumulate, bf16* out_tiled, const bf16* elu_aux, const char* name, cudaStream_t st) {
    TcGemm t{};
    t.A = A; t.M = M; t.K = K; t.a_mode = AM_TILED;
    bind_packed(t, g, pk);
    if (!with_bias) t.bias = nullptr;
    t.elu = elu ? 1 : 0;
    t.N = N;
    if (out_tiled) {
        t.out_mode = OUT_TILED; t.out = out_tiled; t.out_nkc = (int)tc::pad16((uint32_t)N) / 8; t.elu_aux = elu_aux; t.zero_tail = 1;
    } else {
        t.out_mode = OUT_F32; t.out = out_f32; t.ldo = N;
        if (accumulate) { t.residual = out_f32; t.ldr = N; }
    }
    SF_TRY(tc_gemm_plan(&t));
    return launch_tc_gemm(t, name, st);
}
// ---- MLP ------------------------------------------------------------------------------------------------------------------
struct MlpBwdPlan { PackedGemm w1, w2t, w1t; size_t off_pk, off_n, off_h, off_g, off_gh, off_gn, total; };
static MlpBwdPlan mlp_bwd_tc_plan(const sf_mlp_params* p) {
    MlpBwdPlan m{};
    Carver pc;
    m.w1 = plan_packed(pc, p->hidden, p->C);      // hpre = n W1^T + b1
    m.w2t = plan_packed(pc, p->hidden, p->C);     // g_h  = gout W2        (image of W2^T: rows = hidden units, k = output channels)
    m.w1t = plan_packed(pc, p->C, p->hidden);     // g_n  = g_h W1         (image of W1^T: rows = input channels, k = hidden units)
    Carver c;
    m.off_pk = c.take(pc.off);
    m.off_n = c.take(tiled_bytes(p->M, p->C));
    m.off_h = c.take(tiled_bytes(p->M, p->hidden));
    m.off_g = c.take(tiled_bytes(p->M, p->C));
    m.off_gh = c.take(tiled_bytes(p->M, p->hidden));
    m.off_gn = c.take((size_t)p->M * p->C * sizeof(float));
    m.total = c.off;
    return m;
}
static bool mlp_bwd_tc_ok(const sf_mlp_params* p) {
    return bwd_tc_enabled() && p->C % 4 == 0 && (int)tc::pad16((uint32_t)p->C) <= TC_MAX_KPAD && p->hidden % 4 == 0 &&
           aligned16(p->in) && p->M < 2147483647LL;
}
static int mlp_bwd_tc(const sf_mlp_bwd_params* bp, void* ws_ptr, size_t ws_bytes, cudaStream_t st) {
    const sf_mlp_params* p = &bp->fwd;
    const long long M = p->M;
    const int C = p->C, H = p->hidden;
    const MlpBwdPlan m = mlp_bwd_tc_plan(p);
    if (ws_bytes < m.total || !ws_ptr) { set_error("sf_mlp_bwd: workspace too small (%zu B given, %zu needed)", ws_bytes, m.total); return SF_ERR_WORKSPACE; }
    char* base = reinterpret_cast<char*>(ws_ptr);
    char* pk = base + m.off_pk;
    bf16* n_t = reinterpret_cast<bf16*>(base + m.off_n);
    bf16* h_t = reinterpret_cast<bf16*>(base + m.off_h);
    bf16* g_t = reinterpret_cast<bf16*>(base + m.off_g);
    bf16* gh_t = reinterpret_cast<bf16*>(base + m.off_gh);
    float* gn = reinterpret_cast<float*>(base + m.off_gn);
    PackJobs jobs{};
    pack_fwd(jobs, p->w1, p->b1, H, C, m.w1, pk);
    pack_tr(jobs, p->w2, C, H, m.w2t, pk);
    pack_tr(jobs, p->w1, H, C, m.w1t, pk);
    SF_TRY(launch_pack_jobs(jobs, st));
    // recompute: n = LN(x) (or x), a = ELU(n W1^T + b1), both kept as bf16 tiles
    SF_TRY(launch_ln_to_tiled(p->in, p->ln_gamma, p->ln_beta, n_t, M, C, p->ln_eps, st));
    SF_TRY(tc_gemm_tiled(n_t, M, C, H, m.w1, pk, true, true, nullptr, false, h_t, nullptr, "bwd_tc_recompute_mlp1", st));
    SF_TRY(launch_ln_to_tiled(bp->gout, nullptr, nullptr, g_t, M, C, 0.f, st));
    // g_h = (gout W2) o ELU'(pre), ELU' read off a
    SF_TRY(tc_gemm_tiled(g_t, M, C, H, m.w2t, pk, false, false, nullptr, false, gh_t, h_t, "bwd_tc_dx_mlp2", st));
    SF_TRY(launch_tc_wgrad(g_t, h_t, bp->g_w2, bp->g_b2, M, C, H, "bwd_tc_wgrad", st));      // gW2 = gout^T a, gb2
    SF_TRY(launch_tc_wgrad(gh_t, n_t, bp->g_w1, bp->g_b1, M, H, C, "bwd_tc_wgrad", st));    // gW1 = g_h^T n, gb1
    float* gdst = p->ln_gamma ? gn : bp->g_in;
    SF_TRY(tc_gemm_tiled(gh_t, M, H, C, m.w1t, pk, false, false, gdst, false, nullptr, nullptr, "bwd_tc_dx_mlp1", st));
    if (p->ln_gamma) SF_TRY(launch_ln_bwd(p->in, p->ln_gamma, p->ln_beta, gn, bp->g_in, bp->g_ln_gamma, bp->g_ln_beta, M, C, p->ln_eps, false, bp->add_to_g_in, st));
    else if (bp->add_to_g_in) SF_TRY(sf_add(bp->g_in, bp->add_to_g_in, bp->g_in, M * C, (void*)st));
    return SF_OK;
}


// ---- window attention ------------------------------------------------------------------------------------------------------
// The eight fp32 tensors the attention-core adjoint works on are column blocks of ONE [tokens x 8*inner] buffer
// [Q | K | V | gO | O | dQ | dK | dV], and dQ | dK | dV are cast into ONE bf16 tiled tensor, so that the projections
// run as stacked GEMMs: q|k|v recompute in one launch (self) or two (cross: q from one source, k|v from the other), the
// data gradient as one K-concatenated GEMM per source, the three weight gradients as one stacked k_tc_wgrad per source.
static inline void wa_bwd_ln_plan(const sf_window_attn_params* p, bool* need_q, bool* need_kv, bool* share);
struct WaBwdPlan {
    PackedGemm wqkv, wq, wkv, wot, wqkvt, wqt, wkvt;
    bool share;
    size_t off_pk, off_nq, off_nkv, off_g, off_o, off_dqkv, off_f32, off_gn, total;
};
static WaBwdPlan wa_bwd_tc_plan(const sf_window_attn_params* p) {
    WaBwdPlan w{};
    const long long M = (long long)p->B * p->Hp * p->Wp;
    const int C = p->C, inner = p->num_heads * p->head_dim;
    bool nq, nkv;
    wa_bwd_ln_plan(p, &nq, &nkv, &w.share);
    Carver pc;
    if (w.share) {
        w.wqkv = plan_packed(pc, 3 * inner, C);     // q | k | v = n [Wq; Wk; Wv]^T + b
        w.wqkvt = plan_packed(pc, C, 3 * inner);    // g_n = [dQ | dK | dV] [Wq; Wk; Wv]
    } else {
        w.wq = plan_packed(pc, inner, C); w.wkv = plan_packed(pc, 2 * inner, C);
        w.wqt = plan_packed(pc, C, inner); w.wkvt = plan_packed(pc, C, 2 * inner);
    }
    w.wot = plan_packed(pc, inner, C);              // g_O = gout W_o
    Carver c;
    w.off_pk = c.take(pc.off);
    w.off_nq = c.take(tiled_bytes(M, C));
    w.off_nkv = c.take(w.share ? 0 : tiled_bytes(M, C));
    w.off_g = c.take(tiled_bytes(M, C));
    w.off_o = c.take(tiled_bytes(M, inner));
    w.off_dqkv = c.take(tiled_bytes(M, 3 * inner));
    w.off_f32 = c.take(align_up((size_t)M * 4 * inner * sizeof(float)));
    w.off_gn = c.take(2 * align_up((size_t)M * C * sizeof(float)));
    w.total = c.off;
    return w;
}
static bool wa_bwd_tc_ok(const sf_window_attn_params* p) {
    const int inner = p->num_heads * p->head_dim;
    const long long M = (long long)p->B * p->Hp * p->Wp;
    return bwd_tc_enabled() && p->C % 4 == 0 && (int)tc::pad16((uint32_t)p->C) <= TC_MAX_KPAD && inner % 8 == 0 &&
           (int)tc::pad16((uint32_t)inner) <= TC_MAX_KPAD && aligned16(p->q_src) && aligned16(p->kv_src) && M < 2147483647LL &&
           M * 4 * inner < (1LL << 32);
}
static int launch_attn_core_bwd(const float* Q, const float* K, const float* V, const float* gO, float* dQ, float* dK, float* dV,
                                const float* table, float* gtable, const WinGeom& g, int inner, int nh, int d, cudaStream_t st, bool mma,
                                float* O_out);
static int window_attn_bwd_tc(const sf_window_attn_bwd_params* bp, void* ws_ptr, size_t ws_bytes, cudaStream_t st) {
    const sf_window_attn_params* p = &bp->fwd;
    const long long M = (long long)p->B * p->Hp * p->Wp;
    const int C = p->C, inner = p->num_heads * p->head_dim, ld = 4 * inner;
    const WaBwdPlan w = wa_bwd_tc_plan(p);
    if (ws_bytes < w.total || !ws_ptr) { set_error("sf_window_attn_bwd: workspace too small (%zu B given, %zu needed)", ws_bytes, w.total); return SF_ERR_WORKSPACE; }
    char* base = reinterpret_cast<char*>(ws_ptr);
    char* pk = base + w.off_pk;
    bf16* nq_t = reinterpret_cast<bf16*>(base + w.off_nq);
    bf16* nkv_t = w.share ? nq_t : reinterpret_cast<bf16*>(base + w.off_nkv);
    bf16* g_t = reinterpret_cast<bf16*>(base + w.off_g);
    bf16* o_t = reinterpret_cast<bf16*>(base + w.off_o);
    bf16* dqkv_t = reinterpret_cast<bf16*>(base + w.off_dqkv);
    float* F = reinterpret_cast<float*>(base + w.off_f32);
    float *Q = F, *K = F + inner, *V = F + 2 * inner, *gO = F + 3 * inner;
    const size_t cs = align_up((size_t)M * C * sizeof(float));
    float* gnq = reinterpret_cast<float*>(base + w.off_gn);
    float* gnkv = reinterpret_cast<float*>(base + w.off_gn + cs);
    bool need_q, need_kv, share;
    wa_bwd_ln_plan(p, &need_q, &need_kv, &share);
    const bool self_src = p->kv_src == p->q_src;
    SF_CHECK_ARG(share || bp->g_kv_src || self_src, "sf_window_attn_bwd: g_kv_src is required for cross attention");
    const float* Wqkv[3] = {p->wq, p->wk, p->wv};
    const float* bqkv[3] = {p->bq, p->bk, p->bv};
    PackJobs jobs{};
    if (share) {
        pack_fwd_stack(jobs, 3, Wqkv, bqkv, inner, C, w.wqkv, pk);
        pack_tr_stack(jobs, 3, Wqkv, inner, C, w.wqkvt, pk);
    } else {
        pack_fwd(jobs, p->wq, p->bq, inner, C, w.wq, pk);
        pack_fwd_stack(jobs, 2, Wqkv + 1, bqkv + 1, inner, C, w.wkv, pk);
        pack_tr(jobs, p->wq, inner, C, w.wqt, pk);
        pack_tr_stack(jobs, 2, Wqkv + 1, inner, C, w.wkvt, pk);
    }
    pack_tr(jobs, p->wo, C, inner, w.wot, pk);
    SF_TRY(launch_pack_jobs(jobs, st));
    // a GEMM whose fp32 output is a column block of F
    auto gemm_to_F = [&](const bf16* A, const PackedGemm& g, int N, float* out, const char* name) -> int {
        TcGemm t{};
        t.A = A; t.M = M; t.K = C; t.a_mode = AM_TILED;
        bind_packed(t, g, pk);
        t.N = N; t.out_mode = OUT_F32; t.out = out; t.ldo = ld;
        SF_TRY(tc_gemm_plan(&t));
        return launch_tc_gemm(t, name, st);
    };
    // ---- recompute the forward: normalised operands (bf16 tiles), q | k | v (fp32, for the attention-core adjoint) -------------
    SF_TRY(launch_ln_to_tiled(p->q_src, p->ln_q_gamma, p->ln_q_beta, nq_t, M, C, p->ln_eps, st));
    if (share) {
        SF_TRY(gemm_to_F(nq_t, w.wqkv, 3 * inner, Q, "bwd_tc_recompute_qkv"));
    } else {
        SF_TRY(launch_ln_to_tiled(p->kv_src, p->ln_kv_gamma, p->ln_kv_beta, nkv_t, M, C, p->ln_eps, st));
        SF_TRY(gemm_to_F(nq_t, w.wq, inner, Q, "bwd_tc_recompute_qkv"));
        SF_TRY(gemm_to_F(nkv_t, w.wkv, 2 * inner, K, "bwd_tc_recompute_qkv"));
    }
    WinGeom geom = make_geom(p->B, p->Hp, p->Wp, p->wsh, p->wsw, p->shift);
    SF_CHECK_ARG(attn_core_bwd_mma_supported(geom, p->head_dim, p->num_heads), "sf_window_attn_bwd: head shape not built for the tensor-core adjoint");
    // ---- output projection: gradient w.r.t. O ---------------------------------------------------------------------------------
    SF_TRY(launch_ln_to_tiled(bp->gout, nullptr, nullptr, g_t, M, C, 0.f, st));
    {
        TcGemm t{};
        t.A = g_t; t.M = M; t.K = C; t.a_mode = AM_TILED;
        bind_packed(t, w.wot, pk);
        t.bias = nullptr;
        t.N = inner; t.out_mode = OUT_F32; t.out = gO; t.ldo = ld;
        SF_TRY(tc_gemm_plan(&t));
        SF_TRY(launch_tc_gemm(t, "bwd_tc_dx_proj", st));
    }
    // ---- attention core (fp16 mma.sync adjoint; also delivers O = P V) ---------------------------------------------------------------
    // its outputs (O = P V, dQ | dK | dV) leave as bf16 tiled tensors: the operands of the GEMMs below, no cast pass.
    // Tail rows of the last tile and the padding columns are zeroed first (the weight-gradient GEMM sums over all rows).
    const int nkc3 = (int)tc::pad16((uint32_t)(3 * inner)) / 8, nkc1 = (int)tc::pad16((uint32_t)inner) / 8;
    {
        const long long tiles = (M + 127) / 128;
        const bool tail = (M & 127) != 0;
        const bool padc = tc::pad16((uint32_t)inner) != (uint32_t)inner || tc::pad16((uint32_t)(3 * inner)) != (uint32_t)(3 * inner);
        if (padc) {
            if (cudaMemsetAsync(o_t, 0, tiled_bytes(M, inner), st) != cudaSuccess || cudaMemsetAsync(dqkv_t, 0, tiled_bytes(M, 3 * inner), st) != cudaSuccess) {
                set_error("sf_window_attn_bwd: cudaMemsetAsync failed"); return SF_ERR_CUDA;
            }
        } else if (tail) {
            if (cudaMemsetAsync(o_t + (size_t)(tiles - 1) * nkc1 * 1024, 0, (size_t)nkc1 * 2048, st) != cudaSuccess ||
                cudaMemsetAsync(dqkv_t + (size_t)(tiles - 1) * nkc3 * 1024, 0, (size_t)nkc3 * 2048, st) != cudaSuccess) {
                set_error("sf_window_attn_bwd: cudaMemsetAsync failed"); return SF_ERR_CUDA;
            }
        }
    }
    AttnBwdTiledOut to{dqkv_t, o_t, (unsigned)nkc3, (unsigned)nkc1, inner};
    SF_TRY(launch_attn_core_bwd_mma(Q, K, V, gO, nullptr, nullptr, nullptr, nullptr, p->bias_table, bp->g_bias_table, geom, inner, p->num_heads,
                                    p->head_dim, st, ld, &to));
    // ---- weight gradients -------------------------------------------------------------------------------------------------------
    SF_TRY(launch_tc_wgrad(g_t, o_t, bp->g_wo, bp->g_bo, M, C, inner, "bwd_tc_wgrad", st));
    if (share) {
        TcWgradArgs a{};
        a.G = dqkv_t; a.A = nq_t; a.M = M; a.N = 3 * inner; a.K = C; a.nout = 3;
        a.Wg[0] = bp->g_wq; a.Wg[1] = bp->g_wk; a.Wg[2] = bp->g_wv;
        a.bias_grad[0] = bp->g_bq; a.bias_grad[1] = bp->g_bk; a.bias_grad[2] = bp->g_bv;
        SF_TRY(launch_tc_wgrad_ex(a, "bwd_tc_wgrad", st));
    } else {
        TcWgradArgs a{};
        a.G = dqkv_t; a.A = nq_t; a.M = M; a.N = inner; a.K = C; a.nout = 1; a.g_tile_nkc = nkc3; a.g_kc0 = 0;
        a.Wg[0] = bp->g_wq; a.bias_grad[0] = bp->g_bq;
        SF_TRY(launch_tc_wgrad_ex(a, "bwd_tc_wgrad", st));
        TcWgradArgs b2{};
        b2.G = dqkv_t; b2.A = nkv_t; b2.M = M; b2.N = 2 * inner; b2.K = C; b2.nout = 2; b2.g_tile_nkc = nkc3; b2.g_kc0 = inner / 8;
        b2.Wg[0] = bp->g_wk; b2.Wg[1] = bp->g_wv; b2.bias_grad[0] = bp->g_bk; b2.bias_grad[1] = bp->g_bv;
        SF_TRY(launch_tc_wgrad_ex(b2, "bwd_tc_wgrad", st));
    }
    // ---- gradients w.r.t. the (normalised) operands: K-concatenated GEMMs over [dQ | dK | dV] ---------------------------------------
    auto gemm_dx = [&](int kc0, int K, const PackedGemm& g, float* out) -> int {
        TcGemm t{};
        t.A = dqkv_t; t.M = M; t.K = K; t.a_mode = AM_TILED; t.a_tile_nkc = nkc3; t.a_kc0 = kc0;
        bind_packed(t, g, pk);
        t.bias = nullptr;
        t.N = C; t.out_mode = OUT_F32; t.out = out; t.ldo = C;
        SF_TRY(tc_gemm_plan(&t));
        return launch_tc_gemm(t, "bwd_tc_dx_qkv", st);
    };
    float* gq_dst = need_q ? gnq : bp->g_q_src;
    if (share) {
        SF_TRY(gemm_dx(0, 3 * inner, w.wqkvt, gq_dst));
    } else {
        float* gkv_dst = need_kv ? gnkv : (bp->g_kv_src ? bp->g_kv_src : gnkv);
        SF_TRY(gemm_dx(0, inner, w.wqt, gq_dst));
        SF_TRY(gemm_dx(inner / 8, 2 * inner, w.wkvt, gkv_dst));
    }
    // ---- LayerNorm adjoints (as in the exact path) ----------------------------------------------------------------------------------
    if (need_q) SF_TRY(launch_ln_bwd(p->q_src, p->ln_q_gamma, p->ln_q_beta, gnq, bp->g_q_src, bp->g_ln_q_gamma, bp->g_ln_q_beta, M, C, p->ln_eps, false, bp->add_to_g_q_src, st));
    else if (bp->add_to_g_q_src) SF_TRY(sf_add(bp->g_q_src, bp->add_to_g_q_src, bp->g_q_src, M * C, (void*)st));
    if (!share) {
        float* dst = bp->g_kv_src ? bp->g_kv_src : bp->g_q_src;
        const bool accum = bp->g_kv_src == nullptr;
        if (need_kv) {
            SF_TRY(launch_ln_bwd(p->kv_src, p->ln_kv_gamma, p->ln_kv_beta, gnkv, dst, bp->g_ln_kv_gamma, bp->g_ln_kv_beta, M, C, p->ln_eps, false, accum ? dst : nullptr, st));
        } else if (accum) {
            SF_TRY(sf_add(bp->g_q_src, gnkv, bp->g_q_src, M * C, (void*)st));
        }
    }
    return SF_OK;
}

// =============================================================================================
// operator backward: window attention
// =============================================================================================
static inline void wa_bwd_ln_plan(const sf_window_attn_params* p, bool* need_q, bool* need_kv, bool* share) {
    const bool self_attn = (p->kv_src == p->q_src);
    const bool same_ln = self_attn && p->ln_q_gamma == p->ln_kv_gamma && p->ln_q_beta == p->ln_kv_beta;
    *need_q = p->ln_q_gamma != nullptr;
    *share = same_ln;                           // kv operand is the very tensor the q operand is
    *need_kv = p->ln_kv_gamma != nullptr && !same_ln;
}

size_t window_attn_bwd_ws(const sf_window_attn_bwd_params* bp) {
    const sf_window_attn_params* p = &bp->fwd;
    if (p->precision == SF_PREC_BF16 && wa_bwd_tc_ok(p) &&
        attn_core_bwd_mma_supported(make_geom(p->B, p->Hp, p->Wp, p->wsh, p->wsw, p->shift), p->head_dim, p->num_heads))
        return wa_bwd_tc_plan(p).total;
    const size_t M = (size_t)p->B * p->Hp * p->Wp, inner = (size_t)p->num_heads * p->head_dim;
    return 8 * align_up(M * inner * sizeof(float)) + 4 * align_up(M * p->C * sizeof(float));
}

int window_attn_bwd(const sf_window_attn_bwd_params* bp, void* ws_ptr, size_t ws_bytes, cudaStream_t st) {
    const bool tf = bp->fwd.precision == SF_PREC_BF16;   // bf16 operators: tensor-core GEMMs in the backward pass
    const sf_window_attn_params* p = &bp->fwd;
    if (tf && wa_bwd_tc_ok(p) && attn_core_bwd_mma_supported(make_geom(p->B, p->Hp, p->Wp, p->wsh, p->wsw, p->shift), p->head_dim, p->num_heads))
        return window_attn_bwd_tc(bp, ws_ptr, ws_bytes, st);
    const long long M = (long long)p->B * p->Hp * p->Wp;
    const int C = p->C, inner = p->num_heads * p->head_dim;
    Workspace ws(ws_ptr, ws_bytes);
    float* Q = ws.take<float>((size_t)M * inner);
    float* K = ws.take<float>((size_t)M * inner);
    float* V = ws.take<float>((size_t)M * inner);
    float* O = ws.take<float>((size_t)M * inner);
    float* gO = ws.take<float>((size_t)M * inner);
    float* dQ = ws.take<float>((size_t)M * inner);
    float* dK = ws.take<float>((size_t)M * inner);
    float* dV = ws.take<float>((size_t)M * inner);
    float* nqb = ws.take<float>((size_t)M * C);
    float* nkvb = ws.take<float>((size_t)M * C);
    float* gnq = ws.take<float>((size_t)M * C);
    float* gnkv = ws.take<float>((size_t)M * C);
    if (!Q || !gnkv) { set_error("sf_window_attn_bwd: workspace too small (%zu B given)", ws_bytes); return SF_ERR_WORKSPACE; }
    bool need_q, need_kv, share;
    wa_bwd_ln_plan(p, &need_q, &need_kv, &share);
    const bool self_src = p->kv_src == p->q_src;
    SF_CHECK_ARG(share || bp->g_kv_src || self_src, "sf_window_attn_bwd: g_kv_src is required for cross attention");
    // ---- recompute the forward ------------------------------------------------------------------------
    const float* nq = p->q_src;
    const float* nkv = p->kv_src;
    if (need_q) { SF_TRY(launch_layernorm(p->q_src, p->ln_q_gamma, p->ln_q_beta, nqb, M, C, p->ln_eps, 0, nullptr, st)); nq = nqb; }
    if (share) nkv = nq;
    else if (need_kv) { SF_TRY(launch_layernorm(p->kv_src, p->ln_kv_gamma, p->ln_kv_beta, nkvb, M, C, p->ln_eps, 0, nullptr, st)); nkv = nkvb; }
    GemmBatch gb{};
    gb.p[0] = GemmProblem{nq, p->wq, p->bq, nullptr, Q};
    gb.p[1] = GemmProblem{nkv, p->wk, p->bk, nullptr, K};
    gb.p[2] = GemmProblem{nkv, p->wv, p->bv, nullptr, V};
    SF_TRY(tf ? gemm_tf32_nt(gb, 3, M, inner, C, st) : launch_gemm_tn(gb, 3, M, inner, C, false, st));
    WinGeom geom = make_geom(p->B, p->Hp, p->Wp, p->wsh, p->wsw, p->shift);
    // the tensor-core adjoint of the attention core also delivers O = P V (one more product on fragments it already holds)
    const bool fused_o = tf && attn_core_bwd_mma_supported(geom, p->head_dim, p->num_heads);
    if (!fused_o) SF_TRY(launch_attn_core_f32(Q, K, V, O, p->bias_table, geom, inner, p->num_heads, p->head_dim, st));
    // ---- output projection: gradient w.r.t. O -----------------------------------------------------------
    SF_TRY(launch_gemm_nn(bp->gout, p->wo, nullptr, gO, M, inner, C, false, st, tf));
    // ---- attention core ---------------------------------------------------------------------------------
    SF_TRY(launch_attn_core_bwd(Q, K, V, gO, dQ, dK, dV, p->bias_table, bp->g_bias_table, geom, inner, p->num_heads, p->head_dim, st, tf,
                                fused_o ? O : nullptr));
    SF_TRY(launch_gemm_tn_reduce(bp->gout, O, bp->g_wo, M, C, inner, false, st, tf, bp->g_bo));
    // ---- projections --------------------------------------------------------------------------------------
    SF_TRY(launch_gemm_tn_reduce(dQ, nq, bp->g_wq, M, inner, C, false, st, tf, bp->g_bq));
    SF_TRY(launch_gemm_tn_reduce(dK, nkv, bp->g_wk, M, inner, C, false, st, tf, bp->g_bk));
    SF_TRY(launch_gemm_tn_reduce(dV, nkv, bp->g_wv, M, inner, C, false, st, tf, bp->g_bv));
    // gradients w.r.t. the (normalised) operands
    float* gq_dst = need_q ? gnq : bp->g_q_src;
    SF_TRY(launch_gemm_nn(dQ, p->wq, nullptr, gq_dst, M, C, inner, false, st, tf));
    if (share) {
        SF_TRY(launch_gemm_nn(dK, p->wk, nullptr, gq_dst, M, C, inner, true, st, tf));
        SF_TRY(launch_gemm_nn(dV, p->wv, nullptr, gq_dst, M, C, inner, true, st, tf));
    } else {
        // distinct kv operand (other tensor and/or other LayerNorm)
        float* gkv_dst = need_kv ? gnkv : (bp->g_kv_src ? bp->g_kv_src : gnkv);
        SF_TRY(launch_gemm_nn(dK, p->wk, nullptr, gkv_dst, M, C, inner, false, st, tf));
        SF_TRY(launch_gemm_nn(dV, p->wv, nullptr, gkv_dst, M, C, inner, true, st, tf));
    }
    // ---- LayerNorm adjoints -----------------------------------------------------------------------------
    // add_to_g_q_src: the residual branch's gradient rides along instead of a separate add kernel
    if (need_q) SF_TRY(launch_ln_bwd(p->q_src, p->ln_q_gamma, p->ln_q_beta, gnq, bp->g_q_src, bp->g_ln_q_gamma, bp->g_ln_q_beta, M, C, p->ln_eps, false, bp->add_to_g_q_src, st));
    else if (bp->add_to_g_q_src) SF_TRY(sf_add(bp->g_q_src, bp->add_to_g_q_src, bp->g_q_src, M * C, (void*)st));
    if (!share) {
        // where does the kv-side gradient land?  g_kv_src if given, else (same source tensor) it is added to g_q_src
        float* dst = bp->g_kv_src ? bp->g_kv_src : bp->g_q_src;
        const bool accum = bp->g_kv_src == nullptr;
        if (need_kv) {
            SF_TRY(launch_ln_bwd(p->kv_src, p->ln_kv_gamma, p->ln_kv_beta, gnkv, dst, bp->g_ln_kv_gamma, bp->g_ln_kv_beta, M, C, p->ln_eps, false, accum ? dst : nullptr, st));
        } else if (accum) {
            SF_TRY(sf_add(bp->g_q_src, gnkv, bp->g_q_src, M * C, (void*)st));
        }
    }
    return SF_OK;
}

// =============================================================================================
// operator backward: MLP
// =============================================================================================
size_t mlp_bwd_ws(const sf_mlp_bwd_params* bp) {
    const sf_mlp_params* p = &bp->fwd;
    if (p->precision == SF_PREC_BF16 && mlp_bwd_tc_ok(p)) return mlp_bwd_tc_plan(p).total;
    return 2 * align_up((size_t)p->M * p->hidden * sizeof(float)) + 2 * align_up((size_t)p->M * p->C * sizeof(float));
}

int mlp_bwd(const sf_mlp_bwd_params* bp, void* ws_ptr, size_t ws_bytes, cudaStream_t st) {
    const bool tf = bp->fwd.precision == SF_PREC_BF16;   // bf16 operators: tensor-core GEMMs in the backward pass
    const sf_mlp_params* p = &bp->fwd;
    if (tf && mlp_bwd_tc_ok(p)) return mlp_bwd_tc(bp, ws_ptr, ws_bytes, st);
    const long long M = p->M;
    const int C = p->C, H = p->hidden;
    Workspace ws(ws_ptr, ws_bytes);
    float* hpre = ws.take<float>((size_t)M * H);
    float* gh = ws.take<float>((size_t)M * H);
    float* nb = ws.take<float>((size_t)M * C);
    float* gn = ws.take<float>((size_t)M * C);
    if (!hpre || !gn) { set_error("sf_mlp_bwd: workspace too small (%zu B given)", ws_bytes); return SF_ERR_WORKSPACE; }
    const float* n = p->in;
    if (p->ln_gamma) { SF_TRY(launch_layernorm(p->in, p->ln_gamma, p->ln_beta, nb, M, C, p->ln_eps, 0, nullptr, st)); n = nb; }
    GemmBatch g1{};
    g1.p[0] = GemmProblem{n, p->w1, p->b1, nullptr, hpre};
    SF_TRY(tf ? gemm_tf32_nt(g1, 1, M, H, C, st) : launch_gemm_tn(g1, 1, M, H, C, false, st));
    SF_TRY(launch_gemm_tn_reduce(bp->gout, hpre, bp->g_w2, M, C, H, true, st, tf, bp->g_b2));      // gW2 = gout^T ELU(hpre), gb2
    SF_TRY(launch_gemm_nn(bp->gout, p->w2, hpre, gh, M, H, C, false, st, tf));           // g_hpre = (gout W2) o ELU'(hpre)
    SF_TRY(launch_gemm_tn_reduce(gh, n, bp->g_w1, M, H, C, false, st, tf, bp->g_b1));
    float* gdst = p->ln_gamma ? gn : bp->g_in;
    SF_TRY(launch_gemm_nn(gh, p->w1, nullptr, gdst, M, C, H, false, st, tf));
    if (p->ln_gamma) SF_TRY(launch_ln_bwd(p->in, p->ln_gamma, p->ln_beta, gn, bp->g_in, bp->g_ln_gamma, bp->g_ln_beta, M, C, p->ln_eps, false, bp->add_to_g_in, st));
    else if (bp->add_to_g_in) SF_TRY(sf_add(bp->g_in, bp->add_to_g_in, bp->g_in, M * C, (void*)st));
    return SF_OK;
}

// =============================================================================================
// operator backward: patch layers
// =============================================================================================
size_t patch_bwd_ws(const sf_patch_bwd_params* bp) {
    const sf_patch_params* p = &bp->fwd;
    const int mm = p->mh * p->mw;
    const size_t Mr = p->encoder ? (size_t)p->B * (p->H / p->mh) * (p->W / p->mw) : (size_t)p->B * p->H * p->W;
    const size_t K = p->encoder ? (size_t)mm * p->Cin : (size_t)p->Cin, N = p->encoder ? (size_t)p->Cout : (size_t)mm * p->Cout;
    return 2 * align_up(Mr * K * sizeof(float)) + 3 * align_up(Mr * N * sizeof(float));
}

int patch_bwd(const sf_patch_bwd_params* bp, void* ws_ptr, size_t ws_bytes, cudaStream_t st) {
    const bool tf = bp->fwd.precision == SF_PREC_BF16;   // bf16 operators: TF32 tensor-core GEMMs in the backward pass
    const sf_patch_params* p = &bp->fwd;
    const int mm = p->mh * p->mw;
    const long long Mr = p->encoder ? (long long)p->B * (p->H / p->mh) * (p->W / p->mw) : (long long)p->B * p->H * p->W;
    const int K = p->encoder ? mm * p->Cin : p->Cin, N = p->encoder ? p->Cout : mm * p->Cout;
    Workspace ws(ws_ptr, ws_bytes);
    float* Abuf = ws.take<float>((size_t)Mr * K);
    float* gA = ws.take<float>((size_t)Mr * K);
    float* lin = ws.take<float>((size_t)Mr * N);
    float* glin = ws.take<float>((size_t)Mr * N);
    float* gpost = ws.take<float>((size_t)Mr * N);
    if (!Abuf || !gpost) { set_error("sf_patch_bwd: workspace too small (%zu B given)", ws_bytes); return SF_ERR_WORKSPACE; }
    const float* A = p->in;
    if (p->encoder) { SF_TRY(sf_patch_merge(p->in, Abuf, p->B, p->H, p->W, p->Cin, p->mh, p->mw, (void*)st)); A = Abuf; }
    GemmBatch g{};
    g.p[0] = GemmProblem{A, p->w, p->b, nullptr, lin};
    SF_TRY(tf ? gemm_tf32_nt(g, 1, Mr, N, K, st) : launch_gemm_tn(g, 1, Mr, N, K, false, st));
    // gradient w.r.t. the LayerNorm output rows (ELU' is applied inside the LN backward kernel)
    const float* gy = bp->gout;
    if (!p->encoder) { SF_TRY(sf_patch_merge(bp->gout, gpost, p->B, p->H * p->mh, p->W * p->mw, p->Cout, p->mh, p->mw, (void*)st)); gy = gpost; }
    SF_TRY(launch_ln_bwd(lin, p->ln_gamma, p->ln_beta, gy, glin, bp->g_ln_gamma, bp->g_ln_beta, Mr, N, p->ln_eps, true, nullptr, st));
    SF_TRY(launch_gemm_tn_reduce(glin, A, bp->g_w, Mr, N, K, false, st, tf, bp->g_b));
    if (p->encoder) {
        SF_TRY(launch_gemm_nn(glin, p->w, nullptr, gA, Mr, K, N, false, st, tf));
        SF_TRY(sf_patch_unmerge(gA, bp->g_in, p->B, p->H / p->mh, p->W / p->mw, p->Cin, p->mh, p->mw, (void*)st));
    } else {
        SF_TRY(launch_gemm_nn(glin, p->w, nullptr, bp->g_in, Mr, K, N, false, st, tf));
    }
    return SF_OK;
}

// =============================================================================================
// operator backward: final head (a013:126-152)
// =============================================================================================
static constexpr int HB_THREADS = 256;
static constexpr int HB_MAXK = 7;

// forward conv1 recompute: t[pix] = conv(x,y) (same arithmetic as k_head_conv1)
__global__ void k_hb_conv1(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ w1, const float* __restrict__ b1,
                           float2* __restrict__ t, int B, int H, int W, int ks) {
    const int pad = ks / 2;
    long long total = (long long)B * H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % W);
        long long p = i / W;
        int r = (int)(p % H);
        long long b = p / H;
        const float* xb = x + b * H * W;
        const float* yb = y + b * H * W;
        float o0 = b1[0], o1 = b1[1];
        for (int dr = 0; dr < ks; dr++) {
            int rr = reflect_both(r + dr - pad, H);
            for (int dc = 0; dc < ks; dc++) {
                int cc = reflect_both(c + dc - pad, W);
                float xv = xb[(long long)rr * W + cc], yv = yb[(long long)rr * W + cc];
                int wi = dr * ks + dc;
                o0 = fmaf(w1[(0 * 2 + 0) * ks * ks + wi], xv, o0); o0 = fmaf(w1[(0 * 2 + 1) * ks * ks + wi], yv, o0);
                o1 = fmaf(w1[(1 * 2 + 0) * ks * ks + wi], xv, o1); o1 = fmaf(w1[(1 * 2 + 1) * ks * ks + wi], yv, o1);
            }
        }
        t[i] = make_float2(o0, o1);
    }
}

__device__ __forceinline__ void block_atomic_add(float v, float* dst, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) a += red[k];
        atomicAdd(dst, a);
    }
}

// conv2 adjoint: ga[p'] += gout[p] * w2[c][tap] (scatter, reflect), gw2[c][tap] += gout[p] * a[p'], gb2 += gout
__global__ void k_hb_conv2_bwd(const float* __restrict__ gout, const float2* __restrict__ t, const float* __restrict__ affine,
                               const float* __restrict__ w2, float* __restrict__ ga, float* __restrict__ gw2, float* __restrict__ gb2,
                               int B, int H, int W, int ks) {
    __shared__ float red[HB_THREADS / 32];
    const float sc0 = affine[0], sh0 = affine[1], sc1 = affine[2], sh1 = affine[3];
    const int pad = ks / 2;
    long long total = (long long)B * H * W;
    float gb = 0.f;
    float gw[2 * HB_MAXK * HB_MAXK];
    for (int i = 0; i < 2 * ks * ks; i++) gw[i] = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % W);
        long long p = i / W;
        int r = (int)(p % H);
        long long b = p / H;
        const float go = gout[i];
        gb += go;
        for (int dr = 0; dr < ks; dr++) {
            int rr = reflect_both(r + dr - pad, H);
            for (int dc = 0; dc < ks; dc++) {
                int cc = reflect_both(c + dc - pad, W);
                long long q = (b * H + rr) * W + cc;
                float2 tv = t[q];
                float a0 = elu1(fmaf(tv.x, sc0, sh0)), a1 = elu1(fmaf(tv.y, sc1, sh1));
                gw[dr * ks + dc] += go * a0;
                gw[ks * ks + dr * ks + dc] += go * a1;
                atomicAdd(&ga[2 * q], go * w2[dr * ks + dc]);
                atomicAdd(&ga[2 * q + 1], go * w2[ks * ks + dr * ks + dc]);
            }
        }
    }
    block_atomic_add(gb, gb2, red);
    for (int i = 0; i < 2 * ks * ks; i++) block_atomic_add(gw[i], &gw2[i], red);
}

// gz = ga * ELU'(z); per-channel sums: stats[0..1] = sum gz, stats[2..3] = sum gz * zhat
__global__ void k_hb_bn_stats(const float* __restrict__ ga, const float2* __restrict__ t, const float* __restrict__ affine,
                              const float* __restrict__ mean, const float* __restrict__ invstd, float* __restrict__ gz,
                              float* __restrict__ stats, long long total) {
    __shared__ float red[HB_THREADS / 32];
    const float sc0 = affine[0], sh0 = affine[1], sc1 = affine[2], sh1 = affine[3];
    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        float2 tv = t[i];
        float g0 = ga[2 * i] * elu_grad(fmaf(tv.x, sc0, sh0)), g1 = ga[2 * i + 1] * elu_grad(fmaf(tv.y, sc1, sh1));
        gz[2 * i] = g0; gz[2 * i + 1] = g1;
        s0 += g0; s1 += g1;
        q0 += g0 * (tv.x - mean[0]) * invstd[0];
        q1 += g1 * (tv.y - mean[1]) * invstd[1];
    }
    block_atomic_add(s0, &stats[0], red);
    block_atomic_add(s1, &stats[1], red);
    block_atomic_add(q0, &stats[2], red);
    block_atomic_add(q1, &stats[3], red);
}

// gt = BN adjoint of gz (training: batch statistics; eval: gz * scale); then conv1 adjoint (scatter)
__global__ void k_hb_conv1_bwd(const float* __restrict__ gz, const float2* __restrict__ t, const float* __restrict__ stats,
                               const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ invstd,
                               const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ w1,
                               float* __restrict__ gx, float* __restrict__ gy, float* __restrict__ gw1, float* __restrict__ gb1,
                               int B, int H, int W, int ks, int training) {
    __shared__ float red[HB_THREADS / 32];
    const int pad = ks / 2;
    long long total = (long long)B * H * W;
    const float invn = 1.f / (float)total;
    float gb[2] = {0.f, 0.f};
    float gw[4 * HB_MAXK * HB_MAXK];
    for (int i = 0; i < 4 * ks * ks; i++) gw[i] = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % W);
        long long p = i / W;
        int r = (int)(p % H);
        long long b = p / H;
        float2 tv = t[i];
        float gt[2];
#pragma unroll
        for (int ch = 0; ch < 2; ch++) {
            float g = gz[2 * i + ch];
            float tvc = ch == 0 ? tv.x : tv.y;
            if (training) {
                float zh = (tvc - mean[ch]) * invstd[ch];
                gt[ch] = gamma[ch] * invstd[ch] * (g - stats[ch] * invn - zh * stats[2 + ch] * invn);
            } else {
                gt[ch] = g * gamma[ch] * invstd[ch];
            }
            gb[ch] += gt[ch];
        }
        const float* xb = x + b * H * W;
        const float* yb = y + b * H * W;
        for (int dr = 0; dr < ks; dr++) {
            int rr = reflect_both(r + dr - pad, H);
            for (int dc = 0; dc < ks; dc++) {
                int cc = reflect_both(c + dc - pad, W);
                long long q = (b * H + rr) * W + cc;
                int wi = dr * ks + dc;
                float xv = xb[(long long)rr * W + cc], yv = yb[(long long)rr * W + cc];
                gw[(0 * 2 + 0) * ks * ks + wi] += gt[0] * xv; gw[(0 * 2 + 1) * ks * ks + wi] += gt[0] * yv;
                gw[(1 * 2 + 0) * ks * ks + wi] += gt[1] * xv; gw[(1 * 2 + 1) * ks * ks + wi] += gt[1] * yv;
                atomicAdd(&gx[q], gt[0] * w1[(0 * 2 + 0) * ks * ks + wi] + gt[1] * w1[(1 * 2 + 0) * ks * ks + wi]);
                atomicAdd(&gy[q], gt[0] * w1[(0 * 2 + 1) * ks * ks + wi] + gt[1] * w1[(1 * 2 + 1) * ks * ks + wi]);
            }
        }
    }
    block_atomic_add(gb[0], &gb1[0], red);
    block_atomic_add(gb[1], &gb1[1], red);
    for (int i = 0; i < 4 * ks * ks; i++) block_atomic_add(gw[i], &gw1[i], red);
}

// tiny helper: affine / mean / invstd vectors for the backward, and the BN affine gradients
__global__ void k_hb_prepare(const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ running_mean,
                             const float* __restrict__ running_var, const float* __restrict__ save_mean,
                             const float* __restrict__ save_invstd, float* __restrict__ vec, float eps, int training) {
    // vec: [0..3] affine (sc0, sh0, sc1, sh1), [4..5] mean, [6..7] invstd, [8..11] stats (zeroed)
    int c = threadIdx.x;
    if (c < 2) {
        float mean = training ? save_mean[c] : running_mean[c];
        float invstd = training ? save_invstd[c] : 1.0f / sqrtf(running_var[c] + eps);
        float sc = gamma[c] * invstd;
        vec[2 * c] = sc;
        vec[2 * c + 1] = beta[c] - mean * sc;
        vec[4 + c] = mean;
        vec[6 + c] = invstd;
    }
    if (c < 4) vec[8 + c] = 0.f;
}

__global__ void k_hb_bn_param_grads(const float* __restrict__ stats, float* __restrict__ ggamma, float* __restrict__ gbeta) {
    int c = threadIdx.x;
    if (c < 2) {
        if (gbeta) atomicAdd(&gbeta[c], stats[c]);
        if (ggamma) atomicAdd(&ggamma[c], stats[2 + c]);
    }
}

__global__ void k_zero(float* __restrict__ p, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = 0.f;
}

size_t head_bwd_ws(const sf_head_bwd_params* bp) {
    const sf_head_params* p = &bp->fwd;
    const size_t total = (size_t)p->B * p->H * p->W;
    return 3 * align_up(total * 2 * sizeof(float)) + align_up(16 * sizeof(float));
}

int head_bwd(const sf_head_bwd_params* bp, void* ws_ptr, size_t ws_bytes, cudaStream_t st) {
    const sf_head_params* p = &bp->fwd;
    const long long total = (long long)p->B * p->H * p->W;
    SF_CHECK_ARG(bp->g_w1 && bp->g_b1 && bp->g_w2 && bp->g_b2, "sf_head_bwd: weight gradient buffers are required");
    SF_CHECK_ARG(!p->training || (p->save_mean && p->save_invstd), "sf_head_bwd: training needs the saved batch statistics");
    Workspace ws(ws_ptr, ws_bytes);
    float2* t = ws.take<float2>((size_t)total);
    float* ga = ws.take<float>((size_t)total * 2);
    float* gz = ws.take<float>((size_t)total * 2);
    float* vec = ws.take<float>(16);
    if (!t || !vec) { set_error("sf_head_bwd: workspace too small (%zu B given)", ws_bytes); return SF_ERR_WORKSPACE; }
    long long nb = (total + HB_THREADS - 1) / HB_THREADS;
    if (nb > (long long)sm_count() * 8) nb = (long long)sm_count() * 8;
    const int blocks = (int)nb;
    ProfScope ps("bwd_final_head", 200.0 * (double)total, 60.0 * (double)total, st);
    k_hb_prepare<<<1, 32, 0, st>>>(p->bn_gamma, p->bn_beta, p->running_mean, p->running_var, p->save_mean, p->save_invstd, vec, p->bn_eps, p->training);
    SF_CHECK_LAUNCH("hb_prepare");
    k_hb_conv1<<<blocks, HB_THREADS, 0, st>>>(p->x, p->y, p->w1, p->b1, t, p->B, p->H, p->W, p->ksize);
    SF_CHECK_LAUNCH("hb_conv1");
    k_zero<<<blocks, HB_THREADS, 0, st>>>(ga, total * 2);
    SF_CHECK_LAUNCH("hb_zero");
    k_zero<<<blocks, HB_THREADS, 0, st>>>(bp->g_x, total);
    SF_CHECK_LAUNCH("hb_zero");
    k_zero<<<blocks, HB_THREADS, 0, st>>>(bp->g_y, total);
    SF_CHECK_LAUNCH("hb_zero");
    k_hb_conv2_bwd<<<blocks, HB_THREADS, 0, st>>>(bp->gout, t, vec, p->w2, ga, bp->g_w2, bp->g_b2, p->B, p->H, p->W, p->ksize);
    SF_CHECK_LAUNCH("hb_conv2_bwd");
    k_hb_bn_stats<<<blocks, HB_THREADS, 0, st>>>(ga, t, vec, vec + 4, vec + 6, gz, vec + 8, total);
    SF_CHECK_LAUNCH("hb_bn_stats");
    k_hb_bn_param_grads<<<1, 32, 0, st>>>(vec + 8, bp->g_bn_gamma, bp->g_bn_beta);
    SF_CHECK_LAUNCH("hb_bn_param_grads");
    k_hb_conv1_bwd<<<blocks, HB_THREADS, 0, st>>>(gz, t, vec + 8, p->bn_gamma, vec + 4, vec + 6, p->x, p->y, p->w1, bp->g_x, bp->g_y,
                                                  bp->g_w1, bp->g_b1, p->B, p->H, p->W, p->ksize, p->training);
    SF_CHECK_LAUNCH("hb_conv1_bwd");
    return SF_OK;
}

}  // namespace sf

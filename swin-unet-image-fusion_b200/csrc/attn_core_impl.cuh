// Attention core of the bf16 path (a001:317-354): per window and head
//     S = (Q K^T) d^-1/2 + bias ; masked -> -1e10 ; P = softmax(S) ; O = P V
// on bf16 Q/K/V produced by the projection GEMM, fp32 math, bf16 output.
//
// Why CUDA cores: with 49-token windows and head_dim 3..48 the two GEMMs are 2*49*49*d MACs per
// (window, head) while the softmax needs 49*49 exponentials -- for d <= 12 (93% of all windows of
// the 256x256 model) the MUFU/ALU work of the softmax dominates, not the MACs.  The kernel is
// therefore organised around the softmax: one thread per (query row, head), K/V of the window in
// shared memory as float4 so that a warp (32 query rows of one head) reads them as broadcasts,
// scores kept in registers between the max pass and the exp/PV pass, exp2 with log2(e) folded
// into q and the bias matrix.  Cyclic shift, window partition, head split, window reverse and
// un-shift are the index function win_token_src() (no materialised copies).
//
// CTAs are persistent (grid = SMs x resident CTAs) and loop over windows; the (T x T) bias
// matrix is gathered from the (2ws-1)^2 table once per CTA.
// (templates shared by attn_core.cu and attn_core_b.cu .. attn_core_e.cu: the fully unrolled per-shape kernels compile for
// minutes, so their instantiations are spread over five translation units that build in parallel)
#pragma once
#include "bf16_kernels.cuh"
#include "tc_common.cuh"

namespace sf {
using bf16 = __nv_bfloat16;

static __device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct AttnArgs {
    const bf16* qkv;     // [Mtok][ld]: q at column 0, k at koff, v at voff; 16-bit (bf16 or fp16) elements
    int qkv_fp16;
    long long ld;
    int koff, voff;
    bf16* O;             // row-major [Mtok][ldo]            (o_nkc == 0)
    long long ldo;       //   or UMMA-tiled: chunk (tile, kc, r) at ((tile*o_nkc + kc)*128 + r)*8 elements
    int o_nkc;
    const float* table;  // (2wsh-1, 2wsw-1) fp32
    WinGeom g;
    int nh, d, dp;       // dp = head_dim rounded up to a multiple of 4 (smem row of one head)
    long long nwin;
    float scale_log2e;   // d^-1/2 * log2(e)
};

struct AttnSmem {
    float* biasm;        // [T][TS]  (TS odd -> conflict-free when lanes differ in the query row)
    float* Ks;           // [T][nh][dp]
    float* Vs;           // [T][nh][dp]
    float* Qs;           // [T][inner+1]   (SMALL variant only)
    long long* rows;     // [T]
    int* regs;           // [T]
    bf16* Ob;            // [T][inner]     (SMALL variant only)
};

__host__ __device__ static inline int odd_up(int v) { return v | 1; }

template <bool SMALL>
__host__ __device__ static inline size_t attn_carve(AttnSmem* s, uint8_t* base, int T, int nh, int dp, int inner, bool stage_qo = SMALL) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~(size_t)15; return o; };
    size_t o_b = take((size_t)T * odd_up(T) * 4), o_k = take((size_t)T * nh * dp * 4), o_v = take((size_t)T * nh * dp * 4);
    size_t o_q = stage_qo ? take((size_t)T * (inner + 1) * 4) : 0;
    size_t o_r = take((size_t)T * 8), o_g = take((size_t)T * 4);
    size_t o_o = stage_qo ? take((size_t)T * inner * 2) : 0;
    if (s) {
        s->biasm = reinterpret_cast<float*>(base + o_b);
        s->Ks = reinterpret_cast<float*>(base + o_k);
        s->Vs = reinterpret_cast<float*>(base + o_v);
        s->Qs = reinterpret_cast<float*>(base + o_q);
        s->rows = reinterpret_cast<long long*>(base + o_r);
        s->regs = reinterpret_cast<int*>(base + o_g);
        s->Ob = reinterpret_cast<bf16*>(base + o_o);
    }
    return off;
}

__device__ __forceinline__ float ld16f(const bf16* p, int fp16) {
    return fp16 ? __half2float(*reinterpret_cast<const __half*>(p)) : __bfloat162float(*p);
}
__device__ __forceinline__ void unpack8(const uint4& raw, int fp16, float (&f)[8]) {
    if (fp16) {
        const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
        for (int e = 0; e < 4; e++) { float2 p = __half22float2(h2[e]); f[2 * e] = p.x; f[2 * e + 1] = p.y; }
    } else {
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int e = 0; e < 4; e++) { float2 p = __bfloat1622float2(h2[e]); f[2 * e] = p.x; f[2 * e + 1] = p.y; }
    }
}

__device__ __forceinline__ long long o_elem_offset(const AttnArgs& a, long long tok, int col) {
    if (a.o_nkc == 0) return tok * a.ldo + col;
    long long tile = tok >> 7;
    int r = (int)(tok & 127);
    return ((tile * a.o_nkc + (col >> 3)) * 128 + r) * 8 + (col & 7);
}

// -------------------------------------------------------------------------------------------------
// SMALL variant: window of exactly TT tokens, head_dim <= DP <= 16 (DP a multiple of 4).
// smem K/V layout [head][token][DP] fp32: for a fixed thread the address of key j is base + j*DP
// floats -- an immediate offset in the fully unrolled loops, no address arithmetic.
// -------------------------------------------------------------------------------------------------
template <int DP, int TT, bool MASKED>
__device__ __forceinline__ void attn_row_small(const float* __restrict__ brow, const float4* __restrict__ K4,
                                               const float4* __restrict__ V4, const int* __restrict__ regs, int qreg,
                                               const float (&q)[DP], float (&acc)[DP], float& sum) {
    float sc[TT];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < TT; j++) {
        float v = brow[j];
#pragma unroll
        for (int c4 = 0; c4 < DP / 4; c4++) {
            float4 k = K4[j * (DP / 4) + c4];
            v = fmaf(q[c4 * 4], k.x, v); v = fmaf(q[c4 * 4 + 1], k.y, v);
            v = fmaf(q[c4 * 4 + 2], k.z, v); v = fmaf(q[c4 * 4 + 3], k.w, v);
        }
        if (MASKED) { if (regs[j] != qreg) v = -1.4426950e10f; }
        sc[j] = v;
        mx = fmaxf(mx, v);
    }
#pragma unroll
    for (int j = 0; j < TT; j++) {
        float p = ex2(sc[j] - mx);
        sum += p;
#pragma unroll
        for (int c4 = 0; c4 < DP / 4; c4++) {
            float4 vv = V4[j * (DP / 4) + c4];
            acc[c4 * 4] = fmaf(p, vv.x, acc[c4 * 4]); acc[c4 * 4 + 1] = fmaf(p, vv.y, acc[c4 * 4 + 1]);
            acc[c4 * 4 + 2] = fmaf(p, vv.z, acc[c4 * 4 + 2]); acc[c4 * 4 + 3] = fmaf(p, vv.w, acc[c4 * 4 + 3]);
        }
    }
}

template <int DP, int TT>
__global__ void __launch_bounds__(DP <= 16 ? 416 : 256, DP <= 8 ? 2 : 1) k_attn_small(AttnArgs a) {
    constexpr bool STAGE = DP <= 16;   // Q and O go through smem (coalesced 16-byte global accesses)
    extern __shared__ __align__(16) uint8_t smraw[];
    AttnSmem s;
    const WinGeom& g = a.g;
    constexpr int T = TT, TS = TT | 1;
    const int nh = a.nh, d = a.d, inner = nh * d;
    attn_carve<true>(&s, smraw, T, nh, DP, inner, STAGE);
    const int tw = 2 * g.wsw - 1;
    constexpr float LOG2E = 1.4426950408889634f;
    for (int i = threadIdx.x; i < T * T; i += blockDim.x) {
        int qi = i / T, kj = i - qi * T;
        s.biasm[qi * TS + kj] = LOG2E * a.table[(kj / g.wsw - qi / g.wsw + g.wsh - 1) * tw + (kj % g.wsw - qi % g.wsw + g.wsw - 1)];
    }
    for (int i = threadIdx.x; i < T * nh * DP; i += blockDim.x) { s.Ks[i] = 0.f; s.Vs[i] = 0.f; }   // zero the DP > d padding once
    const bool vec_in = ((inner & 7) == 0) && ((a.ld & 7) == 0) && ((a.koff & 7) == 0) && ((a.voff & 7) == 0);
    const bool vec_out = STAGE && ((inner & 7) == 0) && (a.o_nkc != 0 || (a.ldo & 7) == 0);

    for (long long win = blockIdx.x; win < a.nwin; win += gridDim.x) {
        __syncthreads();  // previous window's smem fully consumed
        int flag = 0;
        for (int t = threadIdx.x; t < T; t += blockDim.x) {
            int rg;
            s.rows[t] = win_token_src(g, (int)win, t, &rg);
            s.regs[t] = rg;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < T; t += blockDim.x) flag |= (s.regs[t] != s.regs[0]);
        if (vec_in) {   // coalesced 16-byte chunks of each token row: q | k | v
            const int nch = inner >> 3, per_tok = (STAGE ? 3 : 2) * nch;
            for (int i = threadIdx.x; i < T * per_tok; i += blockDim.x) {
                int t = i / per_tok, c = i - t * per_tok;
                int which = c / nch, ch = c - which * nch;
                if (!STAGE) which += 1;   // only k and v are staged
                const int col = (which == 0 ? 0 : (which == 1 ? a.koff : a.voff)) + ch * 8;
                uint4 raw = *reinterpret_cast<const uint4*>(a.qkv + s.rows[t] * a.ld + col);
                float f[8];
                unpack8(raw, a.qkv_fp16, f);
                if (which == 0) {
#pragma unroll
                    for (int e = 0; e < 8; e++) s.Qs[t * (inner + 1) + ch * 8 + e] = f[e];
                } else {
                    float* dst = which == 1 ? s.Ks : s.Vs;
                    int h = (ch * 8) / d, dd = ch * 8 - h * d;
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        dst[(h * T + t) * DP + dd] = f[e];
                        if (++dd == d) { dd = 0; h++; }
                    }
                }
            }
        } else {
            for (int i = threadIdx.x; i < T * inner; i += blockDim.x) {
                int t = i / inner, cc = i - t * inner, h = cc / d, dd = cc - h * d;
                const bf16* src = a.qkv + s.rows[t] * a.ld + cc;
                s.Ks[(h * T + t) * DP + dd] = ld16f(src + a.koff, a.qkv_fp16);
                s.Vs[(h * T + t) * DP + dd] = ld16f(src + a.voff, a.qkv_fp16);
                if (STAGE) s.Qs[t * (inner + 1) + cc] = ld16f(src, a.qkv_fp16);
            }
        }
        const bool has_mask = __syncthreads_or(flag) != 0;

        for (int item = threadIdx.x; item < nh * T; item += blockDim.x) {   // one thread per (head, query row)
            const int head = item / T, qi = item - head * T;
            float q[DP], acc[DP];
#pragma unroll
            for (int dd = 0; dd < DP; dd++) {
                acc[dd] = 0.f;
                float qv = 0.f;
                if (dd < d) qv = STAGE ? s.Qs[qi * (inner + 1) + head * d + dd] : ld16f(a.qkv + s.rows[qi] * a.ld + head * d + dd, a.qkv_fp16);
                q[dd] = qv * a.scale_log2e;
            }
            const float4* K4 = reinterpret_cast<const float4*>(s.Ks + (size_t)head * T * DP);
            const float4* V4 = reinterpret_cast<const float4*>(s.Vs + (size_t)head * T * DP);
            float sum = 0.f;
            if (has_mask) attn_row_small<DP, TT, true>(s.biasm + qi * TS, K4, V4, s.regs, s.regs[qi], q, acc, sum);
            else attn_row_small<DP, TT, false>(s.biasm + qi * TS, K4, V4, s.regs, 0, q, acc, sum);
            const float inv = 1.f / sum;
            if (vec_out) {
#pragma unroll
                for (int dd = 0; dd < DP; dd++)
                    if (dd < d) s.Ob[qi * inner + head * d + dd] = __float2bfloat16_rn(acc[dd] * inv);
            } else {
#pragma unroll
                for (int dd = 0; dd < DP; dd++)
                    if (dd < d) a.O[o_elem_offset(a, s.rows[qi], head * d + dd)] = __float2bfloat16_rn(acc[dd] * inv);
            }
        }
        if (a.o_nkc * 8 > inner) {  // K padding of the tiled O operand must be exact zeros for the next GEMM
            const int npad = a.o_nkc * 8 - inner;
            for (int i = threadIdx.x; i < T * npad; i += blockDim.x) {
                int t = i / npad, c = inner + (i - t * npad);
                a.O[o_elem_offset(a, s.rows[t], c)] = __float2bfloat16_rn(0.f);
            }
        }
        if (vec_out) {
            __syncthreads();
            const int nch = inner >> 3;
            for (int i = threadIdx.x; i < T * nch; i += blockDim.x) {
                int t = i / nch, ch = i - t * nch;
                *reinterpret_cast<uint4*>(a.O + o_elem_offset(a, s.rows[t], ch * 8)) = *reinterpret_cast<const uint4*>(s.Ob + t * inner + ch * 8);
            }
        }
    }
}

template <int DP, int TT>
static int launch_attn_small(const AttnArgs& a, cudaStream_t st) {
    const int inner = a.nh * a.d;
    size_t smem = attn_carve<true>(nullptr, nullptr, TT, a.nh, DP, inner, DP <= 16);
    SF_CHECK_ARG(smem <= 227 * 1024, "attention core: %d heads x %d dims need %zu B of shared memory", a.nh, a.d, smem);
    static DeviceOnce configured;
    if (smem > 48 * 1024 && configured.need()) {
        cudaError_t e = cudaFuncSetAttribute(k_attn_small<DP, TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { set_error("attention core: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SF_ERR_CUDA; }
        configured.done();
    }
    int threads = (a.nh * TT + 31) / 32 * 32;
    const int maxthr = DP <= 16 ? 416 : 256;
    if (threads > maxthr) threads = maxthr;
    if (threads < 64) threads = 64;
    int per_sm = (int)(227 * 1024 / (smem + 1024));
    if (per_sm > (DP <= 8 ? 2 : 1)) per_sm = (DP <= 8 ? 2 : 1);
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)sm_count() * per_sm;
    if (grid > a.nwin) grid = a.nwin;
    const double mtok = (double)a.nwin * TT;
    ProfScope ps(prof_name("attn_core_small_c%d", inner), 4.0 * TT * mtok * inner, 8.0 * mtok * inner, st);
    k_attn_small<DP, TT><<<(unsigned)grid, threads, smem, st>>>(a);
    SF_CHECK_LAUNCH("attn_core_small");
    return SF_OK;
}

// -------------------------------------------------------------------------------------------------
// generic variant (any window size, head_dim <= 64): two passes, scores recomputed
// -------------------------------------------------------------------------------------------------
template <int DMAX, int TT, bool SMALL>
__global__ void __launch_bounds__(SMALL ? 448 : 256) k_attn_core(AttnArgs a) {
    extern __shared__ __align__(16) uint8_t smraw[];
    AttnSmem s;
    const WinGeom& g = a.g;
    const int T = g.T, nh = a.nh, d = a.d, dp = a.dp, inner = nh * d, TS = odd_up(T);
    attn_carve<SMALL>(&s, smraw, T, nh, dp, inner);
    const int tw = 2 * g.wsw - 1;
    constexpr float LOG2E = 1.4426950408889634f;

    for (int i = threadIdx.x; i < T * T; i += blockDim.x) {
        int qi = i / T, kj = i - qi * T;
        s.biasm[qi * TS + kj] = LOG2E * a.table[(kj / g.wsw - qi / g.wsw + g.wsh - 1) * tw + (kj % g.wsw - qi % g.wsw + g.wsw - 1)];
    }
    // zero the padding lanes of K/V once (dp > d)
    for (int i = threadIdx.x; i < T * nh * dp; i += blockDim.x) { s.Ks[i] = 0.f; s.Vs[i] = 0.f; }

    const bool vec_in = ((inner & 7) == 0) && ((a.ld & 7) == 0) && ((a.koff & 7) == 0) && ((a.voff & 7) == 0);
    const bool vec_out = SMALL && ((inner & 7) == 0) && (a.o_nkc != 0 || (a.ldo & 7) == 0);

    for (long long win = blockIdx.x; win < a.nwin; win += gridDim.x) {
        __syncthreads();  // previous window's smem fully consumed
        int flag = 0;
        for (int t = threadIdx.x; t < T; t += blockDim.x) {
            int rg;
            s.rows[t] = win_token_src(g, (int)win, t, &rg);
            s.regs[t] = rg;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < T; t += blockDim.x) flag |= (s.regs[t] != s.regs[0]);
        // ---- stage K, V (and Q) of the window: coalesced 16-byte chunks of each token row ---------
        if (vec_in) {
            const int nch = inner >> 3;
            const int per_tok = (SMALL ? 3 : 2) * nch;
            for (int i = threadIdx.x; i < T * per_tok; i += blockDim.x) {
                int t = i / per_tok, c = i - t * per_tok;
                int which = c / nch, ch = c - which * nch;  // SMALL: 0=q 1=k 2=v ; else 0=k 1=v
                if (!SMALL) which += 1;
                const int col = (which == 0 ? 0 : (which == 1 ? a.koff : a.voff)) + ch * 8;
                uint4 raw = *reinterpret_cast<const uint4*>(a.qkv + s.rows[t] * a.ld + col);
                float f[8];
                unpack8(raw, a.qkv_fp16, f);
                if (which == 0) {
#pragma unroll
                    for (int e = 0; e < 8; e++) s.Qs[t * (inner + 1) + ch * 8 + e] = f[e];
                } else {
                    float* dst = which == 1 ? s.Ks : s.Vs;
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        int cc = ch * 8 + e, h = cc / d, dd = cc - h * d;
                        dst[(t * nh + h) * dp + dd] = f[e];
                    }
                }
            }
        } else {
            for (int i = threadIdx.x; i < T * inner; i += blockDim.x) {
                int t = i / inner, cc = i - t * inner, h = cc / d, dd = cc - h * d;
                const bf16* src = a.qkv + s.rows[t] * a.ld + cc;
                s.Ks[(t * nh + h) * dp + dd] = ld16f(src + a.koff, a.qkv_fp16);
                s.Vs[(t * nh + h) * dp + dd] = ld16f(src + a.voff, a.qkv_fp16);
                if (SMALL) s.Qs[t * (inner + 1) + cc] = ld16f(src, a.qkv_fp16);
            }
        }
        const bool has_mask = __syncthreads_or(flag) != 0;

        // ---- one thread per (head, query row) --------------------------------------------------------
        for (int item = threadIdx.x; item < nh * T; item += blockDim.x) {
            const int head = item / T, qi = item - head * T;
            float q[DMAX], acc[DMAX];
#pragma unroll
            for (int dd = 0; dd < DMAX; dd++) {
                acc[dd] = 0.f;
                float qv = 0.f;
                if (dd < d) qv = SMALL ? s.Qs[qi * (inner + 1) + head * d + dd] : ld16f(a.qkv + s.rows[qi] * a.ld + head * d + dd, a.qkv_fp16);
                q[dd] = qv * a.scale_log2e;
            }
            const int qreg = s.regs[qi];
            const float* brow = s.biasm + qi * TS;
            const float4* K4 = reinterpret_cast<const float4*>(s.Ks) + head * (dp >> 2);
            const float4* V4 = reinterpret_cast<const float4*>(s.Vs) + head * (dp >> 2);
            const int jstride = nh * (dp >> 2);
            float mx = -INFINITY, sum = 0.f;
            if (SMALL) {
                float sc[TT];
#pragma unroll
                for (int j = 0; j < TT; j++) {
                    if (j < T) {
                        float v = brow[j];
#pragma unroll
                        for (int c4 = 0; c4 < DMAX / 4; c4++) {
                            if (c4 * 4 < dp) {
                                float4 k = K4[j * jstride + c4];
                                v = fmaf(q[c4 * 4], k.x, v); v = fmaf(q[c4 * 4 + 1], k.y, v);
                                v = fmaf(q[c4 * 4 + 2], k.z, v); v = fmaf(q[c4 * 4 + 3], k.w, v);
                            }
                        }
                        if (has_mask && s.regs[j] != qreg) v = -1.4426950e10f;
                        sc[j] = v;
                        mx = fmaxf(mx, v);
                    }
                }
#pragma unroll
                for (int j = 0; j < TT; j++) {
                    if (j < T) {
                        float p = ex2(sc[j] - mx);
                        sum += p;
#pragma unroll
                        for (int c4 = 0; c4 < DMAX / 4; c4++) {
                            if (c4 * 4 < dp) {
                                float4 vv = V4[j * jstride + c4];
                                acc[c4 * 4] = fmaf(p, vv.x, acc[c4 * 4]); acc[c4 * 4 + 1] = fmaf(p, vv.y, acc[c4 * 4 + 1]);
                                acc[c4 * 4 + 2] = fmaf(p, vv.z, acc[c4 * 4 + 2]); acc[c4 * 4 + 3] = fmaf(p, vv.w, acc[c4 * 4 + 3]);
                            }
                        }
                    }
                }
            } else {
                for (int pass = 0; pass < 2; pass++) {
                    for (int j = 0; j < T; j++) {
                        float v = brow[j];
#pragma unroll
                        for (int c4 = 0; c4 < DMAX / 4; c4++) {
                            if (c4 * 4 < dp) {
                                float4 k = K4[j * jstride + c4];
                                v = fmaf(q[c4 * 4], k.x, v); v = fmaf(q[c4 * 4 + 1], k.y, v);
                                v = fmaf(q[c4 * 4 + 2], k.z, v); v = fmaf(q[c4 * 4 + 3], k.w, v);
                            }
                        }
                        if (has_mask && s.regs[j] != qreg) v = -1.4426950e10f;
                        if (pass == 0) {
                            mx = fmaxf(mx, v);
                        } else {
                            float p = ex2(v - mx);
                            sum += p;
#pragma unroll
                            for (int c4 = 0; c4 < DMAX / 4; c4++) {
                                if (c4 * 4 < dp) {
                                    float4 vv = V4[j * jstride + c4];
                                    acc[c4 * 4] = fmaf(p, vv.x, acc[c4 * 4]); acc[c4 * 4 + 1] = fmaf(p, vv.y, acc[c4 * 4 + 1]);
                                    acc[c4 * 4 + 2] = fmaf(p, vv.z, acc[c4 * 4 + 2]); acc[c4 * 4 + 3] = fmaf(p, vv.w, acc[c4 * 4 + 3]);
                                }
                            }
                        }
                    }
                }
            }
            const float inv = 1.f / sum;
            if (vec_out) {
#pragma unroll
                for (int dd = 0; dd < DMAX; dd++)
                    if (dd < d) s.Ob[qi * inner + head * d + dd] = __float2bfloat16_rn(acc[dd] * inv);
            } else {
#pragma unroll
                for (int dd = 0; dd < DMAX; dd++)
                    if (dd < d) a.O[o_elem_offset(a, s.rows[qi], head * d + dd)] = __float2bfloat16_rn(acc[dd] * inv);
            }
        }
        if (a.o_nkc * 8 > inner) {  // K padding of the tiled O operand must be exact zeros for the next GEMM
            const int npad = a.o_nkc * 8 - inner;
            for (int i = threadIdx.x; i < T * npad; i += blockDim.x) {
                int t = i / npad, c = inner + (i - t * npad);
                a.O[o_elem_offset(a, s.rows[t], c)] = __float2bfloat16_rn(0.f);
            }
        }
        if (vec_out) {
            __syncthreads();
            const int nch = inner >> 3;
            for (int i = threadIdx.x; i < T * nch; i += blockDim.x) {
                int t = i / nch, ch = i - t * nch;
                *reinterpret_cast<uint4*>(a.O + o_elem_offset(a, s.rows[t], ch * 8)) = *reinterpret_cast<const uint4*>(s.Ob + t * inner + ch * 8);
            }
        }
    }
}

template <int DMAX, int TT, bool SMALL>
static int launch_attn_t(const AttnArgs& a, cudaStream_t st) {
    const int inner = a.nh * a.d;
    size_t smem = attn_carve<SMALL>(nullptr, nullptr, a.g.T, a.nh, a.dp, inner);
    SF_CHECK_ARG(smem <= 227 * 1024, "attention core: window of %d tokens x %d channels needs %zu B of shared memory", a.g.T, inner, smem);
    static DeviceOnce configured;
    if (smem > 48 * 1024 && configured.need()) {
        cudaError_t e = cudaFuncSetAttribute(k_attn_core<DMAX, TT, SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { set_error("attention core: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SF_ERR_CUDA; }
        configured.done();
    }
    const int items = a.nh * a.g.T;
    const int maxthr = SMALL ? 448 : 256;
    int threads = (items + 31) / 32 * 32;
    if (threads > maxthr) threads = maxthr;
    if (threads < 64) threads = 64;
    int per_sm = (int)(227 * 1024 / (smem + 1024));
    if (per_sm > 2048 / threads) per_sm = 2048 / threads;
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)sm_count() * per_sm;
    if (grid > a.nwin) grid = a.nwin;
    const double mtok = (double)a.nwin * a.g.T;
    ProfScope ps(prof_name("attn_core_generic_c%d", inner), 4.0 * a.g.T * mtok * inner, 8.0 * mtok * inner, st);
    k_attn_core<DMAX, TT, SMALL><<<(unsigned)grid, threads, smem, st>>>(a);
    SF_CHECK_LAUNCH("attn_core");
    return SF_OK;
}

// one dispatcher per translation unit: returns SF_ERR_UNSUPPORTED when the shape belongs to another unit
int attn_core_dispatch_a(const AttnArgs& a, cudaStream_t st);   // 7x7 windows, head_dim <= 16
int attn_core_dispatch_b(const AttnArgs& a, cudaStream_t st);   // 7x7 windows, head_dim 17..64 (forwards to d / e above 32)
int attn_core_dispatch_c(const AttnArgs& a, cudaStream_t st);   // 8x8 windows and the general kernel
int attn_core_dispatch_d(const AttnArgs& a, cudaStream_t st);   // 7x7 windows, head_dim 33..48
int attn_core_dispatch_e(const AttnArgs& a, cudaStream_t st);   // 7x7 windows, head_dim 49..64

}  // namespace sf

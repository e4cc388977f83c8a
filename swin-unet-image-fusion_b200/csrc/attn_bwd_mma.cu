// Attention-core backward on tensor cores for 7x7 windows and the head dims of the default model (d = 3, 6, 12, 24, 48;
// the head dimension is processed in chunks of 16).  Adjoint of a001:317-354:
//   P = softmax(scale Q K^T + bias -> mask) ; dV = P^T dO ; dP = dO V^T ; dS = P o (dP - rowsum(P o dP)) ;
//   dtable[idx(i,j)] += dS ; dQ = scale dS K ; dK = scale dS^T Q.
// One CTA of four warps owns one (window, head) at a time; warp m owns the 16-query tile m.  The q/k/v/dO rows (gathered through the shift + window-partition
// index math) are converted to fp16 in shared memory, row-major [token][16] and transposed [dim][token]; the
// five products run on mma.sync.m16n8k16 (fp16 operands = the 10-bit mantissa of the TF32 GEMMs around this kernel,
// fp32 accumulation):
//   per 16-query tile:  S = Q K^T, dP = dO V^T (accumulator fragments) -> softmax / dS in registers (quad shuffles)
//                       dQ = dS K            (dS re-used in place as the A fragment)
//                       dK^T += Q^T dS, dV^T += dO^T P   (B fragments = 8x8 register transposes of dS / P: movmatrix)
// dO is normalised per (window, head) by an exact power of two (max |dO| -> [1,2)) so that tiny loss gradients keep
// their full fp16 mantissa; every output is multiplied back.  dK^T / dV^T partials of the four warps are summed through
// shared memory.  A lane meets the same 28 (query, key) positions on every item, so the 13x13 table gradient is
// accumulated in registers and leaves the CTA once (shared-memory reduction, one global atomicAdd per entry and CTA).
#include <cuda_fp16.h>
#include "bwd_kernels.cuh"

namespace sf {
namespace {

constexpr int T = 49, TW = 13;
constexpr int RSW = 8;      // row-major operand row = 16 halves = 8 words
constexpr int TSW = 36;     // transposed operand row = 72 halves = 36 words (64 tokens + pad: (4 g + t) % 32 distinct banks)
constexpr int WARPS = 4;
constexpr float LOG2E = 1.4426950408889634f;

// the head dimension is handled in KC chunks of 16 (d = 24 -> 2, d = 48 -> 3): every operand array exists per chunk
template <int KC>
struct __align__(16) ItemSmem {
    uint32_t q[KC][64 * RSW], k[KC][64 * RSW], v[KC][64 * RSW], g[KC][64 * RSW];
    uint32_t qt[KC][16 * TSW], kt[KC][16 * TSW], gt[KC][16 * TSW], vt[KC][16 * TSW];
    float4 red[WARPS][14][32];      // per-warp partial dK^T / dV^T accumulator fragments
    uint32_t tok2[2][64];           // token index (row of the un-shifted map) of every token of the window, same double buffering
    uint32_t rows2[2][64];          // element offset (token row * inner, < 2^32: checked by the launcher) of every token of the
                                    // window; double-buffered by item parity: the tail of item i reads them while item i+1 is set up
    int reg2[2][64];
    float tab[176], gtab[176];
    float wmax[WARPS];
};
constexpr int OPERAND_WORDS = 4 * 64 * RSW + 4 * 16 * TSW;   // per chunk

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// 8x8 b16 transpose across the warp: in: lane (gq, tq) holds M[gq][2tq..2tq+1]; out: M[2tq..2tq+1][gq]
__device__ __forceinline__ uint32_t movm_trans(uint32_t x) {
    uint32_t y;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// element offset inside a bf16 tensor in the UMMA-tiled layout [tile of 128 rows][8-column chunk][row][8]
__device__ __forceinline__ uint32_t tiled_off(uint32_t t, uint32_t col, uint32_t nkc) {
    return ((t >> 7) * nkc + (col >> 3)) * 1024u + (t & 127u) * 8u + (col & 7u);
}
__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// gather one operand: element e = token * D + dd (flat over the 49 x D block), threads take e = tid + 128 i
template <int D, int NE>
__device__ __forceinline__ void load_flat(const float* __restrict__ src, const uint32_t* rows, uint32_t hoff, float (&r)[NE]) {
#pragma unroll
    for (int i = 0; i < NE; i++) {
        const int e = threadIdx.x + WARPS * 32 * i;
        const int tk = e / D, dd = e - tk * D;
        r[i] = e < T * D ? __ldg(src + (rows[tk] + hoff + (uint32_t)dd)) : 0.f;
    }
}
template <int D, int NE>
__device__ __forceinline__ void store_flat(const float (&r)[NE], __half* rowmajor, __half* transposed, float mul) {
#pragma unroll
    for (int i = 0; i < NE; i++) {
        const int e = threadIdx.x + WARPS * 32 * i;
        const int tk = e / D, dd = e - tk * D;
        if (e < T * D) {
            const __half h = __float2half_rn(r[i] * mul);
            const int c = dd >> 4, dl = dd & 15;           // chunk, dim inside the chunk
            rowmajor[c * (64 * 2 * RSW) + tk * 2 * RSW + dl] = h;
            transposed[c * (16 * 2 * TSW) + dl * 2 * TSW + tk] = h;
        }
    }
}

// One CTA (4 warps) per (window, head); warp m owns query rows 16 m .. 16 m + 15, so every lane meets the same 28
// (query, key) positions on every item and the bias-table gradient is accumulated in 28 registers.
template <int D>
__global__ void __launch_bounds__(WARPS * 32, D <= 16 ? 3 : 2) k_attn_bwd_mma(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V,
                                                             const float* __restrict__ gO, float* __restrict__ dQ, float* __restrict__ dK,
                                                             float* __restrict__ dV, float* __restrict__ O, const float* __restrict__ table,
                                                             float* __restrict__ gtable, WinGeom g, int inner, int nh, float scale, long long nitems,
                                                             AttnBwdTiledOut to) {
    constexpr int NE = (T * D + WARPS * 32 - 1) / (WARPS * 32);
    constexpr int KC = (D + 15) / 16;     // 16-wide chunks of the head dimension
    extern __shared__ __align__(16) unsigned char smraw[];
    ItemSmem<KC>* ws = reinterpret_cast<ItemSmem<KC>*>(smraw);
    float2* red2 = reinterpret_cast<float2*>(&ws->red[0][0][0]);   // narrow heads (D <= 8) exchange float2 partials: [warp * 14 + fragment][lane]
    const int m = threadIdx.x >> 5, lane = threadIdx.x & 31;
    {
        uint32_t* ops = reinterpret_cast<uint32_t*>(ws);                 // q, k, v, g, qt, kt, gt, vt are contiguous
        for (int i = threadIdx.x; i < KC * OPERAND_WORDS; i += WARPS * 32) ops[i] = 0u;
    }
    ws->reg2[0][threadIdx.x & 63] = 0;
    ws->reg2[1][threadIdx.x & 63] = 0;
    for (int i = threadIdx.x; i < 169; i += WARPS * 32) { ws->tab[i] = table[i] * LOG2E; ws->gtab[i] = 0.f; }
    const float scale_l2 = scale * LOG2E;
    const int gq = lane >> 2, tq = lane & 3;
    const int r0 = 16 * m + gq, r1 = r0 + 8;
    // rows >= 49 are computed on clamped indices and zeroed
    const int rc0 = min(r0, T - 1), rc1 = min(r1, T - 1);
    const int qo0 = (6 - rc0 / 7) * TW + 6 - rc0 % 7, qo1 = (6 - rc1 / 7) * TW + 6 - rc1 % 7;
    const float rv0 = r0 < T ? 1.f : 0.f, rv1 = r1 < T ? 1.f : 0.f;
    // column part of the bias-table index: key j = 8 n + 2 tq + e -> (j / 7) * 13 + j % 7  (a001:100-144, key - query)
    int cj[7][2];
#pragma unroll
    for (int n = 0; n < 7; n++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int j = min(8 * n + 2 * tq + e, T - 1);
            cj[n][e] = (j / 7) * TW + j % 7;
        }
    float tl[7][4];     // table gradient of this lane's 28 (query, key) positions, summed over items
#pragma unroll
    for (int n = 0; n < 7; n++)
#pragma unroll
        for (int c = 0; c < 4; c++) tl[n][c] = 0.f;

    // Software pipeline: the operand elements of the NEXT item are gathered into registers (pa..pd) while the current
    // item is computed.  rows2 / reg2 are double-buffered by item parity.
    float pa[NE], pb[NE], pc[NE], pd[NE];
    if (blockIdx.x < nitems) {
        const int win = (int)(blockIdx.x / nh), head = (int)(blockIdx.x - (long long)win * nh);
        if (threadIdx.x < T) {
            int rg;
            const long long tk = win_token_src(g, win, threadIdx.x, &rg);
            ws->tok2[0][threadIdx.x] = (uint32_t)tk;
            ws->rows2[0][threadIdx.x] = (uint32_t)(tk * inner);
            ws->reg2[0][threadIdx.x] = rg;
        }
        __syncthreads();
        load_flat<D, NE>(gO, ws->rows2[0], (uint32_t)(head * D), pd);
        load_flat<D, NE>(Q, ws->rows2[0], (uint32_t)(head * D), pa);
        load_flat<D, NE>(K, ws->rows2[0], (uint32_t)(head * D), pb);
        load_flat<D, NE>(V, ws->rows2[0], (uint32_t)(head * D), pc);
    }
    int parity = 0;
    for (long long item = blockIdx.x; item < nitems; item += gridDim.x, parity ^= 1) {
        const int win = (int)(item / nh), head = (int)(item - (long long)win * nh);
        const int hoff = head * D;
        const uint32_t* rows = ws->rows2[parity];
        const uint32_t* tok = ws->tok2[parity];
        const int* reg = ws->reg2[parity];
        const long long nitem = item + gridDim.x;
        const int nwin = (int)(nitem / nh), nhead = (int)(nitem - (long long)nwin * nh);
        float inv_sc;
        {
            // every warp left the previous item's operands behind at its last barrier (the tail only reads `red` and rows)
            float mx = 0.f;
#pragma unroll
            for (int i = 0; i < NE; i++) mx = fmaxf(mx, fabsf(pd[i]));
#pragma unroll
            for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if (lane == 0) ws->wmax[m] = mx;
            store_flat<D, NE>(pa, reinterpret_cast<__half*>(ws->q[0]), reinterpret_cast<__half*>(ws->qt[0]), 1.f);
            store_flat<D, NE>(pb, reinterpret_cast<__half*>(ws->k[0]), reinterpret_cast<__half*>(ws->kt[0]), 1.f);
            store_flat<D, NE>(pc, reinterpret_cast<__half*>(ws->v[0]), reinterpret_cast<__half*>(ws->vt[0]), 1.f);
            if (nitem < nitems && threadIdx.x < T) {     // token rows of the next item (other parity: the previous item's tail is over)
                int rg;
                const long long tk = win_token_src(g, nwin, threadIdx.x, &rg);
                ws->tok2[parity ^ 1][threadIdx.x] = (uint32_t)tk;
                ws->rows2[parity ^ 1][threadIdx.x] = (uint32_t)(tk * inner);
                ws->reg2[parity ^ 1][threadIdx.x] = rg;
            }
            __syncthreads();
            mx = fmaxf(fmaxf(ws->wmax[0], ws->wmax[1]), fmaxf(ws->wmax[2], ws->wmax[3]));
            const int ex = (__float_as_int(mx) >> 23) & 0xff;        // mx = 1.f * 2^(ex-127)
            const float sc = __int_as_float((254 - ex) << 23);        // 2^(127-ex): max |dO| * sc in [1, 2)
            inv_sc = __int_as_float(ex << 23);                        // (mx == 0: every output is 0 * 0)
            store_flat<D, NE>(pd, reinterpret_cast<__half*>(ws->g[0]), reinterpret_cast<__half*>(ws->gt[0]), sc);
            if (nitem < nitems) {
                load_flat<D, NE>(gO, ws->rows2[parity ^ 1], (uint32_t)(nhead * D), pd);
                load_flat<D, NE>(Q, ws->rows2[parity ^ 1], (uint32_t)(nhead * D), pa);
                load_flat<D, NE>(K, ws->rows2[parity ^ 1], (uint32_t)(nhead * D), pb);
                load_flat<D, NE>(V, ws->rows2[parity ^ 1], (uint32_t)(nhead * D), pc);
            }
        }
        __syncthreads();
        const bool has_mask = reg[0] != reg[T - 1];   // region ids grow along both axes of a window
        // packed region ids of this lane's 14 key columns (4 bits each), only on windows the shift mask touches
        uint32_t creg0 = 0, creg1 = 0;
        if (has_mask) {
#pragma unroll
            for (int n = 0; n < 7; n++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const uint32_t r = (uint32_t)reg[8 * n + 2 * tq + e];
                    if (n < 4) creg0 |= r << (4 * (2 * n + e)); else creg1 |= r << (4 * (2 * (n - 4) + e));
                }
        }
        float s[7][4], dp[7][4];
#pragma unroll
        for (int n = 0; n < 7; n++)
#pragma unroll
            for (int c = 0; c < 4; c++) { s[n][c] = 0.f; dp[n][c] = 0.f; }
#pragma unroll
        for (int kc = 0; kc < KC; kc++) {
            const uint32_t* qh = ws->q[kc];
            const uint32_t* gh = ws->g[kc];
            const uint32_t qa0 = qh[r0 * RSW + tq], qa1 = qh[r1 * RSW + tq], qa2 = qh[r0 * RSW + tq + 4], qa3 = qh[r1 * RSW + tq + 4];
            const uint32_t ga0 = gh[r0 * RSW + tq], ga1 = gh[r1 * RSW + tq], ga2 = gh[r0 * RSW + tq + 4], ga3 = gh[r1 * RSW + tq + 4];
#pragma unroll
            for (int n = 0; n < 7; n++) {
                const int kr = (8 * n + gq) * RSW + tq;
                mma16816(s[n], qa0, qa1, qa2, qa3, ws->k[kc][kr], ws->k[kc][kr + 4]);
                mma16816(dp[n], ga0, ga1, ga2, ga3, ws->v[kc][kr], ws->v[kc][kr + 4]);
            }
        }
        // ---- softmax rows r0 (c = 0,1) and r1 (c = 2,3)
        const uint32_t rr0 = (uint32_t)reg[rc0], rr1 = (uint32_t)reg[rc1];
        // logits in the log2 domain: (raw * scale + bias) * log2(e), bias table pre-multiplied in shared memory
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int n = 0; n < 7; n++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                s[n][e] = fmaf(s[n][e], scale_l2, ws->tab[qo0 + cj[n][e]]);
                s[n][2 + e] = fmaf(s[n][2 + e], scale_l2, ws->tab[qo1 + cj[n][e]]);
            }
        if (has_mask) {      // CTA-uniform: only windows on the shifted frame's seams carry more than one region
#pragma unroll
            for (int n = 0; n < 7; n++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const uint32_t cr = ((n < 4 ? creg0 >> (4 * (2 * n + e)) : creg1 >> (4 * (2 * (n - 4) + e))) & 15u);
                    if (cr != rr0) s[n][e] = -1e10f;
                    if (cr != rr1) s[n][2 + e] = -1e10f;
                }
        }
        // key columns 49..55 exist only in tile 6 (its lanes with 2 tq + e >= 1)
        if (tq != 0) { s[6][0] = -INFINITY; s[6][2] = -INFINITY; }
        s[6][1] = -INFINITY; s[6][3] = -INFINITY;
#pragma unroll
        for (int n = 0; n < 7; n++)
#pragma unroll
            for (int e = 0; e < 2; e++) { mx0 = fmaxf(mx0, s[n][e]); mx1 = fmaxf(mx1, s[n][2 + e]); }
        mx0 = quad_max(mx0); mx1 = quad_max(mx1);
        float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
        for (int n = 0; n < 7; n++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const float e0 = ex2_approx(s[n][e] - mx0), e1 = ex2_approx(s[n][2 + e] - mx1);
                s[n][e] = e0; s[n][2 + e] = e1;
                sum0 += e0; sum1 += e1;
            }
        const float inv0 = rv0 / quad_sum(sum0), inv1 = rv1 / quad_sum(sum1);
        float dl0 = 0.f, dl1 = 0.f;
#pragma unroll
        for (int n = 0; n < 7; n++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                s[n][e] *= inv0; s[n][2 + e] *= inv1;
                dl0 = fmaf(s[n][e], dp[n][e], dl0); dl1 = fmaf(s[n][2 + e], dp[n][2 + e], dl1);
            }
        dl0 = quad_sum(dl0); dl1 = quad_sum(dl1);
        uint32_t ph[7][2], dsh[7][2];
#pragma unroll
        for (int n = 0; n < 7; n++) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const float d0 = s[n][e] * (dp[n][e] - dl0), d1 = s[n][2 + e] * (dp[n][2 + e] - dl1);
                dp[n][e] = d0; dp[n][2 + e] = d1;
                tl[n][e] = fmaf(d0, inv_sc, tl[n][e]);
                tl[n][2 + e] = fmaf(d1, inv_sc, tl[n][2 + e]);
            }
            ph[n][0] = pack_h2(s[n][0], s[n][1]); ph[n][1] = pack_h2(s[n][2], s[n][3]);
            dsh[n][0] = pack_h2(dp[n][0], dp[n][1]); dsh[n][1] = pack_h2(dp[n][2], dp[n][3]);
        }
        // ---- O rows of this tile = P V (the forward output the projection's weight gradient needs; the forward saves nothing)
        //      and dQ rows = scale * dS K; k index = key: tiles 2kk, 2kk+1 of P / dS form one 16-deep slice
#pragma unroll
        for (int kc = 0; kc < KC; kc++) {
            constexpr int NDT_MAX = 2;
            const int ndt = (D - 16 * kc >= 9) ? 2 : 1;        // 8-wide dim tiles present in this chunk (compile time after unrolling)
            float o[NDT_MAX][4], dq[NDT_MAX][4];
#pragma unroll
            for (int nd = 0; nd < NDT_MAX; nd++)
#pragma unroll
                for (int c = 0; c < 4; c++) { o[nd][c] = 0.f; dq[nd][c] = 0.f; }
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
                const uint32_t p0 = ph[2 * kk][0], p1 = ph[2 * kk][1];
                const uint32_t p2 = kk < 3 ? ph[(2 * kk + 1) % 7][0] : 0u, p3 = kk < 3 ? ph[(2 * kk + 1) % 7][1] : 0u;
                const uint32_t a0 = dsh[2 * kk][0], a1 = dsh[2 * kk][1];
                const uint32_t a2 = kk < 3 ? dsh[(2 * kk + 1) % 7][0] : 0u, a3 = kk < 3 ? dsh[(2 * kk + 1) % 7][1] : 0u;
#pragma unroll
                for (int nd = 0; nd < NDT_MAX; nd++) {
                    if (nd >= ndt) continue;
                    const int kr = (8 * nd + gq) * TSW + 8 * kk + tq;
                    if (O || to.o) mma16816(o[nd], p0, p1, p2, p3, ws->vt[kc][kr], ws->vt[kc][kr + 4]);
                    mma16816(dq[nd], a0, a1, a2, a3, ws->kt[kc][kr], ws->kt[kc][kr + 4]);
                }
            }
            const float f = scale * inv_sc;
#pragma unroll
            for (int nd = 0; nd < NDT_MAX; nd++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int dd = 16 * kc + 8 * nd + 2 * tq + e;
                    if (nd < ndt && dd < D) {
                        if (to.dqkv) {   // bf16, UMMA-tiled: the operands the projection GEMMs of the backward bulk-copy
                            if (r0 < T) {
                                to.dqkv[tiled_off(tok[r0], (uint32_t)(hoff + dd), to.nkc3)] = __float2bfloat16_rn(dq[nd][e] * f);
                                to.o[tiled_off(tok[r0], (uint32_t)(hoff + dd), to.nkc1)] = __float2bfloat16_rn(o[nd][e]);
                            }
                            if (r1 < T) {
                                to.dqkv[tiled_off(tok[r1], (uint32_t)(hoff + dd), to.nkc3)] = __float2bfloat16_rn(dq[nd][2 + e] * f);
                                to.o[tiled_off(tok[r1], (uint32_t)(hoff + dd), to.nkc1)] = __float2bfloat16_rn(o[nd][2 + e]);
                            }
                        } else {
                            if (r0 < T) {
                                const uint32_t o0 = rows[r0] + (uint32_t)(hoff + dd);
                                dQ[o0] = dq[nd][e] * f;
                                if (O) O[o0] = o[nd][e];
                            }
                            if (r1 < T) {
                                const uint32_t o1 = rows[r1] + (uint32_t)(hoff + dd);
                                dQ[o1] = dq[nd][2 + e] * f;
                                if (O) O[o1] = o[nd][2 + e];
                            }
                        }
                    }
                }
        }
        // ---- dK^T = Q^T dS, dV^T = dO^T P per 16-wide chunk of the head dimension (M = head dim, N = keys, K = queries):
        //      every warp contributes the partial over its 16 queries, the four partials are summed through `red`
#pragma unroll
        for (int kc = 0; kc < KC; kc++) {
            {
                const int ar = gq * TSW + 8 * m + tq;
                const uint32_t* qth = ws->qt[kc];
                const uint32_t* gth = ws->gt[kc];
                const uint32_t qa0 = qth[ar], qa1 = qth[ar + 8 * TSW], qa2 = qth[ar + 4], qa3 = qth[ar + 8 * TSW + 4];
                const uint32_t ga0 = gth[ar], ga1 = gth[ar + 8 * TSW], ga2 = gth[ar + 4], ga3 = gth[ar + 8 * TSW + 4];
#pragma unroll
                for (int n = 0; n < 7; n++) {
                    float pk[4] = {0.f, 0.f, 0.f, 0.f}, pv[4] = {0.f, 0.f, 0.f, 0.f};
                    mma16816(pk, qa0, qa1, qa2, qa3, movm_trans(dsh[n][0]), movm_trans(dsh[n][1]));
                    mma16816(pv, ga0, ga1, ga2, ga3, movm_trans(ph[n][0]), movm_trans(ph[n][1]));
                    if constexpr (D <= 8) {     // fragment rows gq + 8 (dims 8..15) do not exist and only lanes gq < D hold real dims:
                        if (gq < D) {           // 8-byte stores from a quarter of the lanes = a quarter of the shared-memory wavefronts
                            red2[(m * 14 + n) * 32 + lane] = make_float2(pk[0], pk[1]);
                            red2[(m * 14 + 7 + n) * 32 + lane] = make_float2(pv[0], pv[1]);
                        }
                    } else {
                        ws->red[m][n][lane] = make_float4(pk[0], pk[1], pk[2], pk[3]);
                        ws->red[m][7 + n][lane] = make_float4(pv[0], pv[1], pv[2], pv[3]);
                    }
                }
            }
            __syncthreads();
            // warp m finishes fragments m, m+4, m+8, m+12: (row = dim gq / gq+8 of the chunk, col = key 8 n + 2 tq + e)
#pragma unroll
            for (int i = m; i < 14; i += WARPS) {
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                if constexpr (D <= 8) {
                    if (gq < D) {
                        const float2 p0 = red2[i * 32 + lane], p1 = red2[(14 + i) * 32 + lane], p2 = red2[(28 + i) * 32 + lane], p3 = red2[(42 + i) * 32 + lane];
                        acc[0] = (p0.x + p1.x) + (p2.x + p3.x); acc[1] = (p0.y + p1.y) + (p2.y + p3.y);
                    }
                } else {
                    const float4 p0 = ws->red[0][i][lane], p1 = ws->red[1][i][lane], p2 = ws->red[2][i][lane], p3 = ws->red[3][i][lane];
                    acc[0] = (p0.x + p1.x) + (p2.x + p3.x); acc[1] = (p0.y + p1.y) + (p2.y + p3.y);
                    acc[2] = (p0.z + p1.z) + (p2.z + p3.z); acc[3] = (p0.w + p1.w) + (p2.w + p3.w);
                }
                const bool isk = i < 7;
                const int n = isk ? i : i - 7;
                float* dst = isk ? dK : dV;
                const float f = isk ? scale * inv_sc : inv_sc;
                const int d0 = 16 * kc + gq;
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int key = 8 * n + 2 * tq + e;
                    if (key < T) {
                        if (to.dqkv) {
                            const uint32_t cb = (uint32_t)((isk ? to.inner : 2 * to.inner) + hoff);
                            if (d0 < D) to.dqkv[tiled_off(tok[key], cb + (uint32_t)d0, to.nkc3)] = __float2bfloat16_rn(acc[e] * f);
                            if (d0 + 8 < D) to.dqkv[tiled_off(tok[key], cb + (uint32_t)d0 + 8u, to.nkc3)] = __float2bfloat16_rn(acc[2 + e] * f);
                        } else {
                            const uint32_t o = rows[key] + (uint32_t)hoff;
                            if (d0 < D) dst[o + d0] = acc[e] * f;
                            if (d0 + 8 < D) dst[o + d0 + 8] = acc[2 + e] * f;
                        }
                    }
                }
            }
            if (kc + 1 < KC) __syncthreads();     // `red` is rewritten by the next chunk (the next item passes two barriers first)
        }
    }
    // ---- table gradient: registers -> shared -> one global atomic per entry and CTA
    __syncthreads();
    if (gtable) {
#pragma unroll
        for (int n = 0; n < 7; n++)
#pragma unroll
            for (int e = 0; e < 2; e++)
                if (8 * n + 2 * tq + e < T) {
                    if (r0 < T) atomicAdd(&ws->gtab[qo0 + cj[n][e]], tl[n][e]);
                    if (r1 < T) atomicAdd(&ws->gtab[qo1 + cj[n][e]], tl[n][2 + e]);
                }
        __syncthreads();
        for (int i = threadIdx.x; i < 169; i += WARPS * 32)
            if (ws->gtab[i] != 0.f) atomicAdd(&gtable[i], ws->gtab[i]);
    }
}

template <int D>
int launch_one(const float* Q, const float* K, const float* V, const float* gO, float* dQ, float* dK, float* dV, float* O, const float* table,
               float* gtable, const WinGeom& g, int inner, int nh, long long nitems, cudaStream_t st, const AttnBwdTiledOut& to) {
    const size_t smem = sizeof(ItemSmem<(D + 15) / 16>);
    static DeviceOnce configured;
    if (configured.need()) {
        cudaError_t e = cudaFuncSetAttribute(k_attn_bwd_mma<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("attention backward (mma): cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SF_ERR_CUDA; }
        configured.done();
    }
    long long grid = (long long)sm_count() * (D <= 16 ? 3 : 2);
    if (grid > nitems) grid = nitems;
    k_attn_bwd_mma<D><<<(unsigned)grid, WARPS * 32, smem, st>>>(Q, K, V, gO, dQ, dK, dV, O, table, gtable, g, inner, nh, 1.0f / sqrtf((float)D), nitems, to);
    SF_CHECK_LAUNCH("bwd_attn_core_mma");
    return SF_OK;
}

}  // namespace

bool attn_core_bwd_mma_supported(const WinGeom& g, int d, int nh) {
    // element offsets inside the kernel are 32-bit
    return g.wsh == 7 && g.wsw == 7 && (d == 3 || d == 6 || d == 12 || d == 24 || d == 48) &&
           (long long)g.B * g.Hp * g.Wp * nh * d < (1LL << 32);
}

int launch_attn_core_bwd_mma(const float* Q, const float* K, const float* V, const float* gO, float* dQ, float* dK, float* dV, float* O,
                             const float* table, float* gtable, const WinGeom& g, int inner, int nh, int d, cudaStream_t st, int ld,
                             const AttnBwdTiledOut* tiled) {
    AttnBwdTiledOut to{};
    if (tiled) {
        to = *tiled;
        SF_CHECK_ARG(to.dqkv && to.o && to.inner == inner, "attention backward (mma): tiled outputs need both tensors");
    } else {
        SF_CHECK_ARG(dQ && dK && dV, "attention backward (mma): missing outputs");
    }
    // ld: row stride (elements) shared by all eight tensors -- they may be column blocks of one [tokens x ld] buffer
    if (ld <= 0) ld = inner;
    SF_CHECK_ARG((long long)g.B * g.Hp * g.Wp * ld < (1LL << 32), "attention backward (mma): %d-wide rows exceed the 32-bit offset range", ld);
    const long long nitems = (long long)g.B * g.nWh * g.nWw * nh;
    const double mtok = (double)g.B * g.Hp * g.Wp;
    ProfScope ps(prof_name("bwd_attn_core_mma_d%d", d), (O ? 14.0 : 12.0) * g.T * mtok * inner, (O ? 32.0 : 28.0) * mtok * inner, st);
    switch (d) {
        case 3: return launch_one<3>(Q, K, V, gO, dQ, dK, dV, O, table, gtable, g, ld, nh, nitems, st, to);
        case 6: return launch_one<6>(Q, K, V, gO, dQ, dK, dV, O, table, gtable, g, ld, nh, nitems, st, to);
        case 12: return launch_one<12>(Q, K, V, gO, dQ, dK, dV, O, table, gtable, g, ld, nh, nitems, st, to);
        case 24: return launch_one<24>(Q, K, V, gO, dQ, dK, dV, O, table, gtable, g, ld, nh, nitems, st, to);
        case 48: return launch_one<48>(Q, K, V, gO, dQ, dK, dV, O, table, gtable, g, ld, nh, nitems, st, to);
    }
    set_error("attention backward (mma): head_dim %d is not built", d);
    return SF_ERR_INVALID;
}

}  // namespace sf

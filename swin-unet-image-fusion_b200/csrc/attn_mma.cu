// Attention core on tensor cores for 7x7 windows (a001:317-354), fp16 q/k/v -> bf16 O.
//
// Per (window, head): S = Q K^T and O = P V are 49x49xd / 49xdx49 products -- far too small for a
// 128-row tcgen05 tile (one UMMA would be >60% padding and the accumulator would have to round-trip
// through TMEM for the softmax), so they run as warp-level m16n8k16 HMMA tiles whose accumulator
// fragments stay in registers, where the softmax is applied directly (FlashAttention-2 style):
//
//   warp task = (head, slab of 16 query rows):  7 MMAs  S[16 x 56] = Q_slab K^T   (keys padded 49 -> 56)
//       s = acc * (d^-1/2 log2 e) + bias_frag        bias fragment of the slab lives in registers for the
//       [shift mask on boundary windows]             whole kernel (one table shared by all heads, a001:113-144)
//       row max / row sum: 2 quad shuffles each; p = ex2(s - max)
//       P (fp16) re-used in place as the A fragments of  O[16 x d] = P V   (4 k-steps)
//
// CTAs are persistent over windows.  Per window the CTA stages the 49 token rows (coalesced 16-byte
// loads, shift / partition as index math) into fragment-friendly shared-memory layouts:
// Q [head][row][dim], K [head][key][dim] (row stride padded: conflict-free fragment loads) and
// V^T [head][dim][key].  O goes through shared memory and is written in the UMMA-tiled layout the
// projection GEMM bulk-copies.
#include <cuda_fp16.h>
#include "bf16_kernels.cuh"
#include "tc_common.cuh"

namespace sf {

static constexpr int MT = 49;        // tokens per window
static constexpr int MROWS = 64;     // query rows padded to 4 slabs of 16
static constexpr int MKEYS = 56;     // keys padded to 7 n-tiles of 8
static constexpr int BS = 56;        // bias row stride (floats)
static constexpr int VS = 72;        // V^T row stride (halves): 64 keys + 8 -> conflict-free B fragments
static constexpr int AM_THREADS = 256;

struct AttnMmaArgs {
    const __half* qkv; long long ld; int koff, voff;
    int qkv_nkc;                     // > 0: qkv is UMMA-tiled ([tile][chunk][row][8]) with this many chunks per tile; 0: rows of ld halves
    bf16* O; int o_nkc;              // UMMA-tiled output, o_nkc k-chunks per 128-token tile
    const float* table;
    WinGeom g;
    int nh, d;
    long long nwin;
    float scale_log2e;
};

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ const __half* qkv_chunk_ptr(const AttnMmaArgs& a, long long tok, int col) {
    if (a.qkv_nkc == 0) return a.qkv + tok * a.ld + col;
    return a.qkv + (((tok >> 7) * a.qkv_nkc + (col >> 3)) * 128 + (tok & 127)) * 8 + (col & 7);
}

template <int KSTEPS, int NDT>
struct AttnMmaSmem {
    static constexpr int DS = 16 * KSTEPS + 8;   // Q / K row stride in halves (padded)
    static constexpr int DV = 8 * NDT;           // V^T rows per head
    __host__ __device__ static size_t bytes(int nh, int inner) {
        size_t b = (size_t)MROWS * BS * 4;                   // bias matrix
        b += (size_t)nh * MROWS * DS * 2;                    // Q
        b += (size_t)nh * MKEYS * DS * 2;                    // K
        b += (size_t)nh * DV * VS * 2;                       // V^T
        b += (size_t)MT * 8 + (size_t)MT * 4 + 12;           // rows, regions
        b = (b + 15) & ~(size_t)15;
        b += (size_t)MT * inner * 2;                         // O staging (bf16)
        return b + 16;
    }
};

template <int KSTEPS, int NDT>
__global__ void __launch_bounds__(AM_THREADS, (KSTEPS == 1 && NDT <= 2) ? 3 : 1) k_attn_mma(AttnMmaArgs a) {
    using L = AttnMmaSmem<KSTEPS, NDT>;
    constexpr int DS = L::DS, DV = L::DV;
    extern __shared__ __align__(16) uint8_t smraw[];
    const WinGeom& g = a.g;
    const int nh = a.nh, d = a.d, inner = nh * d;
    float* biasm = reinterpret_cast<float*>(smraw);
    __half* Qs = reinterpret_cast<__half*>(biasm + MROWS * BS);
    __half* Ks = Qs + (size_t)nh * MROWS * DS;
    __half* Vt = Ks + (size_t)nh * MKEYS * DS;
    long long* rows = reinterpret_cast<long long*>((reinterpret_cast<uintptr_t>(Vt + (size_t)nh * DV * VS) + 7) & ~(uintptr_t)7);
    int* regs = reinterpret_cast<int*>(rows + MT);
    bf16* Ob = reinterpret_cast<bf16*>((reinterpret_cast<uintptr_t>(regs + MT) + 15) & ~(uintptr_t)15);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gq = lane >> 2, tq = lane & 3;   // fragment coordinates: row group / column pair
    const int tw = 2 * g.wsw - 1;
    constexpr float LOG2E = 1.4426950408889634f;
    constexpr float NEG = -1e30f;

    // ---- once per CTA: bias matrix (x log2e; padded key columns = -inf), zero the operand padding ----
    for (int i = tid; i < MROWS * BS; i += AM_THREADS) {
        int qi = i / BS, kj = i - qi * BS;
        float v = NEG;
        if (kj < MT) v = (qi < MT) ? LOG2E * a.table[(kj / g.wsw - qi / g.wsw + g.wsh - 1) * tw + (kj % g.wsw - qi % g.wsw + g.wsw - 1)] : 0.f;
        biasm[i] = v;
    }
    {
        uint32_t* z = reinterpret_cast<uint32_t*>(Qs);
        const size_t nwords = ((size_t)nh * MROWS * DS + (size_t)nh * MKEYS * DS + (size_t)nh * DV * VS) / 2;
        for (size_t i = tid; i < nwords; i += AM_THREADS) z[i] = 0u;
    }
    __syncthreads();
    // this warp's slab of 16 query rows
    const int slab = warp & 3;
    const int r0 = slab * 16 + gq, r1 = r0 + 8;
    const __half2 qscale = __float2half2_rn(a.scale_log2e);
    const bool vec_in = ((inner & 7) == 0) && (a.qkv_nkc != 0 || (a.ld & 7) == 0) && ((a.koff & 7) == 0) && ((a.voff & 7) == 0);
    const int nch = inner >> 3;
    const int nchunks = MT * 3 * nch;
    const int nWimg = g.nWh * g.nWw;

    // A thread always stages the same chunks (i = tid + k*256) of every window: decode them once.
    // desc: token | which<<8 (0 q, 1 k, 2 v) ; col: first column in the qkv row ; idx: first smem element
    auto decode = [&](int i, int& tok, int& which, int& col, int& idx, int& dd0) {
        tok = i / (3 * nch);
        int c = i - tok * 3 * nch;
        which = c / nch;
        int ch = c - which * nch;
        col = (which == 0 ? 0 : (which == 1 ? a.koff : a.voff)) + ch * 8;
        int h = (ch * 8) / d;
        dd0 = ch * 8 - h * d;
        idx = which == 0 ? (h * MROWS + tok) * DS + dd0 : (which == 1 ? (h * MKEYS + tok) * DS + dd0 : (h * DV + dd0) * VS + tok);
    };
    int p_tok[2], p_which[2], p_col[2], p_idx[2], p_dd[2];
    // source row of token (ti, tj) of window (b, wh, ww): shift / partition as index math without divisions
    auto src_row = [&](int b, int wh, int ww, int tok) -> long long {
        const int ti = tok / 7, tj = tok - ti * 7;            // compile-time divisor
        int r = wh * 7 + ti + g.sh, c = ww * 7 + tj + g.sw;
        if (r >= g.Hp) r -= g.Hp;
        if (c >= g.Wp) c -= g.Wp;
        return ((long long)b * g.Hp + r) * g.Wp + c;
    };
#pragma unroll
    for (int k = 0; k < 2; k++) {
        p_tok[k] = 0; p_which[k] = 0; p_col[k] = 0; p_idx[k] = 0; p_dd[k] = 0;
        if (tid + k * AM_THREADS < nchunks) decode(tid + k * AM_THREADS, p_tok[k], p_which[k], p_col[k], p_idx[k], p_dd[k]);
    }
    auto scatter = [&](const uint4& raw, int which, int idx, int dd) {
        uint4 rs = raw;
        if (which == 0) {   // scores live in the log2 domain: q *= d^-1/2 * log2(e)
            __half2* h2 = reinterpret_cast<__half2*>(&rs);
#pragma unroll
            for (int e = 0; e < 4; e++) h2[e] = __hmul2(h2[e], qscale);
        }
        const __half* hv = reinterpret_cast<const __half*>(&rs);
        __half* dst = which == 0 ? Qs : (which == 1 ? Ks : Vt);
        const int step = which == 2 ? VS : 1;                                   // next dim of the same head
        const int wrap = which == 0 ? MROWS * DS - (d - 1) : (which == 1 ? MKEYS * DS - (d - 1) : (DV - (d - 1)) * VS);   // first dim of the next head
#pragma unroll
        for (int e = 0; e < 8; e++) {
            dst[idx] = hv[e];
            if (++dd == d) { dd = 0; idx += wrap; } else idx += step;
        }
    };

    // software pipeline: the first two chunks of the NEXT window are fetched into registers while the
    // current window is being computed, so their global latency is never exposed
    uint4 raw[2];
    auto prefetch = [&](long long w) {
        if (!vec_in || w >= a.nwin) return;
        const int b = (int)(w / nWimg), wi = (int)(w - (long long)b * nWimg), wh = wi / g.nWw, ww = wi - wh * g.nWw;
#pragma unroll
        for (int k = 0; k < 2; k++)
            if (tid + k * AM_THREADS < nchunks)
                raw[k] = *reinterpret_cast<const uint4*>(qkv_chunk_ptr(a, src_row(b, wh, ww, p_tok[k]), p_col[k]));
    };
    prefetch(blockIdx.x);

    for (long long win = blockIdx.x; win < a.nwin; win += gridDim.x) {
        // boundary windows of the shifted frame are the only ones whose tokens span several regions (a001:222-247)
        const int wb = (int)(win / nWimg), wimg = (int)(win - (long long)wb * nWimg), wwh = wimg / g.nWw, www = wimg - wwh * g.nWw;
        const bool has_mask = g.shift && (wwh == g.nWh - 1 || www == g.nWw - 1);
        __syncthreads();   // previous window: all fragments consumed, O staging drained
        for (int t = tid; t < MT; t += AM_THREADS) {
            int rg;
            rows[t] = win_token_src(g, (int)win, t, &rg);
            regs[t] = rg;
        }
        // ---- stage q, k, v: 16-byte chunks of the token rows -> fragment layouts ---------------------------
        if (vec_in) {
#pragma unroll
            for (int k = 0; k < 2; k++)
                if (tid + k * AM_THREADS < nchunks) scatter(raw[k], p_which[k], p_idx[k], p_dd[k]);
            for (int i = tid + 2 * AM_THREADS; i < nchunks; i += AM_THREADS) {
                int tok, which, col, idx, dd;
                decode(i, tok, which, col, idx, dd);
                uint4 r = *reinterpret_cast<const uint4*>(qkv_chunk_ptr(a, src_row(wb, wwh, www, tok), col));
                scatter(r, which, idx, dd);
            }
        } else {
            for (int i = tid; i < MT * inner; i += AM_THREADS) {
                int t = i / inner, cc = i - t * inner, h = cc / d, dd = cc - h * d;
                const long long tokr = src_row(wb, wwh, www, t);
                Qs[((size_t)h * MROWS + t) * DS + dd] = __hmul(*qkv_chunk_ptr(a, tokr, cc), __low2half(qscale));
                Ks[((size_t)h * MKEYS + t) * DS + dd] = *qkv_chunk_ptr(a, tokr, a.koff + cc);
                Vt[((size_t)h * DV + dd) * VS + t] = *qkv_chunk_ptr(a, tokr, a.voff + cc);
            }
        }
        __syncthreads();
        prefetch(win + gridDim.x);

        // ---- warp tasks: (head, slab) ----------------------------------------------------------------------
        for (int head = warp >> 2; head < nh; head += AM_THREADS / 128) {
            const __half* Qh = Qs + (size_t)head * MROWS * DS;
            const __half* Kh = Ks + (size_t)head * MKEYS * DS;
            const __half* Vh = Vt + (size_t)head * DV * VS;
            float s[7][4];   // accumulators start at bias * log2e (padded key columns: -1e30)
#pragma unroll
            for (int nt = 0; nt < 7; nt++) {
                float2 lo = *reinterpret_cast<const float2*>(&biasm[r0 * BS + nt * 8 + 2 * tq]);
                float2 hi = *reinterpret_cast<const float2*>(&biasm[r1 * BS + nt * 8 + 2 * tq]);
                s[nt][0] = lo.x; s[nt][1] = lo.y; s[nt][2] = hi.x; s[nt][3] = hi.y;
            }
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ks++) {
                uint32_t af[4];
                af[0] = *reinterpret_cast<const uint32_t*>(Qh + r0 * DS + ks * 16 + 2 * tq);
                af[1] = *reinterpret_cast<const uint32_t*>(Qh + r1 * DS + ks * 16 + 2 * tq);
                af[2] = *reinterpret_cast<const uint32_t*>(Qh + r0 * DS + ks * 16 + 8 + 2 * tq);
                af[3] = *reinterpret_cast<const uint32_t*>(Qh + r1 * DS + ks * 16 + 8 + 2 * tq);
#pragma unroll
                for (int nt = 0; nt < 7; nt++) {
                    const __half* kp = Kh + (nt * 8 + gq) * DS + ks * 16 + 2 * tq;
                    uint32_t b0 = *reinterpret_cast<const uint32_t*>(kp), b1 = *reinterpret_cast<const uint32_t*>(kp + 8);
                    mma16816(s[nt], af, b0, b1);
                }
            }
            // shift mask ; row max
            float m0 = NEG, m1 = NEG;
            if (has_mask) {
                const int rg0 = r0 < MT ? regs[r0] : -1, rg1 = r1 < MT ? regs[r1] : -1;
#pragma unroll
                for (int nt = 0; nt < 7; nt++) {
                    const int c0 = nt * 8 + 2 * tq;
                    const int kg0 = c0 < MT ? regs[c0] : -2, kg1 = c0 + 1 < MT ? regs[c0 + 1] : -2;
                    if (kg0 != rg0 && c0 < MT) s[nt][0] = -1.4426950e10f;
                    if (kg1 != rg0 && c0 + 1 < MT) s[nt][1] = -1.4426950e10f;
                    if (kg0 != rg1 && c0 < MT) s[nt][2] = -1.4426950e10f;
                    if (kg1 != rg1 && c0 + 1 < MT) s[nt][3] = -1.4426950e10f;
                }
            }
#pragma unroll
            for (int nt = 0; nt < 7; nt++) {
                m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
                m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
            }
            m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
            m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
            float l0 = 0.f, l1 = 0.f;
            uint32_t pf[7][2];   // P as fp16 pairs: [nt][0] = row r0, [nt][1] = row r1
#pragma unroll
            for (int nt = 0; nt < 7; nt++) {
                float p0 = ex2f(s[nt][0] - m0), p2 = ex2f(s[nt][2] - m1);
                // n-tile 6 holds keys 48..55: only key 48 exists, its odd column needs no exponential
                float p1 = nt < 6 ? ex2f(s[nt][1] - m0) : 0.f, p3 = nt < 6 ? ex2f(s[nt][3] - m1) : 0.f;
                l0 += p0 + p1; l1 += p2 + p3;
                pf[nt][0] = pack_h2(p0, p1);
                pf[nt][1] = pack_h2(p2, p3);
            }
            l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
            l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
            // O = P V : the score fragments of n-tiles (2j, 2j+1) are the A fragment of key step j
            float o[NDT][4];
#pragma unroll
            for (int dt = 0; dt < NDT; dt++) { o[dt][0] = 0.f; o[dt][1] = 0.f; o[dt][2] = 0.f; o[dt][3] = 0.f; }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint32_t af[4];
                af[0] = pf[2 * j][0]; af[1] = pf[2 * j][1];
                af[2] = (2 * j + 1 < 7) ? pf[2 * j + 1][0] : 0u;
                af[3] = (2 * j + 1 < 7) ? pf[2 * j + 1][1] : 0u;
#pragma unroll
                for (int dt = 0; dt < NDT; dt++) {
                    const __half* vp = Vh + (dt * 8 + gq) * VS + j * 16 + 2 * tq;
                    uint32_t b0 = *reinterpret_cast<const uint32_t*>(vp), b1 = *reinterpret_cast<const uint32_t*>(vp + 8);
                    mma16816(o[dt], af, b0, b1);
                }
            }
            const float i0 = __frcp_rn(l0), i1 = __frcp_rn(l1);
#pragma unroll
            for (int dt = 0; dt < NDT; dt++) {
                const int dd = dt * 8 + 2 * tq;
                if (r0 < MT) {
                    if (dd < d) Ob[r0 * inner + head * d + dd] = __float2bfloat16_rn(o[dt][0] * i0);
                    if (dd + 1 < d) Ob[r0 * inner + head * d + dd + 1] = __float2bfloat16_rn(o[dt][1] * i0);
                }
                if (r1 < MT) {
                    if (dd < d) Ob[r1 * inner + head * d + dd] = __float2bfloat16_rn(o[dt][2] * i1);
                    if (dd + 1 < d) Ob[r1 * inner + head * d + dd + 1] = __float2bfloat16_rn(o[dt][3] * i1);
                }
            }
        }
        __syncthreads();
        // ---- O -> global, UMMA-tiled (chunk (tile, kc, r) at ((tile*o_nkc + kc)*128 + r)*8 elements) ----------
        for (int i = tid; i < MT * a.o_nkc; i += AM_THREADS) {
            int t = i / a.o_nkc, kc = i - t * a.o_nkc;
            const long long tok = rows[t];
            uint4 v = make_uint4(0, 0, 0, 0);
            if (kc * 8 + 8 <= inner) {
                v = *reinterpret_cast<const uint4*>(Ob + t * inner + kc * 8);
            } else if (kc * 8 < inner) {
                bf16 tmp[8];
                for (int e = 0; e < 8; e++) tmp[e] = kc * 8 + e < inner ? Ob[t * inner + kc * 8 + e] : __float2bfloat16_rn(0.f);
                v = *reinterpret_cast<const uint4*>(tmp);
            }
            *reinterpret_cast<uint4*>(a.O + (((tok >> 7) * a.o_nkc + kc) * 128 + (tok & 127)) * 8) = v;
        }
    }
}

template <int KSTEPS, int NDT>
static int launch_attn_mma_t(const AttnMmaArgs& a, cudaStream_t st) {
    const int inner = a.nh * a.d;
    const size_t smem = AttnMmaSmem<KSTEPS, NDT>::bytes(a.nh, inner);
    SF_CHECK_ARG(smem <= 227 * 1024, "attention core: %d heads x %d dims need %zu B of shared memory", a.nh, a.d, smem);
    static thread_local bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_attn_mma<KSTEPS, NDT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) { set_error("attention core: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SF_ERR_CUDA; }
        configured = true;
    }
    int per_sm = (int)(227 * 1024 / (smem + 1024));
    const int cap = (KSTEPS == 1 && NDT <= 2) ? 3 : 1;
    if (per_sm > cap) per_sm = cap;
    if (per_sm < 1) per_sm = 1;
    long long grid = 148LL * per_sm;
    if (grid > a.nwin) grid = a.nwin;
    const double mtok = (double)a.nwin * MT;
    ProfScope ps(prof_name("attn_core_mma_c%d", inner), 4.0 * MT * mtok * inner, 8.0 * mtok * inner, st);
    k_attn_mma<KSTEPS, NDT><<<(unsigned)grid, AM_THREADS, smem, st>>>(a);
    SF_CHECK_LAUNCH("attn_core_mma");
    return SF_OK;
}

bool attn_mma_supported(const WinGeom& g, int d) { return g.T == MT && g.wsh == 7 && g.wsw == 7 && d <= 48; }

// returns SF_ERR_UNSUPPORTED (without setting an error) when the shape is not covered: the caller
// then uses the CUDA-core kernels of attn_core.cu
int launch_attn_core_mma(const bf16* qkv_fp16, long long ld, int qkv_nkc, int koff, int voff, bf16* O, int o_nkc, const float* table,
                         const WinGeom& g, int nh, int d, cudaStream_t st) {
    if (!(g.T == MT && g.wsh == 7 && g.wsw == 7 && o_nkc > 0 && d <= 48)) return SF_ERR_UNSUPPORTED;
    AttnMmaArgs a{};
    a.qkv = reinterpret_cast<const __half*>(qkv_fp16); a.ld = ld; a.qkv_nkc = qkv_nkc; a.koff = koff; a.voff = voff; a.O = O; a.o_nkc = o_nkc;
    a.table = table; a.g = g; a.nh = nh; a.d = d;
    a.nwin = (long long)g.B * g.nWh * g.nWw;
    a.scale_log2e = 1.4426950408889634f / sqrtf((float)d);
    if (a.nwin > 2147483647LL) return SF_ERR_UNSUPPORTED;
    if (d <= 8) return launch_attn_mma_t<1, 1>(a, st);
    if (d <= 16) return launch_attn_mma_t<1, 2>(a, st);
    if (d <= 24) return launch_attn_mma_t<2, 3>(a, st);
    if (d <= 32) return launch_attn_mma_t<2, 4>(a, st);
    return launch_attn_mma_t<3, 6>(a, st);
}

}  // namespace sf

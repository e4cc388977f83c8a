// Device code shared by the fused window-attention kernels (wa_fused.cu: bulk-synchronous, two CTAs per SM;
// wa_ws.cu: warp-specialised, helper warps + attention warps): tile geometry, the gather / LayerNorm producer,
// the cp.async row prefetch and the HMMA attention core reading fp16 q|k|v rows from shared memory.
#pragma once
#include "bf16_kernels.cuh"
#include "hmma_util.cuh"
#include "tc_common.cuh"
#include <type_traits>

namespace sf {
using namespace tc;

static constexpr int WF_THREADS = 256;
static constexpr int WF_T = 49;                       // tokens per window
static constexpr int WF_WIN = 2;                      // windows per tile
static constexpr int WF_ROWS = WF_WIN * WF_T;         // 98 of the 128 UMMA rows carry tokens
static constexpr uint32_t WF_LBO = lbo_padded(128);   // thread-written operands: 2064 B between k-chunks
static constexpr uint32_t WF_SBO = 128;
static constexpr size_t WF_SMEM_LIMIT = 113 * 1024;   // two CTAs per SM

__host__ __device__ static inline uint32_t wf_al(uint32_t v) { return (v + 127u) & ~127u; }

struct WfSmem { uint32_t wq, wkv, wo, bias, a1q, a1kv, a2, qkv, raw, raw_stride, bars, total; };

// DP4: heads of d <= 3 dims padded to 4 columns (HW = 32); otherwise padded to 8 (HW = 64).
// raw: DP4 only -- two staging buffers for the fp32 source rows of the next tiles (cp.async prefetch one tile ahead);
// the 16-byte-head flavour has no room for them next to its 42 KB of q|k|v rows at two CTAs per SM.
template <bool DP4>
__host__ __device__ static inline WfSmem wf_layout(int Kpad, int N2, bool self_attn, int C) {
    constexpr uint32_t HW = DP4 ? 32 : 64, NQKV = 3 * HW, PITCH = DP4 ? 208 : 432;
    WfSmem s{};
    uint32_t o = 0;
    const uint32_t kc = (uint32_t)Kpad >> 3;
    s.wq = o;   o += wf_al(kc * (self_attn ? NQKV : HW) * 16u);
    s.wkv = o;  o += wf_al(self_attn ? 0u : kc * 2u * HW * 16u);
    s.wo = o;   o += wf_al((HW >> 3) * (uint32_t)N2 * 16u);
    s.bias = o; o += wf_al((NQKV + (uint32_t)N2) * 4u);
    s.a1q = o;  o += wf_al(kc * WF_LBO);
    s.a1kv = o; o += wf_al(self_attn ? 0u : kc * WF_LBO);
    s.a2 = o;   o += wf_al((HW >> 3) * WF_LBO);
    s.qkv = o;  o += wf_al((uint32_t)WF_ROWS * PITCH);
    s.raw_stride = DP4 ? (self_attn ? 1u : 2u) * wf_al((uint32_t)WF_ROWS * (uint32_t)C * 4u) : 0u;   // one tile: q rows [+ k/v rows]
    s.raw = o;  o += 2u * s.raw_stride;
    s.bars = o; o += 64;
    s.total = o;
    return s;
}

struct WaFused {
    const float* q_src; const float* kv_src; const float* residual; float* out;
    const float* ln_q_g; const float* ln_q_b; const float* ln_kv_g; const float* ln_kv_b; float eps;
    const bf16* Wq; const float* bq;      // self: stacked q|k|v image (N = 3*HW); cross: q image (N = HW)
    const bf16* Wkv; const float* bkv;    // cross only: k|v image (N = 2*HW)
    const bf16* Wo; const float* bo;      // projection image (N = N2, K = HW)
    const float* table;                   // 13 x 13 relative-position bias table
    WinOrder wo;
    int nwin, C, d, Kpad, N2, self_attn;
    int debug;   // bit 0: skip the attention core (timing experiments only: SWINFUSE_WF_DEBUG)
};

__device__ __forceinline__ uint2 lds64(const __half* p) { return *reinterpret_cast<const uint2*>(p); }
__device__ __forceinline__ uint32_t lds32(const __half* p) { return *reinterpret_cast<const uint32_t*>(p); }

// ---- phase A: gather + LayerNorm + bf16 -> UMMA A operand ---------------------------------------------------------
// LPR lanes per token row, each lane up to four float4 (q4 = l, l + LPR, ...): 16-byte loads, the two lanes of a 32-byte
// sector sit next to each other.  Rows of a window are 7 runs of 7 contiguous tokens of the source map.
// STAGED: the rows were brought to shared memory by wf_prefetch (row r of the tile at r * C floats).
template <int LPR, bool LN, bool STAGED, int NT = WF_THREADS>
__device__ __forceinline__ void wf_produce(uint8_t* sA, const float* __restrict__ src, const float* __restrict__ g,
                                           const float* __restrict__ b, float eps, const WinOrder& wo, uint32_t m0, int nrows,
                                           int C, int tid) {
    const int nf4 = C >> 2;
    const float invc = 1.f / (float)C;
#pragma unroll 1
    for (int base = 0; base < WF_ROWS * LPR; base += NT) {
        const int item = base + tid;
        const int r = item / LPR, l = item & (LPR - 1);
        const bool ok = r < nrows;
        const long long tok = STAGED ? (long long)r : (ok ? win_order_token(wo, m0 + (uint32_t)r) : 0);
        const float4* row = reinterpret_cast<const float4*>(src + tok * C);
        float4 v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int q4 = l + i * LPR;
            v[i] = (ok && q4 < nf4) ? row[q4] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (LN) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 4; i++) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
            for (int o = LPR >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const float mean = s * invc;
            float ss = 0.f;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                if (l + i * LPR < nf4) {
                    const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
                    ss += (dx * dx + dy * dy) + (dz * dz + dw * dw);
                }
            }
#pragma unroll
            for (int o = LPR >> 1; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            const float rstd = rsqrtf(ss * invc + eps);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int q4 = l + i * LPR;
                if (q4 < nf4) {
                    const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + q4), bb = __ldg(reinterpret_cast<const float4*>(b) + q4);
                    v[i].x = (v[i].x - mean) * rstd * gg.x + bb.x;
                    v[i].y = (v[i].y - mean) * rstd * gg.y + bb.y;
                    v[i].z = (v[i].z - mean) * rstd * gg.z + bb.z;
                    v[i].w = (v[i].w - mean) * rstd * gg.w + bb.w;
                }
            }
        }
        if (ok) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int q4 = l + i * LPR;
                if (q4 < nf4)
                    *reinterpret_cast<uint2*>(sA + (uint32_t)(q4 >> 1) * WF_LBO + (uint32_t)r * 16u + (uint32_t)(q4 & 1) * 8u) =
                        make_uint2(pack_bf16x2(v[i].x, v[i].y), pack_bf16x2(v[i].z, v[i].w));
            }
        }
    }
}

// cp.async gather of the token rows of one tile (window order) into a staging buffer: 16 bytes per copy, issued a
// whole tile ahead so that the HBM latency is spent under the previous tile's attention core
template <int NT = WF_THREADS>
__device__ __forceinline__ void wf_prefetch(uint8_t* raw, const float* __restrict__ src, const WinOrder& wo, uint32_t m0, int nrows,
                                            int C, int tid) {
    const int nf4 = C >> 2;
    for (int idx = tid; idx < nrows * nf4; idx += NT) {
        const int r = idx / nf4, q4 = idx - r * nf4;
        const long long tok = win_order_token(wo, m0 + (uint32_t)r);
        cp_async16(raw + ((uint32_t)r * (uint32_t)C + (uint32_t)q4 * 4u) * 4u, src + tok * C + q4 * 4);
    }
}

// ---- softmax without the row maximum ------------------------------------------------------------------------------------
// Scores arrive in the log2 domain (q pre-scaled by d^-1/2 log2 e, bias x log2 e).  softmax is shift invariant, and
// P = 2^s needs no shift at all while every row keeps its dominant terms inside the fp32 / bf16 exponent range: P and v are
// bf16 MMA operands (same exponent range as fp32), the row sum l rides on the ones column of v in the fp32 accumulator.
// A row whose l left [2^-100, 2^100] (|score| beyond ~100: overflow, inf, NaN, or everything flushed to zero) sends the
// warp through the exact path once more: row maximum subtracted first (a001:343, torch softmax).  The fast path saves the
// max tree, two shuffle rounds and one FADD per score, and it removes the longest dependency chain of the core.
static constexpr float WF_L_MIN = 7.8886090522101181e-31f;   // 2^-100
static constexpr float WF_L_MAX = 1.2676506002282294e30f;    // 2^100
static constexpr float WF_MASKED = -1.4426950e10f;           // the reference overwrites masked scores with -1e10 (a001:310), log2 domain

__device__ __forceinline__ void wf_apply_mask(float (&s)[7][4], uint32_t m0, uint32_t m1) {
    if (m0 | m1) {
#pragma unroll
        for (int nt = 0; nt < 7; nt++) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                if ((m0 >> (2 * nt + e)) & 1u) s[nt][e] = WF_MASKED;
                if ((m1 >> (2 * nt + e)) & 1u) s[nt][2 + e] = WF_MASKED;
            }
        }
    }
}
// P = 2^(s - x) as bf16 A fragments; EXACT: x = row maxima (quad reduction), else x = 0 and no subtraction is emitted.
// n-tile 6 holds keys 48..55: only key 48 (column 0, lanes tq == 0) is real; the padded ones carry bias -1e30 -> 2^s = 0
template <bool EXACT>
__device__ __forceinline__ void wf_softmax_p(const float (&s)[7][4], uint32_t (&pf)[7][2]) {
    float x0 = 0.f, x1 = 0.f;
    if (EXACT) {
        x0 = fmaxf(s[0][0], s[0][1]); x1 = fmaxf(s[0][2], s[0][3]);
#pragma unroll
        for (int nt = 1; nt < 7; nt++) {
            x0 = max3f(x0, s[nt][0], s[nt][1]);
            x1 = max3f(x1, s[nt][2], s[nt][3]);
        }
        x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, 1)); x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, 2));
        x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, 1)); x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, 2));
    }
#pragma unroll
    for (int nt = 0; nt < 7; nt++) {
        const float p0 = ex2f(EXACT ? s[nt][0] - x0 : s[nt][0]), p2 = ex2f(EXACT ? s[nt][2] - x1 : s[nt][2]);
        const float p1 = nt < 6 ? ex2f(EXACT ? s[nt][1] - x0 : s[nt][1]) : 0.f, p3 = nt < 6 ? ex2f(EXACT ? s[nt][3] - x1 : s[nt][3]) : 0.f;
        pf[nt][0] = pack_bf16x2(p0, p1);
        pf[nt][1] = pack_bf16x2(p2, p3);
    }
}
__device__ __forceinline__ bool wf_l_bad(float l) { return !(l > WF_L_MIN && l < WF_L_MAX); }

// ---- phase D, d <= 3 (8-byte heads): four heads per warp pass (see k_attn_pack4 in attn_frag.cu) ------------------
// warp = (slab, head parity par); lane (gq, tq) loads the 8 bytes of head hj = 2tq + par of its rows, so one set of
// LDS covers four heads; head 2j + par lives in the k-slots fed by lanes tq == j (A fragment = lane select).
// q, k rows are fp16, v rows bf16 (phase C).
template <int PH, int HW>
__device__ __forceinline__ void wf_attn_pack4(const __half* __restrict__ wbase, uint8_t* __restrict__ sA2, int rowbase,
                                              const float (&bias)[7][4], uint32_t m0, uint32_t m1, int d, int r0, int gq, int tq, int par) {
    const int r1 = r0 + 8;
    const int hj = 2 * tq + par;
    const bool r0ok = r0 < WF_T, r1ok = r1 < WF_T;
    const int r0c = r0ok ? r0 : 48, r1c = r1ok ? r1 : r0c;   // rows that do not exist read an existing one (never stored)
    const __half* pq = wbase + r0c * PH + hj * 4;
    const int q1off = (r1c - r0c) * PH;
    const __half* pk = wbase + gq * PH + HW + hj * 4;
    const __half* pv = wbase + gq * PH + 2 * HW + hj * 4;
    const int t6off = gq == 0 ? 48 * PH : 0;   // key tile 6: key 48 for its lane group, any existing row for the others (bias -1e30)
    uint2 q[2], k[7];
    uint32_t vt[7][2];
    q[0] = lds64(pq); q[1] = lds64(pq + q1off);
#pragma unroll
    for (int nt = 0; nt < 7; nt++) {
        k[nt] = lds64(pk + (nt < 6 ? nt * 8 * PH : t6off));
        const uint2 vr = lds64(pv + (nt < 6 ? nt * 8 * PH : t6off));
        vt[nt][0] = movm_trans(vr.x);   // B fragments of P V, shared by the four heads: [key tile][dims 0,1 / 2,3]
        vt[nt][1] = movm_trans(vr.y);
    }
    uint8_t* o0 = sA2 + (uint32_t)(hj >> 1) * WF_LBO + (uint32_t)(rowbase + r0) * 16u + (uint32_t)(hj & 1) * 8u;
    // One pass over the four heads.  The fast pass (EXACT = false) is straight-line code: key tiles stream through
    // QK^T -> 2^s -> P V two at a time (no row maximum means no dependency between key tiles), so the scheduler overlaps the
    // MUFU work of one head with the tensor-core work of the next; the row-sum check is one vote per task.
    auto pass = [&](auto exact) -> bool {
        constexpr bool EXACT = decltype(exact)::value;
        bool bad = false;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const bool mine = tq == j;
            const uint32_t a0 = mine ? q[0].x : 0u, a1 = mine ? q[1].x : 0u, a2 = mine ? q[0].y : 0u, a3 = mine ? q[1].y : 0u;
            float ox[4], oy[4];   // dims 0,1 / 2,3 of head 2*(col/2)+par; rows r0 | r1
            if (EXACT) {
                float s[7][4];
#pragma unroll
                for (int nt = 0; nt < 7; nt++) mma16816_cd(s[nt], a0, a1, a2, a3, k[nt].x, k[nt].y, bias[nt][0], bias[nt][1], bias[nt][2], bias[nt][3]);
                wf_apply_mask(s, m0, m1);
                uint32_t pf[7][2];
                wf_softmax_p<true>(s, pf);
                mma16816_bf16_z(ox, pf[0][0], pf[0][1], pf[1][0], pf[1][1], vt[0][0], vt[1][0]);
                mma16816_bf16_z(oy, pf[0][0], pf[0][1], pf[1][0], pf[1][1], vt[0][1], vt[1][1]);
#pragma unroll
                for (int jj = 1; jj < 4; jj++) {
                    const uint32_t p0 = pf[2 * jj][0], p1 = pf[2 * jj][1];
                    const uint32_t p2 = (2 * jj + 1 < 7) ? pf[2 * jj + 1][0] : 0u, p3 = (2 * jj + 1 < 7) ? pf[2 * jj + 1][1] : 0u;
                    mma16816_bf16(ox, p0, p1, p2, p3, vt[2 * jj][0], (2 * jj + 1 < 7) ? vt[2 * jj + 1][0] : 0u);
                    mma16816_bf16(oy, p0, p1, p2, p3, vt[2 * jj][1], (2 * jj + 1 < 7) ? vt[2 * jj + 1][1] : 0u);
                }
            } else {
#pragma unroll
                for (int jj = 0; jj < 4; jj++) {   // k-step jj of P V = key tiles 2jj, 2jj + 1
                    uint32_t pf[2][2];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int nt = 2 * jj + h;
                        if (nt < 7) {
                            float s[4];
                            mma16816_cd(s, a0, a1, a2, a3, k[nt].x, k[nt].y, bias[nt][0], bias[nt][1], bias[nt][2], bias[nt][3]);
                            if (m0 | m1) {
#pragma unroll
                                for (int e = 0; e < 2; e++) {
                                    if ((m0 >> (2 * nt + e)) & 1u) s[e] = WF_MASKED;
                                    if ((m1 >> (2 * nt + e)) & 1u) s[2 + e] = WF_MASKED;
                                }
                            }
                            // n-tile 6 holds keys 48..55: only key 48 (column 0, lanes tq == 0) is real; the others carry bias -1e30
                            pf[h][0] = pack_bf16x2(ex2f(s[0]), nt < 6 ? ex2f(s[1]) : 0.f);
                            pf[h][1] = pack_bf16x2(ex2f(s[2]), nt < 6 ? ex2f(s[3]) : 0.f);
                        } else {
                            pf[h][0] = 0u; pf[h][1] = 0u;
                        }
                    }
                    const uint32_t b0x = vt[2 * jj][0], b0y = vt[2 * jj][1];
                    const uint32_t b1x = (2 * jj + 1 < 7) ? vt[2 * jj + 1][0] : 0u, b1y = (2 * jj + 1 < 7) ? vt[2 * jj + 1][1] : 0u;
                    if (jj == 0) {
                        mma16816_bf16_z(ox, pf[0][0], pf[0][1], pf[1][0], pf[1][1], b0x, b1x);
                        mma16816_bf16_z(oy, pf[0][0], pf[0][1], pf[1][0], pf[1][1], b0y, b1y);
                    } else {
                        mma16816_bf16(ox, pf[0][0], pf[0][1], pf[1][0], pf[1][1], b0x, b1x);
                        mma16816_bf16(oy, pf[0][0], pf[0][1], pf[1][0], pf[1][1], b0y, b1y);
                    }
                }
            }
            // softmax row sums = the ones column of v (dim d of every head: zero weights, bias 1); valid on the lanes tq == j
            const float l0 = d == 3 ? oy[1] : (d == 2 ? oy[0] : ox[1]);
            const float l1 = d == 3 ? oy[3] : (d == 2 ? oy[2] : ox[3]);
            if (mine) {   // this lane's accumulator columns are its head: dims (ox[0],ox[1],oy[0],oy[1]) of row r0, [2],[3] of row r1
                if (!EXACT) bad = wf_l_bad(l0) || wf_l_bad(l1);
                const float i0 = rcpf(l0), i1 = rcpf(l1);
                if (r0ok) *reinterpret_cast<uint2*>(o0) = make_uint2(pack_bf16x2(ox[0] * i0, ox[1] * i0), pack_bf16x2(oy[0] * i0, oy[1] * i0));
                if (r1ok) *reinterpret_cast<uint2*>(o0 + 128) = make_uint2(pack_bf16x2(ox[2] * i1, ox[3] * i1), pack_bf16x2(oy[2] * i1, oy[3] * i1));
            }
        }
        return bad;
    };
    if (__any_sync(0xffffffffu, pass(std::false_type{}))) pass(std::true_type{});   // exact pass rewrites the task's O rows
}

// ---- phase D, 5 <= d <= 7 (16-byte heads): one head per warp pass (k_attn_frag<1, 1, false, 8, ., true>) ---------------
// warp = (slab, head parity sub) works through heads sub, sub + 2, sub + 4, sub + 6.
template <int PH, int HW>
__device__ __forceinline__ void wf_attn_dp8(const __half* __restrict__ wbase, uint8_t* __restrict__ sA2, int rowbase,
                                            const float (&bias)[7][4], uint32_t m0, uint32_t m1, int d, int r0, int gq, int tq, int sub,
                                            int lane) {
    const int r1 = r0 + 8;
    const bool r0ok = r0 < WF_T, r1ok = r1 < WF_T;
    const int r0c = r0ok ? r0 : 48, r1c = r1ok ? r1 : r0c;
    const __half* pq = wbase + r0c * PH + 4 * tq;
    const int q1off = (r1c - r0c) * PH;
    const __half* pk = wbase + gq * PH + HW + 4 * tq;
    const __half* pv = wbase + gq * PH + 2 * HW + 2 * tq;
    const int t6off = gq == 0 ? 48 * PH : 0;
    const int lsrc = (lane & ~3) | ((d & 7) >> 1);   // quad lane that holds column d (the ones column) of the head's d-tile
    // one pass over the warp's four heads; the fast pass (EXACT = false) streams the key tiles two at a time through
    // QK^T -> 2^s -> P V and checks the row sums with ONE vote per task (see wf_attn_pack4)
    auto pass = [&](auto exact) -> bool {
        constexpr bool EXACT = decltype(exact)::value;
        bool bad = false;
#pragma unroll 2
        for (int hi = 0; hi < 4; hi++) {
            const int head = sub + 2 * hi;
            const int uoff = head * 8;
            uint2 q0 = lds64(pq + uoff), q1 = lds64(pq + q1off + uoff);
            // k-slots 8..15 of the single k-step hold the NEXT head's columns: zero them on the Q side
            if (tq >= 2) { q0 = make_uint2(0u, 0u); q1 = make_uint2(0u, 0u); }
            float o[4];
            if (EXACT) {
                float s[7][4];
                uint32_t vb[7];
#pragma unroll
                for (int nt = 0; nt < 7; nt++) {
                    const uint2 kk = lds64(pk + (nt < 6 ? nt * 8 * PH : t6off) + uoff);
                    mma16816_cd(s[nt], q0.x, q1.x, q0.y, q1.y, kk.x, kk.y, bias[nt][0], bias[nt][1], bias[nt][2], bias[nt][3]);
                    vb[nt] = movm_trans(lds32(pv + (nt < 6 ? nt * 8 * PH : t6off) + uoff));
                }
                wf_apply_mask(s, m0, m1);
                uint32_t pf[7][2];
                wf_softmax_p<true>(s, pf);
                mma16816_bf16_z(o, pf[0][0], pf[0][1], pf[1][0], pf[1][1], vb[0], vb[1]);
#pragma unroll
                for (int j = 1; j < 4; j++) {
                    const uint32_t a2 = (2 * j + 1 < 7) ? pf[2 * j + 1][0] : 0u, a3 = (2 * j + 1 < 7) ? pf[2 * j + 1][1] : 0u;
                    mma16816_bf16(o, pf[2 * j][0], pf[2 * j][1], a2, a3, vb[2 * j], (2 * j + 1 < 7) ? vb[2 * j + 1] : 0u);
                }
            } else {
#pragma unroll
                for (int jj = 0; jj < 4; jj++) {   // k-step jj of P V = key tiles 2jj, 2jj + 1
                    uint32_t pf[2][2], vb[2];
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int nt = 2 * jj + h;
                        if (nt < 7) {
                            const uint2 kk = lds64(pk + (nt < 6 ? nt * 8 * PH : t6off) + uoff);
                            vb[h] = movm_trans(lds32(pv + (nt < 6 ? nt * 8 * PH : t6off) + uoff));
                            float s[4];
                            mma16816_cd(s, q0.x, q1.x, q0.y, q1.y, kk.x, kk.y, bias[nt][0], bias[nt][1], bias[nt][2], bias[nt][3]);
                            if (m0 | m1) {
#pragma unroll
                                for (int e = 0; e < 2; e++) {
                                    if ((m0 >> (2 * nt + e)) & 1u) s[e] = WF_MASKED;
                                    if ((m1 >> (2 * nt + e)) & 1u) s[2 + e] = WF_MASKED;
                                }
                            }
                            pf[h][0] = pack_bf16x2(ex2f(s[0]), nt < 6 ? ex2f(s[1]) : 0.f);
                            pf[h][1] = pack_bf16x2(ex2f(s[2]), nt < 6 ? ex2f(s[3]) : 0.f);
                        } else {
                            pf[h][0] = 0u; pf[h][1] = 0u; vb[h] = 0u;
                        }
                    }
                    if (jj == 0) mma16816_bf16_z(o, pf[0][0], pf[0][1], pf[1][0], pf[1][1], vb[0], vb[1]);
                    else mma16816_bf16(o, pf[0][0], pf[0][1], pf[1][0], pf[1][1], vb[0], vb[1]);
                }
            }
            // row sums = column d of P V (the ones column): quad lane (d % 8) / 2, element d % 2
            const float c0 = (d & 1) ? o[1] : o[0], c1 = (d & 1) ? o[3] : o[2];
            const float l0 = __shfl_sync(0xffffffffu, c0, lsrc), l1 = __shfl_sync(0xffffffffu, c1, lsrc);
            if (!EXACT) bad = bad || wf_l_bad(l0) || wf_l_bad(l1);
            const float i0 = rcpf(l0), i1 = rcpf(l1);
            uint8_t* o0 = sA2 + (uint32_t)head * WF_LBO + (uint32_t)(rowbase + r0) * 16u + (uint32_t)tq * 4u;
            if (r0ok) *reinterpret_cast<uint32_t*>(o0) = pack_bf16x2(o[0] * i0, o[1] * i0);
            if (r1ok) *reinterpret_cast<uint32_t*>(o0 + 128) = pack_bf16x2(o[2] * i1, o[3] * i1);
        }
        return bad;
    };
    if (__any_sync(0xffffffffu, pass(std::false_type{}))) pass(std::true_type{});   // exact pass rewrites the task's O rows
}

// wa_ws.cu: warp-specialised flavour for 8-byte heads (helper warps + attention warps, one CTA per SM)
bool wa_ws_supported(const WinGeom& g, int C, int nh, int d, bool self_attn);
int launch_wa_ws(const WaFused& a, cudaStream_t st);

}  // namespace sf

// Fused window-attention operator for the narrow stages (C <= 64, 8 heads, 7x7 windows):
//
//     out = residual + W_o . Attn( LN_q(q_src), LN_kv(kv_src) ) + b_o          (a001:448-474 + a004:29-38)
//
// ONE persistent kernel; q|k|v and the attention output O never leave the SM.  A tile is two windows
// (98 token rows of a 128-row UMMA tile); per tile the CTA (8 warps, two CTAs per SM so that one CTA's
// memory / tensor-core latencies are covered by the other's softmax work) runs
//
//   A  gather    token rows of the two windows straight from the un-shifted, un-partitioned fp32 map (cyclic shift
//                a001:419-446 and window partition a001:154-172 are index math: win_order_token), LayerNorm in
//                registers (2 or 4 lanes per row), bf16 -> A1 in shared memory, UMMA K-major layout
//   B  tcgen05   D1[128 x 3*HW] = A1 . W_qkv^T  (cross attention: q from A1(q_src), k|v from A1(kv_src)); weights
//                resident in shared memory for the life of the CTA (one bulk copy), accumulator in TMEM
//   C  q|k|v     tcgen05.ld D1 -> + bias -> fp16 rows in shared memory (head-padded columns, q pre-scaled by
//                d^-1/2 log2 e and the ones column of v folded into the packed weights: see bf16_path.cu)
//   D  core      per (window, 16-row slab, head): S = QK^T + bias -> shift mask -> softmax -> PV on m16n8k16 HMMA
//                fragments held in registers (the 49x49xd products are far too small for a 128-row UMMA tile and
//                the softmax wants the scores in registers, where mma.sync leaves them); operands are LDS from
//                the fp16 rows; O -> bf16 A2 in shared memory, UMMA layout
//   E  tcgen05   D2[128 x C] = A2 . W_o^T
//   F  scatter   tcgen05.ld D2 -> + b_o + residual -> fp32 rows written back window-reversed and un-shifted
//                (a001:373-398, 442-445)
//
// HBM traffic per call is the algorithmic minimum: every source row read once (the residual re-read is an L2 hit
// a few microseconds after the gather), every output row written once, weights once per CTA.
#include <cstdlib>
#include "bf16_kernels.cuh"
#include "hmma_util.cuh"
#include "tc_common.cuh"

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
template <int LPR, bool LN, bool STAGED>
__device__ __forceinline__ void wf_produce(uint8_t* sA, const float* __restrict__ src, const float* __restrict__ g,
                                           const float* __restrict__ b, float eps, const WinOrder& wo, uint32_t m0, int nrows,
                                           int C, int tid) {
    const int nf4 = C >> 2;
    const float invc = 1.f / (float)C;
#pragma unroll 1
    for (int base = 0; base < WF_ROWS * LPR; base += WF_THREADS) {
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
__device__ __forceinline__ void wf_prefetch(uint8_t* raw, const float* __restrict__ src, const WinOrder& wo, uint32_t m0, int nrows,
                                            int C, int tid) {
    const int nf4 = C >> 2;
    for (int idx = tid; idx < nrows * nf4; idx += WF_THREADS) {
        const int r = idx / nf4, q4 = idx - r * nf4;
        const long long tok = win_order_token(wo, m0 + (uint32_t)r);
        cp_async16(raw + ((uint32_t)r * (uint32_t)C + (uint32_t)q4 * 4u) * 4u, src + tok * C + q4 * 4);
    }
}

// ---- phase D, d <= 3 (8-byte heads): four heads per warp pass (see k_attn_pack4 in attn_frag.cu) ------------------
// warp = (slab, head parity par); lane (gq, tq) loads the 8 bytes of head hj = 2tq + par of its rows, so one set of
// LDS covers four heads; head 2j + par lives in the k-slots fed by lanes tq == j (A fragment = lane select).
template <int PH, int HW>
__device__ __forceinline__ void wf_attn_pack4(const __half* __restrict__ wbase, uint8_t* __restrict__ sA2, int rowbase,
                                              const float (&bias)[7][4], uint32_t m0, uint32_t m1, int d, int r0, int gq, int tq, int par) {
    constexpr float MASKED = -1.4426950e10f;   // the reference overwrites masked scores with -1e10 (a001:310), log2 domain
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
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const bool mine = tq == j;
        const uint32_t a0 = mine ? q[0].x : 0u, a1 = mine ? q[1].x : 0u, a2 = mine ? q[0].y : 0u, a3 = mine ? q[1].y : 0u;
        float s[7][4];
#pragma unroll
        for (int nt = 0; nt < 7; nt++) { s[nt][0] = bias[nt][0]; s[nt][1] = bias[nt][1]; s[nt][2] = bias[nt][2]; s[nt][3] = bias[nt][3]; }
#pragma unroll
        for (int nt = 0; nt < 7; nt++) mma16816(s[nt], a0, a1, a2, a3, k[nt].x, k[nt].y);
        if (m0 | m1) {
#pragma unroll
            for (int nt = 0; nt < 7; nt++) {
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    if ((m0 >> (2 * nt + e)) & 1u) s[nt][e] = MASKED;
                    if ((m1 >> (2 * nt + e)) & 1u) s[nt][2 + e] = MASKED;
                }
            }
        }
        float x0 = fmaxf(s[0][0], s[0][1]), x1 = fmaxf(s[0][2], s[0][3]);
#pragma unroll
        for (int nt = 1; nt < 7; nt++) {
            x0 = max3f(x0, s[nt][0], s[nt][1]);
            x1 = max3f(x1, s[nt][2], s[nt][3]);
        }
        x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, 1)); x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, 2));
        x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, 1)); x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, 2));
        uint32_t pf[7][2];
#pragma unroll
        for (int nt = 0; nt < 7; nt++) {
            // n-tile 6 holds keys 48..55: only key 48 (column 0, lanes tq == 0) is real; the padded ones give ex2(-1e30) = 0
            const float p0 = ex2f(s[nt][0] - x0), p2 = ex2f(s[nt][2] - x1);
            const float p1 = nt < 6 ? ex2f(s[nt][1] - x0) : 0.f, p3 = nt < 6 ? ex2f(s[nt][3] - x1) : 0.f;
            pf[nt][0] = pack_h2(p0, p1);
            pf[nt][1] = pack_h2(p2, p3);
        }
        float ox[4] = {0.f, 0.f, 0.f, 0.f}, oy[4] = {0.f, 0.f, 0.f, 0.f};   // dims 0,1 / 2,3 of head 2*(col/2)+par; rows r0 | r1
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
            const uint32_t p0 = pf[2 * jj][0], p1 = pf[2 * jj][1];
            const uint32_t p2 = (2 * jj + 1 < 7) ? pf[2 * jj + 1][0] : 0u, p3 = (2 * jj + 1 < 7) ? pf[2 * jj + 1][1] : 0u;
            mma16816(ox, p0, p1, p2, p3, vt[2 * jj][0], (2 * jj + 1 < 7) ? vt[2 * jj + 1][0] : 0u);
            mma16816(oy, p0, p1, p2, p3, vt[2 * jj][1], (2 * jj + 1 < 7) ? vt[2 * jj + 1][1] : 0u);
        }
        if (mine) {   // this lane's accumulator columns are its head: dims (ox[0],ox[1],oy[0],oy[1]) of row r0, [2],[3] of row r1
            // softmax row sums = the ones column of v (dim d of every head: zero weights, bias 1)
            const float l0 = d == 3 ? oy[1] : (d == 2 ? oy[0] : ox[1]);
            const float l1 = d == 3 ? oy[3] : (d == 2 ? oy[2] : ox[3]);
            const float i0 = rcpf(l0), i1 = rcpf(l1);
            if (r0ok) *reinterpret_cast<uint2*>(o0) = make_uint2(pack_bf16x2(ox[0] * i0, ox[1] * i0), pack_bf16x2(oy[0] * i0, oy[1] * i0));
            if (r1ok) *reinterpret_cast<uint2*>(o0 + 128) = make_uint2(pack_bf16x2(ox[2] * i1, ox[3] * i1), pack_bf16x2(oy[2] * i1, oy[3] * i1));
        }
    }
}

// ---- phase D, 5 <= d <= 7 (16-byte heads): one head per warp pass (k_attn_frag<1, 1, false, 8, ., true>) ---------------
// warp = (slab, head parity sub) works through heads sub, sub + 2, sub + 4, sub + 6.
template <int PH, int HW>
__device__ __forceinline__ void wf_attn_dp8(const __half* __restrict__ wbase, uint8_t* __restrict__ sA2, int rowbase,
                                            const float (&bias)[7][4], uint32_t m0, uint32_t m1, int d, int r0, int gq, int tq, int sub,
                                            int lane) {
    constexpr float MASKED = -1.4426950e10f;
    const int r1 = r0 + 8;
    const bool r0ok = r0 < WF_T, r1ok = r1 < WF_T;
    const int r0c = r0ok ? r0 : 48, r1c = r1ok ? r1 : r0c;
    const __half* pq = wbase + r0c * PH + 4 * tq;
    const int q1off = (r1c - r0c) * PH;
    const __half* pk = wbase + gq * PH + HW + 4 * tq;
    const __half* pv = wbase + gq * PH + 2 * HW + 2 * tq;
    const int t6off = gq == 0 ? 48 * PH : 0;
    const int lsrc = (lane & ~3) | ((d & 7) >> 1);   // quad lane that holds column d (the ones column) of the head's d-tile
#pragma unroll 1
    for (int hi = 0; hi < 4; hi++) {
        const int head = sub + 2 * hi;
        const int uoff = head * 8;
        uint2 q0 = lds64(pq + uoff), q1 = lds64(pq + q1off + uoff);
        // k-slots 8..15 of the single k-step hold the NEXT head's columns: zero them on the Q side
        if (tq >= 2) { q0 = make_uint2(0u, 0u); q1 = make_uint2(0u, 0u); }
        float s[7][4];
#pragma unroll
        for (int nt = 0; nt < 7; nt++) { s[nt][0] = bias[nt][0]; s[nt][1] = bias[nt][1]; s[nt][2] = bias[nt][2]; s[nt][3] = bias[nt][3]; }
#pragma unroll
        for (int nt = 0; nt < 7; nt++) {
            const uint2 kk = lds64(pk + (nt < 6 ? nt * 8 * PH : t6off) + uoff);
            mma16816(s[nt], q0.x, q1.x, q0.y, q1.y, kk.x, kk.y);
        }
        uint32_t vb[7];
#pragma unroll
        for (int kt = 0; kt < 7; kt++) vb[kt] = movm_trans(lds32(pv + (kt < 6 ? kt * 8 * PH : t6off) + uoff));
        if (m0 | m1) {
#pragma unroll
            for (int nt = 0; nt < 7; nt++) {
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    if ((m0 >> (2 * nt + e)) & 1u) s[nt][e] = MASKED;
                    if ((m1 >> (2 * nt + e)) & 1u) s[nt][2 + e] = MASKED;
                }
            }
        }
        float x0 = fmaxf(s[0][0], s[0][1]), x1 = fmaxf(s[0][2], s[0][3]);
#pragma unroll
        for (int nt = 1; nt < 7; nt++) {
            x0 = max3f(x0, s[nt][0], s[nt][1]);
            x1 = max3f(x1, s[nt][2], s[nt][3]);
        }
        x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, 1)); x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, 2));
        x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, 1)); x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, 2));
        uint32_t pf[7][2];
#pragma unroll
        for (int nt = 0; nt < 7; nt++) {
            const float p0 = ex2f(s[nt][0] - x0), p2 = ex2f(s[nt][2] - x1);
            const float p1 = nt < 6 ? ex2f(s[nt][1] - x0) : 0.f, p3 = nt < 6 ? ex2f(s[nt][3] - x1) : 0.f;
            pf[nt][0] = pack_h2(p0, p1);
            pf[nt][1] = pack_h2(p2, p3);
        }
        float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t a2 = (2 * j + 1 < 7) ? pf[2 * j + 1][0] : 0u, a3 = (2 * j + 1 < 7) ? pf[2 * j + 1][1] : 0u;
            mma16816(o, pf[2 * j][0], pf[2 * j][1], a2, a3, vb[2 * j], (2 * j + 1 < 7) ? vb[2 * j + 1] : 0u);
        }
        // row sums = column d of P V (the ones column): quad lane (d % 8) / 2, element d % 2
        const float c0 = (d & 1) ? o[1] : o[0], c1 = (d & 1) ? o[3] : o[2];
        const float i0 = rcpf(__shfl_sync(0xffffffffu, c0, lsrc)), i1 = rcpf(__shfl_sync(0xffffffffu, c1, lsrc));
        uint8_t* o0 = sA2 + (uint32_t)head * WF_LBO + (uint32_t)(rowbase + r0) * 16u + (uint32_t)tq * 4u;
        if (r0ok) *reinterpret_cast<uint32_t*>(o0) = pack_bf16x2(o[0] * i0, o[1] * i0);
        if (r1ok) *reinterpret_cast<uint32_t*>(o0 + 128) = pack_bf16x2(o[2] * i1, o[3] * i1);
    }
}

template <bool DP4>
__global__ void __launch_bounds__(WF_THREADS, 2) k_wa_fused(WaFused p) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int HW = DP4 ? 32 : 64, NQKV = 3 * HW;
    constexpr uint32_t PITCH = DP4 ? 208 : 432;   // bytes per fp16 q|k|v row (16-byte multiple, rotates banks row to row)
    constexpr int PH = (int)PITCH / 2;
    constexpr int LPR = DP4 ? 2 : 4;
    constexpr bool RAW = DP4;                     // staged source rows (cp.async one tile ahead)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gq = lane >> 2, tq = lane & 3;
    const int Kpad = p.Kpad, N2 = p.N2;
    const bool self_attn = p.self_attn != 0;
    const WfSmem L = wf_layout<DP4>(Kpad, N2, self_attn, p.C);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
    uint64_t* w_full = bars;
    uint64_t* d1_full = bars + 1;
    uint64_t* d2_full = bars + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
    constexpr uint32_t D2_COL = ((uint32_t)NQKV + 31u) & ~31u;
    const uint32_t ncols = tmem_cols_pow2(D2_COL + (((uint32_t)N2 + 31u) & ~31u));
    const int ntiles = (p.nwin + WF_WIN - 1) / WF_WIN;
    const uint32_t kc1 = (uint32_t)Kpad >> 3;
    const uint32_t wq_bytes = kc1 * (self_attn ? (uint32_t)NQKV : (uint32_t)HW) * 16u;
    const uint32_t wkv_bytes = self_attn ? 0u : kc1 * 2u * (uint32_t)HW * 16u;
    const uint32_t wo_bytes = ((uint32_t)HW >> 3) * (uint32_t)N2 * 16u;

    if (tid == 0) {
        mbar_init(w_full, 1); mbar_init(d1_full, 1); mbar_init(d2_full, 1);
        fence_mbar_init();
        // the weight images stay in shared memory for the life of the CTA
        mbar_arrive_expect_tx(w_full, wq_bytes + wkv_bytes + wo_bytes);
        bulk_g2s(smem + L.wq, p.Wq, wq_bytes, w_full);
        if (!self_attn) bulk_g2s(smem + L.wkv, p.Wkv, wkv_bytes, w_full);
        bulk_g2s(smem + L.wo, p.Wo, wo_bytes, w_full);
    }
    if (warp == 1) tmem_alloc(tmem_slot, ncols);
    {
        // A1 / A2: rows 98..127 and the K padding columns are never written again and must be zero
        uint4* z = reinterpret_cast<uint4*>(smem + L.a1q);
        const int n16 = (int)((L.qkv - L.a1q) >> 4);
        for (int i = tid; i < n16; i += WF_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
        float* sb = reinterpret_cast<float*>(smem + L.bias);
        for (int i = tid; i < NQKV; i += WF_THREADS) sb[i] = self_attn ? __ldg(p.bq + i) : (i < HW ? __ldg(p.bq + i) : __ldg(p.bkv + i - HW));
        for (int i = tid; i < N2; i += WF_THREADS) sb[NQKV + i] = __ldg(p.bo + i);
    }
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const float* sbias = reinterpret_cast<const float*>(smem + L.bias);

    // ---- per-warp constants of the attention core: slab of 16 query rows, head parity -------------------------------
    const int slab = warp & 3, par = warp >> 2;
    const int r0 = slab * 16 + gq;
    float bias[7][4];
    const SlabMask sm = slab_bias_and_mask(bias, p.table, r0, r0 + 8, tq);
    const WinGeom& g = p.wo.g;
    const int rb = warp & 3, eg = warp >> 2;   // TMEM lane quarter (hardware: warp id % 4) / which half of the columns
    const int row = rb * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(rb * 32) << 16);
    const uint32_t raw_kv = self_attn ? 0u : L.raw_stride / 2u;   // k/v rows follow the q rows inside a staging buffer

    auto tile_rows = [&](int t) { return min(WF_WIN, p.nwin - t * WF_WIN) * WF_T; };
    // staged source rows of tile t -> staging buffer `buf` (no wait)
    auto prefetch = [&](int t, uint32_t buf) {
        uint8_t* raw = smem + L.raw + buf * L.raw_stride;
        wf_prefetch(raw, p.q_src, p.wo, (uint32_t)t * WF_ROWS, tile_rows(t), p.C, tid);
        if (!self_attn) wf_prefetch(raw + raw_kv, p.kv_src, p.wo, (uint32_t)t * WF_ROWS, tile_rows(t), p.C, tid);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // phase A of tile t: rows -> LayerNorm -> bf16 A1 (UMMA layout)
    auto produce = [&](int t, uint32_t buf) {
        const uint32_t m0row = (uint32_t)t * WF_ROWS;
        const int nrows = tile_rows(t);
        const float* qs = RAW ? reinterpret_cast<const float*>(smem + L.raw + buf * L.raw_stride) : p.q_src;
        const float* kvs = RAW ? reinterpret_cast<const float*>(smem + L.raw + buf * L.raw_stride + raw_kv) : p.kv_src;
        if (p.ln_q_g) wf_produce<LPR, true, RAW>(smem + L.a1q, qs, p.ln_q_g, p.ln_q_b, p.eps, p.wo, m0row, nrows, p.C, tid);
        else wf_produce<LPR, false, RAW>(smem + L.a1q, qs, nullptr, nullptr, p.eps, p.wo, m0row, nrows, p.C, tid);
        if (!self_attn) {
            if (p.ln_kv_g) wf_produce<LPR, true, RAW>(smem + L.a1kv, kvs, p.ln_kv_g, p.ln_kv_b, p.eps, p.wo, m0row, nrows, p.C, tid);
            else wf_produce<LPR, false, RAW>(smem + L.a1kv, kvs, nullptr, nullptr, p.eps, p.wo, m0row, nrows, p.C, tid);
        }
        fence_async_smem();
    };
    // phase B: q|k|v projection of the tile in A1 -> D1 (one thread)
    auto issue_mma1 = [&]() {
        const uint32_t a1q = smem_u32(smem + L.a1q), wq = smem_u32(smem + L.wq);
        if (self_attn) {
            const uint32_t lbo_w = (uint32_t)NQKV * 16u, idesc = make_idesc_bf16(128, (uint32_t)NQKV);
            for (uint32_t ks = 0; ks < (uint32_t)Kpad >> 4; ks++)
                umma_bf16(tmem_base, make_smem_desc(a1q + ks * 2u * WF_LBO, WF_LBO, WF_SBO),
                          make_smem_desc(wq + ks * 2u * lbo_w, lbo_w, WF_SBO), idesc, ks > 0);
        } else {
            const uint32_t a1kv = smem_u32(smem + L.a1kv), wkv = smem_u32(smem + L.wkv);
            const uint32_t lbo_q = (uint32_t)HW * 16u, lbo_kv = 2u * (uint32_t)HW * 16u;
            const uint32_t idq = make_idesc_bf16(128, (uint32_t)HW), idkv = make_idesc_bf16(128, 2u * (uint32_t)HW);
            for (uint32_t ks = 0; ks < (uint32_t)Kpad >> 4; ks++)
                umma_bf16(tmem_base, make_smem_desc(a1q + ks * 2u * WF_LBO, WF_LBO, WF_SBO),
                          make_smem_desc(wq + ks * 2u * lbo_q, lbo_q, WF_SBO), idq, ks > 0);
            for (uint32_t ks = 0; ks < (uint32_t)Kpad >> 4; ks++)
                umma_bf16(tmem_base + (uint32_t)HW, make_smem_desc(a1kv + ks * 2u * WF_LBO, WF_LBO, WF_SBO),
                          make_smem_desc(wkv + ks * 2u * lbo_kv, lbo_kv, WF_SBO), idkv, ks > 0);
        }
        umma_commit(d1_full);
    };
    // phase C: D1 -> + bias -> fp16 q|k|v rows in shared memory
    auto qkv_rows = [&](int nrows) {
        uint8_t* qrow = smem + L.qkv + (uint32_t)row * PITCH;
#pragma unroll
        for (int c16 = 0; c16 < NQKV / 2; c16 += 16) {
            const int col = eg * (NQKV / 2) + c16;
            float v[16];
            tmem_ld16(tlane + (uint32_t)col, v);
            if (row < nrows) {
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const float4 bb = *reinterpret_cast<const float4*>(sbias + col + i);
                    v[i] += bb.x; v[i + 1] += bb.y; v[i + 2] += bb.z; v[i + 3] += bb.w;
                }
                uint4* dst = reinterpret_cast<uint4*>(qrow + col * 2);
                dst[0] = make_uint4(pack_f16x2(v[0], v[1]), pack_f16x2(v[2], v[3]), pack_f16x2(v[4], v[5]), pack_f16x2(v[6], v[7]));
                dst[1] = make_uint4(pack_f16x2(v[8], v[9]), pack_f16x2(v[10], v[11]), pack_f16x2(v[12], v[13]), pack_f16x2(v[14], v[15]));
            }
        }
        tc_fence_before_sync();
    };

    // ---- software pipeline over this CTA's tiles t_0, t_1, ... (t_i = blockIdx.x + i * gridDim.x) ------------------------
    //   iteration i:   [prefetch rows of t_{i+2}]  A(t_{i+1})  D(t_i)  |sync|  issue E(t_i), B(t_{i+1})   F(t_i)  C(t_{i+1})  |sync|
    // so the gather latency of a tile is spent two attention phases earlier, the q|k|v GEMM of the next tile runs under
    // the scatter of this one, and a tile costs two block barriers.
    int t = blockIdx.x;
    if (t < ntiles) {
        if (RAW) {
            prefetch(t, 0u);
            if (t + (int)gridDim.x < ntiles) prefetch(t + gridDim.x, 1u);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            if (t + (int)gridDim.x >= ntiles) asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
        }
        produce(t, 0u);
        __syncthreads();
        if (tid == 0) {
            mbar_wait(w_full, 0);
            tc_fence_after_sync();
            issue_mma1();
        }
        __syncwarp();
        mbar_wait_relaxed(d1_full, 0u);
        __syncwarp();
        tc_fence_after_sync();
        qkv_rows(tile_rows(t));
    }
    uint32_t it = 0;
#pragma unroll 1
    for (; t < ntiles; t += gridDim.x, it++) {
        const int tn = t + gridDim.x;
        const bool has_next = tn < ntiles;
        const int nw = min(WF_WIN, p.nwin - t * WF_WIN);
        const int nrows = nw * WF_T;
        const uint32_t m0row = (uint32_t)t * WF_ROWS;
        if (RAW) asm volatile("cp.async.wait_group 0;" ::: "memory");   // this thread's share of the rows of t_{i+1} has landed
        __syncthreads();   // q|k|v rows of t_i and the staged rows of t_{i+1} are visible to every warp
        if (RAW && tn + (int)gridDim.x < ntiles) prefetch(tn + gridDim.x, it & 1u);   // buffer of t_i: consumed an iteration ago
        if (has_next) produce(tn, (it + 1u) & 1u);

        // ---- D: attention core of t_i, window by window ---------------------------------------------------------------------
        if (!(p.debug & 1)) {
#pragma unroll 1
            for (int w = 0; w < nw; w++) {
                uint32_t m0 = 0, m1 = 0;
                if (g.shift) {   // boundary windows of the shifted frame are the only ones whose tokens span several regions
                    const uint32_t win = (uint32_t)(t * WF_WIN + w);
                    const uint32_t wi = win - fdiv(win, p.wo.dnW) * (uint32_t)(g.nWh * g.nWw);
                    const uint32_t wh = fdiv(wi, p.wo.dnWw), ww = wi - wh * (uint32_t)g.nWw;
                    if (wh == (uint32_t)g.nWh - 1) { m0 |= sm.mh0; m1 |= sm.mh1; }
                    if (ww == (uint32_t)g.nWw - 1) { m0 |= sm.mw0; m1 |= sm.mw1; }
                }
                const __half* wbase = reinterpret_cast<const __half*>(smem + L.qkv) + w * WF_T * PH;
                if (DP4) wf_attn_pack4<PH, HW>(wbase, smem + L.a2, w * WF_T, bias, m0, m1, p.d, r0, gq, tq, par);
                else wf_attn_dp8<PH, HW>(wbase, smem + L.a2, w * WF_T, bias, m0, m1, p.d, r0, gq, tq, par, lane);
            }
        }
        fence_async_smem();
        __syncthreads();   // A2(t_i) and A1(t_{i+1}) complete; q|k|v rows and staged rows consumed

        // ---- E(t_i) and B(t_{i+1}) on tcgen05 ------------------------------------------------------------------------------------
        if (tid == 0) {
            tc_fence_after_sync();
            const uint32_t a2 = smem_u32(smem + L.a2), wo = smem_u32(smem + L.wo);
            const uint32_t lbo_o = (uint32_t)N2 * 16u, idesc = make_idesc_bf16(128, (uint32_t)N2);
            for (uint32_t ks = 0; ks < (uint32_t)HW >> 4; ks++)
                umma_bf16(tmem_base + D2_COL, make_smem_desc(a2 + ks * 2u * WF_LBO, WF_LBO, WF_SBO),
                          make_smem_desc(wo + ks * 2u * lbo_o, lbo_o, WF_SBO), idesc, ks > 0);
            umma_commit(d2_full);
            if (has_next) issue_mma1();
        }
        // ---- F: D2 + b_o + residual -> fp32 rows, scattered back (window reverse + un-shift as index math) -------------------------
        // destination row and residual (an L2 hit: the gather read it moments ago) are fetched while the MMA runs
        const long long mo = row < nrows ? win_order_token(p.wo, m0row + (uint32_t)row) : 0;
        float4 res[2][4];
#pragma unroll
        for (int gi = 0; gi < 2; gi++) {
            const int c16 = eg * 16 + gi * 32;
#pragma unroll
            for (int i = 0; i < 4; i++)
                res[gi][i] = (p.residual && row < nrows && c16 + i * 4 + 4 <= p.C)
                                 ? *reinterpret_cast<const float4*>(p.residual + mo * p.C + c16 + i * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncwarp();
        mbar_wait_relaxed(d2_full, it & 1u);
        __syncwarp();
        tc_fence_after_sync();
#pragma unroll
        for (int gi = 0; gi < 2; gi++) {
            const int c16 = eg * 16 + gi * 32;
            if (c16 < N2) {   // uniform per warp
                float v[16];
                tmem_ld16(tlane + D2_COL + (uint32_t)c16, v);
                if (row < nrows) {
                    float* o = p.out + mo * p.C + c16;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        if (c16 + i * 4 + 4 <= p.C) {
                            const float4 bb = *reinterpret_cast<const float4*>(sbias + NQKV + c16 + i * 4);
                            const float4 rr = res[gi][i];
                            *reinterpret_cast<float4*>(o + i * 4) = make_float4(v[i * 4] + bb.x + rr.x, v[i * 4 + 1] + bb.y + rr.y,
                                                                              v[i * 4 + 2] + bb.z + rr.z, v[i * 4 + 3] + bb.w + rr.w);
                        }
                    }
                }
            }
        }
        tc_fence_before_sync();
        // ---- C(t_{i+1}): its q|k|v rows replace those of t_i (consumed before the barrier above) ----------------------------------
        if (has_next) {
            mbar_wait_relaxed(d1_full, (it + 1u) & 1u);
            __syncwarp();
            tc_fence_after_sync();
            qkv_rows(tile_rows(tn));
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, ncols);
}

// ---------------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------------
bool wa_fused_supported(const WinGeom& g, int C, int nh, int d) {
    static const bool off = [] { const char* e = getenv("SWINFUSE_WA_FUSED"); return e && e[0] == '0'; }();
    if (off) return false;
    if (!attn_frag_supported(g, nh, d) || nh != 8) return false;
    const int dp = qkvh_dp(d);
    if (!(d <= 3 || (dp == 8 && d < 8))) return false;   // the softmax row sums ride on a spare (ones) column of every v head
    if (C % 4 != 0 || (int)pad16((uint32_t)C) > 64) return false;
    const int Kpad = (int)pad16((uint32_t)C), N2 = Kpad;
    const size_t need = d <= 3 ? wf_layout<true>(Kpad, N2, false, C).total : wf_layout<false>(Kpad, N2, false, C).total;
    return need <= WF_SMEM_LIMIT;
}

template <bool DP4>
static int launch_wa_fused_t(const WaFused& a, cudaStream_t st) {
    const WfSmem L = wf_layout<DP4>(a.Kpad, a.N2, a.self_attn != 0, a.C);
    static DeviceOnce configured;
    if (configured.need()) {
        cudaError_t e = cudaFuncSetAttribute(k_wa_fused<DP4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WF_SMEM_LIMIT);
        if (e != cudaSuccess) { set_error("wa_fused: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SF_ERR_CUDA; }
        configured.done();
    }
    const long long ntiles = ((long long)a.nwin + WF_WIN - 1) / WF_WIN;
    long long grid = 2LL * sm_count();
    if (grid > ntiles) grid = ntiles;
    const double mtok = (double)a.nwin * WF_T;
    const int inner = 8 * a.d;
    // algorithmic work: q|k|v + output projections 8*M*C*inner, QK^T + PV 4*49*M*inner;
    // bytes: each source row read once, each output row written once (fp32), + the residual when it is a third tensor
    const double maps = (a.self_attn ? 2.0 : 3.0) + ((a.residual && a.residual != a.q_src && a.residual != a.kv_src) ? 1.0 : 0.0);
    ProfScope ps(prof_name("wa_fused_c%d", a.C), 8.0 * mtok * a.C * inner + 4.0 * WF_T * mtok * inner, 4.0 * mtok * a.C * maps, st);
    k_wa_fused<DP4><<<(unsigned)grid, WF_THREADS, L.total, st>>>(a);
    SF_CHECK_LAUNCH("wa_fused");
    return SF_OK;
}

int launch_wa_fused(const sf_window_attn_params* p, const WinGeom& g, bool self_attn, const bf16* Wq, const float* bq, const bf16* Wkv,
                    const float* bkv, const bf16* Wo, const float* bo, cudaStream_t st) {
    SF_CHECK_ARG(wa_fused_supported(g, p->C, p->num_heads, p->head_dim), "wa_fused: unsupported shape C=%d heads=%d dim=%d", p->C,
                 p->num_heads, p->head_dim);
    const long long nwin = (long long)g.B * g.nWh * g.nWw;
    SF_CHECK_ARG(nwin * WF_T < 2147483647LL, "wa_fused: %lld windows exceed the index range", nwin);
    WaFused a{};
    a.q_src = p->q_src; a.kv_src = p->kv_src; a.residual = p->residual; a.out = p->out;
    a.ln_q_g = p->ln_q_gamma; a.ln_q_b = p->ln_q_beta; a.ln_kv_g = p->ln_kv_gamma; a.ln_kv_b = p->ln_kv_beta; a.eps = p->ln_eps;
    a.Wq = Wq; a.bq = bq; a.Wkv = Wkv; a.bkv = bkv; a.Wo = Wo; a.bo = bo; a.table = p->bias_table;
    a.wo = make_winorder(g);
    a.nwin = (int)nwin; a.C = p->C; a.d = p->head_dim; a.Kpad = (int)pad16((uint32_t)p->C); a.N2 = a.Kpad; a.self_attn = self_attn ? 1 : 0;
    static const int dbg = [] { const char* e = getenv("SWINFUSE_WF_DEBUG"); return e ? atoi(e) : 0; }();
    a.debug = dbg;
    return p->head_dim <= 3 ? launch_wa_fused_t<true>(a, st) : launch_wa_fused_t<false>(a, st);
}

}  // namespace sf

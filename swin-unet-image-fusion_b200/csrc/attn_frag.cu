// Attention core for 7x7 windows on tensor cores (a001:317-354), fp16 q/k/v rows -> bf16 O.
//
// Per (window, head): S = Q K^T and O = P V are 49x49xd / 49xdx49 products -- far too small for a
// 128-row tcgen05 tile (one UMMA would be >60% padding and the accumulator would have to round-trip
// through TMEM for the softmax), so they run as warp-level m16n8k16 HMMA tiles whose accumulator
// fragments stay in registers, where the softmax is applied directly (FlashAttention-2 style).
//
// There is no shared memory and no block-level synchronisation in this kernel.  The q/k/v projection
// GEMM (tc_gemm.cu) runs on rows in WINDOW ORDER (row = window*49 + token; its A producers apply the
// cyclic shift and the window partition as index math) and writes fp16 rows
//     [ q heads | k heads | v heads ],  every head padded to dp columns,  q pre-multiplied by d^-1/2 log2 e
// so a window is 49 consecutive rows and every operand fragment is a direct global load (L1-resident:
// one CTA works through all heads and row slabs of a window before moving on):
//
//   CTA  = one window at a time (persistent);  warp = (slab of 16 query rows, head parity)
//   warp task = (window, head, slab)
//       S[16 x 56] = Q_slab K^T        7 n-tiles x KSTEPS HMMAs, accumulators start at bias * log2 e
//                                      (the slab's bias fragment lives in registers for the whole kernel:
//                                       one 13x13 table shared by all heads and windows, a001:113-144)
//       shift mask                     boundary windows only: two 28-bit key patterns per lane (a001:222-315)
//       row max / row sum              2 quad shuffles each;  p = ex2(s - max)
//       O[16 x dp] = P V               P (fp16) re-used in place as the A fragments; V fragments are loaded
//                                      row-major and transposed in registers (movmatrix)
//
// Dot products do not care in which order k is summed, so the k index of the S MMAs is permuted to
// make the loads wide: lane (gq, tq) feeds k-slots {2tq, 2tq+1} and {2tq+8, 2tq+9} from the 4
// consecutive head dims 4tq..4tq+3 of a row (one LDG.64 for both registers, A and B alike).
// O is written bf16 in the UMMA-tiled layout, window order, same head padding (the padded dims are
// exact zeros: they come from zero V columns), which is the A operand the projection GEMM bulk-copies.
#include <cuda_fp16.h>
#include "bf16_kernels.cuh"
#include "tc_common.cuh"
#include "hmma_util.cuh"

namespace sf {

static constexpr int FT = 49;          // tokens per window
static constexpr int AF_THREADS = 256; // 8 warps: slab = warp % 4, head parity = warp / 4

struct AttnFragArgs {
    const __half* qkv; int ld;        // fp16 rows, ld = 3 * hw
    bf16* O; int o_nkc;               // hw / 8 k-chunks per 128-row tile
    const float* table;
    int nh, d, hw;                    // hw = nh * dp
    int nWh, nWw, shift;
    int nwin;
    FastDiv dnW, dnWw;
};

// operand fragments of one (window, head) for one slab, as loaded (V still row-major)
template <int KSTEPS, int NDT>
struct Frags {
    uint2 q[KSTEPS][2];     // rows r0 / r1: .x -> k-slots {2tq,2tq+1}, .y -> {2tq+8,2tq+9}
    uint2 k[KSTEPS][7];     // n-tile nt: key nt*8 + gq
    uint32_t v[NDT][7];     // d-tile dt, key tile kt: V[key kt*8 + gq][dims dt*8 + 2tq, +1]
};

// Per-lane pointers into the window-0 / head-0 rows (q row r0, k row gq, v row gq, lane column offsets
// folded in); every other operand address is one of these plus a warp-uniform (window, head) offset
// plus a compile-time immediate -- the row pitch LD is a template constant.  Loads are unconditional:
// lanes whose row does not exist (query rows / keys >= 49) read another, existing row of the window
// instead (finite values that never reach the output: such keys carry bias -1e30 -> probability 0,
// such query rows are not stored), which saves the predicate logic and the zero fill of the registers.
struct LanePtrs {
    const __half* q; const __half* k; const __half* v;
    int q1off;                 // element offset of row r1 from row r0 (0 when r1 does not exist)
    int t6off;                 // element offset of key tile 6 (keys 48..55): 48 rows for the lane group of key 48, else 0
    bool r0ok, r1ok;           // rows r0 / r1 < 49
};

template <int KSTEPS, int NDT, bool DP4, int LD>
__device__ __forceinline__ void load_frags(Frags<KSTEPS, NDT>& f, const LanePtrs& lp, size_t uoff, int tq) {
    constexpr int dp = DP4 ? 4 : 8 * NDT;
    const __half* q = lp.q + uoff;
    const __half* k = lp.k + uoff;
    const __half* v = lp.v + uoff;
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ks++) {
        f.q[ks][0] = ldg64(q + ks * 16);
        f.q[ks][1] = ldg64(q + lp.q1off + ks * 16);
        // k-slots beyond the head's dp columns hold the next head's values: zero them on the Q side
        if (ks * 16 + 16 > dp && ks * 16 + 4 * tq >= dp) { f.q[ks][0] = make_uint2(0u, 0u); f.q[ks][1] = make_uint2(0u, 0u); }
#pragma unroll
        for (int nt = 0; nt < 7; nt++) f.k[ks][nt] = ldg64(k + (nt < 6 ? nt * 8 * LD : lp.t6off) + ks * 16);
    }
#pragma unroll
    for (int dt = 0; dt < NDT; dt++) {
#pragma unroll
        for (int kt = 0; kt < 7; kt++) f.v[dt][kt] = ldg32(v + (kt < 6 ? kt * 8 * LD : lp.t6off) + dt * 8);
    }
}

// one warp task: S = Q K^T + bias, mask, softmax on the fragments, O = P V, store
template <int KSTEPS, int NDT, bool DP4, int NH, bool ONES>
__device__ __forceinline__ void attn_task(const Frags<KSTEPS, NDT>& cur, const float (&bias)[7][4], uint32_t m0, uint32_t m1,
                                          const AttnFragArgs& a, const LanePtrs& lp, int win, int head, int r0, int tq, int lsrc) {
    constexpr int dp = DP4 ? 4 : 8 * NDT;
    constexpr int ONKC = NH * dp / 8;
    constexpr float MASKED = -1.4426950e10f;   // the reference overwrites masked scores with -1e10 (a001:310), log2 domain
    float s[7][4];
#pragma unroll
    for (int nt = 0; nt < 7; nt++) { s[nt][0] = bias[nt][0]; s[nt][1] = bias[nt][1]; s[nt][2] = bias[nt][2]; s[nt][3] = bias[nt][3]; }
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ks++)
#pragma unroll
        for (int nt = 0; nt < 7; nt++)
            mma16816(s[nt], cur.q[ks][0].x, cur.q[ks][1].x, cur.q[ks][0].y, cur.q[ks][1].y, cur.k[ks][nt].x, cur.k[ks][nt].y);
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
    // ---- softmax on the fragments ------------------------------------------------------------------------
    float x0 = fmaxf(s[0][0], s[0][1]), x1 = fmaxf(s[0][2], s[0][3]);
#pragma unroll
    for (int nt = 1; nt < 7; nt++) {
        x0 = max3f(x0, s[nt][0], s[nt][1]);
        x1 = max3f(x1, s[nt][2], s[nt][3]);
    }
    x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, 1)); x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, 2));
    x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, 1)); x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, 2));
    float l0 = 0.f, l1 = 0.f;
    uint32_t pf[7][2];   // P as fp16 pairs: [nt][0] = row r0, [nt][1] = row r1
#pragma unroll
    for (int nt = 0; nt < 7; nt++) {
        // n-tile 6 holds keys 48..55: only key 48 (column 0, lanes tq == 0) is real; the padded ones give ex2(-1e30) = 0
        const float p0 = ex2f(s[nt][0] - x0), p2 = ex2f(s[nt][2] - x1);
        const float p1 = nt < 6 ? ex2f(s[nt][1] - x0) : 0.f, p3 = nt < 6 ? ex2f(s[nt][3] - x1) : 0.f;
        if (!ONES) { l0 += p0 + p1; l1 += p2 + p3; }
        pf[nt][0] = pack_h2(p0, p1);
        pf[nt][1] = pack_h2(p2, p3);
    }
    // ---- O = P V : the score fragments of n-tiles (2j, 2j+1) are the A fragment of key step j ------------------
    float o[NDT][4];
#pragma unroll
    for (int dt = 0; dt < NDT; dt++) { o[dt][0] = 0.f; o[dt][1] = 0.f; o[dt][2] = 0.f; o[dt][3] = 0.f; }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t a0 = pf[2 * j][0], a1 = pf[2 * j][1];
        const uint32_t a2 = (2 * j + 1 < 7) ? pf[2 * j + 1][0] : 0u, a3 = (2 * j + 1 < 7) ? pf[2 * j + 1][1] : 0u;
#pragma unroll
        for (int dt = 0; dt < NDT; dt++) {
            const uint32_t b0 = movm_trans(cur.v[dt][2 * j]);
            const uint32_t b1 = (2 * j + 1 < 7) ? movm_trans(cur.v[dt][2 * j + 1]) : 0u;
            mma16816(o[dt], a0, a1, a2, a3, b0, b1);
        }
    }
    if (ONES) {   // row sums = column d of P V: last d-tile, quad lane (d%8)/2, element d%2
        const float c0 = (a.d & 1) ? o[NDT - 1][1] : o[NDT - 1][0], c1 = (a.d & 1) ? o[NDT - 1][3] : o[NDT - 1][2];
        l0 = __shfl_sync(0xffffffffu, c0, lsrc);
        l1 = __shfl_sync(0xffffffffu, c1, lsrc);
    } else {
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    }
    // ---- O rows -> global, UMMA-tiled in window order: row m = win*49 + r, column head*dp + dim ---------------------
    // element (m, col) at ((m >> 7) * ONKC + (col >> 3)) * 1024 + (m & 127) * 8 + (col & 7)
    const float i0 = rcpf(l0), i1 = rcpf(l1);
    if (!DP4 || tq < 2) {
        const uint32_t mrow0 = (uint32_t)win * FT + (uint32_t)r0;
#pragma unroll
        for (int half = 0; half < 2; half++) {
            if (half ? lp.r1ok : lp.r0ok) {
                const uint32_t m = mrow0 + 8 * half;
                const float inv = half ? i1 : i0;
#pragma unroll
                for (int dt = 0; dt < NDT; dt++) {
                    const int col = head * dp + dt * 8 + 2 * tq;
                    const uint32_t off = ((m >> 7) * ONKC + (uint32_t)(col >> 3)) * 1024u + (m & 127u) * 8u + (uint32_t)(col & 7);
                    *reinterpret_cast<uint32_t*>(a.O + off) = tc::pack_bf16x2(o[dt][2 * half] * inv, o[dt][2 * half + 1] * inv);
                }
            }
        }
    }
}

template <int KSTEPS, int NDT, bool DP4, int NH, bool PREFETCH, bool ONES>
__global__ void __launch_bounds__(AF_THREADS, PREFETCH ? 2 : 1) k_attn_frag(AttnFragArgs a) {
    constexpr int dp = DP4 ? 4 : 8 * NDT;
    constexpr int HW = NH * dp, LD = 3 * HW;
    constexpr int HPW = NH / 2;                // heads per warp and window (head parity = warp / 4)
    static_assert(NH % 2 == 0, "two head streams per CTA");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gq = lane >> 2, tq = lane & 3;   // fragment coordinates: row group / column pair
    const int slab = warp & 3, sub = warp >> 2;
    const int r0 = slab * 16 + gq, r1 = r0 + 8;
    constexpr float LOG2E = 1.4426950408889634f;
    constexpr float NEG = -1e30f;

    // ---- once per warp: bias fragment of this slab (x log2e; padded keys = -inf) and the key patterns of the shift mask
    float bias[7][4];
    uint32_t kh = 0, kw = 0;   // bit 2nt+e: key nt*8 + 2tq + e lies in the upper part (>= 4) of the window along H / W
#pragma unroll
    for (int nt = 0; nt < 7; nt++) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int key = nt * 8 + 2 * tq + e;
            float b0 = NEG, b1 = NEG;
            if (key < FT) {
                const int kr = key / 7, kc = key - kr * 7;
                b0 = r0 < FT ? LOG2E * __ldg(a.table + (kr - r0 / 7 + 6) * 13 + (kc - r0 % 7 + 6)) : 0.f;
                b1 = r1 < FT ? LOG2E * __ldg(a.table + (kr - r1 / 7 + 6) * 13 + (kc - r1 % 7 + 6)) : 0.f;
                if (kr >= 4) kh |= 1u << (2 * nt + e);
                if (kc >= 4) kw |= 1u << (2 * nt + e);
            }
            bias[nt][e] = b0;
            bias[nt][2 + e] = b1;
        }
    }
    // masked-slot patterns of this lane's two rows for a window in the last window row / column (a001:222-272:
    // tokens 0..3 and 4..6 of such a window lie in different regions along that axis)
    const uint32_t mh0 = (r0 < FT && r0 / 7 >= 4) ? ~kh : kh, mw0 = (r0 < FT && r0 % 7 >= 4) ? ~kw : kw;
    const uint32_t mh1 = (r1 < FT && r1 / 7 >= 4) ? ~kh : kh, mw1 = (r1 < FT && r1 % 7 >= 4) ? ~kw : kw;

    LanePtrs lp;
    lp.r0ok = r0 < FT; lp.r1ok = r1 < FT;
    const int r0c = lp.r0ok ? r0 : 48, r1c = lp.r1ok ? r1 : r0c;   // clamped to existing rows
    lp.q = a.qkv + r0c * LD + 4 * tq;
    lp.q1off = (r1c - r0c) * LD;
    lp.k = a.qkv + gq * LD + HW + 4 * tq;
    lp.v = a.qkv + gq * LD + 2 * HW + 2 * tq;
    lp.t6off = gq == 0 ? 48 * LD : 0;
    const int lsrc = (lane & ~3) | ((a.d & 7) >> 1);   // quad lane that holds column d of the last d-tile (ONES)

    // the warp's tasks: windows blockIdx.x, +gridDim.x, ...; per window the heads sub, sub+2, ... (HPW of them).
    // Two fragment sets alternate (loads of task i+1 are in flight while task i is computed), the loop is
    // unrolled by two so that no register copies are needed.
    Frags<KSTEPS, NDT> fa, fb;
    int win = blockIdx.x;
    if (win >= a.nwin) return;
    load_frags<KSTEPS, NDT, DP4, LD>(fa, lp, (size_t)win * (FT * LD) + sub * dp, tq);
    for (; win < a.nwin; win += gridDim.x) {
        uint32_t m0 = 0, m1 = 0;
        if (a.shift) {   // boundary windows of the shifted frame are the only ones whose tokens span several regions
            const uint32_t wi = (uint32_t)win - fdiv((uint32_t)win, a.dnW) * (uint32_t)(a.nWh * a.nWw);
            const uint32_t wh = fdiv(wi, a.dnWw), ww = wi - wh * (uint32_t)a.nWw;
            if (wh == (uint32_t)a.nWh - 1) { m0 |= mh0; m1 |= mh1; }
            if (ww == (uint32_t)a.nWw - 1) { m0 |= mw0; m1 |= mw1; }
        }
        const size_t wbase = (size_t)win * (FT * LD);
        const int wnext = win + gridDim.x;
#pragma unroll
        for (int i = 0; i < HPW; i += 2) {
            const int h0 = sub + 2 * i, h1 = h0 + 2;
            // task (win, h0) from fa; its successor (win, h1), or the next window's first head, goes to fb
            if (i + 1 < HPW) load_frags<KSTEPS, NDT, DP4, LD>(fb, lp, wbase + h1 * dp, tq);
            else if (wnext < a.nwin) load_frags<KSTEPS, NDT, DP4, LD>(fb, lp, (size_t)wnext * (FT * LD) + sub * dp, tq);
            attn_task<KSTEPS, NDT, DP4, NH, ONES>(fa, bias, m0, m1, a, lp, win, h0, r0, tq, lsrc);
            if (i + 1 < HPW) {
                if (i + 2 < HPW) load_frags<KSTEPS, NDT, DP4, LD>(fa, lp, wbase + (h1 + 2) * dp, tq);
                else if (wnext < a.nwin) load_frags<KSTEPS, NDT, DP4, LD>(fa, lp, (size_t)wnext * (FT * LD) + sub * dp, tq);
                attn_task<KSTEPS, NDT, DP4, NH, ONES>(fb, bias, m0, m1, a, lp, win, h1, r0, tq, lsrc);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// d <= 3 (head = 4 padded columns = 8 bytes): four heads per warp pass.
//
// With 8-byte heads a per-head fragment load touches eight 128-byte lines for 64 useful bytes and the
// L1 tag stage becomes the limiter.  Here one warp = (slab, head parity e) works through the four
// heads 2j+e of a window with ONE set of loads: lane (gq, tq) loads the 8 bytes of head 2tq+e of its
// row -- q rows r0/r1, k and v rows gq of every key tile -- so a warp-wide load covers four heads.
//   S_h   : head h = 2j+e lives in the k-slots fed by lanes tq == j, so the A fragment is the q
//           registers of those lanes and zero elsewhere (a SEL); the K registers are used as they are
//           (the other lanes' k-slots multiply zeros).  Nothing moves between lanes.
//   P V   : movmatrix of the v registers gives B fragments whose column n = 2j+c is dim c (.x: 0,1 /
//           .y: 2,3) of head 2j+e -- the same two B fragments serve all four heads; head h's result
//           and its row sum (ones column, dim d) land in the accumulators of lanes tq == j, which store.
// ---------------------------------------------------------------------------------------------
template <int NH>
__global__ void __launch_bounds__(AF_THREADS, 2) k_attn_pack4(AttnFragArgs a) {
    constexpr int dp = 4, HW = NH * dp, LD = 3 * HW, ONKC = HW / 8;
    constexpr int HPW = NH / 2;
    static_assert(HPW == 4, "one head per lane of a quad");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gq = lane >> 2, tq = lane & 3;
    const int slab = warp & 3, par = warp >> 2;
    const int r0 = slab * 16 + gq, r1 = r0 + 8;
    const int hj = 2 * tq + par;               // the head this lane loads and stores
    constexpr float LOG2E = 1.4426950408889634f;
    constexpr float NEG = -1e30f;
    constexpr float MASKED = -1.4426950e10f;

    float bias[7][4];
    uint32_t kh = 0, kw = 0;
#pragma unroll
    for (int nt = 0; nt < 7; nt++) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int key = nt * 8 + 2 * tq + e;
            float b0 = NEG, b1 = NEG;
            if (key < FT) {
                const int kr = key / 7, kc = key - kr * 7;
                b0 = r0 < FT ? LOG2E * __ldg(a.table + (kr - r0 / 7 + 6) * 13 + (kc - r0 % 7 + 6)) : 0.f;
                b1 = r1 < FT ? LOG2E * __ldg(a.table + (kr - r1 / 7 + 6) * 13 + (kc - r1 % 7 + 6)) : 0.f;
                if (kr >= 4) kh |= 1u << (2 * nt + e);
                if (kc >= 4) kw |= 1u << (2 * nt + e);
            }
            bias[nt][e] = b0;
            bias[nt][2 + e] = b1;
        }
    }
    const uint32_t mh0 = (r0 < FT && r0 / 7 >= 4) ? ~kh : kh, mw0 = (r0 < FT && r0 % 7 >= 4) ? ~kw : kw;
    const uint32_t mh1 = (r1 < FT && r1 / 7 >= 4) ? ~kh : kh, mw1 = (r1 < FT && r1 % 7 >= 4) ? ~kw : kw;

    const bool r0ok = r0 < FT, r1ok = r1 < FT;
    const int r0c = r0ok ? r0 : 48, r1c = r1ok ? r1 : r0c;   // rows that do not exist read an existing one (never stored)
    const __half* pq = a.qkv + r0c * LD + hj * dp;
    const int q1off = (r1c - r0c) * LD;
    const __half* pk = a.qkv + gq * LD + HW + hj * dp;
    const __half* pv = a.qkv + gq * LD + 2 * HW + hj * dp;
    const int t6off = gq == 0 ? 48 * LD : 0;   // key tile 6: key 48 for its lane group, any existing row for the others (bias -1e30)
    const int ocol = hj * dp;
    const uint32_t ocoff = (uint32_t)(ocol >> 3) * 1024u + (uint32_t)(ocol & 7);

    int win = blockIdx.x;
    if (win >= a.nwin) return;
    uint2 q[2], k[7], vraw[7];
    {
        const size_t wb = (size_t)win * (FT * LD);
        q[0] = ldg64(pq + wb); q[1] = ldg64(pq + wb + q1off);
#pragma unroll
        for (int nt = 0; nt < 7; nt++) {
            k[nt] = ldg64(pk + wb + (nt < 6 ? nt * 8 * LD : t6off));
            vraw[nt] = ldg64(pv + wb + (nt < 6 ? nt * 8 * LD : t6off));
        }
    }
    for (; win < a.nwin; win += gridDim.x) {
        const int wnext = win + gridDim.x;
        const bool more = wnext < a.nwin;
        const size_t wbn = (size_t)(more ? wnext : win) * (FT * LD);
        // B fragments of P V, shared by the four heads: [key tile][dims 0,1 / 2,3]
        uint32_t vt[7][2];
#pragma unroll
        for (int kt = 0; kt < 7; kt++) { vt[kt][0] = movm_trans(vraw[kt].x); vt[kt][1] = movm_trans(vraw[kt].y); }
        if (more) {
#pragma unroll
            for (int kt = 0; kt < 7; kt++) vraw[kt] = ldg64(pv + wbn + (kt < 6 ? kt * 8 * LD : t6off));
        }
        uint32_t m0 = 0, m1 = 0;
        if (a.shift) {
            const uint32_t wi = (uint32_t)win - fdiv((uint32_t)win, a.dnW) * (uint32_t)(a.nWh * a.nWw);
            const uint32_t wh = fdiv(wi, a.dnWw), ww = wi - wh * (uint32_t)a.nWw;
            if (wh == (uint32_t)a.nWh - 1) { m0 |= mh0; m1 |= mh1; }
            if (ww == (uint32_t)a.nWw - 1) { m0 |= mw0; m1 |= mw1; }
        }
        const uint32_t mrow0 = (uint32_t)win * FT + (uint32_t)r0, mrow1 = mrow0 + 8;
        const uint32_t off0 = (mrow0 >> 7) * (ONKC * 1024u) + (mrow0 & 127u) * 8u + ocoff;
        const uint32_t off1 = (mrow1 >> 7) * (ONKC * 1024u) + (mrow1 & 127u) * 8u + ocoff;
#pragma unroll
        for (int j = 0; j < HPW; j++) {
            const bool mine = tq == j;
            const uint32_t a0 = mine ? q[0].x : 0u, a1 = mine ? q[1].x : 0u, a2 = mine ? q[0].y : 0u, a3 = mine ? q[1].y : 0u;
            float s[7][4];
#pragma unroll
            for (int nt = 0; nt < 7; nt++) { s[nt][0] = bias[nt][0]; s[nt][1] = bias[nt][1]; s[nt][2] = bias[nt][2]; s[nt][3] = bias[nt][3]; }
#pragma unroll
            for (int nt = 0; nt < 7; nt++) mma16816(s[nt], a0, a1, a2, a3, k[nt].x, k[nt].y);
            if (j == HPW - 1 && more) {   // q and k are dead: fetch the next window's
                q[0] = ldg64(pq + wbn); q[1] = ldg64(pq + wbn + q1off);
#pragma unroll
                for (int nt = 0; nt < 7; nt++) k[nt] = ldg64(pk + wbn + (nt < 6 ? nt * 8 * LD : t6off));
            }
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
            float ox[4] = {0.f, 0.f, 0.f, 0.f}, oy[4] = {0.f, 0.f, 0.f, 0.f};   // dims 0,1 / 2,3 of head 2*(col/2)+par; rows r0 | r1
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
                const uint32_t p0 = pf[2 * jj][0], p1 = pf[2 * jj][1];
                const uint32_t p2 = (2 * jj + 1 < 7) ? pf[2 * jj + 1][0] : 0u, p3 = (2 * jj + 1 < 7) ? pf[2 * jj + 1][1] : 0u;
                mma16816(ox, p0, p1, p2, p3, vt[2 * jj][0], (2 * jj + 1 < 7) ? vt[2 * jj + 1][0] : 0u);
                mma16816(oy, p0, p1, p2, p3, vt[2 * jj][1], (2 * jj + 1 < 7) ? vt[2 * jj + 1][1] : 0u);
            }
            if (mine) {   // this lane's accumulator columns are its head: dims (ox[0],ox[1],oy[0],oy[1]) of row r0, [2],[3] of row r1
                const float l0 = a.d == 3 ? oy[1] : (a.d == 2 ? oy[0] : ox[1]);
                const float l1 = a.d == 3 ? oy[3] : (a.d == 2 ? oy[2] : ox[3]);
                const float i0 = rcpf(l0), i1 = rcpf(l1);
                if (r0ok) *reinterpret_cast<uint2*>(a.O + off0) = make_uint2(tc::pack_bf16x2(ox[0] * i0, ox[1] * i0), tc::pack_bf16x2(oy[0] * i0, oy[1] * i0));
                if (r1ok) *reinterpret_cast<uint2*>(a.O + off1) = make_uint2(tc::pack_bf16x2(ox[2] * i1, ox[3] * i1), tc::pack_bf16x2(oy[2] * i1, oy[3] * i1));
            }
        }
    }
}

template <int KSTEPS, int NDT, bool DP4, bool PREFETCH>
static int launch_attn_frag_t(const AttnFragArgs& a, cudaStream_t st) {
    constexpr int dp = DP4 ? 4 : 8 * NDT;
    long long grid = (long long)sm_count() * (PREFETCH ? 2 : 1);
    if (grid > a.nwin) grid = a.nwin;
    const int inner = a.nh * a.d;
    const double mtok = (double)a.nwin * FT;
    // algorithmic work: QK^T + PV = 4*49*N*C FLOP; bytes = q,k,v fp16 in + O bf16 out
    ProfScope ps(prof_name("attn_core_frag_c%d", inner), 4.0 * FT * mtok * inner, 8.0 * mtok * inner, st);
    if (a.d < dp) k_attn_frag<KSTEPS, NDT, DP4, 8, PREFETCH, true><<<(unsigned)grid, AF_THREADS, 0, st>>>(a);
    else k_attn_frag<KSTEPS, NDT, DP4, 8, PREFETCH, false><<<(unsigned)grid, AF_THREADS, 0, st>>>(a);
    SF_CHECK_LAUNCH("attn_core_frag");
    return SF_OK;
}

bool attn_frag_supported(const WinGeom& g, int nh, int d) {
    return g.T == FT && g.wsh == 7 && g.wsw == 7 && d >= 1 && d <= 48 && nh == 8;   // the row pitch is a compile-time constant
}

int launch_attn_frag(const __half* qkv, int ld, bf16* O, const float* table, const WinGeom& g, int nh, int d, cudaStream_t st) {
    if (!attn_frag_supported(g, nh, d)) { set_error("attention core: unsupported window / head shape"); return SF_ERR_UNSUPPORTED; }
    AttnFragArgs a{};
    const int dp = qkvh_dp(d);
    a.qkv = qkv; a.ld = ld; a.O = O; a.hw = nh * dp; a.o_nkc = a.hw / 8; a.table = table;
    a.nh = nh; a.d = d; a.nWh = g.nWh; a.nWw = g.nWw; a.shift = g.shift;
    const long long nwin = (long long)g.B * g.nWh * g.nWw;
    SF_CHECK_ARG(ld == 3 * a.hw, "attention core: q|k|v rows must be %d columns wide", 3 * a.hw);
    SF_CHECK_ARG((nwin * FT + 127) / 128 * 128 * (long long)ld < 2147483647LL, "attention core: %lld windows exceed the index range", nwin);
    a.nwin = (int)nwin;
    a.dnW = make_fastdiv((uint32_t)(g.nWh * g.nWw));
    a.dnWw = make_fastdiv((uint32_t)g.nWw);
    if (d <= 3) {   // 8-byte heads: four heads per warp pass
        long long grid = 296;
        if (grid > a.nwin) grid = a.nwin;
        const double mtok = (double)a.nwin * FT;
        ProfScope ps(prof_name("attn_core_frag_c%d", nh * d), 4.0 * FT * mtok * nh * d, 8.0 * mtok * nh * d, st);
        k_attn_pack4<8><<<(unsigned)grid, AF_THREADS, 0, st>>>(a);
        SF_CHECK_LAUNCH("attn_core_pack4");
        return SF_OK;
    }
    if (d <= 4) return launch_attn_frag_t<1, 1, true, true>(a, st);
    if (d <= 8) return launch_attn_frag_t<1, 1, false, true>(a, st);
    if (d <= 16) return launch_attn_frag_t<1, 2, false, true>(a, st);
    if (d <= 24) return launch_attn_frag_t<2, 3, false, false>(a, st);
    if (d <= 32) return launch_attn_frag_t<2, 4, false, false>(a, st);
    if (d <= 40) return launch_attn_frag_t<3, 5, false, false>(a, st);
    return launch_attn_frag_t<3, 6, false, false>(a, st);
}

}  // namespace sf

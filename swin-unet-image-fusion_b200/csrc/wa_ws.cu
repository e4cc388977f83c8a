// Warp-specialised fused window attention for 8-byte heads (d <= 3, C <= 32: stage 0 of the model, the longest operator of
// the forward pass).  Same mathematics and the same data path as wa_fused.cu -- gather + LayerNorm -> q|k|v on tcgen05 ->
// HMMA attention core on q|k (fp16) / v (bf16) rows in shared memory -> projection on tcgen05 -> + bias + residual, scattered back
// (a001:448-474 + a004:29-38) -- but the two kinds of work no longer take turns:
//
//   warps 0-3   helpers    everything that is latency: bulk-copy gather of the next tiles' rows (two tiles ahead, one copy per
//                          window row; a tile stays staged until its scatter has taken the residual: x is read once),
//                          LayerNorm -> A1, tcgen05.mma issue, TMEM -> fp16 q|k|v rows, TMEM -> + b_o + residual -> HBM.
//                          They run one tile ahead of the attention warps and one to three tiles behind them
//                          (iteration j: C(j-1) | A(j), barrier, B(j) | E(j-2) | F(j-3) | gather of tile j+2 -- the rows the
//                          attention warps wait for come first, the gather's address arithmetic last), every buffer between
//                          the two groups (q|k rows fp16 / v rows bf16, A1, A2, both TMEM accumulators) is double buffered.
//   warps 4-15  attention  the softmax-bound core and nothing else: warp = (window of the tile, 16-row slab 0..2,
//                          head parity), four heads per pass (wf_attn_pack4).  They never wait for HBM, the tensor
//                          core or TMEM -- only for the q|k|v rows of the next tile, which the helpers finish while the
//                          current tile is being computed.
//
// The 49th token of a window no longer costs a quarter of the core: rows 48..63 of the fourth slab held ONE real row.
// Instead, one warp per window runs a "token 48" task in which the 16 MMA rows are the 8 HEADS of that token:
// S[h][key] = sum_(h',dd) A[h][(h',dd)] K[key][(h',dd)] with A block-diagonal (A[h][(h',dd)] = q48[h][dd] iff h' == h), and
// O[h][(h',dd)] = P[h] V, of which the diagonal blocks h' == h are kept.  The task rotates over the six warps of a window.
#include <cstdlib>
#include "wa_common.cuh"

namespace sf {

static constexpr int WS_HELPERS = 4;                 // helper warps (TMEM lane quarter = warp id % 4)
static constexpr int WS_ATT = 12;                    // attention warps: 2 windows x 3 slabs x 2 head parities
static constexpr int WS_THREADS = (WS_HELPERS + WS_ATT) * 32;
static constexpr int WS_HT = WS_HELPERS * 32;
static constexpr int WS_PD = 2;                      // tiles the gather runs ahead of the LayerNorm stage
static constexpr int WS_RS = WS_PD + 4;              // staging slots: a tile stays until its scatter stage (3 iterations later) took the residual
static constexpr size_t WS_SMEM_LIMIT = 227 * 1024;
static constexpr int WS_HW = 32, WS_NQKV = 96;
static constexpr uint32_t WS_PITCH = 208;            // bytes per fp16 q|k|v row
static constexpr int WS_PH = WS_PITCH / 2;
static constexpr uint32_t WS_D1_STRIDE = 128, WS_D2_COL = 256, WS_D2_STRIDE = 32, WS_TMEM_COLS = 512;

struct WsSmem { uint32_t wq, wkv, wo, bias, b48, a1q[2], a1kv[2], a2[2], qkv[2], raw, raw_stride, bars, total; };

__host__ __device__ static inline WsSmem ws_layout(int Kpad, int N2, bool self_attn, int C) {
    WsSmem s{};
    uint32_t o = 0;
    const uint32_t kc = (uint32_t)Kpad >> 3;
    s.wq = o;   o += wf_al(kc * (self_attn ? (uint32_t)WS_NQKV : (uint32_t)WS_HW) * 16u);
    s.wkv = o;  o += wf_al(self_attn ? 0u : kc * 2u * WS_HW * 16u);
    s.wo = o;   o += wf_al((WS_HW >> 3) * (uint32_t)N2 * 16u);
    s.bias = o; o += wf_al((WS_NQKV + (uint32_t)N2) * 4u);
    s.b48 = o;  o += wf_al(56u * 4u);
    for (int i = 0; i < 2; i++) { s.a1q[i] = o; o += wf_al(kc * WF_LBO); }
    for (int i = 0; i < 2; i++) { s.a1kv[i] = o; o += wf_al(self_attn ? 0u : kc * WF_LBO); }
    for (int i = 0; i < 2; i++) { s.a2[i] = o; o += wf_al((WS_HW >> 3) * WF_LBO); }
    for (int i = 0; i < 2; i++) { s.qkv[i] = o; o += wf_al((uint32_t)WF_ROWS * WS_PITCH); }
    s.raw_stride = (self_attn ? 1u : 2u) * wf_al((uint32_t)WF_ROWS * (uint32_t)C * 4u);
    s.raw = o;  o += WS_RS * s.raw_stride;
    s.bars = o; o += 256;
    s.total = o;
    return s;
}

__device__ __forceinline__ void ws_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ws_helper_bar() { asm volatile("bar.sync 1, %0;" ::"n"(WS_HT) : "memory"); }

// ---- token 48 of a window, all 8 heads in one warp task (see the header comment) ------------------------------------------
// lane (gq, tq): MMA row gq = head gq.  k-slot permutation as in wf_attn_pack4: lane tq feeds k-slots {2tq, 2tq+1, 2tq+8,
// 2tq+9} of k-step s from the four columns 16s + 4tq .. +3 = the four dims of head 4s + tq.
__device__ __forceinline__ void ws_attn_tok48(const __half* __restrict__ wbase, uint8_t* __restrict__ sA2, int rowbase,
                                              const float* __restrict__ sb48, bool mask_h, bool mask_w, int d, int gq, int tq, int lane) {
    constexpr int PH = WS_PH, HW = WS_HW;
    constexpr float MASKED = WF_MASKED;
    const uint2 q = lds64(wbase + 48 * PH + gq * 4);   // q of token 48, head gq (pre-scaled by d^-1/2 log2 e)
    const int t6row = gq == 0 ? 48 : gq;               // key tile 6: key 48 for its lane group, an existing row elsewhere (bias -1e30)
    float s[7][4];
#pragma unroll
    for (int nt = 0; nt < 7; nt++) {
        const float2 b = *reinterpret_cast<const float2*>(sb48 + nt * 8 + 2 * tq);
        s[nt][0] = b.x; s[nt][1] = b.y; s[nt][2] = 0.f; s[nt][3] = 0.f;
    }
#pragma unroll
    for (int ks = 0; ks < 2; ks++) {
        const bool mine = gq == 4 * ks + tq;
        const uint32_t a0 = mine ? q.x : 0u, a2 = mine ? q.y : 0u;
#pragma unroll
        for (int nt = 0; nt < 7; nt++) {
            const uint2 kk = lds64(wbase + (nt < 6 ? nt * 8 + gq : t6row) * PH + HW + 16 * ks + 4 * tq);
            mma16816(s[nt], a0, 0u, a2, 0u, kk.x, kk.y);
        }
    }
    if (mask_h | mask_w) {   // token 48 = (row 6, column 6) of the window: in the upper part along both axes
#pragma unroll
        for (int nt = 0; nt < 7; nt++) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int key = nt * 8 + 2 * tq + e;
                const int kr = key / 7, kc = key - kr * 7;
                if (key < WF_T && ((mask_h && kr < 4) || (mask_w && kc < 4))) s[nt][e] = MASKED;
            }
        }
    }
    float x0 = fmaxf(s[0][0], s[0][1]);
#pragma unroll
    for (int nt = 1; nt < 7; nt++) x0 = max3f(x0, s[nt][0], s[nt][1]);
    x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, 1)); x0 = fmaxf(x0, __shfl_xor_sync(0xffffffffu, x0, 2));
    uint32_t pf[7];
#pragma unroll
    for (int nt = 0; nt < 7; nt++) {
        const float p0 = ex2f(s[nt][0] - x0);
        const float p1 = nt < 6 ? ex2f(s[nt][1] - x0) : 0.f;
        pf[nt] = pack_bf16x2(p0, p1);   // v rows are bf16 (phase C)
    }
    // O[h][(h', dd)] = P[h] V: n-tile c holds heads 2c, 2c+1; V fragments are row-major loads transposed in registers
    float o[4][4];
#pragma unroll
    for (int c = 0; c < 4; c++) { o[c][0] = 0.f; o[c][1] = 0.f; o[c][2] = 0.f; o[c][3] = 0.f; }
#pragma unroll
    for (int jj = 0; jj < 4; jj++) {
        const uint32_t a0 = pf[2 * jj], a2 = (2 * jj + 1 < 7) ? pf[2 * jj + 1] : 0u;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            // key tile 6 (jj == 3, first half) holds key 48 only: the other lane groups read an existing row (their P is 0)
            const uint32_t b0 = movm_trans(lds32(wbase + (jj < 3 ? 2 * jj * 8 + gq : t6row) * PH + 2 * HW + 8 * c + 2 * tq));
            const uint32_t b1 = jj < 3 ? movm_trans(lds32(wbase + ((2 * jj + 1) * 8 + gq) * PH + 2 * HW + 8 * c + 2 * tq)) : 0u;
            mma16816_bf16(o[c], a0, 0u, a2, 0u, b0, b1);
        }
    }
    // head gq lives in n-tile gq >> 1, columns 4 (gq & 1) .. + 3: lanes tq = 2 (gq & 1) (dims 0, 1) and + 1 (dims 2, 3)
    float v0 = o[0][0], v1 = o[0][1];
#pragma unroll
    for (int c = 1; c < 4; c++)
        if ((gq >> 1) == c) { v0 = o[c][0]; v1 = o[c][1]; }
    // softmax row sum = the ones column (dim d) of the head: lane holding dims (d & ~1, d | 1), element d & 1
    const int lsrc = (lane & ~3) | (2 * (gq & 1) + (d >> 1));
    const float l = __shfl_sync(0xffffffffu, (d & 1) ? v1 : v0, lsrc);
    const float inv = rcpf(l);
    if ((tq >> 1) == (gq & 1)) {
        uint8_t* dst = sA2 + (uint32_t)(gq >> 1) * WF_LBO + (uint32_t)(rowbase + 48) * 16u + (uint32_t)(4 * (gq & 1) + 2 * (tq & 1)) * 2u;
        *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2(v0 * inv, v1 * inv);
    }
}

__global__ void __launch_bounds__(WS_THREADS, 1) k_wa_ws(WaFused p) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int HW = WS_HW, NQKV = WS_NQKV, PH = WS_PH;
    constexpr uint32_t PITCH = WS_PITCH;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Kpad = p.Kpad, N2 = p.N2;
    const bool self_attn = p.self_attn != 0;
    const WsSmem L = ws_layout(Kpad, N2, self_attn, p.C);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
    uint64_t* w_full = bars;             // weights landed
    uint64_t* d1_full = bars + 1;        // [2] q|k|v accumulator of a tile complete
    uint64_t* d2_full = bars + 3;        // [2] projection accumulator complete
    uint64_t* qkv_full = bars + 5;       // [2] helpers -> attention: fp16 q|k|v rows of a tile written
    uint64_t* qkv_empty = bars + 7;      // [2] attention -> helpers: rows consumed
    uint64_t* a2_full = bars + 9;        // [2] attention -> issuer: O of a tile written (A2)
    uint64_t* a2_empty = bars + 11;      // [2] tensor core -> attention: A2 consumed by the projection MMA
    uint64_t* raw_full = bars + 13;      // [RS] bulk-copied source rows of a tile landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13 + WS_RS);
    const int ntiles = (p.nwin + WF_WIN - 1) / WF_WIN;
    const int J = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;   // tiles of this CTA
    const uint32_t kc1 = (uint32_t)Kpad >> 3;
    const uint32_t wq_bytes = kc1 * (self_attn ? (uint32_t)NQKV : (uint32_t)HW) * 16u;
    const uint32_t wkv_bytes = self_attn ? 0u : kc1 * 2u * (uint32_t)HW * 16u;
    const uint32_t wo_bytes = ((uint32_t)HW >> 3) * (uint32_t)N2 * 16u;

    if (tid == 0) {
        mbar_init(w_full, 1);
        for (int i = 0; i < 2; i++) {
            mbar_init(&d1_full[i], 1); mbar_init(&d2_full[i], 1);
            mbar_init(&qkv_full[i], WS_HELPERS); mbar_init(&qkv_empty[i], WS_ATT);
            mbar_init(&a2_full[i], WS_ATT); mbar_init(&a2_empty[i], 1);
        }
        for (int i = 0; i < WS_RS; i++) mbar_init(&raw_full[i], 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(w_full, wq_bytes + wkv_bytes + wo_bytes);
        bulk_g2s(smem + L.wq, p.Wq, wq_bytes, w_full);
        if (!self_attn) bulk_g2s(smem + L.wkv, p.Wkv, wkv_bytes, w_full);
        bulk_g2s(smem + L.wo, p.Wo, wo_bytes, w_full);
    }
    if (warp == 1) tmem_alloc(tmem_slot, WS_TMEM_COLS);
    {
        // A1 / A2 (both stages): rows 98..127 and the K padding columns are never written again and must be zero
        uint4* z = reinterpret_cast<uint4*>(smem + L.a1q[0]);
        const int n16 = (int)((L.qkv[0] - L.a1q[0]) >> 4);
        for (int i = tid; i < n16; i += WS_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
        float* sb = reinterpret_cast<float*>(smem + L.bias);
        for (int i = tid; i < NQKV; i += WS_THREADS) sb[i] = self_attn ? __ldg(p.bq + i) : (i < HW ? __ldg(p.bq + i) : __ldg(p.bkv + i - HW));
        for (int i = tid; i < N2; i += WS_THREADS) sb[NQKV + i] = __ldg(p.bo + i);
        // bias row of query token 48 (window row 6, column 6) against the 49 keys, x log2 e; padded keys -1e30 (a001:113-144)
        float* b48 = reinterpret_cast<float*>(smem + L.b48);
        for (int key = tid; key < 56; key += WS_THREADS) {
            const int kr = key / 7, kc = key - kr * 7;
            b48[key] = key < WF_T ? 1.4426950408889634f * __ldg(p.table + (kr - 6 + 6) * 13 + (kc - 6 + 6)) : -1e30f;
        }
    }
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const float* sbias = reinterpret_cast<const float*>(smem + L.bias);
    const WinGeom& g = p.wo.g;
    auto tile_of = [&](int j) { return (int)blockIdx.x + j * (int)gridDim.x; };
    auto tile_rows = [&](int t) { return min(WF_WIN, p.nwin - t * WF_WIN) * WF_T; };

    if (warp < WS_HELPERS) {
        // =========================== helper warps ===========================================================================
        const int rb = warp;                      // TMEM lane quarter
        const int row = rb * 32 + lane;
        const uint32_t tlane = tmem_base + ((uint32_t)(rb * 32) << 16);
        const uint32_t raw_kv = self_attn ? 0u : L.raw_stride / 2u;
        const bool res_staged = p.residual == p.q_src;   // x + Attn(LN(x), ..): the residual rows are the staged q rows
        const int C = p.C, nf4 = C >> 2;
        // Gather of this CTA's j-th tile into staging slot j % RS by the bulk-copy engine: lane l of warp 1 owns window row
        // l (7 tokens = one contiguous run of the source map, two runs where the cyclic shift wraps the last window column).
        auto prefetch = [&](int j) {
            if (j >= J || warp != 1) return;
            const int t = tile_of(j);
            const int nw = min(WF_WIN, p.nwin - t * WF_WIN);
            const uint32_t slot = (uint32_t)(j % WS_RS);
            uint8_t* raw = smem + L.raw + slot * L.raw_stride;
            if (lane == 0) mbar_arrive_expect_tx(&raw_full[slot], (uint32_t)(nw * WF_T * C * 4) * (self_attn ? 1u : 2u));
            __syncwarp();
            if (lane < nw * 7) {
                const int w = lane / 7, i = lane - 7 * w;
                const uint32_t win = (uint32_t)(t * WF_WIN + w);
                const uint32_t b = fdiv(win, p.wo.dnW), wi = win - b * (uint32_t)(g.nWh * g.nWw);
                const uint32_t wh = fdiv(wi, p.wo.dnWw), ww = wi - wh * (uint32_t)g.nWw;
                const int sr = shift_src((int)(wh * g.wsh) + i, g.Hp, g.sh);
                const int c0 = shift_src((int)(ww * g.wsw), g.Wp, g.sw);
                const int run1 = min(7, g.Wp - c0);
                const long long base = ((long long)b * g.Hp + sr) * g.Wp;
                const uint32_t doff = (uint32_t)((w * WF_T + i * 7) * C * 4);
                bulk_g2s(raw + doff, p.q_src + (base + c0) * C, (uint32_t)(run1 * C * 4), &raw_full[slot]);
                if (run1 < 7) bulk_g2s(raw + doff + (uint32_t)(run1 * C * 4), p.q_src + base * C, (uint32_t)((7 - run1) * C * 4), &raw_full[slot]);
                if (!self_attn) {
                    bulk_g2s(raw + raw_kv + doff, p.kv_src + (base + c0) * C, (uint32_t)(run1 * C * 4), &raw_full[slot]);
                    if (run1 < 7) bulk_g2s(raw + raw_kv + doff + (uint32_t)(run1 * C * 4), p.kv_src + base * C, (uint32_t)((7 - run1) * C * 4), &raw_full[slot]);
                }
            }
        };
        // staged fp32 row -> LayerNorm -> bf16 A1 row (one thread per row: 98 of the 128 helper threads)
        auto produce = [&](uint8_t* sA, const uint8_t* raw, const float* lg, const float* lb, int nrows) {
            if (row >= nrows) return;
            const float4* src = reinterpret_cast<const float4*>(raw + (uint32_t)row * (uint32_t)C * 4u);
            float4 v[8];
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = i < nf4 ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
            if (lg) {
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int i = 0; i < 8; i++) { s0 += v[i].x + v[i].y; s1 += v[i].z + v[i].w; }
                const float invc = 1.f / (float)C, mean = (s0 + s1) * invc;
                float q0 = 0.f, q1 = 0.f;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if (i < nf4) {
                        const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
                        q0 += dx * dx + dy * dy; q1 += dz * dz + dw * dw;
                    }
                }
                const float rstd = rsqrtf((q0 + q1) * invc + p.eps);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if (i < nf4) {
                        const float4 gg = __ldg(reinterpret_cast<const float4*>(lg) + i), bb = __ldg(reinterpret_cast<const float4*>(lb) + i);
                        v[i].x = (v[i].x - mean) * rstd * gg.x + bb.x;
                        v[i].y = (v[i].y - mean) * rstd * gg.y + bb.y;
                        v[i].z = (v[i].z - mean) * rstd * gg.z + bb.z;
                        v[i].w = (v[i].w - mean) * rstd * gg.w + bb.w;
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 4; c++) {
                if (2 * c < nf4)
                    *reinterpret_cast<uint4*>(sA + (uint32_t)c * WF_LBO + (uint32_t)row * 16u) =
                        make_uint4(pack_bf16x2(v[2 * c].x, v[2 * c].y), pack_bf16x2(v[2 * c].z, v[2 * c].w),
                                   pack_bf16x2(v[2 * c + 1].x, v[2 * c + 1].y), pack_bf16x2(v[2 * c + 1].z, v[2 * c + 1].w));
            }
        };
        // Operand descriptors of the two stages, built once: the issuing thread sits on the helpers' critical chain (A -> barrier ->
        // issue -> C -> issue -> F per tile) and spent ~900 cycles per tile assembling the same eight 64-bit words (r2 phase trace).
        uint64_t dA1[2][2], dW1[2], dA2[2][2], dW2[2];
        {
            const uint32_t lbo_w = (uint32_t)NQKV * 16u, lbo_o = (uint32_t)N2 * 16u;
#pragma unroll
            for (int ks = 0; ks < 2; ks++) {
                dW1[ks] = make_smem_desc(smem_u32(smem + L.wq) + (uint32_t)ks * 2u * lbo_w, lbo_w, WF_SBO);
                dW2[ks] = make_smem_desc(smem_u32(smem + L.wo) + (uint32_t)ks * 2u * lbo_o, lbo_o, WF_SBO);
#pragma unroll
                for (int sg = 0; sg < 2; sg++) {
                    dA1[sg][ks] = make_smem_desc(smem_u32(smem + L.a1q[sg]) + (uint32_t)ks * 2u * WF_LBO, WF_LBO, WF_SBO);
                    dA2[sg][ks] = make_smem_desc(smem_u32(smem + L.a2[sg]) + (uint32_t)ks * 2u * WF_LBO, WF_LBO, WF_SBO);
                }
            }
        }
        const bool fast_desc = self_attn && Kpad == 32;   // two k-steps per product (HW = 32 always)
        for (int j = 0; j < WS_PD; j++) prefetch(j);
#pragma unroll 1
        for (int j = 0; j < J + 3; j++) {
            // Order inside an iteration: C(j-1) comes FIRST.  The attention warps are waiting for exactly these rows (qkv_full), the
            // accumulator they come from was committed an iteration ago, and nothing of A(j) / B(j) is needed for them -- with C after
            // the barrier and the MMA issue, the rows arrived ~2,500 cycles later than they could (r2 phase trace).
            // ---- C(j-1): D1 -> + bias -> fp16 q|k|v rows of the stage the attention warps read next -----------------------------
            if (j >= 1 && j - 1 < J) {
                const int jc = j - 1;
                const uint32_t st = (uint32_t)jc & 1u, par = ((uint32_t)jc >> 1) & 1u;
                const int nrows = tile_rows(tile_of(jc));
                mbar_wait_relaxed(&d1_full[st], par);
                mbar_wait_relaxed(&qkv_empty[st], par ^ 1u);   // rows of tile jc - 2 consumed
                __syncwarp();
                tc_fence_after_sync();
                uint8_t* qrow = smem + L.qkv[st] + (uint32_t)row * PITCH;
#pragma unroll
                for (int half = 0; half < ((p.debug & 4) ? 0 : 2); half++) {   // three TMEM loads in flight per wait
                    uint32_t r[3][16];
#pragma unroll
                    for (int u = 0; u < 3; u++) tmem_ld16_issue(tlane + st * WS_D1_STRIDE + (uint32_t)(half * 48 + u * 16), r[u]);
                    tmem_ld_wait();
                    if (row < nrows) {
#pragma unroll
                        for (int u = 0; u < 3; u++) {
                            const int c16 = half * 48 + u * 16;
                            float v[16];
#pragma unroll
                            for (int i = 0; i < 16; i += 4) {
                                const float4 bb = *reinterpret_cast<const float4*>(sbias + c16 + i);
                                v[i] = __uint_as_float(r[u][i]) + bb.x; v[i + 1] = __uint_as_float(r[u][i + 1]) + bb.y;
                                v[i + 2] = __uint_as_float(r[u][i + 2]) + bb.z; v[i + 3] = __uint_as_float(r[u][i + 3]) + bb.w;
                            }
                            uint4* dst = reinterpret_cast<uint4*>(qrow + c16 * 2);
                            if (c16 < 2 * HW) {   // q, k: fp16
                                dst[0] = make_uint4(pack_f16x2(v[0], v[1]), pack_f16x2(v[2], v[3]), pack_f16x2(v[4], v[5]), pack_f16x2(v[6], v[7]));
                                dst[1] = make_uint4(pack_f16x2(v[8], v[9]), pack_f16x2(v[10], v[11]), pack_f16x2(v[12], v[13]), pack_f16x2(v[14], v[15]));
                            } else {              // v: bf16, the B operand of P V with P = 2^s (wf_softmax_p)
                                dst[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
                                dst[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
                            }
                        }
                    }
                }
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) ws_arrive(&qkv_full[st]);
            }
            // ---- A(j), B(j): staged rows -> LayerNorm -> A1[j & 1]; q|k|v projection -> D1[j & 1] --------------------------------
            if (j < J) {
                const int t = tile_of(j);
                const uint32_t st = (uint32_t)j & 1u;
                const uint32_t slot = (uint32_t)(j % WS_RS);
                const int nrows = tile_rows(t);
                mbar_wait_relaxed(&raw_full[slot], (uint32_t)(j / WS_RS) & 1u);
                const uint8_t* raw = smem + L.raw + slot * L.raw_stride;
                if (!(p.debug & 2)) {
                    produce(smem + L.a1q[st], raw, p.ln_q_g, p.ln_q_b, nrows);
                    if (!self_attn) produce(smem + L.a1kv[st], raw + raw_kv, p.ln_kv_g, p.ln_kv_b, nrows);
                }
                fence_async_smem();
            }
            ws_helper_bar();   // A1 of tile j complete; every helper has finished iteration j - 1 (F(j-4) released its staging slot)
            if (j < J) {
                if (tid == 0) {
                    const uint32_t st = (uint32_t)j & 1u;
                    if (j == 0) mbar_wait(w_full, 0);
                    tc_fence_after_sync();
                    const uint32_t a1q = smem_u32(smem + L.a1q[st]), wq = smem_u32(smem + L.wq);
                    const uint32_t d1 = tmem_base + st * WS_D1_STRIDE;
                    if (fast_desc) {
                        const uint32_t idesc = make_idesc_bf16(128, (uint32_t)NQKV);
                        umma_bf16(d1, st ? dA1[1][0] : dA1[0][0], dW1[0], idesc, false);
                        umma_bf16(d1, st ? dA1[1][1] : dA1[0][1], dW1[1], idesc, true);
                    } else if (self_attn) {
                        const uint32_t lbo_w = (uint32_t)NQKV * 16u, idesc = make_idesc_bf16(128, (uint32_t)NQKV);
                        for (uint32_t ks = 0; ks < (uint32_t)Kpad >> 4; ks++)
                            umma_bf16(d1, make_smem_desc(a1q + ks * 2u * WF_LBO, WF_LBO, WF_SBO),
                                      make_smem_desc(wq + ks * 2u * lbo_w, lbo_w, WF_SBO), idesc, ks > 0);
                    } else {
                        const uint32_t a1kv = smem_u32(smem + L.a1kv[st]), wkv = smem_u32(smem + L.wkv);
                        const uint32_t lbo_q = (uint32_t)HW * 16u, lbo_kv = 2u * (uint32_t)HW * 16u;
                        const uint32_t idq = make_idesc_bf16(128, (uint32_t)HW), idkv = make_idesc_bf16(128, 2u * (uint32_t)HW);
                        for (uint32_t ks = 0; ks < (uint32_t)Kpad >> 4; ks++)
                            umma_bf16(d1, make_smem_desc(a1q + ks * 2u * WF_LBO, WF_LBO, WF_SBO),
                                      make_smem_desc(wq + ks * 2u * lbo_q, lbo_q, WF_SBO), idq, ks > 0);
                        for (uint32_t ks = 0; ks < (uint32_t)Kpad >> 4; ks++)
                            umma_bf16(d1 + (uint32_t)HW, make_smem_desc(a1kv + ks * 2u * WF_LBO, WF_LBO, WF_SBO),
                                      make_smem_desc(wkv + ks * 2u * lbo_kv, lbo_kv, WF_SBO), idkv, ks > 0);
                    }
                    umma_commit(&d1_full[st]);
                }
                __syncwarp();
            }
            // ---- E(j-2): the attention warps have written O of tile j-2 -> projection on tcgen05 --------------------------------------
            if (j >= 2 && j - 2 < J && tid == 0) {
                const int je = j - 2;
                const uint32_t st = (uint32_t)je & 1u, par = ((uint32_t)je >> 1) & 1u;
                mbar_wait(&a2_full[st], par);
                tc_fence_after_sync();
                const uint32_t idesc = make_idesc_bf16(128, (uint32_t)N2);
                umma_bf16(tmem_base + WS_D2_COL + st * WS_D2_STRIDE, st ? dA2[1][0] : dA2[0][0], dW2[0], idesc, false);
                umma_bf16(tmem_base + WS_D2_COL + st * WS_D2_STRIDE, st ? dA2[1][1] : dA2[0][1], dW2[1], idesc, true);
                umma_commit(&d2_full[st]);
                umma_commit(&a2_empty[st]);
            }
            __syncwarp();
            // ---- F(j-3): D2 + b_o + residual -> fp32 rows, scattered back (window reverse + un-shift as index math) --------------------
            if (j >= 3 && j - 3 < J && !(p.debug & 8)) {
                const int jf = j - 3;
                const uint32_t st = (uint32_t)jf & 1u, par = ((uint32_t)jf >> 1) & 1u;
                const int t = tile_of(jf);
                const int nrows = tile_rows(t);
                const long long mo = row < nrows ? win_order_token(p.wo, (uint32_t)t * WF_ROWS + (uint32_t)row) : 0;
                // the residual row: still in its staging slot when it is the q source (x read from HBM exactly once), else from L2 / HBM
                const float* rsrc = res_staged ? reinterpret_cast<const float*>(smem + L.raw + (uint32_t)(jf % WS_RS) * L.raw_stride) + (size_t)row * C
                                               : (p.residual ? p.residual + mo * C : nullptr);
                float4 res[8];
#pragma unroll
                for (int i = 0; i < 8; i++)
                    res[i] = (rsrc && row < nrows && i < nf4) ? *reinterpret_cast<const float4*>(rsrc + i * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                mbar_wait_relaxed(&d2_full[st], par);
                __syncwarp();
                tc_fence_after_sync();
                uint32_t r[2][16];
                tmem_ld16_issue(tlane + WS_D2_COL + st * WS_D2_STRIDE, r[0]);
                if (N2 > 16) tmem_ld16_issue(tlane + WS_D2_COL + st * WS_D2_STRIDE + 16u, r[1]);
                tmem_ld_wait();
                if (row < nrows) {
                    float* o = p.out + mo * C;
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        if (i < nf4) {
                            const float4 bb = *reinterpret_cast<const float4*>(sbias + NQKV + i * 4);
                            const uint32_t* rr = r[i >> 2] + (i & 3) * 4;
                            *reinterpret_cast<float4*>(o + i * 4) = make_float4(__uint_as_float(rr[0]) + bb.x + res[i].x, __uint_as_float(rr[1]) + bb.y + res[i].y,
                                                                              __uint_as_float(rr[2]) + bb.z + res[i].z, __uint_as_float(rr[3]) + bb.w + res[i].w);
                        }
                    }
                }
                tc_fence_before_sync();
            }
            // gather of tile j + PD: issued last -- warp 1's address arithmetic for it used to sit between the barrier and its share of
            // the q|k|v rows, i.e. on the path to qkv_full; the copy itself has two tiles of time.  (Its slot held tile j + PD - RS =
            // j - 4, scattered in iteration j - 1 by every helper: all of them passed this iteration's barrier since.)
            if (j < J) prefetch(j + WS_PD);
        }
    } else {
        // =========================== attention warps ===============================================================================
        const int a = warp - WS_HELPERS;          // 0..11
        const int w = a / 6, a6 = a - 6 * w;      // window of the tile, position among its six warps
        const int slab = a6 >> 1, par = a6 & 1;
        const int gq = lane >> 2, tq = lane & 3;
        const int r0 = slab * 16 + gq;
        float bias[7][4];
        const SlabMask sm = slab_bias_and_mask(bias, p.table, r0, r0 + 8, tq);
        const float* sb48 = reinterpret_cast<const float*>(smem + L.b48);
#pragma unroll 1
        for (int j = 0; j < J; j++) {
            const uint32_t st = (uint32_t)j & 1u, par2 = ((uint32_t)j >> 1) & 1u;
            const int t = tile_of(j);
            const int nw = min(WF_WIN, p.nwin - t * WF_WIN);
            mbar_wait_relaxed(&qkv_full[st], par2);
            mbar_wait_relaxed(&a2_empty[st], par2 ^ 1u);   // the projection MMA of tile j - 2 has consumed this A2 stage
            __syncwarp();
            if (w < nw && !(p.debug & 1)) {
                bool mh = false, mw = false;
                if (g.shift) {   // boundary windows of the shifted frame are the only ones whose tokens span several regions
                    const uint32_t win = (uint32_t)(t * WF_WIN + w);
                    const uint32_t wi = win - fdiv(win, p.wo.dnW) * (uint32_t)(g.nWh * g.nWw);
                    const uint32_t wh = fdiv(wi, p.wo.dnWw), ww = wi - wh * (uint32_t)g.nWw;
                    mh = wh == (uint32_t)g.nWh - 1;
                    mw = ww == (uint32_t)g.nWw - 1;
                }
                const uint32_t m0 = (mh ? sm.mh0 : 0u) | (mw ? sm.mw0 : 0u), m1 = (mh ? sm.mh1 : 0u) | (mw ? sm.mw1 : 0u);
                const __half* wbase = reinterpret_cast<const __half*>(smem + L.qkv[st]) + w * WF_T * PH;
                wf_attn_pack4<PH, HW>(wbase, smem + L.a2[st], w * WF_T, bias, m0, m1, p.d, r0, gq, tq, par);
                if (a6 == j % 6) ws_attn_tok48(wbase, smem + L.a2[st], w * WF_T, sb48, mh, mw, p.d, gq, tq, lane);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) { ws_arrive(&a2_full[st]); ws_arrive(&qkv_empty[st]); }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, WS_TMEM_COLS);
}

bool wa_ws_supported(const WinGeom& g, int C, int nh, int d, bool self_attn) {
    static const bool off = [] { const char* e = getenv("SWINFUSE_WA_WS"); return e && e[0] == '0'; }();
    if (off || !wa_fused_supported(g, C, nh, d) || d > 3) return false;
    // measured at the stage-0 shape (B = 64, 133 x 133, C = 24): self attention 356 us here vs 386 us bulk-synchronous; cross
    // attention (two sources to normalise, two more MMAs per tile in the four helper warps) 470 us vs 433 us -> cross stays
    // with wa_fused.cu unless forced (SWINFUSE_WA_WS=2)
    static const bool force = [] { const char* e = getenv("SWINFUSE_WA_WS"); return e && e[0] == '2'; }();
    if (!self_attn && !force) return false;
    const int Kpad = (int)pad16((uint32_t)C);
    return ws_layout(Kpad, Kpad, self_attn, C).total <= WS_SMEM_LIMIT;
}

int launch_wa_ws(const WaFused& a, cudaStream_t st) {
    const WsSmem L = ws_layout(a.Kpad, a.N2, a.self_attn != 0, a.C);
    SF_CHECK_ARG(L.total <= WS_SMEM_LIMIT, "wa_ws: tile does not fit shared memory");
    static DeviceOnce configured;
    if (configured.need()) {
        cudaError_t e = cudaFuncSetAttribute(k_wa_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_SMEM_LIMIT);
        if (e != cudaSuccess) { set_error("wa_ws: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SF_ERR_CUDA; }
        configured.done();
    }
    const long long ntiles = ((long long)a.nwin + WF_WIN - 1) / WF_WIN;
    long long grid = sm_count();
    if (grid > ntiles) grid = ntiles;
    const double mtok = (double)a.nwin * WF_T;
    const int inner = 8 * a.d;
    const double maps = (a.self_attn ? 2.0 : 3.0) + ((a.residual && a.residual != a.q_src && a.residual != a.kv_src) ? 1.0 : 0.0);
    ProfScope ps(prof_name("wa_fused_c%d", a.C), 8.0 * mtok * a.C * inner + 4.0 * WF_T * mtok * inner, 4.0 * mtok * a.C * maps, st);
    k_wa_ws<<<(unsigned)grid, WS_THREADS, L.total, st>>>(a);
    SF_CHECK_LAUNCH("wa_ws");
    return SF_OK;
}

}  // namespace sf

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
#include "wa_common.cuh"

namespace sf {
using namespace tc;

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
                if (col < 2 * HW) {   // q, k: fp16
                    dst[0] = make_uint4(pack_f16x2(v[0], v[1]), pack_f16x2(v[2], v[3]), pack_f16x2(v[4], v[5]), pack_f16x2(v[6], v[7]));
                    dst[1] = make_uint4(pack_f16x2(v[8], v[9]), pack_f16x2(v[10], v[11]), pack_f16x2(v[12], v[13]), pack_f16x2(v[14], v[15]));
                } else {              // v: bf16, the B operand of P V with P = 2^s (wf_softmax_p)
                    dst[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
                    dst[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
                }
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
#pragma unroll
            for (int w = 0; w < WF_WIN; w++) {
                if (w >= nw) break;
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
    if (wa_ws_supported(g, p->C, p->num_heads, p->head_dim, self_attn)) return launch_wa_ws(a, st);
    return p->head_dim <= 3 ? launch_wa_fused_t<true>(a, st) : launch_wa_fused_t<false>(a, st);
}

}  // namespace sf

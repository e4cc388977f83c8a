// SF_PREC_BF16 operator implementations: which tensor-core kernels run for each fused operator, how
// the weights are packed (once, when the caller provides a `packed` buffer; per call otherwise) and
// how the caller's workspace is carved.  Every linear layer goes through tcgen05.mma (tc_gemm.cu);
// the per-window attention core runs on HMMA tiles (attn_frag.cu); LayerNorm, softmax, bias, ELU and
// the residual stream stay fp32.
#include <cstdlib>
#include "bf16_kernels.cuh"
#include "fp32_kernels.cuh"
#include "tc_common.cuh"

namespace sf {

// rows at least this wide are normalised by the LayerNorm pre-pass (k_ln_to_tiled) instead of the GEMM's own producers
static int ln_prepass_min_c() {
    static const int v = [] { const char* e = getenv("SWINFUSE_LN_PREPASS_MIN_C"); return e ? atoi(e) : TC_LN_PREPASS_MIN_C; }();
    return v;
}

using namespace tc;

// =============================================================================================
// window attention:  [LN] QKV GEMM (fp16 rows) -> HMMA attention core (bf16, UMMA-tiled O)
//                    -> output projection GEMM (+bias +residual, fp32 rows)
// =============================================================================================
struct WaPlan {
    bool self_attn;
    bool frag;                    // 7x7 windows: window-order GEMMs + HMMA attention core (attn_frag.cu)
    bool fused;                   // narrow stages (C <= 64): the whole operator is ONE kernel (wa_fused.cu), nothing staged in HBM
    int inner, dp, hw;            // frag: padded head width of the q / k / v / O columns, hw = heads * dp
    WinGeom geom;
    PackedGemm q, kv, o;          // q: stacked q|k|v for self attention
    size_t packed_bytes;
    bool prepass_q, prepass_kv;
    size_t off_qkv, off_o, off_packed, off_nq, off_nkv, total;   // workspace
};

static WaPlan wa_plan(const sf_window_attn_params* p) {
    WaPlan w{};
    const long long M = (long long)p->B * p->Hp * p->Wp;
    w.self_attn = p->kv_src == p->q_src && p->ln_q_gamma == p->ln_kv_gamma && p->ln_q_beta == p->ln_kv_beta;
    w.inner = p->num_heads * p->head_dim;
    w.geom = make_geom(p->B, p->Hp, p->Wp, p->wsh, p->wsw, p->shift);
    w.frag = attn_frag_supported(w.geom, p->num_heads, p->head_dim) && M < 2147483647LL / 64;
    w.fused = w.frag && wa_fused_supported(w.geom, p->C, p->num_heads, p->head_dim);
    w.dp = qkvh_dp(p->head_dim);
    w.hw = p->num_heads * w.dp;
    Carver pc;
    if (w.frag) {
        w.q = plan_packed(pc, w.self_attn ? 3 * w.hw : w.hw, p->C);
        if (!w.self_attn) w.kv = plan_packed(pc, 2 * w.hw, p->C);
    } else {
        w.q = plan_packed(pc, w.self_attn ? 3 * w.inner : w.inner, p->C);
        if (!w.self_attn) w.kv = plan_packed(pc, 2 * w.inner, p->C);
    }
    w.o = plan_packed(pc, p->C, w.frag ? w.hw : w.inner);
    w.packed_bytes = pc.off;
    Carver c;
    const int ocols = w.frag ? w.hw : w.inner;   // columns of q, k, v and O rows
    w.off_qkv = c.take(w.fused ? 0 : tiled_elems(M, 3 * ocols) * sizeof(bf16));   // fp16 rows
    w.off_o = c.take(w.fused ? 0 : tiled_elems(M, ocols) * sizeof(bf16));
    w.off_packed = c.take(p->packed ? 0 : w.packed_bytes);
    w.prepass_q = p->ln_q_gamma != nullptr && p->C >= ln_prepass_min_c();
    w.prepass_kv = !w.self_attn && p->ln_kv_gamma != nullptr && p->C >= ln_prepass_min_c();
    w.off_nq = c.take(w.prepass_q ? tiled_elems(M, p->C) * sizeof(bf16) : 0);
    w.off_nkv = c.take(w.prepass_kv ? tiled_elems(M, p->C) * sizeof(bf16) : 0);
    w.total = c.off;
    return w;
}

static int wa_check(const sf_window_attn_params* p) {
    if (p->C % 4 != 0 || (int)pad16((uint32_t)p->C) > TC_MAX_KPAD || !aligned16(p->q_src) || !aligned16(p->kv_src) || !aligned16(p->out) ||
        (p->residual && !aligned16(p->residual))) {
        set_error("bf16 window attention supports C %% 4 == 0, C <= %d and 16-byte aligned maps (got C=%d, heads*dim=%d)", TC_MAX_KPAD,
                  p->C, p->num_heads * p->head_dim);
        return SF_ERR_UNSUPPORTED;
    }
    return SF_OK;
}

size_t window_attn_ws_bf16(const sf_window_attn_params* p) { return wa_plan(p).total; }
size_t window_attn_packed_bytes_bf16(const sf_window_attn_params* p) { return wa_plan(p).packed_bytes; }

int window_attn_pack_bf16(const sf_window_attn_params* p, void* packed, size_t bytes, cudaStream_t st) {
    SF_TRY(wa_check(p));
    const WaPlan w = wa_plan(p);
    if (bytes < w.packed_bytes) { set_error("sf_window_attn_pack: buffer too small (%zu B given, %zu needed)", bytes, w.packed_bytes); return SF_ERR_WORKSPACE; }
    char* base = reinterpret_cast<char*>(packed);
    const int inner = w.inner, C = p->C;
    if (w.frag) {
        // N axis = [q heads | k heads | v heads], every head padded to dp columns; q (weights and bias) carries d^-1/2 * log2(e):
        // the attention core works in the log2 domain (a001:32-34,335 applies the scale after the product -- same value)
        const float qs = 1.4426950408889634f / sqrtf((float)p->head_dim);
        if (w.self_attn) {
            PackSrc s{{p->wq, p->wk, p->wv}, {p->bq, p->bk, p->bv}};
            PackMap m{3, {w.hw, w.hw, w.hw}, {1, 1, 1}, {qs, 1.f, 1.f}, p->head_dim, w.dp, 0, 0, {0, 0, 1}};
            SF_TRY(launch_pack_mapped(s, m, C, (bf16*)(base + w.q.off_w), (float*)(base + w.q.off_b), w.q.nch, w.q.ks, w.q.nc, w.q.nslabs, st));
        } else {
            PackSrc s1{{p->wq, nullptr, nullptr}, {p->bq, nullptr, nullptr}};
            PackMap m1{1, {w.hw, 0, 0}, {1, 0, 0}, {qs, 1.f, 1.f}, p->head_dim, w.dp, 0, 0, {0, 0, 0}};
            SF_TRY(launch_pack_mapped(s1, m1, C, (bf16*)(base + w.q.off_w), (float*)(base + w.q.off_b), w.q.nch, w.q.ks, w.q.nc, w.q.nslabs, st));
            PackSrc s2{{p->wk, p->wv, nullptr}, {p->bk, p->bv, nullptr}};
            PackMap m2{2, {w.hw, w.hw, 0}, {1, 1, 0}, {1.f, 1.f, 1.f}, p->head_dim, w.dp, 0, 0, {0, 1, 0}};
            SF_TRY(launch_pack_mapped(s2, m2, C, (bf16*)(base + w.kv.off_w), (float*)(base + w.kv.off_b), w.kv.nch, w.kv.ks, w.kv.nc, w.kv.nslabs, st));
        }
    } else if (w.self_attn) {
        PackSrc s{{p->wq, p->wk, p->wv}, {p->bq, p->bk, p->bv}};
        SF_TRY(launch_pack(s, 3, inner, C, (bf16*)(base + w.q.off_w), (float*)(base + w.q.off_b), w.q.nch, w.q.ks, w.q.nc, w.q.nslabs, st));
    } else {
        PackSrc s1{{p->wq, nullptr, nullptr}, {p->bq, nullptr, nullptr}};
        SF_TRY(launch_pack(s1, 1, inner, C, (bf16*)(base + w.q.off_w), (float*)(base + w.q.off_b), w.q.nch, w.q.ks, w.q.nc, w.q.nslabs, st));
        PackSrc s2{{p->wk, p->wv, nullptr}, {p->bk, p->bv, nullptr}};
        SF_TRY(launch_pack(s2, 2, inner, C, (bf16*)(base + w.kv.off_w), (float*)(base + w.kv.off_b), w.kv.nch, w.kv.ks, w.kv.nc, w.kv.nslabs, st));
    }
    PackSrc so{{p->wo, nullptr, nullptr}, {p->bo, nullptr, nullptr}};
    if (w.frag) {   // the K axis of the projection follows O's head padding
        PackMap mo{1, {C, 0, 0}, {0, 0, 0}, {1.f, 1.f, 1.f}, 0, 0, p->head_dim, w.dp, {0, 0, 0}};
        SF_TRY(launch_pack_mapped(so, mo, inner, (bf16*)(base + w.o.off_w), (float*)(base + w.o.off_b), w.o.nch, w.o.ks, w.o.nc, w.o.nslabs, st));
    } else {
        SF_TRY(launch_pack(so, 1, C, inner, (bf16*)(base + w.o.off_w), (float*)(base + w.o.off_b), w.o.nch, w.o.ks, w.o.nc, w.o.nslabs, st));
    }
    return SF_OK;
}

int window_attn_fwd_bf16(const sf_window_attn_params* p, void* ws_ptr, size_t ws_bytes, cudaStream_t st) {
    SF_TRY(wa_check(p));
    const WaPlan w = wa_plan(p);
    const int inner = w.inner, C = p->C;
    if (ws_bytes < w.total || (w.total && !ws_ptr)) { set_error("sf_window_attn_fwd: workspace too small (%zu B given, %zu needed)", ws_bytes, w.total); return SF_ERR_WORKSPACE; }
    char* base = reinterpret_cast<char*>(ws_ptr);
    const long long M = (long long)p->B * p->Hp * p->Wp;
    bf16* qkv = reinterpret_cast<bf16*>(base + w.off_qkv);
    bf16* O = reinterpret_cast<bf16*>(base + w.off_o);
    const char* pk = reinterpret_cast<const char*>(p->packed);
    if (!pk) {
        SF_TRY(window_attn_pack_bf16(p, base + w.off_packed, w.packed_bytes, st));
        pk = base + w.off_packed;
    }
    if (w.fused) {
        // LN -> q|k|v projection -> attention core -> output projection -> + residual, scattered back: one kernel
        const bf16* Wkv = w.self_attn ? nullptr : reinterpret_cast<const bf16*>(pk + w.kv.off_w);
        const float* bkv = w.self_attn ? nullptr : reinterpret_cast<const float*>(pk + w.kv.off_b);
        return launch_wa_fused(p, w.geom, w.self_attn, reinterpret_cast<const bf16*>(pk + w.q.off_w), reinterpret_cast<const float*>(pk + w.q.off_b),
                               Wkv, bkv, reinterpret_cast<const bf16*>(pk + w.o.off_w), reinterpret_cast<const float*>(pk + w.o.off_b), st);
    }
    // projections -> q|k|v (fp16).  7x7 windows: the GEMMs run on rows in window order (A producers gather, shift and
    // partition as index math) and write per-(window, head) blobs for the HMMA core; otherwise plain fp16 rows.
    const WinGeom& geom = w.geom;
    const WinOrder wo = make_winorder(geom);
    const int ocols = w.frag ? w.hw : inner;   // columns of q, k, v (head padded when w.frag)
    TcGemm g{};
    g.M = M; g.K = C; g.lda = C; g.eps = p->ln_eps; g.out_fp16 = 1;
    g.out_mode = OUT_BF16; g.out = qkv; g.ldo = 3 * ocols;
    g.N = w.self_attn ? 3 * ocols : ocols;
    if (w.frag) { g.win_order = 1; g.wo = wo; }
    g.A = p->q_src; g.ln_g = p->ln_q_gamma; g.ln_b = p->ln_q_beta; g.a_mode = p->ln_q_gamma ? AM_F32_LN : AM_F32;
    bind_packed(g, w.q, pk);
    g.out_col0 = 0;
    if (w.prepass_q) {   // wide rows: LayerNorm once into the UMMA-tiled layout, GEMM streams it
        bf16* nq = reinterpret_cast<bf16*>(base + w.off_nq);
        SF_TRY(launch_ln_to_tiled(p->q_src, p->ln_q_gamma, p->ln_q_beta, nq, M, C, p->ln_eps, st, w.frag ? &wo : nullptr));
        g.A = nq; g.a_mode = AM_TILED;
    }
    SF_TRY(tc_gemm_plan(&g));
    SF_TRY(launch_tc_gemm(g, prof_name(w.self_attn ? "tc_gemm_qkv_c%d" : "tc_gemm_q_c%d", C), st));
    if (!w.self_attn) {
        TcGemm k = g;
        k.A = p->kv_src; k.ln_g = p->ln_kv_gamma; k.ln_b = p->ln_kv_beta; k.a_mode = p->ln_kv_gamma ? AM_F32_LN : AM_F32;
        bind_packed(k, w.kv, pk);
        k.out_col0 = ocols; k.N = 2 * ocols;
        if (w.prepass_kv) {
            bf16* nkv = reinterpret_cast<bf16*>(base + w.off_nkv);
            SF_TRY(launch_ln_to_tiled(p->kv_src, p->ln_kv_gamma, p->ln_kv_beta, nkv, M, C, p->ln_eps, st, w.frag ? &wo : nullptr));
            k.A = nkv; k.a_mode = AM_TILED;
        }
        SF_TRY(tc_gemm_plan(&k));
        SF_TRY(launch_tc_gemm(k, prof_name("tc_gemm_kv_c%d", C), st));
    }
    // attention core -> O (bf16, UMMA-tiled so the projection GEMM can bulk-copy it; window order when w.frag)
    if (w.frag) {
        SF_TRY(launch_attn_frag(reinterpret_cast<const __half*>(qkv), 3 * ocols, O, p->bias_table, geom, p->num_heads, p->head_dim, st));
    } else {
        const int o_nkc = (int)pad16((uint32_t)inner) / 8;
        SF_TRY(launch_attn_core_bf16(qkv, 1, 3 * inner, inner, 2 * inner, O, 0, o_nkc, p->bias_table, geom, p->num_heads, p->head_dim, st));
    }
    // output projection (+ residual) -> out fp32 rows (scattered back: window reverse and un-shift, a001:373-398,442-445)
    TcGemm o{};
    o.M = M; o.K = ocols; o.A = O; o.a_mode = AM_TILED; o.out_mode = OUT_F32;
    bind_packed(o, w.o, pk);
    o.residual = p->residual; o.ldr = C;
    o.out = p->out; o.ldo = C; o.out_col0 = 0; o.N = C;
    if (w.frag) { o.win_order = 1; o.wo = wo; }
    SF_TRY(tc_gemm_plan(&o));
    SF_TRY(launch_tc_gemm(o, prof_name("tc_gemm_proj_c%d", C), st));
    return SF_OK;
}

// =============================================================================================
// MLP: narrow stages (C <= 64) run the fused persistent kernel of tc_mlp.cu (hidden activation stays on chip);
// wider ones GEMM1 (LN, +b1, ELU, bf16 UMMA-tiled hidden) + GEMM2 (bulk-copied A, +b2 +residual).
// =============================================================================================
struct MlpPlan { PackedGemm g1, g2; bool prepass; size_t packed_bytes, off_h, off_packed, off_n, total; };

static MlpPlan mlp_plan(const sf_mlp_params* p) {
    MlpPlan m{};
    Carver pc;
    m.g1 = plan_packed(pc, p->hidden, p->C);
    m.g2 = plan_packed(pc, p->C, p->hidden);
    m.packed_bytes = pc.off;
    Carver c;
    m.off_h = c.take(tiled_elems(p->M, p->hidden) * sizeof(bf16));
    m.off_packed = c.take(p->packed ? 0 : m.packed_bytes);
    m.prepass = p->ln_gamma != nullptr && p->C >= ln_prepass_min_c();
    m.off_n = c.take(m.prepass ? tiled_elems(p->M, p->C) * sizeof(bf16) : 0);
    m.total = c.off;
    return m;
}

static int mlp_check(const sf_mlp_params* p) {
    if (p->C % 4 != 0 || (int)pad16((uint32_t)p->C) > TC_MAX_KPAD || !aligned16(p->in) || !aligned16(p->out) || (p->residual && !aligned16(p->residual))) {
        set_error("bf16 MLP supports C %% 4 == 0, C <= %d and 16-byte aligned maps (got C=%d, hidden=%d)", TC_MAX_KPAD, p->C, p->hidden);
        return SF_ERR_UNSUPPORTED;
    }
    return SF_OK;
}

size_t mlp_ws_bf16(const sf_mlp_params* p) { return mlp_plan(p).total; }
size_t mlp_packed_bytes_bf16(const sf_mlp_params* p) { return mlp_plan(p).packed_bytes; }

int mlp_pack_bf16(const sf_mlp_params* p, void* packed, size_t bytes, cudaStream_t st) {
    SF_TRY(mlp_check(p));
    const MlpPlan m = mlp_plan(p);
    if (bytes < m.packed_bytes) { set_error("sf_mlp_pack: buffer too small (%zu B given, %zu needed)", bytes, m.packed_bytes); return SF_ERR_WORKSPACE; }
    char* base = reinterpret_cast<char*>(packed);
    PackSrc s1{{p->w1, nullptr, nullptr}, {p->b1, nullptr, nullptr}};
    SF_TRY(launch_pack(s1, 1, p->hidden, p->C, (bf16*)(base + m.g1.off_w), (float*)(base + m.g1.off_b), m.g1.nch, m.g1.ks, m.g1.nc, m.g1.nslabs, st));
    PackSrc s2{{p->w2, nullptr, nullptr}, {p->b2, nullptr, nullptr}};
    SF_TRY(launch_pack(s2, 1, p->C, p->hidden, (bf16*)(base + m.g2.off_w), (float*)(base + m.g2.off_b), m.g2.nch, m.g2.ks, m.g2.nc, m.g2.nslabs, st));
    return SF_OK;
}

int mlp_fwd_bf16(const sf_mlp_params* p, void* ws_ptr, size_t ws_bytes, cudaStream_t st) {
    SF_TRY(mlp_check(p));
    const MlpPlan m = mlp_plan(p);
    if (ws_bytes < m.total || !ws_ptr) { set_error("sf_mlp_fwd: workspace too small (%zu B given, %zu needed)", ws_bytes, m.total); return SF_ERR_WORKSPACE; }
    char* base = reinterpret_cast<char*>(ws_ptr);
    const char* pk = reinterpret_cast<const char*>(p->packed);
    if (!pk) {
        SF_TRY(mlp_pack_bf16(p, base + m.off_packed, m.packed_bytes, st));
        pk = base + m.off_packed;
    }
    if (tc_mlp_supported(p->C, p->hidden) && m.g1.nc == 1 && m.g2.nc == 1) {
        TcMlp t{};
        t.x = p->in; t.residual = p->residual; t.out = p->out; t.M = p->M;
        t.C = p->C; t.Cpad = m.g1.kpad; t.hidden = p->hidden; t.Hpad = m.g1.nch;
        t.ln_g = p->ln_gamma; t.ln_b = p->ln_beta; t.eps = p->ln_eps;
        t.max_stages = 8;
        if (const char* e = getenv("SWINFUSE_MLP_STAGES")) t.max_stages = atoi(e) < 2 ? 2 : (atoi(e) > 8 ? 8 : atoi(e));
        t.W1p = reinterpret_cast<const bf16*>(pk + m.g1.off_w); t.b1 = reinterpret_cast<const float*>(pk + m.g1.off_b);
        t.W2p = reinterpret_cast<const bf16*>(pk + m.g2.off_w); t.b2 = reinterpret_cast<const float*>(pk + m.g2.off_b);
        if (m.g2.nch != t.Cpad || m.g2.kpad != t.Hpad) { set_error("mlp: packed weight plan does not match the fused kernel"); return SF_ERR_INVALID; }
        return launch_tc_mlp(t, st);
    }
    bf16* hid = reinterpret_cast<bf16*>(base + m.off_h);
    TcGemm g1{};
    g1.A = p->in; g1.M = p->M; g1.K = p->C; g1.lda = p->C; g1.a_mode = p->ln_gamma ? AM_F32_LN : AM_F32; g1.out_mode = OUT_TILED;
    g1.ln_g = p->ln_gamma; g1.ln_b = p->ln_beta; g1.eps = p->ln_eps;
    bind_packed(g1, m.g1, pk);
    g1.elu = 1;
    g1.out = hid; g1.N = p->hidden; g1.out_nkc = (int)pad16((uint32_t)p->hidden) / 8;
    if (m.prepass) {
        bf16* nb = reinterpret_cast<bf16*>(base + m.off_n);
        SF_TRY(launch_ln_to_tiled(p->in, p->ln_gamma, p->ln_beta, nb, p->M, p->C, p->ln_eps, st));
        g1.A = nb; g1.a_mode = AM_TILED;
    }
    SF_TRY(tc_gemm_plan(&g1));
    SF_TRY(launch_tc_gemm(g1, prof_name("tc_gemm_mlp1_c%d", p->C), st));
    TcGemm g2{};
    g2.A = hid; g2.M = p->M; g2.K = p->hidden; g2.a_mode = AM_TILED; g2.out_mode = OUT_F32;
    bind_packed(g2, m.g2, pk);
    g2.residual = p->residual; g2.ldr = p->C;
    g2.out = p->out; g2.ldo = p->C; g2.N = p->C;
    SF_TRY(tc_gemm_plan(&g2));
    SF_TRY(launch_tc_gemm(g2, prof_name("tc_gemm_mlp2_c%d", p->C), st));
    return SF_OK;
}

// =============================================================================================
// patch layers: tcgen05 GEMM (gather / plain prologue) -> fp32 rows -> LayerNorm(+ELU)(+un-merge)
// =============================================================================================
struct PatchPlan { bool tc; int K, N; long long Mrows; PackedGemm g; size_t packed_bytes, off_lin, off_packed, off_f32, total; };

static PatchPlan patch_plan(const sf_patch_params* p) {
    PatchPlan q{};
    const int mm = p->mh * p->mw;
    if (p->encoder) { q.K = mm * p->Cin; q.N = p->Cout; q.Mrows = (long long)p->B * (p->H / p->mh) * (p->W / p->mw); }
    else { q.K = p->Cin; q.N = mm * p->Cout; q.Mrows = (long long)p->B * p->H * p->W; }
    // layers outside the tensor-core tile limits run the fp32 kernels (higher precision, same ABI)
    q.tc = (int)pad16((uint32_t)q.K) <= TC_MAX_KPAD && (p->encoder || p->Cin % 4 == 0) && aligned16(p->in);
    Carver c;
    if (q.tc) {
        Carver pc;
        q.g = plan_packed(pc, q.N, q.K);
        q.packed_bytes = pc.off;
        q.off_lin = c.take((size_t)q.Mrows * q.N * sizeof(float));
        q.off_packed = c.take(p->packed ? 0 : q.packed_bytes);
    } else {
        q.off_f32 = c.take(patch_ws_f32(p));
    }
    q.total = c.off;
    return q;
}

size_t patch_ws_bf16(const sf_patch_params* p) { return patch_plan(p).total; }
size_t patch_packed_bytes_bf16(const sf_patch_params* p) { return patch_plan(p).packed_bytes; }

int patch_pack_bf16(const sf_patch_params* p, void* packed, size_t bytes, cudaStream_t st) {
    const PatchPlan q = patch_plan(p);
    if (!q.tc) return SF_OK;
    if (bytes < q.packed_bytes) { set_error("sf_patch_pack: buffer too small (%zu B given, %zu needed)", bytes, q.packed_bytes); return SF_ERR_WORKSPACE; }
    char* base = reinterpret_cast<char*>(packed);
    PackSrc s{{p->w, nullptr, nullptr}, {p->b, nullptr, nullptr}};
    return launch_pack(s, 1, q.N, q.K, (bf16*)(base + q.g.off_w), (float*)(base + q.g.off_b), q.g.nch, q.g.ks, q.g.nc, q.g.nslabs, st);
}

int patch_fwd_bf16(const sf_patch_params* p, void* ws_ptr, size_t ws_bytes, cudaStream_t st) {
    const PatchPlan q = patch_plan(p);
    if (ws_bytes < q.total || !ws_ptr) { set_error("sf_patch_fwd: workspace too small (%zu B given, %zu needed)", ws_bytes, q.total); return SF_ERR_WORKSPACE; }
    char* base = reinterpret_cast<char*>(ws_ptr);
    if (!q.tc) return patch_fwd_f32(p, base + q.off_f32, ws_bytes - q.off_f32, st);
    const char* pk = reinterpret_cast<const char*>(p->packed);
    if (!pk) {
        SF_TRY(patch_pack_bf16(p, base + q.off_packed, q.packed_bytes, st));
        pk = base + q.off_packed;
    }
    float* lin = reinterpret_cast<float*>(base + q.off_lin);
    TcGemm g{};
    g.A = p->in; g.M = q.Mrows; g.K = q.K; g.lda = q.K; g.out_mode = OUT_F32;
    g.a_mode = p->encoder ? AM_MERGE : AM_F32;
    bind_packed(g, q.g, pk);
    g.out = lin; g.ldo = q.N; g.N = q.N;
    g.Hf = p->H; g.Wf = p->W; g.Cin = p->Cin; g.mh = p->mh; g.mw = p->mw;
    SF_TRY(tc_gemm_plan(&g));
    SF_TRY(launch_tc_gemm(g, p->encoder ? "tc_gemm_patch_merge" : "tc_gemm_patch_expand", st));
    if (p->encoder) {
        SF_TRY(launch_layernorm(lin, p->ln_gamma, p->ln_beta, p->out, q.Mrows, q.N, p->ln_eps, 1, nullptr, st));
    } else {
        UnmergeGeom ug{p->H, p->W, p->mh, p->mw, p->Cout};
        SF_TRY(launch_layernorm(lin, p->ln_gamma, p->ln_beta, p->out, q.Mrows, q.N, p->ln_eps, 1, &ug, st));
    }
    return SF_OK;
}

}  // namespace sf

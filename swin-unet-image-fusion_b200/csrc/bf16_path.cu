// SF_PREC_BF16 operator implementations: which tcgen05 kernels run for each fused operator and how
// the caller's workspace is carved.  Every linear layer goes through tcgen05.mma (tc_gemm.cu /
// tc_mlp.cu); LayerNorm, softmax, bias, ELU and the residual stream stay fp32.
#include "bf16_kernels.cuh"
#include "fp32_kernels.cuh"
#include "tc_common.cuh"

namespace sf {
using namespace tc;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static size_t tiled_elems(long long M, int cols) { return (size_t)((M + 127) / 128) * 128 * pad16((uint32_t)cols); }

struct Carver {
    size_t off = 0;
    size_t take(size_t bytes) { size_t o = off; off += align_up(bytes); return o; }
};

// =============================================================================================
// window attention:  pack -> [LN] QKV GEMM (bf16 rows) -> attention core (bf16, UMMA-tiled O)
//                    -> output projection GEMM (+bias +residual, fp32 rows)
// =============================================================================================
struct WaPlan {
    bool self_attn;
    int inner;
    int nch_q, nc_q, nch_kv, nc_kv, nch_o, nc_o;
    int kpad_c, kpad_i;
    size_t off_qkv, off_o, off_wq, off_wkv, off_wo, off_bq, off_bkv, off_bo, total;
};

static WaPlan wa_plan(const sf_window_attn_params* p) {
    WaPlan w{};
    const long long M = (long long)p->B * p->Hp * p->Wp;
    w.self_attn = p->kv_src == p->q_src && p->ln_q_gamma == p->ln_kv_gamma && p->ln_q_beta == p->ln_kv_beta;
    w.inner = p->num_heads * p->head_dim;
    w.kpad_c = (int)pad16((uint32_t)p->C);
    w.kpad_i = (int)pad16((uint32_t)w.inner);
    tc_gemm_pick_nchunk(w.self_attn ? 3 * w.inner : w.inner, &w.nch_q, &w.nc_q);
    if (!w.self_attn) tc_gemm_pick_nchunk(2 * w.inner, &w.nch_kv, &w.nc_kv);
    tc_gemm_pick_nchunk(p->C, &w.nch_o, &w.nc_o);
    Carver c;
    w.off_qkv = c.take((size_t)M * 3 * w.inner * sizeof(bf16));
    w.off_o = c.take(tiled_elems(M, w.inner) * sizeof(bf16));
    w.off_wq = c.take((size_t)w.nc_q * w.nch_q * w.kpad_c * sizeof(bf16));
    w.off_wkv = c.take((size_t)w.nc_kv * w.nch_kv * w.kpad_c * sizeof(bf16) + 16);
    w.off_wo = c.take((size_t)w.nc_o * w.nch_o * w.kpad_i * sizeof(bf16));
    w.off_bq = c.take((size_t)w.nc_q * w.nch_q * sizeof(float));
    w.off_bkv = c.take((size_t)w.nc_kv * w.nch_kv * sizeof(float) + 16);
    w.off_bo = c.take((size_t)w.nc_o * w.nch_o * sizeof(float));
    w.total = c.off;
    return w;
}

size_t window_attn_ws_bf16(const sf_window_attn_params* p) { return wa_plan(p).total; }

int window_attn_fwd_bf16(const sf_window_attn_params* p, void* ws_ptr, size_t ws_bytes, cudaStream_t st) {
    const WaPlan w = wa_plan(p);
    const int inner = w.inner, C = p->C;
    if (C % 4 != 0 || w.kpad_c > TC_MAX_KPAD || !aligned16(p->q_src) || !aligned16(p->kv_src) || !aligned16(p->out) ||
        (p->residual && !aligned16(p->residual))) {
        set_error("bf16 window attention supports C %% 4 == 0, C <= %d and 16-byte aligned maps (got C=%d, heads*dim=%d)", TC_MAX_KPAD, C, inner);
        return SF_ERR_UNSUPPORTED;
    }
    if (ws_bytes < w.total || !ws_ptr) { set_error("sf_window_attn_fwd: workspace too small (%zu B given, %zu needed)", ws_bytes, w.total); return SF_ERR_WORKSPACE; }
    char* base = reinterpret_cast<char*>(ws_ptr);
    const long long M = (long long)p->B * p->Hp * p->Wp;
    bf16* qkv = reinterpret_cast<bf16*>(base + w.off_qkv);
    bf16* O = reinterpret_cast<bf16*>(base + w.off_o);
    bf16* wq = reinterpret_cast<bf16*>(base + w.off_wq);
    bf16* wkv = reinterpret_cast<bf16*>(base + w.off_wkv);
    bf16* wo = reinterpret_cast<bf16*>(base + w.off_wo);
    float* bq = reinterpret_cast<float*>(base + w.off_bq);
    float* bkv = reinterpret_cast<float*>(base + w.off_bkv);
    float* bo = reinterpret_cast<float*>(base + w.off_bo);

    // projections -> qkv [M][3*inner] bf16 rows
    TcGemm g{};
    g.M = M; g.K = C; g.lda = C; g.eps = p->ln_eps; g.out_mode = OUT_BF16; g.out_fp16 = 1;
    g.out = qkv; g.ldo = 3 * inner;
    g.A = p->q_src; g.ln_g = p->ln_q_gamma; g.ln_b = p->ln_q_beta; g.a_mode = p->ln_q_gamma ? AM_F32_LN : AM_F32;
    g.Wp = wq; g.NCH = w.nch_q; g.n_chunks = w.nc_q; g.bias = bq; g.out_col0 = 0; g.N = w.self_attn ? 3 * inner : inner;
    SF_TRY(tc_gemm_plan(&g));
    if (w.self_attn) {
        PackSrc s{{p->wq, p->wk, p->wv}, {p->bq, p->bk, p->bv}};
        SF_TRY(launch_pack(s, 3, inner, C, wq, bq, g.NCH, g.KS, g.n_chunks, g.n_slabs, st));
    } else {
        PackSrc s{{p->wq, nullptr, nullptr}, {p->bq, nullptr, nullptr}};
        SF_TRY(launch_pack(s, 1, inner, C, wq, bq, g.NCH, g.KS, g.n_chunks, g.n_slabs, st));
    }
    SF_TRY(launch_tc_gemm(g, prof_name(w.self_attn ? "tc_gemm_qkv_c%d" : "tc_gemm_q_c%d", C), st));
    if (!w.self_attn) {
        TcGemm k = g;
        k.A = p->kv_src; k.ln_g = p->ln_kv_gamma; k.ln_b = p->ln_kv_beta; k.a_mode = p->ln_kv_gamma ? AM_F32_LN : AM_F32;
        k.Wp = wkv; k.NCH = w.nch_kv; k.n_chunks = w.nc_kv; k.bias = bkv; k.out_col0 = inner; k.N = 2 * inner;
        SF_TRY(tc_gemm_plan(&k));
        PackSrc s{{p->wk, p->wv, nullptr}, {p->bk, p->bv, nullptr}};
        SF_TRY(launch_pack(s, 2, inner, C, wkv, bkv, k.NCH, k.KS, k.n_chunks, k.n_slabs, st));
        SF_TRY(launch_tc_gemm(k, prof_name("tc_gemm_kv_c%d", C), st));
    }
    // attention core -> O (bf16, UMMA-tiled so the projection GEMM can bulk-copy it)
    WinGeom geom = make_geom(p->B, p->Hp, p->Wp, p->wsh, p->wsw, p->shift);
    SF_TRY(launch_attn_core_bf16(qkv, 1, 3 * inner, inner, 2 * inner, O, 0, w.kpad_i / 8, p->bias_table, geom, p->num_heads, p->head_dim, st));
    // output projection (+ residual) -> out fp32 rows
    TcGemm o{};
    o.M = M; o.K = inner; o.A = O; o.a_mode = AM_TILED; o.out_mode = OUT_F32;
    o.Wp = wo; o.NCH = w.nch_o; o.n_chunks = w.nc_o; o.bias = bo; o.residual = p->residual; o.ldr = C;
    o.out = p->out; o.ldo = C; o.out_col0 = 0; o.N = C;
    SF_TRY(tc_gemm_plan(&o));
    PackSrc so{{p->wo, nullptr, nullptr}, {p->bo, nullptr, nullptr}};
    SF_TRY(launch_pack(so, 1, C, inner, wo, bo, o.NCH, o.KS, o.n_chunks, o.n_slabs, st));
    SF_TRY(launch_tc_gemm(o, prof_name("tc_gemm_proj_c%d", C), st));
    return SF_OK;
}

// =============================================================================================
// MLP: C <= 64  -> fused kernel (hidden activation never leaves the SM)
//      C  > 64  -> GEMM1 (LN, +b1, ELU, bf16 UMMA-tiled hidden) + GEMM2 (bulk-copied A, +b2 +residual)
// =============================================================================================
struct MlpPlan {
    bool fused;
    int Cpad, HC, n_hc;                 // fused
    int nch1, nc1, nch2, nc2, hpad;     // two-GEMM
    size_t off_w1, off_w2, off_b1, off_b2, off_h, total;
};

static MlpPlan mlp_plan(const sf_mlp_params* p) {
    MlpPlan m{};
    m.Cpad = (int)pad16((uint32_t)p->C);
    m.hpad = (int)pad16((uint32_t)p->hidden);
    m.fused = false;   // the non-persistent fused kernel (tc_mlp.cu) loses to two pipelined GEMMs; see DESIGN.md
    Carver c;
    if (m.fused) {
        m.HC = tc_mlp_pick_hc(m.Cpad, p->hidden);
        m.n_hc = m.HC ? (m.hpad + m.HC - 1) / m.HC : 0;
        m.off_w1 = c.take((size_t)m.n_hc * m.HC * m.Cpad * sizeof(bf16));
        m.off_w2 = c.take((size_t)m.n_hc * m.HC * m.Cpad * sizeof(bf16));
        m.off_b1 = c.take((size_t)m.n_hc * m.HC * sizeof(float));
    } else {
        tc_gemm_pick_nchunk(p->hidden, &m.nch1, &m.nc1);
        tc_gemm_pick_nchunk(p->C, &m.nch2, &m.nc2);
        m.off_w1 = c.take((size_t)m.nc1 * m.nch1 * m.Cpad * sizeof(bf16));
        m.off_w2 = c.take((size_t)m.nc2 * m.nch2 * m.hpad * sizeof(bf16));
        m.off_b1 = c.take((size_t)m.nc1 * m.nch1 * sizeof(float));
        m.off_b2 = c.take((size_t)m.nc2 * m.nch2 * sizeof(float));
        m.off_h = c.take(tiled_elems(p->M, p->hidden) * sizeof(bf16));
    }
    m.total = c.off;
    return m;
}

size_t mlp_ws_bf16(const sf_mlp_params* p) { return mlp_plan(p).total; }

int mlp_fwd_bf16(const sf_mlp_params* p, void* ws_ptr, size_t ws_bytes, cudaStream_t st) {
    const MlpPlan m = mlp_plan(p);
    if (p->C % 4 != 0 || m.Cpad > TC_MAX_KPAD || (m.fused && m.HC == 0) || !aligned16(p->in) || !aligned16(p->out) ||
        (p->residual && !aligned16(p->residual))) {
        set_error("bf16 MLP supports C %% 4 == 0, C <= %d and 16-byte aligned maps (got C=%d, hidden=%d)", TC_MAX_KPAD, p->C, p->hidden);
        return SF_ERR_UNSUPPORTED;
    }
    if (ws_bytes < m.total || !ws_ptr) { set_error("sf_mlp_fwd: workspace too small (%zu B given, %zu needed)", ws_bytes, m.total); return SF_ERR_WORKSPACE; }
    char* base = reinterpret_cast<char*>(ws_ptr);
    bf16* w1 = reinterpret_cast<bf16*>(base + m.off_w1);
    bf16* w2 = reinterpret_cast<bf16*>(base + m.off_w2);
    float* b1 = reinterpret_cast<float*>(base + m.off_b1);
    PackSrc s1{{p->w1, nullptr, nullptr}, {p->b1, nullptr, nullptr}};
    PackSrc s2{{p->w2, nullptr, nullptr}, {p->b2, nullptr, nullptr}};
    if (m.fused) {
        SF_TRY(launch_pack(s1, 1, p->hidden, p->C, w1, b1, m.HC, m.Cpad, m.n_hc, 1, st));          // rows = hidden chunks
        SF_TRY(launch_pack(s2, 1, p->C, p->hidden, w2, nullptr, m.Cpad, m.HC, 1, m.n_hc, st));      // k    = hidden chunks
        TcMlp t{};
        t.x = p->in; t.residual = p->residual; t.out = p->out; t.M = p->M;
        t.C = p->C; t.Cpad = m.Cpad; t.hidden = p->hidden; t.HC = m.HC; t.n_hc = m.n_hc;
        t.ln_g = p->ln_gamma; t.ln_b = p->ln_beta; t.eps = p->ln_eps;
        t.W1p = w1; t.W2p = w2; t.b1 = b1; t.b2 = p->b2;
        return launch_tc_mlp(t, st);
    }
    float* b2 = reinterpret_cast<float*>(base + m.off_b2);
    bf16* hid = reinterpret_cast<bf16*>(base + m.off_h);
    TcGemm g1{};
    g1.A = p->in; g1.M = p->M; g1.K = p->C; g1.lda = p->C; g1.a_mode = p->ln_gamma ? AM_F32_LN : AM_F32; g1.out_mode = OUT_TILED;
    g1.ln_g = p->ln_gamma; g1.ln_b = p->ln_beta; g1.eps = p->ln_eps;
    g1.Wp = w1; g1.NCH = m.nch1; g1.n_chunks = m.nc1; g1.bias = b1; g1.elu = 1;
    g1.out = hid; g1.N = p->hidden; g1.out_nkc = m.hpad / 8;
    SF_TRY(tc_gemm_plan(&g1));
    SF_TRY(launch_pack(s1, 1, p->hidden, p->C, w1, b1, g1.NCH, g1.KS, g1.n_chunks, g1.n_slabs, st));
    SF_TRY(launch_tc_gemm(g1, prof_name("tc_gemm_mlp1_c%d", p->C), st));
    TcGemm g2{};
    g2.A = hid; g2.M = p->M; g2.K = p->hidden; g2.a_mode = AM_TILED; g2.out_mode = OUT_F32;
    g2.Wp = w2; g2.NCH = m.nch2; g2.n_chunks = m.nc2; g2.bias = b2; g2.residual = p->residual; g2.ldr = p->C;
    g2.out = p->out; g2.ldo = p->C; g2.N = p->C;
    SF_TRY(tc_gemm_plan(&g2));
    SF_TRY(launch_pack(s2, 1, p->C, p->hidden, w2, b2, g2.NCH, g2.KS, g2.n_chunks, g2.n_slabs, st));
    SF_TRY(launch_tc_gemm(g2, prof_name("tc_gemm_mlp2_c%d", p->C), st));
    return SF_OK;
}

// =============================================================================================
// patch layers: tcgen05 GEMM (gather / plain prologue) -> fp32 rows -> LayerNorm(+ELU)(+un-merge)
// =============================================================================================
struct PatchPlan { bool tc; int K, N, nch, nc; long long Mrows; size_t off_lin, off_w, off_b, off_f32, total; };

static PatchPlan patch_plan(const sf_patch_params* p) {
    PatchPlan q{};
    const int mm = p->mh * p->mw;
    if (p->encoder) { q.K = mm * p->Cin; q.N = p->Cout; q.Mrows = (long long)p->B * (p->H / p->mh) * (p->W / p->mw); }
    else { q.K = p->Cin; q.N = mm * p->Cout; q.Mrows = (long long)p->B * p->H * p->W; }
    // layers outside the tensor-core tile limits run the fp32 kernels (higher precision, same ABI)
    q.tc = (int)pad16((uint32_t)q.K) <= TC_MAX_KPAD && (p->encoder || p->Cin % 4 == 0) && aligned16(p->in);
    Carver c;
    if (q.tc) {
        tc_gemm_pick_nchunk(q.N, &q.nch, &q.nc);
        q.off_lin = c.take((size_t)q.Mrows * q.N * sizeof(float));
        q.off_w = c.take((size_t)q.nc * q.nch * pad16((uint32_t)q.K) * sizeof(bf16));
        q.off_b = c.take((size_t)q.nc * q.nch * sizeof(float));
    } else {
        q.off_f32 = c.take(patch_ws_f32(p));
    }
    q.total = c.off;
    return q;
}

size_t patch_ws_bf16(const sf_patch_params* p) { return patch_plan(p).total; }

int patch_fwd_bf16(const sf_patch_params* p, void* ws_ptr, size_t ws_bytes, cudaStream_t st) {
    const PatchPlan q = patch_plan(p);
    if (ws_bytes < q.total || !ws_ptr) { set_error("sf_patch_fwd: workspace too small (%zu B given, %zu needed)", ws_bytes, q.total); return SF_ERR_WORKSPACE; }
    char* base = reinterpret_cast<char*>(ws_ptr);
    if (!q.tc) return patch_fwd_f32(p, base + q.off_f32, ws_bytes - q.off_f32, st);
    float* lin = reinterpret_cast<float*>(base + q.off_lin);
    bf16* wp = reinterpret_cast<bf16*>(base + q.off_w);
    float* bp = reinterpret_cast<float*>(base + q.off_b);
    TcGemm g{};
    g.A = p->in; g.M = q.Mrows; g.K = q.K; g.lda = q.K; g.out_mode = OUT_F32;
    g.a_mode = p->encoder ? AM_MERGE : AM_F32;
    g.Wp = wp; g.NCH = q.nch; g.n_chunks = q.nc; g.bias = bp; g.out = lin; g.ldo = q.N; g.N = q.N;
    g.Hf = p->H; g.Wf = p->W; g.Cin = p->Cin; g.mh = p->mh; g.mw = p->mw;
    SF_TRY(tc_gemm_plan(&g));
    PackSrc s{{p->w, nullptr, nullptr}, {p->b, nullptr, nullptr}};
    SF_TRY(launch_pack(s, 1, q.N, q.K, wp, bp, g.NCH, g.KS, g.n_chunks, g.n_slabs, st));
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

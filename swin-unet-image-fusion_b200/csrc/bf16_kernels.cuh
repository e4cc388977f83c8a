// bf16 tensor-core (tcgen05) path: shared declarations.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace sf {
using bf16 = __nv_bfloat16;

static constexpr int TC_MAX_KPAD = 384;   // widest K whose A tile is kept resident in shared memory

// ---- operator entry points (api.cu dispatches here for SF_PREC_BF16) ---------------------------
size_t window_attn_ws_bf16(const sf_window_attn_params* p);
int window_attn_fwd_bf16(const sf_window_attn_params* p, void* ws, size_t ws_bytes, cudaStream_t st);
size_t mlp_ws_bf16(const sf_mlp_params* p);
int mlp_fwd_bf16(const sf_mlp_params* p, void* ws, size_t ws_bytes, cudaStream_t st);
size_t patch_ws_bf16(const sf_patch_params* p);
int patch_fwd_bf16(const sf_patch_params* p, void* ws, size_t ws_bytes, cudaStream_t st);

size_t window_attn_packed_bytes_bf16(const sf_window_attn_params* p);
int window_attn_pack_bf16(const sf_window_attn_params* p, void* packed, size_t bytes, cudaStream_t st);
size_t mlp_packed_bytes_bf16(const sf_mlp_params* p);
int mlp_pack_bf16(const sf_mlp_params* p, void* packed, size_t bytes, cudaStream_t st);
size_t patch_packed_bytes_bf16(const sf_patch_params* p);
int patch_pack_bf16(const sf_patch_params* p, void* packed, size_t bytes, cudaStream_t st);

// ---- weight packing (tc_gemm.cu) -------------------------------------------------------------------
// Stacks up to 3 [Neach x K] fp32 matrices along N, splits rows into n_chunks of NR and columns
// into k_chunks of KR, and writes for every (jn, jk) a bf16 UMMA operand image img[kc][r][8]
// (= W[jn*NR + r][jk*KR + kc*8 + e], zero outside), images ordered jn-major; optional stacked,
// zero-padded fp32 bias [n_chunks*NR].
struct PackSrc { const float* w[3]; const float* b[3]; };
int launch_pack(const PackSrc& src, int nsrc, int Neach, int K, bf16* out, float* bias_out, int NR, int KR, int n_chunks,
                int k_chunks, cudaStream_t st, int transposed = 0);
// Up to MAX images in one launch: image i is that of W_i ([N][K] row-major, or -- transposed -- stored [K][N]) with its
// zero-padded bias; plan fields as launch_pack's (NR = n-chunk rows, KR = k-slab width).
// Up to three sources are stacked: along N (plain: rows [s*each, (s+1)*each) come from w[s], b[s]) or along K (transposed:
// k in [s*each, (s+1)*each) comes from w[s], each stored [each][N]).
struct PackJob { const float* w[3]; const float* b[3]; int each; int N, K, transposed; bf16* out; float* bias_out; int NR, KR, n_chunks, k_chunks; };
struct PackJobs { static constexpr int MAX = 8; int n; PackJob job[MAX]; };
int launch_pack_jobs(const PackJobs& jobs, cudaStream_t st);
// Same images, but the N axis is a concatenation of sources with their own layout: source s has
// rows[s] packed rows; head_padded[s] != 0 means packed row n' = h*dp + dd maps to source row h*d + dd
// (zero when dd >= d); weights and bias of source s are multiplied by scale[s].
// kdp > 0: the K axis is head padded too (packed column k' = h*kdp + dd <- source column h*kd + dd).
// ones_pad[s] != 0: the first padding row of every head of source s (dd == d, needs d < dp) gets zero weights
// and bias 1 -- a column of ones in the GEMM output (the attention core reads its softmax row sums off it).
struct PackMap { int nsrc; int rows[3]; int head_padded[3]; float scale[3]; int d, dp; int kd, kdp; int ones_pad[3]; };
int launch_pack_mapped(const PackSrc& src, const PackMap& map, int K, bf16* out, float* bias_out, int NR, int KR, int n_chunks,
                       int k_chunks, cudaStream_t st);

// ---- persistent token GEMM (tc_gemm.cu) ---------------------------------------------------------------
enum { AM_F32 = 0, AM_F32_LN = 1, AM_TILED = 2, AM_MERGE = 3 };
enum { OUT_F32 = 0, OUT_BF16 = 1, OUT_TILED = 2 };

// head-padded column layout shared by the window-attention GEMMs and the HMMA attention core
// (attn_frag.cu): a head of d dims occupies dp columns (zeros above d), dp = 4 or a multiple of 8
static inline int qkvh_dp(int d) { return d <= 4 ? 4 : (d + 7) / 8 * 8; }

struct TcGemm {
    // ---- caller fills ----
    const void* A;            // fp32 rows (AM_F32, AM_F32_LN), fp32 fine map (AM_MERGE), bf16 UMMA-tiled (AM_TILED)
    long long M;
    int K;
    long long lda;            // elements (fp32 row modes)
    int a_mode, out_mode;
    const float* ln_g; const float* ln_b; float eps;
    const bf16* Wp;           // packed weights: image [KS/8][NCH][8] per (n-chunk, k-slab)
    int NCH, n_chunks;
    const float* bias;        // [n_chunks*NCH] zero padded, or null
    int elu;
    int out_fp16;             // OUT_BF16 rows are written as IEEE fp16 instead of bf16
    const float* residual; long long ldr;   // OUT_F32 only
    void* out; long long ldo; int out_col0; int N;
    int out_nkc;              // OUT_TILED: k-chunks per tile of the destination (= pad16(total columns)/8)
    int Hf, Wf, Cin, mh, mw;  // AM_MERGE: fine map (B,Hf,Wf,Cin) and merging factors
    const bf16* elu_aux;      // OUT_TILED: multiply by ELU'(pre) derived from this saved activation ELU(pre), same tiled layout as out
    int zero_tail;            // OUT_TILED: rows M .. end of the last tile are written as zeros
    int win_order;            // GEMM rows are in window order: fp32 A producers gather source rows, OUT_F32 scatters rows
    WinOrder wo;
    // ---- tc_gemm_plan fills ----
    int a_tile_nkc, a_kc0;    // AM_TILED, optional: A is the chunk range [a_kc0, ..) of a tiled tensor with a_tile_nkc chunks per tile
    int Kpad, KS, n_slabs, a_nkc, NA, NS, n_groups, chunks_per_group;
    int coal;                 // OUT_F32: rows leave through the shared-memory transposition (tc_gemm_plan sets it)
};
void tc_gemm_pick_nchunk(int Ntot, int* NCH, int* n_chunks);
int tc_gemm_pick_ks(int Kpad);          // k-slab width: largest of {64,48,32,16} dividing Kpad
int tc_gemm_plan(TcGemm* p);          // needs K, a_mode, NCH, n_chunks, M
int launch_tc_gemm(const TcGemm& p, const char* name, cudaStream_t st);
// fp32 rows -> LayerNorm -> bf16 UMMA-tiled (pre-pass for wide rows, see tc_gemm.cu)
// `wo` != null: output row m is the token win_order_token(*wo, m) of `in` (window order)
// ld_in: row stride of `in` (0 = C); out_nkc / out_kc0: the C columns are written as the chunk range [out_kc0, ..) of a wider
// tiled tensor with out_nkc chunks per tile (0 = a tensor of its own)
int launch_ln_to_tiled(const float* in, const float* gamma, const float* beta, bf16* out, long long M, int C, float eps, cudaStream_t st,
                       const WinOrder* wo = nullptr, long long ld_in = 0, int out_nkc = 0, int out_kc0 = 0);
static constexpr int TC_LN_PREPASS_MIN_C = 96;   // rows at least this wide are normalised by the pre-pass

// ---- workspace carving and packed-weight plans shared by the operator implementations (bf16_path.cu, tc_bwd.cu) ----
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline size_t tiled_elems(long long M, int cols) { return (size_t)((M + 127) / 128) * 128 * tc::pad16((uint32_t)cols); }

struct Carver {
    size_t off = 0;
    size_t take(size_t bytes) { size_t o = off; off += align_up(bytes); return o; }
};

// plan of one GEMM's packed weights: bf16 images + padded fp32 bias inside the packed buffer
struct PackedGemm { int nch, nc, kpad, ks, nslabs; size_t off_w, off_b; };

static inline PackedGemm plan_packed(Carver& c, int N, int K) {
    PackedGemm g{};
    tc_gemm_pick_nchunk(N, &g.nch, &g.nc);
    g.kpad = (int)tc::pad16((uint32_t)K);
    g.ks = tc_gemm_pick_ks(g.kpad);
    g.nslabs = g.kpad / g.ks;
    g.off_w = c.take((size_t)g.nc * g.nch * g.kpad * sizeof(bf16));
    g.off_b = c.take((size_t)g.nc * g.nch * sizeof(float));
    return g;
}

static inline void bind_packed(TcGemm& t, const PackedGemm& g, const char* packed_base) {
    t.Wp = reinterpret_cast<const bf16*>(packed_base + g.off_w);
    t.bias = reinterpret_cast<const float*>(packed_base + g.off_b);
    t.NCH = g.nch;
    t.n_chunks = g.nc;
}


// ---- fused small-channel MLP (tc_mlp.cu) -------------------------------------------------------------
// Persistent kernel with both weight matrices resident in shared memory; the hidden activation stays on chip.
struct TcMlp {
    const float* x; const float* residual; float* out;
    long long M;
    int C, Cpad, hidden, Hpad;   // Cpad = pad16(C), Hpad = pad16(hidden)
    const float* ln_g; const float* ln_b; float eps;
    const bf16* W1p;   // UMMA image [Cpad/8][Hpad][8]   (rows = hidden unit, k = input channel)
    const bf16* W2p;   // UMMA image [Hpad/8][Cpad][8]   (rows = output channel, k = hidden unit)
    const float* b1;   // [Hpad] zero padded
    const float* b2;   // [Cpad] zero padded
    int max_stages;    // depth cap of the raw-tile ring (<= 8)
    int bulk_out;      // output rows are completed in place in the ring slot and leave by bulk copy (set by launch_tc_mlp)
};
bool tc_mlp_supported(int C, int hidden);
int launch_tc_mlp(const TcMlp& t, cudaStream_t st);

// ---- attention core (attn_core.cu) ----------------------------------------------------------------------
// o_nkc == 0: O row-major [Mtok][ldo]; else O in the UMMA-tiled layout with o_nkc k-chunks per tile
int launch_attn_core_bf16(const bf16* qkv, int qkv_fp16, long long ld, int koff, int voff, bf16* O, long long ldo, int o_nkc,
                          const float* table, const WinGeom& g, int nh, int d, cudaStream_t st);

// HMMA attention core for 7x7 windows (attn_frag.cu).  qkv: fp16 rows in window order (row = window*49 + token),
// ld columns = [q heads | k heads | v heads], every head padded to dp = qkvh_dp(d) columns, q pre-scaled by
// d^-1/2 * log2(e).  O: bf16, UMMA-tiled (nh*dp/8 k-chunks per 128-row tile), same row order and head padding.
bool attn_frag_supported(const WinGeom& g, int nh, int d);
int launch_attn_frag(const __half* qkv, int ld, bf16* O, const float* table, const WinGeom& g, int nh, int d, cudaStream_t st);

// Fused window-attention operator for C <= 64, 8 heads, 7x7 windows (wa_fused.cu): LayerNorm, q|k|v projection (tcgen05),
// attention core (HMMA), output projection (tcgen05), bias + residual, window gather / scatter -- one persistent kernel.
// Weight images and biases are the ones window_attn_pack_bf16 writes for the 7x7 (frag) plan.
bool wa_fused_supported(const WinGeom& g, int C, int nh, int d);
int launch_wa_fused(const sf_window_attn_params* p, const WinGeom& g, bool self_attn, const bf16* Wq, const float* bq, const bf16* Wkv,
                    const float* bkv, const bf16* Wo, const float* bo, cudaStream_t st);

// ---- backward pass on tcgen05 (tc_wgrad.cu, tc_bwd.cu) -----------------------------------------------------------------
// Wg[N][K] += G^T A,  bias_grad[N] += column sums of G;  G: [M x N], A: [M x K], both bf16 UMMA-tiled with zero tail rows
int launch_tc_wgrad(const bf16* G, const bf16* A, float* Wg, float* bias_grad, long long M, int N, int K, const char* name, cudaStream_t st);
// General form: the N rows of the product are `nout` stacked weight matrices (N / nout rows each, e.g. [dWq; dWk; dWv]) written
// to separate tensors; G / A may be chunk ranges [kc0, ..) of wider tiled tensors (tile_nkc chunks per tile; 0 = own tensor).
struct TcWgradArgs {
    const bf16* G; const bf16* A; long long M; int N, K;
    int nout; float* Wg[3]; float* bias_grad[3];
    int g_tile_nkc, g_kc0, a_tile_nkc, a_kc0;
};
int launch_tc_wgrad_ex(const TcWgradArgs& a, const char* name, cudaStream_t st);

}  // namespace sf

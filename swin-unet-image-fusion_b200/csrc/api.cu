// C ABI entry points: argument validation and dispatch on `precision`.
#include <atomic>
#include <cstring>
#include <mutex>
#include <map>
#include <string>
#include <vector>
#include "fp32_kernels.cuh"
#include "bf16_kernels.cuh"
#include "bwd_kernels.cuh"

namespace sf {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};   // process-wide: autograd runs backward kernels on its own thread

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int& c = cached[dev & 63];
    if (c == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        c = v;
    }
    return c;
}
void count_launch(int n) { g_launches += n; }

const char* prof_name(const char* fmt, int v) {
    static std::map<std::string, std::string>* pool = nullptr;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (!pool) pool = new std::map<std::string, std::string>();
    char buf[64];
    snprintf(buf, sizeof(buf), fmt, v);
    auto it = pool->find(buf);
    if (it == pool->end()) it = pool->emplace(buf, buf).first;
    return it->second.c_str();
}

struct ProfRecord { const char* name; double flops, bytes; cudaEvent_t e0, e1; };
static std::atomic<bool> g_prof_on{false};
static std::vector<ProfRecord>* g_prof = nullptr;
static std::mutex g_prof_mu;

ProfScope::ProfScope(const char* name, double flops, double bytes, cudaStream_t stream) : rec(-1), st(stream) {
    if (!g_prof_on) return;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return;
    ProfRecord r{name, flops, bytes, nullptr, nullptr};
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
    cudaEventRecord(r.e0, stream);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_prof) g_prof = new std::vector<ProfRecord>();
    g_prof->push_back(r);
    rec = (int)g_prof->size() - 1;
}
ProfScope::~ProfScope() {
    if (rec < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (g_prof && rec < (int)g_prof->size()) cudaEventRecord((*g_prof)[rec].e1, st);
}

static int check_wa(const sf_window_attn_params* p, const char* who) {
    SF_CHECK_ARG(p, "%s: null params", who);
    SF_CHECK_ARG(p->q_src && p->kv_src && p->out && p->wq && p->wk && p->wv && p->wo && p->bo && p->bias_table, "%s: null tensor pointer", who);
    SF_CHECK_ARG(p->B > 0 && p->Hp > 0 && p->Wp > 0 && p->C > 0 && p->num_heads > 0 && p->head_dim > 0 && p->wsh > 0 && p->wsw > 0, "%s: non-positive dimension", who);
    SF_CHECK_ARG(p->Hp % p->wsh == 0 && p->Wp % p->wsw == 0, "%s: map (%d,%d) is not a multiple of the window (%d,%d); run MyPadding first (a006)", who, p->Hp, p->Wp, p->wsh, p->wsw);
    SF_CHECK_ARG((p->ln_q_gamma == nullptr) == (p->ln_q_beta == nullptr) && (p->ln_kv_gamma == nullptr) == (p->ln_kv_beta == nullptr), "%s: LayerNorm gamma/beta must be given together", who);
    SF_CHECK_ARG(p->precision == SF_PREC_FP32 || p->precision == SF_PREC_BF16, "%s: unknown precision %d", who, p->precision);
    return SF_OK;
}

static int check_mlp(const sf_mlp_params* p, const char* who) {
    SF_CHECK_ARG(p, "%s: null params", who);
    SF_CHECK_ARG(p->in && p->out && p->w1 && p->b1 && p->w2 && p->b2, "%s: null tensor pointer", who);
    SF_CHECK_ARG(p->M > 0 && p->C > 0 && p->hidden > 0, "%s: non-positive dimension", who);
    SF_CHECK_ARG((p->ln_gamma == nullptr) == (p->ln_beta == nullptr), "%s: LayerNorm gamma/beta must be given together", who);
    SF_CHECK_ARG(p->precision == SF_PREC_FP32 || p->precision == SF_PREC_BF16, "%s: unknown precision %d", who, p->precision);
    return SF_OK;
}

static int check_patch(const sf_patch_params* p, const char* who) {
    SF_CHECK_ARG(p, "%s: null params", who);
    SF_CHECK_ARG(p->in && p->out && p->w && p->b && p->ln_gamma && p->ln_beta, "%s: null tensor pointer", who);
    SF_CHECK_ARG(p->B > 0 && p->H > 0 && p->W > 0 && p->Cin > 0 && p->Cout > 0 && p->mh > 0 && p->mw > 0, "%s: non-positive dimension", who);
    if (p->encoder) SF_CHECK_ARG(p->H % p->mh == 0 && p->W % p->mw == 0, "%s: map (%d,%d) is not a multiple of the merging size (%d,%d); run MyPadding first (a006)", who, p->H, p->W, p->mh, p->mw);
    SF_CHECK_ARG(p->precision == SF_PREC_FP32 || p->precision == SF_PREC_BF16, "%s: unknown precision %d", who, p->precision);
    return SF_OK;
}

static int check_head(const sf_head_params* p, const char* who) {
    SF_CHECK_ARG(p, "%s: null params", who);
    SF_CHECK_ARG(p->x && p->y && p->out && p->w1 && p->b1 && p->bn_gamma && p->bn_beta && p->running_mean && p->running_var && p->w2 && p->b2, "%s: null tensor pointer", who);
    SF_CHECK_ARG(p->B > 0 && p->H > 0 && p->W > 0, "%s: non-positive dimension", who);
    SF_CHECK_ARG(p->ksize >= 1 && p->ksize <= 7 && (p->ksize & 1), "%s: kernel size %d unsupported (odd, <= 7)", who, p->ksize);
    SF_CHECK_ARG(p->ksize / 2 < p->H && p->ksize / 2 < p->W, "%s: reflect padding needs pad < size", who);
    return SF_OK;
}

}  // namespace sf

using namespace sf;

extern "C" {

int sf_abi_version(void) { return SF_ABI_VERSION; }
const char* sf_last_error(void) { return g_err; }
long long sf_launch_count(void) { return g_launches; }
void sf_reset_launch_count(void) { g_launches = 0; }

int sf_profile_enable(int on) {
    g_prof_on = on != 0;
    return SF_OK;
}

int sf_profile_summary(sf_profile_entry* out, int max_entries) {
    SF_CHECK_ARG(out && max_entries > 0, "sf_profile_summary: bad args");
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_prof) return 0;
    std::map<std::string, sf_profile_entry> agg;
    for (auto& r : *g_prof) {
        float ms = 0.f;
        cudaEventSynchronize(r.e1);
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
        auto& e = agg[r.name];
        if (e.launches == 0) { memset(&e, 0, sizeof(e)); strncpy(e.name, r.name, sizeof(e.name) - 1); }
        e.launches += 1; e.total_ms += ms; e.flops += r.flops; e.bytes += r.bytes;
    }
    g_prof->clear();
    int n = 0;
    for (auto& kv : agg) { if (n < max_entries) out[n++] = kv.second; }
    return n;
}

int sf_layernorm(const float* in, const float* gamma, const float* beta, float* out, long long M, int C, float eps, int act, void* stream) {
    SF_CHECK_ARG(in && gamma && beta && out && M > 0 && C > 0, "sf_layernorm: bad args");
    return launch_layernorm(in, gamma, beta, out, M, C, eps, act, nullptr, as_stream(stream));
}

size_t sf_window_attn_workspace_bytes(const sf_window_attn_params* p) {
    if (check_wa(p, "sf_window_attn_workspace_bytes") != SF_OK) return 0;
    return p->precision == SF_PREC_BF16 ? window_attn_ws_bf16(p) : window_attn_ws_f32(p);
}
int sf_window_attn_fwd(const sf_window_attn_params* p, void* ws, size_t ws_bytes, void* stream) {
    SF_TRY(check_wa(p, "sf_window_attn_fwd"));
    SF_CHECK_ARG(ws || ws_bytes == 0, "sf_window_attn_fwd: null workspace");
    if (p->precision == SF_PREC_BF16) return window_attn_fwd_bf16(p, ws, ws_bytes, as_stream(stream));
    return window_attn_fwd_f32(p, ws, ws_bytes, as_stream(stream));
}
size_t sf_window_attn_packed_bytes(const sf_window_attn_params* p) {
    if (check_wa(p, "sf_window_attn_packed_bytes") != SF_OK || p->precision != SF_PREC_BF16) return 0;
    return window_attn_packed_bytes_bf16(p);
}
int sf_window_attn_pack(const sf_window_attn_params* p, void* packed, size_t bytes, void* stream) {
    SF_TRY(check_wa(p, "sf_window_attn_pack"));
    SF_CHECK_ARG(p->precision == SF_PREC_BF16 && packed, "sf_window_attn_pack: only SF_PREC_BF16 weights are packed");
    return window_attn_pack_bf16(p, packed, bytes, as_stream(stream));
}
size_t sf_mlp_packed_bytes(const sf_mlp_params* p) {
    if (check_mlp(p, "sf_mlp_packed_bytes") != SF_OK || p->precision != SF_PREC_BF16) return 0;
    return mlp_packed_bytes_bf16(p);
}
int sf_mlp_pack(const sf_mlp_params* p, void* packed, size_t bytes, void* stream) {
    SF_TRY(check_mlp(p, "sf_mlp_pack"));
    SF_CHECK_ARG(p->precision == SF_PREC_BF16 && packed, "sf_mlp_pack: only SF_PREC_BF16 weights are packed");
    return mlp_pack_bf16(p, packed, bytes, as_stream(stream));
}
size_t sf_patch_packed_bytes(const sf_patch_params* p) {
    if (check_patch(p, "sf_patch_packed_bytes") != SF_OK || p->precision != SF_PREC_BF16) return 0;
    return patch_packed_bytes_bf16(p);
}
int sf_patch_pack(const sf_patch_params* p, void* packed, size_t bytes, void* stream) {
    SF_TRY(check_patch(p, "sf_patch_pack"));
    SF_CHECK_ARG(p->precision == SF_PREC_BF16 && packed, "sf_patch_pack: only SF_PREC_BF16 weights are packed");
    return patch_pack_bf16(p, packed, bytes, as_stream(stream));
}

size_t sf_window_attn_bwd_workspace_bytes(const sf_window_attn_bwd_params* p) {
    if (!p || check_wa(&p->fwd, "sf_window_attn_bwd_workspace_bytes") != SF_OK) return 0;
    return window_attn_bwd_ws(p);
}
int sf_window_attn_bwd(const sf_window_attn_bwd_params* p, void* ws, size_t ws_bytes, void* stream) {
    SF_CHECK_ARG(p, "sf_window_attn_bwd: null params");
    SF_TRY(check_wa(&p->fwd, "sf_window_attn_bwd"));
    SF_CHECK_ARG(p->gout && p->g_q_src, "sf_window_attn_bwd: gout and g_q_src are required");
    return window_attn_bwd(p, ws, ws_bytes, as_stream(stream));
}

size_t sf_mlp_workspace_bytes(const sf_mlp_params* p) {
    if (check_mlp(p, "sf_mlp_workspace_bytes") != SF_OK) return 0;
    return p->precision == SF_PREC_BF16 ? mlp_ws_bf16(p) : mlp_ws_f32(p);
}
int sf_mlp_fwd(const sf_mlp_params* p, void* ws, size_t ws_bytes, void* stream) {
    SF_TRY(check_mlp(p, "sf_mlp_fwd"));
    if (p->precision == SF_PREC_BF16) return mlp_fwd_bf16(p, ws, ws_bytes, as_stream(stream));
    return mlp_fwd_f32(p, ws, ws_bytes, as_stream(stream));
}
size_t sf_mlp_bwd_workspace_bytes(const sf_mlp_bwd_params* p) {
    if (!p || check_mlp(&p->fwd, "sf_mlp_bwd_workspace_bytes") != SF_OK) return 0;
    return mlp_bwd_ws(p);
}
int sf_mlp_bwd(const sf_mlp_bwd_params* p, void* ws, size_t ws_bytes, void* stream) {
    SF_CHECK_ARG(p, "sf_mlp_bwd: null params");
    SF_TRY(check_mlp(&p->fwd, "sf_mlp_bwd"));
    SF_CHECK_ARG(p->gout && p->g_in, "sf_mlp_bwd: gout and g_in are required");
    return mlp_bwd(p, ws, ws_bytes, as_stream(stream));
}

size_t sf_patch_workspace_bytes(const sf_patch_params* p) {
    if (check_patch(p, "sf_patch_workspace_bytes") != SF_OK) return 0;
    return p->precision == SF_PREC_BF16 ? patch_ws_bf16(p) : patch_ws_f32(p);
}
int sf_patch_fwd(const sf_patch_params* p, void* ws, size_t ws_bytes, void* stream) {
    SF_TRY(check_patch(p, "sf_patch_fwd"));
    if (p->precision == SF_PREC_BF16) return patch_fwd_bf16(p, ws, ws_bytes, as_stream(stream));
    return patch_fwd_f32(p, ws, ws_bytes, as_stream(stream));
}
size_t sf_patch_bwd_workspace_bytes(const sf_patch_bwd_params* p) {
    if (!p || check_patch(&p->fwd, "sf_patch_bwd_workspace_bytes") != SF_OK) return 0;
    return patch_bwd_ws(p);
}
int sf_patch_bwd(const sf_patch_bwd_params* p, void* ws, size_t ws_bytes, void* stream) {
    SF_CHECK_ARG(p, "sf_patch_bwd: null params");
    SF_TRY(check_patch(&p->fwd, "sf_patch_bwd"));
    SF_CHECK_ARG(p->gout && p->g_in, "sf_patch_bwd: gout and g_in are required");
    return patch_bwd(p, ws, ws_bytes, as_stream(stream));
}

size_t sf_head_workspace_bytes(const sf_head_params* p) {
    if (check_head(p, "sf_head_workspace_bytes") != SF_OK) return 0;
    return head_ws(p);
}
int sf_head_fwd(const sf_head_params* p, void* ws, size_t ws_bytes, void* stream) {
    SF_TRY(check_head(p, "sf_head_fwd"));
    SF_CHECK_ARG(!p->training || (p->save_mean && p->save_invstd), "sf_head_fwd: training needs save_mean / save_invstd");
    return head_fwd(p, ws, ws_bytes, as_stream(stream));
}
size_t sf_head_bwd_workspace_bytes(const sf_head_bwd_params* p) {
    if (!p || check_head(&p->fwd, "sf_head_bwd_workspace_bytes") != SF_OK) return 0;
    return head_bwd_ws(p);
}
int sf_head_bwd(const sf_head_bwd_params* p, void* ws, size_t ws_bytes, void* stream) {
    SF_CHECK_ARG(p, "sf_head_bwd: null params");
    SF_TRY(check_head(&p->fwd, "sf_head_bwd"));
    SF_CHECK_ARG(p->gout && p->g_x && p->g_y, "sf_head_bwd: gout, g_x and g_y are required");
    return head_bwd(p, ws, ws_bytes, as_stream(stream));
}

}  // extern "C"

// Attention core of the bf16 path for the shapes the window-order fragment kernels (attn_frag.cu) do not cover:
// dispatch over the per-shape instantiations of attn_core_impl.cuh (built in three translation units).
#include "attn_core_impl.cuh"

namespace sf {

int attn_core_dispatch_a(const AttnArgs& a, cudaStream_t st) {
    const int d = a.d;
    if (d <= 4) return launch_attn_small<4, 49>(a, st);
    if (d <= 8) return launch_attn_small<8, 49>(a, st);
    if (d <= 12) return launch_attn_small<12, 49>(a, st);
    return launch_attn_small<16, 49>(a, st);
}

int launch_attn_core_bf16(const bf16* qkv, int qkv_fp16, long long ld, int koff, int voff, bf16* O, long long ldo, int o_nkc,
                          const float* table, const WinGeom& g, int nh, int d, cudaStream_t st) {
    AttnArgs a{};
    a.qkv = qkv; a.qkv_fp16 = qkv_fp16; a.ld = ld; a.koff = koff; a.voff = voff; a.O = O; a.ldo = ldo; a.o_nkc = o_nkc; a.table = table;
    a.g = g; a.nh = nh; a.d = d; a.dp = (d + 3) & ~3;
    a.nwin = (long long)g.B * g.nWh * g.nWw;
    a.scale_log2e = 1.4426950408889634f / sqrtf((float)d);
    SF_CHECK_ARG(a.nwin <= 2147483647LL, "attention core: too many windows");
    if (d > 64) {
        set_error("attention core: head_dim %d > 64 is not supported", d);
        return SF_ERR_UNSUPPORTED;
    }
    if (g.T == 49 && d <= 16) return attn_core_dispatch_a(a, st);
    if (g.T == 49) return attn_core_dispatch_b(a, st);
    return attn_core_dispatch_c(a, st);
}

}  // namespace sf

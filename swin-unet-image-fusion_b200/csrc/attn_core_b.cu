// attn_core_impl.cuh instantiations: 7x7 windows, head_dim 17..32 (wider heads: attn_core_d.cu, attn_core_e.cu)
#include "attn_core_impl.cuh"

namespace sf {

int attn_core_dispatch_b(const AttnArgs& a, cudaStream_t st) {
    const int d = a.d;
    if (d <= 24) return launch_attn_small<24, 49>(a, st);
    if (d <= 32) return launch_attn_small<32, 49>(a, st);
    if (d <= 48) return attn_core_dispatch_d(a, st);
    return attn_core_dispatch_e(a, st);
}

}  // namespace sf

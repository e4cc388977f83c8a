// attn_core_impl.cuh instantiations: 8x8 windows (one-thread-per-row kernel) and the general kernel for any other window
#include "attn_core_impl.cuh"

namespace sf {

int attn_core_dispatch_c(const AttnArgs& a, cudaStream_t st) {
    const int d = a.d;
    if (a.g.T == 64 && d <= 4) return launch_attn_small<4, 64>(a, st);
    if (a.g.T == 64 && d <= 8) return launch_attn_small<8, 64>(a, st);
    if (a.g.T == 64 && d <= 16) return launch_attn_small<16, 64>(a, st);
    if (d <= 16) return launch_attn_t<16, 1, false>(a, st);
    if (d <= 32) return launch_attn_t<32, 1, false>(a, st);
    return launch_attn_t<64, 1, false>(a, st);
}

}  // namespace sf

// attn_core_impl.cuh instantiation: 7x7 windows, head_dim 49..64
#include "attn_core_impl.cuh"

namespace sf {
int attn_core_dispatch_e(const AttnArgs& a, cudaStream_t st) { return launch_attn_small<64, 49>(a, st); }
}  // namespace sf

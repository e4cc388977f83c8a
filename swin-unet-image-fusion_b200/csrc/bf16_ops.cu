// bf16 tensor-core path -- placeholder until the tcgen05 kernels land: every entry reports
// SF_ERR_UNSUPPORTED (an error, never a silent fallback to fp32).
#include "bf16_kernels.cuh"
namespace sf {
size_t window_attn_ws_bf16(const sf_window_attn_params*) { return 0; }
int window_attn_fwd_bf16(const sf_window_attn_params*, void*, size_t, cudaStream_t) { set_error("bf16 window attention is not built"); return SF_ERR_UNSUPPORTED; }
size_t mlp_ws_bf16(const sf_mlp_params*) { return 0; }
int mlp_fwd_bf16(const sf_mlp_params*, void*, size_t, cudaStream_t) { set_error("bf16 MLP is not built"); return SF_ERR_UNSUPPORTED; }
size_t patch_ws_bf16(const sf_patch_params*) { return 0; }
int patch_fwd_bf16(const sf_patch_params*, void*, size_t, cudaStream_t) { set_error("bf16 patch layer is not built"); return SF_ERR_UNSUPPORTED; }
}  // namespace sf

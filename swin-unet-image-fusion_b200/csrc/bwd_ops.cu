// Backward kernels -- placeholder: every entry reports SF_ERR_UNSUPPORTED.
#include "bwd_kernels.cuh"
namespace sf {
size_t window_attn_bwd_ws(const sf_window_attn_bwd_params*) { return 0; }
int window_attn_bwd(const sf_window_attn_bwd_params*, void*, size_t, cudaStream_t) { set_error("window attention backward is not built"); return SF_ERR_UNSUPPORTED; }
size_t mlp_bwd_ws(const sf_mlp_bwd_params*) { return 0; }
int mlp_bwd(const sf_mlp_bwd_params*, void*, size_t, cudaStream_t) { set_error("MLP backward is not built"); return SF_ERR_UNSUPPORTED; }
size_t patch_bwd_ws(const sf_patch_bwd_params*) { return 0; }
int patch_bwd(const sf_patch_bwd_params*, void*, size_t, cudaStream_t) { set_error("patch layer backward is not built"); return SF_ERR_UNSUPPORTED; }
size_t head_bwd_ws(const sf_head_bwd_params*) { return 0; }
int head_bwd(const sf_head_bwd_params*, void*, size_t, cudaStream_t) { set_error("head backward is not built"); return SF_ERR_UNSUPPORTED; }
}  // namespace sf

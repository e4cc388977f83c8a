// bf16 tensor-core (tcgen05) implementations of the fused operators.
#pragma once
#include "common.cuh"

namespace sf {
size_t window_attn_ws_bf16(const sf_window_attn_params* p);
int window_attn_fwd_bf16(const sf_window_attn_params* p, void* ws, size_t ws_bytes, cudaStream_t st);
size_t mlp_ws_bf16(const sf_mlp_params* p);
int mlp_fwd_bf16(const sf_mlp_params* p, void* ws, size_t ws_bytes, cudaStream_t st);
size_t patch_ws_bf16(const sf_patch_params* p);
int patch_fwd_bf16(const sf_patch_params* p, void* ws, size_t ws_bytes, cudaStream_t st);
}  // namespace sf

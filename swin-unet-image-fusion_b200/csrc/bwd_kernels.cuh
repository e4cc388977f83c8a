// Backward (gradient) implementations of the fused operators.
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace sf {
size_t window_attn_bwd_ws(const sf_window_attn_bwd_params* p);
int window_attn_bwd(const sf_window_attn_bwd_params* p, void* ws, size_t ws_bytes, cudaStream_t st);
size_t mlp_bwd_ws(const sf_mlp_bwd_params* p);
int mlp_bwd(const sf_mlp_bwd_params* p, void* ws, size_t ws_bytes, cudaStream_t st);
size_t patch_bwd_ws(const sf_patch_bwd_params* p);
int patch_bwd(const sf_patch_bwd_params* p, void* ws, size_t ws_bytes, cudaStream_t st);
size_t head_bwd_ws(const sf_head_bwd_params* p);
int head_bwd(const sf_head_bwd_params* p, void* ws, size_t ws_bytes, cudaStream_t st);
// attn_bwd_mma.cu: tensor-core attention-core backward (7x7 windows, head_dim 3 / 6 / 12)
bool attn_core_bwd_mma_supported(const WinGeom& g, int d, int nh);
// Optional outputs as bf16 UMMA-tiled tensors (instead of fp32 rows): dQ | dK | dV as the column blocks [0, inner), [inner, 2 inner),
// [2 inner, 3 inner) of one tensor with nkc3 chunks per 128-row tile, O in a tensor with nkc1 chunks per tile.  Rows are token
// rows of the un-shifted map.  The caller zeroes the tail rows of the last tile.
struct AttnBwdTiledOut { __nv_bfloat16* dqkv; __nv_bfloat16* o; unsigned nkc3, nkc1; int inner; };
// O (optional): also writes the forward output P V, so that the caller need not recompute the attention core
int launch_attn_core_bwd_mma(const float* Q, const float* K, const float* V, const float* gO, float* dQ, float* dK, float* dV, float* O,
                             const float* table, float* gtable, const WinGeom& g, int inner, int nh, int d, cudaStream_t st, int ld = 0,
                             const AttnBwdTiledOut* tiled = nullptr);
// gemm_tf32.cu: TF32 tensor-core GEMMs for the backward pass of SF_PREC_BF16 operators
struct GemmBatch;
int gemm_tf32_nn(const float* A, const float* B, const float* aux, float* C, long long M, int N, int K, bool accum, cudaStream_t st);
int gemm_tf32_nt(const GemmBatch& batch, int nbatch, long long M, int N, int K, cudaStream_t st);
int gemm_tf32_wgrad(const float* G, const float* A, float* Wg, float* bias_grad, long long M, int N, int K, bool elu_a, cudaStream_t st);
}  // namespace sf

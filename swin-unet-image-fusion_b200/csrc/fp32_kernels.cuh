// Declarations shared by the fp32 forward / backward translation units.
#pragma once
#include "common.cuh"

namespace sf {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct GemmProblem {
    const float* A;         // [M][K]
    const float* W;         // [N][K]
    const float* bias;      // [N] or null
    const float* residual;  // [M][N] or null
    float* C;               // [M][N]
};
struct GemmBatch { GemmProblem p[4]; };

// geometry of the "anti patch merging" scatter (a011:111-117): coarse map (Hc,Wc), factors, Cout
struct UnmergeGeom { int Hc, Wc, mh, mw, Cout; };

int launch_layernorm(const float* in, const float* gamma, const float* beta, float* out, long long M, int C, float eps,
                     int act, const UnmergeGeom* ug, cudaStream_t st);
int launch_gemm_tn(const GemmBatch& batch, int nbatch, long long M, int N, int K, bool elu, cudaStream_t st);
int launch_attn_core_f32(const float* Q, const float* K, const float* V, float* O, const float* table, const WinGeom& g,
                         int inner, int nh, int d, cudaStream_t st);

size_t window_attn_ws_f32(const sf_window_attn_params* p);
int window_attn_fwd_f32(const sf_window_attn_params* p, void* ws, size_t ws_bytes, cudaStream_t st);
size_t mlp_ws_f32(const sf_mlp_params* p);
int mlp_fwd_f32(const sf_mlp_params* p, void* ws, size_t ws_bytes, cudaStream_t st);
size_t patch_ws_f32(const sf_patch_params* p);
int patch_fwd_f32(const sf_patch_params* p, void* ws, size_t ws_bytes, cudaStream_t st);
size_t head_ws(const sf_head_params* p);
int head_fwd(const sf_head_params* p, void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace sf

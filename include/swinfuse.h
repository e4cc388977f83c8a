/*
 * swinfuse.h -- C ABI of libswinfuse.so: the B200 (sm_100a) kernels behind the Swin-UNet
 * fusion hot path of RainbowZL0/swin-unet-image-fusion.
 *
 * The reference has no FFI layer: its operator API is the nn.Module surface consumed by
 * a016_train.py / a017_test.py.  Each entry point below replaces the ATen call sequence of
 * one reference module `forward` (cited as aNNN:line = /root/reference/aNNN_*.py).  The
 * Python drop-in modules (swin-unet-image-fusion_b200/dropin/aNNN_*.py) bind these through
 * ctypes; see INTEGRATION.md for the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - Plain C: raw device pointers, ints, POD structs.  No torch / C++ types.
 *   - Feature maps are fp32, channels-last: T[b][row][col][ch] contiguous ("NHWC").  A
 *     PyTorch (B,C,H,W) tensor in torch.channels_last memory format has exactly this
 *     layout; sf_nchw_to_nhwc / sf_nhwc_to_nchw convert contiguous NCHW tensors.
 *   - Weights are passed exactly as the reference stores them (nn.Linear weight [out][in],
 *     1x1 nn.Conv2d weight [out][in][1][1] == [out][in]), fp32, contiguous.
 *   - The caller owns every buffer, including the workspace (size from the matching
 *     *_workspace_bytes).  The library never allocates, frees or synchronises; every call
 *     only enqueues kernels on `stream` (a cudaStream_t passed as void*) and is CUDA-graph
 *     capture safe.
 *   - Return value: 0 on success, negative sf_status otherwise; sf_last_error() returns a
 *     thread-local message.  Unsupported shapes are errors, never fallbacks.
 *   - `precision`: SF_PREC_FP32 = fp32 FFMA arithmetic (<=1e-4 rel. vs the reference);
 *     SF_PREC_BF16 = bf16 tensor-core operands with fp32 accumulation, fp32 LayerNorm /
 *     softmax / residual stream (<=2e-2 rel.).
 */
#ifndef SWINFUSE_H
#define SWINFUSE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SF_ABI_VERSION 1

typedef enum {
    SF_OK = 0,
    SF_ERR_INVALID = -1,      /* bad argument / unsupported shape */
    SF_ERR_WORKSPACE = -2,    /* workspace too small */
    SF_ERR_CUDA = -3,         /* a CUDA runtime call failed */
    SF_ERR_UNSUPPORTED = -4   /* feature not built */
} sf_status;

typedef enum { SF_PREC_FP32 = 0, SF_PREC_BF16 = 1 } sf_precision;

int sf_abi_version(void);
const char* sf_last_error(void);
/* number of kernels this library has enqueued from the calling thread since the last reset
 * (bench.py's `gpu_launches`). */
long long sf_launch_count(void);
void sf_reset_launch_count(void);

/* Live per-kernel timing for roofline accounting (bench.py).  While enabled, every kernel the
 * library enqueues from the calling thread (outside stream capture) is bracketed by a CUDA-event
 * pair on its stream.  sf_profile_summary synchronises those events, aggregates by kernel name
 * (launch count, summed device time, summed ALGORITHMIC flops and bytes as stated by each
 * launcher), clears the records and returns the number of entries written. */
typedef struct {
    char name[48];
    long long launches;
    double total_ms;
    double flops;
    double bytes;
} sf_profile_entry;
int sf_profile_enable(int on);
int sf_profile_summary(sf_profile_entry* out, int max_entries);

/* ---------------------------------------------------------------------------------------
 * Index / permutation kernels (bit-exact)
 * ------------------------------------------------------------------------------------- */

/* contiguous NCHW <-> NHWC */
int sf_nchw_to_nhwc(const float* in, float* out, int B, int C, int H, int W, void* stream);
int sf_nhwc_to_nchw(const float* in, float* out, int B, int C, int H, int W, void* stream);

/* Inference edges of a017_test.py, batched on the device.
 * sf_bgr_to_ycrcb: a015_dataset.py:86-93 (cv2.cvtColor(vis, COLOR_BGR2YCrCb) on uint8, OpenCV's 14-bit fixed point,
 *   bit-exact) + a015:56-60 (v2.ToImage, v2.ToDtype(float32, scale=True): u8 * fl32(1/255)) + the split of
 *   a017:68.  bgr (B,H,W,3) uint8 -> y (B,1,H,W), crcb (B,2,H,W) fp32.
 * sf_ycrcb_to_rgb: a017:83-88 (clamp_(fus_y,0,1); concat with CrCb; cv2.cvtColor(float32, COLOR_YCrCb2RGB)).
 *   fus_y (B,1,H,W), crcb (B,2,H,W) -> rgb (B,3,H,W) fp32, bit-exact with OpenCV's fused-multiply-add path. */
int sf_bgr_to_ycrcb(const uint8_t* bgr, float* y, float* crcb, int B, int H, int W, void* stream);
int sf_ycrcb_to_rgb(const float* fus_y, const float* crcb, float* rgb, int B, int H, int W, void* stream);

/* MyPadding encoder branch, a006:111-131: reflect pad bottom/right.
 * in (B,H,W,C) -> out (B,H+pad_down,W+pad_right,C); out[L+k] = in[L-2-k]; needs pad < L. */
int sf_pad_reflect(const float* in, float* out, int B, int H, int W, int C, int pad_down, int pad_right,
                   void* stream);
/* adjoint of sf_pad_reflect (autograd of a006:128-131): gin = crop(gout) + mirrored rows/cols */
int sf_pad_reflect_bwd(const float* gout, float* gin, int B, int H, int W, int C, int pad_down, int pad_right,
                       void* stream);

/* MyPadding decoder branch, a006:133-146: in (B,H,W,C) -> out (B,H-crop_down,W-crop_right,C).
 * If `add` != NULL (same shape as out) it is added: the U-Net skip `x += history_x`,
 * a013:222-225, happens right after the crop of the previous decoder stage. */
int sf_crop(const float* in, const float* add, float* out, int B, int H, int W, int C, int crop_down,
            int crop_right, void* stream);
/* adjoint of the crop: zero-extends gout (B,H-cd,W-cr,C) to gin (B,H,W,C) */
int sf_crop_bwd(const float* gout, float* gin, int B, int H, int W, int C, int crop_down, int crop_right,
                void* stream);

/* a011:87-93  b c (H ph)(W pw) -> b (ph pw c) H W, channels-last:
 * out[b][Y][X][(ph*mw+pw)*C+c] = in[b][Y*mh+ph][X*mw+pw][c];  in (B,H,W,C), H%mh==0, W%mw==0 */
int sf_patch_merge(const float* in, float* out, int B, int H, int W, int C, int mh, int mw, void* stream);
/* a011:111-117, inverse of the above: in (B,H,W,mh*mw*C) -> out (B,H*mh,W*mw,C) */
int sf_patch_unmerge(const float* in, float* out, int B, int H, int W, int C, int mh, int mw, void* stream);

/* a001:165-172 (+ a001:442-445 when shift!=0): window partition of the (cyclically shifted) map.
 * in (B,Hp,Wp,C) -> out (B*nWh*nWw, wsh*wsw, C); shifted[r][c] = in[(r+wsh/2)%Hp][(c+wsw/2)%Wp] */
int sf_window_partition(const float* in, float* out, int B, int Hp, int Wp, int C, int wsh, int wsw, int shift,
                        void* stream);
/* a001:390-398 (+ a001:471-473): inverse scatter */
int sf_window_reverse(const float* in, float* out, int B, int Hp, int Wp, int C, int wsh, int wsw, int shift,
                      void* stream);
/* a001:217-272: the (nW, t, t) shift mask as bytes (1 = masked), for tests. */
int sf_shift_mask(uint8_t* out, int Hp, int Wp, int wsh, int wsw, void* stream);
/* a001:127-144: the (t, t) relative-position bias gathered from the (2wsh-1, 2wsw-1) table. */
int sf_relative_position_bias(const float* table, float* out, int wsh, int wsw, void* stream);

/* ---------------------------------------------------------------------------------------
 * Fused floating-point operators
 * ------------------------------------------------------------------------------------- */

/* my_layer_norm, a004:54-72: LN over C (eps, affine) of each of M = B*H*W tokens.
 * act != 0 applies ELU(alpha=1) afterwards. */
int sf_layernorm(const float* in, const float* gamma, const float* beta, float* out, long long M, int C,
                 float eps, int act, void* stream);

/* WindowAttention.forward(q, k, v) with k is v, a001:448-474, optionally fused with the
 * pre-norm and residual of AddAndLayerNormWithOtherModule (a004:29-38):
 *   out = [residual +] Proj( softmax( (Q K^T) d^-1/2 + bias, mask ) V )
 *   Q = LNq?(q_src) Wq^T + bq,  K = LNkv?(kv_src) Wk^T + bk,  V = LNkv?(kv_src) Wv^T + bv
 * Shift, window partition, head split, window reverse and un-shift are index math. */
typedef struct {
    const float* q_src;    /* (B,Hp,Wp,C) */
    const float* kv_src;   /* (B,Hp,Wp,C); == q_src for self attention */
    const float* residual; /* (B,Hp,Wp,C) or NULL */
    float* out;            /* (B,Hp,Wp,C) */
    const float* ln_q_gamma;  const float* ln_q_beta;    /* (C) or NULL: no LN on q_src */
    const float* ln_kv_gamma; const float* ln_kv_beta;   /* (C) or NULL */
    const float* wq; const float* bq;   /* (nh*d, C), (nh*d) or NULL bias */
    const float* wk; const float* bk;
    const float* wv; const float* bv;
    const float* wo; const float* bo;   /* (C, nh*d), (C) */
    const float* bias_table;            /* (2wsh-1, 2wsw-1) */
    int B, Hp, Wp, C, num_heads, head_dim, wsh, wsw;
    int shift;        /* use_cyclic_shift */
    float ln_eps;
    int precision;    /* sf_precision */
    const void* packed; /* optional: weights pre-packed by sf_window_attn_pack (SF_PREC_BF16); NULL = pack per call */
} sf_window_attn_params;
/* SF_PREC_BF16 consumes the weights as bf16 tensor-core operand images.  Packing them is pure data
 * movement that only depends on the weights, so a caller whose weights do not change between calls
 * (inference) packs once into a buffer it owns and passes it in `packed`. */
size_t sf_window_attn_packed_bytes(const sf_window_attn_params* p);
int sf_window_attn_pack(const sf_window_attn_params* p, void* packed, size_t packed_bytes, void* stream);
size_t sf_window_attn_workspace_bytes(const sf_window_attn_params* p);
int sf_window_attn_fwd(const sf_window_attn_params* p, void* workspace, size_t workspace_bytes, void* stream);

/* Gradients of the above (autograd of a001:448-474 + a004:29-38; row a18 of SURVEY section 8).
 * Recomputes the forward intermediates from q_src / kv_src.  All g* outputs are WRITTEN
 * (not accumulated) except where noted; NULL outputs are skipped. */
typedef struct {
    sf_window_attn_params fwd;
    const float* gout;        /* (B,Hp,Wp,C): gradient w.r.t. out */
    float* g_q_src;           /* (B,Hp,Wp,C) */
    float* g_kv_src;          /* (B,Hp,Wp,C); for self attention pass NULL: it is folded into g_q_src */
    float* g_ln_q_gamma;  float* g_ln_q_beta;
    float* g_ln_kv_gamma; float* g_ln_kv_beta;
    float* g_wq; float* g_bq; float* g_wk; float* g_bk; float* g_wv; float* g_bv; float* g_wo; float* g_bo;
    float* g_bias_table;
    const float* add_to_g_q_src;  /* optional (B,Hp,Wp,C): added to g_q_src -- pass gout when fwd.residual was q_src itself
                                     (x + Attn(LN(x)), a004:29-38), so the two gradient branches of x leave as one tensor */
} sf_window_attn_bwd_params;
size_t sf_window_attn_bwd_workspace_bytes(const sf_window_attn_bwd_params* p);
int sf_window_attn_bwd(const sf_window_attn_bwd_params* p, void* workspace, size_t workspace_bytes, void* stream);

/* AutoPathMLP.sequence_{x,y}, a003:21-31, optionally fused with a004:29-38:
 *   out = [residual +] W2 ELU( W1 LN?(in) + b1 ) + b2      (1x1 convs == per-token linears) */
typedef struct {
    const float* in;        /* (M, C) */
    const float* residual;  /* (M, C) or NULL */
    float* out;             /* (M, C) */
    const float* ln_gamma; const float* ln_beta;  /* (C) or NULL */
    const float* w1; const float* b1;   /* (hidden, C), (hidden) */
    const float* w2; const float* b2;   /* (C, hidden), (C) */
    long long M;
    int C, hidden;
    float ln_eps;
    int precision;
    const void* packed;     /* optional, see sf_window_attn_params.packed */
} sf_mlp_params;
size_t sf_mlp_packed_bytes(const sf_mlp_params* p);
int sf_mlp_pack(const sf_mlp_params* p, void* packed, size_t packed_bytes, void* stream);
size_t sf_mlp_workspace_bytes(const sf_mlp_params* p);
int sf_mlp_fwd(const sf_mlp_params* p, void* workspace, size_t workspace_bytes, void* stream);

typedef struct {
    sf_mlp_params fwd;
    const float* gout;
    float* g_in;
    float* g_ln_gamma; float* g_ln_beta;
    float* g_w1; float* g_b1; float* g_w2; float* g_b2;
    const float* add_to_g_in;     /* optional (M,C): added to g_in (gout when fwd.residual was the input itself) */
} sf_mlp_bwd_params;
size_t sf_mlp_bwd_workspace_bytes(const sf_mlp_bwd_params* p);
int sf_mlp_bwd(const sf_mlp_bwd_params* p, void* workspace, size_t workspace_bytes, void* stream);

/* PatchMergingAndLinearLayer, a011:236-264, one path.
 * encoder != 0:  in (B,H,W,Cin) -> merge (mh,mw) -> conv1x1 (mh*mw*Cin -> Cout) -> LN(Cout) -> ELU
 *                out (B,H/mh,W/mw,Cout)
 * encoder == 0:  in (B,H,W,Cin) -> conv1x1 (Cin -> mh*mw*Cout) -> LN(mh*mw*Cout) -> unmerge -> ELU
 *                out (B,H*mh,W*mw,Cout)            ("anti patch merging") */
typedef struct {
    const float* in;
    float* out;
    const float* w; const float* b;               /* conv1x1 weight (out_ch, in_ch), bias */
    const float* ln_gamma; const float* ln_beta;  /* (conv out_ch) */
    int B, H, W, Cin, Cout, mh, mw;
    int encoder;
    float ln_eps;
    int precision;
    const void* packed;     /* optional, see sf_window_attn_params.packed */
} sf_patch_params;
size_t sf_patch_packed_bytes(const sf_patch_params* p);
int sf_patch_pack(const sf_patch_params* p, void* packed, size_t packed_bytes, void* stream);
size_t sf_patch_workspace_bytes(const sf_patch_params* p);
int sf_patch_fwd(const sf_patch_params* p, void* workspace, size_t workspace_bytes, void* stream);

typedef struct {
    sf_patch_params fwd;
    const float* gout;
    float* g_in;
    float* g_w; float* g_b; float* g_ln_gamma; float* g_ln_beta;
} sf_patch_bwd_params;
size_t sf_patch_bwd_workspace_bytes(const sf_patch_bwd_params* p);
int sf_patch_bwd(const sf_patch_bwd_params* p, void* workspace, size_t workspace_bytes, void* stream);

/* MyModel.do_final_layer, a013:126-152:
 *   cat(x,y) -> Conv2d(2,2,k,reflect) -> BatchNorm2d(2) -> ELU -> Conv2d(2,1,k,reflect)
 * x, y, out: (B,H,W) single-channel maps.  training != 0: batch statistics are used and
 * running_mean / running_var are updated in place (momentum, unbiased variance) exactly as
 * nn.BatchNorm2d does; save_mean / save_invstd (2 floats each) receive the batch statistics
 * for the backward pass. */
typedef struct {
    const float* x; const float* y;
    float* out;
    const float* w1; const float* b1;   /* (2,2,k,k), (2) */
    const float* bn_gamma; const float* bn_beta;
    float* running_mean; float* running_var;
    float* save_mean; float* save_invstd;   /* may be NULL when training == 0 */
    const float* w2; const float* b2;   /* (1,2,k,k), (1) */
    int B, H, W, ksize;
    int training;
    float bn_eps, bn_momentum;
} sf_head_params;
size_t sf_head_workspace_bytes(const sf_head_params* p);
int sf_head_fwd(const sf_head_params* p, void* workspace, size_t workspace_bytes, void* stream);

typedef struct {
    sf_head_params fwd;
    const float* gout;
    float* g_x; float* g_y;
    float* g_w1; float* g_b1; float* g_bn_gamma; float* g_bn_beta; float* g_w2; float* g_b2;
} sf_head_bwd_params;
size_t sf_head_bwd_workspace_bytes(const sf_head_bwd_params* p);
int sf_head_bwd(const sf_head_bwd_params* p, void* workspace, size_t workspace_bytes, void* stream);

/* One Adam step (torch.optim.Adam semantics, a016:67: no weight decay, no amsgrad) over flat fp32
 * buffers of n elements: g' = grad * grad_scale (1/world_size after the data-parallel all-reduce);
 * m = b1 m + (1-b1) g'; v = b2 v + (1-b2) g'^2; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps).
 * The scalar hyper-parameters are doubles (as the Python optimizer holds them): 1-b, the bias corrections and the
 * step size are formed in double and rounded to fp32 once, which is what makes the result match torch to 1e-6. */
int sf_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, double lr, double beta1,
                 double beta2, double eps, int step, double grad_scale, void* stream);

/* Fusion loss of a008_loss.py (MyLoss.calcu_total_loss, a008:226-282, constants A000_CONFIG.py:34-52) and its
 * gradient w.r.t. the fused image, SURVEY 8(a) row a19.  fusion / ir / vis: contiguous (B,1,H,W) fp32.
 *   loss[0] = r_ssim * loss[1] + r_texture * loss[2] + r_intensity * loss[3]
 *   loss[1] = ssim_scale * [w_ir MS(f, ir) + (1 - w_ir) MS(f, vis)]   (kornia MS_SSIMLoss() defaults, a008:24,109-110)
 *   loss[2] = texture_scale * mean |Sobel(f) - max(Sobel(ir), Sobel(vis))|   (kornia Sobel() defaults, a008:37,187-198)
 *   loss[3] = intensity_scale * mean |f - max(ir, vis)|                      (a008:218-224)
 * f = clamp(fusion, 0, 1) when clamp01 != 0 (a016:153 clamps the model output before the loss; its gradient mask is
 * then applied to g_fusion).  g_fusion (B,1,H,W) receives d loss[0] / d fusion; NULL = value only.
 * The kornia arithmetic is restated (kornia is not vendored by the reference): see oracle/kornia_restatement.py. */
typedef struct {
    const float* fusion; const float* ir; const float* vis;
    float* loss;        /* 4 floats */
    float* total;       /* or NULL: a second copy of loss[0] (lets the caller own the scalar as its own buffer) */
    float* g_fusion;    /* or NULL */
    int B, H, W;
    int clamp01;
    float w_ir, ssim_scale, texture_scale, intensity_scale, r_ssim, r_texture, r_intensity;
} sf_fusion_loss_params;
size_t sf_fusion_loss_workspace_bytes(const sf_fusion_loss_params* p);
int sf_fusion_loss(const sf_fusion_loss_params* p, void* workspace, size_t workspace_bytes, void* stream);

/* out[i] = in[i] * scalar[0]  (chain rule with the upstream gradient of the scalar loss, read on the device) */
int sf_scale_by_scalar(const float* in, const float* scalar, float* out, long long n, void* stream);

/* out[i] = a[i] + b[i]  (U-Net skip when no crop precedes it) */
int sf_add(const float* a, const float* b, float* out, long long n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SWINFUSE_H */

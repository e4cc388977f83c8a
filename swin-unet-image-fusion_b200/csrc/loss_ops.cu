// Fusion loss of a008_loss.py and its gradient w.r.t. the fused image, on the device:
//   L = r_s * ssim_scale * [w_ir MS(f, ir) + (1 - w_ir) MS(f, vis)]          a008:89-131, 226-282, A000_CONFIG.py:34-52
//     + r_t * texture_scale * mean |Sobel(f) - max(Sobel(ir), Sobel(vis))|    a008:161-199
//     + r_i * intensity_scale * mean |f - max(ir, vis)|                       a008:202-224
// with f = clamp(x, 0, 1) when `clamp01` is set (a016:153).  MS(.,.) is kornia's MS_SSIMLoss with default
// arguments (a008:24) and Sobel is kornia.filters.Sobel with default arguments (a008:37), restated in
// oracle/kornia_restatement.py (kornia is a third-party dependency the reference does not pin or vendor).
//
// The kornia formulation is fifteen dense 33x33 windows per call; here (identical mathematics)
//   * the Gaussian windows are applied separably (vertical pass to HBM planes, horizontal pass from shared memory),
//   * the three duplicated copies of every sigma are computed once (lm = l^3, PIcs = (prod cs)^3),
//   * mu_f and E[f^2] are shared by the IR and the visible call,
//   * taps whose weight underflows fp32 relative to the centre are skipped (sigma 0.5: radius 4, sigma 1: radius 8).
// The gradient uses the self-adjointness of a zero-padded symmetric blur: per sigma four per-pixel gradient
// planes (d/d mu_f, d/d E[f^2], d/d E[f ir], d/d E[f vis]) are blurred again (two passes) and combined with f, ir, vis.
// Every reduction is a two-stage deterministic sum (per-block partials, one final block in double precision).
#include "common.cuh"

namespace sf {

static constexpr int NS = 5, RAD = 16, TAPS = 33;
static constexpr int NV = 42;     // vertical-pass planes: 8 moments x 5 sigmas + |f-ir|, |f-vis| at sigma 8
static constexpr int NG = 20;     // gradient planes: 5 sigmas x {mu_f, E[f^2], E[f ir], E[f vis]}
static constexpr float MS_C1 = 0.01f * 0.01f, MS_C2 = 0.03f * 0.03f, MS_ALPHA = 0.025f, MS_COMP = 200.0f;
static constexpr float SOBEL_EPS = 1e-6f;

struct GaussW { float w[NS][TAPS]; };   // travels as a kernel parameter: unrolled taps become constant-bank operands

__host__ __device__ constexpr int tap_radius(int s) { return s == 0 ? 4 : (s == 1 ? 8 : RAD); }

static GaussW make_gauss() {
    const float sig[NS] = {0.5f, 1.0f, 2.0f, 4.0f, 8.0f};
    GaussW g;
    for (int s = 0; s < NS; s++) {
        float sum = 0.f;
        for (int t = 0; t < TAPS; t++) {
            const float c = (float)(t - RAD);
            g.w[s][t] = expf(-(c * c) / (float)(2.0 * (double)sig[s] * (double)sig[s]));
            sum += g.w[s][t];
        }
        for (int t = 0; t < TAPS; t++) g.w[s][t] /= sum;
    }
    return g;
}

__device__ __forceinline__ float clamp_if(float v, int on) { return on ? fminf(fmaxf(v, 0.f), 1.f) : v; }
__device__ __forceinline__ float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

// sum over the block, result valid in thread 0; NW = warps per block
template <int NW>
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float r = 0.f;
    if (threadIdx.x == 0)
        for (int i = 0; i < NW; i++) r += red[i];
    return r;
}

// ---------------------------------------------------------------------------------------------------------
// K0: Sobel magnitudes (replicate border, normalised 3x3 kernels), texture / intensity partial sums, and the
//     per-pixel texture gradient w.r.t. the two Sobel responses of f (tx, ty).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void sobel3(const float* __restrict__ p, long long base, int y, int x, int H, int W, int cl, float& gx, float& gy) {
    const int y0 = max(y - 1, 0), y2 = min(y + 1, H - 1), x0 = max(x - 1, 0), x2 = min(x + 1, W - 1);
    const float* r0 = p + base + (long long)y0 * W;
    const float* r1 = p + base + (long long)y * W;
    const float* r2 = p + base + (long long)y2 * W;
    const float a00 = clamp_if(__ldg(r0 + x0), cl), a01 = clamp_if(__ldg(r0 + x), cl), a02 = clamp_if(__ldg(r0 + x2), cl);
    const float a10 = clamp_if(__ldg(r1 + x0), cl), a12 = clamp_if(__ldg(r1 + x2), cl);
    const float a20 = clamp_if(__ldg(r2 + x0), cl), a21 = clamp_if(__ldg(r2 + x), cl), a22 = clamp_if(__ldg(r2 + x2), cl);
    gx = ((a02 - a00) + 2.f * (a12 - a10) + (a22 - a20)) * 0.125f;
    gy = ((a20 - a00) + 2.f * (a21 - a01) + (a22 - a02)) * 0.125f;
}

__global__ void __launch_bounds__(256) k_loss_point(const float* __restrict__ xf, const float* __restrict__ ir, const float* __restrict__ vis,
                                                    float* __restrict__ tx, float* __restrict__ ty, float* __restrict__ partial,
                                                    int H, int W, long long N, float ktex, int cl) {
    __shared__ float red[8];
    const long long i = blockIdx.x * 256LL + threadIdx.x;
    float tsum = 0.f, isum = 0.f;
    if (i < N) {
        const long long HW = (long long)H * W;
        const long long b = i / HW;
        const int r = (int)(i - b * HW), y = r / W, x = r - y * W;
        float fx, fy, ax, ay, vx, vy;
        sobel3(xf, b * HW, y, x, H, W, cl, fx, fy);
        sobel3(ir, b * HW, y, x, H, W, 0, ax, ay);
        sobel3(vis, b * HW, y, x, H, W, 0, vx, vy);
        const float sf_ = sqrtf(fx * fx + fy * fy + SOBEL_EPS);
        const float sa = sqrtf(ax * ax + ay * ay + SOBEL_EPS), sv = sqrtf(vx * vx + vy * vy + SOBEL_EPS);
        const float dt = sf_ - fmaxf(sa, sv);
        tsum = fabsf(dt);
        if (tx) {
            const float k = ktex * sgn(dt) / sf_;
            tx[i] = k * fx;
            ty[i] = k * fy;
        }
        const float f = clamp_if(__ldg(xf + i), cl);
        isum = fabsf(f - fmaxf(__ldg(ir + i), __ldg(vis + i)));
    }
    const float t = block_sum<8>(tsum, red);
    const float s = block_sum<8>(isum, red);
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = t; partial[2 * blockIdx.x + 1] = s; }
}

// ---------------------------------------------------------------------------------------------------------
// K1: vertical Gaussian pass of the 8 moment planes (f, f^2, ir, ir^2, f ir, vis, vis^2, f vis) at 5 sigmas and of
//     |f-ir|, |f-vis| at sigma 8.  Tile = 32 columns x 8 rows per CTA (the 40 input rows are shared through L1).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_loss_vblur(const float* __restrict__ xf, const float* __restrict__ ir, const float* __restrict__ vis,
                                                    float* __restrict__ V, const GaussW gw, int H, int W, long long N, int cl) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const long long base = (long long)blockIdx.z * H * W;
    float acc[8][NS], al[2] = {0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 8; k++)
#pragma unroll
        for (int s = 0; s < NS; s++) acc[k][s] = 0.f;
#pragma unroll
    for (int t = -RAD; t <= RAD; t++) {
        const int yy = y + t;
        if (yy < 0 || yy >= H) continue;
        const long long o = base + (long long)yy * W + x;
        const float f = clamp_if(__ldg(xf + o), cl), a = __ldg(ir + o), c = __ldg(vis + o);
        const float pr[8] = {f, f * f, a, a * a, f * a, c, c * c, f * c};
#pragma unroll
        for (int s = 0; s < NS; s++) {
            if (t < -tap_radius(s) || t > tap_radius(s)) continue;
            const float w = gw.w[s][t + RAD];
#pragma unroll
            for (int k = 0; k < 8; k++) acc[k][s] = fmaf(w, pr[k], acc[k][s]);
        }
        al[0] = fmaf(gw.w[NS - 1][t + RAD], fabsf(f - a), al[0]);
        al[1] = fmaf(gw.w[NS - 1][t + RAD], fabsf(f - c), al[1]);
    }
    const long long o = base + (long long)y * W + x;
#pragma unroll
    for (int k = 0; k < 8; k++)
#pragma unroll
        for (int s = 0; s < NS; s++) V[(long long)(k * NS + s) * N + o] = acc[k][s];
    V[(long long)40 * N + o] = al[0];
    V[(long long)41 * N + o] = al[1];
}

// ---------------------------------------------------------------------------------------------------------
// K2: horizontal pass from shared memory, then the MS-SSIM + L1 value of the pixel and the 20 gradient planes.
//     One CTA = 128 consecutive pixels of one image row.
// ---------------------------------------------------------------------------------------------------------
static constexpr int HB = 128, HBP = HB + 2 * RAD;

template <int NP>
__device__ __forceinline__ void stage_rows(float* sm, const float* __restrict__ P, long long N, long long rowbase, int x0, int W) {
    for (int i = threadIdx.x; i < NP * HBP; i += HB) {
        const int pl = i / HBP, j = i - pl * HBP, xs = x0 - RAD + j;
        sm[i] = (xs >= 0 && xs < W) ? __ldg(P + (long long)pl * N + rowbase + xs) : 0.f;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(HB) k_loss_hblur_maps(const float* __restrict__ V, float* __restrict__ G, float* __restrict__ partial,
                                                        const GaussW gw, int H, int W, long long N, float w_ir, float gcoef) {
    __shared__ float sm[NV * HBP];
    __shared__ float red[4];
    const int x0 = blockIdx.x * HB, x = x0 + threadIdx.x;
    const long long rowbase = ((long long)blockIdx.z * H + blockIdx.y) * W;
    stage_rows<NV>(sm, V, N, rowbase, x0, W);
    float val = 0.f;
    if (x < W) {
        float m[8][NS], l1[2] = {0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 8; k++)
#pragma unroll
            for (int s = 0; s < NS; s++) {
                float a = 0.f;
                const float* row = sm + (k * NS + s) * HBP + threadIdx.x + RAD;
#pragma unroll
                for (int t = -tap_radius(s); t <= tap_radius(s); t++) a = fmaf(gw.w[s][t + RAD], row[t], a);
                m[k][s] = a;
            }
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const float* row = sm + (40 + j) * HBP + threadIdx.x + RAD;
#pragma unroll
            for (int t = -RAD; t <= RAD; t++) l1[j] = fmaf(gw.w[NS - 1][t + RAD], row[t], l1[j]);
        }
        // planes: 0 mu_f, 1 E[f^2], 2 mu_ir, 3 E[ir^2], 4 E[f ir], 5 mu_vis, 6 E[vis^2], 7 E[f vis]
        float ga[NS], gb[NS], gc[2][NS];
#pragma unroll
        for (int s = 0; s < NS; s++) { ga[s] = 0.f; gb[s] = 0.f; }
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const float wy = j == 0 ? w_ir : 1.f - w_ir;
            const int pm = j == 0 ? 2 : 5;
            float cs[NS], iden[NS];
#pragma unroll
            for (int s = 0; s < NS; s++) {
                const float a = m[0][s], my = m[pm][s];
                const float num = 2.f * (m[pm + 2][s] - a * my) + MS_C2;
                const float den = (m[1][s] - a * a) + (m[pm + 1][s] - my * my) + MS_C2;
                iden[s] = 1.f / den;
                cs[s] = num * iden[s];
            }
            const float a4 = m[0][NS - 1], m4 = m[pm][NS - 1];
            const float ilden = 1.f / (a4 * a4 + m4 * m4 + MS_C1);
            const float l = (2.f * a4 * m4 + MS_C1) * ilden;
            const float p3 = cs[0] * cs[1] * cs[2] * cs[3] * cs[4];
            const float lm = l * l * l, pics = p3 * p3 * p3;
            val += wy * MS_COMP * (MS_ALPHA * (1.f - lm * pics) + (1.f - MS_ALPHA) * l1[j]);
            // d(pixel value)/dT = -COMP ALPHA ;  T = l^3 (prod cs)^3
            const float up = -gcoef * wy * MS_COMP * MS_ALPHA;
            const float dT_dl = 3.f * l * l * pics;
            const float others[NS] = {cs[1] * cs[2] * cs[3] * cs[4], cs[0] * cs[2] * cs[3] * cs[4], cs[0] * cs[1] * cs[3] * cs[4],
                                      cs[0] * cs[1] * cs[2] * cs[4], cs[0] * cs[1] * cs[2] * cs[3]};
#pragma unroll
            for (int s = 0; s < NS; s++) {
                const float ucs = up * lm * 3.f * p3 * p3 * others[s];
                const float a = m[0][s], my = m[pm][s];
                ga[s] += ucs * (2.f * a * cs[s] - 2.f * my) * iden[s];
                gb[s] -= ucs * cs[s] * iden[s];
                gc[j][s] = ucs * 2.f * iden[s];
            }
            ga[NS - 1] += up * dT_dl * (2.f * m4 - 2.f * a4 * l) * ilden;
        }
        if (G) {
            const long long o = rowbase + x;
#pragma unroll
            for (int s = 0; s < NS; s++) {
                G[(long long)(s * 4 + 0) * N + o] = ga[s];
                G[(long long)(s * 4 + 1) * N + o] = gb[s];
                G[(long long)(s * 4 + 2) * N + o] = gc[0][s];
                G[(long long)(s * 4 + 3) * N + o] = gc[1][s];
            }
        }
    }
    const float t = block_sum<4>(val, red);
    if (threadIdx.x == 0) partial[((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = t;
}

// ---------------------------------------------------------------------------------------------------------
// K3: vertical pass of the 20 gradient planes (each with its own sigma)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_loss_vblur_grad(const float* __restrict__ G, float* __restrict__ VG, const GaussW gw, int H, int W, long long N) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const long long base = (long long)blockIdx.z * H * W;
    float acc[NG];
#pragma unroll
    for (int k = 0; k < NG; k++) acc[k] = 0.f;
#pragma unroll
    for (int t = -RAD; t <= RAD; t++) {
        const int yy = y + t;
        if (yy < 0 || yy >= H) continue;
        const long long o = base + (long long)yy * W + x;
#pragma unroll
        for (int s = 0; s < NS; s++) {
            if (t < -tap_radius(s) || t > tap_radius(s)) continue;
            const float w = gw.w[s][t + RAD];
#pragma unroll
            for (int k = 0; k < 4; k++) acc[s * 4 + k] = fmaf(w, __ldg(G + (long long)(s * 4 + k) * N + o), acc[s * 4 + k]);
        }
    }
    const long long o = base + (long long)y * W + x;
#pragma unroll
    for (int k = 0; k < NG; k++) VG[(long long)k * N + o] = acc[k];
}

// ---------------------------------------------------------------------------------------------------------
// K4: horizontal pass of the gradient planes and the whole gradient w.r.t. x:
//     sum_s [A_s + 2 f B_s + ir Cir_s + vis Cvis_s]  +  L1 branch (sign(f - y) * blur_8(const))
//     + adjoint of the replicate-border Sobel (gather form)  +  intensity sign  ; zero where the clamp saturates.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HB) k_loss_grad_final(const float* __restrict__ VG, const float* __restrict__ xf, const float* __restrict__ ir,
                                                        const float* __restrict__ vis, const float* __restrict__ tx, const float* __restrict__ ty,
                                                        float* __restrict__ gout, const GaussW gw, int H, int W, long long N, float w_ir,
                                                        float kl1, float kint, int cl) {
    __shared__ float sm[NG * HBP];
    const int x0 = blockIdx.x * HB, x = x0 + threadIdx.x, y = blockIdx.y;
    const long long imgbase = (long long)blockIdx.z * H * W, rowbase = imgbase + (long long)y * W;
    stage_rows<NG>(sm, VG, N, rowbase, x0, W);
    if (x >= W) return;
    const float raw = __ldg(xf + rowbase + x), f = clamp_if(raw, cl), a = __ldg(ir + rowbase + x), c = __ldg(vis + rowbase + x);
    float g = 0.f;
#pragma unroll
    for (int s = 0; s < NS; s++) {
        float b[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float* row = sm + (s * 4 + k) * HBP + threadIdx.x + RAD;
#pragma unroll
            for (int t = -tap_radius(s); t <= tap_radius(s); t++) b[k] = fmaf(gw.w[s][t + RAD], row[t], b[k]);
        }
        g += b[0] + 2.f * f * b[1] + a * b[2] + c * b[3];
    }
    // blur_8 of a constant plane under zero padding = ev(y) * eh(x)
    float ev = 0.f, eh = 0.f;
#pragma unroll
    for (int t = -RAD; t <= RAD; t++) {
        if (y + t >= 0 && y + t < H) ev += gw.w[NS - 1][t + RAD];
        if (x + t >= 0 && x + t < W) eh += gw.w[NS - 1][t + RAD];
    }
    g += kl1 * ev * eh * (w_ir * sgn(f - a) + (1.f - w_ir) * sgn(f - c));
    // Sobel adjoint: every output pixel q in the 3x3 neighbourhood whose (clamped) stencil reads pixel (y, x)
    for (int dy = -1; dy <= 1; dy++) {
        const int qy = y + dy;
        if (qy < 0 || qy >= H) continue;
        for (int dx = -1; dx <= 1; dx++) {
            const int qx = x + dx;
            if (qx < 0 || qx >= W) continue;
            float cx = 0.f, cy = 0.f;   // accumulated Sobel weights of q's taps that land on (y, x)
#pragma unroll
            for (int oy = -1; oy <= 1; oy++)
#pragma unroll
                for (int ox = -1; ox <= 1; ox++) {
                    if (min(max(qy + oy, 0), H - 1) != y || min(max(qx + ox, 0), W - 1) != x) continue;
                    cx += (float)(ox * (oy == 0 ? 2 : 1)) * 0.125f;
                    cy += (float)(oy * (ox == 0 ? 2 : 1)) * 0.125f;
                }
            const long long q = imgbase + (long long)qy * W + qx;
            g = fmaf(cx, __ldg(tx + q), g);
            g = fmaf(cy, __ldg(ty + q), g);
        }
    }
    g += kint * sgn(f - fmaxf(a, c));
    if (cl && (raw < 0.f || raw > 1.f)) g = 0.f;
    gout[rowbase + x] = g;
}

// K5: final sums (one block, double precision) -> loss[0..3] = total, ssim term, texture term, intensity term (scaled as a008:245-256)
__global__ void __launch_bounds__(1024) k_loss_reduce(const float* __restrict__ p_point, int n_point, const float* __restrict__ p_ssim, int n_ssim,
                                                      float* __restrict__ loss, float* __restrict__ total, double inv_n, float ssim_scale, float texture_scale,
                                                      float intensity_scale, float r_s, float r_t, float r_i) {
    __shared__ double red[3][32];
    double t = 0.0, i = 0.0, s = 0.0;
    for (int k = threadIdx.x; k < n_point; k += 1024) { t += p_point[2 * k]; i += p_point[2 * k + 1]; }
    for (int k = threadIdx.x; k < n_ssim; k += 1024) s += p_ssim[k];
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        t += __shfl_xor_sync(0xffffffffu, t, o);
        i += __shfl_xor_sync(0xffffffffu, i, o);
        s += __shfl_xor_sync(0xffffffffu, s, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = t; red[1][threadIdx.x >> 5] = i; red[2][threadIdx.x >> 5] = s; }
    __syncthreads();
    if (threadIdx.x == 0) {
        t = i = s = 0.0;
        for (int k = 0; k < 32; k++) { t += red[0][k]; i += red[1][k]; s += red[2][k]; }
        const float ls = (float)(s * inv_n) * ssim_scale, lt = (float)(t * inv_n) * texture_scale, li = (float)(i * inv_n) * intensity_scale;
        loss[0] = ls * r_s + lt * r_t + li * r_i;
        loss[1] = ls; loss[2] = lt; loss[3] = li;
        if (total) *total = loss[0];
    }
}

__global__ void k_scale_by_scalar(const float* __restrict__ in, const float* __restrict__ scalar, float* __restrict__ out, long long n) {
    const float s = __ldg(scalar);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = in[i] * s;
}

struct LossPlan {
    long long N;
    int n_point, n_ssim;
    size_t off_V, off_G, off_tx, off_ty, off_pp, off_ps, total;
};
static LossPlan loss_plan(const sf_fusion_loss_params* p) {
    LossPlan l;
    l.N = (long long)p->B * p->H * p->W;
    l.n_point = (int)((l.N + 255) / 256);
    l.n_ssim = ceil_div(p->W, HB) * p->H * p->B;
    size_t o = 0;
    l.off_V = o;  o += align_up((size_t)NV * l.N * sizeof(float));     // the vertical gradient pass reuses this region
    l.off_G = o;  o += align_up((size_t)NG * l.N * sizeof(float));
    l.off_tx = o; o += align_up((size_t)l.N * sizeof(float));
    l.off_ty = o; o += align_up((size_t)l.N * sizeof(float));
    l.off_pp = o; o += align_up((size_t)2 * l.n_point * sizeof(float));
    l.off_ps = o; o += align_up((size_t)l.n_ssim * sizeof(float));
    l.total = o;
    return l;
}

}  // namespace sf

extern "C" size_t sf_fusion_loss_workspace_bytes(const sf_fusion_loss_params* p) {
    if (!p || p->B <= 0 || p->H <= 0 || p->W <= 0) return 0;
    return sf::loss_plan(p).total;
}

extern "C" int sf_fusion_loss(const sf_fusion_loss_params* p, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace sf;
    SF_CHECK_ARG(p, "sf_fusion_loss: null params");
    SF_CHECK_ARG(p->fusion && p->ir && p->vis && p->loss, "sf_fusion_loss: null tensor pointer");
    SF_CHECK_ARG(p->B > 0 && p->H > 0 && p->W > 0, "sf_fusion_loss: empty image (%d,%d,%d)", p->B, p->H, p->W);
    SF_CHECK_ARG(p->H <= 65535 && p->B <= 65535, "sf_fusion_loss: H and B must be <= 65535 (grid limits), got %d, %d", p->H, p->B);
    const LossPlan l = loss_plan(p);
    SF_CHECK_ARG(workspace && workspace_bytes >= l.total, "sf_fusion_loss: workspace too small (%zu B given, %zu B needed)", workspace_bytes, l.total);
    cudaStream_t st = as_stream(stream);
    char* ws = static_cast<char*>(workspace);
    float* V = reinterpret_cast<float*>(ws + l.off_V);
    float* G = reinterpret_cast<float*>(ws + l.off_G);
    float* tx = reinterpret_cast<float*>(ws + l.off_tx);
    float* ty = reinterpret_cast<float*>(ws + l.off_ty);
    float* pp = reinterpret_cast<float*>(ws + l.off_pp);
    float* psm = reinterpret_cast<float*>(ws + l.off_ps);
    static const GaussW gw = make_gauss();
    const bool grad = p->g_fusion != nullptr;
    const double inv_n = 1.0 / (double)l.N;
    const int cl = p->clamp01 ? 1 : 0;
    const dim3 gv(ceil_div(p->W, 32), ceil_div(p->H, 8), p->B), gh(ceil_div(p->W, HB), p->H, p->B);
    {
        ProfScope ps("loss_sobel_intensity", 80.0 * l.N, (grad ? 20.0 : 12.0) * l.N, st);
        k_loss_point<<<l.n_point, 256, 0, st>>>(p->fusion, p->ir, p->vis, grad ? tx : nullptr, grad ? ty : nullptr, pp, p->H, p->W, l.N,
                                                (float)(p->texture_scale * p->r_texture * inv_n), cl);
        SF_CHECK_LAUNCH("loss_sobel_intensity");
    }
    {
        ProfScope ps("loss_vblur", 2.0 * 1066.0 * l.N, (12.0 + 4.0 * NV) * l.N, st);
        k_loss_vblur<<<gv, 256, 0, st>>>(p->fusion, p->ir, p->vis, V, gw, p->H, p->W, l.N, cl);
        SF_CHECK_LAUNCH("loss_vblur");
    }
    {
        ProfScope ps("loss_hblur_maps", 2.0 * 1066.0 * l.N, (4.0 * NV + (grad ? 4.0 * NG : 0.0)) * l.N, st);
        k_loss_hblur_maps<<<gh, HB, 0, st>>>(V, grad ? G : nullptr, psm, gw, p->H, p->W, l.N, p->w_ir, (float)(p->ssim_scale * p->r_ssim * inv_n));
        SF_CHECK_LAUNCH("loss_hblur_maps");
    }
    k_loss_reduce<<<1, 1024, 0, st>>>(pp, l.n_point, psm, l.n_ssim, p->loss, p->total, inv_n, p->ssim_scale, p->texture_scale, p->intensity_scale,
                                      p->r_ssim, p->r_texture, p->r_intensity);
    SF_CHECK_LAUNCH("loss_reduce");
    if (grad) {
        {
            ProfScope ps("loss_vblur_grad", 2.0 * 500.0 * l.N, 8.0 * NG * l.N, st);
            k_loss_vblur_grad<<<gv, 256, 0, st>>>(G, V, gw, p->H, p->W, l.N);
            SF_CHECK_LAUNCH("loss_vblur_grad");
        }
        ProfScope ps("loss_grad_final", 2.0 * 500.0 * l.N, (4.0 * NG + 24.0) * l.N, st);
        k_loss_grad_final<<<gh, HB, 0, st>>>(V, p->fusion, p->ir, p->vis, tx, ty, p->g_fusion, gw, p->H, p->W, l.N, p->w_ir,
                                             (float)(p->ssim_scale * p->r_ssim * inv_n * MS_COMP * (1.0 - MS_ALPHA)),
                                             (float)(p->intensity_scale * p->r_intensity * inv_n), cl);
        SF_CHECK_LAUNCH("loss_grad_final");
    }
    return SF_OK;
}

extern "C" int sf_scale_by_scalar(const float* in, const float* scalar, float* out, long long n, void* stream) {
    using namespace sf;
    SF_CHECK_ARG(in && scalar && out && n > 0, "sf_scale_by_scalar: null pointer or empty tensor");
    long long b = (n + 255) / 256;
    if (b > (long long)sm_count() * 8) b = (long long)sm_count() * 8;
    ProfScope ps("scale_by_scalar", 0.0, 8.0 * n, as_stream(stream));
    k_scale_by_scalar<<<(unsigned)b, 256, 0, as_stream(stream)>>>(in, scalar, out, n);
    SF_CHECK_LAUNCH("sf_scale_by_scalar");
    return SF_OK;
}

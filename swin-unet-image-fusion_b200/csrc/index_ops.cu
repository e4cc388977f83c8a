// Index / permutation kernels: pure data movement, bit-exact with the reference's
// einops / F.pad / torch.roll (SURVEY.md section 8 rows a2, a3, a13, a14/a15 index parts).
// All are HBM-bound: one pass, channel-innermost so a warp touches contiguous bytes,
// 128-bit accesses when C % 4 == 0, grid sized to a multiple of the SM count.
#include <cmath>
#include "common.cuh"

namespace sf {

static constexpr int kThreads = 256;

static inline int grid_for(long long work_items) {
    // persistent-ish grid: enough CTAs to fill 148 SMs x 8 resident CTAs, grid-stride inside
    long long blocks = (work_items + kThreads - 1) / kThreads;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

template <int V> struct VecT;
template <> struct VecT<1> { using type = float; };
template <> struct VecT<4> { using type = float4; };

__device__ __forceinline__ float vadd(float a, float b) { return a + b; }
__device__ __forceinline__ float4 vadd(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float vzero(float) { return 0.f; }
__device__ __forceinline__ float4 vzero(float4) { return make_float4(0.f, 0.f, 0.f, 0.f); }

// ---- reflect pad / crop -----------------------------------------------------------------
template <int V>
__global__ void k_pad_reflect(const typename VecT<V>::type* __restrict__ in, typename VecT<V>::type* __restrict__ out,
                              int B, int H, int W, int Cv, int Ho, int Wo) {
    long long total = (long long)B * Ho * Wo * Cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int ch = (int)(i % Cv);
        long long p = i / Cv;
        int c = (int)(p % Wo); p /= Wo;
        int r = (int)(p % Ho);
        int b = (int)(p / Ho);
        int sr = reflect_hi(r, H), sc = reflect_hi(c, W);
        out[i] = in[(((long long)b * H + sr) * W + sc) * Cv + ch];
    }
}

template <int V>
__global__ void k_pad_reflect_bwd(const typename VecT<V>::type* __restrict__ gout, typename VecT<V>::type* __restrict__ gin,
                                  int B, int H, int W, int Cv, int Ho, int Wo) {
    using T = typename VecT<V>::type;
    long long total = (long long)B * H * W * Cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int ch = (int)(i % Cv);
        long long p = i / Cv;
        int c = (int)(p % W); p /= W;
        int r = (int)(p % H);
        int b = (int)(p / H);
        // padded rows that read input row r: r itself and (if in range) 2(H-1)-r
        int r2 = 2 * (H - 1) - r, c2 = 2 * (W - 1) - c;
        bool hr = (r2 >= H && r2 < Ho), hc = (c2 >= W && c2 < Wo);
        const T* g = gout + ((long long)b * Ho) * Wo * Cv;
        T acc = g[((long long)r * Wo + c) * Cv + ch];
        if (hc) acc = vadd(acc, g[((long long)r * Wo + c2) * Cv + ch]);
        if (hr) {
            acc = vadd(acc, g[((long long)r2 * Wo + c) * Cv + ch]);
            if (hc) acc = vadd(acc, g[((long long)r2 * Wo + c2) * Cv + ch]);
        }
        gin[i] = acc;
    }
}

template <int V, bool ADD>
__global__ void k_crop(const typename VecT<V>::type* __restrict__ in, const typename VecT<V>::type* __restrict__ add,
                       typename VecT<V>::type* __restrict__ out, int B, int H, int W, int Cv, int Ho, int Wo) {
    long long total = (long long)B * Ho * Wo * Cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int ch = (int)(i % Cv);
        long long p = i / Cv;
        int c = (int)(p % Wo); p /= Wo;
        int r = (int)(p % Ho);
        int b = (int)(p / Ho);
        auto v = in[(((long long)b * H + r) * W + c) * Cv + ch];
        if (ADD) v = vadd(v, add[i]);
        out[i] = v;
    }
}

template <int V>
__global__ void k_crop_bwd(const typename VecT<V>::type* __restrict__ gout, typename VecT<V>::type* __restrict__ gin,
                           int B, int H, int W, int Cv, int Ho, int Wo) {
    using T = typename VecT<V>::type;
    long long total = (long long)B * H * W * Cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int ch = (int)(i % Cv);
        long long p = i / Cv;
        int c = (int)(p % W); p /= W;
        int r = (int)(p % H);
        int b = (int)(p / H);
        T v = vzero(T());
        if (r < Ho && c < Wo) v = gout[(((long long)b * Ho + r) * Wo + c) * Cv + ch];
        gin[i] = v;
    }
}

// ---- patch merge / unmerge ----------------------------------------------------------------
// MERGE=true : out (B,H/mh,W/mw,mh*mw*C) <- in (B,H,W,C)
// MERGE=false: out (B,H*mh,W*mw,C)       <- in (B,H,W,mh*mw*C)    (H,W = coarse size)
template <int V, bool MERGE>
__global__ void k_patch_rearrange(const typename VecT<V>::type* __restrict__ in, typename VecT<V>::type* __restrict__ out,
                                  int B, int Hc, int Wc, int Cv, int mh, int mw) {
    // iterate over the FINE map (B, Hc*mh, Wc*mw, Cv): both directions touch each fine element once
    int Hf = Hc * mh, Wf = Wc * mw;
    long long total = (long long)B * Hf * Wf * Cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long idx = i;
        if (MERGE) {
            // enumerate in OUTPUT (coarse) order so the writes are contiguous
            int cc = (int)(idx % ((long long)mh * mw * Cv));
            long long p = idx / ((long long)mh * mw * Cv);
            int X = (int)(p % Wc); p /= Wc;
            int Y = (int)(p % Hc);
            int b = (int)(p / Hc);
            int q = cc / Cv, ch = cc - q * Cv;
            int ph = q / mw, pw = q - ph * mw;
            out[i] = in[(((long long)b * Hf + (Y * mh + ph)) * Wf + (X * mw + pw)) * Cv + ch];
        } else {
            int ch = (int)(idx % Cv);
            long long p = idx / Cv;
            int c = (int)(p % Wf); p /= Wf;
            int r = (int)(p % Hf);
            int b = (int)(p / Hf);
            int Y = r / mh, ph = r - Y * mh, X = c / mw, pw = c - X * mw;
            out[i] = in[(((long long)b * Hc + Y) * Wc + X) * ((long long)mh * mw * Cv) + (long long)(ph * mw + pw) * Cv + ch];
        }
    }
}

// ---- window partition / reverse -------------------------------------------------------------
template <int V, bool PARTITION>
__global__ void k_window(const typename VecT<V>::type* __restrict__ in, typename VecT<V>::type* __restrict__ out,
                         WinGeom g, int Cv) {
    long long total = (long long)g.B * g.Hp * g.Wp * Cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int ch = (int)(i % Cv);
        long long p = i / Cv;  // = win * T + t
        int t = (int)(p % g.T);
        int win = (int)(p / g.T);
        long long src = win_token_src(g, win, t, nullptr);
        if (PARTITION) out[i] = in[src * Cv + ch];
        else out[src * Cv + ch] = in[i];
    }
}

__global__ void k_shift_mask(uint8_t* __restrict__ out, WinGeom g) {
    int nW = g.nWh * g.nWw;
    long long total = (long long)nW * g.T * g.T;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int j = (int)(i % g.T);
        long long p = i / g.T;
        int q = (int)(p % g.T);
        int w = (int)(p / g.T);
        int rq, rj;
        win_token_src(g, w, q, &rq);
        win_token_src(g, w, j, &rj);
        out[i] = rq != rj;
    }
}

__global__ void k_rel_bias(const float* __restrict__ table, float* __restrict__ out, int wsh, int wsw) {
    int T = wsh * wsw;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < T * T; i += gridDim.x * blockDim.x) {
        int q = i / T, k = i - q * T;
        int dr = k / wsw - q / wsw + (wsh - 1), dc = k % wsw - q % wsw + (wsw - 1);
        out[i] = table[dr * (2 * wsw - 1) + dc];
    }
}

// ---- NCHW <-> NHWC ---------------------------------------------------------------------------
// per image: transpose a (R x S) row-major matrix into (S x R)
__global__ void k_transpose(const float* __restrict__ in, float* __restrict__ out, int R, int S) {
    __shared__ float tile[32][33];
    const float* src = in + (long long)blockIdx.z * R * S;
    float* dst = out + (long long)blockIdx.z * R * S;
    int s0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        int r = r0 + dy, s = s0 + threadIdx.x;
        if (r < R && s < S) tile[dy][threadIdx.x] = src[(long long)r * S + s];
    }
    __syncthreads();
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        int s = s0 + dy, r = r0 + threadIdx.x;
        if (r < R && s < S) dst[(long long)s * R + r] = tile[threadIdx.x][dy];
    }
}

__global__ void k_add(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = a[i] + b[i];
}

__global__ void k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                       float step_size, float b1, float b2, float omb1, float omb2, float inv_sqrt_bc2, float eps, float gscale) {
    // omb1 / omb2 = 1 - beta, formed in double on the host as torch.optim.Adam does (1 - 0.999f in float is off by 5e-5 relative)
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float gi = g[i] * gscale;
        float mi = b1 * m[i] + omb1 * gi;
        float vi = b2 * v[i] + omb2 * gi * gi;
        m[i] = mi;
        v[i] = vi;
        p[i] -= step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
    }
}

static inline bool vec4_ok(int C, const void* a, const void* b, const void* c = nullptr) {
    auto al = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    return (C % 4 == 0) && al(a) && al(b) && al(c);
}

}  // namespace sf

using namespace sf;

extern "C" {

int sf_pad_reflect(const float* in, float* out, int B, int H, int W, int C, int pd, int pr, void* stream) {
    SF_CHECK_ARG(in && out && B > 0 && H > 0 && W > 0 && C > 0, "sf_pad_reflect: bad shape");
    SF_CHECK_ARG(pd >= 0 && pr >= 0 && pd < H && pr < W, "sf_pad_reflect: reflect pad (%d,%d) must be smaller than the map (%d,%d)", pd, pr, H, W);
    int Ho = H + pd, Wo = W + pr;
    ProfScope ps("pad_reflect", 0.0, 4.0 * B * C * ((double)H * W + (double)Ho * Wo), as_stream(stream));
    if (vec4_ok(C, in, out)) {
        long long n = (long long)B * Ho * Wo * (C / 4);
        k_pad_reflect<4><<<grid_for(n), kThreads, 0, as_stream(stream)>>>((const float4*)in, (float4*)out, B, H, W, C / 4, Ho, Wo);
    } else {
        long long n = (long long)B * Ho * Wo * C;
        k_pad_reflect<1><<<grid_for(n), kThreads, 0, as_stream(stream)>>>(in, out, B, H, W, C, Ho, Wo);
    }
    SF_CHECK_LAUNCH("sf_pad_reflect");
    return SF_OK;
}

int sf_pad_reflect_bwd(const float* gout, float* gin, int B, int H, int W, int C, int pd, int pr, void* stream) {
    SF_CHECK_ARG(gout && gin && B > 0 && H > 0 && W > 0 && C > 0 && pd >= 0 && pr >= 0 && pd < H && pr < W, "sf_pad_reflect_bwd: bad shape");
    int Ho = H + pd, Wo = W + pr;
    ProfScope ps("pad_reflect_bwd", 0.0, 4.0 * B * C * ((double)H * W + (double)Ho * Wo), as_stream(stream));
    if (vec4_ok(C, gout, gin)) {
        long long n = (long long)B * H * W * (C / 4);
        k_pad_reflect_bwd<4><<<grid_for(n), kThreads, 0, as_stream(stream)>>>((const float4*)gout, (float4*)gin, B, H, W, C / 4, Ho, Wo);
    } else {
        long long n = (long long)B * H * W * C;
        k_pad_reflect_bwd<1><<<grid_for(n), kThreads, 0, as_stream(stream)>>>(gout, gin, B, H, W, C, Ho, Wo);
    }
    SF_CHECK_LAUNCH("sf_pad_reflect_bwd");
    return SF_OK;
}

int sf_crop(const float* in, const float* add, float* out, int B, int H, int W, int C, int cd, int cr, void* stream) {
    SF_CHECK_ARG(in && out && B > 0 && C > 0 && cd >= 0 && cr >= 0 && cd < H && cr < W, "sf_crop: bad shape");
    int Ho = H - cd, Wo = W - cr;
    bool v4 = vec4_ok(C, in, out, add);
    int Cv = v4 ? C / 4 : C;
    long long n = (long long)B * Ho * Wo * Cv;
    cudaStream_t st = as_stream(stream);
    ProfScope ps("crop", 0.0, 4.0 * B * C * (double)Ho * Wo * (add ? 3.0 : 2.0), st);
    if (v4) {
        if (add) k_crop<4, true><<<grid_for(n), kThreads, 0, st>>>((const float4*)in, (const float4*)add, (float4*)out, B, H, W, Cv, Ho, Wo);
        else k_crop<4, false><<<grid_for(n), kThreads, 0, st>>>((const float4*)in, nullptr, (float4*)out, B, H, W, Cv, Ho, Wo);
    } else {
        if (add) k_crop<1, true><<<grid_for(n), kThreads, 0, st>>>(in, add, out, B, H, W, Cv, Ho, Wo);
        else k_crop<1, false><<<grid_for(n), kThreads, 0, st>>>(in, nullptr, out, B, H, W, Cv, Ho, Wo);
    }
    SF_CHECK_LAUNCH("sf_crop");
    return SF_OK;
}

int sf_crop_bwd(const float* gout, float* gin, int B, int H, int W, int C, int cd, int cr, void* stream) {
    SF_CHECK_ARG(gout && gin && B > 0 && C > 0 && cd >= 0 && cr >= 0 && cd < H && cr < W, "sf_crop_bwd: bad shape");
    int Ho = H - cd, Wo = W - cr;
    ProfScope ps("crop_bwd", 0.0, 4.0 * B * C * ((double)H * W + (double)Ho * Wo), as_stream(stream));
    if (vec4_ok(C, gout, gin)) {
        long long n = (long long)B * H * W * (C / 4);
        k_crop_bwd<4><<<grid_for(n), kThreads, 0, as_stream(stream)>>>((const float4*)gout, (float4*)gin, B, H, W, C / 4, Ho, Wo);
    } else {
        long long n = (long long)B * H * W * C;
        k_crop_bwd<1><<<grid_for(n), kThreads, 0, as_stream(stream)>>>(gout, gin, B, H, W, C, Ho, Wo);
    }
    SF_CHECK_LAUNCH("sf_crop_bwd");
    return SF_OK;
}

int sf_patch_merge(const float* in, float* out, int B, int H, int W, int C, int mh, int mw, void* stream) {
    SF_CHECK_ARG(in && out && B > 0 && C > 0 && mh > 0 && mw > 0 && H % mh == 0 && W % mw == 0,
                 "sf_patch_merge: (%d,%d) not divisible by merging size (%d,%d)", H, W, mh, mw);
    ProfScope ps("patch_merge", 0.0, 8.0 * B * C * (double)H * W, as_stream(stream));
    if (vec4_ok(C, in, out)) {
        long long n = (long long)B * H * W * (C / 4);
        k_patch_rearrange<4, true><<<grid_for(n), kThreads, 0, as_stream(stream)>>>((const float4*)in, (float4*)out, B, H / mh, W / mw, C / 4, mh, mw);
    } else {
        long long n = (long long)B * H * W * C;
        k_patch_rearrange<1, true><<<grid_for(n), kThreads, 0, as_stream(stream)>>>(in, out, B, H / mh, W / mw, C, mh, mw);
    }
    SF_CHECK_LAUNCH("sf_patch_merge");
    return SF_OK;
}

int sf_patch_unmerge(const float* in, float* out, int B, int H, int W, int C, int mh, int mw, void* stream) {
    SF_CHECK_ARG(in && out && B > 0 && C > 0 && mh > 0 && mw > 0 && H > 0 && W > 0, "sf_patch_unmerge: bad shape");
    ProfScope ps("patch_unmerge", 0.0, 8.0 * B * C * (double)H * W * mh * mw, as_stream(stream));
    if (vec4_ok(C, in, out)) {
        long long n = (long long)B * H * mh * W * mw * (C / 4);
        k_patch_rearrange<4, false><<<grid_for(n), kThreads, 0, as_stream(stream)>>>((const float4*)in, (float4*)out, B, H, W, C / 4, mh, mw);
    } else {
        long long n = (long long)B * H * mh * W * mw * C;
        k_patch_rearrange<1, false><<<grid_for(n), kThreads, 0, as_stream(stream)>>>(in, out, B, H, W, C, mh, mw);
    }
    SF_CHECK_LAUNCH("sf_patch_unmerge");
    return SF_OK;
}

static int window_common(const char* name, int B, int Hp, int Wp, int C, int wsh, int wsw) {
    SF_CHECK_ARG(B > 0 && C > 0 && wsh > 0 && wsw > 0 && Hp > 0 && Wp > 0, "%s: bad shape", name);
    SF_CHECK_ARG(Hp % wsh == 0 && Wp % wsw == 0, "%s: map (%d,%d) is not a multiple of the window (%d,%d)", name, Hp, Wp, wsh, wsw);
    return SF_OK;
}

int sf_window_partition(const float* in, float* out, int B, int Hp, int Wp, int C, int wsh, int wsw, int shift, void* stream) {
    SF_TRY(window_common("sf_window_partition", B, Hp, Wp, C, wsh, wsw));
    SF_CHECK_ARG(in && out, "sf_window_partition: null pointer");
    WinGeom g = make_geom(B, Hp, Wp, wsh, wsw, shift);
    ProfScope ps("window_partition", 0.0, 8.0 * B * C * (double)Hp * Wp, as_stream(stream));
    if (vec4_ok(C, in, out)) {
        long long n = (long long)B * Hp * Wp * (C / 4);
        k_window<4, true><<<grid_for(n), kThreads, 0, as_stream(stream)>>>((const float4*)in, (float4*)out, g, C / 4);
    } else {
        long long n = (long long)B * Hp * Wp * C;
        k_window<1, true><<<grid_for(n), kThreads, 0, as_stream(stream)>>>(in, out, g, C);
    }
    SF_CHECK_LAUNCH("sf_window_partition");
    return SF_OK;
}

int sf_window_reverse(const float* in, float* out, int B, int Hp, int Wp, int C, int wsh, int wsw, int shift, void* stream) {
    SF_TRY(window_common("sf_window_reverse", B, Hp, Wp, C, wsh, wsw));
    SF_CHECK_ARG(in && out, "sf_window_reverse: null pointer");
    WinGeom g = make_geom(B, Hp, Wp, wsh, wsw, shift);
    ProfScope ps("window_reverse", 0.0, 8.0 * B * C * (double)Hp * Wp, as_stream(stream));
    if (vec4_ok(C, in, out)) {
        long long n = (long long)B * Hp * Wp * (C / 4);
        k_window<4, false><<<grid_for(n), kThreads, 0, as_stream(stream)>>>((const float4*)in, (float4*)out, g, C / 4);
    } else {
        long long n = (long long)B * Hp * Wp * C;
        k_window<1, false><<<grid_for(n), kThreads, 0, as_stream(stream)>>>(in, out, g, C);
    }
    SF_CHECK_LAUNCH("sf_window_reverse");
    return SF_OK;
}

int sf_shift_mask(uint8_t* out, int Hp, int Wp, int wsh, int wsw, void* stream) {
    SF_TRY(window_common("sf_shift_mask", 1, Hp, Wp, 1, wsh, wsw));
    SF_CHECK_ARG(out, "sf_shift_mask: null pointer");
    WinGeom g = make_geom(1, Hp, Wp, wsh, wsw, 1);
    long long n = (long long)g.nWh * g.nWw * g.T * g.T;
    k_shift_mask<<<grid_for(n), kThreads, 0, as_stream(stream)>>>(out, g);
    SF_CHECK_LAUNCH("sf_shift_mask");
    return SF_OK;
}

int sf_relative_position_bias(const float* table, float* out, int wsh, int wsw, void* stream) {
    SF_CHECK_ARG(table && out && wsh > 0 && wsw > 0, "sf_relative_position_bias: bad args");
    k_rel_bias<<<grid_for((long long)wsh * wsw * wsh * wsw), kThreads, 0, as_stream(stream)>>>(table, out, wsh, wsw);
    SF_CHECK_LAUNCH("sf_relative_position_bias");
    return SF_OK;
}

static int transpose_launch(const char* name, const float* in, float* out, int B, int R, int S, void* stream) {
    SF_CHECK_ARG(in && out && B > 0 && R > 0 && S > 0 && B <= 65535, "%s: bad shape", name);
    dim3 grid(ceil_div(S, 32), ceil_div(R, 32), B), block(32, 8);
    ProfScope ps("layout_transpose", 0.0, 8.0 * B * (double)R * S, as_stream(stream));
    SF_CHECK_ARG(grid.y <= 65535, "%s: map too large", name);
    k_transpose<<<grid, block, 0, as_stream(stream)>>>(in, out, R, S);
    SF_CHECK_LAUNCH(name);
    return SF_OK;
}

int sf_nchw_to_nhwc(const float* in, float* out, int B, int C, int H, int W, void* stream) {
    return transpose_launch("sf_nchw_to_nhwc", in, out, B, C, H * W, stream);
}
int sf_nhwc_to_nchw(const float* in, float* out, int B, int C, int H, int W, void* stream) {
    return transpose_launch("sf_nhwc_to_nchw", in, out, B, H * W, C, stream);
}

int sf_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, double lr, double beta1, double beta2,
                 double eps, int step, double grad_scale, void* stream) {
    SF_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1, "sf_adam_step: bad args");
    // scalars arrive and are combined in double, as torch.optim.Adam forms them in Python, and are rounded to fp32 once
    const double bc1 = 1.0 - pow(beta1, step), bc2 = 1.0 - pow(beta2, step);
    ProfScope ps("adam_step", 12.0 * (double)n, 28.0 * (double)n, as_stream(stream));
    k_adam<<<grid_for(n), kThreads, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, (float)(lr / bc1), (float)beta1, (float)beta2,
                                                             (float)(1.0 - beta1), (float)(1.0 - beta2), (float)(1.0 / sqrt(bc2)), (float)eps,
                                                             (float)grad_scale);
    SF_CHECK_LAUNCH("sf_adam_step");
    return SF_OK;
}

int sf_add(const float* a, const float* b, float* out, long long n, void* stream) {
    SF_CHECK_ARG(a && b && out && n > 0, "sf_add: bad args");
    ProfScope ps("add", 0.0, 12.0 * (double)n, as_stream(stream));
    k_add<<<grid_for(n), kThreads, 0, as_stream(stream)>>>(a, b, out, n);
    SF_CHECK_LAUNCH("sf_add");
    return SF_OK;
}

}  // extern "C"

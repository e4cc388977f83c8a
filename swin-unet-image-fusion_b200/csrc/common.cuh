// Shared host/device helpers for libswinfuse (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstdint>
#include "../../include/swinfuse.h"

namespace sf {

// ---- error / bookkeeping (thread local; never throws across the ABI) --------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
// interned printf-style kernel label for ProfScope (stable pointer for the life of the process)
const char* prof_name(const char* fmt, int v);

// Optional per-kernel timing (sf_profile_enable): a ProfScope brackets the launches issued while
// it is alive with a CUDA-event pair on the launching stream and books the algorithmic FLOPs /
// bytes the launcher states, so bench.py can report roofline numbers measured live.
struct ProfScope {
    int rec;
    cudaStream_t st;
    ProfScope(const char* name, double flops, double bytes, cudaStream_t stream);
    ~ProfScope();
};

#define SF_CHECK_ARG(cond, ...)                      \
    do {                                             \
        if (!(cond)) {                               \
            ::sf::set_error(__VA_ARGS__);            \
            return SF_ERR_INVALID;                   \
        }                                            \
    } while (0)

// checks the launch itself (cudaPeekAtLastError does not clear sticky errors and is legal
// during stream capture)
#define SF_CHECK_LAUNCH(name)                                                              \
    do {                                                                                   \
        cudaError_t e__ = cudaPeekAtLastError();                                           \
        ::sf::count_launch();                                                              \
        if (e__ != cudaSuccess) {                                                          \
            ::sf::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));       \
            return SF_ERR_CUDA;                                                            \
        }                                                                                  \
    } while (0)

#define SF_TRY(expr)                 \
    do {                             \
        int rc__ = (expr);           \
        if (rc__ != SF_OK) return rc__; \
    } while (0)

// SMs of the current device (148 on B200), cached per device: persistent grids are sized in multiples of it
int sm_count();
// One-time per-DEVICE set-up (cudaFuncSetAttribute is per context, not per host thread): a static instance per call
// site remembers which devices have been configured.  A race between two host threads only repeats the cheap call.
struct DeviceOnce {
    unsigned long long mask = 0;
    int dev = 0;
    bool need() { cudaGetDevice(&dev); return !((mask >> (dev & 63)) & 1ull); }
    void done() { mask |= 1ull << (dev & 63); }
};

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }
static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// bump allocator over the caller's workspace
struct Workspace {
    char* base;
    size_t size, off;
    Workspace(void* p, size_t n) : base((char*)p), size(n), off(0) {}
    template <typename T>
    T* take(size_t count) {
        size_t bytes = align_up(count * sizeof(T));
        if (off + bytes > size) return nullptr;
        T* r = reinterpret_cast<T*>(base + off);
        off += bytes;
        return r;
    }
};

// ---- index math shared by every kernel ---------------------------
// F.pad(mode="reflect") on the high side, a006:128-131: out[L+k] = in[L-2-k]
__host__ __device__ __forceinline__ int reflect_hi(int i, int n) { return i < n ? i : 2 * (n - 1) - i; }
// full reflect (both sides) for the 3x3 head convs, a013:126-148
__host__ __device__ __forceinline__ int reflect_both(int i, int n) {
    if (i < 0) i = -i;
    return i < n ? i : 2 * (n - 1) - i;
}
// torch.roll(x, -s): shifted[r] = x[(r+s) % n], a001:432-445
__host__ __device__ __forceinline__ int shift_src(int r, int n, int s) {
    int v = r + s;
    return v >= n ? v - n : v;
}
// region id along one axis of the shifted frame, a001:225-234
__host__ __device__ __forceinline__ int shift_region(int r, int n, int w, int s) {
    return r < n - w ? 0 : (r < n - s ? 1 : 2);
}

struct WinGeom {
    int B, Hp, Wp, wsh, wsw, nWh, nWw, T, sh, sw;  // sh/sw = 0 when not shifted
    int shift;
};
static inline WinGeom make_geom(int B, int Hp, int Wp, int wsh, int wsw, int shift) {
    WinGeom g;
    g.B = B; g.Hp = Hp; g.Wp = Wp; g.wsh = wsh; g.wsw = wsw;
    g.nWh = Hp / wsh; g.nWw = Wp / wsw; g.T = wsh * wsw;
    g.shift = shift ? 1 : 0;
    g.sh = shift ? wsh / 2 : 0; g.sw = shift ? wsw / 2 : 0;
    return g;
}
// window `win` (0..B*nWh*nWw), token t -> flat token index (b*Hp + r)*Wp + c in the UN-shifted map,
// and the region id of the token in the shifted frame (a001:165-172, 222-247, 442-445)
__device__ __forceinline__ long long win_token_src(const WinGeom& g, int win, int t, int* region) {
    int nW = g.nWh * g.nWw;
    int b = win / nW, w = win - b * nW;
    int wh = w / g.nWw, ww = w - wh * g.nWw;
    int ti = t / g.wsw, tj = t - ti * g.wsw;
    int r = wh * g.wsh + ti, c = ww * g.wsw + tj;
    if (region) *region = g.shift ? 3 * shift_region(r, g.Hp, g.wsh, g.sh) + shift_region(c, g.Wp, g.wsw, g.sw) : 0;
    int sr = shift_src(r, g.Hp, g.sh), sc = shift_src(c, g.Wp, g.sw);
    return ((long long)b * g.Hp + sr) * g.Wp + sc;
}

// ---- division by a run-time constant (n < 2^31): q = umulhi(n, mul) >> shr ------------------
struct FastDiv {
    uint32_t mul, shr, d;
};
static inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.d = d;
    if (d <= 1) { f.mul = 0; f.shr = 0; return f; }
    uint32_t lg = 0;
    while ((1ull << lg) < d) lg++;
    const uint32_t p = 31 + lg;
    f.mul = (uint32_t)(((1ull << p) + d - 1) / d);
    f.shr = p - 32;
    return f;
}
__host__ __device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
#ifdef __CUDA_ARCH__
    return f.d <= 1 ? n : (__umulhi(n, f.mul) >> f.shr);
#else
    return f.d <= 1 ? n : (uint32_t)(((uint64_t)n * f.mul) >> 32) >> f.shr;
#endif
}

// "window order" of the tokens of a (B,Hp,Wp) map: row m = window * T + t with the windows and
// tokens numbered as in win_token_src.  The window-attention GEMMs of the bf16 path run on rows in
// this order (gathered by the A producers, scattered back by the projection epilogue), so that a
// window's q/k/v are contiguous for the attention core.
struct WinOrder {
    WinGeom g;
    FastDiv dT, dnW, dnWw, dwsw;
};
static inline WinOrder make_winorder(const WinGeom& g) {
    WinOrder o;
    o.g = g;
    o.dT = make_fastdiv((uint32_t)g.T);
    o.dnW = make_fastdiv((uint32_t)(g.nWh * g.nWw));
    o.dnWw = make_fastdiv((uint32_t)g.nWw);
    o.dwsw = make_fastdiv((uint32_t)g.wsw);
    return o;
}
// row m (window order) -> flat token index in the un-shifted map; *win / *tok receive (window, token in window)
__device__ __forceinline__ long long win_order_token(const WinOrder& o, uint32_t m, uint32_t* win = nullptr, uint32_t* tok = nullptr) {
    const WinGeom& g = o.g;
    const uint32_t w = fdiv(m, o.dT), t = m - w * (uint32_t)g.T;
    if (win) *win = w;
    if (tok) *tok = t;
    const uint32_t b = fdiv(w, o.dnW), wi = w - b * (uint32_t)(g.nWh * g.nWw);
    const uint32_t wh = fdiv(wi, o.dnWw), ww = wi - wh * (uint32_t)g.nWw;
    const uint32_t ti = fdiv(t, o.dwsw), tj = t - ti * (uint32_t)g.wsw;
    const int sr = shift_src((int)(wh * g.wsh + ti), g.Hp, g.sh), sc = shift_src((int)(ww * g.wsw + tj), g.Wp, g.sw);
    return ((long long)b * g.Hp + sr) * g.Wp + sc;
}

__device__ __forceinline__ float elu1(float v) { return v > 0.f ? v : expm1f(v); }

}  // namespace sf

// Warp-level HMMA (mma.sync m16n8k16) and softmax helpers shared by the attention-core kernels
// (attn_frag.cu: stand-alone core; wa_fused.cu: the fused window-attention operator).
#pragma once
#include <cuda_fp16.h>
#include <cstdint>
#include "tc_common.cuh"

namespace sf {

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcpf(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float max3f(float a, float b, float c) {
    float y;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
    return y;
}
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// D = A B + C with C and D in different registers: the bias fragment (or zero) is the C operand itself, no copy into the
// accumulator first
__device__ __forceinline__ void mma16816_cd(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1,
                                            float c0, float c1, float c2, float c3) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(c0), "f"(c1), "f"(c2), "f"(c3));
}
// bf16 operands (same fragment layout): P V of the no-max softmax fast path, where P = 2^s spans the fp32 exponent range
__device__ __forceinline__ void mma16816_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma16816_bf16_z(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f));
}
// 8x8 b16 transpose across the warp: in: lane (gq, tq) holds M[gq][2tq..2tq+1]; out: M[2tq..2tq+1][gq]
__device__ __forceinline__ uint32_t movm_trans(uint32_t x) {
    uint32_t y;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint2 ldg64(const __half* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }
__device__ __forceinline__ uint32_t ldg32(const __half* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }

// Bias fragment of one 16-row slab of a 7x7 window (x log2 e; padded keys 49..55 = -1e30) and the key patterns of the
// shift mask, in m16n8k16 accumulator coordinates: lane (gq, tq) holds rows r0 = slab*16 + gq and r1 = r0 + 8,
// keys nt*8 + 2tq + e.  bias[i,j] = table[r_j - r_i + 6][c_j - c_i + 6]: key minus query (a001:113-144).
// mh / mw: the masked-slot pattern of the lane's two rows in a window of the LAST window row / column of the shifted
// frame (a001:222-272: tokens 0..3 and 4..6 of such a window lie in different regions along that axis).
struct SlabMask { uint32_t mh0, mw0, mh1, mw1; };
__device__ __forceinline__ SlabMask slab_bias_and_mask(float (&bias)[7][4], const float* __restrict__ table, int r0, int r1, int tq) {
    constexpr int FT7 = 49;
    constexpr float LOG2E = 1.4426950408889634f;
    constexpr float NEG = -1e30f;
    uint32_t kh = 0, kw = 0;   // bit 2nt+e: key nt*8 + 2tq + e lies in the upper part (>= 4) of the window along H / W
#pragma unroll
    for (int nt = 0; nt < 7; nt++) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int key = nt * 8 + 2 * tq + e;
            float b0 = NEG, b1 = NEG;
            if (key < FT7) {
                const int kr = key / 7, kc = key - kr * 7;
                b0 = r0 < FT7 ? LOG2E * __ldg(table + (kr - r0 / 7 + 6) * 13 + (kc - r0 % 7 + 6)) : 0.f;
                b1 = r1 < FT7 ? LOG2E * __ldg(table + (kr - r1 / 7 + 6) * 13 + (kc - r1 % 7 + 6)) : 0.f;
                if (kr >= 4) kh |= 1u << (2 * nt + e);
                if (kc >= 4) kw |= 1u << (2 * nt + e);
            }
            bias[nt][e] = b0;
            bias[nt][2 + e] = b1;
        }
    }
    SlabMask m;
    m.mh0 = (r0 < FT7 && r0 / 7 >= 4) ? ~kh : kh; m.mw0 = (r0 < FT7 && r0 % 7 >= 4) ? ~kw : kw;
    m.mh1 = (r1 < FT7 && r1 / 7 >= 4) ? ~kh : kh; m.mw1 = (r1 < FT7 && r1 % 7 >= 4) ? ~kw : kw;
    return m;
}

}  // namespace sf

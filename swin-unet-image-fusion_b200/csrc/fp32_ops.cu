// fp32 (FFMA) forward kernels: the <=1e-4 precision mode of every fused operator.
// These are the correctness anchor of the library; the bf16 tcgen05 path (bf16_ops.cu)
// shares the index math in common.cuh and is checked against the same oracle.
#include "fp32_kernels.cuh"

namespace sf {

// =============================================================================================
// LayerNorm over C per token (my_layer_norm, a004:54-72), optional ELU, optional un-merge scatter
// =============================================================================================
// One warp per row.  Rows are short (C <= 1536) so the three passes hit L1.
template <bool ACT, bool UNMERGE>
__global__ void k_ln_rows(const float* __restrict__ in, const float* __restrict__ gamma, const float* __restrict__ beta,
                          float* __restrict__ out, long long M, int C, float eps, UnmergeGeom ug) {
    int lane = threadIdx.x & 31;
    long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    long long nrows_per_pass = ((long long)gridDim.x * blockDim.x) >> 5;
    for (; row < M; row += nrows_per_pass) {
        const float* x = in + row * C;
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s += x[c];
        s = warp_sum(s);
        float mean = s / (float)C;
        float v = 0.f;
        for (int c = lane; c < C; c += 32) { float d = x[c] - mean; v += d * d; }
        v = warp_sum(v);
        float rstd = rsqrtf(v / (float)C + eps);
        if (!UNMERGE) {
            float* o = out + row * C;
            for (int c = lane; c < C; c += 32) {
                float y = (x[c] - mean) * rstd * gamma[c] + beta[c];
                o[c] = ACT ? elu1(y) : y;
            }
        } else {
            // row = (b, Y, X) of the coarse map; channel n = q*Cout + ch -> fine pixel (Y*mh+ph, X*mw+pw)
            int X = (int)(row % ug.Wc);
            long long p = row / ug.Wc;
            int Y = (int)(p % ug.Hc);
            long long b = p / ug.Hc;
            int Hf = ug.Hc * ug.mh, Wf = ug.Wc * ug.mw;
            for (int c = lane; c < C; c += 32) {
                float y = (x[c] - mean) * rstd * gamma[c] + beta[c];
                int q = c / ug.Cout, ch = c - q * ug.Cout;
                int ph = q / ug.mw, pw = q - ph * ug.mw;
                out[((b * Hf + (Y * ug.mh + ph)) * Wf + (X * ug.mw + pw)) * ug.Cout + ch] = ACT ? elu1(y) : y;
            }
        }
    }
}


// Short rows (C <= 64, C % 4 == 0): one thread per row, the row in registers (float4 loads, no
// shuffles) -- the warp-per-row kernel above leaves most lanes idle there (C = 4 or 24).
template <bool ACT, bool UNMERGE, int NF4>
__global__ void k_ln_rows_small(const float* __restrict__ in, const float* __restrict__ gamma, const float* __restrict__ beta,
                                float* __restrict__ out, long long M, int C, float eps, UnmergeGeom ug) {
    const int nf4 = C >> 2;
    for (long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x; row < M; row += (long long)gridDim.x * blockDim.x) {
        float4 v[NF4];
        const float4* src = reinterpret_cast<const float4*>(in + row * C);
#pragma unroll
        for (int i = 0; i < NF4; i++) v[i] = i < nf4 ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NF4; i++) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        const float mean = s / (float)C;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NF4; i++) {
            if (i < nf4) {
                float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
                q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
            }
        }
        const float rstd = rsqrtf(q / (float)C + eps);
        long long obase = row * C;
        int Cout = C, ostride = 0, Wf = 0;
        if (UNMERGE) {
            int X = (int)(row % ug.Wc);
            long long p = row / ug.Wc;
            int Y = (int)(p % ug.Hc);
            long long b = p / ug.Hc;
            Wf = ug.Wc * ug.mw;
            obase = ((b * (ug.Hc * ug.mh) + (long long)Y * ug.mh) * Wf + (long long)X * ug.mw) * ug.Cout;
            Cout = ug.Cout;
            ostride = 1;
        }
#pragma unroll
        for (int i = 0; i < NF4; i++) {
            if (i < nf4) {
                float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i), bb = __ldg(reinterpret_cast<const float4*>(beta) + i);
                float y[4] = {(v[i].x - mean) * rstd * g.x + bb.x, (v[i].y - mean) * rstd * g.y + bb.y,
                              (v[i].z - mean) * rstd * g.z + bb.z, (v[i].w - mean) * rstd * g.w + bb.w};
#pragma unroll
                for (int e = 0; e < 4; e++) if (ACT) y[e] = elu1(y[e]);
                if (!ostride) {
                    *reinterpret_cast<float4*>(out + obase + i * 4) = make_float4(y[0], y[1], y[2], y[3]);
                } else {
                    // channel c = q*Cout + ch -> fine pixel (ph, pw) = (q / mw, q % mw), channel ch
#pragma unroll
                    for (int e = 0; e < 4; e++) {
                        int c = i * 4 + e, qq = c / Cout, ch = c - qq * Cout;
                        int ph = qq / ug.mw, pw = qq - ph * ug.mw;
                        out[obase + ((long long)ph * Wf + pw) * Cout + ch] = y[e];
                    }
                }
            }
        }
    }
}

template <int NF4>
static void launch_ln_small(const float* in, const float* gamma, const float* beta, float* out, long long M, int C, float eps,
                            int act, const UnmergeGeom* ug, cudaStream_t st) {
    long long blocks = (M + 127) / 128;
    if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
    UnmergeGeom g = ug ? *ug : UnmergeGeom{0, 0, 0, 0, 0};
    if (ug) {
        if (act) k_ln_rows_small<true, true, NF4><<<(int)blocks, 128, 0, st>>>(in, gamma, beta, out, M, C, eps, g);
        else k_ln_rows_small<false, true, NF4><<<(int)blocks, 128, 0, st>>>(in, gamma, beta, out, M, C, eps, g);
    } else {
        if (act) k_ln_rows_small<true, false, NF4><<<(int)blocks, 128, 0, st>>>(in, gamma, beta, out, M, C, eps, g);
        else k_ln_rows_small<false, false, NF4><<<(int)blocks, 128, 0, st>>>(in, gamma, beta, out, M, C, eps, g);
    }
}

// Wide rows (64 < C <= 768, C % 4 == 0): LPR lanes per row, NPER float4 per lane held in registers (one pass over the row,
// 16-byte accesses, 32 / LPR rows per warp) -- the scalar warp-per-row kernel above reads the row three times out of L1.
template <bool ACT, bool UNMERGE, int LPR, int NPER>
__global__ void __launch_bounds__(256) k_ln_rows_vec(const float* __restrict__ in, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                     float* __restrict__ out, long long M, int C, float eps, UnmergeGeom ug) {
    const int nf4 = C >> 2;
    const int lane = threadIdx.x & 31, l = lane & (LPR - 1), sub = lane / LPR;
    constexpr int RPW = 32 / LPR;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const float invc = 1.f / (float)C;
    for (long long r0 = warp0 * RPW; r0 < M; r0 += nwarps * RPW) {
        const long long row = r0 + sub;
        const bool ok = row < M;
        const float4* src = reinterpret_cast<const float4*>(in + (ok ? row : 0) * C);
        float4 v[NPER];
#pragma unroll
        for (int i = 0; i < NPER; i++) v[i] = (ok && l + i * LPR < nf4) ? src[l + i * LPR] : make_float4(0.f, 0.f, 0.f, 0.f);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NPER; i++) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
        for (int o = LPR >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s * invc;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NPER; i++) {
            if (l + i * LPR < nf4) {
                const float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
                q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
            }
        }
#pragma unroll
        for (int o = LPR >> 1; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = rsqrtf(q * invc + eps);
        if (!ok) continue;
        long long obase = row * C;
        int Wf = 0;
        if (UNMERGE) {   // row = (b, Y, X) of the coarse map; channel n = q*Cout + ch -> fine pixel (Y*mh+ph, X*mw+pw), channel ch
            const int X = (int)(row % ug.Wc);
            const long long p = row / ug.Wc;
            const int Y = (int)(p % ug.Hc);
            const long long b = p / ug.Hc;
            Wf = ug.Wc * ug.mw;
            obase = ((b * (ug.Hc * ug.mh) + (long long)Y * ug.mh) * Wf + (long long)X * ug.mw) * ug.Cout;
        }
#pragma unroll
        for (int i = 0; i < NPER; i++) {
            const int q4 = l + i * LPR;
            if (q4 < nf4) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + q4), bb = __ldg(reinterpret_cast<const float4*>(beta) + q4);
                float4 y = make_float4((v[i].x - mean) * rstd * g.x + bb.x, (v[i].y - mean) * rstd * g.y + bb.y,
                                       (v[i].z - mean) * rstd * g.z + bb.z, (v[i].w - mean) * rstd * g.w + bb.w);
                if (ACT) { y.x = elu1(y.x); y.y = elu1(y.y); y.z = elu1(y.z); y.w = elu1(y.w); }
                if (!UNMERGE) {
                    *reinterpret_cast<float4*>(out + obase + q4 * 4) = y;
                } else {   // Cout % 4 == 0: the four channels of a float4 belong to one fine pixel
                    const int c = q4 * 4, qq = c / ug.Cout, ch = c - qq * ug.Cout;
                    const int ph = qq / ug.mw, pw = qq - ph * ug.mw;
                    *reinterpret_cast<float4*>(out + obase + ((long long)ph * Wf + pw) * ug.Cout + ch) = y;
                }
            }
        }
    }
}

template <int LPR, int NPER>
static void launch_ln_vec(const float* in, const float* gamma, const float* beta, float* out, long long M, int C, float eps,
                          int act, const UnmergeGeom* ug, cudaStream_t st) {
    constexpr int RPW = 32 / LPR;
    long long blocks = ((M + RPW - 1) / RPW * 32 + 255) / 256;
    if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
    if (blocks < 1) blocks = 1;
    UnmergeGeom g = ug ? *ug : UnmergeGeom{0, 0, 0, 0, 0};
    if (ug) {
        if (act) k_ln_rows_vec<true, true, LPR, NPER><<<(int)blocks, 256, 0, st>>>(in, gamma, beta, out, M, C, eps, g);
        else k_ln_rows_vec<false, true, LPR, NPER><<<(int)blocks, 256, 0, st>>>(in, gamma, beta, out, M, C, eps, g);
    } else {
        if (act) k_ln_rows_vec<true, false, LPR, NPER><<<(int)blocks, 256, 0, st>>>(in, gamma, beta, out, M, C, eps, g);
        else k_ln_rows_vec<false, false, LPR, NPER><<<(int)blocks, 256, 0, st>>>(in, gamma, beta, out, M, C, eps, g);
    }
}

int launch_layernorm(const float* in, const float* gamma, const float* beta, float* out, long long M, int C, float eps,
                     int act, const UnmergeGeom* ug, cudaStream_t st) {
    const int threads = 256;
    long long blocks = (M * 32 + threads - 1) / threads;
    if (blocks > (long long)sm_count() * 32) blocks = (long long)sm_count() * 32;
    if (blocks < 1) blocks = 1;
    UnmergeGeom g = ug ? *ug : UnmergeGeom{0, 0, 0, 0, 0};
    ProfScope ps(act ? "layernorm_elu" : "layernorm", 8.0 * (double)M * C, 8.0 * (double)M * C, st);
    const bool al = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(gamma) |
                      reinterpret_cast<uintptr_t>(beta)) & 15) == 0;
    if (C % 4 == 0 && C <= 64 && al && M > 0) {
        if (C <= 16) launch_ln_small<4>(in, gamma, beta, out, M, C, eps, act, ug, st);
        else if (C <= 32) launch_ln_small<8>(in, gamma, beta, out, M, C, eps, act, ug, st);
        else launch_ln_small<16>(in, gamma, beta, out, M, C, eps, act, ug, st);
        SF_CHECK_LAUNCH("layernorm_small");
        return SF_OK;
    }
    if (C % 4 == 0 && C <= 768 && al && M > 0 && (!ug || ug->Cout % 4 == 0)) {
        const int nf4 = C >> 2;
        if (nf4 <= 24) launch_ln_vec<8, 3>(in, gamma, beta, out, M, C, eps, act, ug, st);
        else if (nf4 <= 48) launch_ln_vec<16, 3>(in, gamma, beta, out, M, C, eps, act, ug, st);
        else if (nf4 <= 96) launch_ln_vec<32, 3>(in, gamma, beta, out, M, C, eps, act, ug, st);
        else launch_ln_vec<32, 6>(in, gamma, beta, out, M, C, eps, act, ug, st);
        SF_CHECK_LAUNCH("layernorm_vec");
        return SF_OK;
    }
    if (ug) {
        if (act) k_ln_rows<true, true><<<(int)blocks, threads, 0, st>>>(in, gamma, beta, out, M, C, eps, g);
        else k_ln_rows<false, true><<<(int)blocks, threads, 0, st>>>(in, gamma, beta, out, M, C, eps, g);
    } else {
        if (act) k_ln_rows<true, false><<<(int)blocks, threads, 0, st>>>(in, gamma, beta, out, M, C, eps, g);
        else k_ln_rows<false, false><<<(int)blocks, threads, 0, st>>>(in, gamma, beta, out, M, C, eps, g);
    }
    SF_CHECK_LAUNCH("layernorm");
    return SF_OK;
}

// =============================================================================================
// SGEMM  C[M,N] = A[M,K] * W[N,K]^T (+bias)(ELU)(+residual); up to 4 problems per launch (z)
// =============================================================================================
static constexpr int BM = 64, BN = 64, BK = 16, PADM = 4;

template <bool ELU>
__global__ void __launch_bounds__(256) k_gemm_tn(GemmBatch batch, long long M, int N, int K) {
    __shared__ __align__(16) float As[BK][BM + PADM];
    __shared__ __align__(16) float Ws[BK][BN + PADM];
    const GemmProblem pb = batch.p[blockIdx.z];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int lrow = tid >> 2, lk = (tid & 3) * 4;  // each thread stages 4 consecutive k of one row
    const bool vecA = ((K & 3) == 0) && ((reinterpret_cast<uintptr_t>(pb.A) & 15) == 0);
    const bool vecW = ((K & 3) == 0) && ((reinterpret_cast<uintptr_t>(pb.W) & 15) == 0);

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += BK) {
        {   // A tile
            long long m = m0 + lrow;
            int k = k0 + lk;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (m < M) {
                const float* src = pb.A + m * K + k;
                if (vecA && k + 3 < K) {
                    float4 t = *reinterpret_cast<const float4*>(src);
                    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; e++) if (k + e < K) v[e] = src[e];
                }
            }
#pragma unroll
            for (int e = 0; e < 4; e++) As[lk + e][lrow] = v[e];
        }
        {   // W tile
            int n = n0 + lrow;
            int k = k0 + lk;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (n < N) {
                const float* src = pb.W + (long long)n * K + k;
                if (vecW && k + 3 < K) {
                    float4 t = *reinterpret_cast<const float4*>(src);
                    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; e++) if (k + e < K) v[e] = src[e];
                }
            }
#pragma unroll
            for (int e = 0; e < 4; e++) Ws[lk + e][lrow] = v[e];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; k++) {
            float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            float4 w = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
            float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        long long m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (pb.bias) v += pb.bias[n];
            if (ELU) v = elu1(v);
            if (pb.residual) v += pb.residual[m * N + n];
            pb.C[m * N + n] = v;
        }
    }
}

int launch_gemm_tn(const GemmBatch& batch, int nbatch, long long M, int N, int K, bool elu, cudaStream_t st) {
    SF_CHECK_ARG(nbatch >= 1 && nbatch <= 4 && M > 0 && N > 0 && K > 0, "gemm: bad shape M=%lld N=%d K=%d", M, N, K);
    long long gx = (M + BM - 1) / BM;
    SF_CHECK_ARG(gx <= 2147483647LL, "gemm: M too large");
    dim3 grid((unsigned)gx, (unsigned)((N + BN - 1) / BN), (unsigned)nbatch);
    double gbytes = 0.0;
    for (int z = 0; z < nbatch; z++)
        gbytes += 4.0 * ((double)M * K + (double)N * K + (double)M * N * (batch.p[z].residual ? 2.0 : 1.0));
    ProfScope ps("gemm_tn_f32", 2.0 * (double)M * N * K * nbatch, gbytes, st);
    if (elu) k_gemm_tn<true><<<grid, 256, 0, st>>>(batch, M, N, K);
    else k_gemm_tn<false><<<grid, 256, 0, st>>>(batch, M, N, K);
    SF_CHECK_LAUNCH("gemm_tn");
    return SF_OK;
}

// =============================================================================================
// Attention core (a001:317-354) on token-major Q,K,V [Mtok][inner]:
//   per (window, head): S = (Q K^T) * scale + bias ; masked -> -1e10 ; P = softmax ; O = P V
// The cyclic shift, window partition, head split, window reverse and un-shift are the index
// function win_token_src(); O is written at the token's un-shifted position.
// =============================================================================================
template <int DMAX>
__global__ void k_attn_core_f32(const float* __restrict__ Q, const float* __restrict__ Kt, const float* __restrict__ V,
                                float* __restrict__ O, const float* __restrict__ table, WinGeom g, int inner,
                                int head_dim, float scale) {
    extern __shared__ float smem[];
    const int T = g.T;
    float* Ks = smem;                       // [T][DMAX]
    float* Vs = Ks + (size_t)T * DMAX;      // [T][DMAX]
    float* tab = Vs + (size_t)T * DMAX;     // [(2wsh-1)*(2wsw-1)]
    long long* rows = reinterpret_cast<long long*>(tab + (((2 * g.wsh - 1) * (2 * g.wsw - 1) + 1) & ~1));  // [T]
    int* regs = reinterpret_cast<int*>(rows + T);  // [T]
    const int win = blockIdx.x, head = blockIdx.y;
    const int tw = 2 * g.wsw - 1;

    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        int rg;
        rows[t] = win_token_src(g, win, t, &rg);
        regs[t] = rg;
    }
    for (int i = threadIdx.x; i < (2 * g.wsh - 1) * tw; i += blockDim.x) tab[i] = table[i];
    __syncthreads();
    for (int i = threadIdx.x; i < T * DMAX; i += blockDim.x) {
        int t = i / DMAX, dd = i - t * DMAX;
        float kv = 0.f, vv = 0.f;
        if (dd < head_dim) {
            long long off = rows[t] * inner + head * head_dim + dd;
            kv = Kt[off];
            vv = V[off];
        }
        Ks[i] = kv;
        Vs[i] = vv;
    }
    __syncthreads();

    for (int qi = threadIdx.x; qi < T; qi += blockDim.x) {
        float q[DMAX];
        const long long qoff = rows[qi] * inner + head * head_dim;
#pragma unroll
        for (int dd = 0; dd < DMAX; dd++) q[dd] = dd < head_dim ? Q[qoff + dd] : 0.f;
        const int qr = qi / g.wsw, qc = qi - qr * g.wsw, qreg = regs[qi];
        // pass 1: row maximum (same order of operations as the reference: matmul, *scale, +bias, mask)
        float mx = -INFINITY;
        for (int j = 0; j < T; j++) {
            float s = 0.f;
#pragma unroll
            for (int dd = 0; dd < DMAX; dd++) s = fmaf(q[dd], Ks[j * DMAX + dd], s);
            int jr = j / g.wsw, jc = j - jr * g.wsw;
            s = s * scale + tab[(jr - qr + g.wsh - 1) * tw + (jc - qc + g.wsw - 1)];
            if (regs[j] != qreg) s = -1e10f;
            mx = fmaxf(mx, s);
        }
        // pass 2: exp, sum, P V
        float acc[DMAX];
#pragma unroll
        for (int dd = 0; dd < DMAX; dd++) acc[dd] = 0.f;
        float sum = 0.f;
        for (int j = 0; j < T; j++) {
            float s = 0.f;
#pragma unroll
            for (int dd = 0; dd < DMAX; dd++) s = fmaf(q[dd], Ks[j * DMAX + dd], s);
            int jr = j / g.wsw, jc = j - jr * g.wsw;
            s = s * scale + tab[(jr - qr + g.wsh - 1) * tw + (jc - qc + g.wsw - 1)];
            if (regs[j] != qreg) s = -1e10f;
            float p = expf(s - mx);
            sum += p;
#pragma unroll
            for (int dd = 0; dd < DMAX; dd++) acc[dd] = fmaf(p, Vs[j * DMAX + dd], acc[dd]);
        }
        float inv = 1.f / sum;
#pragma unroll
        for (int dd = 0; dd < DMAX; dd++)
            if (dd < head_dim) O[qoff + dd] = acc[dd] * inv;
    }
}

size_t attn_core_smem_bytes(const WinGeom& g, int dmax) {
    size_t tabn = ((2 * g.wsh - 1) * (2 * g.wsw - 1) + 1) & ~1;
    return (size_t)g.T * dmax * 2 * sizeof(float) + tabn * sizeof(float) + (size_t)g.T * (sizeof(long long) + sizeof(int));
}

template <int DMAX>
static int launch_attn_core_t(const float* Q, const float* K, const float* V, float* O, const float* table,
                              const WinGeom& g, int inner, int nh, int d, cudaStream_t st) {
    size_t smem = attn_core_smem_bytes(g, DMAX);
    SF_CHECK_ARG(smem <= 200 * 1024, "attention core: window of %d tokens x head_dim %d needs %zu B of shared memory", g.T, d, smem);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_attn_core_f32<DMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("attention core: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return SF_ERR_CUDA; }
    }
    long long nwin = (long long)g.B * g.nWh * g.nWw;
    SF_CHECK_ARG(nwin <= 2147483647LL && nh <= 65535, "attention core: too many windows/heads");
    int threads = g.T <= 64 ? 64 : (g.T <= 128 ? 128 : 256);
    dim3 grid((unsigned)nwin, (unsigned)nh);
    const double mtok = (double)nwin * g.T;
    ProfScope ps("attn_core_f32", 4.0 * g.T * mtok * nh * d, 16.0 * mtok * inner, st);
    k_attn_core_f32<DMAX><<<grid, threads, smem, st>>>(Q, K, V, O, table, g, inner, d, 1.0f / sqrtf((float)d));
    SF_CHECK_LAUNCH("attn_core_f32");
    return SF_OK;
}

int launch_attn_core_f32(const float* Q, const float* K, const float* V, float* O, const float* table, const WinGeom& g,
                         int inner, int nh, int d, cudaStream_t st) {
    if (d <= 4) return launch_attn_core_t<4>(Q, K, V, O, table, g, inner, nh, d, st);
    if (d <= 8) return launch_attn_core_t<8>(Q, K, V, O, table, g, inner, nh, d, st);
    if (d <= 16) return launch_attn_core_t<16>(Q, K, V, O, table, g, inner, nh, d, st);
    if (d <= 32) return launch_attn_core_t<32>(Q, K, V, O, table, g, inner, nh, d, st);
    if (d <= 64) return launch_attn_core_t<64>(Q, K, V, O, table, g, inner, nh, d, st);
    set_error("attention core: head_dim %d > 64 is not supported", d);
    return SF_ERR_INVALID;
}

// =============================================================================================
// Final head (a013:126-152): conv kxk reflect (2->2) -> BN(2) -> ELU -> conv kxk reflect (2->1)
// =============================================================================================
static constexpr int HEAD_THREADS = 256;
static constexpr int HEAD_MAXK = 7;

// conv1 -> t[b][r][c][2]; when STATS, per-block partial (sum0,sum1,sq0,sq1) -> partials[block][4]
// ELU(alpha = 1) with the exponential on ex2.approx (2^-22 relative): the head applies it 2 (eval) or 18 (train) times per pixel
__device__ __forceinline__ float head_elu(float v) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * 1.4426950408889634f));
    return v > 0.f ? v : e - 1.f;
}

// KS: compile-time kernel size (0 = use the run-time `ks`).  ACT (eval mode, running statistics): the BatchNorm affine
// (computed beforehand by k_head_bn_affine) and the ELU are applied here, once per pixel, and conv2 reads activations.
template <bool STATS, int KS, bool ACT>
__global__ void k_head_conv1(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ w1,
                             const float* __restrict__ b1, float2* __restrict__ t, float* __restrict__ partials,
                             const float* __restrict__ affine, int B, int H, int W, int ks_rt) {
    const int ks = KS ? KS : ks_rt;
    __shared__ float wsm[2 * 2 * HEAD_MAXK * HEAD_MAXK];
    __shared__ float red[4][HEAD_THREADS / 32];
    for (int i = threadIdx.x; i < 4 * ks * ks; i += blockDim.x) wsm[i] = w1[i];
    __syncthreads();
    const int pad = ks / 2;
    long long total = (long long)B * H * W;
    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % W);
        long long p = i / W;
        int r = (int)(p % H);
        long long b = p / H;
        const float* xb = x + b * H * W;
        const float* yb = y + b * H * W;
        float o0 = b1[0], o1 = b1[1];
#pragma unroll
        for (int dr = 0; dr < ks; dr++) {
            int rr = reflect_both(r + dr - pad, H);
#pragma unroll
            for (int dc = 0; dc < ks; dc++) {
                int cc = reflect_both(c + dc - pad, W);
                float xv = xb[(long long)rr * W + cc], yv = yb[(long long)rr * W + cc];
                int wi = dr * ks + dc;
                // w1[o][i][dr][dc]
                o0 = fmaf(wsm[(0 * 2 + 0) * ks * ks + wi], xv, o0);
                o0 = fmaf(wsm[(0 * 2 + 1) * ks * ks + wi], yv, o0);
                o1 = fmaf(wsm[(1 * 2 + 0) * ks * ks + wi], xv, o1);
                o1 = fmaf(wsm[(1 * 2 + 1) * ks * ks + wi], yv, o1);
            }
        }
        if (ACT) { o0 = head_elu(fmaf(o0, affine[0], affine[1])); o1 = head_elu(fmaf(o1, affine[2], affine[3])); }
        t[i] = make_float2(o0, o1);
        if (STATS) { s0 += o0; s1 += o1; q0 += o0 * o0; q1 += o1 * o1; }
    }
    if (STATS) {
        s0 = warp_sum(s0); s1 = warp_sum(s1); q0 = warp_sum(q0); q1 = warp_sum(q1);
        int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        if (lane == 0) { red[0][wid] = s0; red[1][wid] = s1; red[2][wid] = q0; red[3][wid] = q1; }
        __syncthreads();
        if (threadIdx.x < 4) {
            float a = 0.f;
            for (int k = 0; k < HEAD_THREADS / 32; k++) a += red[threadIdx.x][k];
            partials[(long long)blockIdx.x * 4 + threadIdx.x] = a;
        }
    }
}

// one block: BN scale/shift for both channels -> affine[4] = {scale0, shift0, scale1, shift1}
__global__ void k_head_bn_affine(const float* __restrict__ partials, int nblocks, long long count,
                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                 float* __restrict__ running_mean, float* __restrict__ running_var,
                                 float* __restrict__ save_mean, float* __restrict__ save_invstd, float* __restrict__ affine,
                                 float eps, float momentum, int training) {
    __shared__ double red[4][32];
    int ch = threadIdx.x >> 5, lane = threadIdx.x & 31;  // 4 warps: sum0,sum1,sq0,sq1
    if (training) {
        double a = 0.0;
        for (int i = lane; i < nblocks; i += 32) a += (double)partials[(long long)i * 4 + ch];
        red[ch][lane] = a;
        __syncthreads();
        if (threadIdx.x < 2) {
            double s = 0.0, q = 0.0;
            for (int i = 0; i < 32; i++) { s += red[threadIdx.x][i]; q += red[2 + threadIdx.x][i]; }
            double mean = s / (double)count;
            double var = q / (double)count - mean * mean;
            if (var < 0.0) var = 0.0;
            float invstd = (float)(1.0 / sqrt(var + (double)eps));
            int c = threadIdx.x;
            float sc = gamma[c] * invstd;
            affine[2 * c] = sc;
            affine[2 * c + 1] = beta[c] - (float)mean * sc;
            if (save_mean) save_mean[c] = (float)mean;
            if (save_invstd) save_invstd[c] = invstd;
            double unbiased = count > 1 ? var * (double)count / (double)(count - 1) : var;
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
        }
    } else if (threadIdx.x < 2) {
        int c = threadIdx.x;
        float invstd = 1.0f / sqrtf(running_var[c] + eps);
        float sc = gamma[c] * invstd;
        affine[2 * c] = sc;
        affine[2 * c + 1] = beta[c] - running_mean[c] * sc;
    }
}

// ACT: t holds conv1 outputs, BatchNorm affine + ELU are applied per tap (train mode: the statistics come after conv1);
// !ACT: t already holds activations
template <int KS, bool ACT>
__global__ void k_head_conv2(const float2* __restrict__ t, const float* __restrict__ affine, const float* __restrict__ w2,
                             const float* __restrict__ b2, float* __restrict__ out, int B, int H, int W, int ks_rt) {
    const int ks = KS ? KS : ks_rt;
    __shared__ float wsm[2 * HEAD_MAXK * HEAD_MAXK];
    for (int i = threadIdx.x; i < 2 * ks * ks; i += blockDim.x) wsm[i] = w2[i];
    __syncthreads();
    const float sc0 = affine[0], sh0 = affine[1], sc1 = affine[2], sh1 = affine[3];
    const int pad = ks / 2;
    long long total = (long long)B * H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % W);
        long long p = i / W;
        int r = (int)(p % H);
        long long b = p / H;
        const float2* tb = t + b * H * W;
        float o = b2[0];
#pragma unroll
        for (int dr = 0; dr < ks; dr++) {
            int rr = reflect_both(r + dr - pad, H);
#pragma unroll
            for (int dc = 0; dc < ks; dc++) {
                int cc = reflect_both(c + dc - pad, W);
                float2 v = tb[(long long)rr * W + cc];
                float a0 = ACT ? head_elu(fmaf(v.x, sc0, sh0)) : v.x, a1 = ACT ? head_elu(fmaf(v.y, sc1, sh1)) : v.y;
                o = fmaf(wsm[dr * ks + dc], a0, o);
                o = fmaf(wsm[ks * ks + dr * ks + dc], a1, o);
            }
        }
        out[i] = o;
    }
}

static int head_blocks(long long total) {
    long long blocks = (total + HEAD_THREADS - 1) / HEAD_THREADS;
    if (blocks > (long long)sm_count() * 8) blocks = (long long)sm_count() * 8;
    return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace sf

using namespace sf;

// =============================================================================================
// C ABI: fp32 implementations (dispatched from api.cu)
// =============================================================================================
namespace sf {

static inline void wa_ln_plan(const sf_window_attn_params* p, bool* need_q, bool* need_kv, bool* share) {
    const bool self_attn = (p->kv_src == p->q_src);
    const bool same_ln = self_attn && p->ln_q_gamma == p->ln_kv_gamma && p->ln_q_beta == p->ln_kv_beta;
    *need_q = p->ln_q_gamma != nullptr;
    *share = same_ln && *need_q;                 // LN(kv_src) is the same tensor as LN(q_src)
    *need_kv = p->ln_kv_gamma != nullptr && !*share;
}

size_t window_attn_ws_f32(const sf_window_attn_params* p) {
    size_t M = (size_t)p->B * p->Hp * p->Wp, inner = (size_t)p->num_heads * p->head_dim;
    bool nq, nkv, share;
    wa_ln_plan(p, &nq, &nkv, &share);
    size_t n = 4 * align_up(M * inner * sizeof(float));  // Q, K, V, O
    if (nq) n += align_up(M * p->C * sizeof(float));
    if (nkv) n += align_up(M * p->C * sizeof(float));
    return n;
}

int window_attn_fwd_f32(const sf_window_attn_params* p, void* ws_ptr, size_t ws_bytes, cudaStream_t st) {
    const long long M = (long long)p->B * p->Hp * p->Wp;
    const int C = p->C, inner = p->num_heads * p->head_dim;
    Workspace ws(ws_ptr, ws_bytes);
    float* Q = ws.take<float>((size_t)M * inner);
    float* K = ws.take<float>((size_t)M * inner);
    float* V = ws.take<float>((size_t)M * inner);
    float* O = ws.take<float>((size_t)M * inner);
    const float* nq = p->q_src;
    const float* nkv = p->kv_src;
    bool need_q, need_kv, share;
    wa_ln_plan(p, &need_q, &need_kv, &share);
    float* bufq = need_q ? ws.take<float>((size_t)M * C) : nullptr;
    float* bufkv = need_kv ? ws.take<float>((size_t)M * C) : nullptr;
    if (!Q || !K || !V || !O || (need_q && !bufq) || (need_kv && !bufkv)) {
        set_error("sf_window_attn_fwd: workspace too small (%zu B given)", ws_bytes);
        return SF_ERR_WORKSPACE;
    }
    if (need_q) {
        SF_TRY(launch_layernorm(p->q_src, p->ln_q_gamma, p->ln_q_beta, bufq, M, C, p->ln_eps, 0, nullptr, st));
        nq = bufq;
    }
    if (share) nkv = bufq;
    if (need_kv) {
        SF_TRY(launch_layernorm(p->kv_src, p->ln_kv_gamma, p->ln_kv_beta, bufkv, M, C, p->ln_eps, 0, nullptr, st));
        nkv = bufkv;
    }
    GemmBatch gb{};
    gb.p[0] = GemmProblem{nq, p->wq, p->bq, nullptr, Q};
    gb.p[1] = GemmProblem{nkv, p->wk, p->bk, nullptr, K};
    gb.p[2] = GemmProblem{nkv, p->wv, p->bv, nullptr, V};
    SF_TRY(launch_gemm_tn(gb, 3, M, inner, C, false, st));
    WinGeom g = make_geom(p->B, p->Hp, p->Wp, p->wsh, p->wsw, p->shift);
    SF_TRY(launch_attn_core_f32(Q, K, V, O, p->bias_table, g, inner, p->num_heads, p->head_dim, st));
    GemmBatch go{};
    go.p[0] = GemmProblem{O, p->wo, p->bo, p->residual, p->out};
    SF_TRY(launch_gemm_tn(go, 1, M, C, inner, false, st));
    return SF_OK;
}

size_t mlp_ws_f32(const sf_mlp_params* p) {
    size_t n = align_up((size_t)p->M * p->hidden * sizeof(float));
    if (p->ln_gamma) n += align_up((size_t)p->M * p->C * sizeof(float));
    return n;
}

int mlp_fwd_f32(const sf_mlp_params* p, void* ws_ptr, size_t ws_bytes, cudaStream_t st) {
    Workspace ws(ws_ptr, ws_bytes);
    float* Hd = ws.take<float>((size_t)p->M * p->hidden);
    float* nb = p->ln_gamma ? ws.take<float>((size_t)p->M * p->C) : nullptr;
    if (!Hd || (p->ln_gamma && !nb)) {
        set_error("sf_mlp_fwd: workspace too small (%zu B given)", ws_bytes);
        return SF_ERR_WORKSPACE;
    }
    const float* a = p->in;
    if (p->ln_gamma) {
        SF_TRY(launch_layernorm(p->in, p->ln_gamma, p->ln_beta, nb, p->M, p->C, p->ln_eps, 0, nullptr, st));
        a = nb;
    }
    GemmBatch g1{};
    g1.p[0] = GemmProblem{a, p->w1, p->b1, nullptr, Hd};
    SF_TRY(launch_gemm_tn(g1, 1, p->M, p->hidden, p->C, true, st));
    GemmBatch g2{};
    g2.p[0] = GemmProblem{Hd, p->w2, p->b2, p->residual, p->out};
    SF_TRY(launch_gemm_tn(g2, 1, p->M, p->C, p->hidden, false, st));
    return SF_OK;
}

size_t patch_ws_f32(const sf_patch_params* p) {
    int mm = p->mh * p->mw;
    if (p->encoder) {
        size_t Mo = (size_t)p->B * (p->H / p->mh) * (p->W / p->mw);
        return align_up(Mo * mm * p->Cin * sizeof(float)) + align_up(Mo * p->Cout * sizeof(float));
    }
    size_t M = (size_t)p->B * p->H * p->W;
    return align_up(M * mm * p->Cout * sizeof(float));
}

int patch_fwd_f32(const sf_patch_params* p, void* ws_ptr, size_t ws_bytes, cudaStream_t st) {
    Workspace ws(ws_ptr, ws_bytes);
    const int mm = p->mh * p->mw;
    if (p->encoder) {
        const long long Mo = (long long)p->B * (p->H / p->mh) * (p->W / p->mw);
        float* merged = ws.take<float>((size_t)Mo * mm * p->Cin);
        float* lin = ws.take<float>((size_t)Mo * p->Cout);
        if (!merged || !lin) { set_error("sf_patch_fwd: workspace too small (%zu B given)", ws_bytes); return SF_ERR_WORKSPACE; }
        SF_TRY(sf_patch_merge(p->in, merged, p->B, p->H, p->W, p->Cin, p->mh, p->mw, (void*)st));
        GemmBatch g{};
        g.p[0] = GemmProblem{merged, p->w, p->b, nullptr, lin};
        SF_TRY(launch_gemm_tn(g, 1, Mo, p->Cout, mm * p->Cin, false, st));
        SF_TRY(launch_layernorm(lin, p->ln_gamma, p->ln_beta, p->out, Mo, p->Cout, p->ln_eps, 1, nullptr, st));
    } else {
        const long long M = (long long)p->B * p->H * p->W;
        float* lin = ws.take<float>((size_t)M * mm * p->Cout);
        if (!lin) { set_error("sf_patch_fwd: workspace too small (%zu B given)", ws_bytes); return SF_ERR_WORKSPACE; }
        GemmBatch g{};
        g.p[0] = GemmProblem{p->in, p->w, p->b, nullptr, lin};
        SF_TRY(launch_gemm_tn(g, 1, M, mm * p->Cout, p->Cin, false, st));
        UnmergeGeom ug{p->H, p->W, p->mh, p->mw, p->Cout};
        SF_TRY(launch_layernorm(lin, p->ln_gamma, p->ln_beta, p->out, M, mm * p->Cout, p->ln_eps, 1, &ug, st));
    }
    return SF_OK;
}

size_t head_ws(const sf_head_params* p) {
    long long total = (long long)p->B * p->H * p->W;
    return align_up((size_t)total * 2 * sizeof(float)) + align_up((size_t)head_blocks(total) * 4 * sizeof(float)) + align_up(4 * sizeof(float));
}

int head_fwd(const sf_head_params* p, void* ws_ptr, size_t ws_bytes, cudaStream_t st) {
    long long total = (long long)p->B * p->H * p->W;
    int nb = head_blocks(total);
    Workspace ws(ws_ptr, ws_bytes);
    float2* t = ws.take<float2>((size_t)total);
    float* partials = ws.take<float>((size_t)nb * 4);
    float* affine = ws.take<float>(4);
    if (!t || !partials || !affine) { set_error("sf_head_fwd: workspace too small (%zu B given)", ws_bytes); return SF_ERR_WORKSPACE; }
    ProfScope ps("final_head", (double)total * (4.0 * 2 * p->ksize * p->ksize + 2.0 * 2 * p->ksize * p->ksize), 12.0 * (double)total, st);
    const bool k3 = p->ksize == 3;
    if (p->training) {
        // batch statistics: conv1 (+ partial sums) -> scale / shift -> conv2 applies BatchNorm + ELU per tap
        if (k3) k_head_conv1<true, 3, false><<<nb, HEAD_THREADS, 0, st>>>(p->x, p->y, p->w1, p->b1, t, partials, nullptr, p->B, p->H, p->W, 3);
        else k_head_conv1<true, 0, false><<<nb, HEAD_THREADS, 0, st>>>(p->x, p->y, p->w1, p->b1, t, partials, nullptr, p->B, p->H, p->W, p->ksize);
        SF_CHECK_LAUNCH("head_conv1");
        k_head_bn_affine<<<1, 128, 0, st>>>(partials, nb, total, p->bn_gamma, p->bn_beta, p->running_mean, p->running_var,
                                            p->save_mean, p->save_invstd, affine, p->bn_eps, p->bn_momentum, 1);
        SF_CHECK_LAUNCH("head_bn_affine");
        if (k3) k_head_conv2<3, true><<<nb, HEAD_THREADS, 0, st>>>(t, affine, p->w2, p->b2, p->out, p->B, p->H, p->W, 3);
        else k_head_conv2<0, true><<<nb, HEAD_THREADS, 0, st>>>(t, affine, p->w2, p->b2, p->out, p->B, p->H, p->W, p->ksize);
    } else {
        // running statistics: scale / shift first, BatchNorm + ELU once per pixel inside conv1, conv2 reads activations
        k_head_bn_affine<<<1, 128, 0, st>>>(partials, nb, total, p->bn_gamma, p->bn_beta, p->running_mean, p->running_var,
                                            p->save_mean, p->save_invstd, affine, p->bn_eps, p->bn_momentum, 0);
        SF_CHECK_LAUNCH("head_bn_affine");
        if (k3) k_head_conv1<false, 3, true><<<nb, HEAD_THREADS, 0, st>>>(p->x, p->y, p->w1, p->b1, t, partials, affine, p->B, p->H, p->W, 3);
        else k_head_conv1<false, 0, true><<<nb, HEAD_THREADS, 0, st>>>(p->x, p->y, p->w1, p->b1, t, partials, affine, p->B, p->H, p->W, p->ksize);
        SF_CHECK_LAUNCH("head_conv1");
        if (k3) k_head_conv2<3, false><<<nb, HEAD_THREADS, 0, st>>>(t, affine, p->w2, p->b2, p->out, p->B, p->H, p->W, 3);
        else k_head_conv2<0, false><<<nb, HEAD_THREADS, 0, st>>>(t, affine, p->w2, p->b2, p->out, p->B, p->H, p->W, p->ksize);
    }
    SF_CHECK_LAUNCH("head_conv2");
    return SF_OK;
}

}  // namespace sf

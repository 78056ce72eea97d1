// Fused KV-cached decoder step, version 2  (reference math: ResidualDecoderAttentionBlock::forward
// src/model/mod.rs:345-350, MultiHead{Self,Cross}Attention::forward :428-436/:482-490, qkv_attention
// :493-533, MLP::forward :376-382, TextDecoder::forward :141-157, beamsearch_next
// src/transcribe.rs:271-276).
//
// decoder.cu runs one step as ~9 small kernels per layer; at batch 3 each of them is pure launch +
// prologue latency.  Here a layer is THREE kernels, each ending in a "last block resolves" reduction
// (atomic ticket, deterministic summation order) that leaves the complete residual stream x in HBM/L2:
//
//   self block   grid (H, row groups): LN(x) -> q,k,v of ONE head (192 weight rows) -> append k,v to
//                the cache -> softmax(q K^T) V over the row's ancestry -> this head's slice of the
//                out-projection (W_o[:, h*64:(h+1)*64]) -> partial y[h];   resolve: x += b_o + sum_h y[h]
//   cross block  grid (H, splits, row groups): LN(x) -> q of one head -> split-KV attention over the
//                window's encoder K/V (unnormalised o, m, l) -> W_o slice -> partial y[h][s];
//                resolve: x += b_o + sum_h (sum_s e^(m_s-M) y[h][s]) / (sum_s e^(m_s-M) l_s)
//   mlp block    grid (4d/32, row groups): LN(x) -> 32 hidden features -> GELU -> their 32 columns of
//                W_2 -> partial y[c];                                        resolve: x += b_2 + sum_c y[c]
//   logits       persistent grid: LN(x) -> tied-embedding GEMV (the dominant kernel: V*d weights) with
//                fused special-token mask and per-CTA online (max, sum-exp, top candidates)
//   finish       per row: merge partials -> log-softmax of the candidates -> k best (ties -> lower id,
//                what beam.rs:81-110 keeps) -> greedy bookkeeping, position advance
//
// All arithmetic is fp32 (weights fp16-exact or fp32), so results match decoder.cu / the oracle up to
// fp32 summation order.
#include <cfloat>
#include <climits>

#include "decoder.h"
#include "wb_internal.h"

namespace wb {

namespace {

constexpr int NT = 256;   // threads per CTA
constexpr int NW = 8;     // warps per CTA

__device__ __forceinline__ float gelu_erf(float x) {
    const float t = __fadd_rn(erff(__fdiv_rn(x, 1.41421356237309504880f)), 1.0f);
    return __fdiv_rn(__fmul_rn(x, t), 2.0f);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ void load8(const __half* p, float (&w)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h[i]);
        w[2 * i] = f.x;
        w[2 * i + 1] = f.y;
    }
}
__device__ __forceinline__ void load8(const float* p, float (&w)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
    w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}

// LayerNorm of RC rows of x into shared memory (burn 0.9 form, see encoder.cu); one warp per row.
template <int RC>
__device__ __forceinline__ void ln_rows(const float* __restrict__ x, int r0, int R, int d, const float* __restrict__ g,
                                        const float* __restrict__ b, float eps, int eps_outside, float* xs) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int rr = warp; rr < RC; rr += NW) {
        float* xr = xs + rr * d;
        const int r = r0 + rr;
        if (r >= R) {
            for (int c = lane; c < d; c += 32) xr[c] = 0.0f;
            continue;
        }
        const float* src = x + (int64_t)r * d;
        float s = 0.0f;
        for (int c = lane; c < d; c += 32) {
            const float v = __ldcg(src + c);
            xr[c] = v;
            s += v;
        }
        s = warp_sum(s);
        const float mean = __fdiv_rn(s, (float)d);
        float q = 0.0f;
        for (int c = lane; c < d; c += 32) {
            const float dv = __fsub_rn(xr[c], mean);
            xr[c] = dv;
            q = __fadd_rn(q, __fmul_rn(dv, dv));
        }
        q = warp_sum(q);
        const float var = __fdiv_rn(q, (float)d);
        const float den = eps_outside ? __fadd_rn(__fsqrt_rn(var), eps) : __fsqrt_rn(__fadd_rn(var, eps));
        for (int c = lane; c < d; c += 32) xr[c] = __fadd_rn(__fmul_rn(__fdiv_rn(xr[c], den), g[c]), b[c]);
    }
}

// acc[q][rr] = sum_k W[row_q][k] * xs[rr][k] for NF weight rows at once (all lanes get the sums)
template <typename WT, int RC, int NF>
__device__ __forceinline__ void dot_rows(const WT* (&wrow)[NF], const float* xs, int K, float (&acc)[NF][RC]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int q = 0; q < NF; ++q)
#pragma unroll
        for (int rr = 0; rr < RC; ++rr) acc[q][rr] = 0.0f;
    for (int v = lane; v < K / 8; v += 32) {
        float w[NF][8];
#pragma unroll
        for (int q = 0; q < NF; ++q) load8(wrow[q] + v * 8, w[q]);
#pragma unroll
        for (int rr = 0; rr < RC; ++rr) {
            const float4 x0 = *reinterpret_cast<const float4*>(xs + rr * K + v * 8);
            const float4 x1 = *reinterpret_cast<const float4*>(xs + rr * K + v * 8 + 4);
            const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
            for (int q = 0; q < NF; ++q)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[q][rr] = fmaf(w[q][i], xv[i], acc[q][rr]);
        }
    }
#pragma unroll
    for (int q = 0; q < NF; ++q)
#pragma unroll
        for (int rr = 0; rr < RC; ++rr) acc[q][rr] = warp_sum(acc[q][rr]);
}

// y[rr][n] = sum_{c<KS} W[n][col0 + c] * a[rr][c] for all n < N (a slice of the columns of W);
// LPR lanes cooperate on one weight row (KS values = LPR * 8).  Output: ypart[rr*ld + n].
template <typename WT, int RC, int KS>
__device__ __forceinline__ void slice_matvec(const WT* __restrict__ W, int N, int K, int col0, const float* a_s /*[RC][KS]*/,
                                             float* __restrict__ ypart, int64_t row_ld, int r0, int R) {
    constexpr int LPR = KS / 8;           // lanes per weight row
    constexpr int RPI = 32 / LPR;         // weight rows per warp instruction
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane / LPR, l = lane % LPR;
    for (int nb = warp * RPI; nb < N; nb += NW * RPI) {
        const int n = nb + sub;
        float w[8];
        if (n < N) load8(W + (int64_t)n * K + col0 + l * 8, w);
        else {
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = 0.0f;
        }
#pragma unroll
        for (int rr = 0; rr < RC; ++rr) {
            const float4 a0 = *reinterpret_cast<const float4*>(a_s + rr * KS + l * 8);
            const float4 a1 = *reinterpret_cast<const float4*>(a_s + rr * KS + l * 8 + 4);
            float s = w[0] * a0.x;
            s = fmaf(w[1], a0.y, s); s = fmaf(w[2], a0.z, s); s = fmaf(w[3], a0.w, s);
            s = fmaf(w[4], a1.x, s); s = fmaf(w[5], a1.y, s); s = fmaf(w[6], a1.z, s); s = fmaf(w[7], a1.w, s);
#pragma unroll
            for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (l == 0 && n < N && r0 + rr < R) ypart[(int64_t)(r0 + rr) * row_ld + n] = s;
        }
    }
}

// Returns true in exactly one CTA of the grid: the last one to arrive.  All partial results
// written before the call are visible to that CTA (read them with __ldcg).
__device__ __forceinline__ bool last_block_ticket(unsigned int* counter, unsigned int total) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(counter, 1u);
        s_last = (prev == total - 1u) ? 1 : 0;
        if (s_last) *counter = 0u;   // re-arm for the next launch (nobody else touches it any more)
    }
    __syncthreads();
    if (s_last) __threadfence();
    return s_last != 0;
}

// ------------------------------------------------------------------------------------------------
// self-attention block
template <typename WT, int RC>
__global__ void __launch_bounds__(NT)
dec2_self_kernel(const Dec2SelfArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int d = a.d, h = blockIdx.x, r0 = blockIdx.y * RC, R = a.R;
    const int p = *a.pos;
    float* xs = sm;                       // [RC][d]
    float* qs = xs + RC * d;              // [RC][64]
    float* ks = qs + RC * 64;             // [RC][64]  new key (scaled)
    float* vs = ks + RC * 64;             // [RC][64]  new value
    float* at = vs + RC * 64;             // [RC][64]  attention output
    float* sc = at + RC * 64;             // [RC][t_max + 1] scores / weights
    float* wm = sc + RC * (a.t_max + 1);  // [RC][NW] per-warp max
    float* wl = wm + RC * NW;             // [RC][NW] per-warp sum
    float* wo = wl + RC * NW;             // [RC][NW][64] per-warp unnormalised output
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const WT* Wqkv = reinterpret_cast<const WT*>(a.Wqkv);
    const WT* Wo = reinterpret_cast<const WT*>(a.Wo);

    ln_rows<RC>(a.x, r0, R, d, a.ln_g, a.ln_b, a.ln_eps, a.eps_outside, xs);
    __syncthreads();
    // ---- q, k, v of head h: 192 weight rows, 4 at a time per warp
    for (int f0 = warp * 4; f0 < 192; f0 += NW * 4) {
        const WT* rows[4];
        int nn[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int f = f0 + q;
            nn[q] = (f >> 6) * d + h * 64 + (f & 63);
            rows[q] = Wqkv + (int64_t)nn[q] * d;
        }
        float acc[4][RC];
        dot_rows<WT, RC, 4>(rows, xs, d, acc);
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int rr = 0; rr < RC; ++rr)
                if (lane == q * RC + rr) {
                    const int f = f0 + q, part = f >> 6, j = f & 63;
                    float v = __fadd_rn(acc[q][rr], a.bqkv[nn[q]]);
                    if (part < 2) v = __fmul_rn(v, a.qk_scale);
                    (part == 0 ? qs : (part == 1 ? ks : vs))[rr * 64 + j] = v;
                    const int r = r0 + rr;
                    if (part > 0 && r < R) {
                        float* cache = part == 1 ? a.kc : a.vc;
                        cache[((int64_t)r * a.t_max + p) * d + h * 64 + j] = v;
                    }
                }
    }
    __syncthreads();
    // ---- attention of each row over positions 0..p (position p comes from shared memory)
    constexpr int WPR = NW / RC;   // warps per row
    {
        const int rr = warp / WPR, ws = warp % WPR;
        const int r = r0 + rr;
        const int n_keys = p + 1;
        const int* anc = (a.anc && r < R) ? a.anc + (int64_t)r * a.t_max : nullptr;
        const float* qrow = qs + rr * 64;
        float lmax = -INFINITY;
        if (r < R) {
            for (int j = ws * 32 + lane; j < n_keys; j += WPR * 32) {
                float s = 0.0f;
                if (j == p) {
#pragma unroll 16
                    for (int c = 0; c < 64; ++c) s = fmaf(qrow[c], ks[rr * 64 + c], s);
                } else {
                    const int src = anc ? anc[j] : r;
                    const float4* kp = reinterpret_cast<const float4*>(a.kc + ((int64_t)src * a.t_max + j) * d + h * 64);
                    float4 kv[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c) kv[c] = __ldcg(kp + c);
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const float4 q4 = *reinterpret_cast<const float4*>(qrow + c * 4);
                        s = fmaf(q4.x, kv[c].x, s); s = fmaf(q4.y, kv[c].y, s);
                        s = fmaf(q4.z, kv[c].z, s); s = fmaf(q4.w, kv[c].w, s);
                    }
                }
                sc[rr * (a.t_max + 1) + j] = s;
                lmax = fmaxf(lmax, s);
            }
        }
        lmax = warp_max(lmax);
        // weights relative to this warp's max, then weighted values: lane owns dims lane, lane+32
        float l = 0.0f, o0 = 0.0f, o1 = 0.0f;
        if (r < R && lmax > -INFINITY) {
            for (int jb = ws * 32; jb < n_keys; jb += WPR * 32) {
                const int jn = min(32, n_keys - jb);
                __syncwarp();
                for (int jj = 0; jj < jn; ++jj) {
                    const int j = jb + jj;
                    const float e = expf(sc[rr * (a.t_max + 1) + j] - lmax);
                    l += e;
                    float v0, v1;
                    if (j == p) {
                        v0 = vs[rr * 64 + lane];
                        v1 = vs[rr * 64 + 32 + lane];
                    } else {
                        const int src = anc ? anc[j] : r;
                        const float* vp = a.vc + ((int64_t)src * a.t_max + j) * d + h * 64;
                        v0 = __ldcg(vp + lane);
                        v1 = __ldcg(vp + 32 + lane);
                    }
                    o0 = fmaf(e, v0, o0);
                    o1 = fmaf(e, v1, o1);
                }
            }
        }
        if (lane == 0) { wm[rr * NW + ws] = lmax; wl[rr * NW + ws] = l; }
        wo[(rr * NW + ws) * 64 + lane] = o0;
        wo[(rr * NW + ws) * 64 + 32 + lane] = o1;
    }
    __syncthreads();
    for (int i = tid; i < RC * 64; i += NT) {
        const int rr = i >> 6, c = i & 63;
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < WPR; ++w) M = fmaxf(M, wm[rr * NW + w]);
        float L = 0.0f, o = 0.0f;
#pragma unroll
        for (int w = 0; w < WPR; ++w) {
            const float m = wm[rr * NW + w];
            const float sc_w = m > -INFINITY ? expf(m - M) : 0.0f;
            L += sc_w * wl[rr * NW + w];
            o += sc_w * wo[(rr * NW + w) * 64 + c];
        }
        at[i] = (r0 + rr < R) ? __fdiv_rn(o, L) : 0.0f;
    }
    __syncthreads();
    // ---- this head's 64 columns of the out-projection
    slice_matvec<WT, RC, 64>(Wo, d, d, h * 64, at, a.ypart + (int64_t)h * R * d, d, r0, R);
    // ---- resolve: x += b_o + sum_h y[h]
    if (last_block_ticket(a.counter, gridDim.x * gridDim.y)) {
        const int H = gridDim.x;
        for (int i = tid; i < R * d; i += NT) {
            float s = 0.0f;
            for (int hh = 0; hh < H; ++hh) s += __ldcg(a.ypart + (int64_t)hh * R * d + i);
            a.x[i] = __fadd_rn(a.x[i], __fadd_rn(s, a.bo[i % d]));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// cross-attention block
template <typename WT, int RC>
__global__ void __launch_bounds__(NT)
dec2_cross_kernel(const Dec2CrossArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int d = a.d, h = blockIdx.x, sp = blockIdx.y, r0 = blockIdx.z * RC, R = a.R;
    const int H = gridDim.x, S = gridDim.y;
    float* xs = sm;                        // [RC][d]
    float* qs = xs + RC * d;               // [RC][64]
    float* at = qs + RC * 64;              // [RC][64]  unnormalised split output
    float* sc = at + RC * 64;              // [RC][KMAX] scores
    float* wm = sc + RC * a.kmax;          // [RC][NW]
    float* wl = wm + RC * NW;
    float* wo = wl + RC * NW;              // [RC][NW][64]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const WT* Wq = reinterpret_cast<const WT*>(a.Wq);
    const WT* Wo = reinterpret_cast<const WT*>(a.Wo);

    ln_rows<RC>(a.x, r0, R, d, a.ln_g, a.ln_b, a.ln_eps, a.eps_outside, xs);
    __syncthreads();
    for (int f0 = warp * 4; f0 < 64; f0 += NW * 4) {
        const WT* rows[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) rows[q] = Wq + (int64_t)(h * 64 + f0 + q) * d;
        float acc[4][RC];
        dot_rows<WT, RC, 4>(rows, xs, d, acc);
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int rr = 0; rr < RC; ++rr)
                if (lane == q * RC + rr)
                    qs[rr * 64 + f0 + q] = __fmul_rn(__fadd_rn(acc[q][rr], a.bq[h * 64 + f0 + q]), a.qk_scale);
    }
    __syncthreads();
    constexpr int WPR = NW / RC;
    {
        const int rr = warp / WPR, ws = warp % WPR;
        const int r = r0 + rr;
        float lmax = -INFINITY, l = 0.0f, o0 = 0.0f, o1 = 0.0f;
        int kb = 0, nk = 0;
        const float* kbase = nullptr;
        if (r < R) {
            const int w = a.row_window[r];
            const int T = a.win_T[w];
            const int per = (T + S - 1) / S;
            kb = sp * per;
            nk = max(0, min(T, kb + per) - kb);
            kbase = a.ckv + (a.win_row_off[w] + kb) * (int64_t)(2 * d) + h * 64;   // K at +0, V at +d
        }
        const float* qrow = qs + rr * 64;
        for (int j = ws * 32 + lane; j < nk; j += WPR * 32) {
            const float4* kp = reinterpret_cast<const float4*>(kbase + (int64_t)j * 2 * d);
            float4 kv[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) kv[c] = __ldg(kp + c);
            float s = 0.0f;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const float4 q4 = *reinterpret_cast<const float4*>(qrow + c * 4);
                s = fmaf(q4.x, kv[c].x, s); s = fmaf(q4.y, kv[c].y, s);
                s = fmaf(q4.z, kv[c].z, s); s = fmaf(q4.w, kv[c].w, s);
            }
            sc[rr * a.kmax + j] = s;
            lmax = fmaxf(lmax, s);
        }
        lmax = warp_max(lmax);
        if (lmax > -INFINITY) {
            for (int jb = ws * 32; jb < nk; jb += WPR * 32) {
                const int jn = min(32, nk - jb);
                __syncwarp();
                int jj = 0;
                for (; jj + 4 <= jn; jj += 4) {   // 4 keys in flight
                    float e[4], v0[4], v1[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float* vp = kbase + (int64_t)(jb + jj + u) * 2 * d + d;
                        v0[u] = __ldg(vp + lane);
                        v1[u] = __ldg(vp + 32 + lane);
                        e[u] = expf(sc[rr * a.kmax + jb + jj + u] - lmax);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        l += e[u];
                        o0 = fmaf(e[u], v0[u], o0);
                        o1 = fmaf(e[u], v1[u], o1);
                    }
                }
                for (; jj < jn; ++jj) {
                    const float* vp = kbase + (int64_t)(jb + jj) * 2 * d + d;
                    const float e = expf(sc[rr * a.kmax + jb + jj] - lmax);
                    l += e;
                    o0 = fmaf(e, __ldg(vp + lane), o0);
                    o1 = fmaf(e, __ldg(vp + 32 + lane), o1);
                }
            }
        }
        if (lane == 0) { wm[rr * NW + ws] = lmax; wl[rr * NW + ws] = l; }
        wo[(rr * NW + ws) * 64 + lane] = o0;
        wo[(rr * NW + ws) * 64 + 32 + lane] = o1;
    }
    __syncthreads();
    for (int i = tid; i < RC * 64; i += NT) {
        const int rr = i >> 6, c = i & 63;
        float M = -INFINITY;
#pragma unroll
        for (int w = 0; w < WPR; ++w) M = fmaxf(M, wm[rr * NW + w]);
        float L = 0.0f, o = 0.0f;
#pragma unroll
        for (int w = 0; w < WPR; ++w) {
            const float m = wm[rr * NW + w];
            const float sc_w = m > -INFINITY ? expf(m - M) : 0.0f;
            L += sc_w * wl[rr * NW + w];
            o += sc_w * wo[(rr * NW + w) * 64 + c];
        }
        at[i] = o;   // unnormalised, relative to M
        if (c == 0 && r0 + rr < R) {
            a.part_m[((int64_t)h * S + sp) * R + r0 + rr] = M;
            a.part_l[((int64_t)h * S + sp) * R + r0 + rr] = L;
        }
    }
    __syncthreads();
    slice_matvec<WT, RC, 64>(Wo, d, d, h * 64, at, a.ypart + ((int64_t)h * S + sp) * R * d, d, r0, R);
    // ---- resolve: x += b_o + sum_h (sum_s e^(m_s - M_h) y[h][s]) / (sum_s e^(m_s - M_h) l_s)
    if (last_block_ticket(a.counter, gridDim.x * gridDim.y * gridDim.z)) {
        float* wgt = sm;   // [H][S][R] normalised weights (shared memory is free now)
        for (int i = tid; i < H * R; i += NT) {
            const int hh = i / R, r = i % R;
            float M = -INFINITY;
            for (int s = 0; s < S; ++s) M = fmaxf(M, __ldcg(a.part_m + ((int64_t)hh * S + s) * R + r));
            float L = 0.0f;
            for (int s = 0; s < S; ++s) {
                const float m = __ldcg(a.part_m + ((int64_t)hh * S + s) * R + r);
                const float e = m > -INFINITY ? expf(m - M) : 0.0f;
                wgt[(hh * S + s) * R + r] = e;
                L += e * __ldcg(a.part_l + ((int64_t)hh * S + s) * R + r);
            }
            for (int s = 0; s < S; ++s) wgt[(hh * S + s) * R + r] = __fdiv_rn(wgt[(hh * S + s) * R + r], L);
        }
        __syncthreads();
        for (int i = tid; i < R * d; i += NT) {
            const int r = i / d;
            float s = 0.0f;
            for (int hs = 0; hs < H * S; ++hs) s = fmaf(wgt[hs * R + r], __ldcg(a.ypart + (int64_t)hs * R * d + i), s);
            a.x[i] = __fadd_rn(a.x[i], __fadd_rn(s, a.bo[i % d]));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// MLP block: 32 hidden features per CTA
constexpr int MLP_SLICE = 32;

template <typename WT, int RC>
__global__ void __launch_bounds__(NT)
dec2_mlp_kernel(const Dec2MlpArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int d = a.d, c0 = blockIdx.x * MLP_SLICE, r0 = blockIdx.y * RC, R = a.R;
    float* xs = sm;                 // [RC][d]
    float* hs = xs + RC * d;        // [RC][32]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const WT* W1 = reinterpret_cast<const WT*>(a.W1);
    const WT* W2 = reinterpret_cast<const WT*>(a.W2);
    ln_rows<RC>(a.x, r0, R, d, a.ln_g, a.ln_b, a.ln_eps, a.eps_outside, xs);
    __syncthreads();
    {
        const int f0 = warp * 4;    // 8 warps x 4 = 32 features
        const WT* rows[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) rows[q] = W1 + (int64_t)(c0 + f0 + q) * d;
        float acc[4][RC];
        dot_rows<WT, RC, 4>(rows, xs, d, acc);
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int rr = 0; rr < RC; ++rr)
                if (lane == q * RC + rr) hs[rr * MLP_SLICE + f0 + q] = gelu_erf(__fadd_rn(acc[q][rr], a.b1[c0 + f0 + q]));
    }
    __syncthreads();
    slice_matvec<WT, RC, MLP_SLICE>(W2, d, 4 * d, c0, hs, a.ypart + (int64_t)blockIdx.x * R * d, d, r0, R);
    if (last_block_ticket(a.counter, gridDim.x * gridDim.y)) {
        const int NC = gridDim.x;
        for (int i = tid; i < R * d; i += NT) {
            float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
            int c = 0;
            for (; c + 4 <= NC; c += 4) {
                s0 += __ldcg(a.ypart + (int64_t)(c + 0) * R * d + i);
                s1 += __ldcg(a.ypart + (int64_t)(c + 1) * R * d + i);
                s2 += __ldcg(a.ypart + (int64_t)(c + 2) * R * d + i);
                s3 += __ldcg(a.ypart + (int64_t)(c + 3) * R * d + i);
            }
            for (; c < NC; ++c) s0 += __ldcg(a.ypart + (int64_t)c * R * d + i);
            a.x[i] = __fadd_rn(a.x[i], __fadd_rn((s0 + s1) + (s2 + s3), a.b2[i % d]));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// logits: persistent CTAs over the vocabulary + online softmax partials + top candidates
template <int KC>
struct Cand {
    float v[KC];
    int i[KC];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int k = 0; k < KC; ++k) { v[k] = -INFINITY; i[k] = INT_MAX; }
    }
    // keep the KC best by (value desc, index asc)
    __device__ __forceinline__ void push(float val, int idx) {
        if (!(val > v[KC - 1] || (val == v[KC - 1] && idx < i[KC - 1]))) return;
        v[KC - 1] = val;
        i[KC - 1] = idx;
#pragma unroll
        for (int k = KC - 1; k > 0; --k) {
            const bool better = v[k] > v[k - 1] || (v[k] == v[k - 1] && i[k] < i[k - 1]);
            if (better) {
                const float tv = v[k]; v[k] = v[k - 1]; v[k - 1] = tv;
                const int ti = i[k]; i[k] = i[k - 1]; i[k - 1] = ti;
            }
        }
    }
};

template <typename WT, int RC>
__global__ void __launch_bounds__(NT)
dec2_logits_kernel(const Dec2LogitsArgs a) {
    extern __shared__ __align__(16) float sm[];
    constexpr int KC = DEC2_KC;
    const int d = a.d, R = a.R, V = a.V;
    float* xs = sm;                                   // [RC][d]
    float* red_m = xs + RC * d;                       // [NW][RC]
    float* red_s = red_m + NW * RC;                   // [NW][RC]
    float* red_v = red_s + NW * RC;                   // [NW][RC][KC]
    int* red_i = reinterpret_cast<int*>(red_v + NW * RC * KC);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const WT* E = reinterpret_cast<const WT*>(a.E);
    const int p = *a.pos;
    const bool use_mask = a.is_special != nullptr && (a.mask_mode == 1 || (a.mask_mode == 2 && p + 1 <= 5));

    for (int r0 = 0; r0 < R; r0 += RC) {
        __syncthreads();
        ln_rows<RC>(a.x, r0, R, d, a.ln_g, a.ln_b, a.ln_eps, a.eps_outside, xs);
        __syncthreads();
        float m_run = -INFINITY, s_run = 0.0f;   // lane rr (< RC) tracks row rr
        Cand<KC> cand;
        cand.init();
        const int n_groups = (V + 3) / 4;
        for (int grp = blockIdx.x * NW + warp; grp < n_groups; grp += gridDim.x * NW) {
            const int n0 = grp * 4;
            const WT* rows[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) rows[q] = E + (int64_t)min(n0 + q, V - 1) * d;
            float acc[4][RC];
            dot_rows<WT, RC, 4>(rows, xs, d, acc);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int n = n0 + q;
                if (n >= V) continue;
                const bool masked = use_mask && a.is_special[n];
#pragma unroll
                for (int rr = 0; rr < RC; ++rr) {
                    if (lane == rr && r0 + rr < R) {
                        const float raw = acc[q][rr];
                        if (a.logits_out) a.logits_out[(int64_t)(r0 + rr) * V + n] = raw;
                        const float v = masked ? __fadd_rn(raw, -INFINITY) : raw;
                        if (v > -INFINITY) {
                            if (v > m_run) { s_run = s_run * expf(m_run - v) + 1.0f; m_run = v; }
                            else s_run += expf(v - m_run);
                        }
                        cand.push(v, n);
                    }
                }
            }
        }
        if (lane < RC) {
            red_m[warp * RC + lane] = m_run;
            red_s[warp * RC + lane] = s_run;
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                red_v[(warp * RC + lane) * KC + k] = cand.v[k];
                red_i[(warp * RC + lane) * KC + k] = cand.i[k];
            }
        }
        __syncthreads();
        if (tid < RC && r0 + tid < R) {
            const int rr = tid;
            float M = -INFINITY;
            for (int w = 0; w < NW; ++w) M = fmaxf(M, red_m[w * RC + rr]);
            float S = 0.0f;
            Cand<KC> best;
            best.init();
            for (int w = 0; w < NW; ++w) {
                const float m = red_m[w * RC + rr];
                if (m > -INFINITY) S += red_s[w * RC + rr] * expf(m - M);
                for (int k = 0; k < KC; ++k) best.push(red_v[(w * RC + rr) * KC + k], red_i[(w * RC + rr) * KC + k]);
            }
            const int64_t o = (int64_t)blockIdx.x * R + r0 + rr;
            a.part_m[o] = M;
            a.part_s[o] = S;
#pragma unroll
            for (int k = 0; k < KC; ++k) { a.part_v[o * KC + k] = best.v[k]; a.part_i[o * KC + k] = best.i[k]; }
        }
    }
}

// finish: one CTA per row merges the logits partials
__global__ void __launch_bounds__(NT)
dec2_finish_kernel(const Dec2FinishArgs a) {
    constexpr int KC = DEC2_KC;
    __shared__ float s_f[NW];
    __shared__ int s_i[NW];
    __shared__ float s_bf;
    __shared__ int s_bi;
    const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int R = gridDim.x, NP = a.n_parts;
    const int p = *a.pos;
    float mx = -INFINITY;
    for (int c = tid; c < NP; c += NT) mx = fmaxf(mx, a.part_m[(int64_t)c * R + r]);
    mx = warp_max(mx);
    if (lane == 0) s_f[warp] = mx;
    __syncthreads();
    mx = s_f[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) mx = fmaxf(mx, s_f[w]);
    __syncthreads();
    float se = 0.0f;
    for (int c = tid; c < NP; c += NT) {
        const float m = a.part_m[(int64_t)c * R + r];
        if (m > -INFINITY) se += a.part_s[(int64_t)c * R + r] * expf(m - mx);
    }
    se = warp_sum(se);
    if (lane == 0) s_f[warp] = se;
    __syncthreads();
    se = 0.0f;
#pragma unroll
    for (int w = 0; w < NW; ++w) se += s_f[w];
    const float lse = logf(se);
    __syncthreads();
    // candidates -> log-probs (x - max) - lse, ranked by (log-prob desc, id asc)
    float prev_v = INFINITY;
    int prev_i = -1;
    for (int kk = 0; kk < a.k; ++kk) {
        float bv = -INFINITY;
        int bi = INT_MAX;
        for (int c = tid; c < NP * KC; c += NT) {
            const int part = c / KC, k = c % KC;
            const int idx = a.part_i[((int64_t)part * R + r) * KC + k];
            if (idx == INT_MAX) continue;
            const float v = __fsub_rn(__fsub_rn(a.part_v[((int64_t)part * R + r) * KC + k], mx), lse);
            const bool after_prev = v < prev_v || (v == prev_v && idx > prev_i);
            if (after_prev && (v > bv || (v == bv && idx < bi))) { bv = v; bi = idx; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { s_f[warp] = bv; s_i[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < NW; ++w)
                if (s_f[w] > bv || (s_f[w] == bv && s_i[w] < bi)) { bv = s_f[w]; bi = s_i[w]; }
            s_bf = bv;
            s_bi = bi;
            a.topk_id[(int64_t)r * a.k + kk] = bi == INT_MAX ? -1 : bi;
            a.topk_lp[(int64_t)r * a.k + kk] = bv;
        }
        __syncthreads();
        prev_v = s_bf;
        prev_i = s_bi;
        __syncthreads();
    }
    if (tid == 0) {
        if (a.greedy && !a.finished[r]) {   // beam_search with beam_size 1 (beam.rs:9-37)
            const int best = a.topk_id[(int64_t)r * a.k];
            a.tokens[(int64_t)r * a.t_max + p + 1] = best;
            a.lengths[r] = p + 2;
            a.cur_tok[r] = best;
            if (best == a.eot) a.finished[r] = 1;
        }
        __threadfence();
        const unsigned int prev = atomicAdd(a.counter, 1u);
        if (prev == (unsigned int)R - 1u) {   // last row: advance the position, count unfinished rows
            *a.counter = 0u;
            __threadfence();
            int c = 0;
            for (int rr = 0; rr < R; ++rr) c += a.greedy ? (__ldcg(a.finished + rr) ? 0 : 1) : 1;
            *a.n_unfinished = c;
            *a.pos = p + 1;
        }
    }
}

template <typename KernelT>
void set_smem(KernelT k, size_t smem) {
    if (smem > 48 * 1024) WB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
}

template <typename WT, int RC>
void launch_self_t(const Dec2SelfArgs& a, int H, cudaStream_t st) {
    const size_t smem = sizeof(float) * ((size_t)RC * a.d + 4 * RC * 64 + (size_t)RC * (a.t_max + 1) + 2 * RC * NW + (size_t)RC * NW * 64);
    auto k = dec2_self_kernel<WT, RC>;
    set_smem(k, smem);
    k<<<dim3(H, (a.R + RC - 1) / RC), NT, smem, st>>>(a);
    WB_LAUNCH_CHECK();
}
template <typename WT, int RC>
void launch_cross_t(const Dec2CrossArgs& a, int H, int S, cudaStream_t st) {
    size_t fl = (size_t)RC * a.d + 2 * RC * 64 + (size_t)RC * a.kmax + 2 * RC * NW + (size_t)RC * NW * 64;
    fl = std::max(fl, (size_t)H * S * a.R);   // resolve weights reuse the buffer
    const size_t smem = sizeof(float) * fl;
    auto k = dec2_cross_kernel<WT, RC>;
    set_smem(k, smem);
    k<<<dim3(H, S, (a.R + RC - 1) / RC), NT, smem, st>>>(a);
    WB_LAUNCH_CHECK();
}
template <typename WT, int RC>
void launch_mlp_t(const Dec2MlpArgs& a, cudaStream_t st) {
    const size_t smem = sizeof(float) * ((size_t)RC * a.d + RC * MLP_SLICE);
    auto k = dec2_mlp_kernel<WT, RC>;
    set_smem(k, smem);
    k<<<dim3(4 * a.d / MLP_SLICE, (a.R + RC - 1) / RC), NT, smem, st>>>(a);
    WB_LAUNCH_CHECK();
}
template <typename WT, int RC>
void launch_logits_t(const Dec2LogitsArgs& a, int n_ctas, cudaStream_t st) {
    const size_t smem = sizeof(float) * ((size_t)RC * a.d + 2 * NW * RC + 2 * (size_t)NW * RC * DEC2_KC);
    auto k = dec2_logits_kernel<WT, RC>;
    set_smem(k, smem);
    k<<<n_ctas, NT, smem, st>>>(a);
    WB_LAUNCH_CHECK();
}

int pick_rc(int R) { return R <= 1 ? 1 : (R <= 2 ? 2 : (R <= 4 ? 4 : 8)); }

}  // namespace

#define WB_DISPATCH(FN, ...)                                         \
    do {                                                             \
        const int rc_ = pick_rc(a.R);                                \
        if (w_half) {                                                \
            if (rc_ == 1) FN<__half, 1>(__VA_ARGS__);                \
            else if (rc_ == 2) FN<__half, 2>(__VA_ARGS__);           \
            else if (rc_ == 4) FN<__half, 4>(__VA_ARGS__);           \
            else FN<__half, 8>(__VA_ARGS__);                         \
        } else {                                                     \
            if (rc_ == 1) FN<float, 1>(__VA_ARGS__);                 \
            else if (rc_ == 2) FN<float, 2>(__VA_ARGS__);            \
            else if (rc_ == 4) FN<float, 4>(__VA_ARGS__);            \
            else FN<float, 8>(__VA_ARGS__);                          \
        }                                                            \
    } while (0)

void launch_dec2_self(const Dec2SelfArgs& a, int H, bool w_half, cudaStream_t st) { WB_DISPATCH(launch_self_t, a, H, st); }
void launch_dec2_cross(const Dec2CrossArgs& a, int H, int S, bool w_half, cudaStream_t st) {
    WB_DISPATCH(launch_cross_t, a, H, S, st);
}
void launch_dec2_mlp(const Dec2MlpArgs& a, bool w_half, cudaStream_t st) { WB_DISPATCH(launch_mlp_t, a, st); }
void launch_dec2_logits(const Dec2LogitsArgs& a, int n_ctas, bool w_half, cudaStream_t st) {
    WB_DISPATCH(launch_logits_t, a, n_ctas, st);
}
void launch_dec2_finish(const Dec2FinishArgs& a, int R, cudaStream_t st) {
    dec2_finish_kernel<<<R, NT, 0, st>>>(a);
    WB_LAUNCH_CHECK();
}

}  // namespace wb

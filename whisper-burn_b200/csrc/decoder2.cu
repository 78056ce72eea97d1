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

// Dot products of a block of weight rows with the RC activation rows in shared memory.
// 8 lanes share one weight row (16 bytes each per step), so a warp instruction covers 4 rows and
// the cross-lane reduction is 3 shuffles for 4 rows at once.  G row-groups are processed together
// and the K loop is unrolled so that >= 2*G independent 16-byte loads are in flight per lane.
// After the call lane (sub*8) holds acc[g][rr] for row (group g, sub).
template <typename WT, int RC, int G>
__device__ __forceinline__ void dot_groups(const WT* (&wrow)[G], const float* xs, int K, float (&acc)[G][RC]) {
    const int l = threadIdx.x & 7;
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
        for (int rr = 0; rr < RC; ++rr) acc[g][rr] = 0.0f;
#pragma unroll 2
    for (int k0 = l * 8; k0 < K; k0 += 64) {
        float w[G][8];
#pragma unroll
        for (int g = 0; g < G; ++g) load8(wrow[g] + k0, w[g]);
#pragma unroll
        for (int rr = 0; rr < RC; ++rr) {
            const float4 x0 = *reinterpret_cast<const float4*>(xs + rr * K + k0);
            const float4 x1 = *reinterpret_cast<const float4*>(xs + rr * K + k0 + 4);
            const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
            for (int g = 0; g < G; ++g)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[g][rr] = fmaf(w[g][i], xv[i], acc[g][rr]);
        }
    }
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
        for (int rr = 0; rr < RC; ++rr) {
            float v = acc[g][rr];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            acc[g][rr] = v;
        }
}

// y[rr][n] = sum_{c<KS} W[n][col0 + c] * a[rr][c] for all n < N (a column slice of W, KS = 32 or 64);
// KS/8 lanes per weight row, U row-groups in flight.  Output: ypart[(r0+rr)*row_ld + n].
template <typename WT, int RC, int KS>
__device__ __forceinline__ void slice_matvec(const WT* __restrict__ W, int N, int K, int col0, const float* a_s /*[RC][KS]*/,
                                             float* __restrict__ ypart, int64_t row_ld, int r0, int R) {
    constexpr int LPR = KS / 8;           // lanes per weight row
    constexpr int RPI = 32 / LPR;         // weight rows per warp instruction
    constexpr int U = 4;                  // row-groups in flight
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane / LPR, l = lane % LPR;
    float av[RC][8];
#pragma unroll
    for (int rr = 0; rr < RC; ++rr) {
        const float4 a0 = *reinterpret_cast<const float4*>(a_s + rr * KS + l * 8);
        const float4 a1 = *reinterpret_cast<const float4*>(a_s + rr * KS + l * 8 + 4);
        av[rr][0] = a0.x; av[rr][1] = a0.y; av[rr][2] = a0.z; av[rr][3] = a0.w;
        av[rr][4] = a1.x; av[rr][5] = a1.y; av[rr][6] = a1.z; av[rr][7] = a1.w;
    }
    for (int nb = warp * RPI; nb < N; nb += NW * RPI * U) {
        float w[U][8];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int n = nb + u * NW * RPI + sub;
            if (n < N) load8(W + (int64_t)n * K + col0 + l * 8, w[u]);
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) w[u][i] = 0.0f;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int n = nb + u * NW * RPI + sub;
#pragma unroll
            for (int rr = 0; rr < RC; ++rr) {
                float s = w[u][0] * av[rr][0];
#pragma unroll
                for (int i = 1; i < 8; ++i) s = fmaf(w[u][i], av[rr][i], s);
#pragma unroll
                for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (l == 0 && n < N && r0 + rr < R) ypart[(int64_t)(r0 + rr) * row_ld + n] = s;
            }
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

// One-pass (online softmax) attention of one query row over n keys; 4 lanes per key (16 dims each),
// 8 keys per warp step, 2 steps unrolled -> 16 independent 16-byte loads in flight per lane.
// key j lives at kptr(j) / vptr(j) (64 floats, 16-byte aligned); keys j = first, first+stride, ...
// Result: every lane holds (m, l) of the warp and its 16 dims of the unnormalised output o[16]
// (dims (lane&3)*16 .. +15), identical across the 8 key sub-groups.
struct AttnAcc {
    float m, l, o[16];
};
template <typename KF, typename VF>
__device__ __forceinline__ void attn_warp(const float* qrow, int n_keys, int first, int stride, KF&& kptr, VF&& vptr,
                                          AttnAcc& A) {
    const int lane = threadIdx.x & 31, sub = lane >> 2, l4 = lane & 3;
    float q[16];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float4 t = *reinterpret_cast<const float4*>(qrow + l4 * 16 + c * 4);
        q[c * 4] = t.x; q[c * 4 + 1] = t.y; q[c * 4 + 2] = t.z; q[c * 4 + 3] = t.w;
    }
    A.m = -INFINITY;
    A.l = 0.0f;
#pragma unroll
    for (int c = 0; c < 16; ++c) A.o[c] = 0.0f;
    constexpr int UK = 2;
    for (int jb = first; jb < n_keys; jb += stride * 8 * UK) {   // warp-uniform trip count
        float4 kk[UK][4], vv[UK][4];
        bool ok[UK];
#pragma unroll
        for (int u = 0; u < UK; ++u) {
            const int j = jb + (u * 8 + sub) * stride;
            ok[u] = j < n_keys;
            if (ok[u]) {
                const float4* kp = reinterpret_cast<const float4*>(kptr(j)) + l4 * 4;
                const float4* vp = reinterpret_cast<const float4*>(vptr(j)) + l4 * 4;
#pragma unroll
                for (int c = 0; c < 4; ++c) { kk[u][c] = __ldcg(kp + c); vv[u][c] = __ldcg(vp + c); }
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) { kk[u][c] = make_float4(0.f, 0.f, 0.f, 0.f); vv[u][c] = kk[u][c]; }
            }
        }
#pragma unroll
        for (int u = 0; u < UK; ++u) {
            float s = 0.0f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                s = fmaf(q[c * 4], kk[u][c].x, s); s = fmaf(q[c * 4 + 1], kk[u][c].y, s);
                s = fmaf(q[c * 4 + 2], kk[u][c].z, s); s = fmaf(q[c * 4 + 3], kk[u][c].w, s);
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (ok[u]) {
                const float mn = fmaxf(A.m, s);
                const float corr = expf(A.m - mn);   // exp(-inf) = 0 on the first key
                const float e = expf(s - mn);
                A.l = A.l * corr + e;
                A.m = mn;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    A.o[c * 4] = fmaf(e, vv[u][c].x, A.o[c * 4] * corr);
                    A.o[c * 4 + 1] = fmaf(e, vv[u][c].y, A.o[c * 4 + 1] * corr);
                    A.o[c * 4 + 2] = fmaf(e, vv[u][c].z, A.o[c * 4 + 2] * corr);
                    A.o[c * 4 + 3] = fmaf(e, vv[u][c].w, A.o[c * 4 + 3] * corr);
                }
            }
        }
    }
    // merge the 8 key sub-groups of the warp (lanes with equal lane&3)
#pragma unroll
    for (int off = 4; off < 32; off <<= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, A.m, off);
        const float l2 = __shfl_xor_sync(0xffffffffu, A.l, off);
        const float mn = fmaxf(A.m, m2);
        const float c1 = A.m > -INFINITY ? expf(A.m - mn) : 0.0f;
        const float c2 = m2 > -INFINITY ? expf(m2 - mn) : 0.0f;
        A.l = A.l * c1 + l2 * c2;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const float o2 = __shfl_xor_sync(0xffffffffu, A.o[c], off);
            A.o[c] = A.o[c] * c1 + o2 * c2;
        }
        A.m = mn;
    }
}

// Combines the per-warp results of one row (WPR warps) into at[rr][64]; normalised iff `normalise`.
// wm/wl: [RC][NW], wo: [RC][NW][64] in shared memory; returns (M, L) through smem scalars.
template <int RC>
__device__ __forceinline__ void attn_store_warp(const AttnAcc& A, int rr, int ws, float* wm, float* wl, float* wo) {
    const int lane = threadIdx.x & 31;
    if (lane < 4) {
#pragma unroll
        for (int c = 0; c < 16; ++c) wo[(rr * NW + ws) * 64 + lane * 16 + c] = A.o[c];
    }
    if (lane == 0) { wm[rr * NW + ws] = A.m; wl[rr * NW + ws] = A.l; }
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
    float* at = qs + RC * 64;             // [RC][64]  attention output
    float* wm = at + RC * 64;             // [RC][NW] per-warp max
    float* wl = wm + RC * NW;             // [RC][NW] per-warp sum
    float* wo = wl + RC * NW;             // [RC][NW][64] per-warp unnormalised output
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sub = lane >> 3, l8 = lane & 7;
    const WT* Wqkv = reinterpret_cast<const WT*>(a.Wqkv);
    const WT* Wo = reinterpret_cast<const WT*>(a.Wo);

    ln_rows<RC>(a.x, r0, R, d, a.ln_g, a.ln_b, a.ln_eps, a.eps_outside, xs);
    __syncthreads();
    // ---- q, k, v of head h: 3 blocks of 64 contiguous weight rows; a warp takes 2 row-groups (8 rows) at a time
    for (int f0 = warp * 8; f0 < 192; f0 += NW * 8) {
        const WT* rows[2];
        int nn[2], ff[2];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            ff[g] = f0 + g * 4 + sub;
            nn[g] = (ff[g] >> 6) * d + h * 64 + (ff[g] & 63);
            rows[g] = Wqkv + (int64_t)nn[g] * d;
        }
        float acc[2][RC];
        dot_groups<WT, RC, 2>(rows, xs, d, acc);
        if (l8 == 0) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const int part = ff[g] >> 6, j = ff[g] & 63;
                const float bias = a.bqkv[nn[g]];
#pragma unroll
                for (int rr = 0; rr < RC; ++rr) {
                    float v = __fadd_rn(acc[g][rr], bias);
                    if (part < 2) v = __fmul_rn(v, a.qk_scale);
                    const int r = r0 + rr;
                    if (part == 0) qs[rr * 64 + j] = v;
                    else if (r < R) (part == 1 ? a.kc : a.vc)[((int64_t)r * a.t_max + p) * d + h * 64 + j] = v;
                }
            }
        }
    }
    __syncthreads();   // q in smem, this step's k/v visible in global memory to the whole CTA
    // ---- attention of each row over positions 0..p of its ancestry
    constexpr int WPR = NW / RC;   // warps per row
    {
        const int rr = warp / WPR, ws = warp % WPR;
        const int r = r0 + rr;
        if (rr < RC) {
            AttnAcc A;
            if (r < R) {
                const int* anc = a.anc ? a.anc + (int64_t)r * a.t_max : nullptr;
                const float* kc = a.kc + h * 64;
                const float* vc = a.vc + h * 64;
                const int t_max = a.t_max;
                auto kp = [&](int j) { return kc + ((int64_t)((anc && j < p) ? anc[j] : r) * t_max + j) * d; };
                auto vp = [&](int j) { return vc + ((int64_t)((anc && j < p) ? anc[j] : r) * t_max + j) * d; };
                attn_warp(qs + rr * 64, p + 1, ws, WPR, kp, vp, A);
            } else {
                A.m = -INFINITY; A.l = 0.0f;
#pragma unroll
                for (int c = 0; c < 16; ++c) A.o[c] = 0.0f;
            }
            attn_store_warp<RC>(A, rr, ws, wm, wl, wo);
        }
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
    float* wm = at + RC * 64;              // [RC][NW]
    float* wl = wm + RC * NW;
    float* wo = wl + RC * NW;              // [RC][NW][64]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sub = lane >> 3, l8 = lane & 7;
    const WT* Wq = reinterpret_cast<const WT*>(a.Wq);
    const WT* Wo = reinterpret_cast<const WT*>(a.Wo);

    ln_rows<RC>(a.x, r0, R, d, a.ln_g, a.ln_b, a.ln_eps, a.eps_outside, xs);
    __syncthreads();
    {   // q of head h: 64 rows = 16 row-groups, 2 per warp
        const int f0 = warp * 8;
        const WT* rows[2];
        int ff[2];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            ff[g] = h * 64 + f0 + g * 4 + sub;
            rows[g] = Wq + (int64_t)ff[g] * d;
        }
        float acc[2][RC];
        dot_groups<WT, RC, 2>(rows, xs, d, acc);
        if (l8 == 0) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const float bias = a.bq[ff[g]];
#pragma unroll
                for (int rr = 0; rr < RC; ++rr) qs[rr * 64 + (ff[g] - h * 64)] = __fmul_rn(__fadd_rn(acc[g][rr], bias), a.qk_scale);
            }
        }
    }
    __syncthreads();
    constexpr int WPR = NW / RC;
    {
        const int rr = warp / WPR, ws = warp % WPR;
        const int r = r0 + rr;
        if (rr < RC) {
            AttnAcc A;
            if (r < R) {
                const int w = a.row_window[r];
                const int T = a.win_T[w];
                const int per = (T + S - 1) / S;
                const int kb = sp * per;
                const int nk = max(0, min(T, kb + per) - kb);
                const float* kbase = a.ckv + (a.win_row_off[w] + kb) * (int64_t)(2 * d) + h * 64;   // K at +0, V at +d
                const int64_t ld = 2 * (int64_t)d;
                auto kp = [&](int j) { return kbase + j * ld; };
                auto vp = [&](int j) { return kbase + j * ld + d; };
                attn_warp(qs + rr * 64, nk, ws, WPR, kp, vp, A);
            } else {
                A.m = -INFINITY; A.l = 0.0f;
#pragma unroll
                for (int c = 0; c < 16; ++c) A.o[c] = 0.0f;
            }
            attn_store_warp<RC>(A, rr, ws, wm, wl, wo);
        }
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
            float s0 = 0.0f, s1 = 0.0f;
            int hs = 0;
            for (; hs + 2 <= H * S; hs += 2) {
                s0 = fmaf(wgt[hs * R + r], __ldcg(a.ypart + (int64_t)hs * R * d + i), s0);
                s1 = fmaf(wgt[(hs + 1) * R + r], __ldcg(a.ypart + (int64_t)(hs + 1) * R * d + i), s1);
            }
            for (; hs < H * S; ++hs) s0 = fmaf(wgt[hs * R + r], __ldcg(a.ypart + (int64_t)hs * R * d + i), s0);
            a.x[i] = __fadd_rn(a.x[i], __fadd_rn(s0 + s1, a.bo[i % d]));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// MLP block: 64 hidden features per CTA
constexpr int MLP_SLICE = 64;

template <typename WT, int RC>
__global__ void __launch_bounds__(NT)
dec2_mlp_kernel(const Dec2MlpArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int d = a.d, c0 = blockIdx.x * MLP_SLICE, r0 = blockIdx.y * RC, R = a.R;
    float* xs = sm;                 // [RC][d]
    float* hs = xs + RC * d;        // [RC][64]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sub = lane >> 3, l8 = lane & 7;
    const WT* W1 = reinterpret_cast<const WT*>(a.W1);
    const WT* W2 = reinterpret_cast<const WT*>(a.W2);
    ln_rows<RC>(a.x, r0, R, d, a.ln_g, a.ln_b, a.ln_eps, a.eps_outside, xs);
    __syncthreads();
    {   // 64 features = 16 row-groups, 2 per warp
        const int f0 = warp * 8;
        const WT* rows[2];
        int ff[2];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            ff[g] = c0 + f0 + g * 4 + sub;
            rows[g] = W1 + (int64_t)ff[g] * d;
        }
        float acc[2][RC];
        dot_groups<WT, RC, 2>(rows, xs, d, acc);
        if (l8 == 0) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const float bias = a.b1[ff[g]];
#pragma unroll
                for (int rr = 0; rr < RC; ++rr) hs[rr * MLP_SLICE + (ff[g] - c0)] = gelu_erf(__fadd_rn(acc[g][rr], bias));
            }
        }
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
    constexpr int G = 2;                              // row-groups (4 rows each) per warp step
    const int d = a.d, R = a.R, V = a.V;
    float* xs = sm;                                   // [RC][d]
    float* red_m = xs + RC * d;                       // [NW][RC]
    float* red_s = red_m + NW * RC;                   // [NW][RC]
    float* red_v = red_s + NW * RC;                   // [NW][RC][KC]
    int* red_i = reinterpret_cast<int*>(red_v + NW * RC * KC);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sub = lane >> 3;
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
        const int n_steps = (V + 4 * G - 1) / (4 * G);
        for (int stp = blockIdx.x * NW + warp; stp < n_steps; stp += gridDim.x * NW) {
            const int n0 = stp * 4 * G;
            const WT* rows[G];
#pragma unroll
            for (int g = 0; g < G; ++g) rows[g] = E + (int64_t)min(n0 + g * 4 + sub, V - 1) * d;
            float acc[G][RC];
            dot_groups<WT, RC, G>(rows, xs, d, acc);
            // lane (sub*8) holds the 4*G logits of this step for every row; hand row rr to lane rr
#pragma unroll
            for (int g = 0; g < G; ++g)
#pragma unroll
                for (int s4 = 0; s4 < 4; ++s4) {
                    const int n = n0 + g * 4 + s4;
                    const bool masked = n < V && use_mask && a.is_special[n];
#pragma unroll
                    for (int rr = 0; rr < RC; ++rr) {
                        const float raw = __shfl_sync(0xffffffffu, acc[g][rr], s4 * 8);
                        if (lane == rr && n < V && r0 + rr < R) {
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
    const size_t smem = sizeof(float) * ((size_t)RC * a.d + 2 * RC * 64 + 2 * RC * NW + (size_t)RC * NW * 64);
    auto k = dec2_self_kernel<WT, RC>;
    set_smem(k, smem);
    k<<<dim3(H, (a.R + RC - 1) / RC), NT, smem, st>>>(a);
    WB_LAUNCH_CHECK();
}
template <typename WT, int RC>
void launch_cross_t(const Dec2CrossArgs& a, int H, int S, cudaStream_t st) {
    size_t fl = (size_t)RC * a.d + 2 * RC * 64 + 2 * RC * NW + (size_t)RC * NW * 64;
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

int pick_rc(int R) { return R <= 1 ? 1 : (R <= 2 ? 2 : (R <= 4 ? 4 : 8)); }   // NW % RC == 0

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

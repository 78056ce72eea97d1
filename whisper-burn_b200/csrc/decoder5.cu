// Batched persistent decoder on the tensor cores (9..32 rows per launch: small.en / medium / large batches and beams).
//
// Same single-launch structure, stage list and reference math as decoder3.cu (TextDecoder::forward
// src/model/mod.rs:131-157, blocks :345-350, attention :428-533, MLP :376-382, search closure
// src/transcribe.rs:253-307), but every linear layer is a swap-AB tensor-core product instead of a per-row
// FMA GEMV, so the weights are streamed ONCE per step for all rows:
//   * a warp owns a 16-feature tile of W[N][K] (fp16, exact) and multiplies it with ALL rows of the batch:
//     mma.sync.m16n8k16 with M = 16 output features, N = 8 batch rows per n-tile (up to 4 n-tiles), K = 16;
//   * the fp32 activations are split into fp16 hi + fp16 (lo * 2^11) planes (22 mantissa bits; products with
//     the fp16 weights are exact, accumulation is fp32) and staged in shared memory in FRAGMENT ORDER, so a B
//     fragment is one conflict-free 16-byte shared load; the A fragments come straight from global memory
//     as 16-byte loads (a K permutation inside each 32-column chunk makes 8 consecutive halves of a weight
//     row the a0..a3 registers of two MMAs), prefetched BEFORE the grid barrier that precedes the stage;
//   * the 8 warps of a CTA split K; partial tiles are reduced through shared memory in a fixed order;
//   * MLP2 (K = 4d) is split over CTAs into 4 slabs of d: the partial sums are folded into x by the next
//     consumer of x (deterministic order), so every stage is ONE dependent round trip behind its barrier;
//   * logits: E[V][d] is streamed once through the same MMA path into an L2-resident [R][V] buffer; a second
//     stage turns it into per-slice (max, sum-exp, top candidates) records, the finish stage is decoder3's.
// Requirements: fp16-exact weights, d % 256 == 0, R <= 32.  Everything else falls back to decoder3.cu.
#include <cooperative_groups.h>
#include <cuda_fp16.h>

#include "dec_common.cuh"

namespace wb {

namespace {

constexpr int RED_LD = 20;   // padded feature stride of the cross-warp reduction buffer (conflict-free fragment stores)
constexpr int MAXCH = 5;     // 32-column chunks per warp and slab: slab = d <= 1280 -> d / 256 <= 5
constexpr int GC = 4;        // logits: chunks per prefetch group

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

// 4 consecutive fp32 values -> fp16 hi and fp16 (residual * 2048); x == hi + lo / 2048 up to 2^-23 |x|
__device__ __forceinline__ void split4(const float4 v, uint2& hi, uint2& lo) {
    const __half2 h01 = __floats2half2_rn(v.x, v.y), h23 = __floats2half2_rn(v.z, v.w);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn((v.x - f01.x) * 2048.0f, (v.y - f01.y) * 2048.0f);
    const __half2 l23 = __floats2half2_rn((v.z - f23.x) * 2048.0f, (v.w - f23.y) * 2048.0f);
    hi = make_uint2(h2_bits(h01), h2_bits(h23));
    lo = make_uint2(h2_bits(l01), h2_bits(l23));
}

// Fragment-order planes: element (row, col) of the staged [8*NT8][KS] activation slab lives in the uint4
//   ((row / 8) * nchunks + col / 32) * 32 + (row % 8) * 4 + (col % 32) / 8,   halves (col % 8)
// i.e. lane (g = row % 8, t) of the MMA finds the 8 halves x[row][chunk*32 + t*8 .. +8) in ONE 16-byte word.
__device__ __forceinline__ void store_frag(uint4* xhi, uint4* xlo, int nchunks, int row, int col, const float4 v) {
    uint2 hi, lo;
    split4(v, hi, lo);
    const int idx = ((row >> 3) * nchunks + (col >> 5)) * 32 + (row & 7) * 4 + ((col & 31) >> 3);
    const int half = (col & 7) >> 2;
    reinterpret_cast<uint2*>(xhi + idx)[half] = hi;
    reinterpret_cast<uint2*>(xlo + idx)[half] = lo;
}

// LayerNorm (burn 0.9 form, dec_common.cuh stage_ln) of all 8*NT8 rows into the planes; warp per row, the row
// stays in registers.  load(r, c4) returns the float4 at columns 4*c4.. of row r; rows >= R are zero.
template <int NT8, typename LoadF>
__device__ __forceinline__ void stage_ln_frag(LoadF&& load, int R, int d, const float* __restrict__ g, const float* __restrict__ b,
                                              float eps, int eps_outside, uint4* xhi, uint4* xlo) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nv = d / 4, nchunks = d / 32;
    for (int r = warp; r < NT8 * 8; r += NW) {
        if (r >= R) {
            for (int c = lane; c < nv; c += 32) store_frag(xhi, xlo, nchunks, r, c * 4, make_float4(0.f, 0.f, 0.f, 0.f));
            continue;
        }
        float4 v[LN_V4];
#pragma unroll
        for (int i = 0; i < LN_V4; ++i) {
            const int c = i * 32 + lane;
            v[i] = c < nv ? load(r, c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float sum = 0.0f;
#pragma unroll
        for (int i = 0; i < LN_V4; ++i) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        sum = warp_sum(sum);
        const float mean = __fdiv_rn(sum, (float)d);
        float q = 0.0f;
#pragma unroll
        for (int i = 0; i < LN_V4; ++i) {
            const int c = i * 32 + lane;
            if (c < nv) {
                v[i].x = __fsub_rn(v[i].x, mean); v[i].y = __fsub_rn(v[i].y, mean);
                v[i].z = __fsub_rn(v[i].z, mean); v[i].w = __fsub_rn(v[i].w, mean);
                q = __fadd_rn(q, __fmul_rn(v[i].x, v[i].x)); q = __fadd_rn(q, __fmul_rn(v[i].y, v[i].y));
                q = __fadd_rn(q, __fmul_rn(v[i].z, v[i].z)); q = __fadd_rn(q, __fmul_rn(v[i].w, v[i].w));
            }
        }
        q = warp_sum(q);
        const float var = __fdiv_rn(q, (float)d);
        const float den = eps_outside ? __fadd_rn(__fsqrt_rn(var), eps) : __fsqrt_rn(__fadd_rn(var, eps));
#pragma unroll
        for (int i = 0; i < LN_V4; ++i) {
            const int c = i * 32 + lane;
            if (c < nv) {
                const float4 g4 = __ldg(reinterpret_cast<const float4*>(g) + c);
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(b) + c);
                float4 o;
                o.x = __fadd_rn(__fmul_rn(__fdiv_rn(v[i].x, den), g4.x), b4.x);
                o.y = __fadd_rn(__fmul_rn(__fdiv_rn(v[i].y, den), g4.y), b4.y);
                o.z = __fadd_rn(__fmul_rn(__fdiv_rn(v[i].z, den), g4.z), b4.z);
                o.w = __fadd_rn(__fmul_rn(__fdiv_rn(v[i].w, den), g4.w), b4.w);
                store_frag(xhi, xlo, nchunks, r, c * 4, o);
            }
        }
    }
}

// plain copy of src[r][col0 .. col0 + KS) (row stride ld, L2) into the planes; 8 independent loads in flight
template <int NT8>
__device__ __forceinline__ void stage_copy_frag(const float* src, int64_t ld, int col0, int KS, int R, uint4* xhi, uint4* xlo) {
    const int per_row = KS / 4, n4 = NT8 * 8 * per_row, nchunks = KS / 32;
    for (int i0 = threadIdx.x; i0 < n4; i0 += NT * 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * NT;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < n4) {
                const int r = i / per_row, c = i % per_row;
                if (r < R) v[u] = __ldcg(reinterpret_cast<const float4*>(src + (int64_t)r * ld + col0) + c);
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * NT;
            if (i < n4) store_frag(xhi, xlo, nchunks, i / per_row, (i % per_row) * 4, v[u]);
        }
    }
}

// A fragments of one (16-feature tile, K slice of this warp): rows g and g+8, MAXCH chunks of 32 columns
struct AFrag {
    uint4 lo[MAXCH], hi[MAXCH];   // "lo" = feature row g, "hi" = feature row g + 8
    float bias;
};
// item = slab * n_tiles + tile; this warp's slice of the slab = columns [warp * KS / 8, +KS / 8)
__device__ __forceinline__ void load_afrag(const __half* __restrict__ W, const float* __restrict__ bias, int K, int KS, int n_tiles,
                                           int item, AFrag& f) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int tile = item % n_tiles, slab = item / n_tiles;
    const int nch = KS >> 8;
    const int k0 = slab * KS + warp * (KS >> 3) + t * 8;
    const uint4* p0 = reinterpret_cast<const uint4*>(W + (int64_t)(tile * 16 + g) * K + k0);
    const uint4* p1 = reinterpret_cast<const uint4*>(W + (int64_t)(tile * 16 + g + 8) * K + k0);
#pragma unroll
    for (int c = 0; c < MAXCH; ++c) {
        if (c < nch) {
            f.lo[c] = __ldg(p0 + c * 4);
            f.hi[c] = __ldg(p1 + c * 4);
        }
    }
    f.bias = (bias != nullptr && slab == 0) ? __ldg(bias + tile * 16 + (threadIdx.x & 15)) : 0.0f;
}

// One linear layer for all rows: out[r][n] = sum_k x[r][k] W[n][k] (+ bias on slab 0).
//   stage(slab): fills the planes with columns [slab*KS, +KS) of the input (called between two __syncthreads);
//   pre(n, r)  : optional early load (e.g. the residual) issued before the MMAs;
//   emit(n, r, slab, value, pre_value): called for r < R by the thread that owns (feature n, row r).
template <int NT8, typename StageF, typename PreF, typename EmitF>
__device__ __forceinline__ void gemm_phase(const __half* __restrict__ W, const float* __restrict__ bias, int N, int K, int KS, int R,
                                           AFrag& pf, StageF&& stage, PreF&& pre, EmitF&& emit, uint4* xhi, uint4* xlo, float* red) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int n_tiles = N / 16, n_items = n_tiles * (K / KS);
    const int nch = KS >> 8, nchunks = KS >> 5;
    constexpr int NE = (16 * NT8 * 8 + NT - 1) / NT;
    int staged_slab = -1;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int tile = item % n_tiles, slab = item / n_tiles;
        AFrag cur = pf;
        if (item + (int)gridDim.x < n_items) load_afrag(W, bias, K, KS, n_tiles, item + gridDim.x, pf);   // next item of this CTA
        float prev[NE];
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            const int idx = tid + e * NT;
            prev[e] = (idx < 16 * NT8 * 8 && (idx >> 4) < R) ? pre(tile * 16 + (idx & 15), idx >> 4) : 0.0f;
        }
        if (slab != staged_slab) {
            __syncthreads();
            stage(slab);
            staged_slab = slab;
        }
        __syncthreads();
        float ah[NT8][4], al[NT8][4];
#pragma unroll
        for (int j = 0; j < NT8; ++j)
#pragma unroll
            for (int c = 0; c < 4; ++c) { ah[j][c] = 0.0f; al[j][c] = 0.0f; }
#pragma unroll
        for (int c = 0; c < MAXCH; ++c) {
            if (c < nch) {
                const int chunk = warp * nch + c;
#pragma unroll
                for (int j = 0; j < NT8; ++j) {
                    const uint4 bh = xhi[(j * nchunks + chunk) * 32 + lane];
                    const uint4 bl = xlo[(j * nchunks + chunk) * 32 + lane];
                    mma16816(ah[j], cur.lo[c].x, cur.hi[c].x, cur.lo[c].y, cur.hi[c].y, bh.x, bh.y);
                    mma16816(ah[j], cur.lo[c].z, cur.hi[c].z, cur.lo[c].w, cur.hi[c].w, bh.z, bh.w);
                    mma16816(al[j], cur.lo[c].x, cur.hi[c].x, cur.lo[c].y, cur.hi[c].y, bl.x, bl.y);
                    mma16816(al[j], cur.lo[c].z, cur.hi[c].z, cur.lo[c].w, cur.hi[c].w, bl.z, bl.w);
                }
            }
        }
        // C fragment: c0,c1 -> (feature g, rows 2t, 2t+1), c2,c3 -> (feature g+8, rows 2t, 2t+1)
        float* my = red + warp * (NT8 * 8 * RED_LD);
#pragma unroll
        for (int j = 0; j < NT8; ++j) {
            const int r0 = j * 8 + 2 * t;
            my[r0 * RED_LD + g] = fmaf(al[j][0], 1.0f / 2048.0f, ah[j][0]);
            my[(r0 + 1) * RED_LD + g] = fmaf(al[j][1], 1.0f / 2048.0f, ah[j][1]);
            my[r0 * RED_LD + g + 8] = fmaf(al[j][2], 1.0f / 2048.0f, ah[j][2]);
            my[(r0 + 1) * RED_LD + g + 8] = fmaf(al[j][3], 1.0f / 2048.0f, ah[j][3]);
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            const int idx = tid + e * NT;
            if (idx < 16 * NT8 * 8) {
                const int f = idx & 15, r = idx >> 4;
                float s = 0.0f;
#pragma unroll
                for (int w = 0; w < NW; ++w) s += red[w * (NT8 * 8 * RED_LD) + r * RED_LD + f];
                if (r < R) emit(tile * 16 + f, r, slab, s + cur.bias, prev[e]);
            }
        }
    }
    __syncthreads();
}

// =====================================================================================================
template <int NT8, int KC, typename KVT>
__global__ void __launch_bounds__(NT, 1)
dec5_kernel(const Dec3Args a) {
    extern __shared__ __align__(16) float sm[];
    const int d = a.d, H = a.H, L = a.L, V = a.V, R = a.R, t_max = a.t_max;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gw = blockIdx.x * NW + warp, n_gw = gridDim.x * NW;
    constexpr int RP = NT8 * 8;
    uint4* xhi = reinterpret_cast<uint4*>(sm);            // [NT8][d/32][32] fragment-order fp16 hi plane
    uint4* xlo = xhi + NT8 * (d / 32) * 32;               // same, residual * 2^11
    float* red = reinterpret_cast<float*>(xlo + NT8 * (d / 32) * 32);   // [NW][RP][RED_LD]; also cross-merge weights
    float* qs = red + NW * RP * RED_LD;                   // [64]
    float* wm = qs + 64;                                  // [NW]
    float* wl = wm + NW;                                  // [NW]
    float* wo = wl + NW;                                  // [NW][64]
    float* ao = wo + NW * 64;                             // [64]
    float* ML = ao + 64;                                  // [2]
    int* tok_s = reinterpret_cast<int*>(ML + 2);          // [RP]
    unsigned int gen = 0;
    int tr_n = 0;
    WB_TRACE();
    const float scale = a.qk_scale;
    const int S = a.n_splits;
    const float* yp = a.ypart;                            // [4][R][d] MLP2 partial sums of the previous layer
    const int64_t yps = (int64_t)R * d;
    auto no_pre = [](int, int) { return 0.0f; };
    // The residual stream ping-pongs between two buffers: the MLP2 partial sums of layer l-1 are folded into x by
    // the LayerNorm staging of layer l (every CTA, fixed order); CTA r % grid publishes the folded row into the
    // OTHER buffer, so no CTA can read a row that was already folded.  xc = current buffer.
    float* xc = a.x;
    float* xo = a.x2;
    auto fold = [&](int r, int c4) {
        float4 v = __ldcg(reinterpret_cast<const float4*>(xc + (int64_t)r * d) + c4);
        const float4 p0 = __ldcg(reinterpret_cast<const float4*>(yp + (int64_t)r * d) + c4);
        const float4 p1 = __ldcg(reinterpret_cast<const float4*>(yp + yps + (int64_t)r * d) + c4);
        const float4 p2 = __ldcg(reinterpret_cast<const float4*>(yp + 2 * yps + (int64_t)r * d) + c4);
        const float4 p3 = __ldcg(reinterpret_cast<const float4*>(yp + 3 * yps + (int64_t)r * d) + c4);
        const float4 s = make_float4(__fadd_rn(__fadd_rn(p0.x, p1.x), __fadd_rn(p2.x, p3.x)), __fadd_rn(__fadd_rn(p0.y, p1.y), __fadd_rn(p2.y, p3.y)),
                                     __fadd_rn(__fadd_rn(p0.z, p1.z), __fadd_rn(p2.z, p3.z)), __fadd_rn(__fadd_rn(p0.w, p1.w), __fadd_rn(p2.w, p3.w)));
        return make_float4(__fadd_rn(v.x, s.x), __fadd_rn(v.y, s.y), __fadd_rn(v.z, s.z), __fadd_rn(v.w, s.w));
    };
    auto load_x_folded = [&](int r, int c4) {     // P1 of layers > 0: fold and publish into the other buffer
        const float4 v = fold(r, c4);
        if (r % (int)gridDim.x == (int)blockIdx.x) *reinterpret_cast<float4*>(xo + (int64_t)r * d + c4 * 4) = v;
        return v;
    };
    auto load_x = [&](int r, int c4) { return __ldcg(reinterpret_cast<const float4*>(xc + (int64_t)r * d) + c4); };

    AFrag pf;
    for (int step = 0; step < a.n_steps; ++step) {
        const int p = a.pos0 + step;
        const bool want_logits = p >= a.logits_from;
        for (int l = 0; l < L; ++l) {
            const Dec3Layer& W = a.layers[l];
            KVT* kcl = reinterpret_cast<KVT*>(a.kc) + (size_t)l * a.Rmax * t_max * d;
            KVT* vcl = reinterpret_cast<KVT*>(a.vc) + (size_t)l * a.Rmax * t_max * d;
            // ================= P1: q | k | v = LN(x) Wqkv + b   (mod.rs:429-431)
            if (l == 0) {
                load_afrag(reinterpret_cast<const __half*>(W.Wqkv), W.bqkv, d, d, 3 * d / 16, blockIdx.x % (3 * d / 16), pf);
                if (tid < RP) tok_s[tid] = tid < R ? (a.use_cur_tok ? __ldcg(a.cur_tok + tid) : __ldcg(a.tokens + (int64_t)tid * t_max + p)) : 0;
            }
            {
                const float* pe = a.pos_emb + (int64_t)p * d;
                auto load_emb = [&](int r, int c4) {   // x = tok_emb[token] + pos_emb[p]  (mod.rs:141-146)
                    const float4 e4 = __ldg(reinterpret_cast<const float4*>(a.tok_emb + (int64_t)tok_s[r] * d) + c4);
                    const float4 p4 = __ldg(reinterpret_cast<const float4*>(pe) + c4);
                    const float4 v = make_float4(__fadd_rn(e4.x, p4.x), __fadd_rn(e4.y, p4.y), __fadd_rn(e4.z, p4.z), __fadd_rn(e4.w, p4.w));
                    if (r % (int)gridDim.x == (int)blockIdx.x) *reinterpret_cast<float4*>(xc + (int64_t)r * d + c4 * 4) = v;
                    return v;
                };
                auto stage = [&](int) {
                    if (l == 0) stage_ln_frag<NT8>(load_emb, R, d, W.ln1_g, W.ln1_b, W.ln1_eps, a.eps_outside, xhi, xlo);
                    else stage_ln_frag<NT8>(load_x_folded, R, d, W.ln1_g, W.ln1_b, W.ln1_eps, a.eps_outside, xhi, xlo);
                };
                gemm_phase<NT8>(reinterpret_cast<const __half*>(W.Wqkv), W.bqkv, 3 * d, d, d, R, pf, stage, no_pre,
                                [&](int n, int r, int, float v, float) {
                                    if (n < 2 * d) v = __fmul_rn(v, scale);
                                    if (n < d) a.q[(int64_t)r * d + n] = v;
                                    else if (n < 2 * d) kcl[((int64_t)r * t_max + p) * d + (n - d)] = (KVT)v;
                                    else vcl[((int64_t)r * t_max + p) * d + (n - 2 * d)] = (KVT)v;
                                }, xhi, xlo, red);
                if (l > 0) { float* tmp = xc; xc = xo; xo = tmp; }   // the folded rows were published into the other buffer
            }
            if ((int)blockIdx.x < d / 16) load_afrag(reinterpret_cast<const __half*>(W.Wo), W.bo, d, d, d / 16, blockIdx.x, pf);
            WB_TRACE();
            grid_sync(a.bar, gen);
            WB_TRACE();
            // ================= P2: self attention over positions 0..p of the row's ancestry (mask == causal)
            for (int u = blockIdx.x; u < R * H; u += gridDim.x) {
                const int r = u / H, h = u % H;
                if (tid < 16) *reinterpret_cast<float4*>(qs + tid * 4) = __ldcg(reinterpret_cast<const float4*>(a.q + (int64_t)r * d + h * 64) + tid);
                __syncthreads();
                const int* anc = a.anc ? a.anc + (int64_t)r * t_max : nullptr;
                const KVT* kb = kcl + h * 64;
                const KVT* vb = vcl + h * 64;
                auto kp = [&](int j) { return kb + ((int64_t)((anc && j < p) ? __ldcg(anc + j) : r) * t_max + j) * d; };
                auto vp = [&](int j) { return vb + ((int64_t)((anc && j < p) ? __ldcg(anc + j) : r) * t_max + j) * d; };
                attn_cta(qs, p + 1, kp, vp, wm, wl, wo, ao, ML);
                if (tid < 64) a.att[(int64_t)r * d + h * 64 + tid] = __fdiv_rn(ao[tid], ML[1]);
                __syncthreads();
            }
            WB_TRACE();
            grid_sync(a.bar, gen);
            WB_TRACE();
            // ================= P3: x += att Wo + bo   (mod.rs:435, :346)
            gemm_phase<NT8>(reinterpret_cast<const __half*>(W.Wo), W.bo, d, d, d, R, pf,
                            [&](int) { stage_copy_frag<NT8>(a.att, d, 0, d, R, xhi, xlo); },
                            [&](int n, int r) { return __ldcg(xc + (int64_t)r * d + n); },
                            [&](int n, int r, int, float v, float xold) { xc[(int64_t)r * d + n] = __fadd_rn(xold, v); }, xhi, xlo, red);
            if ((int)blockIdx.x < d / 16) load_afrag(reinterpret_cast<const __half*>(W.Wcq), W.bcq, d, d, d / 16, blockIdx.x, pf);
            WB_TRACE();
            grid_sync(a.bar, gen);
            WB_TRACE();
            // ================= P4: cross query = LN(x) Wq + b   (mod.rs:483)
            gemm_phase<NT8>(reinterpret_cast<const __half*>(W.Wcq), W.bcq, d, d, d, R, pf,
                            [&](int) { stage_ln_frag<NT8>(load_x, R, d, W.ln2_g, W.ln2_b, W.ln2_eps, a.eps_outside, xhi, xlo); }, no_pre,
                            [&](int n, int r, int, float v, float) { a.q[(int64_t)r * d + n] = __fmul_rn(v, scale); }, xhi, xlo, red);
            if ((int)blockIdx.x < d / 16) load_afrag(reinterpret_cast<const __half*>(W.Wco), W.bco, d, d, d / 16, blockIdx.x, pf);
            WB_TRACE();
            grid_sync(a.bar, gen);
            WB_TRACE();
            // ================= P5: cross attention, split over the window's encoder positions
            {
                const KVT* ckvl = reinterpret_cast<const KVT*>(a.ckv) + (size_t)l * a.Mcap * 2 * d;
                for (int u = blockIdx.x; u < R * H * S; u += gridDim.x) {
                    const int sp = u % S, h = (u / S) % H, r = u / (S * H);
                    if (tid < 16) *reinterpret_cast<float4*>(qs + tid * 4) = __ldcg(reinterpret_cast<const float4*>(a.q + (int64_t)r * d + h * 64) + tid);
                    __syncthreads();
                    const int w = __ldcg(a.row_window + r);
                    const int T = a.win_T[w];
                    const int per = (T + S - 1) / S;
                    const int kb0 = sp * per;
                    const int nk = max(0, min(T, kb0 + per) - kb0);
                    const KVT* kbase = ckvl + (a.win_row_off[w] + kb0) * (int64_t)(2 * d) + h * 64;
                    const int64_t ld = 2 * (int64_t)d;
                    auto kp = [&](int j) { return kbase + j * ld; };
                    auto vp = [&](int j) { return kbase + j * ld + d; };
                    attn_cta(qs, nk, kp, vp, wm, wl, wo, ao, ML);
                    const int64_t o = ((int64_t)r * H + h) * S + sp;
                    if (tid < 64) a.part_o[o * 64 + tid] = ao[tid];
                    if (tid == 0) { a.part_m[o] = nk > 0 ? ML[0] : -INFINITY; a.part_l[o] = ML[1]; }
                    __syncthreads();
                }
            }
            WB_TRACE();
            grid_sync(a.bar, gen);
            WB_TRACE();
            // ================= P6: x += merge(cross partials) Wo + bo   (mod.rs:489, :347)
            gemm_phase<NT8>(reinterpret_cast<const __half*>(W.Wco), W.bco, d, d, d, R, pf,
                            [&](int) {
                                float* wn = red;   // [RP][H][S] normalised split weights
                                for (int i = tid; i < RP * H; i += NT) {
                                    const int r = i / H, h = i % H;
                                    if (r < R) {
                                        const int64_t o = ((int64_t)r * H + h) * S;
                                        float pm[16], pl[16];
#pragma unroll
                                        for (int s = 0; s < 16; ++s) {
                                            pm[s] = s < S ? __ldcg(a.part_m + o + s) : -INFINITY;
                                            pl[s] = s < S ? __ldcg(a.part_l + o + s) : 0.0f;
                                        }
                                        float M = -INFINITY;
#pragma unroll
                                        for (int s = 0; s < 16; ++s) M = fmaxf(M, pm[s]);
                                        float den = 0.0f;
#pragma unroll
                                        for (int s = 0; s < 16; ++s) {
                                            pm[s] = pm[s] > -INFINITY ? expf(pm[s] - M) : 0.0f;
                                            den += pm[s] * pl[s];
                                        }
#pragma unroll
                                        for (int s = 0; s < 16; ++s)
                                            if (s < S) wn[i * S + s] = __fdiv_rn(pm[s], den);
                                    } else {
                                        for (int s = 0; s < S; ++s) wn[i * S + s] = 0.0f;
                                    }
                                }
                                __syncthreads();
                                for (int i = tid; i < RP * d / 4; i += NT) {   // 4 consecutive dims of one (row, head)
                                    const int r = (i * 4) / d, c = (i * 4) % d, rc = min(r, R - 1);
                                    const int h = c / 64;
                                    const float4* po = reinterpret_cast<const float4*>(a.part_o + (((int64_t)rc * H + h) * S) * 64 + (c & 63));
                                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
                                    for (int s = 0; s < S; ++s) {
                                        const float4 v = __ldcg(po + s * 16);
                                        const float wgt = wn[(r * H + h) * S + s];
                                        acc.x = fmaf(wgt, v.x, acc.x); acc.y = fmaf(wgt, v.y, acc.y);
                                        acc.z = fmaf(wgt, v.z, acc.z); acc.w = fmaf(wgt, v.w, acc.w);
                                    }
                                    store_frag(xhi, xlo, d / 32, r, c, acc);
                                }
                            },
                            [&](int n, int r) { return __ldcg(xc + (int64_t)r * d + n); },
                            [&](int n, int r, int, float v, float xold) { xc[(int64_t)r * d + n] = __fadd_rn(xold, v); }, xhi, xlo, red);
            load_afrag(reinterpret_cast<const __half*>(W.W1), W.b1, d, d, 4 * d / 16, blockIdx.x % (4 * d / 16), pf);
            WB_TRACE();
            grid_sync(a.bar, gen);
            WB_TRACE();
            // ================= P7: hid = gelu(LN(x) W1 + b1)   (mod.rs:377-378)
            gemm_phase<NT8>(reinterpret_cast<const __half*>(W.W1), W.b1, 4 * d, d, d, R, pf,
                            [&](int) { stage_ln_frag<NT8>(load_x, R, d, W.ln3_g, W.ln3_b, W.ln3_eps, a.eps_outside, xhi, xlo); }, no_pre,
                            [&](int n, int r, int, float v, float) { a.hid[(int64_t)r * 4 * d + n] = gelu_erf(v); }, xhi, xlo, red);
            load_afrag(reinterpret_cast<const __half*>(W.W2), W.b2, 4 * d, d, d / 16, blockIdx.x % (4 * (d / 16)), pf);
            WB_TRACE();
            grid_sync(a.bar, gen);
            WB_TRACE();
            // ================= P8: MLP2 partials: ypart[slab] = hid[:, slab] W2[:, slab]^T (+ b2 on slab 0); folded into x by the next LN
            gemm_phase<NT8>(reinterpret_cast<const __half*>(W.W2), W.b2, d, 4 * d, d, R, pf,
                            [&](int slab) { stage_copy_frag<NT8>(a.hid, 4 * d, slab * d, d, R, xhi, xlo); }, no_pre,
                            [&](int n, int r, int slab, float v, float) { a.ypart[slab * yps + (int64_t)r * d + n] = v; }, xhi, xlo, red);
            if (l + 1 < L) {
                const Dec3Layer& Wn = a.layers[l + 1];
                load_afrag(reinterpret_cast<const __half*>(Wn.Wqkv), Wn.bqkv, d, d, 3 * d / 16, blockIdx.x % (3 * d / 16), pf);
            }
            WB_TRACE();
            grid_sync(a.bar, gen);
            WB_TRACE();
        }
        if (want_logits) {
            // ================= logits = LN(x) tok_emb^T (mod.rs:155-156) -> lgbuf[R][V] (L2 resident)
            float* lgbuf = a.lgbuf;
            {
                __syncthreads();
                stage_ln_frag<NT8>(fold, R, d, a.lnf_g, a.lnf_b, a.lnf_eps, a.eps_outside, xhi, xlo);   // nobody reads x afterwards: no publish
                __syncthreads();
                const __half* E = reinterpret_cast<const __half*>(a.E);
                const int g = lane >> 2, t = lane & 3;
                const int n_tiles = (V + 15) / 16, ngrp = d / (32 * GC), nchunks = d / 32;
                const int my_tiles = gw < n_tiles ? (n_tiles - gw + n_gw - 1) / n_gw : 0;
                const int total = my_tiles * ngrp;
                uint4 A0[GC][2], A1[GC][2];
                auto load_grp = [&](int it, uint4 (&A)[GC][2]) {
                    const int tile = gw + (it / ngrp) * n_gw, grp = it % ngrp;
                    const int ra = min(tile * 16 + g, V - 1), rb = min(tile * 16 + g + 8, V - 1);
                    const uint4* pa = reinterpret_cast<const uint4*>(E + (int64_t)ra * d + grp * (32 * GC) + t * 8);
                    const uint4* pb = reinterpret_cast<const uint4*>(E + (int64_t)rb * d + grp * (32 * GC) + t * 8);
#pragma unroll
                    for (int c = 0; c < GC; ++c) { A[c][0] = __ldg(pa + c * 4); A[c][1] = __ldg(pb + c * 4); }
                };
                if (total > 0) load_grp(0, A0);
                float ah[NT8][4], al[NT8][4];
                for (int it = 0; it < total; ++it) {
                    const int grp = it % ngrp;
                    if (grp == 0) {
#pragma unroll
                        for (int j = 0; j < NT8; ++j)
#pragma unroll
                            for (int c = 0; c < 4; ++c) { ah[j][c] = 0.0f; al[j][c] = 0.0f; }
                    }
                    if (it + 1 < total) load_grp(it + 1, A1);
#pragma unroll
                    for (int c = 0; c < GC; ++c) {
                        const int chunk = grp * GC + c;
#pragma unroll
                        for (int j = 0; j < NT8; ++j) {
                            const uint4 bh = xhi[(j * nchunks + chunk) * 32 + lane];
                            const uint4 bl = xlo[(j * nchunks + chunk) * 32 + lane];
                            mma16816(ah[j], A0[c][0].x, A0[c][1].x, A0[c][0].y, A0[c][1].y, bh.x, bh.y);
                            mma16816(ah[j], A0[c][0].z, A0[c][1].z, A0[c][0].w, A0[c][1].w, bh.z, bh.w);
                            mma16816(al[j], A0[c][0].x, A0[c][1].x, A0[c][0].y, A0[c][1].y, bl.x, bl.y);
                            mma16816(al[j], A0[c][0].z, A0[c][1].z, A0[c][0].w, A0[c][1].w, bl.z, bl.w);
                        }
                    }
                    if (grp == ngrp - 1) {
                        const int n0 = (gw + (it / ngrp) * n_gw) * 16;
#pragma unroll
                        for (int j = 0; j < NT8; ++j) {
                            const int r0 = j * 8 + 2 * t;
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                const int n = n0 + g + (c >> 1) * 8, r = r0 + (c & 1);
                                if (n < V && r < R) lgbuf[(int64_t)r * V + n] = fmaf(al[j][c], 1.0f / 2048.0f, ah[j][c]);
                            }
                        }
                    }
#pragma unroll
                    for (int c = 0; c < GC; ++c) { A0[c][0] = A1[c][0]; A0[c][1] = A1[c][1]; }
                }
            }
            WB_TRACE();
            grid_sync(a.bar, gen);
            WB_TRACE();
            // ================= per (row, slice): special-token mask (transcribe.rs:271-275), max, sum-exp, top candidates
            const int NSL = a.lg_slices;
            {
                const bool use_mask = a.is_special != nullptr && (a.mask_mode == 1 || (a.mask_mode == 2 && p + 1 <= 5));
                const int per = (V + NSL - 1) / NSL;
                for (int u = blockIdx.x; u < R * NSL; u += gridDim.x) {
                    const int r = u / NSL, sl = u % NSL;
                    const int n_begin = sl * per, n_end = min(V, n_begin + per);
                    const float* row = lgbuf + (int64_t)r * V;
                    float m_run = -INFINITY, s_run = 0.0f;
                    Cand<KC> cand;
                    cand.init();
                    for (int n0 = n_begin + tid; n0 < n_end; n0 += NT * 8) {
                        float raw[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) raw[i] = (n0 + i * NT < n_end) ? __ldcg(row + n0 + i * NT) : 0.0f;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int n = n0 + i * NT;
                            if (n < n_end) {
                                const float v = (use_mask && a.is_special[n]) ? __fadd_rn(raw[i], -INFINITY) : raw[i];
                                if (v > -INFINITY) {
                                    if (v > m_run) { s_run = s_run * expf(m_run - v) + 1.0f; m_run = v; }
                                    else s_run += expf(v - m_run);
                                }
                                cand.push(v, n);
                            }
                        }
                    }
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        const float m2 = __shfl_xor_sync(0xffffffffu, m_run, off);
                        const float s2 = __shfl_xor_sync(0xffffffffu, s_run, off);
                        float cv[KC];
                        int ci[KC];
#pragma unroll
                        for (int k = 0; k < KC; ++k) { cv[k] = __shfl_xor_sync(0xffffffffu, cand.v[k], off); ci[k] = __shfl_xor_sync(0xffffffffu, cand.i[k], off); }
                        const float mn = fmaxf(m_run, m2);
                        const float e1 = m_run > -INFINITY ? expf(m_run - mn) : 0.0f;
                        const float e2 = m2 > -INFINITY ? expf(m2 - mn) : 0.0f;
                        s_run = s_run * e1 + s2 * e2;
                        m_run = mn;
#pragma unroll
                        for (int k = 0; k < KC; ++k) cand.push(cv[k], ci[k]);
                    }
                    float* rec = red + warp * (2 + 2 * KC);
                    if (lane == 0) {
                        rec[0] = m_run;
                        rec[1] = s_run;
#pragma unroll
                        for (int k = 0; k < KC; ++k) { rec[2 + k] = cand.v[k]; rec[2 + KC + k] = __int_as_float(cand.i[k]); }
                    }
                    __syncthreads();
                    if (tid == 0) {
                        float M = -INFINITY;
                        for (int w = 0; w < NW; ++w) M = fmaxf(M, red[w * (2 + 2 * KC)]);
                        float Ssum = 0.0f;
                        Cand<KC> best;
                        best.init();
                        for (int w = 0; w < NW; ++w) {
                            const float* rc = red + w * (2 + 2 * KC);
                            if (rc[0] > -INFINITY) Ssum += rc[1] * expf(rc[0] - M);
#pragma unroll
                            for (int k = 0; k < KC; ++k) best.push(rc[2 + k], __float_as_int(rc[2 + KC + k]));
                        }
                        const int64_t o = (int64_t)sl * R + r;
                        a.lg_m[o] = M;
                        a.lg_s[o] = Ssum;
#pragma unroll
                        for (int k = 0; k < KC; ++k) { a.lg_v[o * KC + k] = best.v[k]; a.lg_i[o * KC + k] = best.i[k]; }
                    }
                    __syncthreads();
                }
            }
            WB_TRACE();
            grid_sync(a.bar, gen);
            WB_TRACE();
            // ================= finish: log_softmax of the candidates, k best (ties -> lower id), greedy bookkeeping
            for (int r = blockIdx.x; r < R; r += gridDim.x) {
                float* s_f = wm;   // [NW] scratch
                int* s_i = reinterpret_cast<int*>(wl);
                const int NP = NSL;
                float mx = -INFINITY;
                for (int c = tid; c < NP; c += NT) mx = fmaxf(mx, __ldcg(a.lg_m + (int64_t)c * R + r));
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                if (lane == 0) s_f[warp] = mx;
                __syncthreads();
                mx = s_f[0];
#pragma unroll
                for (int w = 1; w < NW; ++w) mx = fmaxf(mx, s_f[w]);
                __syncthreads();
                float se = 0.0f;
                for (int c = tid; c < NP; c += NT) {
                    const float m = __ldcg(a.lg_m + (int64_t)c * R + r);
                    if (m > -INFINITY) se += __ldcg(a.lg_s + (int64_t)c * R + r) * expf(m - mx);
                }
                se = warp_sum(se);
                if (lane == 0) s_f[warp] = se;
                __syncthreads();
                se = 0.0f;
#pragma unroll
                for (int w = 0; w < NW; ++w) se += s_f[w];
                const float lse = logf(se);
                __syncthreads();
                float prev_v = INFINITY;
                int prev_i = -1;
                for (int kk = 0; kk < a.k; ++kk) {
                    float bv = -INFINITY;
                    int bi = INT_MAX;
                    for (int c = tid; c < NP * KC; c += NT) {
                        const int part = c / KC, k = c % KC;
                        const int idx = __ldcg(a.lg_i + ((int64_t)part * R + r) * KC + k);
                        if (idx == INT_MAX) continue;
                        const float v = __fsub_rn(__fsub_rn(__ldcg(a.lg_v + ((int64_t)part * R + r) * KC + k), mx), lse);
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
                    bv = s_f[0];
                    bi = s_i[0];
#pragma unroll
                    for (int w = 1; w < NW; ++w)
                        if (s_f[w] > bv || (s_f[w] == bv && s_i[w] < bi)) { bv = s_f[w]; bi = s_i[w]; }
                    __syncthreads();
                    if (tid == 0) {
                        a.topk_id[(int64_t)r * a.k + kk] = bi == INT_MAX ? -1 : bi;
                        a.topk_lp[(int64_t)r * a.k + kk] = bv;
                        if (kk == 0 && a.greedy && !__ldcg(a.finished + r)) {   // beam.rs:9-37 with beam_size 1
                            a.tokens[(int64_t)r * t_max + p + 1] = bi;
                            a.lengths[r] = p + 2;
                            if (bi == a.eot) a.finished[r] = 1;
                        }
                    }
                    prev_v = bv;
                    prev_i = bi;
                }
            }
            WB_TRACE();
            grid_sync(a.bar, gen);
            WB_TRACE();
            if (a.greedy) {   // stop as soon as every search has produced EOT (beam.rs:22-27)
                int live = 0;
                for (int r = 0; r < R; ++r) live += __ldcg(a.finished + r) ? 0 : 1;
                if (live == 0) {
                    if (blockIdx.x == 0 && tid == 0) { *a.pos = p + 1; *a.n_unfinished = 0; *a.steps_done = step + 1; }
                    return;
                }
            }
        }
    }
    if (blockIdx.x == 0 && tid == 0) {
        *a.pos = a.pos0 + a.n_steps;
        int live = 0;
        for (int r = 0; r < R; ++r) live += (a.greedy && __ldcg(a.finished + r)) ? 0 : 1;
        *a.n_unfinished = live;
        *a.steps_done = a.n_steps;
    }
}

size_t dec5_smem_bytes(int d, int NT8) {
    return (size_t)2 * NT8 * (d / 32) * 32 * 16 + sizeof(float) * ((size_t)NW * NT8 * 8 * RED_LD + 64 + 2 * NW + NW * 64 + 64 + 2 + NT8 * 8 + 8);
}

template <int NT8, int KC, typename KVT>
bool launch5_t(const Dec3Args& a, int n_ctas, cudaStream_t st) {
    const size_t smem = dec5_smem_bytes(a.d, NT8);
    auto k = dec5_kernel<NT8, KC, KVT>;
    static size_t configured = 0;   // per instantiation
    if (configured != smem) {
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        int per_sm = 0;
        WB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, NT, smem));
        if (per_sm < 1) return false;
        configured = smem;
    }
    void* args[] = {(void*)&a};
    WB_CUDA(cudaLaunchCooperativeKernel((void*)k, dim3(n_ctas), dim3(NT), args, smem, st));
    WB_LAUNCH_CHECK();
    return true;
}

}  // namespace

// Returns false when this configuration is not covered (caller falls back to decoder3.cu).
bool launch_dec5(const Dec3Args& a, int n_ctas, bool w_half, cudaStream_t st) {
    if (!w_half || a.R < 1 || a.R > 32 || a.d % 256 != 0 || a.d > 1280 || a.H * 64 != a.d) return false;
    if (a.lgbuf == nullptr || a.ypart == nullptr || a.lg_slices < 1) return false;
    if ((size_t)a.R * a.H * a.n_splits > (size_t)NW * ((a.R + 7) / 8) * 8 * RED_LD) return false;   // cross-merge weights live in the reduction buffer
    const int nt8 = (a.R + 7) / 8;
    const bool wide = a.k > 1;
#define WB_D5(NT8_)                                                                                          \
    do {                                                                                                     \
        if (a.kv_half) return wide ? launch5_t<NT8_, 8, __half>(a, n_ctas, st) : launch5_t<NT8_, 2, __half>(a, n_ctas, st); \
        return wide ? launch5_t<NT8_, 8, float>(a, n_ctas, st) : launch5_t<NT8_, 2, float>(a, n_ctas, st);    \
    } while (0)
    if (nt8 == 1) WB_D5(1);
    if (nt8 == 2) WB_D5(2);
    if (nt8 == 3) WB_D5(3);
    WB_D5(4);
#undef WB_D5
}

}  // namespace wb

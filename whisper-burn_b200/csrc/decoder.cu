// KV-cached single-position decoder step for sm_100a  (reference: TextDecoder::forward
// src/model/mod.rs:131-157, ResidualDecoderAttentionBlock::forward :345-350, qkv_attention :493-533,
// and the per-step closure beamsearch_next src/transcribe.rs:253-307).
//
// The reference recomputes the whole prefix for every beam on every step.  Here one
// step processes ONE new position for R rows (rows = live beams of all windows in flight):
//
//   dec_embed            x[r] = tok_emb[token[r]] + pos_emb[p]                       (mod.rs:141-146)
//   per layer:
//     gemv<LN,QKV>       q | k | v = LN(x) W + b; q,k scaled dh^-0.25; k,v appended to the self cache
//     self_attn          softmax(q K^T) V over positions 0..p of the row's ancestry   (mask == causal)
//     gemv<MERGE,RESID>  x += attn W_o + b_o
//     gemv<LN,Q>         cross query
//     cross_attn         split-KV over the window's T encoder positions (K/V projected once per window)
//     gemv<MERGE,RESID>  x += cross W_o + b_o
//     gemv<LN,GELU>      h = gelu(LN(x) W_1 + b_1)
//     gemv<COPY,RESID>   x += h W_2 + b_2
//   gemv<LN,LOGITS>      logits = LN(x) tok_emb^T   (last position only, mod.rs:155-156)
//   logsoftmax_topk      + special-token mask, log_softmax, k best ids        (transcribe.rs:271-276)
//
// The GEMVs are HBM/L2-bandwidth bound: weights are read once per step for all rows (fp16 storage
// when the checkpoint is fp16-exact, else fp32), 128-bit loads, fp32 accumulate, warp per output
// feature.  Algorithmic bytes per step = sum of weight bytes + cross K/V + self K/V.
#include <cfloat>
#include <climits>

#include "wb_internal.h"
#include "decoder.h"

namespace wb {

namespace {

__device__ __forceinline__ float gelu_erf(float x) {
    const float t = __fadd_rn(erff(__fdiv_rn(x, 1.41421356237309504880f)), 1.0f);
    return __fdiv_rn(__fmul_rn(x, t), 2.0f);
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

// ---------------------------------------------------------------------------------------------
__global__ void dec_embed_kernel(const int* __restrict__ tok, const float* __restrict__ emb,
                                 const float* __restrict__ pos_emb, const int* __restrict__ pos_ptr,
                                 float* __restrict__ x, int d) {
    const int r = blockIdx.x;
    const int p = *pos_ptr;
    const float* e = emb + (int64_t)tok[r] * d;
    const float* pe = pos_emb + (int64_t)p * d;
    for (int c = threadIdx.x; c < d; c += blockDim.x) x[(int64_t)r * d + c] = __fadd_rn(e[c], pe[c]);
}

// ---------------------------------------------------------------------------------------------
constexpr int GV_WARPS = 8;

template <typename WT, int RC, int RPW>
__global__ void __launch_bounds__(GV_WARPS * 32)
dec_gemv_kernel(const GemvArgs a) {
    extern __shared__ __align__(16) float xs[];   // [RC][K]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = a.K;
    const int p = a.pos ? *a.pos : 0;
    const WT* W = reinterpret_cast<const WT*>(a.W);
    const int n_base = (blockIdx.x * GV_WARPS + warp) * RPW;

    for (int r0 = 0; r0 < a.R; r0 += RC) {
        __syncthreads();
        // ---- input stage: RC rows into shared memory
        for (int rr = warp; rr < RC; rr += GV_WARPS) {
            const int r = r0 + rr;
            float* xr = xs + rr * K;
            if (r >= a.R) {
                for (int c = lane; c < K; c += 32) xr[c] = 0.0f;
            } else if (a.in_mode == IN_LN) {
                const float* src = a.in + (int64_t)r * K;
                float s = 0.0f;
                for (int c = lane; c < K; c += 32) s += src[c];
                s = warp_sum(s);
                const float mean = __fdiv_rn(s, (float)K);
                float q = 0.0f;
                for (int c = lane; c < K; c += 32) {
                    const float dv = __fsub_rn(src[c], mean);
                    q = __fadd_rn(q, __fmul_rn(dv, dv));
                }
                q = warp_sum(q);
                const float var = __fdiv_rn(q, (float)K);
                const float den = a.eps_outside ? __fadd_rn(__fsqrt_rn(var), a.ln_eps) : __fsqrt_rn(__fadd_rn(var, a.ln_eps));
                for (int c = lane; c < K; c += 32)
                    xr[c] = __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(src[c], mean), den), a.ln_g[c]), a.ln_b[c]);
            } else if (a.in_mode == IN_COPY) {
                const float* src = a.in + (int64_t)r * K;
                for (int c = lane; c < K; c += 32) xr[c] = src[c];
            } else {   // IN_ATTN_MERGE: combine split-KV partials (o unnormalised, m, l) per head
                const int H = K / 64, S = a.n_splits;
                for (int h = 0; h < H; ++h) {
                    const float* pm = a.part_m + ((int64_t)r * H + h) * S;
                    const float* pl = a.part_l + ((int64_t)r * H + h) * S;
                    const float* po = a.part_o + ((int64_t)r * H + h) * S * 64;
                    float M = -INFINITY;
                    for (int s = 0; s < S; ++s) M = fmaxf(M, pm[s]);
                    float den = 0.0f, v0 = 0.0f, v1 = 0.0f;
                    for (int s = 0; s < S; ++s) {
                        const float wgt = expf(pm[s] - M);
                        den += wgt * pl[s];
                        v0 += wgt * po[s * 64 + lane];
                        v1 += wgt * po[s * 64 + 32 + lane];
                    }
                    xr[h * 64 + lane] = __fdiv_rn(v0, den);
                    xr[h * 64 + 32 + lane] = __fdiv_rn(v1, den);
                }
            }
        }
        __syncthreads();
        // ---- RPW output features per warp, all RC rows
        float acc[RPW][RC];
#pragma unroll
        for (int q = 0; q < RPW; ++q)
#pragma unroll
            for (int rr = 0; rr < RC; ++rr) acc[q][rr] = 0.0f;
        if (n_base < a.N) {
            for (int v = lane; v < K / 8; v += 32) {
                float w[RPW][8];
#pragma unroll
                for (int q = 0; q < RPW; ++q) {
                    const int n = min(n_base + q, a.N - 1);
                    load8(W + (int64_t)n * K + v * 8, w[q]);
                }
#pragma unroll
                for (int rr = 0; rr < RC; ++rr) {
                    const float4 x0 = *reinterpret_cast<const float4*>(xs + rr * K + v * 8);
                    const float4 x1 = *reinterpret_cast<const float4*>(xs + rr * K + v * 8 + 4);
                    const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                    for (int q = 0; q < RPW; ++q)
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[q][rr] = fmaf(w[q][i], xv[i], acc[q][rr]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < RPW; ++q)
#pragma unroll
            for (int rr = 0; rr < RC; ++rr) acc[q][rr] = warp_sum(acc[q][rr]);
        // ---- epilogue: lane (q*RC + rr) owns output (q, rr)
        float mine = 0.0f;
#pragma unroll
        for (int q = 0; q < RPW; ++q)
#pragma unroll
            for (int rr = 0; rr < RC; ++rr)
                if (lane == q * RC + rr) mine = acc[q][rr];
        if (lane < RPW * RC) {
            const int q = lane / RC, rr = lane % RC;
            const int n = n_base + q, r = r0 + rr;
            if (n < a.N && r < a.R) {
                float v = a.bias ? __fadd_rn(mine, a.bias[n]) : mine;
                if (a.act == ACT_GELU) v = gelu_erf(v);
                int si = 0;
                if (a.n_seg > 1 && n >= a.seg[1].begin) si = 1;
                if (a.n_seg > 2 && n >= a.seg[2].begin) si = 2;
                const GemvSeg sg = a.seg[si];
                if (sg.scale != 1.0f) v = __fmul_rn(v, sg.scale);
                float* dst = sg.out + (int64_t)r * sg.row_stride + (int64_t)p * sg.pos_stride + (n - sg.begin);
                if (a.residual) v = __fadd_rn(*dst, v);
                *dst = v;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Attention of one new query against cached keys; writes split partials (o unnormalised, m, l).
// grid (n_head, R, n_splits), 128 threads.  Keys [k_begin, k_end) of this split:
//   self : row pointer K + (anc[j] * t_max + j) * d, j in 0..p          (anc == null -> the row itself)
//   cross: row pointer K + (row_off[w] + j) * ld,     j in 0..T_w
constexpr int DA_THREADS = 128;
constexpr int DA_MAX_KEYS = 512;   // keys per split

__global__ void __launch_bounds__(DA_THREADS)
dec_attn_kernel(const DecAttnArgs a) {
    __shared__ float sc[DA_MAX_KEYS];
    __shared__ __align__(16) float red[8][64];
    __shared__ float s_red[4];
    const int h = blockIdx.x, r = blockIdx.y, sp = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = gridDim.x, S = gridDim.z;
    const int p = *a.pos;

    int n_keys;
    const float* kbase;
    const float* vbase;
    int64_t ld;
    const int* anc = nullptr;
    if (a.is_cross) {
        const int w = a.row_window[r];
        n_keys = a.win_T[w];
        ld = a.kv_ld;
        kbase = a.K + a.win_row_off[w] * ld + h * 64;
        vbase = a.V + a.win_row_off[w] * ld + h * 64;
    } else {
        n_keys = p + 1;
        ld = a.kv_ld;   // == d
        kbase = a.K + h * 64;
        vbase = a.V + h * 64;
        anc = a.anc ? a.anc + (int64_t)r * a.t_max : nullptr;
    }
    const int per = (n_keys + S - 1) / S;
    const int k_begin = sp * per;
    const int k_end = min(n_keys, k_begin + per);
    const int nk = max(0, k_end - k_begin);
    const int64_t out_idx = ((int64_t)r * H + h) * S + sp;

    // q slice: lane%16 holds 4 dims
    const int l16 = lane & 15, half = lane >> 4;
    const float4 q4 = *reinterpret_cast<const float4*>(a.q + (int64_t)r * a.q_ld + h * 64 + l16 * 4);

    auto key_row = [&](int j) -> int64_t {
        if (a.is_cross) return (int64_t)j * ld;
        const int src = anc ? anc[j] : r;
        return ((int64_t)src * a.t_max + j) * ld;
    };

    // ---- scores: half-warp per key (16 lanes x float4 = one 256-byte head slice)
    float lmax = -INFINITY;
    for (int jb = warp * 2; jb < nk; jb += (DA_THREADS / 32) * 2) {   // warp-uniform trip count
        const int j = jb + half;
        const bool valid = j < nk;
        float4 k4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) k4 = __ldg(reinterpret_cast<const float4*>(kbase + key_row(k_begin + j) + l16 * 4));
        float s = q4.x * k4.x;
        s = fmaf(q4.y, k4.y, s);
        s = fmaf(q4.z, k4.z, s);
        s = fmaf(q4.w, k4.w, s);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (valid) {
            if (l16 == 0) sc[j] = s;
            lmax = fmaxf(lmax, s);
        }
    }
    lmax = warp_max(lmax);
    if (lane == 0) s_red[warp] = lmax;
    __syncthreads();
    const float M = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
    // ---- weights and weighted values: thread (jg = tid/16, c4 = tid%16)
    const int jg = tid >> 4, c4 = tid & 15;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float lsum = 0.0f;
    for (int j = jg; j < nk; j += DA_THREADS / 16) {
        const float e = expf(sc[j] - M);
        lsum += e;
        const float4 v4 = __ldg(reinterpret_cast<const float4*>(vbase + key_row(k_begin + j) + c4 * 4));
        acc.x = fmaf(e, v4.x, acc.x);
        acc.y = fmaf(e, v4.y, acc.y);
        acc.z = fmaf(e, v4.z, acc.z);
        acc.w = fmaf(e, v4.w, acc.w);
    }
    *reinterpret_cast<float4*>(&red[jg][c4 * 4]) = acc;
    // lsum: every c4 lane of a jg group has the same value; reduce across jg via smem
    __syncthreads();
    if (c4 == 0) sc[jg] = lsum;   // sc is free now (all reads done before the barrier above)
    __syncthreads();
    if (tid < 64) {
        float o = 0.0f;
#pragma unroll
        for (int g = 0; g < 8; ++g) o += red[g][tid];
        a.part_o[out_idx * 64 + tid] = o;
    }
    if (tid == 0) {
        float l = 0.0f;
#pragma unroll
        for (int g = 0; g < 8; ++g) l += sc[g];
        a.part_m[out_idx] = nk > 0 ? M : -INFINITY;
        a.part_l[out_idx] = l;
    }
}

// ---------------------------------------------------------------------------------------------
// Special-token mask + log_softmax + top-k of one logits row (transcribe.rs:271-276 + the
// selection that beam.rs:81-110 would make: larger log-prob first, on ties the lower id).
// grid R, 1024 threads.  In greedy mode (tokens != null) the winner becomes the row's next token.
constexpr int LS_THREADS = 1024;

struct KeyVal {
    float v;
    int i;
};
__device__ __forceinline__ bool kv_better(float v, int i, float bv, int bi) {
    return v > bv || (v == bv && i < bi);
}

__global__ void __launch_bounds__(LS_THREADS)
logsoftmax_topk_kernel(const LogSoftmaxArgs a) {
    __shared__ float s_f[32];
    __shared__ int s_i[32];
    __shared__ float s_bcast_f;
    __shared__ int s_bcast_i;
    const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int V = a.V;
    const int p = *a.pos;
    const bool use_mask = a.is_special != nullptr && (a.mask_mode == 1 || (a.mask_mode == 2 && p + 1 <= 5));
    const float* x = a.logits + (int64_t)r * V;

    auto val = [&](int i) -> float {
        float v = x[i];
        if (use_mask && a.is_special[i]) v = __fadd_rn(v, -INFINITY);
        return v;
    };
    // max
    float mx = -INFINITY;
    for (int i = tid; i < V; i += LS_THREADS) mx = fmaxf(mx, val(i));
    mx = warp_max(mx);
    if (lane == 0) s_f[warp] = mx;
    __syncthreads();
    if (warp == 0) {
        float t = s_f[lane];
        t = warp_max(t);
        if (lane == 0) s_bcast_f = t;
    }
    __syncthreads();
    mx = s_bcast_f;
    __syncthreads();
    // sum exp
    float se = 0.0f;
    for (int i = tid; i < V; i += LS_THREADS) se += expf(__fsub_rn(val(i), mx));
    se = warp_sum(se);
    if (lane == 0) s_f[warp] = se;
    __syncthreads();
    if (warp == 0) {
        float t = s_f[lane];
        t = warp_sum(t);
        if (lane == 0) s_bcast_f = logf(t);
    }
    __syncthreads();
    const float lse = s_bcast_f;
    __syncthreads();
    if (a.logprob_out) {   // full row of log-probs (stateless forward path / tests)
        float* lp = a.logprob_out + (int64_t)r * V;
        for (int i = tid; i < V; i += LS_THREADS) lp[i] = __fsub_rn(__fsub_rn(val(i), mx), lse);
    }
    // top-k by repeated arg-best with exclusion of everything at or before the previous winner
    float prev_v = INFINITY;
    int prev_i = -1;
    for (int kk = 0; kk < a.k; ++kk) {
        float bv = -INFINITY;
        int bi = INT_MAX;
        for (int i = tid; i < V; i += LS_THREADS) {
            const float v = __fsub_rn(__fsub_rn(val(i), mx), lse);
            const bool after_prev = v < prev_v || (v == prev_v && i > prev_i);
            if (after_prev && kv_better(v, i, bv, bi)) {
                bv = v;
                bi = i;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (kv_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { s_f[warp] = bv; s_i[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            bv = s_f[lane];
            bi = s_i[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (kv_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) { s_bcast_f = bv; s_bcast_i = bi; }
        }
        __syncthreads();
        prev_v = s_bcast_f;
        prev_i = s_bcast_i;
        __syncthreads();
        if (tid == 0) {
            a.topk_id[(int64_t)r * a.k + kk] = prev_i == INT_MAX ? -1 : prev_i;
            a.topk_lp[(int64_t)r * a.k + kk] = prev_v;
        }
    }
    // greedy bookkeeping: beam_search with beam_size 1 (beam.rs:9-37) == argmax until EOT / max_depth
    if (a.greedy && tid == 0) {
        const int best = a.topk_id[(int64_t)r * a.k];
        if (!a.finished[r]) {
            a.tokens[(int64_t)r * a.t_max + p + 1] = best;
            a.lengths[r] = p + 2;
            a.cur_tok[r] = best;
            if (best == a.eot) a.finished[r] = 1;
        }
    }
}

// pos += 1; n_unfinished for the host's early-exit poll
__global__ void dec_advance_kernel(int* pos, const int* finished, int R, int* n_unfinished) {
    if (threadIdx.x == 0) {
        int c = 0;
        for (int r = 0; r < R; ++r) c += finished ? (finished[r] ? 0 : 1) : 1;
        *n_unfinished = c;
        *pos = *pos + 1;
    }
}

// ancestry table for beam reordering: anc_new[r][0..p) = anc_old[parent[r]][0..p), anc_new[r][p] = r
__global__ void dec_reorder_kernel(const int* __restrict__ anc_old, int* __restrict__ anc_new,
                                   const int* __restrict__ parent, const int* __restrict__ pos_ptr, int t_max) {
    const int r = blockIdx.x;
    const int p = *pos_ptr;
    const int* src = anc_old + (int64_t)parent[r] * t_max;
    int* dst = anc_new + (int64_t)r * t_max;
    for (int j = threadIdx.x; j < p; j += blockDim.x) dst[j] = src[j];
    if (threadIdx.x == 0) dst[p] = r;
}

__global__ void dec_anc_identity_kernel(int* __restrict__ anc, int t_max) {
    for (int j = threadIdx.x; j < t_max; j += blockDim.x) anc[(int64_t)blockIdx.x * t_max + j] = blockIdx.x;
}

template <typename WT, int RC>
void launch_gemv_t(const GemvArgs& a, cudaStream_t st) {
    constexpr int RPW = (RC <= 2) ? 4 : (RC <= 4 ? 2 : 1);
    const int feats_per_cta = GV_WARPS * RPW;
    const size_t smem = (size_t)RC * a.K * sizeof(float);
    auto kern = dec_gemv_kernel<WT, RC, RPW>;
    if (smem > 48 * 1024) WB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(a.N + feats_per_cta - 1) / feats_per_cta, GV_WARPS * 32, smem, st>>>(a);
    WB_LAUNCH_CHECK();
}

}  // namespace

void launch_dec_embed(const int* tok, const float* emb, const float* pos_emb, const int* pos_ptr, float* x, int R,
                      int d, cudaStream_t st) {
    dec_embed_kernel<<<R, 128, 0, st>>>(tok, emb, pos_emb, pos_ptr, x, d);
    WB_LAUNCH_CHECK();
}

void launch_dec_gemv(const GemvArgs& a, bool w_half, cudaStream_t st) {
    WB_REQUIRE(a.K % 8 == 0, "gemv: K must be a multiple of 8");
    // rows per pass: as many as fit in ~96 KB of shared memory, at most 8
    int rc = 8;
    while (rc > 1 && ((size_t)rc * a.K * sizeof(float) > 96 * 1024 || rc / 2 >= a.R)) rc >>= 1;
#define WB_GV(RCV)                                                  \
    if (w_half) launch_gemv_t<__half, RCV>(a, st);                  \
    else launch_gemv_t<float, RCV>(a, st);
    switch (rc) {
        case 1: WB_GV(1); break;
        case 2: WB_GV(2); break;
        case 4: WB_GV(4); break;
        default: WB_GV(8); break;
    }
#undef WB_GV
}

void launch_dec_attn(const DecAttnArgs& a, int n_head, int R, int n_splits, cudaStream_t st) {
    dim3 grid(n_head, R, n_splits);
    dec_attn_kernel<<<grid, DA_THREADS, 0, st>>>(a);
    WB_LAUNCH_CHECK();
}

int dec_attn_max_keys_per_split() { return DA_MAX_KEYS; }

void launch_logsoftmax_topk(const LogSoftmaxArgs& a, int R, cudaStream_t st) {
    logsoftmax_topk_kernel<<<R, LS_THREADS, 0, st>>>(a);
    WB_LAUNCH_CHECK();
}

void launch_dec_advance(int* pos, const int* finished, int R, int* n_unfinished, cudaStream_t st) {
    dec_advance_kernel<<<1, 32, 0, st>>>(pos, finished, R, n_unfinished);
    WB_LAUNCH_CHECK();
}

void launch_dec_anc_identity(int* anc, int R, int t_max, cudaStream_t st) {
    dec_anc_identity_kernel<<<R, 128, 0, st>>>(anc, t_max);
    WB_LAUNCH_CHECK();
}

void launch_dec_reorder(const int* anc_old, int* anc_new, const int* parent, const int* pos_ptr, int R, int t_max,
                        cudaStream_t st) {
    dec_reorder_kernel<<<R, 128, 0, st>>>(anc_old, anc_new, parent, pos_ptr, t_max);
    WB_LAUNCH_CHECK();
}

}  // namespace wb

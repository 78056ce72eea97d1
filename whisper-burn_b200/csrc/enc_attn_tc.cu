// Encoder self-attention on the tensor cores (reference op: qkv_attention, src/model/mod.rs:493-533, non-causal, head dim 64).
//
// Flash-attention structure (one CTA = 64 queries of one (window, head), 4 warps x 16 query rows, key/value tiles of 64
// positions streamed through a double-buffered cp.async ring), but at fp32-class accuracy: the reference computes in f32 and the
// parity bar is 2e-5 of scale, so every operand is a PAIR of fp16 planes  x = hi + lo / 2048  (hi = fp16(x),
// lo = fp16((x - hi) * 2048): 22 mantissa bits, the decoder5.cu split) and every product is three mma.sync.m16n8k16 terms
//     Q.K^T = Qh.Kh + (Qh.Kl + Ql.Kh) / 2048          P.V = Ph.Vh + (Ph.Vl + Pl.Vh) / 2048
// (the dropped lo.lo term is 2^-22 relative) with fp32 accumulation in two accumulators (main / correction) and the softmax in
// fp32 exactly as burn's activation::softmax composes it (exp(x - max) / sum, online over the key tiles).
// q | k | v arrive as fp16 planes [rows][3d] written by the QKV GEMM epilogue (q and k already carry dh^-0.25, mod.rs:503-514);
// the output leaves as the fp16 planes [rows][d] the out-projection GEMM consumes.
#include <cuda_fp16.h>

#include "wb_internal.h"

namespace wb {

namespace {

constexpr int TQ = 64, TK = 64, HD = 64;
constexpr int AT_THREADS = 128;
constexpr int TILE_B = 64 * 128;                       // one [64][64] fp16 tile = 8 KB (rows of 128 bytes)
constexpr size_t AT_SMEM = 2 * TILE_B + 2 * 4 * TILE_B;   // Q hi/lo + 2 stages x (K hi, K lo, V hi, V lo) = 80 KB

__device__ __forceinline__ uint32_t s_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp16(uint32_t dst, const void* src, bool ok) {
    const int sz = ok ? 16 : 0;   // zero-fill out-of-range rows
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
// tile [64 rows][64 halves]: 16-byte chunk c of row r sits at r * 128 + ((c ^ (r & 7)) << 4)  (conflict-free ldmatrix)
__device__ __forceinline__ uint32_t tile_off(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr));
}
__device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(__half a, __half b) {
    const __half2 h = __halves2half2(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
// two fp32 values -> fp16 hi pair and fp16 (residual * 2048) pair
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
    const __half hx = __float2half_rn(x), hy = __float2half_rn(y);
    hi = pack2(hx, hy);
    lo = pack2(__float2half_rn((x - __half2float(hx)) * 2048.0f), __float2half_rn((y - __half2float(hy)) * 2048.0f));
}

__global__ void __launch_bounds__(AT_THREADS)
enc_attn_tc_kernel(const __half* __restrict__ qkv_hi, const __half* __restrict__ qkv_lo, __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                   const AttnWindow* __restrict__ wins, int d) {
    extern __shared__ __align__(128) uint8_t sm[];
    const AttnWindow win = wins[blockIdx.z];
    const int q0 = blockIdx.x * TQ;
    if (q0 >= win.T) return;
    const int h = blockIdx.y, T = win.T;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int64_t ld = 3 * (int64_t)d;
    const __half* bh = qkv_hi + win.row_off * ld + h * HD;
    const __half* bl = qkv_lo + win.row_off * ld + h * HD;
    const uint32_t sQ = s_addr(sm), sKV = sQ + 2 * TILE_B;

    // ---- Q tile (hi, lo) and the first K/V tiles
    auto load_tile = [&](uint32_t dst, const __half* src, int row0) {   // [64][64] halves from rows row0.. of a [.][3d] plane
        for (int i = tid; i < 64 * 8; i += AT_THREADS) {
            const int r = i >> 3, c = i & 7;
            const bool ok = row0 + r < T;
            cp16(dst + tile_off(r, c), src + (int64_t)(ok ? row0 + r : 0) * ld + c * 8, ok);
        }
    };
    auto load_kv = [&](int stage, int k0) {
        const uint32_t base = sKV + stage * 4 * TILE_B;
        load_tile(base, bh + d, k0);
        load_tile(base + TILE_B, bl + d, k0);
        load_tile(base + 2 * TILE_B, bh + 2 * d, k0);
        load_tile(base + 3 * TILE_B, bl + 2 * d, k0);
    };
    load_tile(sQ, bh, q0);
    load_tile(sQ + TILE_B, bl, q0);
    load_kv(0, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");

    float o_m[8][4], o_c[8][4], m_row[2] = {-INFINITY, -INFINITY}, l_row[2] = {0.0f, 0.0f};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { o_m[i][j] = 0.0f; o_c[i][j] = 0.0f; }
    uint32_t qh[4][4], ql[4][4];

    const int n_tiles = (T + TK - 1) / TK;
    for (int it = 0; it < n_tiles; ++it) {
        if (it + 1 < n_tiles) load_kv((it + 1) & 1, (it + 1) * TK);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        if (it == 0) {   // A fragments of this warp's 16 query rows, all four 16-dim k-steps
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const int r = warp * 16 + (lane & 15), c = ks * 2 + (lane >> 4);
                ldsm4(sQ + tile_off(r, c), qh[ks][0], qh[ks][1], qh[ks][2], qh[ks][3]);
                ldsm4(sQ + TILE_B + tile_off(r, c), ql[ks][0], ql[ks][1], ql[ks][2], ql[ks][3]);
            }
        }
        const uint32_t sK = sKV + (it & 1) * 4 * TILE_B, sV = sK + 2 * TILE_B;
        // ---- S = Q K^T for 16 rows x 64 keys: main and correction accumulators
        float s_m[8][4], s_c[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) { s_m[i][j] = 0.0f; s_c[i][j] = 0.0f; }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {   // two 8-key n-tiles per ldmatrix.x4
                // B fragment (k = dim, n = key) from K[key][dim]: matrices (keys 0-7, dims 0-7), (keys 0-7, dims 8-15), (keys 8-15, ...)
                const int r = np * 16 + (lane & 7) + ((lane >> 4) << 3), c = ks * 2 + ((lane >> 3) & 1);
                uint32_t kh0, kh1, kh2, kh3, kl0, kl1, kl2, kl3;
                ldsm4(sK + tile_off(r, c), kh0, kh1, kh2, kh3);
                ldsm4(sK + TILE_B + tile_off(r, c), kl0, kl1, kl2, kl3);
                mma(s_m[2 * np], qh[ks], kh0, kh1);
                mma(s_m[2 * np + 1], qh[ks], kh2, kh3);
                mma(s_c[2 * np], qh[ks], kl0, kl1);
                mma(s_c[2 * np + 1], qh[ks], kl2, kl3);
                mma(s_c[2 * np], ql[ks], kh0, kh1);
                mma(s_c[2 * np + 1], ql[ks], kh2, kh3);
            }
        }
        // ---- online softmax over the tile (rows g and g + 8 of the warp's 16; a row lives in the 4 lanes of a quad)
        const int k0 = it * TK;
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float v = fmaf(s_c[i][j], 1.0f / 2048.0f, s_m[i][j]);
                if (k0 + i * 8 + 2 * t + (j & 1) >= T) v = -INFINITY;
                s_m[i][j] = v;
                mx[j >> 1] = fmaxf(mx[j >> 1], v);
            }
        float corr[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            mx[e] = fmaxf(mx[e], __shfl_xor_sync(0xffffffffu, mx[e], 1));
            mx[e] = fmaxf(mx[e], __shfl_xor_sync(0xffffffffu, mx[e], 2));
            const float mn = fmaxf(m_row[e], mx[e]);
            corr[e] = expf(m_row[e] - mn);
            m_row[e] = mn;
        }
        float ps[2] = {0.0f, 0.0f};
        uint32_t ph[4][4], pl[4][4];   // P as A fragments of the P.V product: k-step = 16 keys = S n-tiles 2ks, 2ks + 1
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float p0 = expf(s_m[i][0] - m_row[0]), p1 = expf(s_m[i][1] - m_row[0]);
            const float p2 = expf(s_m[i][2] - m_row[1]), p3 = expf(s_m[i][3] - m_row[1]);
            ps[0] += p0 + p1;
            ps[1] += p2 + p3;
            // A fragment: a0 = (row g, k 2t..), a1 = (row g+8, k 2t..), a2 = (row g, k 2t+8..), a3 = (row g+8, k 2t+8..)
            const int ks = i >> 1, hi_half = i & 1;
            split2(p0, p1, ph[ks][hi_half * 2], pl[ks][hi_half * 2]);
            split2(p2, p3, ph[ks][hi_half * 2 + 1], pl[ks][hi_half * 2 + 1]);
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            ps[e] += __shfl_xor_sync(0xffffffffu, ps[e], 1);
            ps[e] += __shfl_xor_sync(0xffffffffu, ps[e], 2);
            l_row[e] = l_row[e] * corr[e] + ps[e];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            o_m[i][0] *= corr[0]; o_m[i][1] *= corr[0]; o_m[i][2] *= corr[1]; o_m[i][3] *= corr[1];
            o_c[i][0] *= corr[0]; o_c[i][1] *= corr[0]; o_c[i][2] *= corr[1]; o_c[i][3] *= corr[1];
        }
        // ---- O += P V
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {   // two 8-dim n-tiles per ldmatrix.x4.trans
                // B fragment (k = key, n = dim) from V[key][dim] through the transposing load:
                // matrices (keys 0-7, dims 0-7), (keys 8-15, dims 0-7), (keys 0-7, dims 8-15), (keys 8-15, dims 8-15)
                const int r = ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), c = np * 2 + (lane >> 4);
                uint32_t vh0, vh1, vh2, vh3, vl0, vl1, vl2, vl3;
                ldsm4t(sV + tile_off(r, c), vh0, vh1, vh2, vh3);
                ldsm4t(sV + TILE_B + tile_off(r, c), vl0, vl1, vl2, vl3);
                mma(o_m[2 * np], ph[ks], vh0, vh1);
                mma(o_m[2 * np + 1], ph[ks], vh2, vh3);
                mma(o_c[2 * np], ph[ks], vl0, vl1);
                mma(o_c[2 * np + 1], ph[ks], vl2, vl3);
                mma(o_c[2 * np], pl[ks], vh0, vh1);
                mma(o_c[2 * np + 1], pl[ks], vh2, vh3);
            }
        }
        __syncthreads();   // every warp is done with this stage before the next prefetch overwrites it
    }
    // ---- normalise (softmax denominator) and write the fp16 planes of the out-projection's input
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int q = q0 + warp * 16 + g + 8 * e;
        if (q >= T) continue;
        const float inv = l_row[e];
        __half* oh = out_hi + (win.row_off + q) * (int64_t)d + h * HD;
        __half* ol = out_lo + (win.row_off + q) * (int64_t)d + h * HD;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float a = __fdiv_rn(fmaf(o_c[i][2 * e], 1.0f / 2048.0f, o_m[i][2 * e]), inv);
            const float b = __fdiv_rn(fmaf(o_c[i][2 * e + 1], 1.0f / 2048.0f, o_m[i][2 * e + 1]), inv);
            uint32_t hi, lo;
            split2(a, b, hi, lo);
            *reinterpret_cast<uint32_t*>(oh + i * 8 + 2 * t) = hi;
            *reinterpret_cast<uint32_t*>(ol + i * 8 + 2 * t) = lo;
        }
    }
}

}  // namespace

void launch_encoder_attention_tc(const __half* qkv_hi, const __half* qkv_lo, __half* out_hi, __half* out_lo, const AttnWindow* win_dev,
                                 int n_windows, int max_T, int d, int n_head, cudaStream_t st) {
    WB_REQUIRE(d == n_head * HD, "attention: head dim must be 64");
    static bool attr_set[16] = {};
    int dev = 0;
    WB_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 16 && !attr_set[dev]) {
        WB_CUDA(cudaFuncSetAttribute(enc_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AT_SMEM));
        attr_set[dev] = true;
    }
    dim3 grid((max_T + TQ - 1) / TQ, n_head, n_windows);
    enc_attn_tc_kernel<<<grid, AT_THREADS, AT_SMEM, st>>>(qkv_hi, qkv_lo, out_hi, out_lo, win_dev, d);
    WB_LAUNCH_CHECK();
}

}  // namespace wb

// Tensor-core GEMM for the encoder: tcgen05.mma (kind::tf32) fed by TMA, accumulators in TMEM.
// (reference ops: burn nn::Linear / Conv1d at src/model/mod.rs:243-244, :376-382, :429-435, :484-485)
//
//   C[g][m][n] = epi( sum_k A[g][m][k] * B[n][k] )      same contract and epilogues as gemm.cu
//
// Precision: the reference computes in f32 and the parity bar is identical greedy tokens, so single-pass
// TF32 (10-bit mantissa) is not acceptable.  Every weight of a released Whisper checkpoint is
// fp16-representable, hence EXACT in TF32; the activations are split into two TF32 planes
// A = A_hi + A_lo (A_hi = rna_tf32(A), A_lo = A - A_hi, ~22 mantissa bits together) by the producing
// kernel, and the k-loop runs over [A_hi | A_lo] against the same B tile: 2 MMAs per k-block, fp32
// accumulation in TMEM.  Result error is fp32-class (measured vs the oracle in tests/).
//
// Kernel shape (one 128 x BN output tile per CTA, 192 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor (3-D maps: k, row, window) into a 4-stage
//               128B-swizzled shared-memory ring, mbarrier expect_tx / complete_tx
//   warp 1      TMEM allocation + single-thread tcgen05.mma issue (UMMA 128 x BN x 8, 4 per k-block),
//               tcgen05.commit releases ring slots and finally signals the epilogue
//   warps 2-5   epilogue: tcgen05.ld (32 lanes x 16 columns) -> bias / GELU / q,k scale / pos-emb /
//               residual -> global stores
// The conv stems use the same kernel: their A rows are overlapping windows of a token-major buffer,
// expressed as a tensor map whose row stride is smaller than the row length.
#include <cuda.h>

#include <cstring>

#include "wb_internal.h"

namespace wb {

namespace {

constexpr int TC_BM = 128, TC_BK = 32, TC_STAGES = 4;
constexpr int TC_THREADS = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::tf32, single CTA
__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_c),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, SWIZZLE_128B operand tile (rows of 128 bytes, 8-row groups of 1024 bytes)
__device__ __forceinline__ uint64_t make_smem_desc(const void* p) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(p) & 0x3FFFF) >> 4);   // start address >> 4        bits [0,14)
    d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major) = 1
    d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset: 8 rows * 128 B    bits [32,46)
    d |= (uint64_t)1 << 46;                          // descriptor version 1 (sm_100)
    d |= (uint64_t)2 << 61;                          // layout type SWIZZLE_128B
    return d;
}

__device__ __forceinline__ float gelu_erf(float x) {
    const float t = __fadd_rn(erff(__fdiv_rn(x, 1.41421356237309504880f)), 1.0f);
    return __fdiv_rn(__fmul_rn(x, t), 2.0f);
}

struct TcArgs {
    float* C;
    float* C_lo;      // when set: C receives the TF32 hi plane of the result and C_lo the lo plane (feeds the next GEMM)
    __half* C16;
    const float* bias;
    const float* residual;
    const float* pos;
    const GemmGroup* groups;   // device array or null
    GemmGroup single;
    int64_t ldc;
    int N, K;
    int act;
    float scale;
    int scale_cols;
    int a_rows_per_group;      // tensor-map row coordinate of group g's first A row = g-dim coordinate (3-D map)
};

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tf32_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                    const __grid_constant__ CUtensorMap map_b, const TcArgs g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr int A_BYTES = TC_BM * TC_BK * 4;   // 16 KB
    constexpr int B_BYTES = BN * TC_BK * 4;
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = base;
    uint8_t* sb = base + TC_STAGES * A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sb + TC_STAGES * B_BYTES);
    uint64_t* empty = full + TC_STAGES;
    uint64_t* tmem_full = empty + TC_STAGES;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const GemmGroup grp = g.groups ? g.groups[blockIdx.z] : g.single;
    const int m0 = blockIdx.y * TC_BM;
    if (m0 >= grp.rows) return;   // uniform per CTA
    const int n0 = blockIdx.x * BN;
    const int nkb = (g.K + TC_BK - 1) / TC_BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // TMEM: BN fp32 accumulator columns (power of two >= 32)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_c = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer: k-blocks of A_hi then A_lo, each against the same B k-block
            for (int i = 0; i < 2 * nkb; ++i) {
                const int s = i % TC_STAGES;
                const uint32_t ph = (i / TC_STAGES) & 1;
                mbar_wait(empty + s, ph ^ 1);
                mbar_expect_tx(full + s, A_BYTES + B_BYTES);
                const int kb = i < nkb ? i : i - nkb;
                tma_load_3d(sa + s * A_BYTES, i < nkb ? &map_a_hi : &map_a_lo, full + s, kb * TC_BK, m0, blockIdx.z);
                tma_load_2d(sb + s * B_BYTES, &map_b, full + s, kb * TC_BK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer (one thread)
            // instruction descriptor: D=F32 (1<<4), A=TF32 (2<<7), B=TF32 (2<<10), K-major both, N>>3 at 17, M>>4 at 24
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            for (int i = 0; i < 2 * nkb; ++i) {
                const int s = i % TC_STAGES;
                const uint32_t ph = (i / TC_STAGES) & 1;
                mbar_wait(full + s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t da = make_smem_desc(sa + s * A_BYTES);
                const uint64_t db = make_smem_desc(sb + s * B_BYTES);
#pragma unroll
                for (int k = 0; k < TC_BK / 8; ++k)   // UMMA_K = 8 tf32 = 32 bytes -> +2 in the (>>4) address field
                    umma_tf32(tmem_c, da + 2 * k, db + 2 * k, idesc, (i | k) != 0 ? 1u : 0u);
                umma_commit(empty + s);   // frees the ring slot when these MMAs have read it
            }
            umma_commit(tmem_full);       // accumulator complete
        }
    } else {
        // ===== epilogue warps 2..5: TMEM lane quarter = warp % 4
        const int q = warp & 3;
        const int m = m0 + q * 32 + lane;
        mbar_wait(tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const bool row_ok = m < grp.rows;
        const int64_t crow = grp.c_off + (int64_t)m * g.ldc;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
            uint32_t r[16];
            const uint32_t taddr = tmem_c + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row_ok) {
#pragma unroll
                for (int j4 = 0; j4 < 16; j4 += 4) {
                    const int n = n0 + c0 + j4;
                    float v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float t = __uint_as_float(r[j4 + j]);
                        if (g.bias) t = __fadd_rn(t, __ldg(g.bias + n + j));
                        if (g.act == ACT_GELU) t = gelu_erf(t);
                        if (n + j < g.scale_cols) t = __fmul_rn(t, g.scale);
                        v[j] = t;
                    }
                    if (g.pos) {
                        const float4 p4 = __ldg(reinterpret_cast<const float4*>(g.pos + (int64_t)m * g.N + n));
                        v[0] = __fadd_rn(v[0], p4.x); v[1] = __fadd_rn(v[1], p4.y);
                        v[2] = __fadd_rn(v[2], p4.z); v[3] = __fadd_rn(v[3], p4.w);
                    }
                    if (g.residual) {
                        const float4 r4 = *reinterpret_cast<const float4*>(g.residual + crow + n);
                        v[0] = __fadd_rn(r4.x, v[0]); v[1] = __fadd_rn(r4.y, v[1]);
                        v[2] = __fadd_rn(r4.z, v[2]); v[3] = __fadd_rn(r4.w, v[3]);
                    }
                    if (g.C16) {
                        *reinterpret_cast<__half2*>(g.C16 + crow + n) = __floats2half2_rn(v[0], v[1]);
                        *reinterpret_cast<__half2*>(g.C16 + crow + n + 2) = __floats2half2_rn(v[2], v[3]);
                    } else if (g.C_lo) {
                        float h[4], l[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint32_t t;
                            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v[j]));
                            h[j] = __uint_as_float(t);
                            l[j] = __fsub_rn(v[j], h[j]);
                        }
                        *reinterpret_cast<float4*>(g.C + crow + n) = make_float4(h[0], h[1], h[2], h[3]);
                        *reinterpret_cast<float4*>(g.C_lo + crow + n) = make_float4(l[0], l[1], l[2], l[3]);
                    } else {
                        *reinterpret_cast<float4*>(g.C + crow + n) = make_float4(v[0], v[1], v[2], v[3]);
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_c), "n"(BN) : "memory");
    }
}

// A -> (hi, lo) TF32 planes: hi = round-to-nearest TF32, lo = A - hi (exact in fp32)
__global__ void split_tf32_kernel(const float4* __restrict__ src, float4* __restrict__ hi, float4* __restrict__ lo, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = src[i];
        float4 h, l;
        uint32_t t;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.x)); h.x = __uint_as_float(t); l.x = __fsub_rn(v.x, h.x);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.y)); h.y = __uint_as_float(t); l.y = __fsub_rn(v.y, h.y);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.z)); h.z = __uint_as_float(t); l.z = __fsub_rn(v.z, h.z);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.w)); h.w = __uint_as_float(t); l.w = __fsub_rn(v.w, h.w);
        hi[i] = h;
        lo[i] = l;
    }
}

// ---- host: tensor maps ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        WB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) fail(WB_ERR_CUDA, "cuTensorMapEncodeTiled not available");
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

// fp32 tensor [dim2][dim1][dim0] with element strides (1, s1, s2); box (b0, b1, 1); 128B swizzle
CUtensorMap make_map(const float* base, uint64_t dim0, uint64_t dim1, uint64_t dim2, uint64_t s1, uint64_t s2, uint32_t b0,
                     uint32_t b1, int rank) {
    CUtensorMap m;
    std::memset(&m, 0, sizeof(m));
    cuuint64_t dims[3] = {dim0, dim1, dim2};
    cuuint64_t strides[2] = {s1 * sizeof(float), s2 * sizeof(float)};
    cuuint32_t box[3] = {b0, b1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, (void*)base, dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) fail(WB_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
    return m;
}

template <int BN>
void launch_tc_t(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& b, const TcArgs& a, dim3 grid, cudaStream_t st) {
    constexpr size_t smem = 1024 + (size_t)TC_STAGES * (TC_BM * TC_BK * 4 + BN * TC_BK * 4) + 256;
    static bool configured = false;
    if (!configured) {
        WB_CUDA(cudaFuncSetAttribute(gemm_tf32_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    gemm_tf32_tc_kernel<BN><<<grid, TC_THREADS, smem, st>>>(ah, al, b, a);
    WB_LAUNCH_CHECK();
}

}  // namespace

void launch_split_tf32(const float* src, float* hi, float* lo, int64_t n, cudaStream_t st) {
    WB_REQUIRE(n % 4 == 0, "split: length must be a multiple of 4");
    const int64_t n4 = n / 4;
    if (n4 == 0) return;
    const int blocks = (int)std::min<int64_t>((n4 + 255) / 256, 148 * 8);
    split_tf32_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<float4*>(hi),
                                               reinterpret_cast<float4*>(lo), n4);
    WB_LAUNCH_CHECK();
}

bool gemm_tc_supported(const GemmParams& p) {
    return p.N % 64 == 0 && p.K % 4 == 0 && p.lda % 4 == 0 && p.ldc % 4 == 0;
}

// p.A = hi plane, a_lo = lo plane (same geometry).  Groups: host mirror needed for the tensor-map extents:
// group g's A rows start at element offset g * a_group_stride (p.groups[g].a_off must equal that).
void launch_gemm_tc(const GemmParams& p, const float* a_lo, int64_t a_group_stride, int a_rows_total_per_group, cudaStream_t st) {
    WB_REQUIRE(gemm_tc_supported(p), "gemm_tc: unsupported shape");
    if (p.max_rows <= 0) return;
    const int ng = p.groups ? p.n_groups : 1;
    // A: dims (K, rows, groups), strides (lda, group stride)
    const uint64_t rows = (uint64_t)a_rows_total_per_group;
    const uint64_t gstride = ng > 1 ? (uint64_t)a_group_stride : (uint64_t)p.lda * rows;
    const CUtensorMap ah = make_map(p.A, (uint64_t)p.K, rows, (uint64_t)ng, (uint64_t)p.lda, gstride, TC_BK, TC_BM, 3);
    const CUtensorMap al = make_map(a_lo, (uint64_t)p.K, rows, (uint64_t)ng, (uint64_t)p.lda, gstride, TC_BK, TC_BM, 3);
    const int BN = (p.N % 128 == 0 && (int64_t)(p.N / 128) * ((p.max_rows + TC_BM - 1) / TC_BM) * ng >= 96) ? 128 : 64;
    const CUtensorMap b = make_map(p.B, (uint64_t)p.K, (uint64_t)p.N, 1, (uint64_t)p.K, (uint64_t)p.K * p.N, TC_BK, (uint32_t)BN, 2);
    TcArgs a;
    a.C = p.C; a.C_lo = p.C_lo; a.C16 = p.C16; a.bias = p.bias; a.residual = p.residual; a.pos = p.pos; a.groups = p.groups;
    a.single = GemmGroup{0, 0, p.max_rows};
    a.ldc = p.ldc; a.N = p.N; a.K = p.K; a.act = p.act; a.scale = p.scale; a.scale_cols = p.scale_cols;
    a.a_rows_per_group = a_rows_total_per_group;
    dim3 grid(p.N / BN, (p.max_rows + TC_BM - 1) / TC_BM, ng);
    if (BN == 128) launch_tc_t<128>(ah, al, b, a, grid, st);
    else launch_tc_t<64>(ah, al, b, a, grid, st);
}

}  // namespace wb

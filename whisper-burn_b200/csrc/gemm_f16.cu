// Tensor-core GEMM of the encoder: tcgen05.mma kind::f16 fed by TMA, two fp32 accumulators in TMEM.
// (reference ops: burn nn::Linear / Conv1d at src/model/mod.rs:243-244, :376-382, :429-435, :484-485)
//
//   C[g][m][n] = epi( sum_k A[g][m][k] * B[n][k] )      same contract and epilogues as gemm.cu
//
// Precision: the reference computes in f32 and the parity bar is identical greedy tokens (encoder output within 2e-5 of scale),
// so a single fp16 / bf16 pass is not acceptable.  Every weight of a released Whisper checkpoint is fp16-representable (they
// are stored in fp16), so B is EXACT in fp16; the fp32 activations travel between the encoder kernels as a PAIR of fp16 planes
//     A = A_hi + A_lo / 2048,   A_hi = fp16(A),  A_lo = fp16((A - A_hi) * 2048)        (22 mantissa bits, decoder5.cu's split)
// written by the producing kernel (LayerNorm, attention, the GELU epilogue below).  Every product A_hi*B, A_lo*B is exact in the
// fp32 accumulator; the two planes accumulate into TWO TMEM accumulators (hi at column 0, lo at column BN) that the epilogue
// combines as hi + lo / 2048.  Against the TF32 hi/lo formulation this kernel replaces: half the bytes per k-block through
// shared memory, the weight tile loaded once instead of twice, and twice the MMA rate.
//
// Kernel shape (one 128 x BN output tile per CTA, 192 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor (3-D maps: k, row, window) of A_hi, A_lo (128 x 64 halves each) and B
//               (BN x 64) into a 2-stage 128B-swizzled ring, mbarrier expect_tx / complete_tx (two CTAs per SM: four stages in flight)
//   warp 1      TMEM allocation (2 * BN columns) + single-thread tcgen05.mma issue (UMMA 128 x BN x 16, 4 + 4 per k-block),
//               tcgen05.commit releases ring slots and finally signals the epilogue
//   warps 2-5   epilogue: tcgen05.ld (32 lanes x 16 columns, both accumulators) -> bias / GELU / q,k scale / pos-emb /
//               residual -> fp32 rows and / or fp16 hi/lo planes for the next kernel
// The conv stems use the same kernel: their A rows are overlapping windows of a token-major buffer, expressed as a tensor map
// whose row stride is smaller than the row length.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstring>
#include <mutex>

#include "wb_internal.h"

namespace wb {

namespace {

constexpr int F_BM = 128, F_BK = 64, F_STAGES = 2;   // 2 stages x 48 KB: TWO CTAs per SM (2 x 256 TMEM columns), so one tile's epilogue
                                                     // (TMEM -> registers -> GELU -> global) overlaps the other's TMA / MMA main loop
constexpr int F_THREADS = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)), "l"(map),
                 "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)), "l"(map),
                 "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::f16 (fp16 operands, fp32 accumulate), single CTA
__device__ __forceinline__ void umma_f16(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
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
__device__ __forceinline__ void split_pair(float x, float y, __half2& hi, __half2& lo) {
    hi = __floats2half2_rn(x, y);
    const float2 f = __half22float2(hi);
    lo = __floats2half2_rn((x - f.x) * 2048.0f, (y - f.y) * 2048.0f);
}

struct F16Args {
    float* C;                 // fp32 result rows (may be null when only planes are wanted)
    __half *P_hi, *P_lo;      // fp16 hi / lo planes of the result (may be null); same row geometry as C
    const float* bias;
    const float* residual;
    const float* pos;
    const GemmGroup* groups;  // device array or null
    GemmGroup single;
    int64_t ldc;
    int N, K;
    int act;
    float scale;
    int scale_cols;
};

template <int BN>
__global__ void __launch_bounds__(F_THREADS, 2)
gemm_f16_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                   const __grid_constant__ CUtensorMap map_b, const F16Args g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr int A_BYTES = F_BM * F_BK * 2;   // 16 KB per plane
    constexpr int B_BYTES = BN * F_BK * 2;
    constexpr int STAGE = 2 * A_BYTES + B_BYTES;
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(base + F_STAGES * STAGE);
    uint64_t* empty = full + F_STAGES;
    uint64_t* tmem_full = empty + F_STAGES;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const GemmGroup grp = g.groups ? g.groups[blockIdx.z] : g.single;
    const int m0 = blockIdx.y * F_BM;
    if (m0 >= grp.rows) return;   // uniform per CTA
    const int n0 = blockIdx.x * BN;
    const int nkb = (g.K + F_BK - 1) / F_BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        for (int s = 0; s < F_STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // TMEM: 2 * BN fp32 accumulator columns (power of two >= 32)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(2 * BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_c = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer
            for (int i = 0; i < nkb; ++i) {
                const int s = i % F_STAGES;
                const uint32_t ph = (i / F_STAGES) & 1;
                mbar_wait(empty + s, ph ^ 1);
                mbar_expect_tx(full + s, STAGE);
                uint8_t* st = base + s * STAGE;
                tma_load_3d(st, &map_a_hi, full + s, i * F_BK, m0, blockIdx.z);
                tma_load_3d(st + A_BYTES, &map_a_lo, full + s, i * F_BK, m0, blockIdx.z);
                tma_load_2d(st + 2 * A_BYTES, &map_b, full + s, i * F_BK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer (one thread)
            // instruction descriptor: D = F32 (1 << 4), A = B = F16 (format 0), K-major both, N >> 3 at bit 17, M >> 4 at bit 24
            const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(F_BM >> 4) << 24);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % F_STAGES;
                const uint32_t ph = (i / F_STAGES) & 1;
                mbar_wait(full + s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint8_t* st = base + s * STAGE;
                const uint64_t dah = make_smem_desc(st), dal = make_smem_desc(st + A_BYTES), db = make_smem_desc(st + 2 * A_BYTES);
#pragma unroll
                for (int k = 0; k < F_BK / 16; ++k) {   // UMMA_K = 16 halves = 32 bytes -> +2 in the (>>4) address field
                    umma_f16(tmem_c, dah + 2 * k, db + 2 * k, idesc, (i | k) != 0 ? 1u : 0u);
                    umma_f16(tmem_c + BN, dal + 2 * k, db + 2 * k, idesc, (i | k) != 0 ? 1u : 0u);
                }
                umma_commit(empty + s);   // frees the ring slot when these MMAs have read it
            }
            umma_commit(tmem_full);       // accumulators complete
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
            uint32_t rh[16], rl[16];
            const uint32_t taddr = tmem_c + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(rh[0]), "=r"(rh[1]), "=r"(rh[2]), "=r"(rh[3]), "=r"(rh[4]), "=r"(rh[5]), "=r"(rh[6]), "=r"(rh[7]), "=r"(rh[8]),
                  "=r"(rh[9]), "=r"(rh[10]), "=r"(rh[11]), "=r"(rh[12]), "=r"(rh[13]), "=r"(rh[14]), "=r"(rh[15])
                : "r"(taddr)
                : "memory");
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(rl[0]), "=r"(rl[1]), "=r"(rl[2]), "=r"(rl[3]), "=r"(rl[4]), "=r"(rl[5]), "=r"(rl[6]), "=r"(rl[7]), "=r"(rl[8]),
                  "=r"(rl[9]), "=r"(rl[10]), "=r"(rl[11]), "=r"(rl[12]), "=r"(rl[13]), "=r"(rl[14]), "=r"(rl[15])
                : "r"(taddr + (uint32_t)BN)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row_ok) {
#pragma unroll
                for (int j4 = 0; j4 < 16; j4 += 4) {
                    const int n = n0 + c0 + j4;
                    float v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float t = fmaf(__uint_as_float(rl[j4 + j]), 1.0f / 2048.0f, __uint_as_float(rh[j4 + j]));
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
                    if (g.C) *reinterpret_cast<float4*>(g.C + crow + n) = make_float4(v[0], v[1], v[2], v[3]);
                    if (g.P_hi) {
                        __half2 h0, l0, h1, l1;
                        split_pair(v[0], v[1], h0, l0);
                        split_pair(v[2], v[3], h1, l1);
                        uint2 uh, ul;
                        uh.x = *reinterpret_cast<uint32_t*>(&h0); uh.y = *reinterpret_cast<uint32_t*>(&h1);
                        ul.x = *reinterpret_cast<uint32_t*>(&l0); ul.y = *reinterpret_cast<uint32_t*>(&l1);
                        *reinterpret_cast<uint2*>(g.P_hi + crow + n) = uh;
                        *reinterpret_cast<uint2*>(g.P_lo + crow + n) = ul;
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_c), "n"(2 * BN) : "memory");
    }
}

// fp32 -> (hi, lo) fp16 planes
__global__ void split_f16_kernel(const float4* __restrict__ src, uint2* __restrict__ hi, uint2* __restrict__ lo, int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = src[i];
        __half2 h0, l0, h1, l1;
        split_pair(v.x, v.y, h0, l0);
        split_pair(v.z, v.w, h1, l1);
        uint2 uh, ul;
        uh.x = *reinterpret_cast<uint32_t*>(&h0); uh.y = *reinterpret_cast<uint32_t*>(&h1);
        ul.x = *reinterpret_cast<uint32_t*>(&l0); ul.y = *reinterpret_cast<uint32_t*>(&l1);
        hi[i] = uh;
        lo[i] = ul;
    }
}

// ---- host: tensor maps ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static std::once_flag once;
    static EncodeTiledFn fn = nullptr;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
    });
    if (!fn) fail(WB_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    return fn;
}

// fp16 tensor [dim2][dim1][dim0] with element strides (1, s1, s2); box (b0, b1, 1); 128B swizzle
CUtensorMap make_map(const __half* base, uint64_t dim0, uint64_t dim1, uint64_t dim2, uint64_t s1, uint64_t s2, uint32_t b0, uint32_t b1, int rank) {
    CUtensorMap m;
    std::memset(&m, 0, sizeof(m));
    cuuint64_t dims[3] = {dim0, dim1, dim2};
    cuuint64_t strides[2] = {s1 * sizeof(__half), s2 * sizeof(__half)};
    cuuint32_t box[3] = {b0, b1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) fail(WB_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
    return m;
}

template <int BN>
void launch_f16_t(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& b, const F16Args& a, dim3 grid, cudaStream_t st) {
    constexpr size_t smem = 1024 + (size_t)F_STAGES * (2 * F_BM * F_BK * 2 + BN * F_BK * 2) + 256;
    static std::mutex mu;
    static bool configured[16] = {};   // per device ordinal
    int dev = 0;
    WB_CUDA(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lock(mu);
        if (dev >= 0 && dev < 16 && !configured[dev]) {
            WB_CUDA(cudaFuncSetAttribute(gemm_f16_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured[dev] = true;
        }
    }
    gemm_f16_tc_kernel<BN><<<grid, F_THREADS, smem, st>>>(ah, al, b, a);
    WB_LAUNCH_CHECK();
}

}  // namespace

void launch_split_f16(const float* src, __half* hi, __half* lo, int64_t n, cudaStream_t st) {
    WB_REQUIRE(n % 4 == 0, "split: length must be a multiple of 4");
    const int64_t n4 = n / 4;
    if (n4 == 0) return;
    const int blocks = (int)std::min<int64_t>((n4 + 255) / 256, 148 * 8);
    split_f16_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<uint2*>(hi), reinterpret_cast<uint2*>(lo), n4);
    WB_LAUNCH_CHECK();
}

bool gemm_f16_supported(const GemmF16Params& p) { return p.N % 64 == 0 && p.K % 8 == 0 && p.lda % 8 == 0 && p.ldc % 4 == 0; }

// Tensor maps depend only on (pointers, shapes): a plan is built once per GEMM site and geometry and reused by every launch.
struct GemmF16Plan::Impl {
    CUtensorMap ah, al, b;
    F16Args args;
    dim3 grid;
    int BN;
};
GemmF16Plan::GemmF16Plan() = default;
GemmF16Plan::~GemmF16Plan() { delete impl; }

void GemmF16Plan::build(const GemmF16Params& p, int64_t a_group_stride, int a_rows_total_per_group) {
    WB_REQUIRE(gemm_f16_supported(p), "gemm_f16: unsupported shape");
    delete impl;
    impl = new Impl();
    const int ng = p.groups ? p.n_groups : 1;
    const uint64_t rows = (uint64_t)a_rows_total_per_group;
    const uint64_t gstride = ng > 1 ? (uint64_t)a_group_stride : (uint64_t)p.lda * rows;
    impl->ah = make_map(p.A_hi, (uint64_t)p.K, rows, (uint64_t)ng, (uint64_t)p.lda, gstride, F_BK, F_BM, 3);
    impl->al = make_map(p.A_lo, (uint64_t)p.K, rows, (uint64_t)ng, (uint64_t)p.lda, gstride, F_BK, F_BM, 3);
    impl->BN = (p.N % 128 == 0 && (int64_t)(p.N / 128) * ((p.max_rows + F_BM - 1) / F_BM) * ng >= 96) ? 128 : 64;
    impl->b = make_map(p.B, (uint64_t)p.K, (uint64_t)p.N, 1, (uint64_t)p.K, (uint64_t)p.K * p.N, F_BK, (uint32_t)impl->BN, 2);
    F16Args& a = impl->args;
    a.C = p.C; a.P_hi = p.P_hi; a.P_lo = p.P_lo; a.bias = p.bias; a.residual = p.residual; a.pos = p.pos; a.groups = p.groups;
    a.single = GemmGroup{0, 0, p.max_rows};
    a.ldc = p.ldc; a.N = p.N; a.K = p.K; a.act = p.act; a.scale = p.scale; a.scale_cols = p.scale_cols;
    impl->grid = dim3(p.N / impl->BN, (p.max_rows + F_BM - 1) / F_BM, ng);
    key_rows = p.max_rows;
    key_groups = ng;
}

void GemmF16Plan::launch(cudaStream_t st) const {
    if (!impl || impl->args.single.rows <= 0 || impl->grid.y == 0) return;
    if (impl->BN == 128) launch_f16_t<128>(impl->ah, impl->al, impl->b, impl->args, impl->grid, st);
    else launch_f16_t<64>(impl->ah, impl->al, impl->b, impl->args, impl->grid, st);
}

}  // namespace wb

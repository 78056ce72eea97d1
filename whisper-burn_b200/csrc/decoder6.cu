// Head-fused cluster decoder for L2-resident models (d = 128 / 384: tiny.en and the test models), greedy path.
//
// Same math and single-launch structure as decoder3.cu (TextDecoder::forward src/model/mod.rs:131-157, blocks :345-350,
// attention :428-533, MLP :376-382, search closure src/transcribe.rs:253-307; prefill + every greedy step in one kernel), but
// the per-layer chain is cut from 8 cluster-wide stages (decoder4.cu) to THREE exchanges, and every linear layer runs on the
// 5th-generation tensor cores:
//
//   one thread-block cluster of CS = H * HS CTAs owns one batch row; CTA (h, hs) owns attention head h.
//   phase 1  x -> LN1 -> q_h | k_h | v_h (192 weight rows) -> causal self attention of head h -> the head's K-slice of the
//            out projection: y = Wo[:, 64h..64h+64) . o_h   (a partial d-vector)
//   phase 2  x -> LN2 -> cross query of head h -> cross attention of head h over its share of the window's keys
//            (head-major K/V block streamed by bulk copies) -> y = Wco[:, 64h..) . o_h (un-normalised, with its (max, sum))
//   phase 3  x -> LN3 -> a 4d/CS slice of the MLP hidden layer (GELU) -> y = W2[:, slice] . hid_slice
//   after each phase every CTA sends its partial record (y, max, sum) to ALL CTAs of the cluster with ONE bulk shared-memory ->
//   distributed-shared-memory copy per destination (cp.async.bulk.shared::cluster.shared::cta) that signals the destination's
//   mbarrier with complete_tx; the receiver adds bias + the weighted partials (softmax merge of the key splits happens here,
//   in a fixed order, identically in every CTA) and owns a full copy of the residual stream again.  No hardware cluster barrier
//   inside the step, no cross-thread release/acquire chains: data and its "ready" signal travel together.
//
//   Linear layers = swap-AB tcgen05.mma (kind::f16, M = 128 weight rows, N = 16, K = 16): the weight slices of this CTA are
//   pre-packed per (layer, CTA) as 128-row x 64-column slabs in the canonical K-major 128B-swizzled shared-memory image
//   (dec6_pack_kernel), so a slab is ONE 16 KB bulk copy (TMA engine) into a ring slot and IS the A operand; the activation is
//   the B operand: 16 rows of which row 0 = fp16(x) and row 1 = fp16((x - hi) * 2048) (the decoder5.cu split: exact products,
//   fp32 accumulation), the rest zero.  The accumulator [128 lanes][16 columns] lives in TMEM (two of them, ping-pong); thread
//   `row` of an epilogue warp group reads its lane with tcgen05.ld, combines hi + lo / 2048 and applies bias / scale / GELU.
//   Weights do not depend on activations, so they never wait for the chain: a PRODUCER warp streams weight slabs and the cross
//   K/V block through the ring (full / empty mbarriers), an MMA warp (one elected thread) issues the tensor-core instructions
//   as slabs land and releases the slots with tcgen05.commit; the 8 consumer warps run the dependent chain: LayerNorm,
//   attention (8 lanes per key), epilogues, the exchange.
//
//   Only the vocabulary projection is chip-wide (as in decoder4.cu: bulk-copy ring of contiguous half-tiles of the tied
//   embedding, mma.sync swap-AB with fp16 hi/lo activation planes, fused mask / online softmax / arg-max), behind ONE grid
//   barrier; the per-row finish is done by the LAST CTA to deliver its records (ticket), which then releases a flag.
//
// Requirements: fp16-exact weights, d in {128, 384}, greedy (k = 1), identity ancestry, R <= 24 rows, t_max <= 128.
// Everything else is handled by decoder5.cu / decoder3.cu.
#include <cooperative_groups.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "dec_common.cuh"

namespace cg = cooperative_groups;

namespace wb {

namespace {

constexpr int NCW = 8;                    // consumer warps (threads 0..255)
constexpr int NPROD = 1;                  // producer warps (more than one issuing warp does not raise the stream rate: the tensor cores' shared-memory
                                          // A-operand read, ~0.3 us per 16 KB slab, is what paces the ring; measured, profiles/r02_dec6_mma_trace.txt)
constexpr int W_PROD = NCW, W_MMA = NCW + NPROD;
constexpr int NTH6 = (NCW + NPROD + 1) * 32;      // + producer warps + MMA warp
constexpr int SLOT = 16384;               // bytes per ring slot = one 128-row x 64-column fp16 slab
constexpr int NSLOT = 8;
constexpr int LG_NBUF = 2;                // logits stage: ring slots per warp (aliases the weight ring)
constexpr int BX_SLAB = 2048;             // B operand: 16 rows x 128 bytes per 64-column slab
constexpr uint32_t TMEM_COLS = 32;        // two 16-column accumulators

template <int D, int HS>
struct Geo {
    static constexpr int H = D / 64, CS = H * HS, NS = 4 * D / CS, SEND = D + 4;
    static constexpr int pad128(int n) { return (n + 127) / 128 * 128; }
    // packed weight segments of one (layer, rank), bytes: [tiles of 128 rows (the last one 64 rows when N % 128 == 64)][K / 64 slabs][rows x 128 B]
    static constexpr int OFF_QKV = 0, OFF_O = OFF_QKV + 192 * D * 2, OFF_CQ = OFF_O + D * 64 * 2, OFF_CO = OFF_CQ + 64 * D * 2,
                         OFF_W1 = OFF_CO + D * 64 * 2, OFF_W2 = OFF_W1 + NS * D * 2, PACK = OFF_W2 + D * NS * 2;
    static_assert(D % 128 == 0 && NS % 128 == 0, "only the 192- and 64-row segments end in a 64-row tile");
    // parameter block of one (layer, rank), floats
    static constexpr int P_LN1G = 0, P_LN1B = D, P_LN2G = 2 * D, P_LN2B = 3 * D, P_LN3G = 4 * D, P_LN3B = 5 * D, P_BO = 6 * D,
                         P_BCO = 7 * D, P_B2 = 8 * D, P_BQKV = 9 * D, P_BCQ = 9 * D + 192, P_B1 = 9 * D + 256, P_EPS = 9 * D + 256 + NS, PARAMS = 9 * D + 256 + NS + 4;
    static constexpr int KMAX = D > NS ? D : NS;   // widest B operand
    static_assert(PARAMS % 4 == 0 && SEND % 4 == 0 && NS % 64 == 0 && D % 64 == 0, "layout");
};

// ---- small PTX helpers -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(bar)) : "memory"); }
__device__ long long g_watchdog = 20000000000LL;   // SM clocks a wait may last before the kernel traps (fail loudly instead of hanging the GPU); host-settable
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const long long WATCHDOG = g_watchdog;
    const uint32_t b = s32(bar);
    const long long t0 = clock64();
    for (;;) {
        uint32_t done;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(b), "r"(parity) : "memory");
        if (done) return;
        if (clock64() - t0 > WATCHDOG) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
// local shared memory -> the same offset in CTA `rank` of the cluster, completion on that CTA's mbarrier
__device__ __forceinline__ void bulk_s2peer(void* dst_local, const void* src, uint32_t bytes, uint64_t* bar_local, uint32_t rank) {
    uint32_t rd, rb;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rd) : "r"(s32(dst_local)), "r"(rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(s32(bar_local)), "r"(rank));
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(rd), "r"(s32(src)), "r"(bytes), "r"(rb) : "memory");
}
__device__ __forceinline__ void bar_consumers() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void bar_all() { asm volatile("bar.sync 2, %0;" ::"n"(NTH6) : "memory"); }
__device__ __forceinline__ float group8_sum(float v) {   // sum over the 8 lanes that share lane >> 3 (all lanes converged)
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}
// ---- tcgen05 (see gemm_f16.cu for the same descriptors in a GEMM) ------------------------------------------------------------
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
// the MMA thread polls without reading the clock on every probe (a probe of a pending barrier suspends the thread for a while
// in hardware, so 2^26 failed probes are many seconds): fail loudly instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_lean(uint64_t* bar, uint32_t parity) {
    const uint32_t b = s32(bar);
#pragma unroll 1
    for (int i = 0; i < (1 << 26); ++i) {
        uint32_t done;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(b), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void umma_f16_first(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc) {   // D = A * B
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, 0, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_c),
        "l"(desc_a), "l"(desc_b), "r"(idesc)
        : "memory");
}
__device__ __forceinline__ void umma_f16_acc(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc) {     // D += A * B
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.eq.b32 p, 0, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_c),
        "l"(desc_a), "l"(desc_b), "r"(idesc)
        : "memory");
}
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
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);         // start address >> 4        bits [0,14)
    d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major) = 1
    d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset: 8 rows * 128 B    bits [32,46)
    d |= (uint64_t)1 << 46;                          // descriptor version 1 (sm_100)
    d |= (uint64_t)2 << 61;                          // layout type SWIZZLE_128B
    return d;
}
// byte offset of element (row, k) inside a K-major 128B-swizzled operand whose 64-column slabs are `slab_bytes` apart
__device__ __host__ __forceinline__ uint32_t sw128_off(int row, int k, int slab_bytes) {
    return (uint32_t)((k >> 6) * slab_bytes + row * 128 + ((((k & 63) >> 3) ^ (row & 7)) << 4) + (k & 7) * 2);
}

// which of the 64 head dims is element i (0..7) of lane l8: two 16-byte chunks (fp32: l8 and l8 + 8) / one (fp16: l8)
template <typename KVT>
__device__ __forceinline__ int hdim(int l8, int i) {
    if constexpr (sizeof(KVT) == 4) return (i < 4 ? 4 * l8 : 32 + 4 * l8) + (i & 3);
    else return 8 * l8 + i;
}
// the lane's 8 elements of a 64-dim K or V row; par = 1 on odd positions of the head-major cross layout (XOR-4 chunk swizzle)
template <typename KVT>
__device__ __forceinline__ void load_row8(const KVT* row, int l8, int par, float (&f)[8], bool smem) {
    if constexpr (sizeof(KVT) == 4) {
        const float4* p = reinterpret_cast<const float4*>(row);
        const float4 a = smem ? p[l8 ^ (4 * par)] : __ldcg(p + (l8 ^ (4 * par)));
        const float4 b = smem ? p[(l8 + 8) ^ (4 * par)] : __ldcg(p + ((l8 + 8) ^ (4 * par)));
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    } else {
        const uint4* p = reinterpret_cast<const uint4*>(row);
        const uint4 u = smem ? p[l8 ^ (4 * par)] : __ldcg(p + (l8 ^ (4 * par)));
        cvt8(u, f);
    }
}

struct Softmax8 {   // online softmax state of one (warp, rg) key slot; o = the lane's 8 dims
    float m, l, o[8];
    __device__ __forceinline__ void init() {
        m = -INFINITY; l = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = 0.0f;
    }
    __device__ __forceinline__ void add(float s, const float (&v)[8]) {
        const float mn = fmaxf(m, s);
        const float corr = expf(m - mn), e = expf(s - mn);
        l = l * corr + e;
        m = mn;
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaf(e, v[i], o[i] * corr);
    }
    __device__ __forceinline__ void merge_xor(int off) {   // all lanes converged
        const float m2 = __shfl_xor_sync(0xffffffffu, m, off), l2 = __shfl_xor_sync(0xffffffffu, l, off);
        const float mn = fmaxf(m, m2);
        const float c1 = m > -INFINITY ? expf(m - mn) : 0.0f, c2 = m2 > -INFINITY ? expf(m2 - mn) : 0.0f;
        l = l * c1 + l2 * c2;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float o2 = __shfl_xor_sync(0xffffffffu, o[i], off);
            o[i] = o[i] * c1 + o2 * c2;
        }
        m = mn;
    }
};

// ---- weight / parameter packing (once per session) ---------------------------------------------------------------------
// Segment [N][K] (fp16, source rows n0 + (r / piece) * piece_stride + r % piece, columns k0 .. k0 + K of a [.][ldk] matrix) as
// tiles of 128 rows (a last tile of 64 rows when N % 128 == 64), each tile as K / 64 slabs of rows x 128 B in the K-major
// 128B-swizzled shared-memory image: 16-byte unit (row r, chunk c) of a slab at r * 128 + ((c ^ (r & 7)) << 4).  A slab is what
// one bulk copy moves and what one group of four tcgen05.mma (M = 128 or 64, K = 16 each) reads.
struct PackSeg {
    const __half* src;
    int ldk, n0, k0, N, K;
    int piece, piece_stride;
    int64_t dst_off;   // bytes
};
__global__ void dec6_pack_kernel(const PackSeg* segs, int n_segs, uint8_t* dst) {
    for (int s = blockIdx.y; s < n_segs; s += gridDim.y) {
        const PackSeg g = segs[s];
        const int nslab = g.K / 64;
        const int64_t n16 = (int64_t)g.N * g.K / 8;   // 16-byte units; N is a multiple of 64
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
            // source order: unit i = (row, column chunk)
            const int row = (int)(i / (g.K / 8)), cc = (int)(i % (g.K / 8));
            const int t = row >> 7, r = row & 127, sl = cc >> 3, c = cc & 7;
            const int trows = min(128, g.N - t * 128);                       // rows of this tile: 128 or 64
            const int64_t tile_base = (int64_t)t * 128 * g.K * 2;            // full tiles precede
            const int srow = g.n0 + (row / g.piece) * g.piece_stride + row % g.piece;
            const uint4 v = *reinterpret_cast<const uint4*>(g.src + (int64_t)srow * g.ldk + g.k0 + cc * 8);
            *reinterpret_cast<uint4*>(dst + g.dst_off + tile_base + (int64_t)sl * trows * 128 + r * 128 + ((c ^ (r & 7)) << 4)) = v;
        }
        (void)nslab;
    }
}
struct ParamSeg {
    const float* src;
    int n;
    int64_t dst_off;   // floats
};
__global__ void dec6_param_kernel(const ParamSeg* segs, int n_segs, float* dst) {
    for (int s = blockIdx.x; s < n_segs; s += gridDim.x) {
        const ParamSeg g = segs[s];
        for (int i = threadIdx.x; i < g.n; i += blockDim.x) dst[g.dst_off + i] = g.src[i];
    }
}

// ---- building blocks of a phase.  Code size is latency here (decoder5.cu, lesson 2): the layer body runs once per layer and
// position, i.e. mostly from a cold instruction cache, so every block exists ONCE (noinline, runtime shapes) and the phases are
// short call sequences.

// the B operand of the next linear layer: value v of column k -> fp16 hi in row 0, fp16 (residual * 2048) in row 1
__device__ __forceinline__ void bx_store(uint8_t* bx, int k, float v) {
    const __half h = __float2half_rn(v);
    *reinterpret_cast<__half*>(bx + sw128_off(0, k, BX_SLAB)) = h;
    *reinterpret_cast<__half*>(bx + sw128_off(1, k, BX_SLAB)) = __float2half_rn((v - __half2float(h)) * 2048.0f);
}

struct Pipe {        // shared-memory handles of the ring / tensor-core pipeline
    uint8_t* ring;   // [NSLOT][SLOT]
    uint64_t *full, *empty;      // per slot
    uint64_t* b_ready;           // B operand written (consumers -> MMA warp)
    uint64_t *acc_full, *acc_free;   // [2] accumulator ping-pong (MMA warp <-> epilogue warp groups)
    uint32_t tmem;               // base of the 32 allocated TMEM columns
};
struct Counters {    // progress counters every role keeps in registers (identical sequences by construction)
    uint32_t n;      // ring chunks consumed / issued
    uint32_t tile;   // output tiles
    uint32_t gemv;   // linear layers
};

enum { EM_PLAIN = 0, EM_QKV = 1, EM_CQ = 2, EM_HID = 3 };
template <typename KVT>
struct GemvOut {
    int mode;
    const float* bias;
    float scale;
    float* out;        // EM_PLAIN: y[n]; EM_QKV: qkv_s[n]; EM_CQ: q2_s[n]; EM_HID: hid_s[n]
    KVT *kdst, *vdst;  // EM_QKV: this position's 64-element head slice of the self K / V cache
};

// Consumer side of one linear layer y[n] = sum_k W[n][k] x[k] (n < N, N padded to tiles of 128 rows): the B operand has been
// written; hand it to the MMA warp, then the warp group (tile & 1) reads accumulator (tile & 1) -- TMEM lane = weight row --
// and applies the epilogue.  Ends with a barrier of the consumer warps.
template <typename KVT>
__device__ __noinline__ Counters gemv_epi6(const Pipe P, Counters c, int N, int n_slabs, const GemvOut<KVT> o) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes of the B operand -> tensor-core (async proxy) reads
    bar_consumers();
    if (tid == 0) mbar_arrive(P.b_ready);
    const int n_tiles = (N + 127) >> 7;
#pragma unroll 1
    for (int t = 0; t < n_tiles; ++t) {
        const uint32_t T = c.tile + (uint32_t)t, g = T & 1;
        if ((uint32_t)(warp >> 2) != g) continue;
        mbar_wait(P.acc_full + g, (T >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t r0, r1;
        const uint32_t taddr = P.tmem + ((uint32_t)((warp & 3) * 32) << 16) + g * 16;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(P.acc_free + g);                 // 4 warps: the accumulator may be overwritten
        // M = 128: TMEM lane = tile row; M = 64 (last tile of a 192- / 64-row segment): the 64 rows sit in lanes 0..15 of each quadrant
        const bool half_tile = N - t * 128 < 128;
        const int row = t * 128 + (half_tile ? (warp & 3) * 16 + lane : (warp & 3) * 32 + lane);
        if (row < N && (!half_tile || lane < 16)) {
            const float s = fmaf(__uint_as_float(r1), 1.0f / 2048.0f, __uint_as_float(r0));
            if (o.mode == EM_PLAIN) {
                o.out[row] = s;
            } else if (o.mode == EM_QKV) {       // mod.rs:429-431; q and k carry (d/H)^-0.25 each (:500-503)
                float v = __fadd_rn(s, o.bias[row]);
                if (row < 128) v = __fmul_rn(v, o.scale);
                if (row >= 64) {
                    const KVT r = (KVT)v;          // fp16 cache: round-to-nearest where the value enters the cache
                    (row < 128 ? o.kdst : o.vdst)[row & 63] = r;
                    v = (float)r;
                }
                o.out[row] = v;
            } else if (o.mode == EM_CQ) {        // cross query (mod.rs:483)
                o.out[row] = __fmul_rn(__fadd_rn(s, o.bias[row]), o.scale);
            } else {                             // gelu(LN(x) W1 + b1) (mod.rs:377-378)
                o.out[row] = gelu_erf(__fadd_rn(s, o.bias[row]));
            }
        }
    }
    c.tile += (uint32_t)n_tiles;
    c.n += (uint32_t)(n_tiles * n_slabs);
    c.gemv += 1;
    bar_consumers();
    return c;
}

// MMA warp (one thread): the tensor-core side of the same linear layer
__device__ __forceinline__ Counters gemv_mma6(const Pipe P, Counters c, int N, int n_slabs, uint32_t bx_addr, unsigned long long* tr, int& tn, int tcap) {
    // instruction descriptor: D = F32 (1 << 4), A = B = F16 (format 0), K-major both, N >> 3 = 2 at bit 17, M >> 4 at bit 24
    const int n_tiles = (N + 127) >> 7;
    const uint64_t desc_ring0 = make_smem_desc(s32(P.ring)), desc_bx0 = make_smem_desc(bx_addr);
    mbar_wait_lean(P.b_ready, c.gemv & 1);
    if (tr && tn < tcap) tr[tn++] = (gtime() << 2) | 1ull;   // debug trace: B operand ready
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
    for (int t = 0; t < n_tiles; ++t) {
        const uint32_t T = c.tile, g = T & 1, u = T >> 1;
        const uint32_t idesc = (1u << 4) | (2u << 17) | ((uint32_t)(min(128, N - t * 128) >> 4) << 24);
        if (u >= 1) mbar_wait_lean(P.acc_free + g, (u - 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int s = 0; s < n_slabs; ++s) {
            const uint32_t slot = c.n % NSLOT;
            mbar_wait_lean(P.full + slot, (c.n / NSLOT) & 1);
            if (tr && tn < tcap) tr[tn++] = (gtime() << 2) | 2ull;   // debug trace: slab landed
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // descriptors: ring slots are SLOT apart, B slabs BX_SLAB apart (address field = bytes >> 4)
            const uint64_t da = desc_ring0 + (uint64_t)(slot * (SLOT >> 4)), db = desc_bx0 + (uint64_t)(s * (BX_SLAB >> 4));
            const uint32_t acc = P.tmem + g * 16;
            if (s == 0) umma_f16_first(acc, da, db, idesc); else umma_f16_acc(acc, da, db, idesc);
            umma_f16_acc(acc, da + 2, db + 2, idesc);   // UMMA_K = 16 halves = 32 bytes -> +2 in the (>> 4) address field
            umma_f16_acc(acc, da + 4, db + 4, idesc);
            umma_f16_acc(acc, da + 6, db + 6, idesc);
            umma_commit(P.empty + slot);   // the slab may be overwritten when these MMAs have read it
            ++c.n;
        }
        umma_commit(P.acc_full + g);       // accumulator complete
        ++c.tile;
    }
    ++c.gemv;
    return c;
}

// LayerNorm (burn 0.9 form, dec_common.cuh stage_ln) of the row x_s[D] by the 8 consumer warps: every warp computes the
// statistics for itself (no block reduction), thread t normalises elements t, t + 256 and writes them as the B operand.
template <int D>
__device__ __noinline__ void ln6(const float* x_s, const float* g, const float* b, float eps, int eps_outside, uint8_t* bx) {
    const int tid = threadIdx.x, lane = tid & 31;
    float v[D / 32];
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < D / 32; ++i) { v[i] = x_s[lane + 32 * i]; sum += v[i]; }
    sum = warp_sum(sum);
    const float mean = __fdiv_rn(sum, (float)D);
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < D / 32; ++i) { const float dv = __fsub_rn(v[i], mean); q = __fadd_rn(q, __fmul_rn(dv, dv)); }
    q = warp_sum(q);
    const float var = __fdiv_rn(q, (float)D);
    const float den = eps_outside ? __fadd_rn(__fsqrt_rn(var), eps) : __fsqrt_rn(__fadd_rn(var, eps));
#pragma unroll 1
    for (int c = tid; c < D; c += 256) bx_store(bx, c, __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(x_s[c], mean), den), g[c]), b[c]));
}

// merges the 8 per-warp attention records (wm, wl, wo) into the head's un-normalised output, written as the B operand of the
// out projection, and its (max, sum) record
__device__ __noinline__ void attn_merge6(const float* wm, const float* wl, const float* wo, uint8_t* bx, float* rec, float active) {
    const int tid = threadIdx.x;
    bar_consumers();
    if (tid < 64) {
        float M = -INFINITY;
#pragma unroll
        for (int w2 = 0; w2 < NCW; ++w2) M = fmaxf(M, wm[w2]);
        float Ls = 0.0f, o = 0.0f;
#pragma unroll
        for (int w2 = 0; w2 < NCW; ++w2) {
            const float sc = wm[w2] > -INFINITY ? expf(wm[w2] - M) : 0.0f;
            Ls += sc * wl[w2];
            o += sc * wo[w2 * 64 + tid];
        }
        bx_store(bx, tid, o);
        if (tid == 0) { rec[0] = M; rec[1] = Ls; rec[2] = active; rec[3] = 0.0f; }
    }
}

enum { MODE_SUM = 0, MODE_ATTN = 1 };

// x += bias + sum over sources of weight * partial (fixed order, identical in every CTA of the cluster).  pr = the phase's
// records [CS][D + 4] (y, max, sum, active); MODE_ATTN: softmax merge of the key splits of each head on the receiving side
// (mod.rs:516-527 computed in pieces); MODE_SUM: plain sum.  Ends with a barrier.
template <int D, int HS>
__device__ __noinline__ void combine6(const float* pr, uint64_t* bar, uint32_t parity, int mode, const float* bias, float* x_s, float* wsrc_s) {
    constexpr int H = D / 64, CS = H * HS, SEND = D + 4;
    const int tid = threadIdx.x;
    if (tid == 0) mbar_expect_tx(bar, CS * SEND * 4);
    mbar_wait(bar, parity);
    if (tid < CS) {
        float wgt = 1.0f;
        if (mode == MODE_ATTN) {
            const int hh = tid % H;
            float M = -INFINITY;
#pragma unroll
            for (int s = 0; s < HS; ++s)
                if (pr[(hh + H * s) * SEND + D + 2] != 0.0f) M = fmaxf(M, pr[(hh + H * s) * SEND + D]);
            float den = 0.0f;
#pragma unroll
            for (int s = 0; s < HS; ++s) {
                const float* rec = pr + (hh + H * s) * SEND + D;
                if (rec[2] != 0.0f && rec[0] > -INFINITY) den += expf(rec[0] - M) * rec[1];
            }
            const float* me = pr + tid * SEND + D;
            wgt = (me[2] != 0.0f && me[0] > -INFINITY) ? __fdiv_rn(expf(me[0] - M), den) : 0.0f;
        }
        wsrc_s[tid] = wgt;
    }
    bar_consumers();
#pragma unroll 1
    for (int c = tid; c < D; c += 256) {
        float acc = bias[c];
#pragma unroll 4
        for (int s = 0; s < CS; ++s) acc = fmaf(wsrc_s[s], pr[s * SEND + c], acc);
        x_s[c] = __fadd_rn(x_s[c], acc);
    }
    bar_consumers();
}

// causal self attention of one head over positions 0..p (mod.rs:428-436 with the mask of :535-544 = "keys <= p"): 32 (warp, rg)
// key slots, 8 lanes per key; keys < p come from the cache (L2), key p from shared memory.  Leaves per-warp records in wm/wl/wo.
template <typename KVT>
__device__ __noinline__ void self_attn6(const float* qkv_s, const KVT* kbase, const KVT* vbase, int ld, int p, float* wm, float* wl, float* wo) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, rg = lane >> 3, l8 = lane & 7;
    float q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = qkv_s[hdim<KVT>(l8, i)];
    Softmax8 A;
    A.init();
    float kf[4][8], vf[4][8];
    const int j0 = warp * 4 + rg;
#pragma unroll
    for (int u = 0; u < 4; ++u) {                          // all loads first (t_max <= 128 -> at most 4 keys per slot)
        const int j = j0 + 32 * u;
        if (j < p) {
            load_row8<KVT>(kbase + (int64_t)j * ld, l8, 0, kf[u], false);
            load_row8<KVT>(vbase + (int64_t)j * ld, l8, 0, vf[u], false);
        } else if (j == p) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { kf[u][i] = qkv_s[64 + hdim<KVT>(l8, i)]; vf[u][i] = qkv_s[128 + hdim<KVT>(l8, i)]; }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) { kf[u][i] = 0.0f; vf[u][i] = 0.0f; }
        }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(q[i], kf[u][i], s);
        s = group8_sum(s);
        if (j0 + 32 * u <= p) A.add(s, vf[u]);
    }
    A.merge_xor(8);
    A.merge_xor(16);
    if (rg == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) wo[warp * 64 + hdim<KVT>(l8, i)] = A.o[i];
        if (l8 == 0) { wm[warp] = A.m; wl[warp] = A.l; }
    }
}

// cross attention of one head over keys [k_begin, k_end) of the window (mod.rs:482-490), the head-major K/V block arriving through
// the ring in chunks of KPC keys; 8 lanes per key.  A slot is released by the LAST of the 8 warps to finish with it (the slots'
// empty barriers take one arrival, as tcgen05.commit gives them for weight slabs).  (Waiting for several chunks at once to batch
// the per-key latency chains was measured SLOWER, 6 -> 12 us per layer: the chunks arrive one per ~0.3 us and the batch waits for
// the last one.)  Returns the ring counter.
template <typename KVT>
__device__ __noinline__ uint32_t cross_attn6(const Pipe P, uint32_t n, int* slot_cnt, const float* q2_s, int k_begin, int k_end, float* wm, float* wl, float* wo) {
    constexpr int ROWB = 128 * (int)sizeof(KVT), KPC = SLOT / ROWB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, rg = lane >> 3, l8 = lane & 7;
    float q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = q2_s[hdim<KVT>(l8, i)];
    Softmax8 A;
    A.init();
#pragma unroll 1
    for (int k0 = k_begin; k0 < k_end; k0 += KPC) {
        const int nk = min(KPC, k_end - k0);
        const uint32_t slot = n % NSLOT;
        mbar_wait(P.full + slot, (n / NSLOT) & 1);
        const uint8_t* blk = P.ring + slot * SLOT;
#pragma unroll 2
        for (int kk = warp * 4 + rg; kk < nk; kk += 32) {
            const int par = (k0 + kk) & 1;
            const KVT* rowp = reinterpret_cast<const KVT*>(blk + (size_t)kk * ROWB);
            float kf[8], vf[8];
            load_row8<KVT>(rowp, l8, par, kf, true);
            load_row8<KVT>(rowp + 64, l8, par, vf, true);
            float s = 0.0f;
#pragma unroll
            for (int i = 0; i < 8; ++i) s = fmaf(q[i], kf[i], s);
            const unsigned int gm = 0xffu << (lane & 24);
            s += __shfl_xor_sync(gm, s, 1);
            s += __shfl_xor_sync(gm, s, 2);
            s += __shfl_xor_sync(gm, s, 4);
            A.add(s, vf);
        }
        __syncwarp();
        if (lane == 0 && atomicAdd(slot_cnt + slot, 1) == NCW - 1) {
            slot_cnt[slot] = 0;
            mbar_arrive(P.empty + slot);
        }
        ++n;
    }
    __syncwarp();
    A.merge_xor(8);
    A.merge_xor(16);
    if (rg == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) wo[warp * 64 + hdim<KVT>(l8, i)] = A.o[i];
        if (l8 == 0) { wm[warp] = A.m; wl[warp] = A.l; }
    }
    return n;
}

// =====================================================================================================================
template <int D, int HS, int NT8, typename KVT>
__global__ void __launch_bounds__(NTH6, 1)
dec6_kernel(const Dec3Args a) {
    using G = Geo<D, HS>;
    constexpr int H = G::H, CS = G::CS, NS = G::NS, SEND = G::SEND, PARAMS = G::PARAMS;
    constexpr int KPC = SLOT / (128 * (int)sizeof(KVT));   // cross keys per ring chunk
    constexpr int ROWB = 128 * (int)sizeof(KVT);
    extern __shared__ __align__(1024) unsigned char smraw_[];
    uint8_t* smraw = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smraw_) + 1023) & ~(uintptr_t)1023);
    cg::cluster_group cl = cg::this_cluster();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = (int)cl.block_rank(), h = rank % H, hs = rank / H;
    const int cluster_id = blockIdx.x / CS, n_clusters = gridDim.x / CS;
    const int L = a.L, V = a.V, R = a.R, t_max = a.t_max;

    // ---- shared memory carve-up (ring and B operand 1024-byte aligned: swizzle atoms)
    uint8_t* ring_mem = smraw;                                              // [NSLOT][SLOT]
    uint8_t* bx = ring_mem + NSLOT * SLOT;                                  // [KMAX / 64][16 rows][128 B]  B operand (rows 0, 1 live)
    float* params = reinterpret_cast<float*>(bx + (G::KMAX / 64) * BX_SLAB);   // [2][PARAMS]
    float* part = params + 2 * PARAMS;                                      // [2][CS][SEND] partial records of the cluster
    float* y_s = part + 2 * CS * SEND;                                      // [2][SEND]     this CTA's outgoing record
    float* x_s = y_s + 2 * SEND;                                            // [D]           residual stream (own copy)
    float* qkv_s = x_s + D;                                                 // [192]         q_h | k_h | v_h of the current position
    float* q2_s = qkv_s + 192;                                              // [64]
    float* hid_s = q2_s + 64;                                               // [NS]          MLP hidden slice
    float* wsrc_s = hid_s + NS;                                             // [16]          merge weights of the sources
    float* wm = wsrc_s + 16;                                                // [8]
    float* wl = wm + 8;                                                     // [8]
    float* wo = wl + 8;                                                     // [8][64]
    int* ctl = reinterpret_cast<int*>(wo + 512);                            // [4] stop flag, is_last, tmem base
    int* slot_cnt = ctl + 4;                                                // [NSLOT] warps done with a K/V chunk
    uint64_t* bars = reinterpret_cast<uint64_t*>(slot_cnt + NSLOT);
    uint64_t* full = bars;                    // [NSLOT]
    uint64_t* empty = full + NSLOT;           // [NSLOT]
    uint64_t* pfull = empty + NSLOT;          // [2] parameter block landed
    uint64_t* pfree = pfull + 2;              // [2] consumers are done with the parameter block
    uint64_t* pbar = pfree + 2;               // [2] partial records of a phase landed
    uint64_t* b_ready = pbar + 2;             // [1]
    uint64_t* acc_full = b_ready + 1;         // [2]
    uint64_t* acc_free = acc_full + 2;        // [2]
    uint64_t* lg_bar = acc_free + 2;          // [NCW][LG_NBUF] logits stage
    // logits-stage scratch aliases the (then dead) parameter / partial buffers
    uint4* pl_hi = reinterpret_cast<uint4*>(params);                        // [NT8][D/32][32] fp16 hi plane of the LayerNorm rows, fragment order
    uint4* pl_lo = pl_hi + NT8 * (D / 32) * 32;
    float* red = reinterpret_cast<float*>(pl_lo + NT8 * (D / 32) * 32);     // [NCW][8 * NT8][4]

    if (tid == 0) {
        for (int i = 0; i < NSLOT; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); slot_cnt[i] = 0; }
        for (int i = 0; i < 2; ++i) { mbar_init(pfull + i, 1); mbar_init(pfree + i, NCW); mbar_init(pbar + i, 1); mbar_init(acc_full + i, 1); mbar_init(acc_free + i, 4); }
        mbar_init(b_ready, 1);
        for (int i = 0; i < NCW * LG_NBUF; ++i) mbar_init(lg_bar + i, 1);
        ctl[0] = 0; ctl[1] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < (G::KMAX / 64) * BX_SLAB / 16; i += NTH6) reinterpret_cast<uint4*>(bx)[i] = make_uint4(0, 0, 0, 0);   // rows 2..15 stay zero
    if (warp == W_MMA) {   // TMEM: two 16-column fp32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(ctl + 2)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cl.sync();   // every CTA's mbarriers exist before any peer signals them; TMEM address published
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    Pipe P;
    P.ring = ring_mem; P.full = full; P.empty = empty; P.b_ready = b_ready; P.acc_full = acc_full; P.acc_free = acc_free;
    P.tmem = *reinterpret_cast<volatile uint32_t*>(ctl + 2);

    const uint8_t* pack = reinterpret_cast<const uint8_t*>(a.d6_pack);
    const float* gparams = a.d6_params;
    constexpr int S_D = D / 64, S_NS = NS / 64;   // K slabs of a linear layer

    if (warp >= W_PROD && warp < W_MMA) {
        // ===================================================== PRODUCERS: weight slabs, parameters and cross K/V, in consumer order;
        // producer warp pw issues chunks n with n % NPROD == pw (every producer walks the whole schedule)
        const uint32_t pw = (uint32_t)(warp - W_PROD);
        uint32_t n = 0, pl = 0;
        auto push = [&](const void* src, uint32_t bytes) {
            if (n % NPROD == pw) {
                const uint32_t slot = n % NSLOT;
                if (n >= NSLOT) mbar_wait(empty + slot, ((n / NSLOT) - 1) & 1);
                mbar_expect_tx(full + slot, bytes);
                bulk_g2s(ring_mem + slot * SLOT, src, bytes, full + slot);
            }
            ++n;
        };
        for (int step = 0; step < a.n_steps; ++step) {
            if (lane == 0) {
                for (int row = cluster_id; row < R; row += n_clusters) {
                    const int w = __ldg(a.row_window + row);
                    const int T = __ldg(a.win_T + w);
                    const int per = (T + HS - 1) / HS, k_begin = min(T, hs * per), k_end = min(T, k_begin + per);
                    for (int l = 0; l < L; ++l) {
                        if (pw == 0) {
                            const uint32_t b = pl & 1, u = pl >> 1;
                            if (u >= 1) mbar_wait(pfree + b, (u - 1) & 1);
                            mbar_expect_tx(pfull + b, PARAMS * 4);
                            bulk_g2s(params + b * PARAMS, gparams + ((size_t)l * CS + rank) * PARAMS, PARAMS * 4, pfull + b);
                        }
                        ++pl;
                        const uint8_t* base = pack + ((size_t)l * CS + rank) * G::PACK;
                        auto seg = [&](int off, int N, int K) {   // tiles of 128 rows (last one 64), K / 64 slabs each, one bulk copy per slab
                            const uint8_t* src = base + off;
                            for (int r0 = 0; r0 < N; r0 += 128) {
                                const uint32_t bytes = (uint32_t)min(128, N - r0) * 128;
                                for (int sl = 0; sl < K / 64; ++sl, src += bytes) push(src, bytes);
                            }
                        };
                        if (hs == 0) { seg(G::OFF_QKV, 192, D); seg(G::OFF_O, D, 64); }
                        seg(G::OFF_CQ, 64, D);
                        {
                            const KVT* kv = reinterpret_cast<const KVT*>(a.ckv) + (size_t)l * a.Mcap * 2 * D + __ldg(a.win_row_off + w) * (int64_t)(2 * D) +
                                            ((int64_t)h * T + k_begin) * 128;
                            for (int k0 = k_begin; k0 < k_end; k0 += KPC) push(kv + (int64_t)(k0 - k_begin) * 128, (uint32_t)(min(KPC, k_end - k0) * ROWB));
                        }
                        seg(G::OFF_CO, D, 64);
                        seg(G::OFF_W1, NS, D);
                        seg(G::OFF_W2, D, NS);
                    }
                }
            }
            __syncwarp();
            bar_all();   // the ring is lent to the logits stage until the consumers finish the step
            if (*reinterpret_cast<volatile int*>(ctl) != 0) break;
        }
    } else if (warp == W_MMA) {
        // ===================================================== MMA warp: one thread issues every tcgen05.mma of the step
        Counters c{0, 0, 0};
        const uint32_t bxa = s32(bx);
        unsigned long long* mtr = (a.trace && blockIdx.x == 0) ? a.trace + a.trace_cap / 2 : nullptr;   // second half of the trace buffer
        int mtn = 0;
        const int mcap = a.trace_cap / 2;
        for (int step = 0; step < a.n_steps; ++step) {
            if (lane == 0) {
                for (int row = cluster_id; row < R; row += n_clusters) {
                    const int w = __ldg(a.row_window + row);
                    const int T = __ldg(a.win_T + w);
                    const int per = (T + HS - 1) / HS, k_begin = min(T, hs * per), k_end = min(T, k_begin + per);
                    const uint32_t n_kv = (uint32_t)((k_end - k_begin + KPC - 1) / KPC);
                    for (int l = 0; l < L; ++l) {
                        if (hs == 0) { c = gemv_mma6(P, c, 192, S_D, bxa, mtr, mtn, mcap); c = gemv_mma6(P, c, D, 1, bxa, mtr, mtn, mcap); }
                        c = gemv_mma6(P, c, 64, S_D, bxa, mtr, mtn, mcap);
                        c.n += n_kv;                       // K/V chunks are consumed by the attention warps
                        c = gemv_mma6(P, c, D, 1, bxa, mtr, mtn, mcap);
                        c = gemv_mma6(P, c, NS, S_D, bxa, mtr, mtn, mcap);
                        c = gemv_mma6(P, c, D, S_NS, bxa, mtr, mtn, mcap);
                    }
                }
            }
            __syncwarp();
            bar_all();
            if (*reinterpret_cast<volatile int*>(ctl) != 0) break;
        }
    } else {
        // ===================================================== CONSUMERS
        Counters c{0, 0, 0};
        uint32_t pl = 0;     // parameter blocks consumed
        uint32_t ph = 0;     // exchange phases completed (buffer = ph & 1)
        unsigned int gen = 0;
        unsigned int lg_count = 0;
        int tr_n = 0;
        const float scale = a.qk_scale;
        const int gw = blockIdx.x * NCW + warp, n_gw = gridDim.x * NCW;

        auto trace = [&]() {
            if (a.trace && blockIdx.x == 0 && tid == 0 && tr_n < a.trace_cap / 2) a.trace[tr_n++] = gtime();
        };
        // sends y_s[ph & 1] (D values + max, sum, active) to every CTA of the cluster
        auto send = [&]() {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            bar_consumers();
            if (tid < CS) {   // one issuing thread per destination
                const uint32_t b = ph & 1;
                bulk_s2peer(part + (b * CS + rank) * SEND, y_s + b * SEND, SEND * 4, pbar + b, (uint32_t)tid);
            }
        };
        auto combine = [&](int mode, const float* bias) {
            combine6<D, HS>(part + (ph & 1) * CS * SEND, pbar + (ph & 1), (ph >> 1) & 1, mode, bias, x_s, wsrc_s);
            ++ph;
        };

#pragma unroll 1
        for (int step = 0; step < a.n_steps; ++step) {
            const int p = a.pos0 + step;
            const bool want_logits = p >= a.logits_from;
#pragma unroll 1
            for (int row = cluster_id; row < R; row += n_clusters) {
                // ---- embed (mod.rs:141-146): every CTA of the cluster builds its own copy of x
                {
                    const int tok = __ldcg(a.tokens + (int64_t)row * t_max + p);
                    for (int c2 = tid; c2 < D; c2 += 256) x_s[c2] = __fadd_rn(__ldg(a.tok_emb + (int64_t)tok * D + c2), __ldg(a.pos_emb + (int64_t)p * D + c2));
                }
                const int w = __ldg(a.row_window + row);
                const int T = __ldg(a.win_T + w);
                const int per = (T + HS - 1) / HS, k_begin = min(T, hs * per), k_end = min(T, k_begin + per);
                bar_consumers();
                trace();
#pragma unroll 1
                for (int l = 0; l < L; ++l) {
                    KVT* kcl = reinterpret_cast<KVT*>(a.kc) + (size_t)l * a.Rmax * t_max * D;
                    KVT* vcl = reinterpret_cast<KVT*>(a.vc) + (size_t)l * a.Rmax * t_max * D;
                    // ================= phase 1: x += MLP of the previous layer; self attention of head h (mod.rs:346)
                    if (l > 0) {
                        combine(MODE_SUM, params + ((pl + 1) & 1) * PARAMS + G::P_B2);   // previous layer's block: bias of its MLP2
                        __syncwarp();
                        if (lane == 0) mbar_arrive(pfree + ((pl + 1) & 1));               // previous layer's parameters are dead now
                    }
                    trace();   // [t1] records of the previous phase combined
                    mbar_wait(pfull + (pl & 1), (pl >> 1) & 1);
                    const float* prm = params + (pl & 1) * PARAMS;
                    ++pl;
                    float* ys = y_s + (ph & 1) * SEND;
                    if (hs == 0) {
                        ln6<D>(x_s, prm + G::P_LN1G, prm + G::P_LN1B, prm[G::P_EPS + 0], a.eps_outside, bx);
                        KVT* kd = kcl + ((int64_t)row * t_max + p) * D + h * 64;
                        KVT* vd = vcl + ((int64_t)row * t_max + p) * D + h * 64;
                        c = gemv_epi6<KVT>(P, c, 192, S_D, GemvOut<KVT>{EM_QKV, prm + G::P_BQKV, scale, qkv_s, kd, vd});
                        trace();   // [t2] q | k | v done
                        self_attn6<KVT>(qkv_s, kcl + (int64_t)row * t_max * D + h * 64, vcl + (int64_t)row * t_max * D + h * 64, D, p, wm, wl, wo);
                        attn_merge6(wm, wl, wo, bx, ys + D, 1.0f);
                        trace();   // [t3] self attention done
                        c = gemv_epi6<KVT>(P, c, D, 1, GemvOut<KVT>{EM_PLAIN, nullptr, 1.0f, ys, nullptr, nullptr});
                    } else {
                        for (int c2 = tid; c2 < SEND; c2 += 256) ys[c2] = 0.0f;    // inactive in this phase: weight 0, zeros
                    }
                    trace();   // [t4] out-projection slice done
                    send();
                    trace();   // [t5] sent
                    // ================= phase 2: x += self-attention output; cross attention of head h over this CTA's keys (mod.rs:347)
                    combine(MODE_ATTN, prm + G::P_BO);
                    trace();   // [t6] combined
                    ys = y_s + (ph & 1) * SEND;
                    ln6<D>(x_s, prm + G::P_LN2G, prm + G::P_LN2B, prm[G::P_EPS + 1], a.eps_outside, bx);
                    c = gemv_epi6<KVT>(P, c, 64, S_D, GemvOut<KVT>{EM_CQ, prm + G::P_BCQ, scale, q2_s, nullptr, nullptr});
                    trace();   // [t7] cross query done
                    c.n = cross_attn6<KVT>(P, c.n, slot_cnt, q2_s, k_begin, k_end, wm, wl, wo);
                    attn_merge6(wm, wl, wo, bx, ys + D, (k_end > k_begin) ? 1.0f : 0.0f);
                    trace();   // [t8] cross attention done
                    c = gemv_epi6<KVT>(P, c, D, 1, GemvOut<KVT>{EM_PLAIN, nullptr, 1.0f, ys, nullptr, nullptr});
                    trace();   // [t9]
                    send();
                    trace();   // [t10]
                    // ================= phase 3: x += cross-attention output; MLP slice (mod.rs:348, :376-382)
                    combine(MODE_ATTN, prm + G::P_BCO);
                    trace();   // [t11]
                    ys = y_s + (ph & 1) * SEND;
                    ln6<D>(x_s, prm + G::P_LN3G, prm + G::P_LN3B, prm[G::P_EPS + 2], a.eps_outside, bx);
                    c = gemv_epi6<KVT>(P, c, NS, S_D, GemvOut<KVT>{EM_HID, prm + G::P_B1, 1.0f, hid_s, nullptr, nullptr});
                    for (int c2 = tid; c2 < NS; c2 += 256) bx_store(bx, c2, hid_s[c2]);   // every MMA of the W1 product has completed: the B operand may change
                    trace();   // [t12] hidden slice done
                    c = gemv_epi6<KVT>(P, c, D, S_NS, GemvOut<KVT>{EM_PLAIN, nullptr, 1.0f, ys, nullptr, nullptr});
                    if (tid == 0) { ys[D] = 0.0f; ys[D + 1] = 1.0f; ys[D + 2] = 1.0f; ys[D + 3] = 0.0f; }
                    trace();   // [t13]
                    send();
                    trace();   // [t14]
                }
                // ---- x += MLP of the last layer; rank 0 publishes the row for the vocabulary projection
                combine(MODE_SUM, params + ((pl + 1) & 1) * PARAMS + G::P_B2);
                __syncwarp();
                if (lane == 0) mbar_arrive(pfree + ((pl + 1) & 1));
                if (want_logits && rank == 0)
                    for (int c = tid; c < D; c += 256) a.x[(int64_t)row * D + c] = x_s[c];
                bar_consumers();
            }
            int stop = 0;
            if (want_logits) {
                // ---- vocabulary tiles of this warp.  CTAs of clusters without a row have nothing to do until the rows are published:
                // they take the first LG_NBUF half-tiles into their ring BEFORE the grid barrier (the embedding matrix does not depend on
                // the activations), and their warps get `na` extra tile each (stage A) so that the stream that remains after the
                // barrier is spread evenly (stage B: round robin over all warps).
                const int lg_g = lane >> 2, lg_t = lane & 3;
                constexpr int KH = D / 2, NCH = KH / 32;
                constexpr uint32_t BLKB = 16 * KH * 2;
                constexpr int RINGW = NSLOT * SLOT / NCW;
                static_assert(LG_NBUF * (int)BLKB <= RINGW, "logits ring");
                const __half* Et = reinterpret_cast<const __half*>(a.E_tiled);
                const int v_tiles = (V + 15) / 16;
                const bool idle_cta = cluster_id >= R;
                const int n_idle_w = max(0, n_clusters - R) * CS * NCW;
                const int na = (n_idle_w > 0 && 2 * n_idle_w <= v_tiles) ? 1 : 0;   // = what an idle warp has in its ring when the barrier opens
                const int tiles_a = na * n_idle_w, tiles_b = v_tiles - tiles_a;
                const int iw = ((cluster_id - R) * CS + rank) * NCW + warp;                 // index among the idle warps
                const int my_a = idle_cta ? na : 0;
                const int my_tiles = my_a + (gw < tiles_b ? (tiles_b - gw + n_gw - 1) / n_gw : 0);
                const int total = my_tiles * 2;
                auto tile_of = [&](int i) { return i < my_a ? iw + i * n_idle_w : tiles_a + gw + (i - my_a) * n_gw; };
                uint8_t* wring = ring_mem + (size_t)warp * RINGW;
                uint64_t* wbar = lg_bar + warp * LG_NBUF;
                auto issue = [&](int it) {
                    if (it < total && lane == 0) {
                        const int vt = tile_of(it >> 1);
                        const int slot = (int)((lg_count + (unsigned int)it) % LG_NBUF);
                        mbar_expect_tx(wbar + slot, BLKB);
                        bulk_g2s(wring + (size_t)slot * BLKB, Et + ((int64_t)vt * 2 + (it & 1)) * 16 * KH, BLKB, wbar + slot);
                    }
                };
                if (idle_cta) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < LG_NBUF; ++j) issue(j);
                }
                // ---- the one grid barrier of the step: every row's x is published
                trace();
                if (tid == 0) {
                    ++gen;
                    __threadfence();
                    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(a.bar) : "memory");
                    const unsigned int target = gen * gridDim.x;
                    const long long t0 = clock64(), wd = g_watchdog;
                    while (ld_acquire(a.bar) < target)
                        if (clock64() - t0 > wd) __trap();
                } else {
                    ++gen;
                }
                bar_consumers();
                trace();
                // ================= logits (all CTAs): LN(x) tok_emb^T + mask + online softmax + arg-max (mod.rs:155-156, transcribe.rs:271-276)
                const bool use_mask = a.is_special != nullptr && (a.mask_mode == 1 || (a.mask_mode == 2 && p + 1 <= 5));
                // LayerNorm rows straight into fp16 hi / lo planes in MMA fragment order (decoder5.cu); rows >= R are zero
                for (int r = warp; r < 8 * NT8; r += NCW) {
                    constexpr int NV = D / 128;   // float4 per lane
                    float4 v[NV];
                    if (r < R) {
#pragma unroll
                        for (int i = 0; i < NV; ++i) v[i] = __ldcg(reinterpret_cast<const float4*>(a.x + (int64_t)r * D) + lane + 32 * i);
                        float sum = 0.0f;
#pragma unroll
                        for (int i = 0; i < NV; ++i) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
                        sum = warp_sum(sum);
                        const float mean = __fdiv_rn(sum, (float)D);
                        float q = 0.0f;
#pragma unroll
                        for (int i = 0; i < NV; ++i) {
                            v[i].x = __fsub_rn(v[i].x, mean); v[i].y = __fsub_rn(v[i].y, mean); v[i].z = __fsub_rn(v[i].z, mean); v[i].w = __fsub_rn(v[i].w, mean);
                            q = __fadd_rn(q, __fmul_rn(v[i].x, v[i].x)); q = __fadd_rn(q, __fmul_rn(v[i].y, v[i].y));
                            q = __fadd_rn(q, __fmul_rn(v[i].z, v[i].z)); q = __fadd_rn(q, __fmul_rn(v[i].w, v[i].w));
                        }
                        q = warp_sum(q);
                        const float var = __fdiv_rn(q, (float)D);
                        const float den = a.eps_outside ? __fadd_rn(__fsqrt_rn(var), a.lnf_eps) : __fsqrt_rn(__fadd_rn(var, a.lnf_eps));
#pragma unroll
                        for (int i = 0; i < NV; ++i) {
                            const float4 g4 = __ldg(reinterpret_cast<const float4*>(a.lnf_g) + lane + 32 * i);
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.lnf_b) + lane + 32 * i);
                            float4 o;
                            o.x = __fadd_rn(__fmul_rn(__fdiv_rn(v[i].x, den), g4.x), b4.x);
                            o.y = __fadd_rn(__fmul_rn(__fdiv_rn(v[i].y, den), g4.y), b4.y);
                            o.z = __fadd_rn(__fmul_rn(__fdiv_rn(v[i].z, den), g4.z), b4.z);
                            o.w = __fadd_rn(__fmul_rn(__fdiv_rn(v[i].w, den), g4.w), b4.w);
                            store_frag(pl_hi, pl_lo, D / 32, r, (lane + 32 * i) * 4, o);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < NV; ++i) store_frag(pl_hi, pl_lo, D / 32, r, (lane + 32 * i) * 4, make_float4(0.f, 0.f, 0.f, 0.f));
                    }
                }
                bar_consumers();
                trace();
                {
                    // Swap-AB tensor-core product (decoder4.cu): a warp owns tiles of 16 vocabulary rows (M), 8 batch rows per n-tile
                    // (N), K = D; the matrix is streamed as contiguous half-tiles [16][D/2] (one bulk copy each) through this warp's
                    // share of the ring, LG_NBUF - 1 copies in flight.
                    const int g = lg_g, t = lg_t;
                    float m_run[NT8][2], s_run[NT8][2], bv[NT8][2];
                    int bi[NT8][2];
#pragma unroll
                    for (int j = 0; j < NT8; ++j)
#pragma unroll
                        for (int e = 0; e < 2; ++e) { m_run[j][e] = -INFINITY; s_run[j][e] = 0.0f; bv[j][e] = -INFINITY; bi[j][e] = INT_MAX; }
                    if (!idle_cta) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the ring was last written by bulk copies and read through the generic proxy
#pragma unroll
                        for (int j = 0; j < LG_NBUF; ++j) issue(j);
                    }
                    float ah[NT8][4], al[NT8][4];
#pragma unroll 1
                    for (int it = 0; it < total; ++it) {
                        const unsigned int cnt = lg_count + (unsigned int)it;
                        const int slot = (int)(cnt % LG_NBUF);
                        mbar_wait(wbar + slot, (cnt / LG_NBUF) & 1);
                        const uint8_t* blk = wring + (size_t)slot * BLKB;
                        const int half = it & 1;
                        if (half == 0) {
#pragma unroll
                            for (int j = 0; j < NT8; ++j)
#pragma unroll
                                for (int c = 0; c < 4; ++c) { ah[j][c] = 0.0f; al[j][c] = 0.0f; }
                        }
#pragma unroll
                        for (int c = 0; c < NCH; ++c) {
                            const uint4 a0 = *reinterpret_cast<const uint4*>(blk + (size_t)g * (KH * 2) + c * 64 + t * 16);
                            const uint4 a8 = *reinterpret_cast<const uint4*>(blk + (size_t)(g + 8) * (KH * 2) + c * 64 + t * 16);
                            const int chunk = half * NCH + c;
#pragma unroll
                            for (int j = 0; j < NT8; ++j) {
                                const uint4 bh = pl_hi[(j * (D / 32) + chunk) * 32 + lane];
                                const uint4 bl = pl_lo[(j * (D / 32) + chunk) * 32 + lane];
                                mma16816(ah[j], a0.x, a8.x, a0.y, a8.y, bh.x, bh.y);
                                mma16816(ah[j], a0.z, a8.z, a0.w, a8.w, bh.z, bh.w);
                                mma16816(al[j], a0.x, a8.x, a0.y, a8.y, bl.x, bl.y);
                                mma16816(al[j], a0.z, a8.z, a0.w, a8.w, bl.z, bl.w);
                            }
                        }
                        if (half == 1) {
                            // C fragment: c0,c1 -> (vocabulary row g, batch rows 2t, 2t+1), c2,c3 -> (row g+8, same batch rows)
                            const int n0 = tile_of(it >> 1) * 16;
#pragma unroll
                            for (int j = 0; j < NT8; ++j)
#pragma unroll
                                for (int c = 0; c < 4; ++c) {
                                    const int n = n0 + g + (c >> 1) * 8, e = c & 1;
                                    if (n < V && j * 8 + 2 * t + e < R) {
                                        const float raw = fmaf(al[j][c], 1.0f / 2048.0f, ah[j][c]);
                                        const float v = (use_mask && a.is_special[n]) ? __fadd_rn(raw, -INFINITY) : raw;
                                        if (v > -INFINITY) {
                                            if (v > m_run[j][e]) { s_run[j][e] = s_run[j][e] * expf(m_run[j][e] - v) + 1.0f; m_run[j][e] = v; }
                                            else s_run[j][e] += expf(v - m_run[j][e]);
                                        }
                                        if (v > bv[j][e] || (v == bv[j][e] && n < bi[j][e])) { bv[j][e] = v; bi[j][e] = n; }
                                    }
                                }
                        }
                        __syncwarp();                 // every lane is done with the slot
                        issue(it + LG_NBUF);
                    }
                    lg_count += (unsigned int)total;
                    trace();
                    // merge the 8 lanes that share t (batch rows 2t, 2t+1 of every n-tile), then the 8 warps through shared memory
#pragma unroll
                    for (int j = 0; j < NT8; ++j)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
#pragma unroll
                            for (int off = 4; off < 32; off <<= 1) {
                                const float m2 = __shfl_xor_sync(0xffffffffu, m_run[j][e], off), s2 = __shfl_xor_sync(0xffffffffu, s_run[j][e], off);
                                const float v2 = __shfl_xor_sync(0xffffffffu, bv[j][e], off);
                                const int i2 = __shfl_xor_sync(0xffffffffu, bi[j][e], off);
                                const float mn = fmaxf(m_run[j][e], m2);
                                s_run[j][e] = (m_run[j][e] > -INFINITY ? s_run[j][e] * expf(m_run[j][e] - mn) : 0.0f) + (m2 > -INFINITY ? s2 * expf(m2 - mn) : 0.0f);
                                m_run[j][e] = mn;
                                if (v2 > bv[j][e] || (v2 == bv[j][e] && i2 < bi[j][e])) { bv[j][e] = v2; bi[j][e] = i2; }
                            }
                            if (g == 0) {
                                float* rec = red + (warp * 8 * NT8 + j * 8 + 2 * t + e) * 4;
                                rec[0] = m_run[j][e]; rec[1] = s_run[j][e]; rec[2] = bv[j][e]; rec[3] = __int_as_float(bi[j][e]);
                            }
                        }
                    bar_consumers();
                    if (tid < R) {
                        float M = -INFINITY;
                        for (int w2 = 0; w2 < NCW; ++w2) M = fmaxf(M, red[(w2 * 8 * NT8 + tid) * 4]);
                        float Ssum = 0.0f, best_v = -INFINITY;
                        int best_i = INT_MAX;
                        for (int w2 = 0; w2 < NCW; ++w2) {
                            const float* rec = red + (w2 * 8 * NT8 + tid) * 4;
                            if (rec[0] > -INFINITY) Ssum += rec[1] * expf(rec[0] - M);
                            const int ci = __float_as_int(rec[3]);
                            if (rec[2] > best_v || (rec[2] == best_v && ci < best_i)) { best_v = rec[2]; best_i = ci; }
                        }
                        const int64_t o = (int64_t)blockIdx.x * R + tid;
                        a.lg_m[o] = M;
                        a.lg_s[o] = Ssum;
                        a.lg_v[o] = best_v;
                        a.lg_i[o] = best_i;
                    }
                }
                trace();
                // ================= finish (greedy: beam.rs:9-37 with beam_size 1) by the LAST CTA to deliver its records
                bar_consumers();
                if (tid == 0) {
                    __threadfence();
                    const unsigned int ticket = atomicAdd(a.bar + 1, 1u);
                    ctl[1] = (ticket == gen * gridDim.x - 1) ? 1 : 0;
                }
                bar_consumers();
                if (ctl[1]) {
                    __threadfence();
                    for (int r = warp; r < R; r += NCW) {
                        // <= 160 co-resident CTAs: at most 5 records per lane, every load issued before any use (one L2 round trip)
                        const int NP = gridDim.x;
                        float rm[5], rs[5], rv[5];
                        int ri[5];
#pragma unroll
                        for (int k = 0; k < 5; ++k) {
                            const int c = min(lane + 32 * k, NP - 1);
                            rm[k] = __ldcg(a.lg_m + (int64_t)c * R + r);
                            rs[k] = __ldcg(a.lg_s + (int64_t)c * R + r);
                            rv[k] = __ldcg(a.lg_v + (int64_t)c * R + r);
                            ri[k] = __ldcg(a.lg_i + (int64_t)c * R + r);
                        }
                        float mx = -INFINITY;
#pragma unroll
                        for (int k = 0; k < 5; ++k)
                            if (lane + 32 * k < NP) mx = fmaxf(mx, rm[k]);
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                        float se = 0.0f, bvv = -INFINITY;
                        int bii = INT_MAX;
#pragma unroll
                        for (int k = 0; k < 5; ++k) {
                            if (lane + 32 * k < NP) {
                                if (rm[k] > -INFINITY) se += rs[k] * expf(rm[k] - mx);
                                if (ri[k] != INT_MAX && (rv[k] > bvv || (rv[k] == bvv && ri[k] < bii))) { bvv = rv[k]; bii = ri[k]; }
                            }
                        }
                        se = warp_sum(se);
                        const float lse = logf(se);
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            const float ov = __shfl_xor_sync(0xffffffffu, bvv, o);
                            const int oi = __shfl_xor_sync(0xffffffffu, bii, o);
                            if (ov > bvv || (ov == bvv && oi < bii)) { bvv = ov; bii = oi; }
                        }
                        if (lane == 0) {
                            a.topk_id[r] = bii == INT_MAX ? -1 : bii;
                            a.topk_lp[r] = __fsub_rn(__fsub_rn(bvv, mx), lse);
                            if (!__ldcg(a.finished + r)) {
                                a.tokens[(int64_t)r * t_max + p + 1] = bii;
                                a.lengths[r] = p + 2;
                                if (bii == a.eot) a.finished[r] = 1;
                            }
                        }
                    }
                    bar_consumers();
                    if (tid == 0) {
                        __threadfence();
                        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a.bar + 2), "r"(gen) : "memory");
                    }
                }
                if (tid == 0) {
                    const long long t0 = clock64(), wd = g_watchdog;
                    while (ld_acquire(a.bar + 2) < gen)
                        if (clock64() - t0 > wd) __trap();
                }
                bar_consumers();
                trace();
                int live = 0;
                for (int r = 0; r < R; ++r) live += __ldcg(a.finished + r) ? 0 : 1;
                if (live == 0) {
                    stop = 1;
                    if (blockIdx.x == 0 && tid == 0) { *a.pos = p + 1; *a.n_unfinished = 0; *a.steps_done = step + 1; }
                }
            }
            if (tid == 0) ctl[0] = stop;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic accesses to the aliased buffers before the next step's bulk copies
            bar_all();   // hands the ring back to the producer; it reads the stop flag after this barrier
            if (stop) break;
            if (step + 1 == a.n_steps && blockIdx.x == 0 && tid == 0) {
                *a.pos = a.pos0 + a.n_steps;
                int live = 0;
                for (int r = 0; r < R; ++r) live += __ldcg(a.finished + r) ? 0 : 1;
                *a.n_unfinished = live;
                *a.steps_done = a.n_steps;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cl.sync();   // no CTA leaves while a peer may still address its shared memory
    if (warp == W_MMA) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(P.tmem), "n"(TMEM_COLS) : "memory");
    }
}

template <int D, int HS, int NT8>
constexpr size_t dec6_smem() {
    using G = Geo<D, HS>;
    return 1024 + (size_t)NSLOT * SLOT + (size_t)(G::KMAX / 64) * BX_SLAB +
           sizeof(float) * ((size_t)2 * G::PARAMS + 2 * G::CS * G::SEND + 2 * G::SEND + D + 192 + 64 + G::NS + 16 + 8 + 8 + 512 + 4 + NSLOT) +
           8 * (size_t)(2 * NSLOT + 6 + 5 + NCW * LG_NBUF) + 64;
}

struct LaunchState {
    int clusters = 0;        // 0 unknown, > 0 co-resident clusters to launch, -1 unsupported
    bool cooperative = true;
};
std::mutex g_mu;

template <int D, int HS, int NT8, typename KVT>
bool launch6_t(const Dec3Args& a, cudaStream_t st) {
    using G = Geo<D, HS>;
    static_assert((size_t)2 * NT8 * (D / 32) * 32 * 16 + (size_t)NCW * 8 * NT8 * 16 <= sizeof(float) * (2 * G::PARAMS + 2 * G::CS * G::SEND), "logits scratch must fit the aliased buffers");
    auto k = dec6_kernel<D, HS, NT8, KVT>;
    const size_t smem = dec6_smem<D, HS, NT8>();
    static LaunchState state[16];   // per device ordinal
    static bool wd_set[16] = {};
    int dev = 0;
    WB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 16) return false;
    std::lock_guard<std::mutex> lock(g_mu);
    LaunchState& S = state[dev];
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(NTH6);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = G::CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeCooperative;
    attr[1].val.cooperative = 1;
    cfg.attrs = attr;
    if (S.clusters == 0) {
        if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            (G::CS > 8 && cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)) {
            cudaGetLastError();
            S.clusters = -1;
            return false;
        }
        int n_clusters = 0;
        cfg.gridDim = dim3(G::CS);
        cfg.numAttrs = 1;
        const cudaError_t oe = cudaOccupancyMaxActiveClusters(&n_clusters, k, &cfg);
        if (getenv("WB200_VERBOSE")) fprintf(stderr, "[wb] dec6<D=%d,HS=%d,NT8=%d>: smem %zu B, max active clusters %d (%s)\n", D, HS, NT8, smem, n_clusters, cudaGetErrorString(oe));
        if (oe != cudaSuccess || n_clusters < 1) {
            cudaGetLastError();
            S.clusters = -1;
            return false;
        }
        S.clusters = n_clusters;
    }
    if (S.clusters < 0) return false;
    if (!wd_set[dev]) {
        if (const char* e = getenv("WB200_WATCHDOG_MS")) {   // 0 = never trap
            const long long ms = atoll(e);
            const long long cyc = ms <= 0 ? (1LL << 62) : ms * 2000000LL;
            WB_CUDA(cudaMemcpyToSymbol(g_watchdog, &cyc, sizeof(cyc)));
        }
        wd_set[dev] = true;
    }
    // every launched cluster must be co-resident (grid barrier): launch what the device holds; rows beyond that are looped over
    const int n_cl = S.clusters;
    cfg.gridDim = dim3(n_cl * G::CS);
    cudaError_t e = cudaErrorUnknown;
    if (S.cooperative && getenv("WB200_NO_COOP")) S.cooperative = false;   // profilers cannot replay cooperative cluster launches
    if (S.cooperative) {
        cfg.numAttrs = 2;
        e = cudaLaunchKernelEx(&cfg, k, a);
        if (e != cudaSuccess) {   // cooperative + cluster rejected by this driver: plain cluster launch (co-residency from the occupancy query)
            cudaGetLastError();
            S.cooperative = false;
            if (getenv("WB200_VERBOSE")) fprintf(stderr, "[wb] dec6: cooperative cluster launch rejected (%s), using a plain cluster launch\n", cudaGetErrorString(e));
        }
    }
    if (!S.cooperative) {
        cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, k, a);
    }
    WB_CUDA(e);
    WB_LAUNCH_CHECK();
    return true;
}

}  // namespace

// ---- host: packed weights / parameter blocks of a model (built once per session) ----------------------------------------------
template <int D, int HS>
static void build_pack_t(const std::vector<Dec6LayerSrc>& layers, DevBuf<uint8_t>& pack, DevBuf<float>& params, cudaStream_t st) {
    using G = Geo<D, HS>;
    const int L = (int)layers.size();
    pack.alloc((size_t)L * G::CS * G::PACK);
    params.alloc((size_t)L * G::CS * G::PARAMS);
    std::vector<PackSeg> ps;
    std::vector<ParamSeg> qs;
    std::vector<float> eps_h((size_t)L * 4, 0.0f);
    for (int l = 0; l < L; ++l) { eps_h[(size_t)l * 4] = layers[(size_t)l].ln1_eps; eps_h[(size_t)l * 4 + 1] = layers[(size_t)l].ln2_eps; eps_h[(size_t)l * 4 + 2] = layers[(size_t)l].ln3_eps; }
    DevBuf<float> eps_d;
    eps_d.alloc(eps_h.size());
    WB_CUDA(cudaMemcpyAsync(eps_d.p, eps_h.data(), eps_h.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    for (int l = 0; l < L; ++l) {
        const Dec6LayerSrc& S = layers[(size_t)l];
        for (int r = 0; r < G::CS; ++r) {
            const int h = r % G::H;
            const int64_t base = ((int64_t)l * G::CS + r) * G::PACK;
            const int64_t pb = ((int64_t)l * G::CS + r) * G::PARAMS;
            // q_h | k_h | v_h: three 64-row pieces of the fused [3d][d] matrix, d rows apart
            ps.push_back(PackSeg{S.Wqkv, D, h * 64, 0, 192, D, 64, D, base + G::OFF_QKV});
            ps.push_back(PackSeg{S.Wo, D, 0, h * 64, D, 64, D, 0, base + G::OFF_O});
            ps.push_back(PackSeg{S.Wcq, D, h * 64, 0, 64, D, 64, 0, base + G::OFF_CQ});
            ps.push_back(PackSeg{S.Wco, D, 0, h * 64, D, 64, D, 0, base + G::OFF_CO});
            ps.push_back(PackSeg{S.W1, D, r * G::NS, 0, G::NS, D, G::NS, 0, base + G::OFF_W1});
            ps.push_back(PackSeg{S.W2, 4 * D, 0, r * G::NS, D, G::NS, D, 0, base + G::OFF_W2});
            qs.push_back(ParamSeg{S.ln1_g, D, pb + G::P_LN1G}); qs.push_back(ParamSeg{S.ln1_b, D, pb + G::P_LN1B});
            qs.push_back(ParamSeg{S.ln2_g, D, pb + G::P_LN2G}); qs.push_back(ParamSeg{S.ln2_b, D, pb + G::P_LN2B});
            qs.push_back(ParamSeg{S.ln3_g, D, pb + G::P_LN3G}); qs.push_back(ParamSeg{S.ln3_b, D, pb + G::P_LN3B});
            qs.push_back(ParamSeg{S.bo, D, pb + G::P_BO}); qs.push_back(ParamSeg{S.bco, D, pb + G::P_BCO}); qs.push_back(ParamSeg{S.b2, D, pb + G::P_B2});
            for (int part = 0; part < 3; ++part) qs.push_back(ParamSeg{S.bqkv + part * D + h * 64, 64, pb + G::P_BQKV + part * 64});
            qs.push_back(ParamSeg{S.bcq + h * 64, 64, pb + G::P_BCQ});
            qs.push_back(ParamSeg{S.b1 + r * G::NS, G::NS, pb + G::P_B1});
            qs.push_back(ParamSeg{eps_d.p + (size_t)l * 4, 4, pb + G::P_EPS});
        }
    }
    DevBuf<PackSeg> dps;
    DevBuf<ParamSeg> dqs;
    dps.alloc(ps.size());
    dqs.alloc(qs.size());
    WB_CUDA(cudaMemcpyAsync(dps.p, ps.data(), ps.size() * sizeof(PackSeg), cudaMemcpyHostToDevice, st));
    WB_CUDA(cudaMemcpyAsync(dqs.p, qs.data(), qs.size() * sizeof(ParamSeg), cudaMemcpyHostToDevice, st));
    dec6_pack_kernel<<<dim3(32, (unsigned)std::min<size_t>(ps.size(), 1024)), 256, 0, st>>>(dps.p, (int)ps.size(), pack.p);
    WB_LAUNCH_CHECK();
    dec6_param_kernel<<<(unsigned)std::min<size_t>(qs.size(), 2048), 128, 0, st>>>(dqs.p, (int)qs.size(), params.p);
    WB_LAUNCH_CHECK();
    WB_CUDA(cudaStreamSynchronize(st));   // the descriptor arrays go out of scope
}

bool dec6_supported(int d, int H) { return (d == 128 || d == 384) && H * 64 == d; }

int dec6_pick_hs(int d, int R) {
    const char* e = getenv("WB200_DEC6_HS");
    if (e && (e[0] == '1' || e[0] == '2')) return e[0] - '0';
    (void)d;
    (void)R;
    return 1;   // one CTA per head; WB200_DEC6_HS=2 selects the two-CTAs-per-head shape (keys and MLP slices split) for <= 8 rows:
                // measured 150 vs 178 us per position at 3 rows, both behind decoder4.cu's 131 us, which therefore keeps <= 7 rows
}

void dec6_build_pack(int d, int hs, const std::vector<Dec6LayerSrc>& layers, DevBuf<uint8_t>& pack, DevBuf<float>& params, cudaStream_t st) {
    if (d == 384 && hs == 1) build_pack_t<384, 1>(layers, pack, params, st);
    else if (d == 384 && hs == 2) build_pack_t<384, 2>(layers, pack, params, st);
    else if (d == 128 && hs == 1) build_pack_t<128, 1>(layers, pack, params, st);
    else if (d == 128 && hs == 2) build_pack_t<128, 2>(layers, pack, params, st);
    else fail(WB_ERR_UNSUPPORTED, "dec6: unsupported width");
}

// Returns false when this configuration is not covered (caller falls back to decoder5.cu / decoder3.cu).
bool launch_dec6(const Dec3Args& a, int hs, bool w_half, cudaStream_t st) {
    if (!w_half || a.R > 24 || a.R < 1 || a.k != 1 || !a.greedy || a.use_cur_tok || a.anc != nullptr || a.logits_out != nullptr) return false;
    if (a.H * 64 != a.d || a.E_tiled == nullptr || a.d6_pack == nullptr || a.d6_params == nullptr || !a.ckv_hm || a.t_max > 128) return false;
    if (hs == 2 && a.R > 8) return false;
#define WB_D6(DD, HS_, NT8_) (a.kv_half ? launch6_t<DD, HS_, NT8_, __half>(a, st) : launch6_t<DD, HS_, NT8_, float>(a, st))
    if (a.d == 384) {
        if (hs == 2) return WB_D6(384, 2, 1);
        return a.R <= 8 ? WB_D6(384, 1, 1) : WB_D6(384, 1, 3);
    }
    if (a.d == 128) {
        if (hs == 2) return WB_D6(128, 2, 1);
        return a.R <= 8 ? WB_D6(128, 1, 1) : WB_D6(128, 1, 3);
    }
#undef WB_D6
    return false;
}

}  // namespace wb

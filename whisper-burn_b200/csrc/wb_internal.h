// Internal declarations shared by the translation units of libwhisper_b200.so.
// Nothing here is part of the C ABI (include/whisper_b200.h).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/whisper_b200.h"

namespace wb {

// ---- errors -----------------------------------------------------------------------------
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
void set_last_error(const std::string& m);
[[noreturn]] void fail(int code, const std::string& m);

#define WB_CUDA(expr)                                                                           \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            ::wb::fail(_e == cudaErrorMemoryAllocation ? WB_ERR_OOM : WB_ERR_CUDA,              \
                       std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                           std::to_string(__LINE__) + ")");                                     \
    } while (0)

#define WB_REQUIRE(cond, msg)                                   \
    do {                                                        \
        if (!(cond)) ::wb::fail(WB_ERR_INVALID_ARG, (msg));     \
    } while (0)

// kernels launched by this library (bench.py's gpu_launches)
extern thread_local int64_t g_launch_count;
#define WB_LAUNCH_CHECK()                     \
    do {                                      \
        ++::wb::g_launch_count;               \
        WB_CUDA(cudaPeekAtLastError());       \
    } while (0)

// ---- per-device launch configuration cache -----------------------------------------------------
// cudaFuncSetAttribute / occupancy results belong to a DEVICE: one slot per device ordinal, guarded by a mutex
// (a process may create models on several devices and sessions from several host threads).
struct PerDeviceConfig {
    std::mutex mu;
    size_t value[16] = {};
    // runs configure() when the current device has not been configured for `want` yet; returns what configure() returned
    template <typename F>
    bool ensure(size_t want, F&& configure) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return configure();
        std::lock_guard<std::mutex> lock(mu);
        if (value[dev] == want) return true;
        if (!configure()) return false;
        value[dev] = want;
        return true;
    }
};

// ---- device buffer ----------------------------------------------------------------------
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    void alloc(size_t count) {
        release();
        if (count == 0) return;
        WB_CUDA(cudaMalloc((void**)&p, count * sizeof(T)));
        n = count;
    }
    void ensure(size_t count) {
        if (count > n) alloc(count);
    }
    void zero(cudaStream_t st) {
        if (p) WB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), st));
    }
};

// ---- constant tables of the frontend (host-built, f32 op order of audio.rs) ----------------
constexpr int N_FFT = 400, HOP = 160, N_MELS = 80, N_FREQ = 201;
constexpr int KPAD = 208;                 // 201 padded to a multiple of 8
struct FrontendTables {
    std::vector<float> hann;              // [400]            audio.rs:272-278
    std::vector<float> basis_t;           // [400][2*KPAD]    audio.rs:349-364, transposed: [j][cos k.. | sin k..]
    std::vector<float> mel_filt;          // [80][201]        audio.rs:67-143
    int mel_lo[N_MELS], mel_hi[N_MELS];   // non-zero tap range of every filter
};
const FrontendTables& frontend_tables();

// ---- model ----------------------------------------------------------------------------------
struct LinearW {       // y = x @ W + b, stored transposed: w[n][k] (k contiguous)
    float* w32 = nullptr;   // fp32 [N][K]   (encoder GEMMs, and decoder when !fp16_exact)
    __half* w16 = nullptr;  // fp16 [N][K]   (decoder GEMVs when fp16_exact)
    float* b = nullptr;     // [N] (never null; zeros when the reference has no bias)
    int n = 0, k = 0;
};
struct LayerNormW {
    float* g = nullptr;
    float* b = nullptr;
    float eps = 1e-5f;
};
struct EncBlockW {
    LayerNormW attn_ln, mlp_ln;
    LinearW qkv;        // [3d][d]  rows: q | k | v
    LinearW out, mlp1, mlp2;
};
struct DecBlockW {
    LayerNormW attn_ln, cross_ln, mlp_ln;
    LinearW qkv;        // self-attention q | k | v
    LinearW out;
    LinearW cq;         // cross query
    LinearW ckv;        // cross k | v  [2d][d]  (applied to the encoder output once per window)
    LinearW cout, mlp1, mlp2;
};

struct Model {
    wb_dims dims{};
    int device = 0;
    bool finalized = false;
    bool fp16_exact = false;
    int ln_eps_outside = 1;     // burn 0.9: (x-mean)/(sqrt(var)+eps)
    std::map<std::string, std::pair<std::vector<int64_t>, std::vector<float>>> host;  // until finalize

    std::vector<void*> allocs;  // everything cudaMalloc'ed for weights
    // frontend tables
    float* basis_t = nullptr;   // [400][416]
    float* mel_filt = nullptr;  // [80][201]
    int* mel_range = nullptr;   // [80][2]
    // encoder
    LinearW conv1;              // [d][3*80]   k = kk*80 + c
    LinearW conv2;              // [d][3*d]    k = kk*d + c
    float* enc_pos = nullptr;   // [n_audio_ctx][d]
    std::vector<EncBlockW> enc;
    LayerNormW ln_post;
    // decoder
    float* tok_emb32 = nullptr;   // [V][d] (always kept: embedding lookup + fp32 logits path)
    __half* tok_emb16 = nullptr;  // [V][d] when fp16_exact
    __half* tok_emb16_tiled = nullptr;   // [ceil(V/16)][2][16][d/2] half-tiles for decoder4.cu (d = 128 / 384 only)
    float* dec_pos = nullptr;     // [n_text_ctx][d]
    std::vector<DecBlockW> dec;
    LayerNormW dec_ln;
    float* zero_bias = nullptr;   // [max(4d, V)] zeros

    ~Model();
};

// ---- GEMM (gemm.cu) -------------------------------------------------------------------------
// C[g][m][n] = epilogue( sum_k A[g][m][k] * B[n][k] ), fp32.  Rows are grouped (one group per
// audio window) so that A rows may overlap (lda < K: the conv stems read a sliding 3-row
// window of a token-major buffer); plain GEMMs use one group.
struct GemmGroup {
    int64_t a_off;   // element offset of the group's first A row
    int64_t c_off;   // element offset of the group's first C row (also residual)
    int rows;
};
enum { ACT_NONE = 0, ACT_GELU = 1 };
struct GemmParams {
    const float* A = nullptr;
    int64_t lda = 0;
    const float* B = nullptr;      // [N][K]
    float* C = nullptr;
    __half* C16 = nullptr;         // optional: store fp16-rounded results here instead of C (fp16 K/V cache)
    int64_t ldc = 0;
    int N = 0, K = 0;
    const float* bias = nullptr;   // [N] or null
    int act = ACT_NONE;
    float scale = 1.0f;            // applied to columns < scale_cols after bias (q/k pre-scaling)
    int scale_cols = 0;
    const float* residual = nullptr;  // same layout as C, or null
    const float* pos = nullptr;       // [rows][N] added after activation, indexed by the row inside the group
    const GemmGroup* groups = nullptr;   // device array
    int n_groups = 1;
    int max_rows = 0;                 // max rows over groups
};
void launch_gemm(const GemmParams& p, cudaStream_t st);

// ---- tensor-core GEMM on fp16 hi/lo planes (gemm_f16.cu) ----------------------------------------------------
// A = A_hi + A_lo / 2048 (fp16 planes, row-major [rows][lda]), B = fp16 [N][K] (exact weights); see gemm_f16.cu.
struct GemmF16Params {
    const __half *A_hi = nullptr, *A_lo = nullptr;
    int64_t lda = 0;
    const __half* B = nullptr;        // [N][K]
    float* C = nullptr;               // fp32 result rows, or null
    __half *P_hi = nullptr, *P_lo = nullptr;   // fp16 planes of the result, or null
    int64_t ldc = 0;
    int N = 0, K = 0;
    const float* bias = nullptr;
    int act = ACT_NONE;
    float scale = 1.0f;
    int scale_cols = 0;
    const float* residual = nullptr;
    const float* pos = nullptr;
    const GemmGroup* groups = nullptr;
    int n_groups = 1;
    int max_rows = 0;
};
bool gemm_f16_supported(const GemmF16Params& p);
// Tensor maps and launch geometry of one GEMM site, built once per (site, geometry) instead of per launch
class GemmF16Plan {
public:
    GemmF16Plan();
    ~GemmF16Plan();
    GemmF16Plan(const GemmF16Plan&) = delete;
    GemmF16Plan& operator=(const GemmF16Plan&) = delete;
    void build(const GemmF16Params& p, int64_t a_group_stride, int a_rows_total_per_group);
    void launch(cudaStream_t st) const;
    bool matches(int max_rows, int n_groups) const { return impl != nullptr && key_rows == max_rows && key_groups == n_groups; }
private:
    struct Impl;
    Impl* impl = nullptr;
    int key_rows = -1, key_groups = -1;
};
void launch_split_f16(const float* src, __half* hi, __half* lo, int64_t n, cudaStream_t st);

// ---- frontend (logmel.cu) ---------------------------------------------------------------------
struct LogMelWindow {
    int64_t wave_off;   // element offset into the wave buffer
    int n_samples;
    int n_frames;       // n_samples / 160 frames enter the max of audio.rs:50
    int n_store;        // frames written out (<= n_frames; mels_to_text clips, transcribe.rs:171-173)
    int max_slot;       // windows sharing a slot share the global max of audio.rs:50
    int64_t out_off;    // element offset of frame 0's row in the token-major output (row = 80 floats)
};
// raw pass: log10(max(mel,1e-10)) token-major + per-slot max; finalize: clamp/normalise in place.
void launch_logmel(const Model& m, const float* wave, const LogMelWindow* win_dev, int n_windows,
                   int max_frames, float* mel_rows, int* max_slots, int n_slots, cudaStream_t st);
// token-major rows [n][80] -> channel-major [80][n] (the reference's [B,80,F] layout) and back
void launch_rows_to_chan(const float* rows, float* chan, int n_frames, cudaStream_t st);
void launch_chan_to_rows(const float* chan, float* rows, int n_frames, int64_t chan_stride, cudaStream_t st);

// ---- encoder pieces (encoder.cu) -----------------------------------------------------------------
// head-major re-layout of one layer's cross K|V rows (see encoder.cu)
void launch_ckv_relayout(const float* src, void* dst, bool dst_half, const int64_t* win_row_off, const int* win_T, int n_windows,
                         int64_t M, int d, cudaStream_t st);
void launch_layernorm(const float* x, float* y, const LayerNormW& ln, int rows, int d, int eps_outside, cudaStream_t st);
struct AttnWindow {
    int64_t row_off;   // first packed row of the window
    int T;
};
// non-causal multi-head attention over packed rows; qkv [rows][3d] (q,k pre-scaled), out [rows][d]
void launch_encoder_attention(const float* qkv, float* out, const AttnWindow* win_dev, int n_windows, int max_T, int d, int n_head,
                              cudaStream_t st);
// tensor-core version on fp16 hi/lo planes (enc_attn_tc.cu): q | k | v planes [rows][3d] -> output planes [rows][d]
void launch_encoder_attention_tc(const __half* qkv_hi, const __half* qkv_lo, __half* out_hi, __half* out_lo, const AttnWindow* win_dev,
                                 int n_windows, int max_T, int d, int n_head, cudaStream_t st);
// LayerNorm whose output leaves as fp16 hi/lo planes (and optionally as fp32 rows too: y may be null)
void launch_layernorm_f16(const float* x, float* y, __half* y_hi, __half* y_lo, const LayerNormW& ln, int rows, int d, int eps_outside,
                          cudaStream_t st);

}  // namespace wb

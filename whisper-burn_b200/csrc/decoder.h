// Argument blocks of the persistent decoder kernels.  Internal.
#pragma once
#include <stdint.h>

#include "wb_internal.h"

namespace wb {

constexpr int DEC_KC = 8;   // top candidates a persistent decoder keeps per record (k <= 7)

// ---- persistent cooperative decoder (decoder3.cu) ------------------------------------------------------
struct Dec3Layer {
    const float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *ln3_g, *ln3_b;
    float ln1_eps, ln2_eps, ln3_eps;
    const void *Wqkv, *Wo, *Wcq, *Wco, *W1, *W2;     // [N][K] fp16 or fp32
    const float *bqkv, *bo, *bcq, *bco, *b1, *b2;
};
// decoder5.cu stage descriptors, built on the host: entry [l * 16 + slot] for the stage slots of layer l (LN1, QKV, SELF, OUT, LN2,
// CQ, CROSS, COUT, LN3, MLP1, MLP2 = 0..10), entries [L * 16 + 11] / [L * 16 + 12] for the final LayerNorm and the logits.
enum { D5_KIND_LN = 0, D5_KIND_ATTN, D5_KIND_GEMM };
enum { D5_ST_LN_EMB = 0, D5_ST_LN_FOLD, D5_ST_LN_FOLD_NOPUB, D5_ST_LN_X, D5_ST_PLANES, D5_ST_CROSS };
enum { D5_EM_QKV = 0, D5_EM_RESID, D5_EM_CQ, D5_EM_HID, D5_EM_PART, D5_EM_LOGITS };
struct Dec5Desc {
    int kind = D5_KIND_ATTN;
    const void* W = nullptr;       // [N][n_slabs * d] fp16
    const float* bias = nullptr;
    int N = 0, n_slabs = 1, stage = 0, emit = 0;
    const float *g = nullptr, *b = nullptr;   // LayerNorm parameters (D5_KIND_LN)
    float eps = 0.0f;
    int src = 0;                   // input planes of a linear stage: 1 = attention output, 2 = MLP hidden, 3 = LayerNorm output
    int ks = 0;                    // K slab width of a linear stage (= columns staged per item); d except MLP2 when 4d splits into 3
    int n_fold = 0;                // LayerNorm stage: K-slab partial sums (ypart) folded into x first (the producing stage's n_slabs)
    int pad_[2] = {0, 0};          // sizeof % 16 == 0: the table is copied to shared memory in 16-byte words
};
static_assert(sizeof(Dec5Desc) % 16 == 0, "Dec5Desc must be a whole number of 16-byte words");
struct Dec3Args {
    int R = 0, Rmax = 0, d = 0, H = 0, L = 0, V = 0, t_max = 0;
    int64_t Mcap = 0;
    int eps_outside = 1;
    float qk_scale = 1.0f;
    const Dec3Layer* layers = nullptr;    // device array [L]
    const float* tok_emb = nullptr;       // fp32 [V][d] (embedding lookup)
    const float* pos_emb = nullptr;
    const void* E = nullptr;              // logits matrix [V][d] fp16 or fp32
    const void* E_tiled = nullptr;        // decoder4.cu: the same matrix as contiguous half-tiles [ceil(V/16)][2][16][d/2] fp16
    const float *lnf_g = nullptr, *lnf_b = nullptr;
    float lnf_eps = 1e-5f;
    // state
    float *x = nullptr, *q = nullptr, *att = nullptr, *hid = nullptr;
    float* ypart = nullptr;                 // decoder5.cu: MLP2 partial sums [n_slabs][R][d], folded by the next LayerNorm stage
    const Dec5Desc* d5 = nullptr;           // decoder5.cu: device array [L * 8 + 1]
    void *att_pl = nullptr, *hid_pl = nullptr;   // decoder5.cu: fragment-order fp16 hi/lo planes of the attention output / MLP hidden layer
    float* lgbuf = nullptr;                 // decoder5.cu: [R][V] logits scratch (== logits_out when that is requested)
    int lg_slices = 1;                      // decoder5.cu: vocabulary slices per row in the softmax / candidate stage
    void *kc = nullptr, *vc = nullptr;    // [L][Rmax][t_max][d]  fp32 or fp16 (kv_half)
    const void* ckv = nullptr;            // [L][Mcap][2d]
    int ckv_hm = 0;                       // 1: head-major cross K/V (encoder.cu ckv_relayout_kernel), 0: GEMM row-major order
    int kv_half = 0;
    int kv_row0 = 0;                      // decoder5.cu row groups: local row r of this launch is cache row r + kv_row0 (ancestry entries are absolute)
    const int* row_window = nullptr;
    const int64_t* win_row_off = nullptr;
    const int* win_T = nullptr;
    const int* anc = nullptr;
    int n_splits = 1;
    float *part_o = nullptr, *part_m = nullptr, *part_l = nullptr;   // [R][H][S][64], [R][H][S]
    // tokens / control
    int* tokens = nullptr;                // [Rmax][t_max]
    const int* cur_tok = nullptr;
    int use_cur_tok = 0;
    int pos0 = 0, n_steps = 1, logits_from = 0;
    const uint8_t* is_special = nullptr;
    int mask_mode = 0;
    int k = 1, greedy = 0, eot = -1;
    int *lengths = nullptr, *finished = nullptr;
    int *topk_id = nullptr;
    float* topk_lp = nullptr;
    float* logits_out = nullptr;
    float *lg_m = nullptr, *lg_s = nullptr, *lg_v = nullptr;
    int* lg_i = nullptr;
    int *pos = nullptr, *n_unfinished = nullptr, *steps_done = nullptr;
    unsigned int* bar = nullptr;          // [4] arrival count, generation / ticket, finish flag (decoder6.cu)
    const void* d6_pack = nullptr;        // decoder6.cu: packed weight slices [L][CS][PACK] bytes
    const float* d6_params = nullptr;     // decoder6.cu: parameter blocks [L][CS][PARAMS] floats
    unsigned long long* trace = nullptr;  // optional: stage / barrier timestamps of CTA 0 (ns)
    int trace_cap = 0;
};
void launch_dec3(const Dec3Args& a, int n_ctas, bool w_half, cudaStream_t st);
// cluster / DSMEM version (decoder4.cu); returns false when the configuration is not covered
bool launch_dec4(const Dec3Args& a, bool w_half, cudaStream_t st);
// batched tensor-core version (decoder5.cu); returns false when the configuration is not covered
bool launch_dec5(const Dec3Args& a, int n_ctas, bool w_half, cudaStream_t st);
size_t dec5_plane_uint4(int d);   // uint4 elements of one global activation plane
// head-fused cluster decoder (decoder6.cu): greedy, d in {128, 384}, <= 24 rows; hs = CTAs per attention head (1 or 2)
struct Dec6LayerSrc {   // device pointers of one decoder block (fp16 [N][K] weights, fp32 biases / LayerNorm parameters)
    const __half *Wqkv, *Wo, *Wcq, *Wco, *W1, *W2;
    const float *bqkv, *bo, *bcq, *bco, *b1, *b2;
    const float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *ln3_g, *ln3_b;
    float ln1_eps, ln2_eps, ln3_eps;
};
bool dec6_supported(int d, int H);
int dec6_pick_hs(int d, int R);
void dec6_build_pack(int d, int hs, const std::vector<Dec6LayerSrc>& layers, DevBuf<uint8_t>& pack, DevBuf<float>& params, cudaStream_t st);
bool launch_dec6(const Dec3Args& a, int hs, bool w_half, cudaStream_t st);

void launch_dec_anc_identity(int* anc, int R, int t_max, cudaStream_t st);
void launch_dec_reorder(const int* anc_old, int* anc_new, const int* parent, const int* pos_ptr, int R, int t_max,
                        cudaStream_t st);

}  // namespace wb

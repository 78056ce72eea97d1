// Decoding session: all device state between prep_audio and the emitted token ids.  Internal.
#pragma once
#include <memory>
#include <vector>

#include "decoder.h"
#include "wb_internal.h"

namespace wb {

constexpr int MEL_PADDING = 10;   // transcribe.rs:33

struct Session {
    Model* m = nullptr;
    cudaStream_t st = nullptr;
    int max_windows = 0, max_beams = 0, t_max = 0, kv_dtype = WB_KV_F32;
    int Rmax = 0;        // max_windows * max_beams decode rows
    int TmS = 0;         // rows per window in the token-major mel / conv1 buffers (n_audio_ctx + 2 halo rows)
    int Tcap = 0;        // max encoder positions per window
    int64_t Mcap = 0;    // max packed encoder rows
    int kmax = 8;        // candidates per row the step API can return (k <= 7)

    // ---- geometry of the windows currently encoded (host mirrors)
    int n_windows = 0;
    std::vector<int> win_F, win_Tm, win_T;
    std::vector<int64_t> win_row_off;
    int64_t M_tot = 0;
    int max_T = 0, max_Tm = 0;
    bool encoded = false;

    // ---- device: descriptors
    DevBuf<LogMelWindow> d_lmwin;
    DevBuf<GemmGroup> d_g1, d_g2;
    DevBuf<AttnWindow> d_awin;
    DevBuf<int64_t> d_win_row_off;
    DevBuf<int> d_win_T;
    // ---- device: frontend + encoder activations
    DevBuf<float> wave;
    DevBuf<int> max_slots;
    DevBuf<float> mel_rows, x, xa;           // token-major log-mel rows, residual stream, encoder output
    DevBuf<float> h1, xn, att, qkv, hid;     // fp32 activations of the CUDA-core encoder (weights that are not fp16-exact)
    // tensor-core path (fp16-exact weights): every GEMM input travels as a pair of fp16 planes (gemm_f16.cu)
    DevBuf<__half> mel_h, mel_l, h1_h, h1_l, xn_h, xn_l, qkv_h, qkv_l, att_h, att_l, hid_h, hid_l, xa_h, xa_l;
    std::vector<std::unique_ptr<GemmF16Plan>> enc_plans;   // conv1, conv2, per encoder layer qkv / out / mlp1 / mlp2, per decoder layer cross K|V
    bool conv_tc_ok = true;            // cleared if the driver rejects the overlapping-row tensor maps of the conv stems
    bool use_tc = true;                // WB200_GEMM=simt forces the fp32 CUDA-core GEMM
    void run_encoder_f16();            // tensor-core encoder
    void run_encoder_f32();            // fp32 CUDA-core encoder (weights that are not fp16-exact)
    DevBuf<float> ckv;     // [L][Mcap][2d]  cross keys (scaled) | values, projected once per window
    DevBuf<float> ckv_tmp; // [Mcap][2d] one layer's projection in GEMM (row-major) order, before the head-major re-layout
    bool ckv_hm = true;    // cross K/V stored head-major (what the persistent decoders stream)
    // ---- device: decode state
    DevBuf<float> kc, vc;  // [L][Rmax][t_max][d] self keys (scaled) / values
    DevBuf<__half> kc16, vc16, ckv16;   // fp16 caches (WB_KV_F16)
    DevBuf<float> dx, dq, dhid, logits;
    DevBuf<Dec5Desc> d5_desc;         // decoder5.cu stage descriptors, d x d projections unsplit
    DevBuf<Dec5Desc> d5_desc_split;   // the same with the d x d projections as K slabs (launches with unsplit cross attention); may be empty
    DevBuf<uint4> att_pl, hid_pl;   // decoder5.cu activation planes
    DevBuf<float> part_o, part_m, part_l;
    DevBuf<int> tokens, lengths, cur_tok, finished, row_window, anc0, anc1, parent, pos, n_unfinished, topk_id;
    DevBuf<float> topk_lp;
    DevBuf<uint8_t> is_special;
    bool have_special = false;
    int anc_cur = 0;       // which ancestry table is current
    bool anc_identity = true;
    int R = 0;             // live rows
    int host_pos = 0;      // host mirror of *pos
    // pinned host staging
    int* h_int = nullptr;      // [4 * Rmax + 16]
    float* h_float = nullptr;  // [Rmax * kmax]
    // timings of the last transcribe call
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    float last_ms[4] = {0, 0, 0, 0};
    int64_t last_steps = 0;
    int last_groups = 1;         // row groups (launches) of the last decode
    int last_decoder = 0;        // which persistent decoder the last launch used (6, 5, 4 or 3); 0 = none yet
    int dec_version = 4;         // 4 = best persistent decoder for the shape (decoder6 / decoder5 / decoder4), 3 = force the grid-barrier fallback (WB200_DECODER=3)
    int n_sm = 0;
    DevBuf<Dec3Layer> d3_layers;
    DevBuf<uint8_t> d6_pack[3];     // decoder6.cu packed weight slices, index = CTAs per head (1, 2), built on first use
    DevBuf<float> d6_params[3];
    bool use_dec6 = true;           // WB200_DEC6=0 disables the head-fused cluster decoder
    bool force_dec6 = false;        // WB200_DEC6=force: also for <= 7 rows (decoder4.cu's range)
    DevBuf<unsigned int> d3_bar;
    DevBuf<int> steps_done;
    DevBuf<float> datt;
    DevBuf<unsigned long long> d3_trace;   // debug: WB200_TRACE=1
    void launch_v3(int R_, int pos0, int n_steps, int logits_from, bool use_cur_tok, int mask_mode, int k, bool greedy,
                   int eot);
    bool full_logits = false;    // also write raw logits [R][V] (stateless forward_decoder)
    int n_logit_ctas = 0;
    DevBuf<float> ypart, lg_m, lg_s, lg_v;
    DevBuf<int> lg_i;
    std::vector<cudaEvent_t> prof_ev;   // wb_session_profile_decode: launch begin / end
    void profile_decode(const int64_t* prompt, int64_t prompt_len, int n_steps, int64_t eot, float* logits_ms, float* step_ms);

    Session(Model* model, int64_t max_windows, int64_t max_beams, int64_t max_text_len, int kv_dtype);
    ~Session();
    Session(const Session&) = delete;
    Session& operator=(const Session&) = delete;

    // waveforms already in `wave` (device) at offsets[i], lens[i] samples each
    void encode_from_device_wave(const float* wave_dev, const int64_t* offsets, const int64_t* lens, int64_t n);
    void encode_waveforms_host(const float* const* waves, const int64_t* lens, int64_t n);
    void encode_mels_host(const float* mel, int64_t n, int64_t n_mels, int64_t n_ctx);
    // xa rows provided directly (stateless forward_decoder): n windows of T rows each
    void load_encoder_output_host(const float* xa_host, int64_t n, int64_t T);
    void run_encoder();      // conv stems .. ln_post .. cross K/V, from mel_rows
    void run_cross_kv();

    void set_special(const uint8_t* is_special_host);
    void begin(const int64_t* prompt, int64_t prompt_len, bool prefill = true);
    // one decoder position for R rows; tokens come from cur_tok
    void step_core(bool with_logits, int mask_mode, int k, bool greedy, int eot);
    void step_beams(int64_t n_rows, const int32_t* window_of_row, const int32_t* parent_row, const int64_t* token,
                    int apply_mask, int k, int64_t* topk_ids_out, float* topk_lp_out);
    // greedy loop on the device; returns per-window token lists
    void greedy_decode(const int64_t* prompt, int64_t prompt_len, int max_depth, int64_t eot,
                       std::vector<std::vector<int64_t>>& out);
};

// host pipeline (transcribe.cu)
void transcribe_windows(Session& s, int beam_size, int max_depth, const wb_special_ids& ids,
                        const uint8_t* is_special, std::vector<std::vector<int64_t>>& out);
std::vector<std::pair<int64_t, int64_t>> window_bounds(int64_t n_samples, int64_t sample_rate, int64_t window_len);
bool find_chunk_overlap(const int64_t* prev, int64_t n_prev, const int64_t* curr, int64_t n_curr, int64_t max_n_offsets,
                        int64_t min_n_overlaps, int64_t* prev_index, int64_t* curr_index);

// wav.cu: load_audio_waveform (src/bin/transcribe/main.rs:31-55)
void load_wav(const std::string& path, bool strict, std::vector<float>& out, int64_t& sample_rate, int& channels);

// npytree.cu: the reference's model-file format (src/model/load.rs, python/dump.py)
void npy_tree_probe(const std::string& dir, wb_dims& dims);
void npy_tree_load(Model& m, const std::string& dir);

// model.cu
void model_set_tensor(Model& m, const char* path, const float* data, const int64_t* shape, int ndim);
void model_finalize(Model& m);
const std::string& last_error_string();

}  // namespace wb

/* whisper_b200.h -- C ABI of the B200-native Whisper hot path (libwhisper_b200.so).
 *
 * The reference (Gadersd/whisper-burn) has no FFI: its hot path is Rust generic over
 * burn::tensor::backend::Backend.  This header is the boundary a thin Rust shim (rust/ in
 * this repo) binds with `extern "C"` so that whisper-burn's own public
 * functions keep their signatures while every tensor op runs in hand-written sm_100a CUDA:
 *
 *   audio::max_waveform_samples        src/audio.rs:12-17          -> wb_max_waveform_samples
 *   audio::prep_audio                  src/audio.rs:34-56          -> wb_prep_audio
 *   WhisperConfig / Whisper            src/model/mod.rs:16-71      -> wb_model_*
 *   model::load::load_whisper          src/model/load.rs:295-310   -> wb_model_set_tensor (same npy-tree paths)
 *   Whisper::forward_encoder           src/model/mod.rs:52-54      -> wb_forward_encoder
 *   Whisper::forward_decoder           src/model/mod.rs:56-62      -> wb_forward_decoder
 *   Whisper::{encoder,decoder}_ctx_size src/model/mod.rs:64-70     -> wb_model_get_dims
 *   beamsearch_next closure            src/transcribe.rs:253-307   -> wb_session_step (KV-cached, top-k only)
 *   beam::beam_search(_step)           src/beam.rs:9-79            -> wb_beam_* (host C++, same tie-breaks)
 *   mels_to_text (token part)          src/transcribe.rs:148-383   -> wb_transcribe_windows
 *   waveform_to_text (token part)      src/transcribe.rs:23-74     -> wb_waveform_to_tokens
 *   find_chunk_overlap                 src/transcribe.rs:76-110    -> wb_find_chunk_overlap
 *   first_repetition_end / repetition_period / find_repeated_tokens_index
 *                                      src/transcribe.rs:385-447   -> wb_first_repetition_end, wb_repetition_period,
 *                                                                     wb_find_repeated_tokens_index (compiled but unused there)
 *
 * Conventions:
 *   - plain pointers and sizes only; host buffers are caller-owned and only read/written
 *     during the call; every call is synchronous from the caller's point of view.
 *   - every function returns an int status (WB_OK == 0).  WB_ERR_INVALID_ARG marks what the
 *     reference treats as a contract violation (assert!/panic: shapes, n < 400 samples, ...);
 *     the Rust shim turns it into panic!, everything else into Err.  wb_last_error() returns a
 *     thread-local message.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     WB_ERR_CUDA.
 *   - a wb_model is immutable after wb_model_finalize and may be shared by threads; a
 *     wb_session (KV caches, workspaces, one CUDA stream) is used by one thread at a time.
 *   - `*_dev` variants take device pointers on the model's device (inputs resident in HBM).
 */
#ifndef WHISPER_B200_H
#define WHISPER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WB_OK 0
#define WB_ERR_INVALID_ARG 1   /* reference would assert!/panic */
#define WB_ERR_CUDA 2          /* no device / CUDA runtime failure */
#define WB_ERR_OOM 3
#define WB_ERR_STATE 4         /* call order violated (e.g. model not finalized) */
#define WB_ERR_UNSUPPORTED 5

#define WB_KV_F32 0            /* reference numerics */
#define WB_KV_F16 1            /* fp16 K/V cache (north_star); rounding restated by the oracle's kv_dtype="f16" */

typedef struct wb_model wb_model;
typedef struct wb_session wb_session;

/* WhisperConfig = AudioEncoderConfig + TextDecoderConfig (src/model/mod.rs:16-39,73-80,164-171) */
typedef struct wb_dims {
    int32_t n_mels, n_audio_ctx, n_audio_state, n_audio_head, n_audio_layer;
    int32_t n_vocab, n_text_ctx, n_text_state, n_text_head, n_text_layer;
} wb_dims;

/* the ids mels_to_text looks up in the tokenizer (src/transcribe.rs:179-185); the tokenizer
 * itself (src/token.rs) stays on the caller's side of the boundary */
typedef struct wb_special_ids {
    int64_t sot, lang, transcribe, notimestamps, eot;
} wb_special_ids;

const char* wb_version(void);
const char* wb_last_error(void);
int wb_device_count(int* n_out);

/* ---- audio.rs ---------------------------------------------------------------------- */
/* audio.rs:12-17 */
int64_t wb_max_waveform_samples(int64_t n_frame_max);
/* audio.rs:34-56: wave [n_batch, n_samples] -> mel_out [n_batch, 80, n_samples/160] (f32, row-major).
 * The max of audio.rs:50 is taken over the whole call (all batches), as in the reference.
 * WB_ERR_INVALID_ARG if n_samples < 400 (audio.rs:292). */
int wb_prep_audio(int device, const float* wave, int64_t n_batch, int64_t n_samples,
                  float* mel_out, int64_t* n_frames_out);
int wb_prep_audio_dev(int device, const float* wave_dev, int64_t n_batch, int64_t n_samples,
                      float* mel_out_dev, int64_t* n_frames_out);

/* ---- model (src/model/mod.rs, src/model/load.rs) ------------------------------------- */
int wb_model_create(const wb_dims* dims, int device, wb_model** out);
/* `path` is the reference's npy-tree path without ".npy" (load.rs:29-45, dump.py:130-213), e.g.
 * "encoder/block_0/attn/query/weight".  Linear weights are burn layout [d_in, d_out]
 * (dump.py:141-145), Conv1d weights [out, in, k], embeddings [rows, d], LayerNorm eps is a
 * 1-element tensor.  Data is copied. */
int wb_model_set_tensor(wb_model* m, const char* path, const float* data, const int64_t* shape, int ndim);
/* model::load::load_whisper (src/model/load.rs:295-310): builds a finalized model from the npy tree that
 * python/dump.py writes (one f32 .npy per tensor, payload = [dims..., values...], scalars as [1.0, value]);
 * the dimensions come from the tree itself.  wb_npy_tree_probe only reads the dimensions (host only). */
int wb_npy_tree_probe(const char* dir, wb_dims* dims_out);
int wb_model_load_npy_tree(const char* dir, int device, int ln_eps_outside, wb_model** out);
/* burn's nn::LayerNorm is third-party and un-vendored: burn 0.9 (the reference's pin, Cargo.lock:242-244)
 * normalises by (sqrt(var) + eps); later burn releases by sqrt(var + eps).  outside != 0 selects the
 * former (default).  Must be called before wb_model_finalize. */
int wb_model_set_layernorm_eps_mode(wb_model* m, int outside);
/* Validates that every tensor of the tree is present, uploads and re-lays the weights. */
int wb_model_finalize(wb_model* m);
void wb_model_destroy(wb_model* m);
int wb_model_get_dims(const wb_model* m, wb_dims* out);
/* 1 if every weight is exactly representable in fp16 (true for OpenAI checkpoints) and the
 * compact fp16 decoder weight storage is in use, 0 if weights are kept in fp32. */
int wb_model_weights_fp16_exact(const wb_model* m);

/* mod.rs:52-54 / 228-260: mel [n_batch, n_mels, n_ctx] -> out [n_batch, (n_ctx-1)/2+1, d].
 * WB_ERR_INVALID_ARG if n_mels != 80-config or n_ctx > n_audio_ctx (mod.rs:231-241). */
int wb_forward_encoder(wb_model* m, const float* mel, int64_t n_batch, int64_t n_mels, int64_t n_ctx,
                       float* out);
/* mod.rs:56-62 / 131-157, stateless: tokens [n_batch, seq_len] (i64), encoder_output
 * [n_batch, n_audio_ctx_used, d] -> logits_out [n_batch, seq_len, n_vocab].
 * WB_ERR_INVALID_ARG if seq_len > n_text_ctx (mod.rs:134-139). */
int wb_forward_decoder(wb_model* m, const int64_t* tokens, int64_t n_batch, int64_t seq_len,
                       const float* encoder_output, int64_t n_enc_ctx, float* logits_out);

/* ---- KV-cached decoding session -------------------------------------------------------- */
/* A session holds, for up to max_windows audio windows x max_beams live beams each: encoder
 * output, per-layer cross K/V (computed once per window), per-layer self K/V for
 * max_text_len positions, and all workspaces.  max_beams <= 7 and k <= 7 in wb_session_step (the decoders keep 8 candidates per
 * record; the reference searches with width 5, src/transcribe.rs:232); larger values are rejected with WB_ERR_INVALID_ARG. */
int wb_session_create(wb_model* m, int64_t max_windows, int64_t max_beams, int64_t max_text_len,
                      int kv_dtype, wb_session** out);
void wb_session_destroy(wb_session* s);
/* prep_audio + mel padding of mels_to_text (transcribe.rs:161-177) + forward_encoder + cross
 * K/V for n_windows waveforms; waves[i] has lens[i] samples (ragged; each >= 400). */
int wb_session_encode_waveforms(wb_session* s, const float* const* waves, const int64_t* lens,
                                int64_t n_windows);
/* same, windows already on the device, concatenated: window i = wave_dev[offsets[i] .. +lens[i]) */
int wb_session_encode_waveforms_dev(wb_session* s, const float* wave_dev, const int64_t* offsets,
                                    const int64_t* lens, int64_t n_windows);
/* forward_encoder + cross K/V from caller-provided mels [n_windows, n_mels, n_ctx] (no padding added) */
int wb_session_encode_mels(wb_session* s, const float* mel, int64_t n_windows, int64_t n_mels, int64_t n_ctx);
/* copies the session's padded mel [n_windows, 80, n_ctx] / encoder output [n_ctx_enc, d] of one window */
int wb_session_get_mel(wb_session* s, int64_t window, float* mel_out, int64_t capacity, int64_t* n_ctx_out);
int wb_session_get_encoder_output(wb_session* s, int64_t window, float* out, int64_t capacity, int64_t* n_ctx_out);
/* Starts decoding: clears the self K/V and feeds prompt[0 .. prompt_len-1) to one beam per window. */
int wb_session_begin(wb_session* s, const int64_t* prompt, int64_t prompt_len);
/* One beamsearch_next evaluation (transcribe.rs:253-307) for n_rows live beams:
 *   row r continues cache row parent_row[r] of window window_of_row[r] with token[r];
 *   apply_special_mask != 0 adds -inf on ids with is_special[id] != 0 (transcribe.rs:271-275);
 *   returns, per row, the k best (token id, f32 log-prob) of log_softmax over the vocabulary,
 *   ordered best first, ties broken towards the lower id (what beam.rs:81-110 keeps).
 * Rows of one window must be contiguous and use slots 0..n-1 of that window in order. */
int wb_session_step(wb_session* s, int64_t n_rows, const int32_t* window_of_row, const int32_t* parent_row,
                    const int64_t* token, int apply_special_mask, const uint8_t* is_special,
                    int k, int64_t* topk_ids_out, float* topk_logprob_out);

/* ---- transcribe.rs (token side) ---------------------------------------------------------- */
/* mels_to_text for a batch of independent windows: encode, then beam::beam_search with
 * beam_size / max_depth (reference: 5 / 100; greedy = beam_size 1).  tokens_out is
 * [n_windows, capacity]; each row gets prompt + generated ids (incl. EOT if reached). */
int wb_transcribe_windows(wb_session* s, const float* const* waves, const int64_t* lens, int64_t n_windows,
                          int beam_size, int max_depth, const wb_special_ids* ids, const uint8_t* is_special,
                          int64_t* tokens_out, int64_t capacity, int64_t* lens_out);
int wb_transcribe_windows_dev(wb_session* s, const float* wave_dev, const int64_t* offsets, const int64_t* lens,
                              int64_t n_windows, int beam_size, int max_depth, const wb_special_ids* ids,
                              const uint8_t* is_special, int64_t* tokens_out, int64_t capacity, int64_t* lens_out);
/* waveform_to_text without detokenisation: windowing (transcribe.rs:114-138), per-window
 * decoding, overlap merge (transcribe.rs:56-63).  Writes the merged ids.
 * sample_rate must be 16000 (WB_ERR_INVALID_ARG otherwise): the log-mel tables are the 16 kHz ones the reference's binary
 * always uses (src/bin/transcribe/main.rs:38-41); wb_prep_audio likewise assumes 16 kHz input. */
int wb_waveform_to_tokens(wb_session* s, const float* waveform, int64_t n_samples, int64_t sample_rate,
                          int beam_size, int max_depth, const wb_special_ids* ids, const uint8_t* is_special,
                          int64_t* tokens_out, int64_t capacity, int64_t* n_tokens_out);
/* The same for n_waveforms independent waveforms at once (the unit BASELINE.json shards over GPUs: "8x30 s chunks
 * batched"): the windows of all waveforms are decoded in one batch, each waveform's windows are merged in order.
 * tokens_out is [n_waveforms][capacity], n_tokens_out [n_waveforms]. */
int wb_waveforms_to_tokens(wb_session* s, const float* const* waveforms, const int64_t* n_samples, int64_t n_waveforms,
                           int64_t sample_rate, int beam_size, int max_depth, const wb_special_ids* ids,
                           const uint8_t* is_special, int64_t* tokens_out, int64_t capacity, int64_t* n_tokens_out);
/* transcribe.rs:114-138: number of windows and their [start, end) bounds */
int64_t wb_window_count(int64_t n_samples, int64_t sample_rate, int64_t window_len);
int wb_window_bounds(int64_t n_samples, int64_t sample_rate, int64_t window_len, int64_t* starts, int64_t* ends);
/* transcribe.rs:76-110; returns 1 and fills the indices if an overlap was found, else 0 */
int wb_find_chunk_overlap(const int64_t* prev, int64_t n_prev, const int64_t* curr, int64_t n_curr,
                          int64_t max_n_offsets, int64_t min_n_overlaps, int64_t* prev_index, int64_t* curr_index);
/* Repetition heuristics the reference compiles but calls only from its commented-out greedy loop (transcribe.rs:314-380).
 * transcribe.rs:385-393: position after the last mismatch between a block of `period` tokens and the block before it, walking
 * back from the end; `period` when none; -1 where the reference's usize arithmetic underflows (period > n). */
int64_t wb_first_repetition_end(const int64_t* tokens, int64_t n, int64_t period);
/* transcribe.rs:395-419: the period of a suffix repeated at least min_repetitions times before itself, 0 for None. */
int64_t wb_repetition_period(const int64_t* tokens, int64_t n, int64_t min_repetitions);
/* transcribe.rs:421-447: 1 and (first_repeat_index, end) when at least min_repeat_count earlier windows equal the last window of
 * window_size tokens, 0 for None, -1 where the reference unwraps a second repeat that does not exist. */
int wb_find_repeated_tokens_index(const int64_t* tokens, int64_t n, int64_t window_size, int64_t min_repeat_count,
                                  int64_t* first_repeat_index, int64_t* end);

/* ---- beam.rs (host) ------------------------------------------------------------------------ */
/* get_top_elements (beam.rs:81-110) on f64 scores: writes the indices of the kept elements in
 * the reference's output order (ascending score); returns how many were kept. */
int64_t wb_beam_get_top_elements(const double* scores, int64_t n, int64_t num, int64_t* idx_out);

/* beam::beam_search (beam.rs:9-37) with a table-driven `next` (see csrc/api.cu): the complete host search, for tests.  Returns the
 * length of the best sequence written to seq_out, or -1 on bad arguments. */
int64_t wb_beam_search_table(const double* table, int64_t n_ctx, int64_t n_vocab, int64_t first_token, int64_t eot,
                             int64_t beam_size, int64_t max_depth, int64_t* seq_out, int64_t capacity);

/* ---- transcribe binary helpers (host) ---------------------------------------------------------- */
/* load_audio_waveform (src/bin/transcribe/main.rs:31-55): PCM int samples / (2^(bits-1) - 1), float samples as they
 * are, interleaved.  strict_16k_mono != 0 enforces the reference's asserts (16 kHz, one channel) as WB_ERR_INVALID_ARG.
 * out may be NULL to query the sample count. */
int wb_load_wav(const char* path, int strict_16k_mono, float* out, int64_t capacity, int64_t* n_samples_out,
                int64_t* sample_rate_out, int* channels_out);

/* ---- measurement ----------------------------------------------------------------------------- */
/* Diagnostics: which persistent decoder kernel the last decode launch used: 6 = head-fused cluster decoder (decoder6.cu),
 * 5 = batched tensor-core (decoder5.cu), 4 = cluster/DSMEM (decoder4.cu), 3 = grid-barrier FMA fallback (decoder3.cu), 0 = none yet. */
int wb_session_last_decoder(const wb_session* s);
/* kernels launched by this library on this thread's sessions since the last reset */
int64_t wb_kernel_launch_count(void);
void wb_kernel_launch_count_reset(void);
/* device-side duration (CUDA events on the session stream) of the phases of the last
 * wb_transcribe_windows* call, in milliseconds: [0]=log-mel, [1]=encoder+cross-KV, [2]=decode, [3]=total */
int wb_session_last_timings(wb_session* s, float* ms_out4);
int wb_session_last_steps(wb_session* s, int64_t* n_steps_out);
/* roofline aid: re-runs n_steps greedy decoder steps on the currently encoded windows with CUDA
 * events (session stream) around the dominant kernel of the step -- the logits GEMV -- and around
 * each whole step; returns the average durations in milliseconds. */
int wb_session_profile_decode(wb_session* s, const wb_special_ids* ids, int n_steps, float* logits_kernel_ms,
                              float* step_ms);

#ifdef __cplusplus
}
#endif
#endif /* WHISPER_B200_H */

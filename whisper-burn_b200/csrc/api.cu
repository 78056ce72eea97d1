// extern "C" boundary (include/whisper_b200.h).  No exceptions cross it: every entry point maps
// wb::Error / std::exception to a status code and a thread-local message.
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <new>

#include "../host/beam.hpp"
#include "../host/repeat.hpp"
#include "session.h"

struct wb_model {
    wb::Model impl;
    // the stateless entry points (wb_forward_encoder / wb_forward_decoder) keep ONE session between calls instead of paying
    // dozens of cudaMalloc per call; it is rebuilt only when a call needs more windows or positions than it holds
    std::mutex fwd_mu;
    std::unique_ptr<wb::Session> fwd;
    wb::Session& forward_session(int64_t n_windows, int64_t text_len) {
        if (!fwd || fwd->max_windows < n_windows || fwd->t_max < text_len) {
            fwd.reset();
            fwd.reset(new wb::Session(&impl, std::max<int64_t>(n_windows, 1), 1, std::max<int64_t>(text_len, 2), WB_KV_F32));
        }
        return *fwd;
    }
    ~wb_model() { fwd.reset(); }   // sessions die before their model
};
struct wb_session {
    std::unique_ptr<wb::Session> impl;
    wb_model* model;
};

namespace {

template <typename F>
int guarded(F&& f) {
    try {
        f();
        return WB_OK;
    } catch (const wb::Error& e) {
        wb::set_last_error(e.what());
        return e.code;
    } catch (const std::bad_alloc&) {
        wb::set_last_error("host out of memory");
        return WB_ERR_OOM;
    } catch (const std::exception& e) {
        wb::set_last_error(e.what());
        return WB_ERR_CUDA;
    }
}

void require_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        wb::fail(WB_ERR_CUDA, std::string("no CUDA device available (there is no CPU fallback): ") + cudaGetErrorString(e));
    }
    if (device < 0 || device >= n) wb::fail(WB_ERR_INVALID_ARG, "device index out of range");
    WB_CUDA(cudaSetDevice(device));
}

void copy_tokens_out(const std::vector<std::vector<int64_t>>& toks, int64_t* tokens_out, int64_t capacity,
                     int64_t* lens_out) {
    for (size_t w = 0; w < toks.size(); ++w) {
        WB_REQUIRE((int64_t)toks[w].size() <= capacity, "tokens_out capacity too small");
        for (size_t i = 0; i < toks[w].size(); ++i) tokens_out[(int64_t)w * capacity + (int64_t)i] = toks[w][i];
        lens_out[w] = (int64_t)toks[w].size();
    }
}

void collect_timings(wb::Session& s) {
    WB_CUDA(cudaStreamSynchronize(s.st));
    cudaEventElapsedTime(&s.last_ms[0], s.ev[0], s.ev[1]);
    cudaEventElapsedTime(&s.last_ms[1], s.ev[1], s.ev[2]);
    cudaEventElapsedTime(&s.last_ms[2], s.ev[2], s.ev[3]);
    cudaEventElapsedTime(&s.last_ms[3], s.ev[0], s.ev[3]);
}

// prep_audio on the device; `wave` and `mel_out` are device pointers
void prep_audio_device(wb::Model& model, const float* wave_dev, int64_t n_batch, int64_t n_samples, float* mel_out_dev,
                       int64_t* n_frames_out, cudaStream_t st) {
    using namespace wb;
    WB_REQUIRE(n_batch >= 1, "prep_audio: n_batch must be >= 1");
    WB_REQUIRE(n_samples >= N_FFT, "prep_audio: waveform shorter than n_fft (audio.rs:292)");
    WB_REQUIRE(n_samples < ((int64_t)1 << 30), "prep_audio: waveform too long");
    const int F = (int)(n_samples / HOP);
    std::vector<LogMelWindow> lw((size_t)n_batch);
    for (int64_t b = 0; b < n_batch; ++b)
        lw[(size_t)b] = LogMelWindow{b * n_samples, (int)n_samples, F, F, 0, b * (int64_t)F * N_MELS};   // one global max (audio.rs:50)
    DevBuf<LogMelWindow> dwin;
    DevBuf<int> slot;
    DevBuf<float> rows;
    dwin.alloc((size_t)n_batch);
    slot.alloc(1);
    rows.alloc((size_t)n_batch * F * N_MELS + 4);
    WB_CUDA(cudaMemcpyAsync(dwin.p, lw.data(), lw.size() * sizeof(LogMelWindow), cudaMemcpyHostToDevice, st));
    if (F > 0) {
        launch_logmel(model, wave_dev, dwin.p, (int)n_batch, F, rows.p, slot.p, 1, st);
        for (int64_t b = 0; b < n_batch; ++b)
            launch_rows_to_chan(rows.p + b * (int64_t)F * N_MELS, mel_out_dev + b * (int64_t)N_MELS * F, F, st);
    }
    WB_CUDA(cudaStreamSynchronize(st));
    if (n_frames_out) *n_frames_out = F;
}

// a model holding only the frontend tables (prep_audio needs no weights); one per device, uploaded on first use
struct FrontendOnly {
    wb::Model m;
    explicit FrontendOnly(int device) {
        using namespace wb;
        m.device = device;
        const FrontendTables& ft = frontend_tables();
        auto up = [&](const void* src, size_t bytes) {
            void* p = nullptr;
            WB_CUDA(cudaMalloc(&p, bytes));
            m.allocs.push_back(p);
            WB_CUDA(cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice));
            return p;
        };
        m.basis_t = (float*)up(ft.basis_t.data(), ft.basis_t.size() * sizeof(float));
        m.mel_filt = (float*)up(ft.mel_filt.data(), ft.mel_filt.size() * sizeof(float));
        std::vector<int> rng(2 * N_MELS);
        for (int i = 0; i < N_MELS; ++i) { rng[2 * i] = ft.mel_lo[i]; rng[2 * i + 1] = ft.mel_hi[i]; }
        m.mel_range = (int*)up(rng.data(), rng.size() * sizeof(int));
    }
};

FrontendOnly& frontend_for(int device) {
    static std::mutex mu;
    static std::map<int, std::unique_ptr<FrontendOnly>> cache;
    std::lock_guard<std::mutex> lock(mu);
    auto& slot = cache[device];
    if (!slot) slot.reset(new FrontendOnly(device));
    return *slot;
}

}  // namespace

extern "C" {

const char* wb_version(void) { return "whisper_b200 0.1.0 (sm_100a)"; }
const char* wb_last_error(void) { return wb::last_error_string().c_str(); }

int wb_device_count(int* n_out) {
    return guarded([&] {
        WB_REQUIRE(n_out != nullptr, "null output");
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); n = 0; }
        *n_out = n;
    });
}

int64_t wb_max_waveform_samples(int64_t n_frame_max) {
    // audio.rs:12-17: HOP * (n_frame_max + 1) + is_odd(N_FFT) - 1
    return (int64_t)wb::HOP * (n_frame_max + 1) + (wb::N_FFT % 2) - 1;
}

int wb_prep_audio(int device, const float* wave, int64_t n_batch, int64_t n_samples, float* mel_out,
                  int64_t* n_frames_out) {
    return guarded([&] {
        WB_REQUIRE(wave && mel_out, "prep_audio: null pointer");
        WB_REQUIRE(n_batch >= 1, "prep_audio: n_batch must be >= 1");
        WB_REQUIRE(n_samples >= wb::N_FFT, "prep_audio: waveform shorter than n_fft (audio.rs:292)");
        require_device(device);
        FrontendOnly& fe = frontend_for(device);
        const int64_t F = n_samples / wb::HOP;
        wb::DevBuf<float> dw, dm;
        dw.alloc((size_t)(n_batch * n_samples));
        dm.alloc((size_t)(n_batch * wb::N_MELS * F) + 4);
        WB_CUDA(cudaMemcpy(dw.p, wave, (size_t)(n_batch * n_samples) * sizeof(float), cudaMemcpyHostToDevice));
        prep_audio_device(fe.m, dw.p, n_batch, n_samples, dm.p, n_frames_out, nullptr);
        WB_CUDA(cudaMemcpy(mel_out, dm.p, (size_t)(n_batch * wb::N_MELS * F) * sizeof(float), cudaMemcpyDeviceToHost));
    });
}

int wb_prep_audio_dev(int device, const float* wave_dev, int64_t n_batch, int64_t n_samples, float* mel_out_dev,
                      int64_t* n_frames_out) {
    return guarded([&] {
        WB_REQUIRE(wave_dev && mel_out_dev, "prep_audio: null pointer");
        require_device(device);
        FrontendOnly& fe = frontend_for(device);
        prep_audio_device(fe.m, wave_dev, n_batch, n_samples, mel_out_dev, n_frames_out, nullptr);
    });
}

int wb_model_create(const wb_dims* dims, int device, wb_model** out) {
    return guarded([&] {
        WB_REQUIRE(dims && out, "model_create: null pointer");
        WB_REQUIRE(dims->n_mels > 0 && dims->n_audio_ctx > 0 && dims->n_audio_state > 0 && dims->n_audio_head > 0 &&
                       dims->n_audio_layer > 0 && dims->n_vocab > 0 && dims->n_text_ctx > 0 && dims->n_text_state > 0 &&
                       dims->n_text_head > 0 && dims->n_text_layer > 0,
                   "model_create: non-positive dimension");
        require_device(device);
        wb_model* m = new wb_model();
        m->impl.dims = *dims;
        m->impl.device = device;
        *out = m;
    });
}

int wb_model_set_tensor(wb_model* m, const char* path, const float* data, const int64_t* shape, int ndim) {
    return guarded([&] {
        WB_REQUIRE(m != nullptr, "null model");
        wb::model_set_tensor(m->impl, path, data, shape, ndim);
    });
}

int wb_npy_tree_probe(const char* dir, wb_dims* dims_out) {
    return guarded([&] {
        WB_REQUIRE(dir && dims_out, "npy_tree_probe: null pointer");
        wb::npy_tree_probe(dir, *dims_out);
    });
}

int wb_model_load_npy_tree(const char* dir, int device, int ln_eps_outside, wb_model** out) {
    return guarded([&] {
        WB_REQUIRE(dir && out, "model_load_npy_tree: null pointer");
        wb_dims dims{};
        wb::npy_tree_probe(dir, dims);
        require_device(device);
        std::unique_ptr<wb_model> m(new wb_model());
        m->impl.dims = dims;
        m->impl.device = device;
        m->impl.ln_eps_outside = ln_eps_outside ? 1 : 0;
        wb::npy_tree_load(m->impl, dir);
        wb::model_finalize(m->impl);
        *out = m.release();
    });
}

int wb_model_set_layernorm_eps_mode(wb_model* m, int outside) {
    return guarded([&] {
        WB_REQUIRE(m != nullptr, "null model");
        if (m->impl.finalized) wb::fail(WB_ERR_STATE, "model already finalized");
        m->impl.ln_eps_outside = outside ? 1 : 0;
    });
}

int wb_model_finalize(wb_model* m) {
    return guarded([&] {
        WB_REQUIRE(m != nullptr, "null model");
        wb::model_finalize(m->impl);
    });
}

void wb_model_destroy(wb_model* m) { delete m; }

int wb_model_get_dims(const wb_model* m, wb_dims* out) {
    return guarded([&] {
        WB_REQUIRE(m && out, "null pointer");
        *out = m->impl.dims;
    });
}

int wb_model_weights_fp16_exact(const wb_model* m) { return (m && m->impl.fp16_exact) ? 1 : 0; }

int wb_forward_encoder(wb_model* m, const float* mel, int64_t n_batch, int64_t n_mels, int64_t n_ctx, float* out) {
    return guarded([&] {
        WB_REQUIRE(m && mel && out, "forward_encoder: null pointer");
        WB_REQUIRE(n_batch >= 1, "forward_encoder: n_batch must be >= 1");
        std::lock_guard<std::mutex> lock(m->fwd_mu);
        wb::Session& s = m->forward_session(n_batch, 2);
        s.encode_mels_host(mel, n_batch, n_mels, n_ctx);
        const int d = m->impl.dims.n_audio_state;
        WB_CUDA(cudaMemcpy(out, s.xa.p, (size_t)s.M_tot * d * sizeof(float), cudaMemcpyDeviceToHost));
    });
}

int wb_forward_decoder(wb_model* m, const int64_t* tokens, int64_t n_batch, int64_t seq_len, const float* encoder_output,
                       int64_t n_enc_ctx, float* logits_out) {
    return guarded([&] {
        WB_REQUIRE(m && tokens && encoder_output && logits_out, "forward_decoder: null pointer");
        const wb_dims& D = m->impl.dims;
        WB_REQUIRE(n_batch >= 1 && seq_len >= 1, "forward_decoder: empty input");
        WB_REQUIRE(seq_len <= D.n_text_ctx, "Token sequence length must not exceed n_text_ctx (mod.rs:134-139)");
        const int V = D.n_vocab;
        std::lock_guard<std::mutex> lock(m->fwd_mu);
        wb::Session& s = m->forward_session(n_batch, std::min<int64_t>(D.n_text_ctx, std::max<int64_t>(seq_len, 2)));
        s.load_encoder_output_host(encoder_output, n_batch, n_enc_ctx);
        struct FullLogits { wb::Session& s; explicit FullLogits(wb::Session& x) : s(x) { s.full_logits = true; } ~FullLogits() { s.full_logits = false; } } full_guard(s);
        // position by position through the cached step; logits of every position are kept
        wb::DevBuf<float> all;
        all.alloc((size_t)n_batch * seq_len * V);
        s.R = (int)n_batch;
        std::vector<int> rw((size_t)n_batch), tk((size_t)n_batch);
        for (int64_t r = 0; r < n_batch; ++r) rw[(size_t)r] = (int)r;
        WB_CUDA(cudaMemcpyAsync(s.row_window.p, rw.data(), rw.size() * sizeof(int), cudaMemcpyHostToDevice, s.st));
        WB_CUDA(cudaMemsetAsync(s.pos.p, 0, sizeof(int), s.st));
        WB_CUDA(cudaMemsetAsync(s.finished.p, 0, sizeof(int) * s.Rmax, s.st));
        s.anc_identity = true;
        s.host_pos = 0;
        for (int64_t p = 0; p < seq_len; ++p) {
            for (int64_t r = 0; r < n_batch; ++r) {
                const int64_t t = tokens[r * seq_len + p];
                WB_REQUIRE(t >= 0 && t < V, "forward_decoder: token out of range");
                tk[(size_t)r] = (int)t;
            }
            WB_CUDA(cudaMemcpyAsync(s.cur_tok.p, tk.data(), tk.size() * sizeof(int), cudaMemcpyHostToDevice, s.st));
            WB_CUDA(cudaStreamSynchronize(s.st));
            s.step_core(true, 0, 1, false, -1);
            for (int64_t r = 0; r < n_batch; ++r)
                WB_CUDA(cudaMemcpyAsync(all.p + ((size_t)r * seq_len + p) * V, s.logits.p + (size_t)r * V,
                                        (size_t)V * sizeof(float), cudaMemcpyDeviceToDevice, s.st));
        }
        WB_CUDA(cudaMemcpyAsync(logits_out, all.p, all.n * sizeof(float), cudaMemcpyDeviceToHost, s.st));
        WB_CUDA(cudaStreamSynchronize(s.st));
    });
}

int wb_session_create(wb_model* m, int64_t max_windows, int64_t max_beams, int64_t max_text_len, int kv_dtype,
                      wb_session** out) {
    return guarded([&] {
        WB_REQUIRE(m && out, "session_create: null pointer");
        std::unique_ptr<wb_session> s(new wb_session());
        s->model = m;
        s->impl.reset(new wb::Session(&m->impl, max_windows, max_beams, max_text_len, kv_dtype));
        *out = s.release();
    });
}

void wb_session_destroy(wb_session* s) { delete s; }

int wb_session_encode_waveforms(wb_session* s, const float* const* waves, const int64_t* lens, int64_t n_windows) {
    return guarded([&] {
        WB_REQUIRE(s && waves && lens, "encode: null pointer");
        s->impl->encode_waveforms_host(waves, lens, n_windows);
        WB_CUDA(cudaStreamSynchronize(s->impl->st));
    });
}

int wb_session_encode_waveforms_dev(wb_session* s, const float* wave_dev, const int64_t* offsets, const int64_t* lens,
                                    int64_t n_windows) {
    return guarded([&] {
        WB_REQUIRE(s && wave_dev && offsets && lens, "encode: null pointer");
        s->impl->encode_from_device_wave(wave_dev, offsets, lens, n_windows);
        WB_CUDA(cudaStreamSynchronize(s->impl->st));
    });
}

int wb_session_encode_mels(wb_session* s, const float* mel, int64_t n_windows, int64_t n_mels, int64_t n_ctx) {
    return guarded([&] {
        WB_REQUIRE(s && mel, "encode: null pointer");
        s->impl->encode_mels_host(mel, n_windows, n_mels, n_ctx);
    });
}

int wb_session_get_mel(wb_session* s, int64_t window, float* mel_out, int64_t capacity, int64_t* n_ctx_out) {
    return guarded([&] {
        WB_REQUIRE(s && mel_out && n_ctx_out, "get_mel: null pointer");
        wb::Session& S = *s->impl;
        if (!S.encoded) wb::fail(WB_ERR_STATE, "get_mel: nothing encoded");
        WB_REQUIRE(window >= 0 && window < S.n_windows, "get_mel: window out of range");
        const int Tm = S.win_Tm[(size_t)window];
        WB_REQUIRE(capacity >= (int64_t)Tm * wb::N_MELS, "get_mel: capacity too small");
        std::vector<float> rows((size_t)Tm * wb::N_MELS);
        WB_CUDA(cudaMemcpy(rows.data(), S.mel_rows.p + ((int64_t)window * S.TmS + 1) * wb::N_MELS,
                           rows.size() * sizeof(float), cudaMemcpyDeviceToHost));
        for (int t = 0; t < Tm; ++t)
            for (int c = 0; c < wb::N_MELS; ++c) mel_out[(int64_t)c * Tm + t] = rows[(size_t)t * wb::N_MELS + c];
        *n_ctx_out = Tm;
    });
}

int wb_session_get_encoder_output(wb_session* s, int64_t window, float* out, int64_t capacity, int64_t* n_ctx_out) {
    return guarded([&] {
        WB_REQUIRE(s && out && n_ctx_out, "get_encoder_output: null pointer");
        wb::Session& S = *s->impl;
        if (!S.encoded) wb::fail(WB_ERR_STATE, "get_encoder_output: nothing encoded");
        WB_REQUIRE(window >= 0 && window < S.n_windows, "get_encoder_output: window out of range");
        const int d = S.m->dims.n_audio_state;
        const int T = S.win_T[(size_t)window];
        WB_REQUIRE(capacity >= (int64_t)T * d, "get_encoder_output: capacity too small");
        WB_CUDA(cudaMemcpy(out, S.xa.p + S.win_row_off[(size_t)window] * d, (size_t)T * d * sizeof(float),
                           cudaMemcpyDeviceToHost));
        *n_ctx_out = T;
    });
}

int wb_session_begin(wb_session* s, const int64_t* prompt, int64_t prompt_len) {
    return guarded([&] {
        WB_REQUIRE(s && prompt, "begin: null pointer");
        s->impl->begin(prompt, prompt_len);
    });
}

int wb_session_step(wb_session* s, int64_t n_rows, const int32_t* window_of_row, const int32_t* parent_row,
                    const int64_t* token, int apply_special_mask, const uint8_t* is_special, int k,
                    int64_t* topk_ids_out, float* topk_logprob_out) {
    return guarded([&] {
        WB_REQUIRE(s && window_of_row && parent_row && token && topk_ids_out && topk_logprob_out, "step: null pointer");
        if (is_special) s->impl->set_special(is_special);
        WB_REQUIRE(!apply_special_mask || s->impl->have_special, "step: special mask requested but no is_special bitmap given");
        s->impl->step_beams(n_rows, window_of_row, parent_row, token, apply_special_mask, k, topk_ids_out, topk_logprob_out);
    });
}

int wb_transcribe_windows(wb_session* s, const float* const* waves, const int64_t* lens, int64_t n_windows,
                          int beam_size, int max_depth, const wb_special_ids* ids, const uint8_t* is_special,
                          int64_t* tokens_out, int64_t capacity, int64_t* lens_out) {
    return guarded([&] {
        WB_REQUIRE(s && waves && lens && ids && is_special && tokens_out && lens_out, "transcribe: null pointer");
        s->impl->encode_waveforms_host(waves, lens, n_windows);
        std::vector<std::vector<int64_t>> toks;
        wb::transcribe_windows(*s->impl, beam_size, max_depth, *ids, is_special, toks);
        copy_tokens_out(toks, tokens_out, capacity, lens_out);
        collect_timings(*s->impl);
    });
}

int wb_transcribe_windows_dev(wb_session* s, const float* wave_dev, const int64_t* offsets, const int64_t* lens,
                              int64_t n_windows, int beam_size, int max_depth, const wb_special_ids* ids,
                              const uint8_t* is_special, int64_t* tokens_out, int64_t capacity, int64_t* lens_out) {
    return guarded([&] {
        WB_REQUIRE(s && wave_dev && offsets && lens && ids && is_special && tokens_out && lens_out, "transcribe: null pointer");
        s->impl->encode_from_device_wave(wave_dev, offsets, lens, n_windows);
        std::vector<std::vector<int64_t>> toks;
        wb::transcribe_windows(*s->impl, beam_size, max_depth, *ids, is_special, toks);
        copy_tokens_out(toks, tokens_out, capacity, lens_out);
        collect_timings(*s->impl);
    });
}

int64_t wb_window_count(int64_t n_samples, int64_t sample_rate, int64_t window_len) {
    return (int64_t)wb::window_bounds(n_samples, sample_rate, window_len).size();
}

int wb_window_bounds(int64_t n_samples, int64_t sample_rate, int64_t window_len, int64_t* starts, int64_t* ends) {
    return guarded([&] {
        WB_REQUIRE(starts && ends, "window_bounds: null pointer");
        const auto b = wb::window_bounds(n_samples, sample_rate, window_len);
        for (size_t i = 0; i < b.size(); ++i) { starts[i] = b[i].first; ends[i] = b[i].second; }
    });
}

// windows of ALL waveforms are decoded together in batches of the session's capacity (they are independent,
// SURVEY.md F9), then each waveform's windows are merged in order exactly like the reference's sequential
// loop (transcribe.rs:42-71)
static void waveforms_to_tokens(wb::Session& S, const float* const* waveforms, const int64_t* n_samples, int64_t n_waveforms,
                                int64_t sample_rate, int beam_size, int max_depth, const wb_special_ids& ids,
                                const uint8_t* is_special, std::vector<std::vector<int64_t>>& out) {
    // the frontend tables (mel filterbank, DFT bins) are the 16 kHz ones: the reference builds them from the caller's rate
    // (audio.rs:44, 67-143) but its binary only ever passes 16 kHz (src/bin/transcribe/main.rs:38-41 asserts it)
    WB_REQUIRE(sample_rate == 16000, "waveform_to_tokens: only 16 kHz input is supported (frontend tables are built for 16 kHz)");
    const int64_t window_len = wb_max_waveform_samples(S.m->dims.n_audio_ctx - wb::MEL_PADDING);   // transcribe.rs:32-34
    std::vector<const float*> ptrs;
    std::vector<int64_t> lens;
    std::vector<int> owner;
    for (int64_t w = 0; w < n_waveforms; ++w)
        for (const auto& b : wb::window_bounds(n_samples[w], sample_rate, window_len)) {
            ptrs.push_back(waveforms[w] + b.first);
            lens.push_back(b.second - b.first);
            owner.push_back((int)w);
        }
    out.assign((size_t)n_waveforms, {});
    for (size_t b0 = 0; b0 < ptrs.size(); b0 += (size_t)S.max_windows) {
        const size_t nb = std::min(ptrs.size() - b0, (size_t)S.max_windows);
        S.encode_waveforms_host(ptrs.data() + b0, lens.data() + b0, (int64_t)nb);
        std::vector<std::vector<int64_t>> toks;
        wb::transcribe_windows(S, beam_size, max_depth, ids, is_special, toks);
        for (size_t i = 0; i < nb; ++i) {
            std::vector<int64_t>& tokens = out[(size_t)owner[b0 + i]];
            const auto& nt = toks[i];
            int64_t pi = 0, ci = 0;
            if (wb::find_chunk_overlap(tokens.data(), (int64_t)tokens.size(), nt.data(), (int64_t)nt.size(), 40, 3, &pi, &ci)) {
                tokens.resize((size_t)pi);                                    // transcribe.rs:59-60
                tokens.insert(tokens.end(), nt.begin() + ci, nt.end());
            } else {
                tokens.insert(tokens.end(), nt.begin(), nt.end());
            }
        }
    }
    collect_timings(S);
}

int wb_waveform_to_tokens(wb_session* s, const float* waveform, int64_t n_samples, int64_t sample_rate, int beam_size,
                          int max_depth, const wb_special_ids* ids, const uint8_t* is_special, int64_t* tokens_out,
                          int64_t capacity, int64_t* n_tokens_out) {
    return guarded([&] {
        WB_REQUIRE(s && waveform && ids && is_special && tokens_out && n_tokens_out, "waveform_to_tokens: null pointer");
        std::vector<std::vector<int64_t>> out;
        waveforms_to_tokens(*s->impl, &waveform, &n_samples, 1, sample_rate, beam_size, max_depth, *ids, is_special, out);
        WB_REQUIRE((int64_t)out[0].size() <= capacity, "tokens_out capacity too small");
        std::memcpy(tokens_out, out[0].data(), out[0].size() * sizeof(int64_t));
        *n_tokens_out = (int64_t)out[0].size();
    });
}

int wb_waveforms_to_tokens(wb_session* s, const float* const* waveforms, const int64_t* n_samples, int64_t n_waveforms,
                           int64_t sample_rate, int beam_size, int max_depth, const wb_special_ids* ids,
                           const uint8_t* is_special, int64_t* tokens_out, int64_t capacity, int64_t* n_tokens_out) {
    return guarded([&] {
        WB_REQUIRE(s && waveforms && n_samples && ids && is_special && tokens_out && n_tokens_out, "waveforms_to_tokens: null pointer");
        WB_REQUIRE(n_waveforms >= 1, "waveforms_to_tokens: n_waveforms must be >= 1");
        std::vector<std::vector<int64_t>> out;
        waveforms_to_tokens(*s->impl, waveforms, n_samples, n_waveforms, sample_rate, beam_size, max_depth, *ids, is_special, out);
        for (int64_t w = 0; w < n_waveforms; ++w) {
            WB_REQUIRE((int64_t)out[(size_t)w].size() <= capacity, "tokens_out capacity (per waveform) too small");
            std::memcpy(tokens_out + w * capacity, out[(size_t)w].data(), out[(size_t)w].size() * sizeof(int64_t));
            n_tokens_out[w] = (int64_t)out[(size_t)w].size();
        }
    });
}

int wb_find_chunk_overlap(const int64_t* prev, int64_t n_prev, const int64_t* curr, int64_t n_curr, int64_t max_n_offsets,
                          int64_t min_n_overlaps, int64_t* prev_index, int64_t* curr_index) {
    int64_t pi = 0, ci = 0;
    const bool found = wb::find_chunk_overlap(prev, n_prev, curr, n_curr, max_n_offsets, min_n_overlaps, &pi, &ci);
    if (found) {
        if (prev_index) *prev_index = pi;
        if (curr_index) *curr_index = ci;
    }
    return found ? 1 : 0;
}

int64_t wb_first_repetition_end(const int64_t* tokens, int64_t n, int64_t period) {
    if ((!tokens && n > 0) || n < 0) return -1;
    return wb::repeat::first_repetition_end(tokens, n, period);
}

int64_t wb_repetition_period(const int64_t* tokens, int64_t n, int64_t min_repetitions) {
    if ((!tokens && n > 0) || n < 0) return -1;
    return wb::repeat::repetition_period(tokens, n, min_repetitions);
}

int wb_find_repeated_tokens_index(const int64_t* tokens, int64_t n, int64_t window_size, int64_t min_repeat_count, int64_t* first_repeat_index,
                                  int64_t* end) {
    if ((!tokens && n > 0) || n < 0 || !first_repeat_index || !end) return -1;
    return wb::repeat::find_repeated_tokens_index(tokens, n, window_size, min_repeat_count, first_repeat_index, end);
}

int64_t wb_beam_get_top_elements(const double* scores, int64_t n, int64_t num, int64_t* idx_out) {
    if (!scores || !idx_out || n < 0 || num < 0) return -1;
    std::vector<double> v(scores, scores + n);
    const auto top = wb::beam::get_top_elements(v, [](double s) { return s; }, (size_t)num);
    for (size_t i = 0; i < top.size(); ++i) idx_out[i] = (int64_t)top[i];
    return (int64_t)top.size();
}

int wb_load_wav(const char* path, int strict_16k_mono, float* out, int64_t capacity, int64_t* n_samples_out, int64_t* sample_rate_out,
                int* channels_out) {
    return guarded([&] {
        WB_REQUIRE(path && n_samples_out, "load_wav: null pointer");
        std::vector<float> v;
        int64_t sr = 0;
        int ch = 0;
        wb::load_wav(path, strict_16k_mono != 0, v, sr, ch);
        *n_samples_out = (int64_t)v.size();
        if (sample_rate_out) *sample_rate_out = sr;
        if (channels_out) *channels_out = ch;
        if (out) {
            WB_REQUIRE(capacity >= (int64_t)v.size(), "load_wav: capacity too small");
            std::memcpy(out, v.data(), v.size() * sizeof(float));
        }
    });
}

int wb_session_last_decoder(const wb_session* s) { return s ? s->impl->last_decoder : -1; }

// beam::beam_search (src/beam.rs:9-37) over a TABLE-driven `next`: the continuation log-prob of token v after a beam whose last
// token is t and whose length is n is table[((t * 131 + n) % n_ctx) * n_vocab + v] (added to the beam's cumulative f64 log-prob);
// a beam is finished when its last token is eot.  Host only: lets the CPU tests drive the complete C++ search (host/beam.hpp:
// step, carry of finished beams, both tie-break rules) against the oracle without a GPU.
int64_t wb_beam_search_table(const double* table, int64_t n_ctx, int64_t n_vocab, int64_t first_token, int64_t eot, int64_t beam_size,
                             int64_t max_depth, int64_t* seq_out, int64_t capacity) {
    if (!table || !seq_out || n_ctx < 1 || n_vocab < 1 || beam_size < 1 || max_depth < 0) return -1;
    using Node = wb::beam::BeamNode<int64_t>;
    auto next = [&](const std::vector<Node>& beams) {
        std::vector<std::vector<std::pair<int64_t, double>>> out(beams.size());
        for (size_t b = 0; b < beams.size(); ++b) {
            const int64_t t = beams[b].seq.back(), n = (int64_t)beams[b].seq.size();
            const double* row = table + ((t * 131 + n) % n_ctx) * n_vocab;
            out[b].reserve((size_t)n_vocab);
            for (int64_t v = 0; v < n_vocab; ++v) out[b].emplace_back(v, beams[b].log_prob + row[v]);
        }
        return out;
    };
    auto fin = [&](const std::vector<int64_t>& seq) { return !seq.empty() && seq.back() == eot; };
    std::vector<Node> init(1);
    init[0].seq = {first_token};
    init[0].log_prob = 0.0;
    const std::vector<int64_t> best = wb::beam::beam_search(init, next, fin, (size_t)beam_size, (size_t)max_depth);
    if ((int64_t)best.size() > capacity) return -1;
    for (size_t i = 0; i < best.size(); ++i) seq_out[i] = best[i];
    return (int64_t)best.size();
}

int64_t wb_kernel_launch_count(void) { return wb::g_launch_count; }
void wb_kernel_launch_count_reset(void) { wb::g_launch_count = 0; }

int wb_session_last_timings(wb_session* s, float* ms_out4) {
    return guarded([&] {
        WB_REQUIRE(s && ms_out4, "null pointer");
        for (int i = 0; i < 4; ++i) ms_out4[i] = s->impl->last_ms[i];
    });
}

int wb_session_profile_decode(wb_session* s, const wb_special_ids* ids, int n_steps, float* logits_kernel_ms,
                              float* step_ms) {
    return guarded([&] {
        WB_REQUIRE(s && ids && logits_kernel_ms && step_ms, "null pointer");
        const int64_t prompt[4] = {ids->sot, ids->lang, ids->transcribe, ids->notimestamps};
        s->impl->profile_decode(prompt, 4, n_steps, ids->eot, logits_kernel_ms, step_ms);
    });
}

int wb_session_last_steps(wb_session* s, int64_t* n_steps_out) {
    return guarded([&] {
        WB_REQUIRE(s && n_steps_out, "null pointer");
        *n_steps_out = s->impl->last_steps;
    });
}

}  // extern "C"

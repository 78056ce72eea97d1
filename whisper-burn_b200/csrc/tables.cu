// Host-built constant tables of the log-mel frontend.
//
// The reference rebuilds these with f32 tensor ops on every prep_audio call
// (src/audio.rs:67-143 mel filterbank, :272-278 Hann window, :349-364 DFT angle matrix).
// They only depend on constants (sr = 16 kHz, n_fft = 400, n_mels = 80), so they are built
// once per process here -- in the SAME f32 operation order, because that order is part of
// the reference's numerics: the DFT angles are formed in f32 (k * fl32(2pi/400) * j, up to
// ~1253 rad), so the reference's twiddles are not the exact ones.  Transcendentals are
// evaluated in double on the f32 argument and rounded once (<= 0.5 ulp; libtorch's f32
// sin/cos/exp are <= 1 ulp).
#include <cmath>

#include "wb_internal.h"

namespace wb {

static float hz_to_mel_f64(double freq, double* out) {
    // audio.rs:198-230 (htk = false), f64 scalar
    const double f_min = 0.0, f_sp = 200.0 / 3.0;
    const double min_log_hz = 1000.0;
    const double min_log_mel = (min_log_hz - f_min) / f_sp;
    const double logstep = std::log(6.4) / 27.0;
    double mel = freq >= min_log_hz ? min_log_mel + std::log(freq / min_log_hz) / logstep : (freq - f_min) / f_sp;
    *out = mel;
    return (float)mel;
}

static FrontendTables build_tables() {
    FrontendTables t;
    // ---- Hann: sin(i * fl32(pi/400))^2   (audio.rs:272-278)
    t.hann.resize(N_FFT);
    const float c_h = (float)(M_PI / (double)N_FFT);
    for (int i = 0; i < N_FFT; ++i) {
        volatile float a = (float)i * c_h;
        float s = (float)std::sin((double)a);
        volatile float w = s * s;
        t.hann[i] = w;
    }
    // ---- DFT basis (audio.rs:349-364), stored transposed [j][cos(0..KPAD) | sin(0..KPAD)]
    t.basis_t.assign((size_t)N_FFT * 2 * KPAD, 0.0f);
    const float coe = (float)(M_PI * 2.0 / (double)N_FFT);
    for (int k = 0; k < N_FREQ; ++k) {
        volatile float kc = (float)k * coe;
        for (int j = 0; j < N_FFT; ++j) {
            volatile float a = kc * (float)j;
            float c = (float)std::cos((double)a);
            float s = (float)std::sin((double)a);
            volatile float cw = c * t.hann[j];
            volatile float sw = s * (-t.hann[j]);
            t.basis_t[(size_t)j * 2 * KPAD + k] = cw;
            t.basis_t[(size_t)j * 2 * KPAD + KPAD + k] = sw;
        }
    }
    // ---- mel filterbank (audio.rs:67-143), all f32
    const int n_mel_f = N_MELS + 2;
    double min_mel, max_mel;
    hz_to_mel_f64(0.0, &min_mel);
    hz_to_mel_f64(8000.0, &max_mel);
    const float step = (float)((max_mel - min_mel) / (double)(n_mel_f - 1));
    const float min_mel32 = (float)min_mel;
    const float f_sp = (float)(200.0 / 3.0);
    const float min_log_mel = (float)(1000.0 / (200.0 / 3.0));
    const float logstep = (float)(std::log(6.4) / 27.0);
    std::vector<float> mel_f(n_mel_f);
    for (int i = 0; i < n_mel_f; ++i) {
        volatile float mel = (float)i * step;
        mel = mel + min_mel32;
        // mel_to_hz_tensor (audio.rs:232-266): blend through a 0/1 mask
        float log_t = (mel >= min_log_mel) ? 1.0f : 0.0f;
        volatile float e_arg = (mel - min_log_mel);
        e_arg = e_arg * logstep;
        volatile float e = (float)std::exp((double)e_arg);
        e = e * 1000.0f;
        volatile float a = log_t * e;
        volatile float lin = mel * f_sp;
        lin = lin + 0.0f;
        volatile float b = (-log_t + 1.0f) * lin;
        volatile float f = a + b;
        mel_f[i] = f;
    }
    t.mel_filt.assign((size_t)N_MELS * N_FREQ, 0.0f);
    for (int m = 0; m < N_MELS; ++m) {
        volatile float fdiff_lo = mel_f[m + 1] - mel_f[m];
        volatile float fdiff_hi = mel_f[m + 2] - mel_f[m + 1];
        volatile float en = mel_f[m + 2] - mel_f[m];
        en = 1.0f / en;        // powf(-1.0)
        en = en * 2.0f;
        int lo = N_FREQ, hi = 0;
        for (int k = 0; k < N_FREQ; ++k) {
            volatile float fft = (float)k * 40.0f;     // fl32(16000/400) = 40
            volatile float r0 = mel_f[m] - fft;
            volatile float r2 = mel_f[m + 2] - fft;
            volatile float lower = (-r0) / fdiff_lo;
            volatile float upper = r2 / fdiff_hi;
            // tensor_min(lower, upper) = -(relu((-lower) - (-upper)) + (-upper))   helper.rs:16-22
            volatile float dd = (-lower) - (-upper);
            volatile float rl = dd > 0.0f ? (float)dd : 0.0f;
            volatile float mn = -(rl + (-upper));
            volatile float wv = mn > 0.0f ? (float)mn : 0.0f;   // relu
            wv = wv * en;
            t.mel_filt[(size_t)m * N_FREQ + k] = wv;
            if (wv != 0.0f) {
                if (k < lo) lo = k;
                if (k + 1 > hi) hi = k + 1;
            }
        }
        if (hi <= lo) { lo = 0; hi = 0; }
        t.mel_lo[m] = lo;
        t.mel_hi[m] = hi;
    }
    return t;
}

const FrontendTables& frontend_tables() {
    static const FrontendTables t = build_tables();
    return t;
}

}  // namespace wb

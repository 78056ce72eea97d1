// WAV ingestion for the transcribe caller: the reference's load_audio_waveform
// (src/bin/transcribe/main.rs:31-55, hound 3.5.0): PCM int samples are scaled by 1 / (2^(bits-1) - 1)
// (32 767 for 16 bit, NOT 32 768), IEEE-float samples are taken as they are; the reference asserts a 16 kHz,
// single-channel file.  Host code only.
#include <cstdio>
#include <cstring>
#include <fstream>

#include "session.h"

namespace wb {

namespace {
uint32_t rd32(const unsigned char* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
uint16_t rd16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
}  // namespace

// Reads the whole file; returns interleaved samples as f32.  strict: enforce the reference's asserts.
void load_wav(const std::string& path, bool strict, std::vector<float>& out, int64_t& sample_rate, int& channels) {
    std::ifstream f(path, std::ios::binary);
    if (!f) fail(WB_ERR_STATE, "cannot open " + path);
    std::vector<unsigned char> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if (buf.size() < 12 || std::memcmp(buf.data(), "RIFF", 4) != 0 || std::memcmp(buf.data() + 8, "WAVE", 4) != 0)
        fail(WB_ERR_INVALID_ARG, "not a RIFF/WAVE file: " + path);
    int fmt = 0, bits = 0;
    channels = 0;
    sample_rate = 0;
    const unsigned char* data = nullptr;
    size_t data_len = 0;
    for (size_t pos = 12; pos + 8 <= buf.size();) {
        const uint32_t len = rd32(buf.data() + pos + 4);
        const unsigned char* body = buf.data() + pos + 8;
        const size_t avail = buf.size() - (pos + 8);
        if (std::memcmp(buf.data() + pos, "fmt ", 4) == 0) {
            if (len < 16 || avail < 16) fail(WB_ERR_INVALID_ARG, "truncated fmt chunk: " + path);
            fmt = rd16(body);
            channels = rd16(body + 2);
            sample_rate = rd32(body + 4);
            bits = rd16(body + 14);
            if (fmt == 0xFFFE && len >= 26 && avail >= 26) fmt = rd16(body + 24);   // WAVE_FORMAT_EXTENSIBLE: first two bytes of the sub-format GUID
        } else if (std::memcmp(buf.data() + pos, "data", 4) == 0) {
            data = body;
            data_len = std::min((size_t)len, avail);
            break;
        }
        pos += 8 + (size_t)len + (len & 1);
    }
    if (!data || channels <= 0 || bits <= 0) fail(WB_ERR_INVALID_ARG, "WAV without fmt/data chunk: " + path);
    if (strict) {   // main.rs:43-44
        if (sample_rate != 16000) fail(WB_ERR_INVALID_ARG, "The audio sample rate must be 16k.");
        if (channels != 1) fail(WB_ERR_INVALID_ARG, "The audio must be single-channel.");
    }
    const int bytes = (bits + 7) / 8;
    const size_t n = data_len / (size_t)bytes;
    out.resize(n);
    if (fmt == 3) {   // SampleFormat::Float (main.rs:49)
        if (bits != 32) fail(WB_ERR_UNSUPPORTED, "only 32-bit float WAV is supported");
        std::memcpy(out.data(), data, n * 4);
    } else if (fmt == 1) {   // SampleFormat::Int (main.rs:50-53): s as f32 / max_int_val as f32
        const float max_int_val = (float)((1u << (bits - 1)) - 1u);
        for (size_t i = 0; i < n; ++i) {
            const unsigned char* p = data + i * bytes;
            int32_t s;
            if (bytes == 1) s = (int32_t)p[0] - 128;                                   // 8-bit WAV is unsigned; hound returns it signed
            else if (bytes == 2) s = (int16_t)rd16(p);
            else if (bytes == 3) s = (int32_t)((p[0] | (p[1] << 8) | (p[2] << 16)) << 8) >> 8;
            else s = (int32_t)rd32(p);
            out[i] = (float)s / max_int_val;
        }
    } else {
        fail(WB_ERR_UNSUPPORTED, "unsupported WAV format tag");
    }
}

}  // namespace wb

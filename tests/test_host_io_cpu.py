"""CPU tests of the transcribe-binary helpers either side of the hot path: WAV ingestion
(src/bin/transcribe/main.rs:31-55) and the tokenizer bridge (src/token.rs:26-47, src/transcribe.rs:179-185,243-251)."""
import json
import struct
import wave

import numpy as np
import pytest

import wb200  # noqa: F401
from whisper_burn_b200 import ffi, tokens, wav


def write_wav(path, samples_i16, rate=16000, channels=1):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(channels)
        w.setsampwidth(2)
        w.setframerate(rate)
        w.writeframes(np.asarray(samples_i16, dtype="<i2").tobytes())


def test_wav_int16_scaled_by_32767(tmp_path):
    s = np.array([0, 1, -1, 32767, -32768, 12345, -20000], dtype=np.int16)
    write_wav(tmp_path / "a.wav", s)
    got, sr = wav.load_audio_waveform(tmp_path / "a.wav")
    assert sr == 16000 and got.dtype == np.float32
    want = s.astype(np.float32) / np.float32(32767.0)          # main.rs:46-53: 2^(bits-1) - 1, NOT 32768
    assert np.array_equal(got, want) and got[4] < -1.0           # the most negative sample falls just below -1


def test_wav_reference_asserts(tmp_path):
    write_wav(tmp_path / "r.wav", np.zeros(10, np.int16), rate=22050)
    with pytest.raises(ffi.WbError) as e:
        wav.load_audio_waveform(tmp_path / "r.wav")
    assert e.value.code == ffi.WB_ERR_INVALID_ARG and "16k" in e.value.msg              # main.rs:43
    write_wav(tmp_path / "c.wav", np.zeros(10, np.int16), channels=2)
    with pytest.raises(ffi.WbError) as e:
        wav.load_audio_waveform(tmp_path / "c.wav")
    assert "single-channel" in e.value.msg                                             # main.rs:44
    got, sr = wav.load_audio_waveform(tmp_path / "r.wav", strict=False)
    assert sr == 22050 and len(got) == 10
    got, _ = wav.load_audio_waveform(tmp_path / "c.wav", strict=False)
    assert got.shape == (5, 2)
    with pytest.raises(ffi.WbError) as e:
        wav.load_audio_waveform(tmp_path / "missing.wav")
    assert e.value.code == ffi.WB_ERR_STATE


def test_wav_float32_and_24bit_and_extra_chunks(tmp_path):
    f = np.array([0.0, 0.5, -0.25, 1.0], dtype="<f4")
    hdr = b"RIFF" + struct.pack("<I", 4 + 8 + 16 + 8 + 6 + 8 + f.nbytes) + b"WAVE"
    hdr += b"fmt " + struct.pack("<IHHIIHH", 16, 3, 1, 16000, 16000 * 4, 4, 32)
    hdr += b"LIST" + struct.pack("<I", 5) + b"abcde\0"            # odd-sized chunk + pad byte is skipped
    (tmp_path / "f.wav").write_bytes(hdr + b"data" + struct.pack("<I", f.nbytes) + f.tobytes())
    got, _ = wav.load_audio_waveform(tmp_path / "f.wav")
    assert np.array_equal(got, f.astype(np.float32))            # SampleFormat::Float: samples as they are (main.rs:49)
    ints = [0, 1, -1, 8388607, -8388608]
    body = b"".join(struct.pack("<i", v)[:3] for v in ints)
    hdr = b"RIFF" + struct.pack("<I", 4 + 8 + 16 + 8 + len(body)) + b"WAVE" + b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, 16000, 48000, 3, 24)
    (tmp_path / "i24.wav").write_bytes(hdr + b"data" + struct.pack("<I", len(body)) + body)
    got, _ = wav.load_audio_waveform(tmp_path / "i24.wav")
    assert np.array_equal(got, np.array(ints, np.float32) / np.float32(8388607.0))


def make_tokenizer_json(path, n_text=50, multilingual=False):
    vocab = {f"t{i}": i for i in range(n_text)}
    names = ["<|endoftext|>", "<|startoftranscript|>"] + (["<|en|>", "<|de|>"] if multilingual else ["<|en|>"]) + \
            ["<|translate|>", "<|transcribe|>", "<|startoflm|>", "<|startofprev|>", "<|nospeech|>", "<|notimestamps|>", "<|0.00|>", "<|0.02|>"]
    added = [{"id": n_text + i, "content": c, "special": True, "single_word": False, "lstrip": False, "rstrip": False, "normalized": False}
             for i, c in enumerate(names)]
    added.append({"id": n_text + len(names), "content": "extra-word", "special": False})
    path.write_text(json.dumps({"version": "1.0", "added_tokens": added, "model": {"type": "BPE", "vocab": vocab, "merges": []}}))
    return n_text, names


def test_tokenizer_bridge_ids_and_bitmap(tmp_path):
    n_text, names = make_tokenizer_json(tmp_path / "tokenizer.json", multilingual=True)
    sp = tokens.special_tokens(tmp_path / "tokenizer.json", "de")
    assert sp.eot == n_text and sp.sot == n_text + 1 and sp.lang == n_text + 3          # token_to_id of the strings at token.rs:280-294
    assert sp.prompt() == [sp.sot, sp.lang, sp.transcribe, sp.notimestamps]              # transcribe.rs:203
    v = tokens.vocab_size(tmp_path / "tokenizer.json")
    assert v == n_text + len(names) + 1 == sp.n_vocab                                  # get_vocab_size(true), token.rs:45-47
    bm = tokens.is_special_bitmap(tmp_path / "tokenizer.json")
    assert bm.sum() == len(names) and bm[:n_text].sum() == 0 and bm[-1] == 0            # the non-special added token decodes to text
    assert np.array_equal(tokens.is_special_bitmap(tmp_path / "tokenizer.json", n_vocab=v + 7)[:v], bm)
    with pytest.raises(KeyError):
        tokens.special_tokens(tmp_path / "tokenizer.json", "zz")                        # special_token(..).unwrap() panics (transcribe.rs:182)

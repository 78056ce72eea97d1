//! Thin Rust surface over `include/whisper_b200.h` that keeps whisper-burn's public names
//! (reference: src/lib.rs:1-6 exports `audio, helper, model, token, transcribe, beam`).
//!
//! What changes for a caller of the reference:
//!   * `Whisper<B>` is no longer generic over a burn `Backend`; it owns a `wb_model*`.
//!   * tensors crossing the API are plain host `Vec<f32>` + shape (the reference's `Tensor<B, D>` are
//!     device handles of the backend that this crate replaces).
//!   * `Gpt2Tokenizer` (src/token.rs) stays the reference's own code; this crate asks it for 5 ids and
//!     the `is_special` bitmap only (transcribe.rs:179-185, 243-251).
//! Contract violations that `assert!`/panic in the reference (audio.rs:292, mod.rs:134-139, 231-241)
//! come back as WB_ERR_INVALID_ARG and are turned into `panic!` here; CUDA/OOM failures become `Err`.
#![allow(non_camel_case_types)]

use std::ffi::{c_char, c_int, c_void, CStr, CString};

pub mod ffi {
    use super::*;
    #[repr(C)]
    #[derive(Clone, Copy, Debug)]
    pub struct wb_dims {
        pub n_mels: i32, pub n_audio_ctx: i32, pub n_audio_state: i32, pub n_audio_head: i32, pub n_audio_layer: i32,
        pub n_vocab: i32, pub n_text_ctx: i32, pub n_text_state: i32, pub n_text_head: i32, pub n_text_layer: i32,
    }
    #[repr(C)]
    #[derive(Clone, Copy, Debug)]
    pub struct wb_special_ids { pub sot: i64, pub lang: i64, pub transcribe: i64, pub notimestamps: i64, pub eot: i64 }

    extern "C" {
        pub fn wb_last_error() -> *const c_char;
        pub fn wb_max_waveform_samples(n_frame_max: i64) -> i64;
        pub fn wb_prep_audio(device: c_int, wave: *const f32, n_batch: i64, n_samples: i64, mel_out: *mut f32, n_frames_out: *mut i64) -> c_int;
        pub fn wb_model_create(dims: *const wb_dims, device: c_int, out: *mut *mut c_void) -> c_int;
        pub fn wb_model_set_tensor(m: *mut c_void, path: *const c_char, data: *const f32, shape: *const i64, ndim: c_int) -> c_int;
        pub fn wb_npy_tree_probe(dir: *const c_char, dims_out: *mut wb_dims) -> c_int;
        pub fn wb_model_load_npy_tree(dir: *const c_char, device: c_int, ln_eps_outside: c_int, out: *mut *mut c_void) -> c_int;
        pub fn wb_load_wav(path: *const c_char, strict_16k_mono: c_int, out: *mut f32, capacity: i64, n_samples_out: *mut i64, sample_rate_out: *mut i64, channels_out: *mut c_int) -> c_int;
        pub fn wb_model_set_layernorm_eps_mode(m: *mut c_void, outside: c_int) -> c_int;
        pub fn wb_model_finalize(m: *mut c_void) -> c_int;
        pub fn wb_model_destroy(m: *mut c_void);
        pub fn wb_forward_encoder(m: *mut c_void, mel: *const f32, n_batch: i64, n_mels: i64, n_ctx: i64, out: *mut f32) -> c_int;
        pub fn wb_forward_decoder(m: *mut c_void, tokens: *const i64, n_batch: i64, seq_len: i64, enc: *const f32, n_enc_ctx: i64, logits_out: *mut f32) -> c_int;
        pub fn wb_session_create(m: *mut c_void, max_windows: i64, max_beams: i64, max_text_len: i64, kv_dtype: c_int, out: *mut *mut c_void) -> c_int;
        pub fn wb_session_destroy(s: *mut c_void);
        pub fn wb_waveform_to_tokens(s: *mut c_void, waveform: *const f32, n_samples: i64, sample_rate: i64, beam_size: c_int, max_depth: c_int,
                                     ids: *const wb_special_ids, is_special: *const u8, tokens_out: *mut i64, capacity: i64, n_tokens_out: *mut i64) -> c_int;
        pub fn wb_beam_get_top_elements(scores: *const f64, n: i64, num: i64, idx_out: *mut i64) -> i64;
        pub fn wb_beam_search_table(table: *const f64, n_ctx: i64, n_vocab: i64, first_token: i64, eot: i64, beam_size: i64, max_depth: i64, seq_out: *mut i64, capacity: i64) -> i64;
    }
}

pub type Error = Box<dyn std::error::Error + Send + Sync>;   // token::Result's error type (src/token.rs:6)

fn check(status: c_int) -> Result<(), Error> {
    if status == 0 { return Ok(()); }
    let msg = unsafe { CStr::from_ptr(ffi::wb_last_error()) }.to_string_lossy().into_owned();
    if status == 1 { panic!("{}", msg); }                       // WB_ERR_INVALID_ARG == reference assert!/panic
    Err(format!("whisper_b200 status {}: {}", status, msg).into())
}

pub mod audio {
    use super::*;
    /// src/audio.rs:12-17
    pub fn max_waveform_samples(n_frame_max: usize) -> usize { unsafe { ffi::wb_max_waveform_samples(n_frame_max as i64) as usize } }
    /// src/audio.rs:34-56: waveform [n_batch, n_samples] -> (mel [n_batch, 80, n_frames], n_frames)
    pub fn prep_audio(waveform: &[f32], n_batch: usize, _sample_rate: f64) -> Result<(Vec<f32>, usize), Error> {
        let n = waveform.len() / n_batch;
        let mut out = vec![0f32; n_batch * 80 * (n / 160)];
        let mut nf = 0i64;
        check(unsafe { ffi::wb_prep_audio(0, waveform.as_ptr(), n_batch as i64, n as i64, out.as_mut_ptr(), &mut nf) })?;
        Ok((out, nf as usize))
    }
}

pub mod model {
    use super::*;
    pub use ffi::wb_dims as WhisperConfig;   // src/model/mod.rs:16-39

    /// src/model/mod.rs:41-71
    pub struct Whisper { pub(crate) h: *mut c_void, pub config: WhisperConfig }
    unsafe impl Send for Whisper {}
    unsafe impl Sync for Whisper {}          // immutable after finalize
    impl Drop for Whisper { fn drop(&mut self) { unsafe { ffi::wb_model_destroy(self.h) } } }

    impl Whisper {
        /// Builds the model from the reference's npy tree (src/model/load.rs:19-310): `tensors` yields
        /// (path without ".npy", shape, values) for every file of the tree.
        pub fn from_tensors<'a, I: IntoIterator<Item = (&'a str, &'a [i64], &'a [f32])>>(config: WhisperConfig, tensors: I) -> Result<Self, Error> {
            let mut h = std::ptr::null_mut();
            check(unsafe { ffi::wb_model_create(&config, 0, &mut h) })?;
            let w = Whisper { h, config };
            for (path, shape, data) in tensors {
                let c = CString::new(path)?;
                check(unsafe { ffi::wb_model_set_tensor(w.h, c.as_ptr(), data.as_ptr(), shape.as_ptr(), shape.len() as c_int) })?;
            }
            check(unsafe { ffi::wb_model_finalize(w.h) })?;
            Ok(w)
        }
        /// model::load::load_whisper (src/model/load.rs:295-310): the directory python/dump.py writes.
        pub fn load_npy_tree(dir: &str) -> Result<Self, Error> {
            let c = CString::new(dir)?;
            let mut config: WhisperConfig = unsafe { std::mem::zeroed() };
            check(unsafe { ffi::wb_npy_tree_probe(c.as_ptr(), &mut config) })?;
            let mut h = std::ptr::null_mut();
            check(unsafe { ffi::wb_model_load_npy_tree(c.as_ptr(), 0, 1, &mut h) })?;
            Ok(Whisper { h, config })
        }
        /// mod.rs:52-54: mel [n_batch, 80, n_ctx] -> [n_batch, (n_ctx-1)/2+1, d]
        pub fn forward_encoder(&self, mel: &[f32], n_batch: usize, n_ctx: usize) -> Result<Vec<f32>, Error> {
            let t = (n_ctx - 1) / 2 + 1;
            let mut out = vec![0f32; n_batch * t * self.config.n_audio_state as usize];
            check(unsafe { ffi::wb_forward_encoder(self.h, mel.as_ptr(), n_batch as i64, 80, n_ctx as i64, out.as_mut_ptr()) })?;
            Ok(out)
        }
        /// mod.rs:56-62 (stateless): tokens [n_batch, seq_len], encoder_output [n_batch, n_enc_ctx, d] -> logits [n_batch, seq_len, n_vocab]
        pub fn forward_decoder(&self, tokens: &[i64], n_batch: usize, encoder_output: &[f32], n_enc_ctx: usize) -> Result<Vec<f32>, Error> {
            let seq_len = tokens.len() / n_batch;
            let mut out = vec![0f32; n_batch * seq_len * self.config.n_vocab as usize];
            check(unsafe { ffi::wb_forward_decoder(self.h, tokens.as_ptr(), n_batch as i64, seq_len as i64, encoder_output.as_ptr(), n_enc_ctx as i64, out.as_mut_ptr()) })?;
            Ok(out)
        }
        pub fn encoder_ctx_size(&self) -> usize { self.config.n_audio_ctx as usize }   // mod.rs:64-66
        pub fn decoder_ctx_size(&self) -> usize { self.config.n_text_ctx as usize }    // mod.rs:68-70
    }
}

pub mod beam {
    use super::*;
    /// src/beam.rs:81-110 over f64 scores (indices of the kept elements, ascending score).  The search loop
    /// itself (beam.rs:9-79) runs inside the library (host/beam.hpp) with the same tie-breaks.
    pub fn get_top_elements(scores: &[f64], num: usize) -> Vec<usize> {
        let mut idx = vec![0i64; num.max(1)];
        let n = unsafe { ffi::wb_beam_get_top_elements(scores.as_ptr(), scores.len() as i64, num as i64, idx.as_mut_ptr()) };
        idx[..n.max(0) as usize].iter().map(|&i| i as usize).collect()
    }
}

pub mod transcribe {
    use super::*;
    /// What `mels_to_text` looks up in the tokenizer (src/transcribe.rs:179-185) + `is_special` for every id (:243-251).
    pub struct SpecialTokens { pub ids: ffi::wb_special_ids, pub is_special: Vec<u8> }

    /// src/transcribe.rs:23-74 without detokenisation: merged token ids of the whole waveform
    /// (the caller detokenises with the reference's own `Gpt2Tokenizer::decode`).
    pub fn waveform_to_tokens(whisper: &model::Whisper, sp: &SpecialTokens, waveform: Vec<f32>, sample_rate: usize) -> Result<Vec<usize>, Error> {
        let (beam_size, max_depth) = (5, 100);                                // transcribe.rs:232-233
        let mut s = std::ptr::null_mut();
        check(unsafe { ffi::wb_session_create(whisper.h, 8, beam_size as i64, (4 + max_depth + 1) as i64, 0, &mut s) })?;
        let cap = (waveform.len() / 1000 + 2) * (4 + max_depth as usize + 1) + 16;
        let mut out = vec![0i64; cap];
        let mut n = 0i64;
        let st = unsafe { ffi::wb_waveform_to_tokens(s, waveform.as_ptr(), waveform.len() as i64, sample_rate as i64, beam_size, max_depth,
                                                     &sp.ids, sp.is_special.as_ptr(), out.as_mut_ptr(), cap as i64, &mut n) };
        unsafe { ffi::wb_session_destroy(s) };
        check(st)?;
        Ok(out[..n as usize].iter().map(|&t| t as usize).collect())
    }
}

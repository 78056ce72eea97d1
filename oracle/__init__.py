"""CPU oracle for the whisper-burn hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

This package is a CPU restatement (PyTorch-CPU fp32 + numpy) of the reference's
algorithm for the one hot path the build accelerates:

    audio::prep_audio  ->  Whisper::forward_encoder  ->  loop { Whisper::forward_decoder
    -> log_softmax -> beam::beam_search_step }          (reference: src/audio.rs,
    src/helper.rs, src/model/mod.rs, src/beam.rs, src/transcribe.rs)

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker or the timed CPU
baseline.  Nothing under ``whisper-burn_b200/`` imports it; the product path fails
loudly when the CUDA library is missing.

PARITY UNPINNED.  The reference is Rust on top of un-vendored crates (burn 0.9.0 @
fb2a71bb -> burn-tch -> tch 0.13 -> libtorch, Cargo.lock:242-244,333,3319,3579); it has
no tests, no golden vectors and cannot be compiled here (no cargo/rustc, no network).
The arithmetic lives in libtorch, so the restatement vehicle is PyTorch-CPU fp32 -- the
same library family the reference's own CPU path (TchBackend<f32> on TchDevice::Cpu,
src/bin/convert/main.rs:32-33) executes.  Third-party numerics that could not be read
from source (burn LayerNorm eps placement, burn softmax/log_softmax composition, burn
GELU form) are restated from the published burn 0.9 behaviour and exposed as switches
(see oracle.model.OracleOptions); DESIGN.md lists them.
"""

from . import audio, beam, model, synth, transcribe  # noqa: F401

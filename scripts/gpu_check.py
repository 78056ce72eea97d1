"""Stage-by-stage GPU vs oracle comparison (development aid; the pytest -m gpu suite is the gate)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import wb200  # noqa
from whisper_burn_b200 import audio, model, transcribe, ffi
from oracle import audio as oa, model as om, synth, transcribe as ot

def stats(name, a, b):
    d = np.abs(a - b)
    print(f"{name}: shape {a.shape} max_abs {d.max():.3e} rel_to_scale {d.max()/max(np.abs(b).max(),1e-30):.3e} mean_abs {d.mean():.3e}", flush=True)

model_name = sys.argv[1] if len(sys.argv) > 1 else "test-a"
wave = synth.chunk_waveform(0)[:238559]
mel = audio.prep_audio(wave[None]); mref = oa.prep_audio(torch.from_numpy(wave)[None]).numpy()
stats("mel[238559]", mel, mref)
for n, kind in [(400, "noise"), (16000, "chirp"), (98882, "mix"), (4000, "click")]:
    w = synth.waveform(n, seed=3, kind=kind)
    stats(f"mel[{n},{kind}]", audio.prep_audio(w[None]), oa.prep_audio(torch.from_numpy(w)[None]).numpy())
wb = np.stack([synth.waveform(16000, seed=s) for s in (1, 2)])
stats("mel batch2", audio.prep_audio(wb), oa.prep_audio(torch.from_numpy(wb)).numpy())

dims, w_np, w_t = synth.make_weights(model_name)
sp = synth.special_tokens(dims)
is_special = (np.arange(dims.n_vocab) >= sp.first_special).astype(np.uint8)
t = time.time(); wh = model.Whisper(dims, w_np); print("model load", time.time() - t, "fp16 exact", wh.weights_fp16_exact, flush=True)
melp = ot.pad_mel(torch.from_numpy(mref), dims.n_audio_ctx)
enc = wh.forward_encoder(melp.numpy()); eref = om.forward_encoder(w_t, dims, melp).numpy()
stats("encoder", enc, eref)
m2 = melp[:, :, :301]
stats("encoder Tm=301", wh.forward_encoder(m2.numpy()), om.forward_encoder(w_t, dims, m2).numpy())
toks = np.array([sp.prompt() + [5, 17, 99, 3], sp.prompt() + [8, 1, 2, 300]], dtype=np.int64)
xa2 = np.concatenate([eref, eref[:, ::-1].copy()], 0)
lg = wh.forward_decoder(toks, xa2); lref = om.forward_decoder(w_t, dims, torch.from_numpy(toks), torch.from_numpy(xa2)).numpy()
stats("decoder logits", lg, lref)
sess = transcribe.Session(wh, max_windows=3, max_beams=5, max_text_len=4 + 100 + 1)
for bs, depth in [(1, 30), (5, 12)]:
    t = time.time()
    got = sess.transcribe_windows([wave, wave[:98882]], sp, is_special, beam_size=bs, max_depth=depth)
    dt = time.time() - t
    want = [ot.mels_to_tokens(w_t, dims, sp, oa.prep_audio(torch.from_numpy(x)[None]), beam_size=bs, max_depth=depth) for x in (wave, wave[:98882])]
    print(f"beam {bs} depth {depth}: identical={got == want} time {dt:.3f}s timings {sess.last_timings_ms()} steps {sess.last_steps()}", flush=True)
    if got != want:
        for g, w_ in zip(got, want): print(" got ", g, "\n want", w_)
stats("session mel", sess.get_mel(0), ot.pad_mel(torch.from_numpy(mref), dims.n_audio_ctx)[0].numpy())
print("launches", ffi.lib().wb_kernel_launch_count())

"""One batched decode (decoder5.cu) for ncu: python scripts/prof_decode5.py [model] [n_chunks] [kv] [depth]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import wb200  # noqa
from whisper_burn_b200 import ffi, model, synth, transcribe
name = sys.argv[1] if len(sys.argv) > 1 else "small.en"
n_chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 8
kv = sys.argv[3] if len(sys.argv) > 3 else "f32"
depth = int(sys.argv[4]) if len(sys.argv) > 4 else 24
dims, w_np = synth.make_weights(name)
sp = synth.special_tokens(dims)
wh = model.Whisper(dims, w_np)
waves = []
for c in range(n_chunks):
    chunk = synth.chunk_waveform(c)
    waves += [chunk[:238559], chunk[190559:429118], chunk[381118:]]
sess = transcribe.Session(wh, len(waves), 1, 4 + depth + 1, kv_dtype=ffi.WB_KV_F16 if kv == "f16" else ffi.WB_KV_F32)
toks = sess.transcribe_windows(waves, sp, sp.is_special_bitmap(), beam_size=1, max_depth=depth)
print("decoder", sess.last_decoder(), "rows", len(waves), "tokens", sum(len(t) for t in toks))

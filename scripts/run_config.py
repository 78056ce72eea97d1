"""Per-GPU workload of a BASELINE.json config on one GPU: python scripts/run_config.py MODEL N_CHUNKS BEAM KV DEPTH [check]
Times the hot path (CUDA events on the library stream) and, with `check`, compares the token ids of the default decoder
(decoder4/5) with the grid-barrier FMA decoder (WB200_DECODER=3) on the same inputs."""
import json, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import wb200  # noqa
from whisper_burn_b200 import ffi, model, synth, transcribe

name, n_chunks, beam, kv, depth = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5])
check = len(sys.argv) > 6 and sys.argv[6] == "check"
t0 = time.time()
dims, w_np = synth.make_weights(name)
sp = synth.special_tokens(dims)
wh = model.Whisper(dims, w_np)
del w_np
print(f"model {name} built in {time.time() - t0:.1f} s, fp16-exact weights {wh.weights_fp16_exact}", flush=True)
waves = []
for c in range(n_chunks):
    chunk = synth.chunk_waveform(c)
    waves += [chunk[:238559], chunk[190559:429118], chunk[381118:]]
kvd = ffi.WB_KV_F16 if kv == "f16" else ffi.WB_KV_F32


def run():
    sess = transcribe.Session(wh, len(waves), max(beam, 1), 4 + depth + 1, kv_dtype=kvd)
    toks = sess.transcribe_windows(waves, sp, sp.is_special_bitmap(), beam_size=beam, max_depth=depth)   # warm-up
    toks = sess.transcribe_windows(waves, sp, sp.is_special_bitmap(), beam_size=beam, max_depth=depth)
    return toks, sess.last_timings_ms(), sess.last_decoder()


toks, t, dec = run()
out = {"model": name, "chunks": n_chunks, "windows": len(waves), "beam": beam, "kv": kv, "max_depth": depth, "decoder": dec,
       "ms": t, "audio_s_per_s": 30.0 * n_chunks / (t["total"] / 1e3), "tokens_checksum": int(sum(sum(x) for x in toks) % (1 << 31))}
if check:
    os.environ["WB200_DECODER"] = "3"
    toks3, t3, dec3 = run()
    out["grid_barrier_decoder_ms"] = t3
    out["same_tokens_as_decoder3"] = toks3 == toks
print(json.dumps(out))

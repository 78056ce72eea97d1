import os, sys, time, subprocess
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "torch threads", torch.get_num_threads(), flush=True)
print(subprocess.run("lscpu | grep -E 'Model name|^CPU\\(s\\)|Thread|Socket'; cat /sys/fs/cgroup/cpu.max 2>/dev/null", shell=True, capture_output=True, text=True).stdout, flush=True)
from oracle import audio, synth, transcribe
dims, w_np, w = synth.make_weights("tiny.en")
sp = synth.special_tokens(dims)
x = synth.chunk_waveform(0)[:238559]
for nt in (4, 8, 16, 32, 64):
    if nt > (os.cpu_count() or 1): break
    torch.set_num_threads(nt)
    t = time.time(); mel = audio.prep_audio(torch.from_numpy(x)[None]); t1 = time.time() - t
    t = time.time(); transcribe.mels_to_tokens(w, dims, sp, mel, beam_size=1, max_depth=10, use_cache=False); t2 = time.time() - t
    print(f"threads {nt}: prep_audio {t1:.2f}s  encoder+10 no-cache steps {t2:.2f}s", flush=True)

"""Stress: many back-to-back transcribe calls (host PCM) on whisper-large-v3; reports the first failure.
    python tools/stress_transcribe.py [batch] [iters] [lean]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from b200asr.config import PRESETS
from b200asr.engine import WhisperEngine
from b200asr.synth import synth_batch, synth_whisper_checkpoint
from b200asr.weights import fold_whisper

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 300
lean = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dims = PRESETS["whisper-large-v3"]
tensors = fold_whisper(synth_whisper_checkpoint(dims, 20260, pos_scale=100.0), dims, [1, 2, 7], [220, 50257])
MAXB = int(sys.argv[4]) if len(sys.argv) > 4 else B
eng = WhisperEngine(dims, tensors, precision="bf16", max_batch=MAXB, max_samples=128000)
del tensors
try:
    eng.set_option("stream_lean", lean)
except Exception:
    pass
prompt = [50258, 50259, 50360, 50364]
eng.set_decode_options(stop_ids=[], generate_limit=33)
pcm = synth_batch(B, 128000)
ref = None
t0 = time.time()
for i in range(iters):
    try:
        toks = eng.transcribe(pcm, prompt, max_new=33)
    except Exception as ex:
        print(f"FAILED at iteration {i}: {ex}", flush=True)
        sys.exit(1)
    if ref is None:
        ref = toks
    elif toks != ref:
        print(f"MISMATCH at iteration {i}", flush=True)
        sys.exit(2)
print(f"batch {B} lean {lean}: {iters} iterations ok, {1e3 * (time.time() - t0) / iters:.2f} ms each", flush=True)

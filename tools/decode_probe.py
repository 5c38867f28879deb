"""Decode-step timing probe on whisper-large-v3 (bf16): per-phase barrier-to-barrier times of the
persistent kernel and ms/step for a few prefetch settings."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from b200asr.config import PRESETS
from b200asr.engine import WhisperEngine
from b200asr.synth import synth_batch, synth_whisper_checkpoint
from b200asr.weights import fold_whisper

preset = sys.argv[1] if len(sys.argv) > 1 else "whisper-large-v3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dims = PRESETS[preset]
raw = synth_whisper_checkpoint(dims, 20260)
tensors = fold_whisper(raw, dims, [1, 2, 7], [220, 50257 if dims.vocab > 50257 else 2])
del raw
eng = WhisperEngine(dims, tensors, precision="bf16", max_batch=B, max_samples=128000)
del tensors
prompt = [50258, 50259, 50360, 50364] if dims.vocab > 50364 else [3, 10, 11, 12]
pcm = synth_batch(B, 128000)
eng.set_decode_options(stop_ids=[], generate_limit=33)
eng.upload_pcm(pcm)
stream = torch.cuda.ExternalStream(eng.stream_ptr)

def timed(fn, n=1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(stream)
    for _ in range(n): fn()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

eng.encode_resident()
print("encoder ms", timed(eng.encode_resident, 3), flush=True)
for mega, ahead in ((1, 0), (1, 16), (1, 32), (1, 64), (0, 0)):
    eng.set_option("mega", mega); eng.set_option("pf_ahead_mb", ahead)
    eng.prefill(prompt, want_logits=False); eng.decode(max_steps=4)
    eng.prefill(prompt, want_logits=False)
    t_pre = timed(lambda: eng.prefill(prompt, want_logits=False))
    t_dec = timed(lambda: eng.decode(max_steps=32)) / 32
    t_all = timed(lambda: eng.transcribe_resident(prompt, max_new=33), 3)
    print(f"mega={mega} pf_ahead={ahead}MB: prefill {t_pre:.3f} ms, decode {t_dec:.4f} ms/step, transcribe {t_all:.2f} ms", flush=True)
eng.set_option("mega", 1); eng.set_option("pf_ahead_mb", 32); eng.set_option("mega_timing", 1)
eng.prefill(prompt, want_logits=False)
eng.decode(max_steps=3)
t = eng.get_stage("mega_timing", 16384)
L = dims.dec_layers
per = 8 * L + 1
names = ["qkv", "self", "out", "cq", "cross", "cout", "fc1", "fc2"]
print("stamps", len(t), "step totals us:", [round(float(t[i * per:(i + 1) * per].sum()), 1) for i in range(len(t) // per)])
step = t[per:2 * per]
for j, nme in enumerate(names):
    print(f"  {nme:6s} mean {step[j:8 * L:8].mean():7.2f} us  min {step[j:8 * L:8].min():7.2f}  max {step[j:8 * L:8].max():7.2f}")
print(f"  head   {step[8 * L]:7.2f} us")

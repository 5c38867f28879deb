"""whisper-large-v3 decode step: bf16 vs FP8 (E4M3 weights, per-row scale) streaming kernel -- ms/step, logits distance, tokens.
    python tools/fp8_probe.py [batch]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from b200asr.config import PRESETS
from b200asr.engine import WhisperEngine
from b200asr.synth import synth_batch, synth_whisper_checkpoint
from b200asr.weights import fold_whisper

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dims = PRESETS["whisper-large-v3"]
raw = synth_whisper_checkpoint(dims, 20260, pos_scale=100.0)
tensors = fold_whisper(raw, dims, [1, 2, 7], [220, 50257])
del raw
eng = WhisperEngine(dims, tensors, precision="bf16", max_batch=B, max_samples=128000)
del tensors
prompt = [50258, 50259, 50360, 50364]
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
res = {}
for name, f8 in (("bf16", 0), ("fp8", 1)):
    eng.set_option("fp8", f8)
    eng.prefill(prompt, want_logits=False); eng.decode(max_steps=4)
    t_pre = timed(lambda: eng.prefill(prompt, want_logits=False), 3)
    eng.prefill(prompt, want_logits=False)
    t_dec = timed(lambda: eng.decode(max_steps=32)) / 32
    t_all = timed(lambda: eng.transcribe_resident(prompt, max_new=33), 3)
    toks = eng.transcribe_resident(prompt, max_new=33)
    lg, _ = eng.prefill(prompt)
    lgs = [lg.copy()]
    for t in toks[0][:6]:
        l2, _ = eng.decode_step(token_in=np.full(B, t, np.int32))
        lgs.append(l2.copy())
    res[name] = (toks, np.stack(lgs))
    print(f"{name}: prefill {t_pre:.3f} ms, decode {t_dec:.4f} ms/step, transcribe {t_all:.2f} ms", flush=True)
a, b = res["bf16"], res["fp8"]
for u in range(B):
    m = 0
    for x, y in zip(a[0][u], b[0][u]):
        if x != y: break
        m += 1
    print(f"utt {u}: bf16/fp8 greedy prefix match {m}/{len(a[0][u])}, distinct ids {len(set(a[0][u]))}")
d = np.abs(a[1] - b[1])
print("teacher-forced (bf16 stream) logits: max |d| =", float(d.max()), "row std", float(a[1].std(axis=-1).mean()),
      "top-2 margin min/median", float(np.min(np.sort(a[1], -1)[..., -1] - np.sort(a[1], -1)[..., -2])),
      float(np.median(np.sort(a[1], -1)[..., -1] - np.sort(a[1], -1)[..., -2])))

"""Decode-step timing probe on whisper-large-v3 (bf16): split-K tensor-core streaming kernel (decoder_stream.cu)
vs the previous streaming kernel (decoder_ring.cu): ms/step, per-phase times, token agreement.
    python tools/stream_probe.py [preset] [batch] [l2_hint]"""
import sys
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
toks = {}
modes = [("stream", 1, 1), ("stream-nohint", 1, 0)]
if B <= 4: modes.append(("ring", 0, 0))
for name, st, hint in modes:
    eng.set_option("stream", st); eng.set_option("stream_l2_hint", hint)
    eng.prefill(prompt, want_logits=False); eng.decode(max_steps=4)
    t_pre = timed(lambda: eng.prefill(prompt, want_logits=False), 3)
    eng.prefill(prompt, want_logits=False)
    t_dec = timed(lambda: eng.decode(max_steps=32)) / 32
    t_all = timed(lambda: eng.transcribe_resident(prompt, max_new=33), 3)
    toks[name] = eng.transcribe_resident(prompt, max_new=33)
    print(f"{name}: prefill {t_pre:.3f} ms, decode {t_dec:.4f} ms/step, transcribe {t_all:.2f} ms", flush=True)
if "ring" in toks:
    for b in range(B):
        m = 0
        for x, y in zip(toks["stream"][b], toks["ring"][b]):
            if x != y: break
            m += 1
        print(f"utt {b}: stream/ring greedy prefix match {m}/{len(toks['ring'][b])}", toks["stream"][b][:8])
eng.set_option("stream", 1); eng.set_option("stream_l2_hint", 1); eng.set_option("mega_timing", 1)
eng.prefill(prompt, want_logits=False)
eng.decode(max_steps=3)
t = eng.get_stage("mega_timing", 16384)
L = dims.dec_layers
per = 8 * L + 2
names = ["qkv", "self", "out", "cq", "cross", "cout", "fc1", "fc2"]
print("stamps", len(t), "step totals us:", [round(float(t[i * per:(i + 1) * per].sum()), 1) for i in range(len(t) // per)])
step = t[per:2 * per]
for j, nme in enumerate(names):
    print(f"  {nme:6s} mean {step[j:8 * L:8].mean():7.2f} us  min {step[j:8 * L:8].min():7.2f}  max {step[j:8 * L:8].max():7.2f}")
print(f"  head   {step[8 * L]:7.2f} us   exchange {step[8 * L + 1]:7.2f} us")
raw = eng.get_stage("mega_timing_raw", 2 * 16384).view(np.uint32).astype(np.uint64)
raw = (raw[0::2] | (raw[1::2] << np.uint64(32)))
for kind, nme in ((0, "self"), (1, "cross")):
    cyc, cnt = int(raw[16384 - 4 + kind * 2]), int(raw[16384 - 3 + kind * 2])
    n, late = cnt & 0xffffffff, cnt >> 32
    if n: print(f"  {nme}-attention K stages over 3 steps: {n} waits, {late} not ready on arrival, mean wait {cyc / n:.0f} cycles")

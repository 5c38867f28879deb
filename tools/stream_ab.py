"""A/B of streaming-decode-kernel variants selected by set_option("stream_debug", bits): ms/step each (two rounds).
    python tools/stream_ab.py [batch] [bits,bits,...]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from b200asr.config import PRESETS
from b200asr.engine import WhisperEngine
from b200asr.synth import synth_batch, synth_whisper_checkpoint
from b200asr.weights import fold_whisper

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
variants = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1]
dims = PRESETS["whisper-large-v3"]
raw = synth_whisper_checkpoint(dims, 20260)
tensors = fold_whisper(raw, dims, [1, 2, 7], [220, 50257])
del raw
eng = WhisperEngine(dims, tensors, precision="bf16", max_batch=B, max_samples=128000)
del tensors
prompt = [50258, 50259, 50360, 50364]
eng.set_decode_options(stop_ids=[], generate_limit=33)
eng.encode(synth_batch(B, 128000))
stream = torch.cuda.ExternalStream(eng.stream_ptr)

def timed(fn, n=1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(stream)
    for _ in range(n): fn()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

ref = None
for rep in range(2):
    for v in variants:
        eng.set_option("stream_debug", v)
        eng.prefill(prompt, want_logits=False); eng.decode(max_steps=4)
        eng.prefill(prompt, want_logits=False)
        t = timed(lambda: eng.decode(max_steps=32)) / 32
        toks = eng.transcribe_resident(prompt, max_new=12)
        if ref is None: ref = toks
        print(f"B={B} debug={v:3d}: decode {t:.4f} ms/step  tokens {'same' if toks == ref else 'DIFFER'}", flush=True)

"""whisper-large-v3 decode step: the lean instantiation of the streaming kernel (rarely used branches compiled out) vs the full one.
    python tools/lean_probe.py [batch]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from b200asr.config import PRESETS
from b200asr.engine import WhisperEngine
from b200asr.synth import synth_batch, synth_whisper_checkpoint
from b200asr.weights import fold_whisper

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dims = PRESETS["whisper-large-v3"]
tensors = fold_whisper(synth_whisper_checkpoint(dims, 20260, pos_scale=100.0), dims, [1, 2, 7], [220, 50257])
eng = WhisperEngine(dims, tensors, precision="bf16", max_batch=B, max_samples=128000)
del tensors
prompt = [50258, 50259, 50360, 50364]
eng.set_decode_options(stop_ids=[], generate_limit=33)
eng.upload_pcm(synth_batch(B, 128000))
stream = torch.cuda.ExternalStream(eng.stream_ptr)

def timed(fn, n=1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(stream)
    for _ in range(n): fn()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

eng.encode_resident()
res = {}
for lean in (0, 1, 0, 1):
    eng.set_option("stream_lean", lean)
    eng.prefill(prompt, want_logits=False); eng.decode(max_steps=4)
    eng.prefill(prompt, want_logits=False)
    t_dec = timed(lambda: eng.decode(max_steps=32)) / 32
    t_all = timed(lambda: eng.transcribe_resident(prompt, max_new=33), 3)
    res[lean] = eng.transcribe_resident(prompt, max_new=33)
    print(f"lean={lean}: decode {t_dec:.4f} ms/step, transcribe {t_all:.2f} ms", flush=True)
print("tokens equal:", res[0] == res[1])

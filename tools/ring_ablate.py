"""Timing ablation of the streaming decode kernel: ms/step with parts of a phase skipped (results are garbage
when a bit is set; only the time matters).  bits: 1 no exchange spin, 2 no LN stats, 4 no dot math, 8 no weight
streaming, 16 no attention phases."""
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
raw = synth_whisper_checkpoint(dims, 20260)
tensors = fold_whisper(raw, dims, [1, 2, 7], [220, 50257])
del raw
eng = WhisperEngine(dims, tensors, precision="bf16", max_batch=B, max_samples=128000)
del tensors
prompt = [50258, 50259, 50360, 50364]
eng.set_decode_options(stop_ids=[], generate_limit=33)
eng.upload_pcm(synth_batch(B, 128000))
eng.encode_resident()
stream = torch.cuda.ExternalStream(eng.stream_ptr)

def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(stream); fn(); e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1)

for dbg in [0, 8, 4 | 8, 1 | 2 | 4 | 8, 7]:
    eng.set_option("ring_debug", dbg)
    eng.prefill(prompt, want_logits=False); eng.decode(max_steps=4)
    eng.prefill(prompt, want_logits=False)
    t = timed(lambda: eng.decode(max_steps=32)) / 32
    print(f"debug={dbg:2d}: {t:.4f} ms/step", flush=True)

"""One profiled pass of the hot path (whisper-large-v3 bf16, batch 1, 8 s) for ncu:
    ncu --profile-from-start off ... python tools/profile_step.py
Only the pass between cudaProfilerStart/Stop is captured."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
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
eng.set_decode_options(stop_ids=[], generate_limit=33)
eng.upload_pcm(synth_batch(B, 128000))
eng.transcribe_resident(prompt, max_new=33)
torch.cuda.synchronize()
torch.cuda.profiler.start()
toks = eng.transcribe_resident(prompt, max_new=33)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("tokens", toks[0][:6], "launches", eng.kernel_launches)

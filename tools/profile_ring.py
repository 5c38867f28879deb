"""One profiled greedy-decode launch of the streaming kernel (whisper-large-v3 bf16) for ncu:
    ncu --profile-from-start off ... python tools/profile_ring.py [batch] [steps] [fp8]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from b200asr.config import PRESETS
from b200asr.engine import WhisperEngine
from b200asr.synth import synth_batch, synth_whisper_checkpoint
from b200asr.weights import fold_whisper

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
fp8 = int(sys.argv[3]) if len(sys.argv) > 3 else 0       # 1: E4M3 decoder weights (set_option("fp8", 1))
dims = PRESETS["whisper-large-v3"]
raw = synth_whisper_checkpoint(dims, 20260)
tensors = fold_whisper(raw, dims, [1, 2, 7], [220, 50257])
del raw
eng = WhisperEngine(dims, tensors, precision="bf16", max_batch=B, max_samples=128000)
del tensors
prompt = [50258, 50259, 50360, 50364]
eng.set_decode_options(stop_ids=[], generate_limit=64)
eng.set_option("fp8", fp8)
eng.upload_pcm(synth_batch(B, 128000))
eng.encode_resident()
eng.prefill(prompt, want_logits=False)
eng.decode(max_steps=4)
eng.prefill(prompt, want_logits=False)
torch.cuda.synchronize()
torch.cuda.profiler.start()
toks = eng.decode(max_steps=steps)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("tokens", toks[0][:8])

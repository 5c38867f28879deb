"""Time the Whisper-large-v3 4-token prefill through the grid-barrier kernel (mega=1) and through the per-op kernels (mega=0)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, __file__.rsplit("/", 2)[0])
from b200asr.config import WHISPER_LARGE_V3 as DIMS
from b200asr.engine import WhisperEngine
from b200asr.synth import synth_pcm, synth_whisper_checkpoint
from b200asr.weights import fold_whisper

PROMPT = [50258, 50259, 50360, 50364]
tensors = fold_whisper(synth_whisper_checkpoint(DIMS, 20260), DIMS, [1, 2, 7], [220, 50257])
eng = WhisperEngine(DIMS, tensors, precision="bf16", max_batch=1, max_samples=128000)
del tensors
pcm = synth_pcm(0, 128000)
eng.set_decode_options(stop_ids=[], generate_limit=0)
stream = torch.cuda.ExternalStream(eng.stream_ptr)
for mega in (1, 0):
    eng.set_option("mega", mega)
    ts = []
    for i in range(6):
        eng.encode(pcm)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(stream)
        lg, tok = eng.prefill(PROMPT, want_logits=False)
        e1.record(stream); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"mega={mega}: prefill {np.median(ts[2:]):.3f} ms (token {int(tok[0])})", flush=True)
eng.close()

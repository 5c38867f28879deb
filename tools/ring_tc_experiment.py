import sys
sys.path.insert(0, "/root/repo")
import torch
from b200asr.config import PRESETS
from b200asr.engine import WhisperEngine
from b200asr.synth import synth_batch, synth_whisper_checkpoint
from b200asr.weights import fold_whisper
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dims = PRESETS["whisper-large-v3"]
raw = synth_whisper_checkpoint(dims, 20260)
tensors = fold_whisper(raw, dims, [1, 2, 7], [220, 50257]); del raw
eng = WhisperEngine(dims, tensors, precision="bf16", max_batch=B, max_samples=128000); del tensors
prompt = [50258, 50259, 50360, 50364]
eng.set_decode_options(stop_ids=[], generate_limit=33)
eng.upload_pcm(synth_batch(B, 128000)); eng.encode_resident()
stream = torch.cuda.ExternalStream(eng.stream_ptr)
def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(stream); fn(); e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for name, tc, dbg in (("tc", 1, 0), ("tc single-copy (garbage)", 1, 32), ("simt", 0, 0), ("tc", 1, 0)):
    eng.set_option("ring_tc", tc); eng.set_option("ring_debug", dbg)
    eng.prefill(prompt, want_logits=False); eng.decode(max_steps=4)
    eng.prefill(prompt, want_logits=False)
    print(f"B={B} {name}: {timed(lambda: eng.decode(max_steps=32)) / 32:.4f} ms/step", flush=True)

import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from b200asr import qwen as qw
from b200asr.synth import synth_batch
dims = qw.QWEN3_ASR_0_6B
prompt = qw.QwenPrompt(qw.QWEN3_PROMPT.head_ids, qw.QWEN3_PROMPT.suffix_ids, qw.QWEN3_PROMPT.tail_ids, ())
tensors = qw.fold_qwen(qw.synth_qwen_checkpoint(dims, 1), dims)
for B in (4, 8):
    eng = qw.QwenEngine(dims, tensors, prompt, precision="bf16", max_batch=B, max_samples=480000)
    pcm = synth_batch(B, 480000)
    eng.upload(pcm)
    for pdl in (1, 2, 0):
        eng.set_option("pdl", pdl)
        for _ in range(2): eng.transcribe_resident(max_new=128)
        torch.cuda.synchronize(); t = time.time()
        for _ in range(3): eng.transcribe_resident(max_new=128)
        torch.cuda.synchronize(); dt = (time.time() - t) / 3
        print(f"B={B} pdl={pdl}: {dt*1e3:.1f} ms per batch", flush=True)
    eng.close()

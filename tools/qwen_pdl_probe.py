"""Qwen3-ASR-0.6B decode step at 3-4 clips: programmatic dependent launch for every decode-step kernel ("pdl" = 2) against
the default rule (dependents only at 1-2 clips)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from b200asr import qwen as qw
from b200asr.synth import synth_pcm

dims = qw.QWEN3_ASR_0_6B
prompt = qw.QwenPrompt(qw.QWEN3_PROMPT.head_ids, qw.QWEN3_PROMPT.suffix_ids, qw.QWEN3_PROMPT.tail_ids, ())
tensors = qw.fold_qwen(qw.synth_qwen_checkpoint(dims, 20261), dims)
for nb in (3, 4, 8):
    eng = qw.QwenEngine(dims, tensors, prompt, precision="bf16", max_batch=nb, max_samples=480000)
    pcm = np.stack([synth_pcm(10 + i, 480000) for i in range(nb)])
    eng.upload(pcm)
    out = {}
    for pdl in (1, 2, 1, 2):
        eng.set_option("pdl", pdl)
        ts = {}
        for mx in (8, 128):
            eng.transcribe_resident(max_new=mx)
            best = 1e9
            for _ in range(3):
                t0 = time.perf_counter(); toks = eng.transcribe_resident(max_new=mx); best = min(best, time.perf_counter() - t0)
            ts[mx] = best
        out[pdl] = toks
        print(f"batch {nb} pdl {pdl}: {ts[128] * 1e3:.2f} ms / 128 tokens, {(ts[128] - ts[8]) / 120 * 1e3:.4f} ms per step", flush=True)
    print(f"batch {nb}: streams equal: {out[1] == out[2]}")
    eng.close()

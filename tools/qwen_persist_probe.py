"""Qwen3-ASR-0.6B decode step: persistent decode-layer kernel (option "persist") against the per-launch CUDA graph."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from b200asr import qwen as qw
from b200asr.synth import synth_pcm

dims = qw.QWEN3_ASR_0_6B
prompt = qw.QwenPrompt(qw.QWEN3_PROMPT.head_ids, qw.QWEN3_PROMPT.suffix_ids, qw.QWEN3_PROMPT.tail_ids, ())
tensors = qw.fold_qwen(qw.synth_qwen_checkpoint(dims, 20261), dims)
for nb in (1, 2, 4):
    eng = qw.QwenEngine(dims, tensors, prompt, precision="bf16", max_batch=nb, max_samples=480000)
    pcm = np.stack([synth_pcm(10 + i, 480000) for i in range(nb)])
    eng.upload(pcm)
    out = {}
    for persist, dbg in ((0, 0), (1, 0), (2, 0), (1, 3)):
        eng.set_option("persist", persist); eng.set_option("persist_dbg", dbg)
        ts = {}
        for mx in (8, 128):
            eng.transcribe_resident(max_new=mx)
            best = 1e9
            for _ in range(3):
                t0 = time.perf_counter(); toks = eng.transcribe_resident(max_new=mx); best = min(best, time.perf_counter() - t0)
            ts[mx] = best
        out[persist] = toks
        l0 = eng.kernel_launches; eng.transcribe_resident(max_new=8); nl = eng.kernel_launches - l0
        print(f"batch {nb} persist {persist} dbg {dbg} ({nl} launches / 8 tokens): {ts[128] * 1e3:.2f} ms / 128 tokens, {(ts[128] - ts[8]) / 120 * 1e3:.4f} ms per step", flush=True)
    same = sum(a == b for a, b in zip(out[0], out[1]))
    print(f"batch {nb}: greedy streams equal for {same}/{nb} clips; first diff at", [next((i for i, (x, y) in enumerate(zip(a, b)) if x != y), -1) for a, b in zip(out[0], out[1])])
    eng.close()

"""Where a phase of the persistent Qwen3-ASR decode-layer kernel spends its time: block 0's %globaltimer stamps."""
import sys
sys.path.insert(0, ".")
import numpy as np
from b200asr import qwen as qw
from b200asr.synth import synth_pcm

dims = qw.QWEN3_ASR_0_6B
prompt = qw.QwenPrompt(qw.QWEN3_PROMPT.head_ids, qw.QWEN3_PROMPT.suffix_ids, qw.QWEN3_PROMPT.tail_ids, ())
tensors = qw.fold_qwen(qw.synth_qwen_checkpoint(dims, 20261), dims)
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 1
eng = qw.QwenEngine(dims, tensors, prompt, precision="bf16", max_batch=nb, max_samples=480000)
eng.set_option("persist_timing", 1)
pcm = np.stack([synth_pcm(10 + i, 480000) for i in range(nb)])
eng.transcribe(pcm, max_new=16)
t = eng.get_stage("persist_timing", 1024).view(np.uint64).astype(np.int64)
t = t - t[0]
# per layer: 4 linear phases x 4 stamps (entry, waited, staged, math done) + attention 3 stamps (entry, waited, done) = 19
names = ["qkv"] * 4 + ["attn"] * 3 + ["o"] * 4 + ["gate_up"] * 4 + ["down"] * 4
lab = ["entry", "waited", "staged", "done"] * 1 + ["entry", "waited", "done"] + ["entry", "waited", "staged", "done"] * 3
for l in (0, 1, 2, 25):
    base = l * 19
    print(f"layer {l}: " + "  ".join(f"{names[i]}.{lab[i]}={int(t[base + i] - t[base])}" for i in range(19)))
print("layer period (ns):", [int(t[(l + 1) * 19] - t[l * 19]) for l in range(0, 25, 5)])

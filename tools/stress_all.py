"""Repeat-run stress of the other engine paths (every result must equal the first): Whisper ragged batch, FP8 weights, the step
API, SenseVoice, Paraformer, Qwen3-ASR.   python tools/stress_all.py [iters]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from b200asr.config import PRESETS
from b200asr.engine import WhisperEngine
from b200asr.synth import synth_batch, synth_pcm, synth_whisper_checkpoint
from b200asr.weights import fold_whisper

N = int(sys.argv[1]) if len(sys.argv) > 1 else 300


def repeat(name, fn, n):
    t0 = time.time()
    first = fn()
    for i in range(n):
        if fn() != first:
            print(f"{name}: MISMATCH at {i}", flush=True); sys.exit(2)
    print(f"{name}: {n} runs ok, {1e3 * (time.time() - t0) / (n + 1):.2f} ms each", flush=True)


dims = PRESETS["whisper-large-v3"]
tensors = fold_whisper(synth_whisper_checkpoint(dims, 20260, pos_scale=100.0), dims, [1, 2, 7], [220, 50257])
eng = WhisperEngine(dims, tensors, precision="bf16", max_batch=4, max_samples=128000)
del tensors
prompt = [50258, 50259, 50360, 50364]
eng.set_decode_options(stop_ids=[], generate_limit=33)
clips = [synth_pcm(i, n) for i, n in enumerate((128000, 96000, 64160, 112000))]
pcm, lens = WhisperEngine.pad_ragged(clips)
repeat("whisper ragged batch 4", lambda: eng.transcribe(pcm, prompt, max_new=33, lens=lens), N)
eng.set_option("fp8", 1)
one = synth_batch(1, 128000)
repeat("whisper fp8 batch 1", lambda: eng.transcribe(one, prompt, max_new=33), N)
four = synth_batch(4, 128000)
repeat("whisper fp8 batch 4", lambda: eng.transcribe(four, prompt, max_new=33), N // 2)
eng.set_option("fp8", 0)


def step_api():
    eng.encode(one)
    _, tok = eng.prefill(prompt)
    out = [int(tok[0])]
    for _ in range(12):
        _, tok = eng.decode_step(want_logits=False)
        out.append(int(tok[0]))
    return out


repeat("whisper step api", step_api, N // 3)
eng.set_decode_options(stop_ids=[], generate_limit=33, repeat_penalty=0.8, penalty_range=5)
repeat("whisper penalty-greedy batch 4", lambda: eng.transcribe(four, prompt, max_new=33), N // 2)
eng.close()

from b200asr import paraformer as pf, sensevoice as sv
rng = np.random.default_rng(0)
audio = (rng.standard_normal((8, 128000)) * 2500).astype(np.int16)
D = sv.SENSEVOICE_SMALL
se = sv.SenseVoiceEngine(D, sv.fold_sensevoice(sv.synth_sensevoice_checkpoint(D, 0), D, 128000), precision="bf16", max_batch=8, max_samples=128000)
repeat("sensevoice batch 8", lambda: se.run(audio, 0), N)
se.close()
P = pf.PARAFORMER_LARGE
pe = pf.ParaformerEngine(P, pf.fold_paraformer(pf.synth_paraformer_checkpoint(P, 0), P, 128000), precision="bf16", max_batch=8, max_samples=128000)
repeat("paraformer batch 8", lambda: pe.run(audio), N)
pe.close()
from b200asr import qwen as qw
qd = qw.QWEN3_ASR_0_6B
qprompt = qw.QwenPrompt(qw.QWEN3_PROMPT.head_ids, qw.QWEN3_PROMPT.suffix_ids, qw.QWEN3_PROMPT.tail_ids, ())
qe = qw.QwenEngine(qd, qw.fold_qwen(qw.synth_qwen_checkpoint(qd, 0), qd), qprompt, precision="bf16", max_batch=2, max_samples=480000)
clip = synth_batch(2, 480000)
repeat("qwen3-asr batch 2", lambda: qe.transcribe(clip, (), (), max_new=64), max(10, N // 10))
qe.close()

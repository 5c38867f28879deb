"""Encoder / non-autoregressive forward time with programmatic dependent launch on and off (GEMMs + LayerNorm kernels).
    python tools/pdl_probe.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from b200asr.config import PRESETS
from b200asr.engine import WhisperEngine
from b200asr.synth import synth_batch, synth_whisper_checkpoint
from b200asr.weights import fold_whisper


def timed(stream, fn, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(stream)
    for _ in range(n): fn()
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


dims = PRESETS["whisper-large-v3"]
tensors = fold_whisper(synth_whisper_checkpoint(dims, 20260), dims, [1, 2, 7], [220, 50257])
for B in (1, 4):
    eng = WhisperEngine(dims, tensors, precision="bf16", max_batch=B, max_samples=128000)
    stream = torch.cuda.ExternalStream(eng.stream_ptr)
    eng.upload_pcm(synth_batch(B, 128000))
    for pdl in (1, 0, 1):
        eng.set_option("pdl", pdl)
        for _ in range(3): eng.encode_resident()
        print(f"whisper encoder batch {B} pdl={pdl}: {timed(stream, eng.encode_resident, 10):.3f} ms", flush=True)
    eng.close()
del tensors
from b200asr import paraformer as pf, sensevoice as sv
for name, mod, D, mk in (("sensevoice", sv, sv.SENSEVOICE_SMALL, lambda d: sv.SenseVoiceEngine), ("paraformer", pf, pf.PARAFORMER_LARGE, lambda d: pf.ParaformerEngine)):
    raw = (sv.synth_sensevoice_checkpoint if name == "sensevoice" else pf.synth_paraformer_checkpoint)(D, 0)
    fold = (sv.fold_sensevoice if name == "sensevoice" else pf.fold_paraformer)(raw, D, 128000)
    for B in (1, 8):
        eng = mk(D)(D, fold, precision="bf16", max_batch=B, max_samples=128000)
        stream = torch.cuda.ExternalStream(eng.stream_ptr)
        rng = np.random.default_rng(0)
        pcm = (rng.standard_normal((B, 128000)) * 2500).astype(np.int16)
        eng.upload(pcm, 0)
        for pdl in (1, 0, 1):
            eng.set_option("pdl", pdl)
            for _ in range(3): eng.run_resident()
            print(f"{name} batch {B} pdl={pdl}: {timed(stream, eng.run_resident, 10):.3f} ms", flush=True)
        eng.close()

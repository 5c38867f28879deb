"""Shared helpers for the -m gpu parity tests (CUDA engine vs the CPU oracle / goldens)."""
from pathlib import Path

import numpy as np

from oracle import whisper_oracle as wo
from b200asr.config import WHISPER_TINY_TEST
from b200asr.weights import fold_whisper
from b200asr.engine import WhisperEngine

GOLD = sorted((Path(__file__).parent / "golden").glob("whisper_tiny_case*.npz"))
NO_SPEECH = 13


def load_case(path):
    g = dict(np.load(path))
    raw = wo.make_raw_weights(wo.TINY_TEST, int(g["seed"]))
    tensors = fold_whisper(raw, WHISPER_TINY_TEST, g["suppress"].tolist(), g["begin_suppress"].tolist())
    return g, raw, tensors


def make_engine(tensors, precision, max_batch=1, max_samples=32000, tc=True):
    return WhisperEngine(WHISPER_TINY_TEST, tensors, precision=precision, max_batch=max_batch,
                         max_samples=max_samples, use_tensor_cores=tc)


def maxdiff(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))))


def dequantised_e4m3(tensors):
    """The engine's quantiser (csrc/decoder_stream.cu: quant_rows_e4m3_kernel) replayed on the host: bf16 weights, row scale =
    amax / 448, E4M3 round-to-nearest-even of w * (1 / scale); returns the tensors with the decoder matrices replaced by
    scale * e4m3(...)."""
    import torch
    out = dict(tensors)
    for name, w in tensors.items():
        if not (name == "dec.embed" or (name.startswith("dec.L") and name.endswith(".w"))):
            continue
        t = torch.from_numpy(np.ascontiguousarray(w, np.float32))
        shape = t.shape
        t = t.reshape(shape[0], -1).to(torch.bfloat16).to(torch.float32)
        amax = t.abs().amax(dim=1, keepdim=True)
        sc = torch.where(amax > 0, amax / 448.0, torch.ones_like(amax))
        q = (t * (1.0 / sc)).to(torch.float8_e4m3fn).to(torch.float32)
        out[name] = (q * sc).reshape(shape).numpy()
    return out

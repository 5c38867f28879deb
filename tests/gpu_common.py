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

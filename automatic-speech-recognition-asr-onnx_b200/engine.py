"""Python face of the engine: a thin, typed wrapper over the C ABI.

Host buffers are numpy arrays (or pinned torch tensors viewed as numpy); all
device memory and all compute live inside libb200asr.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np

from . import _cabi
from .config import WhisperDims


class B200AsrError(RuntimeError):
    pass


def _i32p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_cabi._I32P)


def _f32p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_cabi._F32P)


class WhisperEngine:
    """One engine per GPU.  Mirrors the three ORT sessions of the reference driver
    (/root/reference/Whisper/Inference_Whisper_ONNX.py:312-315): encode = probe
    session's encoder half, prefill, decode."""

    def __init__(self, dims: WhisperDims, tensors: Dict[str, np.ndarray], *, precision: str = "bf16",
                 max_batch: int = 1, max_samples: int = 480000, device: int = 0, use_tensor_cores: bool = True):
        self.lib = _cabi.load()
        self.dims = dims
        self.precision = precision
        cfg = _cabi.Config(
            n_mels=dims.n_mels, d_model=dims.d_model, n_heads=dims.n_heads, ffn=dims.ffn,
            enc_layers=dims.enc_layers, dec_layers=dims.dec_layers, vocab=dims.vocab, max_source=dims.max_source,
            max_target=dims.max_target, n_fft=dims.n_fft, hop=dims.hop, max_batch=max_batch, max_samples=max_samples,
            precision={"f32": _cabi.PRECISION_F32, "bf16": _cabi.PRECISION_BF16}[precision], device=device,
            use_tensor_cores=1 if use_tensor_cores else 0)
        h = C.c_void_p()
        rc = self.lib.b200asr_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise B200AsrError(f"b200asr_create failed ({rc}): {self.lib.b200asr_last_error(None).decode()}")
        self.h = h
        self.max_batch = max_batch
        self.batch = 0
        self.T_enc = 0                         # encoder positions of the clips currently resident
        for name, arr in tensors.items():
            a = np.ascontiguousarray(arr, dtype=np.float32)
            self._ck(self.lib.b200asr_set_tensor(self.h, name.encode(), _f32p(a), a.size))
        self._ck(self.lib.b200asr_finalize_weights(self.h))

    # -- plumbing -----------------------------------------------------------
    def _ck(self, rc: int):
        if rc != 0:
            raise B200AsrError(f"b200asr error {rc}: {self.lib.b200asr_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200asr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream_ptr(self) -> int:
        return int(self.lib.b200asr_stream(self.h) or 0)

    @property
    def kernel_launches(self) -> int:
        return int(self.lib.b200asr_kernel_launches(self.h))

    def synchronize(self):
        self._ck(self.lib.b200asr_synchronize(self.h))

    def set_option(self, key: str, value: int):
        self._ck(self.lib.b200asr_set_option(self.h, key.encode(), int(value)))

    @staticmethod
    def _pcm(pcm: np.ndarray):
        pcm = np.ascontiguousarray(pcm)
        if pcm.ndim == 1:
            pcm = pcm[None]
        if pcm.ndim == 3:                      # reference layout [B, 1, N]
            pcm = pcm.reshape(pcm.shape[0], pcm.shape[-1])
        if pcm.dtype == np.int16:
            code = _cabi.PCM_I16
        elif pcm.dtype == np.float32:
            code = _cabi.PCM_F32
        else:
            raise TypeError(f"PCM dtype must be int16 or float32, got {pcm.dtype}")
        return pcm, code

    # -- encoder ------------------------------------------------------------
    def _lens(self, lens, pcm: np.ndarray) -> np.ndarray:
        lens = np.ascontiguousarray(np.asarray(lens, np.int32).reshape(-1))
        if lens.shape[0] != pcm.shape[0]:
            raise ValueError(f"lens has {lens.shape[0]} entries for a batch of {pcm.shape[0]}")
        return lens

    @staticmethod
    def pad_ragged(clips: Sequence[np.ndarray]):
        """Clips of different lengths -> ([B][longest] array, lens) for the `lens=` arguments below."""
        lens = np.asarray([len(c) for c in clips], np.int32)
        out = np.zeros((len(clips), int(lens.max())), np.asarray(clips[0]).dtype)
        for b, c in enumerate(clips):
            out[b, :len(c)] = c
        return out, lens

    def valid_positions(self, lens) -> np.ndarray:
        """Encoder positions of each clip of a ragged batch (rows beyond are padding)."""
        return (np.asarray(lens, np.int64) // self.dims.hop + 1) // 2

    def encode(self, pcm: np.ndarray, lens=None):
        """lens: samples per clip for a ragged batch (pcm rows zero-padded to the longest clip)."""
        pcm, code = self._pcm(pcm)
        self.batch = pcm.shape[0]
        self.T_enc = (pcm.shape[1] // self.dims.hop + 1) // 2
        if lens is None:
            self._ck(self.lib.b200asr_encode(self.h, pcm.ctypes.data_as(C.c_void_p), code, pcm.shape[0], pcm.shape[1]))
        else:
            self._ck(self.lib.b200asr_encode_ragged(self.h, pcm.ctypes.data_as(C.c_void_p), code, pcm.shape[0], pcm.shape[1],
                                                    _i32p(self._lens(lens, pcm))))

    def upload_pcm(self, pcm: np.ndarray, lens=None):
        pcm, code = self._pcm(pcm)
        self.batch = pcm.shape[0]
        self.T_enc = (pcm.shape[1] // self.dims.hop + 1) // 2
        if lens is None:
            self._ck(self.lib.b200asr_upload_pcm(self.h, pcm.ctypes.data_as(C.c_void_p), code, pcm.shape[0], pcm.shape[1]))
        else:
            self._ck(self.lib.b200asr_upload_pcm_ragged(self.h, pcm.ctypes.data_as(C.c_void_p), code, pcm.shape[0],
                                                        pcm.shape[1], _i32p(self._lens(lens, pcm))))

    def encode_resident(self):
        self._ck(self.lib.b200asr_encode_resident(self.h))

    # -- decoder ------------------------------------------------------------
    def set_decode_options(self, stop_ids: Sequence[int] = (), generate_limit: int = 0, repeat_penalty: float = 1.0,
                           penalty_range: int = 20):
        s = np.asarray(list(stop_ids), dtype=np.int32)
        self._ck(self.lib.b200asr_set_decode_options(self.h, _i32p(s) if s.size else None, s.size, generate_limit,
                                                     float(repeat_penalty), penalty_range))

    def set_sampling(self, temperature: float = 0.0, top_k: int = 10, top_p: float = 0.95, repetition_penalty: float = 1.0,
                     seed: int = 0, noise: Optional[np.ndarray] = None):
        """TOPK_TOPP_SAMPLING head (Whisper/Export_Whisper.py:263-307); temperature <= 0 = argmax heads.
        noise: optional uniform numbers [launches][max_batch][top_k] for reproducible runs."""
        n = None
        rows = 0
        if noise is not None:
            n = np.ascontiguousarray(noise, dtype=np.float32).reshape(-1, self.max_batch, top_k)
            rows = n.shape[0]
        self._ck(self.lib.b200asr_set_sampling(self.h, float(temperature), int(top_k), float(top_p),
                                               float(repetition_penalty), int(seed), _f32p(n), rows))

    def _prompt(self, prompt) -> np.ndarray:
        p = np.asarray(prompt, dtype=np.int32)
        if p.ndim == 1:
            p = np.tile(p[None], (self.batch, 1))
        if p.shape[0] != self.batch:
            raise ValueError("prompt batch mismatch")
        return np.ascontiguousarray(p)

    def prefill(self, prompt, want_logits: bool = True):
        p = self._prompt(prompt)
        logits = np.empty((self.batch, self.dims.vocab), np.float32) if want_logits else None
        first = np.empty(self.batch, np.int32)
        self._ck(self.lib.b200asr_prefill(self.h, _i32p(p), p.shape[1], _f32p(logits), _i32p(first)))
        return logits, first

    def decode_step(self, token_in=None, want_logits: bool = True):
        t = None if token_in is None else np.ascontiguousarray(np.asarray(token_in, dtype=np.int32).reshape(self.batch))
        logits = np.empty((self.batch, self.dims.vocab), np.float32) if want_logits else None
        tok = np.empty(self.batch, np.int32)
        self._ck(self.lib.b200asr_decode_step(self.h, _i32p(t), _f32p(logits), _i32p(tok)))
        return logits, tok

    def decode(self, max_steps: int = -1):
        ld = self.dims.max_target
        toks = np.zeros((self.batch, ld), np.int32)
        lens = np.zeros(self.batch, np.int32)
        self._ck(self.lib.b200asr_decode(self.h, max_steps, _i32p(toks), ld, _i32p(lens)))
        return [toks[b, :lens[b]].tolist() for b in range(self.batch)]

    def no_speech_prob(self, no_speech_token: int) -> np.ndarray:
        out = np.empty(self.batch, np.float32)
        self._ck(self.lib.b200asr_no_speech_prob(self.h, no_speech_token, _f32p(out)))
        return out

    # -- whole path ---------------------------------------------------------
    def transcribe(self, pcm: np.ndarray, prompt, max_new: int = 0, out_tokens: Optional[np.ndarray] = None,
                   out_lens: Optional[np.ndarray] = None, lens=None):
        pcm, code = self._pcm(pcm)
        self.batch = pcm.shape[0]
        self.T_enc = (pcm.shape[1] // self.dims.hop + 1) // 2
        p = self._prompt(prompt)
        ld = self.dims.max_target
        toks = out_tokens if out_tokens is not None else np.zeros((self.batch, ld), np.int32)
        nout = out_lens if out_lens is not None else np.zeros(self.batch, np.int32)
        if lens is None:
            self._ck(self.lib.b200asr_transcribe(self.h, pcm.ctypes.data_as(C.c_void_p), code, pcm.shape[0], pcm.shape[1],
                                                 _i32p(p), p.shape[1], max_new, _i32p(toks), toks.shape[1], _i32p(nout)))
        else:
            self._ck(self.lib.b200asr_transcribe_ragged(self.h, pcm.ctypes.data_as(C.c_void_p), code, pcm.shape[0], pcm.shape[1],
                                                        _i32p(self._lens(lens, pcm)), _i32p(p), p.shape[1], max_new,
                                                        _i32p(toks), toks.shape[1], _i32p(nout)))
        return [toks[b, :nout[b]].tolist() for b in range(self.batch)]

    def transcribe_resident(self, prompt, max_new: int = 0):
        p = self._prompt(prompt)
        ld = self.dims.max_target
        toks = np.zeros((self.batch, ld), np.int32)
        lens = np.zeros(self.batch, np.int32)
        self._ck(self.lib.b200asr_transcribe_resident(self.h, _i32p(p), p.shape[1], max_new, _i32p(toks), ld, _i32p(lens)))
        return [toks[b, :lens[b]].tolist() for b in range(self.batch)]

    # -- introspection ------------------------------------------------------
    def get_stage(self, name: str, capacity: int) -> np.ndarray:
        out = np.empty(capacity, np.float32)
        n = C.c_int64(0)
        self._ck(self.lib.b200asr_get_stage(self.h, name.encode(), _f32p(out), capacity, C.byref(n)))
        return out[:n.value]


def test_gemm(M: int, N: int, K: int, A: np.ndarray, B: np.ndarray, bias=None, residual=None, act: int = 0,
              impl: str = "tc", device: int = 0) -> np.ndarray:
    lib = _cabi.load()
    A = np.ascontiguousarray(A, np.float32); B = np.ascontiguousarray(B, np.float32)
    bias = None if bias is None else np.ascontiguousarray(bias, np.float32)
    residual = None if residual is None else np.ascontiguousarray(residual, np.float32)
    out = np.empty((M, N), np.float32)
    err = C.create_string_buffer(512)
    rc = lib.b200asr_test_gemm(device, 1 if impl == "tc" else 0, M, N, K, _f32p(A), _f32p(B), _f32p(bias), _f32p(residual),
                               act, _f32p(out), err, 512)
    if rc != 0:
        raise B200AsrError(f"test_gemm failed ({rc}): {err.value.decode()}")
    return out

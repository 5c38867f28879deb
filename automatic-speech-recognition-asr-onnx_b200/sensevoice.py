"""SenseVoiceSmall on the B200 engine: weight folds, the ctypes face of the non-autoregressive C ABI and the host
loop of the reference driver (/root/reference/SenseVoice/Inference_SenseVoice_ONNX.py).

Folds follow the exporter (/root/reference/SenseVoice/Export_SenseVoice.py): front-end constants :139-169, prompt /
position tables :171-206 (fp16-rounded language embeddings and sinusoids, sqrt(d)-scaled embeddings and CMVN scale
:361-364), SANM folds :208-220 (d_head^-0.25 on the q and k rows, +1 on the FSMN centre tap, linear_out's bias moved
onto the FSMN conv).  Nothing here imports oracle/.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, asdict
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _cabi
from .engine import B200AsrError

LANGUAGE_PROFILES = (            # (code, name, aliases, prompt token id); selector index = row (Export_SenseVoice.py:38-50)
    ("auto", "Automatic language detection", ["automatic", "detect"], 0),
    ("zh", "Chinese", ["Chinese", "Mandarin", "zh-CN", "中文"], 3),
    ("en", "English", ["English", "en-US"], 4),
    ("yue", "Cantonese", ["Cantonese", "zh-yue", "粤语", "粵語"], 7),
    ("ja", "Japanese", ["Japanese", "jp", "日本語"], 11),
    ("ko", "Korean", ["Korean", "kr", "한국어"], 12),
    ("nospeech", "No speech", ["no-speech", "silence"], 13),
)
LANGUAGE_PROMPT_TOKEN_IDS = tuple(p[3] for p in LANGUAGE_PROFILES)
SYSTEM_PROMPT_IDS = (1, 2, 14)   # use_emo=True (:172)


@dataclass(frozen=True)
class SenseVoiceDims:
    n_mels: int = 80
    nfft: int = 512
    win: int = 400
    hop: int = 160
    lfr_m: int = 7
    lfr_n: int = 6
    d_model: int = 512
    n_heads: int = 4
    ffn: int = 2048
    n_blocks0: int = 1
    n_blocks: int = 49
    n_tp_blocks: int = 20
    vocab: int = 25055
    blank_id: int = 0
    fsmn_kernel: int = 11
    n_embed: int = 16
    ln_eps: float = 1e-12
    pre_emphasis: float = 0.97
    sample_rate: int = 16000

    @property
    def feat(self) -> int:
        return self.n_mels * self.lfr_m

    @property
    def head_dim(self) -> int:
        return self.d_model // self.n_heads

    @property
    def total_blocks(self) -> int:
        return self.n_blocks0 + self.n_blocks + self.n_tp_blocks

    def frames(self, n_samples: int) -> int:
        return (n_samples - self.win) // self.hop + 1

    def lfr_frames(self, n_samples: int) -> int:
        return (self.frames(n_samples) + self.lfr_n - 1) // self.lfr_n

    def to_dict(self):
        return asdict(self)


SENSEVOICE_SMALL = SenseVoiceDims()
SENSEVOICE_TINY_TEST = SenseVoiceDims(d_model=128, n_heads=2, ffn=256, n_blocks0=1, n_blocks=2, n_tp_blocks=1, vocab=300)
PRESETS = {"sensevoice-small": SENSEVOICE_SMALL, "sensevoice-tiny-test": SENSEVOICE_TINY_TEST}


def synth_sensevoice_checkpoint(d: SenseVoiceDims, seed: int) -> Dict[str, torch.Tensor]:
    """Seeded random checkpoint (no real weights offline); same draw order as the test oracle's generator so parity
    tests can build both sides from one seed."""
    g = torch.Generator().manual_seed(seed)
    raw: Dict[str, torch.Tensor] = {}

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    raw["embed"] = rn(d.n_embed, d.feat, std=0.5)
    raw["cmvn_means"] = rn(d.feat, std=1.0) - 8.0
    raw["cmvn_vars"] = 0.1 + 0.05 * torch.rand(d.feat, generator=g)
    for i in range(d.total_blocks):
        din = d.feat if i == 0 else d.d_model
        p = f"blk{i}."
        raw[p + "norm1.g"] = 1.0 + rn(din, std=0.1); raw[p + "norm1.b"] = rn(din, std=0.1)
        raw[p + "qkv.w"] = rn(3 * d.d_model, din, std=din ** -0.5); raw[p + "qkv.b"] = rn(3 * d.d_model, std=0.1)
        raw[p + "fsmn.w"] = rn(d.d_model, d.fsmn_kernel, std=0.2)
        raw[p + "out.w"] = rn(d.d_model, d.d_model, std=d.d_model ** -0.5); raw[p + "out.b"] = rn(d.d_model, std=0.1)
        raw[p + "norm2.g"] = 1.0 + rn(d.d_model, std=0.1); raw[p + "norm2.b"] = rn(d.d_model, std=0.1)
        raw[p + "w1.w"] = rn(d.ffn, d.d_model, std=d.d_model ** -0.5); raw[p + "w1.b"] = rn(d.ffn, std=0.1)
        raw[p + "w2.w"] = rn(d.d_model, d.ffn, std=d.ffn ** -0.5); raw[p + "w2.b"] = rn(d.d_model, std=0.1)
    for n in ("after_norm", "tp_norm"):
        raw[n + ".g"] = 1.0 + rn(d.d_model, std=0.1); raw[n + ".b"] = rn(d.d_model, std=0.1)
    raw["ctc.w"] = rn(d.vocab, d.d_model, std=d.d_model ** -0.5 * 3.0); raw["ctc.b"] = rn(d.vocab, std=0.5)
    return raw


def _fbank_kernel(d: SenseVoiceDims) -> torch.Tensor:
    F = d.nfft // 2 + 1
    window = torch.hamming_window(d.win, periodic=False, alpha=0.54, beta=0.46, dtype=torch.float32)
    omega = (2.0 * torch.pi / d.nfft) * torch.arange(F, dtype=torch.float32).unsqueeze(1) * torch.arange(d.win, dtype=torch.float32).unsqueeze(0)

    def fold(basis):      # pre-emphasis with replicate boundary, then per-frame DC removal (:154-158)
        nxt = torch.cat([basis[:, 1:], torch.zeros_like(basis[:, :1])], dim=1)
        out = basis - d.pre_emphasis * nxt
        out[:, 0] = out[:, 0] - d.pre_emphasis * basis[:, 0]
        return out - out.mean(dim=1, keepdim=True)

    return torch.cat([fold(torch.cos(omega) * window), fold(-torch.sin(omega) * window)], dim=0).contiguous()


def _mel_filters(d: SenseVoiceDims) -> torch.Tensor:
    import torchaudio.compliance.kaldi as kaldi
    banks, _ = kaldi.get_mel_banks(d.n_mels, d.nfft, float(d.sample_rate), 20.0, 0.0, 100.0, -500.0, 1.0)
    return torch.nn.functional.pad(banks, (0, 1), value=0.0).transpose(0, 1).contiguous()


def _position_table(n_pos: int, d: SenseVoiceDims) -> torch.Tensor:
    feat = d.feat
    inc = torch.log(torch.tensor([10000.0], dtype=torch.float32)) / (feat / 2 - 1)
    inv = torch.exp(torch.arange(feat / 2, dtype=torch.float32) * (-inc)).reshape(1, -1)
    st = torch.arange(1, n_pos + 1, dtype=torch.float32).reshape(-1, 1) * inv
    return torch.cat([torch.sin(st), torch.cos(st)], dim=1).half().float()


def fold_sensevoice(raw: Dict[str, torch.Tensor], d: SenseVoiceDims, max_samples: int) -> Dict[str, np.ndarray]:
    """Checkpoint tensors -> engine tensors (fp32 numpy), named as include/b200asr.h lists them."""
    out: Dict[str, torch.Tensor] = {}
    scale = float(d.d_model) ** 0.5
    embed = raw["embed"] * scale
    n_prompt = 1 + len(SYSTEM_PROMPT_IDS)
    max_lfr = d.lfr_frames(max_samples)
    pos = _position_table(max_lfr + n_prompt, d)
    out["fbank_kernel"] = _fbank_kernel(d)
    out["mel_filters"] = _mel_filters(d)
    out["language_embed"] = embed[list(LANGUAGE_PROMPT_TOKEN_IDS)].half().float() + pos[:1]
    out["system_embed"] = embed[list(SYSTEM_PROMPT_IDS)] + pos[1:n_prompt]
    out["cmvn_means"] = raw["cmvn_means"]
    out["cmvn_vars"] = raw["cmvn_vars"] * scale
    out["speech_position"] = pos[n_prompt:]
    f = float(d.head_dim) ** -0.25
    centre = (d.fsmn_kernel - 1) // 2
    for i in range(d.total_blocks):
        p = f"blk{i}."
        for k in ("norm1.g", "norm1.b", "norm2.g", "norm2.b", "w1.w", "w1.b", "w2.w", "w2.b", "out.w"):
            out[p + k] = raw[p + k]
        w = raw[p + "qkv.w"].clone(); b = raw[p + "qkv.b"].clone()
        w[:-d.d_model] *= f; b[:-d.d_model] *= f
        out[p + "qkv.w"], out[p + "qkv.b"] = w, b
        fs = raw[p + "fsmn.w"].clone(); fs[:, centre] += 1.0
        out[p + "fsmn.w"] = fs
        out[p + "fsmn.b"] = raw[p + "out.b"]
    for n in ("after_norm.g", "after_norm.b", "tp_norm.g", "tp_norm.b", "ctc.w", "ctc.b"):
        out[n] = raw[n]
    return {k: np.ascontiguousarray(v.detach().float().numpy()) for k, v in out.items()}


def build_supported_languages() -> Dict[str, dict]:
    """The `supported_languages` metadata catalog of the exporter (:299-310)."""
    return {code: {"name": name, "aliases": aliases, "selector_index": i, "prompt_token_ids": [tok]}
            for i, (code, name, aliases, tok) in enumerate(LANGUAGE_PROFILES)}


class SenseVoiceEngine:
    """One engine per GPU; `run` = the single InferenceSession.run of the reference script (:303)."""

    def __init__(self, dims: SenseVoiceDims, tensors: Dict[str, np.ndarray], *, precision: str = "f32", max_batch: int = 1,
                 max_samples: int = 480000, device: int = 0, use_tensor_cores: bool = True):
        self.lib = _cabi.load()
        self.dims = dims
        self.max_batch = max_batch
        self.max_samples = max_samples
        cfg = _cabi.NarConfig(kind=0, n_mels=dims.n_mels, nfft=dims.nfft, win=dims.win, hop=dims.hop, lfr_m=dims.lfr_m,
                              lfr_n=dims.lfr_n, d_model=dims.d_model, n_heads=dims.n_heads, ffn=dims.ffn,
                              n_blocks0=dims.n_blocks0, n_blocks=dims.n_blocks, n_tp_blocks=dims.n_tp_blocks, vocab=dims.vocab,
                              blank_id=dims.blank_id, n_prompt=1 + len(SYSTEM_PROMPT_IDS), n_lang=len(LANGUAGE_PROFILES),
                              fsmn_kernel=dims.fsmn_kernel, max_batch=max_batch, max_samples=max_samples,
                              precision={"f32": _cabi.PRECISION_F32, "bf16": _cabi.PRECISION_BF16}[precision], device=device,
                              use_tensor_cores=1 if use_tensor_cores else 0, ln_eps=dims.ln_eps)
        h = C.c_void_p()
        rc = self.lib.b200asr_nar_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise B200AsrError(f"b200asr_nar_create failed ({rc}): {self.lib.b200asr_nar_last_error(None).decode()}")
        self.h = h
        self.batch = 0
        self.n_samples = 0
        for name, arr in tensors.items():
            a = np.ascontiguousarray(arr, dtype=np.float32)
            self._ck(self.lib.b200asr_nar_set_tensor(self.h, name.encode(), a.ctypes.data_as(_cabi._F32P), a.size))
        self._ck(self.lib.b200asr_nar_finalize_weights(self.h))

    def _ck(self, rc: int):
        if rc != 0:
            raise B200AsrError(f"b200asr error {rc}: {self.lib.b200asr_nar_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200asr_nar_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def kernel_launches(self) -> int:
        return int(self.lib.b200asr_nar_kernel_launches(self.h))

    def set_option(self, key: str, value: int):
        self._ck(self.lib.b200asr_nar_set_option(self.h, key.encode(), int(value)))

    @property
    def stream_ptr(self) -> int:
        return int(self.lib.b200asr_nar_stream(self.h) or 0)

    def run(self, pcm: np.ndarray, language_idx=0, out_tokens: Optional[np.ndarray] = None,
            out_lens: Optional[np.ndarray] = None, clip_lens=None) -> List[List[int]]:
        """pcm [B][N] (or [N], or the reference's [B,1,N]): int16, or float32 carrying int16-range values.
        clip_lens: samples per clip of a ragged batch (pcm rows zero-padded to the longest clip)."""
        pcm = np.ascontiguousarray(pcm)
        if pcm.ndim == 1:
            pcm = pcm[None]
        if pcm.ndim == 3:
            pcm = pcm.reshape(pcm.shape[0], pcm.shape[-1])
        if pcm.dtype == np.int16:
            code = _cabi.PCM_I16
        elif pcm.dtype == np.float32:
            code = _cabi.PCM_F32
        else:
            raise TypeError(f"PCM dtype must be int16 or float32, got {pcm.dtype}")
        B, N = pcm.shape
        lang = np.ascontiguousarray(np.broadcast_to(np.asarray(language_idx, dtype=np.int32).reshape(-1), (B,)) if np.ndim(language_idx) == 0
                                    else np.asarray(language_idx, dtype=np.int32).reshape(B))
        ld = self.dims.lfr_frames(self.max_samples) + 1 + len(SYSTEM_PROMPT_IDS)
        toks = out_tokens if out_tokens is not None else np.zeros((B, ld), np.int32)
        lens = out_lens if out_lens is not None else np.zeros(B, np.int32)
        if clip_lens is None:
            self._ck(self.lib.b200asr_nar_run(self.h, pcm.ctypes.data_as(C.c_void_p), code, B, N, lang.ctypes.data_as(_cabi._I32P),
                                              toks.ctypes.data_as(_cabi._I32P), toks.shape[1], lens.ctypes.data_as(_cabi._I32P)))
        else:
            cl = np.ascontiguousarray(np.asarray(clip_lens, np.int32).reshape(B))
            self._ck(self.lib.b200asr_nar_run_ragged(self.h, pcm.ctypes.data_as(C.c_void_p), code, B, N, cl.ctypes.data_as(_cabi._I32P),
                                                     lang.ctypes.data_as(_cabi._I32P), toks.ctypes.data_as(_cabi._I32P), toks.shape[1],
                                                     lens.ctypes.data_as(_cabi._I32P)))
        self.batch, self.n_samples = B, N
        return [toks[b, :lens[b]].tolist() for b in range(B)]

    def upload(self, pcm: np.ndarray, language_idx=0):
        pcm = np.ascontiguousarray(pcm)
        if pcm.ndim == 1:
            pcm = pcm[None]
        code = _cabi.PCM_I16 if pcm.dtype == np.int16 else _cabi.PCM_F32
        B, N = pcm.shape
        lang = np.ascontiguousarray(np.broadcast_to(np.asarray(language_idx, dtype=np.int32).reshape(-1), (B,)))
        self._ck(self.lib.b200asr_nar_upload(self.h, pcm.ctypes.data_as(C.c_void_p), code, B, N, lang.ctypes.data_as(_cabi._I32P)))
        self.batch, self.n_samples = B, N

    def run_resident(self) -> List[List[int]]:
        ld = self.dims.lfr_frames(self.max_samples) + 1 + len(SYSTEM_PROMPT_IDS)
        toks = np.zeros((self.batch, ld), np.int32)
        lens = np.zeros(self.batch, np.int32)
        self._ck(self.lib.b200asr_nar_run_resident(self.h, toks.ctypes.data_as(_cabi._I32P), ld, lens.ctypes.data_as(_cabi._I32P)))
        return [toks[b, :lens[b]].tolist() for b in range(self.batch)]

    def get_stage(self, name: str, capacity: int) -> np.ndarray:
        out = np.empty(capacity, np.float32)
        n = C.c_int64(0)
        self._ck(self.lib.b200asr_nar_get_stage(self.h, name.encode(), out.ctypes.data_as(_cabi._F32P), capacity, C.byref(n)))
        return out[:n.value]


def transcribe_clip(engine: SenseVoiceEngine, raw_audio_int16: np.ndarray, language: str = "auto", *, sample_rate: int = 16000):
    """Host loop of Inference_SenseVoice_ONNX.py:262-310 for one clip: language selector -> one run -> token ids, RTF.
    (SenseVoice takes int16-range samples: audio_pcm_scale = 1, :401.)"""
    import time
    from .ort_io import resolve_supported_language
    code, entry = resolve_supported_language(build_supported_languages(), language)
    pcm = np.ascontiguousarray(np.asarray(raw_audio_int16, dtype=np.int16).reshape(1, -1))
    t0 = time.time()
    tokens = engine.run(pcm, entry["selector_index"])[0]
    elapsed = time.time() - t0
    audio_s = pcm.shape[1] / sample_rate
    return dict(tokens=tokens, language=code, elapsed_s=elapsed, rtf=elapsed / audio_s)


def transcribe_long(engine, raw_audio_int16: np.ndarray, language: str = "auto", *, input_audio_length: Optional[int] = None,
                    sliding_window: int = 0, sample_rate: int = 16000):
    """The scripts' window loop (Inference_SenseVoice_ONNX.py:236-260,290-307; Paraformer :233-257,284-300): a graph exported
    with a static audio length sees the clip as windows of `input_audio_length` samples, stride `sliding_window` (<= 0: the
    window length), the tail zero-padded; the per-window token ids are concatenated.  `input_audio_length=None` = the dynamic
    axis (one window = the whole clip).  The windows are independent, so they go through the engine as batches of up to
    `engine.max_batch` clips instead of one run per window.  Works for SenseVoiceEngine and ParaformerEngine."""
    import time
    from .ort_io import resolve_supported_language
    from .whisper_infer import plan_windows
    sel = 0
    code = None
    if not isinstance(engine, _paraformer_engine_type()):
        code, entry = resolve_supported_language(build_supported_languages(), language)
        sel = entry["selector_index"]
    pcm = np.asarray(raw_audio_int16, dtype=np.int16).reshape(-1)
    audio_len = pcm.shape[0]
    win = audio_len if input_audio_length is None else int(input_audio_length)
    if win > engine.max_samples:
        raise ValueError(f"window of {win} samples exceeds the engine's max_samples {engine.max_samples}")
    n_win, stride, aligned = plan_windows(audio_len, win, sliding_window)
    padded = np.zeros(aligned, np.int16)
    padded[:audio_len] = pcm
    clips = np.stack([padded[i * stride:i * stride + win] for i in range(n_win)])
    t0 = time.time()
    tokens: List[int] = []
    per_window: List[List[int]] = []
    for i in range(0, n_win, engine.max_batch):
        for toks in engine.run(clips[i:i + engine.max_batch], sel):
            per_window.append(toks)
            tokens.extend(toks)
    elapsed = time.time() - t0
    return dict(tokens=tokens, per_window=per_window, windows=n_win, language=code, elapsed_s=elapsed,
                rtf=elapsed / (audio_len / sample_rate))


def _paraformer_engine_type():
    from .paraformer import ParaformerEngine
    return ParaformerEngine


"""Paraformer (non-streaming) on the B200 engine: weight folds, the ctypes face of the non-autoregressive C ABI and the
host loop of the reference driver (/root/reference/Paraformer/Non-Streaming/Inference_Paraformer_ONNX.py).

Folds follow PARAFORMER.__init__ (/root/reference/Paraformer/Non-Streaming/Export_Paraformer.py:385-465): every in-block
LayerNorm affine is absorbed into the Linear that consumes it (float64, rounded once, :245-272), d_k^-0.25 goes onto the
q and k rows, the FSMN identity onto the centre tap (:305-312), CMVN mean x scale + sinusoid position into one additive
table (:461-465).  Nothing here imports oracle/.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, asdict
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _cabi
from .engine import B200AsrError
from .sensevoice import SenseVoiceEngine


@dataclass(frozen=True)
class ParaformerDims:
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
    dec_att_blocks: int = 16
    dec_ffn_blocks: int = 1
    dec_ffn: int = 2048
    vocab: int = 8404
    fsmn_kernel: int = 11
    cif_kernel: int = 3
    tail_threshold: float = 0.45
    ln_eps: float = 1e-12
    dec_ln_eps: float = 1e-12
    pre_emphasis: float = 0.97
    sample_rate: int = 16000

    @property
    def feat(self) -> int:
        return self.n_mels * self.lfr_m

    @property
    def head_dim(self) -> int:
        return self.d_model // self.n_heads

    @property
    def enc_blocks(self) -> int:
        return self.n_blocks0 + self.n_blocks

    def frames(self, n_samples: int) -> int:
        return (n_samples - self.win) // self.hop + 1

    def lfr_frames(self, n_samples: int) -> int:
        return (self.frames(n_samples) + self.lfr_n - 1) // self.lfr_n

    def to_dict(self):
        return asdict(self)


PARAFORMER_LARGE = ParaformerDims()
PARAFORMER_TINY_TEST = ParaformerDims(d_model=128, n_heads=2, ffn=256, n_blocks0=1, n_blocks=2, dec_att_blocks=2,
                                      dec_ffn_blocks=1, dec_ffn=256, vocab=300)
PRESETS = {"paraformer-large": PARAFORMER_LARGE, "paraformer-tiny-test": PARAFORMER_TINY_TEST}


def synth_paraformer_checkpoint(d: ParaformerDims, seed: int) -> Dict[str, torch.Tensor]:
    """Seeded random checkpoint; same draw order as the test oracle's generator."""
    g = torch.Generator().manual_seed(seed)
    raw: Dict[str, torch.Tensor] = {}

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    def norm(name, n):
        raw[name + ".g"] = 1.0 + rn(n, std=0.1); raw[name + ".b"] = rn(n, std=0.1)

    D = d.d_model
    raw["cmvn_means"] = rn(d.feat, std=1.0) - 8.0
    raw["cmvn_vars"] = 0.1 + 0.05 * torch.rand(d.feat, generator=g)
    for i in range(d.enc_blocks):
        din = d.feat if i == 0 else D
        p = f"enc{i}."
        norm(p + "norm1", din)
        raw[p + "qkv.w"] = rn(3 * D, din, std=din ** -0.5); raw[p + "qkv.b"] = rn(3 * D, std=0.1)
        raw[p + "fsmn.w"] = rn(D, d.fsmn_kernel, std=0.2)
        raw[p + "out.w"] = rn(D, D, std=D ** -0.5); raw[p + "out.b"] = rn(D, std=0.1)
        norm(p + "norm2", D)
        raw[p + "w1.w"] = rn(d.ffn, D, std=D ** -0.5); raw[p + "w1.b"] = rn(d.ffn, std=0.1)
        raw[p + "w2.w"] = rn(D, d.ffn, std=d.ffn ** -0.5); raw[p + "w2.b"] = rn(D, std=0.1)
    norm("enc_after_norm", D)
    raw["cif.conv.w"] = rn(D, D, d.cif_kernel, std=(D * d.cif_kernel) ** -0.5); raw["cif.conv.b"] = rn(D, std=0.1)
    raw["cif.out.w"] = rn(1, D, std=D ** -0.5 * 2.0); raw["cif.out.b"] = rn(1, std=0.1) - 0.3
    for i in range(d.dec_att_blocks + d.dec_ffn_blocks):
        p = f"dec{i}."
        norm(p + "norm1", D)
        raw[p + "w1.w"] = rn(d.dec_ffn, D, std=D ** -0.5); raw[p + "w1.b"] = rn(d.dec_ffn, std=0.1)
        norm(p + "ffn_norm", d.dec_ffn)
        raw[p + "w2.w"] = rn(D, d.dec_ffn, std=d.dec_ffn ** -0.5)
        if i < d.dec_att_blocks:
            norm(p + "norm2", D)
            raw[p + "fsmn.w"] = rn(D, d.fsmn_kernel, std=0.2)
            norm(p + "norm3", D)
            raw[p + "q.w"] = rn(D, D, std=D ** -0.5); raw[p + "q.b"] = rn(D, std=0.1)
            raw[p + "kv.w"] = rn(2 * D, D, std=D ** -0.5); raw[p + "kv.b"] = rn(2 * D, std=0.1)
            raw[p + "cout.w"] = rn(D, D, std=D ** -0.5); raw[p + "cout.b"] = rn(D, std=0.1)
    norm("dec_after_norm", D)
    raw["out.w"] = rn(d.vocab, D, std=D ** -0.5 * 3.0); raw["out.b"] = rn(d.vocab, std=0.5)
    return raw


def _kaldi_stft_kernel(d: ParaformerDims) -> torch.Tensor:
    """Hamming-windowed one-sided DFT basis times (pre-emphasis @ DC-removal), [2F][win] (:326-343)."""
    win = d.win
    window = torch.hamming_window(win, periodic=False, alpha=0.54, beta=0.46)
    omega = (2.0 * torch.pi / d.nfft) * torch.arange(d.nfft // 2 + 1, dtype=torch.float32).unsqueeze(1) * torch.arange(win, dtype=torch.float32).unsqueeze(0)
    dc = torch.eye(win) - torch.full((win, win), 1.0 / win)
    prev = torch.zeros((win, win)); prev[0, 0] = 1.0; prev[1:, :-1] = torch.eye(win - 1)
    ft = torch.matmul(torch.eye(win) - float(d.pre_emphasis) * prev, dc)
    return torch.cat([torch.matmul(torch.cos(omega) * window.unsqueeze(0), ft),
                      torch.matmul(-torch.sin(omega) * window.unsqueeze(0), ft)], dim=0).contiguous()


def _absorb(g, b, w, bias, scale=1.0):
    """Linear(LN_affine(x)) == Linear'(LN_plain(x)): scale the outputs, bias += W @ beta, W *= gamma; float64, one rounding."""
    W = w.to(torch.float64)
    B = bias.to(torch.float64) if bias is not None else torch.zeros(w.shape[0], dtype=torch.float64)
    s = torch.as_tensor(scale, dtype=torch.float64)
    if s.ndim == 0:
        W = W * s; B = B * s
    else:
        W = W * s.reshape(-1).unsqueeze(1); B = B * s.reshape(-1)
    B = B + torch.matmul(W, b.to(torch.float64))
    W = W * g.to(torch.float64).unsqueeze(0)
    return W.to(torch.float32), B.to(torch.float32)


def fold_paraformer(raw: Dict[str, torch.Tensor], d: ParaformerDims, max_samples: int) -> Dict[str, np.ndarray]:
    import torchaudio.compliance.kaldi as kaldi
    out: Dict[str, torch.Tensor] = {}
    D = d.d_model
    scale = float(D) ** 0.5
    max_lfr = d.lfr_frames(max_samples)
    out["fbank_kernel"] = _kaldi_stft_kernel(d)
    banks, _ = kaldi.get_mel_banks(d.n_mels, d.nfft, d.sample_rate, 20.0, 0.0, 100.0, -500.0, 1.0)
    out["mel_filters"] = torch.nn.functional.pad(banks, (0, 1), mode="constant", value=0.0).to(torch.float32).t().contiguous()
    cv = raw["cmvn_vars"] * scale
    out["cmvn_vars"] = cv
    inc = torch.log(torch.tensor([10000], dtype=torch.float32)) / (d.feat / 2 - 1)
    inv = torch.exp(torch.arange(d.feat / 2).type(torch.float32) * (-inc)).reshape(1, -1)
    st = torch.arange(1, max_lfr + 1, dtype=torch.int32).type(torch.float32).reshape(-1, 1) * inv
    pos = torch.cat([torch.sin(st), torch.cos(st)], dim=1)
    out["encoder_input_bias"] = (raw["cmvn_means"].to(torch.float64) * cv.to(torch.float64) + pos.to(torch.float64)).to(torch.float32)
    f = float(d.head_dim ** (-0.25))
    c = d.fsmn_kernel // 2
    for i in range(d.enc_blocks):
        p = f"enc{i}."
        qk = torch.ones(3 * D, dtype=torch.float64); qk[:-D] = f
        out[p + "qkv.w"], out[p + "qkv.b"] = _absorb(raw[p + "norm1.g"], raw[p + "norm1.b"], raw[p + "qkv.w"], raw[p + "qkv.b"], qk)
        out[p + "w1.w"], out[p + "w1.b"] = _absorb(raw[p + "norm2.g"], raw[p + "norm2.b"], raw[p + "w1.w"], raw[p + "w1.b"])
        fs = raw[p + "fsmn.w"].to(torch.float64); fs[:, c] += 1.0
        out[p + "fsmn.w"] = fs.to(torch.float32)
        for k in ("out.w", "out.b", "w2.w", "w2.b"):
            out[p + k] = raw[p + k]
    out["enc_after_norm.g"], out["enc_after_norm.b"] = raw["enc_after_norm.g"], raw["enc_after_norm.b"]
    for k in ("cif.conv.w", "cif.conv.b", "cif.out.w", "cif.out.b"):
        out[k] = raw[k]
    for i in range(d.dec_att_blocks + d.dec_ffn_blocks):
        p = f"dec{i}."
        out[p + "w1.w"], out[p + "w1.b"] = _absorb(raw[p + "norm1.g"], raw[p + "norm1.b"], raw[p + "w1.w"], raw[p + "w1.b"])
        out[p + "w2.w"], out[p + "w2.b"] = _absorb(raw[p + "ffn_norm.g"], raw[p + "ffn_norm.b"], raw[p + "w2.w"], None)
        if i < d.dec_att_blocks:
            out[p + "norm2.g"], out[p + "norm2.b"] = raw[p + "norm2.g"], raw[p + "norm2.b"]
            fs = raw[p + "fsmn.w"].to(torch.float64); fs[:, c] += 1.0
            out[p + "fsmn.w"] = fs.to(torch.float32)
            out[p + "q.w"], out[p + "q.b"] = _absorb(raw[p + "norm3.g"], raw[p + "norm3.b"], raw[p + "q.w"], raw[p + "q.b"], f)
            kv = torch.ones(2 * D, dtype=torch.float64); kv[:D] = f
            out[p + "kv.w"] = (raw[p + "kv.w"].to(torch.float64) * kv.unsqueeze(1)).to(torch.float32)
            out[p + "kv.b"] = (raw[p + "kv.b"].to(torch.float64) * kv).to(torch.float32)
            out[p + "cout.w"], out[p + "cout.b"] = raw[p + "cout.w"], raw[p + "cout.b"]
    out["out.w"], out["out.b"] = _absorb(raw["dec_after_norm.g"], raw["dec_after_norm.b"], raw["out.w"], raw["out.b"])
    return {k: np.ascontiguousarray(v.detach().float().numpy()) for k, v in out.items()}


class ParaformerEngine(SenseVoiceEngine):
    """`run(pcm)` = the single InferenceSession.run of Inference_Paraformer_ONNX.py:293 (outputs token_ids, num_id)."""

    def __init__(self, dims: ParaformerDims, tensors: Dict[str, np.ndarray], *, precision: str = "f32", max_batch: int = 1,
                 max_samples: int = 480000, device: int = 0, use_tensor_cores: bool = True):
        self.lib = _cabi.load()
        self.dims = dims
        self.max_batch = max_batch
        self.max_samples = max_samples
        cfg = _cabi.NarConfig(kind=1, n_mels=dims.n_mels, nfft=dims.nfft, win=dims.win, hop=dims.hop, lfr_m=dims.lfr_m,
                              lfr_n=dims.lfr_n, d_model=dims.d_model, n_heads=dims.n_heads, ffn=dims.ffn,
                              n_blocks0=dims.n_blocks0, n_blocks=dims.n_blocks, n_tp_blocks=0, vocab=dims.vocab, blank_id=0,
                              n_prompt=0, n_lang=0, fsmn_kernel=dims.fsmn_kernel, max_batch=max_batch, max_samples=max_samples,
                              precision={"f32": _cabi.PRECISION_F32, "bf16": _cabi.PRECISION_BF16}[precision], device=device,
                              use_tensor_cores=1 if use_tensor_cores else 0, ln_eps=dims.ln_eps,
                              dec_att_blocks=dims.dec_att_blocks, dec_ffn_blocks=dims.dec_ffn_blocks, dec_ffn=dims.dec_ffn,
                              cif_kernel=dims.cif_kernel, tail_threshold=dims.tail_threshold, dec_ln_eps=dims.dec_ln_eps)
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

    def _ld(self) -> int:
        return self.dims.lfr_frames(self.max_samples) + 1

    def run(self, pcm: np.ndarray, language_idx=0, out_tokens: Optional[np.ndarray] = None,
            out_lens: Optional[np.ndarray] = None, clip_lens=None) -> List[List[int]]:
        """clip_lens: samples per clip of a ragged batch (pcm rows zero-padded to the longest clip)."""
        pcm = np.ascontiguousarray(pcm)
        if pcm.ndim == 1:
            pcm = pcm[None]
        if pcm.ndim == 3:
            pcm = pcm.reshape(pcm.shape[0], pcm.shape[-1])
        if pcm.dtype not in (np.int16, np.float32):
            raise TypeError(f"PCM dtype must be int16 or float32, got {pcm.dtype}")
        code = _cabi.PCM_I16 if pcm.dtype == np.int16 else _cabi.PCM_F32
        B, N = pcm.shape
        toks = out_tokens if out_tokens is not None else np.zeros((B, self._ld()), np.int32)
        lens = out_lens if out_lens is not None else np.zeros(B, np.int32)
        if clip_lens is None:
            self._ck(self.lib.b200asr_nar_run(self.h, pcm.ctypes.data_as(C.c_void_p), code, B, N, None,
                                              toks.ctypes.data_as(_cabi._I32P), toks.shape[1], lens.ctypes.data_as(_cabi._I32P)))
        else:
            cl = np.ascontiguousarray(np.asarray(clip_lens, np.int32).reshape(B))
            self._ck(self.lib.b200asr_nar_run_ragged(self.h, pcm.ctypes.data_as(C.c_void_p), code, B, N, cl.ctypes.data_as(_cabi._I32P), None,
                                                     toks.ctypes.data_as(_cabi._I32P), toks.shape[1], lens.ctypes.data_as(_cabi._I32P)))
        self.batch, self.n_samples = B, N
        return [toks[b, :lens[b]].tolist() for b in range(B)]


def tokens_to_text(tokens: List[int], vocab: List[str], decode_mode: str = "zh") -> str:
    """Vocab lookup of the driver (:86-89, :296): zh joins characters, en joins BPE pieces on the '@@ ' continuation marker."""
    pieces = [vocab[t] for t in tokens]
    if decode_mode == "en":
        return " ".join(pieces).replace("@@ ", "")
    return "".join(pieces)

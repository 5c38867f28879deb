"""Qwen3-ASR on the B200 engine: weight folds, the ctypes face of the `b200asr_qwen_*` C ABI and the host loop of the
reference driver (/root/reference/Qwen_ASR/Inference_Qwen_ASR_ONNX.py).

Folds follow the exporter (/root/reference/Qwen_ASR/Export_Qwen_ASR.py): encoder `_fuse_encoder_weights` :829-848
(fused QKV, LayerNorm affines absorbed into the next Linear, sqrt(scaling) on the q and k rows, ln_post into proj1),
decoder `_fuse_weights` :1141-1190 (fused QKV with the input RMS-norm weight, head_dim^-0.25 into the QK-norm weights,
fused gate_up with the post-attention RMS-norm weight, final norm kept), rotary table :977-984, sinusoid positions
:399-405.  The exporter's optional quantisation reorders (:1192-1257) are exact permutations absorbed into the weights
and change no result, so they are not applied.  Nothing here imports oracle/.
"""
from __future__ import annotations

import ctypes as C
import time
from dataclasses import dataclass, asdict
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _cabi
from .engine import B200AsrError
from .weights import hann_dft_kernel, slaney_mel_filterbank

CHUNK = 100                  # mel frames per conv chunk (2 * n_window, :744)
CHUNK_TOKENS = 13            # tokens per full chunk (:519-527)


@dataclass(frozen=True)
class QwenDims:
    n_mels: int = 128
    nfft: int = 400
    hop: int = 160
    sample_rate: int = 16000
    enc_layers: int = 18
    enc_d: int = 896
    enc_heads: int = 14
    enc_ffn: int = 3584
    conv_ch: int = 480
    out_dim: int = 1024
    chunks_per_window: int = 8
    max_source_positions: int = 1500
    enc_ln_eps: float = 1e-5
    vocab: int = 151936
    hidden: int = 1024
    inter: int = 3072
    dec_layers: int = 28
    heads: int = 16
    kv_heads: int = 8
    head_dim: int = 128
    rope_theta: float = 1000000.0
    rms_eps: float = 1e-6
    max_seq_len: int = 1024

    @property
    def enc_head_dim(self) -> int:
        return self.enc_d // self.enc_heads

    @property
    def conv_freq(self) -> int:
        return (((self.n_mels + 1) // 2 + 1) // 2 + 1) // 2

    def audio_tokens(self, n_samples: int) -> int:
        """Tokens the audio tower yields for a clip (`_get_feat_extract_output_lengths`, :519-527)."""
        frames = n_samples // self.hop
        full, rem = divmod(frames, CHUNK)
        n = 0
        if rem > 0:
            n = ((((rem - 1) // 2 + 1) - 1) // 2 + 1 - 1) // 2 + 1
        return full * CHUNK_TOKENS + n

    def to_dict(self):
        return asdict(self)


QWEN3_ASR_0_6B = QwenDims()
QWEN3_ASR_1_7B = QwenDims(enc_layers=24, enc_d=1024, enc_heads=16, enc_ffn=4096, out_dim=2048, hidden=2048, inter=6144)
QWEN_TINY_TEST = QwenDims(enc_layers=2, enc_d=128, enc_heads=2, enc_ffn=256, conv_ch=16, out_dim=128, vocab=512, hidden=128,
                          inter=256, dec_layers=2, heads=4, kv_heads=2, head_dim=64, max_seq_len=256)
PRESETS = {"qwen3-asr-0.6b": QWEN3_ASR_0_6B, "qwen3-asr-1.7b": QWEN3_ASR_1_7B, "qwen3-asr-tiny-test": QWEN_TINY_TEST}


@dataclass(frozen=True)
class QwenPrompt:
    """Token ids the exporter bakes around the audio (:1540-1586) and the stop set (:1503)."""
    head_ids: Sequence[int]      # <|im_start|> system \n
    suffix_ids: Sequence[int]    # <|im_end|> \n <|im_start|> user \n <|audio_start|>
    tail_ids: Sequence[int]      # <|audio_end|> <|im_end|> \n <|im_start|> assistant \n "language "
    stop_ids: Sequence[int]      # <|endoftext|>, <|im_end|>


# Qwen3-ASR tokenizer ids (Qwen2 BPE vocabulary + the ASR special tokens); override from the checkpoint's tokenizer
# when one is present.
QWEN3_PROMPT = QwenPrompt(head_ids=(151644, 8948, 198), suffix_ids=(151645, 198, 151644, 872, 198, 151669),
                          tail_ids=(151670, 151645, 198, 151644, 77091, 198, 11528, 220), stop_ids=(151643, 151645))
TINY_PROMPT = QwenPrompt(head_ids=(500, 501, 502), suffix_ids=(503, 502, 500, 504, 502, 505),
                         tail_ids=(506, 503, 502, 500, 507, 502, 508, 509), stop_ids=(510, 503))


def synth_qwen_checkpoint(d: QwenDims, seed: int) -> Dict[str, torch.Tensor]:
    """Seeded random checkpoint under the Hugging Face names of the exporter's skeleton model (:311-516); same draw
    order as the test oracle's generator so parity tests can build both sides from one seed."""
    g = torch.Generator().manual_seed(seed)
    raw: Dict[str, torch.Tensor] = {}

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    def lin(name, nout, nin, bias=True, std=None):
        raw[name + ".weight"] = rn(nout, nin, std=std if std is not None else nin ** -0.5)
        if bias:
            raw[name + ".bias"] = rn(nout, std=0.05)

    def ln(name, n):
        raw[name + ".weight"] = 1.0 + rn(n, std=0.1)
        raw[name + ".bias"] = rn(n, std=0.05)

    Cc = d.conv_ch
    a = "thinker.audio_tower."
    raw[a + "conv2d1.weight"] = rn(Cc, 1, 3, 3, std=1.0 / 3.0); raw[a + "conv2d1.bias"] = rn(Cc, std=0.05)
    raw[a + "conv2d2.weight"] = rn(Cc, Cc, 3, 3, std=(9 * Cc) ** -0.5); raw[a + "conv2d2.bias"] = rn(Cc, std=0.05)
    raw[a + "conv2d3.weight"] = rn(Cc, Cc, 3, 3, std=(9 * Cc) ** -0.5); raw[a + "conv2d3.bias"] = rn(Cc, std=0.05)
    lin(a + "conv_out", d.enc_d, Cc * d.conv_freq, bias=False)
    for i in range(d.enc_layers):
        p = f"{a}layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            lin(p + "self_attn." + n, d.enc_d, d.enc_d)
        ln(p + "self_attn_layer_norm", d.enc_d)
        lin(p + "fc1", d.enc_ffn, d.enc_d)
        lin(p + "fc2", d.enc_d, d.enc_ffn)
        ln(p + "final_layer_norm", d.enc_d)
    ln(a + "ln_post", d.enc_d)
    lin(a + "proj1", d.enc_d, d.enc_d)
    lin(a + "proj2", d.out_dim, d.enc_d)
    t = "thinker.model."
    raw[t + "embed_tokens.weight"] = rn(d.vocab, d.hidden, std=0.5)
    qd, kd = d.heads * d.head_dim, d.kv_heads * d.head_dim
    for i in range(d.dec_layers):
        p = f"{t}layers.{i}."
        lin(p + "self_attn.q_proj", qd, d.hidden, bias=False)
        lin(p + "self_attn.k_proj", kd, d.hidden, bias=False)
        lin(p + "self_attn.v_proj", kd, d.hidden, bias=False)
        lin(p + "self_attn.o_proj", d.hidden, qd, bias=False)
        raw[p + "self_attn.q_norm.weight"] = 1.0 + rn(d.head_dim, std=0.1)
        raw[p + "self_attn.k_norm.weight"] = 1.0 + rn(d.head_dim, std=0.1)
        lin(p + "mlp.gate_proj", d.inter, d.hidden, bias=False)
        lin(p + "mlp.up_proj", d.inter, d.hidden, bias=False)
        lin(p + "mlp.down_proj", d.hidden, d.inter, bias=False)
        raw[p + "input_layernorm.weight"] = 1.0 + rn(d.hidden, std=0.1)
        raw[p + "post_attention_layernorm.weight"] = 1.0 + rn(d.hidden, std=0.1)
    raw[t + "norm.weight"] = 1.0 + rn(d.hidden, std=0.1)
    raw["thinker.lm_head.weight"] = rn(d.vocab, d.hidden, std=d.hidden ** -0.5 * 2.0)
    return raw


def _sinusoids(length: int, channels: int) -> torch.Tensor:
    inc = np.log(10000.0) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2).float())
    st = torch.arange(length)[:, None] * inv[None, :]
    return torch.cat([torch.sin(st), torch.cos(st)], dim=1)


def fold_qwen(state: Dict[str, torch.Tensor], d: QwenDims, tie_lm_head: bool = False) -> Dict[str, np.ndarray]:
    """HF state dict (fp32 tensors) -> the tensors include/b200asr.h names.  `tie_lm_head` drops lm_head.w so the
    engine projects with the embedding table (tied checkpoints store the table once, :1179-1190)."""
    st = {k: (v.detach().float() if isinstance(v, torch.Tensor) else torch.from_numpy(np.array(v, dtype=np.float32))) for k, v in state.items()}
    out: Dict[str, torch.Tensor] = {}
    out["stft_kernel"] = torch.from_numpy(hann_dft_kernel(d.nfft, 1.0))
    out["mel_fbank"] = torch.from_numpy(slaney_mel_filterbank(d.nfft // 2 + 1, d.n_mels, d.sample_rate))
    a = "thinker.audio_tower."
    out["conv1.w"] = st[a + "conv2d1.weight"].reshape(d.conv_ch, 9)
    out["conv1.b"] = st[a + "conv2d1.bias"]
    for i in (2, 3):
        out[f"conv{i}.w"] = st[f"{a}conv2d{i}.weight"]
        out[f"conv{i}.b"] = st[f"{a}conv2d{i}.bias"]
    out["conv_out.w"] = st[a + "conv_out.weight"]
    out["enc_pos"] = _sinusoids(d.max_source_positions, d.enc_d)[:CHUNK_TOKENS]
    s = float(d.enc_head_dim) ** -0.25
    for i in range(d.enc_layers):
        p = f"{a}layers.{i}."
        W = torch.cat([st[p + f"self_attn.{n}_proj.weight"] for n in "qkv"], dim=0)
        b = torch.cat([st[p + f"self_attn.{n}_proj.bias"] for n in "qkv"], dim=0)
        g1, b1 = st[p + "self_attn_layer_norm.weight"], st[p + "self_attn_layer_norm.bias"]
        b = b + W @ b1
        W = W * g1.unsqueeze(0)
        W[: 2 * d.enc_d] *= s
        b[: 2 * d.enc_d] *= s
        out[f"enc{i}.qkv.w"], out[f"enc{i}.qkv.b"] = W, b
        out[f"enc{i}.out.w"], out[f"enc{i}.out.b"] = st[p + "self_attn.out_proj.weight"], st[p + "self_attn.out_proj.bias"]
        g2, b2 = st[p + "final_layer_norm.weight"], st[p + "final_layer_norm.bias"]
        out[f"enc{i}.fc1.w"] = st[p + "fc1.weight"] * g2.unsqueeze(0)
        out[f"enc{i}.fc1.b"] = st[p + "fc1.bias"] + st[p + "fc1.weight"] @ b2
        out[f"enc{i}.fc2.w"], out[f"enc{i}.fc2.b"] = st[p + "fc2.weight"], st[p + "fc2.bias"]
    gp, bp = st[a + "ln_post.weight"], st[a + "ln_post.bias"]
    out["proj1.w"] = st[a + "proj1.weight"] * gp.unsqueeze(0)
    out["proj1.b"] = st[a + "proj1.bias"] + st[a + "proj1.weight"] @ bp
    out["proj2.w"], out["proj2.b"] = st[a + "proj2.weight"], st[a + "proj2.bias"]
    t = "thinker.model."
    out["embed.w"] = st[t + "embed_tokens.weight"]
    if not tie_lm_head and "thinker.lm_head.weight" in st:
        out["lm_head.w"] = st["thinker.lm_head.weight"]
    out["final_norm.g"] = st[t + "norm.weight"]
    qs = float(d.head_dim) ** -0.25
    for i in range(d.dec_layers):
        p = f"{t}layers.{i}."
        W = torch.cat([st[p + f"self_attn.{n}_proj.weight"] for n in "qkv"], dim=0)
        out[f"dec{i}.qkv.w"] = W * st[p + "input_layernorm.weight"].unsqueeze(0)
        out[f"dec{i}.qk_norm.g"] = torch.stack([st[p + "self_attn.q_norm.weight"] * qs, st[p + "self_attn.k_norm.weight"] * qs])
        out[f"dec{i}.o.w"] = st[p + "self_attn.o_proj.weight"]
        g2 = st[p + "post_attention_layernorm.weight"].unsqueeze(0)
        out[f"dec{i}.gate_up.w"] = torch.cat([st[p + "mlp.gate_proj.weight"] * g2, st[p + "mlp.up_proj.weight"] * g2], dim=0)
        out[f"dec{i}.down.w"] = st[p + "mlp.down_proj.weight"]
    inv_freq = 1.0 / (d.rope_theta ** (torch.arange(0, d.head_dim, 2, dtype=torch.int64).float() / d.head_dim))
    theta = torch.arange(d.max_seq_len, dtype=torch.float32).unsqueeze(-1) * inv_freq
    out["rope_cos"], out["rope_sin"] = torch.cos(theta), torch.sin(theta)
    return {k: np.ascontiguousarray(v.numpy(), dtype=np.float32) for k, v in out.items()}


def _i32(ids) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(list(ids), dtype=np.int32).reshape(-1))


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_cabi._I32P) if a.size else None


class QwenEngine:
    """One engine per GPU.  `encode` + `prefill` = the merged prefill session of the reference script (:656); `decode_step`
    = embed + merged decode session (:703-717); `transcribe` runs the whole greedy loop on the device."""

    def __init__(self, dims: QwenDims, tensors: Dict[str, np.ndarray], prompt: QwenPrompt, *, precision: str = "f32",
                 max_batch: int = 1, max_samples: int = 480000, device: int = 0, use_tensor_cores: bool = True):
        self.lib = _cabi.load()
        self.dims = dims
        self.prompt = prompt
        self.max_batch = max_batch
        self.max_samples = max_samples
        cfg = _cabi.QwenConfig(device=device, max_batch=max_batch, max_samples=max_samples,
                               precision={"f32": _cabi.PRECISION_F32, "bf16": _cabi.PRECISION_BF16}[precision],
                               use_tensor_cores=1 if use_tensor_cores else 0, n_mels=dims.n_mels, n_fft=dims.nfft, hop=dims.hop,
                               enc_layers=dims.enc_layers, enc_d=dims.enc_d, enc_heads=dims.enc_heads, enc_ffn=dims.enc_ffn,
                               conv_ch=dims.conv_ch, out_dim=dims.out_dim, chunks_per_window=dims.chunks_per_window,
                               enc_ln_eps=dims.enc_ln_eps, vocab=dims.vocab, hidden=dims.hidden, inter=dims.inter,
                               dec_layers=dims.dec_layers, heads=dims.heads, kv_heads=dims.kv_heads, head_dim=dims.head_dim,
                               max_seq_len=dims.max_seq_len, rms_eps=dims.rms_eps)
        h = C.c_void_p()
        rc = self.lib.b200asr_qwen_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise B200AsrError(f"b200asr_qwen_create failed ({rc}): {self.lib.b200asr_qwen_create_error().decode()}")
        self.h = h
        self.batch = 0
        self.n_prompt = 0
        self.repeat_penalty, self.penalty_range = 1.0, 10
        for name, arr in tensors.items():
            a = np.ascontiguousarray(arr, dtype=np.float32)
            self._ck(self.lib.b200asr_qwen_set_tensor(self.h, name.encode(), a.ctypes.data_as(_cabi._F32P), a.size))
        self._ck(self.lib.b200asr_qwen_finalize_weights(self.h))
        hd, sf, tl, sp = _i32(prompt.head_ids), _i32(prompt.suffix_ids), _i32(prompt.tail_ids), _i32(prompt.stop_ids)
        self._ck(self.lib.b200asr_qwen_set_prompt(self.h, _ptr(hd), hd.size, _ptr(sf), sf.size, _ptr(tl), tl.size, _ptr(sp), sp.size))

    def _ck(self, rc: int):
        if rc != 0:
            raise B200AsrError(f"b200asr error {rc}: {self.lib.b200asr_qwen_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200asr_qwen_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def kernel_launches(self) -> int:
        return int(self.lib.b200asr_qwen_kernel_launches(self.h))

    def set_option(self, key: str, value: int):
        self._ck(self.lib.b200asr_qwen_set_option(self.h, key.encode(), int(value)))

    def set_decode_options(self, repeat_penalty: float = 1.0, penalty_range: int = 10):
        """repeat_penalty 1.0 = greedy; anything else = the script's penalty-greedy strategy (:90-91,369-376)."""
        self._ck(self.lib.b200asr_qwen_set_decode_options(self.h, float(repeat_penalty), int(penalty_range)))
        self.repeat_penalty, self.penalty_range = float(repeat_penalty), int(penalty_range)

    def set_sampling(self, temperature: float = 0.0, top_k: int = 10, top_p: float = 0.95, repetition_penalty: float = 1.0,
                     seed: int = 0, noise: Optional[np.ndarray] = None):
        """temperature > 0 = the script's sampling strategy (USE_SAMPLING, :85-89); noise [launch][max_batch][top_k] in (0,1)
        replaces the head's random draw for reproducible runs."""
        nz, rows = None, 0
        if noise is not None:
            nz = np.ascontiguousarray(noise, dtype=np.float32)
            if nz.ndim != 3 or nz.shape[1] != self.max_batch or nz.shape[2] != top_k:
                raise ValueError("noise must be [launch][max_batch][top_k]")
            rows = nz.shape[0]
        self._ck(self.lib.b200asr_qwen_set_sampling(self.h, float(temperature), int(top_k), float(top_p), float(repetition_penalty),
                                                    int(seed), nz.ctypes.data_as(_cabi._F32P) if nz is not None else None, rows))

    @property
    def stream_ptr(self) -> int:
        return int(self.lib.b200asr_qwen_stream(self.h) or 0)

    @staticmethod
    def _pcm(pcm: np.ndarray):
        pcm = np.ascontiguousarray(pcm)
        if pcm.ndim == 1:
            pcm = pcm[None]
        if pcm.ndim == 3:
            pcm = pcm.reshape(pcm.shape[0], pcm.shape[-1])
        if pcm.dtype == np.int16:
            return pcm, _cabi.PCM_I16
        if pcm.dtype == np.float32:
            return pcm, _cabi.PCM_F32
        raise TypeError(f"PCM dtype must be int16 or float32, got {pcm.dtype}")

    @staticmethod
    def pad_ragged(clips: Sequence[np.ndarray]):
        """Clips of different lengths -> ([B][longest] zero-padded PCM, lens) for the `lens=` arguments below."""
        clips = [np.asarray(c).reshape(-1) for c in clips]
        lens = np.array([c.size for c in clips], np.int32)
        pcm = np.zeros((len(clips), int(lens.max())), clips[0].dtype)
        for b, c in enumerate(clips):
            pcm[b, :c.size] = c
        return pcm, lens

    def _lens(self, pcm: np.ndarray, lens) -> np.ndarray:
        lens = np.ascontiguousarray(np.asarray(lens, dtype=np.int32).reshape(-1))
        if lens.size != pcm.shape[0]:
            raise ValueError(f"lens has {lens.size} entries for a batch of {pcm.shape[0]}")
        return lens

    def encode(self, pcm: np.ndarray, query_ids: Sequence[int] = (), language_tail_ids: Sequence[int] = (), lens=None):
        """pcm [B][N]: int16, or float32 already in [-1,1] (audio_pcm_scale 32768).  Returns the prompt length; with `lens`
        (ragged batch: samples per clip, N = the longest) the list of per-clip prompt lengths."""
        pcm, code = self._pcm(pcm)
        q, l = _i32(query_ids), _i32(language_tail_ids)
        if lens is not None:
            lens = self._lens(pcm, lens)
            npr = np.zeros(pcm.shape[0], np.int32)
            self._ck(self.lib.b200asr_qwen_encode_ragged(self.h, pcm.ctypes.data_as(C.c_void_p), code, pcm.shape[0], pcm.shape[1],
                                                         lens.ctypes.data_as(_cabi._I32P), _ptr(q), q.size, _ptr(l), l.size,
                                                         npr.ctypes.data_as(_cabi._I32P)))
            self.batch, self.n_prompt = pcm.shape[0], int(npr.max())
            return npr.tolist()
        n = C.c_int32(0)
        self._ck(self.lib.b200asr_qwen_encode(self.h, pcm.ctypes.data_as(C.c_void_p), code, pcm.shape[0], pcm.shape[1], _ptr(q), q.size,
                                              _ptr(l), l.size, C.byref(n)))
        self.batch, self.n_prompt = pcm.shape[0], int(n.value)
        return self.n_prompt

    def prefill(self, want_logits: bool = True):
        lg = np.empty((self.batch, self.dims.vocab), np.float32) if want_logits else None
        tok = np.zeros(self.batch, np.int32)
        self._ck(self.lib.b200asr_qwen_prefill(self.h, lg.ctypes.data_as(_cabi._F32P) if want_logits else None, tok.ctypes.data_as(_cabi._I32P)))
        return lg, tok

    def decode_step(self, token_in: Optional[np.ndarray] = None, want_logits: bool = True):
        lg = np.empty((self.batch, self.dims.vocab), np.float32) if want_logits else None
        tok = np.zeros(self.batch, np.int32)
        tin = None if token_in is None else np.ascontiguousarray(np.asarray(token_in, dtype=np.int32).reshape(self.batch))
        self._ck(self.lib.b200asr_qwen_decode_step(self.h, None if tin is None else tin.ctypes.data_as(_cabi._I32P),
                                                   lg.ctypes.data_as(_cabi._F32P) if want_logits else None, tok.ctypes.data_as(_cabi._I32P)))
        return lg, tok

    def decode(self, max_new: int = -1) -> List[List[int]]:
        ld = self.dims.max_seq_len
        toks = np.zeros((self.batch, ld), np.int32)
        lens = np.zeros(self.batch, np.int32)
        self._ck(self.lib.b200asr_qwen_decode(self.h, max_new, toks.ctypes.data_as(_cabi._I32P), ld, lens.ctypes.data_as(_cabi._I32P)))
        return [toks[b, :lens[b]].tolist() for b in range(self.batch)]

    def transcribe(self, pcm: np.ndarray, query_ids: Sequence[int] = (), language_tail_ids: Sequence[int] = (), max_new: int = -1,
                   out_tokens: Optional[np.ndarray] = None, out_lens: Optional[np.ndarray] = None, lens=None) -> List[List[int]]:
        """`lens` (samples per clip, pcm padded to the longest) = a ragged batch: every clip gets the tokens it has alone."""
        pcm, code = self._pcm(pcm)
        q, l = _i32(query_ids), _i32(language_tail_ids)
        B = pcm.shape[0]
        ld = self.dims.max_seq_len
        toks = out_tokens if out_tokens is not None else np.zeros((B, ld), np.int32)
        clip_lens = None if lens is None else self._lens(pcm, lens)
        lens = out_lens if out_lens is not None else np.zeros(B, np.int32)
        if clip_lens is not None:
            self._ck(self.lib.b200asr_qwen_transcribe_ragged(self.h, pcm.ctypes.data_as(C.c_void_p), code, B, pcm.shape[1],
                                                             clip_lens.ctypes.data_as(_cabi._I32P), _ptr(q), q.size, _ptr(l), l.size,
                                                             max_new, toks.ctypes.data_as(_cabi._I32P), toks.shape[1],
                                                             lens.ctypes.data_as(_cabi._I32P)))
            self.batch = B
            return [toks[b, :lens[b]].tolist() for b in range(B)]
        self._ck(self.lib.b200asr_qwen_transcribe(self.h, pcm.ctypes.data_as(C.c_void_p), code, B, pcm.shape[1], _ptr(q), q.size, _ptr(l),
                                                  l.size, max_new, toks.ctypes.data_as(_cabi._I32P), toks.shape[1],
                                                  lens.ctypes.data_as(_cabi._I32P)))
        self.batch = B
        return [toks[b, :lens[b]].tolist() for b in range(B)]

    def upload(self, pcm: np.ndarray, lens=None):
        pcm, code = self._pcm(pcm)
        if lens is not None:
            lens = self._lens(pcm, lens)
            self._ck(self.lib.b200asr_qwen_upload_ragged(self.h, pcm.ctypes.data_as(C.c_void_p), code, pcm.shape[0], pcm.shape[1],
                                                         lens.ctypes.data_as(_cabi._I32P)))
        else:
            self._ck(self.lib.b200asr_qwen_upload(self.h, pcm.ctypes.data_as(C.c_void_p), code, pcm.shape[0], pcm.shape[1]))
        self.batch = pcm.shape[0]

    def transcribe_resident(self, query_ids: Sequence[int] = (), language_tail_ids: Sequence[int] = (), max_new: int = -1) -> List[List[int]]:
        q, l = _i32(query_ids), _i32(language_tail_ids)
        ld = self.dims.max_seq_len
        toks = np.zeros((self.batch, ld), np.int32)
        lens = np.zeros(self.batch, np.int32)
        self._ck(self.lib.b200asr_qwen_transcribe_resident(self.h, _ptr(q), q.size, _ptr(l), l.size, max_new,
                                                           toks.ctypes.data_as(_cabi._I32P), ld, lens.ctypes.data_as(_cabi._I32P)))
        return [toks[b, :lens[b]].tolist() for b in range(self.batch)]

    def get_stage(self, name: str, capacity: int) -> np.ndarray:
        out = np.empty(capacity, np.float32)
        n = C.c_int64(0)
        self._ck(self.lib.b200asr_qwen_get_stage(self.h, name.encode(), out.ctypes.data_as(_cabi._F32P), capacity, C.byref(n)))
        return out[:n.value]


REPEAT_PENALTY = 0.8         # script defaults (Inference_Qwen_ASR_ONNX.py:90-91): penalty-greedy out of the box
PENALTY_RANGE = 10


def transcribe_clip(engine: QwenEngine, raw_audio_int16: np.ndarray, *, query_ids: Sequence[int] = (),
                    language_tail_ids: Sequence[int] = (), sample_rate: int = 16000, step_through_host: bool = False,
                    repeat_penalty: float = REPEAT_PENALTY, penalty_range: int = PENALTY_RANGE):
    """Host loop of Inference_Qwen_ASR_ONNX.py:586-745 for one clip: int16 PCM in, token ids + timing out.
    `step_through_host` walks the script's protocol call by call (prefill, then one decode_step per token with the
    stop test and generation_limit on the host, :666-737); the default runs the same loop on the device."""
    pcm = np.asarray(raw_audio_int16, dtype=np.int16).reshape(1, -1)[:, :engine.max_samples]
    engine.set_decode_options(repeat_penalty, penalty_range)
    t0 = time.time()
    if not step_through_host:
        tokens = engine.transcribe(pcm, query_ids, language_tail_ids)[0]
    else:
        n_prompt = engine.encode(pcm, query_ids, language_tail_ids)
        limit = max(engine.dims.max_seq_len - 10 - n_prompt, 0)
        stop = set(int(s) for s in engine.prompt.stop_ids)
        tokens: List[int] = []
        if limit > 0:
            _, tok = engine.prefill(want_logits=False)
            sel = int(tok[0])
            count = 0
            if sel not in stop:
                count = 1
                tokens.append(sel)
            while count < limit and sel not in stop:
                _, tok = engine.decode_step(want_logits=False)
                sel = int(tok[0])
                if sel not in stop:
                    count += 1
                    tokens.append(sel)
    wall = time.time() - t0
    audio_s = pcm.shape[1] / float(sample_rate)
    return {"tokens": tokens, "wall_s": wall, "rtf": wall / audio_s if audio_s > 0 else float("inf")}


def transcribe_clips(engine: QwenEngine, clips: Sequence[np.ndarray], *, query_ids: Sequence[int] = (),
                     language_tail_ids: Sequence[int] = (), sample_rate: int = 16000, repeat_penalty: float = REPEAT_PENALTY,
                     penalty_range: int = PENALTY_RANGE, max_batch: Optional[int] = None):
    """`transcribe_clip` for a list of clips of any lengths: ragged batches of up to `max_batch` (default: the engine's) clips,
    neighbours in length together; every clip gets the tokens it gets alone (the script runs one clip per call,
    Inference_Qwen_ASR_ONNX.py:586-745).  Returns one {"tokens", "wall_s", "rtf"} per clip, in input order; a batch's wall time
    is shared by its clips in proportion to their audio."""
    from .sharding import ragged_batches
    clips = [np.asarray(c, dtype=np.int16).reshape(-1)[:engine.max_samples] for c in clips]
    nb = max(1, min(int(max_batch or engine.max_batch), engine.max_batch))
    engine.set_decode_options(repeat_penalty, penalty_range)
    out: List[Optional[dict]] = [None] * len(clips)
    for group in ragged_batches(range(len(clips)), [c.size for c in clips], nb):
        pcm, lens = QwenEngine.pad_ragged([clips[i] for i in group])
        t0 = time.time()
        toks = engine.transcribe(pcm, query_ids, language_tail_ids, lens=lens)
        wall = time.time() - t0
        total = float(lens.sum())
        for i, t, n in zip(group, toks, lens):
            share = wall * float(n) / total
            audio_s = float(n) / float(sample_rate)
            out[i] = {"tokens": t, "wall_s": share, "rtf": share / audio_s if audio_s > 0 else float("inf")}
    return out


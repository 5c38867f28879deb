"""Checkpoint -> engine tensor set.

The reference bakes three folds into the ONNX initialisers at export time
(/root/reference/Whisper/Export_Whisper.py:376-420 encoder, :527-550 decoder):
  * q/k/v fused into one Linear, with head_dim**-0.25 multiplied into the q and
    k rows (and the q bias) so no scale is applied at run time;
  * every in-layer LayerNorm affine absorbed into the Linear that follows it
    (:215-225:  b += W @ beta, then W *= gamma);
  * all decoder layers' cross-attention K (pre-scaled) and V projections
    concatenated into one [2*L*d, d] Linear applied to the final encoder state.
This module performs the same folds on an HF-named Whisper state dict and adds
the engine's own layout changes: Conv1d weights become [out, k*C_in + c] so the
conv stem runs as a GEMM over a strided view of the time-major activations.

It also reads/writes the flat weight-blob file the engine loads in deployment.
"""
from __future__ import annotations

import json
import math
import struct
from typing import Dict, Iterable, Optional, Sequence

import numpy as np

from .config import WhisperDims

SUPPRESS_VALUE = -128.0
BLOB_MAGIC = b"B200ASR1"


def _np(x) -> np.ndarray:
    if isinstance(x, np.ndarray):
        return x.astype(np.float32, copy=False)
    return x.detach().cpu().numpy().astype(np.float32, copy=False)   # torch tensor


def hann_dft_kernel(n_fft: int, input_scale: float = 1.0) -> np.ndarray:
    """[2F, n_fft] windowed DFT basis, built with the same fp32 op order as
    /root/reference/Whisper/STFT_Process.py:136-150 (torch is used so cos/sin
    match the exported initialiser bit for bit)."""
    import torch
    f_bins = n_fft // 2 + 1
    t = torch.arange(n_fft, dtype=torch.float32).unsqueeze(0)
    f = torch.arange(f_bins, dtype=torch.float32).unsqueeze(1)
    omega = (2.0 * torch.pi / n_fft) * f * t
    win = torch.hann_window(n_fft, periodic=True).float() * float(input_scale)
    k = torch.cat([torch.cos(omega) * win.unsqueeze(0), -torch.sin(omega) * win.unsqueeze(0)], dim=0)
    return k.contiguous().numpy()


def slaney_mel_filterbank(n_freqs: int, n_mels: int, sample_rate: int) -> np.ndarray:
    """[n_mels, n_freqs]; equals torchaudio melscale_fbanks(..., 'slaney', 'slaney').T used at
    /root/reference/Whisper/Export_Whisper.py:357-362 (computed through torchaudio so the
    fp32 rounding is the exporter's)."""
    import torchaudio
    fb = torchaudio.functional.melscale_fbanks(n_freqs, 0, sample_rate // 2, n_mels, sample_rate, "slaney", "slaney")
    return fb.transpose(0, 1).contiguous().numpy()


def _absorb(ln_w, ln_b, w, b):
    b = b + w @ ln_b
    w = w * ln_w[None, :]
    return w, b


def fold_whisper(state: Dict[str, object], dims: WhisperDims,
                 suppress_tokens: Optional[Sequence[int]] = None,
                 begin_suppress_tokens: Iterable[int] = (),
                 input_scale: float = 1.0) -> Dict[str, np.ndarray]:
    d, L = dims.d_model, dims.dec_layers
    scale = np.float32(float(dims.head_dim) ** -0.25)
    g = lambda k: _np(state[k])
    out: Dict[str, np.ndarray] = {}
    out["stft_kernel"] = hann_dft_kernel(dims.n_fft, input_scale)
    out["mel_fbank"] = slaney_mel_filterbank(dims.n_fft // 2 + 1, dims.n_mels, dims.sample_rate)

    def qkv(prefix):
        qw, kw, vw = g(prefix + "q_proj.weight"), g(prefix + "k_proj.weight"), g(prefix + "v_proj.weight")
        qb, vb = g(prefix + "q_proj.bias"), g(prefix + "v_proj.bias")
        kb = g(prefix + "k_proj.bias") if (prefix + "k_proj.bias") in state else np.zeros(d, np.float32)
        w = np.concatenate([qw * scale, kw * scale, vw], axis=0)
        b = np.concatenate([qb * scale, kb, vb], axis=0)
        return w, b

    e = "model.encoder."
    # Conv1d [out, in, k] -> GEMM weight [out, k*in + c] (time-major im2col view)
    out["enc.conv1.w"] = np.ascontiguousarray(g(e + "conv1.weight").transpose(0, 2, 1).reshape(d, -1))
    out["enc.conv1.b"] = g(e + "conv1.bias")
    out["enc.conv2.w"] = np.ascontiguousarray(g(e + "conv2.weight").transpose(0, 2, 1).reshape(d, -1))
    out["enc.conv2.b"] = g(e + "conv2.bias")
    out["enc.pos"] = g(e + "embed_positions.weight")
    for i in range(dims.enc_layers):
        p = f"{e}layers.{i}."
        w, b = qkv(p + "self_attn.")
        w, b = _absorb(g(p + "self_attn_layer_norm.weight"), g(p + "self_attn_layer_norm.bias"), w, b)
        out[f"enc.L{i}.qkv.w"], out[f"enc.L{i}.qkv.b"] = w, b
        out[f"enc.L{i}.out.w"], out[f"enc.L{i}.out.b"] = g(p + "self_attn.out_proj.weight"), g(p + "self_attn.out_proj.bias")
        w, b = _absorb(g(p + "final_layer_norm.weight"), g(p + "final_layer_norm.bias"), g(p + "fc1.weight"), g(p + "fc1.bias"))
        out[f"enc.L{i}.fc1.w"], out[f"enc.L{i}.fc1.b"] = w, b
        out[f"enc.L{i}.fc2.w"], out[f"enc.L{i}.fc2.b"] = g(p + "fc2.weight"), g(p + "fc2.bias")
    out["enc.ln_post.g"], out["enc.ln_post.b"] = g(e + "layer_norm.weight"), g(e + "layer_norm.bias")

    dd = "model.decoder."
    kws, kbs, vws, vbs = [], [], [], []
    for i in range(L):
        p = f"{dd}layers.{i}.encoder_attn."
        kws.append(g(p + "k_proj.weight") * scale)
        kb = g(p + "k_proj.bias") if (p + "k_proj.bias") in state else np.zeros(d, np.float32)
        kbs.append(kb * scale)
        vws.append(g(p + "v_proj.weight"))
        vbs.append(g(p + "v_proj.bias"))
    out["enc.cross_kv.w"] = np.concatenate(kws + vws, axis=0)
    out["enc.cross_kv.b"] = np.concatenate(kbs + vbs, axis=0)

    out["dec.embed"] = g(dd + "embed_tokens.weight")
    out["dec.pos"] = g(dd + "embed_positions.weight")
    for i in range(L):
        p = f"{dd}layers.{i}."
        w, b = qkv(p + "self_attn.")
        w, b = _absorb(g(p + "self_attn_layer_norm.weight"), g(p + "self_attn_layer_norm.bias"), w, b)
        out[f"dec.L{i}.qkv.w"], out[f"dec.L{i}.qkv.b"] = w, b
        out[f"dec.L{i}.out.w"], out[f"dec.L{i}.out.b"] = g(p + "self_attn.out_proj.weight"), g(p + "self_attn.out_proj.bias")
        w, b = _absorb(g(p + "encoder_attn_layer_norm.weight"), g(p + "encoder_attn_layer_norm.bias"),
                       g(p + "encoder_attn.q_proj.weight") * scale, g(p + "encoder_attn.q_proj.bias") * scale)
        out[f"dec.L{i}.cq.w"], out[f"dec.L{i}.cq.b"] = w, b
        out[f"dec.L{i}.cout.w"], out[f"dec.L{i}.cout.b"] = g(p + "encoder_attn.out_proj.weight"), g(p + "encoder_attn.out_proj.bias")
        w, b = _absorb(g(p + "final_layer_norm.weight"), g(p + "final_layer_norm.bias"), g(p + "fc1.weight"), g(p + "fc1.bias"))
        out[f"dec.L{i}.fc1.w"], out[f"dec.L{i}.fc1.b"] = w, b
        out[f"dec.L{i}.fc2.w"], out[f"dec.L{i}.fc2.b"] = g(p + "fc2.weight"), g(p + "fc2.bias")
    out["dec.ln.g"], out["dec.ln.b"] = g(dd + "layer_norm.weight"), g(dd + "layer_norm.bias")
    sup = np.zeros(dims.vocab, np.float32)
    if suppress_tokens is not None and len(suppress_tokens):
        sup[np.asarray(list(suppress_tokens), dtype=np.int64)] = SUPPRESS_VALUE
    out["dec.suppress_bias"] = sup
    beg = np.zeros(dims.vocab, np.float32)
    ids = [int(t) for t in begin_suppress_tokens if 0 <= int(t) < dims.vocab]
    if ids:
        beg[ids] = -np.inf
    out["dec.begin_suppress_bias"] = beg
    return {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in out.items()}


# --------------------------------------------------------------------------
# flat weight blob:  magic | u64 header_len | JSON header | 64B-aligned fp32 data
# --------------------------------------------------------------------------
def save_blob(path, tensors: Dict[str, np.ndarray], dims: WhisperDims, metadata: Optional[dict] = None) -> None:
    table, offset = {}, 0
    for name, arr in tensors.items():
        offset = (offset + 63) // 64 * 64
        table[name] = {"shape": list(arr.shape), "offset": offset, "numel": int(arr.size)}
        offset += arr.size * 4
    header = json.dumps({"dims": dims.to_dict(), "tensors": table, "metadata": metadata or {}}).encode()
    with open(path, "wb") as f:
        f.write(BLOB_MAGIC)
        f.write(struct.pack("<Q", len(header)))
        f.write(header)
        base = f.tell()
        base_pad = (base + 63) // 64 * 64
        f.write(b"\0" * (base_pad - base))
        pos = 0
        for name, arr in tensors.items():
            off = table[name]["offset"]
            f.write(b"\0" * (off - pos))
            f.write(np.ascontiguousarray(arr, dtype=np.float32).tobytes())
            pos = off + arr.size * 4


def load_blob(path):
    """Returns (dims, {name: np.memmap view}, metadata); tensors are mapped, not copied."""
    with open(path, "rb") as f:
        if f.read(8) != BLOB_MAGIC:
            raise ValueError(f"{path}: not a b200asr weight blob")
        (hlen,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(hlen))
        base = (f.tell() + 63) // 64 * 64
    mm = np.memmap(path, dtype=np.uint8, mode="r")
    tensors = {}
    for name, ent in header["tensors"].items():
        start = base + ent["offset"]
        tensors[name] = mm[start:start + ent["numel"] * 4].view(np.float32).reshape(ent["shape"])
    return WhisperDims(**header["dims"]), tensors, header.get("metadata", {})

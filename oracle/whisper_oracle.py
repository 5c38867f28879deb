"""CPU oracle for the Whisper hot path (TEST INFRASTRUCTURE ONLY).

This file is a plain fp32 restatement, on the CPU, of the arithmetic that the
reference's exported Whisper graphs perform inside ``InferenceSession.run()``:
STFT power -> log-mel -> conv stem -> encoder layers -> fused cross-KV ->
embed/position -> decoder layers with a growing self-KV cache -> logits ->
begin-suppress / argmax heads, plus the greedy host loop around them.

It is the *checker*.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import it.  The
product (``automatic-speech-recognition-asr-onnx_b200``) never does.

Pinning status: the reference ships no golden vectors or tests for this path
(SURVEY.md section 4 / 8c).  The oracle is pinned instead against the
reference's own graph-defining ``torch.nn.Module`` wrappers
(``WHISPER_ENCODER`` / ``WHISPER_DECODER`` / ``WHISPER_PREFILL`` / ... AST-
extracted from /root/reference/Whisper/Export_Whisper.py and executed in the
dev container by ``oracle/gen_golden.py``); the resulting vectors are
committed under ``tests/golden/`` and re-checked by
``tests/test_oracle_golden.py``.  ONNX Runtime itself is not installed, so the
ORT-vs-nn.Module deltas (tanh-GELU approximation option, denormal flush, fp16
KV default) are NOT pinned -- see DESIGN.md "Oracle".

Every function cites the reference lines it follows (paths relative to
/root/reference).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

LN_EPS = 1e-5          # HF Whisper nn.LayerNorm default (modules read norm.eps)
MASK_VALUE = -128.0    # Whisper/Export_Whisper.py:471-474 (causal mask fill)
SUPPRESS_VALUE = -128.0  # Whisper/Export_Whisper.py:517-520


@dataclass(frozen=True)
class WhisperDims:
    n_mels: int = 128
    d_model: int = 1280
    n_heads: int = 20
    ffn: int = 5120
    enc_layers: int = 32
    dec_layers: int = 32
    vocab: int = 51866
    max_source: int = 1500
    max_target: int = 448
    n_fft: int = 400
    hop: int = 160
    sample_rate: int = 16000

    @property
    def head_dim(self) -> int:
        return self.d_model // self.n_heads

    def to_dict(self):
        return asdict(self)


LARGE_V3 = WhisperDims()
TINY_TEST = WhisperDims(n_mels=128, d_model=256, n_heads=4, ffn=512, enc_layers=2,
                        dec_layers=2, vocab=1000)


# --------------------------------------------------------------------------
# Seeded synthetic checkpoint (HF WhisperForConditionalGeneration key names)
# --------------------------------------------------------------------------
def make_raw_weights(dims: WhisperDims, seed: int, pos_scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """Deterministic random checkpoint with HF state-dict names.

    pos_scale multiplies the decoder's learned position table after the draw (the draw order never changes).  With
    the default scales a random decoder is in the "ordered" regime: its output barely depends on the fed-back token
    or the position, so greedy streams repeat one id; pos_scale = 100 makes the position term comparable to the
    32 layers of residual updates and the streams take 20+ distinct ids in 33 steps, without scaling the weight
    matrices (which would push the net into the chaotic regime where rounding errors explode).

    There are no real checkpoints offline, so weights are drawn
    here.  Scales are chosen so activations stay O(1): Linear ~ N(0, 1/fan_in),
    LayerNorm gamma ~ 1 + 0.1 N, beta ~ 0.1 N (non-trivial so the LN-affine
    folds are exercised), token embedding ~ 0.05 N (logit std ~ 1.8).
    """
    g = torch.Generator().manual_seed(int(seed))
    d, f = dims.d_model, dims.ffn
    w: Dict[str, torch.Tensor] = {}

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std

    def linear(prefix, out_f, in_f, bias=True):
        w[prefix + ".weight"] = rn(out_f, in_f, std=1.0 / math.sqrt(in_f))
        if bias:
            w[prefix + ".bias"] = rn(out_f, std=0.1)

    def norm(prefix):
        w[prefix + ".weight"] = 1.0 + rn(d, std=0.1)
        w[prefix + ".bias"] = rn(d, std=0.1)

    def attn(prefix):
        linear(prefix + ".k_proj", d, d, bias=False)   # HF Whisper: k_proj has no bias
        linear(prefix + ".v_proj", d, d)
        linear(prefix + ".q_proj", d, d)
        linear(prefix + ".out_proj", d, d)

    e = "model.encoder."
    w[e + "conv1.weight"] = rn(d, dims.n_mels, 3, std=1.0 / math.sqrt(3 * dims.n_mels))
    w[e + "conv1.bias"] = rn(d, std=0.1)
    w[e + "conv2.weight"] = rn(d, d, 3, std=1.0 / math.sqrt(3 * d))
    w[e + "conv2.bias"] = rn(d, std=0.1)
    w[e + "embed_positions.weight"] = rn(dims.max_source, d, std=0.1)
    for i in range(dims.enc_layers):
        p = f"{e}layers.{i}."
        attn(p + "self_attn")
        norm(p + "self_attn_layer_norm")
        linear(p + "fc1", f, d)
        linear(p + "fc2", d, f)
        norm(p + "final_layer_norm")
    norm(e + "layer_norm")

    dd = "model.decoder."
    w[dd + "embed_tokens.weight"] = rn(dims.vocab, d, std=0.05)
    w[dd + "embed_positions.weight"] = rn(dims.max_target, d, std=0.05) * float(pos_scale)
    for i in range(dims.dec_layers):
        p = f"{dd}layers.{i}."
        attn(p + "self_attn")
        norm(p + "self_attn_layer_norm")
        attn(p + "encoder_attn")
        norm(p + "encoder_attn_layer_norm")
        linear(p + "fc1", f, d)
        linear(p + "fc2", d, f)
        norm(p + "final_layer_norm")
    norm(dd + "layer_norm")
    w["proj_out.weight"] = w[dd + "embed_tokens.weight"]   # tied (Shared_Merged.py:285-300)
    return w


# --------------------------------------------------------------------------
# Front-end constants
# --------------------------------------------------------------------------
def stft_kernel(n_fft: int = 400, input_scale: float = 1.0) -> torch.Tensor:
    """Windowed DFT basis [2*(n_fft/2+1), n_fft]: rows 0..F-1 = hann*cos, rows F.. = -hann*sin.

    Follows Whisper/STFT_Process.py:136-150 (omega built in fp32 as
    (2*pi/n_fft) * f * t; periodic Hann; input_scale folded into the window).
    """
    f_bins = n_fft // 2 + 1
    omega_factor = 2.0 * torch.pi / n_fft
    t = torch.arange(n_fft, dtype=torch.float32).unsqueeze(0)
    f = torch.arange(f_bins, dtype=torch.float32).unsqueeze(1)
    omega = omega_factor * f * t
    window = torch.hann_window(n_fft, periodic=True).float() * float(input_scale)
    wc = torch.cos(omega) * window.unsqueeze(0)
    ws = -torch.sin(omega) * window.unsqueeze(0)
    return torch.cat([wc, ws], dim=0).contiguous()


def mel_filterbank(n_freqs: int, n_mels: int, sample_rate: int) -> torch.Tensor:
    """Slaney-scale, slaney-normalised triangular filterbank, shape [n_mels, n_freqs].

    Restates torchaudio.functional.melscale_fbanks(n_freqs, 0, sr/2, n_mels, sr,
    'slaney', 'slaney').T as used at Whisper/Export_Whisper.py:357-362
    (torchaudio is a third-party dependency of the reference; algorithm =
    librosa/Slaney Auditory Toolbox mel: linear below 1 kHz, log above).
    """
    f_min, f_max = 0.0, float(sample_rate // 2)
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)

    def hz_to_mel(freq: float) -> float:
        f_sp = 200.0 / 3
        mels = freq / f_sp
        min_log_hz = 1000.0
        min_log_mel = min_log_hz / f_sp
        logstep = math.log(6.4) / 27.0
        if freq >= min_log_hz:
            mels = min_log_mel + math.log(freq / min_log_hz) / logstep
        return mels

    def mel_to_hz(mels: torch.Tensor) -> torch.Tensor:
        f_sp = 200.0 / 3
        freqs = f_sp * mels
        min_log_hz = 1000.0
        min_log_mel = min_log_hz / f_sp
        logstep = math.log(6.4) / 27.0
        log_t = mels >= min_log_mel
        freqs[log_t] = min_log_hz * torch.exp(logstep * (mels[log_t] - min_log_mel))
        return freqs

    m_pts = torch.linspace(hz_to_mel(f_min), hz_to_mel(f_max), n_mels + 2)
    f_pts = mel_to_hz(m_pts)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)          # [n_freqs, n_mels+2]
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(torch.zeros(1), torch.min(down, up))           # [n_freqs, n_mels]
    enorm = 2.0 / (f_pts[2 : n_mels + 2] - f_pts[:n_mels])
    fb = fb * enorm.unsqueeze(0)
    return fb.t().contiguous().float()


# --------------------------------------------------------------------------
# Weight folding (restates WHISPER_ENCODER/_DECODER._fuse_weights)
# --------------------------------------------------------------------------
def _absorb_ln(ln_w, ln_b, lin_w, lin_b):
    """Whisper/Export_Whisper.py:215-225: b += W @ beta (pre-scaled W), then W *= gamma."""
    lin_b = lin_b + torch.matmul(lin_w, ln_b)
    lin_w = lin_w * ln_w.unsqueeze(0)
    return lin_w, lin_b


def _fused_qkv(raw, p, d, scale):
    """Whisper/Export_Whisper.py:381-389 / 534-542."""
    qw, kw, vw = raw[p + "q_proj.weight"], raw[p + "k_proj.weight"], raw[p + "v_proj.weight"]
    qb = raw[p + "q_proj.bias"]
    kb = raw.get(p + "k_proj.bias", torch.zeros(d))
    vb = raw[p + "v_proj.bias"]
    w = torch.cat([qw, kw, vw], dim=0).clone()
    b = torch.cat([qb, kb, vb], dim=0).clone()
    w[: 2 * d] *= scale          # q and k weights
    b[:d] *= scale               # q bias only (":388": k bias is not scaled; it is zero in Whisper)
    return w, b


def fold_weights(raw: Dict[str, torch.Tensor], dims: WhisperDims,
                 suppress_tokens: Optional[Sequence[int]] = None,
                 begin_suppress_tokens: Sequence[int] = (),
                 input_scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """Folded tensors in the reference's own layouts (conv weights [out, in, k])."""
    d, L = dims.d_model, dims.dec_layers
    scale = float(dims.head_dim ** -0.25)
    fw: Dict[str, torch.Tensor] = {}
    fw["stft_kernel"] = stft_kernel(dims.n_fft, input_scale)
    fw["mel_fbank"] = mel_filterbank(dims.n_fft // 2 + 1, dims.n_mels, dims.sample_rate)
    e = "model.encoder."
    for k in ("conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias"):
        fw["enc." + k] = raw[e + k].clone()
    fw["enc.pos"] = raw[e + "embed_positions.weight"].clone()
    for i in range(dims.enc_layers):
        p = f"{e}layers.{i}."
        w, b = _fused_qkv(raw, p + "self_attn.", d, scale)
        w, b = _absorb_ln(raw[p + "self_attn_layer_norm.weight"], raw[p + "self_attn_layer_norm.bias"], w, b)
        fw[f"enc.L{i}.qkv.w"], fw[f"enc.L{i}.qkv.b"] = w, b
        fw[f"enc.L{i}.out.w"] = raw[p + "self_attn.out_proj.weight"].clone()
        fw[f"enc.L{i}.out.b"] = raw[p + "self_attn.out_proj.bias"].clone()
        w, b = _absorb_ln(raw[p + "final_layer_norm.weight"], raw[p + "final_layer_norm.bias"],
                          raw[p + "fc1.weight"], raw[p + "fc1.bias"])
        fw[f"enc.L{i}.fc1.w"], fw[f"enc.L{i}.fc1.b"] = w, b
        fw[f"enc.L{i}.fc2.w"] = raw[p + "fc2.weight"].clone()
        fw[f"enc.L{i}.fc2.b"] = raw[p + "fc2.bias"].clone()
    fw["enc.ln_post.g"] = raw[e + "layer_norm.weight"].clone()
    fw["enc.ln_post.b"] = raw[e + "layer_norm.bias"].clone()
    # fused cross-KV: all K projections (pre-scaled) then all V (Export_Whisper.py:393-417)
    dd = "model.decoder."
    kws, kbs, vws, vbs = [], [], [], []
    for i in range(L):
        p = f"{dd}layers.{i}.encoder_attn."
        kws.append(raw[p + "k_proj.weight"] * scale)
        kbs.append(raw.get(p + "k_proj.bias", torch.zeros(d)) * scale)
        vws.append(raw[p + "v_proj.weight"])
        vbs.append(raw[p + "v_proj.bias"])
    fw["enc.cross_kv.w"] = torch.cat(kws + vws, dim=0)
    fw["enc.cross_kv.b"] = torch.cat(kbs + vbs, dim=0)
    # decoder (Export_Whisper.py:527-550)
    fw["dec.embed"] = raw[dd + "embed_tokens.weight"].clone()
    fw["dec.pos"] = raw[dd + "embed_positions.weight"].clone()
    for i in range(L):
        p = f"{dd}layers.{i}."
        w, b = _fused_qkv(raw, p + "self_attn.", d, scale)
        w, b = _absorb_ln(raw[p + "self_attn_layer_norm.weight"], raw[p + "self_attn_layer_norm.bias"], w, b)
        fw[f"dec.L{i}.qkv.w"], fw[f"dec.L{i}.qkv.b"] = w, b
        fw[f"dec.L{i}.out.w"] = raw[p + "self_attn.out_proj.weight"].clone()
        fw[f"dec.L{i}.out.b"] = raw[p + "self_attn.out_proj.bias"].clone()
        w, b = _absorb_ln(raw[p + "encoder_attn_layer_norm.weight"], raw[p + "encoder_attn_layer_norm.bias"],
                          raw[p + "encoder_attn.q_proj.weight"] * scale, raw[p + "encoder_attn.q_proj.bias"] * scale)
        fw[f"dec.L{i}.cq.w"], fw[f"dec.L{i}.cq.b"] = w, b
        fw[f"dec.L{i}.cout.w"] = raw[p + "encoder_attn.out_proj.weight"].clone()
        fw[f"dec.L{i}.cout.b"] = raw[p + "encoder_attn.out_proj.bias"].clone()
        w, b = _absorb_ln(raw[p + "final_layer_norm.weight"], raw[p + "final_layer_norm.bias"],
                          raw[p + "fc1.weight"], raw[p + "fc1.bias"])
        fw[f"dec.L{i}.fc1.w"], fw[f"dec.L{i}.fc1.b"] = w, b
        fw[f"dec.L{i}.fc2.w"] = raw[p + "fc2.weight"].clone()
        fw[f"dec.L{i}.fc2.b"] = raw[p + "fc2.bias"].clone()
    fw["dec.ln.g"] = raw[dd + "layer_norm.weight"].clone()
    fw["dec.ln.b"] = raw[dd + "layer_norm.bias"].clone()
    sup = torch.zeros(dims.vocab)
    if suppress_tokens is not None and len(suppress_tokens):
        sup[torch.as_tensor(list(suppress_tokens), dtype=torch.long)] = SUPPRESS_VALUE
    fw["dec.suppress_bias"] = sup                                  # Export_Whisper.py:517-520
    beg = torch.zeros(dims.vocab)
    ids = [int(t) for t in begin_suppress_tokens if 0 <= int(t) < dims.vocab]
    if ids:
        beg[ids] = float("-inf")                                   # Export_Whisper.py:228-237
    fw["dec.begin_suppress_bias"] = beg
    return fw


# --------------------------------------------------------------------------
# Forward passes
# --------------------------------------------------------------------------
def stft_power(audio: torch.Tensor, kernel: torch.Tensor, n_fft: int, hop: int) -> torch.Tensor:
    """audio [1,1,N] -> power [1, n_fft/2+1, N//hop].

    Whisper/STFT_Process.py:96-102 (right pad shortened by one hop = drop last
    frame), :224-246 (reflect pad, Conv1d, sum of squares of packed re/im).
    """
    half = n_fft // 2
    right = half - hop
    left_pad = audio[..., 1 : half + 1].flip(2)
    if right:
        right_pad = audio[..., -(right + 1) : -1].flip(2)
        x = torch.cat([left_pad, audio, right_pad], dim=2)
    else:
        x = torch.cat([left_pad, audio], dim=2)
    packed = F.conv1d(x, kernel.unsqueeze(1), stride=hop)
    packed = packed.reshape(1, 2, half + 1, -1)
    return torch.sum(packed * packed, dim=1)


def log_mel(power: torch.Tensor, fbank: torch.Tensor) -> torch.Tensor:
    """Whisper/Export_Whisper.py:425-427: log10(clamp(fbank@power,1e-10)), max(x, x.max()-8), (x+4)/4."""
    mel = torch.matmul(fbank.unsqueeze(0), power).clamp(min=1e-10).log10()
    mel = torch.maximum(mel, mel.max() - 8.0)
    return (mel + 4.0) * 0.25


def _ln(x: torch.Tensor) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), eps=LN_EPS)


def encoder(audio: torch.Tensor, fw: Dict[str, torch.Tensor], dims: WhisperDims,
            keep_stages: bool = False):
    """audio f32 [1,1,N] (already divided by 32768 unless input_scale was folded).

    Returns (keys, values, stages): keys[l] (H, dh, T), values[l] (H, T, dh) --
    Whisper/Export_Whisper.py:422-447.
    """
    H, dh, d, L = dims.n_heads, dims.head_dim, dims.d_model, dims.dec_layers
    stages = {}
    power = stft_power(audio.float(), fw["stft_kernel"], dims.n_fft, dims.hop)
    mel = log_mel(power, fw["mel_fbank"])
    h = F.gelu(F.conv1d(mel, fw["enc.conv1.weight"], fw["enc.conv1.bias"], padding=1))
    h = F.gelu(F.conv1d(h, fw["enc.conv2.weight"], fw["enc.conv2.bias"], stride=2, padding=1)).transpose(1, 2)
    h = h + fw["enc.pos"][: h.shape[1]]
    if keep_stages:
        stages.update(power=power, mel=mel, stem=h.clone())
    for i in range(dims.enc_layers):
        qkv = F.linear(_ln(h), fw[f"enc.L{i}.qkv.w"], fw[f"enc.L{i}.qkv.b"])
        qkv = qkv.view(-1, 3 * H, dh).transpose(0, 1)
        q, k, v = qkv.split(H, dim=0)
        a = torch.matmul(torch.softmax(torch.matmul(q, k.transpose(1, 2)), dim=-1), v)
        a = a.transpose(0, 1).reshape(1, -1, d)
        ha = F.linear(a, fw[f"enc.L{i}.out.w"], fw[f"enc.L{i}.out.b"]) + h
        f1 = F.gelu(F.linear(_ln(ha), fw[f"enc.L{i}.fc1.w"], fw[f"enc.L{i}.fc1.b"]))
        h = ha + F.linear(f1, fw[f"enc.L{i}.fc2.w"], fw[f"enc.L{i}.fc2.b"])
        if keep_stages and (i == 0 or i == dims.enc_layers - 1):
            stages[f"enc_layer{i}"] = h.clone()
    h = F.layer_norm(h, (d,), fw["enc.ln_post.g"], fw["enc.ln_post.b"], LN_EPS)
    if keep_stages:
        stages["enc_out"] = h.clone()
    ckv = F.linear(h, fw["enc.cross_kv.w"], fw["enc.cross_kv.b"])
    keys, values = ckv.split(L * d, dim=-1)
    keys = keys.reshape(-1, L, H, dh).permute(1, 2, 3, 0).unbind(0)
    values = values.reshape(-1, L, H, dh).permute(1, 2, 0, 3).unbind(0)
    return list(keys), list(values), stages


def decoder(token_ids: torch.Tensor, history_len: int,
            self_k: List[torch.Tensor], self_v: List[torch.Tensor],
            cross_k: List[torch.Tensor], cross_v: List[torch.Tensor],
            fw: Dict[str, torch.Tensor], dims: WhisperDims):
    """One decoder launch on ``n`` new tokens (prefill n>=1 with causal mask, decode n=1).

    token_ids int [B, n]; self_k[l] [B,H,dh,hist], self_v[l] [B,H,hist,dh].
    Embedding: Export_Whisper.py:450-458; positions + mask: :461-497 (mask =
    triu(-128, diagonal=1) sliced [:n, :kv]; decode mask is zeros);
    layers: :614-667.  Returns (new_self_k, new_self_v, logits[B, vocab]).
    """
    H, dh, d = dims.n_heads, dims.head_dim, dims.d_model
    B, n = token_ids.shape
    kv = history_len + n
    x = fw["dec.embed"][token_ids.long()] + fw["dec.pos"][history_len:kv].unsqueeze(0)
    if n > 1:
        mask = torch.triu(torch.full((1, dims.max_target, dims.max_target), MASK_VALUE), diagonal=1)[:, :n, :kv]
    else:
        mask = torch.zeros(1, 1, 1)
    new_k, new_v = [], []
    for i in range(dims.dec_layers):
        qkv = F.linear(_ln(x), fw[f"dec.L{i}.qkv.w"], fw[f"dec.L{i}.qkv.b"])
        qkv = qkv.view(B, -1, 3 * H, dh).transpose(1, 2)
        q, k, v = qkv.split(H, dim=1)
        k = torch.cat((self_k[i], k.transpose(-1, -2)), dim=-1)
        v = torch.cat((self_v[i], v), dim=-2)
        new_k.append(k)
        new_v.append(v)
        a = torch.matmul(torch.softmax(torch.matmul(q, k) + mask, dim=-1), v)
        a = a.transpose(1, 2).reshape(B, -1, d)
        xa = F.linear(a, fw[f"dec.L{i}.out.w"], fw[f"dec.L{i}.out.b"]) + x
        q = F.linear(_ln(xa), fw[f"dec.L{i}.cq.w"], fw[f"dec.L{i}.cq.b"]).view(B, -1, H, dh).transpose(1, 2)
        a = torch.matmul(torch.softmax(torch.matmul(q, cross_k[i]), dim=-1), cross_v[i])
        xc = F.linear(a.transpose(1, 2).reshape(B, -1, d), fw[f"dec.L{i}.cout.w"], fw[f"dec.L{i}.cout.b"]) + xa
        f1 = F.gelu(F.linear(_ln(xc), fw[f"dec.L{i}.fc1.w"], fw[f"dec.L{i}.fc1.b"]))
        x = xc + F.linear(f1, fw[f"dec.L{i}.fc2.w"], fw[f"dec.L{i}.fc2.b"])
    last = F.layer_norm(x[:, -1], (d,), fw["dec.ln.g"], fw["dec.ln.b"], LN_EPS)
    logits = F.linear(last, fw["dec.embed"]) + fw["dec.suppress_bias"]
    return new_k, new_v, logits


def empty_self_kv(dims: WhisperDims, batch: int = 1):
    """Export_Whisper.py:841-845: K (B,H,dh,0), V (B,H,0,dh)."""
    k = [torch.zeros(batch, dims.n_heads, dims.head_dim, 0) for _ in range(dims.dec_layers)]
    v = [torch.zeros(batch, dims.n_heads, 0, dims.head_dim) for _ in range(dims.dec_layers)]
    return k, v


# --------------------------------------------------------------------------
# Heads (Export_Whisper.py:228-348)
# --------------------------------------------------------------------------
def begin_suppress(logits, fw):
    return logits + fw["dec.begin_suppress_bias"]


def argmax_head(logits) -> torch.Tensor:
    return torch.argmax(logits, dim=-1, keepdim=True).int()


def apply_penalty(logits, save_id, penalty_value: float, penalty_range: int):
    """Export_Whisper.py:318-331: multiply logits of the last penalty_range saved ids."""
    idx = save_id[:, -penalty_range:].long()
    pen = logits.gather(1, idx) * penalty_value
    return logits.scatter(1, idx, pen)


def topk_topp_sample(logits, temperature: float, top_k: int, top_p: float, repetition_penalty: float,
                     previous_ids, noise):
    """TOPK_TOPP_SAMPLING.forward (Whisper/Export_Whisper.py:281-307) with the uniform noise supplied by the
    caller instead of drawn by torch.rand_like (:298-300), so a run is reproducible: HF-style repetition penalty on
    every previously selected id (negative logits * p, positive / p, gather-then-scatter), / temperature, sorted
    top-k, softmax, keep while (cumsum - prob) <= top_p, Gumbel-max over the kept scores.
    logits [B, vocab], previous_ids int [B, n], noise [B, top_k] in (0, 1).  Returns (sampled_id [B,1] int32, save_id)."""
    rp = torch.tensor(float(repetition_penalty), dtype=torch.float32)
    prev = previous_ids.long()
    prev_logits = torch.gather(logits, 1, prev)
    prev_scores = torch.where(prev_logits < 0.0, prev_logits * rp, prev_logits * torch.reciprocal(rp))
    scores = torch.scatter(logits, 1, prev, prev_scores)
    scores = scores * torch.reciprocal(torch.tensor(float(temperature), dtype=torch.float32))
    sorted_scores, sorted_indices = torch.topk(scores, k=int(top_k), dim=-1, largest=True, sorted=True)
    probs = torch.softmax(sorted_scores, dim=-1)
    cums = torch.cumsum(probs, dim=-1)
    keep = (cums - probs) <= top_p
    sorted_scores = torch.where(keep, sorted_scores, torch.tensor(float("-inf")))
    u = torch.clamp(torch.as_tensor(noise, dtype=torch.float32), 1.0e-7, 1.0 - 1.0e-7)
    gumbel = -torch.log(-torch.log(u))
    winner = torch.argmax(sorted_scores + gumbel, dim=-1, keepdim=True)
    sampled = torch.gather(sorted_indices, 1, winner).int()
    return sampled, torch.cat([previous_ids.int(), sampled], dim=-1)


def sampling_transcribe(pcm_int16: np.ndarray, fw, dims: "WhisperDims", prompt: Sequence[int], stop_tokens: Sequence[int],
                        max_new: int, temperature: float, top_k: int, top_p: float, repetition_penalty: float,
                        noise: np.ndarray):
    """The `sampling` strategy of the driver (Inference_Whisper_ONNX.py:294-304, graphs Shared_Merged.py:925-975):
    begin-suppress + sampling head on the prefill logits with an empty history, then one sampling head per decode
    launch fed with every id selected so far.  noise [launch][top_k]."""
    stop = set(int(s) for s in stop_tokens)
    audio = prepare_audio(pcm_int16)
    ck, cv, _ = encoder(audio, fw, dims)
    sk, sv = empty_self_kv(dims)
    ids = torch.tensor([list(prompt)], dtype=torch.int32)
    limit = min(max(0, dims.max_target - ids.shape[-1]), int(max_new))
    sk, sv, logits = decoder(ids, 0, sk, sv, ck, cv, fw, dims)
    kv_len = ids.shape[-1]
    save_id = torch.zeros(1, 0, dtype=torch.int32)
    sampled, save_id = topk_topp_sample(begin_suppress(logits, fw), temperature, top_k, top_p, repetition_penalty, save_id,
                                        noise[0:1])
    selected = int(sampled[0, 0])
    tokens, generated, step = [], 0, 0
    if selected not in stop and limit > 0:
        generated = 1
        tokens.append(selected)
    while generated < limit and selected not in stop:
        sk, sv, logits = decoder(torch.tensor([[selected]], dtype=torch.int32), kv_len, sk, sv, ck, cv, fw, dims)
        kv_len += 1
        step += 1
        sampled, save_id = topk_topp_sample(logits, temperature, top_k, top_p, repetition_penalty, save_id, noise[step:step + 1])
        selected = int(sampled[0, 0])
        if selected not in stop:
            generated += 1
            tokens.append(selected)
    return dict(tokens=tokens, selected=save_id[0].tolist())


def no_speech_prob(logits, suppress_tokens, no_speech_token: int) -> torch.Tensor:
    """Export_Whisper.py:334-348."""
    unsup = torch.zeros(1, logits.shape[-1])
    if suppress_tokens is not None and len(suppress_tokens):
        unsup[:, torch.as_tensor(list(suppress_tokens), dtype=torch.long)] = 128.0
    return torch.softmax(logits + unsup, dim=-1)[:, no_speech_token]


def detect_language(logits_row: np.ndarray, language_token_ids: Sequence[int]) -> int:
    """Whisper/Inference_Whisper_ONNX.py:794-797."""
    ids = np.asarray(list(language_token_ids), dtype=np.int64)
    return int(ids[np.argmax(logits_row[ids])])


# --------------------------------------------------------------------------
# Host loop (Whisper/Inference_Whisper_ONNX.py:584-663, 766-827), greedy strategy
# --------------------------------------------------------------------------
def prepare_audio(pcm_int16: np.ndarray, audio_pcm_scale: int = 32768) -> torch.Tensor:
    """Inference_Whisper_ONNX.py:103-126 (F32 input, USE_NORMALISE_AUDIO=False)."""
    a = pcm_int16.astype(np.float32)
    a *= np.float32(1.0 / audio_pcm_scale)
    return torch.from_numpy(np.ascontiguousarray(a)).reshape(1, 1, -1)


def greedy_transcribe(pcm_int16: np.ndarray, fw, dims: WhisperDims, prompt: Sequence[int],
                      stop_tokens: Sequence[int], max_new: Optional[int] = None,
                      forced_tokens: Optional[Sequence[int]] = None,
                      repeat_penalty: float = 1.0, penalty_range: int = 20,
                      return_logits: bool = True):
    """1 encoder + 1 prefill + N decode launches (DETECT_LANGUAGE=False,
    NO_SPEECH_DETECTION=False).  repeat_penalty == 1.0 -> `greedy` graphs,
    otherwise `penalty_greedy` (penalty_value bound to 1.0 until
    generated_count >= PENALTY_RANGE, Inference_Whisper_ONNX.py:629-633).

    ``forced_tokens`` teacher-forces the fed-back token stream
    while still recording the argmax the graph would have produced.
    Returns dict(tokens, step_logits, selected).
    """
    stop = set(int(s) for s in stop_tokens)
    audio = prepare_audio(pcm_int16)
    ck, cv, _ = encoder(audio, fw, dims)
    sk, sv = empty_self_kv(dims)
    ids = torch.tensor([list(prompt)], dtype=torch.int32)
    limit = max(0, dims.max_target - ids.shape[-1])
    if max_new is not None:
        limit = min(limit, int(max_new))
    step_logits, selected_all, tokens = [], [], []
    use_pen = repeat_penalty != 1.0
    sk, sv, logits = decoder(ids, 0, sk, sv, ck, cv, fw, dims)
    kv_len = ids.shape[-1]
    if return_logits:
        step_logits.append(logits[0].numpy().copy())
    head = begin_suppress(logits, fw)
    save_id = torch.zeros(1, 0, dtype=torch.int32)
    selected = int(argmax_head(head)[0, 0])
    save_id = torch.cat([save_id, torch.tensor([[selected]], dtype=torch.int32)], dim=-1)
    selected_all.append(selected)
    generated = 0
    if selected not in stop and limit > 0:
        generated = 1
        tokens.append(selected)
    step = 0
    while generated < limit and selected not in stop:
        feed = selected if forced_tokens is None else int(forced_tokens[step])
        if forced_tokens is not None:
            save_id[0, -1] = feed
        sk, sv, logits = decoder(torch.tensor([[feed]], dtype=torch.int32), kv_len, sk, sv, ck, cv, fw, dims)
        kv_len += 1
        if return_logits:
            step_logits.append(logits[0].numpy().copy())
        head = logits
        if use_pen:
            pv = repeat_penalty if generated >= penalty_range else 1.0
            head = apply_penalty(logits, save_id, pv, penalty_range)
        selected = int(argmax_head(head)[0, 0])
        save_id = torch.cat([save_id, torch.tensor([[selected]], dtype=torch.int32)], dim=-1)
        selected_all.append(selected)
        if selected not in stop:
            generated += 1
            tokens.append(selected)
        step += 1
    return dict(tokens=tokens, selected=selected_all,
                step_logits=np.stack(step_logits) if step_logits else None)

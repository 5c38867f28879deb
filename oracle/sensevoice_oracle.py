"""CPU restatement of the SenseVoiceSmall graph the reference exports -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file; the product path
(libb200asr.so behind b200asr.sensevoice) never does.

Follows /root/reference/SenseVoice/Export_SenseVoice.py:
  front end   SENSE_VOICE.__init__ :139-169 (folded Kaldi fbank Conv1d kernel, mel filterbank, LFR index table),
              forward :271-283 (power, log-mel, LFR gather, CMVN + position, prompt rows)
  prompts     :171-206 (language / system embeddings, sinusoid table, both rounded through fp16)
  SANM block  _prepare_sanm_for_export :208-220 (d^-0.25 on q and k, FSMN centre tap + 1, linear_out bias moved
              onto the FSMN conv), sanm_block :227-258, encode :260-269
  CTC         forward :285-296 (argmax, keep id != next id (circular) and id != blank)
Pinned against the reference module itself by oracle/gen_sensevoice_golden.py -> tests/golden/sensevoice_tiny_*.npz.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict
from typing import Dict, List

import numpy as np
import torch
import torchaudio.compliance.kaldi as kaldi

LANGUAGE_PROMPT_TOKEN_IDS = (0, 3, 4, 7, 11, 12, 13)      # Export_SenseVoice.py:38-50 (auto, zh, en, yue, ja, ko, nospeech)
SYSTEM_PROMPT_IDS_EMO = (1, 2, 14)                         # :172 (use_emo=True)


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

    def to_dict(self):
        return asdict(self)


SENSEVOICE_SMALL = SenseVoiceDims()
TINY_TEST = SenseVoiceDims(d_model=128, n_heads=2, ffn=256, n_blocks0=1, n_blocks=2, n_tp_blocks=1, vocab=300)


def n_frames(n_samples: int, d: SenseVoiceDims) -> int:
    return (n_samples - d.win) // d.hop + 1


def n_lfr(frames: int, d: SenseVoiceDims) -> int:
    return (frames + d.lfr_n - 1) // d.lfr_n


# ---------------------------------------------------------------------------------------------
# seeded synthetic checkpoint (no real weights offline): tensors named after what they are
# ---------------------------------------------------------------------------------------------
def make_raw_weights(d: SenseVoiceDims, seed: int) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    raw: Dict[str, torch.Tensor] = {}

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    raw["embed"] = rn(d.n_embed, d.feat, std=0.5)
    raw["cmvn_means"] = rn(d.feat, std=1.0) - 8.0           # log-mel of int16-range audio sits around 8..20
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


# ---------------------------------------------------------------------------------------------
# constants the exporter bakes (:139-206)
# ---------------------------------------------------------------------------------------------
def fbank_kernel(d: SenseVoiceDims) -> torch.Tensor:
    """[2F][win]: Hamming(symmetric) x one-sided DFT over nfft, with per-frame pre-emphasis (replicate boundary)
    and DC removal folded in (:147-160)."""
    F = d.nfft // 2 + 1
    window = torch.hamming_window(d.win, periodic=False, alpha=0.54, beta=0.46, dtype=torch.float32)
    freqs = torch.arange(F, dtype=torch.float32).unsqueeze(1)
    samples = torch.arange(d.win, dtype=torch.float32).unsqueeze(0)
    omega = (2.0 * torch.pi / d.nfft) * freqs * samples
    cos_b = torch.cos(omega) * window
    sin_b = -torch.sin(omega) * window

    def fold(basis):
        shifted = torch.cat([basis[:, 1:], torch.zeros_like(basis[:, :1])], dim=1)
        out = basis - d.pre_emphasis * shifted
        out[:, 0] = out[:, 0] - d.pre_emphasis * basis[:, 0]
        return out - out.mean(dim=1, keepdim=True)

    return torch.cat([fold(cos_b), fold(sin_b)], dim=0).contiguous()


def mel_filters(d: SenseVoiceDims) -> torch.Tensor:
    """[F][n_mels]: Kaldi triangular bank (20 Hz .. Nyquist) padded with a zero Nyquist column (:165-166)."""
    banks, _ = kaldi.get_mel_banks(d.n_mels, d.nfft, float(d.sample_rate), 20.0, 0.0, 100.0, -500.0, 1.0)
    return torch.nn.functional.pad(banks, (0, 1), value=0.0).transpose(0, 1).contiguous()


def position_table(n_pos: int, d: SenseVoiceDims) -> torch.Tensor:
    """Sinusoid table over positions 1..n_pos, depth = feat, rounded through fp16 (:186-195)."""
    feat = d.feat
    positions = torch.arange(1, n_pos + 1, dtype=torch.float32)
    inc = torch.log(torch.tensor([10000.0], dtype=torch.float32)) / (feat / 2 - 1)
    inv = torch.exp(torch.arange(feat / 2, dtype=torch.float32) * (-inc)).reshape(1, -1)
    st = positions.reshape(-1, 1) * inv
    return torch.cat([torch.sin(st), torch.cos(st)], dim=1).half().float()


def fold_weights(raw: Dict[str, torch.Tensor], d: SenseVoiceDims, max_lfr: int) -> Dict[str, torch.Tensor]:
    """Checkpoint tensors -> the tensors the exported graph holds (and the engine takes)."""
    fw: Dict[str, torch.Tensor] = {}
    scale = float(d.d_model) ** 0.5                                   # :361-364 embed and CMVN scale pre-multiplied by sqrt(d)
    embed = raw["embed"] * scale
    fw["fbank_kernel"] = fbank_kernel(d)
    fw["mel_filters"] = mel_filters(d)
    n_prompt = 1 + len(SYSTEM_PROMPT_IDS_EMO)
    pos = position_table(max_lfr + n_prompt, d)
    fw["language_embed"] = embed[list(LANGUAGE_PROMPT_TOKEN_IDS)].half().float() + pos[:1]       # :174,200
    fw["system_embed"] = embed[list(SYSTEM_PROMPT_IDS_EMO)] + pos[1:n_prompt]                     # :173,201
    fw["cmvn_means"] = raw["cmvn_means"].clone()
    fw["cmvn_vars"] = raw["cmvn_vars"] * scale
    fw["speech_position"] = pos[n_prompt:].contiguous()
    f = float(d.head_dim) ** -0.25
    c = (d.fsmn_kernel - 1) // 2
    for i in range(d.total_blocks):
        p = f"blk{i}."
        for k in ("norm1.g", "norm1.b", "norm2.g", "norm2.b", "w1.w", "w1.b", "w2.w", "w2.b", "out.w"):
            fw[p + k] = raw[p + k].clone()
        w = raw[p + "qkv.w"].clone(); b = raw[p + "qkv.b"].clone()
        w[:-d.d_model] *= f; b[:-d.d_model] *= f                         # :213-214 q and k rows only
        fw[p + "qkv.w"], fw[p + "qkv.b"] = w, b
        fs = raw[p + "fsmn.w"].clone(); fs[:, c] += 1.0                  # :215 fsmn(v) + v
        fw[p + "fsmn.w"] = fs
        fw[p + "fsmn.b"] = raw[p + "out.b"].clone()                      # :216-217 linear_out bias rides on the conv
    for n in ("after_norm.g", "after_norm.b", "tp_norm.g", "tp_norm.b", "ctc.w", "ctc.b"):
        fw[n] = raw[n].clone()
    return fw


# ---------------------------------------------------------------------------------------------
# forward
# ---------------------------------------------------------------------------------------------
def log_mel(audio: torch.Tensor, fw, d: SenseVoiceDims) -> torch.Tensor:
    """audio [1,1,N] int16-range floats -> [frames][n_mels]  (:274-277)."""
    F = d.nfft // 2 + 1
    spec = torch.nn.functional.conv1d(audio.float(), fw["fbank_kernel"].unsqueeze(1), stride=d.hop)
    re, im = torch.split(spec * spec, F, dim=1)
    power = (re + im).transpose(1, 2)
    eps = float(torch.finfo(torch.float32).eps)
    return torch.matmul(power, fw["mel_filters"]).clamp(min=eps).log()[0]


def lfr_cmvn(mel: torch.Tensor, fw, d: SenseVoiceDims, language_idx: int) -> torch.Tensor:
    """[frames][n_mels] -> [4 + T_lfr][feat]  (:278-285)."""
    frames = mel.shape[0]
    T = n_lfr(frames, d)
    idx = torch.arange(0, T * d.lfr_n, d.lfr_n).unsqueeze(1) + torch.arange(d.lfr_m) - (d.lfr_m - 1) // 2
    idx = idx.clamp(min=0).clamp(max=frames - 1)
    x = mel[idx].reshape(T, d.feat)
    x = (x + fw["cmvn_means"]) * fw["cmvn_vars"]
    x = x + fw["speech_position"][:T]
    return torch.cat([fw["language_embed"][language_idx:language_idx + 1], fw["system_embed"], x], dim=0)


def _ln(x, g, b, eps):
    return torch.nn.functional.layer_norm(x, (x.shape[-1],), g, b, eps)


def sanm_block(x: torch.Tensor, fw, d: SenseVoiceDims, i: int) -> torch.Tensor:
    p = f"blk{i}."
    T = x.shape[0]
    H, dh, D = d.n_heads, d.head_dim, d.d_model
    qkv = torch.nn.functional.linear(_ln(x, fw[p + "norm1.g"], fw[p + "norm1.b"], d.ln_eps), fw[p + "qkv.w"], fw[p + "qkv.b"])
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    qh = q.reshape(T, H, dh).permute(1, 0, 2); kh = k.reshape(T, H, dh).permute(1, 0, 2); vh = v.reshape(T, H, dh).permute(1, 0, 2)
    ctx = torch.matmul(torch.softmax(torch.matmul(qh, kh.transpose(-2, -1)), dim=-1), vh).permute(1, 0, 2).reshape(T, D)
    c = (d.fsmn_kernel - 1) // 2
    mem = torch.nn.functional.conv1d(v.t().unsqueeze(0), fw[p + "fsmn.w"].unsqueeze(1), fw[p + "fsmn.b"], padding=c,
                                     groups=D)[0].t()
    att = torch.nn.functional.linear(ctx, fw[p + "out.w"]) + mem
    if x.shape[-1] == D:
        att = att + x
    h = torch.relu(torch.nn.functional.linear(_ln(att, fw[p + "norm2.g"], fw[p + "norm2.b"], d.ln_eps), fw[p + "w1.w"], fw[p + "w1.b"]))
    return att + torch.nn.functional.linear(h, fw[p + "w2.w"], fw[p + "w2.b"])


def encode(x: torch.Tensor, fw, d: SenseVoiceDims, stages: Dict[str, torch.Tensor] | None = None) -> torch.Tensor:
    n_main = d.n_blocks0 + d.n_blocks
    for i in range(n_main):
        x = sanm_block(x, fw, d, i)
        if stages is not None and i == 0:
            stages["block0"] = x.clone()
    x = _ln(x, fw["after_norm.g"], fw["after_norm.b"], d.ln_eps)
    if stages is not None:
        stages["after_norm"] = x.clone()
    for i in range(n_main, d.total_blocks):
        x = sanm_block(x, fw, d, i)
    return _ln(x, fw["tp_norm.g"], fw["tp_norm.b"], d.ln_eps)


def ctc_collapse(ids: torch.Tensor, blank_id: int) -> List[int]:
    """Keep frame t when ids[t] != ids[(t + 1) % T] and ids[t] != blank (:289-294: the comparison is with the NEXT
    frame, circularly)."""
    nxt = torch.cat([ids[1:], ids[:1]], dim=0)
    keep = (ids != nxt) & (ids != blank_id)
    return [int(v) for v in ids[keep]]


def transcribe(pcm: np.ndarray, fw, d: SenseVoiceDims, language_idx: int = 0, return_stages: bool = False):
    """pcm: int16 (or int16-range float) samples of one clip -> token ids."""
    audio = torch.as_tensor(np.asarray(pcm), dtype=torch.float32).reshape(1, 1, -1)
    stages: Dict[str, torch.Tensor] = {}
    mel = log_mel(audio, fw, d)
    x = lfr_cmvn(mel, fw, d, language_idx)
    enc = encode(x, fw, d, stages)
    logits = torch.nn.functional.linear(enc, fw["ctc.w"], fw["ctc.b"])
    ids = logits.argmax(dim=-1)
    tokens = ctc_collapse(ids, d.blank_id)
    if return_stages:
        stages.update(mel=mel, feats=x, enc_out=enc, logits=logits, frame_ids=ids)
        return tokens, stages
    return tokens

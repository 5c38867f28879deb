"""CPU restatement of the Paraformer (non-streaming) graph the reference exports -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file.

Follows /root/reference/Paraformer/Non-Streaming/Export_Paraformer.py:
  front end   create_kaldi_stft_kernel :326-343, KaldiFbank :346-364, LFR gather forward :474-480
  folds       PARAFORMER.__init__ :385-465 (d_k^-0.25 on q, k rows; every in-block LayerNorm affine absorbed into the Linear
              that consumes it in float64, rounded once: absorb_layer_norm_affine :245-272, fold_linear_output_scale :222-242;
              FSMN identity folded into the centre tap :305-312; encoder_input_bias = means * vars + position :461-465)
  encoder     forward :483-497 (SANM blocks, after_norm)
  CIF         forward :499-521 (alpha = sigmoid(linear(relu(conv k3))), tail 0.45, float64 prefix sum, fire where the floor
              advances, weighted prefix-sum differences)
  decoder     forward :523-563 (per block FFN with inner LayerNorm -> FSMN over tokens + residual -> cross-attention; one
              FFN-only block; output layer with after_norm folded; zero-fire guard)
Pinned against the reference module by oracle/gen_paraformer_golden.py -> tests/golden/paraformer_tiny_*.npz.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict
from typing import Dict, List

import numpy as np
import torch
import torchaudio.compliance.kaldi as kaldi


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

    def to_dict(self):
        return asdict(self)


PARAFORMER_LARGE = ParaformerDims()
TINY_TEST = ParaformerDims(d_model=128, n_heads=2, ffn=256, n_blocks0=1, n_blocks=2, dec_att_blocks=2, dec_ffn_blocks=1,
                           dec_ffn=256, vocab=300)


def make_raw_weights(d: ParaformerDims, seed: int) -> Dict[str, torch.Tensor]:
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


# ---------------------------------------------------------------------------------------------
def stft_kernel(d: ParaformerDims) -> torch.Tensor:
    """[2F][win]: Hamming window x DFT basis, times (pre-emphasis matrix @ DC-removal matrix)  (:326-343)."""
    win = d.win
    window = torch.hamming_window(win, periodic=False, alpha=0.54, beta=0.46)
    freq = torch.arange(d.nfft // 2 + 1, dtype=torch.float32).unsqueeze(1)
    t = torch.arange(win, dtype=torch.float32).unsqueeze(0)
    omega = (2.0 * torch.pi / d.nfft) * freq * t
    real_b = torch.cos(omega) * window.unsqueeze(0)
    imag_b = -torch.sin(omega) * window.unsqueeze(0)
    dc = torch.eye(win) - torch.full((win, win), 1.0 / win)
    prev = torch.zeros((win, win)); prev[0, 0] = 1.0; prev[1:, :-1] = torch.eye(win - 1)
    pre = torch.eye(win) - float(d.pre_emphasis) * prev
    ft = torch.matmul(pre, dc)
    return torch.cat([torch.matmul(real_b, ft), torch.matmul(imag_b, ft)], dim=0).contiguous()


def mel_bins(d: ParaformerDims) -> torch.Tensor:
    """[n_mels][F] (:352-353)."""
    m, _ = kaldi.get_mel_banks(d.n_mels, d.nfft, d.sample_rate, 20.0, 0.0, 100.0, -500.0, 1.0)
    return torch.nn.functional.pad(m, (0, 1), mode="constant", value=0.0).to(torch.float32)


def sinusoid(n_pos: int, depth: int) -> torch.Tensor:
    pos = torch.arange(1, n_pos + 1, dtype=torch.int32).unsqueeze(0).type(torch.float32)
    inc = torch.log(torch.tensor([10000], dtype=torch.float32)) / (depth / 2 - 1)
    inv = torch.exp(torch.arange(depth / 2).type(torch.float32) * (-inc)).reshape(1, -1)
    st = pos.reshape(1, -1, 1) * inv.reshape(1, 1, -1)
    return torch.cat([torch.sin(st), torch.cos(st)], dim=2)[0]


def _absorb(norm_g, norm_b, w, b, scale=1.0):
    """absorb_layer_norm_affine (:245-272): output scale first, then bias += W @ beta, W *= gamma; float64, rounded once."""
    W = w.to(torch.float64)
    B = b.to(torch.float64) if b is not None else torch.zeros(w.shape[0], dtype=torch.float64)
    s = torch.as_tensor(scale, dtype=torch.float64)
    if s.ndim == 0:
        W = W * s; B = B * s
    else:
        W = W * s.reshape(-1).unsqueeze(1); B = B * s.reshape(-1)
    B = B + torch.matmul(W, norm_b.to(torch.float64))
    W = W * norm_g.to(torch.float64).unsqueeze(0)
    return W.to(torch.float32), B.to(torch.float32)


def _scale_out(w, b, scale):
    W = w.to(torch.float64) * scale.unsqueeze(1)
    B = b.to(torch.float64) * scale
    return W.to(torch.float32), B.to(torch.float32)


def fold_weights(raw: Dict[str, torch.Tensor], d: ParaformerDims, max_lfr: int) -> Dict[str, torch.Tensor]:
    fw: Dict[str, torch.Tensor] = {}
    D = d.d_model
    scale = float(D) ** 0.5
    fw["fbank_kernel"] = stft_kernel(d)
    fw["mel_filters"] = mel_bins(d).t().contiguous()                              # [F][n_mels]
    cv = raw["cmvn_vars"] * scale                                                 # :590 CMVN scale x sqrt(d)
    fw["cmvn_vars"] = cv
    pos = sinusoid(max_lfr, d.feat)
    fw["encoder_input_bias"] = (raw["cmvn_means"].to(torch.float64) * cv.to(torch.float64) + pos.to(torch.float64)).to(torch.float32)
    f = float(d.head_dim ** (-0.25))
    c = d.fsmn_kernel // 2
    for i in range(d.enc_blocks):
        p = f"enc{i}."
        qk = torch.ones(3 * D, dtype=torch.float64); qk[:-D] = f
        fw[p + "qkv.w"], fw[p + "qkv.b"] = _absorb(raw[p + "norm1.g"], raw[p + "norm1.b"], raw[p + "qkv.w"], raw[p + "qkv.b"], qk)
        fw[p + "w1.w"], fw[p + "w1.b"] = _absorb(raw[p + "norm2.g"], raw[p + "norm2.b"], raw[p + "w1.w"], raw[p + "w1.b"])
        fs = raw[p + "fsmn.w"].to(torch.float64); fs[:, c] += 1.0
        fw[p + "fsmn.w"] = fs.to(torch.float32)
        for k in ("out.w", "out.b", "w2.w", "w2.b"):
            fw[p + k] = raw[p + k].clone()
    fw["enc_after_norm.g"], fw["enc_after_norm.b"] = raw["enc_after_norm.g"].clone(), raw["enc_after_norm.b"].clone()
    for k in ("cif.conv.w", "cif.conv.b", "cif.out.w", "cif.out.b"):
        fw[k] = raw[k].clone()
    for i in range(d.dec_att_blocks + d.dec_ffn_blocks):
        p = f"dec{i}."
        fw[p + "w1.w"], fw[p + "w1.b"] = _absorb(raw[p + "norm1.g"], raw[p + "norm1.b"], raw[p + "w1.w"], raw[p + "w1.b"])
        fw[p + "w2.w"], fw[p + "w2.b"] = _absorb(raw[p + "ffn_norm.g"], raw[p + "ffn_norm.b"], raw[p + "w2.w"], None)
        if i < d.dec_att_blocks:
            fw[p + "norm2.g"], fw[p + "norm2.b"] = raw[p + "norm2.g"].clone(), raw[p + "norm2.b"].clone()   # feeds the FSMN conv: stays affine
            fs = raw[p + "fsmn.w"].to(torch.float64); fs[:, c] += 1.0
            fw[p + "fsmn.w"] = fs.to(torch.float32)
            fw[p + "q.w"], fw[p + "q.b"] = _absorb(raw[p + "norm3.g"], raw[p + "norm3.b"], raw[p + "q.w"], raw[p + "q.b"], f)
            kv = torch.ones(2 * D, dtype=torch.float64); kv[:D] = f
            fw[p + "kv.w"], fw[p + "kv.b"] = _scale_out(raw[p + "kv.w"], raw[p + "kv.b"], kv)
            fw[p + "cout.w"], fw[p + "cout.b"] = raw[p + "cout.w"].clone(), raw[p + "cout.b"].clone()
    fw["out.w"], fw["out.b"] = _absorb(raw["dec_after_norm.g"], raw["dec_after_norm.b"], raw["out.w"], raw["out.b"])
    return fw


# ---------------------------------------------------------------------------------------------
def _ln(x, eps, g=None, b=None):
    return torch.nn.functional.layer_norm(x, (x.shape[-1],), g, b, eps)


def log_mel(audio: torch.Tensor, fw, d: ParaformerDims) -> torch.Tensor:
    F = d.nfft // 2 + 1
    st = torch.nn.functional.conv1d(audio.float(), fw["fbank_kernel"].unsqueeze(1), stride=d.hop)
    re, im = torch.split(st * st, F, dim=1)
    mel = torch.matmul(fw["mel_filters"].t().unsqueeze(0), re + im)
    eps = torch.tensor(torch.finfo(torch.float32).eps, dtype=torch.float32)
    return torch.maximum(mel, eps).log().transpose(1, 2)[0]


def encoder(mel: torch.Tensor, fw, d: ParaformerDims) -> torch.Tensor:
    frames = mel.shape[0]
    T = (frames + d.lfr_n - 1) // d.lfr_n
    idx = (torch.arange(0, T * d.lfr_n, d.lfr_n).unsqueeze(1) + torch.arange(d.lfr_m) - (d.lfr_m - 1) // 2).clamp(min=0)
    idx = torch.minimum(idx.reshape(-1), torch.tensor(frames - 1))
    x = mel[idx].reshape(T, d.feat)
    enc = x * fw["cmvn_vars"] + fw["encoder_input_bias"][:T]
    D, H, dh = d.d_model, d.n_heads, d.head_dim
    c = d.fsmn_kernel // 2
    for i in range(d.enc_blocks):
        p = f"enc{i}."
        qkv = torch.nn.functional.linear(_ln(enc, d.ln_eps), fw[p + "qkv.w"], fw[p + "qkv.b"])
        v = qkv[:, 2 * D:]
        q, k, vh = torch.split(qkv.view(-1, 3 * H, dh).transpose(0, 1), H, dim=0)
        ctx = torch.matmul(torch.softmax(torch.matmul(q, k.transpose(1, 2)), dim=-1), vh).transpose(0, 1).reshape(-1, D)
        fsmn = torch.nn.functional.conv1d(v.t().unsqueeze(0), fw[p + "fsmn.w"].unsqueeze(1), None, padding=c, groups=D)[0].t()
        att = torch.nn.functional.linear(ctx, fw[p + "out.w"], fw[p + "out.b"]) + fsmn
        enc = enc + att if enc.shape[-1] == D else att
        enc = enc + torch.nn.functional.linear(torch.relu(torch.nn.functional.linear(_ln(enc, d.ln_eps), fw[p + "w1.w"], fw[p + "w1.b"])),
                                               fw[p + "w2.w"], fw[p + "w2.b"])
    return _ln(enc, d.ln_eps, fw["enc_after_norm.g"], fw["enc_after_norm.b"])


def cif(enc_out: torch.Tensor, fw, d: ParaformerDims):
    """-> (acoustic_embeds [N][D], N, alphas [T])."""
    D = d.d_model
    conv = torch.relu(torch.nn.functional.conv1d(enc_out.t().unsqueeze(0), fw["cif.conv.w"], fw["cif.conv.b"],
                                                 padding=(d.cif_kernel - 1) // 2))[0].t()
    alphas = torch.sigmoid(torch.nn.functional.linear(conv, fw["cif.out.w"], fw["cif.out.b"])).squeeze(-1)
    a = torch.cat([alphas, torch.tensor([d.tail_threshold], dtype=torch.float32)], dim=-1)
    hidden = torch.cat([enc_out, torch.zeros(1, D)], dim=0)
    prefix = torch.cumsum(a, dim=-1, dtype=torch.float64).float()
    fl = torch.floor(prefix)
    dis = torch.cat([torch.zeros(1), fl[:-1]], dim=0)
    fire = torch.nonzero(fl > dis, as_tuple=False).squeeze(1)
    ph = torch.cumsum(a.unsqueeze(-1) * hidden, dim=0)
    frames = ph[fire]
    remains = (prefix - fl)[fire]
    completed = frames - remains.unsqueeze(1) * hidden[fire]
    completed = torch.cat([torch.zeros(1, D), completed], dim=0)
    return completed[1:] - completed[:-1], int(fl[-1].to(torch.int32)), alphas


def decoder(acoustic: torch.Tensor, n_tok: int, memory: torch.Tensor, fw, d: ParaformerDims) -> torch.Tensor:
    D, H, dh = d.d_model, d.n_heads, d.head_dim
    c = d.fsmn_kernel // 2
    safe = max(n_tok, 1)
    dec = torch.cat([acoustic, torch.zeros(1, D)], dim=0)[:safe]
    for i in range(d.dec_att_blocks):
        p = f"dec{i}."
        h = torch.relu(torch.nn.functional.linear(_ln(dec, d.dec_ln_eps), fw[p + "w1.w"], fw[p + "w1.b"]))
        x = torch.nn.functional.linear(_ln(h, d.dec_ln_eps), fw[p + "w2.w"], fw[p + "w2.b"])
        sa_in = _ln(x, d.dec_ln_eps, fw[p + "norm2.g"], fw[p + "norm2.b"])
        fsmn = torch.nn.functional.conv1d(sa_in.t().unsqueeze(0), fw[p + "fsmn.w"].unsqueeze(1), None, padding=c, groups=D)[0].t()
        x = dec + fsmn
        q = torch.nn.functional.linear(_ln(x, d.dec_ln_eps), fw[p + "q.w"], fw[p + "q.b"]).view(-1, H, dh).transpose(0, 1)
        kv = torch.nn.functional.linear(memory, fw[p + "kv.w"], fw[p + "kv.b"])
        k = kv[:, :D].reshape(-1, H, dh).transpose(0, 1)
        v = kv[:, D:].reshape(-1, H, dh).transpose(0, 1)
        co = torch.matmul(torch.softmax(torch.matmul(q, k.transpose(1, 2)), dim=-1), v).transpose(0, 1).reshape(-1, D)
        dec = x + torch.nn.functional.linear(co, fw[p + "cout.w"], fw[p + "cout.b"])
    for i in range(d.dec_att_blocks, d.dec_att_blocks + d.dec_ffn_blocks):
        p = f"dec{i}."
        h = torch.relu(torch.nn.functional.linear(_ln(dec, d.dec_ln_eps), fw[p + "w1.w"], fw[p + "w1.b"]))
        dec = torch.nn.functional.linear(_ln(h, d.dec_ln_eps), fw[p + "w2.w"], fw[p + "w2.b"])
    return torch.nn.functional.linear(_ln(dec, d.dec_ln_eps), fw["out.w"], fw["out.b"])


def transcribe(pcm: np.ndarray, fw, d: ParaformerDims, return_stages: bool = False):
    audio = torch.as_tensor(np.asarray(pcm), dtype=torch.float32).reshape(1, 1, -1)
    mel = log_mel(audio, fw, d)
    enc = encoder(mel, fw, d)
    acoustic, n_tok, alphas = cif(enc, fw, d)
    logits = decoder(acoustic, n_tok, enc, fw, d)
    tokens = [int(t) for t in logits.argmax(dim=-1)[:n_tok]]
    if return_stages:
        return tokens, dict(mel=mel, enc_out=enc, alphas=alphas, acoustic=acoustic, logits=logits, n_tok=n_tok)
    return tokens

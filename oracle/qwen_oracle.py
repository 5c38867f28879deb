"""CPU restatement of the reference's Qwen3-ASR path.  TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference leg may import this file; the product path never does.

What it restates (reference file:line):
  * front end + audio encoder + prompt concat: QWEN3_ASR_ENCODER.forward, Qwen_ASR/Export_Qwen_ASR.py:850-927
    (weight folds `_fuse_encoder_weights` :829-848, chunk bookkeeping :744-760, key mask table :768-775,
    `_get_feat_extract_output_lengths` :519-527); STFT = Qwen_ASR/STFT_Process.py 'stft_B' with reflect centre pad and
    the last frame dropped (:104-170) = the Whisper front end of oracle/whisper_oracle.py (same constants: n_fft 400,
    hop 160, periodic Hann, slaney mel x128), so those functions are reused.
  * rotary table + causal mask: QWEN3_ASR_ROTARY_MASK_PREFILL / _DECODE :933-1025
  * decoder: QWEN3_ASR_DECODER_MAIN._fuse_weights :1141-1190, forward :1265-1336 (RMS norm = SimplifiedLayerNormalization
    :1043-1077; QK-norm with d^-0.25 folded :1156-1163; rotate_half = flip of the two halves with the sign in the sine
    table :1259-1265,968-975; GQA by grouping the query heads :1301-1309; SwiGLU :1327-1329; final norm keeps gamma :1179-1190)
  * heads: ARGMAX :1418-1420, CONCAT_EMBED :1428-1435; host loop: Qwen_ASR/Inference_Qwen_ASR_ONNX.py:656-737
    (generation_limit = max_seq_len - 10 - prompt_len :666; stop on any stop id; first token counted :683-687).

Parity pin: oracle/gen_qwen_golden.py runs the reference's own classes (AST-extracted, seeded tiny checkpoint) and
asserts this file reproduces them before tests/golden/qwen_tiny_case*.npz are written; tests/test_qwen_cpu.py re-checks
the oracle against those files on every run.  The reference holds no golden vectors of its own for this path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from oracle import whisper_oracle as wo

MASK_VALUE = -128.0          # Export_Qwen_ASR.py:775 (encoder key mask), :959-963 (causal mask)
CHUNK = 100                  # mel frames per conv chunk = 2 * n_window (:744); the length formula hard-codes 100 -> 13
CHUNK_TOKENS = 13


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
    chunks_per_window: int = 8          # n_window_infer // (2 * n_window)
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
    def conv_freq(self) -> int:          # mel bins after three stride-2 convs
        return (((self.n_mels + 1) // 2 + 1) // 2 + 1) // 2

    @property
    def tokens_per_window(self) -> int:
        return self.chunks_per_window * CHUNK_TOKENS


QWEN3_ASR_0_6B = QwenDims()
TINY_TEST = QwenDims(enc_layers=2, enc_d=128, enc_heads=2, enc_ffn=256, conv_ch=16, out_dim=128, vocab=512, hidden=128,
                     inter=256, dec_layers=2, heads=4, kv_heads=2, head_dim=64, max_seq_len=256)


@dataclass(frozen=True)
class QwenPrompt:
    """Token ids the exporter bakes around the audio (Export_Qwen_ASR.py:1540-1586)."""
    head_ids: Sequence[int]
    suffix_ids: Sequence[int]
    tail_ids: Sequence[int]
    stop_ids: Sequence[int]


TINY_PROMPT = QwenPrompt(head_ids=(500, 501, 502), suffix_ids=(503, 502, 500, 504, 502, 505),
                         tail_ids=(506, 503, 502, 500, 507, 502, 508, 509), stop_ids=(510, 503))


def aftercnn_len(n: int) -> int:
    """Tokens a chunk of n <= 100 valid mel frames yields (:519-527 restricted to one chunk)."""
    if n >= CHUNK:
        return CHUNK_TOKENS
    if n <= 0:
        return 0
    a = (n - 1) // 2 + 1
    b = (a - 1) // 2 + 1
    return (b - 1) // 2 + 1


def audio_token_count(n_samples: int, d: QwenDims) -> int:
    frames = n_samples // d.hop
    full, rem = divmod(frames, CHUNK)
    return full * CHUNK_TOKENS + aftercnn_len(rem)


def make_raw_weights(d: QwenDims, seed: int) -> Dict[str, torch.Tensor]:
    """Seeded synthetic checkpoint with the Hugging Face tensor names the exporter's skeleton model uses (:311-516)."""
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

    C = d.conv_ch
    a = "thinker.audio_tower."
    raw[a + "conv2d1.weight"] = rn(C, 1, 3, 3, std=1.0 / 3.0); raw[a + "conv2d1.bias"] = rn(C, std=0.05)
    raw[a + "conv2d2.weight"] = rn(C, C, 3, 3, std=(9 * C) ** -0.5); raw[a + "conv2d2.bias"] = rn(C, std=0.05)
    raw[a + "conv2d3.weight"] = rn(C, C, 3, 3, std=(9 * C) ** -0.5); raw[a + "conv2d3.bias"] = rn(C, std=0.05)
    lin(a + "conv_out", d.enc_d, C * d.conv_freq, bias=False)
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


def sinusoid_positions(length: int, channels: int) -> torch.Tensor:
    """SinusoidsPositionEmbedding (:399-405)."""
    inc = np.log(10000.0) / (channels // 2 - 1)
    inv = torch.exp(-inc * torch.arange(channels // 2).float())
    st = torch.arange(length)[:, None] * inv[None, :]
    return torch.cat([torch.sin(st), torch.cos(st)], dim=1)


def rotary_tables(d: QwenDims):
    """cos/sin [max_seq_len][head_dim/2] (:977-984); inv_freq as in the rotary module's default branch (:455-458)."""
    dim = d.head_dim
    inv_freq = 1.0 / (d.rope_theta ** (torch.arange(0, dim, 2, dtype=torch.int64).float() / dim))
    theta = torch.arange(d.max_seq_len, dtype=torch.float32).unsqueeze(-1) * inv_freq
    return torch.cos(theta), torch.sin(theta)


def fold_weights(raw: Dict[str, torch.Tensor], d: QwenDims, audio_int16: bool = False) -> Dict[str, torch.Tensor]:
    """Folded tensors under the names the CUDA engine binds (b200asr/qwen.py:fold_qwen does the same fold for the
    product; this copy exists so the checker stays independent of the package)."""
    fw: Dict[str, torch.Tensor] = {}
    fw["stft_kernel"] = wo.stft_kernel(d.nfft, (1.0 / 32768.0) if audio_int16 else 1.0)
    fw["mel_fbank"] = wo.mel_filterbank(d.nfft // 2 + 1, d.n_mels, d.sample_rate)
    a = "thinker.audio_tower."
    for i in (1, 2, 3):
        fw[f"conv{i}.w"] = raw[f"{a}conv2d{i}.weight"].clone()
        fw[f"conv{i}.b"] = raw[f"{a}conv2d{i}.bias"].clone()
    fw["conv_out.w"] = raw[a + "conv_out.weight"].clone()
    fw["enc_pos"] = sinusoid_positions(d.max_source_positions, d.enc_d)[:CHUNK_TOKENS].clone()
    s = float(d.enc_head_dim) ** -0.25                                   # sqrt(scaling) on q and k (:838-843)
    for i in range(d.enc_layers):
        p = f"{a}layers.{i}."
        W = torch.cat([raw[p + f"self_attn.{n}_proj.weight"] for n in "qkv"], dim=0)
        b = torch.cat([raw[p + f"self_attn.{n}_proj.bias"] for n in "qkv"], dim=0)
        g1, b1 = raw[p + "self_attn_layer_norm.weight"], raw[p + "self_attn_layer_norm.bias"]
        b = b + W @ b1
        W = W * g1.unsqueeze(0)
        W[: 2 * d.enc_d] *= s
        b[: 2 * d.enc_d] *= s
        fw[f"enc{i}.qkv.w"], fw[f"enc{i}.qkv.b"] = W, b
        fw[f"enc{i}.out.w"], fw[f"enc{i}.out.b"] = raw[p + "self_attn.out_proj.weight"].clone(), raw[p + "self_attn.out_proj.bias"].clone()
        g2, b2 = raw[p + "final_layer_norm.weight"], raw[p + "final_layer_norm.bias"]
        fw[f"enc{i}.fc1.w"] = raw[p + "fc1.weight"] * g2.unsqueeze(0)
        fw[f"enc{i}.fc1.b"] = raw[p + "fc1.bias"] + raw[p + "fc1.weight"] @ b2
        fw[f"enc{i}.fc2.w"], fw[f"enc{i}.fc2.b"] = raw[p + "fc2.weight"].clone(), raw[p + "fc2.bias"].clone()
    gp, bp = raw[a + "ln_post.weight"], raw[a + "ln_post.bias"]
    fw["proj1.w"] = raw[a + "proj1.weight"] * gp.unsqueeze(0)
    fw["proj1.b"] = raw[a + "proj1.bias"] + raw[a + "proj1.weight"] @ bp
    fw["proj2.w"], fw["proj2.b"] = raw[a + "proj2.weight"].clone(), raw[a + "proj2.bias"].clone()
    t = "thinker.model."
    fw["embed.w"] = raw[t + "embed_tokens.weight"].clone()
    fw["lm_head.w"] = raw["thinker.lm_head.weight"].clone()
    fw["final_norm.g"] = raw[t + "norm.weight"].clone()
    qs = float(d.head_dim) ** -0.25
    for i in range(d.dec_layers):
        p = f"{t}layers.{i}."
        W = torch.cat([raw[p + f"self_attn.{n}_proj.weight"] for n in "qkv"], dim=0)
        fw[f"dec{i}.qkv.w"] = W * raw[p + "input_layernorm.weight"].unsqueeze(0)
        fw[f"dec{i}.qk_norm.g"] = torch.stack([raw[p + "self_attn.q_norm.weight"] * qs, raw[p + "self_attn.k_norm.weight"] * qs])
        fw[f"dec{i}.o.w"] = raw[p + "self_attn.o_proj.weight"].clone()
        g2 = raw[p + "post_attention_layernorm.weight"].unsqueeze(0)
        fw[f"dec{i}.gate_up.w"] = torch.cat([raw[p + "mlp.gate_proj.weight"] * g2, raw[p + "mlp.up_proj.weight"] * g2], dim=0)
        fw[f"dec{i}.down.w"] = raw[p + "mlp.down_proj.weight"].clone()
    cos, sin = rotary_tables(d)
    fw["rope_cos"], fw["rope_sin"] = cos, sin
    return fw


def _ln(x, eps):
    return F.layer_norm(x, (x.shape[-1],), None, None, eps)


def _rms(x, eps):
    return x * torch.rsqrt(x.pow(2).mean(dim=-1, keepdim=True) + eps)


def features(audio: torch.Tensor, fw, d: QwenDims) -> torch.Tensor:
    """audio [1,1,N] (already scaled unless the int16 scale is folded) -> input_features [n_mels, frames] (:850-859)."""
    power = wo.stft_power(audio, fw["stft_kernel"], d.nfft, d.hop)
    return wo.log_mel(power, fw["mel_fbank"])[0]


def audio_encoder(feat: torch.Tensor, fw, d: QwenDims, stages: Dict[str, torch.Tensor] | None = None) -> torch.Tensor:
    """input_features [n_mels, frames] -> audio_hidden [n_tokens, out_dim] (:860-924)."""
    frames = feat.shape[1]
    n_chunks = (frames + CHUNK - 1) // CHUNK
    padded = F.pad(feat, (0, n_chunks * CHUNK - frames))
    chunks = padded.reshape(d.n_mels, n_chunks, CHUNK).permute(1, 0, 2).unsqueeze(1)          # [chunks,1,mels,100]
    lens = [aftercnn_len(min(max(frames - c * CHUNK, 0), CHUNK)) for c in range(n_chunks)]
    x = chunks
    for i in (1, 2, 3):
        x = F.gelu(F.conv2d(x, fw[f"conv{i}.w"], fw[f"conv{i}.b"], stride=2, padding=1), approximate="tanh")
    if stages is not None:
        stages["conv3"] = x.clone()                                                             # [chunks,C,16,13]
    x = x.permute(0, 3, 1, 2).reshape(n_chunks, CHUNK_TOKENS, -1) @ fw["conv_out.w"].T
    h = x + fw["enc_pos"].unsqueeze(0)
    if stages is not None:
        stages["stem"] = h.clone()
    cpw = d.chunks_per_window
    n_win = (n_chunks + cpw - 1) // cpw
    pad_chunks = n_win * cpw - n_chunks
    h = torch.cat([h, torch.zeros(pad_chunks, CHUNK_TOKENS, d.enc_d)], dim=0).reshape(n_win, d.tokens_per_window, d.enc_d)
    lens_p = lens + [0] * pad_chunks
    valid = [sum(lens_p[w * cpw:(w + 1) * cpw]) for w in range(n_win)]
    tpw = d.tokens_per_window
    mask = torch.zeros(n_win, 1, 1, tpw)
    for w, n in enumerate(valid):
        mask[w, 0, 0, n:] = MASK_VALUE
    H, dh = d.enc_heads, d.enc_head_dim
    for i in range(d.enc_layers):
        qkv = _ln(h, d.enc_ln_eps) @ fw[f"enc{i}.qkv.w"].T + fw[f"enc{i}.qkv.b"]
        q, k, v = [t.reshape(n_win, tpw, H, dh).transpose(1, 2) for t in qkv.split(d.enc_d, dim=-1)]
        att = torch.softmax(q @ k.transpose(-1, -2) + mask, dim=-1) @ v
        h = h + att.transpose(1, 2).reshape(n_win, tpw, d.enc_d) @ fw[f"enc{i}.out.w"].T + fw[f"enc{i}.out.b"]
        y = F.gelu(_ln(h, d.enc_ln_eps) @ fw[f"enc{i}.fc1.w"].T + fw[f"enc{i}.fc1.b"], approximate="tanh")
        h = h + y @ fw[f"enc{i}.fc2.w"].T + fw[f"enc{i}.fc2.b"]
    y = F.gelu(_ln(h, d.enc_ln_eps) @ fw["proj1.w"].T + fw["proj1.b"], approximate="tanh")
    y = (y @ fw["proj2.w"].T + fw["proj2.b"]).reshape(-1, d.out_dim)
    return y[: sum(lens)]


def build_prompt_embed(audio_hidden: torch.Tensor, fw, prompt: QwenPrompt, query_ids: Sequence[int] = (),
                       language_tail_ids: Sequence[int] = ()) -> torch.Tensor:
    """[head | query | suffix | audio | tail | language tail] (:925, CONCAT_EMBED :1428-1435)."""
    E = fw["embed.w"]
    emb = lambda ids: E[torch.tensor(list(ids), dtype=torch.long)] if len(ids) else torch.zeros(0, E.shape[1])
    return torch.cat([emb(prompt.head_ids), emb(query_ids), emb(prompt.suffix_ids), audio_hidden, emb(prompt.tail_ids),
                      emb(language_tail_ids)], dim=0)


def decoder(x: torch.Tensor, history_len: int, kv: List[torch.Tensor], fw, d: QwenDims,
            stages: Dict[str, torch.Tensor] | None = None):
    """x [n, hidden] new-token embeddings at positions history_len.. -> (logits [vocab] of the last row, kv).
    kv[i] = K [kv_heads, len, dh], kv[L+i] = V (same layout; the reference keeps K transposed, :1299-1312)."""
    n = x.shape[0]
    H, KH, dh, L = d.heads, d.kv_heads, d.head_dim, d.dec_layers
    G = H // KH
    half = dh // 2
    cos = fw["rope_cos"][history_len:history_len + n].unsqueeze(1)          # [n,1,half]
    sin = fw["rope_sin"][history_len:history_len + n].unsqueeze(1)
    rows = torch.arange(history_len, history_len + n).unsqueeze(1)
    cols = torch.arange(history_len + n).unsqueeze(0)
    mask = torch.where(cols <= rows, torch.tensor(0.0), torch.tensor(MASK_VALUE))   # [n, kv]
    new_kv = list(kv)
    for i in range(L):
        qkv = _rms(x, d.rms_eps) @ fw[f"dec{i}.qkv.w"].T
        qkv = qkv.reshape(n, H + 2 * KH, dh)
        qk, v = qkv[:, :H + KH], qkv[:, H + KH:]
        g = torch.cat([fw[f"dec{i}.qk_norm.g"][0].expand(H, dh), fw[f"dec{i}.qk_norm.g"][1].expand(KH, dh)], dim=0)
        qk = _rms(qk, d.rms_eps) * g
        x1, x2 = qk[..., :half], qk[..., half:]
        qk = torch.cat([x1 * cos - x2 * sin, x2 * cos + x1 * sin], dim=-1)
        q, k = qk[:, :H], qk[:, H:]
        K = torch.cat([kv[i], k.transpose(0, 1)], dim=1)                    # [KH, len, dh]
        V = torch.cat([kv[L + i], v.transpose(0, 1)], dim=1)
        new_kv[i], new_kv[L + i] = K, V
        qg = q.reshape(n, KH, G, dh).permute(1, 2, 0, 3)                    # [KH, G, n, dh]
        att = torch.softmax(qg @ K.unsqueeze(1).transpose(-1, -2) + mask, dim=-1) @ V.unsqueeze(1)   # [KH,G,n,dh]
        ctx = att.permute(2, 0, 1, 3).reshape(n, H * dh)
        x = x + ctx @ fw[f"dec{i}.o.w"].T
        gu = _rms(x, d.rms_eps) @ fw[f"dec{i}.gate_up.w"].T
        x = x + (F.silu(gu[:, :d.inter]) * gu[:, d.inter:]) @ fw[f"dec{i}.down.w"].T
        if stages is not None:
            stages[f"dec{i}"] = x.clone()
    last = _rms(x[-1], d.rms_eps) * fw["final_norm.g"]
    return fw["lm_head.w"] @ last, new_kv


def empty_kv(d: QwenDims) -> List[torch.Tensor]:
    return [torch.zeros(d.kv_heads, 0, d.head_dim) for _ in range(2 * d.dec_layers)]


def prepare_audio(pcm_int16: np.ndarray) -> torch.Tensor:
    """int16 -> [-1,1] float (prepare_audio_input with audio_pcm_scale 32768, Inference_Qwen_ASR_ONNX.py:592-596)."""
    return torch.from_numpy(pcm_int16.astype(np.float32) / 32768.0).reshape(1, 1, -1)


def apply_penalty(logits: torch.Tensor, save_id: Sequence[int], penalty_value: float, penalty_range: int) -> torch.Tensor:
    """APPLY_PENALTY / PENALIZE_LOGITS (Export_Qwen_ASR.py:669-694,1403-1415): the logits of the last `penalty_range`
    selected ids are multiplied by `penalty_value`; gather-then-scatter, so a repeated id is scaled once."""
    ids = list(save_id)[-penalty_range:]
    out = logits.clone()
    for i in set(ids):
        out[i] = logits[i] * penalty_value
    return out


def greedy_transcribe(pcm_int16: np.ndarray, fw, d: QwenDims, prompt: QwenPrompt, query_ids: Sequence[int] = (),
                      language_tail_ids: Sequence[int] = (), max_new: int | None = None, forced: Sequence[int] | None = None,
                      return_stages: bool = False, repeat_penalty: float = 1.0, penalty_range: int = 10):
    """Greedy host loop of Inference_Qwen_ASR_ONNX.py:656-737.  `forced` feeds the given ids instead of the arg-max
    (teacher forcing for logit comparison; stop test skipped).  repeat_penalty != 1 = the script's penalty_greedy
    strategy (its default, REPEAT_PENALTY = 0.8 / PENALTY_RANGE = 10, :90-91,369-376): the prefill head is a plain
    arg-max, every decode step scales the logits of the last `penalty_range` selected ids first
    (merge_decode_penalty_greedy, Qwen_ASR/Shared_Merged.py:806-840)."""
    with torch.no_grad():
        st: Dict[str, torch.Tensor] = {}
        feat = features(prepare_audio(pcm_int16), fw, d)
        audio_hidden = audio_encoder(feat, fw, d, st)
        emb = build_prompt_embed(audio_hidden, fw, prompt, query_ids, language_tail_ids)
        n_prompt = emb.shape[0]
        limit = max(d.max_seq_len - 10 - n_prompt, 0)
        if max_new is not None:
            limit = min(limit, max_new)
        logits, kv = decoder(emb, 0, empty_kv(d), fw, d, st)
        all_logits = [logits]
        tok = int(torch.argmax(logits))
        tokens: List[int] = []
        count = 0
        stop = set(int(s) for s in prompt.stop_ids)
        kv_len = n_prompt
        if forced is not None:
            for t in forced:
                logits, kv = decoder(fw["embed.w"][int(t)].unsqueeze(0), kv_len, kv, fw, d)
                kv_len += 1
                all_logits.append(logits)
        else:
            save_id = [tok]
            if tok not in stop:
                count = 1
                tokens.append(tok)
            while count < limit and tok not in stop:
                logits, kv = decoder(fw["embed.w"][tok].unsqueeze(0), kv_len, kv, fw, d)
                kv_len += 1
                all_logits.append(logits)
                if repeat_penalty != 1.0:
                    logits = apply_penalty(logits, save_id, repeat_penalty, penalty_range)
                tok = int(torch.argmax(logits))
                save_id.append(tok)
                if tok not in stop:
                    count += 1
                    tokens.append(tok)
        if return_stages:
            st.update(features=feat, audio_hidden=audio_hidden, prompt_embed=emb, logits=torch.stack(all_logits))
            return tokens, st
        return tokens


def sampling_transcribe(pcm_int16: np.ndarray, fw, d: QwenDims, prompt: QwenPrompt, query_ids: Sequence[int], language_tail_ids: Sequence[int],
                        max_new: int, temperature: float, top_k: int, top_p: float, repetition_penalty: float, noise: np.ndarray):
    """The script's `sampling` strategy (Inference_Qwen_ASR_ONNX.py:369-376,640-644,697-737): TOPK_TOPP_SAMPLING
    (Export_Qwen_ASR.py:1348-1400 -- line for line the Whisper head, restated in whisper_oracle.topk_topp_sample and pinned
    there to the reference class) on the prefill logits with an empty history, then on every decode step with all ids
    selected so far.  noise [launch][top_k] replaces the head's rand_like draw."""
    with torch.no_grad():
        feat = features(prepare_audio(pcm_int16), fw, d)
        emb = build_prompt_embed(audio_encoder(feat, fw, d), fw, prompt, query_ids, language_tail_ids)
        limit = min(max(d.max_seq_len - 10 - emb.shape[0], 0), int(max_new))
        stop = set(int(s) for s in prompt.stop_ids)
        logits, kv = decoder(emb, 0, empty_kv(d), fw, d)
        kv_len = emb.shape[0]
        save_id = torch.zeros(1, 0, dtype=torch.int32)
        sampled, save_id = wo.topk_topp_sample(logits.unsqueeze(0), temperature, top_k, top_p, repetition_penalty, save_id, noise[0:1])
        tok = int(sampled[0, 0])
        tokens, count, step = [], 0, 0
        if tok not in stop and limit > 0:
            count = 1
            tokens.append(tok)
        while count < limit and tok not in stop:
            logits, kv = decoder(fw["embed.w"][tok].unsqueeze(0), kv_len, kv, fw, d)
            kv_len += 1
            step += 1
            sampled, save_id = wo.topk_topp_sample(logits.unsqueeze(0), temperature, top_k, top_p, repetition_penalty, save_id,
                                                   noise[step:step + 1])
            tok = int(sampled[0, 0])
            if tok not in stop:
                count += 1
                tokens.append(tok)
        return tokens

"""Synthetic 16 kHz PCM clips for tests and bench.

clip i = int16(clip(round(N(0,1)*1638 + tones), -32768, 32767)), generator
seeded with 1234+i; three sinusoids (220/440/1760 Hz, amplitude 3000) keep the
mel spectrum from being flat.
"""
import numpy as np
import torch


def synth_pcm(index: int, n_samples: int = 128000, sample_rate: int = 16000,
              tones: bool = True) -> np.ndarray:
    g = torch.Generator().manual_seed(1234 + int(index))
    x = torch.randn(n_samples, generator=g, dtype=torch.float32) * 1638.0
    if tones:
        t = torch.arange(n_samples, dtype=torch.float32) / float(sample_rate)
        for k, f in enumerate((220.0, 440.0, 1760.0)):
            x = x + 3000.0 * torch.sin(2.0 * torch.pi * f * t + 0.5 * k + 0.1 * index)
    x = torch.clamp(torch.round(x), -32768, 32767)
    return x.to(torch.int16).numpy()


def synth_batch(batch: int, n_samples: int = 128000, first_index: int = 0) -> np.ndarray:
    return np.stack([synth_pcm(first_index + i, n_samples) for i in range(batch)])


def synth_whisper_checkpoint(dims, seed: int, pos_scale: float = 1.0):
    """Seeded random Whisper checkpoint with HF state-dict key names (no real
    checkpoints exist offline).  Linear ~ N(0, 1/fan_in), LayerNorm gamma ~ 1+0.1N,
    beta ~ 0.1N, token embedding ~ 0.05N; draw order is part of the contract
    (tests pin it against the oracle's generator).  pos_scale multiplies the decoder position table after the draw:
    100 gives non-degenerate greedy streams (20+ distinct ids in 33 steps) while the random net stays in the regime
    where rounding errors do not amplify (oracle/whisper_oracle.py: make_raw_weights)."""
    import math
    g = torch.Generator().manual_seed(int(seed))
    d, f = dims.d_model, dims.ffn
    w = {}

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
        linear(prefix + ".k_proj", d, d, bias=False)
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
        attn(p + "self_attn"); norm(p + "self_attn_layer_norm")
        linear(p + "fc1", f, d); linear(p + "fc2", d, f); norm(p + "final_layer_norm")
    norm(e + "layer_norm")
    dd = "model.decoder."
    w[dd + "embed_tokens.weight"] = rn(dims.vocab, d, std=0.05)
    w[dd + "embed_positions.weight"] = rn(dims.max_target, d, std=0.05) * float(pos_scale)
    for i in range(dims.dec_layers):
        p = f"{dd}layers.{i}."
        attn(p + "self_attn"); norm(p + "self_attn_layer_norm")
        attn(p + "encoder_attn"); norm(p + "encoder_attn_layer_norm")
        linear(p + "fc1", f, d); linear(p + "fc2", d, f); norm(p + "final_layer_norm")
    norm(dd + "layer_norm")
    w["proj_out.weight"] = w[dd + "embed_tokens.weight"]
    return w

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

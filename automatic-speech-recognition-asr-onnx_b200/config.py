"""Model dimensions of the Whisper family as the engine sees them."""
from __future__ import annotations

from dataclasses import dataclass, asdict


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


WHISPER_LARGE_V3 = WhisperDims()
WHISPER_LARGE_V3_TURBO = WhisperDims(dec_layers=4)
WHISPER_TINY_TEST = WhisperDims(d_model=256, n_heads=4, ffn=512, enc_layers=2, dec_layers=2, vocab=1000)

PRESETS = {
    "whisper-large-v3": WHISPER_LARGE_V3,
    "whisper-large-v3-turbo": WHISPER_LARGE_V3_TURBO,
    "whisper-tiny-test": WHISPER_TINY_TEST,
}

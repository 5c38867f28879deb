"""Host-side logic of the Qwen3-ASR ragged path (no GPU): `QwenEngine.pad_ragged`, `transcribe_clips` batching / ordering /
truncation, and the per-clip audio-token bookkeeping the engine lays its prompts out with (Export_Qwen_ASR.py:519-527)."""
import numpy as np
import pytest

from b200asr import qwen as qw


class _FakeEngine:
    """Stands in for QwenEngine above the C ABI: 'tokens' of a clip = [its length, its first sample]."""
    max_batch = 3
    max_samples = 50

    def __init__(self):
        self.calls = []
        self.options = None

    def set_decode_options(self, repeat_penalty, penalty_range):
        self.options = (repeat_penalty, penalty_range)

    def transcribe(self, pcm, query_ids=(), language_tail_ids=(), max_new=-1, lens=None):
        assert lens is not None and pcm.shape[1] == int(max(lens)) and pcm.dtype == np.int16
        self.calls.append(list(map(int, lens)))
        for b, n in enumerate(lens):
            assert not pcm[b, n:].any()                        # padding is zeros
        return [[int(n), int(pcm[b, 0])] for b, n in enumerate(lens)]


def test_pad_ragged_layout():
    clips = [np.arange(1, 6, dtype=np.int16), np.arange(1, 3, dtype=np.int16), np.arange(1, 10, dtype=np.int16)]
    pcm, lens = qw.QwenEngine.pad_ragged(clips)
    assert pcm.shape == (3, 9) and pcm.dtype == np.int16 and lens.tolist() == [5, 2, 9] and lens.dtype == np.int32
    assert pcm[1].tolist() == [1, 2, 0, 0, 0, 0, 0, 0, 0] and pcm[2].tolist() == list(range(1, 10))
    f, fl = qw.QwenEngine.pad_ragged([np.ones(3, np.float32), np.ones((1, 1, 4), np.float32)])
    assert f.dtype == np.float32 and f.shape == (2, 4) and fl.tolist() == [3, 4]


def test_transcribe_clips_batches_by_length_and_keeps_input_order():
    eng = _FakeEngine()
    lens = [10, 40, 12, 60, 39, 11, 8]                        # 60 is cut to the engine's 50 samples
    clips = [np.full(n, i + 1, np.int16) for i, n in enumerate(lens)]
    out = qw.transcribe_clips(eng, clips, repeat_penalty=0.9, penalty_range=7)
    assert eng.options == (0.9, 7)
    assert eng.calls == [[50, 40, 39], [12, 11, 10], [8]]    # neighbours in length share a batch, longest first
    assert [r["tokens"] for r in out] == [[min(n, 50), i + 1] for i, n in enumerate(lens)]
    assert all(r["wall_s"] >= 0 and r["rtf"] >= 0 for r in out)
    eng2 = _FakeEngine()
    qw.transcribe_clips(eng2, clips, max_batch=2)
    assert [len(c) for c in eng2.calls] == [2, 2, 2, 1]
    eng3 = _FakeEngine()
    qw.transcribe_clips(eng3, clips, max_batch=99)            # never above the engine's own capacity
    assert max(len(c) for c in eng3.calls) == 3


@pytest.mark.parametrize("n_samples,want", [(480, 1), (1599, 2), (16000, 13), (31999, 26), (128000, 104), (128160, 105), (480000, 390)])
def test_audio_token_count(n_samples, want):
    """100-frame chunks -> 13 tokens each; the tail chunk through three stride-2 convolutions (ceil halving)."""
    d = qw.QWEN3_ASR_0_6B
    assert d.audio_tokens(n_samples) == want
    frames = n_samples // d.hop
    full, rem = divmod(frames, 100)
    tail = 0
    if rem:
        tail = rem
        for _ in range(3):
            tail = (tail - 1) // 2 + 1
    assert want == full * 13 + tail

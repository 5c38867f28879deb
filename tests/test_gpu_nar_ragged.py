"""Ragged batches for the non-autoregressive models (b200asr_nar_run_ragged): SenseVoice and Paraformer clips of different
lengths in one batch.  The reference graphs take one clip with a dynamic audio axis (SenseVoice/Export_SenseVoice.py:19,379;
Paraformer/Non-Streaming/Export_Paraformer.py:74,603), so the statement to hold is "every clip of the batch gets what it gets
alone": its own fbank frame count, LFR tail (the clamped gather repeats the clip's OWN last frame), FSMN / CIF-conv zero
padding, attention key range, CTC roll and CIF tail position.

fp32: tokens identical to the single-clip run AND to the CPU oracle, encoder rows bit-identical; bf16: same token counts,
encoder rows within 2e-2 (measured on B200: 0.0 -- the bound leaves room for a differently tiled attention launch), written here."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import paraformer_oracle as po, sensevoice_oracle as so
from b200asr import paraformer as pf, sensevoice as sv
from b200asr.engine import WhisperEngine

pytestmark = pytest.mark.gpu
MAX_SAMPLES = 160000
LENS = [40000, 16000, 52345, 9000]          # 248 / 98 / 325 / 54 fbank frames: LFR tails of every residue, shortest = 9 LFR rows


def _clips(seed):
    rng = np.random.default_rng(seed)
    return [(rng.standard_normal(n) * 2500).clip(-32768, 32767).astype(np.int16) for n in LENS]


@pytest.mark.parametrize("precision", ["f32", "bf16"])
def test_sensevoice_ragged_equals_single_and_oracle(precision):
    D = sv.SENSEVOICE_TINY_TEST
    raw = sv.synth_sensevoice_checkpoint(D, 3)
    eng = sv.SenseVoiceEngine(D, sv.fold_sensevoice(raw, D, MAX_SAMPLES), precision=precision, max_batch=4, max_samples=MAX_SAMPLES)
    clips = _clips(11)
    pcm, lens = WhisperEngine.pad_ragged(clips)
    pcm[1, LENS[1]:] = 1234                   # what lies beyond a clip's end must not matter
    langs = [0, 2, 1, 3]
    toks = eng.run(pcm, langs, clip_lens=lens)
    NP = 1 + len(sv.SYSTEM_PROMPT_IDS)
    T = D.lfr_frames(int(lens.max())) + NP
    enc = eng.get_stage("enc_out", 4 * T * D.d_model).reshape(4, T, D.d_model).copy()
    fw = so.fold_weights(so.make_raw_weights(so.TINY_TEST, 3), so.TINY_TEST, D.lfr_frames(MAX_SAMPLES))
    for b, clip in enumerate(clips):
        single = eng.run(clip, langs[b])[0]
        Tb = D.lfr_frames(len(clip)) + NP
        es = eng.get_stage("enc_out", Tb * D.d_model).reshape(Tb, D.d_model)
        de = float(np.abs(enc[b, :Tb] - es).max())
        print(precision, "clip", b, "rows", Tb, "ragged vs single |d enc_out| =", de, "tokens", len(toks[b]), len(single))
        if precision == "f32":
            assert de == 0.0
            assert toks[b] == single
            with torch.no_grad():
                assert toks[b] == so.transcribe(clip, fw, so.TINY_TEST, langs[b])
        else:
            assert de <= 2e-2
            assert abs(len(toks[b]) - len(single)) <= max(2, len(single) // 10)
    # equal lengths through the ragged entry point = the uniform path
    same = np.stack([clips[0], clips[0][::-1].copy()])
    assert eng.run(same, [0, 0], clip_lens=[LENS[0], LENS[0]]) == eng.run(same, [0, 0])
    with pytest.raises(Exception):
        eng.run(pcm, langs, clip_lens=[40000, 16000, 52345, 100])       # shorter than one analysis window
    eng.close()


@pytest.mark.parametrize("precision", ["f32", "bf16"])
@pytest.mark.parametrize("batched", [1, 0])
def test_paraformer_ragged_equals_single_and_oracle(precision, batched):
    D = pf.PARAFORMER_TINY_TEST
    raw = pf.synth_paraformer_checkpoint(D, 5)
    eng = pf.ParaformerEngine(D, pf.fold_paraformer(raw, D, MAX_SAMPLES), precision=precision, max_batch=4, max_samples=MAX_SAMPLES)
    eng.set_option("batched_decoder", batched)
    clips = _clips(12)
    pcm, lens = WhisperEngine.pad_ragged(clips)
    pcm[3, LENS[3]:] = -777
    toks = eng.run(pcm, clip_lens=lens)
    T = D.lfr_frames(int(lens.max()))
    enc = eng.get_stage("enc_out", 4 * T * D.d_model).reshape(4, T, D.d_model).copy()
    fw = po.fold_weights(po.make_raw_weights(po.TINY_TEST, 5), po.TINY_TEST, D.lfr_frames(MAX_SAMPLES))
    for b, clip in enumerate(clips):
        single = eng.run(clip)[0]
        Tb = D.lfr_frames(len(clip))
        es = eng.get_stage("enc_out", Tb * D.d_model).reshape(Tb, D.d_model)
        de = float(np.abs(enc[b, :Tb] - es).max())
        print(precision, "batched" if batched else "per-clip", "clip", b, "rows", Tb, "|d enc_out| =", de, "tokens", len(toks[b]), len(single))
        if precision == "f32":
            assert de == 0.0
            assert toks[b] == single
            with torch.no_grad():
                assert toks[b] == po.transcribe(clip, fw, po.TINY_TEST)
        else:
            assert de <= 2e-2
            assert len(toks[b]) == len(single)
    eng.close()

"""Pins oracle/whisper_oracle.py against vectors minted from the reference's own
nn.Module wrappers (oracle/gen_golden.py).  CPU only."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import whisper_oracle as wo

GOLD = sorted((Path(__file__).parent / "golden").glob("whisper_tiny_case*.npz"))
EOS = 2
NO_SPEECH = 13


@pytest.fixture(scope="module", params=GOLD, ids=[p.stem for p in GOLD])
def case(request):
    g = dict(np.load(request.param))
    dims = wo.TINY_TEST
    raw = wo.make_raw_weights(dims, int(g["seed"]))
    fw = wo.fold_weights(raw, dims, g["suppress"].tolist(), g["begin_suppress"].tolist())
    return g, dims, fw


def test_golden_files_present():
    assert len(GOLD) == 3


def test_frontend_constants(case):
    g, dims, fw = case
    np.testing.assert_allclose(fw["mel_fbank"].numpy()[:, ::8], g["mel_fbank_ref"], atol=1e-7, rtol=1e-6)
    np.testing.assert_allclose(fw["stft_kernel"].numpy()[::16], g["stft_kernel_ref"], atol=1e-7)


def test_encoder_stages(case):
    g, dims, fw = case
    audio = wo.prepare_audio(g["pcm"])
    with torch.no_grad():
        ck, cv, st = wo.encoder(audio, fw, dims, keep_stages=True)
    p = st["power"][0, :, ::4].numpy()
    np.testing.assert_allclose(p, g["power_sub"], rtol=1e-4, atol=1e-6 * float(g["power_sub"].max()))
    np.testing.assert_allclose(st["mel"][0].numpy(), g["mel"], atol=1e-5)
    np.testing.assert_allclose(st["stem"][0].numpy(), g["stem"], atol=2e-5)
    np.testing.assert_allclose(st["enc_layer0"][0].numpy(), g["enc_layer0"], atol=1e-4)
    np.testing.assert_allclose(st["enc_out"][0].numpy(), g["enc_out"], atol=1e-4)
    np.testing.assert_allclose(ck[0].numpy(), g["cross_k_layer0"], atol=1e-4)
    np.testing.assert_allclose(cv[-1].numpy(), g["cross_v_last"], atol=1e-4)
    assert ck[0].shape[-1] == (len(g["pcm"]) // 160 + 1) // 2


def test_free_running_greedy(case):
    g, dims, fw = case
    with torch.no_grad():
        r = wo.greedy_transcribe(g["pcm"], fw, dims, g["prompt"].tolist(), stop_tokens=[], max_new=7)
    np.testing.assert_allclose(r["step_logits"], g["free_logits"], atol=1e-3)   # north_star fp32 tolerance
    assert r["selected"] == g["free_tokens"].tolist()


def test_teacher_forced_logits(case):
    g, dims, fw = case
    with torch.no_grad():
        r = wo.greedy_transcribe(g["pcm"], fw, dims, g["prompt"].tolist(), stop_tokens=[], max_new=7,
                                 forced_tokens=g["forced_tokens"].tolist())
    np.testing.assert_allclose(r["step_logits"], g["forced_logits"], atol=1e-3)


def test_penalty_greedy(case):
    g, dims, fw = case
    with torch.no_grad():
        r = wo.greedy_transcribe(g["pcm"], fw, dims, g["prompt"].tolist(), stop_tokens=[], max_new=7,
                                 repeat_penalty=0.8, penalty_range=3)
    assert r["selected"] == g["penalty_tokens"].tolist()


def test_probe_heads(case):
    g, dims, fw = case
    audio = wo.prepare_audio(g["pcm"])
    with torch.no_grad():
        ck, cv, _ = wo.encoder(audio, fw, dims)
        sk, sv = wo.empty_self_kv(dims)
        _, _, logits = wo.decoder(torch.tensor([[3]], dtype=torch.int32), 0, sk, sv, ck, cv, fw, dims)
        p = wo.no_speech_prob(logits, g["suppress"].tolist(), NO_SPEECH)
    np.testing.assert_allclose(logits[0].numpy(), g["probe_logits"], atol=1e-3)
    np.testing.assert_allclose(p.numpy(), g["no_speech_prob"], rtol=1e-3, atol=1e-7)
    assert wo.detect_language(logits[0].numpy(), g["lang_ids"].tolist()) == int(g["detected_language"])


def test_self_kv_cache_layout(case):
    g, dims, fw = case
    audio = wo.prepare_audio(g["pcm"])
    with torch.no_grad():
        ck, cv, _ = wo.encoder(audio, fw, dims)
        sk, sv = wo.empty_self_kv(dims)
        sk, sv, lg = wo.decoder(torch.tensor([g["prompt"].tolist()], dtype=torch.int32), 0, sk, sv, ck, cv, fw, dims)
        hist = 4
        toks = g["free_tokens"].tolist()
        for t in toks[:6]:
            sk, sv, lg = wo.decoder(torch.tensor([[t]], dtype=torch.int32), hist, sk, sv, ck, cv, fw, dims)
            hist += 1
    np.testing.assert_allclose(sk[-1][0].numpy(), g["self_k_last_layer"], atol=1e-4)
    np.testing.assert_allclose(sv[0][0].numpy(), g["self_v_layer0"], atol=1e-4)

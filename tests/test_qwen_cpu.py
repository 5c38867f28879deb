"""Qwen3-ASR on CPU: the oracle against the golden vectors minted from the reference classes (oracle/gen_qwen_golden.py),
the product-side folds against the oracle's, and the bookkeeping helpers.  No CUDA calls."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import qwen_oracle as qo
from b200asr import qwen as qw

GOLD = sorted((Path(__file__).parent / "golden").glob("qwen_tiny_case*.npz"))


def test_goldens_present():
    assert len(GOLD) == 4


@pytest.mark.parametrize("path", GOLD, ids=[p.stem for p in GOLD])
def test_oracle_reproduces_reference_golden(path):
    g = dict(np.load(path))
    d = qo.TINY_TEST
    fw = qo.fold_weights(qo.make_raw_weights(d, int(g["seed"])), d)
    q, l = g["query_ids"].tolist(), g["language_tail_ids"].tolist()
    toks, st = qo.greedy_transcribe(g["pcm"], fw, d, qo.TINY_PROMPT, q, l, max_new=int(g["max_new"]), return_stages=True)
    assert toks == g["tokens"].tolist()
    np.testing.assert_allclose(st["features"].numpy(), g["features"], atol=1e-5)
    np.testing.assert_allclose(st["audio_hidden"].numpy(), g["audio_hidden"], atol=1e-4)
    assert st["prompt_embed"].shape[0] == int(g["n_prompt"]) and st["audio_hidden"].shape[0] == int(g["n_audio"])
    pen = qo.greedy_transcribe(g["pcm"], fw, d, qo.TINY_PROMPT, q, l, max_new=int(g["penalty_max_new"]), repeat_penalty=0.8, penalty_range=10)
    assert pen == g["penalty_tokens"].tolist()
    _, sf = qo.greedy_transcribe(g["pcm"], fw, d, qo.TINY_PROMPT, q, l, forced=g["forced_tokens"].tolist(), return_stages=True)
    np.testing.assert_allclose(sf["logits"].numpy(), g["forced_logits"], atol=1e-3)


def test_product_checkpoint_and_folds_match_oracle():
    d, pd = qo.TINY_TEST, qw.QWEN_TINY_TEST
    raw_o, raw_p = qo.make_raw_weights(d, 3), qw.synth_qwen_checkpoint(pd, 3)
    assert raw_o.keys() == raw_p.keys()
    for k in raw_o:
        assert torch.equal(raw_o[k], raw_p[k]), k
    fo, fp = qo.fold_weights(raw_o, d), qw.fold_qwen(raw_p, pd)
    for k, v in fp.items():
        ref = fo[k].numpy()
        if k == "conv1.w":
            ref = ref.reshape(pd.conv_ch, 9)
        np.testing.assert_allclose(v, ref, atol=2e-6, err_msg=k)
    assert set(fo) == set(fp)
    assert "lm_head.w" not in qw.fold_qwen(raw_p, pd, tie_lm_head=True)


@pytest.mark.parametrize("n", [400, 1599, 1600, 15999, 16000, 16160, 32000, 130000, 171360, 480000])
def test_audio_token_count(n):
    assert qw.QWEN_TINY_TEST.audio_tokens(n) == qo.audio_token_count(n, qo.TINY_TEST)
    frames = n // 160
    lens = torch.tensor([frames])
    leave = lens % 100                                    # the reference formula (Export_Qwen_ASR.py:519-527), restated
    f1 = torch.clamp(leave - 1, min=0) // 2 + 1
    f1 = f1 * (leave > 0)
    f2 = (torch.clamp(f1 - 1, min=0) // 2 + 1) * (f1 > 0)
    f3 = (torch.clamp(f2 - 1, min=0) // 2 + 1) * (f2 > 0)
    assert qw.QWEN_TINY_TEST.audio_tokens(n) == int(f3 + (lens // 100) * 13)


def test_presets_and_prompt():
    assert qw.QWEN3_ASR_0_6B.enc_head_dim == 64 and qw.QWEN3_ASR_0_6B.conv_freq == 16
    assert qw.QWEN3_ASR_0_6B.out_dim == qw.QWEN3_ASR_0_6B.hidden
    assert qw.QWEN3_ASR_1_7B.out_dim == qw.QWEN3_ASR_1_7B.hidden
    assert tuple(qw.TINY_PROMPT.head_ids) == tuple(qo.TINY_PROMPT.head_ids)
    assert tuple(qw.TINY_PROMPT.tail_ids) == tuple(qo.TINY_PROMPT.tail_ids)
    assert tuple(qw.TINY_PROMPT.stop_ids) == tuple(qo.TINY_PROMPT.stop_ids)

"""CPU tests for the SenseVoice row (a13): the oracle against goldens minted from the reference SENSE_VOICE module
(oracle/gen_sensevoice_golden.py), and the product's weight folds / synthetic checkpoint against the oracle's."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import sensevoice_oracle as so
from b200asr import sensevoice as sv

GOLD = sorted((Path(__file__).parent / "golden").glob("sensevoice_tiny_case*.npz"))


@pytest.mark.parametrize("path", GOLD, ids=[p.stem for p in GOLD])
def test_oracle_matches_reference_golden(path):
    g = dict(np.load(path))
    raw = so.make_raw_weights(so.TINY_TEST, int(g["seed"]))
    fw = so.fold_weights(raw, so.TINY_TEST, int(g["max_lfr"]))
    with torch.no_grad():
        toks, st = so.transcribe(g["pcm"], fw, so.TINY_TEST, int(g["language_idx"]), return_stages=True)
    assert toks == g["tokens"].tolist() and len(toks) == int(g["num"][0])
    np.testing.assert_allclose(st["mel"].numpy(), g["mel"], atol=1e-4)
    np.testing.assert_allclose(st["feats"].numpy(), g["feats"], atol=1e-4)
    np.testing.assert_allclose(st["enc_out"].numpy(), g["enc_out"], atol=1e-3)
    np.testing.assert_allclose(st["logits"].numpy()[:, :64], g["logits_sub"], atol=1e-3)
    assert st["frame_ids"].numpy().tolist() == g["frame_ids"].tolist()


def test_ctc_collapse_is_circular_next_frame_rule():
    ids = torch.tensor([5, 5, 0, 5, 7, 7, 0, 5])
    # keep t when ids[t] != ids[t+1 (circular)] and != blank: t=1 (5!=0), t=3 (5!=7), t=5 (7!=0); t=7: next is ids[0]=5 -> dropped
    assert so.ctc_collapse(ids, 0) == [5, 5, 7]
    assert so.ctc_collapse(torch.tensor([3]), 0) == []           # single frame equals its own circular neighbour
    assert so.ctc_collapse(torch.tensor([0, 0, 0]), 0) == []


def test_product_folds_equal_oracle_folds():
    d, o = sv.SENSEVOICE_TINY_TEST, so.TINY_TEST
    assert d.to_dict() == o.to_dict()
    raw_p = sv.synth_sensevoice_checkpoint(d, 4)
    raw_o = so.make_raw_weights(o, 4)
    assert raw_p.keys() == raw_o.keys() and all(torch.equal(raw_p[k], raw_o[k]) for k in raw_p)
    max_samples = 64000
    fp = sv.fold_sensevoice(raw_p, d, max_samples)
    fo = so.fold_weights(raw_o, o, d.lfr_frames(max_samples))
    assert fp.keys() == fo.keys()
    for k in fp:
        assert np.array_equal(fp[k], fo[k].numpy()), k
    cat = sv.build_supported_languages()
    assert [cat[c]["selector_index"] for c in ("auto", "zh", "en", "yue", "ja", "ko", "nospeech")] == list(range(7))
    assert cat["en"]["prompt_token_ids"] == [4]

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


# ---- Paraformer (a14) -------------------------------------------------------------------------------------------
from oracle import paraformer_oracle as po       # noqa: E402
from b200asr import paraformer as pf             # noqa: E402

PGOLD = sorted((Path(__file__).parent / "golden").glob("paraformer_tiny_case*.npz"))


@pytest.mark.parametrize("path", PGOLD, ids=[p.stem for p in PGOLD])
def test_paraformer_oracle_matches_reference_golden(path):
    g = dict(np.load(path))
    raw = po.make_raw_weights(po.TINY_TEST, int(g["seed"]))
    fw = po.fold_weights(raw, po.TINY_TEST, int(g["max_lfr"]))
    with torch.no_grad():
        toks, st = po.transcribe(g["pcm"], fw, po.TINY_TEST, return_stages=True)
    assert toks == g["tokens"].tolist() and st["n_tok"] == int(g["num"][0])
    np.testing.assert_allclose(st["mel"].numpy(), g["mel"], atol=1e-4)


def test_cif_fires_where_the_float64_prefix_floor_advances():
    d = po.TINY_TEST
    enc = torch.arange(6 * d.d_model, dtype=torch.float32).reshape(6, d.d_model) / 100.0
    fw = {"cif.conv.w": torch.zeros(d.d_model, d.d_model, 3), "cif.conv.b": torch.zeros(d.d_model),
          "cif.out.w": torch.zeros(1, d.d_model), "cif.out.b": torch.tensor([0.0])}      # alpha = 0.5 everywhere
    ac, n, alphas = po.cif(enc, fw, d)
    assert torch.allclose(alphas, torch.full((6,), 0.5))
    assert n == 3 and ac.shape == (3, d.d_model)                 # 6 x 0.5 + tail 0.45 = 3.45
    np.testing.assert_allclose(ac[0].numpy(), (0.5 * enc[0] + 0.5 * enc[1]).numpy(), rtol=1e-6)   # remains = 0 at the fire
    fw["cif.out.b"] = torch.tensor([-20.0])                      # alpha ~ 0: only the tail, nothing fires
    ac, n, _ = po.cif(enc, fw, d)
    assert n == 0 and ac.shape[0] == 0


def test_paraformer_product_folds_equal_oracle_folds():
    d, o = pf.PARAFORMER_TINY_TEST, po.TINY_TEST
    assert d.to_dict() == o.to_dict()
    raw_p, raw_o = pf.synth_paraformer_checkpoint(d, 2), po.make_raw_weights(o, 2)
    assert raw_p.keys() == raw_o.keys() and all(torch.equal(raw_p[k], raw_o[k]) for k in raw_p)
    fp = pf.fold_paraformer(raw_p, d, 64000)
    fo = po.fold_weights(raw_o, o, d.lfr_frames(64000))
    assert fp.keys() == fo.keys()
    for k in fp:
        assert np.array_equal(fp[k], fo[k].numpy()), k
    assert pf.tokens_to_text([0, 1, 2], ["hel@@", "lo", "world"], "en") == "hello world"
    assert pf.tokens_to_text([0, 1], ["你", "好"], "zh") == "你好"


def test_transcribe_long_window_plan_with_stub_engine():
    """Host-side window loop of the SenseVoice / Paraformer scripts (Inference_SenseVoice_ONNX.py:236-260,290-307) on a stub
    engine: window count, stride, zero-padded tail, batching by max_batch, order of concatenation."""
    import numpy as np
    from b200asr import sensevoice as sv

    class Stub:
        max_batch, max_samples = 2, 50000
        def __init__(self):
            self.calls = []
        def run(self, clips, sel):
            self.calls.append((clips.shape, sel))
            return [[int(c[0]), int(c[-1]), int(np.count_nonzero(c))] for c in clips]        # first sample, last sample, non-zero count

    pcm = np.arange(1, 100001, dtype=np.int64).astype(np.int16)          # wraps, but deterministic
    eng = Stub()
    res = sv.transcribe_long(eng, pcm, "Korean", input_audio_length=32000, sliding_window=20000)
    # reference arithmetic: ceil((100000 - 32000) / 20000) + 1 = 5 windows, aligned length 4 * 20000 + 32000 = 112000
    assert res["windows"] == 5 and res["language"] == "ko"
    assert [c[0] for c in eng.calls] == [(2, 32000), (2, 32000), (1, 32000)] and all(c[1] == 5 for c in eng.calls)
    padded = np.zeros(112000, np.int16); padded[:100000] = pcm
    want = []
    for i in range(5):
        w = padded[i * 20000:i * 20000 + 32000]
        want += [int(w[0]), int(w[-1]), int(np.count_nonzero(w))]
    assert res["tokens"] == want and len(res["per_window"]) == 5
    # dynamic axis: one window = the whole clip; a window longer than the engine accepts is refused
    eng = Stub(); eng.max_samples = 200000
    assert sv.transcribe_long(eng, pcm, "auto")["windows"] == 1 and eng.calls[0][0] == (1, 100000)
    import pytest
    with pytest.raises(ValueError, match="max_samples"):
        sv.transcribe_long(Stub(), pcm, "auto")

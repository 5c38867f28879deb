"""Paraformer on the CUDA engine (csrc/sanm.cu, kind = PARAFORMER) against goldens minted from the reference PARAFORMER
module.  fp32: encoder output / alphas / acoustic embeddings / logits within 2e-3, CIF token count and token ids exact.
bf16: token count may only differ when the float64 prefix sum lands within 0.02 of an integer (not the case for the
goldens); ids equal wherever the oracle's top-2 logit margin exceeds 0.5."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import paraformer_oracle as po
from b200asr import paraformer as pf

pytestmark = pytest.mark.gpu
GOLD = sorted((Path(__file__).parent / "golden").glob("paraformer_tiny_case*.npz"))
D = pf.PARAFORMER_TINY_TEST
MAX_SAMPLES = 160000


def _engine(seed, precision, max_batch=1):
    raw = pf.synth_paraformer_checkpoint(D, seed)
    return pf.ParaformerEngine(D, pf.fold_paraformer(raw, D, MAX_SAMPLES), precision=precision, max_batch=max_batch,
                               max_samples=MAX_SAMPLES)


@pytest.mark.parametrize("path", GOLD, ids=[p.stem for p in GOLD])
def test_paraformer_f32_vs_reference_golden(path):
    g = dict(np.load(path))
    eng = _engine(int(g["seed"]), "f32")
    toks = eng.run(g["pcm"])[0]
    T = g["enc_out"].shape[0]
    n = int(g["num"][0])
    np.testing.assert_allclose(eng.get_stage("mel", g["mel"].size).reshape(g["mel"].shape), g["mel"], atol=2e-3)
    np.testing.assert_allclose(eng.get_stage("enc_out", T * D.d_model).reshape(T, D.d_model), g["enc_out"], atol=2e-3)
    np.testing.assert_allclose(eng.get_stage("alphas", T + 1)[:T], g["alphas"], atol=1e-4)
    assert int(eng.get_stage("n_tok", 1)[0]) == n
    ac = eng.get_stage("acoustic", (T + 1) * D.d_model).reshape(T + 1, D.d_model)[:n]
    np.testing.assert_allclose(ac, g["acoustic"], atol=2e-3)
    lg = eng.get_stage("dec_logits", max(n, 1) * D.vocab).reshape(max(n, 1), D.vocab)
    np.testing.assert_allclose(lg[:, :64], g["logits_sub"], atol=3e-3)
    assert toks == g["tokens"].tolist()
    assert eng.run(g["pcm"].astype(np.float32))[0] == g["tokens"].tolist()
    eng.close()


def test_paraformer_bf16_and_batch():
    g = dict(np.load(GOLD[0]))
    raw = po.make_raw_weights(po.TINY_TEST, int(g["seed"]))
    fw = po.fold_weights(raw, po.TINY_TEST, D.lfr_frames(MAX_SAMPLES))
    rng = np.random.default_rng(2)
    clips = (rng.standard_normal((3, 36000)) * 2500).clip(-32768, 32767).astype(np.int16)
    with torch.no_grad():
        want = [po.transcribe(clips[i], fw, po.TINY_TEST, return_stages=True) for i in range(3)]
    eng = _engine(int(g["seed"]), "f32", max_batch=3)
    assert eng.run(clips) == [w[0] for w in want]                 # ragged token counts across the batch
    eng.close()
    eng = _engine(int(g["seed"]), "bf16", max_batch=3)
    got = eng.run(clips)
    for i in range(3):
        toks, st = want[i]
        assert len(got[i]) == len(toks)
        lg = st["logits"].numpy()[:len(toks)]
        top2 = np.sort(lg, axis=-1)[:, -2:]
        safe = (top2[:, 1] - top2[:, 0]) > 0.5
        assert np.array_equal(np.asarray(got[i])[safe], np.asarray(toks)[safe])
    eng.close()


@pytest.mark.parametrize("precision", ["f32", "bf16"])
def test_paraformer_batched_decoder_equals_per_clip(precision):
    """The stacked-row decoder (one pass over all clips, segment-aware FSMN / cross-attention) against the per-clip
    decoder of the same engine: fp32 same tokens and equal to the oracle clip by clip (ragged counts, silence = zero fires);
    bf16 same token counts and at most 10 % of the ids differ (written here; random-weight logits have near-ties)."""
    g = dict(np.load(GOLD[1]))
    rng = np.random.default_rng(7)
    clips = (rng.standard_normal((4, 52000)) * 2500).clip(-32768, 32767).astype(np.int16)
    clips[2] = 0                                             # a silent clip: CIF may fire nothing -> the zero-fire guard row
    clips[3, 20000:] //= 8
    out = {}
    for batched in (1, 0):
        eng = _engine(int(g["seed"]), precision, max_batch=4)
        eng.set_option("batched_decoder", batched)
        out[batched] = (eng.run(clips), eng.kernel_launches)
        eng.close()
    if precision == "f32":
        assert out[1][0] == out[0][0]
    else:       # the stacked path keeps the attention probabilities in fp32 (the per-clip path rounds them to bf16): near-ties may flip
        assert [len(t) for t in out[1][0]] == [len(t) for t in out[0][0]]
        a, b = np.concatenate([np.asarray(t) for t in out[1][0]]), np.concatenate([np.asarray(t) for t in out[0][0]])
        assert (a != b).mean() <= 0.1
    assert out[1][1] < out[0][1]
    print(precision, "tokens per clip", [len(t) for t in out[1][0]], "launches", out[1][1], "vs", out[0][1])
    if precision == "f32":
        fw = po.fold_weights(po.make_raw_weights(po.TINY_TEST, int(g["seed"])), po.TINY_TEST, D.lfr_frames(MAX_SAMPLES))
        with torch.no_grad():
            assert out[1][0] == [po.transcribe(clips[i], fw, po.TINY_TEST) for i in range(4)]


def test_paraformer_sliding_windows_vs_oracle():
    from b200asr import sensevoice as sv
    g = dict(np.load(GOLD[0]))
    rng = np.random.default_rng(4)
    pcm = (rng.standard_normal(90000) * 2500).clip(-32768, 32767).astype(np.int16)
    eng = _engine(int(g["seed"]), "f32", max_batch=2)
    res = sv.transcribe_long(eng, pcm, input_audio_length=40000, sliding_window=0)
    fw = po.fold_weights(po.make_raw_weights(po.TINY_TEST, int(g["seed"])), po.TINY_TEST, D.lfr_frames(MAX_SAMPLES))
    padded = np.zeros(120000, np.int16)
    padded[:90000] = pcm
    want = []
    with torch.no_grad():
        for i in range(3):
            want.extend(po.transcribe(padded[i * 40000:(i + 1) * 40000], fw, po.TINY_TEST))
    assert res["windows"] == 3 and res["tokens"] == want
    eng.close()


def test_paraformer_cli_end_to_end_from_funasr_folder_and_wav(tmp_path, capsys):
    """`python -m b200asr.cli paraformer --model-folder F --audio x.wav`: FunASR-style folder + vocabulary + WAV in, the
    script's `ASR Result` block out (zh mode: vocabulary pieces joined); the ids behind it must be the oracle's."""
    import json, wave
    from funasr_folders import write_paraformer_folder
    from b200asr import cli, paraformer as pfm
    g = dict(np.load(GOLD[0]))
    seed = int(g["seed"])
    write_paraformer_folder(tmp_path, D, pfm.synth_paraformer_checkpoint(D, seed))
    vocab = [f"<{i}>" for i in range(D.vocab)]
    (tmp_path / "tokens.json").write_text(json.dumps(vocab))
    with wave.open(str(tmp_path / "clip.wav"), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000); w.writeframes(g["pcm"].astype("<i2").tobytes())
    rc = cli.main(["paraformer", "--model-folder", str(tmp_path), "--audio", str(tmp_path / "clip.wav"), "--precision", "f32"])
    out = capsys.readouterr().out
    assert rc == 0 and "ASR Result:" in out and "RTF:" in out
    text = out.split("ASR Result:\n")[1].split("\n")[0]
    assert text == "".join(vocab[t] for t in g["tokens"].tolist())

"""SenseVoice on the CUDA engine (csrc/sanm.cu) against goldens minted from the reference SENSE_VOICE module.
fp32 mode: stage tensors to 2e-3 (log-mel of int16-range audio is O(10), logits O(5)), frame ids and tokens exact.
bf16 mode (tcgen05 GEMMs + fused attention): encoder output to 0.12 on O(1) LayerNorm outputs, frame ids equal
wherever the golden top-2 logit margin exceeds 0.5, written here."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import sensevoice_oracle as so
from b200asr import sensevoice as sv

pytestmark = pytest.mark.gpu
GOLD = sorted((Path(__file__).parent / "golden").glob("sensevoice_tiny_case*.npz"))
D = sv.SENSEVOICE_TINY_TEST
MAX_SAMPLES = 160000


def _engine(seed, precision, max_batch=1):
    raw = sv.synth_sensevoice_checkpoint(D, seed)
    return sv.SenseVoiceEngine(D, sv.fold_sensevoice(raw, D, MAX_SAMPLES), precision=precision, max_batch=max_batch,
                               max_samples=MAX_SAMPLES)


@pytest.mark.parametrize("path", GOLD, ids=[p.stem for p in GOLD])
def test_sensevoice_f32_vs_reference_golden(path):
    g = dict(np.load(path))
    eng = _engine(int(g["seed"]), "f32")
    toks = eng.run(g["pcm"], int(g["language_idx"]))[0]
    T = g["feats"].shape[0]
    mel = eng.get_stage("mel", g["mel"].size).reshape(g["mel"].shape)
    np.testing.assert_allclose(mel, g["mel"], atol=2e-3)
    feats = eng.get_stage("feats", g["feats"].size).reshape(g["feats"].shape)
    np.testing.assert_allclose(feats, g["feats"], atol=2e-3)
    enc = eng.get_stage("enc_out", T * D.d_model).reshape(T, D.d_model)
    np.testing.assert_allclose(enc, g["enc_out"], atol=2e-3)
    logits = eng.get_stage("logits", T * D.vocab).reshape(T, D.vocab)
    np.testing.assert_allclose(logits[:, :64], g["logits_sub"], atol=2e-3)
    ids = eng.get_stage("frame_ids", T).astype(np.int32)
    assert ids.tolist() == g["frame_ids"].tolist()
    assert toks == g["tokens"].tolist()
    # float32 input carrying int16-range values (the reference's F32 audio mode, audio_pcm_scale = 1)
    assert eng.run(g["pcm"].astype(np.float32), int(g["language_idx"]))[0] == g["tokens"].tolist()
    eng.close()


@pytest.mark.parametrize("path", GOLD[:2], ids=[p.stem for p in GOLD[:2]])
def test_sensevoice_bf16_vs_reference_golden(path):
    g = dict(np.load(path))
    eng = _engine(int(g["seed"]), "bf16")
    eng.run(g["pcm"], int(g["language_idx"]))
    T = g["feats"].shape[0]
    enc = eng.get_stage("enc_out", T * D.d_model).reshape(T, D.d_model)
    d = float(np.abs(enc - g["enc_out"]).max())
    print("bf16 enc_out max|d| =", d)
    assert d <= 0.12
    raw = so.make_raw_weights(so.TINY_TEST, int(g["seed"]))
    fw = so.fold_weights(raw, so.TINY_TEST, int(g["max_lfr"]))
    with torch.no_grad():
        _, st = so.transcribe(g["pcm"], fw, so.TINY_TEST, int(g["language_idx"]), return_stages=True)
    ref = st["logits"].numpy()
    top2 = np.sort(ref, axis=-1)[:, -2:]
    safe = (top2[:, 1] - top2[:, 0]) > 0.5
    ids = eng.get_stage("frame_ids", T).astype(np.int32)
    assert safe.sum() > 0 and np.array_equal(ids[safe], ref.argmax(-1)[safe])
    eng.close()


def test_sensevoice_batch_and_language_selector():
    g = dict(np.load(GOLD[0]))
    eng = _engine(int(g["seed"]), "f32", max_batch=3)
    raw = so.make_raw_weights(so.TINY_TEST, int(g["seed"]))
    fw = so.fold_weights(raw, so.TINY_TEST, D.lfr_frames(MAX_SAMPLES))
    rng = np.random.default_rng(1)
    clips = (rng.standard_normal((3, 40000)) * 2500).clip(-32768, 32767).astype(np.int16)
    langs = [0, 2, 6]
    got = eng.run(clips, langs)
    with torch.no_grad():
        want = [so.transcribe(clips[i], fw, so.TINY_TEST, langs[i]) for i in range(3)]
    assert got == want
    r = sv.transcribe_clip(eng, clips[1], "English")
    assert r["language"] == "en" and r["tokens"] == want[1] and r["rtf"] > 0
    with pytest.raises(Exception, match="language_idx"):
        eng.run(clips[:1], 9)
    eng.close()


@pytest.mark.parametrize("n_samples", [40000, 83000, 123520, 240000, 255000, 330000, 480000])
def test_sensevoice_bf16_head128_fused_attention(n_samples):
    """Production head width (128 = two swizzle tiles per operand in attention_tc.cu) on a two-head model: the fused
    tcgen05 attention against the three-launch CUDA-core path of the same engine (3e-2 on O(1) LayerNorm outputs, the
    two differ in where P is normalised) and against the fp32 oracle (0.12, the bf16 bound of this file).  The clip
    lengths put T below one 128-key box, at a non-multiple of 16, across the box edge, near the 256-key limit of the
    single-pass kernel, and beyond it (two-pass streaming kernel: 269, 347 and 504 positions)."""
    import dataclasses
    maxs = max(245000, n_samples)       # beyond 256 positions the 128-wide heads take the two-pass streaming kernel (30 s = 504)
    dims = dataclasses.replace(D, d_model=256, n_heads=2, ffn=512)
    odims = dataclasses.replace(so.TINY_TEST, d_model=256, n_heads=2, ffn=512)
    assert dims.head_dim == 128
    raw = sv.synth_sensevoice_checkpoint(dims, 5)
    tensors = sv.fold_sensevoice(raw, dims, maxs)
    rng = np.random.default_rng(n_samples)
    clips = (rng.standard_normal((2, n_samples)) * 2500).clip(-32768, 32767).astype(np.int16)
    T = dims.lfr_frames(n_samples) + 4
    outs = []
    for fused in (1, 0):
        eng = sv.SenseVoiceEngine(dims, tensors, precision="bf16", max_batch=2, max_samples=maxs)
        eng.set_option("attn_tc", fused)
        eng.set_option("graph", 0)
        eng.run(clips, [0, 3])
        enc = eng.get_stage("enc_out", 2 * T * 256).reshape(2, T, 256)
        outs.append((enc.copy(), eng.kernel_launches))
        eng.close()
    d = float(np.abs(outs[0][0] - outs[1][0]).max())
    print(f"T={T}: fused vs unfused enc_out max|d| = {d:.4f}; launches {outs[0][1]} vs {outs[1][1]}")
    assert np.isfinite(outs[0][0]).all()
    assert d <= 3e-2
    assert outs[0][1] < outs[1][1]
    fw = so.fold_weights(so.make_raw_weights(odims, 5), odims, dims.lfr_frames(maxs))
    with torch.no_grad():
        _, st = so.transcribe(clips[1], fw, odims, 3, return_stages=True)
    ref = st["enc_out"].numpy()
    do = float(np.abs(outs[0][0][1] - ref).max())
    print("vs fp32 oracle max|d| =", do)
    assert do <= 0.12


def test_sensevoice_pdl_gemm_launch_is_transparent():
    g = dict(np.load(GOLD[1]))
    outs = []
    for pdl, graph in ((1, 1), (0, 1), (1, 0)):
        eng = _engine(int(g["seed"]), "bf16")
        eng.set_option("pdl", pdl)
        eng.set_option("graph", graph)
        toks = eng.run(g["pcm"], int(g["language_idx"]))[0]
        T = g["feats"].shape[0]
        outs.append((eng.get_stage("logits", T * D.vocab).copy(), toks))
        eng.close()
    for o in outs[1:]:
        assert np.array_equal(o[0], outs[0][0]) and o[1] == outs[0][1]


@pytest.mark.parametrize("window,stride", [(32000, 0), (32000, 20000), (None, 0)])
def test_sensevoice_sliding_windows_vs_oracle(window, stride):
    """The script's window loop (static-length export): windows batched through the engine == the oracle run window by window
    on the zero-padded clip, ids concatenated in order."""
    g = dict(np.load(GOLD[0]))
    rng = np.random.default_rng(9)
    pcm = (rng.standard_normal(100000) * 2500).clip(-32768, 32767).astype(np.int16)
    eng = _engine(int(g["seed"]), "f32", max_batch=3)
    res = sv.transcribe_long(eng, pcm, "English", input_audio_length=window, sliding_window=stride)
    fw = so.fold_weights(so.make_raw_weights(so.TINY_TEST, int(g["seed"])), so.TINY_TEST, D.lfr_frames(MAX_SAMPLES))
    win = len(pcm) if window is None else window
    st = win if stride <= 0 else stride
    n_win = 1 if len(pcm) <= win else int(np.ceil((len(pcm) - win) / st)) + 1
    padded = np.zeros((n_win - 1) * st + win, np.int16)
    padded[:len(pcm)] = pcm
    want = []
    with torch.no_grad():
        for i in range(n_win):
            want.extend(so.transcribe(padded[i * st:i * st + win], fw, so.TINY_TEST, 2))
    assert res["windows"] == n_win and res["language"] == "en"
    assert res["tokens"] == want
    eng.close()


def test_sensevoice_cli_end_to_end_from_funasr_folder_and_wav(tmp_path, capsys):
    """`python -m b200asr.cli sensevoice --model-folder F --audio x.wav`: FunASR-style folder (model.pt, am.mvn, config.yaml) and
    a 16-bit WAV in, the script's `ASR Result` block out; without the SentencePiece model the ids are printed -- they must be
    the oracle's."""
    import wave
    from b200asr import cli
    g = dict(np.load(GOLD[0]))
    seed = int(g["seed"])
    raw = sv.synth_sensevoice_checkpoint(D, seed)
    from funasr_folders import write_sensevoice_folder
    write_sensevoice_folder(tmp_path, D, raw)
    with wave.open(str(tmp_path / "clip.wav"), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000); w.writeframes(g["pcm"].astype("<i2").tobytes())
    rc = cli.main(["sensevoice", "--model-folder", str(tmp_path), "--audio", str(tmp_path / "clip.wav"), "--precision", "f32",
                   "--language", "auto"])
    out = capsys.readouterr().out
    assert rc == 0 and "ASR Result:" in out and "RTF:" in out
    ids = [int(x) for x in out.split("ASR Result:\n")[1].split("\n")[0].split()]
    fw = so.fold_weights(so.make_raw_weights(so.TINY_TEST, seed), so.TINY_TEST, D.lfr_frames(len(g["pcm"])))
    with torch.no_grad():
        want = so.transcribe(g["pcm"], fw, so.TINY_TEST, 0)
    assert ids == want


def test_nar_session_shim_drives_like_the_script():
    """ORT-shaped session over the engine: the names, shapes and call sequence of Inference_SenseVoice_ONNX.py:196-305."""
    from b200asr.session import NarSessions, OrtValue
    g = dict(np.load(GOLD[0]))
    eng = _engine(int(g["seed"]), "f32")
    sess = NarSessions(eng, {"sample_rate": "16000"}).session
    assert [m.name for m in sess.get_inputs()] == ["audio", "language_idx"]
    assert [m.name for m in sess.get_outputs()] == ["token_ids", "num_id"]
    assert sess.get_inputs()[0].shape == [1, 1, "audio_len"] and sess.get_inputs()[0].type == "tensor(int16)"
    assert sess.get_modelmeta().custom_metadata_map["sample_rate"] == "16000"
    binding = sess.io_binding()
    binding.bind_cpu_input("audio", g["pcm"].reshape(1, 1, -1))
    binding.bind_ortvalue_input("language_idx", OrtValue.ortvalue_from_numpy(np.array([int(g["language_idx"])], np.int32)))
    binding._iobinding.bind_output("token_ids", None)
    sess.run_with_iobinding(binding)
    assert binding.get_outputs()[0].numpy().tolist() == g["tokens"].tolist()
    tok, num = sess.run(None, {"audio": g["pcm"].reshape(1, 1, -1), "language_idx": np.array([int(g["language_idx"])], np.int32)})
    assert tok.tolist() == g["tokens"].tolist() and num.tolist() == [len(g["tokens"])]
    with pytest.raises(ValueError, match="no input named"):
        binding.bind_cpu_input("audio_in", g["pcm"])
    with pytest.raises(ValueError, match=r"\[1, 1, audio_len\]"):
        sess.run(None, {"audio": g["pcm"], "language_idx": np.array([0], np.int32)})
    eng.close()

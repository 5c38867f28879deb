"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol the
header declares, the product's weight folds equal the oracle's, the blob format round-trips."""
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import whisper_oracle as wo
from b200asr import _cabi
from b200asr.config import WHISPER_TINY_TEST, PRESETS
from b200asr.synth import synth_pcm, synth_whisper_checkpoint
from b200asr import weights as W

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "b200asr.h").read_text()
    declared = set(re.findall(r"\b(b200asr_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = _cabi.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/b200asr.h but not exported"
    assert declared == set(_cabi.SYMBOLS), declared ^ set(_cabi.SYMBOLS)


def test_engine_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from b200asr.engine import WhisperEngine, B200AsrError
    with pytest.raises(B200AsrError, match="no CPU fallback"):
        WhisperEngine(WHISPER_TINY_TEST, {})


def test_synth_checkpoint_matches_oracle_generator():
    a = synth_whisper_checkpoint(WHISPER_TINY_TEST, 3)
    b = wo.make_raw_weights(wo.TINY_TEST, 3)
    assert a.keys() == b.keys()
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_fold_matches_oracle_fold():
    raw = wo.make_raw_weights(wo.TINY_TEST, 1)
    sup, beg = [1, 5, 900], [220, 2]
    mine = W.fold_whisper(raw, WHISPER_TINY_TEST, sup, beg)
    ref = wo.fold_weights(raw, wo.TINY_TEST, sup, beg)
    for k, v in ref.items():
        r = v.numpy()
        if k.startswith("enc.conv") and k.endswith("weight"):
            name = k.replace(".weight", ".w")
            r = np.ascontiguousarray(r.transpose(0, 2, 1).reshape(r.shape[0], -1))     # [out, k*Cin + c]
        elif k.startswith("enc.conv") and k.endswith("bias"):
            name = k.replace(".bias", ".b")
        else:
            name = k
        assert name in mine, name
        np.testing.assert_allclose(mine[name], r, rtol=1e-6, atol=1e-7, err_msg=name)
    assert set(mine) == {k.replace(".weight", ".w").replace(".bias", ".b") if k.startswith("enc.conv") else k for k in ref}


def test_blob_roundtrip(tmp_path):
    raw = synth_whisper_checkpoint(WHISPER_TINY_TEST, 0)
    t = W.fold_whisper(raw, WHISPER_TINY_TEST, [1], [2])
    path = tmp_path / "w.b200asr"
    W.save_blob(path, t, WHISPER_TINY_TEST, {"sample_rate": 16000})
    dims, t2, meta = W.load_blob(path)
    assert dims == WHISPER_TINY_TEST and meta["sample_rate"] == 16000
    assert set(t2) == set(t)
    for k in t:
        assert np.array_equal(np.asarray(t2[k]), t[k], equal_nan=True), k


def test_synth_pcm_is_deterministic_int16():
    a, b = synth_pcm(3, 16000), synth_pcm(3, 16000)
    assert a.dtype == np.int16 and np.array_equal(a, b) and a.std() > 1000
    assert "whisper-large-v3" in PRESETS


def test_header_is_a_plain_c_header():
    """include/b200asr.h is the C ABI a cgo / JNI / ctypes binding includes: it must compile as C99 and as C++ on its own."""
    import shutil
    import subprocess
    hdr = str(ROOT / "include" / "b200asr.h")
    if shutil.which("gcc") is None:
        pytest.skip("gcc not on PATH")
    subprocess.run(["gcc", "-x", "c", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", hdr], check=True)
    subprocess.run(["g++", "-x", "c++", "-std=c++17", "-fsyntax-only", hdr], check=True)


def test_engines_of_every_family_fail_loudly_without_gpu():
    """No CPU fallback anywhere: creating a SenseVoice / Paraformer / Qwen3-ASR engine without a B200 raises."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from b200asr import paraformer as pfm, qwen as qw, sensevoice as sv
    from b200asr.engine import B200AsrError
    with pytest.raises(B200AsrError):
        sv.SenseVoiceEngine(sv.SENSEVOICE_TINY_TEST, {}, max_samples=32000)
    with pytest.raises(B200AsrError):
        pfm.ParaformerEngine(pfm.PARAFORMER_TINY_TEST, {}, max_samples=32000)
    with pytest.raises(B200AsrError):
        qw.QwenEngine(qw.QWEN_TINY_TEST, {}, qw.TINY_PROMPT, max_samples=32000)

"""CPU tests of the host side of the Whisper drop-in: ORT_IO helpers (a16), driver helpers (a12) and the sampling
head of the oracle (a11), each against golden vectors minted from the reference's own code by
oracle/gen_heads_golden.py (and, when /root/reference is present, against the reference module directly)."""
import json
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from b200asr import ort_io
from b200asr.whisper_infer import plan_windows, prepare_audio_input, remove_repeated_parts
from oracle import whisper_oracle as wo

GOLD = Path(__file__).parent / "golden"
REF = Path("/root/reference")


def meta(name, shape, typ):
    return SimpleNamespace(name=name, shape=shape, type=typ)


def test_sampling_head_oracle_vs_reference_golden():
    g = np.load(GOLD / "whisper_heads.npz")
    for i in range(int(g["n"])):
        sid, save = wo.topk_topp_sample(torch.from_numpy(g[f"logits_{i}"]), float(g[f"t_{i}"]), int(g[f"k_{i}"]),
                                        float(g[f"p_{i}"]), float(g[f"rp_{i}"]), torch.from_numpy(g[f"prev_{i}"]),
                                        g[f"noise_{i}"])
        assert int(sid[0, 0]) == int(g[f"sampled_{i}"]), i
        assert save.numpy().tolist() == g[f"save_{i}"].tolist()


def test_driver_helpers_vs_reference_golden():
    host = json.loads((GOLD / "host_logic.json").read_text())
    pcm = np.load(GOLD / "host_pcm.npy").reshape(1, 1, -1)
    for c in host["prepare"]:
        out = prepare_audio_input(pcm, np.dtype(c["dtype"]), audio_pcm_scale=32768, use_normalise_audio=c["normalise"])
        assert out.dtype == np.dtype(c["dtype"]) and out.flags.c_contiguous
        assert np.isclose(float(np.abs(out.astype(np.float64)).sum()), c["checksum"], rtol=1e-6)
        assert np.allclose(out.reshape(-1)[:8], c["head"], rtol=1e-6)
    for c in host["repeat"]:
        arr = np.asarray(c["ids"])
        assert [int(x) for x in remove_repeated_parts(arr, c["thr"], arr.shape[-1])] == c["out"]
    for c in host["windows"]:
        assert plan_windows(c["audio_len"], c["input_len"], c["sliding"]) == (c["windows"], c["stride"], c["aligned"])


def test_ort_io_shapes_and_arrays():
    m = meta("x", [1, "n", 4], "tensor(float)")
    assert ort_io.numpy_dtype(m) == np.float32 and ort_io.numpy_dtype("tensor(int64)") == np.int64
    with pytest.raises(KeyError):
        ort_io.numpy_dtype("tensor(complex64)")
    assert ort_io.resolve_shape(m, axes={1: 7}) == (1, 7, 4)
    assert ort_io.resolve_shape(m, symbols={"n": 3}) == (1, 3, 4)
    with pytest.raises(ValueError):
        ort_io.resolve_shape(m)
    a = ort_io.array_for(m, np.arange(8, dtype=np.int64).reshape(1, 2, 4))
    assert a.dtype == np.float32 and a.shape == (1, 2, 4) and a.flags.c_contiguous
    kv = meta("in_de_key_layer_0", ["batch", 4, 64, "history_len"], "tensor(float)")
    assert ort_io.filled_for(kv, axes={0: 1, 3: 0}).shape == (1, 4, 64, 0)
    with pytest.raises(ValueError, match="provide axes"):
        ort_io.array_for(meta("s", ["a", "b"], "tensor(int32)"), [1, 2, 3])
    assert ort_io.array_for(meta("s", ["a", "b"], "tensor(int32)"), [1, 2, 3], axes={0: 3, 1: 1}).shape == (3, 1)
    assert ort_io.scalar_for(meta("s", [], "tensor(int64)"), 5).shape == ()
    assert ort_io.scalar_for(meta("s", [1], "tensor(float)"), 0.5).tolist() == [0.5]
    assert set(ort_io.metadata_by_name([m, kv])) == {"x", "in_de_key_layer_0"}


def test_ort_io_metadata_catalog():
    md = {"max_seq_len": "448", "ids": "1,2,,3", "special_token_ids": json.dumps({"decoder_start": 3, "stop": [2]}),
          "supported_languages": json.dumps({" en ": {"name": " English ", "aliases": [" english", "eng"], "token_id": 10},
                                             "zh": {"aliases": ["chinese", "mandarin"], "token_id": 11},
                                             "yue": {"aliases": ["chinese"], "token_id": 12}})}
    assert ort_io.metadata_int(md, "max_seq_len") == 448 and ort_io.metadata_int_list(md, "ids") == [1, 2, 3]
    assert ort_io.load_special_token_ids(md)["stop"] == [2]
    cat = ort_io.load_supported_languages(md)
    assert list(cat) == ["en", "zh", "yue"] and cat["en"]["name"] == "English" and cat["zh"]["prompt_token_ids"] == []
    assert ort_io.resolve_supported_language(cat, "EN")[0] == "en"
    assert ort_io.resolve_supported_language(cat, "Mandarin")[0] == "zh"
    with pytest.raises(ValueError, match="Unsupported language"):
        ort_io.resolve_supported_language(cat, "chinese")          # ambiguous alias
    with pytest.raises(ValueError):
        ort_io.resolve_supported_language(cat, "klingon")


@pytest.mark.skipif(not (REF / "ORT_IO.py").exists(), reason="reference checkout not present (GPU box)")
def test_ort_io_matches_reference_module():
    sys.path.insert(0, str(REF))
    try:
        import ORT_IO as ref
    finally:
        sys.path.pop(0)
    rng = np.random.default_rng(0)
    metas = [meta("a", [1, 1, "n"], "tensor(int16)"), meta("b", ["batch", 4, "h", 64], "tensor(float16)"),
             meta("c", [1], "tensor(int64)"), meta("d", [], "tensor(float)")]
    for mm in metas:
        assert ort_io.numpy_dtype(mm) == ref.numpy_dtype(mm)
    v = rng.integers(-5, 5, size=(1, 1, 9))
    assert np.array_equal(ort_io.array_for(metas[0], v), ref.array_for(metas[0], v))
    assert np.array_equal(ort_io.filled_for(metas[1], 2, axes={0: 1, 2: 3}), ref.filled_for(metas[1], 2, axes={0: 1, 2: 3}))
    for mm, val in ((metas[2], 7), (metas[3], 0.25)):
        assert np.array_equal(ort_io.scalar_for(mm, val), ref.scalar_for(mm, val))
    for fn in ("array_for",):
        with pytest.raises(ValueError) as e1:
            getattr(ort_io, fn)(metas[1], [1.0])
        with pytest.raises(ValueError) as e2:
            getattr(ref, fn)(metas[1], [1.0])
        assert str(e1.value) == str(e2.value)

"""The reference driver's own functions against the product's session facade (CPU; needs the reference checkout).

oracle/gen_script_golden.py AST-extracts `_plan_merged_io`, `_probe_prefill`, `_prefill`, `_decode_tokens`, `_run_no_speech`
(+ helpers) from /root/reference/Whisper/Inference_Whisper_ONNX.py and runs them unmodified against b200asr.session
(over the oracle-backed stand-in engine, oracle/cpu_engine.py).  Here: (1) that run reproduces the committed streams
(tests/golden/whisper_script.json), (2) the restated per-clip loop tests/script_loop.py -- what the GPU test uses, since the
GPU box has no reference checkout -- produces exactly the same streams through the same facade, (3) the plain greedy
stream equals the oracle's own greedy loop."""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

REF = Path("/root/reference/Whisper/Inference_Whisper_ONNX.py")
GOLD = Path(__file__).parent / "golden"
pytestmark = pytest.mark.skipif(not REF.exists(), reason="reference checkout not present (GPU box)")


def _factory(g):
    from oracle import whisper_oracle as wo
    from oracle.cpu_engine import OracleWhisperEngine
    from b200asr.config import WHISPER_TINY_TEST
    fw = wo.fold_weights(wo.make_raw_weights(wo.TINY_TEST, int(g["seed"])), wo.TINY_TEST, g["suppress"].tolist(), g["begin_suppress"].tolist())
    return OracleWhisperEngine(WHISPER_TINY_TEST, fw, g["suppress"].tolist())


def _golden():
    return json.loads((GOLD / "whisper_script.json").read_text())


@pytest.mark.parametrize("i", range(4))
def test_reference_functions_reproduce_committed_streams(i):
    from oracle import gen_script_golden as gs
    gold = _golden()["cases"][i]
    res = gs.run_case(_factory, gold["config"])
    assert res["tokens"] == gold["result"]["tokens"]
    assert res["detected_language_token"] == gold["result"]["detected_language_token"]
    if gold["result"]["no_speech_probability"] is not None:
        assert abs(res["no_speech_probability"] - gold["result"]["no_speech_probability"]) < 1e-6


@pytest.mark.parametrize("i", range(4))
def test_restated_loop_equals_reference_functions(i):
    from b200asr.session import WhisperSessions
    from script_loop import drive_case
    meta = _golden()
    gold = meta["cases"][i]
    cfg = gold["config"]
    g = dict(np.load(GOLD / f"whisper_tiny_case{cfg['case']}.npz"))
    S = WhisperSessions(_factory(g), {}, strategy=cfg["strategy"], no_speech_token=meta["no_speech_token"],
                        repeat_penalty=cfg["repeat_penalty"], penalty_range=cfg["penalty_range"])
    res = drive_case(S, g, cfg, meta)
    for k in ("tokens", "detected_language_token", "decode_steps"):
        assert res.get(k) == gold["result"].get(k), k
    if cfg["stop_at"] is not None:
        assert res["stop_token"] == gold["result"]["stop_token"]


def test_plain_greedy_stream_equals_oracle_loop():
    from oracle import whisper_oracle as wo
    meta = _golden()
    gold = meta["cases"][0]
    g = dict(np.load(GOLD / "whisper_tiny_case0.npz"))
    fw = wo.fold_weights(wo.make_raw_weights(wo.TINY_TEST, int(g["seed"])), wo.TINY_TEST, g["suppress"].tolist(), g["begin_suppress"].tolist())
    with torch.no_grad():
        ref = wo.greedy_transcribe(g["pcm"], fw, wo.TINY_TEST, g["prompt"].tolist(), stop_tokens=[], max_new=meta["max_seq_len"] - 4,
                                   return_logits=False)
    assert ref["tokens"] == gold["result"]["tokens"]

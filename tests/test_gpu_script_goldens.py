"""The CUDA engine behind the ORT-shaped facade against token streams minted by the REFERENCE DRIVER'S OWN FUNCTIONS.

tests/golden/whisper_script.json was produced by oracle/gen_script_golden.py: `_plan_merged_io`, `_probe_prefill`, `_prefill`,
`_decode_tokens`, `_run_no_speech` AST-extracted from Whisper/Inference_Whisper_ONNX.py and run unmodified against
b200asr.session.  The GPU box has no reference checkout, so the engine is driven here by the restated per-clip loop
(tests/script_loop.py), which tests/test_script_goldens_cpu.py holds to those functions on the same facade.  fp32 engine:
identical streams, language token and no-speech probability (1e-4); bf16 engine (the streaming tcgen05 decode kernel):
identical streams on these cases."""
import json
from pathlib import Path

import numpy as np
import pytest

from gpu_common import load_case, make_engine
from b200asr.session import WhisperSessions
from script_loop import drive_case

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"
META = json.loads((GOLD / "whisper_script.json").read_text())


@pytest.mark.parametrize("precision", ["f32", "bf16"])
@pytest.mark.parametrize("i", range(4))
def test_engine_behind_facade_reproduces_reference_function_streams(i, precision):
    gold = META["cases"][i]
    cfg = gold["config"]
    g, raw, tensors = load_case(GOLD / f"whisper_tiny_case{cfg['case']}.npz")
    eng = make_engine(tensors, precision)
    S = WhisperSessions(eng, {}, strategy=cfg["strategy"], no_speech_token=META["no_speech_token"],
                        repeat_penalty=cfg["repeat_penalty"], penalty_range=cfg["penalty_range"])
    res = drive_case(S, g, cfg, META)
    eng.close()
    assert res["tokens"] == gold["result"]["tokens"]
    assert res["detected_language_token"] == gold["result"]["detected_language_token"]
    assert res["decode_steps"] == gold["result"]["decode_steps"]
    if gold["result"]["no_speech_probability"] is not None:
        assert abs(res["no_speech_probability"] - gold["result"]["no_speech_probability"]) < (1e-4 if precision == "f32" else 5e-3)

"""TEST INFRASTRUCTURE ONLY.  Runs the reference driver's OWN functions against the session facade and records what they
produce.

    python -m oracle.gen_script_golden        (needs /root/reference; writes tests/golden/whisper_script.json)

`_plan_merged_io`, `_probe_prefill`, `_prefill`, `_decode_tokens`, `_run_no_speech` and their small helpers
(/root/reference/Whisper/Inference_Whisper_ONNX.py:247-268,323-392,410-435,437-663,691-699) are AST-extracted from the
script (it cannot be imported: module-level code parses a command line, loads ONNX files and runs the examples) and
executed UNMODIFIED in a namespace where
  * `onnxruntime` and `C` resolve to b200asr.session (OrtValue, OrtDevice) -- the product's ORT-shaped facade,
  * PROBE_SESSION / PREFILL_SESSION / DECODE_SESSION / NO_SPEECH_SESSION are the facade's sessions over an engine,
  * the ORT_IO helpers are the reference's own ORT_IO.py, the script's constants are set as the script sets them.
Here (no GPU) the engine behind the facade is the CPU stand-in oracle/cpu_engine.py; on the GPU box
tests/test_gpu_script_goldens.py puts the CUDA engine behind the same facade, drives it with the product's host loop and
must reproduce these streams.  The per-clip glue between the functions (probe -> language arg-max -> no-speech -> prefill
-> decode) follows the script's main loop :766-823 line by line."""
from __future__ import annotations

import ast
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
SCRIPT = REF / "Whisper" / "Inference_Whisper_ONNX.py"
WANT = ["_run", "_in_names", "_out_names", "_ort_value", "_bind_device_outputs", "_plan_merged_io", "_self_kv_sequence_axis",
        "_empty_self_kv", "_bind_typed", "_prefill", "_probe_prefill", "_decode_static_inputs", "_bind_sampling_controls",
        "_decode_tokens", "_run_no_speech"]


def extract_functions():
    tree = ast.parse(SCRIPT.read_text())
    found = {n.name: n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in WANT}
    missing = [w for w in WANT if w not in found]
    if missing:
        raise SystemExit(f"reference script no longer defines {missing}")
    return ast.Module(body=[found[w] for w in WANT], type_ignores=[])


def build_namespace(sessions, *, strategy, repeat_penalty, penalty_range, no_speech):
    sys.path.insert(0, str(REF))
    import ORT_IO as ref_io                       # the reference's own helper module (pure numpy)
    from b200asr import session as shim
    ns = {"np": np, "time": time, "onnxruntime": shim, "C": shim,
          "DEVICE_TYPE": "cpu", "DEVICE_ID": 0, "ORT_DEVICE": shim.OrtDevice(), "RUN_OPTIONS": None,
          "STRATEGY": strategy, "REPEAT_PENALTY": repeat_penalty, "PENALTY_RANGE": penalty_range,
          "TEMPERATURE": 1.0, "TOP_K": 10, "TOP_P": 0.95, "SAMPLING_REPETITION_PENALTY": 1.0,
          "PROBE_SESSION": sessions.probe, "PREFILL_SESSION": sessions.prefill, "DECODE_SESSION": sessions.decode,
          "NO_SPEECH_SESSION": sessions.no_speech if no_speech else None}
    for name in ("array_for", "filled_for", "scalar_for", "metadata_by_name", "is_dynamic_dim", "numpy_dtype"):
        ns[name] = getattr(ref_io, name)
    exec(compile(extract_functions(), str(SCRIPT), "exec"), ns)
    # the module-level plan / metadata tables of the script (:394-406,420-423), built with the extracted functions
    ns["PROBE_PLAN"] = ns["_plan_merged_io"](sessions.probe, strategy, is_decode=False)
    ns["PREFILL_PLAN"] = ns["_plan_merged_io"](sessions.prefill, strategy, is_decode=False)
    ns["DECODE_PLAN"] = ns["_plan_merged_io"](sessions.decode, strategy, is_decode=True)
    ns["KV_NUM_TENSORS"] = len(ns["DECODE_PLAN"]["state_inputs"])
    ns["PREFILL_INPUT_META"] = ref_io.metadata_by_name(sessions.prefill.get_inputs())
    ns["PROBE_INPUT_META"] = ref_io.metadata_by_name(sessions.probe.get_inputs())
    ns["DECODE_INPUT_META"] = ref_io.metadata_by_name(sessions.decode.get_inputs())
    ns["DECODE_OUTPUT_META"] = ref_io.metadata_by_name(sessions.decode.get_outputs())
    for k in ("DECODE", "PREFILL", "PROBE"):
        ns[f"{k}_OUTPUT_INDEX"] = {name: i for i, name in enumerate(ns[f"{k}_PLAN"]["outputs"])}
    ns["NO_SPEECH_INPUT_META"] = sessions.no_speech.get_inputs()[0] if no_speech else None
    return ns


def run_clip(ns, pcm_i16, *, start, lang_id, task, notimestamps, stop_tokens, max_seq_len, detect_language, no_speech,
             language_token_ids, threshold=0.6):
    """One window of the script's main loop (:766-823) over the extracted functions."""
    ref_io_array_for, filled_for = ns["array_for"], ns["filled_for"]
    meta = ns["PROBE_INPUT_META"]["audio"]
    audio = np.asarray(pcm_i16, np.int16).reshape(1, 1, -1)
    audio_buffer = ns["_ort_value"](filled_for(meta, axes={0: 1, 1: 1, 2: audio.shape[-1]}))
    needs_probe = detect_language or no_speech
    probe_vals = [[start]] if needs_probe else [[start, lang_id, task, notimestamps]]
    probe_tokens = ref_io_array_for(ns["PROBE_INPUT_META"][ns["PROBE_PLAN"]["token_input"]], probe_vals, axes={0: 1, 1: len(probe_vals[0])})
    probe_outputs = ns["_probe_prefill"](audio_buffer, audio, probe_tokens)
    cross_kv = {dn: probe_outputs[ns["PROBE_OUTPUT_INDEX"][pn]]
                for pn, dn in zip(ns["PROBE_PLAN"]["cross_outputs"], ns["PREFILL_PLAN"]["cross_inputs"])}
    out = {"detected_language_token": None, "no_speech_probability": None}
    if needs_probe:
        det = probe_outputs[ns["PROBE_OUTPUT_INDEX"][ns["PROBE_PLAN"]["raw_logits_output"]]]
        if detect_language:
            logits = det.numpy().reshape(-1)
            ids = np.asarray(language_token_ids, np.int64)
            lang_id = int(ids[np.argmax(logits[ids])])
            out["detected_language_token"] = lang_id
        if no_speech:
            p = ns["_run_no_speech"](det)
            out["no_speech_probability"] = p
            if p >= threshold:
                out["tokens"] = []
                return out
    prompt_vals = [[start, lang_id, task, notimestamps]]
    prompt = ref_io_array_for(ns["PREFILL_INPUT_META"][ns["PREFILL_PLAN"]["token_input"]], prompt_vals, axes={0: 1, 1: 4})
    prefill_outputs = (ns["_prefill"](prompt, cross_kv) if needs_probe
                       else [probe_outputs[ns["PROBE_OUTPUT_INDEX"][n]] for n in ns["PREFILL_PLAN"]["outputs"]])
    limit = max(0, max_seq_len - prompt.shape[-1])
    toks, steps, _ = ns["_decode_tokens"](prefill_outputs, cross_kv, limit, set(stop_tokens))
    out["tokens"] = [int(t) for t in toks]
    out["decode_steps"] = int(steps)
    return out


CASES = [  # (golden case index, strategy, repeat_penalty, penalty_range, detect_language, no_speech, stop token taken from the free stream)
    dict(case=0, strategy="greedy", repeat_penalty=1.0, penalty_range=20, detect_language=False, no_speech=False, stop_at=None),
    dict(case=1, strategy="greedy", repeat_penalty=1.0, penalty_range=20, detect_language=True, no_speech=True, stop_at=None),
    dict(case=2, strategy="penalty_greedy", repeat_penalty=0.8, penalty_range=3, detect_language=False, no_speech=False, stop_at=None),
    dict(case=0, strategy="penalty_greedy", repeat_penalty=0.7, penalty_range=5, detect_language=True, no_speech=False, stop_at=6),
]
MAX_SEQ_LEN = 24                      # keeps the tiny-model streams short (the script's MAX_SEQ_LEN is metadata-driven)
NO_SPEECH_TOKEN = 13
LANG_IDS = [10, 11, 14, 15]


def run_case(engine_factory, c, golden_dir=ROOT / "tests" / "golden"):
    """engine_factory(tensors_or_fw_inputs) -> engine; shared by the generator (CPU stand-in) and the checks."""
    from b200asr.session import WhisperSessions
    g = dict(np.load(golden_dir / f"whisper_tiny_case{c['case']}.npz"))
    eng = engine_factory(g)
    prompt = [int(t) for t in g["prompt"].reshape(-1)]
    S = WhisperSessions(eng, {}, strategy=c["strategy"], no_speech_token=NO_SPEECH_TOKEN, repeat_penalty=c["repeat_penalty"],
                        penalty_range=c["penalty_range"])
    ns = build_namespace(S, strategy=c["strategy"], repeat_penalty=c["repeat_penalty"], penalty_range=c["penalty_range"],
                         no_speech=c["no_speech"])
    kw = dict(start=prompt[0], lang_id=prompt[1], task=prompt[2], notimestamps=prompt[3], max_seq_len=MAX_SEQ_LEN,
              detect_language=c["detect_language"], no_speech=c["no_speech"], language_token_ids=LANG_IDS, threshold=2.0)
    res = run_clip(ns, g["pcm"], stop_tokens=[], **kw)
    if c["stop_at"] is not None:                # a second pass with a stop token taken from the free-running stream
        stop = res["tokens"][c["stop_at"]]
        res2 = run_clip(ns, g["pcm"], stop_tokens=[stop], **kw)
        res = {**res2, "stop_token": int(stop), "free_tokens": res["tokens"]}
    return res


def main():
    from oracle import whisper_oracle as wo
    from oracle.cpu_engine import OracleWhisperEngine
    from b200asr.config import WHISPER_TINY_TEST

    def factory(g):
        raw = wo.make_raw_weights(wo.TINY_TEST, int(g["seed"]))
        fw = wo.fold_weights(raw, wo.TINY_TEST, g["suppress"].tolist(), g["begin_suppress"].tolist())
        return OracleWhisperEngine(WHISPER_TINY_TEST, fw, g["suppress"].tolist())

    out = []
    for c in CASES:
        r = run_case(factory, c)
        out.append({"config": c, "result": r})
        print(c, "->", {k: v for k, v in r.items() if k != "free_tokens"})
    (ROOT / "tests" / "golden" / "whisper_script.json").write_text(json.dumps(
        {"source": "reference functions of Whisper/Inference_Whisper_ONNX.py run against b200asr.session over oracle/cpu_engine.py",
         "max_seq_len": MAX_SEQ_LEN, "no_speech_token": NO_SPEECH_TOKEN, "language_token_ids": LANG_IDS, "cases": out}, indent=1))


if __name__ == "__main__":
    main()

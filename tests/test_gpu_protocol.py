"""GPU tests of the reference-facing layers above the C ABI: the sampling head (a11), the default driver protocol
(probe -> language -> no-speech -> prefill -> decode, a12/f1) and the ORT-shaped session shim (a17), each against
the CPU oracle on the tiny seeded model in fp32 (tokens exact, logits 1e-3)."""
import json

import numpy as np
import pytest
import torch

from gpu_common import GOLD, NO_SPEECH, load_case, make_engine, maxdiff
from oracle import whisper_oracle as wo
from b200asr.ort_io import array_for, filled_for, metadata_by_name, scalar_for
from b200asr.session import OrtValue, WhisperSessions
from b200asr.whisper_infer import InferenceOptions, WhisperPipeline

pytestmark = pytest.mark.gpu


def _oracle_weights(g):
    raw = wo.make_raw_weights(wo.TINY_TEST, int(g["seed"]))
    return wo.fold_weights(raw, wo.TINY_TEST, g["suppress"].tolist(), g["begin_suppress"].tolist())


def _metadata(g):
    langs = {f"l{int(t)}": {"name": f"lang{int(t)}", "aliases": [f"alias{int(t)}"], "token_id": int(t)} for t in g["lang_ids"]}
    p = g["prompt"].tolist()
    langs["en"] = {"name": "English", "aliases": ["english"], "token_id": int(p[1])}
    return {"audio_pcm_scale": "32768", "max_seq_len": "448", "sample_rate": "16000",
            "special_token_ids": json.dumps({"decoder_start": int(p[0]), "tasks": {"transcribe": int(p[2]), "translate": 9},
                                             "no_timestamps": int(p[3]), "stop": [2], "no_speech": NO_SPEECH}),
            "supported_languages": json.dumps(langs)}


@pytest.mark.parametrize("cfg", [(0.8, 10, 0.95, 1.0), (1.3, 5, 0.6, 1.4), (0.5, 64, 0.99, 1.1)])
def test_sampling_matches_oracle(cfg):
    t, k, p, rp = cfg
    g, raw, tensors = load_case(GOLD[0])
    fw = _oracle_weights(g)
    noise = np.random.default_rng(5).random((12, k)).astype(np.float32)
    with torch.no_grad():
        ref = wo.sampling_transcribe(g["pcm"], fw, wo.TINY_TEST, g["prompt"].tolist(), [], 9, t, k, p, rp, noise)
    eng = make_engine(tensors, "f32")
    eng.set_decode_options(stop_ids=[], generate_limit=9)
    eng.set_sampling(temperature=t, top_k=k, top_p=p, repetition_penalty=rp, noise=noise.reshape(12, 1, k))
    got = eng.transcribe(g["pcm"], g["prompt"], max_new=9)[0]
    assert got == ref["tokens"]
    assert len(set(got)) > 1 or k == 1
    # back to argmax heads
    eng.set_sampling(temperature=0.0)
    eng.set_decode_options(stop_ids=[], generate_limit=7)
    assert eng.transcribe(g["pcm"], g["prompt"], max_new=7)[0] == g["free_tokens"].tolist()
    # without supplied noise the head still samples inside the top-k set and is reproducible per seed
    eng.set_sampling(temperature=t, top_k=k, top_p=p, repetition_penalty=rp, seed=11)
    a = eng.transcribe(g["pcm"], g["prompt"], max_new=7)[0]
    b = eng.transcribe(g["pcm"], g["prompt"], max_new=7)[0]
    assert a == b and len(a) == 7
    eng.close()


@pytest.mark.parametrize("path", GOLD[:2], ids=[p.stem for p in GOLD[:2]])
def test_default_protocol_pipeline(path):
    g, raw, tensors = load_case(path)
    fw = _oracle_weights(g)
    md = _metadata(g)
    eng = make_engine(tensors, "f32")
    opt = InferenceOptions(REPEAT_PENALTY=0.8, PENALTY_RANGE=3, NO_SPEECH_THRESHOLD=2.0)   # never classify as silence here
    pipe = WhisperPipeline(eng, md, opt)
    assert pipe.strategy == "penalty_greedy"
    eng_limit = 9
    pipe.max_seq_len = 448
    res = pipe.transcribe_pcm(g["pcm"])
    assert res.language_token == int(g["detected_language"])
    np.testing.assert_allclose(res.no_speech_probability, float(g["no_speech_prob"][0]), rtol=2e-3, atol=1e-7)
    prompt = [int(g["prompt"][0]), res.language_token, int(g["prompt"][2]), int(g["prompt"][3])]
    with torch.no_grad():
        ref = wo.greedy_transcribe(g["pcm"], fw, wo.TINY_TEST, prompt, stop_tokens=[2], max_new=24,
                                   repeat_penalty=0.8, penalty_range=3, return_logits=False)
    assert res.tokens[:len(ref["tokens"])] == ref["tokens"] or res.tokens[:24] == ref["tokens"][:24]
    assert "RTF:" in pipe.report(res, "text")
    # silence path: threshold below the measured probability -> no transcription, like the reference (:800-805)
    pipe2 = WhisperPipeline(eng, md, InferenceOptions(NO_SPEECH_THRESHOLD=0.0))
    r2 = pipe2.transcribe_pcm(g["pcm"])
    assert r2.no_speech and r2.tokens == []
    # sliding windows: two windows of 1 s each decode twice
    pipe3 = WhisperPipeline(eng, md, InferenceOptions(REPEAT_PENALTY=1.0, DETECT_LANGUAGE=False, NO_SPEECH_DETECTION=False,
                                                      INPUT_AUDIO_LENGTH=16000, SLIDING_WINDOW=16000))
    eng.set_decode_options(stop_ids=[2], generate_limit=5)
    r3 = pipe3.transcribe_pcm(g["pcm"])
    assert r3.windows == 2
    eng.close()


@pytest.mark.parametrize("probe", [True, False])
def test_long_form_windows_batched_equal_sequential(probe):
    """Long-form audio: the windows after the first go through the engine as batches (`BATCH_WINDOWS`; the reference runs them one
    by one, Inference_Whisper_ONNX.py:766-827, and they share no state).  Seven 0.5 s windows with stride 0.45 s on a batch-4 engine:
    ids, window count and detected language identical to the sequential loop."""
    from b200asr.synth import synth_pcm
    g, raw, tensors = load_case(GOLD[2])
    md = _metadata(g)
    eng = make_engine(tensors, "f32", max_batch=4, max_samples=64000)
    pcm = synth_pcm(77, 50000)
    out = {}
    for batched in (True, False):
        opt = InferenceOptions(REPEAT_PENALTY=0.8, PENALTY_RANGE=3, NO_SPEECH_THRESHOLD=2.0, DETECT_LANGUAGE=probe, NO_SPEECH_DETECTION=probe,
                               INPUT_AUDIO_LENGTH=8000, SLIDING_WINDOW=7200, BATCH_WINDOWS=batched)
        pipe = WhisperPipeline(eng, md, opt)
        pipe.max_seq_len = 448
        r = pipe.transcribe_pcm(pcm)
        out[batched] = (r.tokens, r.windows, r.language_token, r.decode_steps)
    assert out[True] == out[False]
    assert out[True][1] == 7 and len(out[True][0]) > 7
    eng.close()


def test_session_shim_runs_reference_shaped_loop():
    """The probe / prefill / decode call sequence of Inference_Whisper_ONNX.py:437-663 against the shim."""
    g, raw, tensors = load_case(GOLD[1])
    md = _metadata(g)
    eng = make_engine(tensors, "f32")
    eng.set_decode_options(stop_ids=[], generate_limit=0)
    S = WhisperSessions(eng, md, strategy="greedy", no_speech_token=NO_SPEECH)
    PROBE, PREFILL, DECODE = S.probe, S.prefill, S.decode
    in_probe, in_pre, in_dec = (metadata_by_name(s.get_inputs()) for s in (PROBE, PREFILL, DECODE))
    out_probe = [m.name for m in PROBE.get_outputs()]
    out_pre = [m.name for m in PREFILL.get_outputs()]
    out_dec = [m.name for m in DECODE.get_outputs()]
    L = 2
    assert [n for n in in_dec][:2 * L] == [f"in_de_key_layer_{i}" for i in range(L)] + [f"in_de_value_layer_{i}" for i in range(L)]
    assert out_pre[-3:] == ["argmax_max_logits_idx", "logits", "prefill_kv_seq_len"]
    assert out_dec[-2:] == ["argmax_max_logits_idx", "decode_kv_seq_len_next"]
    assert DECODE.get_modelmeta().custom_metadata_map["max_seq_len"] == "448"

    def bind_common(binding, meta, ids):
        for name in [n for n in meta if n.startswith("in_de_")]:
            axis = 3 if "key" in name else 2
            binding.bind_ortvalue_input(name, OrtValue.ortvalue_from_numpy(filled_for(meta[name], axes={0: 1, axis: 0})))
        binding.bind_ortvalue_input("embed_input_ids", OrtValue.ortvalue_from_numpy(array_for(meta["embed_input_ids"], ids, axes={0: 1, 1: len(ids[0])})))
        binding.bind_ortvalue_input("prefill_ids_len", OrtValue.ortvalue_from_numpy(scalar_for(meta["prefill_ids_len"], len(ids[0]))))
        binding.bind_ortvalue_input("prefill_history_len", OrtValue.ortvalue_from_numpy(scalar_for(meta["prefill_history_len"], 0)))

    # probe([SOT]) with the audio
    b = PROBE.io_binding()
    audio = OrtValue.ortvalue_from_numpy(filled_for(in_probe["audio"], axes={2: len(g["pcm"])}))
    audio.update_inplace(array_for(in_probe["audio"], g["pcm"].reshape(1, 1, -1)))
    b.bind_ortvalue_input("audio", audio)
    bind_common(b, in_probe, [[int(g["prompt"][0])]])
    for n in out_probe:
        b._iobinding.bind_output(n, None)
    PROBE.run_with_iobinding(b)
    outs = dict(zip(out_probe, b.get_outputs()))
    assert maxdiff(outs["logits"].numpy()[0], g["probe_logits"]) <= 1e-3
    nb = S.no_speech.io_binding()
    nb.bind_ortvalue_input("logits", outs["logits"])
    S.no_speech.run_with_iobinding(nb)
    np.testing.assert_allclose(nb.get_outputs()[0].numpy(), g["no_speech_prob"], rtol=2e-3, atol=1e-7)
    np.testing.assert_allclose(outs["encoder_en_key_layer_0"].numpy(), g["cross_k_layer0"], atol=1e-3)
    cross = {n.replace("encoder_", ""): v for n, v in outs.items() if n.startswith("encoder_en_")}
    # prefill with the full prompt
    b = PREFILL.io_binding()
    for n, v in cross.items():
        b.bind_ortvalue_input(n, v)
    bind_common(b, in_pre, [g["prompt"].tolist()])
    PREFILL.run_with_iobinding(b)
    pre = dict(zip(out_pre, b.get_outputs()))
    tokens = [int(pre["argmax_max_logits_idx"].numpy().reshape(-1)[0])]
    state = [pre[n] for n in out_pre if n.startswith("out_de_")]
    next_token, kv_len = pre["argmax_max_logits_idx"], pre["prefill_kv_seq_len"]
    bindings = [DECODE.io_binding(), DECODE.io_binding()]
    for step in range(6):
        b = bindings[step & 1]
        for n, v in cross.items():
            b.bind_ortvalue_input(n, v)
        b.bind_ortvalue_input("embed_input_ids", next_token)
        b.bind_ortvalue_input("decode_kv_seq_len", kv_len)
        for n, v in zip([m for m in in_dec if m.startswith("in_de_")], state):
            b.bind_ortvalue_input(n, v)
        b.clear_binding_outputs()
        DECODE.run_with_iobinding(b)
        o = dict(zip(out_dec, b.get_outputs()))
        state = [o[n] for n in out_dec if n.startswith("out_de_")]
        next_token, kv_len = o["argmax_max_logits_idx"], o["decode_kv_seq_len_next"]
        tokens.append(int(next_token.numpy().reshape(-1)[0]))
    assert tokens == g["free_tokens"].tolist()
    assert int(kv_len.numpy()[0]) == 4 + 6
    np.testing.assert_allclose(state[L - 1].numpy()[0], g["self_k_last_layer"], atol=1e-3)      # out_de_key_layer_{L-1}
    # stale handles are refused
    with pytest.raises(ValueError, match="previous launch"):
        b2 = DECODE.io_binding()
        for n, v in cross.items():
            b2.bind_ortvalue_input(n, v)
        b2.bind_ortvalue_input("embed_input_ids", next_token)
        b2.bind_ortvalue_input("decode_kv_seq_len", kv_len)
        for n, v in zip([m for m in in_dec if m.startswith("in_de_")], [pre[n] for n in out_pre if n.startswith("out_de_")]):
            b2.bind_ortvalue_input(n, v)
        DECODE.run_with_iobinding(b2)
    with pytest.raises(ValueError, match="unbound inputs"):
        DECODE.run_with_iobinding(DECODE.io_binding())
    eng.close()


def test_cli_end_to_end_from_hf_folder_and_wav(tmp_path, capsys):
    """`python -m b200asr.cli whisper --model-folder F --audio clip.wav` on a synthetic HF checkpoint folder + WAV file:
    checkpoint ingest -> folds -> engine -> default protocol -> the script's report block."""
    import wave
    from b200asr import cli, ingest
    from b200asr.config import WHISPER_TINY_TEST as dims
    from b200asr.synth import synth_whisper_checkpoint
    g, raw_o, tensors = load_case(GOLD[0])
    raw = synth_whisper_checkpoint(dims, int(g["seed"]))
    cfg = {"num_mel_bins": dims.n_mels, "d_model": dims.d_model, "encoder_attention_heads": dims.n_heads,
           "decoder_attention_heads": dims.n_heads, "encoder_ffn_dim": dims.ffn, "decoder_ffn_dim": dims.ffn,
           "encoder_layers": dims.enc_layers, "decoder_layers": dims.dec_layers, "vocab_size": dims.vocab,
           "max_source_positions": dims.max_source, "max_target_positions": dims.max_target}
    p = g["prompt"].tolist()
    gen = {"suppress_tokens": g["suppress"].tolist(), "begin_suppress_tokens": g["begin_suppress"].tolist(),
           "lang_to_id": {f"<|l{int(t)}|>": int(t) for t in g["lang_ids"]} | {"<|en|>": p[1]},
           "task_to_id": {"transcribe": p[2]}, "no_timestamps_token_id": p[3], "decoder_start_token_id": p[0], "eos_token_id": 2,
           "no_speech_token_id": NO_SPEECH}
    (tmp_path / "config.json").write_text(json.dumps(cfg))
    (tmp_path / "generation_config.json").write_text(json.dumps(gen))
    ingest.write_safetensors(tmp_path / "model.safetensors", {k: v.numpy() for k, v in raw.items()})
    with wave.open(str(tmp_path / "clip.wav"), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000)
        w.writeframes(g["pcm"].astype("<i2").tobytes())
    rc = cli.main(["whisper", "--model-folder", str(tmp_path), "--audio", str(tmp_path / "clip.wav"), "--precision", "f32",
                   "--set", "REPEAT_PENALTY=1.0", "NO_SPEECH_THRESHOLD=2.0"])
    out = capsys.readouterr().out
    assert rc == 0 and "ASR Result:" in out and "RTF:" in out and "Detected Language:" in out
    ids = [int(t) for t in out.split("ASR Result:\n")[1].split("\n")[0].split()]
    # greedy, language detected from the probe: the oracle with the same prompt gives the same ids (EOS = 2 stops both)
    det = int(g["detected_language"])
    with torch.no_grad():
        ref = wo.greedy_transcribe(g["pcm"], _oracle_weights(g), wo.TINY_TEST, [p[0], det, p[2], p[3]], stop_tokens=[2], max_new=40,
                                   return_logits=False)
    assert ids[:40] == ref["tokens"][:40]

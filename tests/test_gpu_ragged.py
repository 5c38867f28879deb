"""Ragged batches through the C ABI (b200asr_encode_ragged / _transcribe_ragged): clips of different lengths in one
batch.  The reference's audio axis is dynamic (Whisper/Export_Whisper.py:743), i.e. a caller with mixed lengths runs the
graph clip by clip; the engine must give every clip of a ragged batch what it gets alone: its own reflect pad, log-mel
maximum, conv zero padding and attention key range.

Checked three ways: (1) against the CPU oracle run on each clip alone (fp32 engine, north_star's 1e-3 on logits, tokens
exact), (2) against the engine's own single-clip results (fp32: 1e-4, measured 0; bf16: 2e-3 on logits, measured
1.4e-4 with a bit-identical encoder output -- the decode kernel's row class differs between batch 4 and batch 1), (3) every decoder path that
carries the per-clip key count (tensor-core streaming kernel, grid-barrier kernel, per-op graph) agrees."""
import numpy as np
import pytest

from gpu_common import GOLD, load_case, make_engine, maxdiff
from oracle import whisper_oracle as wo
from b200asr.engine import WhisperEngine
from b200asr.synth import synth_pcm

pytestmark = pytest.mark.gpu

LENS = [24160, 16000, 31840, 8000]          # 151 / 100 / 199 / 50 mel frames: odd and even, shortest = 25 encoder positions


def _clips():
    return [synth_pcm(40 + i, n) for i, n in enumerate(LENS)]


def _run(eng, pcm, prompt, forced, lens=None):
    eng.encode(pcm, lens=lens)
    eng.set_decode_options(stop_ids=[])
    logits, tok = eng.prefill(prompt)
    out = [logits.copy()]
    for t in forced:
        logits, tok = eng.decode_step(token_in=np.full(eng.batch, t, np.int32))
        out.append(logits.copy())
    return np.stack(out, axis=1)        # [B, steps, vocab]


def test_ragged_f32_vs_oracle_per_clip():
    g, raw, tensors = load_case(GOLD[0])
    fw = wo.fold_weights(raw, wo.TINY_TEST, g["suppress"].tolist(), g["begin_suppress"].tolist())
    clips = _clips()
    pcm, lens = WhisperEngine.pad_ragged(clips)
    # what lies beyond a clip's end in its row must not matter
    rng = np.random.default_rng(0)
    for b, n in enumerate(LENS):
        pcm[b, n:] = rng.integers(-3000, 3000, pcm.shape[1] - n)
    forced = g["forced_tokens"].tolist()[:3]
    eng = make_engine(tensors, "f32", max_batch=4)
    lg = _run(eng, pcm, g["prompt"], forced, lens=lens)
    T = eng.T_enc
    enc = eng.get_stage("enc_out", 4 * T * 256).reshape(4, T, 256)
    eng.set_decode_options(stop_ids=[], generate_limit=5)
    toks = eng.transcribe(pcm, g["prompt"], max_new=5, lens=lens)
    worst = 0.0
    for b, clip in enumerate(clips):
        ref = wo.greedy_transcribe(clip, fw, wo.TINY_TEST, g["prompt"].tolist(), stop_tokens=[], max_new=4, forced_tokens=forced)
        d = maxdiff(lg[b], np.stack(ref["step_logits"]))
        worst = max(worst, d)
        assert d <= 1e-3, (b, d)
        free = wo.greedy_transcribe(clip, fw, wo.TINY_TEST, g["prompt"].tolist(), stop_tokens=[], max_new=5, return_logits=False)
        assert toks[b] == free["tokens"], b
    print("ragged f32 vs oracle max |dlogit| =", worst)
    # padding rows exist but are finite (they are never attended to)
    assert np.isfinite(enc).all()
    eng.close()


@pytest.mark.parametrize("precision", ["f32", "bf16"])
def test_ragged_batch_equals_single(precision):
    g, raw, tensors = load_case(GOLD[1])
    clips = _clips()
    pcm, lens = WhisperEngine.pad_ragged(clips)
    forced = g["forced_tokens"].tolist()[:3]
    eng = make_engine(tensors, precision, max_batch=4)
    lb = _run(eng, pcm, g["prompt"], forced, lens=lens)
    T = eng.T_enc
    enc = eng.get_stage("enc_out", 4 * T * 256).reshape(4, T, 256).copy()
    tv = eng.valid_positions(lens)
    tol_l, tol_e = (1e-4, 1e-4) if precision == "f32" else (2e-3, 1e-3)    # measured on B200: f32 0 / 0, bf16 1.4e-4 / 0
    for b, clip in enumerate(clips):
        ls = _run(eng, clip, g["prompt"], forced)
        es = eng.get_stage("enc_out", int(tv[b]) * 256).reshape(int(tv[b]), 256)
        d, de = maxdiff(lb[b], ls[0]), maxdiff(enc[b, :tv[b]], es)
        print(precision, "clip", b, "ragged vs single: |dlogit|", d, "|d enc_out|", de)
        assert d <= tol_l and de <= tol_e
    eng.set_decode_options(stop_ids=[], generate_limit=6)
    tb = eng.transcribe(pcm, g["prompt"], max_new=6, lens=lens)
    ts = [eng.transcribe(c, g["prompt"], max_new=6)[0] for c in clips]
    if precision == "f32":
        assert tb == ts
    # a batch whose lens are all equal to the stride is the uniform path
    same = np.stack([clips[0], clips[0][::-1].copy()])
    assert eng.transcribe(same, g["prompt"], max_new=6, lens=[LENS[0], LENS[0]]) == eng.transcribe(same, g["prompt"], max_new=6)
    eng.close()


def test_ragged_decoder_paths_agree_bf16():
    """streaming tcgen05 kernel, grid-barrier kernel and per-op graph all mask a clip's cross-attention at its own length."""
    g, raw, tensors = load_case(GOLD[2])
    pcm, lens = WhisperEngine.pad_ragged(_clips())
    forced = g["forced_tokens"].tolist()[:3]
    outs = []
    for opts in ({}, {"stream": 0}, {"stream": 0, "mega": 0}):
        eng = make_engine(tensors, "bf16", max_batch=4)
        for k, v in opts.items():
            eng.set_option(k, v)
        outs.append(_run(eng, pcm, g["prompt"], forced, lens=lens))
        eng.close()
    for o in outs[1:]:
        d = maxdiff(outs[0], o)
        print("decoder paths on a ragged batch: max |dlogit| =", d)
        assert d <= 1e-2


def test_ragged_after_uniform_on_the_same_engine_and_shape():
    """The per-op decoder replays one CUDA graph per step; its cross-attention nodes carry the per-clip key-count pointer, so a
    ragged batch after a uniform batch of the same shape (and back) must not reuse the other's graph."""
    g, raw, tensors = load_case(GOLD[0])
    clips = _clips()[:2]
    pcm, lens = WhisperEngine.pad_ragged(clips)
    forced = g["forced_tokens"].tolist()[:3]
    for opts in ({"stream": 0, "mega": 0}, {"stream": 0}, {}):
        eng = make_engine(tensors, "bf16", max_batch=2)
        for k, v in opts.items():
            eng.set_option(k, v)
        u1 = _run(eng, pcm, g["prompt"], forced)                   # uniform: both rows at full length (padding is audio here)
        r1 = _run(eng, pcm, g["prompt"], forced, lens=lens)
        u2 = _run(eng, pcm, g["prompt"], forced)
        r2 = _run(eng, pcm, g["prompt"], forced, lens=lens)
        assert np.array_equal(u1, u2) and np.array_equal(r1, r2), opts
        assert maxdiff(u1[1], r1[1]) > 1e-3, opts                  # the shorter clip really is treated differently
        eng.close()


def test_ragged_argument_checks():
    g, raw, tensors = load_case(GOLD[0])
    eng = make_engine(tensors, "bf16", max_batch=2)
    pcm = np.zeros((2, 16000), np.int16)
    with pytest.raises(Exception):
        eng.encode(pcm, lens=[16000, 100])           # shorter than one FFT window
    with pytest.raises(Exception):
        eng.encode(pcm, lens=[16000, 16001])         # longer than the row stride
    with pytest.raises(Exception):
        eng.encode(pcm, lens=[8000, 8000])           # stride is not the longest clip
    with pytest.raises(ValueError):
        eng.encode(pcm, lens=[16000])
    eng.close()


def test_pipeline_transcribe_batch_equals_clip_by_clip():
    """WhisperPipeline.transcribe_batch: the reference's default protocol (probe -> language -> no-speech -> prefill -> penalty-greedy
    decode, Inference_Whisper_ONNX.py:766-827) for four clips of different lengths in ONE ragged batch gives every clip the
    result of running it alone through `transcribe_pcm` (fp32 engine: ids, detected language and no-speech probability)."""
    from b200asr.cli import whisper_metadata
    from b200asr.config import WHISPER_TINY_TEST as dims
    from b200asr.whisper_infer import InferenceOptions, WhisperPipeline
    g, raw, tensors = load_case(GOLD[0])
    p = g["prompt"].tolist()
    gen = {"lang_to_id": {f"<|l{int(t)}|>": int(t) for t in g["lang_ids"]} | {"<|en|>": p[1]}, "task_to_id": {"transcribe": p[2]},
           "no_timestamps_token_id": p[3], "decoder_start_token_id": p[0], "eos_token_id": 2, "no_speech_token_id": 13}
    md = whisper_metadata(dims, gen)
    clips = _clips()
    eng = make_engine(tensors, "f32", max_batch=4)
    pipe = WhisperPipeline(eng, md, InferenceOptions(NO_SPEECH_THRESHOLD=2.0, PENALTY_RANGE=3))
    batch = pipe.transcribe_batch(clips)
    for b, clip in enumerate(clips):
        one = pipe.transcribe_pcm(clip)
        assert batch[b].tokens == one.tokens, b
        assert batch[b].language_token == one.language_token
        assert abs(batch[b].no_speech_probability - one.no_speech_probability) <= 1e-5
    assert len({r.language_token for r in batch}) >= 1 and all(len(r.tokens) > 0 for r in batch)
    eng.close()


def test_cli_batch_mode_end_to_end(tmp_path, capsys):
    """`python -m b200asr.cli whisper --model-folder F --audio a.wav b.wav c.wav --batch 4`: three WAV files of different lengths
    in one ragged batch; every file's report block carries the ids the oracle gives that clip alone (language detected per clip)."""
    import json
    import wave
    import torch
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
           "no_speech_token_id": 13}
    (tmp_path / "config.json").write_text(json.dumps(cfg))
    (tmp_path / "generation_config.json").write_text(json.dumps(gen))
    ingest.write_safetensors(tmp_path / "model.safetensors", {k: v.numpy() for k, v in raw.items()})
    clips = _clips()[:3]
    paths = []
    for i, c in enumerate(clips):
        path = tmp_path / f"clip{i}.wav"
        with wave.open(str(path), "wb") as w:
            w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000)
            w.writeframes(c.astype("<i2").tobytes())
        paths.append(str(path))
    rc = cli.main(["whisper", "--model-folder", str(tmp_path), "--audio", *paths, "--precision", "f32", "--batch", "4",
                   "--set", "REPEAT_PENALTY=1.0", "NO_SPEECH_THRESHOLD=2.0"])
    out = capsys.readouterr().out
    assert rc == 0 and out.count("ASR Result:") == 3
    blocks = out.split("Test Input Audio: ")[1:]
    fw = wo.fold_weights(raw_o, wo.TINY_TEST, g["suppress"].tolist(), g["begin_suppress"].tolist())
    lang_names = {v: k[2:-2] for k, v in gen["lang_to_id"].items()}
    for blk, clip in zip(blocks, clips):
        ids = [int(t) for t in blk.split("ASR Result:\n")[1].split("\n")[0].split()]
        with torch.no_grad():
            probe = wo.greedy_transcribe(clip, fw, wo.TINY_TEST, [p[0]], stop_tokens=[], max_new=1, return_logits=True)
        lang_ids = g["lang_ids"].astype(int).tolist() + [p[1]]
        det = lang_ids[int(np.argmax(np.asarray(probe["step_logits"][0])[lang_ids]))]
        assert f"Detected Language: {lang_names[det]}" in blk
        with torch.no_grad():
            ref = wo.greedy_transcribe(clip, fw, wo.TINY_TEST, [p[0], det, p[2], p[3]], stop_tokens=[2], max_new=40, return_logits=False)
        assert ids[:40] == ref["tokens"][:40]

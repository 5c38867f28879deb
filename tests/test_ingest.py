"""CPU tests of checkpoint / audio ingest: the numpy safetensors reader against the `safetensors`
library, an HF Whisper folder round trip into the engine's folded tensor set, WAV reading and resampling."""
import json
import wave

import numpy as np
import pytest
import torch

from b200asr import ingest
from b200asr.config import WHISPER_TINY_TEST
from b200asr.synth import synth_whisper_checkpoint
from b200asr.weights import fold_whisper


def test_safetensors_reader_matches_library(tmp_path):
    st = pytest.importorskip("safetensors.torch")
    t = {"a.weight": torch.randn(3, 5), "b": torch.randn(7).to(torch.bfloat16), "c": torch.arange(6, dtype=torch.int64).reshape(2, 3),
         "h": torch.randn(2, 2).half()}
    p = tmp_path / "m.safetensors"
    st.save_file(t, str(p), metadata={"format": "pt"})
    got = ingest.read_safetensors(p)
    assert set(got) == set(t)
    assert np.array_equal(got["a.weight"], t["a.weight"].numpy())
    assert np.array_equal(got["b"], t["b"].float().numpy())            # bf16 widened exactly
    assert np.array_equal(got["c"], t["c"].numpy()) and got["h"].dtype == np.float16
    # and our writer is readable by the library
    p2 = tmp_path / "w.safetensors"
    ingest.write_safetensors(p2, {"x": np.arange(12, dtype=np.float32).reshape(3, 4)}, {"k": "v"})
    back = st.load_file(str(p2))
    assert torch.equal(back["x"], torch.arange(12, dtype=torch.float32).reshape(3, 4))
    with pytest.raises(ValueError):
        (tmp_path / "bad").write_bytes(b"\x01\x02")
        ingest.read_safetensors(tmp_path / "bad")


def test_hf_folder_to_engine_tensors(tmp_path):
    dims = WHISPER_TINY_TEST
    raw = synth_whisper_checkpoint(dims, 5)
    cfg = {"num_mel_bins": dims.n_mels, "d_model": dims.d_model, "encoder_attention_heads": dims.n_heads,
           "decoder_attention_heads": dims.n_heads, "encoder_ffn_dim": dims.ffn, "decoder_ffn_dim": dims.ffn,
           "encoder_layers": dims.enc_layers, "decoder_layers": dims.dec_layers, "vocab_size": dims.vocab,
           "max_source_positions": dims.max_source, "max_target_positions": dims.max_target}
    (tmp_path / "config.json").write_text(json.dumps(cfg))
    (tmp_path / "generation_config.json").write_text(json.dumps({"suppress_tokens": [1, 5], "begin_suppress_tokens": [220, 2]}))
    ingest.write_safetensors(tmp_path / "model.safetensors", {k: v.numpy() for k, v in raw.items() if k != "proj_out.weight"})
    d2, state, gen = ingest.load_hf_whisper(tmp_path)
    assert d2 == dims and gen["suppress_tokens"] == [1, 5]
    a = fold_whisper(state, d2, gen["suppress_tokens"], gen["begin_suppress_tokens"])
    b = fold_whisper(raw, dims, [1, 5], [220, 2])
    assert a.keys() == b.keys() and all(np.array_equal(a[k], b[k]) for k in a)


def _write_wav(path, x, rate, width=2, nch=1):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(nch); w.setsampwidth(width); w.setframerate(rate)
        w.writeframes(x.tobytes())


def test_wav_reader_and_resampler(tmp_path):
    rate = 16000
    t = np.arange(rate) / rate
    x = (8000 * np.sin(2 * np.pi * 440 * t)).astype("<i2")
    _write_wav(tmp_path / "m.wav", x, rate)
    pcm, r = ingest.read_wav(tmp_path / "m.wav")
    assert r == rate and np.array_equal(pcm, x)
    st = np.stack([x, -x // 2], axis=1).astype("<i2")                      # stereo: channel mean
    _write_wav(tmp_path / "s.wav", st, rate, nch=2)
    pcm2, _ = ingest.read_wav(tmp_path / "s.wav")
    assert np.array_equal(pcm2, (x.astype(np.int32) + (-x // 2).astype(np.int32)) // 2)
    x32 = (x.astype(np.int32) << 16).astype("<i4")                         # 32-bit -> 16-bit
    _write_wav(tmp_path / "w.wav", x32, rate, width=4)
    pcm3, _ = ingest.read_wav(tmp_path / "w.wav")
    assert np.array_equal(pcm3, x)
    x48 = (8000 * np.sin(2 * np.pi * 440 * np.arange(48000) / 48000)).astype("<i2")
    y = ingest.to_model_rate(x48, 48000, 16000)
    assert y.dtype == np.int16 and abs(len(y) - 16000) <= 1
    assert np.abs(y[200:-200].astype(np.int32) - x[200:len(y) - 200]).max() <= 120     # same tone after 3:1 decimation
    assert ingest.to_model_rate(x, 16000) is not None and np.array_equal(ingest.to_model_rate(x, 16000), x)


def test_cli_metadata_and_options():
    from b200asr.cli import _options, whisper_metadata
    from b200asr.ort_io import load_special_token_ids, load_supported_languages
    gen = {"lang_to_id": {"<|en|>": 50259, "<|zh|>": 50260}, "task_to_id": {"transcribe": 50360, "translate": 50359},
           "no_timestamps_token_id": 50364, "decoder_start_token_id": 50258, "eos_token_id": 50257}
    md = whisper_metadata(WHISPER_TINY_TEST, gen)
    sp = load_special_token_ids(md)
    assert sp["decoder_start"] == 50258 and sp["stop"] == [50257] and sp["no_speech"] == 50363 and sp["tasks"]["transcribe"] == 50360
    assert load_supported_languages(md)["zh"]["token_id"] == 50260 and md["max_seq_len"] == "448"
    o = _options(["REPEAT_PENALTY=1.0", "DETECT_LANGUAGE=0", "PENALTY_RANGE=7", "TARGET_LANGUAGE=zh"])
    assert o.REPEAT_PENALTY == 1.0 and o.DETECT_LANGUAGE is False and o.PENALTY_RANGE == 7 and o.TARGET_LANGUAGE == "zh"
    with pytest.raises(SystemExit):
        _options(["NOPE=1"])


def test_qwen_hf_folder_to_engine_tensors(tmp_path):
    """A Qwen3-ASR checkpoint folder (config.json + safetensors under the HF names) -> dims + folded engine tensors, equal to
    folding the in-memory state dict; a tied checkpoint (no lm_head) folds without lm_head.w."""
    import json
    from b200asr import qwen as qw
    d = qw.QWEN_TINY_TEST
    raw = qw.synth_qwen_checkpoint(d, 4)
    cfg = {"thinker_config": {
        "audio_config": {"num_mel_bins": d.n_mels, "encoder_layers": d.enc_layers, "encoder_attention_heads": d.enc_heads,
                         "encoder_ffn_dim": d.enc_ffn, "d_model": d.enc_d, "max_source_positions": d.max_source_positions,
                         "n_window": 50, "n_window_infer": 800, "output_dim": d.out_dim, "downsample_hidden_size": d.conv_ch},
        "text_config": {"vocab_size": d.vocab, "hidden_size": d.hidden, "intermediate_size": d.inter, "num_hidden_layers": d.dec_layers,
                        "num_attention_heads": d.heads, "num_key_value_heads": d.kv_heads, "head_dim": d.head_dim,
                        "rope_theta": d.rope_theta, "rms_norm_eps": d.rms_eps}}}
    (tmp_path / "config.json").write_text(json.dumps(cfg))
    ingest.write_safetensors(tmp_path / "model.safetensors", {k: v.numpy() for k, v in raw.items()})
    dims, state, tied = ingest.load_hf_qwen3_asr(tmp_path)
    assert not tied
    assert dims.to_dict() == {**d.to_dict(), "max_seq_len": qw.QwenDims().max_seq_len}
    got, want = qw.fold_qwen(state, d), qw.fold_qwen(raw, d)
    assert got.keys() == want.keys()
    for k in want:
        np.testing.assert_array_equal(got[k], want[k], err_msg=k)
    ingest.write_safetensors(tmp_path / "model.safetensors", {k: v.numpy() for k, v in raw.items() if k != "thinker.lm_head.weight"})
    _, state2, tied2 = ingest.load_hf_qwen3_asr(tmp_path)
    assert tied2 and "lm_head.w" not in qw.fold_qwen(state2, d, tie_lm_head=tied2)
    with pytest.raises(ValueError, match="n_window"):
        cfg["thinker_config"]["audio_config"]["n_window"] = 100
        ingest.qwen_dims_from_hf_config(cfg)


def test_qwen_prompt_from_tokenizer():
    """Prompt layout of Export_Qwen_ASR.py:1540-1586 from a tokenizer object (stub with the calls the exporter makes)."""
    class Tok:
        words = {"system": [11], "user": [12], "assistant": [13], "\n": [14], "language ": [15, 16], "English": [17], "Chinese": [18, 19]}
        def get_vocab(self):
            return {"<|im_start|>": 1, "<|im_end|>": 2, "<|audio_start|>": 3, "<|audio_end|>": 4, "<|endoftext|>": 5, "<asr_text>": 6}
        def encode(self, text, add_special_tokens=False):
            return self.words[text]
    prompt, tails = ingest.qwen_prompt_from_tokenizer(Tok(), ["English", "Chinese"])
    assert prompt.head_ids == (1, 11, 14)
    assert prompt.suffix_ids == (2, 14, 1, 12, 14, 3)
    assert prompt.tail_ids == (4, 2, 14, 1, 13, 14, 15, 16)
    assert prompt.stop_ids == (5, 2)
    assert tails == {"English": [17, 6], "Chinese": [18, 19, 6]}


def test_funasr_sensevoice_folder_to_engine_tensors(tmp_path):
    """A FunASR SenseVoiceSmall folder (`model.pt` with the module-tree keys the exporter walks + Kaldi `am.mvn`) -> dims and
    the checkpoint dict `fold_sensevoice` takes; folding it equals folding the in-memory checkpoint."""
    from b200asr import sensevoice as sv
    d = sv.SENSEVOICE_TINY_TEST
    raw = sv.synth_sensevoice_checkpoint(d, 2)
    from funasr_folders import write_sensevoice_folder
    write_sensevoice_folder(tmp_path, d, raw)
    dims, got = ingest.load_funasr_sensevoice(tmp_path)
    assert dims.n_heads == d.n_heads
    assert (dims.d_model, dims.ffn, dims.n_blocks0, dims.n_blocks, dims.n_tp_blocks, dims.vocab, dims.fsmn_kernel, dims.n_mels) == \
           (d.d_model, d.ffn, d.n_blocks0, d.n_blocks, d.n_tp_blocks, d.vocab, d.fsmn_kernel, d.n_mels)
    a, b = sv.fold_sensevoice(got, d, 32000), sv.fold_sensevoice(raw, d, 32000)
    assert a.keys() == b.keys()
    for k in b:
        np.testing.assert_allclose(a[k], b[k], rtol=2e-7, atol=1e-7, err_msg=k)
    with pytest.raises(ValueError, match="AddShift"):
        (tmp_path / "bad.mvn").write_text("<Nnet>\n</Nnet>\n")
        ingest.read_kaldi_cmvn(tmp_path / "bad.mvn")


def test_funasr_paraformer_folder_to_engine_tensors(tmp_path):
    """A FunASR Paraformer folder -> dims and the checkpoint dict `fold_paraformer` takes; folding equals the in-memory fold."""
    from funasr_folders import write_paraformer_folder
    from b200asr import paraformer as pfm
    d = pfm.PARAFORMER_TINY_TEST
    raw = pfm.synth_paraformer_checkpoint(d, 3)
    write_paraformer_folder(tmp_path, d, raw)
    dims, got = ingest.load_funasr_paraformer(tmp_path)
    assert dims == d
    assert got.keys() == raw.keys()
    a, b = pfm.fold_paraformer(got, d, 32000), pfm.fold_paraformer(raw, d, 32000)
    for k in b:
        np.testing.assert_allclose(a[k], b[k], rtol=2e-6, atol=1e-6, err_msg=k)
    (tmp_path / "tokens.json").write_text(json.dumps(["<blank>", "a", "b@@", "c"]))
    assert ingest.read_vocab(tmp_path / "tokens.json") == ["<blank>", "a", "b@@", "c"]
    (tmp_path / "Vocab_Paraformer.txt").write_text("x\ny\n", encoding="utf-8")
    assert ingest.read_vocab(tmp_path / "Vocab_Paraformer.txt") == ["x", "y"]

"""Checkpoint and audio ingest for the drop-in scripts (the parts that need no ONNX Runtime).

* `read_safetensors` / `write_safetensors`: the safetensors container (8-byte little-endian header length, JSON header
  {name: {dtype, shape, data_offsets}}, raw little-endian tensor bytes) read with numpy only, so an HF Whisper folder
  (`config.json` + `model.safetensors`) loads without transformers: `load_hf_whisper(folder)` -> (dims, state dict),
  which `weights.fold_whisper` turns into engine tensors -- the same folds the exporter bakes into the ONNX initialisers
  (/root/reference/Whisper/Export_Whisper.py:376-420, 527-550), applied to the checkpoint the exporter starts from
  (:14, `whisper-large-v3-turbo` by default).
* `read_wav` + `to_model_rate`: what the driver gets from pydub at
  /root/reference/Whisper/Inference_Whisper_ONNX.py:731-732 (`AudioSegment.from_file(...).set_channels(1)
  .set_frame_rate(SAMPLE_RATE).set_sample_width(2)`) for PCM WAV input: mono mix-down, polyphase resampling to the model
  rate, int16.  Compressed formats need a decoder the image does not have and are out of scope.
"""
from __future__ import annotations

import json
import struct
import wave
from math import gcd
from pathlib import Path
from typing import Dict, Tuple

import numpy as np

from .config import WhisperDims

_ST_DTYPES = {"F32": np.float32, "F16": np.float16, "F64": np.float64, "I64": np.int64, "I32": np.int32, "I16": np.int16,
              "I8": np.int8, "U8": np.uint8, "BOOL": np.bool_}


def _bf16_to_f32(raw: np.ndarray) -> np.ndarray:
    return (raw.astype(np.uint32) << 16).view(np.float32)


def read_safetensors(path) -> Dict[str, np.ndarray]:
    """name -> array (bf16 tensors come back as float32)."""
    data = Path(path).read_bytes()
    if len(data) < 8:
        raise ValueError(f"{path}: not a safetensors file")
    (n,) = struct.unpack("<Q", data[:8])
    if n <= 0 or 8 + n > len(data):
        raise ValueError(f"{path}: bad safetensors header length {n}")
    header = json.loads(data[8:8 + n].decode("utf-8"))
    base = 8 + n
    out: Dict[str, np.ndarray] = {}
    for name, info in header.items():
        if name == "__metadata__":
            continue
        lo, hi = info["data_offsets"]
        shape = tuple(info["shape"])
        buf = memoryview(data)[base + lo:base + hi]
        if info["dtype"] == "BF16":
            arr = _bf16_to_f32(np.frombuffer(buf, dtype="<u2"))
        elif info["dtype"] in _ST_DTYPES:
            arr = np.frombuffer(buf, dtype=np.dtype(_ST_DTYPES[info["dtype"]]).newbyteorder("<"))
        else:
            raise ValueError(f"{path}: tensor {name!r} has unsupported dtype {info['dtype']}")
        if arr.size != int(np.prod(shape, dtype=np.int64)):
            raise ValueError(f"{path}: tensor {name!r} has {arr.size} elements, header says {shape}")
        out[name] = arr.reshape(shape)
    return out


def write_safetensors(path, tensors: Dict[str, np.ndarray], metadata: Dict[str, str] | None = None) -> None:
    """Writer for the same container (fixtures, converted checkpoints)."""
    rev = {np.dtype(v): k for k, v in _ST_DTYPES.items()}
    header, blobs, off = {}, [], 0
    for name, arr in tensors.items():
        a = np.ascontiguousarray(arr)
        b = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
        header[name] = {"dtype": rev[np.dtype(a.dtype)], "shape": list(a.shape), "data_offsets": [off, off + len(b)]}
        blobs.append(b)
        off += len(b)
    if metadata:
        header["__metadata__"] = dict(metadata)
    h = json.dumps(header, separators=(",", ":")).encode("utf-8")
    h += b" " * ((8 - len(h) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(h)))
        f.write(h)
        for b in blobs:
            f.write(b)


def dims_from_hf_config(cfg: dict) -> WhisperDims:
    """HF `config.json` of a Whisper checkpoint -> engine dimensions."""
    if cfg.get("encoder_attention_heads") != cfg.get("decoder_attention_heads") or cfg.get("encoder_ffn_dim") != cfg.get("decoder_ffn_dim"):
        raise ValueError("encoder and decoder must share heads and ffn width")
    return WhisperDims(n_mels=int(cfg["num_mel_bins"]), d_model=int(cfg["d_model"]), n_heads=int(cfg["encoder_attention_heads"]),
                       ffn=int(cfg["encoder_ffn_dim"]), enc_layers=int(cfg["encoder_layers"]), dec_layers=int(cfg["decoder_layers"]),
                       vocab=int(cfg["vocab_size"]), max_source=int(cfg["max_source_positions"]),
                       max_target=int(cfg["max_target_positions"]))


def load_hf_whisper(folder) -> Tuple[WhisperDims, Dict[str, np.ndarray], dict]:
    """(dims, HF-named state dict, generation config) from a `WhisperForConditionalGeneration` checkpoint folder; sharded
    checkpoints (`model.safetensors.index.json`) are followed."""
    folder = Path(folder)
    cfg = json.loads((folder / "config.json").read_text())
    dims = dims_from_hf_config(cfg)
    index = folder / "model.safetensors.index.json"
    files = sorted(set(json.loads(index.read_text())["weight_map"].values())) if index.exists() else ["model.safetensors"]
    state: Dict[str, np.ndarray] = {}
    for fn in files:
        state.update(read_safetensors(folder / fn))
    if "proj_out.weight" not in state and "model.decoder.embed_tokens.weight" in state:
        state["proj_out.weight"] = state["model.decoder.embed_tokens.weight"]          # tied head stored once
    gen = {}
    if (folder / "generation_config.json").exists():
        gen = json.loads((folder / "generation_config.json").read_text())
    return dims, state, gen


# ---------------------------------------------------------------------------------------------
def read_wav(path) -> Tuple[np.ndarray, int]:
    """PCM WAV (8/16/24/32-bit integer) -> (mono int16 samples, sample rate).  Channels are averaged like pydub's
    `set_channels(1)`, wider samples are shifted down to 16 bits like `set_sample_width(2)`."""
    with wave.open(str(path), "rb") as w:
        nch, width, rate, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
        if w.getcomptype() != "NONE":
            raise ValueError(f"{path}: compressed WAV ({w.getcomptype()}) is not supported")
        raw = w.readframes(n)
    if width == 1:
        x = (np.frombuffer(raw, dtype=np.uint8).astype(np.int32) - 128) << 8
    elif width == 2:
        x = np.frombuffer(raw, dtype="<i2").astype(np.int32)
    elif width == 3:
        b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        x = ((b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)) << 8 >> 8) >> 8
    elif width == 4:
        x = np.frombuffer(raw, dtype="<i4").astype(np.int64) >> 16
    else:
        raise ValueError(f"{path}: unsupported sample width {width}")
    x = x.reshape(-1, nch)
    mono = x[:, 0] if nch == 1 else x.sum(axis=1) // nch
    return np.clip(mono, -32768, 32767).astype(np.int16), int(rate)


def to_model_rate(pcm: np.ndarray, rate: int, target: int = 16000) -> np.ndarray:
    """Polyphase resampling to the model's sample rate (metadata `sample_rate`), int16 in / int16 out."""
    if rate == target:
        return np.ascontiguousarray(pcm, dtype=np.int16)
    from scipy.signal import resample_poly
    g = gcd(int(rate), int(target))
    y = resample_poly(pcm.astype(np.float64), target // g, rate // g)
    return np.clip(np.rint(y), -32768, 32767).astype(np.int16)


# ---------------------------------------------------------------------------------------------
def qwen_dims_from_hf_config(cfg: dict):
    """HF `config.json` of a Qwen3-ASR checkpoint (thinker_config.{audio_config,text_config}; field names as the
    exporter's config classes, Qwen_ASR/Export_Qwen_ASR.py:145-236) -> engine dimensions."""
    from .qwen import CHUNK, QwenDims
    th = cfg.get("thinker_config", cfg)
    a, t = th["audio_config"], th["text_config"]
    n_window = int(a.get("n_window", 50))
    if 2 * n_window != CHUNK:
        raise ValueError(f"n_window {n_window}: the exporter's length formula assumes {CHUNK}-frame chunks (Export_Qwen_ASR.py:519-527)")
    rope = t.get("rope_theta")
    if rope is None and t.get("rope_scaling"):
        rope = t["rope_scaling"].get("rope_theta")
    hidden, heads = int(t["hidden_size"]), int(t["num_attention_heads"])
    return QwenDims(n_mels=int(a.get("num_mel_bins", 128)), enc_layers=int(a["encoder_layers"]), enc_d=int(a["d_model"]),
                    enc_heads=int(a["encoder_attention_heads"]), enc_ffn=int(a["encoder_ffn_dim"]),
                    conv_ch=int(a.get("downsample_hidden_size", 480)), out_dim=int(a["output_dim"]),
                    chunks_per_window=int(a.get("n_window_infer", 400)) // CHUNK,
                    max_source_positions=int(a.get("max_source_positions", 1500)), vocab=int(t["vocab_size"]), hidden=hidden,
                    inter=int(t["intermediate_size"]), dec_layers=int(t["num_hidden_layers"]), heads=heads,
                    kv_heads=int(t.get("num_key_value_heads", heads)), head_dim=int(t.get("head_dim", hidden // heads)),
                    rope_theta=float(rope if rope is not None else 10000.0), rms_eps=float(t.get("rms_norm_eps", 1e-6)))


def load_hf_qwen3_asr(folder):
    """(dims, HF-named state dict, tied?) from a Qwen3-ASR checkpoint folder (what the exporter loads, :1470-1480); sharded
    checkpoints are followed.  `tied` = the checkpoint stores no separate lm_head (tie_word_embeddings)."""
    folder = Path(folder)
    dims = qwen_dims_from_hf_config(json.loads((folder / "config.json").read_text()))
    index = folder / "model.safetensors.index.json"
    files = sorted(set(json.loads(index.read_text())["weight_map"].values())) if index.exists() else ["model.safetensors"]
    state: Dict[str, np.ndarray] = {}
    for fn in files:
        state.update(read_safetensors(folder / fn))
    tied = "thinker.lm_head.weight" not in state
    return dims, state, tied


def qwen_prompt_from_tokenizer(tokenizer, languages=()):
    """The ids the exporter bakes around the audio and the per-language tails (Export_Qwen_ASR.py:1500-1586):
    returns (QwenPrompt, {language name: prompt_token_ids})."""
    from .qwen import QwenPrompt
    vocab = tokenizer.get_vocab()
    enc = lambda text: [int(i) for i in tokenizer.encode(text, add_special_tokens=False)]
    im_start, im_end = int(vocab["<|im_start|>"]), int(vocab["<|im_end|>"])
    nl = enc("\n")[0]
    head = [im_start, enc("system")[0], nl]
    suffix = [im_end, nl, im_start, enc("user")[0], nl, int(vocab["<|audio_start|>"])]
    tail = [int(vocab["<|audio_end|>"]), im_end, nl, im_start, enc("assistant")[0], nl] + enc("language ")
    stop = [int(vocab["<|endoftext|>"]), im_end]
    asr_text = [int(vocab["<asr_text>"])]
    tails = {name: enc(name) + asr_text for name in languages}
    return QwenPrompt(tuple(head), tuple(suffix), tuple(tail), tuple(stop)), tails


# ---------------------------------------------------------------------------------------------
def read_kaldi_cmvn(path):
    """`am.mvn` of a FunASR model folder (Kaldi nnet text: <AddShift> row = negated means, <Rescale> row = inverse standard
    deviations) -> (means, vars) as the front end applies them, (x + means) * vars -- what `frontend.cmvn` holds in the
    exporter (SenseVoice/Export_SenseVoice.py:363-364)."""
    lines = Path(path).read_text().splitlines()
    means = scales = None
    for i, line in enumerate(lines):
        tag = line.split()[0] if line.split() else ""
        if tag in ("<AddShift>", "<Rescale>") and i + 1 < len(lines):
            items = lines[i + 1].split()
            lo, hi = items.index("[") + 1, items.index("]")
            vals = np.asarray([float(v) for v in items[lo:hi]], dtype=np.float32)
            if tag == "<AddShift>":
                means = vals
            else:
                scales = vals
    if means is None or scales is None:
        raise ValueError(f"{path}: no <AddShift> / <Rescale> rows")
    return means, scales


def load_funasr_sensevoice(folder):
    """(dims, checkpoint dict under the names `sensevoice.fold_sensevoice` expects) from a FunASR SenseVoiceSmall folder
    (`model.pt` + `am.mvn`, what `AutoModel(model=...)` loads at Export_SenseVoice.py:356-364).  State-dict keys follow the
    module tree the exporter walks (:208-266): encoder.encoders0 / encoders / tp_encoders [.self_attn.{linear_q_k_v,
    linear_out, fsmn_block}, .feed_forward.{w_1, w_2}, .norm1, .norm2], encoder.after_norm, encoder.tp_norm, ctc.ctc_lo, embed."""
    import torch
    from .sensevoice import SenseVoiceDims
    folder = Path(folder)
    sd = torch.load(folder / "model.pt", map_location="cpu", weights_only=True)
    sd = sd.get("state_dict", sd)
    t = lambda k: sd[k].detach().float()
    groups = []
    for grp in ("encoders0", "encoders", "tp_encoders"):
        n = 0
        while f"encoder.{grp}.{n}.norm1.weight" in sd:
            groups.append(f"encoder.{grp}.{n}.")
            n += 1
    n0 = sum(1 for g in groups if ".encoders0." in g)
    ntp = sum(1 for g in groups if ".tp_encoders." in g)
    d_model = int(t(groups[0] + "self_attn.linear_out.weight").shape[0])
    means, scales = read_kaldi_cmvn(folder / "am.mvn")
    feat = int(means.shape[0])
    ffn = int(t(groups[0] + "feed_forward.w_1.weight").shape[0])
    k = int(t(groups[0] + "self_attn.fsmn_block.weight").shape[-1])
    vocab = int(t("ctc.ctc_lo.weight").shape[0])
    n_embed = int(t("embed.weight").shape[0])
    base = SenseVoiceDims()
    heads = base.n_heads
    if (folder / "config.yaml").exists():            # FunASR model config: encoder_conf.attention_heads (4 for SenseVoiceSmall)
        import yaml
        conf = yaml.safe_load((folder / "config.yaml").read_text()) or {}
        heads = int((conf.get("encoder_conf") or {}).get("attention_heads", heads))
    if feat % base.lfr_m:
        raise ValueError("am.mvn width is not a multiple of the LFR stack")
    dims = SenseVoiceDims(n_mels=feat // base.lfr_m, d_model=d_model, n_heads=heads, ffn=ffn, n_blocks0=n0,
                          n_blocks=len(groups) - n0 - ntp, n_tp_blocks=ntp, vocab=vocab, fsmn_kernel=k, n_embed=n_embed)
    raw = {"embed": t("embed.weight"), "cmvn_means": torch.from_numpy(means), "cmvn_vars": torch.from_numpy(scales)}
    for i, g in enumerate(groups):
        p = f"blk{i}."
        raw[p + "norm1.g"], raw[p + "norm1.b"] = t(g + "norm1.weight"), t(g + "norm1.bias")
        raw[p + "norm2.g"], raw[p + "norm2.b"] = t(g + "norm2.weight"), t(g + "norm2.bias")
        raw[p + "qkv.w"], raw[p + "qkv.b"] = t(g + "self_attn.linear_q_k_v.weight"), t(g + "self_attn.linear_q_k_v.bias")
        raw[p + "out.w"], raw[p + "out.b"] = t(g + "self_attn.linear_out.weight"), t(g + "self_attn.linear_out.bias")
        raw[p + "fsmn.w"] = t(g + "self_attn.fsmn_block.weight").reshape(d_model, k)
        raw[p + "w1.w"], raw[p + "w1.b"] = t(g + "feed_forward.w_1.weight"), t(g + "feed_forward.w_1.bias")
        raw[p + "w2.w"], raw[p + "w2.b"] = t(g + "feed_forward.w_2.weight"), t(g + "feed_forward.w_2.bias")
    for n, key in (("after_norm", "encoder.after_norm"), ("tp_norm", "encoder.tp_norm")):
        raw[n + ".g"], raw[n + ".b"] = t(key + ".weight"), t(key + ".bias")
    raw["ctc.w"], raw["ctc.b"] = t("ctc.ctc_lo.weight"), t("ctc.ctc_lo.bias")
    return dims, raw


def load_funasr_paraformer(folder):
    """(dims, checkpoint dict under the names `paraformer.fold_paraformer` expects) from a FunASR Paraformer-large folder
    (`model.pt` + `am.mvn` [+ `config.yaml`]).  Keys follow the module tree `PARAFORMER.__init__` walks
    (Paraformer/Non-Streaming/Export_Paraformer.py:385-465): encoder.encoders0 / encoders, encoder.after_norm,
    predictor.{cif_conv1d, cif_output}, decoder.decoders [.feed_forward.{w_1, w_2, norm}, .norm1-3, .self_attn.fsmn_block,
    .src_attn.{linear_q, linear_k_v, linear_out}], decoder.decoders3, decoder.after_norm, decoder.output_layer."""
    import torch
    from .paraformer import ParaformerDims
    folder = Path(folder)
    sd = torch.load(folder / "model.pt", map_location="cpu", weights_only=True)
    sd = sd.get("state_dict", sd)
    t = lambda k: sd[k].detach().float()

    def count(prefix, probe):
        n = 0
        while f"{prefix}.{n}.{probe}" in sd:
            n += 1
        return n

    n0, n1 = count("encoder.encoders0", "norm1.weight"), count("encoder.encoders", "norm1.weight")
    na, nf = count("decoder.decoders", "norm1.weight"), count("decoder.decoders3", "norm1.weight")
    enc = [f"encoder.encoders0.{i}." for i in range(n0)] + [f"encoder.encoders.{i}." for i in range(n1)]
    dec = [f"decoder.decoders.{i}." for i in range(na)] + [f"decoder.decoders3.{i}." for i in range(nf)]
    means, scales = read_kaldi_cmvn(folder / "am.mvn")
    base = ParaformerDims()
    heads, tail = base.n_heads, base.tail_threshold
    if (folder / "config.yaml").exists():
        import yaml
        conf = yaml.safe_load((folder / "config.yaml").read_text()) or {}
        heads = int((conf.get("encoder_conf") or {}).get("attention_heads", heads))
        tail = float((conf.get("predictor_conf") or {}).get("tail_threshold", tail))
    D = int(t(enc[0] + "self_attn.linear_out.weight").shape[0])
    feat = int(means.shape[0])
    dims = ParaformerDims(n_mels=feat // base.lfr_m, d_model=D, n_heads=heads, ffn=int(t(enc[0] + "feed_forward.w_1.weight").shape[0]),
                          n_blocks0=n0, n_blocks=n1, dec_att_blocks=na, dec_ffn_blocks=nf,
                          dec_ffn=int(t(dec[0] + "feed_forward.w_1.weight").shape[0]), vocab=int(t("decoder.output_layer.weight").shape[0]),
                          fsmn_kernel=int(t(enc[0] + "self_attn.fsmn_block.weight").shape[-1]),
                          cif_kernel=int(t("predictor.cif_conv1d.weight").shape[-1]), tail_threshold=tail)
    raw = {"cmvn_means": torch.from_numpy(means), "cmvn_vars": torch.from_numpy(scales)}

    def norm(dst, src):
        raw[dst + ".g"], raw[dst + ".b"] = t(src + ".weight"), t(src + ".bias")

    for i, g in enumerate(enc):
        p = f"enc{i}."
        norm(p + "norm1", g + "norm1"); norm(p + "norm2", g + "norm2")
        raw[p + "qkv.w"], raw[p + "qkv.b"] = t(g + "self_attn.linear_q_k_v.weight"), t(g + "self_attn.linear_q_k_v.bias")
        raw[p + "out.w"], raw[p + "out.b"] = t(g + "self_attn.linear_out.weight"), t(g + "self_attn.linear_out.bias")
        raw[p + "fsmn.w"] = t(g + "self_attn.fsmn_block.weight").reshape(D, -1)
        raw[p + "w1.w"], raw[p + "w1.b"] = t(g + "feed_forward.w_1.weight"), t(g + "feed_forward.w_1.bias")
        raw[p + "w2.w"], raw[p + "w2.b"] = t(g + "feed_forward.w_2.weight"), t(g + "feed_forward.w_2.bias")
    norm("enc_after_norm", "encoder.after_norm")
    raw["cif.conv.w"], raw["cif.conv.b"] = t("predictor.cif_conv1d.weight"), t("predictor.cif_conv1d.bias")
    raw["cif.out.w"], raw["cif.out.b"] = t("predictor.cif_output.weight"), t("predictor.cif_output.bias")
    for i, g in enumerate(dec):
        p = f"dec{i}."
        norm(p + "norm1", g + "norm1"); norm(p + "ffn_norm", g + "feed_forward.norm")
        raw[p + "w1.w"], raw[p + "w1.b"] = t(g + "feed_forward.w_1.weight"), t(g + "feed_forward.w_1.bias")
        raw[p + "w2.w"] = t(g + "feed_forward.w_2.weight")
        if i < na:
            norm(p + "norm2", g + "norm2"); norm(p + "norm3", g + "norm3")
            raw[p + "fsmn.w"] = t(g + "self_attn.fsmn_block.weight").reshape(D, -1)
            raw[p + "q.w"], raw[p + "q.b"] = t(g + "src_attn.linear_q.weight"), t(g + "src_attn.linear_q.bias")
            raw[p + "kv.w"], raw[p + "kv.b"] = t(g + "src_attn.linear_k_v.weight"), t(g + "src_attn.linear_k_v.bias")
            raw[p + "cout.w"], raw[p + "cout.b"] = t(g + "src_attn.linear_out.weight"), t(g + "src_attn.linear_out.bias")
    norm("dec_after_norm", "decoder.after_norm")
    raw["out.w"], raw["out.b"] = t("decoder.output_layer.weight"), t("decoder.output_layer.bias")
    return dims, raw


def read_vocab(path):
    """Vocabulary of the Paraformer driver: one token per line (`Vocab_Paraformer.txt`, Inference_Paraformer_ONNX.py:181-183),
    or FunASR's `tokens.json` list."""
    path = Path(path)
    if path.suffix == ".json":
        return [str(x) for x in json.loads(path.read_text(encoding="utf-8"))]
    return [line.rstrip("\n") for line in path.read_text(encoding="utf-8").splitlines()]

